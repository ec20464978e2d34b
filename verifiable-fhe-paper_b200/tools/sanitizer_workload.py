import sys
sys.path.insert(0, '.')
import numpy as np
import vfhe_b200 as V
ctx = V.Context(0)
rng = np.random.default_rng(1)
for (lg, C, r, h, co) in [(8, 135, 3, 4, False), (9, 20, 3, 4, True), (10, 7, 2, 0, False), (3, 5, 1, 2, False), (0, 3, 3, 1, False), (16, 3, 3, 4, False)]:
    cols = rng.integers(0, 2**64, size=(C, 1 << lg), dtype=np.uint64)
    f = V.PolynomialBatch.from_coeffs if co else V.PolynomialBatch.from_values
    b = f(cols, r, False, h, ctx=ctx)
    rb = V.commit_resident(cols, r, False, h, co, ctx=ctx)
    idx = rng.integers(0, (1 << lg) << r, size=5, dtype=np.uint64)
    rb.merkle_tree.get_many(idx); rb.merkle_tree.prove_many(idx); rb.eval_ext2(rng.integers(0, 2**64, size=(2, 2), dtype=np.uint64)); rb.close()
v = rng.integers(0, 2**64, size=1 << 12, dtype=np.uint64)
V.fft(v, ctx); V.ifft(v, ctx); V.coset_fft(v, 7, ctx)
vals = rng.integers(0, 2**64, size=(1 << 10, 2), dtype=np.uint64)
V.fri_layer_commit(vals, 4, 2, ctx); V.fri_fold(vals, 4, (3, 4), 49, ctx)
V.fri_proof_of_work(rng.integers(0, 2**63, size=12, dtype=np.uint64), 3, 8, ctx=ctx)
V.MerkleTree.new(rng.integers(0, 2**64, size=(1 << 12, 33), dtype=np.uint64), 3, ctx)
# multi-context commit (row ranges) and a 2^16-row, 70-column batch: chunked upload, persistent NTT
# passes over several tiles per CTA (cp.async staging buffers reused), two-stream LDE
cols = rng.integers(0, 2**64, size=(70, 1 << 16), dtype=np.uint64)
ctxs = [V.Context(0), V.Context(0)]
V.PolynomialBatch.from_values(cols, 1, False, 2, ctxs=ctxs)
V.PolynomialBatch.from_values(cols, 1, False, 2, ctx=ctx)
print("sanitizer workload done")
# round 2, second half: pageable columns through the pinned staging ring with chunk-wise sponge hashing
# (hash_leaves_part; 70 columns = chunks of 32 + 32 + 6), sharded resident batches, the all-oracle
# openings / query calls and the table-driven openings kernel
cols70 = [rng.integers(0, 2**64, size=1 << 13, dtype=np.uint64) for _ in range(70)]
mat70 = np.stack(cols70)
import ctypes
u64p = V._lib.u64p
for threads in (1, 3):
    ctx.set_host_threads(threads)
    colp = (u64p * 70)(*[a.ctypes.data_as(u64p) for a in cols70])
    cap = np.empty((16, 4), np.uint64)
    h = ctypes.c_void_p()
    ctx.check(ctx.lib.vpbs_batch_commit(ctx.handle, colp, 70, 13, 3, 4, 0, None, cap.ctypes.data_as(u64p), ctypes.byref(h), None))
    ctx.lib.vpbs_batch_destroy(h)
V.PolynomialBatch.from_values(mat70, 3, False, 4, ctx=ctx)
for rank in range(4):
    c = V.Context(0)
    c.set_shard(rank, 4)
    rb = V.commit_resident(mat70[:20], 3, False, 4, ctx=c)
    first, nl = rb.shard
    idx = rng.integers(first, first + nl, size=6, dtype=np.uint64)
    rb2 = V.commit_resident(mat70[20:36], 3, False, 4, True, ctx=c)
    V.open_all_at_leaves([rb, rb2], idx)
    V.open_all_at_points([rb, rb2], rng.integers(0, 2**64, size=(2, 2), dtype=np.uint64))
    rb.get_lde_rows(V.reverse_bits(first, 16), 4, 5)
    rb.download(); rb.close(); rb2.close(); c.close()
print("sanitizer workload (second half) done")
# the device quotient: permutation terms + tail, gate terms from the host and from a program, and the
# two-halves form on sharded batches
n13 = 1 << 10
wires = rng.integers(0, 2**64, size=(20, n13), dtype=np.uint64)
cs = rng.integers(0, 2**64, size=(14, n13), dtype=np.uint64)
k_is = V.get_unique_coset_shifts(n13, 8)
bb, gg, aa = (rng.integers(1, V.P, size=2, dtype=np.uint64) for _ in range(3))
Bd = V.GateProgramBuilder()
acc = Bd.mul(Bd.wire(0), Bd.const(1))
acc = Bd.mad(acc, Bd.wire(19), Bd.imm(12345))
Bd.emit(0, Bd.sub(acc, Bd.pih(2)))
Bd.emit(3, Bd.add(Bd.wire(7), Bd.const(0)))
Bd.end_gate(Bd.selector_filter(0, 1, range(3), True))
for shard in ((0, 1), (1, 2)):
    c = V.Context(0)
    c.set_shard(*shard)
    wb, cb = V.commit_resident(wires, 3, False, 4, ctx=c), V.commit_resident(cs, 3, False, 4, ctx=c)
    sg = V.Sigmas(cs[6:], k_is, c)
    zb = V.commit_zs_partial_products(wb, sg, bb, gg, 4, 3, 4)
    prog = Bd.build(c)
    if shard[1] == 1:
        V.commit_quotient_polys(cb, 6, wb, zb, k_is, 4, 3, bb, gg, aa, 3, 4, program=prog, public_inputs_hash=bb.repeat(2)).close()
        V.commit_quotient_polys(cb, 6, wb, zb, k_is, 4, 2, bb, gg, aa, 3, 4,
                                gate_terms=rng.integers(0, 2**64, size=(2, n13 << 2), dtype=np.uint64)).close()
    else:
        import torch
        sp = V.ShardedProof(0, 1, torch.device("cuda", 0))
        h, cap = sp.quotient_polys(c, cb, 6, wb, zb, k_is, 4, 3, bb, gg, aa, 3, 4, 10, program=prog)
        c.lib.vpbs_batch_destroy(h)
    for x in (wb, cb, zb):
        x.close()
    sg.close(); prog.close(); c.close()
print("sanitizer workload (quotient) done")
