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
