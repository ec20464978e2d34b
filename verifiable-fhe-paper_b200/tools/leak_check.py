"""Developer script: device memory stays flat over many resident commits / FRI chains / sigma uploads."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
import vfhe_b200 as V
ctx = V.Context(0)
rng = np.random.default_rng(5)
cols = rng.integers(0, 2**64, size=(70, 1 << 13), dtype=np.uint64)
sig = rng.integers(0, 2**64, size=(40, 1 << 13), dtype=np.uint64)
k = V.get_unique_coset_shifts(1 << 13, 40)
def step():
    b = V.commit_resident(cols, 3, False, 4, False, ctx=ctx)
    s = V.Sigmas(sig, k, ctx)
    z = V.commit_zs_partial_products(b, s, [3, 4], [5, 6], 8, 3, 4)
    f = V.FriCommitPhase.from_openings([b, z], [[(0, j) for j in range(70)] + [(1, j) for j in range(10)], [(1, 0), (1, 1)]],
                                       np.array([[1, 2], [3, 4]], np.uint64), np.array([5, 6], np.uint64), 3)
    f.commit_layer(4, 4); f.fold([7, 8]); f.commit_layer(4, 4); f.fold([9, 10]); f.final_poly(); f.query(0, [1, 2, 3]); f.close()
    b.merkle_tree.get_many(np.arange(5, dtype=np.uint64)); b.eval_ext2(np.array([[1, 2]], np.uint64))
    V.PolynomialBatch.from_values(cols, 3, False, 4, ctx=ctx)
    # second half of round 2: sigma-carrying constants batch, device quotient from a freshly uploaded gate
    # program, all-oracle openings, staged (pageable) upload with chunk-wise sponge hashing
    cb = V.commit_resident(sig, 3, False, 4, ctx=ctx)
    B = V.GateProgramBuilder()
    B.emit(0, B.mad(B.mul(B.wire(0), B.wire(1)), B.wire(2), B.imm(7)))
    B.end_gate(B.selector_filter(0, 1, range(3), False))
    prog = B.build(ctx)
    q = V.commit_quotient_polys(cb, 0, b, z, k, 8, 3, [3, 4], [5, 6], [7, 8], 3, 4, program=prog)
    V.open_all_at_points([b, z, q], np.array([[1, 2], [3, 4]], np.uint64))
    V.open_all_at_leaves([b, z, q], np.arange(7, dtype=np.uint64))
    prog.close(); q.close(); cb.close()
    s.close(); z.close(); b.close()
import psutil
for _ in range(5):
    step()
torch.cuda.synchronize()
rss0 = psutil.Process().memory_info().rss
free0 = torch.cuda.mem_get_info()[0]
for _ in range(300):
    step()
torch.cuda.synchronize()
free1 = torch.cuda.mem_get_info()[0]
print("free before %.1f MiB, after %.1f MiB, delta %.1f MiB" % (free0 / 2**20, free1 / 2**20, (free0 - free1) / 2**20))
rss1 = psutil.Process().memory_info().rss
print("host RSS before %.1f MiB, after %.1f MiB, threads %d" % (rss0 / 2**20, rss1 / 2**20, psutil.Process().num_threads()))
assert free0 - free1 < 64 * 2**20, "device memory grows"
assert rss1 - rss0 < 256 * 2**20, "host memory grows"
print("leak check ok")
