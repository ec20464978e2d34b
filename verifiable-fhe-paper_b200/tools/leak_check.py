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
    s.close(); z.close(); b.close()
for _ in range(5):
    step()
torch.cuda.synchronize()
free0 = torch.cuda.mem_get_info()[0]
for _ in range(300):
    step()
torch.cuda.synchronize()
free1 = torch.cuda.mem_get_info()[0]
print("free before %.1f MiB, after %.1f MiB, delta %.1f MiB" % (free0 / 2**20, free1 / 2**20, (free0 - free1) / 2**20))
assert free0 - free1 < 64 * 2**20, "device memory grows"
print("leak check ok")
