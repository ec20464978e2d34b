"""Developer script for compute-sanitizer --tool racecheck / synccheck: the TMA NTT passes (strided,
leaf-order and natural-order final passes; several tiles per CTA) and the openings / permutation
kernels that synchronise through shared memory, on small shapes."""
import sys
sys.path.insert(0, '.')
import numpy as np
import vfhe_b200 as V
ctx = V.Context(0)
rng = np.random.default_rng(2)
cols = rng.integers(0, 2**64, size=(20, 1 << 16), dtype=np.uint64)
b = V.commit_resident(cols, 1, False, 2, False, ctx=ctx)          # IFFT + 2 LDE blocks through the r16t passes
z = V.commit_resident(cols[:16], 1, False, 2, True, ctx=ctx)
fri = V.FriCommitPhase.from_openings([b, z], [[(0, j) for j in range(20)] + [(1, j) for j in range(16)], [(1, 0)]],
                                     rng.integers(0, 2**63, size=(2, 2), dtype=np.uint64),
                                     rng.integers(0, 2**63, size=2, dtype=np.uint64), 1)
fri.commit_layer(4, 2); fri.fold(np.array([3, 4], np.uint64)); fri.final_poly(); fri.close()
sig = V.Sigmas(rng.integers(0, 2**64, size=(8, 1 << 12), dtype=np.uint64), V.get_unique_coset_shifts(1 << 12, 8), ctx)
V.all_wires_permutation_partial_products(rng.integers(0, 2**64, size=(8, 1 << 12), dtype=np.uint64), sig,
                                         np.array([5, 6], np.uint64), np.array([7, 8], np.uint64), 8)
sig.close(); b.close(); z.close()
print("racecheck workload done")
