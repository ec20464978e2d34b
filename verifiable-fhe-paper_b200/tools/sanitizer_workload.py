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
