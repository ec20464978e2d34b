import sys; sys.path.insert(0, '.')
import numpy as np, time
import vfhe_b200 as V
if len(sys.argv) > 2:  # A/B of interpreter variants: a library built with other -DVPBS_PROG_THREADS / _POINTS
    V._lib.LIB_PATH = sys.argv[2]
ctx = V.Context(0)
rng = np.random.default_rng(1)
log_n = 16; n = 1 << log_n
wires = V.synthetic_columns(135, n, 1); cs = V.synthetic_columns(85, n, 2)
k_is = V.get_unique_coset_shifts(n, 80)
wb = V.commit_resident(wires, 3, False, 4, ctx=ctx); cb = V.commit_resident(cs, 3, False, 4, ctx=ctx)
sg = V.Sigmas(cs[5:], k_is, ctx)
b, g, a = (rng.integers(1, V.P, size=2, dtype=np.uint64) for _ in range(3))
zb = V.commit_zs_partial_products(wb, sg, b, g, 8, 3, 4)
B = V.GateProgramBuilder(); R = lambda i: (0, i)
for i in range(12): B.into(i, B.ADD, B.wire(i), B.imm(0))
j = 0
while len(B.code) < int(sys.argv[1]):
    for i in range(12):
        if j < 123:
            B.emit(j, B.into(24, B.SUB, R(i), B.wire(12 + j % 120))); j += 1
        B.into(24, B.MUL, R(i), R(i)); B.into(25, B.MUL, R(24), R(24)); B.into(25, B.MUL, R(25), R(24)); B.into(12 + i, B.MUL, R(25), R(i))
    for r in range(12):
        B.into(26 + r, B.MUL, R(12), B.imm(int(rng.integers(1, 64))))
        for i in range(1, 12):
            B.mad(R(26 + r), R(12 + i), B.imm(int(rng.integers(1, 64))))
    for r in range(12): B.into(r, B.ADD, R(26 + r), B.imm(int(rng.integers(1, 2**62))))
B.end_gate(R(0))
prog = B.build(ctx, 123)
for rep in range(3):
    t0 = time.perf_counter()
    qb = V.commit_quotient_polys(cb, 5, wb, zb, k_is, 8, 3, b, g, a, 3, 4, program=prog)
    print("quotient with %d-instruction program: %.3f ms" % (len(B.code), (time.perf_counter() - t0) * 1e3))
    qb.close()
