import sys; sys.path.insert(0, '.')
import numpy as np
import vfhe_b200 as V
ctx = V.Context(0)
rng = np.random.default_rng(2)
for lg in (5, 10, 14):
    n = 1 << lg
    wires = rng.integers(0, 2**64, size=(12, n), dtype=np.uint64)
    cs = rng.integers(0, 2**64, size=(11, n), dtype=np.uint64)
    k_is = V.get_unique_coset_shifts(n, 8)
    b, g, a = (rng.integers(1, V.P, size=2, dtype=np.uint64) for _ in range(3))
    wb, cb = V.commit_resident(wires, 3, False, 4, ctx=ctx), V.commit_resident(cs, 3, False, 4, ctx=ctx)
    sg = V.Sigmas(cs[3:], k_is, ctx)
    zb = V.commit_zs_partial_products(wb, sg, b, g, 4, 3, 4)
    B = V.GateProgramBuilder()
    acc = B.mul(B.wire(0), B.const(1))
    for j in range(6):
        acc = B.mad(acc, B.wire(1 + j), B.imm(3 + j))
    B.emit(0, acc); B.emit(2, B.sub(acc, B.pih(1)))
    B.end_gate(B.selector_filter(0, 1, range(3), False))
    prog = B.build(ctx)
    qb = V.commit_quotient_polys(cb, 3, wb, zb, k_is, 4, 3, b, g, a, 3, 4, program=prog, public_inputs_hash=b.repeat(2))
    V.open_all_at_points([wb, cb, zb, qb], rng.integers(0, 2**64, size=(2, 2), dtype=np.uint64))
    V.open_all_at_leaves([wb, cb, zb, qb], rng.integers(0, n << 3, size=5, dtype=np.uint64))
    for x in (wb, cb, zb, qb): x.close()
    sg.close(); prog.close()
print("racecheck workload (second half) done")
