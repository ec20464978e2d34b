#!/bin/bash
# round-2 ncu captures of the kernels added in the second half: quotient terms, gate-program interpreter
# (final form), chunk-wise sponge hashing, openings
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T=verifiable-fhe-paper_b200/tools
ncu --set full --import-source on --clock-control none -k regex:"quotient_permutation_terms|gate_program_eval" -c 2 -f -o gpurun_out/r2_quotient_kernels python $T/gate_program_workload.py 2000 > gpurun_out/n_ncu1.log 2>&1; tail -1 gpurun_out/n_ncu1.log
cat > /tmp/wl.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, ctypes
import vfhe_b200 as V
ctx = V.Context(0)
n = 1 << 16
cols = [V.synthetic_columns(1, n, 100 + c)[0].copy() for c in range(128)]
u64p = V._lib.u64p
colp = (u64p * 128)(*[a.ctypes.data_as(u64p) for a in cols])
cap = np.empty((16, 4), np.uint64)
for _ in range(2):
    h = ctypes.c_void_p()
    ctx.check(ctx.lib.vpbs_batch_commit(ctx.handle, colp, 128, 16, 3, 4, 0, None, cap.ctypes.data_as(u64p), ctypes.byref(h), None))
    out = np.empty((2, 128, 2), np.uint64)
    pts = np.array([[3, 5], [7, 11]], np.uint64)
    ctx.check(ctx.lib.vpbs_batch_eval_ext2(h, pts.ctypes.data_as(u64p), 2, out.ctypes.data_as(u64p)))
    ctx.lib.vpbs_batch_destroy(h)
print("done")
PY
ncu --set full --import-source on --clock-control none -k regex:"hash_leaves_part|eval_ext2" -s 5 -c 5 -f -o gpurun_out/r2_sponge_openings_kernels python /tmp/wl.py > gpurun_out/n_ncu2.log 2>&1; tail -1 gpurun_out/n_ncu2.log
