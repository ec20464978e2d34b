"""Developer script: where the end-to-end time of the resident host path goes (wall clock per call)."""
import sys, time, ctypes
sys.path.insert(0, '.')
import numpy as np
import vfhe_b200 as V
ctx = V.Context(0)
lib, u64p = ctx.lib, V._lib.u64p
C, lg, r, h = 128, 16, 3, 4
n, m = 1 << lg, (1 << lg) << r
p = lib.vpbs_host_alloc(C * n * 8)
cols = np.ctypeslib.as_array((ctypes.c_uint64 * (C * n)).from_address(p)).reshape(C, n)
cols[:] = V.synthetic_columns(C, n)
colp = (u64p * C)(*[cols[c].ctypes.data_as(u64p) for c in range(C)])
cap = np.empty((16, 4), np.uint64); zeta = np.array([[3, 4], [5, 6]], np.uint64)
op = np.empty((2, C, 2), np.uint64); idx = np.arange(28, dtype=np.uint64) * 12345 % m
rows = np.empty((28, C), np.uint64); sib = np.empty((28, lg + r - h, 4), np.uint64)
st = V.VpbsStats()
acc = {}
def t(name, f):
    t0 = time.perf_counter(); f(); acc[name] = acc.get(name, 0) + time.perf_counter() - t0
for it in range(13):
    if it == 3: acc.clear()
    hb = ctypes.c_void_p()
    t("commit", lambda: ctx.check(lib.vpbs_batch_commit(ctx.handle, colp, C, lg, r, h, 0, None, cap.ctypes.data_as(u64p), ctypes.byref(hb), ctypes.byref(st))))
    t("eval", lambda: ctx.check(lib.vpbs_batch_eval_ext2(hb, zeta.ctypes.data_as(u64p), 2, op.ctypes.data_as(u64p))))
    t("rows", lambda: ctx.check(lib.vpbs_batch_get_leaves(hb, idx.ctypes.data_as(u64p), 28, rows.ctypes.data_as(u64p))))
    t("prove", lambda: ctx.check(lib.vpbs_batch_prove(hb, idx.ctypes.data_as(u64p), 28, sib.ctypes.data_as(u64p))))
    t("destroy", lambda: lib.vpbs_batch_destroy(hb))
print({k: round(v * 100, 3) for k, v in acc.items()}, "ms per call; last commit stats:", {k: round(v, 3) for k, v in st.as_dict().items()})
