import sys, time, ctypes
sys.path.insert(0, '.')
import numpy as np, torch
import vfhe_b200 as V
ctx = V.Context(0)
for (C, lg) in ((128, 16), (135, 16), (20, 16), (16, 16), (85, 16), (135, 13)):
    n = 1 << lg; m = n << 3
    cols = torch.from_numpy(V.synthetic_columns(C, n).view(np.int64)).cuda()
    coeffs = torch.empty((C, n), dtype=torch.int64, device='cuda')
    leaves = torch.empty((m, C), dtype=torch.int64, device='cuda')
    digests = torch.empty((2 * (m - 16), 4), dtype=torch.int64, device='cuda')
    cap = torch.empty((16, 4), dtype=torch.int64, device='cuda')
    torch.cuda.synchronize()
    for it in range(4):
        st = V.commit_device(ctx, cols.data_ptr(), C, lg, 3, 4, False, coeffs.data_ptr(), leaves.data_ptr(), digests.data_ptr(), cap.data_ptr(), want_stats=True)
    print(C, lg, {k: round(v, 3) for k, v in st.items()}, flush=True)
