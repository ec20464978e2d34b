"""Developer script for profilers: a few device-resident commits of the microbench shape
(2^16 x 128, rate_bits 3, cap_height 4).  Usage: ncu ... python commit_workload.py [reps]"""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
import vfhe_b200 as V
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ctx = V.Context(0)
C, lg = 128, 16
n = 1 << lg; m = n << 3
cols = torch.from_numpy(V.synthetic_columns(C, n).view(np.int64)).cuda()
coeffs = torch.empty((C, n), dtype=torch.int64, device='cuda')
leaves = torch.empty((m, C), dtype=torch.int64, device='cuda')
digests = torch.empty((2 * (m - 16), 4), dtype=torch.int64, device='cuda')
cap = torch.empty((16, 4), dtype=torch.int64, device='cuda')
for it in range(reps):
    V.commit_device(ctx, cols.data_ptr(), C, lg, 3, 4, False, coeffs.data_ptr(), leaves.data_ptr(), digests.data_ptr(), cap.data_ptr())
ctx.sync()
print("cap0", hex(int(cap[0, 0].item()) & (2**64 - 1)))
