import sys, os
sys.path.insert(0, '.')
import numpy as np, torch
import vfhe_b200 as V
ctx = V.Context(0)
C, lg = 128, 16
n = 1 << lg; m = n << 3
cols = torch.from_numpy(V.synthetic_columns(C, n).view(np.int64)).cuda()
coeffs = torch.empty((C, n), dtype=torch.int64, device='cuda')
leaves = torch.empty((m, C), dtype=torch.int64, device='cuda')
digests = torch.empty((2 * (m - 16), 4), dtype=torch.int64, device='cuda')
cap = torch.empty((16, 4), dtype=torch.int64, device='cuda')
torch.cuda.synchronize()
best = 1e9
for it in range(8):
    st = V.commit_device(ctx, cols.data_ptr(), C, lg, 3, 4, False, coeffs.data_ptr(), leaves.data_ptr(), digests.data_ptr(), cap.data_ptr(), want_stats=True)
    best = min(best, st['merkle_ms'] - st['leaf_hash_ms'])
print(os.environ.get('VPBS_COOP_LOG'), 'tree levels ms', round(best, 4), 'cap0', hex(int(cap[0,0].item()) & (2**64-1)), flush=True)
