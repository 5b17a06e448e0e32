import sys; sys.path.insert(0, '.')
import numpy as np, torch
from recbole_fairrec_b200 import _lib, kernels
n, ni, d, K = [int(x) for x in sys.argv[1:5]]
mode = _lib.SCORE_TC_3XTF32 if sys.argv[5] == 'tc' else _lib.SCORE_EXACT_FP32
g = torch.Generator(device='cuda').manual_seed(0)
U = (torch.randn(n + 1, d, device='cuda', generator=g) * 0.3)
I = (torch.randn(ni, d, device='cuda', generator=g) * 0.3)
users = torch.arange(1, n + 1, dtype=torch.int32, device='cuda')
hist_off = torch.arange(0, (n + 1) * 4, 4, dtype=torch.int64, device='cuda')[:n + 1]
hist_items = torch.sort(torch.randint(1, ni, (n, 4), device='cuda', generator=g), dim=1).values.to(torch.int32).reshape(-1).contiguous()
for _ in range(3):
    ids, sc = kernels.fullsort_topk(U, I, users, hist_off, hist_items, K, _lib.TRANSFORM_CLAMP_DIV, 5.0, 0, mode)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
ids, sc = kernels.fullsort_topk(U, I, users, hist_off, hist_items, K, _lib.TRANSFORM_CLAMP_DIV, 5.0, 0, mode)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
print(f'{sys.argv[5]} n={n} ni={ni} d={d}: {ms:.3f} ms  {2.0*n*ni*d/ms/1e9:.1f} TFLOP/s fp32-equivalent')
