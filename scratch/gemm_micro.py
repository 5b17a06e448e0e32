import sys, torch
sys.path.insert(0, '.')
from recbole_fairrec_b200 import _lib, ops
dev = torch.device('cuda')
def run(M, K, N, act=2):
    X = torch.randn(M, K, device=dev, requires_grad=True); W = torch.randn(N, K, device=dev, requires_grad=True)
    b = torch.randn(N, device=dev, requires_grad=True)
    for _ in range(3):
        Y = ops.LinearAct.apply(X, W, b, act, 0.0, 0); Y.sum().backward()
    _lib.profile_enable(True)
    for _ in range(20):
        Y = ops.LinearAct.apply(X, W, b, act, 0.0, 0); Y.backward(torch.ones_like(Y))
    prof = _lib.profile_report(); _lib.profile_enable(False)
    fl = 2.0 * M * K * N
    print(f"M={M} K={K} N={N}: " + "  ".join(f"{k}={1e3*v[1]/v[0]:.1f}us({fl/(v[1]/v[0]*1e-3)/1e12:.2f}TF)" for k, v in prof.items()))
for shp in [(9748,64,128),(9748,128,64),(9748,64,64),(2048,64,128),(2048,128,256),(2048,256,128),(2048,128,128),(2048,64,32),(2048,32,1)]:
    run(*shp)
