import sys, torch
sys.path.insert(0, '.')
from recbole_fairrec_b200 import ops
dev = torch.device('cuda')
M, K, N = 9748, 64, 128
X = torch.randn(M, K, device=dev, requires_grad=True); W = torch.randn(N, K, device=dev, requires_grad=True)
b = torch.randn(N, device=dev, requires_grad=True)
for _ in range(4):
    Y = ops.LinearAct.apply(X, W, b, 2, 0.0, 0); Y.backward(torch.ones_like(Y))
torch.cuda.synchronize()
