import sys, torch
sys.path.insert(0, '.')
from recbole_fairrec_b200 import ops
dev = torch.device('cuda')
for M, K, N in ((2048, 128, 256), (9748, 64, 128)):
    X = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev)
    for _ in range(4):
        Y = ops.LinearAct.apply(X, W, b, 2, 0.0, 0)
torch.cuda.synchronize()
