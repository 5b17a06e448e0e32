import sys, faulthandler; faulthandler.enable()
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from recbole_fairrec_b200 import _lib, kernels
import test_fullsort_eval_gpu as T
n_users, n_items, d, n_eval, K = [int(x) for x in sys.argv[1:6]]
U, I, users, ho, hi, po, pi, sst = T.random_eval_case(1, n_users, n_items, d, n_eval)
data, ev, Ud, Id = T.build(U, I, users, ho, hi, po, pi, sst, [K])
args = (Ud, Id, data.users, data.hist_off, data.hist_items)
ids_e, sc_e = kernels.fullsort_topk(*args, K, _lib.TRANSFORM_NONE, 1.0)
torch.cuda.synchronize(); print('exact done', flush=True)
ids_t, sc_t = kernels.fullsort_topk(*args, K, _lib.TRANSFORM_NONE, 1.0, 0, _lib.SCORE_TC_3XTF32)
torch.cuda.synchronize(); print('tc done', flush=True)
se, st = sc_e.cpu().numpy(), sc_t.cpu().numpy()
print('max abs score diff', np.abs(se - st).max(), 'ids equal frac', (ids_e == ids_t).float().mean().item())
print('exact row0', se[0][:6], ids_e[0][:6].tolist()); print('tc    row0', st[0][:6], ids_t[0][:6].tolist())
print('exact rowL', se[-1][:6], ids_e[-1][:6].tolist()); print('tc    rowL', st[-1][:6], ids_t[-1][:6].tolist())
