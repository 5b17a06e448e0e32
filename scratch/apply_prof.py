"""scale-out shaped FOCF steps (2M users x 262k items, d=128, 2^18-row batches) for ncu captures of the HBM-bound
kernels; synthetic interactions are drawn directly (no uniqueness pass) to keep the setup short."""
import sys; sys.path.insert(0, '.')
import numpy as np, torch
import recbole_fairrec_b200 as pkg
from recbole_fairrec_b200 import synth
dev = torch.device('cuda')
nu, ni, d, batch, n_inter = 2_000_001, 262_145, 128, 1 << 18, 8_000_000
rng = np.random.default_rng(0)
iid = (rng.lognormal(4.5, 1.4, n_inter) % (ni - 1)).astype(np.int32) + 1
uid = rng.integers(1, nu, n_inter).astype(np.int32)
rating = rng.integers(1, 6, n_inter).astype(np.float32)
gender = (rng.random(nu) < 0.28).astype(np.float32) + 1
cfg = pkg.Config(embedding_size=d, fair_objective='value', train_batch_size=batch, device=dev)
train = pkg.TrainData(uid, iid, rating, gender, nu, ni, dev)
loader = pkg.FOCFDataLoader(cfg, train, mode='fast', seed=1)
model = pkg.FOCF(cfg, synth.SynthDataset(nu, ni, 5.0)).to(dev)
model.init_adam(lr=1e-3, weight_decay=1e-3)
n = 0
for inter in loader:
    model.train_step(inter)
    n += 1
    if n >= int(sys.argv[1]):
        break
torch.cuda.synchronize()
model.check_flags()
print('steps', n, 'rows', len(inter['user_id']))
