import sys, faulthandler; faulthandler.enable()
sys.path.insert(0, '.')
import numpy as np, torch
import recbole_fairrec_b200 as pkg
from recbole_fairrec_b200 import synth
dev = torch.device('cuda')
G = int(sys.argv[1]); order = sys.argv[2]; nu, ni, nint, d, batch = 6041, 3707, int(sys.argv[3]), 64, 2048
uid, iid, rating, gender = synth.interactions(nu, ni, nint, 2020)
cfg = pkg.Config(embedding_size=d, fair_objective='value', train_batch_size=batch, device=dev)
train = pkg.TrainData(uid, iid, rating, gender, nu, ni, dev)
loader = pkg.FOCFDataLoader(cfg, train, mode='fast', seed=1)
print('max_batch', loader.max_batch, 'len', len(loader), flush=True)
model = pkg.FOCF(cfg, synth.SynthDataset(nu, ni, 5.0)).to(dev)
model.init_adam(lr=1e-3, weight_decay=1e-3)
losses = torch.zeros(len(loader) * 2 + 16, device=dev)
runner = model.planned_runner(loader, losses, graph_steps=G)
torch.cuda.synchronize(); print('captured', flush=True)
if order == 'big_first':
    runner.run(G); torch.cuda.synchronize(); print('big ok', flush=True)
for i in range(50):
    runner.run(1)
torch.cuda.synchronize(); print('ones ok', flush=True)
runner.run(G); torch.cuda.synchronize(); print('big after ones ok', flush=True)
for i in range(300):
    runner.run(G)
torch.cuda.synchronize(); print('many big ok', losses[:5].tolist(), flush=True)
model.check_flags()
