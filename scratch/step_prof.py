import sys
sys.path.insert(0, '.')
import numpy as np, torch
import recbole_fairrec_b200 as pkg
from recbole_fairrec_b200 import synth
dev = torch.device('cuda')
nu, ni, nint, d, batch = 6041, 3707, 1000209, 64, 2048
uid, iid, rating, gender = synth.interactions(nu, ni, nint, 2020)
cfg = pkg.Config(embedding_size=d, fair_objective='value', train_batch_size=batch, device=dev)
train = pkg.TrainData(uid, iid, rating, gender, nu, ni, dev)
loader = pkg.FOCFDataLoader(cfg, train, mode='fast', seed=1)
model = pkg.FOCF(cfg, synth.SynthDataset(nu, ni, 5.0)).to(dev)
model.init_adam(lr=1e-3, weight_decay=1e-3)
losses = torch.zeros(len(loader) * 2 + 16, device=dev)
runner = model.planned_runner(loader, losses, graph_steps=2)
st = model._graph_step
for i in range(int(sys.argv[1])):
    model._engine().run_planned(st)
torch.cuda.synchronize()
print('rows', runner.plan['batch_rows'][:12])
