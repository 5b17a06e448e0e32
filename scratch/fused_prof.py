import sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
import recbole_fairrec_b200 as pkg
from recbole_fairrec_b200 import synth
dev = torch.device('cuda', 0)
w, train, valid, test, gender = bench.make_workload('ml1m')
cfg = pkg.Config(embedding_size=64, fair_objective="value", fair_weight=1.0, train_batch_size=2048, device=dev)
tdata = pkg.TrainData(train[0], train[1], train[2], gender, w["n_users"], w["n_items"], dev)
loader = pkg.FOCFDataLoader(cfg, tdata, mode="fast", seed=1)
model = pkg.FOCF(cfg, synth.SynthDataset(w["n_users"], w["n_items"], 5.0)).to(dev)
model.init_adam(lr=1e-3, weight_decay=1e-3)
it = iter(loader)
for _ in range(12):
    model.train_step(next(it))
torch.cuda.synchronize()
