#!/usr/bin/env python
"""us/step of the persistent FOCF epoch kernel at the ML-1M shape (one launch of N steps between two CUDA events).
FR_FOCF_EPOCH_SKIP=<mask> leaves phases out (timing experiments only; results are then meaningless):
1 forward, 2 statistics, 4 gradients, 8 Adam of touched rows, 16 Adam of untouched rows, 32 row-stamp prefetch, 64 barrier 2."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import bench_ml1m as bm
import recbole_fairrec_b200 as pkg
from recbole_fairrec_b200 import synth

dev = torch.device("cuda", 0)
w, train, valid, test, gender = bm.make_workload("ml1m")
cfg = pkg.Config(embedding_size=w["d"], fair_objective="value", fair_weight=1.0, topk=[10], valid_metric="NDCG@10",
                 train_batch_size=w["batch"], learning_rate=1e-3, weight_decay=1e-3, device=dev, seed=2020)
tdata = pkg.TrainData(train[0], train[1], train[2], gender, w["n_users"], w["n_items"], dev)
loader = pkg.FOCFDataLoader(cfg, tdata, mode="fast", seed=2020)
model = pkg.FOCF(cfg, synth.SynthDataset(w["n_users"], w["n_items"], 5.0)).to(dev)
model.init_adam(lr=1e-3, weight_decay=1e-3)
losses = torch.zeros(len(loader) * 2 + 16, device=dev)
runner = model.epoch_runner(loader, losses)
runner.run(16)
torch.cuda.synchronize()
best = None
for rep in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    runner.run(256)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e3 / 256
    best = t if best is None else min(best, t)
print(f"skip={os.environ.get('FR_FOCF_EPOCH_SKIP', '0'):>4}  {best:.2f} us/step")
