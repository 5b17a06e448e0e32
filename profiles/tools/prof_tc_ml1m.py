#!/usr/bin/env python
"""The tensor-core full-sort scorer at the ML-1M shape (6,006 users x 3,707 items, d = 64, K = 10) for ncu:
    ncu --set full --import-source on -k regex:k_fullsort_tc --launch-skip 2 -c 1 -o gpurun_out/x python profiles/tools/prof_tc_ml1m.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

import bench_ml1m as bm
import recbole_fairrec_b200 as pkg
from recbole_fairrec_b200 import synth

dev = torch.device("cuda", 0)
w, train, valid, test, gender = bm.make_workload("ml1m")
cfg = pkg.Config(embedding_size=w["d"], fair_objective="value", topk=[10], valid_metric="NDCG@10", device=dev, score_mode="tc")
users, hist, pos = synth.eval_lists(train, valid, test, "valid")
edata = pkg.EvalData(users, hist, pos, {"gender": gender.astype(np.int64)}, dev)
counts = np.bincount(train[1], minlength=w["n_items"])
ev = pkg.FullSortEvaluator(cfg, w["n_items"], {int(i): int(c) for i, c in enumerate(counts) if c > 0})
g = torch.Generator(device=dev).manual_seed(1)
U = torch.randn(w["n_users"], w["d"], device=dev, generator=g) * 0.3
I = torch.randn(w["n_items"], w["d"], device=dev, generator=g) * 0.3
for _ in range(4):
    ev.collect(U, I, edata, 5.0)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    ev.collect(U, I, edata, 5.0)
b.record()
torch.cuda.synchronize()
print(f"eager pass: {a.elapsed_time(b) / 10:.3f} ms")
