#!/usr/bin/env python
"""Phase trace of the cooperative FOCF fused step at the ML-1M shape (fr_focf_step_trace; FR_FOCF_TRACE=1): every CTA
stamps %globaltimer at its 8 phase boundaries; printed per phase as (first CTA to get there, last CTA) in microseconds
from the earliest start stamp of the launch, plus the per-phase duration of the slowest / median CTA.

    FR_FOCF_TRACE=1 python profiles/tools/trace_fused.py
"""
import ctypes
import os
import sys

os.environ.setdefault("FR_FOCF_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

import recbole_fairrec_b200 as pkg
from recbole_fairrec_b200 import _lib, synth

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench_ml1m as bm

dev = torch.device("cuda", 0)
w, train, valid, test, gender = bm.make_workload("ml1m")
cfg = pkg.Config(embedding_size=w["d"], fair_objective="value", fair_weight=1.0, topk=[10], valid_metric="NDCG@10",
                 train_batch_size=w["batch"], learning_rate=1e-3, weight_decay=1e-3, device=dev, seed=2020)
tdata = pkg.TrainData(train[0], train[1], train[2], gender, w["n_users"], w["n_items"], dev)
loader = pkg.FOCFDataLoader(cfg, tdata, mode="fast", seed=2020)
model = pkg.FOCF(cfg, synth.SynthDataset(w["n_users"], w["n_items"], 5.0)).to(dev)
model.init_adam(lr=1e-3, weight_decay=1e-3)
losses = torch.zeros(len(loader) * 2 + 16, device=dev)
runner = model.planned_runner(loader, losses, graph_steps=8, persistent=False)
runner.run(16)
torch.cuda.synchronize()
lib = _lib.load()
names = ["forward", "barrier1", "stats", "grads", "barrier2", "adam", "barrier3"]
acc = []
for rep in range(6):
    runner.eager_steps(1) if rep % 2 == 0 else runner.run(1)
    torch.cuda.synchronize()
    n_cta = 144
    buf = (ctypes.c_uint64 * (8 * n_cta))()
    _lib.check(lib.fr_focf_step_trace(buf, 8 * n_cta), "fr_focf_step_trace")
    t = np.array(list(buf), dtype=np.int64).reshape(n_cta, 8)
    t0 = t[:, 0].min()
    rel = (t - t0) / 1e3
    dur = np.diff(rel, axis=1)
    print(f"rep {rep} ({'eager' if rep % 2 == 0 else 'graph'}): step {rel[:, 7].max():.1f} us; boundaries (min..max over CTAs): "
          + " ".join(f"{rel[:, k].min():.1f}..{rel[:, k].max():.1f}" for k in range(8)))
    print("    per-phase duration median / max over CTAs: "
          + " ".join(f"{n}={np.median(dur[:, k]):.1f}/{dur[:, k].max():.1f}" for k, n in enumerate(names)))

# ---- the persistent epoch kernel (fr_focf_epoch_run): timing of one launch of many steps + phase trace of one step
runner = model.epoch_runner(loader, losses)
if runner is None:
    print("epoch kernel: shape not eligible")
    sys.exit(0)
for steps in (64, 256):
    runner.run(8)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    runner.run(steps)
    e1.record()
    torch.cuda.synchronize()
    print(f"epoch kernel: {steps} steps in one launch: {e0.elapsed_time(e1) * 1e3 / steps:.2f} us/step")
n_cta, n_prod = 148, 2 * runner.state["n_slots"]
buf = (ctypes.c_uint64 * (16 * n_cta))()
_lib.check(lib.fr_focf_epoch_trace(buf, 16 * n_cta), "fr_focf_epoch_trace")
t = np.array(list(buf), dtype=np.int64).reshape(n_cta, 16)
c = t[n_prod:, :8]
t0 = c[:, 0].min()
rel = (c - t0) / 1e3
dur = np.diff(rel, axis=1)
print(f"epoch kernel, traced step: {rel[:, 7].max():.1f} us; boundaries (min..max over compute CTAs): "
      + " ".join(f"{rel[:, k].min():.1f}..{rel[:, k].max():.1f}" for k in range(8)))
print("    per-phase duration median / max over CTAs: "
      + " ".join(f"{n}={np.median(dur[:, k]):.1f}/{dur[:, k].max():.1f}" for k, n in enumerate(names)))
for k in range(n_prod):
    p = (t[k, :4] - t[k, 0]) / 1e3
    print(f"    producer CTA {k} ({'user' if k & 1 else 'item'} side of slot {k >> 1}): gathered {p[1]:.1f}, sorted {p[2]:.1f}, "
          f"published {p[3]:.1f} us after its start")
