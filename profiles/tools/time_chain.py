#!/usr/bin/env python
"""Device time of one forward + backward (and of a no-grad forward) of the fused MLP chain kernels (csrc/mlp_chain.cu) at the
shapes of PFCN_MLP.yaml / FairGo_PMF.yaml, each captured in a CUDA graph and replayed (no host time in the numbers).

    python profiles/tools/time_chain.py
    ncu --set full --clock-control none --import-source on -k regex:k_mlp_chain -c 6 -o gpurun_out/r02_chain \
        python profiles/tools/time_chain.py --only dis
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from recbole_fairrec_b200.layers import MLPLayers
from recbole_fairrec_b200 import ops
dev = torch.device("cuda", 0)
def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
def case(name, layers, bn, drop, M, act="leakyrelu", group=1):
    mods = [MLPLayers(layers, dropout=drop, activation=act, bn=bn, init_method="norm").to(dev).train() for _ in range(group)]
    x = torch.randn(M, layers[0], device=dev, requires_grad=True)
    gys = [torch.randn(M, layers[-1], device=dev) for _ in range(group)]
    def step():
        ys = ops.mlp_chain(mods, [x])
        torch.autograd.backward(ys, gys)
    # graph-captured fwd+bwd to remove host overhead
    ops.init_autograd_thread(dev)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
        with torch.cuda.graph(g, stream=s):
            step()
    torch.cuda.current_stream().wait_stream(s)
    t = timeit(g.replay)
    with torch.no_grad():
        gf = torch.cuda.CUDAGraph()
        with torch.cuda.stream(s):
            ops.mlp_chain(mods, [x])
            with torch.cuda.graph(gf, stream=s):
                ops.mlp_chain(mods, [x])
        torch.cuda.current_stream().wait_stream(s)
        tf = timeit(gf.replay)
    print(f"{name:28s} M={M:6d} group={group} fwd+bwd {t:8.1f} us   fwd(no grad) {tf:8.1f} us", flush=True)
import argparse
ap = argparse.ArgumentParser()
ap.add_argument("--only", default=None)
ONLY = ap.parse_args().only
_case = case
def case(name, *a, **k):
    if ONLY is None or ONLY == name:
        _case(name, *a, **k)
case("one layer 64->64", [64, 64], False, 0.0, 2048)
case("one layer 64->64 bn", [64, 64], True, 0.0, 2048)
case("filter", [64, 128, 64], True, 0.0, 2048)
case("tower", [128, 64, 32, 16, 1], False, 0.2, 2048, act="relu")
case("tower 4096", [128, 64, 32, 16, 1], False, 0.2, 4096, act="relu")
case("dis", [64, 128, 256, 128, 128, 64, 32, 1], True, 0.3, 2048)
case("dis x3", [64, 128, 256, 128, 128, 64, 32, 1], True, 0.3, 2048, group=3)
case("fairgo filter", [64, 128, 64], True, 0.0, 9748)
case("fairgo dis", [64, 16, 8, 4, 1], False, 0.0, 9748)
