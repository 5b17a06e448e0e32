#!/usr/bin/env python
"""Phase-by-phase %globaltimer trace of CTA 0 of the chain kernels (fr_mlp_chain_trace; FR_CHAIN_TRACE=1 forward, =2 backward):
microseconds between consecutive stamps = [phase work, barrier wait, ...].

    FR_CHAIN_TRACE=2 python profiles/tools/trace_chain.py
"""
import sys, os, ctypes
os.environ.setdefault("FR_CHAIN_TRACE", "2")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from recbole_fairrec_b200.layers import MLPLayers
from recbole_fairrec_b200 import ops, _lib
dev = torch.device("cuda", 0)
lib = _lib.load()
def case(name, layers, bn, drop, M, n, group=1):
    mods = [MLPLayers(layers, dropout=drop, activation="leakyrelu", bn=bn, init_method="norm").to(dev).train() for _ in range(group)]
    x = torch.randn(M, layers[0], device=dev, requires_grad=True)
    gys = [torch.randn(M, layers[-1], device=dev) for _ in range(group)]
    for _ in range(3):
        ys = ops.mlp_chain(mods, [x]); torch.autograd.backward(ys, gys)
    torch.cuda.synchronize()
    buf = (ctypes.c_uint64 * n)()
    fn = lib.fr_mlp_chain_trace
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    fn(buf, n)
    t = list(buf)
    print(name, [round((b - a) / 1e3, 1) for a, b in zip(t[:-1], t[1:])], "total", round((t[-1]-t[0])/1e3,1))
case("one layer bn", [64, 64], True, 0.0, 2048, 9)
case("tower", [128, 64, 32, 16, 1], False, 0.2, 2048, 12)
case("dis", [64, 128, 256, 128, 128, 64, 32, 1], True, 0.3, 2048, 36)
case("dis x3", [64, 128, 256, 128, 128, 64, 32, 1], True, 0.3, 2048, 36, group=3)
