"""Cancellation-aware tolerance for gradient parity tests (test infrastructure).

A weight gradient is a sum over the batch rows, dW[n,k] = sum_r dZ[r,n] * X[r,k]; when its terms nearly cancel (the BPR
tower sees every user once with +g and once with -g) no float32 evaluation -- the reference's sgemm included -- can hold a
bound relative to the VALUE of the sum, only one relative to the sum of the terms' magnitudes,
A[n,k] = sum_r |dZ[r,n]| * |X[r,k]| (the classical rounding bound of a dot product).  `track_abs_terms()` evaluates A
for every Linear weight / bias and BatchNorm weight / bias that a functional torch oracle touches, by running that
oracle ONCE in float64 with `F.linear` / `F.batch_norm` swapped for autograd Functions that also accumulate |dZ|^T |X|
in their backward.  Nothing here depends on the host's BLAS, thread count or float32 kernels: the yardstick is a float64
quantity of the problem itself."""
import contextlib

import torch
import torch.nn.functional as F


class _Tracker:
    def __init__(self):
        self.abs = {}

    def add(self, t, a):
        if t is None:
            return
        k = id(t)
        self.abs[k] = a if k not in self.abs else self.abs[k] + a

    def named(self, st):
        """{state key: A tensor} for the tensors of `st` that were touched"""
        return {k: self.abs[id(v)].numpy() for k, v in st.items() if id(v) in self.abs}


@contextlib.contextmanager
def track_abs_terms():
    tr = _Tracker()
    lin0, bn0 = F.linear, F.batch_norm

    class LinAbs(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, W, b, refs):
            ctx.save_for_backward(x, W)
            ctx.refs = refs
            return lin0(x, W, b)

        @staticmethod
        def backward(ctx, dz):
            x, W = ctx.saved_tensors
            Wref, bref = ctx.refs
            dz2, x2 = dz.reshape(-1, dz.shape[-1]), x.reshape(-1, x.shape[-1])
            tr.add(Wref, dz2.abs().t() @ x2.abs())
            tr.add(bref, dz2.abs().sum(0))
            return dz @ W, dz2.t() @ x2, (dz2.sum(0) if bref is not None else None), None

    class AffineAbs(torch.autograd.Function):
        @staticmethod
        def forward(ctx, xh, g, b, refs):
            ctx.save_for_backward(xh, g)
            ctx.refs = refs
            return xh * g + b

        @staticmethod
        def backward(ctx, dy):
            xh, g = ctx.saved_tensors
            tr.add(ctx.refs[0], (dy * xh).abs().sum(0))
            tr.add(ctx.refs[1], dy.abs().sum(0))
            return dy * g, (dy * xh).sum(0), dy.sum(0), None

    def linear(x, W, b=None):
        return LinAbs.apply(x, W, b, (W, b))

    def batch_norm(x, rm, rv, weight=None, bias=None, training=False, momentum=0.1, eps=1e-5):
        xh = bn0(x, rm, rv, None, None, training, momentum, eps)
        if weight is None:
            return xh
        return AffineAbs.apply(xh, weight, bias, (weight, bias))

    F.linear, F.batch_norm = linear, batch_norm
    try:
        yield tr
    finally:
        F.linear, F.batch_norm = lin0, bn0


def gradient_tolerances(loss_fn, st64, keys, rtol, delta=2.0 ** -20, trials=3, factor=4.0, seed=0):
    """Absolute tolerance per parameter tensor for comparing a float32 gradient with the float64 one, from float64
    quantities only (so it is the same number on every host):

        tol[k] = max( rtol * max(max|g64[k]|, max A[k]),  factor * max_trials max|g64'[k] - g64[k]| )

    * first term: `rtol` relative to the VALUE where the sum is well conditioned, relative to the sum of the terms'
      magnitudes A (see track_abs_terms) where it cancels;
    * second term: how far the EXACT gradient itself moves when every floating-point input is perturbed by a relative
      `delta` (default 2^-20 = 9.5e-7: the error a float32 / 3xTF32 forward pass through a handful of layers carries).  It
      stays below the first term for smooth, well-conditioned tensors and grows exactly where a float32 evaluation cannot be pinned: an
      activation within `delta` of a ReLU / LeakyReLU kink takes the other slope, and a deep BatchNorm stack amplifies
      the rounding of its batch statistics.
    loss_fn(st) -> 0-dim float64 tensor; st64: {key: float64 / integer tensors}; keys: the parameters to differentiate.
    Returns ({key: float64 gradient as numpy or None}, {key: absolute tolerance})."""
    import numpy as np

    def grads(st, tracker=None):
        st = {k: v.clone() for k, v in st.items()}
        for k in keys:
            st[k].requires_grad_(True)
        loss_fn(st).backward()
        g = {k: (st[k].grad.numpy().copy() if st[k].grad is not None else None) for k in keys}
        return g, (tracker.named(st) if tracker is not None else None)

    with track_abs_terms() as tr:
        g64, A = grads(st64, tr)
    move = {k: 0.0 for k in keys}
    for t in range(trials):
        gen = torch.Generator().manual_seed(seed + t)
        stp = {k: (v * (1 + delta * (2 * torch.rand(v.shape, generator=gen, dtype=torch.float64) - 1))
                   if v.is_floating_point() else v) for k, v in st64.items()}
        gp, _ = grads(stp)
        for k in keys:
            if g64[k] is not None and gp[k] is not None:
                move[k] = max(move[k], float(np.abs(gp[k] - g64[k]).max()))
    tol = {}
    for k in keys:
        if g64[k] is None:
            continue
        scale = float(np.abs(g64[k]).max())
        if k in A:
            scale = max(scale, float(A[k].max()))
        tol[k] = max(rtol * scale, factor * move[k])
    return g64, tol
