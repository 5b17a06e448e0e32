"""The fused whole-chain kernels (csrc/mlp_chain.cu: fr_mlp_chain_forward / fr_mlp_chain_backward, tcgen05 3xTF32 for the
forward, data-gradient and weight-gradient GEMMs) against
  * the float64 evaluation of the SAME torch module tree (MLPLayers keeps the reference's nn.Sequential of
    Dropout / Linear / BatchNorm1d / activation, recbole/model/layers.py:30-85), dropout off: outputs, input gradient and every
    parameter gradient within 1e-5 of the tensor's scale, BatchNorm running statistics within 1e-6;
  * the per-layer kernels of this package with dropout ON (both paths hash the same (seed, element) counters, so the masks
    are identical): same bound.
Shapes: the filter, discriminator and scorer chains of PFCN_MLP.yaml / FairGo_PMF.yaml, ragged row counts, 1/7/21-wide heads,
three discriminators sharing one launch."""
import copy

import numpy as np
import pytest
import torch

from abs_terms import track_abs_terms

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _pkg():
    import recbole_fairrec_b200 as pkg
    return pkg


def _close(a, b, what, tol=RTOL):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    scale = max(float(b.abs().max()), 1e-30)
    err = float((a - b).abs().max()) / scale
    assert err <= tol, f"{what}: {err:.3e} of scale {scale:.3e}"


def _close_terms(mine, ref, terms, what, tol=RTOL):
    """a batch-summed gradient against its float64 value, relative to the sum of the magnitudes of its terms
    (tests/abs_terms.py): the rounding bound of a dot product, which stays meaningful when the terms cancel"""
    mine, ref, terms = mine.detach().double().cpu(), ref.detach().double().cpu(), terms.detach().double().cpu()
    err = float(((mine - ref).abs() / (terms + 1e-30)).max())
    assert err <= tol, f"{what}: {err:.3e} of the terms' magnitude"


def _make(layers, dropout, act, bn, seed, dev, calibrate_on=None):
    from recbole_fairrec_b200.layers import MLPLayers
    torch.manual_seed(seed)
    m = MLPLayers(layers, dropout=dropout, activation=act, bn=bn, init_method="norm").to(dev)
    with torch.no_grad():       # weights of a trained-looking scale (init 0.01 makes everything tiny and BN-dominated)
        for p in m.parameters():
            if p.dim() == 2:
                p.copy_(torch.randn_like(p) / np.sqrt(p.shape[1]))
            else:
                p.copy_(torch.randn_like(p) * 0.1 + (1.0 if p.abs().mean() > 0.5 else 0.0))
    if calibrate_on is not None and act in ("relu", "leakyrelu"):
        _keep_off_the_kink(m, calibrate_on)
    return m


@torch.no_grad()
def _keep_off_the_kink(m, x):
    """relu / leakyrelu are not differentiable at 0: a pre-activation within float32 rounding of 0 may fall on different sides
    in the float32 kernels and in the float64 yardstick, and through BatchNorm's batch sums one such element moves the
    gradients of a whole column by ~1/sqrt(M) (measured: 1-2 flipped signs among 6e5 outputs -> 6e-2 on the input gradient).
    The additive term in front of every activation (BatchNorm beta, else the Linear bias) is therefore set so that column n
    of the pre-activation has mean +-5 standard deviations (signs alternate: both slopes are exercised) on the test input."""
    import torch.nn as nn
    h = x.double()
    mods = list(m.mlp_layers)
    k = 0
    while k < len(mods):
        lin = mods[k + 1]
        k += 2
        v = h @ lin.weight.double().t() + lin.bias.double()
        shift = lin.bias
        if k < len(mods) and isinstance(mods[k], nn.BatchNorm1d):
            bnm = mods[k]
            k += 1
            v = (v - v.mean(0)) / torch.sqrt(v.var(0, unbiased=False) + bnm.eps) * bnm.weight.double() + bnm.bias.double()
            shift = bnm.bias
        if k < len(mods) and not isinstance(mods[k], nn.Dropout):
            actm = mods[k]
            k += 1
            sign = torch.where(torch.arange(v.shape[1], device=v.device) % 2 == 0, 1.0, -1.0).double()
            delta = sign * 5.0 * v.std(0, unbiased=False).clamp_min(1e-3) - v.mean(0)
            shift.add_(delta.float())
            v = actm(v + delta.float().double())
        h = v


CASES = [
    ("filter", [64, 128, 64], 0.0, "leakyrelu", True, 2048),
    ("filter_ragged", [64, 128, 64], 0.0, "leakyrelu", True, 300),
    ("dis_bin", [64, 128, 256, 128, 128, 64, 32, 1], 0.0, "leakyrelu", True, 2048),
    ("dis_7", [64, 128, 256, 128, 128, 64, 32, 7], 0.0, "leakyrelu", True, 1000),
    ("tower", [128, 64, 32, 16, 1], 0.0, "relu", False, 2048),
    ("fairgo_dis", [64, 16, 8, 4, 1], 0.0, "leakyrelu", False, 9748),
    ("fairgo_filter", [64, 128, 64], 0.0, "leakyrelu", True, 9748),
    ("tanh_sigmoid", [32, 48, 21], 0.0, "tanh", True, 130),
    ("sigmoid", [16, 100, 20], 0.0, "sigmoid", False, 129),
]


@pytest.mark.parametrize("name,layers,dropout,act,bn,M", CASES, ids=[c[0] for c in CASES])
def test_chain_matches_float64_module(name, layers, dropout, act, bn, M):
    from recbole_fairrec_b200 import ops
    dev = torch.device("cuda", 0)
    torch.manual_seed(5)
    x = torch.randn(M, layers[0], device=dev)
    gy = torch.randn(M, layers[-1], device=dev) / M
    m = _make(layers, dropout, act, bn, 11, dev, calibrate_on=x).train()
    ref = copy.deepcopy(m).double()
    x1 = x.clone().requires_grad_(True)
    ops.set_chain_enabled(True)
    launches0 = _pkg()._lib.launch_count()
    y = m(x1)
    y.backward(gy)
    assert _pkg()._lib.launch_count() - launches0 == 2, "one forward and one backward launch"
    x2 = x.double().requires_grad_(True)
    with track_abs_terms() as tr:
        yr = ref.mlp_layers(x2)
        yr.backward(gy.double())
    _close(y, yr, "output")
    _close(x1.grad, x2.grad, "input gradient")
    for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        _close_terms(p.grad, q.grad, tr.abs[id(q)], f"grad {n}")
    for (n, b), (_, c) in zip(m.named_buffers(), ref.named_buffers()):
        if "num_batches" in n:
            assert int(b) == int(c) == 1
        else:
            _close(b, c, n, 1e-5)


def test_eval_mode_and_no_grad_training_mode():
    from recbole_fairrec_b200 import ops
    dev = torch.device("cuda", 0)
    m = _make([64, 128, 256, 128, 128, 64, 32, 1], 0.3, "tanh", True, 3, dev)
    with torch.no_grad():
        for b in m.buffers():
            if b.dtype == torch.float32:
                b.copy_(torch.rand_like(b) + 0.5)
    ref = copy.deepcopy(m).double()
    x = torch.randn(70000, 64, device=dev)          # > one eval chunk
    m.eval(), ref.eval()
    for q in m.parameters():
        q.requires_grad_(False)
    _close(m(x[:100]), ref.mlp_layers(x[:100].double()), "eval output (grad mode on, frozen parameters)")
    with torch.no_grad():
        _close(m(x), ref.mlp_layers(x.double()), "eval output")
    # training-mode forward without autograd (the discriminator phase evaluates the filters this way): batch statistics,
    # running statistics advance; dropout off in this comparison
    m2 = _make([64, 128, 64], 0.0, "leakyrelu", True, 4, dev, calibrate_on=x[:2048]).train()
    r2 = copy.deepcopy(m2).double()
    with torch.no_grad():
        _close(m2(x[:2048]), r2.mlp_layers(x[:2048].double()), "no-grad training output")
    for (n, b), (_, c) in zip(m2.named_buffers(), r2.named_buffers()):
        if "num_batches" not in n:
            _close(b, c, n, 1e-5)


def test_chain_with_dropout_matches_the_per_layer_kernels():
    from recbole_fairrec_b200 import ops
    dev = torch.device("cuda", 0)
    for layers, p, act, bn, M in (([128, 64, 32, 16, 1], 0.2, "relu", False, 2048),
                                  ([64, 128, 256, 128, 128, 64, 32, 1], 0.3, "leakyrelu", True, 2048)):
        x = torch.randn(M, layers[0], device=dev)
        m = _make(layers, p, act, bn, 21, dev, calibrate_on=x).train()
        m_b = copy.deepcopy(m)
        gy = torch.randn(M, layers[-1], device=dev) / M
        outs = []
        for mod, fused in ((m, True), (m_b, False)):
            ops.set_chain_enabled(fused)
            ops._seed_counter = __import__("itertools").count(1000)
            xx = x.clone().requires_grad_(True)
            y = mod(xx)
            y.backward(gy)
            outs.append((y, xx.grad, [q.grad for q in mod.parameters()], [b for b in mod.buffers()]))
        ops.set_chain_enabled(True)
        assert float((outs[0][0] == 0).float().mean()) < 0.9
        _close(outs[0][0], outs[1][0], "output")
        _close(outs[0][1], outs[1][1], "input gradient")
        scale = max(float(b.abs().max()) for b in outs[1][2])
        for a, b in zip(outs[0][2], outs[1][2]):      # both float32: bound relative to the largest gradient of the chain
            assert float((a - b).abs().max()) <= 2e-5 * scale
        for a, b in zip(outs[0][3], outs[1][3]):
            _close(a.float(), b.float(), "buffer", 1e-5)


def test_three_discriminators_share_one_launch():
    from recbole_fairrec_b200 import ops
    dev = torch.device("cuda", 0)
    heads = (1, 7, 21)
    M = 2048
    x = torch.randn(M, 64, device=dev)
    mods = [_make([64, 128, 256, 128, 128, 64, 32, h], 0.0, "leakyrelu", True, 30 + h, dev, calibrate_on=x).train()
            for h in heads]
    refs = [copy.deepcopy(m).double() for m in mods]
    gys = [torch.randn(M, h, device=dev) / M for h in heads]
    x1 = x.clone().requires_grad_(True)
    ops.set_chain_enabled(True)
    n0 = _pkg()._lib.launch_count()
    ys = ops.mlp_chain(mods, [x1])
    assert ys is not None
    torch.autograd.backward(ys, gys)
    assert _pkg()._lib.launch_count() - n0 == 2
    x2 = x.double().requires_grad_(True)
    with track_abs_terms() as tr:
        yrs = [r.mlp_layers(x2) for r in refs]
        torch.autograd.backward(yrs, [g.double() for g in gys])
    for y, yr in zip(ys, yrs):
        _close(y, yr, "output")
    _close(x1.grad, x2.grad, "summed input gradient")
    for m, r in zip(mods, refs):
        for (n, p), (_, q) in zip(m.named_parameters(), r.named_parameters()):
            _close_terms(p.grad, q.grad, tr.abs[id(q)], f"grad {n}")


def test_run_to_run_bit_stability():
    from recbole_fairrec_b200 import ops
    dev = torch.device("cuda", 0)
    x = torch.randn(2048, 64, device=dev)
    m = _make([64, 128, 256, 128, 128, 64, 32, 1], 0.0, "leakyrelu", True, 8, dev).train()
    gy = torch.randn(2048, 1, device=dev)
    res = []
    for _ in range(2):
        xx = x.clone().requires_grad_(True)
        m.zero_grad()
        y = m(xx)
        y.backward(gy)
        res.append([y.clone(), xx.grad.clone()] + [p.grad.clone() for p in m.parameters()])
    for a, b in zip(*res):
        assert torch.equal(a, b)


@pytest.mark.parametrize("world", [2, 4])
def test_data_parallel_chain_equals_the_single_gpu_chain(world):
    """fr_mlp_chain_*_dp with the ranks emulated on one device (ops.chain_dp_emulate: the host sequences the launch segments
    instead of the cross-GPU flag barriers; same kernels, same exchange layout): every rank holds M / world rows, the
    BatchNorm column sums cross the ranks through the exchange buffers.  With M / world a multiple of 128 the global
    row-block order equals the single-GPU one, so outputs, input gradients and running statistics are BIT-identical to the
    single-GPU chain on the whole batch (dropout included: the masks hash the global row index); the parameter gradients
    are the sum of the ranks' shares (another association of the same terms: 1e-6)."""
    import itertools
    from recbole_fairrec_b200 import ops
    dev = torch.device("cuda", 0)
    M = 2048
    for layers, p, heads in (([64, 128, 64], 0.0, None), ([64, 128, 256, 128, 128, 64, 32], 0.3, (1, 7))):
        torch.manual_seed(3)
        x = torch.randn(M, layers[0], device=dev)
        outs = heads or (layers[-1],)
        base = layers if heads is None else layers
        mods = [_make((base + [h]) if heads else base, p, "leakyrelu", True, 40 + h, dev).train() for h in outs]
        gys = [torch.randn(M, (h if heads else layers[-1]), device=dev) / M for h in outs]
        replicas = [[copy.deepcopy(m) for m in mods] for _ in range(world)]
        # single GPU
        ops._seed_counter = itertools.count(5000)
        x1 = x.clone().requires_grad_(True)
        ys = ops.mlp_chain(mods, [x1])
        torch.autograd.backward(ys, gys)
        # data parallel
        ops._seed_counter = itertools.count(5000)
        Ml = M // world
        res = ops.chain_dp_emulate(replicas, [[x[r * Ml:(r + 1) * Ml].contiguous()] for r in range(world)],
                                   [[g[r * Ml:(r + 1) * Ml].contiguous() for g in gys] for r in range(world)])
        for c in range(len(mods)):
            assert torch.equal(torch.cat([res[r][0][c] for r in range(world)]), ys[c]), "outputs"
        assert torch.equal(torch.cat([res[r][1][0] for r in range(world)]), x1.grad), "input gradient"
        params = [q for m in mods for q in ops._chain_params(ops._pairs_of(m))]
        for i, q in enumerate(params):
            total = sum(res[r][2][i] for r in range(world))
            scale = max(float(q.grad.abs().max()), 1e-12)
            assert float((total - q.grad).abs().max()) <= 2e-6 * scale + 1e-9, i
        for m, reps in zip(mods, zip(*replicas)):
            for rep in reps:
                for (n, b), (_, b2) in zip(m.named_buffers(), rep.named_buffers()):
                    assert torch.equal(b, b2), n
