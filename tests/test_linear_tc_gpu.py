"""fr_linear_forward on the tensor cores (linear_tc.cu: tcgen05.mma kind::tf32, TMA, TMEM, 3xTF32) against a float64
reference and against the CUDA-core kernel of the same entry point (allow_tensor_cores = 0), incl. ragged M, every
activation, missing bias and the dropout mask shared with the backward kernels."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def run(lib, X, W, b, act, drop_p=0.0, seed=0, tc=True):
    from recbole_fairrec_b200._lib import check, ptr, stream_ptr
    M, K = X.shape
    N = W.shape[0]
    Y = torch.empty((M, N), dtype=torch.float32, device=X.device)
    on_tc = lib.fr_linear_uses_tensor_cores(M, K, N) if tc else 0
    check(lib.fr_linear_forward(ptr(X), ptr(W), ptr(b), ptr(Y), M, K, N, act, float(drop_p), seed, None, 0, 1 if tc else 0,
                                stream_ptr()), "fr_linear_forward")
    return Y, on_tc


ACTS = {0: lambda x: x, 1: torch.relu, 2: lambda x: torch.nn.functional.leaky_relu(x, 0.01), 3: torch.sigmoid, 4: torch.tanh}


@pytest.mark.parametrize("M,K,N", [(2048, 64, 128), (2048, 128, 256), (2048, 256, 128), (9748, 64, 64), (300, 32, 32),
                                   (1, 128, 64), (129, 96, 160), (513, 512, 96)])
def test_tc_linear_matches_float64(M, K, N):
    from recbole_fairrec_b200._lib import load
    lib = load()
    g = torch.Generator(device="cuda").manual_seed(M + K + N)
    X = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    for act in (0, 2, 3):
        Y, nws = run(lib, X, W, b, act)
        assert nws > 0, "shape should be tensor-core eligible"
        ref = ACTS[act](X.double() @ W.double().t() + b.double())
        err = (Y.double() - ref).abs().max().item() / ref.abs().max().item()
        assert err < 5e-6, (act, err)     # 3xTF32 drops the lo.lo products (2^-22 each) and accumulates in fp32
        Yc, _ = run(lib, X, W, b, act, tc=False)       # CUDA-core kernel: same result up to fp32 summation order
        assert (Y - Yc).abs().max().item() / ref.abs().max().item() < 5e-6
    Y, _ = run(lib, X, W, None, 1)
    ref = torch.relu(X.double() @ W.double().t())
    assert (Y.double() - ref).abs().max().item() / ref.abs().max().item() < 5e-6
    Y2, _ = run(lib, X, W, None, 1)
    assert torch.equal(Y, Y2)                          # run-to-run bit stability


def test_tc_linear_shapes_outside_the_tile_rules_fall_back():
    from recbole_fairrec_b200._lib import load
    lib = load()
    assert lib.fr_linear_uses_tensor_cores(2048, 64, 16) == 0
    assert lib.fr_linear_uses_tensor_cores(2048, 48, 64) == 0
    assert lib.fr_linear_uses_tensor_cores(2048, 512, 64) == 1
    X = torch.randn(100, 48, device="cuda")
    W = torch.randn(7, 48, device="cuda")
    Y, nws = run(lib, X, W, None, 0)
    assert nws == 0 and torch.allclose(Y, X @ W.t(), atol=1e-4)


def test_tc_linear_dropout_mask_is_the_backward_kernels_mask():
    """the hi/lo split applies the same counter-based mask as the CUDA-core loader: TC forward == CUDA-core forward for
    one seed, and autograd through ops.LinearAct stays consistent (finite-difference check of one weight)"""
    from recbole_fairrec_b200 import ops
    from recbole_fairrec_b200._lib import load
    lib = load()
    g = torch.Generator(device="cuda").manual_seed(3)
    X = torch.randn(512, 64, device="cuda", generator=g)
    W = torch.randn(128, 64, device="cuda", generator=g) / 8
    b = torch.randn(128, device="cuda", generator=g)
    Y, _ = run(lib, X, W, b, 2, drop_p=0.3, seed=1234)
    Yc, _ = run(lib, X, W, b, 2, drop_p=0.3, seed=1234, tc=False)
    assert (Y - Yc).abs().max().item() < 1e-5 * Yc.abs().max().item()
    Y0, _ = run(lib, X, W, b, 2)
    assert (Y - Y0).abs().max().item() > 1e-2           # the mask did something
    Wp = W.clone().requires_grad_(True)
    out = ops.LinearAct.apply(X, Wp, b, 0, 0.3, 99)     # identity activation: the loss is linear in W, so the
    out.sum().backward()                                 # finite difference is exact up to rounding
    eps = 0.5
    W2 = W.clone()
    W2[5, 7] += eps
    fd = (ops.LinearAct.apply(X, W2, b, 0, 0.3, 99).double().sum() - out.double().sum()).item() / eps
    assert abs(fd - Wp.grad[5, 7].item()) < 1e-3 * max(1.0, abs(fd))
