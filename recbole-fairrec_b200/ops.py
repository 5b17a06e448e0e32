"""torch.autograd.Functions over the generic layer kernels of the C ABI (fr_linear_*, fr_batchnorm_*, fr_gather_rows /
fr_scatter_rows_dense, fr_bpr_loss, fr_sigmoid_bce_loss, fr_softmax_ce_loss): the pieces the PFCN / FairGo filter,
discriminator and scorer MLPs are chained from.  torch supplies memory and the autograd tape only."""
import itertools

import torch

from ._lib import check, load, ptr, stream_ptr

ACT = {None: 0, "none": 0, "relu": 1, "leakyrelu": 2, "sigmoid": 3, "tanh": 4}
_seed_counter = itertools.count(1)


def next_seed():
    return (torch.initial_seed() * 0x9E3779B1 + next(_seed_counter) * 0x85EBCA77) & 0xFFFFFFFFFFFFFFFF


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


class LinearAct(torch.autograd.Function):
    """act(dropout(X) @ W.T + b)  (layers.py:60-68 without BatchNorm)"""

    @staticmethod
    def forward(ctx, X, W, b, act, drop_p, seed):
        lib = load()
        X, W = X.contiguous(), W.contiguous()
        M, K = X.shape
        N = W.shape[0]
        Y = torch.empty((M, N), dtype=torch.float32, device=X.device)
        check(lib.fr_linear_forward(ptr(X), ptr(W), ptr(b), ptr(Y), M, K, N, act, float(drop_p), seed, 0, stream_ptr()),
              "fr_linear_forward")
        ctx.save_for_backward(X, W, Y)
        ctx.cfg = (act, float(drop_p), seed, b is not None)
        return Y

    @staticmethod
    def backward(ctx, dY):
        lib = load()
        X, W, Y = ctx.saved_tensors
        act, drop_p, seed, has_b = ctx.cfg
        M, K = X.shape
        N = W.shape[0]
        dY = dY.contiguous()
        dX = torch.empty_like(X) if ctx.needs_input_grad[0] else None
        dW = torch.empty_like(W)
        db = torch.empty(N, dtype=torch.float32, device=X.device) if has_b else None
        ws = _ws(lib.fr_linear_backward_workspace_bytes(M, K, N), X.device)
        check(lib.fr_linear_backward(ptr(X), ptr(W), ptr(Y), ptr(dY), M, K, N, act, drop_p, seed, 0, ptr(dX), ptr(dW),
                                     ptr(db), ptr(ws), ws.numel(), stream_ptr()), "fr_linear_backward")
        return dX, dW, db, None, None, None


class BatchNormAct(torch.autograd.Function):
    """act(BatchNorm1d(X))  (layers.py:64-68); running statistics are updated in place in training mode"""

    @staticmethod
    def forward(ctx, X, gamma, beta, rmean, rvar, momentum, eps, training, act):
        lib = load()
        X = X.contiguous()
        M, N = X.shape
        Y = torch.empty_like(X)
        sm = torch.empty(N, dtype=torch.float32, device=X.device)
        si = torch.empty(N, dtype=torch.float32, device=X.device)
        check(lib.fr_batchnorm_forward(ptr(X), ptr(gamma), ptr(beta), ptr(rmean), ptr(rvar), M, N, float(momentum),
                                       float(eps), 1 if training else 0, act, ptr(Y), ptr(sm), ptr(si), stream_ptr()),
              "fr_batchnorm_forward")
        ctx.save_for_backward(X, Y, gamma, sm, si)
        ctx.cfg = (act, training)
        return Y

    @staticmethod
    def backward(ctx, dY):
        lib = load()
        X, Y, gamma, sm, si = ctx.saved_tensors
        act, training = ctx.cfg
        if not training:
            raise NotImplementedError("BatchNormAct backward is implemented for training mode (batch statistics)")
        M, N = X.shape
        dX, dg, db = torch.empty_like(X), torch.empty_like(gamma), torch.empty_like(gamma)
        check(lib.fr_batchnorm_backward(ptr(X), ptr(Y), ptr(dY.contiguous()), ptr(gamma), ptr(sm), ptr(si), M, N, act,
                                        ptr(dX), ptr(dg), ptr(db), stream_ptr()), "fr_batchnorm_backward")
        return dX, dg, db, None, None, None, None, None, None


class GatherRows(torch.autograd.Function):
    """nn.Embedding forward / dense backward"""

    @staticmethod
    def forward(ctx, T, idx):
        lib = load()
        M, d = idx.numel(), T.shape[1]
        out = torch.empty((M, d), dtype=torch.float32, device=T.device)
        check(lib.fr_gather_rows(ptr(T), ptr(idx), M, d, ptr(out), d, 0, stream_ptr()), "fr_gather_rows")
        ctx.save_for_backward(idx)
        ctx.shape = T.shape
        return out

    @staticmethod
    def backward(ctx, dX):
        lib = load()
        (idx,) = ctx.saved_tensors
        n_rows, d = ctx.shape
        M = idx.numel()
        dX = dX.contiguous()
        dT = torch.empty((n_rows, d), dtype=torch.float32, device=dX.device)
        ws = _ws(lib.fr_scatter_rows_workspace_bytes(M), dX.device)
        check(lib.fr_scatter_rows_dense(ptr(idx), ptr(dX), d, 0, M, d, n_rows, ptr(dT), ptr(ws), ws.numel(), stream_ptr()),
              "fr_scatter_rows_dense")
        return dT, None


class _ScalarLoss(torch.autograd.Function):
    @staticmethod
    def backward(ctx, g):
        return tuple(None if t is None else t * g for t in ctx.grads)


class BprLoss(_ScalarLoss):
    """recbole/model/loss.py:44-46"""

    @staticmethod
    def forward(ctx, pos, neg):
        lib = load()
        shape_p, shape_n = pos.shape, neg.shape
        pos, neg = pos.contiguous().view(-1), neg.contiguous().view(-1)
        loss = torch.empty(1, dtype=torch.float32, device=pos.device)
        dp, dn = torch.empty_like(pos), torch.empty_like(neg)
        check(lib.fr_bpr_loss(ptr(pos), ptr(neg), pos.numel(), ptr(loss), ptr(dp), ptr(dn), stream_ptr()), "fr_bpr_loss")
        ctx.grads = (dp.view(shape_p), dn.view(shape_n))
        return loss.view(())


class SigmoidBce(_ScalarLoss):
    """nn.BCELoss()(nn.Sigmoid()(z), y)  (pfcn_mlp.py:206-207)"""

    @staticmethod
    def forward(ctx, z, y):
        lib = load()
        shape = z.shape
        zf = z.contiguous().view(-1)
        loss = torch.empty(1, dtype=torch.float32, device=z.device)
        dz = torch.empty_like(zf)
        check(lib.fr_sigmoid_bce_loss(ptr(zf), ptr(y.contiguous().view(-1)), zf.numel(), ptr(loss), ptr(dz), stream_ptr()),
              "fr_sigmoid_bce_loss")
        ctx.grads = (dz.view(shape), None)
        return loss.view(())


class SoftmaxCe(_ScalarLoss):
    """nn.CrossEntropyLoss()(Z, y)  (pfcn_mlp.py:209)"""

    @staticmethod
    def forward(ctx, Z, y):
        lib = load()
        Z = Z.contiguous()
        M, C = Z.shape
        loss = torch.empty(1, dtype=torch.float32, device=Z.device)
        dZ = torch.empty_like(Z)
        check(lib.fr_softmax_ce_loss(ptr(Z), ptr(y.contiguous()), M, C, ptr(loss), ptr(dZ), stream_ptr()),
              "fr_softmax_ce_loss")
        ctx.grads = (dZ, None)
        return loss.view(())
