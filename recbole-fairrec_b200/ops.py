"""torch.autograd.Functions over the generic layer kernels of the C ABI (fr_linear_*, fr_batchnorm_*, fr_gather_rows /
fr_scatter_rows_dense, fr_bpr_loss, fr_sigmoid_bce_loss, fr_softmax_ce_loss): the pieces the PFCN / FairGo filter,
discriminator and scorer MLPs are chained from.  torch supplies memory and the autograd tape only."""
import itertools

import torch

from ._lib import check, load, ptr, stream_ptr

ACT = {None: 0, "none": 0, "relu": 1, "leakyrelu": 2, "sigmoid": 3, "tanh": 4}
_seed_counter = itertools.count(1)
_seed_dev = {}


def next_seed():
    return (torch.initial_seed() * 0x9E3779B1 + next(_seed_counter) * 0x85EBCA77) & 0xFFFFFFFFFFFFFFFF


def seed_dev(device):
    """device-resident dropout seed offset (added to every layer's host seed): graphed.GraphedStep bumps it inside the
    captured graph so that replays draw fresh masks"""
    key = (device.type, device.index)
    if key not in _seed_dev:
        _seed_dev[key] = torch.zeros(1, dtype=torch.int64, device=device)
    return _seed_dev[key]


_tickets = {}


def tickets(device):
    """self-resetting ticket words of the fused weight-gradient reduction (one buffer per device: the layer kernels of
    a step all run on one stream)"""
    key = (device.type, device.index)
    if key not in _tickets:
        _tickets[key] = torch.zeros(1024, dtype=torch.int32, device=device)
    return _tickets[key]


def bump_seed(device, inc=0x9E3779B97F4A7C15):
    check(load().fr_bump_u64(ptr(seed_dev(device)), inc & 0xFFFFFFFFFFFFFFFF, stream_ptr()), "fr_bump_u64")


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
        sd = seed_dev(X.device) if drop_p > 0 else None
        check(lib.fr_linear_forward(ptr(X), ptr(W), ptr(b), ptr(Y), M, K, N, act, float(drop_p), seed, ptr(sd), 0, 1,
                                    stream_ptr()), "fr_linear_forward")
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
        sd = seed_dev(X.device) if drop_p > 0 else None
        check(lib.fr_linear_backward(ptr(X), ptr(W), ptr(Y), ptr(dY), M, K, N, act, drop_p, seed, ptr(sd), 0, ptr(dX),
                                     ptr(dW), ptr(db), ptr(tickets(X.device)), ptr(ws), ws.numel(), stream_ptr()),
              "fr_linear_backward")
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
        ws = _ws(lib.fr_batchnorm_workspace_bytes(N), X.device)
        check(lib.fr_batchnorm_forward(ptr(X), ptr(gamma), ptr(beta), ptr(rmean), ptr(rvar), M, N, float(momentum),
                                       float(eps), 1 if training else 0, act, ptr(Y), ptr(sm), ptr(si), ptr(ws),
                                       ws.numel(), stream_ptr()), "fr_batchnorm_forward")
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
        ws = _ws(lib.fr_batchnorm_workspace_bytes(N), X.device)
        check(lib.fr_batchnorm_backward(ptr(X), ptr(Y), ptr(dY.contiguous()), ptr(gamma), ptr(sm), ptr(si), M, N, act,
                                        ptr(dX), ptr(dg), ptr(db), ptr(ws), ws.numel(), stream_ptr()),
              "fr_batchnorm_backward")
        return dX, dg, db, None, None, None, None, None, None


_chain_bars = {}
_chain_on = None
CHAIN_MAX_TRAIN_ROWS = 32768
CHAIN_EVAL_CHUNK = 65536


def chain_enabled():
    """the fused whole-chain kernels (mlp_chain.cu) are the default; FR_MLP_CHAIN=0 selects the per-layer kernels"""
    global _chain_on
    if _chain_on is None:
        import os
        _chain_on = os.environ.get("FR_MLP_CHAIN", "1") != "0"
    return _chain_on


def set_chain_enabled(on):
    global _chain_on
    _chain_on = bool(on)


def chain_bars(device):
    """the two self-resetting grid-barrier words of the chain kernels (one buffer per device: chains run on one stream)"""
    key = (device.type, device.index)
    if key not in _chain_bars:
        _chain_bars[key] = torch.zeros(2, dtype=torch.int32, device=device)
    return _chain_bars[key]


_thread_state = __import__("threading").local()


def _bind_thread():
    """fr_thread_init once per Python thread, outside stream captures (see include/fairrec_b200.h)"""
    if not getattr(_thread_state, "bound", False) and not torch.cuda.is_current_stream_capturing():
        check(load().fr_thread_init(), "fr_thread_init")
        _thread_state.bound = True


class _TouchBackwardThread(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.clone()

    @staticmethod
    def backward(ctx, g):
        _bind_thread()
        return g


def init_autograd_thread(device):
    """run one trivial backward so that the autograd worker thread of `device` holds a CUDA context BEFORE a step is
    captured into a CUDA graph (graphed.GraphedStep captures without an eager warm-up pass)"""
    x = torch.zeros(1, device=device, requires_grad=True)
    _TouchBackwardThread.apply(x).sum().backward()


class SeqChain:
    """a chain that is not an MLPLayers module (e.g. FairGo's aggregation nn.Sequential(Linear, act, Linear, act, Linear)):
    pairs = [(nn.Linear, nn.BatchNorm1d | None, activation code, dropout p)]"""

    def __init__(self, pairs, training):
        self._pairs, self.training = list(pairs), bool(training)

    def chain_pairs(self):
        return self._pairs

    def parameters(self):
        for lin, bn, _, _ in self._pairs:
            yield from lin.parameters()
            if bn is not None:
                yield from bn.parameters()


def _pairs_of(mod):
    """[(Linear, BatchNorm1d | None, activation code, dropout p)] of an MLPLayers module or a SeqChain"""
    import torch.nn as nn
    if hasattr(mod, "chain_pairs"):
        return mod.chain_pairs()
    mods = list(mod.mlp_layers)
    pairs, k = [], 0
    act = ACT[mod.activation]
    while k < len(mods):
        lin = mods[k + 1]
        k += 2
        bn = None
        if k < len(mods) and isinstance(mods[k], nn.BatchNorm1d):
            bn = mods[k]
            k += 1
        if k < len(mods) and not isinstance(mods[k], nn.Dropout):
            k += 1
        pairs.append((lin, bn, act, float(mod.dropout)))
    return pairs


_bn_repeat = 1     # see bn_repeat()


class bn_repeat:
    """Context: the fused chain forwards launched inside apply their BatchNorm running-statistics update `n` times
    (fr_chain_layer.bn_repeat).  For a DETERMINISTIC chain (no dropout) evaluated n times on the same batch in training
    mode -- the reference's PFCN loss runs the filter twice per step (pfcn_mlp.py:177-193) -- one launch with n = 2 leaves
    the same outputs and the same buffers as two launches."""

    def __init__(self, n):
        self.n = int(n)

    def __enter__(self):
        global _bn_repeat
        self.prev, _bn_repeat = _bn_repeat, self.n
        return self

    def __exit__(self, *a):
        global _bn_repeat
        _bn_repeat = self.prev


def _chain_layers(mod, training, seeds=None):
    """ctypes layer array of a chain (pointers filled, gradients not) + its pairs; (None, pairs) = outside the rules"""
    from ._lib import ChainLayer
    pairs = _pairs_of(mod)
    arr = (ChainLayer * 8)()
    if len(pairs) > 8:
        return None, pairs
    for i, (lin, bn, act, p) in enumerate(pairs):
        L = arr[i]
        L.K, L.N, L.act, L.has_bn = lin.in_features, lin.out_features, act, 1 if bn is not None else 0
        L.drop_p = float(p) if training else 0.0
        L.seed = seeds[i] if seeds is not None else 0
        L.W = lin.weight.data_ptr()
        L.b = lin.bias.data_ptr() if lin.bias is not None else None
        if bn is not None:
            if bn.momentum is None or not bn.affine or not bn.track_running_stats:
                return None, pairs
            L.bn_eps, L.bn_momentum = float(bn.eps), float(bn.momentum)
            L.gamma, L.beta = bn.weight.data_ptr(), bn.bias.data_ptr()
            L.running_mean, L.running_var = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
            L.num_batches_tracked = bn.num_batches_tracked.data_ptr()
            L.bn_repeat = _bn_repeat
    return arr, pairs


def _chain_params(pairs):
    out = []
    for lin, bn, _, _ in pairs:
        out.append(lin.weight)
        if lin.bias is not None:
            out.append(lin.bias)
        if bn is not None:
            out += [bn.weight, bn.bias]
    return out


class ChainDP:
    """Data-parallel context of the chain kernels (fr_mlp_chain_*_dp): this process holds rows [rank * M, (rank + 1) * M)
    of every batch; BatchNorm column sums cross the ranks through NVLink peer memory.  `group` = torch.distributed process
    group (one process per GPU): the exchange buffers are CUDA-IPC mapped into the peers once, here."""

    XCHG_BYTES = 32 << 20

    def __init__(self, rank, world, group=None, device=None):
        import ctypes

        import torch.distributed as dist
        self.rank, self.world, self.group = rank, world, group
        self.lib = load()
        own = ctypes.c_void_p()
        check(self.lib.fr_xchg_alloc(self.XCHG_BYTES, ctypes.byref(own)), "fr_xchg_alloc")
        self.own = own.value
        self.peers = [None] * world
        self.peers[rank] = self.own
        self.opened = []
        h = ctypes.create_string_buffer(64)
        check(self.lib.fr_xchg_export(self.own, h), "fr_xchg_export")
        handles = [None] * world
        dist.all_gather_object(handles, bytes(h.raw), group=group)
        for k, raw in enumerate(handles):
            if k == rank:
                continue
            pp = ctypes.c_void_p()
            check(self.lib.fr_xchg_open(ctypes.create_string_buffer(raw, 64), ctypes.byref(pp)), "fr_xchg_open")
            self.peers[k] = pp.value
            self.opened.append(pp.value)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        dist.barrier(group=group)

    def struct(self, segment=-1, parity=0):
        from ._lib import ChainDp
        d = ChainDp()
        d.rank, d.world, d.xchg_bytes, d.barriers, d.segment, d.parity = self.rank, self.world, self.XCHG_BYTES, 1, -1, 0
        for k in range(self.world):
            d.xchg[k] = self.peers[k]
        d.status_flags = self.status.data_ptr()
        return d

    def all_reduce_grads(self, params):
        """sum the ranks' shares of the gradients (one NCCL all-reduce over a flat bucket)"""
        import torch.distributed as dist
        gs = [p.grad for p in params if p.grad is not None]
        if not gs:
            return
        flat = torch.cat([g.reshape(-1) for g in gs])
        dist.all_reduce(flat, group=self.group)
        off = 0
        for g in gs:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()

    def close(self):
        for pp in self.opened:
            self.lib.fr_xchg_close(pp)
        self.opened = []
        if self.own:
            self.lib.fr_xchg_free(self.own)
            self.own = None


_chain_dp = None


def set_chain_dp(ctx):
    """install (or clear, with None) the data-parallel context used by every chain call of this process"""
    global _chain_dp
    _chain_dp = ctx


def _chain_build_forward(mods, xs, training, need_grad, seeds_per_mod):
    """ctypes chain array of a forward call: outputs and workspaces allocated -> (chains, ys, workspaces, metas)"""
    from ._lib import Chain
    lib = load()
    M, dev = xs[0].shape[0], xs[0].device
    chains = (Chain * len(mods))()
    keep, ys, metas = [], [], []
    for c, mod in enumerate(mods):
        seeds = seeds_per_mod[c]
        arr, pairs = _chain_layers(mod, training, seeds)
        C = chains[c]
        C.n_layers = len(pairs)
        for i in range(len(pairs)):
            C.layer[i] = arr[i]
        x = xs[c if len(xs) > 1 else 0]
        y = torch.empty((M, pairs[-1][0].out_features), dtype=torch.float32, device=dev)
        nb = lib.fr_mlp_chain_workspace_bytes(arr, len(pairs), M, 1 if training else 0, 1 if need_grad else 0, 0)
        ws = torch.empty(nb, dtype=torch.uint8, device=dev)
        C.X, C.ldx, C.Y = x.data_ptr(), x.shape[1], y.data_ptr()
        C.fwd_ws, C.fwd_ws_bytes = ws.data_ptr(), nb
        keep.append(ws)
        ys.append(y)
        metas.append((seeds, pairs))
    return chains, ys, keep, metas


def _chain_build_backward(mods, xs, ys, dys, metas, fwd_ws, want_dx):
    """ctypes chain array of a backward call: gradient outputs and scratch allocated
    -> (chains, parameter gradients in _chain_params order, per-chain dX, dX_sum, tensors to keep alive)"""
    from ._lib import Chain
    lib = load()
    M, dev = xs[0].shape[0], xs[0].device
    chains = (Chain * len(mods))()
    keep, grads_p, dxs = [], [], []
    shared = len(xs) == 1 and len(mods) > 1
    for c, mod in enumerate(mods):
        seeds, pairs = metas[c]
        arr, _ = _chain_layers(mod, True, seeds)
        C = chains[c]
        C.n_layers = len(pairs)
        x = xs[c if len(xs) > 1 else 0]
        dy = dys[c]
        dy = torch.zeros_like(ys[c]) if dy is None else dy.contiguous()
        for i, (lin, bn, _, _) in enumerate(pairs):
            dW = torch.empty_like(lin.weight)
            arr[i].dW = dW.data_ptr()
            grads_p.append(dW)
            if lin.bias is not None:
                db = torch.empty_like(lin.bias)
                arr[i].db = db.data_ptr()
                grads_p.append(db)
            if bn is not None:
                dg, dbt = torch.empty_like(bn.weight), torch.empty_like(bn.bias)
                arr[i].dgamma, arr[i].dbeta = dg.data_ptr(), dbt.data_ptr()
                grads_p += [dg, dbt]
            C.layer[i] = arr[i]
        nb = lib.fr_mlp_chain_workspace_bytes(arr, len(pairs), M, 1, 1, 1)
        ws = torch.empty(nb, dtype=torch.uint8, device=dev)
        dx = torch.empty((M, pairs[0][0].in_features), dtype=torch.float32, device=dev) if want_dx[c] else None
        C.X, C.ldx, C.Y, C.dY = x.data_ptr(), x.shape[1], ys[c].data_ptr(), dy.data_ptr()
        C.dX = dx.data_ptr() if dx is not None else None
        C.fwd_ws, C.fwd_ws_bytes = fwd_ws[c].data_ptr(), fwd_ws[c].numel()
        C.bwd_ws, C.bwd_ws_bytes = ws.data_ptr(), nb
        keep += [ws, dy]
        dxs.append(dx)
    dx_sum = torch.empty_like(dxs[0]) if (shared and want_dx[0]) else None
    return chains, grads_p, dxs, dx_sum, keep


class MLPChainGroup(torch.autograd.Function):
    """Whole MLPLayers chains (layers.py:30-85) -- up to four over the same batch rows -- as ONE forward launch and ONE
    backward launch of the tcgen05 chain kernels (fr_mlp_chain_forward / fr_mlp_chain_backward); with a data-parallel
    context installed (set_chain_dp) the launches are cut where BatchNorm statistics cross the ranks."""

    @staticmethod
    def forward(ctx, spec, *tensors):
        import ctypes
        lib = load()
        mods, n_x, training = spec["mods"], spec["n_x"], spec["training"]
        xs = [t.contiguous() for t in tensors[:n_x]]
        M, dev = xs[0].shape[0], xs[0].device
        need_grad = spec["grad"]        # (ctx.needs_input_grad ignores torch.no_grad())
        seeds = [[next_seed() for _ in range(n)] for n in spec["n_layers"]]
        chains, ys, keep, metas = _chain_build_forward(mods, xs, training, need_grad, seeds)
        dp = _chain_dp.struct() if _chain_dp is not None else None
        check(lib.fr_mlp_chain_forward_dp(chains, len(mods), M, 1 if training else 0, 1 if need_grad else 0,
                                          ptr(seed_dev(dev)), ptr(chain_bars(dev)),
                                          ctypes.byref(dp) if dp is not None else None, stream_ptr()), "fr_mlp_chain_forward")
        if need_grad:
            if not training:
                raise NotImplementedError("MLPChainGroup backward is implemented for training mode")
            ctx.spec, ctx.metas, ctx.ws = spec, metas, keep
            ctx.save_for_backward(*xs, *ys)
        return tuple(ys)

    @staticmethod
    def backward(ctx, *dys):
        import ctypes
        lib = load()
        _bind_thread()
        spec = ctx.spec
        mods, n_x = spec["mods"], spec["n_x"]
        saved = ctx.saved_tensors
        xs, ys = saved[:n_x], saved[n_x:]
        M, dev = xs[0].shape[0], xs[0].device
        shared = n_x == 1 and len(mods) > 1
        want_dx = [ctx.needs_input_grad[1 + (c if n_x > 1 else 0)] for c in range(len(mods))]
        chains, grads_p, dxs, dx_sum, keep = _chain_build_backward(mods, xs, ys, dys, ctx.metas, ctx.ws, want_dx)
        dp = _chain_dp.struct() if _chain_dp is not None else None
        check(lib.fr_mlp_chain_backward_dp(chains, len(mods), M, ptr(seed_dev(dev)), ptr(dx_sum), ptr(chain_bars(dev)),
                                           ctypes.byref(dp) if dp is not None else None, stream_ptr()),
              "fr_mlp_chain_backward")
        gx = [dx_sum] if shared else (dxs if n_x > 1 else [dxs[0]])
        return (None, *gx, *grads_p)


def chain_dp_emulate(rank_mods, rank_xs, rank_dys):
    """Single-process emulation of a data-parallel chain call (test / profiling harness): P emulated ranks on ONE device,
    the host running segment after segment, rank after rank, instead of the flag barriers (fr_chain_dp.barriers = 0).
    rank_mods[r]: chain modules of rank r (replicas); rank_xs[r] / rank_dys[r]: its inputs / output gradients.
    -> (ys[r][c], dxs[r], grads[r]) with grads[r] in _chain_params order (this rank's share)."""
    import ctypes

    from ._lib import ChainDp
    lib = load()
    world = len(rank_mods)
    dev = rank_xs[0][0].device
    M = rank_xs[0][0].shape[0]
    bufs = [torch.zeros(ChainDP.XCHG_BYTES, dtype=torch.uint8, device=dev) for _ in range(world)]
    status = torch.zeros(1, dtype=torch.int32, device=dev)

    def dp_struct(r, seg, parity):
        d = ChainDp()
        d.rank, d.world, d.xchg_bytes, d.barriers, d.segment, d.parity = r, world, ChainDP.XCHG_BYTES, 0, seg, parity
        for k in range(world):
            d.xchg[k] = bufs[k].data_ptr()
        d.status_flags = status.data_ptr()
        return d

    seeds = [[next_seed() for _ in _pairs_of(m)] for m in rank_mods[0]]      # the ranks share the dropout seeds
    fw = [_chain_build_forward(rank_mods[r], rank_xs[r], True, True, seeds) for r in range(world)]
    nseg = lib.fr_mlp_chain_segments(fw[0][0], len(rank_mods[0]), 1, 0, world)
    exch = 0
    for seg in range(nseg):
        for r in range(world):
            d = dp_struct(r, seg, (exch + seg) & 1)
            check(lib.fr_mlp_chain_forward_dp(fw[r][0], len(rank_mods[r]), M, 1, 1, ptr(seed_dev(dev)), ptr(chain_bars(dev)),
                                              ctypes.byref(d), stream_ptr()), "fr_mlp_chain_forward_dp")
    exch += nseg - 1
    outs = []
    bw = []
    for r in range(world):
        chains, ys, keep, metas = fw[r]
        want = [True] * len(rank_mods[r])
        bw.append(_chain_build_backward(rank_mods[r], rank_xs[r], ys, rank_dys[r], metas, keep, want))
    nseg_b = lib.fr_mlp_chain_segments(bw[0][0], len(rank_mods[0]), 1, 1, world)
    for seg in range(nseg_b):
        for r in range(world):
            d = dp_struct(r, seg, (exch + seg) & 1)
            check(lib.fr_mlp_chain_backward_dp(bw[r][0], len(rank_mods[r]), M, ptr(seed_dev(dev)), ptr(bw[r][3]),
                                               ptr(chain_bars(dev)), ctypes.byref(d), stream_ptr()),
                  "fr_mlp_chain_backward_dp")
    torch.cuda.synchronize()
    for r in range(world):
        shared = len(rank_xs[r]) == 1 and len(rank_mods[r]) > 1
        outs.append((fw[r][1], [bw[r][3]] if shared else bw[r][2], bw[r][1]))
    return outs


def _chain_check(mods, xs):
    """(training, grad, n_layers) when the fused chain kernels take `mods` over `xs`, else None (see mlp_chain)"""
    if not chain_enabled() or not xs[0].is_cuda or len(mods) > 4 or len(xs) not in (1, len(mods)):
        return None
    M = xs[0].shape[0]
    training = mods[0].training
    grad = torch.is_grad_enabled() and (any(x.requires_grad for x in xs) or
                                        any(p.requires_grad for m in mods for p in m.parameters()))
    if any(m.training != training for m in mods) or any(x.dim() != 2 or x.shape[0] != M for x in xs):
        return None
    if grad and (not training or M > CHAIN_MAX_TRAIN_ROWS):
        return None
    if training and M > CHAIN_MAX_TRAIN_ROWS:
        return None
    lib = load()
    n_layers, total = [], 0
    for c, m in enumerate(mods):
        arr, pairs = _chain_layers(m, training)
        if arr is None or not lib.fr_mlp_chain_eligible(arr, len(pairs), M):
            return None
        if pairs[0][0].in_features != xs[c if len(xs) > 1 else 0].shape[1]:
            return None
        n_layers.append(len(pairs))
        total += len(pairs)
    if total > 24:
        return None
    return training, grad, n_layers


def mlp_chain_would_fuse(mods, xs):
    """True when MLPLayers.forward / mlp_chain run these modules on the fused chain kernels (which honour bn_repeat)"""
    return _chain_check(mods, xs) is not None


def mlp_chain(mods, xs):
    """Run MLPLayers modules `mods` over inputs `xs` (one shared input, or one per module; all [M, K]) with the fused chain
    kernels.  Returns the list of outputs, or None when the configuration is outside the kernels' rules (the caller then
    uses the per-layer path)."""
    chk = _chain_check(mods, xs)
    if chk is None:
        return None
    training, grad, n_layers = chk
    M = xs[0].shape[0]
    params = [p for m, nl in zip(mods, n_layers) for p in _chain_params(_chain_layers(m, training)[1])]
    if not training and M > CHAIN_EVAL_CHUNK:
        outs = [[] for _ in mods]
        for a in range(0, M, CHAIN_EVAL_CHUNK):
            part = mlp_chain(mods, [x[a:a + CHAIN_EVAL_CHUNK] for x in xs])
            for o, p_ in zip(outs, part):
                o.append(p_)
        return [torch.cat(o) for o in outs]
    spec = {"mods": list(mods), "n_x": len(xs), "training": training, "n_layers": n_layers, "grad": bool(grad)}
    return list(MLPChainGroup.apply(spec, *xs, *params))


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


class RowDot(torch.autograd.Function):
    """torch.mul(a, b).sum(-1)  (pfcn_pmf.py:172, fairgo_pmf.py:169)"""

    @staticmethod
    def forward(ctx, A, B):
        lib = load()
        A, B = A.contiguous(), B.contiguous()
        M, d = A.shape
        out = torch.empty(M, dtype=torch.float32, device=A.device)
        check(lib.fr_rowdot_forward(ptr(A), ptr(B), M, d, ptr(out), stream_ptr()), "fr_rowdot_forward")
        ctx.save_for_backward(A, B)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = load()
        A, B = ctx.saved_tensors
        M, d = A.shape
        dA = torch.empty_like(A) if ctx.needs_input_grad[0] else None
        dB = torch.empty_like(B) if ctx.needs_input_grad[1] else None
        check(lib.fr_rowdot_backward(ptr(A), ptr(B), ptr(dout.contiguous()), M, d, ptr(dA), ptr(dB), stream_ptr()),
              "fr_rowdot_backward")
        return dA, dB


class CosineSim(torch.autograd.Function):
    """nn.CosineSimilarity()(a, b)  (pfcn_dmf.py:176,190-191)"""
    EPS = 1e-8

    @staticmethod
    def forward(ctx, A, B):
        lib = load()
        A, B = A.contiguous(), B.contiguous()
        M, d = A.shape
        out = torch.empty(M, dtype=torch.float32, device=A.device)
        norms = torch.empty(2 * M, dtype=torch.float32, device=A.device)
        check(lib.fr_cosine_forward(ptr(A), ptr(B), M, d, CosineSim.EPS, ptr(out), ptr(norms), stream_ptr()),
              "fr_cosine_forward")
        ctx.save_for_backward(A, B, out, norms)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = load()
        A, B, out, norms = ctx.saved_tensors
        M, d = A.shape
        dA = torch.empty_like(A) if ctx.needs_input_grad[0] else None
        dB = torch.empty_like(B) if ctx.needs_input_grad[1] else None
        check(lib.fr_cosine_backward(ptr(A), ptr(B), ptr(out), ptr(norms), ptr(dout.contiguous()), M, d, CosineSim.EPS,
                                     ptr(dA), ptr(dB), stream_ptr()), "fr_cosine_backward")
        return dA, dB


class BprOuter(torch.autograd.Function):
    """BPRLoss over the [B,B] broadcast score matrices of PFCN_BiasedMF (pfcn_biasedmf.py:189-192).
    Inputs: dotp[B], dotn[B], user_bias[B], pos_item_bias[B], neg_item_bias[B], global_bias[1]."""

    @staticmethod
    def forward(ctx, dotp, dotn, ub, pib, nib, gb):
        lib = load()
        B = dotp.numel()
        args = [t.contiguous().view(-1) for t in (dotp, dotn, ub, pib, nib, gb)]
        dev = dotp.device
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        g = [torch.empty(B, dtype=torch.float32, device=dev) for _ in range(5)]
        check(lib.fr_bpr_outer_loss(*[ptr(t) for t in args], B, ptr(loss), ptr(g[0]), ptr(g[1]), ptr(g[2]), ptr(g[3]),
                                    ptr(g[4]), stream_ptr()), "fr_bpr_outer_loss")
        ctx.shapes = (dotp.shape, dotn.shape, ub.shape, pib.shape, nib.shape, gb.shape)
        ctx.save_for_backward(g[0], g[1], g[2], g[3])
        return loss.view(())

    @staticmethod
    def backward(ctx, gl):
        d_dotp, d_dotn, d_pib, d_nib = ctx.saved_tensors
        s = ctx.shapes
        zeros_ub = torch.zeros(s[2], dtype=torch.float32, device=gl.device)    # cancels between pos and neg
        zeros_gb = torch.zeros(s[5], dtype=torch.float32, device=gl.device)
        return ((d_dotp * gl).view(s[0]), (d_dotn * gl).view(s[1]), zeros_ub, (d_pib * gl).view(s[3]),
                (d_nib * gl).view(s[4]), zeros_gb)


def _scaled_sum(tensors, scale, divide):
    import ctypes
    lib = load()
    arr = (ctypes.c_void_p * len(tensors))(*[ptr(t) for t in tensors])
    out = torch.empty_like(tensors[0])
    check(lib.fr_scaled_sum(arr, len(tensors), out.numel(), float(scale), 1 if divide else 0, ptr(out), stream_ptr()),
          "fr_scaled_sum")
    return out


class SumDiv(torch.autograd.Function):
    """(x0 + x1 + ...) / denom: cm-mode filter averaging (pfcn_mlp.py:158-165, fairgo_pmf.py:163-168)"""

    @staticmethod
    def forward(ctx, denom, *xs):
        ctx.denom = float(denom)
        ctx.n = len(xs)
        return _scaled_sum([x.contiguous() for x in xs], denom, True)

    @staticmethod
    def backward(ctx, dY):
        g = _scaled_sum([dY.contiguous()], ctx.denom, True)
        return (None,) + (g,) * ctx.n


class WeightedSum(torch.autograd.Function):
    """(x0 + x1 + ...) * scale (WAP mean over the ego-network layers, fairgo_pmf.py:206-208)"""

    @staticmethod
    def forward(ctx, scale, *xs):
        ctx.scale = float(scale)
        ctx.n = len(xs)
        return _scaled_sum([x.contiguous() for x in xs], scale, False)

    @staticmethod
    def backward(ctx, dY):
        g = _scaled_sum([dY.contiguous()], ctx.scale, False)
        return (None,) + (g,) * ctx.n


class ConcatCols(torch.autograd.Function):
    """torch.cat(xs, dim=1)"""

    @staticmethod
    def forward(ctx, *xs):
        lib = load()
        M = xs[0].shape[0]
        widths = [x.shape[1] for x in xs]
        W = sum(widths)
        out = torch.empty((M, W), dtype=torch.float32, device=xs[0].device)
        c0 = 0
        for x, w in zip(xs, widths):
            x = x.contiguous()
            check(lib.fr_copy_cols(ptr(x), w, 0, ptr(out), W, c0, M, w, stream_ptr()), "fr_copy_cols")
            c0 += w
        ctx.widths = widths
        return out

    @staticmethod
    def backward(ctx, dY):
        lib = load()
        dY = dY.contiguous()
        M, W = dY.shape
        outs, c0 = [], 0
        for i, w in enumerate(ctx.widths):
            if ctx.needs_input_grad[i]:
                g = torch.empty((M, w), dtype=torch.float32, device=dY.device)
                check(lib.fr_copy_cols(ptr(dY), W, c0, ptr(g), w, 0, M, w, stream_ptr()), "fr_copy_cols")
                outs.append(g)
            else:
                outs.append(None)
            c0 += w
        return tuple(outs)


class MseLoss(_ScalarLoss):
    """nn.MSELoss()(pred, target)  (fairgo_pmf.py:170-171)"""

    @staticmethod
    def forward(ctx, pred, target):
        lib = load()
        shape = pred.shape
        p = pred.contiguous().view(-1)
        loss = torch.empty(1, dtype=torch.float32, device=p.device)
        dp = torch.empty_like(p)
        check(lib.fr_mse_loss(ptr(p), ptr(target.contiguous().view(-1)), p.numel(), ptr(loss), ptr(dp), stream_ptr()),
              "fr_mse_loss")
        ctx.grads = (dp.view(shape), None)
        return loss.view(())


class Act(torch.autograd.Function):
    """a stand-alone activation module (nn.Sigmoid() etc. outside an MLPLayers)"""

    @staticmethod
    def forward(ctx, x, act):
        lib = load()
        x = x.contiguous()
        y = torch.empty_like(x)
        check(lib.fr_act_forward(ptr(x), act, x.numel(), ptr(y), stream_ptr()), "fr_act_forward")
        ctx.save_for_backward(y)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, dY):
        lib = load()
        (y,) = ctx.saved_tensors
        dX = torch.empty_like(y)
        check(lib.fr_act_backward(ptr(dY.contiguous()), ptr(y), ctx.act, y.numel(), ptr(dX), stream_ptr()),
              "fr_act_backward")
        return dX, None


class Dropout(torch.autograd.Function):
    """inverted dropout with a counter-based mask (seed + device seed offset, element index): the backward pass applies the
    same mask to the incoming gradient"""

    @staticmethod
    def forward(ctx, x, p, seed):
        x = x.contiguous()
        y = torch.empty_like(x)
        check(load().fr_dropout(ptr(x), float(p), seed, ptr(seed_dev(x.device)), x.numel(), ptr(y), stream_ptr()),
              "fr_dropout")
        ctx.cfg = (float(p), seed)
        return y

    @staticmethod
    def backward(ctx, dy):
        p, seed = ctx.cfg
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        check(load().fr_dropout(ptr(dy), p, seed, ptr(seed_dev(dy.device)), dy.numel(), ptr(dx), stream_ptr()),
              "fr_dropout")
        return dx, None, None


def clamp_div(x, hi):
    """clamp(x, 0, hi) / hi (inference only: fairgo_pmf.py:248, focf.py:150)"""
    lib = load()
    x = x.contiguous()
    y = torch.empty_like(x)
    check(lib.fr_clamp_div(ptr(x), float(hi), x.numel(), ptr(y), stream_ptr()), "fr_clamp_div")
    return y


class SpmmMatrix:
    """A static sparse matrix in CSR on the device together with the chunk plans of A and A^T (spmm.cu)."""

    def __init__(self, csr, device, chunk=128):
        import ctypes

        import numpy as np
        from ._lib import SpmmPlan
        self.shape = csr.shape
        self._keep = []
        self.plans = []
        lib = load()
        for mat in (csr, csr.T.tocsr()):
            mat.sort_indices()
            row_off = np.ascontiguousarray(mat.indptr, dtype=np.int64)
            n_rows = mat.shape[0]
            sizes = [ctypes.c_int64() for _ in range(4)]
            check(lib.fr_spmm_plan_sizes(row_off.ctypes.data, n_rows, chunk, *[ctypes.byref(s) for s in sizes]),
                  "fr_spmm_plan_sizes")
            nc, nm, ns, ne = [s.value for s in sizes]
            host = [np.zeros(max(nc, 1), np.int32) for _ in range(4)] + [np.zeros(max(nm, 1), np.int32),
                                                                         np.zeros(nm + 1, np.int32),
                                                                         np.zeros(max(ne, 1), np.int32)]
            check(lib.fr_spmm_plan_fill(row_off.ctypes.data, n_rows, chunk, *[h.ctypes.data for h in host]),
                  "fr_spmm_plan_fill")
            dev = [torch.from_numpy(h).to(device) for h in host]
            col = torch.from_numpy(np.ascontiguousarray(mat.indices, dtype=np.int32)).to(device)
            val = torch.from_numpy(np.ascontiguousarray(mat.data, dtype=np.float32)).to(device)
            plan = SpmmPlan(*[t.data_ptr() for t in dev], nc, nm, ns, ne)
            self._keep.append((dev, col, val))
            self.plans.append((plan, col, val, ns))

    def apply(self, X, transpose=False):
        import ctypes
        lib = load()
        plan, col, val, ns = self.plans[1 if transpose else 0]
        X = X.contiguous()
        d = X.shape[1]
        Y = torch.empty((self.shape[1] if transpose else self.shape[0], d), dtype=torch.float32, device=X.device)
        partial = torch.empty(max(ns, 1) * d, dtype=torch.float32, device=X.device)
        check(lib.fr_spmm_csr(ctypes.byref(plan), ptr(col), ptr(val), ptr(X), d, ptr(Y), ptr(partial), stream_ptr()),
              "fr_spmm_csr")
        return Y


class Spmm(torch.autograd.Function):
    """torch.sparse.mm(A, X) with a static A  (fairgo_pmf.py:201)"""

    @staticmethod
    def forward(ctx, X, mat):
        ctx.mat = mat
        return mat.apply(X, False)

    @staticmethod
    def backward(ctx, dY):
        return ctx.mat.apply(dY, True), None


def biased_score(dot, ub, ib, gb, act):
    """act(dot + b_u + b_i + b_g) (inference: pfcn_biasedmf.py:170-181)"""
    lib = load()
    dot = dot.contiguous().view(-1)
    out = torch.empty_like(dot)
    check(lib.fr_biased_score(ptr(dot), ptr(ub.contiguous().view(-1)), ptr(ib.contiguous().view(-1)),
                              ptr(gb.detach().contiguous().view(-1)), dot.numel(), act, ptr(out), stream_ptr()),
          "fr_biased_score")
    return out


class AdamGroup:
    """torch.optim.Adam(params, lr, weight_decay) semantics (betas 0.9/0.999, eps 1e-8, L2 form, per-parameter step
    counts, parameters without a gradient skipped) on fr_adam_multi: one launch per 48 parameter tensors.  The step
    counts live on the device (bumped on the stream right before each update), so a step captured in a CUDA graph
    keeps the right bias correction on every replay."""

    def __init__(self, params, lr=1e-3, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8):
        self.params = [p for p in params]
        self.lr, self.weight_decay, self.betas, self.eps = lr, weight_decay, betas, eps
        self.state = {}
        self._steps = None          # int32 [n_params] on the device
        self._index = {id(p): i for i, p in enumerate(self.params)}

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    def _ensure(self, p):
        st = self.state.get(p)
        if st is None:
            st = self.state[p] = {"exp_avg": torch.zeros_like(p.data), "exp_avg_sq": torch.zeros_like(p.data)}
        if self._steps is None:
            self._steps = torch.zeros(len(self.params), dtype=torch.int32, device=p.device)
        return st

    def init_state(self):
        """allocate every moment buffer up front (needed before capturing a step in a CUDA graph)"""
        for p in self.params:
            if p.requires_grad:
                self._ensure(p)

    def step(self):
        from ._lib import AdamEntry
        lib = load()
        entries = []
        for p in self.params:
            if p.grad is None or not p.requires_grad:
                continue
            st = self._ensure(p)
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            st["_g"] = g                                    # keep alive until the kernel ran
            entries.append(AdamEntry(p.data.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(),
                                     st["exp_avg_sq"].data_ptr(), p.numel(), 0,
                                     self._steps.data_ptr() + 4 * self._index[id(p)]))
        if not entries:
            return
        arr = (AdamEntry * len(entries))(*entries)
        check(lib.fr_adam_multi(arr, len(entries), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                stream_ptr()), "fr_adam_multi")

    def state_dict(self):
        steps = self._steps.cpu().tolist() if self._steps is not None else [0] * len(self.params)
        return {"state": {i: {"step": steps[i], **{k: v for k, v in self.state[p].items() if k != "_g"}}
                          for i, p in enumerate(self.params) if p in self.state},
                "param_groups": [{"lr": self.lr, "weight_decay": self.weight_decay, "betas": self.betas, "eps": self.eps}]}

    def load_state_dict(self, sd):
        for i, st in sd["state"].items():
            p = self.params[int(i)]
            cur = self._ensure(p)
            cur["exp_avg"].copy_(st["exp_avg"])
            cur["exp_avg_sq"].copy_(st["exp_avg_sq"])
            self._steps[int(i)] = int(st.get("step", 0))
