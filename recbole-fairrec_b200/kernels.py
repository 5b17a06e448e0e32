"""Tensor-level wrappers over the C ABI (include/fairrec_b200.h).  torch is used for device memory and
streams only; every computation below is one of this package's own CUDA kernels."""
import ctypes

import torch

from . import _lib
from ._lib import FocfStep, FullSort, check, load, ptr, stream_ptr


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def sort_pairs(keys, vals=None, key_bits=32):
    """stable radix sort of uint32 keys (stored in int32 tensors) with their values; vals=None -> identity"""
    lib = load()
    n = keys.numel()
    ko, vo = torch.empty_like(keys), torch.empty_like(keys)
    ws = _ws(lib.fr_sort_pairs_workspace_bytes(n), keys.device)
    check(lib.fr_sort_pairs_u32(ptr(keys), ptr(vals), ptr(ko), ptr(vo), n, key_bits, ptr(ws), ws.numel(),
                                stream_ptr()), "fr_sort_pairs_u32")
    return ko, vo


def pair_scores(U, I, uid, iid, transform=_lib.TRANSFORM_NONE, max_rating=1.0):
    lib = load()
    out = torch.empty(uid.numel(), dtype=torch.float32, device=U.device)
    check(lib.fr_pair_scores(ptr(U), ptr(I), ptr(uid), ptr(iid), uid.numel(), U.shape[1], transform,
                             float(max_rating), ptr(out), stream_ptr()), "fr_pair_scores")
    return out


def fill_adam(s, adam):
    """Adam hyper-parameters, moments and (lazy_exact) per-row state of `adam` (FOCF.init_adam) into a FocfStep"""
    s.mU, s.vU, s.mI, s.vI = ptr(adam["mU"]), ptr(adam["vU"]), ptr(adam["mI"]), ptr(adam["vI"])
    s.lr, s.beta1, s.beta2 = float(adam["lr"]), float(adam["beta1"]), float(adam["beta2"])
    s.eps, s.weight_decay = float(adam["eps"]), float(adam["weight_decay"])
    if adam.get("mode", "dense_exact") == "lazy_exact":
        s.adam_mode = _lib.ADAM_LAZY_EXACT
        s.last_step_u, s.last_step_i, s.adam_scalars = ptr(adam["last_u"]), ptr(adam["last_i"]), ptr(adam["scalars"])
        s.scalars_cap = adam["scalars"].numel() // 2
        s.scalars_filled = ctypes.pointer(adam["filled"])
    else:
        s.adam_mode = _lib.ADAM_DENSE_EXACT
    s.no_fused = 1 if adam.get("no_fused") else 0


class FocfEngine:
    """Owns the workspace of the FOCF training kernels for one (U, I) pair and drives
    fr_focf_forward / backward / adam / train_step."""

    def __init__(self, n_users, n_items, d, max_batch, device):
        self.lib = load()
        self.n_users, self.n_items, self.d = int(n_users), int(n_items), int(d)
        self.device = device
        self.max_batch = 0
        self.ws = None
        self.flags = torch.zeros(1, dtype=torch.int32, device=device)
        self.loss = torch.zeros(1, dtype=torch.float32, device=device)
        self._stage = None          # persistent device staging buffer of train_step_packed
        self._fast = self._fast_key = None
        self._ensure(max_batch)

    def _ensure(self, B):
        if B <= self.max_batch:
            return
        B = int(B * 1.25) + 64
        nbytes = self.lib.fr_focf_workspace_bytes(self.n_users, self.n_items, self.d, B)
        self.ws = _ws(nbytes, self.device)
        check(self.lib.fr_focf_workspace_init(ptr(self.ws), self.ws.numel(), self.n_users, self.n_items, self.d, B,
                                              stream_ptr()), "fr_focf_workspace_init")
        self.max_batch = B
        self.pred_buf = torch.empty(B, dtype=torch.float32, device=self.device)
        self._fast = None           # the persistent argument struct points into the buffers just replaced

    def _step(self, U, I, batch, objective, fair_weight, loss_out=None, norm=None):
        uid, iid, rating, sst, contiguous = batch
        B = uid.numel()
        self._ensure(B)
        s = FocfStep()
        s.U, s.I = ptr(U), ptr(I)
        s.n_users, s.n_items, s.d = self.n_users, self.n_items, self.d
        s.uid, s.iid, s.rating, s.sst, s.B = ptr(uid), ptr(iid), ptr(rating), ptr(sst), B
        s.items_contiguous = 1 if contiguous else 0
        s.objective = objective
        s.fair_weight = float(fair_weight)
        s.pred = ptr(self.pred_buf)
        s.loss = ptr(self.loss if loss_out is None else loss_out)
        s.status_flags = ptr(self.flags)
        s.workspace, s.workspace_bytes = ptr(self.ws), self.ws.numel()
        if norm is not None:
            s.norm_B, s.norm_J = int(norm[0]), int(norm[1])
        return s

    def forward(self, U, I, batch, objective, fair_weight, loss_out=None, norm=None):
        s = self._step(U, I, batch, objective, fair_weight, loss_out, norm)
        check(self.lib.fr_focf_forward(ctypes.byref(s), stream_ptr()), "fr_focf_forward")
        return s

    def backward(self, s, dU, dI, grad_scale=1.0):
        s.dU, s.dI = ptr(dU), ptr(dI)
        check(self.lib.fr_focf_backward(ctypes.byref(s), float(grad_scale), stream_ptr()), "fr_focf_backward")

    def train_step(self, U, I, adam, batch, objective, fair_weight, loss_out=None):
        """adam: dict(mU, vU, mI, vI, step, lr, beta1, beta2, eps, weight_decay)"""
        s = self._step(U, I, batch, objective, fair_weight, loss_out)
        fill_adam(s, adam)
        s.step = int(adam["step"])
        check(self.lib.fr_focf_train_step(ctypes.byref(s), stream_ptr()), "fr_focf_train_step")
        return s

    def adam_flush(self, U, I, adam):
        """lazy_exact: replay the pending steps of every row (fr_focf_adam_flush); a no-op in dense_exact mode"""
        if adam.get("mode", "dense_exact") != "lazy_exact" or int(adam["step"]) < 1:
            return
        s = FocfStep()
        s.U, s.I = ptr(U), ptr(I)
        s.n_users, s.n_items, s.d = self.n_users, self.n_items, self.d
        fill_adam(s, adam)
        s.step = int(adam["step"])
        check(self.lib.fr_focf_adam_flush(ctypes.byref(s), stream_ptr()), "fr_focf_adam_flush")

    def train_step_packed(self, U, I, adam, packed, contiguous, objective, fair_weight, loss_out):
        """The eager fused step for a host batch packed into ONE pinned buffer (focf.pack_host_batch: int32 user ids |
        int32 item ids | f32 ratings | f32 attribute values, n entries each).  The batch is staged into a persistent
        device buffer with a single async H2D copy and the argument struct lives across steps: only the batch pointers /
        size, the Adam step count and the loss pointer are rewritten, so the per-step host work is one copy, a handful of
        integer stores and one library call (the generic path builds four tensor views and a fresh struct per step, which
        costs more host time than the step takes on the device).  One staging buffer is enough: the copy of step t+1 is
        enqueued behind the kernels of step t on the same stream."""
        buf, n = packed
        n = int(n)
        self._ensure(n)
        nbytes = 16 * n
        if buf.numel() != nbytes:
            raise ValueError("packed batch: the buffer must hold exactly 16 bytes per entry")
        st = self._stage
        if st is None or st.numel() < nbytes:
            st = self._stage = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=self.device)
        key = (ptr(U), ptr(I), ptr(adam["mU"]), ptr(adam["vU"]), ptr(adam["mI"]), ptr(adam["vI"]), objective,
               float(fair_weight), adam["lr"], adam["beta1"], adam["beta2"], adam["eps"], adam["weight_decay"])
        s = self._fast
        if s is None or self._fast_key != key:
            s = FocfStep()
            s.U, s.I, s.mU, s.vU, s.mI, s.vI = key[:6]
            s.n_users, s.n_items, s.d = self.n_users, self.n_items, self.d
            s.objective, s.fair_weight = objective, float(fair_weight)
            s.pred, s.status_flags = ptr(self.pred_buf), ptr(self.flags)
            s.workspace, s.workspace_bytes = ptr(self.ws), self.ws.numel()
            s.lr, s.beta1, s.beta2 = float(adam["lr"]), float(adam["beta1"]), float(adam["beta2"])
            s.eps, s.weight_decay = float(adam["eps"]), float(adam["weight_decay"])
            self._fast, self._fast_key = s, key
        st[:nbytes].copy_(buf, non_blocking=True)
        base = ptr(st)
        s.uid, s.iid, s.rating, s.sst, s.B = base, base + 4 * n, base + 8 * n, base + 12 * n, n
        s.items_contiguous = 1 if contiguous else 0
        s.loss = ptr(loss_out)
        s.step = int(adam["step"])
        check(self.lib.fr_focf_train_step(ctypes.byref(s), stream_ptr()), "fr_focf_train_step")
        return s

    def train_steps_host(self, U, I, adam, packed_batches, contiguous, objective, fair_weight):
        """fr_focf_train_steps_host: the optimisation steps of a LIST of host batches (each packed into one pinned buffer like
        train_step_packed's) in ONE library call -- per step one H2D copy, the fused step, one D2H copy of its loss; the
        host loop (copies, launches, waiting for the previous step's loss) runs in the library instead of the interpreter.
        adam["step"] is the count BEFORE the first batch and is advanced by the caller.  Returns the pinned float32 tensor
        of the per-step losses (complete on return)."""
        k = len(packed_batches)
        rows = [int(n) for _, n in packed_batches]
        for (buf, _), n in zip(packed_batches, rows):
            if buf.numel() != 16 * n or buf.is_cuda:
                raise ValueError("packed batch: a host buffer of exactly 16 bytes per entry is expected")
        if k == 0:
            return torch.zeros(0, dtype=torch.float32)
        self._ensure(max(rows))
        nbytes = 16 * max(rows)
        st = self._stage
        if st is None or st.numel() < nbytes:
            st = self._stage = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=self.device)
        if getattr(self, "_loss2", None) is None:
            self._loss2 = torch.zeros(2, dtype=torch.float32, device=self.device)
        host = getattr(self, "_loss_host", None)
        if host is None or host.numel() < k:
            host = self._loss_host = torch.zeros(max(k, 64), dtype=torch.float32).pin_memory()
        s = FocfStep()
        s.U, s.I = ptr(U), ptr(I)
        fill_adam(s, adam)
        s.n_users, s.n_items, s.d = self.n_users, self.n_items, self.d
        s.objective, s.fair_weight = objective, float(fair_weight)
        s.pred, s.status_flags = ptr(self.pred_buf), ptr(self.flags)
        s.workspace, s.workspace_bytes = ptr(self.ws), self.ws.numel()
        s.items_contiguous = 1 if contiguous else 0
        s.step = int(adam["step"]) + 1
        ptrs = (ctypes.c_void_p * k)(*[buf.data_ptr() for buf, _ in packed_batches])
        nrow = (ctypes.c_int32 * k)(*rows)
        check(self.lib.fr_focf_train_steps_host(ctypes.byref(s), k, ptrs, nrow, ptr(st), st.numel(), ptr(self._loss2),
                                                host.data_ptr(), stream_ptr()), "fr_focf_train_steps_host")
        return host[:k]

    def set_counters(self, plan_cursor=-1, adam_step=None, stride=0):
        """device-resident counters of the workspace; None / negative cursor / stride 0 leave a counter unchanged"""
        check(self.lib.fr_focf_set_counters(ptr(self.ws), self.ws.numel(), self.n_users, self.n_items, self.d,
                                            self.max_batch, int(plan_cursor),
                                            -2147483648 if adam_step is None else int(adam_step), int(stride),
                                            stream_ptr()), "fr_focf_set_counters")

    def planned_step(self, U, I, adam, plan, train, objective, fair_weight, loss_buf):
        """fr_focf_step for planned batches (device-resident cursor / batch size / Adam step): the struct is the same
        for every batch of the epoch, so the launch sequence can be captured once in a CUDA graph and replayed.
        plan: dict(desc, items, offs, len, cols=(uid, iid, rating, sst) scratch columns); train: TrainData."""
        uid, iid, rating, sst = plan["cols"]
        cap = uid.numel()
        if cap > self.max_batch:
            self._ensure(cap)
        s = FocfStep()
        s.U, s.I = ptr(U), ptr(I)
        s.n_users, s.n_items, s.d = self.n_users, self.n_items, self.d
        s.uid, s.iid, s.rating, s.sst, s.B = ptr(uid), ptr(iid), ptr(rating), ptr(sst), cap
        s.items_contiguous = 1
        s.objective, s.fair_weight = objective, float(fair_weight)
        s.pred, s.loss, s.status_flags = ptr(self.pred_buf), ptr(loss_buf), ptr(self.flags)
        s.workspace, s.workspace_bytes = ptr(self.ws), self.ws.numel()
        s.mU, s.vU, s.mI, s.vI = ptr(adam["mU"]), ptr(adam["vU"]), ptr(adam["mI"]), ptr(adam["vI"])
        s.step = 0   # device-resident counter
        s.lr, s.beta1, s.beta2 = float(adam["lr"]), float(adam["beta1"]), float(adam["beta2"])
        s.eps, s.weight_decay = float(adam["eps"]), float(adam["weight_decay"])
        s.plan_desc, s.plan_items, s.plan_offs = ptr(plan["desc"]), ptr(plan["items"]), ptr(plan["offs"])
        s.plan_len = int(plan["len"])
        s.item_off, s.train_uid = ptr(train.item_off), ptr(train.train_uid)
        s.train_rating, s.sst_of_user = ptr(train.train_rating), ptr(train.sst_of_user)
        return s

    def run_planned(self, s):
        check(self.lib.fr_focf_train_step(ctypes.byref(s), stream_ptr()), "fr_focf_train_step")

    def run_prepare(self, s):
        check(self.lib.fr_focf_step_prepare(ctypes.byref(s), stream_ptr()), "fr_focf_step_prepare")

    def run_compute(self, s):
        check(self.lib.fr_focf_step_compute(ctypes.byref(s), stream_ptr()), "fr_focf_step_compute")

    def adam_dense(self, U, I, dU, dI, adam):
        s = FocfStep()
        s.U, s.I, s.dU, s.dI = ptr(U), ptr(I), ptr(dU), ptr(dI)
        s.n_users, s.n_items, s.d = self.n_users, self.n_items, self.d
        s.mU, s.vU, s.mI, s.vI = ptr(adam["mU"]), ptr(adam["vU"]), ptr(adam["mI"]), ptr(adam["vI"])
        s.step = int(adam["step"])
        s.lr, s.beta1, s.beta2 = float(adam["lr"]), float(adam["beta1"]), float(adam["beta2"])
        s.eps, s.weight_decay = float(adam["eps"]), float(adam["weight_decay"])
        s.workspace, s.workspace_bytes = ptr(self.ws), self.ws.numel()
        check(self.lib.fr_focf_adam(ctypes.byref(s), stream_ptr()), "fr_focf_adam")

    def read_flags(self):
        """host read of the device status word (synchronises); returns and clears it"""
        f = int(self.flags.item())
        if f:
            self.flags.zero_()
        return f


def fullsort_topk(U, I_shard, users, hist_off, hist_items, K, transform, max_rating, item_base=0,
                  score_mode=_lib.SCORE_EXACT_FP32):
    """returns (topk_id int32 [n,K], topk_score f32 [n,K]) of this item shard"""
    lib = load()
    n, d, nl = users.numel(), U.shape[1], I_shard.shape[0]
    ids = torch.empty((n, K), dtype=torch.int32, device=U.device)
    sc = torch.empty((n, K), dtype=torch.float32, device=U.device)
    ws = _ws(lib.fr_fullsort_workspace_bytes(n, K, nl, d, int(score_mode)), U.device)
    a = FullSort()
    a.U, a.I_shard, a.d, a.n_items_local, a.item_base = ptr(U), ptr(I_shard), d, nl, int(item_base)
    a.users, a.n, a.hist_off, a.hist_items = ptr(users), n, ptr(hist_off), ptr(hist_items)
    a.K, a.transform, a.max_rating, a.score_mode = int(K), int(transform), float(max_rating), int(score_mode)
    a.topk_id, a.topk_score, a.workspace, a.workspace_bytes = ptr(ids), ptr(sc), ptr(ws), ws.numel()
    check(lib.fr_fullsort_topk(ctypes.byref(a), stream_ptr()), "fr_fullsort_topk")
    return ids, sc


def topk_merge(ids_in, scores_in):
    """[P, n, K] per-shard lists -> global [n, K]"""
    lib = load()
    P, n, K = ids_in.shape
    ids = torch.empty((n, K), dtype=torch.int32, device=ids_in.device)
    sc = torch.empty((n, K), dtype=torch.float32, device=ids_in.device)
    check(lib.fr_topk_merge(ptr(ids_in), ptr(scores_in), P, n, K, ptr(ids), ptr(sc), stream_ptr()), "fr_topk_merge")
    return ids, sc


def hits(topk_id, pos_off, pos_items):
    lib = load()
    n, K = topk_id.shape
    out = torch.empty((n, K + 1), dtype=torch.int32, device=topk_id.device)
    check(lib.fr_hits(ptr(topk_id), n, K, ptr(pos_off), ptr(pos_items), ptr(out), stream_ptr()), "fr_hits")
    return out


def topk_metric_sums(rec_topk):
    """float64 [4, K]: SUM over users of NDCG/Recall/Hit/MRR@1..K"""
    lib = load()
    n, K = rec_topk.shape[0], rec_topk.shape[1] - 1
    out = torch.empty((4, K), dtype=torch.float64, device=rec_topk.device)
    ws = _ws(lib.fr_topk_metrics_workspace_bytes(n, K), rec_topk.device)
    check(lib.fr_topk_metrics(ptr(rec_topk), n, K, ptr(out), ptr(ws), ws.numel(), stream_ptr()), "fr_topk_metrics")
    return out


def rec_item_stats(topk_id, n_items, is_popular=None):
    lib = load()
    n, K = topk_id.shape
    cnt = torch.empty((K, n_items), dtype=torch.int32, device=topk_id.device)
    pop = torch.empty(K, dtype=torch.int64, device=topk_id.device)
    check(lib.fr_rec_item_stats(ptr(topk_id), n, K, n_items, ptr(is_popular), ptr(cnt), ptr(pop), stream_ptr()),
          "fr_rec_item_stats")
    return cnt, pop


def gini_at_k(item_pos_count, k, n_users):
    lib = load()
    n_items = item_pos_count.shape[1]
    out = torch.empty(1, dtype=torch.float64, device=item_pos_count.device)
    ws = _ws(lib.fr_gini_workspace_bytes(n_items), item_pos_count.device)
    check(lib.fr_gini_at_k(ptr(item_pos_count), n_items, int(k), int(n_users), ptr(out), ptr(ws), ws.numel(),
                           stream_ptr()), "fr_gini_at_k")
    return out


def item_group_stats(pos_items, pos_score, group, n_items, G):
    """float64 [n_items, G, 2] = (sum score, count) per item x group over the positives"""
    lib = load()
    n_pos = pos_items.numel()
    out = torch.empty((n_items, G, 2), dtype=torch.float64, device=pos_items.device)
    ws = _ws(lib.fr_item_group_stats_workspace_bytes(n_pos, n_items, G), pos_items.device)
    check(lib.fr_item_group_stats(ptr(pos_items), ptr(pos_score), ptr(group), n_pos, n_items, G, ptr(out), ptr(ws),
                                  ws.numel(), stream_ptr()), "fr_item_group_stats")
    return out


def item_group_plan(pos_items, n_items):
    """the item-sorted view of a positive list (uint8 plan buffer): depends on the evaluation data only, build it once"""
    lib = load()
    n_pos = pos_items.numel()
    plan = torch.empty(max(int(lib.fr_item_group_plan_bytes(n_pos)), 256), dtype=torch.uint8, device=pos_items.device)
    ws = _ws(lib.fr_item_group_plan_workspace_bytes(n_pos), pos_items.device)
    check(lib.fr_item_group_plan(ptr(pos_items), n_pos, n_items, ptr(plan), plan.numel(), ptr(ws), ws.numel(), stream_ptr()),
          "fr_item_group_plan")
    return plan


def item_group_stats_planned(plan, pos_score, group, n_items, G):
    """item_group_stats() over a positive list whose plan exists: the segment reduction alone"""
    lib = load()
    n_pos = pos_score.numel()
    out = torch.empty((n_items, G, 2), dtype=torch.float64, device=pos_score.device)
    check(lib.fr_item_group_stats_planned(ptr(plan), plan.numel(), ptr(pos_score), ptr(group), n_pos, n_items, G, ptr(out),
                                          stream_ptr()), "fr_item_group_stats_planned")
    return out


def fairness_metrics(stats):
    """float64 [7]: DF, value, absolute, under, over, nonparity, J"""
    lib = load()
    n_items, G = stats.shape[0], stats.shape[1]
    out = torch.empty(7, dtype=torch.float64, device=stats.device)
    ws = _ws(lib.fr_fairness_metrics_workspace_bytes(n_items, G), stats.device)
    check(lib.fr_fairness_metrics(ptr(stats), n_items, G, ptr(out), ptr(ws), ws.numel(), stream_ptr()),
          "fr_fairness_metrics")
    return out


def unfairness_sampled(stats_all, stats_pos):
    """float64 [5]: value, absolute, under, over unfairness of the sampled-negative mode, number of items in the union"""
    lib = load()
    n_items = stats_all.shape[0]
    out = torch.empty(5, dtype=torch.float64, device=stats_all.device)
    ws = _ws(lib.fr_unfairness_sampled_workspace_bytes(n_items), stats_all.device)
    check(lib.fr_unfairness_sampled(ptr(stats_all), ptr(stats_pos), n_items, ptr(out), ptr(ws), ws.numel(), stream_ptr()),
          "fr_unfairness_sampled")
    return out
