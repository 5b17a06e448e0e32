"""FOCF (MF + fairness regulariser) -- drop-in for recbole/model/fair_recommender/focf.py:24-178.

Same plugin surface (class name, `input_type`, ctor `(config, dataset)`, `calculate_loss / predict /
full_sort_predict`, parameter names `user_embedding_layer.weight` / `item_embedding_layer.weight` so
reference checkpoints load), but every computation runs in this package's sm_100a kernels through the
C ABI.  Two ways to train:

  * compat : `loss = model.calculate_loss(interaction); loss.backward(); optimizer.step()` -- what the
    reference's own `Trainer` does (trainer.py:184-196).  `calculate_loss` is a torch.autograd.Function
    whose backward materialises the dense embedding gradients exactly like nn.Embedding does.
  * fused  : `model.train_step(interaction)` -- forward + backward + dense Adam (L2 form) in one kernel
    chain; the dense gradient never exists.  Used by `FOCFTrainer` when `learner: adam`.

There is no CPU path: the model must live on a CUDA device.
"""
import os

import torch
import torch.nn as nn

from . import _lib
from .kernels import FocfEngine, pair_scores

POINTWISE = "POINTWISE"


def _xavier_normal_initialization(module):
    # recbole/model/init.py:15-31
    if isinstance(module, nn.Embedding):
        nn.init.xavier_normal_(module.weight.data)


_NO_FAST_HOST_STEP = bool(os.environ.get("FR_FOCF_NO_FAST_HOST_STEP"))      # debugging switch: generic path for packed batches


def batch_columns(interaction, uid_f, iid_f, rating_f, sst_f, device):
    """Interaction -> the four int32/float32 device columns of include/fairrec_b200.h:fr_focf_step.
    Host batches whose four columns are views of ONE pinned buffer (pack_host_batch) move with a single H2D copy."""
    packed = getattr(interaction, "packed_host", None)
    if packed is not None:
        buf, n = packed
        dev = torch.empty(buf.numel(), dtype=torch.uint8, device=device)
        dev.copy_(buf, non_blocking=True)
        return (dev[:4 * n].view(torch.int32), dev[4 * n:8 * n].view(torch.int32), dev[8 * n:12 * n].view(torch.float32),
                dev[12 * n:16 * n].view(torch.float32), bool(getattr(interaction, "items_contiguous", False)))

    def col(name, dtype):
        t = interaction[name]
        return t.to(device=device, dtype=dtype, non_blocking=True).contiguous()
    return (col(uid_f, torch.int32), col(iid_f, torch.int32), col(rating_f, torch.float32),
            col(sst_f, torch.float32), bool(getattr(interaction, "items_contiguous", False)))


def pack_host_batch(uid, iid, rating, sst, fields, items_contiguous=True):
    """Host-side FOCF batch whose columns (int32 user ids, int32 item ids, float32 ratings, float32 attribute values)
    are views of one pinned buffer: an Interaction like any other, but `train_step` moves it with ONE H2D copy."""
    from .interaction import Interaction
    n = int(uid.numel())
    buf = torch.empty(16 * n, dtype=torch.uint8).pin_memory()
    cols = (buf[:4 * n].view(torch.int32), buf[4 * n:8 * n].view(torch.int32), buf[8 * n:12 * n].view(torch.float32),
            buf[12 * n:16 * n].view(torch.float32))
    for dst, src in zip(cols, (uid, iid, rating, sst)):
        dst.copy_(src)
    inter = Interaction(dict(zip(fields, cols)))
    inter.items_contiguous = items_contiguous
    inter.packed_host = (buf, n)
    return inter


class _FocfLoss(torch.autograd.Function):
    """focf.py:152-169 calculate_loss with autograd semantics (dense grads for both embedding tables)."""

    @staticmethod
    def forward(ctx, U, I, model, batch):
        eng = model._engine()
        loss = torch.empty(1, dtype=torch.float32, device=U.device)
        ctx.step = eng.forward(U.detach(), I.detach(), batch, model._objective, model.fair_weight, loss_out=loss)
        ctx.model, ctx.batch, ctx.tables = model, batch, (U, I)
        return loss.view(())

    @staticmethod
    def backward(ctx, grad_out):
        U, I = ctx.tables
        dU, dI = torch.empty_like(U), torch.empty_like(I)
        ctx.model._engine().backward(ctx.step, dU, dI, float(grad_out))
        return dU, dI, None, None


class FOCF(nn.Module):
    input_type = POINTWISE
    type = "GENERAL"

    def __init__(self, config, dataset):
        super().__init__()
        # abstract_recommender.py:92-104
        self.USER_ID = config["USER_ID_FIELD"]
        self.ITEM_ID = config["ITEM_ID_FIELD"]
        self.NEG_ITEM_ID = config["NEG_PREFIX"] + self.ITEM_ID
        self.n_users = dataset.num(self.USER_ID)
        self.n_items = dataset.num(self.ITEM_ID)
        self.device = config["device"]
        # focf.py:35-48
        self.embedding_size = config["embedding_size"]
        self.RATING = config["RATING_FIELD"]
        self.SST_FIELD = config["sst_attr_list"][0]
        self.fair_weight = float(config["fair_weight"])
        self.max_rating = float(dataset.inter_feat[self.RATING].max())
        self.fair_objective = config["fair_objective"].strip().lower()
        if self.fair_objective not in _lib.OBJECTIVES:
            raise ValueError("you must set config['fair_objective'] be one of (none,"
                             "value,absolute,under,over,nonparity)")
        self._objective = _lib.OBJECTIVES[self.fair_objective]
        if self.embedding_size % 4 != 0:
            raise ValueError("fairrec_b200 FOCF needs embedding_size to be a multiple of 4")
        self.user_embedding_layer = nn.Embedding(self.n_users, self.embedding_size)
        self.item_embedding_layer = nn.Embedding(self.n_items, self.embedding_size)
        self.apply(_xavier_normal_initialization)
        self._eng = None
        self._adam = None

    # ------------------------------------------------------------------ plumbing
    def _engine(self):
        U = self.user_embedding_layer.weight
        if not U.is_cuda:
            raise _lib.FairRecLibraryError("FOCF (fairrec_b200) runs on CUDA only: move the model to a cuda device")
        if self._eng is None or self._eng.device != U.device:
            self._eng = FocfEngine(self.n_users, self.n_items, self.embedding_size, 4096, U.device)
        return self._eng

    def _batch(self, interaction):
        return batch_columns(interaction, self.USER_ID, self.ITEM_ID, self.RATING, self.SST_FIELD,
                             self.user_embedding_layer.weight.device)

    def other_parameter(self):
        return dict()

    def load_other_parameter(self, para):
        if para:
            for k, v in para.items():
                setattr(self, k, v)

    # ------------------------------------------------------------------ reference API
    def forward(self, user, item):
        """focf.py:136-143; returns (pred, user_embedding, item_embedding)"""
        U, I = self.user_embedding_layer.weight, self.item_embedding_layer.weight
        pred = pair_scores(U.detach(), I.detach(), user.to(torch.int32).contiguous(),
                           item.to(torch.int32).contiguous())
        return pred, U[user.long()], I[item.long()]

    def predict(self, interaction):
        """focf.py:145-150"""
        U, I = self.user_embedding_layer.weight, self.item_embedding_layer.weight
        dev = U.device
        user = interaction[self.USER_ID].to(device=dev, dtype=torch.int32).contiguous()
        item = interaction[self.ITEM_ID].to(device=dev, dtype=torch.int32).contiguous()
        return pair_scores(U.detach(), I.detach(), user, item, _lib.TRANSFORM_CLAMP_DIV, self.max_rating)

    def calculate_loss(self, interaction):
        """focf.py:152-169; 0-dim tensor connected to both embedding tables"""
        return _FocfLoss.apply(self.user_embedding_layer.weight, self.item_embedding_layer.weight, self,
                               self._batch(interaction))

    def full_sort_predict(self, interaction):
        """focf.py:171-178: dense [b * n_items] scores (compat API; the fused evaluator never calls this)"""
        U, I = self.user_embedding_layer.weight, self.item_embedding_layer.weight
        user = interaction[self.USER_ID].to(device=U.device, dtype=torch.int32)
        uid = user.repeat_interleave(self.n_items).contiguous()
        iid = torch.arange(self.n_items, dtype=torch.int32, device=U.device).repeat(user.numel()).contiguous()
        return pair_scores(U.detach(), I.detach(), uid, iid, _lib.TRANSFORM_CLAMP_DIV, self.max_rating)

    # ------------------------------------------------------------------ fused training step
    def init_adam(self, lr=1e-3, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8, mode="dense_exact", max_steps=1 << 20):
        """state of torch.optim.Adam(params, lr, weight_decay) as built by trainer.py:139.

        mode "dense_exact": every row of both tables moves at every step, as in the reference (rows without a data gradient
        still see weight_decay * p and the decay of their moments).  mode "lazy_exact": the same numbers bit for bit, but a
        row is only brought up to date when a batch touches it or at `flush_adam()` -- by replaying the steps it missed in
        registers -- so a step streams the touched rows instead of all of them (include/fairrec_b200.h: enum fr_adam_mode).
        Everything that READS the tables outside train_step (predict, full_sort_predict, evaluation, state_dict) must see
        flushed tables: the trainer calls flush_adam() before evaluating / checkpointing."""
        if mode not in ("dense_exact", "lazy_exact"):
            raise ValueError("adam mode must be dense_exact or lazy_exact")
        U, I = self.user_embedding_layer.weight, self.item_embedding_layer.weight
        self._adam = dict(mU=torch.zeros_like(U), vU=torch.zeros_like(U), mI=torch.zeros_like(I),
                          vI=torch.zeros_like(I), step=0, lr=lr, beta1=betas[0], beta2=betas[1], eps=eps,
                          weight_decay=weight_decay, mode=mode)
        if mode == "lazy_exact":
            import ctypes
            self._adam.update(last_u=torch.zeros(U.shape[0], dtype=torch.int32, device=U.device),
                              last_i=torch.zeros(I.shape[0], dtype=torch.int32, device=U.device),
                              scalars=torch.zeros(2 * int(max_steps), dtype=torch.float32, device=U.device),
                              filled=ctypes.c_int32(0))
        return self._adam

    def _moment_key(self):
        """addresses of the Adam moments: part of every captured graph's key (init_adam() allocates new ones)"""
        a = self._adam
        return (a["mU"].data_ptr(), a["vU"].data_ptr(), a["mI"].data_ptr(), a["vI"].data_ptr())

    def flush_adam(self):
        """lazy_exact: apply the pending updates of every row (no-op otherwise).  After it the tables, moments and step
        count are exactly those of dense_exact training."""
        if self._adam is not None:
            self._engine().adam_flush(self.user_embedding_layer.weight.data, self.item_embedding_layer.weight.data, self._adam)

    def train_step(self, interaction, loss_out=None):
        """One fused optimisation step (trainer.py:183-196 for `learner: adam`).  Returns the device tensor
        holding the loss; nothing synchronises."""
        adam = self._adam
        if adam is None:
            raise RuntimeError("call init_adam() (FOCFTrainer does) before train_step()")
        packed = getattr(interaction, "packed_host", None)
        if packed is None or _NO_FAST_HOST_STEP or adam.get("mode") == "lazy_exact":
            return self._train_step_generic(interaction, loss_out)
        # host batch in one pinned buffer: persistent staging buffer + persistent argument struct (kernels.py); nothing
        # here is seen by autograd (`.data` tensors, one copy, one library call), so no grad-mode bookkeeping either
        eng = self._engine()
        adam["step"] += 1
        out = eng.loss if loss_out is None else loss_out
        mods = self._modules          # plain dict lookups instead of nn.Module.__getattr__
        eng.train_step_packed(mods["user_embedding_layer"]._parameters["weight"].data,
                              mods["item_embedding_layer"]._parameters["weight"].data, adam, packed,
                              bool(getattr(interaction, "items_contiguous", False)), self._objective, self.fair_weight, out)
        return out

    def train_steps_host(self, interactions):
        """One fused optimisation step per HOST batch of `interactions` (pack_host_batch Interactions), driven by ONE call
        into the library (include/fairrec_b200.h:fr_focf_train_steps_host): per step the batch is copied host -> device,
        stepped, and its loss copied device -> host; the loop of trainer.py:181-196 runs in C instead of the interpreter.
        Returns the per-step losses as a pinned host tensor, complete on return."""
        adam = self._adam
        if adam is None:
            raise RuntimeError("call init_adam() (FOCFTrainer does) before train_steps_host()")
        packed = [getattr(it, "packed_host", None) for it in interactions]
        if any(p is None for p in packed):
            raise ValueError("train_steps_host takes host batches built by pack_host_batch")
        contiguous = all(bool(getattr(it, "items_contiguous", False)) for it in interactions)
        mods = self._modules
        losses = self._engine().train_steps_host(mods["user_embedding_layer"]._parameters["weight"].data,
                                                 mods["item_embedding_layer"]._parameters["weight"].data, adam, packed,
                                                 contiguous, self._objective, self.fair_weight)
        adam["step"] += len(packed)
        return losses

    @torch.no_grad()
    def _train_step_generic(self, interaction, loss_out=None):
        eng = self._engine()
        self._adam["step"] += 1
        out = eng.loss if loss_out is None else loss_out
        eng.train_step(self.user_embedding_layer.weight.data, self.item_embedding_layer.weight.data, self._adam,
                       self._batch(interaction), self._objective, self.fair_weight, loss_out=out)
        return out

    @torch.no_grad()
    def dp_train_step(self, interaction, norm, group, loss_out=None):
        """Data-parallel step: this rank's whole-item batch, global normalisers `norm = (B_total,
        J_total)`, dense gradient shares summed with ONE NCCL all-reduce, then the same dense Adam on every replica.
        Equals the fused single-GPU step on the union of the ranks' batches (up to summation order)."""
        import torch.distributed as dist
        if self._adam is None:
            raise RuntimeError("call init_adam() before dp_train_step()")
        eng = self._engine()
        U, I = self.user_embedding_layer.weight.data, self.item_embedding_layer.weight.data
        if getattr(self, "_dp_grad", None) is None:
            self._dp_grad = torch.empty(U.numel() + I.numel(), dtype=torch.float32, device=U.device)
        dU = self._dp_grad[:U.numel()].view_as(U)
        dI = self._dp_grad[U.numel():].view_as(I)
        out = eng.loss if loss_out is None else loss_out
        st = eng.forward(U, I, self._batch(interaction), self._objective, self.fair_weight, loss_out=out, norm=norm)
        eng.backward(st, dU, dI, 1.0)
        dist.all_reduce(self._dp_grad, group=group)
        self._adam["step"] += 1
        eng.adam_dense(U, I, dU, dI, self._adam)
        return out

    @torch.no_grad()
    def epoch_runner(self, loader, loss_buf, n_slots=6, plan_buf=0):
        """Plan an epoch of `loader` on the device and return a runner whose `.run(k)` executes the next k steps as ONE
        persistent cooperative launch (fr_focf_epoch_run: producer CTAs build the batches ahead, compute CTAs keep their
        share of the tables and Adam moments in shared memory; bit-identical to the stepwise path), or None when the
        shape is not eligible (tables too large for shared-memory residency, batches over 8192 rows, d > 128)."""
        import ctypes
        if self._adam is None:
            raise RuntimeError("call init_adam() before epoch_runner()")
        if self._adam.get("mode", "dense_exact") != "dense_exact":
            return None
        eng = self._engine()
        U, I = self.user_embedding_layer.weight.data, self.item_embedding_layer.weight.data
        plan = loader.plan_epoch_device(plan_buf)
        if loss_buf.numel() < plan["len"]:
            raise ValueError("loss buffer shorter than the epoch")
        cap = plan["cols"][0].numel()
        if cap > 8192 or self.embedding_size > 128:
            return None
        st = getattr(self, "_ep_state", None)
        if st is None or st["device"] != U.device or st["cap"] != cap or st["n_slots"] != n_slots:
            engines = [FocfEngine(self.n_users, self.n_items, self.embedding_size, cap, U.device) for _ in range(n_slots)]
            for e in engines:
                e.flags = eng.flags                            # one status word for every slot
            cols = [tuple(torch.empty_like(c) for c in plan["cols"]) for _ in range(n_slots)]
            st = self._ep_state = dict(device=U.device, cap=cap, n_slots=n_slots, engines=engines, cols=cols,
                                       sync=torch.zeros(16, dtype=torch.int32, device=U.device), per_buf={})
        # the argument structs point into the plan buffers and the loss buffer: one set per (plan buffer set)
        pb = st["per_buf"].setdefault(plan_buf, dict(key=None))
        key = (plan["generation"], loss_buf.data_ptr(), U.data_ptr(), I.data_ptr(), plan["len"], id(self._adam["mU"]))
        if pb["key"] != key:
            steps = (_lib.FocfStep * n_slots)()
            for k, (e, c) in enumerate(zip(st["engines"], st["cols"])):
                s = e.planned_step(U, I, self._adam, dict(plan, cols=c), loader.train, self._objective, self.fair_weight,
                                   loss_buf)
                ctypes.memmove(ctypes.byref(steps[k]), ctypes.byref(s), ctypes.sizeof(s))
            pb["steps"], pb["key"] = steps, key
            pb["eligible"] = bool(eng.lib.fr_focf_epoch_eligible(ctypes.cast(steps, ctypes.c_void_p), n_slots))
        if not pb["eligible"]:
            return None
        return _EpochRunner(self, plan, st, pb["steps"])

    @torch.no_grad()
    def train_epochs_planned(self, loader, n_epochs):
        """`n_epochs` epochs of trainer.py:181-196 with the host's share overlapped: while epoch e runs on the device (one
        persistent launch) the host draws epoch e + 1 (two plan buffer sets and two loss buffers alternate) and reads the
        losses of epoch e - 1.  Returns (list of per-epoch loss sums, total steps, total interactions); raises what the
        reference raises for NaN losses.  Falls back to train_epoch_planned per epoch where the epoch kernel is not eligible."""
        dev = self.user_embedding_layer.weight.device
        n_max = len(loader) + 8
        if getattr(self, "_ep_loss", None) is None or self._ep_loss[0].numel() < n_max:
            self._ep_loss = [torch.zeros(n_max, device=dev) for _ in range(2)]
            self._ep_loss_host = [torch.zeros(n_max).pin_memory() for _ in range(2)]
            self._ep_done = [torch.cuda.Event() for _ in range(2)]
        sums, steps, rows, pending = [], 0, 0, None

        def consume(b, n):
            self._ep_done[b].synchronize()
            lh = self._ep_loss_host[b][:n]
            if bool(torch.isnan(lh).any()):
                raise ValueError("Training loss is nan")          # trainer.py:286-288
            sums.append(float(lh.double().sum()))

        for e in range(n_epochs):
            b = e & 1
            runner = self.epoch_runner(loader, self._ep_loss[b], plan_buf=b)
            if runner is None:
                if pending is not None:
                    consume(*pending)
                    pending = None
                n, r = self.train_epoch_planned(loader, self._ep_loss[b], persistent=False)
                sums.append(float(self._ep_loss[b][:n].double().sum()))
            else:
                n = runner.plan["len"]
                r = runner.run(n)
                self._ep_loss_host[b][:n].copy_(self._ep_loss[b][:n], non_blocking=True)
                self._ep_done[b].record()
                if pending is not None:
                    consume(*pending)
                pending = (b, n)
            steps += n
            rows += r
        if pending is not None:
            consume(*pending)
        return sums, steps, rows

    @torch.no_grad()
    def planned_runner(self, loader, loss_buf, graph_steps=8, persistent=None):
        """Plan an epoch of `loader` on the device and return a runner whose `.run(k)` executes the next k fused steps
        with NO per-step host work.  persistent (default: on unless FR_FOCF_NO_EPOCH_KERNEL=1): where the shape allows it
        the runner is `epoch_runner`'s -- k steps = one persistent launch; otherwise (and below) the stepwise path:
        the step is captured once into CUDA graphs and replayed; batch size, batch cursor
        and Adam step count are device resident.  loss_buf[cursor] receives each step's loss.

        Two workspaces alternate the batches of the plan (even batches -> workspace 0, odd -> workspace 1): the
        preparation of batch t+1 (gather + sort / segments / row stamps, independent of the tables) runs on a second
        stream while batch t computes (forward + loss + gradients + Adam), so the captured graph of `graph_steps` steps
        has a critical path of one compute per step instead of prepare + compute."""
        if self._adam is None:
            raise RuntimeError("call init_adam() before planned_runner()")
        if persistent is None:
            persistent = os.environ.get("FR_FOCF_NO_EPOCH_KERNEL", "0") != "1"
        if persistent:
            runner = self.epoch_runner(loader, loss_buf)
            if runner is not None:
                return runner
        eng = self._engine()
        U, I = self.user_embedding_layer.weight.data, self.item_embedding_layer.weight.data
        if getattr(self, "_eng2", None) is None or self._eng2.device != U.device:
            self._eng2 = FocfEngine(self.n_users, self.n_items, self.embedding_size, 4096, U.device)
            self._eng2.flags = eng.flags                      # one status word for both
            self._cols2 = None
        graph_steps = max(2, graph_steps + (graph_steps & 1))   # even: the pipelined graph starts on workspace 0
        plan = loader.plan_epoch_device()
        if loss_buf.numel() < plan["len"]:
            raise ValueError("loss buffer shorter than the epoch")
        cap = plan["cols"][0].numel()
        if self._cols2 is None or self._cols2[0].numel() != cap:
            self._cols2 = tuple(torch.empty_like(c) for c in plan["cols"])
        engines = (eng, self._eng2)
        key = (plan["generation"], loss_buf.data_ptr(), U.data_ptr(), I.data_ptr(), graph_steps, plan["len"],
               eng.ws.data_ptr(), self._eng2.ws.data_ptr(), cap) + self._moment_key()
        runner = _PlannedRunner(self, plan)
        T = self._adam["step"]
        if getattr(self, "_graph_key", None) != key:
            plans = (plan, dict(plan, cols=self._cols2))
            sts = [e.planned_step(U, I, self._adam, p, loader.train, self._objective, self.fair_weight, loss_buf)
                   for e, p in zip(engines, plans)]
            key = key[:6] + (eng.ws.data_ptr(), self._eng2.ws.data_ptr(), cap) + self._moment_key()   # planned_step may have grown a workspace
            # eager first step on each workspace: module loading, function attributes
            # (an epoch of ONE batch warms up workspace 0 only: cursor 1 would wrap to batch 0 and train it twice)
            eng.set_counters(plan_cursor=0, adam_step=T, stride=1)
            eng.run_planned(sts[0])
            warm = min(2, plan["len"])
            if warm == 2:
                self._eng2.set_counters(plan_cursor=1, adam_step=T + 1, stride=1)
                self._eng2.run_planned(sts[1])
            runner.cursor = warm
            T += warm
            self._adam["step"] = T
            torch.cuda.synchronize()
            self._graphs = {}
            side = getattr(self, "_prep_stream", None) or torch.cuda.Stream(device=U.device)
            self._prep_stream = side
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):   # capture does not execute
                main = torch.cuda.current_stream()
                done = [None, None]
                side.wait_stream(main)
                for t in range(graph_steps):
                    k = t & 1
                    with torch.cuda.stream(side):
                        if done[k] is not None:
                            side.wait_event(done[k])          # workspace k is free once its previous compute finished
                        engines[k].run_prepare(sts[k])
                        ready = torch.cuda.Event()
                        ready.record(side)
                    main.wait_event(ready)
                    engines[k].run_compute(sts[k])
                    done[k] = torch.cuda.Event()
                    done[k].record(main)
                main.wait_stream(side)
            self._graphs[graph_steps] = g
            for k in (0, 1):            # single steps (tail of an epoch): prepare + compute back to back
                g1 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g1):
                    engines[k].run_planned(sts[k])
                self._graphs[("one", k)] = g1
            self._graph_key, self._graph_step, self._graph_len = key, sts[0], graph_steps
        # workspace k serves the batches k, k+2, ...: cursor and Adam count advance by 2 per use
        c = runner.cursor
        k0 = c & 1                      # workspace of the next batch
        for k, e in enumerate(engines):
            first = c + ((k - k0) & 1)  # first batch this workspace will see
            e.set_counters(plan_cursor=first, adam_step=T + (first - c) - 1, stride=2)
        return runner

    @torch.no_grad()
    def dp_planned_runner(self, loader, loss_buf, group, graph_steps=8):
        """Data-parallel twin of planned_runner: this rank's planned whole-item batches, the global normalisers of
        every planned batch in a device array (summed over the ranks once per epoch), and the step
        forward -> backward (dense gradient shares) -> ONE NCCL all-reduce -> dense Adam captured into a CUDA graph of
        `graph_steps` steps.  loss_buf[cursor] receives this rank's loss share of each step."""
        import torch.distributed as dist
        if self._adam is None:
            raise RuntimeError("call init_adam() before dp_planned_runner()")
        eng = self._engine()
        U, I = self.user_embedding_layer.weight.data, self.item_embedding_layer.weight.data
        dev = U.device
        plan = loader.plan_epoch_device()
        n = torch.tensor([plan["len"]], dtype=torch.int64, device=dev)
        dist.all_reduce(n, op=dist.ReduceOp.MIN, group=group)      # ranks may have planned different batch counts
        L = int(n.item())
        plan = dict(plan, len=L, rows=int(sum(plan["batch_rows"][:L])))
        if loss_buf.numel() < L:
            raise ValueError("loss buffer shorter than the epoch")
        if getattr(self, "_dp_norm", None) is None or self._dp_norm.shape[0] < L:
            self._dp_norm = torch.zeros((int(L * 1.5) + 64, 2), dtype=torch.int32, device=dev)
        norms = plan["desc"][:L][:, [3, 2]].contiguous()            # (B, J) of this rank's batches
        dist.all_reduce(norms, group=group)
        self._dp_norm[:L].copy_(norms)
        if getattr(self, "_dp_grad", None) is None:
            self._dp_grad = torch.empty(U.numel() + I.numel(), dtype=torch.float32, device=dev)
        dU, dI = self._dp_grad[:U.numel()].view_as(U), self._dp_grad[U.numel():].view_as(I)
        key = ("dp", plan["generation"], loss_buf.data_ptr(), U.data_ptr(), I.data_ptr(), graph_steps, L,
               self._dp_norm.data_ptr())
        runner = _DpPlannedRunner(self, plan, group)
        if getattr(self, "_dp_graph_key", None) != key or eng.ws.data_ptr() != getattr(self, "_dp_graph_ws", None):
            st = eng.planned_step(U, I, self._adam, plan, loader.train, self._objective, self.fair_weight, loss_buf)
            st.norm_dev = self._dp_norm.data_ptr()
            st.dU, st.dI = dU.data_ptr(), dI.data_ptr()
            self._dp_step = st
            eng.set_counters(plan_cursor=0, adam_step=self._adam["step"], stride=1)
            runner.eager_steps(1)                                   # NCCL communicator, module loading
            torch.cuda.synchronize()
            self._dp_graphs = {}
            for g_steps in sorted({graph_steps, 1}):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for _ in range(g_steps):
                        runner._enqueue()
                self._dp_graphs[g_steps] = g
            self._dp_graph_key, self._dp_graph_len, self._dp_graph_ws = key, graph_steps, eng.ws.data_ptr()
        eng.set_counters(plan_cursor=runner.cursor, adam_step=self._adam["step"], stride=1)
        return runner

    def release_graphs(self):
        """drop every captured CUDA graph (call before torch.distributed.destroy_process_group(): graphs that captured
        NCCL kernels keep the communicator busy and make the teardown hang)"""
        import gc
        self._dp_graphs, self._dp_graph_key = {}, None
        self._graphs, self._graph_key = {}, None
        gc.collect()
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    @torch.no_grad()
    def train_epoch_planned(self, loader, loss_buf, graph_steps=8, persistent=None):
        """One epoch through planned_runner.  Returns (number of steps, number of interactions)."""
        runner = self.planned_runner(loader, loss_buf, graph_steps, persistent)
        n = runner.plan["len"]
        runner.run(n - runner.cursor)
        return n, runner.plan["rows"]

    def check_flags(self):
        """Raise what the reference would have raised for a faulty batch (synchronises the device)."""
        f = self._engine().read_flags()
        if f & _lib.FLAG_TOO_MANY_GROUPS:
            raise IndexError("index 2 is out of bounds for dimension 1 with size 2 "
                             "(more than two sensitive-attribute values in a batch, focf.py:86)")
        if f & _lib.FLAG_SINGLE_GROUP:
            raise IndexError("index 1 is out of bounds for dimension 0 with size 1 (focf.py:130)")
        if f & _lib.FLAG_NAN_LOSS:
            raise ValueError("Training loss is nan")  # trainer.py:286-288


class _DpPlannedRunner:
    def __init__(self, model, plan, group):
        self.model, self.plan, self.group, self.cursor = model, plan, group, 0

    def _enqueue(self):
        import ctypes

        import torch.distributed as dist
        m = self.model
        eng, st = m._engine(), m._dp_step
        _lib.check(eng.lib.fr_focf_forward(ctypes.byref(st), _lib.stream_ptr()), "fr_focf_forward")
        _lib.check(eng.lib.fr_focf_backward(ctypes.byref(st), 1.0, _lib.stream_ptr()), "fr_focf_backward")
        dist.all_reduce(m._dp_grad, group=self.group)
        _lib.check(eng.lib.fr_focf_adam(ctypes.byref(st), _lib.stream_ptr()), "fr_focf_adam")

    def _account(self, k):
        n = self.plan["len"]
        rows = sum(self.plan["batch_rows"][(self.cursor + i) % n] for i in range(k))
        self.cursor += k
        self.model._adam["step"] += k
        return rows

    def eager_steps(self, k):
        for _ in range(k):
            self._enqueue()
        return self._account(k)

    def run(self, k):
        m = self.model
        G = m._dp_graph_len
        for _ in range(k // G):
            m._dp_graphs[G].replay()
        for _ in range(k % G):
            m._dp_graphs[1].replay()
        return self._account(k)


class _EpochRunner:
    """k planned steps = ONE launch of the persistent epoch kernel (fr_focf_epoch_run)"""

    def __init__(self, model, plan, state, steps):
        self.model, self.plan, self.state, self.steps, self.cursor = model, plan, state, steps, 0

    def run(self, k):
        import ctypes
        m, st = self.model, self.state
        k = max(int(k), 0)
        if k == 0:
            return 0
        _lib.check(m._engine().lib.fr_focf_epoch_run(ctypes.cast(self.steps, ctypes.c_void_p), st["n_slots"], self.cursor,
                                                     k, m._adam["step"] + 1, _lib.ptr(st["sync"]), _lib.stream_ptr()),
                   "fr_focf_epoch_run")
        n = self.plan["len"]
        rows = sum(self.plan["batch_rows"][(self.cursor + i) % n] for i in range(k))
        self.cursor += k
        m._adam["step"] += k
        return rows

    eager_steps = run


class _PlannedRunner:
    def __init__(self, model, plan):
        self.model, self.plan, self.cursor = model, plan, 0

    def eager_steps(self, k):
        """the next k steps launched one by one on workspace 0 (profiling: graph replays bypass the launch hooks)"""
        m = self.model
        eng = m._engine()
        eng.set_counters(plan_cursor=self.cursor, adam_step=m._adam["step"], stride=1)
        for _ in range(k):
            eng.run_planned(m._graph_step)
        n = self.plan["len"]
        rows = sum(self.plan["batch_rows"][(self.cursor + i) % n] for i in range(k))
        self.cursor += k
        m._adam["step"] += k
        return rows

    def run(self, k):
        """execute the next k planned steps (graph replays only); returns the interactions they cover"""
        m = self.model
        G = m._graph_len
        big = m._graphs[G]
        k = max(int(k), 0)
        left, c = k, self.cursor
        while left > 0:
            if (c & 1) == 0 and left >= G:
                big.replay()
                c += G
                left -= G
            else:
                m._graphs[("one", c & 1)].replay()
                c += 1
                left -= 1
        n = self.plan["len"]
        rows = sum(self.plan["batch_rows"][(self.cursor + i) % n] for i in range(k))
        self.cursor += k
        m._adam["step"] += k
        return rows
