"""Row-sharded FOCF training over NVLink peer memory -- host side of include/fairrec_b200.h:fr_focf_shard_step.

Replaces, for tables that are distributed over the GPUs of one box, the loop body of `Trainer._train_epoch`
(recbole/trainer/trainer.py:181-196) together with `FOCFDataLoader._next_batch_data`
(recbole/data/dataloader/focf_dataloader.py:37-50).  One process per GPU; rank r owns rows r, r+P, ... of the user table, the
item table and the Adam moments and holds the train interactions of its own users.  The batch is the reference's (all train
rows of the drawn items, the same draw list on every rank); see csrc/focf_shard.cu for what crosses NVLink.

`torch.distributed` is plumbing only: it carries the 64-byte CUDA IPC handles of the exchange buffers at start-up (and the
row gather for evaluation / checkpoints).  No collective runs inside a training step -- the step's exchanges are stores into
peer memory issued by the step's own kernels.

`ShardedGroupEmu` runs P ranks inside ONE process on ONE device (the "peers" are then plain device pointers and the host
sequences the phases instead of the cross-GPU barriers): this is how the single-GPU test box checks sharded == unsharded.
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib
from ._lib import FocfShardStep, check, load, ptr, stream_ptr


def local_rows(n, rank, world):
    """number of rows r, r+world, ... below n"""
    return (n - rank + world - 1) // world if n > rank else 0


class ShardedTrainData:
    """This rank's slice of the train split: the interactions of the users it owns (uid % world == rank), sorted by GLOBAL
    item id (stable, like focf_dataloader.py:12), user ids stored as LOCAL rows (uid // world)."""

    def __init__(self, uid, iid, rating, sst_of_user, n_users, n_items, rank, world, device, global_item_count=None,
                 fields=("user_id", "item_id", "rating", "gender")):
        self.rank, self.world, self.device = int(rank), int(world), device
        self.n_users, self.n_items = int(n_users), int(n_items)
        self.n_users_loc, self.n_items_loc = local_rows(self.n_users, rank, world), local_rows(self.n_items, rank, world)
        self.fields = fields
        on_dev = torch.is_tensor(uid) and uid.is_cuda
        if on_dev:
            iid = iid.long()
            if global_item_count is None:
                global_item_count = torch.bincount(iid, minlength=self.n_items).cpu().numpy()
            mine = (uid % world) == rank
            lu, li, lr = (uid[mine] // world).to(torch.int32), iid[mine], rating[mine]
            order = torch.sort(li, stable=True).indices
            self.train_uid = lu[order].contiguous()
            self.train_rating = lr[order].to(torch.float32).contiguous()
            local_count = torch.bincount(li, minlength=self.n_items).cpu().numpy().astype(np.int64)
            sst = sst_of_user.to(device=device, dtype=torch.float32)
            self.sst_of_user = sst[rank::world].contiguous()
        else:
            uid, iid, rating = np.asarray(uid), np.asarray(iid), np.asarray(rating)
            if global_item_count is None:
                global_item_count = np.bincount(iid, minlength=self.n_items)
            mine = (uid % world) == rank
            lu, li, lr = uid[mine] // world, iid[mine], rating[mine]
            order = np.argsort(li, kind="stable")
            self.train_uid = torch.as_tensor(lu[order].astype(np.int32)).to(device)
            self.train_rating = torch.as_tensor(lr[order].astype(np.float32)).to(device)
            local_count = np.bincount(li, minlength=self.n_items).astype(np.int64)
            self.sst_of_user = torch.as_tensor(np.asarray(sst_of_user, dtype=np.float32)[rank::world].copy()).to(device)
        self.item_count_h = np.asarray(global_item_count, dtype=np.int64)      # GLOBAL counts: batch boundaries
        self.local_count_h = local_count                                        # this rank's rows per item
        off = np.zeros(self.n_items + 1, np.int64)
        off[1:] = np.cumsum(local_count)
        if off[-1] >= 2 ** 31:
            raise ValueError("the device-side batch builder indexes the train split with int32 offsets")
        self.item_off = torch.as_tensor(off.astype(np.int32)).to(device)
        self.n_rows_loc = int(off[-1])
        self.n_rows = int(self.item_count_h.sum())
        self.item_uniques = np.nonzero(self.item_count_h)[0]
        self.max_rating = 5.0


class ShardedFOCFLoader:
    """The reference's batch draws (focf_dataloader.py:37-50, "fast" mode of dataloader.FOCFDataLoader: a fresh uniform
    permutation of the items present in train per batch, items taken until the batch holds >= train_batch_size rows) made
    from the GLOBAL item counts with the same seed on every rank, so all ranks draw the same list; each rank derives its own
    row offsets from its local counts."""

    def __init__(self, train_batch_size, data, seed):
        self.step, self.data = int(train_batch_size), data
        self._rng = np.random.default_rng(seed)
        mean = max(1.0, data.n_rows / max(1, len(data.item_uniques)))
        self.J_cap = int(min(len(data.item_uniques), math.ceil(4.0 * self.step / mean) + 1024))
        self.max_batch_loc = 0

    def __len__(self):
        return math.ceil(self.data.n_rows / self.step)

    def plan(self, n_batches):
        """draws for the next n_batches; returns a list of dicts (host) and uploads the index arrays with one copy each"""
        d = self.data
        items, offs, slots, desc = [], [], [], []
        pi = po = 0
        for _ in range(int(n_batches)):
            perm = self._rng.permutation(d.item_uniques)
            csum = np.cumsum(d.item_count_h[perm])
            J = min(int(np.searchsorted(csum, self.step, side="left")) + 1, len(perm))
            it = perm[:J].astype(np.int64)
            if J > self.J_cap:
                raise ValueError(f"a batch drew {J} items, more than the exchange buffers were sized for ({self.J_cap})")
            off = np.zeros(J + 1, np.int64)
            off[1:] = np.cumsum(d.local_count_h[it])
            owner = it % d.world
            slot = np.zeros(J, np.int64)
            for o in range(d.world):
                m = owner == o
                slot[m] = np.arange(int(m.sum()))
            desc.append(dict(items_pos=pi, offs_pos=po, J=J, B_loc=int(off[-1]), B_glob=int(csum[J - 1])))
            items.append(it); offs.append(off); slots.append(slot)
            pi += J
            po += J + 1
            self.max_batch_loc = max(self.max_batch_loc, int(off[-1]))
        dev = d.device
        def up(a):
            t = torch.from_numpy(np.concatenate(a).astype(np.int32))
            return t.pin_memory().to(dev, non_blocking=True) if torch.device(dev).type == "cuda" else t

        return dict(desc=desc, items=up(items), offs=up(offs), slots=up(slots),
                    rows_glob=sum(b["B_glob"] for b in desc))


class ShardedFOCF:
    """One rank of the row-sharded FOCF trainer: local table shards + moments, workspace, exchange memory, step struct."""

    def __init__(self, data, d, objective="value", fair_weight=1.0, lr=1e-3, weight_decay=1e-3, betas=(0.9, 0.999),
                 eps=1e-8, adam_mode="dense_exact", J_cap=1024, max_batch_loc=1 << 16, max_steps=1 << 16):
        self.lib = load()
        self.data, self.d = data, int(d)
        self.rank, self.world, dev = data.rank, data.world, data.device
        if self.world > _lib.MAX_RANKS:
            raise ValueError(f"at most {_lib.MAX_RANKS} ranks")
        if objective == "nonparity":
            raise ValueError("the nonparity objective needs batch-global group means and is not available row-sharded")
        self.objective, self.fair_weight = _lib.OBJECTIVES[objective], float(fair_weight)
        self.adam = dict(lr=lr, beta1=betas[0], beta2=betas[1], eps=eps, weight_decay=weight_decay, step=0)
        self.adam_mode = {"dense_exact": _lib.ADAM_DENSE_EXACT, "lazy_exact": _lib.ADAM_LAZY_EXACT}[adam_mode]
        nu, ni = data.n_users_loc, data.n_items_loc
        z = lambda n: torch.zeros((n, self.d), dtype=torch.float32, device=dev)
        self.U, self.I = z(nu), z(ni)
        self.mU, self.vU, self.mI, self.vI = z(nu), z(nu), z(ni), z(ni)
        self.flags = torch.zeros(1, dtype=torch.int32, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.J_cap = int(J_cap)
        self._alloc_batch(max_batch_loc)
        if self.adam_mode == _lib.ADAM_LAZY_EXACT:
            self.last_u = torch.zeros(nu, dtype=torch.int32, device=dev)
            self.last_i = torch.zeros(ni, dtype=torch.int32, device=dev)
            self.scalars = torch.zeros(2 * int(max_steps), dtype=torch.float32, device=dev)
            self._filled = ctypes.c_int32(0)
        self.xchg_bytes = self.lib.fr_focf_shard_xchg_bytes(self.world, self.J_cap, self.d)
        own = ctypes.c_void_p()
        check(self.lib.fr_xchg_alloc(self.xchg_bytes, ctypes.byref(own)), "fr_xchg_alloc")
        self.xchg_own = own.value
        self.peers = [None] * self.world
        self.peers[self.rank] = self.xchg_own
        self._opened = []
        self.barriers = 0
        self._s = None

    # ------------------------------------------------------------------ buffers
    def _alloc_batch(self, B):
        dev = self.data.device
        B = int(B * 1.1) + 64
        self.cap = B
        self.cols = (torch.empty(B, dtype=torch.int32, device=dev), torch.empty(B, dtype=torch.int32, device=dev),
                     torch.empty(B, dtype=torch.float32, device=dev), torch.empty(B, dtype=torch.float32, device=dev))
        self.pred = torch.empty(B, dtype=torch.float32, device=dev)
        n = self.lib.fr_focf_shard_workspace_bytes(self.data.n_users_loc, self.data.n_items_loc, self.d, B, self.J_cap)
        self.ws = torch.empty(n, dtype=torch.uint8, device=dev)
        check(self.lib.fr_focf_shard_workspace_init(ptr(self.ws), n, self.data.n_users_loc, self.data.n_items_loc, self.d, B,
                                                    self.J_cap, stream_ptr()), "fr_focf_shard_workspace_init")

    def set_tables(self, U_full, I_full):
        """load this rank's rows of full tables (tests, checkpoints)"""
        self.U.copy_(torch.as_tensor(U_full)[self.rank::self.world].to(self.U.device))
        self.I.copy_(torch.as_tensor(I_full)[self.rank::self.world].to(self.I.device))

    def init_xavier(self, n_users, n_items, seed):
        """recbole/model/init.py:15-31 xavier_normal_ for the local rows: std = sqrt(2 / (rows + d)) of the FULL table"""
        g = torch.Generator(device=self.U.device).manual_seed(int(seed) * 1009 + self.rank)
        self.U.normal_(0.0, math.sqrt(2.0 / (n_users + self.d)), generator=g)
        self.I.normal_(0.0, math.sqrt(2.0 / (n_items + self.d)), generator=g)

    # ------------------------------------------------------------------ exchange set-up
    def connect(self, group=None):
        """exchange the CUDA IPC handles of the exchange buffers over torch.distributed and map every peer's buffer"""
        import torch.distributed as dist
        h = ctypes.create_string_buffer(64)
        check(self.lib.fr_xchg_export(self.xchg_own, h), "fr_xchg_export")
        mine = torch.tensor(list(h.raw), dtype=torch.uint8, device=self.U.device)
        allh = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allh, mine, group=group)
        for k in range(self.world):
            if k == self.rank:
                continue
            raw = bytes(allh[k].cpu().tolist())
            p = ctypes.c_void_p()
            check(self.lib.fr_xchg_open(ctypes.create_string_buffer(raw, 64), ctypes.byref(p)), "fr_xchg_open")
            self.peers[k] = p.value
            self._opened.append(p.value)
        self.barriers = 1
        self._s = None
        dist.barrier(group=group)

    def close(self):
        torch.cuda.synchronize()
        for p in self._opened:
            self.lib.fr_xchg_close(p)
        self._opened = []
        if self.xchg_own:
            self.lib.fr_xchg_free(self.xchg_own)
            self.xchg_own = None

    # ------------------------------------------------------------------ step
    def _struct(self):
        s = self._s
        if s is not None:
            return s
        d = self.data
        s = FocfShardStep()
        s.U, s.I, s.mU, s.vU, s.mI, s.vI = (ptr(t) for t in (self.U, self.I, self.mU, self.vU, self.mI, self.vI))
        s.n_users_loc, s.n_items_loc, s.n_items, s.d = d.n_users_loc, d.n_items_loc, d.n_items, self.d
        s.rank, s.world = self.rank, self.world
        s.item_off, s.train_uid, s.train_rating, s.sst_of_user = (ptr(d.item_off), ptr(d.train_uid), ptr(d.train_rating),
                                                                   ptr(d.sst_of_user))
        s.objective, s.fair_weight, s.adam_mode = self.objective, self.fair_weight, self.adam_mode
        a = self.adam
        s.lr, s.beta1, s.beta2, s.eps, s.weight_decay = (float(a["lr"]), float(a["beta1"]), float(a["beta2"]), float(a["eps"]),
                                                         float(a["weight_decay"]))
        if self.adam_mode == _lib.ADAM_LAZY_EXACT:
            s.last_step_u, s.last_step_i, s.adam_scalars = ptr(self.last_u), ptr(self.last_i), ptr(self.scalars)
            s.scalars_cap = self.scalars.numel() // 2
            s.scalars_filled = ctypes.pointer(self._filled)
        s.uid, s.iid, s.rating, s.sst = (ptr(c) for c in self.cols)
        s.pred, s.loss, s.status_flags = ptr(self.pred), ptr(self.loss), ptr(self.flags)
        s.workspace, s.workspace_bytes = ptr(self.ws), self.ws.numel()
        for k in range(self.world):
            s.xchg[k] = self.peers[k]
        s.J_cap, s.barriers = self.J_cap, self.barriers
        self._s = s
        return s

    def _set_batch(self, s, plan, k):
        b = plan["desc"][k]
        if b["B_loc"] > self.cap:
            raise RuntimeError(f"local batch of {b['B_loc']} rows exceeds the columns ({self.cap}): raise max_batch_loc")
        base_i, base_o = plan["items"].data_ptr(), plan["offs"].data_ptr()
        s.draw_items = base_i + 4 * b["items_pos"]
        s.draw_off = base_o + 4 * b["offs_pos"]
        s.draw_slot = plan["slots"].data_ptr() + 4 * b["items_pos"]
        s.J, s.B_loc, s.B_glob = b["J"], b["B_loc"], b["B_glob"]

    def _set_stage(self, s, plan, k, parity):
        b = plan["desc"][k]
        s.stage_items = plan["items"].data_ptr() + 4 * b["items_pos"]
        s.stage_J, s.stage_parity = b["J"], parity

    def stage(self, plan, k):
        """before the first step of a run: push the rows of batch k's items (into the half batch k will read)"""
        s = self._struct()
        s.step = self.adam["step"] + 1
        self._set_stage(s, plan, k, (self.adam["step"] + 1) & 1)
        check(self.lib.fr_focf_shard_step_run(ctypes.byref(s), _lib.SHARD_STAGE, stream_ptr()), "fr_focf_shard_step_run")

    def run(self, plan, k, phases, loss_out=None, next_k=None):
        """phases of the step on batch k of `plan` (optimizer step = adam['step'] + 1; the caller advances the count with
        `advance()` once phase C has run everywhere).  SHARD_STAGE in `phases` stages batch next_k for the following step."""
        s = self._struct()
        t = self.adam["step"] + 1
        s.step, s.parity = t, t & 1
        self._set_batch(s, plan, k)
        s.loss = ptr(self.loss if loss_out is None else loss_out)
        if phases & _lib.SHARD_STAGE:
            self._set_stage(s, plan, k if next_k is None else next_k, (t + 1) & 1)
        check(self.lib.fr_focf_shard_step_run(ctypes.byref(s), int(phases), stream_ptr()), "fr_focf_shard_step_run")

    def advance(self):
        self.adam["step"] += 1

    def local_batch_to_host(self, plan, k):
        """this rank's rows of batch k as ONE pinned host buffer (int32 local user rows | int32 draw positions | f32 ratings |
        f32 attribute values) -- the host-side Interaction of the sharded step, for end-to-end runs that start from host
        batches like the reference's loop does (trainer.py:181-184)"""
        b = plan["desc"][k]
        n = b["B_loc"]
        s = self._struct()
        self._set_batch(s, plan, k)
        buf = torch.empty(16 * max(n, 1), dtype=torch.uint8).pin_memory()
        if n:
            d = self.data
            uid, iid, rating, sst = (c[:n] for c in self.cols)
            check(self.lib.fr_focf_gather_batch(ptr(d.item_off), ptr(d.train_uid), ptr(d.train_rating), ptr(d.sst_of_user),
                                                s.draw_items, s.draw_off, b["J"], ptr(uid), ptr(iid), ptr(rating), ptr(sst),
                                                stream_ptr()), "fr_focf_gather_batch")
            # fr_focf_gather_batch writes item ids; the sharded step indexes the staged rows by draw position
            off = plan["offs"][b["offs_pos"]:b["offs_pos"] + b["J"] + 1].long()
            pos = torch.repeat_interleave(torch.arange(b["J"], device=off.device, dtype=torch.int32), off[1:] - off[:-1])
            torch.cuda.synchronize()
            for j, col in enumerate((uid, pos, rating, sst)):
                buf[4 * n * j:4 * n * (j + 1)].copy_(col.contiguous().view(torch.uint8).cpu())
        return buf, n

    def train_step_host(self, plan, k, host_batch, next_k=None, loss_out=None):
        """one whole step from a HOST batch of this rank's rows (local_batch_to_host layout): one H2D copy, then the step"""
        buf, n = host_batch
        if n:
            for j, col in enumerate(self.cols):
                col[:n].view(torch.uint8).copy_(buf[4 * n * j:4 * n * (j + 1)], non_blocking=True)
        s = self._struct()
        s.prebuilt = 1
        try:
            self.train_step(plan, k, next_k, loss_out)
        finally:
            s.prebuilt = 0

    def train_step(self, plan, k, next_k=None, loss_out=None):
        """one whole step (real multi-process run: the phases are separated by cross-GPU barriers inside the call)"""
        ph = _lib.SHARD_A | _lib.SHARD_B | _lib.SHARD_C | (_lib.SHARD_STAGE if next_k is not None else 0)
        self.run(plan, k, ph, loss_out, next_k)
        self.advance()

    def flush(self):
        """lazy_exact: bring every local row up to the current optimizer step (call before the tables are read)"""
        if self.adam_mode != _lib.ADAM_LAZY_EXACT or self.adam["step"] < 1:
            return
        s = self._struct()
        s.step = self.adam["step"]
        check(self.lib.fr_focf_shard_step_run(ctypes.byref(s), _lib.SHARD_FLUSH, stream_ptr()), "fr_focf_shard_step_run")

    def check_flags(self):
        f = int(self.flags.item())
        if f:
            self.flags.zero_()
        if f & _lib.FLAG_XCHG_TIMEOUT:
            raise RuntimeError("row-sharded step: a peer did not arrive at a cross-GPU barrier within 20 s")
        if f & _lib.FLAG_TOO_MANY_GROUPS:
            raise IndexError("index 2 is out of bounds for dimension 1 with size 2 "
                             "(more than two sensitive-attribute values in a batch, focf.py:86)")
        if f & _lib.FLAG_NAN_LOSS:
            raise ValueError("Training loss is nan")

    def full_tables(self, group=None):
        """(U, I) with all rows, assembled over torch.distributed (evaluation, checkpoints, parity checks)"""
        import torch.distributed as dist
        self.flush()
        out = []
        for loc, n in ((self.U, self.data.n_users), (self.I, self.data.n_items)):
            rows = local_rows(n, 0, self.world)
            pad = torch.zeros((rows, self.d), dtype=torch.float32, device=loc.device)
            pad[:loc.shape[0]] = loc
            parts = [torch.empty_like(pad) for _ in range(self.world)]
            dist.all_gather(parts, pad, group=group)
            full = torch.stack(parts, dim=1).reshape(rows * self.world, self.d)[:n]
            out.append(full.contiguous())
        return out


class ShardedGroupEmu:
    """P emulated ranks in ONE process on ONE device.  The exchange buffers of the "peers" are ordinary allocations of the
    same device and the host sequences the phases (every rank's A, then every rank's B, ...) where a real run has cross-GPU
    barriers -- the kernels, the exchange layout and the arithmetic are exactly those of the multi-process run."""

    def __init__(self, uid, iid, rating, sst_of_user, n_users, n_items, world, device, d, train_batch_size, seed, **kw):
        self.world = world
        datas = [ShardedTrainData(uid, iid, rating, sst_of_user, n_users, n_items, r, world, device) for r in range(world)]
        self.loaders = [ShardedFOCFLoader(train_batch_size, dt, seed) for dt in datas]
        J_cap = self.loaders[0].J_cap
        max_loc = max(int(dt.n_rows_loc) for dt in datas)
        max_loc = min(max_loc, train_batch_size + int(datas[0].item_count_h.max()))
        self.ranks = [ShardedFOCF(dt, d, J_cap=J_cap, max_batch_loc=max_loc, **kw) for dt in datas]
        for a in self.ranks:
            for k, b in enumerate(self.ranks):
                a.peers[k] = b.xchg_own
            a._s = None
        self.n_users, self.n_items = n_users, n_items

    def set_tables(self, U, I):
        for r in self.ranks:
            r.set_tables(U, I)

    def plan(self, n):
        return [ld.plan(n) for ld in self.loaders]

    def train(self, plans, losses=None):
        n = len(plans[0]["desc"])
        for r, p in zip(self.ranks, plans):
            r.stage(p, 0)
        for k in range(n):
            for ph in (_lib.SHARD_A, _lib.SHARD_B, _lib.SHARD_C):
                for r, p in zip(self.ranks, plans):
                    r.run(p, k, ph, None if losses is None else losses[r.rank][k:k + 1])
            for r in self.ranks:
                r.advance()
            if k + 1 < n:
                for r, p in zip(self.ranks, plans):
                    r.stage(p, k + 1)

    def full_tables(self):
        outs = []
        for r in self.ranks:
            r.flush()
        for attr, n in (("U", self.n_users), ("I", self.n_items)):
            full = torch.zeros((n, self.ranks[0].d), dtype=torch.float32, device=self.ranks[0].U.device)
            for r in self.ranks:
                full[r.rank::self.world] = getattr(r, attr)
            outs.append(full)
        return outs

    def close(self):
        for r in self.ranks:
            r.close()


def selfcheck(rank, world, device, group=None, adam_mode="lazy_exact", n_users=20001, n_items=3001, d=128, n_inter=400_000,
              batch=1 << 15, steps=4, seed=7, rtol=1e-5):
    """Multi-process correctness check of the row-sharded step (run by bench.py's `dp_check` on the driver's multi-GPU
    lines and by tests/test_dp_gpu.py): every rank trains `steps` batches through the real path (CUDA IPC exchange memory,
    cross-GPU barriers); rank 0 then replays the same batches through the single-GPU fused step (FOCF.train_step, dense
    Adam) and compares the losses and the gathered tables.  Returns a dict; `pass` is the verdict (valid on rank 0)."""
    import torch.distributed as dist
    from . import synth
    from .config import Config
    from .dataloader import FOCFDataLoader, TrainData
    from .focf import FOCF
    from .interaction import Interaction
    data = synth.device_interactions(n_users, n_items, n_inter, seed, device)
    tr_u, tr_i, tr_r = data["train"]
    gender = data["gender"]
    sd = ShardedTrainData(tr_u, tr_i, tr_r.float(), gender, n_users, n_items, rank, world, device)
    loader = ShardedFOCFLoader(batch, sd, seed)
    plan = loader.plan(steps)
    model = ShardedFOCF(sd, d, objective="value", fair_weight=1.0, adam_mode=adam_mode, J_cap=loader.J_cap,
                        max_batch_loc=max(loader.max_batch_loc, 1), max_steps=steps + 8)
    g = torch.Generator(device=device).manual_seed(seed)
    U0 = torch.randn((n_users, d), generator=g, device=device) * 0.2
    I0 = torch.randn((n_items, d), generator=g, device=device) * 0.2
    model.set_tables(U0, I0)
    losses = torch.zeros(steps, device=device)
    out = {"world": world, "steps": steps, "adam_mode": adam_mode, "shape": [n_users, n_items, d, batch]}
    try:
        model.connect(group)
        model.stage(plan, 0)
        for k in range(steps):
            model.train_step(plan, k, next_k=k + 1 if k + 1 < steps else None, loss_out=losses[k:k + 1])
        model.check_flags()
        U, I = model.full_tables(group)
        torch.cuda.synchronize()
    finally:
        model.close()
    if rank == 0:
        cfg = Config(embedding_size=d, fair_objective="value", fair_weight=1.0, device=device,
                     train_batch_size=max(b["B_glob"] for b in plan["desc"]))
        train = TrainData.from_device(tr_u, tr_i.long(), tr_r, gender, n_users, n_items)
        items = plan["items"].cpu().numpy()
        draws = np.concatenate([np.r_[items[b["items_pos"]:b["items_pos"] + b["J"]], -1] for b in plan["desc"]])
        ld = FOCFDataLoader(cfg, train, draws=draws)
        ref = FOCF(cfg, synth.SynthDataset(n_users, n_items, 5.0)).to(device)
        with torch.no_grad():
            ref.user_embedding_layer.weight.copy_(U0)
            ref.item_embedding_layer.weight.copy_(I0)
        ref.init_adam(lr=1e-3, weight_decay=1e-3)
        ref_losses = torch.zeros(steps, device=device)
        it, of, bs = ld.plan_epoch(steps)
        d_it, d_of = torch.from_numpy(it).to(device), torch.from_numpy(of).to(device)
        uf, itf, rf, sf = train.fields
        for k, b in enumerate(bs):
            u, i, r, s_ = ld.gather(d_it, d_of, b)
            inter = Interaction({uf: u, itf: i, rf: r, sf: s_})
            inter.items_contiguous = True
            ref.train_step(inter, loss_out=ref_losses[k:k + 1])
        ref.check_flags()

        def rel(a, b):
            return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

        out["loss_rel_err"] = float(((losses - ref_losses).abs() / ref_losses.abs()).max())
        out["U_rel_err"] = rel(U, ref.user_embedding_layer.weight.data)
        out["I_rel_err"] = rel(I, ref.item_embedding_layer.weight.data)
        out["pass"] = bool(max(out["loss_rel_err"], out["U_rel_err"], out["I_rel_err"]) < rtol)
        out["what"] = ("row-sharded step over CUDA-IPC peer memory vs the single-GPU fused step on the same batches: losses "
                       "and updated tables within %g relative" % rtol)
    if world > 1:
        dist.barrier(group=group)
    return out
