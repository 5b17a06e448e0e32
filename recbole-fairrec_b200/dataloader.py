"""Device-side FOCF batch builder -- replaces FOCFDataLoader (recbole/data/dataloader/focf_dataloader.py:5-50)
and the per-batch `Dataset.__getitem__` join (84 % of the reference's train epoch).

The train split is sorted by item once (stable, like focf_dataloader.py:12 `dataset.sort(by=ITEM_ID)`) and kept
on the device in CSC form; a batch is "all train rows of randomly drawn items until >= train_batch_size rows"
(focf_dataloader.py:37-50).  The item draws of a whole epoch are made on the host in one go and uploaded with ONE
copy; each batch is then materialised by the `fr_focf_gather_batch` kernel.

Draw modes:
  * "reference" -- calls `np.random.choice(candidates, 1, False)` exactly like focf_dataloader.py:43, so with the
    same numpy global seed the batches are IDENTICAL to the reference's (parity tests; O(n_items) per draw).
  * "fast"      -- one permutation of the candidate items per batch from a private Generator (same distribution:
    uniform without replacement inside a batch, fresh for every batch).
  * explicit `draws` (array, -1 terminates a batch) -- replay of recorded draws (golden fixtures).
"""
import math

import numpy as np
import torch

from . import _lib
from ._lib import check, load, ptr, stream_ptr
from .interaction import Interaction


class TrainData:
    """Item-sorted, device-resident train split + what the trainer needs from `train_data.dataset`."""

    def __init__(self, uid, iid, rating, sst_of_user, n_users, n_items, device, uid_field="user_id",
                 iid_field="item_id", rating_field="rating", sst_field="gender"):
        uid, iid = np.asarray(uid), np.asarray(iid)
        order = np.argsort(iid, kind="stable")                      # interaction.py:334
        self.uid_h = np.ascontiguousarray(uid[order], dtype=np.int32)
        self.iid_h = np.ascontiguousarray(iid[order], dtype=np.int32)
        self.rating_h = np.ascontiguousarray(np.asarray(rating)[order], dtype=np.float32)
        self.n_users, self.n_items, self.device = int(n_users), int(n_items), device
        self.n_rows = len(self.uid_h)
        self.item_count_h = np.bincount(self.iid_h, minlength=self.n_items).astype(np.int64)
        self.item_off_h = np.zeros(self.n_items + 1, np.int32)
        self.item_off_h[1:] = np.cumsum(self.item_count_h)
        self.item_uniques = np.nonzero(self.item_count_h)[0]        # focf_dataloader.py:14
        self.fields = (uid_field, iid_field, rating_field, sst_field)
        self.max_rating = float(self.rating_h.max())
        t = lambda a, dt: torch.as_tensor(a, dtype=dt).to(device)
        self.item_off = t(self.item_off_h, torch.int32)
        self.train_uid = t(self.uid_h, torch.int32)
        self.train_rating = t(self.rating_h, torch.float32)
        self.sst_of_user = t(np.asarray(sst_of_user, dtype=np.float32), torch.float32)

    @classmethod
    def from_device(cls, uid, iid, rating, sst_of_user, n_users, n_items, uid_field="user_id", iid_field="item_id",
                    rating_field="rating", sst_field="gender"):
        """The same layout built ON THE DEVICE from device tensors (scale-out shapes: 10^9 rows never visit the host):
        stable sort by item, CSC offsets from a bincount; only the per-item counts (n_items values) go to the host,
        where the batch draws are planned."""
        self = cls.__new__(cls)
        dev = uid.device
        order = torch.sort(iid, stable=True).indices
        self.train_uid = uid[order].to(torch.int32).contiguous()
        self.train_rating = rating[order].to(torch.float32).contiguous()
        del order
        counts = torch.bincount(iid, minlength=int(n_items))
        self.item_count_h = counts.cpu().numpy().astype(np.int64)
        self.item_off_h = np.zeros(int(n_items) + 1, np.int64)
        self.item_off_h[1:] = np.cumsum(self.item_count_h)
        if self.item_off_h[-1] >= 2 ** 31:
            raise ValueError("the device-side batch builder indexes the train split with int32 offsets")
        self.item_off_h = self.item_off_h.astype(np.int32)
        self.item_uniques = np.nonzero(self.item_count_h)[0]
        self.item_off = torch.from_numpy(self.item_off_h).to(dev)
        self.sst_of_user = sst_of_user.to(device=dev, dtype=torch.float32).contiguous()
        self.n_users, self.n_items, self.device = int(n_users), int(n_items), dev
        self.n_rows = int(self.train_uid.numel())
        self.fields = (uid_field, iid_field, rating_field, sst_field)
        self.max_rating = float(self.train_rating.max())
        self.uid_h = self.iid_h = self.rating_h = None
        return self

    # the two `dataset` members the reference trainer/model read (collector.py:91-93, focf.py:40)
    @property
    def item_counter(self):
        return {int(i): int(self.item_count_h[i]) for i in self.item_uniques}


class FOCFDataLoader:
    def __init__(self, config, train, mode="fast", draws=None, seed=None, partition=None):
        """partition=(rank, world): data-parallel training -- this rank draws only items of its slice
        item_uniques[rank::world] (whole items stay on one rank, so item x group sums need no exchange) and an epoch
        has ceil(N_train / (train_batch_size * world)) global steps of `world` local batches each."""
        self.config, self.train = config, train
        self.partition = partition
        self.candidates = train.item_uniques if partition is None else train.item_uniques[partition[0]::partition[1]]
        self.dataset = train
        self.step = int(config["train_batch_size"])
        self.mode = mode
        self._replay = None if draws is None else np.asarray(draws)
        self._rng = np.random.default_rng(config["seed"] if seed is None else seed)
        # fast mode: items drawn per batch at first (3x what an average batch needs, at least 16)
        mean_cnt = max(float(train.n_rows) / max(len(self.candidates), 1), 1.0)
        self._prefix = max(16, int(3 * self.step / mean_cnt) + 8)
        self.lib = load()
        dev = train.device
        # worst case: step-1 rows + the most popular item
        self.max_batch = self.step + int(train.item_count_h.max())
        self._cols = (torch.empty(self.max_batch, dtype=torch.int32, device=dev),
                      torch.empty(self.max_batch, dtype=torch.int32, device=dev),
                      torch.empty(self.max_batch, dtype=torch.float32, device=dev),
                      torch.empty(self.max_batch, dtype=torch.float32, device=dev))
        self._replay_pos = 0
        self._pool0 = np.sort(np.asarray(self.candidates, np.int64))     # items present in train (this rank's slice)

    def __len__(self):
        world = 1 if self.partition is None else self.partition[1]
        return math.ceil(self.train.n_rows / (self.step * world))    # abstract_dataloader.py:67-68

    # ------------------------------------------------------------------ host-side draws
    def _draw_batch(self):
        tr = self.train
        if self._replay is not None:
            end = self._replay_pos
            while self._replay[end] != -1:
                end += 1
            items = self._replay[self._replay_pos:end]
            self._replay_pos = end + 1
            return np.asarray(items, dtype=np.int64)
        if self.mode == "reference":
            # focf_dataloader.py:38-47 in effect, on numpy's global RNG: `np.random.choice(pool, 1, False)` IS
            # `pool[np.random.permutation(len(pool))[:1]]` (legacy RandomState.choice without replacement), so the same
            # stream is consumed and the same items come out; the pool of still-selectable items (ascending, like
            # `select_item[is_select]`) is kept as an array instead of being re-masked for every draw
            pool = self._pool0
            cnt, items = 0, []
            while cnt < self.step:
                if len(pool) == 0:
                    raise ValueError("a must be non-empty")          # what np.random.choice raises in the reference
                j = int(np.random.permutation(len(pool))[0])
                iid = int(pool[j])
                cnt += int(tr.item_count_h[iid])
                pool = np.delete(pool, j)
                items.append(iid)
            if len(pool) == 0:          # focf_dataloader.py:43 `or not any(is_select)`: a batch that uses up every item
                raise ValueError("a must be non-empty")          # makes the reference draw from an empty pool
            return np.asarray(items, dtype=np.int64)
        # whole items in uniformly random order until the batch holds `step` rows = a prefix of a uniform permutation of
        # the candidates.  Only a short prefix is ever used (step / mean item count items), so only a short prefix is
        # drawn: k distinct positions in random order (Floyd's algorithm inside Generator.choice, O(k)); in the rare case
        # that they do not suffice the permutation is continued over the remaining candidates.
        cand = self.candidates
        n = len(cand)
        k = min(n, self._prefix)
        idx = self._rng.choice(n, size=k, replace=False, shuffle=True)
        csum = np.cumsum(tr.item_count_h[cand[idx]])
        if csum[-1] < self.step and k < n:
            rest = np.setdiff1d(np.arange(n), idx, assume_unique=True)
            idx = np.concatenate([idx, self._rng.permutation(rest)])
            csum = np.cumsum(tr.item_count_h[cand[idx]])
        j = int(np.searchsorted(csum, self.step, side="left")) + 1
        return cand[idx[:j]].astype(np.int64)

    def plan_epoch(self, n_batches=None):
        """Draw every batch of the epoch (or only the first n_batches); returns (draw_items, draw_off, batches) where
        batches is a list of (first draw index, first offset index, J, B) and draw_off holds, per batch, J+1 positions."""
        items, offs, batches = [], [], []
        pos_items = pos_offs = 0
        for _ in range(len(self) if n_batches is None else int(n_batches)):
            it = self._draw_batch()
            cnt = self.train.item_count_h[it]
            off = np.zeros(len(it) + 1, np.int64)
            off[1:] = np.cumsum(cnt)
            batches.append((pos_items, pos_offs, len(it), int(off[-1])))
            items.append(it)
            offs.append(off)
            pos_items += len(it)
            pos_offs += len(it) + 1
        return (np.concatenate(items).astype(np.int32), np.concatenate(offs).astype(np.int32), batches)

    # ------------------------------------------------------------------ device-side materialisation
    def gather(self, d_items, d_offs, batch):
        """run fr_focf_gather_batch for one planned batch; returns the four columns (views of reused buffers)"""
        pi, po, J, B = batch
        tr = self.train
        uid, iid, rating, sst = (c[:B] for c in self._cols)
        check(self.lib.fr_focf_gather_batch(ptr(tr.item_off), ptr(tr.train_uid), ptr(tr.train_rating),
                                            ptr(tr.sst_of_user), d_items[pi:pi + J].data_ptr(),
                                            d_offs[po:po + J + 1].data_ptr(), J, ptr(uid), ptr(iid), ptr(rating),
                                            ptr(sst), stream_ptr()), "fr_focf_gather_batch")
        return uid, iid, rating, sst

    def plan_epoch_device(self, buf=0):
        """Draw the whole epoch on the host and place the plan in PERSISTENT device buffers (their addresses stay the
        same from epoch to epoch, so a captured CUDA graph of the step keeps pointing at them).  Returns the plan dict
        consumed by FocfEngine.planned_step; plan["rows"] is the host-side total number of interactions.
        buf: which of the loader's plan buffer sets to fill -- a caller that plans epoch e + 1 on the host while epoch e
        still runs on the device (FOCF.train_epochs_planned) alternates 0 / 1."""
        items, offs, batches = self.plan_epoch()
        desc = np.asarray(batches, dtype=np.int32).reshape(-1, 4)
        dev = self.train.device
        if not hasattr(self, "_plans"):
            self._plans = {}
        p = self._plans.get(buf)
        if p is None or p["items"].numel() < len(items) or p["offs"].numel() < len(offs) or p["desc"].shape[0] < len(desc):
            grow = lambda n: int(n * 1.5) + 64
            p = dict(items=torch.zeros(grow(len(items)), dtype=torch.int32, device=dev),
                     offs=torch.zeros(grow(len(offs)), dtype=torch.int32, device=dev),
                     desc=torch.zeros((grow(len(desc)), 4), dtype=torch.int32, device=dev),
                     cols=self._cols, generation=(0 if p is None else p["generation"] + 1), buf=buf)
            self._plans[buf] = p
        self._plan = p          # the plan drawn last
        if p.get("copied") is not None:
            p["copied"].synchronize()          # the previous epoch's copies have read the pinned staging buffers
        for name, arr in (("items", items), ("offs", offs), ("desc", desc)):
            host = p.get("host_" + name)
            if host is None or host.shape[0] < len(arr):   # pinned staging, allocated once (pin_memory() per epoch costs more than the copy)
                host = p["host_" + name] = torch.empty((p[name].shape[0],) + tuple(arr.shape[1:]), dtype=torch.int32).pin_memory()
            host[:len(arr)].copy_(torch.from_numpy(arr))
            p[name][:len(arr)].copy_(host[:len(arr)], non_blocking=True)
        p["copied"] = torch.cuda.Event()
        p["copied"].record()
        p["len"], p["rows"], p["batch_rows"] = len(desc), int(desc[:, 3].sum()), desc[:, 3].tolist()
        return p

    def __iter__(self):
        items, offs, batches = self.plan_epoch()
        dev = self.train.device
        d_items = torch.from_numpy(items).pin_memory().to(dev, non_blocking=True)
        d_offs = torch.from_numpy(offs).pin_memory().to(dev, non_blocking=True)
        uf, itf, rf, sf = self.train.fields
        for b in batches:
            uid, iid, rating, sst = self.gather(d_items, d_offs, b)
            inter = Interaction({uf: uid, itf: iid, rf: rating, sf: sst})
            inter.items_contiguous = True
            yield inter
