"""CPU port of the reference's hot loops in the reference's own vocabulary of PyTorch/NumPy ops, multi-threaded,
used ONLY as the timed CPU baseline (`bench.py` cpu_baseline / `--impl reference`) and checked against the golden
fixtures in tests/test_oracle_golden.py.  kind = "port": the reference itself is Python and cannot travel to the GPU
box, so this file restates its per-batch op sequence with stock torch CPU kernels:

  train : FOCFDataLoader._next_batch_data (focf_dataloader.py:37-50: np.random.choice + np.where over the whole
          split + fancy-index join) -> FOCF.calculate_loss (focf.py:152-169: embedding, mul/sum, torch.unique x2,
          index_put_ x3, smooth_l1) -> backward -> torch.optim.Adam(weight_decay) dense step -> loss.item()
  eval  : FullSortEvalDataLoader batches of max(eval_batch_size // n_items, 1) users (general_dataloader.py:209-245)
          -> mm + clamp (focf.py:171-178) -> -inf masks (trainer.py:435-438) -> topk / pos-matrix / gather
          (collector.py:141-189) with per-batch torch.cat (collector.py:46-52) -> numpy metrics (metrics.py)

TEST INFRASTRUCTURE ONLY -- never imported by the product package.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import metrics_oracle as mo


class TorchFOCF(torch.nn.Module):
    def __init__(self, U0, I0, objective="value", fair_weight=1.0, max_rating=5.0):
        super().__init__()
        self.user_emb = torch.nn.Embedding.from_pretrained(torch.as_tensor(U0).clone(), freeze=False)
        self.item_emb = torch.nn.Embedding.from_pretrained(torch.as_tensor(I0).clone(), freeze=False)
        self.objective, self.fair_weight, self.max_rating = objective, fair_weight, max_rating

    def scores(self, user, item):
        return (self.user_emb(user) * self.item_emb(item)).sum(-1)

    def _cells(self, pred, item, rating, sst):
        _, g = torch.unique(sst, return_inverse=True)
        items, j = torch.unique(item, return_inverse=True)
        shape = (len(items), 2)
        sp = torch.zeros(shape).index_put_((j, g), pred, accumulate=True)
        st = torch.zeros(shape).index_put_((j, g), rating, accumulate=True)
        n = torch.zeros(shape).index_put_((j, g), torch.ones(len(pred)), accumulate=True) + 1e-5
        return sp / n, st / n

    def fair(self, pred, item, rating, sst):
        if self.objective == "none":
            return 0.0
        if self.objective == "nonparity":
            v = torch.unique(sst)
            return F.smooth_l1_loss(pred[sst == v[0]].mean(), pred[sst == v[1]].mean())
        P, T = self._cells(pred, item, rating, sst)
        zero = torch.tensor(0.0)
        D = {"value": lambda: P - T, "absolute": lambda: (P - T).abs(),
             "under": lambda: torch.where(T - P > zero, T - P, zero),
             "over": lambda: torch.where(P - T > zero, P - T, zero)}[self.objective]()
        x = (D[:, 0] - D[:, 1]).abs()
        return F.smooth_l1_loss(x, torch.zeros_like(x))

    def loss(self, user, item, rating, sst):
        pred = self.scores(user, item)
        return F.mse_loss(pred, rating) + self.fair_weight * self.fair(pred, item, rating, sst)

    def full_sort(self, users):
        s = torch.mm(self.user_emb(users), self.item_emb.weight.t()).view(-1)
        return torch.clamp(s, min=0.0, max=self.max_rating) / self.max_rating


class RefStyleLoader:
    """the reference's batch construction, including its O(N_train)-per-item scan"""

    def __init__(self, train_u, train_i, train_r, sst_of_user, n_items, step):
        o = np.argsort(train_i, kind="stable")
        self.u = torch.as_tensor(np.asarray(train_u)[o]).long()
        self.i = torch.as_tensor(np.asarray(train_i)[o]).long()
        self.r = torch.as_tensor(np.asarray(train_r)[o]).float()
        self.sst = torch.as_tensor(np.asarray(sst_of_user)).long()
        self.n_items, self.step = n_items, step
        self.item_uniques = np.unique(self.i.numpy())

    def next_batch(self):
        sel = np.zeros(self.n_items, bool)
        sel[self.item_uniques] = True
        cand = np.arange(self.n_items)
        cnt, rows = 0, []
        while cnt < self.step:
            it = np.random.choice(cand[sel], 1, False)[0]
            idx = np.where(self.i == it)[0]
            cnt += len(idx)
            sel[it] = False
            rows.extend(idx)
        u = self.u[rows]
        return u, self.i[rows], self.r[rows], self.sst[u]


def train_steps(model, loader, n_steps, lr=1e-3, weight_decay=1e-3, prebuilt=None):
    """trainer.py:181-196.  Returns (#interactions processed, summed loss)."""
    opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=weight_decay)
    rows, total = 0, 0.0
    for s in range(n_steps):
        u, i, r, g = prebuilt[s] if prebuilt is not None else loader.next_batch()
        opt.zero_grad()
        loss = model.loss(u, i, r, g)
        total += loss.item()
        loss.backward()
        opt.step()
        rows += len(u)
    return rows, total


@torch.no_grad()
def evaluate(model, users, hist, pos, sst_of_user, n_items, topk, count_items, users_per_batch, popularity_ratio=0.1):
    """trainer.py:505-512 + collector.py + metrics.py for the eval users given (lists per user)."""
    K = max(topk)
    acc = {}

    def update(name, t):  # collector.py:46-52: concatenate on the host every batch
        acc[name] = t.clone() if name not in acc else torch.cat((acc[name], t), dim=0)

    sst_t = torch.as_tensor(np.asarray(sst_of_user))
    for b0 in range(0, len(users), users_per_batch):
        bu = users[b0:b0 + users_per_batch]
        h = [torch.as_tensor(np.asarray(x)).long() for x in hist[b0:b0 + users_per_batch]]
        p = [torch.as_tensor(np.asarray(x)).long() for x in pos[b0:b0 + users_per_batch]]
        hu = torch.cat([torch.full_like(x, k) for k, x in enumerate(h)])
        pu = torch.cat([torch.full_like(x, k) for k, x in enumerate(p)])
        hi, pi = torch.cat(h), torch.cat(p)
        ut = torch.as_tensor(np.asarray(bu)).long()
        s = model.full_sort(ut).view(-1, n_items)
        s[:, 0] = -np.inf
        s[hu, hi] = -np.inf
        _, idx = torch.topk(s, K, dim=-1)
        update("rec.items", idx)
        pm = torch.zeros_like(s, dtype=torch.int)
        pm[pu, pi] = 1
        update("rec.topk", torch.cat((torch.gather(pm, 1, idx), pm.sum(dim=1, keepdim=True)), dim=1))
        update("rec.positive_score", s[pu, pi])
        update("data.positive_i", pi)
        update("data.sst", sst_t[ut][pu])
    st = {k: v.numpy() for k, v in acc.items()}
    return mo.evaluate(st, list(topk), n_items, count_items, popularity_ratio), st
