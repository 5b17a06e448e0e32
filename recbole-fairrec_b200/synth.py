"""Synthetic data of the benchmark shapes (the real ML-1M .inter is a missing blob in the
reference tree, and there is no network).  Pure numpy on the host; deterministic per seed."""
import numpy as np


class SynthDataset:
    """What FOCF.__init__ reads from `dataset` (abstract_recommender.py:96-97, focf.py:40)."""

    def __init__(self, n_users, n_items, max_rating):
        import torch
        self._n = {"user_id": n_users, "item_id": n_items}
        self.inter_feat = {"rating": torch.tensor([1.0, float(max_rating)])}

    def num(self, field):
        return self._n[field]


def interactions(n_users, n_items, n_inter, seed=2020, user_sigma=1.0, item_sigma=1.4):
    """Unique (user, item) pairs with log-normal user activity / item popularity, ratings 1..5 with ML-1M's
    marginal (p = .056 .107 .261 .349 .227), a binary gender per user ~ Bernoulli(0.28) stored as token ids
    1/2 (row 0 = [PAD]).  ids start at 1."""
    rng = np.random.default_rng(seed)
    pu = rng.lognormal(4.5, user_sigma, n_users - 1)
    pi = rng.lognormal(4.5, item_sigma, n_items - 1)
    pu, pi = pu / pu.sum(), pi / pi.sum()
    pairs = np.zeros(0, np.int64)
    need = n_inter
    while need > 0:
        m = int(need * 1.3) + 1024
        u = rng.choice(n_users - 1, size=m, p=pu) + 1
        i = rng.choice(n_items - 1, size=m, p=pi) + 1
        pairs = np.unique(np.concatenate([pairs, u.astype(np.int64) * n_items + i]))
        need = n_inter - len(pairs)
    pairs = rng.permutation(pairs)[:n_inter]
    uid, iid = (pairs // n_items).astype(np.int32), (pairs % n_items).astype(np.int32)
    rating = rng.choice(np.arange(1, 6), size=n_inter, p=[.056, .107, .261, .349, .227]).astype(np.float32)
    gender = (rng.random(n_users) < 0.28).astype(np.float32) + 1.0
    gender[0] = 0.0
    return uid, iid, rating, gender


def split_by_user(uid, iid, rating, ratios=(0.8, 0.1, 0.1), seed=2020):
    """dataset.py:1362-1396 split_by_ratio grouped by user on a shuffled order (RO): per user, the first 80 % of
    its rows go to train, the next 10 % to valid, the rest to test."""
    rng = np.random.default_rng(seed)
    p = rng.permutation(len(uid))
    uid, iid, rating = uid[p], iid[p], rating[p]
    order = np.argsort(uid, kind="stable")
    uid, iid, rating = uid[order], iid[order], rating[order]
    starts = np.flatnonzero(np.r_[True, uid[1:] != uid[:-1]])
    lens = np.diff(np.r_[starts, len(uid)])
    rank = np.arange(len(uid)) - np.repeat(starts, lens)
    tot = np.repeat(lens, lens)
    # dataset.py:1340-1360 _calcu_split_ids: cnt = [int(r * tot)]; cnt[0] = tot - sum(cnt[1:])
    n_va, n_te = (ratios[1] * tot).astype(np.int64), (ratios[2] * tot).astype(np.int64)
    n_tr = tot - n_va - n_te
    part = np.where(rank < n_tr, 0, np.where(rank < n_tr + n_va, 1, 2))
    return [(uid[part == k], iid[part == k], rating[part == k]) for k in range(3)]


def eval_lists(train, valid, test, phase="valid"):
    """general_dataloader.py:173-207 + sampler.py:243-264: eval users (ascending), per-user positives of the phase
    and history = used_ids[phase] - positives (valid: train; test: train + valid)."""
    tr_u, tr_i, _ = train
    ev_u, ev_i, _ = valid if phase == "valid" else test
    used_u, used_i = (tr_u, tr_i) if phase == "valid" else (np.r_[tr_u, valid[0]], np.r_[tr_i, valid[1]])

    def group(u, i):
        o = np.argsort(u, kind="stable")
        u, i = u[o], i[o]
        s = np.flatnonzero(np.r_[True, u[1:] != u[:-1]])
        return u[s], np.split(i, s[1:])

    users, pos = group(ev_u, ev_i)
    hu, hl = group(used_u, used_i)
    hmap = dict(zip(hu.tolist(), hl))
    hist = [hmap.get(int(u), np.zeros(0, np.int32)) for u in users]
    return users, hist, pos


def device_interactions(n_users, n_items, n_inter, seed, device, chunk=100_000_000):
    """BASELINE.json configs[4]-style data synthesised ON THE DEVICE (1e9 rows never visit the host):
    user activity and item popularity log-normal (inverse-CDF sampling), ratings 1..5 with ML-1M's marginal, a binary
    gender ~ Bernoulli(0.28) stored as 1/2 (row 0 = [PAD]); each interaction goes to train / valid / test with probability
    .8 / .1 / .1 (the reference splits 8:1:1 per user; the per-interaction draw has the same shape).  (user, item) pairs
    are not de-duplicated (collision probability ~1e-4 at the full shape).  The same seed gives the same tensors on every
    rank.  Returns dict(train=(uid, iid, rating u8), valid=(uid, iid), gender f32[n_users])."""
    import numpy as np
    import torch
    g = torch.Generator(device=device).manual_seed(int(seed))

    def cdf(n, sigma):
        w = torch.empty(n, device=device).log_normal_(4.5, sigma, generator=g).double()
        return (torch.cumsum(w, 0) / w.sum()).float()

    ucdf, icdf = cdf(n_users - 1, 1.0), cdf(n_items - 1, 1.4)
    rcdf = torch.tensor(np.cumsum([.056, .107, .261, .349, .227]), dtype=torch.float32, device=device)
    gender = (torch.rand(n_users, device=device, generator=g) < 0.28).float() + 1.0
    gender[0] = 0.0
    tr_u, tr_i, tr_r, va_u, va_i = [], [], [], [], []
    for lo in range(0, n_inter, chunk):
        m = min(chunk, n_inter - lo)
        u = (torch.searchsorted(ucdf, torch.rand(m, device=device, generator=g)).clamp_(max=n_users - 2) + 1).to(torch.int32)
        i = (torch.searchsorted(icdf, torch.rand(m, device=device, generator=g)).clamp_(max=n_items - 2) + 1).to(torch.int32)
        r = (torch.searchsorted(rcdf, torch.rand(m, device=device, generator=g)).clamp_(max=4) + 1).to(torch.uint8)
        part = torch.rand(m, device=device, generator=g)
        tr = part < 0.8
        va = (part >= 0.8) & (part < 0.9)
        tr_u.append(u[tr]); tr_i.append(i[tr]); tr_r.append(r[tr]); va_u.append(u[va]); va_i.append(i[va])
        del u, i, r, part, tr, va
    return dict(train=(torch.cat(tr_u), torch.cat(tr_i), torch.cat(tr_r)), valid=(torch.cat(va_u), torch.cat(va_i)),
                gender=gender)
