"""Atomic-file datasets with sensitive attributes -> the device-resident layouts of the hot path.

Replaces, for the fairness configs, the ingestion chain of the reference: `Dataset._load_feat` (recbole/data/dataset/
dataset.py:385-454: `pd.read_csv(engine='python')`), `_remap_ID_all` (920-974: `pd.factorize`, id 0 = '[PAD]'),
`_user_item_feat_preparation` (488-507: feature row index == id), `build` (1467-1514: RO shuffle = `torch.randperm`,
`split_by_ratio` grouped by user, 1362-1396, `_calcu_split_ids` 1339-1360) and `Sampler.get_used_ids`
(recbole/sampler/sampler.py:243-264) / `FullSortEvalDataLoader.__init__` (general_dataloader.py:173-207) -- with the same
id assignment, the same shuffle and the same per-user cut, so that the splits (and everything downstream) are IDENTICAL
to the reference's for a given seed (tests/test_atomic.py pins this on ml-100k against tensors exported from the
reference).  All of it is host-side numpy (vectorised; no per-row Python loops), done once per run.

File format (RecBole atomic files): TSV with a `name:type` header, types token / float / token_seq / float_seq
(sequences are not used by the fairness configs and are skipped)."""
import os

import numpy as np
import torch

from .interaction import Interaction


def read_atomic(path, usecols=None, sep="\t"):
    """-> (columns: {field: np.ndarray}, types: {field: 'token' | 'float'}).  Token columns come back as str arrays (ids are
    assigned by first appearance, so the textual token is what matters), float columns as float64 like pandas gives."""
    import pandas as pd
    with open(path, "r", encoding="utf-8") as f:
        header = f.readline().rstrip("\n").split(sep)
    names, types = zip(*[h.split(":") for h in header])
    keep = [n for n, t in zip(names, types) if t in ("token", "float") and (usecols is None or n in usecols)]
    dtype = {n: (str if t == "token" else np.float64) for n, t in zip(names, types) if n in keep}
    df = pd.read_csv(path, sep=sep, header=0, names=list(names), usecols=keep, dtype=dtype, engine="c",
                     keep_default_na=False, na_values=[""])
    return {n: df[n].to_numpy() for n in keep}, {n: t for n, t in zip(names, types) if n in keep}


def factorize(chunks):
    """pd.factorize over the concatenation of `chunks`, ids + 1 (0 = '[PAD]'), split back (dataset.py:952-974)"""
    import pandas as pd
    tokens = np.concatenate(chunks)
    ids, mp = pd.factorize(tokens)
    out = np.split(ids + 1, np.cumsum([len(c) for c in chunks])[:-1])
    return out, np.array(["[PAD]"] + list(mp), dtype=object)


def calcu_split_counts(tot, ratios):
    """dataset.py:1339-1360 _calcu_split_ids, vectorised over groups: tot [n_groups] -> counts [n_groups, len(ratios)]"""
    tot = np.asarray(tot, np.int64)
    r = np.asarray(ratios, np.float64) / np.sum(ratios)
    cnt = np.stack([(r[i] * tot).astype(np.int64) for i in range(len(r))], axis=1)
    cnt[:, 0] = tot - cnt[:, 1:].sum(axis=1)
    for i in range(1, len(r)):
        frac = r[-i] * tot
        move = (cnt[:, 0] > 1) & (frac > 0) & (frac < 1)
        cnt[move, -i] += 1
        cnt[move, 0] -= 1
    return cnt


class AtomicDataset:
    """The slice of recbole.data.dataset.Dataset the hot path reads: `num`, `inter_feat`, `get_user_feature`,
    `inter_matrix`, `field2id_token`, plus `build()` -> three column dicts (train / valid / test)."""

    def __init__(self, config):
        self.config = config
        name = config["dataset"]
        root = os.path.join(config["data_path"] or "dataset", name)
        self.uid_field, self.iid_field = config["USER_ID_FIELD"], config["ITEM_ID_FIELD"]
        self.rating_field = config["RATING_FIELD"]
        load_col = config["load_col"] or {}
        inter, itypes = read_atomic(os.path.join(root, name + ".inter"), load_col.get("inter"))
        user_path, item_path = os.path.join(root, name + ".user"), os.path.join(root, name + ".item")
        user, utypes = read_atomic(user_path, load_col.get("user")) if os.path.exists(user_path) and \
            (not load_col or "user" in load_col) else ({}, {})
        item, _ = read_atomic(item_path, load_col.get("item")) if os.path.exists(item_path) and "item" in load_col \
            else ({}, {})
        self.field2id_token = {}
        # ---- id remap: interactions first, then the feature file (dataset.py:894-918)
        chunks = [inter[self.uid_field]] + ([user[self.uid_field]] if self.uid_field in user else [])
        ids, self.field2id_token[self.uid_field] = factorize(chunks)
        inter_u = ids[0]
        feat_u = ids[1] if len(ids) > 1 else None
        chunks = [inter[self.iid_field]] + ([item[self.iid_field]] if self.iid_field in item else [])
        ids, self.field2id_token[self.iid_field] = factorize(chunks)
        inter_i = ids[0]
        self.user_num, self.item_num = len(self.field2id_token[self.uid_field]), len(self.field2id_token[self.iid_field])
        cols = {self.uid_field: inter_u.astype(np.int64), self.iid_field: inter_i.astype(np.int64)}
        for f, t in itypes.items():
            if f in (self.uid_field, self.iid_field):
                continue
            cols[f] = inter[f].astype(np.float32) if t == "float" else factorize([inter[f]])[0][0].astype(np.int64)
        # ---- label by threshold (dataset.py:865-892; the rating column is kept)
        thr = config["threshold"]
        if thr:
            (field, value), = thr.items()
            cols[config["LABEL_FIELD"]] = (cols[field] >= value).astype(np.float32)
        self.inter = cols
        # ---- user features: row index == user id, row 0 = [PAD] (dataset.py:488-507); token attributes get ids by first
        # appearance in the FILE (dataset.py:927-929), float attributes stay as they are
        self.user_feat = {self.uid_field: np.arange(self.user_num, dtype=np.int64)}
        if feat_u is not None:
            for f, t in utypes.items():
                if f == self.uid_field:
                    continue
                if t == "token":
                    vals, self.field2id_token[f] = factorize([user[f]])
                    col = np.zeros(self.user_num, np.int64)
                    col[feat_u] = vals[0]
                else:
                    col = np.zeros(self.user_num, np.float32)
                    col[feat_u] = user[f].astype(np.float32)
                self.user_feat[f] = col

    # ------------------------------------------------------------------ the reference's Dataset surface
    def num(self, field):
        if field == self.uid_field:
            return self.user_num
        if field == self.iid_field:
            return self.item_num
        return len(self.field2id_token[field])

    @property
    def inter_feat(self):
        return {k: torch.from_numpy(v) for k, v in self.inter.items()}

    def get_user_feature(self):
        return Interaction({k: torch.from_numpy(v) for k, v in self.user_feat.items()})

    def inter_matrix(self, form="coo", value_field=None):
        """dataset.py:1633-1651 (of whatever interactions this object holds: the train split after build())"""
        import scipy.sparse as sp
        src = getattr(self, "_matrix_src", self.inter)
        data = src[value_field] if value_field else np.ones(len(src[self.uid_field]), np.float32)
        m = sp.coo_matrix((data, (src[self.uid_field], src[self.iid_field])), shape=(self.user_num, self.item_num))
        return m if form == "coo" else m.tocsr()

    def __len__(self):
        return len(self.inter[self.uid_field])

    # ------------------------------------------------------------------ ordering + splitting
    def build(self):
        """eval_args {order: RO, split: {RS: [a, b, c]}, group_by: user} (the setting of all eight fairness YAMLs):
        shuffle with torch.randperm (global torch RNG, as seeded by init_seed), then per user -- users in order of first
        appearance in the shuffled data, rows in shuffled order -- the first a/(a+b+c) go to train, etc."""
        ea = self.config["eval_args"] or {}
        if (ea.get("order") or "RO") != "RO" or "RS" not in (ea.get("split") or {"RS": [8, 1, 1]}) or \
                (ea.get("group_by") or "user") != "user":
            raise NotImplementedError("eval_args other than order RO / split RS / group_by user are outside the fairness configs")
        ratios = (ea.get("split") or {"RS": [8, 1, 1]})["RS"]
        n = len(self)
        perm = torch.randperm(n).numpy()                      # interaction.py:293-297
        cols = {k: v[perm] for k, v in self.inter.items()}
        u = cols[self.uid_field]
        first = np.full(self.user_num, n, np.int64)
        np.minimum.at(first, u, np.arange(n))
        order = np.argsort(first[u], kind="stable")          # groups by first appearance, shuffled order inside
        us = u[order]
        starts = np.flatnonzero(np.r_[True, us[1:] != us[:-1]])
        lens = np.diff(np.r_[starts, n])
        cnt = calcu_split_counts(lens, ratios)
        rank = np.arange(n) - np.repeat(starts, lens)
        edges = np.cumsum(cnt, axis=1)
        part = (rank[:, None] >= np.repeat(edges, lens, axis=0)).sum(axis=1)
        splits = []
        for p in range(len(ratios)):
            idx = order[part == p]
            splits.append({k: v[idx] for k, v in cols.items()})
        self._matrix_src = splits[0]
        return splits


def used_and_positive_lists(splits, phase, uf="user_id", itf="item_id"):
    """sampler.py:243-264 + general_dataloader.py:173-207: eval users (ascending), their positives of the phase (in split
    order) and history = items used in the EARLIER phases (valid: train; test: train + valid)."""
    ev = splits[1] if phase == "valid" else splits[2]
    used = [splits[0]] + ([splits[1]] if phase == "test" else [])
    ku = np.concatenate([s[uf] for s in used])
    ki = np.concatenate([s[itf] for s in used])

    def group(u, i):
        o = np.argsort(u, kind="stable")
        u, i = u[o], i[o]
        s = np.flatnonzero(np.r_[True, u[1:] != u[:-1]]) if len(u) else np.zeros(0, np.int64)
        return u[s], np.split(i, s[1:])

    users, pos = group(ev[uf], ev[itf])
    hu, hl = group(ku, ki)
    hmap = dict(zip(hu.tolist(), hl))
    hist = [hmap.get(int(x), np.zeros(0, np.int64)) for x in users]
    return users, hist, pos
