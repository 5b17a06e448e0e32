"""Sampled-negative ("uni100") ranking evaluation -- replaces NegSampleEvalDataLoader batches
(recbole/data/dataloader/general_dataloader.py:68-158) + Trainer._neg_sample_batch_eval (recbole/trainer/trainer.py:441-456)
+ Collector.eval_batch_collect (recbole/evaluator/collector.py:131-205) + the mode-independent metrics with one pass of
device kernels over ALL eval users.  `uni100` is the default evaluation mode of all eight model YAMLs and the only
evaluation the reference defines for the PFCN family.

The reference scatters every user's candidate scores into a dense [users, n_items] row of -inf and runs torch.topk on
it; here the dense rows never exist: candidates live in a CSR (positives first, then the sampled negatives, the
dataloader's own order), their scores come from the model's `predict` (any model) or the pair-score kernel (dot-product
models), and `fr_sampled_topk` extracts the K best per user in the canonical order (score desc, item id asc).

Metrics: all 12 of the model YAMLs.  NDCG / Recall / Hit / MRR, GiniIndex, PopularityPercentage, DifferentialFairness and
NonParityUnfairness are defined as in full mode.  Value / Absolute / Under / Over unfairness follow the sampled mode of
metrics.py:935-978 (and siblings): the item set is the union of the positives' items and of the items of each positive's
FIRST sampled negative (`interaction[item][P:2P]`, collector.py:190-199), a negative contributing its score (for the
positive's user) and a count, but no "true" mass.

One deliberate difference from the reference, only visible when its dataloader packs SEVERAL users into one batch (the
default eval_batch_size does): its collector indexes the batch interaction as if it held one user -- `interaction[sst][:P]`
is mostly the first user's attribute (collector.py:203-205) and `interaction[item][P:2P]` straddles the users' blocks, so
the negative scores it reads are -inf entries of other users' rows and the four unfairness values come out inf / NaN.  Here
every positive carries its OWN user's attribute and is paired with ITS first negative: with one user per batch this IS the
reference (tests/golden/uni_eval_uni100.npz, uni_eval_uni100_unfair.npz), with several it is the value the reference means."""
from collections import OrderedDict

import numpy as np
import torch

from . import _lib, kernels
from ._lib import check, load, ptr, stream_ptr
from .evaluator import FAIR_KEYS, FAIR_SLOTS, TOPK_ROWS, FullSortEvaluator

UNFAIR4 = ("valueunfairness", "absoluteunfairness", "underunfairness", "overunfairness")


def sample_negatives(pos_lists, used_lists, n_items, neg_num, rng):
    """Sampler.sample_by_key_ids (recbole/sampler/sampler.py:145-197) in effect: for every positive `neg_num` items drawn
    uniformly from the items the user has NOT used (ids 1..n_items-1), laid out like the dataloader does
    (`[j * p + k]` = j-th negative of positive k, abstract_dataloader.py:190-198)."""
    out = []
    draw = getattr(rng, "integers", None) or rng.randint      # np.random.Generator, or the seeded np.random module
    for pos, used in zip(pos_lists, used_lists):
        banned = np.zeros(n_items, bool)
        banned[0] = True
        banned[np.asarray(used, np.int64)] = True
        banned[np.asarray(pos, np.int64)] = True
        need = neg_num * len(pos)
        got = np.zeros(0, np.int64)
        while len(got) < need:            # rejection sampling, like the reference
            c = draw(1, n_items, size=2 * (need - len(got)) + 8)
            got = np.concatenate([got, c[~banned[c]]])
        out.append(got[:need])
    return out


class AliasSampler:
    """The reference's popularity-biased item sampler (sampler.py:72-120): Walker alias table over the item ids of ALL
    interactions (train, valid, test rows in that order, sampler.py:234-238), keys in order of first appearance, the two
    work queues first-in first-out -- and `sampling(n)` = ONE `np.random.randint(0, n_keys, n)` followed by ONE
    `np.random.random(n)` on numpy's global RNG, so that a seed gives the reference's draws."""

    def __init__(self, item_ids):
        import pandas as pd
        codes, keys = pd.factorize(np.asarray(item_ids, np.int64))
        cnt = np.bincount(codes, minlength=len(keys))
        prob = (cnt / len(codes) * len(keys)).tolist()
        alias = [-1] * len(keys)
        large = [i for i, p in enumerate(prob) if p > 1]
        small = [i for i, p in enumerate(prob) if p < 1]
        a = b = 0                                   # queue heads (the reference pops from the front of two lists)
        while a < len(large) and b < len(small):
            l, s_ = large[a], small[b]
            a, b = a + 1, b + 1
            alias[s_] = l
            prob[l] = prob[l] - (1 - prob[s_])
            if prob[l] < 1:
                small.append(l)
            elif prob[l] > 1:
                large.append(l)
        self.keys = np.asarray(keys, np.int64)
        self.prob = np.asarray(prob, np.float64)
        self.alias = np.where(np.asarray(alias) >= 0, self.keys[np.maximum(alias, 0)], -1)      # item ids (-1: never read)

    def sampling(self, n):
        idx = np.random.randint(0, len(self.keys), n)
        p = np.random.random(n)
        return np.where(self.prob[idx] > p, self.keys[idx], self.alias[idx]).astype(np.int64)


def uniform_sampling(n_items):
    """sampler.py:240-241"""
    return lambda n: np.random.randint(1, n_items, n)


def sample_negatives_reference(pos_lists, used_lists, n_items, neg_num, sampling=None):
    """The reference's own draws (sampler.py:159-175 through NegSampleEvalDataLoader._next_batch_data,
    general_dataloader.py:128-140): users in evaluation order, for each ONE call `np.random.randint(1, n_items, p * neg_num)`
    on numpy's global RNG, then the entries that hit a used item (train + the evaluated split, positives included) are
    redrawn together until none is left.  Same calls on the same stream => the same negatives as the reference for a seed
    (oracle/fuzz_loaders.py checks that against the live reference).  The reference draws them anew at EVERY evaluation;
    `ResamplingEvalSource` does the same."""
    out = []
    sampling = sampling or uniform_sampling(n_items)          # `pop<N>` modes pass AliasSampler.sampling
    for pos, used in zip(pos_lists, used_lists):
        banned = np.zeros(n_items, bool)
        banned[np.asarray(used, np.int64)] = True
        banned[np.asarray(pos, np.int64)] = True
        value = sampling(neg_num * len(pos))
        check = np.flatnonzero(banned[value])
        while len(check) > 0:
            redraw = sampling(len(check))
            value[check] = redraw
            check = check[banned[redraw]]
        out.append(value.astype(np.int64))
    return out


class ResamplingEvalSource:
    """Evaluation split of the `uni<N>` mode whose negatives are drawn again at every evaluation, like the reference's
    NegSampleEvalDataLoader does while it iterates: trainers call `.resample()` and evaluate the SampledEvalData it returns."""

    def __init__(self, users, pos_lists, used_lists, sst_of_user, n_items, neg_num, device, sampling=None):
        self.users, self.pos, self.used, self.sst = users, pos_lists, used_lists, sst_of_user
        self.n_items, self.neg_num, self.device = int(n_items), int(neg_num), device
        self.sampling = sampling                  # None: uniform (`uni<N>`); AliasSampler.sampling for `pop<N>`

    def __len__(self):
        return len(self.users)

    def resample(self):
        neg = sample_negatives_reference(self.pos, self.used, self.n_items, self.neg_num, self.sampling)
        return SampledEvalData(self.users, self.pos, neg, self.sst, self.device)

    def resample_tiled(self, copies):
        """ONE draw of negatives, the user list repeated `copies` times (copy-major): the layout in which the reference's
        PFCN validation scores every batch under each attribute subset and collects everything into one struct
        (trainer.py:1010-1023).  Returns (data, candidate rows per copy)."""
        neg = sample_negatives_reference(self.pos, self.used, self.n_items, self.neg_num, self.sampling)
        data = SampledEvalData(np.tile(np.asarray(self.users, np.int64), copies), list(self.pos) * copies, neg * copies,
                               self.sst, self.device)
        return data, int(data.cand_uid.numel()) // copies


class SampledEvalData:
    """Device-resident candidate lists: per eval user its positives followed by its sampled negatives."""

    def __init__(self, users, pos_lists, neg_lists, sst_of_user, device):
        users = np.asarray(users, np.int64)
        n = len(users)
        self.n, self.device = n, device
        n_pos = np.array([len(p) for p in pos_lists], np.int64)
        cand = [np.concatenate([np.asarray(p, np.int64), np.asarray(q, np.int64)]) for p, q in zip(pos_lists, neg_lists)]
        cand_off = np.zeros(n + 1, np.int64)
        cand_off[1:] = np.cumsum([len(c) for c in cand])
        t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(device)
        self.users = t(users, torch.int32)
        self.cand_off = t(cand_off, torch.int64)
        self.cand_items = t(np.concatenate(cand), torch.int32)
        self.cand_uid = t(np.repeat(users, np.diff(cand_off)), torch.int32)
        self.n_pos_of_user = t(n_pos, torch.int32)
        # positions of the positives inside the candidate arrays (user-major, dataloader order)
        pos_idx = np.concatenate([cand_off[k] + np.arange(n_pos[k]) for k in range(n)]) if n else np.zeros(0, np.int64)
        self.pos_idx = t(pos_idx, torch.int64)
        # ... and of each positive's FIRST sampled negative (negatives are laid out draw-major: [j * P + k] = j-th negative of
        # positive k, general_dataloader.py / abstract_dataloader.py `times` layout), where a user has negatives at all
        has_neg = np.array([len(q) >= len(p) > 0 for p, q in zip(pos_lists, neg_lists)], bool)
        neg_idx = np.concatenate([cand_off[k] + n_pos[k] + np.arange(n_pos[k]) for k in range(n)]) if n else np.zeros(0, np.int64)
        self.first_neg_idx = t(neg_idx, torch.int64) if bool(has_neg.all()) else None
        self.pos_items = t(np.concatenate([np.asarray(p, np.int64) for p in pos_lists]), torch.int32)
        self.n_pos = int(n_pos.sum())
        pos_row = np.repeat(np.arange(n), n_pos)
        self.pos_row = t(pos_row, torch.int64)
        self.sst_value, self.group_of_pos, self.n_groups = {}, {}, {}
        for attr, per_user in sst_of_user.items():
            vals = np.asarray(per_user)[users]
            uniq, inv = np.unique(vals[pos_row], return_inverse=True)
            self.sst_value[attr] = torch.as_tensor(vals)
            self.group_of_pos[attr] = t(inv, torch.int32)
            self.n_groups[attr] = len(uniq)


def sampled_topk(data, scores, K, n_items):
    """fr_sampled_topk -> (topk_id int32 [n,K], topk_score f32 [n,K], rec_topk int32 [n,K+1])"""
    lib = load()
    dev = scores.device
    ids = torch.empty((data.n, K), dtype=torch.int32, device=dev)
    sc = torch.empty((data.n, K), dtype=torch.float32, device=dev)
    rec = torch.empty((data.n, K + 1), dtype=torch.int32, device=dev)
    check(lib.fr_sampled_topk(ptr(data.cand_off), ptr(data.cand_items), ptr(scores.contiguous()), ptr(data.n_pos_of_user),
                              data.n, int(K), int(n_items), ptr(ids), ptr(sc), ptr(rec), stream_ptr()), "fr_sampled_topk")
    return ids, sc, rec


class SampledEvaluator(FullSortEvaluator):
    """evaluate(score_fn, data) -> the reference's metric dict (same keys, same rounding).  score_fn(uid int32 [C],
    iid int32 [C]) -> float32 [C] scores of the candidate pairs; `dot_scorer` builds one for dot-product models."""

    def __init__(self, config, n_items, train_item_count=None):
        super().__init__(config, n_items, train_item_count)

    @staticmethod
    def dot_scorer(U, I, max_rating=None, transform=None):
        if transform is None:
            transform = _lib.TRANSFORM_CLAMP_DIV if max_rating is not None else _lib.TRANSFORM_NONE
        return lambda uid, iid: kernels.pair_scores(U, I, uid, iid, transform, 1.0 if max_rating is None else max_rating)

    @torch.no_grad()
    def collect(self, score_fn, data):
        C, chunk = data.cand_uid.numel(), int(self.config["sampled_chunk_rows"] or (1 << 24))
        if C <= chunk:
            scores = score_fn(data.cand_uid, data.cand_items).view(-1).to(torch.float32)
        else:       # bound the activations of MLP scorers (PFCN predict over ~1e7 candidate rows); rows are independent
            scores = torch.cat([score_fn(data.cand_uid[a:a + chunk], data.cand_items[a:a + chunk]).view(-1).to(torch.float32)
                                for a in range(0, C, chunk)])
        return self.collect_scores(scores, data)

    @torch.no_grad()
    def collect_scores(self, scores, data):
        """the pass over given candidate scores (float32 [C], candidate order of `data`)"""
        ids, sc, rec_topk = sampled_topk(data, scores, self.K, self.n_items)
        pos_score = scores[data.pos_idx].contiguous()
        out = {"topk_id": ids, "topk_score": sc, "rec_topk": rec_topk, "pos_score": pos_score}
        need = set(self.metrics)
        if need & set(TOPK_ROWS):
            out["topk_sums"] = kernels.topk_metric_sums(rec_topk)
        if need & {"giniindex", "popularitypercentage"}:
            cnt, pop = kernels.rec_item_stats(ids, self.n_items, self._popular_mask(scores.device))
            out["pop_hits"] = pop
            if "giniindex" in need:
                out["gini"] = {k: kernels.gini_at_k(cnt, k, data.n) for k in self.topk}
        if need & set(FAIR_SLOTS):
            out["fair"] = {}
            for ai, attr in enumerate(self.sst_attr_list):
                grp, G = data.group_of_pos[attr], data.n_groups[attr]
                stats = kernels.item_group_stats(data.pos_items, pos_score, grp, self.n_items, G)
                fair = kernels.fairness_metrics(stats)
                if ai == 0 and need & set(UNFAIR4) and G == 2:
                    # metrics.py:935-978 with mode != 'full': positives + each positive's first negative (see module docstring)
                    if data.first_neg_idx is None:
                        raise ValueError("the sampled-mode unfairness metrics need at least one negative per positive")
                    neg_score = scores[data.first_neg_idx]
                    neg_items = data.cand_items[data.first_neg_idx]
                    both = kernels.item_group_stats(torch.cat([data.pos_items, neg_items]).contiguous(),
                                                    torch.cat([pos_score, neg_score]).contiguous(),
                                                    torch.cat([grp, grp]).contiguous(), self.n_items, 2)
                    fair[1:5] = kernels.unfairness_sampled(both, stats)[:4]
                out["fair"][attr] = fair
        self.last = out
        return out

    def evaluate(self, score_fn, data):
        return self.finalize(self.collect(score_fn, data), data)


__all__ = ["SampledEvalData", "SampledEvaluator", "ResamplingEvalSource", "AliasSampler", "sample_negatives", "sample_negatives_reference",
           "sampled_topk", "OrderedDict", "FAIR_KEYS"]
