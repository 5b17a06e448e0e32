"""Fused full-sort fair evaluation -- replaces the per-batch loop of recbole/trainer/trainer.py:505-512
(`full_sort_predict` -> mask -> `Collector.eval_batch_collect` -> `Evaluator.evaluate`) with one pass of
device kernels over ALL eval users: scoring + mask + streaming top-K (never materialising scores), hit
bits, positive scores, and the 12 metrics of properties/model/FOCF.yaml:29-30 accumulated on the device.

`EvalData` is the device-resident form of what `FullSortEvalDataLoader` (general_dataloader.py:161-253)
yields batch by batch: eval users, per-user history CSR (`used_ids - positives`, 201-207) and positives CSR.
`FullSortEvaluator.evaluate()` returns the reference's metric dict (same keys, same rounding) and can emit
the collector's `DataStruct` entries (collector.py:141-189) for cross-checking against metrics.py.

Multi-GPU: the item table is row-sharded over the ranks of a torch.distributed group;
each rank scores its shard, the per-shard top-K lists are all-gathered (NCCL) and merged under the total
order (score desc, item id asc); per-item x group statistics are computed by the shard that owns the item
and summed with one all-reduce.
"""
from collections import OrderedDict

import os

import numpy as np
import torch

from . import _lib, kernels

TOPK_ROWS = {"ndcg": 0, "recall": 1, "hit": 2, "mrr": 3}
FAIR_SLOTS = {"differentialfairness": 0, "valueunfairness": 1, "absoluteunfairness": 2, "underunfairness": 3,
              "overunfairness": 4, "nonparityunfairness": 5}
FAIR_KEYS = {  # metrics.py:855, 924, 1021, 1117, 1213, 1308
    "differentialfairness": "Differential Fairness of sensitive attribute {}",
    "valueunfairness": "Value Unfairness of sensitive attribute {}",
    "absoluteunfairness": "Absolute Unfairness of sensitive attribute {}",
    "underunfairness": "Underestimation Unfairness of sensitive attribute {}",
    "overunfairness": "Overestimation Unfairness of sensitive attribute {}",
    "nonparityunfairness": "NonParity Unfairness of sensitive attribute {}",
}
SUPPORTED = set(TOPK_ROWS) | set(FAIR_SLOTS) | {"giniindex", "popularitypercentage"}


class EvalData:
    """Device-resident full-sort eval split.  All ids are int32, offsets int64."""

    def __init__(self, users, hist_lists, pos_lists, sst_of_user, device):
        """users: eval user ids (ascending, like general_dataloader.py:188); hist_lists / pos_lists: one
        array of item ids per eval user; sst_of_user: {attr: array indexed by USER ID}."""
        users = np.asarray(users, dtype=np.int64)
        n = len(users)
        self.n = n
        self.device = device
        hist_sorted = [np.sort(np.asarray(h, dtype=np.int64)) for h in hist_lists]
        pos_orig = [np.asarray(p, dtype=np.int64) for p in pos_lists]
        pos_sorted = [np.sort(p) for p in pos_orig]
        hist_off = np.zeros(n + 1, np.int64)
        pos_off = np.zeros(n + 1, np.int64)
        hist_off[1:] = np.cumsum([len(h) for h in hist_sorted])
        pos_off[1:] = np.cumsum([len(p) for p in pos_orig])
        cat = lambda ls: np.concatenate(ls) if len(ls) and sum(map(len, ls)) else np.zeros(0, np.int64)
        self.n_pos = int(pos_off[-1])
        pos_row = np.repeat(np.arange(n), np.diff(pos_off))
        t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(device)
        self.users = t(users, torch.int32)
        self.hist_off, self.hist_items = t(hist_off, torch.int64), t(cat(hist_sorted), torch.int32)
        self.pos_off = t(pos_off, torch.int64)
        self.pos_items_sorted = t(cat(pos_sorted), torch.int32)   # for the hit-bit binary search
        self.pos_items = t(cat(pos_orig), torch.int32)            # emission order of data.positive_i
        self.pos_row = t(pos_row, torch.int64)
        self.pos_uid = t(users[pos_row], torch.int32)
        if self.hist_items.numel() == 0:
            self.hist_items = torch.zeros(1, dtype=torch.int32, device=device)
        # sensitive attributes: value per eval user, group id = rank among the values present in the
        # positives (np.unique(sst_value, return_inverse=True), metrics.py:941, 1327)
        self.sst_value, self.group_of_pos, self.n_groups = {}, {}, {}
        for attr, per_user in sst_of_user.items():
            vals = np.asarray(per_user)[users]
            uniq, inv = np.unique(vals[pos_row], return_inverse=True)
            self.sst_value[attr] = torch.as_tensor(vals)
            self.group_of_pos[attr] = t(inv, torch.int32)
            self.n_groups[attr] = len(uniq)

    @classmethod
    def from_device(cls, hist_uid, hist_iid, pos_uid, pos_iid, sst_of_user, n_users, n_items):
        """The same CSR layout built ON THE DEVICE from interaction columns (scale-out shapes): hist_* = the interactions
        already used in earlier phases (general_dataloader.py:201-207), pos_* = the phase's own interactions.  Eval users =
        users with at least one positive, ascending; every per-user list ascending by item id.
        sst_of_user: {attr: device tensor indexed by user id}."""
        self = cls.__new__(cls)
        dev = pos_uid.device
        shift = max(int(n_items - 1).bit_length(), 1)
        mask = (1 << shift) - 1

        def grouped(u, i):
            key = torch.sort((u.to(torch.int64) << shift) | i.to(torch.int64)).values
            return (key >> shift), (key & mask).to(torch.int32)

        pu, pi = grouped(pos_uid, pos_iid)
        pcnt = torch.bincount(pu, minlength=int(n_users))
        users = torch.nonzero(pcnt, as_tuple=False).view(-1)
        n = int(users.numel())
        self.n, self.device = n, dev
        self.users = users.to(torch.int32)
        self.pos_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        self.pos_off[1:] = torch.cumsum(pcnt[users], 0)
        self.pos_items = pi.contiguous()
        self.pos_items_sorted = self.pos_items
        self.n_pos = int(pi.numel())
        self.pos_row = torch.repeat_interleave(torch.arange(n, device=dev), pcnt[users])
        self.pos_uid = pu.to(torch.int32).contiguous()
        del pu
        hu, hi = grouped(hist_uid, hist_iid)
        keep = pcnt[hu] > 0
        hcnt = torch.bincount(hu[keep], minlength=int(n_users))
        self.hist_items = hi[keep].contiguous()
        del hu, hi, keep
        self.hist_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        self.hist_off[1:] = torch.cumsum(hcnt[users], 0)
        if self.hist_items.numel() == 0:
            self.hist_items = torch.zeros(1, dtype=torch.int32, device=dev)
        self.sst_value, self.group_of_pos, self.n_groups = {}, {}, {}
        for attr, per_user in sst_of_user.items():
            vals = per_user.to(dev)[users]
            uniq, inv = torch.unique(vals[self.pos_row], return_inverse=True)
            self.sst_value[attr] = vals.cpu()
            self.group_of_pos[attr] = inv.to(torch.int32).contiguous()
            self.n_groups[attr] = int(uniq.numel())
        return self

    @classmethod
    def from_reference_loader(cls, loader, sst_attr_list, device):
        """Adopt a reference FullSortEvalDataLoader (general_dataloader.py:173-199): uid_list,
        uid2history_item, uid2positive_item, user_df."""
        users = loader.uid_list.numpy()
        hist = [loader.uid2history_item[u].numpy() for u in users]
        pos = [loader.uid2positive_item[u].numpy() for u in users]
        feat = loader.dataset.get_user_feature()
        sst = {a: feat[a].numpy() for a in sst_attr_list}
        return cls(users, hist, pos, sst, device)


class FullSortEvaluator:
    def __init__(self, config, n_items, train_item_count=None, group=None):
        """train_item_count: {item id: #train interactions} (collector.py:91-93 data.count_items).
        group: optional torch.distributed process group for item-sharded evaluation."""
        self.config = config
        self.metrics = [m.lower() for m in config["metrics"]]
        bad = [m for m in self.metrics if m not in SUPPORTED]
        if bad:
            raise NotImplementedError(f"metrics {bad} are outside the fairness hot path (SURVEY.md section 2, row 10)")
        self.topk = list(config["topk"]) if isinstance(config["topk"], (list, tuple)) else [config["topk"]]
        self.K = max(self.topk)
        self.decimal = config["metric_decimal_place"] if config["metric_decimal_place"] is not None else 4
        self.sst_attr_list = list(config["sst_attr_list"])
        self.n_items = int(n_items)
        # score_mode: "tc" = the tensor-core contraction (tcgen05 + TMA, 3xTF32: scores within ~2e-6 of float32, ids equal to
        # the exact mode's outside float32-level near ties), "exact" = the bit-defined float32 FMA chain the oracle reproduces
        # (the pinned parity mode), "auto" (default) = tc whenever the shape allows it (d in {32, 64, 96, 128}, K <= 64)
        mode = config["score_mode"] or "auto"
        if mode == "auto":
            d = config["embedding_size"]
            mode = "tc" if (d is not None and d % 32 == 0 and d <= 128 and self.K <= 64) else "exact"
        self.score_mode = {"exact": _lib.SCORE_EXACT_FP32, "tc": _lib.SCORE_TC_3XTF32}[mode]
        self.score_mode_name = mode
        self.group = group
        # the three metric chains behind the scorer on forked streams (FR_EVAL_STREAMS=0: one stream, as before)
        self.branch_streams = os.environ.get("FR_EVAL_STREAMS", "1") != "0"
        self._side = None
        self._is_popular = None
        self._pop_host = self._popular_items(train_item_count, config["popularity_ratio"])
        self.last = None

    @staticmethod
    def _popular_items(count_items, ratio):
        """metrics.py:786-794"""
        if not count_items:
            return None
        if ratio is None or ratio <= 0:
            ratio = 0.1
        if ratio > 1:
            return np.array([i for i, c in count_items.items() if c >= ratio], dtype=np.int64)
        items = np.fromiter(count_items.keys(), dtype=np.int64, count=len(count_items))
        cnts = np.fromiter(count_items.values(), dtype=np.int64, count=len(count_items))
        order = np.lexsort((items, cnts))[::-1]          # (count, id) descending
        cut = max(int(len(items) * ratio), 1)
        return items[order[:cut]]

    def _popular_mask(self, device):
        if self._is_popular is None and self._pop_host is not None:
            m = torch.zeros(self.n_items, dtype=torch.uint8)
            m[torch.as_tensor(self._pop_host)] = 1
            self._is_popular = m.to(device)
        return self._is_popular

    # ------------------------------------------------------------------ the fused pass
    @torch.no_grad()
    def collect(self, U, I, data, max_rating, transform=_lib.TRANSFORM_CLAMP_DIV):
        """Runs the kernels; returns a dict of DEVICE tensors (no host sync)."""
        K = self.K
        fork = None
        # (only where the pass is a chain of short launches; behind a long scorer launch the side branches would just take
        # SMs and bandwidth from it)
        if self.group is None and self.branch_streams and data.n * self.n_items <= (1 << 27):
            self._popular_mask(U.device)            # (built on the current stream before anything forks)
            cur = torch.cuda.current_stream(U.device)
            if getattr(self, "_side", None) is None or self._side[0].device != U.device:
                self._side = (torch.cuda.Stream(device=U.device), torch.cuda.Stream(device=U.device))
            fork = (cur,) + self._side
            self._side[0].wait_stream(cur)
        if self.group is None:
            ids, sc = kernels.fullsort_topk(U, I, data.users, data.hist_off, data.hist_items, K, transform, max_rating,
                                            0, self.score_mode)
        else:
            import torch.distributed as dist
            world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
            lo, hi = shard_bounds(self.n_items, world, rank)
            ids_l, sc_l = kernels.fullsort_topk(U, I[lo:hi], data.users, data.hist_off, data.hist_items, K, transform,
                                                max_rating, lo, self.score_mode)
            ids_all = torch.empty((world,) + tuple(ids_l.shape), dtype=ids_l.dtype, device=ids_l.device)
            sc_all = torch.empty((world,) + tuple(sc_l.shape), dtype=sc_l.dtype, device=sc_l.device)
            dist.all_gather_into_tensor(ids_all, ids_l, group=self.group)
            dist.all_gather_into_tensor(sc_all, sc_l, group=self.group)
            ids, sc = kernels.topk_merge(ids_all, sc_all)
        # The pass behind the scorer is three independent chains of small launches: the positives' scores -> item x group
        # sums -> fairness metrics (needs nothing from the scorer), the top-K ids -> hit bits -> NDCG / Recall / Hit / MRR
        # sums, and the top-K ids -> recommendation histogram -> Gini / popularity.  Unsharded passes fork them onto side
        # streams (captured as parallel branches by collect_graphed): at the ML-1M shape the pass is latency-bound and the
        # fairness chain hides under the scorer.
        need = set(self.metrics)
        out = {"topk_id": ids, "topk_score": sc}

        def fair_branch():
            pos_score = kernels.pair_scores(U, I, data.pos_uid, data.pos_items, transform, max_rating)
            out["pos_score"] = pos_score
            if need & set(FAIR_SLOTS):
                out["fair"] = {}
                for ai, attr in enumerate(self.sst_attr_list):
                    G = data.n_groups[attr]
                    if self.group is None:
                        plan = self._ig_plan(data, None)
                        stats = kernels.item_group_stats_planned(plan["plan"], pos_score, data.group_of_pos[attr], self.n_items, G)
                    else:
                        import torch.distributed as dist
                        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
                        plan = self._ig_plan(data, shard_bounds(self.n_items, world, rank))
                        if plan["n"] > 0:     # the positives whose item this rank owns (a host constant of the plan: no sync)
                            if attr not in plan["group"]:
                                plan["group"][attr] = data.group_of_pos[attr].index_select(0, plan["idx"]).contiguous()
                            stats = kernels.item_group_stats_planned(plan["plan"], pos_score.index_select(0, plan["idx"]),
                                                                     plan["group"][attr], self.n_items, G)
                        else:
                            stats = torch.zeros((self.n_items, G, 2), dtype=torch.float64, device=U.device)
                        dist.all_reduce(stats, group=self.group)  # disjoint supports: x + 0, order-independent
                    out["fair"][attr] = kernels.fairness_metrics(stats)

        def rec_branch():
            if need & {"giniindex", "popularitypercentage"}:
                cnt, pop = kernels.rec_item_stats(ids, self.n_items, self._popular_mask(U.device))
                out["pop_hits"] = pop
                if "giniindex" in need:
                    out["gini"] = {k: kernels.gini_at_k(cnt, k, data.n) for k in self.topk}

        def hit_branch():
            rec_topk = kernels.hits(ids, data.pos_off, data.pos_items_sorted)
            out["rec_topk"] = rec_topk
            if need & set(TOPK_ROWS):
                out["topk_sums"] = kernels.topk_metric_sums(rec_topk)

        if fork is None:
            hit_branch()
            fair_branch()
            rec_branch()
        else:
            cur, s_fair, s_rec = fork
            with torch.cuda.stream(s_fair):         # (forked before the scorer was enqueued)
                fair_branch()
            s_rec.wait_stream(cur)                  # the ids exist on `cur` from here on
            with torch.cuda.stream(s_rec):
                rec_branch()
            hit_branch()
            cur.wait_stream(s_fair)
            cur.wait_stream(s_rec)
        self.last = out
        return out

    def _ig_plan(self, data, bounds):
        """The item-sorted view of the positive list (fr_item_group_plan) -- of the positives whose item lies in `bounds` for
        an item-sharded pass -- depends on the evaluation data only: built once per EvalData and kept on it."""
        cache = data.__dict__.setdefault("_ig_plans", {})
        key = (self.n_items, bounds)
        if key not in cache:
            if bounds is None:
                cache[key] = {"plan": kernels.item_group_plan(data.pos_items, self.n_items), "n": data.n_pos}
            else:
                lo, hi = bounds
                idx = torch.nonzero((data.pos_items >= lo) & (data.pos_items < hi), as_tuple=False).view(-1)
                n_own = int(idx.numel())
                plan = kernels.item_group_plan(data.pos_items.index_select(0, idx).contiguous(), self.n_items) if n_own else None
                cache[key] = {"plan": plan, "n": n_own, "idx": idx, "group": {}}
        return cache[key]

    def collect_graphed(self, U, I, data, max_rating, transform=_lib.TRANSFORM_CLAMP_DIV):
        """collect() captured once into a CUDA graph and replayed: the pass is ~35 small launches, so at the ML-1M
        shape the host-side launch cost would otherwise exceed the device time.  The graph is keyed on the table and
        eval-data addresses (weights change in place between evaluations, addresses do not)."""
        if self.group is not None:
            return self.collect(U, I, data, max_rating, transform)
        key = (U.data_ptr(), I.data_ptr(), id(data), float(max_rating), int(transform), data.users.data_ptr())
        g = getattr(self, "_graphs", None)
        if g is None:
            g = self._graphs = {}
        if key not in g:
            self.collect(U, I, data, max_rating, transform)      # eager warm-up (lazy module load, popularity mask)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self.collect(U, I, data, max_rating, transform)
            if len(g) >= 4:
                g.pop(next(iter(g)))
            g[key] = (graph, out)
        graph, out = g[key]
        graph.replay()
        self.last = out
        return out

    def evaluate(self, U, I, data, max_rating, transform=_lib.TRANSFORM_CLAMP_DIV):
        """evaluator.py:28-42: OrderedDict metric -> value, keys and rounding as the reference."""
        use_graph = self.config["cuda_graph"] is None or self.config["cuda_graph"]
        out = self.collect_graphed(U, I, data, max_rating, transform) if use_graph \
            else self.collect(U, I, data, max_rating, transform)
        return self.finalize(out, data)

    def finalize(self, out, data, rounded=True):
        n = data.n
        rnd = (lambda v: round(v, self.decimal)) if rounded else (lambda v: v)
        topk_sums = out["topk_sums"].cpu().numpy() if "topk_sums" in out else None
        pop = out["pop_hits"].cpu().numpy() if "pop_hits" in out else None
        fair = {a: v.cpu().numpy() for a, v in out.get("fair", {}).items()}
        res = OrderedDict()
        for m in self.metrics:
            if m in TOPK_ROWS:
                for k in self.topk:
                    res[f"{m}@{k}"] = rnd(float(topk_sums[TOPK_ROWS[m], k - 1] / n))
            elif m == "giniindex":
                for k in self.topk:
                    res[f"giniindex@{k}"] = rnd(float(out["gini"][k].item()))
            elif m == "popularitypercentage":
                for k in self.topk:
                    res[f"popularitypercentage@{k}"] = rnd(float(pop[:k].sum() / (n * k)))
            else:
                # DifferentialFairness / NonParity report every attribute, the other four only the first
                attrs = self.sst_attr_list if m in ("differentialfairness", "nonparityunfairness") \
                    else self.sst_attr_list[:1]
                for attr in attrs:
                    v = float(fair[attr][FAIR_SLOTS[m]])
                    if np.isnan(v):
                        if m == "nonparityunfairness":
                            raise ValueError(f"there is only one value for {attr} sensitive attribute")
                        raise ValueError("sensitive attribute must be binary")   # metrics.py:951-952
                    res[FAIR_KEYS[m].format(attr)] = rnd(v)
        return res

    def data_struct(self, data):
        """The collector's DataStruct entries (collector.py:141-189) of the last collect(), on the host."""
        out = self.last
        st = {
            "rec.items": out["topk_id"].cpu().to(torch.int64),
            "rec.topk": out["rec_topk"].cpu(),
            "rec.positive_score": out["pos_score"].cpu(),
            "data.positive_i": data.pos_items.cpu().to(torch.int64),
        }
        for attr in self.sst_attr_list:
            st["data." + attr] = data.sst_value[attr][data.pos_row.cpu()]
        return st


def shard_bounds(n_items, world, rank):
    """contiguous item-id ranges so that the tie rule (lowest id) survives the merge"""
    per = (n_items + world - 1) // world
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)
