"""GPU parity of the fused full-sort evaluation against (1) the reference's golden fixtures and (2) the
oracle (bit-exact: indices, hit bits, counts, and -- in exact scoring mode -- the scores themselves)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import fullsort_oracle as fs
from oracle import metrics_oracle as mo

pytestmark = pytest.mark.gpu
RTOL = 1e-5
HERE = os.path.dirname(__file__)
EVAL = sorted(glob.glob(os.path.join(HERE, "golden", "focf_eval_*.npz")))


def build(U, I, users, hist_off, hist_items, pos_off, pos_items, sst_of_user, topk, count_items=None, **cfg):
    import recbole_fairrec_b200 as pkg
    dev = torch.device("cuda")
    hist = [hist_items[hist_off[r]:hist_off[r + 1]] for r in range(len(users))]
    pos = [pos_items[pos_off[r]:pos_off[r + 1]] for r in range(len(users))]
    data = pkg.EvalData(users, hist, pos, {"gender": sst_of_user}, dev)
    conf = pkg.Config(topk=list(topk), metric_decimal_place=12, device=dev, **cfg)
    ev = pkg.FullSortEvaluator(conf, I.shape[0], count_items)
    return data, ev, torch.from_numpy(U).cuda(), torch.from_numpy(I).cuda()


@pytest.mark.parametrize("path", EVAL, ids=[os.path.basename(p)[10:-4] for p in EVAL])
def test_golden_fullsort_eval(path):
    g = np.load(path)
    K, users = int(g["K"]), g["eval_users"]
    count_items = {int(i): int(c) for i, c in g["train_count_items"]}
    data, ev, U, I = build(g["U"], g["I"], users, g["hist_off"], g["hist_items"], g["pos_off"], g["pos_items"],
                           g["sst_of_user"], g["topk"], count_items)
    res = ev.evaluate(U, I, data, float(g["max_rating"]))
    st = ev.data_struct(data)
    # ---- against the oracle: bit-exact everything (canonical tie order included)
    scores = fs.mask_history(fs.full_sort_scores(g["U"], g["I"], users, float(g["max_rating"])),
                             g["hist_off"], g["hist_items"])
    ost = fs.collect(scores, K, g["pos_off"], g["pos_items"], g["sst_of_user"][users])
    np.testing.assert_array_equal(st["rec.items"].numpy(), ost["rec.items"])
    np.testing.assert_array_equal(st["rec.topk"].numpy(), ost["rec.topk"])
    np.testing.assert_array_equal(st["rec.positive_score"].numpy(), ost["rec.positive_score"])
    np.testing.assert_array_equal(st["data.positive_i"].numpy(), ost["data.positive_i"])
    np.testing.assert_array_equal(ev.last["topk_score"].cpu().numpy(), fs.topk_canonical(scores, K)[1])
    # ---- against the reference's own collector output where torch.topk is well defined (no near ties)
    _, vals = fs.topk_canonical(scores, K + 1)
    clean = np.all(np.abs(np.diff(vals, axis=1)) > 1e-6, axis=1)
    np.testing.assert_array_equal(st["rec.items"].numpy()[clean], g["rec_items"][clean])
    np.testing.assert_array_equal(st["rec.topk"].numpy()[clean], g["rec_topk"][clean])
    np.testing.assert_allclose(st["rec.positive_score"].numpy(), g["rec_positive_score"], rtol=1e-5, atol=1e-7)
    np.testing.assert_array_equal(st["data.gender"].numpy(), g["data_sst"])
    # ---- metrics: device accumulation vs the oracle on the same struct, and vs the reference when comparable
    omet = mo.evaluate(ost, [int(k) for k in g["topk"]], g["I"].shape[0], count_items, 0.1)
    assert list(res.keys()) == list(omet.keys()) == [str(k) for k in g["metric_names"]]
    for k in res:
        assert abs(res[k] - omet[k]) <= RTOL * max(abs(omet[k]), 1e-12) + 1e-12, (k, res[k], omet[k])
    if clean.all():
        for (k, v), ref in zip(res.items(), g["metric_values"]):
            assert abs(v - ref) <= RTOL * max(abs(ref), 1e-12) + 1e-12, (k, v, ref)


def random_eval_case(seed, n_users, n_items, d, n_eval, scale=0.3):
    rng = np.random.default_rng(seed)
    U = (rng.standard_normal((n_users, d)) * scale).astype(np.float32)
    I = (rng.standard_normal((n_items, d)) * scale).astype(np.float32)
    users = np.sort(rng.choice(np.arange(1, n_users), n_eval, replace=False))
    hist, pos = [], []
    for _ in users:
        used = rng.choice(np.arange(1, n_items), int(rng.integers(2, min(200, n_items - 1))), replace=False)
        npos = int(rng.integers(1, min(20, len(used))))
        pos.append(used[:npos])
        hist.append(used[npos:])
    hist_off = np.r_[0, np.cumsum([len(h) for h in hist])]
    pos_off = np.r_[0, np.cumsum([len(p) for p in pos])]
    sst = rng.integers(1, 3, n_users)
    return U, I, users, hist_off, np.concatenate(hist), pos_off, np.concatenate(pos), sst


@pytest.mark.parametrize("n_users,n_items,d,n_eval,K", [(700, 1683, 64, 650, 10), (3000, 3707, 64, 2900, 10),
                                                       (500, 9000, 128, 333, 20), (200, 130, 16, 150, 5),
                                                       (300, 1000, 96, 257, 50)])
def test_oracle_fullsort_random(n_users, n_items, d, n_eval, K):
    U, I, users, ho, hi, po, pi, sst = random_eval_case(n_items + d, n_users, n_items, d, n_eval)
    data, ev, Ud, Id = build(U, I, users, ho, hi, po, pi, sst, [K])
    ev.collect(Ud, Id, data, 5.0)
    st = ev.data_struct(data)
    scores = fs.mask_history(fs.full_sort_scores(U, I, users, 5.0), ho, hi)
    ost = fs.collect(scores, K, po, pi, sst[users])
    np.testing.assert_array_equal(st["rec.items"].numpy(), ost["rec.items"])
    np.testing.assert_array_equal(st["rec.topk"].numpy(), ost["rec.topk"])
    np.testing.assert_array_equal(st["rec.positive_score"].numpy(), ost["rec.positive_score"])
    np.testing.assert_array_equal(ev.last["topk_score"].cpu().numpy(), fs.topk_canonical(scores, K)[1])
    res = ev.finalize(ev.last, data, rounded=False)
    # metrics that do not need train counts
    for name, fn in (("ndcg", mo.ndcg), ("recall", mo.recall)):
        pos_index, pos_len = ost["rec.topk"][:, :K].astype(bool), ost["rec.topk"][:, K]
        want = fn(pos_index, pos_len).mean(axis=0)[K - 1]
        assert abs(res[f"{name}@{K}"] - want) <= RTOL * max(abs(want), 1e-12)
    want = mo.gini(ost["rec.items"], n_items)
    assert abs(res[f"giniindex@{K}"] - want) <= 1e-12 + RTOL * abs(want)
    for key, fn in (("Value Unfairness", mo.value_unfairness), ("Absolute Unfairness", mo.absolute_unfairness),
                    ("Underestimation Unfairness", mo.under_unfairness),
                    ("Overestimation Unfairness", mo.over_unfairness),
                    ("Differential Fairness", mo.differential_fairness)):
        want = fn(ost["rec.positive_score"], ost["data.positive_i"], ost["data.sst"])
        got = res[f"{key} of sensitive attribute gender"]
        assert abs(got - want) <= RTOL * max(abs(want), 1e-12) + 1e-12, (key, got, want)
    want = mo.nonparity(ost["rec.positive_score"], ost["data.sst"])
    got = res["NonParity Unfairness of sensitive attribute gender"]
    assert abs(got - want) <= 1e-5 * max(abs(want), 1e-3), (got, want)   # reference accumulates this one in fp32


def test_item_sharded_merge_equals_unsharded():
    """The multi-GPU protocol on one device: P contiguous item shards -> local top-K -> fr_topk_merge."""
    from recbole_fairrec_b200 import _lib, kernels
    from recbole_fairrec_b200.evaluator import shard_bounds
    U, I, users, ho, hi, po, pi, sst = random_eval_case(3, 900, 2500, 64, 800)
    data, ev, Ud, Id = build(U, I, users, ho, hi, po, pi, sst, [10])
    full_ids, full_sc = kernels.fullsort_topk(Ud, Id, data.users, data.hist_off, data.hist_items, 10,
                                              _lib.TRANSFORM_CLAMP_DIV, 5.0)
    for P in (2, 3, 8):
        ids, scs = [], []
        for r in range(P):
            lo, hi_ = shard_bounds(2500, P, r)
            a, b = kernels.fullsort_topk(Ud, Id[lo:hi_].contiguous(), data.users, data.hist_off, data.hist_items, 10,
                                         _lib.TRANSFORM_CLAMP_DIV, 5.0, item_base=lo)
            ids.append(a)
            scs.append(b)
        mi, ms = kernels.topk_merge(torch.stack(ids).contiguous(), torch.stack(scs).contiguous())
        assert torch.equal(mi, full_ids) and torch.equal(ms, full_sc)


def test_clamp_ties_lowest_item_id_wins():
    """clamp(0,max) creates exact ties at 0.0 and 1.0 (SURVEY.md hard part 3): canonical order is id ascending."""
    from recbole_fairrec_b200 import _lib, kernels
    rng = np.random.default_rng(0)
    U = (rng.standard_normal((40, 16)) * 3).astype(np.float32)      # large dots: most scores clamp to 0 or 1
    I = (rng.standard_normal((500, 16)) * 3).astype(np.float32)
    users = np.arange(1, 40)
    ho = np.zeros(40, np.int64)
    ids, sc = kernels.fullsort_topk(torch.from_numpy(U).cuda(), torch.from_numpy(I).cuda(),
                                    torch.from_numpy(users.astype(np.int32)).cuda(), torch.from_numpy(ho).cuda(),
                                    torch.zeros(1, dtype=torch.int32, device="cuda"), 10, _lib.TRANSFORM_CLAMP_DIV, 5.0)
    scores = fs.mask_history(fs.full_sort_scores(U, I, users, 5.0), ho, np.zeros(0, np.int64))
    oi, os_ = fs.topk_canonical(scores, 10)
    np.testing.assert_array_equal(ids.cpu().numpy(), oi)
    np.testing.assert_array_equal(sc.cpu().numpy(), os_)
    assert (os_ == 1.0).sum() > 100   # the case really is tie-heavy


def test_kat2_metric_kernels():
    """SURVEY.md section 4 KAT-2 through the metric kernels."""
    from recbole_fairrec_b200 import kernels
    rec_topk = torch.tensor([[1, 0, 1, 2], [0, 0, 0, 1], [0, 1, 0, 3]], dtype=torch.int32, device="cuda")
    sums = kernels.topk_metric_sums(rec_topk).cpu().numpy() / 3
    np.testing.assert_allclose(sums[0, 2], np.mean([0.91972079, 0, 0.29608191]), atol=1e-8)
    np.testing.assert_allclose(sums[1, 2], np.mean([1, 0, 1 / 3]))
    np.testing.assert_allclose(sums[2, 2], np.mean([1, 0, 1]))
    np.testing.assert_allclose(sums[3, 2], np.mean([1, 0, 0.5]))
    score = torch.tensor([0.9, 0.2, 0.4, 1.0, 0.0, 0.6, 0.7], device="cuda")
    item = torch.tensor([1, 1, 2, 2, 2, 3, 3], dtype=torch.int32, device="cuda")
    grp = torch.tensor([0, 1, 0, 1, 1, 0, 0], dtype=torch.int32, device="cuda")
    out = kernels.fairness_metrics(kernels.item_group_stats(item, score, grp, 5, 2)).cpu().numpy()
    np.testing.assert_allclose(out[0], 0.507108, rtol=1e-5)
    np.testing.assert_allclose(out[1:4], 0.3833292371278623, rtol=1e-6)
    assert out[4] == 0.0 and out[6] == 3
    np.testing.assert_allclose(out[5], 0.25, rtol=1e-6)
    items = torch.tensor([[1, 2, 3], [1, 2, 4], [1, 3, 2]], dtype=torch.int32, device="cuda")
    pop = torch.zeros(5, dtype=torch.uint8, device="cuda")
    pop[[1, 3]] = 1   # counts {1:10,2:5,3:5,4:1}, ratio .5 -> top 2 by (count,id) desc = items 1 and 3
    cnt, hits = kernels.rec_item_stats(items, 5, pop)
    np.testing.assert_allclose(kernels.gini_at_k(cnt, 3, 3).item(), 0.35555555555, rtol=1e-9)
    h = hits.cpu().numpy()
    np.testing.assert_allclose([h[:k].sum() / (3 * k) for k in (1, 2, 3)], [1, 2 / 3, 5 / 9])


@pytest.mark.parametrize("n_users", [5000, 20000])
@pytest.mark.parametrize("n_items", [1, 2, 777, 4096, 8192, 8193, 30000])
def test_gini_small_and_radix_paths_vs_numpy(n_items, n_users):
    """fr_gini_at_k: a one-CTA histogram over the count values when there are fewer than 8192 users, else a one-CTA
    shared-memory sort up to 8192 items, else the radix sort -- the same integer sum every way (metrics.py:644-661 on the
    summed rank rows)."""
    from recbole_fairrec_b200 import kernels
    rng = np.random.default_rng(n_items)
    K = 3
    cnt = rng.integers(0, 40, size=(K, n_items)).astype(np.int32)
    cnt[:, : max(1, n_items // 50)] = n_users // K   # a few very popular items (the largest admissible count)
    cnt[:, rng.random(n_items) < 0.3] = 0
    for k in (1, 3):
        c = np.sort(cnt[:k].sum(axis=0).astype(np.int64))
        want = float(((2 * np.arange(1, n_items + 1) - n_items - 1) * c).sum()) / float(n_users * k) / float(n_items)
        got = kernels.gini_at_k(torch.from_numpy(cnt).cuda(), k, n_users).item()
        assert got == want, (n_items, k, got, want)


def test_item_group_stats_planned_equals_unplanned():
    """fr_item_group_plan + fr_item_group_stats_planned (the sort of the positives hoisted out of the pass) = fr_item_group_stats,
    bit for bit, and the plan is reusable with new scores."""
    from recbole_fairrec_b200 import kernels
    g = torch.Generator(device="cuda").manual_seed(5)
    n_pos, n_items, G = 50_000, 3707, 3
    item = torch.randint(1, n_items, (n_pos,), device="cuda", generator=g, dtype=torch.int32)
    grp = torch.randint(0, G, (n_pos,), device="cuda", generator=g, dtype=torch.int32)
    plan = kernels.item_group_plan(item, n_items)
    for _ in range(2):
        score = torch.rand(n_pos, device="cuda", generator=g)
        a = kernels.item_group_stats(item, score, grp, n_items, G)
        b = kernels.item_group_stats_planned(plan, score, grp, n_items, G)
        assert torch.equal(a, b)


@pytest.mark.parametrize("n_users,n_items,d,n_eval,K", [(700, 1683, 64, 650, 10), (3000, 3707, 64, 2900, 10),
                                                       (500, 9000, 128, 333, 20), (400, 1000, 32, 257, 5),
                                                       (300, 2100, 96, 129, 10), (350, 1500, 64, 300, 16),
                                                       (350, 1500, 64, 300, 12)])
def test_tc_scorer_matches_exact_scorer(n_users, n_items, d, n_eval, K):
    """FR_SCORE_TC_3XTF32 (tcgen05 + TMA, 3xTF32) against the bit-defined exact scorer: scores within fp32-level
    tolerance; ids identical wherever the exact top-(K+1) is separated by more than that tolerance (near-tie
    protocol of SURVEY.md hard part 3), same score multiset elsewhere."""
    from recbole_fairrec_b200 import _lib, kernels
    U, I, users, ho, hi, po, pi, sst = random_eval_case(n_items + d + 1, n_users, n_items, d, n_eval)
    data, ev, Ud, Id = build(U, I, users, ho, hi, po, pi, sst, [K])
    args = (Ud, Id, data.users, data.hist_off, data.hist_items)
    ids_e, sc_e = kernels.fullsort_topk(*args, K + 1, _lib.TRANSFORM_CLAMP_DIV, 5.0)
    ids_t, sc_t = kernels.fullsort_topk(*args, K, _lib.TRANSFORM_CLAMP_DIV, 5.0, 0, _lib.SCORE_TC_3XTF32)
    torch.cuda.synchronize()
    sc_e, sc_t, ids_e, ids_t = sc_e.cpu().numpy(), sc_t.cpu().numpy(), ids_e.cpu().numpy(), ids_t.cpu().numpy()
    hist = [set(hi[ho[r]:ho[r + 1]].tolist()) for r in range(len(users))]

    def check(ids_e, sc_e, ids_t, sc_t, transform, max_rating, tol):
        """every row, clean or not: the K scores agree position by position (same multiset); each id the tensor-core scorer
        returns is a legal candidate whose EXACT score is the one reported (within tol) and reaches the exact K-th score
        (within tol), no id twice; rows whose exact top-(K+1) is separated by more than 2*tol must match id by id"""
        np.testing.assert_allclose(sc_t, sc_e[:, :K], rtol=0, atol=tol)
        n = ids_t.shape[0]
        uid = torch.from_numpy(np.repeat(users.astype(np.int32), K)).cuda()
        exact_of_picks = kernels.pair_scores(Ud, Id, uid, torch.from_numpy(ids_t.reshape(-1).astype(np.int32)).cuda(),
                                             transform, max_rating).cpu().numpy().reshape(n, K)
        np.testing.assert_allclose(exact_of_picks, sc_t, rtol=0, atol=tol)
        assert np.all(exact_of_picks >= sc_e[:, K - 1:K] - tol)
        for r in range(n):
            row = ids_t[r].tolist()
            assert len(set(row)) == K and 0 not in row and not (set(row) & hist[r]), r
        clean = np.all(np.abs(np.diff(sc_e, axis=1)) > 2 * tol, axis=1)
        assert clean.any()
        np.testing.assert_array_equal(ids_t[clean], ids_e[clean, :K])

    check(ids_e, sc_e, ids_t, sc_t, _lib.TRANSFORM_CLAMP_DIV, 5.0, 4e-6)
    # raw-dot variant (no clamp): exercises the filter with negative scores
    ids_e, sc_e = kernels.fullsort_topk(*args, K + 1, _lib.TRANSFORM_NONE, 1.0)
    ids_t, sc_t = kernels.fullsort_topk(*args, K, _lib.TRANSFORM_NONE, 1.0, 0, _lib.SCORE_TC_3XTF32)
    sc_e, sc_t, ids_e, ids_t = sc_e.cpu().numpy(), sc_t.cpu().numpy(), ids_e.cpu().numpy(), ids_t.cpu().numpy()
    check(ids_e, sc_e, ids_t, sc_t, _lib.TRANSFORM_NONE, 1.0, 2e-5)


def test_tc_evaluator_metrics_close_to_exact():
    """whole fused evaluation with score_mode: tc -- metrics within 1e-5 of the exact mode's on a tie-free case"""
    U, I, users, ho, hi, po, pi, sst = random_eval_case(77, 1500, 2500, 64, 1400)
    res = {}
    for mode in ("exact", "tc"):
        data, ev, Ud, Id = build(U, I, users, ho, hi, po, pi, sst, [10], score_mode=mode)
        ev.collect(Ud, Id, data, 5.0)
        res[mode] = ev.finalize(ev.last, data, rounded=False)
    for k, v in res["exact"].items():
        tol = 3.0 / 1400 if "@" in k else 1e-5 * max(abs(v), 1e-3)   # a near-tie flip moves a top-K metric by O(1/n)
        assert abs(res["tc"][k] - v) <= tol, (k, res["tc"][k], v)


def test_tc_scorer_continues_into_masked_items_like_topk_of_the_masked_row():
    """users with fewer than K unmasked items: torch.topk over the masked row (trainer.py:435-438, collector.py:147)
    returns -inf entries; canonical order among them = ascending item id.  Tensor-core scorer == exact scorer == oracle,
    unsharded and item-sharded (3 shards, ragged)."""
    from recbole_fairrec_b200 import _lib, kernels
    rng = np.random.default_rng(5)
    n_users, n_items, d, K = 260, 300, 64, 10
    U = (rng.standard_normal((n_users, d)) * 0.3).astype(np.float32)
    I = (rng.standard_normal((n_items, d)) * 0.3).astype(np.float32)
    users = np.arange(1, n_users)
    hist = []
    for r, _ in enumerate(users):
        free = [0, 3, 9, 10, 25][r % 5] if r < 200 else 150            # unmasked items left to this user
        keep = rng.choice(np.arange(1, n_items), free, replace=False)
        hist.append(np.setdiff1d(np.arange(1, n_items), keep))
    ho = np.r_[0, np.cumsum([len(h) for h in hist])]
    hi = np.concatenate(hist)
    data, ev, Ud, Id = build(U, I, users, ho, hi, np.zeros(len(users) + 1, np.int64), np.zeros(0, np.int64),
                             rng.integers(1, 3, n_users), [K])
    args = (Ud, Id, data.users, data.hist_off, data.hist_items)
    ids_e, sc_e = kernels.fullsort_topk(*args, K, _lib.TRANSFORM_NONE, 1.0)
    ids_t, sc_t = kernels.fullsort_topk(*args, K, _lib.TRANSFORM_NONE, 1.0, 0, _lib.SCORE_TC_3XTF32)
    ids_e, sc_e, ids_t, sc_t = ids_e.cpu().numpy(), sc_e.cpu().numpy(), ids_t.cpu().numpy(), sc_t.cpu().numpy()
    oid, osc = fs.topk_canonical(fs.mask_history(fs.full_sort_scores(U, I, users, None), ho, hi), K)
    np.testing.assert_array_equal(ids_e, oid)                         # exact scorer == oracle, bit for bit
    np.testing.assert_array_equal(sc_e, osc)
    np.testing.assert_array_equal(ids_t, ids_e)                       # scores are well separated at this size
    np.testing.assert_array_equal(np.isinf(sc_t), np.isinf(sc_e))
    np.testing.assert_allclose(sc_t[np.isfinite(sc_t)], sc_e[np.isfinite(sc_e)], atol=2e-5, rtol=0)
    # row by row against first principles: the best unmasked ids by score, then masked ids ascending (0 first)
    raw = U[users].astype(np.float64) @ I.astype(np.float64).T
    for r in (0, 1, 2, 3, 4, 205):
        masked = np.zeros(n_items, bool)
        masked[0] = True
        masked[hist[r]] = True
        un = np.flatnonzero(~masked)
        want = list(un[np.argsort(-raw[r, un], kind="stable")][:K]) + list(np.flatnonzero(masked))
        np.testing.assert_array_equal(ids_t[r], want[:K])
    # item-sharded: per-shard lists merged under the same total order
    cuts = [0, 97, 211, n_items]
    parts_i, parts_s = [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        i_, s_ = kernels.fullsort_topk(Ud, Id[a:b].contiguous(), data.users, data.hist_off, data.hist_items, K,
                                       _lib.TRANSFORM_NONE, 1.0, a, _lib.SCORE_TC_3XTF32)
        parts_i.append(i_)
        parts_s.append(s_)
    mi, ms = kernels.topk_merge(torch.stack(parts_i), torch.stack(parts_s))
    np.testing.assert_array_equal(mi.cpu().numpy(), ids_e)
