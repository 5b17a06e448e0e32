"""CPU: size-independent properties of the oracle (the checker the GPU parity tests trust) -- invariances and bounds that
must hold for ANY input, on seeded random cases: the canonical top-K is a total order and survives item sharding, user
permutation does not move a metric, swapping the two groups does not move a symmetric fairness metric, the bounded
metrics stay in their ranges, and the degenerate cases the reference defines (identical groups, single positive) give the
values its formulas imply.  These are the properties the full-size runs (bench_scaleout.py) are checked with as well."""
import numpy as np
import pytest

from oracle import fullsort_oracle as fs
from oracle import metrics_oracle as mo


def case(seed, n=60, n_items=120, d=8, K=7):
    rng = np.random.default_rng(seed)
    U = (rng.standard_normal((n + 1, d)) * 0.5).astype(np.float32)
    I = (rng.standard_normal((n_items, d)) * 0.5).astype(np.float32)
    users = np.arange(1, n + 1)
    hist, pos = [], []
    for _ in users:
        used = rng.choice(np.arange(1, n_items), int(rng.integers(3, 40)), replace=False)
        k = int(rng.integers(1, min(8, len(used))))
        pos.append(np.sort(used[:k]))
        hist.append(np.sort(used[k:]))
    ho = np.r_[0, np.cumsum([len(h) for h in hist])]
    po = np.r_[0, np.cumsum([len(p) for p in pos])]
    sst = rng.integers(1, 3, n + 1)
    return U, I, users, ho, np.concatenate(hist), po, np.concatenate(pos), sst, K


@pytest.mark.parametrize("seed", range(4))
def test_canonical_topk_is_a_total_order_and_survives_item_sharding(seed):
    U, I, users, ho, hi, po, pi, sst, K = case(seed)
    s = fs.mask_history(fs.full_sort_scores(U, I, users, 5.0), ho, hi)       # clamp -> exact ties at 0
    ids, vals = fs.topk_canonical(s, K)
    assert (np.diff(vals, axis=1) <= 0).all()
    tie = np.diff(vals, axis=1) == 0
    assert (np.diff(ids, axis=1)[tie] > 0).all()                             # ties broken by ascending item id
    # every item outside the list is not better than the last one inside, under (score desc, id asc)
    for r in range(0, len(users), 7):
        rest = np.setdiff1d(np.arange(s.shape[1]), ids[r])
        worst_s, worst_i = vals[r, -1], ids[r, -1]
        assert not ((s[r, rest] > worst_s) | ((s[r, rest] == worst_s) & (rest < worst_i))).any()
    # contiguous item shards, per-shard canonical lists (global ids), merged under the same order
    cuts = [0, 31, 77, s.shape[1]]
    cand_i = np.concatenate([fs.topk_canonical(s[:, a:b], min(K, b - a))[0] + a for a, b in zip(cuts[:-1], cuts[1:])], axis=1)
    cand_s = np.take_along_axis(s, cand_i, axis=1)
    order = np.lexsort((cand_i, -cand_s), axis=1)[:, :K]
    np.testing.assert_array_equal(np.take_along_axis(cand_i, order, axis=1), ids)


@pytest.mark.parametrize("seed", range(4))
def test_metrics_are_invariant_under_user_permutation_and_bounded(seed):
    U, I, users, ho, hi, po, pi, sst, K = case(10 + seed)
    st = fs.full_sort_eval(U, I, users, 5.0, ho, hi, po, pi, sst, K)
    counts = {int(i): int(c) for i, c in enumerate(np.bincount(hi, minlength=I.shape[0])) if c}
    a = mo.evaluate(st, [3, K], I.shape[0], counts, 0.1)
    perm = np.random.default_rng(seed).permutation(len(users))
    hist = [hi[ho[r]:ho[r + 1]] for r in perm]
    pos = [pi[po[r]:po[r + 1]] for r in perm]
    st2 = fs.full_sort_eval(U, I, users[perm], 5.0, np.r_[0, np.cumsum([len(h) for h in hist])], np.concatenate(hist),
                            np.r_[0, np.cumsum([len(p) for p in pos])], np.concatenate(pos), sst, K)
    b = mo.evaluate(st2, [3, K], I.shape[0], counts, 0.1)
    for k in a:
        assert abs(a[k] - b[k]) <= 1e-6 * max(abs(a[k]), 1.0), k             # float32 sums in another order
    for k, v in a.items():
        if "@" in k:
            assert 0.0 <= v <= 1.0, (k, v)
        else:
            assert v >= 0.0, (k, v)
    assert a[f"hit@{K}"] >= a["hit@3"] and a[f"recall@{K}"] >= a["recall@3"] and a["mrr@3"] <= a["hit@3"]


@pytest.mark.parametrize("seed", range(4))
def test_fairness_metrics_symmetries_and_degenerate_cases(seed):
    rng = np.random.default_rng(20 + seed)
    p = 400
    score = rng.random(p).astype(np.float32)
    iids = rng.integers(1, 30, p)
    sst = rng.integers(1, 3, p)
    swapped = 3 - sst
    for f in (mo.value_unfairness, mo.absolute_unfairness, mo.under_unfairness, mo.over_unfairness):
        assert abs(f(score, iids, sst) - f(score, iids, swapped)) < 1e-12    # |D0 - D1| is symmetric in the groups
    assert abs(mo.nonparity(score, sst) - mo.nonparity(score, swapped)) < 1e-7
    assert abs(mo.differential_fairness(score, iids, sst) - mo.differential_fairness(score, iids, swapped)) < 1e-6
    # two groups with identical (item, score) multisets: every between-group difference vanishes
    score2, iids2 = np.r_[score, score], np.r_[iids, iids]
    sst2 = np.r_[np.ones(p, np.int64), 2 * np.ones(p, np.int64)]
    assert mo.value_unfairness(score2, iids2, sst2) < 1e-12 and mo.absolute_unfairness(score2, iids2, sst2) < 1e-12
    assert mo.nonparity(score2, sst2) < 1e-7 and mo.differential_fairness(score2, iids2, sst2) < 1e-6
    # over / under split the signed error: predictions above the 0/1 "true" mean T = cnt / (cnt + 1e-5) (~1) never happen
    # for scores in [0, 1) -> over-estimation unfairness is exactly 0 (KAT-2 shows the same)
    assert mo.over_unfairness(score * 0.99, iids, sst) == 0.0


def test_gini_and_popularity_limits():
    n_items = 50
    same = np.tile(np.arange(1, 6), (40, 1))                 # everyone gets the same five items: maximal concentration
    spread = (np.arange(200).reshape(40, 5) % (n_items - 1)) + 1     # 200 slots round-robin over the 49 real items
    g_same, g_spread = mo.gini(same, n_items), mo.gini(spread, n_items)
    assert 0.0 <= g_spread < g_same <= 1.0
    counts = {i: 100 - i for i in range(1, n_items)}         # item 1 most popular
    top = mo.popular_items(counts, 0.1)
    assert sorted(top) == [1, 2, 3, 4] and len(top) == max(int(len(counts) * 0.1), 1)
    pp = mo.popularity_percentage(same, counts, 0.1)         # 4 of the 5 recommended items are popular
    assert pp.shape == same.shape and abs(pp.mean(axis=0)[-1] - 0.8) < 1e-12


def test_gini_from_a_histogram_of_count_values_equals_the_sorted_definition():
    """The identity `k_gini_hist` (csrc/metrics.cu) rests on: the h items whose count is v occupy the ascending positions
    P + 1 .. P + h (P = items with a smaller count), so sum_p (2p - N - 1) c_(p) = sum_v v (2 (h P + h (h + 1) / 2) - h (N + 1))
    -- integer arithmetic, no sort -- and the result is the oracle's Gini (metrics.py:644-661)."""
    rng = np.random.default_rng(4)
    for n_users, n_items, K in ((50, 40, 3), (300, 1000, 5), (2000, 257, 10)):
        rec = np.stack([rng.choice(np.arange(1, n_items), size=K, replace=False, p=None) for _ in range(n_users)])
        counts = np.bincount(rec.reshape(-1), minlength=n_items).astype(np.int64)      # <= n_users each
        assert counts.max() <= n_users
        c_sorted = np.sort(counts)
        want_int = int(((2 * np.arange(1, n_items + 1) - n_items - 1) * c_sorted).sum())
        hist = np.bincount(counts, minlength=n_users + 1).astype(np.int64)
        P = np.concatenate([[0], np.cumsum(hist)[:-1]])
        v = np.arange(hist.size, dtype=np.int64)
        got_int = int((v * (2 * (hist * P + hist * (hist + 1) // 2) - hist * (n_items + 1))).sum())
        assert got_int == want_int
        got = got_int / float(n_users * K) / float(n_items)
        np.testing.assert_allclose(got, mo.gini(rec, n_items), rtol=1e-12)
