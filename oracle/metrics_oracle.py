"""CPU restatement (numpy) of the 12 metrics of recbole/properties/model/FOCF.yaml:29-30, computed from
the collector outputs in full mode.  Values are returned UNROUNDED (the reference rounds to
`metric_decimal_place`, base_metric.py:81).

TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py cpu_baseline) -- never imported by the product.

Parity status: PINNED against tests/golden/focf_eval_*.npz (the reference's own Evaluator output at 12
decimals) and KAT-2 of SURVEY.md section 4, in tests/test_oracle_golden.py.
Line numbers below refer to /root/reference/recbole/evaluator/metrics.py unless stated otherwise.
"""
from collections import OrderedDict

import numpy as np


def _topk_mean(per_user, topk):
    """base_metric.py:59-82 TopkMetric.topk_result -- mean over users of column k-1."""
    avg = per_user.mean(axis=0)
    return {k: float(avg[k - 1]) for k in topk}


def hit(pos_index):                                     # :63-65
    return (np.cumsum(pos_index, axis=1) > 0).astype(int)


def mrr(pos_index):                                     # :89-97
    idxs = pos_index.argmax(axis=1)
    out = np.zeros(pos_index.shape, dtype=np.float64)
    for row, idx in enumerate(idxs):
        out[row, idx:] = 1.0 / (idx + 1) if pos_index[row, idx] > 0 else 0.0
    return out


def recall(pos_index, pos_len):                         # :160-161
    return np.cumsum(pos_index, axis=1) / pos_len.reshape(-1, 1)


def ndcg(pos_index, pos_len):                           # :187-203
    K = pos_index.shape[1]
    idcg_len = np.minimum(pos_len, K)
    ranks = np.arange(1, K + 1, dtype=np.float64)
    idcg = np.tile(np.cumsum(1.0 / np.log2(ranks + 1)), (pos_index.shape[0], 1))
    for row, idx in enumerate(idcg_len):
        idcg[row, idx:] = idcg[row, idx - 1]
    dcg = np.cumsum(np.where(pos_index, 1.0 / np.log2(ranks + 1), 0), axis=1)
    return dcg / idcg


def gini(item_matrix, num_items):                       # :644-661
    _, counts = np.unique(item_matrix.flatten(), return_counts=True)
    sorted_count = np.sort(counts)
    m = sorted_count.shape[0]
    total = item_matrix.shape[0] * item_matrix.shape[1]
    idx = np.arange(num_items - m + 1, num_items + 1)
    return float(np.sum((2 * idx - num_items - 1) * sorted_count) / total / num_items)


def popular_items(count_items, popularity_ratio):       # :786-794
    """count_items: dict item -> train count.  Returns the set of 'popular' items."""
    if popularity_ratio is None or popularity_ratio <= 0:
        popularity_ratio = 0.1                          # :763-765
    if popularity_ratio > 1:
        return {i for i, c in count_items.items() if c >= popularity_ratio}
    srt = sorted(count_items.items(), key=lambda kv: (kv[1], kv[0]), reverse=True)
    cut = max(int(len(srt) * popularity_ratio), 1)
    return {i for i, _ in srt[:cut]}


def popularity_percentage(item_matrix, count_items, popularity_ratio):     # :772-804
    pop = popular_items(count_items, popularity_ratio)
    value = np.isin(item_matrix, np.fromiter(pop, dtype=np.int64, count=len(pop))).astype(np.int64)
    return value.cumsum(axis=1) / np.arange(1, value.shape[1] + 1)


def nonparity(score, sst_value):                        # :860-881 (float32 like the reference)
    uniq = np.unique(sst_value)
    if len(uniq) < 2:
        raise ValueError("there is only one value for sensitive attribute")
    avgs = [np.mean(score[sst_value == s]) for s in uniq]
    if len(uniq) == 2:
        return float(np.abs(avgs[0] - avgs[1]))
    return float(np.std(avgs))


def _item_group_means(pos_score, pos_iids, sst_value):
    """:935-973 (full mode): P = sum score/(cnt+1e-5), T = cnt/(cnt+1e-5), float64."""
    suniq, sidx = np.unique(sst_value, return_inverse=True)
    iuniq, iidx = np.unique(pos_iids, return_inverse=True)
    if len(suniq) != 2:
        raise ValueError("sensitive attribute must be binary")          # :951-952
    P = np.zeros((len(iuniq), 2))
    n = np.zeros((len(iuniq), 2))
    np.add.at(P, (iidx, sidx), pos_score.astype(np.float64))
    np.add.at(n, (iidx, sidx), 1.0)
    T = n.copy()
    n += 1e-5
    return P / n, T / n


def value_unfairness(pos_score, pos_iids, sst_value):   # :975-978
    P, T = _item_group_means(pos_score, pos_iids, sst_value)
    D = P - T
    return float(np.mean(np.abs(D[:, 0] - D[:, 1])))


def absolute_unfairness(pos_score, pos_iids, sst_value):    # :1071-1074
    P, T = _item_group_means(pos_score, pos_iids, sst_value)
    D = np.abs(P - T)
    return float(np.mean(np.abs(D[:, 0] - D[:, 1])))


def under_unfairness(pos_score, pos_iids, sst_value):   # :1167-1170
    P, T = _item_group_means(pos_score, pos_iids, sst_value)
    D = np.where((T - P) > 0, T - P, 0)
    return float(np.mean(np.abs(D[:, 0] - D[:, 1])))


def over_unfairness(pos_score, pos_iids, sst_value):    # :1263-1266
    P, T = _item_group_means(pos_score, pos_iids, sst_value)
    D = np.where((P - T) > 0, P - T, 0)
    return float(np.mean(np.abs(D[:, 0] - D[:, 1])))


def differential_fairness(score, iids, sst_value):      # :1313-1341
    suniq, sidx = np.unique(sst_value, return_inverse=True)
    iuniq, iidx = np.unique(iids, return_inverse=True)
    J, G = len(iuniq), len(suniq)
    S = np.zeros((J, G), np.float64)
    n = np.zeros((J, G), np.float64)
    np.add.at(S, (iidx, sidx), score.astype(np.float64))
    np.add.at(n, (iidx, sidx), 1.0)
    alpha = 1.0 / J
    M = ((S + alpha) / (n + 1.0)).astype(np.float32)
    eps = np.zeros(J, np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(G):
            for j in range(i + 1, G):
                e = np.abs(np.log(M[:, i]) - np.log(M[:, j]))
                eps = np.where(e > eps, e, eps)
    return float(eps.mean())


def evaluate(struct, topk, num_items, count_items, popularity_ratio=0.1, sst_name="gender"):
    """recbole/evaluator/evaluator.py:28-42 for the 12 FOCF metrics; keys as the reference emits them."""
    rec_topk = struct["rec.topk"]
    K = rec_topk.shape[1] - 1
    pos_index = rec_topk[:, :K].astype(bool)
    pos_len = rec_topk[:, K]
    items = struct["rec.items"]
    score, iids, sst = struct["rec.positive_score"], struct["data.positive_i"], struct["data.sst"]
    out = OrderedDict()
    for name, per_user in (("ndcg", ndcg(pos_index, pos_len)), ("recall", recall(pos_index, pos_len)),
                           ("hit", hit(pos_index)), ("mrr", mrr(pos_index))):
        for k, v in _topk_mean(per_user, topk).items():
            out[f"{name}@{k}"] = v
    out[f"Differential Fairness of sensitive attribute {sst_name}"] = differential_fairness(score, iids, sst)
    for k in topk:
        out[f"giniindex@{k}"] = gini(items[:, :k], num_items)
    for k, v in _topk_mean(popularity_percentage(items, count_items, popularity_ratio), topk).items():
        out[f"popularitypercentage@{k}"] = v
    out[f"Value Unfairness of sensitive attribute {sst_name}"] = value_unfairness(score, iids, sst)
    out[f"Absolute Unfairness of sensitive attribute {sst_name}"] = absolute_unfairness(score, iids, sst)
    out[f"Underestimation Unfairness of sensitive attribute {sst_name}"] = under_unfairness(score, iids, sst)
    out[f"Overestimation Unfairness of sensitive attribute {sst_name}"] = over_unfairness(score, iids, sst)
    out[f"NonParity Unfairness of sensitive attribute {sst_name}"] = nonparity(score, sst)
    return out
