"""CPU restatement (numpy) of the reference's sampled-negative ("uni100") ranking evaluation.
TEST INFRASTRUCTURE ONLY (tests/, smoke, bench cpu legs) -- never imported by the product package.

Parity status: PINNED against tests/golden/uni_eval_*.npz (generated from the unmodified reference by
oracle/gen_golden.py `uni`) in tests/test_oracle_golden.py.

Restates (paths relative to /root/reference):
  recbole/data/dataloader/general_dataloader.py:128-152  per user: [positives ; neg_num negatives per positive]
  recbole/trainer/trainer.py:441-456  _neg_sample_batch_eval: predict the candidates, scatter into a [users, n_items]
                                      matrix of -inf (duplicate candidates collapse: same pair, same score)
  recbole/evaluator/collector.py:141-153, 179  topk -> rec.items, hit bits | number of (distinct) positives -> rec.topk,
                                      rec.positive_score
Ties (and the -inf filler when a user has fewer than K distinct candidates) follow the canonical order of the north star:
score descending, item id ascending -- torch.topk leaves that order unspecified.
"""
import numpy as np

from . import fullsort_oracle as fs


def candidate_lists(pos_off, pos_items, neg_items, neg_num):
    """per eval user: (positives, negatives) in the dataloader's order"""
    out = []
    for k in range(len(pos_off) - 1):
        p0, p1 = int(pos_off[k]), int(pos_off[k + 1])
        out.append((pos_items[p0:p1], neg_items[p0 * neg_num:p1 * neg_num]))
    return out


def dense_rows(U, I, users, cands, n_items, max_rating):
    """trainer.py:452-455 with FOCF.predict (focf.py:145-150) as the scorer"""
    rows = np.full((len(users), n_items), -np.inf, np.float32)
    for k, (u, (pos, neg)) in enumerate(zip(users, cands)):
        items = np.concatenate([pos, neg]).astype(np.int64)
        uid = np.full(len(items), u, np.int64)
        rows[k, items] = fs.pair_scores(U, I, uid, items, max_rating)
    return rows


def collect(rows, cands, K):
    """collector.py:141-153, 179 -> (rec.items [n,K], rec.topk [n,K+1], rec.positive_score [n_pos])"""
    ids, _ = fs.topk_canonical(rows, K)
    n = rows.shape[0]
    rec_topk = np.zeros((n, K + 1), np.int32)
    pos_score = []
    for k, (pos, _) in enumerate(cands):
        pset = set(int(x) for x in pos)
        rec_topk[k, :K] = [1 if int(i) in pset else 0 for i in ids[k]]
        rec_topk[k, K] = len(pset)
        pos_score.append(rows[k, np.asarray(pos, np.int64)])
    return ids, rec_topk, np.concatenate(pos_score).astype(np.float32)


def reference_sst_of_pos(users, cands, sst_of_user, users_per_batch):
    """collector.py:203-205 in sampled mode takes `interaction[sst][arange(len(positive_u))]`: the attribute of the FIRST
    P rows of the batch interaction ([u1 positives ; u1 negatives ; u2 positives ; ...]), where P = positives of the
    whole batch -- correct only for single-user batches; with several users per batch the rows mostly belong to the
    first user.  Restated as is, so that the reference's DifferentialFairness / NonParity values can be pinned."""
    out = []
    for b0 in range(0, len(users), users_per_batch):
        bu, bc = users[b0:b0 + users_per_batch], cands[b0:b0 + users_per_batch]
        row_user = np.concatenate([np.full(len(p) + len(n), u) for u, (p, n) in zip(bu, bc)])
        P = sum(len(p) for p, _ in bc)
        out.append(np.asarray(sst_of_user)[row_user[:P]])
    return np.concatenate(out)


def metrics(ids, rec_topk, pos_score, pos_items, sst_of_pos, topk, num_items, count_items, popularity_ratio=0.1,
            sst_name="gender"):
    """evaluator.py:28-42 for the metrics that do not depend on the evaluation mode (the four *Unfairness metrics read
    rec.negative_score / data.negative_i in sampled mode, collector.py:190-199, and are not restated)"""
    from collections import OrderedDict

    from . import metrics_oracle as mo
    K = rec_topk.shape[1] - 1
    pos_index, pos_len = rec_topk[:, :K].astype(bool), rec_topk[:, K]
    out = OrderedDict()
    for name, per_user in (("ndcg", mo.ndcg(pos_index, pos_len)), ("recall", mo.recall(pos_index, pos_len)),
                           ("hit", mo.hit(pos_index)), ("mrr", mo.mrr(pos_index))):
        for k, v in mo._topk_mean(per_user, topk).items():
            out[f"{name}@{k}"] = v
    out[f"Differential Fairness of sensitive attribute {sst_name}"] = mo.differential_fairness(pos_score, pos_items, sst_of_pos)
    for k in topk:
        out[f"giniindex@{k}"] = mo.gini(ids[:, :k], num_items)
    for k, v in mo._topk_mean(mo.popularity_percentage(ids, count_items, popularity_ratio), topk).items():
        out[f"popularitypercentage@{k}"] = v
    out[f"NonParity Unfairness of sensitive attribute {sst_name}"] = mo.nonparity(pos_score, sst_of_pos)
    return out


def unfairness_sampled(pos_score, pos_items, neg_score, neg_items, sst_of_pos):
    """metrics.py:935-978, 1031-1074, 1127-1170, 1224-1266 with mode != 'full' (float64, like the reference's numpy loops):
    item set = union of positive and paired-negative items; negatives add score + count under the positive's group, no
    "true" mass.  Returns (value, absolute, under, over)."""
    sst_unique, sst_idx = np.unique(sst_of_pos, return_inverse=True)
    if len(sst_unique) != 2:
        raise ValueError("sensitive attribute must be binary")
    items, idx = np.unique(np.concatenate((pos_items, neg_items)), return_inverse=True)
    P = len(pos_items)
    pred = np.zeros((len(items), 2))
    num = np.zeros((len(items), 2))
    true = np.zeros((len(items), 2))
    np.add.at(pred, (idx[:P], sst_idx), np.asarray(pos_score, np.float64))
    np.add.at(num, (idx[:P], sst_idx), 1.0)
    np.add.at(true, (idx[:P], sst_idx), 1.0)
    np.add.at(pred, (idx[P:], sst_idx), np.asarray(neg_score, np.float64))
    np.add.at(num, (idx[P:], sst_idx), 1.0)
    num += 1e-5
    pred /= num
    true /= num
    out = []
    for D in (pred - true, np.abs(pred - true), np.where(true - pred > 0, true - pred, 0.0),
              np.where(pred - true > 0, pred - true, 0.0)):
        out.append(float(np.mean(np.abs(D[:, 0] - D[:, 1]))))
    return tuple(out)
