"""CPU restatement of the reference's full-sort evaluation hand-off (scores -> mask -> top-K ->
collector outputs).

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline).  Never imported by
the product package.

Parity status: PINNED against tests/golden/focf_eval_*.npz (generated from the unmodified reference by
oracle/gen_golden.py) in tests/test_oracle_golden.py: bit-exact `rec.items` / `rec.topk` on tie-free
rows, canonical (score desc, item-id asc) on tied rows where torch.topk itself is unspecified.

The dot product is evaluated as the k-ascending fused-multiply-add chain (C: oracle/c/fairrec_oracle.c,
fmaf) so that the CUDA "exact" scoring mode is bit-identical to it.  If the C library has not been
built, a float64-emulated chain is used (identical except for ~2^-29-probability double roundings).
"""
import ctypes
import os

import numpy as np

F32 = np.float32
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def clib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libfairrec_oracle.so")
        if not os.path.exists(path):
            return None
        _LIB = ctypes.CDLL(path)
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def full_sort_scores(U, I, users, max_rating=None):
    """recbole/model/fair_recommender/focf.py:171-178 -- clamp(U[users] @ I.T, 0, max_rating) / max_rating.
    max_rating=None returns the raw dot products."""
    U = np.ascontiguousarray(U, F32)
    I = np.ascontiguousarray(I, F32)
    users = np.ascontiguousarray(users, np.int64)
    n, Ni, d = len(users), I.shape[0], I.shape[1]
    lib = clib()
    out = np.empty((n, Ni), F32)
    if lib is not None:
        lib.oracle_full_sort_scores(_p(U), _p(I), _p(users), ctypes.c_int64(n), ctypes.c_int64(Ni),
                                    ctypes.c_int64(d), ctypes.c_int(0 if max_rating is None else 1),
                                    ctypes.c_float(0.0 if max_rating is None else max_rating), _p(out))
        return out
    acc = np.zeros((n, Ni), F32)
    Uu = U[users].astype(np.float64)
    I64 = I.astype(np.float64)
    for k in range(d):
        acc = (acc.astype(np.float64) + Uu[:, k, None] * I64[None, :, k]).astype(F32)
    if max_rating is not None:
        acc = (np.clip(acc, F32(0), F32(max_rating)) / F32(max_rating)).astype(F32)
    return acc


def pair_scores(U, I, uid, iid, max_rating=None):
    """Scores of explicit (user, item) pairs with the same arithmetic as full_sort_scores."""
    U = np.ascontiguousarray(U, F32)
    I = np.ascontiguousarray(I, F32)
    uid = np.ascontiguousarray(uid, np.int64)
    iid = np.ascontiguousarray(iid, np.int64)
    lib = clib()
    out = np.empty(len(uid), F32)
    if lib is not None:
        lib.oracle_pair_scores(_p(U), _p(I), _p(uid), _p(iid), ctypes.c_int64(len(uid)),
                               ctypes.c_int64(U.shape[1]), ctypes.c_int(0 if max_rating is None else 1),
                               ctypes.c_float(0.0 if max_rating is None else max_rating), _p(out))
        return out
    acc = np.zeros(len(uid), F32)
    for k in range(U.shape[1]):
        acc = (acc.astype(np.float64) + U[uid, k].astype(np.float64) * I[iid, k].astype(np.float64)).astype(F32)
    if max_rating is not None:
        acc = (np.clip(acc, F32(0), F32(max_rating)) / F32(max_rating)).astype(F32)
    return acc


def mask_history(scores, hist_off, hist_items):
    """recbole/trainer/trainer.py:435-438 -- scores[:,0] = -inf ; scores[history_u, history_i] = -inf."""
    scores = scores.copy()
    scores[:, 0] = -np.inf
    for r in range(scores.shape[0]):
        scores[r, hist_items[hist_off[r]:hist_off[r + 1]]] = -np.inf
    return scores


def topk_canonical(scores, K):
    """recbole/evaluator/collector.py:143 torch.topk restated with the canonical total order
    (score desc, item id asc); torch.topk leaves the order of tied entries unspecified."""
    order = np.argsort(-scores, axis=1, kind="stable")[:, :K]
    return order.astype(np.int64), np.take_along_axis(scores, order, axis=1)


def collect(scores, K, pos_off, pos_items, sst_of_row):
    """recbole/evaluator/collector.py:141-189 (full mode) for ONE concatenated block of eval users.
    Returns the DataStruct entries: rec.items i64[n,K], rec.topk i32[n,K+1], rec.positive_score f32[p],
    data.positive_i i64[p], data.<sst>[p]."""
    n = scores.shape[0]
    items, _ = topk_canonical(scores, K)
    pos_u = np.repeat(np.arange(n), np.diff(pos_off))
    pos_matrix = np.zeros(scores.shape, np.int32)
    pos_matrix[pos_u, pos_items] = 1
    hit = np.take_along_axis(pos_matrix, items, axis=1)
    rec_topk = np.concatenate([hit, pos_matrix.sum(axis=1, keepdims=True)], axis=1).astype(np.int32)
    return {
        "rec.items": items,
        "rec.topk": rec_topk,
        "rec.positive_score": scores[pos_u, pos_items].astype(F32),
        "data.positive_i": np.asarray(pos_items, np.int64),
        "data.sst": np.asarray(sst_of_row)[pos_u],
    }


def full_sort_eval(U, I, users, max_rating, hist_off, hist_items, pos_off, pos_items, sst_of_user, K):
    """trainer.py:505-510 over all eval users at once (the reference walks them in batches of
    max(eval_batch_size // n_items, 1) users; the concatenation is identical)."""
    s = mask_history(full_sort_scores(U, I, users, max_rating), hist_off, hist_items)
    return collect(s, K, pos_off, pos_items, np.asarray(sst_of_user)[users])


def threaded_topk(U, I, users, max_rating, hist_off, hist_items, K):
    """Multi-threaded C path (scoring + mask + top-K per user), the CPU baseline of bench.py."""
    lib = clib()
    if lib is None:
        raise RuntimeError("oracle C library not built: run `make -C oracle`")
    U = np.ascontiguousarray(U, F32)
    I = np.ascontiguousarray(I, F32)
    users = np.ascontiguousarray(users, np.int64)
    hist_off = np.ascontiguousarray(hist_off, np.int64)
    hist_items = np.ascontiguousarray(hist_items, np.int64)
    n = len(users)
    ids = np.empty((n, K), np.int64)
    vals = np.empty((n, K), F32)
    lib.oracle_full_sort_topk(_p(U), _p(I), _p(users), ctypes.c_int64(n), ctypes.c_int64(I.shape[0]),
                              ctypes.c_int64(I.shape[1]), ctypes.c_int(1), ctypes.c_float(max_rating),
                              _p(hist_off), _p(hist_items), ctypes.c_int64(K), _p(ids), _p(vals))
    return ids, vals
