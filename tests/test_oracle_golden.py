"""Pins the oracle (oracle/*.py, oracle/c) against fixtures produced by the UNMODIFIED reference
(tests/golden/*.npz, made by oracle/gen_golden.py) and against KAT-1/KAT-2 of SURVEY.md section 4."""
import glob
import os

import numpy as np
import pytest

from oracle import focf_oracle as fo
from oracle import fullsort_oracle as fs
from oracle import metrics_oracle as mo

RTOL = 1e-5  # north star: losses, metrics and updated embeddings within 1e-5 relative


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


TRAIN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "focf_train_*.npz")))


@pytest.mark.parametrize("path", TRAIN, ids=[os.path.basename(p)[11:-4] for p in TRAIN])
def test_focf_train_matches_reference(path):
    g = np.load(path)
    obj, fw = str(g["objective"]), float(g["fair_weight"])
    batches = [(g[f"uid{s}"], g[f"iid{s}"], g[f"rating{s}"], g[f"sst{s}"]) for s in range(int(g["n_steps"]))]
    uid, iid, r, sst = batches[0]
    pred, coef, dU, dI = fo.grads(g["U0"], g["I0"], uid, iid, r, sst, obj, fw)
    assert rel_err(pred, g["pred0"]) < RTOL
    assert rel_err(fo.predict(g["U0"], g["I0"], uid, iid, float(g["max_rating"])), g["predict0"]) < RTOL
    assert rel_err(dU, g["dU0"]) < RTOL
    assert rel_err(dI, g["dI0"]) < RTOL
    losses, U, I, mU, vU, mI, vI = fo.train_steps(g["U0"], g["I0"], batches, obj, fw, float(g["lr"]), float(g["wd"]))
    np.testing.assert_allclose(losses, g["losses"], rtol=RTOL)
    for mine, ref in ((U, "U_final"), (I, "I_final"), (mU, "mU_final"), (mI, "mI_final"),
                      (vU, "vU_final"), (vI, "vI_final")):
        assert rel_err(mine, g[ref]) < RTOL, ref


def test_kat1_focf():
    """SURVEY.md section 4, KAT-1."""
    U = (np.arange(28, dtype=np.float32).reshape(7, 4) / 10 - 1.0).astype(np.float32)
    I = (np.flip(np.arange(20, dtype=np.float32).reshape(5, 4), 0) / 8 - 0.9).astype(np.float32)
    uid = np.array([1, 2, 3, 4, 5, 6, 1, 3])
    iid = np.array([1, 1, 1, 2, 2, 2, 3, 3])
    r = np.array([5, 3, 1, 4, 2, 5, 3, 4], np.float32)
    g = np.array([1, 2, 1, 2, 2, 1, 1, 2])
    np.testing.assert_allclose(fo.forward(U, I, uid, iid),
                               [-1.355, -0.095, 1.165, 0.925, 1.385, 1.845, 0.445, -0.235], atol=2e-6)
    want = dict(none=11.780424118, value=12.443744659, absolute=12.443744659, under=12.443744659,
                over=11.780424118, nonparity=11.780874252)
    for obj, v in want.items():
        np.testing.assert_allclose(fo.calculate_loss(U, I, uid, iid, r, g, obj, 1.0), v, rtol=1e-6)
    _, _, dU, dI = fo.grads(U, I, uid, iid, r, g, "value", 1.0)
    np.testing.assert_allclose(dU[1], [-0.83108360, -1.06785524, -1.30462682, -1.54139841], rtol=1e-5)
    np.testing.assert_allclose(dI[1], [1.11625004, 0.88412505, 0.65200001, 0.41987500], rtol=1e-5)
    _, _, dU, _ = fo.grads(U, I, uid, iid, r, g, "none", 1.0)
    np.testing.assert_allclose(dU[1], [-0.69775009, -0.97618753, -1.25462508, -1.53306258], rtol=1e-5)
    s = fs.full_sort_scores(U, I, np.array([1, 2]), 5.0)
    np.testing.assert_allclose(s, [[0, 0, 0, 0.089, 0.269], [0, 0, 0.001, 0.021, 0.041]], atol=2e-7)


def test_kat2_metrics():
    """SURVEY.md section 4, KAT-2."""
    hitm = np.array([[1, 0, 1], [0, 0, 0], [0, 1, 0]], bool)
    pos_len = np.array([2, 1, 3])
    np.testing.assert_allclose(mo.ndcg(hitm, pos_len)[:, 2], [0.91972079, 0, 0.29608191], atol=1e-8)
    np.testing.assert_allclose(mo.recall(hitm, pos_len)[:, 2], [1, 0, 1 / 3])
    np.testing.assert_allclose(mo.hit(hitm)[:, 2], [1, 0, 1])
    np.testing.assert_allclose(mo.mrr(hitm)[:, 2], [1, 0, 0.5])
    score = np.array([0.9, 0.2, 0.4, 1.0, 0.0, 0.6, 0.7], np.float32)
    item = np.array([1, 1, 2, 2, 2, 3, 3])
    sst = np.array([1, 2, 1, 2, 2, 1, 1])
    np.testing.assert_allclose(mo.differential_fairness(score, item, sst), 0.507108, rtol=1e-5)
    np.testing.assert_allclose(mo.nonparity(score, sst), 0.24999997, rtol=1e-6)
    for f in (mo.value_unfairness, mo.absolute_unfairness, mo.under_unfairness):
        np.testing.assert_allclose(f(score, item, sst), 0.3833292371278623, rtol=1e-12)
    assert mo.over_unfairness(score, item, sst) == 0.0
    items = np.array([[1, 2, 3], [1, 2, 4], [1, 3, 2]])
    np.testing.assert_allclose(mo.gini(items, 5), 0.35555555555, rtol=1e-9)
    pp = mo.popularity_percentage(items, {1: 10, 2: 5, 3: 5, 4: 1}, 0.5).mean(axis=0)
    np.testing.assert_allclose(pp, [1, 2 / 3, 5 / 9], rtol=1e-12)


EVAL = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "focf_eval_*.npz")))


@pytest.mark.parametrize("path", EVAL, ids=[os.path.basename(p)[10:-4] for p in EVAL])
def test_fullsort_eval_matches_reference(path):
    g = np.load(path)
    K = int(g["K"])
    users = g["eval_users"]
    scores = fs.mask_history(fs.full_sort_scores(g["U"], g["I"], users, float(g["max_rating"])),
                             g["hist_off"], g["hist_items"])
    st = fs.collect(scores, K, g["pos_off"], g["pos_items"], g["sst_of_user"][users])
    # positive scores: the reference used MKL sgemm, we use an fma chain -> a few ulp
    np.testing.assert_allclose(st["rec.positive_score"], g["rec_positive_score"], rtol=1e-5, atol=1e-7)
    np.testing.assert_array_equal(st["data.positive_i"], g["data_positive_i"])
    np.testing.assert_array_equal(st["data.sst"], g["data_sst"])
    np.testing.assert_array_equal(st["rec.topk"][:, K], g["rec_topk"][:, K])          # pos_len
    # rows whose top-(K+1) scores are pairwise separated by > 1e-6 must match torch.topk bit-exactly;
    # torch.topk's order among exactly-tied scores is unspecified, there the canonical rule decides.
    _, vals = fs.topk_canonical(scores, K + 1)
    gap = np.abs(np.diff(vals, axis=1))
    clean = np.all(gap > 1e-6, axis=1)
    assert clean.sum() >= (len(users) // 2 if "tiefree" in path or "float" in path else 1)
    np.testing.assert_array_equal(st["rec.items"][clean], g["rec_items"][clean])
    np.testing.assert_array_equal(st["rec.topk"][clean], g["rec_topk"][clean])
    for r in np.where(~clean)[0]:   # tied rows: same score multiset, ids ascending inside ties
        ref_scores = scores[r, g["rec_items"][r]]
        np.testing.assert_array_equal(np.sort(ref_scores), np.sort(vals[r, :K]))
        mine = st["rec.items"][r]
        for a, b in zip(range(K - 1), range(1, K)):
            if vals[r, a] == vals[r, b]:
                assert mine[a] < mine[b]
    if clean.all():
        count_items = {int(i): int(c) for i, c in g["train_count_items"]}
        res = mo.evaluate(st, [int(k) for k in g["topk"]], g["I"].shape[0], count_items, 0.1)
        assert list(res.keys()) == [str(k) for k in g["metric_names"]]
        for (k, v), ref in zip(res.items(), g["metric_values"]):
            assert abs(v - ref) <= RTOL * max(abs(ref), 1e-12) + 1e-12, (k, v, ref)


@pytest.mark.parametrize("path", EVAL, ids=[os.path.basename(p)[10:-4] for p in EVAL])
def test_metrics_match_reference_on_reference_struct(path):
    """Feed the reference's own collector outputs to the restated metrics: isolates metrics.py parity
    from top-K tie handling."""
    g = np.load(path)
    st = {"rec.items": g["rec_items"], "rec.topk": g["rec_topk"], "rec.positive_score": g["rec_positive_score"],
          "data.positive_i": g["data_positive_i"], "data.sst": g["data_sst"]}
    count_items = {int(i): int(c) for i, c in g["train_count_items"]}
    res = mo.evaluate(st, [int(k) for k in g["topk"]], g["I"].shape[0], count_items, 0.1)
    for (k, v), ref in zip(res.items(), g["metric_values"]):
        assert abs(v - ref) <= RTOL * max(abs(ref), 1e-12) + 1e-12, (k, v, ref)


def test_c_oracle_matches_numpy_chain():
    if fs.clib() is None:
        pytest.skip("oracle C library not built")
    rng = np.random.default_rng(0)
    U = rng.standard_normal((20, 24)).astype(np.float32)
    I = rng.standard_normal((33, 24)).astype(np.float32)
    users = np.arange(1, 20)
    a = fs.full_sort_scores(U, I, users, 5.0)
    acc = np.zeros((19, 33), np.float32)
    for k in range(24):
        acc = (acc.astype(np.float64) + U[users, k, None].astype(np.float64) * I[None, :, k]).astype(np.float32)
    b = (np.clip(acc, 0, 5) / np.float32(5)).astype(np.float32)
    np.testing.assert_array_equal(a, b)
    hist_off = np.arange(0, 20 * 3, 3)[:20]
    hist_items = rng.integers(1, 33, size=hist_off[-1])
    ids, vals = fs.threaded_topk(U, I, users, 5.0, hist_off, hist_items, 7)
    ids2, vals2 = fs.topk_canonical(fs.mask_history(a, hist_off, hist_items), 7)
    np.testing.assert_array_equal(ids, ids2)
    np.testing.assert_array_equal(vals, vals2)


@pytest.mark.parametrize("name", ["value", "absolute", "nonparity", "none"])
def test_torch_port_matches_reference(name):
    """the timed CPU baseline (oracle/torch_port.py) computes what the reference computes"""
    import torch
    from oracle import torch_port as tp
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"focf_train_{name}.npz"))
    model = tp.TorchFOCF(g["U0"], g["I0"], str(g["objective"]), float(g["fair_weight"]))
    pre = [tuple(torch.as_tensor(g[f"{k}{s}"]) for k in ("uid", "iid", "rating", "sst")) for s in range(3)]
    rows, total = tp.train_steps(model, None, 3, float(g["lr"]), float(g["wd"]), prebuilt=pre)
    np.testing.assert_allclose(total, float(g["losses"].astype(np.float64).sum()), rtol=1e-6)
    assert rel_err(model.user_emb.weight.detach().numpy(), g["U_final"]) < 1e-6
    assert rel_err(model.item_emb.weight.detach().numpy(), g["I_final"]) < 1e-6


def test_torch_port_eval_matches_reference():
    import torch
    from oracle import torch_port as tp
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "focf_eval_tiefree.npz"))
    model = tp.TorchFOCF(g["U"], g["I"], "none", 1.0, float(g["max_rating"]))
    users = g["eval_users"]
    hist = [g["hist_items"][g["hist_off"][r]:g["hist_off"][r + 1]] for r in range(len(users))]
    pos = [g["pos_items"][g["pos_off"][r]:g["pos_off"][r + 1]] for r in range(len(users))]
    count_items = {int(i): int(c) for i, c in g["train_count_items"]}
    res, st = tp.evaluate(model, users, hist, pos, g["sst_of_user"], g["I"].shape[0], [int(k) for k in g["topk"]],
                          count_items, 7)
    np.testing.assert_array_equal(st["rec.items"], g["rec_items"])
    for (k, v), ref in zip(res.items(), g["metric_values"]):
        assert abs(v - ref) <= 1e-6 * max(abs(ref), 1e-12) + 1e-12, (k, v, ref)


NFCF = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "nfcf_train_*.npz")))


@pytest.mark.parametrize("path", NFCF, ids=[os.path.basename(p)[11:-4] for p in NFCF])
def test_nfcf_oracle_matches_reference(path):
    from oracle import nfcf_oracle as no
    g = np.load(path)
    L = int(g["n_layers"])
    Ws, bs = [g[f"W{k}_0"] for k in range(L)], [g[f"b{k}_0"] for k in range(L)]
    batches = [(g[f"uid{s}"], g[f"iid{s}"], g[f"label{s}"], g[f"sst{s}"]) for s in range(int(g["n_steps"]))]
    fair, fw = bool(g["fair"]), float(g["fair_weight"])
    loss, p, dU, dI, dWs, dbs = no.loss_and_grads(g["U0"], g["I0"], Ws, bs, *batches[0], fair, fw)
    assert rel_err(p, g["pred0"]) < RTOL
    np.testing.assert_allclose(loss, g["losses"][0], rtol=RTOL)
    assert rel_err(dI, g["dI0"]) < RTOL
    if "dU0" in g:
        assert rel_err(dU, g["dU0"]) < RTOL
    for k in range(L):
        assert rel_err(dWs[k], g[f"dW{k}_0"]) < RTOL and rel_err(dbs[k], g[f"db{k}_0"]) < RTOL, k
    losses, U, I, Wf, bf = no.train_steps(g["U0"], g["I0"], Ws, bs, batches, fair, fw, float(g["lr"]), float(g["wd"]),
                                          bool(g["user_frozen"]))
    np.testing.assert_allclose(losses, g["losses"], rtol=RTOL)
    assert rel_err(U, g["U_final"]) < RTOL and rel_err(I, g["I_final"]) < RTOL
    for k in range(L):
        assert rel_err(Wf[k], g[f"W{k}_final"]) < RTOL and rel_err(bf[k], g[f"b{k}_final"]) < RTOL, k


PFCN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "pfcn_*.npz")))


@pytest.mark.parametrize("path", PFCN, ids=[os.path.basename(p)[5:-4] for p in PFCN])
def test_pfcn_oracle_matches_reference(path):
    """oracle/pfcn_oracle.py replayed over the alternating filter / discriminator schedule of the fixture"""
    from oracle import pfcn_oracle as po
    g = np.load(path)
    losses, grads0, pred0, final = po.replay(g)
    np.testing.assert_allclose(losses, g["losses"], rtol=RTOL)
    assert rel_err(pred0, g["predict0"]) < RTOL
    n = 0
    for k in g.files:
        if k.startswith("grad_") and k.endswith("@0"):
            assert rel_err(grads0[k[5:-2]], g[k]) < RTOL, k
            n += 1
        if k.endswith("@final") and "num_batches_tracked" not in k:
            assert rel_err(final[k[:-6]], g[k]) < RTOL, k
            n += 1
    assert n > 40


FAIRGO = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "fairgo_*.npz")))


@pytest.mark.parametrize("path", FAIRGO, ids=[os.path.basename(p)[7:-4] for p in FAIRGO])
def test_fairgo_oracle_matches_reference(path):
    """oracle/fairgo_oracle.py replayed over the fixture's pretrain + alternating fine-tune schedule"""
    from oracle import fairgo_oracle as go
    g = np.load(path)
    L = go.norm_matrix(g["train_u"], g["train_i"], g["train_r"], int(g["n_users"]), int(g["n_items"])).tocsr()
    import scipy.sparse as sp
    ref = sp.csr_matrix((g["norm_val"], (g["norm_row"], g["norm_col"])), shape=L.shape)
    assert abs(L - ref).max() <= 1e-6 * abs(ref).max() and (L != 0).sum() == (ref != 0).sum()
    losses, grads, extra, pretrained, final = go.replay(g)
    np.testing.assert_allclose(losses, g["losses"], rtol=RTOL)
    assert rel_err(extra["predict"], g["predict_ft0"]) < RTOL
    assert rel_err(extra["full_sort"], g["full_sort_ft0"]) < RTOL
    n = 0
    for k in g.files:
        if k.startswith("grad_") and k.endswith("@ft0"):
            assert rel_err(grads[k[5:-4]], g[k]) < RTOL, k
            n += 1
        if k.endswith("@pretrained"):
            assert rel_err(pretrained[k[:-11]], g[k]) < RTOL, k
        if k.endswith("@final"):
            assert rel_err(final[k[:-6]], g[k]) < RTOL, k
            n += 1
    assert n > 20


UNI = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "uni_eval_*.npz")))


@pytest.mark.parametrize("path", UNI, ids=[os.path.basename(p)[9:-4] for p in UNI])
def test_sampled_eval_oracle_matches_reference(path):
    """oracle/sampled_oracle.py (uni100 ranking evaluation) against the reference's _neg_sample_batch_eval + Collector +
    Evaluator; rows whose top-(K+1) scores are separated by > 1e-6 bit-exactly, tied rows by the canonical rule"""
    from oracle import sampled_oracle as so
    g = np.load(path)
    users, K = g["eval_users"], int(max(g["topk"]))
    cands = so.candidate_lists(g["pos_off"], g["pos_items"], g["neg_items"], int(g["neg_num"]))
    rows = so.dense_rows(g["U"], g["I"], users, cands, g["I"].shape[0], float(g["max_rating"]))
    ids, rec_topk, pos_score = so.collect(rows, cands, K)
    np.testing.assert_allclose(pos_score, g["rec_positive_score"], rtol=1e-5, atol=1e-7)
    np.testing.assert_array_equal(rec_topk[:, K], g["rec_topk"][:, K])
    _, vals = fs.topk_canonical(rows, K + 1)
    clean = np.all(np.abs(np.diff(vals, axis=1)) > 1e-6, axis=1) & np.isfinite(vals).all(axis=1)
    assert clean.sum() >= len(users) // 2
    np.testing.assert_array_equal(ids[clean], g["rec_items"][clean])
    np.testing.assert_array_equal(rec_topk[clean], g["rec_topk"][clean])
    if clean.all():
        count_items = {int(i): int(c) for i, c in g["train_count_items"]}
        # the reference attributes the sensitive attribute per BATCH ROW (collector.py:203-205): equal to the per-user
        # attribution only for single-user batches (fixture uni100); restated for the batched fixtures
        upb = int(g["users_per_batch"]) if "users_per_batch" in g else \
            {"uni100": 1, "uni100_batched": 3, "uni20_small_catalog": 2}[os.path.basename(path)[9:-4]]
        sst_of_pos = so.reference_sst_of_pos(users, cands, g["sst_of_user"], upb)
        if upb == 1:
            np.testing.assert_array_equal(sst_of_pos, np.repeat(g["sst_of_user"][users], np.diff(g["pos_off"])))
        res = so.metrics(ids, rec_topk, pos_score, g["pos_items"], sst_of_pos, [int(k) for k in g["topk"]],
                         g["I"].shape[0], count_items)
        if any("Value Unfairness" in str(k) for k in g["metric_names"]):
            # sampled mode of the four unfairness metrics: each positive paired with its FIRST negative (collector.py:190-199)
            n_pos = np.diff(g["pos_off"])
            neg_items = np.concatenate([np.asarray(c[1])[:p] for c, p in zip(cands, n_pos)])
            neg_score = np.concatenate([rows[r, np.asarray(c[1])[:p]] for r, (c, p) in enumerate(zip(cands, n_pos))])
            vals4 = so.unfairness_sampled(pos_score, g["pos_items"], neg_score, neg_items, sst_of_pos)
            for name, v in zip(("Value", "Absolute", "Underestimation", "Overestimation"), vals4):
                res[f"{name} Unfairness of sensitive attribute gender"] = v
        for k, ref in zip(g["metric_names"], g["metric_values"]):
            assert abs(res[str(k)] - ref) <= 1e-5 * max(abs(ref), 1e-12) + 1e-9, (k, res[str(k)], ref)


def test_gcn_restatement_equals_the_dense_formula_and_the_products_adjacency():
    """oracle/fairgo_oracle.gcn_forward (edge-list form of GCNConv: self loops, symmetric normalisation by in-degree,
    aggregate x W^T, + b, act between layers) == relu(A_hat (x W0^T) + b0) ... with the DENSE A_hat = D^-1/2 (A + I) D^-1/2,
    and the adjacency the product builds on the host (fairgo.gcn_norm_csr) is that same matrix.  PARITY UNPINNED against
    torch_geometric itself (absent from the build container): this pins the two restatements on each other only."""
    import scipy.sparse as sp
    import torch
    from oracle import fairgo_oracle as go
    from recbole_fairrec_b200.fairgo import gcn_norm_csr
    rng = np.random.default_rng(0)
    nu, ni, d = 30, 20, 8
    tu, ti = rng.integers(1, nu, 200), rng.integers(1, ni, 200)          # repeated pairs included
    tr = rng.integers(1, 6, 200).astype(np.float32)
    N = nu + ni
    A = np.zeros((N, N))
    for u, i, r in zip(tu, ti, tr):
        A[u, nu + i] += r
        A[nu + i, u] += r
    A += np.eye(N)
    dinv = A.sum(1) ** -0.5
    A_hat = dinv[:, None] * A * dinv[None, :]
    got = gcn_norm_csr(sp.coo_matrix((tr, (tu, ti)), shape=(nu, ni)), nu, ni).toarray()
    np.testing.assert_allclose(got, A_hat, rtol=2e-7, atol=1e-9)             # stored in float32
    x = torch.randn(N, d, dtype=torch.float64)
    W = [torch.randn(16, d, dtype=torch.float64), torch.randn(12, 16, dtype=torch.float64), torch.randn(d, 12, dtype=torch.float64)]
    b = [torch.randn(16, dtype=torch.float64), torch.randn(12, dtype=torch.float64), torch.randn(d, dtype=torch.float64)]
    ei, ew = go.gcn_edges(tu, ti, tr, nu, ni)
    At = torch.from_numpy(A_hat)
    h = x
    for k in range(3):
        h = At @ (h @ W[k].t()) + b[k]
        if k < 2:
            h = torch.relu(h)
    out = go.gcn_forward(x, ei, ew.double(), W, b)
    assert (out - h).abs().max().item() < 1e-12
    # isolated nodes keep their self loop: user 0 ([PAD]) has no edges -> out row = W-chain of its own x
    assert A_hat[0, 0] == 1.0 and np.count_nonzero(A_hat[0]) == 1
