"""GPU parity of the FairGo family (full-table filter MLPs on fr_linear_*, D^-1 A aggregation on fr_spmm_csr, WAP / LBA /
LVA heads, fr_mse_loss / fr_sigmoid_bce_loss / fr_softmax_ce_loss, fr_adam_multi -- all through the C ABI) against the
fixtures generated from the unmodified reference (tests/golden/fairgo_*.npz) and against oracle/fairgo_oracle.py."""
import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import fairgo_oracle as go

pytestmark = pytest.mark.gpu
RTOL = 1e-5          # north star: losses and updated parameters within 1e-5 relative
HERE = os.path.dirname(__file__)
FAIRGO = sorted(glob.glob(os.path.join(HERE, "golden", "fairgo_*.npz")))


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


class GraphDataset:
    def __init__(self, n_users, n_items, feats, tu, ti, tr):
        import recbole_fairrec_b200 as pkg
        self._n = {"user_id": n_users, "item_id": n_items}
        self._feat = pkg.Interaction({"user_id": torch.arange(n_users), **{k: torch.from_numpy(v) for k, v in feats.items()}})
        self._coo = sp.coo_matrix((tr, (tu, ti)), shape=(n_users, n_items))
        self.inter_feat = {"rating": torch.tensor([1.0, 5.0])}

    def num(self, f):
        return self._n[f]

    def get_user_feature(self):
        return self._feat

    def inter_matrix(self, form="coo", value_field=None):
        return self._coo


def owners(model):
    out = {f"filter_{k}": m for k, m in model.filter_layer_dict.items()}
    out.update({f"dis_{k}": m for k, m in model.dis_layer_dict.items()})
    return out


def load_state(model, st):
    with torch.no_grad():
        model.load_state_dict({k[5:]: v for k, v in st.items() if k.startswith("base.")})
        for name, mod in owners(model).items():
            mod.load_state_dict({k[len(name) + 1:]: v for k, v in st.items() if k.startswith(name + ".")})


def dump_state(model):
    out = {f"base.{k}": v for k, v in model.state_dict().items()}
    for name, mod in owners(model).items():
        out.update({f"{name}.{k}": v for k, v in mod.state_dict().items()})
    return {k: v.detach().cpu().numpy() for k, v in out.items()}


def build(g, cls="FairGo_PMF", **kw):
    import recbole_fairrec_b200 as pkg
    feats = {"gender": np.array(g["gender"]), "age": np.array(g["age"])}
    cfg = pkg.Config(embedding_size=int(g["d"]), sst_attr_list=["gender", "age"], n_layers=int(g["n_layers"]),
                     activation="leakyrelu", dis_hidden_size_list=[16, 8], filter_hidden_size_list=[32, 16],
                     fair_weight=float(g["fair_weight"]), load_pretrain_weight=False, aggr_method=str(g["aggr"]),
                     vs_weights=[float(x) for x in g["vs_weights"]], device=torch.device("cuda"), learning_rate=1e-3,
                     weight_decay=1e-4, train_epoch_interval=1, pretrain_epochs=1, **kw)
    ds = GraphDataset(int(g["n_users"]), int(g["n_items"]), feats, g["train_u"], g["train_i"], g["train_r"])
    model = getattr(pkg, cls)(cfg, ds).to(torch.device("cuda"))
    return cfg, model, feats


@pytest.mark.parametrize("path", FAIRGO, ids=[os.path.basename(p)[7:-4] for p in FAIRGO])
def test_fairgo_matches_reference(path):
    import recbole_fairrec_b200 as pkg
    g = np.load(path)
    cfg, model, feats = build(g)
    load_state(model, go.load_state(g, "init"))
    # the normalised adjacency equals the reference's (fairgo_pmf.py:100-127)
    ref = sp.csr_matrix((g["norm_val"], (g["norm_row"], g["norm_col"])), shape=model._norm_csr.shape)
    assert abs(model._norm_csr - ref).max() <= 1e-6 * abs(ref).max()
    trainer = pkg.FairGoTrainer(cfg, model)
    assert model.train_stage == "pretrain"
    model.train()
    P = int(g["pretrain_steps"])
    final64 = go.replay(g, dtype=torch.float64)[4]
    losses = []
    for s in range(P + 2 * int(g["n_rounds"])):
        u = g[f"user_id{s}"]
        inter = pkg.Interaction({"user_id": torch.from_numpy(u), "item_id": torch.from_numpy(g[f"item_id{s}"]),
                                 "rating": torch.from_numpy(g[f"rating{s}"]),
                                 "gender": torch.from_numpy(feats["gender"][u]), "age": torch.from_numpy(feats["age"][u])})
        sst_list = [str(x) for x in g[f"sst_list{s}"]]
        if s < P:
            fn, opt = model.calculate_loss, trainer.optimizer_pretrain
        else:
            if s == P:
                model.train_stage = "finetune"
                model._ego = None
                mid = dump_state(model)
                for k in g.files:
                    if k.endswith("@pretrained"):
                        assert rel_err(mid[k[:-11]], g[k]) < RTOL, k
            fn, opt = ((model.calculate_loss, trainer.optimizer_filter) if (s - P) % 2 == 0 else
                       (model.calculate_dis_loss, trainer.optimizer_dis))
        opt.zero_grad()
        loss = fn(inter, sst_list if s >= P else None)
        loss.backward()
        if s == P:
            n = 0
            for name, mod in owners(model).items():
                for kk, p in mod.named_parameters():
                    key = f"grad_{name}.{kk}@ft0"
                    if key in g.files:
                        assert rel_err(p.grad.cpu().numpy(), g[key]) < RTOL, key
                        n += 1
            assert n >= 6
            assert rel_err(model.predict(inter).cpu().numpy(), g["predict_ft0"]) < RTOL
            fs = model.full_sort_predict(pkg.Interaction({"user_id": torch.tensor([1, 2, 5])}))
            assert rel_err(fs.cpu().numpy(), g["full_sort_ft0"]) < RTOL
        opt.step()
        losses.append(loss.item())
    np.testing.assert_allclose(losses, g["losses"], rtol=RTOL)
    final = dump_state(model)
    n = 0
    for k in g.files:
        if k.endswith("@final"):
            tol = max(RTOL, 3.0 * rel_err(g[k], final64[k[:-6]]))      # see tests/test_pfcn_gpu.py
            assert rel_err(final[k[:-6]], final64[k[:-6]]) < tol, (k, tol)
            n += 1
    assert n > 20


def test_spmm_matches_scipy_on_skewed_rows():
    """fr_spmm_csr with rows from empty to thousands of nnz (multi-chunk partials), d = 64 and d = 20; A and A^T"""
    from recbole_fairrec_b200 import ops
    rng = np.random.default_rng(2)
    n = 3000
    lens = np.minimum((rng.lognormal(3.0, 1.6, n)).astype(np.int64), n)
    lens[:5] = [0, 1, 2999, 128, 129]
    rows = np.repeat(np.arange(n), lens)
    cols = np.concatenate([rng.choice(n, l, replace=False) for l in lens])
    A = sp.csr_matrix((rng.standard_normal(len(rows)).astype(np.float32), (rows, cols)), shape=(n, n))
    mat = ops.SpmmMatrix(A, torch.device("cuda"))
    for d in (64, 20):
        X = rng.standard_normal((n, d)).astype(np.float32)
        Xg = torch.from_numpy(X).cuda()
        for tr in (False, True):
            want = ((A.T if tr else A).astype(np.float64) @ X.astype(np.float64))
            got = mat.apply(Xg, tr).cpu().numpy()
            assert rel_err(got, want) < 2e-6
            assert np.array_equal(got, mat.apply(Xg, tr).cpu().numpy())      # run-to-run bit stability


def test_fairgo_gcn_finetune_equals_pmf():
    import recbole_fairrec_b200 as pkg
    g = np.load(FAIRGO[0])
    _, pmf, feats = build(g)
    _, gcn, _ = build(g, cls="FairGo_GCN", hidden_channels=32, gcn_n_layers=2, gcn_dropout=0.2, gcn_act="relu")
    st = go.load_state(g, "pretrained")
    load_state(pmf, st)
    gcn.load_state_dict({k[5:]: v for k, v in st.items() if k.startswith("base.")}, strict=False)   # + its own gcn.*
    for name, mod in owners(gcn).items():
        mod.load_state_dict({k[len(name) + 1:]: v for k, v in st.items() if k.startswith(name + ".")})
    assert sorted(k for k in gcn.state_dict() if k.startswith("gcn.")) == \
        ["gcn.convs.0.bias", "gcn.convs.0.lin.weight", "gcn.convs.1.bias", "gcn.convs.1.lin.weight"]
    u = g["user_id2"]
    inter = pkg.Interaction({"user_id": torch.from_numpy(u), "item_id": torch.from_numpy(g["item_id2"]),
                             "rating": torch.from_numpy(g["rating2"]), "gender": torch.from_numpy(feats["gender"][u]),
                             "age": torch.from_numpy(feats["age"][u])})
    pmf.train_stage = gcn.train_stage = "finetune"
    assert pmf.calculate_loss(inter, ["gender", "age"]).item() == gcn.calculate_loss(inter, ["gender", "age"]).item()


@pytest.mark.parametrize("layers,hidden", [(2, 32), (3, 16), (1, 32)])
def test_fairgo_gcn_pretrain_matches_the_restated_gcn(layers, hidden):
    """pretrain stage of FairGo_GCN (fairgo_gcn.py:175-176, 190-199): rating MSE on GCN(ego embeddings) -- loss, gradients
    of both tables and of every GCN parameter, and three Adam steps, against oracle/fairgo_oracle.gcn_forward (the
    published GCNConv algorithm in its own order A_hat (x W^T) + b, torch CPU autograd + torch.optim.Adam)"""
    import recbole_fairrec_b200 as pkg
    g = np.load(FAIRGO[0])
    cfg, model, feats = build(g, cls="FairGo_GCN", hidden_channels=hidden, gcn_n_layers=layers, gcn_dropout=0.0,
                              gcn_act="relu")
    trainer = pkg.FairGoTrainer(cfg, model)
    assert model.train_stage == "pretrain"
    nu, ni = int(g["n_users"]), int(g["n_items"])
    with torch.no_grad():
        for conv in model.gcn.convs:
            conv.bias.copy_(torch.linspace(-0.1, 0.1, conv.bias.numel()))       # non-trivial biases
    ref = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()
           if k.startswith("gcn.") or k.endswith("embedding_layer.weight")}
    ei, ew = go.gcn_edges(g["train_u"], g["train_i"], g["train_r"], nu, ni)
    names = [f"gcn.convs.{k}" for k in range(layers)]
    opt = torch.optim.Adam(list(ref.values()), lr=1e-3, weight_decay=1e-4)
    u, i, r = (torch.from_numpy(np.asarray(g[k])) for k in ("user_id2", "item_id2", "rating2"))
    inter = pkg.Interaction({"user_id": u, "item_id": i, "rating": r})
    model.train()
    for step in range(3):
        x = torch.cat([ref["user_embedding_layer.weight"], ref["item_embedding_layer.weight"]])
        out = go.gcn_forward(x, ei, ew, [ref[n + ".lin.weight"] for n in names], [ref[n + ".bias"] for n in names])
        want = torch.nn.functional.mse_loss((out[u.long()] * out[nu + i.long()]).sum(-1), r.float())
        opt.zero_grad()
        want.backward()
        trainer.optimizer_pretrain.zero_grad()
        loss = model.calculate_loss(inter, None)
        loss.backward()
        assert abs(loss.item() - want.item()) <= RTOL * abs(want.item()), (step, loss.item(), want.item())
        if step == 0:
            for k, p in model.named_parameters():
                if k in ref:
                    assert rel_err(p.grad.cpu().numpy(), ref[k].grad.numpy()) < 2e-5, k
        opt.step()
        trainer.optimizer_pretrain.step()
    for k, p in model.named_parameters():
        if k in ref:
            # Adam's first steps move every element by ~lr whatever the gradient's size: compare the UPDATE
            assert np.abs(p.detach().cpu().numpy() - ref[k].detach().numpy()).max() < 1e-5, k
    # evaluation of the stage scores with the GCN output (full_sort_predict -> forward, fairgo_gcn.py:264-271)
    model.eval()
    U, I = model.filtered_tables()
    out = go.gcn_forward(torch.cat([ref["user_embedding_layer.weight"], ref["item_embedding_layer.weight"]]).detach(), ei, ew,
                         [ref[n + ".lin.weight"].detach() for n in names], [ref[n + ".bias"].detach() for n in names])
    assert rel_err(torch.cat([U, I]).cpu().numpy(), out.numpy()) < 1e-4


def test_dropout_op_mask_statistics_and_backward():
    from recbole_fairrec_b200 import ops
    x = torch.randn(4000, 32, device="cuda", requires_grad=True)
    y = ops.Dropout.apply(x, 0.2, 12345)
    kept = (y != 0)
    assert abs(kept.float().mean().item() - 0.8) < 0.01
    torch.testing.assert_close(y[kept], (x.detach() / 0.8)[kept])
    gy = torch.randn_like(y)
    y.backward(gy)
    torch.testing.assert_close(x.grad, torch.where(kept, gy / 0.8, torch.zeros_like(gy)))
    y2 = ops.Dropout.apply(x.detach(), 0.2, 12346)
    assert ((y2 != 0) != kept).float().mean().item() > 0.2          # another seed, another mask
    model_input = torch.ones(8, 4, device="cuda")
    assert torch.equal(ops.Dropout.apply(model_input, 0.0, 1), model_input)


def test_fairgo_trainer_runs_ml1m_widths():
    """ML-1M widths (d=64, filters [64,128,64,64], discriminators [64,16,8,4,{1|C}], LBA,
    2 layers): pretrain + alternating epochs run, losses finite, the filtered tables feed the fused full-sort evaluator"""
    import recbole_fairrec_b200 as pkg
    rng = np.random.default_rng(4)
    nu, ni, d, n_inter, B = 1500, 700, 64, 60000, 2048
    pairs = rng.permutation((nu - 1) * (ni - 1))[:n_inter]
    tu, ti = (pairs // (ni - 1) + 1).astype(np.int64), (pairs % (ni - 1) + 1).astype(np.int64)
    tr = rng.integers(1, 6, n_inter).astype(np.float32)
    feats = {"gender": rng.integers(0, 2, nu).astype(np.float32), "age": rng.integers(0, 7, nu).astype(np.float32)}
    cfg = pkg.Config(embedding_size=d, sst_attr_list=list(feats), n_layers=2, activation="leakyrelu",
                     dis_hidden_size_list=[16, 8, 4], filter_hidden_size_list=[128, 64], fair_weight=0.1,
                     load_pretrain_weight=False, aggr_method="LBA", vs_weights=[4, 1], device=torch.device("cuda"),
                     learning_rate=1e-3, weight_decay=1e-4, train_epoch_interval=1, pretrain_epochs=2)
    torch.manual_seed(0)
    np.random.seed(0)
    model = pkg.FairGo_PMF(cfg, GraphDataset(nu, ni, feats, tu, ti, tr)).to(torch.device("cuda"))
    trainer = pkg.FairGoTrainer(cfg, model)

    def batches():
        for b in range(0, n_inter, B):
            u = tu[b:b + B]
            yield pkg.Interaction({"user_id": torch.from_numpy(u), "item_id": torch.from_numpy(ti[b:b + B]),
                                   "rating": torch.from_numpy(tr[b:b + B]),
                                   **{a: torch.from_numpy(feats[a][u]) for a in feats}})

    pl = trainer.pretrain(list(batches()))
    assert pl[1] < pl[0]
    for epoch in range(2):
        dl, fl = trainer._train_epoch(list(batches()), epoch)
        assert np.isfinite(dl) and np.isfinite(fl)
    U, I = model.filtered_tables()
    assert U.shape == (nu, d) and I.shape == (ni, d) and torch.isfinite(U).all() and torch.isfinite(I).all()


def test_fairgo_trainer_fit_and_fused_full_sort_eval():
    """FairGoTrainer.fit / evaluate: the fused full-sort evaluator on the filtered tables equals the oracle's canonical
    top-K on the dense scores of full_sort_predict"""
    import recbole_fairrec_b200 as pkg
    from oracle import fullsort_oracle as fs
    g = np.load(FAIRGO[0])
    cfg, model, feats = build(g, topk=[5], valid_metric="NDCG@5", epochs=2, stopping_step=5)
    load_state(model, go.load_state(g, "pretrained"))
    trainer = pkg.FairGoTrainer(cfg, model)
    model.train_stage = "finetune"
    nu, ni = int(g["n_users"]), int(g["n_items"])
    rng = np.random.default_rng(0)
    tu, ti = g["train_u"], g["train_i"]
    users = np.arange(1, nu)
    hist = [np.unique(ti[tu == u]) for u in users]
    pos = [np.setdiff1d(rng.choice(np.arange(1, ni), 4, replace=False), h)[:3] for h in hist]
    pos = [p if len(p) else np.setdiff1d(np.arange(1, ni), h)[:1] for p, h in zip(pos, hist)]
    data = pkg.EvalData(users, hist, pos, {"gender": feats["gender"], "age": feats["age"]}, torch.device("cuda"))

    def batches():
        for s in range(2, 6):
            u = g[f"user_id{s}"]
            yield pkg.Interaction({"user_id": torch.from_numpy(u), "item_id": torch.from_numpy(g[f"item_id{s}"]),
                                   "rating": torch.from_numpy(g[f"rating{s}"]),
                                   "gender": torch.from_numpy(feats["gender"][u]), "age": torch.from_numpy(feats["age"][u])})

    np.random.seed(1)
    cfg["sst_attr_list"] = ["gender", "age"]
    best, res = trainer.fit(list(batches()), data, train_item_count={int(i): int(c) for i, c in
                                                                       enumerate(np.bincount(ti, minlength=ni)) if c})
    assert res is not None and 0.0 <= res["ndcg@5"] <= 1.0 and np.isfinite(best)
    out = trainer.evaluate(data)
    dense = model.full_sort_predict(pkg.Interaction({"user_id": torch.from_numpy(users)})).view(len(users), ni).cpu().numpy()
    hist_off = np.r_[0, np.cumsum([len(h) for h in hist])]
    ids, _ = fs.topk_canonical(fs.mask_history(dense, hist_off, np.concatenate(hist)), 5)
    got = trainer.evaluator.last["topk_id"].cpu().numpy()
    # the dense scores come from a different summation order than the scorer's fma chain: compare outside near-ties
    srt = -np.sort(-fs.mask_history(dense, hist_off, np.concatenate(hist)), axis=1)[:, :6]
    clear = (np.abs(np.diff(srt, axis=1)) > 1e-6).all(axis=1)
    assert clear.sum() > len(users) // 2 and np.array_equal(got[clear], ids[clear])
    assert set(out) >= {"ndcg@5", "giniindex@5", "Differential Fairness of sensitive attribute gender"}
