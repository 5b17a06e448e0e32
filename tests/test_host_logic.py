"""CPU tests of the host-side logic around the kernels (no CUDA calls): the SpMM chunk planner of the C ABI, the
normalised FairGo adjacency, negative sampling / candidate layout of the sampled evaluation, and the module trees
(state_dict names) of the MLP families against the names recorded from the reference."""
import ctypes
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

HERE = os.path.dirname(__file__)


def test_spmm_planner_covers_every_row_once():
    from recbole_fairrec_b200._lib import load
    lib = load()
    rng = np.random.default_rng(0)
    lens = np.minimum(rng.lognormal(3.0, 1.8, 500).astype(np.int64), 5000)
    lens[:4] = [0, 1, 128, 129]
    row_off = np.zeros(len(lens) + 1, np.int64)
    row_off[1:] = np.cumsum(lens)
    chunk = 128
    sizes = [ctypes.c_int64() for _ in range(4)]
    assert lib.fr_spmm_plan_sizes(row_off.ctypes.data, len(lens), chunk, *[ctypes.byref(s) for s in sizes]) == 0
    nc, nm, ns, ne = [s.value for s in sizes]
    assert nc == int(np.ceil(lens / chunk).sum()) and ne == int((lens == 0).sum())
    assert nm == int((lens > chunk).sum()) and ns == int(np.ceil(lens[lens > chunk] / chunk).sum())
    host = [np.zeros(max(nc, 1), np.int32) for _ in range(4)] + [np.zeros(max(nm, 1), np.int32), np.zeros(nm + 1, np.int32),
                                                                np.zeros(max(ne, 1), np.int32)]
    assert lib.fr_spmm_plan_fill(row_off.ctypes.data, len(lens), chunk, *[h.ctypes.data for h in host]) == 0
    crow, cbeg, cend, cslot, mrow, mfirst, erow = host
    covered = np.zeros(row_off[-1], np.int32)
    for c in range(nc):
        assert 0 < cend[c] - cbeg[c] <= chunk
        assert row_off[crow[c]] <= cbeg[c] and cend[c] <= row_off[crow[c] + 1]
        covered[cbeg[c]:cend[c]] += 1
    assert (covered == 1).all()
    # multi-chunk rows own consecutive slots, in chunk order; single-chunk rows write Y directly (slot -1)
    for r in range(nm):
        sl = [cslot[c] for c in range(nc) if crow[c] == mrow[r]]
        assert sl == list(range(mfirst[r], mfirst[r + 1]))
    single = lens[crow[:nc]] <= chunk
    assert (cslot[:nc][single] == -1).all() and (cslot[:nc][~single] >= 0).all()
    assert sorted(erow[:ne].tolist()) == np.nonzero(lens == 0)[0].tolist()


def test_fairgo_normalised_adjacency_matches_the_oracle():
    from oracle import fairgo_oracle as go
    from recbole_fairrec_b200.fairgo import norm_rating_csr
    rng = np.random.default_rng(1)
    nu, ni, n = 40, 30, 300
    pairs = rng.permutation((nu - 1) * (ni - 1))[:n]
    tu, ti = pairs // (ni - 1) + 1, pairs % (ni - 1) + 1
    tr = rng.integers(1, 6, n).astype(np.float32)
    mine = norm_rating_csr(sp.coo_matrix((tr, (tu, ti)), shape=(nu, ni)), nu, ni)
    ref = go.norm_matrix(tu, ti, tr, nu, ni).tocsr()
    assert abs(mine - ref).max() <= 1e-7 and mine.nnz == ref.nnz == 2 * n
    assert np.allclose(np.asarray(mine.sum(axis=1)).ravel()[np.asarray(mine.sum(axis=1)).ravel() > 0], 1.0, atol=1e-5)


def test_sample_negatives_and_candidate_layout():
    import recbole_fairrec_b200 as pkg
    rng = np.random.default_rng(2)
    n_items = 50
    pos = [np.array([3, 7]), np.array([10])]
    used = [np.array([1, 2, 3, 7, 8]), np.arange(1, 45)]
    neg = pkg.sample_negatives(pos, used, n_items, 20, rng)
    for p, u, q in zip(pos, used, neg):
        assert len(q) == 20 * len(p) and q.min() >= 1 and q.max() < n_items
        assert not np.isin(q, u).any() and not np.isin(q, p).any()
    data = pkg.SampledEvalData([4, 9], pos, neg, {"gender": np.arange(12) % 2}, torch.device("cpu"))
    off = data.cand_off.numpy()
    assert off.tolist() == [0, 2 + 40, 2 + 40 + 1 + 20] and data.n_pos_of_user.tolist() == [2, 1]
    assert data.cand_items[:2].tolist() == [3, 7] and data.cand_items[off[1]].item() == 10
    assert data.cand_uid[:42].unique().tolist() == [4] and data.pos_idx.tolist() == [0, 1, 42]
    assert data.group_of_pos["gender"].tolist() == [0, 0, 1] and data.n_groups["gender"] == 2


@pytest.mark.parametrize("fixture,prefix,layers,kw", [
    ("pfcn_mlp_sm.npz", "filter_1", [16, 32, 16], dict(activation="leakyrelu", bn=True, init_method="norm")),
    ("pfcn_mlp_sm.npz", "dis_age", [16, 32, 16, 4], dict(activation="leakyrelu", bn=True, init_method="norm")),
    ("pfcn_mlp_sm.npz", "base.mlp_layer", [32, 16, 8, 1], dict()),
    ("fairgo_pmf_lba.npz", "filter_gender", [16, 32, 16, 16], dict(activation="leakyrelu")),
])
def test_mlp_module_tree_has_the_reference_state_dict_names(fixture, prefix, layers, kw):
    """checkpoints are interchangeable: MLPLayers here exposes exactly the parameter / buffer names (and shapes) that the
    reference's MLPLayers (recbole/model/layers.py:30-85) recorded in the golden fixtures"""
    from recbole_fairrec_b200.layers import MLPLayers
    g = np.load(os.path.join(HERE, "golden", fixture))
    want = {k[len(prefix) + 1:-5]: g[k].shape for k in g.files if k.startswith(prefix + ".") and k.endswith("@init")}
    got = {k: tuple(v.shape) for k, v in MLPLayers(layers, **kw).state_dict().items()}
    assert got == {k: tuple(v) for k, v in want.items()}


def test_run_recbole_config_precedence_and_seeding(tmp_path):
    """configurator.py:211-263 as far as the fairness configs need it: defaults < YAML files (in order) < config_dict;
    init_seed seeds python / numpy / torch (utils.py:172-189)"""
    import random
    import yaml
    from recbole_fairrec_b200.quick_start import build_config, init_seed
    a, b = tmp_path / "a.yaml", tmp_path / "b.yaml"
    a.write_text(yaml.safe_dump({"embedding_size": 32, "topk": [3], "fair_weight": 0.5}))
    b.write_text(yaml.safe_dump({"topk": [7], "learning_rate": 0.01}))
    cfg = build_config("FOCF", "ml-100k", [str(a), str(b)], {"learning_rate": 0.1, "device": "cpu"})
    assert (cfg["embedding_size"], cfg["topk"], cfg["fair_weight"], cfg["learning_rate"]) == (32, [7], 0.5, 0.1)
    assert cfg["model"] == "FOCF" and cfg["dataset"] == "ml-100k" and cfg["train_batch_size"] == 2048
    assert cfg["no_such_key"] is None                       # configurator.py:405-409
    # model property defaults sit between the built-ins and the files; --key=value overrides sit on top
    cfg = build_config("NFCF", "ml-100k", [str(b)], {"dropout": 0.5, "device": "cpu"},
                       argv=["--fair_weight=0.3", "--mlp_hidden_size=[32,16]", "--fair_objective=value", "--junk"])
    assert (cfg["mlp_hidden_size"], cfg["dropout"], cfg["fair_weight"], cfg["weight_decay"]) == ([32, 16], 0.5, 0.3, 1e-6)
    assert cfg["topk"] == [7] and cfg["fair_objective"] == "value"
    assert build_config("FairGo_GCN", "x", None, {"device": "cpu"})["hidden_channels"] == 32
    init_seed(123)
    x = (random.random(), np.random.rand(), torch.rand(1).item())
    init_seed(123)
    assert x == (random.random(), np.random.rand(), torch.rand(1).item())


def test_pairwise_loader_negatives_are_unseen_items():
    """abstract_dataloader.py:182-188 + sampler.py:145-197: one uniform negative per positive, never an item of the
    user's train split, never the [PAD] item; every train row appears exactly once per epoch"""
    import os
    from recbole_fairrec_b200.atomic import AtomicDataset
    from recbole_fairrec_b200.quick_start import BatchLoader, build_config, init_seed
    cfg = build_config("PFCN_PMF", "ml-100k", None, dict(
        data_path=os.path.join(os.path.dirname(__file__), "data"), sst_attr_list=["gender"], device="cpu",
        load_col={"inter": ["user_id", "item_id", "rating"], "user": ["user_id", "gender"]}, train_batch_size=4096))
    init_seed(5)
    ds = AtomicDataset(cfg)
    train = ds.build()[0]
    used = set((train["user_id"] * ds.item_num + train["item_id"]).tolist())
    seen = []
    for b in BatchLoader(cfg, ds, train, pairwise=True):
        u, i, n = b["user_id"].numpy(), b["item_id"].numpy(), b["neg_item_id"].numpy()
        assert (n >= 1).all() and (n < ds.item_num).all()
        assert not (set((u * ds.item_num + n).tolist()) & used)
        assert (b["gender"].numpy() == ds.user_feat["gender"][u]).all()
        seen.append(u * ds.item_num + i)
    seen = np.concatenate(seen)
    assert len(seen) == len(train["user_id"]) and set(seen.tolist()) == used


class _StubDataset:
    def __init__(self, nu, ni, feats, coo=None):
        import recbole_fairrec_b200 as pkg
        self._n = {"user_id": nu, "item_id": ni}
        self._feat = pkg.Interaction({"user_id": torch.arange(nu), **{k: torch.from_numpy(v) for k, v in feats.items()}})
        self._coo = coo
        self.inter_feat = {"rating": torch.tensor([1.0, 5.0])}

    def num(self, f):
        return self._n[f]

    def get_user_feature(self):
        return self._feat

    def inter_matrix(self, form="coo", value_field=None):
        return self._coo


def _family(name, seed):
    """(config, model, trainer) of one MLP family on the CPU: construction, state and optimizer bookkeeping need no kernel"""
    import recbole_fairrec_b200 as pkg
    rng = np.random.default_rng(1)
    nu, ni = 40, 30
    feats = {"gender": (rng.random(nu) < 0.4).astype(np.float32), "age": rng.integers(0, 3, nu).astype(np.float32)}
    coo = sp.coo_matrix((rng.integers(1, 6, 200).astype(np.float32), (rng.integers(1, nu, 200), rng.integers(1, ni, 200))),
                        shape=(nu, ni))
    common = dict(embedding_size=16, sst_attr_list=["gender", "age"], device=torch.device("cpu"), learning_rate=1e-3,
                  weight_decay=1e-4, train_epoch_interval=1, activation="leakyrelu", model=name)
    torch.manual_seed(seed)
    if name == "PFCN_MLP":
        cfg = pkg.Config(filter_mode="sm", dropout=0.0, dis_dropout=0.0, dis_weight=1.0, dis_hidden_size_list=[16, 8],
                         mlp_hidden_size_list=[16, 8], **common)
        model = pkg.PFCN_MLP(cfg, _StubDataset(nu, ni, feats))
        return cfg, model, pkg.PFCNTrainer(cfg, model)
    if name == "FairGo_GCN":
        cfg = pkg.Config(n_layers=2, dis_hidden_size_list=[8, 4], filter_hidden_size_list=[16], fair_weight=0.1,
                         load_pretrain_weight=False, aggr_method="LBA", vs_weights=[4, 1], pretrain_epochs=1,
                         hidden_channels=8, gcn_n_layers=2, gcn_dropout=0.0, gcn_act="relu", **common)
        model = pkg.FairGo_GCN(cfg, _StubDataset(nu, ni, feats, coo))
        return cfg, model, pkg.FairGoTrainer(cfg, model)
    cfg = pkg.Config(mlp_hidden_size=[16, 8], dropout=0.0, load_pretrain_path=None, **dict(common, sst_attr_list=["gender"]))
    model = pkg.NFCF(cfg, _StubDataset(nu, ni, feats))
    return cfg, model, pkg.NFCFTrainer(cfg, model)


def _all_state(model, trainer):
    out = {f"model.{k}": v for k, v in model.state_dict().items()}
    for attr in ("filter_layer", "filter_layer_dict", "dis_layer_dict"):
        for k, m in (getattr(model, attr, None) or {}).items():
            out.update({f"{attr}.{k}.{kk}": v for kk, v in m.state_dict().items()})
    for name in ("optimizer", "optimizer_filter", "optimizer_dis", "optimizer_pretrain"):
        opt = getattr(trainer, name, None)
        if opt is not None:
            for i, st in opt.state_dict()["state"].items():
                out.update({f"{name}.{i}.{kk}": torch.as_tensor(v) for kk, v in st.items()})
    return out


def test_fairgo_trainer_loads_the_pretrain_checkpoint(tmp_path):
    """trainer.py:541-549: with `pretrain_model_file_path` the fine-tune stage starts from the checkpointed tables"""
    import recbole_fairrec_b200 as pkg
    cfg, model, _ = _family("FairGo_GCN", 1)
    want = {k: torch.randn_like(v) if v.is_floating_point() else v.clone() for k, v in model.state_dict().items()}
    path = str(tmp_path / "pretrain.pth")
    torch.save({"state_dict": want, "other_parameter": None}, path)
    cfg2, model2, _ = _family("FairGo_GCN", 2)
    assert not torch.equal(model2.user_embedding_layer.weight, want["user_embedding_layer.weight"])
    model2._ego = ("stale", None)
    cfg2["pretrain_model_file_path"] = path
    trainer = pkg.FairGoTrainer(cfg2, model2)
    assert model2.train_stage == "finetune" and model2._ego is None and not hasattr(trainer, "optimizer_pretrain")
    for k, v in model2.state_dict().items():
        assert torch.equal(v, want[k]), k
    cfg2["pretrain_model_file_path"] = str(tmp_path / "missing.pth")
    with pytest.raises(FileNotFoundError):
        pkg.FairGoTrainer(cfg2, model2)


def test_focf_fused_checkpoint_is_a_torch_adam_state_dict_and_resumes_in_place(tmp_path):
    """FOCFTrainer in fused mode (trainer.py:221-284): the 'optimizer' entry has torch.optim.Adam's own layout -- a
    torch.optim.Adam over the reference model's parameters loads it --, a resume copies INTO the existing moment tensors
    (captured graphs hold their addresses), and a state_dict written by torch.optim.Adam (the reference's files) resumes too."""
    import recbole_fairrec_b200 as pkg
    from recbole_fairrec_b200 import synth

    def make(seed):
        torch.manual_seed(seed)
        cfg = pkg.Config(embedding_size=8, fair_objective="value", device=torch.device("cpu"), learning_rate=2e-3,
                         weight_decay=1e-3, checkpoint_dir=str(tmp_path), model="FOCF")
        model = pkg.FOCF(cfg, synth.SynthDataset(30, 20, 5.0))
        return cfg, model, pkg.FOCFTrainer(cfg, model)

    cfg, model, trainer = make(1)
    assert trainer.fused and trainer.optimizer is None
    a = model._adam
    for k in ("mU", "vU", "mI", "vI"):
        a[k].uniform_()
    a["step"] = 7
    trainer.cur_step, trainer.best_valid_score = 1, 0.25
    trainer._save_checkpoint(3)
    ck = torch.load(trainer.saved_model_file, weights_only=False)
    opt = torch.optim.Adam([torch.nn.Parameter(torch.zeros(30, 8)), torch.nn.Parameter(torch.zeros(20, 8))], lr=1.0)
    opt.load_state_dict(ck["optimizer"])                       # what the reference's resume_checkpoint does
    assert opt.param_groups[0]["lr"] == 2e-3 and opt.param_groups[0]["weight_decay"] == 1e-3
    st = [opt.state[p] for p in opt.param_groups[0]["params"]]
    assert float(st[0]["step"]) == 7 and torch.equal(st[0]["exp_avg"], a["mU"]) and torch.equal(st[1]["exp_avg_sq"], a["vI"])
    # resume: in place, every field
    cfg2, model2, trainer2 = make(2)
    ptrs = {k: model2._adam[k].data_ptr() for k in ("mU", "vU", "mI", "vI")}
    trainer2.resume_checkpoint(trainer.saved_model_file)
    for k in ("mU", "vU", "mI", "vI"):
        assert model2._adam[k].data_ptr() == ptrs[k] and torch.equal(model2._adam[k], a[k]), k
    assert model2._adam["step"] == 7 and (trainer2.start_epoch, trainer2.cur_step, trainer2.best_valid_score) == (4, 1, 0.25)
    assert torch.equal(model2.user_embedding_layer.weight, model.user_embedding_layer.weight)
    # a file whose optimizer entry torch.optim.Adam itself wrote
    for p in opt.param_groups[0]["params"]:
        opt.state[p]["exp_avg"].add_(1.0)
    ck["optimizer"] = opt.state_dict()
    path = str(tmp_path / "ref.pth")
    torch.save(ck, path)
    cfg3, model3, trainer3 = make(3)
    trainer3.resume_checkpoint(path)
    assert torch.equal(model3._adam["mU"], a["mU"] + 1.0) and torch.equal(model3._adam["vI"], a["vI"]) and model3._adam["step"] == 7


@pytest.mark.parametrize("name", ["PFCN_MLP", "FairGo_GCN", "NFCF"])
def test_checkpoint_round_trip_of_the_mlp_family_trainers(name, tmp_path):
    """trainer.py:221-284 / 784-830 / 1133-1184: the reference's checkpoint keys, plus the dict-held filter / discriminator
    modules under 'dict_modules'; resume restores model, modules, every optimizer's moments and step counts, and the
    early-stopping bookkeeping"""
    cfg, model, trainer = _family(name, 1)
    for opt_name in ("optimizer", "optimizer_filter", "optimizer_dis", "optimizer_pretrain"):
        opt = getattr(trainer, opt_name, None)
        if opt is not None:                                  # pretend some steps happened
            opt.init_state()
            for k, p in enumerate(opt.params):
                if p in opt.state:
                    opt.state[p]["exp_avg"].normal_()
                    opt.state[p]["exp_avg_sq"].uniform_()
                    opt._steps[k] = 3 + k
    trainer.cur_step, trainer.best_valid_score = 2, 0.125
    path = trainer._save_checkpoint(7, str(tmp_path / "ck.pth"))
    ck = torch.load(path, weights_only=False)
    assert {"config", "epoch", "cur_step", "best_valid_score", "state_dict", "other_parameter", "optimizer"} <= set(ck)
    assert ck["epoch"] == 7 and ck["config"]["model"] == name
    if name != "NFCF":
        assert set(ck["dict_modules"]) and not any(k.startswith(("filter_layer", "dis_layer")) for k in ck["state_dict"])
    want = _all_state(model, trainer)
    cfg2, model2, trainer2 = _family(name, 2)                # other initial weights
    assert any(not torch.equal(v, _all_state(model2, trainer2).get(k, v + 1)) for k, v in want.items() if k.startswith("model."))
    trainer2.resume_checkpoint(path)
    got = _all_state(model2, trainer2)
    assert set(got) == set(want)
    for k, v in want.items():
        assert torch.equal(torch.as_tensor(got[k]).cpu(), torch.as_tensor(v).cpu()), k
    assert (trainer2.start_epoch, trainer2.cur_step, trainer2.best_valid_score) == (8, 2, 0.125)


@pytest.mark.parametrize("name", ["PFCN_MLP", "FairGo_GCN"])
def test_fit_loop_early_stopping_checkpoint_and_resume(name, tmp_path, monkeypatch):
    """the epoch loop around the (faked) train / evaluate calls: best score tracking, stop after `stopping_step`
    non-improving evaluations, check-point only on improvement, resume continues at the next epoch with the restored state"""
    cfg, model, trainer = _family(name, 3)
    cfg["stopping_step"], cfg["epochs"], cfg["valid_metric"] = 2, 20, "NDCG@5"
    scores = [0.10, 0.30, 0.20, 0.25, 0.50, 0.1, 0.1, 0.1, 0.9]
    calls = {"train": [], "saves": []}
    model.train_stage = "finetune"

    def fake_train(self, data, epoch):
        calls["train"].append(epoch)
        return (0.0, 0.0)

    def fake_eval(self, *a, **k):
        return {"ndcg@5": scores[len(calls["train"]) - 1]}

    cls = type(trainer)
    monkeypatch.setattr(cls, "_train_epoch", fake_train)
    monkeypatch.setattr(cls, "evaluate", fake_eval)
    orig_save = cls._save_checkpoint
    monkeypatch.setattr(cls, "_save_checkpoint", lambda self, epoch, f=None: calls["saves"].append(epoch) or
                        orig_save(self, epoch, str(tmp_path / "best.pth")))
    best, res = trainer.fit([None], [None], saved=True)
    # utils.py:97-140: stop once MORE than `stopping_step` evaluations in a row did not improve
    assert calls["train"] == list(range(8)) and calls["saves"] == [0, 1, 4]
    assert best == 0.50 and res == {"ndcg@5": 0.50} and trainer.cur_step == 3
    # resume from the best checkpoint (epoch 4): continues at epoch 5 with best = 0.50 and the counter as saved (0)
    cfg2, model2, trainer2 = _family(name, 4)
    cfg2["stopping_step"], cfg2["epochs"], cfg2["valid_metric"] = 3, 9, "NDCG@5"
    trainer2.resume_checkpoint(str(tmp_path / "best.pth"))
    assert (trainer2.start_epoch, trainer2.best_valid_score, trainer2.cur_step) == (5, 0.50, 0)
    calls["train"].clear()
    scores[:] = [0.2, 0.2, 0.6, 0.1]
    best, res = trainer2.fit([None], [None])
    assert calls["train"] == [5, 6, 7, 8] and best == 0.6


def test_device_built_eval_lists_equal_the_host_built_ones():
    """EvalData from per-user host lists (general_dataloader.py:173-207 order) vs EvalData.from_device from the raw split
    columns (`eval_lists: device` in run_recbole; torch ops only, so it runs on CPU tensors too): same users, CSR offsets,
    sorted histories / positives, groups; the positives' emission order differs (ascending item id on the device)"""
    import os
    import recbole_fairrec_b200 as pkg
    from recbole_fairrec_b200.atomic import AtomicDataset, used_and_positive_lists
    from recbole_fairrec_b200.quick_start import build_config, init_seed
    cfg = build_config("FOCF", "ml-100k", None, dict(
        data_path=os.path.join(os.path.dirname(__file__), "data"), sst_attr_list=["gender"], device="cpu",
        load_col={"inter": ["user_id", "item_id", "rating"], "user": ["user_id", "gender"]}))
    init_seed(3)
    ds = AtomicDataset(cfg)
    splits = ds.build()
    sst = {"gender": ds.user_feat["gender"]}
    cpu = torch.device("cpu")
    for phase in ("valid", "test"):
        users, hist, pos = used_and_positive_lists(splits, phase)
        a = pkg.EvalData(users, hist, pos, sst, cpu)
        ev = splits[1] if phase == "valid" else splits[2]
        used = [splits[0]] + ([splits[1]] if phase == "test" else [])
        t = lambda x: torch.as_tensor(np.ascontiguousarray(x))
        b = pkg.EvalData.from_device(t(np.concatenate([s["user_id"] for s in used])),
                                     t(np.concatenate([s["item_id"] for s in used])), t(ev["user_id"]), t(ev["item_id"]),
                                     {"gender": t(sst["gender"])}, ds.user_num, ds.item_num)
        assert a.n == b.n and a.n_pos == b.n_pos and a.n_groups == b.n_groups
        for f in ("users", "hist_off", "hist_items", "pos_off", "pos_items_sorted", "pos_row", "pos_uid"):
            assert torch.equal(getattr(a, f).to(torch.int64), getattr(b, f).to(torch.int64)), f
        assert torch.equal(b.pos_items, b.pos_items_sorted)
        # group of a positive depends on its user only, and pos_row is identical
        assert torch.equal(a.group_of_pos["gender"], b.group_of_pos["gender"])


def test_sampled_mode_accepts_the_twelve_metric_list_and_keeps_the_reference_keys():
    """the reference's 12-metric YAML list is usable with mode uni<N>: Value / Absolute / Under / Over unfairness take
    their sampled-mode definition (metrics.py:935-978 with mode != 'full'); result keys in the reference's order"""
    import recbole_fairrec_b200 as pkg
    from recbole_fairrec_b200.evaluator import FAIR_KEYS
    cfg = pkg.Config(topk=[2, 3], sst_attr_list=["gender"], device=torch.device("cpu"), metric_decimal_place=6)
    ev = pkg.SampledEvaluator(cfg, 10)

    class D:
        n = 4
    out = {"topk_sums": torch.tensor([[1.0, 2.0, 3.0]] * 4, dtype=torch.float64),
           "pop_hits": torch.tensor([2.0, 1.0, 1.0], dtype=torch.float64),
           "gini": {2: torch.tensor(0.25), 3: torch.tensor(0.5)},
           "fair": {"gender": torch.tensor([0.5, 0.1, 0.2, 0.3, 0.4, 0.125], dtype=torch.float64)}}
    res = ev.finalize(out, D())
    want = []
    for m in cfg["metrics"]:
        m = m.lower()
        want += [FAIR_KEYS[m].format("gender")] if m in FAIR_KEYS else [f"{m}@2", f"{m}@3"]
    assert list(res) == want
    for m, v in (("valueunfairness", 0.1), ("absoluteunfairness", 0.2), ("underunfairness", 0.3), ("overunfairness", 0.4)):
        assert res[FAIR_KEYS[m].format("gender")] == v
    assert res[FAIR_KEYS["differentialfairness"].format("gender")] == 0.5 and res["ndcg@3"] == 0.75
    # the pairing of each positive with its first negative (negatives are draw-major: [j * P + k])
    data = pkg.SampledEvalData([3, 5], [[7, 8], [2]], [[11, 12, 13, 14], [4, 6, 9]], {"gender": np.array([0, 1, 2, 1, 2, 1])},
                               torch.device("cpu"))
    assert data.cand_items[data.first_neg_idx].tolist() == [11, 12, 4] and data.cand_items[data.pos_idx].tolist() == [7, 8, 2]


def test_packed_fast_host_step_passes_the_same_arguments_as_the_generic_step(monkeypatch):
    """FocfEngine.train_step_packed (persistent staging buffer + persistent argument struct) vs the generic
    batch_columns + train_step path, with the library call intercepted: every field of `fr_focf_step` equal, the four
    batch pointers at the same offsets of the staged copy, the staged bytes equal to the host buffer, and the persistent
    struct refreshed when tables, moments, hyper-parameters or the workspace change.  (CPU tensors stand in for device
    memory here: the launch itself is stubbed.)"""
    import recbole_fairrec_b200 as pkg
    from recbole_fairrec_b200 import _lib, focf as focf_mod, kernels, synth
    from recbole_fairrec_b200.interaction import Interaction
    real = _lib.load()
    seen = []

    class FakeLib:
        fr_focf_workspace_bytes = staticmethod(real.fr_focf_workspace_bytes)

        def fr_focf_workspace_init(self, *a):
            return 0

        def fr_focf_train_step(self, ref, stream):
            s = ref._obj
            # (POINTER fields come back as fresh wrapper objects: compare what they point at)
            seen.append({name: (bool(getattr(s, name)) if name == "scalars_filled" else getattr(s, name))
                         for name, _ in _lib.FocfStep._fields_})
            return 0

    fake = FakeLib()
    monkeypatch.setattr(kernels, "load", lambda: fake)
    monkeypatch.setattr(kernels, "stream_ptr", lambda: 0)
    monkeypatch.setattr(kernels, "ptr", lambda t: None if t is None else t.data_ptr())
    cpu = torch.device("cpu")
    cfg = pkg.Config(embedding_size=16, fair_objective="value", fair_weight=0.7, device=cpu)
    m = pkg.FOCF(cfg, synth.SynthDataset(50, 40, 5.0))
    eng = kernels.FocfEngine(50, 40, 16, 64, cpu)
    monkeypatch.setattr(m, "_engine", lambda: eng)
    m.init_adam(lr=2e-3, weight_decay=1e-3)
    rng = np.random.default_rng(0)

    def packed(n):
        buf = torch.from_numpy(rng.integers(0, 255, 16 * n).astype(np.uint8))
        it = Interaction({"user_id": buf[:4 * n].view(torch.int32), "item_id": buf[4 * n:8 * n].view(torch.int32),
                          "rating": buf[8 * n:12 * n].view(torch.float32), "gender": buf[12 * n:].view(torch.float32)})
        it.items_contiguous, it.packed_host = True, (buf, n)
        return it, buf

    loss = torch.zeros(1)
    for n in (37, 37, 52, 300, 41):                     # 300 grows the workspace and the staging buffer
        it, buf = packed(n)
        monkeypatch.setattr(focf_mod, "_NO_FAST_HOST_STEP", True)
        m.train_step(it, loss_out=loss)
        m._adam["step"] -= 1
        monkeypatch.setattr(focf_mod, "_NO_FAST_HOST_STEP", False)
        m.train_step(it, loss_out=loss)
        a, b = seen[-2], seen[-1]
        batch_ptrs = ("uid", "iid", "rating", "sst")
        for k in a:
            if k not in batch_ptrs:
                assert a[k] == b[k], (n, k, a[k], b[k])
        base = eng._stage.data_ptr()
        assert [b[k] - base for k in batch_ptrs] == [0, 4 * n, 8 * n, 12 * n]
        assert [a[k] - a["uid"] for k in batch_ptrs] == [0, 4 * n, 8 * n, 12 * n]
        assert torch.equal(eng._stage[:16 * n], buf) and b["B"] == n and b["step"] == m._adam["step"]
    # the persistent struct follows a change of the Adam state, of the hyper-parameters and of the loss slot
    m.init_adam(lr=5e-4, weight_decay=0.0)
    other = torch.zeros(1)
    it, _ = packed(20)
    m.train_step(it, loss_out=other)
    b = seen[-1]
    assert b["lr"] == 5e-4 and b["weight_decay"] == 0.0 and b["mU"] == m._adam["mU"].data_ptr() and b["step"] == 1
    assert b["loss"] == other.data_ptr() and b["objective"] == m._objective and abs(b["fair_weight"] - 0.7) < 1e-7


def _live_reference():
    import sys
    here = os.path.dirname(__file__)
    sys.path.insert(0, os.path.join(here, "..", "oracle", "ref_shim"))
    import shim
    if not shim.available():
        pytest.skip("reference tree not present")
    sys.path.insert(0, os.path.join(here, "..", "oracle"))
    argv, sys.argv = sys.argv, sys.argv[:1]
    try:
        import gen_golden as gg
    finally:
        sys.argv = argv
    return gg


def test_initial_weights_equal_the_live_references_for_a_seed():
    """build container only: module creation order and initialisers consume torch's RNG exactly like the reference files,
    so `init_seed(s)` + model construction gives the reference's initial weights -- every registered tensor and every
    dict-held filter / discriminator MLP of PFCN_MLP / PMF / BiasedMF / DMF (sm, cm) and FairGo_PMF (LBA, WAP), and NFCF's
    stage-2 construction (checkpoint load, gender-direction projection, frozen user table, re-initialised item table)"""
    import importlib
    import tempfile
    gg = _live_reference()
    import recbole_fairrec_b200 as pkg
    from recbole.data.interaction import Interaction as RefInter
    nu, ni, d = 80, 50, 16
    rng = np.random.default_rng(0)
    gender, age = rng.integers(0, 2, nu).astype(np.float32), rng.integers(0, 4, nu).astype(np.float32)
    coo = sp.coo_matrix((rng.integers(1, 6, 400).astype(np.float32), (rng.integers(1, nu, 400), rng.integers(1, ni, 400))),
                        shape=(nu, ni))

    def dataset(inter_cls):
        class DS:
            inter_feat = {"rating": torch.tensor([1.0, 5.0])}

            def num(self, f):
                return {"user_id": nu, "item_id": ni}[f]

            def inter_matrix(self, form="coo", value_field=None):
                return coo

            def get_user_feature(self):
                return inter_cls({"user_id": torch.arange(nu), "gender": torch.from_numpy(gender), "age": torch.from_numpy(age)})
        return DS()

    def flat(model, attrs):
        out = {f"base.{k}": v.detach().clone() for k, v in model.state_dict().items()}
        for a in attrs:
            for k, m in (getattr(model, a, None) or {}).items():
                out.update({f"{a}.{k}.{kk}": v.detach().clone() for kk, v in m.state_dict().items()})
        return out

    cases = [(n, dict(filter_mode=m, dis_dropout=0.0, dis_weight=10.0, dis_hidden_size_list=[32, 16], activation="leakyrelu",
                      **gg.PFCN_EXTRA[n]), ("filter_layer", "dis_layer_dict"))
             for n in ("PFCN_MLP", "PFCN_PMF", "PFCN_BiasedMF", "PFCN_DMF") for m in ("sm", "cm")]
    cases += [("FairGo_PMF", dict(n_layers=2, activation="leakyrelu", dis_hidden_size_list=[16, 8], filter_hidden_size_list=[32, 16],
                                  fair_weight=0.1, load_pretrain_weight=False, aggr_method=a, vs_weights=[4, 1]),
               ("filter_layer_dict", "dis_layer_dict")) for a in ("LBA", "WAP")]
    for name, extra, attrs in cases:
        ref_cls = getattr(importlib.import_module(f"recbole.model.fair_recommender.{name.lower()}"), name)
        torch.manual_seed(7)
        ref = ref_cls(gg.base_cfg(embedding_size=d, sst_attr_list=["gender", "age"], **extra), dataset(RefInter))
        torch.manual_seed(7)
        mine = getattr(pkg, name)(pkg.Config(embedding_size=d, sst_attr_list=["gender", "age"], device=torch.device("cpu"), **extra),
                                  dataset(pkg.Interaction))
        a, b = flat(ref, attrs), flat(mine, attrs)
        assert set(a) == set(b), (name, set(a) ^ set(b))
        for k in a:
            assert torch.equal(a[k].float(), b[k].float()), (name, extra.get("filter_mode"), k)
    # NFCF stage 2
    from recbole.model.fair_recommender.nfcf import NFCF as RefNFCF
    path = os.path.join(tempfile.mkdtemp(), "ncf.pth")
    torch.save({"state_dict": {"user_embedding.weight": torch.randn(nu, d), "item_embedding.weight": torch.randn(ni, d)}}, path)
    kw = dict(embedding_size=d, mlp_hidden_size=[32, 16], dropout=0.0, fair_weight=0.1, load_pretrain_path=path)
    torch.manual_seed(3)
    ref = RefNFCF(gg.base_cfg(**kw), dataset(RefInter))
    torch.manual_seed(3)
    mine = pkg.NFCF(pkg.Config(sst_attr_list=["gender"], device=torch.device("cpu"), **kw), dataset(pkg.Interaction))
    assert torch.equal(ref.user_embedding.weight.data, mine.user_embedding.weight.data)
    assert torch.equal(ref.item_embedding.weight.data, mine.item_embedding.weight.data)
    assert not mine.user_embedding.weight.requires_grad and mine.item_embedding.weight.requires_grad
    for r, m in zip([x for x in ref.mlp_layers.mlp_layers if isinstance(x, torch.nn.Linear)], mine.mlp_layers.linears()):
        assert torch.equal(r.weight.data, m.weight.data) and torch.equal(r.bias.data, m.bias.data)


def test_fairgo_pretrain_keeps_the_best_validated_state(monkeypatch):
    """trainer.py:606-685: validation after every pretrain epoch, early stopping, and the BEST pretrained tables go into the
    fine-tune stage"""
    cfg, model, trainer = _family("FairGo_GCN", 5)
    cfg["stopping_step"], cfg["valid_metric"] = 1, "NDCG@5"
    trainer.pretrain_epochs = 10
    scores = iter([0.1, 0.4, 0.3, 0.2, 0.9])
    calls = []

    def fake_pass(self, data, loss_func, optimizer, sst_list):
        with torch.no_grad():
            self.model.user_embedding_layer.weight.add_(1.0)     # "training" moves the table by one per epoch
        calls.append("train")
        return 1.0

    monkeypatch.setattr(type(trainer), "_pass", fake_pass)
    monkeypatch.setattr(type(trainer), "evaluate", lambda self, data: {"ndcg@5": next(scores)})
    u0 = model.user_embedding_layer.weight.detach().clone()
    losses = trainer.pretrain([None], valid_data=[None])
    # epochs 1..4 ran (0.4 stands for two evaluations with stopping_step 1); the state of epoch 2 is restored
    assert len(losses) == 4 and model.train_stage == "finetune" and trainer.pretrain_valid_score == 0.4
    torch.testing.assert_close(model.user_embedding_layer.weight.detach(), u0 + 2.0)
    # without validation data: the plain loop, last state
    cfg2, model2, trainer2 = _family("FairGo_GCN", 6)
    monkeypatch.setattr(type(trainer2), "_pass", fake_pass)
    u0 = model2.user_embedding_layer.weight.detach().clone()
    trainer2.pretrain([None], epochs=3)
    torch.testing.assert_close(model2.user_embedding_layer.weight.detach(), u0 + 3.0)


def test_pfcn_validation_scores_every_attribute_subset_on_one_draw_of_negatives(monkeypatch):
    """trainer.py:985-1030: the candidate lists are tiled once per non-empty attribute subset (same negatives), tile s is
    scored with subset s, and ONE pass runs over the whole; trainer.py:1072-1086: the test evaluation is one pass per subset"""
    import recbole_fairrec_b200 as pkg
    from recbole_fairrec_b200.sampled_eval import ResamplingEvalSource, SampledEvaluator
    cfg, model, trainer = _family("PFCN_MLP", 8)                         # attributes gender, age
    assert trainer.attribute_subsets() == [["gender"], ["age"], ["gender", "age"]]
    rng = np.random.default_rng(0)
    users = np.array([3, 5, 9, 11])
    pos = [np.array([1, 2]), np.array([4]), np.array([7, 8, 9]), np.array([2])]
    used = [np.array([5, 6]), np.array([1, 2, 3]), np.array([10]), np.array([20, 21])]
    sst = {"gender": rng.integers(0, 2, 40).astype(np.float32), "age": rng.integers(0, 3, 40).astype(np.float32)}
    src = ResamplingEvalSource(users, pos, used, sst, 30, 4, torch.device("cpu"))
    np.random.seed(1)
    data, per_copy = src.resample_tiled(3)
    assert data.n == 12 and per_copy == sum(len(p) * 5 for p in pos) and data.cand_uid.numel() == 3 * per_copy
    assert torch.equal(data.cand_items[:per_copy], data.cand_items[per_copy:2 * per_copy])          # same negatives per tile
    assert torch.equal(data.users, torch.tensor(list(users) * 3, dtype=torch.int32))
    calls, captured = [], {}

    def fake_predict(inter, sst_list=None):
        calls.append((list(sst_list), int(inter["user_id"].numel())))
        return torch.full((inter["user_id"].numel(),), float(len(calls)))

    monkeypatch.setattr(model, "predict", fake_predict)
    monkeypatch.setattr(SampledEvaluator, "collect_scores", lambda self, scores, d: captured.update(scores=scores, data=d) or {})
    monkeypatch.setattr(SampledEvaluator, "finalize", lambda self, out, d, rounded=True: {"ndcg@5": 0.5})
    cfg["metrics"] = ["NDCG"]
    np.random.seed(1)
    res = trainer.pfcn_evaluate(src)
    assert res == {"ndcg@5": 0.5} and calls == [(["gender"], per_copy), (["age"], per_copy), (["gender", "age"], per_copy)]
    sc = captured["scores"]
    assert sc.numel() == 3 * per_copy and torch.equal(sc, torch.repeat_interleave(torch.tensor([1.0, 2.0, 3.0]), per_copy))
    assert torch.equal(captured["data"].cand_items, data.cand_items)                              # same RNG state, same draw
    # test evaluation: one pass (and one draw) per subset, keyed like the reference
    monkeypatch.setattr(type(trainer), "evaluate", lambda self, d, sst_list=None, c=None: {"subset": list(sst_list)})
    out = trainer.evaluate_subsets(src)
    assert list(out) == ["sm-['gender']", "sm-['age']", "sm-['gender', 'age']"] and out["sm-['age']"] == {"subset": ["age"]}


def test_loader_streams_equal_the_live_references(tmp_path):
    """build container only (oracle/fuzz_loaders.py): pairwise and pointwise train batches over three passes (in-place
    cumulative shuffles, negatives redrawn on collision) and the per-evaluation `uni<N>` / `pop<N>` negatives, with
    `neg_sampling` uniform and popularity with one and several negatives per positive, identical to the reference's TrainDataLoader / NegSampleEvalDataLoader for a seed"""
    _live_reference()
    import subprocess
    import sys
    here = os.path.dirname(__file__)
    r = subprocess.run([sys.executable, os.path.join(here, "..", "oracle", "fuzz_loaders.py"), "7", "3"], capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0 and "bad: 0" in r.stdout and r.stdout.count("IDENTICAL") == 6, r.stdout[-2000:] + r.stderr[-2000:]


def test_plugin_discovery_by_name_like_the_reference():
    """utils/utils.py:51-94: get_model finds the class called like the model, get_trainer `<name>Trainer` and otherwise the
    base trainer; an unknown model name raises the reference's ValueError"""
    import pytest
    import recbole_fairrec_b200 as pkg
    for name, trainer in (("FOCF", pkg.FOCFTrainer), ("PFCN_MLP", pkg.PFCNTrainer), ("PFCN_PMF", pkg.PFCNTrainer),
                          ("PFCN_BiasedMF", pkg.PFCNTrainer), ("PFCN_DMF", pkg.PFCNTrainer), ("FairGo_PMF", pkg.FairGoTrainer),
                          ("FairGo_GCN", pkg.FairGoTrainer), ("NFCF", pkg.NFCFTrainer)):
        cls = pkg.get_model(name)
        assert cls is getattr(pkg, name) and cls.__name__ == name and cls.type == "GENERAL"
        assert pkg.get_trainer(cls.type, name) is trainer
    assert pkg.get_trainer("GENERAL", "SomethingElse") is pkg.FOCFTrainer          # the base-trainer fall-through
    for bad in ("BPR", "FOCFTrainer", "focf"):
        with pytest.raises(ValueError, match="is not the name of an existing model"):
            pkg.get_model(bad)


def test_run_recbole_wiring_of_every_family_with_faked_epochs(tmp_path, monkeypatch):
    """run_recbole end to end on the raw ml-100k files for every branch of quick_start.run_recbole -- config, ingestion,
    model and trainer discovery, model construction, loaders (a few batches are drawn), per-evaluation negatives, the fit
    loops with check-pointing, best-model reload and the test evaluation -- with the kernel-calling methods (`_train_epoch`,
    `_pass`, `evaluate`) replaced by fakes, so the host logic around the CUDA calls runs without a GPU.  The same
    configurations run for real in tests/test_run_recbole_gpu.py."""
    from oracle import make_test_data as mtd
    import recbole_fairrec_b200 as pkg
    from recbole_fairrec_b200.quick_start import run_recbole
    base = dict(RATING_FIELD="rating", LABEL_FIELD="label", threshold={"rating": 3.0}, sst_attr_list=["gender"],
                load_col={"inter": ["user_id", "item_id", "rating"], "user": ["user_id", "gender"], "item": ["item_id"]},
                embedding_size=64, seed=2020, verbose=False, device="cpu", epochs=2, topk=[5], valid_metric="NDCG@5",
                data_path=mtd.float_gender_copy(str(tmp_path / "data")), checkpoint_dir=str(tmp_path),
                eval_args={"split": {"RS": [8, 1, 1]}, "group_by": "user", "order": "RO", "mode": "full"})
    uni = dict(base["eval_args"], mode="uni100")
    seen = []

    def fake_train(self, data, epoch, *a, **k):
        if isinstance(self, pkg.FOCFTrainer):
            n = sum(len(data._draw_batch()) > 0 for _ in range(3))          # the host side of the FOCF loader
        else:
            n = sum(1 for _ in zip(range(3), data))
        seen.append(("train", epoch, n))
        return 0.3 if isinstance(self, (pkg.FOCFTrainer, pkg.NFCFTrainer)) else (0.1, 0.2)

    def fake_pass(self, data, *a, **k):
        return 0.5 * sum(1 for _ in zip(range(2), data))

    def fake_eval(self, eval_data, *a, **k):
        data = eval_data.resample() if hasattr(eval_data, "resample") else eval_data
        seen.append(("eval", type(data).__name__, len(data.users) if hasattr(data, "users") else None))
        return {"ndcg@5": 0.1 + 0.01 * len(seen)}

    for cls in (pkg.FOCFTrainer, pkg.PFCNTrainer, pkg.FairGoTrainer, pkg.NFCFTrainer):
        monkeypatch.setattr(cls, "_train_epoch", fake_train)
        monkeypatch.setattr(cls, "evaluate", fake_eval)
    monkeypatch.setattr(pkg.PFCNTrainer, "_pass", fake_pass)
    monkeypatch.setattr(pkg.FairGoTrainer, "_pass", fake_pass)
    dis = dict(dis_dropout=0.0, dis_weight=1.0, dis_hidden_size_list=[32, 16], activation="leakyrelu")
    fairgo = dict(n_layers=2, activation="leakyrelu", dis_hidden_size_list=[16, 8, 4], filter_hidden_size_list=[128, 64],
                  fair_weight=0.1, load_pretrain_weight=False, aggr_method="LBA", vs_weights=[4, 1], pretrain_epochs=3)
    runs = [("FOCF", dict(fair_objective="value"), False, 3),
            ("FOCF", dict(fair_objective="value", eval_args=uni, focf_draw_mode="fast"), False, 3),
            ("PFCN_PMF", dict(dis, filter_mode="sm", eval_args=uni), False, 3),
            ("PFCN_MLP", dict(dis, filter_mode="cm", mlp_hidden_size=[32, 16], dropout=0.0, eval_args=uni), True, 3),
            ("PFCN_BiasedMF", dict(dis, filter_mode="sm", neg_sampling={"popularity": 1},          # popularity-biased negatives
                                   eval_args=dict(uni, mode="pop50")), False, 3),                  # in training and evaluation
            ("FairGo_PMF", fairgo, False, 6),          # + one validation per pretrain epoch
            ("FairGo_GCN", dict(fairgo, gcn_n_layers=2, hidden_channels=32, gcn_dropout=0.2, gcn_act="relu", eval_args=uni), True, 6),
            ("NFCF", dict(dropout=0.2, fair_weight=0.1, mlp_hidden_size=[128, 64], eval_args=uni, load_pretrain_path=None), True, 3)]
    for name, extra, saved, n_eval in runs:
        seen.clear()
        out = run_recbole(name, "ml-100k", None, dict(base, **extra), saved=saved)
        evals = [s for s in seen if s[0] == "eval"]
        assert len(evals) == n_eval, (name, seen)
        assert all(s[2] > 0 for s in seen if s[0] == "train"), (name, seen)
        assert out["best_valid_score"] == max(0.1 + 0.01 * (k + 1) for k, s in enumerate(seen[:-1]) if s[0] == "eval"), name
        assert set(out) >= {"best_valid_score", "valid_score_bigger", "best_valid_result", "test_result"}
        if saved:
            assert os.path.exists(out["saved_model_file"]), name
        if name == "NFCF":          # stage 2 reads the stage-1 checkpoint (nfcf.py:49-51)
            out2 = run_recbole(name, "ml-100k", None, dict(dict(base, **extra), load_pretrain_path=out["saved_model_file"]))
            assert out2["test_result"]["ndcg@5"] > 0


def test_host_batch_step_loop_passes_the_generic_steps_arguments(monkeypatch):
    """FOCF.train_steps_host -> fr_focf_train_steps_host with the library call intercepted: the step template carries the
    fields the generic per-batch step passes (tables, moments, hyper-parameters, workspace, objective), the optimizer step
    of the first batch is adam.step + 1 and the count advances by the number of batches, and the batch table holds the
    host buffers and their row counts.  (CPU tensors stand in for device memory: the launch is stubbed.)"""
    import recbole_fairrec_b200 as pkg
    from recbole_fairrec_b200 import _lib, focf as focf_mod, kernels, synth
    from recbole_fairrec_b200.interaction import Interaction
    real = _lib.load()
    steps, loops = [], []

    class FakeLib:
        fr_focf_workspace_bytes = staticmethod(real.fr_focf_workspace_bytes)

        def fr_focf_workspace_init(self, *a):
            return 0

        def fr_focf_train_step(self, ref, stream):
            steps.append({name: getattr(ref._obj, name) for name, _ in _lib.FocfStep._fields_})
            return 0

        def fr_focf_train_steps_host(self, ref, k, ptrs, nrow, stage, stage_bytes, loss_dev, loss_host, stream):
            loops.append(dict({name: getattr(ref._obj, name) for name, _ in _lib.FocfStep._fields_}, k=k,
                              ptrs=[ptrs[j] for j in range(k)], rows=[nrow[j] for j in range(k)], stage=stage,
                              stage_bytes=stage_bytes, loss_dev=loss_dev, loss_host=loss_host))
            return 0

    fake = FakeLib()
    monkeypatch.setattr(kernels, "load", lambda: fake)
    monkeypatch.setattr(kernels, "stream_ptr", lambda: 0)
    monkeypatch.setattr(kernels, "ptr", lambda t: None if t is None else t.data_ptr())
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    cpu = torch.device("cpu")
    cfg = pkg.Config(embedding_size=16, fair_objective="absolute", fair_weight=0.4, device=cpu)
    m = pkg.FOCF(cfg, synth.SynthDataset(50, 40, 5.0))
    eng = kernels.FocfEngine(50, 40, 16, 64, cpu)
    monkeypatch.setattr(m, "_engine", lambda: eng)
    m.init_adam(lr=2e-3, weight_decay=1e-3)
    rng = np.random.default_rng(0)
    batches = []
    for n in (37, 120, 52):
        buf = torch.from_numpy(rng.integers(0, 255, 16 * n).astype(np.uint8))
        it = Interaction({"user_id": buf[:4 * n].view(torch.int32), "item_id": buf[4 * n:8 * n].view(torch.int32),
                          "rating": buf[8 * n:12 * n].view(torch.float32), "gender": buf[12 * n:].view(torch.float32)})
        it.items_contiguous, it.packed_host = True, (buf, n)
        batches.append(it)
    m._adam["step"] = 5
    losses = m.train_steps_host(batches)
    assert m._adam["step"] == 8 and losses.shape == (3,) and losses.dtype == torch.float32
    t = loops[-1]
    assert (t["k"], t["rows"], t["step"]) == (3, [37, 120, 52], 6)
    assert t["ptrs"] == [b.packed_host[0].data_ptr() for b in batches]
    assert t["stage"] == eng._stage.data_ptr() and t["stage_bytes"] == eng._stage.numel() >= 16 * 120
    assert t["loss_host"] == losses.data_ptr() and t["loss_dev"] == eng._loss2.data_ptr()
    assert not t["plan_desc"] and not t["B_dev"]
    # same template as the generic step on one of the batches (which runs after the loop grew the workspace to 120 rows)
    monkeypatch.setattr(focf_mod, "_NO_FAST_HOST_STEP", True)
    m.train_step(batches[1], loss_out=torch.zeros(1))
    g = steps[-1]
    for k in g:
        if k not in ("uid", "iid", "rating", "sst", "B", "loss", "step", "scalars_filled"):   # (the last: a POINTER wrapper)
            assert g[k] == t[k], (k, g[k], t[k])
    assert g["step"] == 9
    with pytest.raises(ValueError):
        m.train_steps_host([Interaction({"user_id": torch.zeros(3, dtype=torch.int32)})])
    assert m.train_steps_host([]).numel() == 0 and m._adam["step"] == 9


def test_host_batch_step_loop_validates_its_arguments_before_touching_the_device():
    """fr_focf_train_steps_host refuses bad arguments with FR_ERR_INVALID and a message -- checked before any CUDA call, so
    this runs without a GPU"""
    from recbole_fairrec_b200 import _lib
    lib = _lib.load()
    s = _lib.FocfStep()
    s.step = 1
    buf = (ctypes.c_uint8 * 64)()
    ptrs = (ctypes.c_void_p * 2)(ctypes.addressof(buf), ctypes.addressof(buf))
    rows = (ctypes.c_int32 * 2)(4, 4)
    stage, loss = ctypes.addressof(buf), ctypes.addressof(buf)
    assert lib.fr_focf_train_steps_host(None, 1, ptrs, rows, stage, 64, loss, loss, None) == _lib.FR_ERR_INVALID
    assert lib.fr_focf_train_steps_host(ctypes.byref(s), -1, ptrs, rows, stage, 64, loss, loss, None) == _lib.FR_ERR_INVALID
    assert lib.fr_focf_train_steps_host(ctypes.byref(s), 0, None, None, None, 0, None, None, None) == _lib.FR_OK
    assert lib.fr_focf_train_steps_host(ctypes.byref(s), 2, ptrs, rows, None, 64, loss, loss, None) == _lib.FR_ERR_INVALID
    rows[1] = 5                                             # 80 bytes > the 64-byte staging buffer
    assert lib.fr_focf_train_steps_host(ctypes.byref(s), 2, ptrs, rows, stage, 64, loss, loss, None) == _lib.FR_ERR_INVALID
    assert b"exceeds the staging buffer" in lib.fr_last_error()
    rows[1] = 0
    assert lib.fr_focf_train_steps_host(ctypes.byref(s), 2, ptrs, rows, stage, 64, loss, loss, None) == _lib.FR_ERR_INVALID
    rows[1] = 4
    s.step = 0                                              # device-resident step counts belong to the planned epoch
    assert lib.fr_focf_train_steps_host(ctypes.byref(s), 2, ptrs, rows, stage, 64, loss, loss, None) == _lib.FR_ERR_INVALID


def test_epoch_kernel_entry_validates_its_slots_before_touching_the_device():
    """fr_focf_epoch_run / fr_focf_epoch_eligible refuse malformed slot arrays with FR_ERR_INVALID and a message -- checked
    before any CUDA call, so this runs without a GPU"""
    from recbole_fairrec_b200 import _lib
    lib = _lib.load()
    buf = (ctypes.c_uint8 * 256)()
    p = ctypes.addressof(buf)

    def slot(k, **kw):
        s = _lib.FocfStep()
        for f in ("U", "I", "mU", "vU", "mI", "vI", "loss", "status_flags", "plan_desc", "plan_items", "plan_offs", "item_off",
                  "train_uid", "train_rating", "sst_of_user"):
            setattr(s, f, p)
        for f in ("uid", "iid", "rating", "sst", "pred", "workspace"):          # per-slot buffers must differ
            setattr(s, f, p + 16 * (k + 1))
        s.n_users, s.n_items, s.d, s.B, s.plan_len, s.objective, s.workspace_bytes = 10, 10, 32, 64, 3, 1, 1 << 20
        for key, v in kw.items():
            setattr(s, key, v)
        return s

    def arr(*slots):
        a = (_lib.FocfStep * len(slots))()
        for k, s in enumerate(slots):
            ctypes.memmove(ctypes.byref(a[k]), ctypes.byref(s), ctypes.sizeof(s))
        return ctypes.cast(a, ctypes.c_void_p), a

    run = lambda a, n, first=0, steps=2, t=1, sync=p: lib.fr_focf_epoch_run(a, n, first, steps, t, sync, None)
    good, keep = arr(slot(0), slot(1), slot(2))
    assert run(good, 3, steps=0) == _lib.FR_OK                                   # nothing to do
    assert run(None, 3) == _lib.FR_ERR_INVALID
    assert run(good, 1) == _lib.FR_ERR_INVALID and b"slots" in lib.fr_last_error()
    assert run(good, 9) == _lib.FR_ERR_INVALID
    assert run(good, 3, t=0) == _lib.FR_ERR_INVALID                              # optimizer steps are 1-based
    assert run(good, 3, sync=None) == _lib.FR_ERR_INVALID
    bad, keep2 = arr(slot(0), slot(1, d=64))                                     # a slot with another shape
    assert run(bad, 2) == _lib.FR_ERR_INVALID and b"differs from slot 0" in lib.fr_last_error()
    bad, keep3 = arr(slot(0), slot(0))                                           # two slots on the same buffers
    assert run(bad, 2) == _lib.FR_ERR_INVALID and b"share" in lib.fr_last_error()
    bad, keep4 = arr(slot(0, plan_desc=None), slot(1, plan_desc=None))           # not a planned epoch
    assert run(bad, 2) == _lib.FR_ERR_INVALID and b"planned epoch" in lib.fr_last_error()
    bad, keep5 = arr(slot(0, B=9000), slot(1, B=9000))                           # batches beyond the producers' sort capacity
    assert run(bad, 2) == _lib.FR_ERR_INVALID and b"batch capacity" in lib.fr_last_error()
    bad, keep6 = arr(slot(0, adam_mode=1), slot(1, adam_mode=1))                 # lazy_exact belongs to the stepwise path
    assert run(bad, 2) == _lib.FR_ERR_INVALID
    assert lib.fr_focf_epoch_eligible(bad, 2) == 0


def test_alias_sampler_follows_item_popularity_and_the_reference_call_pattern():
    """sampler.py:72-120: keys in order of first appearance, a valid alias table (every column sums to 1 with its alias),
    draws proportional to the interaction counts, and exactly one randint + one random call per `sampling(n)`"""
    from recbole_fairrec_b200.sampled_eval import AliasSampler
    rng = np.random.default_rng(3)
    items = rng.choice(np.arange(1, 40), 5000, p=np.arange(1, 40) / np.arange(1, 40).sum())
    a = AliasSampler(items)
    _, first = np.unique(items, return_index=True)
    assert a.keys.tolist() == items[np.sort(first)].tolist()
    # mass conservation of the alias construction: P(key) = (prob[key] + sum of (1 - prob[j]) over j aliased to key) / n_keys
    n = len(a.keys)
    mass = a.prob.clip(max=1.0).copy()
    for j in range(n):
        if a.prob[j] < 1 and a.alias[j] >= 0:
            mass[np.flatnonzero(a.keys == a.alias[j])[0]] += 1 - a.prob[j]
    counts = np.array([np.sum(items == k) for k in a.keys])
    np.testing.assert_allclose(mass / n, counts / len(items), atol=1e-12)
    np.random.seed(5)
    draw = a.sampling(200000)
    freq = np.array([np.mean(draw == k) for k in a.keys])
    assert np.abs(freq - counts / len(items)).max() < 4e-3 and set(draw.tolist()) <= set(items.tolist())
    np.random.seed(5)
    idx, p = np.random.randint(0, n, 17), np.random.random(17)
    np.random.seed(5)
    assert a.sampling(17).tolist() == np.where(a.prob[idx] > p, a.keys[idx], a.alias[idx]).tolist()
