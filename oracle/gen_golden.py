#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED
reference (/root/reference, imported through oracle/ref_shim/shim.py).

TEST INFRASTRUCTURE ONLY (never imported by the product package).  Run in the build
container (the GPU box has no /root/reference; it only reads the committed .npz):

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz

What is pinned, and by which reference code:
  focf_train_*.npz   recbole/model/fair_recommender/focf.py:75-169 (+autograd) and
                     torch.optim.Adam as built by recbole/trainer/trainer.py:139
  focf_eval_*.npz    focf.py:171-178, trainer.py:420-439 (_full_sort_batch_eval),
                     evaluator/collector.py:131-205, evaluator/metrics.py (12 metrics)
  ml100k_*.npz       the whole pipeline on the bundled ml-100k (quick_start.py:20-71):
                     split tensors, the item draws of FOCFDataLoader, per-epoch loss and
                     valid/test metric dicts.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "ref_shim"))
import shim  # noqa: E402

shim.install()
import torch  # noqa: E402
from recbole.data.interaction import Interaction  # noqa: E402
from recbole.evaluator import Collector, Evaluator  # noqa: E402
from recbole.model.fair_recommender.focf import FOCF  # noqa: E402
from recbole.trainer import Trainer  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
METRICS12 = ["NDCG", "Recall", "Hit", "MRR", "DifferentialFairness", "GiniIndex", "PopularityPercentage",
             "ValueUnfairness", "AbsoluteUnfairness", "UnderUnfairness", "OverUnfairness", "NonParityUnfairness"]


class Cfg(dict):
    """dict that answers None for missing keys, like recbole Config.__getitem__ (configurator.py:405-409)."""

    def __getitem__(self, k):
        return self.get(k, None)


class FakeDataset:
    def __init__(self, n_users, n_items, max_rating):
        self._n = {"user_id": n_users, "item_id": n_items}
        self.inter_feat = {"rating": torch.tensor([1.0, float(max_rating)])}

    def num(self, f):
        return self._n[f]


def base_cfg(**kw):
    c = Cfg(USER_ID_FIELD="user_id", ITEM_ID_FIELD="item_id", NEG_PREFIX="neg_", RATING_FIELD="rating",
            LABEL_FIELD="label", device=torch.device("cpu"), embedding_size=16, sst_attr_list=["gender"],
            fair_weight=1.0, fair_objective="value", metric_decimal_place=12, topk=[10], metrics=METRICS12,
            eval_args={"mode": "full"}, popularity_ratio=0.1)
    c.update(kw)
    return c


def make_focf_batches(rng, n_users, n_items, n_steps, target_rows, single_group_step=None, shuffle=False,
                      float_sst=False):
    """Batches shaped like FOCFDataLoader's (focf_dataloader.py:37-50): whole items, item-contiguous,
    unique (user,item) pairs."""
    gender_of_user = rng.integers(1, 3, size=n_users)  # token ids 1/2 (row 0 = [PAD] unused)
    batches = []
    for s in range(n_steps):
        items = rng.permutation(np.arange(1, n_items))
        uid, iid = [], []
        for it in items:
            cnt = int(rng.integers(1, max(2, n_users // 3)))
            us = rng.choice(np.arange(1, n_users), size=cnt, replace=False)
            uid.append(us)
            iid.append(np.full(cnt, it))
            if sum(map(len, uid)) >= target_rows:
                break
        uid = np.concatenate(uid).astype(np.int64)
        iid = np.concatenate(iid).astype(np.int64)
        rating = rng.integers(1, 6, size=len(uid)).astype(np.float32)
        g = gender_of_user[uid].astype(np.int64)
        if single_group_step == s:
            g[:] = 2
        if shuffle:
            p = rng.permutation(len(uid))
            uid, iid, rating, g = uid[p], iid[p], rating[p], g[p]
        if float_sst:
            g = (g - 1).astype(np.float32)
        batches.append((uid, iid, rating, g))
    return batches


def run_focf_train(name, objective, n_users=50, n_items=40, d=16, n_steps=3, target_rows=260, seed=0, scale=0.6,
                   fair_weight=1.0, lr=1e-3, wd=1e-3, **bk):
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    cfg = base_cfg(embedding_size=d, fair_objective=objective, fair_weight=fair_weight)
    model = FOCF(cfg, FakeDataset(n_users, n_items, 5.0))
    with torch.no_grad():
        model.user_embedding_layer.weight.copy_(torch.from_numpy(
            (rng.standard_normal((n_users, d)) * scale).astype(np.float32)))
        model.item_embedding_layer.weight.copy_(torch.from_numpy(
            (rng.standard_normal((n_items, d)) * scale).astype(np.float32)))
    U0 = model.user_embedding_layer.weight.detach().numpy().copy()
    I0 = model.item_embedding_layer.weight.detach().numpy().copy()
    opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=wd)  # trainer.py:139
    batches = make_focf_batches(rng, n_users, n_items, n_steps, target_rows, **bk)
    out = dict(U0=U0, I0=I0, n_steps=n_steps, objective=objective, fair_weight=fair_weight, lr=lr, wd=wd,
               max_rating=5.0)
    losses = []
    for s, (uid, iid, rating, g) in enumerate(batches):
        inter = Interaction({"user_id": torch.from_numpy(uid), "item_id": torch.from_numpy(iid),
                             "rating": torch.from_numpy(rating), "gender": torch.from_numpy(g)})
        opt.zero_grad()
        loss = model.calculate_loss(inter)  # focf.py:152-169
        loss.backward()
        if s == 0:
            pred, _, _ = model.forward(inter["user_id"], inter["item_id"])
            out["pred0"] = pred.detach().numpy().copy()
            out["dU0"] = model.user_embedding_layer.weight.grad.numpy().copy()
            out["dI0"] = model.item_embedding_layer.weight.grad.numpy().copy()
            out["predict0"] = model.predict(inter).detach().numpy().copy()  # focf.py:145-150
        opt.step()
        losses.append(loss.item())
        out[f"uid{s}"], out[f"iid{s}"], out[f"rating{s}"], out[f"sst{s}"] = uid, iid, rating, g
    out["losses"] = np.array(losses, dtype=np.float32)
    out["U_final"] = model.user_embedding_layer.weight.detach().numpy().copy()
    out["I_final"] = model.item_embedding_layer.weight.detach().numpy().copy()
    st = opt.state[model.user_embedding_layer.weight]
    out["mU_final"], out["vU_final"] = st["exp_avg"].numpy().copy(), st["exp_avg_sq"].numpy().copy()
    st = opt.state[model.item_embedding_layer.weight]
    out["mI_final"], out["vI_final"] = st["exp_avg"].numpy().copy(), st["exp_avg_sq"].numpy().copy()
    np.savez_compressed(os.path.join(OUT, f"focf_train_{name}.npz"), **out)
    print(f"focf_train_{name}: losses={losses}")


class FakeTrainer:
    """Just the attributes Trainer._full_sort_batch_eval (trainer.py:420-439) touches."""

    def __init__(self, model, n_items):
        self.model, self.device, self.tot_item_num = model, torch.device("cpu"), n_items


def run_focf_eval(name, n_users=60, n_items=97, d=16, K=10, topk=(5, 10), seed=1, scale=0.55, positive_only=False,
                  users_per_batch=7, float_sst=False, n_groups=2):
    """Full-sort eval of the reference: full_sort_predict -> mask -> Collector -> 12 metrics."""
    rng = np.random.default_rng(seed)
    cfg = base_cfg(embedding_size=d, topk=list(topk), metric_decimal_place=12)
    model = FOCF(cfg, FakeDataset(n_users, n_items, 5.0))
    U = (rng.standard_normal((n_users, d)) * scale).astype(np.float32)
    It = (rng.standard_normal((n_items, d)) * scale).astype(np.float32)
    if positive_only:  # tie-free variant: all dots strictly inside (0, max_rating)
        U, It = np.abs(U) * 0.5 + 0.01, np.abs(It) * 0.5 + 0.01
    with torch.no_grad():
        model.user_embedding_layer.weight.copy_(torch.from_numpy(U))
        model.item_embedding_layer.weight.copy_(torch.from_numpy(It))
    model.eval()
    sst_of_user = rng.integers(1, n_groups + 1, size=n_users)
    if float_sst:
        sst_of_user = (sst_of_user - 1).astype(np.float32)
    # eval users: all but pad and a few users without positives
    eval_users = np.array([u for u in range(1, n_users) if u % 9 != 0], dtype=np.int64)
    hist, pos = {}, {}
    for u in eval_users:
        n_used = int(rng.integers(3, 30))
        used = rng.choice(np.arange(1, n_items), size=n_used, replace=False)
        n_pos = int(rng.integers(1, min(6, n_used)))
        pos[u], hist[u] = used[:n_pos], used[n_pos:]
    train_item_count = {int(i): int(c) for i, c in zip(np.arange(1, n_items), rng.integers(1, 50, n_items - 1))}

    trainer = FakeTrainer(model, n_items)
    collector = Collector(cfg)
    collector.data_struct.set("data.num_items", n_items)  # collector.py:86-88
    from collections import Counter
    collector.data_struct.set("data.count_items", Counter(train_item_count))  # collector.py:91-93
    with torch.no_grad():
        for b0 in range(0, len(eval_users), users_per_batch):
            bu = eval_users[b0:b0 + users_per_batch]
            inter = Interaction({"user_id": torch.from_numpy(bu),
                                 "gender": torch.from_numpy(sst_of_user[bu])})
            hu = torch.cat([torch.full((len(hist[u]),), i, dtype=torch.int64) for i, u in enumerate(bu)])
            hi = torch.cat([torch.from_numpy(hist[u]) for u in bu])
            pu = torch.cat([torch.full((len(pos[u]),), i, dtype=torch.int64) for i, u in enumerate(bu)])
            pi = torch.cat([torch.from_numpy(pos[u]) for u in bu])
            inter, scores, pu, pi = Trainer._full_sort_batch_eval(trainer, (inter, (hu, hi), pu, pi))
            collector.eval_batch_collect(scores, inter, pu, pi)
    struct = collector.get_data_struct()
    result = Evaluator(cfg).evaluate(struct)
    out = dict(U=U, I=It, max_rating=5.0, K=K, topk=np.array(topk), eval_users=eval_users,
               sst_of_user=sst_of_user,
               hist_off=np.cumsum([0] + [len(hist[u]) for u in eval_users]),
               hist_items=np.concatenate([hist[u] for u in eval_users]),
               pos_off=np.cumsum([0] + [len(pos[u]) for u in eval_users]),
               pos_items=np.concatenate([pos[u] for u in eval_users]),
               train_count_items=np.array(sorted(train_item_count.items()), dtype=np.int64),
               rec_items=struct.get("rec.items").numpy(), rec_topk=struct.get("rec.topk").numpy(),
               rec_positive_score=struct.get("rec.positive_score").numpy(),
               data_positive_i=struct.get("data.positive_i").numpy(),
               data_sst=struct.get("data.gender").numpy(),
               metric_names=np.array(list(result.keys())),
               metric_values=np.array([float(v) for v in result.values()], dtype=np.float64))
    np.savez_compressed(os.path.join(OUT, f"focf_eval_{name}.npz"), **out)
    print(f"focf_eval_{name}:", {k: float(v) for k, v in result.items()})


def run_uni_eval(name, n_users=50, n_items=400, d=16, topk=(5, 10), seed=51, neg_num=100, users_per_batch=1,
                 transform="clamp", all_metrics=False):
    """Sampled-negative (uni100) ranking evaluation of the reference: NegSampleEvalDataLoader's batch layout
    (general_dataloader.py:128-152: per user [positives ; neg_num negatives per positive], row index per interaction),
    Trainer._neg_sample_batch_eval (trainer.py:441-456: predict + scatter into a -inf [users, n_items] matrix),
    Collector.eval_batch_collect and the mode-independent metrics.  Negatives are drawn like
    Sampler.sample_by_key_ids (uniform over the items not used by the user), recorded in the fixture."""
    rng = np.random.default_rng(seed)
    metrics = ["NDCG", "Recall", "Hit", "MRR", "DifferentialFairness", "GiniIndex", "PopularityPercentage",
               "NonParityUnfairness"]
    if all_metrics:   # + the four unfairness metrics in their sampled mode (well defined with one user per batch)
        metrics = METRICS12
    cfg = base_cfg(embedding_size=d, topk=list(topk), metric_decimal_place=12, metrics=metrics,
                   eval_args={"mode": f"uni{neg_num}"})
    model = FOCF(cfg, FakeDataset(n_users, n_items, 5.0))
    U = (np.abs(rng.standard_normal((n_users, d))) * 0.35 + 0.01).astype(np.float32)
    It = (np.abs(rng.standard_normal((n_items, d))) * 0.35 + 0.01).astype(np.float32)
    with torch.no_grad():
        model.user_embedding_layer.weight.copy_(torch.from_numpy(U))
        model.item_embedding_layer.weight.copy_(torch.from_numpy(It))
    model.eval()
    sst_of_user = rng.integers(1, 3, size=n_users)
    eval_users = np.array([u for u in range(1, n_users) if u % 7 != 0], dtype=np.int64)
    pos, neg = {}, {}
    for u in eval_users:
        n_used = int(rng.integers(4, 40))
        used = rng.choice(np.arange(1, n_items), size=n_used, replace=False)
        pos[u] = used[:int(rng.integers(1, min(6, n_used)))]
        free = np.setdiff1d(np.arange(1, n_items), used)
        neg[u] = rng.choice(free, size=neg_num * len(pos[u]), replace=True)     # [j * p + k] = j-th negative of positive k
    train_item_count = {int(i): int(c) for i, c in zip(np.arange(1, n_items), rng.integers(1, 50, n_items - 1))}
    trainer = FakeTrainer(model, n_items)
    trainer.test_batch_size = 1 << 30
    trainer.config = cfg
    from recbole.utils import EvaluatorType
    cfg["eval_type"] = EvaluatorType.RANKING
    collector = Collector(cfg)
    collector.data_struct.set("data.num_items", n_items)
    from collections import Counter
    collector.data_struct.set("data.count_items", Counter(train_item_count))
    with torch.no_grad():
        for b0 in range(0, len(eval_users), users_per_batch):
            bu = eval_users[b0:b0 + users_per_batch]
            uid = np.concatenate([np.full(len(pos[u]) * (neg_num + 1), u) for u in bu])
            iid = np.concatenate([np.concatenate([pos[u], neg[u]]) for u in bu])
            row = np.concatenate([np.full(len(pos[u]) * (neg_num + 1), i) for i, u in enumerate(bu)])
            inter = Interaction({"user_id": torch.from_numpy(uid), "item_id": torch.from_numpy(iid),
                                 "gender": torch.from_numpy(sst_of_user[uid])})
            pu = torch.cat([torch.full((len(pos[u]),), i, dtype=torch.int64) for i, u in enumerate(bu)])
            pi = torch.cat([torch.from_numpy(pos[u]) for u in bu])
            inter, scores, pu, pi = Trainer._neg_sample_batch_eval(trainer, (inter, torch.from_numpy(row), pu, pi))
            collector.eval_batch_collect(scores, inter, pu, pi)
    struct = collector.get_data_struct()
    result = Evaluator(cfg).evaluate(struct)
    out = dict(U=U, I=It, max_rating=5.0, topk=np.array(topk), eval_users=eval_users, sst_of_user=sst_of_user,
               users_per_batch=users_per_batch,
               neg_num=neg_num, pos_off=np.cumsum([0] + [len(pos[u]) for u in eval_users]),
               pos_items=np.concatenate([pos[u] for u in eval_users]),
               neg_items=np.concatenate([neg[u] for u in eval_users]),
               train_count_items=np.array(sorted(train_item_count.items()), dtype=np.int64),
               rec_items=struct.get("rec.items").numpy(), rec_topk=struct.get("rec.topk").numpy(),
               rec_positive_score=struct.get("rec.positive_score").numpy(),
               metric_names=np.array(list(result.keys())),
               metric_values=np.array([float(v) for v in result.values()], dtype=np.float64))
    np.savez_compressed(os.path.join(OUT, f"uni_eval_{name}.npz"), **out)
    print(f"uni_eval_{name}:", {k: float(v) for k, v in result.items()})


def run_ml100k(epochs=2):
    """End-to-end on the bundled ml-100k through run_recbole's own steps (quick_start.py:20-71); records
    the split tensors, the FOCFDataLoader item draws, the per-epoch train loss and the metric dicts."""
    import tempfile
    import yaml
    from recbole.config import Config
    from recbole.data import create_dataset, data_preparation
    from recbole.utils import init_seed, get_model, get_trainer
    import recbole.data.dataloader.focf_dataloader as fdl

    cfg_dict = dict(
        data_path=os.path.join(shim.REFERENCE_ROOT, "recbole/dataset_example/"), RATING_FIELD="rating",
        LABEL_FIELD="label", threshold={"rating": 3.0},
        load_col={"inter": ["user_id", "item_id", "rating"], "user": ["user_id", "gender"], "item": ["item_id"]},
        sst_attr_list=["gender"], fair_weight=1.0, neg_sampling=None, weight_decay=0.001,
        embedding_size=64, epochs=epochs, topk=[10], valid_metric="NDCG@10", metrics=METRICS12, seed=2020,
        use_gpu=False, state="WARNING", show_progress=False, metric_decimal_place=12,
        eval_args={"split": {"RS": [8, 1, 1]}, "group_by": "user", "order": "RO", "mode": "full"})
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.chdir(tmp)
    try:
        # fair_objective must come from a YAML file
        with open("c.yaml", "w") as f:
            yaml.safe_dump(dict(cfg_dict, fair_objective="value"), f)
        sys.argv = sys.argv[:1]
        config = Config(model="FOCF", dataset="ml-100k", config_file_list=["c.yaml"])
        init_seed(config["seed"], config["reproducibility"])
        dataset = create_dataset(config)
        train_data, valid_data, test_data = data_preparation(config, dataset)
        model = get_model("FOCF")(config, train_data.dataset).to(config["device"])
        U0 = model.user_embedding_layer.weight.detach().numpy().copy()
        I0 = model.item_embedding_layer.weight.detach().numpy().copy()
        trainer = get_trainer(config["MODEL_TYPE"], config["model"])(config, model)

        draws = []  # item ids drawn by FOCFDataLoader._next_batch_data, in order, -1 = batch end
        orig_next = fdl.FOCFDataLoader._next_batch_data
        orig_choice = np.random.choice

        def rec_choice(*a, **k):
            r = orig_choice(*a, **k)
            draws.append(int(r[0]))
            return r

        def rec_next(self):
            np.random.choice = rec_choice
            try:
                return orig_next(self)
            finally:
                np.random.choice = orig_choice
                draws.append(-1)

        fdl.FOCFDataLoader._next_batch_data = rec_next
        trainer.eval_collector.data_collect(train_data)
        losses, valids = [], []
        for ep in range(epochs):
            losses.append(trainer._train_epoch(train_data, ep))
            valids.append(trainer.evaluate(valid_data, load_best_model=False))
        test = trainer.evaluate(test_data, load_best_model=False)
        fdl.FOCFDataLoader._next_batch_data = orig_next

        def split(dl):
            f = dl.dataset.inter_feat
            return (f["user_id"].numpy().astype(np.int32), f["item_id"].numpy().astype(np.int32),
                    f["rating"].numpy().astype(np.float32))

        tr, va, te = split(train_data), split(valid_data), split(test_data)
        gender = dataset.get_user_feature()["gender"].numpy().copy()
        gender[0] = 0  # pad row (pandas-3 CoW leaves INT64_MIN there; never referenced)
        out = dict(n_users=dataset.user_num, n_items=dataset.item_num,
                   train_u=tr[0], train_i=tr[1], train_r=tr[2], valid_u=va[0], valid_i=va[1],
                   test_u=te[0], test_i=te[1], gender=gender.astype(np.int8),
                   draws=np.array(draws, dtype=np.int32), U0=U0, I0=I0,
                   epoch_losses=np.array(losses, dtype=np.float64),
                   metric_names=np.array(list(test.keys())),
                   valid_metrics=np.array([[float(v) for v in r.values()] for r in valids]),
                   test_metrics=np.array([float(v) for v in test.values()]),
                   U_final=model.user_embedding_layer.weight.detach().numpy().copy(),
                   I_final=model.item_embedding_layer.weight.detach().numpy().copy(),
                   max_rating=float(model.max_rating), train_batch_size=config["train_batch_size"])
        np.savez_compressed(os.path.join(OUT, "ml100k_focf_value.npz"), **out)
        print("ml100k: losses", losses, "test", {k: float(v) for k, v in test.items()})
    finally:
        Dataset._fill_nan = stock_fill_nan
        os.chdir(cwd)


def run_nfcf(name, fair, n_users=60, n_items=45, d=16, hidden=(32, 16), n_steps=3, B=96, seed=21, fair_weight=0.1, n_groups=2,
             lr=1e-3, wd=1e-6):
    """NFCF (recbole/model/fair_recommender/nfcf.py): NCF tower + BCE (+ differential-fairness regulariser when a
    pre-trained checkpoint was loaded: reset_params de-biases the user table and freezes it)."""
    import tempfile
    from recbole.model.fair_recommender.nfcf import NFCF

    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    gender = rng.integers(1, 1 + n_groups, size=n_users)   # n_groups > 2: nfcf.py:91-95 takes the max over all pairs

    class DS(FakeDataset):
        def get_user_feature(self):
            return Interaction({"user_id": torch.arange(n_users), "gender": torch.from_numpy(gender)})

    cfg = base_cfg(embedding_size=d, mlp_hidden_size=list(hidden), dropout=0.0, fair_weight=fair_weight,
                   load_pretrain_path=None, LABEL_FIELD="label")
    pre = NFCF(cfg, DS(n_users, n_items, 5.0))
    with torch.no_grad():   # make the scores spread out (default N(0,1) embeddings saturate ReLU towers less predictably)
        pre.user_embedding.weight.mul_(0.5)
        pre.item_embedding.weight.mul_(0.5)
    if fair:
        path = os.path.join(tempfile.mkdtemp(), "ncf.pth")
        torch.save({"state_dict": pre.state_dict()}, path)
        cfg = base_cfg(**{**cfg, "load_pretrain_path": path})
        model = NFCF(cfg, DS(n_users, n_items, 5.0))
        with torch.no_grad():
            model.item_embedding.weight.mul_(0.5)
    else:
        model = pre
    lin = [m for m in model.mlp_layers.mlp_layers if isinstance(m, torch.nn.Linear)]
    with torch.no_grad():   # the tower ends in ReLU (layers.py:66-68): keep most logits alive so gradients are non-trivial
        lin[-1].bias.fill_(0.3)
    out = dict(U0=model.user_embedding.weight.detach().numpy().copy(),
               I0=model.item_embedding.weight.detach().numpy().copy(), n_layers=len(lin), fair=int(fair),
               fair_weight=fair_weight, lr=lr, wd=wd, n_steps=n_steps,
               user_frozen=int(not model.user_embedding.weight.requires_grad))
    for k, l in enumerate(lin):
        out[f"W{k}_0"], out[f"b{k}_0"] = l.weight.detach().numpy().copy(), l.bias.detach().numpy().copy()
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=lr, weight_decay=wd)
    losses = []
    for s in range(n_steps):
        half = B // 2
        u = rng.integers(1, n_users, half)
        ipos = rng.integers(1, max(2, n_items // 3), half)   # few distinct positive items -> populated item x group cells
        ineg = rng.integers(1, n_items, half)
        uid = np.concatenate([u, u]).astype(np.int64)
        iid = np.concatenate([ipos, ineg]).astype(np.int64)
        label = np.concatenate([np.ones(half), np.zeros(half)]).astype(np.float32)
        g = gender[uid].astype(np.int64)
        inter = Interaction({"user_id": torch.from_numpy(uid), "item_id": torch.from_numpy(iid),
                             "label": torch.from_numpy(label), "gender": torch.from_numpy(g)})
        opt.zero_grad()
        loss = model.calculate_loss(inter)
        loss.backward()
        if s == 0:
            out["pred0"] = model.predict(inter).detach().numpy().copy()
            out["dI0"] = model.item_embedding.weight.grad.numpy().copy()
            if model.user_embedding.weight.grad is not None:
                out["dU0"] = model.user_embedding.weight.grad.numpy().copy()
            for k, l in enumerate(lin):
                out[f"dW{k}_0"], out[f"db{k}_0"] = l.weight.grad.numpy().copy(), l.bias.grad.numpy().copy()
        opt.step()
        losses.append(loss.item())
        out[f"uid{s}"], out[f"iid{s}"], out[f"label{s}"], out[f"sst{s}"] = uid, iid, label, g
    out["losses"] = np.array(losses, np.float32)
    out["U_final"] = model.user_embedding.weight.detach().numpy().copy()
    out["I_final"] = model.item_embedding.weight.detach().numpy().copy()
    for k, l in enumerate(lin):
        out[f"W{k}_final"], out[f"b{k}_final"] = l.weight.detach().numpy().copy(), l.bias.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, f"nfcf_train_{name}.npz"), **out)
    print(f"nfcf_train_{name}: losses={losses}")


def _dump_mlp(prefix, mlp, out, tag):
    for k, v in mlp.state_dict().items():
        out[f"{prefix}.{k}@{tag}"] = v.detach().numpy().copy()


PFCN_EXTRA = {
    "PFCN_MLP": dict(dropout=0.0, mlp_hidden_size_list=[16, 8]),
    "PFCN_PMF": dict(),
    "PFCN_BiasedMF": dict(),
    "PFCN_DMF": dict(num_layers=2, mlp_dropout=0.0, mlp_activation="leakyrelu", dis_activation="leakyrelu"),
}


def run_pfcn(model_name, name, filter_mode="sm", n_users=80, n_items=50, d=16, B=128, seed=31, n_rounds=2):
    """PFCN_* (recbole/model/fair_recommender/pfcn_*.py) driven like PFCNTrainer (trainer.py:875-898, 1189-1235):
    alternating filter+base steps on `bpr - dis_weight*dis` and discriminator steps on `dis`; dropout off (the
    reference's Philox masks cannot be reproduced), BatchNorm in training mode.  State keys: `base.<state_dict key>` for
    the registered modules, `filter_<idx>.…` / `dis_<attr>.…` for the dict-held MLPs."""
    import importlib
    cls = getattr(importlib.import_module(f"recbole.model.fair_recommender.{model_name.lower()}"), model_name)

    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    gender = rng.integers(0, 2, n_users).astype(np.float32)
    age = rng.integers(0, 4, n_users).astype(np.float32)

    class DS(FakeDataset):
        def get_user_feature(self):
            return Interaction({"user_id": torch.arange(n_users), "gender": torch.from_numpy(gender),
                                "age": torch.from_numpy(age)})

    cfg = base_cfg(embedding_size=d, sst_attr_list=["gender", "age"], filter_mode=filter_mode, dis_dropout=0.0,
                   dis_weight=10.0, dis_hidden_size_list=[32, 16], activation="leakyrelu", **PFCN_EXTRA[model_name])
    model = cls(cfg, DS(n_users, n_items, 5.0))
    with torch.no_grad():   # default inits (N(0,1) embeddings, N(0,0.01) MLPs) give vanishing signals: widen them
        for k_, p_ in model.named_parameters():
            if "embedding" in k_:
                p_.mul_(0.5)
            elif "bias" in k_ and p_.dim() == 2:                    # nn.Embedding(n, 1) bias tables
                p_.mul_(0.3)
        mlps = list(model.filter_layer.values()) + list(model.dis_layer_dict.values())
        mlps += [getattr(model, a_) for a_ in ("user_mlp", "item_mlp") if hasattr(model, a_)]
        for m in mlps:
            for p_ in m.parameters():
                if p_.dim() == 2:
                    p_.copy_(torch.randn_like(p_) * (1.0 / np.sqrt(p_.shape[1])))
        if hasattr(model, "mlp_layer"):
            for lin in [m for m in model.mlp_layer.mlp_layers if isinstance(m, torch.nn.Linear)]:
                lin.bias.add_(0.2)
    out = dict(model=model_name, filter_mode=filter_mode, d=d, gender=gender, age=age, n_users=n_users, n_items=n_items,
               n_rounds=n_rounds)

    def dump(tag):
        for k_, v_ in model.state_dict().items():
            out[f"base.{k_}@{tag}"] = v_.detach().numpy().copy()
        for k_, f in model.filter_layer.items():
            _dump_mlp(f"filter_{k_}", f, out, tag)
        for k_, f in model.dis_layer_dict.items():
            _dump_mlp(f"dis_{k_}", f, out, tag)

    dump("init")
    base = list(model.parameters())
    fparams = [p_ for f in model.filter_layer.values() for p_ in f.parameters()]
    dparams = [p_ for f in model.dis_layer_dict.values() for p_ in f.parameters()]
    opt_f = torch.optim.Adam(base + fparams, lr=1e-3, weight_decay=1e-4)
    opt_d = torch.optim.Adam(dparams, lr=1e-3, weight_decay=1e-4)
    model.train()
    for m in list(model.filter_layer.values()) + list(model.dis_layer_dict.values()):
        m.train()
    sst_lists = [["gender", "age"], ["age"], ["gender"]]
    losses = []
    for r in range(n_rounds):
        for phase, (fn, opt) in enumerate(((model.calculate_loss, opt_f), (model.calculate_dis_loss, opt_d))):
            s = 2 * r + phase
            u = rng.integers(1, n_users, B)
            inter = Interaction({"user_id": torch.from_numpy(u.astype(np.int64)),
                                 "item_id": torch.from_numpy(rng.integers(1, n_items, B).astype(np.int64)),
                                 "neg_item_id": torch.from_numpy(rng.integers(1, n_items, B).astype(np.int64)),
                                 "gender": torch.from_numpy(gender[u]), "age": torch.from_numpy(age[u])})
            sst_list = sst_lists[r % len(sst_lists)]
            opt.zero_grad()
            loss = fn(inter, sst_list)
            loss.backward()
            if s == 0:
                for k_, p_ in model.named_parameters():
                    if p_.grad is not None:
                        out[f"grad_base.{k_}@0"] = p_.grad.numpy().copy()
                idx = sum(model.sst_dict[a] for a in sst_list) if filter_mode == "sm" else model.sst_dict[sst_list[0]]
                for k_, p_ in model.filter_layer[idx].named_parameters():
                    out[f"grad_filter_{idx}.{k_}@0"] = p_.grad.numpy().copy()
                with torch.no_grad():
                    out["predict0"] = model.predict(inter, sst_list).numpy().copy()   # train-mode batch statistics
                    # NB: this extra train-mode forward also advances the BatchNorm running statistics; replays do the same
            opt.step()
            losses.append(loss.item())
            for k_ in ("user_id", "item_id", "neg_item_id"):
                out[f"{k_}{s}"] = inter[k_].numpy()
            out[f"sst_list{s}"] = np.array(sst_list)
    out["losses"] = np.array(losses, np.float32)
    dump("final")
    np.savez_compressed(os.path.join(OUT, f"pfcn_{name}.npz"), **out)
    print(f"pfcn_{name}: losses={losses}")


def run_all_pfcn():
    for old in ("pfcn_mlp_sm.npz", "pfcn_mlp_cm.npz"):
        if os.path.exists(os.path.join(OUT, old)):
            os.remove(os.path.join(OUT, old))
    run_pfcn("PFCN_MLP", "mlp_sm", "sm")
    run_pfcn("PFCN_MLP", "mlp_cm", "cm", seed=32)
    run_pfcn("PFCN_PMF", "pmf_sm", "sm", seed=33)
    run_pfcn("PFCN_PMF", "pmf_cm", "cm", seed=34)
    run_pfcn("PFCN_BiasedMF", "biasedmf_sm", "sm", seed=35)
    run_pfcn("PFCN_DMF", "dmf_cm", "cm", seed=36)


def run_fairgo(name, aggr, n_layers=2, n_users=70, n_items=45, d=16, B=96, n_inter=600, seed=41, n_rounds=2,
               pretrain_steps=2):
    """FairGo_PMF (recbole/model/fair_recommender/fairgo_pmf.py) driven like FairGo_PMFTrainer (trainer.py:534-704,
    837-848): `pretrain_steps` pretrain steps on the rating loss (optimizer over the two embedding tables), then the
    alternating fine-tune schedule (filter optimizer on `mse - fair_weight*dis`, discriminator (+aggr_layer for LBA)
    optimizer on `dis`).  The model's MLPs have no dropout / BatchNorm, so the run is deterministic."""
    import scipy.sparse as sp_
    from recbole.model.fair_recommender.fairgo_pmf import FairGo_PMF

    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    gender = rng.integers(0, 2, n_users).astype(np.float32)
    age = rng.integers(0, 4, n_users).astype(np.float32)
    pairs = rng.permutation((n_users - 1) * (n_items - 1))[:n_inter]
    tu, ti = (pairs // (n_items - 1) + 1).astype(np.int64), (pairs % (n_items - 1) + 1).astype(np.int64)
    tr = rng.integers(1, 6, n_inter).astype(np.float32)

    class DS(FakeDataset):
        def get_user_feature(self):
            return Interaction({"user_id": torch.arange(n_users), "gender": torch.from_numpy(gender),
                                "age": torch.from_numpy(age)})

        def inter_matrix(self, form="coo", value_field=None):
            return sp_.coo_matrix((tr, (tu, ti)), shape=(n_users, n_items))

    cfg = base_cfg(embedding_size=d, sst_attr_list=["gender", "age"], n_layers=n_layers, activation="leakyrelu",
                   dis_hidden_size_list=[16, 8], filter_hidden_size_list=[32, 16], fair_weight=0.5,
                   load_pretrain_weight=False, aggr_method=aggr, vs_weights=[4, 1] if n_layers == 2 else [3, 2, 1])
    model = FairGo_PMF(cfg, DS(n_users, n_items, 5.0))
    with torch.no_grad():
        model.user_embedding_layer.weight.mul_(0.6)
        model.item_embedding_layer.weight.mul_(0.6)
    out = dict(aggr=aggr, n_layers=n_layers, d=d, gender=gender, age=age, n_users=n_users, n_items=n_items,
               n_rounds=n_rounds, pretrain_steps=pretrain_steps, train_u=tu, train_i=ti, train_r=tr,
               vs_weights=np.array(cfg["vs_weights"], np.float32), fair_weight=0.5)
    L = model.norm_rating_matrix.coalesce()
    out["norm_row"], out["norm_col"] = L.indices()[0].numpy().copy(), L.indices()[1].numpy().copy()
    out["norm_val"] = L.values().numpy().copy()

    def dump(tag):
        for k_, v_ in model.state_dict().items():
            out[f"base.{k_}@{tag}"] = v_.detach().numpy().copy()
        for k_, f in model.filter_layer_dict.items():
            _dump_mlp(f"filter_{k_}", f, out, tag)
        for k_, f in model.dis_layer_dict.items():
            _dump_mlp(f"dis_{k_}", f, out, tag)

    dump("init")
    opt_p = torch.optim.Adam([model.user_embedding_layer.weight, model.item_embedding_layer.weight], lr=1e-3,
                             weight_decay=1e-4)
    dparams = [p_ for f in model.dis_layer_dict.values() for p_ in f.parameters()]
    if aggr == "LBA":
        dparams += list(model.aggr_layer.parameters())
    opt_d = torch.optim.Adam(dparams, lr=1e-3, weight_decay=1e-4)
    opt_f = torch.optim.Adam([p_ for f in model.filter_layer_dict.values() for p_ in f.parameters()], lr=1e-3,
                             weight_decay=1e-4)
    model.train()

    def batch(s):
        sel = rng.integers(0, n_inter, B)
        u = tu[sel]
        inter = Interaction({"user_id": torch.from_numpy(u), "item_id": torch.from_numpy(ti[sel]),
                             "rating": torch.from_numpy(tr[sel]), "gender": torch.from_numpy(gender[u]),
                             "age": torch.from_numpy(age[u])})
        for k_ in ("user_id", "item_id", "rating"):
            out[f"{k_}{s}"] = inter[k_].numpy()
        return inter

    losses = []
    s = 0
    model.train_stage = "pretrain"
    for _ in range(pretrain_steps):
        inter = batch(s)
        opt_p.zero_grad()
        loss = model.calculate_loss(inter, None)
        loss.backward()
        opt_p.step()
        losses.append(loss.item())
        out[f"sst_list{s}"] = np.array([], dtype="<U6")
        s += 1
    dump("pretrained")
    model.train_stage = "finetune"
    sst_lists = [["gender", "age"], ["age"], ["gender"]]
    for r in range(n_rounds):
        for phase, (fn, opt) in enumerate(((model.calculate_loss, opt_f), (model.calculate_dis_loss, opt_d))):
            inter = batch(s)
            sst_list = sst_lists[r % len(sst_lists)]
            opt.zero_grad()
            loss = fn(inter, sst_list)
            loss.backward()
            if r == 0 and phase == 0:
                for k_, f in model.filter_layer_dict.items():
                    for kk, p_ in f.named_parameters():
                        if p_.grad is not None:
                            out[f"grad_filter_{k_}.{kk}@ft0"] = p_.grad.numpy().copy()
                with torch.no_grad():
                    out["predict_ft0"] = model.predict(inter).numpy().copy()
                    out["full_sort_ft0"] = model.full_sort_predict(
                        Interaction({"user_id": torch.tensor([1, 2, 5])})).numpy().copy()
            opt.step()
            losses.append(loss.item())
            out[f"sst_list{s}"] = np.array(sst_list)
            s += 1
    out["losses"] = np.array(losses, np.float32)
    dump("final")
    np.savez_compressed(os.path.join(OUT, f"fairgo_{name}.npz"), **out)
    print(f"fairgo_{name}: losses={losses}")


def run_all_fairgo():
    run_fairgo("pmf_lba", "LBA")
    run_fairgo("pmf_wap", "WAP", seed=42)
    run_fairgo("pmf_lva", "LVA", seed=43)
    run_fairgo("pmf_lba_3layers", "LBA", n_layers=3, seed=44)


def run_ingest():
    """The reference's create_dataset + data_preparation on the `messy` atomic files (oracle/make_test_data.write_messy)
    under each option set of make_test_data.INGEST_CASES -> tests/golden/ingest_<case>.npz (outputs only)."""
    import tempfile
    import yaml
    from recbole.config import Config
    from recbole.data import create_dataset, data_preparation
    from recbole.utils import init_seed
    import make_test_data as mtd
    root = tempfile.mkdtemp()
    name = mtd.write_messy(root)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    from recbole.data.dataset import Dataset
    from recbole.utils import FeatureType
    stock_fill_nan = Dataset._fill_nan

    def fill_nan(self):        # dataset.py:554-575 with assignments instead of `fillna(inplace=True)` (pandas 3 copy-on-write)
        for feat_name in self.feat_name_list:
            feat = getattr(self, feat_name)
            for field in feat:
                ftype = self.field2type[field]
                if ftype == FeatureType.TOKEN:
                    feat[field] = feat[field].fillna(value=0)
                elif ftype == FeatureType.FLOAT:
                    feat[field] = feat[field].fillna(value=feat[field].mean())
                else:
                    dtype = np.int64 if ftype == FeatureType.TOKEN_SEQ else float
                    feat[field] = feat[field].apply(lambda x: np.array([], dtype=dtype) if isinstance(x, float) else x)

    only = set(sys.argv[2:])
    try:
        for case, opts in mtd.INGEST_CASES.items():
            if only and case not in only:
                continue
            Dataset._fill_nan = fill_nan if case in mtd.INGEST_PATCH_FILL_NAN else stock_fill_nan
            with open("c.yaml", "w") as f:
                yaml.safe_dump(dict(mtd.INGEST_BASE, **opts, data_path=root, use_gpu=False, state="WARNING",
                                    show_progress=False, neg_sampling=None, fair_objective="value"), f)
            sys.argv = sys.argv[:1]
            config = Config(model="FOCF", dataset=name, config_file_list=["c.yaml"])
            init_seed(config["seed"], config["reproducibility"])
            dataset = create_dataset(config)
            built = dataset.build()                       # the split itself (data_preparation re-sorts train for FOCF)
            out = dict(n_users=dataset.user_num, n_items=dataset.item_num,
                       user_tokens=np.array([str(t) for t in dataset.field2id_token["user_id"]]),
                       item_tokens=np.array([str(t) for t in dataset.field2id_token["item_id"]]))
            for k, part in zip(("train", "valid", "test"), built):
                f_ = part.inter_feat
                for col in ("user_id", "item_id", "rating", "timestamp", "label"):
                    out[f"{k}_{col}"] = f_[col].numpy() if len(f_) else np.zeros(0)
            uf = dataset.get_user_feature()
            for col in ("gender", "age", "occupation"):
                out["user_" + col] = uf[col].numpy()[1:]
            for idf in (config["preload_weight"] or {}):
                out["preload_" + idf] = dataset.get_preload_weight(idf)
            np.savez_compressed(os.path.join(OUT, f"ingest_{case}.npz"), **out)
            print("ingest", case, dataset.user_num, dataset.item_num, [len(p.inter_feat) for p in built])
    finally:
        Dataset._fill_nan = stock_fill_nan
        os.chdir(cwd)


from make_test_data import FAMILY_E2E, FAMILY_E2E_BASE, float_gender_copy  # noqa: E402


def run_family_e2e(model_name):
    """The reference's run_recbole steps (quick_start.py:20-71) for one MLP family on ml-100k from the atomic files:
    per-epoch train losses, per-epoch validation results and the test result -> tests/golden/e2e_<model>.npz"""
    import tempfile
    import yaml
    from recbole.config import Config
    from recbole.data import create_dataset, data_preparation
    from recbole.utils import init_seed, get_model, get_trainer
    root = float_gender_copy(tempfile.mkdtemp())
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    try:
        with open("c.yaml", "w") as f:
            yaml.safe_dump(dict(FAMILY_E2E_BASE, **FAMILY_E2E[model_name], data_path=root, use_gpu=False, state="WARNING",
                                show_progress=False), f)
        sys.argv = sys.argv[:1]
        config = Config(model=model_name, dataset="ml-100k", config_file_list=["c.yaml"])
        init_seed(config["seed"], config["reproducibility"])
        dataset = create_dataset(config)
        train_data, valid_data, test_data = data_preparation(config, dataset)
        model = get_model(config["model"])(config, train_data.dataset).to(config["device"])
        trainer = get_trainer(config["MODEL_TYPE"], config["model"])(config, model)
        trainer.eval_collector.data_collect(train_data)
        # per-step losses of the first epoch (the trajectories of these models are chaotic: BatchNorm stacks + Adam amplify
        # float32 rounding, so only the first steps of a run are comparable at tight tolerance)
        steps = []
        o_loss = model.calculate_loss
        o_dis = getattr(model, "calculate_dis_loss", None)
        model.calculate_loss = lambda *a, **k: (lambda v: (steps.append(("f", float(v))), v)[1])(o_loss(*a, **k))
        if o_dis is not None:
            model.calculate_dis_loss = lambda *a, **k: (lambda v: (steps.append(("d", float(v))), v)[1])(o_dis(*a, **k))
        losses, valids = [], []
        first_epoch_steps = None
        for ep in range(config["epochs"]):
            loss = trainer._train_epoch(train_data, ep)
            if first_epoch_steps is None:
                first_epoch_steps = list(steps)
            losses.append([float(x) for x in loss] if isinstance(loss, tuple) else [float(loss)])
            _, res = trainer._valid_epoch(valid_data)
            valids.append(res)
        test = trainer.evaluate(test_data, load_best_model=False)
        if model_name.startswith("PFCN") and len(test) == 1:         # {'sm-[gender]': metrics}
            test = next(iter(test.values()))
        names = list(test.keys())
        f_steps = [v for t_, v in first_epoch_steps if t_ == "f"]
        d_steps = [v for t_, v in first_epoch_steps if t_ == "d"][-len(f_steps):] if o_dis is not None else []
        out = dict(epoch_losses=np.array(losses, np.float64), metric_names=np.array(names),
                   first_epoch_step_losses=np.array(f_steps, np.float64), first_epoch_dis_step_losses=np.array(d_steps, np.float64),
                   valid_metrics=np.array([[float(r[k]) for k in names] for r in valids]),
                   test_metrics=np.array([float(test[k]) for k in names]), n_users=dataset.user_num, n_items=dataset.item_num)
        np.savez_compressed(os.path.join(OUT, f"e2e_{model_name.lower()}.npz"), **out)
        print(f"e2e {model_name}: losses {losses} valid {[dict(zip(names, v)) for v in out['valid_metrics']]}")
    finally:
        os.chdir(cwd)


def run_focf_uni_e2e():
    """FOCF on ml-100k through the reference's run_recbole steps with the evaluation mode of its own YAML (`uni<N>`,
    FOCF.yaml:28): the numpy RNG stream interleaves FOCFDataLoader's item draws with the negatives the evaluation loader
    draws at every validation -> tests/golden/e2e_focf_uni.npz (per-epoch losses, validation / test metrics)"""
    import tempfile
    import yaml
    from recbole.config import Config
    from recbole.data import create_dataset, data_preparation
    from recbole.utils import init_seed, get_model, get_trainer
    from make_test_data import FOCF_UNI_E2E
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    try:
        with open("c.yaml", "w") as f:
            yaml.safe_dump(dict(FOCF_UNI_E2E, use_gpu=False, state="WARNING", show_progress=False), f)
        sys.argv = sys.argv[:1]
        config = Config(model="FOCF", dataset="ml-100k", config_file_list=["c.yaml"])
        init_seed(config["seed"], config["reproducibility"])
        dataset = create_dataset(config)
        train_data, valid_data, test_data = data_preparation(config, dataset)
        model = get_model("FOCF")(config, train_data.dataset).to(config["device"])
        trainer = get_trainer(config["MODEL_TYPE"], config["model"])(config, model)
        trainer.eval_collector.data_collect(train_data)
        losses, valids = [], []
        for ep in range(config["epochs"]):
            losses.append(float(trainer._train_epoch(train_data, ep)))
            valids.append(trainer.evaluate(valid_data, load_best_model=False))
        test = trainer.evaluate(test_data, load_best_model=False)
        names = list(test.keys())
        np.savez_compressed(os.path.join(OUT, "e2e_focf_uni.npz"), epoch_losses=np.array(losses), metric_names=np.array(names),
                            valid_metrics=np.array([[float(r[k]) for k in names] for r in valids]),
                            test_metrics=np.array([float(test[k]) for k in names]))
        print("e2e focf uni:", losses, {k: float(v) for k, v in test.items()})
    finally:
        os.chdir(cwd)


def run_fairgo_e2e():
    """FairGo_PMF on ml-100k through the reference's FairGoTrainer.fit steps (trainer.py:583-592): validated pretrain with
    early stopping and best-checkpoint reload (606-685), reset_params, alternating fine-tune epochs with a validation after
    each, test evaluation; full-sort mode, all 12 metrics -> tests/golden/e2e_fairgo_pmf.npz"""
    import tempfile
    import yaml
    from recbole.config import Config
    from recbole.data import create_dataset, data_preparation
    from recbole.utils import init_seed, get_model, get_trainer
    from make_test_data import FAIRGO_E2E
    root = float_gender_copy(tempfile.mkdtemp())
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    try:
        with open("c.yaml", "w") as f:
            yaml.safe_dump(dict(FAIRGO_E2E, data_path=root, use_gpu=False, state="WARNING", show_progress=False,
                                save_sst_embed=False), f)
        sys.argv = sys.argv[:1]
        config = Config(model="FairGo_PMF", dataset="ml-100k", config_file_list=["c.yaml"])
        init_seed(config["seed"], config["reproducibility"])
        dataset = create_dataset(config)
        train_data, valid_data, test_data = data_preparation(config, dataset)
        model = get_model("FairGo_PMF")(config, train_data.dataset).to(config["device"])
        trainer = get_trainer(config["MODEL_TYPE"], config["model"])(config, model)
        per_epoch = []           # validation of every pretrain epoch (the reference keeps only the best)
        o_valid = trainer._valid_epoch
        trainer._valid_epoch = lambda *a, **k: (lambda r: (per_epoch.append(r[1]), r)[1])(o_valid(*a, **k))
        pre_score, pre_result = trainer.pretrain(train_data, valid_data, verbose=False, saved=True)
        trainer._valid_epoch = o_valid
        pre_losses = [float(trainer.train_loss_dict[k]) for k in sorted(trainer.train_loss_dict)]
        trainer.reset_params()
        trainer.eval_collector.data_collect(train_data)
        losses, valids = [], []
        for ep in range(config["epochs"]):
            dis_loss, filter_loss = trainer._train_epoch(train_data, ep)
            losses.append([float(dis_loss), float(filter_loss)])
            _, res = trainer._valid_epoch(valid_data)
            valids.append(res)
        test = trainer.evaluate(test_data, load_best_model=False)
        names = list(test.keys())
        np.savez_compressed(os.path.join(OUT, "e2e_fairgo_pmf.npz"), pretrain_losses=np.array(pre_losses),
                            pretrain_best_valid=float(pre_score), epoch_losses=np.array(losses), metric_names=np.array(names),
                            pretrain_valid=np.array([float(pre_result[k]) for k in names]),
                            pretrain_valid_per_epoch=np.array([[float(r[k]) for k in names] for r in per_epoch]),
                            pretrain_best_epoch=int(np.argmax([float(r["ndcg@5"]) for r in per_epoch])),
                            valid_metrics=np.array([[float(r[k]) for k in names] for r in valids]),
                            test_metrics=np.array([float(test[k]) for k in names]))
        print("e2e fairgo:", pre_losses, pre_score, losses, float(test["ndcg@5"]))
    finally:
        os.chdir(cwd)


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "e2e_fairgo":
        run_fairgo_e2e()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "e2e":
        for m in FAMILY_E2E:
            run_family_e2e(m)
        run_focf_uni_e2e()
        run_fairgo_e2e()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "e2e_focf":
        run_focf_uni_e2e()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "ingest":
        run_ingest()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "uni":
        run_uni_eval("uni100", users_per_batch=1)
        run_uni_eval("uni100_batched", users_per_batch=3, seed=52)
        run_uni_eval("uni20_small_catalog", n_items=60, neg_num=20, users_per_batch=2, seed=53)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "uni_unfair":
        run_uni_eval("uni100_unfair", users_per_batch=1, seed=54, all_metrics=True)
        run_uni_eval("uni10_unfair_small", n_items=90, neg_num=10, users_per_batch=1, seed=55, all_metrics=True)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "fairgo":
        run_all_fairgo()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "pfcn":
        run_all_pfcn()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "nfcf":
        run_nfcf("ncf", fair=False)
        run_nfcf("fair", fair=True)
        run_nfcf("fair_d64", fair=True, n_users=200, n_items=120, d=64, hidden=(128, 64), B=512, seed=22)
        return
    for obj in ["none", "value", "absolute", "under", "over", "nonparity"]:
        run_focf_train(obj, obj, seed=10 + len(obj))
    run_focf_train("value_d64", "value", n_users=300, n_items=120, d=64, target_rows=900, seed=3, scale=0.3)
    run_focf_train("value_single_group", "value", single_group_step=1, seed=4)
    run_focf_train("value_shuffled", "value", shuffle=True, seed=5)
    run_focf_train("value_float_sst", "value", float_sst=True, seed=6, fair_weight=0.5)
    run_focf_eval("ties", seed=1)
    run_focf_eval("tiefree", seed=2, positive_only=True)
    run_focf_eval("float_sst", seed=3, positive_only=True, float_sst=True, d=64, n_users=130, n_items=300,
                  users_per_batch=13)
    run_ml100k()
    run_nfcf("ncf", fair=False)
    run_nfcf("fair", fair=True)
    run_nfcf("fair_d64", fair=True, n_users=200, n_items=120, d=64, hidden=(128, 64), B=512, seed=22)
    run_nfcf("fair_g3", fair=True, n_users=120, n_items=60, d=32, hidden=(64, 32), B=384, seed=23, n_groups=3, fair_weight=0.5)
    run_nfcf("fair_g5", fair=True, n_users=150, n_items=40, d=16, hidden=(32, 16), B=512, seed=24, n_groups=5, fair_weight=1.0)
    run_all_pfcn()
    run_all_fairgo()
    run_uni_eval("uni100", users_per_batch=1)
    run_uni_eval("uni100_batched", users_per_batch=3, seed=52)
    run_uni_eval("uni20_small_catalog", n_items=60, neg_num=20, users_per_batch=2, seed=53)
    run_uni_eval("uni100_unfair", users_per_batch=1, seed=54, all_metrics=True)
    run_uni_eval("uni10_unfair_small", n_items=90, neg_num=10, users_per_batch=1, seed=55, all_metrics=True)
    run_ingest()


if __name__ == "__main__":
    main()
