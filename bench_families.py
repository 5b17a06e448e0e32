"""bench legs for the PFCN and FairGo families (BASELINE.json configs[2] and configs[3]) -- imported by bench.py.

A "step" here is what one batch costs in an epoch where both phases run (trainer.py:875-898 / 687-704): one
filter-phase step (calculate_loss -> backward -> Adam over base + filters) plus one discriminator-phase step
(calculate_dis_loss -> backward -> Adam over the discriminators) on a 2048-row batch of the ML-1M shape.
`value` = interactions/s over those two phases, device-timed with CUDA events, batches resident in HBM;
`e2e`   = the same through the public model / trainer API with HOST (pinned) batches, H2D per step and the loss read
          back every phase (trainer.py:191);
`cpu_baseline` = oracle/pfcn_oracle.py / oracle/fairgo_oracle.py (the reference's op sequence on stock torch CPU kernels,
          all host threads) on a bounded sample of the same steps.
"""
import os
import statistics
import time

import numpy as np

ML1M = dict(n_users=6041, n_items=3707, d=64, batch=2048)


def _user_feats(rng, nu):
    return {"gender": (rng.random(nu) < 0.28).astype(np.float32), "age": rng.integers(0, 7, nu).astype(np.float32),
            "occupation": rng.integers(0, 21, nu).astype(np.float32)}


class _DS:
    def __init__(self, nu, ni, feats, coo=None):
        import torch

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


def _timed_steps(step_fn, batches_dev, n_warm, n_steps, flush):
    import torch
    for k in range(n_warm):
        step_fn(batches_dev[k % len(batches_dev)])
    torch.cuda.synchronize()
    ms = []
    for k in range(n_steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step_fn(batches_dev[k % len(batches_dev)])
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return ms


def _graphed_legs(pkg, dev, flush, hosts, devs, n_steps, phases, sst_list):
    """(device-timed ms per step, e2e seconds per step) of the CUDA-graph-replayed step: both phases per batch.
    e2e: pinned host batch -> H2D into the graph's static buffers -> replay -> loss D2H + sync, every phase."""
    import torch
    from recbole_fairrec_b200.graphed import GraphedStep
    graphs = [GraphedStep(fn, opt, sst_list, devs[0], dev) for fn, opt in phases]

    def step(inter):
        return [g.run(inter) for g in graphs]

    for k in range(3):
        step(devs[k % len(devs)])
    torch.cuda.synchronize()
    ms = []
    for k in range(n_steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step(devs[k % len(devs)])
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(n_steps):
        inter = pkg.Interaction(hosts[k % len(hosts)])
        for g in graphs:
            _ = g.run(inter).item()
    torch.cuda.synchronize()
    return ms, (time.perf_counter() - t0) / n_steps


def _profile(step_fn, batch, n=5):
    from recbole_fairrec_b200 import _lib
    _lib.profile_enable(True)
    for _ in range(n):
        step_fn(batch)
    prof = _lib.profile_report()
    _lib.profile_enable(False)
    tot = sum(v[1] for v in prof.values()) or 1.0
    shares = {k: round(v[1] / tot, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]}
    global LAST_COUNTS     # launches per step and mean microseconds per launch of every library kernel of the step
    LAST_COUNTS = {k: [round(v[0] / n, 2), round(1e3 * v[1] / max(v[0], 1), 1)] for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    return shares, sum(v[0] for v in prof.values()) / n, tot / n


LAST_COUNTS = {}


def bench_pfcn(dev, flush, n_steps=30, cpu=True, seed=2020):
    import torch

    import recbole_fairrec_b200 as pkg
    w = ML1M
    rng = np.random.default_rng(seed)
    feats = _user_feats(rng, w["n_users"])
    attrs = list(feats)
    cfg = pkg.Config(embedding_size=w["d"], sst_attr_list=attrs, filter_mode="sm", dropout=0.2, dis_dropout=0.3,
                     dis_weight=10.0, dis_hidden_size_list=[128, 256, 128, 128, 64, 32], mlp_hidden_size_list=[64, 32, 16],
                     activation="leakyrelu", device=dev, learning_rate=1e-3, weight_decay=1e-4, train_epoch_interval=1)
    torch.manual_seed(seed)
    model = pkg.PFCN_MLP(cfg, _DS(w["n_users"], w["n_items"], feats)).to(dev)
    trainer = pkg.PFCNTrainer(cfg, model)
    model.train()
    sst_list = ["gender", "occupation"]
    B = w["batch"]

    def host_batch():
        u = rng.integers(1, w["n_users"], B)
        return {"user_id": torch.from_numpy(u).pin_memory(),
                "item_id": torch.from_numpy(rng.integers(1, w["n_items"], B)).pin_memory(),
                "neg_item_id": torch.from_numpy(rng.integers(1, w["n_items"], B)).pin_memory(),
                **{a: torch.from_numpy(feats[a][u]).pin_memory() for a in attrs}}

    hosts = [host_batch() for _ in range(8)]
    devs = [pkg.Interaction({k: v.to(dev) for k, v in h.items()}) for h in hosts]

    def step(inter, read_loss=False):
        out = []
        for fn, opt in ((model.calculate_loss, trainer.optimizer_filter), (model.calculate_dis_loss, trainer.optimizer_dis)):
            opt.zero_grad()
            loss = fn(inter, sst_list)
            if read_loss:
                out.append(loss.item())
            loss.backward()
            opt.step()
        return out

    eager_ms = _timed_steps(step, devs, 3, n_steps, flush)
    shares, launches, kernel_ms = _profile(step, devs[0])
    ms, t_e2e = _graphed_legs(pkg, dev, flush, hosts, devs, n_steps,
                              [(model.calculate_loss, trainer.optimizer_filter), (model.calculate_dis_loss, trainer.optimizer_dis)],
                              sst_list)
    h2d = sum(v.numel() * v.element_size() for v in hosts[0].values())
    out = {"metric": "PFCN_MLP train interactions/s", "value": B / (statistics.mean(ms) * 1e-3), "unit": "interactions/s",
           "ms_per_step": statistics.mean(ms), "steps": n_steps, "launch_mode": "cuda graph replay (one per phase)",
           "eager_ms_per_step": statistics.mean(eager_ms),
           "config": {"workload": "pfcn_mlp_ml1m", **w, "filter_mode": "sm (7 filters)", "sst_list": sst_list,
                      "attrs": {a: int(len(np.unique(feats[a][1:]))) for a in attrs}, "dropout": [0.2, 0.3],
                      "step": "filter-phase step + discriminator-phase step on one batch"},
           "e2e": {"value": B / t_e2e, "unit": "interactions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8},
           "gpu_launches_per_step": launches, "kernel_ms_per_step": kernel_ms, "kernel_shares": shares,
           "kernel_launches_and_us": dict(LAST_COUNTS)}
    if cpu:
        out["cpu_baseline"] = cpu_pfcn(model, feats, hosts, sst_list)
    return out


def cpu_pfcn(model, feats, hosts, sst_list, n=3):
    """the reference's op sequence (oracle/pfcn_oracle.py) on stock torch CPU kernels, all host threads"""
    import torch
    from oracle import pfcn_oracle as po
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    st = {f"base.{k}": v.detach().cpu().clone() for k, v in model.state_dict().items()}
    for name, mods in (("filter", model.filter_layer), ("dis", model.dis_layer_dict)):
        for k, m in mods.items():
            st.update({f"{name}_{k}.{kk}": v.detach().cpu().clone() for kk, v in m.state_dict().items()})
    fkeys, dkeys = po.param_groups(st)
    for k in fkeys + dkeys:
        st[k].requires_grad_(True)
    opt_f = torch.optim.Adam([st[k] for k in fkeys], lr=1e-3, weight_decay=1e-4)
    opt_d = torch.optim.Adam([st[k] for k in dkeys], lr=1e-3, weight_decay=1e-4)
    attrs = list(feats)
    sst_dict = {s: 2 ** i for i, s in enumerate(attrs)}
    sst_size = {s: len(np.unique(feats[s][1:])) for s in attrs}

    def step(h):
        labels = {a: h[a] for a in attrs}
        opt_f.zero_grad()
        po.calculate_loss(st, "PFCN_MLP", h["user_id"], h["item_id"], h["neg_item_id"], labels, sst_list, sst_dict, sst_size,
                          "sm", 7, "leakyrelu", 10.0).backward()
        opt_f.step()
        opt_d.zero_grad()
        po.dis_loss(st, "PFCN_MLP", h["user_id"], labels, sst_list, sst_dict, sst_size, "sm", 7, "leakyrelu").backward()
        opt_d.step()

    step(hosts[0])
    t0 = time.perf_counter()
    for k in range(n):
        step(hosts[(k + 1) % len(hosts)])
    dt = (time.perf_counter() - t0) / n
    B = hosts[0]["user_id"].numel()
    return {"value": B / dt, "unit": "interactions/s", "cores": cores, "kind": "port",
            "sample": f"{n} filter+discriminator steps of {B} rows (dropout off in the port)", "ms_per_step": 1e3 * dt}


def bench_fairgo(dev, flush, n_steps=20, cpu=True, seed=2020):
    import scipy.sparse as sp
    import torch

    import recbole_fairrec_b200 as pkg
    from recbole_fairrec_b200 import synth
    w = ML1M
    rng = np.random.default_rng(seed)
    feats = _user_feats(rng, w["n_users"])
    attrs = ["gender", "age"]
    feats = {a: feats[a] for a in attrs}
    uid, iid, rating, _ = synth.interactions(w["n_users"], w["n_items"], 1_000_209, seed)
    train = synth.split_by_user(uid, iid, rating, seed=seed)[0]
    coo = sp.coo_matrix((train[2], (train[0].astype(np.int64), train[1].astype(np.int64))), shape=(w["n_users"], w["n_items"]))
    cfg = pkg.Config(embedding_size=w["d"], sst_attr_list=attrs, n_layers=2, activation="leakyrelu",
                     dis_hidden_size_list=[16, 8, 4], filter_hidden_size_list=[128, 64], fair_weight=0.1,
                     load_pretrain_weight=False, aggr_method="LBA", vs_weights=[4, 1], device=dev, learning_rate=1e-3,
                     weight_decay=1e-4, train_epoch_interval=1, pretrain_epochs=1)
    torch.manual_seed(seed)
    model = pkg.FairGo_PMF(cfg, _DS(w["n_users"], w["n_items"], feats, coo)).to(dev)
    with torch.no_grad():
        model.user_embedding_layer.weight.mul_(0.3)
        model.item_embedding_layer.weight.mul_(0.3)
    trainer = pkg.FairGoTrainer(cfg, model)
    model.train_stage = "finetune"
    model.train()
    B = w["batch"]
    n_train = len(train[0])

    def host_batch():
        sel = rng.integers(0, n_train, B)
        u = train[0][sel].astype(np.int64)
        return {"user_id": torch.from_numpy(u).pin_memory(), "item_id": torch.from_numpy(train[1][sel].astype(np.int64)).pin_memory(),
                "rating": torch.from_numpy(train[2][sel]).pin_memory(),
                **{a: torch.from_numpy(feats[a][u]).pin_memory() for a in attrs}}

    hosts = [host_batch() for _ in range(8)]
    devs = [pkg.Interaction({k: v.to(dev) for k, v in h.items()}) for h in hosts]

    def step(inter, read_loss=False):
        out = []
        for fn, opt in ((model.calculate_loss, trainer.optimizer_filter), (model.calculate_dis_loss, trainer.optimizer_dis)):
            opt.zero_grad()
            loss = fn(inter, attrs)
            if read_loss:
                out.append(loss.item())
            loss.backward()
            opt.step()
        return out

    eager_ms = _timed_steps(step, devs, 3, n_steps, flush)
    shares, launches, kernel_ms = _profile(step, devs[0])
    model._matrix()
    ms, t_e2e = _graphed_legs(pkg, dev, flush, hosts, devs, n_steps,
                              [(model.calculate_loss, trainer.optimizer_filter), (model.calculate_dis_loss, trainer.optimizer_dis)],
                              attrs)
    N = w["n_users"] + w["n_items"]
    nnz = model._norm_csr.nnz
    out = {"metric": "FairGo_PMF (LBA) fine-tune interactions/s", "value": B / (statistics.mean(ms) * 1e-3),
           "unit": "interactions/s", "ms_per_step": statistics.mean(ms), "steps": n_steps,
           "launch_mode": "cuda graph replay (one per phase)", "eager_ms_per_step": statistics.mean(eager_ms),
           "config": {"workload": "fairgo_pmf_lba_ml1m", **w, "n_layers": 2, "graph_rows": N, "graph_nnz": int(nnz),
                      "filters": "[64,128,64,64] x 2 attributes over all graph rows", "discriminators": "[64,16,8,4,{1|7}]",
                      "step": "filter-phase step + discriminator-phase step on one batch (each runs the full-table "
                              "filters and 2 SpMM layers forward and backward)"},
           "e2e": {"value": B / t_e2e, "unit": "interactions/s",
                   "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in hosts[0].values()), "d2h_bytes_per_step": 8},
           "gpu_launches_per_step": launches, "kernel_ms_per_step": kernel_ms, "kernel_shares": shares}
    if cpu:
        out["cpu_baseline"] = cpu_fairgo(model, feats, train, hosts, attrs)
    return out


def cpu_fairgo(model, feats, train, hosts, attrs, n=2):
    import torch
    from oracle import fairgo_oracle as go
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    st = {f"base.{k}": v.detach().cpu().clone() for k, v in model.state_dict().items()}
    for name, mods in (("filter", model.filter_layer_dict), ("dis", model.dis_layer_dict)):
        for k, m in mods.items():
            st.update({f"{name}_{k}.{kk}": v.detach().cpu().clone() for kk, v in m.state_dict().items()})
    for v in st.values():
        v.requires_grad_(True)
    nu, ni = model.n_users, model.n_items
    L = go.to_torch_sparse(go.norm_matrix(train[0].astype(np.int64), train[1].astype(np.int64), train[2], nu, ni))
    dkeys = [k for k in st if k.startswith("dis_") or k.startswith("base.aggr_layer")]
    fkeys = [k for k in st if k.startswith("filter_")]
    opt_d = torch.optim.Adam([st[k] for k in dkeys], lr=1e-3, weight_decay=1e-4)
    opt_f = torch.optim.Adam([st[k] for k in fkeys], lr=1e-3, weight_decay=1e-4)
    sst_size = {s: len(np.unique(feats[s][1:])) for s in attrs}
    assert attrs == go.ATTRS

    def step(h):
        labels = {a: h[a] for a in attrs}
        opt_f.zero_grad()
        go.calculate_loss(st, L, "finetune", h["user_id"], h["item_id"], h["rating"], labels, attrs, sst_size, nu, 2, "LBA",
                          [0.8, 0.2], 0.1).backward()
        opt_f.step()
        opt_d.zero_grad()
        go.dis_loss(st, L, h["user_id"], labels, attrs, sst_size, nu, 2, "LBA", [0.8, 0.2]).backward()
        opt_d.step()

    step(hosts[0])
    t0 = time.perf_counter()
    for k in range(n):
        step(hosts[(k + 1) % len(hosts)])
    dt = (time.perf_counter() - t0) / n
    B = hosts[0]["user_id"].numel()
    return {"value": B / dt, "unit": "interactions/s", "cores": cores, "kind": "port",
            "sample": f"{n} filter+discriminator fine-tune steps of {B} rows", "ms_per_step": 1e3 * dt}


def bench_sampled_eval(dev, flush, cpu=True, seed=2020, neg_num=100, K=10):
    """Sampled-negative (uni100) fair ranking evaluation at the ML-1M shape: all valid users, 100
    uniform negatives per positive, FOCF scorer; users/s of scoring + candidate top-K + metrics, device-timed."""
    import torch

    import recbole_fairrec_b200 as pkg
    from recbole_fairrec_b200 import synth
    w = ML1M
    uid, iid, rating, gender = synth.interactions(w["n_users"], w["n_items"], 1_000_209, seed)
    train, valid, test = synth.split_by_user(uid, iid, rating, seed=seed)
    users, hist, pos = synth.eval_lists(train, valid, test, "valid")
    rng = np.random.default_rng(seed)
    t0 = time.perf_counter()
    neg = pkg.sample_negatives(pos, hist, w["n_items"], neg_num, rng)
    t_sample = time.perf_counter() - t0
    data = pkg.SampledEvalData(users, pos, neg, {"gender": gender.astype(np.int64)}, dev)
    metrics = ["NDCG", "Recall", "Hit", "MRR", "DifferentialFairness", "GiniIndex", "PopularityPercentage",
               "NonParityUnfairness"]
    cfg = pkg.Config(topk=[K], metrics=metrics, sst_attr_list=["gender"], device=dev, eval_args={"mode": f"uni{neg_num}"})
    counts = np.bincount(train[1], minlength=w["n_items"])
    ev = pkg.SampledEvaluator(cfg, w["n_items"], {int(i): int(c) for i, c in enumerate(counts) if c > 0})
    g = torch.Generator(device=dev).manual_seed(seed)
    U = torch.randn(w["n_users"], w["d"], device=dev, generator=g) * 0.2
    I = torch.randn(w["n_items"], w["d"], device=dev, generator=g) * 0.2
    score_fn = pkg.SampledEvaluator.dot_scorer(U, I, 5.0)
    res = ev.evaluate(score_fn, data)
    for _ in range(3):      # warm-up passes: the first pass after evaluate() re-grows the caching allocator's pools (a 15 ms
        ev.collect(score_fn, data)   # cudaMalloc stall on the host, seen as 1.3 / 3.5 / 6 / 20 ms means in earlier lines)
    torch.cuda.synchronize()
    ms = []
    for _ in range(7):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ev.collect(score_fn, data)
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    ms_all, ms = ms, [statistics.median(ms)]
    from recbole_fairrec_b200 import _lib
    _lib.profile_enable(True)
    ev.collect(score_fn, data)
    prof = _lib.profile_report()
    _lib.profile_enable(False)
    tot = sum(v[1] for v in prof.values()) or 1.0
    n_cand = int(data.cand_items.numel())
    out = {"metric": "sampled-negative (uni100) fair-eval users/s", "value": data.n / (statistics.mean(ms) * 1e-3),
           "unit": "users/s", "ms_per_pass": statistics.mean(ms), "ms_passes": [round(x, 3) for x in ms_all],
           "timing": "median of 7 passes after 3 warm-up passes, L2 flushed before each", "n_users": data.n, "candidates": n_cand,
           "config": {"workload": "focf_ml1m_uni100", **w, "neg_per_positive": neg_num, "K": K},
           "kernel_shares": {k: round(v[1] / tot, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:6]},
           "host_negative_sampling_s": t_sample, "ndcg@10": float(res[f"ndcg@{K}"])}
    if cpu:
        from oracle import sampled_oracle as so
        n_cpu = 300
        Uc, Ic = U.cpu().numpy(), I.cpu().numpy()
        cands = [(np.asarray(p), np.asarray(q)) for p, q in zip(pos[:n_cpu], neg[:n_cpu])]
        t0 = time.perf_counter()
        rows = so.dense_rows(Uc, Ic, users[:n_cpu], cands, w["n_items"], 5.0)
        so.collect(rows, cands, K)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n_cpu / dt, "unit": "users/s", "cores": 1, "kind": "port",
                               "sample": f"{n_cpu} users: dense -inf rows + canonical top-K + hit bits (oracle/sampled_oracle.py)"}
    return out


def bench_nfcf(dev, flush, n_steps=30, cpu=True, seed=2020):
    """NFCF stage 2 (nfcf.py:99-110 with a pretrain loaded: BCE + fair_weight * differential fairness, user table
    frozen) at the ML-1M shape, NFCF.yaml widths: batches of 1024 positives + 1024 uniform negatives (labels 1 | 0), one
    step = calculate_loss -> backward -> Adam over the item table and the tower (NFCFTrainer's loop body)."""
    import tempfile

    import torch

    import recbole_fairrec_b200 as pkg
    w = ML1M
    rng = np.random.default_rng(seed)
    feats = {"gender": (rng.random(w["n_users"]) < 0.28).astype(np.float32)}
    base = dict(embedding_size=w["d"], sst_attr_list=["gender"], mlp_hidden_size=[128, 64], dropout=0.2, fair_weight=0.1,
                device=dev, learning_rate=1e-3, weight_decay=1e-6)
    ds = _DS(w["n_users"], w["n_items"], feats)
    torch.manual_seed(seed)
    stage1 = pkg.NFCF(pkg.Config(load_pretrain_path=None, **base), ds)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "ncf.pth")
        torch.save({"state_dict": stage1.state_dict()}, path)
        cfg = pkg.Config(load_pretrain_path=path, **base)
        model = pkg.NFCF(cfg, ds).to(dev)
    trainer = pkg.NFCFTrainer(cfg, model)
    model.train()
    B = w["batch"]

    def host_batch():
        u = rng.integers(1, w["n_users"], B // 2)
        u = np.concatenate([u, u])
        label = np.zeros(B, np.float32)
        label[:B // 2] = 1.0
        return {"user_id": torch.from_numpy(u).pin_memory(),
                "item_id": torch.from_numpy(rng.integers(1, w["n_items"], B)).pin_memory(),
                "label": torch.from_numpy(label).pin_memory(), "gender": torch.from_numpy(feats["gender"][u]).pin_memory()}

    hosts = [host_batch() for _ in range(8)]
    devs = [pkg.Interaction({k: v.to(dev) for k, v in h.items()}) for h in hosts]

    def step(inter, read_loss=False):
        trainer.optimizer.zero_grad()
        loss = model.calculate_loss(inter)
        out = loss.item() if read_loss else None
        loss.backward()
        trainer.optimizer.step()
        return out

    ms = _timed_steps(step, devs, 3, n_steps, flush)
    shares, launches, kernel_ms = _profile(step, devs[0])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(n_steps):
        step(pkg.Interaction(hosts[k % len(hosts)]), read_loss=True)
    torch.cuda.synchronize()
    t_e2e = (time.perf_counter() - t0) / n_steps
    model.check_flags()
    h2d = sum(v.numel() * v.element_size() for v in hosts[0].values())
    out = {"metric": "NFCF (stage 2) train interactions/s", "value": B / (statistics.mean(ms) * 1e-3), "unit": "interactions/s",
           "ms_per_step": statistics.mean(ms), "steps": n_steps, "launch_mode": "stream launches",
           "config": {"workload": "nfcf_ml1m", **w, "mlp_hidden_size": [128, 64], "dropout": 0.2, "fair_weight": 0.1,
                      "step": "BCE + differential-fairness loss -> backward -> Adam (item table + tower; user table frozen)"},
           "e2e": {"value": B / t_e2e, "unit": "interactions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
           "gpu_launches_per_step": launches, "kernel_ms_per_step": kernel_ms, "kernel_shares": shares}
    if cpu:
        out["cpu_baseline"] = cpu_nfcf(model, hosts)
    return out


def cpu_nfcf(model, hosts, n=3):
    """oracle/nfcf_oracle.py (numpy restatement of the reference step; BLAS threads) on a bounded sample of the same steps"""
    from oracle import nfcf_oracle as no
    f = lambda t: t.detach().cpu().numpy().copy()
    U, I = f(model.user_embedding.weight), f(model.item_embedding.weight)
    lin = list(model.mlp_layers.linears())
    Ws, bs = [f(m.weight) for m in lin], [f(m.bias) for m in lin]
    batches = [(h["user_id"].numpy(), h["item_id"].numpy(), h["label"].numpy(), h["gender"].numpy())
               for h in hosts[:n + 1]]
    no.train_steps(U, I, Ws, bs, batches[:1], True, 0.1, 1e-3, 1e-6, True)
    t0 = time.perf_counter()
    no.train_steps(U, I, Ws, bs, batches[1:], True, 0.1, 1e-3, 1e-6, True)
    dt = (time.perf_counter() - t0) / n
    B = hosts[0]["user_id"].numel()
    return {"value": B / dt, "unit": "interactions/s", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": f"{n} steps of {B} rows (dropout off in the port)", "ms_per_step": 1e3 * dt}


def bench_pfcn_dp(dev, rank, world, group, n_steps=20, seed=2020):
    """PFCN_MLP data parallel (bench.py --gpus N, all ranks call this): every rank trains on ITS 2048 rows of a global batch
    of 2048 * world rows (weak scaling), BatchNorm column sums through NVLink peer memory (fr_mlp_chain_*_dp), one NCCL
    all-reduce of the gradient shares per phase.  `value` = global interactions/s (eager launches, device-timed, max over
    ranks); `dp_check` = first-step gradients and the losses of two steps against the single-GPU trainer on the whole
    global batch (rank 0 computes the reference)."""
    import copy

    import torch
    import torch.distributed as dist

    import recbole_fairrec_b200 as pkg
    from recbole_fairrec_b200 import ops
    w = ML1M
    rng = np.random.default_rng(seed)
    feats = _user_feats(rng, w["n_users"])
    attrs = list(feats)
    cfg = pkg.Config(embedding_size=w["d"], sst_attr_list=attrs, filter_mode="sm", dropout=0.2, dis_dropout=0.3,
                     dis_weight=10.0, dis_hidden_size_list=[128, 256, 128, 128, 64, 32], mlp_hidden_size_list=[64, 32, 16],
                     activation="leakyrelu", device=dev, learning_rate=1e-3, weight_decay=1e-4, train_epoch_interval=1)
    torch.manual_seed(seed)
    model = pkg.PFCN_MLP(cfg, _DS(w["n_users"], w["n_items"], feats)).to(dev)
    ref = copy.deepcopy(model)
    for src, dst in zip(model._dict_modules(), ref._dict_modules()):
        dst.load_state_dict(src.state_dict())
    B = w["batch"] * world
    sst_list = ["gender", "occupation"]

    def batch():
        u = rng.integers(1, w["n_users"], B)
        return pkg.Interaction({"user_id": torch.from_numpy(u).to(dev),
                                "item_id": torch.from_numpy(rng.integers(1, w["n_items"], B)).to(dev),
                                "neg_item_id": torch.from_numpy(rng.integers(1, w["n_items"], B)).to(dev),
                                **{a: torch.from_numpy(feats[a][u]).to(dev) for a in attrs}})

    batches = [batch() for _ in range(4)]
    dp = ops.ChainDP(rank, world, group, dev)
    trainer = pkg.PFCNTrainer(cfg, model, dp=dp)
    model.train()

    def step(m, tr, inter, shard, keep=None):
        out = []
        for fn, opt in ((m.calculate_loss, tr.optimizer_filter), (m.calculate_dis_loss, tr.optimizer_dis)):
            opt.zero_grad()
            loss = tr._dp_loss(fn)(tr._shard(inter), sst_list) if shard else fn(inter, sst_list)
            loss.backward()
            if keep is not None and not keep:
                if shard:
                    dp.all_reduce_grads(opt.opt.params)
                keep.extend(None if q.grad is None else q.grad.detach().clone() for q in (opt.opt.params if shard else opt.params))
                (opt.opt if shard else opt).step()
            else:
                opt.step()
            out.append(loss.detach())
        return out

    import itertools
    ops._seed_counter = itertools.count(777)
    g_dp = []
    l_dp = []
    for k in range(2):
        for t in step(model, trainer, batches[k], True, g_dp):
            t = t.clone().double()
            dist.all_reduce(t, group=group)
            l_dp.append(float(t))
    check = None
    ops.set_chain_dp(None)
    if rank == 0:
        ops._seed_counter = itertools.count(777)
        ref.train()
        rt = pkg.PFCNTrainer(cfg, ref)
        g_ref, l_ref = [], []
        for k in range(2):
            l_ref += [float(t) for t in step(ref, rt, batches[k], False, g_ref)]
        gerr = 0.0
        for a, b in zip(g_dp, g_ref):
            if a is not None and float(b.abs().max()) > 1e-7:
                gerr = max(gerr, float((a - b).abs().max() / b.abs().max()))
        lerr = max(abs(a - b) / max(abs(b), 1e-12) for a, b in zip(l_dp, l_ref))
        check = {"world": world, "first_step_grad_rel_err": gerr, "loss_rel_err_2_steps": lerr,
                 "pass": bool(gerr < 5e-5 and lerr < 1e-4),
                 "what": "PFCNTrainer(dp) on 2048 rows per rank vs the single-GPU trainer on the whole global batch (dropout on: "
                         "the masks hash the global row index, so both runs draw the same masks)"}
    ops.set_chain_dp(dp)
    dist.barrier(group=group)
    for k in range(3):
        step(model, trainer, batches[k % 4], True)
    torch.cuda.synchronize()
    dist.barrier(group=group)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(n_steps):
        step(model, trainer, batches[k % 4], True)
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n_steps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    ms_eager = float(t)
    # the same step as two CUDA-graph replays per batch (filter phase, discriminator phase): NCCL all-reduce and the flag
    # barriers are captured with the kernels
    ms, mode = ms_eager, "eager (host-paced)"
    try:
        from recbole_fairrec_b200.graphed import GraphedStep
        shard0 = trainer._shard(batches[0])
        graphs = [GraphedStep(trainer._dp_loss(fn), opt, sst_list, shard0, dev)
                  for fn, opt in ((model.calculate_loss, trainer.optimizer_filter),
                                  (model.calculate_dis_loss, trainer.optimizer_dis))]
        shards = [trainer._shard(bt) for bt in batches]
        for k in range(3):
            for g in graphs:
                g.run(shards[k % 4])
        torch.cuda.synchronize()
        dist.barrier(group=group)
        a.record()
        for k in range(n_steps):
            for g in graphs:
                g.run(shards[k % 4])
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / n_steps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        ms, mode = float(t), "cuda graph replay (one per phase; NCCL + flag barriers captured)"
        del graphs
        torch.cuda.synchronize()
    except Exception as e:
        mode = f"eager (host-paced; graph capture failed: {type(e).__name__}: {str(e)[:120]})"
    dist.barrier(group=group)
    timeout = int(dp.status.item())
    ops.set_chain_dp(None)
    dp.close()
    return {"metric": "PFCN_MLP train interactions/s (data parallel)", "value": B / (ms * 1e-3), "unit": "interactions/s",
            "ms_per_step": ms, "eager_ms_per_step": ms_eager, "steps": n_steps, "global_batch": B, "launch_mode": mode,
            "exchange": "BatchNorm column sums: peer-memory stores + flag barriers inside the step; gradient shares: one NCCL "
                        "all-reduce per phase", "xchg_timeout_flag": timeout, "dp_check": check}


def main():
    """stand-alone / as bench.py's `families` block (own process): ONE JSON line {leg: result}"""
    import argparse
    import json
    import sys

    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    dev = torch.device("cuda", 0)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    out = {}
    saved = os.dup(1)
    os.dup2(2, 1)          # keep stdout for the ONE JSON line
    try:
        for name, fn in (("pfcn_mlp", bench_pfcn), ("fairgo_pmf", bench_fairgo), ("nfcf", bench_nfcf),
                         ("sampled_eval", bench_sampled_eval)):
            try:
                out[name] = fn(dev, flush, cpu=not args.no_cpu_baseline)
            except Exception as e:
                out[name] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
