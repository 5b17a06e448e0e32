"""CPU restatement (stock torch CPU fp32 ops, functional, no reference import) of the FairGo_PMF training math: rating
MSE on (filtered) embeddings, node-level and ego-network-level discriminator losses over D^-1 A aggregations.
TEST INFRASTRUCTURE ONLY (tests/, smoke, bench cpu legs) -- never imported by the product package.

Parity status: PINNED against tests/golden/fairgo_pmf_*.npz (generated from the unmodified reference by
oracle/gen_golden.py `fairgo`) in tests/test_oracle_golden.py.  FairGo_GCN's fine-tune stage is the same code path; its
torch_geometric GCN pretrain is third-party, absent and unpinned in the reference -> not restated (parity unpinned).

Restates (paths relative to /root/reference):
  recbole/model/fair_recommender/fairgo_pmf.py:100-127  get_norm_rating_matrix: L = D^-1 A, diag = rowsum + 1e-7 (float32)
  fairgo_pmf.py:159-171   forward: fine-tune = sum of the selected filters over ALL rows / total number of filters
  fairgo_pmf.py:173-188   calculate_loss = MSE(dot, rating) [- fair_weight * dis_loss in the fine-tune stage]
  fairgo_pmf.py:190-236   calculate_dis_loss: n_layers x sparse.mm, WAP / LBA / LVA aggregation, BCE / CE heads (the
                          multi-class ego-network head goes through a sigmoid first, line 231)
  fairgo_pmf.py:238-257   predict / full_sort_predict: clamp(., 0, max_rating) / max_rating
  recbole/model/layers.py:58-70  MLPLayers without BatchNorm: Linear -> activation after every layer
State layout: `base.<state_dict key>` (user/item embedding tables, aggr_layer.{0,2,4}), `filter_<attr>.…`, `dis_<attr>.…`.
"""
import numpy as np
import scipy.sparse as sp
import torch
import torch.nn.functional as F

ATTRS = ["gender", "age"]


def norm_matrix(train_u, train_i, train_r, n_users, n_items):
    """fairgo_pmf.py:100-127 as a torch sparse COO tensor (float32)"""
    n = n_users + n_items
    row = np.concatenate([train_u, train_i + n_users])
    col = np.concatenate([train_i + n_users, train_u])
    val = np.concatenate([train_r, train_r]).astype(np.float32)
    A = sp.csr_matrix((val, (row, col)), shape=(n, n), dtype=np.float32)
    diag = (np.asarray(A.sum(axis=1)).ravel().astype(np.float32) + np.float32(1e-7)).astype(np.float32)
    L = sp.coo_matrix(sp.diags((np.float32(1.0) / diag).astype(np.float32)).astype(np.float32) @ A)
    return L


def to_torch_sparse(L, dtype=torch.float32):
    idx = torch.from_numpy(np.stack([L.row, L.col]).astype(np.int64))
    return torch.sparse_coo_tensor(idx, torch.from_numpy(L.data).to(dtype), L.shape).coalesce()


def mlp(x, st, prefix, act=F.leaky_relu):
    """layers.py:58-70 (no BatchNorm, dropout 0): module indices Dropout 3l, Linear 3l+1, activation 3l+2"""
    n = len([k for k in st if k.startswith(prefix + ".mlp_layers.") and k.endswith(".weight")])
    for l in range(n):
        x = act(F.linear(x, st[f"{prefix}.mlp_layers.{3 * l + 1}.weight"], st[f"{prefix}.mlp_layers.{3 * l + 1}.bias"]))
    return x


def forward(st, stage, sst_list, n_users):
    """fairgo_pmf.py:159-171"""
    all_e = torch.cat([st["base.user_embedding_layer.weight"], st["base.item_embedding_layer.weight"]], dim=0)
    if stage == "finetune":
        temp = None
        for s in sst_list:
            f = mlp(all_e, st, f"filter_{s}")
            temp = f if temp is None else temp + f
        all_e = temp / len(ATTRS)
    return all_e[:n_users], all_e[n_users:]


def dis_loss(st, L, uid, labels, sst_list, sst_size, n_users, n_layers, aggr, vs_weights):
    """fairgo_pmf.py:190-236"""
    ua, ia = forward(st, "finetune", sst_list, n_users)
    node_e = ua[uid]
    all_e = torch.cat([ua, ia], dim=0)
    graph = []
    for _ in range(n_layers):
        all_e = torch.sparse.mm(L, all_e)
        graph.append(all_e)
    lva = aggr == "LVA" and n_layers > 1
    if n_layers == 1:
        g = graph[0]
    elif aggr == "WAP":
        g = torch.stack(graph, dim=1).mean(dim=1)
    elif aggr == "LBA":
        x = torch.cat(graph, dim=1)
        x = F.leaky_relu(F.linear(x, st["base.aggr_layer.0.weight"], st["base.aggr_layer.0.bias"]))
        x = F.leaky_relu(F.linear(x, st["base.aggr_layer.2.weight"], st["base.aggr_layer.2.bias"]))
        g = F.linear(x, st["base.aggr_layer.4.weight"], st["base.aggr_layer.4.bias"])
    else:
        g = [e[:n_users][uid] for e in graph]
    if not lva:
        local_e = g[:n_users][uid]
    node, local = 0.0, 0.0
    for s in sst_list:
        if sst_size[s] == 2:
            y = labels[s].to(node_e.dtype).unsqueeze(1)
            node = node + F.binary_cross_entropy(torch.sigmoid(mlp(node_e, st, f"dis_{s}")), y)
            if lva:
                for k, w in enumerate(vs_weights):
                    local = local + w * F.binary_cross_entropy(torch.sigmoid(mlp(g[k], st, f"dis_{s}")), y)
            else:
                local = local + F.binary_cross_entropy(torch.sigmoid(mlp(local_e, st, f"dis_{s}")), y)
        else:
            y = labels[s].long()
            node = node + F.cross_entropy(mlp(node_e, st, f"dis_{s}"), y)
            if lva:
                for k, w in enumerate(vs_weights):
                    local = local + w * F.cross_entropy(torch.sigmoid(mlp(g[k], st, f"dis_{s}")), y)
            else:
                local = local + F.cross_entropy(torch.sigmoid(mlp(local_e, st, f"dis_{s}")), y)
    return node + local


def calculate_loss(st, L, stage, uid, iid, rating, labels, sst_list, sst_size, n_users, n_layers, aggr, vs_weights,
                   fair_weight):
    """fairgo_pmf.py:173-188"""
    ua, ia = forward(st, stage, sst_list, n_users)
    mse = F.mse_loss((ua[uid] * ia[iid]).sum(-1), rating.to(ua.dtype))
    if stage == "finetune":
        return mse - fair_weight * dis_loss(st, L, uid, labels, sst_list, sst_size, n_users, n_layers, aggr, vs_weights)
    return mse


def predict(st, uid, iid, n_users, max_rating=5.0):
    """fairgo_pmf.py:238-248 (fine-tune stage: all filters)"""
    ua, ia = forward(st, "finetune", ATTRS, n_users)
    return torch.clamp((ua[uid] * ia[iid]).sum(1), 0.0, max_rating) / max_rating


def full_sort_predict(st, users, n_users, max_rating=5.0):
    """fairgo_pmf.py:250-257"""
    ua, ia = forward(st, "finetune", ATTRS, n_users)
    return torch.clamp((ua[users] @ ia.t()).view(-1), 0.0, max_rating) / max_rating


def load_state(g, tag, dtype=torch.float32):
    st = {}
    for k in g.files:
        if k.endswith("@" + tag) and not k.startswith("grad_"):
            v = torch.from_numpy(np.array(g[k]))
            st[k[: -len(tag) - 1]] = v.to(dtype) if v.is_floating_point() else v
    return st


def replay(g, lr=1e-3, wd=1e-4, dtype=torch.float32):
    """Re-run the fixture's schedule (gen_golden.run_fairgo).  Returns (losses, filter grads of the first fine-tune
    step, predict / full-sort at that point, pretrained state, final state)."""
    st = load_state(g, "init", dtype)
    n_users, n_items, n_layers, aggr = int(g["n_users"]), int(g["n_items"]), int(g["n_layers"]), str(g["aggr"])
    feats = {"gender": np.array(g["gender"]), "age": np.array(g["age"])}
    sst_size = {s: len(np.unique(feats[s][1:])) for s in ATTRS}
    w = np.array(g["vs_weights"], np.float32)
    vs = [float(x) for x in (w / w.sum(dtype=np.float32))]
    L = to_torch_sparse(norm_matrix(np.array(g["train_u"]), np.array(g["train_i"]), np.array(g["train_r"]), n_users,
                                    n_items), dtype)
    for v in st.values():
        if v.is_floating_point():
            v.requires_grad_(True)
    emb = ["base.user_embedding_layer.weight", "base.item_embedding_layer.weight"]
    dkeys = [k for k in st if k.startswith("dis_")] + ([k for k in st if k.startswith("base.aggr_layer")] if aggr == "LBA" else [])
    fkeys = [k for k in st if k.startswith("filter_")]
    opt_p = torch.optim.Adam([st[k] for k in emb], lr=lr, weight_decay=wd)
    opt_d = torch.optim.Adam([st[k] for k in dkeys], lr=lr, weight_decay=wd)
    opt_f = torch.optim.Adam([st[k] for k in fkeys], lr=lr, weight_decay=wd)
    P, fw = int(g["pretrain_steps"]), float(g["fair_weight"])
    losses, grads, extra, pretrained = [], {}, {}, None
    for s in range(P + 2 * int(g["n_rounds"])):
        uid = torch.from_numpy(np.array(g[f"user_id{s}"]))
        iid = torch.from_numpy(np.array(g[f"item_id{s}"]))
        rating = torch.from_numpy(np.array(g[f"rating{s}"]))
        labels = {a: torch.from_numpy(feats[a][uid.numpy()]) for a in ATTRS}
        sst_list = [str(x) for x in g[f"sst_list{s}"]]
        if s < P:
            opt, stage, full = opt_p, "pretrain", True
        else:
            if s == P:
                pretrained = {k: v.detach().clone().numpy() for k, v in st.items()}
            stage = "finetune"
            opt, full = (opt_f, True) if (s - P) % 2 == 0 else (opt_d, False)
        opt.zero_grad()
        if full:
            loss = calculate_loss(st, L, stage, uid, iid, rating, labels, sst_list, sst_size, n_users, n_layers, aggr, vs, fw)
        else:
            loss = dis_loss(st, L, uid, labels, sst_list, sst_size, n_users, n_layers, aggr, vs)
        loss.backward()
        if s == P:
            grads = {k: st[k].grad.detach().numpy().copy() for k in fkeys if st[k].grad is not None}
            with torch.no_grad():
                extra["predict"] = predict(st, uid, iid, n_users).numpy().copy()
                extra["full_sort"] = full_sort_predict(st, torch.tensor([1, 2, 5]), n_users).numpy().copy()
        opt.step()
        losses.append(loss.item())
    return np.array(losses, np.float32), grads, extra, pretrained, {k: v.detach().numpy() for k, v in st.items()}


def gcn_edges(train_u, train_i, train_r, n_users, n_items):
    """the edge list fairgo_gcn.py:60-66 hands to the GCN: (u -> n_users + i) and (n_users + i -> u), weight = rating"""
    u = torch.as_tensor(np.asarray(train_u), dtype=torch.long)
    i = torch.as_tensor(np.asarray(train_i), dtype=torch.long) + n_users
    w = torch.as_tensor(np.asarray(train_r, np.float32))
    return torch.stack([torch.cat([u, i]), torch.cat([i, u])]), torch.cat([w, w])


def gcn_forward(x, edge_index, edge_weight, weights, biases, act=F.relu, dropout_masks=None):
    """torch_geometric.nn.GCN (BasicGNN over GCNConv; third-party, absent here -- restated from its published algorithm,
    Kipf & Welling 2017): per layer, self loops of weight 1 are added, deg_i = sum of the weights of the edges INTO i,
    norm_e = deg^-1/2[src] * w_e * deg^-1/2[dst]; out_i = sum_{e: dst = i} norm_e * (x W^T)[src] + b; between layers act
    then dropout (`dropout_masks[k]`, already scaled, multiplies the output of layer k), nothing after the last layer.
    PARITY UNPINNED against torch_geometric itself."""
    N = x.shape[0]
    loops = torch.arange(N)
    src = torch.cat([edge_index[0], loops])
    dst = torch.cat([edge_index[1], loops])
    w = torch.cat([edge_weight.to(x.dtype), torch.ones(N, dtype=x.dtype)])
    deg = torch.zeros(N, dtype=x.dtype).index_add_(0, dst, w)
    dinv = deg.pow(-0.5)
    dinv[torch.isinf(dinv)] = 0
    norm = dinv[src] * w * dinv[dst]
    for k, (W, b) in enumerate(zip(weights, biases)):
        h = x @ W.t()
        x = torch.zeros(N, W.shape[0], dtype=x.dtype).index_add_(0, dst, norm[:, None] * h[src]) + b
        if k + 1 < len(weights):
            x = act(x)
            if dropout_masks is not None:
                x = x * dropout_masks[k]
    return x
