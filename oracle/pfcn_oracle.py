"""CPU restatement (stock torch CPU fp32 ops, functional, no reference import) of the PFCN family's training math:
filter MLPs -> scorer (NCF tower / dot / biased dot / cosine) -> BPR, discriminator MLPs -> BCE / CrossEntropy.
TEST INFRASTRUCTURE ONLY (tests/, smoke, bench cpu legs) -- never imported by the product package.

Parity status: PINNED against tests/golden/pfcn_{mlp,pmf,biasedmf,dmf}_*.npz (generated from the unmodified reference by
oracle/gen_golden.py `pfcn`) in tests/test_oracle_golden.py.  A floating-point path: per the tier rules this oracle is
a torch fp32 reference (autograd supplies the backward of the restated forward).

Restates (paths relative to /root/reference):
  recbole/model/layers.py:58-70                      MLPLayers: Dropout -> Linear -> [BatchNorm1d] -> activation per layer
  recbole/model/loss.py:44-46                        BPRLoss: -mean(log(1e-10 + sigmoid(pos - neg)))
  recbole/model/fair_recommender/pfcn_mlp.py:145-167 forward: sm = one filter per attribute subset (index sum 2^attr);
                                                     cm = sum of single-attribute filters / TOTAL filter count
  pfcn_mlp.py:177-193                                calculate_loss = bpr - dis_weight * dis
  pfcn_mlp.py:195-211                                calculate_dis_loss: binary -> BCE(sigmoid(z), y), else CE(z, y.long())
  pfcn_pmf.py:176-193, pfcn_biasedmf.py:183-199 ([B]+[B,1] broadcast kept), pfcn_dmf.py:144-199   the other scorers
State layout: dict name -> torch tensor: `base.<state_dict key of the model>` for the registered modules
(`base.user_embedding.weight`, `base.mlp_layer.mlp_layers.<k>.weight`, …), `filter_<idx>.…` / `dis_<attr>.…` for the
dict-held MLPs.
"""
import numpy as np
import torch
import torch.nn.functional as F

ACTS = {"relu": F.relu, "leakyrelu": F.leaky_relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh, None: lambda x: x}


def mlp_forward(x, st, prefix, n_layers, bn, act, training=True, momentum=0.1, eps=1e-5):
    """layers.py:58-70 with dropout p = 0.  Module indices follow nn.Sequential: per layer Dropout, Linear, [BN], [act]."""
    stride = 2 + (1 if bn else 0) + (1 if act is not None else 0)
    for l in range(n_layers):
        k = l * stride + 1
        x = F.linear(x, st[f"{prefix}.mlp_layers.{k}.weight"], st[f"{prefix}.mlp_layers.{k}.bias"])
        if bn:
            p = f"{prefix}.mlp_layers.{k + 1}"
            x = F.batch_norm(x, st[p + ".running_mean"], st[p + ".running_var"], st[p + ".weight"], st[p + ".bias"],
                             training, momentum, eps)
        x = ACTS[act](x)
    return x


def n_layers_of(st, prefix):
    return len([k for k in st if k.startswith(prefix + ".mlp_layers.") and k.endswith(".weight")
                and st[k].dim() == 2])


def user_base(st, model, uid, act):
    ue = st["base." + EMB[model][0] + ".weight"][uid]
    if model == "PFCN_DMF":                       # pfcn_dmf.py:146-147
        ue = mlp_forward(ue, st, "base.user_mlp", n_layers_of(st, "base.user_mlp"), False, act)
    return ue


def item_base(st, model, iid, act):
    ie = st["base." + EMB[model][1] + ".weight"][iid]
    if model == "PFCN_DMF":                       # pfcn_dmf.py:150-151
        ie = mlp_forward(ie, st, "base.item_mlp", n_layers_of(st, "base.item_mlp"), False, act)
    return ie


EMB = {"PFCN_MLP": ("user_embedding", "item_embedding"), "PFCN_PMF": ("user_embedding_layer", "item_embedding_layer"),
       "PFCN_BiasedMF": ("user_embedding_layer", "item_embedding_layer"),
       "PFCN_DMF": ("user_embedding_layer", "item_embedding_layer")}


def filtered_user(st, model, uid, sst_list, sst_dict, filter_mode, n_filters, act, training=True):
    """pfcn_mlp.py:145-167 (identical in the four files)"""
    ue = user_base(st, model, uid, act)
    if filter_mode == "none":
        return ue
    if filter_mode == "sm":
        idx = sum(sst_dict[s] for s in sst_list)
        return mlp_forward(ue, st, f"filter_{idx}", 2, True, act, training)
    acc = None
    for s in sst_list:
        e = mlp_forward(ue, st, f"filter_{sst_dict[s]}", 2, True, act, training)
        acc = e if acc is None else acc + e
    return acc / n_filters


def tower(st, ue, ie):
    """pfcn_mlp.py:63,172: tower on [user || item], ReLU after every layer including the last"""
    return mlp_forward(torch.cat((ue, ie), dim=1), st, "base.mlp_layer", n_layers_of(st, "base.mlp_layer"), False, "relu")


def dis_loss(st, model, uid, labels, sst_list, sst_dict, sst_size, filter_mode, n_filters, act):
    """pfcn_mlp.py:195-211"""
    ue = filtered_user(st, model, uid, sst_list, sst_dict, filter_mode, n_filters, act)
    loss = 0.0
    for s in sst_list:
        z = mlp_forward(ue, st, f"dis_{s}", n_layers_of(st, f"dis_{s}"), True, act)
        if sst_size[s] == 2:
            loss = loss + F.binary_cross_entropy(torch.sigmoid(z), labels[s].to(z.dtype).view(-1, 1))
        else:
            loss = loss + F.cross_entropy(z, labels[s].long())
    return loss


def bpr(pos, neg):
    """loss.py:44-46"""
    return -torch.log(1e-10 + torch.sigmoid(pos - neg)).mean()


def calculate_loss(st, model, uid, pos, neg, labels, sst_list, sst_dict, sst_size, filter_mode, n_filters, act, dis_weight):
    """pfcn_mlp.py:177-193 / pfcn_pmf.py:176-193 / pfcn_biasedmf.py:183-199 / pfcn_dmf.py:180-199"""
    ue = filtered_user(st, model, uid, sst_list, sst_dict, filter_mode, n_filters, act)
    pe, ne = item_base(st, model, pos, act), item_base(st, model, neg, act)
    if model == "PFCN_MLP":
        loss = bpr(tower(st, ue, pe), tower(st, ue, ne))
    elif model == "PFCN_PMF":
        loss = bpr((ue * pe).sum(-1), (ue * ne).sum(-1))
    elif model == "PFCN_BiasedMF":
        ub, gb = st["base.user_bias.weight"][uid], st["base.global_bias"]
        pib, nib = st["base.item_bias.weight"][pos], st["base.item_bias.weight"][neg]
        # [B] + [B,1] broadcasts to [B,B] (pfcn_biasedmf.py:189-190), kept as in the reference
        loss = bpr((ue * pe).sum(-1) + ub + pib + gb, (ue * ne).sum(-1) + ub + nib + gb)
    else:
        loss = bpr(F.cosine_similarity(ue, pe) * 10, F.cosine_similarity(ue, ne) * 10)
    if filter_mode == "none":
        return loss
    return loss - dis_weight * dis_loss(st, model, uid, labels, sst_list, sst_dict, sst_size, filter_mode, n_filters, act)


def predict(st, model, uid, iid, sst_list, sst_dict, filter_mode, n_filters, act):
    """pfcn_mlp.py:169-175 etc. (training-mode batch statistics when called on a model in train mode)"""
    ue = filtered_user(st, model, uid, sst_list, sst_dict, filter_mode, n_filters, act)
    ie = item_base(st, model, iid, act)
    if model == "PFCN_MLP":
        return torch.sigmoid(tower(st, ue, ie))
    if model == "PFCN_PMF":
        return torch.sigmoid((ue * ie).sum(-1, keepdim=True))
    if model == "PFCN_BiasedMF":
        return torch.sigmoid((ue * ie).sum(-1, keepdim=True) + st["base.user_bias.weight"][uid]
                             + st["base.item_bias.weight"][iid] + st["base.global_bias"])
    return torch.sigmoid(F.cosine_similarity(ue, ie))


# ---------------------------------------------------------------- fixture replay helpers
def load_state(g, tag):
    st = {}
    for k in g.files:
        if k.endswith("@" + tag) and not k.startswith("grad_"):
            st[k[: -len(tag) - 1]] = torch.from_numpy(np.array(g[k]))
    return st


def param_groups(st):
    """(filter-optimizer params, discriminator-optimizer params) as in PFCN_*Trainer (trainer.py:1189-1235): base =
    model.parameters(); filters join the base; discriminators apart.  BatchNorm buffers are not parameters."""
    def is_param(k):
        return not (k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"))
    base = [k for k in st if is_param(k) and k.startswith("base.")]
    filt = [k for k in st if is_param(k) and k.startswith("filter_")]
    dis = [k for k in st if is_param(k) and k.startswith("dis_")]
    return base + filt, dis


def fixture_setup(g):
    filter_mode, attrs = str(g["filter_mode"]), ["gender", "age"]
    if filter_mode == "sm":
        sst_dict, n_filters = {s: 2 ** i for i, s in enumerate(attrs)}, 2 ** len(attrs) - 1
    else:
        sst_dict, n_filters = {s: i + 1 for i, s in enumerate(attrs)}, len(attrs)
    feats = {"gender": np.array(g["gender"]), "age": np.array(g["age"])}
    sst_size = {s: len(np.unique(feats[s][1:])) for s in attrs}
    return filter_mode, attrs, sst_dict, n_filters, feats, sst_size


def replay(g, lr=1e-3, wd=1e-4, dis_weight=10.0, act="leakyrelu", dtype=torch.float32):
    """Re-run the fixture's alternating schedule (gen_golden.run_pfcn) on the restatement.
    Returns (losses, grads of step 0, predict at step 0, final state)."""
    st = load_state(g, "init")
    st = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in st.items()}
    model = str(g["model"])
    fkeys, dkeys = param_groups(st)
    for k in fkeys + dkeys:
        st[k].requires_grad_(True)
    filter_mode, attrs, sst_dict, n_filters, feats, sst_size = fixture_setup(g)
    opt_f = torch.optim.Adam([st[k] for k in fkeys], lr=lr, weight_decay=wd)
    opt_d = torch.optim.Adam([st[k] for k in dkeys], lr=lr, weight_decay=wd)
    losses, grads0, pred0 = [], {}, None
    for s in range(2 * int(g["n_rounds"])):
        uid = torch.from_numpy(np.array(g[f"user_id{s}"]))
        pos = torch.from_numpy(np.array(g[f"item_id{s}"]))
        labels = {a: torch.from_numpy(feats[a][uid.numpy()]) for a in attrs}
        sst_list = [str(x) for x in g[f"sst_list{s}"]]
        opt = opt_f if s % 2 == 0 else opt_d
        opt.zero_grad()
        if s % 2 == 0:
            loss = calculate_loss(st, model, uid, pos, torch.from_numpy(np.array(g[f"neg_item_id{s}"])), labels, sst_list,
                                  sst_dict, sst_size, filter_mode, n_filters, act, dis_weight)
        else:
            loss = dis_loss(st, model, uid, labels, sst_list, sst_dict, sst_size, filter_mode, n_filters, act)
        loss.backward()
        if s == 0:
            grads0 = {k: st[k].grad.detach().numpy().copy() for k in fkeys if st[k].grad is not None}
            with torch.no_grad():       # the fixture's extra train-mode forward (advances the running statistics too)
                pred0 = predict(st, model, uid, pos, sst_list, sst_dict, filter_mode, n_filters, act).numpy().copy()
        opt.step()
        losses.append(loss.item())
    return np.array(losses, np.float32), grads0, pred0, {k: v.detach().numpy() for k, v in st.items()}
