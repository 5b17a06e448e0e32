"""CPU restatement (stock torch CPU fp32 ops, functional, no reference import) of the PFCN_MLP training math:
filter MLPs -> NCF-style scorer -> BPR, discriminator MLPs -> BCE / CrossEntropy.
TEST INFRASTRUCTURE ONLY (tests/, smoke, bench cpu legs) -- never imported by the product package.

Parity status: PINNED against tests/golden/pfcn_mlp_{sm,cm}.npz (generated from the unmodified reference by
oracle/gen_golden.py `pfcn`) in tests/test_oracle_golden.py.  A floating-point path: per the tier rules this oracle is
a torch fp32 reference (autograd supplies the backward of the restated forward).

Restates (paths relative to /root/reference):
  recbole/model/layers.py:58-70                      MLPLayers: Dropout -> Linear -> [BatchNorm1d] -> activation per layer
  recbole/model/loss.py:44-46                        BPRLoss: -mean(log(1e-10 + sigmoid(pos - neg)))
  recbole/model/fair_recommender/pfcn_mlp.py:145-167 forward: sm = one filter per attribute subset (index sum 2^attr);
                                                     cm = sum of single-attribute filters / TOTAL filter count
  pfcn_mlp.py:177-193                                calculate_loss = bpr - dis_weight * dis
  pfcn_mlp.py:195-211                                calculate_dis_loss: binary -> BCE(sigmoid(z), y), else CE(z, y.long())
State layout: dict name -> torch tensor with the reference's state_dict names prefixed by the owner
(`user_embedding`, `item_embedding`, `mlp_layer.mlp_layers.<k>.weight`, `filter_<idx>.…`, `dis_<attr>.…`).
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


def filtered_user(st, uid, sst_list, sst_dict, filter_mode, n_filters, act, training=True):
    """pfcn_mlp.py:145-167"""
    ue = st["user_embedding"][uid]
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


def score(st, ue, ie):
    """pfcn_mlp.py:63,172: tower on [user || item], ReLU after every layer including the last"""
    return mlp_forward(torch.cat((ue, ie), dim=1), st, "mlp_layer", n_layers_of(st, "mlp_layer"), False, "relu")


def dis_loss(st, uid, labels, sst_list, sst_dict, sst_size, filter_mode, n_filters, act):
    """pfcn_mlp.py:195-211"""
    ue = filtered_user(st, uid, sst_list, sst_dict, filter_mode, n_filters, act)
    loss = 0.0
    for s in sst_list:
        z = mlp_forward(ue, st, f"dis_{s}", n_layers_of(st, f"dis_{s}"), True, act)
        if sst_size[s] == 2:
            loss = loss + F.binary_cross_entropy(torch.sigmoid(z), labels[s].to(z.dtype).view(-1, 1))
        else:
            loss = loss + F.cross_entropy(z, labels[s].long())
    return loss


def calculate_loss(st, uid, pos, neg, labels, sst_list, sst_dict, sst_size, filter_mode, n_filters, act, dis_weight):
    """pfcn_mlp.py:177-193"""
    ue = filtered_user(st, uid, sst_list, sst_dict, filter_mode, n_filters, act)
    p = score(st, ue, st["item_embedding"][pos])
    n = score(st, ue, st["item_embedding"][neg])
    bpr = -torch.log(1e-10 + torch.sigmoid(p - n)).mean()
    if filter_mode == "none":
        return bpr
    return bpr - dis_weight * dis_loss(st, uid, labels, sst_list, sst_dict, sst_size, filter_mode, n_filters, act)


def predict(st, uid, iid, sst_list, sst_dict, filter_mode, n_filters, act):
    """pfcn_mlp.py:169-175 (training-mode batch statistics when called on a model in train mode)"""
    ue = filtered_user(st, uid, sst_list, sst_dict, filter_mode, n_filters, act)
    return torch.sigmoid(score(st, ue, st["item_embedding"][iid]))


# ---------------------------------------------------------------- fixture replay helpers
def load_state(g, tag):
    st = {}
    for k in g.files:
        if k.endswith("@" + tag) and not k.startswith("grad_"):
            st[k[: -len(tag) - 1]] = torch.from_numpy(np.array(g[k]))
    return st


def param_groups(st):
    """(filter-optimizer params, discriminator-optimizer params) as in PFCN_MLPTrainer (trainer.py:1189-1198): base =
    embeddings + scorer; filters join the base; discriminators apart.  BatchNorm buffers are not parameters."""
    def is_param(k):
        return not (k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"))
    base = [k for k in st if is_param(k) and (k in ("user_embedding", "item_embedding") or k.startswith("mlp_layer."))]
    filt = [k for k in st if is_param(k) and k.startswith("filter_")]
    dis = [k for k in st if is_param(k) and k.startswith("dis_")]
    return base + filt, dis


def replay(g, lr=1e-3, wd=1e-4, dis_weight=10.0, act="leakyrelu"):
    """Re-run the fixture's alternating schedule (gen_golden.run_pfcn_mlp) on the restatement.
    Returns (losses, grads of step 0, final state)."""
    st = load_state(g, "init")
    fkeys, dkeys = param_groups(st)
    for k in fkeys + dkeys:
        st[k].requires_grad_(True)
    filter_mode = str(g["filter_mode"])
    attrs = ["gender", "age"]
    if filter_mode == "sm":
        sst_dict, n_filters = {s: 2 ** i for i, s in enumerate(attrs)}, 2 ** len(attrs) - 1
    else:
        sst_dict, n_filters = {s: i + 1 for i, s in enumerate(attrs)}, len(attrs)
    feats = {"gender": np.array(g["gender"]), "age": np.array(g["age"])}
    sst_size = {s: len(np.unique(feats[s][1:])) for s in attrs}
    opt_f = torch.optim.Adam([st[k] for k in fkeys], lr=lr, weight_decay=wd)
    opt_d = torch.optim.Adam([st[k] for k in dkeys], lr=lr, weight_decay=wd)
    losses, grads0 = [], {}
    for s in range(2 * int(g["n_rounds"])):
        uid = torch.from_numpy(np.array(g[f"user_id{s}"]))
        labels = {a: torch.from_numpy(feats[a][uid.numpy()]) for a in attrs}
        sst_list = [str(x) for x in g[f"sst_list{s}"]]
        opt = opt_f if s % 2 == 0 else opt_d
        opt.zero_grad()
        if s % 2 == 0:
            loss = calculate_loss(st, uid, torch.from_numpy(np.array(g[f"item_id{s}"])),
                                  torch.from_numpy(np.array(g[f"neg_item_id{s}"])), labels, sst_list, sst_dict,
                                  sst_size, filter_mode, n_filters, act, dis_weight)
        else:
            loss = dis_loss(st, uid, labels, sst_list, sst_dict, sst_size, filter_mode, n_filters, act)
        loss.backward()
        if s == 0:
            grads0 = {k: st[k].grad.detach().numpy().copy() for k in fkeys if st[k].grad is not None}
        opt.step()
        losses.append(loss.item())
    return np.array(losses, np.float32), grads0, {k: v.detach().numpy() for k, v in st.items()}
