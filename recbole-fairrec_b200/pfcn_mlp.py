"""PFCN_MLP (personalised fairness filters + discriminators over an NCF-style scorer) -- drop-in for
recbole/model/fair_recommender/pfcn_mlp.py:23-232, computed by this package's kernels (layers.MLPLayers, ops.*).

Kept from the reference, on purpose: the filter / discriminator MLPs live in plain Python dicts (`filter_layer`,
`dis_layer_dict`; pfcn_mlp.py:111-143) -- so, like there, they are not part of `state_dict()`; `sm` mode owns one
filter per non-empty attribute subset (index = sum of 2^attr, pfcn_mlp.py:74-78,152-157); `cm` mode averages the
single-attribute filters over the TOTAL number of filters (pfcn_mlp.py:158-165); binary attributes get one logit +
BCE on float 0/1 labels, the others CrossEntropy on `.long()` labels (pfcn_mlp.py:203-209).
`full_sort_predict` of the reference is broken for this model (SURVEY.md section 7 hard part 7): here it raises
NotImplementedError, which is the trainer's cue to fall back to `predict` over all items (trainer.py:425-433).
"""
import numpy as np
import torch
import torch.nn as nn

from . import ops
from .layers import MLPLayers


class PFCN_MLP(nn.Module):
    input_type = "PAIRWISE"
    type = "GENERAL"

    def __init__(self, config, dataset):
        super().__init__()
        self.USER_ID, self.ITEM_ID = config["USER_ID_FIELD"], config["ITEM_ID_FIELD"]
        self.POS_ITEM_ID = self.ITEM_ID
        self.NEG_ITEM_ID = config["NEG_PREFIX"] + self.ITEM_ID
        self.n_users, self.n_items = dataset.num(self.USER_ID), dataset.num(self.ITEM_ID)
        self.device = config["device"]
        self.filter_mode = config["filter_mode"].lower()
        if self.filter_mode not in ("cm", "sm", "none"):
            raise AssertionError("filter_mode must be cm, sm or none")
        self.sst_attrs = list(config["sst_attr_list"])
        self.embedding_size = config["embedding_size"]
        if self.filter_mode != "none":
            self.dis_drop_out = config["dis_dropout"]
            self.dis_weight = config["dis_weight"]
            self.dis_hidden_size_list = list(config["dis_hidden_size_list"])
        self.activation = config["activation"]
        self.dropout = config["dropout"]
        self.mlp_hidden_size_list = list(config["mlp_hidden_size_list"])
        self.filter_num, self.sst_dict = self._get_filter_info()
        self.sst_size = self._get_sst_size(dataset.get_user_feature())
        if self.filter_mode != "none":
            self.filter_layer = self.init_filter()
            self.dis_layer_dict = self.init_dis_layer()
        self.user_embedding = nn.Embedding(self.n_users, self.embedding_size)
        self.item_embedding = nn.Embedding(self.n_items, self.embedding_size)
        self.mlp_layer = MLPLayers([self.embedding_size * 2] + self.mlp_hidden_size_list + [1], dropout=self.dropout)

    # ------------------------------------------------------------------ construction (pfcn_mlp.py:67-143)
    def _get_filter_info(self):
        if self.filter_mode == "cm":
            return len(self.sst_attrs), {s: i + 1 for i, s in enumerate(self.sst_attrs)}
        if self.filter_mode == "sm":
            return 2 ** len(self.sst_attrs) - 1, {s: int(2 ** i) for i, s in enumerate(self.sst_attrs)}
        return 0, {}

    def _get_sst_size(self, user_feature):
        size = {}
        for sst in self.sst_attrs:
            if sst not in user_feature.columns:
                raise ValueError(f"{sst} sensitive attribute not in user feature")
            size[sst] = len(user_feature[sst][1:].unique())
        return size

    def init_filter(self):
        e = self.embedding_size
        return {i + 1: MLPLayers([e, e * 2, e], activation=self.activation, bn=True, init_method="norm").to(self.device)
                for i in range(self.filter_num)}

    def init_dis_layer(self):
        out = {}
        for sst in self.sst_attrs:
            c = self.sst_size[sst]
            out[sst] = MLPLayers([self.embedding_size] + self.dis_hidden_size_list + [1 if c == 2 else c],
                                 dropout=self.dis_drop_out, activation=self.activation, bn=True,
                                 init_method="norm").to(self.device)
        return out

    def to(self, *a, **k):           # the dict-held sub-networks follow the model (the reference leaves them behind)
        super().to(*a, **k)
        if self.filter_mode != "none":
            for m in list(self.filter_layer.values()) + list(self.dis_layer_dict.values()):
                m.to(*a, **k)
        return self

    def train(self, mode=True):
        super().train(mode)
        if self.filter_mode != "none":
            for m in list(self.filter_layer.values()) + list(self.dis_layer_dict.values()):
                m.train(mode)
        return self

    def other_parameter(self):
        return dict()

    def load_other_parameter(self, para):
        return

    # ------------------------------------------------------------------ forward pieces
    def _ids(self, t):
        return t.to(device=self.user_embedding.weight.device, dtype=torch.int32).contiguous()

    def forward(self, user, item=None, sst_list=None):
        """pfcn_mlp.py:145-167"""
        user_embed = ops.GatherRows.apply(self.user_embedding.weight, self._ids(user))
        item_embed = None if item is None else ops.GatherRows.apply(self.item_embedding.weight, self._ids(item))
        if self.filter_mode == "none":
            return user_embed, item_embed
        if self.filter_mode == "sm":
            idx = sum(self.sst_dict[s] for s in sst_list)
            return self.filter_layer[idx](user_embed), item_embed
        acc = None
        for s in sst_list:
            e = self.filter_layer[self.sst_dict[s]](user_embed)
            acc = e if acc is None else acc + e
        return acc / len(self.filter_layer), item_embed

    def _score(self, user_embed, item_embed):
        return self.mlp_layer(torch.cat((user_embed, item_embed), dim=1))

    def predict(self, interaction, sst_list=None):
        """pfcn_mlp.py:169-175"""
        u, i = self.forward(interaction[self.USER_ID], interaction[self.ITEM_ID], sst_list)
        return torch.sigmoid(self._score(u, i))

    def calculate_loss(self, interaction, sst_list=None):
        """pfcn_mlp.py:177-193: BPR(pos, neg) - dis_weight * discriminator loss"""
        user_embed, pos_embed = self.forward(interaction[self.USER_ID], interaction[self.POS_ITEM_ID], sst_list)
        neg_embed = ops.GatherRows.apply(self.item_embedding.weight, self._ids(interaction[self.NEG_ITEM_ID]))
        bpr = ops.BprLoss.apply(self._score(user_embed, pos_embed), self._score(user_embed, neg_embed))
        if self.filter_mode != "none":
            return bpr - self.dis_weight * self.calculate_dis_loss(interaction, sst_list)
        return bpr

    def calculate_dis_loss(self, interaction, sst_list=None):
        """pfcn_mlp.py:195-211"""
        user_embed, _ = self.forward(interaction[self.USER_ID], None, sst_list)
        dev = user_embed.device
        loss = 0.0
        for sst in sst_list:
            z = self.dis_layer_dict[sst](user_embed)
            if self.sst_size[sst] == 2:
                loss = loss + ops.SigmoidBce.apply(z, interaction[sst].to(device=dev, dtype=torch.float32))
            else:
                loss = loss + ops.SoftmaxCe.apply(z, interaction[sst].to(device=dev, dtype=torch.int32))
        return loss

    def full_sort_predict(self, interaction, sst_list=None):
        raise NotImplementedError("PFCN full-sort scoring is undefined in the reference (pfcn_mlp.py:213-222 passes a "
                                  "tuple where a tensor is expected); use predict() over all items")

    def get_sst_embed(self, user_data, sst_list=None):
        """pfcn_mlp.py:224-232"""
        ret = {}
        user_indices = torch.arange(1, self.n_users)
        sst_list = self.sst_attrs if self.filter_mode == "none" else sst_list
        for sst in sst_list:
            ret[sst] = user_data[sst][user_indices - 1]
        ret["embedding"], _ = self.forward(user_indices.to(self.device), None, sst_list)
        return ret


class PFCN_MLPTrainer:
    """The alternating schedule of PFCNTrainer / PFCN_MLPTrainer (trainer.py:865-898, 1189-1198): per epoch a random
    non-empty attribute subset; every `train_epoch_interval`-th epoch one pass on `bpr - dis_weight * dis` with the
    filter optimizer (embeddings + filters + scorer), then always one pass on `dis` with the discriminator optimizer."""

    def __init__(self, config, model):
        import torch.optim as optim
        self.config, self.model = config, model
        self.filter_mode = config["filter_mode"].lower()
        self.train_epoch_interval = config["train_epoch_interval"] or 1
        lr, wd = config["learning_rate"], config["weight_decay"] or 0.0
        base = [model.user_embedding.weight, model.item_embedding.weight] + list(model.mlp_layer.parameters())
        if self.filter_mode != "none":
            self.sst_attrs = list(config["sst_attr_list"])
            fparams = [p for f in model.filter_layer.values() for p in f.parameters()]
            dparams = [p for d in model.dis_layer_dict.values() for p in d.parameters()]
            self.optimizer_filter = optim.Adam(base + fparams, lr=lr, weight_decay=wd)
            self.optimizer_dis = optim.Adam(dparams, lr=lr, weight_decay=wd)
        else:
            self.optimizer_filter = optim.Adam(base, lr=lr, weight_decay=wd)

    def _pass(self, train_data, loss_func, optimizer, sst_list):
        self.model.train()
        total = None
        for interaction in train_data:
            optimizer.zero_grad()
            loss = loss_func(interaction, sst_list)
            v = loss.item()
            if v != v:
                raise ValueError("Training loss is nan")
            total = v if total is None else total + v
            loss.backward()
            optimizer.step()
        return total

    def _train_epoch(self, train_data, epoch_idx):
        if self.filter_mode == "none":
            return self._pass(train_data, self.model.calculate_loss, self.optimizer_filter, None)
        mask = np.zeros(len(self.sst_attrs))
        while mask.sum() == 0:
            mask = np.random.choice([0, 1], len(self.sst_attrs))
        sst_list = [s for s, m in zip(self.sst_attrs, mask) if m != 0]
        filter_loss = 0.0
        if epoch_idx % self.train_epoch_interval == 0:
            filter_loss = self._pass(train_data, self.model.calculate_loss, self.optimizer_filter, sst_list)
        dis_loss = self._pass(train_data, self.model.calculate_dis_loss, self.optimizer_dis, sst_list)
        return filter_loss, dis_loss
