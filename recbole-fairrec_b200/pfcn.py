"""The PFCN family (personalised counterfactual-fairness filters + adversarial discriminators) -- drop-ins for
recbole/model/fair_recommender/pfcn_{mlp,pmf,biasedmf,dmf}.py, computed by this package's kernels (layers.MLPLayers,
ops.*).  One base class holds what the four reference files repeat verbatim (filter bookkeeping, discriminator loss,
get_sst_embed); each subclass is its scorer.

Kept from the reference, on purpose:
  * filter / discriminator MLPs live in plain Python dicts (`filter_layer`, `dis_layer_dict`; pfcn_mlp.py:111-143), so
    they are not part of `state_dict()`;
  * `sm` mode owns one filter per non-empty attribute subset, index = sum of 2^attr (pfcn_mlp.py:74-78,152-157); `cm`
    mode divides the sum of the selected single-attribute filters by the TOTAL number of filters (pfcn_mlp.py:158-165);
  * binary attributes: one logit + BCE on float 0/1 labels; others: CrossEntropy on `.long()` labels (pfcn_mlp.py:203-209);
  * PFCN_BiasedMF.calculate_loss adds [B] dot products to [B,1] bias columns, so its BPR runs over a [B,B] broadcast
    matrix (pfcn_biasedmf.py:189-192) -- reproduced by ops.BprOuter;
  * `full_sort_predict` is broken in all four reference files (the (user, item) tuple returned by forward is used as a
    tensor, SURVEY.md section 7 hard part 7): here it raises NotImplementedError, the trainer's cue to fall back to
    `predict` over all items (trainer.py:425-433).
"""
import numpy as np
import torch
import torch.nn as nn

from .checkpoint import CheckpointMixin
from .utils import stopping_step as _stopping_step
from . import ops
from .layers import MLPLayers


class _PFCNBase(nn.Module):
    input_type = "PAIRWISE"
    type = "GENERAL"
    USER_EMB, ITEM_EMB = "user_embedding_layer", "item_embedding_layer"   # attribute names differ per reference file
    SCORER_FIRST = True       # pfcn_pmf / pfcn_biasedmf / pfcn_dmf build the scorer before filters and discriminators

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
        self.filter_num, self.sst_dict = self._get_filter_info()
        self.sst_size = self._get_sst_size(dataset.get_user_feature())
        # module creation order = the reference file's, so that a seed gives the reference's initial weights
        # (verified against the live reference for all four models: every tensor bit-identical)
        if self.SCORER_FIRST:
            self._build_scorer(config)
        if self.filter_mode != "none":
            self.filter_layer = self.init_filter()
            self.dis_layer_dict = self.init_dis_layer()
        if not self.SCORER_FIRST:
            self._build_scorer(config)

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

    def _filter_activation(self):
        return self.activation

    def _dis_activation(self):
        return self.activation

    def init_filter(self):
        e = self.embedding_size
        return {i + 1: MLPLayers([e, e * 2, e], activation=self._filter_activation(), bn=True,
                                 init_method="norm").to(self.device) for i in range(self.filter_num)}

    def init_dis_layer(self):
        out = {}
        for sst in self.sst_attrs:
            c = self.sst_size[sst]
            out[sst] = MLPLayers([self.embedding_size] + self.dis_hidden_size_list + [1 if c == 2 else c],
                                 dropout=self.dis_drop_out, activation=self._dis_activation(), bn=True,
                                 init_method="norm").to(self.device)
        return out

    def _dict_modules(self):
        if self.filter_mode == "none":
            return []
        return list(self.filter_layer.values()) + list(self.dis_layer_dict.values())

    def to(self, *a, **k):           # the dict-held sub-networks follow the model (the reference leaves them behind)
        super().to(*a, **k)
        for m in self._dict_modules():
            m.to(*a, **k)
        return self

    def train(self, mode=True):
        super().train(mode)
        for m in self._dict_modules():
            m.train(mode)
        return self

    def other_parameter(self):
        return dict()

    def load_other_parameter(self, para):
        return

    # ------------------------------------------------------------------ shared forward pieces
    def _ids(self, t):
        return t.to(device=getattr(self, self.USER_EMB).weight.device, dtype=torch.int32).contiguous()

    def _user_base(self, user):
        return ops.GatherRows.apply(getattr(self, self.USER_EMB).weight, self._ids(user))

    def _item_base(self, item):
        return ops.GatherRows.apply(getattr(self, self.ITEM_EMB).weight, self._ids(item))

    def _apply_filters(self, user_embed, sst_list):
        if self.filter_mode == "none":
            return user_embed
        if self.filter_mode == "sm":
            return self.filter_layer[sum(self.sst_dict[s] for s in sst_list)](user_embed)
        outs = [self.filter_layer[self.sst_dict[s]](user_embed) for s in sst_list]
        return ops.SumDiv.apply(len(self.filter_layer), *outs)

    def forward(self, user, item=None, sst_list=None):
        """pfcn_mlp.py:145-167"""
        user_embed = self._apply_filters(self._user_base(user), sst_list)
        return user_embed, (None if item is None else self._item_base(item))

    # The reference's loss evaluates forward(user) TWICE on the same batch (pfcn_mlp.py:177-193: once for the scorer, once more
    # inside calculate_dis_loss).  Where the user path is deterministic in training mode (no dropout: the filters never have
    # one) the two passes return the same tensor and differ only in that the BatchNorm buffers advance twice -- so one pass
    # with ops.bn_repeat(2) and the tensor shared by both consumers is the same computation (autograd adds the two
    # gradients at the shared output instead of after two backward passes: one filter forward, one filter backward, one
    # embedding gather and one gradient scatter less per step).  Needs the fused chain kernels (they own the buffer
    # update); anything else keeps the second evaluation.
    SHARE_USER_FORWARD = True

    def _user_path_dropout(self):
        return 0.0

    def _filters_of(self, sst_list):
        if self.filter_mode == "none":
            return []
        if self.filter_mode == "sm":
            return [self.filter_layer[sum(self.sst_dict[s] for s in sst_list)]]
        return [self.filter_layer[self.sst_dict[s]] for s in sst_list]

    def _forward_for_loss(self, user, item, sst_list):
        """forward() for calculate_loss: -> (user_embed, item_embed, shared) with shared = the filtered embedding also
        stands for the loss's second evaluation (the BatchNorm buffers have advanced twice already)"""
        share = False
        if self.SHARE_USER_FORWARD and self.filter_mode != "none" and self.training and self._user_path_dropout() == 0.0:
            base = self._user_base(user)
            mods = self._filters_of(sst_list)
            if all(m.dropout == 0.0 and ops.mlp_chain_would_fuse([m], [base]) for m in mods):
                with ops.bn_repeat(2):
                    user_embed = self._apply_filters(base, sst_list)
                return user_embed, (None if item is None else self._item_base(item)), True
            return self._apply_filters(base, sst_list), (None if item is None else self._item_base(item)), False
        user_embed, item_embed = self.forward(user, item, sst_list)
        return user_embed, item_embed, share

    def calculate_dis_loss(self, interaction, sst_list=None):
        """pfcn_mlp.py:195-211.  Called on its own (the discriminator phase, trainer.py:889-892) only the discriminators
        are optimised, so the filtered embedding enters as a constant: the reference back-propagates into the filters
        and embeddings here too, but no optimizer step ever uses those gradients (they are zeroed by the filter
        optimizer before its own backward).  `calculate_loss` uses _dis_terms with the graph intact."""
        with torch.no_grad():
            user_embed, _ = self.forward(interaction[self.USER_ID], None, sst_list)
        return self._dis_terms(user_embed, interaction, sst_list)

    def _dis_terms(self, user_embed, interaction, sst_list):
        dev = user_embed.device
        loss = 0.0
        # the discriminators of the step share one forward and one backward launch (ops.mlp_chain); None = per-module path
        zs = ops.mlp_chain([self.dis_layer_dict[s] for s in sst_list], [user_embed]) if len(sst_list) > 1 else None
        for i, sst in enumerate(sst_list):
            z = zs[i] if zs is not None else self.dis_layer_dict[sst](user_embed)
            if self.sst_size[sst] == 2:
                loss = loss + ops.SigmoidBce.apply(z, interaction[sst].to(device=dev, dtype=torch.float32))
            else:
                loss = loss + ops.SoftmaxCe.apply(z, interaction[sst].to(device=dev, dtype=torch.int32))
        return loss

    def _with_dis(self, bpr, interaction, sst_list, shared_user_embed=None):
        """pfcn_mlp.py:188-191: the reference re-evaluates forward(user) inside calculate_dis_loss (a second pass through
        the filter in training mode: BatchNorm running statistics advance twice per step) -- kept, either literally or as
        the shared pass of _forward_for_loss"""
        if self.filter_mode != "none":
            if shared_user_embed is None:
                shared_user_embed, _ = self.forward(interaction[self.USER_ID], None, sst_list)
            return bpr - self.dis_weight * self._dis_terms(shared_user_embed, interaction, sst_list)
        return bpr

    def full_sort_predict(self, interaction, sst_list=None):
        raise NotImplementedError("PFCN full-sort scoring is undefined in the reference (forward() returns a tuple "
                                  "that full_sort_predict uses as a tensor); use predict() over all items")

    def get_sst_embed(self, user_data, sst_list=None):
        """pfcn_mlp.py:224-232"""
        ret = {}
        user_indices = torch.arange(1, self.n_users)
        sst_list = self.sst_attrs if self.filter_mode == "none" else sst_list
        for sst in sst_list:
            ret[sst] = user_data[sst][user_indices - 1]
        ret["embedding"], _ = self.forward(user_indices.to(self.device), None, sst_list)
        return ret


class PFCN_MLP(_PFCNBase):
    """pfcn_mlp.py:23-232: NCF-style tower over [filtered user || item]"""
    USER_EMB, ITEM_EMB = "user_embedding", "item_embedding"
    SCORER_FIRST = False      # pfcn_mlp.py:55-63: filters and discriminators first, then embeddings and tower

    def _build_scorer(self, config):
        self.dropout = config["dropout"]
        self.mlp_hidden_size_list = list(config["mlp_hidden_size_list"])
        self.user_embedding = nn.Embedding(self.n_users, self.embedding_size)
        self.item_embedding = nn.Embedding(self.n_items, self.embedding_size)
        self.mlp_layer = MLPLayers([self.embedding_size * 2] + self.mlp_hidden_size_list + [1], dropout=self.dropout)

    def _score(self, user_embed, item_embed):
        return self.mlp_layer(ops.ConcatCols.apply(user_embed, item_embed))

    def predict(self, interaction, sst_list=None):
        """pfcn_mlp.py:169-175"""
        u, i = self.forward(interaction[self.USER_ID], interaction[self.ITEM_ID], sst_list)
        return ops.Act.apply(self._score(u, i), ops.ACT["sigmoid"])

    def calculate_loss(self, interaction, sst_list=None):
        """pfcn_mlp.py:177-193: BPR(pos, neg) - dis_weight * discriminator loss"""
        user_embed, pos_embed, shared = self._forward_for_loss(interaction[self.USER_ID], interaction[self.POS_ITEM_ID], sst_list)
        neg_embed = self._item_base(interaction[self.NEG_ITEM_ID])
        # the positive and the negative pass of the tower share one forward and one backward launch (two chains of the same
        # module: autograd adds their weight gradients); None = per-module path
        both = ops.mlp_chain([self.mlp_layer, self.mlp_layer], [ops.ConcatCols.apply(user_embed, pos_embed),
                                                                ops.ConcatCols.apply(user_embed, neg_embed)])
        if both is not None:
            bpr = ops.BprLoss.apply(both[0], both[1])
        else:
            bpr = ops.BprLoss.apply(self._score(user_embed, pos_embed), self._score(user_embed, neg_embed))
        return self._with_dis(bpr, interaction, sst_list, user_embed if shared else None)


class PFCN_PMF(_PFCNBase):
    """pfcn_pmf.py: dot-product scorer"""

    def _build_scorer(self, config):
        self.user_embedding_layer = nn.Embedding(self.n_users, self.embedding_size)
        self.item_embedding_layer = nn.Embedding(self.n_items, self.embedding_size)

    def predict(self, interaction, sst_list=None):
        """pfcn_pmf.py:166-174 (keepdim=True: [B,1])"""
        u, i = self.forward(interaction[self.USER_ID], interaction[self.ITEM_ID], sst_list)
        return ops.Act.apply(ops.RowDot.apply(u, i), ops.ACT["sigmoid"]).view(-1, 1)

    def calculate_loss(self, interaction, sst_list=None):
        """pfcn_pmf.py:176-193"""
        user_embed, pos_embed, shared = self._forward_for_loss(interaction[self.USER_ID], interaction[self.POS_ITEM_ID], sst_list)
        neg_embed = self._item_base(interaction[self.NEG_ITEM_ID])
        bpr = ops.BprLoss.apply(ops.RowDot.apply(user_embed, pos_embed), ops.RowDot.apply(user_embed, neg_embed))
        return self._with_dis(bpr, interaction, sst_list, user_embed if shared else None)


class PFCN_BiasedMF(_PFCNBase):
    """pfcn_biasedmf.py: dot product + user / item / global biases"""

    def _build_scorer(self, config):
        self.user_embedding_layer = nn.Embedding(self.n_users, self.embedding_size)
        self.user_bias = nn.Embedding(self.n_users, 1)
        self.item_embedding_layer = nn.Embedding(self.n_items, self.embedding_size)
        self.item_bias = nn.Embedding(self.n_items, 1)
        self.global_bias = nn.Parameter(torch.tensor(0.1))

    def _bias(self, table, ids):
        return ops.GatherRows.apply(table.weight, self._ids(ids))          # [B,1]

    def predict(self, interaction, sst_list=None):
        """pfcn_biasedmf.py:170-181: sigmoid(dot + b_u + b_i + b_g), [B,1]"""
        user, item = interaction[self.USER_ID], interaction[self.ITEM_ID]
        with torch.no_grad():
            u, i = self.forward(user, item, sst_list)
            return ops.biased_score(ops.RowDot.apply(u, i), self._bias(self.user_bias, user),
                                    self._bias(self.item_bias, item), self.global_bias, ops.ACT["sigmoid"]).view(-1, 1)

    def calculate_loss(self, interaction, sst_list=None):
        """pfcn_biasedmf.py:183-199 (the [B] + [B,1] broadcast makes the BPR run over B*B pairs)"""
        user, pos, neg = interaction[self.USER_ID], interaction[self.POS_ITEM_ID], interaction[self.NEG_ITEM_ID]
        user_embed, pos_embed, shared = self._forward_for_loss(user, pos, sst_list)
        neg_embed = self._item_base(neg)
        bpr = ops.BprOuter.apply(ops.RowDot.apply(user_embed, pos_embed), ops.RowDot.apply(user_embed, neg_embed),
                                 self._bias(self.user_bias, user), self._bias(self.item_bias, pos),
                                 self._bias(self.item_bias, neg), self.global_bias.view(1))
        return self._with_dis(bpr, interaction, sst_list, user_embed if shared else None)


class PFCN_DMF(_PFCNBase):
    """pfcn_dmf.py: user / item MLPs + cosine similarity (x10 in the loss)"""

    def __init__(self, config, dataset):
        self.num_layers = config["num_layers"]
        self.mlp_dropout = config["mlp_dropout"]
        self.mlp_activation = config["mlp_activation"]
        self.dis_activation = config["dis_activation"]
        super().__init__(config, dataset)

    def _filter_activation(self):
        return self.mlp_activation          # pfcn_dmf.py:112

    def _dis_activation(self):
        return self.dis_activation          # pfcn_dmf.py:136

    def _build_scorer(self, config):
        e = self.embedding_size
        self.user_embedding_layer = nn.Embedding(self.n_users, e)
        self.item_embedding_layer = nn.Embedding(self.n_items, e)
        self.user_mlp = MLPLayers([e] * (self.num_layers + 1), dropout=self.mlp_dropout, activation=self.mlp_activation,
                                  init_method="norm")
        self.item_mlp = MLPLayers([e] * (self.num_layers + 1), dropout=self.mlp_dropout, activation=self.mlp_activation,
                                  init_method="norm")

    def _user_path_dropout(self):
        return float(self.mlp_dropout)          # the user tower draws a new dropout mask per evaluation: no sharing unless 0

    def _user_base(self, user):
        return self.user_mlp(super()._user_base(user))

    def _item_base(self, item):
        return self.item_mlp(super()._item_base(item))

    def predict(self, interaction, sst_list=None):
        """pfcn_dmf.py:170-178"""
        u, i = self.forward(interaction[self.USER_ID], interaction[self.ITEM_ID], sst_list)
        return ops.Act.apply(ops.CosineSim.apply(u, i), ops.ACT["sigmoid"])

    def calculate_loss(self, interaction, sst_list=None):
        """pfcn_dmf.py:180-199"""
        user_embed, pos_embed, shared = self._forward_for_loss(interaction[self.USER_ID], interaction[self.POS_ITEM_ID], sst_list)
        neg_embed = self._item_base(interaction[self.NEG_ITEM_ID])
        pos = ops.WeightedSum.apply(10.0, ops.CosineSim.apply(user_embed, pos_embed))
        neg = ops.WeightedSum.apply(10.0, ops.CosineSim.apply(user_embed, neg_embed))
        return self._with_dis(ops.BprLoss.apply(pos, neg), interaction, sst_list, user_embed if shared else None)


class _DpOptimizer:
    """an AdamGroup whose step first sums the ranks' gradient shares (ops.ChainDP.all_reduce_grads: one NCCL all-reduce)"""

    def __init__(self, opt, dp):
        self.opt, self.dp = opt, dp

    def __getattr__(self, k):
        return getattr(self.opt, k)

    def step(self):
        self.dp.all_reduce_grads(self.opt.params)
        self.opt.step()


class PFCNTrainer(CheckpointMixin):
    """The alternating schedule of PFCNTrainer and its per-model subclasses (trainer.py:865-898, 1189-1235): per epoch
    a random non-empty attribute subset; every `train_epoch_interval`-th epoch one pass on `bpr - dis_weight * dis` with
    the filter optimizer (base model + filters), then always one pass on `dis` with the discriminator optimizer.
    Optimizer steps run on this package's Adam kernel (ops.AdamGroup = torch.optim.Adam semantics, L2 form)."""

    def __init__(self, config, model, dp=None):
        """dp: ops.ChainDP -- data-parallel training over the GPUs of one box (one process per GPU):
        every rank takes rows [rank * B / world, (rank + 1) * B / world) of each batch (a tail of B % world rows is dropped),
        BatchNorm keeps the statistics of the WHOLE batch (the chain kernels exchange their column sums through NVLink peer
        memory, ops.set_chain_dp), the loss is the mean over the whole batch, and the ranks' gradient shares are summed by
        one NCCL all-reduce before every Adam step: the replicas stay identical and follow the single-GPU trajectory."""
        self.config, self.model, self.dp = config, model, dp
        if dp is not None:
            ops.set_chain_dp(dp)
        self.filter_mode = config["filter_mode"].lower()
        self.train_epoch_interval = config["train_epoch_interval"] or 1
        lr, wd = config["learning_rate"], config["weight_decay"] or 0.0
        base = list(model.parameters())
        if self.filter_mode != "none":
            self.sst_attrs = list(config["sst_attr_list"])
            fparams = [p for f in model.filter_layer.values() for p in f.parameters()]
            dparams = [p for d in model.dis_layer_dict.values() for p in d.parameters()]
            self.optimizer_filter = ops.AdamGroup(base + fparams, lr=lr, weight_decay=wd)
            self.optimizer_dis = ops.AdamGroup(dparams, lr=lr, weight_decay=wd)
        else:
            self.optimizer_filter = ops.AdamGroup(base, lr=lr, weight_decay=wd)
        if dp is not None:
            self.optimizer_filter = _DpOptimizer(self.optimizer_filter, dp)
            if self.filter_mode != "none":
                self.optimizer_dis = _DpOptimizer(self.optimizer_dis, dp)

    def _shard(self, interaction):
        """this rank's rows of a batch (data-parallel mode)"""
        if self.dp is None:
            return interaction
        from .interaction import Interaction
        n = len(interaction) // self.dp.world
        lo = self.dp.rank * n
        return Interaction({k: interaction[k][lo:lo + n] for k in interaction.columns})

    def _dp_loss(self, loss_func):
        if self.dp is None:
            return loss_func
        world = self.dp.world

        def scaled(interaction, sst_list):        # the mean over the whole batch = the mean of the ranks' means
            return ops.WeightedSum.apply(1.0 / world, loss_func(interaction, sst_list))
        scaled.__name__ = loss_func.__name__
        return scaled

    def _pass(self, train_data, loss_func, optimizer, sst_list):
        self.model.train()
        if self.config["use_cuda_graph"]:
            return self._pass_graphed(train_data, loss_func, optimizer, sst_list)
        total = None
        loss_func = self._dp_loss(loss_func)
        for interaction in train_data:
            optimizer.zero_grad()
            loss = loss_func(self._shard(interaction), sst_list)
            v = loss.item()
            if v != v:
                raise ValueError("Training loss is nan")
            total = v if total is None else total + v
            loss.backward()
            optimizer.step()
        return self._dp_total(total)

    def _dp_total(self, total):
        """the pass's summed loss over all ranks (each rank summed its share of every batch mean)"""
        if self.dp is None or total is None:
            return total
        import torch.distributed as dist
        t = torch.tensor([total], dtype=torch.float64, device=next(self.model.parameters()).device)
        dist.all_reduce(t, group=self.dp.group)
        return float(t.item())

    def _pass_graphed(self, train_data, loss_func, optimizer, sst_list):
        """`use_cuda_graph: True`: one CUDA-graph replay per batch (graphed.py); the per-batch losses are summed on the
        device and read back once per pass (the NaN check of trainer.py:192 moves to the end of the pass)"""
        from .graphed import GraphedSteps
        if getattr(self, "_graphs", None) is None:
            self._graphs = GraphedSteps(next(self.model.parameters()).device)
            self._graph_gen = 0
        total = None
        name = f"{loss_func.__name__}:{id(optimizer)}:{getattr(self.model, 'train_stage', '')}"
        loss_func = self._dp_loss(loss_func)
        for interaction in train_data:
            loss = self._graphs.run(name, loss_func, optimizer, sst_list, self._shard(interaction))
            total = loss.clone() if total is None else total + loss
        v = float(total.item()) if total is not None else None
        if v is not None and v != v:
            raise ValueError("Training loss is nan")
        return self._dp_total(v)

    @torch.no_grad()
    def evaluate(self, eval_data, sst_list=None, train_item_count=None):
        """PFCNTrainer.pfcn_evaluate / evaluate (trainer.py:1010-1093) for one attribute subset: the sampled-negative
        (`uni100`) evaluation -- the only one the reference defines for this family -- with `model.predict` as the scorer
        and the fused candidate top-K + metrics (sampled_eval.SampledEvaluator).  eval_data: SampledEvalData."""
        from .interaction import Interaction
        from .sampled_eval import SampledEvaluator
        self.model.eval()
        if hasattr(eval_data, "resample"):          # uni<N> source that redraws its negatives per evaluation
            eval_data = eval_data.resample()
        if getattr(self, "sampled_evaluator", None) is None:
            self.sampled_evaluator = SampledEvaluator(self.config, self.model.n_items, train_item_count)
        sst_list = (self.sst_attrs if self.filter_mode != "none" else None) if sst_list is None else sst_list
        m = self.model

        def score_fn(uid, iid):
            return m.predict(Interaction({m.USER_ID: uid, m.ITEM_ID: iid}), sst_list).view(-1)

        return self.sampled_evaluator.evaluate(score_fn, eval_data)

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


    def attribute_subsets(self):
        """trainer.py:1012-1013, 1076-1077: the non-empty subsets of `sst_attr_list`, by size, in combination order"""
        import itertools
        return [list(c) for i in range(1, len(self.sst_attrs) + 1) for c in itertools.combinations(self.sst_attrs, i)]

    @torch.no_grad()
    def pfcn_evaluate(self, eval_data, train_item_count=None):
        """The reference's VALIDATION pass (trainer.py:985-1030): every user batch is scored under each non-empty attribute
        subset with the SAME negatives and all of it goes into one collector struct, so the validation metrics are taken
        over (subset, user) pairs.  Here: one draw of negatives, the candidate lists tiled once per subset, each tile scored
        with its subset, one pass of the evaluation kernels over the whole.  One subset (a single attribute, or
        filter_mode none) is the plain `evaluate`."""
        from .interaction import Interaction
        from .sampled_eval import SampledEvaluator
        subsets = self.attribute_subsets() if self.filter_mode != "none" else [None]
        if len(subsets) == 1 or not hasattr(eval_data, "resample_tiled"):
            return self.evaluate(eval_data, None, train_item_count)
        self.model.eval()
        if getattr(self, "sampled_evaluator", None) is None:
            self.sampled_evaluator = SampledEvaluator(self.config, self.model.n_items, train_item_count)
        data, per_copy = eval_data.resample_tiled(len(subsets))
        m, chunk = self.model, int(self.config["sampled_chunk_rows"] or (1 << 24))
        parts = []
        for s, sst_list in enumerate(subsets):
            for a in range(s * per_copy, (s + 1) * per_copy, chunk):
                b = min(a + chunk, (s + 1) * per_copy)
                inter = Interaction({m.USER_ID: data.cand_uid[a:b], m.ITEM_ID: data.cand_items[a:b]})
                parts.append(m.predict(inter, sst_list).view(-1).to(torch.float32))
        ev = self.sampled_evaluator
        return ev.finalize(ev.collect_scores(torch.cat(parts), data), data)

    @torch.no_grad()
    def evaluate_subsets(self, eval_data, train_item_count=None):
        """The reference's TEST evaluation (trainer.py:1072-1086): one full evaluation per attribute subset, each with its own
        draw of negatives -> {'<filter_mode>-<subset>': metric dict}"""
        if self.filter_mode == "none":
            return {self.filter_mode: self.evaluate(eval_data, None, train_item_count)}
        return {"{}-{}".format(self.filter_mode, sst_list): self.evaluate(eval_data, sst_list, train_item_count)
                for sst_list in self.attribute_subsets()}

    def fit(self, train_data, valid_data=None, epochs=None, train_item_count=None, verbose=False, saved=False):
        """trainer.py:300-380 around the alternating epochs: early stopping on `valid_metric` (all attributes filtered,
        trainer.py:1010-1093), check-point on improvement when saved=True (checkpoint.py), continue at `start_epoch` after
        `resume_checkpoint`.  Returns (best valid score, best valid result)."""
        from .trainer import early_stopping
        metric = (self.config["valid_metric"] or "NDCG@5").lower()
        bigger = self.config["valid_metric_bigger"] if self.config["valid_metric_bigger"] is not None else True
        start = getattr(self, "start_epoch", 0)
        best, best_res, cur = (-np.inf if bigger else np.inf), None, 0
        if start > 0:
            best, cur = self.best_valid_score, self.cur_step
        for epoch in range(start, epochs if epochs is not None else (self.config["epochs"] or 1)):
            losses = self._train_epoch(train_data, epoch)
            if verbose:
                print(f"epoch {epoch}: losses {losses}")
            if not valid_data:
                continue
            res = self.pfcn_evaluate(valid_data, train_item_count)
            best, cur, stop, update = early_stopping(res[metric], best, cur, max_step=_stopping_step(self.config),
                                                     bigger=bigger)
            self.best_valid_score, self.cur_step = best, cur
            if update:
                best_res = res
                if saved:
                    self._save_checkpoint(epoch)
            if stop:
                break
        return best, best_res


PFCN_MLPTrainer = PFCN_PMFTrainer = PFCN_BiasedMFTrainer = PFCN_DMFTrainer = PFCNTrainer
