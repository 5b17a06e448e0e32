"""The FairGo family (graph-based fair representation: filters over the whole ego-embedding table + node-level and
ego-network-level discriminators) -- drop-ins for recbole/model/fair_recommender/fairgo_{pmf,gcn}.py, computed by this
package's kernels (layers.MLPLayers over fr_linear_*, ops.Spmm over fr_spmm_csr, ops.*).

Kept from the reference, on purpose:
  * filters / discriminators live in plain Python dicts (`filter_layer_dict`, `dis_layer_dict`; fairgo_pmf.py:141-157),
    so they are not part of `state_dict()`; `aggr_layer` IS a registered nn.Sequential (Linear, act, Linear, act, Linear);
  * the filters run over ALL n_users + n_items rows every step and their sum is divided by the TOTAL number of filters
    (fairgo_pmf.py:163-168);
  * the multi-class ego-network loss applies a sigmoid before CrossEntropy while the node-level one does not
    (fairgo_pmf.py:231 vs 226);
  * `D^-1 A` uses rowsum + 1e-7 in float32 (fairgo_pmf.py:100-127).
Differences, documented: in the fine-tune stage the embedding tables are constants (no optimizer of the reference's
FairGo trainers contains them), so this implementation does not back-propagate into them; `calculate_loss` evaluates
`forward(sst_list)` once and shares it between the rating term and the discriminator term (the reference evaluates the
identical expression twice).
FairGo_GCN: the fine-tune stage is identical to FairGo_PMF's.  Its pretrain stage calls torch_geometric.nn.GCN, a
third-party dependency that is neither vendored nor pinned by the reference (parity unpinned); it
is restated from the published algorithm on this package's SpMM / linear kernels (see FairGo_GCN), pinned on the oracle's
restatement only.
"""
import numpy as np
import scipy.sparse as sp
import torch
import torch.nn as nn

from .checkpoint import CheckpointMixin
from .utils import stopping_step as _stopping_step
from . import ops
from .layers import _ACT_MODULES, MLPLayers


def norm_rating_csr(rating_coo, n_users, n_items):
    """fairgo_pmf.py:100-127 get_norm_rating_matrix: L = D^-1 A over the bipartite graph, float32, as CSR"""
    r = rating_coo.tocoo()
    row = np.concatenate([r.row, r.col + n_users]).astype(np.int64)
    col = np.concatenate([r.col + n_users, r.row]).astype(np.int64)
    val = np.concatenate([r.data, r.data]).astype(np.float32)
    n = n_users + n_items
    A = sp.csr_matrix((val, (row, col)), shape=(n, n), dtype=np.float32)   # unique (user, item) pairs: no duplicates to sum
    diag = (np.asarray(A.sum(axis=1)).ravel().astype(np.float32) + np.float32(1e-7)).astype(np.float32)
    inv = (np.float32(1.0) / diag).astype(np.float32)
    L = sp.diags(inv).astype(np.float32) @ A
    return sp.csr_matrix(L, dtype=np.float32)


class FairGo_PMF(nn.Module):
    input_type = "POINTWISE"
    type = "GENERAL"

    def __init__(self, config, dataset):
        super().__init__()
        self.USER_ID, self.ITEM_ID = config["USER_ID_FIELD"], config["ITEM_ID_FIELD"]
        self.RATING = config["RATING_FIELD"]
        self.n_users, self.n_items = dataset.num(self.USER_ID), dataset.num(self.ITEM_ID)
        self.device = config["device"]
        self.n_layers = config["n_layers"]
        self.act = config["activation"]
        self.embedding_size = config["embedding_size"]
        self.dis_hidden_size_list = list(config["dis_hidden_size_list"])
        self.filter_hidden_size_list = list(config["filter_hidden_size_list"])
        self.sst_attrs = list(config["sst_attr_list"])
        self.fair_weight = config["fair_weight"]
        self.load_pretrain_weight = config["load_pretrain_weight"]
        self.train_stage = None
        self.aggr_method = config["aggr_method"].upper()
        self.vs_weights = None
        if config["vs_weights"] is not None:
            w = np.asarray(config["vs_weights"], np.float32)
            self.vs_weights = (w / w.sum(dtype=np.float32)).astype(np.float32)
            if self.aggr_method == "LVA":
                assert self.n_layers == len(self.vs_weights), "n_layers should be equal to length of vs_weights"
        self.max_rating = float(dataset.inter_feat[self.RATING].max())
        self.rating_matrix = dataset.inter_matrix(form="coo", value_field=self.RATING).astype(np.float32)
        self.sst_size = self._get_sst_size(dataset.get_user_feature())

        self.user_embedding_layer = nn.Embedding(self.n_users, self.embedding_size, padding_idx=0)
        self.item_embedding_layer = nn.Embedding(self.n_items, self.embedding_size, padding_idx=0)
        if self.load_pretrain_weight:
            self.user_embedding_layer.weight.data.copy_(torch.from_numpy(dataset.get_preload_weight("uid")))
            self.item_embedding_layer.weight.data.copy_(torch.from_numpy(dataset.get_preload_weight("iid")))
        self.dis_layer_dict = self.init_dis_layers()
        self.filter_layer_dict = self.init_filter_layers()
        e = self.embedding_size
        act = _ACT_MODULES[self.act.lower()]
        self.aggr_layer = nn.Sequential(nn.Linear(self.n_layers * e, e), act(), nn.Linear(e, e), act(), nn.Linear(e, e))
        self._norm_csr = norm_rating_csr(self.rating_matrix, self.n_users, self.n_items)
        self.norm_rating_matrix = None          # ops.SpmmMatrix, built on the model's device at first use
        self._ego = None

    # ------------------------------------------------------------------ construction (fairgo_pmf.py:78-157)
    def _get_sst_size(self, user_feature):
        size = {}
        for sst in self.sst_attrs:
            if sst not in user_feature.columns:
                raise ValueError(f"{sst} sensitive attribute not in user feature")
            size[sst] = len(user_feature[sst][1:].unique())
        return size

    def init_dis_layers(self):
        return {sst: MLPLayers([self.embedding_size] + self.dis_hidden_size_list +
                               [1 if self.sst_size[sst] == 2 else self.sst_size[sst]], activation=self.act).to(self.device)
                for sst in self.sst_attrs}

    def init_filter_layers(self):
        return {sst: MLPLayers([self.embedding_size] + self.filter_hidden_size_list + [self.embedding_size],
                               activation=self.act).to(self.device) for sst in self.sst_attrs}

    def _dict_modules(self):
        return list(self.filter_layer_dict.values()) + list(self.dis_layer_dict.values())

    def to(self, *a, **k):
        super().to(*a, **k)
        for m in self._dict_modules():
            m.to(*a, **k)
        self.norm_rating_matrix = None
        self._ego = None
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

    # ------------------------------------------------------------------ forward pieces
    def _dev(self):
        return self.user_embedding_layer.weight.device

    def _ids(self, t):
        return t.to(device=self._dev(), dtype=torch.int32).contiguous()

    def _matrix(self):
        if self.norm_rating_matrix is None:
            self.norm_rating_matrix = ops.SpmmMatrix(self._norm_csr, self._dev())
        return self.norm_rating_matrix

    def get_ego_embeddings(self):
        """fairgo_pmf.py:128-138; the [N, d] concatenation is cached while the tables do not change (fine-tune stage)"""
        U, I = self.user_embedding_layer.weight, self.item_embedding_layer.weight
        key = (U.data_ptr(), I.data_ptr(), U._version, I._version)
        if self._ego is None or self._ego[0] != key:
            self._ego = (key, torch.cat([U.detach(), I.detach()], dim=0).contiguous())
        return self._ego[1]

    def _aggr(self, x):
        """nn.Sequential(Linear, act, Linear, act, Linear) (fairgo_pmf.py:66-70) on the linear kernels"""
        act = ops.ACT[self.act.lower()]
        lins = [m for m in self.aggr_layer if isinstance(m, nn.Linear)]
        fused = ops.mlp_chain([ops.SeqChain([(lin, None, act if k + 1 < len(lins) else 0, 0.0) for k, lin in enumerate(lins)],
                                            self.training)], [x])
        if fused is not None:       # the three layers in one forward / one backward launch (mlp_chain.cu)
            return fused[0]
        for k, lin in enumerate(lins):
            x = ops.LinearAct.apply(x, lin.weight, lin.bias, act if k + 1 < len(lins) else 0, 0.0, 0)
        return x

    def _forward_all(self, sst_list=None):
        """fairgo_pmf.py:159-169: the [N, d] table the step works on (filtered in the fine-tune stage)"""
        if self.train_stage == "finetune":
            sst_list = self.sst_attrs if sst_list is None else sst_list
            ego = self.get_ego_embeddings()
            outs = [self.filter_layer_dict[s](ego) for s in sst_list]
            return ops.SumDiv.apply(len(self.filter_layer_dict), *outs)
        return None

    def forward(self, sst_list=None):
        """fairgo_pmf.py:159-171 -> (user_all_embeddings [n_users, d], item_all_embeddings [n_items, d])"""
        table = self._forward_all(sst_list)
        if table is None:
            return self.user_embedding_layer.weight, self.item_embedding_layer.weight
        return table[: self.n_users], table[self.n_users:]

    def _rows(self, table, interaction):
        """(user rows, item rows) of the working table for a batch; item ids are offset by n_users in the [N, d] table"""
        uid, iid = self._ids(interaction[self.USER_ID]), self._ids(interaction[self.ITEM_ID])
        if table is None:       # pretrain stage: the two parameter tables
            return (ops.GatherRows.apply(self.user_embedding_layer.weight, uid),
                    ops.GatherRows.apply(self.item_embedding_layer.weight, iid))
        return ops.GatherRows.apply(table, uid), ops.GatherRows.apply(table, iid + self.n_users)

    def _rating_loss(self, interaction, table):
        u, i = self._rows(table, interaction)
        rating = interaction[self.RATING].to(device=self._dev(), dtype=torch.float32)
        return ops.MseLoss.apply(ops.RowDot.apply(u, i), rating)

    def calculate_loss(self, interaction, sst_list=None):
        """fairgo_pmf.py:173-188: mse - fair_weight * (node + ego-network discriminator loss) in the fine-tune stage"""
        table = self._forward_all(sst_list)
        mse = self._rating_loss(interaction, table)
        if self.train_stage == "finetune":
            return mse - self.fair_weight * self._dis_loss(interaction, sst_list, table)
        return mse

    def calculate_dis_loss(self, interaction, sst_list):
        """fairgo_pmf.py:190-236"""
        # the discriminator phase optimises the discriminators (+ aggr_layer) only: the filtered table enters as a
        # constant (the reference also back-propagates into the filters here, but nothing ever uses those gradients)
        with torch.no_grad():
            table = self._forward_all(sst_list)
            if table is None:
                table = torch.cat([self.user_embedding_layer.weight, self.item_embedding_layer.weight], dim=0)
        return self._dis_loss(interaction, sst_list, table)

    def _dis_loss(self, interaction, sst_list, table):
        dev = self._dev()
        uid = self._ids(interaction[self.USER_ID])
        user_node = ops.GatherRows.apply(table, uid)
        mat = self._matrix()
        graph, cur = [], table
        for _ in range(self.n_layers):
            cur = ops.Spmm.apply(cur, mat)
            graph.append(cur)
        lva = self.aggr_method == "LVA" and self.n_layers > 1
        if self.n_layers == 1:
            all_graph = graph[0]
        elif self.aggr_method == "WAP":
            all_graph = ops.SumDiv.apply(self.n_layers, *graph)
        elif self.aggr_method == "LBA":
            all_graph = self._aggr(ops.ConcatCols.apply(*graph))
        elif lva:
            all_graph = [ops.GatherRows.apply(g, uid) for g in graph]      # user rows are the first n_users of [N, d]
        else:
            raise ValueError(f"unknown aggr_method {self.aggr_method}")
        if not lva:
            user_local = ops.GatherRows.apply(all_graph, uid)
        node, local = 0.0, 0.0
        sig = ops.ACT["sigmoid"]
        for sst in sst_list:
            dis = self.dis_layer_dict[sst]
            # the discriminator's passes over the node rows and the ego-network rows share one forward and one backward
            # launch (chains of the same module: autograd adds their weight gradients); None = per-module path
            ins = [user_node] + (list(all_graph) if lva else [user_local])
            zs = ops.mlp_chain([dis] * len(ins), ins) if len(ins) <= 4 else None
            if zs is None:
                zs = [dis(t) for t in ins]
            if self.sst_size[sst] == 2:
                y = interaction[sst].to(device=dev, dtype=torch.float32)
                node = node + ops.SigmoidBce.apply(zs[0], y)
                if lva:
                    for k, w in enumerate(self.vs_weights):
                        local = local + float(w) * ops.SigmoidBce.apply(zs[1 + k], y)
                else:
                    local = local + ops.SigmoidBce.apply(zs[1], y)
            else:
                y = interaction[sst].to(device=dev, dtype=torch.int32)
                node = node + ops.SoftmaxCe.apply(zs[0], y)
                if lva:
                    for k, w in enumerate(self.vs_weights):
                        local = local + float(w) * ops.SoftmaxCe.apply(ops.Act.apply(zs[1 + k], sig), y)
                else:
                    local = local + ops.SoftmaxCe.apply(ops.Act.apply(zs[1], sig), y)
        return node + local

    def predict(self, interaction):
        """fairgo_pmf.py:238-248"""
        with torch.no_grad():
            user_all, item_all = self.forward()
            u = ops.GatherRows.apply(user_all, self._ids(interaction[self.USER_ID]))
            i = ops.GatherRows.apply(item_all, self._ids(interaction[self.ITEM_ID]))
            return ops.clamp_div(ops.RowDot.apply(u, i), self.max_rating)

    def filtered_tables(self):
        """(U', I') = forward(all attributes): what full_sort_predict scores with (fairgo_pmf.py:250-257); feed them to
        evaluator.FullSortEvaluator (transform clamp/max) for the fused full-sort fair evaluation"""
        with torch.no_grad():
            user_all, item_all = self.forward()
            return user_all.contiguous(), item_all.contiguous()

    def full_sort_predict(self, interaction):
        """fairgo_pmf.py:250-257: clamp(U'[users] . I'^T, 0, max_rating) / max_rating, flattened"""
        from . import _lib, kernels
        U, I = self.filtered_tables()
        users = self._ids(interaction[self.USER_ID])
        uid = users.repeat_interleave(self.n_items).contiguous()
        iid = torch.arange(self.n_items, dtype=torch.int32, device=U.device).repeat(users.numel()).contiguous()
        return kernels.pair_scores(U, I, uid, iid, _lib.TRANSFORM_CLAMP_DIV, self.max_rating)

    def get_sst_embed(self, user_data, sst_list=None):
        """fairgo_pmf.py:259-269"""
        ret = {}
        user_indices = torch.arange(1, self.n_users)
        sst_list = self.sst_attrs if sst_list is None else sst_list
        for sst in sst_list:
            ret[sst] = user_data[sst][user_indices - 1]
        with torch.no_grad():
            user_all, _ = self.forward()
        ret["embedding"] = user_all[user_indices.to(user_all.device)]
        return ret


def gcn_norm_csr(rating_matrix, n_users, n_items):
    """Kipf & Welling's renormalised adjacency with edge weights, as torch_geometric's GCNConv builds it (gcn_norm with
    add_self_loops, fill value 1): A_hat = D^-1/2 (A + I) D^-1/2 over the bipartite rating graph of fairgo_gcn.py:60-66
    (user u <-> node n_users + i, weight = rating, both directions); deg_i = sum_j (A + I)_ij; isolated nodes keep their
    self loop.  Host-side scipy, once per model."""
    import scipy.sparse as sp
    N = n_users + n_items
    R = rating_matrix.tocoo()
    A = sp.coo_matrix((np.r_[R.data, R.data].astype(np.float64),
                       (np.r_[R.row, R.col + n_users], np.r_[R.col + n_users, R.row])), shape=(N, N)).tocsr()
    A = A + sp.identity(N, dtype=np.float64, format="csr")
    deg = np.asarray(A.sum(axis=1)).ravel()
    with np.errstate(divide="ignore"):
        dinv = np.where(deg > 0, deg ** -0.5, 0.0)
    return sp.csr_matrix(sp.diags(dinv) @ A @ sp.diags(dinv), dtype=np.float32)


class _GCNConv(nn.Module):
    """parameters of one torch_geometric GCNConv (`lin.weight` glorot-uniform [out, in], `bias` zeros)"""

    def __init__(self, n_in, n_out):
        super().__init__()
        self.lin = nn.Linear(n_in, n_out, bias=False)
        nn.init.xavier_uniform_(self.lin.weight)
        self.bias = nn.Parameter(torch.zeros(n_out))


class _GCN(nn.Module):
    """parameter tree of torch_geometric.nn.GCN(in, hidden, num_layers, out, dropout, act): `convs.<k>.lin.weight`,
    `convs.<k>.bias`; widths in -> hidden x (num_layers - 1) -> out"""

    def __init__(self, n_in, hidden, n_out, num_layers, dropout, act):
        super().__init__()
        dims = [n_in] + [hidden] * (num_layers - 1) + [n_out]
        self.convs = nn.ModuleList(_GCNConv(a, b) for a, b in zip(dims[:-1], dims[1:]))
        self.dropout = float(dropout or 0.0)
        self.act = (act or "relu").lower()


class FairGo_GCN(FairGo_PMF):
    """fairgo_gcn.py.  Fine-tune stage: identical to FairGo_PMF (the reference applies its filters to the RAW embedding
    tables there, fairgo_gcn.py:177-186; the GCN only shapes the pretrain stage).  Pretrain stage (fairgo_gcn.py:175-176):
    rating MSE on `GCN(ego embeddings)`, the GCN being torch_geometric.nn.GCN in the reference -- a third-party module the
    reference neither vendors nor pins.  It is restated here from the published algorithm (Kipf & Welling 2017 as
    implemented by GCNConv / BasicGNN: x <- A_hat (x W^T) + b, then act and dropout between layers, none after the last)
    and computed as LinearAct(Spmm(A_hat, dropout(x)), W, b, act) -- aggregation first, which is the same map by
    linearity -- on this package's kernels.  Parity for this stage is pinned on the oracle's restatement only
    (oracle/fairgo_oracle.gcn_forward), not on torch_geometric itself (absent from the build container)."""

    def __init__(self, config, dataset):
        super().__init__(config, dataset)
        self.gcn = _GCN(self.embedding_size, int(config["hidden_channels"] or 32), self.embedding_size,
                        int(config["gcn_n_layers"] or 2), config["gcn_dropout"], config["gcn_act"])
        self._gcn_csr = gcn_norm_csr(self.rating_matrix, self.n_users, self.n_items)
        self._gcn_mat = None

    def to(self, *a, **k):
        super().to(*a, **k)
        self._gcn_mat = None
        return self

    def _gcn_forward(self, x):
        if self._gcn_mat is None:
            self._gcn_mat = ops.SpmmMatrix(self._gcn_csr, self._dev())
        g = self.gcn
        act, last = ops.ACT[g.act], len(g.convs) - 1
        for k, conv in enumerate(g.convs):
            if k > 0 and self.training and g.dropout > 0:
                x = ops.Dropout.apply(x, g.dropout, ops.next_seed())
            x = ops.Spmm.apply(x, self._gcn_mat)
            x = ops.LinearAct.apply(x, conv.lin.weight, conv.bias, act if k < last else 0, 0.0, 0)
        return x

    def _forward_all(self, sst_list=None):
        if self.train_stage == "pretrain":
            ego = torch.cat([self.user_embedding_layer.weight, self.item_embedding_layer.weight], dim=0)
            return self._gcn_forward(ego)
        return super()._forward_all(sst_list)


class FairGoTrainer(CheckpointMixin):
    """FairGoTrainer / FairGo_PMFTrainer / FairGo_GCNTrainer (trainer.py:534-862): optional pretrain of the embedding
    tables on the rating loss, then the alternating fine-tune schedule -- per epoch a random non-empty attribute subset;
    every `train_epoch_interval`-th epoch one pass on `mse - fair_weight * dis` with the filter optimizer, then always one
    pass on `dis` with the discriminator optimizer (+ aggr_layer when aggr_method is LBA)."""

    def __init__(self, config, model):
        self.config, self.model = config, model
        self.train_epoch_interval = config["train_epoch_interval"] or 1
        self.sst_attrs = list(config["sst_attr_list"])
        lr, wd = config["learning_rate"], config["weight_decay"] or 0.0
        self.load_pretrain_weight = config["load_pretrain_weight"]
        if config["pretrain_model_file_path"] is not None:
            # trainer.py:541-549: the fine-tune stage starts from the checkpointed pretrain model
            checkpoint = torch.load(config["pretrain_model_file_path"], map_location=model._dev(), weights_only=False)
            model.load_state_dict(checkpoint["state_dict"])
            model.load_other_parameter(checkpoint.get("other_parameter"))
            model._ego = None           # the tables changed under the cached [N, d] concatenation
            self.saved_pretrain_model_file = config["pretrain_model_file_path"]
            model.train_stage = "finetune"
        elif self.load_pretrain_weight:
            model.train_stage = "finetune"
        else:
            model.train_stage = "pretrain"
            self.pretrain_epochs = config["pretrain_epochs"]
            if self.pretrain_epochs is None:
                raise ValueError("pretrain_epochs must be set when neither load_pretrain_weight nor "
                                 "pretrain_model_file_path is given (FairGo_*.yaml: 600)")
            # the reference's pretrain optimizer holds model.parameters(); the ones that receive a gradient in this stage:
            self.optimizer_pretrain = ops.AdamGroup([model.user_embedding_layer.weight, model.item_embedding_layer.weight] +
                                                    (list(model.gcn.parameters()) if hasattr(model, "gcn") else []),
                                                    lr=lr, weight_decay=wd)
        dparams = [p for m in model.dis_layer_dict.values() for p in m.parameters()]
        if str(config["aggr_method"]).upper() == "LBA":
            dparams += list(model.aggr_layer.parameters())
        self.optimizer_dis = ops.AdamGroup(dparams, lr=lr, weight_decay=wd)
        self.optimizer_filter = ops.AdamGroup([p for m in model.filter_layer_dict.values() for p in m.parameters()],
                                              lr=lr, weight_decay=wd)

    def _pass(self, train_data, loss_func, optimizer, sst_list):
        self.model.train()
        if self.config["use_cuda_graph"]:
            return self._pass_graphed(train_data, loss_func, optimizer, sst_list)
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

    def pretrain(self, train_data, epochs=None, valid_data=None):
        """trainer.py:606-685: rating-loss epochs over the embedding tables (and the GCN of FairGo_GCN).  With `valid_data`
        the reference's bookkeeping applies: evaluate after every epoch, early stopping on `valid_metric` after
        `stopping_step` non-improving evaluations, and the BEST pretrained state -- not the last -- goes into the
        fine-tune stage (the reference reloads its best pretrain checkpoint, trainer.py:677-679; kept in memory here)."""
        from .trainer import early_stopping
        self.model.train_stage = "pretrain"
        n = epochs if epochs is not None else self.pretrain_epochs
        losses = []
        if not valid_data:
            losses = [self._pass(train_data, self.model.calculate_loss, self.optimizer_pretrain, None) for _ in range(n)]
        else:
            metric = (self.config["valid_metric"] or "NDCG@5").lower()
            bigger = self.config["valid_metric_bigger"] if self.config["valid_metric_bigger"] is not None else True
            best, cur, best_state = (-np.inf if bigger else np.inf), 0, None
            for _ in range(n):
                losses.append(self._pass(train_data, self.model.calculate_loss, self.optimizer_pretrain, None))
                res = self.evaluate(valid_data)            # pretrain stage: the raw tables (FairGo_GCN: the GCN output)
                best, cur, stop, update = early_stopping(res[metric], best, cur,
                                                         max_step=_stopping_step(self.config), bigger=bigger)
                if update:
                    best_state = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
                if stop:
                    break
            if best_state is not None:
                self.model.load_state_dict(best_state)
            self.pretrain_valid_score = best
        self.model.train_stage = "finetune"
        self.model._ego = None          # the tables moved under the cached [N, d] concatenation
        return losses

    def _pass_graphed(self, train_data, loss_func, optimizer, sst_list):
        """`use_cuda_graph: True`: one CUDA-graph replay per batch (graphed.py); the per-batch losses are summed on the
        device and read back once per pass (the NaN check of trainer.py:192 moves to the end of the pass)"""
        from .graphed import GraphedSteps
        if getattr(self, "_graphs", None) is None:
            self._graphs = GraphedSteps(next(self.model.parameters()).device)
            self._graph_gen = 0
        total = None
        name = f"{loss_func.__name__}:{id(optimizer)}:{getattr(self.model, 'train_stage', '')}"
        for interaction in train_data:
            loss = self._graphs.run(name, loss_func, optimizer, sst_list, interaction)
            total = loss.clone() if total is None else total + loss
        v = float(total.item()) if total is not None else None
        if v is not None and v != v:
            raise ValueError("Training loss is nan")
        return v

    def _train_epoch(self, train_data, epoch_idx):
        """trainer.py:687-704 -> (dis_loss, filter_loss)"""
        mask = np.zeros(len(self.sst_attrs))
        while mask.sum() == 0:
            mask = np.random.choice([0, 1], len(self.sst_attrs))
        sst_list = [s for s, m in zip(self.sst_attrs, mask) if m != 0]
        filter_loss = 0.0
        if epoch_idx % self.train_epoch_interval == 0:
            filter_loss = self._pass(train_data, self.model.calculate_loss, self.optimizer_filter, sst_list)
        dis_loss = self._pass(train_data, self.model.calculate_dis_loss, self.optimizer_dis, sst_list)
        return dis_loss, filter_loss


    # ------------------------------------------------------------------ evaluation / fit (trainer.py:739-772, 332-418)
    def data_collect(self, train_item_count):
        """collector.py:80-95: {item id: #train interactions} for the popularity metric"""
        self._train_item_count = dict(train_item_count) if train_item_count is not None else None
        self.evaluator = self.sampled_evaluator = None

    @torch.no_grad()
    def evaluate(self, eval_data):
        """Full-sort fair evaluation of the CURRENT stage's tables (pretrain: the raw embeddings; fine-tune: the tables
        filtered by all attributes, fairgo_pmf.py:250-257) with the fused evaluator: scoring + mask + top-K + the 12
        metrics of FairGo_PMF.yaml.  eval_data: evaluator.EvalData or a reference FullSortEvalDataLoader."""
        from .evaluator import EvalData, FullSortEvaluator
        from .sampled_eval import SampledEvalData, SampledEvaluator
        self.model.eval()
        if hasattr(eval_data, "resample"):          # uni<N> source that redraws its negatives per evaluation
            eval_data = eval_data.resample()
        if isinstance(eval_data, SampledEvalData):      # eval_args.mode uni<N> (the FairGo YAMLs' default): the same
            if getattr(self, "sampled_evaluator", None) is None:      # tables, scored on the candidate pairs only
                self.sampled_evaluator = SampledEvaluator(self.config, self.model.n_items,
                                                          getattr(self, "_train_item_count", None))
            U, I = self.model.filtered_tables()
            return self.sampled_evaluator.evaluate(
                SampledEvaluator.dot_scorer(U.detach(), I.detach(), self.model.max_rating), eval_data)
        if getattr(self, "evaluator", None) is None:
            self.evaluator = FullSortEvaluator(self.config, self.model.n_items, getattr(self, "_train_item_count", None))
        if not isinstance(eval_data, EvalData):
            cache = self.__dict__.setdefault("_eval_cache", {})
            if id(eval_data) not in cache:
                cache[id(eval_data)] = EvalData.from_reference_loader(eval_data, self.sst_attrs, self.model._dev())
            eval_data = cache[id(eval_data)]
        U, I = self.model.filtered_tables()
        return self.evaluator.evaluate(U.detach(), I.detach(), eval_data, self.model.max_rating)

    def fit(self, train_data, valid_data=None, epochs=None, train_item_count=None, verbose=False, saved=False):
        """pretrain (unless embeddings were given) + fine-tune with early stopping on `valid_metric`; train_data is any
        re-iterable of Interactions (user_id, item_id, rating, <sst>...).  Returns (best valid score, best valid result).
        saved=True check-points every improvement (checkpoint.py); after `resume_checkpoint` the loop continues at
        `start_epoch` with the restored early-stopping state."""
        from .trainer import early_stopping
        if train_item_count is not None:
            self.data_collect(train_item_count)
        resumed = getattr(self, "start_epoch", 0) > 0
        if self.model.train_stage == "pretrain" and not resumed:
            self.pretrain(train_data, valid_data=valid_data)
        self.model.train_stage = "finetune"
        metric = (self.config["valid_metric"] or "NDCG@5").lower()
        bigger = self.config["valid_metric_bigger"] if self.config["valid_metric_bigger"] is not None else True
        best, best_res, cur = (-np.inf if bigger else np.inf), None, 0
        if resumed:
            best, cur = self.best_valid_score, self.cur_step
        for epoch in range(getattr(self, "start_epoch", 0), epochs if epochs is not None else (self.config["epochs"] or 1)):
            dis_loss, filter_loss = self._train_epoch(train_data, epoch)
            if verbose:
                print(f"epoch {epoch}: filter loss {filter_loss:.4f}, discriminator loss {dis_loss:.4f}")
            if not valid_data:
                continue
            res = self.evaluate(valid_data)
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


FairGo_PMFTrainer = FairGo_GCNTrainer = FairGoTrainer
