"""NFCF (NCF tower + differential-fairness regulariser) -- drop-in for recbole/model/fair_recommender/nfcf.py:17-115.

Same plugin surface and the same parameter names as the reference (`user_embedding.weight`, `item_embedding.weight`,
`mlp_layers.mlp_layers.{1,4,7,...}.weight/bias`, so reference checkpoints load, including the pre-trained NCF that
`reset_params` de-biases), but forward, loss and backward run in this package's kernels (`fr_nfcf_forward` /
`fr_nfcf_backward`): embedding gather + concat, the tower's fused Linear+ReLU GEMMs, sigmoid + BCE, the item x group
sums of the regulariser over the batch's positives (sorted-segment reduction), and the dense embedding gradients.
`calculate_loss` is a torch.autograd.Function, so the reference's own `Trainer` (torch.optim.Adam over
`model.parameters()`) drives it unchanged.  CUDA only -- no CPU fallback.
"""
import ctypes

import torch
import torch.nn as nn

from .checkpoint import CheckpointMixin
from .utils import stopping_step as _stopping_step
from . import _lib
from ._lib import MlpTower, NfcfStep, check, load, ptr, stream_ptr

ACT = {"none": 0, "relu": 1, "leakyrelu": 2, "sigmoid": 3, "tanh": 4}


def _mlp_modules(layers, dropout):
    """the module layout of recbole/model/layers.py:58-70 (parameter holder; compute happens in the kernels)"""
    mods = []
    for i, o in zip(layers[:-1], layers[1:]):
        mods += [nn.Dropout(p=dropout), nn.Linear(i, o), nn.ReLU()]
    return nn.Sequential(*mods)


class _Tower(nn.Module):
    def __init__(self, layers, dropout):
        super().__init__()
        self.layers = layers
        self.mlp_layers = _mlp_modules(layers, dropout)

    def linears(self):
        return [m for m in self.mlp_layers if isinstance(m, nn.Linear)]


class _NfcfLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, batch, U, I, *wb):
        s, keep = model._step(U.detach(), I.detach(), [t.detach() for t in wb], batch)
        check(model._lib.fr_nfcf_forward(ctypes.byref(s), stream_ptr()), "fr_nfcf_forward")
        ctx.model, ctx.s, ctx.keep, ctx.shapes = model, s, keep, (U, I, wb)
        return keep["loss"].view(())

    @staticmethod
    def backward(ctx, grad_out):
        U, I, wb = ctx.shapes
        s, model = ctx.s, ctx.model
        need_u = U.requires_grad
        dU = torch.empty_like(U) if need_u else None
        dI = torch.empty_like(I)
        grads = [torch.empty_like(t) for t in wb]
        s.dU, s.dI = ptr(dU), ptr(dI)
        L = len(wb) // 2
        for l in range(L):
            s.dW[l], s.db[l] = ptr(grads[2 * l]), ptr(grads[2 * l + 1])
        check(model._lib.fr_nfcf_backward(ctypes.byref(s), float(grad_out), stream_ptr()), "fr_nfcf_backward")
        return (None, None, dU, dI, *grads)


class NFCF(nn.Module):
    input_type = "POINTWISE"
    type = "GENERAL"

    def __init__(self, config, dataset):
        super().__init__()
        self.USER_ID, self.ITEM_ID = config["USER_ID_FIELD"], config["ITEM_ID_FIELD"]
        self.n_users, self.n_items = dataset.num(self.USER_ID), dataset.num(self.ITEM_ID)
        self.device = config["device"]
        self.LABEL = config["LABEL_FIELD"]                                   # nfcf.py:27-36
        self.embedding_size = config["embedding_size"]
        self.mlp_hidden_size = list(config["mlp_hidden_size"])
        self.dropout = float(config["dropout"] or 0.0)
        self.sst_attr = config["sst_attr_list"][0]
        self.fair_weight = float(config["fair_weight"])
        self.load_pretrain_path = config["load_pretrain_path"]
        if self.embedding_size % 4 != 0:
            raise ValueError("fairrec_b200 NFCF needs embedding_size to be a multiple of 4")
        self.user_embedding = nn.Embedding(self.n_users, self.embedding_size)   # nfcf.py:39-44
        self.item_embedding = nn.Embedding(self.n_items, self.embedding_size)
        self.mlp_layers = _Tower([2 * self.embedding_size] + self.mlp_hidden_size + [1], self.dropout)
        self._lib = load()
        self._ws = None
        self._flags = None
        self._calls = 0
        if self.load_pretrain_path is not None:
            self.reset_params(self.load_pretrain_path, dataset.get_user_feature()[1:])

    def reset_params(self, pretrain_path, user_data):
        """nfcf.py:49-67: load the pre-trained NCF, project the gender direction out of the user table, freeze it,
        re-initialise the item table.  One-off setup on the host side of the path (plain torch)."""
        checkpoint = torch.load(pretrain_path, weights_only=False)
        self.load_state_dict(checkpoint["state_dict"], strict=False)
        sst = user_data[self.sst_attr]
        vals = torch.unique(sst)
        emb = self.user_embedding.weight.data[1:].clone()
        e1, e2 = emb[sst == vals[0]].mean(dim=0), emb[sst == vals[1]].mean(dim=0)
        bias = (e1 - e2) / torch.linalg.norm(e1 - e2, keepdim=True)
        self.user_embedding.weight.data[1:] = emb - torch.mul(emb, bias).sum(dim=1, keepdim=True) * bias
        self.user_embedding.weight.requires_grad = False
        self.item_embedding = nn.Embedding(self.n_items, self.embedding_size)

    # ------------------------------------------------------------------ plumbing
    def _step(self, U, I, wb, batch):
        uid, iid, label, sst = batch
        if not U.is_cuda:
            raise _lib.FairRecLibraryError("NFCF (fairrec_b200) runs on CUDA only: move the model to a cuda device")
        M = uid.numel()
        s = NfcfStep()
        s.U, s.I, s.n_users, s.n_items, s.d = ptr(U), ptr(I), self.n_users, self.n_items, self.embedding_size
        s.uid, s.iid, s.label, s.sst, s.M = ptr(uid), ptr(iid), ptr(label), ptr(sst), M
        t = s.tower
        layers = self.mlp_layers.layers
        t.n_layers = len(layers) - 1
        for k, v in enumerate(layers):
            t.dims[k] = v
        for l in range(t.n_layers):
            t.W[l], t.b[l] = ptr(wb[2 * l].contiguous()), ptr(wb[2 * l + 1].contiguous())
        t.act, t.dropout = ACT["relu"], self.dropout
        s.use_df = 1 if (self.load_pretrain_path is not None and sst is not None) else 0
        s.fair_weight = self.fair_weight
        s.training = 1 if self.training else 0
        self._calls += 1
        s.seed = (torch.initial_seed() * 1000003 + self._calls) & 0xFFFFFFFFFFFFFFFF
        need = self._lib.fr_nfcf_workspace_bytes(ctypes.byref(t), M)
        if self._ws is None or self._ws.numel() < need or self._ws.device != U.device:
            self._ws = torch.empty(int(need * 1.25) + 256, dtype=torch.uint8, device=U.device)
            self._flags = torch.zeros(1, dtype=torch.int32, device=U.device)
        keep = {"loss": torch.empty(1, dtype=torch.float32, device=U.device),
                "pred": torch.empty(M, dtype=torch.float32, device=U.device), "batch": batch, "wb": wb}
        s.pred, s.loss, s.status_flags = ptr(keep["pred"]), ptr(keep["loss"]), ptr(self._flags)
        s.workspace, s.workspace_bytes = ptr(self._ws), self._ws.numel()
        return s, keep

    def _batch(self, interaction, need_label=True):
        dev = self.user_embedding.weight.device
        col = lambda name, dt: interaction[name].to(device=dev, dtype=dt, non_blocking=True).contiguous()
        uid, iid = col(self.USER_ID, torch.int32), col(self.ITEM_ID, torch.int32)
        label = col(self.LABEL, torch.float32) if need_label and self.LABEL in interaction else \
            torch.zeros(uid.numel(), dtype=torch.float32, device=dev)
        sst = col(self.sst_attr, torch.float32) if self.sst_attr in interaction else None
        return uid, iid, label, sst

    def _params(self):
        return [p for lin in self.mlp_layers.linears() for p in (lin.weight, lin.bias)]

    def other_parameter(self):
        return dict()

    def load_other_parameter(self, para):
        return

    # ------------------------------------------------------------------ reference API
    def calculate_loss(self, interaction):
        """nfcf.py:99-110"""
        return _NfcfLoss.apply(self, self._batch(interaction), self.user_embedding.weight, self.item_embedding.weight,
                               *self._params())

    @torch.no_grad()
    def predict(self, interaction):
        """nfcf.py:112-115"""
        was = self.training
        self.training = False
        try:
            s, keep = self._step(self.user_embedding.weight.data, self.item_embedding.weight.data,
                                 [p.data for p in self._params()], self._batch(interaction, need_label=False))
            s.use_df = 0
            check(self._lib.fr_nfcf_forward(ctypes.byref(s), stream_ptr()), "fr_nfcf_forward")
        finally:
            self.training = was
        return keep["pred"]

    def forward(self, user, item):
        """nfcf.py:69-74"""
        from .interaction import Interaction
        return self.predict(Interaction({self.USER_ID: user, self.ITEM_ID: item}))

    def check_flags(self):
        if self._flags is not None and int(self._flags.item()) & _lib.FLAG_TOO_MANY_GROUPS:
            self._flags.zero_()
            raise NotImplementedError("NFCF regulariser: more than 32 distinct sensitive-attribute values among the positives "
                                      "of a batch (the kernel keeps one lane per group)")


class NFCFTrainer(CheckpointMixin):
    """The base `Trainer` (recbole/trainer/trainer.py:100-260, the one `get_trainer` hands NFCF: utils.py:74-94) for the
    two NFCF stages: stage 1 (`load_pretrain_path: ~`) trains the plain NCF tower and `fit(saved=True)` writes the
    `{'state_dict': ...}` checkpoint stage 2 reads (nfcf.py:49-51); stage 2 (path given) fine-tunes with the user table
    frozen and the differential-fairness regulariser on.  Optimizer: this package's Adam kernel (ops.AdamGroup = the
    reference's torch.optim.Adam, L2 form) over the parameters that require a gradient; evaluation: the sampled-negative
    (`uni<N>`) evaluator with `model.predict` as the scorer, as NFCF.yaml configures it."""

    def __init__(self, config, model):
        from . import ops
        self.config, self.model = config, model
        self.optimizer = ops.AdamGroup([p for p in model.parameters() if p.requires_grad],
                                       lr=config["learning_rate"] or 1e-3, weight_decay=config["weight_decay"] or 0.0)
        self.sampled_evaluator = None
        self.saved_model_file = None

    def _train_epoch(self, train_data, epoch_idx):
        """trainer.py:160-199: sum of the batch losses of the epoch"""
        self.model.train()
        total = None
        for interaction in train_data:
            self.optimizer.zero_grad()
            loss = self.model.calculate_loss(interaction)
            total = loss.detach() if total is None else total + loss.detach()
            loss.backward()
            self.optimizer.step()
        self.model.check_flags()
        value = float(total.item()) if total is not None else 0.0
        if value != value:
            raise ValueError("Training loss is nan")
        return value

    @torch.no_grad()
    def evaluate(self, eval_data, train_item_count=None):
        from .interaction import Interaction
        from .sampled_eval import SampledEvaluator
        self.model.eval()
        if hasattr(eval_data, "resample"):          # uni<N> source that redraws its negatives per evaluation
            eval_data = eval_data.resample()
        if self.sampled_evaluator is None:
            self.sampled_evaluator = SampledEvaluator(self.config, self.model.n_items, train_item_count)
        m = self.model
        return self.sampled_evaluator.evaluate(
            lambda uid, iid: m.predict(Interaction({m.USER_ID: uid, m.ITEM_ID: iid})).view(-1), eval_data)

    def fit(self, train_data, valid_data=None, saved=False, train_item_count=None, verbose=False):
        """trainer.py:300-380: early stopping on `valid_metric`; the best model is check-pointed when saved=True (the
        reference's checkpoint layout, checkpoint.py: its 'state_dict' is what stage 2 loads, nfcf.py:49-51)"""
        import os
        from .trainer import early_stopping
        metric = (self.config["valid_metric"] or "NDCG@5").lower()
        bigger = self.config["valid_metric_bigger"] if self.config["valid_metric_bigger"] is not None else True
        self.best_valid_score, best_res, self.cur_step = (-float("inf") if bigger else float("inf")), None, 0
        if saved:
            root = self.config["checkpoint_dir"] or "saved"
            os.makedirs(root, exist_ok=True)
            self.saved_model_file = os.path.join(root, f"{self.config['model']}-{os.getpid()}.pth")
        for epoch in range(getattr(self, "start_epoch", 0), self.config["epochs"] or 1):
            loss = self._train_epoch(train_data, epoch)
            if verbose:
                print(f"epoch {epoch}: train loss {loss:.4f}")
            if not valid_data:
                if saved:
                    self._save_checkpoint(epoch)
                continue
            res = self.evaluate(valid_data, train_item_count)
            self.best_valid_score, self.cur_step, stop, update = early_stopping(
                res[metric], self.best_valid_score, self.cur_step, max_step=_stopping_step(self.config), bigger=bigger)
            if update:
                best_res = res
                if saved:
                    self._save_checkpoint(epoch)
            if stop:
                break
        return self.best_valid_score, best_res
