"""MLPLayers (recbole/model/layers.py:30-85) on this package's kernels.

The module tree is IDENTICAL to the reference's (`mlp_layers = nn.Sequential(Dropout, Linear, [BatchNorm1d],
[activation], ...)`), so parameter / buffer names and therefore state_dicts are interchangeable; only `forward` differs:
each (Dropout, Linear, BatchNorm1d?, activation?) group runs as fr_linear_forward (+ fr_batchnorm_forward) with the
activation fused into whichever kernel comes last."""
import torch.nn as nn
from torch.nn.init import normal_

from . import ops

_ACT_MODULES = {"sigmoid": nn.Sigmoid, "tanh": nn.Tanh, "relu": nn.ReLU, "leakyrelu": nn.LeakyReLU}


class MLPLayers(nn.Module):
    def __init__(self, layers, dropout=0.0, activation="relu", bn=False, init_method=None):
        super().__init__()
        self.layers, self.dropout, self.use_bn, self.init_method = list(layers), dropout, bn, init_method
        self.activation = None if activation is None or activation.lower() == "none" else activation.lower()
        if self.activation is not None and self.activation not in _ACT_MODULES:
            raise NotImplementedError(f"activation [{activation}] is not built into the kernels")
        mods = []
        for i, o in zip(self.layers[:-1], self.layers[1:]):
            mods.append(nn.Dropout(p=dropout))
            mods.append(nn.Linear(i, o))
            if bn:
                mods.append(nn.BatchNorm1d(num_features=o))
            if self.activation is not None:
                mods.append(_ACT_MODULES[self.activation]())
        self.mlp_layers = nn.Sequential(*mods)
        if init_method is not None:
            self.apply(self.init_weights)

    def init_weights(self, module):                      # layers.py:76-82
        if isinstance(module, nn.Linear):
            if self.init_method == "norm":
                normal_(module.weight.data, 0, 0.01)
            if module.bias is not None:
                module.bias.data.fill_(0.0)

    def forward(self, x):
        fused = ops.mlp_chain([self], [x]) if x.dim() == 2 else None     # the whole chain in one launch (mlp_chain.cu)
        if fused is not None:
            return fused[0]
        if ops._chain_dp is not None and self.use_bn and self.training:
            raise NotImplementedError("data-parallel BatchNorm needs the fused chain kernels (layer widths <= 256, K % 4 == 0)")
        act = ops.ACT[self.activation]
        mods = list(self.mlp_layers)
        k = 0
        while k < len(mods):
            lin = mods[k + 1]
            k += 2
            bn = None
            if k < len(mods) and isinstance(mods[k], nn.BatchNorm1d):
                bn = mods[k]
                k += 1
            if k < len(mods) and not isinstance(mods[k], nn.Dropout):
                k += 1                                   # the activation module (fused)
            p = self.dropout if self.training else 0.0
            x = ops.LinearAct.apply(x, lin.weight, lin.bias, 0 if bn is not None else act, p, ops.next_seed())
            if bn is not None:
                x = ops.BatchNormAct.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum, bn.eps,
                                           self.training, act)
                if self.training:
                    bn.num_batches_tracked += 1
        return x
