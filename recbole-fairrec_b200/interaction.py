"""Minimal host-side batch container with the interface the hot path uses from the reference's
`Interaction` (recbole/data/interaction.py:43-200): a dict of equally long tensors with `[]`, `.to()`,
`len()`, `.columns` and row indexing.  The reference's own Interaction objects are accepted everywhere
this one is (duck typing); this class exists so the package runs without the reference installed."""
import numpy as np
import torch


class Interaction:
    def __init__(self, interaction):
        self.interaction = {}
        for k, v in interaction.items():
            if isinstance(v, np.ndarray):
                v = torch.from_numpy(v)
            elif isinstance(v, (list, tuple)):
                v = torch.as_tensor(v)
            elif not isinstance(v, torch.Tensor):
                raise ValueError(f"The type of {k}[{type(v)}] is not supported!")
            self.interaction[k] = v
        self.length = max((v.shape[0] for v in self.interaction.values()), default=0)

    def __getitem__(self, index):
        if isinstance(index, str):
            return self.interaction[index]
        return Interaction({k: v[index] for k, v in self.interaction.items()})

    def __contains__(self, item):
        return item in self.interaction

    def __iter__(self):
        return iter(self.interaction)

    def __len__(self):
        return self.length

    @property
    def columns(self):
        return list(self.interaction.keys())

    def to(self, device, selected_field=None):
        sel = set(self.interaction) if selected_field is None else (
            {selected_field} if isinstance(selected_field, str) else set(selected_field))
        out = Interaction({k: (v.to(device, non_blocking=True) if k in sel else v)
                           for k, v in self.interaction.items()})
        for attr in ("items_contiguous",):
            if hasattr(self, attr):
                setattr(out, attr, getattr(self, attr))
        return out

    def cpu(self):
        return self.to("cpu")

    def update(self, other):
        for k in other.interaction:
            self.interaction[k] = other.interaction[k]
