"""Checkpoint / resume for the PFCN, FairGo and NFCF trainers -- the reference's file layout (trainer.py:221-240, 784-805,
1133-1184): {'config', 'epoch', 'cur_step', 'best_valid_score', 'state_dict', 'other_parameter', 'optimizer',
'optimizer_filter', 'optimizer_dis'}.

Reference quirk kept: the filter / discriminator MLPs live in plain dicts outside `state_dict()`, so the
reference's own keys do not hold them -- a resumed reference run restarts them from their random initialisation.  Their
weights are stored here under the additional key 'dict_modules' (ignored by the reference's loader, which reads keys by
name) and restored when present, so that a resume of THIS package continues the same trajectory."""
import os

import torch


def _cpu(obj):
    if torch.is_tensor(obj):
        return obj.detach().cpu()
    if isinstance(obj, dict):
        return {k: _cpu(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_cpu(v) for v in obj)
    return obj


class CheckpointMixin:
    """expects: self.config, self.model; optional self.optimizer / optimizer_filter / optimizer_dis / optimizer_pretrain
    (ops.AdamGroup), self.cur_step, self.best_valid_score, self.start_epoch"""

    _OPTIMIZERS = ("optimizer", "optimizer_filter", "optimizer_dis", "optimizer_pretrain")

    def _dict_modules(self):
        m = self.model
        out = {}
        for attr in ("filter_layer", "filter_layer_dict", "dis_layer_dict"):
            mods = getattr(m, attr, None)
            if isinstance(mods, dict):
                out[attr] = {k: _cpu(v.state_dict()) for k, v in mods.items()}
        return out

    def _save_checkpoint(self, epoch, saved_model_file=None):
        path = saved_model_file or getattr(self, "saved_model_file", None)
        if path is None:
            root = self.config["checkpoint_dir"] or "saved"
            path = self.saved_model_file = os.path.join(root, f"{self.config['model']}-{os.getpid()}.pth")
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        state = {"config": {k: (str(v) if isinstance(v, torch.device) else v) for k, v in dict(self.config).items()},
                 "epoch": epoch, "cur_step": getattr(self, "cur_step", 0),
                 "best_valid_score": getattr(self, "best_valid_score", None),
                 "state_dict": _cpu(self.model.state_dict()), "other_parameter": self.model.other_parameter(),
                 "dict_modules": self._dict_modules()}
        for name in self._OPTIMIZERS:
            opt = getattr(self, name, None)
            state[name] = _cpu(opt.state_dict()) if opt is not None else None
        torch.save(state, path)
        return path

    def resume_checkpoint(self, resume_file):
        ck = torch.load(str(resume_file), weights_only=False)
        self.saved_model_file = str(resume_file)
        self.start_epoch, self.cur_step = ck["epoch"] + 1, ck["cur_step"]
        self.best_valid_score = ck["best_valid_score"]
        self.model.load_state_dict(ck["state_dict"])
        self.model.load_other_parameter(ck.get("other_parameter"))
        for attr, mods in (ck.get("dict_modules") or {}).items():
            for k, sd in mods.items():
                getattr(self.model, attr)[k].load_state_dict(sd)
        for name in self._OPTIMIZERS:
            opt = getattr(self, name, None)
            if opt is not None and ck.get(name) is not None:
                opt.load_state_dict(ck[name])
        return ck
