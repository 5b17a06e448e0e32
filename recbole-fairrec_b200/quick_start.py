"""run_recbole(model, dataset, config_file_list, config_dict) -- the reference's entry point
(recbole/quick_start/quick_start.py:20-71, driven by run_recbole.py:16-26) over this package: YAML config ->
init_seed -> atomic-file dataset -> split -> device-resident loaders -> model -> trainer.fit -> trainer.evaluate(test).

Config precedence mirrors configurator.py:211-263 as far as the fairness configs need it: built-in defaults (config.py
DEFAULTS = overall.yaml / sample.yaml) < the model's hyper-parameter defaults (config.MODEL_DEFAULTS = the model / training
keys of properties/model/<Model>.yaml) < the YAML files of `config_file_list` (in order) < `config_dict` < `--key=value`
command-line overrides.

Negatives follow `neg_sampling: {uniform | popularity: n}` in training and `eval_args.mode: uni<N> | pop<N>` in evaluation
(popularity = the reference's alias table over the items of all interactions, sampler.py:72-120; same RNG calls, same draws).
Supported: FOCF (eval mode `full` or `uni<N>`), PFCN_MLP / PFCN_PMF / PFCN_BiasedMF / PFCN_DMF (pairwise batches with one
uniform negative per positive, `uni<N>` evaluation), FairGo_PMF / FairGo_GCN (pointwise batches, full-sort or `uni<N>`) and
NFCF (both stages: positives + uniform negatives with 1 | 0 labels, `uni<N>` evaluation; `saved=True` writes the stage-1
checkpoint that `load_pretrain_path` reads in stage 2).  With the
same seed the splits, the initial weights and FOCF's batch draws are identical to the reference's (tests/test_atomic.py,
tests/test_run_recbole_gpu.py)."""
import random
from logging import getLogger

import numpy as np
import torch
import yaml

from .atomic import AtomicDataset, used_and_positive_lists
from .config import Config
from .interaction import Interaction


def init_seed(seed, reproducibility=True):
    """recbole/utils/utils.py:172-189"""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.benchmark = not reproducibility
    torch.backends.cudnn.deterministic = bool(reproducibility)


def parse_argv_overrides(argv):
    """`--key=value` command-line overrides (configurator.py:167-172): values are read as YAML scalars / lists / dicts
    (the reference eval()s them, with the `value` pitfall of SURVEY.md section 5 (d); a YAML parse has the same results for
    numbers, booleans, lists and dicts, and keeps plain words as strings)"""
    out = {}
    for arg in argv or []:
        if not arg.startswith("--") or "=" not in arg:
            continue
        key, value = arg[2:].split("=", 1)
        out[key] = yaml.safe_load(value)
    return out


def build_config(model, dataset, config_file_list=None, config_dict=None, argv=None):
    from .config import MODEL_DEFAULTS
    merged = dict(MODEL_DEFAULTS.get(model, {}))
    for path in config_file_list or []:
        with open(path, "r", encoding="utf-8") as f:
            merged.update(yaml.safe_load(f) or {})
    merged.update(config_dict or {})
    merged.update(parse_argv_overrides(argv))
    merged["model"], merged["dataset"] = model, dataset
    cfg = Config(**merged)
    if "device" not in merged:
        cfg["device"] = torch.device("cuda" if torch.cuda.is_available() and cfg["use_gpu"] is not False else "cpu")
    return cfg


class BatchLoader:
    """TrainDataLoader (recbole/data/dataloader/general_dataloader.py:23-65) for the pointwise / pairwise families:
    shuffled fixed-size batches of the train split with the users' sensitive attributes joined and, for pairwise models,
    `neg_sampling: {uniform: n}` negatives (abstract_dataloader.py:182-188; uniform over the items the user has not
    interacted with in the train split, by rejection like sampler.py:145-197)."""

    def __init__(self, config, ds, split, pairwise, shuffle=True, pointwise_neg=False, sampling=None):
        self.cfg, self.ds, self.split, self.pairwise, self.shuffle = config, ds, split, pairwise, shuffle
        self.pointwise_neg = pointwise_neg
        self.batch_size = int(config["train_batch_size"])
        ns = config["neg_sampling"] or {}
        self.neg_num = int(next(iter(ns.values()))) if ns and (pairwise or pointwise_neg) else 1      # `by` (configurator.py:355-366)
        # general_dataloader.py:40-50: positives and their negatives together fill one train_batch_size
        times = self.neg_num if pairwise else (1 + self.neg_num if pointwise_neg else 1)
        self.batch_size = max(self.batch_size // times, 1)
        self.n = len(split[ds.uid_field])
        self._order = np.arange(self.n)
        self.attrs = [a for a in (config["sst_attr_list"] or []) if a in ds.user_feat]
        self.neg_prefix = config["NEG_PREFIX"] or "neg_"
        self.sampling = sampling or (lambda n: np.random.randint(1, self.ds.item_num, size=n))      # sampler.py:240-241
        if pairwise or pointwise_neg:
            key = split[ds.uid_field].astype(np.int64) * ds.item_num + split[ds.iid_field]
            self._used = np.sort(key)

    def __len__(self):
        return (self.n + self.batch_size - 1) // self.batch_size

    def _negatives(self, u):
        """sampler.py:145-197 for the users of one batch: `neg_num` draws per row, laid out draw-major (all first draws,
        then all second draws, ...), entries that hit an item of the user's train split redrawn together until none is left"""
        u = np.tile(u, self.neg_num)
        neg = self.sampling(len(u))
        while True:
            key = u.astype(np.int64) * self.ds.item_num + neg
            pos = np.searchsorted(self._used, key)
            bad = (pos < len(self._used)) & (self._used[np.minimum(pos, len(self._used) - 1)] == key)
            if not bad.any():
                return neg
            neg[bad] = self.sampling(int(bad.sum()))

    def __iter__(self):
        # the reference shuffles the train split IN PLACE at the start of every pass (abstract_dataloader.py:88-91 ->
        # interaction.py:293-297): each permutation applies to the order the previous pass left behind
        if self.shuffle:
            self._order = self._order[torch.randperm(self.n).numpy()]
        order = self._order
        uf, itf = self.ds.uid_field, self.ds.iid_field
        for b0 in range(0, self.n, self.batch_size):
            idx = order[b0:b0 + self.batch_size]
            u = self.split[uf][idx]
            cols = {k: torch.from_numpy(np.ascontiguousarray(v[idx])) for k, v in self.split.items()}
            for a in self.attrs:
                cols[a] = torch.from_numpy(self.ds.user_feat[a][u])
            if self.pairwise:                 # abstract_dataloader.py:190-198: rows repeated `neg_num` times + the negatives
                if self.neg_num > 1:
                    cols = {k: v.repeat(self.neg_num) for k, v in cols.items()}
                cols[self.neg_prefix + itf] = torch.from_numpy(self._negatives(u))
            if self.pointwise_neg:            # abstract_dataloader.py:200-208: rows repeated, items replaced, labels 1 | 0
                cols = {k: v.repeat(1 + self.neg_num) for k, v in cols.items()}
                cols[itf][len(u):] = torch.from_numpy(self._negatives(u))
                label = torch.zeros((1 + self.neg_num) * len(u))
                label[:len(u)] = 1.0
                cols[self.cfg["LABEL_FIELD"] or "label"] = label
            yield Interaction(cols)


def _load_best(trainer, saved):
    """quick_start.py:61 / trainer.py:476-482: the test split is evaluated with the best check-pointed model"""
    import os
    path = getattr(trainer, "saved_model_file", None)
    if saved and path and os.path.exists(path):
        trainer.resume_checkpoint(path)


def run_recbole(model=None, dataset=None, config_file_list=None, config_dict=None, saved=False, argv=None):
    """quick_start.py:20-71 -> {'best_valid_score', 'valid_score_bigger', 'best_valid_result', 'test_result'}"""
    import recbole_fairrec_b200 as pkg
    from .sampled_eval import ResamplingEvalSource, SampledEvalData, sample_negatives
    from .utils import get_model, get_trainer
    cfg = build_config(model, dataset, config_file_list, config_dict, argv)
    init_seed(cfg["seed"], cfg["reproducibility"] if cfg["reproducibility"] is not None else True)
    logger = getLogger()
    ds = AtomicDataset(cfg)
    splits = ds.build()
    dev = cfg["device"]
    uf, itf, rf = ds.uid_field, ds.iid_field, ds.rating_field
    attrs = list(cfg["sst_attr_list"] or [])
    train = splits[0]
    counts = np.bincount(train[itf], minlength=ds.item_num)
    item_counter = {int(i): int(c) for i, c in enumerate(counts) if c > 0}
    mode = str((cfg["eval_args"] or {}).get("mode", "full"))

    class TrainView:                         # what the models read from `train_data.dataset`
        num = staticmethod(ds.num)
        inter_feat = {k: torch.from_numpy(v) for k, v in train.items()}
        get_user_feature = staticmethod(ds.get_user_feature)
        inter_matrix = staticmethod(ds.inter_matrix)
        get_preload_weight = staticmethod(ds.get_preload_weight)

    sst_of_user = {a: ds.user_feat[a] for a in attrs}
    _alias = []

    def pop_sampling():        # one alias table over the items of all three splits (data/utils.py:244-265, sampler.py:234-238)
        if not _alias:
            from .sampled_eval import AliasSampler
            _alias.append(AliasSampler(np.concatenate([s[itf] for s in splits])))
        return _alias[0].sampling

    ns = cfg["neg_sampling"] or {}
    if ns and next(iter(ns)) not in ("uniform", "popularity"):
        raise ValueError(f"The distribution [{next(iter(ns))}] of neg_sampling should in ['uniform', 'popularity']")
    train_sampling = pop_sampling() if "popularity" in ns else None

    def eval_data(phase):
        if mode == "full" and cfg["eval_lists"] == "device":
            # large datasets: upload the split columns and group them on the device (EvalData.from_device: one sort per
            # side instead of per-user host lists).  Assumes unique (user, item) pairs, i.e. history and positives disjoint.
            t = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
            ev = splits[1] if phase == "valid" else splits[2]
            used = [splits[0]] + ([splits[1]] if phase == "test" else [])
            return pkg.EvalData.from_device(t(np.concatenate([s[uf] for s in used])), t(np.concatenate([s[itf] for s in used])),
                                            t(ev[uf]), t(ev[itf]), {a: t(v) for a, v in sst_of_user.items()},
                                            ds.user_num, ds.item_num)
        users, hist, pos = used_and_positive_lists(splits, phase)
        if mode == "full":
            return pkg.EvalData(users, hist, pos, sst_of_user, dev)
        if mode[:3] not in ("uni", "pop") or not mode[3:].isdigit():
            raise ValueError(f"the mode [{mode}] in eval_args is not supported.")       # configurator.py:380-390
        neg_num = int(mode[3:])
        if cfg["eval_neg_resample"] is False and mode[:3] == "uni":          # one fixed draw for all evaluations
            return SampledEvalData(users, pos, sample_negatives(pos, hist, ds.item_num, neg_num, np.random), sst_of_user, dev)
        # like the reference: negatives drawn anew at every evaluation, by the same calls on numpy's global RNG
        return ResamplingEvalSource(users, pos, hist, sst_of_user, ds.item_num, neg_num, dev,
                                    sampling=pop_sampling() if mode[:3] == "pop" else None)

    name = cfg["model"]
    if name == "FOCF":
        first = attrs[0]
        tdata = pkg.TrainData(train[uf], train[itf], train[rf], ds.user_feat[first].astype(np.float32), ds.user_num,
                              ds.item_num, dev, uf, itf, rf, first)
        loader = pkg.FOCFDataLoader(cfg, tdata, mode=cfg["focf_draw_mode"] or "reference")
        net = get_model(name)(cfg, TrainView).to(dev)
        trainer = get_trainer(net.type, name)(cfg, net)
        valid, test = eval_data("valid"), eval_data("test")      # negatives (uni<N>) drawn before training, like the
        best, best_res = trainer.fit(loader, valid, saved=saved, verbose=cfg["verbose"] is not False)   # reference's samplers
        test_res = trainer.evaluate(test, load_best_model=bool(saved))      # quick_start.py:61: the best model when saved
    elif name.startswith("PFCN_"):
        net = get_model(name)(cfg, TrainView).to(dev)
        trainer = get_trainer(net.type, name)(cfg, net)
        loader = BatchLoader(cfg, ds, train, pairwise=True, sampling=train_sampling)
        if mode == "full":
            raise NotImplementedError("PFCN full-sort evaluation is undefined in the reference; use eval_args.mode uni100")
        valid, test = eval_data("valid"), eval_data("test")
        best, best_res = trainer.fit(loader, valid, train_item_count=item_counter, saved=saved,
                                     verbose=cfg["verbose"] is not False)
        _load_best(trainer, saved)
        # several attributes: one evaluation per attribute subset, keyed like the reference's ('sm-[...]', trainer.py:1086);
        # a single subset comes back flat
        multi = net.filter_mode != "none" and len(trainer.attribute_subsets()) > 1
        test_res = trainer.evaluate_subsets(test, item_counter) if multi else trainer.evaluate(test, None, item_counter)
    elif name in ("FairGo_PMF", "FairGo_GCN"):
        net = get_model(name)(cfg, TrainView).to(dev)
        trainer = get_trainer(net.type, name)(cfg, net)
        # the FairGo YAMLs leave `neg_sampling: {uniform: 1}` in force: the reference's pointwise loader appends one sampled
        # item per interaction (same user, same rating column, label 0), abstract_dataloader.py:200-208
        loader = BatchLoader(cfg, ds, train, pairwise=False, pointwise_neg=cfg["neg_sampling"] is not None,
                             sampling=train_sampling)
        valid, test = eval_data("valid"), eval_data("test")
        best, best_res = trainer.fit(loader, valid, train_item_count=item_counter, saved=saved)
        _load_best(trainer, saved)
        test_res = trainer.evaluate(test)
    elif name == "NFCF":
        net = get_model(name)(cfg, TrainView).to(dev)
        trainer = get_trainer(net.type, name)(cfg, net)
        loader = BatchLoader(cfg, ds, train, pairwise=False, pointwise_neg=cfg["neg_sampling"] is not None,
                             sampling=train_sampling)
        if mode == "full":
            raise NotImplementedError("NFCF defines no full_sort_predict in the reference; use eval_args.mode uni100")
        valid, test = eval_data("valid"), eval_data("test")
        best, best_res = trainer.fit(loader, valid, saved=saved, train_item_count=item_counter,
                                     verbose=cfg["verbose"] is not False)
        _load_best(trainer, saved)
        test_res = trainer.evaluate(test, item_counter)
        if saved:
            logger.info("saved %s", trainer.saved_model_file)
    else:
        raise ValueError(f"unknown model {name}")
    bigger = cfg["valid_metric_bigger"]
    out = {"best_valid_score": best, "valid_score_bigger": True if bigger is None else bool(bigger), "best_valid_result": best_res, "test_result": test_res}
    if getattr(trainer, "saved_model_file", None):
        out["saved_model_file"] = trainer.saved_model_file
    return out
