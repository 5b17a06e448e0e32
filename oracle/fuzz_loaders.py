#!/usr/bin/env python
"""Differential check of the train / evaluation loaders of `run_recbole` against the LIVE reference (build container only):
with the same seed, the pairwise (PFCN) and pointwise (NFCF) train batches -- in-place cumulative shuffles, uniform
negatives redrawn on collision -- and the `uni<N>` evaluation negatives (drawn anew at every evaluation) must come out
IDENTICAL to the reference's TrainDataLoader / NegSampleEvalDataLoader, pass after pass, because they make the same calls
on the same torch / numpy RNG streams.
    python oracle/fuzz_loaders.py [seed] [passes]
TEST INFRASTRUCTURE ONLY."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.dirname(HERE), HERE, os.path.join(HERE, "ref_shim")):
    sys.path.insert(0, p)
_args = sys.argv[1:]
sys.argv = sys.argv[:1]
import shim  # noqa: E402

shim.install()
import logging  # noqa: E402
import warnings  # noqa: E402

import torch  # noqa: E402
import yaml  # noqa: E402
from recbole.config import Config  # noqa: E402
from recbole.data import create_dataset, data_preparation  # noqa: E402
from recbole.utils import init_seed as ref_seed  # noqa: E402

import make_test_data as mtd  # noqa: E402
from recbole_fairrec_b200.atomic import AtomicDataset, used_and_positive_lists  # noqa: E402
from recbole_fairrec_b200.quick_start import BatchLoader, build_config, init_seed  # noqa: E402
from recbole_fairrec_b200.sampled_eval import AliasSampler, sample_negatives_reference  # noqa: E402

warnings.filterwarnings("ignore")
logging.disable(logging.CRITICAL)


def main():
    seed = int(_args[0]) if _args else 2020
    passes = int(_args[1]) if len(_args) > 1 else 3
    root = tempfile.mkdtemp()
    name = mtd.write_messy(root)
    os.chdir(tempfile.mkdtemp())
    bad = 0
    for model, pairwise, neg_num, bs, dist, by in (("PFCN_PMF", True, 20, 512, "uniform", 1), ("NFCF", False, 50, 300, "uniform", 1),
                                                   ("PFCN_PMF", True, 10, 700, "popularity", 1), ("NFCF", False, 30, 256, "popularity", 1),
                                                   ("PFCN_PMF", True, 5, 600, "uniform", 3), ("NFCF", False, 5, 500, "popularity", 4)):
        pop = dist == "popularity"
        base = dict(mtd.INGEST_BASE, **mtd.INGEST_CASES["defaults"], seed=seed, train_batch_size=bs,
                    user_inter_num_interval="[3,inf)")
        base["eval_args"] = dict(base["eval_args"], mode=f"{'pop' if pop else 'uni'}{neg_num}")
        extra = dict(filter_mode="none", dis_hidden_size_list=[8], activation="leakyrelu", embedding_size=8) if pairwise else \
            dict(mlp_hidden_size=[8], dropout=0.0, embedding_size=8, fair_weight=0.1)
        with open("c.yaml", "w") as f:
            yaml.safe_dump(dict(base, **extra, data_path=root, use_gpu=False, state="CRITICAL", show_progress=False,
                                neg_sampling={dist: by}, eval_batch_size=4096), f)
        config = Config(model=model, dataset=name, config_file_list=["c.yaml"])
        ref_seed(config["seed"], config["reproducibility"])
        train_r, valid_r, _ = data_preparation(config, create_dataset(config))
        ref_stream = []
        for _ in range(passes):
            for b in train_r:
                ref_stream.append(("train", b["user_id"].numpy().copy(), b["item_id"].numpy().copy(),
                                   (b["neg_item_id"] if pairwise else b["label"]).numpy().copy()))
            for cur, idx_list, pu, pi in valid_r:
                ref_stream.append(("valid", cur["user_id"].numpy().copy(), cur["item_id"].numpy().copy(), None))
        cfg = build_config(model, name, None, dict(base, **extra, data_path=root, device="cpu", neg_sampling={dist: by}))
        init_seed(cfg["seed"])
        ds = AtomicDataset(cfg)
        splits = ds.build()
        sampling = AliasSampler(np.concatenate([sp["item_id"] for sp in splits])).sampling if pop else None
        loader = BatchLoader(cfg, ds, splits[0], pairwise=pairwise, pointwise_neg=not pairwise, sampling=sampling)
        users, hist, pos = used_and_positive_lists(splits, "valid")
        mine = []
        for _ in range(passes):
            for b in loader:
                mine.append(("train", b["user_id"].numpy(), b["item_id"].numpy(),
                             (b["neg_item_id"] if pairwise else b["label"]).numpy()))
            neg = sample_negatives_reference(pos, hist, ds.item_num, neg_num, sampling)
            # the reference's evaluation batches: per user [positives ; negatives], several users per batch
            mine.append(("valid-all", np.concatenate([np.repeat(u, len(p) * (neg_num + 1)) for u, p in zip(users, pos)]),
                         np.concatenate([np.r_[p, q] for p, q in zip(pos, neg)]), None))
        # regroup the reference's valid batches of each pass into one record
        ref2, buf = [], None
        for rec in ref_stream:
            if rec[0] == "valid":
                buf = rec if buf is None else ("valid", np.r_[buf[1], rec[1]], np.r_[buf[2], rec[2]], None)
            else:
                if buf is not None:
                    ref2.append(("valid-all",) + buf[1:])
                    buf = None
                ref2.append(rec)
        if buf is not None:
            ref2.append(("valid-all",) + buf[1:])
        same = len(ref2) == len(mine)
        first = None
        for k, (a, b) in enumerate(zip(ref2, mine)):
            ok = a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and \
                (a[3] is None or np.array_equal(a[3], b[3]))
            if not ok and first is None:
                first = (k, a[0], len(a[1]), len(b[1]))
            same &= ok
        print(model, dist, by, "records", len(ref2), len(mine), "IDENTICAL" if same else f"DIFFERENT at {first}")
        bad += not same
    print("bad:", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
