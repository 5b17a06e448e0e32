#!/usr/bin/env python
"""Differential fuzz of the atomic-file ingestion against the LIVE reference (build container only):
random combinations of the `_data_filtering` options, orderings and splitting modes on the `messy` dataset of
make_test_data.py; ids and the three splits must be bit-identical, errors must coincide.
    python oracle/fuzz_ingest.py [option seed] [data seed] [trials]
TEST INFRASTRUCTURE ONLY.  (115 random option sets were run when this was written: 0 mismatches.)"""
import os
import random
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.dirname(HERE), HERE, os.path.join(HERE, 'ref_shim')):
    sys.path.insert(0, p)
import numpy as np
import shim; shim.install()
import yaml, torch
from recbole.config import Config
from recbole.data import create_dataset
from recbole.utils import init_seed as ref_seed
import make_test_data as mtd
from recbole_fairrec_b200.atomic import AtomicDataset
from recbole_fairrec_b200.quick_start import build_config, init_seed
import warnings; warnings.filterwarnings("ignore")
import logging; logging.disable(logging.CRITICAL)

def rand_opts(rng):
    o = {}
    lo = rng.choice([0,1,2,3,5,8]); hi = rng.choice(["inf", 30, 60, 200])
    o["user_inter_num_interval"] = rng.choice([None, f"[{lo},{hi})", f"({lo},{hi}]", f"[{lo},inf)"])
    lo = rng.choice([0,1,2,4,6]); hi = rng.choice(["inf", 40, 120])
    o["item_inter_num_interval"] = rng.choice([None, f"[{lo},{hi})", f"[{lo},{hi}]"])
    o["rm_dup_inter"] = rng.choice([None, "first", "last"])
    o["filter_inter_by_user_or_item"] = rng.choice([True, False])
    vi = {}
    if rng.random() < 0.4: vi["rating"] = rng.choice(["[2,5]", "(1,4]", "[1,3);[4,5]"])
    if rng.random() < 0.3: vi["timestamp"] = rng.choice(["[1000,1200)", "(1100,1399]"])
    if rng.random() < 0.3: vi["occupation"] = [f"occ{k}" for k in rng.sample(range(9), 5)]
    if rng.random() < 0.2: vi["genre"] = [f"g{k}" for k in rng.sample(range(5), 3)]
    if rng.random() < 0.2: vi["age"] = "[1,5]"
    o["val_interval"] = vi or None
    order = rng.choice(["RO", "TO"])
    split = rng.choice([{"RS":[8,1,1]}, {"RS":[7,2,1]}, {"RS":[6,2,2]}, {"LS":"valid_and_test"}, {"LS":"valid_only"}, {"LS":"test_only"}])
    group = "user" if "LS" in split else rng.choice(["user", "none"])
    o["eval_args"] = {"split": split, "group_by": group, "order": order, "mode": "full"}
    r = rng.random()
    if r < 0.25: o["normalize_all"] = True
    elif r < 0.55: o["normalize_field"] = rng.sample(["rating", "timestamp", "age", "gender", "occupation"], rng.choice([1, 2, 3]))
    if rng.random() < 0.15: o["benchmark_filename"] = rng.choice([["train", "valid", "test"], ["test", "train"], ["valid"]])
    return o

from recbole.data.dataset import Dataset
from recbole.utils import FeatureType
def _fill_nan(self):        # dataset.py:554-575 with assignments instead of `fillna(inplace=True)` (a no-op under pandas 3, which
    for feat_name in self.feat_name_list:      # would turn every normalised user column into NaN through the [PAD] row)
        feat = getattr(self, feat_name)
        for field in feat:
            ftype = self.field2type[field]
            if ftype == FeatureType.TOKEN: feat[field] = feat[field].fillna(value=0)
            elif ftype == FeatureType.FLOAT: feat[field] = feat[field].fillna(value=feat[field].mean())
            else:
                dtype = np.int64 if ftype == FeatureType.TOKEN_SEQ else float
                feat[field] = feat[field].apply(lambda x: np.array([], dtype=dtype) if isinstance(x, float) else x)
Dataset._fill_nan = _fill_nan

root = tempfile.mkdtemp(); name = mtd.write_messy(root, seed=int(sys.argv[2]) if len(sys.argv)>2 else 3)
os.chdir(tempfile.mkdtemp())
rng = random.Random(int(sys.argv[1]) if len(sys.argv)>1 else 0)
bad = 0
for trial in range(int(sys.argv[3]) if len(sys.argv)>3 else 25):
    opts = rand_opts(rng)
    cfgd = dict(mtd.INGEST_BASE, **opts)
    ref_err = ours_err = None
    try:
        with open("c.yaml","w") as f: yaml.safe_dump(dict(cfgd, data_path=root, use_gpu=False, state="CRITICAL", show_progress=False, neg_sampling=None, fair_objective="value"), f)
        sys.argv = sys.argv[:1]
        config = Config(model="FOCF", dataset=name, config_file_list=["c.yaml"])
        ref_seed(config["seed"], config["reproducibility"])
        ds_r = create_dataset(config); built = ds_r.build()
    except Exception as e:
        ref_err = f"{type(e).__name__}: {str(e)[:80]}"
    try:
        cfg = build_config("FOCF", name, None, dict(cfgd, data_path=root, device="cpu"))
        init_seed(cfg["seed"])
        ds = AtomicDataset(cfg); splits = ds.build()
    except Exception as e:
        ours_err = f"{type(e).__name__}: {str(e)[:80]}"
    if ref_err or ours_err:
        status = "both-error" if (ref_err and ours_err) else "MISMATCH-ERROR"
        if status != "both-error": bad += 1
        print(trial, status, "| ref:", ref_err, "| ours:", ours_err, "|", opts)
        continue
    ok = (ds.user_num, ds.item_num) == (ds_r.user_num, ds_r.item_num)
    why = "" if ok else f"nums {(ds.user_num, ds.item_num)} vs {(ds_r.user_num, ds_r.item_num)}"
    if ok:
        for k, part in zip(range(3), built):
            f_ = part.inter_feat
            for col in ("user_id","item_id","rating","timestamp","label"):
                a = splits[k][col]; b = f_[col].numpy() if len(f_) else np.zeros(0)
                if len(a) != len(b) or not np.array_equal(a, b):
                    ok = False; why = f"split {k} col {col} len {len(a)} vs {len(b)}"; break
            if not ok: break
    if ok:
        uf = ds_r.get_user_feature()
        for col in ("gender", "age", "occupation"):
            if not np.array_equal(ds.user_feat[col], uf[col].numpy()):
                ok = False; why = f"user feature {col}"; break
    if not ok:
        bad += 1
        print(trial, "MISMATCH", why, "|", opts)
    else:
        print(trial, "ok", ds.user_num, ds.item_num, [len(s['user_id']) for s in splits])
print("bad:", bad)
