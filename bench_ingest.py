"""Dataset ingestion at the ML-1M shape: atomic files -> id remap -> labels -> RO shuffle -> per-user
RS split -> evaluation lists, host side only (no GPU needed).

    python bench_ingest.py                 # this package's ingestion (recbole_fairrec_b200.atomic), one JSON line
    python bench_ingest.py --reference     # + the unmodified reference's create_dataset / data_preparation on the same
                                           #   files (needs /root/reference: build container only), and a bit-for-bit
                                           #   comparison of the three splits and the evaluation lists

The files are synthetic (6040 users, 3706 items, 1,000,209 unique pairs, ML-1M's rating marginal;
gender / age / occupation as 0/1-or-small float columns like dataset/ml-1M/ml-1M.user) and are written to a temp dir."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def write_files(root, name="ml-1M-synth", n_users=6041, n_items=3707, n_inter=1_000_209, seed=2020):
    from recbole_fairrec_b200 import synth
    os.makedirs(os.path.join(root, name), exist_ok=True)
    uid, iid, rating, gender = synth.interactions(n_users, n_items, n_inter, seed)
    rng = np.random.default_rng(seed + 1)
    order = rng.permutation(len(uid))                      # file order is arbitrary, like the raw MovieLens dump
    ts = rng.integers(956_703_932, 1_046_454_590, len(uid))
    inter_path = os.path.join(root, name, name + ".inter")
    try:                                   # pyarrow's writer: seconds instead of minutes at 1e7+ rows
        import pyarrow as pa
        import pyarrow.csv as pc
        tab = pa.table({"u": uid[order], "i": iid[order], "r": rating[order].astype(np.int64), "t": ts})
        with open(inter_path, "wb") as f:
            f.write(b"user_id:token\titem_id:token\trating:float\ttimestamp:float\n")
            pc.write_csv(tab, f, pc.WriteOptions(include_header=False, delimiter="\t"))
    except ImportError:
        with open(inter_path, "w") as f:
            f.write("user_id:token\titem_id:token\trating:float\ttimestamp:float\n")
            np.savetxt(f, np.stack([uid[order], iid[order], rating[order].astype(np.int64), ts], 1), fmt="%d", delimiter="\t")
    with open(os.path.join(root, name, name + ".user"), "w") as f:
        f.write("user_id:token\tgender:float\tage:float\toccupation:float\n")
        users = np.arange(1, n_users)
        np.savetxt(f, np.stack([users, (np.asarray(gender)[1:] == 2).astype(np.int64), rng.integers(0, 7, n_users - 1),
                                rng.integers(0, 21, n_users - 1)], 1), fmt="%d", delimiter="\t")
    with open(os.path.join(root, name, name + ".item"), "w") as f:
        f.write("item_id:token\n")
        np.savetxt(f, np.arange(1, n_items), fmt="%d")
    return name


CFG = dict(RATING_FIELD="rating", LABEL_FIELD="label", threshold={"rating": 3.0}, seed=2020,
           load_col={"inter": ["user_id", "item_id", "rating"], "user": ["user_id", "gender", "age", "occupation"],
                     "item": ["item_id"]},
           sst_attr_list=["gender", "age", "occupation"],
           eval_args={"split": {"RS": [8, 1, 1]}, "group_by": "user", "order": "RO", "mode": "full"})


def ours(root, name):
    from recbole_fairrec_b200.atomic import AtomicDataset, used_and_positive_lists
    from recbole_fairrec_b200.quick_start import build_config, init_seed
    import pandas  # noqa: F401  (module imports stay outside the timed region in both arms)
    t = {}
    t0 = time.perf_counter()
    cfg = build_config("FOCF", name, None, dict(CFG, data_path=root, device="cpu"))
    init_seed(cfg["seed"])
    ds = AtomicDataset(cfg)
    t["load_remap_s"] = time.perf_counter() - t0
    t1 = time.perf_counter()
    splits = ds.build()
    t["shuffle_split_s"] = time.perf_counter() - t1
    t2 = time.perf_counter()
    lists = {ph: used_and_positive_lists(splits, ph) for ph in ("valid", "test")}
    t["eval_lists_s"] = time.perf_counter() - t2
    t["total_s"] = time.perf_counter() - t0
    return t, ds, splits, lists


def reference(root, name):
    sys.path.insert(0, os.path.join(HERE, "oracle", "ref_shim"))
    import shim
    shim.install()
    import yaml
    from recbole.config import Config
    from recbole.data import create_dataset, data_preparation
    from recbole.utils import init_seed
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    try:
        with open("c.yaml", "w") as f:
            yaml.safe_dump(dict(CFG, data_path=root, use_gpu=False, state="WARNING", show_progress=False, neg_sampling=None,
                                fair_objective="value"), f)
        sys.argv = sys.argv[:1]
        t = {}
        t0 = time.perf_counter()
        config = Config(model="FOCF", dataset=name, config_file_list=["c.yaml"])
        init_seed(config["seed"], config["reproducibility"])
        dataset = create_dataset(config)
        t["create_dataset_s"] = time.perf_counter() - t0
        t1 = time.perf_counter()
        train, valid, test = data_preparation(config, dataset)
        t["data_preparation_s"] = time.perf_counter() - t1
        t["total_s"] = time.perf_counter() - t0
        return t, dataset, (train, valid, test)
    finally:
        os.chdir(cwd)


def compare(ds, splits, lists, rds, rloaders):
    """bit-for-bit: id counts, the three splits (train: after the reference loader's stable item sort), user features,
    the evaluation lists (positives as sets: the reference stores them in Python-set iteration order)"""
    assert (ds.user_num, ds.item_num) == (rds.user_num, rds.item_num)
    tr = rloaders[0].dataset.inter_feat
    o = np.argsort(splits[0]["item_id"], kind="stable")
    for f in ("user_id", "item_id", "rating", "label"):
        np.testing.assert_array_equal(splits[0][f][o], tr[f].numpy())
    for k, loader in ((1, rloaders[1]), (2, rloaders[2])):
        ev = loader.dataset.inter_feat
        o = np.argsort(splits[k]["user_id"], kind="stable")
        for f in ("user_id", "item_id", "rating"):
            np.testing.assert_array_equal(splits[k][f][o], ev[f].numpy())
    uf = rds.get_user_feature()
    for a in ("gender", "age", "occupation"):
        np.testing.assert_array_equal(ds.user_feat[a][1:], uf[a].numpy()[1:])
    for ph, loader in (("valid", rloaders[1]), ("test", rloaders[2])):
        users, hist, pos = lists[ph]
        np.testing.assert_array_equal(users, np.asarray(loader.uid_list))
        for r in range(0, len(users), 97):
            u = int(users[r])
            assert set(pos[r].tolist()) == set(loader.uid2positive_item[u].tolist())
            assert set(hist[r].tolist()) == set(loader.uid2history_item[u].tolist())
    return True


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", action="store_true")
    ap.add_argument("--out", default=None)
    ap.add_argument("--rows", type=int, default=1_000_209)
    ap.add_argument("--users", type=int, default=6041)
    ap.add_argument("--items", type=int, default=3707)
    a = ap.parse_args()
    root = tempfile.mkdtemp()
    name = write_files(root, n_users=a.users, n_items=a.items, n_inter=a.rows)
    t, ds, splits, lists = ours(root, name)
    shape = "ML-1M shape" if a.rows == 1_000_209 else f"{a.rows} rows x {a.users - 1} users x {a.items - 1} items"
    line = {"metric": f"dataset ingestion (atomic files -> splits + eval lists), {shape}", "unit": "s",
            "higher_is_better": False, "value": round(t["total_s"], 3), "phases": {k: round(v, 3) for k, v in t.items()},
            "rows": int(len(ds)), "rows_per_s": round(len(ds) / t["total_s"]), "n_users": ds.user_num, "n_items": ds.item_num,
            "cores": os.cpu_count()}
    if a.reference:
        rt, rds, rloaders = reference(root, name)
        line["reference"] = {k: round(v, 3) for k, v in rt.items()}
        line["identical_to_reference"] = compare(ds, splits, lists, rds, rloaders)
    s = json.dumps(line)
    print(s)
    if a.out:
        with open(a.out, "w") as f:
            f.write(s + "\n")


if __name__ == "__main__":
    main()
