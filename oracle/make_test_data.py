#!/usr/bin/env python
"""Trimmed copies of the bundled ml-100k atomic files for the end-to-end tests (tests/data/ml-100k/): only the columns
the FOCF config loads (inter: user_id, item_id, rating; user: user_id, gender; item: item_id), same row order and the
same `name:type` headers, so ids / shuffles / splits come out exactly as from the originals.
TEST INFRASTRUCTURE ONLY.  Source: /root/reference/recbole/dataset_example/ml-100k/ (MovieLens-100k data files)."""
import os

SRC = "/root/reference/recbole/dataset_example/ml-100k"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "data", "ml-100k")


def trim(name, keep):
    with open(os.path.join(SRC, name)) as f:
        header = f.readline().rstrip("\n").split("\t")
        idx = [i for i, h in enumerate(header) if h.split(":")[0] in keep]
        lines = ["\t".join(header[i] for i in idx)]
        for line in f:
            parts = line.rstrip("\n").split("\t")
            lines.append("\t".join(parts[i] for i in idx))
    with open(os.path.join(DST, name), "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    os.makedirs(DST, exist_ok=True)
    trim("ml-100k.inter", ("user_id", "item_id", "rating"))
    trim("ml-100k.user", ("user_id", "gender"))
    trim("ml-100k.item", ("item_id",))
    print({n: os.path.getsize(os.path.join(DST, n)) for n in os.listdir(DST)})


def write_messy(root, name="messy", seed=3, n_users=300, n_items=200, n_inter=6000):
    """A small atomic dataset exercising every filtering branch of the reference's Dataset (dataset.py:160-181): duplicated
    (user, item) pairs, tied timestamps, rows with a missing id, users / items absent from the feature files, feature rows
    without interactions, float and token attributes.  Deterministic; used by gen_golden.py (`ingest`) and
    tests/test_atomic.py so that the fixtures hold only the reference's OUTPUTS."""
    import numpy as np
    rng = np.random.default_rng(seed)
    d = os.path.join(root, name)
    os.makedirs(d, exist_ok=True)
    pu = rng.lognormal(0, 1.0, n_users)
    pi = rng.lognormal(0, 1.2, n_items)
    u = rng.choice(n_users, n_inter, p=pu / pu.sum())
    i = rng.choice(n_items, n_inter, p=pi / pi.sum())
    r = rng.integers(1, 6, n_inter)
    t = rng.integers(1000, 1400, n_inter)                   # many ties
    with open(os.path.join(d, name + ".inter"), "w") as f:
        f.write("user_id:token\titem_id:token\trating:float\ttimestamp:float\n")
        for k in range(n_inter):
            us = "" if k % 997 == 5 else f"u{u[k]}"
            it = "" if k in (1, 3) else f"i{i[k]}"      # before the first row without a user id: see atomic.data_filtering
            f.write(f"{us}\t{it}\t{r[k]}\t{t[k]}\n")
    in_file = np.sort(rng.choice(n_users, n_users - 20, replace=False))
    with open(os.path.join(d, name + ".user"), "w") as f:
        f.write("user_id:token\tgender:float\tage:float\toccupation:token\n")
        for k in rng.permutation(np.r_[in_file, np.arange(n_users, n_users + 15)]):
            f.write(f"u{k}\t{int(rng.integers(0, 2))}\t{int(rng.integers(0, 7))}\tocc{int(rng.integers(0, 9))}\n")
    keep_items = np.sort(rng.choice(n_items, n_items - 10, replace=False))
    with open(os.path.join(d, name + ".item"), "w") as f:
        f.write("item_id:token\tgenre:token\n")
        for k in rng.permutation(np.r_[keep_items, np.arange(n_items, n_items + 10)]):
            f.write(f"i{k}\tg{int(rng.integers(0, 5))}\n")
    # pretrained-embedding files of FairGo's `load_pretrain_weight` route (FairGo_PMF.yaml:12-17): id alias + float_seq
    for suf, idf, valf, prefix, ids in (("user_emb", "uid", "user_emb", "u", np.r_[rng.permutation(n_users)[:250], [900, 901, 902]]),
                                        ("item_emb", "iid", "item_emb", "i", np.r_[rng.permutation(n_items)[:170], [700, 701]])):
        with open(os.path.join(d, f"{name}.{suf}"), "w") as f:
            f.write(f"{idf}:token\t{valf}:float_seq\n")
            for k in ids:
                f.write(f"{prefix}{k}\t" + " ".join(f"{v:.4f}" for v in rng.standard_normal(8)) + "\n")
    # pre-split interaction files of `benchmark_filename: [train, valid, test]` (dataset.py:265-285): the complete rows of
    # the .inter file dealt out 8 : 1 : 1 by row index (no RNG draws: the files above stay what they were)
    parts = {p: open(os.path.join(d, f"{name}.{p}.inter"), "w") for p in ("train", "valid", "test")}
    for f in parts.values():
        f.write("user_id:token\titem_id:token\trating:float\ttimestamp:float\n")
    for k in range(n_inter):
        if k % 997 == 5 or k in (1, 3):
            continue
        parts["train" if k % 10 < 8 else ("valid" if k % 10 == 8 else "test")].write(f"u{u[k]}\ti{i[k]}\t{r[k]}\t{t[k]}\n")
    for f in parts.values():
        f.close()
    return name


INGEST_BASE = dict(RATING_FIELD="rating", LABEL_FIELD="label", threshold={"rating": 3.0}, seed=2020, TIME_FIELD="timestamp",
                   load_col={"inter": ["user_id", "item_id", "rating", "timestamp"],
                             "user": ["user_id", "gender", "age", "occupation"], "item": ["item_id", "genre"]},
                   sst_attr_list=["gender"])
PRELOAD = dict(additional_feat_suffix=["user_emb", "item_emb"], alias_of_user_id=["uid"], alias_of_item_id=["iid"],
               preload_weight={"uid": "user_emb", "iid": "item_emb"})
INGEST_CASES = {
    "preload": dict(PRELOAD, load_col=dict(INGEST_BASE["load_col"], user_emb=["uid", "user_emb"], item_emb=["iid", "item_emb"]),
                    eval_args={"split": {"RS": [8, 1, 1]}, "group_by": "user", "order": "RO", "mode": "full"}),
    "defaults": dict(eval_args={"split": {"RS": [8, 1, 1]}, "group_by": "user", "order": "RO", "mode": "full"}),
    "kcore_to_ls": dict(user_inter_num_interval="[5,inf)", item_inter_num_interval="[3,120]", rm_dup_inter="first",
                        val_interval={"rating": "[2,5]"},
                        eval_args={"split": {"LS": "valid_and_test"}, "group_by": "user", "order": "TO", "mode": "full"}),
    "nogroup_keepall": dict(filter_inter_by_user_or_item=False, rm_dup_inter="last", user_inter_num_interval="[1,inf)",
                            item_inter_num_interval="[1,inf)", val_interval={"occupation": ["occ1", "occ2", "occ3", "occ4", "occ5"]},
                            eval_args={"split": {"RS": [7, 2, 1]}, "group_by": "none", "order": "RO", "mode": "full"}),
    "ls_valid_only_to": dict(user_inter_num_interval="(2,60)",
                             eval_args={"split": {"LS": "valid_only"}, "group_by": "user", "order": "TO", "mode": "full"}),
    "ls_test_only_ro": dict(item_inter_num_interval="[2,inf)", val_interval={"timestamp": "[1000,1100);(1200,1399]"},
                            eval_args={"split": {"LS": "test_only"}, "group_by": "user", "order": "RO", "mode": "full"}),
    # min-max normalisation of chosen float fields (interaction and user side; the label is taken before it)
    "normalize_fields": dict(normalize_field=["rating", "age", "occupation"], rm_dup_inter="first",
                             eval_args={"split": {"RS": [8, 1, 1]}, "group_by": "user", "order": "TO", "mode": "full"}),
    # ... of every float field of the inter / user / item files -- the preloaded float_seq matrices are NOT touched
    "normalize_all": dict(PRELOAD, normalize_all=True,
                          load_col=dict(INGEST_BASE["load_col"], user_emb=["uid", "user_emb"], item_emb=["iid", "item_emb"]),
                          eval_args={"split": {"RS": [8, 1, 1]}, "group_by": "user", "order": "RO", "mode": "full"}),
    # pre-split files: no filtering, no shuffle, the parts come back as they are
    "benchmark_files": dict(benchmark_filename=["train", "valid", "test"], normalize_field=["timestamp"],
                            user_inter_num_interval="[50,inf)",        # ignored by the reference for benchmark files
                            eval_args={"split": {"RS": [8, 1, 1]}, "group_by": "user", "order": "RO", "mode": "full"}),
}
# cases generated with a copy-on-write-safe `_fill_nan` (SURVEY.md 8c caveat: under pandas 3 the reference's
# `fillna(inplace=True)` is a no-op, the [PAD] row keeps NaN and a min-max over the column turns ALL of it into NaN)
INGEST_PATCH_FILL_NAN = ("normalize_fields", "normalize_all")


FAMILY_E2E = {
    # dropout-free variants: the reference's Philox dropout masks cannot be reproduced, everything else (id remap, shuffle,
    # split, initial weights, in-place epoch shuffles, uniform negatives, per-evaluation negatives) follows from the seed
    "PFCN_PMF": dict(filter_mode="sm", dis_dropout=0.0, dis_weight=10, dis_hidden_size_list=[128, 256, 128, 128, 64, 32],
                     activation="leakyrelu", train_epoch_interval=1, weight_decay=0.0001, embedding_size=64),
    "NFCF": dict(mlp_hidden_size=[128, 64], dropout=0.0, fair_weight=0.1, weight_decay=1e-6, embedding_size=64,
                 load_pretrain_path=None),
}
FAMILY_E2E_BASE = dict(
    RATING_FIELD="rating", LABEL_FIELD="label", threshold={"rating": 3.0}, sst_attr_list=["gender"],
    load_col={"inter": ["user_id", "item_id", "rating"], "user": ["user_id", "gender"]}, neg_sampling={"uniform": 1},
    epochs=2, train_batch_size=2048, learning_rate=0.001, topk=[5], valid_metric="NDCG@5", seed=2020,
    metrics=["NDCG", "Recall", "Hit", "MRR", "GiniIndex", "PopularityPercentage"], metric_decimal_place=12,
    eval_args={"split": {"RS": [8, 1, 1]}, "group_by": "user", "order": "RO", "mode": "uni20"})


def float_gender_copy(dst_root):
    """tests/data/ml-100k with the gender token (M / F) as a 0 / 1 float column: what the PFCN / NFCF discriminators and
    regularisers need; tests/test_run_recbole_gpu.py builds the same files"""
    src = DST
    dst = os.path.join(dst_root, "ml-100k")
    os.makedirs(dst, exist_ok=True)
    import shutil
    for f in ("ml-100k.inter", "ml-100k.item"):
        shutil.copy(os.path.join(src, f), os.path.join(dst, f))
    lines = open(os.path.join(src, "ml-100k.user")).read().splitlines()
    with open(os.path.join(dst, "ml-100k.user"), "w") as f:
        f.write("\n".join(["user_id:token\tgender:float"] +
                          [l.split("\t")[0] + "\t" + ("1" if l.split("\t")[1] == "F" else "0") for l in lines[1:]]) + "\n")
    return dst_root


FOCF_UNI_E2E = dict(
    data_path=os.path.dirname(DST), RATING_FIELD="rating", LABEL_FIELD="label", threshold={"rating": 3.0},
    load_col={"inter": ["user_id", "item_id", "rating"], "user": ["user_id", "gender"], "item": ["item_id"]},
    sst_attr_list=["gender"], fair_objective="value", fair_weight=1.0, neg_sampling=None, weight_decay=0.001,
    learning_rate=0.001, embedding_size=64, epochs=2, train_batch_size=2048, topk=[5], valid_metric="NDCG@5", seed=2020,
    metrics=["NDCG", "Recall", "Hit", "MRR", "GiniIndex", "PopularityPercentage"], metric_decimal_place=12,
    eval_args={"split": {"RS": [8, 1, 1]}, "group_by": "user", "order": "RO", "mode": "uni20"})


FAIRGO_E2E = dict(
    RATING_FIELD="rating", LABEL_FIELD="label", threshold={"rating": 3.0}, sst_attr_list=["gender"],
    load_col={"inter": ["user_id", "item_id", "rating"], "user": ["user_id", "gender"]}, neg_sampling={"uniform": 1},
    load_pretrain_weight=False, aggr_method="LBA", vs_weights=[4, 1], embedding_size=64, n_layers=2,
    dis_hidden_size_list=[16, 8, 4], filter_hidden_size_list=[128, 64], activation="leakyrelu", fair_weight=0.1,
    pretrain_epochs=3, train_epoch_interval=1, weight_decay=0.0001, learning_rate=0.001, epochs=2, train_batch_size=2048,
    stopping_step=10, topk=[5], valid_metric="NDCG@5", seed=2020, metric_decimal_place=12,
    metrics=["NDCG", "Recall", "Hit", "MRR", "DifferentialFairness", "GiniIndex", "PopularityPercentage", "ValueUnfairness",
             "AbsoluteUnfairness", "UnderUnfairness", "OverUnfairness", "NonParityUnfairness"],
    eval_args={"split": {"RS": [8, 1, 1]}, "group_by": "user", "order": "RO", "mode": "full"})
