"""CPU: the atomic-file front end (atomic.py) reproduces the reference's data pipeline bit for bit -- id assignment,
RO shuffle, 8:1:1 split grouped by user, token ids of the sensitive attribute, model initialisation and FOCF's batch
draws -- against tensors exported from the UNMODIFIED reference's run on the bundled ml-100k
(tests/golden/ml100k_focf_value.npz; the data files under tests/data/ml-100k are column-trimmed copies)."""
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(__file__)
G = os.path.join(HERE, "golden", "ml100k_focf_value.npz")


def make(device=torch.device("cpu")):
    import recbole_fairrec_b200 as pkg
    from recbole_fairrec_b200.quick_start import init_seed
    g = np.load(G)
    cfg = pkg.Config(dataset="ml-100k", data_path=os.path.join(HERE, "data"), threshold={"rating": 3.0},
                     load_col={"inter": ["user_id", "item_id", "rating"], "user": ["user_id", "gender"], "item": ["item_id"]},
                     eval_args={"split": {"RS": [8, 1, 1]}, "group_by": "user", "order": "RO", "mode": "full"},
                     embedding_size=64, device=device, train_batch_size=int(g["train_batch_size"]), seed=2020)
    init_seed(2020)
    ds = pkg.AtomicDataset(cfg)
    return g, cfg, ds, ds.build()


def test_ids_split_and_attribute_match_the_reference():
    g, cfg, ds, (tr, va, te) = make()
    assert (ds.user_num, ds.item_num, len(ds)) == (int(g["n_users"]), int(g["n_items"]), 100000)
    assert np.array_equal(ds.user_feat["gender"], g["gender"])            # token ids by first appearance: M=1, F=2
    o = np.argsort(tr["item_id"], kind="stable")                          # FOCFDataLoader sorts the split by item
    assert np.array_equal(tr["user_id"][o], g["train_u"]) and np.array_equal(tr["item_id"][o], g["train_i"])
    assert np.array_equal(tr["rating"][o], g["train_r"])
    for mine, (gu, gi) in ((va, ("valid_u", "valid_i")), (te, ("test_u", "test_i"))):
        o = np.argsort(mine["user_id"], kind="stable")                    # the eval loader sorts by user
        assert np.array_equal(mine["user_id"][o], g[gu]) and np.array_equal(mine["item_id"][o], g[gi])
    assert np.array_equal(tr["label"], (tr["rating"] >= 3).astype(np.float32))


def test_model_init_and_batch_draws_match_the_reference():
    import recbole_fairrec_b200 as pkg
    g, cfg, ds, (tr, va, te) = make()

    class View:
        num = staticmethod(ds.num)
        inter_feat = {"rating": torch.from_numpy(tr["rating"])}

    model = pkg.FOCF(cfg, View)
    assert np.array_equal(model.user_embedding_layer.weight.detach().numpy(), g["U0"])
    assert np.array_equal(model.item_embedding_layer.weight.detach().numpy(), g["I0"])
    train = pkg.TrainData(tr["user_id"], tr["item_id"], tr["rating"], ds.user_feat["gender"].astype(np.float32),
                          ds.user_num, ds.item_num, torch.device("cpu"))
    loader = pkg.FOCFDataLoader(cfg, train, mode="reference")
    draws = []
    for _ in range(len(loader)):
        draws += list(loader._draw_batch()) + [-1]
    assert np.array_equal(np.array(draws), g["draws"][:len(draws)])


def test_split_counts_rule():
    from recbole_fairrec_b200.atomic import calcu_split_counts
    # dataset.py:1339-1360: parts other than the first are rounded down, but a part that would be empty borrows one
    got = calcu_split_counts(np.array([1, 2, 5, 9, 10, 20, 23]), [8, 1, 1])
    want = []
    for tot in [1, 2, 5, 9, 10, 20, 23]:
        r = [0.8, 0.1, 0.1]
        cnt = [int(r[i] * tot) for i in range(3)]
        cnt[0] = tot - sum(cnt[1:])
        for i in range(1, 3):
            if cnt[0] <= 1:
                break
            if 0 < r[-i] * tot < 1:
                cnt[-i] += 1
                cnt[0] -= 1
        want.append(cnt)
    assert got.tolist() == want


def test_pointwise_loader_fills_batches_with_positives_and_unseen_negatives():
    """abstract_dataloader.py:200-208 + general_dataloader.py:40-50: B/2 positives + B/2 negatives, labels 1 | 0, negatives
    never among the user's train items, sensitive attribute joined for both halves"""
    from recbole_fairrec_b200.quick_start import BatchLoader, build_config, init_seed
    from recbole_fairrec_b200.atomic import AtomicDataset
    cfg = build_config("NFCF", "ml-100k", None, dict(
        data_path=os.path.join(os.path.dirname(__file__), "data"), threshold={"rating": 3.0}, sst_attr_list=["gender"],
        load_col={"inter": ["user_id", "item_id", "rating"], "user": ["user_id", "gender"]}, train_batch_size=512))
    init_seed(7)
    ds = AtomicDataset(cfg)
    train = ds.build()[0]
    loader = BatchLoader(cfg, ds, train, pairwise=False, pointwise_neg=True)
    used = set((train["user_id"] * ds.item_num + train["item_id"]).tolist())
    seen = 0
    for k, b in enumerate(loader):
        u, i, y, g = (b[c].numpy() for c in ("user_id", "item_id", "label", "gender"))
        h = len(u) // 2
        assert len(u) <= 512 and (y[:h] == 1).all() and (y[h:] == 0).all() and (u[:h] == u[h:]).all()
        assert all(int(a) * ds.item_num + int(c) in used for a, c in zip(u[:h], i[:h]))
        assert not any(int(a) * ds.item_num + int(c) in used for a, c in zip(u[h:], i[h:])) and (i[h:] >= 1).all()
        assert (g == ds.user_feat["gender"][u]).all()
        seen += h
        if k == 3:
            break
    assert seen == 4 * 256 and len(loader) == -(-len(train["user_id"]) // 256)


def test_ingestion_identical_to_the_live_reference_on_synthetic_files(tmp_path):
    """build container only (skipped where /root/reference is absent): the unmodified reference's create_dataset +
    data_preparation on freshly written synthetic atomic files (3 float attributes, 60k rows) vs atomic.py -- splits,
    user features and evaluation lists bit for bit (bench_ingest.py does the same at the ML-1M shape)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle", "ref_shim"))
    import shim
    if not shim.available():
        pytest.skip("reference tree not present")
    import bench_ingest as bi
    name = bi.write_files(str(tmp_path), "tiny-synth", n_users=801, n_items=501, n_inter=60_000, seed=11)
    _, ds, splits, lists = bi.ours(str(tmp_path), name)
    _, rds, rloaders = bi.reference(str(tmp_path), name)
    assert bi.compare(ds, splits, lists, rds, rloaders)


def test_eval_lists_with_duplicated_pairs_match_the_live_reference(tmp_path):
    """build container only: the `messy` files hold repeated (user, item) pairs, so an item can sit in a user's train
    rows AND among the positives of the evaluated split -- the reference then keeps it out of the history
    (general_dataloader.py:201-207); eval users, positives and histories of both phases vs the reference's loaders"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle", "ref_shim"))
    import shim
    if not shim.available():
        pytest.skip("reference tree not present")
    sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
    import bench_ingest as bi
    import make_test_data as mtd
    name = mtd.write_messy(str(tmp_path))
    old = dict(bi.CFG)
    try:
        bi.CFG.update(mtd.INGEST_BASE, **mtd.INGEST_CASES["defaults"])
        _, ds, splits, lists = bi.ours(str(tmp_path), name)
        _, rds, rloaders = bi.reference(str(tmp_path), name)
    finally:
        bi.CFG.clear()
        bi.CFG.update(old)
    n_dup = 0
    for ph, loader in (("valid", rloaders[1]), ("test", rloaders[2])):
        users, hist, pos = lists[ph]
        np.testing.assert_array_equal(users, np.asarray(loader.uid_list))
        for r, u in enumerate(users.tolist()):
            assert set(pos[r].tolist()) == set(loader.uid2positive_item[u].tolist())
            assert set(hist[r].tolist()) == set(loader.uid2history_item[u].tolist()), (ph, u)
        tr = set(zip(splits[0]["user_id"].tolist(), splits[0]["item_id"].tolist()))
        ev = splits[1] if ph == "valid" else splits[2]
        n_dup += len(tr & set(zip(ev["user_id"].tolist(), ev["item_id"].tolist())))
    assert n_dup > 0                      # the case is actually exercised


def _ingest_case_names():
    import sys
    sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
    import make_test_data as mtd
    return list(mtd.INGEST_CASES)


@pytest.mark.parametrize("case", _ingest_case_names())
def test_filtering_ordering_and_splitting_options_match_the_reference(case, tmp_path):
    """every `_data_filtering` branch (missing ids, rm_dup_inter, val_interval on float and token fields,
    filter_inter_by_user_or_item, k-core intervals), order RO | TO, split RS (grouped or not) | LS in its three modes:
    ids, id->token maps, the three splits (row order included) and the user features, bit for bit against the outputs of
    the unmodified reference on the same files (tests/golden/ingest_<case>.npz, oracle/gen_golden.py ingest)."""
    import make_test_data as mtd
    from recbole_fairrec_b200.atomic import AtomicDataset
    from recbole_fairrec_b200.quick_start import build_config, init_seed
    g = np.load(os.path.join(HERE, "golden", f"ingest_{case}.npz"))
    name = mtd.write_messy(str(tmp_path))
    cfg = build_config("FOCF", name, None, dict(mtd.INGEST_BASE, **mtd.INGEST_CASES[case], data_path=str(tmp_path),
                                                device="cpu"))
    init_seed(cfg["seed"])
    ds = AtomicDataset(cfg)
    assert (ds.user_num, ds.item_num) == (int(g["n_users"]), int(g["n_items"]))
    np.testing.assert_array_equal(np.array([str(t) for t in ds.field2id_token["user_id"]]), g["user_tokens"])
    np.testing.assert_array_equal(np.array([str(t) for t in ds.field2id_token["item_id"]]), g["item_tokens"])
    splits = ds.build()
    for k, part in zip(("train", "valid", "test"), splits):
        for col in ("user_id", "item_id", "rating", "timestamp", "label"):
            np.testing.assert_array_equal(part[col], g[f"{k}_{col}"], err_msg=f"{case} {k} {col}")
    for idf in (cfg["preload_weight"] or {}):              # pretrained-embedding matrices, row = remapped id
        np.testing.assert_array_equal(ds.get_preload_weight(idf), g["preload_" + idf], err_msg=idf)
    for col in ("gender", "age", "occupation"):
        want = g["user_" + col].copy()
        if want.dtype.kind == "f":
            # users without a feature row: the reference fills the column mean (dataset.py:571-572); under pandas 3 that
            # `fillna(inplace=True)` is a no-op and the fixture holds NaN there (SURVEY.md 8c caveat)
            want[np.isnan(want)] = np.float32(np.nanmean(want.astype(np.float64)))
        else:
            want[want < 0] = 0                               # same no-op for token columns: INT64_MIN instead of [PAD]
        np.testing.assert_array_equal(ds.user_feat[col][1:], want, err_msg=col)


def test_split_counts_follow_the_scalar_rule_for_all_small_groups():
    """dataset.py:1339-1360 restated scalar-wise (cnt = floor(ratio * tot), first part takes the rest, then every part
    whose exact share lies in (0, 1) steals one row from the first while the first has more than one) vs the vectorised
    calcu_split_counts, exhaustively for group sizes 0..400 and several ratio triples"""
    from recbole_fairrec_b200.atomic import calcu_split_counts

    def scalar(tot, ratios):
        r = [x / sum(ratios) for x in ratios]
        cnt = [int(x * tot) for x in r]
        cnt[0] = tot - sum(cnt[1:])
        for i in range(1, len(r)):
            if cnt[0] <= 1:
                break
            if 0 < r[-i] * tot < 1:
                cnt[-i] += 1
                cnt[0] -= 1
        return cnt

    tots = np.arange(0, 401)
    for ratios in ([8, 1, 1], [7, 2, 1], [0.6, 0.2, 0.2], [98, 1, 1], [1, 1, 1], [9, 1]):
        got = calcu_split_counts(tots, ratios)
        assert (got.sum(axis=1) == tots).all() and (got >= 0).all()
        for t in tots:
            assert got[t].tolist() == scalar(int(t), ratios), (t, ratios)


def test_interval_parsing_and_kcore_fixed_point():
    from recbole_fairrec_b200.atomic import data_filtering, parse_intervals, within_intervals
    from recbole_fairrec_b200.config import Config
    iv = parse_intervals("(0,1];[3,inf)")
    assert iv == [("(", 0.0, 1.0, "]"), ("[", 3.0, float("inf"), ")")]
    np.testing.assert_array_equal(within_intervals([0, 0.5, 1, 2, 3, 1e9], iv), [False, True, True, False, True, True])
    assert parse_intervals(None) is None
    rng = np.random.default_rng(0)
    for trial in range(5):
        n = 3000
        inter = {"user_id": np.array([f"u{k}" for k in rng.zipf(1.6, n) % 200], dtype=object),
                 "item_id": np.array([f"i{k}" for k in rng.zipf(1.4, n) % 150], dtype=object),
                 "rating": rng.integers(1, 6, n).astype(np.float64)}
        cfg = Config(user_inter_num_interval="[4,inf)", item_inter_num_interval="[3,200]", device="cpu")
        out, _, _ = data_filtering(cfg, inter, None, None, {"rating": "float"}, "user_id", "item_id")
        ku, ki = out["user_id"].tokens().astype(str), out["item_id"].tokens().astype(str)
        _, cu = np.unique(ku, return_counts=True)
        _, ci = np.unique(ki, return_counts=True)
        assert cu.min() >= 4 and ci.min() >= 3 and ci.max() <= 200      # a fixed point of both interval filters
        # and the kept rows are a subsequence of the input (order preserved)
        key_in = [a + "|" + b for a, b in zip(inter["user_id"], inter["item_id"])]
        key_out = [a + "|" + b for a, b in zip(ku, ki)]
        it = iter(key_in)
        assert all(k in it for k in key_out)
