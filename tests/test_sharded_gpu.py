"""Row-sharded FOCF step (fr_focf_shard_step_run, csrc/focf_shard.cu) and lazy-exact Adam (fr_adam_mode), through the C ABI.

The single-GPU test box runs P emulated ranks on one device (ShardedGroupEmu: same kernels, same exchange layout, the host
sequences the phases where a multi-process run has cross-GPU barriers) and compares with the single-GPU fused step of
FOCF.train_step on the SAME batches: losses and updated tables within 1e-5 (the only difference is the order in which the
partial item x group sums and the item gradients of the ranks are added).  lazy_exact must equal dense_exact BIT FOR BIT.
The real multi-process path (CUDA IPC + barriers) is covered by tests/test_dp_gpu.py (needs >= 2 GPUs) and bench.py's dp_check.
Reference semantics: recbole/trainer/trainer.py:181-196, recbole/model/fair_recommender/focf.py:75-169."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def synth_case(seed, n_users, n_items, n_inter, d):
    from recbole_fairrec_b200 import synth
    uid, iid, rating, gender = synth.interactions(n_users, n_items, n_inter, seed, item_sigma=1.0)
    rng = np.random.default_rng(seed)
    U0 = (rng.standard_normal((n_users, d)) * 0.2).astype(np.float32)
    I0 = (rng.standard_normal((n_items, d)) * 0.2).astype(np.float32)
    return uid, iid, rating, gender, U0, I0


def single_gpu_reference(uid, iid, rating, gender, U0, I0, n_users, n_items, d, batches, objective, fair_weight, mode="dense_exact",
                         flush_every=None, no_fused=False):
    """FOCF.train_step on the batches `batches` (lists of drawn item ids) -> losses, U, I, model"""
    import recbole_fairrec_b200 as pkg
    from recbole_fairrec_b200.synth import SynthDataset
    dev = torch.device("cuda")
    counts = np.bincount(iid, minlength=n_items)
    cfg = pkg.Config(embedding_size=d, fair_objective=objective, fair_weight=fair_weight, device=dev,
                     train_batch_size=int(max(counts[np.asarray(b)].sum() for b in batches)))   # sizes the loader's columns
    train = pkg.TrainData(uid, iid, rating, gender, n_users, n_items, dev)
    draws = np.concatenate([np.r_[b, -1] for b in batches])
    loader = pkg.FOCFDataLoader(cfg, train, draws=draws)
    model = pkg.FOCF(cfg, SynthDataset(n_users, n_items, 5.0))
    with torch.no_grad():
        model.user_embedding_layer.weight.copy_(torch.from_numpy(U0))
        model.item_embedding_layer.weight.copy_(torch.from_numpy(I0))
    model = model.to(dev)
    model.init_adam(lr=1e-3, weight_decay=1e-3, mode=mode, max_steps=4096)
    model._adam["no_fused"] = no_fused
    losses = torch.zeros(len(batches), device=dev)
    items, offs, planned = loader.plan_epoch(len(batches))
    d_items, d_offs = torch.from_numpy(items).to(dev), torch.from_numpy(offs).to(dev)
    uf, itf, rf, sf = train.fields
    for k, b in enumerate(planned):
        u, i, r, s = loader.gather(d_items, d_offs, b)
        inter = pkg.Interaction({uf: u, itf: i, rf: r, sf: s})
        inter.items_contiguous = True
        model.train_step(inter, loss_out=losses[k:k + 1])
        if flush_every and (k + 1) % flush_every == 0:
            model.flush_adam()
    model.flush_adam()
    model.check_flags()
    return (losses.cpu().numpy(), model.user_embedding_layer.weight.detach().cpu().numpy(),
            model.item_embedding_layer.weight.detach().cpu().numpy(), model)


@pytest.mark.parametrize("objective", ["value", "absolute", "under", "over", "none"])
def test_lazy_exact_adam_is_bit_identical_to_dense_exact(objective):
    """a row's update while no batch touches it is a recurrence in (p, m, v, t) alone: replaying it when the row is next
    touched (or at a flush) must leave EXACTLY the tables, moments and losses of the reference's dense Adam"""
    n_users, n_items, d = 1500, 300, 64
    uid, iid, rating, gender, U0, I0 = synth_case(11, n_users, n_items, 40000, d)
    rng = np.random.default_rng(3)
    present = np.unique(iid)
    batches = [rng.choice(present, int(rng.integers(2, 9)), replace=False) for _ in range(24)]   # few items: most rows idle
    out = {}
    for mode, fe in (("dense_exact", None), ("lazy_exact", None), ("lazy_exact", 5)):
        # no_fused: the small-batch cooperative kernel of dense_exact adds the batch statistics in another order
        out[(mode, fe)] = single_gpu_reference(uid, iid, rating, gender, U0, I0, n_users, n_items, d, batches, objective, 0.7,
                                               mode, fe, no_fused=True)
    base = out[("dense_exact", None)]
    for key in (("lazy_exact", None), ("lazy_exact", 5)):
        got = out[key]
        np.testing.assert_array_equal(got[0], base[0])
        np.testing.assert_array_equal(got[1], base[1])
        np.testing.assert_array_equal(got[2], base[2])
        for m in ("mU", "vU", "mI", "vI"):
            assert torch.equal(got[3]._adam[m], base[3]._adam[m]), m
    assert not np.array_equal(base[1], U0)


def test_lazy_exact_200_random_steps_bit_identical():
    """a 200-step random run, larger batches (general multi-kernel preparation path), d = 128"""
    n_users, n_items, d = 4000, 500, 128
    uid, iid, rating, gender, U0, I0 = synth_case(5, n_users, n_items, 120000, d)
    rng = np.random.default_rng(9)
    present = np.unique(iid)
    batches = [rng.choice(present, int(rng.integers(1, 60)), replace=False) for _ in range(200)]
    a = single_gpu_reference(uid, iid, rating, gender, U0, I0, n_users, n_items, d, batches, "value", 1.0, "dense_exact",
                             no_fused=True)
    b = single_gpu_reference(uid, iid, rating, gender, U0, I0, n_users, n_items, d, batches, "value", 1.0, "lazy_exact", 37)
    for x, y in zip(a[:3], b[:3]):
        np.testing.assert_array_equal(x, y)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("adam_mode", ["dense_exact", "lazy_exact"])
def test_row_sharded_step_equals_single_gpu_step(world, adam_mode):
    import recbole_fairrec_b200 as pkg
    n_users, n_items, d, B = 1200, 260, 64, 3000
    uid, iid, rating, gender, U0, I0 = synth_case(world + 20, n_users, n_items, 30000, d)
    dev = torch.device("cuda")
    emu = pkg.ShardedGroupEmu(uid, iid, rating, gender, n_users, n_items, world, dev, d, B, seed=4, objective="value",
                              fair_weight=0.8, adam_mode=adam_mode, max_steps=256)
    try:
        emu.set_tables(U0, I0)
        n_steps = 6
        plans = emu.plan(n_steps)
        losses = [torch.zeros(n_steps, device=dev) for _ in range(world)]
        emu.train(plans, losses)
        for r in emu.ranks:
            r.check_flags()
        U, I = emu.full_tables()
        torch.cuda.synchronize()
        items = plans[0]["items"].cpu().numpy()
        batches = [items[b["items_pos"]:b["items_pos"] + b["J"]] for b in plans[0]["desc"]]
        # every rank drew the same items and the local rows add up to the global batch
        for p in plans[1:]:
            assert np.array_equal(p["items"].cpu().numpy(), items)
        for k in range(n_steps):
            assert sum(p["desc"][k]["B_loc"] for p in plans) == plans[0]["desc"][k]["B_glob"]
        lo, Ur, Ir, _ = single_gpu_reference(uid, iid, rating, gender, U0, I0, n_users, n_items, d, batches, "value", 0.8)
        for r in range(world):                       # identical loss on every rank
            np.testing.assert_array_equal(losses[r].cpu().numpy(), losses[0].cpu().numpy())
        np.testing.assert_allclose(losses[0].cpu().numpy(), lo, rtol=RTOL)
        assert rel_err(U.cpu().numpy(), Ur) < RTOL and rel_err(I.cpu().numpy(), Ir) < RTOL
    finally:
        emu.close()


def test_row_sharded_step_with_a_rank_that_has_no_rows_and_single_gender_ranks():
    """ragged case: 8 ranks, tiny batches -- some ranks hold no row of a batch, some only one attribute value; objectives
    other than value; run-to-run bit stability"""
    import recbole_fairrec_b200 as pkg
    n_users, n_items, d, B = 90, 40, 32, 12
    uid, iid, rating, gender, U0, I0 = synth_case(77, n_users, n_items, 400, d)
    dev = torch.device("cuda")
    for objective in ("absolute", "over"):
        runs = []
        for rep in range(2):
            emu = pkg.ShardedGroupEmu(uid, iid, rating, gender, n_users, n_items, 8, dev, d, B, seed=1, objective=objective,
                                      fair_weight=1.0, max_steps=64)
            try:
                emu.set_tables(U0, I0)
                plans = emu.plan(5)
                losses = [torch.zeros(5, device=dev) for _ in range(8)]
                emu.train(plans, losses)
                for r in emu.ranks:
                    r.check_flags()
                U, I = emu.full_tables()
                runs.append((losses[0].cpu().numpy(), U.cpu().numpy(), I.cpu().numpy(), plans))
            finally:
                emu.close()
        np.testing.assert_array_equal(runs[0][1], runs[1][1])
        np.testing.assert_array_equal(runs[0][2], runs[1][2])
        plans = runs[0][3]
        assert any(p["desc"][k]["B_loc"] == 0 for p in plans for k in range(5))
        items = plans[0]["items"].cpu().numpy()
        batches = [items[b["items_pos"]:b["items_pos"] + b["J"]] for b in plans[0]["desc"]]
        lo, Ur, Ir, _ = single_gpu_reference(uid, iid, rating, gender, U0, I0, n_users, n_items, d, batches, objective, 1.0)
        np.testing.assert_allclose(runs[0][0], lo, rtol=RTOL)
        assert rel_err(runs[0][1], Ur) < RTOL and rel_err(runs[0][2], Ir) < RTOL
