"""GPU parity of the FOCF training step (C ABI -> sm_100a kernels) against
  (1) the golden fixtures produced by the unmodified reference (tests/golden/focf_train_*.npz), and
  (2) the numpy oracle on larger seeded inputs.
Tolerance: 1e-5 relative (north star: losses and updated embeddings within 1e-5 relative)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import focf_oracle as fo

pytestmark = pytest.mark.gpu
RTOL = 1e-5
HERE = os.path.dirname(__file__)
TRAIN = sorted(glob.glob(os.path.join(HERE, "golden", "focf_train_*.npz")))


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def make_model(U0, I0, objective, fair_weight, max_rating=5.0):
    import recbole_fairrec_b200 as pkg
    from recbole_fairrec_b200.synth import SynthDataset
    cfg = pkg.Config(embedding_size=U0.shape[1], fair_objective=objective, fair_weight=fair_weight,
                     device=torch.device("cuda"))
    model = pkg.FOCF(cfg, SynthDataset(U0.shape[0], I0.shape[0], max_rating))
    with torch.no_grad():
        model.user_embedding_layer.weight.copy_(torch.from_numpy(U0))
        model.item_embedding_layer.weight.copy_(torch.from_numpy(I0))
    return model.cuda()


def make_inter(uid, iid, rating, sst, contiguous):
    import recbole_fairrec_b200 as pkg
    inter = pkg.Interaction({"user_id": torch.from_numpy(np.asarray(uid)), "item_id": torch.from_numpy(np.asarray(iid)),
                             "rating": torch.from_numpy(np.asarray(rating, np.float32)),
                             "gender": torch.from_numpy(np.asarray(sst))})
    inter.items_contiguous = contiguous
    return inter


def is_contiguous(iid):
    iid = np.asarray(iid)
    heads = np.r_[True, iid[1:] != iid[:-1]]
    return len(np.unique(iid)) == heads.sum()


@pytest.mark.parametrize("path", TRAIN, ids=[os.path.basename(p)[11:-4] for p in TRAIN])
@pytest.mark.parametrize("promise_contiguous", [False, True])
def test_golden_compat_path(path, promise_contiguous):
    """calculate_loss().backward() + torch.optim.Adam, exactly the reference Trainer's sequence."""
    g = np.load(path)
    obj, fw = str(g["objective"]), float(g["fair_weight"])
    contiguous = promise_contiguous and is_contiguous(g["iid0"])
    if promise_contiguous and not contiguous:
        pytest.skip("batch is not item-contiguous")
    model = make_model(g["U0"], g["I0"], obj, fw)
    opt = torch.optim.Adam(model.parameters(), lr=float(g["lr"]), weight_decay=float(g["wd"]))
    losses = []
    for s in range(int(g["n_steps"])):
        inter = make_inter(g[f"uid{s}"], g[f"iid{s}"], g[f"rating{s}"], g[f"sst{s}"],
                           promise_contiguous and is_contiguous(g[f"iid{s}"]))
        opt.zero_grad()
        loss = model.calculate_loss(inter)
        loss.backward()
        if s == 0:
            pred = model._engine().pred_buf[:len(g["uid0"])].cpu().numpy()
            assert rel_err(pred, g["pred0"]) < RTOL
            assert rel_err(model.user_embedding_layer.weight.grad.cpu().numpy(), g["dU0"]) < RTOL
            assert rel_err(model.item_embedding_layer.weight.grad.cpu().numpy(), g["dI0"]) < RTOL
            assert rel_err(model.predict(inter).cpu().numpy(), g["predict0"]) < RTOL
        opt.step()
        losses.append(loss.item())
    model.check_flags()
    np.testing.assert_allclose(losses, g["losses"], rtol=RTOL)
    assert rel_err(model.user_embedding_layer.weight.detach().cpu().numpy(), g["U_final"]) < RTOL
    assert rel_err(model.item_embedding_layer.weight.detach().cpu().numpy(), g["I_final"]) < RTOL


@pytest.mark.parametrize("path", TRAIN, ids=[os.path.basename(p)[11:-4] for p in TRAIN])
def test_golden_fused_path(path):
    """model.train_step(): forward + backward + dense Adam fused, no dense gradient."""
    g = np.load(path)
    obj, fw = str(g["objective"]), float(g["fair_weight"])
    model = make_model(g["U0"], g["I0"], obj, fw)
    adam = model.init_adam(lr=float(g["lr"]), weight_decay=float(g["wd"]))
    losses = torch.zeros(int(g["n_steps"]), device="cuda")
    for s in range(int(g["n_steps"])):
        inter = make_inter(g[f"uid{s}"], g[f"iid{s}"], g[f"rating{s}"], g[f"sst{s}"], is_contiguous(g[f"iid{s}"]))
        model.train_step(inter, loss_out=losses[s:s + 1])
    model.check_flags()
    np.testing.assert_allclose(losses.cpu().numpy(), g["losses"], rtol=RTOL)
    for mine, ref in ((model.user_embedding_layer.weight, "U_final"), (model.item_embedding_layer.weight, "I_final"),
                      (adam["mU"], "mU_final"), (adam["mI"], "mI_final"), (adam["vU"], "vU_final"),
                      (adam["vI"], "vI_final")):
        assert rel_err(mine.detach().cpu().numpy(), g[ref]) < RTOL, ref


def random_case(seed, n_users, n_items, d, B, shuffle=False):
    rng = np.random.default_rng(seed)
    U = (rng.standard_normal((n_users, d)) * 0.3).astype(np.float32)
    I = (rng.standard_normal((n_items, d)) * 0.3).astype(np.float32)
    gender = rng.integers(1, 3, n_users)
    uid, iid = [], []
    for it in rng.permutation(np.arange(1, n_items)):
        # a few very popular items so that segments span many 32-entry chunks
        cnt = int(min(n_users - 1, rng.integers(1, 40) if rng.random() > 0.05 else rng.integers(200, 1500)))
        uid.append(rng.choice(np.arange(1, n_users), cnt, replace=False))
        iid.append(np.full(cnt, it))
        if sum(map(len, uid)) >= B:
            break
    uid, iid = np.concatenate(uid), np.concatenate(iid)
    r = rng.integers(1, 6, len(uid)).astype(np.float32)
    if shuffle:
        p = rng.permutation(len(uid))
        uid, iid, r = uid[p], iid[p], r[p]
    return U, I, uid.astype(np.int64), iid.astype(np.int64), r, gender[uid].astype(np.int64)


@pytest.mark.parametrize("objective", ["none", "value", "absolute", "under", "over", "nonparity"])
@pytest.mark.parametrize("d,B,shuffle", [(64, 5000, False), (128, 3000, True), (16, 700, False), (256, 2100, False)])
def test_oracle_two_fused_steps(objective, d, B, shuffle):
    U0, I0, uid, iid, r, sst = random_case(sum(map(ord, objective)) + d, 3000, 800, d, B, shuffle)
    batches = [(uid, iid, r, sst), (uid[::-1].copy(), iid[::-1].copy(), r[::-1].copy(), sst[::-1].copy())]
    losses_o, U_o, I_o, mU, vU, mI, vI = fo.train_steps(U0, I0, batches, objective, 0.7, 1e-3, 1e-3)
    model = make_model(U0, I0, objective, 0.7)
    adam = model.init_adam(lr=1e-3, weight_decay=1e-3)
    losses = torch.zeros(2, device="cuda")
    for s, (a, b, c, e) in enumerate(batches):
        model.train_step(make_inter(a, b, c, e, is_contiguous(b)), loss_out=losses[s:s + 1])
    model.check_flags()
    np.testing.assert_allclose(losses.cpu().numpy(), losses_o, rtol=RTOL)
    # Adam divides by sqrt(v): an element whose two gradients nearly cancel amplifies float32 rounding (on this very
    # case stock torch CPU ops differ from the numpy oracle by 8e-6).  The tolerance is 1e-5 unless the case itself is
    # worse conditioned than that, measured as the float32 oracle's own distance from the float64 oracle.
    _, U64, I64, mU64, vU64, mI64, vI64 = fo.train_steps_f64(U0, I0, batches, objective, fair_weight=0.7, lr=1e-3,
                                                             weight_decay=1e-3)

    def close(mine, o32, o64, what):
        tol = max(RTOL, 8 * rel_err(o32, o64))
        err = rel_err(mine.detach().cpu().numpy(), o64)
        assert err < tol, (what, err, tol)

    close(model.user_embedding_layer.weight, U_o, U64, "U")
    close(model.item_embedding_layer.weight, I_o, I64, "I")
    close(adam["mU"], mU, mU64, "mU")
    close(adam["mI"], mI, mI64, "mI")
    close(adam["vU"], vU, vU64, "vU")
    close(adam["vI"], vI, vI64, "vI")


def test_adam_kernel_alone():
    """fr_focf_adam from a given dense gradient == torch.optim.Adam's update rule (oracle adam_step)"""
    rng = np.random.default_rng(1)
    U0 = rng.standard_normal((500, 64)).astype(np.float32)
    I0 = rng.standard_normal((300, 64)).astype(np.float32)
    model = make_model(U0, I0, "none", 1.0)
    adam = model.init_adam(lr=1e-3, weight_decay=1e-3)
    Uo, Io = U0.copy(), I0.copy()
    mU, vU, mI, vI = np.zeros_like(U0), np.zeros_like(U0), np.zeros_like(I0), np.zeros_like(I0)
    for step in range(1, 4):
        gU = (rng.standard_normal(U0.shape) * 10.0 ** rng.integers(-6, 1, U0.shape)).astype(np.float32)
        gI = (rng.standard_normal(I0.shape) * 10.0 ** rng.integers(-6, 1, I0.shape)).astype(np.float32)
        adam["step"] = step
        model._engine().adam_dense(model.user_embedding_layer.weight.data, model.item_embedding_layer.weight.data,
                                   torch.from_numpy(gU).cuda(), torch.from_numpy(gI).cuda(), adam)
        fo.adam_step(Uo, gU, mU, vU, step, 1e-3, 0.9, 0.999, 1e-8, 1e-3)
        fo.adam_step(Io, gI, mI, vI, step, 1e-3, 0.9, 0.999, 1e-8, 1e-3)
    assert rel_err(model.user_embedding_layer.weight.detach().cpu().numpy(), Uo) < 1e-6
    assert rel_err(model.item_embedding_layer.weight.detach().cpu().numpy(), Io) < 1e-6
    assert rel_err(adam["vU"].cpu().numpy(), vU) < 1e-6 and rel_err(adam["mI"].cpu().numpy(), mI) < 1e-6


def test_gradients_match_oracle_large_batch():
    """multi-tile sort path (B > 2048) + dense-gradient read-out"""
    U0, I0, uid, iid, r, sst = random_case(11, 6041, 3707, 64, 9000)
    model = make_model(U0, I0, "value", 1.0)
    loss = model.calculate_loss(make_inter(uid, iid, r, sst, True))
    loss.backward()
    pred, coef, dU, dI = fo.grads(U0, I0, uid, iid, r, sst, "value", 1.0)
    assert rel_err(model.user_embedding_layer.weight.grad.cpu().numpy(), dU) < RTOL
    assert rel_err(model.item_embedding_layer.weight.grad.cpu().numpy(), dI) < RTOL
    np.testing.assert_allclose(loss.item(), fo.calculate_loss(U0, I0, uid, iid, r, sst, "value", 1.0), rtol=RTOL)


def test_run_to_run_bit_stability():
    U0, I0, uid, iid, r, sst = random_case(5, 3000, 800, 64, 6000)
    outs = []
    for _ in range(2):
        model = make_model(U0, I0, "value", 1.0)
        model.init_adam(lr=1e-3, weight_decay=1e-3)
        for _s in range(3):
            model.train_step(make_inter(uid, iid, r, sst, True))
        outs.append((model.user_embedding_layer.weight.detach().cpu().numpy().copy(),
                     model.item_embedding_layer.weight.detach().cpu().numpy().copy()))
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    np.testing.assert_array_equal(outs[0][1], outs[1][1])


def test_fault_flags():
    """>2 attribute values -> IndexError like focf.py:86; single value + nonparity -> IndexError (focf.py:130)"""
    U0, I0, uid, iid, r, sst = random_case(2, 300, 100, 16, 400)
    sst3 = sst.copy()
    sst3[::3] = 3
    model = make_model(U0, I0, "value", 1.0)
    model.calculate_loss(make_inter(uid, iid, r, sst3, True))
    with pytest.raises(IndexError):
        model.check_flags()
    model = make_model(U0, I0, "nonparity", 1.0)
    model.calculate_loss(make_inter(uid, iid, r, np.ones_like(sst), True))
    with pytest.raises(IndexError):
        model.check_flags()
    # single group with an item objective is legal (everything lands in column 0)
    model = make_model(U0, I0, "value", 1.0)
    loss = model.calculate_loss(make_inter(uid, iid, r, np.ones_like(sst), True))
    model.check_flags()
    np.testing.assert_allclose(loss.item(), fo.calculate_loss(U0, I0, uid, iid, r, np.ones_like(sst), "value", 1.0),
                               rtol=RTOL)


def _ml_like_loader(seed, n_users=900, n_items=400, n_inter=60000, batch=2048, d=32):
    import recbole_fairrec_b200 as pkg
    from recbole_fairrec_b200 import synth
    dev = torch.device("cuda")
    uid, iid, rating, gender = synth.interactions(n_users, n_items, n_inter, seed, item_sigma=1.0)
    cfg = pkg.Config(embedding_size=d, fair_objective="value", fair_weight=1.0, train_batch_size=batch, device=dev,
                     learning_rate=1e-3, weight_decay=1e-3, seed=seed, epochs=2)
    train = pkg.TrainData(uid, iid, rating, gender, n_users, n_items, dev)
    rng = np.random.default_rng(seed)
    U0 = (rng.standard_normal((n_users, d)) * 0.2).astype(np.float32)
    I0 = (rng.standard_normal((n_items, d)) * 0.2).astype(np.float32)
    return cfg, train, U0, I0


def test_planned_graph_epoch_equals_stepwise_epoch():
    """The CUDA-graph replay over a planned epoch (device-resident batch size / cursor / Adam step, fused single-CTA
    preparation) is bit-identical to the same batches fed one fr_focf_train_step at a time."""
    import recbole_fairrec_b200 as pkg
    cfg, train, U0, I0 = _ml_like_loader(3)
    results = []
    for planned in (False, True):
        loader = pkg.FOCFDataLoader(cfg, train, mode="fast", seed=11)
        model = make_model(U0, I0, "value", 1.0)
        model.init_adam(lr=1e-3, weight_decay=1e-3)
        n = len(loader)
        losses = torch.zeros(2 * n, device="cuda")
        for ep in range(2):
            if planned:
                k, rows = model.train_epoch_planned(loader, losses[ep * n:(ep + 1) * n], graph_steps=4)
                assert k == n
            else:
                for k, inter in enumerate(loader):
                    model.train_step(inter, loss_out=losses[ep * n + k:ep * n + k + 1])
        model.check_flags()
        results.append((losses.cpu().numpy().copy(), model.user_embedding_layer.weight.detach().cpu().numpy().copy(),
                        model.item_embedding_layer.weight.detach().cpu().numpy().copy(), model._adam["step"]))
    assert results[0][3] == results[1][3] == 2 * n
    np.testing.assert_array_equal(results[0][0], results[1][0])
    np.testing.assert_array_equal(results[0][1], results[1][1])
    np.testing.assert_array_equal(results[0][2], results[1][2])


def test_planned_graph_epoch_with_large_batches_equals_stepwise_epoch():
    """Batches beyond the fused single-CTA preparation (> 8192 rows: multi-kernel sort / segments, multi-launch compute):
    the pipelined graphs of the planned runner -- preparation of batch t + 1 on a second stream under the compute of batch
    t, the configs[4] regime of bench.py's `pipelined` block -- are bit-identical to one fr_focf_train_step per batch."""
    import recbole_fairrec_b200 as pkg
    cfg, train, U0, I0 = _ml_like_loader(9, n_users=4000, n_items=700, n_inter=260000, batch=30000, d=64)
    results = []
    for planned in (False, True):
        loader = pkg.FOCFDataLoader(cfg, train, mode="fast", seed=5)
        assert loader.max_batch > 8192
        model = make_model(U0, I0, "value", 1.0)
        model.init_adam(lr=1e-3, weight_decay=1e-3)
        n = len(loader)
        assert n >= 5
        losses = torch.zeros(2 * n, device="cuda")
        for ep in range(2):
            if planned:
                k, rows = model.train_epoch_planned(loader, losses[ep * n:(ep + 1) * n], graph_steps=4, persistent=False)
                assert k == n
            else:
                for k, inter in enumerate(loader):
                    model.train_step(inter, loss_out=losses[ep * n + k:ep * n + k + 1])
        model.check_flags()
        results.append((losses.cpu().numpy().copy(), model.user_embedding_layer.weight.detach().cpu().numpy().copy(),
                        model.item_embedding_layer.weight.detach().cpu().numpy().copy(), model._adam["step"]))
    assert results[0][3] == results[1][3] == 2 * n
    np.testing.assert_array_equal(results[0][0], results[1][0])
    np.testing.assert_array_equal(results[0][1], results[1][1])
    np.testing.assert_array_equal(results[0][2], results[1][2])


def test_planned_epoch_of_a_single_batch_equals_train_step():
    """train rows <= train_batch_size: the epoch plan holds ONE batch; the planned runner's warm-up must not run it twice
    and the host's Adam step count must stay equal to the device's over several epochs"""
    import recbole_fairrec_b200 as pkg
    cfg, train, U0, I0 = _ml_like_loader(5, n_users=300, n_items=60, n_inter=1500, batch=4096)
    results = []
    for planned in (False, True):
        loader = pkg.FOCFDataLoader(cfg, train, mode="fast", seed=3)
        assert len(loader) == 1
        model = make_model(U0, I0, "value", 1.0)
        model.init_adam(lr=1e-3, weight_decay=1e-3)
        losses = torch.zeros(3, device="cuda")
        for ep in range(3):
            if planned:
                k, _ = model.train_epoch_planned(loader, losses[ep:ep + 1], graph_steps=4)
                assert k == 1
            else:
                for inter in loader:
                    model.train_step(inter, loss_out=losses[ep:ep + 1])
        model.check_flags()
        results.append((losses.cpu().numpy().copy(), model.user_embedding_layer.weight.detach().cpu().numpy().copy(),
                        model.item_embedding_layer.weight.detach().cpu().numpy().copy(), model._adam["step"]))
    assert results[0][3] == results[1][3] == 3
    np.testing.assert_array_equal(results[0][0], results[1][0])
    np.testing.assert_array_equal(results[0][1], results[1][1])
    np.testing.assert_array_equal(results[0][2], results[1][2])


def test_small_and_general_preparation_agree():
    """B <= 8192 takes the fused single-CTA preparation, larger batches the multi-kernel path: same results"""
    from recbole_fairrec_b200 import kernels
    U0, I0, uid, iid, r, sst = random_case(21, 3000, 800, 64, 7000, shuffle=True)
    outs = []
    for pad in (0, 3000):   # padding rows from extra items push the batch over the 8192 limit
        if pad:
            rng = np.random.default_rng(0)
            eu = rng.integers(1, 3000, pad)
            ei = rng.integers(1, 800, pad)
            u2, i2 = np.r_[uid, eu], np.r_[iid, ei]
            r2, s2 = np.r_[r, rng.integers(1, 6, pad).astype(np.float32)], np.r_[sst, rng.integers(1, 3, pad)]
        else:
            u2, i2, r2, s2 = uid, iid, r, sst
        model = make_model(U0, I0, "value", 1.0)
        loss = model.calculate_loss(make_inter(u2, i2, r2, s2, False))
        loss.backward()
        o = fo.grads(U0, I0, u2, i2, r2, s2, "value", 1.0)
        assert rel_err(model.user_embedding_layer.weight.grad.cpu().numpy(), o[2]) < RTOL
        assert rel_err(model.item_embedding_layer.weight.grad.cpu().numpy(), o[3]) < RTOL
        np.testing.assert_allclose(loss.item(), fo.calculate_loss(U0, I0, u2, i2, r2, s2, "value", 1.0), rtol=RTOL)
