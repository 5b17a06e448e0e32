"""The persistent epoch kernel (fr_focf_epoch_run, csrc/focf_epoch.cu) against the stepwise path: the same planned batches
through ONE cooperative launch (producer CTAs build the batches, compute CTAs keep tables and Adam moments in shared memory)
must leave bit-identical losses, tables and moments (trainer.py:181-196 over focf_dataloader.py:37-50)."""
import numpy as np
import pytest
import torch

from test_focf_train_gpu import make_model

pytestmark = pytest.mark.gpu


def _setup(seed, n_users, n_items, n_inter, batch, d, objective):
    import recbole_fairrec_b200 as pkg
    from recbole_fairrec_b200 import synth
    dev = torch.device("cuda")
    uid, iid, rating, gender = synth.interactions(n_users, n_items, n_inter, seed, item_sigma=1.0)
    cfg = pkg.Config(embedding_size=d, fair_objective=objective, fair_weight=0.7, train_batch_size=batch, device=dev,
                     learning_rate=1e-3, weight_decay=1e-3, seed=seed, epochs=2)
    train = pkg.TrainData(uid, iid, rating, gender, n_users, n_items, dev)
    rng = np.random.default_rng(seed)
    U0 = (rng.standard_normal((n_users, d)) * 0.2).astype(np.float32)
    I0 = (rng.standard_normal((n_items, d)) * 0.2).astype(np.float32)
    return cfg, train, U0, I0


def _state(model, losses):
    a = model._adam
    return [losses.cpu().numpy().copy()] + [t.detach().cpu().numpy().copy() for t in (
        model.user_embedding_layer.weight, model.item_embedding_layer.weight, a["mU"], a["vU"], a["mI"], a["vI"])]


def _run(cfg, train, U0, I0, objective, persistent, splits, loader_seed=11):
    """two epochs; `splits`: how the steps of an epoch are cut into runner.run(k) calls (None: one call per epoch)"""
    import recbole_fairrec_b200 as pkg
    loader = pkg.FOCFDataLoader(cfg, train, mode="fast", seed=loader_seed)
    model = make_model(U0, I0, objective, 0.7)
    model.init_adam(lr=1e-3, weight_decay=1e-3)
    n = len(loader)
    losses = torch.zeros(2 * n, device="cuda")
    kinds = set()
    for ep in range(2):
        runner = model.planned_runner(loader, losses[ep * n:(ep + 1) * n], graph_steps=4, persistent=persistent)
        kinds.add(type(runner).__name__)
        left = n - runner.cursor
        for k in (splits or [left]):
            k = min(k, left)
            runner.run(k)
            left -= k
        runner.run(left)
    model.check_flags()
    torch.cuda.synchronize()
    assert model._adam["step"] == 2 * n
    return _state(model, losses), kinds, n


@pytest.mark.parametrize("objective,d,shape", [
    ("value", 32, (900, 400, 60000, 2048)),
    ("absolute", 64, (1200, 300, 50000, 1024)),
    ("under", 128, (700, 250, 30000, 2048)),
    ("over", 64, (2500, 1500, 90000, 4096)),
    ("nonparity", 32, (900, 400, 40000, 2048)),
    ("none", 64, (600, 200, 20000, 512)),
])
def test_epoch_kernel_is_bit_identical_to_the_stepwise_path(objective, d, shape):
    cfg, train, U0, I0 = _setup(7, *shape, d, objective)
    ref, kinds_ref, n = _run(cfg, train, U0, I0, objective, False, None)
    out, kinds, _ = _run(cfg, train, U0, I0, objective, True, None)
    assert kinds_ref == {"_PlannedRunner"} and kinds == {"_EpochRunner"}
    assert n >= 5
    for a, b, name in zip(ref, out, ("losses", "U", "I", "mU", "vU", "mI", "vI")):
        np.testing.assert_array_equal(a, b, err_msg=name)


def test_epoch_kernel_in_pieces_and_with_few_slots():
    """runs of 1, 2, 3, ... steps per launch (each launch re-loads and writes back the resident state; slots are re-armed),
    also with the minimum of two slots"""
    import recbole_fairrec_b200 as pkg
    cfg, train, U0, I0 = _setup(3, 900, 400, 60000, 2048, 32, "value")
    ref, _, n = _run(cfg, train, U0, I0, "value", False, None)
    out, kinds, _ = _run(cfg, train, U0, I0, "value", True, [1, 2, 3, 1, 5])
    assert kinds == {"_EpochRunner"}
    for a, b, name in zip(ref, out, ("losses", "U", "I", "mU", "vU", "mI", "vI")):
        np.testing.assert_array_equal(a, b, err_msg=name)
    # two slots
    loader = pkg.FOCFDataLoader(cfg, train, mode="fast", seed=11)
    model = make_model(U0, I0, "value", 0.7)
    model.init_adam(lr=1e-3, weight_decay=1e-3)
    losses = torch.zeros(2 * n, device="cuda")
    for ep in range(2):
        runner = model.epoch_runner(loader, losses[ep * n:(ep + 1) * n], n_slots=2)
        assert runner is not None
        runner.run(n)
    out2 = _state(model, losses)
    for a, b, name in zip(ref, out2, ("losses", "U", "I", "mU", "vU", "mI", "vI")):
        np.testing.assert_array_equal(a, b, err_msg=name + " (2 slots)")


def test_epoch_kernel_then_stepwise_steps_on_the_same_model():
    """an epoch through the persistent kernel, then eager fr_focf_train_step calls: the workspaces are left consistent"""
    import recbole_fairrec_b200 as pkg
    cfg, train, U0, I0 = _setup(5, 900, 400, 60000, 2048, 32, "value")
    res = []
    for persistent in (False, True):
        loader = pkg.FOCFDataLoader(cfg, train, mode="fast", seed=4)
        model = make_model(U0, I0, "value", 0.7)
        model.init_adam(lr=1e-3, weight_decay=1e-3)
        n = len(loader)
        losses = torch.zeros(n + 4, device="cuda")
        model.train_epoch_planned(loader, losses[:n], graph_steps=4, persistent=persistent)
        for k, inter in enumerate(loader):
            if k == 4:
                break
            model.train_step(inter, loss_out=losses[n + k:n + k + 1])
        model.check_flags()
        res.append(_state(model, losses))
    for a, b, name in zip(res[0], res[1], ("losses", "U", "I", "mU", "vU", "mI", "vI")):
        np.testing.assert_array_equal(a, b, err_msg=name)


def test_tables_too_large_for_residency_fall_back_to_the_stepwise_runner():
    import recbole_fairrec_b200 as pkg
    cfg, train, U0, I0 = _setup(9, 60000, 3000, 50000, 2048, 128, "value")
    loader = pkg.FOCFDataLoader(cfg, train, mode="fast", seed=1)
    model = make_model(U0, I0, "value", 0.7)
    model.init_adam(lr=1e-3, weight_decay=1e-3)
    losses = torch.zeros(len(loader), device="cuda")
    assert model.epoch_runner(loader, losses) is None
    runner = model.planned_runner(loader, losses, graph_steps=4)
    assert type(runner).__name__ == "_PlannedRunner"
    runner.run(3)
    model.check_flags()


def test_pipelined_epochs_equal_epoch_by_epoch_training():
    """train_epochs_planned (host draws epoch e + 1 while epoch e runs; two plan / loss buffer sets) == one
    train_epoch_planned call per epoch on a loader with the same seed"""
    import recbole_fairrec_b200 as pkg
    cfg, train, U0, I0 = _setup(13, 900, 400, 60000, 2048, 32, "value")
    res = []
    for pipelined in (False, True):
        loader = pkg.FOCFDataLoader(cfg, train, mode="fast", seed=5)
        model = make_model(U0, I0, "value", 0.7)
        model.init_adam(lr=1e-3, weight_decay=1e-3)
        n = len(loader)
        if pipelined:
            sums, steps, rows = model.train_epochs_planned(loader, 3)
            assert steps == 3 * n and len(sums) == 3
        else:
            losses = torch.zeros(n, device="cuda")
            sums = []
            for ep in range(3):
                model.train_epoch_planned(loader, losses)
                sums.append(float(losses[:n].double().sum()))
        model.check_flags()
        torch.cuda.synchronize()
        res.append((sums, _state(model, torch.zeros(1, device="cuda"))[1:]))
    assert res[0][0] == res[1][0]
    for a, b, name in zip(res[0][1], res[1][1], ("U", "I", "mU", "vU", "mI", "vI")):
        np.testing.assert_array_equal(a, b, err_msg=name)
