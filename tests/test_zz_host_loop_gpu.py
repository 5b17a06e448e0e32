"""GPU parity of FOCF.train_steps_host (include/fairrec_b200.h:fr_focf_train_steps_host -- the loop of trainer.py:181-196
over HOST batches run inside the library: per step one H2D copy, the fused step, one D2H copy of the loss) against
  (1) the golden fixtures of the unmodified reference (tests/golden/focf_train_*.npz: losses and final tables, 1e-5
      relative) and
  (2) the per-batch path FOCF.train_step on the same batches (bit-equal losses, tables and Adam moments).
The file sorts last on purpose: the entry point was written after the round's last GPU session, and a failure here must
not hide the rest of the suite from a `-x` run."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
RTOL = 1e-5
HERE = os.path.dirname(__file__)
TRAIN = sorted(glob.glob(os.path.join(HERE, "golden", "focf_train_*.npz")))


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def make_model(U0, I0, objective, fair_weight):
    import recbole_fairrec_b200 as pkg
    from recbole_fairrec_b200.synth import SynthDataset
    cfg = pkg.Config(embedding_size=U0.shape[1], fair_objective=objective, fair_weight=fair_weight,
                     device=torch.device("cuda"))
    model = pkg.FOCF(cfg, SynthDataset(U0.shape[0], I0.shape[0], 5.0))
    with torch.no_grad():
        model.user_embedding_layer.weight.copy_(torch.from_numpy(U0))
        model.item_embedding_layer.weight.copy_(torch.from_numpy(I0))
    return model.cuda()


def host_batches(g):
    import recbole_fairrec_b200 as pkg
    out = []
    for s in range(int(g["n_steps"])):
        iid = np.asarray(g[f"iid{s}"])
        heads = np.r_[True, iid[1:] != iid[:-1]]
        t = lambda a, dt: torch.from_numpy(np.asarray(a)).to(dt)
        out.append(pkg.pack_host_batch(t(g[f"uid{s}"], torch.int32), t(iid, torch.int32), t(g[f"rating{s}"], torch.float32),
                                       t(g[f"sst{s}"], torch.float32), fields=("user_id", "item_id", "rating", "gender"),
                                       items_contiguous=bool(len(np.unique(iid)) == heads.sum())))
    return out


@pytest.mark.parametrize("path", TRAIN, ids=[os.path.basename(p)[11:-4] for p in TRAIN])
def test_host_batch_loop_matches_the_reference_fixture_and_the_per_batch_path(path, monkeypatch):
    import recbole_fairrec_b200.focf as focf_mod
    monkeypatch.setattr(focf_mod, "_NO_FAST_HOST_STEP", True)          # per-batch reference = the generic host path
    g = np.load(path)
    obj, fw = str(g["objective"]), float(g["fair_weight"])
    batches = host_batches(g)
    loop = make_model(g["U0"], g["I0"], obj, fw)
    loop.init_adam(lr=float(g["lr"]), weight_decay=float(g["wd"]))
    losses = loop.train_steps_host(batches).clone()
    loop.check_flags()
    assert loop._adam["step"] == len(batches)
    np.testing.assert_allclose(losses.numpy(), g["losses"], rtol=RTOL)
    assert rel_err(loop.user_embedding_layer.weight.detach().cpu().numpy(), g["U_final"]) < RTOL
    assert rel_err(loop.item_embedding_layer.weight.detach().cpu().numpy(), g["I_final"]) < RTOL
    # the per-batch path on the same batches: same kernels, same arguments => the same bits
    step = make_model(g["U0"], g["I0"], obj, fw)
    step.init_adam(lr=float(g["lr"]), weight_decay=float(g["wd"]))
    ref = [float(step.train_step(b).item()) for b in batches]
    assert [float(x) for x in losses] == ref
    for a, b in ((loop.user_embedding_layer.weight, step.user_embedding_layer.weight),
                 (loop.item_embedding_layer.weight, step.item_embedding_layer.weight),
                 (loop._adam["mU"], step._adam["mU"]), (loop._adam["vI"], step._adam["vI"])):
        assert torch.equal(a.detach(), b.detach())


def test_host_batch_loop_on_ragged_batches_and_a_second_call(monkeypatch):
    """batch sizes that change from step to step (the staging buffer and the workspace grow), a second call continuing the
    optimizer step count, an empty list, and the library's refusal of a device buffer"""
    import recbole_fairrec_b200 as pkg
    import recbole_fairrec_b200.focf as focf_mod
    monkeypatch.setattr(focf_mod, "_NO_FAST_HOST_STEP", True)
    rng = np.random.default_rng(7)
    nu, ni, d = 300, 200, 32
    U0 = (rng.standard_normal((nu, d)) * 0.3).astype(np.float32)
    I0 = (rng.standard_normal((ni, d)) * 0.3).astype(np.float32)
    gender = rng.integers(1, 3, nu)

    def batch(n_items):
        uid, iid = [], []
        for it in rng.permutation(np.arange(1, ni))[:n_items]:
            c = int(rng.integers(1, 40))
            uid.append(rng.choice(np.arange(1, nu), c, replace=False))
            iid.append(np.full(c, it))
        uid, iid = np.concatenate(uid), np.concatenate(iid)
        t = torch.from_numpy
        return pkg.pack_host_batch(t(uid).to(torch.int32), t(iid).to(torch.int32),
                                   t(rng.integers(1, 6, len(uid)).astype(np.float32)), t(gender[uid].astype(np.float32)),
                                   fields=("user_id", "item_id", "rating", "gender"))

    batches = [batch(k) for k in (5, 60, 1, 150, 20, 20)]
    loop, step = make_model(U0, I0, "value", 1.0), make_model(U0, I0, "value", 1.0)
    for m in (loop, step):
        m.init_adam(lr=1e-3, weight_decay=1e-3)
    a = loop.train_steps_host(batches[:4]).clone()
    assert loop.train_steps_host([]).numel() == 0
    b = loop.train_steps_host(batches[4:]).clone()
    ref = [float(step.train_step(x).item()) for x in batches]
    assert [float(x) for x in torch.cat([a, b])] == ref and loop._adam["step"] == 6
    assert torch.equal(loop.user_embedding_layer.weight.detach(), step.user_embedding_layer.weight.detach())
    assert torch.equal(loop.item_embedding_layer.weight.detach(), step.item_embedding_layer.weight.detach())
    bad = batch(3)
    bad.packed_host = (bad.packed_host[0].cuda(), bad.packed_host[1])
    with pytest.raises(ValueError):
        loop.train_steps_host([bad])


@pytest.mark.parametrize("path", TRAIN[:3], ids=[os.path.basename(p)[11:-4] for p in TRAIN[:3]])
def test_packed_per_batch_path_equals_the_generic_one(path, monkeypatch):
    """FocfEngine.train_step_packed (persistent staging buffer + argument struct, also written after the last GPU session)
    against the generic per-batch path on the same host batches: the same bits"""
    import recbole_fairrec_b200.focf as focf_mod
    g = np.load(path)
    obj, fw = str(g["objective"]), float(g["fair_weight"])
    batches = host_batches(g)
    out = []
    for generic in (True, False):
        monkeypatch.setattr(focf_mod, "_NO_FAST_HOST_STEP", generic)
        m = make_model(g["U0"], g["I0"], obj, fw)
        m.init_adam(lr=float(g["lr"]), weight_decay=float(g["wd"]))
        losses = [float(m.train_step(b).item()) for b in batches]
        m.check_flags()
        out.append((losses, m.user_embedding_layer.weight.detach().clone(), m.item_embedding_layer.weight.detach().clone()))
    assert out[0][0] == out[1][0] and torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][2], out[1][2])
    np.testing.assert_allclose(out[1][0], g["losses"], rtol=RTOL)
