"""GPU parity of NFCF (fr_nfcf_forward / fr_nfcf_backward through the C ABI) against the fixtures generated from the
unmodified reference (tests/golden/nfcf_train_*.npz) and against the numpy oracle."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import nfcf_oracle as no

pytestmark = pytest.mark.gpu
RTOL = 1e-5
HERE = os.path.dirname(__file__)
NFCF = sorted(glob.glob(os.path.join(HERE, "golden", "nfcf_train_*.npz")))


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def final_state_fp64(g, fair):
    """the fixture's schedule on the numpy oracle evaluated in float64 (module-level dtype switched for the call): the
    yardstick for how well conditioned the multi-step Adam result is"""
    from oracle import focf_oracle as fo
    L = int(g["n_layers"])
    old = (no.F32, fo.F32)
    no.F32 = fo.F32 = np.float64
    try:
        f64 = lambda a: np.asarray(a, np.float64)
        batches = [(g[f"uid{s}"], g[f"iid{s}"], g[f"label{s}"], g[f"sst{s}"]) for s in range(int(g["n_steps"]))]
        _, U, I, Ws, bs = no.train_steps(f64(g["U0"]), f64(g["I0"]), [f64(g[f"W{k}_0"]) for k in range(L - 1 + 1)],
                                         [f64(g[f"b{k}_0"]) for k in range(L - 1 + 1)], batches, fair,
                                         float(g["fair_weight"]), float(g["lr"]), float(g["wd"]), fair)
    finally:
        no.F32, fo.F32 = old
    return U, I, Ws, bs


def close_as_reference(mine, ref, truth64):
    """1e-5 relative; widened only where Adam's m/sqrt(v) normalisation puts the REFERENCE itself further than that from
    the float64 evaluation of the same schedule (then: as close to the float64 result as the reference, x3)"""
    tol = max(RTOL, 3.0 * rel_err(ref, truth64))
    return rel_err(mine, truth64) < tol


def make_model(g, fair, dropout=0.0):
    import recbole_fairrec_b200 as pkg
    from recbole_fairrec_b200.synth import SynthDataset
    L = int(g["n_layers"])
    d = g["U0"].shape[1]
    hidden = [g[f"W{k}_0"].shape[0] for k in range(L - 1)]
    cfg = pkg.Config(embedding_size=d, mlp_hidden_size=hidden, dropout=dropout, fair_weight=float(g["fair_weight"]),
                     load_pretrain_path=None, device=torch.device("cuda"))
    model = pkg.NFCF(cfg, SynthDataset(g["U0"].shape[0], g["I0"].shape[0], 5.0))
    with torch.no_grad():
        model.user_embedding.weight.copy_(torch.from_numpy(g["U0"]))
        model.item_embedding.weight.copy_(torch.from_numpy(g["I0"]))
        for k, lin in enumerate(model.mlp_layers.linears()):
            lin.weight.copy_(torch.from_numpy(g[f"W{k}_0"]))
            lin.bias.copy_(torch.from_numpy(g[f"b{k}_0"]))
    if fair:   # what reset_params leaves behind: regulariser on, user table frozen
        model.load_pretrain_path = "loaded"
        model.user_embedding.weight.requires_grad = False
    return model.cuda()


def make_inter(g, s):
    import recbole_fairrec_b200 as pkg
    return pkg.Interaction({"user_id": torch.from_numpy(g[f"uid{s}"]), "item_id": torch.from_numpy(g[f"iid{s}"]),
                            "label": torch.from_numpy(g[f"label{s}"]), "gender": torch.from_numpy(g[f"sst{s}"])})


@pytest.mark.parametrize("path", NFCF, ids=[os.path.basename(p)[11:-4] for p in NFCF])
def test_nfcf_matches_reference(path):
    g = np.load(path)
    fair, L = bool(g["fair"]), int(g["n_layers"])
    model = make_model(g, fair)
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=float(g["lr"]),
                           weight_decay=float(g["wd"]))          # trainer.py:139
    losses = []
    for s in range(int(g["n_steps"])):
        inter = make_inter(g, s)
        opt.zero_grad()
        loss = model.calculate_loss(inter)
        loss.backward()
        if s == 0:
            assert rel_err(model.predict(inter).cpu().numpy(), g["pred0"]) < RTOL
            assert rel_err(model.item_embedding.weight.grad.cpu().numpy(), g["dI0"]) < RTOL
            if "dU0" in g:
                assert rel_err(model.user_embedding.weight.grad.cpu().numpy(), g["dU0"]) < RTOL
            for k, lin in enumerate(model.mlp_layers.linears()):
                assert rel_err(lin.weight.grad.cpu().numpy(), g[f"dW{k}_0"]) < RTOL, k
                assert rel_err(lin.bias.grad.cpu().numpy(), g[f"db{k}_0"]) < RTOL, k
        opt.step()
        losses.append(loss.item())
    model.check_flags()
    np.testing.assert_allclose(losses, g["losses"], rtol=RTOL)
    U64, I64, W64, b64 = final_state_fp64(g, fair)
    assert close_as_reference(model.item_embedding.weight.detach().cpu().numpy(), g["I_final"], I64)
    assert close_as_reference(model.user_embedding.weight.detach().cpu().numpy(), g["U_final"], U64)
    for k, lin in enumerate(model.mlp_layers.linears()):
        assert close_as_reference(lin.weight.detach().cpu().numpy(), g[f"W{k}_final"], W64[k]), k
        assert close_as_reference(lin.bias.detach().cpu().numpy(), g[f"b{k}_final"], b64[k]), k


def test_nfcf_oracle_large_batch():
    """ML-1M-like widths (d=64, tower 128-128-64-1), 6000-row batch spanning several weight-gradient chunks"""
    rng = np.random.default_rng(3)
    nu, ni, d, B = 3000, 900, 64, 6000
    g = {"U0": (rng.standard_normal((nu, d)) * 0.3).astype(np.float32),
         "I0": (rng.standard_normal((ni, d)) * 0.3).astype(np.float32), "n_layers": 3, "fair_weight": 0.1}
    dims = [2 * d, 128, 64, 1]
    for k in range(3):
        g[f"W{k}_0"] = (rng.standard_normal((dims[k + 1], dims[k])) * (1.0 / np.sqrt(dims[k]))).astype(np.float32)
        g[f"b{k}_0"] = (rng.standard_normal(dims[k + 1]) * 0.1 + (0.3 if k == 2 else 0)).astype(np.float32)
    half = B // 2
    u = rng.integers(1, nu, half)
    uid = np.r_[u, u].astype(np.int64)
    iid = np.r_[rng.integers(1, ni // 4, half), rng.integers(1, ni, half)].astype(np.int64)
    label = np.r_[np.ones(half), np.zeros(half)].astype(np.float32)
    sst = rng.integers(1, 3, nu)[uid].astype(np.int64)
    g.update(uid0=uid, iid0=iid, label0=label, sst0=sst)
    model = make_model(g, True)
    loss = model.calculate_loss(make_inter(g, 0))
    loss.backward()
    Ws, bs = [g[f"W{k}_0"] for k in range(3)], [g[f"b{k}_0"] for k in range(3)]
    lo, p, dU, dI, dWs, dbs = no.loss_and_grads(g["U0"], g["I0"], Ws, bs, uid, iid, label, sst, True, 0.1)
    np.testing.assert_allclose(loss.item(), lo, rtol=RTOL)
    assert rel_err(model.item_embedding.weight.grad.cpu().numpy(), dI) < RTOL
    for k, lin in enumerate(model.mlp_layers.linears()):
        assert rel_err(lin.weight.grad.cpu().numpy(), dWs[k]) < RTOL, k
        assert rel_err(lin.bias.grad.cpu().numpy(), dbs[k]) < RTOL, k


def test_nfcf_dropout_trains_and_is_reproducible_per_step():
    """dropout > 0: the counter-based mask is the same in forward and backward of one step (finite-difference check of
    one weight) and differs between steps"""
    g = np.load(NFCF[0])
    model = make_model(g, False, dropout=0.3)
    model.train()
    inter = make_inter(g, 0)
    l1 = model.calculate_loss(inter).item()
    l2 = model.calculate_loss(inter).item()
    assert l1 != l2          # a new mask per call
    model.eval()
    e1, e2 = model.calculate_loss(inter).item(), model.calculate_loss(inter).item()
    assert e1 == e2          # no dropout in eval mode
