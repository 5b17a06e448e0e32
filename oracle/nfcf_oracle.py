"""CPU restatement (numpy, float32) of the reference's NFCF training step: NCF tower + BCE + differential-fairness
regulariser.  TEST INFRASTRUCTURE ONLY (tests/, smoke, bench cpu legs) -- never imported by the product package.

Parity status: PINNED against tests/golden/nfcf_train_*.npz (generated from the unmodified reference by
oracle/gen_golden.py) in tests/test_oracle_golden.py.

Restates (paths relative to /root/reference):
  recbole/model/fair_recommender/nfcf.py:69-74   forward: emb || emb -> MLPLayers -> sigmoid
  recbole/model/layers.py:58-70                  MLPLayers: Dropout -> Linear -> activation after EVERY layer (ReLU)
  nfcf.py:76-97                                  get_differential_fairness over the batch's positives
  nfcf.py:99-110                                 calculate_loss = BCE + fair_weight * DF (when a pretrain was loaded)
"""
import numpy as np

F32 = np.float32


def tower_forward(x, Ws, bs):
    """returns (output after the last ReLU, list of post-activation layer outputs)"""
    acts, h = [], x.astype(F32)
    for W, b in zip(Ws, bs):
        h = np.maximum(h @ W.T + b, F32(0)).astype(F32)      # Linear -> ReLU, also after the last layer
        acts.append(h)
    return h, acts


def forward(U, I, Ws, bs, uid, iid):
    """nfcf.py:69-74"""
    x = np.concatenate([U[uid], I[iid]], axis=1).astype(F32)
    z, acts = tower_forward(x, Ws, bs)
    p = (F32(1) / (F32(1) + np.exp(-z[:, 0]))).astype(F32)
    return p, x, acts


def bce(p, y):
    """nn.BCELoss: mean of -(y log p + (1-y) log(1-p)), logs clamped at -100"""
    lp = np.maximum(np.log(p), F32(-100))
    l1 = np.maximum(np.log(F32(1) - p), F32(-100))
    return F32(-(y * lp + (F32(1) - y) * l1).mean(dtype=F32))


def differential_fairness(p, iid, label, sst):
    """nfcf.py:76-97.  Returns (value, d value / d p) -- the gradient restates autograd: torch.where routes the
    gradient to the branch taken, abs'(0) = 0."""
    pos = label == 1
    idx = np.flatnonzero(pos)
    sv, g = np.unique(sst[pos], return_inverse=True)
    iv, j = np.unique(iid[pos], return_inverse=True)
    J, G = len(iv), len(sv)
    S = np.zeros((J, G), F32)
    N = np.zeros((J, G), F32)
    np.add.at(S, (j, g), p[pos])
    np.add.at(N, (j, g), F32(1))
    alpha = F32(1.0 / J)
    M = ((S + alpha) / (N + F32(1))).astype(F32)
    eps = np.zeros(J, F32)
    dM = np.zeros((J, G), F32)                       # d eps_j / d M[j, :] of the pair that currently holds the max
    for a in range(G):
        for b in range(a + 1, G):
            diff = (np.log(M[:, a]) - np.log(M[:, b])).astype(F32)
            e = np.abs(diff)
            take = e > eps
            eps = np.where(take, e, eps)
            sgn = np.sign(diff)
            dM[take] = 0
            dM[take, a] = (sgn / M[:, a])[take]
            dM[take, b] = (-sgn / M[:, b])[take]
    val = F32(eps.mean(dtype=F32))
    dp = np.zeros_like(p)
    dp[idx] = (dM / (N + F32(1)))[j, g] / F32(J)
    return val, dp


def loss_and_grads(U, I, Ws, bs, uid, iid, label, sst, fair, fair_weight):
    """nfcf.py:99-110 + autograd.  Returns (loss, p, dU, dI, dWs, dbs)."""
    p, x, acts = forward(U, I, Ws, bs, uid, iid)
    y = label.astype(F32)
    B = len(p)
    loss = bce(p, y)
    dp = (-(y / p) + (F32(1) - y) / (F32(1) - p)).astype(F32) / F32(B)
    if fair:
        val, ddf = differential_fairness(p, iid, label, sst)
        loss = F32(loss + F32(fair_weight) * val)
        dp = (dp + F32(fair_weight) * ddf).astype(F32)
    dz = (dp * p * (F32(1) - p))[:, None].astype(F32)                # through the sigmoid
    dWs, dbs = [None] * len(Ws), [None] * len(Ws)
    d = dz
    for k in range(len(Ws) - 1, -1, -1):
        d = (d * (acts[k] > 0)).astype(F32)                            # through the ReLU of layer k
        inp = acts[k - 1] if k > 0 else x
        dWs[k] = (d.T @ inp).astype(F32)
        dbs[k] = d.sum(axis=0, dtype=F32)
        d = (d @ Ws[k]).astype(F32)
    dU, dI = np.zeros_like(U), np.zeros_like(I)
    E = U.shape[1]
    np.add.at(dU, uid, d[:, :E])
    np.add.at(dI, iid, d[:, E:])
    return loss, p, dU, dI, dWs, dbs


def train_steps(U, I, Ws, bs, batches, fair, fair_weight, lr, wd, user_frozen):
    """trainer.py:181-196 with torch.optim.Adam over the parameters that require grad."""
    from .focf_oracle import adam_step
    U, I = U.copy(), I.copy()
    Ws, bs = [w.copy() for w in Ws], [b.copy() for b in bs]
    params = ([] if user_frozen else [U]) + [I] + [t for pair in zip(Ws, bs) for t in pair]
    m = [np.zeros_like(t) for t in params]
    v = [np.zeros_like(t) for t in params]
    losses = []
    for t, (uid, iid, label, sst) in enumerate(batches, start=1):
        loss, _, dU, dI, dWs, dbs = loss_and_grads(U, I, Ws, bs, uid, iid, label, sst, fair, fair_weight)
        grads = ([] if user_frozen else [dU]) + [dI] + [g for pair in zip(dWs, dbs) for g in pair]
        for prm, g, mm, vv in zip(params, grads, m, v):
            adam_step(prm, g, mm, vv, t, lr, 0.9, 0.999, 1e-8, wd)
        losses.append(loss)
    return np.array(losses, F32), U, I, Ws, bs
