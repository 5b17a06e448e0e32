"""CPU restatement (numpy, float32) of the reference's FOCF training step.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg as the checker -- never by the product package (recbole-fairrec_b200/),
which must fail loudly when the CUDA library is missing.

Parity status: PINNED.  tests/test_oracle_golden.py checks every function here against
tests/golden/focf_train_*.npz, which oracle/gen_golden.py produced by running the
unmodified reference (recbole FOCF + torch.optim.Adam) in the build container, and against
KAT-1 of SURVEY.md section 4.

Each function cites the reference lines it restates (paths relative to /root/reference).
"""
import numpy as np

F32 = np.float32
OBJECTIVES = ("none", "value", "absolute", "under", "over", "nonparity")


def forward(U, I, uid, iid):
    """recbole/model/fair_recommender/focf.py:136-143 -- pred_b = sum_k U[u_b,k] * I[i_b,k]."""
    return (U[uid] * I[iid]).sum(axis=-1, dtype=F32).astype(F32)


def predict(U, I, uid, iid, max_rating):
    """focf.py:145-150 -- clamp(pred, 0, max_rating) / max_rating."""
    p = forward(U, I, uid, iid)
    return (np.clip(p, F32(0.0), F32(max_rating)) / F32(max_rating)).astype(F32)


def group_index(sst):
    """focf.py:77 -- torch.unique(sst, return_inverse=True): the group of a row is the RANK of its
    attribute value among the values present in the batch (a single-valued batch is all group 0)."""
    vals, inv = np.unique(np.asarray(sst), return_inverse=True)
    return vals, inv.astype(np.int64)


def item_group_stats(pred, iid, rating, sst):
    """focf.py:75-91 get_item_ratings -- per (unique item j, group g): Sp, St, n = count + 1e-5.
    Returns (P, T, n, item_inverse, group_inverse, J).  More than two attribute values make the
    reference's index_put_ into a [J,2] tensor fail; we raise the same way (IndexError)."""
    vals, ginv = group_index(sst)
    if len(vals) > 2:
        raise IndexError("index out of range: more than two sensitive-attribute values in the batch")
    items, jinv = np.unique(np.asarray(iid), return_inverse=True)
    J = len(items)
    Sp = np.zeros((J, 2), F32)
    St = np.zeros((J, 2), F32)
    n = np.zeros((J, 2), F32)
    np.add.at(Sp, (jinv, ginv), pred.astype(F32))      # index_put_(accumulate=True): sequential in b
    np.add.at(St, (jinv, ginv), np.asarray(rating, F32))
    np.add.at(n, (jinv, ginv), F32(1.0))
    n = (n + F32(1e-5)).astype(F32)
    return (Sp / n).astype(F32), (St / n).astype(F32), n, jinv, ginv, J


def _smooth_l1(x):
    """F.smooth_l1_loss(x, 0, beta=1, reduction='none') for x >= 0."""
    x = x.astype(F32)
    return np.where(x < F32(1.0), F32(0.5) * x * x, x - F32(0.5)).astype(F32)


def _dmat(objective, P, T):
    """focf.py:93-125 -- the per-(item, group) quantity D whose group gap is penalised."""
    if objective == "value":
        return (P - T).astype(F32)
    if objective == "absolute":
        return np.abs(P - T).astype(F32)
    if objective == "under":
        return np.where((T - P) > 0, T - P, F32(0.0)).astype(F32)
    if objective == "over":
        return np.where((P - T) > 0, P - T, F32(0.0)).astype(F32)
    raise ValueError(objective)


def fair_loss(objective, pred, iid, rating, sst):
    """focf.py:93-134.  Returns the scalar fairness loss (float32)."""
    objective = objective.strip().lower()
    if objective == "none":
        return F32(0.0)
    if objective == "nonparity":                      # focf.py:127-134
        vals = np.unique(np.asarray(sst))
        if len(vals) < 2:
            raise IndexError("index 1 is out of bounds: a single sensitive-attribute value in the batch")
        a1 = pred[np.asarray(sst) == vals[0]].mean(dtype=F32)
        a2 = pred[np.asarray(sst) == vals[1]].mean(dtype=F32)
        return _smooth_l1(np.abs(np.asarray(a1 - a2, F32)))[()]
    if objective not in OBJECTIVES:
        raise ValueError("you must set config['fair_objective'] be one of (none,"
                         "value,absolute,under,over,nonparity)")          # focf.py:67-68
    P, T, _, _, _, _ = item_group_stats(pred, iid, rating, sst)
    D = _dmat(objective, P, T)
    x = np.abs(D[:, 0] - D[:, 1]).astype(F32)
    return _smooth_l1(x).mean(dtype=F32)


def calculate_loss(U, I, uid, iid, rating, sst, objective="none", fair_weight=1.0):
    """focf.py:152-169 -- MSE(pred, rating) + fair_weight * fair."""
    pred = forward(U, I, uid, iid)
    r = np.asarray(rating, F32)
    mse = ((pred - r) * (pred - r)).mean(dtype=F32)
    fl = fair_loss(objective, pred, iid, r, sst)
    return F32(mse + F32(fair_weight) * fl)


def dloss_dpred(pred, iid, rating, sst, objective="none", fair_weight=1.0):
    """Autograd of focf.py:152-169 restated analytically:
    abs'(0)=0, `where` routes gradient only where its condition holds, smooth_l1' = x (x<1) else 1."""
    objective = objective.strip().lower()
    r = np.asarray(rating, F32)
    B = len(pred)
    coef = (F32(2.0) * (pred - r) / F32(B)).astype(F32)
    if objective == "none":
        return coef
    fw = F32(fair_weight)
    if objective == "nonparity":
        vals = np.unique(np.asarray(sst))
        m0, m1 = np.asarray(sst) == vals[0], np.asarray(sst) == vals[1]
        a1, a2 = pred[m0].mean(dtype=F32), pred[m1].mean(dtype=F32)
        z = F32(a1 - a2)
        hp = z if abs(z) < 1 else np.sign(z)
        coef = coef.copy()
        coef[m0] += fw * F32(hp) / F32(m0.sum())
        coef[m1] -= fw * F32(hp) / F32(m1.sum())
        return coef.astype(F32)
    P, T, n, jinv, ginv, J = item_group_stats(pred, iid, r, sst)
    D = _dmat(objective, P, T)
    z = (D[:, 0] - D[:, 1]).astype(F32)
    x = np.abs(z)
    hp = np.where(x < 1, x, F32(1.0)).astype(F32)
    q = (hp * np.sign(z) / F32(J)).astype(F32)             # dL_fair/dD[:,0] ; dD[:,1] gets -q
    if objective == "value":
        dDdP = np.ones_like(P)
    elif objective == "absolute":
        dDdP = np.sign(P - T).astype(F32)
    elif objective == "under":
        dDdP = np.where((T - P) > 0, F32(-1.0), F32(0.0)).astype(F32)
    else:
        dDdP = np.where((P - T) > 0, F32(1.0), F32(0.0)).astype(F32)
    dLdP = np.stack([q, -q], axis=1) * dDdP * fw           # [J,2]
    return (coef + (dLdP / n)[jinv, ginv]).astype(F32)


def grads(U, I, uid, iid, rating, sst, objective="none", fair_weight=1.0):
    """Dense embedding gradients as nn.Embedding (non-sparse) backward produces them:
    dU[u] += coef_b * I[i_b], dI[i] += coef_b * U[u_b], accumulated in batch order."""
    pred = forward(U, I, uid, iid)
    coef = dloss_dpred(pred, iid, rating, sst, objective, fair_weight)
    dU = np.zeros_like(U)
    dI = np.zeros_like(I)
    np.add.at(dU, uid, coef[:, None] * I[iid])
    np.add.at(dI, iid, coef[:, None] * U[uid])
    return pred, coef, dU, dI


def adam_step(p, g, m, v, step, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
    """torch.optim.Adam (_single_tensor_adam, amsgrad=False, L2 form) as built by
    recbole/trainer/trainer.py:139.  `step` is the 1-based step count AFTER increment.  In place."""
    g = (g + F32(weight_decay) * p).astype(F32) if weight_decay != 0 else g.astype(F32)
    m += F32(1 - beta1) * (g - m)                          # exp_avg.lerp_(grad, 1-beta1)
    v *= F32(beta2)
    v += F32(1 - beta2) * g * g                            # addcmul_(grad, grad, value=1-beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    step_size = lr / bc1
    denom = (np.sqrt(v) / F32(bc2 ** 0.5) + F32(eps)).astype(F32)
    p += F32(-step_size) * (m / denom)                     # addcdiv_(exp_avg, denom, value=-step_size)
    return p, m, v


def train_steps(U, I, batches, objective, fair_weight=1.0, lr=1e-3, weight_decay=0.0, beta1=0.9, beta2=0.999,
                eps=1e-8):
    """trainer.py:181-196 for a list of (uid, iid, rating, sst) batches: loss -> backward -> dense Adam
    over BOTH tables (every row moves every step because of L2 weight decay).  Returns the per-step
    losses and the final (U, I, mU, vU, mI, vI)."""
    U, I = U.copy(), I.copy()
    mU, vU, mI, vI = (np.zeros_like(U), np.zeros_like(U), np.zeros_like(I), np.zeros_like(I))
    losses = []
    for t, (uid, iid, rating, sst) in enumerate(batches, start=1):
        losses.append(calculate_loss(U, I, uid, iid, rating, sst, objective, fair_weight))
        _, _, dU, dI = grads(U, I, uid, iid, rating, sst, objective, fair_weight)
        adam_step(U, dU, mU, vU, t, lr, beta1, beta2, eps, weight_decay)
        adam_step(I, dI, mI, vI, t, lr, beta1, beta2, eps, weight_decay)
    return np.array(losses, F32), U, I, mU, vU, mI, vI


def train_steps_f64(U, I, batches, objective, **kw):
    """The same restatement evaluated in float64: the well-conditioned answer.  Tests use
    err(fp32 oracle, fp64 oracle) as the conditioning yardstick of a case -- Adam divides by sqrt(v), so an
    element whose gradient nearly cancels amplifies ordinary float32 rounding by orders of magnitude."""
    global F32
    keep = F32
    F32 = np.float64
    try:
        b64 = [(u, i, np.asarray(r, np.float64), s) for u, i, r, s in batches]
        return train_steps(np.asarray(U, np.float64), np.asarray(I, np.float64), b64, objective, **kw)
    finally:
        F32 = keep
