#!/usr/bin/env python
"""Differential fuzz of the PFCN / FairGo / NFCF / sampled-evaluation oracles against the LIVE reference (build
container only): the generators of oracle/gen_golden.py (which drive the unmodified reference models, trainers' schedules,
`torch.optim.Adam`, `_neg_sample_batch_eval`, `Collector`, `Evaluator`) are run with random shapes / seeds / modes into a
temp dir, and every fixture is put through the SAME checks that tests/test_oracle_golden.py applies to the committed ones.
    python oracle/fuzz_families.py [seed] [trials per family]
TEST INFRASTRUCTURE ONLY."""
import contextlib
import io
import os
import sys
import tempfile
import traceback

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
_args = sys.argv[1:]
sys.argv = sys.argv[:1]
import gen_golden as gg  # noqa: E402
import test_oracle_golden as T  # noqa: E402


def check_nfcf(path):
    """tests/test_oracle_golden.test_nfcf_oracle_matches_reference with the conditioning yardstick of the GPU tests: Adam
    divides by sqrt(v), so on a random case a dead ReLU unit can put the float32 reference itself 1e-4 away from the
    float64 evaluation of the same schedule; the oracle has to be as close to float64 as the reference is (x3), or 1e-5"""
    from oracle import nfcf_oracle as no
    g = np.load(path)
    rel = lambda a, b: float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() /
                             max(np.abs(np.asarray(b, np.float64)).max(), 1e-30))
    f64 = lambda a: np.asarray(a, np.float64)
    L = int(g["n_layers"])
    Ws, bs = [g[f"W{k}_0"] for k in range(L)], [g[f"b{k}_0"] for k in range(L)]
    batches = [(g[f"uid{s}"], g[f"iid{s}"], g[f"label{s}"], g[f"sst{s}"]) for s in range(int(g["n_steps"]))]
    fair, fw, lr, wd = bool(g["fair"]), float(g["fair_weight"]), float(g["lr"]), float(g["wd"])
    losses, U, I, Wf, bf = no.train_steps(g["U0"], g["I0"], Ws, bs, batches, fair, fw, lr, wd, fair)
    _, _, I64, W64, b64 = no.train_steps(f64(g["U0"]), f64(g["I0"]), [f64(w) for w in Ws], [f64(b) for b in bs], batches,
                                         fair, fw, lr, wd, fair)
    assert rel(losses, g["losses"]) < 1e-5
    pairs = [(I, g["I_final"], I64)] + [(Wf[k], g[f"W{k}_final"], W64[k]) for k in range(L)] + \
            [(bf[k], g[f"b{k}_final"], b64[k]) for k in range(L)]
    for mine, ref, hi in pairs:
        assert rel(mine, ref) <= max(1e-5, 3 * rel(ref, hi)), (rel(mine, ref), rel(ref, hi))


def main():
    seed = int(_args[0]) if _args else 0
    n = int(_args[1]) if len(_args) > 1 else 4
    rng = np.random.default_rng(seed)
    tmp = tempfile.mkdtemp()
    keep, gg.OUT = gg.OUT, tmp
    cases = []
    for t in range(n):
        s = int(rng.integers(0, 1 << 30))
        cases.append(("nfcf", lambda t=t, s=s: gg.run_nfcf(f"z{t}", fair=bool(t % 2), n_users=int(rng.integers(30, 150)),
                                                         n_items=int(rng.integers(20, 90)), d=int(rng.choice([8, 16, 64])),
                                                         hidden=tuple(int(x) for x in rng.choice([16, 32, 64], 2)),
                                                         n_steps=int(rng.integers(1, 4)), B=int(rng.integers(40, 300)), seed=s),
                      f"nfcf_train_z{t}.npz", check_nfcf))
        model = ["PFCN_MLP", "PFCN_PMF", "PFCN_BiasedMF", "PFCN_DMF"][t % 4]
        mode = ["sm", "cm"][int(rng.integers(0, 2))]
        cases.append((f"pfcn {model} {mode}", lambda t=t, s=s, model=model, mode=mode: gg.run_pfcn(
            model, f"z{t}", filter_mode=mode, n_users=int(rng.integers(40, 120)), n_items=int(rng.integers(30, 80)),
            d=int(rng.choice([8, 16])), B=int(rng.integers(48, 200)), seed=s, n_rounds=int(rng.integers(1, 3))),
            f"pfcn_z{t}.npz", T.test_pfcn_oracle_matches_reference))
        aggr = ["LBA", "WAP", "LVA"][t % 3]
        layers = int(rng.integers(1, 4)) if aggr != "LVA" else 2
        cases.append((f"fairgo {aggr} {layers}", lambda t=t, s=s, aggr=aggr, layers=layers: gg.run_fairgo(
            f"z{t}", aggr, n_layers=layers, n_users=int(rng.integers(40, 100)), n_items=int(rng.integers(30, 70)),
            d=int(rng.choice([8, 16])), B=int(rng.integers(48, 160)), n_inter=int(rng.integers(300, 900)), seed=s),
            f"fairgo_z{t}.npz", T.test_fairgo_oracle_matches_reference))
        cases.append(("uni eval", lambda t=t, s=s: gg.run_uni_eval(
            f"z{t}", n_users=int(rng.integers(20, 80)), n_items=int(rng.integers(60, 500)), d=int(rng.choice([8, 16])),
            seed=s, neg_num=int(rng.choice([10, 50, 100])), users_per_batch=int(rng.integers(1, 5))),
            f"uni_eval_z{t}.npz", T.test_sampled_eval_oracle_matches_reference))
    bad = 0
    try:
        for name, gen, fname, check in cases:
            try:
                with contextlib.redirect_stdout(io.StringIO()):
                    gen()
            except Exception as e:
                print(name, "generator raised (reference-side):", type(e).__name__, str(e)[:100])
                continue
            try:
                check(os.path.join(tmp, fname))
                print(name, "ok")
            except Exception:
                bad += 1
                print(name, "BAD")
                traceback.print_exc(limit=3)
    finally:
        gg.OUT = keep
    print("bad:", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
