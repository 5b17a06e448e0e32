#!/usr/bin/env python
"""BASELINE.json configs[4] at FULL size on one B200: FOCF on 10M users x 1M items, d=128, 1e9 synthetic interactions.

    python bench_scaleout.py [--inter 1e9] [--users 1e7] [--items 1e6] [--batch 1048576] [--steps 10] [--eval-users N]

Everything is synthesised, split, sorted and indexed ON THE DEVICE: user activity and item
popularity log-normal (inverse-CDF sampling), ratings 1..5 with ML-1M's marginal, a binary gender ~ Bernoulli(0.28);
each interaction goes to train / valid / test with probability .8 / .1 / .1 (the reference splits 8:1:1 per user; the
per-interaction draw has the same shape).  (user, item) pairs are not de-duplicated (collision probability ~1e-4).
Prints ONE JSON line: FOCF train interactions/s (dense-exact Adam, batch rows as given; every step streams the 16.9 GB of
tables + moments, far beyond L2) with the k_apply HBM roofline, and full-sort fair-eval users/s over ALL users with a valid
interaction (tcgen05 3xTF32 scorer + mask + top-K + the 12 metrics) with the tensor roofline.
Not the driver's default bench (that is bench.py at the ML-1M shape): run explicitly; results are kept under profiles/.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    import torch

    import bench
    import recbole_fairrec_b200 as pkg
    from recbole_fairrec_b200 import _lib, synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--inter", type=float, default=1e9)
    ap.add_argument("--users", type=float, default=1e7)
    ap.add_argument("--items", type=float, default=1e6)
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--batch", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--eval-users", type=float, default=0, help="cap on evaluated users (0 = all)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    nu, ni, n_inter, d = int(args.users) + 1, int(args.items) + 1, int(args.inter), args.d
    g = torch.Generator(device=dev).manual_seed(2020)
    t_all = time.perf_counter()

    def cdf(n, sigma):
        w = torch.empty(n, device=dev).log_normal_(4.5, sigma, generator=g).double()
        return (torch.cumsum(w, 0) / w.sum()).float()

    ucdf, icdf = cdf(nu - 1, 1.0), cdf(ni - 1, 1.4)
    rcdf = torch.tensor(np.cumsum([.056, .107, .261, .349, .227]), dtype=torch.float32, device=dev)
    gender = (torch.rand(nu, device=dev, generator=g) < 0.28).float() + 1.0
    gender[0] = 0.0
    tr_u, tr_i, tr_r, va_u, va_i = [], [], [], [], []
    chunk = 100_000_000
    for lo in range(0, n_inter, chunk):
        m = min(chunk, n_inter - lo)
        u = (torch.searchsorted(ucdf, torch.rand(m, device=dev, generator=g)).clamp_(max=nu - 2) + 1).to(torch.int32)
        i = (torch.searchsorted(icdf, torch.rand(m, device=dev, generator=g)).clamp_(max=ni - 2) + 1).to(torch.int32)
        r = (torch.searchsorted(rcdf, torch.rand(m, device=dev, generator=g)).clamp_(max=4) + 1).to(torch.uint8)
        part = torch.rand(m, device=dev, generator=g)
        tr = part < 0.8
        va = (part >= 0.8) & (part < 0.9)
        tr_u.append(u[tr]); tr_i.append(i[tr]); tr_r.append(r[tr]); va_u.append(u[va]); va_i.append(i[va])
        del u, i, r, part, tr, va
    tr_u, tr_i, tr_r = torch.cat(tr_u), torch.cat(tr_i), torch.cat(tr_r)
    va_u, va_i = torch.cat(va_u), torch.cat(va_i)
    torch.cuda.synchronize()
    t_synth = time.perf_counter() - t_all

    t0 = time.perf_counter()
    tdata = pkg.TrainData.from_device(tr_u, tr_i.long(), tr_r, gender, nu, ni)
    torch.cuda.synchronize()
    t_csc = time.perf_counter() - t0
    cfg = pkg.Config(embedding_size=d, fair_objective="value", fair_weight=1.0, topk=[10], valid_metric="NDCG@10",
                     train_batch_size=args.batch, learning_rate=1e-3, weight_decay=1e-3, device=dev, seed=2020,
                     score_mode="tc", cuda_graph=False)
    loader = pkg.FOCFDataLoader(cfg, tdata, mode="fast", seed=2020)
    with torch.device(dev):
        model = pkg.FOCF(cfg, synth.SynthDataset(nu, ni, 5.0))
    model.init_adam(lr=1e-3, weight_decay=1e-3)
    n_steps = args.steps + args.warmup + 4
    items, offs, batches = loader.plan_epoch(n_steps)
    d_items, d_offs = torch.from_numpy(items).to(dev), torch.from_numpy(offs).to(dev)
    uf, itf, rf, sf = tdata.fields
    losses = torch.zeros(n_steps, device=dev)

    def step(k):
        uid, iid, rating, sst = loader.gather(d_items, d_offs, batches[k])
        inter = pkg.Interaction({uf: uid, itf: iid, rf: rating, sf: sst})
        inter.items_contiguous = True
        model.train_step(inter, loss_out=losses[k:k + 1])
        return batches[k][3]

    for k in range(args.warmup):
        step(k)
    torch.cuda.synchronize()
    model.check_flags()
    evs, rows = [], 0
    with bench.ClockSampler(0) as clocks:
        for k in range(args.warmup, args.warmup + args.steps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rows += step(k)
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        step_ms = [x.elapsed_time(y) for x, y in evs]
        _lib.profile_enable(True)
        rows_p = sum(step(args.warmup + args.steps + k) for k in range(4))
        prof = _lib.profile_report()
        _lib.profile_enable(False)
        # ------------------------------------------------------------ evaluation
        t0 = time.perf_counter()
        if args.eval_users:
            keep = va_u <= int(args.eval_users)
            va_u, va_i = va_u[keep], va_i[keep]
        edata = pkg.EvalData.from_device(tr_u, tr_i, va_u, va_i, {sf: gender}, nu, ni)
        torch.cuda.synchronize()
        t_csr = time.perf_counter() - t0
        del tr_u, tr_i, tr_r, va_u, va_i
        torch.cuda.empty_cache()
        counts = tdata.item_count_h
        evaluator = pkg.FullSortEvaluator(cfg, ni, {int(i): int(counts[i]) for i in tdata.item_uniques})
        Uw, Iw = model.user_embedding_layer.weight.data, model.item_embedding_layer.weight.data
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        _lib.profile_enable(True)
        a.record()
        res = evaluator.evaluate(Uw, Iw, edata, 5.0)
        b.record()
        torch.cuda.synchronize()
        prof_e = _lib.profile_report()
        _lib.profile_enable(False)
        eval_ms = a.elapsed_time(b)
    hbm, bf16, peak_src = bench.peaks()
    n_tab = nu + ni
    alg_apply = 24.0 * n_tab * d
    cnt, tot = prof["k_apply<fr::kAdamFused>"]
    ach = alg_apply / (tot / cnt * 1e-3) / 1e9
    B_avg = rows / args.steps
    step_bytes = (16.0 * d + 16.0) * B_avg + alg_apply
    mean_ms = sum(step_ms) / len(step_ms)
    ptot = sum(v[1] for v in prof.values()) or 1.0
    tc_cnt, tc_tot = prof_e.get("k_fullsort_tc", (1, float("nan")))
    tc_peak = bf16 / 2.0 / 3.0
    eval_flops = 2.0 * edata.n * ni * d
    etot = sum(v[1] for v in prof_e.values()) or 1.0
    out = {
        "metric": "FOCF train interactions/s", "value": rows / (sum(step_ms) / 1e3), "unit": "interactions/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (generated on the device)",
        "config": {"workload": "focf_scaleout_full (BASELINE.json configs[4])", "n_users": nu, "n_items": ni,
                   "n_inter": n_inter, "n_train": tdata.n_rows, "d": d, "train_batch_size": args.batch,
                   "avg_batch_rows": B_avg, "fair_objective": "value",
                   "optimizer": "adam(lr=1e-3, weight_decay=1e-3) dense-exact",
                   "l2": "no flush needed: every step streams 24*(Nu+Ni)*d = %.1f GB of tables + moments" % (alg_apply / 1e9)},
        "clocks": clocks.summary(),
        "roofline": {"bound": "hbm", "kernel": "k_apply<fr::kAdamFused>", "achieved": ach, "peak": hbm, "unit": "GB/s",
                     "frac": ach / hbm, "traffic": None, "algorithmic_bytes_per_launch": alg_apply,
                     "avg_launch_us": 1e3 * tot / cnt, "share_of_step": tot / ptot, "peak_source": peak_src},
        "step_roofline": {"algorithmic_bytes_per_step": step_bytes, "achieved": step_bytes / (mean_ms * 1e-3) / 1e9,
                          "unit": "GB/s", "frac": step_bytes / (mean_ms * 1e-3) / 1e9 / hbm},
        "kernel_shares": {k: round(v[1] / ptot, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]},
        "eval": {"metric": "full-sort fair-eval users/s", "value": edata.n / (eval_ms / 1e3), "unit": "users/s",
                 "n_users": edata.n, "n_pos": edata.n_pos, "history_entries": int(edata.hist_items.numel()),
                 "ms_per_pass": eval_ms, "score_mode": "tc_3xtf32",
                 "roofline": {"bound": "tensor", "kernel": "k_fullsort_tc", "achieved": eval_flops / (tc_tot / tc_cnt * 1e-3) / 1e12,
                              "peak": tc_peak, "unit": "TFLOP/s (fp32-equivalent; 3 TF32 MMAs per product)",
                              "frac": eval_flops / (tc_tot / tc_cnt * 1e-3) / 1e12 / tc_peak,
                              "avg_launch_us": 1e3 * tc_tot / tc_cnt,
                              "peak_source": f"{peak_src}: bf16 {bf16} / 2 (tf32) / 3 (3xTF32)"},
                 "kernel_shares": {k: round(v[1] / etot, 4) for k, v in sorted(prof_e.items(), key=lambda kv: -kv[1][1])[:6]},
                 "metrics": {k: float(v) for k, v in res.items()}},
        "setup_s": {"synthesis": t_synth, "train_csc_sort": t_csc, "eval_csr_build": t_csr},
        "hbm_allocated_gb": torch.cuda.max_memory_allocated() / 1e9,
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
