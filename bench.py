#!/usr/bin/env python
"""bench.py -- FOCF train interactions/s and full-sort fair-eval users/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scale F]

Headline workload = BASELINE.json configs[4], the configuration the metric is quoted on "at 1/2/4/8 B200" and the only one
BASELINE.json shards: FOCF on 10M users x 1M items, d = 128, 1e9 synthetic interactions (generated on the device), value
objective, Adam(lr 1e-3, weight_decay 1e-3), train_batch_size 2^20 rows PER GPU (weak scaling: global batch = N x 2^20, one
optimizer step per global batch like the reference's loop).  It fits one GPU (74 GB), so N = 1 runs it whole; N > 1 runs the
row-sharded step (recbole-fairrec_b200/sharded.py, csrc/focf_shard.cu: rank r owns rows r::N of both tables and moments,
touched item rows / partial item x group sums / partial item gradients travel as stores into NVLink peer memory, no
collective inside a step).  `--scale F` shrinks users / items / interactions / batch by F (smoke runs).

ONE JSON line on stdout (rank 0):
  value        dense_exact device-resident steps: exactly K steps between two CUDA events on the launching stream, barrier +
               synchronize on both sides, max over ranks.  Every step streams >= 24 * rows * d bytes of tables + moments per
               GPU (33.8 GB at N = 1), far beyond the 126 MB L2: no flush needed.
  lazy_exact   the same K steps with adam_mode lazy_exact (bit-identical tables, untouched rows replayed on demand), the
               closing flush INSIDE the timed region (it carries the deferred work of the K steps)
  e2e          the same metric through the public API with HOST batches: pinned host batch -> H2D -> step -> loss D2H every
               step (one step in flight), wall clock, max over ranks
  roofline     dominant kernel of the timed steps: algorithmic bytes per launch / its mean CUDA-event duration (library
               profiler, a second pass over K more batches in the same regime) vs MEASURED_PEAKS.json
  eval         full-sort fair evaluation (tcgen05 3xTF32 scorer + history mask + top-K + the 12 metrics) of the first
               --eval-users users with a valid interaction against all 1M items; item table sharded over the ranks for N > 1
  ml1m         BASELINE.json configs[1] (6,040 x 3,706, d = 64, batch 2048) on one GPU: bench_ml1m.py (N = 1 only)
  families     BASELINE.json configs[2] / [3] and the NFCF / sampled-evaluation legs: bench_families.py (N = 1 only)
  dp_check     N > 1: the row-sharded step over CUDA-IPC peer memory vs the single-GPU step on the same batches, and the
               item-sharded evaluation vs the unsharded one (ids bit-equal, metrics equal), on a reduced shape
  cpu_baseline N = 1: the reference arm's figure (below) measured in the same run
`--impl reference`: the UNMODIFIED reference (oracle/_ref = copy of its `recbole` package, imported through
oracle/ref_shim) on the host cores: recbole.model.fair_recommender.focf.FOCF.calculate_loss + backward +
torch.optim.Adam(weight_decay) + loss.item() on pre-built host Interactions -- the bracket of `e2e` -- on a 1/S-scale
replica of the workload (users, items and batch divided by S, same d: the reference's step cost is linear in both, so
interactions/s carries over), kind "reference".
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCALEOUT = dict(n_users=10_000_001, n_items=1_000_001, n_inter=1_000_000_000, d=128, batch=1 << 20, K=10)


def scaled(scale):
    w = dict(SCALEOUT)
    if scale != 1.0:
        w["n_users"] = max(int((w["n_users"] - 1) * scale), 1000) + 1
        w["n_items"] = max(int((w["n_items"] - 1) * scale), 200) + 1
        w["n_inter"] = max(int(w["n_inter"] * scale), 100_000)
        w["batch"] = max(int(w["batch"] * scale), 4096)
    return w


def peaks():
    """(HBM GB/s, dense bf16 TFLOP/s, source): MEASURED_PEAKS.json (driver-written) or the recipe's fallback"""
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), float(p["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 7700.0 * 0.85, 2250.0 * 0.75, "fallback (B200_PROFILING.md: 85 % of 7.7 TB/s, 75 % of 2.25 PFLOP/s)"


class ClockSampler:
    """nvidia-smi SM clock + throttle reasons sampled in a thread DURING the timed region"""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self._stop = index, [], threading.Event()

    def __enter__(self):
        def loop():
            while not self._stop.is_set():
                r = self._read()
                if r:
                    self.samples.append(r)
                self._stop.wait(0.05)
        self._t = threading.Thread(target=loop, daemon=True)
        self._t.start()
        return self

    def _read(self):
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=2).stdout.strip().split(",")
            return [x.strip() for x in out] if len(out) == 6 else None
        except Exception:
            return None

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=3)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ reference arm
def reference_arm(args, w, scale_div=8, quiet=False):
    """The unmodified reference's FOCF step on the host cores (see the module docstring)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shim"))
    import shim
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kind = "reference"
    try:
        shim.install()
        from recbole.data.interaction import Interaction
        from recbole.model.fair_recommender.focf import FOCF as RefFOCF
    except Exception as e:               # oracle/_ref missing (it is made by __graft_entry__.build()): the restated port
        kind = "port"
        why = str(e)[:120]
    nu = (w["n_users"] - 1) // scale_div + 1
    ni = (w["n_items"] - 1) // scale_div + 1
    B = max(w["batch"] // scale_div, 256)
    d = w["d"]
    rng = np.random.default_rng(2020)
    # batches shaped like FOCFDataLoader's: whole items (item-contiguous), log-normal popularity, ~B rows each
    n_steps = args.steps + args.warmup
    pop = rng.lognormal(4.5, 1.4, ni - 1)
    mean_rows = w["n_inter"] * 0.8 / (w["n_items"] - 1)
    cnt_of_item = np.maximum((pop / pop.mean() * mean_rows / 1.0).astype(np.int64), 1)
    cnt_of_item = np.minimum(cnt_of_item, nu - 1)
    ucdf = np.cumsum(rng.lognormal(4.5, 1.0, nu - 1))
    ucdf /= ucdf[-1]
    gender = (rng.random(nu) < 0.28).astype(np.int64) + 1
    batches = []
    for _ in range(n_steps):
        perm = rng.permutation(ni - 1) + 1
        csum = np.cumsum(cnt_of_item[perm - 1])
        J = int(np.searchsorted(csum, B, side="left")) + 1
        items = perm[:J]
        iid = np.repeat(items, cnt_of_item[items - 1])
        uid = np.searchsorted(ucdf, rng.random(len(iid))).clip(max=nu - 2) + 1
        rating = rng.choice(np.arange(1, 6), size=len(iid), p=[.056, .107, .261, .349, .227]).astype(np.float32)
        batches.append((uid.astype(np.int64), iid.astype(np.int64), rating, gender[uid]))
    if kind == "reference":
        class Cfg(dict):
            def __getitem__(self, k):
                return self.get(k, None)

        class DS:
            inter_feat = {"rating": torch.tensor([1.0, 5.0])}

            def num(self, f):
                return {"user_id": nu, "item_id": ni}[f]

        cfg = Cfg(USER_ID_FIELD="user_id", ITEM_ID_FIELD="item_id", NEG_PREFIX="neg_", RATING_FIELD="rating",
                  LABEL_FIELD="label", device=torch.device("cpu"), embedding_size=d, sst_attr_list=["gender"],
                  fair_weight=1.0, fair_objective="value")
        torch.manual_seed(2020)
        model = RefFOCF(cfg, DS())
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-3)        # trainer.py:139
        inters = [Interaction({"user_id": torch.from_numpy(u), "item_id": torch.from_numpy(i),
                               "rating": torch.from_numpy(r), "gender": torch.from_numpy(g)}) for u, i, r, g in batches]

        def step(it):
            opt.zero_grad()
            loss = model.calculate_loss(it)                                           # focf.py:152-169
            loss.backward()
            opt.step()
            return loss.item()                                                        # trainer.py:191
    else:
        from oracle import torch_port as tp
        model = tp.TorchFOCF((rng.standard_normal((nu, d)) * math.sqrt(2.0 / (nu + d))).astype(np.float32),
                             (rng.standard_normal((ni, d)) * math.sqrt(2.0 / (ni + d))).astype(np.float32), "value", 1.0, 5.0)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-3)
        inters = [tuple(torch.from_numpy(x) for x in b) for b in batches]

        def step(it):
            opt.zero_grad()
            loss = model.loss(it[0], it[1], it[2], it[3])
            loss.backward()
            opt.step()
            return loss.item()
    for it in inters[:args.warmup]:
        step(it)
    t0 = time.perf_counter()
    rows = 0
    for it in inters[args.warmup:]:
        v = step(it)
        if v != v:
            raise ValueError("Training loss is nan")
        rows += len(it[0]) if kind == "port" else len(it)
    dt = time.perf_counter() - t0
    val = rows / dt
    sample = (f"{args.steps} steps of a 1/{scale_div}-scale replica of the workload ({nu - 1} users x {ni - 1} items, d {d}, "
              f"~{B} rows per batch = whole items): the reference's step cost (dense autograd gradient + dense Adam over all "
              f"rows, O(B) forward/backward) is linear in both table rows and batch rows, so interactions/s carries over; "
              f"pre-built host Interactions -> calculate_loss -> backward -> Adam -> loss.item() (the bracket of e2e)")
    base = {"value": val, "unit": "interactions/s", "cores": cores, "kind": kind, "sample": sample}
    if kind == "port":
        base["note"] = f"oracle/_ref not importable here ({why}); oracle/torch_port.py restates the same op sequence"
    return {"metric": "FOCF train interactions/s", "value": val, "unit": "interactions/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(w, args.gpus), "cpu_baseline": base,
            "e2e": {"value": val, "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def workload_config(w, world):
    return {"workload": "focf_scaleout (BASELINE.json configs[4])" if w["n_inter"] == SCALEOUT["n_inter"]
            else "focf_scaleout REDUCED by --scale (not the BASELINE configuration)",
            "n_users": w["n_users"], "n_items": w["n_items"], "n_inter": w["n_inter"], "d": w["d"],
            "train_batch_size": w["batch"] * world, "fair_objective": "value",
            "optimizer": "adam(lr=1e-3, weight_decay=1e-3)"}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args, w):
    import torch
    import torch.distributed as dist

    import recbole_fairrec_b200 as pkg
    from recbole_fairrec_b200 import _lib, sharded, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        import datetime
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
        group = dist.group.WORLD

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def reduce_(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    def note(msg):
        if rank == 0:
            print(f"[bench {time.perf_counter() - t_all:7.1f}s] {msg}", file=sys.stderr, flush=True)

    vmax = lambda x: reduce_(x, dist.ReduceOp.MAX)
    vsum = lambda x: reduce_(x, dist.ReduceOp.SUM)
    hbm, bf16, peak_src = peaks()
    t_all = time.perf_counter()
    nu, ni, d, K = w["n_users"], w["n_items"], w["d"], w["K"]
    W_, K_ = args.warmup, args.steps
    data = synth.device_interactions(nu, ni, w["n_inter"], 2020, dev)
    tr_u, tr_i, tr_r = data["train"]
    va_u, va_i = data["valid"]
    gender = data["gender"]
    torch.cuda.synchronize()
    t_synth = time.perf_counter() - t_all
    cfg = pkg.Config(embedding_size=d, fair_objective="value", fair_weight=1.0, topk=[K], valid_metric=f"NDCG@{K}",
                     train_batch_size=w["batch"], learning_rate=1e-3, weight_decay=1e-3, device=dev, seed=2020,
                     score_mode="tc", cuda_graph=False)
    n_plan = W_ + 3 * K_ + 4
    out_modes, prof_modes = {}, {}
    uniq_users = None

    # ============================================================ training
    if world == 1:
        tdata = pkg.TrainData.from_device(tr_u, tr_i.long(), tr_r, gender, nu, ni)
        n_train = tdata.n_rows
        loader = pkg.FOCFDataLoader(cfg, tdata, mode="fast", seed=2020)
        with torch.device(dev):
            model = pkg.FOCF(cfg, synth.SynthDataset(nu, ni, 5.0))
        items, offs, batches = loader.plan_epoch(n_plan)
        d_items, d_offs = torch.from_numpy(items).to(dev), torch.from_numpy(offs).to(dev)
        uf, itf, rf, sf = tdata.fields
        losses = torch.zeros(n_plan, device=dev)

        def make_step(model):
            def step(k):
                uid, iid, rating, sst = loader.gather(d_items, d_offs, batches[k])
                inter = pkg.Interaction({uf: uid, itf: iid, rf: rating, sf: sst})
                inter.items_contiguous = True
                model.train_step(inter, loss_out=losses[k:k + 1])
                return batches[k][3]
            return step

        def run_mode(mode):
            model.init_adam(lr=1e-3, weight_decay=1e-3, mode=mode, max_steps=n_plan + 8)
            step = make_step(model)
            for k in range(W_):
                step(k)
            model.flush_adam()
            barrier()
            model.check_flags()
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            launches0 = _lib.launch_count()
            a.record()
            rows = sum(step(W_ + k) for k in range(K_))
            b.record()
            model.flush_adam()
            c.record()
            barrier()
            launches = _lib.launch_count() - launches0
            ms_steps, ms_all = a.elapsed_time(b), a.elapsed_time(c)
            _lib.profile_enable(True)
            rows_p = sum(step(W_ + K_ + k) for k in range(K_))
            model.flush_adam()
            prof = _lib.profile_report()
            _lib.profile_enable(False)
            return dict(rows=rows, ms_steps=ms_steps, ms_all=ms_all, launches=launches, rows_p=rows_p), prof

        u0 = loader.gather(d_items, d_offs, batches[0])[0]
        uniq_users = int(torch.unique(u0).numel())
        J_avg = float(np.mean([b[2] for b in batches]))
        with ClockSampler(local) as clocks:
            out_modes["dense_exact"], prof_modes["dense_exact"] = run_mode("dense_exact")
        out_modes["lazy_exact"], prof_modes["lazy_exact"] = run_mode("lazy_exact")
        rows_per_gpu_tab = nu + ni
        parallelism = "single GPU"
        # ---- the same dense_exact steps through the planned runner (FOCF.planned_runner, persistent=False): the whole epoch
        # planned on the device, the step captured into CUDA graphs of 8 with the preparation of batch t + 1 (gather, sort,
        # segments: independent of the tables) on a second stream under the Adam sweep of batch t
        pipelined = None
        if os.environ.get("FR_BENCH_PIPELINED", "1") != "0":
            try:
                model.init_adam(lr=1e-3, weight_decay=1e-3, mode="dense_exact", max_steps=n_plan + 4096)
                lbuf = torch.zeros(8192, device=dev)
                runner = model.planned_runner(loader, lbuf, graph_steps=8, persistent=False)
                runner.run(W_ + ((W_ + runner.cursor) & 1))          # (an even cursor: the 8-step graphs start on workspace 0)
                torch.cuda.synchronize()
                c0 = runner.cursor
                pa, pb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                pa.record()
                prow = runner.run(K_)
                pb.record()
                torch.cuda.synchronize()
                model.check_flags()
                pms = pa.elapsed_time(pb)
                pl = lbuf[c0:c0 + K_].cpu().numpy()
                pipelined = {"value": prow / (pms * 1e-3), "unit": "interactions/s", "ms_per_step": pms / K_, "steps": K_,
                             "rows": int(prow), "losses_finite": bool(np.isfinite(pl).all()), "loss_first_last": [float(pl[0]), float(pl[-1])],
                             "what": "FOCF.planned_runner(persistent=False): CUDA graphs of 8 steps, prepare(t + 1) on a second stream "
                                     "under compute(t); same kernels and arithmetic as `value` (bit-identical steps: "
                                     "tests/test_focf_train_gpu.py::test_planned_graph_epoch_equals_stepwise_epoch)"}
            except Exception as e:
                pipelined = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
    else:
        sd = sharded.ShardedTrainData(tr_u, tr_i, tr_r.float(), gender, nu, ni, rank, world, dev)
        n_train = sd.n_rows
        loader = sharded.ShardedFOCFLoader(w["batch"] * world, sd, 2020)
        plan = loader.plan(n_plan)
        losses = torch.zeros(n_plan, device=dev)
        uf, itf, rf, sf = "user_id", "item_id", "rating", "gender"
        model = None
        pipelined = None

        def run_mode(mode):
            m = sharded.ShardedFOCF(sd, d, objective="value", fair_weight=1.0, lr=1e-3, weight_decay=1e-3, adam_mode=mode,
                                    J_cap=loader.J_cap, max_batch_loc=loader.max_batch_loc, max_steps=n_plan + 8)
            try:
                m.init_xavier(nu, ni, 2020)
                m.connect(group)
                m.stage(plan, 0)
                step = lambda k: m.train_step(plan, k, next_k=k + 1, loss_out=losses[k:k + 1])
                for k in range(W_):
                    step(k)
                m.flush()
                barrier()
                m.check_flags()
                a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                launches0 = _lib.launch_count()
                a.record()
                for k in range(K_):
                    step(W_ + k)
                b.record()
                m.flush()
                c.record()
                barrier()
                launches = _lib.launch_count() - launches0
                rows = sum(plan["desc"][W_ + k]["B_glob"] for k in range(K_))
                ms_steps, ms_all = a.elapsed_time(b), a.elapsed_time(c)
                _lib.profile_enable(True)
                for k in range(K_):
                    step(W_ + K_ + k)
                m.flush()
                prof = _lib.profile_report()
                _lib.profile_enable(False)
                barrier()
                m.check_flags()
                # ---- end to end from host batches (this rank's rows of each batch in pinned host memory)
                e2e = None
                if mode == "dense_exact":
                    base = W_ + 2 * K_
                    hbs = [m.local_batch_to_host(plan, base + k) for k in range(K_)]
                    loss_dev = [torch.zeros(1, device=dev) for _ in range(2)]
                    loss_host = [torch.zeros(1).pin_memory() for _ in range(2)]
                    done = [torch.cuda.Event() for _ in range(2)]
                    barrier()
                    t0 = time.perf_counter()
                    pending = None
                    for k, hb in enumerate(hbs):
                        slot = k & 1
                        m.train_step_host(plan, base + k, hb, next_k=base + k + 1, loss_out=loss_dev[slot])
                        loss_host[slot].copy_(loss_dev[slot], non_blocking=True)
                        done[slot].record()
                        if pending is not None:
                            done[pending].synchronize()
                            if float(loss_host[pending]) != float(loss_host[pending]):
                                raise ValueError("Training loss is nan")
                        pending = slot
                    done[pending].synchronize()
                    barrier()
                    t_e = vmax(time.perf_counter() - t0)
                    rows_e = sum(plan["desc"][base + k]["B_glob"] for k in range(K_))
                    e2e = {"value": rows_e / t_e, "unit": "interactions/s",
                           "h2d_bytes_per_step": int(np.mean([hb[0].numel() for hb in hbs])), "d2h_bytes_per_step": 4,
                           "mode": "per rank: its rows of the batch in ONE pinned host buffer -> H2D -> row-sharded step -> loss "
                                   "D2H every step; the host waits for step t's loss after enqueuing step t+1",
                           "api": "ShardedFOCF.train_step_host(plan, k, host_batch)"}
                    m.check_flags()
            finally:
                m.close()
            return dict(rows=rows, ms_steps=ms_steps, ms_all=ms_all, launches=launches, rows_p=rows, e2e=e2e), prof

        with ClockSampler(local) as clocks:
            out_modes["dense_exact"], prof_modes["dense_exact"] = run_mode("dense_exact")
        out_modes["lazy_exact"], prof_modes["lazy_exact"] = run_mode("lazy_exact")
        rows_per_gpu_tab = sharded.local_rows(nu, rank, world) + sharded.local_rows(ni, rank, world)
        J_avg = float(np.mean([b["J"] for b in plan["desc"]]))
        parallelism = (f"row-sharded x{world}: rank r owns rows r::{world} of both tables and the Adam moments and the train rows "
                       f"of its users; per step the drawn items' rows, the partial item x group sums and the partial item "
                       f"gradients (O(J d), J ~ {J_avg:.0f}) cross NVLink as peer-memory stores of the step's kernels, 3 flag "
                       f"barriers per step, no collective")

    note("training modes done")
    # ---- headline numbers (dense_exact) and the lazy block
    dn, lz = out_modes["dense_exact"], out_modes["lazy_exact"]
    t_dev = vmax(dn["ms_steps"] / 1e3)
    value = dn["rows"] / t_dev
    B_avg = dn["rows"] / K_ / world          # rows per GPU per step

    def roofline_of(prof, n_steps, rows_tab, B_gpu, touched_users):
        """dominant kernel + algorithmic bytes per launch"""
        rows_touched = touched_users + J_avg
        alg = {   # by kernel-name prefix; kernels launched once per table side report the mean over their two launches
            "k_apply": 24.0 * rows_tab * d,
            "k_shard_apply_users": 24.0 * (rows_tab - ni / world) * d, "k_shard_apply_items": 24.0 * (ni / world) * d,
            "k_forward": (8.0 * d + 16.0) * B_gpu, "k_shard_forward": (8.0 * d + 16.0) * B_gpu,
            "k_segment_grads": 8.0 * d * B_gpu + 4.0 * d * rows_touched,
            "k_shard_grads": 8.0 * d * B_gpu + 4.0 * d * rows_touched,
            "(k_adam_lazy<0, false>)": 24.0 * d * rows_touched / 2, "(k_adam_lazy<0, true>)": 28.0 * d * rows_touched / 2,
            "(k_adam_lazy<2, false>)": 24.0 * rows_tab * d / 2,
        }
        tot = sum(v[1] for v in prof.values()) or 1.0
        shares = {k: round(v[1] / tot, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}

        def key_of(k):
            return next((a for a in alg if k.startswith(a)), None)

        dom = next((k for k in shares if key_of(k)), None)
        if dom is None:
            return None, shares
        cnt, ms = prof[dom]
        ach = alg[key_of(dom)] / (ms / cnt * 1e-3) / 1e9
        return ({"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                 "traffic": None, "algorithmic_bytes_per_launch": alg[key_of(dom)], "avg_launch_us": 1e3 * ms / cnt,
                 "launches_profiled": cnt, "share_of_step": shares[dom], "peak_source": peak_src,
                 "from": f"library CUDA-event profiler over {n_steps} further steps of the same loop (same batches regime, no "
                         f"flush needed: every launch streams far more than the 126 MB L2)"}, shares)

    touched = (uniq_users or B_avg * 0.5)
    roof, shares = roofline_of(prof_modes["dense_exact"], K_, rows_per_gpu_tab, B_avg, touched)
    if roof and world == 1 and roof["kernel"].startswith("k_apply") and w["n_inter"] == SCALEOUT["n_inter"]:
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this shape from the committed `ncu --set full`
        # capture (profiles/r02_ncu_summary.md, last section): 17.4823 GB + 16.8369 GB per launch
        roof["traffic"] = 34.3192e9
        roof["traffic_note"] = ("ncu --set full capture of this kernel at this shape (profiles/r02_ncu_summary.md): DRAM read "
                                "17.48 GB + write 16.84 GB per launch = 1.016 x the algorithmic bytes")
    roof_lz, shares_lz = roofline_of(prof_modes["lazy_exact"], K_, rows_per_gpu_tab, B_avg, touched)
    t_lz = vmax(lz["ms_all"] / 1e3)
    lazy = {"value": lz["rows"] / t_lz, "unit": "interactions/s", "ms_per_step": 1e3 * t_lz / K_,
            "steps_only_ms_per_step": vmax(lz["ms_steps"]) / K_, "speedup_vs_dense_exact": (lz["rows"] / t_lz) / value,
            "what": "adam_mode lazy_exact: tables, moments and losses BIT-IDENTICAL to dense_exact (tests/test_sharded_gpu.py); a "
                    "row is replayed through the steps it missed when a batch next touches it, or at the flush.  The timed region "
                    "is the K steps PLUS the closing flush, which carries the deferred updates of all rows for those K steps (the "
                    "replay is float32 ALU work, ~35 instructions per parameter and step, instead of 24 bytes of HBM traffic).",
            "roofline": roof_lz, "kernel_shares": dict(list(shares_lz.items())[:8])}

    # ============================================================ e2e (N = 1): host batches through FOCF.train_step
    if world == 1:
        model.init_adam(lr=1e-3, weight_decay=1e-3, mode="dense_exact")
        base = W_ + 2 * K_
        hbs = []
        for k in range(K_):
            cols = loader.gather(d_items, d_offs, batches[base + k])
            hbs.append(pkg.pack_host_batch(*[c.cpu() for c in cols], fields=(uf, itf, rf, sf)))
        loss_dev = [torch.zeros(1, device=dev) for _ in range(2)]
        loss_host = [torch.zeros(1).pin_memory() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        model.train_step(hbs[0], loss_out=loss_dev[0])          # staging buffers
        barrier()
        t0 = time.perf_counter()
        pending, rows_e = None, 0
        for k, hb in enumerate(hbs):
            slot = k & 1
            model.train_step(hb, loss_out=loss_dev[slot])       # the single H2D copy of the packed batch happens inside
            loss_host[slot].copy_(loss_dev[slot], non_blocking=True)
            done[slot].record()
            if pending is not None:
                done[pending].synchronize()
                if float(loss_host[pending]) != float(loss_host[pending]):
                    raise ValueError("Training loss is nan")
            pending = slot
            rows_e += len(hb)
        done[pending].synchronize()
        barrier()
        t_e = time.perf_counter() - t0
        e2e = {"value": rows_e / t_e, "unit": "interactions/s",
               "h2d_bytes_per_step": int(np.mean([hb.packed_host[0].numel() for hb in hbs])), "d2h_bytes_per_step": 4,
               "mode": "host batch (one pinned buffer: int32 user | int32 item | f32 rating | f32 attribute) -> H2D -> step -> "
                       "loss D2H every step; the host waits for step t's loss after enqueuing step t+1",
               "api": "FOCF.train_step(interaction) with adam_mode dense_exact"}
        model.check_flags()
        del hbs
    else:
        e2e = dn["e2e"]

    note("e2e done")
    # ============================================================ evaluation
    ev = None
    try:
        n_eval_cap = int(args.eval_users)
        keep = va_u <= n_eval_cap
        edata = pkg.EvalData.from_device(tr_u, tr_i, va_u[keep], va_i[keep], {sf: gender}, nu, ni)
        del keep
        counts = torch.bincount(tr_i.long(), minlength=ni).cpu().numpy()
        del tr_u, tr_i, tr_r, va_u, va_i, data
        if world == 1:
            Uw, Iw = model.user_embedding_layer.weight.data, model.item_embedding_layer.weight.data
        else:
            g = torch.Generator(device=dev).manual_seed(77)
            Uw = torch.randn((nu, d), generator=g, device=dev) * math.sqrt(2.0 / (nu + d))
            Iw = torch.randn((ni, d), generator=g, device=dev) * math.sqrt(2.0 / (ni + d))
        torch.cuda.empty_cache()
        evaluator = pkg.FullSortEvaluator(cfg, ni, {int(i): int(c) for i, c in enumerate(counts) if c > 0}, group=group)
        evaluator.evaluate(Uw, Iw, edata, 5.0)                   # warm-up pass (popularity mask, module load)
        barrier()
        n_pass = 2
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        for _ in range(n_pass):
            res = evaluator.evaluate(Uw, Iw, edata, 5.0)         # returns the metric dict on the host (D2H inside)
        b.record()
        barrier()
        t_wall = vmax((time.perf_counter() - t0) / n_pass)
        t_evd = vmax(a.elapsed_time(b) / 1e3 / n_pass)
        _lib.profile_enable(True)
        evaluator.collect(Uw, Iw, edata, 5.0)
        prof_e = _lib.profile_report()
        _lib.profile_enable(False)
        tc_cnt, tc_ms = prof_e.get("k_fullsort_tc", (1, float("nan")))
        tc_peak = bf16 / 2.0 / 3.0
        try:
            bf16_sus = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained") or 0.0)
        except Exception:
            bf16_sus = 0.0
        flops = 2.0 * edata.n * (ni / world) * d
        etot = sum(v[1] for v in prof_e.values()) or 1.0
        ach = flops / (tc_ms / tc_cnt * 1e-3) / 1e12
        ev = {"metric": "full-sort fair-eval users/s", "value": edata.n / t_evd, "unit": "users/s", "n_users": edata.n,
              "n_items": ni, "n_pos": edata.n_pos, "history_entries": int(edata.hist_items.numel()),
              "ms_per_pass": 1e3 * t_evd, "score_mode": "tc_3xtf32",
              "e2e": {"value": edata.n / t_wall, "unit": "users/s",
                      "what": "FullSortEvaluator.evaluate(U, I, data): wall clock incl. the read-back of the metric sums"},
              "sharding": None if world == 1 else f"item table in {world} contiguous id ranges, NCCL all-gather of the per-shard "
              f"top-{K} + merge, all-reduce of the item x group sums",
              "roofline": {"bound": "tensor", "kernel": "k_fullsort_tc", "achieved": ach, "peak": tc_peak,
                           "unit": "TFLOP/s (fp32-equivalent; 3 TF32 MMAs per product)", "frac": ach / tc_peak,
                           "avg_launch_us": 1e3 * tc_ms / tc_cnt, "share_of_pass": round(tc_ms / etot, 4),
                           "peak_source": f"{peak_src}: bf16 {bf16} / 2 (tf32) / 3 (3xTF32)",
                           "nominal_peak": 2250.0 / 2.0 / 3.0, "frac_of_nominal": ach / (2250.0 / 2.0 / 3.0),
                           "peak_sustained": bf16_sus / 6.0 if bf16_sus else None,
                           "frac_of_sustained": ach / (bf16_sus / 6.0) if bf16_sus else None,
                           "note": "peak = the BURST bf16 figure / 6 although a pass is a 0.5-s kernel (the sustained figure / 6 "
                                   "is peak_sustained); the same kernel timed alone with idle gaps reaches 297-300 TFLOP/s "
                                   "(profiles/r02_tc_scorer.md), back to back the board's power management takes ~10 %; "
                                   "nominal_peak = 2.25 PFLOP/s bf16 / 2 / 3 (B200_PROFILING.md)"},
              "kernel_shares": {k: round(v[1] / etot, 4) for k, v in sorted(prof_e.items(), key=lambda kv: -kv[1][1])[:6]},
              "metrics": {k: float(v) for k, v in res.items()}}
        del edata, evaluator
    except Exception as e:
        ev = {"error": f"{type(e).__name__}: {str(e)[:300]}"}

    note("evaluation done")
    # ============================================================ dp_check (N > 1)
    dp_check = None
    if world > 1:
        try:
            torch.cuda.empty_cache()
            dp_check = {"train": sharded.selfcheck(rank, world, dev, group, adam_mode="dense_exact"),
                        "train_lazy": sharded.selfcheck(rank, world, dev, group, adam_mode="lazy_exact"),
                        "eval": eval_selfcheck(rank, world, dev, group)}
            if rank == 0:
                dp_check["pass"] = all(bool(v.get("pass")) for v in dp_check.values() if isinstance(v, dict))
        except Exception as e:
            dp_check = {"pass": False, "error": f"{type(e).__name__}: {str(e)[:300]}"}

    # ============================================================ N > 1: PFCN_MLP data parallel (configs[2] shape per rank)
    families_dp = None
    if world > 1 and not args.no_families:
        try:
            import bench_families
            model = loader = tdata = make_step = run_mode = None
            Uw = Iw = d_items = d_offs = None
            import gc
            gc.collect()
            torch.cuda.empty_cache()
            note("families_dp block (PFCN_MLP data parallel)")
            families_dp = {"pfcn_mlp": bench_families.bench_pfcn_dp(dev, rank, world, group)}
        except Exception as e:
            families_dp = {"error": f"{type(e).__name__}: {str(e)[:300]}"}

    # ============================================================ N = 1 only: configs[1], families, CPU baseline
    ml1m = families = cpu = None
    if world == 1 and rank == 0:
        model = loader = tdata = make_step = run_mode = None       # (closures hold the tables)
        Uw = Iw = d_items = d_offs = None
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        # the two side blocks run in their OWN processes after this one has released the GPU memory: a failure there cannot
        # take the headline line down, and their CUDA-graph pools / side streams start from a clean context
        def side_block(script, extra):
            cmd = [sys.executable, os.path.join(ROOT, script), "--steps", str(max(K_, 20)), "--warmup", str(max(W_, 5))] + extra
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
            lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            if r.returncode != 0 or not lines:
                return {"error": f"{script} exited with {r.returncode}: {r.stderr[-300:]}"}
            return json.loads(lines[-1])

        if not args.no_ml1m:
            try:
                note("ml1m block (bench_ml1m.py, own process)")
                ml1m = side_block("bench_ml1m.py", ["--no-probe", "--no-families"] +
                                  (["--no-cpu-baseline"] if args.no_cpu_baseline else []))
            except Exception as e:
                ml1m = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
        if not args.no_families:
            try:
                note("families block (bench_families.py, own process)")
                families = side_block("bench_families.py", ["--no-cpu-baseline"] if args.no_cpu_baseline else [])
            except Exception as e:
                families = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
        if not args.no_cpu_baseline:
            try:
                note("cpu baseline (reference arm)")
                sub = argparse.Namespace(steps=min(K_, 10), warmup=2, gpus=1)
                cpu = reference_arm(sub, w)["cpu_baseline"]
            except Exception as e:
                cpu = {"error": f"{type(e).__name__}: {str(e)[:300]}"}

    out = {
        "metric": "FOCF train interactions/s", "value": value, "unit": "interactions/s", "n_gpus": world, "steps": K_,
        "warmup": W_, "ms_per_step": 1e3 * t_dev / K_, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (generated on the device, seed 2020)",
        "gpu_launches": int(dn["launches"]),
        # `config` is the workload both arms run (identical dicts in this line and in --impl reference's); how THIS arm runs it
        # is `config_detail`
        "config": workload_config(w, world),
        "config_detail": {"n_train": int(n_train), "avg_batch_rows_per_gpu": B_avg,
                          "adam_mode": "dense_exact (every row moves every step, as torch.optim.Adam does in the reference)",
                          "l2": f"no flush needed: every step streams 24 * rows * d = {24.0 * rows_per_gpu_tab * d / 1e9:.1f} GB "
                                f"of tables + moments per GPU, far beyond the 126 MB L2",
                          "loop": ("value = one FOCF.train_step per batch on one stream (prepare -> forward -> loss -> gradients -> Adam "
                                   "in series); FOCFTrainer._train_epoch itself runs the planned runner of the `pipelined` block "
                                   "(same kernels, the next batch's preparation under the Adam sweep)") if world == 1 else
                                  "row-sharded step, the next batch staged during phase C (next_k)",
                          "parallelism": parallelism},
        "clocks": clocks.summary(),
        "roofline": roof,
        "kernel_shares": dict(list(shares.items())[:10]),
        "cpu_baseline": cpu,
        "e2e": e2e,
        "lazy_exact": lazy,
        "pipelined": pipelined if world == 1 else None,
        "eval": ev,
        "dp_check": dp_check,
        "ml1m": ml1m,
        "families": families,
        "families_dp": families_dp,
        "setup_s": {"synthesis": t_synth, "total_before_json": time.perf_counter() - t_all},
        "hbm_allocated_gb": torch.cuda.max_memory_allocated() / 1e9,
    }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out if rank == 0 else None


def eval_selfcheck(rank, world, dev, group, n_users=3001, n_items=4099, d=64, K=10):
    """item-sharded evaluation (NCCL all-gather + merge, all-reduce of the item x group sums) vs the unsharded pass on the
    same tables: top-K ids bit-equal, the 12 metrics equal (SURVEY.md section 4 test plan iv)"""
    import torch
    import recbole_fairrec_b200 as pkg
    from recbole_fairrec_b200 import synth
    uid, iid, rating, gender = synth.interactions(n_users, n_items, 60000, 3)
    tr, va, te = synth.split_by_user(uid, iid, rating, seed=3)
    users, hist, pos = synth.eval_lists(tr, va, te, "valid")
    cfg = pkg.Config(embedding_size=d, topk=[K], device=dev, metric_decimal_place=12, score_mode="exact", cuda_graph=False)
    edata = pkg.EvalData(users, hist, pos, {"gender": gender.astype(np.int64)}, dev)
    g = torch.Generator(device=dev).manual_seed(5)
    U = torch.randn((n_users, d), generator=g, device=dev) * 0.3
    I = torch.randn((n_items, d), generator=g, device=dev) * 0.3
    counts = np.bincount(tr[1], minlength=n_items)
    ci = {int(i): int(c) for i, c in enumerate(counts) if c > 0}
    sharded_ev = pkg.FullSortEvaluator(cfg, n_items, ci, group=group)
    res_s = sharded_ev.evaluate(U, I, edata, 5.0)
    ids_s = sharded_ev.last["topk_id"].clone()
    out = {"world": world, "shape": [n_users, n_items, d, K]}
    if rank == 0:
        single = pkg.FullSortEvaluator(cfg, n_items, ci, group=None)
        res_1 = single.evaluate(U, I, edata, 5.0)
        out["ids_bit_equal"] = bool(torch.equal(ids_s, single.last["topk_id"]))
        out["metrics_max_abs_diff"] = max(abs(float(res_s[k]) - float(res_1[k])) for k in res_1)
        out["pass"] = out["ids_bit_equal"] and out["metrics_max_abs_diff"] <= 1e-9
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (smoke runs; not the BASELINE config)")
    ap.add_argument("--eval-users", type=float, default=1 << 19, help="evaluate users with id <= this")
    ap.add_argument("--ref-scale-div", type=int, default=8, help="reference arm: 1/S-scale replica of the workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ml1m", action="store_true")
    ap.add_argument("--no-families", action="store_true")
    args = ap.parse_args()
    w = scaled(args.scale)
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        print(json.dumps(reference_arm(args, w, args.ref_scale_div)), flush=True)
        return
    args.warmup = max(args.warmup, 3)
    # libraries (NCCL's version banner, ...) may write to stdout: keep fd 1 for the ONE JSON line
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        out = run_ours(args, w)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
