#!/usr/bin/env python
"""Profiling driver for the row-sharded FOCF step on ONE GPU (ncu cannot wrap a multi-rank command): P emulated ranks
(recbole-fairrec_b200/sharded.py:ShardedGroupEmu -- the same kernels and exchange layout as the multi-process run, the host
sequencing the phases) on a 1/5-scale replica of BASELINE.json configs[4].

    ncu --set full --clock-control none --import-source on -k regex:k_shard -s <n> -c <m> -o gpurun_out/r02_shard \
        python profiles/tools/prof_shard_emu.py [--world 2] [--mode dense_exact|lazy_exact] [--steps 4]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    import recbole_fairrec_b200 as pkg
    from recbole_fairrec_b200 import synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=2)
    ap.add_argument("--mode", default="dense_exact")
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--scale", type=float, default=0.2)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    nu, ni = int(10_000_000 * args.scale) + 1, int(1_000_000 * args.scale) + 1
    data = synth.device_interactions(nu, ni, int(1_000_000_000 * args.scale), 2020, dev)
    u, i, r = data["train"]
    emu = pkg.ShardedGroupEmu(u, i, r.float(), data["gender"], nu, ni, args.world, dev, 128,
                              int((1 << 20) * args.scale) * args.world, seed=2020, adam_mode=args.mode, max_steps=64)
    try:
        for rk in emu.ranks:
            rk.init_xavier(nu, ni, 2020)
        plans = emu.plan(args.steps)
        emu.train(plans)
        emu.full_tables()
        torch.cuda.synchronize()
        for rk in emu.ranks:
            rk.check_flags()
        print("ok: %d steps, world %d, %s, J ~ %d, local rows ~ %d" % (
            args.steps, args.world, args.mode, plans[0]["desc"][0]["J"], plans[0]["desc"][0]["B_loc"]))
    finally:
        emu.close()


if __name__ == "__main__":
    main()
