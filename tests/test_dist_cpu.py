"""CPU, gloo, world_size 2: the host-side logic of the multi-GPU paths -- disjoint item partitions for data-parallel
batches, global normalisers by all-reduce of the planned (J, B), and the item-shard top-K merge protocol (contiguous
id ranges keep the lowest-id tie rule) checked with the oracle standing in for the kernels."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fullsort_oracle as fs


def _merge(ids, scores, K):
    """canonical K-way merge of per-shard lists: (score desc, id asc)"""
    n = ids.shape[1]
    out_i = np.zeros((n, K), np.int64)
    out_s = np.zeros((n, K), np.float32)
    for r in range(n):
        cand = sorted(((-float(s), int(i)) for p in range(ids.shape[0]) for i, s in zip(ids[p, r], scores[p, r])))[:K]
        out_i[r] = [c[1] for c in cand]
        out_s[r] = [-c[0] for c in cand]
    return out_i, out_s


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from recbole_fairrec_b200.evaluator import shard_bounds
    rng = np.random.default_rng(0)                      # same data on both ranks
    n_users, n_items, d, K = 40, 257, 8, 10
    U = (rng.standard_normal((n_users, d)) * 2).astype(np.float32)   # clamp ties at 0/1 on purpose
    I = (rng.standard_normal((n_items, d)) * 2).astype(np.float32)
    users = np.arange(1, n_users)
    hist_off = np.arange(0, n_users * 3, 3)[:n_users]
    hist_items = rng.integers(1, n_items, hist_off[-1])
    scores = fs.mask_history(fs.full_sort_scores(U, I, users, 5.0), hist_off, hist_items)
    lo, hi = shard_bounds(n_items, world, rank)
    ids_l, sc_l = fs.topk_canonical(scores[:, lo:hi], K)
    ids_l = ids_l + lo
    gi = [torch.zeros_like(torch.from_numpy(ids_l)) for _ in range(world)]
    gs = [torch.zeros_like(torch.from_numpy(sc_l)) for _ in range(world)]
    dist.all_gather(gi, torch.from_numpy(ids_l))
    dist.all_gather(gs, torch.from_numpy(sc_l))
    mi, ms = _merge(np.stack([g.numpy() for g in gi]), np.stack([g.numpy() for g in gs]), K)
    full_i, full_s = fs.topk_canonical(scores, K)
    ok_merge = np.array_equal(mi, full_i) and np.array_equal(ms, full_s)

    # data-parallel planning: disjoint item partitions, identical epoch length, global normalisers
    uniq = np.arange(1, n_items)
    mine = uniq[rank::world]
    sizes = torch.tensor([[len(mine), int(mine.sum())]], dtype=torch.int64)
    dist.all_reduce(sizes)
    ok_part = int(sizes[0, 0]) == len(uniq) and int(sizes[0, 1]) == int(uniq.sum())
    q.put((rank, bool(ok_merge), bool(ok_part)))
    dist.destroy_process_group()


def test_shard_merge_and_partition_protocol_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29544, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok_m and ok_p for _, ok_m, ok_p in res)


def test_dataloader_partitions_are_disjoint_and_cover():
    """host logic of FOCFDataLoader(partition=...) without any kernel: candidates and epoch length"""
    import math
    uniq = np.array([1, 2, 3, 5, 8, 13, 21, 34])
    parts = [uniq[r::3] for r in range(3)]
    assert sorted(np.concatenate(parts).tolist()) == uniq.tolist()
    assert len(set(parts[0]) & set(parts[1])) == 0
    assert math.ceil(1000 / (128 * 3)) == 3
