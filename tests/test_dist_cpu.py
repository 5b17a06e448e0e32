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


# ------------------------------------------------------------------ row-sharded FOCF training: host-side protocol
def _numpy_sharded_step(datas, plans, k, U, I, world, fair_weight=1.0):
    """numpy restatement of one row-sharded step's EXCHANGE protocol (csrc/focf_shard.cu) from the ranks' plans: every rank's
    partial item x group sums indexed by draw position, summed in rank order -> loss; checked against the single-process
    oracle on the union batch.  Returns (loss_sharded, batch columns of the union in rank order)."""
    J = plans[0]["desc"][k]["J"]
    part = np.zeros((world, J, 7), np.float64)
    cols = []
    for r in range(world):
        b = plans[r]["desc"][k]
        items = plans[r]["items"][b["items_pos"]:b["items_pos"] + J].numpy()
        off = plans[r]["offs"][b["offs_pos"]:b["offs_pos"] + J + 1].numpy()
        d = datas[r]
        item_off = d.item_off.numpy()
        for j, it in enumerate(items):
            lo = item_off[it]
            n = off[j + 1] - off[j]
            lu = d.train_uid.numpy()[lo:lo + n]
            rt = d.train_rating.numpy()[lo:lo + n]
            g = d.sst_of_user.numpy()[lu]
            gu = lu.astype(np.int64) * world + r                       # back to global user ids
            pred = (U[gu] * I[it]).sum(1)
            for grp, val in ((0, 1.0), (1, 2.0)):
                m = g == val
                part[r, j, grp] = pred[m].sum()
                part[r, j, 2 + grp] = rt[m].sum()
                part[r, j, 4 + grp] = m.sum()
            part[r, j, 6] = ((pred - rt) ** 2).sum()
            cols.append((gu, np.full(n, it), rt, g))
    tot = part.sum(0)                                                  # rank order
    B = plans[0]["desc"][k]["B_glob"]
    n = tot[:, 4:6] + 1e-5
    D = tot[:, 0:2] / n - tot[:, 2:4] / n
    x = np.abs(D[:, 0] - D[:, 1])
    hx = np.where(x < 1, 0.5 * x * x, x - 0.5)
    return tot[:, 6].sum() / B + fair_weight * hx.mean(), cols


def _shard_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from recbole_fairrec_b200 import synth
    from recbole_fairrec_b200.sharded import ShardedFOCFLoader, ShardedTrainData, local_rows
    nu, ni = 301, 83
    uid, iid, rating, gender = synth.interactions(nu, ni, 4000, 9, item_sigma=1.0)
    cpu = torch.device("cpu")
    sd = ShardedTrainData(uid, iid, rating, gender, nu, ni, rank, world, cpu)
    plan = ShardedFOCFLoader(500, sd, seed=3).plan(4)
    # every rank draws the SAME items (same seed, GLOBAL item counts); local rows add up to the global batch
    mine = [plan["items"].numpy().tolist(), [b["B_loc"] for b in plan["desc"]], [b["B_glob"] for b in plan["desc"]],
            sd.n_rows_loc, sd.n_users_loc, sd.n_items_loc]
    allp = [None] * world
    dist.all_gather_object(allp, mine)
    ok = all(p[0] == allp[0][0] and p[2] == allp[0][2] for p in allp)
    ok &= all(sum(p[1][k] for p in allp) == allp[0][2][k] for k in range(4))
    ok &= sum(p[3] for p in allp) == len(uid)
    ok &= sum(p[4] for p in allp) == nu and sum(p[5] for p in allp) == ni
    ok &= sd.n_users_loc == local_rows(nu, rank, world)
    # owner slots: position of each drawn item among its owner's items, in draw order
    b0 = plan["desc"][0]
    items = plan["items"][:b0["J"]].numpy()
    slots = plan["slots"][:b0["J"]].numpy()
    for o in range(world):
        ok &= slots[items % world == o].tolist() == list(range(int((items % world == o).sum())))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_row_sharded_plans_agree_across_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, 29546, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)


def test_row_sharded_exchange_protocol_equals_the_oracle_on_the_union_batch():
    """one process, three emulated ranks, numpy only: partial item x group sums by draw position, summed in rank order, give
    the loss the oracle computes for the union batch (what k_shard_stats_push / k_shard_stats_reduce implement)"""
    from oracle import focf_oracle as fo
    from recbole_fairrec_b200 import synth
    from recbole_fairrec_b200.sharded import ShardedFOCFLoader, ShardedTrainData
    nu, ni, d, world = 211, 61, 8, 3
    uid, iid, rating, gender = synth.interactions(nu, ni, 3000, 4, item_sigma=1.0)
    cpu = torch.device("cpu")
    datas = [ShardedTrainData(uid, iid, rating, gender, nu, ni, r, world, cpu) for r in range(world)]
    plans = [ShardedFOCFLoader(400, dt, seed=5).plan(2) for dt in datas]
    rng = np.random.default_rng(0)
    U = (rng.standard_normal((nu, d)) * 0.4).astype(np.float32)
    I = (rng.standard_normal((ni, d)) * 0.4).astype(np.float32)
    for k in range(2):
        loss_s, cols = _numpy_sharded_step(datas, plans, k, U.astype(np.float64), I.astype(np.float64), world)
        u = np.concatenate([c[0] for c in cols]); i = np.concatenate([c[1] for c in cols])
        r = np.concatenate([c[2] for c in cols]); g = np.concatenate([c[3] for c in cols])
        assert len(u) == plans[0]["desc"][k]["B_glob"]
        want = fo.calculate_loss(U, I, u, i, r, g.astype(np.int64), "value", 1.0)
        np.testing.assert_allclose(loss_s, want, rtol=2e-6)
