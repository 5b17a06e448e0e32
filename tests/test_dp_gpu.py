"""Multi-GPU (needs >= 2 CUDA devices; run with `gpurun --gpus 2`): the data-parallel FOCF step over NCCL equals the
fused single-GPU step on the union of the ranks' batches, and item-sharded evaluation equals unsharded evaluation."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out):
    import torch.distributed as dist
    import recbole_fairrec_b200 as pkg
    from recbole_fairrec_b200 import synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    nu, ni, d = 700, 300, 32
    uid, iid, rating, gender = synth.interactions(nu, ni, 40000, 5, item_sigma=1.0)
    tr, va, te = synth.split_by_user(uid, iid, rating, seed=5)
    cfg = pkg.Config(embedding_size=d, fair_objective="value", fair_weight=1.0, train_batch_size=1024, device=dev,
                     learning_rate=1e-3, weight_decay=1e-3, epochs=1, topk=[10], metric_decimal_place=12)
    train = pkg.TrainData(tr[0], tr[1], tr[2], gender, nu, ni, dev)
    rng = np.random.default_rng(0)
    U0 = (rng.standard_normal((nu, d)) * 0.2).astype(np.float32)
    I0 = (rng.standard_normal((ni, d)) * 0.2).astype(np.float32)

    def fresh():
        m = pkg.FOCF(cfg, synth.SynthDataset(nu, ni, 5.0))
        with torch.no_grad():
            m.user_embedding_layer.weight.copy_(torch.from_numpy(U0))
            m.item_embedding_layer.weight.copy_(torch.from_numpy(I0))
        return m.to(dev)

    # ---- data-parallel epoch
    model = fresh()
    trainer = pkg.FOCFTrainer(cfg, model, group=dist.group.WORLD)
    loader = pkg.FOCFDataLoader(cfg, train, mode="fast", seed=100 + rank, partition=(rank, world))
    loss_dp = trainer._train_epoch(loader, 0)
    U_dp = model.user_embedding_layer.weight.detach().cpu().numpy()
    I_dp = model.item_embedding_layer.weight.detach().cpu().numpy()

    # ---- the same global batches on one GPU: union of the ranks' draws, step by step (same seeds => same draws)
    loaders = [pkg.FOCFDataLoader(cfg, train, mode="fast", seed=100 + r, partition=(r, world)) for r in range(world)]
    plans = [l.plan_epoch() for l in loaders]
    ref = fresh()
    ref.init_adam(lr=1e-3, weight_decay=1e-3)
    n = len(loaders[0])
    losses = torch.zeros(n, device=dev)
    for k in range(n):
        cols = []
        for l, (items, offs, batches) in zip(loaders, plans):
            c = l.gather(torch.from_numpy(items).to(dev), torch.from_numpy(offs).to(dev), batches[k])
            cols.append([x.clone() for x in c])
        u, i, r, s = (torch.cat([c[j] for c in cols]) for j in range(4))
        inter = pkg.Interaction({"user_id": u, "item_id": i, "rating": r, "gender": s})
        inter.items_contiguous = True
        ref.train_step(inter, loss_out=losses[k:k + 1])
    loss_ref = float(losses.cpu().numpy().astype(np.float64).sum())
    U_ref = ref.user_embedding_layer.weight.detach().cpu().numpy()
    I_ref = ref.item_embedding_layer.weight.detach().cpu().numpy()

    # ---- item-sharded evaluation vs unsharded
    users, hist, pos = synth.eval_lists(tr, va, te, "valid")
    data = pkg.EvalData(users, hist, pos, {"gender": gender.astype(np.int64)}, dev)
    ev1 = pkg.FullSortEvaluator(cfg, ni, train.item_counter)
    evp = pkg.FullSortEvaluator(cfg, ni, train.item_counter, group=dist.group.WORLD)
    Uw, Iw = model.user_embedding_layer.weight.data, model.item_embedding_layer.weight.data
    r1 = ev1.evaluate(Uw, Iw, data, 5.0)
    rp = evp.evaluate(Uw, Iw, data, 5.0)
    ids_equal = bool(torch.equal(ev1.last["topk_id"], evp.last["topk_id"]))
    model.release_graphs()      # captured NCCL kernels must go before the communicator
    if rank == 0:
        out.put(dict(loss_dp=loss_dp, loss_ref=loss_ref, U_dp=U_dp, U_ref=U_ref, I_dp=I_dp, I_ref=I_ref, r1=dict(r1),
                     rp=dict(rp), ids_equal=ids_equal))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_dp_step_and_sharded_eval_match_single_gpu():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29533, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    np.testing.assert_allclose(res["loss_dp"], res["loss_ref"], rtol=1e-5)
    for a, b in (("U_dp", "U_ref"), ("I_dp", "I_ref")):
        err = np.abs(res[a] - res[b]).max() / np.abs(res[b]).max()
        assert err < 1e-5, (a, err)
    assert res["ids_equal"]
    for k in res["r1"]:
        assert abs(res["r1"][k] - res["rp"][k]) <= 1e-9 * max(abs(res["r1"][k]), 1.0), k


def _shard_worker(rank, world, port, out, adam_mode):
    import torch.distributed as dist
    from recbole_fairrec_b200 import sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    res = sharded.selfcheck(rank, world, dev, dist.group.WORLD, adam_mode=adam_mode, n_users=5001, n_items=901, d=64,
                            n_inter=120000, batch=1 << 13, steps=5)
    if rank == 0:
        out.update(res)
    dist.destroy_process_group()


@pytest.mark.parametrize("adam_mode", ["dense_exact", "lazy_exact"])
def test_row_sharded_step_multi_process(adam_mode):
    """the REAL exchange path: one process per GPU, CUDA IPC peer memory, cross-GPU flag barriers (csrc/focf_shard.cu)"""
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_shard_worker, args=(world, 29541 + (adam_mode == "lazy_exact"), out, adam_mode), nprocs=world, join=True)
    assert out["pass"], dict(out)
