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


def _pfcn_worker(rank, world, port, out):
    """PFCN_MLP, data parallel over `world` GPUs (PFCNTrainer(dp=ops.ChainDP): rows of every batch split over the ranks,
    BatchNorm column sums through NVLink peer memory, gradient shares summed by NCCL) against the single-GPU trainer on
    the whole batches: first-step gradients, the losses of three alternating filter / discriminator steps, BatchNorm
    running statistics."""
    import copy

    import torch.distributed as dist
    import recbole_fairrec_b200 as pkg
    from recbole_fairrec_b200 import ops
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from test_pfcn_gpu import UserFeatDataset
    nu, ni, d, B = 600, 400, 64, 1024
    rng = np.random.default_rng(7)
    feats = {"gender": (rng.random(nu) < 0.3).astype(np.float32), "age": rng.integers(0, 7, nu).astype(np.float32)}
    cfg = pkg.Config(embedding_size=d, sst_attr_list=list(feats), filter_mode="sm", dropout=0.0, dis_dropout=0.0,
                     dis_weight=10.0, dis_hidden_size_list=[128, 256, 128, 128, 64, 32], mlp_hidden_size_list=[64, 32, 16],
                     activation="leakyrelu", device=dev, learning_rate=1e-3, weight_decay=1e-4, train_epoch_interval=1)
    torch.manual_seed(11)
    model = pkg.PFCN_MLP(cfg, UserFeatDataset(nu, ni, feats)).to(dev)
    ref = copy.deepcopy(model)
    for src, dst in zip(model._dict_modules(), ref._dict_modules()):
        dst.load_state_dict(src.state_dict())
    batches = []
    for _ in range(3):
        u = rng.integers(1, nu, B)
        batches.append(pkg.Interaction({"user_id": torch.from_numpy(u).to(dev),
                                        "item_id": torch.from_numpy(rng.integers(1, ni, B)).to(dev),
                                        "neg_item_id": torch.from_numpy(rng.integers(1, ni, B)).to(dev),
                                        **{a: torch.from_numpy(feats[a][u]).to(dev) for a in feats}}))
    sst = ["gender", "age"]

    def run(m, trainer, shard):
        losses, grads0 = [], None
        for k, inter in enumerate(batches):
            for fn, opt in ((m.calculate_loss, trainer.optimizer_filter), (m.calculate_dis_loss, trainer.optimizer_dis)):
                opt.zero_grad()
                loss = trainer._dp_loss(fn)(trainer._shard(inter), sst) if shard else fn(inter, sst)
                loss.backward()
                if shard:
                    trainer.dp.all_reduce_grads(opt.opt.params)
                if grads0 is None:
                    ps = opt.opt.params if shard else opt.params
                    grads0 = [None if p.grad is None else p.grad.detach().clone() for p in ps]
                (opt.opt if shard else opt).step()
                t = loss.detach().clone().double()
                if shard:
                    dist.all_reduce(t)
                losses.append(float(t))
        return losses, grads0

    dp = ops.ChainDP(rank, world, dist.group.WORLD, dev)
    model.train()
    l_dp, g_dp = run(model, pkg.PFCNTrainer(cfg, model, dp=dp), True)
    ops.set_chain_dp(None)
    ref.train()
    l_ref, g_ref = run(ref, pkg.PFCNTrainer(cfg, ref), False)
    gerr = 0.0
    for a, b in zip(g_dp, g_ref):
        if a is None or float(b.abs().max()) < 1e-7:
            continue
        gerr = max(gerr, float((a - b).abs().max() / b.abs().max()))
    berr = 0.0
    for ma, mb in zip(model._dict_modules(), ref._dict_modules()):
        for (n, x), (_, y) in zip(ma.named_buffers(), mb.named_buffers()):
            if x.dtype == torch.float32 and "running_mean" not in n:     # running means carry the arbitrary pre-BN bias
                berr = max(berr, float((x - y).abs().max() / y.abs().max().clamp_min(1e-12)))
    if rank == 0:
        out.update(dict(l_dp=l_dp, l_ref=l_ref, gerr=gerr, berr=berr, timeout=int(dp.status.item())))
    dp.close()
    dist.destroy_process_group()


def test_pfcn_data_parallel_matches_single_gpu():
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_pfcn_worker, args=(world, 29561, out), nprocs=world, join=True)
    assert out["timeout"] == 0
    # first batch (one filter step, one discriminator step): 1e-5; afterwards Adam has amplified the rounding of the
    # all-reduce's other summation order (m / sqrt(v) on near-zero gradients), as between any two float32 runs: 1e-4
    np.testing.assert_allclose(out["l_dp"][:2], out["l_ref"][:2], rtol=1e-5)
    np.testing.assert_allclose(out["l_dp"], out["l_ref"], rtol=1e-4)
    assert out["gerr"] < 2e-5, out["gerr"]
    assert out["berr"] < 1e-5, out["berr"]
