import os, sys, numpy as np, torch
sys.path.insert(0, '.')
import torch.distributed as dist
import recbole_fairrec_b200 as pkg
from recbole_fairrec_b200 import synth
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29544", RANK="0", WORLD_SIZE="1")
torch.cuda.set_device(0); dev = torch.device("cuda", 0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
nu, ni, d = 700, 300, 32
uid, iid, rating, gender = synth.interactions(nu, ni, 40000, 5, item_sigma=1.0)
tr, va, te = synth.split_by_user(uid, iid, rating, seed=5)
cfg = pkg.Config(embedding_size=d, fair_objective="value", fair_weight=1.0, train_batch_size=1024, device=dev, learning_rate=1e-3, weight_decay=1e-3, epochs=1)
train = pkg.TrainData(tr[0], tr[1], tr[2], gender, nu, ni, dev)
rng = np.random.default_rng(0)
U0 = (rng.standard_normal((nu, d)) * 0.2).astype(np.float32); I0 = (rng.standard_normal((ni, d)) * 0.2).astype(np.float32)
def fresh():
    m = pkg.FOCF(cfg, synth.SynthDataset(nu, ni, 5.0))
    with torch.no_grad():
        m.user_embedding_layer.weight.copy_(torch.from_numpy(U0)); m.item_embedding_layer.weight.copy_(torch.from_numpy(I0))
    return m.to(dev)
res = {}
for mode in ("graph", "eager", "single"):
    model = fresh()
    c = pkg.Config(**{**dict(cfg), "cuda_graph": mode != "eager"})
    trainer = pkg.FOCFTrainer(c, model, group=dist.group.WORLD if mode != "single" else None)
    loader = pkg.FOCFDataLoader(c, train, mode="fast", seed=100, partition=(0, 1) if mode != "single" else None)
    print(mode, "start", flush=True)
    loss = trainer._train_epoch(loader, 0)
    res[mode] = (loss, model.user_embedding_layer.weight.detach().cpu().numpy().copy())
    print(mode, loss, flush=True)
for m in ("eager", "single"):
    print(m, abs(res["graph"][0] - res[m][0]) / abs(res[m][0]), np.abs(res["graph"][1] - res[m][1]).max() / np.abs(res[m][1]).max())
dist.destroy_process_group()
