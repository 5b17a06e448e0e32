import sys; sys.path.insert(0, '.')
import numpy as np, torch
import recbole_fairrec_b200 as fr
from recbole_fairrec_b200 import synth
dev = torch.device("cuda")
uid, iid, rating, gender = synth.interactions(6041, 3707, 1_000_209)
train, valid, test = synth.split_by_user(uid, iid, rating)
cfg = fr.Config(embedding_size=64, fair_objective="value", fair_weight=1.0, topk=[10], valid_metric="NDCG@10",
                train_batch_size=2048, learning_rate=1e-3, weight_decay=1e-3, device=dev, score_mode="tc", epochs=3)
tdata = fr.TrainData(*train, gender, 6041, 3707, dev)
loader = fr.FOCFDataLoader(cfg, tdata)
model = fr.FOCF(cfg, synth.SynthDataset(6041, 3707, 5.0)).to(dev)
trainer = fr.FOCFTrainer(cfg, model)
users, hist, pos = synth.eval_lists(train, valid, test, "valid")
vdata = fr.EvalData(users, hist, pos, {"gender": gender.astype(np.int64)}, dev)
import time; t = time.time()
best_score, best_result = trainer.fit(loader, vdata, saved=False, verbose=False)
print("README example ok:", best_score, {k: best_result[k] for k in list(best_result)[:3]}, "losses", trainer.train_loss_dict, "%.2fs" % (time.time() - t))
