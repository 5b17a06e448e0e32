import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import recbole_fairrec_b200 as pkg
from test_focf_epoch_gpu import _setup
from test_focf_train_gpu import make_model
cfg, train, U0, I0 = _setup(3, 400, 150, 9000, 1024, 32, "value")
loader = pkg.FOCFDataLoader(cfg, train, mode="fast", seed=11)
model = make_model(U0, I0, "value", 0.7)
model.init_adam(lr=1e-3, weight_decay=1e-3)
n = len(loader)
losses = torch.zeros(2 * n, device="cuda")
runner = model.epoch_runner(loader, losses[:n], n_slots=int(sys.argv[1]) if len(sys.argv) > 1 else 6)
runner.run(n)
torch.cuda.synchronize()
model.check_flags()
print("epoch of", n, "steps ok; loss[0..2]", losses[:3].tolist())
