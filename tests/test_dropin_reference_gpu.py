"""The drop-in claim of INTEGRATION.md section 1, exercised with the UNMODIFIED reference code (oracle/_ref, a copy of its
`recbole` package made by oracle/make_ref.py, imported through oracle/ref_shim): the product's FOCF model is handed to
the reference's OWN `Config` / `create_dataset` / `data_preparation` / `FOCFDataLoader` / `Trainer._train_epoch` /
`Trainer.evaluate` (`_full_sort_batch_eval` -> `Collector` -> `Evaluator`, recbole/trainer/trainer.py:155-204, 420-515) on
the bundled ml-100k files, and must reproduce what the reference's own FOCF produced with the same seed
(tests/golden/ml100k_focf_value.npz).  Runs in a subprocess: importing the reference replaces `recbole` and patches numpy /
torch.load for the whole interpreter."""
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(ROOT, "oracle", "_ref", "recbole")

SCRIPT = textwrap.dedent('''
    import json, os, sys, tempfile
    import numpy as np
    ROOT, OUT = sys.argv[1], sys.argv[2]
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shim"))
    import shim
    shim.install()
    import torch, yaml
    from recbole.config import Config
    from recbole.data import create_dataset, data_preparation
    from recbole.trainer import Trainer
    from recbole.utils import init_seed
    import recbole_fairrec_b200 as pkg

    METRICS12 = ["NDCG", "Recall", "Hit", "MRR", "DifferentialFairness", "GiniIndex", "PopularityPercentage",
                 "ValueUnfairness", "AbsoluteUnfairness", "UnderUnfairness", "OverUnfairness", "NonParityUnfairness"]
    cfg = dict(data_path=os.path.join(ROOT, "tests", "data") + "/", RATING_FIELD="rating", LABEL_FIELD="label",
               threshold={"rating": 3.0}, fair_objective="value",
               load_col={"inter": ["user_id", "item_id", "rating"], "user": ["user_id", "gender"], "item": ["item_id"]},
               sst_attr_list=["gender"], fair_weight=1.0, neg_sampling=None, weight_decay=0.001, embedding_size=64, epochs=2,
               topk=[10], valid_metric="NDCG@10", metrics=METRICS12, seed=2020, use_gpu=True, gpu_id=0, state="WARNING",
               show_progress=False, metric_decimal_place=12,
               eval_args={"split": {"RS": [8, 1, 1]}, "group_by": "user", "order": "RO", "mode": "full"})
    os.chdir(tempfile.mkdtemp())
    with open("c.yaml", "w") as f:
        yaml.safe_dump(cfg, f)
    sys.argv = sys.argv[:1]
    config = Config(model="FOCF", dataset="ml-100k", config_file_list=["c.yaml"])
    init_seed(config["seed"], config["reproducibility"])
    dataset = create_dataset(config)
    train_data, valid_data, test_data = data_preparation(config, dataset)
    model = pkg.FOCF(config, train_data.dataset).to(config["device"])          # the PRODUCT model, the reference's ctor call
    U0 = model.user_embedding_layer.weight.detach().cpu().numpy().copy()
    trainer = Trainer(config, model)                                            # the REFERENCE trainer, stock
    trainer.eval_collector.data_collect(train_data)
    losses, valids = [], []
    for ep in range(2):
        losses.append(float(trainer._train_epoch(train_data, ep)))
        valids.append({k: float(v) for k, v in trainer.evaluate(valid_data, load_best_model=False).items()})
    test = {k: float(v) for k, v in trainer.evaluate(test_data, load_best_model=False).items()}
    json.dump({"losses": losses, "valid": valids, "test": test, "U0_sum": float(np.abs(U0).sum()),
               "launches": int(pkg._lib.launch_count()), "device": str(model.user_embedding_layer.weight.device)},
              open(OUT, "w"))
''')


@pytest.mark.skipif(not os.path.isdir(REF), reason="oracle/_ref missing: run __graft_entry__.build() where /root/reference exists")
def test_product_focf_under_the_reference_trainer_collector_and_evaluator(tmp_path):
    out = tmp_path / "res.json"
    script = tmp_path / "run.py"
    script.write_text(SCRIPT)
    env = dict(os.environ, FAIRREC_REFERENCE_ROOT=os.path.join(ROOT, "oracle", "_ref"))
    r = subprocess.run([sys.executable, str(script), ROOT, str(out)], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.load(open(out))
    g = np.load(os.path.join(HERE, "golden", "ml100k_focf_value.npz"))
    assert res["device"].startswith("cuda") and res["launches"] > 100           # the product's kernels did the work
    # same seed -> same initial tables as the reference's own FOCF drew
    np.testing.assert_allclose(res["U0_sum"], float(np.abs(g["U0"]).sum()), rtol=1e-6)
    # the reference's loader draws the reference's batches: the epoch losses of its own run within 1e-5
    np.testing.assert_allclose(res["losses"], g["epoch_losses"], rtol=1e-5)
    names = [str(k) for k in g["metric_names"]]
    assert list(res["test"].keys()) == names
    n_eval = 943
    for k, ref in zip(names, g["test_metrics"]):
        tol = 3.0 / n_eval if "@" in k else 1e-4 * max(abs(ref), 1e-3)           # cf. tests/test_run_recbole_gpu.py
        assert abs(res["test"][k] - ref) <= tol, (k, res["test"][k], ref)
    for ep in range(2):
        for k, ref in zip(names, g["valid_metrics"][ep]):
            tol = 3.0 / n_eval if "@" in k else 1e-4 * max(abs(ref), 1e-3)
            assert abs(res["valid"][ep][k] - ref) <= tol, (ep, k)
