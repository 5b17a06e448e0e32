"""FOCFTrainer -- the two hot loops of recbole/trainer/trainer.py (`_train_epoch` 155-204, `evaluate` 458-515)
behind the reference's trainer interface (`fit` 332-418 / `evaluate`), running on this package's kernels.

`get_trainer` in the reference resolves `<Model>Trainer` by name (utils/utils.py:86-94), so a class called
FOCFTrainer placed in `recbole.trainer` is picked up automatically for `model: FOCF`.

Differences from the reference loop that do not change results:
  * the loss of every step stays on the device; the epoch total is formed on the host at the end of the epoch
    as the same left-to-right Python-float sum of float32 step losses the reference builds with `.item()`
    (trainer.py:191) -- no per-batch device->host sync;
  * `learner: adam` runs the fused step (dense Adam with L2 weight decay, every row, exactly trainer.py:139);
    other learners use `calculate_loss().backward()` + the torch optimizer, like the reference;
  * evaluation is one fused pass over all eval users (evaluator.py) instead of 1-2 users per batch.
"""
import os
import time
from logging import getLogger

import numpy as np
import torch
import torch.optim as optim

from .evaluator import EvalData, FullSortEvaluator


def early_stopping(value, best, cur_step, max_step, bigger=True):
    """recbole/utils/utils.py:97-140"""
    stop_flag = update_flag = False
    if (value > best) if bigger else (value < best):
        cur_step, best, update_flag = 0, value, True
    else:
        cur_step += 1
        stop_flag = cur_step > max_step
    return best, cur_step, stop_flag, update_flag


class FOCFTrainer:
    def __init__(self, config, model, group=None):
        self.config, self.model = config, model
        self.logger = getLogger()
        self.learner = (config["learner"] or "adam").lower()
        self.learning_rate = config["learning_rate"]
        self.epochs = config["epochs"]
        self.eval_step = min(config["eval_step"] or 1, self.epochs)
        self.stopping_step = config["stopping_step"]
        self.clip_grad_norm = config["clip_grad_norm"]
        self.valid_metric = (config["valid_metric"] or "NDCG@5").lower()
        self.valid_metric_bigger = config["valid_metric_bigger"] if config["valid_metric_bigger"] is not None else True
        self.weight_decay = config["weight_decay"] or 0.0
        self.device = config["device"]
        self.checkpoint_dir = config["checkpoint_dir"] or "saved"
        self.saved_model_file = os.path.join(self.checkpoint_dir, "{}-{}.pth".format(
            config["model"] or "FOCF", time.strftime("%b-%d-%Y_%H-%M-%S")))
        self.start_epoch, self.cur_step = 0, 0
        self.best_valid_score = -np.inf if self.valid_metric_bigger else np.inf
        self.best_valid_result = None
        self.train_loss_dict = dict()
        self.group = group
        self.adam_mode = config["adam_mode"] or "dense_exact"
        self.fused = self.learner == "adam" and not self.clip_grad_norm and self.adam_mode in ("dense_exact", "lazy_exact")
        self.optimizer = None if self.fused else self._build_optimizer()
        if self.fused:
            model.init_adam(lr=self.learning_rate, weight_decay=self.weight_decay, mode=self.adam_mode)
        self.evaluator = None
        self._train_item_count = None
        self._eval_cache = {}
        self._loss_buf = None

    def _build_optimizer(self):
        """trainer.py:114-153"""
        p, lr, wd = self.model.parameters(), self.learning_rate, self.weight_decay
        if self.learner == "adam":
            return optim.Adam(p, lr=lr, weight_decay=wd)
        if self.learner == "sgd":
            return optim.SGD(p, lr=lr, weight_decay=wd)
        if self.learner == "adagrad":
            return optim.Adagrad(p, lr=lr, weight_decay=wd)
        if self.learner == "rmsprop":
            return optim.RMSprop(p, lr=lr, weight_decay=wd)
        if self.learner == "sparse_adam":
            return optim.SparseAdam(p, lr=lr)
        self.logger.warning("Received unrecognized optimizer, set default Adam optimizer")
        return optim.Adam(p, lr=lr)

    # ------------------------------------------------------------------ hot loop 1
    def _train_epoch(self, train_data, epoch_idx, loss_func=None, show_progress=False):
        """trainer.py:155-204.  Returns the epoch's summed loss (Python float)."""
        self.model.train()
        n_steps = len(train_data)
        if self._loss_buf is None or self._loss_buf.numel() < n_steps:
            self._loss_buf = torch.zeros(max(n_steps, 1), dtype=torch.float32, device=self.device)
        from .dataloader import FOCFDataLoader
        if self.group is not None and isinstance(train_data, FOCFDataLoader) and train_data.partition is not None:
            return self._train_epoch_dp(train_data)
        # the planned runner: batches of up to 8192 rows as one persistent launch per epoch (fr_focf_epoch_run) or pipelined
        # graphs of the fused step; larger batches as pipelined graphs of the multi-launch step (the preparation of batch
        # t + 1 under the Adam sweep of batch t: bench.py's `pipelined` block at configs[4]) -- bit-identical to the loop
        # below in every case (tests/test_focf_train_gpu.py, tests/test_focf_epoch_gpu.py)
        if self.fused and self.adam_mode == "dense_exact" and isinstance(train_data, FOCFDataLoader) \
                and (self.config["cuda_graph"] is None or self.config["cuda_graph"]):
            k, _ = self.model.train_epoch_planned(train_data, self._loss_buf)
            return self._finish_epoch(k)
        k = 0
        for interaction in train_data:
            if k >= self._loss_buf.numel():
                self._loss_buf = torch.cat([self._loss_buf, torch.zeros_like(self._loss_buf)])
            slot = self._loss_buf[k:k + 1]
            if self.fused:
                self.model.train_step(interaction, loss_out=slot)
            else:
                self.optimizer.zero_grad()
                loss = (loss_func or self.model.calculate_loss)(interaction.to(self.device))
                slot.copy_(loss.detach().view(1))
                loss.backward()
                if self.clip_grad_norm:
                    torch.nn.utils.clip_grad_norm_(self.model.parameters(), **self.clip_grad_norm)
                self.optimizer.step()
            k += 1
        if self.fused:
            self.model.flush_adam()     # lazy_exact: the tables are read next (evaluation, checkpoint); no-op otherwise
        return self._finish_epoch(k)

    def _train_epoch_dp(self, train_data):
        """Data-parallel epoch: every rank plans its own whole-item batches, the per-step (J, B) are summed over the
        ranks once per epoch (they are host-known at planning time), each step all-reduces the gradient shares, and
        the per-step loss shares are summed once at the end of the epoch."""
        import torch.distributed as dist
        if not self.fused:
            raise NotImplementedError("data-parallel FOCF training uses the library's Adam (learner: adam)")
        if (self.config["cuda_graph"] is None or self.config["cuda_graph"]) and self.config["dp_cuda_graph"] is not False:
            # the data-parallel step (forward -> backward -> NCCL all-reduce -> Adam) of 8 planned batches per captured
            # CUDA graph; call model.release_graphs() before destroying the process group (captured NCCL kernels keep
            # the communicator busy)
            runner = self.model.dp_planned_runner(train_data, self._loss_buf, self.group)
            k = runner.plan["len"]
            runner.run(k - runner.cursor)
            dist.all_reduce(self._loss_buf[:k], group=self.group)
            return self._finish_epoch(k)
        items, offs, batches = train_data.plan_epoch()
        dev = self.device
        sizes = torch.tensor([[b[2], b[3]] for b in batches], dtype=torch.int64, device=dev)
        dist.all_reduce(sizes, group=self.group)
        norms = sizes.cpu().numpy()
        d_items = torch.from_numpy(items).to(dev)
        d_offs = torch.from_numpy(offs).to(dev)
        uf, itf, rf, sf = train_data.train.fields
        from .interaction import Interaction
        for k, b in enumerate(batches):
            uid, iid, rating, sst = train_data.gather(d_items, d_offs, b)
            inter = Interaction({uf: uid, itf: iid, rf: rating, sf: sst})
            inter.items_contiguous = True
            self.model.dp_train_step(inter, (int(norms[k, 1]), int(norms[k, 0])), self.group,
                                     loss_out=self._loss_buf[k:k + 1])
        dist.all_reduce(self._loss_buf[:len(batches)], group=self.group)
        return self._finish_epoch(len(batches))

    def _finish_epoch(self, k):
        losses = self._loss_buf[:k].cpu().numpy()          # the epoch's single device->host sync
        self.model.check_flags()
        total = None
        for v in losses:                                   # trainer.py:191: total_loss + losses.item()
            total = float(v) if total is None else total + float(v)
        if total is not None and np.isnan(total):
            raise ValueError("Training loss is nan")        # trainer.py:286-288
        return total

    # ------------------------------------------------------------------ hot loop 2
    def _eval_data(self, eval_data):
        if isinstance(eval_data, EvalData):
            return eval_data
        key = id(eval_data)
        if key not in self._eval_cache:  # a reference FullSortEvalDataLoader
            self._eval_cache[key] = EvalData.from_reference_loader(eval_data, self.config["sst_attr_list"], self.device)
        return self._eval_cache[key]

    def data_collect(self, train_data):
        """collector.py:80-95: what the metrics need from the training data"""
        ds = train_data.dataset
        self._train_item_count = dict(ds.item_counter)
        self.evaluator = FullSortEvaluator(self.config, self.model.n_items, self._train_item_count, group=self.group)

    @torch.no_grad()
    def evaluate(self, eval_data, load_best_model=False, model_file=None, show_progress=False):
        """trainer.py:458-515"""
        if not eval_data:
            return
        if load_best_model:
            checkpoint = torch.load(model_file or self.saved_model_file, weights_only=False)
            self.model.load_state_dict(checkpoint["state_dict"])
            self.model.load_other_parameter(checkpoint.get("other_parameter"))
        self.model.eval()
        from .sampled_eval import SampledEvalData, SampledEvaluator
        if hasattr(eval_data, "resample"):          # uni<N> source that redraws its negatives per evaluation
            eval_data = eval_data.resample()
        if isinstance(eval_data, SampledEvalData):      # eval_args.mode uni<N>: sampled-negative ranking evaluation
            if getattr(self, "sampled_evaluator", None) is None:
                self.sampled_evaluator = SampledEvaluator(self.config, self.model.n_items, self._train_item_count)
            score_fn = SampledEvaluator.dot_scorer(self.model.user_embedding_layer.weight.data,
                                                   self.model.item_embedding_layer.weight.data, self.model.max_rating)
            return self.sampled_evaluator.evaluate(score_fn, eval_data)
        if self.evaluator is None:
            self.evaluator = FullSortEvaluator(self.config, self.model.n_items, self._train_item_count, group=self.group)
        data = self._eval_data(eval_data)
        return self.evaluator.evaluate(self.model.user_embedding_layer.weight.data,
                                       self.model.item_embedding_layer.weight.data, data, self.model.max_rating)

    # ------------------------------------------------------------------ bookkeeping (cold)
    def _adam_state_dict(self):
        """The fused step's Adam state in torch.optim.Adam's own state_dict layout (parameter order = the reference model's:
        user table, item table -- focf.py:43-44), so that the reference's resume_checkpoint (trainer.py:258-284) loads a
        file written here and vice versa.  lazy_exact keeps nothing extra: the trainer flushes before it saves."""
        a = self.model._adam
        step = torch.tensor(float(a["step"]))
        group = {"lr": a["lr"], "betas": (a["beta1"], a["beta2"]), "eps": a["eps"], "weight_decay": a["weight_decay"],
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                 "fused": None, "params": [0, 1]}
        return {"state": {0: {"step": step.clone(), "exp_avg": a["mU"].detach().cpu().clone(),
                              "exp_avg_sq": a["vU"].detach().cpu().clone()},
                          1: {"step": step.clone(), "exp_avg": a["mI"].detach().cpu().clone(),
                              "exp_avg_sq": a["vI"].detach().cpu().clone()}},
                "param_groups": [group], "fairrec_b200": {"adam_mode": a.get("mode", "dense_exact")}}

    def _load_adam_state_dict(self, sd):
        """Into the EXISTING moment tensors (captured graphs and persistent argument structs hold their addresses).  Accepts
        torch.optim.Adam's layout (ours and the reference's files) and the private dict earlier versions of this trainer wrote."""
        a = self.model._adam
        if "state" in sd and "param_groups" in sd:
            st = sd["state"]
            if len(st) == 0:                 # a checkpoint written before the first optimizer step
                for k in ("mU", "vU", "mI", "vI"):
                    a[k].zero_()
                a["step"] = 0
            else:
                for idx, (m, v) in ((0, ("mU", "vU")), (1, ("mI", "vI"))):
                    a[m].copy_(st[idx]["exp_avg"])
                    a[v].copy_(st[idx]["exp_avg_sq"])
                a["step"] = int(float(st[0]["step"]))
            g = sd["param_groups"][0]
            a["lr"], a["eps"], a["weight_decay"] = g["lr"], g["eps"], g["weight_decay"]
            a["beta1"], a["beta2"] = g["betas"]
        else:
            for k in ("mU", "vU", "mI", "vI"):
                a[k].copy_(sd[k])
            for k in ("step", "lr", "beta1", "beta2", "eps", "weight_decay"):
                if k in sd:
                    a[k] = sd[k]
        if a.get("mode") == "lazy_exact":    # the file holds flushed tables: every row is current for `step`
            a["last_u"].fill_(int(a["step"]))
            a["last_i"].fill_(int(a["step"]))

    def _save_checkpoint(self, epoch):
        """trainer.py:221-240"""
        os.makedirs(self.checkpoint_dir, exist_ok=True)
        if self.optimizer is None and self.model._adam.get("mode") == "lazy_exact":
            self.model.flush_adam()
        opt_state = self.optimizer.state_dict() if self.optimizer is not None else self._adam_state_dict()
        torch.save({"config": dict(self.config), "epoch": epoch, "cur_step": self.cur_step,
                    "best_valid_score": self.best_valid_score, "state_dict": self.model.state_dict(),
                    "other_parameter": self.model.other_parameter(), "optimizer": opt_state}, self.saved_model_file)

    def resume_checkpoint(self, resume_file):
        """trainer.py:258-284"""
        ck = torch.load(str(resume_file), weights_only=False)
        self.saved_model_file = str(resume_file)
        self.start_epoch, self.cur_step = ck["epoch"] + 1, ck["cur_step"]
        self.best_valid_score = ck["best_valid_score"]
        self.model.load_state_dict(ck["state_dict"])
        self.model.load_other_parameter(ck.get("other_parameter"))
        if self.optimizer is not None:
            self.optimizer.load_state_dict(ck["optimizer"])
        else:
            self._load_adam_state_dict(ck["optimizer"])

    def fit(self, train_data, valid_data=None, verbose=True, saved=True, show_progress=False, callback_fn=None):
        """trainer.py:332-418"""
        self.data_collect(train_data)
        for epoch_idx in range(self.start_epoch, self.epochs):
            t0 = time.time()
            train_loss = self._train_epoch(train_data, epoch_idx, show_progress=show_progress)
            self.train_loss_dict[epoch_idx] = train_loss
            if verbose:
                des = self.config["loss_decimal_place"] or 4
                self.logger.info(("epoch %d training [time: %.2fs, train loss: %." + str(des) + "f]") %
                                 (epoch_idx, time.time() - t0, train_loss))
            if self.eval_step <= 0 or not valid_data:
                if saved:
                    self._save_checkpoint(epoch_idx)
                continue
            if (epoch_idx + 1) % self.eval_step == 0:
                t1 = time.time()
                valid_result = self.evaluate(valid_data, load_best_model=False)
                valid_score = valid_result[self.valid_metric]
                self.best_valid_score, self.cur_step, stop_flag, update_flag = early_stopping(
                    valid_score, self.best_valid_score, self.cur_step, max_step=self.stopping_step,
                    bigger=self.valid_metric_bigger)
                if verbose:
                    self.logger.info("epoch %d evaluating [time: %.2fs, valid_score: %f]" %
                                     (epoch_idx, time.time() - t1, valid_score))
                    self.logger.info("valid result: \n" + "    ".join(f"{k} : {v}" for k, v in valid_result.items()))
                if update_flag:
                    if saved:
                        self._save_checkpoint(epoch_idx)
                    self.best_valid_result = valid_result
                if callback_fn:
                    callback_fn(epoch_idx, valid_score)
                if stop_flag:
                    if verbose:
                        self.logger.info("Finished training, best eval result in epoch %d" %
                                         (epoch_idx - self.cur_step * self.eval_step))
                    break
        return self.best_valid_score, self.best_valid_result
