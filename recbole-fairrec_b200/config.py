"""Tiny stand-in for the reference's Config (recbole/config/configurator.py:27-437) carrying only the
keys the hot path reads.  `__getitem__` answers None for missing keys like configurator.py:405-409.  The
reference's real Config object is accepted everywhere this one is."""
import torch

DEFAULTS = dict(
    USER_ID_FIELD="user_id", ITEM_ID_FIELD="item_id", NEG_PREFIX="neg_", RATING_FIELD="rating", LABEL_FIELD="label",
    embedding_size=64, sst_attr_list=["gender"], fair_objective="none", fair_weight=1.0,
    learner="adam", learning_rate=0.001, weight_decay=0.001, epochs=300, train_batch_size=2048,
    eval_batch_size=4096, eval_step=1, stopping_step=10, clip_grad_norm=None, loss_decimal_place=4,
    metric_decimal_place=4, topk=[5], valid_metric="NDCG@5", valid_metric_bigger=True, popularity_ratio=0.1,
    metrics=["NDCG", "Recall", "Hit", "MRR", "DifferentialFairness", "GiniIndex", "PopularityPercentage",
             "ValueUnfairness", "AbsoluteUnfairness", "UnderUnfairness", "OverUnfairness", "NonParityUnfairness"],
    eval_args={"split": {"RS": [8, 1, 1]}, "group_by": "user", "order": "RO", "mode": "full"},
    seed=2020, score_mode="exact", adam_mode="dense_exact", neg_sampling={"uniform": 1},
    TIME_FIELD="timestamp", filter_inter_by_user_or_item=True, user_inter_num_interval="[0,inf)",
    item_inter_num_interval="[0,inf)",
)


class Config(dict):
    def __init__(self, **kw):
        super().__init__(DEFAULTS)
        self.update(kw)
        if "device" not in self:
            self["device"] = torch.device("cuda" if torch.cuda.is_available() else "cpu")

    def __getitem__(self, k):
        return self.get(k, None)
