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
    seed=2020, score_mode="auto", adam_mode="dense_exact", neg_sampling={"uniform": 1},
    TIME_FIELD="timestamp", filter_inter_by_user_or_item=True, user_inter_num_interval="[0,inf)",
    item_inter_num_interval="[0,inf)",
)


# Hyper-parameter defaults of the reference's per-model property files (recbole/properties/model/<Model>.yaml), the layer
# configurator.py:211-257 puts between overall.yaml and the user's files.  Model / training keys only: the data keys of
# those files (load_col, threshold, ...) are dead in the reference too (sample.yaml overrides them), and
# eval_args stays with DEFAULTS (full-sort) because the sampled mode refuses four of the default metrics (sampled_eval.py).
_DIS = [128, 256, 128, 128, 64, 32]
MODEL_DEFAULTS = {
    "FOCF": dict(embedding_size=64, fair_objective="none", fair_weight=1.0, neg_sampling=None, weight_decay=0.001),
    "NFCF": dict(embedding_size=64, mlp_hidden_size=[128, 64], dropout=0.2, fair_weight=0.1, weight_decay=1e-6),
    "PFCN_MLP": dict(embedding_size=64, mlp_hidden_size_list=[64, 32, 16], dis_hidden_size_list=_DIS, activation="leakyrelu",
                     filter_mode="sm", dropout=0.2, dis_dropout=0.3, train_epoch_interval=5, weight_decay=0.0001,
                     dis_weight=10.0),
    "PFCN_PMF": dict(embedding_size=64, dis_hidden_size_list=_DIS, activation="leakyrelu", filter_mode="none",
                     dis_dropout=0.3, train_epoch_interval=5, weight_decay=0.0001, dis_weight=10),
    "PFCN_BiasedMF": dict(embedding_size=64, dis_hidden_size_list=_DIS, activation="leakyrelu", filter_mode="none",
                          dis_dropout=0.3, train_epoch_interval=5, weight_decay=0.0001, dis_weight=10),
    "PFCN_DMF": dict(embedding_size=64, num_layers=3, dis_hidden_size_list=_DIS, mlp_activation="relu",
                     dis_activation="leakyrelu", filter_mode="sm", mlp_dropout=0.2, dis_dropout=0.3, train_epoch_interval=5,
                     weight_decay=0.001, dis_weight=10),
    "FairGo_PMF": dict(load_pretrain_weight=False, aggr_method="LBA", vs_weights=[4, 1], embedding_size=64,
                       dis_hidden_size_list=[16, 8, 4], filter_hidden_size_list=[128, 64], activation="leakyrelu",
                       fair_weight=0.1, pretrain_epochs=600, train_epoch_interval=5, weight_decay=0.0001),
    "FairGo_GCN": dict(aggr_method="LBA", vs_weights=[4, 1], embedding_size=64, n_layers=2, dis_hidden_size_list=[16, 8, 4],
                       filter_hidden_size_list=[128, 64], activation="leakyrelu", fair_weight=0.1, gcn_n_layers=2,
                       hidden_channels=32, gcn_dropout=0.2, gcn_act="relu", pretrain_epochs=600, train_epoch_interval=5,
                       weight_decay=0.0001),
}


class Config(dict):
    def __init__(self, **kw):
        super().__init__(DEFAULTS)
        self.update(kw)
        if "device" not in self:
            self["device"] = torch.device("cuda" if torch.cuda.is_available() else "cpu")

    def __getitem__(self, k):
        return self.get(k, None)
