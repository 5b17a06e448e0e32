"""ctypes binding of libfairrec_b200.so (the C ABI declared in include/fairrec_b200.h).

There is NO CPU fallback: if the shared library is missing the import of the product package fails with
a build hint, and every wrapper refuses tensors that are not on a CUDA device.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfairrec_b200.so")

FR_OK, FR_ERR_INVALID, FR_ERR_CUDA, FR_ERR_WORKSPACE, FR_ERR_UNSUPPORTED = 0, 1, 2, 3, 4      # include/fairrec_b200.h: enum fr_status
FLAG_TOO_MANY_GROUPS, FLAG_SINGLE_GROUP, FLAG_NAN_LOSS, FLAG_XCHG_TIMEOUT = 1, 2, 4, 8
ADAM_DENSE_EXACT, ADAM_LAZY_EXACT = 0, 1                                   # enum fr_adam_mode
SHARD_STAGE, SHARD_A, SHARD_B, SHARD_C, SHARD_FLUSH = 1, 2, 4, 8, 16      # enum fr_shard_phase
MAX_RANKS = 8
OBJECTIVES = {"none": 0, "value": 1, "absolute": 2, "under": 3, "over": 4, "nonparity": 5}
TRANSFORM_NONE, TRANSFORM_CLAMP_DIV, TRANSFORM_SIGMOID = 0, 1, 2
SCORE_EXACT_FP32, SCORE_TC_3XTF32 = 0, 1


class FairRecLibraryError(RuntimeError):
    pass


class FocfStep(Structure):
    """mirror of `struct fr_focf_step` (include/fairrec_b200.h)"""
    _fields_ = [
        ("U", c_void_p), ("I", c_void_p), ("n_users", c_int32), ("n_items", c_int32), ("d", c_int32),
        ("uid", c_void_p), ("iid", c_void_p), ("rating", c_void_p), ("sst", c_void_p), ("B", c_int32),
        ("items_contiguous", c_int32), ("objective", c_int32), ("fair_weight", c_float),
        ("pred", c_void_p), ("loss", c_void_p), ("status_flags", c_void_p),
        ("mU", c_void_p), ("vU", c_void_p), ("mI", c_void_p), ("vI", c_void_p),
        ("step", c_int32), ("lr", c_double), ("beta1", c_double), ("beta2", c_double), ("eps", c_double),
        ("weight_decay", c_double), ("dU", c_void_p), ("dI", c_void_p),
        ("workspace", c_void_p), ("workspace_bytes", c_size_t),
        ("B_dev", c_void_p), ("plan_desc", c_void_p), ("plan_items", c_void_p), ("plan_offs", c_void_p),
        ("plan_len", c_int32), ("item_off", c_void_p), ("train_uid", c_void_p), ("train_rating", c_void_p),
        ("sst_of_user", c_void_p), ("norm_B", c_int32), ("norm_J", c_int32), ("norm_dev", c_void_p),
        ("adam_mode", c_int32), ("last_step_u", c_void_p), ("last_step_i", c_void_p), ("adam_scalars", c_void_p),
        ("scalars_cap", c_int32), ("scalars_filled", POINTER(c_int32)), ("no_fused", c_int32),
    ]


class FocfShardStep(Structure):
    """mirror of `struct fr_focf_shard_step` (include/fairrec_b200.h)"""
    _fields_ = [
        ("U", c_void_p), ("I", c_void_p), ("mU", c_void_p), ("vU", c_void_p), ("mI", c_void_p), ("vI", c_void_p),
        ("n_users_loc", c_int32), ("n_items_loc", c_int32), ("n_items", c_int32), ("d", c_int32),
        ("rank", c_int32), ("world", c_int32),
        ("item_off", c_void_p), ("train_uid", c_void_p), ("train_rating", c_void_p), ("sst_of_user", c_void_p),
        ("draw_items", c_void_p), ("draw_off", c_void_p), ("draw_slot", c_void_p),
        ("J", c_int32), ("B_loc", c_int32), ("B_glob", c_int32), ("parity", c_int32),
        ("stage_items", c_void_p), ("stage_J", c_int32), ("stage_parity", c_int32),
        ("objective", c_int32), ("fair_weight", c_float),
        ("adam_mode", c_int32), ("step", c_int32),
        ("lr", c_double), ("beta1", c_double), ("beta2", c_double), ("eps", c_double), ("weight_decay", c_double),
        ("last_step_u", c_void_p), ("last_step_i", c_void_p), ("adam_scalars", c_void_p), ("scalars_cap", c_int32),
        ("scalars_filled", POINTER(c_int32)),
        ("uid", c_void_p), ("iid", c_void_p), ("rating", c_void_p), ("sst", c_void_p), ("pred", c_void_p),
        ("loss", c_void_p), ("status_flags", c_void_p), ("workspace", c_void_p), ("workspace_bytes", c_size_t),
        ("xchg", c_void_p * 8), ("J_cap", c_int32), ("barriers", c_int32), ("prebuilt", c_int32),
    ]


class MlpTower(Structure):
    """mirror of `struct fr_mlp_tower`"""
    _fields_ = [("n_layers", c_int32), ("dims", c_int32 * 9), ("W", c_void_p * 8), ("b", c_void_p * 8),
                ("act", c_int32), ("dropout", c_float)]


class NfcfStep(Structure):
    """mirror of `struct fr_nfcf_step`"""
    _fields_ = [
        ("U", c_void_p), ("I", c_void_p), ("n_users", c_int32), ("n_items", c_int32), ("d", c_int32),
        ("uid", c_void_p), ("iid", c_void_p), ("label", c_void_p), ("sst", c_void_p), ("M", c_int64),
        ("tower", MlpTower), ("use_df", c_int32), ("fair_weight", c_float), ("training", c_int32), ("seed", c_uint64),
        ("pred", c_void_p), ("loss", c_void_p), ("status_flags", c_void_p), ("dU", c_void_p), ("dI", c_void_p),
        ("dW", c_void_p * 8), ("db", c_void_p * 8), ("workspace", c_void_p), ("workspace_bytes", c_size_t),
    ]


class FullSort(Structure):
    """mirror of `struct fr_fullsort`"""
    _fields_ = [
        ("U", c_void_p), ("I_shard", c_void_p), ("d", c_int32), ("n_items_local", c_int32), ("item_base", c_int32),
        ("users", c_void_p), ("n", c_int32), ("hist_off", c_void_p), ("hist_items", c_void_p), ("K", c_int32),
        ("transform", c_int32), ("max_rating", c_float), ("score_mode", c_int32),
        ("topk_id", c_void_p), ("topk_score", c_void_p), ("workspace", c_void_p), ("workspace_bytes", c_size_t),
    ]


class AdamEntry(Structure):
    """mirror of `struct fr_adam_entry`"""
    _fields_ = [("p", c_void_p), ("g", c_void_p), ("m", c_void_p), ("v", c_void_p), ("n", c_int64), ("step", c_int32),
                ("step_dev", c_void_p)]


class SpmmPlan(Structure):
    """mirror of `struct fr_spmm_plan`"""
    _fields_ = [("chunk_row", c_void_p), ("chunk_begin", c_void_p), ("chunk_end", c_void_p), ("chunk_slot", c_void_p),
                ("multi_row", c_void_p), ("multi_first", c_void_p), ("empty_row", c_void_p),
                ("n_chunks", c_int64), ("n_multi", c_int64), ("n_slots", c_int64), ("n_empty", c_int64)]


class ChainLayer(Structure):
    """mirror of `struct fr_chain_layer`"""
    _fields_ = [("K", c_int32), ("N", c_int32), ("act", c_int32), ("has_bn", c_int32), ("drop_p", c_float),
                ("bn_eps", c_float), ("bn_momentum", c_float), ("bn_repeat", c_int32), ("seed", c_uint64), ("W", c_void_p), ("b", c_void_p),
                ("gamma", c_void_p), ("beta", c_void_p), ("running_mean", c_void_p), ("running_var", c_void_p),
                ("num_batches_tracked", c_void_p), ("dW", c_void_p), ("db", c_void_p), ("dgamma", c_void_p),
                ("dbeta", c_void_p)]


class Chain(Structure):
    """mirror of `struct fr_chain`"""
    _fields_ = [("n_layers", c_int32), ("layer", ChainLayer * 8), ("X", c_void_p), ("ldx", c_int32), ("Y", c_void_p),
                ("dY", c_void_p), ("dX", c_void_p), ("fwd_ws", c_void_p), ("fwd_ws_bytes", c_size_t),
                ("bwd_ws", c_void_p), ("bwd_ws_bytes", c_size_t)]


class ChainDp(Structure):
    """mirror of `struct fr_chain_dp`"""
    _fields_ = [("rank", c_int32), ("world", c_int32), ("xchg", c_void_p * 8), ("xchg_bytes", c_size_t),
                ("barriers", c_int32), ("segment", c_int32), ("parity", c_int32), ("status_flags", c_void_p)]


# name -> (restype, argtypes); also the list the ABI-surface test checks against the header
SIGNATURES = {
    "fr_abi_version": (c_int, []),
    "fr_last_error": (c_char_p, []),
    "fr_launch_count": (c_uint64, []),
    "fr_profile_enable": (None, [c_int]),
    "fr_profile_report": (c_int, [c_char_p, c_size_t]),
    "fr_sort_pairs_workspace_bytes": (c_size_t, [c_int64]),
    "fr_sort_pairs_u32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_size_t, c_void_p]),
    "fr_focf_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32, c_int32]),
    "fr_focf_workspace_init": (c_int, [c_void_p, c_size_t, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "fr_focf_set_counters": (c_int, [c_void_p, c_size_t, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                     c_void_p]),
    "fr_focf_step_prepare": (c_int, [POINTER(FocfStep), c_void_p]),
    "fr_focf_step_compute": (c_int, [POINTER(FocfStep), c_void_p]),
    "fr_focf_forward": (c_int, [POINTER(FocfStep), c_void_p]),
    "fr_focf_backward": (c_int, [POINTER(FocfStep), c_float, c_void_p]),
    "fr_focf_adam": (c_int, [POINTER(FocfStep), c_void_p]),
    "fr_focf_train_step": (c_int, [POINTER(FocfStep), c_void_p]),
    "fr_focf_train_steps_host": (c_int, [POINTER(FocfStep), c_int32, POINTER(c_void_p), POINTER(c_int32), c_void_p, c_size_t,
                                 c_void_p, c_void_p, c_void_p]),
    "fr_focf_adam_flush": (c_int, [POINTER(FocfStep), c_void_p]),
    "fr_xchg_alloc": (c_int, [c_size_t, POINTER(c_void_p)]),
    "fr_xchg_free": (c_int, [c_void_p]),
    "fr_xchg_export": (c_int, [c_void_p, c_void_p]),
    "fr_xchg_open": (c_int, [c_void_p, POINTER(c_void_p)]),
    "fr_xchg_close": (c_int, [c_void_p]),
    "fr_focf_shard_xchg_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "fr_focf_shard_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32, c_int32, c_int32]),
    "fr_focf_shard_workspace_init": (c_int, [c_void_p, c_size_t, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "fr_focf_shard_step_run": (c_int, [POINTER(FocfShardStep), c_int32, c_void_p]),
    "fr_focf_gather_batch": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p]),
    "fr_pair_scores": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_float, c_void_p,
                               c_void_p]),
    "fr_nfcf_workspace_bytes": (c_size_t, [POINTER(MlpTower), c_int64]),
    "fr_nfcf_forward": (c_int, [POINTER(NfcfStep), c_void_p]),
    "fr_nfcf_backward": (c_int, [POINTER(NfcfStep), c_float, c_void_p]),
    "fr_adam_dense": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_double, c_double, c_double,
                              c_double, c_double, c_void_p]),
    "fr_linear_uses_tensor_cores": (c_int, [c_int64, c_int32, c_int32]),
    "fr_linear_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_float,
                                  c_uint64, c_void_p, c_int32, c_int32, c_void_p]),
    "fr_bump_u64": (c_int, [c_void_p, c_uint64, c_void_p]),
    "fr_linear_backward_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32]),
    "fr_linear_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_float,
                                   c_uint64, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                   c_void_p]),
    "fr_batchnorm_workspace_bytes": (c_size_t, [c_int32]),
    "fr_batchnorm_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_float, c_float,
                                     c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "fr_batchnorm_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32,
                                      c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "fr_gather_rows": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_int32, c_int32, c_void_p]),
    "fr_scatter_rows_workspace_bytes": (c_size_t, [c_int64]),
    "fr_scatter_rows_dense": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int64, c_int32, c_int32, c_void_p, c_void_p,
                                      c_size_t, c_void_p]),
    "fr_bpr_loss": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fr_sigmoid_bce_loss": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "fr_softmax_ce_loss": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p]),
    "fr_adam_multi": (c_int, [POINTER(AdamEntry), c_int32, c_double, c_double, c_double, c_double, c_double, c_void_p]),
    "fr_biased_score": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    "fr_rowdot_forward": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    "fr_rowdot_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p]),
    "fr_cosine_forward": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_float, c_void_p, c_void_p, c_void_p]),
    "fr_cosine_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_float, c_void_p,
                                   c_void_p, c_void_p]),
    "fr_bpr_outer_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fr_scaled_sum": (c_int, [POINTER(c_void_p), c_int32, c_int64, c_float, c_int32, c_void_p, c_void_p]),
    "fr_copy_cols": (c_int, [c_void_p, c_int32, c_int32, c_void_p, c_int32, c_int32, c_int64, c_int32, c_void_p]),
    "fr_mse_loss": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "fr_act_forward": (c_int, [c_void_p, c_int32, c_int64, c_void_p, c_void_p]),
    "fr_act_backward": (c_int, [c_void_p, c_void_p, c_int32, c_int64, c_void_p, c_void_p]),
    "fr_dropout": (c_int, [c_void_p, c_float, c_uint64, c_void_p, c_int64, c_void_p, c_void_p]),
    "fr_clamp_div": (c_int, [c_void_p, c_float, c_int64, c_void_p, c_void_p]),
    "fr_spmm_plan_sizes": (c_int, [c_void_p, c_int32, c_int32, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64),
                                   POINTER(c_int64)]),
    "fr_spmm_plan_fill": (c_int, [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p]),
    "fr_spmm_csr": (c_int, [POINTER(SpmmPlan), c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p]),
    "fr_fullsort_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32, c_int32, c_int32]),
    "fr_fullsort_topk": (c_int, [POINTER(FullSort), c_void_p]),
    "fr_sampled_topk": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                c_void_p, c_void_p]),
    "fr_topk_merge": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "fr_hits": (c_int, [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fr_topk_metrics_workspace_bytes": (c_size_t, [c_int32, c_int32]),
    "fr_topk_metrics": (c_int, [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_size_t, c_void_p]),
    "fr_rec_item_stats": (c_int, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fr_gini_workspace_bytes": (c_size_t, [c_int32]),
    "fr_gini_at_k": (c_int, [c_void_p, c_int32, c_int32, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "fr_item_group_stats_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32]),
    "fr_item_group_plan_bytes": (c_size_t, [c_int64]),
    "fr_item_group_plan_workspace_bytes": (c_size_t, [c_int64]),
    "fr_item_group_plan": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p]),
    "fr_item_group_stats_planned": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p,
                                            c_void_p]),
    "fr_item_group_stats": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p,
                                    c_size_t, c_void_p]),
    "fr_fairness_metrics_workspace_bytes": (c_size_t, [c_int32, c_int32]),
    "fr_unfairness_sampled_workspace_bytes": (c_size_t, [c_int32]),
    "fr_unfairness_sampled": (c_int, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_size_t, c_void_p]),
    "fr_thread_init": (c_int, []),
    "fr_mlp_chain_trace": (c_int, [c_void_p, c_int32]),
    "fr_focf_step_trace": (c_int, [c_void_p, c_int32]),
    "fr_focf_epoch_eligible": (c_int, [c_void_p, c_int32]),
    "fr_focf_epoch_run": (c_int, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "fr_focf_epoch_trace": (c_int, [c_void_p, c_int32]),
    "fr_mlp_chain_eligible": (c_int, [POINTER(ChainLayer), c_int32, c_int64]),
    "fr_mlp_chain_workspace_bytes": (c_size_t, [POINTER(ChainLayer), c_int32, c_int64, c_int32, c_int32, c_int32]),
    "fr_mlp_chain_forward": (c_int, [POINTER(Chain), c_int32, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "fr_mlp_chain_backward": (c_int, [POINTER(Chain), c_int32, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fr_mlp_chain_segments": (c_int, [POINTER(Chain), c_int32, c_int32, c_int32, c_int32]),
    "fr_mlp_chain_forward_dp": (c_int, [POINTER(Chain), c_int32, c_int64, c_int32, c_int32, c_void_p, c_void_p,
                                        POINTER(ChainDp), c_void_p]),
    "fr_mlp_chain_backward_dp": (c_int, [POINTER(Chain), c_int32, c_int64, c_void_p, c_void_p, c_void_p, POINTER(ChainDp),
                                         c_void_p]),
    "fr_fairness_metrics": (c_int, [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_size_t, c_void_p]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises FairRecLibraryError with a build hint when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FairRecLibraryError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -c 'import __graft_entry__ as g; "
            f"g.build()'` (or `make -C recbole-fairrec_b200/csrc`). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = library/header mismatch
        fn.restype, fn.argtypes = res, args
    if lib.fr_abi_version() != 1:
        raise FairRecLibraryError(f"ABI version mismatch: library {lib.fr_abi_version()}, binding 1")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != FR_OK:
        msg = load().fr_last_error().decode("utf-8", "replace")
        raise FairRecLibraryError(f"{what} failed with status {rc}: {msg}")


def ptr(t):
    """device pointer of a CUDA tensor (None -> NULL)"""
    if t is None:
        return None
    if not t.is_cuda:
        raise FairRecLibraryError("fairrec_b200 kernels take CUDA tensors only (no CPU fallback)")
    if not t.is_contiguous():
        raise FairRecLibraryError("fairrec_b200 kernels take contiguous tensors")
    return t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def launch_count():
    return int(load().fr_launch_count())


def profile_enable(on):
    load().fr_profile_enable(1 if on else 0)


def profile_report():
    """{kernel: (launches, total_ms)} of the launches recorded since profile_enable(True); synchronises"""
    buf = ctypes.create_string_buffer(1 << 16)
    load().fr_profile_report(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.rsplit(",", 2)
        out[name] = (int(cnt), float(ms))
    return out
