/* fairrec_b200.h -- C ABI of libfairrec_b200.so (hand-written sm_100a CUDA kernels).
 *
 * This is the drop-in boundary for the hot path of RecBole-FairRec: the
 * reference is pure Python/PyTorch, so the "FFI" a maintainer binds is ctypes;
 * every entry point below names the reference code it replaces (paths relative to the reference
 * root, file:line).
 *
 * Conventions
 *  - Plain C: device pointers + sizes, no torch types.  All pointers are DEVICE pointers unless
 *    the name ends in _host.  Every function enqueues work on `stream` (a cudaStream_t passed as
 *    void*) and returns immediately; nothing here synchronises the device or allocates memory.
 *  - Return value: FR_OK or an FR_ERR_* code; fr_last_error() gives the message of the last failure
 *    on the calling thread.  Data-dependent faults the reference raises as Python exceptions (e.g.
 *    more than two sensitive-attribute values in a FOCF batch -> IndexError at focf.py:86) are
 *    recorded in a caller-owned device word `status_flags` (FR_FLAG_* bits) that the host reads at
 *    its next natural synchronisation point.
 *  - ids are int32 (the reference's int64 Interaction columns are narrowed by the host mirror),
 *    embeddings/ratings/scores float32 row-major, exactly the reference's dtypes for arithmetic.
 *  - Workspaces are caller-allocated; fr_*_workspace_bytes() tells the size.  They may be reused
 *    between calls on the same stream.
 */
#ifndef FAIRREC_B200_H_
#define FAIRREC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FR_ABI_VERSION 1

enum fr_status {
  FR_OK = 0,
  FR_ERR_INVALID = 1,   /* bad argument (null pointer, unsupported size) */
  FR_ERR_CUDA = 2,      /* a CUDA runtime call failed */
  FR_ERR_WORKSPACE = 3, /* workspace too small */
  FR_ERR_UNSUPPORTED = 4
};

/* bits of the device-side status word */
enum fr_flag {
  FR_FLAG_TOO_MANY_GROUPS = 1, /* >2 distinct sensitive-attribute values in a FOCF batch (focf.py:81-86) */
  FR_FLAG_SINGLE_GROUP = 2,    /* nonparity objective with <2 values (focf.py:129-130 IndexError) */
  FR_FLAG_NAN_LOSS = 4,        /* trainer.py:286-288 _check_nan */
  FR_FLAG_XCHG_TIMEOUT = 8     /* a peer did not arrive at a cross-GPU barrier of the row-sharded step within 20 s */
};

/* focf.py:52-68 get_loss_fun */
enum fr_objective {
  FR_OBJ_NONE = 0, FR_OBJ_VALUE = 1, FR_OBJ_ABSOLUTE = 2, FR_OBJ_UNDER = 3, FR_OBJ_OVER = 4, FR_OBJ_NONPARITY = 5
};

/* score transform applied after the dot product */
enum fr_transform {
  FR_TRANSFORM_NONE = 0,      /* raw dot (focf.py:140) */
  FR_TRANSFORM_CLAMP_DIV = 1, /* clamp(x,0,max_rating)/max_rating (focf.py:150,178) */
  FR_TRANSFORM_SIGMOID = 2    /* pfcn_pmf.py:216 / nfcf.py:73 */
};

/* arithmetic of the full-sort scoring contraction */
enum fr_score_mode {
  FR_SCORE_EXACT_FP32 = 0, /* k-ascending fmaf chain on CUDA cores: bit-identical to oracle/c */
  FR_SCORE_TC_3XTF32 = 1   /* tcgen05 tensor cores, 3xTF32 split (fp32-level accuracy, not bit-defined) */
};

int fr_abi_version(void);
const char *fr_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
uint64_t fr_launch_count(void);
/* Measurement aid (bench.py roofline): while enabled, every kernel launch of this library is bracketed by two
 * CUDA events on its stream; fr_profile_report synchronises, writes "kernel,count,total_ms" lines into buf and
 * returns the number of distinct kernels.  Off by default (zero overhead). */
void fr_profile_enable(int on);
int fr_profile_report(char *buf, size_t buf_bytes);

/* ------------------------------------------------------------------------------------------------
 * Sorting / segmentation primitives (replace torch.unique(return_inverse=True), focf.py:77-78, and the
 * implicit sort inside nn.Embedding's dense backward; also used by the metric kernels).
 * Stable LSD radix sort of (key, value) pairs; key_bits = number of significant key bits.
 * ---------------------------------------------------------------------------------------------- */
size_t fr_sort_pairs_workspace_bytes(int64_t n);
int fr_sort_pairs_u32(const uint32_t *keys_in, const uint32_t *vals_in, uint32_t *keys_out, uint32_t *vals_out,
                      int64_t n, int key_bits, void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * FOCF training step.
 *
 * One struct describes the model state and one batch; the same struct drives
 *   fr_focf_forward   -> focf.py:136-143 forward + 75-134 fairness objective + 152-169 calculate_loss
 *   fr_focf_backward  -> autograd of the above + nn.Embedding dense backward (trainer.py:193)
 *   fr_focf_adam      -> torch.optim.Adam dense step incl. L2 weight decay (trainer.py:139,196)
 *   fr_focf_train_step = forward + backward + adam fused (the dense gradient is never materialised)
 * ---------------------------------------------------------------------------------------------- */
typedef struct fr_focf_step {
  /* model (focf.py:43-44): row-major float32 tables, row 0 is the [PAD] id */
  float *U;            /* [n_users, d] */
  float *I;            /* [n_items, d] */
  int32_t n_users, n_items, d;
  /* batch (data/interaction.py Interaction columns of focf_dataloader.py:49) */
  const int32_t *uid;  /* [B] */
  const int32_t *iid;  /* [B] */
  const float *rating; /* [B] RATING_FIELD */
  const float *sst;    /* [B] sensitive attribute VALUE (token id or 0/1 float), focf.py:77 */
  int32_t B;
  int32_t items_contiguous; /* 1: rows of one item are adjacent (FOCFDataLoader batches) -> no item sort */
  /* objective */
  int32_t objective;   /* enum fr_objective */
  float fair_weight;   /* focf.py:166 */
  /* outputs */
  float *pred;         /* [B] forward scores */
  float *loss;         /* [1] calculate_loss value (written at loss[0]) */
  int32_t *status_flags; /* [1] OR-ed FR_FLAG_* */
  /* Adam state (only read by fr_focf_adam / fr_focf_train_step) */
  float *mU, *vU, *mI, *vI;
  int32_t step;        /* 1-based optimizer step t; <= 0: use the workspace's device-resident counter
                          (fr_focf_set_counters), advanced by every fr_focf_train_step -- CUDA-graph replay */
  double lr, beta1, beta2, eps, weight_decay; /* doubles: torch derives 1-beta, lr/(1-beta1^t) in double */
  /* dense gradients (only written by fr_focf_backward; may be NULL for fr_focf_train_step) */
  float *dU, *dI;
  /* scratch */
  void *workspace;
  size_t workspace_bytes;
  /* --- optional: device-resident batch description, so that ONE captured CUDA graph of the step can be replayed on
   * every batch of an epoch (batch sizes differ from step to step).  B above is then the CAPACITY of the columns. */
  const int32_t *B_dev;       /* [1] actual number of rows of a caller-built batch (NULL: B is exact) */
  /* planned epoch (focf_dataloader.py:37-50 for a whole epoch, drawn ahead): when plan_desc != NULL the step first
   * materialises its batch itself (fr_focf_gather_batch semantics) into uid/iid/rating/sst (caller-owned scratch
   * columns of capacity B) from descriptor row (cursor % plan_len) and writes its loss to loss[cursor % plan_len]
   * (loss must hold plan_len floats); the cursor
   * lives in the workspace (fr_focf_set_counters) and advances by one per forward. */
  const int32_t *plan_desc;   /* [plan_len, 4] = {first draw index, first offset index, J, B} */
  const int32_t *plan_items;  /* concatenated drawn item ids */
  const int32_t *plan_offs;   /* concatenated per-batch exclusive row offsets (J+1 each) */
  int32_t plan_len;
  const int32_t *item_off;    /* CSC of the item-sorted train split, as in fr_focf_gather_batch */
  const int32_t *train_uid;
  const float *train_rating;
  const float *sst_of_user;
  /* --- optional: data-parallel step.  Every rank holds whole items (disjoint item partitions), so the item x group
   * sums are rank-local; only the two means are global: norm_B / norm_J (> 0) replace this rank's B and J as the
   * normalisers of MSE and of the fairness mean.  The loss written is this rank's share (sum over ranks = loss);
   * gradients from fr_focf_backward are this rank's share (all-reduce, then fr_focf_adam on every replica). */
  int32_t norm_B, norm_J;
  /* planned data-parallel steps (CUDA-graph replay): [plan_len, 2] = (B_total, J_total) of every planned batch, read at
   * the device-resident cursor instead of norm_B / norm_J */
  const int32_t *norm_dev;
  /* --- optional: lazy_exact Adam (enum fr_adam_mode; 0 = dense_exact).  Needs step >= 1 (host-provided), the general
   * (non-cooperative) step path, and three caller-owned buffers: last_step_u/i [n_users] / [n_items] uint32 zeroed at the
   * start, adam_scalars [2 * scalars_cap] floats and the host counter scalars_filled (0 at the start). */
  int32_t adam_mode;
  uint32_t *last_step_u, *last_step_i;
  float *adam_scalars;
  int32_t scalars_cap;
  int32_t *scalars_filled;
  /* 1: never take the single-launch cooperative path for small batches.  Both paths are deterministic, but they add the
   * item x group sums in different orders; set it to compare dense_exact with lazy_exact (always multi-launch) bit for bit. */
  int32_t no_fused;
} fr_focf_step;

size_t fr_focf_workspace_bytes(int32_t n_users, int32_t n_items, int32_t d, int32_t max_batch);
/* zero the persistent part of a fresh workspace (row-stamp tables); call once after allocation */
int fr_focf_workspace_init(void *workspace, size_t workspace_bytes, int32_t n_users, int32_t n_items, int32_t d,
                           int32_t max_batch, void *stream);
/* set the workspace's device-resident counters: the planned-batch cursor (< 0: unchanged), the Adam step count used
 * when fr_focf_step.step <= 0 (INT32_MIN: unchanged) and how far both advance per step (< 1: unchanged; 2 when two
 * workspaces alternate the batches of one plan, see fr_focf_step_prepare / fr_focf_step_compute) */
int fr_focf_set_counters(void *workspace, size_t workspace_bytes, int32_t n_users, int32_t n_items, int32_t d,
                         int32_t max_batch, int32_t plan_cursor, int32_t adam_step, int32_t stride, void *stream);
int fr_focf_forward(const fr_focf_step *s, void *stream);
/* needs the workspace left by fr_focf_forward for the same batch; grad_scale = upstream dL (usually 1) */
int fr_focf_backward(const fr_focf_step *s, float grad_scale, void *stream);
int fr_focf_adam(const fr_focf_step *s, void *stream);
int fr_focf_train_step(const fr_focf_step *s, void *stream);
/* The training loop body of trainer.py:181-196 for batches that are still in HOST memory -- the entry a plugin calls
 * with the reference's CPU-side Interaction columns: n_steps optimisation steps, batch k = batch_rows[k] rows packed at
 * host_batches[k] (pinned memory for asynchronous copies) as int32 user ids | int32 item ids | float rating | float sst,
 * 16 bytes per row.  Per step: ONE host->device copy into `stage` (device, stage_bytes >= 16 * largest batch),
 * fr_focf_train_step on it with optimizer step tmpl->step + k, ONE device->host copy of the loss into loss_host[k]
 * (loss_dev: 2 device floats).  The host waits for step k-1's loss after enqueuing step k and returns when the last one
 * has arrived (so loss_host is complete on return).  tmpl: every field of the step except uid / iid / rating / sst / B /
 * loss; tmpl->step = 1-based optimizer step of the first batch; no planned epoch (plan_desc, B_dev NULL). */
int fr_focf_train_steps_host(const fr_focf_step *tmpl, int32_t n_steps, const void *const *host_batches,
                             const int32_t *batch_rows, void *stage, size_t stage_bytes, float *loss_dev,
                             float *loss_host, void *stream);
/* fr_focf_train_step in two halves, for overlapping the preparation of batch t+1 (its own stream and workspace) with the
 * compute of batch t: prepare = planned batch gather + sort / segments / row stamps (does not touch the embedding
 * tables); compute = forward + loss + gradients + Adam (one cooperative launch for small batches) */
int fr_focf_step_prepare(const fr_focf_step *s, void *stream);
int fr_focf_step_compute(const fr_focf_step *s, void *stream);

/* focf_dataloader.py:37-50 (FOCFDataLoader._next_batch_data) + dataset join of the user feature: build the
 * batch = all train rows of the drawn items.  The train split is device resident, sorted by item:
 * item_off [n_items+1] offsets into train_uid / train_rating; sst_of_user [n_users].  draw_items [J] are the
 * drawn item ids, draw_off [J+1] the exclusive prefix sum of their row counts (batch positions). */
int fr_focf_gather_batch(const int32_t *item_off, const int32_t *train_uid, const float *train_rating,
                         const float *sst_of_user, const int32_t *draw_items, const int32_t *draw_off, int32_t J,
                         int32_t *uid, int32_t *iid, float *rating, float *sst, void *stream);

/* focf.py:145-150 predict / pair scores with the EXACT (fmaf-chain) arithmetic of the scorer;
 * also produces rec.positive_score (collector.py:179) */
int fr_pair_scores(const float *U, const float *I, const int32_t *uid, const int32_t *iid, int64_t n, int32_t d,
                   int32_t transform, float max_rating, float *out, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Small-MLP towers (recbole/model/layers.py:30-85 MLPLayers) and the NFCF step built on them.
 * ---------------------------------------------------------------------------------------------- */
enum fr_activation { FR_ACT_NONE = 0, FR_ACT_RELU = 1, FR_ACT_LEAKYRELU = 2, FR_ACT_SIGMOID = 3, FR_ACT_TANH = 4 };

/* MLPLayers(layers=dims[0..n_layers], dropout, activation): per layer Dropout -> Linear(W[l] [dims[l+1], dims[l]],
 * b[l]) -> activation (also after the LAST layer, layers.py:66-68) */
typedef struct fr_mlp_tower {
  int32_t n_layers;
  int32_t dims[9];
  const float *W[8];
  const float *b[8];
  int32_t act;     /* enum fr_activation */
  float dropout;   /* p of the Dropout in front of every Linear (training only) */
} fr_mlp_tower;

/* NFCF (recbole/model/fair_recommender/nfcf.py): p = sigmoid(tower(U[uid] || I[iid])) (69-74),
 * loss = BCELoss(p, label) (+ fair_weight * differential fairness over the batch's positives, 76-110) */
typedef struct fr_nfcf_step {
  const float *U, *I;          /* [n_users, d], [n_items, d] */
  int32_t n_users, n_items, d;
  const int32_t *uid, *iid;    /* [M] */
  const float *label;          /* [M] LABEL_FIELD (1 / 0) */
  const float *sst;            /* [M] sensitive attribute value (use_df only) */
  int64_t M;
  fr_mlp_tower tower;          /* dims[0] == 2*d, dims[n_layers] == 1 */
  int32_t use_df;              /* nfcf.py:106-110: regulariser active (a pre-trained model was loaded) */
  float fair_weight;
  int32_t training;            /* dropout active */
  uint64_t seed;               /* dropout stream of this step */
  float *pred;                 /* [M] out: p (nfcf.py:112-115 predict) */
  float *loss;                 /* [1] out */
  int32_t *status_flags;       /* FR_FLAG_TOO_MANY_GROUPS: the kernels implement the binary-attribute case */
  /* backward outputs (fr_nfcf_backward): dense embedding grads (dU may be NULL: frozen table, nfcf.py:66) + tower */
  float *dU, *dI;
  float *dW[8];
  float *db[8];
  void *workspace;
  size_t workspace_bytes;
} fr_nfcf_step;

size_t fr_nfcf_workspace_bytes(const fr_mlp_tower *t, int64_t M);
int fr_nfcf_forward(const fr_nfcf_step *s, void *stream);
/* needs the workspace left by fr_nfcf_forward of the same batch */
int fr_nfcf_backward(const fr_nfcf_step *s, float grad_scale, void *stream);
/* torch.optim.Adam (L2 weight-decay form) on one flat parameter tensor (trainer.py:139 for the tower parameters) */
int fr_adam_dense(float *p, const float *g, float *m, float *v, int64_t n, int32_t step, double lr, double beta1,
                  double beta2, double eps, double weight_decay, void *stream);

/* torch.optim.Adam (trainer.py:139, 1189-1235) over MANY small parameter tensors in one launch per 48 entries; every entry
 * carries its own step count (torch keeps state['step'] per parameter and skips parameters without a gradient) */
typedef struct fr_adam_entry {
  float *p;
  const float *g;
  float *m, *v;
  int64_t n;
  int32_t step;      /* 1-based step count of THIS parameter after the update (used when step_dev is NULL) */
  int32_t *step_dev; /* optional device-resident counter: incremented on the stream right before the update and used
                        instead of `step`, so that a captured CUDA graph advances the bias correction on every replay */
} fr_adam_entry;
int fr_adam_multi(const fr_adam_entry *entries_host, int32_t n_entries, double lr, double beta1, double beta2, double eps,
                  double weight_decay, void *stream);

/* ---- generic layer ops: the pieces of MLPLayers (layers.py:58-70) for the PFCN / FairGo filter, discriminator and
 * scorer MLPs.  Each is one layer's forward or backward; the host mirror chains them (torch.autograd.Functions). */
/* Y = act(dropout(X) . W^T + b): nn.Dropout -> nn.Linear -> activation (layers.py:60-68 without BatchNorm) */
/* dropout masks are a counter-based hash of (seed + *seed_dev, layer, element); seed_dev (may be NULL) is a device-resident
 * offset so that replays of a captured CUDA graph draw fresh masks (bump it with fr_bump_u64 inside the graph) */
/* Shapes with K % 32 == 0 and N % 32 == 0 run on the tensor cores when allow_tensor_cores != 0 (tcgen05.mma kind::tf32
 * fed by TMA, 3xTF32 split done in shared memory, accumulator in TMEM; linear_tc.cu): one launch, no workspace; results
 * agree with the CUDA-core kernel to fp32 rounding.  Other shapes use the CUDA-core kernel. */
int fr_linear_uses_tensor_cores(int64_t M, int32_t K, int32_t N);
int fr_linear_forward(const float *X, const float *W, const float *b, float *Y, int64_t M, int32_t K, int32_t N, int32_t act,
                      float drop_p, uint64_t seed, const uint64_t *seed_dev, int32_t layer, int32_t allow_tensor_cores,
                      void *stream);
int fr_bump_u64(uint64_t *counter_dev, uint64_t inc, void *stream);
size_t fr_linear_backward_workspace_bytes(int64_t M, int32_t K, int32_t N);
/* dY is the gradient w.r.t. the layer OUTPUT Y (post-activation); dX may be NULL */
int fr_linear_backward(const float *X, const float *W, const float *Y, const float *dY, int64_t M, int32_t K, int32_t N,
                       int32_t act, float drop_p, uint64_t seed, const uint64_t *seed_dev, int32_t layer, float *dX, float *dW,
                       float *db, int32_t *tickets, void *workspace, size_t workspace_bytes, void *stream);
/* tickets (may be NULL): int32[1024] zero-initialised ONCE by the caller and then left alone (self-resetting); lets the
 * weight-gradient kernel add its chunk partials itself instead of two extra reduction launches.  One buffer per stream. */
/* nn.BatchNorm1d (layers.py:64-65) fused with the following activation; training mode uses batch statistics and
 * updates the running ones (momentum 0.1, unbiased variance), eval mode uses the running statistics */
size_t fr_batchnorm_workspace_bytes(int32_t N);
int fr_batchnorm_forward(const float *X, const float *gamma, const float *beta, float *running_mean, float *running_var,
                         int64_t M, int32_t N, float momentum, float eps, int32_t training, int32_t act, float *Y,
                         float *save_mean, float *save_invstd, void *workspace, size_t workspace_bytes, void *stream);
int fr_batchnorm_backward(const float *X, const float *Y, const float *dY, const float *gamma, const float *save_mean,
                          const float *save_invstd, int64_t M, int32_t N, int32_t act, float *dX, float *dgamma,
                          float *dbeta, void *workspace, size_t workspace_bytes, void *stream);
/* nn.Embedding lookup written into columns [col0, col0+d) of a wider row-major matrix (torch.cat for free) */
int fr_gather_rows(const float *T, const int32_t *idx, int64_t M, int32_t d, float *out, int32_t ld_out, int32_t col0,
                   void *stream);
/* dense nn.Embedding backward from columns [col0, col0+d) of the row gradients (sorted-segment sums, batch order) */
size_t fr_scatter_rows_workspace_bytes(int64_t M);
int fr_scatter_rows_dense(const int32_t *idx, const float *dX, int32_t ldx, int32_t col0, int64_t M, int32_t d,
                          int32_t n_rows, float *dT, void *workspace, size_t workspace_bytes, void *stream);
/* recbole/model/loss.py:44-46 BPRLoss; pfcn_mlp.py:206-209 BCELoss(sigmoid(z), y) and CrossEntropyLoss */
int fr_bpr_loss(const float *pos, const float *neg, int64_t M, float *loss, float *dpos, float *dneg, void *stream);
int fr_sigmoid_bce_loss(const float *z, const float *y, int64_t M, float *loss, float *dz, void *stream);
int fr_softmax_ce_loss(const float *Z, const int32_t *y, int64_t M, int32_t C, float *loss, float *dZ, void *stream);

/* ---- whole MLP chains in one launch (mlp_chain.cu): recbole/model/layers.py:30-85 MLPLayers -- [Dropout -> Linear ->
 * BatchNorm1d? -> activation?] x L -- forward as ONE persistent cooperative kernel and backward as ONE, every GEMM (forward,
 * data gradient, weight gradient) on tcgen05 (3xTF32, TMA-fed, TMEM accumulators); BatchNorm batch statistics and bias
 * gradients are float64 column sums added in tile order.  Up to FR_CHAIN_MAX_CHAINS chains over the same M rows (the
 * discriminators of one PFCN step, pfcn_mlp.py:195-211) share a launch.  Rules: every K % 4 == 0, widths <= 256,
 * <= FR_CHAIN_MAX_LAYERS layers per chain, <= 24 layers per launch (fr_mlp_chain_eligible).
 * Dropout masks are the ones of fr_linear_forward for the same (seed, seed_dev): hash of (seed, 0, row * K + k). */
#define FR_CHAIN_MAX_LAYERS 8
#define FR_CHAIN_MAX_CHAINS 4
typedef struct fr_chain_layer {
  int32_t K, N;              /* in / out features of the Linear */
  int32_t act;               /* activation code of fr_linear_forward, applied after the Linear (or its BatchNorm) */
  int32_t has_bn;
  float drop_p;              /* dropout on the layer INPUT (training only) */
  float bn_eps, bn_momentum;
  int32_t bn_repeat;         /* training forward: apply the running-statistics update this many times (0 or 1 = once).  2 = the
                              * reference's PFCN loss, which evaluates the deterministic filter twice on the same batch
                              * (pfcn_mlp.py:177-193): same output, the running statistics advance twice */
  uint64_t seed;             /* dropout seed of this layer */
  const float *W, *b;        /* [N,K], [N] or NULL */
  const float *gamma, *beta; /* BatchNorm affine */
  float *running_mean, *running_var;
  int64_t *num_batches_tracked; /* may be NULL; += 1 per training forward */
  float *dW, *db, *dgamma, *dbeta; /* backward outputs (db / dgamma / dbeta may be NULL) */
} fr_chain_layer;
typedef struct fr_chain {
  int32_t n_layers;
  fr_chain_layer layer[FR_CHAIN_MAX_LAYERS];
  const float *X;            /* [M, ldx] input (first layer's K columns used) */
  int32_t ldx;               /* 0 = K of the first layer */
  float *Y;                  /* [M, N_last] output (forward writes; backward reads it for the last activation) */
  const float *dY;           /* backward: gradient w.r.t. Y */
  float *dX;                 /* backward: [M, K_first] gradient w.r.t. X, or NULL */
  void *fwd_ws;              /* forward workspace: written by the forward call, read by the backward call */
  size_t fwd_ws_bytes;
  void *bwd_ws;              /* backward scratch */
  size_t bwd_ws_bytes;
} fr_chain;
/* bind the device's primary context to the calling thread: call once (outside any stream capture) on a thread that will
 * call fr_mlp_chain_* without having used the CUDA runtime before (e.g. torch's autograd worker thread) */
int fr_thread_init(void);
int fr_mlp_chain_eligible(const fr_chain_layer *layers, int32_t n_layers, int64_t M);
/* diagnostic: with FR_CHAIN_TRACE=1 (forward) / =2 (backward) in the environment, CTA 0 of every chain launch stamps
 * %globaltimer at its phase boundaries; this copies the first n (<= 256) stamps of the LAST launch to the host */
int fr_mlp_chain_trace(uint64_t *out_host, int32_t n);
/* backward == 0: the forward workspace for (training, need_grad); backward != 0: the backward scratch */
size_t fr_mlp_chain_workspace_bytes(const fr_chain_layer *layers, int32_t n_layers, int64_t M, int32_t training,
                                    int32_t need_grad, int32_t backward);
/* barrier_words: uint32[2] zero-initialised ONCE by the caller (self-resetting), one buffer per stream.
 * training != 0: BatchNorm uses batch statistics and advances the running ones, dropout is on; need_grad != 0 keeps what
 * fr_mlp_chain_backward needs in fwd_ws. */
int fr_mlp_chain_forward(const fr_chain *chains, int32_t n_chains, int64_t M, int32_t training, int32_t need_grad,
                         const uint64_t *seed_dev, uint32_t *barrier_words, void *stream);
/* after fr_mlp_chain_forward(training = 1, need_grad = 1) with the same chains / workspaces; dX_sum (may be NULL): the
 * fixed-order sum of the chains' dX (they then must all have a dX buffer and the same input width) */
int fr_mlp_chain_backward(const fr_chain *chains, int32_t n_chains, int64_t M, const uint64_t *seed_dev, float *dX_sum,
                          uint32_t *barrier_words, void *stream);

/* ---- row-wise scorers and glue ops of the PFCN / FairGo families (layer_ops.cu) */
/* out[m] = sum_k A[m,k] * B[m,k]: torch.mul(u, i).sum(-1) (pfcn_pmf.py:172,183-184; fairgo_pmf.py:169) */
int fr_rowdot_forward(const float *A, const float *B, int64_t M, int32_t d, float *out, void *stream);
int fr_rowdot_backward(const float *A, const float *B, const float *dout, int64_t M, int32_t d, float *dA, float *dB,
                       void *stream);
/* nn.CosineSimilarity(dim=1, eps) (pfcn_dmf.py:176,190-191); norms = float[2*M] kept for the backward */
int fr_cosine_forward(const float *A, const float *B, int64_t M, int32_t d, float eps, float *out, float *norms,
                      void *stream);
int fr_cosine_backward(const float *A, const float *B, const float *out, const float *norms, const float *dout, int64_t M,
                       int32_t d, float eps, float *dA, float *dB, void *stream);
/* pfcn_biasedmf.py:189-192: the [B] dots broadcast against [B,1] bias columns give [B,B] score matrices
 *   pos[i,j] = dotp[j] + ub[i] + pib[i] + gb, neg[i,j] = dotn[j] + ub[i] + nib[i] + gb; loss = BPRLoss over all B*B pairs.
 * Outputs the loss and the gradients of the four inputs that do not cancel; row_scratch = float[B]. */
int fr_bpr_outer_loss(const float *dotp, const float *dotn, const float *ub, const float *pib, const float *nib,
                      const float *gb, int32_t B, float *loss, float *d_dotp, float *d_dotn, float *d_pib, float *d_nib,
                      float *row_scratch, void *stream);
/* out = (x0 + x1 + ... ) * scale, or / scale when divide != 0; xs_host = HOST array of n_terms (<= 8) device pointers
 * (cm-mode filter averaging pfcn_mlp.py:158-165, fairgo_pmf.py:163-168) */
int fr_scaled_sum(const float *const *xs_host, int32_t n_terms, int64_t n, float scale, int32_t divide, float *out,
                  void *stream);
/* dst[m, col_dst : col_dst+ncols] = src[m, col_src : col_src+ncols] (torch.cat / torch.split along dim 1) */
int fr_copy_cols(const float *src, int32_t ld_src, int32_t col_src, float *dst, int32_t ld_dst, int32_t col_dst, int64_t M,
                 int32_t ncols, void *stream);
/* nn.MSELoss() (fairgo_pmf.py:170-171): loss and d loss / d pred */
int fr_mse_loss(const float *pred, const float *target, int64_t M, float *loss, float *dpred, void *stream);
/* stand-alone activation (act codes of fr_linear_forward) and its backward through the OUTPUT y */
int fr_act_forward(const float *x, int32_t act, int64_t n, float *y, void *stream);
int fr_act_backward(const float *dY, const float *Y, int32_t act, int64_t n, float *dX, void *stream);
/* stand-alone inverted dropout, y = x * mask(seed + *seed_dev, i) / (1 - p); applied to dY it is its own backward.
   Replaces the F.dropout between the convolutions of torch_geometric.nn.GCN (fairgo_gcn.py:52-57, pretrain stage). */
int fr_dropout(const float *x, float p, uint64_t seed, const uint64_t *seed_dev, int64_t n, float *y, void *stream);
/* out[m] = act(dot[m] + ub[m] + ib[m] + gb[0]) (pfcn_biasedmf.py:170-181 predict) */
int fr_biased_score(const float *dot, const float *ub, const float *ib, const float *gb, int64_t M, int32_t act, float *out,
                    void *stream);
/* clamp(x, 0, hi) / hi (fairgo_pmf.py:248 predict) */
int fr_clamp_div(const float *x, float hi, int64_t n, float *y, void *stream);

/* ---- CSR SpMM for the FairGo ego-network aggregation (spmm.cu): Y = A . X, fairgo_pmf.py:199-202.
 * The matrix is static: plan once on the host (rows cut into chunks of <= chunk nnz), copy the plan arrays to the
 * device, then call fr_spmm_csr every step (backward = the same call on the CSR of A^T). */
int fr_spmm_plan_sizes(const int64_t *row_off_host, int32_t n_rows, int32_t chunk, int64_t *n_chunks, int64_t *n_multi,
                       int64_t *n_slots, int64_t *n_empty);
int fr_spmm_plan_fill(const int64_t *row_off_host, int32_t n_rows, int32_t chunk, int32_t *chunk_row_host,
                      int32_t *chunk_begin_host, int32_t *chunk_end_host, int32_t *chunk_slot_host, int32_t *multi_row_host,
                      int32_t *multi_first_host /* n_multi + 1 */, int32_t *empty_row_host);
typedef struct fr_spmm_plan {
  const int32_t *chunk_row, *chunk_begin, *chunk_end, *chunk_slot; /* device, [n_chunks] */
  const int32_t *multi_row, *multi_first;                          /* device, [n_multi], [n_multi + 1] */
  const int32_t *empty_row;                                        /* device, [n_empty] */
  int64_t n_chunks, n_multi, n_slots, n_empty;
} fr_spmm_plan;
/* partial = float[n_slots * d] scratch (may be NULL when n_slots == 0) */
int fr_spmm_csr(const fr_spmm_plan *plan, const int32_t *col, const float *val, const float *X, int32_t d, float *Y,
                float *partial, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Full-sort evaluation: scoring contraction fused with the pad/history mask and a streaming top-K.
 * Replaces focf.py:171-178 full_sort_predict + trainer.py:435-438 mask + collector.py:143-153 topk /
 * pos-matrix / gather.  Scores are never materialised.
 *
 * Item sharding: the call scores items [item_base, item_base + n_items_local) held in I_shard; ids in
 * hist/pos CSR and in the outputs are GLOBAL item ids.  Total order: (score desc, item id asc).
 * ---------------------------------------------------------------------------------------------- */
typedef struct fr_fullsort {
  const float *U;          /* [n_users_total, d] */
  const float *I_shard;    /* [n_items_local, d] rows item_base.. */
  int32_t d, n_items_local, item_base;
  const int32_t *users;    /* [n] eval user ids (general_dataloader.py:188 uid_list) */
  int32_t n;
  const int64_t *hist_off; /* [n+1] CSR of history items per eval user, item ids ASCENDING per user */
  const int32_t *hist_items;
  int32_t K;               /* max(topk), 1..64 */
  int32_t transform;       /* enum fr_transform */
  float max_rating;
  int32_t score_mode;      /* enum fr_score_mode */
  int32_t *topk_id;        /* [n, K] out */
  float *topk_score;       /* [n, K] out */
  void *workspace;
  size_t workspace_bytes;
} fr_fullsort;

size_t fr_fullsort_workspace_bytes(int32_t n, int32_t K, int32_t n_items_local, int32_t d, int32_t score_mode);
int fr_fullsort_topk(const fr_fullsort *a, void *stream);

/* merge P per-shard top-K lists [P, n, K] (as gathered by NCCL all-gather) into the global top-K */
int fr_topk_merge(const int32_t *ids_in, const float *scores_in, int32_t P, int32_t n, int32_t K,
                  int32_t *ids_out, float *scores_out, void *stream);

/* collector.py:147-153: hit bits of the top-K against the positives CSR (ascending item ids per user)
 * -> rec.topk int32 [n, K+1] = [hit bits | pos_len] */
int fr_hits(const int32_t *topk_id, int32_t n, int32_t K, const int64_t *pos_off, const int32_t *pos_items,
            int32_t *rec_topk, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Metric accumulation (evaluator/metrics.py, the 12 metrics of properties/model/FOCF.yaml:29-30).
 * ---------------------------------------------------------------------------------------------- */
/* metrics.py:63-65,89-97,160-161,187-203 + base_metric.py:59-82: per-user Hit/MRR/Recall/NDCG@1..K summed
 * over users in float64.  sums_out: [4, K] double (rows: ndcg, recall, hit, mrr) = SUM over users. */
size_t fr_topk_metrics_workspace_bytes(int32_t n, int32_t K);
int fr_topk_metrics(const int32_t *rec_topk, int32_t n, int32_t K, double *sums_out, void *workspace,
                    size_t workspace_bytes, void *stream);

/* metrics.py:644-661 (GiniIndex) and 772-820 (PopularityPercentage):
 * item_pos_count_out [K, n_items] int32: how often each item was recommended AT rank e (row e)
 * pop_hits_out [K] int64: number of users whose rank-e recommendation is a popular item
 * (PopularityPercentage@k = sum_{e<k} pop_hits[e] / (n * k); is_popular may be NULL) */
int fr_rec_item_stats(const int32_t *topk_id, int32_t n, int32_t K, int32_t n_items, const uint8_t *is_popular,
                      int32_t *item_pos_count_out, int64_t *pop_hits_out, void *stream);
/* Gini@k from the first k rows of item_pos_count (sorts the summed counts; integer arithmetic until the
 * final division): gini_out[0] double */
size_t fr_gini_workspace_bytes(int32_t n_items);
int fr_gini_at_k(const int32_t *item_pos_count, int32_t n_items, int32_t k_rows, int64_t n_users, double *gini_out,
                 void *workspace, size_t workspace_bytes, void *stream);

/* Sampled-negative ("uni100") ranking evaluation -- trainer.py:441-456 (_neg_sample_batch_eval: candidate scores
 * scattered into a [users, n_items] matrix of -inf) + collector.py:141-153 (topk, hit bits, number of positives) without
 * the dense matrix.  Candidates of user u = cand_items[cand_off[u] .. cand_off[u+1]) with their scores; the first
 * n_pos_of_user[u] of them are the positives (general_dataloader.py:128-152 layout).  Canonical order (score desc, item
 * id asc); duplicate candidates collapse; fewer than K distinct candidates -> lowest non-candidate ids as -inf filler.
 * rec_topk [n, K+1] = hit bits | number of distinct positives. */
int fr_sampled_topk(const int64_t *cand_off, const int32_t *cand_items, const float *cand_scores,
                    const int32_t *n_pos_of_user, int32_t n, int32_t K, int32_t n_items, int32_t *topk_id,
                    float *topk_score, int32_t *rec_topk, void *stream);

/* metrics.py:860-881, 935-1266, 1313-1341: per (positive item, group) sums over the eval positives.
 * group[p] in [0,G).  Outputs (float64): stats [n_items, G, 2] = (sum score, count) per item x group.
 * Deterministic (sorted-segment reduction, no float atomics). */
size_t fr_item_group_stats_workspace_bytes(int64_t n_pos, int32_t n_items, int32_t G);
int fr_item_group_stats(const int32_t *pos_items, const float *pos_score, const int32_t *group, int64_t n_pos,
                        int32_t n_items, int32_t G, double *stats_out, void *workspace, size_t workspace_bytes,
                        void *stream);
/* The same in two parts: the item-sorted view of the positive list (order, segments) depends on the evaluation data only,
 * not on the scores, so it is built once (fr_item_group_plan into a caller-owned plan buffer of fr_item_group_plan_bytes)
 * and every evaluation pass runs the segment reduction alone (fr_item_group_stats_planned; same sums, same order). */
size_t fr_item_group_plan_bytes(int64_t n_pos);
size_t fr_item_group_plan_workspace_bytes(int64_t n_pos);
int fr_item_group_plan(const int32_t *pos_items, int64_t n_pos, int32_t n_items, void *plan, size_t plan_bytes,
                       void *workspace, size_t workspace_bytes, void *stream);
int fr_item_group_stats_planned(const void *plan, size_t plan_bytes, const float *pos_score, const int32_t *group,
                                int64_t n_pos, int32_t n_items, int32_t G, double *stats_out, void *stream);
/* reduce stats -> fairness metrics.  out[0..6] double: DifferentialFairness (metrics.py:1313-1341), Value,
 * Absolute, Under, Over (935-1266; NaN unless G == 2), NonParity (860-881), n_distinct_positive_items */
int fr_fairness_metrics(const double *stats, int32_t n_items, int32_t G, double *out, void *workspace,
                        size_t workspace_bytes, void *stream);
size_t fr_fairness_metrics_workspace_bytes(int32_t n_items, int32_t G);
/* Value / Absolute / Under / Over unfairness in the sampled-negative (`uni<N>`) evaluation mode, metrics.py:935-978 and
 * siblings with mode != 'full' (collector.py:190-205: each positive is paired with its FIRST sampled negative, scored for the
 * positive's user): stats_all / stats_pos [n_items, 2, 2] = fr_item_group_stats over positives + paired negatives / over the
 * positives only (binary attribute).  out[0..3] = the four metrics, out[4] = number of items in the union. */
size_t fr_unfairness_sampled_workspace_bytes(int32_t n_items);
int fr_unfairness_sampled(const double *stats_all, const double *stats_pos, int32_t n_items, double *out, void *workspace,
                          size_t workspace_bytes, void *stream);


/* ----------------------------------------------------------------------------------------------
 * Row-sharded FOCF training step over NVLink peer memory (SURVEY.md 8e "FOCF training", rows 1-2; the
 * data-parallel form of trainer.py:181-196 for tables that do not belong on one GPU).
 *
 * World of P <= FR_MAX_RANKS processes, one GPU each.  Rank r OWNS rows r, r+P, r+2P, ... of the user table, the
 * item table and their Adam moments (local row = global row / P) and holds the train interactions of ITS users,
 * in CSC form by GLOBAL item id.  A batch is the reference's: all train rows of the drawn items
 * (focf_dataloader.py:37-50); the draw list is the same on every rank, each rank materialises the rows of its own
 * users.  Per step, user-side work (gather, gradient segments, Adam) is entirely local; what crosses NVLink is
 *   (1) the current item rows of the J drawn items, pushed by their owners into every rank's staging table
 *       [J, d] (the forward / backward kernels then index items by draw position),
 *   (2) each rank's partial item x group sums [J, 8 floats] (all ranks reduce them in rank order -> identical
 *       fairness terms and loss everywhere),
 *   (3) each rank's partial item gradients [J, d], pushed to the item's owner, which adds them in rank order and
 *       applies Adam to its rows.
 * Nothing else is exchanged: J is ~10^3 rows when the batch is 10^6 interactions.  All transfers are plain stores
 * into peer memory issued by the kernels themselves, separated by cross-GPU barriers (one flag word per peer,
 * release/acquire at system scope; a barrier that does not complete within 20 s raises FR_FLAG_XCHG_TIMEOUT
 * instead of hanging).  Every reduction has a fixed order: results are bit-stable run to run and equal to the
 * single-GPU step on the same batch up to summation order.
 *
 * Exchange memory: ONE device allocation per rank (fr_xchg_alloc) laid out by fr_focf_shard_xchg_bytes, identical
 * on all ranks; peers map it through CUDA IPC (fr_xchg_export / fr_xchg_open; the handles travel over whatever
 * channel the host has -- torch.distributed here).  For single-process tests `xchg[k]` may simply be the
 * allocations of P emulated ranks on one device; then the caller sequences the phases itself (no barrier:
 * `barriers = 0`).
 * ---------------------------------------------------------------------------------------------- */
#define FR_MAX_RANKS 8

int fr_xchg_alloc(size_t bytes, void **ptr_out);                 /* cudaMalloc + zero */
int fr_xchg_free(void *ptr);
int fr_xchg_export(void *ptr, void *handle64_out);               /* 64-byte CUDA IPC handle of a fr_xchg_alloc buffer */
int fr_xchg_open(const void *handle64, void **peer_ptr_out);     /* map a peer's buffer into this process */
int fr_xchg_close(void *peer_ptr);

/* ---- data-parallel MLP chains: every rank runs fr_mlp_chain_*_dp on ITS rows [rank * M, (rank + 1) * M) of a global batch
 * of world * M rows (recbole/trainer/trainer.py:865-898 sharded over the GPUs of one box).  BatchNorm keeps single-GPU
 * batch semantics (recbole/model/layers.py:62-70): the per-row-block column sums (sum z, sum z^2; backward: sum g, sum g*xhat)
 * are written by the GEMM epilogues straight into EVERY rank's exchange memory (NVLink peer stores), the launch is cut
 * where they cross the ranks, a flag barrier (st.release.sys / ld.acquire.sys on peer memory) sits between the
 * segments, and every rank then adds the world * ceil(M/128) partials in global row-block order: the statistics are
 * bit-identical on all ranks and to the single-GPU run when M % 128 == 0.  Dropout masks use the global row index.
 * Gradient outputs are this rank's share: the host sums them over the ranks (NCCL all-reduce).
 * Exchange memory: one fr_xchg_alloc buffer per rank, mapped by the peers (fr_xchg_export / fr_xchg_open);
 * 256 + 2 * sum over BatchNorm layers of world * ceil(M/128) * N * 16 bytes. */
typedef struct fr_chain_dp {
  int32_t rank, world;
  void *xchg[FR_MAX_RANKS];
  size_t xchg_bytes;
  int32_t barriers;      /* 1: flag barriers between the segments (one process per GPU).  0: single-process emulation -- the
                            caller runs segment after segment, rank after rank, and passes `segment` and `parity` */
  int32_t segment;       /* barriers == 0: the segment to run (-1: all, for chains without BatchNorm) */
  int32_t parity;        /* barriers == 0: exchange region written by this segment (reads use the other one) */
  int32_t *status_flags; /* device word: FR_FLAG_XCHG_TIMEOUT when a peer never arrived (may be NULL) */
} fr_chain_dp;
int fr_mlp_chain_segments(const fr_chain *chains, int32_t n_chains, int32_t training, int32_t backward, int32_t world);
int fr_mlp_chain_forward_dp(const fr_chain *chains, int32_t n_chains, int64_t M, int32_t training, int32_t need_grad,
                            const uint64_t *seed_dev, uint32_t *barrier_words, const fr_chain_dp *dp, void *stream);
int fr_mlp_chain_backward_dp(const fr_chain *chains, int32_t n_chains, int64_t M, const uint64_t *seed_dev, float *dX_sum,
                             uint32_t *barrier_words, const fr_chain_dp *dp, void *stream);

enum fr_shard_phase {
  FR_SHARD_STAGE = 1, /* push the rows of stage_items that this rank owns into every rank's staging table (+ barrier) */
  FR_SHARD_A = 2,     /* batch gather, sort / segments, forward, partial item x group sums pushed (+ barrier) */
  FR_SHARD_B = 4,     /* reduce the sums -> loss + fairness terms; gradient segments; partial item gradients pushed (+ barrier) */
  FR_SHARD_C = 8,     /* Adam: this rank's user rows, the item rows it owns */
  FR_SHARD_FLUSH = 16 /* lazy_exact only: bring every local row up to optimizer step `step` (before the tables are read) */
};

enum fr_adam_mode {
  FR_ADAM_DENSE_EXACT = 0, /* every row moves every step (torch.optim.Adam with weight_decay, trainer.py:139) */
  FR_ADAM_LAZY_EXACT = 1   /* a row is brought up to date when a batch next touches it (or at FR_SHARD_FLUSH), by replaying the
                              missed steps (gradient = weight_decay * p) with the same float32 operations: bit-identical to
                              DENSE_EXACT, without the per-step sweep over all rows */
};

typedef struct fr_focf_shard_step {
  /* this rank's rows of the tables and moments */
  float *U, *I, *mU, *vU, *mI, *vI;
  int32_t n_users_loc, n_items_loc, n_items, d; /* n_items: GLOBAL item count (CSC length) */
  int32_t rank, world;
  /* this rank's interactions, CSC by global item id; user ids are LOCAL rows */
  const int32_t *item_off;     /* [n_items + 1] */
  const int32_t *train_uid;
  const float *train_rating;
  const float *sst_of_user;    /* [n_users_loc] */
  /* current batch */
  const int32_t *draw_items;   /* [J] global ids of the drawn items, draw order (the same list on every rank) */
  const int32_t *draw_off;     /* [J+1] exclusive prefix of THIS RANK's row counts of the drawn items */
  const int32_t *draw_slot;    /* [J] position of item j among the batch items of its owner (owner = id % world) */
  int32_t J, B_loc, B_glob;    /* drawn items; this rank's rows (= draw_off[J], may be 0); rows over all ranks */
  int32_t parity;              /* step & 1: which half of the double-buffered exchange memory this step uses */
  /* items whose rows FR_SHARD_STAGE pushes (the NEXT batch's draw list when combined with A|B|C; into parity ^ 1 then) */
  const int32_t *stage_items;
  int32_t stage_J, stage_parity;
  /* objective (nonparity needs batch-global means: not available sharded) */
  int32_t objective;
  float fair_weight;
  /* Adam */
  int32_t adam_mode;           /* enum fr_adam_mode */
  int32_t step;                /* 1-based optimizer step of this batch */
  double lr, beta1, beta2, eps, weight_decay;
  uint32_t *last_step_u, *last_step_i; /* lazy_exact: [n_users_loc], [n_items_loc] step each row is current for (zeroed at start) */
  float *adam_scalars;         /* lazy_exact: [2 * scalars_cap] per-step (-lr / (1 - beta1^t), sqrt(1 - beta2^t)), filled on demand */
  int32_t scalars_cap;
  int32_t *scalars_filled;     /* lazy_exact: HOST int, number of steps already tabulated (updated by the call) */
  /* columns of the materialised local batch (capacity >= B_loc) + outputs */
  int32_t *uid, *iid;          /* iid holds DRAW POSITIONS j, not item ids */
  float *rating, *sst, *pred;
  float *loss;                 /* [1] the batch loss (identical on every rank) */
  int32_t *status_flags;
  void *workspace;             /* fr_focf_shard_workspace_bytes, zeroed once by fr_focf_shard_workspace_init */
  size_t workspace_bytes;
  /* exchange memory: xchg[k] = rank k's allocation as mapped in THIS process (xchg[rank] = own) */
  void *xchg[FR_MAX_RANKS];
  int32_t J_cap;               /* capacity the exchange layout was sized for (>= every batch's J) */
  int32_t barriers;            /* 1: cross-GPU barriers after STAGE / A / B (real multi-process run); 0: caller sequences phases */
  int32_t prebuilt;            /* 1: uid / iid (draw positions) / rating / sst already hold this rank's rows of the batch (a host
                                  batch copied in by the caller, interaction.py:174-200 `.to(device)`): phase A skips its gather */
} fr_focf_shard_step;

size_t fr_focf_shard_xchg_bytes(int32_t world, int32_t J_cap, int32_t d);
size_t fr_focf_shard_workspace_bytes(int32_t n_users_loc, int32_t n_items_loc, int32_t d, int32_t max_batch, int32_t J_cap);
int fr_focf_shard_workspace_init(void *workspace, size_t workspace_bytes, int32_t n_users_loc, int32_t n_items_loc, int32_t d,
                                 int32_t max_batch, int32_t J_cap, void *stream);
/* run the phases in `phases` (OR of enum fr_shard_phase) in the order A, B, C, STAGE, FLUSH */
int fr_focf_shard_step_run(const fr_focf_shard_step *s, int32_t phases, void *stream);

/* ------------------------------------------------------------------------------------------------
 * A planned FOCF epoch (or any run of its steps) as ONE persistent cooperative launch (csrc/focf_epoch.cu): the body of
 * Trainer._train_epoch (trainer.py:181-196) over the batches of FOCFDataLoader._next_batch_data
 * (focf_dataloader.py:37-50).  For the latency-bound regime: batch capacity <= 8192 rows, d <= 128, dense_exact Adam,
 * tables small enough for every compute CTA to keep its share of [U; I] and of both moments in shared memory
 * (fr_focf_epoch_eligible).  Results (tables, moments, losses) are bit-identical to fr_focf_train_step over the same batches.
 *
 * slots[0..n_slots): n_slots (2..8; 4 is a good value) PLANNED fr_focf_step structs that differ only in their batch
 * columns (uid / iid / rating / sst / pred, capacity B each) and workspace: producer CTAs fill slot (k mod n_slots) with
 * batch first_batch + k of the plan while earlier steps compute.  The steps run are the plan rows
 * (first_batch + k) mod plan_len, k = 0..n_steps-1; losses go to loss[(first_batch + k) mod plan_len]; adam_step is the
 * 1-based optimizer step of the first of them.  The workspaces' device-resident cursor / Adam counters are neither read
 * nor advanced.  sync_words: 64 bytes of device memory owned by the caller (zeroed by the call).
 * ---------------------------------------------------------------------------------------------- */
int fr_focf_epoch_eligible(const fr_focf_step *slots, int32_t n_slots);
int fr_focf_epoch_run(const fr_focf_step *slots, int32_t n_slots, int32_t first_batch, int32_t n_steps,
                      int32_t adam_step, void *sync_words, void *stream);
/* diagnostic (FR_FOCF_TRACE=1; FR_FOCF_TRACE_STEP selects the step, default 8): [CTA][16] %globaltimer stamps of the last
 * epoch launch -- compute CTAs: step start | forward done | barrier 1 passed | statistics done | gradients done | barrier 2
 * passed | Adam done (thread 0) | barrier 3 passed; producer CTAs: start | gathered | sorted | published.
 * (%globaltimer reads of ~140 CTAs in the same microsecond serialise: per-CTA phase durations are reliable, the cross-CTA
 * alignment of simultaneous stamps is not.) */
int fr_focf_epoch_trace(uint64_t *out_host, int32_t n);

/* diagnostic: with FR_FOCF_TRACE=1 in the environment every CTA of the cooperative fused step stamps %globaltimer at its 8
 * phase boundaries (start | forward done | barrier 1 passed | statistics done | gradients done | barrier 2 passed | Adam done |
 * barrier 3 passed); this copies the first n (<= 2048) stamps ([CTA][8]) of the LAST launch to the host */
int fr_focf_step_trace(uint64_t *out_host, int32_t n);

/* lazy_exact for the single-GPU step: fr_focf_train_step with fr_focf_step.adam_mode = FR_ADAM_LAZY_EXACT updates only the
 * rows the batch touches; fr_focf_adam_flush brings every row of both tables up to step `step` (call before anything
 * reads the tables: evaluation, checkpoint, predict). */
int fr_focf_adam_flush(const fr_focf_step *s, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FAIRREC_B200_H_ */
