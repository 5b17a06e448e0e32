// Row-sharded FOCF training step over NVLink peer memory (include/fairrec_b200.h: fr_focf_shard_step).
//
// Reference being replaced: the body of Trainer._train_epoch (recbole/trainer/trainer.py:181-196) -- calculate_loss
// (recbole/model/fair_recommender/focf.py:136-169), backward, optimizer.step -- for embedding tables whose rows are
// distributed over P GPUs (rank r owns rows r, r+P, ...), with the batch of focf_dataloader.py:37-50 split by USER owner.
//
// Why this partition: a FOCF batch is "every train row of ~J drawn items" (J ~ 10^3 for 10^6 rows), so the item side of a
// step is tiny and the user side is everything.  Splitting the rows by user owner keeps the user gather, the user
// gradient segments and the user Adam update local; what has to cross NVLink is O(J * d) per step:
//   STAGE : owners push the current rows of the drawn items into every rank's staging table xI[par][J, d]
//   A     : each rank pushes its partial item x group sums (8 floats per drawn item) into slot `rank` of every rank's xS
//   B     : each rank pushes its partial item gradient rows into slot `rank` of the OWNER's xG
// Every transfer is a plain 128-bit store into peer memory issued by the producing kernel; consumers read local memory
// after a cross-GPU barrier (k_xbar).  Reductions over ranks run in rank order on every consumer: bit-stable, identical on
// all ranks.  The exchange memory is double buffered by step parity, so one barrier per phase is enough.
//
// HBM-bound work, no tensor cores: rows move as float4 lane accesses, warp per drawn item for the exchange kernels.
#include <stdlib.h>
#include <string.h>

#include "focf_device.cuh"

namespace fr {

// ------------------------------------------------------------------------------------------ exchange memory layout
struct XchgLayout {
  size_t flags, epoch, hdr, xI, xS, xG, total;   // byte offsets
  size_t xI_par, xS_par, xS_slot, xG_par, xG_slot, hdr_par;
};

static inline XchgLayout xchg_layout(int world, int J_cap, int d) {
  XchgLayout L;
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t off = 0;
  L.flags = off; off += up(sizeof(unsigned long long) * FR_MAX_RANKS);
  L.epoch = off; off += up(sizeof(unsigned long long));
  L.hdr_par = up(sizeof(uint32_t) * 4 * FR_MAX_RANKS);
  L.hdr = off; off += 2 * L.hdr_par;
  L.xI_par = up(sizeof(float) * (size_t)J_cap * d);
  L.xI = off; off += 2 * L.xI_par;
  L.xS_slot = up(sizeof(float) * (size_t)J_cap * 8);
  L.xS_par = L.xS_slot * world;
  L.xS = off; off += 2 * L.xS_par;
  L.xG_slot = up(sizeof(float) * (size_t)J_cap * d);
  L.xG_par = L.xG_slot * world;
  L.xG = off; off += 2 * L.xG_par;
  L.total = off;
  return L;
}

struct Peers {
  char *base[FR_MAX_RANKS];
};

// ------------------------------------------------------------------------------------------ cross-GPU barrier
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// One CTA, thread k talks to rank k: publish "I have arrived at barrier #e" in slot `rank` of k's flag array, then wait
// until k has published e in mine.  Everything this rank stored into peer memory in earlier kernels of the stream is
// ordered before the flag (kernel boundary + fence + release); the data is in the consumer's own memory when it sees it.
__global__ void k_xbar(Peers px, size_t flags_off, size_t epoch_off, int rank, int world, int32_t *status) {
  __shared__ unsigned long long e;
  unsigned long long *epoch = (unsigned long long *)(px.base[rank] + epoch_off);
  if (threadIdx.x == 0) e = *epoch + 1ull;
  __syncthreads();
  const int k = threadIdx.x;
  if (k < world) {
    __threadfence_system();
    st_release_sys((unsigned long long *)(px.base[k] + flags_off) + rank, e);
    const unsigned long long *mine = (const unsigned long long *)(px.base[rank] + flags_off) + k;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(mine) < e) {
      if (global_ns() - t0 > 20000000000ull) {   // 20 s: a peer died; do not hang the GPU
        atomicOr(status, FR_FLAG_XCHG_TIMEOUT);
        break;
      }
    }
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x == 0) *epoch = e;
}

// ------------------------------------------------------------------------------------------ workspace
struct ShardWs {
  FocfWs f;             // the single-GPU step's scratch, item side used per local segment
  int32_t *seg_of_j;    // [J_cap] local item segment of draw position j, -1 if this rank has no row of it
  uint2 *row_tab_i;     // [n_items_loc] {stamp, owner slot} of the batch that last touched the owned item row (dense mode)
  float *cseg_j;        // [J_cap, 2] fairness terms by draw position
  float *red_part;      // [kRedMaxBlocks, 2] per-CTA partial sums of k_shard_stats_reduce
};
constexpr int kRedMaxBlocks = 128;

static ShardWs carve_shard(Carver &c, int n_users_loc, int n_items_loc, int d, int B, int J_cap) {
  ShardWs w;
  // persistent parts first (their offsets must not depend on the batch size): the owned-item stamps, then the single-GPU
  // layout, which itself starts with the user stamps and the control block
  w.row_tab_i = c.take<uint2>(n_items_loc < 1 ? 1 : n_items_loc);
  w.f = carve(c, n_users_loc, 1, d, B);   // item stamps of the single-GPU layout are not used here
  w.seg_of_j = c.take<int32_t>(J_cap);
  w.cseg_j = c.take<float>(2 * (size_t)J_cap);
  w.red_part = c.take<float>(2 * kRedMaxBlocks);
  return w;
}

// ------------------------------------------------------------------------------------------ kernels
// focf_dataloader.py:37-50 for this rank's users: rows of the drawn items, item column = DRAW POSITION (the staging
// table is indexed by it)
__global__ void __launch_bounds__(256)
    k_shard_gather(const int32_t *__restrict__ item_off, const int32_t *__restrict__ train_uid,
                   const float *__restrict__ train_rating, const float *__restrict__ sst_of_user,
                   const int32_t *__restrict__ draw_items, const int32_t *__restrict__ draw_off, int J,
                   int32_t *__restrict__ uid, int32_t *__restrict__ jid, float *__restrict__ rating,
                   float *__restrict__ sst) {
  const int B = draw_off[J];
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < B; p += gridDim.x * blockDim.x) {
    int lo = 0, hi = J;  // invariant: draw_off[lo] <= p < draw_off[hi]; the last such lo is the (non-empty) item of row p
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (draw_off[mid] <= p) lo = mid; else hi = mid;
    }
    const int src = item_off[draw_items[lo]] + (p - draw_off[lo]);
    const int u = train_uid[src];
    uid[p] = u;
    jid[p] = lo;
    rating[p] = train_rating[src];
    sst[p] = sst_of_user[u];
  }
}

// draw position -> local segment; owned items -> {stamp, slot} for the dense sweep
__global__ void __launch_bounds__(256)
    k_shard_segmap(const int32_t *__restrict__ jid, const int32_t *__restrict__ segoff_i, const int32_t *__restrict__ n_seg,
                   int has_rows, const int32_t *__restrict__ draw_items, const int32_t *__restrict__ draw_slot, int J,
                   int rank, int world, const uint32_t *__restrict__ ctrl, int32_t *__restrict__ seg_of_j,
                   uint2 *__restrict__ row_tab_i) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < J) {
    const int it = draw_items[t];
    if (it % world == rank) row_tab_i[it / world] = make_uint2(ctrl[CTRL_STAMP], (uint32_t)draw_slot[t]);
  }
  if (has_rows && t < *n_seg) seg_of_j[jid[segoff_i[t]]] = t;
}

__global__ void __launch_bounds__(256)
    k_shard_forward(const float *__restrict__ U, const float *__restrict__ xI, const int32_t *__restrict__ uid,
                    const int32_t *__restrict__ jid, const float *__restrict__ sst, int B, int d, float *__restrict__ pred,
                    uint32_t *__restrict__ ctrl) {
  forward_body(U, xI, uid, jid, sst, B, d, pred, ctrl);
}

__global__ void __launch_bounds__(kLossThreads) k_shard_loss_records(LossArgs a) { loss_phase1(a, a.B); }

// partial item x group sums of every draw position (zeros where this rank has no row) -> slot `rank` of every rank's xS;
// header {local min, local max of the attribute (order-encoded), rows} -> slot `rank` of every rank's header block.
// CTA b owns the draw positions b, b + G, ...: cooperative record sum (popular items span thousands of records).
__global__ void __launch_bounds__(kSegThreads)
    k_shard_stats_push(LossArgs a, const int32_t *__restrict__ seg_of_j, int J, int B_loc, Peers px, size_t xS_off,
                       size_t hdr_off, int rank, int world) {
  __shared__ float sh[kSegThreads / 32][8];
  if (blockIdx.x == 0 && (int)threadIdx.x < world) {
    uint4 h = make_uint4(B_loc > 0 ? a.ctrl[CTRL_MIN] : 0xffffffffu, B_loc > 0 ? a.ctrl[CTRL_MAX] : 0u, (uint32_t)B_loc, 0u);
    *((uint4 *)(px.base[threadIdx.x] + hdr_off) + rank) = h;
  }
  for (int j = blockIdx.x; j < J; j += gridDim.x) {
    const int s = B_loc > 0 ? seg_of_j[j] : -1;      // uniform over the CTA
    float v[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (s >= 0) segment_record_sum(a, s, v, sh);
    if ((int)threadIdx.x < world) {   // xS_off already selects the parity half and slot `rank`
      float *dst = (float *)(px.base[threadIdx.x] + xS_off) + (size_t)j * 8;
      *(float4 *)dst = make_float4(v[0], v[1], v[2], v[3]);
      *(float4 *)(dst + 4) = make_float4(v[4], v[5], v[6], 0.f);
    }
  }
}

// every rank: add the P partial records of each draw position in rank order (groups re-based on the GLOBAL minimum of
// the attribute), fairness terms per draw position and per local segment; per-CTA partial sums of the squared errors and
// the smooth-L1 terms, which the LAST CTA to finish (ticket) adds in CTA order -> the batch loss + control-block hand-over
constexpr int kRedThreads = 256;
__global__ void __launch_bounds__(kRedThreads)
    k_shard_stats_reduce(const char *__restrict__ xS, size_t xS_slot, const uint4 *__restrict__ hdr, int world, int J,
                         int B_loc, int B_glob, int objective, float fair_weight, const int32_t *__restrict__ seg_of_j,
                         float *__restrict__ cseg_j, float *__restrict__ cseg, float *__restrict__ cglob,
                         float *__restrict__ part /* [gridDim.x, 2] */, float *__restrict__ loss,
                         uint32_t *__restrict__ ctrl, int32_t *__restrict__ flags) {
  __shared__ float sh[33];
  __shared__ uint32_t s_min, s_max;
  __shared__ int s_swap[FR_MAX_RANKS];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) {
    uint32_t lo = 0xffffffffu, hi = 0u;
    for (int k = 0; k < world; ++k) {
      if (hdr[k].z == 0u) continue;   // a rank without rows in this batch
      lo = min(lo, hdr[k].x);
      hi = max(hi, hdr[k].y);
    }
    bool bad = false;
    for (int k = 0; k < world; ++k) {
      s_swap[k] = 0;
      if (hdr[k].z == 0u) continue;
      // rank k put the rows whose value equals ITS minimum into column 0: if that is not the global minimum, its column 0
      // is the global column 1 (and its column 1 must be empty)
      s_swap[k] = hdr[k].x != lo;
      bad |= (hdr[k].x != lo && hdr[k].x != hi) || (hdr[k].y != lo && hdr[k].y != hi);
    }
    if (bad && objective != FR_OBJ_NONE && blockIdx.x == 0) atomicOr(flags, FR_FLAG_TOO_MANY_GROUPS);   // focf.py:81-86
    s_min = lo;
    s_max = hi;
  }
  __syncthreads();
  const float Bn = (float)B_glob, Jn = (float)J;
  float w_sq = 0.f, w_hx = 0.f;
  for (int j = blockIdx.x * nw + wib; j < J; j += gridDim.x * nw) {
    float v[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x;
    if (lane < world) {
      const float *src = (const float *)(xS + (size_t)lane * xS_slot) + (size_t)j * 8;
      x = *(const float4 *)src;
      y = *(const float4 *)(src + 4);
      if (s_swap[lane]) {   // this rank's column 0 holds the global group 1
        x = make_float4(x.y, x.x, x.w, x.z);
        y = make_float4(y.y, y.x, y.z, 0.f);
      }
    }
    for (int k = 0; k < world; ++k) {   // rank order
      v[0] += __shfl_sync(0xffffffffu, x.x, k); v[1] += __shfl_sync(0xffffffffu, x.y, k);
      v[2] += __shfl_sync(0xffffffffu, x.z, k); v[3] += __shfl_sync(0xffffffffu, x.w, k);
      v[4] += __shfl_sync(0xffffffffu, y.x, k); v[5] += __shfl_sync(0xffffffffu, y.y, k);
      v[6] += __shfl_sync(0xffffffffu, y.z, k);
    }
    if (lane == 0) {
      float hx = 0.f, cs0 = 0.f, cs1 = 0.f;
      w_sq += v[6];
      if (objective >= FR_OBJ_VALUE && objective <= FR_OBJ_OVER)
        segment_terms(objective, fair_weight, Jn, v[0], v[1], v[2], v[3], v[4], v[5], hx, cs0, cs1);
      w_hx += hx;
      cseg_j[2 * j] = cs0;
      cseg_j[2 * j + 1] = cs1;
      const int s = B_loc > 0 ? seg_of_j[j] : -1;
      if (s >= 0) {
        cseg[2 * s] = cs0;
        cseg[2 * s + 1] = cs1;
      }
    }
  }
  const float sq = block_sum_1024(w_sq, sh), hx = block_sum_1024(w_hx, sh);
  if (threadIdx.x == 0) {
    part[2 * blockIdx.x] = sq;
    part[2 * blockIdx.x + 1] = hx;
    __threadfence();
    is_last = atomicAdd(&ctrl[CTRL_TICKET], 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last || threadIdx.x >= 32) return;
  __threadfence();
  float tsq = 0.f, thx = 0.f;
  for (unsigned b = lane; b < gridDim.x; b += 32) {   // fixed order whichever CTA ends up last: lane-strided, then the shuffle tree
    const float2 q = __ldcg((const float2 *)(part + 2 * b));
    tsq += q.x;
    thx += q.y;
  }
  tsq = warp_sum(tsq);
  thx = warp_sum(thx);
  if (lane != 0) return;
  float l = tsq / Bn;
  if (objective >= FR_OBJ_VALUE && objective <= FR_OBJ_OVER) l += fair_weight * (thx / Jn);
  loss[0] = l;
  if (l != l) atomicOr(flags, FR_FLAG_NAN_LOSS);
  cglob[0] = 0.f;
  cglob[1] = 0.f;
  ctrl[CTRL_SAVED_MIN] = s_min;   // the backward groups rows by the GLOBAL minimum
  ctrl[CTRL_SAVED_MAX] = s_max;
  ctrl[CTRL_MIN] = 0xffffffffu;
  ctrl[CTRL_MAX] = 0u;
  ctrl[CTRL_TICKET] = 0u;
  ctrl[CTRL_STAMP] += 1u;
}

template <int kRowVecs>
__global__ void __launch_bounds__(256, kRowVecs == 1 ? 4 : 2) k_shard_grads(GradArgs a, int nchunk) {
  grads_chunk<kRowVecs>(a, nchunk, (blockIdx.x * blockDim.x + threadIdx.x) >> 5);
}

// this rank's partial gradient row of every draw position (zero where it has no row) -> slot `rank` of the OWNER's xG,
// at the item's owner slot.  One CTA per draw position: a popular item's rows span thousands of gradient chunks, whose
// partials the 8 warps add in a fixed two-level order (warp w takes chunks c0+1+w, c0+1+w+8, ... in increasing order,
// then tail + the 8 warp sums in warp order) -- deterministic, and 8 dependent-load chains instead of one.
__global__ void __launch_bounds__(256)
    k_shard_igrad_push(const int32_t *__restrict__ seg_of_j, const int32_t *__restrict__ segoff_i,
                       const float *__restrict__ gseg, const float *__restrict__ head, const float *__restrict__ tail,
                       int chunk, int d, const int32_t *__restrict__ draw_items, const int32_t *__restrict__ draw_slot,
                       int J, int B_loc, Peers px, size_t xG_off /* incl. parity and slot `rank` */, int world) {
  __shared__ __align__(16) float part[8][kMaxD];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int j = blockIdx.x;
  if (j >= J) return;
  const int s = B_loc > 0 ? seg_of_j[j] : -1;
  float *dst = (float *)(px.base[draw_items[j] % world] + xG_off) + (size_t)draw_slot[j] * d;
  int c0 = 0, c1 = 0;
  if (s >= 0) {
    const int s0 = segoff_i[s], s1 = segoff_i[s + 1];
    c0 = s0 / chunk;
    c1 = (s1 - 1) / chunk;
  }
  const bool multi = s >= 0 && c1 > c0;
  if (multi) {
    for (int k = lane * 4; k < d; k += 128) {
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int c = c0 + 1 + w; c <= c1; c += 8) g = f4_add(g, __ldg((const float4 *)(head + (size_t)c * d + k)));
      *(float4 *)&part[w][k] = g;
    }
  }
  __syncthreads();
  if (w != 0) return;
  for (int k = lane * 4; k < d; k += 128) {
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (multi) {
      g = *(const float4 *)(tail + (size_t)c0 * d + k);
#pragma unroll
      for (int ww = 0; ww < 8; ++ww) g = f4_add(g, *(const float4 *)&part[ww][k]);
    } else if (s >= 0) {
      g = *(const float4 *)(gseg + (size_t)s * d + k);
    }
    *(float4 *)(dst + k) = g;
  }
}

__global__ void __launch_bounds__(256) k_shard_apply_users(ApplyArgs a) {
  __shared__ float sc[3];
  apply_body<kAdamFused>(a, sc);
}

// dense Adam over the item rows this rank owns: gradient = sum over the P slots (rank order) for a row the batch touched
__global__ void __launch_bounds__(256)
    k_shard_apply_items(float *__restrict__ I, float *__restrict__ mI, float *__restrict__ vI, int n_items_loc, int d,
                        const uint2 *__restrict__ row_tab_i, const uint32_t *__restrict__ ctrl, const char *__restrict__ xG,
                        size_t xG_slot, int world, int step, double lr, double beta1, double beta2, double eps_, double wd_) {
  __shared__ float sc[2];
  if (threadIdx.x == 0) {
    const double t = (double)step;
    const double bc1 = 1.0 - pow(beta1, t);
    const double bc2 = 1.0 - pow(beta2, t);
    sc[0] = (float)(-lr / bc1);
    sc[1] = (float)sqrt(bc2);
  }
  __syncthreads();
  const float neg_step = sc[0], bc2s = sc[1];
  const float w1 = (float)(1.0 - beta1), w2 = (float)(1.0 - beta2), b2 = (float)beta2, wd = (float)wd_, eps = (float)eps_;
  const int dq = d >> 2;
  const size_t nq = (size_t)n_items_loc * dq;
  const uint32_t stamp = ctrl[CTRL_STAMP] - 1u;   // k_shard_stats_reduce already advanced it
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(q / dq), k = (int)(q % dq) * 4;
    float4 *pp = (float4 *)(I + q * 4), *pm = (float4 *)(mI + q * 4), *pv = (float4 *)(vI + q * 4);
    float4 p = *pp, m = ldg_stream(pm), v = ldg_stream(pv);   // (requested before the stamp-dependent branch, see apply_body)
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint2 t = row_tab_i[row];
    if (t.x == stamp)
      for (int r = 0; r < world; ++r)
        g = f4_add(g, *(const float4 *)((const float *)(xG + (size_t)r * xG_slot) + (size_t)t.y * d + k));
    adam1(p.x, m.x, v.x, g.x, wd, w1, b2, w2, bc2s, eps, neg_step);
    adam1(p.y, m.y, v.y, g.y, wd, w1, b2, w2, bc2s, eps, neg_step);
    adam1(p.z, m.z, v.z, g.z, wd, w1, b2, w2, bc2s, eps, neg_step);
    adam1(p.w, m.w, v.w, g.w, wd, w1, b2, w2, bc2s, eps, neg_step);
    *pp = p;
    stg_stream(pm, m);
    stg_stream(pv, v);
  }
}

// STAGE: owners push the current rows of the listed items into every rank's staging table
__global__ void __launch_bounds__(256)
    k_shard_push_rows(const float *__restrict__ I, int d, const int32_t *__restrict__ stage_items, int stage_J, int rank,
                      int world, Peers px, size_t xI_off) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int j = warp; j < stage_J; j += nwarps) {
    const int it = stage_items[j];
    if (it % world != rank) continue;
    const float4 *src = (const float4 *)(I + (size_t)(it / world) * d);
    for (int k = lane; k * 4 < d; k += 32) {
      const float4 v = __ldg(src + k);
      for (int r = 0; r < world; ++r) *((float4 *)((float *)(px.base[r] + xI_off) + (size_t)j * d) + k) = v;
    }
  }
}

// ------------------------------------------------------------------------------------------ host side
static int shard_check(const fr_focf_shard_step *s, const char *who) {
  FR_REQUIRE(s, "%s: null step", who);
  FR_REQUIRE(s->world >= 1 && s->world <= FR_MAX_RANKS && s->rank >= 0 && s->rank < s->world, "%s: bad rank %d / world %d",
             who, s->rank, s->world);
  FR_REQUIRE(s->U && s->I && s->mU && s->vU && s->mI && s->vI && s->workspace && s->status_flags && s->loss,
             "%s: null table / moment / workspace pointer", who);
  FR_REQUIRE(s->d >= 4 && s->d % 4 == 0 && s->d <= kMaxD, "%s: embedding size %d must be a multiple of 4 in [4,%d]", who,
             s->d, kMaxD);
  FR_REQUIRE(s->n_users_loc >= 1 && s->n_items_loc >= 1 && s->n_items >= 1, "%s: empty table", who);
  FR_REQUIRE(s->J_cap >= 1, "%s: J_cap missing", who);
  for (int k = 0; k < s->world; ++k) FR_REQUIRE(s->xchg[k], "%s: exchange memory of rank %d missing", who, k);
  FR_REQUIRE(s->objective >= FR_OBJ_NONE && s->objective <= FR_OBJ_OVER,
             "%s: objective %d is not available row-sharded (nonparity needs batch-global group means)", who, s->objective);
  FR_REQUIRE(s->adam_mode == FR_ADAM_DENSE_EXACT || s->adam_mode == FR_ADAM_LAZY_EXACT, "%s: unknown adam_mode", who);
  if (s->adam_mode == FR_ADAM_LAZY_EXACT)
    FR_REQUIRE(s->last_step_u && s->last_step_i && s->adam_scalars && s->scalars_filled, "%s: lazy_exact state missing", who);
  return FR_OK;
}

static Peers peers_of(const fr_focf_shard_step *s) {
  Peers p;
  for (int k = 0; k < FR_MAX_RANKS; ++k) p.base[k] = k < s->world ? (char *)s->xchg[k] : nullptr;
  return p;
}

static void xbar(const fr_focf_shard_step *s, const XchgLayout &L, cudaStream_t st) {
  if (!s->barriers) return;
  FR_LAUNCH(k_xbar, 1, 32, 0, st, peers_of(s), L.flags, L.epoch, s->rank, s->world, s->status_flags);
}

static LazyCommon lazy_common(const fr_focf_shard_step *s) {
  return LazyCommon{s->d, s->step, s->lr, s->beta1, s->beta2, s->eps, s->weight_decay, s->adam_scalars, s->scalars_cap,
                    s->scalars_filled};
}

static LazyArgs lazy_items(const fr_focf_shard_step *s, const int32_t *items, const int32_t *slots, int n) {
  LazyArgs a = lazy_base(lazy_common(s), s->I, s->mI, s->vI, s->last_step_i);
  a.items = items; a.slots = slots; a.n_items_listed = n; a.rank = s->rank; a.world = s->world;
  return a;
}

// push the rows of stage_items this rank owns into every rank's staging table (parity stage_parity), then barrier
static int phase_stage(const fr_focf_shard_step *s, const XchgLayout &L, bool after_update, cudaStream_t st) {
  FR_REQUIRE(s->stage_items && s->stage_J >= 1 && s->stage_J <= s->J_cap, "fr_focf_shard_step_run: bad stage list");
  if (s->adam_mode == FR_ADAM_LAZY_EXACT) {
    // the rows about to be read by every rank must be current: replay their missed steps first.  Combined with phase C
    // the tables stand after step `step`, a stand-alone STAGE runs before step `step`.
    fr_focf_shard_step t = *s;
    t.step = after_update ? s->step : s->step - 1;
    if (t.step >= 1) {
      int rc = lazy_fill_scalars(lazy_common(&t), st, "fr_focf_shard_step_run");
      if (rc) return rc;
      LazyArgs fl = lazy_items(&t, s->stage_items, nullptr, s->stage_J);
      fl.upto = t.step;
      lazy_listed_items<false>(fl, s->last_step_i, st);
    }
  }
  FR_LAUNCH(k_shard_push_rows, grid_for((int64_t)s->stage_J * 32, 256, kSMs * 8), 256, 0, st, s->I, s->d, s->stage_items,
            s->stage_J, s->rank, s->world, peers_of(s), L.xI + (size_t)(s->stage_parity & 1) * L.xI_par);
  xbar(s, L, st);
  return FR_OK;
}

static int check_batch(const fr_focf_shard_step *s) {
  FR_REQUIRE(s->draw_items && s->draw_off && s->draw_slot && s->item_off && s->train_uid && s->train_rating && s->sst_of_user,
             "fr_focf_shard_step_run: batch description incomplete");
  FR_REQUIRE(s->J >= 1 && s->J <= s->J_cap && s->B_loc >= 0 && s->B_glob >= 1 && s->B_glob >= s->B_loc,
             "fr_focf_shard_step_run: bad batch sizes (J %d, J_cap %d, B_loc %d, B_glob %d)", s->J, s->J_cap, s->B_loc, s->B_glob);
  FR_REQUIRE(s->uid && s->iid && s->rating && s->sst && s->pred, "fr_focf_shard_step_run: batch columns missing");
  FR_REQUIRE(s->step >= 1, "fr_focf_shard_step_run: step must be the 1-based optimizer step");
  return FR_OK;
}

static LossArgs shard_loss_args(const fr_focf_shard_step *s, const ShardWs &w) {
  LossArgs a{};
  a.pred = s->pred; a.rating = s->rating; a.sst = s->sst; a.ord_i = nullptr;
  a.segid_i = w.f.segid_i; a.segoff_i = w.f.segoff_i; a.J = w.f.J; a.B = s->B_loc; a.B_dev = nullptr;
  a.objective = s->objective; a.fair_weight = s->fair_weight;
  a.cseg = w.f.cseg; a.rec_seg = w.f.rec_seg; a.rec_head = w.f.rec_head; a.rec_tail = w.f.rec_tail; a.cglob = w.f.cglob;
  a.loss = s->loss; a.ctrl = w.f.ctrl; a.flags = s->status_flags;
  return a;
}

static int phase_a(const fr_focf_shard_step *s, const ShardWs &w, const XchgLayout &L, cudaStream_t st) {
  const int B = s->B_loc, par = s->parity & 1;
  const float *xI = (const float *)((const char *)s->xchg[s->rank] + L.xI + (size_t)par * L.xI_par);
  FR_CUDA_OK(cudaMemsetAsync(w.seg_of_j, 0xff, sizeof(int32_t) * (size_t)s->J, st));
  if (B > 0) {
    if (!s->prebuilt)
      FR_LAUNCH(k_shard_gather, grid_for((int64_t)B, 256, kSMs * 8), 256, 0, st, s->item_off, s->train_uid, s->train_rating,
                s->sst_of_user, s->draw_items, s->draw_off, s->J, s->uid, s->iid, s->rating, s->sst);
    // item side: rows of one draw position are adjacent and positions ascend -> segments without a sort
    build_segments((const uint32_t *)s->iid, nullptr, B, nullptr, w.f.segid_i, w.f.segoff_i, w.f.J, nullptr, nullptr,
                   w.f.entry_seg, w.f.seg, st);
    // user side: stable sort by local user row (batch order inside a user's segment), row stamps for the dense sweep
    sort_pairs((const uint32_t *)s->uid, nullptr, w.f.skey_u, w.f.ord_u, B, nullptr, bits_for((uint32_t)s->n_users_loc),
               w.f.sort, st);
    build_segments(w.f.skey_u, w.f.ord_u, B, nullptr, w.f.segid_u, w.f.segoff_u, w.f.Ju, w.f.row_tab_u,
                   w.f.ctrl + CTRL_STAMP, nullptr, w.f.seg, st);
    if (s->adam_mode == FR_ADAM_LAZY_EXACT) {   // the forward must read the touched user rows as of step - 1
      int rc = lazy_fill_scalars(lazy_common(s), st, "fr_focf_shard_step_run");
      if (rc) return rc;
      LazyArgs u = lazy_base(lazy_common(s), s->U, s->mU, s->vU, s->last_step_u);
      u.skey = w.f.skey_u; u.segoff = w.f.segoff_u; u.nseg = w.f.Ju;
      lazy_catchup_segments(u, s->last_step_u, B, st);
    }
  }
  FR_LAUNCH(k_shard_segmap, (s->J + 255) / 256, 256, 0, st, s->iid, w.f.segoff_i, w.f.J, B > 0 ? 1 : 0, s->draw_items,
            s->draw_slot, s->J, s->rank, s->world, w.f.ctrl, w.seg_of_j, w.row_tab_i);
  LossArgs la = shard_loss_args(s, w);
  if (B > 0) {
    FR_LAUNCH(k_shard_forward, grid_for((int64_t)B, 32), 256, 0, st, s->U, xI, s->uid, s->iid, s->sst, B, s->d, s->pred,
              w.f.ctrl);
    FR_LAUNCH(k_shard_loss_records, (B + kLossThreads * kLossRows - 1) / (kLossThreads * kLossRows), kLossThreads, 0, st, la);
  }
  FR_LAUNCH(k_shard_stats_push, grid_for(s->J, 1, kSMs * 8), kSegThreads, 0, st, la, w.seg_of_j, s->J, B,
            peers_of(s), L.xS + (size_t)par * L.xS_par + (size_t)s->rank * L.xS_slot, L.hdr + (size_t)par * L.hdr_par,
            s->rank, s->world);
  xbar(s, L, st);
  return FR_OK;
}

static int phase_b(const fr_focf_shard_step *s, const ShardWs &w, const XchgLayout &L, cudaStream_t st) {
  const int B = s->B_loc, par = s->parity & 1;
  const char *own = (const char *)s->xchg[s->rank];
  const float *xI = (const float *)(own + L.xI + (size_t)par * L.xI_par);
  FR_LAUNCH(k_shard_stats_reduce, grid_for(s->J, kRedThreads / 32, kRedMaxBlocks), kRedThreads, 0, st,
            own + L.xS + (size_t)par * L.xS_par, L.xS_slot, (const uint4 *)(own + L.hdr + (size_t)par * L.hdr_par), s->world,
            s->J, B, s->B_glob, s->objective, s->fair_weight, w.seg_of_j, w.cseg_j, w.f.cseg, w.f.cglob, w.red_part, s->loss,
            w.f.ctrl, s->status_flags);
  const int chunk = grad_chunk(B < 1 ? 1 : B);
  if (B > 0) {
    GradArgs ga{};
    ga.U = s->U; ga.I = xI; ga.uid = s->uid; ga.iid = s->iid; ga.rating = s->rating; ga.sst = s->sst; ga.pred = s->pred;
    ga.B = B; ga.d = s->d; ga.B_dev = nullptr; ga.norm_B = s->B_glob; ga.norm_from_ctrl = 0;
    ga.ord_i = nullptr; ga.ord_u = w.f.ord_u;
    ga.segid_i = w.f.segid_i; ga.segoff_i = w.f.segoff_i; ga.segid_u = w.f.segid_u; ga.segoff_u = w.f.segoff_u;
    ga.entry_seg = w.f.entry_seg; ga.cseg = w.f.cseg; ga.cglob = w.f.cglob; ga.ctrl = w.f.ctrl; ga.grad_scale = 1.0f;
    ga.chunk = chunk;
    ga.gseg_i = w.f.gseg_i; ga.head_i = w.f.head_i; ga.tail_i = w.f.tail_i;
    ga.gseg_u = w.f.gseg_u; ga.head_u = w.f.head_u; ga.tail_u = w.f.tail_u;
    ga.pre_handover = 0;
    const int nchunk = (B + chunk - 1) / chunk;
    const int grid = (2 * nchunk + 7) / 8;
    if (s->d <= 128) {
      FR_LAUNCH(k_shard_grads<1>, grid, 256, 0, st, ga, nchunk);
    } else if (s->d <= 256) {
      FR_LAUNCH(k_shard_grads<2>, grid, 256, 0, st, ga, nchunk);
    } else {
      FR_LAUNCH(k_shard_grads<4>, grid, 256, 0, st, ga, nchunk);
    }
  }
  FR_LAUNCH(k_shard_igrad_push, s->J, 256, 0, st, w.seg_of_j, w.f.segoff_i,
            w.f.gseg_i, w.f.head_i, w.f.tail_i, chunk, s->d, s->draw_items, s->draw_slot, s->J, B, peers_of(s),
            L.xG + (size_t)par * L.xG_par + (size_t)s->rank * L.xG_slot, s->world);
  xbar(s, L, st);
  return FR_OK;
}

static int phase_c(const fr_focf_shard_step *s, const ShardWs &w, const XchgLayout &L, cudaStream_t st) {
  const int B = s->B_loc, par = s->parity & 1;
  const char *own = (const char *)s->xchg[s->rank];
  const char *xG = own + L.xG + (size_t)par * L.xG_par;
  const int chunk = grad_chunk(B < 1 ? 1 : B);
  if (s->adam_mode == FR_ADAM_DENSE_EXACT) {
    ApplyArgs aa{};
    aa.U = s->U; aa.mU = s->mU; aa.vU = s->vU; aa.n_users = s->n_users_loc; aa.n_items = 0; aa.d = s->d;
    aa.row_tab_u = w.f.row_tab_u; aa.segoff_u = w.f.segoff_u; aa.gseg_u = w.f.gseg_u; aa.head_u = w.f.head_u;
    aa.tail_u = w.f.tail_u; aa.ctrl = w.f.ctrl; aa.step = s->step;
    aa.lr = s->lr; aa.beta1 = s->beta1; aa.beta2 = s->beta2; aa.eps = s->eps; aa.wd = s->weight_decay;
    aa.chunk = chunk; aa.pre_handover = 0;
    FR_LAUNCH(k_shard_apply_users, grid_for((int64_t)s->n_users_loc * (s->d / 4), 256, kSMs * 16), 256, 0, st, aa);
    FR_LAUNCH(k_shard_apply_items, grid_for((int64_t)s->n_items_loc * (s->d / 4), 256, kSMs * 16), 256, 0, st, s->I, s->mI,
              s->vI, s->n_items_loc, s->d, w.row_tab_i, w.f.ctrl, xG, L.xG_slot, s->world, s->step, s->lr, s->beta1,
              s->beta2, s->eps, s->weight_decay);
    return FR_OK;
  }
  int rc = lazy_fill_scalars(lazy_common(s), st, "fr_focf_shard_step_run");
  if (rc) return rc;
  if (B > 0) {
    LazyArgs u = lazy_base(lazy_common(s), s->U, s->mU, s->vU, s->last_step_u);
    u.skey = w.f.skey_u; u.segoff = w.f.segoff_u; u.nseg = w.f.Ju; u.gseg = w.f.gseg_u; u.head = w.f.head_u;
    u.tail = w.f.tail_u; u.chunk = chunk;
    lazy_apply_segments(u, s->last_step_u, B, st);
  }
  LazyArgs it = lazy_items(s, s->draw_items, s->draw_slot, s->J);
  for (int k = 0; k < s->world; ++k) it.slot_grad[k] = (const float *)(xG + (size_t)k * L.xG_slot);
  lazy_listed_items<true>(it, s->last_step_i, st);
  return FR_OK;
}

}  // namespace fr

extern "C" {

int fr_xchg_alloc(size_t bytes, void **ptr_out) {
  FR_REQUIRE(ptr_out && bytes > 0, "fr_xchg_alloc: bad argument");
  FR_CUDA_OK(cudaMalloc(ptr_out, bytes));
  FR_CUDA_OK(cudaMemset(*ptr_out, 0, bytes));
  return FR_OK;
}

int fr_xchg_free(void *ptr) {
  if (ptr) FR_CUDA_OK(cudaFree(ptr));
  return FR_OK;
}

int fr_xchg_export(void *ptr, void *handle64_out) {
  FR_REQUIRE(ptr && handle64_out, "fr_xchg_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  FR_CUDA_OK(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle64_out, ptr));
  return FR_OK;
}

int fr_xchg_open(const void *handle64, void **peer_ptr_out) {
  FR_REQUIRE(handle64 && peer_ptr_out, "fr_xchg_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  FR_CUDA_OK(cudaIpcOpenMemHandle(peer_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return FR_OK;
}

int fr_xchg_close(void *peer_ptr) {
  if (peer_ptr) FR_CUDA_OK(cudaIpcCloseMemHandle(peer_ptr));
  return FR_OK;
}

size_t fr_focf_shard_xchg_bytes(int32_t world, int32_t J_cap, int32_t d) {
  if (world < 1 || world > FR_MAX_RANKS || J_cap < 1 || d < 4) return 0;
  return fr::xchg_layout(world, J_cap, d).total;
}

size_t fr_focf_shard_workspace_bytes(int32_t n_users_loc, int32_t n_items_loc, int32_t d, int32_t max_batch, int32_t J_cap) {
  fr::Carver c(nullptr, 0);
  fr::carve_shard(c, n_users_loc, n_items_loc, d, max_batch, J_cap);
  return c.off;
}

int fr_focf_shard_workspace_init(void *workspace, size_t workspace_bytes, int32_t n_users_loc, int32_t n_items_loc, int32_t d,
                                 int32_t max_batch, int32_t J_cap, void *stream) {
  FR_REQUIRE(workspace, "fr_focf_shard_workspace_init: null workspace");
  fr::Carver c(workspace, workspace_bytes);
  fr::ShardWs w = fr::carve_shard(c, n_users_loc, n_items_loc, d, max_batch, J_cap);
  if (!c.ok()) {
    fr::set_error("fr_focf_shard_workspace_init: workspace too small (%zu < %zu bytes)", workspace_bytes, c.off);
    return FR_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  FR_CUDA_OK(cudaMemsetAsync(w.f.row_tab_u, 0, sizeof(uint2) * (size_t)n_users_loc, st));
  FR_CUDA_OK(cudaMemsetAsync(w.row_tab_i, 0, sizeof(uint2) * (size_t)(n_items_loc < 1 ? 1 : n_items_loc), st));
  uint32_t ctrl[fr::CTRL_WORDS] = {0};
  ctrl[fr::CTRL_STAMP] = 1u;
  ctrl[fr::CTRL_MIN] = 0xffffffffu;
  ctrl[fr::CTRL_STRIDE] = 1u;
  FR_CUDA_OK(cudaMemcpyAsync(w.f.ctrl, ctrl, sizeof(ctrl), cudaMemcpyHostToDevice, st));
  FR_CUDA_OK(cudaStreamSynchronize(st));
  return FR_OK;
}

int fr_focf_shard_step_run(const fr_focf_shard_step *s, int32_t phases, void *stream) {
  int rc = fr::shard_check(s, "fr_focf_shard_step_run");
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const fr::XchgLayout L = fr::xchg_layout(s->world, s->J_cap, s->d);
  const int step_phases = phases & (FR_SHARD_A | FR_SHARD_B | FR_SHARD_C);
  fr::ShardWs w{};
  if (step_phases) {
    if ((rc = fr::check_batch(s))) return rc;
    fr::Carver c(s->workspace, s->workspace_bytes);
    w = fr::carve_shard(c, s->n_users_loc, s->n_items_loc, s->d, s->B_loc < 1 ? 1 : s->B_loc, s->J_cap);
    if (!c.ok()) {
      fr::set_error("fr_focf_shard_step_run: workspace too small (%zu < %zu bytes)", s->workspace_bytes, c.off);
      return FR_ERR_WORKSPACE;
    }
  }
  if (phases & FR_SHARD_A) if ((rc = fr::phase_a(s, w, L, st))) return rc;
  if (phases & FR_SHARD_B) if ((rc = fr::phase_b(s, w, L, st))) return rc;
  if (phases & FR_SHARD_C) if ((rc = fr::phase_c(s, w, L, st))) return rc;
  if (phases & FR_SHARD_STAGE) if ((rc = fr::phase_stage(s, L, (phases & FR_SHARD_C) != 0, st))) return rc;
  if ((phases & FR_SHARD_FLUSH) && s->adam_mode == FR_ADAM_LAZY_EXACT && s->step >= 1) {
    fr::LazyCommon c = fr::lazy_common(s);
    if ((rc = fr::lazy_fill_scalars(c, st, "fr_focf_shard_step_run"))) return rc;
    fr::lazy_flush_table(fr::lazy_base(c, s->U, s->mU, s->vU, s->last_step_u), s->last_step_u, s->n_users_loc, st);
    fr::lazy_flush_table(fr::lazy_base(c, s->I, s->mI, s->vI, s->last_step_i), s->last_step_i, s->n_items_loc, st);
  }
  FR_LAUNCH_CHECK();
  return FR_OK;
}

}  // extern "C"
