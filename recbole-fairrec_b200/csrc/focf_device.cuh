// Device-side building blocks of the FOCF training step, shared by focf_train.cu (single-GPU step) and
// focf_shard.cu (row-sharded multi-GPU step).  See focf_train.cu for the reference lines each part replaces.
#pragma once
#include "sort.cuh"

namespace fr {

constexpr int kMaxD = 512;            // embedding_size limit (multiple of 4)
constexpr int kChunk = 32;            // max sorted entries per gradient warp (runtime: 8 for small batches)
constexpr int kChunkMin = 8;
// 16 entries per warp for small batches (more warps, shorter dependent chains), 32 for large ones; rows spanning many
// chunks cost k_apply one partial read per chunk, so the chunk must not get too small
static inline int grad_chunk(int B) { return B <= 16384 ? 16 : kChunk; }

enum {
  CTRL_STAMP = 0,      // batch counter: tags row_tab entries, advanced by the last CTA of k_segment_loss
  CTRL_MIN = 1, CTRL_MAX = 2,              // running min / max of the batch's attribute values (order-encoded)
  CTRL_TICKET = 3,
  CTRL_SAVED_MIN = 4, CTRL_SAVED_MAX = 5,  // the same, handed to the backward kernels
  CTRL_CURSOR = 6,     // planned-batch cursor (fr_focf_plan): which batch of the epoch plan comes next
  CTRL_ADAM_T = 7,     // device-resident Adam step count (used when fr_focf_step.step <= 0)
  CTRL_B = 8,          // device-resident batch size (planned batches)
  CTRL_GRID_BAR = 14,  // arrival counter of the fused step's grid barrier: 64-bit (words 14-15, 8-byte aligned), monotonic
  CTRL_STRIDE = 10,    // how far CTRL_CURSOR / CTRL_ADAM_T advance per step (2 when two workspaces alternate batches)
  CTRL_NORM_B = 11, CTRL_NORM_J = 12,   // data-parallel planned steps: the current batch's global normalisers (from norm_dev)
  CTRL_WORDS = 64
};
// batch size: host value unless a device-resident one is given (CUDA-graph replay over batches of varying size)
#define FR_B(host_B, dev_B) ((dev_B) ? *(dev_B) : (host_B))

struct FocfWs {
  // persistent across steps
  uint2 *row_tab_u, *row_tab_i;  // [n_users], [n_items]: {stamp, segment} of the last batch touching the row
  uint32_t *ctrl;                // [CTRL_WORDS]
  // per batch
  uint32_t *skey_i, *ord_i, *skey_u, *ord_u;          // [B] sorted keys / entry order
  int32_t *segid_i, *segoff_i, *J, *segid_u, *segoff_u, *Ju, *entry_seg;
  float *cseg;      // [B,2] additive dL/dpred term of every (item segment, group)
  float *rec_seg;   // [B,8]     loss-statistics record of a segment lying inside one 8-row thread chunk
  float *rec_head;  // [B/8+2,8] record of the run continuing from the previous chunk
  float *rec_tail;  // [B/8+2,8] record of the run continuing into the next chunk
  float *cglob;     // [2]  batch-global additive term per group (nonparity)
  float *gseg_i, *head_i, *tail_i, *gseg_u, *head_u, *tail_u;  // gradient partials [B,d], [B/32+1,d] x2
  float *red_part;  // [148 * 8, 8] per-CTA partial sums of k_segment_reduce
  SortScratch sort;
  SegScratch seg;
};

static FocfWs carve(Carver &c, int n_users, int n_items, int d, int B) {
  FocfWs w;
  w.row_tab_u = c.take<uint2>(n_users);
  w.row_tab_i = c.take<uint2>(n_items);
  w.ctrl = c.take<uint32_t>(CTRL_WORDS);
  const size_t b = (size_t)(B < 1 ? 1 : B), nch = b / kChunkMin + 2;
  w.skey_i = c.take<uint32_t>(b);
  w.ord_i = c.take<uint32_t>(b);
  w.skey_u = c.take<uint32_t>(b);
  w.ord_u = c.take<uint32_t>(b);
  w.segid_i = c.take<int32_t>(b);
  w.segoff_i = c.take<int32_t>(b + 1);
  w.J = c.take<int32_t>(1);
  w.segid_u = c.take<int32_t>(b);
  w.segoff_u = c.take<int32_t>(b + 1);
  w.Ju = c.take<int32_t>(1);
  w.entry_seg = c.take<int32_t>(b);
  w.cseg = c.take<float>(2 * b);
  w.rec_seg = c.take<float>(8 * b);
  w.rec_head = c.take<float>(8 * (b / 8 + 2));
  w.rec_tail = c.take<float>(8 * (b / 8 + 2));
  w.cglob = c.take<float>(2);
  w.gseg_i = c.take<float>(b * d);
  w.head_i = c.take<float>(nch * d);
  w.tail_i = c.take<float>(nch * d);
  w.gseg_u = c.take<float>(b * d);
  w.head_u = c.take<float>(nch * d);
  w.tail_u = c.take<float>(nch * d);
  w.red_part = c.take<float>(8 * 148 * 8);
  w.sort = carve_sort_scratch(c, b);
  w.seg = carve_seg_scratch(c, b);
  return w;
}

// ------------------------------------------------------------------------------------------ forward
// One warp per entry: lanes cover the row with float4 loads (d floats = d/4 lanes per 128 columns).
// Also folds the batch min/max of the sensitive attribute (the "rank among present values" of
// torch.unique, focf.py:77) into two integer atomics.
__device__ __forceinline__ void forward_body(const float *__restrict__ U, const float *__restrict__ I,
                                             const int32_t *__restrict__ uid, const int32_t *__restrict__ iid,
                                             const float *__restrict__ sst, int B, int d, float *__restrict__ pred,
                                             uint32_t *__restrict__ ctrl) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  uint32_t lo = 0xffffffffu, hi = 0u;
  // lane e < 4 holds the ids and the attribute of entry b0 + e; those of the NEXT group of four are fetched before the
  // row loads of the current one are issued, so the id -> row dependency is off the critical path after the first group
  int nu = 0, ni = 0;
  float ns = 0.f;
  int b0 = warp * 4;
  if (lane < 4 && b0 + lane < B) {
    nu = uid[b0 + lane];
    ni = iid[b0 + lane];
    ns = sst[b0 + lane];
  }
  for (; b0 < B; b0 += nwarps * 4) {
    const int cu = nu, ci = ni;
    const float cs = ns;
    const int nb = b0 + nwarps * 4 + lane;
    if (lane < 4 && nb < B) {
      nu = uid[nb];
      ni = iid[nb];
      ns = sst[nb];
    }
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int u = __shfl_sync(0xffffffffu, cu, e), it = __shfl_sync(0xffffffffu, ci, e);
      if (b0 + e < B) {
        const float4 *pu = (const float4 *)(U + (size_t)u * d);
        const float4 *pi = (const float4 *)(I + (size_t)it * d);
        for (int k = lane; k * 4 < d; k += 32) {
          const float4 x = __ldg(pu + k), y = __ldg(pi + k);
          acc[e] = fmaf(x.x, y.x, acc[e]);
          acc[e] = fmaf(x.y, y.y, acc[e]);
          acc[e] = fmaf(x.z, y.z, acc[e]);
          acc[e] = fmaf(x.w, y.w, acc[e]);
        }
      }
    }
    float mine = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float s = warp_sum(acc[e]);
      if (lane == e) mine = s;
    }
    if (lane < 4 && b0 + lane < B) {
      pred[b0 + lane] = mine;
      const uint32_t o = f2ord(cs);
      lo = min(lo, o);
      hi = max(hi, o);
    }
  }
  lo = __reduce_min_sync(0xffffffffu, lo);
  hi = __reduce_max_sync(0xffffffffu, hi);
  if (lane == 0 && hi >= lo) {
    atomicMin(&ctrl[CTRL_MIN], lo);
    atomicMax(&ctrl[CTRL_MAX], hi);
  }
}

// ------------------------------------------------------------------------------------------ segment loss
// Item x group statistics, fairness objective and loss (appendix A.1/A.2 of SURVEY.md) in one launch with an even
// work split that does not depend on item popularity:
//   phase 1 (all CTAs): thread t owns rows 8t..8t+7 of the item-sorted order and accumulates, per run of equal
//            segment, (sum pred, sum rating, count) per group and sum (pred-r)^2.  A segment that lies inside the
//            8 rows is written complete; a run continuing from / into a neighbouring thread leaves a head / tail record.
//   phase 2 (the last CTA to finish, self-resetting ticket): warp per segment adds its records in a fixed order,
//            derives D, smooth-L1 and the additive backward term cseg[j][g]; a fixed-order block reduction gives the loss.
constexpr int kLossRows = 8;       // rows per thread in phase 1
constexpr int kLossThreads = 1024;
constexpr int kLossRec = 8;        // floats per record: sp0 sp1 st0 st1 c0 c1 sq (pad)

struct LossArgs {
  const float *pred, *rating, *sst;
  const uint32_t *ord_i;
  const int32_t *segid_i, *segoff_i, *J;
  int B;
  const int32_t *B_dev;
  int loss_by_cursor;   // planned batches: plan length L > 0 -> write loss[cursor % L] instead of loss[0]
  int advance_adam;     // fused step with the device-resident Adam counter
  int objective;
  int norm_B, norm_J;   // > 0: global batch rows / item count of a data-parallel step (normalisers of the two means)
  const int32_t *norm_dev;   // planned data-parallel steps: [plan_len, 2] = (B_total, J_total) of every planned batch
  float fair_weight;
  float *cseg, *rec_seg, *rec_head, *rec_tail, *cglob, *loss;
  uint32_t *ctrl;
  int32_t *flags;
};

__device__ __forceinline__ float block_sum_1024(float v, float *sh) {  // sh: >= 33 floats
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.f;   // any block size up to 1024 threads
  if (threadIdx.x < 32) {
    t = warp_sum(t);
    if (threadIdx.x == 0) sh[32] = t;
  }
  __syncthreads();
  t = sh[32];
  __syncthreads();
  return t;
}

// fairness objective of ONE item segment (appendix A.1/A.2 of SURVEY.md): smooth-L1 term hx and the additive
// dL/dpred terms cs0 / cs1 of the segment's two groups
__device__ __forceinline__ void segment_terms(int objective, float fair_weight, float Jn, float sp0, float sp1, float st0,
                                              float st1, float c0, float c1, float &hx, float &cs0, float &cs1) {
  const float n0 = c0 + 1e-5f, n1 = c1 + 1e-5f;                       // focf.py:89
  const float P0 = sp0 / n0, P1 = sp1 / n1, T0 = st0 / n0, T1 = st1 / n1;  // focf.py:91
  float D0, D1, dd0, dd1;
  switch (objective) {
    case FR_OBJ_VALUE:    D0 = P0 - T0; D1 = P1 - T1; dd0 = 1.f; dd1 = 1.f; break;
    case FR_OBJ_ABSOLUTE: {
      const float e0 = P0 - T0, e1 = P1 - T1;
      D0 = fabsf(e0); D1 = fabsf(e1);
      dd0 = (e0 > 0.f) - (e0 < 0.f); dd1 = (e1 > 0.f) - (e1 < 0.f);
    } break;
    case FR_OBJ_UNDER: {
      const float e0 = T0 - P0, e1 = T1 - P1;
      D0 = e0 > 0.f ? e0 : 0.f; D1 = e1 > 0.f ? e1 : 0.f;
      dd0 = e0 > 0.f ? -1.f : 0.f; dd1 = e1 > 0.f ? -1.f : 0.f;
    } break;
    default: {
      const float e0 = P0 - T0, e1 = P1 - T1;
      D0 = e0 > 0.f ? e0 : 0.f; D1 = e1 > 0.f ? e1 : 0.f;
      dd0 = e0 > 0.f ? 1.f : 0.f; dd1 = e1 > 0.f ? 1.f : 0.f;
    }
  }
  const float z = D0 - D1, x = fabsf(z);
  hx = x < 1.f ? 0.5f * x * x : x - 0.5f;                              // smooth_l1, beta = 1
  const float hp = (x < 1.f ? x : 1.f) * (float)((z > 0.f) - (z < 0.f));
  const float q = fair_weight * hp / Jn;
  cs0 = q * dd0 / n0;
  cs1 = -q * dd1 / n1;
}

__device__ __forceinline__ void loss_phase1(const LossArgs &a, int B) {
  const float vmin = ord2f(a.ctrl[CTRL_MIN]), vmax = ord2f(a.ctrl[CTRL_MAX]);
  {
    int bad = 0;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t * kLossRows < B; t += gridDim.x * blockDim.x) {
      const int lo = t * kLossRows, hi = lo + kLossRows;
      float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      int cur = a.segid_i[lo];
      auto flush = [&](int sgm) {
        const int s0 = a.segoff_i[sgm], s1 = a.segoff_i[sgm + 1];
        float *dst = (s0 >= lo && s1 <= hi) ? a.rec_seg + (size_t)sgm * kLossRec
                     : (s0 < lo)            ? a.rec_head + (size_t)t * kLossRec
                                            : a.rec_tail + (size_t)t * kLossRec;
        *(float4 *)dst = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *(float4 *)(dst + 4) = make_float4(acc[4], acc[5], acc[6], 0.f);
#pragma unroll
        for (int k = 0; k < 7; ++k) acc[k] = 0.f;
      };
      // issue every load of the 8 rows before consuming any (rows past B re-read row lo and are ignored)
      int sg[kLossRows];
      float prv[kLossRows], rv[kLossRows], svv[kLossRows];
#pragma unroll
      for (int i = 0; i < kLossRows; ++i) {
        const int p = (lo + i < B) ? lo + i : lo;
        const int b = a.ord_i ? (int)a.ord_i[p] : p;
        sg[i] = a.segid_i[p];
        prv[i] = a.pred[b];
        rv[i] = a.rating[b];
        svv[i] = a.sst[b];
      }
#pragma unroll
      for (int i = 0; i < kLossRows; ++i) {
        if (lo + i < B) {
          if (sg[i] != cur) {
            flush(cur);
            cur = sg[i];
          }
          const float pr = prv[i], r = rv[i], sv = svv[i];
          const bool g = sv != vmin;
          bad |= (g && sv != vmax);
          const float df = pr - r;
          acc[6] = fmaf(df, df, acc[6]);
          if (g) {
            acc[1] += pr; acc[3] += r; acc[5] += 1.f;
          } else {
            acc[0] += pr; acc[2] += r; acc[4] += 1.f;
          }
        }
      }
      flush(cur);
    }
    // focf.py:81-86: a third attribute value indexes past the [J,2] tensors (IndexError in the reference).
    // (nonparity in the reference silently keeps the two smallest values, focf.py:129-130; we flag instead.)
    if (a.objective != FR_OBJ_NONE && bad) atomicOr(a.flags, FR_FLAG_TOO_MANY_GROUPS);
  }

}

// Sum of the phase-1 records of ONE item segment by a whole CTA (kSegThreads threads): a popular item's rows span
// thousands of 8-row thread chunks, so the records are strided over all threads (thread t takes records t, t + T, ... in
// increasing order), then combined by the fixed shuffle tree of warp_sum and across warps in warp order -- deterministic.
// Every thread returns with the 7 sums in v.  `sh` holds kSegThreads / 32 rows of 8 floats.
constexpr int kSegThreads = 128;
__device__ __forceinline__ void segment_record_sum(const LossArgs &a, int sgm, float v[7], float (*sh)[8]) {
  const int s0 = a.segoff_i[sgm], s1 = a.segoff_i[sgm + 1];
  const int t0 = s0 / kLossRows, t1 = (s1 - 1) / kLossRows;
  if (t0 == t1) {   // the segment lies inside one thread chunk: one complete record
    const float4 x = *(const float4 *)(a.rec_seg + (size_t)sgm * kLossRec);
    const float4 y = *(const float4 *)(a.rec_seg + (size_t)sgm * kLossRec + 4);
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z;
    return;
  }
#pragma unroll
  for (int k = 0; k < 7; ++k) v[k] = 0.f;
  for (int t = t0 + (int)threadIdx.x; t <= t1; t += kSegThreads) {
    const float *src = (t == t0) ? a.rec_tail + (size_t)t * kLossRec : a.rec_head + (size_t)t * kLossRec;
    const float4 x = *(const float4 *)src, y = *(const float4 *)(src + 4);
    v[0] += x.x; v[1] += x.y; v[2] += x.z; v[3] += x.w; v[4] += y.x; v[5] += y.y; v[6] += y.z;
  }
#pragma unroll
  for (int k = 0; k < 7; ++k) v[k] = warp_sum(v[k]);
  __syncthreads();            // sh may still be read from the previous segment
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 7; ++k) sh[threadIdx.x >> 5][k] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kSegThreads / 32; ++w) t += sh[w][k];
    v[k] = t;
  }
}

// Item x group statistics of the WHOLE batch, computed redundantly by every CTA out of shared memory: the batch is a
// few thousand rows (36 KB of pred / rating / sst), so re-reading it per CTA is cheaper than a grid-wide hand-over --
// it removes the single-CTA reduction phase and one grid barrier from the step.  Rows are staged with one coalesced
// sweep, then a warp owns a segment (lanes stride its rows: popularity skew costs shared-memory, not DRAM, round
// trips).  Every CTA runs the same code on the same data in the same order -> bit-identical cseg everywhere.
// Returns (on thread 0 of CTA 0 only meaningful) the batch loss.
// rest_staged: rating / attribute columns are already in shared memory (persistent epoch kernel: staged while it waits at
// the grid barrier); need_sums == false: this CTA does not need the batch loss (only one CTA writes it), so the block-wide
// sums are skipped unless the objective's backward term depends on them (nonparity).
// pre: values the caller fetched ahead (persistent epoch kernel): segment count, batch min / max of the attribute, and the
// bounds of the warp's first segment.
struct StatsPre {
  int J;
  uint32_t vmin, vmax;   // order-encoded
  int s0, s1;            // bounds of segment (threadIdx.x >> 5)
};
__device__ __forceinline__ float fused_stats(const LossArgs &a, int B, int cap, float *sm, float *sh, float *s_cseg,
                                             float *s_cglob, bool rest_staged = false, bool need_sums = true,
                                             int lead_cta = 0,   // lead_cta: the CTA that raises the status flags
                                             const StatsPre *pre = nullptr) {
  float *s_pred = sm, *s_rat = sm + cap, *s_sst = sm + 2 * cap;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int J = pre ? pre->J : *a.J;
  const float Bn = (float)B, Jn = (float)J;
  const float vmin = ord2f(pre ? pre->vmin : a.ctrl[CTRL_MIN]), vmax = ord2f(pre ? pre->vmax : a.ctrl[CTRL_MAX]);
  // (four rows per thread in flight: a rolled loop costs one L2 round trip per 512 rows)
  for (int p0 = threadIdx.x; p0 < B; p0 += 4 * blockDim.x) {
    int b[4];
    float vp[4], vr[4], vs[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = p0 + i * blockDim.x;
      b[i] = p < B ? (a.ord_i ? (int)a.ord_i[p] : p) : 0;     // item-sorted order (identity for whole-item batches)
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      vp[i] = a.pred[b[i]];
      if (!rest_staged) {
        vr[i] = a.rating[b[i]];
        vs[i] = a.sst[b[i]];
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = p0 + i * blockDim.x;
      if (p < B) {
        s_pred[p] = vp[i];
        if (!rest_staged) {
          s_rat[p] = vr[i];
          s_sst[p] = vs[i];
        }
      }
    }
  }
  __syncthreads();
  float w_sq = 0.f, w_hx = 0.f, w_g0 = 0.f, w_g1 = 0.f, w_n0 = 0.f, w_n1 = 0.f;
  int bad = 0;
  for (int j = wib; j < J; j += nw) {
    const bool first = pre != nullptr && j == wib;
    const int s0 = first ? pre->s0 : a.segoff_i[j], s1 = first ? pre->s1 : a.segoff_i[j + 1];
    float sp0 = 0.f, sp1 = 0.f, st0 = 0.f, st1 = 0.f, c0 = 0.f, c1 = 0.f, sq = 0.f;
#pragma unroll 4
    for (int p = s0 + lane; p < s1; p += 32) {   // (unrolled: the shared-memory loads of a popular item's rows pipeline)
      const float pr = s_pred[p], r = s_rat[p], sv = s_sst[p];
      const bool g = sv != vmin;
      bad |= (g && sv != vmax);
      const float df = pr - r;
      sq = fmaf(df, df, sq);
      if (g) { sp1 += pr; st1 += r; c1 += 1.f; } else { sp0 += pr; st0 += r; c0 += 1.f; }
    }
    sp0 = warp_sum(sp0); sp1 = warp_sum(sp1); st0 = warp_sum(st0); st1 = warp_sum(st1);
    c0 = warp_sum(c0); c1 = warp_sum(c1); sq = warp_sum(sq);
    if (lane == 0) {
      float hx = 0.f, cs0 = 0.f, cs1 = 0.f;
      w_sq += sq;
      if (a.objective >= FR_OBJ_VALUE && a.objective <= FR_OBJ_OVER) {
        segment_terms(a.objective, a.fair_weight, Jn, sp0, sp1, st0, st1, c0, c1, hx, cs0, cs1);
      } else if (a.objective == FR_OBJ_NONPARITY) {
        w_g0 += sp0; w_g1 += sp1; w_n0 += c0; w_n1 += c1;
      }
      w_hx += hx;
      s_cseg[2 * j] = cs0;
      s_cseg[2 * j + 1] = cs1;
    }
  }
  if ((int)blockIdx.x == lead_cta && a.objective != FR_OBJ_NONE && __any_sync(0xffffffffu, bad) && lane == 0)
    atomicOr(a.flags, FR_FLAG_TOO_MANY_GROUPS);
  if (!need_sums && a.objective != FR_OBJ_NONPARITY) {
    if (threadIdx.x == 0) {
      s_cglob[0] = 0.f;
      s_cglob[1] = 0.f;
    }
    __syncthreads();
    return 0.f;
  }
  const float sq = block_sum_1024(w_sq, sh), hx = block_sum_1024(w_hx, sh);
  float g0 = 0.f, g1 = 0.f, n0 = 0.f, n1 = 0.f;
  if (a.objective == FR_OBJ_NONPARITY) {
    g0 = block_sum_1024(w_g0, sh); g1 = block_sum_1024(w_g1, sh);
    n0 = block_sum_1024(w_n0, sh); n1 = block_sum_1024(w_n1, sh);
  }
  float loss = sq / Bn;
  if (threadIdx.x == 0) {
    float cg0 = 0.f, cg1 = 0.f;
    if (a.objective >= FR_OBJ_VALUE && a.objective <= FR_OBJ_OVER) {
      loss += a.fair_weight * (hx / Jn);
    } else if (a.objective == FR_OBJ_NONPARITY) {
      if (n1 == 0.f || n0 == 0.f) {
        if ((int)blockIdx.x == lead_cta) atomicOr(a.flags, FR_FLAG_SINGLE_GROUP);
      } else {
        const float z = g0 / n0 - g1 / n1, x = fabsf(z);
        loss += a.fair_weight * (x < 1.f ? 0.5f * x * x : x - 0.5f);
        const float hp = a.fair_weight * (x < 1.f ? z : (float)((z > 0.f) - (z < 0.f)));
        cg0 = hp / n0;
        cg1 = -hp / n1;
      }
    }
    s_cglob[0] = cg0;
    s_cglob[1] = cg1;
  }
  __syncthreads();
  return loss;
}

// ------------------------------------------------------------------------------------------ gradients
// Sorted-segment reduction of dL/dpred_b * other_row(b).  One warp owns 32 consecutive entries of the
// (item- or user-) sorted order and walks them in order; a row whose segment lies inside the chunk is
// written complete (gseg), a row continuing from / into a neighbouring chunk leaves a head / tail partial
// that k_apply adds up in chunk order.  Uniform work per warp regardless of item popularity.
struct GradArgs {
  const float *U, *I;
  const int32_t *uid, *iid;
  const float *rating, *sst, *pred;
  int B, d;
  const int32_t *B_dev;
  int norm_B;
  int norm_from_ctrl;   // planned data-parallel steps: read the normaliser the loss kernel left in CTRL_NORM_B
  const uint32_t *ord_i, *ord_u;
  const int32_t *segid_i, *segoff_i, *segid_u, *segoff_u, *entry_seg;
  const float *cseg, *cglob;
  const uint32_t *ctrl;
  float grad_scale;
  int chunk;   // sorted entries per warp (8 or 32)
  float *gseg_i, *head_i, *tail_i, *gseg_u, *head_u, *tail_u;
  int pre_handover;   // fused step: the control block has not been handed over yet -> read CTRL_MIN, not CTRL_SAVED_MIN
};

// One warp's chunk of the sorted order in two steps: STAGE reads what the preparation left (sorted entry ids, segment ids,
// the other side's row ids, the entry's item segment) -- nothing that depends on the forward, so the persistent epoch
// kernel issues it ahead of the barrier; RUN forms the coefficients, gathers the rows and writes the partials.
struct ChunkStage {
  int c, pbase, nvalid;        // chunk index within its side, first sorted position, entries (0: nothing to do)
  int my_seg, my_oid, my_b, my_es;   // lane l: segment, other-side row, entry index, item segment of entry pbase + l
  int seg_before, seg_after;   // segments just outside the chunk (-1: none)
  bool user_side;
};

__device__ __forceinline__ ChunkStage stage_chunk(const GradArgs &a, int nchunk, int c) {   // c in [0, 2 * nchunk)
  const int lane = threadIdx.x & 31;
  ChunkStage s;
  s.user_side = c >= nchunk;
  if (s.user_side) c -= nchunk;
  s.c = c;
  s.pbase = c * a.chunk;
  const int B = FR_B(a.B, a.B_dev);
  s.nvalid = s.pbase < B ? min(a.chunk, B - s.pbase) : 0;
  s.my_seg = -1; s.my_oid = 0; s.my_b = 0; s.my_es = 0;
  s.seg_before = s.seg_after = -1;
  if (s.nvalid == 0) return s;
  const uint32_t *ord = s.user_side ? a.ord_u : a.ord_i;
  const int32_t *segid = s.user_side ? a.segid_u : a.segid_i;
  const int32_t *oid = s.user_side ? a.iid : a.uid;
  // the segments just outside the chunk: a row whose segment equals one of them continues from / into a neighbouring
  // chunk -- known without reading the segment offsets at every flush
  int edge = -1;
  if (lane == 0 && s.pbase > 0) edge = segid[s.pbase - 1];
  if (lane == 1 && s.pbase + s.nvalid < B) edge = segid[s.pbase + s.nvalid];
  if (lane < s.nvalid) {   // lane l stages entry pbase + l
    const int p = s.pbase + lane;
    s.my_b = ord ? (int)ord[p] : p;
    s.my_seg = segid[p];
    s.my_oid = oid[s.my_b];
    s.my_es = a.entry_seg[s.my_b];
  }
  s.seg_before = __shfl_sync(0xffffffffu, edge, 0);
  s.seg_after = __shfl_sync(0xffffffffu, edge, 1);
  return s;
}

// kNc: the other side's rows go through the read-only path (__ldg); false in the persistent epoch kernel, where the tables
// change during the launch
template <int kRowVecs, bool kNc = true>
__device__ __forceinline__ void run_chunk(const GradArgs &a, const ChunkStage &s) {
  if (s.nvalid == 0) return;
  const int lane = threadIdx.x & 31;
  const bool user_side = s.user_side;
  const int c = s.c, nvalid = s.nvalid;
  const int B = FR_B(a.B, a.B_dev);
  const float *other = user_side ? a.I : a.U;
  float *gseg = user_side ? a.gseg_u : a.gseg_i;
  float *head = user_side ? a.head_u : a.head_i;
  float *tail = user_side ? a.tail_u : a.tail_i;
  const int d = a.d;
  const float vmin = ord2f(a.ctrl[a.pre_handover ? CTRL_MIN : CTRL_SAVED_MIN]);
  const int nB = a.norm_from_ctrl ? (int)a.ctrl[CTRL_NORM_B] : a.norm_B;
  const int my_seg = s.my_seg, my_oid = s.my_oid, my_b = s.my_b;
  const int seg_before = s.seg_before, seg_after = s.seg_after;
  float4 acc[kRowVecs];
#pragma unroll
  for (int v = 0; v < kRowVecs; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  int cur = __shfl_sync(0xffffffffu, my_seg, 0);

  auto flush = [&](int sg) {
    float *dst = (sg != seg_before && sg != seg_after) ? gseg + (size_t)sg * d     // the segment lies inside the chunk
                 : (sg == seg_before)                   ? head + (size_t)c * d      // it started in an earlier chunk
                                                        : tail + (size_t)c * d;     // it continues into the next one
#pragma unroll
    for (int v = 0; v < kRowVecs; ++v) {
      const int k = lane * 4 + v * 128;
      if (k < d) {
        *(float4 *)(dst + k) = acc[v];
        acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  };

  constexpr int kDepth = kRowVecs == 1 ? 8 : 4;
  float4 x[kDepth][kRowVecs];
  // issue the row loads of kDepth entries before consuming them (memory-level parallelism)
  auto issue = [&](int l0) {
#pragma unroll
    for (int e = 0; e < kDepth; ++e) {
      const int l = l0 + e;
      const int o = __shfl_sync(0xffffffffu, my_oid, l & 31);
      const float4 *row = (const float4 *)(other + (size_t)o * d);
#pragma unroll
      for (int v = 0; v < kRowVecs; ++v) {
        const int k = lane * 4 + v * 128;
        x[e][v] = (l < nvalid && k < d) ? (kNc ? __ldg(row + (k >> 2)) : __ldcg(row + (k >> 2)))
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  };
  // the first rows need the other side's ids only: they are in flight while the entry's dL/dpred coefficient (the
  // entry's columns and its segment's fairness term) is formed
  issue(0);
  float my_coef = 0.f;
  if (lane < nvalid) {
    const int g = a.sst[my_b] != vmin;
    my_coef = (2.f * (a.pred[my_b] - a.rating[my_b]) / (float)(nB > 0 ? nB : B) + a.cseg[2 * s.my_es + g] +
               a.cglob[g]) * a.grad_scale;
  }
  for (int l0 = 0; l0 < nvalid; l0 += kDepth) {
    if (l0) issue(l0);
#pragma unroll
    for (int e = 0; e < kDepth; ++e) {
      const int l = l0 + e;
      const int sg = __shfl_sync(0xffffffffu, my_seg, l & 31);
      const float cf = __shfl_sync(0xffffffffu, my_coef, l & 31);
      if (l < nvalid) {
        if (sg != cur) {
          flush(cur);
          cur = sg;
        }
#pragma unroll
        for (int v = 0; v < kRowVecs; ++v) {
          acc[v].x = fmaf(cf, x[e][v].x, acc[v].x);
          acc[v].y = fmaf(cf, x[e][v].y, acc[v].y);
          acc[v].z = fmaf(cf, x[e][v].z, acc[v].z);
          acc[v].w = fmaf(cf, x[e][v].w, acc[v].w);
        }
      }
    }
  }
  flush(cur);
}

template <int kRowVecs, bool kNc = true>
__device__ __forceinline__ void grads_chunk(const GradArgs &a, int nchunk, int c) {   // c in [0, 2 * nchunk): one warp
  run_chunk<kRowVecs, kNc>(a, stage_chunk(a, nchunk, c));
}

template <int kRowVecs>
__global__ void __launch_bounds__(256, kRowVecs == 1 ? 4 : 2) k_segment_grads(GradArgs a, int nchunk) {
  grads_chunk<kRowVecs>(a, nchunk, (blockIdx.x * blockDim.x + threadIdx.x) >> 5);
}

// ------------------------------------------------------------------------------------------ apply
// One thread per float4 of the concatenated [U ; I] parameter space.  The row's gradient is read from the
// segment partials when the row was touched by this batch (row_tab stamp), else it is zero; then either
//   kAdamFused : torch Adam with L2 weight decay, in place (p, m, v)
//   kDenseOut  : write the dense gradient (compat path: autograd .grad of nn.Embedding)
//   kAdamDense : Adam from a caller-provided dense gradient
enum ApplyMode { kAdamFused = 0, kDenseOut = 1, kAdamDense = 2 };

struct ApplyArgs {
  float *U, *I, *mU, *vU, *mI, *vI, *dU, *dI;
  int n_users, n_items, d;
  const uint2 *row_tab_u, *row_tab_i;
  const int32_t *segoff_u, *segoff_i;
  const float *gseg_u, *head_u, *tail_u, *gseg_i, *head_i, *tail_i;
  const uint32_t *ctrl;
  int step;
  double lr, beta1, beta2, eps, wd;
  int chunk;   // the gradient kernel's chunk size (head/tail partial indexing)
  int pre_handover;   // fused step: CTRL_STAMP / CTRL_ADAM_T still hold the values from before this batch's hand-over
};

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

__device__ __forceinline__ float adam1(float &p, float &m, float &v, float g, float wd, float w1, float b2, float w2,
                                       float bc2s, float eps, float neg_step) {
  g = fmaf(wd, p, g);                       // grad.add(param, alpha=weight_decay)
  m = fmaf(w1, g - m, m);                   // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(w2 * g, g, v * b2);              // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
  const float denom = sqrtf(v) / bc2s + eps;
  p = fmaf(neg_step, m / denom, p);         // param.addcdiv_(exp_avg, denom, value=-step_size)
  return p;
}

template <int kMode>
__device__ __forceinline__ void apply_body(const ApplyArgs &a, float *sc) {
  if (kMode != kDenseOut) {
    if (threadIdx.x == 0) {
      const double t = a.step > 0 ? (double)a.step
                                  : (double)(a.ctrl[CTRL_ADAM_T] + (a.pre_handover ? a.ctrl[CTRL_STRIDE] : 0u));
      const double bc1 = 1.0 - pow(a.beta1, t);
      const double bc2 = 1.0 - pow(a.beta2, t);
      sc[0] = (float)(-a.lr / bc1);
      sc[1] = (float)sqrt(bc2);
    }
    __syncthreads();
  }
  const float neg_step = sc[0], bc2s = sc[1];
  // every scalar is formed in double and rounded once, like the Python floats torch hands to its kernels
  const float w1 = (float)(1.0 - a.beta1), w2 = (float)(1.0 - a.beta2), b2 = (float)a.beta2, wd = (float)a.wd,
              eps = (float)a.eps;
  const int dq = a.d >> 2;
  const size_t nq_u = (size_t)a.n_users * dq, nq = nq_u + (size_t)a.n_items * dq;
  const uint32_t stamp = a.ctrl[CTRL_STAMP] - (a.pre_handover ? 0u : 1u);  // k_segment_loss already advanced it
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (size_t)gridDim.x * blockDim.x) {
    const bool is_item = q >= nq_u;
    const size_t ql = is_item ? q - nq_u : q;
    const int row = (int)(ql / dq), k = (int)(ql % dq) * 4;
    // the element's parameter and moments are requested FIRST: they do not depend on the row stamp, and behind the
    // stamp-dependent branch they were a second DRAM round trip per iteration (k_apply at 85 % of the copy peak)
    float4 *pp = nullptr, *pm = nullptr, *pv = nullptr;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f), m = p, v = p;
    if (kMode != kDenseOut) {
      pp = (float4 *)((is_item ? a.I : a.U) + ql * 4);
      pm = (float4 *)((is_item ? a.mI : a.mU) + ql * 4);
      pv = (float4 *)((is_item ? a.vI : a.vU) + ql * 4);
      p = *pp;
      m = ldg_stream(pm);
      v = ldg_stream(pv);
    }
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kMode == kAdamDense) {
      g = *(const float4 *)((is_item ? a.dI : a.dU) + ql * 4);
    } else {
      const uint2 t = (is_item ? a.row_tab_i : a.row_tab_u)[row];
      if (t.x == stamp) {
        const int32_t *segoff = is_item ? a.segoff_i : a.segoff_u;
        const float *gseg = is_item ? a.gseg_i : a.gseg_u;
        const float *head = is_item ? a.head_i : a.head_u;
        const float *tail = is_item ? a.tail_i : a.tail_u;
        const int s = (int)t.y, s0 = segoff[s], s1 = segoff[s + 1];
        const int c0 = s0 / a.chunk, c1 = (s1 - 1) / a.chunk;
        if (c0 == c1) {
          g = *(const float4 *)(gseg + (size_t)s * a.d + k);
        } else {
          g = *(const float4 *)(tail + (size_t)c0 * a.d + k);
#pragma unroll 8
          for (int c = c0 + 1; c <= c1; ++c) g = f4_add(g, __ldg((const float4 *)(head + (size_t)c * a.d + k)));
        }
      }
    }
    if (kMode == kDenseOut) {
      *(float4 *)((is_item ? a.dI : a.dU) + ql * 4) = g;
    } else {
      adam1(p.x, m.x, v.x, g.x, wd, w1, b2, w2, bc2s, eps, neg_step);
      adam1(p.y, m.y, v.y, g.y, wd, w1, b2, w2, bc2s, eps, neg_step);
      adam1(p.z, m.z, v.z, g.z, wd, w1, b2, w2, bc2s, eps, neg_step);
      adam1(p.w, m.w, v.w, g.w, wd, w1, b2, w2, bc2s, eps, neg_step);
      *pp = p;
      stg_stream(pm, m);
      stg_stream(pv, v);
    }
  }
}

template <int kMode>
__global__ void __launch_bounds__(256) k_apply(ApplyArgs a) {
  __shared__ float sc[3];
  apply_body<kMode>(a, sc);
}

// ------------------------------------------------------------------------------------------ lazy-exact Adam
// torch.optim.Adam with L2 weight decay moves EVERY row every step: a row the batch does not touch still sees the gradient
// wd * p and the decay of its moments (trainer.py:139).  That update is a pure per-element recurrence in
// (p, m, v, t), so it need not be applied when it happens: `last[row]` remembers the step a row is current for, and when a
// batch next touches the row (or at a flush) the missed steps are replayed in registers with the SAME float32 operations
// (adam1 with a zero data gradient, per-step scalars from the table k_adam_scalars fills with the expressions of
// apply_body) -- bit-identical to the dense sweep, without streaming 24 bytes per parameter per step through HBM.
struct LazyArgs {
  float *P, *M, *V;          // one table and its moments
  const uint32_t *last;      // [rows] step each row is current for
  const float *sc;           // scalars table: sc[2t] = -lr / (1 - beta1^t), sc[2t+1] = sqrt(1 - beta2^t)
  int d, step;
  double beta1, beta2, eps, wd;
  // touched rows = the segments of the sorted batch keys (single-GPU step, user side of the sharded step) ...
  const uint32_t *skey;      // row of segment s = skey[segoff[s]]
  const int32_t *segoff, *nseg;
  const float *gseg, *head, *tail;
  int chunk;
  // ... or the drawn items of a row-sharded step that THIS rank owns (items[e] % world == rank; local row = items[e] /
  // world), with the gradient of entry e = sum over the `world` slots (slot order) of slot_grad[k][slots[e]]
  const int32_t *items, *slots;
  int n_items_listed, rank, world;
  const float *slot_grad[8];
  // flush: every row of the table
  int64_t all_rows;
  int upto;                  // catch-up passes: replay up to and including this step
};

__device__ __forceinline__ void lazy_replay(float4 &p, float4 &m, float4 &v, uint32_t from, uint32_t to, const float *sc,
                                            float wd, float w1, float b2, float w2, float eps) {
  for (uint32_t t = from; t <= to; ++t) {   // steps from..to with a zero data gradient
    const float neg_step = __ldg(sc + 2 * t), bc2s = __ldg(sc + 2 * t + 1);
    adam1(p.x, m.x, v.x, 0.f, wd, w1, b2, w2, bc2s, eps, neg_step);
    adam1(p.y, m.y, v.y, 0.f, wd, w1, b2, w2, bc2s, eps, neg_step);
    adam1(p.z, m.z, v.z, 0.f, wd, w1, b2, w2, bc2s, eps, neg_step);
    adam1(p.w, m.w, v.w, 0.f, wd, w1, b2, w2, bc2s, eps, neg_step);
  }
}

// kRows: 0 = the rows of the batch's sorted segments, 1 = the rows of a listed item set that this rank owns, 2 = all rows
// kGrad: true  = optimizer step `step` with the row's data gradient (after replaying whatever is still pending before it)
//        false = catch-up only: replay the pending steps up to and including `upto`
// A training step runs catch-up (upto = step - 1) on the touched rows BEFORE the forward reads them, then the kGrad pass.
template <int kRows, bool kGrad>
__global__ void __launch_bounds__(256) k_adam_lazy(LazyArgs a) {
  const float w1 = (float)(1.0 - a.beta1), w2 = (float)(1.0 - a.beta2), b2 = (float)a.beta2, wd = (float)a.wd,
              eps = (float)a.eps;
  const int dq = a.d >> 2;
  int64_t n;
  if (kRows == 0) n = (int64_t)(*a.nseg) * dq;
  else if (kRows == 1) n = (int64_t)a.n_items_listed * dq;
  else n = a.all_rows * dq;
  const uint32_t upto = kGrad ? (uint32_t)a.step - 1u : (uint32_t)a.upto;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = q / dq;
    const int k = (int)(q % dq) * 4;
    int64_t row;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kRows == 0) {
      const int s = (int)e, s0 = a.segoff[s];
      row = (int64_t)a.skey[s0];
      if (kGrad) {
        const int s1 = a.segoff[s + 1];
        const int c0 = s0 / a.chunk, c1 = (s1 - 1) / a.chunk;
        if (c0 == c1) {
          g = *(const float4 *)(a.gseg + (size_t)s * a.d + k);
        } else {
          g = *(const float4 *)(a.tail + (size_t)c0 * a.d + k);
#pragma unroll 8
          for (int c = c0 + 1; c <= c1; ++c) g = f4_add(g, __ldg((const float4 *)(a.head + (size_t)c * a.d + k)));
        }
      }
    } else if (kRows == 1) {
      const int it = a.items[e];
      if (it % a.world != a.rank) continue;
      row = it / a.world;
      if (kGrad) {
        const size_t slot = (size_t)a.slots[e];
        for (int r = 0; r < a.world; ++r) g = f4_add(g, *(const float4 *)(a.slot_grad[r] + slot * a.d + k));
      }
    } else {
      row = e;
    }
    const uint32_t L = a.last[row];
    if (!kGrad && L >= upto) continue;
    float4 *pp = (float4 *)(a.P + (size_t)row * a.d + k);
    float4 *pm = (float4 *)(a.M + (size_t)row * a.d + k);
    float4 *pv = (float4 *)(a.V + (size_t)row * a.d + k);
    float4 p = *pp, m = *pm, v = *pv;
    lazy_replay(p, m, v, L + 1u, upto, a.sc, wd, w1, b2, w2, eps);
    if (kGrad) {
      const float neg_step = __ldg(a.sc + 2 * a.step), bc2s = __ldg(a.sc + 2 * a.step + 1);
      adam1(p.x, m.x, v.x, g.x, wd, w1, b2, w2, bc2s, eps, neg_step);
      adam1(p.y, m.y, v.y, g.y, wd, w1, b2, w2, bc2s, eps, neg_step);
      adam1(p.z, m.z, v.z, g.z, wd, w1, b2, w2, bc2s, eps, neg_step);
      adam1(p.w, m.w, v.w, g.w, wd, w1, b2, w2, bc2s, eps, neg_step);
    }
    *pp = p;
    *pm = m;
    *pv = v;
  }
}

// after k_adam_lazy (a separate launch: every thread of a row must have read last[row] before it changes)
template <int kRows, bool kGrad>
__global__ void __launch_bounds__(256) k_adam_lazy_mark(LazyArgs a, uint32_t *last) {
  int64_t n;
  if (kRows == 0) n = *a.nseg;
  else if (kRows == 1) n = a.n_items_listed;
  else n = a.all_rows;
  const uint32_t now = kGrad ? (uint32_t)a.step : (uint32_t)a.upto;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t row;
    if (kRows == 0) row = (int64_t)a.skey[a.segoff[e]];
    else if (kRows == 2) row = e;
    else {
      const int it = a.items[e];
      if (it % a.world != a.rank) continue;
      row = it / a.world;
    }
    if (last[row] < now) last[row] = now;
  }
}

// sc[2t], sc[2t+1] for t in [t0, t1]: the expressions of apply_body (double, rounded once)
static __global__ void k_adam_scalars(float *sc, int t0, int t1, double lr, double beta1, double beta2) {
  for (int t = t0 + blockIdx.x * blockDim.x + threadIdx.x; t <= t1; t += gridDim.x * blockDim.x) {
    const double bc1 = 1.0 - pow(beta1, (double)t);
    const double bc2 = 1.0 - pow(beta2, (double)t);
    sc[2 * t] = (float)(-lr / bc1);
    sc[2 * t + 1] = (float)sqrt(bc2);
  }
}

// ---- host helpers of the lazy-exact mode (shared by focf_train.cu and focf_shard.cu)
struct LazyCommon {
  int d, step;
  double lr, beta1, beta2, eps, wd;
  float *sc;
  int sc_cap;
  int32_t *sc_filled;   // HOST counter
};

static inline int lazy_fill_scalars(const LazyCommon &c, cudaStream_t st, const char *who) {
  FR_REQUIRE(c.sc && c.sc_filled && c.step >= 1, "%s: lazy_exact needs adam_scalars, scalars_filled and step >= 1", who);
  FR_REQUIRE(c.step < c.sc_cap, "%s: optimizer step %d exceeds the adam_scalars table (%d steps)", who, c.step, c.sc_cap);
  if (*c.sc_filled < c.step) {
    const int t0 = *c.sc_filled + 1, t1 = c.step;
    FR_LAUNCH(k_adam_scalars, grid_for(t1 - t0 + 1, 128, 64), 128, 0, st, c.sc, t0, t1, c.lr, c.beta1, c.beta2);
    *c.sc_filled = c.step;
  }
  return FR_OK;
}

static inline LazyArgs lazy_base(const LazyCommon &c, float *P, float *M, float *V, const uint32_t *last) {
  LazyArgs a{};
  a.P = P; a.M = M; a.V = V; a.last = last; a.sc = c.sc; a.d = c.d; a.step = c.step;
  a.beta1 = c.beta1; a.beta2 = c.beta2; a.eps = c.eps; a.wd = c.wd;
  return a;
}

// touched rows of one side given as sorted segments (at most max_seg of them; the count is device resident):
// catch-up to step - 1 (before the forward reads the rows) ...
static inline void lazy_catchup_segments(LazyArgs a, uint32_t *last, int64_t max_seg, cudaStream_t st) {
  a.upto = a.step - 1;
  if (a.upto < 1) return;
  FR_LAUNCH((k_adam_lazy<0, false>), grid_for(max_seg * (a.d >> 2), 256, kSMs * 32), 256, 0, st, a);
  FR_LAUNCH((k_adam_lazy_mark<0, false>), grid_for(max_seg, 256, kSMs * 8), 256, 0, st, a, last);
}
// ... and optimizer step `step` with the gradient partials
static inline void lazy_apply_segments(LazyArgs a, uint32_t *last, int64_t max_seg, cudaStream_t st) {
  FR_LAUNCH((k_adam_lazy<0, true>), grid_for(max_seg * (a.d >> 2), 256, kSMs * 32), 256, 0, st, a);
  FR_LAUNCH((k_adam_lazy_mark<0, true>), grid_for(max_seg, 256, kSMs * 8), 256, 0, st, a, last);
}

// owned rows of a listed item set: optimizer step with slot gradients (kGrad) or catch-up to a.upto
template <bool kGrad>
static inline void lazy_listed_items(LazyArgs a, uint32_t *last, cudaStream_t st) {
  if (a.n_items_listed <= 0) return;
  FR_LAUNCH((k_adam_lazy<1, kGrad>), grid_for((int64_t)a.n_items_listed * (a.d >> 2), 256, kSMs * 32), 256, 0, st, a);
  FR_LAUNCH((k_adam_lazy_mark<1, kGrad>), grid_for(a.n_items_listed, 256, kSMs * 8), 256, 0, st, a, last);
}

// every row up to and including step a.step
static inline void lazy_flush_table(LazyArgs a, uint32_t *last, int64_t rows, cudaStream_t st) {
  if (rows <= 0) return;
  a.all_rows = rows;
  a.upto = a.step;
  FR_LAUNCH((k_adam_lazy<2, false>), grid_for(rows * (a.d >> 2), 256, kSMs * 32), 256, 0, st, a);
  FR_LAUNCH((k_adam_lazy_mark<2, false>), grid_for(rows, 256, kSMs * 8), 256, 0, st, a, last);
}

}  // namespace fr
