// FOCF training step for sm_100a: embedding gather + score, item x group fairness regulariser,
// deterministic sorted-segment embedding gradients, dense Adam fused with the gradient read-out.
//
// Reference being replaced (paths relative to the reference root):
//   recbole/model/fair_recommender/focf.py:136-143  forward            -> k_forward
//   focf.py:75-134 get_item_ratings + *_unfairness, 152-169 calculate_loss -> k_segment_loss
//   autograd backward of the above + nn.Embedding dense backward (trainer.py:193) -> k_segment_grads
//   torch.optim.Adam(lr, weight_decay) dense step (trainer.py:139,196)  -> k_apply<...>
//
// Data layout in HBM: U [n_users,d], I [n_items,d] and the Adam moments are row-major float32; a batch is
// four SoA columns (uid, iid, rating, sst).  All of this is HBM/L2-bound gather/scatter + streaming work:
// rows move as 128-bit lane loads (a warp covers 512 contiguous bytes of a row per request), the item x
// group reduction is a warp-per-segment shuffle reduction, and gradients are reduced in sorted-segment
// order with a fixed tree, so results are bit-stable run to run (no floating-point atomics anywhere).
//
// Algorithmic HBM bytes: per interaction 16*d+16 (gather 2 rows fwd, gather 2 rows bwd,
// ids/rating/sst); per step 24*(n_users+n_items)*d for the dense Adam sweep (p,m,v read + write) --
// the dense gradient is never materialised in fr_focf_train_step.
#include <stdlib.h>

#include "sort.cuh"

#include "focf_device.cuh"

namespace fr {


__global__ void __launch_bounds__(256)
    k_forward(const float *__restrict__ U, const float *__restrict__ I, const int32_t *__restrict__ uid,
              const int32_t *__restrict__ iid, const float *__restrict__ sst, int B, const int32_t *__restrict__ B_dev,
              int d, float *__restrict__ pred, uint32_t *__restrict__ ctrl) {
  forward_body(U, I, uid, iid, sst, FR_B(B, B_dev), d, pred, ctrl);
}


// Phase 2 of the loss: CTA b owns the item segments b, b + G, ...: cooperative record sum, fairness terms, additive
// backward terms cseg[j][g]; per-CTA partial sums (squared error, smooth-L1, nonparity group sums) in segment order; the
// LAST CTA to finish (self-resetting ticket) adds the partials in CTA order -> loss, flags, control-block hand-over.
// (The segment -> CTA map depends only on J, so the result is bit-stable run to run.)
constexpr int kSegMaxBlocks = 148 * 8;
__global__ void __launch_bounds__(kSegThreads) k_segment_reduce(LossArgs a, float *__restrict__ part /* [grid, 8] */) {
  __shared__ float sh[kSegThreads / 32][8];
  __shared__ bool is_last;
  const int B = FR_B(a.B, a.B_dev);
  const int J = *a.J;
  int nB = a.norm_B, nJ = a.norm_J;
  if (a.norm_dev) {
    const uint32_t k = a.ctrl[CTRL_CURSOR] % (uint32_t)a.loss_by_cursor;
    nB = a.norm_dev[2 * k];
    nJ = a.norm_dev[2 * k + 1];
  }
  const float Bn = (float)(nB > 0 ? nB : B), Jn = (float)(nJ > 0 ? nJ : J);
  float w_sq = 0.f, w_hx = 0.f, w_g0 = 0.f, w_g1 = 0.f, w_n0 = 0.f, w_n1 = 0.f;   // meaningful on thread 0
  for (int j = blockIdx.x; j < J; j += gridDim.x) {
    float v[7];
    segment_record_sum(a, j, v, sh);
    if (threadIdx.x == 0) {
      w_sq += v[6];
      float hx = 0.f, cs0 = 0.f, cs1 = 0.f;
      if (a.objective >= FR_OBJ_VALUE && a.objective <= FR_OBJ_OVER) {
        segment_terms(a.objective, a.fair_weight, Jn, v[0], v[1], v[2], v[3], v[4], v[5], hx, cs0, cs1);
      } else if (a.objective == FR_OBJ_NONPARITY) {
        w_g0 += v[0]; w_g1 += v[1]; w_n0 += v[4]; w_n1 += v[5];
      }
      w_hx += hx;
      a.cseg[2 * j] = cs0;
      a.cseg[2 * j + 1] = cs1;
    }
  }
  if (threadIdx.x == 0) {
    float *p = part + 8 * (size_t)blockIdx.x;
    p[0] = w_sq; p[1] = w_hx; p[2] = w_g0; p[3] = w_g1; p[4] = w_n0; p[5] = w_n1;
    __threadfence();
    is_last = atomicAdd(&a.ctrl[CTRL_TICKET], 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // the per-CTA partials in a FIXED order, whichever CTA ends up last: thread t adds partials t, t + T, ... (independent
  // loads: a serial walk by one thread costs one L2 round trip per partial -- 0.15 ms for 1184 CTAs), then the fixed
  // shuffle tree of warp_sum, then the warps in warp order
  float q[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (unsigned b = threadIdx.x; b < gridDim.x; b += kSegThreads) {
    const float4 x = __ldcg((const float4 *)(part + 8 * (size_t)b));
    const float2 y = __ldcg((const float2 *)(part + 8 * (size_t)b + 4));
    q[0] += x.x; q[1] += x.y; q[2] += x.z; q[3] += x.w; q[4] += y.x; q[5] += y.y;
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) q[k] = warp_sum(q[k]);
  __syncthreads();   // sh may still be read by segment_record_sum's last combine
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 6; ++k) sh[threadIdx.x >> 5][k] = q[k];
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  float sq = 0.f, hx = 0.f, g0 = 0.f, g1 = 0.f, n0 = 0.f, n1 = 0.f;
#pragma unroll
  for (int w = 0; w < kSegThreads / 32; ++w) {
    sq += sh[w][0]; hx += sh[w][1]; g0 += sh[w][2]; g1 += sh[w][3]; n0 += sh[w][4]; n1 += sh[w][5];
  }
  float loss = sq / Bn;                                                    // nn.MSELoss 'mean'
  float cg0 = 0.f, cg1 = 0.f;
  if (a.objective >= FR_OBJ_VALUE && a.objective <= FR_OBJ_OVER) {
    loss += a.fair_weight * (hx / Jn);                                     // focf.py:166
  } else if (a.objective == FR_OBJ_NONPARITY) {
    if (n1 == 0.f || n0 == 0.f) {
      atomicOr(a.flags, FR_FLAG_SINGLE_GROUP);
    } else {
      const float z = g0 / n0 - g1 / n1, x = fabsf(z);                     // focf.py:131-134
      loss += a.fair_weight * (x < 1.f ? 0.5f * x * x : x - 0.5f);
      const float hp = a.fair_weight * (x < 1.f ? z : (float)((z > 0.f) - (z < 0.f)));
      cg0 = hp / n0;
      cg1 = -hp / n1;
    }
  }
  a.cglob[0] = cg0;
  a.cglob[1] = cg1;
  a.loss[a.loss_by_cursor ? a.ctrl[CTRL_CURSOR] % (uint32_t)a.loss_by_cursor : 0u] = loss;
  if (loss != loss) atomicOr(a.flags, FR_FLAG_NAN_LOSS);
  // hand the group values to the backward kernels, re-arm the control block for the next batch
  a.ctrl[CTRL_SAVED_MIN] = a.ctrl[CTRL_MIN];
  a.ctrl[CTRL_SAVED_MAX] = a.ctrl[CTRL_MAX];
  a.ctrl[CTRL_NORM_B] = (uint32_t)(nB > 0 ? nB : 0);
  a.ctrl[CTRL_NORM_J] = (uint32_t)(nJ > 0 ? nJ : 0);
  a.ctrl[CTRL_MIN] = 0xffffffffu;
  a.ctrl[CTRL_MAX] = 0u;
  a.ctrl[CTRL_TICKET] = 0u;
  a.ctrl[CTRL_STAMP] += 1u;
  a.ctrl[CTRL_CURSOR] += a.ctrl[CTRL_STRIDE];
  if (a.advance_adam) a.ctrl[CTRL_ADAM_T] += a.ctrl[CTRL_STRIDE];
}

// phase 1 (all CTAs): per-thread records over 8 sorted rows each
__global__ void __launch_bounds__(kLossThreads) k_segment_loss(LossArgs a) { loss_phase1(a, FR_B(a.B, a.B_dev)); }

// ------------------------------------------------------------------------------------------ fused step
// Batches at the ML-1M shape are a few thousand rows against ~15 MB of state: every kernel of the step is launch- /
// latency-bound, so forward -> loss -> gradients -> Adam run as ONE cooperative launch (one CTA per SM) separated by
// grid-wide barriers instead of four dependent launches.  The barrier is a monotonically increasing arrival counter in
// the control block (cooperative launch guarantees that all CTAs are co-resident, so spinning cannot deadlock).
__device__ __forceinline__ void grid_barrier(unsigned long long *bar) {
  __syncthreads();
  if (threadIdx.x == 0) {
    // release / acquire at GPU scope (not __threadfence() = fence.sc: ~140 CTAs issuing it at once serialise)
    const unsigned long long n = gridDim.x;
    unsigned long long old, v;
    asm volatile("atom.add.release.gpu.global.u64 %0, [%1], %2;" : "=l"(old) : "l"(bar), "l"(1ull) : "memory");
    const unsigned long long target = (old / n + 1ull) * n;   // 64-bit: never wraps
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(bar) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

struct FusedArgs {
  const float *U, *I;
  const int32_t *uid, *iid;
  const float *sst;
  int B, d;
  const int32_t *B_dev;
  float *pred;
  LossArgs loss;
  GradArgs grad;
  ApplyArgs apply;
  int cap;          // rows the shared-memory staging is sized for (>= any batch)
  uint32_t *ctrl;
  unsigned long long *trace;   // optional (FR_FOCF_TRACE=1): [gridDim.x][8] %globaltimer stamps at the phase boundaries
};
__device__ __forceinline__ void fused_stamp(unsigned long long *trace, int slot) {
  if (trace && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    trace[blockIdx.x * 8 + slot] = t;
  }
}
constexpr int kFusedThreads = 512;
static size_t fused_smem_bytes(int cap) { return ((size_t)5 * cap + 8) * sizeof(float); }

// forward | barrier | batch statistics (per CTA, shared memory) + gradients | barrier | Adam | barrier | hand-over
template <int kRowVecs>
__global__ void __launch_bounds__(kFusedThreads, 1) k_focf_fused_step(FusedArgs f) {
  extern __shared__ __align__(16) float fused_sm[];
  __shared__ float sh[33];
  __shared__ float sc[3];
  float *s_cseg = fused_sm + 3 * f.cap, *s_cglob = fused_sm + 5 * f.cap;
  unsigned long long *bar = (unsigned long long *)(f.ctrl + CTRL_GRID_BAR);
  const int B = FR_B(f.B, f.B_dev);
  fused_stamp(f.trace, 0);
  forward_body(f.U, f.I, f.uid, f.iid, f.sst, B, f.d, f.pred, f.ctrl);
  fused_stamp(f.trace, 1);
  grid_barrier(bar);
  fused_stamp(f.trace, 2);
  const float loss = fused_stats(f.loss, B, f.cap, fused_sm, sh, s_cseg, s_cglob);
  fused_stamp(f.trace, 3);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const LossArgs &a = f.loss;
    a.loss[a.loss_by_cursor ? a.ctrl[CTRL_CURSOR] % (uint32_t)a.loss_by_cursor : 0u] = loss;
    if (loss != loss) atomicOr(a.flags, FR_FLAG_NAN_LOSS);
  }
  {
    GradArgs g = f.grad;
    g.cseg = s_cseg;
    g.cglob = s_cglob;
    const int nchunk = (B + g.chunk - 1) / g.chunk;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int c = warp; c < 2 * nchunk; c += nwarps) grads_chunk<kRowVecs>(g, nchunk, c);
  }
  fused_stamp(f.trace, 4);
  grid_barrier(bar);
  fused_stamp(f.trace, 5);
  apply_body<kAdamFused>(f.apply, sc);
  fused_stamp(f.trace, 6);
  grid_barrier(bar);
  fused_stamp(f.trace, 7);
  if (blockIdx.x == 0 && threadIdx.x == 0) {   // hand the control block over to the next batch (cf. loss_phase2)
    uint32_t *ctrl = f.ctrl;
    ctrl[CTRL_SAVED_MIN] = ctrl[CTRL_MIN];
    ctrl[CTRL_SAVED_MAX] = ctrl[CTRL_MAX];
    ctrl[CTRL_MIN] = 0xffffffffu;
    ctrl[CTRL_MAX] = 0u;
    ctrl[CTRL_STAMP] += 1u;
    ctrl[CTRL_CURSOR] += ctrl[CTRL_STRIDE];
    if (f.loss.advance_adam) ctrl[CTRL_ADAM_T] += ctrl[CTRL_STRIDE];
  }
}

// ------------------------------------------------------------------------------------------ pair scores
// focf.py:145-150 predict / collector.py:179 rec.positive_score: k-ascending fmaf chain per pair, the
// exact arithmetic of the scorer (bit-identical to oracle/c/fairrec_oracle.c).
__device__ __forceinline__ float apply_transform(float x, int transform, float max_rating) {
  if (transform == FR_TRANSFORM_CLAMP_DIV) {
    x = fminf(fmaxf(x, 0.f), max_rating);
    return __fdiv_rn(x, max_rating);
  }
  if (transform == FR_TRANSFORM_SIGMOID) return 1.f / (1.f + expf(-x));
  return x;
}

__global__ void __launch_bounds__(256)
    k_pair_scores(const float *__restrict__ U, const float *__restrict__ I, const int32_t *__restrict__ uid,
                  const int32_t *__restrict__ iid, int64_t n, int d, int transform, float max_rating,
                  float *__restrict__ out) {
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n; b += (int64_t)gridDim.x * blockDim.x) {
    const float4 *pu = (const float4 *)(U + (size_t)uid[b] * d);
    const float4 *pi = (const float4 *)(I + (size_t)iid[b] * d);
    float acc = 0.f;
    for (int k = 0; k < (d >> 2); ++k) {
      const float4 x = __ldg(pu + k), y = __ldg(pi + k);
      acc = fmaf(x.x, y.x, acc);
      acc = fmaf(x.y, y.y, acc);
      acc = fmaf(x.z, y.z, acc);
      acc = fmaf(x.w, y.w, acc);
    }
    out[b] = apply_transform(acc, transform, max_rating);
  }
}


// ------------------------------------------------------------------------------------------ batch gather
// focf_dataloader.py:37-50 + dataset.py __getitem__/join: a FOCF batch is the concatenation of ALL train rows
// of the drawn items.  The train split lives on the device sorted by item (CSC: item_off); one warp copies
// one drawn item's segment (coalesced) and joins the user's sensitive attribute.
// plan_desc == nullptr: (draw_items, draw_off, J) describe the batch directly.
// plan_desc != nullptr: planned epoch -- descriptor row ctrl[CTRL_CURSOR] % plan_len = {item_pos, off_pos, J, B}
// indexes into draw_items / draw_off; the kernel also publishes B in ctrl[CTRL_B] for the rest of the step.
__global__ void __launch_bounds__(256)
    k_gather_batch(const int32_t *__restrict__ item_off, const int32_t *__restrict__ train_uid,
                   const float *__restrict__ train_rating, const float *__restrict__ sst_of_user,
                   const int32_t *__restrict__ draw_items, const int32_t *__restrict__ draw_off, int J,
                   const int32_t *__restrict__ plan_desc, int plan_len, uint32_t *__restrict__ ctrl,
                   int32_t *__restrict__ uid, int32_t *__restrict__ iid, float *__restrict__ rating,
                   float *__restrict__ sst) {
  if (plan_desc) {
    const int32_t *dsc = plan_desc + 4 * (int)(ctrl[CTRL_CURSOR] % (uint32_t)plan_len);
    draw_items += dsc[0];
    draw_off += dsc[1];
    J = dsc[2];
    if (blockIdx.x == 0 && threadIdx.x == 0) ctrl[CTRL_B] = (uint32_t)dsc[3];
  }
  // one thread per output row (work independent of item popularity): the row's drawn item is found by a binary
  // search in the J+1 batch offsets (L1-resident), then the CSC row is copied and the user's attribute joined
  const int B = draw_off[J];
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < B; p += gridDim.x * blockDim.x) {
    int lo = 0, hi = J;  // invariant: draw_off[lo] <= p < draw_off[hi]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (draw_off[mid] <= p) lo = mid; else hi = mid;
    }
    const int it = draw_items[lo];
    const int src = item_off[it] + (p - draw_off[lo]);
    const int u = train_uid[src];
    uid[p] = u;
    iid[p] = it;
    rating[p] = train_rating[src];
    sst[p] = sst_of_user[u];
  }
}

// ------------------------------------------------------------------------------------------ fused preparation
// Batches of up to 8192 rows (every FOCFDataLoader batch at the ML-1M shape): ONE launch of two CTAs does what the
// general path needs ~12 launches for -- CTA 0 the item side, CTA 1 the user side: load keys to shared memory,
// stable LSD radix sort entirely in shared memory (ping-pong buffers, per-warp __match_any_sync ranking), segment
// heads + block scan, and the write-out of sorted keys / order / segment ids / offsets / row stamps.
constexpr int kPsThreads = 1024, kPsItems = 8, kPsMax = kPsThreads * kPsItems, kPsWarps = kPsThreads / 32;

struct PrepSide {
  const int32_t *keys_in;
  uint32_t *skey, *ord;
  int32_t *segid, *segoff, *count;
  uint2 *row_tab;
  int32_t *entry_seg;   // item side only
  int key_bits, do_sort;
};

__global__ void __launch_bounds__(kPsThreads)
    k_prepare_small(PrepSide item, PrepSide user, int B_host, const int32_t *__restrict__ B_dev,
                    const uint32_t *__restrict__ ctrl) {
  extern __shared__ __align__(16) uint32_t ps_smem[];
  uint32_t *kA = ps_smem, *vA = kA + kPsMax, *kB = vA + kPsMax, *vB = kB + kPsMax;
  uint16_t *cnt = (uint16_t *)(vB + kPsMax);           // [kPsWarps][256]
  uint32_t *dbase = (uint32_t *)(cnt + kPsWarps * 256);  // [256]
  uint32_t *wsum = dbase + 256;                          // [32]
  const PrepSide sd = blockIdx.x == 0 ? item : user;
  const int n = FR_B(B_host, B_dev);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t stamp = ctrl[CTRL_STAMP];
  // radix ranking: warp w owns the contiguous keys [w*chunk, (w+1)*chunk), chunk a multiple of 32 chosen so that
  // all 32 warps share the n keys evenly (rounds = chunk/32 <= 8)
  const int rounds = (n + kPsThreads - 1) / kPsThreads, chunk = rounds * 32;

  for (int p = tid; p < n; p += kPsThreads) {
    kA[p] = (uint32_t)sd.keys_in[p];
    vA[p] = (uint32_t)p;
  }
  __syncthreads();

  if (sd.do_sort) {
    const int passes = (sd.key_bits + 7) / 8;
    for (int pass = 0; pass < passes; ++pass) {
      const int shift = 8 * pass;
      for (int i = tid; i < kPsWarps * 256; i += kPsThreads) cnt[i] = 0;
      __syncthreads();
      uint32_t key[kPsItems], val[kPsItems], rnk[kPsItems];
#pragma unroll
      for (int r = 0; r < kPsItems; ++r) {
        if (r >= rounds) break;
        const int p = w * chunk + r * 32 + lane;
        const bool valid = p < n;
        key[r] = valid ? kA[p] : 0u;
        val[r] = valid ? vA[p] : 0u;
        const uint32_t dig = valid ? ((key[r] >> shift) & 255u) : 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, dig);
        const uint32_t before = valid ? cnt[w * 256 + dig] : 0u;
        __syncwarp();
        if (valid && lane == (__ffs(peers) - 1)) cnt[w * 256 + dig] = (uint16_t)(before + __popc(peers));
        __syncwarp();
        rnk[r] = before + __popc(peers & ((1u << lane) - 1u));
      }
      __syncthreads();
      uint32_t tot = 0;
      if (tid < 256) {  // digit tid: exclusive prefix over the warps
#pragma unroll 8
        for (int ww = 0; ww < kPsWarps; ++ww) {
          const uint32_t t = cnt[ww * 256 + tid];
          cnt[ww * 256 + tid] = (uint16_t)tot;
          tot += t;
        }
        uint32_t inc = tot;  // exclusive scan of the 256 digit totals (8 warps)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[w] = inc;
        dbase[tid] = inc - tot;
      }
      __syncthreads();
      if (tid < 256) {
        uint32_t base = 0;
        for (int ww = 0; ww < w; ++ww) base += wsum[ww];
        dbase[tid] += base;
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < kPsItems; ++r) {
        if (r >= rounds) break;
        const int p = w * chunk + r * 32 + lane;
        if (p < n) {
          const uint32_t dig = (key[r] >> shift) & 255u;
          const uint32_t pos = dbase[dig] + cnt[w * 256 + dig] + rnk[r];
          kB[pos] = key[r];
          vB[pos] = val[r];
        }
      }
      __syncthreads();
      uint32_t *t0 = kA; kA = kB; kB = t0;
      t0 = vA; vA = vB; vB = t0;
    }
  }

  // ---- segments: thread tid owns the 8 consecutive sorted positions tid*8 .. tid*8+7
  const int p0 = tid * kPsItems;
  bool head[kPsItems];
  int c = 0;
#pragma unroll
  for (int i = 0; i < kPsItems; ++i) {
    const int p = p0 + i;
    head[i] = p < n && (p == 0 || kA[p] != kA[p - 1]);
    c += head[i];
  }
  int inc = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // wsum reuse
  if (lane == 31) wsum[w] = (uint32_t)inc;
  __syncthreads();
  int run = inc - c;
  for (int ww = 0; ww < w; ++ww) run += (int)wsum[ww];
#pragma unroll
  for (int i = 0; i < kPsItems; ++i) {
    const int p = p0 + i;
    if (p >= n) break;
    const uint32_t k = kA[p];
    if (head[i]) {
      sd.segoff[run] = p;
      sd.row_tab[k] = make_uint2(stamp, (uint32_t)run);
      ++run;
    }
    const int sgm = run - 1;
    sd.segid[p] = sgm;
    sd.skey[p] = k;
    sd.ord[p] = vA[p];
    if (sd.entry_seg) sd.entry_seg[vA[p]] = sgm;
    if (p == n - 1) {
      sd.segoff[sgm + 1] = n;
      sd.count[0] = sgm + 1;
    }
  }
}

__global__ void k_set_ctrl(uint32_t *ctrl, int cursor, int adam_t, int stride) {
  if (cursor >= 0) ctrl[CTRL_CURSOR] = (uint32_t)cursor;
  if (adam_t != INT32_MIN) ctrl[CTRL_ADAM_T] = (uint32_t)adam_t;   // -1 is a legal start with stride 2 (becomes 1 on the first step)
  if (stride >= 1) ctrl[CTRL_STRIDE] = (uint32_t)stride;
}

// ------------------------------------------------------------------------------------------ host side
static bool planned(const fr_focf_step *s) { return s->plan_desc != nullptr; }

static int check_step(const fr_focf_step *s, bool need_adam, const char *who) {
  FR_REQUIRE(s, "%s: null step", who);
  FR_REQUIRE(s->U && s->I && s->uid && s->iid && s->rating && s->sst && s->pred && s->loss && s->status_flags &&
                 s->workspace,
             "%s: null pointer in fr_focf_step", who);
  FR_REQUIRE(s->d >= 4 && s->d % 4 == 0 && s->d <= kMaxD, "%s: embedding size %d must be a multiple of 4 in [4,%d]",
             who, s->d, kMaxD);
  FR_REQUIRE(s->B >= 1 && s->n_users >= 1 && s->n_items >= 1, "%s: empty batch or table", who);
  FR_REQUIRE(s->objective >= FR_OBJ_NONE && s->objective <= FR_OBJ_NONPARITY, "%s: bad objective %d", who,
             s->objective);
  if (need_adam) FR_REQUIRE(s->mU && s->vU && s->mI && s->vI, "%s: Adam state missing", who);
  FR_REQUIRE(!(s->objective == FR_OBJ_NONPARITY && (s->norm_B > 0 || s->norm_J > 0 || s->norm_dev)),
             "%s: the nonparity objective needs batch-global group means and is not available data-parallel", who);
  if (planned(s)) {
    FR_REQUIRE(s->plan_items && s->plan_offs && s->plan_len >= 1 && s->item_off && s->train_uid && s->train_rating &&
                   s->sst_of_user,
               "%s: incomplete batch plan", who);
  }
  return FR_OK;
}

static int carve_checked(const fr_focf_step *s, FocfWs *w, const char *who) {
  Carver c(s->workspace, s->workspace_bytes);
  *w = carve(c, s->n_users, s->n_items, s->d, s->B);
  if (!c.ok()) {
    set_error("%s: workspace too small (%zu < %zu bytes)", who, s->workspace_bytes, c.off);
    return FR_ERR_WORKSPACE;
  }
  return FR_OK;
}

// device-resident batch size: the caller's pointer, or ctrl[CTRL_B] published by the planned gather
static const int32_t *dev_B(const fr_focf_step *s, const FocfWs &w) {
  return planned(s) ? (const int32_t *)(w.ctrl + CTRL_B) : s->B_dev;
}

// batch gather (planned) + sort / segment preparation of both sides
static int prepare_impl(const fr_focf_step *s, const FocfWs &w, cudaStream_t st) {
  const int B = s->B;  // upper bound when dev_B() is set
  const int32_t *Bd = dev_B(s, w);
  const bool contiguous = s->items_contiguous || planned(s);
  if (planned(s)) {
    FR_LAUNCH(k_gather_batch, grid_for((int64_t)B, 256, kSMs * 8), 256, 0, st, s->item_off, s->train_uid, s->train_rating,
              s->sst_of_user, s->plan_items, s->plan_offs, 0, s->plan_desc, s->plan_len, w.ctrl, (int32_t *)s->uid,
              (int32_t *)s->iid, (float *)s->rating, (float *)s->sst);
  }
  const uint32_t *ord_i = contiguous ? nullptr : w.ord_i;
  if (B <= kPsMax) {
    PrepSide pi{s->iid, w.skey_i, w.ord_i, w.segid_i, w.segoff_i, w.J, w.row_tab_i, w.entry_seg,
                bits_for((uint32_t)s->n_items), contiguous ? 0 : 1};
    PrepSide pu{s->uid, w.skey_u, w.ord_u, w.segid_u, w.segoff_u, w.Ju, w.row_tab_u, nullptr,
                bits_for((uint32_t)s->n_users), 1};
    const size_t smem = (size_t)kPsMax * 16 + (size_t)kPsWarps * 256 * 2 + 256 * 4 + 32 * 4;
    static bool attr_set = false;
    if (!attr_set) {
      FR_CUDA_OK(cudaFuncSetAttribute(k_prepare_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set = true;
    }
    FR_LAUNCH(k_prepare_small, 2, kPsThreads, smem, st, pi, pu, B, Bd, w.ctrl);
  } else {
    // item side: segments of equal item id (sorted first unless the loader guarantees adjacency)
    const uint32_t *skey_i = (const uint32_t *)s->iid;
    if (!contiguous) {
      sort_pairs((const uint32_t *)s->iid, nullptr, w.skey_i, w.ord_i, B, Bd, bits_for((uint32_t)s->n_items), w.sort, st);
      skey_i = w.skey_i;
    }
    build_segments(skey_i, ord_i, B, Bd, w.segid_i, w.segoff_i, w.J, w.row_tab_i, w.ctrl + CTRL_STAMP, w.entry_seg,
                   w.seg, st);
    // user side: always sorted (stable => batch order inside a user's segment, like index_add on CPU)
    sort_pairs((const uint32_t *)s->uid, nullptr, w.skey_u, w.ord_u, B, Bd, bits_for((uint32_t)s->n_users), w.sort, st);
    build_segments(w.skey_u, w.ord_u, B, Bd, w.segid_u, w.segoff_u, w.Ju, w.row_tab_u, w.ctrl + CTRL_STAMP, nullptr,
                   w.seg, st);
  }
  return FR_OK;
}

static LossArgs loss_args(const fr_focf_step *s, const FocfWs &w, bool advance_adam) {
  const bool contiguous = s->items_contiguous || planned(s);
  return LossArgs{s->pred, s->rating, s->sst, contiguous ? nullptr : w.ord_i, w.segid_i, w.segoff_i, w.J, s->B, dev_B(s, w),
                  planned(s) ? s->plan_len : 0, advance_adam ? 1 : 0, s->objective, s->norm_B, s->norm_J,
                  planned(s) ? s->norm_dev : nullptr, s->fair_weight,
                  w.cseg, w.rec_seg, w.rec_head, w.rec_tail, w.cglob, s->loss, w.ctrl, s->status_flags};
}

static int forward_only_impl(const fr_focf_step *s, const FocfWs &w, cudaStream_t st, bool advance_adam);
static int forward_impl(const fr_focf_step *s, const FocfWs &w, cudaStream_t st, bool advance_adam) {
  int rc = prepare_impl(s, w, st);
  if (rc) return rc;
  return forward_only_impl(s, w, st, advance_adam);
}

static int forward_only_impl(const fr_focf_step *s, const FocfWs &w, cudaStream_t st, bool advance_adam) {
  const int B = s->B;
  const int32_t *Bd = dev_B(s, w);
  FR_LAUNCH(k_forward, grid_for((int64_t)B, 32), 256, 0, st, s->U, s->I, s->uid, s->iid, s->sst, B, Bd, s->d, s->pred,
            w.ctrl);
  LossArgs la = loss_args(s, w, advance_adam);
  FR_LAUNCH(k_segment_loss, (B + kLossThreads * kLossRows - 1) / (kLossThreads * kLossRows), kLossThreads, 0, st, la);
  // J is device resident: the grid is sized for the batch (a segment holds >= 1 row), capped at 8 CTAs per SM
  FR_LAUNCH(k_segment_reduce, grid_for(B, 64, kSegMaxBlocks), kSegThreads, 0, st, la, w.red_part);
  return FR_OK;
}

static GradArgs grad_args(const fr_focf_step *s, const FocfWs &w, float grad_scale) {
  const uint32_t *ord_i = (s->items_contiguous || planned(s)) ? nullptr : w.ord_i;
  return GradArgs{s->U, s->I, s->uid, s->iid, s->rating, s->sst, s->pred, s->B, s->d, dev_B(s, w), s->norm_B,
                  (planned(s) && s->norm_dev) ? 1 : 0, ord_i, w.ord_u,
                  w.segid_i, w.segoff_i, w.segid_u, w.segoff_u, w.entry_seg, w.cseg, w.cglob, w.ctrl, grad_scale,
                  grad_chunk(s->B), w.gseg_i, w.head_i, w.tail_i, w.gseg_u, w.head_u, w.tail_u};
}

static void grads_impl(const fr_focf_step *s, const FocfWs &w, float grad_scale, cudaStream_t st) {
  GradArgs ga = grad_args(s, w, grad_scale);
  const int nchunk = (s->B + ga.chunk - 1) / ga.chunk;
  const int grid = (2 * nchunk + 7) / 8;
  if (s->d <= 128) {
    FR_LAUNCH(k_segment_grads<1>, grid, 256, 0, st, ga, nchunk);
  } else if (s->d <= 256) {
    FR_LAUNCH(k_segment_grads<2>, grid, 256, 0, st, ga, nchunk);
  } else {
    FR_LAUNCH(k_segment_grads<4>, grid, 256, 0, st, ga, nchunk);
  }
}

static ApplyArgs apply_args(const fr_focf_step *s, const FocfWs &w) {
  return ApplyArgs{s->U, s->I, s->mU, s->vU, s->mI, s->vI, s->dU, s->dI, s->n_users, s->n_items, s->d,
                   w.row_tab_u, w.row_tab_i, w.segoff_u, w.segoff_i, w.gseg_u, w.head_u, w.tail_u,
                   w.gseg_i, w.head_i, w.tail_i, w.ctrl, s->step, s->lr, s->beta1, s->beta2, s->eps, s->weight_decay,
                   grad_chunk(s->B)};
}

static int apply_grid(const fr_focf_step *s) {
  const int64_t nq = ((int64_t)s->n_users + s->n_items) * (s->d / 4);
  // one float4 per thread while the table is small (latency-bound: maximise loads in flight), grid-stride beyond
  return grid_for(nq, 256, kSMs * 16);
}

// the fused cooperative step serves the latency-bound regime: small batches (the single-launch preparation path),
// fused Adam, embedding rows of at most 128 floats
static bool fused_eligible(const fr_focf_step *s) {
  static int disabled = -1;
  if (disabled < 0) {
    const char *e = getenv("FR_FOCF_NO_FUSED_STEP");
    disabled = (e && e[0] == '1') ? 1 : 0;
  }
  return !disabled && !s->no_fused && s->adam_mode == FR_ADAM_DENSE_EXACT && s->B <= kPsMax && s->d <= 128 && s->norm_B == 0 && s->norm_J == 0 && !s->norm_dev &&
         fused_smem_bytes(s->B) <= 200 * 1024;
}

static unsigned long long *g_fused_trace = nullptr;
static int fused_step_impl(const fr_focf_step *s, const FocfWs &w, cudaStream_t st) {
  static int n_sm = 0, per_sm = -1;
  if (per_sm < 0) {
    int dev = 0;
    FR_CUDA_OK(cudaGetDevice(&dev));
    FR_CUDA_OK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    FR_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_focf_fused_step<1>, kFusedThreads, 0));
  }
  if (per_sm < 1 || n_sm < 1) {
    set_error("fr_focf_train_step: fused step kernel does not fit an SM");
    return FR_ERR_UNSUPPORTED;
  }
  const size_t smem = fused_smem_bytes(s->B);
  static size_t smem_set = 0;
  if (smem > smem_set) {
    FR_CUDA_OK(cudaFuncSetAttribute(k_focf_fused_step<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  GradArgs ga = grad_args(s, w, 1.0f);
  ga.pre_handover = 1;
  ApplyArgs aa = apply_args(s, w);
  aa.pre_handover = 1;
  static int trace_on = -1;
  if (trace_on < 0) {
    const char *e = getenv("FR_FOCF_TRACE");
    trace_on = (e && e[0] == '1') ? 1 : 0;
    if (trace_on) FR_CUDA_OK(cudaMalloc(&g_fused_trace, sizeof(unsigned long long) * 8 * 256));
  }
  FusedArgs f{s->U, s->I, s->uid, s->iid, s->sst, s->B, s->d, dev_B(s, w), s->pred, loss_args(s, w, s->step <= 0),
              ga, aa, s->B, w.ctrl, trace_on ? g_fused_trace : nullptr};
  void *args[] = {&f};
  const bool p = prof_on();
  if (p) prof_begin("k_focf_fused_step", st);
  // a few SMs are left to the preparation kernels of the NEXT batch (second stream, fr_focf_step_prepare)
  const int grid = n_sm > 16 ? n_sm - 4 : n_sm;
  cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_focf_fused_step<1>, dim3(grid), dim3(kFusedThreads), args, smem, st);
  if (p) prof_end(st);
  count_launch();
  if (e != cudaSuccess) {
    set_error("cudaLaunchCooperativeKernel(k_focf_fused_step) failed: %s", cudaGetErrorString(e));
    return FR_ERR_CUDA;
  }
  return FR_OK;
}


// lazy_exact: the two sides' touched rows as LazyArgs (segments of the sorted batch keys)
static int lazy_sides(const fr_focf_step *s, const FocfWs &w, cudaStream_t st, const char *who, LazyArgs *u, LazyArgs *i) {
  FR_REQUIRE(s->adam_mode == FR_ADAM_LAZY_EXACT, "%s: unknown adam_mode %d", who, s->adam_mode);
  FR_REQUIRE(s->last_step_u && s->last_step_i && !planned(s) && !s->B_dev,
             "%s: lazy_exact needs last_step_u / last_step_i and host-described batches", who);
  LazyCommon c{s->d, s->step, s->lr, s->beta1, s->beta2, s->eps, s->weight_decay, s->adam_scalars, s->scalars_cap,
               s->scalars_filled};
  int rc = lazy_fill_scalars(c, st, who);
  if (rc) return rc;
  const int chunk = grad_chunk(s->B);
  *u = lazy_base(c, s->U, s->mU, s->vU, s->last_step_u);
  u->skey = w.skey_u; u->segoff = w.segoff_u; u->nseg = w.Ju; u->gseg = w.gseg_u; u->head = w.head_u; u->tail = w.tail_u;
  u->chunk = chunk;
  *i = lazy_base(c, s->I, s->mI, s->vI, s->last_step_i);
  // item side: the keys are sorted into skey_i unless the loader guarantees adjacency (then the batch column itself is
  // sorted by run); the single-launch preparation of small batches always writes skey_i
  i->skey = ((s->items_contiguous || planned(s)) && s->B > kPsMax) ? (const uint32_t *)s->iid : w.skey_i;
  i->segoff = w.segoff_i; i->nseg = w.J; i->gseg = w.gseg_i; i->head = w.head_i; i->tail = w.tail_i; i->chunk = chunk;
  return FR_OK;
}

// lazy_exact, between preparation and forward: the rows this batch touches are brought up to step - 1 (the forward must
// read what the dense sweep would have left)
static int lazy_catchup(const fr_focf_step *s, const FocfWs &w, cudaStream_t st, const char *who) {
  if (s->adam_mode == FR_ADAM_DENSE_EXACT) return FR_OK;
  LazyArgs u, i;
  int rc = lazy_sides(s, w, st, who, &u, &i);
  if (rc) return rc;
  lazy_catchup_segments(u, s->last_step_u, s->B, st);
  lazy_catchup_segments(i, s->last_step_i, s->B, st);
  return FR_OK;
}

// the Adam stage of the general (multi-launch) step: dense sweep, or lazy-exact update of the touched rows only
static int adam_stage(const fr_focf_step *s, const FocfWs &w, cudaStream_t st, const char *who) {
  if (s->adam_mode == FR_ADAM_DENSE_EXACT) {
    // (an L2 prefetch of the next grid-stride iteration's lines was tried: 0.856 of the copy peak against 0.878-0.892 without)
    FR_LAUNCH(k_apply<kAdamFused>, apply_grid(s), 256, 0, st, apply_args(s, w));
    return FR_OK;
  }
  LazyArgs u, i;
  int rc = lazy_sides(s, w, st, who, &u, &i);
  if (rc) return rc;
  lazy_apply_segments(u, s->last_step_u, s->B, st);
  lazy_apply_segments(i, s->last_step_i, s->B, st);
  return FR_OK;
}

}  // namespace fr

extern "C" {

size_t fr_focf_workspace_bytes(int32_t n_users, int32_t n_items, int32_t d, int32_t max_batch) {
  fr::Carver c(nullptr, 0);
  fr::carve(c, n_users, n_items, d, max_batch);
  return c.off;
}

int fr_focf_workspace_init(void *workspace, size_t workspace_bytes, int32_t n_users, int32_t n_items, int32_t d,
                           int32_t max_batch, void *stream) {
  FR_REQUIRE(workspace, "fr_focf_workspace_init: null workspace");
  fr::Carver c(workspace, workspace_bytes);
  fr::FocfWs w = fr::carve(c, n_users, n_items, d, max_batch);
  if (!c.ok()) {
    fr::set_error("fr_focf_workspace_init: workspace too small (%zu < %zu bytes)", workspace_bytes, c.off);
    return FR_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  FR_CUDA_OK(cudaMemsetAsync(w.row_tab_u, 0, sizeof(uint2) * (size_t)n_users, st));
  FR_CUDA_OK(cudaMemsetAsync(w.row_tab_i, 0, sizeof(uint2) * (size_t)n_items, st));
  uint32_t ctrl[fr::CTRL_WORDS] = {0};
  ctrl[fr::CTRL_STAMP] = 1u;
  ctrl[fr::CTRL_MIN] = 0xffffffffu;
  ctrl[fr::CTRL_STRIDE] = 1u;
  FR_CUDA_OK(cudaMemcpyAsync(w.ctrl, ctrl, sizeof(ctrl), cudaMemcpyHostToDevice, st));
  FR_CUDA_OK(cudaStreamSynchronize(st));  // ctrl[] is a stack buffer
  return FR_OK;
}

int fr_focf_set_counters(void *workspace, size_t workspace_bytes, int32_t n_users, int32_t n_items, int32_t d,
                         int32_t max_batch, int32_t plan_cursor, int32_t adam_step, int32_t stride, void *stream) {
  FR_REQUIRE(workspace, "fr_focf_set_counters: null workspace");
  fr::Carver c(workspace, workspace_bytes);
  fr::FocfWs w = fr::carve(c, n_users, n_items, d, max_batch);
  FR_LAUNCH(fr::k_set_ctrl, 1, 1, 0, stream, w.ctrl, plan_cursor, adam_step, stride);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_focf_forward(const fr_focf_step *s, void *stream) {
  int rc = fr::check_step(s, false, "fr_focf_forward");
  if (rc) return rc;
  fr::FocfWs w;
  if ((rc = fr::carve_checked(s, &w, "fr_focf_forward"))) return rc;
  // a planned training step driven by the device-resident counters (data-parallel graph replay: forward -> backward ->
  // all-reduce -> fr_focf_adam): the loss kernel advances the Adam step count together with the batch cursor
  const bool advance = fr::planned(s) && s->step <= 0 && s->mU != nullptr;
  if ((rc = fr::forward_impl(s, w, (cudaStream_t)stream, advance))) return rc;
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_focf_backward(const fr_focf_step *s, float grad_scale, void *stream) {
  int rc = fr::check_step(s, false, "fr_focf_backward");
  if (rc) return rc;
  FR_REQUIRE(s->dU && s->dI, "fr_focf_backward: dU/dI missing");
  fr::FocfWs w;
  if ((rc = fr::carve_checked(s, &w, "fr_focf_backward"))) return rc;
  fr::grads_impl(s, w, grad_scale, (cudaStream_t)stream);
  FR_LAUNCH(fr::k_apply<fr::kDenseOut>, fr::apply_grid(s), 256, 0, stream, fr::apply_args(s, w));
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_focf_adam(const fr_focf_step *s, void *stream) {
  FR_REQUIRE(s && s->U && s->I && s->mU && s->vU && s->mI && s->vI && s->dU && s->dI && s->workspace,
             "fr_focf_adam: null pointer");
  FR_REQUIRE(s->d >= 4 && s->d % 4 == 0, "fr_focf_adam: bad d");   /* step <= 0: the workspace's device-resident count */
  fr::Carver c(s->workspace, s->workspace_bytes);
  fr::FocfWs w = fr::carve(c, s->n_users, s->n_items, s->d, 1);
  FR_LAUNCH(fr::k_apply<fr::kAdamDense>, fr::apply_grid(s), 256, 0, stream, fr::apply_args(s, w));
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_focf_train_step(const fr_focf_step *s, void *stream) {
  int rc = fr::check_step(s, true, "fr_focf_train_step");
  if (rc) return rc;
  fr::FocfWs w;
  if ((rc = fr::carve_checked(s, &w, "fr_focf_train_step"))) return rc;
  if (fr::fused_eligible(s)) {
    if ((rc = fr::prepare_impl(s, w, (cudaStream_t)stream))) return rc;
    if ((rc = fr::fused_step_impl(s, w, (cudaStream_t)stream))) return rc;
    FR_LAUNCH_CHECK();
    return FR_OK;
  }
  if ((rc = fr::prepare_impl(s, w, (cudaStream_t)stream))) return rc;
  if ((rc = fr::lazy_catchup(s, w, (cudaStream_t)stream, "fr_focf_train_step"))) return rc;
  if ((rc = fr::forward_only_impl(s, w, (cudaStream_t)stream, s->step <= 0))) return rc;
  fr::grads_impl(s, w, 1.0f, (cudaStream_t)stream);
  if ((rc = fr::adam_stage(s, w, (cudaStream_t)stream, "fr_focf_train_step"))) return rc;
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_focf_step_trace(uint64_t *out_host, int32_t n) {
  FR_REQUIRE(out_host && n >= 0 && n <= 8 * 256, "fr_focf_step_trace: bad argument");
  FR_REQUIRE(fr::g_fused_trace, "fr_focf_step_trace: set FR_FOCF_TRACE=1 before the first fused step");
  FR_CUDA_OK(cudaDeviceSynchronize());
  FR_CUDA_OK(cudaMemcpy(out_host, fr::g_fused_trace, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToHost));
  return FR_OK;
}

int fr_focf_adam_flush(const fr_focf_step *s, void *stream) {
  FR_REQUIRE(s && s->U && s->I && s->mU && s->vU && s->mI && s->vI, "fr_focf_adam_flush: null pointer");
  if (s->adam_mode == FR_ADAM_DENSE_EXACT || s->step < 1) return FR_OK;   // nothing is ever pending
  FR_REQUIRE(s->last_step_u && s->last_step_i && s->d >= 4 && s->d % 4 == 0, "fr_focf_adam_flush: lazy_exact state missing");
  cudaStream_t st = (cudaStream_t)stream;
  fr::LazyCommon c{s->d, s->step, s->lr, s->beta1, s->beta2, s->eps, s->weight_decay, s->adam_scalars, s->scalars_cap,
                   s->scalars_filled};
  int rc = fr::lazy_fill_scalars(c, st, "fr_focf_adam_flush");
  if (rc) return rc;
  fr::lazy_flush_table(fr::lazy_base(c, s->U, s->mU, s->vU, s->last_step_u), s->last_step_u, s->n_users, st);
  fr::lazy_flush_table(fr::lazy_base(c, s->I, s->mI, s->vI, s->last_step_i), s->last_step_i, s->n_items, st);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

/* trainer.py:181-196 for batches that are still in HOST memory: per step one H2D copy of the packed batch into `stage`,
 * fr_focf_train_step on it, one D2H copy of the loss; the host waits for the loss of step k-1 after it has enqueued step
 * k and returns once the last loss has arrived.  Everything runs on `stream`, so the copy of batch k+1 into the single
 * staging buffer is ordered behind the kernels that read batch k. */
int fr_focf_train_steps_host(const fr_focf_step *tmpl, int32_t n_steps, const void *const *host_batches,
                             const int32_t *batch_rows, void *stage, size_t stage_bytes, float *loss_dev,
                             float *loss_host, void *stream) {
  FR_REQUIRE(tmpl && n_steps >= 0, "fr_focf_train_steps_host: null step template or negative step count");
  if (n_steps == 0) return FR_OK;
  FR_REQUIRE(host_batches && batch_rows && stage && loss_dev && loss_host, "fr_focf_train_steps_host: null pointer");
  FR_REQUIRE(tmpl->step >= 1, "fr_focf_train_steps_host: step must be the 1-based optimizer step of the first batch");
  FR_REQUIRE(!tmpl->plan_desc && !tmpl->B_dev, "fr_focf_train_steps_host: takes host batches, not a planned epoch");
  for (int32_t k = 0; k < n_steps; ++k) {
    FR_REQUIRE(host_batches[k] && batch_rows[k] >= 1, "fr_focf_train_steps_host: batch %d is empty", k);
    FR_REQUIRE((size_t)batch_rows[k] * 16 <= stage_bytes, "fr_focf_train_steps_host: batch %d (%d rows) exceeds the staging buffer (%zu bytes)",
               k, batch_rows[k], stage_bytes);
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  for (int i = 0; i < 2; ++i) {
    cudaError_t e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
    if (e != cudaSuccess) {
      if (i == 1) cudaEventDestroy(ev[0]);
      fr::set_error("fr_focf_train_steps_host: cudaEventCreate failed: %s", cudaGetErrorString(e));
      return FR_ERR_CUDA;
    }
  }
  fr_focf_step s = *tmpl;
  int rc = FR_OK;
  cudaError_t e = cudaSuccess;
  for (int32_t k = 0; k < n_steps; ++k) {
    const size_t n = (size_t)batch_rows[k];
    if ((e = cudaMemcpyAsync(stage, host_batches[k], 16 * n, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
    const char *base = (const char *)stage;
    s.uid = (const int32_t *)base;
    s.iid = (const int32_t *)(base + 4 * n);
    s.rating = (const float *)(base + 8 * n);
    s.sst = (const float *)(base + 12 * n);
    s.B = (int32_t)n;
    s.step = tmpl->step + k;
    s.loss = loss_dev + (k & 1);
    if ((rc = fr_focf_train_step(&s, stream))) break;
    if ((e = cudaMemcpyAsync(loss_host + k, s.loss, sizeof(float), cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
    if ((e = cudaEventRecord(ev[k & 1], st)) != cudaSuccess) break;
    if (k >= 1 && (e = cudaEventSynchronize(ev[(k - 1) & 1])) != cudaSuccess) break;   // loss of step k-1 is on the host
  }
  if (rc == FR_OK && e == cudaSuccess) e = cudaEventSynchronize(ev[(n_steps - 1) & 1]);
  cudaEventDestroy(ev[0]);
  cudaEventDestroy(ev[1]);
  if (rc) return rc;
  if (e != cudaSuccess) {
    fr::set_error("fr_focf_train_steps_host: %s", cudaGetErrorString(e));
    return FR_ERR_CUDA;
  }
  return FR_OK;
}

/* the two halves of fr_focf_train_step, for callers that overlap the preparation of batch t+1 (another stream, another
 * workspace) with the compute of batch t: prepare = batch gather (planned) + sort / segments / row stamps (independent
 * of the embedding tables); compute = forward + loss + gradients + Adam */
int fr_focf_step_prepare(const fr_focf_step *s, void *stream) {
  int rc = fr::check_step(s, true, "fr_focf_step_prepare");
  if (rc) return rc;
  fr::FocfWs w;
  if ((rc = fr::carve_checked(s, &w, "fr_focf_step_prepare"))) return rc;
  if ((rc = fr::prepare_impl(s, w, (cudaStream_t)stream))) return rc;
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_focf_step_compute(const fr_focf_step *s, void *stream) {
  int rc = fr::check_step(s, true, "fr_focf_step_compute");
  if (rc) return rc;
  fr::FocfWs w;
  if ((rc = fr::carve_checked(s, &w, "fr_focf_step_compute"))) return rc;
  if (fr::fused_eligible(s)) {
    if ((rc = fr::fused_step_impl(s, w, (cudaStream_t)stream))) return rc;
  } else {
    if ((rc = fr::lazy_catchup(s, w, (cudaStream_t)stream, "fr_focf_step_compute"))) return rc;
    if ((rc = fr::forward_only_impl(s, w, (cudaStream_t)stream, s->step <= 0))) return rc;
    fr::grads_impl(s, w, 1.0f, (cudaStream_t)stream);
    if ((rc = fr::adam_stage(s, w, (cudaStream_t)stream, "fr_focf_step_compute"))) return rc;
  }
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_pair_scores(const float *U, const float *I, const int32_t *uid, const int32_t *iid, int64_t n, int32_t d,
                   int32_t transform, float max_rating, float *out, void *stream) {
  if (n == 0) return FR_OK;
  FR_REQUIRE(U && I && uid && iid && out && n > 0, "fr_pair_scores: null pointer");
  FR_REQUIRE(d >= 4 && d % 4 == 0, "fr_pair_scores: d=%d must be a multiple of 4", d);
  FR_LAUNCH(fr::k_pair_scores, fr::grid_for(n, 256), 256, 0, stream, U, I, uid, iid, n, d, transform, max_rating,
            out);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_focf_gather_batch(const int32_t *item_off, const int32_t *train_uid, const float *train_rating,
                         const float *sst_of_user, const int32_t *draw_items, const int32_t *draw_off, int32_t J,
                         int32_t *uid, int32_t *iid, float *rating, float *sst, void *stream) {
  FR_REQUIRE(item_off && train_uid && train_rating && sst_of_user && draw_items && draw_off && uid && iid && rating &&
                 sst && J >= 1,
             "fr_focf_gather_batch: bad argument");
  FR_LAUNCH(fr::k_gather_batch, fr::grid_for((int64_t)J * 64, 256, fr::kSMs * 8), 256, 0, stream, item_off, train_uid, train_rating,
            sst_of_user, draw_items, draw_off, J, nullptr, 0, nullptr, uid, iid, rating, sst);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

}  // extern "C"
