// FOCF planned epoch as ONE persistent cooperative launch (the latency-bound regime: ML-1M-shaped tables, batches of a few
// thousand rows).
//
// Reference being replaced: the body of Trainer._train_epoch (trainer.py:181-196) over the batches of
// FOCFDataLoader._next_batch_data (focf_dataloader.py:37-50) -- for every batch: gather the drawn items' rows, forward
// (focf.py:136-143), item x group statistics + fairness objective + loss (focf.py:75-134, 152-169), backward, dense Adam
// with L2 weight decay (trainer.py:139, 196).
//
// The stepwise path (focf_train.cu) spends a step of ~25 us in four dependent phases separated by grid barriers and a kernel
// boundary, each phase a chain of 3-5 dependent L2 round trips over a few thousand rows.  Here the whole epoch is one launch:
//   * PRODUCER CTAs (two per slot: item side, user side) build the batches ahead of the compute: gather from the
//     item-sorted train split, stable radix sort by user in shared memory, segments, row stamps, batch min / max of the
//     attribute, the step's Adam scalars.  n_slots workspaces rotate; a producer re-fills a slot once the step that used
//     it has finished (`done` counter), the compute waits on the slot's `ready` words.
//   * COMPUTE CTAs keep their share of [U; I] and of both Adam moments RESIDENT IN SHARED MEMORY for the whole launch
//     (22 MB at the ML-1M shape = 52 KB per CTA): the dense Adam sweep touches no global memory except the store of the
//     new parameter value; rows the batch does not touch are updated while the CTA waits at the first grid barrier.
//   * everything the later phases need from the producer (row stamps -> gradient segments of the resident rows, the
//     gradient chunk's sorted entries, rating / attribute columns) is fetched before the barrier that precedes its use.
// The arithmetic is the stepwise path's (forward FMA chain + shuffle tree, fused_stats, grads_chunk, adam1 with the same
// scalars): tables, moments and losses are bit-identical to fr_focf_train_step over the same planned batches
// (tests/test_focf_epoch_gpu.py).
#include <stdlib.h>

#include "sort.cuh"

#include "focf_device.cuh"

namespace fr {

constexpr int kEpThreads = 512, kEpWarps = kEpThreads / 32;
constexpr int kEpMaxRows = 8192;      // batch rows (16 sorted keys per producer thread)
constexpr int kEpMaxSlots = 8;
constexpr int kEpMaxR = 4;           // resident float4 elements per compute thread (what shared memory holds at most)
// mailbox of a slot inside its workspace's control block (words the stepwise path does not use)
enum { EP_READY_U = 32, EP_READY_I = 33, EP_B = 34, EP_STAMP = 35 };

struct EpSlot {
  int32_t *uid, *iid;
  float *rating, *sst, *pred;
  FocfWs w;
};

struct EpochArgs {
  float *U, *I, *mU, *vU, *mI, *vI;
  int n_users, n_items, d;
  const int32_t *plan_desc, *plan_items, *plan_offs;
  int plan_len;
  const int32_t *item_off, *train_uid;
  const float *train_rating, *sst_of_user;
  int objective;
  float fair_weight;
  float *loss;
  int32_t *flags;
  double lr, beta1, beta2, eps, wd;
  int first_batch, n_steps, adam_step0;   // adam_step0: 1-based optimizer step of the first batch
  int n_slots, n_prod, cap, chunk, key_bits_u, R;
  unsigned long long *sync;    // [0] arrival counter of the compute CTAs' barrier, [1] steps completed
  unsigned long long *trace;   // optional: [gridDim.x][8] %globaltimer stamps of step `trace_step`
  int trace_step;
  const float *sc_tab;   // [2 * n_steps] Adam scalars of the launch's steps (k_adam_scalars: apply_body's expressions)
  int dbg_skip;   // timing experiments only (FR_FOCF_EPOCH_SKIP bit mask: phases left out; results are then meaningless)
  EpSlot slot[kEpMaxSlots];
};

// spin until `cond` is false.  -DFR_EPOCH_WATCHDOG: give up after ~1 s, say where, and run on (debug builds only)
#ifdef FR_EPOCH_WATCHDOG
#define EP_SPIN(cond, id, info)                                                                     \
  do {                                                                                              \
    unsigned long long _n = 0;                                                                      \
    while (cond) {                                                                                  \
      if (++_n > (1ull << 20)) {                                                                    \
        printf("EP WATCHDOG loop %d block %d info %d\n", (id), (int)blockIdx.x, (int)(info));       \
        break;                                                                                      \
      }                                                                                             \
    }                                                                                               \
  } while (0)
#else
#define EP_SPIN(cond, id, info) \
  do {                          \
    while (cond) {              \
    }                           \
  } while (0)
#endif
__device__ __forceinline__ unsigned long long ep_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint32_t ld_vol(const uint32_t *p) { return *(const volatile uint32_t *)p; }
// Release / acquire at GPU scope for the flags and the barrier counter.  NOT __threadfence(): that is fence.sc
// (MEMBAR.SC.GPU), and ~140 CTAs issuing it in the same microsecond serialise -- 0.9 .. 6.3 us per fence in the phase
// trace, the largest single item of a step.
__device__ __forceinline__ uint32_t ld_acq(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acq(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_rel(uint32_t *p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_rel(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long atom_add_rel(unsigned long long *p, unsigned long long v) {
  unsigned long long old;
  asm volatile("atom.add.release.gpu.global.u64 %0, [%1], %2;" : "=l"(old) : "l"(p), "l"(v) : "memory");
  return old;
}

// ------------------------------------------------------------------------------------------ producer
// Shared memory of a producer CTA (cap = row capacity of the slots):
//   kA vA kB vB [cap] u32 | rnk [cap] u16 | cnt [16][256] u16 | dbase [256] u32 | wsum [32] u32 | s_off [cap + 1] i32
static size_t ep_prod_smem(int cap) {
  return (size_t)cap * 16 + (((size_t)cap * 2 + 15) & ~(size_t)15) + kEpWarps * 256 * 2 + 256 * 4 + 32 * 4 +
         ((size_t)cap + 1) * 4 + 64;
}
static size_t ep_comp_smem(int cap, int R) {
  // pred rating sst [cap] | cseg [2 cap] | cglob [8] floats, then per resident element: p m v (float4) + {row, seg, chunks} (int4)
  return ((size_t)5 * cap + 8) * 4 + 16 + (size_t)R * kEpThreads * (48 + 16);
}

__device__ __forceinline__ void ep_produce(const EpochArgs &a, int k, bool user_side, unsigned char *smem) {
  const int cap = a.cap, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  uint32_t *kA = (uint32_t *)smem, *vA = kA + cap, *kB = vA + cap, *vB = kB + cap;
  uint16_t *rnk = (uint16_t *)(vB + cap);
  uint16_t *cnt = (uint16_t *)((unsigned char *)rnk + (((size_t)cap * 2 + 15) & ~(size_t)15));   // [kEpWarps][256]
  uint32_t *dbase = (uint32_t *)(cnt + kEpWarps * 256);
  uint32_t *wsum = dbase + 256;
  int32_t *s_off = (int32_t *)(wsum + 32);
  __shared__ uint32_t s_mm[2 * kEpWarps];
  const EpSlot &sl = a.slot[k];
  uint32_t *ctrl = sl.w.ctrl;
  const uint32_t stamp_base = ld_vol(ctrl + CTRL_STAMP);
  const bool tr_cta = a.trace != nullptr;
  int uses = 0;
  for (int i = k; i < a.n_steps; i += a.n_slots, ++uses) {
    if (i >= a.n_slots) {   // the slot is free once the step that used it last has finished
      if (tid == 0) {
        EP_SPIN(ld_acq(a.sync + 1) < (unsigned long long)(i - a.n_slots + 1), 1, i);
      }
      __syncthreads();
    }
    const bool tr = tr_cta && i >= a.trace_step && i < a.trace_step + a.n_slots;
    if (tr && tid == 0) a.trace[blockIdx.x * 16 + 0] = ep_now();
    const uint32_t stamp = stamp_base + (uint32_t)uses;
    const int32_t *dsc = a.plan_desc + 4 * (int)((unsigned)(a.first_batch + i) % (unsigned)a.plan_len);
    const int32_t *draw_items = a.plan_items + __ldg(dsc), *draw_off = a.plan_offs + __ldg(dsc + 1);
    const int J = __ldg(dsc + 2), n = __ldg(dsc + 3);
    for (int j = tid; j <= J; j += kEpThreads) s_off[j] = __ldg(draw_off + j);
    __syncthreads();
    // focf_dataloader.py:37-50: the batch = all train rows of the drawn items (k_gather_batch's arithmetic)
    uint32_t lo_o = 0xffffffffu, hi_o = 0u;
    for (int p = tid; p < n; p += kEpThreads) {
      int lo = 0, hi = J;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_off[mid] <= p) lo = mid; else hi = mid;
      }
      const int it = __ldg(draw_items + lo);
      const int src = __ldg(a.item_off + it) + (p - s_off[lo]);
      if (user_side) {
        const int u = __ldg(a.train_uid + src);
        const float sv = __ldg(a.sst_of_user + u);
        sl.uid[p] = u;
        sl.sst[p] = sv;
        kA[p] = (uint32_t)u;
        vA[p] = (uint32_t)p;
        const uint32_t o = f2ord(sv);
        lo_o = min(lo_o, o);
        hi_o = max(hi_o, o);
      } else {
        sl.iid[p] = it;
        sl.rating[p] = __ldg(a.train_rating + src);
        kA[p] = (uint32_t)it;
        vA[p] = (uint32_t)p;
      }
    }
    if (user_side) {
      lo_o = __reduce_min_sync(0xffffffffu, lo_o);
      hi_o = __reduce_max_sync(0xffffffffu, hi_o);
      if (lane == 0) {
        s_mm[w] = lo_o;
        s_mm[kEpWarps + w] = hi_o;
      }
    }
    __syncthreads();
    if (tr && tid == 0) a.trace[blockIdx.x * 16 + 1] = ep_now();

    if (user_side) {   // stable LSD radix sort by user id (k_prepare_small's ranking, 16 warps)
      const int rounds = (n + kEpThreads - 1) / kEpThreads, chunk = rounds * 32;
      const int passes = (a.key_bits_u + 7) / 8;
      for (int pass = 0; pass < passes; ++pass) {
        const int shift = 8 * pass;
        for (int q = tid; q < kEpWarps * 256; q += kEpThreads) cnt[q] = 0;
        __syncthreads();
        for (int r = 0; r < rounds; ++r) {
          const int p = w * chunk + r * 32 + lane;
          const bool valid = p < n;
          const uint32_t dig = valid ? ((kA[p] >> shift) & 255u) : 0xffffffffu;
          const unsigned peers = __match_any_sync(0xffffffffu, dig);
          const uint32_t before = valid ? cnt[w * 256 + dig] : 0u;
          __syncwarp();
          if (valid && lane == (__ffs(peers) - 1)) cnt[w * 256 + dig] = (uint16_t)(before + __popc(peers));
          __syncwarp();
          if (valid) rnk[p] = (uint16_t)(before + __popc(peers & ((1u << lane) - 1u)));
        }
        __syncthreads();
        uint32_t tot = 0;
        if (tid < 256) {   // digit tid: exclusive prefix over the warps, then over the digits
#pragma unroll 8
          for (int ww = 0; ww < kEpWarps; ++ww) {
            const uint32_t t = cnt[ww * 256 + tid];
            cnt[ww * 256 + tid] = (uint16_t)tot;
            tot += t;
          }
          uint32_t inc = tot;
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
        for (int r = 0; r < rounds; ++r) {
          const int p = w * chunk + r * 32 + lane;
          if (p < n) {
            const uint32_t key = kA[p], dig = (key >> shift) & 255u;
            const uint32_t pos = dbase[dig] + cnt[w * 256 + dig] + rnk[p];
            kB[pos] = key;
            vB[pos] = vA[p];
          }
        }
        __syncthreads();
        uint32_t *t0 = kA; kA = kB; kB = t0;
        t0 = vA; vA = vB; vB = t0;
      }
    }
    if (tr && tid == 0) a.trace[blockIdx.x * 16 + 2] = ep_now();

    // segments: thread tid owns ipt consecutive sorted positions
    {
      const int ipt = (n + kEpThreads - 1) / kEpThreads;
      const int p0 = tid * ipt, p1 = min(p0 + ipt, n);
      int c = 0;
      for (int p = p0; p < p1; ++p) c += (p == 0 || kA[p] != kA[p - 1]);
      int inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      __syncthreads();   // wsum reuse
      if (lane == 31) wsum[w] = (uint32_t)inc;
      __syncthreads();
      int run = inc - c;
      for (int ww = 0; ww < w; ++ww) run += (int)wsum[ww];
      int32_t *segid = user_side ? sl.w.segid_u : sl.w.segid_i;
      int32_t *segoff = user_side ? sl.w.segoff_u : sl.w.segoff_i;
      int32_t *count = user_side ? sl.w.Ju : sl.w.J;
      uint2 *row_tab = user_side ? sl.w.row_tab_u : sl.w.row_tab_i;
      uint32_t *skey = user_side ? sl.w.skey_u : sl.w.skey_i;
      for (int p = p0; p < p1; ++p) {
        const uint32_t key = kA[p];
        if (p == 0 || key != kA[p - 1]) {
          segoff[run] = p;
          row_tab[key] = make_uint2(stamp, (uint32_t)run);
          ++run;
        }
        const int sgm = run - 1;
        segid[p] = sgm;
        skey[p] = key;
        if (user_side) sl.w.ord_u[p] = vA[p];
        else sl.w.entry_seg[p] = sgm;      // item side: entry order == sorted order (whole-item batches)
        if (p == n - 1) {
          segoff[sgm + 1] = n;
          count[0] = sgm + 1;
        }
      }
    }
    if (user_side && tid == 0) {
      uint32_t lo = 0xffffffffu, hi = 0u;
      for (int ww = 0; ww < kEpWarps; ++ww) {
        lo = min(lo, s_mm[ww]);
        hi = max(hi, s_mm[kEpWarps + ww]);
      }
      ctrl[CTRL_MIN] = lo;
      ctrl[CTRL_MAX] = hi;
      ctrl[EP_B] = (uint32_t)n;
      ctrl[EP_STAMP] = stamp;
    }
    __syncthreads();
    if (tid == 0) st_rel(ctrl + (user_side ? EP_READY_U : EP_READY_I), (uint32_t)(i + 1));   // (cumulative over the CTA's writes)
    if (tr && tid == 0) a.trace[blockIdx.x * 16 + 3] = ep_now();
  }
  // leave the slot's workspace as the stepwise path expects it: stamp advanced past every use, hand-over words re-armed
  if (user_side && tid == 0) {
    // the last step must have read the words before they are re-armed
    EP_SPIN(ld_acq(a.sync + 1) < (unsigned long long)a.n_steps, 2, a.n_steps);
    ctrl[CTRL_STAMP] = stamp_base + (uint32_t)uses;
    ctrl[CTRL_MIN] = 0xffffffffu;
    ctrl[CTRL_MAX] = 0u;
  }
}

// ------------------------------------------------------------------------------------------ compute
struct EpBar {
  unsigned long long *ctr;
  unsigned long long n, target;   // the counter starts at zero at launch: the k-th barrier is complete at k * n arrivals
};
__device__ __forceinline__ void ep_arrive(EpBar &b) {
  __syncthreads();
  b.target += b.n;
  if (threadIdx.x == 0)   // release (the CTA's writes first), no return value: nothing waits for the round trip
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(b.ctr), "l"(1ull) : "memory");
}
__device__ __forceinline__ void ep_wait(EpBar &b) {
  if (threadIdx.x == 0) {
    EP_SPIN(ld_acq(b.ctr) < b.target, 3, (int)(b.target / b.n));
  }
  __syncthreads();
}

// The gradient chunk's rows in two steps, so that the loads can be issued ahead of the barrier that precedes their use
// (run_chunk's arithmetic: entries in sorted order, acc = fmaf(coef, row, acc), partials flushed at segment ends).
// kHalf (d <= 64): a warp request covers TWO rows -- lanes 0-15 row 2e, lanes 16-31 row 2e+1 -- so the 16 entries of a
// chunk are one round of 8 requests; the odd rows are moved to the accumulating half by a shuffle.  Otherwise x holds the
// first 8 rows and the second round is loaded when the first has been consumed.
template <bool kHalf>
__device__ __forceinline__ void ep_issue_rows(const GradArgs &a, const ChunkStage &s, int l0, float4 (&x)[8]) {
  const int lane = threadIdx.x & 31;
  const float *other = s.user_side ? a.I : a.U;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int l = kHalf ? l0 + 2 * e + (lane >> 4) : l0 + e;
    const int k = kHalf ? (lane & 15) * 4 : lane * 4;
    const int o = __shfl_sync(0xffffffffu, s.my_oid, l & 31);
    x[e] = (l < s.nvalid && k < a.d) ? __ldcg((const float4 *)(other + (size_t)o * a.d) + (k >> 2)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <bool kHalf>
__device__ __forceinline__ void ep_consume_rows(const GradArgs &a, const ChunkStage &s, float4 (&x)[8]) {
  if (s.nvalid == 0) return;
  const int lane = threadIdx.x & 31;
  const int c = s.c, nvalid = s.nvalid, d = a.d, B = a.B;
  float *gseg = s.user_side ? a.gseg_u : a.gseg_i;
  float *head = s.user_side ? a.head_u : a.head_i;
  float *tail = s.user_side ? a.tail_u : a.tail_i;
  const float vmin = ord2f(__ldcg(a.ctrl + CTRL_MIN));
  float my_coef = 0.f;
  if (lane < nvalid) {
    const int gq = a.sst[s.my_b] != vmin;
    my_coef = (2.f * (a.pred[s.my_b] - a.rating[s.my_b]) / (float)B + a.cseg[2 * s.my_es + gq] + a.cglob[gq]) * a.grad_scale;
  }
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int cur = __shfl_sync(0xffffffffu, s.my_seg, 0);
  const int kcol = kHalf ? (lane & 15) * 4 : lane * 4;
  const bool writer = kHalf ? lane < 16 : true;
  auto flush = [&](int sg) {
    float *dst = (sg != s.seg_before && sg != s.seg_after) ? gseg + (size_t)sg * d
                 : (sg == s.seg_before)                     ? head + (size_t)c * d
                                                            : tail + (size_t)c * d;
    if (writer && kcol < d) *(float4 *)(dst + kcol) = acc;
    acc = make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto entry = [&](int l, const float4 &row) {   // l < nvalid is warp-uniform
    const int sg = __shfl_sync(0xffffffffu, s.my_seg, l & 31);
    const float cf = __shfl_sync(0xffffffffu, my_coef, l & 31);
    if (l < nvalid) {
      if (sg != cur) {
        flush(cur);
        cur = sg;
      }
      acc.x = fmaf(cf, row.x, acc.x);
      acc.y = fmaf(cf, row.y, acc.y);
      acc.z = fmaf(cf, row.z, acc.z);
      acc.w = fmaf(cf, row.w, acc.w);
    }
  };
  for (int l0 = 0; l0 < nvalid; l0 += (kHalf ? 16 : 8)) {
    if (l0) ep_issue_rows<kHalf>(a, s, l0, x);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (kHalf) {
        entry(l0 + 2 * e, x[e]);
        float4 odd;   // row 2e+1 lives in the upper half-warp
        odd.x = __shfl_xor_sync(0xffffffffu, x[e].x, 16);
        odd.y = __shfl_xor_sync(0xffffffffu, x[e].y, 16);
        odd.z = __shfl_xor_sync(0xffffffffu, x[e].z, 16);
        odd.w = __shfl_xor_sync(0xffffffffu, x[e].w, 16);
        entry(l0 + 2 * e + 1, odd);
      } else {
        entry(l0 + e, x[e]);
      }
    }
  }
  flush(cur);
}

__global__ void __launch_bounds__(kEpThreads, 1) k_focf_epoch(const __grid_constant__ EpochArgs a) {
  extern __shared__ __align__(16) unsigned char ep_smem[];
  const int tid = threadIdx.x, lane = tid & 31;
  if ((int)blockIdx.x < a.n_prod) {
    ep_produce(a, (int)blockIdx.x >> 1, (blockIdx.x & 1) != 0, ep_smem);
    return;
  }
  __shared__ float sh[33];
  const int cap = a.cap, R = a.R, d = a.d, dq = d >> 2;
  float *fsm = (float *)ep_smem;
  float *s_rat = fsm + cap, *s_sst = fsm + 2 * cap, *s_cseg = fsm + 3 * cap, *s_cglob = fsm + 5 * cap;
  float4 *s_state = (float4 *)(((uintptr_t)(fsm + 5 * cap + 8) + 15) & ~(uintptr_t)15);   // [R][3][kEpThreads]
  int4 *s_meta = (int4 *)(s_state + (size_t)R * 3 * kEpThreads);                          // [R][kEpThreads]
  const int n_comp = (int)gridDim.x - a.n_prod, cta = (int)blockIdx.x - a.n_prod;
  const int NT = n_comp * kEpThreads, g = cta * kEpThreads + tid;
  const int gwarp = g >> 5, nwarps = NT >> 5;
  const long long nq_u = (long long)a.n_users * dq, nq = nq_u + (long long)a.n_items * dq;
  EpBar bar{a.sync, (unsigned long long)n_comp, 0ull};
  const bool tr_cta = a.trace != nullptr;
  const float w1 = (float)(1.0 - a.beta1), w2 = (float)(1.0 - a.beta2), b2 = (float)a.beta2, wd = (float)a.wd,
              eps = (float)a.eps;

  // resident share of [U; I] and both moments: element q = g + r * NT (float4 units)
  for (int r = 0; r < R; ++r) {
    const long long q = (long long)g + (long long)r * NT;
    int4 meta = make_int4(-1, 0, 0, 0);
    if (q < nq) {
      const bool is_item = q >= nq_u;
      const long long ql = is_item ? q - nq_u : q;
      meta.x = (int)(ql / dq) | (is_item ? (int)0x80000000 : 0);
      const float4 *pp = (const float4 *)((is_item ? a.I : a.U) + ql * 4);
      const float4 *pm = (const float4 *)((is_item ? a.mI : a.mU) + ql * 4);
      const float4 *pv = (const float4 *)((is_item ? a.vI : a.vU) + ql * 4);
      s_state[(r * 3 + 0) * kEpThreads + tid] = *pp;
      s_state[(r * 3 + 1) * kEpThreads + tid] = *pm;
      s_state[(r * 3 + 2) * kEpThreads + tid] = *pv;
    }
    s_meta[r * kEpThreads + tid] = meta;
  }

  // What a step needs from its slot is fetched AHEAD of the step.  Two steps ahead (in the shadow of barrier 3 of step
  // i - 2): the slot's mailbox -- row count, stamp, Adam scalars, segment count, attribute range (`nx`).  One step ahead
  // (shadow of barrier 3 of step i - 1): the ids of the warp's first forward group and, for every resident element,
  // whether its row is touched by the batch and which gradient segment / chunks it reads then.
  struct Mail {
    int B, J;
    uint32_t stamp, vmin, vmax;
    float neg_step, bc2s;
  };
  auto poll_ready = [&](int i) {   // one thread polls; the CTA's other threads learn it at the next __syncthreads
    uint32_t *ctrl = a.slot[i % a.n_slots].w.ctrl;
    if (tid == 0) {
      EP_SPIN(ld_acq(ctrl + EP_READY_U) < (uint32_t)(i + 1) || ld_acq(ctrl + EP_READY_I) < (uint32_t)(i + 1), 4, i);
    }
  };
  auto wait_ready = [&](int i) {
    poll_ready(i);
    __syncthreads();
  };
  auto load_mail = [&](int i, Mail &m) {   // ends with the loads in flight; first use is the caller's
    const EpSlot &sl = a.slot[i % a.n_slots];
    uint32_t *ctrl = sl.w.ctrl;
    m.B = (int)__ldcg(ctrl + EP_B);
    m.stamp = __ldcg(ctrl + EP_STAMP);
    m.neg_step = __ldg(a.sc_tab + 2 * i);       // (filled before the launch: double pow() on one thread of a producer
    m.bc2s = __ldg(a.sc_tab + 2 * i + 1);       //  CTA cost 6-12 us per batch on this part's FP64 rate)
    m.J = __ldcg(sl.w.J);
    m.vmin = __ldcg(ctrl + CTRL_MIN);
    m.vmax = __ldcg(ctrl + CTRL_MAX);
  };
  int pre_nu = 0, pre_ni = 0;
  auto fetch_rows = [&](int i, uint32_t stamp) {   // ids of the first forward group + the resident elements' gradient sources
    const EpSlot &sl = a.slot[i % a.n_slots];
    pre_nu = pre_ni = 0;
    if (lane < 4) {   // (rows past the batch end hold stale ids of an earlier batch: never used)
      const int b = min(gwarp * 4 + lane, cap - 1);
      pre_nu = __ldcg(sl.uid + b);
      pre_ni = __ldcg(sl.iid + b);
    }
    if (a.dbg_skip & 32) return;
    // (all row-stamp loads first, then all segment-offset loads: two L2 round trips for the thread's R elements together)
    int mx[kEpMaxR];
    uint2 tab[kEpMaxR];
#pragma unroll
    for (int r = 0; r < kEpMaxR; ++r) {
      mx[r] = r < R ? s_meta[r * kEpThreads + tid].x : -1;
      tab[r] = make_uint2(0u, 0u);
      if (mx[r] != -1) tab[r] = __ldcg((mx[r] < 0 ? sl.w.row_tab_i : sl.w.row_tab_u) + (mx[r] & 0x7fffffff));
    }
    int s0[kEpMaxR], s1[kEpMaxR];
#pragma unroll
    for (int r = 0; r < kEpMaxR; ++r) {
      s0[r] = s1[r] = 0;
      if (mx[r] != -1 && tab[r].x == stamp) {   // touched by this batch: its gradient segment ...
        const int32_t *segoff = mx[r] < 0 ? sl.w.segoff_i : sl.w.segoff_u;
        s0[r] = __ldcg(segoff + tab[r].y);
        s1[r] = __ldcg(segoff + tab[r].y + 1);
      }
    }
#pragma unroll
    for (int r = 0; r < kEpMaxR; ++r) {
      if (mx[r] == -1) continue;
      int4 meta = make_int4(mx[r], -1, 0, 0);
      if (tab[r].x == stamp) {                  // ... and the chunks the segment spans
        meta.y = (int)tab[r].y;
        meta.z = s0[r] / a.chunk;
        meta.w = (s1[r] - 1) / a.chunk;
      }
      s_meta[r * kEpThreads + tid] = meta;
    }
  };
  __shared__ int s_alloc;
  // (with two slots the slot of step i + 2 is the one step i is using: its mailbox cannot be waited for inside step i)
  const bool ahead2 = a.n_slots >= 3 && !(a.dbg_skip & 256);
  auto grad_args = [&](const EpSlot &sl, int B) {
    return GradArgs{a.U, a.I, sl.uid, sl.iid, s_rat, s_sst, fsm, B, d, nullptr, 0, 0, nullptr, sl.w.ord_u,
                    sl.w.segid_i, sl.w.segoff_i, sl.w.segid_u, sl.w.segoff_u, sl.w.entry_seg, s_cseg, s_cglob, sl.w.ctrl, 1.0f,
                    a.chunk, sl.w.gseg_i, sl.w.head_i, sl.w.tail_i, sl.w.gseg_u, sl.w.head_u, sl.w.tail_u, 1};
  };
  auto stage_first = [&](int i, int B) {   // the warp's first gradient chunk of step i: sorted entries, other-side row ids
    const GradArgs ga = grad_args(a.slot[i % a.n_slots], B);
    const int nchunk = (B + a.chunk - 1) / a.chunk;
    return stage_chunk(ga, nchunk, (gwarp < 2 * nchunk && !(a.dbg_skip & 4)) ? gwarp : 2 * nchunk);   // (2 * nchunk: empty)
  };
  Mail cur{}, nx{};
  wait_ready(0);
  load_mail(0, cur);
  fetch_rows(0, cur.stamp);
  ChunkStage cs = stage_first(0, cur.B);
  if (ahead2 && a.n_steps > 1) {
    wait_ready(1);
    load_mail(1, nx);
  }
  const bool half_rows = d <= 64;   // two rows per 512-byte warp request in the gradient gather

  for (int i = 0; i < a.n_steps; ++i) {
    const EpSlot &sl = a.slot[i % a.n_slots];
    uint32_t *ctrl = sl.w.ctrl;
    const bool tr = tr_cta && i == a.trace_step;
    if (tr && tid == 0) a.trace[blockIdx.x * 16 + 0] = ep_now();
    if (tid == 0) s_alloc = 0;
    const int B = cur.B;
    const float neg_step = cur.neg_step, bc2s = cur.bc2s;

    // ---- forward (focf.py:136-143): forward_body's arithmetic, rows through L2 (the tables change during the launch)
    if (!(a.dbg_skip & 1)) {
      int nu = pre_nu, ni = pre_ni;
      for (int b0 = gwarp * 4; b0 < B; b0 += nwarps * 4) {
        const int cu = nu, ci = ni;
        const int nb = b0 + nwarps * 4 + lane;
        if (lane < 4 && nb < B) {
          nu = __ldcg(sl.uid + nb);
          ni = __ldcg(sl.iid + nb);
        }
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int u = __shfl_sync(0xffffffffu, cu, e), it = __shfl_sync(0xffffffffu, ci, e);
          if (b0 + e < B) {
            const float4 *pu = (const float4 *)(a.U + (size_t)u * d);
            const float4 *pi = (const float4 *)(a.I + (size_t)it * d);
            for (int k = lane; k * 4 < d; k += 32) {
              const float4 x = __ldcg(pu + k), y = __ldcg(pi + k);
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
          const float sm = warp_sum(acc[e]);
          if (lane == e) mine = sm;
        }
        if (lane < 4 && b0 + lane < B) sl.pred[b0 + lane] = mine;
      }
    }
    if (tr && tid == 0) a.trace[blockIdx.x * 16 + 1] = ep_now();
    ep_arrive(bar);

    // ---- in the shadow of barrier 1: what the next two phases read -- the rows of the warp's gradient chunk (its
    // entries were staged a barrier earlier; the batch's rows do not change before the Adam phase), the bounds of the
    // warp's first item segment, the rating / attribute columns
    const GradArgs ga = grad_args(sl, B);
    const int nchunk = (B + a.chunk - 1) / a.chunk;
    StatsPre sp{cur.J, cur.vmin, cur.vmax, 0, 0};
    {
      const int wib = tid >> 5;
      if (wib < sp.J) {
        sp.s0 = __ldcg(sl.w.segoff_i + wib);
        sp.s1 = __ldcg(sl.w.segoff_i + wib + 1);
      }
    }
    float4 x[8];
    if (half_rows) ep_issue_rows<true>(ga, cs, 0, x); else ep_issue_rows<false>(ga, cs, 0, x);
    for (int p0 = tid; p0 < B; p0 += 4 * kEpThreads) {
      float vr[4], vs[4];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int p = min(p0 + q4 * kEpThreads, B - 1);
        vr[q4] = __ldcg(sl.rating + p);
        vs[q4] = __ldcg(sl.sst + p);
      }
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int p = p0 + q4 * kEpThreads;
        if (p < B) {
          s_rat[p] = vr[q4];
          s_sst[p] = vs[q4];
        }
      }
    }
    ep_wait(bar);
    if (tr && tid == 0) a.trace[blockIdx.x * 16 + 2] = ep_now();

    // ---- item x group statistics, fairness objective, loss (focf.py:75-134, 152-169): fused_stats, per CTA
    LossArgs la{sl.pred, s_rat, s_sst, nullptr, sl.w.segid_i, sl.w.segoff_i, sl.w.J, B, nullptr, a.plan_len, 0, a.objective,
                0, 0, nullptr, a.fair_weight, nullptr, nullptr, nullptr, nullptr, nullptr, a.loss, ctrl, a.flags};
    const bool loss_cta = cta == n_comp - 1;
    const float loss = (a.dbg_skip & 2) ? 0.f : fused_stats(la, B, cap, fsm, sh, s_cseg, s_cglob, true, loss_cta, (int)gridDim.x - 1, &sp);
    if (loss_cta && tid == 0) {
      a.loss[(unsigned)(a.first_batch + i) % (unsigned)a.plan_len] = loss;
      if (loss != loss) atomicOr(a.flags, FR_FLAG_NAN_LOSS);
    }
    if (tr && tid == 0) a.trace[blockIdx.x * 16 + 3] = ep_now();

    // ---- gradients: the staged chunk (its rows are in registers), then any further ones, over the item- and the
    // user-sorted order; the batch columns come from shared memory
    if (half_rows) ep_consume_rows<true>(ga, cs, x); else ep_consume_rows<false>(ga, cs, x);
    if (!(a.dbg_skip & 4))
      for (int c = gwarp + nwarps; c < 2 * nchunk; c += nwarps) grads_chunk<1, false>(ga, nchunk, c);
    if (tr && tid == 0) a.trace[blockIdx.x * 16 + 4] = ep_now();
    if (!(a.dbg_skip & 64)) ep_arrive(bar);

    // ---- in the shadow of barrier 2: the slot of the step after next is polled for (the barrier's __syncthreads
    // publishes it to the CTA), and Adam of the resident rows this batch does not touch (zero data gradient: nothing of
    // this step is needed, and nobody reads those rows in this step)
    if (ahead2 && i + 2 < a.n_steps) poll_ready(i + 2);
#pragma unroll
    for (int r = 0; r < kEpMaxR; ++r) {
      if (r >= R || (a.dbg_skip & 16)) continue;
      const int4 meta = s_meta[r * kEpThreads + tid];
      if (meta.x == -1 || meta.y >= 0) continue;
      const bool is_item = meta.x < 0;
      float4 p = s_state[(r * 3 + 0) * kEpThreads + tid], m = s_state[(r * 3 + 1) * kEpThreads + tid],
             v = s_state[(r * 3 + 2) * kEpThreads + tid];
      adam1(p.x, m.x, v.x, 0.f, wd, w1, b2, w2, bc2s, eps, neg_step);
      adam1(p.y, m.y, v.y, 0.f, wd, w1, b2, w2, bc2s, eps, neg_step);
      adam1(p.z, m.z, v.z, 0.f, wd, w1, b2, w2, bc2s, eps, neg_step);
      adam1(p.w, m.w, v.w, 0.f, wd, w1, b2, w2, bc2s, eps, neg_step);
      s_state[(r * 3 + 0) * kEpThreads + tid] = p;
      s_state[(r * 3 + 1) * kEpThreads + tid] = m;
      s_state[(r * 3 + 2) * kEpThreads + tid] = v;
      const long long q = (long long)g + (long long)r * NT;
      const long long ql = is_item ? q - nq_u : q;
      *(float4 *)((is_item ? a.I : a.U) + ql * 4) = p;
    }
    if (!(a.dbg_skip & 64)) ep_wait(bar);
    if (tr && tid == 0) a.trace[blockIdx.x * 16 + 5] = ep_now();

    // ---- Adam on the touched resident rows (apply_body's gradient assembly and update).  A popular item's segment
    // spans many chunks: its head partials are brought into shared memory with cp.async first (all in flight at once; the
    // statistics staging area is free now) and then added in chunk order -- the stepwise path's order without one L2
    // round trip per handful of partials.
    if (!(a.dbg_skip & 8)) {
      float4 *stage = (float4 *)fsm;
      const int n_stage = (5 * cap) / 4;
      int sbase[kEpMaxR];
#pragma unroll
      for (int r = 0; r < kEpMaxR; ++r) {
        sbase[r] = -1;
        if (r >= R) continue;
        const int4 meta = s_meta[r * kEpThreads + tid];
        if (meta.x == -1 || meta.y < 0 || meta.w - meta.z < 4) continue;
        const int np = meta.w - meta.z;
        const int base = atomicAdd(&s_alloc, np);
        if (base + np > n_stage) continue;
        sbase[r] = base;
        const bool is_item = meta.x < 0;
        const long long q = (long long)g + (long long)r * NT;
        const int k = (int)((is_item ? q - nq_u : q) % dq) * 4;
        const float *head = is_item ? sl.w.head_i : sl.w.head_u;
        for (int c = 0; c < np; ++c) {
          const unsigned dst = (unsigned)__cvta_generic_to_shared(stage + base + c);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(head + (size_t)(meta.z + 1 + c) * d + k)
                       : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      // (the single-chunk gradients and the tails are loaded while the copies fly)
      float4 gr[kEpMaxR];
#pragma unroll
      for (int r = 0; r < kEpMaxR; ++r) {
        gr[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r >= R) continue;
        const int4 meta = s_meta[r * kEpThreads + tid];
        if (meta.x == -1 || meta.y < 0) continue;
        const bool is_item = meta.x < 0;
        const long long q = (long long)g + (long long)r * NT;
        const int k = (int)((is_item ? q - nq_u : q) % dq) * 4;
        gr[r] = meta.z == meta.w ? __ldcg((const float4 *)((is_item ? sl.w.gseg_i : sl.w.gseg_u) + (size_t)meta.y * d + k))
                                 : __ldcg((const float4 *)((is_item ? sl.w.tail_i : sl.w.tail_u) + (size_t)meta.z * d + k));
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
      for (int r = 0; r < kEpMaxR; ++r) {
        if (r >= R) continue;
        const int4 meta = s_meta[r * kEpThreads + tid];
        if (meta.x == -1 || meta.y < 0) continue;
        const bool is_item = meta.x < 0;
        const long long q = (long long)g + (long long)r * NT;
        const long long ql = is_item ? q - nq_u : q;
        const int k = (int)(ql % dq) * 4;
        const float *head = is_item ? sl.w.head_i : sl.w.head_u;
        float4 gq = gr[r];
        if (meta.z != meta.w) {
          if (sbase[r] >= 0) {
            for (int c = 0; c < meta.w - meta.z; ++c) gq = f4_add(gq, stage[sbase[r] + c]);
          } else {
#pragma unroll 8
            for (int c = meta.z + 1; c <= meta.w; ++c) gq = f4_add(gq, __ldcg((const float4 *)(head + (size_t)c * d + k)));
          }
        }
        float4 p = s_state[(r * 3 + 0) * kEpThreads + tid], m = s_state[(r * 3 + 1) * kEpThreads + tid],
               v = s_state[(r * 3 + 2) * kEpThreads + tid];
        adam1(p.x, m.x, v.x, gq.x, wd, w1, b2, w2, bc2s, eps, neg_step);
        adam1(p.y, m.y, v.y, gq.y, wd, w1, b2, w2, bc2s, eps, neg_step);
        adam1(p.z, m.z, v.z, gq.z, wd, w1, b2, w2, bc2s, eps, neg_step);
        adam1(p.w, m.w, v.w, gq.w, wd, w1, b2, w2, bc2s, eps, neg_step);
        s_state[(r * 3 + 0) * kEpThreads + tid] = p;
        s_state[(r * 3 + 1) * kEpThreads + tid] = m;
        s_state[(r * 3 + 2) * kEpThreads + tid] = v;
        *(float4 *)((is_item ? a.I : a.U) + ql * 4) = p;
      }
    }
    if (tr && tid == 0) a.trace[blockIdx.x * 16 + 6] = ep_now();
    ep_arrive(bar);
    // ---- in the shadow of barrier 3: the mailbox of the step after next, the next step's row stamps / ids / gradient chunk
    if (i + 1 < a.n_steps) {
      if (ahead2) {
        cur = nx;
        if (i + 2 < a.n_steps) {
          if (a.dbg_skip & 64) wait_ready(i + 2);   // (timing experiments without barrier 2: no poll happened there)
          load_mail(i + 2, nx);
        }
      } else {
        wait_ready(i + 1);
        load_mail(i + 1, cur);
      }
      fetch_rows(i + 1, cur.stamp);
      cs = stage_first(i + 1, cur.B);
    }
    ep_wait(bar);
    if (cta == 0 && tid == 0) st_rel(a.sync + 1, (unsigned long long)(i + 1));
    if (tr && tid == 0) a.trace[blockIdx.x * 16 + 7] = ep_now();
  }

  // the moments go back to their tables (the parameters are current in global memory after every step)
  for (int r = 0; r < R; ++r) {
    const int4 meta = s_meta[r * kEpThreads + tid];
    if (meta.x == -1) continue;
    const bool is_item = meta.x < 0;
    const long long q = (long long)g + (long long)r * NT;
    const long long ql = is_item ? q - nq_u : q;
    *(float4 *)((is_item ? a.mI : a.mU) + ql * 4) = s_state[(r * 3 + 1) * kEpThreads + tid];
    *(float4 *)((is_item ? a.vI : a.vU) + ql * 4) = s_state[(r * 3 + 2) * kEpThreads + tid];
  }
}

// ------------------------------------------------------------------------------------------ host side
static unsigned long long *g_epoch_trace = nullptr;
static int g_epoch_trace_step = -1;

struct EpPlan {
  int grid, n_prod, R, cap;
  size_t smem;
};

static int ep_plan(const fr_focf_step *slots, int n_slots, EpPlan *out, const char *who) {
  static int n_sm = 0, smem_max = 0;
  if (!n_sm) {
    int dev = 0;
    FR_CUDA_OK(cudaGetDevice(&dev));
    FR_CUDA_OK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    FR_CUDA_OK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  }
  const fr_focf_step *s = &slots[0];
  out->n_prod = 2 * n_slots;
  out->grid = n_sm;
  const int n_comp = n_sm - out->n_prod;
  if (n_comp < 8) {
    set_error("%s: %d SMs leave no room for %d producer CTAs", who, n_sm, out->n_prod);
    return FR_ERR_UNSUPPORTED;
  }
  const long long nq = ((long long)s->n_users + s->n_items) * (s->d / 4);
  const long long NT = (long long)n_comp * kEpThreads;
  out->R = (int)((nq + NT - 1) / NT);
  out->cap = s->B;
  const size_t sm_c = ep_comp_smem(s->B, out->R), sm_p = ep_prod_smem(s->B);
  out->smem = sm_c > sm_p ? sm_c : sm_p;
  if (out->R > kEpMaxR || out->smem + 1024 > (size_t)smem_max) {
    set_error("%s: tables of %lld parameters and batches of up to %d rows need %zu bytes of shared memory per CTA (max %d)",
              who, nq * 4, s->B, out->smem, smem_max);
    return FR_ERR_UNSUPPORTED;
  }
  return FR_OK;
}

static int ep_check(const fr_focf_step *slots, int n_slots, const char *who) {
  FR_REQUIRE(slots && n_slots >= 2 && n_slots <= kEpMaxSlots, "%s: 2..%d slots expected", who, kEpMaxSlots);
  const fr_focf_step *s0 = &slots[0];
  for (int k = 0; k < n_slots; ++k) {
    const fr_focf_step *s = &slots[k];
    FR_REQUIRE(s->U && s->I && s->mU && s->vU && s->mI && s->vI && s->uid && s->iid && s->rating && s->sst && s->pred &&
                   s->loss && s->status_flags && s->workspace,
               "%s: slot %d: null pointer in fr_focf_step", who, k);
    FR_REQUIRE(s->objective >= FR_OBJ_NONE && s->objective <= FR_OBJ_NONPARITY, "%s: bad objective %d", who, s->objective);
    FR_REQUIRE(s->plan_desc && s->plan_items && s->plan_offs && s->plan_len >= 1 && s->item_off && s->train_uid &&
                   s->train_rating && s->sst_of_user,
               "%s: slot %d: a planned epoch is required", who, k);
    FR_REQUIRE(s->U == s0->U && s->I == s0->I && s->mU == s0->mU && s->vU == s0->vU && s->mI == s0->mI && s->vI == s0->vI &&
                   s->n_users == s0->n_users && s->n_items == s0->n_items && s->d == s0->d && s->B == s0->B &&
                   s->plan_desc == s0->plan_desc && s->plan_len == s0->plan_len && s->loss == s0->loss &&
                   s->objective == s0->objective && s->fair_weight == s0->fair_weight &&
                   s->status_flags == s0->status_flags,
               "%s: slot %d differs from slot 0 in more than its batch columns and workspace", who, k);
    for (int j = 0; j < k; ++j)
      FR_REQUIRE(slots[j].workspace != s->workspace && slots[j].uid != s->uid && slots[j].pred != s->pred,
                 "%s: slots %d and %d share a workspace or batch columns", who, j, k);
  }
  FR_REQUIRE(s0->d % 4 == 0 && s0->d >= 4 && s0->d <= 128, "%s: embedding size %d (multiple of 4, <= 128)", who, s0->d);
  FR_REQUIRE(s0->B >= 1 && s0->B <= kEpMaxRows, "%s: batch capacity %d (<= %d rows)", who, s0->B, kEpMaxRows);
  FR_REQUIRE(s0->adam_mode == FR_ADAM_DENSE_EXACT && s0->norm_B == 0 && s0->norm_J == 0 && !s0->norm_dev,
             "%s: dense_exact Adam on one GPU only", who);
  return FR_OK;
}

}  // namespace fr

extern "C" {

int fr_focf_epoch_eligible(const fr_focf_step *slots, int32_t n_slots) {
  if (fr::ep_check(slots, n_slots, "fr_focf_epoch_eligible")) return 0;
  fr::EpPlan p;
  return fr::ep_plan(slots, n_slots, &p, "fr_focf_epoch_eligible") == FR_OK ? 1 : 0;
}

int fr_focf_epoch_run(const fr_focf_step *slots, int32_t n_slots, int32_t first_batch, int32_t n_steps,
                      int32_t adam_step, void *sync_words, void *stream) {
  using namespace fr;
  if (n_steps == 0) return FR_OK;
  int rc = ep_check(slots, n_slots, "fr_focf_epoch_run");
  if (rc) return rc;
  FR_REQUIRE(first_batch >= 0 && n_steps > 0 && adam_step >= 1 && sync_words, "fr_focf_epoch_run: bad step range or null sync words");
  EpPlan pl;
  if ((rc = ep_plan(slots, n_slots, &pl, "fr_focf_epoch_run"))) return rc;
  const fr_focf_step *s = &slots[0];
  const int max_steps = 4 * s->B;   // the scalars table lives in slot 0's record scratch (8 floats per row of capacity)
  if (n_steps > max_steps) {        // (longer runs: one launch per max_steps steps)
    for (int k = 0; k < n_steps; k += max_steps) {
      const int n = n_steps - k < max_steps ? n_steps - k : max_steps;
      if ((rc = fr_focf_epoch_run(slots, n_slots, first_batch + k, n, adam_step + k, sync_words, stream))) return rc;
    }
    return FR_OK;
  }
  EpochArgs a{};
  a.U = s->U; a.I = s->I; a.mU = s->mU; a.vU = s->vU; a.mI = s->mI; a.vI = s->vI;
  a.n_users = s->n_users; a.n_items = s->n_items; a.d = s->d;
  a.plan_desc = s->plan_desc; a.plan_items = s->plan_items; a.plan_offs = s->plan_offs; a.plan_len = s->plan_len;
  a.item_off = s->item_off; a.train_uid = s->train_uid; a.train_rating = s->train_rating; a.sst_of_user = s->sst_of_user;
  a.objective = s->objective; a.fair_weight = s->fair_weight; a.loss = s->loss; a.flags = s->status_flags;
  a.lr = s->lr; a.beta1 = s->beta1; a.beta2 = s->beta2; a.eps = s->eps; a.wd = s->weight_decay;
  a.first_batch = first_batch; a.n_steps = n_steps; a.adam_step0 = adam_step;
  a.n_slots = n_slots; a.n_prod = pl.n_prod; a.cap = pl.cap; a.chunk = grad_chunk(s->B);
  a.key_bits_u = bits_for((uint32_t)s->n_users); a.R = pl.R;
  a.sync = (unsigned long long *)sync_words;
  static int trace_on = -1;
  if (trace_on < 0) {
    const char *e = getenv("FR_FOCF_TRACE");
    trace_on = (e && e[0] == '1') ? 1 : 0;
    if (trace_on) {
      FR_CUDA_OK(cudaMalloc(&g_epoch_trace, sizeof(unsigned long long) * 16 * 256));
      const char *ts = getenv("FR_FOCF_TRACE_STEP");
      g_epoch_trace_step = ts ? atoi(ts) : 8;
    }
  }
  static int dbg_skip = -1;
  if (dbg_skip < 0) {
    const char *e = getenv("FR_FOCF_EPOCH_SKIP");
    dbg_skip = e ? atoi(e) : 0;
  }
  a.dbg_skip = dbg_skip;
  a.trace = trace_on ? g_epoch_trace : nullptr;
  a.trace_step = g_epoch_trace_step < n_steps ? g_epoch_trace_step : n_steps - 1;
  for (int k = 0; k < n_slots; ++k) {
    const fr_focf_step *sk = &slots[k];
    Carver c(sk->workspace, sk->workspace_bytes);
    a.slot[k].w = carve(c, sk->n_users, sk->n_items, sk->d, sk->B);
    if (!c.ok()) {
      set_error("fr_focf_epoch_run: slot %d: workspace too small (%zu < %zu bytes)", k, sk->workspace_bytes, c.off);
      return FR_ERR_WORKSPACE;
    }
    a.slot[k].uid = (int32_t *)sk->uid; a.slot[k].iid = (int32_t *)sk->iid;
    a.slot[k].rating = (float *)sk->rating; a.slot[k].sst = (float *)sk->sst; a.slot[k].pred = sk->pred;
  }
  cudaStream_t st = (cudaStream_t)stream;
  {   // sc_tab[2k], [2k+1] = the scalars of optimizer step adam_step + k
    float *tab = a.slot[0].w.rec_seg;
    FR_LAUNCH(k_adam_scalars, grid_for(n_steps, 128, 64), 128, 0, st, tab - 2 * (ptrdiff_t)adam_step, adam_step,
              adam_step + n_steps - 1, s->lr, s->beta1, s->beta2);
    a.sc_tab = tab;
  }
  static size_t smem_set = 0;
  if (pl.smem > smem_set) {
    FR_CUDA_OK(cudaFuncSetAttribute(k_focf_epoch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    smem_set = pl.smem;
  }
  int per_sm = 0;
  FR_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_focf_epoch, kEpThreads, pl.smem));
  if (per_sm < 1) {
    set_error("fr_focf_epoch_run: the epoch kernel does not fit an SM with %zu bytes of shared memory", pl.smem);
    return FR_ERR_UNSUPPORTED;
  }
  // barrier / completion counters and the slots' ready words start from zero
  FR_CUDA_OK(cudaMemsetAsync(sync_words, 0, 64, st));
  for (int k = 0; k < n_slots; ++k)
    FR_CUDA_OK(cudaMemsetAsync(a.slot[k].w.ctrl + EP_READY_U, 0, 2 * sizeof(uint32_t), st));
  void *args[] = {&a};
  const bool p = prof_on();
  if (p) prof_begin("k_focf_epoch", st);
  cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_focf_epoch, dim3(pl.grid), dim3(kEpThreads), args, pl.smem, st);
  if (p) prof_end(st);
  count_launch();
  if (e != cudaSuccess) {
    set_error("cudaLaunchCooperativeKernel(k_focf_epoch) failed: %s", cudaGetErrorString(e));
    return FR_ERR_CUDA;
  }
  return FR_OK;
}

int fr_focf_epoch_trace(uint64_t *out_host, int32_t n) {
  FR_REQUIRE(out_host && n >= 0 && n <= 16 * 256, "fr_focf_epoch_trace: bad argument");
  FR_REQUIRE(fr::g_epoch_trace, "fr_focf_epoch_trace: set FR_FOCF_TRACE=1 before the first epoch launch");
  FR_CUDA_OK(cudaDeviceSynchronize());
  FR_CUDA_OK(cudaMemcpy(out_host, fr::g_epoch_trace, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToHost));
  return FR_OK;
}

}  // extern "C"
