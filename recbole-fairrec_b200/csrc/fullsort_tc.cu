// Full-sort scoring on the 5th-generation tensor cores: tcgen05.mma kind::tf32 fed by TMA, accumulators in TMEM,
// 3xTF32 split for fp32-level accuracy, fused with the pad/history mask and the streaming top-K.
//
// Reference being replaced: focf.py:171-178 (mm + clamp/max_rating), trainer.py:435-438 (masks),
// collector.py:143-153 (topk).  This is FR_SCORE_TC_3XTF32; FR_SCORE_EXACT_FP32 (fullsort_eval.cu) is the
// bit-defined mode the parity tests pin, this mode is checked against it with the near-tie protocol.
//
// 3xTF32: x = hi + lo with hi = tf32(x) (round to nearest) and lo = tf32(x - hi) (x - hi is exact in fp32).  U.I^T ~= Uh.Ih + Uh.Il + Ul.Ih, three MMAs per k-step into one fp32 TMEM
// accumulator; the dropped lo.lo term is 2^-22 relative.  The planes are split ONCE per evaluation (k_split_planes),
// not per tile.
//
// CTA = 128 eval users x a contiguous range of item tiles (128 items each); 6 warps:
//   warp 0   TMA producer: streams the item K-blocks (32 k-columns = one 128-byte swizzle atom row) of Ih and Il through a
//            shared-memory ring of up to 6 stages (cp.async.bulk.tensor, mbarrier complete_tx)
//   warp 1   MMA issuer: tcgen05.mma (M=128, N=128, K=8) x 4 k-steps x 3 products per K-block, A operand from TENSOR
//            MEMORY, B from the ring; tcgen05.commit releases the ring slot and, after the last K-block of a tile,
//            publishes the accumulator
//   warp 2-5 epilogue: TMEM lane == user row, so each thread OWNS one user: it first stores its user's two planes into its
//            TMEM lane (tcgen05.st: the A operand of every MMA of the CTA), then per tile tcgen05.ld 32 columns at a time,
//            mask bits from the thread's own walk of its sorted history, a one-compare filter on the raw dot against the
//            raw value of the row's current K-th best, and (rarely) the exact transform + insertion into the row's K-list.
//            No cross-thread synchronisation in the epilogue at all.
// TMEM (512 columns): two accumulators (2 x 128) so the MMAs of tile t+1 overlap the epilogue of tile t, then the user
// tile's hi and lo planes (2 x d <= 256 columns).
// Both issuing warps run their loops warp-uniformly and elect one lane per issue (elect_one(), tc_common.cuh): with the
// loop under `if (lane == 0)` every UTCHMMA paid an ELECT / R2UR.BROADCAST waterfall and the kernel was ISSUE-bound at
// ~106 cycles per 64-cycle MMA (74 % of the 3xTF32 peak; the same time with the item loads or the epilogue switched off,
// FR_TC_DIAG).  Uniform issue: 98 %; A in tensor memory (an SS dispatch reads 8 KB of shared memory per 64 cycles = all
// of the SM's 128 B/clk): 105-107 % of the figure derived from the measured bf16 peak, MMA-only floor 7.3 ms for the
// 8.7 ms of profiles/tools/time_tc_big.py's slab (profiles/r02_tc_scorer.md).
#include "tc_common.cuh"

namespace fr {

constexpr int TC_MAX_STAGES = 6;
constexpr int TC_THREADS = 192;
constexpr int TC_KBLOCK_BYTES = TCN * TCKB * 4;   // 16 KB per plane per K-block
constexpr int TC_ACC_COLS = 2 * TCN;               // TMEM: two accumulators, then the user tile's planes (2 x d columns)

// ---------------------------------------------------------------- the scorer
struct TcArgs {
  const float *uh, *ul;   // the eval users' TF32 planes, row-major [n, d] (k_split_planes)
  const int64_t *hist_off;
  const int32_t *hist_items;
  int n, d, n_items_local, item_base, K, transform;
  float max_rating;
  int tiles_per_split;
  int stages;         // ring depth (2..4, what fits beside the resident user planes)
  int diag;           // timing knob FR_TC_DIAG: bit 0 = no item loads (stale ring), bit 1 = no epilogue reads; results invalid
  int use_scratch;    // 1: per-thread shared-memory scratch rows for the slow path (fast); 0: collective TMEM re-read
  int32_t *out_id;    // [n_splits, n, K]
  float *out_score;
};

__device__ __forceinline__ bool tc_better(float s, int id, float s2, int id2) { return s > s2 || (s == s2 && id < id2); }

__device__ __forceinline__ float tc_transform(float x, int transform, float max_rating) {
  if (transform == FR_TRANSFORM_CLAMP_DIV) return __fdiv_rn(fminf(fmaxf(x, 0.f), max_rating), max_rating);
  if (transform == FR_TRANSFORM_SIGMOID) return 1.f / (1.f + expf(-x));
  return x;
}

// A raw dot that certainly cannot beat a K-th best of transformed score thr_s (ids only grow while a CTA streams its
// item range, so an equal score loses the tie): the transform is monotone, so raw <= bound  =>  s <= thr_s.
__device__ __forceinline__ float tc_raw_bound(float thr_s, int transform, float max_rating) {
  if (thr_s == -INFINITY) return -INFINITY;
  if (transform == FR_TRANSFORM_NONE) return thr_s;
  if (transform == FR_TRANSFORM_CLAMP_DIV) {
    if (thr_s >= 1.f) return INFINITY;                    // saturated: clamp() cannot exceed 1
    const float b = thr_s * max_rating;
    return b - fabsf(b) * 4e-7f;                           // a few ulp below: the exact test decides inside the band
  }
  if (thr_s > 0.f && thr_s < 1.f) {                        // sigmoid
    const float b = logf(thr_s / (1.f - thr_s));
    return b - (fabsf(b) + 1.f) * 1e-4f;
  }
  return -INFINITY;
}

// kRegList (K <= 16): the row's K-list lives in registers and an insertion is a branch-free compare / select network
// (16 compares, 32 selects, no memory round trips); otherwise it lives in shared memory and an insertion is a shift loop of
// dependent LDS / STS pairs -- 40 % of the executed instructions and most of the latency at the ML-1M shape, where a
// catalogue of 3,707 items never lets the threshold warm up.  Same total order, same lists.
// kKReg = register slots of the list (10 for the usual K <= 10, 16 for K <= 16: the network's cost is proportional to the
// slot count), 0 = shared-memory list.
constexpr int TC_KREG_MAX = 16;
template <int kKReg>
__global__ void __launch_bounds__(TC_THREADS, 1)
    k_fullsort_tc(const __grid_constant__ CUtensorMap map_ih, const __grid_constant__ CUtensorMap map_il, TcArgs a) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  const int nkb = a.d / TCKB;                                 // K-blocks per row (d = 32, 64, 96 or 128)
  // carve: [ring: TC_STAGES x (B hi 16 KB + B lo 16 KB)][lists][scratch][barriers]   (the user tile lives in TMEM)
  unsigned char *sB = tc_smem;
  const int TC_STAGES = a.stages;
  constexpr bool kRegList = kKReg > 0;
  constexpr int TC_KREG = kRegList ? kKReg : 1;
  const int list_k = kRegList ? 0 : a.K;                             // (the register list needs no shared-memory rows)
  float *list_s = (float *)(sB + TC_STAGES * 2 * TC_KBLOCK_BYTES);   // [TCM][K]
  int *list_i = (int *)(list_s + TCM * list_k);                      // [TCM][K]
  float *scr = (float *)(list_i + TCM * list_k);                     // [TCM][33] candidate scratch rows (optional)
  uint64_t *bars = (uint64_t *)(((uintptr_t)(scr + (a.use_scratch ? TCM * 33 : 0)) + 7) & ~(uintptr_t)7);
  uint64_t *full = bars, *empty = bars + TC_MAX_STAGES, *tfull = bars + 2 * TC_MAX_STAGES, *tempty = tfull + 2,
           *afull = tempty + 2;
  uint32_t *tmem_slot = (uint32_t *)(afull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u0 = blockIdx.x * TCM;
  const int split = blockIdx.y;
  const int n_tiles_total = (a.n_items_local + TCN - 1) / TCN;
  const int tile_lo = split * a.tiles_per_split;
  const int tile_hi = min(n_tiles_total, tile_lo + a.tiles_per_split);
  const int ntile = max(0, tile_hi - tile_lo);

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], 128);   // every epilogue thread arrives
    }
    mbar_init(afull, 128);          // every epilogue thread has stored its user row into TMEM
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: 2 accumulators x 128 columns + the user tile's two planes (2 x d <= 256 columns): all 512
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================== TMA producer (the whole warp walks the ring, one elected
    // lane issues: see elect_one())
    int s = 0;
    uint32_t ph = 0;
    for (int tt = 0; tt < ntile; ++tt) {
      const int row0 = (tile_lo + tt) * TCN;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty[s], ph ^ 1);
        if (elect_one()) {
          if (a.diag & 1) {
            mbar_arrive(&full[s]);
          } else {
            mbar_expect_tx(&full[s], 2u * TC_KBLOCK_BYTES);
            tma_load_2d(sB + (2 * s) * TC_KBLOCK_BYTES, &map_ih, kb * TCKB, row0, &full[s]);
            tma_load_2d(sB + (2 * s + 1) * TC_KBLOCK_BYTES, &map_il, kb * TCKB, row0, &full[s]);
          }
        }
        __syncwarp();
        if (++s == TC_STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (warp-uniform loop, one elected lane issues)
    const uint32_t idesc = umma_idesc_tf32(TCM, TCN);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);   // (uniform for the compiler)
    mbar_wait(afull, 0);
    tc_fence_after();
    int s = 0;
    uint32_t ph = 0;
    for (int tt = 0; tt < ntile; ++tt) {
      const int buf = tt & 1;
      const uint32_t tph = (tt >> 1) & 1;
      mbar_wait(&tempty[buf], tph ^ 1);      // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_u + (uint32_t)(buf * TCN);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_hi = tmem_u + (uint32_t)(TC_ACC_COLS + kb * TCKB), a_lo = a_hi + (uint32_t)a.d;
          const uint64_t b_hi = umma_desc_sw128(smem_u32(sB + (2 * s) * TC_KBLOCK_BYTES));
          const uint64_t b_lo = umma_desc_sw128(smem_u32(sB + (2 * s + 1) * TC_KBLOCK_BYTES));
#pragma unroll
          for (int k = 0; k < TCKB / 8; ++k) {   // UMMA K = 8 tf32 = 32 bytes along the swizzled row = +2 in the descriptor's
                                                 // 16-byte start-address units (no carry: the blocks are 1 KB aligned)
            const uint64_t off = (uint64_t)(k * 2);
            umma_tf32_ts(tmem_d, a_hi + k * 8, b_hi + off, idesc, (kb | k) ? 1u : 0u);
            umma_tf32_ts(tmem_d, a_lo + k * 8, b_hi + off, idesc, 1u);
            umma_tf32_ts(tmem_d, a_hi + k * 8, b_lo + off, idesc, 1u);
          }
          umma_commit(&empty[s]);              // ring slot reusable once these MMAs have read it
          if (kb == nkb - 1) umma_commit(&tfull[buf]);   // accumulator complete
        }
        __syncwarp();
        if (++s == TC_STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else {
    // ===================================================== epilogue: thread == user row
    const int quarter = warp & 3;               // TMEM lanes [32*quarter, +32) are this warp's
    const int row = quarter * 32 + lane;
    const int r = u0 + row;
    const bool live = r < a.n;
    const int K = a.K;
    float *ls = list_s + row * K;
    int *li = list_i + row * K;
    float *scratch = scr + row * 33;
    {
      // this thread's user row (both planes) into its TMEM lane: the A operand of every MMA of this CTA.  A in tensor
      // memory halves the shared-memory operand traffic of an MMA (8 KB per 64-cycle M=128 N=128 K=8 dispatch is the whole
      // 128 B/clk of the SM: 75-82 cycles per MMA measured, profiles/tools/mma_rate.cu) and frees 2 x d x 512 bytes of shared
      // memory for ring stages.
      const uint32_t ta = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)TC_ACC_COLS;
      for (int pl = 0; pl < 2; ++pl) {
        const float4 *src = (const float4 *)((pl ? a.ul : a.uh) + (size_t)(live ? r : 0) * a.d);
        for (int c = 0; c < nkb; ++c) {
          uint32_t v[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 x = live ? __ldg(src + c * 8 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[4 * j] = __float_as_uint(x.x);
            v[4 * j + 1] = __float_as_uint(x.y);
            v[4 * j + 2] = __float_as_uint(x.z);
            v[4 * j + 3] = __float_as_uint(x.w);
          }
          tmem_st32(ta + (uint32_t)(pl * a.d + c * TCKB), v);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(afull);
    }
    float rs[TC_KREG];
    int ri[TC_KREG];
    // (right-aligned: slot TC_KREG - 1 is the K-th best, so the threshold is a fixed register; the unused leading slots
    // hold an entry that everything loses against)
#pragma unroll
    for (int e = 0; e < TC_KREG; ++e) {
      rs[e] = e < TC_KREG - K ? INFINITY : -INFINITY;
      ri[e] = e < TC_KREG - K ? -1 : 0x7fffffff;
    }
    if (!kRegList) {
      for (int e = 0; e < K; ++e) {
        ls[e] = -INFINITY;
        li[e] = 0x7fffffff;
      }
    }
    float raw_thr = -INFINITY;                  // one-compare filter on the raw dot (tc_raw_bound), -inf while not full
    float thr_s = -INFINITY;
    int thr_i = 0x7fffffff;
    auto try_insert = [&](float raw, int gid) {   // exact transform + total-order insertion into the row's K-list
      const float s = tc_transform(raw, a.transform, a.max_rating);
      if (kRegList) {
        if (tc_better(s, gid, thr_s, thr_i)) {
          // b[e]: the new entry goes before entry e (monotone over the sorted list: false ... false true ... true);
          // entry e becomes entry e-1 where b[e-1], the new entry where b[e] && !b[e-1], itself otherwise
          bool b[TC_KREG];
#pragma unroll
          for (int e = 0; e < TC_KREG; ++e) b[e] = tc_better(s, gid, rs[e], ri[e]);
#pragma unroll
          for (int e = TC_KREG - 1; e >= 1; --e) {
            rs[e] = b[e - 1] ? rs[e - 1] : (b[e] ? s : rs[e]);
            ri[e] = b[e - 1] ? ri[e - 1] : (b[e] ? gid : ri[e]);
          }
          rs[0] = b[0] ? s : rs[0];
          ri[0] = b[0] ? gid : ri[0];
          thr_s = rs[TC_KREG - 1];
          thr_i = ri[TC_KREG - 1];
          if (thr_i != 0x7fffffff) raw_thr = tc_raw_bound(thr_s, a.transform, a.max_rating);
        }
      } else if (tc_better(s, gid, thr_s, thr_i)) {
        int p = K - 1;
        while (p > 0 && tc_better(s, gid, ls[p - 1], li[p - 1])) {
          ls[p] = ls[p - 1];
          li[p] = li[p - 1];
          --p;
        }
        ls[p] = s;
        li[p] = gid;
        thr_s = ls[K - 1];
        thr_i = li[K - 1];
        if (thr_i != 0x7fffffff) raw_thr = tc_raw_bound(thr_s, a.transform, a.max_rating);
      }
    };
    long long hp = 0, hend = 0;
    if (live) {
      long long lo_ = a.hist_off[r], hi_ = a.hist_off[r + 1];
      hend = hi_;
      const int first = a.item_base + tile_lo * TCN;
      while (lo_ < hi_) {
        const long long mid = (lo_ + hi_) >> 1;
        if (a.hist_items[mid] < first) lo_ = mid + 1; else hi_ = mid;
      }
      hp = lo_;
    }
    int hw[8];
    bool hw_valid = false;
    for (int tt = 0; tt < ntile; ++tt) {
      const int buf = tt & 1;
      const uint32_t tph = (tt >> 1) & 1;
      const int tile_base = (tile_lo + tt) * TCN;
      const int g0 = a.item_base + tile_base;
      // mask words of this tile from the thread's own sorted history (+ the [PAD] item, global id 0)
      uint32_t mask[TCN / 32];
#pragma unroll
      for (int w = 0; w < TCN / 32; ++w) mask[w] = 0u;
      if (g0 == 0) mask[0] |= 1u;
      // the row's own sorted history, eight entries per round trip (a one-entry walk is a chain of dependent loads: 21 % of
      // the warp samples at the ML-1M shape, where a user has ~6 history items per tile)
      // (the window is carried across tiles: a tile without history items -- nearly all of them for a large catalogue --
      // costs eight register compares and no load)
      while (true) {
        if (!hw_valid) {
#pragma unroll
          for (int j = 0; j < 8; ++j) hw[j] = (hp + j < hend) ? a.hist_items[hp + j] : 0x7fffffff;
          hw_valid = true;
        }
        int used = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (hw[j] < g0 + TCN) {   // (ascending: the entries of this tile are a prefix of the window)
            used = j + 1;
            if (hw[j] >= g0) {
              const int c = hw[j] - g0;
#pragma unroll
              for (int w = 0; w < TCN / 32; ++w)
                if ((c >> 5) == w) mask[w] |= 1u << (c & 31);
            }
          }
        }
        hp += used;
        if (used) hw_valid = false;
        if (used < 8) break;
      }
      mbar_wait(&tfull[buf], tph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * TCN);
      for (int w = 0; w < ((a.diag & 2) ? 0 : TCN / 32); ++w) {
        uint32_t v[32];
        tmem_ld32(taddr + w * 32, v);
        // fast path: one compare per element -> candidate bit mask (branch free, ~2 instructions per score)
        uint32_t cand = 0u;
#pragma unroll
        for (int c = 0; c < 32; ++c) cand |= (__uint_as_float(v[c]) > raw_thr ? 1u : 0u) << c;
        const int cols = a.n_items_local - (tile_base + w * 32);             // valid columns of this chunk
        const uint32_t colmask = cols >= 32 ? 0xffffffffu : (cols <= 0 ? 0u : ((1u << cols) - 1u));
        const uint32_t mw = w == 0 ? mask[0] : (w == 1 ? mask[1] : (w == 2 ? mask[2] : mask[3]));
        cand &= colmask & ~mw;
        if (!live) cand = 0u;
        // slow path: rare once the row's threshold has warmed up; ONE copy of the insertion code (instruction footprint)
        if (a.use_scratch) {
          if (cand) {   // park the 32 raw scores in this thread's scratch row and walk the candidate bits
#pragma unroll
            for (int c = 0; c < 32; ++c) scratch[c] = __uint_as_float(v[c]);
            while (cand) {
              const int c = __ffs(cand) - 1;
              cand &= cand - 1u;
              const float raw = scratch[c];
              if (raw > raw_thr) try_insert(raw, g0 + w * 32 + c);
            }
          }
        } else if (kRegList) {
          // no scratch rows (their 17 KB buy one more ring stage at d = 128): the candidate's score is picked out of the
          // registers by a select chain -- 32 selects per candidate, which a large catalogue makes rare
          while (cand) {
            const int c = __ffs(cand) - 1;
            cand &= cand - 1u;
            uint32_t x = v[0];
#pragma unroll
            for (int j = 1; j < 32; ++j) x = (j == c) ? v[j] : x;
            const float raw = __uint_as_float(x);
            if (raw > raw_thr) try_insert(raw, g0 + w * 32 + c);
          }
        } else {
          // no room for scratch rows: warp-uniform walk over the union of the lanes' candidate columns, each column
          // re-read for all 32 rows with a collective tcgen05.ld x1
          uint32_t uni = cand;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) uni |= __shfl_xor_sync(0xffffffffu, uni, o);
          while (uni) {
            const int c = __ffs(uni) - 1;
            uni &= uni - 1u;
            uint32_t x;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(x) : "r"(taddr + w * 32 + c));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const float raw = __uint_as_float(x);
            if (((cand >> c) & 1u) && raw > raw_thr) try_insert(raw, g0 + w * 32 + c);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[buf]);
    }
    if (live) {
      if (kRegList) {
#pragma unroll
        for (int e = 0; e < TC_KREG; ++e) {
          if (e >= TC_KREG - K) {
            const size_t o = ((size_t)split * a.n + r) * K + (e - (TC_KREG - K));
            a.out_id[o] = ri[e];
            a.out_score[o] = rs[e];
          }
        }
      } else {
        for (int e = 0; e < K; ++e) {
          const size_t o = ((size_t)split * a.n + r) * K + e;
          a.out_id[o] = li[e];
          a.out_score[o] = ls[e];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---------------------------------------------------------------- host side
size_t tc_plane_bytes(int n, int n_items_local, int d) {
  const size_t a = (((size_t)n * d * 4) + 255) & ~(size_t)255, b = (((size_t)n_items_local * d * 4) + 255) & ~(size_t)255;
  return 2 * a + 2 * b;
}

// Item splits per user tile: one CTA per SM (225 KB of shared memory), so the launch runs in waves of kSMs CTAs and a CTA
// walks ceil(itiles / s) item tiles: minimise waves x tiles per CTA (the smallest s on ties: the merge grows with s).
// (ceil(kSMs / utiles) made 188 CTAs = two waves of 8 tiles at the ML-1M shape; 3 splits are one wave of 10.)
int tc_pick_splits(int n, int n_items_local) {
  const int utiles = (n + TCM - 1) / TCM, itiles = (n_items_local + TCN - 1) / TCN;
  if (const char *e = getenv("FR_TC_SPLITS")) {   // (timing knob)
    const int want = atoi(e);
    if (want >= 1 && want <= 32 && want <= itiles) return want;
  }
  int best = 1;
  long long best_cost = -1;
  for (int s = 1; s <= 32 && s <= itiles; ++s) {
    const long long waves = ((long long)utiles * s + kSMs - 1) / kSMs, per = (itiles + s - 1) / s;
    const long long cost = waves * per;
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = s;
    }
  }
  return best;
}

// planes: workspace region of tc_plane_bytes(); part_id/part_sc: [splits, n, K] (or the outputs when splits == 1)
int tc_launch(const fr_fullsort *a, void *planes, int splits, int32_t *out_id, float *out_sc, cudaStream_t st) {
  const int n = a->n, d = a->d, nl = a->n_items_local;
  const size_t ab = (((size_t)n * d * 4) + 255) & ~(size_t)255, bb = (((size_t)nl * d * 4) + 255) & ~(size_t)255;
  float *Uh = (float *)planes, *Ul = (float *)((char *)planes + ab);
  float *Ih = (float *)((char *)planes + 2 * ab), *Il = (float *)((char *)planes + 2 * ab + bb);
  FR_LAUNCH(k_split_planes2, grid_for(((int64_t)n + nl) * d / 4, 256, kSMs * 16), 256, 0, st, a->U, a->users, (int64_t)n, Uh, Ul,
            a->I_shard, (int64_t)nl, Ih, Il, d);
  CUtensorMap m_ih, m_il;
  if (!make_map(&m_ih, Ih, nl, d) || !make_map(&m_il, Il, nl, d)) {
    set_error("fr_fullsort_topk: cuTensorMapEncodeTiled failed");
    return FR_ERR_CUDA;
  }
  const bool reg_list = a->K <= TC_KREG_MAX;
  size_t fixed = (reg_list ? 0 : (size_t)TCM * a->K * 8) + (size_t)TCM * 33 * 4 + 256;
  int use_scratch = 1;
  int stages = (int)((232448 - 1024 - (long long)fixed) / (2 * TC_KBLOCK_BYTES));   // 227 KB dynamic smem, 1 KB slack
  // drop the scratch rows where that buys a ring stage below the full depth or where nothing fits otherwise
  const int stages_ns = (int)((232448 - 1024 - (long long)(fixed - (size_t)TCM * 33 * 4)) / (2 * TC_KBLOCK_BYTES));
  if (stages < TC_MAX_STAGES && stages_ns > stages && (reg_list || stages < 2)) {
    use_scratch = 0;
    fixed -= (size_t)TCM * 33 * 4;
    stages = stages_ns;
  }
  if (const char *e = getenv("FR_TC_STAGES")) {   // (timing knob: profiles/tools/time_tc_big.py)
    const int want = atoi(e);
    if (want >= 2 && want <= stages) stages = want;
  }
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  if (stages < 2) {
    set_error("fr_fullsort_topk: tensor-core scorer does not fit shared memory for d=%d K=%d", d, a->K);
    return FR_ERR_UNSUPPORTED;
  }
  const size_t smem = fixed + (size_t)stages * 2 * TC_KBLOCK_BYTES + 1024;
  const int itiles = (nl + TCN - 1) / TCN;
  TcArgs t{Uh, Ul, a->hist_off, a->hist_items, n, d, nl, a->item_base, a->K, a->transform, a->max_rating,
           (itiles + splits - 1) / splits, stages, getenv("FR_TC_DIAG") ? atoi(getenv("FR_TC_DIAG")) : 0, use_scratch, out_id, out_sc};
  dim3 grid((n + TCM - 1) / TCM, splits);
  const bool prof = prof_on();   // (one profiler name for both instantiations)
  if (prof) prof_begin("k_fullsort_tc", st);
  if (a->K <= 10) {
    FR_CUDA_OK(cudaFuncSetAttribute(k_fullsort_tc<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_fullsort_tc<10><<<grid, TC_THREADS, smem, st>>>(m_ih, m_il, t);
  } else if (reg_list) {
    FR_CUDA_OK(cudaFuncSetAttribute(k_fullsort_tc<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_fullsort_tc<16><<<grid, TC_THREADS, smem, st>>>(m_ih, m_il, t);
  } else {
    FR_CUDA_OK(cudaFuncSetAttribute(k_fullsort_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_fullsort_tc<0><<<grid, TC_THREADS, smem, st>>>(m_ih, m_il, t);
  }
  if (prof) prof_end(st);
  count_launch();
  return FR_OK;
}

}  // namespace fr
