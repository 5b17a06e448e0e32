// Whole MLP chains (MLPLayers: [Dropout -> Linear -> BatchNorm1d? -> activation?] x L) as ONE persistent cooperative launch
// for the forward pass and ONE for the backward pass, every GEMM -- forward, data gradient AND weight gradient -- on the
// 5th-generation tensor cores (tcgen05.mma kind::tf32 fed by TMA, fp32 accumulators in TMEM, 3xTF32 for fp32-level accuracy).
//
// Reference being replaced: recbole/model/layers.py:30-85 (MLPLayers and its autograd) as used by the filter /
// discriminator / scorer networks of recbole/model/fair_recommender/pfcn_*.py:105-211 and fairgo_*.py:159-236; up to four
// chains over the same batch rows (the discriminators of one step) share a launch.
//
// Structure.  The launch walks a fixed sequence of PHASES separated by grid barriers (cooperative launch, <= one CTA per SM);
// inside a phase the CTAs pull independent work items (item = blockIdx.x, += gridDim.x):
//   forward :  prep (TF32 hi/lo planes of W; import X with the first dropout)
//              per layer  GEMM items (128 rows x NT columns):  Z = Xin . W^T + b  -> BatchNorm partial sums, or act + emit
//                         [BatchNorm items: finalise batch statistics, normalise + act, emit]
//   backward:  prep (transposed TF32 planes of W; import dY through the last activation)
//              per layer (last to first)  [BatchNorm items: dgamma / dbeta, dz]
//                         GEMM items: weight gradient  dW(chunk) = dZ^T . Xin   (256-row chunks -> ordered partials)
//                                     data gradient    dXin = dZ . W  -> dropout mask, act' of the layer below, column sums
//              reduce (ordered sums of the weight-gradient partials, bias gradients in float64, sum of the chains' dX)
// "emit" writes the next layer's operand ALREADY split into TF32 hi/lo planes, row-major (A operand of the next forward
// GEMM) and transposed (B operand of that layer's weight-gradient GEMM); the backward pass writes dZ the same two ways.
// Every GEMM operand is therefore K-major in global memory (L2-resident at these sizes), TMA streams 128-byte-swizzled
// K-blocks of both planes through a 4-stage ring, and one thread issues hi.hi + hi.lo + lo.hi per k-step: no conversion pass
// between TMA and the tensor core.  Warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue / element-wise items
// (thread == tile row == TMEM lane).  Column sums over the batch (BatchNorm statistics, bias gradients) are accumulated in
// float64 per 128-row tile and added in tile order: no floating-point atomics, bit-reproducible run to run.
#include <stdlib.h>
#include <string.h>

#include "act.cuh"
#include "tc_common.cuh"

namespace fr {

constexpr int CH_MAX_LAYERS = FR_CHAIN_MAX_LAYERS;   // per chain
constexpr int CH_MAX_CHAINS = FR_CHAIN_MAX_CHAINS;
constexpr int CH_MAX_TOTAL = 24;                      // layers of all chains of one launch (kernel-parameter budget)
constexpr int CH_MAXW = 256;                          // widest layer
constexpr int CH_THREADS = 192;
constexpr int CH_STAGES = 4;
constexpr int CH_A_PLANE = TCM * TCKB * 4;            // 16 KB: 128 rows x one 128-byte swizzle row
constexpr int CH_NT = 64;                             // widest B tile (rows of the B operand per item)
constexpr int CH_B_PLANE = CH_NT * TCKB * 4;          // 8 KB
constexpr int CH_STAGE = 2 * CH_A_PLANE + 2 * CH_B_PLANE;
constexpr int CH_TMEM_COLS = 64;
constexpr int CH_WCHUNK = 256;                        // batch rows per weight-gradient partial
constexpr int CH_EW = 1024;                           // elements per flat element-wise item

struct ChLayer {
  int K, N, ldn, NT, KT, act, has_bn, first_of_chain, last_of_chain, chain;
  float drop_p, bn_eps, bn_mom;
  unsigned long long seed;
  const float *W, *b, *gamma, *beta;
  float *rmean, *rvar;
  long long *nbt;
  // forward workspace
  float *Wp;      // [2][N][K]   TF32 planes of W
  float *Xin;     // [2][M][K]   this layer's input after dropout
  float *XinT;    // [2][K][Mpad]
  float *Z;       // [M][ldn]    pre-BatchNorm output (training, BatchNorm layers)
  double *stat;   // [RB][N][2]  per-row-block column sums (sum z, sum z^2)
  float *save_mean, *save_invstd;
  // backward workspace
  float *Wt;      // [2][K][ldn] transposed planes
  float *G1;      // [M][ldn]    gradient at the BatchNorm output
  float *DZ;      // [2][M][ldn] gradient at the Linear output
  float *DZT;     // [2][N][Mpad]
  double *bstat;  // [RB][N][2]  (sum g1, sum g1*xhat) or (sum dz, -)
  float *dWpart;  // [chunks][N][K]
  float *dW, *db, *dgamma, *dbeta;
};

struct ChChain {
  int L, first, K0, Nlast, ldx;
  const float *X;
  float *Y;
  const float *dY;
  float *dX;      // per-chain input gradient [M][K0] (NULL: not wanted)
};

struct ChParams {
  int n_chains, Lmax, M, Mpad, RB, training, need_grad, n_total;
  const unsigned long long *seed_dev;
  unsigned *bar;
  float *dX_sum;  // optional: fixed-order sum of the chains' dX
  ChChain chain[CH_MAX_CHAINS];
  ChLayer layer[CH_MAX_TOTAL];
  CUtensorMap map[CH_MAX_TOTAL][4];   // forward: [0] Xin [1] Wp ; backward: [0] DZ [1] Wt [2] DZT [3] XinT
};

// ---------------------------------------------------------------------------------------------- small device helpers
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void bar_epi() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async;" ::: "memory"); }

// grid barrier: arrival counter bar[0] (reset by the last CTA to leave the kernel, see chain_exit)
__device__ __forceinline__ void chain_barrier(unsigned *bar, unsigned &target) {
  proxy_fence();            // this thread's global writes -> visible to later TMA (async-proxy) reads
  __syncthreads();
  target += gridDim.x;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    while (*(volatile unsigned *)bar < target) {}
    __threadfence();
  }
  __syncthreads();
  proxy_fence();
}
__device__ __forceinline__ void chain_exit(unsigned *bar) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(bar + 1, 1u) == gridDim.x - 1) {   // everyone is past the last barrier
      bar[0] = 0;
      bar[1] = 0;
      __threadfence();
    }
  }
}

__device__ __forceinline__ void split_tf32(float x, float &h, float &l) {
  h = rn_tf32(x);
  l = rn_tf32(x - h);
}

struct GemmIt {
  const CUtensorMap *ma, *mb;
  int ay, by, kb0, nkb, nt;
};

struct Pipe {
  unsigned char *stages;
  uint64_t *full, *empty, *tfull, *tempty;
  uint32_t tmem;
};

__device__ __forceinline__ void gemm_produce(const GemmIt &g, const Pipe &p, uint32_t &it) {
  for (int kb = 0; kb < g.nkb; ++kb, ++it) {
    const int s = it % CH_STAGES;
    mbar_wait(&p.empty[s], ((it / CH_STAGES) & 1) ^ 1);
    unsigned char *st = p.stages + (size_t)s * CH_STAGE;
    mbar_expect_tx(&p.full[s], (uint32_t)(2 * CH_A_PLANE + 2 * g.nt * TCKB * 4));
    const int x = (g.kb0 + kb) * TCKB;
    tma_load_3d(st, g.ma, x, g.ay, 0, &p.full[s]);
    tma_load_3d(st + CH_A_PLANE, g.ma, x, g.ay, 1, &p.full[s]);
    tma_load_3d(st + 2 * CH_A_PLANE, g.mb, x, g.by, 0, &p.full[s]);
    tma_load_3d(st + 2 * CH_A_PLANE + CH_B_PLANE, g.mb, x, g.by, 1, &p.full[s]);
  }
}

__device__ __forceinline__ void gemm_mma(const GemmIt &g, const Pipe &p, uint32_t &it, uint32_t &gi) {
  mbar_wait(p.tempty, (gi & 1) ^ 1);   // the epilogue of this CTA's previous GEMM item has drained the accumulator
  tc_fence_after();
  const uint32_t idesc = umma_idesc_tf32(TCM, g.nt);
  for (int kb = 0; kb < g.nkb; ++kb, ++it) {
    const int s = it % CH_STAGES;
    mbar_wait(&p.full[s], (it / CH_STAGES) & 1);
    tc_fence_after();
    const uint32_t a_hi = smem_u32(p.stages + (size_t)s * CH_STAGE), a_lo = a_hi + CH_A_PLANE;
    const uint32_t b_hi = a_hi + 2 * CH_A_PLANE, b_lo = b_hi + CH_B_PLANE;
#pragma unroll
    for (int k = 0; k < TCKB / 8; ++k) {
      const uint32_t off = k * 32;
      umma_tf32(p.tmem, umma_desc_sw128(a_hi + off), umma_desc_sw128(b_hi + off), idesc, (kb == 0 && k == 0) ? 0u : 1u);
      umma_tf32(p.tmem, umma_desc_sw128(a_hi + off), umma_desc_sw128(b_lo + off), idesc, 1u);
      umma_tf32(p.tmem, umma_desc_sw128(a_lo + off), umma_desc_sw128(b_hi + off), idesc, 1u);
    }
    umma_commit(&p.empty[s]);
  }
  umma_commit(p.tfull);
  ++gi;
}

// epilogue-side shared scratch
struct Epi {
  float (*scr)[33];      // [128][33]
  double (*dscr)[4][32]; // [2][4][32]
  float *colv;           // [5][CH_MAXW]
  int er, et;            // tile row of this thread (TMEM lane), linear epilogue thread id
  unsigned long long seedoff;
};

// column sums over the tile rows of a[.] (and of a[.]^2 when kSq) for 32 columns; results land in threads et < 32
template <bool kSq>
__device__ __forceinline__ void colsum32(const Epi &e, const float (&a)[32], double &s_out, double &ss_out) {
#pragma unroll
  for (int c = 0; c < 32; ++c) e.scr[e.er][c] = a[c];
  bar_epi();
  const int q = e.et >> 5, c = e.et & 31;
  double s = 0.0, ss = 0.0;
#pragma unroll 8
  for (int r = 0; r < 32; ++r) {
    const double v = (double)e.scr[q * 32 + r][c];
    s += v;
    if (kSq) ss += v * v;
  }
  e.dscr[0][q][c] = s;
  if (kSq) e.dscr[1][q][c] = ss;
  bar_epi();
  if (e.et < 32) {
    s_out = ((e.dscr[0][0][c] + e.dscr[0][1][c]) + e.dscr[0][2][c]) + e.dscr[0][3][c];
    if (kSq) ss_out = ((e.dscr[1][0][c] + e.dscr[1][1][c]) + e.dscr[1][2][c]) + e.dscr[1][3][c];
  }
  bar_epi();
}

// y[32] = outputs of layer `li` for row m, columns n0..n0+31 (already activated): hand them to the consumer
__device__ __forceinline__ void emit_fwd(const ChParams &P, const Epi &e, int li, int m, bool mvalid, int n0,
                                         const float (&y)[32]) {
  const ChLayer &L = P.layer[li];
  if (L.last_of_chain) {
    if (!mvalid) return;
    float *dst = P.chain[L.chain].Y + (size_t)m * L.N;
#pragma unroll
    for (int c = 0; c < 32; ++c)
      if (n0 + c < L.N) dst[n0 + c] = y[c];
    return;
  }
  const ChLayer &T = P.layer[li + 1];
  const int K = T.K;
  const float p = P.training ? T.drop_p : 0.f;
  const unsigned long long seed = T.seed + e.seedoff;
  float hi[32], lo[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    const int n = n0 + c;
    float x = (n < K && mvalid) ? y[c] : 0.f;
    if (p > 0.f && n < K) x *= drop_scale(seed, 0, (uint32_t)(m * K + n), p);
    split_tf32(x, hi[c], lo[c]);
  }
  if (mvalid) {
    float *r_hi = T.Xin + (size_t)m * K + n0, *r_lo = r_hi + (size_t)P.M * K;
#pragma unroll
    for (int c = 0; c < 32; c += 4)
      if (n0 + c < K) {
        *(float4 *)(r_hi + c) = make_float4(hi[c], hi[c + 1], hi[c + 2], hi[c + 3]);
        *(float4 *)(r_lo + c) = make_float4(lo[c], lo[c + 1], lo[c + 2], lo[c + 3]);
      }
  }
  if (P.need_grad) {
    float *t_hi = T.XinT + (size_t)n0 * P.Mpad + m, *t_lo = t_hi + (size_t)K * P.Mpad;
#pragma unroll
    for (int c = 0; c < 32; ++c)
      if (n0 + c < K) {
        t_hi[(size_t)c * P.Mpad] = hi[c];
        t_lo[(size_t)c * P.Mpad] = lo[c];
      }
  }
}

// ---------------------------------------------------------------------------------------------- item enumeration
__host__ __device__ inline int ch_tiles(int n, int t) { return (n + t - 1) / t; }

// forward GEMM items of layer position l: per chain RB x ceil(N / NT)
__host__ __device__ inline int fwd_gemm_items(const ChParams &P, int l) {
  int n = 0;
  for (int c = 0; c < P.n_chains; ++c)
    if (l < P.chain[c].L) n += P.RB * ch_tiles(P.layer[P.chain[c].first + l].N, P.layer[P.chain[c].first + l].NT);
  return n;
}
__device__ inline bool fwd_gemm_decode(const ChParams &P, int l, int item, int &li, int &rb, int &nt) {
  for (int c = 0; c < P.n_chains; ++c) {
    if (l >= P.chain[c].L) continue;
    const int i = P.chain[c].first + l;
    const int n = P.RB * ch_tiles(P.layer[i].N, P.layer[i].NT);
    if (item < n) {
      li = i;
      rb = item % P.RB;
      nt = item / P.RB;
      return true;
    }
    item -= n;
  }
  return false;
}
__host__ __device__ inline bool fwd_bn_phase(const ChParams &P, int l) {
  if (!P.training) return false;
  for (int c = 0; c < P.n_chains; ++c)
    if (l < P.chain[c].L && P.layer[P.chain[c].first + l].has_bn) return true;
  return false;
}
// BatchNorm items (forward: position l from the start; backward: step s from the end): one per (chain with BN there, row block)
__device__ inline bool bn_decode(const ChParams &P, int pos, bool from_end, int item, int &li, int &rb) {
  for (int c = 0; c < P.n_chains; ++c) {
    const int l = from_end ? P.chain[c].L - 1 - pos : pos;
    if (l < 0 || l >= P.chain[c].L) continue;
    const int i = P.chain[c].first + l;
    if (!P.layer[i].has_bn) continue;
    if (item < P.RB) {
      li = i;
      rb = item;
      return true;
    }
    item -= P.RB;
  }
  return false;
}
__host__ __device__ inline int bn_items(const ChParams &P, int pos, bool from_end) {
  int n = 0;
  for (int c = 0; c < P.n_chains; ++c) {
    const int l = from_end ? P.chain[c].L - 1 - pos : pos;
    if (l >= 0 && l < P.chain[c].L && P.layer[P.chain[c].first + l].has_bn) n += P.RB;
  }
  return n;
}
// flat element-wise items over the layers: blocks of CH_EW elements of an N*K sized array per layer
__host__ __device__ inline int flat_items(const ChParams &P, bool only_dgrad_layers) {
  int n = 0;
  for (int i = 0; i < P.n_total; ++i) {
    if (only_dgrad_layers && P.layer[i].first_of_chain && !P.chain[P.layer[i].chain].dX) continue;
    n += ch_tiles(P.layer[i].N * P.layer[i].K, CH_EW);
  }
  return n;
}
__device__ inline bool flat_decode(const ChParams &P, bool only_dgrad_layers, int item, int &li, int &blk) {
  for (int i = 0; i < P.n_total; ++i) {
    if (only_dgrad_layers && P.layer[i].first_of_chain && !P.chain[P.layer[i].chain].dX) continue;
    const int n = ch_tiles(P.layer[i].N * P.layer[i].K, CH_EW);
    if (item < n) {
      li = i;
      blk = item;
      return true;
    }
    item -= n;
  }
  return false;
}
__host__ __device__ inline bool needs_dgrad(const ChParams &P, int li) {
  return !P.layer[li].first_of_chain || P.chain[P.layer[li].chain].dX != nullptr;
}
// backward GEMM items at step s (layer L-1-s of each chain): weight-gradient tiles, then data-gradient tiles
__host__ __device__ inline int bwd_gemm_items(const ChParams &P, int s) {
  const int chunks = ch_tiles(P.M, CH_WCHUNK);
  int n = 0;
  for (int c = 0; c < P.n_chains; ++c) {
    const int l = P.chain[c].L - 1 - s;
    if (l < 0) continue;
    const int i = P.chain[c].first + l;
    n += ch_tiles(P.layer[i].N, TCM) * ch_tiles(P.layer[i].K, P.layer[i].KT) * chunks;
    if (needs_dgrad(P, i)) n += P.RB * ch_tiles(P.layer[i].K, P.layer[i].KT);
  }
  return n;
}
struct BwdIt {
  int li, kind;      // kind 0: weight gradient (nt128, kt, chunk) ; 1: data gradient (rb, kt)
  int a, b, c;
};
__device__ inline bool bwd_gemm_decode(const ChParams &P, int s, int item, BwdIt &o) {
  const int chunks = ch_tiles(P.M, CH_WCHUNK);
  for (int c = 0; c < P.n_chains; ++c) {
    const int l = P.chain[c].L - 1 - s;
    if (l < 0) continue;
    const int i = P.chain[c].first + l;
    const int kts = ch_tiles(P.layer[i].K, P.layer[i].KT), nts = ch_tiles(P.layer[i].N, TCM);
    const int nw = nts * kts * chunks;
    if (item < nw) {
      o.li = i;
      o.kind = 0;
      o.c = item % chunks;
      o.b = (item / chunks) % kts;
      o.a = item / (chunks * kts);
      return true;
    }
    item -= nw;
    if (needs_dgrad(P, i)) {
      const int nd = P.RB * kts;
      if (item < nd) {
        o.li = i;
        o.kind = 1;
        o.a = item % P.RB;
        o.b = item / P.RB;
        o.c = 0;
        return true;
      }
      item -= nd;
    }
  }
  return false;
}

// ---------------------------------------------------------------------------------------------- BatchNorm helpers
// batch statistics of layer L from the per-row-block partial sums (tile order, float64) -> colv: [0] mean [1] gamma*invstd
// [2] beta [3] invstd; the rb == 0 item also keeps them for the backward pass and advances the running statistics
__device__ __forceinline__ void bn_fwd_finalize(const ChParams &P, const ChLayer &L, const Epi &e, bool owner) {
  for (int n = e.et; n < L.N; n += 128) {
    double s = 0.0, ss = 0.0;
    for (int rb = 0; rb < P.RB; ++rb) {
      s += __ldcg(L.stat + ((size_t)rb * L.N + n) * 2);
      ss += __ldcg(L.stat + ((size_t)rb * L.N + n) * 2 + 1);
    }
    const double mean = s / (double)P.M;
    double var = ss / (double)P.M - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)L.bn_eps));
    e.colv[n] = (float)mean;
    e.colv[CH_MAXW + n] = __ldg(L.gamma + n) * invstd;
    e.colv[2 * CH_MAXW + n] = __ldg(L.beta + n);
    e.colv[3 * CH_MAXW + n] = invstd;
    if (owner) {
      L.save_mean[n] = (float)mean;
      L.save_invstd[n] = invstd;
      const double unb = P.M > 1 ? var * (double)P.M / (double)(P.M - 1) : var;
      L.rmean[n] = (1.f - L.bn_mom) * L.rmean[n] + L.bn_mom * (float)mean;
      L.rvar[n] = (1.f - L.bn_mom) * L.rvar[n] + L.bn_mom * (float)unb;
    }
  }
  if (owner && e.et == 0 && L.nbt) *L.nbt += 1;
  bar_epi();
}

// ============================================================================================== forward kernel
static __global__ void __launch_bounds__(CH_THREADS, 1) k_mlp_chain_fwd(const __grid_constant__ ChParams P) {
  extern __shared__ __align__(1024) unsigned char ch_smem[];
  unsigned char *base = (unsigned char *)(((uintptr_t)ch_smem + 1023) & ~(uintptr_t)1023);
  Pipe pipe;
  pipe.stages = base;
  Epi e;
  e.scr = (float(*)[33])(base + CH_STAGES * CH_STAGE);
  e.dscr = (double(*)[4][32])((unsigned char *)e.scr + 128 * 33 * 4);
  e.colv = (float *)((unsigned char *)e.dscr + 2 * 4 * 32 * 8);
  uint64_t *bars = (uint64_t *)(e.colv + 5 * CH_MAXW);
  pipe.full = bars;
  pipe.empty = bars + CH_STAGES;
  pipe.tfull = bars + 2 * CH_STAGES;
  pipe.tempty = pipe.tfull + 1;
  uint32_t *tmem_slot = (uint32_t *)(pipe.tempty + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < CH_STAGES; ++s) {
      mbar_init(&pipe.full[s], 1);
      mbar_init(&pipe.empty[s], 1);
    }
    mbar_init(pipe.tfull, 1);
    mbar_init(pipe.tempty, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(CH_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pipe.tmem = *tmem_slot;
  const int quarter = warp & 3;
  e.er = quarter * 32 + lane;
  e.et = (int)threadIdx.x - 64;
  e.seedoff = P.seed_dev ? *P.seed_dev : 0ull;
  const uint32_t taddr = pipe.tmem + ((uint32_t)(quarter * 32) << 16);
  uint32_t it = 0, gi = 0;
  unsigned target = 0;

  // ------------------------------------------------------------------ phase 0: weight planes, import of X
  if (warp >= 2) {
    const int n_flat = flat_items(P, false), n_imp = P.n_chains * P.RB;
    for (int item = blockIdx.x; item < n_flat + n_imp; item += gridDim.x) {
      if (item < n_flat) {
        int li, blk;
        flat_decode(P, false, item, li, blk);
        const ChLayer &L = P.layer[li];
        const int n = L.N * L.K;
        for (int j = 0; j < CH_EW / 128; ++j) {
          const int idx = blk * CH_EW + j * 128 + e.et;
          if (idx < n) {
            float h, l;
            split_tf32(__ldg(L.W + idx), h, l);
            L.Wp[idx] = h;
            L.Wp[n + idx] = l;
          }
        }
      } else {
        const int c = (item - n_flat) / P.RB, rb = (item - n_flat) % P.RB;
        const ChChain &C = P.chain[c];
        const int m = rb * 128 + e.er;
        const bool mvalid = m < P.M;
        for (int n0 = 0; n0 < C.K0; n0 += 32) {
          float y[32];
#pragma unroll
          for (int cc = 0; cc < 32; ++cc) y[cc] = (mvalid && n0 + cc < C.K0) ? __ldg(C.X + (size_t)m * C.ldx + n0 + cc) : 0.f;
          // "layer -1": emit into the chain's first layer
          const ChLayer &T = P.layer[C.first];
          const float p = P.training ? T.drop_p : 0.f;
          const unsigned long long seed = T.seed + e.seedoff;
          float hi[32], lo[32];
#pragma unroll
          for (int cc = 0; cc < 32; ++cc) {
            const int n = n0 + cc;
            float x = y[cc];
            if (p > 0.f && n < T.K && mvalid) x *= drop_scale(seed, 0, (uint32_t)(m * T.K + n), p);
            split_tf32(x, hi[cc], lo[cc]);
          }
          if (mvalid) {
            float *r_hi = T.Xin + (size_t)m * T.K + n0, *r_lo = r_hi + (size_t)P.M * T.K;
#pragma unroll
            for (int cc = 0; cc < 32; cc += 4)
              if (n0 + cc < T.K) {
                *(float4 *)(r_hi + cc) = make_float4(hi[cc], hi[cc + 1], hi[cc + 2], hi[cc + 3]);
                *(float4 *)(r_lo + cc) = make_float4(lo[cc], lo[cc + 1], lo[cc + 2], lo[cc + 3]);
              }
          }
          if (P.need_grad) {
            float *t_hi = T.XinT + (size_t)n0 * P.Mpad + m, *t_lo = t_hi + (size_t)T.K * P.Mpad;
#pragma unroll
            for (int cc = 0; cc < 32; ++cc)
              if (n0 + cc < T.K) {
                t_hi[(size_t)cc * P.Mpad] = hi[cc];
                t_lo[(size_t)cc * P.Mpad] = lo[cc];
              }
          }
        }
      }
    }
  }
  chain_barrier(P.bar, target);

  for (int l = 0; l < P.Lmax; ++l) {
    // ---------------------------------------------------------------- GEMM items of layer position l
    const int n_items = fwd_gemm_items(P, l);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int li, rb, nt;
      fwd_gemm_decode(P, l, item, li, rb, nt);
      const ChLayer &L = P.layer[li];
      if (warp == 0) {
        if (lane == 0) {
          GemmIt g{&P.map[li][0], &P.map[li][1], rb * 128, nt * L.NT, 0, ch_tiles(L.K, TCKB), L.NT};
          gemm_produce(g, pipe, it);
        }
      } else if (warp == 1) {
        if (lane == 0) {
          GemmIt g{nullptr, nullptr, 0, 0, 0, ch_tiles(L.K, TCKB), L.NT};
          gemm_mma(g, pipe, it, gi);
        }
      } else {
        mbar_wait(pipe.tfull, gi & 1);
        tc_fence_after();
        const int m = rb * 128 + e.er;
        const bool mvalid = m < P.M;
        const bool bn_batch = L.has_bn && P.training;
        for (int w = 0; w < L.NT; w += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + w, v);
          const int n0 = nt * L.NT + w;
          float z[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const int n = n0 + c;
            const bool ok = (w + c < L.NT) && n < L.N && mvalid;
            z[c] = ok ? __uint_as_float(v[c]) + (L.b ? __ldg(L.b + n) : 0.f) : 0.f;
          }
          if (bn_batch) {
            if (mvalid) {
              float *dst = L.Z + (size_t)m * L.ldn + n0;
#pragma unroll
              for (int c = 0; c < 32; c += 4)
                if (n0 + c < L.ldn && w + c < L.NT) *(float4 *)(dst + c) = make_float4(z[c], z[c + 1], z[c + 2], z[c + 3]);
            }
            double s = 0.0, ss = 0.0;
            colsum32<true>(e, z, s, ss);
            if (e.et < 32 && w + e.et < L.NT && n0 + e.et < L.N) {
              double *st = L.stat + ((size_t)rb * L.N + n0 + e.et) * 2;
              st[0] = s;
              st[1] = ss;
            }
          } else {
            float y[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const int n = n0 + c;
              float t = z[c];
              if (L.has_bn && n < L.N)
                t = (t - __ldg(L.rmean + n)) * (1.f / sqrtf(__ldg(L.rvar + n) + L.bn_eps)) * __ldg(L.gamma + n) +
                    __ldg(L.beta + n);
              y[c] = (w + c < L.NT) ? act_fwd(t, L.act) : 0.f;
            }
            emit_fwd(P, e, li, m, mvalid, n0, y);   // a 16-wide tile is the only tile of a layer with N <= 16
          }
        }
        tc_fence_before();
        mbar_arrive(pipe.tempty);
        ++gi;
      }
    }
    chain_barrier(P.bar, target);
    // ---------------------------------------------------------------- BatchNorm items of layer position l
    if (fwd_bn_phase(P, l)) {
      if (warp >= 2) {
        const int nb = bn_items(P, l, false);
        for (int item = blockIdx.x; item < nb; item += gridDim.x) {
          int li, rb;
          bn_decode(P, l, false, item, li, rb);
          const ChLayer &L = P.layer[li];
          bn_fwd_finalize(P, L, e, rb == 0);
          const int m = rb * 128 + e.er;
          const bool mvalid = m < P.M;
          for (int n0 = 0; n0 < L.N; n0 += 32) {
            float y[32];
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
              float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (mvalid && n0 + c < L.ldn) z4 = __ldcg((const float4 *)(L.Z + (size_t)m * L.ldn + n0 + c));
              const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int n = n0 + c + j;
                y[c + j] = (n < L.N) ? act_fwd((zz[j] - e.colv[n]) * e.colv[CH_MAXW + n] + e.colv[2 * CH_MAXW + n], L.act) : 0.f;
              }
            }
            emit_fwd(P, e, li, m, mvalid, n0, y);
          }
          bar_epi();   // colv is rewritten by the next item
        }
      }
      chain_barrier(P.bar, target);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(pipe.tmem), "r"(CH_TMEM_COLS));
  }
  chain_exit(P.bar);
}

// ============================================================================================== backward kernel
// g1[32] = gradient at the OUTPUT of layer li's BatchNorm (or Linear when it has none), i.e. already through the
// activation, for row m and columns n0..: store what the layer's own backward GEMMs / BatchNorm item need + column sums
__device__ __forceinline__ void bwd_tail(const ChParams &P, const Epi &e, int li, int rb, int m, bool mvalid, int n0,
                                         float (&g1)[32]) {
  const ChLayer &L = P.layer[li];
#pragma unroll
  for (int c = 0; c < 32; ++c)
    if (!mvalid || n0 + c >= L.N) g1[c] = 0.f;
  double s0 = 0.0, s1 = 0.0, dummy = 0.0;
  if (L.has_bn) {
    float gx[32];
    if (mvalid) {
      float *dst = L.G1 + (size_t)m * L.ldn + n0;
#pragma unroll
      for (int c = 0; c < 32; c += 4)
        if (n0 + c < L.ldn) *(float4 *)(dst + c) = make_float4(g1[c], g1[c + 1], g1[c + 2], g1[c + 3]);
    }
#pragma unroll
    for (int c = 0; c < 32; c += 4) {
      float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (mvalid && n0 + c < L.ldn) z4 = __ldg((const float4 *)(L.Z + (size_t)m * L.ldn + n0 + c));
      const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + c + j;
        gx[c + j] = (n < L.N) ? g1[c + j] * ((zz[j] - __ldg(L.save_mean + n)) * __ldg(L.save_invstd + n)) : 0.f;
      }
    }
    colsum32<false>(e, g1, s0, dummy);
    colsum32<false>(e, gx, s1, dummy);
  } else {
    float hi[32], lo[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) split_tf32(g1[c], hi[c], lo[c]);
    if (mvalid) {
      float *r_hi = L.DZ + (size_t)m * L.ldn + n0, *r_lo = r_hi + (size_t)P.M * L.ldn;
#pragma unroll
      for (int c = 0; c < 32; c += 4)
        if (n0 + c < L.ldn) {
          *(float4 *)(r_hi + c) = make_float4(hi[c], hi[c + 1], hi[c + 2], hi[c + 3]);
          *(float4 *)(r_lo + c) = make_float4(lo[c], lo[c + 1], lo[c + 2], lo[c + 3]);
        }
    }
    float *t_hi = L.DZT + (size_t)n0 * P.Mpad + m, *t_lo = t_hi + (size_t)L.N * P.Mpad;
#pragma unroll
    for (int c = 0; c < 32; ++c)
      if (n0 + c < L.N) {
        t_hi[(size_t)c * P.Mpad] = hi[c];
        t_lo[(size_t)c * P.Mpad] = lo[c];
      }
    colsum32<false>(e, g1, s0, dummy);
  }
  if (e.et < 32 && n0 + e.et < L.N) {
    double *st = L.bstat + ((size_t)rb * L.N + n0 + e.et) * 2;
    st[0] = s0;
    st[1] = s1;
  }
}

static __global__ void __launch_bounds__(CH_THREADS, 1) k_mlp_chain_bwd(const __grid_constant__ ChParams P) {
  extern __shared__ __align__(1024) unsigned char ch_smem[];
  unsigned char *base = (unsigned char *)(((uintptr_t)ch_smem + 1023) & ~(uintptr_t)1023);
  Pipe pipe;
  pipe.stages = base;
  Epi e;
  e.scr = (float(*)[33])(base + CH_STAGES * CH_STAGE);
  e.dscr = (double(*)[4][32])((unsigned char *)e.scr + 128 * 33 * 4);
  e.colv = (float *)((unsigned char *)e.dscr + 2 * 4 * 32 * 8);
  uint64_t *bars = (uint64_t *)(e.colv + 5 * CH_MAXW);
  pipe.full = bars;
  pipe.empty = bars + CH_STAGES;
  pipe.tfull = bars + 2 * CH_STAGES;
  pipe.tempty = pipe.tfull + 1;
  uint32_t *tmem_slot = (uint32_t *)(pipe.tempty + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < CH_STAGES; ++s) {
      mbar_init(&pipe.full[s], 1);
      mbar_init(&pipe.empty[s], 1);
    }
    mbar_init(pipe.tfull, 1);
    mbar_init(pipe.tempty, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(CH_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pipe.tmem = *tmem_slot;
  const int quarter = warp & 3;
  e.er = quarter * 32 + lane;
  e.et = (int)threadIdx.x - 64;
  e.seedoff = P.seed_dev ? *P.seed_dev : 0ull;
  const uint32_t taddr = pipe.tmem + ((uint32_t)(quarter * 32) << 16);
  uint32_t it = 0, gi = 0;
  unsigned target = 0;
  const int chunks = ch_tiles(P.M, CH_WCHUNK);

  // ------------------------------------------------------------------ phase 0: transposed weight planes, import of dY
  if (warp >= 2) {
    const int n_flat = flat_items(P, true), n_imp = P.n_chains * P.RB;
    for (int item = blockIdx.x; item < n_flat + n_imp; item += gridDim.x) {
      if (item < n_flat) {
        int li, blk;
        flat_decode(P, true, item, li, blk);
        const ChLayer &L = P.layer[li];
        const int n = L.N * L.K;
        for (int j = 0; j < CH_EW / 128; ++j) {
          const int idx = blk * CH_EW + j * 128 + e.et;   // idx = k * N + n: coalesced writes
          if (idx < n) {
            const int k = idx / L.N, nn = idx % L.N;
            float h, l;
            split_tf32(__ldg(L.W + (size_t)nn * L.K + k), h, l);
            L.Wt[(size_t)k * L.ldn + nn] = h;
            L.Wt[(size_t)L.K * L.ldn + (size_t)k * L.ldn + nn] = l;
          }
        }
      } else {
        const int c = (item - n_flat) / P.RB, rb = (item - n_flat) % P.RB;
        const ChChain &C = P.chain[c];
        const int li = C.first + C.L - 1;
        const ChLayer &L = P.layer[li];
        const int m = rb * 128 + e.er;
        const bool mvalid = m < P.M;
        for (int n0 = 0; n0 < L.N; n0 += 32) {
          float g1[32];
#pragma unroll
          for (int cc = 0; cc < 32; ++cc) {
            const int n = n0 + cc;
            g1[cc] = (mvalid && n < L.N)
                         ? __ldg(C.dY + (size_t)m * L.N + n) * act_bwd(__ldg(C.Y + (size_t)m * L.N + n), L.act)
                         : 0.f;
          }
          bwd_tail(P, e, li, rb, m, mvalid, n0, g1);
        }
      }
    }
  }
  chain_barrier(P.bar, target);

  for (int s = 0; s < P.Lmax; ++s) {
    // ---------------------------------------------------------------- BatchNorm backward items
    const int nb = bn_items(P, s, true);
    if (nb > 0) {
      if (warp >= 2) {
        for (int item = blockIdx.x; item < nb; item += gridDim.x) {
          int li, rb;
          bn_decode(P, s, true, item, li, rb);
          const ChLayer &L = P.layer[li];
          for (int n = e.et; n < L.N; n += 128) {
            double s0 = 0.0, s1 = 0.0;
            for (int r = 0; r < P.RB; ++r) {
              s0 += __ldcg(L.bstat + ((size_t)r * L.N + n) * 2);
              s1 += __ldcg(L.bstat + ((size_t)r * L.N + n) * 2 + 1);
            }
            const float invstd = __ldg(L.save_invstd + n);
            e.colv[n] = __ldg(L.save_mean + n);
            e.colv[CH_MAXW + n] = __ldg(L.gamma + n) * invstd;
            e.colv[2 * CH_MAXW + n] = (float)(s0 / (double)P.M);
            e.colv[3 * CH_MAXW + n] = invstd;
            e.colv[4 * CH_MAXW + n] = (float)(s1 / (double)P.M);
            if (rb == 0) {
              if (L.dbeta) L.dbeta[n] = (float)s0;
              if (L.dgamma) L.dgamma[n] = (float)s1;
              if (L.db) L.db[n] = 0.f;      // a bias in front of BatchNorm has a mathematically zero gradient
            }
          }
          bar_epi();
          const int m = rb * 128 + e.er;
          const bool mvalid = m < P.M;
          for (int n0 = 0; n0 < L.N; n0 += 32) {
            float hi[32], lo[32];
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
              float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = z4;
              if (mvalid && n0 + c < L.ldn) {
                z4 = __ldg((const float4 *)(L.Z + (size_t)m * L.ldn + n0 + c));
                g4 = __ldcg((const float4 *)(L.G1 + (size_t)m * L.ldn + n0 + c));
              }
              const float zz[4] = {z4.x, z4.y, z4.z, z4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int n = n0 + c + j;
                float dz = 0.f;
                if (n < L.N && mvalid) {
                  const float xh = (zz[j] - e.colv[n]) * e.colv[3 * CH_MAXW + n];
                  dz = e.colv[CH_MAXW + n] * (gg[j] - e.colv[2 * CH_MAXW + n] - xh * e.colv[4 * CH_MAXW + n]);
                }
                split_tf32(dz, hi[c + j], lo[c + j]);
              }
            }
            if (mvalid) {
              float *r_hi = L.DZ + (size_t)m * L.ldn + n0, *r_lo = r_hi + (size_t)P.M * L.ldn;
#pragma unroll
              for (int c = 0; c < 32; c += 4)
                if (n0 + c < L.ldn) {
                  *(float4 *)(r_hi + c) = make_float4(hi[c], hi[c + 1], hi[c + 2], hi[c + 3]);
                  *(float4 *)(r_lo + c) = make_float4(lo[c], lo[c + 1], lo[c + 2], lo[c + 3]);
                }
            }
            float *t_hi = L.DZT + (size_t)n0 * P.Mpad + m, *t_lo = t_hi + (size_t)L.N * P.Mpad;
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (n0 + c < L.N) {
                t_hi[(size_t)c * P.Mpad] = hi[c];
                t_lo[(size_t)c * P.Mpad] = lo[c];
              }
          }
          bar_epi();
        }
      }
      chain_barrier(P.bar, target);
    }
    // ---------------------------------------------------------------- GEMM items: weight gradients + data gradients
    const int n_items = bwd_gemm_items(P, s);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      BwdIt b;
      bwd_gemm_decode(P, s, item, b);
      const ChLayer &L = P.layer[b.li];
      GemmIt g;
      if (b.kind == 0) {
        const int kb0 = b.c * (CH_WCHUNK / TCKB);
        int nkb = ch_tiles(P.M, TCKB) - kb0;
        if (nkb > CH_WCHUNK / TCKB) nkb = CH_WCHUNK / TCKB;
        g = GemmIt{&P.map[b.li][2], &P.map[b.li][3], b.a * TCM, b.b * L.KT, kb0, nkb, L.KT};
      } else {
        g = GemmIt{&P.map[b.li][0], &P.map[b.li][1], b.a * 128, b.b * L.KT, 0, ch_tiles(L.N, TCKB), L.KT};
      }
      if (warp == 0) {
        if (lane == 0) gemm_produce(g, pipe, it);
      } else if (warp == 1) {
        if (lane == 0) gemm_mma(g, pipe, it, gi);
      } else {
        mbar_wait(pipe.tfull, gi & 1);
        tc_fence_after();
        if (b.kind == 0) {
          // D[n (lane), k]: partial weight gradient of this 256-row chunk
          const int n = b.a * TCM + e.er;
          float *dst = L.dWpart + ((size_t)b.c * L.N + n) * L.K;
          for (int w = 0; w < L.KT; w += 32) {
            uint32_t v[32];
            tmem_ld32(taddr + w, v);
            const int k0 = b.b * L.KT + w;
            if (n < L.N) {
#pragma unroll
              for (int c = 0; c < 32; c += 4)
                if (k0 + c < L.K && w + c < L.KT)
                  *(float4 *)(dst + k0 + c) = make_float4(__uint_as_float(v[c]), __uint_as_float(v[c + 1]),
                                                          __uint_as_float(v[c + 2]), __uint_as_float(v[c + 3]));
            }
          }
        } else {
          const int rb = b.a, m = rb * 128 + e.er;
          const bool mvalid = m < P.M;
          const float p = L.drop_p;     // backward runs in training mode only
          const unsigned long long seed = L.seed + e.seedoff;
          for (int w = 0; w < L.KT; w += 32) {
            uint32_t v[32];
            tmem_ld32(taddr + w, v);
            const int k0 = b.b * L.KT + w;
            float g1[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const int k = k0 + c;
              float gq = ((w + c < L.KT) && k < L.K && mvalid) ? __uint_as_float(v[c]) : 0.f;
              if (p > 0.f && k < L.K && mvalid) gq *= drop_scale(seed, 0, (uint32_t)(m * L.K + k), p);
              g1[c] = gq;
            }
            if (L.first_of_chain) {
              if (mvalid) {
                float *dst = P.chain[L.chain].dX + (size_t)m * L.K;
#pragma unroll
                for (int c = 0; c < 32; ++c)
                  if (k0 + c < L.K && w + c < L.KT) dst[k0 + c] = g1[c];
              }
            } else {
              const ChLayer &B = P.layer[b.li - 1];     // the layer below: its output (width L.K == B.N) fed this one
              if (mvalid) {
                const float *x_hi = L.Xin + (size_t)m * L.K + k0, *x_lo = x_hi + (size_t)P.M * L.K;
                const float keep = p > 0.f ? 1.f - p : 1.f;
#pragma unroll
                for (int c = 0; c < 32; c += 4)
                  if (k0 + c < L.K && w + c < L.KT) {
                    const float4 h4 = __ldg((const float4 *)(x_hi + c)), l4 = __ldg((const float4 *)(x_lo + c));
                    g1[c] *= act_bwd((h4.x + l4.x) * keep, B.act);
                    g1[c + 1] *= act_bwd((h4.y + l4.y) * keep, B.act);
                    g1[c + 2] *= act_bwd((h4.z + l4.z) * keep, B.act);
                    g1[c + 3] *= act_bwd((h4.w + l4.w) * keep, B.act);
                  }
              }
              // a 16-wide tile must not touch the columns of its neighbour: bwd_tail masks by B.N only, so tiles narrower
              // than 32 columns are restricted to layers whose width is at most the tile (KT == 16 <=> K <= 16)
              bwd_tail(P, e, b.li - 1, rb, m, mvalid, k0, g1);
            }
          }
        }
        tc_fence_before();
        mbar_arrive(pipe.tempty);
        ++gi;
      }
    }
    chain_barrier(P.bar, target);
  }

  // ------------------------------------------------------------------ reduce: dW, db, dX_sum
  if (warp >= 2) {
    const int n_flat = flat_items(P, false);
    const int n_dx = P.dX_sum ? ch_tiles(P.M * P.chain[0].K0, CH_EW) : 0;
    for (int item = blockIdx.x; item < n_flat + P.n_total + n_dx; item += gridDim.x) {
      if (item < n_flat) {
        int li, blk;
        flat_decode(P, false, item, li, blk);
        const ChLayer &L = P.layer[li];
        const int n = L.N * L.K;
        for (int j = 0; j < CH_EW / 128; ++j) {
          const int idx = blk * CH_EW + j * 128 + e.et;
          if (idx < n) {
            float sw = 0.f;
            for (int c = 0; c < chunks; ++c) sw += __ldcg(L.dWpart + (size_t)c * n + idx);
            L.dW[idx] = sw;
          }
        }
      } else if (item < n_flat + P.n_total) {
        const ChLayer &L = P.layer[item - n_flat];
        if (L.db && !L.has_bn)
          for (int n = e.et; n < L.N; n += 128) {
            double sb = 0.0;
            for (int r = 0; r < P.RB; ++r) sb += __ldcg(L.bstat + ((size_t)r * L.N + n) * 2);
            L.db[n] = (float)sb;
          }
      } else {
        const int blk = item - n_flat - P.n_total, n = P.M * P.chain[0].K0;
        for (int j = 0; j < CH_EW / 128; ++j) {
          const int idx = blk * CH_EW + j * 128 + e.et;
          if (idx < n) {
            float sx = 0.f;
            for (int c = 0; c < P.n_chains; ++c) sx += __ldcg(P.chain[c].dX + idx);
            P.dX_sum[idx] = sx;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(pipe.tmem), "r"(CH_TMEM_COLS));
  }
  chain_exit(P.bar);
}

// ============================================================================================== host side
constexpr size_t CH_SMEM = (size_t)CH_STAGES * CH_STAGE + 128 * 33 * 4 + 2 * 4 * 32 * 8 + 5 * CH_MAXW * 4 + 16 * 8 + 1024;

// planes [2][rows][ld] with `inner` valid columns -> boxes of box_rows x 32 columns of one plane
static int g_map_err = 0;
static bool make_map3(CUtensorMap *m, const float *base, int inner, int rows, int ld, size_t plane_elems, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)rows, 2};
  const cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)plane_elems * 4};
  const cuuint32_t box[3] = {(cuuint32_t)TCKB, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = CUDA_SUCCESS;
  for (int attempt = 0; attempt < 2; ++attempt) {
    r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_ERROR_INVALID_CONTEXT) break;
    // a thread that never touched the runtime (torch's autograd worker on its first backward) has no current context:
    // bind the primary one and retry.  Happens on the eager warm-up pass, never inside a stream capture.
    cudaFree(0);
  }
  if (r != CUDA_SUCCESS) {
    g_map_err = (int)r;
    set_error("cuTensorMapEncodeTiled rc=%d (inner=%d rows=%d ld=%d box_rows=%d)", (int)r, inner, rows, ld, box_rows);
  }
  return r == CUDA_SUCCESS;
}

static inline int tile_for(int n) { return n <= 16 ? 16 : (n <= 32 ? 32 : 64); }
static inline int pad4(int n) { return (n + 3) & ~3; }

struct ChSizes {
  int64_t M, Mpad, RB, chunks;
};
static ChSizes ch_sizes(int64_t M) {
  ChSizes s;
  s.M = M;
  s.RB = (M + 127) / 128;
  s.Mpad = s.RB * 128;
  s.chunks = (M + CH_WCHUNK - 1) / CH_WCHUNK;
  return s;
}

static void carve_fwd(Carver &c, ChLayer &L, const ChSizes &z, bool training, bool need_grad) {
  L.Wp = c.take<float>((size_t)2 * L.N * L.K);
  L.Xin = c.take<float>((size_t)2 * z.M * L.K);
  L.XinT = need_grad ? c.take<float>((size_t)2 * L.K * z.Mpad) : nullptr;
  const bool bn = L.has_bn && training;
  L.Z = bn ? c.take<float>((size_t)z.M * L.ldn) : nullptr;
  L.stat = bn ? c.take<double>((size_t)z.RB * L.N * 2) : nullptr;
  L.save_mean = bn ? c.take<float>(L.N) : nullptr;
  L.save_invstd = bn ? c.take<float>(L.N) : nullptr;
}
static void carve_bwd(Carver &c, ChLayer &L, const ChSizes &z) {
  L.Wt = c.take<float>((size_t)2 * L.K * L.ldn);
  L.G1 = L.has_bn ? c.take<float>((size_t)z.M * L.ldn) : nullptr;
  L.DZ = c.take<float>((size_t)2 * z.M * L.ldn);
  L.DZT = c.take<float>((size_t)2 * L.N * z.Mpad);
  L.bstat = c.take<double>((size_t)z.RB * L.N * 2);
  L.dWpart = c.take<float>((size_t)z.chunks * L.N * L.K);
}

static bool layer_ok(const fr_chain_layer &l) {
  return l.K >= 4 && l.K % 4 == 0 && l.K <= CH_MAXW && l.N >= 1 && l.N <= CH_MAXW && l.act >= 0 && l.act <= 4 && l.W &&
         (!l.has_bn || (l.gamma && l.beta && l.running_mean && l.running_var));
}

static int build_params(const fr_chain *chains, int n_chains, int64_t M, int training, int need_grad, bool backward,
                        const uint64_t *seed_dev, uint32_t *bar, float *dX_sum, ChParams &P, const char *who) {
  FR_REQUIRE(chains && n_chains >= 1 && n_chains <= CH_MAX_CHAINS && M >= 1 && M <= (1 << 22) && bar, "%s: bad argument", who);
  memset(&P, 0, sizeof(P));
  const ChSizes z = ch_sizes(M);
  P.n_chains = n_chains;
  P.M = (int)M;
  P.Mpad = (int)z.Mpad;
  P.RB = (int)z.RB;
  P.training = training ? 1 : 0;
  P.need_grad = (need_grad || backward) ? 1 : 0;
  P.seed_dev = (const unsigned long long *)seed_dev;
  P.bar = bar;
  P.dX_sum = dX_sum;
  int total = 0;
  for (int c = 0; c < n_chains; ++c) {
    const fr_chain &C = chains[c];
    FR_REQUIRE(C.n_layers >= 1 && C.n_layers <= CH_MAX_LAYERS && total + C.n_layers <= CH_MAX_TOTAL, "%s: too many layers", who);
    FR_REQUIRE(C.X && C.Y && C.fwd_ws, "%s: chain %d lacks X / Y / fwd_ws", who, c);
    FR_REQUIRE((int64_t)M * CH_MAXW < ((int64_t)1 << 31), "%s: M too large", who);
    ChChain &D = P.chain[c];
    D.L = C.n_layers;
    D.first = total;
    D.K0 = C.layer[0].K;
    D.Nlast = C.layer[C.n_layers - 1].N;
    D.ldx = C.ldx > 0 ? C.ldx : D.K0;
    D.X = C.X;
    D.Y = C.Y;
    D.dY = C.dY;
    D.dX = C.dX;
    if (C.n_layers > P.Lmax) P.Lmax = C.n_layers;
    Carver cf(C.fwd_ws, C.fwd_ws_bytes), cb(C.bwd_ws, C.bwd_ws_bytes);
    for (int l = 0; l < C.n_layers; ++l) {
      const fr_chain_layer &s = C.layer[l];
      FR_REQUIRE(layer_ok(s), "%s: chain %d layer %d is outside the fused kernel's rules (K %% 4 == 0, widths <= 256)", who, c, l);
      FR_REQUIRE(l == 0 || s.K == C.layer[l - 1].N, "%s: chain %d layer %d: K != previous N", who, c, l);
      ChLayer &L = P.layer[total + l];
      L.K = s.K;
      L.N = s.N;
      L.ldn = pad4(s.N);
      L.NT = tile_for(s.N);
      L.KT = tile_for(s.K);
      L.act = s.act;
      L.has_bn = s.has_bn ? 1 : 0;
      L.first_of_chain = l == 0;
      L.last_of_chain = l == C.n_layers - 1;
      L.chain = c;
      L.drop_p = s.drop_p;
      L.bn_eps = s.bn_eps;
      L.bn_mom = s.bn_momentum;
      L.seed = s.seed;
      L.W = s.W;
      L.b = s.b;
      L.gamma = s.gamma;
      L.beta = s.beta;
      L.rmean = s.running_mean;
      L.rvar = s.running_var;
      L.nbt = (long long *)s.num_batches_tracked;
      L.dW = s.dW;
      L.db = s.db;
      L.dgamma = s.dgamma;
      L.dbeta = s.dbeta;
      carve_fwd(cf, L, z, training != 0, P.need_grad != 0);
      if (backward) carve_bwd(cb, L, z);
    }
    if (!cf.ok() || (backward && (!C.bwd_ws || !cb.ok()))) {
      set_error("%s: chain %d workspace too small", who, c);
      return FR_ERR_WORKSPACE;
    }
    if (backward) FR_REQUIRE(C.dY, "%s: chain %d lacks dY", who, c);
    total += C.n_layers;
  }
  P.n_total = total;
  if (backward && dX_sum) {
    for (int c = 0; c < n_chains; ++c)
      FR_REQUIRE(P.chain[c].dX && P.chain[c].K0 == P.chain[0].K0, "%s: dX_sum needs a dX buffer per chain and equal input widths", who);
  }
  // tensor maps
  for (int i = 0; i < total; ++i) {
    ChLayer &L = P.layer[i];
    bool ok = true;
    if (!backward) {
      ok = ok && make_map3(&P.map[i][0], L.Xin, L.K, P.M, L.K, (size_t)P.M * L.K, TCM);
      ok = ok && make_map3(&P.map[i][1], L.Wp, L.K, L.N, L.K, (size_t)L.N * L.K, L.NT);
    } else {
      ok = ok && make_map3(&P.map[i][0], L.DZ, L.N, P.M, L.ldn, (size_t)P.M * L.ldn, TCM);
      ok = ok && make_map3(&P.map[i][1], L.Wt, L.N, L.K, L.ldn, (size_t)L.K * L.ldn, L.KT);
      ok = ok && make_map3(&P.map[i][2], L.DZT, P.M, L.N, P.Mpad, (size_t)L.N * P.Mpad, TCM);
      ok = ok && make_map3(&P.map[i][3], L.XinT, P.M, L.K, P.Mpad, (size_t)L.K * P.Mpad, L.KT);
    }
    if (!ok) return FR_ERR_CUDA;
  }
  return FR_OK;
}

static int coop_grid(const void *kernel, int want) {
  static int per_sm[2] = {-1, -1}, n_sm = 0;
  const int which = kernel == (const void *)k_mlp_chain_fwd ? 0 : 1;
  if (per_sm[which] < 0) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CH_SMEM);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[which], kernel, CH_THREADS, CH_SMEM);
  }
  int cap = per_sm[which] * n_sm;
  if (cap < 1) return 0;
  static int env_cap = -1;     // FR_CHAIN_MAX_GRID: test hook (forces several items per CTA and phase)
  if (env_cap < 0) {
    const char *e = getenv("FR_CHAIN_MAX_GRID");
    env_cap = e ? atoi(e) : 0;
  }
  if (env_cap > 0 && cap > env_cap) cap = env_cap;
  return want < 1 ? 1 : (want > cap ? cap : want);
}

static int launch_chain(const void *kernel, const char *name, ChParams &P, int want, cudaStream_t st) {
  const int grid = coop_grid(kernel, want);
  if (grid < 1) {
    set_error("%s does not fit an SM", name);
    return FR_ERR_UNSUPPORTED;
  }
  void *args[] = {&P};
  const bool p = prof_on();
  if (p) prof_begin(name, st);
  cudaError_t e = cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(CH_THREADS), args, CH_SMEM, st);
  if (p) prof_end(st);
  count_launch();
  if (e != cudaSuccess) {
    set_error("cudaLaunchCooperativeKernel(%s) failed: %s", name, cudaGetErrorString(e));
    return FR_ERR_CUDA;
  }
  return FR_OK;
}

}  // namespace fr

extern "C" {

int fr_thread_init(void) {
  // binds the device's primary context to the calling thread (a thread that never touched the CUDA runtime has none, and
  // the driver-API tensor-map encoder then fails with CUDA_ERROR_INVALID_CONTEXT).  Not legal inside a stream capture.
  return cudaFree(0) == cudaSuccess ? FR_OK : FR_ERR_CUDA;
}

int fr_mlp_chain_eligible(const fr_chain_layer *layers, int32_t n_layers, int64_t M) {
  if (!layers || n_layers < 1 || n_layers > fr::CH_MAX_LAYERS || M < 1 || M > (1 << 22)) return 0;
  for (int l = 0; l < n_layers; ++l) {
    if (!fr::layer_ok(layers[l])) return 0;
    if (l > 0 && layers[l].K != layers[l - 1].N) return 0;
  }
  return 1;
}

size_t fr_mlp_chain_workspace_bytes(const fr_chain_layer *layers, int32_t n_layers, int64_t M, int32_t training,
                                    int32_t need_grad, int32_t backward) {
  fr::Carver c(nullptr, 0);
  const fr::ChSizes z = fr::ch_sizes(M);
  for (int l = 0; l < n_layers; ++l) {
    fr::ChLayer L;
    memset(&L, 0, sizeof(L));
    L.K = layers[l].K;
    L.N = layers[l].N;
    L.ldn = fr::pad4(L.N);
    L.has_bn = layers[l].has_bn ? 1 : 0;
    if (backward)
      fr::carve_bwd(c, L, z);
    else
      fr::carve_fwd(c, L, z, training != 0, need_grad != 0);
  }
  return c.off + 256;
}

int fr_mlp_chain_forward(const fr_chain *chains, int32_t n_chains, int64_t M, int32_t training, int32_t need_grad,
                         const uint64_t *seed_dev, uint32_t *barrier_words, void *stream) {
  static fr::ChParams P;     // 19 KB: not on the stack of a ctypes caller thread
  int rc = fr::build_params(chains, n_chains, M, training, need_grad, false, seed_dev, barrier_words, nullptr, P,
                            "fr_mlp_chain_forward");
  if (rc) return rc;
  int want = P.n_chains * P.RB + fr::flat_items(P, false);
  for (int l = 0; l < P.Lmax; ++l) {
    const int g = fr::fwd_gemm_items(P, l);
    if (g > want) want = g;
  }
  if (want > P.n_chains * P.RB * 4) want = P.n_chains * P.RB * 4;   // barriers cost more with every extra CTA
  return fr::launch_chain((const void *)fr::k_mlp_chain_fwd, "k_mlp_chain_fwd", P, want, (cudaStream_t)stream);
}

int fr_mlp_chain_backward(const fr_chain *chains, int32_t n_chains, int64_t M, const uint64_t *seed_dev, float *dX_sum,
                          uint32_t *barrier_words, void *stream) {
  static fr::ChParams P;
  int rc = fr::build_params(chains, n_chains, M, 1, 1, true, seed_dev, barrier_words, dX_sum, P, "fr_mlp_chain_backward");
  if (rc) return rc;
  int want = P.n_chains * P.RB;
  for (int s = 0; s < P.Lmax; ++s) {
    const int g = fr::bwd_gemm_items(P, s);
    if (g > want) want = g;
  }
  return fr::launch_chain((const void *)fr::k_mlp_chain_bwd, "k_mlp_chain_bwd", P, want, (cudaStream_t)stream);
}

}  // extern "C"
