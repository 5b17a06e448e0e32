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
constexpr int CH_EPI_WARPS = 16;                      // 4 lane quarters x 4 column groups of 8
constexpr int CH_EPI = CH_EPI_WARPS * 32;
constexpr int CH_THREADS = 64 + CH_EPI;
constexpr int CH_STAGES = 4;
constexpr int CH_A_PLANE = TCM * TCKB * 4;            // 16 KB: 128 rows x one 128-byte swizzle row
constexpr int CH_NT = 32;                             // widest B tile (rows of the B operand per item)
constexpr int CH_B_PLANE = CH_NT * TCKB * 4;          // 4 KB
constexpr int CH_STAGE = 2 * CH_A_PLANE + 2 * CH_B_PLANE;
constexpr int CH_TMEM_COLS = 32;
constexpr int CH_WCHUNK = 256;                        // batch rows per weight-gradient partial
constexpr int CH_EW = 1024;                           // elements per flat element-wise item

struct ChLayer {
  int K, N, ldn, NT, KT, act, has_bn, first_of_chain, last_of_chain, chain;
  float drop_p, bn_eps, bn_mom;
  int bn_repeat;
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
  size_t x_off;   // data-parallel: offset of this layer's [RBg][N][2] column sums inside a parity region
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

// everything but the tensor maps: copied into shared memory at kernel start (indexed reads of a 19 KB kernel-parameter block
// thrash the constant cache: every field access in the row-wise loops became a ~200-cycle miss)
struct ChHead {
  int n_chains, Lmax, M, Mpad, RB, training, need_grad, n_total;
  const unsigned long long *seed_dev;
  unsigned *bar;
  float *dX_sum;  // optional: fixed-order sum of the chains' dX
  unsigned long long *trace;   // optional (FR_CHAIN_TRACE=1): %globaltimer stamps of CTA 0 at the phase boundaries
  // data-parallel BatchNorm (world > 1): this rank holds rows [row_base, row_base + M) of a global batch of Mg rows in
  // RBg row blocks; the per-row-block column sums of the BatchNorm layers live in exchange memory (peer[k], written by
  // every rank into every rank's copy) in two parity regions alternated per exchange
  int rank, world, rb_base, RBg, Mg, row_base, phase_lo, phase_hi, x_parity;
  char *peer[FR_MAX_RANKS];
  size_t x_region, x_stride, x_epoch;     // byte offsets in exchange memory: parity-0 region, parity stride, epoch word
  ChChain chain[CH_MAX_CHAINS];
  ChLayer layer[CH_MAX_TOTAL];
};
struct ChParams {
  ChHead h;
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
__device__ __forceinline__ void bar_epi() { asm volatile("bar.sync 1, %0;" ::"n"(CH_EPI) : "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async;" ::: "memory"); }

// grid barrier: arrival counter bar[0] (reset by the last CTA to leave the kernel, see chain_exit)
__device__ __forceinline__ void chain_barrier(unsigned *bar, unsigned &target) {
  proxy_fence();            // this thread's global writes -> visible to later TMA (async-proxy) reads
  __syncthreads();
  target += gridDim.x;
  if (threadIdx.x == 0) {
    // release / acquire at GPU scope, not __threadfence(): that is fence.sc (MEMBAR.SC.GPU), and every CTA of the grid
    // issuing it in the same microsecond serialises (0.9 .. 6.3 us per fence in the FOCF epoch kernel's phase trace)
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    } while (v < target);
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

__device__ __forceinline__ void trace_stamp(const ChHead &P, int &slot) {
  if (P.trace && blockIdx.x == 0 && threadIdx.x == 64) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    P.trace[slot] = t;
  }
  ++slot;
}

// none / relu / leakyrelu as one branch-free expression (slope 1 / 0 / 0.01); sigmoid and tanh take the generic switch on a
// warp-uniform branch OUTSIDE the per-element loops (a per-element switch on a run-time code cost ~250 ns per call here)
__device__ __forceinline__ float act_slope(int act) { return act == ACT_NONE ? 1.f : (act == ACT_RELU ? 0.f : 0.01f); }
__device__ __forceinline__ void act8_fwd(float (&y)[8], int act) {
  if (act <= ACT_LEAKY) {
    const float sl = act_slope(act);
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = y[j] > 0.f ? y[j] : sl * y[j];
  } else {
#pragma unroll 1
    for (int j = 0; j < 8; ++j) y[j] = act_fwd(y[j], act);
  }
}
// g[j] *= act'(y[j]) with y the activation OUTPUT
__device__ __forceinline__ void act8_bwd(float (&g)[8], const float (&y)[8], int act) {
  if (act <= ACT_LEAKY) {
    const float sl = act_slope(act);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] *= y[j] > 0.f ? 1.f : sl;
  } else {
#pragma unroll 1
    for (int j = 0; j < 8; ++j) g[j] *= act_bwd(y[j], act);
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

// Both issuing loops are run by the WHOLE warp, one elected lane per issue (elect_one(), tc_common.cuh): under
// `if (lane == 0)` the operands of UTMALDG / UTCHMMA live in vector registers and every instruction pays an ELECT /
// R2UR.BROADCAST waterfall (110 of them in each chain kernel, ~100 cycles per 16-64-cycle MMA).
__device__ __forceinline__ void gemm_produce(const GemmIt &g, const Pipe &p, uint32_t &it) {
  for (int kb = 0; kb < g.nkb; ++kb, ++it) {
    const int s = it % CH_STAGES;
    mbar_wait(&p.empty[s], ((it / CH_STAGES) & 1) ^ 1);
    if (elect_one()) {
      unsigned char *st = p.stages + (size_t)s * CH_STAGE;
      mbar_expect_tx(&p.full[s], (uint32_t)(2 * CH_A_PLANE + 2 * g.nt * TCKB * 4));
      const int x = (g.kb0 + kb) * TCKB;
      tma_load_3d(st, g.ma, x, g.ay, 0, &p.full[s]);
      tma_load_3d(st + CH_A_PLANE, g.ma, x, g.ay, 1, &p.full[s]);
      tma_load_3d(st + 2 * CH_A_PLANE, g.mb, x, g.by, 0, &p.full[s]);
      tma_load_3d(st + 2 * CH_A_PLANE + CH_B_PLANE, g.mb, x, g.by, 1, &p.full[s]);
    }
    __syncwarp();
  }
}

__device__ __forceinline__ void gemm_mma(const GemmIt &g, const Pipe &p, uint32_t &it, uint32_t &gi) {
  mbar_wait(p.tempty, (gi & 1) ^ 1);   // the epilogue of this CTA's previous GEMM item has drained the accumulator
  tc_fence_after();
  const uint32_t idesc = umma_idesc_tf32(TCM, g.nt);
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, p.tmem, 0);   // (read from shared memory: make it uniform for ptxas)
  for (int kb = 0; kb < g.nkb; ++kb, ++it) {
    const int s = it % CH_STAGES;
    mbar_wait(&p.full[s], (it / CH_STAGES) & 1);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t a0 = smem_u32(p.stages + (size_t)s * CH_STAGE);
      const uint64_t a_hi = umma_desc_sw128(a0), a_lo = umma_desc_sw128(a0 + CH_A_PLANE);
      const uint64_t b_hi = umma_desc_sw128(a0 + 2 * CH_A_PLANE), b_lo = umma_desc_sw128(a0 + 2 * CH_A_PLANE + CH_B_PLANE);
#pragma unroll
      for (int k = 0; k < TCKB / 8; ++k) {
        const uint64_t off = (uint64_t)(k * 2);   // 32 bytes along the swizzled row, in the descriptor's 16-byte units
        umma_tf32(tmem_u, a_hi + off, b_hi + off, idesc, (kb == 0 && k == 0) ? 0u : 1u);
        umma_tf32(tmem_u, a_hi + off, b_lo + off, idesc, 1u);
        umma_tf32(tmem_u, a_lo + off, b_hi + off, idesc, 1u);
      }
      umma_commit(&p.empty[s]);
      if (kb == g.nkb - 1) umma_commit(p.tfull);
    }
    __syncwarp();
  }
  if (g.nkb <= 0) {   // (never for the shapes the chain rules admit; keeps the accumulator handshake whole)
    if (elect_one()) umma_commit(p.tfull);
    __syncwarp();
  }
  ++gi;
}

// epilogue-side shared scratch
struct Epi {
  float (*scr)[33];      // [128][33]  column-sum staging (first quantity)
  float (*scr2)[33];     // [128][33]  (second quantity)
  double (*dscr)[4][32]; // [2][4][32]
  float *colv;           // [5][32]    per-column coefficients of the current BatchNorm item
  int er, et, c8;        // tile row of this thread (TMEM lane), linear epilogue thread id, first of its 8 columns in a chunk
  size_t xw, xr;         // data-parallel: byte offsets of the parity region written / read by this launch
  unsigned long long seedoff;
};

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// The row-wise bodies below handle EIGHT columns per iteration of a rolled loop (#pragma unroll 1): the epilogue code of
// one item runs once, so its size -- not its instruction count -- is what the instruction cache sees; fully unrolled
// 32-column bodies with inlined activations made these kernels 19k instructions long and instruction-fetch bound.

// column sums over the 128 tile rows of what the row threads staged in scr (and scr2 when kTwo); with kSq the second sum
// is the sum of squares of scr.  Results land in threads et < 32 (column et).
template <bool kTwo, bool kSq>
__device__ __forceinline__ void colsum_reduce(const Epi &e, double &s_out, double &ss_out) {
  bar_epi();
  const int q = (e.et >> 5) & 3, c = e.et & 31;
  if (e.et < 128) {
    double s = 0.0, ss = 0.0;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const double v = (double)e.scr[q * 32 + r][c];
      s += v;
      if (kSq) ss += v * v;
      if (kTwo) ss += (double)e.scr2[q * 32 + r][c];
    }
    e.dscr[0][q][c] = s;
    e.dscr[1][q][c] = ss;
  }
  bar_epi();
  if (e.et < 32) {
    s_out = ((e.dscr[0][0][c] + e.dscr[0][1][c]) + e.dscr[0][2][c]) + e.dscr[0][3][c];
    ss_out = ((e.dscr[1][0][c] + e.dscr[1][1][c]) + e.dscr[1][2][c]) + e.dscr[1][3][c];
  }
  bar_epi();
}

// y[8] = input of layer T (before its dropout) for row m, columns n..n+7: dropout, TF32 hi/lo split, row-major planes
// (A operand of T's forward GEMM) and, when a backward pass follows, transposed planes (B operand of T's weight gradient)
__device__ __forceinline__ void emit8_into(const ChHead &P, const Epi &e, const ChLayer &T, int m, bool mvalid, int n,
                                           const float (&y)[8]) {
  const int K = T.K;
  if (n >= K) return;
  const float p = P.training ? T.drop_p : 0.f;
  const unsigned long long seed = T.seed + e.seedoff;
  float hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float x = (n + j < K && mvalid) ? y[j] : 0.f;
    if (p > 0.f && n + j < K) x *= drop_scale(seed, 0, (uint32_t)((P.row_base + m) * K + n + j), p);
    split_tf32(x, hi[j], lo[j]);
  }
  if (mvalid) {
    float *r_hi = T.Xin + (size_t)m * K + n, *r_lo = r_hi + (size_t)P.M * K;
    *(float4 *)r_hi = make_float4(hi[0], hi[1], hi[2], hi[3]);
    *(float4 *)r_lo = make_float4(lo[0], lo[1], lo[2], lo[3]);
    if (n + 4 < K) {
      *(float4 *)(r_hi + 4) = make_float4(hi[4], hi[5], hi[6], hi[7]);
      *(float4 *)(r_lo + 4) = make_float4(lo[4], lo[5], lo[6], lo[7]);
    }
  }
  if (P.need_grad) {
    float *t_hi = T.XinT + (size_t)n * P.Mpad + m, *t_lo = t_hi + (size_t)K * P.Mpad;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (n + j < K) {
        t_hi[(size_t)j * P.Mpad] = hi[j];
        t_lo[(size_t)j * P.Mpad] = lo[j];
      }
  }
}

// y[8] = outputs of layer `li` for row m, columns n..n+7 (already activated): hand them to the consumer
__device__ __forceinline__ void emit8_fwd(const ChHead &P, const Epi &e, int li, int m, bool mvalid, int n,
                                          const float (&y)[8]) {
  const ChLayer &L = P.layer[li];
  if (L.last_of_chain) {
    if (!mvalid || n >= L.N) return;
    float *dst = P.chain[L.chain].Y + (size_t)m * L.N + n;
    if ((L.N & 3) == 0) {
      *(float4 *)dst = make_float4(y[0], y[1], y[2], y[3]);
      if (n + 4 < L.N) *(float4 *)(dst + 4) = make_float4(y[4], y[5], y[6], y[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (n + j < L.N) dst[j] = y[j];
    }
    return;
  }
  emit8_into(P, e, P.layer[li + 1], m, mvalid, n, y);
}

// ---------------------------------------------------------------------------------------------- item enumeration
__host__ __device__ inline int ch_tiles(int n, int t) { return (n + t - 1) / t; }

// forward GEMM items of layer position l: per chain RB x ceil(N / NT)
__host__ __device__ inline int fwd_gemm_items(const ChHead &P, int l) {
  int n = 0;
  for (int c = 0; c < P.n_chains; ++c)
    if (l < P.chain[c].L) n += P.RB * ch_tiles(P.layer[P.chain[c].first + l].N, P.layer[P.chain[c].first + l].NT);
  return n;
}
__device__ inline bool fwd_gemm_decode(const ChHead &P, int l, int item, int &li, int &rb, int &nt) {
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
__host__ __device__ inline bool fwd_bn_phase(const ChHead &P, int l) {
  if (!P.training) return false;
  for (int c = 0; c < P.n_chains; ++c)
    if (l < P.chain[c].L && P.layer[P.chain[c].first + l].has_bn) return true;
  return false;
}
// BatchNorm items (forward: position l from the start; backward: step s from the end): one per (chain with BN there, row
// block, 32-column chunk)
__device__ inline bool bn_decode(const ChHead &P, int pos, bool from_end, int item, int &li, int &rb, int &cc) {
  for (int c = 0; c < P.n_chains; ++c) {
    const int l = from_end ? P.chain[c].L - 1 - pos : pos;
    if (l < 0 || l >= P.chain[c].L) continue;
    const int i = P.chain[c].first + l;
    if (!P.layer[i].has_bn) continue;
    const int n = P.RB * ch_tiles(P.layer[i].N, 32);
    if (item < n) {
      li = i;
      rb = item % P.RB;
      cc = item / P.RB;
      return true;
    }
    item -= n;
  }
  return false;
}
__host__ __device__ inline int bn_items(const ChHead &P, int pos, bool from_end) {
  int n = 0;
  for (int c = 0; c < P.n_chains; ++c) {
    const int l = from_end ? P.chain[c].L - 1 - pos : pos;
    if (l >= 0 && l < P.chain[c].L && P.layer[P.chain[c].first + l].has_bn)
      n += P.RB * ch_tiles(P.layer[P.chain[c].first + l].N, 32);
  }
  return n;
}
// import items (forward: X into the first layer; backward: dY through the last layer): (chain, row block, 32-column chunk)
__host__ __device__ inline int import_width(const ChHead &P, int c, bool bwd) { return bwd ? P.chain[c].Nlast : P.chain[c].K0; }
__host__ __device__ inline int import_items(const ChHead &P, bool bwd) {
  int n = 0;
  for (int c = 0; c < P.n_chains; ++c) n += P.RB * ch_tiles(import_width(P, c, bwd), 32);
  return n;
}
__device__ inline bool import_decode(const ChHead &P, bool bwd, int item, int &c_out, int &rb, int &cc) {
  for (int c = 0; c < P.n_chains; ++c) {
    const int n = P.RB * ch_tiles(import_width(P, c, bwd), 32);
    if (item < n) {
      c_out = c;
      rb = item % P.RB;
      cc = item / P.RB;
      return true;
    }
    item -= n;
  }
  return false;
}
// flat element-wise items over the layers: blocks of CH_EW elements of an N*K sized array per layer
__host__ __device__ inline int flat_items(const ChHead &P, bool only_dgrad_layers) {
  int n = 0;
  for (int i = 0; i < P.n_total; ++i) {
    if (only_dgrad_layers && P.layer[i].first_of_chain && !P.chain[P.layer[i].chain].dX) continue;
    n += ch_tiles(P.layer[i].N * P.layer[i].K, CH_EW);
  }
  return n;
}
__device__ inline bool flat_decode(const ChHead &P, bool only_dgrad_layers, int item, int &li, int &blk) {
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
__host__ __device__ inline bool needs_dgrad(const ChHead &P, int li) {
  return !P.layer[li].first_of_chain || P.chain[P.layer[li].chain].dX != nullptr;
}
// backward GEMM items at step s (layer L-1-s of each chain): weight-gradient tiles, then data-gradient tiles
__host__ __device__ inline int bwd_gemm_items(const ChHead &P, int s) {
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
__device__ inline bool bwd_gemm_decode(const ChHead &P, int s, int item, BwdIt &o) {
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

// (s, ss) += the `count` double pairs at p, p + stride, ... in index order, eight loads in flight (the adds stay ordered)
__device__ __forceinline__ void ordered_sum2(const double *p, size_t stride, int count, double &s, double &ss) {
  int r = 0;
  for (; r + 16 <= count; r += 16) {
    double2 v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) v[u] = __ldcg((const double2 *)(p + (size_t)(r + u) * stride));
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      s += v[u].x;
      ss += v[u].y;
    }
  }
  if (r < count) {          // tail: up to 15 loads in flight, padded with zeros (adding 0.0 does not change the ordered sum)
    double2 v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u)
      v[u] = (r + u < count) ? __ldcg((const double2 *)(p + (size_t)(r + u) * stride)) : make_double2(0.0, 0.0);
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      s += v[u].x;
      ss += v[u].y;
    }
  }
}

// column-sum partials of a BatchNorm layer: private workspace, or (data-parallel) the exchange memory of rank k
__device__ __forceinline__ double *bn_part_w(const ChHead &P, const Epi &e, const ChLayer &L, double *local, int k) {
  return P.world > 1 ? (double *)(P.peer[k] + e.xw + L.x_off) : local;
}
__device__ __forceinline__ const double *bn_part_r(const ChHead &P, const Epi &e, const ChLayer &L, const double *local) {
  return P.world > 1 ? (const double *)(P.peer[P.rank] + e.xr + L.x_off) : local;
}

// ---------------------------------------------------------------------------------------------- BatchNorm helpers
// batch statistics of columns [n0, n0+32) of layer L from the per-row-block partial sums (tile order, float64) -> colv:
// [0] mean [1] gamma*invstd [2] beta [3] invstd, indexed by column - n0; the rb == 0 item also keeps them for the backward
// pass and advances the running statistics
__device__ __forceinline__ void bn_fwd_finalize(const ChHead &P, const ChLayer &L, const Epi &e, int n0, bool owner) {
  const int n = n0 + e.et;
  if (e.et < 32 && n < L.N) {
    const float gam = __ldg(L.gamma + n), bet = __ldg(L.beta + n);       // in flight under the partial sums
    const float rm0 = owner ? L.rmean[n] : 0.f, rv0 = owner ? L.rvar[n] : 0.f;
    double s = 0.0, ss = 0.0;
    ordered_sum2(bn_part_r(P, e, L, L.stat) + (size_t)n * 2, (size_t)L.N * 2, P.RBg, s, ss);
    const double mean = s / (double)P.Mg;
    double var = ss / (double)P.Mg - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)L.bn_eps));
    e.colv[e.et] = (float)mean;
    e.colv[32 + e.et] = gam * invstd;
    e.colv[64 + e.et] = bet;
    e.colv[96 + e.et] = invstd;
    if (owner) {
      L.save_mean[n] = (float)mean;
      L.save_invstd[n] = invstd;
      const double unb = P.Mg > 1 ? var * (double)P.Mg / (double)(P.Mg - 1) : var;
      float rm = rm0, rv = rv0;
      for (int r = 0; r < L.bn_repeat; ++r) {   // (bn_repeat forward passes over the same batch: fr_chain_layer.bn_repeat)
        rm = (1.f - L.bn_mom) * rm + L.bn_mom * (float)mean;
        rv = (1.f - L.bn_mom) * rv + L.bn_mom * (float)unb;
      }
      L.rmean[n] = rm;
      L.rvar[n] = rv;
      if (n == 0 && L.nbt) *L.nbt += L.bn_repeat;
    }
  }
  bar_epi();
}

struct Ctx {     // per-thread kernel state shared by the two kernels
  Pipe pipe;
  Epi e;
  int warp, lane;
  uint32_t taddr;
  uint32_t *tmem_slot;
  const ChHead *head;    // shared-memory copy
};

__device__ __forceinline__ void chain_setup(Ctx &c, unsigned char *smem, const ChHead &P) {
  unsigned char *base = (unsigned char *)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  c.pipe.stages = base;
  c.e.scr = (float(*)[33])(base + CH_STAGES * CH_STAGE);
  c.e.scr2 = (float(*)[33])((unsigned char *)c.e.scr + 128 * 33 * 4);
  c.e.dscr = (double(*)[4][32])((unsigned char *)c.e.scr2 + 128 * 33 * 4);
  c.e.colv = (float *)((unsigned char *)c.e.dscr + 2 * 4 * 32 * 8);
  uint64_t *bars = (uint64_t *)(c.e.colv + 5 * 32);
  c.pipe.full = bars;
  c.pipe.empty = bars + CH_STAGES;
  c.pipe.tfull = bars + 2 * CH_STAGES;
  c.pipe.tempty = c.pipe.tfull + 1;
  c.tmem_slot = (uint32_t *)(c.pipe.tempty + 1);
  {
    uint32_t *dst = (uint32_t *)(((uintptr_t)(c.tmem_slot + 4) + 15) & ~(uintptr_t)15);
    const uint32_t *src = (const uint32_t *)&P;
    for (int i = threadIdx.x; i < (int)(sizeof(ChHead) / 4); i += CH_THREADS) dst[i] = src[i];
    c.head = (const ChHead *)dst;
  }
  c.warp = threadIdx.x >> 5;
  c.lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < CH_STAGES; ++s) {
      mbar_init(&c.pipe.full[s], 1);
      mbar_init(&c.pipe.empty[s], 1);
    }
    mbar_init(c.pipe.tfull, 1);
    mbar_init(c.pipe.tempty, CH_EPI);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (c.warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(c.tmem_slot)),
                 "r"(CH_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.pipe.tmem = *c.tmem_slot;
  const int quarter = c.warp & 3;
  c.e.er = quarter * 32 + c.lane;
  c.e.et = (int)threadIdx.x - 64;
  c.e.c8 = ((c.warp - 2) >> 2) * 8;
  c.e.seedoff = P.seed_dev ? *P.seed_dev : 0ull;
  c.e.xw = c.e.xr = 0;
  if (P.world > 1) {
    const unsigned par = P.x_parity >= 0 ? (unsigned)P.x_parity
                                         : (unsigned)(*(const unsigned long long *)(P.peer[P.rank] + P.x_epoch) & 1ull);
    c.e.xw = P.x_region + (size_t)par * P.x_stride;
    c.e.xr = P.x_region + (size_t)(par ^ 1u) * P.x_stride;
  }
  c.taddr = c.pipe.tmem + ((uint32_t)(quarter * 32) << 16);
}

__device__ __forceinline__ void chain_teardown(Ctx &c, const ChHead &P) {
  tc_fence_before();
  __syncthreads();
  if (c.warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(c.pipe.tmem), "r"(CH_TMEM_COLS));
  }
  chain_exit(P.bar);
}

// TF32 planes of the weights: forward layout [2][N][K] (kT == false) or transposed [2][K][ldn] (kT == true)
template <bool kT>
__device__ __forceinline__ void weight_planes_item(const ChLayer &L, int blk, int et) {
  const int n = L.N * L.K;
#pragma unroll 1
  for (int j = 0; j < CH_EW / CH_EPI; ++j) {
    const int idx = blk * CH_EW + j * CH_EPI + et;
    if (idx >= n) break;
    float h, l;
    if (!kT) {
      split_tf32(__ldg(L.W + idx), h, l);
      L.Wp[idx] = h;
      L.Wp[n + idx] = l;
    } else {
      const int k = idx / L.N, nn = idx % L.N;      // idx = k * N + n: coalesced writes
      split_tf32(__ldg(L.W + (size_t)nn * L.K + k), h, l);
      L.Wt[(size_t)k * L.ldn + nn] = h;
      L.Wt[(size_t)L.K * L.ldn + (size_t)k * L.ldn + nn] = l;
    }
  }
}

// ============================================================================================== forward kernel
static __global__ void __launch_bounds__(CH_THREADS, 1) k_mlp_chain_fwd(const __grid_constant__ ChParams PP) {
  extern __shared__ __align__(1024) unsigned char ch_smem[];
  Ctx cx;
  chain_setup(cx, ch_smem, PP.h);
  const ChHead &P = *cx.head;
  const Pipe &pipe = cx.pipe;
  const Epi &e = cx.e;
  const int warp = cx.warp;
  uint32_t it = 0, gi = 0;
  unsigned target = 0;
  int tslot = 0;
  trace_stamp(P, tslot);

  // phases: 0 = weight planes + import of X ; 1 + 2l = GEMM items of layer position l ; 2 + 2l = its BatchNorm items.
  // A launch runs the window [phase_lo, phase_hi) (the whole chain unless the data-parallel host cuts it where BatchNorm
  // partials cross the ranks); grid barriers separate the phases inside the window.
  const int plo = P.phase_lo, phi = P.phase_hi;
  // ------------------------------------------------------------------ phase 0: weight planes, import of X
  if (warp >= 2 && plo <= 0 && 0 < phi) {
    const int n_flat = flat_items(P, false), n_imp = import_items(P, false);
    for (int item = blockIdx.x; item < n_flat + n_imp; item += gridDim.x) {
      if (item < n_flat) {
        int li, blk;
        flat_decode(P, false, item, li, blk);
        weight_planes_item<false>(P.layer[li], blk, e.et);
      } else {
        int c, rb, cc;
        import_decode(P, false, item - n_flat, c, rb, cc);
        const ChChain &C = P.chain[c];
        const int m = rb * 128 + e.er;
        const bool mvalid = m < P.M;
        const bool vec = (C.ldx & 3) == 0 && (((uintptr_t)C.X) & 15) == 0;
        const int n = cc * 32 + e.c8;
        if (n < C.K0) {
          const float *src = C.X + (size_t)m * C.ldx + n;
          float y[8];
          if (mvalid && vec && n + 8 <= C.K0) {
            const float4 a = __ldg((const float4 *)src), b = __ldg((const float4 *)(src + 4));
            y[0] = a.x; y[1] = a.y; y[2] = a.z; y[3] = a.w; y[4] = b.x; y[5] = b.y; y[6] = b.z; y[7] = b.w;
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = (mvalid && n + j < C.K0) ? __ldg(src + j) : 0.f;
          }
          emit8_into(P, e, P.layer[C.first], m, mvalid, n, y);
        }
      }
    }
  }
  trace_stamp(P, tslot);
  if (plo <= 0 && 1 < phi) chain_barrier(P.bar, target);
  trace_stamp(P, tslot);

  for (int l = 0; l < P.Lmax; ++l) {
    // ---------------------------------------------------------------- GEMM items of layer position l
    const int pg = 1 + 2 * l;
    const int n_items = (plo <= pg && pg < phi) ? fwd_gemm_items(P, l) : 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int li, rb, nt;
      fwd_gemm_decode(P, l, item, li, rb, nt);
      const ChLayer &L = P.layer[li];
      if (warp == 0) {
        GemmIt g{&PP.map[li][0], &PP.map[li][1], rb * 128, nt * L.NT, 0, ch_tiles(L.K, TCKB), L.NT};
        gemm_produce(g, pipe, it);
      } else if (warp == 1) {
        GemmIt g{nullptr, nullptr, 0, 0, 0, ch_tiles(L.K, TCKB), L.NT};
        gemm_mma(g, pipe, it, gi);
      } else {
        mbar_wait(pipe.tfull, gi & 1);
        tc_fence_after();
        trace_stamp(P, tslot);
        const int m = rb * 128 + e.er;
        const bool mvalid = m < P.M;
        const bool bn_batch = L.has_bn && P.training;
        const int n0 = nt * L.NT;
        const int c8 = e.c8;
        if (c8 < L.NT) {
          uint32_t v[8];
          tmem_ld8(cx.taddr + c8, v);
          const int n = n0 + c8;
          float z[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            z[j] = (n + j < L.N && mvalid) ? __uint_as_float(v[j]) + (L.b ? __ldg(L.b + n + j) : 0.f) : 0.f;
          if (bn_batch) {
            if (mvalid) {
              float *dst = L.Z + (size_t)m * L.ldn + n;
              if (n < L.ldn) *(float4 *)dst = make_float4(z[0], z[1], z[2], z[3]);
              if (n + 4 < L.ldn) *(float4 *)(dst + 4) = make_float4(z[4], z[5], z[6], z[7]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) e.scr[e.er][c8 + j] = z[j];
          } else {
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = z[j];
            if (L.has_bn) {      // eval mode: running statistics
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (n + j < L.N)
                  y[j] = (y[j] - __ldg(L.rmean + n + j)) * (1.f / sqrtf(__ldg(L.rvar + n + j) + L.bn_eps)) *
                             __ldg(L.gamma + n + j) + __ldg(L.beta + n + j);
            }
            act8_fwd(y, L.act);
            emit8_fwd(P, e, li, m, mvalid, n, y);
          }
        }
        if (bn_batch) {
          double s = 0.0, ss = 0.0;
          colsum_reduce<false, true>(e, s, ss);
          if (e.et < L.NT && n0 + e.et < L.N) {
            for (int k = 0; k < P.world; ++k) {     // data-parallel: into every rank's copy (NVLink peer stores)
              double *st = bn_part_w(P, e, L, L.stat, k) + ((size_t)(P.rb_base + rb) * L.N + n0 + e.et) * 2;
              st[0] = s;
              st[1] = ss;
            }
          }
        }
        trace_stamp(P, tslot);
        tc_fence_before();
        mbar_arrive(pipe.tempty);
        ++gi;
      }
    }
    trace_stamp(P, tslot);
    if (plo <= pg && pg + 1 < phi) chain_barrier(P.bar, target);
    trace_stamp(P, tslot);
    // ---------------------------------------------------------------- BatchNorm items of layer position l
    if (fwd_bn_phase(P, l) && plo <= pg + 1 && pg + 1 < phi) {
      if (warp >= 2) {
        const int nb = bn_items(P, l, false);
        for (int item = blockIdx.x; item < nb; item += gridDim.x) {
          int li, rb, cc;
          bn_decode(P, l, false, item, li, rb, cc);
          const ChLayer &L = P.layer[li];
          const int n0 = cc * 32;
          const int m = rb * 128 + e.er;
          const bool mvalid = m < P.M;
          const int c8 = e.c8, n = n0 + c8;
          const float *zrow = L.Z + (size_t)m * L.ldn + n;
          float4 za = make_float4(0.f, 0.f, 0.f, 0.f), zb = za;      // the loads run under the statistics' finalisation
          if (mvalid && n < L.ldn) za = __ldcg((const float4 *)zrow);
          if (mvalid && n + 4 < L.ldn) zb = __ldcg((const float4 *)(zrow + 4));
          bn_fwd_finalize(P, L, e, n0, rb == 0);
          if (n < L.N) {
            const float zz[8] = {za.x, za.y, za.z, za.w, zb.x, zb.y, zb.z, zb.w};
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              y[j] = (n + j < L.N) ? (zz[j] - e.colv[c8 + j]) * e.colv[32 + c8 + j] + e.colv[64 + c8 + j] : 0.f;
            act8_fwd(y, L.act);
            emit8_fwd(P, e, li, m, mvalid, n, y);
          }
          bar_epi();   // colv is rewritten by the next item
        }
      }
      trace_stamp(P, tslot);
      if (pg + 2 < phi) chain_barrier(P.bar, target);
      trace_stamp(P, tslot);
    }
  }
  trace_stamp(P, tslot);
  chain_teardown(cx, P);
}

// ============================================================================================== backward kernel
// g1[8] = gradient at the OUTPUT of layer li's BatchNorm (or Linear when it has none), i.e. already through the activation,
// for row m and columns n..n+7 of the 32-column chunk starting at n0: store what the layer's own backward GEMMs /
// BatchNorm item need and stage the column sums (bwd_tail_finish reduces them once the chunk is complete)
__device__ __forceinline__ void bwd_tail8(const ChHead &P, const Epi &e, const ChLayer &L, int m, bool mvalid, int n0, int n,
                                          float (&g1)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (!mvalid || n + j >= L.N) g1[j] = 0.f;
  if (L.has_bn) {
    float4 za = make_float4(0.f, 0.f, 0.f, 0.f), zb = za;
    if (mvalid) {
      float *dst = L.G1 + (size_t)m * L.ldn + n;
      const float *zrow = L.Z + (size_t)m * L.ldn + n;
      if (n < L.ldn) {
        *(float4 *)dst = make_float4(g1[0], g1[1], g1[2], g1[3]);
        za = __ldg((const float4 *)zrow);
      }
      if (n + 4 < L.ldn) {
        *(float4 *)(dst + 4) = make_float4(g1[4], g1[5], g1[6], g1[7]);
        zb = __ldg((const float4 *)(zrow + 4));
      }
    }
    const float zz[8] = {za.x, za.y, za.z, za.w, zb.x, zb.y, zb.z, zb.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      e.scr[e.er][n - n0 + j] = g1[j];
      e.scr2[e.er][n - n0 + j] =
          (n + j < L.N) ? g1[j] * ((zz[j] - __ldg(L.save_mean + n + j)) * __ldg(L.save_invstd + n + j)) : 0.f;
    }
  } else {
    float hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      split_tf32(g1[j], hi[j], lo[j]);
      e.scr[e.er][n - n0 + j] = g1[j];
    }
    if (mvalid) {
      float *r_hi = L.DZ + (size_t)m * L.ldn + n, *r_lo = r_hi + (size_t)P.M * L.ldn;
      if (n < L.ldn) {
        *(float4 *)r_hi = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *(float4 *)r_lo = make_float4(lo[0], lo[1], lo[2], lo[3]);
      }
      if (n + 4 < L.ldn) {
        *(float4 *)(r_hi + 4) = make_float4(hi[4], hi[5], hi[6], hi[7]);
        *(float4 *)(r_lo + 4) = make_float4(lo[4], lo[5], lo[6], lo[7]);
      }
    }
    float *t_hi = L.DZT + (size_t)n * P.Mpad + m, *t_lo = t_hi + (size_t)L.N * P.Mpad;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (n + j < L.N) {
        t_hi[(size_t)j * P.Mpad] = hi[j];
        t_lo[(size_t)j * P.Mpad] = lo[j];
      }
  }
}
// `cols` = how many columns of the chunk the caller staged (the rest of scr is stale and ignored)
__device__ __forceinline__ void bwd_tail_finish(const ChHead &P, const Epi &e, const ChLayer &L, int rb, int n0, int cols) {
  double s0 = 0.0, s1 = 0.0;
  if (L.has_bn)
    colsum_reduce<true, false>(e, s0, s1);
  else
    colsum_reduce<false, false>(e, s0, s1);
  if (e.et < cols && n0 + e.et < L.N) {
    if (L.has_bn) {
      for (int k = 0; k < P.world; ++k) {
        double *st = bn_part_w(P, e, L, L.bstat, k) + ((size_t)(P.rb_base + rb) * L.N + n0 + e.et) * 2;
        st[0] = s0;
        st[1] = s1;
      }
    } else {       // bias gradient partial of a layer without BatchNorm: stays local (the host all-reduces the gradients)
      double *st = L.bstat + ((size_t)rb * L.N + n0 + e.et) * 2;
      st[0] = s0;
      st[1] = 0.0;
    }
  }
}

static __global__ void __launch_bounds__(CH_THREADS, 1) k_mlp_chain_bwd(const __grid_constant__ ChParams PP) {
  extern __shared__ __align__(1024) unsigned char ch_smem[];
  Ctx cx;
  chain_setup(cx, ch_smem, PP.h);
  const ChHead &P = *cx.head;
  const Pipe &pipe = cx.pipe;
  const Epi &e = cx.e;
  const int warp = cx.warp;
  uint32_t it = 0, gi = 0;
  unsigned target = 0;
  const int chunks = ch_tiles(P.M, CH_WCHUNK);
  int tslot = 0;
  trace_stamp(P, tslot);

  // phases: 0 = transposed weight planes + import of dY ; 1 + 2s = BatchNorm backward items of step s (layer L-1-s) ;
  // 2 + 2s = its GEMM items ; 1 + 2 Lmax = reduce.  Window [phase_lo, phase_hi) as in the forward kernel.
  const int plo = P.phase_lo, phi = P.phase_hi;
  // ------------------------------------------------------------------ phase 0: transposed weight planes, import of dY
  if (warp >= 2 && plo <= 0 && 0 < phi) {
    const int n_flat = flat_items(P, true), n_imp = import_items(P, true);
    for (int item = blockIdx.x; item < n_flat + n_imp; item += gridDim.x) {
      if (item < n_flat) {
        int li, blk;
        flat_decode(P, true, item, li, blk);
        weight_planes_item<true>(P.layer[li], blk, e.et);
      } else {
        int c, rb, cc;
        import_decode(P, true, item - n_flat, c, rb, cc);
        const ChChain &C = P.chain[c];
        const ChLayer &L = P.layer[C.first + C.L - 1];
        const int m = rb * 128 + e.er, n0 = cc * 32;
        const bool mvalid = m < P.M;
        const int cols = L.N - n0 < 32 ? L.N - n0 : 32;
        const int n = n0 + e.c8;
        if (n < L.N) {
          float g1[8], yo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const bool ok = mvalid && n + j < L.N;
            g1[j] = ok ? __ldg(C.dY + (size_t)m * L.N + n + j) : 0.f;
            yo[j] = ok ? __ldg(C.Y + (size_t)m * L.N + n + j) : 0.f;
          }
          act8_bwd(g1, yo, L.act);
          bwd_tail8(P, e, L, m, mvalid, n0, n, g1);
        }
        bwd_tail_finish(P, e, L, rb, n0, cols);
      }
    }
  }
  trace_stamp(P, tslot);
  if (plo <= 0 && 1 < phi) chain_barrier(P.bar, target);
  trace_stamp(P, tslot);

  for (int s = 0; s < P.Lmax; ++s) {
    // ---------------------------------------------------------------- BatchNorm backward items
    const int pb = 1 + 2 * s;
    const int nb = (plo <= pb && pb < phi) ? bn_items(P, s, true) : 0;
    if (nb > 0) {
      if (warp >= 2) {
        for (int item = blockIdx.x; item < nb; item += gridDim.x) {
          int li, rb, cc;
          bn_decode(P, s, true, item, li, rb, cc);
          const ChLayer &L = P.layer[li];
          const int n0 = cc * 32;
          const int m = rb * 128 + e.er;
          const bool mvalid = m < P.M;
          const int c8 = e.c8, n = n0 + c8;
          // this thread's row loads run under the finalisation of the column sums
          float4 za = make_float4(0.f, 0.f, 0.f, 0.f), zb = za, ga = za, gb = za;
          if (mvalid && n < L.N && n < L.ldn) {
            za = __ldg((const float4 *)(L.Z + (size_t)m * L.ldn + n));
            ga = __ldcg((const float4 *)(L.G1 + (size_t)m * L.ldn + n));
          }
          if (mvalid && n < L.N && n + 4 < L.ldn) {
            zb = __ldg((const float4 *)(L.Z + (size_t)m * L.ldn + n + 4));
            gb = __ldcg((const float4 *)(L.G1 + (size_t)m * L.ldn + n + 4));
          }
          if (e.et < 32 && n0 + e.et < L.N) {
            const int nc = n0 + e.et;
            const float invstd = __ldg(L.save_invstd + nc), mean = __ldg(L.save_mean + nc), gam = __ldg(L.gamma + nc);
            double s0 = 0.0, s1 = 0.0;
            const double *bs = bn_part_r(P, e, L, L.bstat) + (size_t)nc * 2;
            ordered_sum2(bs, (size_t)L.N * 2, P.RBg, s0, s1);
            e.colv[e.et] = mean;
            e.colv[32 + e.et] = gam * invstd;
            e.colv[64 + e.et] = (float)(s0 / (double)P.Mg);
            e.colv[96 + e.et] = invstd;
            e.colv[128 + e.et] = (float)(s1 / (double)P.Mg);
            if (rb == 0) {
              if (P.world > 1) {      // gradient outputs are this rank's share (the host sums the ranks' gradients)
                s0 = s1 = 0.0;
                ordered_sum2(bs + (size_t)P.rb_base * L.N * 2, (size_t)L.N * 2, P.RB, s0, s1);
              }
              if (L.dbeta) L.dbeta[nc] = (float)s0;
              if (L.dgamma) L.dgamma[nc] = (float)s1;
              if (L.db) L.db[nc] = 0.f;      // a bias in front of BatchNorm has a mathematically zero gradient
            }
          }
          bar_epi();
          if (n < L.N) {
            const float zz[8] = {za.x, za.y, za.z, za.w, zb.x, zb.y, zb.z, zb.w};
            const float gg[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
            float hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float dz = 0.f;
              if (n + j < L.N && mvalid) {
                const float xh = (zz[j] - e.colv[c8 + j]) * e.colv[96 + c8 + j];
                dz = e.colv[32 + c8 + j] * (gg[j] - e.colv[64 + c8 + j] - xh * e.colv[128 + c8 + j]);
              }
              split_tf32(dz, hi[j], lo[j]);
            }
            if (mvalid) {
              float *r_hi = L.DZ + (size_t)m * L.ldn + n, *r_lo = r_hi + (size_t)P.M * L.ldn;
              if (n < L.ldn) {
                *(float4 *)r_hi = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *(float4 *)r_lo = make_float4(lo[0], lo[1], lo[2], lo[3]);
              }
              if (n + 4 < L.ldn) {
                *(float4 *)(r_hi + 4) = make_float4(hi[4], hi[5], hi[6], hi[7]);
                *(float4 *)(r_lo + 4) = make_float4(lo[4], lo[5], lo[6], lo[7]);
              }
            }
            float *t_hi = L.DZT + (size_t)n * P.Mpad + m, *t_lo = t_hi + (size_t)L.N * P.Mpad;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (n + j < L.N) {
                t_hi[(size_t)j * P.Mpad] = hi[j];
                t_lo[(size_t)j * P.Mpad] = lo[j];
              }
          }
          bar_epi();
        }
      }
      trace_stamp(P, tslot);
      if (pb + 1 < phi) chain_barrier(P.bar, target);
      trace_stamp(P, tslot);
    }
    // ---------------------------------------------------------------- GEMM items: weight gradients + data gradients
    const int n_items = (plo <= pb + 1 && pb + 1 < phi) ? bwd_gemm_items(P, s) : 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      BwdIt b;
      bwd_gemm_decode(P, s, item, b);
      const ChLayer &L = P.layer[b.li];
      GemmIt g;
      if (b.kind == 0) {
        const int kb0 = b.c * (CH_WCHUNK / TCKB);
        int nkb = ch_tiles(P.M, TCKB) - kb0;
        if (nkb > CH_WCHUNK / TCKB) nkb = CH_WCHUNK / TCKB;
        g = GemmIt{&PP.map[b.li][2], &PP.map[b.li][3], b.a * TCM, b.b * L.KT, kb0, nkb, L.KT};
      } else {
        g = GemmIt{&PP.map[b.li][0], &PP.map[b.li][1], b.a * 128, b.b * L.KT, 0, ch_tiles(L.N, TCKB), L.KT};
      }
      if (warp == 0) {
        gemm_produce(g, pipe, it);
      } else if (warp == 1) {
        gemm_mma(g, pipe, it, gi);
      } else {
        mbar_wait(pipe.tfull, gi & 1);
        tc_fence_after();
        const int k0 = b.b * L.KT;
        if (b.kind == 0) {
          // D[n (lane), k]: partial weight gradient of this 256-row chunk
          const int n = b.a * TCM + e.er;
          float *dst = L.dWpart + ((size_t)b.c * L.N + n) * L.K + k0;
          const int c8 = e.c8;
          if (c8 < L.KT) {
            uint32_t v[8];
            tmem_ld8(cx.taddr + c8, v);
            if (n < L.N) {
              if (k0 + c8 < L.K)
                *(float4 *)(dst + c8) = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]),
                                                    __uint_as_float(v[3]));
              if (k0 + c8 + 4 < L.K)
                *(float4 *)(dst + c8 + 4) = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]),
                                                        __uint_as_float(v[7]));
            }
          }
        } else {
          const int rb = b.a, m = rb * 128 + e.er;
          const bool mvalid = m < P.M;
          const float p = L.drop_p;     // backward runs in training mode only
          const float keep = p > 0.f ? 1.f - p : 1.f;
          const unsigned long long seed = L.seed + e.seedoff;
          const int cols = L.K - k0 < L.KT ? L.K - k0 : L.KT;
          const int c8 = e.c8, k = k0 + c8;
          if (c8 < L.KT && k < L.K) {
            uint32_t v[8];
            tmem_ld8(cx.taddr + c8, v);
            float g1[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float gq = (k + j < L.K && mvalid) ? __uint_as_float(v[j]) : 0.f;
              if (p > 0.f && k + j < L.K && mvalid) gq *= drop_scale(seed, 0, (uint32_t)((P.row_base + m) * L.K + k + j), p);
              g1[j] = gq;
            }
            if (L.first_of_chain) {
              if (mvalid) {
                float *dst = P.chain[L.chain].dX + (size_t)m * L.K + k;
                *(float4 *)dst = make_float4(g1[0], g1[1], g1[2], g1[3]);
                if (k + 4 < L.K) *(float4 *)(dst + 4) = make_float4(g1[4], g1[5], g1[6], g1[7]);
              }
            } else {
              const ChLayer &B = P.layer[b.li - 1];     // the layer below: its output (width L.K == B.N) fed this one
              if (mvalid) {
                const float *x_hi = L.Xin + (size_t)m * L.K + k, *x_lo = x_hi + (size_t)P.M * L.K;
                const float4 h4 = __ldg((const float4 *)x_hi), l4 = __ldg((const float4 *)x_lo);
                float4 h5 = make_float4(0.f, 0.f, 0.f, 0.f), l5 = h5;
                if (k + 4 < L.K) {
                  h5 = __ldg((const float4 *)(x_hi + 4));
                  l5 = __ldg((const float4 *)(x_lo + 4));
                }
                const float yo[8] = {(h4.x + l4.x) * keep, (h4.y + l4.y) * keep, (h4.z + l4.z) * keep, (h4.w + l4.w) * keep,
                                     (h5.x + l5.x) * keep, (h5.y + l5.y) * keep, (h5.z + l5.z) * keep, (h5.w + l5.w) * keep};
                act8_bwd(g1, yo, B.act);
              }
              bwd_tail8(P, e, B, m, mvalid, k0, k, g1);
            }
          }
          if (!L.first_of_chain) bwd_tail_finish(P, e, P.layer[b.li - 1], rb, k0, cols);
        }
        tc_fence_before();
        mbar_arrive(pipe.tempty);
        ++gi;
      }
    }
    trace_stamp(P, tslot);
    if (plo <= pb + 1 && pb + 2 < phi) chain_barrier(P.bar, target);
    trace_stamp(P, tslot);
  }

  // ------------------------------------------------------------------ reduce: dW, db, dX_sum
  if (warp >= 2 && plo <= 1 + 2 * P.Lmax && 1 + 2 * P.Lmax < phi) {
    const int n_flat = flat_items(P, false);
    const int n_dx = P.dX_sum ? ch_tiles(P.M * P.chain[0].K0, CH_EW) : 0;
    for (int item = blockIdx.x; item < n_flat + P.n_total + n_dx; item += gridDim.x) {
      if (item < n_flat) {
        int li, blk;
        flat_decode(P, false, item, li, blk);
        const ChLayer &L = P.layer[li];
        const int n = L.N * L.K;
#pragma unroll 1
        for (int j = 0; j < CH_EW / CH_EPI; ++j) {
          const int idx = blk * CH_EW + j * CH_EPI + e.et;
          if (idx >= n) break;
          float sw = 0.f;
          int c = 0;
          for (; c + 8 <= chunks; c += 8) {      // 8 loads in flight, added in chunk order
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcg(L.dWpart + (size_t)(c + u) * n + idx);
#pragma unroll
            for (int u = 0; u < 8; ++u) sw += v[u];
          }
          for (; c < chunks; ++c) sw += __ldcg(L.dWpart + (size_t)c * n + idx);
          L.dW[idx] = sw;
        }
      } else if (item < n_flat + P.n_total) {
        const ChLayer &L = P.layer[item - n_flat];
        if (L.db && !L.has_bn)
          for (int n = e.et; n < L.N; n += CH_EPI) {
            double sb = 0.0, unused = 0.0;
            ordered_sum2(L.bstat + (size_t)n * 2, (size_t)L.N * 2, P.RB, sb, unused);
            L.db[n] = (float)sb;
          }
      } else {
        const int blk = item - n_flat - P.n_total, n = P.M * P.chain[0].K0;
#pragma unroll 1
        for (int j = 0; j < CH_EW / CH_EPI; ++j) {
          const int idx = blk * CH_EW + j * CH_EPI + e.et;
          if (idx >= n) break;
          float sx = 0.f;
          for (int c = 0; c < P.n_chains; ++c) sx += __ldcg(P.chain[c].dX + idx);
          P.dX_sum[idx] = sx;
        }
      }
    }
  }
  trace_stamp(P, tslot);
  chain_teardown(cx, P);
}

// ============================================================================================== host side
constexpr size_t CH_SMEM = (size_t)CH_STAGES * CH_STAGE + 2 * 128 * 33 * 4 + 2 * 4 * 32 * 8 + 5 * 32 * 4 + 16 * 8 + sizeof(ChHead) + 64 + 1024;

// planes [2][rows][ld] with `inner` valid columns -> boxes of box_rows x 32 columns of one plane
static int g_map_err = 0;
static unsigned long long *g_trace_buf = nullptr;
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

// B tiles are 16 or 32 rows wide: the work per item is the row-wise epilogue (thread == row), so narrow tiles = more CTAs busy
static inline int tile_for(int n) { return n <= 16 ? 16 : 32; }
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

constexpr size_t X_FLAGS = 0, X_EPOCH = 64, X_REGION = 256;   // layout of the data-parallel exchange memory

static int build_params(const fr_chain *chains, int n_chains, int64_t M, int training, int need_grad, bool backward,
                        const uint64_t *seed_dev, uint32_t *bar, float *dX_sum, const fr_chain_dp *dp, ChParams &PP,
                        const char *who) {
  ChHead &P = PP.h;
  FR_REQUIRE(chains && n_chains >= 1 && n_chains <= CH_MAX_CHAINS && M >= 1 && M <= (1 << 22) && bar, "%s: bad argument", who);
  memset(&PP, 0, sizeof(PP));
  const ChSizes z = ch_sizes(M);
  P.n_chains = n_chains;
  P.M = (int)M;
  P.Mpad = (int)z.Mpad;
  P.RB = (int)z.RB;
  P.world = 1;
  P.RBg = P.RB;
  P.Mg = P.M;
  P.x_parity = -1;
  P.phase_lo = 0;
  P.phase_hi = 1 << 20;
  if (dp && dp->world > 1) {
    FR_REQUIRE(dp->world <= FR_MAX_RANKS && dp->rank >= 0 && dp->rank < dp->world, "%s: bad rank / world", who);
    for (int k = 0; k < dp->world; ++k) FR_REQUIRE(dp->xchg[k], "%s: exchange memory of rank %d missing", who, k);
    P.rank = dp->rank;
    P.world = dp->world;
    P.rb_base = dp->rank * P.RB;           // every rank holds the same number of rows
    P.RBg = dp->world * P.RB;
    P.Mg = dp->world * P.M;
    P.row_base = dp->rank * P.M;
    for (int k = 0; k < dp->world; ++k) P.peer[k] = (char *)dp->xchg[k];
    P.x_epoch = X_EPOCH;
    P.x_region = X_REGION;
    P.x_parity = dp->barriers ? -1 : (dp->parity & 1);
  }
  P.training = training ? 1 : 0;
  P.need_grad = (need_grad || backward) ? 1 : 0;
  P.seed_dev = (const unsigned long long *)seed_dev;
  P.bar = bar;
  P.dX_sum = dX_sum;
  {
    static unsigned long long *trace_buf = nullptr;
    static int trace_on = -1;
    if (trace_on < 0) {
      const char *e = getenv("FR_CHAIN_TRACE");
      trace_on = (e && e[0] == '1') ? 1 : ((e && e[0] == '2') ? 2 : 0);
      if (trace_on) cudaMalloc(&trace_buf, 256 * sizeof(unsigned long long));
    }
    P.trace = (trace_on == 1 && !backward) || (trace_on == 2 && backward) ? trace_buf : nullptr;
    g_trace_buf = trace_buf;
  }
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
      L.bn_repeat = s.bn_repeat > 1 ? s.bn_repeat : 1;
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
  if (P.world > 1) {      // exchange-memory slots of the BatchNorm layers' column-sum partials
    size_t off = 0;
    for (int i = 0; i < total; ++i) {
      P.layer[i].x_off = off;
      if (P.layer[i].has_bn) off += (((size_t)P.RBg * P.layer[i].N * 2 * sizeof(double)) + 255) & ~(size_t)255;
    }
    P.x_stride = off;
    if (X_REGION + 2 * off > dp->xchg_bytes) {
      set_error("%s: exchange memory too small (%zu bytes needed)", who, X_REGION + 2 * off);
      return FR_ERR_WORKSPACE;
    }
  }
  // tensor maps
  for (int i = 0; i < total; ++i) {
    ChLayer &L = P.layer[i];
    bool ok = true;
    if (!backward) {
      ok = ok && make_map3(&PP.map[i][0], L.Xin, L.K, P.M, L.K, (size_t)P.M * L.K, TCM);
      ok = ok && make_map3(&PP.map[i][1], L.Wp, L.K, L.N, L.K, (size_t)L.N * L.K, L.NT);
    } else {
      ok = ok && make_map3(&PP.map[i][0], L.DZ, L.N, P.M, L.ldn, (size_t)P.M * L.ldn, TCM);
      ok = ok && make_map3(&PP.map[i][1], L.Wt, L.N, L.K, L.ldn, (size_t)L.K * L.ldn, L.KT);
      ok = ok && make_map3(&PP.map[i][2], L.DZT, P.M, L.N, P.Mpad, (size_t)L.N * P.Mpad, TCM);
      ok = ok && make_map3(&PP.map[i][3], L.XinT, P.M, L.K, P.Mpad, (size_t)L.K * P.Mpad, L.KT);
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

// ---------------------------------------------------------------- data-parallel: cross-GPU flag barrier between segments
__device__ __forceinline__ unsigned long long chx_ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void chx_st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
struct ChPeers {
  char *base[FR_MAX_RANKS];
};
// thread k talks to rank k: publish "arrived at exchange #e" in slot `rank` of k's flag array, wait until k has published
// e in mine.  The partial sums this rank stored into peer memory in the kernel before are ordered before the flag (kernel
// boundary + system fence + release).  Also advances the epoch word whose parity selects the exchange region.
static __global__ void k_chain_xbar(ChPeers px, int rank, int world, int32_t *status) {
  __shared__ unsigned long long e;
  unsigned long long *epoch = (unsigned long long *)(px.base[rank] + X_EPOCH);
  if (threadIdx.x == 0) e = *epoch + 1ull;
  __syncthreads();
  const int k = threadIdx.x;
  if (k < world) {
    __threadfence_system();
    chx_st_release_sys((unsigned long long *)(px.base[k] + X_FLAGS) + rank, e);
    const unsigned long long *mine = (const unsigned long long *)(px.base[rank] + X_FLAGS) + k;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (chx_ld_acquire_sys(mine) < e) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t - t0 > 20000000000ull) {     // 20 s: a peer died; do not hang the GPU
        if (status) atomicOr(status, FR_FLAG_XCHG_TIMEOUT);
        break;
      }
    }
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x == 0) *epoch = e;
}

// phase windows of a (data-parallel) launch sequence: a cut wherever BatchNorm partials have to cross the ranks, i.e.
// forward: after the GEMM phase of every layer position with a BatchNorm layer; backward: before every BatchNorm phase
static int chain_cuts(const ChHead &P, bool backward, int *cuts /* [2 * CH_MAX_LAYERS + 4] */) {
  int n = 0;
  cuts[n++] = 0;
  if (P.world > 1) {
    for (int l = 0; l < P.Lmax; ++l) {
      if (!backward && fwd_bn_phase(P, l)) cuts[n++] = 2 + 2 * l;
      if (backward && bn_items(P, l, true) > 0) cuts[n++] = 1 + 2 * l;
    }
  }
  cuts[n] = 2 + 2 * P.Lmax;
  return n;     // number of segments; segment i = [cuts[i], cuts[i+1])
}

static int run_chain(const void *kernel, const char *name, ChParams &P, bool backward, int want, const fr_chain_dp *dp,
                     cudaStream_t st) {
  int cuts[2 * CH_MAX_LAYERS + 4];
  const int nseg = chain_cuts(P.h, backward, cuts);
  const bool emu = dp && dp->world > 1 && !dp->barriers;
  for (int i = 0; i < nseg; ++i) {
    if (emu && dp->segment >= 0 && dp->segment != i) continue;
    P.h.phase_lo = cuts[i];
    P.h.phase_hi = cuts[i + 1];
    int rc = launch_chain(kernel, name, P, want, st);
    if (rc) return rc;
    if (P.h.world > 1 && dp->barriers && i + 1 < nseg) {
      ChPeers px;
      for (int k = 0; k < FR_MAX_RANKS; ++k) px.base[k] = k < P.h.world ? P.h.peer[k] : nullptr;
      FR_LAUNCH(k_chain_xbar, 1, 32, 0, st, px, P.h.rank, P.h.world, dp->status_flags);
    }
  }
  FR_LAUNCH_CHECK();
  return FR_OK;
}

}  // namespace fr

extern "C" {

int fr_mlp_chain_trace(uint64_t *out_host, int32_t n) {
  if (!fr::g_trace_buf || n > 256) return FR_ERR_INVALID;
  return cudaMemcpy(out_host, fr::g_trace_buf, (size_t)n * 8, cudaMemcpyDeviceToHost) == cudaSuccess ? FR_OK : FR_ERR_CUDA;
}

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

int fr_mlp_chain_forward_dp(const fr_chain *chains, int32_t n_chains, int64_t M, int32_t training, int32_t need_grad,
                            const uint64_t *seed_dev, uint32_t *barrier_words, const fr_chain_dp *dp, void *stream) {
  static fr::ChParams P;     // 19 KB: not on the stack of a ctypes caller thread
  int rc = fr::build_params(chains, n_chains, M, training, need_grad, false, seed_dev, barrier_words, nullptr, dp, P,
                            "fr_mlp_chain_forward");
  if (rc) return rc;
  int want = fr::import_items(P.h, false) + fr::flat_items(P.h, false);
  for (int l = 0; l < P.h.Lmax; ++l) {
    const int g = fr::fwd_gemm_items(P.h, l);
    if (g > want) want = g;
  }
  return fr::run_chain((const void *)fr::k_mlp_chain_fwd, "k_mlp_chain_fwd", P, false, want, dp, (cudaStream_t)stream);
}

int fr_mlp_chain_backward_dp(const fr_chain *chains, int32_t n_chains, int64_t M, const uint64_t *seed_dev, float *dX_sum,
                             uint32_t *barrier_words, const fr_chain_dp *dp, void *stream) {
  static fr::ChParams P;
  int rc = fr::build_params(chains, n_chains, M, 1, 1, true, seed_dev, barrier_words, dX_sum, dp, P, "fr_mlp_chain_backward");
  if (rc) return rc;
  int want = fr::import_items(P.h, true) + fr::flat_items(P.h, true);
  for (int s = 0; s < P.h.Lmax; ++s) {
    const int g = fr::bwd_gemm_items(P.h, s);
    if (g > want) want = g;
  }
  return fr::run_chain((const void *)fr::k_mlp_chain_bwd, "k_mlp_chain_bwd", P, true, want, dp, (cudaStream_t)stream);
}

int fr_mlp_chain_forward(const fr_chain *chains, int32_t n_chains, int64_t M, int32_t training, int32_t need_grad,
                         const uint64_t *seed_dev, uint32_t *barrier_words, void *stream) {
  return fr_mlp_chain_forward_dp(chains, n_chains, M, training, need_grad, seed_dev, barrier_words, nullptr, stream);
}

int fr_mlp_chain_backward(const fr_chain *chains, int32_t n_chains, int64_t M, const uint64_t *seed_dev, float *dX_sum,
                          uint32_t *barrier_words, void *stream) {
  return fr_mlp_chain_backward_dp(chains, n_chains, M, seed_dev, dX_sum, barrier_words, nullptr, stream);
}

int fr_mlp_chain_segments(const fr_chain *chains, int32_t n_chains, int32_t training, int32_t backward, int32_t world) {
  // number of launch segments of a data-parallel chain call (cuts where BatchNorm partials cross the ranks)
  if (!chains || n_chains < 1 || n_chains > fr::CH_MAX_CHAINS) return -1;
  int n = 1;
  if (world <= 1 || !(training || backward)) return n;
  int Lmax = 0;
  for (int c = 0; c < n_chains; ++c) Lmax = chains[c].n_layers > Lmax ? chains[c].n_layers : Lmax;
  for (int pos = 0; pos < Lmax; ++pos) {
    bool bn = false;
    for (int c = 0; c < n_chains; ++c) {
      const int l = backward ? chains[c].n_layers - 1 - pos : pos;
      if (l >= 0 && l < chains[c].n_layers && chains[c].layer[l].has_bn) bn = true;
    }
    if (bn) ++n;
  }
  return n;
}

}  // extern "C"
