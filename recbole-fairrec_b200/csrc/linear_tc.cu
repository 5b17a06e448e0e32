// MLP layer forward on the 5th-generation tensor cores:  Y[M,N] = act(dropout(X)[M,K] . W[N,K]^T + b)
// tcgen05.mma kind::tf32 fed by TMA, fp32 accumulator in TMEM, 3xTF32 split for fp32-level accuracy (the 1e-5 parity bar).
//
// Reference being replaced: recbole/model/layers.py:60-68 (nn.Dropout -> nn.Linear -> activation) for the filter,
// discriminator and scorer MLPs of pfcn_*.py / fairgo_*.py whenever the layer shape fits the tile rules
// (K % 32 == 0, N % 32 == 0); other shapes (the 16/8/4/1-wide tails of the discriminators)
// stay on the CUDA-core kernel k_gemm (mlp.cu).  Both operands are K-major ([rows, K] row-major), exactly the layout
// of the full-sort scorer (fullsort_tc.cu), whose PTX wrappers this kernel shares (tc_common.cuh).
//
// One CTA = 128 rows of X x a 64- (or 32-) column tile of N x the whole K: warp 0 streams the raw K-blocks (32 k-columns =
// one 128-byte swizzle row) of X and W through a shared-memory ring with TMA, warps 2-5 split them, warp 1 issues
// 4 k-steps x 3 products (hi.hi + hi.lo + lo.hi) per K-block into ONE TMEM accumulator, and warps 2-5 then own one output
// row per thread (TMEM lane): tcgen05.ld 32 columns -> + bias -> activation -> 128-byte row segments of Y.
// The raw fp32 tiles are split into their TF32 hi/lo planes IN shared memory by the four epilogue warps while they wait
// for the accumulator (dropout mask folded in: a counter-based hash of (seed, layer, element), identical to k_gemm's, so
// the backward kernels see the same mask) -- no extra pass over X or W, no workspace, one launch per layer.
#include "act.cuh"
#include "tc_common.cuh"

namespace fr {

constexpr int LT_THREADS = 192;
constexpr int LT_MAX_STAGES = 4;
constexpr int LT_A_BYTES = TCM * TCKB * 4;   // 16 KB per plane per K-block

struct LinTcArgs {
  const float *bias;
  float *Y;
  int M, N, K, NT, act, stages, tmem_cols;
  float drop_p;
  unsigned long long seed;
  const unsigned long long *seed_dev;
  int layer;
};

// split one raw fp32 tile (TMA-written, 128-byte swizzled rows of 32 floats) in place into its TF32 hi plane and write
// the lo plane beside it; position-wise, so the swizzle is irrelevant -- except for the dropout mask, whose counter
// needs the element's logical (row, k): 16-byte chunk c of row r sits at chunk c ^ (r & 7)
__device__ __forceinline__ void split_tile(float4 *raw, float4 *lo, int n_f4, int tid, int row0, int k0, int K, int M,
                                           float drop_p, unsigned long long seed, int layer) {
  for (int q = tid; q < n_f4; q += 128) {
    float4 x = raw[q];
    if (drop_p > 0.f) {
      const int r = q >> 3, c = (q & 7) ^ (r & 7);
      const int m = row0 + r;
      if (m < M) {
        const uint32_t e = (uint32_t)(m * K + k0 + c * 4);
        x.x *= drop_scale(seed, layer, e, drop_p);
        x.y *= drop_scale(seed, layer, e + 1, drop_p);
        x.z *= drop_scale(seed, layer, e + 2, drop_p);
        x.w *= drop_scale(seed, layer, e + 3, drop_p);
      }
    }
    float4 h, l;
    h.x = rn_tf32(x.x); l.x = rn_tf32(x.x - h.x);
    h.y = rn_tf32(x.y); l.y = rn_tf32(x.y - h.y);
    h.z = rn_tf32(x.z); l.z = rn_tf32(x.z - h.z);
    h.w = rn_tf32(x.w); l.w = rn_tf32(x.w - h.w);
    raw[q] = h;
    lo[q] = l;
  }
}

static __global__ void __launch_bounds__(LT_THREADS, 1)
    k_linear_tc(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, LinTcArgs a) {
  extern __shared__ __align__(1024) unsigned char lt_smem[];
  const int nkb = a.K / TCKB;
  const int b_bytes = a.NT * TCKB * 4;                   // one W plane K-block: NT rows x 128 bytes
  const int stage_bytes = 2 * LT_A_BYTES + 2 * b_bytes;  // [X raw->hi][X lo][W raw->hi][W lo]
  uint64_t *bars = (uint64_t *)(lt_smem + (size_t)a.stages * stage_bytes);
  uint64_t *raw = bars, *full = bars + LT_MAX_STAGES, *empty = bars + 2 * LT_MAX_STAGES, *tfull = bars + 3 * LT_MAX_STAGES;
  uint32_t *tmem_slot = (uint32_t *)(tfull + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TCM, n0 = blockIdx.y * a.NT;

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&raw[s], 1);
      mbar_init(&full[s], 128);     // every converter thread arrives
      mbar_init(&empty[s], 1);
    }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(a.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {   // ===================================================== TMA producer: raw fp32 tiles
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % a.stages;
        const uint32_t ph = (kb / a.stages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        unsigned char *st = lt_smem + (size_t)s * stage_bytes;
        mbar_expect_tx(&raw[s], (uint32_t)(LT_A_BYTES + b_bytes));
        tma_load_2d(st, &map_x, kb * TCKB, m0, &raw[s]);
        tma_load_2d(st + 2 * LT_A_BYTES, &map_w, kb * TCKB, n0, &raw[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {   // ===================================================== MMA issuer
      const uint32_t idesc = umma_idesc_tf32(TCM, a.NT);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % a.stages;
        const uint32_t ph = (kb / a.stages) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t x_hi = smem_u32(lt_smem + (size_t)s * stage_bytes), x_lo = x_hi + LT_A_BYTES;
        const uint32_t w_hi = x_hi + 2 * LT_A_BYTES, w_lo = w_hi + b_bytes;
#pragma unroll
        for (int k = 0; k < TCKB / 8; ++k) {   // UMMA K = 8 tf32 = 32 bytes along the swizzled row
          const uint32_t off = k * 32;
          const uint32_t first = (kb == 0 && k == 0) ? 0u : 1u;
          umma_tf32(tmem_base, umma_desc_sw128(x_hi + off), umma_desc_sw128(w_hi + off), idesc, first);
          umma_tf32(tmem_base, umma_desc_sw128(x_hi + off), umma_desc_sw128(w_lo + off), idesc, 1u);
          umma_tf32(tmem_base, umma_desc_sw128(x_lo + off), umma_desc_sw128(w_hi + off), idesc, 1u);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(tfull);
    }
  } else {
    // ======================================================= warps 2-5: TF32 hi/lo split of every stage, then the epilogue
    const int tid = threadIdx.x - 64;
    const unsigned long long seed = a.seed + (a.seed_dev ? *a.seed_dev : 0ull);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % a.stages;
      const uint32_t ph = (kb / a.stages) & 1;
      mbar_wait(&raw[s], ph);
      unsigned char *st = lt_smem + (size_t)s * stage_bytes;
      split_tile((float4 *)st, (float4 *)(st + LT_A_BYTES), LT_A_BYTES / 16, tid, m0, kb * TCKB, a.K, a.M, a.drop_p, seed,
                 a.layer);
      split_tile((float4 *)(st + 2 * LT_A_BYTES), (float4 *)(st + 2 * LT_A_BYTES + b_bytes), b_bytes / 16, tid, 0, 0, a.K,
                 0, 0.f, 0ull, 0);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      mbar_arrive(&full[s]);
    }
    const int quarter = warp & 3;               // TMEM lanes [32*quarter, +32) are this warp's: thread == output row
    const int row = m0 + quarter * 32 + lane;
    mbar_wait(tfull, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    for (int w = 0; w < a.NT / 32; ++w) {
      uint32_t v[32];
      tmem_ld32(taddr + w * 32, v);
      if (row < a.M) {
        const int nb = n0 + w * 32;
        float *dst = a.Y + (size_t)row * a.N + nb;
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          float4 o;
          o.x = act_fwd(__uint_as_float(v[c]) + (a.bias ? __ldg(a.bias + nb + c) : 0.f), a.act);
          o.y = act_fwd(__uint_as_float(v[c + 1]) + (a.bias ? __ldg(a.bias + nb + c + 1) : 0.f), a.act);
          o.z = act_fwd(__uint_as_float(v[c + 2]) + (a.bias ? __ldg(a.bias + nb + c + 2) : 0.f), a.act);
          o.w = act_fwd(__uint_as_float(v[c + 3]) + (a.bias ? __ldg(a.bias + nb + c + 3) : 0.f), a.act);
          *(float4 *)(dst + c) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols));
  }
}

bool tc_linear_eligible(int64_t M, int K, int N) {
  static int disabled = -1;
  if (disabled < 0) {
    const char *e = getenv("FR_LINEAR_NO_TC");
    disabled = (e && e[0] == '1') ? 1 : 0;
  }
  return !disabled && M >= 1 && K % TCKB == 0 && K >= TCKB && K <= 4096 && N % 32 == 0 && N >= 32 &&
         M < (int64_t)1 << 30 && (int64_t)M * K < (int64_t)1 << 32;
}

int tc_linear_forward(const float *X, const float *W, const float *b, float *Y, int64_t M, int K, int N, int act,
                      float drop_p, unsigned long long seed, const unsigned long long *seed_dev, int layer, cudaStream_t st) {
  const int NT = (N % 64 == 0) ? 64 : 32;     // column tile per CTA: small tiles = more CTAs (the layers are latency-bound)
  CUtensorMap m_x, m_w;
  if (!make_map(&m_x, X, M, K, TCM) || !make_map(&m_w, W, N, K, NT)) {
    set_error("fr_linear_forward: cuTensorMapEncodeTiled failed");
    return FR_ERR_CUDA;
  }
  const int nkb = K / TCKB;
  const int stage_bytes = 2 * LT_A_BYTES + 2 * NT * TCKB * 4;
  int stages = (232448 - 2048) / stage_bytes;
  if (stages > LT_MAX_STAGES) stages = LT_MAX_STAGES;
  if (stages > nkb) stages = nkb;
  const size_t smem = (size_t)stages * stage_bytes + 256 + 1024;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    FR_CUDA_OK(cudaFuncSetAttribute(k_linear_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  LinTcArgs a{b, Y, (int)M, N, K, NT, act, stages, NT < 32 ? 32 : NT, drop_p, seed, seed_dev, layer};
  dim3 grid((unsigned)((M + TCM - 1) / TCM), (unsigned)(N / NT));
  FR_LAUNCH(k_linear_tc, grid, LT_THREADS, smem, st, m_x, m_w, a);
  return FR_OK;
}

}  // namespace fr
