// Small-MLP building blocks for the NCF / PFCN towers (kernel K6) on sm_100a, plus the
// NFCF differential-fairness regulariser.
//
// Reference being replaced (paths relative to the reference root):
//   recbole/model/layers.py:58-70 MLPLayers (Dropout -> Linear -> activation per layer) and its autograd
//   recbole/model/fair_recommender/nfcf.py:69-74 forward (emb || emb -> tower -> sigmoid), 99-110 BCE loss,
//   nfcf.py:76-97 get_differential_fairness (item x group sums over the batch's positives via index_put_)
//   nn.Embedding dense backward of the two towers' inputs; torch.optim.Adam for the tower parameters
//
// The towers are tiny (widths <= 256, M = batch rows): every layer is a 64x64x16 register-tiled fp32 GEMM on the CUDA
// cores with the bias / activation / dropout fused into the epilogue or the operand loader.  fp32 FMA keeps the
// 1e-5 parity bar against the reference's sgemm without a 3xTF32 split; at these sizes the step is launch-bound,
// not math-bound.  Weight gradients reduce over the batch in fixed 256-row chunks (partials + ordered sum), the
// item x group statistics reuse the sorted-segment machinery of sort.cu: no floating-point atomics anywhere.
#include "act.cuh"
#include "sort.cuh"

namespace fr {

// ---------------------------------------------------------------- gather / concat of the two embedding rows
__global__ void __launch_bounds__(256)
    k_gather_concat(const float *__restrict__ U, const float *__restrict__ I, const int32_t *__restrict__ uid,
                    const int32_t *__restrict__ iid, int64_t M, int d, float *__restrict__ X) {
  const int dq = d >> 2;
  const int64_t nq = M * 2 * dq;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = q / (2 * dq);
    const int c = (int)(q % (2 * dq));
    const float4 v = c < dq ? __ldg((const float4 *)(U + (size_t)uid[r] * d) + c)
                            : __ldg((const float4 *)(I + (size_t)iid[r] * d) + (c - dq));
    *((float4 *)(X + r * 2 * d) + c) = v;
  }
}

// ---------------------------------------------------------------- tiled GEMM
// C[M,N] = epilogue( sum_k A'[m,k] * B'[k,n] )
//   A' = A[m*lda + k] scaled by the dropout mask of (layer, m*K + k) when drop_p > 0
//   B' = kTB ? B[n*ldb + k] (weight [N,K], forward) : B[k*ldb + n] (weight [K,N] seen from dY.W, backward-data)
//   epilogue: + bias[n], activation; or (backward-data) * dropout mask of the INPUT element and * act'(Yin) of the
//   previous layer's output when `prev_out` is given.
struct GemmArgs {
  const float *A, *B, *bias;
  float *C;
  int M, N, K, lda, ldb, ldc;
  int act;                      // forward epilogue activation
  const float *prev_out;        // backward-data: multiply by act_bwd(prev_out[m,n], prev_act)
  int prev_act;
  float drop_p;                 // forward: mask on A ; backward-data: mask on C (same counter: layer, m*N + n)
  unsigned long long seed;
  int layer;
  int drop_on_c;
  const float *a_out;           // backward-data: A' = A * act_bwd(a_out[m,k], a_act) (activation backward fused in the loader)
  int a_act;
  const unsigned long long *seed_dev;   // optional device-resident seed offset (CUDA-graph replays draw new masks)
};

// 64x64x32 tiles, 4x4 outputs per thread; the next k-slab is fetched into registers while the current one is
// multiplied out of shared memory (these GEMMs are a handful of k-steps long: exposed load latency, not FLOPs, bounds them)
constexpr int kGK = 32;   // k-depth of one slab
template <bool kTB>
__global__ void __launch_bounds__(256) k_gemm(GemmArgs a) {
  __shared__ __align__(16) float As[kGK][64 + 4];
  __shared__ __align__(16) float Bs[kGK][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const unsigned long long seed = a.seed + (a.seed_dev ? *a.seed_dev : 0ull);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // raw operands of the next k-slab: every global load is issued before any is consumed (activation backward and
  // dropout are applied when the registers are parked in shared memory, so no load waits behind a data-dependent op)
  float ra[8], rao[8], rb[8];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int f = threadIdx.x + t * 256;          // A tile: 64 rows x 32 k ; B tile: 32 k x 64 n
      const int r = f >> 5, kk = f & 31;
      const int m = m0 + r, k = k0 + kk;
      const bool va = m < a.M && k < a.K;
      ra[t] = va ? a.A[(size_t)m * a.lda + k] : 0.f;
      rao[t] = (va && a.a_out) ? a.a_out[(size_t)m * a.lda + k] : 1.f;
      int n, kb;
      if (kTB) { n = f >> 5; kb = f & 31; } else { kb = f >> 6; n = f & 63; }
      const int gn = n0 + n, k2 = k0 + kb;
      const bool vb = gn < a.N && k2 < a.K;
      rb[t] = vb ? (kTB ? a.B[(size_t)gn * a.ldb + k2] : a.B[(size_t)k2 * a.ldb + gn]) : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < a.K; k0 += kGK) {
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int f = threadIdx.x + t * 256;
      float v = ra[t];
      if (a.a_out) v *= act_bwd(rao[t], a.a_act);
      if (a.drop_p > 0.f && !a.drop_on_c)
        v *= drop_scale(seed, a.layer, (uint32_t)((m0 + (f >> 5)) * a.K + k0 + (f & 31)), a.drop_p);
      As[f & 31][f >> 5] = v;
      if (kTB) Bs[f & 31][f >> 5] = rb[t]; else Bs[f >> 6][f & 63] = rb[t];
    }
    __syncthreads();
    if (k0 + kGK < a.K) fetch(k0 + kGK);
    const int kmax = min(kGK, a.K - k0);
#pragma unroll 8
    for (int kk = 0; kk < kmax; ++kk) {
      const float4 a4 = *(const float4 *)&As[kk][ty * 4], b4 = *(const float4 *)&Bs[kk][tx * 4];
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= a.N) continue;
      float v = acc[i][j];
      if (a.bias) v += a.bias[n];
      v = act_fwd(v, a.act);
      if (a.prev_out) v *= act_bwd(a.prev_out[(size_t)m * a.ldc + n], a.prev_act);
      if (a.drop_p > 0.f && a.drop_on_c) v *= drop_scale(seed, a.layer, (uint32_t)(m * a.N + n), a.drop_p);
      a.C[(size_t)m * a.ldc + n] = v;
    }
  }
}

// ---------------------------------------------------------------- weight / bias gradients
// dW[n,k] = sum_m dY'[m,n] * Xdrop[m,k] over one row chunk per blockIdx.z -> part[z][N][K]; db likewise.
// dY' = dY * act'(Yout) when y_out is given (the layer's activation backward fused into the loader).
struct WgradArgs {
  const float *dY, *X;
  float *part_w;                // [chunks, N, K]
  double *part_b;               // [chunks, N]: bias-gradient column sums are carried in float64 (a column sum of +-O(1/B)
                                // terms cancels ~1000x in the BPR tower; float64 makes its rounding negligible)
  int M, N, K, ldy, ldx;
  float drop_p;
  unsigned long long seed;
  int layer;
  const float *y_out;
  int y_act;
  const unsigned long long *seed_dev;
  // optional fused chunk reduction: the LAST CTA of an (n,k) tile to finish (self-resetting ticket) adds the tile's chunk
  // partials in chunk order into dW / db -- fixed order whichever CTA ends up last, so still bit-reproducible
  int *tickets;
  float *dW, *db;
  int chunk;                    // rows per chunk (wgrad_chunk(M))
};
constexpr int kWgradChunk = 128;   // smallest chunk: workspace sizing
// batches of a few thousand rows: 128-row chunks (more CTAs, 4 slabs each); full-table passes (FairGo): 256-row chunks
static inline int wgrad_chunk(int64_t M) { return M <= 4096 ? 128 : 256; }

__global__ void __launch_bounds__(256) k_wgrad(WgradArgs a) {
  __shared__ __align__(16) float Ys[kGK][64 + 4];
  __shared__ __align__(16) float Xs[kGK][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int n0 = blockIdx.y * 64, k0 = blockIdx.x * 64;
  const int mlo = blockIdx.z * a.chunk, mhi = min(a.M, mlo + a.chunk);
  const unsigned long long seed = a.seed + (a.seed_dev ? *a.seed_dev : 0ull);
  float acc[4][4];
  double bacc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float ry[8], ryo[8], rx[8];
  auto fetch = [&](int mb) {   // raw loads only (see k_gemm)
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int f = threadIdx.x + t * 256;
      const int m = mb + (f >> 6), c = f & 63;
      const bool vy = m < mhi && n0 + c < a.N, vx = m < mhi && k0 + c < a.K;
      ry[t] = vy ? a.dY[(size_t)m * a.ldy + n0 + c] : 0.f;
      ryo[t] = (vy && a.y_out) ? a.y_out[(size_t)m * a.ldy + n0 + c] : 1.f;
      rx[t] = vx ? a.X[(size_t)m * a.ldx + k0 + c] : 0.f;
    }
  };
  fetch(mlo);
  for (int mb = mlo; mb < mhi; mb += kGK) {
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int f = threadIdx.x + t * 256;
      float y = ry[t], x = rx[t];
      if (a.y_out) y *= act_bwd(ryo[t], a.y_act);
      if (a.drop_p > 0.f) x *= drop_scale(seed, a.layer, (uint32_t)((mb + (f >> 6)) * a.K + k0 + (f & 63)), a.drop_p);
      Ys[f >> 6][f & 63] = y;
      Xs[f >> 6][f & 63] = x;
    }
    __syncthreads();
    if (mb + kGK < mhi) fetch(mb + kGK);
#pragma unroll 8
    for (int mm = 0; mm < kGK; ++mm) {     // rows past mhi were staged as zeros
      const float4 y4 = *(const float4 *)&Ys[mm][ty * 4], x4 = *(const float4 *)&Xs[mm][tx * 4];
      const float yv[4] = {y4.x, y4.y, y4.z, y4.w}, xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (tx == 0) bacc[i] += (double)yv[i];      // warp-uniform per half-warp; only column owners carry the sum
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(yv[i], xv[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
  float *pw = a.part_w + (size_t)blockIdx.z * a.N * a.K;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= a.N) continue;
    if (tx == 0 && blockIdx.x == 0) a.part_b[(size_t)blockIdx.z * a.N + n] = bacc[i];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < a.K) pw[(size_t)n * a.K + k] = acc[i][j];
    }
  }
  if (!a.tickets) return;
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  const int tile = blockIdx.y * gridDim.x + blockIdx.x;
  if (threadIdx.x == 0) is_last = atomicAdd(&a.tickets[tile], 1) == (int)gridDim.z - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const int chunks = gridDim.z;
  const size_t cs = (size_t)a.N * a.K;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= a.N) continue;
    if (tx == 0 && blockIdx.x == 0 && a.db) {
      double sb = 0.0;
#pragma unroll 8
      for (int c = 0; c < chunks; ++c) sb += __ldcg(a.part_b + (size_t)c * a.N + n);
      a.db[n] = (float)sb;
    }
    const int k = k0 + tx * 4;
    if ((a.K & 3) == 0 && k + 3 < a.K) {     // 128-bit reads, 8 chunks in flight, added in chunk order
      const float *src = a.part_w + (size_t)n * a.K + k;
      float4 sw = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
      for (int c = 0; c < chunks; ++c) {
        const float4 v = __ldcg((const float4 *)(src + c * cs));
        sw.x += v.x; sw.y += v.y; sw.z += v.z; sw.w += v.w;
      }
      *(float4 *)(a.dW + (size_t)n * a.K + k) = sw;
    } else {
      for (int j = 0; j < 4; ++j) {
        if (k + j >= a.K) continue;
        float sw = 0.f;
#pragma unroll 8
        for (int c = 0; c < chunks; ++c) sw += __ldcg(a.part_w + c * cs + (size_t)n * a.K + k + j);
        a.dW[(size_t)n * a.K + k + j] = sw;
      }
    }
  }
  if (threadIdx.x == 0) a.tickets[tile] = 0;
}

// out[i] = sum_c part[c][i] in chunk order (float partials: weight gradients; double partials: bias gradients)
template <typename T>
__global__ void k_sum_chunks(const T *__restrict__ part, int chunks, int64_t n, float *__restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    T s = (T)0;
    for (int c = 0; c < chunks; ++c) s += part[(size_t)c * n + i];
    out[i] = (float)s;
  }
}

// ---------------------------------------------------------------- NFCF head: sigmoid + BCE (+ DF regulariser)
// p = sigmoid(z); per-CTA partial of the BCE sum -> part[blk]; dp_bce[b] = (-(y/p) + (1-y)/(1-p)) / B
__global__ void __launch_bounds__(256)
    k_sigmoid_bce(const float *__restrict__ z, const float *__restrict__ label, int M, float *__restrict__ p_out,
                  float *__restrict__ dp, float *__restrict__ part) {
  __shared__ float sh[8];
  const int b = blockIdx.x * 256 + threadIdx.x;
  float l = 0.f;
  if (b < M) {
    const float p = 1.f / (1.f + expf(-z[b])), y = label[b];
    p_out[b] = p;
    l = -(y * fmaxf(logf(p), -100.f) + (1.f - y) * fmaxf(logf(1.f - p), -100.f));   // nn.BCELoss clamps the logs
    dp[b] = (-(y / p) + (1.f - y) / (1.f - p)) / (float)M;
  }
  l = warp_sum(l);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = l;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sh[i];
    part[blockIdx.x] = t;
  }
}

// compact the positives (label == 1): pos_idx[k] = b in batch order (stable), n_pos; single CTA (batches are small)
__global__ void __launch_bounds__(1024) k_compact_pos(const float *__restrict__ label, int M, int32_t *__restrict__ pos_idx,
                                                      int32_t *__restrict__ n_pos) {
  __shared__ int wsum[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int base = 0; base < M; base += 1024) {
    const int b = base + threadIdx.x;
    const int f = (b < M && label[b] == 1.f) ? 1 : 0;
    int inc = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    int wb = 0, tot = 0;
    for (int i = 0; i < 32; ++i) {
      const int t = wsum[i];
      if (i < w) wb += t;
      tot += t;
    }
    const int c = carry;
    if (f) pos_idx[c + wb + inc - 1] = b;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) n_pos[0] = carry;
}

__global__ void k_pos_keys(const int32_t *__restrict__ iid, const int32_t *__restrict__ pos_idx,
                           const int32_t *__restrict__ n_pos, uint32_t *__restrict__ keys) {
  const int n = *n_pos;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) keys[k] = (uint32_t)iid[pos_idx[k]];
}

// nfcf.py:76-97 for a binary attribute.  Warp per item segment of the (item-sorted) positives:
//   M_g = (sum_g p + 1/J) / (cnt_g + 1) ; eps_j = |ln M_0 - ln M_1| ; d eps_j / d p_b = +-sgn / M_g / (cnt_g + 1)
// seg_eps[j] = eps_j ; coef[j][g] = fair_weight * (1/J) * d eps_j / d(sum_g p).  The group of a row is the rank of
// its attribute value among the values present IN THE POSITIVES (torch.unique), found via min/max.
struct DfArgs {
  const float *p, *sst;
  const int32_t *pos_idx, *n_pos;
  const uint32_t *ord;          // item-sorted order of the positives (values index pos_idx)
  const int32_t *seg_off, *n_seg;
  float fair_weight;
  float *seg_eps, *coef;        // [J], [J, G]
  const int32_t *grp, *n_groups; // group (rank of the attribute value) of every positive, in pos_idx order; number of groups
  int32_t *flags;
};

// order-preserving integer keys of the positives' attribute values: sorted + segmented (sort.cu), they give
// torch.unique(sst[pos], return_inverse=True) -- the group of every positive is the rank of its value (nfcf.py:79)
__global__ void k_df_group_keys(const float *__restrict__ sst, const int32_t *__restrict__ pos_idx,
                                const int32_t *__restrict__ n_pos, uint32_t *__restrict__ keys) {
  const int n = *n_pos;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) keys[k] = f2ord(sst[pos_idx[k]]);
}

// nfcf.py:76-97 for any number of attribute values G <= kDfMaxG: per (positive item j, group g) M = (sum p + 1/J) /
// (count + 1); eps_j = max over the pairs g < g' of |ln M_g - ln M_g'|, the gradient going to the FIRST pair (loop order
// of nfcf.py:91-95) that attains it (torch.where(eps > running, ...) replaces only on a strict increase).
// Warp per item segment; one pass over the segment's positives per group (segments are a few rows long).
constexpr int kDfMaxG = 32;
__global__ void __launch_bounds__(256) k_df_segments(DfArgs a) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const int J = *a.n_seg;
  const int G = *a.n_groups;
  if (G > kDfMaxG) {
    if (warp == 0 && lane == 0) atomicOr(a.flags, FR_FLAG_TOO_MANY_GROUPS);
    return;
  }
  const float alpha = 1.f / (float)J;
  for (int j = warp; j < J; j += nwarps) {
    const int p0 = a.seg_off[j], p1 = a.seg_off[j + 1];
    float M = 1.f, cnt = 0.f;       // lane g holds M[j][g] and the count of group g
    for (int g = 0; g < G; ++g) {
      float s = 0.f, c = 0.f;
      for (int q = p0 + lane; q < p1; q += 32) {
        const int k = a.ord[q];
        if (a.grp[k] == g) {
          s += a.p[a.pos_idx[k]];
          c += 1.f;
        }
      }
      s = warp_sum(s);
      c = warp_sum(c);
      if (lane == g) {
        M = (s + alpha) / (c + 1.f);
        cnt = c;
      }
    }
    const float lM = logf(M);
    float eps = 0.f, sg = 0.f;
    int bi = -1, bj = -1;
    for (int x = 0; x < G; ++x) {
      const float lx = __shfl_sync(0xffffffffu, lM, x);
      for (int y = x + 1; y < G; ++y) {
        const float diff = lx - __shfl_sync(0xffffffffu, lM, y);
        const float e = fabsf(diff);
        if (e > eps) {
          eps = e;
          sg = (float)((diff > 0.f) - (diff < 0.f));
          bi = x;
          bj = y;
        }
      }
    }
    if (lane == 0) a.seg_eps[j] = eps;
    if (lane < G) {
      float kf = 0.f;
      if (lane == bi) kf = a.fair_weight * alpha * sg / M / (cnt + 1.f);
      if (lane == bj) kf = -a.fair_weight * alpha * sg / M / (cnt + 1.f);
      a.coef[(size_t)G * j + lane] = kf;
    }
  }
}

// loss = bce_sum / M + fair_weight * mean_j eps_j ; dz = (dp_bce + coef[seg(b)][g(b)]) * p (1 - p)
__global__ void __launch_bounds__(256)
    k_nfcf_finish(const float *__restrict__ bce_part, int n_part, int M, const float *__restrict__ seg_eps,
                  const int32_t *__restrict__ n_seg, int use_df, float fair_weight, float *__restrict__ loss) {
  __shared__ float sh[8];
  float s = 0.f, e = 0.f;
  for (int i = threadIdx.x; i < n_part; i += 256) s += bce_part[i];
  const int J = use_df ? *n_seg : 0;
  for (int j = threadIdx.x; j < J; j += 256) e += seg_eps[j];
  s = warp_sum(s);
  e = warp_sum(e);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  float st = 0.f;
  if (threadIdx.x == 0) for (int i = 0; i < 8; ++i) st += sh[i];
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = e;
  __syncthreads();
  if (threadIdx.x == 0) {
    float et = 0.f;
    for (int i = 0; i < 8; ++i) et += sh[i];
    loss[0] = st / (float)M + (use_df && J > 0 ? fair_weight * (et / (float)J) : 0.f);
  }
}

__global__ void k_df_scatter_coef(DfArgs a, const int32_t *__restrict__ seg_id, float *__restrict__ dp) {
  const int n = *a.n_pos;
  const int G = *a.n_groups;
  if (G > kDfMaxG) return;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
    const int k = a.ord[q];
    dp[a.pos_idx[k]] += a.coef[(size_t)G * seg_id[q] + a.grp[k]];
  }
}

// dz[b] = dp[b] * p (1-p) * grad_scale   (sigmoid backward; the tower's last ReLU is handled by the GEMM epilogue)
__global__ void k_sigmoid_bwd(const float *__restrict__ dp, const float *__restrict__ p, const float *__restrict__ last_out,
                              int M, float grad_scale, float *__restrict__ dz) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < M) dz[b] = dp[b] * p[b] * (1.f - p[b]) * grad_scale * (last_out[b] > 0.f ? 1.f : 0.f);
}

// ---------------------------------------------------------------- dense embedding gradient from row gradients
// dTable[key] = sum of dX[b, col0 : col0+d] over the rows b with that key, in (stable) sorted order; one warp per
// segment.  The dense gradient is zeroed first by the caller (cudaMemsetAsync).
__global__ void __launch_bounds__(256)
    k_segment_sum_rows(const uint32_t *__restrict__ skey, const uint32_t *__restrict__ ord,
                       const int32_t *__restrict__ seg_off, const int32_t *__restrict__ n_seg,
                       const float *__restrict__ dX, int ldx, int col0, int d, float *__restrict__ dTable) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const int S = *n_seg;
  for (int s = warp; s < S; s += nwarps) {
    const int p0 = seg_off[s], p1 = seg_off[s + 1];
    const uint32_t key = skey[p0];
    for (int c = lane; c < d; c += 32) {
      float acc = 0.f;
      for (int q = p0; q < p1; ++q) acc += dX[(size_t)ord[q] * ldx + col0 + c];
      dTable[(size_t)key * d + c] = acc;
    }
  }
}

// ---------------------------------------------------------------- generic dense Adam (torch.optim.Adam, L2 form)
__global__ void __launch_bounds__(256)
    k_adam_flat(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v, int64_t n,
                int step, double lr, double beta1, double beta2, double eps, double wd) {
  __shared__ float sc[2];
  if (threadIdx.x == 0) {
    sc[0] = (float)(-lr / (1.0 - pow(beta1, (double)step)));
    sc[1] = (float)sqrt(1.0 - pow(beta2, (double)step));
  }
  __syncthreads();
  const float neg_step = sc[0], bc2s = sc[1], w1 = (float)(1.0 - beta1), w2 = (float)(1.0 - beta2), b2 = (float)beta2,
              fwd = (float)wd, feps = (float)eps;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = fmaf(fwd, p[i], g[i]);
    float mi = fmaf(w1, gi - m[i], m[i]);
    float vi = fmaf(w2 * gi, gi, v[i] * b2);
    m[i] = mi;
    v[i] = vi;
    p[i] = fmaf(neg_step, mi / (sqrtf(vi) / bc2s + feps), p[i]);
  }
}


// Multi-tensor Adam: one launch updates up to kAdamMulti parameter tensors (the ~100 small weights / biases / BatchNorm
// affine vectors of the PFCN / FairGo MLPs), each with its own step count (torch keeps `state['step']` per parameter and
// skips parameters whose .grad is None).  CTA b serves 1024-element tile (b - first_tile[e]) of entry e.
constexpr int kAdamMulti = 48;
struct AdamMultiArgs {
  float *p[kAdamMulti];
  const float *g[kAdamMulti];
  float *m[kAdamMulti], *v[kAdamMulti];
  int n[kAdamMulti], step[kAdamMulti], first_tile[kAdamMulti + 1];
  int *step_dev[kAdamMulti];     // optional device-resident step counters (bumped by k_adam_bump right before)
  int n_entries;
  double lr, beta1, beta2, eps, wd;
};
__global__ void k_adam_bump(AdamMultiArgs a) {
  const int e = threadIdx.x;
  if (e < a.n_entries && a.step_dev[e]) *a.step_dev[e] += 1;
}
__global__ void k_bump_u64(unsigned long long *c, unsigned long long inc) { *c += inc; }

__global__ void __launch_bounds__(256) k_adam_multi(AdamMultiArgs a) {
  __shared__ float sc[2];
  __shared__ int se;
  if (threadIdx.x == 0) {
    int e = 0;
    while (e + 1 < a.n_entries && (int)blockIdx.x >= a.first_tile[e + 1]) ++e;
    se = e;
    const double t = a.step_dev[e] ? (double)*a.step_dev[e] : (double)a.step[e];
    sc[0] = (float)(-a.lr / (1.0 - pow(a.beta1, t)));
    sc[1] = (float)sqrt(1.0 - pow(a.beta2, t));
  }
  __syncthreads();
  const int e = se;
  const float neg_step = sc[0], bc2s = sc[1], w1 = (float)(1.0 - a.beta1), w2 = (float)(1.0 - a.beta2),
              b2 = (float)a.beta2, fwd = (float)a.wd, feps = (float)a.eps;
  float *p = a.p[e], *m = a.m[e], *v = a.v[e];
  const float *g = a.g[e];
  const int base = ((int)blockIdx.x - a.first_tile[e]) * 1024, hi = min(a.n[e], base + 1024);
  for (int i = base + threadIdx.x; i < hi; i += 256) {
    float gi = fmaf(fwd, p[i], g[i]);
    float mi = fmaf(w1, gi - m[i], m[i]);
    float vi = fmaf(w2 * gi, gi, v[i] * b2);
    m[i] = mi;
    v[i] = vi;
    p[i] = fmaf(neg_step, mi / (sqrtf(vi) / bc2s + feps), p[i]);
  }
}


// ---------------------------------------------------------------- generic layer ops (PFCN / FairGo building blocks)
__global__ void k_act_bwd(const float *__restrict__ dY, const float *__restrict__ Y, int act, int64_t n,
                          float *__restrict__ dpre) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dpre[i] = dY[i] * act_bwd(Y[i], act);
}

// BatchNorm1d over the batch dimension (nn.BatchNorm1d, layers.py:64-65), fused with the following activation.
// Two launches, both spread over (column groups of 32) x (row slices): the batch is a few thousand rows of <= 256
// features, so a column-per-CTA sweep would leave most SMs idle.
//   k_bn_fwd_stats : CTA (cg, sl) -> per column (count, mean, M2) of its row slice, two-pass inside the slice
//   k_bn_fwd_apply : every CTA re-combines the slice statistics of its 32 columns in slice order (Chan's parallel
//                    update: deterministic, no atomics), then normalises its own row slice:
//                    y = act((x - mean) * invstd * gamma + beta); slice 0 also writes save_mean / save_invstd and the
//                    running statistics (momentum update with the unbiased variance).
//   eval mode      : k_bn_fwd_apply alone, with the running statistics.
constexpr int kBnSlices = 32;
__device__ __forceinline__ int bn_slice_rows(int M) { return (M + kBnSlices - 1) / kBnSlices; }

__global__ void __launch_bounds__(256)
    k_bn_fwd_stats(const float *__restrict__ X, int M, int N, float *__restrict__ part /* [slices][N][2]: mean, M2 */) {
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const bool ok = c < N;
  const int rows = bn_slice_rows(M), r0 = blockIdx.y * rows, r1 = min(M, r0 + rows);
  const int cnt = max(r1 - r0, 0);
  float s = 0.f;
#pragma unroll 4
  for (int m = r0 + w; m < r1; m += 8) s += ok ? X[(size_t)m * N + c] : 0.f;
  red[w][lane] = s;
  __syncthreads();
  s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += red[i][lane];
  const float mean = cnt > 0 ? s / (float)cnt : 0.f;
  __syncthreads();
  float q = 0.f;
#pragma unroll 4
  for (int m = r0 + w; m < r1; m += 8) {
    const float dlt = ok ? X[(size_t)m * N + c] - mean : 0.f;
    q = fmaf(dlt, dlt, q);
  }
  red[w][lane] = q;
  __syncthreads();
  if (w == 0 && ok) {
    q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) q += red[i][lane];
    part[((size_t)blockIdx.y * N + c) * 2] = mean;
    part[((size_t)blockIdx.y * N + c) * 2 + 1] = q;
  }
}

__global__ void __launch_bounds__(256)
    k_bn_fwd_apply(const float *__restrict__ X, const float *__restrict__ gamma, const float *__restrict__ beta,
                   float *__restrict__ rmean, float *__restrict__ rvar, int M, int N, float momentum, float eps,
                   int training, int act, const float *__restrict__ part, float *__restrict__ Y,
                   float *__restrict__ save_mean, float *__restrict__ save_invstd) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const bool ok = c < N;
  const int rows = bn_slice_rows(M), r0 = blockIdx.y * rows, r1 = min(M, r0 + rows);
  float mean = 0.f, invstd = 0.f;
  if (ok) {
    if (training) {
      // all slice partials are fetched first (32 independent loads in flight), then combined in slice order with Chan
      // et al.'s pairwise update; empty slices hold (0, 0) and are skipped by the predicate, not by a branch on loads
      float2 pr[kBnSlices];
#pragma unroll
      for (int sl = 0; sl < kBnSlices; ++sl) pr[sl] = __ldg((const float2 *)(part + ((size_t)sl * N + c) * 2));
      float n = 0.f, M2 = 0.f;
#pragma unroll
      for (int sl = 0; sl < kBnSlices; ++sl) {
        const int cb = min(M, (sl + 1) * rows) - sl * rows;
        if (cb > 0) {
          const float nb = (float)cb, nn = n + nb, dlt = pr[sl].x - mean;
          mean += dlt * (nb / nn);
          M2 += pr[sl].y + dlt * dlt * (n * nb / nn);
          n = nn;
        }
      }
      const float var = M2 / (float)M;
      invstd = 1.f / sqrtf(var + eps);
      if (blockIdx.y == 0 && w == 0) {
        save_mean[c] = mean;
        save_invstd[c] = invstd;
        rmean[c] = (1.f - momentum) * rmean[c] + momentum * mean;
        rvar[c] = (1.f - momentum) * rvar[c] + momentum * (M > 1 ? M2 / (float)(M - 1) : var);
      }
    } else {
      mean = rmean[c];
      invstd = 1.f / sqrtf(rvar[c] + eps);
    }
  }
  const float g = ok ? gamma[c] : 0.f, b = ok ? beta[c] : 0.f;
#pragma unroll 4
  for (int m = r0 + w; m < r1; m += 8)
    if (ok) Y[(size_t)m * N + c] = act_fwd(fmaf((X[(size_t)m * N + c] - mean) * invstd, g, b), act);
}

// backward (training mode): dpre = dY * act'(Y); dbeta = sum dpre; dgamma = sum dpre * xhat;
// dX = gamma * invstd / M * (M * dpre - dbeta - xhat * dgamma).  Same slice decomposition: slice partials of the two
// sums, then every CTA adds them in slice order and writes its rows of dX.
__global__ void __launch_bounds__(256)
    k_bn_bwd_stats(const float *__restrict__ X, const float *__restrict__ Y, const float *__restrict__ dY,
                   const float *__restrict__ save_mean, const float *__restrict__ save_invstd, int M, int N, int act,
                   float *__restrict__ part /* [slices][N][2]: sum dpre, sum dpre*xhat */) {
  // the two column sums cancel heavily under BPR (+g / -g per user): float64 accumulators inside the slice, float
  // slice partials (a 1/32 slice of the batch), float64 again for the cross-slice combination in k_bn_bwd_apply
  __shared__ double red[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const bool ok = c < N;
  const int rows = bn_slice_rows(M), r0 = blockIdx.y * rows, r1 = min(M, r0 + rows);
  const float mean = ok ? save_mean[c] : 0.f, invstd = ok ? save_invstd[c] : 0.f;
  double sb = 0.0, sg = 0.0;
#pragma unroll 4
  for (int m = r0 + w; m < r1; m += 8) {
    if (ok) {
      const size_t i = (size_t)m * N + c;
      const float dp = dY[i] * act_bwd(Y[i], act);
      sb += (double)dp;
      sg += (double)(dp * ((X[i] - mean) * invstd));
    }
  }
  red[w][lane] = sb;
  __syncthreads();
  sb = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) sb += red[i][lane];
  __syncthreads();
  red[w][lane] = sg;
  __syncthreads();
  if (w == 0 && ok) {
    sg = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) sg += red[i][lane];
    part[((size_t)blockIdx.y * N + c) * 2] = (float)sb;
    part[((size_t)blockIdx.y * N + c) * 2 + 1] = (float)sg;
  }
}

__global__ void __launch_bounds__(256)
    k_bn_bwd_apply(const float *__restrict__ X, const float *__restrict__ Y, const float *__restrict__ dY,
                   const float *__restrict__ gamma, const float *__restrict__ save_mean,
                   const float *__restrict__ save_invstd, int M, int N, int act, const float *__restrict__ part,
                   float *__restrict__ dX, float *__restrict__ dgamma, float *__restrict__ dbeta) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const bool ok = c < N;
  const int rows = bn_slice_rows(M), r0 = blockIdx.y * rows, r1 = min(M, r0 + rows);
  float sb = 0.f, sg = 0.f, mean = 0.f, invstd = 0.f, g = 0.f;
  if (ok) {
    float2 pr[kBnSlices];
#pragma unroll
    for (int sl = 0; sl < kBnSlices; ++sl) pr[sl] = __ldg((const float2 *)(part + ((size_t)sl * N + c) * 2));
    double sbd = 0.0, sgd = 0.0;
#pragma unroll
    for (int sl = 0; sl < kBnSlices; ++sl) {
      if (sl * rows < M) {
        sbd += (double)pr[sl].x;
        sgd += (double)pr[sl].y;
      }
    }
    sb = (float)sbd;
    sg = (float)sgd;
    mean = save_mean[c]; invstd = save_invstd[c]; g = gamma[c];
    if (blockIdx.y == 0 && w == 0) {
      dbeta[c] = sb;
      dgamma[c] = sg;
    }
  }
  const float k = g * invstd / (float)M;
#pragma unroll 4
  for (int m = r0 + w; m < r1; m += 8) {
    if (ok) {
      const size_t i = (size_t)m * N + c;
      const float dp = dY[i] * act_bwd(Y[i], act), xh = (X[i] - mean) * invstd;
      dX[i] = k * ((float)M * dp - sb - xh * sg);
    }
  }
}

// out[m, col0 : col0 + d] = T[idx[m], :]  (embedding lookup written straight into a slice of a wider row: concat for free)
__global__ void __launch_bounds__(256)
    k_gather_rows(const float *__restrict__ T, const int32_t *__restrict__ idx, int64_t M, int d, float *__restrict__ out,
                  int ld_out, int col0) {
  const int dq = d >> 2;
  const int64_t nq = M * dq;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = q / dq;
    const int c = (int)(q % dq);
    const float4 v = __ldg((const float4 *)(T + (size_t)idx[r] * d) + c);
    float *o = out + r * ld_out + col0 + c * 4;
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
}

__global__ void __launch_bounds__(256)
    k_gather_rows_any(const float *__restrict__ T, const int32_t *__restrict__ idx, int64_t M, int d, float *__restrict__ out,
                      int ld_out, int col0) {
  const int64_t n = M * d;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = q / d;
    const int c = (int)(q % d);
    out[r * ld_out + col0 + c] = __ldg(T + (size_t)idx[r] * d + c);
  }
}

// single-CTA scalar losses with a fixed-order reduction (batches are a few thousand rows)
__device__ __forceinline__ float block_total_1024(float v) {
  __shared__ float sh[33];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = threadIdx.x < 32 ? sh[threadIdx.x] : 0.f;
  if (threadIdx.x < 32) {
    t = warp_sum(t);
    if (threadIdx.x == 0) sh[32] = t;
  }
  __syncthreads();
  t = sh[32];
  __syncthreads();
  return t;
}

// loss.py:44-46 BPRLoss: -mean(log(1e-10 + sigmoid(pos - neg)))
__global__ void __launch_bounds__(1024)
    k_bpr(const float *__restrict__ pos, const float *__restrict__ neg, int M, float *__restrict__ loss,
          float *__restrict__ dpos, float *__restrict__ dneg) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < M; i += 1024) {
    const float sg = 1.f / (1.f + expf(-(pos[i] - neg[i])));
    acc += -logf(1e-10f + sg);
    const float g = -(sg * (1.f - sg)) / (1e-10f + sg) / (float)M;
    dpos[i] = g;
    dneg[i] = -g;
  }
  acc = block_total_1024(acc);
  if (threadIdx.x == 0) loss[0] = acc / (float)M;
}

// nn.BCELoss(sigmoid(z), y) (pfcn_mlp.py:206-207): logs clamped at -100; dz through the sigmoid
__global__ void __launch_bounds__(1024)
    k_sigmoid_bce_loss(const float *__restrict__ z, const float *__restrict__ y, int M, float *__restrict__ loss,
                       float *__restrict__ dz) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < M; i += 1024) {
    const float p = 1.f / (1.f + expf(-z[i])), t = y[i];
    acc += -(t * fmaxf(logf(p), -100.f) + (1.f - t) * fmaxf(logf(1.f - p), -100.f));
    const float den = fmaxf(p * (1.f - p), 1e-12f);
    dz[i] = (p - t) / den * (p * (1.f - p)) / (float)M;
  }
  acc = block_total_1024(acc);
  if (threadIdx.x == 0) loss[0] = acc / (float)M;
}

// nn.CrossEntropyLoss(Z, y) (pfcn_mlp.py:209): mean over rows of -log softmax(Z)[y]
__global__ void __launch_bounds__(1024)
    k_softmax_ce(const float *__restrict__ Z, const int32_t *__restrict__ y, int M, int C, float *__restrict__ loss,
                 float *__restrict__ dZ) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < M; i += 1024) {
    const float *zr = Z + (size_t)i * C;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, zr[c]);
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(zr[c] - mx);
    const float lse = mx + logf(se);
    const int t = y[i];
    acc += lse - zr[t];
    for (int c = 0; c < C; ++c) dZ[(size_t)i * C + c] = (expf(zr[c] - lse) - (c == t ? 1.f : 0.f)) / (float)M;
  }
  acc = block_total_1024(acc);
  if (threadIdx.x == 0) loss[0] = acc / (float)M;
}

// linear_tc.cu
bool tc_linear_eligible(int64_t M, int K, int N);
int tc_linear_forward(const float *X, const float *W, const float *b, float *Y, int64_t M, int K, int N, int act,
                      float drop_p, unsigned long long seed, const unsigned long long *seed_dev, int layer, cudaStream_t st);

static void launch_gemm(bool tb, const GemmArgs &g, cudaStream_t st) {
  dim3 grid((g.N + 63) / 64, (g.M + 63) / 64);
  if (tb) {
    FR_LAUNCH(k_gemm<true>, grid, 256, 0, st, g);
  } else {
    FR_LAUNCH(k_gemm<false>, grid, 256, 0, st, g);
  }
}

struct MlpWs {
  float *X;          // [M, dims[0]] gathered input
  float *act[8];     // post-activation outputs of every layer [M, dims[l+1]]
  float *dA, *dB;    // ping-pong row gradients [M, max width]
  float *part_w;
  double *part_b;
  float *p, *dp, *dz, *bce_part, *seg_eps, *coef, *loss_tmp;
  int32_t *pos_idx, *n_pos, *seg_id, *seg_off, *n_seg;
  uint32_t *keys, *skey, *ord, *gkeys, *gskey, *gord;
  int32_t *gseg_id, *gseg_off, *n_groups, *grp;
  uint32_t *ukey, *uord, *ikey, *iord;
  int32_t *useg_id, *useg_off, *un_seg, *iseg_id, *iseg_off, *in_seg;
  SortScratch sort;
  SegScratch seg;
};

static MlpWs carve_mlp(Carver &c, const fr_mlp_tower *t, int64_t M) {
  MlpWs w;
  int maxw = t->dims[0];
  size_t maxwk = 0;
  for (int l = 0; l < t->n_layers; ++l) {
    maxw = max(maxw, t->dims[l + 1]);
    maxwk = max(maxwk, (size_t)t->dims[l] * t->dims[l + 1]);
  }
  const size_t m = (size_t)(M < 1 ? 1 : M), chunks = (m + kWgradChunk - 1) / kWgradChunk;
  w.X = c.take<float>(m * t->dims[0]);
  for (int l = 0; l < 8; ++l) w.act[l] = l < t->n_layers ? c.take<float>(m * t->dims[l + 1]) : nullptr;
  w.dA = c.take<float>(m * maxw);
  w.dB = c.take<float>(m * maxw);
  w.part_w = c.take<float>(chunks * maxwk);
  w.part_b = c.take<double>(chunks * maxw);
  w.p = c.take<float>(m);
  w.dp = c.take<float>(m);
  w.dz = c.take<float>(m);
  w.bce_part = c.take<float>(m / 256 + 2);
  w.seg_eps = c.take<float>(m);
  w.coef = c.take<float>(32 * m);   // [J, G], G <= kDfMaxG
  w.loss_tmp = c.take<float>(4);
  w.pos_idx = c.take<int32_t>(m);
  w.n_pos = c.take<int32_t>(1);
  w.seg_id = c.take<int32_t>(m);
  w.seg_off = c.take<int32_t>(m + 1);
  w.n_seg = c.take<int32_t>(1);
  w.keys = c.take<uint32_t>(m);
  w.skey = c.take<uint32_t>(m);
  w.ord = c.take<uint32_t>(m);
  w.gkeys = c.take<uint32_t>(m);
  w.gskey = c.take<uint32_t>(m);
  w.gord = c.take<uint32_t>(m);
  w.gseg_id = c.take<int32_t>(m);
  w.gseg_off = c.take<int32_t>(m + 1);
  w.n_groups = c.take<int32_t>(1);
  w.grp = c.take<int32_t>(m);
  w.ukey = c.take<uint32_t>(m);
  w.uord = c.take<uint32_t>(m);
  w.ikey = c.take<uint32_t>(m);
  w.iord = c.take<uint32_t>(m);
  w.useg_id = c.take<int32_t>(m);
  w.useg_off = c.take<int32_t>(m + 1);
  w.un_seg = c.take<int32_t>(1);
  w.iseg_id = c.take<int32_t>(m);
  w.iseg_off = c.take<int32_t>(m + 1);
  w.in_seg = c.take<int32_t>(1);
  w.sort = carve_sort_scratch(c, m);
  w.seg = carve_seg_scratch(c, m);
  return w;
}

static int check_tower(const fr_mlp_tower *t, const char *who) {
  FR_REQUIRE(t && t->n_layers >= 1 && t->n_layers <= 8, "%s: tower needs 1..8 layers", who);
  for (int l = 0; l < t->n_layers; ++l) {
    FR_REQUIRE(t->W[l] && t->b[l] && t->dims[l] >= 1 && t->dims[l + 1] >= 1, "%s: layer %d incomplete", who, l);
  }
  return FR_OK;
}

}  // namespace fr

extern "C" {

size_t fr_nfcf_workspace_bytes(const fr_mlp_tower *t, int64_t M) {
  if (!t) return 0;
  fr::Carver c(nullptr, 0);
  fr::carve_mlp(c, t, M);
  return c.off;
}

// forward: p = sigmoid(tower(U[uid] || I[iid])), loss = BCE(p, label) [+ fair_weight * DF over the positives]
int fr_nfcf_forward(const fr_nfcf_step *s, void *stream) {
  FR_REQUIRE(s && s->U && s->I && s->uid && s->iid && s->label && s->loss && s->pred && s->workspace && s->status_flags,
             "fr_nfcf_forward: null pointer");
  int rc = fr::check_tower(&s->tower, "fr_nfcf_forward");
  if (rc) return rc;
  const fr_mlp_tower *t = &s->tower;
  FR_REQUIRE(s->d % 4 == 0 && t->dims[0] == 2 * s->d && t->dims[t->n_layers] == 1 && s->M >= 1,
             "fr_nfcf_forward: tower must map 2*d -> 1");
  FR_REQUIRE(!s->use_df || s->sst, "fr_nfcf_forward: sst missing");
  fr::Carver c(s->workspace, s->workspace_bytes);
  fr::MlpWs w = fr::carve_mlp(c, t, s->M);
  if (!c.ok()) {
    fr::set_error("fr_nfcf_forward: workspace too small (%zu < %zu)", s->workspace_bytes, c.off);
    return FR_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int M = (int)s->M;
  FR_LAUNCH(fr::k_gather_concat, fr::grid_for((int64_t)M * s->d / 2, 256), 256, 0, st, s->U, s->I, s->uid, s->iid, s->M,
            s->d, w.X);
  const float *in = w.X;
  for (int l = 0; l < t->n_layers; ++l) {
    fr::GemmArgs g{in, t->W[l], t->b[l], w.act[l], M, t->dims[l + 1], t->dims[l], t->dims[l], t->dims[l], t->dims[l + 1],
                   t->act, nullptr, 0, s->training ? t->dropout : 0.f, s->seed, l, 0};
    fr::launch_gemm(true, g, st);
    in = w.act[l];
  }
  const int nblk = (M + 255) / 256;
  FR_LAUNCH(fr::k_sigmoid_bce, nblk, 256, 0, st, w.act[t->n_layers - 1], s->label, M, w.p, w.dp, w.bce_part);
  FR_CUDA_OK(cudaMemcpyAsync(s->pred, w.p, sizeof(float) * M, cudaMemcpyDeviceToDevice, st));
  if (s->use_df) {
    FR_LAUNCH(fr::k_compact_pos, 1, 1024, 0, st, s->label, M, w.pos_idx, w.n_pos);
    FR_LAUNCH(fr::k_pos_keys, fr::grid_for(M, 256), 256, 0, st, s->iid, w.pos_idx, w.n_pos, w.keys);
    fr::sort_pairs(w.keys, nullptr, w.skey, w.ord, M, w.n_pos, fr::bits_for((uint32_t)s->n_items), w.sort, st);
    fr::build_segments(w.skey, w.ord, M, w.n_pos, w.seg_id, w.seg_off, w.n_seg, nullptr, nullptr, nullptr, w.seg, st);
    // groups = torch.unique(sst[pos], return_inverse=True): sort the order-encoded values, segment, scatter the ranks back
    FR_LAUNCH(fr::k_df_group_keys, fr::grid_for(M, 256), 256, 0, st, s->sst, w.pos_idx, w.n_pos, w.gkeys);
    fr::sort_pairs(w.gkeys, nullptr, w.gskey, w.gord, M, w.n_pos, 32, w.sort, st);
    fr::build_segments(w.gskey, w.gord, M, w.n_pos, w.gseg_id, w.gseg_off, w.n_groups, nullptr, nullptr, w.grp, w.seg, st);
    fr::DfArgs da{w.p, s->sst, w.pos_idx, w.n_pos, w.ord, w.seg_off, w.n_seg, s->fair_weight, w.seg_eps, w.coef, w.grp,
                  w.n_groups, s->status_flags};
    FR_LAUNCH(fr::k_df_segments, fr::grid_for(M, 8 * 8, fr::kSMs * 2), 256, 0, st, da);
    FR_LAUNCH(fr::k_df_scatter_coef, fr::grid_for(M, 256), 256, 0, st, da, w.seg_id, w.dp);
  }
  FR_LAUNCH(fr::k_nfcf_finish, 1, 256, 0, st, w.bce_part, nblk, M, w.seg_eps, w.n_seg, s->use_df, s->fair_weight, s->loss);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

// backward of fr_nfcf_forward (same workspace): dense dU (optional), dI, dW[l], db[l]; grad_scale = upstream dL
int fr_nfcf_backward(const fr_nfcf_step *s, float grad_scale, void *stream) {
  FR_REQUIRE(s && s->workspace && s->dI, "fr_nfcf_backward: null pointer");
  const fr_mlp_tower *t = &s->tower;
  fr::Carver c(s->workspace, s->workspace_bytes);
  fr::MlpWs w = fr::carve_mlp(c, t, s->M);
  cudaStream_t st = (cudaStream_t)stream;
  const int M = (int)s->M, L = t->n_layers;
  FR_LAUNCH(fr::k_sigmoid_bwd, (M + 255) / 256, 256, 0, st, w.dp, w.p, w.act[L - 1], M, grad_scale, w.dz);
  const int wchunk = fr::wgrad_chunk(M), chunks = (M + wchunk - 1) / wchunk;
  float *dcur = w.dz;     // gradient w.r.t. the PRE-activation of layer l: [M, dims[l+1]]
  float *bufs[2] = {w.dA, w.dB};
  for (int l = L - 1; l >= 0; --l) {
    const float *inp = l > 0 ? w.act[l - 1] : w.X;
    const int N = t->dims[l + 1], K = t->dims[l];
    fr::WgradArgs wa{dcur, inp, w.part_w, w.part_b, M, N, K, N, K, s->training ? t->dropout : 0.f, s->seed, l,
                     nullptr, 0, nullptr, nullptr, nullptr, nullptr, wchunk};
    dim3 grid((K + 63) / 64, (N + 63) / 64, chunks);
    FR_LAUNCH(fr::k_wgrad, grid, 256, 0, st, wa);
    FR_LAUNCH(fr::k_sum_chunks<float>, fr::grid_for((int64_t)N * K, 256), 256, 0, st, (const float *)w.part_w, chunks,
              (int64_t)N * K, s->dW[l]);
    FR_LAUNCH(fr::k_sum_chunks<double>, 1, 256, 0, st, (const double *)w.part_b, chunks, (int64_t)N, s->db[l]);
    // dInput[M,K] = dcur[M,N] . W[N,K], then through the dropout mask of this layer's input and the previous
    // layer's activation
    float *dnext = bufs[l & 1];
    fr::GemmArgs g{dcur, t->W[l], nullptr, dnext, M, K, N, N, K, K, fr::ACT_NONE, l > 0 ? w.act[l - 1] : nullptr, t->act,
                   s->training ? t->dropout : 0.f, s->seed, l, 1};
    fr::launch_gemm(false, g, st);
    dcur = dnext;
  }
  // dcur = dX [M, 2d] -> dense embedding gradients (stable sort by id, ordered segment sums)
  const int d = s->d;
  if (s->dU) {
    FR_CUDA_OK(cudaMemsetAsync(s->dU, 0, sizeof(float) * (size_t)s->n_users * d, st));
    fr::sort_pairs((const uint32_t *)s->uid, nullptr, w.ukey, w.uord, M, nullptr, fr::bits_for((uint32_t)s->n_users),
                   w.sort, st);
    fr::build_segments(w.ukey, w.uord, M, nullptr, w.useg_id, w.useg_off, w.un_seg, nullptr, nullptr, nullptr, w.seg, st);
    FR_LAUNCH(fr::k_segment_sum_rows, fr::grid_for(M, 8, fr::kSMs * 8), 256, 0, st, w.ukey, w.uord, w.useg_off, w.un_seg,
              (const float *)dcur, 2 * d, 0, d, s->dU);
  }
  FR_CUDA_OK(cudaMemsetAsync(s->dI, 0, sizeof(float) * (size_t)s->n_items * d, st));
  fr::sort_pairs((const uint32_t *)s->iid, nullptr, w.ikey, w.iord, M, nullptr, fr::bits_for((uint32_t)s->n_items), w.sort,
                 st);
  fr::build_segments(w.ikey, w.iord, M, nullptr, w.iseg_id, w.iseg_off, w.in_seg, nullptr, nullptr, nullptr, w.seg, st);
  FR_LAUNCH(fr::k_segment_sum_rows, fr::grid_for(M, 8, fr::kSMs * 8), 256, 0, st, w.ikey, w.iord, w.iseg_off, w.in_seg,
            (const float *)dcur, 2 * d, d, d, s->dI);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_adam_dense(float *p, const float *g, float *m, float *v, int64_t n, int32_t step, double lr, double beta1,
                  double beta2, double eps, double weight_decay, void *stream) {
  FR_REQUIRE(p && g && m && v && n >= 1 && step >= 1, "fr_adam_dense: bad argument");
  FR_LAUNCH(fr::k_adam_flat, fr::grid_for(n, 256, fr::kSMs * 16), 256, 0, stream, p, g, m, v, n, step, lr, beta1, beta2,
            eps, weight_decay);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_adam_multi(const fr_adam_entry *entries_host, int32_t n_entries, double lr, double beta1, double beta2, double eps,
                  double weight_decay, void *stream) {
  FR_REQUIRE(entries_host && n_entries >= 1, "fr_adam_multi: bad argument");
  for (int32_t e0 = 0; e0 < n_entries; e0 += fr::kAdamMulti) {
    fr::AdamMultiArgs a;
    const int cnt = n_entries - e0 < fr::kAdamMulti ? n_entries - e0 : fr::kAdamMulti;
    int tiles = 0;
    bool any_dev = false;
    for (int e = 0; e < cnt; ++e) {
      const fr_adam_entry &x = entries_host[e0 + e];
      FR_REQUIRE(x.p && x.g && x.m && x.v && x.n >= 1 && x.n < (int64_t)INT32_MAX && (x.step >= 1 || x.step_dev),
                 "fr_adam_multi: bad entry");
      a.p[e] = x.p; a.g[e] = x.g; a.m[e] = x.m; a.v[e] = x.v;
      a.n[e] = (int)x.n; a.step[e] = x.step; a.step_dev[e] = x.step_dev; a.first_tile[e] = tiles;
      any_dev |= x.step_dev != nullptr;
      tiles += (int)((x.n + 1023) / 1024);
    }
    a.first_tile[cnt] = tiles;
    a.n_entries = cnt;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = weight_decay;
    if (any_dev) FR_LAUNCH(fr::k_adam_bump, 1, 64, 0, stream, a);
    FR_LAUNCH(fr::k_adam_multi, tiles, 256, 0, stream, a);
  }
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_bump_u64(uint64_t *counter_dev, uint64_t inc, void *stream) {
  FR_REQUIRE(counter_dev, "fr_bump_u64: null pointer");
  FR_LAUNCH(fr::k_bump_u64, 1, 1, 0, stream, (unsigned long long *)counter_dev, (unsigned long long)inc);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

// ---------------------------------------------------------------- generic layer ops
int fr_linear_uses_tensor_cores(int64_t M, int32_t K, int32_t N) { return fr::tc_linear_eligible(M, K, N) ? 1 : 0; }

int fr_linear_forward(const float *X, const float *W, const float *b, float *Y, int64_t M, int32_t K, int32_t N, int32_t act,
                      float drop_p, uint64_t seed, const uint64_t *seed_dev, int32_t layer, int32_t allow_tensor_cores,
                      void *stream) {
  FR_REQUIRE(X && W && Y && M >= 1 && K >= 1 && N >= 1, "fr_linear_forward: bad argument");
  if (allow_tensor_cores && fr::tc_linear_eligible(M, K, N)) {
    // tensor-core path: tcgen05 3xTF32 (linear_tc.cu)
    int rc = fr::tc_linear_forward(X, W, b, Y, M, K, N, act, drop_p, seed, (const unsigned long long *)seed_dev, layer,
                                   (cudaStream_t)stream);
    if (rc) return rc;
    FR_LAUNCH_CHECK();
    return FR_OK;
  }
  fr::GemmArgs g{X, W, b, Y, (int)M, N, K, K, K, N, act, nullptr, 0, drop_p, seed, layer, 0, nullptr, 0,
                 (const unsigned long long *)seed_dev};
  fr::launch_gemm(true, g, (cudaStream_t)stream);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

size_t fr_linear_backward_workspace_bytes(int64_t M, int32_t K, int32_t N) {
  const size_t chunks = ((size_t)M + fr::kWgradChunk - 1) / fr::kWgradChunk;
  return ((size_t)M * N + chunks * (size_t)N * K + 2 * chunks * (size_t)N) * 4 + 1024;
}

int fr_linear_backward(const float *X, const float *W, const float *Y, const float *dY, int64_t M, int32_t K, int32_t N,
                       int32_t act, float drop_p, uint64_t seed, const uint64_t *seed_dev_, int32_t layer, float *dX, float *dW,
                       float *db, int32_t *tickets, void *workspace, size_t workspace_bytes, void *stream) {
  const unsigned long long *seed_dev = (const unsigned long long *)seed_dev_;
  FR_REQUIRE(X && W && Y && dY && dW && workspace && M >= 1, "fr_linear_backward: bad argument");
  if (workspace_bytes < fr_linear_backward_workspace_bytes(M, K, N)) {
    fr::set_error("fr_linear_backward: workspace too small");
    return FR_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int wchunk = fr::wgrad_chunk(M), chunks = (int)((M + wchunk - 1) / wchunk);
  fr::Carver c(workspace, workspace_bytes);
  float *dpre = c.take<float>((size_t)M * N);
  float *part_w = c.take<float>((size_t)chunks * N * K);
  double *part_b = c.take<double>((size_t)chunks * N);
  (void)dpre;
  dim3 grid((K + 63) / 64, (N + 63) / 64, chunks);
  const bool fused = tickets != nullptr && grid.x * grid.y <= 1024;
  fr::WgradArgs wa{dY, X, part_w, part_b, (int)M, N, K, N, K, drop_p, seed, layer, act ? Y : nullptr, act, seed_dev,
                   fused ? tickets : nullptr, dW, db, wchunk};
  FR_LAUNCH(fr::k_wgrad, grid, 256, 0, st, wa);
  if (!fused) {
    FR_LAUNCH(fr::k_sum_chunks<float>, fr::grid_for((int64_t)N * K, 256), 256, 0, st, (const float *)part_w, chunks,
              (int64_t)N * K, dW);
    if (db) FR_LAUNCH(fr::k_sum_chunks<double>, 1, 256, 0, st, (const double *)part_b, chunks, (int64_t)N, db);
  }
  if (dX) {
    fr::GemmArgs g{dY, W, nullptr, dX, (int)M, K, N, N, K, K, fr::ACT_NONE, nullptr, 0, drop_p, seed, layer, 1,
                   act ? Y : nullptr, act, seed_dev};
    fr::launch_gemm(false, g, st);
  }
  FR_LAUNCH_CHECK();
  return FR_OK;
}

size_t fr_batchnorm_workspace_bytes(int32_t N) { return (size_t)fr::kBnSlices * (size_t)N * 2 * sizeof(float) + 256; }

int fr_batchnorm_forward(const float *X, const float *gamma, const float *beta, float *running_mean, float *running_var,
                         int64_t M, int32_t N, float momentum, float eps, int32_t training, int32_t act, float *Y,
                         float *save_mean, float *save_invstd, void *workspace, size_t workspace_bytes, void *stream) {
  FR_REQUIRE(X && gamma && beta && running_mean && running_var && Y && save_mean && save_invstd && M >= 1 && N >= 1,
             "fr_batchnorm_forward: bad argument");
  FR_REQUIRE(!training || (workspace && workspace_bytes >= fr_batchnorm_workspace_bytes(N)),
             "fr_batchnorm_forward: workspace too small");
  dim3 grid((N + 31) / 32, fr::kBnSlices);
  if (training) FR_LAUNCH(fr::k_bn_fwd_stats, grid, 256, 0, stream, X, (int)M, N, (float *)workspace);
  FR_LAUNCH(fr::k_bn_fwd_apply, grid, 256, 0, stream, X, gamma, beta, running_mean, running_var, (int)M, N, momentum, eps,
            training, act, (const float *)workspace, Y, save_mean, save_invstd);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_batchnorm_backward(const float *X, const float *Y, const float *dY, const float *gamma, const float *save_mean,
                          const float *save_invstd, int64_t M, int32_t N, int32_t act, float *dX, float *dgamma,
                          float *dbeta, void *workspace, size_t workspace_bytes, void *stream) {
  FR_REQUIRE(X && Y && dY && gamma && save_mean && save_invstd && dX && dgamma && dbeta && M >= 1,
             "fr_batchnorm_backward: bad argument");
  FR_REQUIRE(workspace && workspace_bytes >= fr_batchnorm_workspace_bytes(N), "fr_batchnorm_backward: workspace too small");
  dim3 grid((N + 31) / 32, fr::kBnSlices);
  FR_LAUNCH(fr::k_bn_bwd_stats, grid, 256, 0, stream, X, Y, dY, save_mean, save_invstd, (int)M, N, act, (float *)workspace);
  FR_LAUNCH(fr::k_bn_bwd_apply, grid, 256, 0, stream, X, Y, dY, gamma, save_mean, save_invstd, (int)M, N, act,
            (const float *)workspace, dX, dgamma, dbeta);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_gather_rows(const float *T, const int32_t *idx, int64_t M, int32_t d, float *out, int32_t ld_out, int32_t col0,
                   void *stream) {
  FR_REQUIRE(T && idx && out && M >= 1 && d >= 1, "fr_gather_rows: bad argument");
  if (d % 4 == 0) {
    FR_LAUNCH(fr::k_gather_rows, fr::grid_for(M * d / 4, 256), 256, 0, stream, T, idx, M, d, out, ld_out, col0);
  } else {   // bias tables ([n,1], pfcn_biasedmf.py:48-51) and other odd widths
    FR_LAUNCH(fr::k_gather_rows_any, fr::grid_for(M * d, 256), 256, 0, stream, T, idx, M, d, out, ld_out, col0);
  }
  FR_LAUNCH_CHECK();
  return FR_OK;
}

size_t fr_scatter_rows_workspace_bytes(int64_t M) {
  fr::Carver c(nullptr, 0);
  const size_t m = (size_t)(M < 1 ? 1 : M);
  c.take<uint32_t>(m); c.take<uint32_t>(m); c.take<int32_t>(m); c.take<int32_t>(m + 1); c.take<int32_t>(1);
  fr::carve_sort_scratch(c, m);
  fr::carve_seg_scratch(c, m);
  return c.off;
}

// dense gradient of an embedding lookup: dT[idx[m], :] += dX[m, col0 : col0 + d], summed per row in batch order
int fr_scatter_rows_dense(const int32_t *idx, const float *dX, int32_t ldx, int32_t col0, int64_t M, int32_t d,
                          int32_t n_rows, float *dT, void *workspace, size_t workspace_bytes, void *stream) {
  FR_REQUIRE(idx && dX && dT && workspace && M >= 1, "fr_scatter_rows_dense: bad argument");
  fr::Carver c(workspace, workspace_bytes);
  const size_t m = (size_t)M;
  uint32_t *skey = c.take<uint32_t>(m), *ord = c.take<uint32_t>(m);
  int32_t *seg_id = c.take<int32_t>(m), *seg_off = c.take<int32_t>(m + 1), *n_seg = c.take<int32_t>(1);
  fr::SortScratch ss = fr::carve_sort_scratch(c, m);
  fr::SegScratch sg = fr::carve_seg_scratch(c, m);
  if (!c.ok()) {
    fr::set_error("fr_scatter_rows_dense: workspace too small");
    return FR_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  FR_CUDA_OK(cudaMemsetAsync(dT, 0, sizeof(float) * (size_t)n_rows * d, st));
  fr::sort_pairs((const uint32_t *)idx, nullptr, skey, ord, M, nullptr, fr::bits_for((uint32_t)n_rows), ss, st);
  fr::build_segments(skey, ord, M, nullptr, seg_id, seg_off, n_seg, nullptr, nullptr, nullptr, sg, st);
  FR_LAUNCH(fr::k_segment_sum_rows, fr::grid_for(M, 8, fr::kSMs * 8), 256, 0, st, skey, ord, seg_off, n_seg, dX, ldx,
            col0, d, dT);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_bpr_loss(const float *pos, const float *neg, int64_t M, float *loss, float *dpos, float *dneg, void *stream) {
  FR_REQUIRE(pos && neg && loss && dpos && dneg && M >= 1, "fr_bpr_loss: bad argument");
  FR_LAUNCH(fr::k_bpr, 1, 1024, 0, stream, pos, neg, (int)M, loss, dpos, dneg);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_sigmoid_bce_loss(const float *z, const float *y, int64_t M, float *loss, float *dz, void *stream) {
  FR_REQUIRE(z && y && loss && dz && M >= 1, "fr_sigmoid_bce_loss: bad argument");
  FR_LAUNCH(fr::k_sigmoid_bce_loss, 1, 1024, 0, stream, z, y, (int)M, loss, dz);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_softmax_ce_loss(const float *Z, const int32_t *y, int64_t M, int32_t C, float *loss, float *dZ, void *stream) {
  FR_REQUIRE(Z && y && loss && dZ && M >= 1 && C >= 2, "fr_softmax_ce_loss: bad argument");
  FR_LAUNCH(fr::k_softmax_ce, 1, 1024, 0, stream, Z, y, (int)M, C, loss, dZ);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

}  // extern "C"
