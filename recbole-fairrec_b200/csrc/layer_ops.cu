// Row-wise scorers and small glue ops of the PFCN / FairGo families on sm_100a.
//
// Reference being replaced (paths relative to the reference root):
//   recbole/model/fair_recommender/pfcn_pmf.py:172,183-184      torch.mul(u, i).sum(-1)               -> fr_rowdot_*
//   recbole/model/fair_recommender/pfcn_dmf.py:176,190-191      nn.CosineSimilarity()(u, i)           -> fr_cosine_*
//   recbole/model/fair_recommender/pfcn_biasedmf.py:189-192     [B] + [B,1] broadcast scores + BPRLoss -> fr_bpr_outer_loss
//   pfcn_*.py cm mode (`user_temp + embed`, `/ len(filter_layer)`), fairgo_*.py:178-180              -> fr_scaled_sum
//   torch.cat / torch.split along dim 1 (pfcn_mlp.py:172, fairgo_*.py:222)                            -> fr_copy_cols
//   nn.MSELoss (fairgo_pmf.py:170-171)                                                                -> fr_mse_loss
//   nn.Sigmoid / activation modules applied outside an MLPLayers, clamp(x,0,max)/max (fairgo_pmf.py:248) -> fr_act_*, fr_clamp_div
// All reductions have a fixed order (no floating-point atomics); all kernels are HBM/latency-bound row sweeps with
// 128-bit lane loads where a row is a whole number of float4.
#include "act.cuh"

namespace fr {

// ---------------------------------------------------------------- row dot
__global__ void __launch_bounds__(256)
    k_rowdot_fwd(const float *__restrict__ A, const float *__restrict__ B, int64_t M, int d, float *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int dq = d >> 2;
  for (int64_t m = warp; m < M; m += nw) {
    const float4 *a = (const float4 *)(A + m * d), *b = (const float4 *)(B + m * d);
    float s = 0.f;
    for (int c = lane; c < dq; c += 32) {
      const float4 x = a[c], y = b[c];
      s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
    }
    s = warp_sum(s);
    if (lane == 0) out[m] = s;
  }
}

__global__ void __launch_bounds__(256)
    k_rowdot_bwd(const float *__restrict__ A, const float *__restrict__ B, const float *__restrict__ dout, int64_t M, int d,
                 float *__restrict__ dA, float *__restrict__ dB) {
  const int dq = d >> 2;
  const int64_t nq = M * dq;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
    const float g = dout[q / dq];
    const float4 x = ((const float4 *)A)[q], y = ((const float4 *)B)[q];
    if (dA) ((float4 *)dA)[q] = make_float4(g * y.x, g * y.y, g * y.z, g * y.w);
    if (dB) ((float4 *)dB)[q] = make_float4(g * x.x, g * x.y, g * x.z, g * x.w);
  }
}

// ---------------------------------------------------------------- cosine similarity (dim = 1, eps = 1e-8)
// out[m] = sum_k (a_k / max(|a|, eps)) * (b_k / max(|b|, eps)) ; norms[m] = (max(|a|,eps), max(|b|,eps)) kept for backward
__global__ void __launch_bounds__(256)
    k_cosine_fwd(const float *__restrict__ A, const float *__restrict__ B, int64_t M, int d, float eps,
                 float *__restrict__ out, float2 *__restrict__ norms) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int dq = d >> 2;
  for (int64_t m = warp; m < M; m += nw) {
    const float4 *a = (const float4 *)(A + m * d), *b = (const float4 *)(B + m * d);
    float sa = 0.f, sb = 0.f;
    for (int c = lane; c < dq; c += 32) {
      const float4 x = a[c], y = b[c];
      sa = fmaf(x.x, x.x, sa); sa = fmaf(x.y, x.y, sa); sa = fmaf(x.z, x.z, sa); sa = fmaf(x.w, x.w, sa);
      sb = fmaf(y.x, y.x, sb); sb = fmaf(y.y, y.y, sb); sb = fmaf(y.z, y.z, sb); sb = fmaf(y.w, y.w, sb);
    }
    const float na = fmaxf(sqrtf(warp_sum(sa)), eps), nb = fmaxf(sqrtf(warp_sum(sb)), eps);
    float s = 0.f;
    for (int c = lane; c < dq; c += 32) {
      const float4 x = a[c], y = b[c];
      s = fmaf(x.x / na, y.x / nb, s); s = fmaf(x.y / na, y.y / nb, s);
      s = fmaf(x.z / na, y.z / nb, s); s = fmaf(x.w / na, y.w / nb, s);
    }
    s = warp_sum(s);
    if (lane == 0) {
      out[m] = s;
      norms[m] = make_float2(na, nb);
    }
  }
}

// d cos / d a_k = (b^_k - cos * a^_k) / |a|   (a^ = a/|a|; when |a| was clamped to eps the a^ term drops)
__global__ void __launch_bounds__(256)
    k_cosine_bwd(const float *__restrict__ A, const float *__restrict__ B, const float *__restrict__ out,
                 const float2 *__restrict__ norms, const float *__restrict__ dout, int64_t M, int d, float eps,
                 float *__restrict__ dA, float *__restrict__ dB) {
  const int64_t n = M * d;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = q / d;
    const float2 nn = norms[m];
    const float g = dout[m], c = out[m];
    const float ah = A[q] / nn.x, bh = B[q] / nn.y;
    if (dA) dA[q] = g * (bh - (nn.x > eps ? c * ah : 0.f)) / nn.x;
    if (dB) dB[q] = g * (ah - (nn.y > eps ? c * bh : 0.f)) / nn.y;
  }
}

// ---------------------------------------------------------------- BPR over the broadcast score matrix of PFCN_BiasedMF
// pfcn_biasedmf.py:189-190 adds a [B] vector of dots to [B,1] bias columns, so the scores are a [B,B] MATRIX:
//   pos[i,j] = ((dotp[j] + ub[i]) + pib[i]) + gb ,  neg[i,j] = ((dotn[j] + ub[i]) + nib[i]) + gb
//   loss = mean_{i,j} -log(1e-10 + sigmoid(pos - neg))               (loss.py:44-46)
// Row pass (CTA per i): row partial of the loss and d/d pib[i] = -d/d nib[i]; column pass (CTA per j) recomputes the
// elements to sum d/d dotp[j] = -d/d dotn[j] over i.  ub and gb cancel: their gradients are exactly zero.
__device__ __forceinline__ float block_sum_256(float v, float *sh) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0)
    for (int i = 0; i < 8; ++i) t += sh[i];
  __syncthreads();
  return t;   // valid on thread 0
}

__device__ __forceinline__ void bpr_elem(float dp, float dn, float ub, float pib, float nib, float gb, float &l, float &g) {
  const float pos = ((dp + ub) + pib) + gb, neg = ((dn + ub) + nib) + gb;
  const float sg = 1.f / (1.f + expf(-(pos - neg)));
  l = -logf(1e-10f + sg);
  g = -(sg * (1.f - sg)) / (1e-10f + sg);
}

__global__ void __launch_bounds__(256)
    k_bpr_outer_rows(const float *__restrict__ dotp, const float *__restrict__ dotn, const float *__restrict__ ub,
                     const float *__restrict__ pib, const float *__restrict__ nib, const float *__restrict__ gb, int B,
                     float *__restrict__ row_loss, float *__restrict__ d_pib, float *__restrict__ d_nib) {
  __shared__ float sh[8];
  const int i = blockIdx.x;
  const float u = ub[i], p = pib[i], n = nib[i], g0 = gb[0];
  float ls = 0.f, gs = 0.f;
  for (int j = threadIdx.x; j < B; j += 256) {
    float l, g;
    bpr_elem(dotp[j], dotn[j], u, p, n, g0, l, g);
    ls += l;
    gs += g;
  }
  ls = block_sum_256(ls, sh);
  gs = block_sum_256(gs, sh);
  if (threadIdx.x == 0) {
    const float inv = 1.f / ((float)B * (float)B);
    row_loss[i] = ls;
    d_pib[i] = gs * inv;
    d_nib[i] = -gs * inv;
  }
}

__global__ void __launch_bounds__(256)
    k_bpr_outer_cols(const float *__restrict__ dotp, const float *__restrict__ dotn, const float *__restrict__ ub,
                     const float *__restrict__ pib, const float *__restrict__ nib, const float *__restrict__ gb, int B,
                     const float *__restrict__ row_loss, float *__restrict__ d_dotp, float *__restrict__ d_dotn,
                     float *__restrict__ loss) {
  __shared__ float sh[8];
  const int j = blockIdx.x;
  const float dp = dotp[j], dn = dotn[j], g0 = gb[0];
  float gs = 0.f, ls = 0.f;
  for (int i = threadIdx.x; i < B; i += 256) {
    float l, g;
    bpr_elem(dp, dn, ub[i], pib[i], nib[i], g0, l, g);
    gs += g;
    if (j == 0) ls += row_loss[i];
  }
  gs = block_sum_256(gs, sh);
  if (j == 0) ls = block_sum_256(ls, sh);
  if (threadIdx.x == 0) {
    const float inv = 1.f / ((float)B * (float)B);
    d_dotp[j] = gs * inv;
    d_dotn[j] = -gs * inv;
    if (j == 0) loss[0] = ls * inv;
  }
}

// ---------------------------------------------------------------- scale * sum of up to 8 equally shaped tensors
struct SumArgs {
  const float *x[8];
  int n_terms;
  float scale;
};
__global__ void __launch_bounds__(256) k_scaled_sum(SumArgs a, int64_t n, float *__restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float s = a.x[0][i];
    for (int t = 1; t < a.n_terms; ++t) s += a.x[t][i];   // left-to-right like `temp + filter(x)` (fairgo_pmf.py:166)
    out[i] = a.scale == 1.f ? s : s * a.scale;
  }
}
// the reference divides (`/ len(filter_layer)`): keep a true division for bit-closeness
__global__ void __launch_bounds__(256) k_sum_div(SumArgs a, int64_t n, float *__restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float s = a.x[0][i];
    for (int t = 1; t < a.n_terms; ++t) s += a.x[t][i];
    out[i] = s / a.scale;
  }
}

// ---------------------------------------------------------------- strided column-block copy (torch.cat / split along dim 1)
__global__ void __launch_bounds__(256)
    k_copy_cols(const float *__restrict__ src, int ld_src, int col_src, float *__restrict__ dst, int ld_dst, int col_dst,
                int64_t M, int ncols) {
  const int64_t n = M * ncols;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = q / ncols;
    const int c = (int)(q % ncols);
    dst[m * ld_dst + col_dst + c] = src[m * ld_src + col_src + c];
  }
}

// ---------------------------------------------------------------- nn.MSELoss (mean), single CTA, fixed order
__global__ void __launch_bounds__(1024)
    k_mse(const float *__restrict__ pred, const float *__restrict__ target, int M, float *__restrict__ loss,
          float *__restrict__ dpred) {
  __shared__ float sh[33];
  float acc = 0.f;
  for (int i = threadIdx.x; i < M; i += 1024) {
    const float e = pred[i] - target[i];
    acc = fmaf(e, e, acc);
    dpred[i] = 2.f * e / (float)M;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = warp_sum(sh[threadIdx.x]);
    if (threadIdx.x == 0) loss[0] = t / (float)M;
  }
}

// ---------------------------------------------------------------- stand-alone activations
__global__ void __launch_bounds__(256) k_act_fwd(const float *__restrict__ x, int act, int64_t n, float *__restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = act_fwd(x[i], act);
}
__global__ void __launch_bounds__(256)
    k_act_bwd_out(const float *__restrict__ dY, const float *__restrict__ Y, int act, int64_t n, float *__restrict__ dX) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dX[i] = dY[i] * act_bwd(Y[i], act);
}
// stand-alone inverted dropout (torch_geometric BasicGNN between its convolutions): y = x * mask / (1 - p); the mask is a
// pure function of (seed, element index), so the backward pass is the same kernel applied to dY
__global__ void __launch_bounds__(256)
    k_dropout(const float *__restrict__ x, float p, unsigned long long seed, const unsigned long long *__restrict__ seed_dev,
              int64_t n, float *__restrict__ y) {
  const unsigned long long s = seed + (seed_dev ? *seed_dev : 0ull);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = x[i] * drop_scale(s, 0x44524f50u, (uint32_t)i, p);
}
__global__ void __launch_bounds__(256)
    k_clamp_div(const float *__restrict__ x, float hi, int64_t n, float *__restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = fminf(fmaxf(x[i], 0.f), hi) / hi;
}

__global__ void __launch_bounds__(256)
    k_biased_score(const float *__restrict__ dot, const float *__restrict__ ub, const float *__restrict__ ib,
                   const float *__restrict__ gb, int64_t M, int act, float *__restrict__ out) {
  const float g = gb[0];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = act_fwd(((dot[i] + ub[i]) + ib[i]) + g, act);
}

}  // namespace fr

extern "C" {

int fr_biased_score(const float *dot, const float *ub, const float *ib, const float *gb, int64_t M, int32_t act, float *out,
                    void *stream) {
  FR_REQUIRE(dot && ub && ib && gb && out && M >= 1, "fr_biased_score: bad argument");
  FR_LAUNCH(fr::k_biased_score, fr::grid_for(M, 256), 256, 0, stream, dot, ub, ib, gb, M, act, out);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_rowdot_forward(const float *A, const float *B, int64_t M, int32_t d, float *out, void *stream) {
  FR_REQUIRE(A && B && out && M >= 1 && d >= 4 && d % 4 == 0, "fr_rowdot_forward: bad argument");
  FR_LAUNCH(fr::k_rowdot_fwd, fr::grid_for(M, 8), 256, 0, stream, A, B, M, d, out);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_rowdot_backward(const float *A, const float *B, const float *dout, int64_t M, int32_t d, float *dA, float *dB,
                       void *stream) {
  FR_REQUIRE(A && B && dout && (dA || dB) && M >= 1 && d % 4 == 0, "fr_rowdot_backward: bad argument");
  FR_LAUNCH(fr::k_rowdot_bwd, fr::grid_for(M * d / 4, 256), 256, 0, stream, A, B, dout, M, d, dA, dB);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_cosine_forward(const float *A, const float *B, int64_t M, int32_t d, float eps, float *out, float *norms,
                      void *stream) {
  FR_REQUIRE(A && B && out && norms && M >= 1 && d >= 4 && d % 4 == 0, "fr_cosine_forward: bad argument");
  FR_LAUNCH(fr::k_cosine_fwd, fr::grid_for(M, 8), 256, 0, stream, A, B, M, d, eps, out, (float2 *)norms);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_cosine_backward(const float *A, const float *B, const float *out, const float *norms, const float *dout, int64_t M,
                       int32_t d, float eps, float *dA, float *dB, void *stream) {
  FR_REQUIRE(A && B && out && norms && dout && (dA || dB) && M >= 1, "fr_cosine_backward: bad argument");
  FR_LAUNCH(fr::k_cosine_bwd, fr::grid_for(M * d, 256), 256, 0, stream, A, B, out, (const float2 *)norms, dout, M, d, eps,
            dA, dB);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_bpr_outer_loss(const float *dotp, const float *dotn, const float *ub, const float *pib, const float *nib,
                      const float *gb, int32_t B, float *loss, float *d_dotp, float *d_dotn, float *d_pib, float *d_nib,
                      float *row_scratch, void *stream) {
  FR_REQUIRE(dotp && dotn && ub && pib && nib && gb && loss && d_dotp && d_dotn && d_pib && d_nib && row_scratch && B >= 1,
             "fr_bpr_outer_loss: bad argument");
  FR_LAUNCH(fr::k_bpr_outer_rows, B, 256, 0, stream, dotp, dotn, ub, pib, nib, gb, B, row_scratch, d_pib, d_nib);
  FR_LAUNCH(fr::k_bpr_outer_cols, B, 256, 0, stream, dotp, dotn, ub, pib, nib, gb, B, (const float *)row_scratch, d_dotp,
            d_dotn, loss);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_scaled_sum(const float *const *xs_host, int32_t n_terms, int64_t n, float scale, int32_t divide, float *out,
                  void *stream) {
  FR_REQUIRE(xs_host && out && n_terms >= 1 && n_terms <= 8 && n >= 1, "fr_scaled_sum: 1..8 terms");
  fr::SumArgs a;
  for (int t = 0; t < 8; ++t) a.x[t] = t < n_terms ? xs_host[t] : nullptr;
  a.n_terms = n_terms;
  a.scale = scale;
  if (divide) {
    FR_LAUNCH(fr::k_sum_div, fr::grid_for(n, 256), 256, 0, stream, a, n, out);
  } else {
    FR_LAUNCH(fr::k_scaled_sum, fr::grid_for(n, 256), 256, 0, stream, a, n, out);
  }
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_copy_cols(const float *src, int32_t ld_src, int32_t col_src, float *dst, int32_t ld_dst, int32_t col_dst, int64_t M,
                 int32_t ncols, void *stream) {
  FR_REQUIRE(src && dst && M >= 1 && ncols >= 1, "fr_copy_cols: bad argument");
  FR_LAUNCH(fr::k_copy_cols, fr::grid_for(M * ncols, 256), 256, 0, stream, src, ld_src, col_src, dst, ld_dst, col_dst, M,
            ncols);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_mse_loss(const float *pred, const float *target, int64_t M, float *loss, float *dpred, void *stream) {
  FR_REQUIRE(pred && target && loss && dpred && M >= 1, "fr_mse_loss: bad argument");
  FR_LAUNCH(fr::k_mse, 1, 1024, 0, stream, pred, target, (int)M, loss, dpred);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_act_forward(const float *x, int32_t act, int64_t n, float *y, void *stream) {
  FR_REQUIRE(x && y && n >= 1, "fr_act_forward: bad argument");
  FR_LAUNCH(fr::k_act_fwd, fr::grid_for(n, 256), 256, 0, stream, x, act, n, y);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_act_backward(const float *dY, const float *Y, int32_t act, int64_t n, float *dX, void *stream) {
  FR_REQUIRE(dY && Y && dX && n >= 1, "fr_act_backward: bad argument");
  FR_LAUNCH(fr::k_act_bwd_out, fr::grid_for(n, 256), 256, 0, stream, dY, Y, act, n, dX);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_dropout(const float *x, float p, uint64_t seed, const uint64_t *seed_dev, int64_t n, float *y, void *stream) {
  FR_REQUIRE(x && y && n >= 1 && n < (1ll << 32) && p >= 0.f && p < 1.f, "fr_dropout: bad argument");
  FR_LAUNCH(fr::k_dropout, fr::grid_for(n, 256), 256, 0, stream, x, p, (unsigned long long)seed,
            (const unsigned long long *)seed_dev, n, y);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_clamp_div(const float *x, float hi, int64_t n, float *y, void *stream) {
  FR_REQUIRE(x && y && n >= 1 && hi > 0.f, "fr_clamp_div: bad argument");
  FR_LAUNCH(fr::k_clamp_div, fr::grid_for(n, 256), 256, 0, stream, x, hi, n, y);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

}  // extern "C"
