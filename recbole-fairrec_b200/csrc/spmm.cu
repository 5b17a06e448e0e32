// CSR SpMM  Y = A . X  for the FairGo ego-network aggregation (kernel K7) on sm_100a.
//
// Reference being replaced (paths relative to the reference root):
//   recbole/model/fair_recommender/fairgo_pmf.py:199-202 / fairgo_gcn.py:213-216
//       for _ in range(n_layers): all_embeddings = torch.sparse.mm(self.norm_rating_matrix, all_embeddings)
//   and its autograd backward (dX = A^T . dY: the caller passes the CSR of A^T).
//   norm_rating_matrix = D^-1 A over the (n_users + n_items) bipartite rating graph (fairgo_pmf.py:100-127).
//
// L2-bound gather: N = n_users + n_items rows of d floats (2.5 MB at the ML-1M shape) are re-read nnz times.  Item rows
// are up to thousands of nnz long while user rows average ~165, so rows are cut into chunks of <= `chunk` nnz by a
// host-side planner (the matrix is static); a group of d/4 lanes owns one chunk and walks its nnz with 8 independent
// 128-bit row loads in flight; rows spanning several chunks go through per-chunk partials that a second kernel adds
// in chunk order.  No atomics: bit-reproducible run to run.
#include "common.cuh"

namespace fr {

struct SpmmArgs {
  const int32_t *chunk_row, *chunk_begin, *chunk_end, *chunk_slot;
  const int32_t *col;
  const float *val, *X;
  float *Y, *partial;
  int n_chunks, d, lanes;       // lanes per chunk: power of two >= d/4, <= 32
};

__global__ void __launch_bounds__(256) k_spmm_chunks(SpmmArgs a) {
  const int per_block = 256 / a.lanes;
  const int sub = threadIdx.x % a.lanes;
  const int dq = a.d >> 2;
  for (int c = blockIdx.x * per_block + threadIdx.x / a.lanes; c < a.n_chunks; c += gridDim.x * per_block) {
    const int p0 = a.chunk_begin[c], p1 = a.chunk_end[c];
    const int slot = a.chunk_slot[c];
    float *dst = slot < 0 ? a.Y + (size_t)a.chunk_row[c] * a.d : a.partial + (size_t)slot * a.d;
    for (int q = sub; q < dq; q += a.lanes) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int p = p0;
      for (; p + 8 <= p1; p += 8) {
        int cj[8];
        float vj[8];
        float4 xj[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          cj[u] = __ldg(a.col + p + u);
          vj[u] = __ldg(a.val + p + u);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) xj[u] = __ldg((const float4 *)(a.X + (size_t)cj[u] * a.d) + q);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          acc.x = fmaf(vj[u], xj[u].x, acc.x); acc.y = fmaf(vj[u], xj[u].y, acc.y);
          acc.z = fmaf(vj[u], xj[u].z, acc.z); acc.w = fmaf(vj[u], xj[u].w, acc.w);
        }
      }
      for (; p < p1; ++p) {
        const float v = __ldg(a.val + p);
        const float4 x = __ldg((const float4 *)(a.X + (size_t)__ldg(a.col + p) * a.d) + q);
        acc.x = fmaf(v, x.x, acc.x); acc.y = fmaf(v, x.y, acc.y); acc.z = fmaf(v, x.z, acc.z); acc.w = fmaf(v, x.w, acc.w);
      }
      ((float4 *)dst)[q] = acc;
    }
  }
}

// rows cut into several chunks: Y[row] = partial[first] + partial[first+1] + ... in chunk order
__global__ void __launch_bounds__(256)
    k_spmm_combine(const int32_t *__restrict__ multi_row, const int32_t *__restrict__ multi_first, int n_multi,
                   const float *__restrict__ partial, int d, float *__restrict__ Y) {
  const int64_t n = (int64_t)n_multi * d;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(q / d), c = (int)(q % d);
    const int s0 = multi_first[r], s1 = multi_first[r + 1];
    float acc = partial[(size_t)s0 * d + c];
    for (int s = s0 + 1; s < s1; ++s) acc += partial[(size_t)s * d + c];
    Y[(size_t)multi_row[r] * d + c] = acc;
  }
}

__global__ void __launch_bounds__(256)
    k_zero_rows(const int32_t *__restrict__ rows, int n, int d, float *__restrict__ Y) {
  const int64_t tot = (int64_t)n * d;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < tot; q += (int64_t)gridDim.x * blockDim.x)
    Y[(size_t)rows[q / d] * d + q % d] = 0.f;
}

}  // namespace fr

extern "C" {

// ---- host-side planner (pure CPU, no CUDA calls): the adjacency is static, plan once per matrix
int fr_spmm_plan_sizes(const int64_t *row_off_host, int32_t n_rows, int32_t chunk, int64_t *n_chunks, int64_t *n_multi,
                       int64_t *n_slots, int64_t *n_empty) {
  FR_REQUIRE(row_off_host && n_rows >= 1 && chunk >= 8 && n_chunks && n_multi && n_slots && n_empty,
             "fr_spmm_plan_sizes: bad argument");
  int64_t c = 0, m = 0, s = 0, e = 0;
  for (int32_t r = 0; r < n_rows; ++r) {
    const int64_t len = row_off_host[r + 1] - row_off_host[r];
    if (len == 0) { ++e; continue; }
    const int64_t k = (len + chunk - 1) / chunk;
    c += k;
    if (k > 1) { ++m; s += k; }
  }
  *n_chunks = c; *n_multi = m; *n_slots = s; *n_empty = e;
  return FR_OK;
}

int fr_spmm_plan_fill(const int64_t *row_off_host, int32_t n_rows, int32_t chunk, int32_t *chunk_row_host,
                      int32_t *chunk_begin_host, int32_t *chunk_end_host, int32_t *chunk_slot_host, int32_t *multi_row_host,
                      int32_t *multi_first_host, int32_t *empty_row_host) {
  FR_REQUIRE(row_off_host && chunk_row_host && chunk_begin_host && chunk_end_host && chunk_slot_host && multi_row_host &&
                 multi_first_host && empty_row_host && row_off_host[n_rows] < (int64_t)INT32_MAX,
             "fr_spmm_plan_fill: bad argument (nnz must fit int32)");
  int64_t c = 0, m = 0, s = 0, e = 0;
  for (int32_t r = 0; r < n_rows; ++r) {
    const int64_t b = row_off_host[r], en = row_off_host[r + 1];
    if (en == b) { empty_row_host[e++] = r; continue; }
    const int64_t k = (en - b + chunk - 1) / chunk;
    if (k > 1) { multi_row_host[m] = r; multi_first_host[m] = (int32_t)s; ++m; }
    for (int64_t i = 0; i < k; ++i) {
      chunk_row_host[c] = r;
      chunk_begin_host[c] = (int32_t)(b + i * chunk);
      chunk_end_host[c] = (int32_t)((b + (i + 1) * chunk < en) ? b + (i + 1) * chunk : en);
      chunk_slot_host[c] = k > 1 ? (int32_t)(s++) : -1;
      ++c;
    }
  }
  multi_first_host[m] = (int32_t)s;
  return FR_OK;
}

int fr_spmm_csr(const fr_spmm_plan *p, const int32_t *col, const float *val, const float *X, int32_t d, float *Y,
                float *partial, void *stream) {
  FR_REQUIRE(p && col && val && X && Y && d >= 4 && d % 4 == 0, "fr_spmm_csr: bad argument");
  FR_REQUIRE(p->n_slots == 0 || partial, "fr_spmm_csr: partial buffer [n_slots, d] required");
  if (p->n_empty > 0)
    FR_LAUNCH(fr::k_zero_rows, fr::grid_for((int64_t)p->n_empty * d, 256), 256, 0, stream, p->empty_row, (int)p->n_empty, d, Y);
  if (p->n_chunks > 0) {
    int lanes = 1;
    while (lanes < d / 4 && lanes < 32) lanes <<= 1;
    fr::SpmmArgs a{p->chunk_row, p->chunk_begin, p->chunk_end, p->chunk_slot, col, val, X, Y, partial, (int)p->n_chunks, d, lanes};
    FR_LAUNCH(fr::k_spmm_chunks, fr::grid_for(p->n_chunks, 256 / lanes, fr::kSMs * 32), 256, 0, stream, a);
  }
  if (p->n_multi > 0)
    FR_LAUNCH(fr::k_spmm_combine, fr::grid_for((int64_t)p->n_multi * d, 256), 256, 0, stream, p->multi_row, p->multi_first,
              (int)p->n_multi, (const float *)partial, d, Y);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

}  // extern "C"
