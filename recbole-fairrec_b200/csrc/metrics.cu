// Metric accumulation on the device (sm_100a) for the 12 metrics of properties/model/FOCF.yaml:29-30.
//
// Reference being replaced (recbole/evaluator/metrics.py unless noted; numpy + Python loops, float64):
//   Hit 63-65, MRR 89-97, Recall 160-161, NDCG 187-203, base_metric.py:59-82 (mean over users)
//   GiniIndex 644-661, PopularityPercentage 772-820
//   NonParity 860-881, Value/Absolute/Under/Over 935-1266 (full mode), DifferentialFairness 1313-1341
//
// Everything is either integer work (histograms, sorted counts: exact) or float64 sums reduced in a fixed
// order (per-CTA partials -> one CTA), so results are bit-stable run to run.  HBM-bound streaming passes.
#include "sort.cuh"

namespace fr {

constexpr int kMaxGroups = 64;

__device__ __forceinline__ double block_sum_d(double v, double *sh) {  // 256 threads
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) t += sh[i];
    sh[8] = t;
  }
  __syncthreads();
  t = sh[8];
  __syncthreads();
  return t;
}

// ---------------------------------------------------------------- NDCG / Recall / Hit / MRR
// rec_topk [n, K+1] = [hit bits | pos_len].  part[blk][4][K] per-CTA sums, then k_reduce_rows.
__global__ void __launch_bounds__(256)
    k_topk_metrics(const int32_t *__restrict__ rec_topk, int n, int K, double *__restrict__ part) {
  // (one __syncthreads for the whole kernel instead of two per (k, metric) block reduction: the sums of a warp go through
  // shuffles into wpart[warp][metric * K + k], the eight warps are added in warp order at the end)
  __shared__ double disc[kMaxGroups];   // 1/log2(r+1), r = 1..K   (K <= 64)
  __shared__ double idcg[kMaxGroups];   // cumulative
  __shared__ double wpart[8][4 * kMaxGroups];
  if (threadIdx.x < K) disc[threadIdx.x] = 1.0 / log2((double)threadIdx.x + 2.0);
  __syncthreads();
  if (threadIdx.x == 0) {
    double run = 0.0;
    for (int r = 0; r < K; ++r) {
      run += disc[r];
      idcg[r] = run;
    }
  }
  __syncthreads();
  const int u = blockIdx.x * 256 + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool ok = u < n;
  const int32_t *row = rec_topk + (size_t)(ok ? u : 0) * (K + 1);
  const int pos_len = ok ? row[K] : 1;
  int cum = 0, first = -1;
  double dcg = 0.0;
  for (int k = 0; k < K; ++k) {
    const int h = ok ? (row[k] != 0) : 0;
    cum += h;
    if (h && first < 0) first = k;
    if (h) dcg += disc[k];
    const int il = min(pos_len, k + 1);                               // metrics.py:188-196
    double v0 = ok ? dcg / idcg[il - 1] : 0.0;                         // ndcg
    double v1 = ok ? (double)cum / (double)pos_len : 0.0;             // recall, metrics.py:161
    double v2 = (ok && cum > 0) ? 1.0 : 0.0;                           // hit, metrics.py:64-65
    double v3 = (ok && first >= 0) ? 1.0 / (double)(first + 1) : 0.0;  // mrr, metrics.py:91-96
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v0 += __shfl_xor_sync(0xffffffffu, v0, o);
      v1 += __shfl_xor_sync(0xffffffffu, v1, o);
      v2 += __shfl_xor_sync(0xffffffffu, v2, o);
      v3 += __shfl_xor_sync(0xffffffffu, v3, o);
    }
    if (lane == 0) {
      wpart[warp][0 * K + k] = v0;
      wpart[warp][1 * K + k] = v1;
      wpart[warp][2 * K + k] = v2;
      wpart[warp][3 * K + k] = v3;
    }
  }
  __syncthreads();
  for (int v = threadIdx.x; v < 4 * K; v += 256) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += wpart[w][v];
    part[(size_t)blockIdx.x * 4 * K + v] = t;
  }
}

// out[v] = sum_{blk} part[blk][v] in block order (one thread per v)
__global__ void k_reduce_rows(const double *__restrict__ part, int nblk, int V, double *__restrict__ out) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) s += part[(size_t)b * V + v];
  out[v] = s;
}

// ---------------------------------------------------------------- recommendation histograms
// item_pos_count[e][item] += 1 for the item at rank e ; pop_hits[e] += is_popular[item]
__global__ void __launch_bounds__(256)
    k_rec_item_stats(const int32_t *__restrict__ topk_id, int n, int K, int n_items,
                     const uint8_t *__restrict__ is_popular, int32_t *__restrict__ item_pos_count,
                     unsigned long long *__restrict__ pop_hits) {
  const int64_t tot = (int64_t)n * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(i % K);
    const int id = topk_id[i];
    if (id < 0 || id >= n_items) continue;  // sentinel of an under-full list
    atomicAdd(&item_pos_count[(size_t)e * n_items + id], 1);
    if (is_popular && is_popular[id]) atomicAdd(&pop_hits[e], 1ull);
  }
}

__global__ void k_sum_count_rows(const int32_t *__restrict__ item_pos_count, int n_items, int k_rows,
                                 uint32_t *__restrict__ keys, unsigned long long *__restrict__ acc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) acc[0] = 0ull;
  if (i >= n_items) return;
  uint32_t c = 0;
  for (int r = 0; r < k_rows; ++r) c += (uint32_t)item_pos_count[(size_t)r * n_items + i];
  keys[i] = c;
}

// metrics.py:656-660: sum_p (2p - Ni - 1) * c_(p) over the ascending counts (zeros contribute nothing, and
// the non-zero counts occupy exactly the positions Ni-m+1..Ni the reference assigns them)
__global__ void __launch_bounds__(256) k_gini_sum(const uint32_t *__restrict__ sorted, int n_items, long long *acc) {
  long long s = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += gridDim.x * blockDim.x)
    s += (2ll * (i + 1) - n_items - 1) * (long long)sorted[i];
  // integer atomics: associative, so the result does not depend on arrival order
  if (s != 0) atomicAdd((unsigned long long *)acc, (unsigned long long)s);
}
__global__ void k_gini_final(const long long *acc, long long total_recs, int n_items, double *out) {
  out[0] = (double)acc[0] / (double)total_recs / (double)n_items;
}

// Small catalogue (n_items <= kGiniSmall): the row sum, the ascending sort and the weighted sum in ONE CTA (bitonic sort in
// shared memory) instead of the nine launches of the radix path -- at the ML-1M shape (3,707 items) an evaluation pass is a
// chain of small launches and those nine were ~12 % of it.  Integer arithmetic until the final division: same bits.
constexpr int kGiniSmall = 8192;
__global__ void __launch_bounds__(1024) k_gini_small(const int32_t *__restrict__ item_pos_count, int n_items, int k_rows,
                                                     long long total_recs, double *__restrict__ out) {
  __shared__ uint32_t key[kGiniSmall];
  __shared__ long long wsum[32];
  int np2 = 1;
  while (np2 < n_items) np2 <<= 1;
  for (int i = threadIdx.x; i < np2; i += blockDim.x) {
    uint32_t c = 0xffffffffu;   // padding sorts to the end
    if (i < n_items) {
      c = 0;
      for (int r = 0; r < k_rows; ++r) c += (uint32_t)item_pos_count[(size_t)r * n_items + i];
    }
    key[i] = c;
  }
  __syncthreads();
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < np2; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const uint32_t a = key[i], b = key[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            key[i] = b;
            key[l] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  long long s = 0;
  for (int i = threadIdx.x; i < n_items; i += blockDim.x) s += (2ll * (i + 1) - n_items - 1) * (long long)key[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += wsum[w];
    out[0] = (double)t / (double)total_recs / (double)n_items;
  }
}

// Few users (n_users < kGiniBins): an item's summed count is at most n_users (a user recommends an item once), so the sort
// is a histogram over count VALUES: the h items of value v occupy the ascending positions P+1 .. P+h (P = items of smaller
// value) and contribute v * (2 * (h P + h (h + 1) / 2) - h (N + 1)).  One CTA: histogram, scan, sum -- a few microseconds
// where the bitonic kernel above needs 78 synchronised stages.
constexpr int kGiniBins = 8192;
__global__ void __launch_bounds__(1024) k_gini_hist(const int32_t *__restrict__ item_pos_count, int n_items, int k_rows,
                                                    long long total_recs, double *__restrict__ out) {
  __shared__ uint32_t hist[kGiniBins];
  __shared__ uint32_t wtot[32];
  __shared__ long long wsum[32];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int i = t; i < kGiniBins; i += blockDim.x) hist[i] = 0u;
  __syncthreads();
  for (int i = t; i < n_items; i += blockDim.x) {
    uint32_t c = 0;
    for (int r = 0; r < k_rows; ++r) c += (uint32_t)item_pos_count[(size_t)r * n_items + i];
    atomicAdd(&hist[c < (uint32_t)kGiniBins ? c : (uint32_t)kGiniBins - 1u], 1u);
  }
  __syncthreads();
  uint32_t h[kGiniBins / 1024], local = 0;
#pragma unroll
  for (int j = 0; j < kGiniBins / 1024; ++j) {
    h[j] = hist[t * (kGiniBins / 1024) + j];
    local += h[j];
  }
  uint32_t incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += x;
  }
  if (lane == 31) wtot[warp] = incl;
  __syncthreads();
  uint32_t base = 0;
  for (int w = 0; w < warp; ++w) base += wtot[w];
  long long P = (long long)(base + incl - local), s = 0;
#pragma unroll
  for (int j = 0; j < kGiniBins / 1024; ++j) {
    const long long v = t * (kGiniBins / 1024) + j, hj = h[j];
    s += v * (2ll * (hj * P + hj * (hj + 1) / 2) - hj * (n_items + 1));
    P += hj;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) wsum[warp] = s;
  __syncthreads();
  if (t == 0) {
    long long tot = 0;
    for (int w = 0; w < 32; ++w) tot += wsum[w];
    out[0] = (double)tot / (double)total_recs / (double)n_items;
  }
}

// ---------------------------------------------------------------- item x group statistics of the positives
// After a stable sort of the positives by item id, one warp per item segment accumulates (sum score, count)
// per group in float64, lanes striding the segment and a fixed shuffle tree at the end.
__global__ void __launch_bounds__(256)
    k_item_group_stats(const uint32_t *__restrict__ sorted_items, const uint32_t *__restrict__ ord,
                       const int32_t *__restrict__ seg_off, const int32_t *__restrict__ nseg,
                       const float *__restrict__ score, const int32_t *__restrict__ group, int G,
                       double *__restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const int S = *nseg;
  for (int s = warp; s < S; s += nwarps) {
    const int p0 = seg_off[s], p1 = seg_off[s + 1];
    const uint32_t item = sorted_items[p0];
    for (int g = 0; g < G; ++g) {
      double sum = 0.0, cnt = 0.0;
      for (int p = p0 + lane; p < p1; p += 32) {
        const uint32_t e = ord[p];
        if (group[e] == g) {
          sum += (double)score[e];
          cnt += 1.0;
        }
      }
      sum = warp_sum(sum);
      cnt = warp_sum(cnt);
      if (lane == 0) {
        stats[((size_t)item * G + g) * 2 + 0] = sum;
        stats[((size_t)item * G + g) * 2 + 1] = cnt;
      }
    }
  }
}

// pass A: per-CTA partials [1 + 2G]: J (items with any positive), S_g, C_g
__global__ void __launch_bounds__(256) k_fair_pass_a(const double *__restrict__ stats, int n_items, int G,
                                                     double *__restrict__ part, unsigned int *ticket) {
  __shared__ double sh[9];
  const int i = blockIdx.x * 256 + threadIdx.x;
  double any = 0.0;
  for (int g = 0; g < G; ++g) {
    const double s = i < n_items ? stats[((size_t)i * G + g) * 2] : 0.0;
    const double c = i < n_items ? stats[((size_t)i * G + g) * 2 + 1] : 0.0;
    if (c > 0.0) any = 1.0;
    const double ss = block_sum_d(s, sh), cc = block_sum_d(c, sh);
    if (threadIdx.x == 0) {
      part[(size_t)blockIdx.x * (1 + 2 * G) + 1 + g] = ss;
      part[(size_t)blockIdx.x * (1 + 2 * G) + 1 + G + g] = cc;
    }
  }
  const double j = block_sum_d(any, sh);
  if (threadIdx.x == 0) part[(size_t)blockIdx.x * (1 + 2 * G)] = j;
  if (ticket && blockIdx.x == 0 && threadIdx.x == 0) *ticket = 0u;   // (pass B's last-CTA ticket: this launch precedes it)
}

// single thread: totals -> out[5] NonParity (metrics.py:872-881), out[6] = J ; glob[0] = J
__device__ void fair_mid(const double *tot, int G, double *__restrict__ out) {
  const double J = tot[0];
  out[6] = J;
  double mean[kMaxGroups];
  int ng = 0;
  for (int g = 0; g < G; ++g)
    if (tot[1 + G + g] > 0.0) mean[ng++] = tot[1 + g] / tot[1 + G + g];
  double np = nan("");
  if (ng == 2) {
    np = fabs(mean[0] - mean[1]);
  } else if (ng > 2) {
    double mu = 0.0, var = 0.0;
    for (int g = 0; g < ng; ++g) mu += mean[g];
    mu /= ng;
    for (int g = 0; g < ng; ++g) var += (mean[g] - mu) * (mean[g] - mu);
    np = sqrt(var / ng);
  }
  out[5] = np;
}

// pass B: per-CTA partials [5]: sum eps_j (DifferentialFairness), sum |D0-D1| for value/absolute/under/over.
// The whole tail of the metric rides in this launch (it used to be four more: reduce, mid, reduce, final -- an evaluation
// pass at the ML-1M shape is a chain of dependent small launches): every CTA adds pass A's partial J's itself (block
// order, the sum k_reduce_rows made), CTA 0 also the group totals and the NonParity / J outputs, and the CTA that draws the
// last ticket adds pass B's partials in block order and writes the five remaining outputs.
__global__ void __launch_bounds__(256) k_fair_pass_b(const double *__restrict__ stats, int n_items, int G,
                                                     const double *__restrict__ partA, int nblk, double *__restrict__ part,
                                                     unsigned int *ticket, double *__restrict__ out) {
  __shared__ double sh[9];
  __shared__ double totA[1 + 2 * kMaxGroups];
  __shared__ bool last;
  const int i = blockIdx.x * 256 + threadIdx.x;
  const int VA = 1 + 2 * G;
  // J is a count (integer-valued: exact in any order), so every CTA adds the partial J's with all its threads -- a serial
  // walk per CTA is O(nblk^2) in total: 0.7 ms at 1 M items; the group totals (order matters) only CTA 0 needs
  double jl = 0.0;
  for (int b = threadIdx.x; b < nblk; b += 256) jl += partA[(size_t)b * VA];
  const double J = block_sum_d(jl, sh);
  if (blockIdx.x == 0) {
    if (threadIdx.x >= 1 && threadIdx.x < VA) {
      double t = 0.0;
      for (int b = 0; b < nblk; ++b) t += partA[(size_t)b * VA + threadIdx.x];
      totA[threadIdx.x] = t;
    }
    if (threadIdx.x == 0) totA[0] = J;
    __syncthreads();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) fair_mid(totA, G, out);
  double eps = 0.0, vv = 0.0, va = 0.0, vu = 0.0, vo = 0.0;
  bool any = false;
  if (i < n_items) {
    for (int g = 0; g < G; ++g) any |= stats[((size_t)i * G + g) * 2 + 1] > 0.0;
  }
  if (any) {
    // metrics.py:1329-1339: M = (sum + 1/J) / (cnt + 1) stored as float32, eps = max_{g<g'} |ln M_g - ln M_g'|
    const double alpha = 1.0 / J;
    float lmin = INFINITY, lmax = -INFINITY;
    for (int g = 0; g < G; ++g) {
      const double s = stats[((size_t)i * G + g) * 2], c = stats[((size_t)i * G + g) * 2 + 1];
      const float M = (float)((s + alpha) / (c + 1.0));
      const float l = (float)log((double)M);
      lmin = fminf(lmin, l);
      lmax = fmaxf(lmax, l);
    }
    eps = (double)(lmax - lmin);
    if (G == 2) {  // metrics.py:965-978 and the three siblings
      const double s0 = stats[((size_t)i * 2 + 0) * 2], c0 = stats[((size_t)i * 2 + 0) * 2 + 1];
      const double s1 = stats[((size_t)i * 2 + 1) * 2], c1 = stats[((size_t)i * 2 + 1) * 2 + 1];
      const double n0 = c0 + 1e-5, n1 = c1 + 1e-5;
      const double P0 = s0 / n0, P1 = s1 / n1, T0 = c0 / n0, T1 = c1 / n1;
      vv = fabs((P0 - T0) - (P1 - T1));
      va = fabs(fabs(P0 - T0) - fabs(P1 - T1));
      vu = fabs(fmax(T0 - P0, 0.0) - fmax(T1 - P1, 0.0));
      vo = fabs(fmax(P0 - T0, 0.0) - fmax(P1 - T1, 0.0));
    }
  }
  const double r0 = block_sum_d(eps, sh), r1 = block_sum_d(vv, sh), r2 = block_sum_d(va, sh), r3 = block_sum_d(vu, sh),
               r4 = block_sum_d(vo, sh);
  if (threadIdx.x == 0) {
    double *o = part + (size_t)blockIdx.x * 5;
    o[0] = r0; o[1] = r1; o[2] = r2; o[3] = r3; o[4] = r4;
    __threadfence();
    last = atomicAdd(ticket, 1u) == (unsigned int)(gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x < 5) {
    __threadfence();
    double t = 0.0;
    for (int b = 0; b < nblk; ++b) t += __ldcg(part + (size_t)b * 5 + threadIdx.x);
    if (threadIdx.x == 0) out[0] = t / J;
    else out[threadIdx.x] = (G == 2) ? t / J : nan("");
  }
}


// Sampled-negative ("uni<N>") mode of Value / Absolute / Under / Over unfairness (metrics.py:935-978 and its three siblings
// with mode != 'full'): the item set is the union of the positives' and the paired negatives' items; a negative adds its
// score and a count to (its item, the group of the positive's user) but nothing to the "true" sum.
// all [n_items, 2, 2] = (sum score, count) over positives AND negatives; pos [n_items, 2, 2] over the positives only.
// per-CTA partials [5]: J (items with any entry), sum |D0 - D1| for value / absolute / under / over.
__global__ void __launch_bounds__(256) k_unfair_sampled_part(const double *__restrict__ all, const double *__restrict__ pos,
                                                             int n_items, double *__restrict__ part) {
  __shared__ double sh[9];
  const int i = blockIdx.x * 256 + threadIdx.x;
  double any = 0.0, vv = 0.0, va = 0.0, vu = 0.0, vo = 0.0;
  if (i < n_items) {
    const double s0 = all[((size_t)i * 2 + 0) * 2], c0 = all[((size_t)i * 2 + 0) * 2 + 1];
    const double s1 = all[((size_t)i * 2 + 1) * 2], c1 = all[((size_t)i * 2 + 1) * 2 + 1];
    if (c0 + c1 > 0.0) {
      any = 1.0;
      const double t0 = pos[((size_t)i * 2 + 0) * 2 + 1], t1 = pos[((size_t)i * 2 + 1) * 2 + 1];
      const double n0 = c0 + 1e-5, n1 = c1 + 1e-5;
      const double P0 = s0 / n0, P1 = s1 / n1, T0 = t0 / n0, T1 = t1 / n1;
      vv = fabs((P0 - T0) - (P1 - T1));
      va = fabs(fabs(P0 - T0) - fabs(P1 - T1));
      vu = fabs(fmax(T0 - P0, 0.0) - fmax(T1 - P1, 0.0));
      vo = fabs(fmax(P0 - T0, 0.0) - fmax(P1 - T1, 0.0));
    }
  }
  const double r0 = block_sum_d(any, sh), r1 = block_sum_d(vv, sh), r2 = block_sum_d(va, sh), r3 = block_sum_d(vu, sh),
               r4 = block_sum_d(vo, sh);
  if (threadIdx.x == 0) {
    double *o = part + (size_t)blockIdx.x * 5;
    o[0] = r0; o[1] = r1; o[2] = r2; o[3] = r3; o[4] = r4;
  }
}

__global__ void k_unfair_sampled_final(const double *__restrict__ tot, double *__restrict__ out) {
  const double J = tot[0];
  for (int m = 0; m < 4; ++m) out[m] = tot[1 + m] / J;
  out[4] = J;
}

}  // namespace fr

extern "C" {

size_t fr_topk_metrics_workspace_bytes(int32_t n, int32_t K) { return (size_t)((n + 255) / 256) * 4 * K * 8 + 256; }

int fr_topk_metrics(const int32_t *rec_topk, int32_t n, int32_t K, double *sums_out, void *workspace,
                    size_t workspace_bytes, void *stream) {
  FR_REQUIRE(rec_topk && sums_out && workspace && n >= 1, "fr_topk_metrics: bad argument");
  FR_REQUIRE(K >= 1 && K <= fr::kMaxGroups, "fr_topk_metrics: K=%d out of range", K);
  const int nblk = (n + 255) / 256;
  if (workspace_bytes < fr_topk_metrics_workspace_bytes(n, K)) {
    fr::set_error("fr_topk_metrics: workspace too small");
    return FR_ERR_WORKSPACE;
  }
  FR_LAUNCH(fr::k_topk_metrics, nblk, 256, 0, stream, rec_topk, n, K, (double *)workspace);
  FR_LAUNCH(fr::k_reduce_rows, (4 * K + 127) / 128, 128, 0, stream, (const double *)workspace, nblk, 4 * K, sums_out);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_rec_item_stats(const int32_t *topk_id, int32_t n, int32_t K, int32_t n_items, const uint8_t *is_popular,
                      int32_t *item_count_out, int64_t *pop_hits_out, void *stream) {
  FR_REQUIRE(topk_id && item_count_out && pop_hits_out && n >= 1 && K >= 1 && n_items >= 1,
             "fr_rec_item_stats: bad argument");
  FR_CUDA_OK(cudaMemsetAsync(item_count_out, 0, sizeof(int32_t) * (size_t)K * n_items, (cudaStream_t)stream));
  FR_CUDA_OK(cudaMemsetAsync(pop_hits_out, 0, sizeof(int64_t) * (size_t)K, (cudaStream_t)stream));
  FR_LAUNCH(fr::k_rec_item_stats, fr::grid_for((int64_t)n * K, 256), 256, 0, stream, topk_id, n, K, n_items,
            is_popular, item_count_out, (unsigned long long *)pop_hits_out);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

size_t fr_gini_workspace_bytes(int32_t n_items) {
  fr::Carver c(nullptr, 0);
  c.take<uint32_t>(n_items);  // keys
  c.take<uint32_t>(n_items);  // sorted keys
  c.take<uint32_t>(n_items);  // sorted vals (unused)
  c.take<long long>(1);
  fr::carve_sort_scratch(c, n_items);
  return c.off;
}

int fr_gini_at_k(const int32_t *item_pos_count, int32_t n_items, int32_t k_rows, int64_t n_users, double *gini_out,
                 void *workspace, size_t workspace_bytes, void *stream) {
  FR_REQUIRE(item_pos_count && gini_out && workspace && n_items >= 1 && k_rows >= 1 && n_users >= 1,
             "fr_gini_at_k: bad argument");
  fr::Carver c(workspace, workspace_bytes);
  uint32_t *keys = c.take<uint32_t>(n_items);
  uint32_t *skeys = c.take<uint32_t>(n_items);
  uint32_t *svals = c.take<uint32_t>(n_items);
  long long *acc = c.take<long long>(1);
  fr::SortScratch ss = fr::carve_sort_scratch(c, n_items);
  if (!c.ok()) {
    fr::set_error("fr_gini_at_k: workspace too small (%zu < %zu)", workspace_bytes, c.off);
    return FR_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (n_users < fr::kGiniBins && n_items <= (1 << 18)) {
    FR_LAUNCH(fr::k_gini_hist, 1, 1024, 0, st, item_pos_count, n_items, k_rows, (long long)(n_users * k_rows), gini_out);
    FR_LAUNCH_CHECK();
    return FR_OK;
  }
  if (n_items <= fr::kGiniSmall) {
    FR_LAUNCH(fr::k_gini_small, 1, 1024, 0, st, item_pos_count, n_items, k_rows, (long long)(n_users * k_rows), gini_out);
    FR_LAUNCH_CHECK();
    return FR_OK;
  }
  FR_LAUNCH(fr::k_sum_count_rows, (n_items + 255) / 256, 256, 0, st, item_pos_count, n_items, k_rows, keys,
            (unsigned long long *)acc);
  const uint64_t maxc = (uint64_t)n_users + 1;
  fr::sort_pairs(keys, nullptr, skeys, svals, n_items, nullptr, fr::bits_for((uint32_t)(maxc > 0xffffffffull ? 0xffffffffu : maxc)),
                 ss, st);
  FR_LAUNCH(fr::k_gini_sum, fr::grid_for(n_items, 256, fr::kSMs * 4), 256, 0, st, skeys, n_items, acc);
  FR_LAUNCH(fr::k_gini_final, 1, 1, 0, st, acc, (long long)(n_users * k_rows), n_items, gini_out);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

size_t fr_item_group_stats_workspace_bytes(int64_t n_pos, int32_t n_items, int32_t G) {
  (void)n_items; (void)G;
  fr::Carver c(nullptr, 0);
  c.take<uint32_t>(n_pos);
  c.take<uint32_t>(n_pos);
  c.take<int32_t>(n_pos);
  c.take<int32_t>(n_pos + 1);
  c.take<int32_t>(1);
  fr::carve_sort_scratch(c, n_pos);
  fr::carve_seg_scratch(c, n_pos);
  return c.off;
}

int fr_item_group_stats(const int32_t *pos_items, const float *pos_score, const int32_t *group, int64_t n_pos,
                        int32_t n_items, int32_t G, double *stats_out, void *workspace, size_t workspace_bytes,
                        void *stream) {
  FR_REQUIRE(pos_items && pos_score && group && stats_out && workspace, "fr_item_group_stats: null pointer");
  FR_REQUIRE(n_pos >= 1 && n_pos < (1ll << 31) && n_items >= 1 && G >= 1 && G <= fr::kMaxGroups,
             "fr_item_group_stats: bad sizes n_pos=%lld G=%d", (long long)n_pos, G);
  fr::Carver c(workspace, workspace_bytes);
  uint32_t *skey = c.take<uint32_t>(n_pos);
  uint32_t *ord = c.take<uint32_t>(n_pos);
  int32_t *segid = c.take<int32_t>(n_pos);
  int32_t *segoff = c.take<int32_t>(n_pos + 1);
  int32_t *nseg = c.take<int32_t>(1);
  fr::SortScratch ss = fr::carve_sort_scratch(c, n_pos);
  fr::SegScratch sg = fr::carve_seg_scratch(c, n_pos);
  if (!c.ok()) {
    fr::set_error("fr_item_group_stats: workspace too small (%zu < %zu)", workspace_bytes, c.off);
    return FR_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  FR_CUDA_OK(cudaMemsetAsync(stats_out, 0, sizeof(double) * 2 * (size_t)n_items * G, st));
  fr::sort_pairs((const uint32_t *)pos_items, nullptr, skey, ord, n_pos, nullptr, fr::bits_for((uint32_t)n_items), ss, st);
  fr::build_segments(skey, ord, n_pos, nullptr, segid, segoff, nseg, nullptr, nullptr, nullptr, sg, st);
  FR_LAUNCH(fr::k_item_group_stats, fr::grid_for(n_pos, 256, fr::kSMs * 8), 256, 0, st, skey, ord, segoff, nseg,
            pos_score, group, G, stats_out);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

namespace {
struct IgPlan {
  uint32_t *skey, *ord;
  int32_t *segoff, *nseg;
  bool ok;
  size_t bytes;
};
IgPlan carve_ig_plan(void *plan, size_t plan_bytes, int64_t n_pos) {
  fr::Carver c(plan, plan_bytes);
  IgPlan p;
  p.skey = c.take<uint32_t>(n_pos);
  p.ord = c.take<uint32_t>(n_pos);
  p.segoff = c.take<int32_t>(n_pos + 1);
  p.nseg = c.take<int32_t>(1);
  p.ok = c.ok();
  p.bytes = c.off;
  return p;
}
}  // namespace

size_t fr_item_group_plan_bytes(int64_t n_pos) { return carve_ig_plan(nullptr, 0, n_pos).bytes; }

size_t fr_item_group_plan_workspace_bytes(int64_t n_pos) {
  fr::Carver c(nullptr, 0);
  c.take<int32_t>(n_pos);
  fr::carve_sort_scratch(c, n_pos);
  fr::carve_seg_scratch(c, n_pos);
  return c.off;
}

int fr_item_group_plan(const int32_t *pos_items, int64_t n_pos, int32_t n_items, void *plan, size_t plan_bytes,
                       void *workspace, size_t workspace_bytes, void *stream) {
  FR_REQUIRE(pos_items && plan && workspace, "fr_item_group_plan: null pointer");
  FR_REQUIRE(n_pos >= 1 && n_pos < (1ll << 31) && n_items >= 1, "fr_item_group_plan: bad sizes n_pos=%lld", (long long)n_pos);
  IgPlan p = carve_ig_plan(plan, plan_bytes, n_pos);
  fr::Carver c(workspace, workspace_bytes);
  int32_t *segid = c.take<int32_t>(n_pos);
  fr::SortScratch ss = fr::carve_sort_scratch(c, n_pos);
  fr::SegScratch sg = fr::carve_seg_scratch(c, n_pos);
  if (!p.ok || !c.ok()) {
    fr::set_error("fr_item_group_plan: plan or workspace too small (%zu < %zu or %zu < %zu)", plan_bytes, p.bytes,
                  workspace_bytes, c.off);
    return FR_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  fr::sort_pairs((const uint32_t *)pos_items, nullptr, p.skey, p.ord, n_pos, nullptr, fr::bits_for((uint32_t)n_items), ss, st);
  fr::build_segments(p.skey, p.ord, n_pos, nullptr, segid, p.segoff, p.nseg, nullptr, nullptr, nullptr, sg, st);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_item_group_stats_planned(const void *plan, size_t plan_bytes, const float *pos_score, const int32_t *group,
                                int64_t n_pos, int32_t n_items, int32_t G, double *stats_out, void *stream) {
  FR_REQUIRE(plan && pos_score && group && stats_out, "fr_item_group_stats_planned: null pointer");
  FR_REQUIRE(n_pos >= 1 && n_pos < (1ll << 31) && n_items >= 1 && G >= 1 && G <= fr::kMaxGroups,
             "fr_item_group_stats_planned: bad sizes n_pos=%lld G=%d", (long long)n_pos, G);
  IgPlan p = carve_ig_plan(const_cast<void *>(plan), plan_bytes, n_pos);
  if (!p.ok) {
    fr::set_error("fr_item_group_stats_planned: plan too small (%zu < %zu)", plan_bytes, p.bytes);
    return FR_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  FR_CUDA_OK(cudaMemsetAsync(stats_out, 0, sizeof(double) * 2 * (size_t)n_items * G, st));
  FR_LAUNCH(fr::k_item_group_stats, fr::grid_for(n_pos, 256, fr::kSMs * 8), 256, 0, st, p.skey, p.ord, p.segoff, p.nseg,
            pos_score, group, G, stats_out);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

size_t fr_fairness_metrics_workspace_bytes(int32_t n_items, int32_t G) {
  const size_t nblk = (size_t)(n_items + 255) / 256;
  return (nblk * (1 + 2 * (size_t)G) + nblk * 5 + (1 + 2 * (size_t)G) + 8) * 8 + 512;
}

int fr_fairness_metrics(const double *stats, int32_t n_items, int32_t G, double *out, void *workspace,
                        size_t workspace_bytes, void *stream) {
  FR_REQUIRE(stats && out && workspace && n_items >= 1 && G >= 1 && G <= fr::kMaxGroups,
             "fr_fairness_metrics: bad argument");
  if (workspace_bytes < fr_fairness_metrics_workspace_bytes(n_items, G)) {
    fr::set_error("fr_fairness_metrics: workspace too small");
    return FR_ERR_WORKSPACE;
  }
  const int nblk = (n_items + 255) / 256, VA = 1 + 2 * G;
  fr::Carver c(workspace, workspace_bytes);
  double *partA = c.take<double>((size_t)nblk * VA);
  double *partB = c.take<double>((size_t)nblk * 5);
  double *totA = c.take<double>(VA);
  double *totB = c.take<double>(8);
  (void)totA;
  unsigned int *ticket = (unsigned int *)totB;
  FR_LAUNCH(fr::k_fair_pass_a, nblk, 256, 0, stream, stats, n_items, G, partA, ticket);
  FR_LAUNCH(fr::k_fair_pass_b, nblk, 256, 0, stream, stats, n_items, G, (const double *)partA, nblk, partB, ticket, out);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

size_t fr_unfairness_sampled_workspace_bytes(int32_t n_items) {
  return ((size_t)(n_items + 255) / 256 * 5 + 8) * 8 + 512;
}

int fr_unfairness_sampled(const double *stats_all, const double *stats_pos, int32_t n_items, double *out, void *workspace,
                          size_t workspace_bytes, void *stream) {
  FR_REQUIRE(stats_all && stats_pos && out && workspace && n_items >= 1, "fr_unfairness_sampled: bad argument");
  if (workspace_bytes < fr_unfairness_sampled_workspace_bytes(n_items)) {
    fr::set_error("fr_unfairness_sampled: workspace too small");
    return FR_ERR_WORKSPACE;
  }
  const int nblk = (n_items + 255) / 256;
  fr::Carver c(workspace, workspace_bytes);
  double *part = c.take<double>((size_t)nblk * 5);
  double *tot = c.take<double>(8);
  FR_LAUNCH(fr::k_unfair_sampled_part, nblk, 256, 0, stream, stats_all, stats_pos, n_items, part);
  FR_LAUNCH(fr::k_reduce_rows, 1, 128, 0, stream, (const double *)part, nblk, 5, tot);
  FR_LAUNCH(fr::k_unfair_sampled_final, 1, 1, 0, stream, (const double *)tot, out);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

}  // extern "C"
