// Full-sort evaluation for sm_100a: U.I^T scoring fused with the pad/history mask and a streaming top-K.
// The [n_users, n_items] score matrix is never materialised.
//
// Reference being replaced (paths relative to the reference root):
//   recbole/model/fair_recommender/focf.py:171-178   full_sort_predict (mm + clamp/max_rating)
//   recbole/trainer/trainer.py:435-438               scores[:,0] = -inf ; scores[history] = -inf
//   recbole/evaluator/collector.py:143-153           topk, pos-matrix scatter, gather (hit bits), pos_len
//
// This file holds the FR_SCORE_EXACT_FP32 scorer: a register-tiled CUDA-core contraction whose every
// output is the k-ascending chain acc = fmaf(u[k], i[k], acc) -- bit-identical to oracle/c, which is what
// makes "top-K indices bit-exact, ties broken by lowest item id" checkable.  (The tcgen05 3xTF32 scorer
// lives in fullsort_tc.cu and shares the epilogue below.)
//
// CTA tile: 64 eval users x 128 items, 256 threads, thread (tx,ty) owns users ty+16m (m<4) and items
// tx+16j (j<8) (interleaved so that the 16 distinct item rows a warp reads per LDS.128 hit all 32 banks).
// The user tile stays in shared memory for the CTA's lifetime; item tiles stream through a 3-stage
// cp.async ring in k-chunks of 64.  Per tile the epilogue applies transform + mask from a shared-memory
// bitmask (built from the per-user sorted history CSR with running pointers), compares against the row's
// current K-th best and pushes the rare survivors into per-row queues that one thread per row merges into
// a sorted K-list under the total order (score desc, item id asc).
#include "common.cuh"

namespace fr {

constexpr int TU = 64;        // users per CTA
constexpr int TI = 128;       // items per tile
constexpr int KC = 64;        // k-chunk (floats) per pipeline stage
constexpr int LDI = KC + 4;   // padded row stride of an item stage (floats)
constexpr int NSTAGE = 3;
constexpr int QCAP = 32;      // candidate queue capacity per row
constexpr int kMaxK = 64;
constexpr int kMaxDExact = 256;
constexpr int kMaxSplits = 32;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

// total order of the ranking: higher score first, then lower item id
__device__ __forceinline__ bool better(float s, int id, float s2, int id2) { return s > s2 || (s == s2 && id < id2); }

__device__ __forceinline__ float score_transform(float x, int transform, float max_rating) {
  if (transform == FR_TRANSFORM_CLAMP_DIV) return __fdiv_rn(fminf(fmaxf(x, 0.f), max_rating), max_rating);
  if (transform == FR_TRANSFORM_SIGMOID) return 1.f / (1.f + expf(-x));
  return x;
}

struct EvalArgs {
  const float *U, *I;
  const int32_t *users;
  const int64_t *hist_off;
  const int32_t *hist_items;
  int n, d, n_items_local, item_base, K, transform;
  float max_rating;
  int tiles_per_split, n_splits;
  int32_t *out_id;   // [n_splits, n, K] (n_splits == 1: the final output)
  float *out_score;
};

// insert (s,id) into the row's sorted K-list (descending in the total order); list[K-1] is the threshold
__device__ __forceinline__ void list_insert(float *ls, int *li, int K, float s, int id) {
  if (!better(s, id, ls[K - 1], li[K - 1])) return;
  int p = K - 1;
  while (p > 0 && better(s, id, ls[p - 1], li[p - 1])) {
    ls[p] = ls[p - 1];
    li[p] = li[p - 1];
    --p;
  }
  ls[p] = s;
  li[p] = id;
}

__global__ void __launch_bounds__(256) k_fullsort_exact(EvalArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int d = a.d, K = a.K, ldu = d + 4, nkc = d / KC + (d % KC ? 1 : 0);
  float *Us = (float *)smem_raw;                       // [TU][ldu]
  float *Is = Us + TU * ldu;                           // [NSTAGE][TI][LDI]
  float *list_s = Is + NSTAGE * TI * LDI;              // [TU][K]
  int *list_i = (int *)(list_s + TU * K);              // [TU][K]
  float *q_s = (float *)(list_i + TU * K);             // [TU][QCAP]
  int *q_i = (int *)(q_s + TU * QCAP);                 // [TU][QCAP]
  int *q_cnt = q_i + TU * QCAP;                        // [TU]
  uint32_t *hmask = (uint32_t *)(q_cnt + TU);          // [TU][TI/32]
  long long *hptr = (long long *)(hmask + TU * (TI / 32));  // [TU] running history pointers (8B aligned)

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int u0 = blockIdx.x * TU;
  const int split = blockIdx.y;
  const int n_tiles_total = (a.n_items_local + TI - 1) / TI;
  const int tile_lo = split * a.tiles_per_split;
  const int tile_hi = min(n_tiles_total, tile_lo + a.tiles_per_split);
  const int ntile = max(0, tile_hi - tile_lo);

  // ---- per-row state
  for (int i = tid; i < TU * K; i += 256) {
    list_s[i] = -INFINITY;
    list_i[i] = 0x7fffffff;
  }
  if (tid < TU) {
    q_cnt[tid] = 0;
    const int r = u0 + tid;
    long long hp = 0;
    if (r < a.n) {  // lower_bound of the first item of this CTA's range in the row's sorted history
      long long lo = a.hist_off[r], hi = a.hist_off[r + 1];
      const int first = a.item_base + tile_lo * TI;
      while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (a.hist_items[mid] < first) lo = mid + 1; else hi = mid;
      }
      hp = lo;
    }
    hptr[tid] = hp;
  }
  // ---- user tile -> smem (rows beyond n are clamped to row u0: computed, never emitted)
  {
    const int dq = d >> 2;
    for (int f = tid; f < TU * dq; f += 256) {
      const int row = f / dq, c4 = f % dq;
      const int r = (u0 + row < a.n) ? u0 + row : u0;
      cp_async16(Us + row * ldu + c4 * 4, a.U + (size_t)a.users[r] * d + c4 * 4);
    }
  }
  const int nstage_total = ntile * nkc;
  auto issue_stage = [&](int q) {
    if (q < nstage_total) {
      const int t = tile_lo + q / nkc, kc = q % nkc;
      const int kw = min(KC, d - kc * KC) >> 2;  // float4 per row in this chunk
      float *dst = Is + (q % NSTAGE) * TI * LDI;
      for (int f = tid; f < TI * kw; f += 256) {
        const int row = f / kw, c4 = f % kw;
        int it = t * TI + row;
        if (it >= a.n_items_local) it = a.n_items_local - 1;
        cp_async16(dst + row * LDI + c4 * 4, a.I + (size_t)it * d + kc * KC + c4 * 4);
      }
    }
    cp_async_commit();
  };
  issue_stage(0);  // group 0 also carries the user tile
  issue_stage(1);

  float acc[4][8];
  for (int tt = 0; tt < ntile; ++tt) {
    const int t = tile_lo + tt;
    const int tile_base = t * TI;  // local item index of the tile's first column
    // history bitmask of this tile: one thread per user row walks its sorted history
    if (tid < TU) {
      uint32_t bits[TI / 32];
#pragma unroll
      for (int w = 0; w < TI / 32; ++w) bits[w] = 0u;
      const int r = u0 + tid;
      if (r < a.n) {
        long long hp = hptr[tid];
        const long long hend = a.hist_off[r + 1];
        const int g0 = a.item_base + tile_base, g1 = g0 + TI;
        while (hp < hend) {
          const int it = a.hist_items[hp];
          if (it >= g1) break;
          if (it >= g0) bits[(it - g0) >> 5] |= 1u << ((it - g0) & 31);
          ++hp;
        }
        hptr[tid] = hp;
      }
#pragma unroll
      for (int w = 0; w < TI / 32; ++w) hmask[tid * (TI / 32) + w] = bits[w];
    }
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[m][j] = 0.f;

    for (int kc = 0; kc < nkc; ++kc) {
      const int q = tt * nkc + kc;
      cp_async_wait<1>();
      __syncthreads();
      issue_stage(q + 2);
      const float *Ist = Is + (q % NSTAGE) * TI * LDI;
      const float *Ust = Us + kc * KC;
      const int kw = min(KC, d - kc * KC);
      for (int k = 0; k < kw; k += 4) {
        float4 u4[4], i4[8];
#pragma unroll
        for (int m = 0; m < 4; ++m) u4[m] = *(const float4 *)(Ust + (ty + 16 * m) * ldu + k);
#pragma unroll
        for (int j = 0; j < 8; ++j) i4[j] = *(const float4 *)(Ist + (tx + 16 * j) * LDI + k);
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float c = acc[m][j];
            c = fmaf(u4[m].x, i4[j].x, c);
            c = fmaf(u4[m].y, i4[j].y, c);
            c = fmaf(u4[m].z, i4[j].z, c);
            c = fmaf(u4[m].w, i4[j].w, c);
            acc[m][j] = c;
          }
      }
    }

    // ---- epilogue: transform, mask, threshold filter
    uint32_t pend = 0;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int ul = ty + 16 * m;
      const bool urow = (u0 + ul) < a.n;
      const float thr_s = list_s[ul * K + K - 1];
      const int thr_i = list_i[ul * K + K - 1];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int il = tx + 16 * j, li = tile_base + il, gid = a.item_base + li;
        float s = score_transform(acc[m][j], a.transform, a.max_rating);
        const bool masked = (gid == 0) || ((hmask[ul * (TI / 32) + (il >> 5)] >> (il & 31)) & 1u);
        if (masked) s = -INFINITY;
        acc[m][j] = s;
        if (urow && li < a.n_items_local && better(s, gid, thr_s, thr_i)) pend |= 1u << (m * 8 + j);
      }
    }
    // ---- push survivors into the row queues; merge; repeat while some queue overflowed
    while (true) {
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int ul = ty + 16 * m;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t bit = 1u << (m * 8 + j);
          if (pend & bit) {
            const int gid = a.item_base + tile_base + tx + 16 * j;
            const float s = acc[m][j];
            if (!better(s, gid, list_s[ul * K + K - 1], list_i[ul * K + K - 1])) {
              pend &= ~bit;  // the threshold has risen past it
            } else {
              const int slot = atomicAdd(&q_cnt[ul], 1);
              if (slot < QCAP) {
                q_s[ul * QCAP + slot] = s;
                q_i[ul * QCAP + slot] = gid;
                pend &= ~bit;
              }
            }
          }
        }
      }
      const int any_pending = __syncthreads_or(pend != 0);
      if (tid < TU) {
        const int c = min(q_cnt[tid], QCAP);
        for (int e = 0; e < c; ++e) list_insert(list_s + tid * K, list_i + tid * K, K, q_s[tid * QCAP + e], q_i[tid * QCAP + e]);
        q_cnt[tid] = 0;
      }
      __syncthreads();
      if (!any_pending) break;
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  // ---- emit the K-lists
  for (int i = tid; i < TU * K; i += 256) {
    const int ul = i / K, e = i % K, r = u0 + ul;
    if (r < a.n) {
      const size_t o = ((size_t)split * a.n + r) * K + e;
      a.out_id[o] = list_i[i];
      a.out_score[o] = list_s[i];
    }
  }
}

// K-way merge of P sorted per-shard (or per-split) lists per user under the same total order.
__global__ void __launch_bounds__(128)
    k_topk_merge(const int32_t *__restrict__ ids_in, const float *__restrict__ sc_in, int P, int n, int K,
                 int32_t *__restrict__ ids_out, float *__restrict__ sc_out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int ptr[kMaxSplits];
  for (int p = 0; p < P; ++p) ptr[p] = 0;
  for (int e = 0; e < K; ++e) {
    int best = -1, bid = 0x7fffffff;
    float bs = -INFINITY;
    for (int p = 0; p < P; ++p) {
      if (ptr[p] >= K) continue;
      const size_t o = ((size_t)p * n + r) * K + ptr[p];
      const float s = sc_in[o];
      const int id = ids_in[o];
      if (best < 0 || better(s, id, bs, bid)) {
        best = p; bs = s; bid = id;
      }
    }
    ids_out[(size_t)r * K + e] = bid;
    sc_out[(size_t)r * K + e] = bs;
    if (best >= 0) ++ptr[best];
  }
}

// Rows the tensor-core scorer left short (a user with fewer than K unmasked items): torch.topk continues into the -inf
// entries (trainer.py:435-438 masks by writing -inf, collector.py:147 takes topk of the whole row), and the canonical
// order among them is ascending item id.  The scorer never ranks masked columns, so they are appended here: the [PAD]
// item (global id 0) and the user's history inside this shard, which the tensor-core path requires sorted.
__global__ void __launch_bounds__(128)
    k_fill_masked(int32_t *__restrict__ ids, float *__restrict__ sc, int n, int K, const int64_t *__restrict__ hist_off,
                  const int32_t *__restrict__ hist_items, int item_base, int n_items_local) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int32_t *row = ids + (size_t)r * K;
  float *rs = sc + (size_t)r * K;
  if (row[K - 1] != 0x7fffffff) return;
  int f = K - 1;
  while (f > 0 && row[f - 1] == 0x7fffffff) --f;
  int prev = -1;
  if (item_base == 0) {
    row[f] = 0;
    rs[f] = -INFINITY;
    ++f;
    prev = 0;
  }
  if (!hist_items) return;
  long long lo = hist_off[r], hi = hist_off[r + 1];
  const long long end = hi;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (hist_items[mid] < item_base) lo = mid + 1; else hi = mid;
  }
  for (long long p = lo; p < end && f < K; ++p) {
    const int it = hist_items[p];
    if (it >= item_base + n_items_local) break;
    if (it == prev) continue;
    row[f] = it;
    rs[f] = -INFINITY;
    ++f;
    prev = it;
  }
}

// collector.py:147-153: rec.topk = [1[topk_id in positives(u)] ... | |positives(u)|]
__global__ void __launch_bounds__(128)
    k_hits(const int32_t *__restrict__ topk_id, int n, int K, const int64_t *__restrict__ pos_off,
           const int32_t *__restrict__ pos_items, int32_t *__restrict__ rec_topk) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * (K + 1)) return;
  const int r = idx / (K + 1), e = idx % (K + 1);
  const long long p0 = pos_off[r], p1 = pos_off[r + 1];
  if (e == K) {
    rec_topk[idx] = (int32_t)(p1 - p0);
    return;
  }
  const int id = topk_id[(size_t)r * K + e];
  long long lo = p0, hi = p1;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (pos_items[mid] < id) lo = mid + 1; else hi = mid;
  }
  rec_topk[idx] = (lo < p1 && pos_items[lo] == id) ? 1 : 0;
}

static size_t exact_smem_bytes(int d, int K) {
  size_t f = (size_t)TU * (d + 4) + (size_t)NSTAGE * TI * LDI + (size_t)TU * K * 2 + (size_t)TU * QCAP * 2 + TU +
             (size_t)TU * (TI / 32);
  return f * 4 + (size_t)TU * 8 + 16;
}

static int pick_splits(int n, int n_items_local) {
  const int utiles = (n + TU - 1) / TU, itiles = (n_items_local + TI - 1) / TI;
  int s = (2 * kSMs + utiles - 1) / utiles;
  if (s > kMaxSplits) s = kMaxSplits;
  if (s > itiles) s = itiles;
  if (s < 1) s = 1;
  return s;
}

// fullsort_tc.cu
size_t tc_plane_bytes(int n, int n_items_local, int d);
int tc_pick_splits(int n, int n_items_local);
int tc_launch(const fr_fullsort *a, void *planes, int splits, int32_t *out_id, float *out_sc, cudaStream_t st);

}  // namespace fr

extern "C" {

size_t fr_fullsort_workspace_bytes(int32_t n, int32_t K, int32_t n_items_local, int32_t d, int32_t score_mode) {
  if (score_mode == FR_SCORE_TC_3XTF32) {
    const int s = fr::tc_pick_splits(n, n_items_local);
    return fr::tc_plane_bytes(n, n_items_local, d) + (s > 1 ? (size_t)s * n * K * 8 + 512 : 256);
  }
  const int s = fr::pick_splits(n, n_items_local);
  return s > 1 ? (size_t)s * n * K * 8 + 512 : 256;
}

int fr_fullsort_topk(const fr_fullsort *a, void *stream) {
  FR_REQUIRE(a && a->U && a->I_shard && a->users && a->hist_off && a->topk_id && a->topk_score,
             "fr_fullsort_topk: null pointer");
  FR_REQUIRE(a->n >= 1 && a->n_items_local >= 1, "fr_fullsort_topk: empty input");
  FR_REQUIRE(a->K >= 1 && a->K <= fr::kMaxK, "fr_fullsort_topk: K=%d out of [1,%d]", a->K, fr::kMaxK);
  FR_REQUIRE(a->d >= 4 && a->d % 4 == 0, "fr_fullsort_topk: d=%d must be a multiple of 4", a->d);
  if (a->score_mode == FR_SCORE_TC_3XTF32) {
    FR_REQUIRE(a->d % 32 == 0 && a->d <= 128, "fr_fullsort_topk: tensor-core scorer needs d in {32,64,96,128} (d=%d)",
               a->d);
    const int splits = fr::tc_pick_splits(a->n, a->n_items_local);
    const size_t planes = fr::tc_plane_bytes(a->n, a->n_items_local, a->d);
    const size_t need = planes + (splits > 1 ? (size_t)splits * a->n * a->K * 8 + 512 : 256);
    if (!a->workspace || a->workspace_bytes < need) {
      fr::set_error("fr_fullsort_topk: workspace too small (%zu < %zu)", a->workspace_bytes, need);
      return FR_ERR_WORKSPACE;
    }
    int32_t *pid = a->topk_id;
    float *psc = a->topk_score;
    if (splits > 1) {
      pid = (int32_t *)((char *)a->workspace + planes);
      psc = (float *)((char *)a->workspace + planes + (((size_t)splits * a->n * a->K * 4 + 255) & ~(size_t)255));
    }
    int rc = fr::tc_launch(a, a->workspace, splits, pid, psc, (cudaStream_t)stream);
    if (rc) return rc;
    if (splits > 1) {
      FR_LAUNCH(fr::k_topk_merge, (a->n + 127) / 128, 128, 0, stream, pid, psc, splits, a->n, a->K, a->topk_id,
                a->topk_score);
    }
    FR_LAUNCH(fr::k_fill_masked, (a->n + 127) / 128, 128, 0, stream, a->topk_id, a->topk_score, a->n, a->K, a->hist_off,
              a->hist_items, a->item_base, a->n_items_local);
    FR_LAUNCH_CHECK();
    return FR_OK;
  }
  if (a->score_mode != FR_SCORE_EXACT_FP32) {
    fr::set_error("fr_fullsort_topk: unknown score_mode %d", a->score_mode);
    return FR_ERR_INVALID;
  }
  if (a->d > fr::kMaxDExact) {
    fr::set_error("fr_fullsort_topk: exact scorer supports d <= %d", fr::kMaxDExact);
    return FR_ERR_UNSUPPORTED;
  }
  const int splits = fr::pick_splits(a->n, a->n_items_local);
  const int itiles = (a->n_items_local + fr::TI - 1) / fr::TI;
  fr::EvalArgs e{a->U, a->I_shard, a->users, a->hist_off, a->hist_items, a->n, a->d, a->n_items_local, a->item_base,
                 a->K, a->transform, a->max_rating, (itiles + splits - 1) / splits, splits, a->topk_id, a->topk_score};
  int32_t *part_id = nullptr;
  float *part_sc = nullptr;
  if (splits > 1) {
    const size_t need = (size_t)splits * a->n * a->K * 8 + 512;
    if (!a->workspace || a->workspace_bytes < need) {
      fr::set_error("fr_fullsort_topk: workspace too small (%zu < %zu)", a->workspace_bytes, need);
      return FR_ERR_WORKSPACE;
    }
    part_id = (int32_t *)a->workspace;
    part_sc = (float *)((char *)a->workspace + (((size_t)splits * a->n * a->K * 4 + 255) & ~(size_t)255));
    e.out_id = part_id;
    e.out_score = part_sc;
  }
  const size_t smem = fr::exact_smem_bytes(a->d, a->K);
  FR_CUDA_OK(cudaFuncSetAttribute(fr::k_fullsort_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((a->n + fr::TU - 1) / fr::TU, splits);
  FR_LAUNCH(fr::k_fullsort_exact, grid, 256, smem, stream, e);
  if (splits > 1) {
    FR_LAUNCH(fr::k_topk_merge, (a->n + 127) / 128, 128, 0, stream, part_id, part_sc, splits, a->n, a->K, a->topk_id,
              a->topk_score);
  }
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_topk_merge(const int32_t *ids_in, const float *scores_in, int32_t P, int32_t n, int32_t K, int32_t *ids_out,
                  float *scores_out, void *stream) {
  FR_REQUIRE(ids_in && scores_in && ids_out && scores_out, "fr_topk_merge: null pointer");
  FR_REQUIRE(P >= 1 && P <= fr::kMaxSplits && n >= 1 && K >= 1, "fr_topk_merge: bad sizes P=%d n=%d K=%d", P, n, K);
  FR_LAUNCH(fr::k_topk_merge, (n + 127) / 128, 128, 0, stream, ids_in, scores_in, P, n, K, ids_out, scores_out);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

int fr_hits(const int32_t *topk_id, int32_t n, int32_t K, const int64_t *pos_off, const int32_t *pos_items,
            int32_t *rec_topk, void *stream) {
  FR_REQUIRE(topk_id && pos_off && pos_items && rec_topk && n >= 1 && K >= 1, "fr_hits: bad argument");
  const int64_t tot = (int64_t)n * (K + 1);
  FR_LAUNCH(fr::k_hits, (int)((tot + 127) / 128), 128, 0, stream, topk_id, n, K, pos_off, pos_items, rec_topk);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Sampled-negative ("uni100") ranking evaluation: top-K over each user's CANDIDATE list.
// Reference being replaced: trainer.py:441-456 (_neg_sample_batch_eval: scatter the candidate scores into a
// [users, n_items] matrix of -inf) + collector.py:141-153 (topk, hit bits, number of positives).  The dense matrix is
// never built: one warp owns a user and extracts the K best candidates in the canonical total order (score desc, item
// id asc) by K rounds of "best entry strictly after the previous pick" -- duplicates of a candidate (same pair, same
// score) collapse by construction, like the scatter.  A user with fewer than K distinct candidates gets the lowest
// non-candidate item ids as filler (score -inf), which is what the canonical order gives on the dense -inf row.
namespace fr {

__device__ __forceinline__ bool cand_after(float s, int id, float ps, int pid) {   // (s,id) strictly after (ps,pid)
  return s < ps || (s == ps && id > pid);
}
__device__ __forceinline__ bool cand_better(float s, int id, float s2, int id2) { return s > s2 || (s == s2 && id < id2); }

__global__ void __launch_bounds__(256)
    k_sampled_topk(const int64_t *__restrict__ cand_off, const int32_t *__restrict__ cand_items,
                   const float *__restrict__ cand_scores, const int32_t *__restrict__ n_pos_of_user, int n, int K,
                   int n_items, int32_t *__restrict__ topk_id, float *__restrict__ topk_score,
                   int32_t *__restrict__ rec_topk) {
  const int lane = threadIdx.x & 31;
  const int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (u >= n) return;
  const int64_t c0 = cand_off[u], c1 = cand_off[u + 1];
  const int np = n_pos_of_user[u];
  float ps = INFINITY;
  int pid = -1;
  bool exhausted = false;
  int fill = 0;   // next filler id to try
  for (int r = 0; r < K; ++r) {
    float bs = -INFINITY;
    int bid = 0x7fffffff;
    if (!exhausted) {
      for (int64_t c = c0 + lane; c < c1; c += 32) {
        const float s = cand_scores[c];
        const int id = cand_items[c];
        if (s == s && cand_after(s, id, ps, pid) && cand_better(s, id, bs, bid)) { bs = s; bid = id; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float s2 = __shfl_xor_sync(0xffffffffu, bs, o);
        const int id2 = __shfl_xor_sync(0xffffffffu, bid, o);
        if (cand_better(s2, id2, bs, bid)) { bs = s2; bid = id2; }
      }
      if (bid == 0x7fffffff) exhausted = true;
    }
    int hit = 0;
    if (!exhausted) {
      ps = bs; pid = bid;
      for (int64_t c = c0 + lane; c < c0 + np; c += 32) hit |= cand_items[c] == bid;
      hit = __any_sync(0xffffffffu, hit);
    } else {   // filler: lowest item id that is not a candidate
      while (true) {
        int is_cand = 0;
        for (int64_t c = c0 + lane; c < c1; c += 32) is_cand |= cand_items[c] == fill;
        if (!__any_sync(0xffffffffu, is_cand) || fill >= n_items) break;
        ++fill;
      }
      bid = fill++;
      bs = -INFINITY;
    }
    if (lane == 0) {
      topk_id[(size_t)u * K + r] = bid;
      topk_score[(size_t)u * K + r] = bs;
      rec_topk[(size_t)u * (K + 1) + r] = hit ? 1 : 0;
    }
  }
  // number of DISTINCT positives (pos_matrix.sum(dim=1), collector.py:150)
  int distinct = 0;
  for (int a = lane; a < np; a += 32) {
    const int id = cand_items[c0 + a];
    bool first = true;
    for (int b = 0; b < a; ++b) first &= cand_items[c0 + b] != id;
    distinct += first;
  }
  distinct = warp_sum(distinct);
  if (lane == 0) rec_topk[(size_t)u * (K + 1) + K] = distinct;
}

}  // namespace fr

extern "C" int fr_sampled_topk(const int64_t *cand_off, const int32_t *cand_items, const float *cand_scores,
                               const int32_t *n_pos_of_user, int32_t n, int32_t K, int32_t n_items, int32_t *topk_id,
                               float *topk_score, int32_t *rec_topk, void *stream) {
  FR_REQUIRE(cand_off && cand_items && cand_scores && n_pos_of_user && topk_id && topk_score && rec_topk && n >= 1 &&
                 K >= 1 && n_items >= K,
             "fr_sampled_topk: bad argument");
  FR_LAUNCH(fr::k_sampled_topk, (n + 7) / 8, 256, 0, stream, cand_off, cand_items, cand_scores, n_pos_of_user, n, K,
            n_items, topk_id, topk_score, rec_topk);
  FR_LAUNCH_CHECK();
  return FR_OK;
}
