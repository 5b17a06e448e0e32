/* CPU restatement (plain C) of the reference's full-sort scoring path, used where the numpy
 * restatement would be too slow or not bit-defined.
 *
 * TEST INFRASTRUCTURE ONLY: linked/loaded only by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs.  The product library never links it.
 *
 * Parity status: PINNED through oracle/fullsort_oracle.py, which tests/test_oracle_golden.py checks
 * against tests/golden/focf_eval_*.npz (produced by the unmodified reference).
 *
 * Restates (paths relative to /root/reference):
 *   recbole/model/fair_recommender/focf.py:171-178   s = clamp(U[u].I[i], 0, max_rating) / max_rating
 *   recbole/trainer/trainer.py:435-438               s[:,0] = -inf ; s[history] = -inf
 *   recbole/evaluator/collector.py:143               top-K indices (canonical order: score desc, id asc)
 * The dot product is defined as the k-ascending chain acc = fmaf(u[k], i[k], acc), acc0 = 0, which is
 * what the CUDA "exact" scoring mode computes, so both sides are bit-identical.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* scores[n, Ni] for users[n]; transform: 0 none, 1 clamp(0,max)/max */
void oracle_full_sort_scores(const float *U, const float *I, const int64_t *users, int64_t n, int64_t Ni,
                             int64_t d, int transform, float max_rating, float *scores) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n; ++r) {
    const float *u = U + users[r] * d;
    for (int64_t i = 0; i < Ni; ++i) {
      const float *v = I + i * d;
      float acc = 0.0f;
      for (int64_t k = 0; k < d; ++k) acc = fmaf(u[k], v[k], acc);
      if (transform == 1) {
        acc = acc < 0.0f ? 0.0f : (acc > max_rating ? max_rating : acc);
        acc = acc / max_rating;
      }
      scores[r * Ni + i] = acc;
    }
  }
}

/* in-place mask: column 0 (the [PAD] item) and every history item -> -inf */
void oracle_mask_history(float *scores, int64_t n, int64_t Ni, const int64_t *hist_off,
                         const int64_t *hist_items) {
  for (int64_t r = 0; r < n; ++r) {
    scores[r * Ni] = -INFINITY;
    for (int64_t p = hist_off[r]; p < hist_off[r + 1]; ++p) scores[r * Ni + hist_items[p]] = -INFINITY;
  }
}

/* canonical top-K per row: (score desc, id asc) total order; O(Ni*K) insertion */
void oracle_topk(const float *scores, int64_t n, int64_t Ni, int64_t K, int64_t *ids, float *vals) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n; ++r) {
    const float *s = scores + r * Ni;
    int64_t *id = ids + r * K;
    float *va = vals + r * K;
    int64_t cnt = 0;
    for (int64_t i = 0; i < Ni; ++i) {
      float x = s[i];
      if (cnt == K && !(x > va[K - 1])) continue; /* equal score, larger id: loses */
      int64_t p = cnt < K ? cnt : K - 1;
      while (p > 0 && x > va[p - 1]) {
        va[p] = va[p - 1];
        id[p] = id[p - 1];
        --p;
      }
      va[p] = x;
      id[p] = i;
      if (cnt < K) ++cnt;
    }
  }
}

/* Fused scoring + mask + top-K for a block of users without materialising [n,Ni] for the caller:
 * used as the multi-threaded CPU baseline ("port") in bench.py.  Same arithmetic as above. */
void oracle_full_sort_topk(const float *U, const float *I, const int64_t *users, int64_t n, int64_t Ni,
                           int64_t d, int transform, float max_rating, const int64_t *hist_off,
                           const int64_t *hist_items, int64_t K, int64_t *ids, float *vals) {
#pragma omp parallel
  {
    float *row = (float *)malloc(sizeof(float) * (size_t)Ni);
#pragma omp for schedule(dynamic, 16)
    for (int64_t r = 0; r < n; ++r) {
      int64_t one = users[r];
      oracle_full_sort_scores(U, I, &one, 1, Ni, d, transform, max_rating, row);
      row[0] = -INFINITY;
      for (int64_t p = hist_off[r]; p < hist_off[r + 1]; ++p) row[hist_items[p]] = -INFINITY;
      oracle_topk(row, 1, Ni, K, ids + r * K, vals + r * K);
    }
    free(row);
  }
}

/* pair scores: pred[b] = chain-fma dot (transform as above); used for rec.positive_score */
void oracle_pair_scores(const float *U, const float *I, const int64_t *uid, const int64_t *iid, int64_t B,
                        int64_t d, int transform, float max_rating, float *out) {
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < B; ++b) {
    const float *u = U + uid[b] * d, *v = I + iid[b] * d;
    float acc = 0.0f;
    for (int64_t k = 0; k < d; ++k) acc = fmaf(u[k], v[k], acc);
    if (transform == 1) {
      acc = acc < 0.0f ? 0.0f : (acc > max_rating ? max_rating : acc);
      acc = acc / max_rating;
    }
    out[b] = acc;
  }
}
