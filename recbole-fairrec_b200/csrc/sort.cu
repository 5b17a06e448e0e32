// Stable LSD radix sort of (u32 key, u32 value) pairs + segment discovery, hand-written for sm_100a.
//
// Replaces, on the hot path: torch.unique(return_inverse=True) (reference focf.py:77-78, sort based),
// the implicit index sort of nn.Embedding's dense backward, and np.unique in the fairness metrics
// (metrics.py:941-945, 1327-1328).  Integer work only: results are bit-exact and deterministic.
//
// Layout: one CTA owns a tile of 2048 consecutive keys.  A pass is
//   hist    : per-CTA digit histogram (shared-memory integer atomics)
//   scan    : warp per digit: exclusive prefix of the digit-major [256][nblk] counts over the blocks + digit totals
//   scatter : each warp ranks its 256 keys with __match_any_sync in 8 rounds of 32 consecutive keys
//             (stable), then writes (key,value) to offset[digit] + rank.
// When the whole input fits one tile the three kernels collapse into one (k_radix_single).
#include <cstdarg>
#include <atomic>
#include <map>
#include <string>
#include <vector>
#include <cstring>

#include "sort.cuh"

namespace fr {

// ------------------------------------------------------------------ error + launch bookkeeping
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// ------------------------------------------------------------------ per-kernel event profiler
struct ProfRec {
  const char *name;
  cudaEvent_t a, b;
};
static bool g_prof = false;
static std::vector<ProfRec> g_recs;
bool prof_on() { return g_prof; }
void prof_begin(const char *kernel, cudaStream_t stream) {
  ProfRec r{kernel, nullptr, nullptr};
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, stream);
  g_recs.push_back(r);
}
void prof_end(cudaStream_t stream) { cudaEventRecord(g_recs.back().b, stream); }

// ------------------------------------------------------------------ radix sort kernels
__global__ void __launch_bounds__(kSortThreads) k_radix_hist(const uint32_t *__restrict__ keys, int64_t n, const int32_t *__restrict__ n_dev,
                                                            int shift, uint32_t *__restrict__ block_hist) {
  if (n_dev) n = *n_dev;
  __shared__ uint32_t h[kRadix];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    int64_t p = base + i * kSortThreads + threadIdx.x;
    if (p < n) atomicAdd(&h[(keys[p] >> shift) & (kRadix - 1)], 1u);
  }
  __syncthreads();
  block_hist[(int64_t)threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];   // digit-major: [256][nblk]
}

// Warp per digit (32 CTAs of 8 warps): the digit's counts hist[d][0..nblk) become their exclusive prefix over the blocks
// and tot[d] their sum; the scatter kernel adds the exclusive scan of tot[] over the digits itself (256 values, per CTA).
// 32 consecutive blocks per warp scan (coalesced), 16 scans' loads in flight.  (A single-CTA scan of the whole table
// is bound by what ONE SM pulls from L2: 33-90 us per pass at 2^20 keys in three variants.)
constexpr int kScanWarps = 8;
__global__ void __launch_bounds__(kScanWarps * 32) k_radix_scan(uint32_t *__restrict__ hist, int64_t nblk,
                                                               uint32_t *__restrict__ tot) {
  const int lane = threadIdx.x & 31, dgt = blockIdx.x * kScanWarps + (threadIdx.x >> 5);
  uint32_t *h = hist + (int64_t)dgt * nblk;
  uint32_t carry = 0;
  for (int64_t c0 = 0; c0 < nblk; c0 += 32 * 16) {
    uint32_t v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int64_t b = c0 + 32 * i + lane;
      v[i] = b < nblk ? h[b] : 0u;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      uint32_t inc = v[i];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      const int64_t b = c0 + 32 * i + lane;
      if (b < nblk) h[b] = carry + inc - v[i];
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
  }
  if (lane == 0) tot[dgt] = carry;
}

// Stable ranking + scatter of one 2048-key tile.  kSingle: the tile is the whole array, compute the digit
// bases in-kernel (no hist/scan kernels needed).
template <bool kSingle>
__global__ void __launch_bounds__(kSortThreads)
    k_radix_scatter(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                    uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int64_t n,
                    const int32_t *__restrict__ n_dev, int shift, const uint32_t *__restrict__ offsets,
                    const uint32_t *__restrict__ digit_tot) {
  if (n_dev) n = *n_dev;
  constexpr int kWarps = kSortThreads / 32;
  __shared__ uint32_t cnt[kWarps][kRadix];
  __shared__ uint32_t digit_base[kRadix];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kWarps * kRadix; i += kSortThreads) (&cnt[0][0])[i] = 0;
  if (!kSingle) {   // exclusive scan of the 256 digit totals (thread d = digit d): the global base of every digit
    __shared__ uint32_t wtot[kWarps];
    const uint32_t t = digit_tot[threadIdx.x];
    uint32_t inc = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) wtot[w] = inc;
    __syncthreads();
    uint32_t base = 0;
#pragma unroll
    for (int ww = 0; ww < kWarps; ++ww)
      if (ww < w) base += wtot[ww];
    digit_base[threadIdx.x] = base + inc - t;
  }
  __syncthreads();

  const int64_t wbase = (int64_t)blockIdx.x * kSortTile + (int64_t)w * (kSortTile / kWarps);
  uint32_t key[kSortItems], val[kSortItems], rnk[kSortItems];
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const int64_t p = wbase + r * 32 + lane;
    const bool valid = p < n;
    key[r] = valid ? keys_in[p] : 0u;
    val[r] = valid ? (vals_in ? vals_in[p] : (uint32_t)p) : 0u;
    const uint32_t dig = valid ? ((key[r] >> shift) & (kRadix - 1)) : 0xffffffffu;
    const unsigned peers = __match_any_sync(0xffffffffu, dig);
    const uint32_t before = valid ? cnt[w][dig] : 0u;
    __syncwarp();
    if (valid && lane == (__ffs(peers) - 1)) cnt[w][dig] = before + __popc(peers);
    __syncwarp();
    rnk[r] = before + __popc(peers & ((1u << lane) - 1u));
  }
  __syncthreads();
  {  // thread d: exclusive prefix over the warps of this CTA, on top of the CTA's global offset for digit d
    const int dgt = threadIdx.x;
    uint32_t run = 0;
    if (!kSingle) run = offsets[(int64_t)dgt * gridDim.x + blockIdx.x] + digit_base[dgt];
    uint32_t tot = 0;
#pragma unroll
    for (int ww = 0; ww < kWarps; ++ww) {
      uint32_t t = cnt[ww][dgt];
      cnt[ww][dgt] = run + tot;
      tot += t;
    }
    if (kSingle) digit_base[dgt] = tot;
  }
  __syncthreads();
  if (kSingle) {
    if (threadIdx.x == 0) {
      uint32_t run = 0;
      for (int i = 0; i < kRadix; ++i) {
        uint32_t t = digit_base[i];
        digit_base[i] = run;
        run += t;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const int64_t p = wbase + r * 32 + lane;
    if (p < n) {
      const uint32_t dig = (key[r] >> shift) & (kRadix - 1);
      const uint32_t pos = cnt[w][dig] + rnk[r] + (kSingle ? digit_base[dig] : 0u);
      keys_out[pos] = key[r];
      vals_out[pos] = val[r];
    }
  }
}

__global__ void k_copy_pairs(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                             uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int64_t n) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    keys_out[p] = keys_in[p];
    vals_out[p] = vals_in ? vals_in[p] : (uint32_t)p;
  }
}

size_t sort_scratch_bytes(int64_t n) {
  Carver c(nullptr, 0);
  carve_sort_scratch(c, n);
  return c.off;
}
SortScratch carve_sort_scratch(Carver &c, int64_t n) {
  SortScratch s;
  s.tmp_keys = c.take<uint32_t>((size_t)n);
  s.tmp_vals = c.take<uint32_t>((size_t)n);
  s.block_hist = c.take<uint32_t>((size_t)(sort_num_blocks(n) + 1) * kRadix);   // + the 256 digit totals
  return s;
}

void sort_pairs(const uint32_t *keys_in, const uint32_t *vals_in, uint32_t *keys_out, uint32_t *vals_out, int64_t n,
                const int32_t *n_dev, int key_bits, const SortScratch &s, cudaStream_t stream) {
  if (n <= 0) return;
  int passes = (key_bits + 7) / 8;
  if (passes < 1) passes = 1;
  const int64_t nblk = sort_num_blocks(n);
  const uint32_t *ki = keys_in, *vi = vals_in;
  for (int p = 0; p < passes; ++p) {
    // ping-pong so that the last pass writes (keys_out, vals_out)
    const bool to_out = ((passes - 1 - p) % 2) == 0;
    uint32_t *ko = to_out ? keys_out : s.tmp_keys;
    uint32_t *vo = to_out ? vals_out : s.tmp_vals;
    const int shift = 8 * p;
    if (nblk == 1) {
      FR_LAUNCH(k_radix_scatter<true>, 1, kSortThreads, 0, stream, ki, vi, ko, vo, n, n_dev, shift, nullptr, nullptr);
    } else {
      FR_LAUNCH(k_radix_hist, (int)nblk, kSortThreads, 0, stream, ki, n, n_dev, shift, s.block_hist);
      uint32_t *digit_tot = s.block_hist + nblk * kRadix;
      FR_LAUNCH(k_radix_scan, kRadix / kScanWarps, kScanWarps * 32, 0, stream, s.block_hist, nblk, digit_tot);
      FR_LAUNCH(k_radix_scatter<false>, (int)nblk, kSortThreads, 0, stream, ki, vi, ko, vo, n, n_dev, shift, s.block_hist,
                digit_tot);
    }
    ki = ko;
    vi = vo;
  }
}

// ------------------------------------------------------------------ segments over sorted keys
__device__ __forceinline__ int block_exclusive_scan_256(int v, int *total) {
  // 256 threads; returns the exclusive prefix of v, *total = block sum (valid in all threads)
  __shared__ int warp_tot[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int t = warp_tot[i];
    if (i < w) base += t;
    tot += t;
  }
  __syncthreads();
  *total = tot;
  return base + inc - v;
}

__device__ __forceinline__ bool is_head(const uint32_t *keys, int64_t p) { return p == 0 || keys[p] != keys[p - 1]; }

// thread t of the CTA owns the kSortItems consecutive positions base + t*8 .. +7
__global__ void __launch_bounds__(kSortThreads) k_seg_count(const uint32_t *__restrict__ keys, int64_t n,
                                                           const int32_t *__restrict__ n_dev,
                                                           uint32_t *__restrict__ blk_cnt) {
  if (n_dev) n = *n_dev;
  const int64_t p0 = (int64_t)blockIdx.x * kSortTile + (int64_t)threadIdx.x * kSortItems;
  int c = 0;
#pragma unroll
  for (int i = 0; i < kSortItems; ++i)
    if (p0 + i < n) c += is_head(keys, p0 + i);
  int tot;
  block_exclusive_scan_256(c, &tot);
  if (threadIdx.x == 0) blk_cnt[blockIdx.x] = (uint32_t)tot;
}

__global__ void __launch_bounds__(1024) k_scan_blocks(uint32_t *__restrict__ blk_cnt, int64_t nblk) {
  // single CTA exclusive scan of nblk counters (chunked by 1024), in place; blk_cnt[nblk] = total
  __shared__ uint32_t wt[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int64_t base = 0; base < nblk; base += 1024) {
    const int64_t i = base + threadIdx.x;
    uint32_t v = i < nblk ? blk_cnt[i] : 0u, inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wt[w] = inc;
    __syncthreads();
    uint32_t wbase = 0, tot = 0;
    for (int k = 0; k < 32; ++k) {
      uint32_t t = wt[k];
      if (k < w) wbase += t;
      tot += t;
    }
    const uint32_t c = carry;
    if (i < nblk) blk_cnt[i] = c + wbase + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) blk_cnt[nblk] = carry;
}

template <bool kSingle>
__global__ void __launch_bounds__(kSortThreads)
    k_seg_assign(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ ord, int64_t n,
                 const int32_t *__restrict__ n_dev, const uint32_t *__restrict__ blk_off,
                 int32_t *__restrict__ seg_id, int32_t *__restrict__ seg_off, int32_t *__restrict__ count_out,
                 uint2 *__restrict__ row_tab, const uint32_t *__restrict__ stamp_ptr,
                 int32_t *__restrict__ entry_seg) {
  if (n_dev) n = *n_dev;
  const uint32_t stamp = stamp_ptr ? *stamp_ptr : 0u;
  if (n <= 0) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      count_out[0] = 0;
      seg_off[0] = 0;
    }
    return;
  }
  const int64_t p0 = (int64_t)blockIdx.x * kSortTile + (int64_t)threadIdx.x * kSortItems;
  bool head[kSortItems];
  int c = 0;
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    head[i] = (p0 + i < n) && is_head(keys, p0 + i);
    c += head[i];
  }
  int tot;
  int run = block_exclusive_scan_256(c, &tot) + (kSingle ? 0 : (int)blk_off[blockIdx.x]);
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const int64_t p = p0 + i;
    if (p >= n) break;
    if (head[i]) {
      seg_off[run] = (int32_t)p;
      if (row_tab) row_tab[keys[p]] = make_uint2(stamp, (uint32_t)run);
      ++run;
    }
    const int s = run - 1;
    seg_id[p] = s;
    if (entry_seg) entry_seg[ord ? ord[p] : (uint32_t)p] = s;
    if (p == n - 1) {
      seg_off[s + 1] = (int32_t)n;
      count_out[0] = s + 1;
    }
  }
}

size_t seg_scratch_bytes(int64_t n) {
  Carver c(nullptr, 0);
  carve_seg_scratch(c, n);
  return c.off;
}
SegScratch carve_seg_scratch(Carver &c, int64_t n) {
  SegScratch s;
  s.blk_cnt = c.take<uint32_t>((size_t)sort_num_blocks(n) + 1);
  return s;
}

void build_segments(const uint32_t *sorted_keys, const uint32_t *ord, int64_t n, const int32_t *n_dev, int32_t *seg_id,
                    int32_t *seg_off, int32_t *count_out, uint2 *row_tab, const uint32_t *stamp_ptr,
                    int32_t *entry_seg, const SegScratch &s, cudaStream_t stream) {
  if (n <= 0) return;
  const int64_t nblk = sort_num_blocks(n);
  if (nblk == 1) {
    FR_LAUNCH(k_seg_assign<true>, 1, kSortThreads, 0, stream, sorted_keys, ord, n, n_dev, nullptr, seg_id,
              seg_off, count_out, row_tab, stamp_ptr, entry_seg);
  } else {
    FR_LAUNCH(k_seg_count, (int)nblk, kSortThreads, 0, stream, sorted_keys, n, n_dev, s.blk_cnt);
    FR_LAUNCH(k_scan_blocks, 1, 1024, 0, stream, s.blk_cnt, nblk);
    FR_LAUNCH(k_seg_assign<false>, (int)nblk, kSortThreads, 0, stream, sorted_keys, ord, n, n_dev, s.blk_cnt,
              seg_id, seg_off, count_out, row_tab, stamp_ptr, entry_seg);
  }
}

}  // namespace fr

// ------------------------------------------------------------------ C ABI
extern "C" {

int fr_abi_version(void) { return FR_ABI_VERSION; }
const char *fr_last_error(void) { return fr::g_err; }
uint64_t fr_launch_count(void) { return fr::g_launches.load(); }

void fr_profile_enable(int on) { fr::g_prof = on != 0; }

// Synchronises the device, folds the recorded launches by kernel name (template arguments stripped) and writes
// "name,count,total_ms\n" lines into buf; returns the number of distinct kernels, clears the records.
int fr_profile_report(char *buf, size_t buf_bytes) {
  cudaDeviceSynchronize();
  std::map<std::string, std::pair<long, double>> acc;
  for (auto &r : fr::g_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      std::string n(r.name);
      size_t p = n.find("fr::");
      if (p == 0) n = n.substr(4);
      auto &e = acc[n];
      e.first += 1;
      e.second += (double)ms;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  fr::g_recs.clear();
  std::string out;
  for (auto &kv : acc) {
    char line[256];
    snprintf(line, sizeof(line), "%s,%ld,%.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    out += line;
  }
  if (buf && buf_bytes) {
    strncpy(buf, out.c_str(), buf_bytes - 1);
    buf[buf_bytes - 1] = 0;
  }
  return (int)acc.size();
}

size_t fr_sort_pairs_workspace_bytes(int64_t n) { return fr::sort_scratch_bytes(n < 1 ? 1 : n); }

int fr_sort_pairs_u32(const uint32_t *keys_in, const uint32_t *vals_in, uint32_t *keys_out, uint32_t *vals_out,
                      int64_t n, int key_bits, void *workspace, size_t workspace_bytes, void *stream) {
  FR_REQUIRE(n >= 0 && n < (1ll << 31), "fr_sort_pairs_u32: n=%lld out of range", (long long)n);
  if (n == 0) return FR_OK;
  FR_REQUIRE(keys_in && keys_out && vals_out && workspace, "fr_sort_pairs_u32: null pointer");
  FR_REQUIRE(key_bits >= 1 && key_bits <= 32, "fr_sort_pairs_u32: key_bits=%d", key_bits);
  fr::Carver c(workspace, workspace_bytes);
  fr::SortScratch s = fr::carve_sort_scratch(c, n);
  if (!c.ok()) {
    fr::set_error("fr_sort_pairs_u32: workspace too small (%zu < %zu)", workspace_bytes, c.off);
    return FR_ERR_WORKSPACE;
  }
  fr::sort_pairs(keys_in, vals_in, keys_out, vals_out, n, nullptr, key_bits, s, (cudaStream_t)stream);
  FR_LAUNCH_CHECK();
  return FR_OK;
}

}  // extern "C"
