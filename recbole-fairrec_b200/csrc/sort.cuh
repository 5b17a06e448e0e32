// Internal interface of sort.cu: stable radix sort of (key,value) pairs and segment discovery over
// sorted keys.  Host functions enqueue kernels on `stream`; no allocation, no synchronisation.
#pragma once
#include "common.cuh"

namespace fr {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;  // 2048 keys per CTA
constexpr int kRadix = 256;

static inline int64_t sort_num_blocks(int64_t n) { return (n + kSortTile - 1) / kSortTile; }
static inline int bits_for(uint32_t max_key_exclusive) {
  int b = 1;
  while (b < 32 && (1ull << b) < (unsigned long long)max_key_exclusive) ++b;
  return b;
}

struct SortScratch {
  uint32_t *tmp_keys;    // [n]
  uint32_t *tmp_vals;    // [n]
  uint32_t *block_hist;  // [nblk * 256]
};
size_t sort_scratch_bytes(int64_t n);
SortScratch carve_sort_scratch(Carver &c, int64_t n);

// vals_in == nullptr means identity values (0..n-1).  Result always lands in keys_out/vals_out.
// n is the host-side (upper bound) count; if n_dev != nullptr the kernels use *n_dev (<= n) instead, so a
// captured CUDA graph can be replayed on batches of different size.
void sort_pairs(const uint32_t *keys_in, const uint32_t *vals_in, uint32_t *keys_out, uint32_t *vals_out, int64_t n,
                const int32_t *n_dev, int key_bits, const SortScratch &s, cudaStream_t stream);

// Segments of equal keys in a sorted key array.
//   seg_id[pos]      segment index of every sorted position
//   seg_off[s]       first position of segment s; seg_off[count] = n
//   count_out[0]     number of segments
//   row_tab[key]     = {*stamp_ptr, s}   (optional, 8 bytes per row: "is this row touched in this batch, where")
//   entry_seg[ord[pos]] = s          (optional; ord == nullptr means identity order)
struct SegScratch {
  uint32_t *blk_cnt;  // [nblk + 1]
};
size_t seg_scratch_bytes(int64_t n);
SegScratch carve_seg_scratch(Carver &c, int64_t n);
void build_segments(const uint32_t *sorted_keys, const uint32_t *ord, int64_t n, const int32_t *n_dev, int32_t *seg_id,
                    int32_t *seg_off, int32_t *count_out, uint2 *row_tab, const uint32_t *stamp_ptr,
                    int32_t *entry_seg, const SegScratch &s, cudaStream_t stream);

}  // namespace fr
