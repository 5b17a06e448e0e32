// Shared helpers for libfairrec_b200 (sm_100a).  Host-side error/launch bookkeeping and small
// device utilities.  No torch, no third-party headers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fairrec_b200.h"

namespace fr {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
// optional per-kernel CUDA-event timing (fr_profile_*): brackets every FR_LAUNCH with two events
bool prof_on();
void prof_begin(const char *kernel, cudaStream_t stream);
void prof_end(cudaStream_t stream);

#define FR_REQUIRE(cond, ...)          \
  do {                                 \
    if (!(cond)) {                     \
      fr::set_error(__VA_ARGS__);      \
      return FR_ERR_INVALID;           \
    }                                  \
  } while (0)

#define FR_CUDA_OK(expr)                                                              \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      fr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return FR_ERR_CUDA;                                                             \
    }                                                                                 \
  } while (0)

// launch + count; errors are collected by FR_LAUNCH_CHECK at the end of each API call
#define FR_LAUNCH(kernel, grid, block, smem, stream, ...)                    \
  do {                                                                       \
    const bool _p = fr::prof_on();                                           \
    if (_p) fr::prof_begin(#kernel, (cudaStream_t)(stream));                 \
    kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__); \
    if (_p) fr::prof_end((cudaStream_t)(stream));                            \
    fr::count_launch();                                                      \
  } while (0)

#define FR_LAUNCH_CHECK()                                                       \
  do {                                                                          \
    cudaError_t _e = cudaPeekAtLastError();                                     \
    if (_e != cudaSuccess) {                                                    \
      fr::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return FR_ERR_CUDA;                                                       \
    }                                                                           \
  } while (0)

// bump allocator over a caller-provided workspace (256-byte aligned slices)
struct Carver {
  char *base;
  size_t off, cap;
  __host__ Carver(void *p, size_t bytes) : base((char *)p), off(0), cap(bytes) {}
  template <typename T>
  __host__ T *take(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
    T *r = (T *)(base ? base + off : nullptr);
    off += bytes;
    return r;
  }
  __host__ bool ok() const { return base == nullptr || off <= cap; }
};

constexpr int kSMs = 148;  // B200

static inline int grid_for(int64_t work_items, int per_block, int max_blocks = kSMs * 16) {
  int64_t g = (work_items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (int)g;
}

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// order-preserving float <-> uint encoding (for integer min/max on floats)
__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
// streaming 128-bit loads/stores that do not pollute L1
__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(float4 *p, const float4 &v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w));
}

}  // namespace fr
