// (the issue loop is warp-uniform with one elected lane: under `if (threadIdx.x == 0)` the operands live in vector registers
// and every UTCHMMA pays an ELECT / R2UR.BROADCAST waterfall -- that, not the tensor pipe, was what the first version timed)
// Micro-benchmark: cycles per tcgen05.mma dispatch for the shapes the tensor-core scorer can use (one CTA per SM on all
// 148 SMs, one thread issuing back-to-back MMAs on a zeroed shared-memory ring, clock64 around the issue loop + the
// final commit).  Answers: is kind::tf32 M=128 N=128 with both operands in shared memory at the 64-cycle floor, and if
// not, does N=256 or A-in-TMEM get there?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I recbole-fairrec_b200/csrc -I include \
//        -o profiles/tools/_bin/mma_rate profiles/tools/mma_rate.cu -lcuda
#include <cstdio>
#include <cstdlib>

#include "tc_common.cuh"

using namespace fr;

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

// mode 0: tf32 SS; 1: tf32 TS (A in TMEM); 2: bf16 SS
__global__ void __launch_bounds__(128, 1) k_rate(int mode, int N, int iters, long long *out, int smem_bytes) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < smem_bytes / 4; i += blockDim.x) ((uint32_t *)sm)[i] = 0u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init(&bar2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = __shfl_sync(0xffffffffu, slot, 0);
  long long t0 = 0;
  if (threadIdx.x < 32) {
   t0 = clock64();
   if (elect_one()) {
    const uint32_t a0 = smem_u32(sm), b0 = smem_u32(sm + 16384);
    uint32_t idesc = umma_idesc_tf32(128, N);
    if (mode == 2) idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int i = 0; i < (mode < 3 ? iters : 0); ++i) {
      const uint32_t off = (i & 3) * 32;
      const uint32_t d = tb + (uint32_t)((i >> 6) & 1) * 0;   // one accumulator
      if (mode == 0) umma_tf32(d, umma_desc_sw128(a0 + off), umma_desc_sw128(b0 + off), idesc, 1u);
      else if (mode == 1) umma_tf32_ts(d, tb + 256 + (i & 3) * 8, umma_desc_sw128(b0 + off), idesc, 1u);
      else if (mode == 2) umma_bf16(d, umma_desc_sw128(a0 + off), umma_desc_sw128(b0 + off), idesc, 1u);
    }
    if (mode >= 3) {
      // the scorer's pattern: per tile 4 K-blocks x 4 k-steps x 3 products over resident A planes (2 x 4 x 16 KB) and a
      // 3-stage B ring (hi 16 KB + lo 16 KB per stage); mode 4 also commits per K-block; mode 5: B-reuse order; mode 6:
      // two accumulators alternating per MMA triple; mode 7: hi.hi only (one product)
      const uint32_t A = smem_u32(sm), B = A + 131072;
      int q = 0;
      for (int t = 0; t < iters / 48; ++t) {
        for (int kb = 0; kb < 4; ++kb, ++q) {
          const int st = q % 3;
          const uint32_t a_hi = A + kb * 16384, a_lo = A + 65536 + kb * 16384, b_hi = B + st * 32768, b_lo = b_hi + 16384;
          for (int k = 0; k < 4; ++k) {
            const uint32_t off = k * 32;
            const uint32_t dd = tb + ((mode == 6) ? (uint32_t)(k & 1) * 128 : (uint32_t)(t & 1) * 128);
            if (mode == 7) {
              umma_tf32(dd, umma_desc_sw128(a_hi + off), umma_desc_sw128(b_hi + off), idesc, 1u);
              umma_tf32(dd, umma_desc_sw128(a_hi + off), umma_desc_sw128(b_hi + off), idesc, 1u);
              umma_tf32(dd, umma_desc_sw128(a_hi + off), umma_desc_sw128(b_hi + off), idesc, 1u);
            } else if (mode == 5) {
              umma_tf32(dd, umma_desc_sw128(a_hi + off), umma_desc_sw128(b_hi + off), idesc, 1u);
              umma_tf32(dd, umma_desc_sw128(a_lo + off), umma_desc_sw128(b_hi + off), idesc, 1u);
              umma_tf32(dd, umma_desc_sw128(a_hi + off), umma_desc_sw128(b_lo + off), idesc, 1u);
            } else {
              umma_tf32(dd, umma_desc_sw128(a_hi + off), umma_desc_sw128(b_hi + off), idesc, 1u);
              umma_tf32(dd, umma_desc_sw128(a_hi + off), umma_desc_sw128(b_lo + off), idesc, 1u);
              umma_tf32(dd, umma_desc_sw128(a_lo + off), umma_desc_sw128(b_hi + off), idesc, 1u);
            }
          }
          if (mode == 4) umma_commit(&bar2);
        }
      }
    }
    umma_commit(&bar);
   }
   __syncwarp();
   mbar_wait(&bar, 0);
   const long long t1 = clock64();
   if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}

int main() {
  long long *out;
  cudaMallocManaged(&out, 148 * sizeof(long long));
  const int SM = 131072 + 98304;
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);
  const int iters = 19200;
  const char *names[8] = {"tf32 SS", "tf32 TS (A in TMEM)", "bf16 SS", "scorer pattern", "scorer pattern+commit", "B-reuse order", "2 accumulators", "one product x3"};
  for (int grid : {148})
    for (int mode = 0; mode < 8; ++mode)
      for (int N : {64, 128, 256}) {
        if (mode >= 3 && N != 128) continue;
        for (int rep = 0; rep < 2; ++rep) {
          k_rate<<<grid, 128, SM>>>(mode, N, iters, out, SM);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("%s N=%d: %s\n", names[mode], N, cudaGetErrorString(e));
            return 1;
          }
        }
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = out[i] > mx ? out[i] : mx;
        printf("grid %3d  %-20s M=128 N=%3d: %.1f cycles per MMA (floor %d)\n", grid, names[mode], N, (double)mx / iters, N / 2);
      }
  return 0;
}
