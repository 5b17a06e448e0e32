// tcgen05 / TMA / mbarrier PTX wrappers, the TF32 hi/lo plane split and the tensor-map helper shared by the tensor-core
// kernels (fullsort_tc.cu: full-sort scorer; linear_tc.cu: MLP layer forward).  sm_100a only.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace fr {

constexpr int TCM = 128;          // rows per CTA tile (UMMA M)
constexpr int TCN = 128;          // scorer: items per tile (UMMA N)
constexpr int TCKB = 32;          // k-columns per K-block: 32 fp32 = 128 bytes = one swizzle-128B row

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int x, int y, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}
// one lane of a converged warp (elect.sync): the code around it stays warp-uniform, so ptxas keeps shared-memory
// addresses, descriptors and TMEM addresses in uniform registers -- UTCHMMA / UTMALDG take their operands from uniform
// registers, and a value that lives in a vector register (anything under `if (lane == 0)`) costs an ELECT / R2UR.BROADCAST
// waterfall loop per instruction (measured: ~106 cycles per 64-cycle MMA in the scorer's issue loop)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, SWIZZLE_128B (cute::UMMA::SmemDescriptor): start address >> 4 in bits
// [0,14), leading byte offset (unused for swizzled K-major, 1) in [16,30), stride byte offset = 1024 B between 8-row
// groups in [32,46), version 1 in [46,48), layout type 2 (SWIZZLE_128B) in [61,64)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 @[4,6), a/b format TF32 = 2 @[7,10)/[10,13),
// a/b K-major = 0 @15/16, n_dim = N>>3 @[17,23), m_dim = M>>4 @[24,29)
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// A operand in tensor memory (lane = row, one 32-bit column per tf32 element), B in shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- plane split (once per evaluation)
// rows: gathered by `rows_idx` (eval users) or identity (items); hi = x rounded to TF32 (nearest), lo = x - hi (exact
// in fp32) rounded to TF32 as well, so that the tensor core's own truncation of the low 13 bits is a no-op
__device__ __forceinline__ float rn_tf32(float x) {
  const uint32_t u = __float_as_uint(x);
  if ((u & 0x7f800000u) == 0x7f800000u) return x;   // inf / nan
  return __uint_as_float((u + 0x1000u) & 0xffffe000u);
}
static __global__ void __launch_bounds__(256)
    k_split_planes(const float *__restrict__ src, const int32_t *__restrict__ rows_idx, int64_t n_rows, int d,
                   float *__restrict__ hi, float *__restrict__ lo) {
  const int dq = d >> 2;
  const int64_t nq = n_rows * dq;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = q / dq;
    const int c = (int)(q % dq);
    const int64_t sr = rows_idx ? (int64_t)rows_idx[r] : r;
    const float4 x = __ldg((const float4 *)(src + sr * d) + c);
    float4 h, l;
    h.x = rn_tf32(x.x); l.x = rn_tf32(x.x - h.x);
    h.y = rn_tf32(x.y); l.y = rn_tf32(x.y - h.y);
    h.z = rn_tf32(x.z); l.z = rn_tf32(x.z - h.z);
    h.w = rn_tf32(x.w); l.w = rn_tf32(x.w - h.w);
    *((float4 *)(hi + r * d) + c) = h;
    *((float4 *)(lo + r * d) + c) = l;
  }
}

// two tables in one launch (the scorer's users and items: one launch less in an evaluation pass)
static __global__ void __launch_bounds__(256)
    k_split_planes2(const float *__restrict__ src0, const int32_t *__restrict__ idx0, int64_t n0, float *__restrict__ hi0,
                    float *__restrict__ lo0, const float *__restrict__ src1, int64_t n1, float *__restrict__ hi1,
                    float *__restrict__ lo1, int d) {
  const int dq = d >> 2;
  const int64_t nq0 = n0 * dq, nq = nq0 + n1 * dq;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
    const bool first = q < nq0;
    const int64_t qq = first ? q : q - nq0;
    const int64_t r = qq / dq;
    const int c = (int)(qq % dq);
    const int64_t sr = (first && idx0) ? (int64_t)idx0[r] : r;
    const float4 x = __ldg((const float4 *)((first ? src0 : src1) + sr * d) + c);
    float4 h, l;
    h.x = rn_tf32(x.x); l.x = rn_tf32(x.x - h.x);
    h.y = rn_tf32(x.y); l.y = rn_tf32(x.y - h.y);
    h.z = rn_tf32(x.z); l.z = rn_tf32(x.z - h.z);
    h.w = rn_tf32(x.w); l.w = rn_tf32(x.w - h.w);
    *((float4 *)((first ? hi0 : hi1) + r * d) + c) = h;
    *((float4 *)((first ? lo0 : lo1) + r * d) + c) = l;
  }
}

// ---------------------------------------------------------------- host side: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st) == cudaSuccess &&
        st == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// row-major fp32 [rows, d] -> boxes of 128 rows x 32 columns (128 bytes), 128-byte swizzle, zero fill out of bounds
static bool make_map(CUtensorMap *m, const float *base, int64_t rows, int d, int box_rows = TCN) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)d * 4};
  const cuuint32_t box[2] = {(cuuint32_t)TCKB, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace fr
