// tc_common.cuh -- tcgen05 / tensor-memory wrappers shared by the tensor-core kernels
// (proj_tc.cu, dense_tc.cu).  PTX ISA syntax as in CUTLASS cute/arch/mma_sm100_umma.hpp and
// copy_sm100.hpp; operands are shared-memory images in the canonical K-major NO-SWIZZLE layout
// ("chunk major": the 16-byte k-chunk c of row r at c * LBO + r * 16, so SBO = 128 B between 8-row
// groups and LBO between the two chunks of one K = 8 MMA).
#pragma once

#include "common.cuh"

namespace dfgnn {

// ---- tcgen05 wrappers (PTX ISA; syntax as in CUTLASS cute/arch/mma_sm100_umma.hpp, copy_sm100.hpp) ----
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // shared-memory matrix descriptor, SWIZZLE_NONE: start address, leading / stride byte offsets in
  // 16-byte units, version 1 (Blackwell) at bits [46, 48)
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
  // c = F32 (1 << 4), a = b = TF32 (2 << 7, 2 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// x = hi + lo with hi = x rounded to TF32 (10 mantissa bits; round half away from zero by an integer add
// on the bit pattern -- cvt.rna.tf32.f32 is a seven-instruction sequence on sm_100a) and lo = x - hi exact
// in fp32.  Rounding instead of truncating halves |lo| (<= 2^-11 |x|) and makes the error the tensor core
// adds by truncating lo to TF32 (<= 2^-21 |x|) sign-symmetric: truncated hi gave every product a one-sided
// error of up to 2^-20, which a sum over 128 same-sign terms does not average away.
__device__ __forceinline__ float round_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void split4(const float4 x, float4& hi, float4& lo) {
  hi.x = round_tf32(x.x);
  hi.y = round_tf32(x.y);
  hi.z = round_tf32(x.z);
  hi.w = round_tf32(x.w);
  lo = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

}  // namespace dfgnn
