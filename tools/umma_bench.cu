// umma_bench.cu -- issue rate of tcgen05.mma kind::tf32 (M = 128) for the operand layouts the
// tensor-core kernels can use: K-major SWIZZLE_NONE (chunk-major images, what proj_tc.cu and
// dense_tc.cu build) against K-major SWIZZLE_128B, N in {64, 128, 256}, both operands in shared
// memory.  One CTA per SM, one thread issues `iters` MMAs back to back on the same operands.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/umma_bench tools/umma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(1024, 1) bench(int n, int swz, int iters, int ksteps, long long* out, int bg_mode) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  __shared__ volatile int s_done;
  if (threadIdx.x == 0) s_done = 0;
  for (int i = threadIdx.x; i < 48 * 1024; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  if (threadIdx.x >= 128) {  // background shared-memory traffic from the other warps (128 KB region behind the operands)
    float4* reg = reinterpret_cast<float4*>(smem + 128 * 1024) + (threadIdx.x - 128);
    float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    int k = 0;
    while (!s_done) {
      if (bg_mode == 1) reg[(k & 3) * 1024] = v;
      else { float4 t = reg[(k & 3) * 1024]; v.x += t.x; }
      ++k;
    }
    if (v.x == 123.456f) out[1] = k;
  }
  if (threadIdx.x == 0) {
    const uint32_t a = smem_u32(smem), b = a + 64 * 1024;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto desc = [&](uint32_t addr, uint32_t rows) -> uint64_t {
      if (swz == 1)  // SWIZZLE_128B, K-major: 8-row atoms of 1024 B
        return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((rows * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
    };
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint32_t off = swz == 1 ? ks * 32 : ks * 2 * 128 * 16;
        const uint64_t da = desc(a + off, 128), db = desc(b + off, 128);
        if (swz == 2)  // A from tensor memory (columns 256 + 8 ks), B from shared memory (SWIZZLE_NONE images)
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem),
                       "r"(tmem + 256 + ks * 8), "l"(db), "r"(idesc), "r"(1u)
                       : "memory");
        else
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
                       "l"(da), "l"(db), "r"(idesc), "r"(1u)
                       : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
    const long long t1 = clock64();
    s_done = 1;
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 512, ksteps = 4;
  for (int swz = 0; swz < 3; ++swz)
    for (int n : {64, 128, 256}) {
      for (int rep = 0; rep < 2; ++rep) bench<<<148, 128, 200 * 1024>>>(n, swz, iters, ksteps, d, 0);
      long long h = 0;
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      cudaError_t e = cudaDeviceSynchronize();
      printf("layout %-12s M=128 N=%3d: %.1f cycles per tcgen05.mma (K = 8, tf32)  floor %d   [%s]\n",
             swz == 2 ? "A in TMEM" : swz ? "SWIZZLE_128B" : "SWIZZLE_NONE", n, (double)h / (iters * ksteps), 128 * n / 256, cudaGetErrorString(e));
    }
  // the same MMAs (SWIZZLE_NONE, N = 256 and 128) with other warps of the CTA writing / reading shared memory
  for (int mode = 1; mode <= 2; ++mode)
    for (int bgw : {4, 8, 16})
      for (int n : {128, 256}) {
        for (int rep = 0; rep < 2; ++rep) bench<<<148, 128 + 32 * bgw, 200 * 1024>>>(n, 0, iters, ksteps, d, mode);
        long long h = 0;
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaDeviceSynchronize();
        printf("%d background warps doing %s: N=%3d %.1f cycles per tcgen05.mma   [%s]\n", bgw, mode == 1 ? "STS.128" : "LDS.128", n,
               (double)h / (iters * ksteps), cudaGetErrorString(e));
      }
  return 0;
}
