// mma_bench.cu -- issue rate of the legacy warp-level tensor-core path (mma.sync.m16n8k8 TF32) on
// sm_100a, register operands only and with B fragments read from padded shared memory + hi/lo
// split (the inner loop shape of a 3xTF32 masked-attention tile).  Developer tool, not product.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/mma_bench tools/mma_bench.cu && tools/mma_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int NACC>
__global__ void __launch_bounds__(256) k_regs(float* out, int iters) {
  float d[NACC][4] = {};
  uint32_t a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3};
  uint32_t b[2] = {threadIdx.x * 3, threadIdx.x * 5};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) mma_tf32(d[i], a, b);
  }
  float s = 0.f;
  for (int i = 0; i < NACC; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 16 x 64 output tile per warp, K = 128: B from smem (row stride 132), 3 MMAs per fragment pair
__global__ void __launch_bounds__(256) k_smem3(float* out, int iters) {
  extern __shared__ float sK[];  // 64 rows x 132
  for (int i = threadIdx.x; i < 64 * 132; i += blockDim.x) sK[i] = (float)(i % 97) * 0.01f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float q[16][4];
  for (int k = 0; k < 16; ++k)
    for (int i = 0; i < 4; ++i) q[k][i] = (float)(lane + k + i) * 0.001f;
  float acc[8][4] = {};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) {
      uint32_t ah[4], al[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t x = __float_as_uint(q[ks][i]);
        ah[i] = x & 0xffffe000u;
        al[i] = __float_as_uint(q[ks][i] - __uint_as_float(ah[i]));
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float* p = sK + (nt * 8 + (lane >> 2)) * 132 + ks * 8 + (lane & 3);
        const float b0 = p[0], b1 = p[4];
        uint32_t bh[2] = {__float_as_uint(b0) & 0xffffe000u, __float_as_uint(b1) & 0xffffe000u};
        uint32_t bl[2] = {__float_as_uint(b0 - __uint_as_float(bh[0])), __float_as_uint(b1 - __uint_as_float(bh[1]))};
        mma_tf32(acc[nt], al, bh);
        mma_tf32(acc[nt], ah, bl);
        mma_tf32(acc[nt], ah, bh);
      }
    }
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int warps : {4, 8, 16, 32}) {
    const int iters = 4096, ctas = 148 * (warps <= 8 ? 1 : warps / 8);
    const int thr = warps <= 8 ? warps * 32 : 256;
    k_regs<8><<<ctas, thr>>>(out, 16);
    cudaEventRecord(e0);
    k_regs<8><<<ctas, thr>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)ctas * (thr / 32) * iters * 8;
    printf("regs only  : %2d warps/SM  %.3f ms  %.1f TF32-TFLOP/s  (%.2f clk/mma/SM @1.9GHz)\n", warps, ms,
           mmas * 2048 / ms / 1e9, ms * 1e-3 * 1.9e9 / (mmas / 148));
  }
  cudaFuncSetAttribute(k_smem3, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 132 * 4);
  for (int ctas_per_sm : {1, 2}) {
    const int iters = 256, ctas = 148 * ctas_per_sm;
    k_smem3<<<ctas, 256, 64 * 132 * 4>>>(out, 4);
    cudaEventRecord(e0);
    k_smem3<<<ctas, 256, 64 * 132 * 4>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)ctas * 8 * iters * 16 * 8 * 3;
    printf("smem 3xTF32: %2d warps/SM  %.3f ms  %.1f TF32-TFLOP/s issued = %.1f TFLOP/s of fp32-grade product  (%.2f clk/mma/SM)\n",
           8 * ctas_per_sm, ms, mmas * 2048 / ms / 1e9, mmas * 2048 / 3 / ms / 1e9, ms * 1e-3 * 1.9e9 / (mmas / 148));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
