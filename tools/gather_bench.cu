// gather_bench.cu -- microbenchmark (developer tool): random row gathers, register-staged
// LDG.128 vs TMA bulk copies (cp.async.bulk, one 256/512-byte row per issuing lane) into a
// per-warp shared-memory ring.  Decides whether the conv kernels should stage neighbour
// rows through the async proxy.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gather_bench gather_bench.cu
//   ./gather_bench [n_rows] [n_edges] [f]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(b))
               : "memory");
}

// ---- register-staged baseline: 8 lanes per row, C rows in flight per group ----
template <int F4, int C>
__global__ void __launch_bounds__(256) ldg_kernel(const float4* __restrict__ X, const int* __restrict__ idx, int n_edges,
                                                  float* __restrict__ out) {
  constexpr int VPL = F4 / 8;
  const int lane = threadIdx.x & 31, gl = lane & 7;
  const int group = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, ngroups = (gridDim.x * blockDim.x) >> 3;
  const int per = (n_edges + ngroups - 1) / ngroups;
  int e = group * per;
  const int e_end = min(n_edges, e + per);
  float acc[4 * VPL] = {};
  for (; e < e_end; e += C) {
    float4 v[C][VPL];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int r = __ldg(idx + min(e + c, e_end - 1));
#pragma unroll
      for (int k = 0; k < VPL; ++k) v[c][k] = __ldg(X + (size_t)r * F4 + k * 8 + gl);
    }
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        acc[4 * k] += v[c][k].x; acc[4 * k + 1] += v[c][k].y; acc[4 * k + 2] += v[c][k].z; acc[4 * k + 3] += v[c][k].w;
      }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4 * VPL; ++i) s += acc[i];
  if (s == 123.456f) out[0] = s;
}

// ---- TMA bulk gather: every lane issues one row copy per stage; ring of S stages per warp ----
template <int F4, int S, int R>  // R rows per stage (<= 32)
__global__ void __launch_bounds__(256) bulk_kernel(const float4* __restrict__ X, const int* __restrict__ idx, int n_edges,
                                                   float* __restrict__ out) {
  constexpr int VPL = F4 / 8, ROWB = F4 * 16;
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, gl = lane & 7, grp = lane >> 3;
  float4* buf = reinterpret_cast<float4*>(smem) + (size_t)w * S * R * F4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)(blockDim.x >> 5) * S * R * ROWB) + w * S;
  if (lane == 0)
    for (int s = 0; s < S; ++s) mbar_init(bars + s, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const int per = ((n_edges + nwarps - 1) / nwarps + R - 1) / R * R;
  const int e0 = warp * per, e_end = min(n_edges, e0 + per);
  const int nchunks = e0 < e_end ? (e_end - e0 + R - 1) / R : 0;
  auto issue = [&](int chunk) {
    const int s = chunk % S;
    const int e = e0 + chunk * R;
    const int cnt = min(R, e_end - e);
    if (lane == 0) mbar_expect_tx(bars + s, (uint32_t)cnt * ROWB);
    __syncwarp();
    if (lane < cnt) {
      const int r = __ldg(idx + e + lane);
      bulk_g2s(buf + (size_t)(s * R + lane) * F4, X + (size_t)r * F4, ROWB, bars + s);
    }
  };
  for (int c = 0; c < S - 1 && c < nchunks; ++c) issue(c);
  float acc[4 * VPL] = {};
  for (int c = 0; c < nchunks; ++c) {
    if (c + S - 1 < nchunks) issue(c + S - 1);
    const int s = c % S;
    mbar_wait(bars + s, (c / S) & 1);
    const int cnt = min(R, e_end - (e0 + c * R));
    for (int j = grp; j < cnt; j += 4) {
      const float4* row = buf + (size_t)(s * R + j) * F4;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const float4 v = row[k * 8 + gl];
        acc[4 * k] += v.x; acc[4 * k + 1] += v.y; acc[4 * k + 2] += v.z; acc[4 * k + 3] += v.w;
      }
    }
    __syncwarp();  // all lanes done reading stage s before it is refilled
  }
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 4 * VPL; ++i) sum += acc[i];
  if (sum == 123.456f) out[0] = sum;
}

template <class F>
static float time_it(F launch, float* flush, size_t flush_n, bool do_flush) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) launch();
  float best = 1e30f, tot = 0.f;
  const int it = 10;
  for (int i = 0; i < it; ++i) {
    if (do_flush) CK(cudaMemsetAsync(flush, i, flush_n));
    cudaEventRecord(a);
    launch();
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    best = ms < best ? ms : best;
    tot += ms;
  }
  return tot / it;
}

template <int F4>
static void run(int n, int ne) {
  const size_t xb = (size_t)n * F4 * 16;
  float4* X; int* idx; float* out; float* flush;
  const size_t flush_n = 256u << 20;
  CK(cudaMalloc(&X, xb)); CK(cudaMalloc(&idx, (size_t)ne * 4)); CK(cudaMalloc(&out, 16)); CK(cudaMalloc(&flush, flush_n));
  CK(cudaMemset(X, 0, xb));
  std::vector<int> h(ne);
  uint64_t z = 88172645463325252ull;
  for (int i = 0; i < ne; ++i) { z ^= z << 13; z ^= z >> 7; z ^= z << 17; h[i] = (int)(z % (uint64_t)n); }
  CK(cudaMemcpy(idx, h.data(), (size_t)ne * 4, cudaMemcpyHostToDevice));
  const double gb = (double)ne * F4 * 16 / 1e9;
  printf("rows %d x %d B (%.1f MB), %d gathers (%.1f MB)\n", n, F4 * 16, xb / 1e6, ne, gb * 1e3);
  for (int flushed = 0; flushed < 2; ++flushed) {
    {
      const int grid = 148 * 8;
      float ms = time_it([&] { ldg_kernel<F4, 4><<<grid, 256>>>(X, idx, ne, out); }, flush, flush_n, flushed);
      printf("  %-34s %s  %8.1f us  %7.1f GB/s\n", "LDG.128 C=4 (8 CTAs/SM)", flushed ? "cold" : "warm", ms * 1e3, gb / ms * 1e3);
    }
    auto bulk = [&](auto kern, const char* name, int S, int R, int ctas) {
      const size_t smem = (size_t)8 * S * R * F4 * 16 + 8 * S * 8;
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      float ms = time_it([&] { kern<<<148 * ctas, 256, smem>>>(X, idx, ne, out); }, flush, flush_n, flushed);
      CK(cudaGetLastError());
      printf("  %-34s %s  %8.1f us  %7.1f GB/s  (smem %zu KB/CTA)\n", name, flushed ? "cold" : "warm", ms * 1e3, gb / ms * 1e3, smem >> 10);
    };
    bulk(bulk_kernel<F4, 2, 32>, "bulk S=2 R=32 x2 CTA", 2, 32, 2);
    bulk(bulk_kernel<F4, 3, 16>, "bulk S=3 R=16 x2 CTA", 3, 16, 2);
    bulk(bulk_kernel<F4, 4, 16>, "bulk S=4 R=16 x2 CTA", 4, 16, 2);
    bulk(bulk_kernel<F4, 4, 8>, "bulk S=4 R=8 x4 CTA", 4, 8, 4);
    bulk(bulk_kernel<F4, 8, 8>, "bulk S=8 R=8 x2 CTA", 8, 8, 2);
  }
  cudaFree(X); cudaFree(idx); cudaFree(out); cudaFree(flush);
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 169343;
  const int ne = argc > 2 ? atoi(argv[2]) : 1166173;
  const int f = argc > 3 ? atoi(argv[3]) : 64;
  if (f == 64) run<16>(n, ne);
  else run<32>(n, ne);
  return 0;
}
