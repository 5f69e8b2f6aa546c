// common.cuh -- shared device helpers for the dfgnn_b200 kernels (sm_100a only).
//
// Design: every conv kernel is a "segmented row-block" kernel.
// A CTA owns RB consecutive rows of the CSR (or columns of the CSC); the CTA's
// contiguous edge range is split EVENLY over its warps, so a warp's range may
// cover several short rows or a slice of one long row.  Each warp walks the
// row pieces inside its range with a flash-style online softmax, neighbour
// feature rows are gathered with 16-byte vector loads spread over LPR lanes,
// and pieces of rows that straddle warps are merged through shared memory.
// Scores never touch HBM (training forward stores them once, as the
// reference's API requires attn_edge).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace dfgnn {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr float kNeg = -1e30f;  // "minus infinity" that survives subtraction
constexpr float kLog2e = 1.4426950408889634f;

// ------------------------------------------------------------------------- //
// Feature-row layouts: how the f floats of one node row are spread over a    //
// lane group.  LPR lanes cooperate on one row and form a "virtual warp" that //
// walks its own slice of edges; a warp holds G = 32/LPR such groups running  //
// in lockstep.  Every lane keeps NR floats of the row in registers.          //
// ------------------------------------------------------------------------- //

// f = 4*F4: 16-byte loads; lane gl of the group owns float4 number v*LPR + gl
// (one load instruction covers LPR*16 contiguous bytes per group).
template <int F4_, int LPR_>
struct VecLayout {
  static_assert(LPR_ >= 1 && LPR_ <= 32 && (LPR_ & (LPR_ - 1)) == 0 && F4_ % LPR_ == 0, "bad layout");
  static constexpr bool kVec = true;
  static constexpr int F4 = F4_;
  static constexpr int LPR = LPR_;
  static constexpr int VPL = F4 / LPR;
  static constexpr int G = 32 / LPR;
  static constexpr int NR = 4 * VPL;

  // `row` = row start + lane_off(gl) bytes (see RowAddr)
  __device__ __forceinline__ static int lane_off(int gl) { return gl * 16; }
  __device__ __forceinline__ static void load(float (&r)[NR], const float* __restrict__ row,
                                              int /*gl*/, int /*f*/) {
    const float4* p = reinterpret_cast<const float4*>(row);
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const float4 t = __ldg(p + v * LPR);
      r[4 * v + 0] = t.x; r[4 * v + 1] = t.y; r[4 * v + 2] = t.z; r[4 * v + 3] = t.w;
    }
  }
  // same from shared memory (`row` already holds the lane offset)
  __device__ __forceinline__ static void load_smem(float (&r)[NR], const float* row) {
    const float4* p = reinterpret_cast<const float4*>(row);
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const float4 t = p[v * LPR];
      r[4 * v + 0] = t.x; r[4 * v + 1] = t.y; r[4 * v + 2] = t.z; r[4 * v + 3] = t.w;
    }
  }
  __device__ __forceinline__ static void store(float* __restrict__ row, const float (&r)[NR],
                                               int /*gl*/, int /*f*/) {
    float4* p = reinterpret_cast<float4*>(row);
#pragma unroll
    for (int v = 0; v < VPL; ++v)
      p[v * LPR] = make_float4(r[4 * v + 0], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
  }
};

// any f <= 32*NT: scalar loads, the whole warp is one group, lane l owns features l, l+32, ...
template <int NT_>
struct ScalarLayout {
  static constexpr bool kVec = false;
  static constexpr int F4 = 0;
  static constexpr int LPR = 32;
  static constexpr int G = 1;
  static constexpr int NR = NT_;

  __device__ __forceinline__ static int lane_off(int gl) { return gl * 4; }
  __device__ __forceinline__ static void load(float (&r)[NR], const float* __restrict__ row,
                                              int gl, int f) {
#pragma unroll
    for (int t = 0; t < NR; ++t) r[t] = (t * 32 + gl) < f ? __ldg(row + t * 32) : 0.f;
  }
  __device__ __forceinline__ static void store(float* __restrict__ row, const float (&r)[NR],
                                               int gl, int f) {
#pragma unroll
    for (int t = 0; t < NR; ++t)
      if ((t * 32 + gl) < f) row[t * 32] = r[t];
  }
};

// Row addressing with ONE 32x32+64 multiply-add per gathered row: every thread keeps
// byte pointers that already contain its head and lane offsets.
template <class L>
struct RowAddr {
  uint32_t stride_b;  // bytes between consecutive node rows = h * f * 4
  int hid, f, gl;
  __device__ __forceinline__ RowAddr(int h, int f_, int hid_, int gl_)
      : stride_b((uint32_t)h * (uint32_t)f_ * 4u), hid(hid_), f(f_), gl(gl_) {}
  // The empty asm keeps base + lane offset together in one 64-bit register pair, so that a
  // gathered row address is a single IMAD.WIDE (row * stride + base) instead of the
  // compiler's (row * stride + lane offset) + uniform base split.
  __device__ __forceinline__ const char* base(const float* t) const {
    const char* b = reinterpret_cast<const char*>(t + (size_t)hid * f) + L::lane_off(gl);
    asm volatile("" : "+l"(b));
    return b;
  }
  __device__ __forceinline__ char* base(float* t) const {
    char* b = reinterpret_cast<char*>(t + (size_t)hid * f) + L::lane_off(gl);
    asm volatile("" : "+l"(b));
    return b;
  }
  __device__ __forceinline__ const float* at(const char* b, int row) const {
    return reinterpret_cast<const float*>(b + (uint64_t)(uint32_t)row * stride_b);
  }
  __device__ __forceinline__ float* at(char* b, int row) const {
    return reinterpret_cast<float*>(b + (uint64_t)(uint32_t)row * stride_b);
  }
};

// edges a lane group keeps in flight per step: bounded by registers (two rows of
// NR floats per edge in the GT kernels) and by the group's chunk of LPR edges.
template <class L>
struct ChunkOf {
  // GT kernels hold two rows of NR floats per entry; NR <= 8 is compiled for 24 resident warps
  // per SM (<= 80 registers), which pays more than a deeper batch (measured on PATTERN / VOC)
  static constexpr int kByReg = L::NR <= 4 ? 4 : (L::NR <= 16 ? 2 : 1);
  static constexpr int kChunk = L::LPR < 8 ? L::LPR : 8;  // edges whose indices a group prefetches
#ifdef DFGNN_GT_C  // developer knob: entries in flight per group in the GT kernels
  static constexpr int C = DFGNN_GT_C;
#else
  static constexpr int C = kByReg < kChunk ? kByReg : kChunk;
#endif
  // kernels that gather ONE row per edge (GAT) can keep twice as many edges in flight
  static constexpr int kByReg1 = L::NR <= 4 ? 8 : (L::NR <= 8 ? 4 : (L::NR <= 16 ? 4 : 2));
  static constexpr int C1 = kByReg1 < L::LPR ? kByReg1 : L::LPR;
};

template <int N>
__device__ __forceinline__ void zero(float (&r)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) r[i] = 0.f;
}

template <int N>
__device__ __forceinline__ float dot(const float (&a)[N], const float (&b)[N]) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) s = fmaf(a[i], b[i], s);
  return s;
}

// all-reduce inside an aligned group of LPR lanes (the warp must be converged)
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int off = LPR / 2; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
  return v;
}
template <int LPR>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
  for (int off = LPR / 2; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, off));
  return v;
}
// group all-reduce that is legal in group-divergent code: only the caller's own group
// takes part (every lane of the group must be there)
template <int LPR>
__device__ __forceinline__ float group_sum_local(float v, int lane) {
  const unsigned mask = LPR == 32 ? kFull : (((1u << (LPR & 31)) - 1u) << (lane & ~(LPR - 1)));
#pragma unroll
  for (int off = LPR / 2; off > 0; off >>= 1) v += __shfl_xor_sync(mask, v, off);
  return v;
}
// value held by lane `src` (0..LPR-1) of the caller's own group
template <int LPR, class T>
__device__ __forceinline__ T group_bcast(T v, int src) {
  return __shfl_sync(kFull, v, src, LPR);
}

// Group totals of FOUR per-lane partial sums with 2 + log2(LPR / 2) shuffles instead of
// 4 * log2(LPR): lanes swap halves of their values (butterfly on the values, not the lanes).
// Returns the total of d[perm4(gl)], perm4(gl) = 2 * (gl & 1) + ((gl >> 1) & 1).
__device__ __forceinline__ int perm4(int gl) { return ((gl & 1) << 1) | ((gl >> 1) & 1); }
template <int LPR>
__device__ __forceinline__ float reduce4_transposed(const float (&d)[4], int gl) {
  static_assert(LPR >= 4, "needs at least four lanes per group");
  const bool o1 = gl & 1, o2 = gl & 2;
  const float s0 = o1 ? d[0] : d[2], s1 = o1 ? d[1] : d[3];
  float k0 = o1 ? d[2] : d[0], k1 = o1 ? d[3] : d[1];
  k0 += __shfl_xor_sync(kFull, s0, 1);
  k1 += __shfl_xor_sync(kFull, s1, 1);
  const float s = o2 ? k0 : k1;
  float k = o2 ? k1 : k0;
  k += __shfl_xor_sync(kFull, s, 2);
#pragma unroll
  for (int off = 4; off < LPR; off <<= 1) k += __shfl_xor_sync(kFull, k, off);
  return k;
}

__device__ __forceinline__ float warp_sum(float v) { return group_sum<32>(v); }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, off));
  return v;
}

// largest r in [0, n) with a[r] <= e, for a non-decreasing smem array a[0..n]
// with a[0] <= e < a[n].  Skips empty rows (a[r] == a[r+1]) by construction.
__device__ __forceinline__ int find_row(const int* a, int n, int e) {
  int lo = 0, hi = n;  // invariant: a[lo] <= e < a[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] <= e) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ float leaky(float x, float slope) { return x > 0.f ? x : x * slope; }

// e^x and 2^x as one MUFU.EX2 (x <= 0 here; results below the normal range flush to 0)
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_exp(float x) { return fast_exp2(x * kLog2e); }
// 16-byte asynchronous global -> shared copy (LDGSTS): no register staging, the issuing thread
// goes on; cp_async_wait_all() + a barrier make the data visible to the CTA.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- TMA bulk copies (1-D cp.async.bulk, SASS UBLKCP) completing on an mbarrier ------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// makes the barrier initialisation visible to the async proxy (the TMA unit)
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// a contiguous range of `bytes` bytes in pieces of at most 16 KB (several requests in flight);
// called by ONE thread that has already done mbar_expect_tx for the total
__device__ __forceinline__ void bulk_g2s_range(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  constexpr uint32_t kPiece = 16 * 1024;
  for (uint32_t off = 0; off < bytes; off += kPiece)
    bulk_g2s(static_cast<char*>(smem_dst) + off, static_cast<const char*>(gmem_src) + off,
             bytes - off < kPiece ? bytes - off : kPiece, bar);
}

// lets the next kernel on the stream start early if it was launched with
// cudaLaunchAttributeProgrammaticStreamSerialization (abi_common.h: launch_overlapped)
__device__ __forceinline__ void allow_dependent_launch() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// Counterpart for a kernel launched that way: returns once every kernel in front of it on the
// stream has completed and its writes are visible.  The "big tiles only" row-block kernels call
// it before they exit, so that THEIR completion implies the staged kernel's: without it a
// dependent grid that finishes early would let later work on the stream run while the primary
// kernel is still writing.  A no-op for a normally launched kernel.
__device__ __forceinline__ void dependency_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// 1/x as one MUFU.RCP (1 ulp)
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// counter-based uniform in (0, 1]: two rounds of a 64-bit mix (splitmix64
// finaliser) over (seed, edge index).  Replaces the cuRAND stream the reference
// draws edge_mask from (fused_gatconv_kernel.cu:1073-1081).
__device__ __forceinline__ float uniform01(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (float)((z >> 40) + 1) * (1.0f / 16777216.0f);  // (0, 1]
}

}  // namespace dfgnn
