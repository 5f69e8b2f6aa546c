// rowblock.cuh -- the segmented row-block scaffolding shared by all conv kernels.
//
// A CTA of kNW warps owns `rb` consecutive segments (rows of a CSR or columns of
// a CSC) and their contiguous range of entries [E0, E1).  That range is cut
// into kNW equal slices, one per warp; a warp walks the *pieces* of segments
// that intersect its slice.  A piece that covers a whole segment is finished by
// the warp on its own; a segment cut by a slice boundary leaves partial results
// in shared-memory slots (at most two per warp: a head piece that continues a
// segment begun by an earlier warp, and a tail piece that begins a segment some
// later warp finishes) which the warp that holds the segment's FIRST piece
// folds together after one __syncthreads().
//
// This is the CSR+COO "hyper" idea of the reference (edge-balanced phase +
// row-parallel phase, fused_gtconv_hyper.cu:63-161) restated so that no
// per-edge score is ever staged in shared memory (no degree limit) and so that
// a super-node row is spread over all warps of the CTA.
#pragma once

#include "common.cuh"

namespace dfgnn {

constexpr int kNW = 8;       // warps per CTA
constexpr int kMaxRB = 64;   // max segments per CTA

struct RowBlock {
  int seg_lb;   // first segment of this CTA
  int nseg;     // segments in this CTA
  int E0, E1;   // entry range of the CTA
  int e, e_end; // entry range of this warp
};

// Loads seg_ptr[seg_lb .. seg_lb+nseg] into s_ptr and computes the warp's slice.
// Contains a __syncthreads().
__device__ __forceinline__ RowBlock rowblock_init(int* s_ptr, const int* __restrict__ seg_ptr,
                                                  int n_seg_total, int rb) {
  RowBlock b;
  b.seg_lb = blockIdx.x * rb;
  b.nseg = min(rb, n_seg_total - b.seg_lb);
  for (int i = threadIdx.x; i <= b.nseg; i += blockDim.x) s_ptr[i] = __ldg(seg_ptr + b.seg_lb + i);
  __syncthreads();
  b.E0 = s_ptr[0];
  b.E1 = s_ptr[b.nseg];
  const int w = threadIdx.x >> 5;
  const int per = (b.E1 - b.E0 + kNW - 1) / kNW;
  b.e = min(b.E1, b.E0 + w * per);
  b.e_end = min(b.E1, b.e + per);
  return b;
}

// A partial-result slot: NV floats per lane (stored [NV][32], conflict free)
// followed by 4 scalars: {a, b, seg (int), unused}.
template <int NV>
struct Slot {
  static constexpr int kFloats = NV * 32 + 4;
  float* base;
  __device__ __forceinline__ Slot(float* smem, int warp, int which)
      : base(smem + (size_t)(warp * 2 + which) * kFloats) {}
  __device__ __forceinline__ int seg() const { return reinterpret_cast<const int*>(base)[NV * 32 + 2]; }
  __device__ __forceinline__ void set_seg(int s) { reinterpret_cast<int*>(base)[NV * 32 + 2] = s; }
  __device__ __forceinline__ float& a() { return base[NV * 32 + 0]; }
  __device__ __forceinline__ float& b() { return base[NV * 32 + 1]; }
  __device__ __forceinline__ float& v(int i, int lane) { return base[i * 32 + lane]; }
};

template <int NV>
__device__ __forceinline__ void slots_clear(float* smem) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane < 2) Slot<NV>(smem, w, lane).set_seg(-1);
}

}  // namespace dfgnn
