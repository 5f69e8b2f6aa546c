// rowblock.cuh -- the segmented row-block scaffolding shared by all conv kernels.
//
// A CTA of kNW warps owns `rb` consecutive segments (rows of a CSR or columns of
// a CSC) and their contiguous range of entries [E0, E1).  Every warp is split
// into G lane groups of LPR lanes ("virtual warps", VW = kNW*G per CTA); the
// entry range is cut into VW equal slices, one per group, and a group walks the
// *pieces* of segments that intersect its slice.  The groups of a warp advance
// in lockstep, one piece at a time, so every shuffle runs with the warp converged.
// A piece that covers a whole segment is finished by its group alone; a segment
// cut by a slice boundary leaves partial results in shared-memory slots (at most
// two per group: a head piece continuing a segment begun by an earlier group,
// and a tail piece beginning a segment a later group finishes), which the group
// holding the segment's FIRST piece folds together after one __syncthreads().
//
// This is the CSR+COO "hyper" idea of the reference (edge-balanced phase +
// row-parallel phase, fused_gtconv_hyper.cu:63-161) restated so that no
// per-edge score is staged in shared memory (no degree limit), a super-node row
// is spread over all groups of the CTA, and short rows run G per warp.
#pragma once

#include "common.cuh"

namespace dfgnn {

#ifndef DFGNN_KNW
#define DFGNN_KNW 8
#endif
constexpr int kNW = DFGNN_KNW;  // warps per CTA
#ifndef DFGNN_GT_WARPS_SMALL
#define DFGNN_GT_WARPS_SMALL 24  // the same for layouts of <= 8 floats per lane
#endif
#ifndef DFGNN_GT_WARPS
#define DFGNN_GT_WARPS 16  // resident warps per SM the GT kernels are compiled for
#endif
constexpr int kMaxRB = 128;   // max segments per CTA

struct RowBlock {
  int seg_lb;   // first segment of this CTA
  int nseg;     // segments in this CTA
  int E0, E1;   // entry range of the CTA
  int e, e_end; // entry range of this lane group
};

// Loads seg_ptr[seg_lb .. seg_lb+nseg] into s_ptr and computes the group's slice.
// Contains a __syncthreads().
// ALIGN > 1 rounds the slice length up to a multiple of ALIGN (staged kernels: whole batches).
template <int G, int ALIGN = 1>
__device__ __forceinline__ RowBlock rowblock_init(int* s_ptr, const int* __restrict__ seg_ptr,
                                                  int n_seg_total, int rb, int vw,
                                                  int tile = blockIdx.x) {
  RowBlock b;
  b.seg_lb = tile * rb;
  b.nseg = min(rb, n_seg_total - b.seg_lb);
  for (int i = threadIdx.x; i <= b.nseg; i += blockDim.x) s_ptr[i] = __ldg(seg_ptr + b.seg_lb + i);
  __syncthreads();
  b.E0 = s_ptr[0];
  b.E1 = s_ptr[b.nseg];
  constexpr int VW = kNW * G;
  int per = (b.E1 - b.E0 + VW - 1) / VW;
  if (ALIGN > 1) per = (per + ALIGN - 1) / ALIGN * ALIGN;
  b.e = min(b.E1, b.E0 + vw * per);
  b.e_end = min(b.E1, b.e + per);
  return b;
}

// A partial-result slot of one lane group: NV floats per lane (stored [NV][LPR])
// followed by 4 scalars: {a, b, seg (int), unused}.
template <int NV, int LPR>
struct Slot {
  static constexpr int kFloats = NV * LPR + 4;
  float* base;
  __device__ __forceinline__ Slot(float* smem, int vw, int which)
      : base(smem + (size_t)(vw * 2 + which) * kFloats) {}
  __device__ __forceinline__ int seg() const { return reinterpret_cast<const int*>(base)[NV * LPR + 2]; }
  __device__ __forceinline__ void set_seg(int s) { reinterpret_cast<int*>(base)[NV * LPR + 2] = s; }
  __device__ __forceinline__ float& a() { return base[NV * LPR + 0]; }
  __device__ __forceinline__ float& b() { return base[NV * LPR + 1]; }
  __device__ __forceinline__ float& v(int i, int gl) { return base[i * LPR + gl]; }
};

// The walk shared by every conv kernel.  One loop iteration = one chunk of <= CH entries
// of every group's CURRENT piece; a group that reaches the end of its piece closes it
// and opens the next one inside the same iteration, so the G groups of a warp stay busy
// whatever the segment lengths are (a warp's trip count is max over its groups of
// sum_pieces ceil(len / CH), and the slices are equal).
//   begin(r)                 group-local (divergent): load the segment's own operands
//   chunk(base, cnt)         warp-converged: may shuffle; cnt == 0 for an idle group
//   end(r, first, last)      group-local: `first`/`last` = the piece contains the
//                            segment's first/last entry (both: the piece is the segment)
template <int CH, class Begin, class Chunk, class End>
__device__ __forceinline__ void walk_pieces(const RowBlock& b, const int* s_ptr, Begin begin,
                                            Chunk chunk, End end) {
  int e = b.e, pend = b.e, pstart = b.e, rs = 0, re = 0;
  int r = e < b.e_end ? find_row(s_ptr, b.nseg, e) : 0;
  while (__any_sync(kFull, e < b.e_end)) {
    const bool act = e < b.e_end;
    if (act && e == pend) {
      while (s_ptr[r + 1] <= e) ++r;
      rs = s_ptr[r];
      re = s_ptr[r + 1];
      pend = min(re, b.e_end);
      pstart = e;
      begin(r);
    }
    const int cnt = act ? min(pend - e, CH) : 0;
    chunk(e, cnt);
    e += cnt;
    if (act && e == pend) end(r, pstart == rs, pend == re);
  }
}

template <int NV, int LPR>
__device__ __forceinline__ void slots_clear(float* smem, int vw, int gl) {
  if (gl < 2) Slot<NV, LPR>(smem, vw, gl).set_seg(-1);
}

// "big tiles only" launches (cap > 0) behind a staged kernel: almost every CTA owns a tile the staged
// kernel has already done.  Decide from two segment pointers before anything is staged in shared memory.
__device__ __forceinline__ bool small_tile_exit(const int* __restrict__ seg_ptr, int n_seg, int rb, int cap) {
  if (cap <= 0) return false;
  const int lb = blockIdx.x * rb, ub = min(lb + rb, n_seg);
  if (__ldg(seg_ptr + ub) - __ldg(seg_ptr + lb) > cap) return false;
  dependency_wait();
  return true;
}

}  // namespace dfgnn
