// staged.cuh -- building blocks of the "staged tile" kernels (short-row graphs).
//
// The segmented row-block kernels (rowblock.cuh) carry an online softmax through
// every row piece; on graphs whose rows are a handful of entries long (arxiv-,
// cora-, PascalVOC-shaped: mean degree 4-7) that bookkeeping costs more issue slots
// than the gathers themselves.  The staged kernels (staged_gat.cuh) split a CTA's tile
// of <= kStageCap entries into phases over per-entry records held in shared memory:
//
//   A  stage     every thread loads the indices of up to kEPT entries, then issues all
//                dependent 4/8-byte gathers at once, and writes {neighbour, scalar} records
//   B  per row   shared memory only: short rows one THREAD each, long rows one WARP each
//                (max / sum / exp weights, or the backward's row sums)
//   C  flat_spmm every lane group walks an EQUAL, whole-batch slice of the tile's entries
//                and accumulates weight * gathered row; the last entry of a row is flagged,
//                so a row boundary costs one short divergent flush; rows cut by a slice
//                boundary are merged through the same shared-memory slots as in
//                rowblock.cuh (deterministic order)
//      flat_sddmm(_sx)  the same walk producing one dot product per entry (backward:
//                g = <dO_i, feat_j>), row operands of the tile staged in shared memory
//
// This is the reference's CSR+COO "hyper" idea (edge-balanced SDDMM into shared
// memory, then row-parallel softmax + SpMM; fused_gtconv_hyper.cu:63-161,
// fused_gatconv_hyper.cu:37-110) with the SpMM edge-balanced as well.  Tiles larger
// than kStageCap (super rows) are left to the row-block kernels, launched behind the
// staged kernel in "big tiles only" mode -- so there is still no degree limit.
#pragma once

#include "rowblock.cuh"

namespace dfgnn {

#ifndef DFGNN_STAGE_CAP
#define DFGNN_STAGE_CAP 2048
#endif
constexpr int kStageCap = DFGNN_STAGE_CAP;  // entries of a CTA tile staged in shared memory

// Staged tile entries.  The arrays hold kStageCap + kStagePad entries: the C entries behind
// the tile must be {idx = 0, weights = 0} (stage_pad) so that the last batch of the last
// slice can be read without bounds checks.
constexpr int kStagePad = 8;
struct __align__(8) Ent1 { int idx; float w; };                    // one weight
struct __align__(16) Ent2 { int idx; float w; float w1; float aux; };  // two weights (or weight + scalar)

template <class E>
__device__ __forceinline__ void stage_pad(E* s_e, int ne) {
  if (threadIdx.x < kStagePad) {
    E z{};
    s_e[ne + threadIdx.x] = z;
  }
}

// floats of the partial-result slots of one CTA (two per lane group)
template <int NV, class L>
constexpr int slot_floats() { return kNW * L::G * 2 * Slot<NV, L::LPR>::kFloats; }

// row operands of a tile are staged in shared memory for vector layouts up to f = 128
template <class L>
constexpr bool stage_x() { return L::kVec && L::NR * L::LPR <= 128; }

enum SpmmMode { kOneOp = 0, kOneOpScalar = 1, kTwoOps = 2 };

// The LAST entry of every segment carries this flag in idx (neighbour ids are < 2^31).
constexpr int kLastFlag = (int)0x80000000u;

// s_next[r] = the next segment after r that has entries (nseg if none); marks segment ends.
template <class E>
__device__ __forceinline__ void stage_mark_ends(const RowBlock& b, const int* s_ptr, E* s_e,
                                                int* s_next) {
  const int r = threadIdx.x;
  if (r < b.nseg) {
    const int rs = s_ptr[r], re = s_ptr[r + 1];
    if (re > rs) s_e[re - 1 - b.E0].idx |= kLastFlag;
    int nx = r + 1;
    while (nx < b.nseg && s_ptr[nx + 1] == s_ptr[nx]) ++nx;
    s_next[r] = nx;
  }
}

// out[seg] = sum_e w[e] * X[idx[e], :] over the staged tile (kOneOp); kOneOpScalar also sums
// w1[e] per segment; kTwoOps walks a second operand matrix X1 with weights w1.
// The group's slice [b.e, b.e_end) is a whole number of C-entry batches (rowblock_init<G, C>;
// the tile's tail is padded by stage_pad), segment ends are flagged (stage_mark_ends).
// store(seg, scalar, acc[NV]) is called for segments that lie inside one slice; pieces of
// split segments are left in the slots (merge with sum_merge_slots after a __syncthreads()).
template <class L, int C, int MODE, class E, class Store>
__device__ __forceinline__ void flat_spmm(const RowBlock& b, const int* s_ptr, const int* s_next,
                                          const E* s_e, const RowAddr<L>& ra, const char* X0,
                                          const char* X1, float* s_slot, int vw, int gl, int f,
                                          Store store) {
  constexpr int NOPS = MODE == kTwoOps ? 2 : 1;
  constexpr int NR = L::NR, NV = NOPS * NR, LPR = L::LPR;
  static_assert(C <= kStagePad, "padding");
  int e = b.e;
  const int e_end = b.e_end;
  if (e >= e_end) return;
  int r = find_row(s_ptr, b.nseg, e);
  bool head = e > s_ptr[r];  // the slice starts inside a segment
  float acc[NV], sc = 0.f;
  zero(acc);
  auto to_slot = [&](int which) {
    Slot<NV, LPR> sl(s_slot, vw, which);
#pragma unroll
    for (int i = 0; i < NV; ++i) sl.v(i, gl) = acc[i];
    if (gl == 0) { sl.a() = sc; sl.set_seg(r); }
  };
  const E* pe = s_e + (e - b.E0);
  for (; e < e_end; e += C, pe += C) {
    E en[C];
#pragma unroll
    for (int c = 0; c < C; ++c) en[c] = pe[c];
    float v0[C][NR], v1[NOPS == 2 ? C : 1][NR];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int row = en[c].idx & ~kLastFlag;
      L::load(v0[c], ra.at(X0, row), gl, f);
      if constexpr (MODE == kTwoOps) L::load(v1[c], ra.at(X1, row), gl, f);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        acc[i] = fmaf(en[c].w, v0[c][i], acc[i]);
        if constexpr (MODE == kTwoOps) acc[NR + i] = fmaf(en[c].w1, v1[c][i], acc[NR + i]);
      }
      if constexpr (MODE == kOneOpScalar) sc += en[c].w1;
      if (en[c].idx < 0) {  // last entry of segment r
        if (head) to_slot(0);
        else store(r, sc, acc);
        zero(acc);
        sc = 0.f;
        head = false;
        r = s_next[r];
      }
    }
  }
  // the slice ends inside a segment (never on the padded tail: the tile's last entry is flagged)
  if (s_e[e_end - 1 - b.E0].idx >= 0) to_slot(head ? 0 : 1);
}

// out(e, <Xrow[seg(e), :], Y[idx[e], :]>) for every staged entry e (tile-relative).
// One lane group per entry, C entries in flight; the row operand of the NEXT segment is
// prefetched into registers when a segment starts.  Warp-converged (shuffles inside).
template <class L, int C, class E, class Out>
__device__ __forceinline__ void flat_sddmm(const RowBlock& b, const int* s_ptr, const E* s_e,
                                           const RowAddr<L>& ra, const char* Xrow, const char* Y,
                                           int gl, int f, Out out) {
  constexpr int NR = L::NR, LPR = L::LPR;
  static_assert(C <= LPR, "one result lane per entry in flight");
  int e = b.e;
  const int e_end = b.e_end;
  int r = 0, row_end = 0, rn = 0;
  float x[NR], xn[NR];
  zero(x);
  zero(xn);
  if (e < e_end) {
    r = find_row(s_ptr, b.nseg, e);
    row_end = s_ptr[r + 1];
    L::load(x, ra.at(Xrow, b.seg_lb + r), gl, f);
    if (row_end < e_end) {
      rn = r + 1;
      while (s_ptr[rn + 1] <= row_end) ++rn;
      L::load(xn, ra.at(Xrow, b.seg_lb + rn), gl, f);
    }
  }
  while (__any_sync(kFull, e < e_end)) {
    // entries behind the slice are valid neighbours too (next slice or stage_pad); a group
    // that has run out of entries parks on the padding
    const E* pe = s_e + ((e < e_end ? e : b.E1) - b.E0);
    int idx[C];
#pragma unroll
    for (int c = 0; c < C; ++c) idx[c] = pe[c].idx & ~kLastFlag;
    float y[C][NR];
#pragma unroll
    for (int c = 0; c < C; ++c) L::load(y[c], ra.at(Y, idx[c]), gl, f);
    float mine = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (e + c == row_end && e + c < e_end) {  // entry e+c opens the next segment
#pragma unroll
        for (int i = 0; i < NR; ++i) x[i] = xn[i];
        r = rn;
        row_end = s_ptr[r + 1];
        if (row_end < e_end) {
          rn = r + 1;
          while (s_ptr[rn + 1] <= row_end) ++rn;
          L::load(xn, ra.at(Xrow, b.seg_lb + rn), gl, f);
        }
      }
      const float d = group_sum<LPR>(dot<NR>(x, y[c]));
      if (gl == c) mine = d;
    }
    if (gl < C && e + gl < e_end) out(e - b.E0 + gl, mine);
    e += C;
  }
}

// flat_sddmm with the row operands of the tile staged in shared memory (s_x: [nseg][f], f = 4 *
// L::F4 floats per row) and flagged segment ends: a row switch is two LDS.128 instead of a
// register copy of a prefetched row.  C == 4 uses the transposed reduction.
template <class L, int C, class E, class Out>
__device__ __forceinline__ void flat_sddmm_sx(const RowBlock& b, const int* s_ptr, const int* s_next,
                                              const E* s_e, const RowAddr<L>& ra, const float* s_x,
                                              const char* Y, int gl, int f, Out out) {
  constexpr int NR = L::NR, LPR = L::LPR;
  static_assert(L::kVec && C <= LPR, "vector layouts only; one result lane per entry in flight");
  int e = b.e;
  const int e_end = b.e_end;
  const float* xl = s_x + 4 * gl;  // this lane's float4 column of every staged row
  float x[NR];
  zero(x);
  int r = 0;
  if (e < e_end) {
    r = find_row(s_ptr, b.nseg, e);
    L::load_smem(x, xl + r * f);
  }
  while (__any_sync(kFull, e < e_end)) {
    // a group that has run out of entries parks on the padding behind the tile
    const E* pe = s_e + ((e < e_end ? e : b.E1) - b.E0);
    int raw[C];
#pragma unroll
    for (int c = 0; c < C; ++c) raw[c] = pe[c].idx;
    float y[C][NR];
#pragma unroll
    for (int c = 0; c < C; ++c) L::load(y[c], ra.at(Y, raw[c] & ~kLastFlag), gl, f);
    float d[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      d[c] = dot<NR>(x, y[c]);
      if (raw[c] < 0) {  // last entry of segment r: switch to the next segment that has entries
        r = s_next[r];
        if (r < b.nseg) L::load_smem(x, xl + r * f);
      }
    }
    if constexpr (C == 4 && LPR >= 4) {
      const float t = reduce4_transposed<LPR>(d, gl);
      if (gl < 4 && e + perm4(gl) < e_end) out(e - b.E0 + perm4(gl), t);
    } else {
      float mine = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float t = group_sum<LPR>(d[c]);
        if (gl == c) mine = t;
      }
      if (gl < C && e + gl < e_end) out(e - b.E0 + gl, mine);
    }
    e += C;
  }
}

}  // namespace dfgnn
