// staged.cuh -- building blocks of the "staged tile" kernels (short-row graphs).
//
// The segmented row-block kernels (rowblock.cuh) carry an online softmax through
// every row piece; on graphs whose rows are a handful of entries long (arxiv-,
// cora-, PascalVOC-shaped: mean degree 4-7) that bookkeeping costs more issue slots
// than the gathers themselves.  The staged kernels split a CTA's tile of <= kStageCap
// entries into three phases with the per-entry scalars held in shared memory:
//
//   A  entry-parallel  one thread per entry: neighbour index, score / probability
//      (GAT), or   flat_sddmm: one lane group per entry, dot of the row operand
//      with the gathered neighbour row (GT scores, dA, GAT g)
//   B  segment-parallel  one lane group per row: max / sum / normalise in smem
//   C  flat_spmm  every lane group walks an EQUAL slice of the tile's entries and
//      accumulates weight * gathered row; a row boundary inside the slice costs one
//      short divergent flush, rows cut by a slice boundary are merged through the
//      same shared-memory slots as in rowblock.cuh (deterministic order).
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

// out[seg] (+)= sum_e w[e] * X[idx[e], :] over the staged tile; NOPS = 2 walks two operand
// matrices with two weight arrays at once; SCALAR also sums s_sc[e] per segment.
// store(seg, scalar, acc[NOPS*NR]) is called for segments finished inside one slice;
// split segments are left in the slots (merge with sum_merge_slots after a __syncthreads()).
template <class L, int C, int NOPS, bool SCALAR, class Store>
__device__ __forceinline__ void flat_spmm(const RowBlock& b, const int* s_ptr, const int* s_idx,
                                          const float* s_w0, const float* s_w1, const float* s_sc,
                                          const RowAddr<L>& ra, const char* X0, const char* X1,
                                          float* s_slot, int vw, int gl, int f, Store store) {
  constexpr int NR = L::NR, NV = NOPS * NR, LPR = L::LPR;
  int e = b.e;
  const int e_end = b.e_end;
  if (e >= e_end) return;
  int r = find_row(s_ptr, b.nseg, e);
  int row_end = s_ptr[r + 1];
  int pend = min(row_end, e_end);
  bool head = e > s_ptr[r];
  float acc[NV], sc = 0.f;
  zero(acc);
  const int last = e_end - 1 - b.E0;
  for (; e < e_end; e += C) {
    int idx[C];
    float w0[C], w1[NOPS == 2 ? C : 1], scc[SCALAR ? C : 1];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int k = min(e - b.E0 + c, last);
      const bool ok = e + c < e_end;
      idx[c] = s_idx[k];
      w0[c] = ok ? s_w0[k] : 0.f;
      if (NOPS == 2) w1[c] = ok ? s_w1[k] : 0.f;
      if (SCALAR) scc[c] = ok ? s_sc[k] : 0.f;
    }
    float v0[C][NR], v1[NOPS == 2 ? C : 1][NR];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      L::load(v0[c], ra.at(X0, idx[c]), gl, f);
      if (NOPS == 2) L::load(v1[c], ra.at(X1, idx[c]), gl, f);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        acc[i] = fmaf(w0[c], v0[c][i], acc[i]);
        if (NOPS == 2) acc[NR + i] = fmaf(w1[c], v1[c][i], acc[NR + i]);
      }
      if (SCALAR) sc += scc[c];
      if (e + c + 1 == pend) {  // last entry of a piece (never true for the padded tail)
        const bool complete = pend == row_end;
        if (!head && complete) {
          store(r, sc, acc);
        } else {
          Slot<NV, LPR> sl(s_slot, vw, head ? 0 : 1);
#pragma unroll
          for (int i = 0; i < NV; ++i) sl.v(i, gl) = acc[i];
          if (gl == 0) { sl.a() = sc; sl.set_seg(r); }
        }
        zero(acc);
        sc = 0.f;
        head = false;
        if (pend < e_end) {
          while (s_ptr[r + 1] <= pend) ++r;
          row_end = s_ptr[r + 1];
          pend = min(row_end, e_end);
        }
      }
    }
  }
}

// s_out[e] = <Xrow[seg(e), :], Y[idx[e], :]> for every staged entry (mul_into: s_out[e] *= ...).
// One lane group per entry, C entries in flight; the row operand of the NEXT segment is
// prefetched into registers when a segment starts.  Warp-converged (shuffles inside).
template <class L, int C>
__device__ __forceinline__ void flat_sddmm(const RowBlock& b, const int* s_ptr, const int* s_idx,
                                           const RowAddr<L>& ra, const char* Xrow, const char* Y,
                                           float* s_out, bool mul_into, int gl, int f) {
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
  const int last = max(e_end - 1 - b.E0, 0);
  while (__any_sync(kFull, e < e_end)) {
    int idx[C];
#pragma unroll
    for (int c = 0; c < C; ++c) idx[c] = s_idx[min(e - b.E0 + c, last)];
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
    if (gl < C && e + gl < e_end) {
      const int k = e - b.E0 + gl;
      s_out[k] = mul_into ? s_out[k] * mine : mine;
    }
    e += C;
  }
}

}  // namespace dfgnn
