// staged_gat.cuh -- staged-tile GAT kernels (see staged.cuh); same maths and parameter
// blocks as gat_fwd_kernel / gat_bwd_row_kernel / gat_bwd_col_kernel.  A CTA whose tile
// holds more than kStageCap entries returns at once: the row-block kernel launched right
// behind it (params.cap = kStageCap) processes exactly those tiles.
#pragma once

#include "bwd_kernels.cuh"
#include "fwd_kernels.cuh"
#include "staged.cuh"

namespace dfgnn {

#ifndef DFGNN_STAGE_WARPS
#define DFGNN_STAGE_WARPS 24  // resident warps per SM the short-row staged kernels are compiled for
#endif
constexpr int kEPT = kStageCap / (kNW * 32);  // staged entries per thread
static_assert(kMaxRB <= kNW * 32, "one thread per row in the prologues");

constexpr int kShortRow = 32;  // rows up to this length are reduced by ONE thread (smem only)

// Per-row passes over the staged tile (shared memory only).  Warps [0, kNW/2) take the short
// rows, one THREAD per row: short_row(r, rs, re).  Warps [kNW/2, kNW) take the rows longer than
// kShortRow, one WARP per row: long_row(r, rs, re, lane) is called by all 32 lanes and may use
// warp_sum / warp_max.  rs / re are tile-relative entry positions.
static_assert(kNW * 32 >= 2 * kMaxRB, "half of the CTA's threads cover the rows of a tile");
template <class Short, class Long>
__device__ __forceinline__ void per_row(const RowBlock& b, const int* s_rp, Short short_row,
                                        Long long_row) {
  constexpr int kHalf = kNW * 32 / 2;
  const int lane = threadIdx.x & 31;
  if ((int)threadIdx.x < kHalf) {
    const int r = threadIdx.x;
    if (r < b.nseg) {
      const int rs = s_rp[r] - b.E0, re = s_rp[r + 1] - b.E0;
      if (re - rs <= kShortRow) short_row(r, rs, re);
    }
  } else {
    const int r0 = threadIdx.x - kHalf - lane;  // first row of this warp's 32
    const int r = r0 + lane;
    const bool is_long = r < b.nseg && (s_rp[r + 1] - s_rp[r]) > kShortRow;
    unsigned m = __ballot_sync(kFull, is_long);
    while (m) {
      const int rr = r0 + __ffs(m) - 1;
      m &= m - 1;
      long_row(rr, s_rp[rr] - b.E0, s_rp[rr + 1] - b.E0, lane);
    }
  }
}

// forward: (A) every thread stages {neighbour, attn_col[neighbour]} of its entries; (B) per
// row: leakyrelu scores -> max -> exp weights, sum; (C) flat SpMM over feat, scaled by 1/sum
// at the row flush.
template <class L, int C>
__global__ void __launch_bounds__(kNW * 32, (L::NR <= 8 ? DFGNN_STAGE_WARPS : 16) / kNW) gat_fwd_staged_kernel(const GatFwdParams p) {
  constexpr int NR = L::NR, LPR = L::LPR, G = L::G, VW = kNW * G;
  __shared__ int s_rp[kMaxRB + 1], s_next[kMaxRB];
  __shared__ float s_inv[kMaxRB], s_ar[kMaxRB];
  __shared__ Ent1 s_e[kStageCap + kStagePad];
  extern __shared__ float s_slot[];

  allow_dependent_launch();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int grp = lane / LPR, gl = lane % LPR, vw = w * G + grp;
  const int hid = blockIdx.y, h = p.h, f = p.f;
  const bool use_mask = p.emask != nullptr;
  const float keep_scale = 1.f / (1.f - p.drop);
  const RowAddr<L> ra(h, f, hid, gl);
  const char* Fb = ra.base(p.feat);
  char* Ob = ra.base(p.out);
  const float* acb = p.ac + hid;

  slots_clear<NR, LPR>(s_slot, vw, gl);
  const RowBlock b = rowblock_init<G, C>(s_rp, p.row_ptr, p.m, p.rb, vw);
  const int ne = b.E1 - b.E0;
  if (ne > kStageCap) return;
  stage_pad(s_e, ne);
  // A: all index loads are issued before the dependent gathers (one DRAM + one L2 latency per CTA)
  {
    const int* colp = p.col_ind + b.E0;
    int cols[kEPT];
    float acv[kEPT];
#pragma unroll
    for (int k = 0; k < kEPT; ++k) {
      if (k * (kNW * 32) >= ne) break;  // uniform: no entry of the tile in this slot
      const int i = threadIdx.x + k * (kNW * 32);
      cols[k] = i < ne ? __ldg(colp + i) : 0;
    }
    if (threadIdx.x < b.nseg) s_ar[threadIdx.x] = __ldg(p.ar + (size_t)(b.seg_lb + threadIdx.x) * h + hid);
#pragma unroll
    for (int k = 0; k < kEPT; ++k) {
      if (k * (kNW * 32) >= ne) break;  // uniform: no entry of the tile in this slot
      const int i = threadIdx.x + k * (kNW * 32);
      acv[k] = i < ne ? __ldg(acb + (size_t)(unsigned)cols[k] * (unsigned)h) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < kEPT; ++k) {
      if (k * (kNW * 32) >= ne) break;  // uniform: no entry of the tile in this slot
      const int i = threadIdx.x + k * (kNW * 32);
      if (i < ne) { Ent1 en; en.idx = cols[k]; en.w = acv[k]; s_e[i] = en; }
    }
  }
  __syncthreads();

  // B: softmax statistics; s_e[i].w becomes exp(x - max) (times the dropout keep factor)
  auto weight = [&](int i, float x, float mx, float& l) {
    float pe = fast_exp(x - mx);
    l += pe;
    if (use_mask) {  // dropout on the attention weights, not on the normaliser
      const size_t eid = (size_t)(b.E0 + i) * h + hid;
      const float u = p.drop > 0.f ? uniform01(p.seed, eid) : 1.f;
      p.emask[eid] = u;
      pe = (u > p.drop) ? pe * keep_scale : 0.f;
    }
    s_e[i].w = pe;
  };
  auto finish_row = [&](int r, float mx, float l) {
    s_inv[r] = l > 0.f ? fast_rcp(l) : 0.f;
    if (p.emax) {  // saved for backward (fused_gatconv_kernel.cu:66-68, 89-91)
      const size_t node = (size_t)(b.seg_lb + r) * h + hid;
      p.emax[node] = l > 0.f ? mx : -1e38f;
      p.esum[node] = l;
    }
  };
  per_row(
      b, s_rp,
      [&](int r, int rs, int re) {
        const float ar_i = s_ar[r];
        float mx = kNeg, l = 0.f;
        for (int i = rs; i < re; ++i) {
          const float x = leaky(ar_i + s_e[i].w, p.slope);
          s_e[i].w = x;
          mx = fmaxf(mx, x);
        }
        for (int i = rs; i < re; ++i) weight(i, s_e[i].w, mx, l);
        finish_row(r, mx, l);
      },
      [&](int r, int rs, int re, int ln) {
        const float ar_i = s_ar[r];
        float mx = kNeg, l = 0.f;
        for (int i = rs + ln; i < re; i += 32) {
          const float x = leaky(ar_i + s_e[i].w, p.slope);
          s_e[i].w = x;
          mx = fmaxf(mx, x);
        }
        mx = warp_max(mx);
        for (int i = rs + ln; i < re; i += 32) weight(i, s_e[i].w, mx, l);
        l = warp_sum(l);
        if (ln == 0) finish_row(r, mx, l);
      });
  for (int r = vw; r < b.nseg; r += VW)
    if (s_rp[r + 1] == s_rp[r]) {  // rows without edges produce zeros
      float z[NR];
      zero(z);
      L::store(ra.at(Ob, b.seg_lb + r), z, gl, f);
    }
  __syncthreads();
  stage_mark_ends(b, s_rp, s_e, s_next);
  __syncthreads();

  // C: aggregation
  auto store = [&](int r, float, float (&acc)[NR]) {
    const float inv = s_inv[r];
#pragma unroll
    for (int i = 0; i < NR; ++i) acc[i] *= inv;
    L::store(ra.at(Ob, b.seg_lb + r), acc, gl, f);
  };
  flat_spmm<L, C, kOneOp>(b, s_rp, s_next, s_e, ra, Fb, nullptr, s_slot, vw, gl, f, store);
  __syncthreads();
  sum_merge_slots<NR, LPR, G>(s_slot, vw, gl, store);
}

// backward, row side: (A) stage {neighbour, attn_col[neighbour], keep factor}, dO rows -> smem;
// (C) flat SDDMM g = <dO_i, feat_j>; (D) per row -> p_e, lrelu', w, de, grad_attn_row, and the
// packed scratch {de_e, keep-scaled p_e} the column side reads.
// Entry fields: w = attn_col[j], then p_e; w1 = lrelu'(x_e); aux = keep_e / (1 - drop); s_g = g_e.
template <class L, int C>
__global__ void __launch_bounds__(kNW * 32, (L::NR <= 8 ? DFGNN_STAGE_WARPS : 16) / kNW) gat_bwd_row_staged_kernel(const GatBwdParams p) {
  constexpr int LPR = L::LPR, G = L::G;
  __shared__ int s_rp[kMaxRB + 1];
  __shared__ float s_ar[kMaxRB], s_mx[kMaxRB], s_inv[kMaxRB];
  __shared__ Ent2 s_e[kStageCap + kStagePad];
  __shared__ float s_g[kStageCap];
  __shared__ int s_next[kMaxRB];
  extern __shared__ float4 s_x4[];  // dO rows of the tile (kStageX)
  constexpr bool kStageX = stage_x<L>();

  allow_dependent_launch();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int grp = lane / LPR, gl = lane % LPR, vw = w * G + grp;
  const int hid = blockIdx.y, h = p.h, f = p.f;
  const float keep_scale = 1.f / (1.f - p.drop);
  const RowAddr<L> ra(h, f, hid, gl);
  const char* Gb = ra.base(p.dO);
  const char* Fb = ra.base(p.feat);
  const float* acb = p.ac + hid;
  float2* scratch = reinterpret_cast<float2*>(p.grad_edge);

  const RowBlock b = rowblock_init<G, C>(s_rp, p.row_ptr, p.m, p.rb, vw);
  const int ne = b.E1 - b.E0;
  if (ne > kStageCap) return;
  stage_pad(s_e, ne);
  if constexpr (kStageX) {  // dO rows of the tile -> shared memory, asynchronously (LDGSTS)
    constexpr int F4 = L::F4 > 0 ? L::F4 : 1;
    const float4* src = reinterpret_cast<const float4*>(p.dO);
    for (int i = threadIdx.x; i < b.nseg * F4; i += kNW * 32) {
      const int r = i / F4, k = i % F4;
      cp_async16(s_x4 + i, src + ((size_t)(b.seg_lb + r) * h + hid) * F4 + k);
    }
  }
  {
    const int* colp = p.col_ind + b.E0;
    int cols[kEPT];
    float acv[kEPT], kmv[kEPT];
#pragma unroll
    for (int k = 0; k < kEPT; ++k) {
      if (k * (kNW * 32) >= ne) break;  // uniform: no entry of the tile in this slot
      const int i = threadIdx.x + k * (kNW * 32);
      cols[k] = i < ne ? __ldg(colp + i) : 0;
      kmv[k] = 1.f;  // keep_e / (1 - drop)
      if (p.emask && i < ne)
        kmv[k] = (__ldg(p.emask + (size_t)(b.E0 + i) * h + hid) > p.drop) ? keep_scale : 0.f;
    }
    if (threadIdx.x < b.nseg) {
      const size_t node = (size_t)(b.seg_lb + threadIdx.x) * h + hid;
      s_ar[threadIdx.x] = __ldg(p.ar + node);
      s_mx[threadIdx.x] = __ldg(p.emax + node);
      s_inv[threadIdx.x] = fast_rcp(__ldg(p.esum + node));  // unused for rows without edges
    }
#pragma unroll
    for (int k = 0; k < kEPT; ++k) {
      if (k * (kNW * 32) >= ne) break;  // uniform: no entry of the tile in this slot
      const int i = threadIdx.x + k * (kNW * 32);
      acv[k] = i < ne ? __ldg(acb + (size_t)(unsigned)cols[k] * (unsigned)h) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < kEPT; ++k) {
      if (k * (kNW * 32) >= ne) break;  // uniform: no entry of the tile in this slot
      const int i = threadIdx.x + k * (kNW * 32);
      if (i < ne) { Ent2 en; en.idx = cols[k]; en.w = acv[k]; en.w1 = 0.f; en.aux = kmv[k]; s_e[i] = en; }
    }
  }
  if constexpr (kStageX) cp_async_wait_all();
  __syncthreads();
  if constexpr (kStageX) {
    stage_mark_ends(b, s_rp, s_e, s_next);
    __syncthreads();
    if (ne > 0)
      flat_sddmm_sx<L, C>(b, s_rp, s_next, s_e, ra, reinterpret_cast<const float*>(s_x4), Fb, gl, f,
                          [&](int k, float d) { s_g[k] = d; });
  } else {
    if (ne > 0)
      flat_sddmm<L, C>(b, s_rp, s_e, ra, Gb, Fb, gl, f, [&](int k, float d) { s_g[k] = d; });
  }
  __syncthreads();

  // Per row: p_e = exp(x_e - max_i) / sum_i, x_e = leakyrelu(ar_i + ac_j) from the staged ac_j;
  // t_e = keep-scaled p_e g_e, w_i = sum_e t_e, de_e = (t_e - w_i p_e) * lrelu'(x_e)
  // (fused_gatconv_kernel.cu:830-864).  The probabilities are only needed here, so they are
  // not a phase of their own.
  auto prob = [&](int i, float ar_i, float mx, float inv, float& wsum) {
    const float x = leaky(ar_i + s_e[i].w, p.slope);
    const float pe = fast_exp(x - mx) * inv;
    s_e[i].w = pe;
    s_e[i].w1 = x < 0.f ? p.slope : 1.f;
    wsum = fmaf(pe * s_e[i].aux, s_g[i], wsum);
  };
  auto edge_grad = [&](int i, float wsum, float& rsum) {
    const Ent2 en = s_e[i];
    const float pk = en.w * en.aux;
    const float de = (pk * s_g[i] - wsum * en.w) * en.w1;
    scratch[(size_t)(b.E0 + i) * h + hid] = make_float2(de, pk);
    rsum += de;
  };
  per_row(
      b, s_rp,
      [&](int r, int rs, int re) {
        const float ar_i = s_ar[r], mx = s_mx[r], inv = s_inv[r];
        float wsum = 0.f, rsum = 0.f;
        for (int i = rs; i < re; ++i) prob(i, ar_i, mx, inv, wsum);
        for (int i = rs; i < re; ++i) edge_grad(i, wsum, rsum);
        p.grad_ar[(size_t)(b.seg_lb + r) * h + hid] = rsum;
      },
      [&](int r, int rs, int re, int ln) {
        const float ar_i = s_ar[r], mx = s_mx[r], inv = s_inv[r];
        float wsum = 0.f, rsum = 0.f;
        for (int i = rs + ln; i < re; i += 32) prob(i, ar_i, mx, inv, wsum);
        wsum = warp_sum(wsum);
        for (int i = rs + ln; i < re; i += 32) edge_grad(i, wsum, rsum);
        rsum = warp_sum(rsum);
        if (ln == 0) p.grad_ar[(size_t)(b.seg_lb + r) * h + hid] = rsum;
      });
}

// backward, column side (CSC): (A) every thread stages {row, keep-scaled p_e, de_e} of its
// entries from the packed scratch of the row side; (C) flat SpMM over dO carrying sum(de).
template <class L, int C>
__global__ void __launch_bounds__(kNW * 32, (L::NR <= 8 ? DFGNN_STAGE_WARPS : 16) / kNW) gat_bwd_col_staged_kernel(const GatBwdParams p) {
  constexpr int NR = L::NR, LPR = L::LPR, G = L::G, VW = kNW * G;
  __shared__ int s_cp[kMaxRB + 1], s_next[kMaxRB];
  __shared__ Ent2 s_e[kStageCap + kStagePad];
  extern __shared__ float s_slot[];

  allow_dependent_launch();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int grp = lane / LPR, gl = lane % LPR, vw = w * G + grp;
  const int hid = blockIdx.y, h = p.h, f = p.f;
  const RowAddr<L> ra(h, f, hid, gl);
  const char* Gb = ra.base(p.dO);
  char* GFb = ra.base(p.grad_feat);
  const float2* scratch = reinterpret_cast<const float2*>(p.grad_edge);

  slots_clear<NR, LPR>(s_slot, vw, gl);
  const RowBlock b = rowblock_init<G, C>(s_cp, p.col_ptr, p.n, p.rb_col, vw);
  const int ne = b.E1 - b.E0;
  if (ne > kStageCap) return;
  stage_pad(s_e, ne);
  {
    int rid[kEPT], eid[kEPT];
#pragma unroll
    for (int k = 0; k < kEPT; ++k) {
      if (k * (kNW * 32) >= ne) break;  // uniform: no entry of the tile in this slot
      const int i = threadIdx.x + k * (kNW * 32);
      rid[k] = i < ne ? __ldg(p.row_ind + b.E0 + i) : 0;
      eid[k] = i < ne ? __ldg(p.permute + b.E0 + i) : 0;
    }
    float2 dp[kEPT];
#pragma unroll
    for (int k = 0; k < kEPT; ++k) {
      if (k * (kNW * 32) >= ne) break;  // uniform: no entry of the tile in this slot
      const int i = threadIdx.x + k * (kNW * 32);
      dp[k] = i < ne ? __ldg(scratch + (size_t)eid[k] * h + hid) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < kEPT; ++k) {
      if (k * (kNW * 32) >= ne) break;  // uniform: no entry of the tile in this slot
      const int i = threadIdx.x + k * (kNW * 32);
      if (i < ne) { Ent2 en; en.idx = rid[k]; en.w = dp[k].y; en.w1 = dp[k].x; en.aux = 0.f; s_e[i] = en; }
    }
  }
  auto store = [&](int c, float dac, float (&acc)[NR]) {
    L::store(ra.at(GFb, b.seg_lb + c), acc, gl, f);
    if (gl == 0) p.grad_ac[(size_t)(b.seg_lb + c) * h + hid] = dac;
  };
  for (int c = vw; c < b.nseg; c += VW)
    if (s_cp[c + 1] == s_cp[c]) {  // columns without entries
      float z[NR];
      zero(z);
      store(c, 0.f, z);
    }
  __syncthreads();
  stage_mark_ends(b, s_cp, s_e, s_next);
  __syncthreads();

  flat_spmm<L, C, kOneOpScalar>(b, s_cp, s_next, s_e, ra, Gb, nullptr, s_slot, vw, gl, f, store);
  __syncthreads();
  sum_merge_slots<NR, LPR, G>(s_slot, vw, gl, store);
}

}  // namespace dfgnn
