// staged_gat.cuh -- staged-tile GAT kernels (see staged.cuh); same maths and parameter
// blocks as gat_fwd_kernel / gat_bwd_row_kernel / gat_bwd_col_kernel.  A CTA whose tile
// holds more than kStageCap entries returns at once: the row-block kernel launched right
// behind it (params.cap = kStageCap) processes exactly those tiles.
#pragma once

#include "bwd_kernels.cuh"
#include "fwd_kernels.cuh"
#include "staged.cuh"

namespace dfgnn {

// forward: A thread per edge -> leakyrelu score; B group per row -> max, sum, exp weights;
// C flat SpMM over feat, scaled by 1/sum at the row flush.
template <class L, int C>
__global__ void __launch_bounds__(kNW * 32, (L::NR <= 8 ? 24 : 16) / kNW) gat_fwd_staged_kernel(const GatFwdParams p) {
  constexpr int NR = L::NR, LPR = L::LPR, G = L::G, VW = kNW * G;
  __shared__ int s_rp[kMaxRB + 1];
  __shared__ float s_ar[kMaxRB], s_inv[kMaxRB];
  __shared__ int s_idx[kStageCap];
  __shared__ float s_w[kStageCap];
  extern __shared__ float s_slot[];

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
  {
    const int lb = blockIdx.x * p.rb, ns = min(p.rb, p.m - lb);
    for (int i = threadIdx.x; i < ns; i += kNW * 32) s_ar[i] = __ldg(p.ar + (size_t)(lb + i) * h + hid);
  }
  const RowBlock b = rowblock_init<G>(s_rp, p.row_ptr, p.m, p.rb, vw);
  const int ne = b.E1 - b.E0;
  if (ne > kStageCap) return;

  // A: scores
  for (int i = threadIdx.x; i < ne; i += kNW * 32) {
    const int col = __ldg(p.col_ind + b.E0 + i);
    const int r = find_row(s_rp, b.nseg, b.E0 + i);
    s_idx[i] = col;
    s_w[i] = leaky(s_ar[r] + __ldg(acb + (size_t)(unsigned)col * (unsigned)h), p.slope);
  }
  __syncthreads();

  // B: per-row softmax statistics; s_w becomes exp(x - max) (times the dropout keep factor)
  for (int r0 = w * G; r0 < b.nseg; r0 += VW) {
    const int rr = r0 + grp;
    const bool valid = rr < b.nseg;
    const int rs = valid ? s_rp[rr] - b.E0 : 0, re = valid ? s_rp[rr + 1] - b.E0 : 0;
    float mx = kNeg;
    for (int i = rs + gl; __any_sync(kFull, i < re); i += LPR)
      if (i < re) mx = fmaxf(mx, s_w[i]);
    mx = group_max<LPR>(mx);
    float l = 0.f;
    for (int i = rs + gl; __any_sync(kFull, i < re); i += LPR)
      if (i < re) {
        float pe = fast_exp(s_w[i] - mx);
        l += pe;
        if (use_mask) {
          const size_t eid = (size_t)(b.E0 + i) * h + hid;
          const float u = p.drop > 0.f ? uniform01(p.seed, eid) : 1.f;
          p.emask[eid] = u;
          pe = (u > p.drop) ? pe * keep_scale : 0.f;
        }
        s_w[i] = pe;
      }
    l = group_sum<LPR>(l);
    if (valid) {
      const size_t node = (size_t)(b.seg_lb + rr) * h + hid;
      if (gl == 0) {
        s_inv[rr] = l > 0.f ? 1.f / l : 0.f;
        if (p.emax) {  // saved for backward (fused_gatconv_kernel.cu:66-68, 89-91)
          p.emax[node] = l > 0.f ? mx : -1e38f;
          p.esum[node] = l;
        }
      }
      if (re == rs) {  // rows without edges produce zeros
        float z[NR];
        zero(z);
        L::store(ra.at(Ob, b.seg_lb + rr), z, gl, f);
      }
    }
  }
  __syncthreads();

  // C: aggregation
  auto store = [&](int r, float, float (&acc)[NR]) {
    const float inv = s_inv[r];
#pragma unroll
    for (int i = 0; i < NR; ++i) acc[i] *= inv;
    L::store(ra.at(Ob, b.seg_lb + r), acc, gl, f);
  };
  flat_spmm<L, C, 1, false>(b, s_rp, s_idx, s_w, nullptr, nullptr, ra, Fb, nullptr, s_slot, vw, gl,
                            f, store);
  __syncthreads();
  sum_merge_slots<NR, LPR, G>(s_slot, vw, gl, store);
}

// backward, row side: A thread per edge -> p, slope factor, keep factor; flat SDDMM
// g = <dO_i, feat_j>; B group per row -> w, de, grad_attn_row.
template <class L, int C>
__global__ void __launch_bounds__(kNW * 32, (L::NR <= 8 ? 24 : 16) / kNW) gat_bwd_row_staged_kernel(const GatBwdParams p) {
  constexpr int LPR = L::LPR, G = L::G, VW = kNW * G;
  __shared__ int s_rp[kMaxRB + 1];
  __shared__ float s_ar[kMaxRB], s_mx[kMaxRB], s_inv[kMaxRB];
  __shared__ int s_idx[kStageCap];
  __shared__ float s_p[kStageCap], s_sl[kStageCap], s_g[kStageCap];

  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int grp = lane / LPR, gl = lane % LPR, vw = w * G + grp;
  const int hid = blockIdx.y, h = p.h, f = p.f;
  const float keep_scale = 1.f / (1.f - p.drop);
  const RowAddr<L> ra(h, f, hid, gl);
  const char* Gb = ra.base(p.dO);
  const char* Fb = ra.base(p.feat);
  const float* acb = p.ac + hid;

  {
    const int lb = blockIdx.x * p.rb, ns = min(p.rb, p.m - lb);
    for (int i = threadIdx.x; i < ns; i += kNW * 32) {
      const size_t node = (size_t)(lb + i) * h + hid;
      s_ar[i] = __ldg(p.ar + node);
      s_mx[i] = __ldg(p.emax + node);
      s_inv[i] = 1.f / __ldg(p.esum + node);  // unused for rows without edges
    }
  }
  const RowBlock b = rowblock_init<G>(s_rp, p.row_ptr, p.m, p.rb, vw);
  const int ne = b.E1 - b.E0;
  if (ne > kStageCap) return;

  for (int i = threadIdx.x; i < ne; i += kNW * 32) {
    const int col = __ldg(p.col_ind + b.E0 + i);
    const int r = find_row(s_rp, b.nseg, b.E0 + i);
    const float x = leaky(s_ar[r] + __ldg(acb + (size_t)(unsigned)col * (unsigned)h), p.slope);
    s_idx[i] = col;
    s_p[i] = fast_exp(x - s_mx[r]) * s_inv[r];
    s_sl[i] = x < 0.f ? p.slope : 1.f;
    float km = 1.f;  // keep_e / (1 - drop)
    if (p.emask) km = (__ldg(p.emask + (size_t)(b.E0 + i) * h + hid) > p.drop) ? keep_scale : 0.f;
    s_g[i] = km;
  }
  __syncthreads();

  if (ne > 0) flat_sddmm<L, C>(b, s_rp, s_idx, ra, Gb, Fb, s_g, true, gl, f);
  __syncthreads();

  // de_e = (p_e g_e - w_i p_e) * lrelu'(x_e), w_i = sum_e p_e g_e, with g_e already keep-scaled
  // (fused_gatconv_kernel.cu:830-864)
  for (int r0 = w * G; r0 < b.nseg; r0 += VW) {
    const int rr = r0 + grp;
    const bool valid = rr < b.nseg;
    const int rs = valid ? s_rp[rr] - b.E0 : 0, re = valid ? s_rp[rr + 1] - b.E0 : 0;
    float wsum = 0.f;
    for (int i = rs + gl; __any_sync(kFull, i < re); i += LPR)
      if (i < re) wsum = fmaf(s_p[i], s_g[i], wsum);
    wsum = group_sum<LPR>(wsum);
    float rsum = 0.f;
    for (int i = rs + gl; __any_sync(kFull, i < re); i += LPR)
      if (i < re) {
        const float pe = s_p[i];
        const float de = (pe * s_g[i] - wsum * pe) * s_sl[i];
        p.grad_edge[(size_t)(b.E0 + i) * h + hid] = de;
        rsum += de;
      }
    rsum = group_sum<LPR>(rsum);
    if (valid && gl == 0) p.grad_ar[(size_t)(b.seg_lb + rr) * h + hid] = rsum;
  }
}

// backward, column side (CSC): A thread per entry -> p (keep-scaled), de; C flat SpMM over dO
// carrying sum(de) per column.
template <class L, int C>
__global__ void __launch_bounds__(kNW * 32, (L::NR <= 8 ? 24 : 16) / kNW) gat_bwd_col_staged_kernel(const GatBwdParams p) {
  constexpr int NR = L::NR, LPR = L::LPR, G = L::G, VW = kNW * G;
  __shared__ int s_cp[kMaxRB + 1];
  __shared__ float s_ac[kMaxRB];
  __shared__ int s_idx[kStageCap];
  __shared__ float s_w[kStageCap], s_de[kStageCap];
  extern __shared__ float s_slot[];

  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int grp = lane / LPR, gl = lane % LPR, vw = w * G + grp;
  const int hid = blockIdx.y, h = p.h, f = p.f;
  const float keep_scale = 1.f / (1.f - p.drop);
  const RowAddr<L> ra(h, f, hid, gl);
  const char* Gb = ra.base(p.dO);
  char* GFb = ra.base(p.grad_feat);

  slots_clear<NR, LPR>(s_slot, vw, gl);
  {
    const int lb = blockIdx.x * p.rb_col, ns = min(p.rb_col, p.n - lb);
    for (int i = threadIdx.x; i < ns; i += kNW * 32) s_ac[i] = __ldg(p.ac + (size_t)(lb + i) * h + hid);
  }
  const RowBlock b = rowblock_init<G>(s_cp, p.col_ptr, p.n, p.rb_col, vw);
  const int ne = b.E1 - b.E0;
  if (ne > kStageCap) return;

  for (int i = threadIdx.x; i < ne; i += kNW * 32) {
    const int rid = __ldg(p.row_ind + b.E0 + i);
    const size_t eid = (size_t)__ldg(p.permute + b.E0 + i) * h + hid;
    const size_t rn = (size_t)rid * h + hid;
    const int c = find_row(s_cp, b.nseg, b.E0 + i);
    const float x = leaky(__ldg(p.ar + rn) + s_ac[c], p.slope);
    float pe = fast_exp(x - __ldg(p.emax + rn)) / __ldg(p.esum + rn);
    if (p.emask) pe = (__ldg(p.emask + eid) > p.drop) ? pe * keep_scale : 0.f;
    s_idx[i] = rid;
    s_w[i] = pe;
    s_de[i] = __ldg(p.grad_edge + eid);
  }
  __syncthreads();

  auto store = [&](int c, float dac, float (&acc)[NR]) {
    L::store(ra.at(GFb, b.seg_lb + c), acc, gl, f);
    if (gl == 0) p.grad_ac[(size_t)(b.seg_lb + c) * h + hid] = dac;
  };
  for (int c = vw; c < b.nseg; c += VW)
    if (s_cp[c + 1] == s_cp[c]) {
      float z[NR];
      zero(z);
      store(c, 0.f, z);
    }
  flat_spmm<L, C, 1, true>(b, s_cp, s_idx, s_w, nullptr, s_de, ra, Gb, nullptr, s_slot, vw, gl, f,
                           store);
  __syncthreads();
  sum_merge_slots<NR, LPR, G>(s_slot, vw, gl, store);
}

}  // namespace dfgnn
