// fwd_kernels.cuh -- fused forward kernels: score -> online softmax -> aggregate.
//
//   dot_fwd_kernel<L, C, AGNN>   GT   : s_ij = <Q_i, K_j> * val_ij,  O_i = sum_j p_ij V_j
//                                AGNN : s_ij = <H_i, H_j> rn_i rn_j, O_i = sum_j p_ij H_j
//                                       (one gathered row serves score AND aggregation)
//   gat_fwd_kernel<L, C>         GAT  : s_ij = leakyrelu(ar_i + ac_j), O_i = sum_j keep_ij p_ij feat_j
//
// Maths spec: SURVEY.md section 8 (restating fused_gtconv_hyper.cu:31-163 and
// fused_gatconv_kernel.cu:24-125).  Schedule: rowblock.cuh (walk_pieces).
#pragma once

#include "rowblock.cuh"

namespace dfgnn {

struct DotFwdParams {
  int m, nnz, h, f, rb;
  const int* row_ptr;
  const int* col_ind;
  const float* val;  // [nnz] or null (GT)
  const float* Q;
  const float* K;
  const float* V;
  const float* rn;   // [m, h] inverse norms (AGNN) or null
  float* out;
  float* attn;       // [h, nnz] or null
  int cap = 0;       // > 0: process only tiles with more than `cap` entries (behind a staged kernel)
};

// fold (m2, l2, acc2) into (m, l, acc): the online-softmax merge, base-2 exponent domain
template <int NR>
__device__ __forceinline__ void softmax_merge2(float& m, float& l, float (&acc)[NR], float m2,
                                               float l2, const float (&acc2)[NR]) {
  const float mn = fmaxf(m, m2);
  const float sa = fast_exp2(m - mn), sb = fast_exp2(m2 - mn);
  l = l * sa + l2 * sb;
#pragma unroll
  for (int i = 0; i < NR; ++i) acc[i] = acc[i] * sa + acc2[i] * sb;
  m = mn;
}

// After the walk: every group that holds the FIRST piece of a split row folds the
// following groups' head pieces into it and finishes the row.  Group-local code
// (no shuffles): groups of a warp may take different trip counts.
template <int NR, int LPR, int G, class Fin>
__device__ __forceinline__ void softmax_merge_slots(float* s_slot, int vw, int gl, Fin fin) {
  Slot<NR, LPR> mine(s_slot, vw, 1);
  const int seg = mine.seg();
  if (seg < 0) return;
  float m = mine.a(), l = mine.b(), acc[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) acc[i] = mine.v(i, gl);
  for (int v2 = vw + 1; v2 < kNW * G; ++v2) {
    Slot<NR, LPR> s(s_slot, v2, 0);
    if (s.seg() != seg) break;
    float acc2[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) acc2[i] = s.v(i, gl);
    softmax_merge2<NR>(m, l, acc, s.a(), s.b(), acc2);
  }
  fin(seg, m, l, acc);
}

// Scores are kept in the base-2 exponent domain (Q is pre-multiplied by log2 e),
// so every exponential is one ex2.approx.
template <class L, int C, bool AGNN>
__global__ void __launch_bounds__(kNW * 32, (L::NR <= 8 ? DFGNN_GT_WARPS_SMALL : DFGNN_GT_WARPS) / kNW) dot_fwd_kernel(const DotFwdParams p) {
  constexpr int NR = L::NR, LPR = L::LPR, G = L::G, VW = kNW * G;
  constexpr int CH = ChunkOf<L>::kChunk;  // edges per index prefetch (<= LPR)
  static_assert(CH % C == 0 && CH <= LPR, "chunking");
  __shared__ int s_rp[kMaxRB + 1];
  __shared__ float s_m[kMaxRB], s_inv[kMaxRB];
  extern __shared__ float s_slot[];

  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int grp = lane / LPR, gl = lane % LPR, vw = w * G + grp;
  const int hid = blockIdx.y, h = p.h, f = p.f;
  const bool train = p.attn != nullptr;
  const bool use_w = AGNN || p.val != nullptr;
  float* attn = train ? p.attn + (size_t)hid * p.nnz : nullptr;
  const RowAddr<L> ra(h, f, hid, gl);
  const char* Qb = ra.base(p.Q);
  const char* Kb = ra.base(p.K);
  const char* Vb = ra.base(p.V);
  char* Ob = ra.base(p.out);

  slots_clear<NR, LPR>(s_slot, vw, gl);
  const RowBlock b = rowblock_init<G>(s_rp, p.row_ptr, p.m, p.rb, vw);
  if (p.cap > 0 && b.E1 - b.E0 <= p.cap) {  // not a big tile: nothing to do here
    dependency_wait();
    return;
  }

  auto finish = [&](int r, float m, float l, float (&acc)[NR]) {
    const float inv = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
    for (int i = 0; i < NR; ++i) acc[i] *= inv;
    L::store(ra.at(Ob, b.seg_lb + r), acc, gl, f);
    if (gl == 0) { s_m[r] = m; s_inv[r] = inv; }
  };

  // rows without edges produce zeros (fused_gtconv_hyper.cu:142-143)
  for (int r = vw; r < b.nseg; r += VW)
    if (s_rp[r + 1] == s_rp[r]) {
      float z[NR];
      zero(z);
      finish(r, kNeg, 0.f, z);
    }

  float q[NR], acc[NR];
  zero(q);
  zero(acc);
  float rn_i = 1.f, m_run = kNeg, l_run = 0.f;
  walk_pieces<CH>(
      b, s_rp,
      [&](int r) {
        L::load(q, ra.at(Qb, b.seg_lb + r), gl, f);
#pragma unroll
        for (int i = 0; i < NR; ++i) q[i] *= kLog2e;
        if (AGNN) rn_i = __ldg(p.rn + (size_t)(b.seg_lb + r) * h + hid);
        zero(acc);
        m_run = kNeg;
        l_run = 0.f;
      },
      [&](int base, int cnt) {
        int my_col = 0;
        float my_w = 1.f, my_sc = 0.f;
        if (gl < cnt) {
          my_col = __ldg(p.col_ind + base + gl);
          if (AGNN) my_w = __ldg(p.rn + (size_t)my_col * h + hid) * rn_i;
          else if (p.val) my_w = __ldg(p.val + base + gl);
        }
#pragma unroll
        for (int s = 0; s < CH; s += C) {
          if (s > 0 && !__any_sync(kFull, s < cnt)) break;
          float kk[AGNN ? 1 : C][NR], vv[C][NR], d[C];
          bool ok[C];
#pragma unroll
          for (int c = 0; c < C; ++c) {
            ok[c] = s + c < cnt;
            // beyond cnt: the chunk's first neighbour again (a valid, cached row; its weight is 0) --
            // cheaper than zeroing the registers and branching around the loads
            const int col = group_bcast<LPR>(my_col, ok[c] ? s + c : 0);
            if (!AGNN) L::load(kk[c], ra.at(Kb, col), gl, f);
            L::load(vv[c], ra.at(Vb, col), gl, f);
          }
          float cm = kNeg;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            float dc = AGNN ? dot<NR>(q, vv[c]) : dot<NR>(q, kk[AGNN ? 0 : c]);
            dc = group_sum<LPR>(dc);
            if (use_w) dc *= group_bcast<LPR>(my_w, s + c);
            if (gl == s + c) my_sc = dc;
            d[c] = dc;
            cm = ok[c] ? fmaxf(cm, dc) : cm;
          }
          const float m_new = fmaxf(m_run, cm);
          const float scale = fast_exp2(m_run - m_new);
#pragma unroll
          for (int i = 0; i < NR; ++i) acc[i] *= scale;
          float ps = 0.f;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float pc = ok[c] ? fast_exp2(d[c] - m_new) : 0.f;
            ps += pc;
#pragma unroll
            for (int i = 0; i < NR; ++i) acc[i] = fmaf(pc, vv[c][i], acc[i]);
          }
          l_run = fmaf(l_run, scale, ps);
          m_run = m_new;
        }
        if (train && gl < cnt) attn[base + gl] = my_sc;  // raw score, normalised below
      },
      [&](int r, bool first, bool last) {
        if (first && last) {
          finish(r, m_run, l_run, acc);
        } else {
          Slot<NR, LPR> sl(s_slot, vw, first ? 1 : 0);
#pragma unroll
          for (int i = 0; i < NR; ++i) sl.v(i, gl) = acc[i];
          if (gl == 0) { sl.a() = m_run; sl.b() = l_run; sl.set_seg(r); }
        }
      });
  __syncthreads();
  softmax_merge_slots<NR, LPR, G>(s_slot, vw, gl, finish);

  if (train) {  // scores -> probabilities (attn_edge of fused_gtconv_hyper.cu:146-149)
    __syncthreads();
    for (int i = b.E0 + threadIdx.x; i < b.E1; i += kNW * 32) {
      const int rr = find_row(s_rp, b.nseg, i);
      attn[i] = fast_exp2(attn[i] - s_m[rr]) * s_inv[rr];
    }
  }
  if (p.cap > 0) dependency_wait();  // see common.cuh
}

// ------------------------------------------------------------------------- //

struct GatFwdParams {
  int m, nnz, h, f, rb;
  const int* row_ptr;
  const int* col_ind;
  const float* ar;    // attn_row [m, h]
  const float* ac;    // attn_col [n, h]
  const float* feat;
  float slope, drop;
  uint64_t seed;
  float* out;
  float* emax;   // [m, h] or null
  float* esum;   // [m, h] or null
  float* emask;  // [nnz, h] or null
  int cap = 0;   // > 0: process only tiles with more than `cap` entries (behind a staged kernel)
};

// Natural-log domain (edge_max / edge_sum are returned to the caller).
template <int NR>
__device__ __forceinline__ void softmax_merge_e(float& m, float& l, float (&acc)[NR], float m2,
                                                float l2, const float (&acc2)[NR]) {
  const float mn = fmaxf(m, m2);
  const float sa = fast_exp(m - mn), sb = fast_exp(m2 - mn);
  l = l * sa + l2 * sb;
#pragma unroll
  for (int i = 0; i < NR; ++i) acc[i] = acc[i] * sa + acc2[i] * sb;
  m = mn;
}

template <class L, int C>
__global__ void __launch_bounds__(kNW * 32, (L::NR <= 8 ? 24 : 16) / kNW) gat_fwd_kernel(const GatFwdParams p) {
  constexpr int NR = L::NR, LPR = L::LPR, G = L::G, VW = kNW * G;
  static_assert(LPR % C == 0, "chunking");
  __shared__ int s_rp[kMaxRB + 1];
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

  if (small_tile_exit(p.row_ptr, p.m, p.rb, p.cap)) return;
  slots_clear<NR, LPR>(s_slot, vw, gl);
  const RowBlock b = rowblock_init<G>(s_rp, p.row_ptr, p.m, p.rb, vw);
  if (p.cap > 0 && b.E1 - b.E0 <= p.cap) {  // not a big tile: nothing to do here
    dependency_wait();
    return;
  }

  auto finish = [&](int r, float m, float l, float (&acc)[NR]) {
    const size_t node = (size_t)(b.seg_lb + r) * h + hid;
    const float inv = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
    for (int i = 0; i < NR; ++i) acc[i] *= inv;
    L::store(ra.at(Ob, b.seg_lb + r), acc, gl, f);
    if (gl == 0 && p.emax) {  // saved for backward (fused_gatconv_kernel.cu:66-68, 89-91)
      p.emax[node] = l > 0.f ? m : -1e38f;
      p.esum[node] = l;
    }
  };

  for (int r = vw; r < b.nseg; r += VW)
    if (s_rp[r + 1] == s_rp[r]) {
      float z[NR];
      zero(z);
      finish(r, kNeg, 0.f, z);
    }

  float ar_i = 0.f, acc[NR], m_run = kNeg, l_lane = 0.f;
  zero(acc);
  walk_pieces<LPR>(
      b, s_rp,
      [&](int r) {
        ar_i = __ldg(p.ar + (size_t)(b.seg_lb + r) * h + hid);
        zero(acc);
        m_run = kNeg;
        l_lane = 0.f;
      },
      [&](int base, int cnt) {
        int my_col = 0;
        float sc = kNeg;
        if (gl < cnt) {  // lane gl scores edge base + gl
          my_col = __ldg(p.col_ind + base + gl);
          sc = leaky(ar_i + __ldg(acb + (size_t)(unsigned)my_col * (unsigned)h), p.slope);
        }
        const float m_new = fmaxf(m_run, group_max<LPR>(sc));
        const float scale = fast_exp(m_run - m_new);
        float pe = gl < cnt ? fast_exp(sc - m_new) : 0.f;
        l_lane = fmaf(l_lane, scale, pe);
        m_run = m_new;
        if (use_mask && gl < cnt) {  // dropout on the attention weights
          const size_t eid = (size_t)(base + gl) * h + hid;
          // attn_drop == 0 keeps every edge: any mask value in (0, 1] says so, skip the generator
          const float u = p.drop > 0.f ? uniform01(p.seed, eid) : 1.f;
          p.emask[eid] = u;
          pe = (u > p.drop) ? pe * keep_scale : 0.f;
        }
#pragma unroll
        for (int i = 0; i < NR; ++i) acc[i] *= scale;
#pragma unroll
        for (int s = 0; s < LPR; s += C) {
          if (s > 0 && !__any_sync(kFull, s < cnt)) break;
          float vv[C][NR], pc[C];
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const int col = group_bcast<LPR>(my_col, s + c < cnt ? s + c : 0);  // beyond cnt: first neighbour again
            pc[c] = group_bcast<LPR>(pe, s + c);  // 0 beyond cnt
            L::load(vv[c], ra.at(Fb, col), gl, f);
          }
#pragma unroll
          for (int c = 0; c < C; ++c)
#pragma unroll
            for (int i = 0; i < NR; ++i) acc[i] = fmaf(pc[c], vv[c][i], acc[i]);
        }
      },
      [&](int r, bool first, bool last) {
        const float l_run = group_sum_local<LPR>(l_lane, lane);
        if (first && last) {
          finish(r, m_run, l_run, acc);
        } else {
          Slot<NR, LPR> sl(s_slot, vw, first ? 1 : 0);
#pragma unroll
          for (int i = 0; i < NR; ++i) sl.v(i, gl) = acc[i];
          if (gl == 0) { sl.a() = m_run; sl.b() = l_run; sl.set_seg(r); }
        }
      });
  __syncthreads();
  {
    Slot<NR, LPR> mine(s_slot, vw, 1);
    const int seg = mine.seg();
    if (seg >= 0) {
      float m = mine.a(), l = mine.b(), acc[NR];
#pragma unroll
      for (int i = 0; i < NR; ++i) acc[i] = mine.v(i, gl);
      for (int v2 = vw + 1; v2 < VW; ++v2) {
        Slot<NR, LPR> s(s_slot, v2, 0);
        if (s.seg() != seg) break;
        float acc2[NR];
#pragma unroll
        for (int i = 0; i < NR; ++i) acc2[i] = s.v(i, gl);
        softmax_merge_e<NR>(m, l, acc, s.a(), s.b(), acc2);
      }
      finish(seg, m, l, acc);
    }
  }
  if (p.cap > 0) dependency_wait();  // see common.cuh
}

// ------------------------------------------------------------------------- //
// Node-wise prologues                                                        //
// ------------------------------------------------------------------------- //

// attn_row[i,h] = <a_l[h], feat[i,h]>, attn_col = <a_r[h], feat[i,h]>
// (fused_gat_dot_attn_weight, fused_gatconv_hyper_v2.cu:212-250). Warp per (node, head).
static __global__ void __launch_bounds__(256) gat_attn_weight_kernel(int m, int h, int f,
                                                                     const float* __restrict__ a_l,
                                                                     const float* __restrict__ a_r,
                                                                     const float* __restrict__ feat,
                                                                     float* __restrict__ attn_row,
                                                                     float* __restrict__ attn_col) {
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= (long long)m * h) return;
  const int hid = (int)(wid % h);
  const float* x = feat + (size_t)wid * f;
  float r = 0.f, c = 0.f;
  if ((f & 3) == 0) {
    for (int j = lane; j < (f >> 2); j += 32) {
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x) + j);
      const float4 lv = __ldg(reinterpret_cast<const float4*>(a_l + (size_t)hid * f) + j);
      const float4 rv = __ldg(reinterpret_cast<const float4*>(a_r + (size_t)hid * f) + j);
      r += xv.x * lv.x + xv.y * lv.y + xv.z * lv.z + xv.w * lv.w;
      c += xv.x * rv.x + xv.y * rv.y + xv.z * rv.z + xv.w * rv.w;
    }
  } else {
    for (int j = lane; j < f; j += 32) {
      const float xv = __ldg(x + j);
      r = fmaf(xv, __ldg(a_l + (size_t)hid * f + j), r);
      c = fmaf(xv, __ldg(a_r + (size_t)hid * f + j), c);
    }
  }
  r = warp_sum(r);
  c = warp_sum(c);
  if (lane == 0) { attn_row[wid] = r; attn_col[wid] = c; }
}

// rn[i,h] = 1 / max(||H[i,h,:]||_2, 1e-12)   (F.normalize, agnn_layer_fused.py:14)
static __global__ void __launch_bounds__(256) inv_norm_kernel(int m, int h, int f,
                                                              const float* __restrict__ H,
                                                              float* __restrict__ rn) {
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= (long long)m * h) return;
  const float* x = H + (size_t)wid * f;
  float s = 0.f;
  if ((f & 3) == 0) {
    for (int j = lane; j < (f >> 2); j += 32) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x) + j);
      s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
  } else {
    for (int j = lane; j < f; j += 32) { const float v = __ldg(x + j); s = fmaf(v, v, s); }
  }
  s = warp_sum(s);
  if (lane == 0) rn[wid] = 1.f / fmaxf(sqrtf(s), 1e-12f);
}

}  // namespace dfgnn
