// fwd_kernels.cuh -- fused forward kernels: score -> online softmax -> aggregate.
//
//   dot_fwd_kernel<L, C, AGNN>   GT   : s_ij = <Q_i, K_j> * val_ij,  O_i = sum_j p_ij V_j
//                                AGNN : s_ij = <H_i, H_j> rn_i rn_j, O_i = sum_j p_ij H_j
//                                       (one gathered row serves score AND aggregation)
//   gat_fwd_kernel<L, C>         GAT  : s_ij = leakyrelu(ar_i + ac_j), O_i = sum_j keep_ij p_ij feat_j
//
// Maths spec: SURVEY.md section 8 (restating fused_gtconv_hyper.cu:31-163 and
// fused_gatconv_kernel.cu:24-125).  Schedule: rowblock.cuh.
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
};

// fold (m2, l2, acc2) into (m, l, acc) -- the usual online-softmax merge
template <int NR>
__device__ __forceinline__ void softmax_merge(float& m, float& l, float (&acc)[NR], float m2,
                                              float l2, const float (&acc2)[NR]) {
  const float mn = fmaxf(m, m2);
  const float sa = __expf(m - mn), sb = __expf(m2 - mn);
  l = l * sa + l2 * sb;
#pragma unroll
  for (int i = 0; i < NR; ++i) acc[i] = acc[i] * sa + acc2[i] * sb;
  m = mn;
}

// After the walk: fold split rows together and write them.  `fin(seg, m, l, acc)`
// is called by the owning warp (all lanes) once per split row.
template <int NR, class Fin>
__device__ __forceinline__ void softmax_merge_slots(float* s_slot, Fin fin) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  Slot<NR> mine(s_slot, w, 1);
  const int seg = mine.seg();
  if (seg < 0) return;
  float m = mine.a(), l = mine.b(), acc[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) acc[i] = mine.v(i, lane);
  for (int w2 = w + 1; w2 < kNW; ++w2) {
    Slot<NR> s(s_slot, w2, 0);
    if (s.seg() != seg) break;
    float acc2[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) acc2[i] = s.v(i, lane);
    softmax_merge<NR>(m, l, acc, s.a(), s.b(), acc2);
  }
  fin(seg, m, l, acc);
}

template <class L, int C, bool AGNN>
__global__ void __launch_bounds__(kNW * 32, 2) dot_fwd_kernel(const DotFwdParams p) {
  constexpr int NR = L::NR, LPR = L::LPR, G = L::G, EPS = G * C;
  static_assert(32 % EPS == 0, "edges per step must divide 32");
  __shared__ int s_rp[kMaxRB + 1];
  __shared__ float s_m[kMaxRB], s_inv[kMaxRB];
  extern __shared__ float s_slot[];

  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int grp = lane / LPR, gl = lane % LPR;
  const int hid = blockIdx.y, h = p.h, f = p.f;
  const bool train = p.attn != nullptr;
  float* attn = train ? p.attn + (size_t)hid * p.nnz : nullptr;

  slots_clear<NR>(s_slot);
  RowBlock b = rowblock_init(s_rp, p.row_ptr, p.m, p.rb);

  auto finish = [&](int r, float m, float l, float (&acc)[NR]) {
    const float inv = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
    for (int i = 0; i < NR; ++i) acc[i] *= inv;
    if (grp == 0) L::store(p.out + ((size_t)(b.seg_lb + r) * h + hid) * f, acc, gl, f);
    if (lane == 0) { s_m[r] = m; s_inv[r] = inv; }
  };

  // rows without edges produce zeros (fused_gtconv_hyper.cu:142-143)
  for (int r = w; r < b.nseg; r += kNW)
    if (s_rp[r + 1] == s_rp[r]) {
      float z[NR];
      zero(z);
      finish(r, kNeg, 0.f, z);
    }

  int e = b.e;
  if (e < b.e_end) {
    int r = find_row(s_rp, b.nseg, e);
    while (e < b.e_end) {
      while (s_rp[r + 1] <= e) ++r;
      const int rs = s_rp[r], re = s_rp[r + 1];
      const int seg_end = min(re, b.e_end);
      const bool starts = (e == rs), ends = (seg_end == re);
      const size_t node = (size_t)(b.seg_lb + r) * h + hid;

      float q[NR], acc[NR];
      L::load(q, p.Q + node * f, gl, f);
      zero(acc);
      const float rn_i = AGNN ? __ldg(p.rn + node) : 1.f;
      const bool use_w = AGNN || p.val != nullptr;
      float m_run = kNeg, l_run = 0.f;

      for (int base = e; base < seg_end; base += 32) {
        const int cnt = min(32, seg_end - base);
        int my_col = 0;
        float my_w = 1.f;
        if (lane < cnt) {
          my_col = __ldg(p.col_ind + base + lane);
          if (AGNN) my_w = __ldg(p.rn + (size_t)my_col * h + hid) * rn_i;
          else if (p.val) my_w = __ldg(p.val + base + lane);
        }
        for (int s = 0; s < cnt; s += EPS) {
          float kk[AGNN ? 1 : C][NR], vv[C][NR], d[C];
          bool ok[C];
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const int idx = s + c * G + grp;
            ok[c] = idx < cnt;
            const int col = __shfl_sync(kFull, my_col, idx);
            const size_t off = ((size_t)col * h + hid) * f;
            if (ok[c]) {
              if (!AGNN) L::load(kk[c], p.K + off, gl, f);
              L::load(vv[c], p.V + off, gl, f);
            } else {
              if (!AGNN) zero(kk[c]);
              zero(vv[c]);
            }
          }
          float cm = kNeg;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const int idx = s + c * G + grp;
            float dc = AGNN ? dot<NR>(q, vv[c]) : dot<NR>(q, kk[AGNN ? 0 : c]);
            dc = group_sum<LPR>(dc);
            if (use_w) dc *= __shfl_sync(kFull, my_w, idx);
            if (train && gl == 0 && ok[c]) attn[base + idx] = dc;
            d[c] = dc;
            cm = ok[c] ? fmaxf(cm, dc) : cm;
          }
          const float m_new = fmaxf(m_run, cm);
          const float scale = __expf(m_run - m_new);
#pragma unroll
          for (int i = 0; i < NR; ++i) acc[i] *= scale;
          float ps = 0.f;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float pc = ok[c] ? __expf(d[c] - m_new) : 0.f;
            ps += pc;
#pragma unroll
            for (int i = 0; i < NR; ++i) acc[i] = fmaf(pc, vv[c][i], acc[i]);
          }
          l_run = l_run * scale + ps;
          m_run = m_new;
        }
      }
      // lane groups saw disjoint edges: fold them
#pragma unroll
      for (int off = LPR; off < 32; off <<= 1) {
        const float m2 = __shfl_xor_sync(kFull, m_run, off);
        const float l2 = __shfl_xor_sync(kFull, l_run, off);
        float acc2[NR];
#pragma unroll
        for (int i = 0; i < NR; ++i) acc2[i] = __shfl_xor_sync(kFull, acc[i], off);
        softmax_merge<NR>(m_run, l_run, acc, m2, l2, acc2);
      }
      if (starts && ends) {
        finish(r, m_run, l_run, acc);
      } else {
        Slot<NR> sl(s_slot, w, starts ? 1 : 0);
#pragma unroll
        for (int i = 0; i < NR; ++i) sl.v(i, lane) = acc[i];
        if (lane == 0) { sl.a() = m_run; sl.b() = l_run; sl.set_seg(r); }
      }
      e = seg_end;
    }
  }
  __syncthreads();
  softmax_merge_slots<NR>(s_slot, finish);

  if (train) {  // scores -> probabilities (attn_edge of fused_gtconv_hyper.cu:146-149)
    __syncthreads();
    for (int i = b.E0 + threadIdx.x; i < b.E1; i += kNW * 32) {
      const int r = find_row(s_rp, b.nseg, i);
      attn[i] = __expf(attn[i] - s_m[r]) * s_inv[r];
    }
  }
}

// ------------------------------------------------------------------------- //

struct GatFwdParams {
  int m, nnz, h, f, rb;
  const int* row_ptr;
  const int* col_ind;
  const float* ar;    // attn_row [m, h]
  const float* ac;    // attn_col [m, h]
  const float* feat;
  float slope, drop;
  uint64_t seed;
  float* out;
  float* emax;   // [m, h] or null
  float* esum;   // [m, h] or null
  float* emask;  // [nnz, h] or null
};

template <class L, int C>
__global__ void __launch_bounds__(kNW * 32, 3) gat_fwd_kernel(const GatFwdParams p) {
  constexpr int NR = L::NR, LPR = L::LPR, G = L::G, EPS = G * C;
  static_assert(32 % EPS == 0, "edges per step must divide 32");
  __shared__ int s_rp[kMaxRB + 1];
  extern __shared__ float s_slot[];

  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int grp = lane / LPR, gl = lane % LPR;
  const int hid = blockIdx.y, h = p.h, f = p.f;
  const bool use_mask = p.emask != nullptr;
  const float keep_scale = 1.f / (1.f - p.drop);

  slots_clear<NR>(s_slot);
  RowBlock b = rowblock_init(s_rp, p.row_ptr, p.m, p.rb);

  auto finish = [&](int r, float m, float l, float (&acc)[NR]) {
    const size_t node = (size_t)(b.seg_lb + r) * h + hid;
    const float inv = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
    for (int i = 0; i < NR; ++i) acc[i] *= inv;
    if (grp == 0) L::store(p.out + node * f, acc, gl, f);
    if (lane == 0 && p.emax) {  // saved for backward (fused_gatconv_kernel.cu:66-68, 89-91)
      p.emax[node] = l > 0.f ? m : -1e38f;
      p.esum[node] = l;
    }
  };

  for (int r = w; r < b.nseg; r += kNW)
    if (s_rp[r + 1] == s_rp[r]) {
      float z[NR];
      zero(z);
      finish(r, kNeg, 0.f, z);
    }

  int e = b.e;
  if (e < b.e_end) {
    int r = find_row(s_rp, b.nseg, e);
    while (e < b.e_end) {
      while (s_rp[r + 1] <= e) ++r;
      const int rs = s_rp[r], re = s_rp[r + 1];
      const int seg_end = min(re, b.e_end);
      const bool starts = (e == rs), ends = (seg_end == re);
      const float ar_i = __ldg(p.ar + (size_t)(b.seg_lb + r) * h + hid);

      float acc[NR];
      zero(acc);
      float m_run = kNeg, l_lane = 0.f;

      for (int base = e; base < seg_end; base += 32) {
        const int cnt = min(32, seg_end - base);
        int my_col = 0;
        float sc = kNeg;
        if (lane < cnt) {
          my_col = __ldg(p.col_ind + base + lane);
          sc = leaky(ar_i + __ldg(p.ac + (size_t)my_col * h + hid), p.slope);
        }
        const float m_new = fmaxf(m_run, warp_max(sc));
        const float scale = __expf(m_run - m_new);
        float pe = lane < cnt ? __expf(sc - m_new) : 0.f;
        l_lane = l_lane * scale + pe;
        m_run = m_new;
        if (use_mask && lane < cnt) {  // dropout on the attention weights
          const size_t eid = (size_t)(base + lane) * h + hid;
          const float u = uniform01(p.seed, eid);
          p.emask[eid] = u;
          pe = (u > p.drop) ? pe * keep_scale : 0.f;
        }
#pragma unroll
        for (int i = 0; i < NR; ++i) acc[i] *= scale;
        for (int s = 0; s < cnt; s += EPS) {
          float vv[C][NR], pc[C];
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const int idx = s + c * G + grp;
            const int col = __shfl_sync(kFull, my_col, idx);
            pc[c] = __shfl_sync(kFull, pe, idx);
            if (idx < cnt) L::load(vv[c], p.feat + ((size_t)col * h + hid) * f, gl, f);
            else zero(vv[c]);
          }
#pragma unroll
          for (int c = 0; c < C; ++c)
#pragma unroll
            for (int i = 0; i < NR; ++i) acc[i] = fmaf(pc[c], vv[c][i], acc[i]);
        }
      }
      float l_run = warp_sum(l_lane);
#pragma unroll
      for (int off = LPR; off < 32; off <<= 1)
#pragma unroll
        for (int i = 0; i < NR; ++i) acc[i] += __shfl_xor_sync(kFull, acc[i], off);

      if (starts && ends) {
        finish(r, m_run, l_run, acc);
      } else {
        Slot<NR> sl(s_slot, w, starts ? 1 : 0);
#pragma unroll
        for (int i = 0; i < NR; ++i) sl.v(i, lane) = acc[i];
        if (lane == 0) { sl.a() = m_run; sl.b() = l_run; sl.set_seg(r); }
      }
      e = seg_end;
    }
  }
  __syncthreads();
  softmax_merge_slots<NR>(s_slot, finish);
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
