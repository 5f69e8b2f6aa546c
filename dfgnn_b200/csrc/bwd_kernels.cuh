// bwd_kernels.cuh -- backward kernels (row side over the CSR, column side over the CSC).
//
// GT / AGNN (restating fused_gtconv_backward.cu:40-191):
//   row side   dA_e = <dO_i, V_j>; t_e = dA_e p_e; s_i = sum_e t_e;
//              dS_e = t_e - s_i p_e -> grad_edge;  dQ_i = sum_e dS_e K_j
//              (computed in ONE pass over the row, see gt_bwd_row_kernel)
//   col side   dV_j = sum_i p_ij dO_i;  dK_j = sum_i dS_ij Q_i   (deterministic, no atomics)
// GAT (restating fused_gatconv_kernel.cu:609-660, 711-865):
//   row side   g_e = keep_e/(1-drop) <dO_i, feat_j>; t_e = p_e g_e; w_i = sum_e t_e;
//              de_e = (t_e - w_i p_e) * (e_ij < 0 ? slope : 1) -> grad_edge;
//              grad_attn_row_i = sum_e de_e
//   col side   grad_feat_j = sum_i keep/(1-drop) p_ij dO_i;  grad_attn_col_j = sum_i de_ij
//              (the reference scatters grad_attn_col with atomicAdd, l.854; here it is
//               a column-side sum through `permute`, bit-reproducible)
// Schedule: rowblock.cuh (walk_pieces).
#pragma once

#include "rowblock.cuh"

namespace dfgnn {

struct GtBwdParams {
  int m, n, nnz, h, f, rb, rb_col;  // m rows, n columns
  const int* row_ptr;
  const int* col_ind;
  const int* col_ptr;
  const int* row_ind;
  const int* val_idx;
  const float* Q;
  const float* K;
  const float* V;
  const float* attn;   // [h, nnz] probabilities
  const float* val;    // [nnz] edge weights of the forward scores, or null (all ones)
  const float* dO;
  float* dQ;
  float* dK;
  float* dV;
  float* grad_edge;    // [h, nnz, 2] scratch: {dS_e, p_e}
  int cap = 0;         // > 0: process only tiles with more than `cap` entries
};

// Sum-merge of split segments: NV floats per lane + one scalar (slot.a).  Group-local.
template <int NV, int LPR, int G, class Fin>
__device__ __forceinline__ void sum_merge_slots(float* s_slot, int vw, int gl, Fin fin) {
  Slot<NV, LPR> mine(s_slot, vw, 1);
  const int seg = mine.seg();
  if (seg < 0) return;
  float a = mine.a(), acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = mine.v(i, gl);
  for (int v2 = vw + 1; v2 < kNW * G; ++v2) {
    Slot<NV, LPR> s(s_slot, v2, 0);
    if (s.seg() != seg) break;
    a += s.a();
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] += s.v(i, gl);
  }
  fin(seg, a, acc);
}

template <class L, int C>
__global__ void __launch_bounds__(kNW * 32, (L::NR <= 8 ? DFGNN_GT_WARPS_SMALL : DFGNN_GT_WARPS) / kNW) gt_bwd_row_kernel(const GtBwdParams p) {
  constexpr int NR = L::NR, LPR = L::LPR, G = L::G, VW = kNW * G;
  constexpr int CH = ChunkOf<L>::kChunk;
  static_assert(CH % C == 0 && CH <= LPR, "chunking");
  __shared__ int s_rp[kMaxRB + 1];
  __shared__ float s_s[kMaxRB];
  extern __shared__ float s_slot[];

  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int grp = lane / LPR, gl = lane % LPR, vw = w * G + grp;
  const int hid = blockIdx.y, h = p.h, f = p.f;
  const float* attn = p.attn + (size_t)hid * p.nnz;
  float2* gedge = reinterpret_cast<float2*>(p.grad_edge) + (size_t)hid * p.nnz;  // {dS_e, p_e}
  const RowAddr<L> ra(h, f, hid, gl);
  const char* Gb = ra.base(p.dO);
  const char* Kb = ra.base(p.K);
  const char* Vb = ra.base(p.V);
  char* DQb = ra.base(p.dQ);

  slots_clear<2 * NR, LPR>(s_slot, vw, gl);
  const RowBlock b = rowblock_init<G>(s_rp, p.row_ptr, p.m, p.rb, vw);
  if (p.cap > 0 && b.E1 - b.E0 <= p.cap) {  // not a big tile: nothing to do here
    dependency_wait();
    return;
  }

  // Per piece: acc2 = [A1c | A2] with A1c = sum_e p_e (dA_e - c) K_e, A2 = sum_e p_e K_e,
  // c = dA of the piece's first edge.  dQ = A1c + (c - s) * A2, which equals
  // sum_e (t_e - s p_e) K_e but keeps the subtraction at the scale of the dS terms
  // (a row with one edge gives exactly 0, like the reference's two-pass form).
  auto write_dq = [&](int r, float s, const float (&dq)[NR]) {
    L::store(ra.at(DQb, b.seg_lb + r), dq, gl, f);
    if (gl == 0) s_s[r] = s;
  };

  for (int r = vw; r < b.nseg; r += VW)
    if (s_rp[r + 1] == s_rp[r]) {
      float z[NR];
      zero(z);
      write_dq(r, 0.f, z);
    }

  float g[NR], acc2[2 * NR];
  zero(g);
  zero(acc2);
  float s_part = 0.f, c_ref = 0.f;
  bool have_c = false;
  walk_pieces<CH>(
      b, s_rp,
      [&](int r) {
        L::load(g, ra.at(Gb, b.seg_lb + r), gl, f);
        zero(acc2);
        s_part = 0.f;
        c_ref = 0.f;
        have_c = false;
      },
      [&](int base, int cnt) {
        int my_col = 0;
        float my_p = 0.f, my_t = 0.f, my_w = 1.f;
        if (gl < cnt) {
          my_col = __ldg(p.col_ind + base + gl);
          my_p = __ldg(attn + base + gl);
          if (p.val) my_w = __ldg(p.val + base + gl);
        }
#pragma unroll
        for (int s = 0; s < CH; s += C) {
          if (s > 0 && !__any_sync(kFull, s < cnt)) break;
          float kk[C][NR], vv[C][NR], dA[C];
#pragma unroll
          for (int c = 0; c < C; ++c) {
            // beyond cnt: the chunk's first neighbour again (valid, cached; probability 0)
            const int col = group_bcast<LPR>(my_col, s + c < cnt ? s + c : 0);
            L::load(vv[c], ra.at(Vb, col), gl, f);
            L::load(kk[c], ra.at(Kb, col), gl, f);
          }
#pragma unroll
          for (int c = 0; c < C; ++c) dA[c] = group_sum<LPR>(dot<NR>(g, vv[c]));
          if (!have_c && cnt > 0) {  // the piece's first edge is slot 0 of its first step
            c_ref = dA[0];
            have_c = true;
          }
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float pc = group_bcast<LPR>(my_p, s + c);  // 0 beyond cnt
            const float t = dA[c] * pc;
            float u = (dA[c] - c_ref) * pc, pw = pc;
            if (p.val) {  // s_e = <Q_i, K_j> * val_e: the K-side factors carry val_e
              const float wv = group_bcast<LPR>(my_w, s + c);
              u *= wv;
              pw *= wv;
            }
            if (gl == s + c) my_t = t;
            s_part += t;
#pragma unroll
            for (int i = 0; i < NR; ++i) {
              acc2[i] = fmaf(u, kk[c][i], acc2[i]);
              acc2[NR + i] = fmaf(pw, kk[c][i], acc2[NR + i]);
            }
          }
        }
        if (gl < cnt) gedge[base + gl] = make_float2(my_t, my_p);  // {t_e, p_e}
      },
      [&](int r, bool first, bool last) {
        if (first && last) {
          float dq[NR];
#pragma unroll
          for (int i = 0; i < NR; ++i) dq[i] = fmaf(c_ref - s_part, acc2[NR + i], acc2[i]);
          write_dq(r, s_part, dq);
        } else {
          Slot<2 * NR, LPR> sl(s_slot, vw, first ? 1 : 0);
#pragma unroll
          for (int i = 0; i < 2 * NR; ++i) sl.v(i, gl) = acc2[i];
          if (gl == 0) { sl.a() = s_part; sl.b() = c_ref; sl.set_seg(r); }
        }
      });
  __syncthreads();
  {  // rows split over groups: s first, then the re-centred vectors
    Slot<2 * NR, LPR> mine(s_slot, vw, 1);
    const int seg = mine.seg();
    if (seg >= 0) {
      float s = mine.a();
      int v_end = vw + 1;
      for (; v_end < VW; ++v_end) {
        Slot<2 * NR, LPR> sl(s_slot, v_end, 0);
        if (sl.seg() != seg) break;
        s += sl.a();
      }
      float dq[NR];
#pragma unroll
      for (int i = 0; i < NR; ++i) dq[i] = fmaf(mine.b() - s, mine.v(NR + i, gl), mine.v(i, gl));
      for (int v2 = vw + 1; v2 < v_end; ++v2) {
        Slot<2 * NR, LPR> sl(s_slot, v2, 0);
        const float dc = sl.b() - s;
#pragma unroll
        for (int i = 0; i < NR; ++i) dq[i] += fmaf(dc, sl.v(NR + i, gl), sl.v(i, gl));
      }
      write_dq(seg, s, dq);
    }
  }
  __syncthreads();
  // t_e -> dS_e = t_e - s_i p_e   (fused_gtconv_backward.cu:171-176); times val_e when the
  // forward scores were weighted, so that the column side sees d(score)/d<Q,K> (the reference
  // drops val in its backward, l.126 -- identical for its all-ones val)
  for (int i = b.E0 + threadIdx.x; i < b.E1; i += kNW * 32) {
    const int rr = find_row(s_rp, b.nseg, i);
    const float2 tp = gedge[i];
    float ds = fmaf(-s_s[rr], tp.y, tp.x);
    if (p.val) ds *= __ldg(p.val + i);
    gedge[i].x = ds;
  }
  if (p.cap > 0) dependency_wait();  // see common.cuh
}

// Column side: segments are CSC columns.  dV_j = sum p dO_i, dK_j = sum dS Q_i.
template <class L, int C>
__global__ void __launch_bounds__(kNW * 32, (L::NR <= 8 ? DFGNN_GT_WARPS_SMALL : DFGNN_GT_WARPS) / kNW) gt_bwd_col_kernel(const GtBwdParams p) {
  constexpr int NR = L::NR, LPR = L::LPR, G = L::G, VW = kNW * G;
  constexpr int CH = ChunkOf<L>::kChunk;
  static_assert(CH % C == 0 && CH <= LPR, "chunking");
  __shared__ int s_cp[kMaxRB + 1];
  extern __shared__ float s_slot[];

  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int grp = lane / LPR, gl = lane % LPR, vw = w * G + grp;
  const int hid = blockIdx.y, h = p.h, f = p.f;
  const float2* gedge = reinterpret_cast<const float2*>(p.grad_edge) + (size_t)hid * p.nnz;
  const RowAddr<L> ra(h, f, hid, gl);
  const char* Gb = ra.base(p.dO);
  const char* Qb = ra.base(p.Q);
  char* DVb = ra.base(p.dV);
  char* DKb = ra.base(p.dK);

  slots_clear<2 * NR, LPR>(s_slot, vw, gl);
  const RowBlock b = rowblock_init<G>(s_cp, p.col_ptr, p.n, p.rb_col, vw);
  if (p.cap > 0 && b.E1 - b.E0 <= p.cap) {  // not a big tile: nothing to do here
    dependency_wait();
    return;
  }

  // acc2 = [dV | dK]
  auto finish = [&](int c, float, float (&acc2)[2 * NR]) {
    float t[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) t[i] = acc2[i];
    L::store(ra.at(DVb, b.seg_lb + c), t, gl, f);
#pragma unroll
    for (int i = 0; i < NR; ++i) t[i] = acc2[NR + i];
    L::store(ra.at(DKb, b.seg_lb + c), t, gl, f);
  };

  for (int c = vw; c < b.nseg; c += VW)
    if (s_cp[c + 1] == s_cp[c]) {
      float z[2 * NR];
      zero(z);
      finish(c, 0.f, z);
    }

  float acc2[2 * NR];
  zero(acc2);
  walk_pieces<CH>(
      b, s_cp, [&](int) { zero(acc2); },
      [&](int base, int cnt) {
        int my_rid = 0;
        float my_p = 0.f, my_ds = 0.f;
        if (gl < cnt) {
          my_rid = __ldg(p.row_ind + base + gl);
          const float2 dp = __ldg(gedge + __ldg(p.val_idx + base + gl));  // {dS_e, p_e} from the row side
          my_p = dp.y;
          my_ds = dp.x;
        }
#pragma unroll
        for (int s = 0; s < CH; s += C) {
          if (s > 0 && !__any_sync(kFull, s < cnt)) break;
          float go[C][NR], qq[C][NR];
#pragma unroll
          for (int c = 0; c < C; ++c) {
            // beyond cnt: the chunk's first entry again (valid, cached; weights 0)
            const int rid = group_bcast<LPR>(my_rid, s + c < cnt ? s + c : 0);
            L::load(go[c], ra.at(Gb, rid), gl, f);
            L::load(qq[c], ra.at(Qb, rid), gl, f);
          }
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float pc = group_bcast<LPR>(my_p, s + c);
            const float ds = group_bcast<LPR>(my_ds, s + c);
#pragma unroll
            for (int i = 0; i < NR; ++i) {
              acc2[i] = fmaf(pc, go[c][i], acc2[i]);
              acc2[NR + i] = fmaf(ds, qq[c][i], acc2[NR + i]);
            }
          }
        }
      },
      [&](int c0, bool first, bool last) {
        if (first && last) {
          finish(c0, 0.f, acc2);
        } else {
          Slot<2 * NR, LPR> sl(s_slot, vw, first ? 1 : 0);
#pragma unroll
          for (int i = 0; i < 2 * NR; ++i) sl.v(i, gl) = acc2[i];
          if (gl == 0) { sl.a() = 0.f; sl.set_seg(c0); }
        }
      });
  __syncthreads();
  sum_merge_slots<2 * NR, LPR, G>(s_slot, vw, gl, finish);
  if (p.cap > 0) dependency_wait();  // see common.cuh
}

// ------------------------------------------------------------------------- //

struct GatBwdParams {
  int m, n, nnz, h, f, rb, rb_col;  // m rows, n columns
  float slope, drop;
  const int* row_ptr;
  const int* col_ind;
  const int* col_ptr;
  const int* row_ind;
  const int* permute;
  const float* emax;
  const float* esum;
  const float* emask;   // [nnz, h] or null (keep all)
  const float* feat;
  const float* ar;
  const float* ac;
  const float* dO;
  float* grad_feat;
  float* grad_ar;
  float* grad_ac;
  float* grad_edge;     // [nnz, h, 2] scratch: {t_e then de_e, keep-scaled p_e}
  int cap = 0;          // > 0: process only tiles with more than `cap` entries
};

template <class L, int C>
__global__ void __launch_bounds__(kNW * 32, (L::NR <= 8 ? 24 : 16) / kNW) gat_bwd_row_kernel(const GatBwdParams p) {
  constexpr int NR = L::NR, LPR = L::LPR, G = L::G, VW = kNW * G;
  static_assert(LPR % C == 0, "chunking");
  __shared__ int s_rp[kMaxRB + 1];
  __shared__ float s_w[kMaxRB];
  extern __shared__ float s_slot[];

  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int grp = lane / LPR, gl = lane % LPR, vw = w * G + grp;
  const int hid = blockIdx.y, h = p.h, f = p.f;
  const float keep_scale = 1.f / (1.f - p.drop);
  const RowAddr<L> ra(h, f, hid, gl);
  const char* Gb = ra.base(p.dO);
  const char* Fb = ra.base(p.feat);
  const float* acb = p.ac + hid;
  float2* scratch = reinterpret_cast<float2*>(p.grad_edge);  // [nnz, h] x {de_e, keep-scaled p_e}

  if (small_tile_exit(p.row_ptr, p.m, p.rb, p.cap)) return;
  slots_clear<1, LPR>(s_slot, vw, gl);
  const RowBlock b = rowblock_init<G>(s_rp, p.row_ptr, p.m, p.rb, vw);
  if (p.cap > 0 && b.E1 - b.E0 <= p.cap) {  // not a big tile: nothing to do here
    dependency_wait();
    return;
  }

  for (int r = vw; r < b.nseg; r += VW)
    if (s_rp[r + 1] == s_rp[r] && gl == 0) s_w[r] = 0.f;

  float g[NR], ar_i = 0.f, mx = 0.f, inv = 0.f, w_lane = 0.f;
  zero(g);
  walk_pieces<LPR>(
      b, s_rp,
      [&](int r) {
        const size_t node = (size_t)(b.seg_lb + r) * h + hid;
        L::load(g, ra.at(Gb, b.seg_lb + r), gl, f);
        ar_i = __ldg(p.ar + node);
        mx = __ldg(p.emax + node);
        inv = 1.f / __ldg(p.esum + node);
        w_lane = 0.f;
      },
      [&](int base, int cnt) {
        int my_col = 0;
        float my_p = 0.f, my_t = 0.f;  // my_p = p_e * keep_e / (1 - drop)
        if (gl < cnt) {
          my_col = __ldg(p.col_ind + base + gl);
          const float sc = leaky(ar_i + __ldg(acb + (size_t)(unsigned)my_col * (unsigned)h), p.slope);
          my_p = fast_exp(sc - mx) * inv;
          if (p.emask)
            my_p = (__ldg(p.emask + (size_t)(base + gl) * h + hid) > p.drop) ? my_p * keep_scale : 0.f;
        }
#pragma unroll
        for (int s = 0; s < LPR; s += C) {
          if (s > 0 && !__any_sync(kFull, s < cnt)) break;
          float ff[C][NR];
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const int col = group_bcast<LPR>(my_col, s + c < cnt ? s + c : 0);  // beyond cnt: first neighbour, unused
            L::load(ff[c], ra.at(Fb, col), gl, f);
          }
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float ge = group_sum<LPR>(dot<NR>(g, ff[c]));
            if (gl == s + c) my_t = ge * my_p;  // lane s+c owns this edge
          }
        }
        if (gl < cnt) scratch[(size_t)(base + gl) * h + hid] = make_float2(my_t, my_p);  // {t_e, keep-scaled p_e}
        w_lane += my_t;
      },
      [&](int r, bool first, bool last) {
        const float w_part = group_sum_local<LPR>(w_lane, lane);
        if (first && last) {
          if (gl == 0) s_w[r] = w_part;
        } else {
          Slot<1, LPR> sl(s_slot, vw, first ? 1 : 0);
          if (gl == 0) { sl.a() = w_part; sl.set_seg(r); }
        }
      });
  __syncthreads();
  {
    auto fin = [&](int rr, float a, float (&)[1]) { if (gl == 0) s_w[rr] = a; };
    sum_merge_slots<1, LPR, G>(s_slot, vw, gl, fin);
  }
  __syncthreads();
  // de_e and grad_attn_row, one row per group, groups of a warp in lockstep
  // (fused_gatconv_kernel.cu:830-864)
  for (int r0 = w * G; r0 < b.nseg; r0 += VW) {
    const int rr = r0 + grp;
    const bool valid = rr < b.nseg;
    const int rs = valid ? s_rp[rr] : 0, re = valid ? s_rp[rr + 1] : 0;
    const size_t node = (size_t)(b.seg_lb + (valid ? rr : 0)) * h + hid;
    const float ar_i = __ldg(p.ar + node);
    const float mx = __ldg(p.emax + node);
    const float inv = re > rs ? 1.f / __ldg(p.esum + node) : 0.f;
    const float wr = valid ? s_w[rr] : 0.f;
    float rsum = 0.f;
    for (int i = rs + gl; __any_sync(kFull, i < re); i += LPR) {
      if (i < re) {
        const int col = __ldg(p.col_ind + i);
        const float x = leaky(ar_i + __ldg(acb + (size_t)(unsigned)col * (unsigned)h), p.slope);
        const float pe = fast_exp(x - mx) * inv;
        const size_t eid = (size_t)i * h + hid;
        float de = fmaf(-wr, pe, scratch[eid].x);
        if (x < 0.f) de *= p.slope;
        scratch[eid].x = de;
        rsum += de;
      }
    }
    rsum = group_sum<LPR>(rsum);
    if (valid && gl == 0) p.grad_ar[node] = rsum;
  }
  if (p.cap > 0) dependency_wait();  // see common.cuh
}

template <class L, int C>
__global__ void __launch_bounds__(kNW * 32, (L::NR * C <= 32 ? 24 : 16) / kNW) gat_bwd_col_kernel(const GatBwdParams p) {
  constexpr int NR = L::NR, LPR = L::LPR, G = L::G, VW = kNW * G;
  static_assert(LPR % C == 0, "chunking");
  __shared__ int s_cp[kMaxRB + 1];
  extern __shared__ float s_slot[];

  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int grp = lane / LPR, gl = lane % LPR, vw = w * G + grp;
  const int hid = blockIdx.y, h = p.h, f = p.f;
  const RowAddr<L> ra(h, f, hid, gl);
  const char* Gb = ra.base(p.dO);
  char* GFb = ra.base(p.grad_feat);
  const float2* scratch = reinterpret_cast<const float2*>(p.grad_edge);

  if (small_tile_exit(p.col_ptr, p.n, p.rb_col, p.cap)) return;
  slots_clear<NR, LPR>(s_slot, vw, gl);
  const RowBlock b = rowblock_init<G>(s_cp, p.col_ptr, p.n, p.rb_col, vw);
  if (p.cap > 0 && b.E1 - b.E0 <= p.cap) {  // not a big tile: nothing to do here
    dependency_wait();
    return;
  }

  auto finish = [&](int c, float dac, float (&acc)[NR]) {
    const size_t node = (size_t)(b.seg_lb + c) * h + hid;
    L::store(ra.at(GFb, b.seg_lb + c), acc, gl, f);
    if (gl == 0) p.grad_ac[node] = dac;
  };

  for (int c = vw; c < b.nseg; c += VW)
    if (s_cp[c + 1] == s_cp[c]) {
      float z[NR];
      zero(z);
      finish(c, 0.f, z);
    }

  float acc[NR], dac_lane = 0.f;
  zero(acc);
  walk_pieces<LPR>(
      b, s_cp,
      [&](int) {
        zero(acc);
        dac_lane = 0.f;
      },
      [&](int base, int cnt) {
        int my_rid = 0;
        float my_p = 0.f;
        if (gl < cnt) {
          my_rid = __ldg(p.row_ind + base + gl);
          const size_t eid = (size_t)__ldg(p.permute + base + gl) * h + hid;
          const float2 dp = __ldg(scratch + eid);  // {de_e, keep-scaled p_e} from the row side
          my_p = dp.y;
          dac_lane += dp.x;
        }
#pragma unroll
        for (int s = 0; s < LPR; s += C) {
          if (s > 0 && !__any_sync(kFull, s < cnt)) break;
          float go[C][NR], pc[C];
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const int rid = group_bcast<LPR>(my_rid, s + c < cnt ? s + c : 0);  // beyond cnt: first entry, weight 0
            pc[c] = group_bcast<LPR>(my_p, s + c);
            L::load(go[c], ra.at(Gb, rid), gl, f);
          }
#pragma unroll
          for (int c = 0; c < C; ++c)
#pragma unroll
            for (int i = 0; i < NR; ++i) acc[i] = fmaf(pc[c], go[c][i], acc[i]);
        }
      },
      [&](int c0, bool first, bool last) {
        const float dac = group_sum_local<LPR>(dac_lane, lane);
        if (first && last) {
          finish(c0, dac, acc);
        } else {
          Slot<NR, LPR> sl(s_slot, vw, first ? 1 : 0);
#pragma unroll
          for (int i = 0; i < NR; ++i) sl.v(i, gl) = acc[i];
          if (gl == 0) { sl.a() = dac; sl.set_seg(c0); }
        }
      });
  __syncthreads();
  sum_merge_slots<NR, LPR, G>(s_slot, vw, gl, finish);
  if (p.cap > 0) dependency_wait();  // see common.cuh
}

}  // namespace dfgnn
