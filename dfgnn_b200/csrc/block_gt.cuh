// block_gt.cuh -- "graph-resident" GT kernels for block-diagonal batches of small graphs.
//
// A batched graph (DGL batch, PATTERN-shaped: ~119 nodes, ~51 neighbours per row, 43 % dense
// blocks) is block diagonal: every neighbour of a row lies in the row's own graph.  One CTA owns
// ONE graph: the two operand blocks its edges gather from (K and V forward / row side, dO and Q
// column side; [nodes, f] fp32, contiguous for h == 1) are copied into shared memory by the TMA
// unit (1-D cp.async.bulk completing on an mbarrier, issued by one thread while the other threads
// load the segment pointers) and every neighbour-row gather of the walk is then an LDS.128 instead
// of an L1/L2 access: each staged row is reused deg(row) ~ 51 times.  The walk itself is the
// segmented schedule of rowblock.cuh (equal entry slices per lane group, split rows merged through
// shared-memory slots), so the maths and its order are those of dot_fwd_kernel / gt_bwd_row_kernel
// / gt_bwd_col_kernel.
//
// The reference's counterpart is the hyper kernel (fused_gtconv_hyper.cu:31-163), which stages the
// SCORES of 8 rows in shared memory and still gathers K and V rows from global memory per edge.
//
// Bound: shared-memory bandwidth, 128 B/clk/SM -- 4 LDS.128 (K/V or dO/Q quarter rows) per edge and
// lane, i.e. 1 KB per edge: 6.2 M edges / 148 SMs * 8 clk = 0.18 ms per kernel on the PATTERN-shaped
// batch at 1.9 GHz (the L1/L2 gather path of the row-block kernels needs 0.39-0.49 ms).
#pragma once

#include "bwd_kernels.cuh"
#include "fwd_kernels.cuh"

namespace dfgnn {

struct BlockPlan {
  const int* blk_ptr;  // [n_blocks + 1] first node of every graph
  int n_blocks;
  int max_nodes;       // largest graph (sizes the shared-memory stage)
};

struct GtBlockFwdParams {
  DotFwdParams c;
  BlockPlan b;
};
struct GtBlockBwdParams {
  GtBwdParams c;
  BlockPlan b;
};

// segment pointers of ONE block -> s_ptr, and the lane group's equal slice of its entries.
// Contains a __syncthreads().
template <int G, int NW>
__device__ __forceinline__ RowBlock segblock_init(int* s_ptr, const int* __restrict__ seg_ptr, int seg_lb,
                                                  int nseg, int vw) {
  RowBlock b;
  b.seg_lb = seg_lb;
  b.nseg = nseg;
  for (int i = threadIdx.x; i <= nseg; i += NW * 32) s_ptr[i] = __ldg(seg_ptr + seg_lb + i);
  __syncthreads();
  b.E0 = s_ptr[0];
  b.E1 = s_ptr[nseg];
  constexpr int VW = NW * G;
  const int per = (b.E1 - b.E0 + VW - 1) / VW;
  b.e = min(b.E1, b.E0 + vw * per);
  b.e_end = min(b.E1, b.e + per);
  return b;
}

// shared-memory carve of the block kernels: [A: mn*f][B: mn*f][slots][s_ptr: mn+1 ints][2 x mn floats]
template <int NV, class L, int NW>
__host__ __device__ constexpr size_t block_slot_floats() {
  return (size_t)NW * L::G * 2 * Slot<NV, L::LPR>::kFloats;
}
template <int NV, class L, int NW>
inline size_t block_smem_bytes(int max_nodes, int f) {
  return ((size_t)2 * max_nodes * f + block_slot_floats<NV, L, NW>() + (size_t)3 * max_nodes + 4) * sizeof(float);
}

// one thread: both operand blocks of the graph -> shared memory (TMA), completion on `bar`
__device__ __forceinline__ void stage_two(float* sA, float* sB, const float* A, const float* B, int seg_lb,
                                          int nseg, int f, uint64_t* bar) {
  const uint32_t bytes = (uint32_t)nseg * (uint32_t)f * 4u;
  mbar_expect_tx(bar, 2 * bytes);
  bulk_g2s_range(sA, A + (size_t)seg_lb * f, bytes, bar);
  bulk_g2s_range(sB, B + (size_t)seg_lb * f, bytes, bar);
}

template <class L, int C, int NW>
__global__ void __launch_bounds__(NW * 32, 1) gt_block_fwd_kernel(const GtBlockFwdParams pp) {
  const DotFwdParams& p = pp.c;
  constexpr int NR = L::NR, LPR = L::LPR, G = L::G, VW = NW * G;
  constexpr int CH = ChunkOf<L>::kChunk;
  static_assert(L::kVec && CH % C == 0 && CH <= LPR, "vector layouts; chunking");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint64_t s_bar;
  const int f = p.f, mn = pp.b.max_nodes;
  float* sK = reinterpret_cast<float*>(smem_raw);
  float* sV = sK + (size_t)mn * f;
  float* s_slot = sV + (size_t)mn * f;
  int* s_rp = reinterpret_cast<int*>(s_slot + block_slot_floats<NR, L, NW>());
  float* s_m = reinterpret_cast<float*>(s_rp + mn + 1);
  float* s_inv = s_m + mn;

  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int grp = lane / LPR, gl = lane % LPR, vw = w * G + grp;
  const int seg_lb = __ldg(pp.b.blk_ptr + blockIdx.x);
  const int nseg = __ldg(pp.b.blk_ptr + blockIdx.x + 1) - seg_lb;
  if (nseg <= 0) return;
  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
    stage_two(sK, sV, p.K, p.V, seg_lb, nseg, f, &s_bar);
  }
  const bool train = p.attn != nullptr;
  const bool use_w = p.val != nullptr;
  float* attn = p.attn;
  const RowAddr<L> ra(1, f, 0, gl);
  const char* Qb = ra.base(p.Q);
  char* Ob = ra.base(p.out);
  const float* kl = sK + 4 * gl;  // this lane's float4 column of every staged row
  const float* vl = sV + 4 * gl;

  slots_clear<NR, LPR>(s_slot, vw, gl);
  const RowBlock b = segblock_init<G, NW>(s_rp, p.row_ptr, seg_lb, nseg, vw);  // barrier: s_bar is initialised for all
  mbar_wait(&s_bar, 0);

  auto finish = [&](int r, float m, float l, float (&acc)[NR]) {
    const float inv = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
    for (int i = 0; i < NR; ++i) acc[i] *= inv;
    L::store(ra.at(Ob, seg_lb + r), acc, gl, f);
    if (gl == 0) { s_m[r] = m; s_inv[r] = inv; }
  };
  for (int r = vw; r < nseg; r += VW)
    if (s_rp[r + 1] == s_rp[r]) {
      float z[NR];
      zero(z);
      finish(r, kNeg, 0.f, z);
    }

  float q[NR], acc[NR];
  zero(q);
  zero(acc);
  float m_run = kNeg, l_run = 0.f;
  walk_pieces<CH>(
      b, s_rp,
      [&](int r) {
        L::load(q, ra.at(Qb, seg_lb + r), gl, f);
#pragma unroll
        for (int i = 0; i < NR; ++i) q[i] *= kLog2e;
        zero(acc);
        m_run = kNeg;
        l_run = 0.f;
      },
      [&](int base, int cnt) {
        int my_off = 0;  // float offset of the neighbour's staged row
        float my_w = 1.f, my_sc = 0.f;
        if (gl < cnt) {
          my_off = (__ldg(p.col_ind + base + gl) - seg_lb) * f;
          if (use_w) my_w = __ldg(p.val + base + gl);
        }
#pragma unroll
        for (int s = 0; s < CH; s += C) {
          if (s > 0 && !__any_sync(kFull, s < cnt)) break;
          float kk[C][NR], vv[C][NR], d[C];
          bool ok[C];
#pragma unroll
          for (int c = 0; c < C; ++c) {
            ok[c] = s + c < cnt;
            const int off = group_bcast<LPR>(my_off, s + c);  // beyond cnt: row 0 of the block, weight 0
            L::load_smem(kk[c], kl + off);
            L::load_smem(vv[c], vl + off);
          }
          float cm = kNeg;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            float dc = group_sum<LPR>(dot<NR>(q, kk[c]));
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
  {  // rows split over groups: the group with the first piece folds the following head pieces
    Slot<NR, LPR> mine(s_slot, vw, 1);
    const int seg = mine.seg();
    if (seg >= 0) {
      float m = mine.a(), l = mine.b(), a[NR];
#pragma unroll
      for (int i = 0; i < NR; ++i) a[i] = mine.v(i, gl);
      for (int v2 = vw + 1; v2 < VW; ++v2) {
        Slot<NR, LPR> s(s_slot, v2, 0);
        if (s.seg() != seg) break;
        float a2[NR];
#pragma unroll
        for (int i = 0; i < NR; ++i) a2[i] = s.v(i, gl);
        softmax_merge2<NR>(m, l, a, s.a(), s.b(), a2);
      }
      finish(seg, m, l, a);
    }
  }
  if (train) {  // scores -> probabilities (attn_edge of fused_gtconv_hyper.cu:146-149)
    __syncthreads();
    for (int i = b.E0 + threadIdx.x; i < b.E1; i += NW * 32) {
      const int rr = find_row(s_rp, nseg, i);
      attn[i] = fast_exp2(attn[i] - s_m[rr]) * s_inv[rr];
    }
  }
}

// Row side of the backward on one graph: V and K staged; see gt_bwd_row_kernel for the maths.
template <class L, int C, int NW>
__global__ void __launch_bounds__(NW * 32, 1) gt_block_bwd_row_kernel(const GtBlockBwdParams pp) {
  const GtBwdParams& p = pp.c;
  constexpr int NR = L::NR, LPR = L::LPR, G = L::G, VW = NW * G;
  constexpr int CH = ChunkOf<L>::kChunk;
  static_assert(L::kVec && CH % C == 0 && CH <= LPR, "vector layouts; chunking");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint64_t s_bar;
  const int f = p.f, mn = pp.b.max_nodes;
  float* sV = reinterpret_cast<float*>(smem_raw);
  float* sK = sV + (size_t)mn * f;
  float* s_slot = sK + (size_t)mn * f;
  int* s_rp = reinterpret_cast<int*>(s_slot + block_slot_floats<2 * NR, L, NW>());
  float* s_s = reinterpret_cast<float*>(s_rp + mn + 1);

  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int grp = lane / LPR, gl = lane % LPR, vw = w * G + grp;
  const int seg_lb = __ldg(pp.b.blk_ptr + blockIdx.x);
  const int nseg = __ldg(pp.b.blk_ptr + blockIdx.x + 1) - seg_lb;
  if (nseg <= 0) return;
  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
    stage_two(sV, sK, p.V, p.K, seg_lb, nseg, f, &s_bar);
  }
  const float* attn = p.attn;
  float2* gedge = reinterpret_cast<float2*>(p.grad_edge);  // {dS_e, p_e}
  const RowAddr<L> ra(1, f, 0, gl);
  const char* Gb = ra.base(p.dO);
  char* DQb = ra.base(p.dQ);
  const float* vl = sV + 4 * gl;
  const float* kl = sK + 4 * gl;

  slots_clear<2 * NR, LPR>(s_slot, vw, gl);
  const RowBlock b = segblock_init<G, NW>(s_rp, p.row_ptr, seg_lb, nseg, vw);
  mbar_wait(&s_bar, 0);

  auto write_dq = [&](int r, float s, const float (&dq)[NR]) {
    L::store(ra.at(DQb, seg_lb + r), dq, gl, f);
    if (gl == 0) s_s[r] = s;
  };
  for (int r = vw; r < nseg; r += VW)
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
        L::load(g, ra.at(Gb, seg_lb + r), gl, f);
        zero(acc2);
        s_part = 0.f;
        c_ref = 0.f;
        have_c = false;
      },
      [&](int base, int cnt) {
        int my_off = 0;
        float my_p = 0.f, my_t = 0.f, my_w = 1.f;
        if (gl < cnt) {
          my_off = (__ldg(p.col_ind + base + gl) - seg_lb) * f;
          my_p = __ldg(attn + base + gl);
          if (p.val) my_w = __ldg(p.val + base + gl);
        }
#pragma unroll
        for (int s = 0; s < CH; s += C) {
          if (s > 0 && !__any_sync(kFull, s < cnt)) break;
          float kk[C][NR], vv[C][NR], dA[C];
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const int off = group_bcast<LPR>(my_off, s + c);  // beyond cnt: row 0, probability 0
            L::load_smem(vv[c], vl + off);
            L::load_smem(kk[c], kl + off);
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
            if (p.val) {
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
  {
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
  for (int i = b.E0 + threadIdx.x; i < b.E1; i += NW * 32) {
    const int rr = find_row(s_rp, nseg, i);
    const float2 tp = gedge[i];
    float ds = fmaf(-s_s[rr], tp.y, tp.x);
    if (p.val) ds *= __ldg(p.val + i);
    gedge[i].x = ds;
  }
}

// Column side on one graph: dO and Q staged; segments are the CSC columns of the graph.
template <class L, int C, int NW>
__global__ void __launch_bounds__(NW * 32, 1) gt_block_bwd_col_kernel(const GtBlockBwdParams pp) {
  const GtBwdParams& p = pp.c;
  constexpr int NR = L::NR, LPR = L::LPR, G = L::G, VW = NW * G;
  constexpr int CH = ChunkOf<L>::kChunk;
  static_assert(L::kVec && CH % C == 0 && CH <= LPR, "vector layouts; chunking");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint64_t s_bar;
  const int f = p.f, mn = pp.b.max_nodes;
  float* sG = reinterpret_cast<float*>(smem_raw);
  float* sQ = sG + (size_t)mn * f;
  float* s_slot = sQ + (size_t)mn * f;
  int* s_cp = reinterpret_cast<int*>(s_slot + block_slot_floats<2 * NR, L, NW>());

  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int grp = lane / LPR, gl = lane % LPR, vw = w * G + grp;
  const int seg_lb = __ldg(pp.b.blk_ptr + blockIdx.x);
  const int nseg = __ldg(pp.b.blk_ptr + blockIdx.x + 1) - seg_lb;
  if (nseg <= 0) return;
  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
    stage_two(sG, sQ, p.dO, p.Q, seg_lb, nseg, f, &s_bar);
  }
  const float2* gedge = reinterpret_cast<const float2*>(p.grad_edge);
  const RowAddr<L> ra(1, f, 0, gl);
  char* DVb = ra.base(p.dV);
  char* DKb = ra.base(p.dK);
  const float* gol = sG + 4 * gl;
  const float* ql = sQ + 4 * gl;

  slots_clear<2 * NR, LPR>(s_slot, vw, gl);
  const RowBlock b = segblock_init<G, NW>(s_cp, p.col_ptr, seg_lb, nseg, vw);
  mbar_wait(&s_bar, 0);

  auto finish = [&](int c, float (&acc2)[2 * NR]) {
    float t[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) t[i] = acc2[i];
    L::store(ra.at(DVb, seg_lb + c), t, gl, f);
#pragma unroll
    for (int i = 0; i < NR; ++i) t[i] = acc2[NR + i];
    L::store(ra.at(DKb, seg_lb + c), t, gl, f);
  };
  for (int c = vw; c < nseg; c += VW)
    if (s_cp[c + 1] == s_cp[c]) {
      float z[2 * NR];
      zero(z);
      finish(c, z);
    }

  float acc2[2 * NR];
  zero(acc2);
  walk_pieces<CH>(
      b, s_cp, [&](int) { zero(acc2); },
      [&](int base, int cnt) {
        int my_off = 0;
        float my_p = 0.f, my_ds = 0.f;
        if (gl < cnt) {
          my_off = (__ldg(p.row_ind + base + gl) - seg_lb) * f;
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
            const int off = group_bcast<LPR>(my_off, s + c);  // beyond cnt: row 0, weights 0
            L::load_smem(go[c], gol + off);
            L::load_smem(qq[c], ql + off);
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
          finish(c0, acc2);
        } else {
          Slot<2 * NR, LPR> sl(s_slot, vw, first ? 1 : 0);
#pragma unroll
          for (int i = 0; i < 2 * NR; ++i) sl.v(i, gl) = acc2[i];
          if (gl == 0) { sl.a() = 0.f; sl.set_seg(c0); }
        }
      });
  __syncthreads();
  {
    Slot<2 * NR, LPR> mine(s_slot, vw, 1);
    const int seg = mine.seg();
    if (seg >= 0) {
      float a[2 * NR];
#pragma unroll
      for (int i = 0; i < 2 * NR; ++i) a[i] = mine.v(i, gl);
      for (int v2 = vw + 1; v2 < VW; ++v2) {
        Slot<2 * NR, LPR> s(s_slot, v2, 0);
        if (s.seg() != seg) break;
#pragma unroll
        for (int i = 0; i < 2 * NR; ++i) a[i] += s.v(i, gl);
      }
      finish(seg, a);
    }
  }
}

// every column id of a block's rows lies inside the block; blk_ptr is increasing from 0 to m.
// flag[0] != 0 on violation; flag[1] = largest block; flag[2]: see below.
static __global__ void block_check_kernel(int n_blocks, int m, const int* __restrict__ blk_ptr,
                                          const int* __restrict__ row_ptr, const int* __restrict__ col_ind,
                                          int* __restrict__ flag) {
  const int b = blockIdx.x;
  const int lo = blk_ptr[b], hi = blk_ptr[b + 1];
  if (threadIdx.x == 0) {
    if (lo < 0 || hi < lo || hi > m || (b == 0 && lo != 0) || (b == n_blocks - 1 && hi != m)) atomicOr(flag, 1);
    else atomicMax(flag + 1, hi - lo);
  }
  if (lo < 0 || hi < lo || hi > m) return;
  const int e0 = row_ptr[lo], e1 = row_ptr[hi];
  for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
    const int c = col_ind[e];
    if (c < lo || c >= hi) { atomicOr(flag, 1); break; }
  }
  // flag[2] != 0: some row's column ids are not strictly ascending (unsorted or duplicate edges);
  // the dense kernels address attn_edge by the rank of a column inside its row and need that order
  for (int r = lo + threadIdx.x; r < hi; r += blockDim.x) {
    const int rs = row_ptr[r], re = row_ptr[r + 1];
    int prev = -1;
    for (int e = rs; e < re; ++e) {
      const int c = col_ind[e];
      if (c <= prev) { atomicOr(flag + 2, 1); break; }
      prev = c;
    }
  }
}

}  // namespace dfgnn
