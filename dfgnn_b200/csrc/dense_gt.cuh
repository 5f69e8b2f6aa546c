// dense_gt.cuh -- tensor-core GT kernels for block-diagonal batches whose graphs are DENSE
// (PATTERN-shaped: ~119 nodes, 43 % of the n x n entries present).
//
// For such a graph the per-edge formulation executes ~40 warp instructions per edge (dot product
// over lane groups, shuffle reductions, online-softmax bookkeeping): the row-block and the
// shared-memory-staged kernels are instruction-issue bound, not memory bound.  A dense masked
// tile does the same work as two small GEMMs per 16-row tile,
//       S = Q_tile K^T   (16 x n x f)      O = P V   (16 x n x f),
// on the tensor cores: warp-level mma.sync.m16n8k8 TF32 with the 3xTF32 split
// (x = hi + lo, a*b ~ hi*hi + hi*lo + lo*hi, fp32 accumulation; relative error ~2^-20 per product),
// which keeps the 1e-4 / 1e-5 parity bar of the fp32 kernels.  2.3x the multiply-adds of the sparse
// form, ~10x fewer instructions.  (tcgen05 is not used here: its operands come straight from
// shared / tensor memory, so the hi/lo split would need every operand twice in shared memory; the
// warp-level path splits in registers.  Measured on B200: 2.1 clk per mma.m16n8k8 per SM,
// tools/mma_bench.cu.)
//
// One CTA per graph: K and V rows are staged in shared memory by per-row TMA bulk copies into a
// row stride of f + 4 floats (conflict-free B-fragment reads); the adjacency becomes a bit mask
// (n x n bits) built from the CSR; each warp owns 16-row tiles of Q and runs a flash-style online
// softmax over 32-column blocks.  The C fragment of S is reused as the A fragment of the second
// product by permuting the summation index (C holds columns {2t, 2t+1}, A wants {t, t+4}: read V
// rows 2t and 2t+1 instead), so P never leaves registers.
//
// Maths and outputs are those of dot_fwd_kernel (out, attn_edge); the reference counterpart is
// fused_gtconv_hyper.cu:31-163.
#pragma once

#include <type_traits>

#include "block_gt.cuh"

namespace dfgnn {

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// x = hi + lo with hi = x truncated to TF32 (10 mantissa bits, one LOP3 -- cvt.rna.tf32.f32 is a
// seven-instruction sequence on sm_100a); lo = x - hi is exact in fp32 and is truncated to TF32 by
// the tensor core (error <= 2^-20 |x|, like the dropped lo * lo term)
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
// d += a * b in 3xTF32
__device__ __forceinline__ void mma_3x(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                       const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
  mma_tf32(d, al, bh);
  mma_tf32(d, ah, bl);
  mma_tf32(d, ah, bh);
}

#ifndef DFGNN_DENSE_CB
#define DFGNN_DENSE_CB 32
#endif
constexpr int kDenseCB = DFGNN_DENSE_CB;  // columns per softmax block (measured: 32 -> 0.354 ms, 64 -> 0.379 ms inference on PATTERN)

// shared-memory carve of the dense kernels
struct DenseSmem {
  int mn8, mn16, W;
  size_t off_b, off_mask, off_pre, off_rp, off_m, off_inv, bytes;
  __host__ __device__ DenseSmem(int max_nodes, int f) {
    mn8 = (max_nodes + 7) & ~7;
    mn16 = (max_nodes + 15) & ~15;
    W = (mn8 + 31) >> 5;
    const size_t ld = f + 4;
    size_t o = (size_t)mn8 * ld * 4;       // operand block A
    off_b = o;
    o += (size_t)mn8 * ld * 4;             // operand block B
    off_mask = o;
    o += (size_t)mn16 * W * 4;
    off_pre = o;
    o += (size_t)mn16 * W * 4;
    off_rp = o;
    o += (size_t)(max_nodes + 1) * 4;
    off_m = o;
    o += (size_t)mn16 * 4;
    off_inv = o;
    o += (size_t)mn16 * 4;
    bytes = (o + 15) & ~(size_t)15;
  }
};

// Stage rows [lb, lb + n) of two [*, F] matrices into padded shared-memory rows (stride F + 4) with
// one TMA bulk copy per row, zero the rows up to the next multiple of 8, and build the adjacency
// bit mask (+ per-word prefix popcounts when RANKS) of the graph from its CSR segment pointers.
// All threads call it; ends with the data visible to everyone.
template <int F, int NW, bool RANKS>
__device__ __forceinline__ void dense_stage(const DenseSmem& L, unsigned char* smem, uint64_t* bar, const float* A,
                                            const float* B, const int* __restrict__ seg_ptr,
                                            const int* __restrict__ idx, int lb, int n) {
  constexpr int LD = F + 4;
  float* sA = reinterpret_cast<float*>(smem);
  float* sB = reinterpret_cast<float*>(smem + L.off_b);
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(smem + L.off_mask);
  uint32_t* s_pre = reinterpret_cast<uint32_t*>(smem + L.off_pre);
  int* s_rp = reinterpret_cast<int*>(smem + L.off_rp);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n8 = (n + 7) & ~7, n16 = (n + 15) & ~15;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    mbar_expect_tx(bar, 2u * (uint32_t)n * F * 4u);
  }
  __syncthreads();
  if (w == 0) {
    for (int r = lane; r < n; r += 32) {
      bulk_g2s(sA + (size_t)r * LD, A + (size_t)(lb + r) * F, F * 4, bar);
      bulk_g2s(sB + (size_t)r * LD, B + (size_t)(lb + r) * F, F * 4, bar);
    }
  }
  for (int i = threadIdx.x; i < (n8 - n) * F; i += NW * 32) {
    const int r = n + i / F, c = i % F;
    sA[(size_t)r * LD + c] = 0.f;
    sB[(size_t)r * LD + c] = 0.f;
  }
  for (int i = threadIdx.x; i <= n; i += NW * 32) s_rp[i] = __ldg(seg_ptr + lb + i);
  for (int i = threadIdx.x; i < n16 * L.W; i += NW * 32) s_mask[i] = 0u;
  __syncthreads();
  // adjacency bits: a warp per row, four rows (eight coalesced index loads) in flight per warp
  for (int r0 = w; r0 < n; r0 += 4 * NW) {
    int j[4][2];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + u * NW;
      const int b = r < n ? s_rp[r] : 0, e = r < n ? s_rp[r + 1] : 0;
      j[u][0] = b + lane < e ? __ldg(idx + b + lane) - lb : -1;
      j[u][1] = b + lane + 32 < e ? __ldg(idx + b + lane + 32) - lb : -1;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + u * NW;
#pragma unroll
      for (int v = 0; v < 2; ++v)
        if (j[u][v] >= 0) atomicOr(s_mask + r * L.W + (j[u][v] >> 5), 1u << (j[u][v] & 31));
      if (r < n)
        for (int e = s_rp[r] + 64 + lane; e < s_rp[r + 1]; e += 32) {  // rows of more than 64 entries
          const int jj = __ldg(idx + e) - lb;
          atomicOr(s_mask + r * L.W + (jj >> 5), 1u << (jj & 31));
        }
    }
  }
  __syncthreads();
  if (RANKS) {
    for (int r = threadIdx.x; r < n16; r += NW * 32) {
      uint32_t acc = 0;
      for (int k = 0; k < L.W; ++k) {
        s_pre[r * L.W + k] = acc;
        acc += __popc(s_mask[r * L.W + k]);
      }
    }
    __syncthreads();
  }
  mbar_wait(bar, 0);
}

template <int F, int NW>
__global__ void __launch_bounds__(NW * 32, 1) gt_dense_fwd_kernel(const GtBlockFwdParams pp) {
  const DotFwdParams& p = pp.c;
  constexpr int LD = F + 4, KS = F / 8, NT = F / 8, CB = kDenseCB, NB = CB / 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint64_t s_bar;
  const DenseSmem L(pp.b.max_nodes, F);
  const float* sK = reinterpret_cast<const float*>(smem_raw);
  const float* sV = reinterpret_cast<const float*>(smem_raw + L.off_b);
  const uint32_t* s_mask = reinterpret_cast<const uint32_t*>(smem_raw + L.off_mask);
  const uint32_t* s_pre = reinterpret_cast<const uint32_t*>(smem_raw + L.off_pre);
  const int* s_rp = reinterpret_cast<const int*>(smem_raw + L.off_rp);
  float* s_m = reinterpret_cast<float*>(smem_raw + L.off_m);
  float* s_inv = reinterpret_cast<float*>(smem_raw + L.off_inv);

  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  const int lb = __ldg(pp.b.blk_ptr + blockIdx.x);
  const int n = __ldg(pp.b.blk_ptr + blockIdx.x + 1) - lb;
  if (n <= 0) return;
  const bool train = p.attn != nullptr;
  dense_stage<F, NW, true>(L, smem_raw, &s_bar, p.K, p.V, p.row_ptr, p.col_ind, lb, n);
  const int n8 = (n + 7) & ~7, W = L.W;

  for (int r0 = w * 16; r0 < n; r0 += NW * 16) {
    const int ra = r0 + g, rb = r0 + g + 8;  // the two rows of this thread's fragments
    // Q fragments (raw fp32, scaled into the base-2 exponent domain), rows beyond the graph are 0
    float qf[KS][4];
    {
      const float* qa = p.Q + (size_t)(lb + ra) * F + t4;
      const float* qb = p.Q + (size_t)(lb + rb) * F + t4;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        qf[ks][0] = ra < n ? __ldg(qa + 8 * ks) * kLog2e : 0.f;
        qf[ks][1] = rb < n ? __ldg(qb + 8 * ks) * kLog2e : 0.f;
        qf[ks][2] = ra < n ? __ldg(qa + 8 * ks + 4) * kLog2e : 0.f;
        qf[ks][3] = rb < n ? __ldg(qb + 8 * ks + 4) * kLog2e : 0.f;
      }
    }
    float o[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_a = kNeg, m_b = kNeg, l_a = 0.f, l_b = 0.f;
    const uint32_t* mrow_a = s_mask + ra * W;  // rows < mn16: always inside the mask array
    const uint32_t* mrow_b = s_mask + rb * W;

    // one 64-column block; FULL = all 8 n8-tiles present (branch-free inner loops)
    auto column_block = [&](int jb, int ntile, auto full_tag) {
      constexpr bool FULL = decltype(full_tag)::value;
      float s[NB][4];
#pragma unroll
      for (int i = 0; i < NB; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      // ---- S = Q K^T ------------------------------------------------------------------
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t ah[4], al[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_tf32(qf[ks][i], ah[i], al[i]);
#pragma unroll
        for (int nt = 0; nt < NB; ++nt) {
          if (FULL || nt < ntile) {
            const float* kp = sK + (size_t)(jb + 8 * nt + g) * LD + 8 * ks + t4;
            uint32_t bh[2], bl[2];
            split_tf32(kp[0], bh[0], bl[0]);
            split_tf32(kp[4], bh[1], bl[1]);
            mma_3x(s[nt], ah, al, bh, bl);
          }
        }
      }
      // ---- mask, raw scores out (training), online softmax ------------------------------
      float bm_a = kNeg, bm_b = kNeg;
#pragma unroll
      for (int nt = 0; nt < NB; ++nt) {
        if (FULL || nt < ntile) {
          const int c0 = jb + 8 * nt + 2 * t4, wd = c0 >> 5, bit = c0 & 31;
          const uint32_t ma = mrow_a[wd], mb = mrow_b[wd];
          const bool v0 = (ma >> bit) & 1u, v1 = (ma >> (bit + 1)) & 1u;
          const bool v2 = (mb >> bit) & 1u, v3 = (mb >> (bit + 1)) & 1u;
          if (train) {  // attn_edge[e] = raw score; e = CSR position of (row, column)
            const uint32_t low = (1u << bit) - 1u;
            if (v0 | v1) {
              const int e = s_rp[ra] + (int)s_pre[ra * W + wd] + __popc(ma & low);
              if (v0) p.attn[e] = s[nt][0];
              if (v1) p.attn[e + (v0 ? 1 : 0)] = s[nt][1];
            }
            if (v2 | v3) {
              const int e = s_rp[rb] + (int)s_pre[rb * W + wd] + __popc(mb & low);
              if (v2) p.attn[e] = s[nt][2];
              if (v3) p.attn[e + (v2 ? 1 : 0)] = s[nt][3];
            }
          }
          s[nt][0] = v0 ? s[nt][0] : kNeg;
          s[nt][1] = v1 ? s[nt][1] : kNeg;
          s[nt][2] = v2 ? s[nt][2] : kNeg;
          s[nt][3] = v3 ? s[nt][3] : kNeg;
          bm_a = fmaxf(bm_a, fmaxf(s[nt][0], s[nt][1]));
          bm_b = fmaxf(bm_b, fmaxf(s[nt][2], s[nt][3]));
        }
      }
      bm_a = fmaxf(bm_a, __shfl_xor_sync(kFull, bm_a, 1));
      bm_a = fmaxf(bm_a, __shfl_xor_sync(kFull, bm_a, 2));
      bm_b = fmaxf(bm_b, __shfl_xor_sync(kFull, bm_b, 1));
      bm_b = fmaxf(bm_b, __shfl_xor_sync(kFull, bm_b, 2));
      const float mn_a = fmaxf(m_a, bm_a), mn_b = fmaxf(m_b, bm_b);
      const float sc_a = fast_exp2(m_a - mn_a), sc_b = fast_exp2(m_b - mn_b);
      m_a = mn_a;
      m_b = mn_b;
      l_a *= sc_a;
      l_b *= sc_b;
#pragma unroll
      for (int i = 0; i < NT; ++i) {
        o[i][0] *= sc_a; o[i][1] *= sc_a; o[i][2] *= sc_b; o[i][3] *= sc_b;
      }
      // probabilities (masked entries: exactly 0; a row without edges so far keeps m = kNeg and
      // every entry masked)
#pragma unroll
      for (int nt = 0; nt < NB; ++nt) {
        if (FULL || nt < ntile) {
          s[nt][0] = s[nt][0] > 0.5f * kNeg ? fast_exp2(s[nt][0] - mn_a) : 0.f;
          s[nt][1] = s[nt][1] > 0.5f * kNeg ? fast_exp2(s[nt][1] - mn_a) : 0.f;
          s[nt][2] = s[nt][2] > 0.5f * kNeg ? fast_exp2(s[nt][2] - mn_b) : 0.f;
          s[nt][3] = s[nt][3] > 0.5f * kNeg ? fast_exp2(s[nt][3] - mn_b) : 0.f;
          l_a += s[nt][0] + s[nt][1];
          l_b += s[nt][2] + s[nt][3];
        }
      }
      // ---- O += P V : the C fragment of S is the A fragment over the permuted k index -------
#pragma unroll
      for (int kk = 0; kk < NB; ++kk) {
        if (FULL || kk < ntile) {
          uint32_t ah[4], al[4];
          split_tf32(s[kk][0], ah[0], al[0]);  // (row a, k-slot t)     = column 2t
          split_tf32(s[kk][2], ah[1], al[1]);  // (row b, k-slot t)
          split_tf32(s[kk][1], ah[2], al[2]);  // (row a, k-slot t + 4) = column 2t + 1
          split_tf32(s[kk][3], ah[3], al[3]);  // (row b, k-slot t + 4)
          const float* vp = sV + (size_t)(jb + 8 * kk + 2 * t4) * LD + g;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            uint32_t bh[2], bl[2];
            split_tf32(vp[8 * nt], bh[0], bl[0]);
            split_tf32(vp[8 * nt + LD], bh[1], bl[1]);
            mma_3x(o[nt], ah, al, bh, bl);
          }
        }
      }
    };
    for (int jb = 0; jb < n8; jb += CB) {
      const int ntile = min(NB, (n8 - jb) >> 3);  // n8-tiles of this column block (warp-uniform)
      if (ntile == NB) column_block(jb, NB, std::true_type{});
      else column_block(jb, ntile, std::false_type{});
    }
    // ---- finish the tile ----------------------------------------------------------------------
    l_a += __shfl_xor_sync(kFull, l_a, 1);
    l_a += __shfl_xor_sync(kFull, l_a, 2);
    l_b += __shfl_xor_sync(kFull, l_b, 1);
    l_b += __shfl_xor_sync(kFull, l_b, 2);
    const float inv_a = l_a > 0.f ? 1.f / l_a : 0.f, inv_b = l_b > 0.f ? 1.f / l_b : 0.f;
    if (t4 == 0) {
      if (ra < n) { s_m[ra] = m_a; s_inv[ra] = inv_a; }
      if (rb < n) { s_m[rb] = m_b; s_inv[rb] = inv_b; }
    }
    float* oa = p.out + (size_t)(lb + ra) * F + 2 * t4;
    float* ob = p.out + (size_t)(lb + rb) * F + 2 * t4;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (ra < n) *reinterpret_cast<float2*>(oa + 8 * nt) = make_float2(o[nt][0] * inv_a, o[nt][1] * inv_a);
      if (rb < n) *reinterpret_cast<float2*>(ob + 8 * nt) = make_float2(o[nt][2] * inv_b, o[nt][3] * inv_b);
    }
  }
  if (train) {  // scores -> probabilities (attn_edge of fused_gtconv_hyper.cu:146-149)
    __syncthreads();
    for (int r = w; r < n; r += NW) {  // a warp per row: coalesced, no search
      const float mr = s_m[r], ir = s_inv[r];
      for (int i = s_rp[r] + lane; i < s_rp[r + 1]; i += 32) p.attn[i] = fast_exp2(p.attn[i] - mr) * ir;
    }
  }
}

}  // namespace dfgnn
