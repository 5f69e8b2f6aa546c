// dense_tc.cu -- GT attention for block-diagonal batches of DENSE small graphs on the 5th-generation
// tensor cores (tcgen05, accumulators in tensor memory).
//
// A PATTERN-shaped batch (1024 graphs of ~119 nodes, 43 % of each n x n block present) is, per graph,
// a small masked dense attention: S = Q K^T, P = softmax over the row's neighbours, O = P V.  The
// per-edge kernels (fwd_kernels.cuh, block_gt.cuh) spend ~40 warp instructions and 1 KB of L1 /
// shared-memory traffic per edge there and are load/store-unit bound at 0.45 ms per forward; as two
// GEMMs per 128-row tile the same work is 96 tcgen05.mma instructions and the kernel is bound by
// reading Q, K, V once and writing O and attn_edge once (0.3 GB per forward on that batch).
//
// fp32 parity (1e-4 relative) from TF32 tensor cores by the 3xTF32 split: x = hi + lo with hi = x
// truncated to TF32, a b ~ a_lo b_hi + a_hi b_lo + a_hi b_hi, fp32 accumulation in tensor memory.
//
// One persistent CTA per SM walks graphs; a graph of n <= 256 nodes is one or two tiles of 128 query
// rows.  Roles (warp specialised, mbarriers only):
//   warps 0-3    softmax / epilogue: thread r owns tile row r = TMEM lane r.  Builds the row's adjacency
//                bit mask from the CSR while the first product runs, then reads S from tensor memory in
//                32-column pieces: pass 1 row max, pass 2 p = 2^(s - max) (unnormalised) split into the A
//                images of the second product, pass 3 (training) normalised probabilities to attn_edge;
//                finally O * (1 / row sum) -> out through a shared-memory staging buffer (coalesced).
//   warp  4      MMA issue (one lane).  Product 1: S[128 x keys] += Q_slice K_slice^T over four 32-wide
//                slices of the feature dimension (two key halves when n > 128).  Product 2:
//                O[128 x 128] += P_slice V_slice over 32-key slices.  3 MMAs (lo*hi, hi*lo, hi*hi) per K = 8.
//   warps 5-16   three loader groups of 4 warps, group g fills ring slot g: global -> registers (issued
//                before the slot is free, so up to three stages of loads are in flight) -> hi / lo split ->
//                canonical K-major images.  Product 1 stages carry Q (A) and K (B) slices; product 2
//                stages carry V TRANSPOSED (B: rows = features, k = keys; read as coalesced scalars,
//                stored as one float4 of four keys) while the softmax threads supply P (A).
// Ring: 3 slots x {A_hi, A_lo, B_hi, B_lo} x (128 rows x 32 floats) = 192 KB.  Tensor memory: S in
// columns [0, 256), O in [256, 384).
//
// Requirements (checked by the block plan, formats.py / dfgnn_block_plan_check): h == 1, f == 128,
// unweighted scores, graphs of at most 256 nodes, column ids strictly ascending inside every row (no
// duplicate edges: a dense mask cannot count an edge twice).  Reference counterpart:
// fused_gtconv_hyper.cu:31-163 (forward) and, for the maths, DFGNN/layers/GT/gtconv_layer.py:28-45.
#include "abi_common.h"
#include "block_gt.cuh"
#include "tc_common.cuh"

namespace dfgnn {

constexpr int kTcM = 128;                        // rows per tile (UMMA M) and rows of every image
constexpr int kTcF = 128;                        // feature width
constexpr int kTcKS = 32;                        // contraction slice of one ring stage
constexpr int kTcSlots = 3;
constexpr int kTcImg = kTcM * kTcKS * 4;         // one image: 128 rows x 32 floats = 16 KB
constexpr int kTcSlotBytes = 4 * kTcImg;         // A_hi | A_lo | B_hi | B_lo
constexpr int kTcMaxNodes = 256;
constexpr int kTcSoftWarps = 4, kTcLoadGroups = kTcSlots, kTcGroupWarps = 4;
constexpr int kTcGroupThreads = kTcGroupWarps * 32;
constexpr int kTcThreads = (kTcSoftWarps + 1 + kTcLoadGroups * kTcGroupWarps) * 32;  // 544
constexpr int kTcMaskW = kTcMaxNodes / 32;       // mask words per row
constexpr int kTcStgLd = 36;                     // epilogue staging row stride (floats)
constexpr uint32_t kTcLBO = kTcM * 16, kTcSBO = 128;
constexpr int kTcColO = 256;                     // first tensor-memory column of O

constexpr size_t kTcOffMask = (size_t)kTcSlots * kTcSlotBytes;
constexpr size_t kTcOffStage = kTcOffMask + (size_t)kTcM * kTcMaskW * 4;
constexpr size_t kTcOffRp = kTcOffStage + (size_t)kTcM * kTcStgLd * 4;
constexpr size_t kTcSmemBytes = kTcOffRp + (size_t)(kTcM + 4) * 4;

// barrier among the 128 threads of the softmax group
__device__ __forceinline__ void soft_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

__global__ void __launch_bounds__(kTcThreads, 1) gt_dense_tc_fwd_kernel(const GtBlockFwdParams pp) {
  const DotFwdParams& p = pp.c;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full_b[kTcSlots], full_a[kTcSlots], empty[kTcSlots], s_full, s_free, o_full, o_free;
  __shared__ uint32_t s_tmem;
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(smem + kTcOffMask);
  float* s_stg = reinterpret_cast<float*>(smem + kTcOffStage);
  int* s_rp = reinterpret_cast<int*>(smem + kTcOffRp);

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < kTcSlots; ++i) {
      mbar_init(&full_b[i], kTcGroupThreads);
      mbar_init(&full_a[i], kTcSoftWarps * 32);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&s_full, 1);
    mbar_init(&o_full, 1);
    mbar_init(&s_free, kTcSoftWarps * 32);
    mbar_init(&o_free, kTcSoftWarps * 32);
    mbar_fence_init();
  }
  if (w == 0) {  // all 512 columns: S (256) + O (128), one CTA per SM
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const int nb = pp.b.n_blocks;

  auto slot_img = [&](int slot, int which) -> float4* {  // which: 0 A_hi, 1 A_lo, 2 B_hi, 3 B_lo
    return reinterpret_cast<float4*>(smem + (size_t)slot * kTcSlotBytes + (size_t)which * kTcImg);
  };

  if (w < kTcSoftWarps) {
    // =============================== softmax / epilogue ===========================================
    const int r = tid;  // tile row = TMEM lane
    const uint32_t lane_base = tmem + ((uint32_t)(w * 32) << 16);
    const bool train = p.attn != nullptr;
    uint32_t sc = 0, tc = 0;
    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
      const int lb = __ldg(pp.b.blk_ptr + b), n = __ldg(pp.b.blk_ptr + b + 1) - lb;
      if (n <= 0) continue;
      const int MT = (n + kTcM - 1) / kTcM, NS2 = (n + 31) >> 5;
      for (int mt = 0; mt < MT; ++mt, ++tc) {
        const int row = mt * kTcM + r;
        // ---- segment pointers and adjacency bits of the tile's rows (while product 1 runs) ----
        s_rp[r] = __ldg(p.row_ptr + lb + min(row, n));
        if (r == 0) s_rp[kTcM] = __ldg(p.row_ptr + lb + min(mt * kTcM + kTcM, n));
        {
          uint4* mz = reinterpret_cast<uint4*>(s_mask + r * kTcMaskW);
          mz[0] = make_uint4(0u, 0u, 0u, 0u);
          mz[1] = make_uint4(0u, 0u, 0u, 0u);
        }
        soft_bar();
        for (int rr0 = w * 32; rr0 < w * 32 + 32; rr0 += 4) {  // a warp per row, four rows in flight
          int j[4][2];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int e0 = s_rp[rr0 + u], e1 = s_rp[rr0 + u + 1];
            j[u][0] = e0 + lane < e1 ? __ldg(p.col_ind + e0 + lane) - lb : -1;
            j[u][1] = e0 + lane + 32 < e1 ? __ldg(p.col_ind + e0 + lane + 32) - lb : -1;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint32_t* mrow = s_mask + (rr0 + u) * kTcMaskW;
#pragma unroll
            for (int v = 0; v < 2; ++v)
              if (j[u][v] >= 0) atomicOr(mrow + (j[u][v] >> 5), 1u << (j[u][v] & 31));
            for (int e = s_rp[rr0 + u] + 64 + lane; e < s_rp[rr0 + u + 1]; e += 32) {
              const int jj = __ldg(p.col_ind + e) - lb;
              atomicOr(mrow + (jj >> 5), 1u << (jj & 31));
            }
          }
        }
        soft_bar();
        uint32_t mw[kTcMaskW];
        {
          const uint4 m0 = *reinterpret_cast<const uint4*>(s_mask + r * kTcMaskW);
          const uint4 m1 = *reinterpret_cast<const uint4*>(s_mask + r * kTcMaskW + 4);
          mw[0] = m0.x; mw[1] = m0.y; mw[2] = m0.z; mw[3] = m0.w;
          mw[4] = m1.x; mw[5] = m1.y; mw[6] = m1.z; mw[7] = m1.w;
        }
        const uint32_t sc2 = sc + 4u * MT;  // stage index of the first product-2 stage of this tile
        // ---- S is complete ----------------------------------------------------------------------
        mbar_wait(&s_full, tc & 1u);
        tc_fence_after();
        float mx = kNeg;
#pragma unroll
        for (int cb = 0; cb < kTcMaskW; ++cb) {
          if (cb < NS2) {
            float s[32];
            tmem_ld32(lane_base + cb * 32, s);
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, ((mw[cb] >> i) & 1u) ? s[i] : kNeg);
          }
        }
        float l = 0.f;
#pragma unroll
        for (int cb = 0; cb < kTcMaskW; ++cb) {
          if (cb < NS2) {
            float s[32];
            tmem_ld32(lane_base + cb * 32, s);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              s[i] = ((mw[cb] >> i) & 1u) ? fast_exp2(s[i] - mx) : 0.f;
              l += s[i];
            }
            const uint32_t st = sc2 + cb, slot = st % kTcSlots, k = st / kTcSlots;
            mbar_wait(&empty[slot], (k & 1u) ^ 1u);
            float4* a_hi = slot_img(slot, 0);
            float4* a_lo = slot_img(slot, 1);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              float4 hi, lo;
              split4(make_float4(s[4 * c], s[4 * c + 1], s[4 * c + 2], s[4 * c + 3]), hi, lo);
              a_hi[c * kTcM + r] = hi;
              a_lo[c * kTcM + r] = lo;
            }
            fence_proxy_async();
            mbar_arrive(&full_a[slot]);
          }
        }
        const float inv = l > 0.f ? 1.f / l : 0.f;
        if (train) {  // attn_edge in CSR order: the row's neighbours are its set mask bits, ascending
          float* ap = p.attn + s_rp[r];
#pragma unroll
          for (int cb = 0; cb < kTcMaskW; ++cb) {
            if (cb < NS2) {  // warp-uniform: tcgen05.ld is a warp-collective
              float s[32];
              tmem_ld32(lane_base + cb * 32, s);
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if ((mw[cb] >> i) & 1u) *ap++ = fast_exp2(s[i] - mx) * inv;
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&s_free);
        // ---- O is complete ----------------------------------------------------------------------
        mbar_wait(&o_full, tc & 1u);
        tc_fence_after();
        float* obase = p.out + (size_t)(lb + mt * kTcM) * kTcF;
        const int rows_here = min(kTcM, n - mt * kTcM);
#pragma unroll 1
        for (int cq = 0; cq < kTcF / 32; ++cq) {
          float y[32];
          tmem_ld32(lane_base + kTcColO + cq * 32, y);
          if (cq == kTcF / 32 - 1) {
            tc_fence_before();
            mbar_arrive(&o_free);
          }
          float4* so = reinterpret_cast<float4*>(s_stg + (size_t)r * kTcStgLd);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            so[i] = make_float4(y[4 * i] * inv, y[4 * i + 1] * inv, y[4 * i + 2] * inv, y[4 * i + 3] * inv);
          soft_bar();
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int i = tid + u * 128, rr = i >> 3, c4 = i & 7;
            if (rr < rows_here)
              *reinterpret_cast<float4*>(obase + (size_t)rr * kTcF + cq * 32 + 4 * c4) =
                  *reinterpret_cast<const float4*>(s_stg + (size_t)rr * kTcStgLd + 4 * c4);
          }
          soft_bar();
        }
        sc = sc2 + NS2;
      }
    }
  } else if (w == kTcSoftWarps) {
    // =============================== MMA issue ====================================================
    if (lane == 0) {
      uint32_t sc = 0, tc = 0, ka = 0;  // ka: bit s = parity of full_a[s]
      constexpr uint32_t idesc2 = umma_idesc_tf32(kTcM, kTcF);
      const uint32_t ring = smem_u32(smem);
      for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        const int lb = __ldg(pp.b.blk_ptr + b), n = __ldg(pp.b.blk_ptr + b + 1) - lb;
        if (n <= 0) continue;
        const int MT = (n + kTcM - 1) / kTcM, NS2 = (n + 31) >> 5;
        for (int mt = 0; mt < MT; ++mt, ++tc) {
          mbar_wait(&s_free, (tc & 1u) ^ 1u);  // the softmax threads have read S of the previous tile
          tc_fence_after();
          for (int kh = 0; kh < MT; ++kh) {
            const int keys = min(kTcM, n - kh * kTcM);
            const uint32_t idesc1 = umma_idesc_tf32(kTcM, (keys + 15) & ~15);
            const uint32_t d = tmem + kh * kTcM;
            for (int q = 0; q < kTcF / kTcKS; ++q, ++sc) {
              const uint32_t slot = sc % kTcSlots, k = sc / kTcSlots;
              mbar_wait(&full_b[slot], k & 1u);
              tc_fence_after();
              const uint32_t base = ring + slot * kTcSlotBytes;
#pragma unroll
              for (int ks = 0; ks < kTcKS / 8; ++ks) {
                const uint32_t off = ks * 2 * kTcLBO;
                const uint64_t dah = umma_desc_kmajor(base + off, kTcLBO, kTcSBO);
                const uint64_t dal = umma_desc_kmajor(base + kTcImg + off, kTcLBO, kTcSBO);
                const uint64_t dbh = umma_desc_kmajor(base + 2 * kTcImg + off, kTcLBO, kTcSBO);
                const uint64_t dbl = umma_desc_kmajor(base + 3 * kTcImg + off, kTcLBO, kTcSBO);
                umma_tf32(d, dal, dbh, idesc1, (q | ks) != 0 ? 1u : 0u);
                umma_tf32(d, dah, dbl, idesc1, 1u);
                umma_tf32(d, dah, dbh, idesc1, 1u);
              }
              umma_commit(&empty[slot]);
            }
          }
          umma_commit(&s_full);
          mbar_wait(&o_free, (tc & 1u) ^ 1u);  // O of the previous tile has been read
          tc_fence_after();
          for (int s = 0; s < NS2; ++s, ++sc) {
            const uint32_t slot = sc % kTcSlots, k = sc / kTcSlots;
            mbar_wait(&full_b[slot], k & 1u);
            mbar_wait(&full_a[slot], (ka >> slot) & 1u);
            ka ^= 1u << slot;
            tc_fence_after();
            const uint32_t base = ring + slot * kTcSlotBytes;
#pragma unroll
            for (int ks = 0; ks < kTcKS / 8; ++ks) {
              const uint32_t off = ks * 2 * kTcLBO;
              const uint64_t dah = umma_desc_kmajor(base + off, kTcLBO, kTcSBO);
              const uint64_t dal = umma_desc_kmajor(base + kTcImg + off, kTcLBO, kTcSBO);
              const uint64_t dbh = umma_desc_kmajor(base + 2 * kTcImg + off, kTcLBO, kTcSBO);
              const uint64_t dbl = umma_desc_kmajor(base + 3 * kTcImg + off, kTcLBO, kTcSBO);
              umma_tf32(tmem + kTcColO, dal, dbh, idesc2, (s | ks) != 0 ? 1u : 0u);
              umma_tf32(tmem + kTcColO, dah, dbl, idesc2, 1u);
              umma_tf32(tmem + kTcColO, dah, dbh, idesc2, 1u);
            }
            umma_commit(&empty[slot]);
          }
          umma_commit(&o_full);
        }
      }
    }
  } else {
    // =============================== loaders ======================================================
    const int g = (w - kTcSoftWarps - 1) / kTcGroupWarps;
    const int lt = tid - (kTcSoftWarps + 1 + g * kTcGroupWarps) * 32;  // thread inside the group
    float4* const a_hi = slot_img(g, 0);
    float4* const a_lo = slot_img(g, 1);
    float4* const b_hi = slot_img(g, 2);
    float4* const b_lo = slot_img(g, 3);
    uint32_t sc = 0;
    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
      const int lb = __ldg(pp.b.blk_ptr + b), n = __ldg(pp.b.blk_ptr + b + 1) - lb;
      if (n <= 0) continue;
      const int MT = (n + kTcM - 1) / kTcM, NS2 = (n + 31) >> 5;
      for (int mt = 0; mt < MT; ++mt) {
        // ---- product 1: Q slice (A, scaled into the base-2 exponent domain) and K slice (B) ----
        for (int kh = 0; kh < MT; ++kh) {
          const int rows_a = min(kTcM, n - mt * kTcM), rows_b = min(kTcM, n - kh * kTcM);
          for (int q = 0; q < kTcF / kTcKS; ++q, ++sc) {
            if ((int)(sc % kTcSlots) != g) continue;
            const float4* qsrc = reinterpret_cast<const float4*>(p.Q + (size_t)(lb + mt * kTcM) * kTcF + q * kTcKS);
            const float4* ksrc = reinterpret_cast<const float4*>(p.K + (size_t)(lb + kh * kTcM) * kTcF + q * kTcKS);
            float4 xa[8], xb[8];
            // float4 i covers row (i & 7) + 8 * (i >> 6), chunk (i >> 3) & 7: a warp reads 8 rows x 64
            // contiguous bytes and writes 4 x 128 contiguous bytes of the chunk-major images
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int i = lt + u * kTcGroupThreads, rr = (i & 7) + 8 * (i >> 6), c = (i >> 3) & 7;
              xa[u] = rr < rows_a ? __ldg(qsrc + (size_t)rr * (kTcF / 4) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
              xb[u] = rr < rows_b ? __ldg(ksrc + (size_t)rr * (kTcF / 4) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            mbar_wait(&empty[g], ((sc / kTcSlots) & 1u) ^ 1u);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int i = lt + u * kTcGroupThreads, rr = (i & 7) + 8 * (i >> 6), c = (i >> 3) & 7;
              float4 hi, lo;
              split4(make_float4(xa[u].x * kLog2e, xa[u].y * kLog2e, xa[u].z * kLog2e, xa[u].w * kLog2e), hi, lo);
              a_hi[c * kTcM + rr] = hi;
              a_lo[c * kTcM + rr] = lo;
              split4(xb[u], hi, lo);
              b_hi[c * kTcM + rr] = hi;
              b_lo[c * kTcM + rr] = lo;
            }
            fence_proxy_async();
            mbar_arrive(&full_b[g]);
          }
        }
        // ---- product 2: V transposed (B: row = feature lt, k = key) -------------------------------
        for (int s = 0; s < NS2; ++s, ++sc) {
          if ((int)(sc % kTcSlots) != g) continue;
          const float* vsrc = p.V + (size_t)(lb + s * kTcKS) * kTcF + lt;
          const int keys = n - s * kTcKS;  // valid keys of this slice (>= 1)
          float4 xv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            xv[u].x = 4 * u + 0 < keys ? __ldg(vsrc + (size_t)(4 * u + 0) * kTcF) : 0.f;
            xv[u].y = 4 * u + 1 < keys ? __ldg(vsrc + (size_t)(4 * u + 1) * kTcF) : 0.f;
            xv[u].z = 4 * u + 2 < keys ? __ldg(vsrc + (size_t)(4 * u + 2) * kTcF) : 0.f;
            xv[u].w = 4 * u + 3 < keys ? __ldg(vsrc + (size_t)(4 * u + 3) * kTcF) : 0.f;
          }
          mbar_wait(&empty[g], ((sc / kTcSlots) & 1u) ^ 1u);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            float4 hi, lo;
            split4(xv[u], hi, lo);
            b_hi[u * kTcM + lt] = hi;
            b_lo[u * kTcM + lt] = lo;
          }
          fence_proxy_async();
          mbar_arrive(&full_b[g]);
        }
      }
    }
  }
  // ---- teardown ------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

static bool dense_tc_supported(int max_nodes, int h, int f) {
  return h == 1 && f == kTcF && max_nodes >= 1 && max_nodes <= kTcMaxNodes;
}

static int sm_count() {
  static const int sms = [] {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v > 0 ? v : 148;
  }();
  return sms;
}

}  // namespace dfgnn

using namespace dfgnn;

extern "C" {

int dfgnn_gt_dense_tc_supported(int max_nodes, int h, int f) { return dense_tc_supported(max_nodes, h, f) ? 1 : 0; }

int dfgnn_gt_dense_tc_forward(int n_blocks, const int32_t* blk_ptr, int max_nodes, int m, int nnz, int h, int f,
                              const int32_t* row_ptr, const int32_t* col_ind, const float* Q, const float* K,
                              const float* V, float* out_feat, float* attn_edge, void* stream) {
  const char* fn = "dfgnn_gt_dense_tc_forward";
  if (int rc = check_common(fn, m, nnz, h, f)) return rc;
  if (m == 0) return DFGNN_OK;
  DFGNN_REQUIRE(blk_ptr, fn); DFGNN_REQUIRE(row_ptr, fn);
  if (nnz > 0) DFGNN_REQUIRE(col_ind, fn);
  DFGNN_REQUIRE(Q, fn); DFGNN_REQUIRE(K, fn); DFGNN_REQUIRE(V, fn); DFGNN_REQUIRE(out_feat, fn);
  if (n_blocks < 1 || !dense_tc_supported(max_nodes, h, f)) {
    set_error("%s: needs h == 1, f == %d and graphs of at most %d nodes (h=%d, f=%d, max_nodes=%d)", fn, kTcF,
              kTcMaxNodes, h, f, max_nodes);
    return DFGNN_ERR_UNSUPPORTED_DIM;
  }
  cudaStream_t st = (cudaStream_t)stream;
  GtBlockFwdParams p{{m, nnz, h, f, 0, row_ptr, col_ind, nullptr, Q, K, V, nullptr, out_feat, nnz > 0 ? attn_edge : nullptr},
                     {blk_ptr, n_blocks, max_nodes}};
  auto kernel = gt_dense_tc_fwd_kernel;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes);
  const int grid = n_blocks < sm_count() ? n_blocks : sm_count();
  kernel<<<grid, kTcThreads, kTcSmemBytes, st>>>(p);
  note_kernel(0, "gt_dense_tc_fwd_kernel");
  return check_launch(fn);
}

}  // extern "C"
