// dense_tc.cu -- GT attention for block-diagonal batches of DENSE small graphs on the 5th-generation
// tensor cores (tcgen05, accumulators in tensor memory).
//
// A PATTERN-shaped batch (1024 graphs of ~119 nodes, 43 % of each n x n block present) is, per graph,
// a small masked dense attention: S = Q K^T, P = softmax over the row's neighbours, O = P V.  The
// per-edge kernels (fwd_kernels.cuh, block_gt.cuh) spend ~40 warp instructions and 1 KB of L1 /
// shared-memory traffic per edge there and are load/store-unit bound at 0.45 ms per forward; as two
// GEMMs per 128-row tile the kernel is bound by reading Q, K, V once and writing O and attn_edge once.
//
// fp32 parity (1e-4 relative) from TF32 tensor cores by the 3xTF32 split: x = hi + lo with hi = x
// rounded to TF32, a b ~ a_hi b_hi + a_hi b_lo + a_lo b_hi, fp32 accumulation in tensor memory.
// A tcgen05.mma kind::tf32 costs ~118 cycles for any N <= 128 and 128 cycles for N = 256 (measured,
// tools/umma_bench.cu), so the hi and lo images of the B operand are STACKED along N: one N = 256 MMA
// gives a_hi b_hi (columns 0-127) and a_hi b_lo (columns 128-255), one N = 128 MMA adds a_lo b_hi to
// the second half (the small cross terms stay together: the tensor core adds with truncation); the two
// column halves are summed when the accumulator is read.
//
// The adjacency of the batch comes as a bitmap (dfgnn_block_adj_bits: 256 bits per row, bit j = key j
// of the row's own graph), a format built once per batch like the CSC: the kernel never reads col_ind.
//
// One persistent CTA per SM walks graphs; a graph of n <= 256 nodes is one or two tiles of 128 query
// rows.  Roles (warp specialised, mbarriers only):
//   warps 0-7    softmax / epilogue, two groups of 4 warps that share every tile: group g takes the
//                32-column pieces cb with (cb & 1) == g; thread r of either group owns tile row r = TMEM
//                lane r.  Pass 1: online row max and row sum of 2^(s - max) straight from tensor memory,
//                combined across the groups through shared memory; pass 2: normalised probabilities,
//                split into the A images of the second product and (training) written to attn_edge
//                through a warp transposition (a store writes two rows' contiguous pieces); finally
//                O -> out through a staging buffer (coalesced stores).
//   warp  8      MMA issue (one lane).  Product 1: S[128 x keys] += Q_slice K_slice^T over slices of the
//                feature dimension; product 2: O[128 x 128] += P_slice V_slice over 32-key slices.
//   warps 9-16   two loader groups of 4 warps taking alternate ring stages: global -> registers (issued
//                before the slot is free) -> hi / lo split -> canonical K-major images.  Product 1 stages
//                carry Q (A) and K (B); product 2 stages carry V TRANSPOSED (B: rows = features, k = keys;
//                read as coalesced scalars, stored as one float4 of four keys) while the softmax threads
//                supply P (A).
// Ring: 3 slots of 64 KB.  Stage formats (all images "chunk major": 16-byte k-chunk c of row r at
// c * rows * 16 + r * 16):
//   narrow product 1 (n <= 128), K = 32:  A_hi | A_lo (128 rows) | B = K_hi over K_lo (256 rows)
//   wide product 1 (n > 128),   K = 16:  A_hi | A_lo (128 rows) | K_hi (256 rows) | K_lo (256 rows)
//   product 2, K = 32 keys:              P_hi | P_lo (128 rows) | B = V^T_hi over V^T_lo (256 rows)
// Tensor memory: S in columns [0, 256) (narrow tiles: two halves to be summed), O in [256, 512) (two
// halves to be summed).
//
// The backward runs on the same machinery: gt_dense_tc_body<true> (row side: dA = dO V^T, dS, dQ = dS K)
// and gt_dense_tc_bwd_col_kernel (column side: dV = P^T dO, dK = dS^T Q), further down.
//
// Requirements (checked by the block plan, formats.py / dfgnn_block_plan_check): h == 1, f == 128,
// unweighted scores, graphs of at most 256 nodes, column ids strictly ascending inside every row (no
// duplicate edges: a bitmap cannot count an edge twice).  Reference counterpart:
// fused_gtconv_hyper.cu:31-163 (forward) and, for the maths, DFGNN/layers/GT/gtconv_layer.py:28-45.
#include <algorithm>
#include <functional>
#include <queue>
#include <utility>
#include <vector>

#include "abi_common.h"
#include "tc_common.cuh"

namespace dfgnn {

constexpr int kTcM = 128;                        // rows per tile (UMMA M)
constexpr int kTcF = 128;                        // feature width
constexpr int kTcSlots = 3;
constexpr int kTcSlotBytes = 64 * 1024;
constexpr int kTcMaxNodes = 256;
constexpr int kTcSoftWarps = 8, kTcLoadGroups = 2, kTcGroupWarps = 4;
constexpr int kTcGroupThreads = kTcGroupWarps * 32;
constexpr int kTcMmaWarp = kTcSoftWarps;
constexpr int kTcThreads = (kTcSoftWarps + 1 + kTcLoadGroups * kTcGroupWarps) * 32;  // 544
constexpr int kTcMaskW = kTcMaxNodes / 32;       // bitmap words per row
constexpr int kTcStgLd = 20;                     // epilogue staging: 16 columns per row + 4 floats of padding
constexpr uint32_t kTcSBO = 128;
constexpr uint32_t kTcLboA = 128 * 16;           // chunk stride of a 128-row image
constexpr uint32_t kTcLboB = 256 * 16;           // chunk stride of a 256-row image
constexpr int kTcColO = 256;                     // first tensor-memory column of O

// slot carve (bytes)
constexpr int kTcN1_Alo = 16 * 1024, kTcN1_B = 32 * 1024;                          // K = 32 stages
constexpr int kTcW1_Alo = 8 * 1024, kTcW1_Bhi = 16 * 1024, kTcW1_Blo = 32 * 1024;  // K = 16 stages

constexpr size_t kTcOffStage = (size_t)kTcSlots * kTcSlotBytes;
constexpr size_t kTcOffPart = kTcOffStage + (size_t)2 * kTcM * kTcStgLd * 4;   // one staging buffer per softmax group
constexpr size_t kTcSmemBytes = kTcOffPart + (size_t)4 * kTcM * 4;             // partial row max / sum of the two groups

// barrier among the 128 threads of one softmax group / among all 256 softmax threads
__device__ __forceinline__ void soft_bar(int group) { asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory"); }
__device__ __forceinline__ void soft_bar_all() { asm volatile("bar.sync 3, 256;" ::: "memory"); }

__device__ __forceinline__ uint32_t sel8(const uint32_t (&a)[8], int i) {  // a[i] without local memory
  uint32_t v = a[0];
#pragma unroll
  for (int k = 1; k < 8; ++k) v = (i == k) ? a[k] : v;
  return v;
}

#ifdef DFGNN_TC_PROF
// developer build: cycles block 0 spends in / waiting for each step (tools/time_tc.py prints them);
// one thread per role: softmax tid 0, MMA tid 256, loader group 0 tid 288
__device__ unsigned long long g_tc_prof[32];
#define TC_ON() (blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 256 || threadIdx.x == 288))
#define TC_TIMED(slot, stmt) do { const long long _t0 = clock64(); stmt; if (TC_ON()) g_tc_prof[slot] += (unsigned long long)(clock64() - _t0); } while (0)
#define TC_MARK(var) const long long var = clock64()
#define TC_SINCE(slot, var) do { if (TC_ON()) g_tc_prof[slot] += (unsigned long long)(clock64() - var); } while (0)
#else
#define TC_TIMED(slot, stmt) do { stmt; } while (0)
#define TC_MARK(var)
#define TC_SINCE(slot, var)
#endif

// the tiles of this CTA, in the order every role walks them.  With a schedule (sched_ptr / sched_idx:
// the graphs of CTA c are sched_idx[sched_ptr[c] .. sched_ptr[c + 1]), balanced by the host from the
// graph sizes) a CTA walks its own list; without one, graphs blockIdx.x, blockIdx.x + gridDim.x, ...
struct TcTiles {
  const int *blk, *sidx;
  int i, i_end, step, mt, MT, lb, n, b;
  __device__ TcTiles(const int* blk_, int nb, const int* sched_ptr, const int* sched_idx) : blk(blk_), sidx(sched_idx), mt(0), MT(0), lb(0), n(0), b(0) {
    if (sched_ptr) {
      i = __ldg(sched_ptr + blockIdx.x) - 1;
      i_end = __ldg(sched_ptr + blockIdx.x + 1);
      step = 1;
    } else {
      i = (int)blockIdx.x - (int)gridDim.x;
      i_end = nb;
      step = gridDim.x;
    }
  }
  __device__ bool next() {
    if (++mt < MT) return true;
    for (i += step; i < i_end; i += step) {
      b = sidx ? __ldg(sidx + i) : i;
      lb = __ldg(blk + b);
      n = __ldg(blk + b + 1) - lb;
      if (n > 0) {
        MT = (n + kTcM - 1) / kTcM;
        mt = 0;
        return true;
      }
    }
    MT = 0;
    return false;
  }
  __device__ int stages1() const { return MT == 1 ? 4 : 8; }   // product-1 stages of a tile
  __device__ int stages2() const { return (n + 31) >> 5; }     // product-2 stages (32-key slices)
};

struct GtTcParams {
  int n_blocks;
  const int* blk_ptr;
  const int* row_ptr;
  const int *sched_ptr, *sched_idx;  // optional balanced schedule (null: round robin)
  const uint32_t* adj_bits;  // [m][8] (forward)
  const float *a1, *b1, *b2;  // forward: Q, K, V; backward row side: dO, V, K
  float *out, *attn;          // forward: out, attn_edge (may be null: inference); backward: dQ, unused
  const float* Pd;            // backward: dense probabilities, tile images (below)
  float* dSd;                 // backward: dense dS, tile images (written here, read by the column side)
  const int* tile_ptr;        // backward: [n_blocks + 1] first row tile of every graph
};
// Dense P / dS work arrays of the backward: one 128-row x 256-column image per row tile, "chunk major" like
// the operand images (4-column chunk c of row r at (c * 128 + r) * 4 floats), so that the row threads of a
// warp (thread = row) read and write 512 contiguous bytes per instruction.
constexpr int kTcTileFloats = kTcM * kTcMaxNodes;
__device__ __forceinline__ size_t tile_at(int tile, int r, int j) {
  return (size_t)tile * kTcTileFloats + (size_t)(j >> 2) * (kTcM * 4) + r * 4 + (j & 3);
}

// BWD = false: forward (S = Q K^T, softmax, O = P V).  BWD = true: row side of the backward with the same
// pipeline -- product 1 dA = dO V^T, then per row s_i = sum_j p_ij dA_ij and dS_ij = p_ij (dA_ij - s_i) with
// the probabilities read from the dense work array (block_attn_dense_kernel), product 2 dQ = dS K; dS is also
// written to its dense work array for the column side (gt_backward: DFGNN/src/fused_gtconv/fused_gtconv.cu).
template <bool BWD>
__device__ __forceinline__ void gt_dense_tc_body(const GtTcParams& p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full_b[kTcSlots], full_a[kTcSlots], empty[kTcSlots], s_full, s_free, o_full, o_free;
  __shared__ uint32_t s_tmem;
  float* s_stg = reinterpret_cast<float*>(smem + kTcOffStage);

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < kTcSlots; ++i) {
      mbar_init(&full_b[i], kTcGroupThreads);
      mbar_init(&full_a[i], 128);  // one softmax group writes a whole P slice
      mbar_init(&empty[i], 1);
    }
    mbar_init(&s_full, 1);
    mbar_init(&o_full, 1);
    mbar_init(&s_free, kTcSoftWarps * 32);
    mbar_init(&o_free, kTcSoftWarps * 32);
    mbar_fence_init();
  }
  if (w == 0) {  // all 512 columns, one CTA per SM
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  TC_MARK(t_kernel);

  if (w < kTcSoftWarps) {
    if constexpr (!BWD) {
    // =============================== softmax / epilogue ===========================================
    const int sg = w >> 2, r = tid & 127;  // group, tile row
    const uint32_t lane_base = tmem + ((uint32_t)((w & 3) * 32) << 16);
    const bool train = p.attn != nullptr;
    float* s_part = reinterpret_cast<float*>(smem + kTcOffPart);  // [max | sum][group][row]
    float* stg = s_stg + (size_t)sg * kTcM * kTcStgLd;            // this group's staging buffer
    float* tp = stg + (size_t)((w & 3) * 32) * kTcStgLd;          // this warp's 32 x 17 transposition tile
    uint32_t sc = 0, tc = 0;

    // bitmap words and first CSR position of this thread's row (and of the tile), loaded one tile ahead
    uint4 pre0 = make_uint4(0u, 0u, 0u, 0u), pre1 = pre0;
    int pre_e0 = 0, pre_t0 = 0;
    auto prefetch_row = [&](const TcTiles& t) {
      const int row = t.mt * kTcM + r;
      pre0 = pre1 = make_uint4(0u, 0u, 0u, 0u);
      pre_t0 = __ldg(p.row_ptr + t.lb + t.mt * kTcM);
      pre_e0 = pre_t0;
      if (row < t.n) {
        const uint4* src = reinterpret_cast<const uint4*>(p.adj_bits + (size_t)(t.lb + row) * kTcMaskW);
        pre0 = __ldg(src);
        pre1 = __ldg(src + 1);
        pre_e0 = __ldg(p.row_ptr + t.lb + row);
      }
    };
    TcTiles nxt(p.blk_ptr, p.n_blocks, p.sched_ptr, p.sched_idx);
    bool have = nxt.next();
    if (have) prefetch_row(nxt);
    while (have) {
      const TcTiles t = nxt;
      uint32_t mw[kTcMaskW] = {pre0.x, pre0.y, pre0.z, pre0.w, pre1.x, pre1.y, pre1.z, pre1.w};
      const int tile_e0 = pre_t0;
      int e_rel = pre_e0 - pre_t0;  // CSR position of the row's next entry, relative to the tile's first (< 2^15)
      have = nxt.next();
      if (have) prefetch_row(nxt);

      const bool narrow = t.MT == 1;
      const int NS2 = t.stages2();
      const uint32_t sc2 = sc + t.stages1();  // stage index of the first product-2 stage of this tile
      auto load_s = [&](int cb, float (&s)[32]) {  // 32 columns of S (narrow tiles: the sum of the two halves)
        tmem_ld32(lane_base + cb * 32, s);
        if (narrow) {
          float s2[32];
          tmem_ld32(lane_base + 128 + cb * 32, s2);
#pragma unroll
          for (int i = 0; i < 32; ++i) s[i] += s2[i];
        }
      };
      // ---- S is complete ----------------------------------------------------------------------
      TC_TIMED(1, mbar_wait(&s_full, tc & 1u));
      tc_fence_after();
      TC_MARK(t_p1);
      // pass 1: row max and row sum of 2^(s - max) over this group's pieces (online), then combined
      float mx = kNeg, l = 0.f;
#pragma unroll 1
      for (int cb = sg; cb < NS2; cb += 2) {
        float s[32];
        load_s(cb, s);
        const uint32_t m = sel8(mw, cb);
        float cm = kNeg;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          s[i] = ((m >> i) & 1u) ? s[i] : kNeg;
          cm = fmaxf(cm, s[i]);
        }
        if (cm > mx) {
          l *= fast_exp2(mx - cm);
          mx = cm;
        }
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {  // masked entries: 2^(-1e30 - max) = 0 (a row without entries keeps l = 0:
          a0 += ((m >> i) & 1u) ? fast_exp2(s[i] - mx) : 0.f;          // its mask words are 0)
          a1 += ((m >> (i + 1)) & 1u) ? fast_exp2(s[i + 1] - mx) : 0.f;
        }
        l += a0 + a1;
      }
      s_part[sg * kTcM + r] = mx;
      s_part[(2 + sg) * kTcM + r] = l;
      soft_bar_all();
      {
        const float mo = s_part[(sg ^ 1) * kTcM + r], lo = s_part[(2 + (sg ^ 1)) * kTcM + r];
        const float mn = fmaxf(mx, mo);
        l = l * fast_exp2(mx - mn) + lo * fast_exp2(mo - mn);
        mx = mn;
      }
      const float inv = l > 0.f ? 1.f / l : 0.f;
      TC_SINCE(4, t_p1);
      TC_MARK(t_p2);
      // pass 2: normalised probabilities -> A images of product 2 and (training) attn_edge.
      // attn_edge in CSR order: a row's neighbours are its set bitmap bits, ascending.  The 32-row x
      // 16-column piece a warp holds (thread = row) is transposed through its slice of the staging buffer
      // so that one store instruction writes the pieces of TWO rows (lanes 0-15 row rr, lanes 16-31 row
      // rr + 16): contiguous floats instead of 32 different rows.
      const int hl = lane & 15, hr = lane >> 4;
      const uint32_t below = (1u << hl) - 1u;
#pragma unroll 1
      for (int cb = 0; cb < NS2; ++cb) {
        const uint32_t m = sel8(mw, cb);
        if ((cb & 1) != sg) {  // the other group's piece (warp-uniform branch)
          e_rel += __popc(m);
          continue;
        }
        float s[32];
        load_s(cb, s);
#pragma unroll
        for (int i = 0; i < 32; ++i) s[i] = ((m >> i) & 1u) ? fast_exp2(s[i] - mx) * inv : 0.f;
        const uint32_t st = sc2 + cb, slot = st % kTcSlots, k = st / kTcSlots;
        TC_TIMED(2, mbar_wait(&empty[slot], (k & 1u) ^ 1u));
        float4* a_hi = reinterpret_cast<float4*>(smem + (size_t)slot * kTcSlotBytes);
        float4* a_lo = reinterpret_cast<float4*>(smem + (size_t)slot * kTcSlotBytes + kTcN1_Alo);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 hi, lo;
          split4(make_float4(s[4 * c], s[4 * c + 1], s[4 * c + 2], s[4 * c + 3]), hi, lo);
          a_hi[c * kTcM + r] = hi;
          a_lo[c * kTcM + r] = lo;
        }
        fence_proxy_async();
        mbar_arrive(&full_a[slot]);
        if (train) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int i = 0; i < 16; ++i) tp[lane * 17 + i] = s[16 * half + i];
            __syncwarp();
            const uint32_t mh = (m >> (16 * half)) & 0xffffu;
            const uint32_t packed = ((uint32_t)e_rel << 16) | mh;
#pragma unroll
            for (int rr = 0; rr < 16; ++rr) {
              const int src = rr + 16 * hr;
              const uint32_t v = __shfl_sync(kFull, packed, src);
              if ((v >> hl) & 1u) p.attn[tile_e0 + (int)(v >> 16) + __popc(v & below)] = tp[src * 17 + hl];
            }
            __syncwarp();
            e_rel += __popc(mh);
          }
        }
      }
      TC_SINCE(5, t_p2);
      tc_fence_before();
      mbar_arrive(&s_free);
      // ---- O is complete ----------------------------------------------------------------------
      TC_TIMED(3, mbar_wait(&o_full, tc & 1u));
      tc_fence_after();
      TC_MARK(t_ep);
      float* obase = p.out + (size_t)(t.lb + t.mt * kTcM) * kTcF;
      const int rows_here = min(kTcM, t.n - t.mt * kTcM);
#pragma unroll 1
      for (int cq = sg; cq < kTcF / 32; cq += 2) {
        float y[32];
        {
          float y2[32];
          tmem_ld32(lane_base + kTcColO + cq * 32, y);
          tmem_ld32(lane_base + kTcColO + 128 + cq * 32, y2);
#pragma unroll
          for (int i = 0; i < 32; ++i) y[i] += y2[i];
        }
        if (cq + 2 >= kTcF / 32) {  // this thread's last read of O
          tc_fence_before();
          mbar_arrive(&o_free);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float4* so = reinterpret_cast<float4*>(stg + (size_t)r * kTcStgLd);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            so[i] = make_float4(y[16 * half + 4 * i], y[16 * half + 4 * i + 1], y[16 * half + 4 * i + 2],
                                y[16 * half + 4 * i + 3]);
          soft_bar(sg);
#pragma unroll
          for (int u = 0; u < 4; ++u) {  // a warp stores 8 rows x 64 contiguous bytes
            const int i = r + u * 128, rr = i >> 2, c4 = i & 3;
            if (rr < rows_here)
              *reinterpret_cast<float4*>(obase + (size_t)rr * kTcF + cq * 32 + 16 * half + 4 * c4) =
                  *reinterpret_cast<const float4*>(stg + (size_t)rr * kTcStgLd + 4 * c4);
          }
          soft_bar(sg);
        }
      }
      TC_SINCE(7, t_ep);
      sc = sc2 + NS2;
      ++tc;
    }
    } else {
    // =============================== backward row side: dS / epilogue ==============================
    const int sg = w >> 2, r = tid & 127;  // group, tile row
    const uint32_t lane_base = tmem + ((uint32_t)((w & 3) * 32) << 16);
    float* s_part = reinterpret_cast<float*>(smem + kTcOffPart);  // [group][row] partial row sums
    float* stg = s_stg + (size_t)sg * kTcM * kTcStgLd;            // this group's staging buffer
    uint32_t sc = 0, tc = 0;
    TcTiles t(p.blk_ptr, p.n_blocks, p.sched_ptr, p.sched_idx);
    while (t.next()) {
      const bool narrow = t.MT == 1;
      const int NS2 = t.stages2();
      const uint32_t sc2 = sc + t.stages1();
      const int row = t.mt * kTcM + r;
      const bool valid = row < t.n;
      const int tile = __ldg(p.tile_ptr + t.b) + t.mt;
      // float4 (chunk c, row r) of this tile's images: c * 128 + r
      const float4* prow = reinterpret_cast<const float4*>(p.Pd + (size_t)tile * kTcTileFloats) + r;
      float4* dsrow = reinterpret_cast<float4*>(p.dSd + (size_t)tile * kTcTileFloats) + r;
      auto load_da = [&](int cb, float (&s)[32]) {  // 32 columns of dA (narrow tiles: the sum of the two halves)
        tmem_ld32(lane_base + cb * 32, s);
        if (narrow) {
          float s2[32];
          tmem_ld32(lane_base + 128 + cb * 32, s2);
#pragma unroll
          for (int i = 0; i < 32; ++i) s[i] += s2[i];
        }
      };
      mbar_wait(&s_full, tc & 1u);
      tc_fence_after();
      // pass 1: s_i = sum_j p_ij dA_ij over this group's pieces (p = 0 where there is no edge)
      float srow = 0.f;
#pragma unroll 1
      for (int cb = sg; cb < NS2; cb += 2) {
        float4 pv[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) pv[c] = valid ? __ldg(prow + (cb * 8 + c) * kTcM) : make_float4(0.f, 0.f, 0.f, 0.f);
        float s[32];
        load_da(cb, s);
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          a0 = fmaf(pv[c].x, s[4 * c], a0);
          a1 = fmaf(pv[c].y, s[4 * c + 1], a1);
          a0 = fmaf(pv[c].z, s[4 * c + 2], a0);
          a1 = fmaf(pv[c].w, s[4 * c + 3], a1);
        }
        srow += a0 + a1;
      }
      s_part[sg * kTcM + r] = srow;
      soft_bar_all();
      srow += s_part[(sg ^ 1) * kTcM + r];
      // pass 2: dS -> A images of product 2 and the dense work array
#pragma unroll 1
      for (int cb = sg; cb < NS2; cb += 2) {
        float4 pv[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) pv[c] = valid ? __ldg(prow + (cb * 8 + c) * kTcM) : make_float4(0.f, 0.f, 0.f, 0.f);
        float s[32];
        load_da(cb, s);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          pv[c].x *= s[4 * c] - srow;
          pv[c].y *= s[4 * c + 1] - srow;
          pv[c].z *= s[4 * c + 2] - srow;
          pv[c].w *= s[4 * c + 3] - srow;
        }
        const uint32_t st = sc2 + cb, slot = st % kTcSlots, k = st / kTcSlots;
        mbar_wait(&empty[slot], (k & 1u) ^ 1u);
        float4* a_hi = reinterpret_cast<float4*>(smem + (size_t)slot * kTcSlotBytes);
        float4* a_lo = reinterpret_cast<float4*>(smem + (size_t)slot * kTcSlotBytes + kTcN1_Alo);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 hi, lo;
          split4(pv[c], hi, lo);
          a_hi[c * kTcM + r] = hi;
          a_lo[c * kTcM + r] = lo;
        }
        fence_proxy_async();
        mbar_arrive(&full_a[slot]);
        if (valid) {
#pragma unroll
          for (int c = 0; c < 8; ++c) dsrow[(cb * 8 + c) * kTcM] = pv[c];
        }
      }
      tc_fence_before();
      mbar_arrive(&s_free);
      // ---- dQ is complete -----------------------------------------------------------------------
      mbar_wait(&o_full, tc & 1u);
      tc_fence_after();
      float* obase = p.out + (size_t)(t.lb + t.mt * kTcM) * kTcF;
      const int rows_here = min(kTcM, t.n - t.mt * kTcM);
#pragma unroll 1
      for (int cq = sg; cq < kTcF / 32; cq += 2) {
        float y[32];
        {
          float y2[32];
          tmem_ld32(lane_base + kTcColO + cq * 32, y);
          tmem_ld32(lane_base + kTcColO + 128 + cq * 32, y2);
#pragma unroll
          for (int i = 0; i < 32; ++i) y[i] += y2[i];
        }
        if (cq + 2 >= kTcF / 32) {  // this thread's last read of the accumulator
          tc_fence_before();
          mbar_arrive(&o_free);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float4* so = reinterpret_cast<float4*>(stg + (size_t)r * kTcStgLd);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            so[i] = make_float4(y[16 * half + 4 * i], y[16 * half + 4 * i + 1], y[16 * half + 4 * i + 2],
                                y[16 * half + 4 * i + 3]);
          soft_bar(sg);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = r + u * 128, rr = i >> 2, c4 = i & 3;
            if (rr < rows_here)
              *reinterpret_cast<float4*>(obase + (size_t)rr * kTcF + cq * 32 + 16 * half + 4 * c4) =
                  *reinterpret_cast<const float4*>(stg + (size_t)rr * kTcStgLd + 4 * c4);
          }
          soft_bar(sg);
        }
      }
      sc = sc2 + NS2;
      ++tc;
    }
    }
  } else if (w == kTcMmaWarp) {
    // =============================== MMA issue ====================================================
    if (lane == 0) {
      uint32_t sc = 0, tc = 0, ka = 0;  // ka: bit s = parity of full_a[s]
      constexpr uint32_t id256 = umma_idesc_tf32(kTcM, 256), id128 = umma_idesc_tf32(kTcM, 128);
      const uint32_t ring = smem_u32(smem);
      TcTiles t(p.blk_ptr, p.n_blocks, p.sched_ptr, p.sched_idx);
      while (t.next()) {
        TC_TIMED(8, mbar_wait(&s_free, (tc & 1u) ^ 1u));  // the softmax threads have read S of the previous tile
        tc_fence_after();
        if (t.MT == 1) {
          // narrow tile: per K = 8, one N = 256 MMA (Q_hi x [K_hi over K_lo]) and one N = 128 MMA (Q_lo x K_hi)
          for (int q = 0; q < 4; ++q, ++sc) {
            const uint32_t slot = sc % kTcSlots, k = sc / kTcSlots;
            TC_TIMED(9, mbar_wait(&full_b[slot], k & 1u));
            tc_fence_after();
            const uint32_t base = ring + slot * kTcSlotBytes;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t dah = umma_desc_kmajor(base + ks * 2 * kTcLboA, kTcLboA, kTcSBO);
              const uint64_t dal = umma_desc_kmajor(base + kTcN1_Alo + ks * 2 * kTcLboA, kTcLboA, kTcSBO);
              const uint64_t db = umma_desc_kmajor(base + kTcN1_B + ks * 2 * kTcLboB, kTcLboB, kTcSBO);
              // the cross terms (a_lo b_hi here, a_hi b_lo from the N = 256 MMA) stay in the second, small-magnitude
              // half: the tensor core adds with truncation, and every add into the full-magnitude half costs
              // up to one ulp of it (measured: alternating the halves made the results worse, tools/tc_error.py)
              umma_tf32(tmem, dah, db, id256, (q | ks) != 0 ? 1u : 0u);
              umma_tf32(tmem + 128, dal, db, id128, 1u);
            }
            umma_commit(&empty[slot]);
          }
        } else {
          // wide tile: up to 256 keys in one MMA, 3 MMAs per K = 8, K = 16 per stage
          // N = every column the softmax threads read (whole 32-column pieces): the K / V rows behind the graph's
          // last key are zero in the images, so those columns are 0, never stale tensor memory (the backward
          // multiplies them by p = 0, which would not survive a NaN)
          const uint32_t idk = umma_idesc_tf32(kTcM, ((t.n + 31) >> 5) << 5);
          for (int q = 0; q < 8; ++q, ++sc) {
            const uint32_t slot = sc % kTcSlots, k = sc / kTcSlots;
            TC_TIMED(9, mbar_wait(&full_b[slot], k & 1u));
            tc_fence_after();
            const uint32_t base = ring + slot * kTcSlotBytes;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint64_t dah = umma_desc_kmajor(base + ks * 2 * kTcLboA, kTcLboA, kTcSBO);
              const uint64_t dal = umma_desc_kmajor(base + kTcW1_Alo + ks * 2 * kTcLboA, kTcLboA, kTcSBO);
              const uint64_t dbh = umma_desc_kmajor(base + kTcW1_Bhi + ks * 2 * kTcLboB, kTcLboB, kTcSBO);
              const uint64_t dbl = umma_desc_kmajor(base + kTcW1_Blo + ks * 2 * kTcLboB, kTcLboB, kTcSBO);
              umma_tf32(tmem, dal, dbh, idk, (q | ks) != 0 ? 1u : 0u);
              umma_tf32(tmem, dah, dbl, idk, 1u);
              umma_tf32(tmem, dah, dbh, idk, 1u);
            }
            umma_commit(&empty[slot]);
          }
        }
        umma_commit(&s_full);
        TC_TIMED(10, mbar_wait(&o_free, (tc & 1u) ^ 1u));  // O of the previous tile has been read
        tc_fence_after();
        const int NS2 = t.stages2();
        for (int s = 0; s < NS2; ++s, ++sc) {
          const uint32_t slot = sc % kTcSlots, k = sc / kTcSlots;
          TC_TIMED(11, mbar_wait(&full_b[slot], k & 1u));
          TC_TIMED(12, mbar_wait(&full_a[slot], (ka >> slot) & 1u));
          ka ^= 1u << slot;
          tc_fence_after();
          const uint32_t base = ring + slot * kTcSlotBytes;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t dah = umma_desc_kmajor(base + ks * 2 * kTcLboA, kTcLboA, kTcSBO);
            const uint64_t dal = umma_desc_kmajor(base + kTcN1_Alo + ks * 2 * kTcLboA, kTcLboA, kTcSBO);
            const uint64_t db = umma_desc_kmajor(base + kTcN1_B + ks * 2 * kTcLboB, kTcLboB, kTcSBO);
            umma_tf32(tmem + kTcColO, dah, db, id256, (s | ks) != 0 ? 1u : 0u);
            umma_tf32(tmem + kTcColO + 128, dal, db, id128, 1u);
          }
          umma_commit(&empty[slot]);
        }
        umma_commit(&o_full);
        ++tc;
      }
    }
  } else {
    // =============================== loaders ======================================================
    const int g = (w - kTcMmaWarp - 1) / kTcGroupWarps;
    const int lt = tid - (kTcMmaWarp + 1 + g * kTcGroupWarps) * 32;  // thread inside the group
    uint32_t sc = 0;
    TcTiles t(p.blk_ptr, p.n_blocks, p.sched_ptr, p.sched_idx);
    while (t.next()) {
      const int rows_a = min(kTcM, t.n - t.mt * kTcM);
      const float* qrow = p.a1 + (size_t)(t.lb + t.mt * kTcM) * kTcF;
      const float* krow = p.b1 + (size_t)t.lb * kTcF;
      constexpr float kScaleA = BWD ? 1.f : kLog2e;  // forward: Q into the base-2 exponent domain
      if (t.MT == 1) {
        // ---- narrow product 1, K = 32: Q slice (A, scaled into the base-2 exponent domain), K slice (B) ----
        for (int q = 0; q < 4; ++q, ++sc) {
          if ((int)(sc % kTcLoadGroups) != g) continue;
          unsigned char* slot = smem + (size_t)(sc % kTcSlots) * kTcSlotBytes;
          float4* const a_hi = reinterpret_cast<float4*>(slot);
          float4* const a_lo = reinterpret_cast<float4*>(slot + kTcN1_Alo);
          float4* const bst = reinterpret_cast<float4*>(slot + kTcN1_B);
          const float4* qsrc = reinterpret_cast<const float4*>(qrow + q * 32);
          const float4* ksrc = reinterpret_cast<const float4*>(krow + q * 32);
          TC_MARK(t_ld1);
          float4 xa[8], xb[8];
          // float4 i covers row (i & 7) + 8 * (i >> 6), chunk (i >> 3) & 7: a warp reads 8 rows x 64
          // contiguous bytes and writes 4 x 128 contiguous bytes of the chunk-major images
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int i = lt + u * kTcGroupThreads, rr = (i & 7) + 8 * (i >> 6), c = (i >> 3) & 7;
            xa[u] = rr < rows_a ? __ldg(qsrc + (size_t)rr * (kTcF / 4) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            xb[u] = rr < t.n ? __ldg(ksrc + (size_t)rr * (kTcF / 4) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          TC_TIMED(16, mbar_wait(&empty[sc % kTcSlots], ((sc / kTcSlots) & 1u) ^ 1u));
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int i = lt + u * kTcGroupThreads, rr = (i & 7) + 8 * (i >> 6), c = (i >> 3) & 7;
            float4 hi, lo;
            split4(make_float4(xa[u].x * kScaleA, xa[u].y * kScaleA, xa[u].z * kScaleA, xa[u].w * kScaleA), hi, lo);
            a_hi[c * 128 + rr] = hi;
            a_lo[c * 128 + rr] = lo;
            split4(xb[u], hi, lo);
            bst[c * 256 + rr] = hi;
            bst[c * 256 + 128 + rr] = lo;
          }
          fence_proxy_async();
          mbar_arrive(&full_b[sc % kTcSlots]);
          TC_SINCE(18, t_ld1);
        }
      } else {
        // ---- wide product 1, K = 16: Q slice (128 rows), K slice (256 rows) ----------------------------
        for (int q = 0; q < 8; ++q, ++sc) {
          if ((int)(sc % kTcLoadGroups) != g) continue;
          unsigned char* slot = smem + (size_t)(sc % kTcSlots) * kTcSlotBytes;
          float4* const a_hi = reinterpret_cast<float4*>(slot);
          float4* const a_lo = reinterpret_cast<float4*>(slot + kTcW1_Alo);
          float4* const b_hi = reinterpret_cast<float4*>(slot + kTcW1_Bhi);
          float4* const b_lo = reinterpret_cast<float4*>(slot + kTcW1_Blo);
          const float4* qsrc = reinterpret_cast<const float4*>(qrow + q * 16);
          const float4* ksrc = reinterpret_cast<const float4*>(krow + q * 16);
          float4 xa[4], xb[8];
          // float4 i covers row (i & 7) + 8 * (i >> 5), chunk (i >> 3) & 3
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int i = lt + u * kTcGroupThreads, rr = (i & 7) + 8 * (i >> 5), c = (i >> 3) & 3;
            if (u < 4) xa[u] = rr < rows_a ? __ldg(qsrc + (size_t)rr * (kTcF / 4) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            xb[u] = rr < t.n ? __ldg(ksrc + (size_t)rr * (kTcF / 4) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          TC_TIMED(16, mbar_wait(&empty[sc % kTcSlots], ((sc / kTcSlots) & 1u) ^ 1u));
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int i = lt + u * kTcGroupThreads, rr = (i & 7) + 8 * (i >> 5), c = (i >> 3) & 3;
            float4 hi, lo;
            if (u < 4) {
              split4(make_float4(xa[u].x * kScaleA, xa[u].y * kScaleA, xa[u].z * kScaleA, xa[u].w * kScaleA), hi, lo);
              a_hi[c * 128 + rr] = hi;
              a_lo[c * 128 + rr] = lo;
            }
            split4(xb[u], hi, lo);
            b_hi[c * 256 + rr] = hi;
            b_lo[c * 256 + rr] = lo;
          }
          fence_proxy_async();
          mbar_arrive(&full_b[sc % kTcSlots]);
        }
      }
      // ---- product 2: V transposed (B: row = feature lt, k = key), hi rows 0-127 over lo rows 128-255 ----
      const int NS2 = t.stages2();
      for (int s = 0; s < NS2; ++s, ++sc) {
        if ((int)(sc % kTcLoadGroups) != g) continue;
        unsigned char* slot = smem + (size_t)(sc % kTcSlots) * kTcSlotBytes;
        float4* const bst = reinterpret_cast<float4*>(slot + kTcN1_B);
        const float* vsrc = p.b2 + (size_t)(t.lb + s * 32) * kTcF + lt;
        const int keys = t.n - s * 32;  // valid keys of this slice (>= 1)
        TC_MARK(t_ld2);
        float4 xv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          xv[u].x = 4 * u + 0 < keys ? __ldg(vsrc + (size_t)(4 * u + 0) * kTcF) : 0.f;
          xv[u].y = 4 * u + 1 < keys ? __ldg(vsrc + (size_t)(4 * u + 1) * kTcF) : 0.f;
          xv[u].z = 4 * u + 2 < keys ? __ldg(vsrc + (size_t)(4 * u + 2) * kTcF) : 0.f;
          xv[u].w = 4 * u + 3 < keys ? __ldg(vsrc + (size_t)(4 * u + 3) * kTcF) : 0.f;
        }
        TC_TIMED(17, mbar_wait(&empty[sc % kTcSlots], ((sc / kTcSlots) & 1u) ^ 1u));
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float4 hi, lo;
          split4(xv[u], hi, lo);
          bst[u * 256 + lt] = hi;
          bst[u * 256 + 128 + lt] = lo;
        }
        fence_proxy_async();
        mbar_arrive(&full_b[sc % kTcSlots]);
        TC_SINCE(19, t_ld2);
      }
    }
  }
#ifdef DFGNN_TC_PROF
  if (TC_ON()) g_tc_prof[threadIdx.x == 288 ? 27 : 24 + (threadIdx.x >> 7)] += (unsigned long long)(clock64() - t_kernel);
#endif
  // ---- teardown ------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

__global__ void __launch_bounds__(kTcThreads, 1) gt_dense_tc_fwd_kernel(const GtTcParams p) { gt_dense_tc_body<false>(p); }
__global__ void __launch_bounds__(kTcThreads, 1) gt_dense_tc_bwd_row_kernel(const GtTcParams p) { gt_dense_tc_body<true>(p); }

// adjacency bitmap of a block-diagonal batch: bit j of row r's 256 bits = (r, first node of r's graph + j)
// is an edge.  One CTA per graph, a warp per row.
static __global__ void block_adj_bits_kernel(int n_blocks, const int* __restrict__ blk_ptr, const int* __restrict__ row_ptr,
                                             const int* __restrict__ col_ind, uint32_t* __restrict__ bits) {
  const int b = blockIdx.x, lb = blk_ptr[b], n = blk_ptr[b + 1] - lb;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int r = w; r < n; r += nw) {
    const int e0 = row_ptr[lb + r], e1 = row_ptr[lb + r + 1];
    uint32_t mine = 0u;  // lane k < 8 ends up with word k
    for (int e = e0 + lane; __any_sync(kFull, e < e1); e += 32) {
      const int j = e < e1 ? col_ind[e] - lb : -1;
#pragma unroll
      for (int k = 0; k < kTcMaskW; ++k) {
        const uint32_t word = __reduce_or_sync(kFull, (j >= 0 && (j >> 5) == k) ? (1u << (j & 31)) : 0u);
        if (lane == k) mine |= word;
      }
    }
    if (lane < kTcMaskW) bits[(size_t)(lb + r) * kTcMaskW + lane] = mine;
  }
}

// ======================================================================================================
// Backward, column side: dV_j = sum_i p_ij dO_i, dK_j = sum_i dS_ij Q_i from the row side's packed scratch
// {dS_e, p_e} (CSR order; gt_bwd_row_kernel or the dense row-side kernel writes it).  Reference counterpart:
// the SpMM-over-CSC half of fused_gtconv_backward (DFGNN/src/fused_gtconv/fused_gtconv.cu, gt_backward).
//
// Work item = (graph, 128-key tile kt): dV[keys x 128] = P^T dO and dK = dS^T Q, contraction over the rows of
// the graph in 16-row slices (one ring stage each).  M = keys, so the A operands are the TRANSPOSED tiles
// P^T / dS^T: element (key j, row i) at chunk i / 4, image row j, word i % 4 -- written straight from the
// CSR scratch through the adjacency bitmap (a warp covers 8 keys x 4 rows = 128 contiguous bytes of an image,
// every element of the image is written, zeros where there is no edge).  B operands: dO^T / Q^T slices
// (rows = features), hi over lo stacked along N like in the forward.
//   warps 0-7   workers, two groups of 4 warps taking alternate stages: per item a table of the tile's bitmap
//               words and CSR positions per row (CSR input), then per stage 16 loads per thread -> hi / lo -> the four A images; after the item's last stage
//               the epilogue (group 0: dV, group 1: dK; two column halves summed, staged, coalesced stores)
//   warp  8     MMA issue: per K = 8 and product one N = 256 MMA (A_hi x [B_hi over B_lo]) + one N = 128 MMA
//               (A_lo x B_hi); dV in tensor-memory columns [0, 256), dK in [256, 512)
//   warps 9-16  loaders of the B images, two groups taking alternate stages (thread = feature, coalesced
//               scalar loads of 16 rows)
constexpr int kBcWorkWarps = 8, kBcLoadWarps = 8;  // two groups of 4 warps each, taking alternate ring stages
constexpr int kBcThreads = (kBcWorkWarps + 1 + kBcLoadWarps) * 32;  // 544
constexpr int kBcA = 8 * 1024;                       // one A image: 4 chunks x 128 rows x 16 B
constexpr int kBcOffB0 = 4 * kBcA, kBcOffB1 = 4 * kBcA + 16 * 1024;  // dO^T / Q^T images (256 rows, hi over lo)
constexpr size_t kBcOffTab = (size_t)kTcSlots * kTcSlotBytes;          // [256 rows][uint4 words | int4 positions]
constexpr size_t kBcOffStage = kBcOffTab + (size_t)kTcMaxNodes * 32;
constexpr size_t kBcSmemBytes = kBcOffStage + (size_t)2 * kTcM * kTcStgLd * 4;

struct TcItems {  // (graph, key tile) items of this CTA; item id = 2 * graph + kt
  const int *blk, *sidx;
  int i, i_end, step, lb, n, kt, b;
  __device__ TcItems(const int* blk_, int nb, const int* sched_ptr, const int* sched_idx) : blk(blk_), sidx(sched_idx), lb(0), n(0), kt(0), b(0) {
    if (sched_ptr) {
      i = __ldg(sched_ptr + blockIdx.x) - 1;
      i_end = __ldg(sched_ptr + blockIdx.x + 1);
      step = 1;
    } else {
      i = (int)blockIdx.x - (int)gridDim.x;
      i_end = 2 * nb;
      step = gridDim.x;
    }
  }
  __device__ bool next() {
    for (i += step; i < i_end; i += step) {
      const int item = sidx ? __ldg(sidx + i) : i;
      b = item >> 1;
      kt = item & 1;
      lb = __ldg(blk + b);
      n = __ldg(blk + b + 1) - lb;
      if (n > kt * kTcM) return true;
    }
    return false;
  }
  __device__ int stages() const { return (n + 15) >> 4; }
};

struct GtTcBwdColParams {
  int n_blocks;
  const int *blk_ptr, *row_ptr, *sched_ptr, *sched_idx;
  const uint32_t* adj_bits;
  const float2* scratch;  // [nnz] {dS_e, p_e} (DENSE = false)
  const float *Pd, *dSd;  // dense probabilities and dS, tile images (DENSE = true)
  const int* tile_ptr;    // [n_blocks + 1] first row tile of every graph (DENSE = true)
  const float *dO, *Q;
  float *dK, *dV;
};

// DENSE = false: P / dS from the packed CSR scratch of the general row-side kernel, through the bitmap.
// DENSE = true: from the dense work arrays the tcgen05 row-side kernel leaves (no bitmap, no CSR).
template <bool DENSE>
__global__ void __launch_bounds__(kBcThreads, 1) gt_dense_tc_bwd_col_kernel(const GtTcBwdColParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full_a[kTcSlots], full_b[kTcSlots], empty[kTcSlots], acc_full, acc_free;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < kTcSlots; ++i) {
      mbar_init(&full_a[i], 128);  // one group of four worker warps
      mbar_init(&full_b[i], 128);  // one group of four loader warps
      mbar_init(&empty[i], 1);
    }
    mbar_init(&acc_full, 1);
    mbar_init(&acc_free, kBcWorkWarps * 32);
    mbar_fence_init();
  }
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  if (w < kBcWorkWarps) {
    // =============================== workers: A images, then the epilogue =========================
    uint4* s_tabw = reinterpret_cast<uint4*>(smem + kBcOffTab);           // bitmap words of the key tile, per row
    int4* s_tabc = reinterpret_cast<int4*>(smem + kBcOffTab) + kTcMaxNodes;  // CSR position of the first edge of each word
    const int eg = w >> 2;                                                 // epilogue group: 0 dV, 1 dK
    float* stg = reinterpret_cast<float*>(smem + kBcOffStage) + (size_t)eg * kTcM * kTcStgLd;
    const uint32_t lane_base = tmem + ((uint32_t)((w & 3) * 32) << 16) + eg * 256;
    const int i_loc = 4 * (w & 3) + (lane & 3);  // this thread's row inside a 16-row slice
    const int j0 = lane >> 2;                    // its key in iteration u: j0 + 8 u
    const uint32_t wg = w >> 2;                  // worker group: takes the stages with (stage & 1) == wg
    uint32_t sc = 0, tc = 0;
    TcItems t(p.blk_ptr, p.n_blocks, p.sched_ptr, p.sched_idx);
    while (t.next()) {
      if constexpr (!DENSE) {
        asm volatile("bar.sync 1, 256;" ::: "memory");  // the previous item's stages have read the tables
        {
          const int row = tid;  // one table row per thread
          uint4 ww = make_uint4(0u, 0u, 0u, 0u);
          int4 cc = make_int4(0, 0, 0, 0);
          if (row < t.n) {
            const uint4* src = reinterpret_cast<const uint4*>(p.adj_bits + (size_t)(t.lb + row) * kTcMaskW);
            int e = __ldg(p.row_ptr + t.lb + row);
            if (t.kt) {
              const uint4 lo = __ldg(src);
              e += __popc(lo.x) + __popc(lo.y) + __popc(lo.z) + __popc(lo.w);
            }
            ww = __ldg(src + t.kt);
            cc.x = e;
            cc.y = cc.x + __popc(ww.x);
            cc.z = cc.y + __popc(ww.y);
            cc.w = cc.z + __popc(ww.z);
          }
          s_tabw[row] = ww;
          s_tabc[row] = cc;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      const int NS = t.stages();
      const int tile0 = DENSE ? __ldg(p.tile_ptr + t.b) : 0;
      for (int s = 0; s < NS; ++s, ++sc) {
        if ((sc & 1u) != wg) continue;
        const uint32_t slot = sc % kTcSlots;
        float2 v[16];
        if constexpr (DENSE) {
          // element (row, key jg = 128 kt + j0 + 8 u) of a tile image: (jg >> 2) * 512 + r * 4 + (jg & 3) floats,
          // i.e. a per-thread base plus u * 1024
          const int row = 16 * s + i_loc, n32 = (t.n + 31) & ~31;
          const size_t base = (size_t)(tile0 + (row >> 7)) * kTcTileFloats +
                              (size_t)(t.kt * 32 + (j0 >> 2)) * (kTcM * 4) + (row & 127) * 4 + (j0 & 3);
          const float* ps = p.dSd + base;
          const float* pp = p.Pd + base;
          const int jmax = row < t.n ? n32 - t.kt * kTcM - j0 : 0;  // keys j0 + 8 u < jmax are valid
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            v[u] = make_float2(0.f, 0.f);
            if (8 * u < jmax) v[u] = make_float2(__ldg(ps + u * (2 * kTcM * 4)), __ldg(pp + u * (2 * kTcM * 4)));
          }
        } else {
          const uint4 ww = s_tabw[16 * s + i_loc];
          const int4 cc = s_tabc[16 * s + i_loc];
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int bit = j0 + 8 * (u & 3);  // key j0 + 8 u lies in word u >> 2
            const uint32_t word = (u >> 2) == 0 ? ww.x : (u >> 2) == 1 ? ww.y : (u >> 2) == 2 ? ww.z : ww.w;
            const int cbase = (u >> 2) == 0 ? cc.x : (u >> 2) == 1 ? cc.y : (u >> 2) == 2 ? cc.z : cc.w;
            v[u] = make_float2(0.f, 0.f);
            if ((word >> bit) & 1u) v[u] = __ldg(p.scratch + cbase + __popc(word & ((1u << bit) - 1u)));
          }
        }
        mbar_wait(&empty[slot], ((sc / kTcSlots) & 1u) ^ 1u);
        float* img = reinterpret_cast<float*>(smem + (size_t)slot * kTcSlotBytes);
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int j = j0 + 8 * u;
          const int at = ((w & 3) * kTcM + j) * 4 + (lane & 3);  // chunk (w & 3), image row j, word i % 4
          const float ph = round_tf32(v[u].y), dh = round_tf32(v[u].x);
          img[at] = ph;
          img[kBcA / 4 + at] = v[u].y - ph;
          img[2 * kBcA / 4 + at] = dh;
          img[3 * kBcA / 4 + at] = v[u].x - dh;
        }
        fence_proxy_async();
        mbar_arrive(&full_a[slot]);
      }
      // ---- epilogue: the accumulators of this item ----------------------------------------------
      mbar_wait(&acc_full, tc & 1u);
      tc_fence_after();
      float* obase = (eg == 0 ? p.dV : p.dK) + (size_t)(t.lb + t.kt * kTcM) * kTcF;
      const int rows_here = min(kTcM, t.n - t.kt * kTcM), r = tid & 127;
#pragma unroll 1
      for (int cq = 0; cq < kTcF / 32; ++cq) {
        float y[32];
        {
          float y2[32];
          tmem_ld32(lane_base + cq * 32, y);
          tmem_ld32(lane_base + 128 + cq * 32, y2);
#pragma unroll
          for (int i = 0; i < 32; ++i) y[i] += y2[i];
        }
        if (cq == kTcF / 32 - 1) {
          tc_fence_before();
          mbar_arrive(&acc_free);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float4* so = reinterpret_cast<float4*>(stg + (size_t)r * kTcStgLd);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            so[i] = make_float4(y[16 * half + 4 * i], y[16 * half + 4 * i + 1], y[16 * half + 4 * i + 2],
                                y[16 * half + 4 * i + 3]);
          soft_bar(eg + 3);  // barrier ids 4 and 5: the 128 threads of this epilogue group
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = r + u * 128, rr = i >> 2, c4 = i & 3;
            if (rr < rows_here)
              *reinterpret_cast<float4*>(obase + (size_t)rr * kTcF + cq * 32 + 16 * half + 4 * c4) =
                  *reinterpret_cast<const float4*>(stg + (size_t)rr * kTcStgLd + 4 * c4);
          }
          soft_bar(eg + 3);
        }
      }
      ++tc;
    }
  } else if (w == kBcWorkWarps) {
    // =============================== MMA issue ====================================================
    if (lane == 0) {
      uint32_t sc = 0, tc = 0;
      constexpr uint32_t id256 = umma_idesc_tf32(kTcM, 256), id128 = umma_idesc_tf32(kTcM, 128);
      const uint32_t ring = smem_u32(smem);
      TcItems t(p.blk_ptr, p.n_blocks, p.sched_ptr, p.sched_idx);
      while (t.next()) {
        mbar_wait(&acc_free, (tc & 1u) ^ 1u);
        tc_fence_after();
        const int NS = t.stages();
        for (int s = 0; s < NS; ++s, ++sc) {
          const uint32_t slot = sc % kTcSlots, k = sc / kTcSlots;
          mbar_wait(&full_b[slot], k & 1u);
          mbar_wait(&full_a[slot], k & 1u);
          tc_fence_after();
          const uint32_t base = ring + slot * kTcSlotBytes;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint32_t ao = ks * 2 * kTcLboA, bo = ks * 2 * kTcLboB;
            const uint64_t dph = umma_desc_kmajor(base + ao, kTcLboA, kTcSBO);
            const uint64_t dpl = umma_desc_kmajor(base + kBcA + ao, kTcLboA, kTcSBO);
            const uint64_t dsh = umma_desc_kmajor(base + 2 * kBcA + ao, kTcLboA, kTcSBO);
            const uint64_t dsl = umma_desc_kmajor(base + 3 * kBcA + ao, kTcLboA, kTcSBO);
            const uint64_t db0 = umma_desc_kmajor(base + kBcOffB0 + bo, kTcLboB, kTcSBO);
            const uint64_t db1 = umma_desc_kmajor(base + kBcOffB1 + bo, kTcLboB, kTcSBO);
            const uint32_t acc = (s | ks) != 0 ? 1u : 0u;
            umma_tf32(tmem, dph, db0, id256, acc);        // dV += P^T_hi [dO_hi | dO_lo]
            umma_tf32(tmem + 128, dpl, db0, id128, 1u);   // dV (second half) += P^T_lo dO_hi
            umma_tf32(tmem + 256, dsh, db1, id256, acc);  // dK += dS^T_hi [Q_hi | Q_lo]
            umma_tf32(tmem + 384, dsl, db1, id128, 1u);   // dK (second half) += dS^T_lo Q_hi
          }
          umma_commit(&empty[slot]);
        }
        umma_commit(&acc_full);
        ++tc;
      }
    }
  } else {
    // =============================== loaders: dO^T and Q^T slices ==================================
    const uint32_t lg = (w - kBcWorkWarps - 1) >> 2;             // loader group: stages with (stage & 1) == lg
    const int lf = (tid - (kBcWorkWarps + 1) * 32) & 127;         // feature
    uint32_t sc = 0;
    TcItems t(p.blk_ptr, p.n_blocks, p.sched_ptr, p.sched_idx);
    while (t.next()) {
      const int NS = t.stages();
      for (int s = 0; s < NS; ++s, ++sc) {
        if ((sc & 1u) != lg) continue;
        const uint32_t slot = sc % kTcSlots;
        const float* gsrc = p.dO + (size_t)(t.lb + 16 * s) * kTcF + lf;
        const float* qsrc = p.Q + (size_t)(t.lb + 16 * s) * kTcF + lf;
        const int rows = t.n - 16 * s;  // valid rows of this slice (>= 1)
        float4 xg[4], xq[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          xg[c].x = 4 * c + 0 < rows ? __ldg(gsrc + (size_t)(4 * c + 0) * kTcF) : 0.f;
          xg[c].y = 4 * c + 1 < rows ? __ldg(gsrc + (size_t)(4 * c + 1) * kTcF) : 0.f;
          xg[c].z = 4 * c + 2 < rows ? __ldg(gsrc + (size_t)(4 * c + 2) * kTcF) : 0.f;
          xg[c].w = 4 * c + 3 < rows ? __ldg(gsrc + (size_t)(4 * c + 3) * kTcF) : 0.f;
          xq[c].x = 4 * c + 0 < rows ? __ldg(qsrc + (size_t)(4 * c + 0) * kTcF) : 0.f;
          xq[c].y = 4 * c + 1 < rows ? __ldg(qsrc + (size_t)(4 * c + 1) * kTcF) : 0.f;
          xq[c].z = 4 * c + 2 < rows ? __ldg(qsrc + (size_t)(4 * c + 2) * kTcF) : 0.f;
          xq[c].w = 4 * c + 3 < rows ? __ldg(qsrc + (size_t)(4 * c + 3) * kTcF) : 0.f;
        }
        mbar_wait(&empty[slot], ((sc / kTcSlots) & 1u) ^ 1u);
        float4* b0 = reinterpret_cast<float4*>(smem + (size_t)slot * kTcSlotBytes + kBcOffB0);
        float4* b1 = reinterpret_cast<float4*>(smem + (size_t)slot * kTcSlotBytes + kBcOffB1);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float4 hi, lo;
          split4(xg[c], hi, lo);
          b0[c * 256 + lf] = hi;
          b0[c * 256 + 128 + lf] = lo;
          split4(xq[c], hi, lo);
          b1[c * 256 + lf] = hi;
          b1[c * 256 + 128 + lf] = lo;
        }
        fence_proxy_async();
        mbar_arrive(&full_b[slot]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

// attn_edge (CSR order) -> dense probabilities as tile images (tile_at; zeros where there is no edge, up to the
// next multiple of 32 columns of each graph).  gridDim.y CTAs share a graph; a warp takes 8 rows at a time:
// lane = (row r8 = lane & 7, 4-column quad q = lane >> 3), 16 columns per step, so that a store instruction
// writes four 128-byte runs of the chunk-major image (8 rows x one float4 each per quad).
static __global__ void block_attn_dense_kernel(const int* __restrict__ blk_ptr, const int* __restrict__ row_ptr,
                                               const uint32_t* __restrict__ bits, const float* __restrict__ attn,
                                               const int* __restrict__ tile_ptr, float* __restrict__ Pd) {
  const int b = blockIdx.x, lb = blk_ptr[b], n = blk_ptr[b + 1] - lb, tile0 = tile_ptr[b];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5, n32 = (n + 31) & ~31;
  const int r8 = lane & 7, q = lane >> 3;
  for (int rb = 8 * (w + nw * blockIdx.y); rb < n; rb += 8 * nw * gridDim.y) {
    const int r = rb + r8;
    if (r >= n) continue;  // no warp-level operations below
    const uint4* src = reinterpret_cast<const uint4*>(bits + (size_t)(lb + r) * kTcMaskW);
    const uint4 w0 = __ldg(src), w1 = __ldg(src + 1);
    const uint32_t wd[kTcMaskW] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    const float* arow = attn + __ldg(row_ptr + lb + r);
    float4* img = reinterpret_cast<float4*>(Pd + (size_t)(tile0 + (r >> 7)) * kTcTileFloats) + (r & 127);
    int before = 0;  // set bits of the words in front of the current one
#pragma unroll
    for (int k = 0; k < kTcMaskW; ++k) {
      if (32 * k < n32) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int bit = 16 * half + 4 * q;  // first of this thread's four columns inside word k
          const uint32_t nib = (wd[k] >> bit) & 0xFu;
          const float* at = arow + before + __popc(wd[k] & ((1u << bit) - 1u));
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (nib & 1u) v.x = __ldg(at);
          if (nib & 2u) v.y = __ldg(at + __popc(nib & 1u));
          if (nib & 4u) v.z = __ldg(at + __popc(nib & 3u));
          if (nib & 8u) v.w = __ldg(at + __popc(nib & 7u));
          img[(size_t)((32 * k + bit) >> 2) * kTcM] = v;
        }
        before += __popc(wd[k]);
      }
    }
  }
}

// Test hook: fills all 512 tensor-memory columns of every SM with NaN bit patterns, so that a kernel that reads
// tensor memory it has not written shows up in the parity tests (tensor memory keeps its contents between kernels).
static __global__ void __launch_bounds__(128, 1) tc_poison_tmem_kernel() {
  __shared__ uint32_t s_tmem;
  const int w = threadIdx.x >> 5;
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem, nan = 0x7fc00000u;
  for (int c = 0; c < 512; c += 8)
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(
                     tmem + ((uint32_t)(w * 32) << 16) + c),
                 "r"(nan)
                 : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
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

// Host-side work lists for the persistent CTAs (no device work): longest processing time first over a cost
// model of the MMA work.  column_items = 0: one entry per graph (forward / backward row side; a graph of more
// than 128 nodes is two row tiles over up to 256 keys); 1: one entry per (graph, 128-key tile), id = 2 * graph
// + tile (backward column side; cost = 16-row slices of the graph + the epilogue).  nodes: host array
// [n_blocks].  ptr_out [n_ctas_out + 1], idx_out [items]; returns the number of CTAs (<= n_ctas), or < 0.
int dfgnn_tc_balanced_lists(int n_blocks, const int32_t* nodes, int n_ctas, int column_items, int32_t* ptr_out,
                            int32_t* idx_out) {
  if (n_blocks < 1 || n_ctas < 1 || !nodes || !ptr_out || !idx_out) return DFGNN_ERR_INVALID_ARGUMENT;
  struct Item { long long cost; int id; };
  std::vector<Item> items;
  items.reserve(2 * (size_t)n_blocks);
  for (int b = 0; b < n_blocks; ++b) {
    const long long n = nodes[b] > 0 ? nodes[b] : 0;
    if (column_items) {
      const long long c = ((n + 15) / 16) * 1000 + 4000;
      items.push_back({c, 2 * b});
      if (n > kTcM) items.push_back({c, 2 * b + 1});
    } else {
      const long long sl = (n + 31) / 32;
      items.push_back({n <= kTcM ? 16 * 246 + sl * 4 * 246 : 2 * (16 * 384 + sl * 4 * 246), b});
    }
  }
  std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.cost > b.cost; });
  const int g = (int)std::min<size_t>(items.size(), (size_t)n_ctas);
  using Load = std::pair<long long, int>;  // (load, cta): smallest load first, ties to the lower CTA
  std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
  for (int c = 0; c < g; ++c) heap.push({0, c});
  std::vector<std::vector<int>> lists(g);
  for (const Item& it : items) {
    Load l = heap.top();
    heap.pop();
    lists[l.second].push_back(it.id);
    heap.push({l.first + it.cost, l.second});
  }
  int at = 0;
  for (int c = 0; c < g; ++c) {
    ptr_out[c] = at;
    for (int id : lists[c]) idx_out[at++] = id;
  }
  ptr_out[g] = at;
  return g;
}

int dfgnn_tc_poison_tmem(void* stream) {
  tc_poison_tmem_kernel<<<4 * sm_count(), 128, 0, (cudaStream_t)stream>>>();
  return check_launch("dfgnn_tc_poison_tmem");
}

int dfgnn_gt_dense_tc_supported(int max_nodes, int h, int f) { return dense_tc_supported(max_nodes, h, f) ? 1 : 0; }

int dfgnn_block_adj_bits(int n_blocks, int max_nodes, int m, int nnz, const int32_t* blk_ptr, const int32_t* row_ptr,
                         const int32_t* col_ind, uint32_t* adj_bits, void* stream) {
  const char* fn = "dfgnn_block_adj_bits";
  if (n_blocks < 1 || m < 0 || nnz < 0 || max_nodes < 1 || max_nodes > kTcMaxNodes) {
    set_error("%s: needs a block plan with graphs of at most %d nodes (n_blocks=%d, max_nodes=%d)", fn, kTcMaxNodes,
              n_blocks, max_nodes);
    return DFGNN_ERR_INVALID_ARGUMENT;
  }
  if (m == 0) return DFGNN_OK;
  DFGNN_REQUIRE(blk_ptr, fn); DFGNN_REQUIRE(row_ptr, fn); DFGNN_REQUIRE(adj_bits, fn);
  if (nnz > 0) DFGNN_REQUIRE(col_ind, fn);
  block_adj_bits_kernel<<<n_blocks, 256, 0, (cudaStream_t)stream>>>(n_blocks, blk_ptr, row_ptr, col_ind, adj_bits);
  return check_launch(fn);
}

int dfgnn_gt_dense_tc_forward(int n_blocks, const int32_t* blk_ptr, int max_nodes, int m, int nnz, int h, int f,
                              const int32_t* row_ptr, const uint32_t* adj_bits, int n_ctas, const int32_t* sched_ptr,
                              const int32_t* sched_idx, const float* Q, const float* K, const float* V, float* out_feat,
                              float* attn_edge, void* stream) {
  const char* fn = "dfgnn_gt_dense_tc_forward";
  if (int rc = check_common(fn, m, nnz, h, f)) return rc;
  if (m == 0) return DFGNN_OK;
  DFGNN_REQUIRE(blk_ptr, fn); DFGNN_REQUIRE(row_ptr, fn); DFGNN_REQUIRE(adj_bits, fn);
  DFGNN_REQUIRE(Q, fn); DFGNN_REQUIRE(K, fn); DFGNN_REQUIRE(V, fn); DFGNN_REQUIRE(out_feat, fn);
  if (n_blocks < 1 || !dense_tc_supported(max_nodes, h, f)) {
    set_error("%s: needs h == 1, f == %d and graphs of at most %d nodes (h=%d, f=%d, max_nodes=%d)", fn, kTcF,
              kTcMaxNodes, h, f, max_nodes);
    return DFGNN_ERR_UNSUPPORTED_DIM;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const bool sched = sched_ptr != nullptr && sched_idx != nullptr && n_ctas >= 1;
  GtTcParams p{n_blocks, blk_ptr, row_ptr, sched ? sched_ptr : nullptr, sched ? sched_idx : nullptr, adj_bits,
               Q, K, V, out_feat, nnz > 0 ? attn_edge : nullptr, nullptr, nullptr, nullptr};
  auto kernel = gt_dense_tc_fwd_kernel;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes);
  const int grid = sched ? n_ctas : (n_blocks < sm_count() ? n_blocks : sm_count());
  kernel<<<grid, kTcThreads, kTcSmemBytes, st>>>(p);
  note_kernel(0, "gt_dense_tc_fwd_kernel");
  return check_launch(fn);
}

int dfgnn_gt_dense_tc_backward_col(int n_blocks, const int32_t* blk_ptr, int max_nodes, int m, int nnz, int h, int f,
                                   const int32_t* row_ptr, const uint32_t* adj_bits, int n_ctas, const int32_t* sched_ptr,
                                   const int32_t* sched_idx, const float* Q, const float* grad_out, const float* grad_edge,
                                   float* grad_K, float* grad_V, void* stream) {
  const char* fn = "dfgnn_gt_dense_tc_backward_col";
  if (int rc = check_common(fn, m, nnz, h, f)) return rc;
  if (m == 0) return DFGNN_OK;
  DFGNN_REQUIRE(blk_ptr, fn); DFGNN_REQUIRE(row_ptr, fn); DFGNN_REQUIRE(adj_bits, fn);
  DFGNN_REQUIRE(Q, fn); DFGNN_REQUIRE(grad_out, fn); DFGNN_REQUIRE(grad_K, fn); DFGNN_REQUIRE(grad_V, fn);
  if (nnz > 0) DFGNN_REQUIRE(grad_edge, fn);
  if (n_blocks < 1 || !dense_tc_supported(max_nodes, h, f)) {
    set_error("%s: needs h == 1, f == %d and graphs of at most %d nodes (h=%d, f=%d, max_nodes=%d)", fn, kTcF,
              kTcMaxNodes, h, f, max_nodes);
    return DFGNN_ERR_UNSUPPORTED_DIM;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const bool sched = sched_ptr != nullptr && sched_idx != nullptr && n_ctas >= 1;
  GtTcBwdColParams p{n_blocks, blk_ptr, row_ptr, sched ? sched_ptr : nullptr, sched ? sched_idx : nullptr, adj_bits,
                     reinterpret_cast<const float2*>(grad_edge), nullptr, nullptr, nullptr, grad_out, Q, grad_K, grad_V};
  auto kernel = gt_dense_tc_bwd_col_kernel<false>;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBcSmemBytes);
  const int grid = sched ? n_ctas : (2 * n_blocks < sm_count() ? 2 * n_blocks : sm_count());
  kernel<<<grid, kBcThreads, kBcSmemBytes, st>>>(p);
  note_kernel(2, "gt_dense_tc_bwd_col_kernel");
  return check_launch(fn);
}

size_t dfgnn_gt_dense_tc_backward_ws_floats(int m, int n_blocks) {
  // two arrays of one 128 x 256 image per row tile; sum over graphs of ceil(n / 128) <= m / 128 + n_blocks
  const size_t tiles = (size_t)(m > 0 ? m : 0) / kTcM + (size_t)(n_blocks > 0 ? n_blocks : 0) + 1;
  return 2 * tiles * kTcTileFloats;
}

int dfgnn_gt_dense_tc_backward(int phases, int n_blocks, const int32_t* blk_ptr, int max_nodes, int m, int nnz, int h,
                               int f, const int32_t* row_ptr, const uint32_t* adj_bits, int n_ctas,
                               const int32_t* sched_ptr, const int32_t* sched_idx, int n_ctas_col,
                               const int32_t* sched_ptr_col, const int32_t* sched_idx_col, const float* Q,
                               const float* K, const float* V, const float* attn_edge, const float* grad_out,
                               float* grad_Q, float* grad_K, float* grad_V, const int32_t* tile_ptr, float* dense_ws,
                               void* stream) {
  const char* fn = "dfgnn_gt_dense_tc_backward";
  if (phases < 1 || phases > 3) { set_error("%s: phases=%d must be 1, 2 or 3", fn, phases); return DFGNN_ERR_INVALID_ARGUMENT; }
  if (int rc = check_common(fn, m, nnz, h, f)) return rc;
  if (m == 0) return DFGNN_OK;
  DFGNN_REQUIRE(blk_ptr, fn); DFGNN_REQUIRE(row_ptr, fn); DFGNN_REQUIRE(adj_bits, fn); DFGNN_REQUIRE(dense_ws, fn);
  DFGNN_REQUIRE(tile_ptr, fn);
  DFGNN_REQUIRE(Q, fn); DFGNN_REQUIRE(K, fn); DFGNN_REQUIRE(V, fn); DFGNN_REQUIRE(grad_out, fn);
  DFGNN_REQUIRE(grad_Q, fn); DFGNN_REQUIRE(grad_K, fn); DFGNN_REQUIRE(grad_V, fn);
  if (nnz > 0) DFGNN_REQUIRE(attn_edge, fn);
  if (n_blocks < 1 || !dense_tc_supported(max_nodes, h, f)) {
    set_error("%s: needs h == 1, f == %d and graphs of at most %d nodes (h=%d, f=%d, max_nodes=%d)", fn, kTcF,
              kTcMaxNodes, h, f, max_nodes);
    return DFGNN_ERR_UNSUPPORTED_DIM;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float* Pd = dense_ws;
  float* dSd = dense_ws + dfgnn_gt_dense_tc_backward_ws_floats(m, n_blocks) / 2;
  if (phases & 1) {
    block_attn_dense_kernel<<<dim3(n_blocks, 4), 256, 0, st>>>(blk_ptr, row_ptr, adj_bits, attn_edge, tile_ptr, Pd);
    if (int rc = check_launch(fn)) return rc;
    const bool sched = sched_ptr != nullptr && sched_idx != nullptr && n_ctas >= 1;
    GtTcParams p{n_blocks, blk_ptr, row_ptr, sched ? sched_ptr : nullptr, sched ? sched_idx : nullptr, adj_bits,
                 grad_out, V, K, grad_Q, nullptr, Pd, dSd, tile_ptr};
    auto kernel = gt_dense_tc_bwd_row_kernel;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes);
    const int grid = sched ? n_ctas : (n_blocks < sm_count() ? n_blocks : sm_count());
    kernel<<<grid, kTcThreads, kTcSmemBytes, st>>>(p);
    note_kernel(1, "gt_dense_tc_bwd_row_kernel");
    if (int rc = check_launch(fn)) return rc;
  }
  if (phases & 2) {
    const bool sched = sched_ptr_col != nullptr && sched_idx_col != nullptr && n_ctas_col >= 1;
    GtTcBwdColParams p{n_blocks, blk_ptr, row_ptr, sched ? sched_ptr_col : nullptr, sched ? sched_idx_col : nullptr, adj_bits,
                       nullptr, Pd, dSd, tile_ptr, grad_out, Q, grad_K, grad_V};
    auto kernel = gt_dense_tc_bwd_col_kernel<true>;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBcSmemBytes);
    const int grid = sched ? n_ctas_col : (2 * n_blocks < sm_count() ? 2 * n_blocks : sm_count());
    kernel<<<grid, kBcThreads, kBcSmemBytes, st>>>(p);
    note_kernel(2, "gt_dense_tc_bwd_col_kernel");
    if (int rc = check_launch(fn)) return rc;
  }
  return DFGNN_OK;
}

#ifdef DFGNN_TC_PROF
__attribute__((visibility("default"))) int dfgnn_tc_prof_read(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_tc_prof, sizeof(g_tc_prof));
  if (reset) {
    unsigned long long z[32] = {0};
    cudaMemcpyToSymbol(g_tc_prof, z, sizeof(z));
  }
  return 0;
}
#endif

}  // extern "C"
