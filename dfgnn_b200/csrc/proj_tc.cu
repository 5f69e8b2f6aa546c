// proj_tc.cu -- the dense projection in front of the conv on the 5th-generation tensor cores.
//
// Reference: q, k, v = Linear(h) (+ q *= d^-0.5) then reshape / transpose to [N, heads, d]
// (DFGNN/layers/GT/gtconv_layer.py:19-27, gtconv_layer_fused.py:20-22), and for GAT
// feat = W(x), attn_row = <a_l, feat>, attn_col = <a_r, feat> (layers/GAT/gatconv_layer_fused.py:
// 121-123; fused logits kernel fused_gatconv_hyper_v2.cu:212-250): three fp32 cuBLAS GEMMs, a
// scale, transposes and two reductions.  Measured on B200 for the PATTERN-shaped batch (121.8 k
// nodes, 128 -> 3 x 128): 0.49 ms in fp32 (SIMT SGEMM), longer than the fused conv forward itself.
//
// Here: ONE kernel, Y = (X W^T + b) * scale written straight into the operands of the conv
// ([N, heads * d] per part, i.e. [N, heads, d]), fp32-grade accuracy from TF32 tensor cores by the
// 3xTF32 split (x = hi + lo; x w ~ hi*hi + hi*lo + lo*hi, fp32 accumulation in tensor memory), and
// the GAT logits reduced in the epilogue from the accumulator rows.
//
//   tcgen05.mma.cta_group::1.kind::tf32, M = 128 (a tile of node rows), N = 128 (or 64) output
//   columns, K = 8 per instruction; operands in shared memory in the canonical K-major
//   no-swizzle layout ("chunk major": 16-byte k-chunk c of row r at c * rows * 16 + r * 16, so
//   SBO = 128 B between 8-row groups and LBO = rows * 16 B between the two chunks of one MMA);
//   accumulators: two blocks of 128 lanes x N columns of tensor memory, read back with
//   tcgen05.ld.32x32b.x32.
//
// A persistent CTA walks over node tiles.  Per tile the X images are built ONCE (producer warps:
// coalesced float4 loads prefetched a tile ahead, hi / lo split in registers, stores into the
// canonical layout) and the pre-split W images (prepared once per weight update,
// dfgnn_proj_pack_weights) stream through a two-stage ring by TMA bulk copies, 32 input columns per
// stage; one thread issues 12 MMAs per stage; the epilogue warps drain accumulator block b while the
// MMAs of block b+1 run.  Roles synchronise through mbarriers only (see the kernel).
#include "abi_common.h"
#include "tc_common.cuh"

namespace dfgnn {

constexpr int kProjM = 128;      // node rows per tile (UMMA M)
constexpr int kProjMaxK = 128;

struct ProjParams {
  int n, k, n_out;        // rows, input width, total output columns (multiple of 64)
  int part_width;         // columns of one output tensor (q | k | v are three parts); multiple of 64
  const float* x;         // [n, k]
  const float* w_img;     // [2][n_out / 64][k / 4][64][4]: hi images then lo images
  const float* bias;      // [n_out] or null
  const float* scale;     // [n_out] or null
  float *out0, *out1, *out2, *out3;  // one [n, part_width] tensor per part (separate members: a
                                     // dynamically indexed array would move the whole block to local memory)
  // GAT logits (optional): per head of width head_dim, attn_row = <a_l, y>, attn_col = <a_r, y>
  int head_dim;           // 0: no logits
  const float* a_l;       // [n_out]
  const float* a_r;       // [n_out]
  float* attn_row;        // [n, n_out / head_dim]
  float* attn_col;
};

// Warp-specialised, persistent: CTA = one SM, walks node tiles; per tile the X images are built once
// and W streams through a 2-stage ring in stages of NS output columns x 32 input columns (hi + lo).
// UMMA N = NS (128 where n_out allows, else 64): small-N MMAs have a fixed minimum cost -- measured,
// N = 32 ran at a quarter of the tensor rate -- so the ring is cut along K, not along N.
//   warps 0-7   epilogue: TMEM lane quarter w % 4, every other 32-column piece (w / 4), tcgen05.ld,
//               bias / scale / logits, 128 bytes of the thread's own output row to global
//   warp  8     MMA issue (one lane): 4 k-steps x 3 tcgen05.mma per ring stage, K/32 stages per
//               accumulator block (NS columns), two accumulator blocks in tensor memory
//   warp  9     TMA producer (one lane): W stage images -> ring, mbarrier complete_tx
//   warps 10-17 X producers: global float4 -> registers (prefetched one tile ahead) -> hi / lo split
//               -> canonical shared-memory images
// mbarriers: a_full / a_empty (X images), b_full / b_empty [2] (W ring), t_full / t_empty [2]
// (accumulator blocks); tcgen05.commit arrives on the "empty" / "full" barriers when the MMAs that
// read the operands / wrote the accumulator have completed.
constexpr int kEpiWarps = 8, kProdWarps = 8;
constexpr int kProjThreads2 = (kEpiWarps + 2 + kProdWarps) * 32;  // 576
constexpr int kStageK = 32;                                        // input columns per W ring stage
constexpr int kProjMaxOut = 1024;                                  // bias / scale staged in shared memory
constexpr int kStageCols = 16, kStageLd = kStageCols + 4;          // epilogue staging: 128 rows x 16 columns per group

// barrier among the 4 warps (128 threads) of one epilogue group
__device__ __forceinline__ void epi_bar_sync(int group) { asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory"); }

template <int K, int NS>
__global__ void __launch_bounds__(kProjThreads2, 1) proj_tf32x3_kernel(const ProjParams p) {
  constexpr int CH = K / 4;                                   // 16-byte k-chunks per row
  constexpr int KQ = K / kStageK;                             // ring stages per accumulator block
  constexpr int PT = kProdWarps * 32;                         // X producer threads
  constexpr int XV = kProjM * CH / PT;                        // float4 of the X tile per producer thread
  constexpr uint32_t A_LBO = kProjM * 16, B_LBO = NS * 16, SBO = 128;
  constexpr uint32_t A_BYTES = kProjM * K * 4;                // one X image
  constexpr uint32_t B_BYTES = NS * kStageK * 4;              // one W stage image
  static_assert(K % kStageK == 0 && K <= kProjMaxK && (kProjM * CH) % PT == 0 && (NS == 64 || NS == 128), "shape");
  extern __shared__ __align__(128) unsigned char smem[];
  float4* sA_hi = reinterpret_cast<float4*>(smem);
  float4* sA_lo = reinterpret_cast<float4*>(smem + A_BYTES);
  unsigned char* sB = smem + 2 * A_BYTES;                     // [2 stages][hi | lo] images
  float* s_bias = reinterpret_cast<float*>(sB + 4 * B_BYTES); // [n_out] bias * scale
  float* s_scale = s_bias + kProjMaxOut;                      // [n_out]
  float* s_stage = s_scale + kProjMaxOut;                     // [2 groups][128][kStageLd]
  __shared__ uint64_t a_full, a_empty, b_full[2], b_empty[2], t_full[2], t_empty[2];
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int n_blocks = p.n_out / NS;                          // accumulator blocks per tile
  const int tiles = (p.n + kProjM - 1) / kProjM;

  for (int i = tid; i < p.n_out; i += kProjThreads2) {  // y = acc * scale + bias * scale
    const float sc = p.scale ? __ldg(p.scale + i) : 1.f;
    s_scale[i] = sc;
    s_bias[i] = (p.bias ? __ldg(p.bias + i) : 0.f) * sc;
  }
  if (tid == 0) {
    mbar_init(&a_full, PT);
    mbar_init(&a_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], kEpiWarps * 32);
    }
    mbar_fence_init();
  }
  if (w == 0) {  // 2 * NS columns of tensor memory: two 128 x NS fp32 accumulator blocks
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(2 * NS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = s_tmem;

  if (w < kEpiWarps) {
    // ================================ epilogue ================================================
    const int q4 = w & 3, half = w >> 2;
    const int erow = q4 * 32 + lane;  // row of the tile = TMEM lane
    uint32_t ab = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      const int row = t * kProjM + erow;
      for (int blk = 0; blk < n_blocks; ++blk, ++ab) {
        const uint32_t st = ab & 1u, ph = (ab >> 1) & 1u;
        mbar_wait(&t_full[st], ph);
        tc_fence_after();
#pragma unroll 1
        for (int cq = half; cq < NS / 32; cq += 2) {
          float y[32];
          tmem_ld32(tmem_d + ((uint32_t)(q4 * 32) << 16) + st * NS + cq * 32, y);
          if (cq + 2 >= NS / 32) {  // this warp's last read of the accumulator block
            tc_fence_before();
            mbar_arrive(&t_empty[st]);
          }
          const int col0 = blk * NS + cq * 32;
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + col0);
          const float4* s4 = reinterpret_cast<const float4*>(s_scale + col0);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b = b4[i], sc = s4[i];
            y[4 * i] = fmaf(y[4 * i], sc.x, b.x);
            y[4 * i + 1] = fmaf(y[4 * i + 1], sc.y, b.y);
            y[4 * i + 2] = fmaf(y[4 * i + 2], sc.z, b.z);
            y[4 * i + 3] = fmaf(y[4 * i + 3], sc.w, b.w);
          }
          {  // 128 rows x 32 columns -> global through this group's staging buffer, 16 columns at a
             // time: a warp store covers 8 rows x 64 contiguous bytes instead of 32 rows x 16 bytes
            const int part = col0 / p.part_width;
            float* obase = (part == 0 ? p.out0 : part == 1 ? p.out1 : part == 2 ? p.out2 : p.out3) + col0 % p.part_width;
            float* stg = s_stage + (size_t)half * kProjM * kStageLd;
            const int gt = tid - half * 128;  // thread index inside the group
#pragma unroll
            for (int sp = 0; sp < 2; ++sp) {
              float4* so = reinterpret_cast<float4*>(stg + (size_t)erow * kStageLd);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                so[i] = make_float4(y[16 * sp + 4 * i], y[16 * sp + 4 * i + 1], y[16 * sp + 4 * i + 2], y[16 * sp + 4 * i + 3]);
              epi_bar_sync(half);
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int i = gt + u * 128, r = i >> 2, c4 = i & 3;
                const int grow = t * kProjM + r;
                if (grow < p.n)
                  *reinterpret_cast<float4*>(obase + (size_t)grow * p.part_width + 16 * sp + 4 * c4) =
                      *reinterpret_cast<const float4*>(stg + (size_t)r * kStageLd + 4 * c4);
              }
              epi_bar_sync(half);
            }
          }
          if (p.head_dim > 0 && row < p.n) {
            const int hd = p.head_dim, heads = p.n_out / hd;
            float pl[32], pr[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              pl[j] = y[j] * __ldg(p.a_l + col0 + j);
              pr[j] = y[j] * __ldg(p.a_r + col0 + j);
            }
            float l8[4], r8[4];
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8) {
              float a = 0.f, b = 0.f;
#pragma unroll
              for (int j = 0; j < 8; ++j) { a += pl[8 * k8 + j]; b += pr[8 * k8 + j]; }
              l8[k8] = a;
              r8[k8] = b;
            }
            const size_t at = (size_t)row * heads + col0 / hd;
            if (hd == 8) {
#pragma unroll
              for (int k8 = 0; k8 < 4; ++k8) { p.attn_row[at + k8] = l8[k8]; p.attn_col[at + k8] = r8[k8]; }
            } else if (hd == 16) {
              p.attn_row[at] = l8[0] + l8[1]; p.attn_row[at + 1] = l8[2] + l8[3];
              p.attn_col[at] = r8[0] + r8[1]; p.attn_col[at + 1] = r8[2] + r8[3];
            } else if (hd == 32) {
              p.attn_row[at] = (l8[0] + l8[1]) + (l8[2] + l8[3]);
              p.attn_col[at] = (r8[0] + r8[1]) + (r8[2] + r8[3]);
            } else {  // heads wider than 32 columns: one atomicAdd per piece (arrays zeroed by the launcher)
              atomicAdd(p.attn_row + at, (l8[0] + l8[1]) + (l8[2] + l8[3]));
              atomicAdd(p.attn_col + at, (r8[0] + r8[1]) + (r8[2] + r8[3]));
            }
          }
        }
        if (half >= NS / 32) {  // NS = 32 would leave the upper half without a piece (not instantiated)
          tc_fence_before();
          mbar_arrive(&t_empty[st]);
        }
      }
    }
  } else if (w == kEpiWarps) {
    // ================================ MMA issue ================================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(kProjM, NS);
      const uint32_t a_hi = smem_u32(sA_hi), a_lo = smem_u32(sA_lo), b0 = smem_u32(sB);
      uint32_t it = 0, ab = 0, tt = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++tt) {
        mbar_wait(&a_full, tt & 1u);
        for (int blk = 0; blk < n_blocks; ++blk, ++ab) {
          const uint32_t as = ab & 1u, aph = (ab >> 1) & 1u;
          mbar_wait(&t_empty[as], aph ^ 1u);
          const uint32_t d = tmem_d + as * NS;
          for (int kq = 0; kq < KQ; ++kq, ++it) {
            const uint32_t st = it & 1u, ph = (it >> 1) & 1u;
            mbar_wait(&b_full[st], ph);
            tc_fence_after();
            const uint32_t b_hi = b0 + st * 2 * B_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
            for (int ks = 0; ks < kStageK / 8; ++ks) {
              const uint32_t ka = (kq * (kStageK / 8) + ks) * 2 * A_LBO;
              const uint64_t dah = umma_desc_kmajor(a_hi + ka, A_LBO, SBO);
              const uint64_t dal = umma_desc_kmajor(a_lo + ka, A_LBO, SBO);
              const uint64_t dbh = umma_desc_kmajor(b_hi + ks * 2 * B_LBO, B_LBO, SBO);
              const uint64_t dbl = umma_desc_kmajor(b_lo + ks * 2 * B_LBO, B_LBO, SBO);
              umma_tf32(d, dal, dbh, idesc, (kq | ks) != 0 ? 1u : 0u);
              umma_tf32(d, dah, dbl, idesc, 1u);
              umma_tf32(d, dah, dbh, idesc, 1u);
            }
            umma_commit(&b_empty[st]);  // the W stage may be refilled
          }
          umma_commit(&t_full[as]);     // the accumulator block is complete
          if (blk == n_blocks - 1) umma_commit(&a_empty);  // the X images may be replaced
        }
      }
    }
  } else if (w == kEpiWarps + 1) {
    // ================================ TMA producer =============================================
    if (lane == 0) {
      const size_t img = (size_t)NS * kStageK;           // floats of one stage image
      const size_t lo_off = (size_t)p.n_out * K;         // lo images follow all hi images
      uint32_t it = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        for (int blk = 0; blk < n_blocks; ++blk) {
          for (int kq = 0; kq < KQ; ++kq, ++it) {
            const uint32_t st = it & 1u, ph = (it >> 1) & 1u;
            mbar_wait(&b_empty[st], ph ^ 1u);
            mbar_expect_tx(&b_full[st], 2 * B_BYTES);
            unsigned char* dst = sB + st * 2 * B_BYTES;
            const float* src = p.w_img + ((size_t)blk * KQ + kq) * img;
            bulk_g2s_range(dst, src, B_BYTES, &b_full[st]);
            bulk_g2s_range(dst + B_BYTES, src + lo_off, B_BYTES, &b_full[st]);
          }
        }
      }
    }
  } else {
    // ================================ X producers ==============================================
    const int ptid = tid - (kEpiWarps + 2) * 32;
    // float4 i = ptid + u * PT covers row (i % 8) + 8 * (i / (8 * CH)), chunk (i / 8) % CH: a warp reads
    // 8 rows x 64 contiguous bytes and writes 512 contiguous bytes of the chunk-major image
    auto piece = [&](int u, int& r, int& c) {
      const int i = ptid + u * PT;
      r = (i & 7) + 8 * (i / (8 * CH));
      c = (i >> 3) % CH;
    };
    float4 xv[XV];
    auto load_tile = [&](int t) {
#pragma unroll
      for (int u = 0; u < XV; ++u) {
        int r, c;
        piece(u, r, c);
        const int row = t * kProjM + r;
        xv[u] = row < p.n ? __ldg(reinterpret_cast<const float4*>(p.x + (size_t)row * K) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    uint32_t tt = 0;
    int t = blockIdx.x;
    if (t < tiles) load_tile(t);
    for (; t < tiles; t += gridDim.x, ++tt) {
      mbar_wait(&a_empty, (tt & 1u) ^ 1u);  // the MMAs of the previous tile have read the images
#pragma unroll
      for (int u = 0; u < XV; ++u) {
        int r, c;
        piece(u, r, c);
        float4 hi, lo;
        split4(xv[u], hi, lo);
        sA_hi[c * kProjM + r] = hi;
        sA_lo[c * kProjM + r] = lo;
      }
      fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's operand reads
      mbar_arrive(&a_full);
      if (t + (int)gridDim.x < tiles) load_tile(t + gridDim.x);  // in flight during this tile's MMAs
    }
  }
  // ---- teardown ------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(2 * NS) : "memory");
}

// W [n_out, k] row major -> hi / lo images in the chunk-major layout of a 64-row slice
// image order: [hi | lo][n_out / ns column blocks][k / 32 stages][8 chunks][ns rows][4 floats]
static __global__ void proj_pack_kernel(int n_out, int k, int ns, const float* __restrict__ W, float* __restrict__ img) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // one float4 (row o, chunk c)
  const int ch = k / 4;
  if (i >= n_out * ch) return;
  const int o = i / ch, c = i % ch;
  const float4 x = __ldg(reinterpret_cast<const float4*>(W + (size_t)o * k) + c);
  float4 hi, lo;
  split4(x, hi, lo);
  const int blk = o / ns, r = o % ns, kq = c / 8, c8 = c % 8;
  const size_t at = (((size_t)blk * (ch / 8) + kq) * 8 + c8) * ns + r;  // float4 index inside the hi images
  reinterpret_cast<float4*>(img)[at] = hi;
  reinterpret_cast<float4*>(img)[(size_t)n_out * ch + at] = lo;
}

// UMMA N of the W ring for a given output width
static inline int proj_ns(int n_out) { return n_out % 128 == 0 ? 128 : 64; }

}  // namespace dfgnn

using namespace dfgnn;

extern "C" {

size_t dfgnn_proj_weight_image_floats(int n_out, int k) { return (size_t)2 * n_out * k; }

int dfgnn_proj_pack_weights(int n_out, int k, const float* W, float* w_img, void* stream) {
  const char* fn = "dfgnn_proj_pack_weights";
  if (n_out < 64 || n_out % 64 != 0 || (k != 32 && k != 64 && k != 128)) {
    set_error("%s: n_out=%d must be a multiple of 64 and k=%d one of 32, 64, 128", fn, n_out, k);
    return DFGNN_ERR_UNSUPPORTED_DIM;
  }
  DFGNN_REQUIRE(W, fn); DFGNN_REQUIRE(w_img, fn);
  const int n4 = n_out * (k / 4);
  proj_pack_kernel<<<(n4 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n_out, k, proj_ns(n_out), W, w_img);
  return check_launch(fn);
}

int dfgnn_proj_forward(int n, int k, int n_out, int part_width, const float* x, const float* w_img,
                       const float* bias, const float* scale, float* out0, float* out1, float* out2, float* out3,
                       int head_dim, const float* a_l, const float* a_r, float* attn_row, float* attn_col,
                       void* stream) {
  const char* fn = "dfgnn_proj_forward";
  if (n < 0 || n_out < 64 || n_out > kProjMaxOut || n_out % 64 != 0 || part_width < 64 || part_width % 64 != 0 || n_out % part_width != 0 ||
      n_out / part_width > 4) {
    set_error("%s: n_out=%d (<= %d) / part_width=%d must be multiples of 64 with at most 4 parts", fn, n_out,
              kProjMaxOut, part_width);
    return DFGNN_ERR_UNSUPPORTED_DIM;
  }
  if (k != 32 && k != 64 && k != 128) {
    set_error("%s: input width k=%d is not supported (32, 64, 128)", fn, k);
    return DFGNN_ERR_UNSUPPORTED_DIM;
  }
  if (n == 0) return DFGNN_OK;
  DFGNN_REQUIRE(x, fn); DFGNN_REQUIRE(w_img, fn); DFGNN_REQUIRE(out0, fn);
  float* outs[4] = {out0, out1, out2, out3};
  for (int i = 0; i < n_out / part_width; ++i)
    if (outs[i] == nullptr) { set_error("%s: output %d is NULL", fn, i); return DFGNN_ERR_INVALID_ARGUMENT; }
  if (head_dim > 0) {
    DFGNN_REQUIRE(a_l, fn); DFGNN_REQUIRE(a_r, fn); DFGNN_REQUIRE(attn_row, fn); DFGNN_REQUIRE(attn_col, fn);
    if (n_out % head_dim != 0 || !(head_dim == 8 || head_dim == 16 || head_dim % 32 == 0)) {
      set_error("%s: head_dim=%d must be 8, 16 or a multiple of 32", fn, head_dim);
      return DFGNN_ERR_UNSUPPORTED_DIM;
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  ProjParams p{n, k, n_out, part_width, x, w_img, bias, scale, out0, out1, out2, out3,
               head_dim, a_l, a_r, attn_row, attn_col};
  if (head_dim > 32) {  // wider heads are summed from several 32-column pieces
    const size_t bytes = (size_t)n * (n_out / head_dim) * sizeof(float);
    cudaMemsetAsync(attn_row, 0, bytes, st);
    cudaMemsetAsync(attn_col, 0, bytes, st);
  }
  static const int sms = [] {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v > 0 ? v : 148;
  }();
  const int tiles = (n + kProjM - 1) / kProjM;
  const int grid = tiles < sms ? tiles : sms;
  const int ns = proj_ns(n_out);
  const size_t smem = (size_t)2 * kProjM * k * 4 + (size_t)4 * ns * kStageK * 4 + (size_t)2 * kProjMaxOut * 4 +
                      (size_t)2 * kProjM * kStageLd * 4;
  auto launch = [&](auto kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kernel<<<grid, kProjThreads2, smem, st>>>(p);
  };
  if (ns == 128) {
    if (k == 128) launch(proj_tf32x3_kernel<128, 128>);
    else if (k == 64) launch(proj_tf32x3_kernel<64, 128>);
    else launch(proj_tf32x3_kernel<32, 128>);
  } else {
    if (k == 128) launch(proj_tf32x3_kernel<128, 64>);
    else if (k == 64) launch(proj_tf32x3_kernel<64, 64>);
    else launch(proj_tf32x3_kernel<32, 64>);
  }
  return check_launch(fn);
}

}  // extern "C"
