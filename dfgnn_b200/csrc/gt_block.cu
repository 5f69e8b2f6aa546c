// gt_block.cu -- entry points of the graph-resident GT kernels (block_gt.cuh) for block-diagonal
// batches of small graphs, and the block-plan validation.
#include "abi_common.h"
#include "block_gt.cuh"
#include "dense_gt.cuh"

namespace dfgnn {

// Mode (dfgnn_set_block_mode; initial value from DFGNN_B200_BLOCK=0|1|2|3): 0 auto, 1 off, 2 the
// shared-memory-staged sparse kernels (block_gt.cuh) whenever they fit, 3 the dense mma.sync
// kernels (dense_gt.cuh) whenever they fit, 4 the dense tcgen05 kernels (dense_tc.cu).  Auto: dense kernels for dense batches (the caller
// checks the density), never the staged sparse ones (measured slower than the row-block kernels on
// the PATTERN-shaped batch: both are instruction-issue bound).
static std::atomic<int>& block_mode() {
  static std::atomic<int> v{[] {
    const char* e = getenv("DFGNN_B200_BLOCK");
    return e ? (e[0] == '0' ? 1 : (e[0] == '2' ? 3 : (e[0] == '3' ? 4 : 2))) : 0;
  }()};
  return v;
}
static int block_override() { return block_mode().load(std::memory_order_relaxed); }

static size_t smem_limit() {
  static const size_t lim = [] {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    return (size_t)(v > 0 ? v : 48 * 1024);
  }();
  return lim;
}

template <class Fn>
static bool dispatch_block_layout(int f, Fn&& fn) {
  if (f == 32) fn(Tag<VecLayout<8, 8>>{});
  else if (f == 64) fn(Tag<VecLayout<16, 8>>{});
  else if (f == 128) fn(Tag<VecLayout<32, 16>>{});
  else return false;
  return true;
}

// warps per CTA for a stage of max_nodes rows: 16 if the slots still fit next to the two operand
// blocks, else 8, else 0 (not supported).  NVF = slot floats per lane in units of NR (1 fwd, 2 bwd).
template <class L, int NVF>
static int pick_nw(int max_nodes, int f) {
  const size_t lim = smem_limit() - 64;  // static: the mbarrier
  if (block_smem_bytes<NVF * L::NR, L, 16>(max_nodes, f) <= lim) return 16;
  if (block_smem_bytes<NVF * L::NR, L, 8>(max_nodes, f) <= lim) return 8;
  return 0;
}

static bool block_supported(int max_nodes, int m, int nnz, int h, int f) {
  if (block_override() != 2 || h != 1 || m <= 0 || nnz <= 0 || max_nodes <= 0) return false;
  bool fits = false;
  if (!dispatch_block_layout(f, [&](auto tag) {
        using L = typename decltype(tag)::type;
        fits = pick_nw<L, 1>(max_nodes, f) > 0 && pick_nw<L, 2>(max_nodes, f) > 0;
      }))
    return false;
  return fits;
}

// dense tensor-core kernels: f in {64, 128}, unweighted scores, stage fits
static bool dense_supported(int max_nodes, int h, int f) {
  if (h != 1 || (f != 64 && f != 128) || max_nodes < 1) return false;
  return DenseSmem(max_nodes, f).bytes <= smem_limit() - 64;
}

template <class K, class P>
static void launch_block(K kernel, int n_blocks, int nw, size_t smem, cudaStream_t st, const P& p) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kernel<<<n_blocks, nw * 32, smem, st>>>(p);
}

}  // namespace dfgnn

using namespace dfgnn;

extern "C" {

int dfgnn_block_plan_check(int n_blocks, int m, int nnz, const int32_t* blk_ptr, const int32_t* row_ptr,
                           const int32_t* col_ind, int32_t* flag_ws, int32_t* max_nodes_out, int32_t* ascending_out,
                           void* stream) {
  const char* fn = "dfgnn_block_plan_check";
  if (n_blocks < 1 || m < 0 || nnz < 0) { set_error("%s: invalid sizes", fn); return DFGNN_ERR_INVALID_ARGUMENT; }
  DFGNN_REQUIRE(blk_ptr, fn); DFGNN_REQUIRE(row_ptr, fn); DFGNN_REQUIRE(flag_ws, fn); DFGNN_REQUIRE(max_nodes_out, fn);
  if (nnz > 0) DFGNN_REQUIRE(col_ind, fn);
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(flag_ws, 0, 3 * sizeof(int32_t), st);
  block_check_kernel<<<n_blocks, 256, 0, st>>>(n_blocks, m, blk_ptr, row_ptr, col_ind, flag_ws);
  if (int rc = check_launch(fn)) return rc;
  int h_flag[3] = {0, 0, 0};
  cudaError_t err = cudaMemcpyAsync(h_flag, flag_ws, 3 * sizeof(int), cudaMemcpyDeviceToHost, st);
  if (err == cudaSuccess) err = cudaStreamSynchronize(st);
  if (err != cudaSuccess) { set_error("%s: %s", fn, cudaGetErrorString(err)); return (int)err; }
  if (h_flag[0]) {
    set_error("%s: the matrix is not block diagonal over the given node ranges", fn);
    return DFGNN_ERR_INVALID_ARGUMENT;
  }
  *max_nodes_out = h_flag[1];
  if (ascending_out) *ascending_out = h_flag[2] ? 0 : 1;
  return DFGNN_OK;
}

int dfgnn_set_block_mode(int mode) {
  if (mode < 0 || mode > 4) return block_override();
  return block_mode().exchange(mode);
}

int dfgnn_gt_block_supported(int max_nodes, int m, int nnz, int h, int f) {
  return block_supported(max_nodes, m, nnz, h, f) ? 1 : 0;
}

int dfgnn_gt_dense_supported(int max_nodes, int h, int f) {
  const int mode = block_override();
  return ((mode == 0 || mode == 3) && dense_supported(max_nodes, h, f)) ? 1 : 0;
}

int dfgnn_gt_dense_forward(int n_blocks, const int32_t* blk_ptr, int max_nodes, int m, int nnz, int h, int f,
                           const int32_t* row_ptr, const int32_t* col_ind, const float* Q, const float* K,
                           const float* V, float* out_feat, float* attn_edge, void* stream) {
  const char* fn = "dfgnn_gt_dense_forward";
  if (int rc = check_common(fn, m, nnz, h, f)) return rc;
  if (m == 0) return DFGNN_OK;
  DFGNN_REQUIRE(blk_ptr, fn); DFGNN_REQUIRE(row_ptr, fn);
  if (nnz > 0) DFGNN_REQUIRE(col_ind, fn);
  DFGNN_REQUIRE(Q, fn); DFGNN_REQUIRE(K, fn); DFGNN_REQUIRE(V, fn); DFGNN_REQUIRE(out_feat, fn);
  if (n_blocks < 1 || !dense_supported(max_nodes, h, f)) {
    set_error("%s: needs h == 1, f in {64, 128} and a stage that fits (h=%d, f=%d, max_nodes=%d)", fn, h, f, max_nodes);
    return DFGNN_ERR_UNSUPPORTED_DIM;
  }
  cudaStream_t st = (cudaStream_t)stream;
  GtBlockFwdParams p{{m, nnz, h, f, 0, row_ptr, col_ind, nullptr, Q, K, V, nullptr, out_feat, nnz > 0 ? attn_edge : nullptr},
                     {blk_ptr, n_blocks, max_nodes}};
  const size_t smem = DenseSmem(max_nodes, f).bytes;
#ifndef DFGNN_DENSE_NW
#define DFGNN_DENSE_NW 8
#endif
  if (f == 128) launch_block(gt_dense_fwd_kernel<128, DFGNN_DENSE_NW>, n_blocks, DFGNN_DENSE_NW, smem, st, p);
  else launch_block(gt_dense_fwd_kernel<64, DFGNN_DENSE_NW>, n_blocks, DFGNN_DENSE_NW, smem, st, p);
  note_kernel(0, "gt_dense_fwd_kernel");
  return check_launch(fn);
}

int dfgnn_gt_block_forward(int n_blocks, const int32_t* blk_ptr, int max_nodes, int m, int nnz, int h, int f,
                           const int32_t* row_ptr, const int32_t* col_ind, const float* val, const float* Q,
                           const float* K, const float* V, float* out_feat, float* attn_edge, void* stream) {
  const char* fn = "dfgnn_gt_block_forward";
  if (int rc = check_common(fn, m, nnz, h, f)) return rc;
  if (m == 0) return DFGNN_OK;
  DFGNN_REQUIRE(blk_ptr, fn); DFGNN_REQUIRE(row_ptr, fn);
  if (nnz > 0) DFGNN_REQUIRE(col_ind, fn);
  DFGNN_REQUIRE(Q, fn); DFGNN_REQUIRE(K, fn); DFGNN_REQUIRE(V, fn); DFGNN_REQUIRE(out_feat, fn);
  if (n_blocks < 1 || h != 1 || max_nodes < 1) {
    set_error("%s: needs h == 1 and a block plan (h=%d, n_blocks=%d, max_nodes=%d)", fn, h, n_blocks, max_nodes);
    return DFGNN_ERR_INVALID_ARGUMENT;
  }
  cudaStream_t st = (cudaStream_t)stream;
  GtBlockFwdParams p{{m, nnz, h, f, 0, row_ptr, col_ind, val, Q, K, V, nullptr, out_feat, nnz > 0 ? attn_edge : nullptr},
                     {blk_ptr, n_blocks, max_nodes}};
  int rc = DFGNN_ERR_UNSUPPORTED_DIM;
  const bool ok = dispatch_block_layout(f, [&](auto tag) {
    using L = typename decltype(tag)::type;
    const int nw = pick_nw<L, 1>(max_nodes, f);
    if (nw == 16) launch_block(gt_block_fwd_kernel<L, 2, 16>, n_blocks, 16, block_smem_bytes<L::NR, L, 16>(max_nodes, f), st, p);
    else if (nw == 8) launch_block(gt_block_fwd_kernel<L, 4, 8>, n_blocks, 8, block_smem_bytes<L::NR, L, 8>(max_nodes, f), st, p);
    else return;
    rc = check_launch(fn);
    note_kernel(0, "gt_block_fwd_kernel");
  });
  if (!ok || rc == DFGNN_ERR_UNSUPPORTED_DIM) {
    set_error("%s: f=%d / max_nodes=%d does not fit the shared-memory stage", fn, f, max_nodes);
    return DFGNN_ERR_UNSUPPORTED_DIM;
  }
  return rc;
}

int dfgnn_gt_block_backward(int phases, int n_blocks, const int32_t* blk_ptr, int max_nodes, int m, int nnz,
                            int h, int f, const int32_t* row_ptr, const int32_t* col_ind, const float* val,
                            const int32_t* col_ptr, const int32_t* row_ind, const int32_t* val_idx,
                            const float* Q, const float* K, const float* V, const float* attn_edge,
                            const float* grad_out, float* grad_Q, float* grad_K, float* grad_V,
                            float* grad_edge, void* stream) {
  const char* fn = "dfgnn_gt_block_backward";
  if (phases < 1 || phases > 3) { set_error("%s: phases=%d must be 1, 2 or 3", fn, phases); return DFGNN_ERR_INVALID_ARGUMENT; }
  if (int rc = check_common(fn, m, nnz, h, f)) return rc;
  if (m == 0) return DFGNN_OK;
  DFGNN_REQUIRE(blk_ptr, fn); DFGNN_REQUIRE(row_ptr, fn); DFGNN_REQUIRE(col_ptr, fn);
  if (nnz > 0) {
    DFGNN_REQUIRE(col_ind, fn); DFGNN_REQUIRE(row_ind, fn); DFGNN_REQUIRE(val_idx, fn);
    DFGNN_REQUIRE(attn_edge, fn); DFGNN_REQUIRE(grad_edge, fn);
  }
  DFGNN_REQUIRE(Q, fn); DFGNN_REQUIRE(K, fn); DFGNN_REQUIRE(V, fn); DFGNN_REQUIRE(grad_out, fn);
  DFGNN_REQUIRE(grad_Q, fn); DFGNN_REQUIRE(grad_K, fn); DFGNN_REQUIRE(grad_V, fn);
  if (n_blocks < 1 || h != 1 || max_nodes < 1) {
    set_error("%s: needs h == 1 and a block plan (h=%d, n_blocks=%d, max_nodes=%d)", fn, h, n_blocks, max_nodes);
    return DFGNN_ERR_INVALID_ARGUMENT;
  }
  cudaStream_t st = (cudaStream_t)stream;
  GtBlockBwdParams p{{m, m, nnz, h, f, 0, 0, row_ptr, col_ind, col_ptr, row_ind, val_idx, Q, K, V, attn_edge, val,
                      grad_out, grad_Q, grad_K, grad_V, grad_edge},
                     {blk_ptr, n_blocks, max_nodes}};
  int rc = DFGNN_ERR_UNSUPPORTED_DIM;
  const bool ok = dispatch_block_layout(f, [&](auto tag) {
    using L = typename decltype(tag)::type;
    const int nw = pick_nw<L, 2>(max_nodes, f);
    if (nw == 0) return;
    const size_t smem = nw == 16 ? block_smem_bytes<2 * L::NR, L, 16>(max_nodes, f)
                                 : block_smem_bytes<2 * L::NR, L, 8>(max_nodes, f);
    rc = DFGNN_OK;
    if (phases & 1) {
      if (nw == 16) launch_block(gt_block_bwd_row_kernel<L, 2, 16>, n_blocks, 16, smem, st, p);
      else launch_block(gt_block_bwd_row_kernel<L, 4, 8>, n_blocks, 8, smem, st, p);
      rc = check_launch(fn);
      note_kernel(1, "gt_block_bwd_row_kernel");
      if (rc) return;
    }
    if (phases & 2) {
      if (nw == 16) launch_block(gt_block_bwd_col_kernel<L, 2, 16>, n_blocks, 16, smem, st, p);
      else launch_block(gt_block_bwd_col_kernel<L, 4, 8>, n_blocks, 8, smem, st, p);
      rc = check_launch(fn);
      note_kernel(2, "gt_block_bwd_col_kernel");
    }
  });
  if (!ok || rc == DFGNN_ERR_UNSUPPORTED_DIM) {
    set_error("%s: f=%d / max_nodes=%d does not fit the shared-memory stage", fn, f, max_nodes);
    return DFGNN_ERR_UNSUPPORTED_DIM;
  }
  return rc;
}

}  // extern "C"
