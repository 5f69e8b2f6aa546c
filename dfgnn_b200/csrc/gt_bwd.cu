// gt_bwd.cu -- GT / AGNN backward entry point of include/dfgnn_b200.h.
#include "abi_common.h"
#include "bwd_kernels.cuh"

using namespace dfgnn;

// col0 / n_sub / nnz_sub: the column side runs on columns [col0, col0 + n_sub) only (n_sub < 0:
// all n columns); nnz_sub = entries of those columns (schedule heuristics only).
static int gt_backward_impl(int phases, int col0, int n_sub, int nnz_sub, int m, int n, int nnz, int h, int f, const int32_t* row_ptr,
                                 const int32_t* col_ind, const int32_t* /*rows*/,
                                 const float* val, const int32_t* col_ptr,
                                 const int32_t* row_ind, const int32_t* val_idx,
                                 int /*smem_consume*/, const float* Q, const float* K,
                                 const float* V, const float* attn_edge, const float* grad_out,
                                 float* grad_Q, float* grad_K, float* grad_V, float* grad_edge,
                                 void* stream) {
  const char* fn = "dfgnn_gt_backward";
  if (phases < 1 || phases > 3) { set_error("%s: phases=%d must be 1, 2 or 3", fn, phases); return DFGNN_ERR_INVALID_ARGUMENT; }
  if (int rc = check_common(fn, m, nnz, h, f)) return rc;
  if (n < 0) { set_error("%s: invalid n=%d", fn, n); return DFGNN_ERR_INVALID_ARGUMENT; }
  if (m == 0 && n == 0) return DFGNN_OK;  // an empty graph: nothing to read or write
  DFGNN_REQUIRE(row_ptr, fn); DFGNN_REQUIRE(col_ptr, fn);
  if (nnz > 0) {
    DFGNN_REQUIRE(col_ind, fn); DFGNN_REQUIRE(row_ind, fn); DFGNN_REQUIRE(val_idx, fn);
    DFGNN_REQUIRE(attn_edge, fn); DFGNN_REQUIRE(grad_edge, fn);
  }
  DFGNN_REQUIRE(Q, fn); DFGNN_REQUIRE(K, fn); DFGNN_REQUIRE(V, fn); DFGNN_REQUIRE(grad_out, fn);
  DFGNN_REQUIRE(grad_Q, fn); DFGNN_REQUIRE(grad_K, fn); DFGNN_REQUIRE(grad_V, fn);
  if (n_sub >= 0 && (col0 < 0 || col0 + n_sub > n)) {
    set_error("%s: column range [%d, %d) outside the %d columns", fn, col0, col0 + n_sub, n);
    return DFGNN_ERR_INVALID_ARGUMENT;
  }
  if (m == 0 && n == 0) return DFGNN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  GtBwdParams p{m, n, nnz, h, f, 8, 8, row_ptr, col_ind, col_ptr, row_ind, val_idx,
                Q, K, V, attn_edge, val, grad_out, grad_Q, grad_K, grad_V, grad_edge};
  int n_c = n, nnz_c = nnz;
  if (n_sub >= 0) {  // column side restricted to a column range (row-partitioned shards, dist.py)
    n_c = n_sub;
    nnz_c = nnz_sub >= 0 ? nnz_sub : nnz;
    p.n = n_sub;
    p.col_ptr = col_ptr + col0;
    p.dK = grad_K + (size_t)col0 * h * f;
    p.dV = grad_V + (size_t)col0 * h * f;
  }
  int rc = DFGNN_OK;
  dispatch_layout(f, [&](auto tag) {
    using L = typename decltype(tag)::type;
    constexpr int C = ChunkOf<L>::C;
    p.rb = pick_rb(m, nnz, L::G);
    p.rb_col = pick_rb(n_c, nnz_c, L::G);
    const dim3 grid((m + p.rb - 1) / p.rb, h);
    const dim3 grid_c((n_c + p.rb_col - 1) / p.rb_col, h);
    const size_t smem = slot_bytes<2 * L::NR, L>();
    ensure_smem(gt_bwd_row_kernel<L, C>, smem);
    ensure_smem(gt_bwd_col_kernel<L, C>, smem);
    if (m > 0 && (phases & 1)) {
      gt_bwd_row_kernel<L, C><<<grid, kNW * 32, smem, st>>>(p);
      rc = check_launch(fn);
      note_kernel(1, "gt_bwd_row_kernel");
      if (rc) return;
    }
    if (n_c > 0 && (phases & 2)) {
      gt_bwd_col_kernel<L, C><<<grid_c, kNW * 32, smem, st>>>(p);
      rc = check_launch(fn);
      note_kernel(2, "gt_bwd_col_kernel");
    }
  }, long_rows(m, nnz));
  return rc;
}

extern "C" {

int dfgnn_gt_backward(int m, int n, int nnz, int h, int f, const int32_t* row_ptr,
                      const int32_t* col_ind, const int32_t* rows, const float* val,
                      const int32_t* col_ptr, const int32_t* row_ind, const int32_t* val_idx,
                      int smem_consume, const float* Q, const float* K, const float* V,
                      const float* attn_edge, const float* grad_out, float* grad_Q, float* grad_K,
                      float* grad_V, float* grad_edge, void* stream) {
  return gt_backward_impl(3, 0, -1, -1, m, n, nnz, h, f, row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx,
                          smem_consume, Q, K, V, attn_edge, grad_out, grad_Q, grad_K, grad_V,
                          grad_edge, stream);
}

int dfgnn_gt_backward_phase(int phases, int m, int n, int nnz, int h, int f, const int32_t* row_ptr,
                            const int32_t* col_ind, const int32_t* rows, const float* val,
                            const int32_t* col_ptr, const int32_t* row_ind, const int32_t* val_idx,
                            int smem_consume, const float* Q, const float* K, const float* V,
                            const float* attn_edge, const float* grad_out, float* grad_Q,
                            float* grad_K, float* grad_V, float* grad_edge, void* stream) {
  return gt_backward_impl(phases, 0, -1, -1, m, n, nnz, h, f, row_ptr, col_ind, rows, val, col_ptr, row_ind,
                          val_idx, smem_consume, Q, K, V, attn_edge, grad_out, grad_Q, grad_K, grad_V,
                          grad_edge, stream);
}

int dfgnn_gt_backward_cols(int col_begin, int n_sub, int nnz_sub, int m, int n, int nnz, int h, int f,
                           const int32_t* row_ptr, const int32_t* col_ind, const int32_t* rows,
                           const float* val, const int32_t* col_ptr, const int32_t* row_ind,
                           const int32_t* val_idx, int smem_consume, const float* Q, const float* K,
                           const float* V, const float* attn_edge, const float* grad_out,
                           float* grad_Q, float* grad_K, float* grad_V, float* grad_edge,
                           void* stream) {
  if (n_sub < 0) { set_error("dfgnn_gt_backward_cols: n_sub=%d must be >= 0", n_sub); return DFGNN_ERR_INVALID_ARGUMENT; }
  return gt_backward_impl(2, col_begin, n_sub, nnz_sub, m, n, nnz, h, f, row_ptr, col_ind, rows, val,
                          col_ptr, row_ind, val_idx, smem_consume, Q, K, V, attn_edge, grad_out,
                          grad_Q, grad_K, grad_V, grad_edge, stream);
}

}  // extern "C"
