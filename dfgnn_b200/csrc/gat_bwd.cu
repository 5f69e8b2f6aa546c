// gat_bwd.cu -- GAT backward entry point of include/dfgnn_b200.h.
#include "abi_common.h"
#include "bwd_kernels.cuh"
#include "staged_gat.cuh"

using namespace dfgnn;

static int gat_backward_impl(int phases, int col0, int n_sub, int nnz_sub, int m, int n, int nnz, int h, int f, float negative_slope,
                                  float attn_drop, const int32_t* row_ptr, const int32_t* col_ind,
                                  const int32_t* col_ptr, const int32_t* row_ind,
                                  const int32_t* permute, const float* edge_max,
                                  const float* edge_sum, const float* edge_mask,
                                  const float* in_feat, const float* attn_row,
                                  const float* attn_col, const float* grad_out, float* grad_feat,
                                  float* grad_attn_row, float* grad_attn_col, float* grad_edge,
                                  void* stream) {
  const char* fn = "dfgnn_gat_backward";
  if (phases < 1 || phases > 3) { set_error("%s: phases=%d must be 1, 2 or 3", fn, phases); return DFGNN_ERR_INVALID_ARGUMENT; }
  if (int rc = check_common(fn, m, nnz, h, f)) return rc;
  if (n < 0) { set_error("%s: invalid n=%d", fn, n); return DFGNN_ERR_INVALID_ARGUMENT; }
  if (m == 0 && n == 0) return DFGNN_OK;  // an empty graph: nothing to read or write
  DFGNN_REQUIRE(row_ptr, fn); DFGNN_REQUIRE(col_ptr, fn);
  if (nnz > 0) {
    DFGNN_REQUIRE(col_ind, fn); DFGNN_REQUIRE(row_ind, fn); DFGNN_REQUIRE(permute, fn);
    DFGNN_REQUIRE(grad_edge, fn);
    if (attn_drop > 0.f) DFGNN_REQUIRE(edge_mask, fn);
  }
  DFGNN_REQUIRE(edge_max, fn); DFGNN_REQUIRE(edge_sum, fn); DFGNN_REQUIRE(in_feat, fn);
  DFGNN_REQUIRE(attn_row, fn); DFGNN_REQUIRE(attn_col, fn); DFGNN_REQUIRE(grad_out, fn);
  DFGNN_REQUIRE(grad_feat, fn); DFGNN_REQUIRE(grad_attn_row, fn); DFGNN_REQUIRE(grad_attn_col, fn);
  if (!(attn_drop >= 0.f && attn_drop < 1.f)) {
    set_error("%s: attn_drop=%g must be in [0, 1)", fn, (double)attn_drop);
    return DFGNN_ERR_INVALID_ARGUMENT;
  }
  if (n_sub >= 0 && (col0 < 0 || col0 + n_sub > n)) {
    set_error("%s: column range [%d, %d) outside the %d columns", fn, col0, col0 + n_sub, n);
    return DFGNN_ERR_INVALID_ARGUMENT;
  }
  if (m == 0 && n == 0) return DFGNN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // with attn_drop == 0 every edge is kept: skip the mask reads entirely
  const float* mask = attn_drop > 0.f ? edge_mask : nullptr;
  GatBwdParams p{m, n, nnz, h, f, 8, 8, negative_slope, attn_drop, row_ptr, col_ind,
                 col_ptr, row_ind, permute, edge_max, edge_sum, mask, in_feat, attn_row,
                 attn_col, grad_out, grad_feat, grad_attn_row, grad_attn_col, grad_edge};
  int n_c = n, nnz_c = nnz;
  if (n_sub >= 0) {  // column side restricted to a column range (row-partitioned shards, dist.py)
    n_c = n_sub;
    nnz_c = nnz_sub >= 0 ? nnz_sub : nnz;
    p.n = n_sub;
    p.col_ptr = col_ptr + col0;
    p.grad_feat = grad_feat + (size_t)col0 * h * f;
    p.grad_ac = grad_attn_col + (size_t)col0 * h;
  }
  int rc = DFGNN_OK;
  dispatch_layout(f, [&](auto tag) {
    using L = typename decltype(tag)::type;
    constexpr int C = ChunkOf<L>::C1;
    const bool staged_r = want_staged(m, nnz), staged_c = want_staged(n_c, nnz_c);
    p.rb = staged_r ? pick_rb_staged(m, nnz) : pick_rb(m, nnz, L::G);
    p.rb_col = staged_c ? pick_rb_staged(n_c, nnz_c) : pick_rb(n_c, nnz_c, L::G);
    const dim3 grid((m + p.rb - 1) / p.rb, h);
    const dim3 grid_c((n_c + p.rb_col - 1) / p.rb_col, h);
    ensure_smem(gat_bwd_col_kernel<L, C>, slot_bytes<L::NR, L>());
    if (m > 0 && (phases & 1)) {
      note_kernel(1, staged_r ? "gat_bwd_row_staged_kernel" : "gat_bwd_row_kernel");
      if (staged_r) {
        const size_t sx = stage_x<L>() ? (size_t)p.rb * f * sizeof(float) : 0;
        ensure_smem(gat_bwd_row_staged_kernel<L, StageChunk<L>::kSddmm>, sx, 44 * 1024);
        gat_bwd_row_staged_kernel<L, StageChunk<L>::kSddmm><<<grid, kNW * 32, sx, st>>>(p);
        rc = check_launch(fn);
        if (rc) return;
        p.cap = kStageCap;
        launch_overlapped(gat_bwd_row_kernel<L, C>, grid, dim3(kNW * 32), slot_bytes<1, L>(), st, p);
      } else {
        gat_bwd_row_kernel<L, C><<<grid, kNW * 32, slot_bytes<1, L>(), st>>>(p);
      }
      rc = check_launch(fn);
      if (rc) return;
    }
    if (n_c > 0 && (phases & 2)) {
      p.cap = 0;
      note_kernel(2, staged_c ? "gat_bwd_col_staged_kernel" : "gat_bwd_col_kernel");
      if (staged_c) {
        ensure_smem(gat_bwd_col_staged_kernel<L, StageChunk<L>::kSpmm>, slot_bytes<L::NR, L>(), 42 * 1024);
        gat_bwd_col_staged_kernel<L, StageChunk<L>::kSpmm><<<grid_c, kNW * 32, slot_bytes<L::NR, L>(), st>>>(p);
        rc = check_launch(fn);
        if (rc) return;
        p.cap = kStageCap;
        launch_overlapped(gat_bwd_col_kernel<L, C>, grid_c, dim3(kNW * 32), slot_bytes<L::NR, L>(), st, p);
      } else {
        gat_bwd_col_kernel<L, C><<<grid_c, kNW * 32, slot_bytes<L::NR, L>(), st>>>(p);
      }
      rc = check_launch(fn);
    }
  });
  return rc;
}

extern "C" {

int dfgnn_gat_backward(int m, int n, int nnz, int h, int f, float negative_slope, float attn_drop,
                       const int32_t* row_ptr, const int32_t* col_ind, const int32_t* col_ptr,
                       const int32_t* row_ind, const int32_t* permute, const float* edge_max,
                       const float* edge_sum, const float* edge_mask, const float* in_feat,
                       const float* attn_row, const float* attn_col, const float* grad_out,
                       float* grad_feat, float* grad_attn_row, float* grad_attn_col, float* grad_edge,
                       void* stream) {
  return gat_backward_impl(3, 0, -1, -1, m, n, nnz, h, f, negative_slope, attn_drop, row_ptr, col_ind, col_ptr,
                           row_ind, permute, edge_max, edge_sum, edge_mask, in_feat, attn_row,
                           attn_col, grad_out, grad_feat, grad_attn_row, grad_attn_col, grad_edge,
                           stream);
}

int dfgnn_gat_backward_phase(int phases, int m, int n, int nnz, int h, int f, float negative_slope,
                             float attn_drop, const int32_t* row_ptr, const int32_t* col_ind,
                             const int32_t* col_ptr, const int32_t* row_ind, const int32_t* permute,
                             const float* edge_max, const float* edge_sum, const float* edge_mask,
                             const float* in_feat, const float* attn_row, const float* attn_col,
                             const float* grad_out, float* grad_feat, float* grad_attn_row,
                             float* grad_attn_col, float* grad_edge, void* stream) {
  return gat_backward_impl(phases, 0, -1, -1, m, n, nnz, h, f, negative_slope, attn_drop, row_ptr, col_ind,
                           col_ptr, row_ind, permute, edge_max, edge_sum, edge_mask, in_feat,
                           attn_row, attn_col, grad_out, grad_feat, grad_attn_row, grad_attn_col,
                           grad_edge, stream);
}

int dfgnn_gat_backward_cols(int col_begin, int n_sub, int nnz_sub, int m, int n, int nnz, int h, int f,
                            float negative_slope, float attn_drop, const int32_t* row_ptr,
                            const int32_t* col_ind, const int32_t* col_ptr, const int32_t* row_ind,
                            const int32_t* permute, const float* edge_max, const float* edge_sum,
                            const float* edge_mask, const float* in_feat, const float* attn_row,
                            const float* attn_col, const float* grad_out, float* grad_feat,
                            float* grad_attn_row, float* grad_attn_col, float* grad_edge,
                            void* stream) {
  if (n_sub < 0) { set_error("dfgnn_gat_backward_cols: n_sub=%d must be >= 0", n_sub); return DFGNN_ERR_INVALID_ARGUMENT; }
  return gat_backward_impl(2, col_begin, n_sub, nnz_sub, m, n, nnz, h, f, negative_slope, attn_drop,
                           row_ptr, col_ind, col_ptr, row_ind, permute, edge_max, edge_sum,
                           edge_mask, in_feat, attn_row, attn_col, grad_out, grad_feat,
                           grad_attn_row, grad_attn_col, grad_edge, stream);
}

}  // extern "C"
