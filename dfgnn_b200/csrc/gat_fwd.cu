// gat_fwd.cu -- GAT forward / inference entry points of include/dfgnn_b200.h.
#include "abi_common.h"
#include "fwd_kernels.cuh"
#include "staged_gat.cuh"

namespace dfgnn {

int launch_gat_fwd(int m, int nnz, int h, int f, const float* ar, const float* ac,
                   const int* row_ptr, const int* col_ind, float slope, const float* feat,
                   float drop, uint64_t seed, float* out, float* emax, float* esum, float* emask,
                   cudaStream_t st, const char* fn) {
  if (int rc = check_common(fn, m, nnz, h, f)) return rc;
  if (m == 0) return DFGNN_OK;  // an empty graph: nothing to read or write
  DFGNN_REQUIRE(ar, fn); DFGNN_REQUIRE(ac, fn); DFGNN_REQUIRE(row_ptr, fn);
  if (nnz > 0) DFGNN_REQUIRE(col_ind, fn);
  DFGNN_REQUIRE(feat, fn); DFGNN_REQUIRE(out, fn);
  if (!(drop >= 0.f && drop < 1.f)) {
    set_error("%s: attn_drop=%g must be in [0, 1)", fn, (double)drop);
    return DFGNN_ERR_INVALID_ARGUMENT;
  }
  if ((emax == nullptr) != (esum == nullptr)) {
    set_error("%s: edge_max and edge_sum must be given together", fn);
    return DFGNN_ERR_INVALID_ARGUMENT;
  }
  if (m == 0) return DFGNN_OK;
  GatFwdParams p{m, nnz, h, f, 8, row_ptr, col_ind, ar, ac, feat,
                 slope, drop, seed, out, emax, esum, emask};
  int rc = DFGNN_OK;
  dispatch_layout(f, [&](auto tag) {
    using L = typename decltype(tag)::type;
    constexpr int C = ChunkOf<L>::C1;
    const bool staged = want_staged(m, nnz);
    note_kernel(0, staged ? "gat_fwd_staged_kernel" : "gat_fwd_kernel");
    p.rb = staged ? pick_rb_staged(m, nnz) : pick_rb(m, nnz, L::G);
    const dim3 grid((m + p.rb - 1) / p.rb, h);
    const size_t smem = slot_bytes<L::NR, L>();
    ensure_smem(gat_fwd_kernel<L, C>, smem);
    if (staged) {
      ensure_smem(gat_fwd_staged_kernel<L, StageChunk<L>::kSpmm>, smem, 24 * 1024);
      gat_fwd_staged_kernel<L, StageChunk<L>::kSpmm><<<grid, kNW * 32, smem, st>>>(p);
      rc = check_launch(fn);
      if (rc) return;
      p.cap = kStageCap;  // tiles the staged kernel skipped
      launch_overlapped(gat_fwd_kernel<L, C>, grid, dim3(kNW * 32), smem, st, p);
    } else {
      gat_fwd_kernel<L, C><<<grid, kNW * 32, smem, st>>>(p);
    }
    rc = check_launch(fn);
  });
  return rc;
}

static int attn_weight(const char* fn, int m, int h, int f, const float* a_l, const float* a_r,
                       const float* feat, float* ar, float* ac, cudaStream_t st) {
  DFGNN_REQUIRE(a_l, fn); DFGNN_REQUIRE(a_r, fn); DFGNN_REQUIRE(feat, fn);
  DFGNN_REQUIRE(ar, fn); DFGNN_REQUIRE(ac, fn);
  if (m == 0) return DFGNN_OK;
  const long long warps = (long long)m * h;
  gat_attn_weight_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(m, h, f, a_l, a_r, feat, ar, ac);
  return check_launch(fn);
}

}  // namespace dfgnn

using namespace dfgnn;

#define GAT_INFER(fn)                                                                      \
  return launch_gat_fwd(m, nnz, h, f, attn_row, attn_col, indptr, indices, negative_slope, \
                        in_feat, 0.f, 0, out_feat, nullptr, nullptr, nullptr,              \
                        (cudaStream_t)stream, fn)

extern "C" {

int dfgnn_gat_forward(int m, int nnz, int h, int f, const float* attn_row, const float* attn_col,
                      const int32_t* row_ptr, const int32_t* col_ind, float negative_slope,
                      const float* in_feat, float attn_drop, uint64_t seed, float* out_feat,
                      float* edge_max, float* edge_sum, float* edge_mask, void* stream) {
  const char* fn = "dfgnn_gat_forward";
  if (m == 0) return check_common(fn, m, nnz, h, f);
  DFGNN_REQUIRE(edge_max, fn); DFGNN_REQUIRE(edge_sum, fn);
  if (attn_drop > 0.f && nnz > 0) DFGNN_REQUIRE(edge_mask, fn);
  return launch_gat_fwd(m, nnz, h, f, attn_row, attn_col, row_ptr, col_ind, negative_slope,
                        in_feat, attn_drop, seed, out_feat, edge_max, edge_sum,
                        nnz > 0 ? edge_mask : nullptr, (cudaStream_t)stream, fn);
}

int dfgnn_gat_inference(int m, int nnz, int h, int f, const float* attn_row, const float* attn_col,
                        const int32_t* indptr, const int32_t* indices, float negative_slope,
                        const float* in_feat, float* out_feat, void* stream) {
  GAT_INFER("dfgnn_gat_inference");
}
int dfgnn_gat_inference_hyper(int, int m, int nnz, int h, int f, const float* attn_row,
                              const float* attn_col, const int32_t* indptr, const int32_t* indices,
                              const int32_t*, float negative_slope, const float* in_feat,
                              float* out_feat, void* stream) {
  GAT_INFER("dfgnn_gat_inference_hyper");
}
int dfgnn_gat_inference_hyper_recompute(int m, int nnz, int h, int f, const float* attn_row,
                                        const float* attn_col, const int32_t* indptr,
                                        const int32_t* indices, float negative_slope,
                                        const float* in_feat, float* out_feat, void* stream) {
  GAT_INFER("dfgnn_gat_inference_hyper_recompute");
}
int dfgnn_gat_inference_softmax(int, int m, int nnz, int h, int f, const float* attn_row,
                                const float* attn_col, const int32_t* indptr,
                                const int32_t* indices, const int32_t*, float negative_slope,
                                const float* in_feat, float* out_feat, void* stream) {
  GAT_INFER("dfgnn_gat_inference_softmax");
}
int dfgnn_gat_inference_softmax_gm(int m, int nnz, int h, int f, const float* attn_row,
                                   const float* attn_col, const int32_t* indptr,
                                   const int32_t* indices, const int32_t*, float negative_slope,
                                   const float* in_feat, float* out_feat, void* stream) {
  GAT_INFER("dfgnn_gat_inference_softmax_gm");
}
int dfgnn_gat_inference_tiling(int m, int nnz, int h, int f, const float* attn_row,
                               const float* attn_col, const int32_t* indptr, const int32_t* indices,
                               float negative_slope, const float* in_feat, float* out_feat,
                               void* stream) {
  GAT_INFER("dfgnn_gat_inference_tiling");
}

int dfgnn_gat_attn_weight(int m, int h, int f, const float* a_l, const float* a_r,
                          const float* in_feat, float* attn_row, float* attn_col, void* stream) {
  const char* fn = "dfgnn_gat_attn_weight";
  if (int rc = check_common(fn, m, 0, h, f)) return rc;
  return attn_weight(fn, m, h, f, a_l, a_r, in_feat, attn_row, attn_col, (cudaStream_t)stream);
}

int dfgnn_gat_inference_hyper_v2(int, int m, int nnz, int h, int f, const float* a_l,
                                 const float* a_r, const int32_t* indptr, const int32_t* indices,
                                 float negative_slope, const float* in_feat, float* attn_row,
                                 float* attn_col, float* out_feat, void* stream) {
  const char* fn = "dfgnn_gat_inference_hyper_v2";
  if (int rc = check_common(fn, m, nnz, h, f)) return rc;
  if (int rc = attn_weight(fn, m, h, f, a_l, a_r, in_feat, attn_row, attn_col, (cudaStream_t)stream))
    return rc;
  GAT_INFER(fn);
}

}  // extern "C"
