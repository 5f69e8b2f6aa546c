// gt_fwd.cu -- GT / AGNN forward entry points of include/dfgnn_b200.h.
#include "abi_common.h"
#include "fwd_kernels.cuh"

namespace dfgnn {

int launch_dot_fwd(bool agnn, int m, int nnz, int h, int f, const int* row_ptr, const int* col_ind,
                   const float* val, const float* Q, const float* K, const float* V,
                   const float* rn, float* out, float* attn, cudaStream_t st, const char* fn) {
  if (int rc = check_common(fn, m, nnz, h, f)) return rc;
  if (m == 0) return DFGNN_OK;
  DotFwdParams p{m, nnz, h, f, 8, row_ptr, col_ind, val, Q, K, V, rn, out, attn};
  int rc = DFGNN_OK;
  dispatch_layout(f, [&](auto tag) {
    using L = typename decltype(tag)::type;
    constexpr int C = ChunkOf<L>::C;
    p.rb = pick_rb(m, nnz, L::G);
    const dim3 grid((m + p.rb - 1) / p.rb, h);
    const size_t smem = slot_bytes<L::NR, L>();
    if (agnn) {
      ensure_smem(dot_fwd_kernel<L, C, true>, smem);
      dot_fwd_kernel<L, C, true><<<grid, kNW * 32, smem, st>>>(p);
    } else {
      ensure_smem(dot_fwd_kernel<L, C, false>, smem);
      dot_fwd_kernel<L, C, false><<<grid, kNW * 32, smem, st>>>(p);
    }
    rc = check_launch(fn);
    note_kernel(0, "dot_fwd_kernel");
  }, long_rows(m, nnz));
  return rc;
}

static int gt_infer(const char* fn, int m, int nnz, int h, int f, const int32_t* indptr,
                    const int32_t* indices, const float* val, const float* Q, const float* K,
                    const float* V, float* out, void* stream) {
  if (m == 0) return check_common(fn, m, nnz, h, f);
  DFGNN_REQUIRE(indptr, fn);
  if (nnz > 0) DFGNN_REQUIRE(indices, fn);
  DFGNN_REQUIRE(Q, fn); DFGNN_REQUIRE(K, fn); DFGNN_REQUIRE(V, fn); DFGNN_REQUIRE(out, fn);
  return launch_dot_fwd(false, m, nnz, h, f, indptr, indices, val, Q, K, V, nullptr, out, nullptr,
                        (cudaStream_t)stream, fn);
}

}  // namespace dfgnn

using namespace dfgnn;

extern "C" {

int dfgnn_gt_hyper_forward(int m, int nnz, int h, int f, const int32_t* row_ptr,
                           const int32_t* col_ind, const int32_t* /*rows*/, const float* val,
                           const int32_t* /*col_ptr*/, const int32_t* /*row_ind*/,
                           const int32_t* /*val_idx*/, int /*smem_consume*/, const float* Q,
                           const float* K, const float* V, float* out_feat, float* attn_edge,
                           void* stream) {
  const char* fn = "dfgnn_gt_hyper_forward";
  if (m == 0) return check_common(fn, m, nnz, h, f);
  DFGNN_REQUIRE(row_ptr, fn);
  if (nnz > 0) { DFGNN_REQUIRE(col_ind, fn); DFGNN_REQUIRE(attn_edge, fn); }
  DFGNN_REQUIRE(Q, fn); DFGNN_REQUIRE(K, fn); DFGNN_REQUIRE(V, fn); DFGNN_REQUIRE(out_feat, fn);
  return launch_dot_fwd(false, m, nnz, h, f, row_ptr, col_ind, val, Q, K, V, nullptr, out_feat,
                        nnz > 0 ? attn_edge : nullptr, (cudaStream_t)stream, fn);
}

int dfgnn_gt_hyper_inference(int m, int nnz, int h, int f, const int32_t* indptr,
                             const int32_t* indices, const int32_t*, const float* val, int,
                             const float* Q, const float* K, const float* V, float* out_feat,
                             void* stream) {
  return gt_infer("dfgnn_gt_hyper_inference", m, nnz, h, f, indptr, indices, val, Q, K, V, out_feat, stream);
}
int dfgnn_gt_softmax_inference(int m, int nnz, int h, int f, const int32_t* indptr,
                               const int32_t* indices, const int32_t*, const float* val, int,
                               const float* Q, const float* K, const float* V, float* out_feat,
                               void* stream) {
  return gt_infer("dfgnn_gt_softmax_inference", m, nnz, h, f, indptr, indices, val, Q, K, V, out_feat, stream);
}
int dfgnn_gt_softmax_gm_inference(int m, int nnz, int h, int f, const int32_t* indptr,
                                  const int32_t* indices, const int32_t*, const float* val,
                                  const float* Q, const float* K, const float* V, float* out_feat,
                                  void* stream) {
  return gt_infer("dfgnn_gt_softmax_gm_inference", m, nnz, h, f, indptr, indices, val, Q, K, V, out_feat, stream);
}
int dfgnn_gt_tiling_inference(int m, int nnz, int h, int f, const int32_t* indptr,
                              const int32_t* indices, const float* val, int, const float* Q,
                              const float* K, const float* V, float* out_feat, void* stream) {
  return gt_infer("dfgnn_gt_tiling_inference", m, nnz, h, f, indptr, indices, val, Q, K, V, out_feat, stream);
}
int dfgnn_gt_csr_inference(int m, int nnz, int h, int f, const int32_t* indptr,
                           const int32_t* indices, const float* val, int, const float* Q,
                           const float* K, const float* V, float* out_feat, void* stream) {
  return gt_infer("dfgnn_gt_csr_inference", m, nnz, h, f, indptr, indices, val, Q, K, V, out_feat, stream);
}
int dfgnn_gt_csr_gm_inference(int m, int nnz, int h, int f, const int32_t* indptr,
                              const int32_t* indices, const float* val, const float* Q,
                              const float* K, const float* V, float* out_feat, void* stream) {
  return gt_infer("dfgnn_gt_csr_gm_inference", m, nnz, h, f, indptr, indices, val, Q, K, V, out_feat, stream);
}

int dfgnn_agnn_forward(int m, int nnz, int h, int f, const int32_t* indptr, const int32_t* indices,
                       const float* H, float* inv_norm, float* out_feat, float* attn_edge,
                       void* stream) {
  const char* fn = "dfgnn_agnn_forward";
  if (int rc = check_common(fn, m, nnz, h, f)) return rc;
  if (m == 0) return DFGNN_OK;
  DFGNN_REQUIRE(indptr, fn);
  if (nnz > 0) DFGNN_REQUIRE(indices, fn);
  DFGNN_REQUIRE(H, fn); DFGNN_REQUIRE(inv_norm, fn); DFGNN_REQUIRE(out_feat, fn);
  cudaStream_t st = (cudaStream_t)stream;
  const long long warps = (long long)m * h;
  inv_norm_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(m, h, f, H, inv_norm);
  if (int rc = check_launch(fn)) return rc;
  return launch_dot_fwd(true, m, nnz, h, f, indptr, indices, nullptr, H, H, H, inv_norm, out_feat,
                        nnz > 0 ? attn_edge : nullptr, st, fn);
}

}  // extern "C"
