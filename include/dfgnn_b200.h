/*
 * dfgnn_b200.h -- C ABI of the B200-native DF-GNN fused attention-conv library.
 *
 * This is the drop-in boundary: one entry point per function the reference's
 * pybind modules export for the hot path (`fused_gtconv`,
 * DFGNN/src/fused_gtconv/fused_gtconv.cpp:577-602; `fused_gatconv`,
 * DFGNN/src/fused_gatconv/fused_gatconv.cpp:355-372), plus the index-format
 * construction the reference delegates to dgl.sparse
 * (DFGNN/layers/util.py:52-162).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer valid on the current CUDA device;
 *     indices are int32, features / edge scalars are float32, contiguous;
 *   - shapes: Q/K/V/feat/out [m, h, f]; attn_row/attn_col/edge_max/edge_sum [m, h];
 *     row_ptr/col_ptr [m+1]; col_ind/row_ind/rows/val/val_idx/permute [nnz];
 *     GT attn_edge [h, nnz]; GAT edge_mask [nnz, h]; backward scratch: see below;
 *   - outputs are CALLER-allocated (the reference's launchers allocate them with
 *     torch::zeros / torch::empty, e.g. fused_gtconv_hyper.cu:688-691); no entry
 *     point requires an output to be pre-zeroed;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and
 *     the call returns without synchronising (the reference launches on the
 *     legacy default stream, DFGNN/src/util/computeUtil.h:257-265);
 *   - return value: 0 on success; DFGNN_ERR_* (< 0) for argument errors;
 *     a positive value is the cudaError_t of a failed launch.
 *     `dfgnn_last_error()` returns a thread-local message for the last failure
 *     (the reference raises TORCH_CHECK / glog CHECK failures instead,
 *     fused_gtconv.cpp:7-13, computeUtil.h:244-265);
 *   - arguments the reference's signatures carry but the new schedules do not
 *     need (`rows`, `smem_consume`) stay in the signatures and may be NULL / 0.
 *   - supported feature widths: any 1 <= f <= 512 (vectorised fast paths for
 *     f in {16, 32, 64, 128, 256, 512}); any h >= 1; no limit on row degree.
 */
#ifndef DFGNN_B200_H_
#define DFGNN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define DFGNN_OK 0
#define DFGNN_ERR_INVALID_ARGUMENT (-1)
#define DFGNN_ERR_UNSUPPORTED_DIM (-2)
#define DFGNN_ERR_WORKSPACE (-3)

#define DFGNN_ABI_VERSION 1

int dfgnn_abi_version(void);
const char *dfgnn_last_error(void);
/* Number of kernels this library has launched in this process (all entry points). */
uint64_t dfgnn_launch_count(void);
/* Name of the kernel the process's last forward (slot 0), backward row-side (slot 1) or
 * backward column-side (slot 2) call dispatched to, e.g. "gat_fwd_staged_kernel"; "" if none. */
const char *dfgnn_last_kernel(int slot);

/* ------------------------------------------------------------------------ */
/* Index-format construction                                                 */
/* ------------------------------------------------------------------------ */

/* Bytes of scratch the two builders below need for a graph of n nodes / nnz edges. */
size_t dfgnn_format_workspace_bytes(int64_t n, int64_t nnz); /* n = max(n_rows, n_cols) */

/*
 * COO -> CSR (+ the COO half of the hyper format).
 * Replaces g_to_SPmatrix + A.csr() + torch.sort(A.row) + A.val[val_idx] of
 * preprocess_CSR / preprocess_Hyper / preprocess_softmax
 * (DFGNN/layers/util.py:52-57, 66-79, 82-100, 145-162).
 * Stable by row: inside a row the input edge order is kept.
 *   row, col : [nnz] int64 (what torch.stack(g.edges()) holds); row < n_rows, col < n_cols
 *              (n_rows == n_cols for the reference's square adjacency; a row-partitioned
 *              shard of a full graph is n_rows_local x n_cols_global)
 *   row_ptr  : [n_rows+1]; col_ind, rows : [nnz]
 *   perm     : [nnz] sorted position -> input edge id (A.csr()'s value_indices), may be NULL
 *   val      : [nnz] float32, filled with 1.0f (A.val[val_idx] of an unweighted graph), may be NULL
 * A COO that is already sorted by row (the usual case) skips the sort.  The call synchronises
 * the stream once (index validation).
 */
int dfgnn_coo_to_csr(int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t *row,
                     const int64_t *col,
                     int32_t *row_ptr, int32_t *col_ind, int32_t *rows, int32_t *perm,
                     float *val, void *workspace, size_t workspace_bytes, void *stream);

/*
 * CSR -> CSC with the CSC-position -> CSR-position map.
 * Replaces dglsp.from_csr(...).csc() of preprocess_Hyper_fw_bw
 * (DFGNN/layers/util.py:136-141) and the scipy tocsc() `permute` of
 * DFGNN/script/train/train_gatconv.py:119-136.  Stable by column.
 */
int dfgnn_csr_to_csc(int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t *row_ptr,
                     const int32_t *col_ind, const int32_t *rows /* [nnz] from dfgnn_coo_to_csr, or NULL */,
                     int32_t *col_ptr, int32_t *row_ind, int32_t *val_idx, void *workspace,
                     size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------ */
/* GT / AGNN  (module `fused_gtconv`)                                        */
/* ------------------------------------------------------------------------ */

/*
 * Training forward.  Replaces gt_hyper_forward
 * (fused_gtconv.cpp:79-116 -> fused_gtconv_hyper.cu:727-760, kernel l.31-163).
 *   out[i] = sum_j softmax_j(<Q_i,K_j> * val_ij) V_j ; attn_edge[h, nnz] = the probabilities.
 * `val` may be NULL (== all ones).  Rows without edges produce 0.
 */
int dfgnn_gt_hyper_forward(int m, int nnz, int h, int f, const int32_t *row_ptr,
                           const int32_t *col_ind, const int32_t *rows, const float *val,
                           const int32_t *col_ptr, const int32_t *row_ind,
                           const int32_t *val_idx, int smem_consume, const float *Q,
                           const float *K, const float *V, float *out_feat, float *attn_edge,
                           void *stream);

/*
 * Backward.  Replaces gt_backward (fused_gtconv.cpp:125-172 ->
 * fused_gtconv_backward.cu:193-265).  grad_edge is scratch of 2 * h * nnz floats
 * ([h, nnz] pairs {dS_e, p_e} that the row-side kernel leaves for the column-side
 * kernel; the reference's scratch is torch::empty({h, nnz}), l.250).  m = rows (row_ptr, Q, grad_out, grad_Q),
 * n = columns (col_ptr, K, V, grad_K, grad_V); the reference reads both sizes too
 * (fused_gtconv_backward.cu:238-239) and they are equal for a square adjacency.
 * `val` (NULL == all ones) weights the scores as in the forward: grad_Q and grad_K carry the
 * factor val_e.  (The reference's backward ignores val, fused_gtconv_backward.cu:126, which is
 * the same thing for the all-ones val its preprocessing produces.)
 */
int dfgnn_gt_backward(int m, int n, int nnz, int h, int f, const int32_t *row_ptr,
                      const int32_t *col_ind, const int32_t *rows, const float *val,
                      const int32_t *col_ptr, const int32_t *row_ind, const int32_t *val_idx,
                      int smem_consume, const float *Q, const float *K, const float *V,
                      const float *attn_edge, const float *grad_out, float *grad_Q,
                      float *grad_K, float *grad_V, float *grad_edge, void *stream);

/*
 * The same with the two kernels of the backward selectable: phases = 1 row side only
 * (grad_Q and the {dS, p} scratch), 2 column side only (grad_K, grad_V; needs the scratch
 * a row-side call left), 3 both.  For per-kernel timing; no reference counterpart.
 */
int dfgnn_gt_backward_phase(int phases, int m, int n, int nnz, int h, int f,
                            const int32_t *row_ptr, const int32_t *col_ind, const int32_t *rows,
                            const float *val, const int32_t *col_ptr, const int32_t *row_ind,
                            const int32_t *val_idx, int smem_consume, const float *Q,
                            const float *K, const float *V, const float *attn_edge,
                            const float *grad_out, float *grad_Q, float *grad_K, float *grad_V,
                            float *grad_edge, void *stream);

/*
 * Column side only, restricted to columns [col_begin, col_begin + n_sub) of the n columns:
 * writes grad_K / grad_V rows of that range (pointers are the FULL arrays).  nnz_sub = entries
 * of those columns (schedule heuristics; < 0: unknown).  Used by the row-partitioned multi-GPU
 * operator (dfgnn_b200/dist.py) to reduce-scatter one column chunk while the next is computed;
 * no reference counterpart (the reference is single-GPU).
 */
int dfgnn_gt_backward_cols(int col_begin, int n_sub, int nnz_sub, int m, int n, int nnz, int h,
                           int f, const int32_t *row_ptr, const int32_t *col_ind,
                           const int32_t *rows, const float *val, const int32_t *col_ptr,
                           const int32_t *row_ind, const int32_t *val_idx, int smem_consume,
                           const float *Q, const float *K, const float *V,
                           const float *attn_edge, const float *grad_out, float *grad_Q,
                           float *grad_K, float *grad_V, float *grad_edge, void *stream);

/* ------------------------------------------------------------------------ */
/* Block-diagonal batches of small graphs (graph-resident kernels)            */
/* ------------------------------------------------------------------------ */
/*
 * The reference batches small graphs with dgl.batch (block-diagonal adjacency; the node ranges
 * are g.batch_num_nodes(), used by its preprocessing at DFGNN/layers/util.py:116-142 only
 * implicitly) and runs the same kernels as on a full graph.  Here a block plan -- blk_ptr
 * [n_blocks + 1], the first node of every graph -- selects kernels that keep one graph's operand
 * blocks in shared memory (csrc/block_gt.cuh).  Results are those of the entry points above.
 *
 * dfgnn_block_plan_check validates that every column id of a block's rows lies inside the block
 * (one small device -> host readback; flag_ws = 3 ints of device scratch) and returns the largest
 * block in *max_nodes_out (host) and, in *ascending_out (host, may be NULL), whether the column ids
 * of every row are strictly ascending (sorted, no duplicate edges) -- the dense kernels need that.  dfgnn_gt_block_supported tells whether the block kernels would
 * be used for this size in the current mode (h == 1, f in {32, 64, 128}, both operand blocks of
 * the largest graph fit shared memory); callers fall back to the general entry points otherwise.
 */
int dfgnn_block_plan_check(int n_blocks, int m, int nnz, const int32_t *blk_ptr,
                           const int32_t *row_ptr, const int32_t *col_ind, int32_t *flag_ws,
                           int32_t *max_nodes_out, int32_t *ascending_out, void *stream);
int dfgnn_gt_block_supported(int max_nodes, int m, int nnz, int h, int f);
/* 0 = automatic choice (default: the dense kernels below for dense batches), 1 = general kernels
 * only, 2 = the shared-memory-staged sparse kernels whenever they fit, 3 = the dense mma.sync
 * kernels whenever they fit, 4 = the dense tcgen05 kernels whenever they fit; returns the previous
 * mode (an out-of-range argument only queries). */
int dfgnn_set_block_mode(int mode);
/* = dfgnn_gt_hyper_forward (attn_edge may be NULL: inference). */
int dfgnn_gt_block_forward(int n_blocks, const int32_t *blk_ptr, int max_nodes, int m, int nnz,
                           int h, int f, const int32_t *row_ptr, const int32_t *col_ind,
                           const float *val, const float *Q, const float *K, const float *V,
                           float *out_feat, float *attn_edge, void *stream);
/*
 * Dense tensor-core variant for batches of DENSE small graphs (csrc/dense_gt.cuh: masked
 * 16-row tiles, mma.sync TF32 with the 3xTF32 split): unweighted scores (val == all ones), h == 1,
 * f in {64, 128}.  Same outputs as dfgnn_gt_hyper_forward.
 */
int dfgnn_gt_dense_supported(int max_nodes, int h, int f);
int dfgnn_gt_dense_forward(int n_blocks, const int32_t *blk_ptr, int max_nodes, int m, int nnz,
                           int h, int f, const int32_t *row_ptr, const int32_t *col_ind,
                           const float *Q, const float *K, const float *V, float *out_feat,
                           float *attn_edge, void *stream);
/*
 * Dense tcgen05 variant (csrc/dense_tc.cu): per 128-row tile S = Q K^T and O = P V as tcgen05.mma
 * kind::tf32 with the 3xTF32 split, accumulators in tensor memory, softmax by one thread per row
 * straight from tensor memory.  h == 1, f == 128, graphs of at most 256 nodes, unweighted scores,
 * strictly ascending column ids per row (dfgnn_block_plan_check).  Same outputs as
 * dfgnn_gt_hyper_forward (attn_edge may be NULL: inference).
 *
 * The adjacency is passed as a bitmap, a format built once per batch (like the CSC of
 * preprocess_Hyper_fw_bw, DFGNN/layers/util.py:116-142): adj_bits [m][8] uint32, bit j of row r =
 * (r, first node of r's graph + j) is an edge.  dfgnn_block_adj_bits builds it from the CSR.
 * The kernel is persistent (one CTA per SM walks several graphs); sched_ptr [n_ctas + 1] / sched_idx
 * [n_blocks] optionally give every CTA its own list of graphs (balanced by the caller from the graph
 * sizes, which it knows from batch_num_nodes); NULL = graphs dealt round robin over the SMs.
 */
int dfgnn_gt_dense_tc_supported(int max_nodes, int h, int f);
/* Test hook: writes NaN into every tensor-memory column of every SM (tensor memory keeps its contents
 * between kernels; the parity tests run the tcgen05 kernels after it). */
int dfgnn_tc_poison_tmem(void *stream);
/* Host-only helper: balanced work lists for the persistent CTAs from the graph sizes (nodes: HOST array
 * [n_blocks]); column_items = 0 lists graphs (forward, backward row side), 1 lists (graph, key tile) items
 * with id = 2 * graph + tile (backward column side).  ptr_out [n_ctas + 1], idx_out [2 * n_blocks] (host);
 * returns the number of CTAs used (<= n_ctas) or a negative error code. */
int dfgnn_tc_balanced_lists(int n_blocks, const int32_t *nodes, int n_ctas, int column_items,
                            int32_t *ptr_out, int32_t *idx_out);
int dfgnn_block_adj_bits(int n_blocks, int max_nodes, int m, int nnz, const int32_t *blk_ptr,
                         const int32_t *row_ptr, const int32_t *col_ind, uint32_t *adj_bits,
                         void *stream);
int dfgnn_gt_dense_tc_forward(int n_blocks, const int32_t *blk_ptr, int max_nodes, int m, int nnz,
                              int h, int f, const int32_t *row_ptr, const uint32_t *adj_bits,
                              int n_ctas, const int32_t *sched_ptr, const int32_t *sched_idx,
                              const float *Q, const float *K, const float *V, float *out_feat,
                              float *attn_edge, void *stream);
/*
 * Column side of the GT backward on tcgen05 (csrc/dense_tc.cu): grad_V = P^T grad_out and
 * grad_K = dS^T Q per (graph, 128-key tile) from the row side's packed scratch grad_edge [nnz][2] =
 * {dS_e, p_e} in CSR order (what dfgnn_gt_backward_phase(1, ...) leaves there).  Same requirements and
 * bitmap / schedule arguments as dfgnn_gt_dense_tc_forward; the schedule lists (graph, key tile) items,
 * item id = 2 * graph + key tile.  Replaces the column-side half of gt_backward
 * (DFGNN/src/fused_gtconv/fused_gtconv.cpp:125-172) for such batches.
 */
int dfgnn_gt_dense_tc_backward_col(int n_blocks, const int32_t *blk_ptr, int max_nodes, int m,
                                   int nnz, int h, int f, const int32_t *row_ptr,
                                   const uint32_t *adj_bits, int n_ctas, const int32_t *sched_ptr,
                                   const int32_t *sched_idx, const float *Q, const float *grad_out,
                                   const float *grad_edge, float *grad_K, float *grad_V,
                                   void *stream);
/*
 * Whole GT backward on tcgen05 for such batches (csrc/dense_tc.cu).  phases & 1: attn_edge is expanded to
 * dense per-graph tiles, then the row side runs with the forward's pipeline (dA = grad_out V^T, dS =
 * P (dA - rowsum(P dA)), grad_Q = dS K) and leaves dS as dense tiles; phases & 2: the column side
 * (grad_V = P^T grad_out, grad_K = dS^T Q) from those tiles.  dense_ws: device scratch of
 * dfgnn_gt_dense_tc_backward_ws_floats(m, n_blocks) floats that must persist between a phases = 1 and a
 * phases = 2 call; tile_ptr [n_blocks + 1]: exclusive prefix sum of ceil(nodes / 128) over the graphs (the
 * dense tiles are stored one 128-row image per row tile).  The first schedule lists graphs (row side), the second (graph, key tile) items (column side).
 * Same results as dfgnn_gt_backward (DFGNN/src/fused_gtconv/fused_gtconv.cpp:125-172).
 */
size_t dfgnn_gt_dense_tc_backward_ws_floats(int m, int n_blocks);
int dfgnn_gt_dense_tc_backward(int phases, int n_blocks, const int32_t *blk_ptr, int max_nodes, int m,
                               int nnz, int h, int f, const int32_t *row_ptr,
                               const uint32_t *adj_bits, int n_ctas, const int32_t *sched_ptr,
                               const int32_t *sched_idx, int n_ctas_col, const int32_t *sched_ptr_col,
                               const int32_t *sched_idx_col, const float *Q, const float *K,
                               const float *V, const float *attn_edge, const float *grad_out,
                               float *grad_Q, float *grad_K, float *grad_V,
                               const int32_t *tile_ptr, float *dense_ws, void *stream);
/* = dfgnn_gt_backward_phase on a square block-diagonal adjacency (n == m). */
int dfgnn_gt_block_backward(int phases, int n_blocks, const int32_t *blk_ptr, int max_nodes, int m,
                            int nnz, int h, int f, const int32_t *row_ptr, const int32_t *col_ind,
                            const float *val, const int32_t *col_ptr, const int32_t *row_ind,
                            const int32_t *val_idx, const float *Q, const float *K, const float *V,
                            const float *attn_edge, const float *grad_out, float *grad_Q,
                            float *grad_K, float *grad_V, float *grad_edge, void *stream);

/*
 * Inference entry points; all compute the same function, the name selects the
 * schedule heuristics.  Replace gt_hyper_inference (fused_gtconv.cpp:278-314),
 * gt_softmax_inference (l.316-352), gt_softmax_gm_inference (l.354-389),
 * gt_tiling_inference (l.244-276), gt_csr_inference (l.174-207),
 * gt_csr_gm_inference (l.209-242); export table l.577-602.
 */
int dfgnn_gt_hyper_inference(int m, int nnz, int h, int f, const int32_t *indptr,
                             const int32_t *indices, const int32_t *rows, const float *val,
                             int smem_consume, const float *Q, const float *K, const float *V,
                             float *out_feat, void *stream);
int dfgnn_gt_softmax_inference(int m, int nnz, int h, int f, const int32_t *indptr,
                               const int32_t *indices, const int32_t *rows, const float *val,
                               int smem_consume, const float *Q, const float *K, const float *V,
                               float *out_feat, void *stream);
int dfgnn_gt_softmax_gm_inference(int m, int nnz, int h, int f, const int32_t *indptr,
                                  const int32_t *indices, const int32_t *rows, const float *val,
                                  const float *Q, const float *K, const float *V, float *out_feat,
                                  void *stream);
int dfgnn_gt_tiling_inference(int m, int nnz, int h, int f, const int32_t *indptr,
                              const int32_t *indices, const float *val, int smem_consume,
                              const float *Q, const float *K, const float *V, float *out_feat,
                              void *stream);
int dfgnn_gt_csr_inference(int m, int nnz, int h, int f, const int32_t *indptr,
                           const int32_t *indices, const float *val, int smem_consume,
                           const float *Q, const float *K, const float *V, float *out_feat,
                           void *stream);
int dfgnn_gt_csr_gm_inference(int m, int nnz, int h, int f, const int32_t *indptr,
                              const int32_t *indices, const float *val, const float *Q,
                              const float *K, const float *V, float *out_feat, void *stream);

/*
 * AGNN fused entry (no reference export; fuses F.normalize of
 * DFGNN/layers/AGNN/agnn_layer_fused.py:14 into the conv so that one gathered
 * row of H serves both the score and the aggregation):
 *   out[i] = sum_j softmax_j(<H_i,H_j> / (max(|H_i|,eps) max(|H_j|,eps))) H_j
 * inv_norm [m, h] is caller-allocated scratch.  attn_edge may be NULL.
 */
int dfgnn_agnn_forward(int m, int nnz, int h, int f, const int32_t *indptr,
                       const int32_t *indices, const float *H, float *inv_norm, float *out_feat,
                       float *attn_edge, void *stream);

/* ------------------------------------------------------------------------ */
/* GAT  (module `fused_gatconv`)                                             */
/* ------------------------------------------------------------------------ */

/*
 * Training forward.  Replaces gat_forward (fused_gatconv.cpp:11-32 ->
 * fused_gatconv_kernel.cu:1062-1129, kernel l.24-125).
 *   e_ij = leakyrelu(attn_row[i] + attn_col[j]); edge_max/edge_sum = row max / sum exp;
 *   out[i] = sum_j keep_ij/(1-drop) * softmax_j(e_ij) * feat[j],
 *   keep_ij = edge_mask[e] > attn_drop, edge_mask ~ U(0,1] written by this call
 *   from a counter-based generator keyed by `seed` (the reference seeds cuRAND
 *   with clock(), l.1073, so masks are comparable only at attn_drop == 0).
 */
int dfgnn_gat_forward(int m, int nnz, int h, int f, const float *attn_row,
                      const float *attn_col, const int32_t *row_ptr, const int32_t *col_ind,
                      float negative_slope, const float *in_feat, float attn_drop,
                      uint64_t seed, float *out_feat, float *edge_max, float *edge_sum,
                      float *edge_mask, void *stream);

/*
 * Backward.  Replaces gat_backward (fused_gatconv.cpp:291-353 ->
 * fused_gatconv_kernel.cu:1171-1244).  grad_edge is scratch of 2 * nnz * h floats
 * ([nnz, h] pairs {de_e, keep-scaled p_e}; the reference's grad_edge_csr, l.1225,
 * holds nnz * h).  grad_attn_col is produced by a deterministic
 * column-side sum (the reference uses atomicAdd, l.854).  m = rows, n = columns
 * (grad_feat, grad_attn_col, attn_col, in_feat have n rows; the rest m).
 */
int dfgnn_gat_backward(int m, int n, int nnz, int h, int f, float negative_slope,
                       float attn_drop,
                       const int32_t *row_ptr, const int32_t *col_ind, const int32_t *col_ptr,
                       const int32_t *row_ind, const int32_t *permute, const float *edge_max,
                       const float *edge_sum, const float *edge_mask, const float *in_feat,
                       const float *attn_row, const float *attn_col, const float *grad_out,
                       float *grad_feat, float *grad_attn_row, float *grad_attn_col,
                       float *grad_edge, void *stream);

/* Row side (1: grad_attn_row + scratch), column side (2: grad_feat, grad_attn_col) or both (3). */
int dfgnn_gat_backward_phase(int phases, int m, int n, int nnz, int h, int f,
                             float negative_slope, float attn_drop, const int32_t *row_ptr,
                             const int32_t *col_ind, const int32_t *col_ptr,
                             const int32_t *row_ind, const int32_t *permute,
                             const float *edge_max, const float *edge_sum,
                             const float *edge_mask, const float *in_feat, const float *attn_row,
                             const float *attn_col, const float *grad_out, float *grad_feat,
                             float *grad_attn_row, float *grad_attn_col, float *grad_edge,
                             void *stream);

/* Column side on columns [col_begin, col_begin + n_sub) only; see dfgnn_gt_backward_cols. */
int dfgnn_gat_backward_cols(int col_begin, int n_sub, int nnz_sub, int m, int n, int nnz, int h,
                            int f, float negative_slope, float attn_drop,
                            const int32_t *row_ptr, const int32_t *col_ind,
                            const int32_t *col_ptr, const int32_t *row_ind,
                            const int32_t *permute, const float *edge_max,
                            const float *edge_sum, const float *edge_mask, const float *in_feat,
                            const float *attn_row, const float *attn_col, const float *grad_out,
                            float *grad_feat, float *grad_attn_row, float *grad_attn_col,
                            float *grad_edge, void *stream);

/*
 * Inference entry points (one function, several schedule names).  Replace
 * gat_inference (fused_gatconv.cpp:225-254), gat_inference_hyper (l.99-124),
 * gat_inference_hyper_recompute (l.126-150), gat_inference_softmax (l.40-68),
 * gat_inference_softmax_gm (l.70-97), gat_inference_tiling (l.196-223);
 * export table l.355-372.
 */
int dfgnn_gat_inference(int m, int nnz, int h, int f, const float *attn_row,
                        const float *attn_col, const int32_t *row_ptr, const int32_t *col_ind,
                        float negative_slope, const float *in_feat, float *out_feat,
                        void *stream);
int dfgnn_gat_inference_hyper(int smem_consume, int m, int nnz, int h, int f,
                              const float *attn_row, const float *attn_col,
                              const int32_t *indptr, const int32_t *indices,
                              const int32_t *rows, float negative_slope, const float *in_feat,
                              float *out_feat, void *stream);
int dfgnn_gat_inference_hyper_recompute(int m, int nnz, int h, int f, const float *attn_row,
                                        const float *attn_col, const int32_t *indptr,
                                        const int32_t *indices, float negative_slope,
                                        const float *in_feat, float *out_feat, void *stream);
int dfgnn_gat_inference_softmax(int smem_consume, int m, int nnz, int h, int f,
                                const float *attn_row, const float *attn_col,
                                const int32_t *indptr, const int32_t *indices,
                                const int32_t *rows, float negative_slope,
                                const float *in_feat, float *out_feat, void *stream);
int dfgnn_gat_inference_softmax_gm(int m, int nnz, int h, int f, const float *attn_row,
                                   const float *attn_col, const int32_t *indptr,
                                   const int32_t *indices, const int32_t *rows,
                                   float negative_slope, const float *in_feat, float *out_feat,
                                   void *stream);
int dfgnn_gat_inference_tiling(int m, int nnz, int h, int f, const float *attn_row,
                               const float *attn_col, const int32_t *row_ptr,
                               const int32_t *col_ind, float negative_slope,
                               const float *in_feat, float *out_feat, void *stream);
/*
 * hyper_v2: attention logits computed inside the call from a_l / a_r [h, f].
 * Replaces gat_inference_hyper_v2 (fused_gatconv.cpp:152-166 ->
 * fused_gatconv_hyper_v2.cu:212-305).  attn_row / attn_col [m, h] are
 * caller-allocated scratch (torch::empty in the reference, l.293-294).
 */
int dfgnn_gat_inference_hyper_v2(int smem_consume, int m, int nnz, int h, int f,
                                 const float *a_l, const float *a_r, const int32_t *indptr,
                                 const int32_t *indices, float negative_slope,
                                 const float *in_feat, float *attn_row, float *attn_col,
                                 float *out_feat, void *stream);

/* The attention-logit prologue on its own (fused_gatconv_hyper_v2.cu:212-250). */
int dfgnn_gat_attn_weight(int m, int h, int f, const float *a_l, const float *a_r,
                          const float *in_feat, float *attn_row, float *attn_col, void *stream);

/* ------------------------------------------------------------------------ */
/* The dense projection in front of the conv (tcgen05 tensor cores, 3xTF32)   */
/* ------------------------------------------------------------------------ */
/*
 * Replaces the three nn.Linear GEMMs + scale + reshape of SparseMHA.prep_qkv
 * (DFGNN/layers/GT/gtconv_layer.py:19-27; fused branch gtconv_layer_fused.py:20-22) and, for GAT,
 * feat = W x plus the logits attn_row = <a_l, feat>, attn_col = <a_r, feat>
 * (layers/GAT/gatconv_layer_fused.py:121-123, fused_gatconv_hyper_v2.cu:212-250) with one kernel:
 *     Y[n, n_out] = (X[n, k] W[n_out, k]^T + bias) * scale
 * at fp32-grade accuracy (TF32 tensor cores, x = hi + lo split, fp32 accumulation in tensor
 * memory).  Y is written as n_out / part_width tensors [n, part_width] (out0..out3; q | k | v are
 * three parts of width heads * d, i.e. already [N, heads, d]).  k in {32, 64, 128}; n_out and
 * part_width multiples of 64; bias / scale [n_out] or NULL.  head_dim > 0 also writes the logits
 * [n, n_out / head_dim] (head_dim 8, 16 or a multiple of 32; a_l, a_r [n_out]).
 *
 * dfgnn_proj_pack_weights splits W into the hi / lo operand images the kernel loads by TMA
 * (w_img: dfgnn_proj_weight_image_floats(n_out, k) floats); call it once per weight update.
 */
size_t dfgnn_proj_weight_image_floats(int n_out, int k);
int dfgnn_proj_pack_weights(int n_out, int k, const float *W, float *w_img, void *stream);
int dfgnn_proj_forward(int n, int k, int n_out, int part_width, const float *x, const float *w_img,
                       const float *bias, const float *scale, float *out0, float *out1, float *out2,
                       float *out3, int head_dim, const float *a_l, const float *a_r,
                       float *attn_row, float *attn_col, void *stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* DFGNN_B200_H_ */
