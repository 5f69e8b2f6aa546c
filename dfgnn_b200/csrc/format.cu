// format.cu -- index-format construction on the GPU (COO -> CSR + rows + perm, CSR -> CSC + val_idx).
//
// Replaces the dgl.sparse calls of DFGNN/layers/util.py:52-162 (A.csr(), A.csc(),
// torch.sort(A.row)).  Integer work, bit-exact by definition:
// both sorts are STABLE least-significant-digit radix sorts on just the
// ceil(log2 n) significant key bits.
#include <cub/device/device_radix_sort.cuh>

#include "abi_common.h"

namespace dfgnn {

static inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static inline int key_bits(int64_t n) {
  int b = 1;
  while (b < 31 && ((int64_t)1 << b) < n) ++b;
  return b;
}

// Validates the endpoints, builds the sort keys, and -- speculating that the COO is already
// sorted by row (what DGL graphs and batched graphs usually are) -- writes the CSR arrays of
// that case directly.  flags[0] = an endpoint is out of range, flags[1] = rows are not
// non-decreasing (the radix sort then overwrites the speculative output).
__global__ void coo_keys_kernel(int64_t nnz, int64_t n, int64_t n_cols, const int64_t* __restrict__ row,
                                const int64_t* __restrict__ col, int32_t* __restrict__ keys,
                                int32_t* __restrict__ ids, int32_t* __restrict__ rows,
                                int32_t* __restrict__ col_ind, int32_t* __restrict__ perm,
                                float* __restrict__ val, int* __restrict__ flags) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnz) return;
  const int64_t r = row[i], c = col[i];
  if (r < 0 || r >= n || c < 0 || c >= n_cols) { flags[0] = 1; keys[i] = 0; }
  else keys[i] = (int32_t)r;
  ids[i] = (int32_t)i;
  if (i > 0 && row[i - 1] > r) flags[1] = 1;
  rows[i] = (int32_t)r;
  col_ind[i] = (int32_t)c;
  if (perm) perm[i] = (int32_t)i;
  if (val) val[i] = 1.0f;
}

// ids = 0..nnz-1; flag[0] = a column id lies outside [0, n_cols)
__global__ void iota_check_kernel(int64_t nnz, int64_t n_cols, const int32_t* __restrict__ col_ind,
                                  int32_t* __restrict__ ids, int* __restrict__ flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnz) return;
  ids[i] = (int32_t)i;
  const int32_t c = col_ind[i];
  if (c < 0 || c >= n_cols) flag[0] = 1;
}

__global__ void gather_col_kernel(int64_t nnz, const int64_t* __restrict__ col,
                                  const int32_t* __restrict__ perm, int32_t* __restrict__ col_ind,
                                  float* __restrict__ val) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnz) return;
  col_ind[i] = (int32_t)col[perm[i]];
  if (val) val[i] = 1.0f;
}

// ptr[s] = first position p with sorted[p] >= s, for s in [0, n]
__global__ void seg_ptr_kernel(int64_t nnz, int64_t n, const int32_t* __restrict__ sorted,
                               int32_t* __restrict__ ptr) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p > nnz) return;
  int64_t prev = p > 0 ? sorted[p - 1] : -1;
  int64_t cur = p < nnz ? sorted[p] : n;
  // keys outside [0, n) are rejected by the callers' validation; never write outside ptr[0..n]
  prev = prev < -1 ? -1 : (prev > n ? n : prev);
  cur = cur < 0 ? 0 : (cur > n ? n : cur);
  for (int64_t s = prev + 1; s <= cur; ++s) ptr[s] = (int32_t)p;
}

// row_ind[p] = rows[val_idx[p]] when the caller has the expanded row ids of the CSR
__global__ void row_gather_kernel(int64_t nnz, const int32_t* __restrict__ rows,
                                  const int32_t* __restrict__ val_idx, int32_t* __restrict__ row_ind) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nnz) row_ind[i] = __ldg(rows + val_idx[i]);
}

// row_ind[p] = the row whose CSR range contains position val_idx[p]
__global__ void row_of_kernel(int64_t nnz, int64_t n, const int32_t* __restrict__ row_ptr,
                              const int32_t* __restrict__ val_idx, int32_t* __restrict__ row_ind) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnz) return;
  const int32_t e = val_idx[i];
  int64_t lo = 0, hi = n;  // row_ptr[lo] <= e < row_ptr[hi]
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(row_ptr + mid) <= e) lo = mid; else hi = mid;
  }
  row_ind[i] = (int32_t)lo;
}

static size_t sort_temp_bytes(int64_t nnz, int bits) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)nnz, 0, bits);
  return bytes;
}

static inline unsigned blocks(int64_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

}  // namespace dfgnn

using namespace dfgnn;

extern "C" {

size_t dfgnn_format_workspace_bytes(int64_t n, int64_t nnz) {
  if (n < 0 || nnz < 0) return 0;
  const size_t e = align_up((size_t)nnz * sizeof(int32_t));
  // keys_in, ids_in, keys_out (csc) / perm (csr), flag, cub temp
  return 3 * e + 256 + align_up(sort_temp_bytes(nnz > 0 ? nnz : 1, key_bits(n))) + 256;
}

int dfgnn_coo_to_csr(int64_t n, int64_t n_cols, int64_t nnz, const int64_t* row, const int64_t* col,
                     int32_t* row_ptr, int32_t* col_ind, int32_t* rows, int32_t* perm, float* val,
                     void* workspace, size_t workspace_bytes, void* stream) {
  const char* fn = "dfgnn_coo_to_csr";
  if (n < 0 || n_cols < 0 || nnz < 0 || n > INT32_MAX || n_cols > INT32_MAX || nnz > INT32_MAX) {
    set_error("%s: n_rows=%lld n_cols=%lld nnz=%lld out of int32 range", fn, (long long)n,
              (long long)n_cols, (long long)nnz);
    return DFGNN_ERR_INVALID_ARGUMENT;
  }
  DFGNN_REQUIRE(row_ptr, fn);
  cudaStream_t st = (cudaStream_t)stream;
  if (nnz == 0) {
    if (cudaMemsetAsync(row_ptr, 0, (size_t)(n + 1) * sizeof(int32_t), st) != cudaSuccess) {
      set_error("%s: memset failed", fn);
      return (int)cudaGetLastError();
    }
    return DFGNN_OK;
  }
  DFGNN_REQUIRE(row, fn); DFGNN_REQUIRE(col, fn); DFGNN_REQUIRE(col_ind, fn); DFGNN_REQUIRE(rows, fn);
  DFGNN_REQUIRE(workspace, fn);
  const int64_t n_max = n > n_cols ? n : n_cols;
  if (workspace_bytes < dfgnn_format_workspace_bytes(n_max, nnz)) {
    set_error("%s: workspace too small (%zu < %zu)", fn, workspace_bytes,
              dfgnn_format_workspace_bytes(n_max, nnz));
    return DFGNN_ERR_WORKSPACE;
  }
  const int bits = key_bits(n);
  const size_t e = align_up((size_t)nnz * sizeof(int32_t));
  char* ws = (char*)workspace;
  int32_t* keys_in = (int32_t*)ws;
  int32_t* ids_in = (int32_t*)(ws + e);
  int32_t* perm_buf = perm ? perm : (int32_t*)(ws + 2 * e);
  int* bad = (int*)(ws + 3 * e);
  void* temp = ws + 3 * e + 256;
  size_t temp_bytes = workspace_bytes - (3 * e + 256);

  cudaMemsetAsync(bad, 0, 2 * sizeof(int), st);
  coo_keys_kernel<<<blocks(nnz), 256, 0, st>>>(nnz, n, n_cols, row, col, keys_in, ids_in, rows, col_ind,
                                               perm_buf, val, bad);
  if (int rc = check_launch(fn)) return rc;
  // one small readback: index validation, and whether the input still has to be sorted
  int h_flags[2] = {0, 0};
  cudaError_t err = cudaMemcpyAsync(h_flags, bad, 2 * sizeof(int), cudaMemcpyDeviceToHost, st);
  if (err == cudaSuccess) err = cudaStreamSynchronize(st);
  if (err != cudaSuccess) { set_error("%s: %s", fn, cudaGetErrorString(err)); return (int)err; }
  if (h_flags[0]) {
    set_error("%s: edge endpoint outside [0, %lld) x [0, %lld)", fn, (long long)n, (long long)n_cols);
    return DFGNN_ERR_INVALID_ARGUMENT;
  }
  if (h_flags[1]) {  // not sorted by row: stable radix sort, then gather the columns
    err = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, rows, ids_in, perm_buf, (int)nnz, 0,
                                          bits, st);
    launch_counter().fetch_add(1);
    if (err != cudaSuccess) { set_error("%s: radix sort: %s", fn, cudaGetErrorString(err)); return (int)err; }
    gather_col_kernel<<<blocks(nnz), 256, 0, st>>>(nnz, col, perm_buf, col_ind, val);
    if (int rc = check_launch(fn)) return rc;
  }
  seg_ptr_kernel<<<blocks(nnz + 1), 256, 0, st>>>(nnz, n, rows, row_ptr);
  return check_launch(fn);
}

int dfgnn_csr_to_csc(int64_t n_rows, int64_t n, int64_t nnz, const int32_t* row_ptr, const int32_t* col_ind,
                     const int32_t* rows, int32_t* col_ptr, int32_t* row_ind, int32_t* val_idx,
                     void* workspace, size_t workspace_bytes, void* stream) {
  const char* fn = "dfgnn_csr_to_csc";
  if (n < 0 || n_rows < 0 || nnz < 0 || n > INT32_MAX || n_rows > INT32_MAX || nnz > INT32_MAX) {
    set_error("%s: n_rows=%lld n_cols=%lld nnz=%lld out of int32 range", fn, (long long)n_rows,
              (long long)n, (long long)nnz);
    return DFGNN_ERR_INVALID_ARGUMENT;
  }
  DFGNN_REQUIRE(col_ptr, fn);
  cudaStream_t st = (cudaStream_t)stream;
  if (nnz == 0) {
    cudaMemsetAsync(col_ptr, 0, (size_t)(n + 1) * sizeof(int32_t), st);
    return DFGNN_OK;
  }
  DFGNN_REQUIRE(row_ptr, fn); DFGNN_REQUIRE(col_ind, fn); DFGNN_REQUIRE(row_ind, fn);
  DFGNN_REQUIRE(val_idx, fn); DFGNN_REQUIRE(workspace, fn);
  const int64_t n_max = n > n_rows ? n : n_rows;
  if (workspace_bytes < dfgnn_format_workspace_bytes(n_max, nnz)) {
    set_error("%s: workspace too small (%zu < %zu)", fn, workspace_bytes,
              dfgnn_format_workspace_bytes(n_max, nnz));
    return DFGNN_ERR_WORKSPACE;
  }
  const int bits = key_bits(n);
  const size_t e = align_up((size_t)nnz * sizeof(int32_t));
  char* ws = (char*)workspace;
  int32_t* ids_in = (int32_t*)ws;
  int32_t* keys_out = (int32_t*)(ws + e);
  int* bad = (int*)(ws + 3 * e);
  void* temp = ws + 3 * e + 256;
  size_t temp_bytes = workspace_bytes - (3 * e + 256);

  cudaMemsetAsync(bad, 0, sizeof(int), st);
  iota_check_kernel<<<blocks(nnz), 256, 0, st>>>(nnz, n, col_ind, ids_in, bad);
  if (int rc = check_launch(fn)) return rc;
  int h_flag = 0;  // one small readback, like dfgnn_coo_to_csr: the sort below trusts the key range
  cudaError_t err = cudaMemcpyAsync(&h_flag, bad, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (err == cudaSuccess) err = cudaStreamSynchronize(st);
  if (err != cudaSuccess) { set_error("%s: %s", fn, cudaGetErrorString(err)); return (int)err; }
  if (h_flag) {
    set_error("%s: column index outside [0, %lld)", fn, (long long)n);
    return DFGNN_ERR_INVALID_ARGUMENT;
  }
  err = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, col_ind, keys_out, ids_in,
                                                    val_idx, (int)nnz, 0, bits, st);
  launch_counter().fetch_add(1);
  if (err != cudaSuccess) { set_error("%s: radix sort: %s", fn, cudaGetErrorString(err)); return (int)err; }
  seg_ptr_kernel<<<blocks(nnz + 1), 256, 0, st>>>(nnz, n, keys_out, col_ptr);
  if (int rc = check_launch(fn)) return rc;
  if (rows) row_gather_kernel<<<blocks(nnz), 256, 0, st>>>(nnz, rows, val_idx, row_ind);
  else row_of_kernel<<<blocks(nnz), 256, 0, st>>>(nnz, n_rows, row_ptr, val_idx, row_ind);
  return check_launch(fn);
}

}  // extern "C"
