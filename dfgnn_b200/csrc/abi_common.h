// abi_common.h -- host-side helpers shared by the translation units that
// implement include/dfgnn_b200.h: error reporting, launch accounting, layout
// dispatch and schedule heuristics.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <type_traits>

#include "../../include/dfgnn_b200.h"
#include "common.cuh"
#include "rowblock.cuh"
#include "staged.cuh"

#ifndef DFGNN_LPR64
#define DFGNN_LPR64 8
#endif
#ifndef DFGNN_LPR128
#define DFGNN_LPR128 16
#endif

namespace dfgnn {

void set_error(const char* fmt, ...);
// which kernel family the last forward (slot 0) / backward row side (1) / column side (2) call of
// this process dispatched to (dfgnn_last_kernel; bench.py labels its per-kernel numbers with it)
void note_kernel(int slot, const char* name);
std::atomic<uint64_t>& launch_counter();

inline int check_launch(const char* what) {
  launch_counter().fetch_add(1, std::memory_order_relaxed);
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(err));
    return (int)err;
  }
  return DFGNN_OK;
}

template <class T>
struct Tag { using type = T; };

// Picks the lane-group layout for a feature width.  Returns false for f > 512.
// 8 lanes per row wherever the row is at least 8 float4 wide: 4 rows per warp
// instruction stream, 3-stage group reductions.
// `long_rows` (mean degree > 128, GT kernels only): f = 128 keeps 8 lanes per row there (16 floats
// per lane, 16 resident warps) -- on super-node graphs the deeper per-lane rows win, on short
// and medium rows 16 lanes x 8 floats at 24 resident warps do (measured: reddit- vs PATTERN- /
// VOC-shaped workloads).
template <class Fn>
inline bool dispatch_layout(int f, Fn&& fn, bool long_rows = false) {
  if (f == 16) fn(Tag<VecLayout<4, 4>>{});
  else if (f == 32) fn(Tag<VecLayout<8, 8>>{});
  else if (f == 64) fn(Tag<VecLayout<16, DFGNN_LPR64>>{});
  else if (f == 128 && long_rows) fn(Tag<VecLayout<32, 8>>{});
  else if (f == 128) fn(Tag<VecLayout<32, DFGNN_LPR128>>{});
  else if (f == 256) fn(Tag<VecLayout<64, 16>>{});
  else if (f == 512) fn(Tag<VecLayout<128, 32>>{});
  else if (f <= 32) fn(Tag<ScalarLayout<1>>{});
  else if (f <= 64) fn(Tag<ScalarLayout<2>>{});
  else if (f <= 96) fn(Tag<ScalarLayout<3>>{});
  else if (f <= 128) fn(Tag<ScalarLayout<4>>{});
  else if (f <= 256) fn(Tag<ScalarLayout<8>>{});
  else if (f <= 512) fn(Tag<ScalarLayout<16>>{});
  else return false;
  return true;
}

// Segments (rows / columns) per CTA: aim at ~96 entries per lane group (measured on the PATTERN-
// and VOC-shaped batches: 16 / 32 / 64 rows per CTA -> 1.33 / 1.29 / 1.35 ms resp. 1.53 / 1.20 / 1.04 ms) (kNW * G groups
// per CTA), but keep at least ~4 CTAs per SM in the grid (148 SMs) so that small graphs
// still fill the chip; [8, kMaxRB].
inline int pick_rb(int m, int nnz, int G) {
  static const int forced = [] { const char* e = getenv("DFGNN_B200_RB"); return e ? atoi(e) : 0; }();
  if (forced >= 8 && forced <= kMaxRB) return forced;  // developer knob
  const double avg = m > 0 ? (double)nnz / (double)m : 0.0;
  int rb = 8;
  while (rb < kMaxRB && avg * rb < 96.0 * kNW * G && (m / (2 * rb)) >= 4 * 148) rb <<= 1;
  return rb;
}

// ---- staged-tile kernels (staged.cuh): short rows ------------------------------------
// Used (GAT kernels) when the mean degree is at most kStagedMaxDegree: measured crossover on
// B200 with 200k-node log-normal graphs, f = 64 / 128 -- staged wins at mean degree 10 and 16
// (by 8 % / 0-1 %), the row-block kernels win from 24 up (PATTERN-shaped, degree 51: by 1.5-1.8x).
// Rows per CTA aim at ~kStageCap/2 entries, keeping >= 2 CTAs per SM in the grid.
// DFGNN_B200_SCHEDULE=rowblock|staged overrides the choice (developer / test knob).
constexpr double kStagedMaxDegree = 16.0;
inline int schedule_override() {
  static const int v = [] {
    const char* e = getenv("DFGNN_B200_SCHEDULE");
    if (!e) return 0;
    return e[0] == 'r' ? 1 : (e[0] == 's' ? 2 : 0);
  }();
  return v;
}
inline bool want_staged(int m, int nnz) {
  const int o = schedule_override();
  if (o) return o == 2;
  return m > 0 && (double)nnz / (double)m <= kStagedMaxDegree;
}
#ifndef DFGNN_STAGE_TARGET
#define DFGNN_STAGE_TARGET (kStageCap / 2)
#endif
inline int pick_rb_staged(int m, int nnz) {
  const double avg = m > 0 ? (double)nnz / (double)m : 0.0;
  int rb = kMaxRB;
  while (rb > 8 && (avg * rb > DFGNN_STAGE_TARGET || (m + rb - 1) / rb < 2 * 148)) rb >>= 1;
  return rb;
}
// entries in flight per lane group in the staged kernels
template <class L>
struct StageChunk {
#ifndef DFGNN_SPMM_C
#define DFGNN_SPMM_C 4
#endif
  static constexpr int kSpmm = L::NR <= 8 ? DFGNN_SPMM_C : (L::NR <= 16 ? 4 : 2);
  static constexpr int kSpmm2 = L::NR <= 8 ? 4 : 2;  // two operand matrices
  static constexpr int kRaw = L::NR <= 8 ? 4 : 2;    // sddmm also holds two row operands
  static constexpr int kSddmm = stage_x<L>() ? 4 : (kRaw < L::LPR ? kRaw : L::LPR);
};

inline bool long_rows(int m, int nnz) { return m > 0 && (double)nnz / (double)m > 128.0; }

inline int check_common(const char* fn, int m, int nnz, int h, int f) {
  if (m < 0 || nnz < 0 || h < 1 || f < 1) {
    set_error("%s: invalid sizes m=%d nnz=%d h=%d f=%d", fn, m, nnz, h, f);
    return DFGNN_ERR_INVALID_ARGUMENT;
  }
  if (f > 512) {
    set_error("%s: feature width f=%d is not supported (max 512)", fn, f);
    return DFGNN_ERR_UNSUPPORTED_DIM;
  }
  if (h > 65535) {
    set_error("%s: h=%d exceeds the grid limit", fn, h);
    return DFGNN_ERR_INVALID_ARGUMENT;
  }
  return DFGNN_OK;
}

#define DFGNN_REQUIRE(ptr, fn)                                        \
  do {                                                                \
    if ((ptr) == nullptr) {                                           \
      ::dfgnn::set_error("%s: argument `%s` is NULL", fn, #ptr);      \
      return DFGNN_ERR_INVALID_ARGUMENT;                              \
    }                                                                 \
  } while (0)

template <int NV, class L>
constexpr size_t slot_bytes() {
  return (size_t)kNW * L::G * 2 * Slot<NV, L::LPR>::kFloats * sizeof(float);
}

// kernels whose partial-result slots exceed the 48 KB default need the opt-in
template <class K>
inline void ensure_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024)
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
// same for kernels that also hold `static_bytes` of static shared memory (staged kernels)
template <class K>
inline void ensure_smem(K kernel, size_t bytes, size_t static_bytes) {
  if (bytes + static_bytes > 48 * 1024)
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// Launch `kernel` so that it may start while the kernel in front of it on the stream is still
// draining (programmatic dependent launch).  Used for the "big tiles only" row-block launch
// behind a staged kernel: the two write disjoint tiles and read only inputs, so there is no
// dependency to wait for; the staged kernel issues griddepcontrol.launch_dependents at entry.
template <class P>
inline void launch_overlapped(void (*kernel)(P), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              const P& p) {
  static const bool off = getenv("DFGNN_B200_NO_PDL") != nullptr;  // developer knob
  if (off) {
    kernel<<<grid, block, smem, st>>>(p);
    return;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, p);
}

// internal launchers (one per kernel family), defined in gt.cu / gat.cu
int launch_dot_fwd(bool agnn, int m, int nnz, int h, int f, const int* row_ptr, const int* col_ind,
                   const float* val, const float* Q, const float* K, const float* V,
                   const float* rn, float* out, float* attn, cudaStream_t st, const char* fn);
int launch_gat_fwd(int m, int nnz, int h, int f, const float* ar, const float* ac,
                   const int* row_ptr, const int* col_ind, float slope, const float* feat,
                   float drop, uint64_t seed, float* out, float* emax, float* esum, float* emask,
                   cudaStream_t st, const char* fn);

}  // namespace dfgnn
