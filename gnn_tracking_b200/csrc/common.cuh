// Shared helpers of the gtb200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/gtb200.h"

namespace gtb {

void set_error(const char* fmt, ...);

inline int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return GTB_OK;
  // "out of memory" must stay in the text: the reference's tolerate_some_oom_errors
  // (utils/oom.py:12-18) looks for it.
  set_error("%s: %s", what, cudaGetErrorString(e));
  return GTB_ERR_CUDA;
}

#define GTB_CHECK_LAUNCH(what)                                   \
  do {                                                           \
    int _rc = ::gtb::check_cuda(cudaGetLastError(), what);       \
    if (_rc != GTB_OK) return _rc;                               \
  } while (0)

#define GTB_REQUIRE(cond, code, ...)      \
  do {                                    \
    if (!(cond)) {                        \
      ::gtb::set_error(__VA_ARGS__);      \
      return code;                        \
    }                                     \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// cudaFuncSetAttribute is a per-device setting: one flag per device and kernel family (the host side makes
// the tensors' device current before it calls in, see ops.on_device)
struct PerDeviceOnce {
  bool done[64] = {false};
  bool* slot() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    return &done[dev];
  }
};

inline int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }
inline int64_t imax64(int64_t a, int64_t b) { return a > b ? a : b; }

__host__ __device__ inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// ---------------------------------------------------------------- packed MLP (FFMA layout)
// Layer l: Wt[Kp_l][Nw_l] (K-major, zero padded) followed by bias[Nw_l].
//   Kp_0 = round_up(K0, 32);  Kp_l = Nw_{l-1}  (the padded width of the previous layer)
//   Nw_l = 64 * NC for every layer except a "narrow" last layer (N <= 8) which uses Nw = 8.
struct FfmaLayout {
  int nc;                         // 64-column groups of the wide layers (1 or 2)
  int kp[GTB_MAX_LAYERS];
  int nw[GTB_MAX_LAYERS];
  int narrow_last;
  size_t w_off[GTB_MAX_LAYERS];   // float offsets
  size_t b_off[GTB_MAX_LAYERS];
  size_t total_floats;
};

__host__ __device__ inline bool ffma_layout(int n_layers, const int32_t* dims, FfmaLayout* L) {
  int maxn = 0;
  const int last = n_layers - 1;
  L->narrow_last = dims[last + 1] <= 8;
  for (int l = 0; l < n_layers; ++l) {
    if (l == last && L->narrow_last) continue;
    if (dims[l + 1] > maxn) maxn = dims[l + 1];
  }
  if (maxn > GTB_MAX_WIDTH) return false;
  L->nc = maxn > 64 ? 2 : 1;
  size_t off = 0;
  for (int l = 0; l < n_layers; ++l) {
    L->kp[l] = (l == 0) ? round_up(dims[0], 32) : L->nw[l - 1];
    L->nw[l] = (l == last && L->narrow_last) ? 8 : 64 * L->nc;
    L->w_off[l] = off;
    off += (size_t)L->kp[l] * L->nw[l];
    L->b_off[l] = off;
    off += L->nw[l];
  }
  L->total_floats = off;
  return true;
}

}  // namespace gtb
