// Shared internals of libttk: error reporting, launch checks, small device helpers.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ttk.h"

void ttk_set_error(const char* fmt, ...);

#define TTK_CHECK_ARG(cond, ...)     \
  do {                               \
    if (!(cond)) {                   \
      ttk_set_error(__VA_ARGS__);    \
      return TTK_ERR_ARG;            \
    }                                \
  } while (0)

#define TTK_CUDA(expr)                                                                        \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      ttk_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__));   \
      return TTK_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define TTK_LAUNCH_CHECK() TTK_CUDA(cudaGetLastError())

static inline int ttk_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// number of SMs of the current device (cached)
int ttk_num_sms();

// Handles own device memory (packed weights): they bind to the device that is current at their first upload and refuse to run on another.
static inline int ttk_bind_device(int* slot, const char* what) {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess) {
    ttk_set_error("%s: cudaGetDevice failed", what);
    return TTK_ERR_CUDA;
  }
  if (*slot < 0) *slot = d;
  if (*slot != d) {
    ttk_set_error("%s: the handle's weights live on device %d but the current device is %d (create one handle per device)", what, *slot, d);
    return TTK_ERR_STATE;
  }
  return TTK_OK;
}

// One-time, per-DEVICE setup (cudaFuncSetAttribute applies to the current device only): `static TtkPerDevice once; if (once.first()) {...}`.
// A process that drives several GPUs runs the setup once on each; racing threads at worst repeat it (the calls are idempotent).
struct TtkPerDevice {
  bool done[64] = {};
  bool first() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};
