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
