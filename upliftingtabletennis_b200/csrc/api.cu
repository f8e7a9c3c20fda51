// libttk: version, error state, device probe, host staging copy.
#include "ttk_internal.h"

#include <string.h>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

static thread_local char g_err[512] = "";

void ttk_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int ttk_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      sms = 148;
  }
  return sms;
}

extern "C" int ttk_version(void) { return TTK_VERSION; }
extern "C" const char* ttk_last_error(void) { return g_err; }

extern "C" int ttk_device_ok(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return 0;
  }
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

// Host copy of one frame into a pinned staging slot with non-temporal stores: the slot is read next by the DMA engine, not by a
// core, so the stores bypass the cache and skip the read-for-ownership of the destination lines (2 instead of 3 memory transfers
// per byte).  Called from Python worker threads through ctypes (which drops the GIL).
extern "C" int ttk_host_copy_stream(void* dst, const void* src, size_t bytes) {
  if (!dst || !src) {
    ttk_set_error("ttk_host_copy_stream: null pointer");
    return TTK_ERR_ARG;
  }
#if defined(__x86_64__)
  if (((uintptr_t)dst & 15) == 0 && bytes >= 4096) {
    const __m128i* s = (const __m128i*)src;
    __m128i* d = (__m128i*)dst;
    const size_t lines = bytes / 64;
    for (size_t i = 0; i < lines; ++i) {
      const __m128i a = _mm_loadu_si128(s + 4 * i), b = _mm_loadu_si128(s + 4 * i + 1), c = _mm_loadu_si128(s + 4 * i + 2),
                    e = _mm_loadu_si128(s + 4 * i + 3);
      _mm_stream_si128(d + 4 * i, a);
      _mm_stream_si128(d + 4 * i + 1, b);
      _mm_stream_si128(d + 4 * i + 2, c);
      _mm_stream_si128(d + 4 * i + 3, e);
    }
    _mm_sfence();
    memcpy((char*)dst + lines * 64, (const char*)src + lines * 64, bytes - lines * 64);
    return TTK_OK;
  }
#endif
  memcpy(dst, src, bytes);
  return TTK_OK;
}

// Make host threads that wait for this device (stream / event synchronisation, blocking copies) sleep instead of spin.  CUDA's default
// spins when there are no more contexts than cores; with one rank per GPU on a host that has two cores per GPU, every rank's waiting
// threads then burn the cores its frame-staging threads need (bench.py at eight ranks: DESIGN.md section 5).  Process-global for the
// device; returns the previous flags in *old_flags (may be null).
extern "C" int ttk_host_blocking_sync(int device, unsigned* old_flags) {
  int cur = 0;
  TTK_CUDA(cudaGetDevice(&cur));
  TTK_CUDA(cudaSetDevice(device));
  unsigned flags = 0;
  TTK_CUDA(cudaGetDeviceFlags(&flags));
  if (old_flags) *old_flags = flags;
  const cudaError_t e = cudaSetDeviceFlags((flags & ~(unsigned)cudaDeviceScheduleMask) | cudaDeviceScheduleBlockingSync);
  cudaSetDevice(cur);
  if (e != cudaSuccess) {
    cudaGetLastError();
    ttk_set_error("ttk_host_blocking_sync: cudaSetDeviceFlags -> %s", cudaGetErrorString(e));
    return TTK_ERR_CUDA;
  }
  return TTK_OK;
}
