// Heatmap decode: argmax (first maximum, NaN counts as maximum, like torch.argmax) ->
// 3x3 zero-padded window -> bounded Gaussian fit (float64 L-BFGS-B that follows SciPy's iteration, lbfgsb4.h)
// -> heatmap->image rescale.
// Reference: tabledetection/helper_tabledetection.py:50-156, balldetection/helper_balldetection.py:29-110.
//
// HBM-bound: every heatmap value is read exactly once with 128-bit loads
// (H*W*4 bytes per map); the fit is O(10^3) flops per map.
#define TTK_WARP_COOPERATIVE_LOSS 1     // the fit runs warp-wide: 9 lanes evaluate the 9 window terms
#include "lbfgsb4.h"
#include "ttk_internal.h"

namespace {

struct Best {
  float v;
  int i;
};

// true when (v,i) beats (bv,bi): larger value, NaN above everything, lower index on ties
__device__ __forceinline__ bool beats(float v, int i, float bv, int bi) {
  const bool vn = v != v, bn = bv != bv;
  if (vn != bn) return vn;
  if (vn) return i < bi;
  return v > bv || (v == bv && i < bi);
}

__device__ __forceinline__ Best warp_best(Best b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, b.v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, b.i, o);
    if (beats(ov, oi, b.v, b.i)) {
      b.v = ov;
      b.i = oi;
    }
  }
  return b;
}

__device__ __forceinline__ Best block_best(Best b, Best* smem) {
  b = warp_best(b);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) smem[warp] = b;
  __syncthreads();
  const int nw = blockDim.x >> 5;
  if (warp == 0) {
    Best c;
    c.v = lane < nw ? smem[lane].v : -INFINITY;
    c.i = lane < nw ? smem[lane].i : 0x7fffffff;
    b = warp_best(c);
  }
  return b;   // valid in warp 0
}

__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// grid (chunks, n_maps); each block scans one contiguous chunk of one map.
template <bool VEC>
__global__ void __launch_bounds__(256) argmax_partial_kernel(const float* __restrict__ maps, int hw, int chunk_elems,
                                                             Best* __restrict__ partial) {
  __shared__ Best smem[8];
  const int map = blockIdx.y;
  const float* base = maps + (size_t)map * hw;
  const int begin = blockIdx.x * chunk_elems;
  const int end = min(hw, begin + chunk_elems);
  Best b;
  b.v = -INFINITY;
  b.i = 0x7fffffff;
  if (VEC) {
    const float4* p4 = reinterpret_cast<const float4*>(base);
    const int b4 = begin >> 2, e4 = end >> 2;    // chunk_elems is a multiple of 4, hw % 4 == 0
    int i = b4 + threadIdx.x;
    // 4 independent 128-bit loads in flight per thread
    for (; i + 3 * 256 < e4; i += 4 * 256) {
      float4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) q[u] = ld_stream(p4 + i + u * 256);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = (i + u * 256) << 2;
        if (beats(q[u].x, e, b.v, b.i)) { b.v = q[u].x; b.i = e; }
        if (beats(q[u].y, e + 1, b.v, b.i)) { b.v = q[u].y; b.i = e + 1; }
        if (beats(q[u].z, e + 2, b.v, b.i)) { b.v = q[u].z; b.i = e + 2; }
        if (beats(q[u].w, e + 3, b.v, b.i)) { b.v = q[u].w; b.i = e + 3; }
      }
    }
    for (; i < e4; i += 256) {
      const float4 q = ld_stream(p4 + i);
      const int e = i << 2;
      if (beats(q.x, e, b.v, b.i)) { b.v = q.x; b.i = e; }
      if (beats(q.y, e + 1, b.v, b.i)) { b.v = q.y; b.i = e + 1; }
      if (beats(q.z, e + 2, b.v, b.i)) { b.v = q.z; b.i = e + 2; }
      if (beats(q.w, e + 3, b.v, b.i)) { b.v = q.w; b.i = e + 3; }
    }
  } else {
    for (int i = begin + threadIdx.x; i < end; i += 256) {
      const float v = __ldg(base + i);
      if (beats(v, i, b.v, b.i)) { b.v = v; b.i = i; }
    }
  }
  b = block_best(b, smem);
  if (threadIdx.x == 0) partial[(size_t)map * gridDim.x + blockIdx.x] = b;
}

// one warp per map: reduce the partials, cut the window, fit, rescale.
__global__ void __launch_bounds__(32) decode_finalize_kernel(const float* __restrict__ maps, int H, int W, int chunks,
                                                             const Best* __restrict__ partial, int variant, double scale_x,
                                                             double scale_y, double* __restrict__ out_xyv,
                                                             int32_t* __restrict__ out_idx, float* __restrict__ out_win) {
  const int map = blockIdx.x;
  Best b;
  b.v = -INFINITY;
  b.i = 0x7fffffff;
  for (int c = threadIdx.x; c < chunks; c += 32) {
    const Best p = partial[(size_t)map * chunks + c];
    if (beats(p.v, p.i, b.v, b.i)) b = p;
  }
  b = warp_best(b);          // every lane holds the winner
  const int idx = b.i;
  const int y = idx / W, x = idx - y * W;
  const float* m = maps + (size_t)map * H * W;
  float wf[9];
  double w[9];
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) {
      const int yy = y - 1 + j, xx = x - 1 + i;
      const float v = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? m[(size_t)yy * W + xx] : 0.f;
      wf[j * 3 + i] = v;
      w[j * 3 + i] = (double)v;
    }
  const TtkLbfgsbResult fit = ttk_lbfgsb_gauss(w, variant);
  double xi, yi;
  if (fit.success) {
    const double xs = (double)(x - 1) + fit.x[0];
    const double ys = (double)(y - 1) + fit.x[1];
    xi = (xs + 0.5) * scale_x - 0.5;
    yi = (ys + 0.5) * scale_y - 0.5;
  } else {
    // reference fallback (helper_tabledetection.py:131-134): mean index of the window maxima,
    // carried in float32 like numpy does for a float32 0-d array combined with Python floats.
    float mx = wf[0];
    bool has_nan = false;
    for (int k = 0; k < 9; ++k) {
      if (wf[k] != wf[k]) has_nan = true;
      mx = fmaxf(mx, wf[k]);
    }
    double sx = 0, sy = 0;
    int n = 0;
    if (!has_nan)
      for (int k = 0; k < 9; ++k)
        if (wf[k] == mx) {
          sx += k % 3;
          sy += k / 3;
          ++n;
        }
    if (n == 0) {
      xi = yi = __longlong_as_double(0x7ff8000000000000LL);
    } else {
      const float xs = __fadd_rn((float)(x - 1), (float)(sx / n));
      const float ys = __fadd_rn((float)(y - 1), (float)(sy / n));
      xi = (double)__fsub_rn(__fmul_rn(__fadd_rn(xs, 0.5f), (float)scale_x), 0.5f);
      yi = (double)__fsub_rn(__fmul_rn(__fadd_rn(ys, 0.5f), (float)scale_y), 0.5f);
    }
  }
  if (threadIdx.x != 0) return;
  out_xyv[map * 3 + 0] = xi;
  out_xyv[map * 3 + 1] = yi;
  out_xyv[map * 3 + 2] = 1.0;   // visibility is always 1 (helper_tabledetection.py:142)
  if (out_idx) out_idx[map] = idx;
  if (out_win)
    for (int k = 0; k < 9; ++k) out_win[map * 9 + k] = wf[k];
}

int decode_chunks(int hw) {
  // ~16 float4 per thread per block of 256 threads, at most 1024 chunks
  long long per = 256LL * 16 * 4;
  long long c = (hw + per - 1) / per;
  if (c < 1) c = 1;
  if (c > 1024) c = 1024;
  return (int)c;
}

}  // namespace

namespace {
bool g_profile = false;
cudaEvent_t g_ev[3] = {nullptr, nullptr, nullptr};
}  // namespace

extern "C" int ttk_decode_set_profile(int enable) {
  if (enable && !g_ev[0])
    for (cudaEvent_t& e : g_ev) TTK_CUDA(cudaEventCreate(&e));
  g_profile = enable != 0;
  return TTK_OK;
}

extern "C" int ttk_decode_profile_read(float* argmax_ms, float* fit_ms) {
  TTK_CHECK_ARG(argmax_ms && fit_ms, "ttk_decode_profile_read: null pointer");
  TTK_CHECK_ARG(g_ev[0], "ttk_decode_profile_read: profiling was never enabled");
  TTK_CUDA(cudaEventSynchronize(g_ev[2]));
  TTK_CUDA(cudaEventElapsedTime(argmax_ms, g_ev[0], g_ev[1]));
  TTK_CUDA(cudaEventElapsedTime(fit_ms, g_ev[1], g_ev[2]));
  return TTK_OK;
}

extern "C" size_t ttk_decode_workspace_bytes(int n_maps, int height, int width) {
  if (n_maps <= 0 || height <= 0 || width <= 0) return 0;
  return (size_t)n_maps * decode_chunks(height * width) * sizeof(Best);
}

extern "C" int ttk_heatmap_decode(const float* heatmaps_dev, int n_maps, int height, int width, int variant, int image_width,
                                  int image_height, double* out_xyv_dev, int32_t* out_idx_dev, float* out_win_dev,
                                  void* workspace_dev, size_t workspace_bytes, void* stream) {
  TTK_CHECK_ARG(n_maps >= 0 && height > 0 && width > 0, "ttk_heatmap_decode: bad sizes");
  TTK_CHECK_ARG((long long)height * width < (1LL << 31), "ttk_heatmap_decode: heatmap too large");
  TTK_CHECK_ARG(variant == TTK_DECODE_TABLE || variant == TTK_DECODE_BALL, "ttk_heatmap_decode: bad variant %d", variant);
  if (n_maps == 0) return TTK_OK;
  TTK_CHECK_ARG(heatmaps_dev && out_xyv_dev && workspace_dev, "ttk_heatmap_decode: null pointer");
  TTK_CHECK_ARG(n_maps <= 65535, "ttk_heatmap_decode: at most 65535 maps per call");
  const int hw = height * width;
  const int chunks = decode_chunks(hw);
  TTK_CHECK_ARG(workspace_bytes >= (size_t)n_maps * chunks * sizeof(Best), "ttk_heatmap_decode: workspace too small");
  int chunk_elems = (hw + chunks - 1) / chunks;
  chunk_elems = (chunk_elems + 3) & ~3;
  cudaStream_t st = (cudaStream_t)stream;
  Best* partial = (Best*)workspace_dev;
  const bool vec = (hw % 4 == 0) && (((uintptr_t)heatmaps_dev & 15) == 0);
  dim3 grid(chunks, n_maps);
  if (g_profile) TTK_CUDA(cudaEventRecord(g_ev[0], st));
  if (vec)
    argmax_partial_kernel<true><<<grid, 256, 0, st>>>(heatmaps_dev, hw, chunk_elems, partial);
  else
    argmax_partial_kernel<false><<<grid, 256, 0, st>>>(heatmaps_dev, hw, chunk_elems, partial);
  TTK_LAUNCH_CHECK();
  if (g_profile) TTK_CUDA(cudaEventRecord(g_ev[1], st));
  decode_finalize_kernel<<<n_maps, 32, 0, st>>>(heatmaps_dev, height, width, chunks, partial, variant,
                                                (double)image_width / width, (double)image_height / height, out_xyv_dev,
                                                out_idx_dev, out_win_dev);
  TTK_LAUNCH_CHECK();
  if (g_profile) TTK_CUDA(cudaEventRecord(g_ev[2], st));
  return TTK_OK;
}
