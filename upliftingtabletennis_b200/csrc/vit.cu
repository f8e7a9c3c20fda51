// ViTPose-small heatmap detector (SURVEY.md section 8 row a4' / 8f row 4).
//   VitPose.forward            balldetection/models/vitpose.py:92-103 (tabledetection/models/vitpose.py for 3 -> 13 channels)
//   ViT.forward                vit_pose/vit_models/backbone/vit.py:375-389; PatchEmbed :208-228; Block :182-205; Attention :143-180
//   TopdownHeatmapSimpleHead   vit_pose/vit_models/head/topdown_heatmap_simple_head.py:188-193, 291-321
// Plan per sub-batch of images (tokens T = images x Hp x Wp, residual stream X float32 [T][384]):
//   im2col(16x16, pad 2) -> GEMM(+bias +pos) -> 12 x [LN -> GEMM qkv -> attention -> GEMM proj (+X) -> LN -> GEMM fc1 GELU ->
//   GEMM fc2 (+X)] -> LN -> 2 x [4 parity gathers -> 4 GEMMs (BN folded, ReLU) scattered to the 2x grid] -> 1x1 conv.
// Every Linear / conv / transposed conv is one GEMM C = act(A W^T + b) (+ R).  dtype TTK_F32 runs the float32 SIMT kernels
// of this file (parity with the CPU reference); TTK_BF16 runs the same plan with bf16 operands on tcgen05 tensor cores
// (gemm_umma.cu, attn_umma.cu), float32 accumulation, float32 residual stream.  TTK_TF32X3 is the reference-precision tensor-core
// path: every GEMM / attention operand is a split float32 pair (hi = tf32(x), lo = tf32(x - hi), written by the producing kernel's
// epilogue) and every product three kind::tf32 MMAs (gemm_umma.cu X3 mode, attn3_umma.cu) -- float32-class results at ~1/3 of the
// TF32 tensor rate instead of the SIMT rate.
#include "vit.h"

#include <math.h>
#include <string.h>

#include <algorithm>

namespace {

constexpr int DIM = 384, DEPTH = 12, HEADS = 12, HD = 32, MLP = 1536, PATCH = 16, PAD = 2, DEC = 256;

// ---- im2col of the patch embedding: A[t][c*256 + ky*16 + kx] = x[b][c][ty*16 - 2 + ky][tx*16 - 2 + kx] ----------------
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  // integer form of cvt.rna.tf32.f32 (round to nearest, ties away) without its special-case handling: 2 ALU ops; lo = x - hi is exact
  // (13 significant bits) and is left unrounded -- kind::tf32 ignores its low bits (2^-22 of x)
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  lo = x - hi;
}
// store four float32 values, or (lo_plane != 0) their tf32 hi / lo split into two planes lo_plane bytes apart
__device__ __forceinline__ void store4(float* o, size_t lo_plane, float a, float b, float c, float d) {
  if (lo_plane == 0) {
    *reinterpret_cast<float4*>(o) = make_float4(a, b, c, d);
    return;
  }
  float h[4], l[4];
  split_tf32(a, h[0], l[0]), split_tf32(b, h[1], l[1]), split_tf32(c, h[2], l[2]), split_tf32(d, h[3], l[3]);
  *reinterpret_cast<float4*>(o) = make_float4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<float4*>(reinterpret_cast<char*>(o) + lo_plane) = make_float4(l[0], l[1], l[2], l[3]);
}

template <typename T>
__global__ void patch_im2col_kernel(const float* __restrict__ x, int images, int C, int H, int W, int hp, int wp, T* __restrict__ A,
                                    size_t lo_plane = 0) {
  // one warp per (token, channel): lane = (ky, half row) reads 8 consecutive pixels and writes 8 consecutive elements, so a warp
  // reads 16 segments of 64 bytes and writes one contiguous run of 256 elements
  const int lane = threadIdx.x & 31, ky = lane >> 1, hx = (lane & 1) * 8;
  const long long total = (long long)images * hp * wp * C;
  for (long long i = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5); i < total; i += (long long)gridDim.x * (blockDim.x >> 5)) {
    const int c = (int)(i % C);
    const long long t = i / C;
    const int tx = (int)(t % wp), ty = (int)((t / wp) % hp), b = (int)(t / ((long long)wp * hp));
    const int y = ty * PATCH - PAD + ky, x0 = tx * PATCH - PAD + hx;
    float v[8];
    const bool row_ok = y >= 0 && y < H;
    const float* row = x + (((size_t)b * C + c) * H + (row_ok ? y : 0)) * W;
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (row_ok && x0 + k >= 0 && x0 + k < W) ? __ldg(row + x0 + k) : 0.f;
    T* o = A + (size_t)t * (C * PATCH * PATCH) + (size_t)c * PATCH * PATCH + ky * PATCH + hx;
    if constexpr (sizeof(T) == 4) {
      store4(o, lo_plane, v[0], v[1], v[2], v[3]);
      store4(o + 4, lo_plane, v[4], v[5], v[6], v[7]);
    } else {
      uint32_t pk[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
        pk[k] = *reinterpret_cast<uint32_t*>(&b2);
      }
      *reinterpret_cast<uint4*>(o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
}

// ---- LayerNorm(eps 1e-6) over 384 channels, one warp per token ------------------------------------------------------------
template <typename T>
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, int rows,
                                 T* __restrict__ out, size_t lo_plane = 0) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* p = reinterpret_cast<const float4*>(x + (size_t)row * DIM);
  float4 v[3];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    v[i] = p[lane + 32 * i];
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mu = s * (1.f / DIM);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    v[i].x -= mu, v[i].y -= mu, v[i].z -= mu, v[i].w -= mu;
    q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = 1.f / sqrtf(q * (1.f / DIM) + 1e-6f);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int c = (lane + 32 * i) * 4;
    const float4 ww = *reinterpret_cast<const float4*>(w + c), bb = *reinterpret_cast<const float4*>(b + c);
    const float o0 = v[i].x * rstd * ww.x + bb.x, o1 = v[i].y * rstd * ww.y + bb.y, o2 = v[i].z * rstd * ww.z + bb.z,
                o3 = v[i].w * rstd * ww.w + bb.w;
    if constexpr (sizeof(T) == 4) {
      store4(reinterpret_cast<float*>(out) + (size_t)row * DIM + c, lo_plane, o0, o1, o2, o3);
    } else {
      __nv_bfloat162 a = __floats2bfloat162_rn(o0, o1), d = __floats2bfloat162_rn(o2, o3);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&a);
      pk.y = *reinterpret_cast<uint32_t*>(&d);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + (size_t)row * DIM + c) = pk;
    }
  }
}

// ---- float32 SIMT GEMM (parity path): 64 x 64 tile, K step 16, 4 x 4 outputs per thread ---------------------------------
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ size_t out_row(const GemmArgs& g, int m) {
  if (g.up_w == 0) return (size_t)m;
  const int x = m % g.up_w, y = (m / g.up_w) % g.up_h, img = m / (g.up_w * g.up_h);
  return ((size_t)img * 2 * g.up_h + 2 * y + g.py) * 2 * g.up_w + 2 * x + g.px;
}

__global__ void __launch_bounds__(256) gemm_f32_kernel(GemmArgs g) {
  __shared__ float As[16][64 + 4], Ws[16][64 + 4];
  const float* A = (const float*)g.A;
  const float* W = (const float*)g.W;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tid = threadIdx.x, tr = tid / 16, tc = tid % 16;
  const int lr = tid / 4, lk = (tid % 4) * 4;          // loader: row lr (0..63), k offset lk
  float acc[4][4] = {};
  for (int k0 = 0; k0 < g.K; k0 += 16) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m0 + lr < g.M) a = *reinterpret_cast<const float4*>(A + (size_t)(m0 + lr) * g.K + k0 + lk);
    const float4 w = *reinterpret_cast<const float4*>(W + (size_t)(n0 + lr) * g.K + k0 + lk);
    As[lk][lr] = a.x, As[lk + 1][lr] = a.y, As[lk + 2][lr] = a.z, As[lk + 3][lr] = a.w;
    Ws[lk][lr] = w.x, Ws[lk + 1][lr] = w.y, Ws[lk + 2][lr] = w.z, Ws[lk + 3][lr] = w.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][tr * 4]);
      const float4 wv = *reinterpret_cast<const float4*>(&Ws[k][tc * 4]);
      const float a4[4] = {av.x, av.y, av.z, av.w}, w4[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], w4[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + tr * 4 + i;
    if (m >= g.M) continue;
    const int n = n0 + tc * 4;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = acc[i][j] + (g.bias ? g.bias[n + j] : 0.f);
      if (g.act == VIT_ACT_GELU) v[j] = gelu_erf(v[j]);
      if (g.act == VIT_ACT_RELU) v[j] = fmaxf(v[j], 0.f);
    }
    if (g.R) {
      const float4 r = *reinterpret_cast<const float4*>(g.R + (size_t)(g.r_mod ? m % g.r_mod : m) * g.N + n);
      v[0] += r.x, v[1] += r.y, v[2] += r.z, v[3] += r.w;
    }
    *reinterpret_cast<float4*>((float*)g.C + out_row(g, m) * g.N + n) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// ---- float32 attention (parity path): one warp per query, keys strided over lanes, online softmax, warp merge -------------
__global__ void __launch_bounds__(256) attention_f32_kernel(const float* __restrict__ qkv, int tokens, float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_idx = blockIdx.x * 8 + warp, head = blockIdx.y, img = blockIdx.z;
  if (q_idx >= tokens) return;
  const size_t base = (size_t)img * tokens;
  const float scale = 0.17677669529663687f;        // 32 ** -0.5, applied to q before q k^T (vit.py:168)
  float q[HD];
  {
    const float4* qp = reinterpret_cast<const float4*>(qkv + (base + q_idx) * (3 * DIM) + head * HD);
#pragma unroll
    for (int i = 0; i < HD / 4; ++i) {
      const float4 t = qp[i];
      q[4 * i] = t.x * scale, q[4 * i + 1] = t.y * scale, q[4 * i + 2] = t.z * scale, q[4 * i + 3] = t.w * scale;
    }
  }
  float m = -INFINITY, l = 0.f, o[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) o[d] = 0.f;
  for (int j = lane; j < tokens; j += 32) {
    const float4* kp = reinterpret_cast<const float4*>(qkv + (base + j) * (3 * DIM) + DIM + head * HD);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < HD / 4; ++i) {
      const float4 t = kp[i];
      s = fmaf(q[4 * i], t.x, s), s = fmaf(q[4 * i + 1], t.y, s), s = fmaf(q[4 * i + 2], t.z, s), s = fmaf(q[4 * i + 3], t.w, s);
    }
    const float mn = fmaxf(m, s), c = expf(m - mn), p = expf(s - mn);
    l = l * c + p;
    const float4* vp = reinterpret_cast<const float4*>(qkv + (base + j) * (3 * DIM) + 2 * DIM + head * HD);
#pragma unroll
    for (int i = 0; i < HD / 4; ++i) {
      const float4 t = vp[i];
      o[4 * i] = o[4 * i] * c + p * t.x, o[4 * i + 1] = o[4 * i + 1] * c + p * t.y;
      o[4 * i + 2] = o[4 * i + 2] * c + p * t.z, o[4 * i + 3] = o[4 * i + 3] * c + p * t.w;
    }
    m = mn;
  }
  float M = m;
#pragma unroll
  for (int s = 16; s; s >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, s));
  const float f = (m == -INFINITY) ? 0.f : expf(m - M);
  l *= f;
#pragma unroll
  for (int s = 16; s; s >>= 1) l += __shfl_xor_sync(0xffffffffu, l, s);
  float mine = 0.f;
#pragma unroll
  for (int d = 0; d < HD; ++d) {
    float v = o[d] * f;
#pragma unroll
    for (int s = 16; s; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if (lane == d) mine = v;
  }
  out[(base + q_idx) * DIM + head * HD + lane] = mine / l;
}

// ---- gather for one output parity of ConvTranspose2d(4, stride 2, padding 1): A[m][tap*Cin + c] ---------------------------
// py = 0: taps (dy, ky) = (0, 1), (-1, 3); py = 1: (0, 2), (1, 0) (same along x); weights are packed in that tap order.
template <typename T>
__global__ void deconv_gather_kernel(const T* __restrict__ in, int images, int H, int W, int Cin, int py, int px, T* __restrict__ A) {
  const int chunks = Cin / 8;
  const long long total = (long long)images * H * W * 4 * chunks;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % chunks);
    const int tap = (int)((i / chunks) % 4);
    const long long m = i / (4LL * chunks);
    const int x = (int)(m % W), y = (int)((m / W) % H), b = (int)(m / ((long long)W * H));
    const int dy = (tap >> 1) == 0 ? 0 : (py ? 1 : -1), dx = (tap & 1) == 0 ? 0 : (px ? 1 : -1);
    const int yy = y + dy, xx = x + dx;
    T* o = A + (size_t)m * (4 * Cin) + (size_t)tap * Cin + ch * 8;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) {
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = (T)0.f;
    } else {
      const T* s = in + (((size_t)b * H + yy) * W + xx) * Cin + ch * 8;
      if constexpr (sizeof(T) == 4) {
        reinterpret_cast<float4*>(o)[0] = reinterpret_cast<const float4*>(s)[0];
        reinterpret_cast<float4*>(o)[1] = reinterpret_cast<const float4*>(s)[1];
      } else {
        *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(s);
      }
    }
  }
}

// ---- final 1x1 conv: [P][256] -> planar heatmaps (images, out_ch, H, W) float32, one warp per pixel ------------------------
template <typename T>
__global__ void final_conv_kernel(const T* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias, int images,
                                  int hw, int out_ch, float* __restrict__ heat) {
  const long long p = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (p >= (long long)images * hw) return;
  float v[8];
  if constexpr (sizeof(T) == 4) {
    const float4 a = reinterpret_cast<const float4*>(in + (size_t)p * DEC)[lane * 2], b = reinterpret_cast<const float4*>(in + (size_t)p * DEC)[lane * 2 + 1];
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
  } else {
    const uint4 u = reinterpret_cast<const uint4*>(in + (size_t)p * DEC)[lane];
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x, v[2 * i + 1] = f.y;
    }
  }
  const int img = (int)(p / hw), pix = (int)(p % hw);
  for (int o = 0; o < out_ch; ++o) {
    const float4 w0 = reinterpret_cast<const float4*>(w + (size_t)o * DEC)[lane * 2], w1 = reinterpret_cast<const float4*>(w + (size_t)o * DEC)[lane * 2 + 1];
    float s = v[0] * w0.x + v[1] * w0.y + v[2] * w0.z + v[3] * w0.w + v[4] * w1.x + v[5] * w1.y + v[6] * w1.z + v[7] * w1.w;
#pragma unroll
    for (int k = 16; k; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
    if (lane == 0) heat[((size_t)img * out_ch + o) * hw + pix] = s + bias[o];
  }
}

__global__ void split_pool_kernel(const float* __restrict__ in, float* __restrict__ hi, float* __restrict__ lo, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) split_tf32(in[i], hi[i], lo[i]);
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = __float2bfloat16(in[i]);
}

std::string fmt(const char* f, int i) {
  char b[160];
  snprintf(b, sizeof b, f, i);
  return b;
}

void add_param(ttk_vit* h, const std::string& name, std::vector<int> shape) {
  VitParam p;
  p.name = name;
  p.shape = shape;
  p.numel = 1;
  for (int s : shape) p.numel *= (size_t)s;
  h->params.push_back(std::move(p));
}

int launch_gemm(ttk_vit* h, const GemmArgs& g, int dtype, cudaStream_t st) {
  ++h->launches;
  if (dtype == TTK_BF16 || dtype == TTK_TF32X3) return ttk_gemm_umma(g, st);
  gemm_f32_kernel<<<dim3(g.N / 64, ttk_cdiv(g.M, 64)), 256, 0, st>>>(g);
  TTK_LAUNCH_CHECK();
  return TTK_OK;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Workspace {              // slices for one sub-batch of `n` images
  float* X;                     // [T][384] float32 residual stream
  char *A, *QKV, *ATT, *HID;    // operand buffers in the path's dtype
  char *D1, *D2;                // deconv outputs NHWC
  size_t pA, pQKV, pATT, pHID, pD1;      // TTK_TF32X3: byte distance between the hi and the lo plane of each operand buffer
  size_t total;
};

Workspace carve(const ttk_vit* h, char* base, int n, int dtype) {
  const bool x3 = dtype == TTK_TF32X3;
  const size_t es = dtype == TTK_BF16 ? 2 : 4, planes = x3 ? 2 : 1;
  const size_t T = (size_t)n * h->tokens;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += align_up(bytes, 1024);
    return p;
  };
  Workspace w;
  w.X = (float*)take(T * DIM * 4);
  // A: LN output [T][384], patch im2col [T][in_ch*256], deconv gathers [T][4*384] and [4T][4*256]
  size_t a_elems = T * (size_t)std::max(std::max(DIM, h->in_ch * PATCH * PATCH), std::max(4 * DIM, 4 * 4 * DEC));
  w.pA = a_elems * es, w.pQKV = T * 3 * DIM * es, w.pATT = T * DIM * es, w.pHID = T * MLP * es, w.pD1 = T * 4 * DEC * es;
  w.A = take(w.pA * planes);
  w.QKV = take(w.pQKV * planes);
  w.ATT = take(w.pATT * planes);
  w.HID = take(w.pHID * planes);
  w.D1 = take(w.pD1 * planes);
  w.D2 = take(T * 16 * DEC * es);
  w.total = off;
  return w;
}

}  // namespace

extern "C" int ttk_vit_create(int in_ch, int out_ch, int height, int width, ttk_vit** out) {
  TTK_CHECK_ARG(out, "ttk_vit_create: null out");
  TTK_CHECK_ARG(in_ch > 0 && in_ch <= 16 && out_ch > 0 && out_ch <= 64, "ttk_vit_create: bad channel counts %d -> %d", in_ch, out_ch);
  TTK_CHECK_ARG(height >= PATCH && width >= PATCH, "ttk_vit_create: bad resolution %d x %d", width, height);
  ttk_vit* h = new ttk_vit();
  h->in_ch = in_ch;
  h->out_ch = out_ch;
  h->height = height;
  h->width = width;
  h->hp = (height + 2 * PAD - PATCH) / PATCH + 1;
  h->wp = (width + 2 * PAD - PATCH) / PATCH + 1;
  h->tokens = h->hp * h->wp;
  // order of the reference's state_dict (oracle/vitpose.py:state_dict_layout)
  const std::string p = "model.backbone.";
  add_param(h, p + "pos_embed", {1, h->tokens + 1, DIM});
  add_param(h, p + "patch_embed.proj.weight", {DIM, in_ch, PATCH, PATCH});
  add_param(h, p + "patch_embed.proj.bias", {DIM});
  for (int i = 0; i < DEPTH; ++i) {
    const std::string b = p + fmt("blocks.%d.", i);
    add_param(h, b + "norm1.weight", {DIM});
    add_param(h, b + "norm1.bias", {DIM});
    add_param(h, b + "attn.qkv.weight", {3 * DIM, DIM});
    add_param(h, b + "attn.qkv.bias", {3 * DIM});
    add_param(h, b + "attn.proj.weight", {DIM, DIM});
    add_param(h, b + "attn.proj.bias", {DIM});
    add_param(h, b + "norm2.weight", {DIM});
    add_param(h, b + "norm2.bias", {DIM});
    add_param(h, b + "mlp.fc1.weight", {MLP, DIM});
    add_param(h, b + "mlp.fc1.bias", {MLP});
    add_param(h, b + "mlp.fc2.weight", {DIM, MLP});
    add_param(h, b + "mlp.fc2.bias", {DIM});
  }
  add_param(h, p + "last_norm.weight", {DIM});
  add_param(h, p + "last_norm.bias", {DIM});
  const std::string k = "model.keypoint_head.";
  int cin = DIM;
  for (int j : {0, 3}) {
    add_param(h, k + fmt("deconv_layers.%d.weight", j), {cin, DEC, 4, 4});
    const std::string bn = k + fmt("deconv_layers.%d.", j + 1);
    add_param(h, bn + "weight", {DEC});
    add_param(h, bn + "bias", {DEC});
    add_param(h, bn + "running_mean", {DEC});
    add_param(h, bn + "running_var", {DEC});
    cin = DEC;
  }
  add_param(h, k + "final_layer.weight", {out_ch, DEC, 1, 1});
  add_param(h, k + "final_layer.bias", {out_ch});
  *out = h;
  return TTK_OK;
}

extern "C" void ttk_vit_destroy(ttk_vit* h) {
  if (!h) return;
  cudaFree(h->f32_pool);
  cudaFree(h->bf16_pool);
  cudaFree(h->x3_pool);
  delete h;
}

extern "C" int ttk_vit_num_params(const ttk_vit* h) { return h ? (int)h->params.size() : 0; }

extern "C" int ttk_vit_param_info(const ttk_vit* h, int i, char* name, int* numel) {
  TTK_CHECK_ARG(h && i >= 0 && i < (int)h->params.size(), "ttk_vit_param_info: bad index %d", i);
  if (name) snprintf(name, 128, "%s", h->params[i].name.c_str());
  if (numel) *numel = (int)h->params[i].numel;
  return TTK_OK;
}

extern "C" int ttk_vit_set_param(ttk_vit* h, int i, const float* data_host, int numel) {
  TTK_CHECK_ARG(h && i >= 0 && i < (int)h->params.size(), "ttk_vit_set_param: bad index %d", i);
  VitParam& p = h->params[i];
  TTK_CHECK_ARG(data_host && (size_t)numel == p.numel, "ttk_vit_set_param: %s expects %zu elements, got %d", p.name.c_str(), p.numel, numel);
  p.host.assign(data_host, data_host + numel);
  p.set = true;
  h->ready = false;
  return TTK_OK;
}

extern "C" int ttk_vit_tokens(const ttk_vit* h, int* hp, int* wp) {
  TTK_CHECK_ARG(h, "ttk_vit_tokens: null handle");
  if (hp) *hp = h->hp;
  if (wp) *wp = h->wp;
  return h->tokens;
}

extern "C" int ttk_vit_set_subbatch(ttk_vit* h, int images) {
  TTK_CHECK_ARG(h && images >= 1 && images <= 64, "ttk_vit_set_subbatch: images must be in [1, 64]");
  h->subbatch = images;
  return TTK_OK;
}

extern "C" int ttk_vit_last_launches(const ttk_vit* h) { return h ? h->launches : 0; }

// Pack all GEMM operands into one float32 pool (+ a bf16 copy): BN of the head folded in float64, transposed-conv weights
// re-ordered per output parity to [Cout][tap][Cin], pos_embed[1:] + pos_embed[:1] pre-added.
static int vit_prepare(ttk_vit* h) {
  for (const VitParam& p : h->params)
    if (!p.set) {
      ttk_set_error("ttk_vit_forward: parameter %s was never set", p.name.c_str());
      return TTK_ERR_STATE;
    }
  std::vector<float> pool;
  auto push = [&](const float* d, size_t n) {
    const size_t off = pool.size();
    pool.insert(pool.end(), d, d + n);
    while (pool.size() % 64) pool.push_back(0.f);      // 256-byte alignment of every operand (TMA needs 16)
    return off;
  };
  auto host = [&](const std::string& n) -> const std::vector<float>& { return h->params[h->find(n)].host; };
  auto lin = [&](const std::string& wn, const std::string& bn, int n, int k) {
    ttk_vit::Lin l;
    l.w_off = push(host(wn).data(), (size_t)n * k);
    l.b_off = push(host(bn).data(), (size_t)n);
    l.n = n;
    l.k = k;
    return l;
  };
  const std::string p = "model.backbone.";
  h->patch = lin(p + "patch_embed.proj.weight", p + "patch_embed.proj.bias", DIM, h->in_ch * PATCH * PATCH);
  {
    const std::vector<float>& pe = host(p + "pos_embed");
    std::vector<float> pos((size_t)h->tokens * DIM);
    for (int t = 0; t < h->tokens; ++t)
      for (int c = 0; c < DIM; ++c) pos[(size_t)t * DIM + c] = (float)((double)pe[(size_t)(t + 1) * DIM + c] + (double)pe[c]);
    h->pos_off = push(pos.data(), pos.size());
  }
  h->blocks.resize(DEPTH);
  for (int i = 0; i < DEPTH; ++i) {
    const std::string b = p + fmt("blocks.%d.", i);
    ttk_vit::Block& B = h->blocks[i];
    B.ln1w = push(host(b + "norm1.weight").data(), DIM);
    B.ln1b = push(host(b + "norm1.bias").data(), DIM);
    B.ln2w = push(host(b + "norm2.weight").data(), DIM);
    B.ln2b = push(host(b + "norm2.bias").data(), DIM);
    B.qkv = lin(b + "attn.qkv.weight", b + "attn.qkv.bias", 3 * DIM, DIM);
    B.proj = lin(b + "attn.proj.weight", b + "attn.proj.bias", DIM, DIM);
    B.fc1 = lin(b + "mlp.fc1.weight", b + "mlp.fc1.bias", MLP, DIM);
    B.fc2 = lin(b + "mlp.fc2.weight", b + "mlp.fc2.bias", DIM, MLP);
  }
  h->lnfw = push(host(p + "last_norm.weight").data(), DIM);
  h->lnfb = push(host(p + "last_norm.bias").data(), DIM);
  const std::string k = "model.keypoint_head.";
  int cin = DIM;
  for (int layer = 0; layer < 2; ++layer) {
    const int j = layer * 3;
    const std::vector<float>& w = host(k + fmt("deconv_layers.%d.weight", j));        // [cin][DEC][4][4]
    const std::string bn = k + fmt("deconv_layers.%d.", j + 1);
    const std::vector<float>&g = host(bn + "weight"), &be = host(bn + "bias"), &mu = host(bn + "running_mean"), &var = host(bn + "running_var");
    std::vector<double> scale(DEC);
    std::vector<float> bias(DEC);
    for (int o = 0; o < DEC; ++o) {
      scale[o] = (double)g[o] / sqrt((double)var[o] + 1e-5);
      bias[o] = (float)((double)be[o] - (double)mu[o] * scale[o]);
    }
    const int kidx[2][2] = {{1, 3}, {2, 0}};           // kernel index of tap 0 / tap 1 for parity 0 / 1
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        std::vector<float> wp((size_t)DEC * 4 * cin);
        for (int o = 0; o < DEC; ++o)
          for (int ty = 0; ty < 2; ++ty)
            for (int tx = 0; tx < 2; ++tx)
              for (int c = 0; c < cin; ++c)
                wp[((size_t)o * 4 + ty * 2 + tx) * cin + c] =
                    (float)((double)w[(((size_t)c * DEC + o) * 4 + kidx[py][ty]) * 4 + kidx[px][tx]] * scale[o]);
        ttk_vit::Lin l;
        l.w_off = push(wp.data(), wp.size());
        l.b_off = push(bias.data(), DEC);
        l.n = DEC;
        l.k = 4 * cin;
        h->deconv[layer][py * 2 + px] = l;
      }
    cin = DEC;
  }
  h->final_w = push(host(k + "final_layer.weight").data(), (size_t)h->out_ch * DEC);
  h->final_b = push(host(k + "final_layer.bias").data(), h->out_ch);
  cudaFree(h->f32_pool);
  cudaFree(h->bf16_pool);
  cudaFree(h->x3_pool);
  h->f32_pool = nullptr;
  h->bf16_pool = nullptr;
  h->x3_pool = nullptr;
  h->pool_bytes = pool.size() * sizeof(float);
  TTK_CUDA(cudaMalloc((void**)&h->f32_pool, pool.size() * sizeof(float)));
  TTK_CUDA(cudaMalloc((void**)&h->bf16_pool, pool.size() * sizeof(__nv_bfloat16)));
  TTK_CUDA(cudaMemcpy(h->f32_pool, pool.data(), pool.size() * sizeof(float), cudaMemcpyHostToDevice));
  TTK_CUDA(cudaMalloc((void**)&h->x3_pool, 2 * pool.size() * sizeof(float)));
  split_pool_kernel<<<1024, 256>>>(h->f32_pool, h->x3_pool, h->x3_pool + pool.size(), pool.size());
  TTK_LAUNCH_CHECK();
  f32_to_bf16_kernel<<<1024, 256>>>(h->f32_pool, h->bf16_pool, pool.size());
  TTK_LAUNCH_CHECK();
  TTK_CUDA(cudaDeviceSynchronize());
  h->ready = true;
  return TTK_OK;
}

extern "C" size_t ttk_vit_workspace_bytes(const ttk_vit* h, int batch, int dtype) {
  if (!h || batch <= 0) return 0;
  return carve(h, nullptr, std::min(batch, h->subbatch), dtype).total;
}

extern "C" int ttk_vit_forward(ttk_vit* h, const float* x_dev, int batch, int dtype, float* heatmaps_dev, void* workspace_dev,
                               size_t workspace_bytes, void* stream) {
  TTK_CHECK_ARG(h, "ttk_vit_forward: null handle");
  if (int rc = ttk_bind_device(&h->device, "ttk_vit_forward")) return rc;
  TTK_CHECK_ARG(dtype == TTK_F32 || dtype == TTK_BF16 || dtype == TTK_TF32X3, "ttk_vit_forward: bad dtype %d", dtype);
  TTK_CHECK_ARG(batch >= 0, "ttk_vit_forward: bad batch");
  if (!h->ready) {
    const int rc = vit_prepare(h);
    if (rc != TTK_OK) return rc;
  }
  h->launches = 0;
  if (batch == 0) return TTK_OK;
  TTK_CHECK_ARG(x_dev && heatmaps_dev && workspace_dev, "ttk_vit_forward: null pointer");
  TTK_CHECK_ARG(workspace_bytes >= ttk_vit_workspace_bytes(h, batch, dtype), "ttk_vit_forward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const bool bf = dtype == TTK_BF16, x3 = dtype == TTK_TF32X3;
  auto wptr = [&](size_t off) -> const void* {
    return bf ? (const void*)(h->bf16_pool + off) : x3 ? (const void*)(h->x3_pool + off) : (const void*)(h->f32_pool + off);
  };
  auto fptr = [&](size_t off) { return (const float*)(h->f32_pool + off); };
  const int hw_out = 16 * h->tokens;

  for (int b0 = 0; b0 < batch; b0 += h->subbatch) {
    const int n = std::min(h->subbatch, batch - b0);
    const int T = n * h->tokens;
    Workspace w = carve(h, (char*)workspace_dev, std::min(batch, h->subbatch), dtype);
    const float* x = x_dev + (size_t)b0 * h->in_ch * h->height * h->width;
    // a_plane / c_plane: plane distance of a split operand (TTK_TF32X3; c_plane != 0 makes the output a split pair)
    auto gemm = [&](const void* A, const ttk_vit::Lin& l, const float* R, void* C, int M, int act, int c_bf16, int up_h = 0, int up_w = 0,
                    int py = 0, int px = 0, size_t a_plane = 0, size_t c_plane = 0) {
      GemmArgs g;
      g.A = A, g.W = wptr(l.w_off), g.bias = fptr(l.b_off), g.R = R, g.C = C, g.M = M, g.N = l.n, g.K = l.k, g.act = act, g.c_bf16 = c_bf16;
      g.up_h = up_h, g.up_w = up_w, g.py = py, g.px = px;
      if (x3) g.x3 = 1, g.a_plane = a_plane, g.w_plane = h->pool_bytes, g.c_split = c_plane != 0, g.c_plane = c_plane;
      return launch_gemm(h, g, dtype, st);
    };
    auto ln = [&](size_t wo, size_t bo) {
      ++h->launches;
      if (bf)
        layernorm_kernel<__nv_bfloat16><<<ttk_cdiv(T, 8), 256, 0, st>>>(w.X, fptr(wo), fptr(bo), T, (__nv_bfloat16*)w.A);
      else
        layernorm_kernel<float><<<ttk_cdiv(T, 8), 256, 0, st>>>(w.X, fptr(wo), fptr(bo), T, (float*)w.A, x3 ? w.pA : 0);
    };
    int rc;
    // patch embedding (+ bias + position): the position table repeats per image and enters as a residual read modulo the token count
    ++h->launches;
    if (bf)
      patch_im2col_kernel<__nv_bfloat16><<<ttk_num_sms() * 8, 256, 0, st>>>(x, n, h->in_ch, h->height, h->width, h->hp, h->wp, (__nv_bfloat16*)w.A);
    else
      patch_im2col_kernel<float><<<ttk_num_sms() * 8, 256, 0, st>>>(x, n, h->in_ch, h->height, h->width, h->hp, h->wp, (float*)w.A, x3 ? w.pA : 0);
    TTK_LAUNCH_CHECK();
    {
      GemmArgs g;
      g.A = w.A, g.W = wptr(h->patch.w_off), g.bias = fptr(h->patch.b_off), g.R = fptr(h->pos_off), g.r_mod = h->tokens, g.C = w.X;
      g.M = T, g.N = h->patch.n, g.K = h->patch.k, g.act = VIT_ACT_NONE, g.c_bf16 = 0, g.up_h = g.up_w = g.py = g.px = 0;
      if (x3) g.x3 = 1, g.a_plane = w.pA, g.w_plane = h->pool_bytes;
      if ((rc = launch_gemm(h, g, dtype, st)) != TTK_OK) return rc;
    }
    for (int i = 0; i < DEPTH; ++i) {
      const ttk_vit::Block& B = h->blocks[i];
      ln(B.ln1w, B.ln1b);
      if (x3) {
        // the V third of the output goes straight to the attention's transposed B operand (in HID, free until fc1)
        GemmArgs g;
        g.A = w.A, g.W = wptr(B.qkv.w_off), g.bias = fptr(B.qkv.b_off), g.R = nullptr, g.C = w.QKV, g.M = T, g.N = B.qkv.n, g.K = B.qkv.k;
        g.act = VIT_ACT_NONE, g.c_bf16 = 0, g.up_h = g.up_w = g.py = g.px = 0;
        g.x3 = 1, g.a_plane = w.pA, g.w_plane = h->pool_bytes, g.c_split = 1, g.c_plane = w.pQKV;
        const int tok_pad = (h->tokens + 7) / 8 * 8;
        g.vt = (float*)w.HID, g.vt_plane = (size_t)n * DIM * tok_pad * 4, g.vt_col0 = 2 * DIM, g.vt_tokens = h->tokens, g.vt_tok_pad = tok_pad;
        if ((rc = launch_gemm(h, g, dtype, st)) != TTK_OK) return rc;
      } else if ((rc = gemm(w.A, B.qkv, nullptr, w.QKV, T, VIT_ACT_NONE, bf)) != TTK_OK) return rc;
      ++h->launches;
      if (bf) {
        if ((rc = ttk_attention_umma((const __nv_bfloat16*)w.QKV, (__nv_bfloat16*)w.ATT, w.A, n, h->tokens, HEADS, HD, st)) != TTK_OK) return rc;
      } else if (x3) {
        if ((rc = ttk_attention3((const float*)w.QKV, w.pQKV, (float*)w.ATT, w.pATT, w.HID, n, h->tokens, HEADS, HD, st, 1)) != TTK_OK) return rc;
      } else {
        attention_f32_kernel<<<dim3(ttk_cdiv(h->tokens, 8), HEADS, n), 256, 0, st>>>((const float*)w.QKV, h->tokens, (float*)w.ATT);
      }
      if ((rc = gemm(w.ATT, B.proj, w.X, w.X, T, VIT_ACT_NONE, 0, 0, 0, 0, 0, w.pATT)) != TTK_OK) return rc;
      ln(B.ln2w, B.ln2b);
      if ((rc = gemm(w.A, B.fc1, nullptr, w.HID, T, VIT_ACT_GELU, bf, 0, 0, 0, 0, w.pA, x3 ? w.pHID : 0)) != TTK_OK) return rc;
      if ((rc = gemm(w.HID, B.fc2, w.X, w.X, T, VIT_ACT_NONE, 0, 0, 0, 0, 0, w.pHID)) != TTK_OK) return rc;
    }
    // last_norm: its output [T][384] is the NHWC feature map (n, hp, wp, 384) -- reuse ATT for it, A for the gathers
    ++h->launches;
    if (bf)
      layernorm_kernel<__nv_bfloat16><<<ttk_cdiv(T, 8), 256, 0, st>>>(w.X, fptr(h->lnfw), fptr(h->lnfb), T, (__nv_bfloat16*)w.ATT);
    else
      layernorm_kernel<float><<<ttk_cdiv(T, 8), 256, 0, st>>>(w.X, fptr(h->lnfw), fptr(h->lnfb), T, (float*)w.ATT, x3 ? w.pATT : 0);
    const char* feat = w.ATT;
    char* outs[2] = {w.D1, w.D2};
    size_t feat_plane = w.pATT;
    int fh = h->hp, fw = h->wp, cin = DIM;
    for (int layer = 0; layer < 2; ++layer) {
      const int M = n * fh * fw;
      for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
          if (bf || x3) {
            // tensor-core paths: TMA gathers the 2x2 taps straight from the feature map (implicit GEMM, no materialised gather)
            GemmArgs g;
            const ttk_vit::Lin& l = h->deconv[layer][py * 2 + px];
            g.A = feat, g.W = wptr(l.w_off), g.bias = fptr(l.b_off), g.R = nullptr, g.C = outs[layer], g.M = M, g.N = l.n, g.K = l.k;
            g.act = VIT_ACT_RELU, g.c_bf16 = bf, g.up_h = fh, g.up_w = fw, g.py = py, g.px = px, g.implicit_c = cin;
            // the first transposed convolution feeds the second (split pair), the second the float32 1x1 convolution
            if (x3) g.x3 = 1, g.a_plane = feat_plane, g.w_plane = h->pool_bytes, g.c_split = layer == 0, g.c_plane = layer == 0 ? w.pD1 : 0;
            if ((rc = launch_gemm(h, g, dtype, st)) != TTK_OK) return rc;
            continue;
          }
          ++h->launches;
          deconv_gather_kernel<float><<<ttk_num_sms() * 8, 256, 0, st>>>((const float*)feat, n, fh, fw, cin, py, px, (float*)w.A);
          if ((rc = gemm(w.A, h->deconv[layer][py * 2 + px], nullptr, outs[layer], M, VIT_ACT_RELU, bf, fh, fw, py, px)) != TTK_OK) return rc;
        }
      feat = outs[layer];
      feat_plane = w.pD1;
      fh *= 2, fw *= 2, cin = DEC;
    }
    ++h->launches;
    float* heat = heatmaps_dev + (size_t)b0 * h->out_ch * hw_out;
    if (bf)
      final_conv_kernel<__nv_bfloat16><<<ttk_cdiv((long long)n * hw_out, 8), 256, 0, st>>>((const __nv_bfloat16*)feat, fptr(h->final_w), fptr(h->final_b), n, hw_out, h->out_ch, heat);
    else
      final_conv_kernel<float><<<ttk_cdiv((long long)n * hw_out, 8), 256, 0, st>>>((const float*)feat, fptr(h->final_w), fptr(h->final_b), n, hw_out, h->out_ch, heat);
    TTK_LAUNCH_CHECK();
  }
  return TTK_OK;
}

// ---- test hooks: the two tensor-core kernels on caller buffers -----------------------------------------------------------------
extern "C" int ttk_vit_debug_gemm(const void* a_dev, const void* w_dev, const float* bias_dev, const float* res_dev, void* c_dev, int m,
                                  int n, int k, int act, int c_bf16, int up_h, int up_w, int py, int px, int dtype, void* stream) {
  TTK_CHECK_ARG(a_dev && w_dev && c_dev && m > 0 && n > 0 && k > 0, "ttk_vit_debug_gemm: bad arguments");
  GemmArgs g;
  g.A = a_dev, g.W = w_dev, g.bias = bias_dev, g.R = res_dev, g.C = c_dev, g.M = m, g.N = n, g.K = k, g.act = act, g.c_bf16 = c_bf16;
  g.up_h = up_h, g.up_w = up_w, g.py = py, g.px = px;
  if (dtype == TTK_TF32X3) {        // split pairs [2][m][k], [2][n][k]; c_bf16 == 2: C is a split pair [2][rows][n] too
    g.x3 = 1, g.a_plane = (size_t)m * k * 4, g.w_plane = (size_t)n * k * 4, g.c_split = c_bf16 == 2, g.c_bf16 = 0;
    g.c_plane = (size_t)(up_w ? 4 : 1) * m * n * 4;
    return ttk_gemm_umma(g, (cudaStream_t)stream);
  }
  if (dtype == TTK_BF16) return ttk_gemm_umma(g, (cudaStream_t)stream);
  TTK_CHECK_ARG(n % 64 == 0 && k % 16 == 0 && !c_bf16, "ttk_vit_debug_gemm: the float32 kernel needs N %% 64 == 0 and K %% 16 == 0");
  gemm_f32_kernel<<<dim3(n / 64, ttk_cdiv(m, 64)), 256, 0, (cudaStream_t)stream>>>(g);
  TTK_LAUNCH_CHECK();
  return TTK_OK;
}

extern "C" int ttk_vit_debug_attention(const void* qkv_dev, void* out_dev, void* scratch_dev, size_t scratch_bytes, int images, int tokens,
                                       int dtype, void* stream) {
  TTK_CHECK_ARG(qkv_dev && out_dev && images > 0 && tokens > 0, "ttk_vit_debug_attention: bad arguments");
  if (dtype == TTK_BF16) {
    TTK_CHECK_ARG(scratch_dev && scratch_bytes >= ttk_attention_umma_scratch_bytes(images, tokens, HEADS, HD), "ttk_vit_debug_attention: scratch too small");
    return ttk_attention_umma((const __nv_bfloat16*)qkv_dev, (__nv_bfloat16*)out_dev, scratch_dev, images, tokens, HEADS, HD, (cudaStream_t)stream);
  }
  if (dtype == TTK_TF32X3) {        // qkv [2][T][1152] / out [2][T][384] split pairs
    TTK_CHECK_ARG(scratch_dev && scratch_bytes >= ttk_attention3_scratch_bytes(images, tokens, HEADS, HD), "ttk_vit_debug_attention: scratch too small");
    const size_t T = (size_t)images * tokens;
    return ttk_attention3((const float*)qkv_dev, T * 3 * DIM * 4, (float*)out_dev, T * DIM * 4, scratch_dev, images, tokens, HEADS, HD,
                          (cudaStream_t)stream);
  }
  attention_f32_kernel<<<dim3(ttk_cdiv(tokens, 8), HEADS, images), 256, 0, (cudaStream_t)stream>>>((const float*)qkv_dev, tokens, (float*)out_dev);
  TTK_LAUNCH_CHECK();
  return TTK_OK;
}
