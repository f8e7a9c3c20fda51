// Multi-head self-attention of the ViTPose backbone with fp32-level results on the tensor cores (3xTF32), the reference-precision
// sibling of attn_umma.cu (same flash-attention schedule, same roles):
//   out = softmax((q * 32^-0.5) k^T) v per (image, head)      vit_pose/vit_models/backbone/vit.py:160-176 (Attention.forward)
// Every operand is a split float32 pair x = hi + lo (hi = tf32(x), lo = tf32(x - hi)); every product is three kind::tf32 MMAs
// (lo*hi + hi*lo + hi*hi, float32 accumulation in TMEM), which leaves ~2^-21 relative error per product -- float32 class.
// qkv [2][images * tokens][3 * 384] float32 planes as the x3 qkv GEMM wrote them; out [2][T][384] planes (the proj GEMM's A operand).
//
// One CTA = one (image, head, 128-query tile); it walks the key tiles (64 keys) once:
//   S   = Q K_j^T           3 x 4 tcgen05.mma 128 x 64 x 8; Q / K_j hi | lo tiles staged by TMA (128-byte rows = 32 floats), S in TMEM
//   P_j = exp2(S*c - m)     two threads per query row (32 keys each): scores to registers, S is then free for the next Q K^T; running
//                           maximum in registers, P_j split and written to TENSOR MEMORY (tcgen05.st; hi and lo, 64 columns each): the
//                           P V MMAs take their A operand from there, so P never crosses shared memory (a float32 P pair through
//                           shared memory was 128 KB of traffic per tile and made the kernel shared-memory bound, 2990 clk / tile)
//   O   = P_j V_j           per k8 step: P_hi x [V_hi ; V_lo] (N = 64, two accumulator blocks) and P_lo x V_hi (N = 32); V staged
//                           TRANSPOSED ([dim][key] hi / lo planes from a small transpose kernel).  Each key tile's
//                           P_j V_j is a FRESH TMEM accumulation that the softmax threads add to register accumulators
//                           (rescaled by 2^(m_old - m_new), round-to-nearest): the tensor core's float32 accumulate truncates, and
//                           a chain over all 45 key tiles (1080 MMAs) measured 1.8e-5 relative error -- per tile it is 24 MMAs.
// TMEM: S 64 + O 64 + P 128 = 256 columns and 98 KB of shared memory (Q pair + 2 K / V^T stages), so TWO CTAs share an SM and one
// CTA's softmax (exponentials, splits, TMEM round trips: a long dependent chain per tile) overlaps the other's MMAs.
#include "umma_prims.h"
#include "vit.h"

namespace {

using namespace umma;

constexpr int HD = 32, BQ = 128, BKEY = 64, NST = 2, THREADS = 320;      // 8 softmax warps, MMA warp, TMA warp
constexpr int Q_PLANE = BQ * 128, Q_BYTES = 2 * Q_PLANE;                 // hi | lo, 128-byte rows
constexpr int K_PLANE = BKEY * 128, K_BYTES = 2 * K_PLANE;
constexpr int VR = HD;                                 // rows of the staged V^T tile
constexpr int V_CHUNK = VR * 128;                      // one 32-key chunk [32 rows][128 B]
constexpr int V_PLANE = 2 * V_CHUNK, V_BYTES = 2 * V_PLANE;
constexpr int KV_BYTES = K_BYTES + V_BYTES;            // 32 KB
constexpr int X_BYTES = 4 * BQ * 4;                    // row maxima (2 parities x 2 halves) exchanged between partner warps
constexpr int SMEM_BYTES = Q_BYTES + NST * KV_BYTES + X_BYTES + 256;
constexpr int TMEM_COLS = 256;
constexpr int COL_S = 0, COL_O = BKEY, COL_PH = COL_O + 2 * VR, COL_PL = COL_PH + BKEY;      // S | O_hi O_lo | P_hi | P_lo
static_assert(2 * (SMEM_BYTES + 1024) <= 233472 && COL_PL + BKEY <= TMEM_COLS, "two CTAs per SM");

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct Attn3Maps {
  CUtensorMap qkv;     // [2][images][tokens][3*dim] float32, box {32, rows, 1, 2}: one for Q (128 rows) ...
  CUtensorMap kk;      // ... one for K (64 rows)
  CUtensorMap vt;      // [2][images][heads*32][tok_pad] float32, box {32, 32, 1, 2}
};

// V^T planes: vt[p][img][head*32 + d][token] = v_p[img][token][head*32 + d] (0 for the padding tokens)
__global__ void __launch_bounds__(256) v_transpose3_kernel(const float* __restrict__ qkv, size_t qkv_plane, int tokens, int tok_pad, int dim,
                                                           float* __restrict__ vt, size_t vt_plane) {
  __shared__ float tile[2][64][HD + 1];
  const int t0 = blockIdx.x * 64, head = blockIdx.y, img = blockIdx.z;
  const int heads = dim / HD;
  for (int i = threadIdx.x; i < 2 * 64 * HD; i += 256) {
    const int p = i / (64 * HD), t = (i / HD) % 64, d = i % HD;
    const float* src = reinterpret_cast<const float*>(reinterpret_cast<const char*>(qkv) + p * qkv_plane);
    tile[p][t][d] = t0 + t < tokens ? src[((size_t)img * tokens + t0 + t) * 3 * dim + 2 * dim + head * HD + d] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * 64 * VR; i += 256) {
    const int p = i / (64 * VR), d = (i / 64) % VR, t = i % 64;
    float* base = reinterpret_cast<float*>(reinterpret_cast<char*>(vt) + p * vt_plane) + ((size_t)img * heads + head) * VR * tok_pad;
    if (t0 + t < tok_pad) base[(size_t)d * tok_pad + t0 + t] = tile[p][t][d];
  }
}

__global__ void __launch_bounds__(THREADS, 2) attention3_umma_kernel(const __grid_constant__ Attn3Maps maps, float* __restrict__ out,
                                                                     size_t out_plane, int tokens, int dim) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + Q_BYTES;
  uint8_t* sX = sKV + NST * KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sX + X_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * NST + 5);
  // K and V^T of a stage have their own full / empty barriers: K_j is free as soon as S_j has been computed (early), V_j only after
  // P_j V_j, and the next K must arrive a tile ahead of the next V -- with one barrier pair per stage the softmax warps spent 27 %
  // of their time waiting for S
  const uint32_t bar_k_full = smem_u32(bars), bar_k_empty = bar_k_full + 8 * NST, bar_v_full = bar_k_empty + 8 * NST,
                 bar_v_empty = bar_v_full + 8 * NST, bar_q = bar_v_empty + 8 * NST, bar_s_full = bar_q + 8, bar_s_empty = bar_s_full + 8,
                 bar_p_full = bar_s_empty + 8, bar_o_full = bar_p_full + 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * BQ, head = blockIdx.y, img = blockIdx.z;
  const int nk = (tokens + BKEY - 1) / BKEY;
  if ((smem_u32(smem) & 1023u) != 0) __trap();          // the swizzled tiles need 1024-byte alignment

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(bar_k_full + 8 * s, 1);
      mbar_init(bar_k_empty + 8 * s, 1);
      mbar_init(bar_v_full + 8 * s, 1);
      mbar_init(bar_v_empty + 8 * s, 1);
    }
    mbar_init(bar_q, 1);
    mbar_init(bar_s_full, 1);
    mbar_init(bar_s_empty, 8);
    mbar_init(bar_p_full, 8);
    mbar_init(bar_o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 9) {
    // TMA producer: Q once, then K_j / V^T_j in tile order into their 2-stage rings
    if (lane == 0) {
      mbar_expect_tx(bar_q, Q_BYTES);
      tma_load_4d(smem_u32(sQ), &maps.qkv, bar_q, head * HD, q0, img, 0);
      for (int j = 0; j < nk; ++j) {
        const uint32_t s = j % NST, ph = ((j / NST) & 1) ^ 1;
        const uint32_t dst = smem_u32(sKV + s * KV_BYTES);
        mbar_wait(bar_k_empty + 8 * s, ph);
        mbar_expect_tx(bar_k_full + 8 * s, K_BYTES);
        tma_load_4d(dst, &maps.kk, bar_k_full + 8 * s, dim + head * HD, j * BKEY, img, 0);
        mbar_wait(bar_v_empty + 8 * s, ph);
        mbar_expect_tx(bar_v_full + 8 * s, V_BYTES);
        // V^T: one box = one 32-key chunk of both planes, so the stage holds [chunk][plane][32 rows][128 B]
        tma_load_4d(dst + K_BYTES, &maps.vt, bar_v_full + 8 * s, j * BKEY, head * VR, img, 0);
        tma_load_4d(dst + K_BYTES + 2 * V_CHUNK, &maps.vt, bar_v_full + 8 * s, j * BKEY + 32, head * VR, img, 0);
      }
    }
  } else if (warp == 8) {
    // MMA warp: all lanes walk the loop with warp-uniform values, one elected lane issues (see conv_umma.cu)
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t tS = tmem_u + COL_S, tO = tmem_u + COL_O, tPh = tmem_u + COL_PH, tPl = tmem_u + COL_PL;
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = make_idesc_tf32(BQ, BKEY), idesc_o = make_idesc_tf32(BQ, VR), idesc_o2 = make_idesc_tf32(BQ, 2 * VR);
    auto issue_s = [&](int j) {
      const uint32_t s = j % NST;
      mbar_wait(bar_k_full + 8 * s, (j / NST) & 1);
      fence_after();
      const uint32_t a0 = smem_u32(sQ) >> 4, b0 = smem_u32(sKV + s * KV_BYTES) >> 4;
#pragma unroll
      for (int k8 = 0; k8 < HD / 8; ++k8) {
        const uint64_t ah = make_desc16<1024, 2>(a0 + k8 * 2), al = make_desc16<1024, 2>(a0 + Q_PLANE / 16 + k8 * 2);
        const uint64_t bh = make_desc16<1024, 2>(b0 + k8 * 2), bl = make_desc16<1024, 2>(b0 + K_PLANE / 16 + k8 * 2);
        if (leader) {
          mma_tf32(tS, al, bh, idesc_s, k8 ? 1u : 0u);
          mma_tf32(tS, ah, bl, idesc_s, 1u);
          mma_tf32(tS, ah, bh, idesc_s, 1u);
        }
      }
      if (leader) {
        commit(bar_s_full);
        commit(bar_k_empty + 8 * s);
      }
    };
    mbar_wait(bar_q, 0);
    issue_s(0);
    for (int j = 0; j < nk; ++j) {
      if (j + 1 < nk) {
        mbar_wait(bar_s_empty, j & 1);           // softmax has read S_j out of TMEM
        fence_after();
        issue_s(j + 1);
      }
      const uint32_t s = j % NST;
      mbar_wait(bar_v_full + 8 * s, (j / NST) & 1);
      mbar_wait(bar_p_full, j & 1);              // P_j is in tensor memory (and O_{j-1} has been consumed)
      fence_after();
      const uint32_t v0 = smem_u32(sKV + s * KV_BYTES + K_BYTES) >> 4;
#pragma unroll
      for (int k8 = 0; k8 < BKEY / 8; ++k8) {
        // V^T: [chunk][plane][32 rows][128 B] -- the hi and lo rows of a chunk are one contiguous 64-row B operand
        const uint64_t vb = make_desc16<1024, 2>(v0 + (k8 >> 2) * (2 * V_CHUNK / 16) + (k8 & 3) * 2);
        if (leader) {
          mma_tf32_ts(tO, tPh + k8 * 8, vb, idesc_o2, k8 ? 1u : 0u);      // this tile's P_hi V_hi | P_hi V_lo (the softmax threads accumulate over tiles)
          mma_tf32_ts(tO, tPl + k8 * 8, vb, idesc_o, 1u);                 // + P_lo V_hi
        }
      }
      if (leader) {
        commit(bar_v_empty + 8 * s);
        commit(bar_o_full);
      }
      __syncwarp();
    }
  } else {
    // (warps 8 and 9 fall through to the common exit)
    // softmax warps 0..7: warp w and w+4 share TMEM lanes 32 (w % 4) .., i.e. the same 32 query rows; `half` picks the 32 key
    // columns (= one swizzled P chunk) and the 16 output dimensions a thread owns
    const uint32_t tS = tmem + COL_S, tO = tmem + COL_O, tPh = tmem + COL_PH, tPl = tmem + COL_PL;
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;                  // query row of this thread = TMEM lane
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const float sl2 = 0.17677669529663687f * 1.4426950408889634f;      // 32^-0.5 * log2(e)
    float m_run = -INFINITY, l_acc = 0.f, o_acc[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) o_acc[d] = 0.f;
    constexpr int HK = BKEY / 2;
    float* xch = reinterpret_cast<float*>(sX);      // [2 parities][2 halves][128 rows] row maxima
    auto take_o = [&](float scale) {                // (accumulators + the finished tile's P V) * scale
      uint32_t v[16], v2[16];
      tmem_ld16(tO + lane_base + half * 16, v);
      tmem_ld16(tO + lane_base + VR + half * 16, v2);                                                              // the P_hi V_lo block
      tmem_wait_ld();
#pragma unroll
      for (int d = 0; d < 16; ++d) o_acc[d] = (o_acc[d] + (__uint_as_float(v[d]) + __uint_as_float(v2[d]))) * scale;
    };
    for (int j = 0; j < nk; ++j) {
      const int valid = min(BKEY, tokens - j * BKEY) - half * HK;        // keys of this thread's half that exist (may be <= 0)
      mbar_wait(bar_s_full, j & 1);
      fence_after();
      uint32_t sv[HK];
      tmem_ld32(tS + lane_base + half * HK, sv);
      tmem_wait_ld();
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_s_empty);
      if (valid < HK) {                                                  // last key tile only (warp-uniform)
#pragma unroll
        for (int i = 0; i < HK; ++i)
          if (i >= valid) sv[i] = 0xff800000u;                           // -inf
      }
      float mx4[4] = {__uint_as_float(sv[0]), __uint_as_float(sv[1]), __uint_as_float(sv[2]), __uint_as_float(sv[3])};
#pragma unroll
      for (int i = 4; i < HK; i += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) mx4[k] = fmaxf(mx4[k], __uint_as_float(sv[i + k]));
      }
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      xch[((j & 1) * 2 + half) * BQ + row] = mx;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      mx = fmaxf(mx, xch[((j & 1) * 2 + (half ^ 1)) * BQ + row]);
      const float m_new = fmaxf(m_run, mx * sl2);   // finite: every key tile has at least one existing key
      const float alpha = ex2(m_run - m_new);       // (0 for j == 0, unused there)
      m_run = m_new;
      // exponentials: they need neither O nor the P buffer, so they overlap the P_{j-1} V_{j-1} MMAs
      float p[HK];
#pragma unroll
      for (int i = 0; i < HK; ++i) p[i] = ex2(fmaf(__uint_as_float(sv[i]), sl2, -m_run));
      {                                              // row sum of this thread's 32 keys (the partner's half joins at the end)
        float s4[4] = {p[0], p[1], p[2], p[3]};
#pragma unroll
        for (int i = 4; i < HK; i += 4) {
#pragma unroll
          for (int k = 0; k < 4; ++k) s4[k] += p[i + k];
        }
        l_acc = fmaf(l_acc, alpha, (s4[0] + s4[1]) + (s4[2] + s4[3]));
      }
      if (j > 0) {                                   // P_{j-1} V_{j-1} has landed: into the register accumulators, on to this tile's scale
        mbar_wait(bar_o_full, (j - 1) & 1);
        fence_after();
        take_o(alpha);
      }
      // P_j -> tensor memory as hi / lo float32 column blocks (lane = query row, column = key): the A operand of the P V MMAs
#pragma unroll
      for (int c = 0; c < HK; c += 16) {
        uint32_t ph[16], pl[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float h, l;
          split_tf32(p[c + i], h, l);
          ph[i] = __float_as_uint(h), pl[i] = __float_as_uint(l);
        }
        tmem_st16(tPh + lane_base + half * HK + c, ph);
        tmem_st16(tPl + lane_base + half * HK + c, pl);
      }
      tmem_wait_st();
      fence_before();                                // the reads of O_{j-1} and the writes of P_j are ordered before the MMAs that follow the arrive
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p_full);
    }
    mbar_wait(bar_o_full, (nk - 1) & 1);
    fence_after();
    take_o(1.f);
    // row sums of the two halves: through the exchange slots of the parity the last tile did NOT use (their last readers are
    // behind that tile's bar.sync)
    xch[((nk & 1) * 2 + half) * BQ + row] = l_acc;
    asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
    l_acc += xch[((nk & 1) * 2 + (half ^ 1)) * BQ + row];
    if (q0 + row < tokens) {
      const float inv = 1.f / l_acc;
      float* op = out + ((size_t)img * tokens + q0 + row) * dim + head * HD + half * 16;
      float* ol = reinterpret_cast<float*>(reinterpret_cast<char*>(op) + out_plane);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_tf32(o_acc[4 * i + e] * inv, h[e], l[e]);
        reinterpret_cast<float4*>(op)[i] = make_float4(h[0], h[1], h[2], h[3]);
        reinterpret_cast<float4*>(ol)[i] = make_float4(l[0], l[1], l[2], l[3]);
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace

size_t ttk_attention3_scratch_bytes(int images, int tokens, int heads, int head_dim) {
  const size_t tok_pad = (size_t)(tokens + 7) / 8 * 8;
  (void)head_dim;
  return 2 * (size_t)images * heads * VR * tok_pad * 4;      // (callers may pass more)
}

int ttk_attention3(const float* qkv, size_t qkv_plane, float* out, size_t out_plane, void* vt_scratch, int images, int tokens, int heads,
                   int head_dim, cudaStream_t st, int v_transposed) {
  if (head_dim != HD || images <= 0 || tokens <= 0 || images > 65535 || qkv_plane % 16 || out_plane % 16) {
    ttk_set_error("ttk_attention3: unsupported shape (head_dim %d, tokens %d, images %d)", head_dim, tokens, images);
    return TTK_ERR_UNSUPPORTED;
  }
  const int dim = heads * HD, tok_pad = (tokens + 7) / 8 * 8;
  static TtkPerDevice attr;
  if (attr.first()) {
    TTK_CUDA(cudaFuncSetAttribute(attention3_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  }
  float* vt = (float*)vt_scratch;
  const size_t vt_plane = (size_t)images * heads * VR * tok_pad * 4;
  if (!v_transposed) {       // (the x3 qkv GEMM of vit.cu writes V^T itself)
    v_transpose3_kernel<<<dim3(ttk_cdiv(tok_pad, 64), heads, images), 256, 0, st>>>(qkv, qkv_plane, tokens, tok_pad, dim, vt, vt_plane);
    TTK_LAUNCH_CHECK();
  }
  Attn3Maps maps;
  bool ok = true;
  {
    cuuint64_t dims[4] = {(cuuint64_t)3 * dim, (cuuint64_t)tokens, (cuuint64_t)images, 2};
    cuuint64_t strides[3] = {(cuuint64_t)3 * dim * 4, (cuuint64_t)tokens * 3 * dim * 4, (cuuint64_t)qkv_plane};
    cuuint32_t box_q[4] = {HD, BQ, 1, 2}, box_k[4] = {HD, BKEY, 1, 2};
    ok = ok && encode_f32(&maps.qkv, qkv, 4, dims, strides, box_q, CU_TENSOR_MAP_SWIZZLE_128B);
    ok = ok && encode_f32(&maps.kk, qkv, 4, dims, strides, box_k, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)tokens, (cuuint64_t)heads * VR, (cuuint64_t)images, 2};
    cuuint64_t strides[3] = {(cuuint64_t)tok_pad * 4, (cuuint64_t)heads * VR * tok_pad * 4, (cuuint64_t)vt_plane};
    cuuint32_t box[4] = {32, VR, 1, 2};
    ok = ok && encode_f32(&maps.vt, vt, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  if (!ok) {
    ttk_set_error("ttk_attention3: cuTensorMapEncodeTiled failed");
    return TTK_ERR_CUDA;
  }
  attention3_umma_kernel<<<dim3(ttk_cdiv(tokens, BQ), heads, images), THREADS, SMEM_BYTES, st>>>(maps, out, out_plane, tokens, dim);
  TTK_LAUNCH_CHECK();
  return TTK_OK;
}
