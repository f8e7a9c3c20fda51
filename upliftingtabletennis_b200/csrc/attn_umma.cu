// Multi-head self-attention of the ViTPose backbone on the 5th-generation tensor cores (flash-attention schedule):
//   out = softmax((q * 32^-0.5) k^T) v per (image, head)      vit_pose/vit_models/backbone/vit.py:160-176 (Attention.forward)
// qkv [images * tokens][3 * 384] bf16 as the qkv GEMM wrote it (q | k | v, head h at columns h*32 of each third).
//
// One CTA = one (image, head, 128-query tile); it walks the key tiles (128 keys) once:
//   S   = Q K_j^T           tcgen05.mma 128 x 128 x 32, Q / K_j tiles staged by TMA (64-byte swizzle), S in TMEM (128 columns)
//   P_j = exp2(S*c - m)     thread r owns query row r: the 128 scores of its TMEM row are pulled into registers at once (S is then
//                           free, so the next Q K^T overlaps the exponentials), running maximum / sum in registers, P_j written as
//                           a bf16 K-major A operand (128-byte swizzle) in shared memory
//   O  += P_j V_j           tcgen05.mma 128 x 48 x 128; V is staged TRANSPOSED ([dim][key], from a small transpose kernel) so that it
//                           is an ordinary K-major B operand, with a row of ones appended so that the row sums of P come out of the
//                           same MMA; O and the row sums accumulate in TMEM (48 columns) over all key tiles and are rescaled in
//                           place only when a row maximum grew by more than 2^8 (lazy rescaling).
// Warps 0-7 are the softmax / correction warps (TMEM lane quarter = warp % 4, two threads per query row: 64 keys and 16 output
// dimensions each), warp 8 issues TMA and MMA.  The kernel needs 160 TMEM columns (256 allocated) and ~93 KB of shared memory, so
// two CTAs share an SM and one CTA's exponentials overlap the other's MMAs.
// The exponentials (128 x 128 per tile on the 16/clk MUFU) bound the kernel, not the tensor pipe (head dimension 32).
#include "umma_prims.h"
#include "vit.h"

namespace {

using namespace umma;

constexpr int HD = 32, BQ = 128, BKEY = 128, NST = 3, THREADS = 288;      // 8 softmax warps + 1 TMA / MMA warp
constexpr int Q_BYTES = BQ * HD * 2;                 // 8 KB, 64-byte rows
constexpr int K_BYTES = BKEY * HD * 2;               // 8 KB
constexpr int VR = 48;                               // rows of the staged V^T tile: 32 dims, a row of ones (row sums of P come out of the
                                                     // same MMA, in float32 and from the bf16-rounded P), 15 rows of zeros (N % 16 == 0)
constexpr int V_BYTES = 2 * VR * 128;                // two 64-key chunks of [48 rows][128 B]
constexpr int KV_BYTES = K_BYTES + V_BYTES;
constexpr int P_BYTES = 2 * BQ * 128;                // two 64-key chunks of [128 rows][128 B]
constexpr int X_BYTES = 4 * BQ * 4;                  // row maxima (2 parities x 2 halves) exchanged between partner warps
constexpr int SMEM_BYTES = 1024 + Q_BYTES + NST * KV_BYTES + P_BYTES + X_BYTES + 256;
constexpr int TMEM_COLS = 256;

__device__ __forceinline__ float ex2(float x) {       // bare MUFU.EX2 (exp2f adds denormal range handling around it)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnMaps {
  CUtensorMap qkv;     // [images][tokens][3*dim], box {32, 128, 1}, SWIZZLE_64B
  CUtensorMap vt;      // [images][heads*48][tokens], box {64, 48, 1}, SWIZZLE_128B
};

// V^T with the extra rows: vt[img][head*48 + d][token] = qkv[img][token][2*dim + head*32 + d] for d < 32, 1 for d == 32, 0 above
__global__ void __launch_bounds__(256) v_transpose_kernel(const __nv_bfloat16* __restrict__ qkv, int tokens, int tok_pad, int dim,
                                                          __nv_bfloat16* __restrict__ vt) {
  __shared__ __nv_bfloat16 tile[64][HD + 2];
  const int t0 = blockIdx.x * 64, head = blockIdx.y, img = blockIdx.z;
  const int heads = dim / HD;
  for (int i = threadIdx.x; i < 64 * HD; i += 256) {
    const int t = i / HD, d = i % HD;
    tile[t][d] = t0 + t < tokens ? qkv[((size_t)img * tokens + t0 + t) * 3 * dim + 2 * dim + head * HD + d] : __float2bfloat16(0.f);
  }
  __syncthreads();
  __nv_bfloat16* base = vt + ((size_t)img * heads + head) * VR * tok_pad;
  for (int i = threadIdx.x; i < 64 * VR; i += 256) {
    const int d = i / 64, t = i % 64;
    if (t0 + t < tok_pad) base[(size_t)d * tok_pad + t0 + t] = d < HD ? tile[t][d] : __float2bfloat16(d == HD && t0 + t < tokens ? 1.f : 0.f);
  }
}

__global__ void __launch_bounds__(THREADS, 2) attention_umma_kernel(const __grid_constant__ AttnMaps maps, __nv_bfloat16* __restrict__ out,
                                                                    int tokens, int dim) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + Q_BYTES;
  uint8_t* sP = sKV + NST * KV_BYTES;
  uint8_t* sX = sP + P_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sX + X_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 5);
  const uint32_t bar_kv_full = smem_u32(bars), bar_kv_empty = bar_kv_full + 8 * NST, bar_q = bar_kv_empty + 8 * NST, bar_s_full = bar_q + 8,
                 bar_s_empty = bar_s_full + 8, bar_p_full = bar_s_empty + 8, bar_o_full = bar_p_full + 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * BQ, head = blockIdx.y, img = blockIdx.z;
  const int nk = (tokens + BKEY - 1) / BKEY;

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(bar_kv_full + 8 * s, 1);
      mbar_init(bar_kv_empty + 8 * s, 1);
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
  const uint32_t tS = tmem, tO = tmem + BKEY;

  if (warp == 8) {
    if (lane == 0) {
      auto load_kv = [&](int j) {
        const uint32_t s = j % NST;
        mbar_expect_tx(bar_kv_full + 8 * s, KV_BYTES);
        const uint32_t dst = smem_u32(sKV + s * KV_BYTES);
        tma_load_3d(dst, &maps.qkv, bar_kv_full + 8 * s, dim + head * HD, j * BKEY, img);
        tma_load_3d(dst + K_BYTES, &maps.vt, bar_kv_full + 8 * s, j * BKEY, head * VR, img);
        tma_load_3d(dst + K_BYTES + VR * 128, &maps.vt, bar_kv_full + 8 * s, j * BKEY + 64, head * VR, img);
      };
      auto issue_s = [&](int j) {
        const uint32_t s = j % NST;
        mbar_wait(bar_kv_full + 8 * s, (j / NST) & 1);
        fence_after();
        const uint32_t a0 = smem_u32(sQ), b0 = smem_u32(sKV + s * KV_BYTES);
#pragma unroll
        for (int k16 = 0; k16 < HD / 16; ++k16)
          mma(tS, make_desc(a0 + k16 * 32, 512, 4), make_desc(b0 + k16 * 32, 512, 4), make_idesc(BQ, BKEY), k16 ? 1u : 0u);
        commit(bar_s_full);
      };
      mbar_expect_tx(bar_q, Q_BYTES);
      tma_load_3d(smem_u32(sQ), &maps.qkv, bar_q, head * HD, q0, img);
      for (int j = 0; j < NST && j < nk; ++j) load_kv(j);
      mbar_wait(bar_q, 0);
      issue_s(0);
      for (int j = 0; j < nk; ++j) {
        if (j + 1 < nk) {
          mbar_wait(bar_s_empty, j & 1);           // softmax has read S_j out of TMEM
          fence_after();
          issue_s(j + 1);
        }
        mbar_wait(bar_p_full, j & 1);              // P_j is in shared memory (and O_{j-1} has been consumed)
        fence_after();
        const uint32_t s = j % NST;
        const uint32_t p0 = smem_u32(sP), v0 = smem_u32(sKV + s * KV_BYTES + K_BYTES);
#pragma unroll
        for (int k16 = 0; k16 < BKEY / 16; ++k16)
          mma(tO, make_desc(p0 + (k16 >> 2) * (BQ * 128) + (k16 & 3) * 32, 1024, 2), make_desc(v0 + (k16 >> 2) * (VR * 128) + (k16 & 3) * 32, 1024, 2),
              make_idesc(BQ, VR), (j | k16) ? 1u : 0u);       // O (and the row sums) accumulate in TMEM over all key tiles
        commit(bar_kv_empty + 8 * s);
        commit(bar_o_full);
        if (j + NST < nk) {                          // refill this stage once P_j V_j has read it
          mbar_wait(bar_kv_empty + 8 * s, (j / NST) & 1);
          load_kv(j + NST);
        }
      }
    }
  } else {
    // softmax warps 0..7: warp w and w+4 share TMEM lanes 32 (w % 4) .., i.e. the same 32 query rows; `half` picks the 64 key
    // columns (= one swizzled P chunk) and the 16 output dimensions a thread owns.  Two threads per row double the warps that
    // can hide each other's TMEM / MUFU latencies.
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;                  // query row of this thread = TMEM lane
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const float sl2 = 0.17677669529663687f * 1.4426950408889634f;      // 32^-0.5 * log2(e)
    float m_run = -INFINITY;                        // the maximum the exponentials are scaled with (may lag the true one, see below)
    constexpr int HK = BKEY / 2;
    const uint32_t p_row = smem_u32(sP) + half * (BQ * 128) + row * 128;
    uint32_t p_addr[HK / 8];                        // the thread's eight 16-byte units of its P row, swizzled
#pragma unroll
    for (int u = 0; u < HK / 8; ++u) p_addr[u] = p_row + (((uint32_t)u ^ (row & 7)) << 4);
    float* xch = reinterpret_cast<float*>(sX);      // [2 parities][2 halves][128 rows] row maxima
    for (int j = 0; j < nk; ++j) {
      const int valid = min(BKEY, tokens - j * BKEY) - half * HK;        // keys of this thread's half that exist (may be <= 0)
      mbar_wait(bar_s_full, j & 1);
      fence_after();
      // the thread's 64 scores into registers, then S is free for the next Q K^T (which overlaps the exponentials below)
      uint32_t sv[HK];
#pragma unroll
      for (int c = 0; c < HK; c += 32) tmem_ld32(tS + lane_base + half * HK + c, sv + c);
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
      for (int i = 4; i < HK; i += 4) {               // four independent chains
#pragma unroll
        for (int k = 0; k < 4; ++k) mx4[k] = fmaxf(mx4[k], __uint_as_float(sv[i + k]));
      }
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      // row maximum over both halves: exchange through shared memory with the partner warp (named barrier per lane quarter)
      xch[((j & 1) * 2 + half) * BQ + row] = mx;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      mx = fmaxf(mx, xch[((j & 1) * 2 + (half ^ 1)) * BQ + row]);
      const float m_true = fmaxf(m_run, mx * sl2);
      // Lazy rescaling: O and the row sums stay in TMEM and are multiplied by 2^(m_run - m_true) only when the maximum grew by more
      // than 8 (log2 units); otherwise P is scaled with the lagging maximum (entries up to 2^8, exact in float32 / bf16 terms since
      // numerator and denominator carry the same factor).  The decision is per row; the TMEM round trip is warp-collective, so a
      // warp takes it when any of its rows needs it (both partner warps see the same rows and decide alike).
      const bool grow = j > 0 && m_true - m_run > 8.0f;
      if (j == 0) m_run = m_true;
      if (__any_sync(0xffffffffu, grow)) {
        mbar_wait(bar_o_full, (j - 1) & 1);          // P_{j-1} V_{j-1} has landed
        fence_after();
        const float alpha = grow ? ex2(m_run - m_true) : 1.f;
        uint32_t v[16];
        tmem_ld16(tO + lane_base + half * 16, v);
        tmem_wait_ld();
#pragma unroll
        for (int d = 0; d < 16; ++d) v[d] = __float_as_uint(__uint_as_float(v[d]) * alpha);
        tmem_st16(tO + lane_base + half * 16, v);
        if (half == 0) {                             // the row-sum column block (32..47) belongs to the first partner
          tmem_ld16(tO + lane_base + HD, v);
          tmem_wait_ld();
#pragma unroll
          for (int d = 0; d < 16; ++d) v[d] = __float_as_uint(__uint_as_float(v[d]) * alpha);
          tmem_st16(tO + lane_base + HD, v);
        }
        tmem_wait_st();
        if (grow) m_run = m_true;
      }
      // exponentials (P_j packed in registers): they need neither O nor the P buffer, so they overlap the P_{j-1} V_{j-1} MMA
      uint32_t pk[HK / 2];
#pragma unroll
      for (int i = 0; i < HK / 2; ++i) {
        const float p0 = ex2(fmaf(__uint_as_float(sv[2 * i]), sl2, -m_run));
        const float p1 = ex2(fmaf(__uint_as_float(sv[2 * i + 1]), sl2, -m_run));
        __nv_bfloat162 b2 = __floats2bfloat162_rn(p0, p1);
        pk[i] = *reinterpret_cast<uint32_t*>(&b2);
      }
      // P_{j-1} has been consumed by its MMA: P_j -> shared memory (bf16, 128-byte swizzle: 16-byte unit u of row r at u ^ (r & 7))
      if (j > 0) {
        mbar_wait(bar_o_full, (j - 1) & 1);
        fence_after();
      }
#pragma unroll
      for (int u = 0; u < HK / 8; ++u) {
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(p_addr[u]), "r"(pk[4 * u]), "r"(pk[4 * u + 1]), "r"(pk[4 * u + 2]), "r"(pk[4 * u + 3]) : "memory");
      }
      fence_before();                                // TMEM writes of a rescale are ordered before the MMA that follows the arrive
      fence_async_smem();                            // P_j visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p_full);
    }
    mbar_wait(bar_o_full, (nk - 1) & 1);
    fence_after();
    uint32_t v[16], w[16];
    tmem_ld16(tO + lane_base + half * 16, v);
    tmem_ld16(tO + lane_base + HD, w);               // column 32: sum_k P[row][k] (both threads of a row read it)
    tmem_wait_ld();
    if (q0 + row < tokens) {
      const float inv = 1.f / __uint_as_float(w[0]);
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(v[2 * i]) * inv, __uint_as_float(v[2 * i + 1]) * inv);
        pk[i] = *reinterpret_cast<uint32_t*>(&b2);
      }
      uint4* op = reinterpret_cast<uint4*>(out + ((size_t)img * tokens + q0 + row) * dim + head * HD + half * 16);
      op[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      op[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace

size_t ttk_attention_umma_scratch_bytes(int images, int tokens, int heads, int head_dim) {
  const size_t tok_pad = (size_t)(tokens + 7) / 8 * 8;
  (void)head_dim;
  return (size_t)images * heads * VR * tok_pad * 2;
}

int ttk_attention_umma(const __nv_bfloat16* qkv, __nv_bfloat16* out, void* vt_scratch, int images, int tokens, int heads, int head_dim,
                       cudaStream_t st) {
  if (head_dim != HD || images <= 0 || tokens <= 0 || images > 65535) {
    ttk_set_error("ttk_attention_umma: unsupported shape (head_dim %d, tokens %d, images %d)", head_dim, tokens, images);
    return TTK_ERR_UNSUPPORTED;
  }
  const int dim = heads * HD, tok_pad = (tokens + 7) / 8 * 8;
  EncodeFn enc = get_encode();
  if (!enc) {
    ttk_set_error("cuTensorMapEncodeTiled is not available from the driver");
    return TTK_ERR_CUDA;
  }
  static TtkPerDevice attr;
  if (attr.first()) {
    TTK_CUDA(cudaFuncSetAttribute(attention_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  }
  __nv_bfloat16* vt = (__nv_bfloat16*)vt_scratch;
  v_transpose_kernel<<<dim3(ttk_cdiv(tok_pad, 64), heads, images), 256, 0, st>>>(qkv, tokens, tok_pad, dim, vt);
  TTK_LAUNCH_CHECK();
  AttnMaps maps;
  cuuint32_t es[3] = {1, 1, 1};
  {
    cuuint64_t dims[3] = {(cuuint64_t)3 * dim, (cuuint64_t)tokens, (cuuint64_t)images};
    cuuint64_t strides[2] = {(cuuint64_t)3 * dim * 2, (cuuint64_t)tokens * 3 * dim * 2};
    cuuint32_t box[3] = {HD, BQ, 1};
    if (enc(&maps.qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(qkv), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      ttk_set_error("ttk_attention_umma: tensor map (qkv) failed");
      return TTK_ERR_CUDA;
    }
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)tokens, (cuuint64_t)heads * VR, (cuuint64_t)images};
    cuuint64_t strides[2] = {(cuuint64_t)tok_pad * 2, (cuuint64_t)heads * VR * tok_pad * 2};
    cuuint32_t box[3] = {64, VR, 1};
    if (enc(&maps.vt, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, vt, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      ttk_set_error("ttk_attention_umma: tensor map (v^T) failed");
      return TTK_ERR_CUDA;
    }
  }
  attention_umma_kernel<<<dim3(ttk_cdiv(tokens, BQ), heads, images), THREADS, SMEM_BYTES, st>>>(maps, out, tokens, dim);
  TTK_LAUNCH_CHECK();
  return TTK_OK;
}
