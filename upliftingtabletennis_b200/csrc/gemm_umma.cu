// bf16 GEMM with fused epilogue on the 5th-generation tensor cores, for the ViTPose detector's Linear / patch-embedding /
// transposed-convolution layers (vit.cu):   C[m][n] = act(sum_k A[m][k] W[n][k] + bias[n]) (+ R[m][n])
//   A [M][K] bf16 (activations, K contiguous), W [N][K] bf16 (torch Linear weight as stored) -- both K-major UMMA operands.
// Persistent CTAs, static tile round-robin (n fastest, so the CTAs working on one 128-row slab of A run together and A is read
// from HBM once).  Roles: warp 0 = TMA producer (128 x 64 A box + BN x 64 W box per stage, 128-byte swizzle), warp 1 = MMA issuer
// (one elected thread: 4 x tcgen05.mma 128 x BN x 16 per stage), warps 2-9 = epilogue (tcgen05.ld -> bias / GELU / ReLU /
// residual -> staged, row-contiguous bf16 or float32 stores; thread = output row, two warps per TMEM lane quarter).  smem ring of STAGES stages, two TMEM accumulators of BN columns
// so the epilogue of tile i overlaps the MMAs of tile i+1.
// BN is 192 where N allows (384, 1152, 1536): one MMA then computes for 96 clk against 80 clk of operand reads
// (10 KB at 128 B/clk), i.e. the tensor pipe, not shared memory, is the limiter; N = 256 uses BN = 256, anything else 128.
#include "umma_prims.h"
#include "vit.h"

namespace {

using namespace umma;

constexpr int BM = 128, THREADS = 320;      // TMA warp, MMA warp, 8 epilogue warps
constexpr int STG_ROW = 80;         // bytes per staged output row piece (64 + 16 padding: conflict-free 16-byte accesses)


// implicit transposed-convolution mode: an M tile is a 16 x 8 pixel patch of the input grid (pixel-major rows of the A operand)
constexpr int PW = 16, PH = 8;

struct GemmMaps {
  CUtensorMap a, w;
};

// WRES: the whole BN x K weight tile (K = 384 = 6 k-blocks) stays in shared memory and the CTA walks M tiles of one N tile, so only
// A streams through the ring.  L2 -> SM operand traffic, not the tensor pipe, bounds the streaming variant (~42 B/clk per SM
// against 40 KB per 384 clk of MMA); with the weights resident it drops from 40 KB to 16 KB per k-block.
// X3 (3xTF32, fp32-level results): operands are split float32 plane pairs (hi | lo, vit.h), a k-block is 32 floats (the same 128-byte
// rows), a stage holds A_hi | A_lo | W_hi | W_lo (one TMA box with a plane dimension per operand) and every k8 step issues three
// kind::tf32 MMAs: lo*hi, hi*lo, hi*hi.  64 KB per stage against 12 x 64 clk of MMA: the L2 -> SM operand stream bounds it.
constexpr int KB_RES = 6;
template <int BN, bool WRES, bool X3 = false>
struct GCfg {
  static_assert(!X3 || !WRES, "the split weights of an N tile do not fit in shared memory");
  static constexpr int BK = X3 ? 32 : 64;            // elements per k-block
  static constexpr int A_PLANE = BM * 128, W_PLANE = BN * 128;
  static constexpr int A_BYTES = (X3 ? 2 : 1) * A_PLANE, W_BYTES = (X3 ? 2 : 1) * W_PLANE;
  static constexpr int STAGE_BYTES = WRES ? A_BYTES : A_BYTES + W_BYTES;
  static constexpr int STAGES = X3 ? 3 : WRES ? (BN == 192 ? 3 : 6) : (BN == 128 ? 6 : BN == 192 ? 5 : 4);
  static constexpr int WRES_BYTES = WRES ? KB_RES * W_BYTES : 0;
  static constexpr int RING_BYTES = WRES_BYTES + STAGES * STAGE_BYTES;
  static constexpr int TMEM_COLS = BN == 128 ? 256 : 512;
  static constexpr int SMEM_BYTES = 1024 + RING_BYTES + 256 + 1024 + 8 * 32 * STG_ROW;      // + barriers, bias tile, output staging
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

// tile walk of one CTA: streaming variant = global round-robin over (m, n) tiles, n fastest; resident variant = M tiles of a fixed n
template <bool WRES>
struct Walk {
  int tn_fixed, first, step, limit, tiles_n;
  __device__ Walk(int tiles_m, int tiles_n_) : tiles_n(tiles_n_) {
    if (WRES) {
      tn_fixed = blockIdx.x % tiles_n;
      first = blockIdx.x / tiles_n;
      step = (gridDim.x - tn_fixed + tiles_n - 1) / tiles_n;
      limit = tiles_m;
    } else {
      tn_fixed = 0;
      first = blockIdx.x;
      step = gridDim.x;
      limit = tiles_m * tiles_n;
    }
  }
  __device__ int tn(int t) const { return WRES ? tn_fixed : t % tiles_n; }
  __device__ int tm(int t) const { return WRES ? t : t / tiles_n; }
};

// exact (erf) GELU with one MUFU: erf(z) = 1 - exp2(-z Q(z)) for 0 <= z <= 4 (Q: degree-5 least-squares fit, |erf error| <= 3.3e-7,
// |GELU error| <= 4e-7, far below the bf16 resolution of the output); erf(z > 4) = 1 to 1.5e-8.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fminf(fabsf(x) * 0.70710678118654752440f, 4.0f);
  float q = -1.588800078e-04f;
  q = fmaf(q, z, 3.746585688e-03f);
  q = fmaf(q, z, -3.103881516e-02f);
  q = fmaf(q, z, 1.498060673e-01f);
  q = fmaf(q, z, 9.181324244e-01f);
  q = fmaf(q, z, 1.627928257e+00f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * q));
  const float erf_abs = 1.f - e;
  return 0.5f * x * (1.f + copysignf(erf_abs, x));
}

template <int BN, bool WRES, int ACT, bool X3>
__global__ void __launch_bounds__(THREADS, 1) gemm_umma_kernel(const __grid_constant__ GemmMaps maps, const GemmArgs g, int tiles_m, int tiles_n) {
  using C = GCfg<BN, WRES, X3>;
  constexpr int BK = C::BK;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = smem + C::WRES_BYTES;             // [resident W k-blocks][A (+ W) stages]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::RING_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 5);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * C::STAGES, bar_tfull = bar_empty + 8 * C::STAGES, bar_tempty = bar_tfull + 16,
                 bar_w = bar_tempty + 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kblocks = g.K / BK;
  const Walk<WRES> walk(tiles_m, tiles_n);

  if (tid == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, 8);
    }
    mbar_init(bar_w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), C::TMEM_COLS);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      if (WRES && walk.first < walk.limit) {          // the weight tile of this CTA's N tile, once
        mbar_expect_tx(bar_w, C::WRES_BYTES);
        for (int kb = 0; kb < KB_RES; ++kb) tma_load_2d(smem_u32(smem + kb * C::W_BYTES), &maps.w, bar_w, kb * BK, walk.tn_fixed * BN);
      }
      for (int tile = walk.first; tile < walk.limit; tile += walk.step) {
        const int tn = walk.tn(tile), tm = walk.tm(tile);
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const uint32_t s = it % C::STAGES, ph = (it / C::STAGES) & 1;
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          mbar_expect_tx(bar_full + 8 * s, C::STAGE_BYTES);
          const uint32_t dst = smem_u32(ring + s * C::STAGE_BYTES);
          if (g.implicit_c) {
            // tap t = (ty, tx): ty = 0 reads the pixel itself, ty = 1 the row below (py = 1) or above (py = 0); same along x.
            // Out-of-image pixels are TMA zero fill.
            const int kpt = g.implicit_c / BK, tap = kb / kpt, kc = kb % kpt;
            const int ptx = (g.up_w + PW - 1) / PW, pty = (g.up_h + PH - 1) / PH;
            const int px0 = (tm % ptx) * PW, py0 = ((tm / ptx) % pty) * PH, img = tm / (ptx * pty);
            const int dy = (tap >> 1) == 0 ? 0 : (g.py ? 1 : -1), dx = (tap & 1) == 0 ? 0 : (g.px ? 1 : -1);
            if (X3) tma_load_5d(dst, &maps.a, bar_full + 8 * s, kc * BK, px0 + dx, py0 + dy, img, 0);
            else tma_load_4d(dst, &maps.a, bar_full + 8 * s, kc * BK, px0 + dx, py0 + dy, img);
          } else if (X3) {
            tma_load_3d(dst, &maps.a, bar_full + 8 * s, kb * BK, tm * BM, 0);
          } else {
            tma_load_2d(dst, &maps.a, bar_full + 8 * s, kb * BK, tm * BM);
          }
          if (X3) tma_load_3d(dst + C::A_BYTES, &maps.w, bar_full + 8 * s, kb * BK, tn * BN, 0);
          else if (!WRES) tma_load_2d(dst + C::A_BYTES, &maps.w, bar_full + 8 * s, kb * BK, tn * BN);
        }
      }
    }
  } else if (warp == 1) {
    // the whole warp walks the loops with warp-uniform values, one elected lane issues (conv_umma.cu explains why)
    {
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      const bool leader = elect_one();
      constexpr uint32_t idesc = X3 ? make_idesc_tf32(BM, BN) : make_idesc(BM, BN);
      uint32_t it = 0, tcount = 0;
      if (WRES && walk.first < walk.limit) mbar_wait(bar_w, 0);
      for (int tile = walk.first; tile < walk.limit; tile += walk.step, ++tcount) {
        const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
        mbar_wait(bar_tempty + 8 * acc, aph ^ 1);        // accumulator drained by the epilogue (passes at once the first two times)
        fence_after();
        const uint32_t d = tmem_u + acc * BN;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const uint32_t s = it % C::STAGES, ph = (it / C::STAGES) & 1;
          mbar_wait(bar_full + 8 * s, ph);
          fence_after();
          const uint32_t a0 = smem_u32(ring + s * C::STAGE_BYTES) >> 4, b0 = WRES ? smem_u32(smem + kb * C::W_BYTES) >> 4 : a0 + C::A_BYTES / 16;
          if (X3) {
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8) {
              const uint64_t ah = make_desc16<1024, 2>(a0 + k8 * 2), al = make_desc16<1024, 2>(a0 + C::A_PLANE / 16 + k8 * 2);
              const uint64_t bh = make_desc16<1024, 2>(b0 + k8 * 2), bl = make_desc16<1024, 2>(b0 + C::W_PLANE / 16 + k8 * 2);
              if (leader) {
                mma_tf32(d, al, bh, idesc, (kb | k8) ? 1u : 0u);
                mma_tf32(d, ah, bl, idesc, 1u);
                mma_tf32(d, ah, bh, idesc, 1u);
              }
            }
          } else {
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16)
              if (leader) mma(d, make_desc16<1024, 2>(a0 + k16 * 2), make_desc16<1024, 2>(b0 + k16 * 2), idesc, (kb | k16) ? 1u : 0u);
          }
          if (leader) commit(bar_empty + 8 * s);
        }
        if (leader) commit(bar_tfull + 8 * acc);
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue warps 2..9: TMEM lane quarter q = warp % 4 (thread = accumulator row), column chunks interleaved between the two
    // warps of a quarter =====
    const int q = warp & 3, half = (warp - 2) >> 2;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    float* sBias = reinterpret_cast<float*>(smem + C::RING_BYTES + 256);
    const uint32_t stg_u32 = smem_u32(smem + C::RING_BYTES + 256 + 1024 + (warp - 2) * (32 * STG_ROW));
    const int et = tid - 64;                        // 0..255 among the epilogue threads
    // kernel parameters in registers (the asm memory clobbers below would otherwise re-read them from the constant bank)
    const uint32_t sbias_u32 = smem_u32(sBias);
    const bool implicit = g.implicit_c > 0;
    const int gM = g.M, gN = g.N, dbg = g.act >> 8, r_mod = g.r_mod, up_w = g.up_w, up_h = g.up_h, upy = g.py, upx = g.px;
    const bool c_bf16 = g.c_bf16 != 0;
    const bool c_split = X3 && g.c_split != 0;
    const long long c_plane = (long long)g.c_plane;
    float* const vt = X3 ? g.vt : nullptr;
    const int vt_col0 = g.vt_col0, vt_tokens = g.vt_tokens, vt_tok_pad = g.vt_tok_pad;
    const size_t vt_plane = g.vt_plane;
    const float* __restrict__ gR = g.R;
    const float* __restrict__ gBias = g.bias;
    uint8_t* const gC = (uint8_t*)g.C;
    const int esz = c_bf16 ? 2 : 4;
    const int rsel = lane >> 2, usel = lane & 3;    // copy-out: 4 lanes per 64-byte row piece, 8 rows per store instruction
    uint32_t tcount = 0;
    for (int tile = walk.first; tile < walk.limit; tile += walk.step, ++tcount) {
      const int tn = walk.tn(tile), tm = walk.tm(tile);
      const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
      const int m_warp = tm * BM + q * 32;
      const int m = m_warp + lane;
      const bool live = m < gM;                    // (residual rows; not used in the implicit mode)
      const int ptx = implicit ? (up_w + PW - 1) / PW : 1, pty = implicit ? (up_h + PH - 1) / PH : 1;
      const int ipx0 = (tm % ptx) * PW, ipy0 = ((tm / ptx) % pty) * PH, iimg = tm / (ptx * pty);
      // output byte offsets of the four rows this lane copies out (row r0 + lane / 4), -1 = row does not exist
      long long obase[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int mr = m_warp + 8 * i + rsel;
        size_t orow = (size_t)mr;
        bool ok = mr < gM;
        if (implicit) {
          const int rt = q * 32 + 8 * i + rsel;          // row within the 16 x 8 pixel patch
          const int x = ipx0 + rt % PW, y = ipy0 + rt / PW;
          ok = x < up_w && y < up_h;
          orow = ((size_t)iimg * 2 * up_h + 2 * y + upy) * 2 * up_w + 2 * x + upx;
        } else if (up_w) {
          const int x = mr % up_w, y = (mr / up_w) % up_h, img = mr / (up_w * up_h);
          orow = ((size_t)img * 2 * up_h + 2 * y + upy) * 2 * up_w + 2 * x + upx;
        }
        obase[i] = ok ? (long long)(orow * gN) * esz + usel * 16 : -1;
      }
      // bias of this tile's columns -> shared memory (read back as broadcasts)
      asm volatile("bar.sync 1, 256;" ::: "memory");             // previous tile's readers are done
      for (int i = et; i < BN; i += 256) sBias[i] = gBias ? __ldg(gBias + tn * BN + i) : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(bar_tfull + 8 * acc, aph);
      fence_after();
      if (dbg & 1) {                                  // measurement aid: mainloop only
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
        continue;
      }
      uint32_t v[32];
      tmem_ld32(lane_base + acc * BN + half * 32, v);
#pragma unroll 1
      for (int c0 = half * 32; c0 < BN; c0 += 64) {
        const int n = tn * BN + c0;
        float bz[32];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(bz[4 * j]), "=f"(bz[4 * j + 1]), "=f"(bz[4 * j + 2]), "=f"(bz[4 * j + 3]) : "r"(sbias_u32 + c0 * 4 + 16 * j));
        float r[32];
        if (gR && live) {
          const float4* rp = reinterpret_cast<const float4*>(gR + (size_t)(r_mod ? m % r_mod : m) * gN + n);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t = rp[j];
            r[4 * j] = t.x, r[4 * j + 1] = t.y, r[4 * j + 2] = t.z, r[4 * j + 3] = t.w;
          }
        }
        tmem_wait_ld();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (c0 + 64 < BN) tmem_ld32(lane_base + acc * BN + c0 + 64, v);      // next chunk's accumulators while this one is processed
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          f[j] += bz[j];
          if (ACT == VIT_ACT_GELU) f[j] = gelu_erf(f[j]);
          if (ACT == VIT_ACT_RELU) f[j] = fmaxf(f[j], 0.f);
        }
        if (gR && live) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] += r[j];
        }
        if (X3 && vt && n >= vt_col0) {
          // V columns of the qkv GEMM: written transposed ([dim][token], the attention's K-major B operand).  Lane = token, so the 32
          // lanes of a warp store 32 consecutive floats of one [dim] row: coalesced without staging.
          if (live) {
            const int img = m / vt_tokens, t = m - img * vt_tokens;
            float* dst = vt + ((size_t)img * (gN - vt_col0) + (n - vt_col0)) * vt_tok_pad + t;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float hi, lo;
              split_tf32(f[j], hi, lo);
              dst[(size_t)j * vt_tok_pad] = hi;
              *reinterpret_cast<float*>(reinterpret_cast<char*>(dst + (size_t)j * vt_tok_pad) + vt_plane) = lo;
            }
          }
          continue;
        }
        // stage 64-byte row pieces of the warp's 32 x 32 block in shared memory, then write them out row-contiguous (4 lanes per row,
        // full 32-byte sectors): bf16 = one piece of 32 columns, float32 = two pieces of 16 columns
        const uint32_t my = stg_u32 + lane * STG_ROW;
        const int pieces = c_bf16 ? 1 : c_split ? 4 : 2;      // split: two float32 pieces of the hi plane, then of the lo plane
#pragma unroll 1
        for (int pc = 0; pc < pieces; ++pc) {
          if (c_bf16) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                __nv_bfloat162 b2 = __floats2bfloat162_rn(f[8 * j + 2 * e], f[8 * j + 2 * e + 1]);
                o[e] = *reinterpret_cast<uint32_t*>(&b2);
              }
              asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(my + 16 * j), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              // f[16 * pc + 4 * j ..]: select without dynamic register indexing
              float a0 = (pc & 1) ? f[16 + 4 * j] : f[4 * j], a1 = (pc & 1) ? f[17 + 4 * j] : f[4 * j + 1],
                    a2 = (pc & 1) ? f[18 + 4 * j] : f[4 * j + 2], a3 = (pc & 1) ? f[19 + 4 * j] : f[4 * j + 3];
              if (c_split) {
                float h0, h1, h2, h3, l0, l1, l2, l3;
                split_tf32(a0, h0, l0), split_tf32(a1, h1, l1), split_tf32(a2, h2, l2), split_tf32(a3, h3, l3);
                const bool lo = pc >= 2;
                a0 = lo ? l0 : h0, a1 = lo ? l1 : h1, a2 = lo ? l2 : h2, a3 = lo ? l3 : h3;
              }
              asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(my + 16 * j), "f"(a0), "f"(a1), "f"(a2), "f"(a3) : "memory");
            }
          }
          __syncwarp();
          uint4 val[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(val[i].x), "=r"(val[i].y), "=r"(val[i].z), "=r"(val[i].w)
                         : "r"(stg_u32 + (8 * i + rsel) * STG_ROW + usel * 16));
          const long long coff = (long long)(n + (pc & 1) * 16) * esz + (pc >= 2 ? c_plane : 0);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (obase[i] >= 0 && !(dbg & 2)) *reinterpret_cast<uint4*>(gC + obase[i] + coff) = val[i];
          __syncwarp();
        }
      }
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, C::TMEM_COLS);
}

template <int BN, bool WRES, int ACT, bool X3>
int launch_act(const GemmArgs& g, cudaStream_t st) {
  using C = GCfg<BN, WRES, X3>;
  constexpr int BK = C::BK;
  static TtkPerDevice attr;
  if (attr.first()) {
    TTK_CUDA(cudaFuncSetAttribute(gemm_umma_kernel<BN, WRES, ACT, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  }
  GemmMaps maps;
  bool ok_a, ok_w;
  int tiles_m = ttk_cdiv(g.M, BM);
  if (X3) {
    // split float32 operands: the plane pair is the outermost map dimension (box 2), so one TMA brings hi and lo of a tile
    const cuuint64_t K4 = (cuuint64_t)g.K * 4;
    if (g.implicit_c) {
      const int c = g.implicit_c, n_img = g.M / (g.up_h * g.up_w);
      cuuint64_t dims[5] = {(cuuint64_t)c, (cuuint64_t)g.up_w, (cuuint64_t)g.up_h, (cuuint64_t)n_img, 2};
      cuuint64_t strides[4] = {(cuuint64_t)c * 4, (cuuint64_t)g.up_w * c * 4, (cuuint64_t)g.up_h * g.up_w * c * 4, (cuuint64_t)g.a_plane};
      cuuint32_t box[5] = {BK, PW, PH, 1, 2};
      ok_a = encode_f32(&maps.a, g.A, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
      tiles_m = n_img * ttk_cdiv(g.up_w, PW) * ttk_cdiv(g.up_h, PH);
    } else {
      cuuint64_t dims[3] = {(cuuint64_t)g.K, (cuuint64_t)g.M, 2};
      cuuint64_t strides[2] = {K4, (cuuint64_t)g.a_plane};
      cuuint32_t box[3] = {BK, BM, 2};
      ok_a = encode_f32(&maps.a, g.A, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    }
    cuuint64_t dims[3] = {(cuuint64_t)g.K, (cuuint64_t)g.N, 2};
    cuuint64_t strides[2] = {K4, (cuuint64_t)g.w_plane};
    cuuint32_t box[3] = {BK, BN, 2};
    ok_w = encode_f32(&maps.w, g.W, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  } else if (g.implicit_c) {
    EncodeFn enc = get_encode();
    const int c = g.implicit_c, n_img = g.M / (g.up_h * g.up_w);
    cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)g.up_w, (cuuint64_t)g.up_h, (cuuint64_t)n_img};
    cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)g.up_w * c * 2, (cuuint64_t)g.up_h * g.up_w * c * 2};
    cuuint32_t box[4] = {BK, PW, PH, 1}, es[4] = {1, 1, 1, 1};
    ok_a = enc && enc(&maps.a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(g.A), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    tiles_m = n_img * ttk_cdiv(g.up_w, PW) * ttk_cdiv(g.up_h, PH);
  } else {
    ok_a = encode_2d(&maps.a, g.A, (uint64_t)g.M, (uint64_t)g.K, (uint64_t)g.K, BK, BM, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  if (!X3) ok_w = encode_2d(&maps.w, g.W, (uint64_t)g.N, (uint64_t)g.K, (uint64_t)g.K, BK, BN, CU_TENSOR_MAP_SWIZZLE_128B);
  if (!ok_a || !ok_w) {
    ttk_set_error("ttk_gemm_umma: cuTensorMapEncodeTiled failed (M %d N %d K %d)", g.M, g.N, g.K);
    return TTK_ERR_CUDA;
  }
  const int tiles_n = g.N / BN;
  const int grid = std::max(1, std::min(tiles_m * tiles_n, ttk_num_sms()));
  gemm_umma_kernel<BN, WRES, ACT, X3><<<grid, THREADS, C::SMEM_BYTES, st>>>(maps, g, tiles_m, tiles_n);
  TTK_LAUNCH_CHECK();
  return TTK_OK;
}

template <int BN, bool WRES, bool X3 = false>
int launch(const GemmArgs& g, cudaStream_t st) {      // the activation is a compile-time parameter of the epilogue
  switch (g.act & 0xff) {
    case VIT_ACT_GELU: return launch_act<BN, WRES, VIT_ACT_GELU, X3>(g, st);
    case VIT_ACT_RELU: return launch_act<BN, WRES, VIT_ACT_RELU, X3>(g, st);
    default: return launch_act<BN, WRES, VIT_ACT_NONE, X3>(g, st);
  }
}

}  // namespace

int ttk_gemm_umma(const GemmArgs& g, cudaStream_t st) {
  const int BK = g.x3 ? 32 : 64;
  if (g.M <= 0 || g.K % BK != 0 || g.N % 128 != 0 || (g.R && g.c_bf16) || (g.x3 && g.c_bf16) || (g.c_split && !g.x3) ||
      (g.x3 && (g.a_plane % 16 || g.w_plane % 16 || g.c_plane % 16)) || (g.vt && (!g.x3 || g.vt_col0 % 128 || g.vt_tokens <= 0 || g.up_w)) ||
      (g.implicit_c && (g.implicit_c % BK != 0 || g.K != 4 * g.implicit_c || g.up_w <= 0 || g.R))) {
    ttk_set_error("ttk_gemm_umma: unsupported shape M %d N %d K %d", g.M, g.N, g.K);
    return TTK_ERR_UNSUPPORTED;
  }
  if (g.x3) return launch<128, false, true>(g, st);
  // K = 384 (qkv, proj, fc1) with enough M tiles per CTA to amortise the weight load: weights resident in shared memory
  if (g.K == KB_RES * BK && g.M >= 16 * BM) return g.N >= 1152 && g.N % 192 == 0 ? launch<192, true>(g, st) : launch<128, true>(g, st);
  if (g.N % 192 == 0) return launch<192, false>(g, st);
  if (g.N % 256 == 0) return launch<256, false>(g, st);
  return launch<128, false>(g, st);
}
