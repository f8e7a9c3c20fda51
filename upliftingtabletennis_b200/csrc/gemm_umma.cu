// bf16 GEMM with fused epilogue on the 5th-generation tensor cores, for the ViTPose detector's Linear / patch-embedding /
// transposed-convolution layers (vit.cu):   C[m][n] = act(sum_k A[m][k] W[n][k] + bias[n]) (+ R[m][n])
//   A [M][K] bf16 (activations, K contiguous), W [N][K] bf16 (torch Linear weight as stored) -- both K-major UMMA operands.
// Persistent CTAs, static tile round-robin (n fastest, so the CTAs working on one 128-row slab of A run together and A is read
// from HBM once).  Roles: warp 0 = TMA producer (128 x 64 A box + BN x 64 W box per stage, 128-byte swizzle), warp 1 = MMA issuer
// (one elected thread: 4 x tcgen05.mma 128 x BN x 16 per stage), warps 2-5 = epilogue (tcgen05.ld -> bias / GELU / ReLU /
// residual -> bf16 or float32 stores; thread = output row).  smem ring of STAGES stages, two TMEM accumulators of BN columns
// so the epilogue of tile i overlaps the MMAs of tile i+1.
// BN is 192 where N allows (384, 1152, 1536): one MMA then computes for 96 clk against 80 clk of operand reads
// (10 KB at 128 B/clk), i.e. the tensor pipe, not shared memory, is the limiter; N = 256 uses BN = 256, anything else 128.
#include "umma_prims.h"
#include "vit.h"

namespace {

using namespace umma;

constexpr int BM = 128, BK = 64, THREADS = 192;
constexpr int STG_ROW = 144;        // bytes per staged output row (128 + 16 padding: conflict-free 16-byte accesses)

struct GemmMaps {
  CUtensorMap a, w;
};

template <int BN>
struct GCfg {
  static constexpr int A_BYTES = BM * BK * 2, W_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + W_BYTES;
  static constexpr int STAGES = BN == 128 ? 6 : BN == 192 ? 5 : 4;
  static constexpr int TMEM_COLS = BN == 128 ? 256 : 512;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + 256 + 1024 + 4 * 32 * STG_ROW;      // + barriers, bias tile, output staging
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

template <int BN>
__global__ void __launch_bounds__(THREADS, 1) gemm_umma_kernel(const __grid_constant__ GemmMaps maps, const GemmArgs g, int tiles_m, int tiles_n) {
  using C = GCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * C::STAGES, bar_tfull = bar_empty + 8 * C::STAGES, bar_tempty = bar_tfull + 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total = tiles_m * tiles_n, kblocks = g.K / BK;

  if (tid == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, 4);
    }
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
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int tn = tile % tiles_n, tm = tile / tiles_n;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const uint32_t s = it % C::STAGES, ph = (it / C::STAGES) & 1;
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          mbar_expect_tx(bar_full + 8 * s, C::STAGE_BYTES);
          const uint32_t dst = smem_u32(smem + s * C::STAGE_BYTES);
          tma_load_2d(dst, &maps.a, bar_full + 8 * s, kb * BK, tm * BM);
          tma_load_2d(dst + C::A_BYTES, &maps.w, bar_full + 8 * s, kb * BK, tn * BN);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      uint32_t it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tcount) {
        const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
        mbar_wait(bar_tempty + 8 * acc, aph ^ 1);        // accumulator drained by the epilogue (passes at once the first two times)
        fence_after();
        const uint32_t d = tmem + acc * BN;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const uint32_t s = it % C::STAGES, ph = (it / C::STAGES) & 1;
          mbar_wait(bar_full + 8 * s, ph);
          fence_after();
          const uint32_t a0 = smem_u32(smem + s * C::STAGE_BYTES), b0 = a0 + C::A_BYTES;
#pragma unroll
          for (int k16 = 0; k16 < BK / 16; ++k16)
            mma(d, make_desc(a0 + k16 * 32, 1024, 2), make_desc(b0 + k16 * 32, 1024, 2), idesc, (kb | k16) ? 1u : 0u);
          commit(bar_empty + 8 * s);
        }
        commit(bar_tfull + 8 * acc);
      }
    }
  } else {
    // ===== epilogue warps 2..5: TMEM lane quarter q = warp % 4; thread = accumulator row =====
    const int q = warp & 3;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    float* sBias = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES + 256);
    uint8_t* stg = smem + C::STAGES * C::STAGE_BYTES + 256 + 1024 + (warp - 2) * (32 * STG_ROW);
    const uint32_t stg_u32 = smem_u32(stg);
    const int et = tid - 64;                        // 0..127 among the epilogue threads
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tcount) {
      const int tn = tile % tiles_n, tm = tile / tiles_n;
      const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
      const int m_warp = tm * BM + q * 32;
      const int m = m_warp + lane;
      const bool live = m < g.M;
      // bias of this tile's columns -> shared memory (read back as broadcasts)
      asm volatile("bar.sync 1, 128;" ::: "memory");             // previous tile's readers are done
      for (int i = et; i < BN; i += 128) sBias[i] = g.bias ? __ldg(g.bias + tn * BN + i) : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(bar_tfull + 8 * acc, aph);
      fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(lane_base + acc * BN + c0, v);
        const int n = tn * BN + c0;
        float r[32];
        if (g.R && live) {
          const float4* rp = reinterpret_cast<const float4*>(g.R + (size_t)m * g.N + n);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t = rp[j];
            r[4 * j] = t.x, r[4 * j + 1] = t.y, r[4 * j + 2] = t.z, r[4 * j + 3] = t.w;
          }
        }
        tmem_wait_ld();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          f[j] = __uint_as_float(v[j]) + sBias[c0 + j];
          if (g.act == VIT_ACT_GELU) f[j] = gelu_erf(f[j]);
          if (g.act == VIT_ACT_RELU) f[j] = fmaxf(f[j], 0.f);
        }
        if (g.R && live) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] += r[j];
        }
        // stage the warp's 32 x 32 block in shared memory, then write it out with row-contiguous (full sector) stores
        const uint32_t my = stg_u32 + lane * STG_ROW;
        if (g.c_bf16) {
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
          for (int j = 0; j < 8; ++j)
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(my + 16 * j), "f"(f[4 * j]), "f"(f[4 * j + 1]), "f"(f[4 * j + 2]), "f"(f[4 * j + 3]) : "memory");
        }
        __syncwarp();
        const int lpr = g.c_bf16 ? 4 : 8;            // lanes per row (16 bytes each)
        const int rpi = 32 / lpr;                    // rows per store instruction
#pragma unroll 1
        for (int r0 = 0; r0 < 32; r0 += rpi) {
          const int rr = r0 + lane / lpr, u = lane % lpr;
          const int mr = m_warp + rr;
          if (mr < g.M) {
            size_t orow = (size_t)mr;
            if (g.up_w) {
              const int x = mr % g.up_w, y = (mr / g.up_w) % g.up_h, img = mr / (g.up_w * g.up_h);
              orow = ((size_t)img * 2 * g.up_h + 2 * y + g.py) * 2 * g.up_w + 2 * x + g.px;
            }
            uint4 val;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w) : "r"(stg_u32 + rr * STG_ROW + u * 16));
            uint8_t* dst = (uint8_t*)g.C + (orow * g.N + n) * (g.c_bf16 ? 2 : 4) + u * 16;
            *reinterpret_cast<uint4*>(dst) = val;
          }
        }
        __syncwarp();
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

template <int BN>
int launch(const GemmArgs& g, cudaStream_t st) {
  using C = GCfg<BN>;
  static bool attr = false;
  if (!attr) {
    TTK_CUDA(cudaFuncSetAttribute(gemm_umma_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr = true;
  }
  GemmMaps maps;
  if (!encode_2d(&maps.a, g.A, (uint64_t)g.M, (uint64_t)g.K, (uint64_t)g.K, BK, BM, CU_TENSOR_MAP_SWIZZLE_128B) ||
      !encode_2d(&maps.w, g.W, (uint64_t)g.N, (uint64_t)g.K, (uint64_t)g.K, BK, BN, CU_TENSOR_MAP_SWIZZLE_128B)) {
    ttk_set_error("ttk_gemm_umma: cuTensorMapEncodeTiled failed (M %d N %d K %d)", g.M, g.N, g.K);
    return TTK_ERR_CUDA;
  }
  const int tiles_m = ttk_cdiv(g.M, BM), tiles_n = g.N / BN;
  const int grid = std::max(1, std::min(tiles_m * tiles_n, ttk_num_sms()));
  gemm_umma_kernel<BN><<<grid, THREADS, C::SMEM_BYTES, st>>>(maps, g, tiles_m, tiles_n);
  TTK_LAUNCH_CHECK();
  return TTK_OK;
}

}  // namespace

int ttk_gemm_umma(const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0 || g.K % BK != 0 || g.N % 128 != 0 || (g.R && g.c_bf16)) {
    ttk_set_error("ttk_gemm_umma: unsupported shape M %d N %d K %d", g.M, g.N, g.K);
    return TTK_ERR_UNSUPPORTED;
  }
  if (g.N % 192 == 0) return launch<192>(g, st);
  if (g.N % 256 == 0) return launch<256>(g, st);
  return launch<128>(g, st);
}
