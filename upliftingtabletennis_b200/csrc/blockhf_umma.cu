// Fused BasicBlock of the 16-channel full-resolution HRNet branch in TF32 (fp32 NHWC16 tensors), with HORIZONTAL tap fusion:
//   y = relu(conv2(relu(conv1(x) + b1)) + b2 + x)              balldetection/models/wasb.py:35-64 (BasicBlock.forward, BN folded)
// block_umma.cu fuses the two convolutions with one MMA per (input row, horizontal tap, k8 step): 84 thin MMAs (N <= 48) per 4-row tile,
// each at the tensor pipe's 45.5 clk floor -- the fused block was tensor bound at twice the HBM time.  Here the three horizontal taps
// share ONE MMA (as in conv_umma.cu's HF mode): lane = input pixel (no shift), B = [W(ky, kx) for kx, ky in range], N = 9 x 16 = 144
// (72 clk), results in three accumulator sets per row that the epilogue adds with a +-1 lane shift:
//   out[x] = set0[x - 1] + set1[x] + set2[x + 1]        (warp shuffles; the edge lanes of a warp exchange through shared memory)
// 28 MMAs per tile instead of 84.  Tile = 4 output rows x 124 output pixels from an 8 x 128 pixel input box: conv1's lanes 1..126 are
// valid intermediate pixels, conv2's lanes 2..125 valid outputs (1280 = 11 tiles either way).
//   M1: conv1, 16 MMAs -> 6 intermediate rows x 3 sets (288 TMEM columns)
//   E1: sets -> + b1, ReLU, zero outside the image, round to TF32 (nearest) -> shared memory in TMA's swizzled pixel-row layout
//   M2: conv2 from that tile, 12 MMAs -> 4 rows x 3 sets (192 columns)      E2: + b2 + x (fp32, from global / L2) -> ReLU -> global
// The residual is the fp32 tensor itself, not the TF32-rounded staged tile, so the block is in the same arithmetic class as the two
// separate convolutions.  Roles: warp 0 TMA, warp 1 MMA issuer, warps 2-5 / 6-9 two epilogue groups (half the rows of E1 / E2 each).
// One tile in flight per CTA; M1 of tile t+1 is issued behind M2 of tile t and runs under E2(t).
#include <algorithm>

#include "hrnet.h"
#include "umma_prims.h"

namespace {

using namespace umma;

constexpr int C = 16, ROWB = 64, BW = 128, WO = 124, R = 4, R1 = R + 2, RX = R + 4, THREADS = 320;
constexpr int SET = 3 * C;                            // accumulator columns of one row: [kx][channel]
constexpr int X_BYTES = RX * BW * ROWB, T_BYTES = R1 * BW * ROWB, W_BYTES = 9 * C * ROWB;      // 64 KB, 48 KB, 9 KB
constexpr int ACC1 = R1 * SET, ACC2 = R * SET;        // 288 + 192 TMEM columns
constexpr int E1_ROWS = R1 / 2, E2_ROWS = R / 2;      // rows per epilogue group
constexpr int XCH_FLOATS = 2 * 4 * 2 * E1_ROWS * C;   // [group][warp][set 0 of lane 31 | set 2 of lane 0][row][channel]
constexpr int SMEM_BYTES = 2 * 10240 + 2 * X_BYTES + T_BYTES + 2 * XCH_FLOATS * 4 + 256;
constexpr uint32_t LAYOUT = 4u;                       // SWIZZLE_64B
static_assert(SMEM_BYTES <= 232448 && ACC1 + ACC2 <= 512, "budget");

struct BlockArgs {
  const void *w1, *w2;               // packed like ttk_conv_umma_pack: [kx][ky][cout][cin] float32 (TF32-rounded)
  const float *b1, *b2;
  const float* x;                    // the block's input (residual)
  float* out;
  int n, h, w;
  int tiles_x, tiles_y, total;
};

// RR output rows from RR + 2 staged rows at `abase`: one MMA per input row and k8 step against [W(ky, kx) for all kx, ky in k0..k1]
template <int RR>
__device__ __forceinline__ void issue_conv_hf(uint32_t abase, uint32_t wbase, uint32_t d_acc, bool leader) {
  constexpr uint32_t RB16 = ROWB / 16;
  const uint32_t a16 = abase >> 4, w16 = wbase >> 4;
#pragma unroll 1
  for (int hr = 0; hr < RR + 2; ++hr) {
    const int yi = hr - 1;
    const int k0 = yi + 2 - RR > 0 ? yi + 2 - RR : 0;
    const int k1 = yi + 1 < 2 ? yi + 1 : 2;
    const uint32_t idesc = make_idesc_tf32(128, (k1 - k0 + 1) * SET);
    const uint32_t d = d_acc + (RR - 2 - yi + k0) * SET;
    const uint32_t arow = a16 + hr * BW * RB16;
    const uint32_t brow = w16 + k0 * SET * RB16;
#pragma unroll
    for (int ks = 0; ks < ROWB / 32; ++ks)
      if (leader) mma_tf32(d, make_desc16<8 * ROWB, LAYOUT>(arow + ks * 2), make_desc16<8 * ROWB, LAYOUT>(brow + ks * 2), idesc, 1u);
  }
}

__global__ void __launch_bounds__(THREADS, 1) blockhf_umma_kernel(const __grid_constant__ CUtensorMap xmap, const BlockArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW1 = smem;
  uint8_t* sW2 = sW1 + 10240;
  uint8_t* sX = sW2 + 10240;                          // two input tiles
  uint8_t* sT = sX + 2 * X_BYTES;                     // the intermediate tile
  float* sXch1 = reinterpret_cast<float*>(sT + T_BYTES);
  float* sXch2 = sXch1 + XCH_FLOATS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sXch2 + XCH_FLOATS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const uint32_t bar_xfull = smem_u32(bars), bar_xempty = bar_xfull + 16, bar_m1 = bar_xempty + 16, bar_st = bar_m1 + 8, bar_m2 = bar_st + 8,
                 bar_init = bar_m2 + 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int my_tiles = a.total > (int)blockIdx.x ? (a.total - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  // ---- one-time setup: weights re-ordered [kx][ky][cout] -> [ky][kx][cout] rows (software swizzle on absolute address bits, as TMA
  // does), zeroed intermediate tile ----
  {
    constexpr int CPR = ROWB / 16;
    for (int which = 0; which < 2; ++which) {
      const uint4* src = reinterpret_cast<const uint4*>(which ? a.w2 : a.w1);
      uint8_t* dstb = which ? sW2 : sW1;
      const uint32_t wb = smem_u32(dstb);
      for (int i = tid; i < W_BYTES / 16; i += THREADS) {
        const int row = i / CPR, co = row % C, kk = row / C, ky = kk / 3, kx = kk % 3;       // destination row (ky, kx, co)
        const int srow = (kx * 3 + ky) * C + co;
        uint32_t addr = wb + row * ROWB + (i % CPR) * 16;
        addr ^= ((addr >> 7) & 3u) << 4;
        *reinterpret_cast<uint4*>(dstb + (addr - wb)) = __ldg(src + srow * CPR + (i % CPR));
      }
    }
    for (int i = tid; i < T_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(sT)[i] = make_uint4(0, 0, 0, 0);
  }
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_xfull + 8 * s, 1);
      mbar_init(bar_xempty + 8 * s, 1);
    }
    mbar_init(bar_m1, 1);
    mbar_init(bar_st, 8);
    mbar_init(bar_m2, 1);
    mbar_init(bar_init, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_async_smem();                                  // weights / zeros: generic-proxy writes -> async proxy (tensor core)
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 512);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int t = 0; t < my_tiles; ++t) {
        const int tile = blockIdx.x + t * gridDim.x;
        const int tx = tile % a.tiles_x, ty = (tile / a.tiles_x) % a.tiles_y, img = tile / (a.tiles_x * a.tiles_y);
        const uint32_t s = t & 1, ph = (t >> 1) & 1;
        mbar_wait(bar_xempty + 8 * s, ph ^ 1);         // conv1 of the tile that used this buffer has read it
        mbar_expect_tx(bar_xfull + 8 * s, X_BYTES);
        tma_load_4d(smem_u32(sX + s * X_BYTES), &xmap, bar_xfull + 8 * s, 0, tx * WO - 2, ty * R - 2, img);
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t acc1 = tmem_u, acc2 = tmem_u + ACC1;
    mbar_wait(bar_init, 0);                            // accumulators zeroed
    fence_after();
    for (int t = 0; t < my_tiles; ++t) {
      const uint32_t s = t & 1;
      mbar_wait(bar_xfull + 8 * s, (t >> 1) & 1);
      fence_after();
      issue_conv_hf<R1>(smem_u32(sX + s * X_BYTES), smem_u32(sW1), acc1, leader);       // (under E2 of the previous tile)
      if (leader) {
        commit(bar_m1);
        commit(bar_xempty + 8 * s);                    // the residual comes from global memory: the input tile is free after conv1
      }
      __syncwarp();
      mbar_wait(bar_st, t & 1);                        // intermediate tile written (and acc1 zeroed again)
      fence_after();
      issue_conv_hf<R>(smem_u32(sT), smem_u32(sW2), acc2, leader);
      if (leader) commit(bar_m2);
      __syncwarp();
    }
  } else {
    const int q = warp & 3, ge = (warp - 2) >> 2;
    const int m = q * 32 + lane;                       // lane of the tile row: input pixel x0 - 2 + m
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t acc1 = tmem + lane_base, acc2 = acc1 + ACC1;
    for (int c = ge * 16; c < ACC1 + ACC2; c += 32) tmem_zero16(acc1 + c);
    tmem_wait_st();
    fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_init);
    float b1r[C], b2r[C];
#pragma unroll
    for (int j = 0; j < C; ++j) b1r[j] = __ldg(a.b1 + j), b2r[j] = __ldg(a.b2 + j);
    const uint32_t st_base = smem_u32(sT);
    const int e1_lo = ge * E1_ROWS, e2_lo = ge * E2_ROWS;
    // exchange slots of this warp / of its neighbours: [group][warp][side][row][channel]
    float* x1 = sXch1 + ((ge * 4 + q) * 2) * E1_ROWS * C;
    const float* x1p = sXch1 + ((ge * 4 + (q > 0 ? q - 1 : 0)) * 2) * E1_ROWS * C;              // set 0 of the previous warp's lane 31
    const float* x1n = sXch1 + ((ge * 4 + (q < 3 ? q + 1 : 3)) * 2 + 1) * E1_ROWS * C;          // set 2 of the next warp's lane 0
    float* x2 = sXch2 + ((ge * 4 + q) * 2) * E1_ROWS * C;
    const float* x2p = sXch2 + ((ge * 4 + (q > 0 ? q - 1 : 0)) * 2) * E1_ROWS * C;
    const float* x2n = sXch2 + ((ge * 4 + (q < 3 ? q + 1 : 3)) * 2 + 1) * E1_ROWS * C;
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = blockIdx.x + t * gridDim.x;
      const int tx = tile % a.tiles_x, ty = (tile / a.tiles_x) % a.tiles_y, img = tile / (a.tiles_x * a.tiles_y);
      const int x0 = tx * WO, y0 = ty * R;
      const int px = x0 - 2 + m;                       // image column of this lane
      // residual rows of E2, requested before anything else (fp32 x from global memory / L2)
      const bool out_lane = m >= 2 && m <= WO + 1 && px < a.w;
      uint32_t res[E2_ROWS][C];
#pragma unroll
      for (int g = 0; g < E2_ROWS; ++g) {
        const int oy = y0 + e2_lo + g;
        if (out_lane && oy < a.h) {
          const float* rp = a.x + (((size_t)img * a.h + oy) * a.w + px) * C;
          ldg256(rp, res[g]);
          ldg256(rp + 8, res[g] + 8);
        }
      }
      // ---- E1: intermediate pixel (y0 - 1 + ri, px), lanes 1..126 ----
      mbar_wait(bar_m1, t & 1);
      fence_after();
#pragma unroll 1
      for (int g = 0; g < E1_ROWS; ++g) {              // pass 1: edge lanes publish set 0 (lane 31) / set 2 (lane 0)
        const uint32_t ta = acc1 + (R1 - 1 - (e1_lo + g)) * SET;
        uint32_t e0[16], e2[16];
        tmem_ld16(ta, e0);
        tmem_ld16(ta + 2 * C, e2);
        tmem_wait_ld();
        if (lane == 31) {
#pragma unroll
          for (int u = 0; u < C / 4; ++u) reinterpret_cast<uint4*>(x1 + g * C)[u] = make_uint4(e0[4 * u], e0[4 * u + 1], e0[4 * u + 2], e0[4 * u + 3]);
        }
        if (lane == 0) {
#pragma unroll
          for (int u = 0; u < C / 4; ++u) reinterpret_cast<uint4*>(x1 + (E1_ROWS + g) * C)[u] = make_uint4(e2[4 * u], e2[4 * u + 1], e2[4 * u + 2], e2[4 * u + 3]);
        }
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + ge) : "memory");
#pragma unroll 1
      for (int g = 0; g < E1_ROWS; ++g) {
        const int ri = e1_lo + g;
        const uint32_t ta = acc1 + (R1 - 1 - ri) * SET;
        uint32_t v0[16], v1[16], v2[16];
        tmem_ld16(ta, v0);
        tmem_ld16(ta + C, v1);
        tmem_ld16(ta + 2 * C, v2);
        tmem_wait_ld();
        tmem_zero16(ta);
        tmem_zero16(ta + C);
        tmem_zero16(ta + 2 * C);
        const int iy = y0 - 1 + ri;
        const bool inside = m >= 1 && m <= BW - 2 && px >= 0 && px < a.w && iy >= 0 && iy < a.h;
        uint32_t pk[C];
        float el[C], er[C];                            // what lanes 0 / 31 take instead of the shuffled value (broadcast vector loads, no branches)
#pragma unroll
        for (int u = 0; u < C / 4; ++u) {
          const float4 p = reinterpret_cast<const float4*>(x1p + g * C)[u], n = reinterpret_cast<const float4*>(x1n + g * C)[u];
          el[4 * u] = p.x, el[4 * u + 1] = p.y, el[4 * u + 2] = p.z, el[4 * u + 3] = p.w;
          er[4 * u] = n.x, er[4 * u + 1] = n.y, er[4 * u + 2] = n.z, er[4 * u + 3] = n.w;
        }
#pragma unroll
        for (int j = 0; j < C; ++j) {
          float left = __shfl_up_sync(0xffffffffu, __uint_as_float(v0[j]), 1);
          float right = __shfl_down_sync(0xffffffffu, __uint_as_float(v2[j]), 1);
          left = lane == 0 ? el[j] : left;
          right = lane == 31 ? er[j] : right;
          const float f = inside ? fmaxf((left + right) + (__uint_as_float(v1[j]) + b1r[j]), 0.f) : 0.f;
          pk[j] = (__float_as_uint(f) + 0x1000u) & 0xFFFFE000u;          // round to TF32 (nearest), the operand format of conv2
        }
        const uint32_t row_ad = st_base + (ri * BW + m) * ROWB;
#pragma unroll
        for (int u = 0; u < ROWB / 16; ++u) {
          uint32_t ad = row_ad + u * 16;
          ad ^= ((ad >> 7) & 3u) << 4;
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(ad), "r"(pk[4 * u]), "r"(pk[4 * u + 1]), "r"(pk[4 * u + 2]), "r"(pk[4 * u + 3]) : "memory");
        }
      }
      tmem_wait_st();
      fence_before();
      fence_async_smem();                              // intermediate tile visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_st);
      // ---- E2: output pixel (y0 + r, px), lanes 2..125 ----
      mbar_wait(bar_m2, t & 1);
      fence_after();
#pragma unroll 1
      for (int g = 0; g < E2_ROWS; ++g) {
        const uint32_t ta = acc2 + (R - 1 - (e2_lo + g)) * SET;
        uint32_t e0[16], e2[16];
        tmem_ld16(ta, e0);
        tmem_ld16(ta + 2 * C, e2);
        tmem_wait_ld();
        if (lane == 31) {
#pragma unroll
          for (int u = 0; u < C / 4; ++u) reinterpret_cast<uint4*>(x2 + g * C)[u] = make_uint4(e0[4 * u], e0[4 * u + 1], e0[4 * u + 2], e0[4 * u + 3]);
        }
        if (lane == 0) {
#pragma unroll
          for (int u = 0; u < C / 4; ++u) reinterpret_cast<uint4*>(x2 + (E1_ROWS + g) * C)[u] = make_uint4(e2[4 * u], e2[4 * u + 1], e2[4 * u + 2], e2[4 * u + 3]);
        }
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + ge) : "memory");
#pragma unroll
      for (int g = 0; g < E2_ROWS; ++g) {
        const int r = e2_lo + g;
        const uint32_t ta = acc2 + (R - 1 - r) * SET;
        uint32_t v0[16], v1[16], v2[16];
        tmem_ld16(ta, v0);
        tmem_ld16(ta + C, v1);
        tmem_ld16(ta + 2 * C, v2);
        tmem_wait_ld();
        tmem_zero16(ta);
        tmem_zero16(ta + C);
        tmem_zero16(ta + 2 * C);
        const int oy = y0 + r;
        uint32_t o[C];
        float el[C], er[C];
#pragma unroll
        for (int u = 0; u < C / 4; ++u) {
          const float4 p = reinterpret_cast<const float4*>(x2p + g * C)[u], n = reinterpret_cast<const float4*>(x2n + g * C)[u];
          el[4 * u] = p.x, el[4 * u + 1] = p.y, el[4 * u + 2] = p.z, el[4 * u + 3] = p.w;
          er[4 * u] = n.x, er[4 * u + 1] = n.y, er[4 * u + 2] = n.z, er[4 * u + 3] = n.w;
        }
#pragma unroll
        for (int j = 0; j < C; ++j) {
          float left = __shfl_up_sync(0xffffffffu, __uint_as_float(v0[j]), 1);
          float right = __shfl_down_sync(0xffffffffu, __uint_as_float(v2[j]), 1);
          left = lane == 0 ? el[j] : left;
          right = lane == 31 ? er[j] : right;
          o[j] = __float_as_uint(fmaxf(((left + right) + (__uint_as_float(v1[j]) + b2r[j])) + __uint_as_float(res[g][j]), 0.f));
        }
        if (out_lane && oy < a.h) {
          float* op = a.out + (((size_t)img * a.h + oy) * a.w + px) * C;
          stg256(op, o);
          stg256(op + 8, o + 8);
        }
      }
      tmem_wait_st();
      fence_before();                                  // acc2 zeroed before conv2 of the next tile (ordered through bar_st)
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

}  // namespace

// y = relu(conv2(relu(conv1(x))) + x) for two 3x3 stride-1 convolutions with 16 (padded) channels on NHWC fp32 tensors, TF32 products
int ttk_blockhf_umma_launch(const TtkConv& c1, const TtkConv& c2, const void* x, void* y, int n, int h, int w, cudaStream_t st) {
  if (c1.k != 3 || c2.k != 3 || c1.stride != 1 || c2.stride != 1 || c1.cin_p != 16 || c1.cout_p != 16 || c2.cin_p != 16 || c2.cout_p != 16 || !c1.w_umma32 ||
      !c2.w_umma32)
    return TTK_ERR_UNSUPPORTED;
  EncodeFn encode = get_encode();
  if (!encode) {
    ttk_set_error("cuTensorMapEncodeTiled is not available from the driver");
    return TTK_ERR_CUDA;
  }
  static TtkPerDevice attr;
  if (attr.first()) TTK_CUDA(cudaFuncSetAttribute(blockhf_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  CUtensorMap map;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)w * C * 4, (cuuint64_t)h * w * C * 4};
  cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)BW, (cuuint32_t)RX, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  // TFLOAT32: TMA rounds the input to TF32 (nearest-even) while it fills shared memory, like the unfused convolutions' input maps
  if (encode(&map, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 4, const_cast<void*>(x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    ttk_set_error("cuTensorMapEncodeTiled failed for the fused block %s", c1.name.c_str());
    return TTK_ERR_CUDA;
  }
  BlockArgs a;
  a.w1 = c1.w_umma32, a.w2 = c2.w_umma32, a.b1 = c1.bias, a.b2 = c2.bias, a.x = (const float*)x, a.out = (float*)y;
  a.n = n, a.h = h, a.w = w;
  a.tiles_x = ttk_cdiv(w, WO), a.tiles_y = ttk_cdiv(h, R);
  a.total = a.tiles_x * a.tiles_y * n;
  const int grid = std::max(1, std::min(a.total, ttk_num_sms()));
  blockhf_umma_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(map, a);
  TTK_LAUNCH_CHECK();
  return TTK_OK;
}
