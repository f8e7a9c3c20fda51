// bf16 implicit-GEMM convolution (3x3 stride 1/2, 1x1) on the 5th-generation tensor cores:
// tcgen05.mma with TMEM accumulators, operands staged by TMA, fused bias / residual / ReLU epilogue.
// Reference ops: the Conv2d + BatchNorm2d(eval) + ReLU (+ residual / fuse sum) groups of
// balldetection/models/wasb.py:35-105, :179-245, :383-416.
//
// Mapping (NHWC bf16 activations):
//   M = 128 consecutive output pixels of one image row, N = Cout (or a 64-wide slice of it), K = taps x Cin.
//   A CTA tile is R output rows x 128 pixels.  ONE TMA box brings the (R+2) x 130 pixel halo of a
//   K-chunk (<= 64 channels = one swizzled row of 32/64/128 bytes per pixel) into shared memory;
//   the 9 taps are NOT re-loaded: tap (ky,kx) of output row r is the same staged tile with the
//   matrix descriptor's start address moved by ((r+ky)*130 + kx) pixel rows.  tools/umma_probe.cu
//   established on a B200 that the 32/64/128-byte swizzles are applied to absolute shared-memory
//   address bits, so row-shifted descriptors (base_offset 0) read exactly what TMA wrote.
//   Stride 2: the input is addressed as four parity sub-grids (even/odd rows x even/odd columns, one
//   tensor map each, strides doubled); every tap then reads one sub-grid with unit pixel stride.
//   Image borders come for free from TMA's zero fill of out-of-bounds coordinates.
//   Vertical tap fusion (3x3 stride 1, Cout <= 64): an MMA costs at least the shared-memory read of its
//   128 x 16 A slab, so N = Cout = 16..64 leaves the tensor pipe mostly fetching.  The accumulators of a
//   tile's R output rows sit side by side in TMEM in REVERSE row order; input row yi and horizontal tap kx
//   are then multiplied ONCE with B = [W(ky=0,kx) | W(ky=1,kx) | W(ky=2,kx)] (N = 3 Cout), which lands in
//   the column blocks of output rows yi+1, yi, yi-1.  3(R+2) wide MMAs replace 9R narrow ones.
//   Accumulators are zeroed by the epilogue after it drains them, so every MMA accumulates.
// Roles (persistent CTA, static tile round-robin): warp 0 = TMA producer, warp 1 = MMA issuer
// (one elected thread), warps 2-5 = epilogue (TMEM -> registers -> bias/residual/ReLU -> bf16 -> global).
// Pipelines: smem full/empty ring over load units, double-buffered TMEM accumulators (tmem full/empty).
#include <cuda.h>

#include <algorithm>
#include <vector>

#include "hrnet.h"

namespace {

constexpr int BW = 128;          // output pixels per tile row = MMA M
constexpr int THREADS = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x4000;\n\t"
      "@p bra LAB_DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "LAB_DONE:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 | LBO | SBO | version 1 | layout
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                                   // LBO (ignored for swizzled K-major layouts)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}
// kind::f16 instruction descriptor: D f32, A/B bf16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

// 256-bit read-only load: the 16 bf16 channels of one pixel (one 32-byte sector)
__device__ __forceinline__ void ldg256(const void* p, uint4& lo, uint4& hi) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
               : "l"(p));
}

__device__ __forceinline__ void tmem_zero16(uint32_t taddr) {
  const uint32_t z = 0;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z) : "memory");
}

constexpr int pow2_cols(int c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }
constexpr int al1024(int b) { return (b + 1023) & ~1023; }

// KS: 1 or 3; S: stride 1 or 2 (3x3 only); CIN: padded input channels; COUT: output channels handled by one CTA
// (blockIdx.y selects the slice when the layer has more); R: output rows per tile; STAGES: smem ring depth;
// RB: reserve shared memory for TMA-staged residual tiles (double buffered with the accumulators);
// KCO: channels per K-chunk when not the default (a narrower chunk buys a taller tile for the same shared memory).
template <int KS, int S, int CIN, int COUT, int R, int STAGES, int RB, int KCO = 0>
struct Cfg {
  static constexpr int KC = KCO ? KCO : (CIN < 64 ? CIN : 64);      // channels per K-chunk = one swizzled smem row
  static constexpr int NKC = CIN / KC;
  static constexpr int ROWB = KC * 2;
  static constexpr int PAD = KS / 2;
  static constexpr bool FUSE = KS == 3 && S == 1 && 3 * COUT <= 256;   // vertical tap fusion
  static constexpr int NPY = S == 2 ? 2 : 1;            // row-parity units per K-chunk
  static constexpr int NBOX = S == 2 ? 2 : 1;           // TMA boxes per unit (column parities)
  static constexpr int TW = S == 2 ? BW + 1 : BW + 2 * PAD;
  static constexpr int TR = S == 2 ? R + 1 : R + 2 * PAD;
  static constexpr int BOX_BYTES = TR * TW * ROWB;
  static constexpr int BOX_AL = al1024(BOX_BYTES);
  static constexpr int STAGE_BYTES = NBOX * BOX_AL;
  static constexpr int UNITS = NKC * NPY;               // load units per tile
  static constexpr int TAPS = KS * KS;
  static constexpr int W_BYTES = TAPS * NKC * COUT * ROWB;
  static constexpr int W_BYTES_AL = al1024(W_BYTES);
  static constexpr int ACC_COLS = R * COUT;
  static constexpr int TMEM_COLS = pow2_cols(2 * ACC_COLS);
  // residual staging: boxes of CB channels (<= 64 = 128 swizzled bytes per pixel) x 128 px x R rows
  static constexpr int CB = COUT < 64 ? COUT : 64;
  static constexpr int NRB = COUT / CB;
  static constexpr int RROWB = CB * 2;
  static constexpr int RBOX_BYTES = R * BW * RROWB;     // multiple of 1024
  static constexpr int RES_BYTES = RB ? 2 * NRB * RBOX_BYTES : 0;
  static constexpr uint32_t RSWZ = RROWB == 32 ? 1u : RROWB == 64 ? 3u : 7u;
  static constexpr int SMEM_BYTES = 1024 + W_BYTES_AL + STAGES * STAGE_BYTES + RES_BYTES + COUT * 4 + 256;
  static constexpr uint32_t LAYOUT = ROWB == 32 ? 6u : ROWB == 64 ? 4u : 2u;     // SWIZZLE_32B / 64B / 128B
  static constexpr uint32_t SWZ = ROWB == 32 ? 1u : ROWB == 64 ? 3u : 7u;
  static_assert(S == 1 || KS == 3, "stride 2 is implemented for 3x3 only");
  static_assert(2 * ACC_COLS <= 512, "accumulators exceed TMEM");
  static_assert(ACC_COLS < 64 || ACC_COLS % 64 == 0, "the epilogue drains whole groups of 64 accumulator columns");
  static_assert(COUT % 16 == 0 && COUT <= 256 && CIN % 16 == 0, "bad channel counts");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

struct KMaps {
  CUtensorMap m[4];            // stride 1: m[0]; stride 2: m[py * 2 + px] (parity sub-grids)
  CUtensorMap res;             // residual tensor (same shape as the output), when staged by TMA
};

struct KArgs {
  const __nv_bfloat16* w;      // packed weights, see ttk_conv_umma_pack
  const float* bias;
  __nv_bfloat16* out;
  const __nv_bfloat16* res[3];
  int rsh[3];
  int nres;
  int res_tma;                 // res[0] (shift 0) arrives through shared memory
  int dual;                    // 1x1 only: K-chunk kc is read from tensor map m[kc] (two concatenated inputs)
  int n, h, w_img, cout_total;
  int relu;
  int tiles_x, tiles_y, total;
};

template <int KS, int S, int CIN, int COUT, int R, int STAGES, int RB, int KCO = 0>
__global__ void __launch_bounds__(THREADS, 1) conv_umma_kernel(const __grid_constant__ KMaps maps, const KArgs a) {
  using C = Cfg<KS, S, CIN, COUT, R, STAGES, RB, KCO>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;
  uint8_t* sA = smem + C::W_BYTES_AL;
  uint8_t* sR = sA + STAGES * C::STAGE_BYTES;           // [acc][box][row][px][CB] swizzled (1024-aligned)
  float* sBias = reinterpret_cast<float*>(sR + C::RES_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + COUT);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 6);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * STAGES, bar_tfull = bar_empty + 8 * STAGES,
                 bar_tempty = bar_tfull + 16, bar_rfull = bar_tempty + 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_off = blockIdx.y * COUT;                  // output-channel slice of this CTA
  const bool res_tma = RB && a.res_tma;

  // ---- one-time setup: weights (software swizzle on absolute address bits, as TMA does), bias, barriers, TMEM ----
  {
    const uint4* wsrc = reinterpret_cast<const uint4*>(a.w) + (size_t)blockIdx.y * (C::W_BYTES / 16);
    const uint32_t wbase = smem_u32(sW);
    constexpr int CPR = C::ROWB / 16;
    for (int i = tid; i < C::W_BYTES / 16; i += THREADS) {
      uint32_t addr = wbase + (i / CPR) * C::ROWB + (i % CPR) * 16;
      addr ^= ((addr >> 7) & C::SWZ) << 4;
      *reinterpret_cast<uint4*>(sW + (addr - wbase)) = __ldg(wsrc + i);
    }
    for (int i = tid; i < COUT; i += THREADS) sBias[i] = a.bias[n_off + i];
  }
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, 4);
      mbar_init(bar_rfull + 8 * i, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // weights: generic-proxy writes -> async proxy (tensor core)
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(C::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < a.total; tile += gridDim.x, ++tcount) {
        const int tx = tile % a.tiles_x, ty = (tile / a.tiles_x) % a.tiles_y, img = tile / (a.tiles_x * a.tiles_y);
        for (int u = 0; u < C::UNITS; ++u, ++it) {
          const int kc = u / C::NPY, py = u % C::NPY;
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          mbar_expect_tx(bar_full + 8 * s, C::NBOX * C::BOX_BYTES);
          const uint32_t dst = smem_u32(sA + s * C::STAGE_BYTES);
          if (S == 1) {
            if (KS == 1 && a.dual)
              tma_load_4d(dst, &maps.m[kc], bar_full + 8 * s, 0, tx * BW, ty * R, img);     // channels beyond the tensor are zero filled
            else
              tma_load_4d(dst, &maps.m[0], bar_full + 8 * s, kc * C::KC, tx * BW - C::PAD, ty * R - C::PAD, img);
          } else {
#pragma unroll
            for (int px = 0; px < 2; ++px)
              tma_load_4d(dst + px * C::BOX_AL, &maps.m[py * 2 + px], bar_full + 8 * s, kc * C::KC, tx * BW - px, ty * R - py, img);
          }
        }
        if (res_tma) {
          // the residual buffer pairs with the accumulator buffer: free once the epilogue of tile-2 has finished
          const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
          mbar_wait(bar_tempty + 8 * acc, aph);
          mbar_expect_tx(bar_rfull + 8 * acc, C::NRB * C::RBOX_BYTES);
#pragma unroll
          for (int b = 0; b < C::NRB; ++b)
            tma_load_4d(smem_u32(sR + (acc * C::NRB + b) * C::RBOX_BYTES), &maps.res, bar_rfull + 8 * acc, n_off + b * C::CB, tx * BW,
                        ty * R, img);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t wbase = smem_u32(sW);
      uint32_t it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < a.total; tile += gridDim.x, ++tcount) {
        const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
        mbar_wait(bar_tempty + 8 * acc, aph);      // accumulators drained and zeroed by the epilogue
        tc_fence_after();
        const uint32_t d_acc = tmem + acc * C::ACC_COLS;
        for (int u = 0; u < C::UNITS; ++u, ++it) {
          const int kc = u / C::NPY, py = u % C::NPY;
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(bar_full + 8 * s, ph);
          tc_fence_after();
          const uint32_t abase = smem_u32(sA + s * C::STAGE_BYTES);
          if (C::FUSE) {
            // input (halo) row hr = yi + 1 feeds output rows yo = yi + 1 - ky; row yo lives in column block R-1-yo
#pragma unroll 1
            for (int hr = 0; hr < R + 2; ++hr) {
              const int yi = hr - 1;
              const int k0 = yi + 2 - R > 0 ? yi + 2 - R : 0;
              const int k1 = yi + 1 < 2 ? yi + 1 : 2;
              const uint32_t idesc = make_idesc(128, (k1 - k0 + 1) * COUT);
              const uint32_t d_tmem = d_acc + (R - 2 - yi + k0) * COUT;
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                const uint32_t arow = abase + (hr * C::TW + kx) * C::ROWB;
                const uint32_t brow = wbase + (((kc * 3 + kx) * 3 + k0) * COUT) * C::ROWB;
#pragma unroll
                for (int k16 = 0; k16 < C::KC / 16; ++k16)
                  umma(d_tmem, make_desc(arow + k16 * 32, 8 * C::ROWB, C::LAYOUT), make_desc(brow + k16 * 32, 8 * C::ROWB, C::LAYOUT), idesc, 1u);
              }
            }
          } else {
            constexpr uint32_t idesc = make_idesc(128, COUT);
#pragma unroll 1
            for (int r = 0; r < R; ++r) {
              const uint32_t d_tmem = d_acc + (R - 1 - r) * COUT;
#pragma unroll
              for (int tap = 0; tap < C::TAPS; ++tap) {
                const int ky = tap / KS, kx = tap % KS;
                uint32_t arow;
                if (S == 1) {
                  arow = abase + ((r + ky) * C::TW + kx) * C::ROWB;
                } else {
                  if ((ky != 1 ? 1 : 0) != py) continue;     // this unit holds the other row parity
                  const int px = kx != 1 ? 1 : 0;
                  arow = abase + px * C::BOX_AL + ((r + (ky == 2 ? 1 : 0)) * C::TW + (kx == 2 ? 1 : 0)) * C::ROWB;
                }
                const uint32_t brow = wbase + ((tap * C::NKC + kc) * COUT) * C::ROWB;
#pragma unroll
                for (int k16 = 0; k16 < C::KC / 16; ++k16)
                  umma(d_tmem, make_desc(arow + k16 * 32, 8 * C::ROWB, C::LAYOUT), make_desc(brow + k16 * 32, 8 * C::ROWB, C::LAYOUT), idesc, 1u);
              }
            }
          }
          umma_commit(bar_empty + 8 * s);          // smem stage reusable once these MMAs have read it
        }
        umma_commit(bar_tfull + 8 * acc);          // accumulators complete
      }
    }
  } else {
    // ===== epilogue warps 2..5: TMEM lane quarter (warp % 4) =====
    const int q = warp & 3;
    const int m = q * 32 + lane;                   // pixel within the tile row = TMEM lane
    constexpr int GCOLS = C::ACC_COLS < 64 ? C::ACC_COLS : 64;      // accumulator columns fetched per TMEM wait
    constexpr int NSUB = GCOLS / 16;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    // zero both accumulator buffers, then open them for the MMA warp
    for (int c = 0; c < 2 * C::ACC_COLS; c += 16) tmem_zero16(lane_base + c);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(bar_tempty);
      mbar_arrive(bar_tempty + 8);
    }
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < a.total; tile += gridDim.x, ++tcount) {
      const int tx = tile % a.tiles_x, ty = (tile / a.tiles_x) % a.tiles_y, img = tile / (a.tiles_x * a.tiles_y);
      const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
      const int ox = tx * BW + m;
      if (res_tma) mbar_wait(bar_rfull + 8 * acc, aph);
      mbar_wait(bar_tfull + 8 * acc, aph);
      tc_fence_after();
#pragma unroll 1
      for (int g0 = 0; g0 < C::ACC_COLS; g0 += GCOLS) {
        uint32_t v[NSUB][16];
        const uint32_t taddr = lane_base + acc * C::ACC_COLS + g0;
#pragma unroll
        for (int sb = 0; sb < NSUB; ++sb) tmem_ld16(taddr + sb * 16, v[sb]);
        // residual operands for the same columns, all fetched before the TMEM wait so that the latencies overlap (the
        // lower-resolution fuse terms used to be loaded one by one at their point of use: 0.25 ms per stage-3/4 fuse conv)
        uint4 rv[NSUB][2], rx[NSUB][2][2];
        const int nres = a.nres;
#pragma unroll
        for (int sb = 0; sb < NSUB; ++sb) {
          const int col = g0 + sb * 16, r = R - 1 - col / COUT, c0 = col % COUT;
          const int oy = ty * R + r;
          const bool live = ox < a.w_img && oy < a.h;
          rv[sb][0] = rv[sb][1] = make_uint4(0, 0, 0, 0);
          if (nres > 0) {
            if (res_tma) {
              const uint32_t base = smem_u32(sR + (acc * C::NRB + c0 / C::CB) * C::RBOX_BYTES);
              uint32_t ad = base + (r * BW + m) * C::RROWB + (c0 % C::CB) * 2;
              ad ^= ((ad >> 7) & C::RSWZ) << 4;
              const uint32_t ad2 = (base + (r * BW + m) * C::RROWB + (c0 % C::CB) * 2 + 16) ^ ((((base + (r * BW + m) * C::RROWB) >> 7) & C::RSWZ) << 4);
              asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(rv[sb][0].x), "=r"(rv[sb][0].y), "=r"(rv[sb][0].z), "=r"(rv[sb][0].w) : "r"(ad));
              asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(rv[sb][1].x), "=r"(rv[sb][1].y), "=r"(rv[sb][1].z), "=r"(rv[sb][1].w) : "r"(ad2));
            } else if (live) {
              const int sh = a.rsh[0];
              const uint4* rp = reinterpret_cast<const uint4*>(
                  a.res[0] + (((size_t)img * (a.h >> sh) + (oy >> sh)) * (a.w_img >> sh) + (ox >> sh)) * a.cout_total + n_off + c0);
              ldg256(rp, rv[sb][0], rv[sb][1]);
            }
#pragma unroll
            for (int rr = 1; rr < 3; ++rr) {
              if (rr < nres && live) {
                const int sh = a.rsh[rr];
                const uint4* rp = reinterpret_cast<const uint4*>(
                    a.res[rr] + (((size_t)img * (a.h >> sh) + (oy >> sh)) * (a.w_img >> sh) + (ox >> sh)) * a.cout_total + n_off + c0);
                ldg256(rp, rx[sb][rr - 1][0], rx[sb][rr - 1][1]);
              }
            }
          }
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int sb = 0; sb < NSUB; ++sb) tmem_zero16(taddr + sb * 16);     // leave the columns zeroed for the next tile
#pragma unroll
        for (int sb = 0; sb < NSUB; ++sb) {
          const int col = g0 + sb * 16, r = R - 1 - col / COUT, c0 = col % COUT;
          const int oy = ty * R + r;
          if (!(ox < a.w_img && oy < a.h)) continue;
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[sb][j]) + sBias[c0 + j];
          if (nres > 0) {
            const uint32_t rw[8] = {rv[sb][0].x, rv[sb][0].y, rv[sb][0].z, rv[sb][0].w, rv[sb][1].x, rv[sb][1].y, rv[sb][1].z, rv[sb][1].w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {            // bf16 -> f32 by bit placement (keeps the conversion pipe free)
              f[2 * j] += __uint_as_float(rw[j] << 16);
              f[2 * j + 1] += __uint_as_float(rw[j] & 0xffff0000u);
            }
#pragma unroll
            for (int rr = 1; rr < 3; ++rr) {
              if (rr >= nres) break;
              const uint4 r0 = rx[sb][rr - 1][0], r1 = rx[sb][rr - 1][1];
              const uint32_t xw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                f[2 * j] += __uint_as_float(xw[j] << 16);
                f[2 * j + 1] += __uint_as_float(xw[j] & 0xffff0000u);
              }
            }
          }
          if (a.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          uint32_t o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            __nv_bfloat162 b2 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
            o[j] = *reinterpret_cast<uint32_t*>(&b2);
          }
          // one 256-bit store: the 16 channels of a pixel are a whole 32-byte sector
          __nv_bfloat16* op = a.out + (((size_t)img * a.h + oy) * a.w_img + ox) * a.cout_total + n_off + c0;
          asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(op), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]),
                       "r"(o[5]), "r"(o[6]), "r"(o[7])
                       : "memory");
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(C::TMEM_COLS));
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    cudaDriverEntryPointQueryResult q;
    void* p = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeFn)p;
  }
  return fn;
}

// output channels one CTA handles: 128-wide 3x3 layers with 64+ input channels are split so the weights fit in shared memory
int cout_tile(const TtkConv& cv) { return (cv.k == 3 && cv.cout_p == 128 && cv.cin_p >= 64) ? 64 : cv.cout_p; }
bool fused_ky(const TtkConv& cv) { return cv.k == 3 && cv.stride == 1 && 3 * cout_tile(cv) <= 256; }
// channels per K-chunk (one swizzled shared-memory row).  The stride-1 3x3 layers with 64+ input channels are bound by their MMA
// count (tensor pipe 73-85 % busy; a 128-pixel x 16-channel MMA holds it ~79 clk at N <= 48, ~1 clk per column beyond) and the halo
// rows of a tile are pure overhead: 3 (R + 2) / R MMAs per K16 step and output row.  32-channel chunks halve the staging boxes so that
// taller tiles fit: transition1.0 R = 3 -> 8, the 64 -> 64 layers R = 2 -> 4, and the 128 -> 128 layers get a second pipeline stage.
int kc_of(const TtkConv& cv) {
  if (cv.k == 3 && cv.stride == 1 && cv.cin_p >= 64) return 32;
  // stride 2 with 64+ input channels (transition1.1 1.74 -> 1.49 ms): two-row tiles and a third pipeline stage in the same shared memory
  if (cv.k == 3 && cv.stride == 2 && cv.cin_p >= 64) return 32;
  return cv.cin_p < 64 ? cv.cin_p : 64;
}

CUtensorMapSwizzle swizzle_for(int row_bytes) {
  return row_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
}

template <int KS, int S, int CIN, int COUT, int R, int STAGES, int RB, int KCO = 0>
int launch(const TtkConv& cv, const ConvLaunch& a, cudaStream_t st) {
  using C = Cfg<KS, S, CIN, COUT, R, STAGES, RB, KCO>;
  if (C::KC != kc_of(cv)) {
    ttk_set_error("conv %s: kernel K-chunk %d differs from the packed weights' %d", cv.name.c_str(), C::KC, kc_of(cv));
    return TTK_ERR_STATE;
  }
  EncodeFn encode = get_encode();
  if (!encode) {
    ttk_set_error("cuTensorMapEncodeTiled is not available from the driver");
    return TTK_ERR_CUDA;
  }
  static bool attr = false;
  if (!attr) {
    TTK_CUDA(cudaFuncSetAttribute(conv_umma_kernel<KS, S, CIN, COUT, R, STAGES, RB, KCO>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr = true;
  }
  if (S == 2 && ((a.hin & 1) || (a.win & 1))) return TTK_ERR_UNSUPPORTED;
  KMaps maps;
  cuuint32_t box[4] = {(cuuint32_t)C::KC, (cuuint32_t)C::TW, (cuuint32_t)C::TR, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  for (int i = 0; i < (S == 2 ? 4 : 1); ++i) {
    const int py = i >> 1, px = i & 1;
    cuuint64_t dims[4] = {(cuuint64_t)CIN, (cuuint64_t)(a.win / S), (cuuint64_t)(a.hin / S), (cuuint64_t)a.n};
    cuuint64_t strides[3] = {(cuuint64_t)CIN * 2 * S, (cuuint64_t)a.win * CIN * 2 * S, (cuuint64_t)a.hin * a.win * CIN * 2};
    void* base = (char*)const_cast<void*>(a.in) + ((size_t)py * a.win + px) * CIN * 2;
    const CUresult r = encode(&maps.m[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              swizzle_for(C::ROWB), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      ttk_set_error("cuTensorMapEncodeTiled failed (%d) for conv %s", (int)r, cv.name.c_str());
      return TTK_ERR_CUDA;
    }
  }
  for (int i = (S == 2 ? 4 : 1); i < 4; ++i) maps.m[i] = maps.m[0];
  const bool res_tma = RB && a.nres > 0 && a.res_shift[0] == 0;
  maps.res = maps.m[0];
  if (res_tma) {
    cuuint64_t dims[4] = {(cuuint64_t)a.cout, (cuuint64_t)a.wout, (cuuint64_t)a.hout, (cuuint64_t)a.n};
    cuuint64_t strides[3] = {(cuuint64_t)a.cout * 2, (cuuint64_t)a.wout * a.cout * 2, (cuuint64_t)a.hout * a.wout * a.cout * 2};
    cuuint32_t rbox[4] = {(cuuint32_t)C::CB, (cuuint32_t)BW, (cuuint32_t)R, 1};
    const CUresult r = encode(&maps.res, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(a.res[0]), dims, strides, rbox, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(C::RROWB), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      ttk_set_error("cuTensorMapEncodeTiled (residual) failed (%d) for conv %s", (int)r, cv.name.c_str());
      return TTK_ERR_CUDA;
    }
  }
  KArgs k;
  k.w = cv.w_umma;
  k.bias = cv.bias;
  k.out = (__nv_bfloat16*)a.out;
  for (int i = 0; i < 3; ++i) {
    k.res[i] = (const __nv_bfloat16*)a.res[i];
    k.rsh[i] = a.res_shift[i];
  }
  k.nres = a.nres;
  k.res_tma = res_tma ? 1 : 0;
  k.dual = 0;
  k.n = a.n;
  k.h = a.hout;
  k.w_img = a.wout;
  k.cout_total = a.cout;
  k.relu = a.relu;
  k.tiles_x = ttk_cdiv(a.wout, BW);
  k.tiles_y = ttk_cdiv(a.hout, R);
  k.total = k.tiles_x * k.tiles_y * a.n;
  const int nsplit = a.cout / COUT;
  // persistent CTAs: as many as fit per SM (shared memory and TMEM columns), never more than there are tiles
  int occ = std::min(232448 / (C::SMEM_BYTES + 1024), 512 / C::TMEM_COLS);
  occ = std::max(1, std::min(occ, 2));
  const int gx = std::max(1, std::min(k.total, ttk_num_sms() * occ / nsplit));
  dim3 grid(gx, nsplit);
  conv_umma_kernel<KS, S, CIN, COUT, R, STAGES, RB, KCO><<<grid, THREADS, C::SMEM_BYTES, st>>>(maps, k);
  TTK_LAUNCH_CHECK();
  return TTK_OK;
}

int launch_dual(const __nv_bfloat16* w_dual, const float* bias_dual, const ConvLaunch& a, cudaStream_t st) {
  constexpr int R = 2, STAGES = 3;
  using C = Cfg<1, 1, 128, 128, R, STAGES, 0>;
  EncodeFn encode = get_encode();
  if (!encode) return TTK_ERR_UNSUPPORTED;
  static bool attr = false;
  if (!attr) {
    TTK_CUDA(cudaFuncSetAttribute(conv_umma_kernel<1, 1, 128, 128, R, STAGES, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr = true;
  }
  KMaps maps;
  cuuint32_t box[4] = {64, (cuuint32_t)BW, (cuuint32_t)R, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  const void* ins[2] = {a.in, a.in2};
  const int cins[2] = {a.cin, a.cin2};
  for (int i = 0; i < 2; ++i) {
    cuuint64_t dims[4] = {(cuuint64_t)cins[i], (cuuint64_t)a.win, (cuuint64_t)a.hin, (cuuint64_t)a.n};
    cuuint64_t strides[3] = {(cuuint64_t)cins[i] * 2, (cuuint64_t)a.win * cins[i] * 2, (cuuint64_t)a.hin * a.win * cins[i] * 2};
    const CUresult r = encode(&maps.m[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ins[i]), dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return TTK_ERR_UNSUPPORTED;      // e.g. a driver that rejects a box wider than the tensor
  }
  maps.m[2] = maps.m[3] = maps.res = maps.m[0];
  KArgs k;
  k.w = w_dual;
  k.bias = bias_dual;
  k.out = (__nv_bfloat16*)a.out;
  for (int i = 0; i < 3; ++i) {
    k.res[i] = nullptr;
    k.rsh[i] = 0;
  }
  k.nres = 0;
  k.res_tma = 0;
  k.dual = 1;
  k.n = a.n;
  k.h = a.hout;
  k.w_img = a.wout;
  k.cout_total = a.cout;
  k.relu = a.relu;
  k.tiles_x = ttk_cdiv(a.wout, BW);
  k.tiles_y = ttk_cdiv(a.hout, R);
  k.total = k.tiles_x * k.tiles_y * a.n;
  const int gx = std::max(1, std::min(k.total, ttk_num_sms()));
  conv_umma_kernel<1, 1, 128, 128, R, STAGES, 0><<<gx, THREADS, C::SMEM_BYTES, st>>>(maps, k);
  TTK_LAUNCH_CHECK();
  return TTK_OK;
}

}  // namespace

int ttk_conv_umma_launch_dual(const __nv_bfloat16* w_dual, const float* bias_dual, const ConvLaunch& a, cudaStream_t st) {
  if (a.cin != 32 || a.cin2 != 64 || a.cout != 128) return TTK_ERR_UNSUPPORTED;
  return launch_dual(w_dual, bias_dual, a, st);
}

// Weights for the tensor-core path: bf16 rows of KC input channels (K-major), zero padded, per output-channel slice:
//   fused 3x3 stride 1 : [slice][k-chunk][kx][ky][cout_tile][KC]   (B = the three vertical taps side by side)
//   otherwise          : [slice][tap][k-chunk][cout_tile][KC]
int ttk_conv_umma_pack(TtkConv& cv, const float* w_host) {
  const int kk = cv.k * cv.k;
  const int KC = kc_of(cv);
  const int nkc = cv.cin_p / KC;
  const int ct = cout_tile(cv);
  const bool fused = fused_ky(cv);
  std::vector<__nv_bfloat16> w((size_t)kk * cv.cin_p * cv.cout_p, __float2bfloat16_rn(0.f));
  for (int co = 0; co < cv.cout; ++co)
    for (int ci = 0; ci < cv.cin; ++ci)
      for (int t = 0; t < kk; ++t) {
        const int kc = ci / KC, c = ci % KC, sl = co / ct, cl = co % ct;
        size_t row;
        if (fused) {
          const int ky = t / 3, kx = t % 3;
          row = (((size_t)sl * nkc + kc) * 3 + kx) * 3 + ky;
        } else {
          row = ((size_t)sl * kk + t) * nkc + kc;
        }
        w[(row * ct + cl) * KC + c] = __float2bfloat16_rn(w_host[((size_t)co * cv.cin + ci) * kk + t]);
      }
  if (!cv.w_umma) TTK_CUDA(cudaMalloc((void**)&cv.w_umma, w.size() * sizeof(__nv_bfloat16)));
  TTK_CUDA(cudaMemcpy(cv.w_umma, w.data(), w.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
  return TTK_OK;
}

int ttk_conv_umma_launch(const TtkConv& cv, const ConvLaunch& a, cudaStream_t st) {
  const int ci = cv.cin_p, co = cout_tile(cv);
#define TTK_UMMA(KS_, S_, CI_, CO_, R_, ST_, RB_) \
  if (cv.k == KS_ && cv.stride == S_ && ci == CI_ && co == CO_) return launch<KS_, S_, CI_, CO_, R_, ST_, RB_>(cv, a, st);
  // 3x3 stride 1
  TTK_UMMA(3, 1, 16, 64, 4, 3, 0)     // stem conv1 (9 -> 64, input padded to 16 channels)
  if (cv.k == 3 && cv.stride == 1 && ci == 64 && co == 64) return launch<3, 1, 64, 64, 4, 3, 0, 32>(cv, a, st);      // stem conv2, quarter-resolution branch
  if (cv.k == 3 && cv.stride == 1 && ci == 32 && co == 32 && a.nres == 0) return launch<3, 1, 32, 32, 8, 2, 0>(cv, a, st);   // no residual tiles to stage: taller tile
  TTK_UMMA(3, 1, 32, 32, 4, 2, 1)     // bottleneck conv2, half-resolution branch
  TTK_UMMA(3, 1, 16, 16, 8, 3, 1)     // full-resolution branch
  if (cv.k == 3 && cv.stride == 1 && ci == 128 && co == 16) return launch<3, 1, 128, 16, 8, 2, 0, 32>(cv, a, st);     // transition1.0
  if (cv.k == 3 && cv.stride == 1 && ci == 128 && co == 64) return launch<3, 1, 128, 64, 2, 2, 0, 32>(cv, a, st);    // eighth-resolution branch (128 -> 128 as two 64-channel output slices)
  // 3x3 stride 2 (transitions and fuse down-paths)
  if (cv.k == 3 && cv.stride == 2 && ci == 128 && co == 32) return launch<3, 2, 128, 32, 2, 3, 0, 32>(cv, a, st);
  TTK_UMMA(3, 2, 16, 16, 4, 3, 0)
  TTK_UMMA(3, 2, 16, 32, 4, 3, 0)
  TTK_UMMA(3, 2, 16, 64, 4, 3, 0)
  TTK_UMMA(3, 2, 16, 128, 2, 3, 0)
  TTK_UMMA(3, 2, 32, 32, 4, 2, 0)
  TTK_UMMA(3, 2, 32, 64, 4, 2, 0)
  TTK_UMMA(3, 2, 32, 128, 2, 2, 0)
  if (cv.k == 3 && cv.stride == 2 && ci == 64 && co == 64) return launch<3, 2, 64, 64, 2, 3, 0, 32>(cv, a, st);       // 64 -> 128 as two output slices
  // 1x1
  TTK_UMMA(1, 1, 64, 32, 4, 2, 0)     // bottleneck conv1, fuse 64 -> 32
  TTK_UMMA(1, 1, 32, 128, 2, 3, 1)    // bottleneck conv3 (+ projection shortcut as residual)
  TTK_UMMA(1, 1, 64, 128, 2, 3, 0)    // bottleneck projection shortcut
  TTK_UMMA(1, 1, 32, 16, 4, 3, 0)     // fuse layers (low -> high resolution)
  TTK_UMMA(1, 1, 64, 16, 4, 3, 0)
  TTK_UMMA(1, 1, 128, 16, 4, 2, 0)
  TTK_UMMA(1, 1, 128, 32, 4, 2, 0)
  TTK_UMMA(1, 1, 128, 64, 2, 2, 0)
#undef TTK_UMMA
  return TTK_ERR_UNSUPPORTED;
}
