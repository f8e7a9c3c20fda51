// Implicit-GEMM convolution (3x3 stride 1/2, 1x1) on the 5th-generation tensor cores:
// tcgen05.mma with TMEM accumulators, operands staged by TMA, fused bias / residual / ReLU epilogue.
// Two arithmetic classes share the kernel (template parameter ESZ = bytes per activation element):
//   ESZ 2: bf16 activations and weights, kind::f16;
//   ESZ 4: fp32 activations in HBM, kind::tf32 -- the class cuDNN uses for the reference's convolutions on a GPU (torch enables
//          TF32 for cuDNN by default).  The tensor map of the input has data type TFLOAT32, so TMA rounds every element to TF32
//          (nearest-even) on its way into shared memory; the weights are rounded once on the host; kind::tf32 ignores the low 13
//          mantissa bits, which are zero by then (tools/tf32_probe.cu, profiles/r02_tf32_probe.log).  Accumulation, bias,
//          residual adds and ReLU are fp32, and the stored activations are full fp32 like the reference's.
// Reference ops: the Conv2d + BatchNorm2d(eval) + ReLU (+ residual / fuse sum) groups of
// balldetection/models/wasb.py:35-105, :179-245, :383-416.
//
// Mapping (NHWC bf16 activations):
//   M = 128 consecutive output pixels of one image row, N = Cout (or a 64-wide slice of it), K = taps x Cin.
//   A CTA tile is R output rows x 128 pixels.  ONE TMA box brings the (R+2) x 130 pixel halo of a
//   K-chunk (<= 128 bytes of channels = one swizzled row of 32/64/128 bytes per pixel) into shared memory;
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

#include <string.h>

#include <algorithm>
#include <vector>

#include "hrnet.h"
#include "umma_prims.h"

namespace {

using namespace umma;

constexpr int BW = 128;          // output pixels per tile row = MMA M
constexpr int THREADS = 192;         // TMA warp, MMA warp, 4 epilogue warps
constexpr int THREADS_HF = 320;      // horizontal tap fusion: 8 epilogue warps (two per TMEM lane quarter, alternating output rows)

constexpr int pow2_cols(int c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }
constexpr int al1024(int b) { return (b + 1023) & ~1023; }

// KS: 1 or 3; S: stride 1 or 2 (3x3 only); CIN: padded input channels; COUT: output channels handled by one CTA
// (blockIdx.y selects the slice when the layer has more); R: output rows per tile; STAGES: smem ring depth;
// RB: reserve shared memory for TMA-staged residual tiles (double buffered with the accumulators);
// KCO: channels per K-chunk when not the default (a narrower chunk buys a taller tile for the same shared memory).
template <int KS, int S, int CIN, int COUT, int R, int STAGES, int RB, int KCO = 0, int ESZ = 2, int HF = 0>
struct Cfg {
  static constexpr int KCMAX = 128 / ESZ;                           // a swizzled smem row holds at most 128 bytes
  static constexpr int KC = KCO ? KCO : (CIN < KCMAX ? CIN : KCMAX);      // channels per K-chunk = one swizzled smem row
  static constexpr int NKC = CIN / KC;
  static constexpr int ROWB = KC * ESZ;
  static constexpr int KSTEPS = ROWB / 32;                          // MMAs per row: K = 16 bf16 or 8 tf32 = 32 bytes each
  static constexpr int PAD = KS / 2;
  static constexpr bool FUSE = KS == 3 && S == 1 && 3 * COUT <= 256;   // vertical tap fusion
  static constexpr int NPY = S == 2 ? 2 : 1;            // row-parity units per K-chunk
  static constexpr int NBOX = S == 2 ? 2 : 1;           // TMA boxes per unit (column parities)
  static constexpr int TW = S == 2 ? BW + 1 : (HF ? BW : BW + 2 * PAD);
  static constexpr int OUTW = HF ? BW - 2 : BW;          // output pixels per tile row (horizontal tap fusion: the edge lanes are halo only)
  static constexpr int TR = S == 2 ? R + 1 : R + 2 * PAD;
  static constexpr int BOX_BYTES = TR * TW * ROWB;
  static constexpr int BOX_AL = al1024(BOX_BYTES);
  static constexpr int STAGE_BYTES = NBOX * BOX_AL;
  static constexpr int UNITS = NKC * NPY;               // load units per tile
  static constexpr int TAPS = KS * KS;
  static constexpr int W_BYTES = TAPS * NKC * COUT * ROWB;
  static constexpr int W_BYTES_AL = al1024(W_BYTES);
  static constexpr int ACC_COLS = (HF ? 3 : 1) * R * COUT;      // HF: three accumulator sets (one per horizontal tap) per output row
  static constexpr int TMEM_COLS = pow2_cols(2 * ACC_COLS);
  // residual staging: boxes of CB channels (<= 128 swizzled bytes per pixel) x 128 px x R rows
  static constexpr int CB = COUT < KCMAX ? COUT : KCMAX;
  static constexpr int NRB = COUT / CB;
  static constexpr int RROWB = CB * ESZ;
  static constexpr int RBOX_BYTES = R * BW * RROWB;     // multiple of 1024
  static constexpr int RES_BYTES = RB ? 2 * NRB * RBOX_BYTES : 0;
  static constexpr uint32_t RSWZ = RROWB == 32 ? 1u : RROWB == 64 ? 3u : 7u;
  static constexpr int XCH_BYTES = HF ? 4 * 2 * R * COUT * 4 : 0;      // HF: edge-lane exchange between the four epilogue warps
  static constexpr int SMEM_BASE = 1024 + W_BYTES_AL + STAGES * STAGE_BYTES + RES_BYTES + COUT * 4 + XCH_BYTES + 256;
  // STG (fp32 layers with >= 32 output channels per CTA, where shared memory allows): a column group (32 channels = 128 contiguous
  // bytes of a pixel) is staged per warp in shared memory and written out row-contiguous -- 8 pixels x 128 bytes per store
  // instruction instead of 32 pixels x 32 bytes, a quarter of the L1 wavefronts.  The wide layers' epilogues are bound by exactly
  // those: stem conv1 ran at 64 % of the HBM rate with l1tex 95 % busy.
  static constexpr int STG_WARP_BYTES = 32 * 128;
  static constexpr bool STG_WANTED = S == 1 && !RB && !HF && ESZ == 4 && COUT % 32 == 0 && R * COUT >= 64;
  static constexpr bool STG = STG_WANTED && SMEM_BASE + 128 + 4 * STG_WARP_BYTES <= 232448;
  static constexpr int SMEM_BYTES = SMEM_BASE + (STG ? 128 + 4 * STG_WARP_BYTES : 0);
  static constexpr uint32_t LAYOUT = ROWB == 32 ? 6u : ROWB == 64 ? 4u : 2u;     // SWIZZLE_32B / 64B / 128B
  static constexpr uint32_t SWZ = ROWB == 32 ? 1u : ROWB == 64 ? 3u : 7u;
  static_assert(S == 1 || KS == 3, "stride 2 is implemented for 3x3 only");
  static_assert(!HF || (KS == 3 && S == 1 && 9 * COUT <= 256 && COUT == 16 && !RB), "horizontal tap fusion: 3x3 stride 1, 16 output channels, no staged residual");
  static_assert(2 * ACC_COLS <= 512, "accumulators exceed TMEM");
  // software-pipelined epilogue (see the kernel): stride-1 layers without TMA-staged residual tiles, whose shared-memory reads have no
  // latency to hide (the full-resolution 16 -> 16 layers ran at 6.5 TB/s without it, 5.7 TB/s with it).  TF32: all of them; bf16: the
  // wide layers (64+ output channels per CTA), whose epilogue -- one warp per scheduler -- is what bounds them
  static constexpr bool PIPE = S == 1 && !RB && !HF && ACC_COLS >= 64 && (ESZ == 4 || COUT >= 64);
  static constexpr int GMAX = (ESZ == 2 && !PIPE) ? 64 : 32;
  static constexpr int GCOLS = ACC_COLS < GMAX ? ACC_COLS : GMAX;     // accumulator columns fetched per TMEM wait
  static_assert(ESZ == 2 || ESZ == 4, "bf16 or tf32-in-fp32 elements");
  static_assert(ACC_COLS % GCOLS == 0, "the epilogue drains whole groups of accumulator columns");
  static_assert(ROWB == 32 || ROWB == 64 || ROWB == 128, "a K-chunk is one 32/64/128-byte swizzled row");
  static_assert(COUT % 16 == 0 && COUT <= 256 && CIN % 16 == 0, "bad channel counts");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

struct KMaps {
  CUtensorMap m[4];            // stride 1: m[0]; stride 2: m[py * 2 + px] (parity sub-grids)
  CUtensorMap res;             // residual tensor (same shape as the output), when staged by TMA
};

struct KArgs {
  const void* w;               // packed weights, see ttk_conv_umma_pack
  const float* bias;
  void* out;
  const void* res[3];
  int rsh[3];
  int nres;
  int res_tma;                 // res[0] (shift 0) arrives through shared memory
  int dual;                    // 1x1 only, two K-concatenated inputs: the first `dual` K-chunks come from tensor map m[0], the rest from m[1]
  int n, h, w_img, cout_total;
  int relu;
  int tiles_x, tiles_y, total;
};

template <int KS, int S, int CIN, int COUT, int R, int STAGES, int RB, int KCO = 0, int ESZ = 2, int HF = 0>
__global__ void __launch_bounds__(HF ? THREADS_HF : THREADS, 1) conv_umma_kernel(const __grid_constant__ KMaps maps, const KArgs a) {
  constexpr int NTHR = HF ? THREADS_HF : THREADS;
  using C = Cfg<KS, S, CIN, COUT, R, STAGES, RB, KCO, ESZ, HF>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;
  uint8_t* sA = smem + C::W_BYTES_AL;
  uint8_t* sR = sA + STAGES * C::STAGE_BYTES;           // [acc][box][row][px][CB] swizzled (1024-aligned)
  float* sBias = reinterpret_cast<float*>(sR + C::RES_BYTES);
  float* sXch = sBias + COUT;                           // HF: [warp][set 0 of lane 31 | set 2 of lane 0][R][COUT]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sXch + C::XCH_BYTES / 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 6);
  // STG: per-warp staging rows of 128 bytes, 128-byte aligned (behind the barriers' 256 bytes)
  uint8_t* sStage = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(bars) + 256 + 127) & ~(uintptr_t)127);
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
    for (int i = tid; i < C::W_BYTES / 16; i += NTHR) {
      uint32_t addr = wbase + (i / CPR) * C::ROWB + (i % CPR) * 16;
      addr ^= ((addr >> 7) & C::SWZ) << 4;
      *reinterpret_cast<uint4*>(sW + (addr - wbase)) = __ldg(wsrc + i);
    }
    for (int i = tid; i < COUT; i += NTHR) sBias[i] = a.bias[n_off + i];
  }
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, HF ? 8 : 4);
      mbar_init(bar_rfull + 8 * i, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_async_smem();              // weights: generic-proxy writes -> async proxy (tensor core)
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), C::TMEM_COLS);
  fence_before();
  __syncthreads();
  fence_after();
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
            if (KS == 1 && a.dual)     // channels beyond the first tensor are zero filled
              tma_load_4d(dst, &maps.m[kc < a.dual ? 0 : 1], bar_full + 8 * s, (kc < a.dual ? kc : kc - a.dual) * C::KC, tx * BW, ty * R, img);
            else
              tma_load_4d(dst, &maps.m[0], bar_full + 8 * s, kc * C::KC, tx * C::OUTW - C::PAD, ty * R - C::PAD, img);
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
    // The WHOLE warp walks the loops with warp-uniform values and one elected lane issues: addresses, descriptors and the TMEM
    // column then live in uniform registers and consecutive tcgen05.mma are 1-3 instructions apart.  With `if (lane == 0)` around
    // the loops the compiler treated every operand as divergent and wrapped each MMA in an ELECT / R2UR.BROADCAST loop (16
    // instructions, ~80 clk per MMA against the tensor pipe's 45 clk for a thin one -- the thin layers were issue bound).
    {
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      const bool leader = elect_one();
      const uint32_t wbase = smem_u32(sW) >> 4;
      auto issue = [leader](uint32_t d, uint64_t da, uint64_t db, uint32_t idesc) {      // every MMA accumulates (the epilogue zeroes what it drains)
        if (leader) {
          if (ESZ == 2) mma(d, da, db, idesc, 1u);
          else mma_tf32(d, da, db, idesc, 1u);
        }
      };
      uint32_t it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < a.total; tile += gridDim.x, ++tcount) {
        const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
        mbar_wait(bar_tempty + 8 * acc, aph);      // accumulators drained and zeroed by the epilogue
        fence_after();
        const uint32_t d_acc = tmem_u + acc * C::ACC_COLS;
        for (int u = 0; u < C::UNITS; ++u, ++it) {
          const int kc = u / C::NPY, py = u % C::NPY;
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(bar_full + 8 * s, ph);
          fence_after();
          const uint32_t abase = smem_u32(sA + s * C::STAGE_BYTES) >> 4;      // 16-byte units from here on (make_desc16)
          constexpr uint32_t RB16 = C::ROWB / 16;
          if (HF) {
            // horizontal + vertical tap fusion: input row hr (lane = input pixel, no shift) times [W(ky, kx) for all kx, ky in k0..k1]:
            // N = 3 (k1 - k0 + 1) Cout lands in the accumulator sets [row block][kx][Cout]; the epilogue adds the three sets of a row
            // with a lane shift of kx - 1
#pragma unroll 1
            for (int hr = 0; hr < R + 2; ++hr) {
              const int yi = hr - 1;
              const int k0 = yi + 2 - R > 0 ? yi + 2 - R : 0;
              const int k1 = yi + 1 < 2 ? yi + 1 : 2;
              const uint32_t idesc = ESZ == 2 ? make_idesc(128, (k1 - k0 + 1) * 3 * COUT) : make_idesc_tf32(128, (k1 - k0 + 1) * 3 * COUT);
              const uint32_t d_tmem = d_acc + (R - 2 - yi + k0) * 3 * COUT;
              const uint32_t arow = abase + hr * C::TW * RB16;
              const uint32_t brow = wbase + ((kc * 3 + k0) * 3 * COUT) * RB16;
#pragma unroll
              for (int ks = 0; ks < C::KSTEPS; ++ks)
                issue(d_tmem, make_desc16<8 * C::ROWB, C::LAYOUT>(arow + ks * 2), make_desc16<8 * C::ROWB, C::LAYOUT>(brow + ks * 2), idesc);
            }
          } else if (C::FUSE) {
            // input (halo) row hr = yi + 1 feeds output rows yo = yi + 1 - ky; row yo lives in column block R-1-yo
#pragma unroll 1
            for (int hr = 0; hr < R + 2; ++hr) {
              const int yi = hr - 1;
              const int k0 = yi + 2 - R > 0 ? yi + 2 - R : 0;
              const int k1 = yi + 1 < 2 ? yi + 1 : 2;
              const uint32_t idesc = ESZ == 2 ? make_idesc(128, (k1 - k0 + 1) * COUT) : make_idesc_tf32(128, (k1 - k0 + 1) * COUT);
              const uint32_t d_tmem = d_acc + (R - 2 - yi + k0) * COUT;
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                const uint32_t arow = abase + (hr * C::TW + kx) * RB16;
                const uint32_t brow = wbase + (((kc * 3 + kx) * 3 + k0) * COUT) * RB16;
#pragma unroll
                for (int ks = 0; ks < C::KSTEPS; ++ks)
                  issue(d_tmem, make_desc16<8 * C::ROWB, C::LAYOUT>(arow + ks * 2), make_desc16<8 * C::ROWB, C::LAYOUT>(brow + ks * 2), idesc);
              }
            }
          } else {
            constexpr uint32_t idesc = ESZ == 2 ? make_idesc(128, COUT) : make_idesc_tf32(128, COUT);
#pragma unroll 1
            for (int r = 0; r < R; ++r) {
              const uint32_t d_tmem = d_acc + (R - 1 - r) * COUT;
#pragma unroll
              for (int tap = 0; tap < C::TAPS; ++tap) {
                const int ky = tap / KS, kx = tap % KS;
                uint32_t arow;
                if (S == 1) {
                  arow = abase + ((r + ky) * C::TW + kx) * RB16;
                } else {
                  if ((ky != 1 ? 1 : 0) != py) continue;     // this unit holds the other row parity
                  const int px = kx != 1 ? 1 : 0;
                  arow = abase + px * (C::BOX_AL / 16) + ((r + (ky == 2 ? 1 : 0)) * C::TW + (kx == 2 ? 1 : 0)) * RB16;
                }
                const uint32_t brow = wbase + ((tap * C::NKC + kc) * COUT) * RB16;
#pragma unroll
                for (int ks = 0; ks < C::KSTEPS; ++ks)
                  issue(d_tmem, make_desc16<8 * C::ROWB, C::LAYOUT>(arow + ks * 2), make_desc16<8 * C::ROWB, C::LAYOUT>(brow + ks * 2), idesc);
              }
            }
          }
          if (leader) commit(bar_empty + 8 * s);          // smem stage reusable once these MMAs have read it
        }
        if (leader) commit(bar_tfull + 8 * acc);          // accumulators complete
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue warps 2..5: TMEM lane quarter (warp % 4) =====
    const int q = warp & 3;
    const int m = q * 32 + lane;                   // pixel within the tile row = TMEM lane
    constexpr int GCOLS = C::GCOLS;                // accumulator columns fetched per TMEM wait
    constexpr int NSUB = GCOLS / 16;
    constexpr int RW = 4 * ESZ;                    // 32-bit words of the 16 channels of a pixel: 8 (bf16, one sector) or 16 (fp32, two)
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    const int eg = (warp - 2) >> 2;                // HF: epilogue group 0 / 1 takes the even / odd output rows of a tile
    // zero both accumulator buffers, then open them for the MMA warp
    if (eg == 0)
      for (int c = 0; c < 2 * C::ACC_COLS; c += 16) tmem_zero16(lane_base + c);
    tmem_wait_st();
    fence_before();
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(bar_tempty);
      mbar_arrive(bar_tempty + 8);
    }
    // f += the 16 channels held in w: bf16 pairs are widened by bit placement (keeps the conversion pipe free)
    auto add_words = [](float* f, const uint32_t* w) {
      if (ESZ == 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          f[2 * j] += __uint_as_float(w[j] << 16);
          f[2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] += __uint_as_float(w[j]);
      }
    };
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < a.total; tile += gridDim.x, ++tcount) {
      const int tx = tile % a.tiles_x, ty = (tile / a.tiles_x) % a.tiles_y, img = tile / (a.tiles_x * a.tiles_y);
      const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
      const int ox = HF ? tx * C::OUTW + m - 1 : tx * BW + m;      // HF: lane m is input pixel tx OUTW - 1 + m and output pixel of the same index
      if (res_tma) mbar_wait(bar_rfull + 8 * acc, aph);
      mbar_wait(bar_tfull + 8 * acc, aph);
      fence_after();
      if constexpr (HF != 0) {
        // ---- horizontal tap fusion: out[r][x] = set0[r][x - 1] + set1[r][x] + set2[r][x + 1]; lanes 0 and 127 are halo only ----
        const uint32_t tbase = lane_base + acc * C::ACC_COLS;
        float* xme = sXch + (q * 2) * R * COUT;
        // pass 1: the edge lanes publish what their neighbours in the adjacent warps need
#pragma unroll 1
        for (int r = eg; r < R; r += 2) {
          uint32_t e0[16], e2[16];
          tmem_ld16(tbase + (R - 1 - r) * 3 * COUT, e0);
          tmem_ld16(tbase + (R - 1 - r) * 3 * COUT + 2 * COUT, e2);
          tmem_wait_ld();
          if (lane == 31) {
#pragma unroll
            for (int j = 0; j < 16; ++j) xme[r * COUT + j] = __uint_as_float(e0[j]);
          }
          if (lane == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) xme[R * COUT + r * COUT + j] = __uint_as_float(e2[j]);
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float* xprev = sXch + ((q > 0 ? q - 1 : 0) * 2) * R * COUT;              // set 0 of the previous warp's lane 31
        const float* xnext = sXch + ((q < 3 ? q + 1 : 3) * 2 + 1) * R * COUT;          // set 2 of the next warp's lane 0
#pragma unroll 1
        for (int r = eg; r < R; r += 2) {
          uint32_t v0[16], v1[16], v2[16];
          const uint32_t ta = tbase + (R - 1 - r) * 3 * COUT;
          tmem_ld16(ta, v0);
          tmem_ld16(ta + COUT, v1);
          tmem_ld16(ta + 2 * COUT, v2);
          tmem_wait_ld();
          tmem_zero16(ta);
          tmem_zero16(ta + COUT);
          tmem_zero16(ta + 2 * COUT);
          const int oy = ty * R + r;
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float left = __shfl_up_sync(0xffffffffu, __uint_as_float(v0[j]), 1);
            float right = __shfl_down_sync(0xffffffffu, __uint_as_float(v2[j]), 1);
            if (lane == 0) left = xprev[r * COUT + j];
            if (lane == 31) right = xnext[r * COUT + j];
            f[j] = (left + __uint_as_float(v1[j])) + right + sBias[j];
          }
          const bool live = m >= 1 && m <= C::OUTW && ox < a.w_img && oy < a.h;
          for (int rr = 0; rr < a.nres; ++rr) {            // (the layer this path serves has no residual; kept for the one-conv test hook)
            if (!live) break;
            const int sh = a.rsh[rr];
            const char* rp = (const char*)a.res[rr] +
                             ((((size_t)img * (a.h >> sh) + (oy >> sh)) * (a.w_img >> sh) + (ox >> sh)) * a.cout_total + n_off) * ESZ;
            uint32_t rw[RW];
#pragma unroll
            for (int i = 0; i < RW / 8; ++i) ldg256(rp + 32 * i, &rw[8 * i]);
            add_words(f, rw);
          }
          if (a.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          if (live) {
            char* op = (char*)a.out + ((((size_t)img * a.h + oy) * a.w_img + ox) * a.cout_total + n_off) * ESZ;
            if (ESZ == 2) {
              uint32_t o[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                __nv_bfloat162 b2 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                o[j] = *reinterpret_cast<uint32_t*>(&b2);
              }
              stg256(op, o);
            } else {
              uint32_t o[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(f[j]);
              stg256(op, o);
              stg256(op + 32, o + 8);
            }
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");        // the exchange buffer is free for the next tile
      } else if constexpr (C::PIPE) {
        // Software-pipelined drain (stride-1 layers without staged residual tiles, at most one residual): the TMEM load and the residual loads of column
        // group g + 1 are in flight while group g is added up and stored, so neither latency is exposed per group (stem conv1 ran at
        // 59 % of the HBM rate waiting for one tcgen05.ld at a time; residuals read from global memory cost ~1 us per group).
        constexpr int G = C::ACC_COLS / GCOLS;
        uint32_t v[2][NSUB][16], rv[2][NSUB][RW];
        const int nres = a.nres;
        const uint32_t tbase = lane_base + acc * C::ACC_COLS;
        auto fetch = [&](int g, uint32_t (&vv)[NSUB][16], uint32_t (&rr)[NSUB][RW]) {
#pragma unroll
          for (int sb = 0; sb < NSUB; ++sb) tmem_ld16(tbase + g * GCOLS + sb * 16, vv[sb]);
#pragma unroll
          for (int sb = 0; sb < NSUB; ++sb) {
            const int col = g * GCOLS + sb * 16, r = R - 1 - col / COUT, c0 = col % COUT;
            const int oy = ty * R + r;
#pragma unroll
            for (int j = 0; j < RW; ++j) rr[sb][j] = 0;
            if (nres > 0) {
              if (res_tma) {
                const uint32_t row = smem_u32(sR + (acc * C::NRB + c0 / C::CB) * C::RBOX_BYTES) + (r * BW + m) * C::RROWB;
                const uint32_t swz = ((row >> 7) & C::RSWZ) << 4;
#pragma unroll
                for (int i = 0; i < RW / 4; ++i) lds128((row + (c0 % C::CB) * ESZ + 16 * i) ^ swz, &rr[sb][4 * i]);
              } else if (ox < a.w_img && oy < a.h) {
                const int sh = a.rsh[0];
                const char* rp = (const char*)a.res[0] +
                                 ((((size_t)img * (a.h >> sh) + (oy >> sh)) * (a.w_img >> sh) + (ox >> sh)) * a.cout_total + n_off + c0) * ESZ;
#pragma unroll
                for (int i = 0; i < RW / 8; ++i) ldg256(rp + 32 * i, &rr[sb][8 * i]);
              }
            }
          }
        };
        fetch(0, v[0], rv[0]);
        tmem_wait_ld();
#pragma unroll
        for (int sb = 0; sb < NSUB; ++sb) tmem_zero16(tbase + sb * 16);
#pragma unroll
        for (int g = 0; g < G; ++g) {
          if (g + 1 < G) fetch(g + 1, v[(g + 1) & 1], rv[(g + 1) & 1]);
          if (C::STG && nres == 0) {         // (with a residual the staged form measured slower: 0.74 vs 0.60 ms for the half-resolution conv2 layers)
            // the group's 32 channels of this lane's pixel -> staging row `lane` (16-byte chunk c at slot c ^ (lane & 7): conflict free),
            // then 4 store instructions of 8 pixels x 128 contiguous bytes each (lane = pixel it * 8 + lane / 4, 32-byte piece lane % 4)
            const int col = g * GCOLS, r = R - 1 - col / COUT, c0 = col % COUT;
            const int oy = ty * R + r;
            const uint32_t srow = smem_u32(sStage) + (warp - 2) * C::STG_WARP_BYTES;
#pragma unroll
            for (int sb = 0; sb < NSUB; ++sb) {
              float f[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[g & 1][sb][j]) + sBias[c0 + sb * 16 + j];
              if (nres > 0) add_words(f, rv[g & 1][sb]);
              if (a.relu) {
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
              }
#pragma unroll
              for (int u = 0; u < 4; ++u)
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(srow + lane * 128 + (((sb * 4 + u) ^ (lane & 7)) << 4)), "f"(f[4 * u]),
                             "f"(f[4 * u + 1]), "f"(f[4 * u + 2]), "f"(f[4 * u + 3])
                             : "memory");
            }
            __syncwarp();
            const int pq = lane & 3;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int p = it * 8 + (lane >> 2);
              uint32_t o[8];
              lds128(srow + p * 128 + (((2 * pq) ^ (p & 7)) << 4), o);
              lds128(srow + p * 128 + (((2 * pq + 1) ^ (p & 7)) << 4), o + 4);
              const int oxp = tx * BW + (warp & 3) * 32 + p;
              if (oxp < a.w_img && oy < a.h)
                stg256((char*)a.out + ((((size_t)img * a.h + oy) * a.w_img + oxp) * a.cout_total + n_off + c0) * 4 + pq * 32, o);
            }
            __syncwarp();
          } else
#pragma unroll
          for (int sb = 0; sb < NSUB; ++sb) {
            const int col = g * GCOLS + sb * 16, r = R - 1 - col / COUT, c0 = col % COUT;
            const int oy = ty * R + r;
            if (!(ox < a.w_img && oy < a.h)) continue;
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[g & 1][sb][j]) + sBias[c0 + j];
            if (nres > 0) add_words(f, rv[g & 1][sb]);
            if (a.relu) {
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            char* op = (char*)a.out + ((((size_t)img * a.h + oy) * a.w_img + ox) * a.cout_total + n_off + c0) * ESZ;
            if (ESZ == 2) {
              uint32_t o[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                __nv_bfloat162 b2 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                o[j] = *reinterpret_cast<uint32_t*>(&b2);
              }
              stg256(op, o);
            } else {
              uint32_t o[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(f[j]);
              stg256(op, o);
              stg256(op + 32, o + 8);
            }
          }
          if (g + 1 < G) {
            tmem_wait_ld();
#pragma unroll
            for (int sb = 0; sb < NSUB; ++sb) tmem_zero16(tbase + (g + 1) * GCOLS + sb * 16);     // leave the columns zeroed for the next tile
          }
        }
      } else {
#pragma unroll 1
      for (int g0 = 0; g0 < C::ACC_COLS; g0 += GCOLS) {
        uint32_t v[NSUB][16];
        const uint32_t taddr = lane_base + acc * C::ACC_COLS + g0;
#pragma unroll
        for (int sb = 0; sb < NSUB; ++sb) tmem_ld16(taddr + sb * 16, v[sb]);
        // residual operands for the same columns, all fetched before the TMEM wait so that the latencies overlap (the
        // lower-resolution fuse terms used to be loaded one by one at their point of use: 0.25 ms per stage-3/4 fuse conv)
        uint32_t rv[NSUB][RW], rx[NSUB][2][RW];
        const int nres = a.nres;
#pragma unroll
        for (int sb = 0; sb < NSUB; ++sb) {
          const int col = g0 + sb * 16, r = R - 1 - col / COUT, c0 = col % COUT;
          const int oy = ty * R + r;
          const bool live = ox < a.w_img && oy < a.h;
#pragma unroll
          for (int j = 0; j < RW; ++j) rv[sb][j] = 0;
          if (nres > 0) {
            if (res_tma) {
              const uint32_t row = smem_u32(sR + (acc * C::NRB + c0 / C::CB) * C::RBOX_BYTES) + (r * BW + m) * C::RROWB;
              const uint32_t swz = ((row >> 7) & C::RSWZ) << 4;       // a pixel row never crosses a 128-byte line: one XOR for all its chunks
#pragma unroll
              for (int i = 0; i < RW / 4; ++i) lds128((row + (c0 % C::CB) * ESZ + 16 * i) ^ swz, &rv[sb][4 * i]);
            } else if (live) {
              const int sh = a.rsh[0];
              const char* rp = (const char*)a.res[0] +
                               ((((size_t)img * (a.h >> sh) + (oy >> sh)) * (a.w_img >> sh) + (ox >> sh)) * a.cout_total + n_off + c0) * ESZ;
#pragma unroll
              for (int i = 0; i < RW / 8; ++i) ldg256(rp + 32 * i, &rv[sb][8 * i]);
            }
#pragma unroll
            for (int rr = 1; rr < 3; ++rr) {
              if (rr < nres && live) {
                const int sh = a.rsh[rr];
                const char* rp = (const char*)a.res[rr] +
                                 ((((size_t)img * (a.h >> sh) + (oy >> sh)) * (a.w_img >> sh) + (ox >> sh)) * a.cout_total + n_off + c0) * ESZ;
#pragma unroll
                for (int i = 0; i < RW / 8; ++i) ldg256(rp + 32 * i, &rx[sb][rr - 1][8 * i]);
              }
            }
          }
        }
        tmem_wait_ld();
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
            add_words(f, rv[sb]);
#pragma unroll
            for (int rr = 1; rr < 3; ++rr) {
              if (rr >= nres) break;
              add_words(f, rx[sb][rr - 1]);
            }
          }
          if (a.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          // 256-bit stores: the 16 channels of a pixel are one (bf16) or two (fp32) whole 32-byte sectors
          char* op = (char*)a.out + ((((size_t)img * a.h + oy) * a.w_img + ox) * a.cout_total + n_off + c0) * ESZ;
          if (ESZ == 2) {
            uint32_t o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              __nv_bfloat162 b2 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
              o[j] = *reinterpret_cast<uint32_t*>(&b2);
            }
            stg256(op, o);
          } else {
            uint32_t o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(f[j]);
            stg256(op, o);
            stg256(op + 32, o + 8);
          }
        }
      }
      }
      tmem_wait_st();
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, C::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
// output channels one CTA handles: 128-wide 3x3 layers with 64+ input channels are split so the weights fit in shared memory
int cout_tile(const TtkConv& cv, int esz) { return (cv.k == 3 && cv.cout_p == 128 && cv.cin_p >= 64) ? (esz == 2 ? 64 : 32) : cv.cout_p; }
bool fused_ky(const TtkConv& cv, int esz) { return cv.k == 3 && cv.stride == 1 && 3 * cout_tile(cv, esz) <= 256; }
// horizontal + vertical tap fusion (one MMA per input row and K step, N = 9 Cout): the thin layer that is bound by its MMA count
// -- TF32 only: with K = 8 per instruction it has twice the MMAs of the bf16 kernel (4.48 -> 3.50 ms per 32 stacks); in bf16 the extra
// epilogue work (three accumulator sets, two shuffles per value, edge-lane exchange) outweighs the saved MMAs (1.74 -> 3.00 ms)
bool fused_h(const TtkConv& cv, int esz) { return esz == 4 && cv.k == 3 && cv.stride == 1 && cv.cin_p == 128 && cv.cout_p == 16; }
// channels per K-chunk (one swizzled shared-memory row).  The stride-1 3x3 layers with 64+ input channels are bound by their MMA
// count (a 128-pixel x 32-byte MMA holds the tensor pipe 45.5 clk at N <= 48 and N / 2 clk from N = 96 on, tools/tf32_probe.cu) and
// the halo rows of a tile are pure overhead: 3 (R + 2) / R MMAs per K step and output row.  Narrow chunks shrink the staging boxes so
// that taller tiles fit: bf16 transition1.0 R = 3 -> 8, the 64 -> 64 layers R = 2 -> 4, and the 128 -> 128 layers get a second
// pipeline stage.  With fp32 elements the 3x3 weights of the wide layers take up to 147 KB, which leaves room for 8-channel chunks.
int kc_of(const TtkConv& cv, int esz) {
  if (esz == 2) {
    if (cv.k == 3 && cv.stride == 1 && cv.cin_p >= 64) return 32;
    // stride 2 with 64+ input channels (transition1.1 1.74 -> 1.49 ms): two-row tiles and a third pipeline stage in the same shared memory
    if (cv.k == 3 && cv.stride == 2 && cv.cin_p >= 64) return 32;
    return cv.cin_p < 64 ? cv.cin_p : 64;
  }
  if (cv.k == 3 && cv.stride == 1 && cv.cin_p == 128 && cv.cout_p == 16) return 16;      // transition1.0: 64-byte rows fill faster than 32-byte ones
  if (cv.k == 3 && cv.cin_p >= 64) return 8;
  if (cv.k == 3 && cv.stride == 2 && cv.cin_p == 32 && cv.cout_p == 128) return 8;
  if (cv.k == 3) return 16;
  return 32;
}

CUtensorMapSwizzle swizzle_for(int row_bytes) {
  return row_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
}

float round_tf32(float v) {      // nearest-even on the 13 dropped mantissa bits (finite inputs)
  uint32_t u;
  memcpy(&u, &v, 4);
  u += 0xfffu + ((u >> 13) & 1u);
  u &= ~0x1fffu;
  memcpy(&v, &u, 4);
  return v;
}

template <int KS, int S, int CIN, int COUT, int R, int STAGES, int RB, int KCO = 0, int ESZ = 2, int HF = 0>
int launch(const TtkConv& cv, const ConvLaunch& a, cudaStream_t st) {
  using C = Cfg<KS, S, CIN, COUT, R, STAGES, RB, KCO, ESZ, HF>;
  if (C::KC != kc_of(cv, ESZ) || COUT != cout_tile(cv, ESZ) || (HF != 0) != fused_h(cv, ESZ)) {
    ttk_set_error("conv %s: kernel K-chunk %d / output slice %d differ from the packed weights' %d / %d", cv.name.c_str(), C::KC, COUT,
                  kc_of(cv, ESZ), cout_tile(cv, ESZ));
    return TTK_ERR_STATE;
  }
  EncodeFn encode = get_encode();
  if (!encode) {
    ttk_set_error("cuTensorMapEncodeTiled is not available from the driver");
    return TTK_ERR_CUDA;
  }
  static TtkPerDevice attr;
  if (attr.first()) {
    TTK_CUDA(cudaFuncSetAttribute(conv_umma_kernel<KS, S, CIN, COUT, R, STAGES, RB, KCO, ESZ, HF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  C::SMEM_BYTES));
  }
  if (S == 2 && ((a.hin & 1) || (a.win & 1))) return TTK_ERR_UNSUPPORTED;
  // the input map's TFLOAT32 type makes TMA round fp32 to TF32 (nearest-even) while it fills shared memory
  const CUtensorMapDataType in_type = ESZ == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32;
  const CUtensorMapDataType res_type = ESZ == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  KMaps maps;
  cuuint32_t box[4] = {(cuuint32_t)C::KC, (cuuint32_t)C::TW, (cuuint32_t)C::TR, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  for (int i = 0; i < (S == 2 ? 4 : 1); ++i) {
    const int py = i >> 1, px = i & 1;
    cuuint64_t dims[4] = {(cuuint64_t)CIN, (cuuint64_t)(a.win / S), (cuuint64_t)(a.hin / S), (cuuint64_t)a.n};
    cuuint64_t strides[3] = {(cuuint64_t)CIN * ESZ * S, (cuuint64_t)a.win * CIN * ESZ * S, (cuuint64_t)a.hin * a.win * CIN * ESZ};
    void* base = (char*)const_cast<void*>(a.in) + ((size_t)py * a.win + px) * CIN * ESZ;
    const CUresult r = encode(&maps.m[i], in_type, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(C::ROWB),
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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
    cuuint64_t strides[3] = {(cuuint64_t)a.cout * ESZ, (cuuint64_t)a.wout * a.cout * ESZ, (cuuint64_t)a.hout * a.wout * a.cout * ESZ};
    cuuint32_t rbox[4] = {(cuuint32_t)C::CB, (cuuint32_t)BW, (cuuint32_t)R, 1};
    const CUresult r = encode(&maps.res, res_type, 4, const_cast<void*>(a.res[0]), dims, strides, rbox, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              swizzle_for(C::RROWB), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      ttk_set_error("cuTensorMapEncodeTiled (residual) failed (%d) for conv %s", (int)r, cv.name.c_str());
      return TTK_ERR_CUDA;
    }
  }
  KArgs k;
  k.w = ESZ == 2 ? (const void*)cv.w_umma : (const void*)cv.w_umma32;
  k.bias = cv.bias;
  k.out = a.out;
  for (int i = 0; i < 3; ++i) {
    k.res[i] = a.res[i];
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
  k.tiles_x = ttk_cdiv(a.wout, C::OUTW);
  k.tiles_y = ttk_cdiv(a.hout, R);
  k.total = k.tiles_x * k.tiles_y * a.n;
  const int nsplit = a.cout / COUT;
  // persistent CTAs: as many as fit per SM (shared memory and TMEM columns), never more than there are tiles
  int occ = std::min(232448 / (C::SMEM_BYTES + 1024), 512 / C::TMEM_COLS);
  occ = std::max(1, std::min(occ, 2));
  const int gx = std::max(1, std::min(k.total, ttk_num_sms() * occ / nsplit));
  dim3 grid(gx, nsplit);
  conv_umma_kernel<KS, S, CIN, COUT, R, STAGES, RB, KCO, ESZ, HF><<<grid, HF ? THREADS_HF : THREADS, C::SMEM_BYTES, st>>>(maps, k);
  TTK_LAUNCH_CHECK();
  return TTK_OK;
}

// Bottleneck tail: out = relu(W3 a + Wd x + bias) with a (32 channels) and x (64 channels) K-concatenated.  bf16: two 64-channel
// K-chunks (the first box reaches past the 32-channel tensor, TMA zero fills); tf32: three 32-channel chunks (a | x[0:32] | x[32:64]).
template <int ESZ>
int launch_dual(const void* w_dual, const float* bias_dual, const ConvLaunch& a, cudaStream_t st) {
  constexpr int R = 2, STAGES = 3, KC = 128 / ESZ, CINK = ESZ == 2 ? 128 : 96;
  using C = Cfg<1, 1, CINK, 128, R, STAGES, 0, 0, ESZ>;
  EncodeFn encode = get_encode();
  if (!encode) return TTK_ERR_UNSUPPORTED;
  static TtkPerDevice attr;
  if (attr.first()) {
    TTK_CUDA(cudaFuncSetAttribute(conv_umma_kernel<1, 1, CINK, 128, R, STAGES, 0, 0, ESZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  }
  KMaps maps;
  cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)BW, (cuuint32_t)R, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  const void* ins[2] = {a.in, a.in2};
  const int cins[2] = {a.cin, a.cin2};
  for (int i = 0; i < 2; ++i) {
    cuuint64_t dims[4] = {(cuuint64_t)cins[i], (cuuint64_t)a.win, (cuuint64_t)a.hin, (cuuint64_t)a.n};
    cuuint64_t strides[3] = {(cuuint64_t)cins[i] * ESZ, (cuuint64_t)a.win * cins[i] * ESZ, (cuuint64_t)a.hin * a.win * cins[i] * ESZ};
    const CUresult r = encode(&maps.m[i], ESZ == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 4, const_cast<void*>(ins[i]),
                              dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return TTK_ERR_UNSUPPORTED;      // e.g. a driver that rejects a box wider than the tensor
  }
  maps.m[2] = maps.m[3] = maps.res = maps.m[0];
  KArgs k;
  k.w = w_dual;
  k.bias = bias_dual;
  k.out = a.out;
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
  conv_umma_kernel<1, 1, CINK, 128, R, STAGES, 0, 0, ESZ><<<gx, THREADS, C::SMEM_BYTES, st>>>(maps, k);
  TTK_LAUNCH_CHECK();
  return TTK_OK;
}

}  // namespace

int ttk_conv_umma_launch_dual(const void* w_dual, const float* bias_dual, const ConvLaunch& a, cudaStream_t st, int esz) {
  if (a.cin != 32 || a.cin2 != 64 || a.cout != 128) return TTK_ERR_UNSUPPORTED;
  return esz == 2 ? launch_dual<2>(w_dual, bias_dual, a, st) : launch_dual<4>(w_dual, bias_dual, a, st);
}

// Host image of the fused bottleneck tail's B operand: [k-chunk][cout 128][KC], chunk 0 = conv3 (32 input channels, zero padded to KC),
// then the projection shortcut's 64 input channels.  bf16: KC = 64 (2 chunks); tf32: KC = 32 (3 chunks), values rounded to TF32.
void ttk_conv_umma_pack_dual(const float* w3, const float* wd, int esz, std::vector<uint8_t>& out) {
  const int KC = 128 / esz, nk = esz == 2 ? 2 : 3;
  out.assign((size_t)nk * 128 * KC * esz, 0);
  auto put = [&](int kc, int co, int c, float v) {
    const size_t i = ((size_t)kc * 128 + co) * KC + c;
    if (esz == 2) reinterpret_cast<__nv_bfloat16*>(out.data())[i] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(out.data())[i] = round_tf32(v);
  };
  for (int co = 0; co < 128; ++co) {
    for (int ci = 0; ci < 32; ++ci) put(0, co, ci, w3[(size_t)co * 32 + ci]);
    for (int ci = 0; ci < 64; ++ci) put(1 + ci / KC, co, ci % KC, wd[(size_t)co * 64 + ci]);
  }
}

// Weights for the tensor-core path: rows of KC input channels (K-major), zero padded, per output-channel slice:
//   fused 3x3 stride 1 : [slice][k-chunk][kx][ky][cout_tile][KC]   (B = the three vertical taps side by side)
//   transition1.0      : [k-chunk][ky][kx][cout][KC]               (B = all nine taps: horizontal + vertical fusion)
//   otherwise          : [slice][tap][k-chunk][cout_tile][KC]
// bf16 (w_umma) and TF32-rounded fp32 (w_umma32) images are both kept; they differ in KC and in the slice width of the 128-wide layers.
int ttk_conv_umma_pack(TtkConv& cv, const float* w_host) {
  const int kk = cv.k * cv.k;
  for (int esz = 2; esz <= 4; esz += 2) {
    const int KC = kc_of(cv, esz);
    const int nkc = cv.cin_p / KC;
    const int ct = cout_tile(cv, esz);
    const bool fused = fused_ky(cv, esz);
    const size_t count = (size_t)kk * cv.cin_p * cv.cout_p;
    std::vector<uint8_t> w(count * esz, 0);
    for (int co = 0; co < cv.cout; ++co)
      for (int ci = 0; ci < cv.cin; ++ci)
        for (int t = 0; t < kk; ++t) {
          const int kc = ci / KC, c = ci % KC, sl = co / ct, cl = co % ct;
          size_t row;
          if (fused_h(cv, esz)) {
            const int ky = t / 3, kx = t % 3;
            row = (((size_t)sl * nkc + kc) * 3 + ky) * 3 + kx;      // [k-chunk][ky][kx][cout]: B of one input row = all nine taps
          } else if (fused) {
            const int ky = t / 3, kx = t % 3;
            row = (((size_t)sl * nkc + kc) * 3 + kx) * 3 + ky;
          } else {
            row = ((size_t)sl * kk + t) * nkc + kc;
          }
          const float v = w_host[((size_t)co * cv.cin + ci) * kk + t];
          const size_t i = (row * ct + cl) * KC + c;
          if (esz == 2) reinterpret_cast<__nv_bfloat16*>(w.data())[i] = __float2bfloat16_rn(v);
          else reinterpret_cast<float*>(w.data())[i] = round_tf32(v);
        }
    void** dst = esz == 2 ? (void**)&cv.w_umma : (void**)&cv.w_umma32;
    if (!*dst) TTK_CUDA(cudaMalloc(dst, w.size()));
    TTK_CUDA(cudaMemcpy(*dst, w.data(), w.size(), cudaMemcpyHostToDevice));
  }
  return TTK_OK;
}

int ttk_conv_umma_launch(const TtkConv& cv, const ConvLaunch& a, cudaStream_t st, int esz) {
  const int ci = cv.cin_p, co = cout_tile(cv, esz);
  const int k = cv.k, s = cv.stride;
  if (esz == 2) {
#define TTK_UMMA(KS_, S_, CI_, CO_, R_, ST_, RB_) \
  if (k == KS_ && s == S_ && ci == CI_ && co == CO_) return launch<KS_, S_, CI_, CO_, R_, ST_, RB_>(cv, a, st);
    // 3x3 stride 1
    TTK_UMMA(3, 1, 16, 64, 4, 3, 0)     // stem conv1 (9 -> 64, input padded to 16 channels)
    if (k == 3 && s == 1 && ci == 64 && co == 64) return launch<3, 1, 64, 64, 4, 3, 0, 32>(cv, a, st);      // stem conv2, quarter-resolution branch
    if (k == 3 && s == 1 && ci == 32 && co == 32 && a.nres == 0) return launch<3, 1, 32, 32, 8, 2, 0>(cv, a, st);   // no residual tiles to stage: taller tile
    TTK_UMMA(3, 1, 32, 32, 4, 2, 1)     // bottleneck conv2, half-resolution branch
    TTK_UMMA(3, 1, 16, 16, 8, 3, 1)     // full-resolution branch
    if (k == 3 && s == 1 && ci == 128 && co == 16) return launch<3, 1, 128, 16, 8, 2, 0, 32>(cv, a, st);     // transition1.0
    if (k == 3 && s == 1 && ci == 128 && co == 64) return launch<3, 1, 128, 64, 2, 2, 0, 32>(cv, a, st);    // eighth-resolution branch (128 -> 128 as two 64-channel output slices)
    // 3x3 stride 2 (transitions and fuse down-paths)
    if (k == 3 && s == 2 && ci == 128 && co == 32) return launch<3, 2, 128, 32, 2, 3, 0, 32>(cv, a, st);
    TTK_UMMA(3, 2, 16, 16, 4, 3, 0)
    TTK_UMMA(3, 2, 16, 32, 4, 3, 0)
    TTK_UMMA(3, 2, 16, 64, 4, 3, 0)
    TTK_UMMA(3, 2, 16, 128, 2, 3, 0)
    TTK_UMMA(3, 2, 32, 32, 4, 2, 0)
    TTK_UMMA(3, 2, 32, 64, 4, 2, 0)
    TTK_UMMA(3, 2, 32, 128, 2, 2, 0)
    if (k == 3 && s == 2 && ci == 64 && co == 64) return launch<3, 2, 64, 64, 2, 3, 0, 32>(cv, a, st);       // 64 -> 128 as two output slices
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
  // ---- TF32 (fp32 activations): same tiles, re-sized for 4-byte elements (shared memory: weights + STAGES boxes + residual tiles) ----
#define TTK_UMMA32(KS_, S_, CI_, CO_, R_, ST_, RB_, KC_) \
  if (k == KS_ && s == S_ && ci == CI_ && co == CO_) return launch<KS_, S_, CI_, CO_, R_, ST_, RB_, KC_, 4>(cv, a, st);
  // 3x3 stride 1
  TTK_UMMA32(3, 1, 16, 64, 4, 3, 0, 16)      // stem conv1
  TTK_UMMA32(3, 1, 64, 64, 4, 3, 0, 8)       // stem conv2, quarter-resolution branch: 147 KB of weights + three 25 KB boxes (no room for store staging; two boxes + staging measured slower: 3.80 vs 3.66 ms)
  if (k == 3 && s == 1 && ci == 32 && co == 32 && a.nres == 0) return launch<3, 1, 32, 32, 8, 2, 0, 16, 4>(cv, a, st);
  TTK_UMMA32(3, 1, 32, 32, 4, 3, 0, 16)      // half-resolution branch (residual read from global memory: no room for staged tiles)
  if (k == 3 && s == 1 && ci == 16 && co == 16 && a.nres == 0) return launch<3, 1, 16, 16, 8, 2, 0, 16, 4>(cv, a, st);     // taller tile (R = 4 with four stages: 0.77 vs 0.66 ms)
  TTK_UMMA32(3, 1, 16, 16, 4, 3, 1, 16)      // full-resolution branch, residual tiles staged by TMA
  if (k == 3 && s == 1 && ci == 128 && co == 16) return launch<3, 1, 128, 16, 4, 3, 0, 16, 4, 1>(cv, a, st);     // transition1.0: nine taps per MMA
  TTK_UMMA32(3, 1, 128, 32, 4, 3, 0, 8)      // eighth-resolution branch (128 -> 128 as four 32-channel output slices)
  // 3x3 stride 2
  TTK_UMMA32(3, 2, 128, 32, 2, 3, 0, 8)
  TTK_UMMA32(3, 2, 16, 16, 4, 2, 0, 16)
  TTK_UMMA32(3, 2, 16, 32, 4, 2, 0, 16)
  TTK_UMMA32(3, 2, 16, 64, 4, 2, 0, 16)
  TTK_UMMA32(3, 2, 16, 128, 2, 2, 0, 16)
  TTK_UMMA32(3, 2, 32, 32, 4, 2, 0, 16)
  TTK_UMMA32(3, 2, 32, 64, 2, 3, 0, 16)
  TTK_UMMA32(3, 2, 32, 128, 2, 3, 0, 8)
  TTK_UMMA32(3, 2, 64, 32, 2, 3, 0, 8)       // 64 -> 128 as four output slices
  // 1x1
  TTK_UMMA32(1, 1, 64, 32, 4, 3, 0, 32)
  TTK_UMMA32(1, 1, 32, 128, 2, 3, 0, 32)
  TTK_UMMA32(1, 1, 64, 128, 2, 3, 0, 32)
  TTK_UMMA32(1, 1, 32, 16, 4, 3, 0, 32)
  TTK_UMMA32(1, 1, 64, 16, 4, 3, 0, 32)
  TTK_UMMA32(1, 1, 128, 16, 4, 3, 0, 32)
  TTK_UMMA32(1, 1, 128, 32, 4, 3, 0, 32)
  TTK_UMMA32(1, 1, 128, 64, 2, 3, 0, 32)
#undef TTK_UMMA32
  return TTK_ERR_UNSUPPORTED;
}
