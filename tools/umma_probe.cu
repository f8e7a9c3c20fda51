// Ground-truth probe for the tcgen05 / TMA conventions used by conv_umma.cu (run once on a B200):
//   1. shared-memory matrix descriptors: K-major, no swizzle and 32/64/128-byte swizzle, with the start
//      address shifted by whole rows (the implicit-GEMM convolution reads the 3 horizontal taps of a
//      staged halo row by shifting the descriptor instead of re-loading), with and without base_offset;
//   2. instruction descriptor, tcgen05.ld lane mapping, commit/mbarrier protocol;
//   3. TMA tile-load bandwidth as a function of the box's inner extent (16 B ... 128 B).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_probe tools/umma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e = (x);                                                              \
    if (e != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);  \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}

// kind::f16, A/B bf16, D f32, both K-major
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct ProbeCfg {
  int mode;        // 0 none, 1 sw32, 2 sw64, 3 sw128
  int shift;       // rows
  int use_base_offset;
  int N;
  int kchunks16;   // number of K=16 MMA steps
};

// One CTA, 128 threads.  A: [rows_total][K] bf16 row-major in global, B: [N][K].  D: [128][N] f32.
__global__ void __launch_bounds__(128) probe_kernel(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, ProbeCfg cfg, int rows_total) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int K = cfg.kchunks16 * 16;
  const int rowbytes = K * 2;
  uint8_t* sA = smem;                          // 1024-aligned
  uint8_t* sB = smem + 48 * 1024;              // 1024-aligned
  const int tid = threadIdx.x;
  const uint32_t swz_mask = cfg.mode == 1 ? 1 : cfg.mode == 2 ? 3 : cfg.mode == 3 ? 7 : 0;
  auto place = [&](uint8_t* base, int r, int c16, int nrows) -> uint8_t* {
    if (cfg.mode == 0) return base + (size_t)c16 * nrows * 16 + (size_t)r * 16;      // [kchunk][row][16B]
    uint32_t a = smem_u32(base) + r * rowbytes + c16 * 16;
    a ^= ((a >> 7) & swz_mask) << 4;                                                 // swizzle on absolute smem address bits
    return base + (a - smem_u32(base));
  };
  for (int i = tid; i < rows_total * (K / 8); i += 128) {
    const int r = i / (K / 8), c = i % (K / 8);
    *reinterpret_cast<uint4*>(place(sA, r, c, rows_total)) = *reinterpret_cast<const uint4*>(A + (size_t)r * K + c * 8);
  }
  for (int i = tid; i < cfg.N * (K / 8); i += 128) {
    const int r = i / (K / 8), c = i % (K / 8);
    *reinterpret_cast<uint4*>(place(sB, r, c, cfg.N)) = *reinterpret_cast<const uint4*>(B + (size_t)r * K + c * 8);
  }
  if (tid == 0) mbar_init(&bar, 1);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core (async proxy)
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, cfg.N);
    const uint32_t layout = cfg.mode == 0 ? 0 : cfg.mode == 1 ? 6 : cfg.mode == 2 ? 4 : 2;
    for (int k = 0; k < cfg.kchunks16; ++k) {
      uint32_t a_addr, b_addr, lbo_a, lbo_b, sbo;
      if (cfg.mode == 0) {
        a_addr = smem_u32(sA) + (2 * k) * rows_total * 16 + cfg.shift * 16;
        b_addr = smem_u32(sB) + (2 * k) * cfg.N * 16;
        lbo_a = rows_total * 16;
        lbo_b = cfg.N * 16;
        sbo = 128;
      } else {
        a_addr = smem_u32(sA) + cfg.shift * rowbytes + k * 32;
        b_addr = smem_u32(sB) + k * 32;
        lbo_a = lbo_b = 16;
        sbo = 8 * rowbytes;
      }
      const uint32_t bo = cfg.use_base_offset ? ((a_addr >> 7) & 7) : 0;
      const uint64_t da = make_desc(a_addr, lbo_a, sbo, layout, bo);
      const uint64_t db = make_desc(b_addr, lbo_b, sbo, layout, 0);
      const uint32_t acc = k > 0 ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
          "l"(da), "l"(db), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = tid >> 5;
  for (int c0 = 0; c0 < cfg.N; c0 += 16) {
    uint32_t v[16];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) D[(size_t)tid * cfg.N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

// ---------------------------------------------------------------------------------------------------
// TMA bandwidth: persistent CTAs stream boxes of an NHWC bf16 tensor into a 4-deep smem ring
// ---------------------------------------------------------------------------------------------------
template <int RANK>
__device__ __forceinline__ void tma_load(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  if (RANK == 3)
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
  else if (RANK == 4)
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
  else
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

template <int RANK>
__global__ void __launch_bounds__(128) tma_bw_kernel(const __grid_constant__ CUtensorMap map, int box_bytes, int tiles_x, int tiles_y,
                                                     int n_img, int box_w, int box_h, int stages, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[4];
  const int stage_bytes = (box_bytes + 1023) & ~1023;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long total = (long long)tiles_x * tiles_y * n_img;
  unsigned long long acc = 0;
  if (threadIdx.x == 0) {
    long long issued = 0, consumed = 0;
    long long t = blockIdx.x;
    uint32_t phase[4] = {0, 0, 0, 0};
    while (true) {
      while (issued - consumed < stages && t < total) {
        const int s = (int)(issued % stages);
        const int tx = (int)(t % tiles_x), ty = (int)((t / tiles_x) % tiles_y), img = (int)(t / ((long long)tiles_x * tiles_y));
        mbar_expect_tx(&full[s], box_bytes);
        // coordinates: channel 0, x, y, (channel group 0), image ; -1 exercises the zero-filled halo
        if (RANK == 4) tma_load<4>(smem + (size_t)s * stage_bytes, &map, &full[s], 0, tx * (box_w - 2) - 1, ty * (box_h - 2) - 1, img, 0);
        else tma_load<5>(smem + (size_t)s * stage_bytes, &map, &full[s], 0, tx * (box_w - 2) - 1, ty * (box_h - 2) - 1, 0, img);
        ++issued;
        t += gridDim.x;
      }
      if (consumed == issued) break;
      const int s = (int)(consumed % stages);
      mbar_wait(&full[s], phase[s]);
      phase[s] ^= 1;
      acc += *reinterpret_cast<volatile unsigned long long*>(smem + (size_t)s * stage_bytes);
      ++consumed;
    }
    sink[blockIdx.x] = acc;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  // ---- part 1: descriptors ----
  const int rows_total = 144, N = 32;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  const char* mode_name[4] = {"none", "sw32", "sw64", "sw128"};
  for (int mode = 0; mode < 4; ++mode) {
    const int kch = mode == 0 ? 4 : mode == 1 ? 1 : mode == 2 ? 2 : 4;
    const int K = kch * 16;
    std::vector<float> hA(rows_total * K), hB(N * K);
    std::vector<__nv_bfloat16> bA(hA.size()), bB(hB.size());
    srand(1234 + mode);
    for (size_t i = 0; i < hA.size(); ++i) { hA[i] = bf16_round((rand() % 2001 - 1000) / 1000.f); bA[i] = __float2bfloat16_rn(hA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { hB[i] = bf16_round((rand() % 2001 - 1000) / 1000.f); bB[i] = __float2bfloat16_rn(hB[i]); }
    __nv_bfloat16 *dA, *dB;
    float* dD;
    CK(cudaMalloc(&dA, bA.size() * 2));
    CK(cudaMalloc(&dB, bB.size() * 2));
    CK(cudaMalloc(&dD, 128 * N * 4));
    CK(cudaMemcpy(dA, bA.data(), bA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, bB.data(), bB.size() * 2, cudaMemcpyHostToDevice));
    const int shifts[6] = {0, 1, 2, 3, 5, 8};
    for (int si = 0; si < 6; ++si)
      for (int ubo = 0; ubo < 2; ++ubo) {
        if (mode == 0 && ubo == 1) continue;
        ProbeCfg cfg = {mode, shifts[si], ubo, N, kch};
        CK(cudaMemset(dD, 0xff, 128 * N * 4));
        probe_kernel<<<1, 128, 96 * 1024>>>(dA, dB, dD, cfg, rows_total);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("desc mode=%s shift=%d base_offset=%d : CUDA error %s\n", mode_name[mode], shifts[si], ubo, cudaGetErrorString(e));
          return 2;
        }
        std::vector<float> hD(128 * N);
        CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)hA[(m + shifts[si]) * K + k] * hB[n * K + k];
            double d = fabs(ref - hD[m * N + n]);
            if (!(d <= 1e30)) d = 1e30;
            if (d > maxerr) maxerr = d;
          }
        printf("desc mode=%-5s shift=%d base_offset=%d : max|err| = %.3g %s\n", mode_name[mode], shifts[si], ubo, maxerr,
               maxerr < 1e-3 ? "OK" : "WRONG");
      }
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dD);
  }
  // ---- part 2: TMA bandwidth vs inner box extent ----
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 3; }
  const int H = 704, W = 1280, NI = 8;
  for (int C : {16, 32, 64}) {
    __nv_bfloat16* t;
    const size_t bytes = (size_t)NI * H * W * C * 2;
    CK(cudaMalloc(&t, bytes));
    CK(cudaMemset(t, 0, bytes));
    unsigned long long* sink;
    CK(cudaMalloc(&sink, 4096 * 8));
    for (int style = 0; style < 2; ++style) {
      // style 0: "interleaved" 5-D view (8 ch | W | H | C/8 | N), 16-byte inner extent, no swizzle
      // style 1: plain 4-D view (C | W | H | N), inner extent C*2 bytes, matching swizzle
      const int box_w = 130, box_h = 6;
      CUtensorMap map;
      CUresult r;
      if (style == 0) {
        cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(C / 8), (cuuint64_t)NI};
        cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, 16, (cuuint64_t)H * W * C * 2};
        cuuint32_t box[5] = {8, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)(C / 8), 1};
        cuuint32_t es[5] = {1, 1, 1, 1, 1};
        r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, t, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      } else {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NI};
        cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
        cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUtensorMapSwizzle sw = C == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : C == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
        r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, t, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      }
      if (r != CUDA_SUCCESS) { printf("tma C=%d style=%d: encode failed (%d)\n", C, style, (int)r); continue; }
      const int box_bytes = box_w * box_h * C * 2;
      const int stage_bytes = (box_bytes + 1023) & ~1023;
      const int stages = stage_bytes * 4 <= 200 * 1024 ? 4 : 2;
      const int smem_bytes = stages * stage_bytes;
      const int tiles_x = W / (box_w - 2), tiles_y = H / (box_h - 2);
      for (int ctas_per_sm = 1; ctas_per_sm <= 2; ++ctas_per_sm) {
        if (smem_bytes * ctas_per_sm > 220 * 1024) continue;
        const int grid = prop.multiProcessorCount * ctas_per_sm;
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
          CK(cudaEventRecord(e0));
          if (style == 0) {
            CK(cudaFuncSetAttribute(tma_bw_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
            tma_bw_kernel<5><<<grid, 128, smem_bytes>>>(map, box_bytes, tiles_x, tiles_y, NI, box_w, box_h, stages, sink);
          } else {
            CK(cudaFuncSetAttribute(tma_bw_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
            tma_bw_kernel<4><<<grid, 128, smem_bytes>>>(map, box_bytes, tiles_x, tiles_y, NI, box_w, box_h, stages, sink);
          }
          CK(cudaEventRecord(e1));
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("tma C=%d style=%d: CUDA error %s\n", C, style, cudaGetErrorString(e)); return 4; }
          float ms;
          CK(cudaEventElapsedTime(&ms, e0, e1));
          if (ms < best) best = ms;
        }
        const double moved = (double)tiles_x * tiles_y * NI * box_bytes;
        printf("tma C=%-3d %s inner=%3d B box=%dx%d ctas/SM=%d : %.3f ms, smem fill %.0f GB/s (tensor %.0f MB => dram-side %.0f GB/s)\n", C,
               style == 0 ? "5D-interleaved/noswz" : "4D-swizzled        ", style == 0 ? 16 : C * 2, box_w, box_h, ctas_per_sm, best,
               moved / best / 1e6, bytes / 1e6, bytes / best / 1e6);
      }
    }
    cudaFree(t);
    cudaFree(sink);
  }
  return 0;
}
