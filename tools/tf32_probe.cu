// Ground-truth probe for the TF32 tensor-core path (run once on a B200):
//   1. what a TMA tile load does to fp32 data when the tensor map's data type is FLOAT32 / TFLOAT32 (bits written to smem);
//   2. how tcgen05.mma.kind::tf32 treats the low 13 mantissa bits of its fp32 operands (truncate or round);
//   3. tensor-pipe time per MMA: kind::tf32 (M128 x N x K8) and kind::f16 (M128 x N x K16) as a function of N.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tf32_probe tools/tf32_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
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
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}
// D f32 (bit 4), A/B format at bits 7 / 10: 1 = bf16 (kind::f16), 2 = tf32 (kind::tf32); K-major
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int fmt) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <int TF32>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  if (TF32)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ---- part 1: TMA load of a [64][32] fp32 tile (128-byte rows, SWIZZLE_128B), raw smem copied out ----
__global__ void __launch_bounds__(128) tma_kernel(const __grid_constant__ CUtensorMap map, uint32_t* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, 64 * 128);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem)),
                 "l"(&map), "r"(smem_u32(&bar)), "r"(0), "r"(0)
                 : "memory");
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < 64 * 32; i += 128) {
    const int r = i / 32, c = i % 32;
    uint32_t a = smem_u32(smem) + r * 128 + c * 4;
    a ^= ((a >> 7) & 7) << 4;
    out[i] = *reinterpret_cast<uint32_t*>(smem + (a - smem_u32(smem)));
  }
}

// ---- part 2/3: one CTA, A [128][32] fp32 (or [128][64] bf16) K-major SWIZZLE_128B rows, B [N][..] likewise ----
template <int TF32>
__global__ void __launch_bounds__(128) mma_kernel(const uint32_t* A, const uint32_t* B, float* D, int N, int reps, long long* clk) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 16 * 1024;
  const int tid = threadIdx.x;
  auto place = [&](uint8_t* base, int r, int c16) -> uint8_t* {
    uint32_t a = smem_u32(base) + r * 128 + c16 * 16;
    a ^= ((a >> 7) & 7) << 4;
    return base + (a - smem_u32(base));
  };
  for (int i = tid; i < 128 * 8; i += 128) *reinterpret_cast<uint4*>(place(sA, i / 8, i % 8)) = reinterpret_cast<const uint4*>(A)[i];
  for (int i = tid; i < N * 8; i += 128) *reinterpret_cast<uint4*>(place(sB, i / 8, i % 8)) = reinterpret_cast<const uint4*>(B)[i];
  if (tid == 0) mbar_init(&bar, 1);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, N, TF32 ? 2 : 1);
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep)
      for (int k = 0; k < 4; ++k)      // 4 steps of 32 bytes = one 128-byte row: K = 32 tf32 or 64 bf16
        umma<TF32>(tmem, make_desc(smem_u32(sA) + k * 32, 1024, 2), make_desc(smem_u32(sB) + k * 32, 1024, 2), idesc, (rep | k) ? 1u : 0u);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    mbar_wait(&bar, 0);
    clk[0] = clock64() - t0;
  }
  __syncthreads();
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = tid >> 5;
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

// ---- part 4: like the convolution's issue loop: every MMA reads a DIFFERENT A slab (start address moved by whole pixel rows), rows of
// ROWB bytes (32 / 64 / 128-byte swizzle), B = N weight rows of the same width ----
template <int TF32>
__global__ void __launch_bounds__(128) mma_shift_kernel(int N, int rowb, int reps, int distinct, long long* clk) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint8_t* sA = smem;                    // 1400 pixel rows
  uint8_t* sB = smem + 180 * 1024;       // N rows
  const int tid = threadIdx.x;
  for (int i = tid; i < (180 * 1024 + 32 * 1024) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u);
  if (tid == 0) mbar_init(&bar, 1);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, N, TF32 ? 2 : 1);
    const uint32_t layout = rowb == 32 ? 6u : rowb == 64 ? 4u : 2u;
    const int ksteps = rowb / 32;
    const long long t0 = clock64();
    int n = 0;
    for (int rep = 0; rep < reps; ++rep)
      for (int j = 0; j < 16; ++j)
        for (int k = 0; k < ksteps; ++k, ++n) {
          const uint32_t arow = smem_u32(sA) + (distinct ? (j * 65 + (rep & 3)) * rowb : 0) + k * 32;
          umma<TF32>(tmem, make_desc(arow, 8 * rowb, layout), make_desc(smem_u32(sB) + k * 32, 8 * rowb, layout), idesc, n ? 1u : 0u);
        }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    mbar_wait(&bar, 0);
    clk[0] = clock64() - t0;
    clk[1] = n;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static float as_f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static uint32_t as_u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float rna_tf32(float f) { return as_f((as_u(f) + 0x1000u) & ~0x1fffu); }
static float trunc_tf32(float f) { return as_f(as_u(f) & ~0x1fffu); }

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sm_%d%d, %d SMs, clock %d kHz\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount, prop.clockRate);
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 3; }
  // ---- part 1 ----
  {
    std::vector<float> h(64 * 32);
    srand(7);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (rand() % 200001 - 100000) / 31337.f;
    h[0] = as_f(0x3f800000u + 0x0fffu);   // just below half a tf32 ulp above 1
    h[1] = as_f(0x3f800000u + 0x1000u);   // exactly half
    h[2] = as_f(0x3f800000u + 0x1001u);   // just above
    h[3] = as_f(0x3f802000u + 0x1000u);   // half, odd tf32 mantissa
    h[4] = as_f(0x00000fffu);             // denormal
    float* d;
    uint32_t* o;
    CK(cudaMalloc(&d, h.size() * 4));
    CK(cudaMalloc(&o, h.size() * 4));
    CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024));
    const CUtensorMapDataType types[3] = {CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32_FTZ};
    const char* names[3] = {"FLOAT32", "TFLOAT32", "TFLOAT32_FTZ"};
    for (int t = 0; t < 3; ++t) {
      CUtensorMap map;
      cuuint64_t dims[2] = {32, 64};
      cuuint64_t strides[1] = {128};
      cuuint32_t box[2] = {32, 64};
      cuuint32_t es[2] = {1, 1};
      CUresult r = encode(&map, types[t], 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("tma %s: encode failed (%d)\n", names[t], (int)r); continue; }
      CK(cudaMemset(o, 0xff, h.size() * 4));
      tma_kernel<<<1, 128, 16 * 1024>>>(map, o);
      CK(cudaDeviceSynchronize());
      std::vector<uint32_t> got(h.size());
      CK(cudaMemcpy(got.data(), o, got.size() * 4, cudaMemcpyDeviceToHost));
      int same = 0, rna = 0, trunc = 0, lowzero = 0;
      for (size_t i = 0; i < h.size(); ++i) {
        same += got[i] == as_u(h[i]);
        rna += got[i] == as_u(rna_tf32(h[i]));
        trunc += got[i] == as_u(trunc_tf32(h[i]));
        lowzero += (got[i] & 0x1fffu) == 0;
      }
      printf("tma %-13s: identical %d, == rna(tf32) %d, == truncated %d, low 13 bits zero %d of %zu; first five %08x %08x %08x %08x %08x\n", names[t], same,
             rna, trunc, lowzero, h.size(), got[0], got[1], got[2], got[3], got[4]);
    }
    cudaFree(d);
    cudaFree(o);
  }
  // ---- part 2: operand rounding of kind::tf32 ----
  {
    const int N = 16;
    std::vector<float> hA(128 * 32, 0.f), hB(N * 32, 0.f);
    srand(11);
    for (auto& v : hA) v = (rand() % 200001 - 100000) / 31337.f;
    for (auto& v : hB) v = (rand() % 200001 - 100000) / 31337.f;
    uint32_t *dA, *dB;
    float* dD;
    long long* dclk;
    CK(cudaMalloc(&dA, hA.size() * 4));
    CK(cudaMalloc(&dB, hB.size() * 4));
    CK(cudaMalloc(&dD, 128 * 256 * 4));
    CK(cudaMalloc(&dclk, 8));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(mma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    mma_kernel<1><<<1, 128, 64 * 1024>>>(dA, dB, dD, N, 1, dclk);
    CK(cudaDeviceSynchronize());
    std::vector<float> hD(128 * N);
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    double e_full = 0, e_rna = 0, e_trunc = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        double full = 0, rna = 0, tr = 0;
        for (int k = 0; k < 32; ++k) {
          full += (double)hA[m * 32 + k] * hB[n * 32 + k];
          rna += (double)rna_tf32(hA[m * 32 + k]) * rna_tf32(hB[n * 32 + k]);
          tr += (double)trunc_tf32(hA[m * 32 + k]) * trunc_tf32(hB[n * 32 + k]);
        }
        e_full = fmax(e_full, fabs(full - hD[m * N + n]));
        e_rna = fmax(e_rna, fabs(rna - hD[m * N + n]));
        e_trunc = fmax(e_trunc, fabs(tr - hD[m * N + n]));
      }
    printf("mma kind::tf32 K=32: max|D - fp32 product| %.3g, max|D - rna operands| %.3g, max|D - truncated operands| %.3g\n", e_full, e_rna, e_trunc);
    // ---- part 3: tensor-pipe time per MMA ----
    const int reps = 2000;
    for (int tf = 1; tf >= 0; --tf)
      for (int n : {16, 32, 48, 64, 96, 128, 192, 256}) {
        if (tf) mma_kernel<1><<<1, 128, 64 * 1024>>>(dA, dB, dD, n, reps, dclk);
        else mma_kernel<0><<<1, 128, 64 * 1024>>>(dA, dB, dD, n, reps, dclk);
        CK(cudaDeviceSynchronize());
        long long c;
        CK(cudaMemcpy(&c, dclk, 8, cudaMemcpyDeviceToHost));
        printf("mma %s M128 N=%-3d K=%d: %.1f clk per MMA (%d MMAs back to back)\n", tf ? "tf32" : "bf16", n, tf ? 8 : 16, (double)c / (reps * 4), reps * 4);
      }
    // ---- part 4 ----
    CK(cudaFuncSetAttribute(mma_shift_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 214 * 1024));
    CK(cudaFuncSetAttribute(mma_shift_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 214 * 1024));
    long long* dclk2;
    CK(cudaMalloc(&dclk2, 16));
    for (int tf = 1; tf >= 0; --tf)
      for (int rowb : {32, 64, 128})
        for (int distinct = 0; distinct < 2; ++distinct)
          for (int n : {16, 48, 96, 144, 192, 256}) {
            if (tf) mma_shift_kernel<1><<<1, 128, 214 * 1024>>>(n, rowb, 200, distinct, dclk2);
            else mma_shift_kernel<0><<<1, 128, 214 * 1024>>>(n, rowb, 200, distinct, dclk2);
            CK(cudaDeviceSynchronize());
            long long c[2];
            CK(cudaMemcpy(c, dclk2, 16, cudaMemcpyDeviceToHost));
            printf("issue loop %s rows of %3d B, %s A slabs, N=%-3d: %.1f clk per MMA\n", tf ? "tf32" : "bf16", rowb, distinct ? "distinct" : "one    ", n, (double)c[0] / c[1]);
          }
  }
  return 0;
}
