// Throughput probe: ex2.approx.ftz.f32 vs ex2.approx.ftz.bf16x2 (exps per clock per SM).  nvcc -arch=sm_100a -o mufu_probe mufu_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_f32(float* out, int iters) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = -0.001f * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_bf16x2(unsigned* out, int iters) {
  unsigned a[8];
  for (int i = 0; i < 8; ++i) a[i] = 0xbf00bf00u + threadIdx.x + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a[i]));
  unsigned s = 0; for (int i = 0; i < 8; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_f16x2(unsigned* out, int iters) {
  unsigned a[8];
  for (int i = 0; i < 8; ++i) a[i] = 0xb800b800u + threadIdx.x + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[i]));
  unsigned s = 0; for (int i = 0; i < 8; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* o; cudaMalloc(&o, 148 * 1024 * 4);
  int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  for (int v = 0; v < 3; ++v) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (v == 0) k_f32<<<148, 1024>>>(o, iters);
      if (v == 1) k_bf16x2<<<148, 1024>>>((unsigned*)o, iters);
      if (v == 2) k_f16x2<<<148, 1024>>>((unsigned*)o, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double instr = 1024.0 * iters * 8;           // thread-level instructions per SM
    double clks = ms * 1e-3 * clk * 1e3;
    printf("%s: %.3f ms, %.2f thread-instr/clk/SM (%s exps/clk/SM %.2f) [clock %d kHz nominal]\n", v == 0 ? "ex2.f32" : v == 1 ? "ex2.bf16x2" : "ex2.f16x2", ms,
           instr / clks, v == 0 ? "" : "2 per instr:", (v == 0 ? 1 : 2) * instr / clks, clk);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
