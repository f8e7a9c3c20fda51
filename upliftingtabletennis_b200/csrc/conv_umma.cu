// bf16 implicit-GEMM convolution on tcgen05 / TMEM fed by TMA (placeholder until the kernel lands:
// every shape reports "unsupported" and the executor runs the SIMT kernel on bf16 storage).
#include "hrnet.h"

int ttk_conv_umma_pack(TtkConv& cv, const float* w_host) {
  (void)cv;
  (void)w_host;
  return TTK_OK;
}

int ttk_conv_umma_launch(const TtkConv& cv, const ConvLaunch& a, cudaStream_t st) {
  (void)cv;
  (void)a;
  (void)st;
  return TTK_ERR_UNSUPPORTED;
}
