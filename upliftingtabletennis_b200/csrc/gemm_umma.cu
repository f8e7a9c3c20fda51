// bf16 tcgen05 GEMM for the ViTPose detector (placeholder until the kernel lands in the next commit).
#include "vit.h"

int ttk_gemm_umma(const GemmArgs&, cudaStream_t) {
  ttk_set_error("ttk_gemm_umma: the bf16 tensor-core path of the ViT detector is not built yet");
  return TTK_ERR_UNSUPPORTED;
}
