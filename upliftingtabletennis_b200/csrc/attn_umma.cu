// bf16 tcgen05 flash attention for the ViTPose detector (placeholder until the kernel lands in the next commit).
#include "vit.h"

int ttk_attention_umma(const __nv_bfloat16*, __nv_bfloat16*, int, int, int, int, cudaStream_t) {
  ttk_set_error("ttk_attention_umma: the bf16 tensor-core path of the ViT detector is not built yet");
  return TTK_ERR_UNSUPPORTED;
}
