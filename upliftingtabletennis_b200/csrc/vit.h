// ViTPose detector internals shared by vit.cu (plan, fp32 parity kernels), gemm_umma.cu (bf16 tcgen05 GEMM) and
// attn_umma.cu (bf16 tcgen05 flash attention).
#pragma once
#include <string>
#include <vector>

#include "ttk_internal.h"

enum { VIT_ACT_NONE = 0, VIT_ACT_GELU = 1, VIT_ACT_RELU = 2 };

// C[m][n] = act(sum_k A[m][k] W[n][k] + bias[n]) (+ R[m][n]); rows of C optionally scattered to the 2x up-sampled grid of a
// transposed convolution: m = (img, y, x) on an up_h x up_w grid -> row (img, 2y + py, 2x + px) of a 2up_h x 2up_w grid.
struct GemmArgs {
  const void* A;          // [M][K]  (f32 or bf16, K contiguous)
  const void* W;          // [N][K]
  const float* bias;      // [N] or null
  const float* R;         // [M][N] float32 residual or null (may alias C when C is float32)
  int r_mod = 0;          // > 0: the residual has r_mod rows and row m reads row m % r_mod (position table repeated per image)
  void* C;                // [M][N] f32 or bf16
  int M, N, K;
  int act;
  int c_bf16;             // output type
  int up_h, up_w, py, px; // up_w == 0: no scatter
  // 3xTF32 mode (tcgen05 path, fp32-level results): A, W (and C when c_split) are pairs of float32 planes -- plane 0 = tf32(x), plane 1 =
  // tf32(x - plane 0), *_plane bytes apart -- and every product is three kind::tf32 MMAs (hi*hi + lo*hi + hi*lo)
  int x3 = 0;
  int c_split = 0;        // x3 only: write C as such a pair (it feeds another x3 GEMM / the attention); 0: plain float32
  size_t a_plane = 0, w_plane = 0, c_plane = 0;
  // x3 qkv GEMM only: columns >= vt_col0 (the V third, one head per 32 columns) are written TRANSPOSED instead, as the attention's B
  // operand: vt[plane][row / vt_tokens][(col - vt_col0)][row % vt_tokens] with rows vt_tok_pad floats long, planes vt_plane bytes apart
  float* vt = nullptr;
  size_t vt_plane = 0;
  int vt_col0 = 0, vt_tokens = 0, vt_tok_pad = 0;
  int implicit_c = 0;     // > 0 (tcgen05 path only): A is the NHWC feature map [img][up_h][up_w][implicit_c] itself and the 2x2 taps of the
                          // transposed convolution's output parity (py, px) are gathered by TMA (K = 4 taps x implicit_c), no materialised gather
};

struct VitParam {
  std::string name;
  std::vector<int> shape;
  size_t numel;
  std::vector<float> host;
  bool set = false;
};

struct ttk_vit {
  int in_ch, out_ch, height, width, hp, wp, tokens;
  std::vector<VitParam> params;
  bool ready = false;
  int device = -1;          // device of the packed weights (ttk_bind_device)
  int launches = 0;
  int subbatch = 16;
  // prepared device weights (float32 and bf16 copies of every GEMM operand)
  float* f32_pool = nullptr;
  __nv_bfloat16* bf16_pool = nullptr;
  float* x3_pool = nullptr;      // [hi pool | lo pool]: tf32 split of f32_pool for the 3xTF32 path, planes pool_bytes apart
  size_t pool_bytes = 0;
  struct Lin {
    size_t w_off, b_off;    // offsets (elements) into the pools; bias always float32
    int n, k;
  };
  Lin patch;
  size_t pos_off;           // [tokens][384] float32: pos_embed[1:] + pos_embed[:1]
  struct Block {
    size_t ln1w, ln1b, ln2w, ln2b;
    Lin qkv, proj, fc1, fc2;
  };
  std::vector<Block> blocks;
  size_t lnfw, lnfb;
  Lin deconv[2][4];         // [layer][parity py*2+px], BN folded, K = 4 taps x Cin
  size_t final_w, final_b;  // [out_ch][256], [out_ch]
  int find(const std::string& n) const {
    for (size_t i = 0; i < params.size(); ++i)
      if (params[i].name == n) return (int)i;
    return -1;
  }
};

// gemm_umma.cu: bf16 (or, x3, split float32) operands through TMA, tcgen05.mma, fp32 accumulation in TMEM.  Returns TTK_ERR_UNSUPPORTED for shapes
// it has no kernel for (the caller then reports the error; there is no silent fallback).
int ttk_gemm_umma(const GemmArgs& g, cudaStream_t st);
// attn_umma.cu: softmax(q k^T / sqrt(32)) v per (image, head) on tensor cores.  qkv [T][3*dim] bf16, out [T][dim] bf16.
// vt_scratch: ttk_attention_umma_scratch_bytes(...) bytes for the transposed V operand.
size_t ttk_attention_umma_scratch_bytes(int images, int tokens, int heads, int head_dim);
int ttk_attention_umma(const __nv_bfloat16* qkv, __nv_bfloat16* out, void* vt_scratch, int images, int tokens, int heads, int head_dim,
                       cudaStream_t st);
// attn3_umma.cu: the same attention with fp32-level results (3xTF32).  qkv / out are split float32 plane pairs ([2][T][3*dim],
// [2][T][dim]; planes qkv_plane / out_plane bytes apart); vt_scratch: ttk_attention3_scratch_bytes(...) bytes.
size_t ttk_attention3_scratch_bytes(int images, int tokens, int heads, int head_dim);
// v_transposed: vt_scratch already holds V^T ([2][images][heads*32][tok_pad], tok_pad = tokens rounded up to 8) -- the qkv GEMM wrote it
int ttk_attention3(const float* qkv, size_t qkv_plane, float* out, size_t out_plane, void* vt_scratch, int images, int tokens, int heads,
                   int head_dim, cudaStream_t st, int v_transposed = 0);
