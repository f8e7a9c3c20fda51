// Uplifting transformer internals shared by uplift.cu (fp32 path) and uplift_tc.cu (bf16 tcgen05 path).
#pragma once
#include <string>
#include <vector>

#include "ttk_internal.h"

struct LayerW {
  const float *ln1w, *ln1b, *qkvw, *qkvb, *projw, *invf, *fc1w, *fc1b, *fc2w, *fc2b, *ln2w, *ln2b;
};
struct HeadW {
  const float *w1, *b1, *w2, *b2, *w3, *b3;
};


struct UpliftParam {
  std::string name;
  int numel;
  float* dev = nullptr;
  bool set = false;
};

struct ttk_uplift {
  int dim, heads, depth, skip;
  std::vector<UpliftParam> params;
  LayerW* layers_dev = nullptr;    // [4 pos + (depth-4) temporal + 4 second]
  bool layers_ready = false;
  int launches = 0;
  // bf16 tensor-core path: every layer's [qkv(384) | proj(128) | fc1(128) | fc2(128)] x 128 weight rows, K-major
  __nv_bfloat16* wmat_dev = nullptr;
  bool wmat_ready = false;
  int find(const std::string& n) const {
    for (size_t i = 0; i < params.size(); ++i)
      if (params[i].name == n) return (int)i;
    return -1;
  }
  const float* dev(const std::string& n) const { return params[find(n)].dev; }
};

struct UpliftIO {
  const float *ball, *table, *mask, *times;
  int batch, T;
  float *rot_out, *pos_out;
  float *X, *table_emb, *second_emb;     // workspace slices
  void* attn_rows;                       // bf16 [batch*T][128]: attention output of the ball tokens in the last table-token layer
};

// uplift_tc.cu
int ttk_uplift_tc_prepare(ttk_uplift* h);
int ttk_uplift_tc_stage(ttk_uplift* h, int mode, const UpliftIO& io, cudaStream_t st);
