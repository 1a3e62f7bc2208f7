// Non-causal joint attention, head_dim 128, ragged query rows against the persistent K/V cache (attention.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace rge {

struct AttnArgs {
  const __nv_bfloat16* Q = nullptr;  // [Sq, H*128]  post-norm, post-RoPE queries of the active rows
  long ldq = 0;
  const __nv_bfloat16* K = nullptr;  // [Skv, H*128] persistent cache (post-norm, post-RoPE)
  long ldk = 0;
  const __nv_bfloat16* V = nullptr;  // [Skv, H*128]
  long ldv = 0;
  __nv_bfloat16* O = nullptr;        // [Sq, ldo], head h written at columns [h*128, h*128+128)
  long ldo = 0;
  int Sq = 0, Skv = 0, H = 0;
  float scale = 0.08838834764831845f;  // 1/sqrt(128)
  // optional scratch (attention_workspace_bytes(H)) that lets the launcher cut the ragged last query tile of every
  // head along K/V when that tile alone would cost a second wave of CTAs; not shared between concurrent launches
  void* workspace = nullptr;
  size_t workspace_bytes = 0;
};

size_t attention_workspace_bytes(int H);

// Dispatches on tuning().attn_kernel: 0 = attention.cu (128-row K/V tiles, P aliased onto S), 1 = attention64.cu
// (64-row K/V tiles, P in TMEM columns of its own: Q K^T of the next tile overlaps the softmax of the current one).
cudaError_t launch_attention(const AttnArgs& a, cudaStream_t stream);
cudaError_t launch_attention128(const AttnArgs& a, cudaStream_t stream);
cudaError_t launch_attention64(const AttnArgs& a, cudaStream_t stream);

}  // namespace rge
