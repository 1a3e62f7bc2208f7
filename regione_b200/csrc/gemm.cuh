// Host-side descriptor of one bf16 GEMM  out = epilogue(A[M,K] @ W[N,K]^T)  on the tcgen05 path (gemm.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace rge {

enum GemmEpilogue {
  EPI_STORE = 0,      // out = bf16(acc + bias)
  EPI_GELU = 1,       // out = bf16(gelu_tanh(bf16(acc + bias)))
  EPI_GATE_RES = 2,   // out = bf16(res + bf16(gate[n] * bf16(acc + bias)))
  EPI_NORM_ROPE = 3,  // per 128-wide head: bias, RMSNorm(eps 1e-6, weight), rotary embedding, store
};

struct GemmArgs {
  // operands
  const __nv_bfloat16* A = nullptr;  // [M, K] row-major, leading dimension lda (elements)
  long lda = 0;
  const __nv_bfloat16* W = nullptr;  // [N, K] row-major (nn.Linear weight), leading dimension ldw
  long ldw = 0;
  int M = 0, N = 0, K = 0;
  const __nv_bfloat16* bias = nullptr;  // [N] or null
  int epilogue = EPI_STORE;
  // output: element (m, n) goes to out[(row_map ? row_map[m] : m) + row_off][col_off + n]
  __nv_bfloat16* out = nullptr;
  long ldo = 0;
  const int* row_map = nullptr;
  int row_off = 0;
  int col_off = 0;
  // EPI_GATE_RES
  const __nv_bfloat16* gate = nullptr;  // [N]
  const __nv_bfloat16* res = nullptr;   // [M, ldr] rows indexed by m
  long ldr = 0;
  // EPI_NORM_ROPE
  const __nv_bfloat16* norm_w = nullptr;  // [128]
  const float2* rope_cs = nullptr;        // [S, 64] (cos, sin) per rotary pair
  const int* rope_map = nullptr;          // rope row = (rope_map ? rope_map[m] : m) + rope_off
  int rope_off = 0;
  long rope_ld = 0;                       // 0: rope_cs is [S][64]; > 0: pair-major [64][rope_ld] (coalesced per warp)
  int flags = 0;                          // bit 0: EPI_STORE stores through fp16 (fused_kernels.py:80)
};

// Launches on `stream`; returns cudaSuccess or the launch/encode error. `num_sms` bounds the persistent grid.
cudaError_t launch_gemm(const GemmArgs& a, int num_sms, cudaStream_t stream);

// Up to 6 independent GEMMs (1-CTA tiles) as one persistent launch; members with M <= 0 are skipped.
cudaError_t launch_gemm_group(const GemmArgs* args, int n, int num_sms, cudaStream_t stream);

bool gemm_args_valid(const GemmArgs& a);   // strides, alignment and epilogue operands acceptable to every kernel
size_t streamk_workspace_bytes(int num_sms);
size_t streamk_flag_bytes(int num_sms);

// Resolves cuTensorMapEncodeTiled through the runtime (no link-time libcuda dependency).
void* get_tensor_map_encoder();

}  // namespace rge
