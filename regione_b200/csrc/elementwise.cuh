// HBM-bound kernels of the RegionE hot path (everything that is not a GEMM or attention).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rge {

// out[m,:] = bf16(bf16(LN(x[m,:])) * bf16(1 + scale) + shift)   LN: no affine, eps 1e-6, fp32 statistics.
cudaError_t launch_ln_modulate(const __nv_bfloat16* x, long ldx, const __nv_bfloat16* scale,
                               const __nv_bfloat16* shift, __nv_bfloat16* out, long ldo, int M, int D,
                               cudaStream_t s);

// out[m,:] = bf16(x[m,:] * rsqrt(mean(x^2) + eps)) * w     (diffusers RMSNorm with weight)
cudaError_t launch_rmsnorm(const __nv_bfloat16* x, long ldx, const __nv_bfloat16* w, __nv_bfloat16* out, long ldo,
                           int M, int D, float eps, cudaStream_t s);
// Norm-rescaled CFG: comb = neg + scale (pos - neg); out = comb * (|pos| / |comb|) per token.
cudaError_t launch_cfg_rescale(const __nv_bfloat16* pos, const __nv_bfloat16* neg, float scale, __nv_bfloat16* out,
                               int M, int Cch, cudaStream_t s);

// Step1X CFG: out[m] = bf16(|pos[m,:] - neg[m,:]|);  out = neg + scale (pos - neg) [/ denom[m]] in bf16 steps.
cudaError_t launch_row_diff_norm(const __nv_bfloat16* pos, const __nv_bfloat16* neg, __nv_bfloat16* out, int M,
                                 int Cch, cudaStream_t s);
cudaError_t launch_cfg_combine(const __nv_bfloat16* pos, const __nv_bfloat16* neg, float scale,
                               const __nv_bfloat16* denom, __nv_bfloat16* out, int M, int Cch, cudaStream_t s);

// Batched GEMV: for every job j, out_j[n] = act_out(W_j[n,:] . act_in(x_j) + b_j[n]); one warp per output row.
struct GemvJob {
  const __nv_bfloat16* W;  // [N, K]
  const __nv_bfloat16* b;  // [N] or null
  const __nv_bfloat16* x;  // [K]
  __nv_bfloat16* out;      // [N]
  int N, K;
  int silu_in, silu_out;
};
cudaError_t launch_gemv_batch(const GemvJob* jobs_dev, int n_jobs, int max_n, cudaStream_t s);

// Sinusoidal timestep projection (256 channels, cos first), rounded to bf16.
cudaError_t launch_timestep_proj(float t, __nv_bfloat16* out256, cudaStream_t s);
// out = bf16(bf16(a + b) + c)
cudaError_t launch_add3(const __nv_bfloat16* a, const __nv_bfloat16* b, const __nv_bfloat16* c, __nv_bfloat16* out,
                        int n, cudaStream_t s);

// Rotary table: ids [S,3] fp32 -> (cos, sin) per rotary pair, axes (16,56,56), theta 10000, fp64 angles.
// ld == 0: cs is [S][64] row-major; ld > 0: pair-major [64][ld] (the GEMM epilogue's coalesced layout).
cudaError_t launch_rope_table(const float* ids, float2* cs, int S, long ld, cudaStream_t s);
// [S][64] row-major table -> pair-major [64][ld]
cudaError_t launch_rope_transpose(const float2* src, float2* dst, int S, long ld, cudaStream_t s);

// sel_all = [0..T-1, T + sel_img[i]]; sel_img == null means identity over n_img.
cudaError_t launch_build_selection(const int* sel_img, int n_img, int T, int* sel_img_out, int* sel_all_out,
                                   cudaStream_t s);

// Row gather / scatter on bf16 rows of `width` elements (width % 8 == 0).
cudaError_t launch_gather_rows(const __nv_bfloat16* src, long lds, const int* ids, int n, int width,
                               __nv_bfloat16* dst, long ldd, cudaStream_t s);
cudaError_t launch_scatter_rows(const __nv_bfloat16* src, long lds, const int* ids, int n, int width,
                                __nv_bfloat16* dst, long ldd, cudaStream_t s);

// Euler update x' = bf16(x + bf16(dt_row * v')), v' = vscale_on ? bf16(v * vscale) : v.
// dt_row = mask ? (mask[m] ? dt : dt_direct) : dt.
cudaError_t launch_euler(const __nv_bfloat16* x, const __nv_bfloat16* v, __nv_bfloat16* out, int M, int Cch,
                         float dt, float dt_direct, const uint8_t* mask, int vscale_on, float vscale,
                         cudaStream_t s);

// Adaptive region partition, part 1: raw mask[m] = cos_sim(x + bf16(dt_final*v), cond) <= thr. sim_out optional.
cudaError_t launch_arp_similarity(const __nv_bfloat16* x, const __nv_bfloat16* v, const __nv_bfloat16* cond,
                                  float dt_final, float thr, uint8_t* mask, float* sim_out, int L, int Cch,
                                  cudaStream_t s);
// part 2: optional erosion(3x3 cross)+dilation(5x5 square) with zero padding, then ordered compaction.
// counts[0] = n_edited, counts[1] = n_unedited.
cudaError_t launch_morph_compact(const uint8_t* mask_in, uint8_t* mask_out, int gh, int gw, int erosion_dilation,
                                 int* edited, int* unedited, int* counts, cudaStream_t s);

// Latent pack / unpack: planar [B, C, H, W] <-> packed [B, (H/2)(W/2), 4C] (packed channel = 4c + 2dy + dx).
cudaError_t launch_pack_latents(const __nv_bfloat16* src, __nv_bfloat16* dst, int B, int C, int H, int W, bool unpack,
                                cudaStream_t s);

}  // namespace rge
