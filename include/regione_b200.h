/* regione_b200 — C ABI of the B200-native RegionE hot path.
 *
 * Loaded with ctypes.CDLL by regione_b200/_lib.py (the host side of the reference is Python; see INTEGRATION.md for
 * the stub a RegionE maintainer would add). Conventions, mirroring SURVEY.md §8(b):
 *   - every entry point returns 0 (RGE_OK) or a negative error code and never throws; rge_last_error() gives text;
 *   - all pointers are raw device pointers unless the name says host; the caller owns every activation, latent and
 *     weight buffer (weights are BORROWED: keep the tensors alive); the library owns workspaces, TMA descriptors
 *     and the Region-Instruction KV cache;
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it unless documented otherwise;
 *   - one handle per (process, device); not thread-safe — the reference keeps a module-global singleton too
 *     (RegionE/FluxKontext/inplace.py:51).
 * Tensors are bf16 row-major unless stated. Reference citations are relative to /root/reference.
 */
#ifndef REGIONE_B200_H
#define REGIONE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RGE_ABI_VERSION 5

enum rge_status {
  RGE_OK = 0,
  RGE_ERR_INVALID = -1,     /* bad argument */
  RGE_ERR_CUDA = -2,        /* CUDA runtime / driver error; text in rge_last_error() */
  RGE_ERR_STATE = -3,       /* call order violated (weights missing, image not begun ...) */
  RGE_ERR_UNSUPPORTED = -4  /* shape outside the kernels' envelope (head_dim != 128 ...) */
};

int rge_abi_version(void);
const char* rge_last_error(void);
/* Run-time tuning knob of the kernels (same names as the RGE_* environment variables they start from, lower case
 * without the prefix: "attn_kernel", "attn_poly", "attn_split", "gemm_bn", "gemm2_bn", "2cta_min_m", "raster", "wide_store", "trim_last", "nvtx"). For
 * benchmarks that sweep variants inside one process; results never depend on them beyond the stated tolerances. */
int rge_set_option(const char* name, int32_t value);

/* ------------------------------------------------------------------------------------------------------------
 * Kernel-level entry points (each replaces one reference op; also what the parity tests call)
 * ---------------------------------------------------------------------------------------------------------- */

enum rge_epilogue { RGE_EPI_STORE = 0, RGE_EPI_GELU = 1, RGE_EPI_GATE_RES = 2, RGE_EPI_NORM_ROPE = 3 };

/* out = epilogue(A[M,K] @ W[N,K]^T + bias); element (m,n) lands at out[(row_map?row_map[m]:m)+row_off][col_off+n].
 * Replaces nn.Linear on the path (inplace.py:715-725, 768-770, 816-820) and, with row_map, the Triton scatter-GEMM
 * `_partially_linear(inputs, weight, bias, index, outputs)` (fused_kernels.py:81-101; index == row_map).
 * RGE_EPI_NORM_ROPE additionally fuses attn.norm_q/norm_k + apply_rotary_emb (inplace.py:760-763, 792-794). */
typedef struct rge_gemm_desc {
  const void* A; int64_t lda;
  const void* W; int64_t ldw;
  const void* bias;
  int32_t M, N, K;
  int32_t epilogue;
  void* out; int64_t ldo;
  const int32_t* row_map; int32_t row_off; int32_t col_off;
  const void* gate; const void* res; int64_t ldr;            /* RGE_EPI_GATE_RES: out = res + gate[n]*(..) */
  const void* norm_w; const float* rope_cs;                   /* RGE_EPI_NORM_ROPE: RMSNorm weight[128], table */
  const int32_t* rope_map; int32_t rope_off;                  /*   rope row = (rope_map?rope_map[m]:m)+rope_off */
  int64_t rope_ld;                                            /*   0: rope_cs is [S][64][2]; > 0: pair-major
                                                                   [64][rope_ld][2] (coalesced per epilogue warp) */
  int32_t flags;                                              /* RGE_GEMM_* bits */
} rge_gemm_desc;
/* RGE_EPI_STORE only: round fp32 -> fp16 -> bf16 like the reference's Triton kernel does before its store
 * (`accumulator.to(tl.float16)`, fused_kernels.py:80) instead of fp32 -> bf16. For differential tests
 * against that kernel; the engine stores fp32 -> bf16 directly. bias, gate, norm_w, out and res must be 16-byte
 * aligned. */
#define RGE_GEMM_FP16_ROUNDTRIP 1
int rge_op_gemm(const rge_gemm_desc* d, void* stream);
/* n (1..6) independent GEMMs as ONE persistent launch: the q / k / v (/ MLP-up) projections of a block, or the image-
 * and text-stream halves of one stage of a double block. Results are identical to n rge_op_gemm calls; members with
 * M <= 0 are skipped. Members may read the same A and may write disjoint regions of the same buffers. */
int rge_op_gemm_group(const rge_gemm_desc* descs, int32_t n, void* stream);

/* O[Sq, H*128] = softmax(Q K^T * scale) V per head, non-causal, head_dim 128; replaces flash_attn_func
 * (inplace.py:796-801). K/V are the persistent cache [Skv, H*128]. */
typedef struct rge_attn_desc {
  const void* Q; int64_t ldq;
  const void* K; int64_t ldk;
  const void* V; int64_t ldv;
  void* O; int64_t ldo;
  int32_t Sq, Skv, H;
  float scale;
  /* Optional device scratch of rge_attention_workspace_bytes(H) bytes (may be NULL / 0). With it, when the full
   * 256-row query tiles of all heads fit into one wave of CTAs and the ragged last tile alone would cost a second one
   * (REGION steps: 512 + ~1064 rows x 24 heads), that tile is cut along K/V and merged by a small second kernel. Must
   * not be shared between attention launches that may run concurrently. */
  void* workspace; int64_t workspace_bytes;
} rge_attn_desc;
int rge_op_attention(const rge_attn_desc* d, void* stream);
int64_t rge_attention_workspace_bytes(int32_t H);

/* out = LayerNorm(x, eps 1e-6, no affine) * (1 + scale) + shift  (diffusers AdaLayerNormZero*). */
int rge_op_ln_modulate(const void* x, int64_t ldx, const void* scale, const void* shift, void* out, int64_t ldo,
                       int32_t M, int32_t D, void* stream);

/* out = RMSNorm(x) * weight over the last dimension (diffusers RMSNorm; Qwen txt_norm, QwenImageEdit/inplace.py:518). */
int rge_op_rmsnorm(const void* x, int64_t ldx, const void* weight, void* out, int64_t ldo, int32_t M, int32_t D,
                   float eps, void* stream);

/* Norm-rescaled classifier-free guidance (QwenImageEdit/inplace.py:386-405): comb = neg + scale * (pos - neg);
 * out = comb * (||pos|| / ||comb||) per token; pos, neg, out: [M, channels]. */
int rge_cfg_rescale(const void* pos, const void* neg, float scale, void* out, int32_t M, int32_t channels,
                    void* stream);

/* Step1X classifier-free guidance (Step1XEdit/inplace.py:388-400): norm_out[m] = ||pos[m,:] - neg[m,:]|| (bf16 [M]);
 * out = neg + scale * (pos - neg) / denom[m]   (denom bf16 [M] = the pipeline's process_diff_norm(diff_norm); NULL
 * = plain CFG for t <= timesteps_truncate). */
int rge_cfg_diff_norm(const void* pos, const void* neg, void* norm_out, int32_t M, int32_t channels, void* stream);
int rge_cfg_combine(const void* pos, const void* neg, float scale, const void* denom, void* out, int32_t M,
                    int32_t channels, void* stream);

/* Rotary table of FluxPosEmbed(theta 10000, axes (16,56,56)): ids fp32 [S,3] -> cs fp32 [S,64,2] = (cos, sin). */
int rge_op_rope_table(const float* ids, float* cs, int32_t S, void* stream);

/* ids_gather / ids_scatter (utils.py:240-279) on bf16 rows of `width` elements. */
int rge_gather_rows(const void* src, int64_t lds, const int32_t* ids, int32_t n, int32_t width, void* dst,
                    int64_t ldd, void* stream);
int rge_scatter_rows(const void* src, int64_t lds, const int32_t* ids, int32_t n, int32_t width, void* dst,
                     int64_t ldd, void* stream);

/* Latent pack / unpack either side of the loop: diffusers FluxKontextPipeline._pack_latents / _unpack_latents as the
 * reference calls them (RegionE/FluxKontext/inplace.py:212-226 via prepare_latents, :398). latents: bf16
 * [batch, channels, height, width] (height, width even); packed: bf16 [batch, (height/2)*(width/2), 4*channels] with
 * packed channel = 4*c + 2*dy + dx. */
int rge_pack_latents(const void* latents, void* packed, int32_t batch, int32_t channels, int32_t height,
                     int32_t width, void* stream);
int rge_unpack_latents(const void* packed, void* latents, int32_t batch, int32_t channels, int32_t height,
                       int32_t width, void* stream);

/* Scheduler step (inplace.py:610-686): x' = bf16(float(x) + bf16(dt_row * v)), dt_row = edited_mask ?
 * (edited_mask[m] ? dt : dt_direct) : dt. With reuse_on != 0 the velocity is the velocity-decay cache reuse
 * v = bf16(cache * ratio) (inplace.py:318) fused in. x, v, out: [M, channels]. */
int rge_euler(const void* x, const void* v, void* out, int32_t M, int32_t channels, float dt, float dt_direct,
              const uint8_t* edited_mask, int32_t reuse_on, float ratio, void* stream);

/* Adaptive region partition, similarity half (utils.py:305-333 + inplace.py:650): one-step estimate
 * x0 = float(x) + bf16(dt_final * v); mask[m] = cosine(x0[m], cond[m]) <= threshold. sim_out (fp32 [L]) optional. */
int rge_partition(const void* x, const void* v, const void* cond, float dt_final, float threshold,
                  uint8_t* mask_out, float* sim_out, int32_t L, int32_t channels, void* stream);

/* Morphology (erosion 3x3 cross, dilation 5x5 square, zero padding; utils.py:215-237) + ascending index lists
 * (utils.py:345-352). counts (device int32[2]) = {n_edited, n_unedited}. mask_out may be NULL. */
int rge_compact(const uint8_t* mask_in, uint8_t* mask_out, int32_t grid_h, int32_t grid_w, int32_t erosion_dilation,
                int32_t* edited_ids, int32_t* unedited_ids, int32_t* counts, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Engine: the patched transformer forward (inplace.py:413-576 + processor :694-824) as one call per step
 * ---------------------------------------------------------------------------------------------------------- */

typedef struct rge_handle rge_handle;

typedef struct rge_config {
  int32_t dim;            /* 3072 */
  int32_t heads;          /* 24 (head_dim must be 128) */
  int32_t n_double;       /* 19 */
  int32_t n_single;       /* 38 */
  int32_t mlp_ratio;      /* 4 */
  int32_t in_channels;    /* 64 */
  int32_t ctx_dim;        /* 4096 */
  int32_t pooled_dim;     /* 768 */
  int32_t txt_len;        /* T */
  int32_t lat_len;        /* L noise tokens */
  int32_t cond_len;       /* C instruction-image tokens */
  int32_t guidance_embeds;
  int32_t n_pass;         /* independent text/KV-cache sets (1; 2 for CFG pairs) */
  int32_t device;
  int32_t external_embed; /* bit 0: rotary table and embedded context come from the family's own pos_embed / text
                             front end through rge_begin_image_ex (Qwen: txt_norm + txt_in + pos_embed; Step1X:
                             connector + context_embedder); RGE_G_CTX_EMBED_* stay unset.
                             bit 1: temb is computed by the family's own modules and handed to rge_dit_step_ex
                             (Step1X: time_embed + vec_embed); RGE_G_TIME1..POOL2 stay unset.
                             pooled_dim = 0 drops the pooled-text term of temb (Qwen). */
  int32_t shared_cache;   /* 1: every pass reads and writes ONE K/V cache set (FLUX true-CFG: the reference's processor
                             owns a single k_cache / v_cache that both the prompt and the negative-prompt forward patch,
                             RegionE/FluxKontext/inplace.py:349-364, :700-749); 0: one cache set per pass (Qwen /
                             Step1X: k_cache_even / _odd, batch rows) */
} rge_config;

enum rge_block_kind { RGE_BLK_GLOBAL = 0, RGE_BLK_DOUBLE = 1, RGE_BLK_SINGLE = 2 };

/* weight slots; *_W are [out, in] nn.Linear weights, *_B biases, NORM_* RMSNorm weights [128] */
enum rge_global_slot {
  RGE_G_X_EMBED_W, RGE_G_X_EMBED_B, RGE_G_CTX_EMBED_W, RGE_G_CTX_EMBED_B,
  RGE_G_TIME1_W, RGE_G_TIME1_B, RGE_G_TIME2_W, RGE_G_TIME2_B,
  RGE_G_GUID1_W, RGE_G_GUID1_B, RGE_G_GUID2_W, RGE_G_GUID2_B,
  RGE_G_POOL1_W, RGE_G_POOL1_B, RGE_G_POOL2_W, RGE_G_POOL2_B,
  RGE_G_NORM_OUT_W, RGE_G_NORM_OUT_B, RGE_G_PROJ_OUT_W, RGE_G_PROJ_OUT_B,
  RGE_G_NUM_SLOTS
};
enum rge_double_slot {
  RGE_D_MOD_W, RGE_D_MOD_B, RGE_D_MOD_CTX_W, RGE_D_MOD_CTX_B,
  RGE_D_Q_W, RGE_D_Q_B, RGE_D_K_W, RGE_D_K_B, RGE_D_V_W, RGE_D_V_B,
  RGE_D_ADD_Q_W, RGE_D_ADD_Q_B, RGE_D_ADD_K_W, RGE_D_ADD_K_B, RGE_D_ADD_V_W, RGE_D_ADD_V_B,
  RGE_D_NORM_Q, RGE_D_NORM_K, RGE_D_NORM_ADD_Q, RGE_D_NORM_ADD_K,
  RGE_D_OUT_W, RGE_D_OUT_B, RGE_D_ADD_OUT_W, RGE_D_ADD_OUT_B,
  RGE_D_FF_UP_W, RGE_D_FF_UP_B, RGE_D_FF_DOWN_W, RGE_D_FF_DOWN_B,
  RGE_D_FFC_UP_W, RGE_D_FFC_UP_B, RGE_D_FFC_DOWN_W, RGE_D_FFC_DOWN_B,
  RGE_D_NUM_SLOTS
};
enum rge_single_slot {
  RGE_S_MOD_W, RGE_S_MOD_B,
  RGE_S_Q_W, RGE_S_Q_B, RGE_S_K_W, RGE_S_K_B, RGE_S_V_W, RGE_S_V_B,
  RGE_S_NORM_Q, RGE_S_NORM_K,
  RGE_S_MLP_W, RGE_S_MLP_B, RGE_S_OUT_W, RGE_S_OUT_B,
  RGE_S_NUM_SLOTS
};

/* Allocates workspaces and the per-layer K/V cache ([n_pass][layer][T+L+C, dim] x2). Synchronous. */
int rge_create(const rge_config* cfg, rge_handle** out);
int rge_destroy(rge_handle* h);
/* Registers one borrowed weight pointer (what enable() reads off the pipeline modules). */
int rge_set_weight(rge_handle* h, int32_t block_kind, int32_t block_index, int32_t slot, const void* ptr);
/* Checks every slot is set and builds the device-side job tables. Synchronous. */
int rge_finalize_weights(rge_handle* h);

/* Per image and pass (≙ MANAGER.refresh, utils.py:437-465, plus the step-invariant front end of the forward):
 * rotary table from txt_ids [T,3] / img_ids [L+C,3] (fp32), context_embedder(enc [T, ctx_dim]) and the
 * guidance / pooled halves of time_text_embed (inplace.py:471-480, 495-499). */
int rge_begin_image(rge_handle* h, int32_t pass, const float* txt_ids, const float* img_ids, const void* enc,
                    const void* pooled, float guidance_x1000, void* stream);

/* One transformer forward on the active image tokens.
 *   x_in   [n_x, in_channels]    packed latents of the active noise tokens
 *   x_cond [n_cond, in_channels] packed instruction-image (condition) tokens that follow them, or NULL / 0: FULL steps
 *                                pass the condition latent here instead of concatenating it behind the noise latent
 *                                (inplace.py:332); n_img = n_x + n_cond active image tokens in total
 *   sel   int32 [n_img] or NULL  position of each active token in the full [L+C] image sequence; NULL = identity
 *                                (FULL step, n_img must be L+C); REGION step: the edited ids (inplace.py:727-732)
 *   timestep_x1000               the bf16-rounded timestep the reference feeds time_text_embed
 *   v_out [n_out, in_channels]   velocity of the first n_out active tokens (the noise tokens; inplace.py:347)
 * K/V rows of the active tokens (and of all text tokens) are written into the persistent cache of every layer,
 * attention runs active-Q x full cache. A FULL step therefore (re)writes the whole cache, which subsumes the
 * reference's "write cache at warmup-1 / refresh" modes (inplace.py:717-725).
 * Stream semantics: independent launches of a block fork onto library-owned side streams (one of them high priority)
 * and join back into `stream` before the call returns to the host, so for the caller the call is asynchronous on
 * `stream` like every other entry point; two handles must not run concurrently on one device. */
int rge_dit_step(rge_handle* h, int32_t pass, const void* x_in, int32_t n_x, const void* x_cond, int32_t n_cond,
                 const int32_t* sel, float timestep_x1000, void* v_out, int32_t n_out, void* stream);

/* Variants for families whose front end is not FLUX's (rge_config.external_embed = 1).
 *   rope_cs       fp32 [T+L+C, 64, 2] (cos, sin) per rotary pair for the FULL key sequence, text rows first — what
 *                 the pipeline's pos_embed returns (Qwen: QwenImageEdit/inplace.py:530-531, 851-855)
 *   ctx_embedded  [T, dim] text tokens after the family's own context embedding (may be NULL in begin and supplied per
 *                 step instead: Step1X's connector depends on the timestep, Step1XEdit/inplace.py:514-520)
 *   temb          [dim] conditioning vector the adaLN modulations are computed from */
/* Text length of one pass (default rge_config.txt_len = the maximum). Step1X-Edit v1p2 runs the cond and uncond
 * prompts with their own lengths (`txt_length` / `neg_txt_length`, Step1XEditV1P2/utils.py:444-445); the pass then
 * uses rows [0, txt_len) for text and [txt_len, txt_len + L + C) for image tokens in its cache and rotary table.
 * Call before rge_begin_image[_ex] of that pass. */
int rge_set_pass_text_len(rge_handle* h, int32_t pass, int32_t txt_len);
int rge_begin_image_ex(rge_handle* h, int32_t pass, const float* rope_cs, const void* ctx_embedded, void* stream);
int rge_dit_step_ex(rge_handle* h, int32_t pass, const void* x_in, int32_t n_x, const void* x_cond, int32_t n_cond,
                    const int32_t* sel, const void* temb, const void* ctx_embedded, void* v_out, int32_t n_out,
                    void* stream);

/* Optional timing of the engine's tensor-core launches with CUDA events on the streams they run on (bench.py
 * roofline). rge_profile_collect synchronises the device and returns, per class c (0 = GEMM, 1 = attention):
 * ms_busy = length of the UNION of the launches' [start, end] intervals (independent GEMMs of one block overlap on the
 * library's side streams), ms_sum = plain sum of the durations, work = summed algorithmic FLOPs (2MNK resp.
 * 4*Sq*Skv*128*H), count = launches; then clears. */
int rge_profile_enable(int32_t on);
int rge_profile_collect(double* ms_busy, double* ms_sum, double* work, int64_t* count);

/* Number of kernels this library has launched since creation of the process (bench.py's gpu_launches). */
int64_t rge_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* REGIONE_B200_H */
