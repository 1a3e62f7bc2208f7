// C ABI (include/regione_b200.h): kernel-level entry points and the engine that runs one patched transformer
// forward (RegionE/FluxKontext/inplace.py:413-576 with the attention processor of :694-824) per call.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <utility>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/regione_b200.h"
#include "attention.cuh"
#include "elementwise.cuh"
#include "gemm.cuh"
#include "tuning.cuh"

using namespace rge;
typedef __nv_bfloat16 bf16;

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

// Optional per-class timing (bench.py's roofline): CUDA events around every GEMM / attention launch of the engine.
enum { PC_GEMM = 0, PC_ATTN = 1, PC_NCLS = 2 };
struct ProfRec { int cls; double work; cudaEvent_t a, b; int m, n, k; };
bool g_prof = false;
cudaEvent_t g_base = nullptr;
std::mutex g_prof_mu;              // the records / event pool may be touched from several host threads
std::vector<ProfRec> g_recs;
std::vector<cudaEvent_t> g_pool;
cudaEvent_t prof_event() {
  cudaEvent_t e;
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_pool.empty()) { e = g_pool.back(); g_pool.pop_back(); return e; }
  }
  cudaEventCreate(&e);
  return e;
}
struct ProfScope {
  cudaStream_t st; int cls; double work; cudaEvent_t a; int m, n, k;
  ProfScope(cudaStream_t s, int c, double w, int m_ = 0, int n_ = 0, int k_ = 0)
      : st(s), cls(c), work(w), a(nullptr), m(m_), n(n_), k(k_) {
    if (g_prof) { a = prof_event(); cudaEventRecord(a, st); }
  }
  ~ProfScope() {
    if (a) {
      cudaEvent_t b = prof_event();
      cudaEventRecord(b, st);
      std::lock_guard<std::mutex> lk(g_prof_mu);
      g_recs.push_back(ProfRec{cls, work, a, b, m, n, k});
    }
  }
};

// NVTX range (header-only nvtx3: a no-op unless a profiler injects itself) around a step / block / stage; enabled
// with RGE_NVTX=1 or rge_set_option("nvtx", 1) so that ncu / nsys timelines can be cut by block and stage.
struct NvtxRange {
  bool on;
  explicit NvtxRange(const char* fmt, int a = 0, int b = 0) : on(tuning().nvtx != 0) {
    if (on) {
      char name[96];
      snprintf(name, sizeof(name), fmt, a, b);
      nvtxRangePushA(name);
    }
  }
  ~NvtxRange() {
    if (on) nvtxRangePop();
  }
};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define RGE_CUDA(expr)                                                                             \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) return fail(RGE_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e));     \
  } while (0)
#define RGE_LAUNCH(expr)                                                                           \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) return fail(RGE_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e));     \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                            \
  } while (0)

int device_sms() {
  static int sms[kMaxDevices] = {};   // per device
  const int dev = current_device();
  if (!sms[dev]) {
    cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    if (sms[dev] <= 0) sms[dev] = 148;
  }
  return sms[dev];
}

GemmArgs to_args(const rge_gemm_desc* d) {
  GemmArgs a;
  a.A = (const bf16*)d->A; a.lda = d->lda;
  a.W = (const bf16*)d->W; a.ldw = d->ldw;
  a.M = d->M; a.N = d->N; a.K = d->K;
  a.bias = (const bf16*)d->bias;
  a.epilogue = d->epilogue;
  a.out = (bf16*)d->out; a.ldo = d->ldo;
  a.row_map = d->row_map; a.row_off = d->row_off; a.col_off = d->col_off;
  a.gate = (const bf16*)d->gate; a.res = (const bf16*)d->res; a.ldr = d->ldr;
  a.norm_w = (const bf16*)d->norm_w; a.rope_cs = (const float2*)d->rope_cs;
  a.rope_map = d->rope_map; a.rope_off = d->rope_off;
  a.rope_ld = d->rope_ld; a.flags = d->flags;
  return a;
}

}  // namespace

extern "C" {

int rge_abi_version(void) { return RGE_ABI_VERSION; }

int rge_set_option(const char* name, int32_t value) {
  if (!name || !set_tuning(name, value)) return fail(RGE_ERR_INVALID, "rge_set_option: unknown option '%s'", name ? name : "");
  return RGE_OK;
}
const char* rge_last_error(void) { return g_err; }
int64_t rge_launch_count(void) { return g_launches.load(); }

int rge_profile_enable(int32_t on) {
  g_prof = on != 0;
  if (g_prof) {
    if (!g_base) RGE_CUDA(cudaEventCreate(&g_base));
    RGE_CUDA(cudaDeviceSynchronize());
    RGE_CUDA(cudaEventRecord(g_base, 0));
    RGE_CUDA(cudaEventSynchronize(g_base));
  }
  return RGE_OK;
}

int rge_profile_collect(double* ms_busy, double* ms_sum, double* work, int64_t* count) {
  if (!ms_busy || !ms_sum || !work || !count) return fail(RGE_ERR_INVALID, "rge_profile_collect: null argument");
  RGE_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_prof_mu);
  std::vector<std::pair<float, float>> iv[PC_NCLS];
  for (int c = 0; c < PC_NCLS; ++c) { ms_busy[c] = 0; ms_sum[c] = 0; work[c] = 0; count[c] = 0; }
  // RGE_PROFILE_DUMP=<file>: append one line per launch (class, stream-ready ms, end ms, M/Sq, N/Skv, K/H) - the
  // timeline used to tune the side-stream fan-out (there is no nsys in this image)
  FILE* dump = nullptr;
  if (const char* path = getenv("RGE_PROFILE_DUMP")) dump = fopen(path, "a");
  for (const ProfRec& r : g_recs) {
    float t0 = 0.f, t1 = 0.f;
    if (g_base && cudaEventElapsedTime(&t0, g_base, r.a) == cudaSuccess &&
        cudaEventElapsedTime(&t1, g_base, r.b) == cudaSuccess) {
      if (dump) fprintf(dump, "%d %.4f %.4f %d %d %d\n", r.cls, t0, t1, r.m, r.n, r.k);
      iv[r.cls].push_back({t0, t1});
      ms_sum[r.cls] += t1 - t0;
      work[r.cls] += r.work;
      count[r.cls] += 1;
    }
    g_pool.push_back(r.a);
    g_pool.push_back(r.b);
  }
  if (dump) { fprintf(dump, "# end of collect\n"); fclose(dump); }
  g_recs.clear();
  // launches of one class overlap on the side streams: the class is "busy" over the union of its intervals
  for (int c = 0; c < PC_NCLS; ++c) {
    std::sort(iv[c].begin(), iv[c].end());
    float lo = 0.f, hi = -1.f;
    for (const auto& x : iv[c]) {
      if (hi < lo || x.first > hi) {
        if (hi >= lo) ms_busy[c] += hi - lo;
        lo = x.first;
        hi = x.second;
      } else if (x.second > hi) {
        hi = x.second;
      }
    }
    if (hi >= lo) ms_busy[c] += hi - lo;
  }
  return RGE_OK;
}

int rge_op_gemm(const rge_gemm_desc* d, void* stream) {
  if (!d || !d->A || !d->W || !d->out) return fail(RGE_ERR_INVALID, "rge_op_gemm: null operand");
  if (d->epilogue < 0 || d->epilogue > 3) return fail(RGE_ERR_INVALID, "rge_op_gemm: bad epilogue %d", d->epilogue);
  if (d->epilogue == RGE_EPI_NORM_ROPE && (!d->norm_w || !d->rope_cs))
    return fail(RGE_ERR_INVALID, "rge_op_gemm: NORM_ROPE needs norm_w and rope_cs");
  const GemmArgs a = to_args(d);
  RGE_LAUNCH(launch_gemm(a, device_sms(), (cudaStream_t)stream));
  return RGE_OK;
}

int rge_op_gemm_group(const rge_gemm_desc* descs, int32_t n, void* stream) {
  if (!descs || n < 1 || n > 6) return fail(RGE_ERR_INVALID, "rge_op_gemm_group: 1..6 descriptors expected");
  GemmArgs a[6];
  for (int i = 0; i < n; ++i) {
    if (descs[i].M > 0 && (!descs[i].A || !descs[i].W || !descs[i].out))
      return fail(RGE_ERR_INVALID, "rge_op_gemm_group: null operand in member %d", i);
    a[i] = to_args(&descs[i]);
  }
  RGE_LAUNCH(launch_gemm_group(a, n, device_sms(), (cudaStream_t)stream));
  return RGE_OK;
}

int rge_op_attention(const rge_attn_desc* d, void* stream) {
  if (!d || !d->Q || !d->K || !d->V || !d->O) return fail(RGE_ERR_INVALID, "rge_op_attention: null operand");
  AttnArgs a;
  a.Q = (const bf16*)d->Q; a.ldq = d->ldq;
  a.K = (const bf16*)d->K; a.ldk = d->ldk;
  a.V = (const bf16*)d->V; a.ldv = d->ldv;
  a.O = (bf16*)d->O; a.ldo = d->ldo;
  a.Sq = d->Sq; a.Skv = d->Skv; a.H = d->H;
  a.scale = d->scale;
  a.workspace = d->workspace;
  a.workspace_bytes = d->workspace && d->workspace_bytes > 0 ? (size_t)d->workspace_bytes : 0;
  RGE_LAUNCH(launch_attention(a, (cudaStream_t)stream));
  return RGE_OK;
}

int64_t rge_attention_workspace_bytes(int32_t H) { return H > 0 ? (int64_t)attention_workspace_bytes(H) : 0; }

int rge_op_ln_modulate(const void* x, int64_t ldx, const void* scale, const void* shift, void* out, int64_t ldo,
                       int32_t M, int32_t D, void* stream) {
  if (!x || !scale || !shift || !out) return fail(RGE_ERR_INVALID, "rge_op_ln_modulate: null operand");
  RGE_LAUNCH(launch_ln_modulate((const bf16*)x, ldx, (const bf16*)scale, (const bf16*)shift, (bf16*)out, ldo, M, D,
                                (cudaStream_t)stream));
  return RGE_OK;
}

int rge_op_rmsnorm(const void* x, int64_t ldx, const void* weight, void* out, int64_t ldo, int32_t M, int32_t D,
                   float eps, void* stream) {
  if (!x || !weight || !out) return fail(RGE_ERR_INVALID, "rge_op_rmsnorm: null operand");
  RGE_LAUNCH(launch_rmsnorm((const bf16*)x, ldx, (const bf16*)weight, (bf16*)out, ldo, M, D, eps,
                            (cudaStream_t)stream));
  return RGE_OK;
}

int rge_cfg_rescale(const void* pos, const void* neg, float scale, void* out, int32_t M, int32_t channels,
                    void* stream) {
  if (M > 0 && (!pos || !neg || !out)) return fail(RGE_ERR_INVALID, "rge_cfg_rescale: null operand");
  RGE_LAUNCH(launch_cfg_rescale((const bf16*)pos, (const bf16*)neg, scale, (bf16*)out, M, channels,
                                (cudaStream_t)stream));
  return RGE_OK;
}

int rge_cfg_diff_norm(const void* pos, const void* neg, void* norm_out, int32_t M, int32_t channels, void* stream) {
  if (M > 0 && (!pos || !neg || !norm_out)) return fail(RGE_ERR_INVALID, "rge_cfg_diff_norm: null operand");
  RGE_LAUNCH(launch_row_diff_norm((const bf16*)pos, (const bf16*)neg, (bf16*)norm_out, M, channels,
                                  (cudaStream_t)stream));
  return RGE_OK;
}

int rge_cfg_combine(const void* pos, const void* neg, float scale, const void* denom, void* out, int32_t M,
                    int32_t channels, void* stream) {
  if (M > 0 && (!pos || !neg || !out)) return fail(RGE_ERR_INVALID, "rge_cfg_combine: null operand");
  RGE_LAUNCH(launch_cfg_combine((const bf16*)pos, (const bf16*)neg, scale, (const bf16*)denom, (bf16*)out, M, channels,
                                (cudaStream_t)stream));
  return RGE_OK;
}

int rge_op_rope_table(const float* ids, float* cs, int32_t S, void* stream) {
  if (!ids || !cs) return fail(RGE_ERR_INVALID, "rge_op_rope_table: null operand");
  RGE_LAUNCH(launch_rope_table(ids, (float2*)cs, S, 0, (cudaStream_t)stream));
  return RGE_OK;
}

int rge_gather_rows(const void* src, int64_t lds, const int32_t* ids, int32_t n, int32_t width, void* dst,
                    int64_t ldd, void* stream) {
  if (n > 0 && (!src || !ids || !dst)) return fail(RGE_ERR_INVALID, "rge_gather_rows: null operand");
  RGE_LAUNCH(launch_gather_rows((const bf16*)src, lds, ids, n, width, (bf16*)dst, ldd, (cudaStream_t)stream));
  return RGE_OK;
}
int rge_scatter_rows(const void* src, int64_t lds, const int32_t* ids, int32_t n, int32_t width, void* dst,
                     int64_t ldd, void* stream) {
  if (n > 0 && (!src || !ids || !dst)) return fail(RGE_ERR_INVALID, "rge_scatter_rows: null operand");
  RGE_LAUNCH(launch_scatter_rows((const bf16*)src, lds, ids, n, width, (bf16*)dst, ldd, (cudaStream_t)stream));
  return RGE_OK;
}

int rge_pack_latents(const void* latents, void* packed, int32_t batch, int32_t channels, int32_t height,
                     int32_t width, void* stream) {
  if (batch > 0 && (!latents || !packed)) return fail(RGE_ERR_INVALID, "rge_pack_latents: null operand");
  if ((height | width) & 1) return fail(RGE_ERR_INVALID, "rge_pack_latents: height and width must be even");
  RGE_LAUNCH(launch_pack_latents((const bf16*)latents, (bf16*)packed, batch, channels, height, width, false,
                                 (cudaStream_t)stream));
  return RGE_OK;
}
int rge_unpack_latents(const void* packed, void* latents, int32_t batch, int32_t channels, int32_t height,
                       int32_t width, void* stream) {
  if (batch > 0 && (!latents || !packed)) return fail(RGE_ERR_INVALID, "rge_unpack_latents: null operand");
  if ((height | width) & 1) return fail(RGE_ERR_INVALID, "rge_unpack_latents: height and width must be even");
  RGE_LAUNCH(launch_pack_latents((const bf16*)packed, (bf16*)latents, batch, channels, height, width, true,
                                 (cudaStream_t)stream));
  return RGE_OK;
}

int rge_euler(const void* x, const void* v, void* out, int32_t M, int32_t channels, float dt, float dt_direct,
              const uint8_t* edited_mask, int32_t reuse_on, float ratio, void* stream) {
  if (M > 0 && (!x || !v || !out)) return fail(RGE_ERR_INVALID, "rge_euler: null operand");
  RGE_LAUNCH(launch_euler((const bf16*)x, (const bf16*)v, (bf16*)out, M, channels, dt, dt_direct, edited_mask,
                          reuse_on, ratio, (cudaStream_t)stream));
  return RGE_OK;
}

int rge_partition(const void* x, const void* v, const void* cond, float dt_final, float threshold,
                  uint8_t* mask_out, float* sim_out, int32_t L, int32_t channels, void* stream) {
  if (!x || !v || !cond || !mask_out) return fail(RGE_ERR_INVALID, "rge_partition: null operand");
  RGE_LAUNCH(launch_arp_similarity((const bf16*)x, (const bf16*)v, (const bf16*)cond, dt_final, threshold, mask_out,
                                   sim_out, L, channels, (cudaStream_t)stream));
  return RGE_OK;
}

int rge_compact(const uint8_t* mask_in, uint8_t* mask_out, int32_t grid_h, int32_t grid_w, int32_t erosion_dilation,
                int32_t* edited_ids, int32_t* unedited_ids, int32_t* counts, void* stream) {
  if (!mask_in || !edited_ids || !unedited_ids || !counts) return fail(RGE_ERR_INVALID, "rge_compact: null operand");
  RGE_LAUNCH(launch_morph_compact(mask_in, mask_out, grid_h, grid_w, erosion_dilation, edited_ids, unedited_ids,
                                  counts, (cudaStream_t)stream));
  return RGE_OK;
}

}  // extern "C"

// ====================================================================================================== engine
struct rge_handle {
  rge_config cfg;
  int T, L, C, S, D, H, Dm, n_layers, num_sms;
  std::vector<const void*> gw, dw, sw;
  bool finalized = false;
  std::vector<char> begun;
  std::vector<int> Tp;   // text length of each pass (<= T): Step1X-Edit v1p2 prompts differ in length
  // workspaces
  bf16 *h = nullptr, *n = nullptr, *q = nullptr, *big = nullptr;
  bf16 *kcache = nullptr, *vcache = nullptr;
  bf16 *mods = nullptr, *small = nullptr;  // small: tproj[256] t1[D] t2[D] temb[D]
  bf16 *ctx = nullptr;                     // [n_pass][T, D]
  bf16 *pass_small = nullptr;              // per pass: gproj[256] g1[D] gemb[D] pooled[pooled_dim] p1[D] pemb[D]
  float2* rope = nullptr;                  // [n_pass][64][S] pair-major (cos, sin): row s of pair p at p * S + s
  float* ids = nullptr;                    // [S, 3] staging
  int *sel_img = nullptr, *sel_all = nullptr;
  GemvJob* jobs = nullptr;                 // [2 + n_mod] step jobs, then per pass 4 image jobs
  int n_mod = 0;
  size_t pass_small_stride = 0;
  // independent GEMMs of one block run on library-owned side streams so that small-M launches (text stream,
  // region steps) fill the SMs the persistent grid of a neighbour leaves idle; joined before every attention
  cudaStream_t aux[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // [3], [4]: text-stream K / V projections
  cudaEvent_t ev_main = nullptr, ev_aux[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}, ev_txt = nullptr;
  bool fanout = true;
  // REGION steps: the attention grid (query-tile pairs x heads) rarely fills a whole number of waves of the SMs. In
  // the single-stream blocks the MLP-up GEMM is independent of attention, so attention runs on a high-priority stream
  // (its CTAs are placed first) and the GEMM, gated on the same events and capped to the SMs the last attention wave
  // leaves idle, runs beside it instead of in front of it.
  cudaStream_t sattn = nullptr;
  cudaEvent_t ev_attn = nullptr, ev_q = nullptr;
  bool fill_attn_tail = true;
  // The adaLN modulation vectors of all blocks are one HBM-bound batched GEMV over 6.5 GB of weights (~0.9 ms) that
  // only depends on the time embedding. The vectors of the first kModHeadBlocks blocks are computed on the caller's
  // stream; the rest runs on this side stream beside those blocks' tensor-bound GEMMs (its CTAs fit next to a GEMM
  // CTA: 16 KB of shared memory) and is joined before the first block that reads it.
  // scratch for the K/V-split of a ragged last query tile (attention.cu); attention launches of one handle never overlap
  void* attn_ws = nullptr;
  size_t attn_ws_bytes = 0;
  cudaStream_t smod = nullptr;
  cudaEvent_t ev_temb = nullptr, ev_mod = nullptr;
  // RGE_GROUPED=1: steps whose GEMMs all take the 1-CTA path (REGION steps: few rows) launch the independent GEMMs of
  // a stage as ONE grouped persistent kernel on the caller's stream instead of fanning them out over the side
  // streams. Off by default: measured 43.5 ms vs 41.5 ms per REGION step (profiles/r01_step_times_grouped*.log) -
  // the pre-attention stage gets faster (119 vs 184 us) but the free-running image / text chains after attention,
  // which overlap with the next block in the fan-out, are serialised at every stage.
  bool grouped = false;
  // RGE_GROUP_QKV (default 1): keep the fan-out but, in REGION-sized steps, launch the q / k / v projections of one
  // stream (image, text, single block) as ONE grouped launch on that stream's chain - fewer launch + prologue +
  // un-overlapped-epilogue costs without serialising independent chains: 41.6 -> 39.4 ms per REGION step
  // (profiles/r02_step_times_launch_variants.log); 2 = image AND text q / k / v of a double block as one launch (40.1 ms);
  // 0 = one launch per projection
  int group_qkv = 1;

  const bf16* G(int slot) const { return (const bf16*)gw[slot]; }
  const bf16* Dw(int b, int slot) const { return (const bf16*)dw[(size_t)b * RGE_D_NUM_SLOTS + slot]; }
  const bf16* Sw(int b, int slot) const { return (const bf16*)sw[(size_t)b * RGE_S_NUM_SLOTS + slot]; }
  int cset(int pass) const { return cfg.shared_cache ? 0 : pass; }
  bf16* kc(int pass, int layer) const { return kcache + ((size_t)cset(pass) * n_layers + layer) * (size_t)S * D; }
  bf16* vc(int pass, int layer) const { return vcache + ((size_t)cset(pass) * n_layers + layer) * (size_t)S * D; }
  bf16* ps(int pass) const { return pass_small + (size_t)pass * pass_small_stride; }
};

namespace {

template <typename T>
cudaError_t dalloc(T** p, size_t count) {
  return cudaMalloc((void**)p, count * sizeof(T));
}

int gemm_group(rge_handle* h, cudaStream_t st, const GemmArgs* a, int n, int sm_cap = 0);

int gemm(rge_handle* h, cudaStream_t st, const bf16* A, long lda, int M, int K, const bf16* W, const bf16* bias, int N,
         int epi, bf16* out, long ldo, const int* row_map, int row_off, int col_off, const bf16* gate = nullptr,
         const bf16* res = nullptr, long ldr = 0, const bf16* norm_w = nullptr, const float2* rope = nullptr,
         const int* rope_map = nullptr, int rope_off = 0, int sm_cap = 0) {
  GemmArgs a;
  a.A = A; a.lda = lda; a.M = M; a.K = K; a.W = W; a.ldw = K; a.N = N; a.bias = bias; a.epilogue = epi;
  a.out = out; a.ldo = ldo; a.row_map = row_map; a.row_off = row_off; a.col_off = col_off;
  a.gate = gate; a.res = res; a.ldr = ldr; a.norm_w = norm_w; a.rope_cs = rope; a.rope_map = rope_map;
  a.rope_off = rope_off;
  if (M <= 0) return RGE_OK;
  return gemm_group(h, st, &a, 1, sm_cap);
}

GemmArgs mk(const bf16* A, long lda, int M, int K, const bf16* W, const bf16* bias, int N, int epi, bf16* out, long ldo,
            const int* row_map, int row_off, int col_off, const bf16* gate = nullptr, const bf16* res = nullptr,
            long ldr = 0, const bf16* norm_w = nullptr, const float2* rope = nullptr, const int* rope_map = nullptr,
            int rope_off = 0, long rope_ld = 0) {
  GemmArgs a;
  a.A = A; a.lda = lda; a.M = M; a.K = K; a.W = W; a.ldw = K; a.N = N; a.bias = bias; a.epilogue = epi;
  a.out = out; a.ldo = ldo; a.row_map = row_map; a.row_off = row_off; a.col_off = col_off;
  a.gate = gate; a.res = res; a.ldr = ldr; a.norm_w = norm_w; a.rope_cs = rope; a.rope_map = rope_map;
  a.rope_off = rope_off; a.rope_ld = rope_ld;
  return a;
}

int gemm_group(rge_handle* h, cudaStream_t st, const GemmArgs* a, int n, int sm_cap) {
  double work = 0;
  for (int i = 0; i < n; ++i)
    if (a[i].M > 0) work += 2.0 * a[i].M * (double)a[i].N * a[i].K;
  if (work == 0) return RGE_OK;
  ProfScope prof(st, PC_GEMM, work, a[0].M, n == 1 ? a[0].N : -n, a[0].K);   // N < 0: a group of |N| members
  if (n == 1) RGE_LAUNCH(launch_gemm(a[0], sm_cap > 0 ? sm_cap : h->num_sms, st));
  else RGE_LAUNCH(launch_gemm_group(a, n, sm_cap > 0 ? sm_cap : h->num_sms, st));
  return RGE_OK;
}

#define RGE_TRY(expr)            \
  do {                           \
    int _r = (expr);             \
    if (_r != RGE_OK) return _r; \
  } while (0)

// One transformer forward's block stack: the launch sequences of a double-stream block and of a
// single-stream block (App. B-2), each in two flavours - fan-out over the library's side streams (default) and one
// grouped launch per stage (RGE_GROUPED=1, REGION-sized steps only).
struct StepRun {
  rge_handle* h;
  cudaStream_t st, sT, sK, sV, sTK, sTV;
  int pass, D, Dm, T, S, M, MA;
  long ldb;                 // `big`: attention output in columns [0, D), MLP hidden in [D, D + Dm)
  const float2* rope;
  bf16 *x_img, *n_img_p, *big_img;
  int idle_sms;             // SMs the last (partial) wave of the attention grid leaves idle
  bool fill_tail;           // ... and whether the single blocks' MLP-up GEMM runs there beside attention

  StepRun(rge_handle* h_, cudaStream_t st_, int pass_, int T_, int M_)
      : h(h_), st(st_), pass(pass_), D(h_->D), Dm(h_->Dm), T(T_), S(T_ + h_->L + h_->C), M(M_), MA(T_ + M_) {
    ldb = D + Dm;
    rope = h->rope + (size_t)pass * h->S * 64;
    x_img = h->h + (size_t)T * D;
    n_img_p = h->n + (size_t)T * D;
    big_img = h->big + (size_t)T * ldb;
    // side streams: sT carries the text chain of the double blocks (and the MLP of the single blocks), sK / sV the
    // image K and V projections, sTK / sTV the text K and V projections
    const bool fan = h->fanout;
    sT = fan ? h->aux[0] : st; sK = fan ? h->aux[1] : st; sV = fan ? h->aux[2] : st;
    sTK = fan ? h->aux[3] : st; sTV = fan ? h->aux[4] : st;
    // the MLP-up GEMM fits into the idle SMs if it takes about one attention-CTA time there (per-SM rates of the two
    // kernels are comparable)
    const int n_attn_ctas = ((MA + 255) / 256) * h->H;
    const int attn_tail = n_attn_ctas % h->num_sms;
    idle_sms = attn_tail ? h->num_sms - attn_tail : 0;
    const double attn_cta_flop = 4.0 * 256.0 * S * 128.0;
    fill_tail = fan && h->fill_attn_tail && idle_sms >= 16 && 2.0 * MA * (double)Dm * D / idle_sms <= 1.3 * attn_cta_flop;
  }

  // makes `to` wait for everything enqueued on `from` so far
  cudaError_t link(cudaStream_t from, cudaEvent_t ev, cudaStream_t to) const {
    if (from == to) return cudaSuccess;
    cudaError_t e = cudaEventRecord(ev, from);
    return e != cudaSuccess ? e : cudaStreamWaitEvent(to, ev, 0);
  }

  // queries = rows [row0, row0 + n_rows) of the [text; image] sequence (default: all active rows)
  // allow_split: the launcher may cut the ragged last query tile along K/V (not beside the capped MLP GEMM, which
  // already fills the attention grid's second wave there)
  int attention(bf16* kc, bf16* vc, cudaStream_t sa, int row0 = 0, int n_rows = -1, bool allow_split = true) const {
    AttnArgs at;
    if (allow_split) { at.workspace = h->attn_ws; at.workspace_bytes = h->attn_ws_bytes; }
    at.Q = h->q + (size_t)row0 * D; at.ldq = D; at.K = kc; at.ldk = D; at.V = vc; at.ldv = D;
    at.O = h->big + (size_t)row0 * ldb; at.ldo = ldb;
    at.Sq = n_rows < 0 ? MA : n_rows; at.Skv = S; at.H = h->H;
    if (at.Sq <= 0) return RGE_OK;
    ProfScope prof(sa, PC_ATTN, 4.0 * at.Sq * (double)at.Skv * 128.0 * at.H, at.Sq, at.Skv, at.H);
    RGE_LAUNCH(launch_attention(at, sa));
    return RGE_OK;
  }

  // ---- the GEMMs of a double block (b) / single block (b); k / v rows are scattered into the cache at T + sel[m]
  GemmArgs img_q(int b) const {
    return mk(n_img_p, D, M, D, h->Dw(b, RGE_D_Q_W), h->Dw(b, RGE_D_Q_B), D, EPI_NORM_ROPE, h->q, D, nullptr, T, 0,
              nullptr, nullptr, 0, h->Dw(b, RGE_D_NORM_Q), rope, h->sel_img, T, h->S);
  }
  GemmArgs img_k(int b, bf16* kc) const {
    return mk(n_img_p, D, M, D, h->Dw(b, RGE_D_K_W), h->Dw(b, RGE_D_K_B), D, EPI_NORM_ROPE, kc, D, h->sel_img, T, 0,
              nullptr, nullptr, 0, h->Dw(b, RGE_D_NORM_K), rope, h->sel_img, T, h->S);
  }
  GemmArgs img_v(int b, bf16* vc) const {
    return mk(n_img_p, D, M, D, h->Dw(b, RGE_D_V_W), h->Dw(b, RGE_D_V_B), D, EPI_STORE, vc, D, h->sel_img, T, 0);
  }
  // text stream q / k / v: recomputed every step (the reference does not cache text K/V)
  GemmArgs txt_q(int b) const {
    return mk(h->n, D, T, D, h->Dw(b, RGE_D_ADD_Q_W), h->Dw(b, RGE_D_ADD_Q_B), D, EPI_NORM_ROPE, h->q, D, nullptr, 0, 0,
              nullptr, nullptr, 0, h->Dw(b, RGE_D_NORM_ADD_Q), rope, nullptr, 0, h->S);
  }
  GemmArgs txt_k(int b, bf16* kc) const {
    return mk(h->n, D, T, D, h->Dw(b, RGE_D_ADD_K_W), h->Dw(b, RGE_D_ADD_K_B), D, EPI_NORM_ROPE, kc, D, nullptr, 0, 0,
              nullptr, nullptr, 0, h->Dw(b, RGE_D_NORM_ADD_K), rope, nullptr, 0, h->S);
  }
  GemmArgs txt_v(int b, bf16* vc) const {
    return mk(h->n, D, T, D, h->Dw(b, RGE_D_ADD_V_W), h->Dw(b, RGE_D_ADD_V_B), D, EPI_STORE, vc, D, nullptr, 0, 0);
  }
  // out projections and feed-forward halves with gate * (.) + residual fused
  GemmArgs img_out(int b, const bf16* gate) const {
    return mk(big_img, ldb, M, D, h->Dw(b, RGE_D_OUT_W), h->Dw(b, RGE_D_OUT_B), D, EPI_GATE_RES, x_img, D, nullptr, 0, 0,
              gate, x_img, D);
  }
  GemmArgs txt_out(int b, const bf16* gate) const {
    return mk(h->big, ldb, T, D, h->Dw(b, RGE_D_ADD_OUT_W), h->Dw(b, RGE_D_ADD_OUT_B), D, EPI_GATE_RES, h->h, D, nullptr,
              0, 0, gate, h->h, D);
  }
  GemmArgs img_up(int b) const {
    return mk(n_img_p, D, M, D, h->Dw(b, RGE_D_FF_UP_W), h->Dw(b, RGE_D_FF_UP_B), Dm, EPI_GELU, big_img, ldb, nullptr, 0,
              D);
  }
  GemmArgs txt_up(int b) const {
    return mk(h->n, D, T, D, h->Dw(b, RGE_D_FFC_UP_W), h->Dw(b, RGE_D_FFC_UP_B), Dm, EPI_GELU, h->big, ldb, nullptr, 0, D);
  }
  GemmArgs img_down(int b, const bf16* gate) const {
    return mk(big_img + D, ldb, M, Dm, h->Dw(b, RGE_D_FF_DOWN_W), h->Dw(b, RGE_D_FF_DOWN_B), D, EPI_GATE_RES, x_img, D,
              nullptr, 0, 0, gate, x_img, D);
  }
  GemmArgs txt_down(int b, const bf16* gate) const {
    return mk(h->big + D, ldb, T, Dm, h->Dw(b, RGE_D_FFC_DOWN_W), h->Dw(b, RGE_D_FFC_DOWN_B), D, EPI_GATE_RES, h->h, D,
              nullptr, 0, 0, gate, h->h, D);
  }
  // single block on [text; image]; selection = [0..T) ++ (T + sel) (inplace.py:730)
  GemmArgs s_q(int b) const {
    return mk(h->n, D, MA, D, h->Sw(b, RGE_S_Q_W), h->Sw(b, RGE_S_Q_B), D, EPI_NORM_ROPE, h->q, D, nullptr, 0, 0, nullptr,
              nullptr, 0, h->Sw(b, RGE_S_NORM_Q), rope, h->sel_all, 0, h->S);
  }
  GemmArgs s_k(int b, bf16* kc) const {
    return mk(h->n, D, MA, D, h->Sw(b, RGE_S_K_W), h->Sw(b, RGE_S_K_B), D, EPI_NORM_ROPE, kc, D, h->sel_all, 0, 0, nullptr,
              nullptr, 0, h->Sw(b, RGE_S_NORM_K), rope, h->sel_all, 0, h->S);
  }
  GemmArgs s_v(int b, bf16* vc) const {
    return mk(h->n, D, MA, D, h->Sw(b, RGE_S_V_W), h->Sw(b, RGE_S_V_B), D, EPI_STORE, vc, D, h->sel_all, 0, 0);
  }
  GemmArgs s_mlp(int b) const {
    return mk(h->n, D, MA, D, h->Sw(b, RGE_S_MLP_W), h->Sw(b, RGE_S_MLP_B), Dm, EPI_GELU, h->big, ldb, nullptr, 0, D);
  }
  GemmArgs s_out(int b, const bf16* gate) const {
    return mk(h->big, ldb, MA, D + Dm, h->Sw(b, RGE_S_OUT_W), h->Sw(b, RGE_S_OUT_B), D, EPI_GATE_RES, h->h, D, nullptr, 0,
              0, gate, h->h, D);
  }
  int one(cudaStream_t s, const GemmArgs& a, int sm_cap = 0) const { return gemm_group(h, s, &a, 1, sm_cap); }
  // grouped q / k / v only where the members take the 1-CTA path anyway (REGION-sized steps, text stream)
  int group_qkv() const { return h->fanout && MA < 2048 ? h->group_qkv : 0; }

  // adaLN vectors of a double block: image stream at mod, text stream at mod + 6 D; each shift, scale, gate x 2
  int double_block_fanout(int b, int layer, const bf16* mod) const {
    NvtxRange range("double block %d (M=%d)", b, M);
    const bf16* cm = mod + 6 * D;
    bf16* kc = h->kc(pass, layer);
    bf16* vc = h->vc(pass, layer);
    // image chain on `st`, text chain on sT, joined around attention
    RGE_LAUNCH(launch_ln_modulate(x_img, D, mod + D, mod, n_img_p, D, M, D, st));
    if (group_qkv() >= 2) {   // all six projections of the block in ONE launch on `st` (the text LayerNorm joins first)
      RGE_LAUNCH(launch_ln_modulate(h->h, D, cm + D, cm, h->n, D, T, D, sT));
      RGE_CUDA(link(sT, h->ev_txt, st));
      const GemmArgs qkv[6] = {img_q(b), img_k(b, kc), img_v(b, vc), txt_q(b), txt_k(b, kc), txt_v(b, vc)};
      RGE_TRY(gemm_group(h, st, qkv, 6));
    } else if (group_qkv()) {
      const GemmArgs iq[3] = {img_q(b), img_k(b, kc), img_v(b, vc)};
      RGE_TRY(gemm_group(h, st, iq, 3));
      RGE_LAUNCH(launch_ln_modulate(h->h, D, cm + D, cm, h->n, D, T, D, sT));
      const GemmArgs tq[3] = {txt_q(b), txt_k(b, kc), txt_v(b, vc)};
      RGE_TRY(gemm_group(h, sT, tq, 3));
      RGE_CUDA(link(sT, h->ev_aux[0], st));
    } else {
      RGE_CUDA(link(st, h->ev_main, sK));
      RGE_CUDA(link(st, h->ev_main, sV));
      RGE_TRY(one(st, img_q(b)));
      RGE_TRY(one(sK, img_k(b, kc)));
      RGE_TRY(one(sV, img_v(b, vc)));
      // the three text projections are small (T rows): on one stream they would run back to back on a mostly idle GPU
      RGE_LAUNCH(launch_ln_modulate(h->h, D, cm + D, cm, h->n, D, T, D, sT));
      RGE_CUDA(link(sT, h->ev_txt, sTK));
      RGE_CUDA(link(sT, h->ev_txt, sTV));
      RGE_TRY(one(sT, txt_q(b)));
      RGE_TRY(one(sTK, txt_k(b, kc)));
      RGE_TRY(one(sTV, txt_v(b, vc)));
      RGE_CUDA(link(sT, h->ev_aux[0], st));
      RGE_CUDA(link(sK, h->ev_aux[1], st));
      RGE_CUDA(link(sV, h->ev_aux[2], st));
      RGE_CUDA(link(sTK, h->ev_aux[3], st));
      RGE_CUDA(link(sTV, h->ev_aux[4], st));
    }
    {
      NvtxRange stage("attention %d x %d", MA, S);
      RGE_TRY(attention(kc, vc, st));
    }
    RGE_CUDA(link(st, h->ev_main, sT));
    NvtxRange stage("out-proj + MLP (image rows %d, text rows %d)", M, T);
    // out projection, LayerNorm, feed-forward: each stream on its own chain
    RGE_TRY(one(st, img_out(b, mod + 2 * D)));
    RGE_TRY(one(sT, txt_out(b, cm + 2 * D)));
    RGE_LAUNCH(launch_ln_modulate(x_img, D, mod + 4 * D, mod + 3 * D, n_img_p, D, M, D, st));
    RGE_LAUNCH(launch_ln_modulate(h->h, D, cm + 4 * D, cm + 3 * D, h->n, D, T, D, sT));
    RGE_TRY(one(st, img_up(b)));
    RGE_TRY(one(sT, txt_up(b)));
    RGE_TRY(one(st, img_down(b, mod + 5 * D)));
    RGE_TRY(one(sT, txt_down(b, cm + 5 * D)));
    return RGE_OK;
  }

  int double_block_grouped(int b, int layer, const bf16* mod) const {
    const bf16* cm = mod + 6 * D;
    bf16* kc = h->kc(pass, layer);
    bf16* vc = h->vc(pass, layer);
    RGE_LAUNCH(launch_ln_modulate(x_img, D, mod + D, mod, n_img_p, D, M, D, st));
    RGE_LAUNCH(launch_ln_modulate(h->h, D, cm + D, cm, h->n, D, T, D, st));
    const GemmArgs qkv[6] = {img_q(b), img_k(b, kc), img_v(b, vc), txt_q(b), txt_k(b, kc), txt_v(b, vc)};
    RGE_TRY(gemm_group(h, st, qkv, 6));
    RGE_TRY(attention(kc, vc, st));
    const GemmArgs outp[2] = {img_out(b, mod + 2 * D), txt_out(b, cm + 2 * D)};
    RGE_TRY(gemm_group(h, st, outp, 2));
    RGE_LAUNCH(launch_ln_modulate(x_img, D, mod + 4 * D, mod + 3 * D, n_img_p, D, M, D, st));
    RGE_LAUNCH(launch_ln_modulate(h->h, D, cm + 4 * D, cm + 3 * D, h->n, D, T, D, st));
    const GemmArgs up[2] = {img_up(b), txt_up(b)};
    RGE_TRY(gemm_group(h, st, up, 2));
    const GemmArgs down[2] = {img_down(b, mod + 5 * D), txt_down(b, cm + 5 * D)};
    RGE_TRY(gemm_group(h, st, down, 2));
    return RGE_OK;
  }

  // attention on the high-priority stream and the MLP-up GEMM, capped to the SMs the last attention wave leaves idle,
  // both start once `ev_q` (and, in the fan-out, the K / V events already recorded) are reached
  int attention_beside_mlp(const GemmArgs& mlp, bf16* kc, bf16* vc, bool wait_kv) const {
    RGE_CUDA(cudaEventRecord(h->ev_q, st));
    for (cudaStream_t s2 : {h->sattn, sT}) {
      RGE_CUDA(cudaStreamWaitEvent(s2, h->ev_q, 0));
      if (wait_kv) {
        RGE_CUDA(cudaStreamWaitEvent(s2, h->ev_aux[1], 0));
        RGE_CUDA(cudaStreamWaitEvent(s2, h->ev_aux[2], 0));
      }
    }
    RGE_TRY(attention(kc, vc, h->sattn, 0, -1, false));
    RGE_TRY(one(sT, mlp, idle_sms));
    RGE_CUDA(link(h->sattn, h->ev_attn, st));
    return RGE_OK;
  }

  // adaLN vectors of a single block at mod: shift, scale, gate
  int single_block_fanout(int b, int layer, const bf16* mod) const {
    NvtxRange range("single block %d (MA=%d)", b, MA);
    bf16* kc = h->kc(pass, layer);
    bf16* vc = h->vc(pass, layer);
    RGE_LAUNCH(launch_ln_modulate(h->h, D, mod + D, mod, h->n, D, MA, D, st));
    if (group_qkv()) {   // q / k / v as one launch on `st`; the MLP GEMM keeps its own stream as below
      if (!fill_tail) {
        RGE_CUDA(link(st, h->ev_main, sT));
        RGE_TRY(one(sT, s_mlp(b)));
      }
      const GemmArgs qkv[3] = {s_q(b), s_k(b, kc), s_v(b, vc)};
      RGE_TRY(gemm_group(h, st, qkv, 3));
      if (!fill_tail) RGE_TRY(attention(kc, vc, st));
      else RGE_TRY(attention_beside_mlp(s_mlp(b), kc, vc, false));
      RGE_CUDA(link(sT, h->ev_aux[0], st));
      return one(st, s_out(b, mod + 2 * D));
    }
    if (!fill_tail) RGE_CUDA(link(st, h->ev_main, sT));
    RGE_CUDA(link(st, h->ev_main, sK));
    RGE_CUDA(link(st, h->ev_main, sV));
    RGE_TRY(one(st, s_q(b)));
    RGE_TRY(one(sK, s_k(b, kc)));
    RGE_TRY(one(sV, s_v(b, vc)));
    if (!fill_tail) {
      // the MLP GEMM (independent of attention, disjoint columns of `big`) may still be running on sT when attention
      // starts: its CTAs and the attention CTAs share the SMs, which fills the partial last wave of either kernel
      RGE_TRY(one(sT, s_mlp(b)));
      RGE_CUDA(link(sK, h->ev_aux[1], st));
      RGE_CUDA(link(sV, h->ev_aux[2], st));
      RGE_TRY(attention(kc, vc, st));
    } else {
      RGE_CUDA(cudaEventRecord(h->ev_aux[1], sK));
      RGE_CUDA(cudaEventRecord(h->ev_aux[2], sV));
      RGE_TRY(attention_beside_mlp(s_mlp(b), kc, vc, true));
    }
    RGE_CUDA(link(sT, h->ev_aux[0], st));
    return one(st, s_out(b, mod + 2 * D));
  }

  // ---- LAST block of the stack: only the first n_out image rows (the noise tokens) reach norm_out / proj_out; the
  // text rows and the instruction-image rows are discarded after it (inplace.py:347, :566-567). K / V still need every
  // row (attention runs over all keys), but queries, attention output, MLP and the output projections are computed for
  // those n_out rows only. Rows are independent in every one of these ops, so the kept rows are bit-identical.
  int single_block_last(int b, int layer, const bf16* mod, int n_out) const {
    NvtxRange range("single block %d, last: %d rows kept", b, n_out);
    bf16* kc = h->kc(pass, layer);
    bf16* vc = h->vc(pass, layer);
    RGE_LAUNCH(launch_ln_modulate(h->h, D, mod + D, mod, h->n, D, MA, D, st));
    RGE_CUDA(link(st, h->ev_main, sK));
    RGE_CUDA(link(st, h->ev_main, sV));
    RGE_CUDA(link(st, h->ev_main, sT));
    GemmArgs q = s_q(b);
    q.A = n_img_p; q.M = n_out; q.row_off = T; q.rope_map = h->sel_all + T;
    RGE_TRY(one(st, q));
    RGE_TRY(one(sK, s_k(b, kc)));
    RGE_TRY(one(sV, s_v(b, vc)));
    GemmArgs mlp = s_mlp(b);
    mlp.A = n_img_p; mlp.M = n_out; mlp.out = big_img;
    RGE_TRY(one(sT, mlp));
    RGE_CUDA(link(sK, h->ev_aux[1], st));
    RGE_CUDA(link(sV, h->ev_aux[2], st));
    RGE_TRY(attention(kc, vc, st, T, n_out));
    RGE_CUDA(link(sT, h->ev_aux[0], st));
    GemmArgs o = s_out(b, mod + 2 * D);
    o.A = big_img; o.M = n_out; o.out = x_img; o.res = x_img;
    return one(st, o);
  }

  int double_block_last(int b, int layer, const bf16* mod, int n_out) const {
    NvtxRange range("double block %d, last: %d rows kept", b, n_out);
    const bf16* cm = mod + 6 * D;
    bf16* kc = h->kc(pass, layer);
    bf16* vc = h->vc(pass, layer);
    RGE_LAUNCH(launch_ln_modulate(x_img, D, mod + D, mod, n_img_p, D, M, D, st));
    RGE_CUDA(link(st, h->ev_main, sK));
    RGE_CUDA(link(st, h->ev_main, sV));
    GemmArgs q = img_q(b);
    q.M = n_out;
    RGE_TRY(one(st, q));
    RGE_TRY(one(sK, img_k(b, kc)));
    RGE_TRY(one(sV, img_v(b, vc)));
    // text keys / values are still attended to; text queries and the whole text chain after attention are not needed
    RGE_LAUNCH(launch_ln_modulate(h->h, D, cm + D, cm, h->n, D, T, D, sT));
    RGE_CUDA(link(sT, h->ev_txt, sTV));
    RGE_TRY(one(sT, txt_k(b, kc)));
    RGE_TRY(one(sTV, txt_v(b, vc)));
    RGE_CUDA(link(sT, h->ev_aux[0], st));
    RGE_CUDA(link(sK, h->ev_aux[1], st));
    RGE_CUDA(link(sV, h->ev_aux[2], st));
    RGE_CUDA(link(sTV, h->ev_aux[4], st));
    RGE_TRY(attention(kc, vc, st, T, n_out));
    GemmArgs o = img_out(b, mod + 2 * D);
    o.M = n_out;
    RGE_TRY(one(st, o));
    RGE_LAUNCH(launch_ln_modulate(x_img, D, mod + 4 * D, mod + 3 * D, n_img_p, D, n_out, D, st));
    GemmArgs up = img_up(b), down = img_down(b, mod + 5 * D);
    up.M = n_out;
    down.M = n_out;
    RGE_TRY(one(st, up));
    return one(st, down);
  }

  int single_block_grouped(int b, int layer, const bf16* mod) const {
    bf16* kc = h->kc(pass, layer);
    bf16* vc = h->vc(pass, layer);
    RGE_LAUNCH(launch_ln_modulate(h->h, D, mod + D, mod, h->n, D, MA, D, st));
    const GemmArgs proj[4] = {s_q(b), s_k(b, kc), s_v(b, vc), s_mlp(b)};
    if (!fill_tail) {
      RGE_TRY(gemm_group(h, st, proj, 4));
      RGE_TRY(attention(kc, vc, st));
    } else {
      RGE_TRY(gemm_group(h, st, proj, 3));
      RGE_TRY(attention_beside_mlp(s_mlp(b), kc, vc, false));
      RGE_CUDA(link(sT, h->ev_aux[0], st));
    }
    return one(st, s_out(b, mod + 2 * D));
  }
};

}  // namespace

extern "C" {

int rge_create(const rge_config* cfg, rge_handle** out) {
  if (!cfg || !out) return fail(RGE_ERR_INVALID, "rge_create: null argument");
  if (cfg->dim <= 0 || cfg->heads <= 0 || cfg->dim != cfg->heads * 128)
    return fail(RGE_ERR_UNSUPPORTED, "rge_create: head_dim must be 128 (dim %d heads %d)", cfg->dim, cfg->heads);
  if (cfg->dim > 4096 || cfg->pooled_dim > 4096 || cfg->dim % 256)
    return fail(RGE_ERR_UNSUPPORTED, "rge_create: dim must be a multiple of 256 and <= 4096");
  if (cfg->in_channels % 32 || cfg->ctx_dim % 8 || cfg->pooled_dim % 8 || cfg->pooled_dim < 0 || cfg->ctx_dim < 0 ||
      cfg->txt_len < 0 || cfg->lat_len <= 0 ||
      cfg->cond_len < 0 || cfg->n_pass < 1 || cfg->mlp_ratio < 1 || cfg->n_double < 0 || cfg->n_single < 0)
    return fail(RGE_ERR_INVALID, "rge_create: bad shape");
  RGE_CUDA(cudaSetDevice(cfg->device));
  if (!get_tensor_map_encoder()) return fail(RGE_ERR_CUDA, "rge_create: cuTensorMapEncodeTiled not available");
  rge_handle* h = new (std::nothrow) rge_handle();
  if (!h) return fail(RGE_ERR_INVALID, "rge_create: out of host memory");
  h->cfg = *cfg;
  h->T = cfg->txt_len; h->L = cfg->lat_len; h->C = cfg->cond_len; h->S = h->T + h->L + h->C;
  h->D = cfg->dim; h->H = cfg->heads; h->Dm = cfg->dim * cfg->mlp_ratio;
  h->n_layers = cfg->n_double + cfg->n_single;
  cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, cfg->device);
  h->gw.assign(RGE_G_NUM_SLOTS, nullptr);
  h->dw.assign((size_t)cfg->n_double * RGE_D_NUM_SLOTS, nullptr);
  h->sw.assign((size_t)cfg->n_single * RGE_S_NUM_SLOTS, nullptr);
  h->begun.assign(cfg->n_pass, 0);
  h->Tp.assign(cfg->n_pass, cfg->txt_len);
  const size_t S = h->S, D = h->D;
  h->n_mod = cfg->n_double * 2 + cfg->n_single + 1;
  const size_t mods_elems = (size_t)cfg->n_double * 12 * D + (size_t)cfg->n_single * 3 * D + 2 * D;
  h->pass_small_stride = 256 + 5 * D + cfg->pooled_dim + 64;
  cudaError_t e = cudaSuccess;
  auto A = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  A(dalloc(&h->h, S * D));
  A(dalloc(&h->n, S * D));
  A(dalloc(&h->q, S * D));
  A(dalloc(&h->big, S * (D + (size_t)h->Dm)));
  const size_t n_sets = cfg->shared_cache ? 1 : cfg->n_pass;
  A(dalloc(&h->kcache, n_sets * h->n_layers * S * D));
  A(dalloc(&h->vcache, n_sets * h->n_layers * S * D));
  A(dalloc(&h->mods, mods_elems));
  A(dalloc(&h->small, 256 + 3 * D));
  A(dalloc(&h->ctx, (size_t)cfg->n_pass * h->T * D + 8));
  A(dalloc(&h->pass_small, (size_t)cfg->n_pass * h->pass_small_stride));
  A(dalloc(&h->rope, (size_t)cfg->n_pass * S * 64));
  A(dalloc(&h->ids, S * 3));
  A(dalloc(&h->sel_img, S));
  A(dalloc(&h->sel_all, S));
  A(dalloc(&h->jobs, (size_t)2 + h->n_mod + 4 * cfg->n_pass));
  for (int i = 0; i < 5; ++i) {
    A(cudaStreamCreateWithFlags(&h->aux[i], cudaStreamNonBlocking));
    A(cudaEventCreateWithFlags(&h->ev_aux[i], cudaEventDisableTiming));
  }
  A(cudaEventCreateWithFlags(&h->ev_txt, cudaEventDisableTiming));
  A(cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming));
  {
    int least = 0, greatest = 0;
    A(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    A(cudaStreamCreateWithPriority(&h->sattn, cudaStreamNonBlocking, greatest));
    A(cudaEventCreateWithFlags(&h->ev_attn, cudaEventDisableTiming));
    A(cudaEventCreateWithFlags(&h->ev_q, cudaEventDisableTiming));
  }
  h->attn_ws_bytes = attention_workspace_bytes(h->H);
  A(cudaMalloc(&h->attn_ws, h->attn_ws_bytes));
  A(cudaStreamCreateWithFlags(&h->smod, cudaStreamNonBlocking));
  A(cudaEventCreateWithFlags(&h->ev_temb, cudaEventDisableTiming));
  A(cudaEventCreateWithFlags(&h->ev_mod, cudaEventDisableTiming));
  if (const char* env = getenv("RGE_NO_FANOUT")) h->fanout = env[0] == '0' || env[0] == 0;
  if (const char* env = getenv("RGE_FILL_ATTN_TAIL")) h->fill_attn_tail = env[0] != '0';
  if (const char* env = getenv("RGE_GROUPED")) h->grouped = env[0] != '0';
  if (const char* env = getenv("RGE_GROUP_QKV")) h->group_qkv = atoi(env);
  if (e != cudaSuccess) {
    rge_destroy(h);
    return fail(RGE_ERR_CUDA, "rge_create: allocation failed: %s", cudaGetErrorString(e));
  }
  // the cache must never expose uninitialised rows to attention
  cudaMemset(h->kcache, 0, n_sets * h->n_layers * S * D * sizeof(bf16));
  cudaMemset(h->vcache, 0, n_sets * h->n_layers * S * D * sizeof(bf16));
  RGE_CUDA(cudaDeviceSynchronize());
  *out = h;
  return RGE_OK;
}

int rge_destroy(rge_handle* h) {
  if (!h) return RGE_OK;
  void* ptrs[] = {h->h, h->n, h->q, h->big, h->kcache, h->vcache, h->mods, h->small, h->ctx, h->pass_small,
                  h->rope, h->ids, h->sel_img, h->sel_all, h->jobs};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (int i = 0; i < 5; ++i) {
    if (h->aux[i]) cudaStreamDestroy(h->aux[i]);
    if (h->ev_aux[i]) cudaEventDestroy(h->ev_aux[i]);
  }
  if (h->ev_txt) cudaEventDestroy(h->ev_txt);
  if (h->ev_main) cudaEventDestroy(h->ev_main);
  if (h->attn_ws) cudaFree(h->attn_ws);
  if (h->smod) cudaStreamDestroy(h->smod);
  if (h->ev_temb) cudaEventDestroy(h->ev_temb);
  if (h->ev_mod) cudaEventDestroy(h->ev_mod);
  if (h->sattn) cudaStreamDestroy(h->sattn);
  if (h->ev_attn) cudaEventDestroy(h->ev_attn);
  if (h->ev_q) cudaEventDestroy(h->ev_q);
  delete h;
  return RGE_OK;
}

int rge_set_weight(rge_handle* h, int32_t kind, int32_t index, int32_t slot, const void* ptr) {
  if (!h || !ptr) return fail(RGE_ERR_INVALID, "rge_set_weight: null argument");
  if (kind == RGE_BLK_GLOBAL) {
    if (slot < 0 || slot >= RGE_G_NUM_SLOTS) return fail(RGE_ERR_INVALID, "rge_set_weight: bad global slot %d", slot);
    h->gw[slot] = ptr;
  } else if (kind == RGE_BLK_DOUBLE) {
    if (index < 0 || index >= h->cfg.n_double || slot < 0 || slot >= RGE_D_NUM_SLOTS)
      return fail(RGE_ERR_INVALID, "rge_set_weight: bad double block %d slot %d", index, slot);
    h->dw[(size_t)index * RGE_D_NUM_SLOTS + slot] = ptr;
  } else if (kind == RGE_BLK_SINGLE) {
    if (index < 0 || index >= h->cfg.n_single || slot < 0 || slot >= RGE_S_NUM_SLOTS)
      return fail(RGE_ERR_INVALID, "rge_set_weight: bad single block %d slot %d", index, slot);
    h->sw[(size_t)index * RGE_S_NUM_SLOTS + slot] = ptr;
  } else {
    return fail(RGE_ERR_INVALID, "rge_set_weight: bad block kind %d", kind);
  }
  h->finalized = false;
  return RGE_OK;
}

int rge_finalize_weights(rge_handle* h) {
  if (!h) return fail(RGE_ERR_INVALID, "rge_finalize_weights: null handle");
  for (int s = 0; s < RGE_G_NUM_SLOTS; ++s) {
    const bool guid = s >= RGE_G_GUID1_W && s <= RGE_G_GUID2_B;
    const bool ctxs = s == RGE_G_CTX_EMBED_W || s == RGE_G_CTX_EMBED_B;
    const bool pool = s >= RGE_G_POOL1_W && s <= RGE_G_POOL2_B;
    const bool tims = s >= RGE_G_TIME1_W && s <= RGE_G_POOL2_B;         // time / guidance / pooled MLPs
    const bool optional = (guid && !h->cfg.guidance_embeds) || (ctxs && (h->cfg.external_embed & 1)) ||
                          (pool && h->cfg.pooled_dim == 0) || (tims && (h->cfg.external_embed & 2));
    if (!h->gw[s] && !optional)
      return fail(RGE_ERR_STATE, "rge_finalize_weights: global slot %d not set", s);
  }
  for (size_t i = 0; i < h->dw.size(); ++i)
    if (!h->dw[i])
      return fail(RGE_ERR_STATE, "rge_finalize_weights: double block %d slot %d not set", (int)(i / RGE_D_NUM_SLOTS),
                  (int)(i % RGE_D_NUM_SLOTS));
  for (size_t i = 0; i < h->sw.size(); ++i)
    if (!h->sw[i])
      return fail(RGE_ERR_STATE, "rge_finalize_weights: single block %d slot %d not set", (int)(i / RGE_S_NUM_SLOTS),
                  (int)(i % RGE_S_NUM_SLOTS));
  const int D = h->D;
  std::vector<GemvJob> jobs((size_t)2 + h->n_mod + 4 * h->cfg.n_pass);
  bf16* tproj = h->small;
  bf16* t1 = tproj + 256;
  bf16* t2 = t1 + D;
  bf16* temb = t2 + D;
  // time_text_embed.timestep_embedder: Linear(256->D) + SiLU, Linear(D->D)   
  jobs[0] = GemvJob{h->G(RGE_G_TIME1_W), h->G(RGE_G_TIME1_B), tproj, t1, D, 256, 0, 1};
  jobs[1] = GemvJob{h->G(RGE_G_TIME2_W), h->G(RGE_G_TIME2_B), t1, t2, D, D, 0, 0};
  // adaLN modulation of every block in one launch: Linear(silu(temb))
  size_t j = 2;
  bf16* m = h->mods;
  for (int b = 0; b < h->cfg.n_double; ++b) {
    jobs[j++] = GemvJob{h->Dw(b, RGE_D_MOD_W), h->Dw(b, RGE_D_MOD_B), temb, m, 6 * D, D, 1, 0};
    m += 6 * D;
    jobs[j++] = GemvJob{h->Dw(b, RGE_D_MOD_CTX_W), h->Dw(b, RGE_D_MOD_CTX_B), temb, m, 6 * D, D, 1, 0};
    m += 6 * D;
  }
  for (int b = 0; b < h->cfg.n_single; ++b) {
    jobs[j++] = GemvJob{h->Sw(b, RGE_S_MOD_W), h->Sw(b, RGE_S_MOD_B), temb, m, 3 * D, D, 1, 0};
    m += 3 * D;
  }
  jobs[j++] = GemvJob{h->G(RGE_G_NORM_OUT_W), h->G(RGE_G_NORM_OUT_B), temb, m, 2 * D, D, 1, 0};
  // per pass: guidance_embedder and text_embedder (pooled) MLPs, evaluated once per image
  const bool has_image_jobs = !(h->cfg.external_embed & 2) && (h->cfg.guidance_embeds || h->cfg.pooled_dim > 0);
  for (int p = 0; p < h->cfg.n_pass && has_image_jobs; ++p) {
    bf16* gproj = h->ps(p);
    bf16* g1 = gproj + 256;
    bf16* gemb = g1 + D;
    bf16* pooled = gemb + D;
    bf16* p1 = pooled + h->cfg.pooled_dim;
    bf16* pemb = p1 + D;
    if (h->cfg.guidance_embeds) {
      jobs[j++] = GemvJob{h->G(RGE_G_GUID1_W), h->G(RGE_G_GUID1_B), gproj, g1, D, 256, 0, 1};
      jobs[j++] = GemvJob{h->G(RGE_G_POOL1_W), h->G(RGE_G_POOL1_B), pooled, p1, D, h->cfg.pooled_dim, 0, 1};
      jobs[j++] = GemvJob{h->G(RGE_G_GUID2_W), h->G(RGE_G_GUID2_B), g1, gemb, D, D, 0, 0};
      jobs[j++] = GemvJob{h->G(RGE_G_POOL2_W), h->G(RGE_G_POOL2_B), p1, pemb, D, D, 0, 0};
    } else {
      jobs[j++] = GemvJob{h->G(RGE_G_POOL1_W), h->G(RGE_G_POOL1_B), pooled, p1, D, h->cfg.pooled_dim, 0, 1};
      jobs[j++] = GemvJob{h->G(RGE_G_POOL1_W), h->G(RGE_G_POOL1_B), pooled, p1, D, h->cfg.pooled_dim, 0, 1};
      jobs[j++] = GemvJob{h->G(RGE_G_POOL2_W), h->G(RGE_G_POOL2_B), p1, pemb, D, D, 0, 0};
      jobs[j++] = GemvJob{h->G(RGE_G_POOL2_W), h->G(RGE_G_POOL2_B), p1, pemb, D, D, 0, 0};
    }
  }
  RGE_CUDA(cudaMemcpy(h->jobs, jobs.data(), jobs.size() * sizeof(GemvJob), cudaMemcpyHostToDevice));
  h->finalized = true;
  return RGE_OK;
}

int rge_begin_image(rge_handle* h, int32_t pass, const float* txt_ids, const float* img_ids, const void* enc,
                    const void* pooled, float guidance_x1000, void* stream) {
  if (!h || !img_ids || !enc || !pooled) return fail(RGE_ERR_INVALID, "rge_begin_image: null argument");
  if (h->T > 0 && !txt_ids) return fail(RGE_ERR_INVALID, "rge_begin_image: null txt_ids");
  if (!h->finalized) return fail(RGE_ERR_STATE, "rge_begin_image: weights not finalized");
  if (pass < 0 || pass >= h->cfg.n_pass) return fail(RGE_ERR_INVALID, "rge_begin_image: bad pass %d", pass);
  if (h->cfg.external_embed) return fail(RGE_ERR_STATE, "rge_begin_image: handle uses rge_begin_image_ex");
  if (h->cfg.pooled_dim == 0) return fail(RGE_ERR_UNSUPPORTED, "rge_begin_image: pooled_dim 0 needs external_embed");
  cudaStream_t st = (cudaStream_t)stream;
  const int D = h->D, T = h->Tp[pass];
  // key-side rotary table over the FULL sequence (MANAGER.image_rotary_emb, inplace.py:499); query rows are
  // looked up in the same table through the selection, which equals pos_embed(gathered ids) (:495-496)
  if (T > 0) RGE_CUDA(cudaMemcpyAsync(h->ids, txt_ids, (size_t)T * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  RGE_CUDA(cudaMemcpyAsync(h->ids + (size_t)T * 3, img_ids, (size_t)(h->L + h->C) * 3 * sizeof(float),
                           cudaMemcpyDeviceToDevice, st));
  RGE_LAUNCH(launch_rope_table(h->ids, h->rope + (size_t)pass * h->S * 64, T + h->L + h->C, h->S, st));
  // context_embedder (inplace.py:480): same value every step of the image, so computed once
  RGE_TRY(gemm(h, st, (const bf16*)enc, h->cfg.ctx_dim, T, h->cfg.ctx_dim, h->G(RGE_G_CTX_EMBED_W),
               h->G(RGE_G_CTX_EMBED_B), D, EPI_STORE, h->ctx + (size_t)pass * h->T * D, D, nullptr, 0, 0));
  // guidance / pooled halves of time_text_embed (inplace.py:475-479)
  bf16* gproj = h->ps(pass);
  bf16* pooled_buf = gproj + 256 + 2 * D;
  RGE_CUDA(cudaMemcpyAsync(pooled_buf, pooled, (size_t)h->cfg.pooled_dim * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
  if (h->cfg.guidance_embeds) RGE_LAUNCH(launch_timestep_proj(guidance_x1000, gproj, st));
  const GemvJob* pj = h->jobs + 2 + h->n_mod + 4 * pass;
  RGE_LAUNCH(launch_gemv_batch(pj, 2, D, st));
  RGE_LAUNCH(launch_gemv_batch(pj + 2, 2, D, st));
  h->begun[pass] = 1;
  return RGE_OK;
}

static int dit_step_impl(rge_handle* h, int32_t pass, const void* x_in, int32_t n_x, const void* x_cond, int32_t n_cond,
                         const int32_t* sel, float timestep_x1000, const void* ext_temb, const void* ext_ctx,
                         void* v_out, int32_t n_out, void* stream) {
  if (!h || (n_out > 0 && !v_out) || (n_x > 0 && !x_in) || (n_cond > 0 && !x_cond))
    return fail(RGE_ERR_INVALID, "rge_dit_step: null argument");
  if (pass < 0 || pass >= h->cfg.n_pass) return fail(RGE_ERR_INVALID, "rge_dit_step: bad pass %d", pass);
  if (!h->finalized || !h->begun[pass]) return fail(RGE_ERR_STATE, "rge_dit_step: rge_begin_image not called");
  if (n_x < 0 || n_cond < 0) return fail(RGE_ERR_INVALID, "rge_dit_step: negative row count");
  const int n_img = n_x + n_cond;   // active image tokens: the rows of x_in, then the rows of x_cond
  if (n_img > h->L + h->C || (!sel && n_img != h->L + h->C))
    return fail(RGE_ERR_INVALID, "rge_dit_step: n_img %d invalid (sel %s, L+C %d)", n_img, sel ? "set" : "null",
                h->L + h->C);
  if (n_out < 0 || n_out > n_img) return fail(RGE_ERR_INVALID, "rge_dit_step: n_out %d > n_img %d", n_out, n_img);
  cudaStream_t st = (cudaStream_t)stream;
  NvtxRange step_range(sel ? "rge_dit_step REGION (%d active image tokens, pass %d)"
                           : "rge_dit_step FULL (%d image tokens, pass %d)", n_img, pass);
  const int D = h->D, T = h->Tp[pass], M = n_img, MA = T + n_img;
  bf16* tproj = h->small;
  bf16* t2 = tproj + 256 + D;
  bf16* temb = t2 + D;
  const bf16* gemb = h->ps(pass) + 256 + D;
  const bf16* pemb = h->ps(pass) + 256 + 2 * D + h->cfg.pooled_dim + D;

  RGE_LAUNCH(launch_build_selection(sel, M, T, h->sel_img, h->sel_all, st));
  if (ext_temb) {
    // the family's own time/text embedding modules produced temb (Step1X connector/vec_embed, Qwen time_text_embed)
    RGE_CUDA(cudaMemcpyAsync(temb, ext_temb, (size_t)D * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
  } else {
    // ---- temb = timestep_embedder(proj(t)) [+ guidance_embedder(..)] + text_embedder(pooled)
    RGE_LAUNCH(launch_timestep_proj(timestep_x1000, tproj, st));
    RGE_LAUNCH(launch_gemv_batch(h->jobs, 1, D, st));
    RGE_LAUNCH(launch_gemv_batch(h->jobs + 1, 1, D, st));
    if (h->cfg.guidance_embeds && h->cfg.pooled_dim > 0) RGE_LAUNCH(launch_add3(t2, gemb, pemb, temb, D, st));
    else if (h->cfg.pooled_dim > 0) RGE_LAUNCH(launch_add3(t2, pemb, nullptr, temb, D, st));
    else if (h->cfg.guidance_embeds) RGE_LAUNCH(launch_add3(t2, gemb, nullptr, temb, D, st));
    else RGE_CUDA(cudaMemcpyAsync(temb, t2, (size_t)D * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
  }
  // ---- adaLN modulation vectors: the first blocks' on `st`, the rest on the side stream (joined in the block loop)
  constexpr int kModHeadBlocks = 2;
  const int n_blocks = h->cfg.n_double + h->cfg.n_single;
  int head_jobs = h->n_mod, head_blocks = n_blocks;   // default: everything on `st`
  if (h->fanout && tuning().split_mod && n_blocks > kModHeadBlocks + 1) {
    head_blocks = kModHeadBlocks;
    const int hd = h->cfg.n_double < head_blocks ? h->cfg.n_double : head_blocks;
    head_jobs = 2 * hd + (head_blocks - hd);           // two jobs per double block, one per single block
  }
  RGE_LAUNCH(launch_gemv_batch(h->jobs + 2, head_jobs, 6 * D, st));
  if (head_jobs < h->n_mod) {
    RGE_CUDA(cudaEventRecord(h->ev_temb, st));
    RGE_CUDA(cudaStreamWaitEvent(h->smod, h->ev_temb, 0));
    RGE_LAUNCH(launch_gemv_batch(h->jobs + 2 + head_jobs, h->n_mod - head_jobs, 6 * D, h->smod));
    RGE_CUDA(cudaEventRecord(h->ev_mod, h->smod));
  }
  // ---- token embedding: text rows [0,T) come from the per-image context embedding, image rows from x_embedder
  if (T > 0)
    RGE_CUDA(cudaMemcpyAsync(h->h, ext_ctx ? (const bf16*)ext_ctx : h->ctx + (size_t)pass * h->T * D,
                             (size_t)T * D * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
  // x_embedder over the two row ranges (the reference concatenates latents and image_latents first, inplace.py:332;
  // reading them through two pointers spares that copy)
  RGE_TRY(gemm(h, st, (const bf16*)x_in, h->cfg.in_channels, n_x, h->cfg.in_channels, h->G(RGE_G_X_EMBED_W),
               h->G(RGE_G_X_EMBED_B), D, EPI_STORE, h->h, D, nullptr, T, 0));
  RGE_TRY(gemm(h, st, (const bf16*)x_cond, h->cfg.in_channels, n_cond, h->cfg.in_channels, h->G(RGE_G_X_EMBED_W),
               h->G(RGE_G_X_EMBED_B), D, EPI_STORE, h->h, D, nullptr, T + n_x, 0));

  StepRun r(h, st, pass, T, M);
  const bf16* mod = h->mods;
  int layer = 0;
  // A step whose widest GEMM (T + M rows) stays below the CTA-pair kernel's threshold is a REGION step (or a small
  // model): with RGE_GROUPED=1 every stage of a block is then ONE grouped launch on `st`. Otherwise the independent
  // GEMMs of a stage fan out over the side streams (the default: see rge_handle::grouped).
  const bool grouped = h->grouped && r.MA < 2048;
  if (!grouped && h->cfg.n_double > 0) RGE_CUDA(r.link(st, h->ev_main, r.sT));
  // the last block of the stack computes only the rows whose output survives (bit-identical; tuning().trim_last)
  const bool trim = tuning().trim_last && h->fanout && !grouped && n_out < MA;
  const int last_double = trim && h->cfg.n_single == 0 ? h->cfg.n_double - 1 : -1;
  const int last_single = trim ? h->cfg.n_single - 1 : -1;
  for (int b = 0; b < h->cfg.n_double; ++b, ++layer, mod += 12 * D) {
    if (layer == head_blocks) RGE_CUDA(cudaStreamWaitEvent(st, h->ev_mod, 0));
    RGE_TRY(b == last_double ? r.double_block_last(b, layer, mod, n_out)
            : grouped        ? r.double_block_grouped(b, layer, mod)
                             : r.double_block_fanout(b, layer, mod));
  }
  if (!grouped && h->cfg.n_double > 0) RGE_CUDA(r.link(r.sT, h->ev_aux[0], st));
  for (int b = 0; b < h->cfg.n_single; ++b, ++layer, mod += 3 * D) {
    if (layer == head_blocks) RGE_CUDA(cudaStreamWaitEvent(st, h->ev_mod, 0));
    RGE_TRY(b == last_single ? r.single_block_last(b, layer, mod, n_out)
            : grouped        ? r.single_block_grouped(b, layer, mod)
                             : r.single_block_fanout(b, layer, mod));
  }
  bf16* x_img = r.x_img;
  bf16* n_img_p = r.n_img_p;
  // ---- norm_out (scale first, then shift) + proj_out on the noise rows only (App. C-9)
  RGE_LAUNCH(launch_ln_modulate(x_img, D, mod, mod + D, n_img_p, D, n_out, D, st));
  RGE_TRY(gemm(h, st, n_img_p, D, n_out, D, h->G(RGE_G_PROJ_OUT_W), h->G(RGE_G_PROJ_OUT_B), h->cfg.in_channels,
               EPI_STORE, (bf16*)v_out, h->cfg.in_channels, nullptr, 0, 0));
  return RGE_OK;
}

int rge_dit_step(rge_handle* h, int32_t pass, const void* x_in, int32_t n_x, const void* x_cond, int32_t n_cond,
                 const int32_t* sel, float timestep_x1000, void* v_out, int32_t n_out, void* stream) {
  if (h && (h->cfg.external_embed & 2)) return fail(RGE_ERR_STATE, "rge_dit_step: handle uses rge_dit_step_ex");
  return dit_step_impl(h, pass, x_in, n_x, x_cond, n_cond, sel, timestep_x1000, nullptr, nullptr, v_out, n_out,
                       stream);
}

int rge_dit_step_ex(rge_handle* h, int32_t pass, const void* x_in, int32_t n_x, const void* x_cond, int32_t n_cond,
                    const int32_t* sel, const void* temb, const void* ctx_embedded, void* v_out, int32_t n_out,
                    void* stream) {
  if (!temb) return fail(RGE_ERR_INVALID, "rge_dit_step_ex: null temb");
  return dit_step_impl(h, pass, x_in, n_x, x_cond, n_cond, sel, 0.f, temb, ctx_embedded, v_out, n_out, stream);
}

int rge_set_pass_text_len(rge_handle* h, int32_t pass, int32_t txt_len) {
  if (!h) return fail(RGE_ERR_INVALID, "rge_set_pass_text_len: null handle");
  if (pass < 0 || pass >= h->cfg.n_pass) return fail(RGE_ERR_INVALID, "rge_set_pass_text_len: bad pass %d", pass);
  if (txt_len < 0 || txt_len > h->T)
    return fail(RGE_ERR_INVALID, "rge_set_pass_text_len: %d outside [0, %d]", txt_len, h->T);
  h->Tp[pass] = txt_len;
  h->begun[pass] = 0;   // the rotary table / context of the pass must be re-supplied for the new length
  return RGE_OK;
}

int rge_begin_image_ex(rge_handle* h, int32_t pass, const float* rope_cs, const void* ctx_embedded, void* stream) {
  if (!h || !rope_cs) return fail(RGE_ERR_INVALID, "rge_begin_image_ex: null argument");
  if (!h->finalized) return fail(RGE_ERR_STATE, "rge_begin_image_ex: weights not finalized");
  if (pass < 0 || pass >= h->cfg.n_pass) return fail(RGE_ERR_INVALID, "rge_begin_image_ex: bad pass %d", pass);
  if (!(h->cfg.external_embed & 1)) return fail(RGE_ERR_STATE, "rge_begin_image_ex: handle uses rge_begin_image");
  if (!(h->cfg.external_embed & 2) && (h->cfg.guidance_embeds || h->cfg.pooled_dim > 0))
    return fail(RGE_ERR_UNSUPPORTED, "rge_begin_image_ex: guidance / pooled embeddings need an external temb");
  cudaStream_t st = (cudaStream_t)stream;
  const int Tp = h->Tp[pass];
  // kept pair-major ([64][S]) inside the library: the 32 rows of an epilogue warp then read each pair coalesced
  RGE_LAUNCH(launch_rope_transpose((const float2*)rope_cs, h->rope + (size_t)pass * h->S * 64, Tp + h->L + h->C, h->S,
                                   st));
  if (ctx_embedded && Tp > 0)
    RGE_CUDA(cudaMemcpyAsync(h->ctx + (size_t)pass * h->T * h->D, ctx_embedded,
                             (size_t)Tp * h->D * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
  h->begun[pass] = 1;
  return RGE_OK;
}

}  // extern "C"
