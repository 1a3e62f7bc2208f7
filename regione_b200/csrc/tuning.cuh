// Process-wide tuning knobs of the kernels. Each knob starts from its environment variable (read once) and can be
// changed at run time through rge_set_option (include/regione_b200.h) so that a benchmark can sweep variants inside one
// process. Defaults are the fastest measured settings (profiles/).
#pragma once
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

namespace rge {

// Function attributes (dynamic shared memory size) are per device: launchers remember them per device index.
constexpr int kMaxDevices = 64;
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev < 0 || dev >= kMaxDevices ? 0 : dev;
}

struct Tuning {
  int attn_poly;     // RGE_ATTN_POLY:   exponential pairs of every 8 evaluated on the FMA pipe (0, 2, 3, 4)
  int attn_split;    // RGE_ATTN_SPLIT:  1 = K/V split of the ragged last query tile when it alone costs a second wave (default)
  int attn_kernel;   // RGE_ATTN_KERNEL: 0 = attention.cu, 1 = attention64.cu (decoupled pipeline), -1 = default
  int gemm_bn;       // RGE_GEMM_BN:     forced tile width of the 1-CTA GEMM, 0 = choose per launch
  int gemm2_bn;      // RGE_GEMM2_BN:    forced tile width of the CTA-pair GEMM (multiple of 16), 0 = choose per launch
  int min_m_2cta;    // RGE_2CTA_MIN_M:  rows from which the CTA-pair GEMM is used, 0 = never, -1 = per-shape rule (default)
  int raster;        // RGE_RASTER:      -1 = choose per launch, 0 = walk down M, 1 = walk along N
  int wide_store;    // RGE_WIDE_STORE:  1 = 256-bit global stores in the GEMM epilogues where rows are 32-byte aligned
  int trim_last;     // RGE_TRIM_LAST:   1 = the last block computes only the rows whose output is kept (default)
  int split_mod;     // RGE_SPLIT_MOD:   1 = modulation GEMV of all but the first blocks on a side stream (default)
  int nvtx;          // RGE_NVTX:        1 = NVTX ranges per step / block / stage (profilers only)
};

inline int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e && e[0] ? atoi(e) : dflt;
}

inline Tuning& tuning() {
  static Tuning t = [] {
    Tuning x;
    x.attn_poly = env_int("RGE_ATTN_POLY", -1);
    x.attn_split = env_int("RGE_ATTN_SPLIT", 1);
    x.attn_kernel = env_int("RGE_ATTN_KERNEL", -1);
    x.gemm_bn = env_int("RGE_GEMM_BN", 0);
    x.gemm2_bn = env_int("RGE_GEMM2_BN", 0);
    x.min_m_2cta = env_int("RGE_2CTA_MIN_M", -1);
    const char* r = getenv("RGE_RASTER");
    x.raster = !r ? -1 : (r[0] == 'n' ? 1 : (r[0] == 'm' ? 0 : -1));
    x.wide_store = env_int("RGE_WIDE_STORE", 1);
    x.trim_last = env_int("RGE_TRIM_LAST", 1);
    x.split_mod = env_int("RGE_SPLIT_MOD", 1);
    x.nvtx = env_int("RGE_NVTX", 0);
    return x;
  }();
  return t;
}

// returns false for an unknown knob
inline bool set_tuning(const char* name, int value) {
  Tuning& t = tuning();
  if (!strcmp(name, "attn_poly")) t.attn_poly = value;
  else if (!strcmp(name, "attn_split")) t.attn_split = value;
  else if (!strcmp(name, "attn_kernel")) t.attn_kernel = value;
  else if (!strcmp(name, "gemm_bn")) t.gemm_bn = value;
  else if (!strcmp(name, "gemm2_bn")) t.gemm2_bn = value;
  else if (!strcmp(name, "2cta_min_m")) t.min_m_2cta = value;
  else if (!strcmp(name, "raster")) t.raster = value;
  else if (!strcmp(name, "wide_store")) t.wide_store = value;
  else if (!strcmp(name, "trim_last")) t.trim_last = value;
  else if (!strcmp(name, "split_mod")) t.split_mod = value;
  else if (!strcmp(name, "nvtx")) t.nvtx = value;
  else return false;
  return true;
}

}  // namespace rge
