// Fused GEMM epilogues shared by the 1-CTA (gemm.cu) and 2-CTA (gemm2.cu) tcgen05 kernels: one thread owns one
// output row of the accumulator tile in TMEM and walks its columns in chunks of 32.
#pragma once
#include <cstdlib>
#include "gemm.cuh"
#include "ptx.cuh"

namespace rge {

struct GemmDev {
  int M, N, K;
  const __nv_bfloat16* bias;
  __nv_bfloat16* out;
  long ldo;
  const int* row_map;
  int row_off, col_off;
  const __nv_bfloat16* gate;
  const __nv_bfloat16* res;
  long ldr;
  const __nv_bfloat16* norm_w;
  const float2* rope_cs;
  const int* rope_map;
  int rope_off;
  int n_fast;   // tile order: 0 = consecutive tiles walk down M (W tile shared, A streamed), 1 = walk along N
};

// Tile order. A wave of concurrent tiles re-streams one operand from L2 / DRAM for every few blocks of the other
// dimension. Walking down M keeps the current W tiles hot and streams A once per ~2 column blocks, which is free while
// A (M x K) fits in the 126 MB L2 (K = 3072: 53 MB) and costs 4-5x the algorithmic DRAM reads when it does not
// (FF-down, K = 12288: 201 MB; single-block proj_out, K = 15360: 267 MB; ncu: 1.28 GB read per launch instead of
// 0.33 GB). Those launches walk along N instead: A is read once band by band and W (75 - 94 MB) is the operand that
// lives in L2. RGE_RASTER=m|n forces one order (tuning).
inline int pick_n_fast(const GemmArgs& a) {
  static int forced = -2;
  if (forced == -2) {
    const char* env = getenv("RGE_RASTER");
    forced = !env ? -1 : (env[0] == 'n' ? 1 : (env[0] == 'm' ? 0 : -1));
  }
  if (forced >= 0) return forced;
  const double a_bytes = 2.0 * a.M * a.K, w_bytes = 2.0 * a.N * a.K;
  return a_bytes > 96e6 && w_bytes < a_bytes ? 1 : 0;
}

inline GemmDev to_dev(const GemmArgs& a) {
  GemmDev p;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.bias = a.bias; p.out = a.out; p.ldo = a.ldo; p.row_map = a.row_map; p.row_off = a.row_off; p.col_off = a.col_off;
  p.gate = a.gate; p.res = a.res; p.ldr = a.ldr;
  p.norm_w = a.norm_w; p.rope_cs = a.rope_cs; p.rope_map = a.rope_map; p.rope_off = a.rope_off;
  p.n_fast = 0;
  return p;
}

__device__ __forceinline__ float ldg_bf16f(const __nv_bfloat16* p) { return __bfloat162float(__ldg(p)); }

// taddr: TMEM address of this thread's row (lane field set) at column 0 of the accumulator; m: global row;
// n_base: first global column of the tile; bn: tile width (multiple of 32; of 128 for EPI_NORM_ROPE). All 32 lanes of
// the warp must call this together.
template <int EPI>
__device__ __forceinline__ void gemm_epilogue_row(const GemmDev& p, uint32_t taddr, int m, int n_base, int bn) {
  const bool valid = m < p.M;
  const long out_row = valid ? (long)((p.row_map ? __ldg(p.row_map + m) : m) + p.row_off) : 0;
  __nv_bfloat16* out_ptr = p.out + out_row * p.ldo + p.col_off;
  if constexpr (EPI == EPI_NORM_ROPE) {
    const long rope_row = valid ? (long)((p.rope_map ? __ldg(p.rope_map + m) : m) + p.rope_off) : 0;
    const float2* cs_row = p.rope_cs + rope_row * 64;
#pragma unroll 1
    for (int h = 0; h < bn / 128; ++h) {
      const int n0 = n_base + h * 128;
      if (n0 >= p.N) break;
      float ss = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(taddr + h * 128 + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float b = p.bias ? ldg_bf16f(p.bias + n0 + c * 32 + j) : 0.f;
          float x = bf16_round(__uint_as_float(v[j]) + b);
          ss += x * x;
        }
      }
      const float rstd = rsqrtf(ss * (1.0f / 128.0f) + 1e-6f);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(taddr + h * 128 + c * 32, v);
        tmem_ld_wait();
        uint32_t o[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const int d = c * 32 + j;
          float b0 = p.bias ? ldg_bf16f(p.bias + n0 + d) : 0.f;
          float b1 = p.bias ? ldg_bf16f(p.bias + n0 + d + 1) : 0.f;
          float x0 = bf16_round(__uint_as_float(v[j]) + b0);
          float x1 = bf16_round(__uint_as_float(v[j + 1]) + b1);
          x0 = bf16_round(bf16_round(x0 * rstd) * ldg_bf16f(p.norm_w + d));
          x1 = bf16_round(bf16_round(x1 * rstd) * ldg_bf16f(p.norm_w + d + 1));
          float2 cs = valid ? __ldg(cs_row + (d >> 1)) : make_float2(1.f, 0.f);
          o[j >> 1] = pack_bf16x2(x0 * cs.x - x1 * cs.y, x1 * cs.x + x0 * cs.y);
        }
        if (valid) {
          uint4* dst = reinterpret_cast<uint4*>(out_ptr + n0 + c * 32);
#pragma unroll
          for (int t = 0; t < 4; ++t) dst[t] = make_uint4(o[4 * t], o[4 * t + 1], o[4 * t + 2], o[4 * t + 3]);
        }
      }
    }
  } else {
#pragma unroll 1
    for (int c = 0; c < bn / 32; ++c) {
      const int n0 = n_base + c * 32;
      if (n0 >= p.N) break;
      uint32_t v[32];
      tmem_ld32(taddr + c * 32, v);
      tmem_ld_wait();
      uint32_t o[16];
      uint4 rv[4];
      if constexpr (EPI == EPI_GATE_RES) {
        if (valid) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.res + (long)m * p.ldr + n0);
#pragma unroll
          for (int t = 0; t < 4; ++t) rv[t] = rp[t];
        } else {
#pragma unroll
          for (int t = 0; t < 4; ++t) rv[t] = make_uint4(0, 0, 0, 0);
        }
      }
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        float b0 = p.bias ? ldg_bf16f(p.bias + n0 + j) : 0.f;
        float b1 = p.bias ? ldg_bf16f(p.bias + n0 + j + 1) : 0.f;
        float x0 = bf16_round(__uint_as_float(v[j]) + b0);
        float x1 = bf16_round(__uint_as_float(v[j + 1]) + b1);
        if constexpr (EPI == EPI_GELU) {
          x0 = gelu_tanh(x0);
          x1 = gelu_tanh(x1);
        } else if constexpr (EPI == EPI_GATE_RES) {
          const uint32_t* rw = reinterpret_cast<const uint32_t*>(rv);
          __nv_bfloat162 rr = *reinterpret_cast<const __nv_bfloat162*>(&rw[j >> 1]);
          x0 = __bfloat162float(rr.x) + bf16_round(ldg_bf16f(p.gate + n0 + j) * x0);
          x1 = __bfloat162float(rr.y) + bf16_round(ldg_bf16f(p.gate + n0 + j + 1) * x1);
        }
        o[j >> 1] = pack_bf16x2(x0, x1);
      }
      if (valid) {
        uint4* dst = reinterpret_cast<uint4*>(out_ptr + n0);
#pragma unroll
        for (int t = 0; t < 4; ++t) dst[t] = make_uint4(o[4 * t], o[4 * t + 1], o[4 * t + 2], o[4 * t + 3]);
      }
    }
  }
}

// Implemented in gemm2.cu: cluster-of-2 kernel (tcgen05 cta_group::2, 256 x 256 tiles). Returns
// cudaErrorNotSupported when the shape is outside its envelope so the caller can fall through to the 1-CTA kernel.
cudaError_t launch_gemm_2cta(const GemmArgs& a, int num_sms, cudaStream_t stream);

}  // namespace rge
