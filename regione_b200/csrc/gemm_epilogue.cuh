// Fused GEMM epilogues shared by the 1-CTA (gemm.cu) and 2-CTA (gemm2.cu) tcgen05 kernels: one thread owns one
// output row of the accumulator tile in TMEM and walks its columns in chunks of 32.
#pragma once
#include <cstdlib>
#include "gemm.cuh"
#include <cuda_fp16.h>
#include "ptx.cuh"
#include "tuning.cuh"

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
  long rope_ld;   // 0: rope_cs is [S][64] row-major; > 0: pair-major [64][rope_ld]
  int flags;      // bit 0: EPI_STORE rounds fp32 -> fp16 -> bf16 (the reference's Triton kernel, fused_kernels.py:80);
                  // bit 1 (set by to_dev): output rows are 32-byte aligned, the epilogues use 256-bit stores
  int n_fast;   // tile order: 0 = consecutive tiles walk down M (W tile shared, A streamed), 1 = walk along N
};

// tile index -> (row block, column block) for the two tile orders
__device__ __forceinline__ void tile_decode(const GemmDev& p, int tile, int num_m, int num_n, int& m_blk, int& n_blk) {
  m_blk = p.n_fast ? tile / num_n : tile % num_m;
  n_blk = p.n_fast ? tile % num_n : tile / num_m;
}

// Tile order. A wave of concurrent tiles re-streams one operand from L2 / DRAM for every few blocks of the other
// dimension. Walking down M keeps the current W tiles hot and streams A once per ~2 column blocks, which is free while
// A (M x K) fits in the 126 MB L2 (K = 3072: 53 MB) and costs 4-5x the algorithmic DRAM reads when it does not
// (FF-down, K = 12288: 201 MB; single-block proj_out, K = 15360: 267 MB; ncu: 1.28 GB read per launch instead of
// 0.33 GB). Those launches walk along N instead: A is read once band by band and W (75 - 94 MB) is the operand that
// lives in L2. RGE_RASTER=m|n forces one order (tuning).
inline int pick_n_fast(const GemmArgs& a) {
  const int forced = tuning().raster;
  if (forced >= 0) return forced;
  const double a_bytes = 2.0 * a.M * a.K, w_bytes = 2.0 * a.N * a.K;
  return a_bytes > 96e6 && w_bytes < a_bytes ? 1 : 0;
}

// Measured and rejected (profiles/r02_n_group_l2_hints_rejected.log): sweeping the row bands once per GROUP of column
// blocks so that a W slice stays in the L2 - it does not (781 vs 809 MB of DRAM reads for FF-down, 1-2 % slower), and
// pinning it with evict-last TMA hints while streaming A evict-first loses the A band's reuse (1011 MB, 9 % slower).
inline GemmDev to_dev(const GemmArgs& a) {
  GemmDev p;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.bias = a.bias; p.out = a.out; p.ldo = a.ldo; p.row_map = a.row_map; p.row_off = a.row_off; p.col_off = a.col_off;
  p.gate = a.gate; p.res = a.res; p.ldr = a.ldr;
  p.norm_w = a.norm_w; p.rope_cs = a.rope_cs; p.rope_map = a.rope_map; p.rope_off = a.rope_off;
  p.rope_ld = a.rope_ld;
  // bit 1: every output row segment the epilogues store starts 32-byte aligned -> 256-bit stores
  const bool wide = tuning().wide_store && (reinterpret_cast<uintptr_t>(a.out) % 32 == 0) && (a.ldo % 16 == 0) &&
                    (a.col_off % 16 == 0);
  p.flags = (a.flags & 1) | (wide ? 2 : 0);
  p.n_fast = 0;
  return p;
}

// Eight consecutive bf16 of a per-column vector (bias, gate, RMSNorm weight) as floats: one 16-byte read-only load,
// same address in every lane (broadcast), instead of eight scalar loads. Pointers are 16-byte aligned (checked on the
// host), column offsets are multiples of 8.
__device__ __forceinline__ void ld_vec8(const __nv_bfloat16* ptr, float (&f)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(ptr));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

// NW 32-bit words of one output row: 256-bit stores (STG.256, one full 32-byte sector per instruction) when the row is
// 32-byte aligned (GemmDev::flags bit 1, checked on the host), 128-bit stores otherwise. NW is a multiple of 4.
template <int NW>
__device__ __forceinline__ void store_row_words(__nv_bfloat16* dst, const uint32_t* o, bool wide) {
  if (wide && NW % 8 == 0) {
#pragma unroll
    for (int t = 0; t < NW / 8; ++t)
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 16 * t), "r"(o[8 * t]),
                   "r"(o[8 * t + 1]), "r"(o[8 * t + 2]), "r"(o[8 * t + 3]), "r"(o[8 * t + 4]), "r"(o[8 * t + 5]),
                   "r"(o[8 * t + 6]), "r"(o[8 * t + 7])
                   : "memory");
  } else {
#pragma unroll
    for (int t = 0; t < NW / 4; ++t)
      reinterpret_cast<uint4*>(dst)[t] = make_uint4(o[4 * t], o[4 * t + 1], o[4 * t + 2], o[4 * t + 3]);
  }
}

// fp32 -> fp16 -> fp32: the rounding the reference's Triton scatter-GEMM applies before its store into the bf16 cache
// (`accumulator.to(tl.float16)`, RegionE/FluxKontext/fused_kernels.py:80); only with RGE_GEMM_FP16_ROUNDTRIP.
__device__ __forceinline__ float fp16_roundtrip(float x) { return __half2float(__float2half_rn(x)); }

// One chunk of NG x 8 columns (32, or 16 for the tail of a tile whose width is not a multiple of 32) of the STORE /
// GELU / GATE_RES epilogues: v = accumulator columns [n0, n0 + 8 NG) of this thread's row, rv = the residual row's same
// columns (GATE_RES only).
template <int EPI, int NG = 4>
__device__ __forceinline__ void epilogue_chunk(const GemmDev& p, const uint32_t (&v)[32], const uint4 (&rv)[4],
                                               __nv_bfloat16* out_ptr, int n0, bool valid) {
  uint32_t o[16];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    float b[8], gt[8];
    if (p.bias) ld_vec8(p.bias + n0 + 8 * g, b);
    if constexpr (EPI == EPI_GATE_RES) ld_vec8(p.gate + n0 + 8 * g, gt);
    const uint32_t rw[4] = {rv[g].x, rv[g].y, rv[g].z, rv[g].w};
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      float x0 = __uint_as_float(v[8 * g + j]) + (p.bias ? b[j] : 0.f);
      float x1 = __uint_as_float(v[8 * g + j + 1]) + (p.bias ? b[j + 1] : 0.f);
      if constexpr (EPI == EPI_STORE) {
        if (p.flags & 1) { x0 = fp16_roundtrip(x0); x1 = fp16_roundtrip(x1); }
      } else {
        bf16_round2(x0, x1);
      }
      if constexpr (EPI == EPI_GELU) {
        x0 = gelu_tanh(x0);
        x1 = gelu_tanh(x1);
      } else if constexpr (EPI == EPI_GATE_RES) {
        float g0 = gt[j] * x0, g1 = gt[j + 1] * x1;
        bf16_round2(g0, g1);
        x0 = __uint_as_float(rw[j >> 1] << 16) + g0;
        x1 = __uint_as_float(rw[j >> 1] & 0xffff0000u) + g1;
      }
      o[4 * g + (j >> 1)] = pack_bf16x2(x0, x1);
    }
  }
  if (valid) store_row_words<4 * NG>(out_ptr + n0, o, (p.flags & 2) != 0);
}

template <int EPI, int NG = 4>
__device__ __forceinline__ void load_residual(const GemmDev& p, uint4 (&rv)[4], int m, int n0, bool valid) {
  if constexpr (EPI == EPI_GATE_RES) {
    if (valid) {
      const uint4* rp = reinterpret_cast<const uint4*>(p.res + (long)m * p.ldr + n0);
#pragma unroll
      for (int t = 0; t < NG; ++t) rv[t] = rp[t];
#pragma unroll
      for (int t = NG; t < 4; ++t) rv[t] = make_uint4(0, 0, 0, 0);
      return;
    }
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) rv[t] = make_uint4(0, 0, 0, 0);
}

// Hands a TMEM accumulator back to the MMA issuer: every lane's tcgen05.ld of it has completed (the caller waited),
// one lane arrives on the `tempty` barrier - of this CTA, or of the pair's leader CTA (cluster_cta >= 0).
struct TmemRelease {
  uint64_t* bar;
  int cluster_cta;
  __device__ __forceinline__ void operator()() const {
    tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
      if (cluster_cta < 0) mbar_arrive(bar);
      else mbar_arrive_cluster(bar, (uint32_t)cluster_cta);
    }
  }
};

// EPI_NORM_ROPE for NH (1 or 2) consecutive heads of this thread's row: per-head RMSNorm of the bf16-rounded
// (acc + bias), weight, rotary embedding, store. ONE pass over TMEM: the rounded values are exactly representable in
// bf16, so the whole head is kept as 64 packed bf16 pairs in registers while the sum of squares accumulates (the
// 32-column TMEM loads are double-buffered), the accumulator is handed back to the MMA issuer (`release`, when this
// is the tile's last head group) BEFORE the normalise / rotate / store pass, and that pass reads each rotary pair once
// for both heads (the rotation depends on the row only), with the next chunk's pairs requested a chunk ahead.
template <int NH>
__device__ __forceinline__ void norm_rope_heads(const GemmDev& p, uint32_t taddr, __nv_bfloat16* out_ptr, int n0,
                                                const float2* cs_base, long cs_step, bool valid,
                                                const TmemRelease* release) {
  uint32_t xp[NH][64];
  float rstd[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    float ss0 = 0.f, ss1 = 0.f;
    uint32_t va[32], vb[32];
    tmem_ld32(taddr + h * 128, va);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      tmem_ld_wait();
      uint32_t (&cur)[32] = (c & 1) ? vb : va;
      if (c < 3) tmem_ld32(taddr + h * 128 + (c + 1) * 32, (c & 1) ? va : vb);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float b[8];
        if (p.bias) ld_vec8(p.bias + n0 + h * 128 + c * 32 + 8 * g, b);
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const float x0 = __uint_as_float(cur[8 * g + j]) + (p.bias ? b[j] : 0.f);
          const float x1 = __uint_as_float(cur[8 * g + j + 1]) + (p.bias ? b[j + 1] : 0.f);
          const uint32_t u = pack_bf16x2(x0, x1);
          xp[h][c * 16 + 4 * g + (j >> 1)] = u;
          const float r0 = __uint_as_float(u << 16), r1 = __uint_as_float(u & 0xffff0000u);
          ss0 = fmaf(r0, r0, ss0);
          ss1 = fmaf(r1, r1, ss1);
        }
      }
    }
    rstd[h] = rsqrtf((ss0 + ss1) * (1.0f / 128.0f) + 1e-6f);
  }
  if (release) (*release)();
  float2 cs[16], cs_next[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) cs[j] = valid ? __ldg(cs_base + (long)j * cs_step) : make_float2(1.f, 0.f);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (c < 3) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        cs_next[j] = valid ? __ldg(cs_base + (long)((c + 1) * 16 + j) * cs_step) : make_float2(1.f, 0.f);
    }
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      uint32_t o[16];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float w[8];
        ld_vec8(p.norm_w + c * 32 + 8 * g, w);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t u = xp[h][c * 16 + 4 * g + j];
          float x0 = __uint_as_float(u << 16) * rstd[h], x1 = __uint_as_float(u & 0xffff0000u) * rstd[h];
          bf16_round2(x0, x1);
          x0 *= w[2 * j];
          x1 *= w[2 * j + 1];
          bf16_round2(x0, x1);
          const float2 t = cs[4 * g + j];
          o[4 * g + j] = pack_bf16x2(x0 * t.x - x1 * t.y, x1 * t.x + x0 * t.y);
        }
      }
      if (valid) store_row_words<16>(out_ptr + n0 + h * 128 + c * 32, o, (p.flags & 2) != 0);
    }
    if (c < 3) {
#pragma unroll
      for (int j = 0; j < 16; ++j) cs[j] = cs_next[j];
    }
  }
}

// taddr: TMEM address of this thread's row (lane field set) at column 0 of the accumulator; m: global row;
// n_base: first global column of the tile; bn: tile width (multiple of 16; of 128 for EPI_NORM_ROPE). All 32 lanes of
// the warp must call this together. kTail16: the tile width may be an odd multiple of 16 (CTA-pair kernel only - the
// extra code path costs the grouped 1-CTA kernel, which inlines all four epilogues, its register budget).
// `release` hands the accumulator back to the MMA issuer; it is called exactly once, as soon as the last TMEM read of
// the tile has completed (EPI_NORM_ROPE: before the store pass).
// kRopeHeads: heads EPI_NORM_ROPE processes together in its single-pass form (2: one rotary read for two heads, ~250
// registers); 0 = the two-pass form with few registers for the grouped kernel, which inlines all four epilogues.
template <int EPI, bool kTail16 = false, int kRopeHeads = 2>
__device__ __forceinline__ void gemm_epilogue_row(const GemmDev& p, uint32_t taddr, int m, int n_base, int bn,
                                                  const TmemRelease& release) {
  const bool valid = m < p.M;
  const long out_row = valid ? (long)((p.row_map ? __ldg(p.row_map + m) : m) + p.row_off) : 0;
  __nv_bfloat16* out_ptr = p.out + out_row * p.ldo + p.col_off;
  if constexpr (EPI == EPI_NORM_ROPE) {
    const long rope_row = valid ? (long)((p.rope_map ? __ldg(p.rope_map + m) : m) + p.rope_off) : 0;
    // rotary table: [S][64] (cos, sin) row-major, or - p.rope_ld > 0 - pair-major [64][rope_ld]: consecutive rows of
    // one pair are then consecutive in memory, so the 32 rows of a warp read 256 contiguous bytes per pair instead of
    // 32 separate sectors
    const float2* cs_base = p.rope_ld > 0 ? p.rope_cs + rope_row : p.rope_cs + rope_row * 64;
    const long cs_step = p.rope_ld > 0 ? p.rope_ld : 1;
    if constexpr (kRopeHeads == 0) {
      // two passes over TMEM, few registers (the grouped kernel inlines all four epilogues)
#pragma unroll 1
      for (int h = 0; h < bn / 128; ++h) {
        const int n0 = n_base + h * 128;
        if (n0 >= p.N) break;
        // pass 1: sum of squares of the bf16-rounded (acc + bias) over the head, two 32-column TMEM loads in flight
        float ss0 = 0.f, ss1 = 0.f;
#pragma unroll 1
        for (int c = 0; c < 4; c += 2) {
          uint32_t v[64];
          tmem_ld32p(taddr + h * 128 + c * 32, v);
          tmem_ld32p(taddr + h * 128 + c * 32 + 32, v + 32);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float b[8];
            if (p.bias) ld_vec8(p.bias + n0 + c * 32 + 8 * g, b);
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              float x0 = __uint_as_float(v[8 * g + j]) + (p.bias ? b[j] : 0.f);
              float x1 = __uint_as_float(v[8 * g + j + 1]) + (p.bias ? b[j + 1] : 0.f);
              bf16_round2(x0, x1);
              ss0 = fmaf(x0, x0, ss0);
              ss1 = fmaf(x1, x1, ss1);
            }
          }
        }
        const float rstd = rsqrtf((ss0 + ss1) * (1.0f / 128.0f) + 1e-6f);
        // pass 2: normalise, weight, rotate, store; the rotary pairs of chunk c are requested before its TMEM load
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld32(taddr + h * 128 + c * 32, v);
          float2 cs[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            cs[j] = valid ? __ldg(cs_base + (long)(c * 16 + j) * cs_step) : make_float2(1.f, 0.f);
          tmem_ld_wait();
          uint32_t o[16];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float b[8], w[8];
            if (p.bias) ld_vec8(p.bias + n0 + c * 32 + 8 * g, b);
            ld_vec8(p.norm_w + c * 32 + 8 * g, w);
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              float x0 = __uint_as_float(v[8 * g + j]) + (p.bias ? b[j] : 0.f);
              float x1 = __uint_as_float(v[8 * g + j + 1]) + (p.bias ? b[j + 1] : 0.f);
              bf16_round2(x0, x1);
              x0 *= rstd;
              x1 *= rstd;
              bf16_round2(x0, x1);
              x0 *= w[j];
              x1 *= w[j + 1];
              bf16_round2(x0, x1);
              const float2 t = cs[4 * g + (j >> 1)];
              o[4 * g + (j >> 1)] = pack_bf16x2(x0 * t.x - x1 * t.y, x1 * t.x + x0 * t.y);
            }
          }
          if (valid) {
            uint4* dst = reinterpret_cast<uint4*>(out_ptr + n0 + c * 32);
#pragma unroll
            for (int t = 0; t < 4; ++t) dst[t] = make_uint4(o[4 * t], o[4 * t + 1], o[4 * t + 2], o[4 * t + 3]);
          }
        }
      }
      release();
    } else {
      int heads = (p.N - n_base) / 128;
      if (heads > bn / 128) heads = bn / 128;
      bool released = false;
#pragma unroll 1
      for (int h = 0; h < heads; h += kRopeHeads) {
        const bool last = h + kRopeHeads >= heads;
        if (kRopeHeads == 2 && h + 1 < heads)
          norm_rope_heads<kRopeHeads>(p, taddr + h * 128, out_ptr, n_base + h * 128, cs_base, cs_step, valid,
                             last ? &release : nullptr);
        else
          norm_rope_heads<1>(p, taddr + h * 128, out_ptr, n_base + h * 128, cs_base, cs_step, valid,
                             last ? &release : nullptr);
        released = released || last;
      }
      if (!released) release();
    }
  } else {
    // software pipeline over 32-column chunks: the TMEM load (and, for GATE_RES, the residual-row loads, whose L2 /
    // DRAM latency one thread per row cannot hide otherwise) of chunk c + 1 are in flight while chunk c is processed
    int cols = p.N - n_base;
    if (cols > bn) cols = bn;
    const int n_chunks = cols / 32;
    const bool tail16 = kTail16 && (cols & 16) != 0;   // CTA-pair kernel: widths 240, 208, 176, ...
    uint32_t va[32], vb[32];
    uint4 ra[4], rb[4];
    if constexpr (!kTail16) {
      if (n_chunks <= 0) { release(); return; }
      tmem_ld32(taddr, va);
      load_residual<EPI>(p, ra, m, n_base, valid);
    } else if (n_chunks > 0) {
      tmem_ld32(taddr, va);
      load_residual<EPI>(p, ra, m, n_base, valid);
    }
#pragma unroll 1
    for (int c = 0; c < n_chunks; c += 2) {
      tmem_ld_wait();
      if (c + 1 < n_chunks) {
        tmem_ld32(taddr + (c + 1) * 32, vb);
        load_residual<EPI>(p, rb, m, n_base + (c + 1) * 32, valid);
      }
      epilogue_chunk<EPI>(p, va, ra, out_ptr, n_base + c * 32, valid);
      if (c + 1 < n_chunks) {
        tmem_ld_wait();
        if (c + 2 < n_chunks) {
          tmem_ld32(taddr + (c + 2) * 32, va);
          load_residual<EPI>(p, ra, m, n_base + (c + 2) * 32, valid);
        }
        epilogue_chunk<EPI>(p, vb, rb, out_ptr, n_base + (c + 1) * 32, valid);
      }
    }
    if constexpr (kTail16) if (tail16) {
      tmem_ld16p(taddr + n_chunks * 32, va);
      load_residual<EPI, 2>(p, ra, m, n_base + n_chunks * 32, valid);
      tmem_ld_wait();
      epilogue_chunk<EPI, 2>(p, va, ra, out_ptr, n_base + n_chunks * 32, valid);
    }
    release();
  }
}

// Implemented in gemm2.cu: cluster-of-2 kernel (tcgen05 cta_group::2, 256 x 256 tiles). Returns
// cudaErrorNotSupported when the shape is outside its envelope so the caller can fall through to the 1-CTA kernel.
cudaError_t launch_gemm_2cta(const GemmArgs& a, int num_sms, cudaStream_t stream);

}  // namespace rge
