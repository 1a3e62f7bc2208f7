// HBM-bound kernels: LayerNorm+adaLN modulation, batched GEMV (adaLN / time-text embedding), rotary table,
// row gather/scatter, Euler / velocity-decay reuse, adaptive region partition, morphology + compaction.
// Reference behaviour restated per kernel; see oracle/ for the CPU restatement the tests compare against.
#include "elementwise.cuh"
#include "ptx.cuh"
#include "tuning.cuh"

namespace rge {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  return u;
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }

// ------------------------------------------------------------------ LayerNorm + modulation
// diffusers AdaLayerNormZero / AdaLayerNormZeroSingle / AdaLayerNormContinuous body:
//   x = LayerNorm(x, eps=1e-6, no affine) * (1 + scale) + shift     (all bf16 tensors, fp32 statistics)
__global__ void __launch_bounds__(256) ln_modulate_kernel(const __nv_bfloat16* __restrict__ x, long ldx,
                                                          const __nv_bfloat16* __restrict__ scale,
                                                          const __nv_bfloat16* __restrict__ shift,
                                                          __nv_bfloat16* __restrict__ out, long ldo, int M, int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + warp;
  if (m >= M) return;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (long)m * ldx);
  const int nv = D >> 3;
  float s = 0.f;
  for (int i = lane; i < nv; i += 32) {
    float f[8];
    unpack8(xr[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += f[j];
  }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
  for (int i = lane; i < nv; i += 32) {
    float f[8];
    unpack8(xr[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float d = f[j] - mean;
      q += d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + 1e-6f);
  const uint4* sc = reinterpret_cast<const uint4*>(scale);
  const uint4* sh = reinterpret_cast<const uint4*>(shift);
  uint4* o = reinterpret_cast<uint4*>(out + (long)m * ldo);
  for (int i = lane; i < nv; i += 32) {
    float f[8], a[8], b[8];
    unpack8(xr[i], f);
    unpack8(__ldg(sc + i), a);
    unpack8(__ldg(sh + i), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float n = bf16_round((f[j] - mean) * rstd);
      float t = bf16_round(n * bf16_round(1.0f + a[j]));
      f[j] = t + b[j];
    }
    o[i] = pack8(f);
  }
}

// Same operation with the row held in registers (one global read): D = 256 * NV.
template <int NV>
__global__ void __launch_bounds__(256) ln_modulate_reg_kernel(const __nv_bfloat16* __restrict__ x, long ldx,
                                                              const __nv_bfloat16* __restrict__ scale,
                                                              const __nv_bfloat16* __restrict__ shift,
                                                              __nv_bfloat16* __restrict__ out, long ldo, int M) {
  constexpr int D = 256 * NV;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + warp;
  if (m >= M) return;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (long)m * ldx);
  uint4 r[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) r[i] = xr[lane + 32 * i];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float f[8];
    unpack8(r[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += f[j];
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float f[8];
    unpack8(r[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = f[j] - mean;
      q += d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + 1e-6f);
  const uint4* sc = reinterpret_cast<const uint4*>(scale);
  const uint4* sh = reinterpret_cast<const uint4*>(shift);
  uint4* o = reinterpret_cast<uint4*>(out + (long)m * ldo);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float f[8], a[8], b[8];
    unpack8(r[i], f);
    unpack8(__ldg(sc + lane + 32 * i), a);
    unpack8(__ldg(sh + lane + 32 * i), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float n = bf16_round((f[j] - mean) * rstd);
      const float t = bf16_round(n * bf16_round(1.0f + a[j]));
      f[j] = t + b[j];
    }
    o[lane + 32 * i] = pack8(f);
  }
}

// Wide rows (D = 1024 * NV, e.g. 3072): four warps share one row, two rows per 256-thread block. Each thread keeps NV
// 128-bit vectors (12 registers at NV = 3), the mean / variance cross the four warps through shared memory, and the
// block stays below 64 registers per thread so that 4+ blocks (32+ warps) are resident per SM: the one-warp-per-row
// variant above needs the whole row plus its scale / shift vectors in one thread's registers (255 at D = 3072, one
// block per SM) and ran at a third of the HBM rate.
template <int NV>
__global__ void __launch_bounds__(256, 4) ln_modulate_wide_kernel(const __nv_bfloat16* __restrict__ x, long ldx,
                                                                  const __nv_bfloat16* __restrict__ scale,
                                                                  const __nv_bfloat16* __restrict__ shift,
                                                                  __nv_bfloat16* __restrict__ out, long ldo, int M) {
  constexpr int D = 1024 * NV;
  __shared__ float red[2][2][4];                 // [pass][row in block][warp of the row]
  const int t = threadIdx.x & 127;               // thread within the row
  const int rb = threadIdx.x >> 7;               // row within the block
  const int w = (threadIdx.x >> 5) & 3;
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * 2 + rb;
  const bool live = m < M;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (long)(live ? m : 0) * ldx);
  uint4 r[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) r[i] = live ? xr[t + 128 * i] : make_uint4(0, 0, 0, 0);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float f[8];
    unpack8(r[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += f[j];
  }
  s = warp_sum(s);
  if (lane == 0) red[0][rb][w] = s;
  __syncthreads();
  const float mean = ((red[0][rb][0] + red[0][rb][1]) + (red[0][rb][2] + red[0][rb][3])) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float f[8];
    unpack8(r[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = f[j] - mean;
      q += d * d;
    }
  }
  q = warp_sum(q);
  if (lane == 0) red[1][rb][w] = q;
  __syncthreads();
  const float rstd = rsqrtf(((red[1][rb][0] + red[1][rb][1]) + (red[1][rb][2] + red[1][rb][3])) * (1.0f / D) + 1e-6f);
  if (!live) return;
  const uint4* sc = reinterpret_cast<const uint4*>(scale);
  const uint4* sh = reinterpret_cast<const uint4*>(shift);
  uint4* o = reinterpret_cast<uint4*>(out + (long)m * ldo);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float f[8], a[8], b[8];
    unpack8(r[i], f);
    unpack8(__ldg(sc + t + 128 * i), a);
    unpack8(__ldg(sh + t + 128 * i), b);
#pragma unroll
    for (int j = 0; j < 8; j += 2) {   // the reference's bf16 rounding points, two elements per packed conversion
      float n0 = (f[j] - mean) * rstd, n1 = (f[j + 1] - mean) * rstd;
      bf16_round2(n0, n1);
      float s0 = 1.0f + a[j], s1 = 1.0f + a[j + 1];
      bf16_round2(s0, s1);
      float t0 = n0 * s0, t1 = n1 * s1;
      bf16_round2(t0, t1);
      f[j] = t0 + b[j];
      f[j + 1] = t1 + b[j + 1];
    }
    o[t + 128 * i] = pack8(f);
  }
}

// ------------------------------------------------------------------ batched GEMV
// One block = 32 consecutive output rows of one job (8 warps x 4 rows): the input vector (with its optional SiLU) is
// staged once per block in shared memory and every lane keeps four independent 128-bit weight loads in flight. The
// per-row accumulation order (lane-strided, then a warp shuffle tree) does not depend on the rows-per-warp factor.
constexpr int kGemvRowsPerWarp = 4;
constexpr int kGemvRowsPerBlock = 8 * kGemvRowsPerWarp;

__global__ void __launch_bounds__(256) gemv_batch_kernel(const GemvJob* __restrict__ jobs) {
  extern __shared__ float xs[];
  const GemvJob job = jobs[blockIdx.y];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (blockIdx.x * kGemvRowsPerBlock >= job.N) return;
  for (int k = threadIdx.x; k < job.K; k += blockDim.x) {
    float v = __bfloat162float(job.x[k]);
    if (job.silu_in) v = bf16_round(silu_f(v));
    xs[k] = v;
  }
  __syncthreads();
  const int n0 = blockIdx.x * kGemvRowsPerBlock + warp * kGemvRowsPerWarp;
  if (n0 >= job.N) return;
  const int nv = job.K >> 3;
  const uint4* wr[kGemvRowsPerWarp];
  float acc[kGemvRowsPerWarp];
#pragma unroll
  for (int r = 0; r < kGemvRowsPerWarp; ++r) {
    wr[r] = reinterpret_cast<const uint4*>(job.W + (long)min(n0 + r, job.N - 1) * job.K);
    acc[r] = 0.f;
  }
  for (int i = lane; i < nv; i += 32) {
    uint4 w[kGemvRowsPerWarp];
#pragma unroll
    for (int r = 0; r < kGemvRowsPerWarp; ++r) w[r] = __ldg(wr[r] + i);
    const float* xv = xs + i * 8;
#pragma unroll
    for (int r = 0; r < kGemvRowsPerWarp; ++r) {
      float f[8];
      unpack8(w[r], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[r] = fmaf(f[j], xv[j], acc[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < kGemvRowsPerWarp; ++r) {
    const float a = warp_sum(acc[r]);
    const int n = n0 + r;
    if (lane == 0 && n < job.N) {
      float y = bf16_round(a + (job.b ? __bfloat162float(job.b[n]) : 0.f));
      if (job.silu_out) y = silu_f(y);
      job.out[n] = __float2bfloat16_rn(y);
    }
  }
}

// ------------------------------------------------------------------ timestep projection / small adds
__global__ void timestep_proj_kernel(float t, __nv_bfloat16* out) {
  const int i = threadIdx.x;  // 0..127
  const float exponent = (-9.210340371976184f * (float)i) / 128.0f;  // -ln(10000) * i / half_dim
  const float arg = t * expf(exponent);
  out[i] = __float2bfloat16_rn(cosf(arg));
  out[128 + i] = __float2bfloat16_rn(sinf(arg));
}
__global__ void add3_kernel(const __nv_bfloat16* a, const __nv_bfloat16* b, const __nv_bfloat16* c,
                            __nv_bfloat16* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = bf16_round(__bfloat162float(a[i]) + __bfloat162float(b[i]));
  if (c) s = s + __bfloat162float(c[i]);
  out[i] = __float2bfloat16_rn(s);
}

// ------------------------------------------------------------------ rotary table (FluxPosEmbed)
// ld == 0: cs is [S][64] row-major; ld > 0: pair-major [64][ld] (what the GEMM epilogue reads coalesced per warp)
__global__ void rope_table_kernel(const float* __restrict__ ids, float2* __restrict__ cs, int S, long ld) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= S * 64) return;
  const int s = ld > 0 ? idx % S : idx >> 6, p = ld > 0 ? idx / S : idx & 63;
  int axis, j, dim;
  if (p < 8) { axis = 0; j = p; dim = 16; }
  else if (p < 36) { axis = 1; j = p - 8; dim = 56; }
  else { axis = 2; j = p - 36; dim = 56; }
  const double freq = 1.0 / pow(10000.0, (double)(2 * j) / (double)dim);
  const double ang = (double)ids[s * 3 + axis] * freq;
  cs[ld > 0 ? (long)p * ld + s : (long)idx] = make_float2((float)cos(ang), (float)sin(ang));
}

// [S][64] row-major -> pair-major [64][ld] through a 32 x 32 shared-memory tile
__global__ void rope_transpose_kernel(const float2* __restrict__ src, float2* __restrict__ dst, int S, long ld) {
  __shared__ float2 tile[32][33];
  const int s0 = blockIdx.x * 32, p0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y)
    if (s0 + r < S) tile[r][threadIdx.x] = src[(long)(s0 + r) * 64 + p0 + threadIdx.x];
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y)
    if (s0 + threadIdx.x < S) dst[(long)(p0 + r) * ld + s0 + threadIdx.x] = tile[threadIdx.x][r];
}

__global__ void build_selection_kernel(const int* sel_img, int n_img, int T, int* sel_img_out, int* sel_all_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < T) sel_all_out[i] = i;
  if (i < n_img) {
    int v = sel_img ? sel_img[i] : i;
    sel_img_out[i] = v;
    sel_all_out[T + i] = T + v;
  }
}

// ------------------------------------------------------------------ row gather / scatter (utils.py:240-279)
template <bool kScatter>
__global__ void move_rows_kernel(const __nv_bfloat16* __restrict__ src, long lds, const int* __restrict__ ids, int n,
                                 int vec_per_row, __nv_bfloat16* __restrict__ dst, long ldd) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)n * vec_per_row) return;
  const int r = (int)(idx / vec_per_row), c = (int)(idx % vec_per_row);
  const int id = ids[r];
  if (kScatter)
    reinterpret_cast<uint4*>(dst + (long)id * ldd)[c] = reinterpret_cast<const uint4*>(src + (long)r * lds)[c];
  else
    reinterpret_cast<uint4*>(dst + (long)r * ldd)[c] = reinterpret_cast<const uint4*>(src + (long)id * lds)[c];
}

// ------------------------------------------------------------------ Euler step (+ velocity-decay reuse)
// inplace.py:610,655-680,686 and :318: fp32 sample, dt*v rounded to bf16 first (0-dim fp32 x bf16 tensor -> bf16).
__global__ void euler_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ v,
                             __nv_bfloat16* __restrict__ out, int M, int vec_per_row, float dt, float dt_direct,
                             const uint8_t* __restrict__ mask, int vscale_on, float vscale) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)M * vec_per_row) return;
  const int m = (int)(idx / vec_per_row);
  const float d = mask ? (mask[m] ? dt : dt_direct) : dt;
  float xf[8], vf[8];
  unpack8(reinterpret_cast<const uint4*>(x)[idx], xf);
  unpack8(reinterpret_cast<const uint4*>(v)[idx], vf);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float vv = vscale_on ? bf16_round(vf[j] * vscale) : vf[j];
    xf[j] = xf[j] + bf16_round(d * vv);
  }
  reinterpret_cast<uint4*>(out)[idx] = pack8(xf);
}

// ------------------------------------------------------------------ adaptive region partition (utils.py:305-333)
// kLanes = channels / 8 lanes own one token (one 128-bit load per tensor per lane, 32 / kLanes tokens per warp);
// reductions over the channel axis are shuffles inside the lane group.
template <int kLanes>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = kLanes / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int kLanes>
__global__ void __launch_bounds__(256) arp_similarity_kernel(const __nv_bfloat16* __restrict__ x,
                                                             const __nv_bfloat16* __restrict__ v,
                                                             const __nv_bfloat16* __restrict__ cond, float dt_final,
                                                             float thr, uint8_t* __restrict__ mask,
                                                             float* __restrict__ sim_out, int L) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;   // one (token, 8-channel vector) per thread
  const long m = idx / kLanes;
  const bool live = m < L;
  float e[8], c[8], xf[8], vf[8];
  if (live) {
    unpack8(reinterpret_cast<const uint4*>(x)[idx], xf);
    unpack8(reinterpret_cast<const uint4*>(v)[idx], vf);
    unpack8(reinterpret_cast<const uint4*>(cond)[idx], c);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) xf[j] = vf[j] = c[j] = 0.f;
  }
  float ss_e = 0.f, ss_c = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    // one-step x0 estimate, fp32: sample + bf16(dt_final * model_output)   (inplace.py:650)
    e[j] = xf[j] + bf16_round(dt_final * vf[j]);
    ss_e += e[j] * e[j];
    ss_c += c[j] * c[j];
  }
  ss_e = group_sum<kLanes>(ss_e);
  ss_c = group_sum<kLanes>(ss_c);
  // F.normalize: fp32 tensor in fp32; the bf16 condition latent in bf16 (norm rounded to bf16, quotient too)
  const float den_e = fmaxf(sqrtf(ss_e), 1e-12f);
  const float den_c = fmaxf(bf16_round(sqrtf(ss_c)), bf16_round(1e-12f));
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) dot += (e[j] / den_e) * bf16_round(c[j] / den_c);
  dot = group_sum<kLanes>(dot);
  if (live && (threadIdx.x % kLanes) == 0) {
    mask[m] = dot <= thr ? 1 : 0;
    if (sim_out) sim_out[m] = dot;
  }
}

// ------------------------------------------------------------------ morphology + ordered compaction
// utils.py:215-237 (erosion 3x3 cross == 5, dilation 5x5 ones > 0, zero padding) and :345-352 (ascending ids).
__global__ void __launch_bounds__(1024) morph_compact_kernel(const uint8_t* __restrict__ mask_in,
                                                             uint8_t* __restrict__ mask_out, int gh, int gw,
                                                             int erosion_dilation, int* __restrict__ edited,
                                                             int* __restrict__ unedited, int* __restrict__ counts) {
  extern __shared__ uint8_t sm[];
  const int L = gh * gw;
  uint8_t* a = sm;
  uint8_t* b = sm + L;
  __shared__ int warp_tot[32];
  for (int i = threadIdx.x; i < L; i += blockDim.x) a[i] = mask_in[i] ? 1 : 0;
  __syncthreads();
  if (erosion_dilation) {
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
      const int r = i / gw, c = i % gw;
      int s = a[i];
      s += (r > 0) ? a[i - gw] : 0;
      s += (r < gh - 1) ? a[i + gw] : 0;
      s += (c > 0) ? a[i - 1] : 0;
      s += (c < gw - 1) ? a[i + 1] : 0;
      b[i] = (s == 5) ? 1 : 0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
      const int r = i / gw, c = i % gw;
      int s = 0;
      for (int dr = -2; dr <= 2; ++dr) {
        const int rr = r + dr;
        if (rr < 0 || rr >= gh) continue;
        for (int dc = -2; dc <= 2; ++dc) {
          const int cc = c + dc;
          if (cc < 0 || cc >= gw) continue;
          s += b[rr * gw + cc];
        }
      }
      a[i] = (s > 0) ? 1 : 0;
    }
    __syncthreads();
  }
  // ordered compaction: thread t owns the contiguous chunk [t*chunk, (t+1)*chunk)
  const int chunk = (L + blockDim.x - 1) / blockDim.x;
  const int beg = threadIdx.x * chunk;
  const int end = min(beg + chunk, L);
  int cnt = 0;
  for (int i = beg; i < end; ++i) cnt += a[i];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = warp_tot[lane];
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    warp_tot[lane] = wi - w;  // exclusive
    if (lane == 31) counts[0] = wi;
    if (lane == 31) counts[1] = L - wi;
  }
  __syncthreads();
  int e_pos = warp_tot[warp] + incl - cnt;  // edited tokens before `beg`
  int u_pos = beg - e_pos;
  for (int i = beg; i < end; ++i) {
    if (mask_out) mask_out[i] = a[i];
    if (a[i]) edited[e_pos++] = i;
    else unedited[u_pos++] = i;
  }
}

// ------------------------------------------------------------------ row RMSNorm with weight (Qwen txt_norm)
// diffusers RMSNorm: fp32 variance, x * rsqrt(var + eps) in fp32, cast to the weight dtype, times weight.
__global__ void __launch_bounds__(256) rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, long ldx,
                                                      const __nv_bfloat16* __restrict__ w,
                                                      __nv_bfloat16* __restrict__ out, long ldo, int M, int D,
                                                      float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + warp;
  if (m >= M) return;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (long)m * ldx);
  const int nv = D >> 3;
  float q = 0.f;
  for (int i = lane; i < nv; i += 32) {
    float f[8];
    unpack8(xr[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) q += f[j] * f[j];
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
  const uint4* wr = reinterpret_cast<const uint4*>(w);
  uint4* o = reinterpret_cast<uint4*>(out + (long)m * ldo);
  for (int i = lane; i < nv; i += 32) {
    float f[8], a[8];
    unpack8(xr[i], f);
    unpack8(__ldg(wr + i), a);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = bf16_round(f[j] * rstd) * a[j];
    o[i] = pack8(f);
  }
}

// ------------------------------------------------------------------ norm-rescaled classifier-free guidance
// RegionE/QwenImageEdit/inplace.py:386-405: comb = neg + s * (pos - neg); out = comb * (||pos|| / ||comb||), every
// intermediate a bf16 tensor (row norms accumulate in fp32 and round to bf16). One warp per token.
template <int kLanes>
__global__ void __launch_bounds__(256) cfg_rescale_kernel(const __nv_bfloat16* __restrict__ pos,
                                                          const __nv_bfloat16* __restrict__ neg, float scale,
                                                          __nv_bfloat16* __restrict__ out, int M) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;   // one (token, 8-channel vector) per thread
  const bool live = idx / kLanes < M;
  float p[8], n[8], comb[8];
  if (live) {
    unpack8(reinterpret_cast<const uint4*>(pos)[idx], p);
    unpack8(reinterpret_cast<const uint4*>(neg)[idx], n);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) { p[j] = 0.f; n[j] = 0.f; }
  }
  float sp = 0.f, sc = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    comb[j] = bf16_round(n[j] + bf16_round(scale * bf16_round(p[j] - n[j])));
    sp += p[j] * p[j];
    sc += comb[j] * comb[j];
  }
  const float cond_norm = bf16_round(sqrtf(group_sum<kLanes>(sp)));
  const float noise_norm = bf16_round(sqrtf(group_sum<kLanes>(sc)));
  const float r = bf16_round(cond_norm / noise_norm);
#pragma unroll
  for (int j = 0; j < 8; ++j) comb[j] *= r;
  if (live) reinterpret_cast<uint4*>(out)[idx] = pack8(comb);
}

// ------------------------------------------------------------------ Step1X classifier-free guidance pieces
// RegionE/Step1XEdit/inplace.py:388-400: diff_norm = ||pos - neg|| per token (bf16), then
// out = neg + scale * (pos - neg) [/ denom], every intermediate a bf16 tensor. One warp per token.
template <int kLanes>
__global__ void __launch_bounds__(256) row_diff_norm_kernel(const __nv_bfloat16* __restrict__ pos,
                                                            const __nv_bfloat16* __restrict__ neg,
                                                            __nv_bfloat16* __restrict__ out, int M) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;   // one (token, 8-channel vector) per thread
  const long m = idx / kLanes;
  const bool live = m < M;
  float p[8], n[8];
  if (live) {
    unpack8(reinterpret_cast<const uint4*>(pos)[idx], p);
    unpack8(reinterpret_cast<const uint4*>(neg)[idx], n);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) { p[j] = 0.f; n[j] = 0.f; }
  }
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float d = bf16_round(p[j] - n[j]);
    ss += d * d;
  }
  ss = group_sum<kLanes>(ss);
  if (live && (threadIdx.x % kLanes) == 0) out[m] = __float2bfloat16_rn(sqrtf(ss));
}
__global__ void cfg_combine_kernel(const __nv_bfloat16* __restrict__ pos, const __nv_bfloat16* __restrict__ neg,
                                   float scale, const __nv_bfloat16* __restrict__ denom,
                                   __nv_bfloat16* __restrict__ out, int M, int vec_per_row) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;   // one 8-channel vector per thread
  if (idx >= (long)M * vec_per_row) return;
  float p[8], n[8];
  unpack8(reinterpret_cast<const uint4*>(pos)[idx], p);
  unpack8(reinterpret_cast<const uint4*>(neg)[idx], n);
  const float den = denom ? __bfloat162float(denom[idx / vec_per_row]) : 1.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float e = bf16_round(scale * bf16_round(p[j] - n[j]));
    if (denom) e = bf16_round(e / den);
    n[j] += e;
  }
  reinterpret_cast<uint4*>(out)[idx] = pack8(n);
}

inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// Latent pack / unpack either side of the loop (diffusers FluxKontextPipeline._pack_latents / _unpack_latents, called at
// RegionE/FluxKontext/inplace.py:212-226 through prepare_latents and at :398): [B, C, H, W] <-> [B, (H/2)(W/2), 4C]
// with packed channel = c*4 + dy*2 + dx. One thread moves one 2x2 patch of one channel (two 4-byte reads, one 8-byte
// write, or the reverse); threads run along x so the planar side is coalesced.
template <bool kUnpack>
__global__ void pack_latents_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int B, int C,
                                    int H2, int W2) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)B * C * H2 * W2;
  if (idx >= total) return;
  const int x = idx % W2;
  const int y = (idx / W2) % H2;
  const int c = (idx / ((long)W2 * H2)) % C;
  const int b = idx / ((long)W2 * H2 * C);
  const long planar = (((long)b * C + c) * (2 * H2) + 2 * y) * (2 * W2) + 2 * x;      // element (b, c, 2y, 2x)
  const long packed = (((long)b * H2 + y) * W2 + x) * (4 * C) + 4 * c;                 // token (y, x), channel 4c
  if (!kUnpack) {
    const uint32_t top = *reinterpret_cast<const uint32_t*>(src + planar);
    const uint32_t bot = *reinterpret_cast<const uint32_t*>(src + planar + 2 * W2);
    *reinterpret_cast<uint2*>(dst + packed) = make_uint2(top, bot);
  } else {
    const uint2 v = *reinterpret_cast<const uint2*>(src + packed);
    *reinterpret_cast<uint32_t*>(dst + planar) = v.x;
    *reinterpret_cast<uint32_t*>(dst + planar + 2 * W2) = v.y;
  }
}

}  // namespace

cudaError_t launch_ln_modulate(const __nv_bfloat16* x, long ldx, const __nv_bfloat16* scale,
                               const __nv_bfloat16* shift, __nv_bfloat16* out, long ldo, int M, int D,
                               cudaStream_t s) {
  if (M <= 0) return cudaSuccess;
  if (D % 8 || ldx % 8 || ldo % 8) return cudaErrorInvalidValue;
  if (D == 3072) ln_modulate_wide_kernel<3><<<cdiv(M, 2), 256, 0, s>>>(x, ldx, scale, shift, out, ldo, M);
  else if (D == 2048) ln_modulate_wide_kernel<2><<<cdiv(M, 2), 256, 0, s>>>(x, ldx, scale, shift, out, ldo, M);
  else if (D == 4096) ln_modulate_wide_kernel<4><<<cdiv(M, 2), 256, 0, s>>>(x, ldx, scale, shift, out, ldo, M);
  else if (D == 256) ln_modulate_reg_kernel<1><<<cdiv(M, 8), 256, 0, s>>>(x, ldx, scale, shift, out, ldo, M);
  else ln_modulate_kernel<<<cdiv(M, 8), 256, 0, s>>>(x, ldx, scale, shift, out, ldo, M, D);
  return cudaGetLastError();
}

cudaError_t launch_rmsnorm(const __nv_bfloat16* x, long ldx, const __nv_bfloat16* w, __nv_bfloat16* out, long ldo,
                           int M, int D, float eps, cudaStream_t s) {
  if (M <= 0) return cudaSuccess;
  if (D % 8 || ldx % 8 || ldo % 8) return cudaErrorInvalidValue;
  rmsnorm_kernel<<<cdiv(M, 8), 256, 0, s>>>(x, ldx, w, out, ldo, M, D, eps);
  return cudaGetLastError();
}

cudaError_t launch_cfg_rescale(const __nv_bfloat16* pos, const __nv_bfloat16* neg, float scale, __nv_bfloat16* out,
                               int M, int Cch, cudaStream_t s) {
  if (M <= 0) return cudaSuccess;
  const int grid = cdiv((long)M * (Cch / 8), 256);
  switch (Cch) {   // lanes per token = channels / 8 must be a power of two (the reference packs 64 channels)
    case 32: cfg_rescale_kernel<4><<<grid, 256, 0, s>>>(pos, neg, scale, out, M); break;
    case 64: cfg_rescale_kernel<8><<<grid, 256, 0, s>>>(pos, neg, scale, out, M); break;
    case 128: cfg_rescale_kernel<16><<<grid, 256, 0, s>>>(pos, neg, scale, out, M); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_row_diff_norm(const __nv_bfloat16* pos, const __nv_bfloat16* neg, __nv_bfloat16* out, int M,
                                 int Cch, cudaStream_t s) {
  if (M <= 0) return cudaSuccess;
  const int grid = cdiv((long)M * (Cch / 8), 256);
  switch (Cch) {
    case 32: row_diff_norm_kernel<4><<<grid, 256, 0, s>>>(pos, neg, out, M); break;
    case 64: row_diff_norm_kernel<8><<<grid, 256, 0, s>>>(pos, neg, out, M); break;
    case 128: row_diff_norm_kernel<16><<<grid, 256, 0, s>>>(pos, neg, out, M); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_cfg_combine(const __nv_bfloat16* pos, const __nv_bfloat16* neg, float scale,
                               const __nv_bfloat16* denom, __nv_bfloat16* out, int M, int Cch, cudaStream_t s) {
  if (M <= 0) return cudaSuccess;
  if (Cch % 8) return cudaErrorInvalidValue;
  cfg_combine_kernel<<<cdiv((long)M * (Cch / 8), 256), 256, 0, s>>>(pos, neg, scale, denom, out, M, Cch / 8);
  return cudaGetLastError();
}

cudaError_t launch_gemv_batch(const GemvJob* jobs_dev, int n_jobs, int max_n, cudaStream_t s) {
  if (n_jobs <= 0) return cudaSuccess;
  dim3 grid(cdiv(max_n, kGemvRowsPerBlock), n_jobs);
  gemv_batch_kernel<<<grid, 256, 4096 * sizeof(float), s>>>(jobs_dev);  // K <= 4096 (checked by the caller)
  return cudaGetLastError();
}

cudaError_t launch_timestep_proj(float t, __nv_bfloat16* out256, cudaStream_t s) {
  timestep_proj_kernel<<<1, 128, 0, s>>>(t, out256);
  return cudaGetLastError();
}

cudaError_t launch_add3(const __nv_bfloat16* a, const __nv_bfloat16* b, const __nv_bfloat16* c, __nv_bfloat16* out,
                        int n, cudaStream_t s) {
  add3_kernel<<<cdiv(n, 256), 256, 0, s>>>(a, b, c, out, n);
  return cudaGetLastError();
}

cudaError_t launch_rope_table(const float* ids, float2* cs, int S, long ld, cudaStream_t s) {
  if (S <= 0) return cudaSuccess;
  rope_table_kernel<<<cdiv((long)S * 64, 256), 256, 0, s>>>(ids, cs, S, ld);
  return cudaGetLastError();
}

cudaError_t launch_rope_transpose(const float2* src, float2* dst, int S, long ld, cudaStream_t s) {
  if (S <= 0) return cudaSuccess;
  rope_transpose_kernel<<<dim3(cdiv(S, 32), 2), dim3(32, 8), 0, s>>>(src, dst, S, ld);
  return cudaGetLastError();
}

cudaError_t launch_build_selection(const int* sel_img, int n_img, int T, int* sel_img_out, int* sel_all_out,
                                   cudaStream_t s) {
  const int n = n_img > T ? n_img : T;
  if (n <= 0) return cudaSuccess;
  build_selection_kernel<<<cdiv(n, 256), 256, 0, s>>>(sel_img, n_img, T, sel_img_out, sel_all_out);
  return cudaGetLastError();
}

cudaError_t launch_pack_latents(const __nv_bfloat16* src, __nv_bfloat16* dst, int B, int C, int H, int W, bool unpack,
                                cudaStream_t s) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return cudaSuccess;
  if ((H & 1) || (W & 1)) return cudaErrorInvalidValue;
  const long total = (long)B * C * (H / 2) * (W / 2);
  if (unpack) pack_latents_kernel<true><<<cdiv(total, 256), 256, 0, s>>>(src, dst, B, C, H / 2, W / 2);
  else pack_latents_kernel<false><<<cdiv(total, 256), 256, 0, s>>>(src, dst, B, C, H / 2, W / 2);
  return cudaGetLastError();
}

cudaError_t launch_gather_rows(const __nv_bfloat16* src, long lds, const int* ids, int n, int width,
                               __nv_bfloat16* dst, long ldd, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  if (width % 8 || lds % 8 || ldd % 8) return cudaErrorInvalidValue;
  move_rows_kernel<false><<<cdiv((long)n * (width / 8), 256), 256, 0, s>>>(src, lds, ids, n, width / 8, dst, ldd);
  return cudaGetLastError();
}
cudaError_t launch_scatter_rows(const __nv_bfloat16* src, long lds, const int* ids, int n, int width,
                                __nv_bfloat16* dst, long ldd, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  if (width % 8 || lds % 8 || ldd % 8) return cudaErrorInvalidValue;
  move_rows_kernel<true><<<cdiv((long)n * (width / 8), 256), 256, 0, s>>>(src, lds, ids, n, width / 8, dst, ldd);
  return cudaGetLastError();
}

cudaError_t launch_euler(const __nv_bfloat16* x, const __nv_bfloat16* v, __nv_bfloat16* out, int M, int Cch,
                         float dt, float dt_direct, const uint8_t* mask, int vscale_on, float vscale,
                         cudaStream_t s) {
  if (M <= 0) return cudaSuccess;
  if (Cch % 8) return cudaErrorInvalidValue;
  euler_kernel<<<cdiv((long)M * (Cch / 8), 256), 256, 0, s>>>(x, v, out, M, Cch / 8, dt, dt_direct, mask, vscale_on,
                                                             vscale);
  return cudaGetLastError();
}

cudaError_t launch_arp_similarity(const __nv_bfloat16* x, const __nv_bfloat16* v, const __nv_bfloat16* cond,
                                  float dt_final, float thr, uint8_t* mask, float* sim_out, int L, int Cch,
                                  cudaStream_t s) {
  if (L <= 0) return cudaSuccess;
  const int grid = cdiv((long)L * (Cch / 8), 256);
  switch (Cch) {   // lanes per token = channels / 8 must be a power of two (the reference packs 64 channels)
    case 32: arp_similarity_kernel<4><<<grid, 256, 0, s>>>(x, v, cond, dt_final, thr, mask, sim_out, L); break;
    case 64: arp_similarity_kernel<8><<<grid, 256, 0, s>>>(x, v, cond, dt_final, thr, mask, sim_out, L); break;
    case 128: arp_similarity_kernel<16><<<grid, 256, 0, s>>>(x, v, cond, dt_final, thr, mask, sim_out, L); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_morph_compact(const uint8_t* mask_in, uint8_t* mask_out, int gh, int gw, int erosion_dilation,
                                 int* edited, int* unedited, int* counts, cudaStream_t s) {
  const int L = gh * gw;
  if (L <= 0 || 2 * L > 96 * 1024) return cudaErrorInvalidValue;
  static bool attr_set[kMaxDevices] = {};   // per device
  const int dev = current_device();
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(morph_compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  morph_compact_kernel<<<1, 1024, 2 * L, s>>>(mask_in, mask_out, gh, gw, erosion_dilation, edited, unedited, counts);
  return cudaGetLastError();
}

}  // namespace rge
