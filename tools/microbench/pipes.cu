// Micro-benchmarks of the per-SM resources the attention kernel's softmax leg competes for (B200, sm_100a):
// TMEM read (tcgen05.ld 32x32b.x32), TMEM write (tcgen05.st), MUFU.EX2, packed FFMA2 - each with 4 and 8 warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I regione_b200/csrc -o /tmp/pipes tools/microbench/pipes.cu && /tmp/pipes
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace rge;

constexpr int kIters = 2000;

// mode 0: four x32 loads + one wait (a 128-column fp32 row per thread); mode 1: stores of the same; mode 2: ex2 on 128
// values; mode 3: 64 FFMA2; mode 4: loads and ex2 interleaved (software-pipelined: load chunk c+1 while exp of chunk c)
__global__ void __launch_bounds__(256, 1) bench(int mode, long long* cycles, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t = slot + (uint32_t((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t v[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) v[i] = __float_as_uint(-0.001f * (threadIdx.x + i));
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  if (mode == 0) {
    for (int it = 0; it < kIters; ++it) {
      tmem_ld32p(t, v); tmem_ld32p(t + 32, v + 32); tmem_ld32p(t + 64, v + 64); tmem_ld32p(t + 96, v + 96);
      tmem_ld_wait();
      acc += __uint_as_float(v[0] ^ v[37] ^ v[64] ^ v[127]);   // static indices: v stays in registers
    }
  } else if (mode == 1) {
    for (int it = 0; it < kIters; ++it) {
      v[0] = it;
      tmem_st32p(t, v); tmem_st32p(t + 32, v + 32); tmem_st32p(t + 64, v + 64); tmem_st32p(t + 96, v + 96);
      tmem_st_wait();
    }
  } else if (mode == 2) {
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
      for (int i = 0; i < 128; ++i) {
        float y;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(__uint_as_float(v[i])));
        v[i] = __float_as_uint(y - 1.0f);
      }
    }
  } else if (mode == 3) {
    uint64_t a = pack2f(1.0001f, 0.9999f), b = pack2f(0.5f, 0.25f);
    uint64_t x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = pack2u(v[2 * i], v[2 * i + 1]);
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = fma2(x[i], a, b);
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) { float lo, hi; unpack2f(x[i], lo, hi); acc += lo + hi; }
  } else if (mode == 4) {
    uint32_t w[32];
    float accs[4] = {0.f, 0.f, 0.f, 0.f};
    tmem_ld32p(t, v);
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld_wait();
        if (c & 1) tmem_ld32p(t + ((c + 1) & 3) * 32, v); else tmem_ld32p(t + ((c + 1) & 3) * 32, w);
        uint32_t* cur = (c & 1) ? w : v;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float y;
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(__uint_as_float(cur[i]) * 1e-30f));
          accs[i & 3] += y;
        }
      }
    }
    tmem_ld_wait();
    acc += accs[0] + accs[1] + accs[2] + accs[3];
  }
  const long long t1 = clock64();
#pragma unroll
  for (int i = 0; i < 128; ++i) acc += __uint_as_float(v[i]);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

int main() {
  long long* cyc;
  float* sink;
  cudaMalloc(&cyc, 148 * sizeof(long long));
  cudaMalloc(&sink, 148 * 256 * sizeof(float));
  const char* names[] = {"tcgen05.ld 4 x (32x32b.x32) + wait   [64 KB per 4 warps]", "tcgen05.st 4 x (32x32b.x32) + wait",
                         "MUFU.EX2 x 128 per thread", "FFMA2 x 64 per thread", "ld(32 cols) pipelined with 32 x EX2, x 4"};
  for (int mode = 0; mode < 5; ++mode)
    for (int threads = 128; threads <= 256; threads += 128) {
      bench<<<148, threads>>>(mode, cyc, sink);
      bench<<<148, threads>>>(mode, cyc, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
      long long h[148], mx = 0;
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      for (int i = 0; i < 148; ++i) if (h[i] > mx) mx = h[i];
      const double per_iter = (double)mx / kIters;
      printf("%-62s %d warps/SM: %8.1f cycles per iteration", names[mode], threads / 32, per_iter);
      if (mode <= 1 || mode == 4) printf("  -> %.0f B/cycle/SM of TMEM traffic", threads * 128.0 * 4 / per_iter);
      if (mode == 2 || mode == 4) printf("  -> %.1f ex2/cycle/SM", threads * 128.0 / per_iter);
      if (mode == 3) printf("  -> %.1f fp32 FMA/cycle/SM", threads * 128.0 / per_iter);
      printf("\n");
    }
  return 0;
}
