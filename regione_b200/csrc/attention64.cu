// tcgen05/TMEM flash attention, decoupled pipeline: active-Q x cached-KV, non-causal, head_dim 128.
//
// Same roles and the same ping-pong of two 128-row query tiles per CTA as attention.cu, with ONE structural change that
// its profile asked for (profiles/r02_ncu_attention_source_hotspots.txt: softmax warps wait for S 43 % of their time,
// tensor pipe and MUFU both 63 % busy - a serial  S ready -> softmax -> P ready -> P V -> Q K^T -> S ready  chain per
// tile with ~600 cycles of barrier / commit latency in it). There, P overwrites the score columns it was computed from,
// so the next Q K^T of a group cannot be issued before that group's P V has consumed P. Here the K/V tile is 64 rows,
// which leaves TMEM room for P in columns of its own:
//
//     S0 [0,64)  S1 [64,128)  P0 [128,160)  P1 [160,192)   O0 [256,384)  O1 [384,512)        (fp32 columns)
//
// so Q_i K_{j+1}^T is issued as soon as softmax group i has pulled S_i(j) into registers (`s_free`), S_i(j+1) is
// waiting when the group finishes tile j, and the group exponentiates tile after tile without waiting for the tensor
// pipe; P_i V_j only has to finish before P_i(j+1) is stored (`o_bar`, practically always long complete). The MMA
// thread issues whatever is ready (non-blocking mbarrier tests), so neither group's MMAs queue behind the other's
// barriers.
//
//   warp 0        TMA producer: Q once, then K_j / V_j 64-row tiles through a 10-slot shared-memory ring
//   warp 1        MMA issuer (one thread)
//   warps 2-3     idle (pad warpgroup 0 so that setmaxnreg can hand its registers to the softmax warpgroups)
//   warps 4-7     softmax group 0 (query tile 0), one thread per query row
//   warps 8-11    softmax group 1 (query tile 1)
//
// Replaces flash_attn_func(q, k, v, causal=False) at RegionE/FluxKontext/inplace.py:796-801 (see attention.cu).
#include "attention.cuh"
#include "ptx.cuh"
#include "tmap.cuh"
#include "tuning.cuh"

namespace rge {

namespace {

constexpr int kThreads = 384;
constexpr int kQTile = 128;                 // query rows per tile, head dim
constexpr int kKV = 64;                     // kv rows per tile
constexpr int kQHalfBytes = 128 * 128;      // one [128 rows x 64 bf16] swizzled half of a Q tile
constexpr int kQTileBytes = 2 * kQHalfBytes;
constexpr int kKVHalfBytes = kKV * 128;     // one [64 rows x 64 bf16] swizzled half of a K or V tile
constexpr int kKVTileBytes = 2 * kKVHalfBytes;
constexpr int kSlots = 10;
constexpr int kSmemBytes = 2 * kQTileBytes + kSlots * kKVTileBytes + 256 + 1024;
constexpr uint32_t kColS = 0, kColP = 128, kColO = 256;

struct Attn64Dev {
  __nv_bfloat16* O;
  long ldo;
  int Sq, Skv;
  float sl2;  // softmax scale * log2(e)
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// non-blocking phase test (mbarrier.test_wait): true once the phase with this parity has completed
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

__global__ void __launch_bounds__(kThreads, 1)
attention64_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const Attn64Dev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_q = smem;
  uint8_t* s_kv = smem + 2 * kQTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_kv + kSlots * kKVTileBytes);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + kSlots;
  uint64_t* s_full = kv_empty + kSlots;   // [group] S_i(j) = Q_i K_j^T is in TMEM
  uint64_t* s_free = s_full + 2;          // [group] the group has S_i(j) in registers: S_i may be overwritten
  uint64_t* p_full = s_free + 2;          // [group] P_i(j) is in TMEM
  uint64_t* o_bar = p_full + 2;           // [group] P_i(j) V_j has completed (O_i quiescent, P_i may be overwritten)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int q0 = blockIdx.x * 2 * kQTile;
  const int n_tiles = (p.Skv + kKV - 1) / kKV;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_bar[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    if (warp == 0) {
      // ------------------------------------------------------------ TMA producer (all lanes loop, one elected lane issues)
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, 2 * kQTileBytes);
        for (int i = 0; i < 2; ++i)
          for (int half = 0; half < 2; ++half)
            tma_load_2d(s_q + i * kQTileBytes + half * kQHalfBytes, &map_q, q_full, head * 128 + half * 64,
                        q0 + i * kQTile);
      }
      __syncwarp();
      for (int t = 0; t < 2 * n_tiles; ++t) {   // ring order K_0, V_0, K_1, V_1, ...
        const int slot = t % kSlots;
        const uint32_t ph = (t / kSlots) & 1;
        mbar_wait(&kv_empty[slot], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&kv_full[slot], kKVTileBytes);
          const CUtensorMap* map = (t & 1) ? &map_v : &map_k;
          for (int half = 0; half < 2; ++half)
            tma_load_2d(s_kv + slot * kKVTileBytes + half * kKVHalfBytes, map, &kv_full[slot], head * 128 + half * 64,
                        (t >> 1) * kKV);
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      // ------------------------------------------------------------ MMA issuer: whatever is ready, in tile order per group.
      // The WHOLE warp runs the (warp-uniform) control flow and one elected lane issues: inside a divergent
      // `if (lane == 0)` region ptxas wraps every tcgen05 instruction in an ELECT / BRA.U.ANY loop and routes its
      // operands through R2UR moves (~12 issue slots per MMA), which made this thread the bottleneck of the kernel.
      {
        constexpr uint32_t idesc_qk = make_idesc_bf16(128, kKV, 0, 0);
        constexpr uint32_t idesc_pv = make_idesc_bf16(128, 128, 0, 1);  // B (= V) is MN-major
        const uint32_t sq_addr = smem_u32(s_q);
        const uint32_t skv_addr = smem_u32(s_kv);
        const uint64_t q_desc[2] = {make_sdesc_sw128(sq_addr, 0, 1024), make_sdesc_sw128(sq_addr + kQTileBytes, 0, 1024)};
        const uint64_t k_desc0 = make_sdesc_sw128(skv_addr, 0, 1024);
        const uint64_t v_desc0 = make_sdesc_sw128(skv_addr, kKVHalfBytes, 1024);
        auto kv_ready = [&](int t) { return mbar_test(&kv_full[t % kSlots], (t / kSlots) & 1); };
        int qk_next[2] = {0, 0}, pv_next[2] = {0, 0};
        mbar_wait(q_full, 0);
        long long spin_t0 = clock64();
        while (pv_next[0] < n_tiles || pv_next[1] < n_tiles) {
          bool progressed = false;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            // S_i(j) = Q_i K_j^T: contraction over d, 8 steps of 16; both operands K-major, 128B-swizzled. Needs K_j in
            // shared memory and S_i(j-1) in the group's registers.
            int j = qk_next[i];
            if (j < n_tiles && (j == 0 || mbar_test(&s_free[i], (j - 1) & 1)) && kv_ready(2 * j)) {
              tc_fence_after();
              const uint64_t kd = k_desc0 + (uint64_t)((((2 * j) % kSlots) * kKVTileBytes) >> 4);
              const bool release = qk_next[i ^ 1] > j;   // both groups have used K_j
              if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                  umma_ss(tmem + kColS + i * kKV, q_desc[i] + (((kk >> 2) * kQHalfBytes + (kk & 3) * 32) >> 4),
                          kd + (((kk >> 2) * kKVHalfBytes + (kk & 3) * 32) >> 4), idesc_qk, kk != 0);
                }
                tc_commit(&s_full[i]);
                if (release) tc_commit(&kv_empty[(2 * j) % kSlots]);
              }
              __syncwarp();
              ++qk_next[i];
              progressed = true;
            }
            // O_i += P_i(j) V_j: contraction over kv, 4 steps of 16 rows (2048 B); V is [kv][d] = MN-major B with two
            // 64-wide d atoms kKVHalfBytes apart (LBO) and 8-row groups 1 KB apart (SBO); P from TMEM
            j = pv_next[i];
            if (j < n_tiles && mbar_test(&p_full[i], j & 1) && kv_ready(2 * j + 1)) {
              tc_fence_after();
              const uint64_t vd = v_desc0 + (uint64_t)((((2 * j + 1) % kSlots) * kKVTileBytes) >> 4);
              const bool release = pv_next[i ^ 1] > j;   // both groups have used V_j
              const uint32_t acc = j > 0;
              if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < kKV / 16; ++kk) {
                  umma_ts(tmem + kColO + i * 128, tmem + kColP + i * (kKV / 2) + kk * 8, vd + ((kk * 2048) >> 4), idesc_pv,
                          acc | (kk != 0));
                }
                tc_commit(&o_bar[i]);
                if (release) tc_commit(&kv_empty[(2 * j + 1) % kSlots]);
              }
              __syncwarp();
              ++pv_next[i];
              progressed = true;
            }
          }
          if (progressed) {
            spin_t0 = clock64();
          } else if (clock64() - spin_t0 > RGE_WAIT_TIMEOUT_CYCLES) {
            __trap();   // a lost arrive would otherwise hang the GPU
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    // ------------------------------------------------------------ softmax groups
    const int grp = (warp - 4) >> 2;  // query tile handled by this warp group
    const int qtr = warp & 3;         // TMEM lane quarter this warp may access
    const int row = q0 + grp * kQTile + qtr * 32 + lane;
    const uint32_t lane_base = uint32_t(qtr * 32) << 16;
    const uint32_t t_s = tmem + lane_base + kColS + grp * kKV;
    const uint32_t t_p = tmem + lane_base + kColP + grp * (kKV / 2);
    const uint32_t t_o = tmem + lane_base + kColO + grp * 128;
    const float sl2 = p.sl2;
    const uint64_t sl2_2 = pack2f(sl2, sl2);
    float m_run = -INFINITY, l_run = 0.f;

    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(&s_full[grp], j & 1);
      tc_fence_after();
      uint32_t v[kKV];
      tmem_ld32p(t_s, v);
      tmem_ld32p(t_s + 32, v + 32);
      tmem_ld_wait();
      // the scores are in registers: Q_i K_{j+1}^T may overwrite S_i while this tile is exponentiated
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[grp]);
      const int n_valid = min(kKV, p.Skv - j * kKV);
      if (n_valid < kKV) {  // KV tail (last tile only): masked columns behave as -inf
#pragma unroll
        for (int jj = 0; jj < kKV; ++jj)
          if (jj >= n_valid) v[jj] = 0xff800000u;
      }
      // row maximum of the raw scores, four independent FMNMX3 chains
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < kKV; jj += 8) {
        mx0 = max3f(mx0, __uint_as_float(v[jj + 0]), __uint_as_float(v[jj + 1]));
        mx1 = max3f(mx1, __uint_as_float(v[jj + 2]), __uint_as_float(v[jj + 3]));
        mx2 = max3f(mx2, __uint_as_float(v[jj + 4]), __uint_as_float(v[jj + 5]));
        mx3 = max3f(mx3, __uint_as_float(v[jj + 6]), __uint_as_float(v[jj + 7]));
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      if (j == 0) {
        m_run = mx;
      } else {
        // lazy reference maximum: rescale O (in TMEM, by this group) only when a row maximum grows by more than 2^8
        const bool need = (mx - m_run) * sl2 > 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          // O_i must be quiescent: P_i(j-1) V_{j-1} complete; P_i(j) V_j cannot be issued before our p_full arrive
          mbar_wait(&o_bar[grp], (j - 1) & 1);
          tc_fence_after();
          const float m_new = need ? mx : m_run;
          const float alpha = ex2f((m_run - m_new) * sl2);
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t o[32];
            tmem_ld32(t_o + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) o[jj] = __float_as_uint(__uint_as_float(o[jj]) * alpha);
            tmem_st32(t_o + c * 32, o);
          }
          tmem_st_wait();
          l_run *= alpha;
          m_run = m_new;
        }
      }
      // P = exp2((s - m) * scale * log2 e) as packed bf16 pairs; scale-and-shift and the row sum as packed FFMA2 / FADD2
      const float neg_m = -m_run * sl2;
      const uint64_t neg_m2 = pack2f(neg_m, neg_m);
      uint64_t sum_a = pack2f(0.f, 0.f), sum_b = sum_a;
#pragma unroll
      for (int jj = 0; jj < kKV; jj += 4) {
        float p0, p1, p2, p3;
        unpack2f(fma2(pack2u(v[jj + 0], v[jj + 1]), sl2_2, neg_m2), p0, p1);
        unpack2f(fma2(pack2u(v[jj + 2], v[jj + 3]), sl2_2, neg_m2), p2, p3);
        p0 = ex2f(p0);
        p1 = ex2f(p1);
        p2 = ex2f(p2);
        p3 = ex2f(p3);
        sum_a = add2(sum_a, pack2f(p0, p1));
        sum_b = add2(sum_b, pack2f(p2, p3));
        v[(jj >> 1) + 0] = pack_bf16x2(p0, p1);   // in place: slot jj/2 <= jj has already been consumed
        v[(jj >> 1) + 1] = pack_bf16x2(p2, p3);
      }
      // P_i's TMEM columns are free once P_i(j-1) V_{j-1} has completed (issued a whole tile ago: normally long done)
      if (j > 0) {
        mbar_wait(&o_bar[grp], (j - 1) & 1);
        tc_fence_after();
      }
      tmem_st32p(t_p, v);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[grp]);
      float s0, s1, s2, s3;
      unpack2f(sum_a, s0, s1);
      unpack2f(sum_b, s2, s3);
      l_run += (s0 + s1) + (s2 + s3);
    }
    // epilogue: O_i / l -> bf16 -> global
    mbar_wait(&o_bar[grp], (n_tiles - 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / l_run;
    const bool valid = row < p.Sq;
    __nv_bfloat16* orow = p.O + (long)(valid ? row : 0) * p.ldo + head * 128;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t o[32];
      tmem_ld32(t_o + c * 32, o);
      tmem_ld_wait();
      if (valid) {
        uint4* dst = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[8 * t + 0]) * inv, __uint_as_float(o[8 * t + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(o[8 * t + 2]) * inv, __uint_as_float(o[8 * t + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(o[8 * t + 4]) * inv, __uint_as_float(o[8 * t + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(o[8 * t + 6]) * inv, __uint_as_float(o[8 * t + 7]) * inv);
          dst[t] = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace

cudaError_t launch_attention64(const AttnArgs& a, cudaStream_t stream) {
  if (a.Sq <= 0 || a.H <= 0) return cudaSuccess;
  if (a.Skv <= 0 || (a.ldq % 8) || (a.ldk % 8) || (a.ldv % 8) || (a.ldo % 8)) return cudaErrorInvalidValue;
  const int dev = current_device();
  static bool attr_set[kMaxDevices] = {};   // per device
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(attention64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  CUtensorMap mq, mk, mv;
  if (!make_tmap_bf16_2d(&mq, a.Q, a.Sq, (uint64_t)a.H * 128, a.ldq, kQTile)) return cudaErrorInvalidValue;
  if (!make_tmap_bf16_2d(&mk, a.K, a.Skv, (uint64_t)a.H * 128, a.ldk, kKV)) return cudaErrorInvalidValue;
  if (!make_tmap_bf16_2d(&mv, a.V, a.Skv, (uint64_t)a.H * 128, a.ldv, kKV)) return cudaErrorInvalidValue;
  Attn64Dev p;
  p.O = a.O;
  p.ldo = a.ldo;
  p.Sq = a.Sq;
  p.Skv = a.Skv;
  p.sl2 = a.scale * 1.4426950408889634f;
  dim3 grid((a.Sq + 2 * kQTile - 1) / (2 * kQTile), a.H);
  attention64_kernel<<<grid, kThreads, kSmemBytes, stream>>>(mq, mk, mv, p);
  return cudaGetLastError();
}

}  // namespace rge
