// tcgen05/TMEM flash attention for the RegionE hot path: active-Q x cached-KV, non-causal, head_dim 128.
//
// One CTA owns 256 query rows (two 128-row tiles) of one head and streams the whole K/V cache of that head:
//   warp 0        TMA producer: Q once, then K_j / V_j tiles through a 5-slot shared-memory ring
//   warp 1        MMA issuer (one thread): S_i = Q_i K_j^T (SS), O_i += P_i V_j (P from TMEM, V MN-major from smem)
//   warps 2-3     idle (pad warpgroup 0 so that setmaxnreg can hand its registers to the softmax warpgroups)
//   warps 4-7     softmax group 0 (query tile 0), one thread per query row, 208 registers
//   warps 8-11    softmax group 1 (query tile 1)
// TMEM (512 columns): S0 | S1 | O0 | O1, 128 fp32 columns each; P_i (bf16 pairs) overwrites the first 64 columns
// of S_i once the scores they replace are in registers. The two tiles ping-pong so the tensor pipe works on one while
// the other's softmax runs. Online softmax uses a lazy reference maximum: O is only rescaled (by the softmax group
// itself, in TMEM) when a row maximum grows by more than 2^8.
//
// Measured and rejected in round 2 (profiles/r02_attn_bench_pipe_poly_variants.log, r02_ncu_attention_source_hotspots.txt):
//   * evaluating 4 of every 8 exponential pairs as a packed-FFMA2 polynomial instead of MUFU.EX2 (`attn_poly` knob):
//     slower than 2 or 3 of 8 (3 is the default since the warp-uniform issue loops: 813 vs 829 us at 8704^2, 222 vs
//     237 us at 1576 x 8704, profiles/r02_attn_bench_early_qk_variants.log) - beyond that the extra ~4 issue slots
//     per offloaded score cost the single softmax warp per scheduler more than the MUFU time saves;
//   * S = Q K^T and the softmax in two software-pipelined 64-column halves (own commit per half, half 1's TMEM loads in
//     flight during half 0's exponentials): 17 % slower - TMEM reads are not a bottleneck (measured 720-920 B/clk/SM,
//     profiles/r02_pipes_microbench.log) and the second barrier round trip per tile lengthens the serial
//     S-ready -> softmax -> P-ready -> MMA chain that bounds this design (41 % of softmax-warp samples wait for S).
//   * issuing the upper 64 score columns of the next tile (which P does not alias) before this tile's P exists, and
//     publishing P in 64 | 32 | 32 pieces (profiles/r02_attn_bench_early_qk_variants.log, source under
//     profiles/rejected/): 3-6 % slower - an N = 64 SS MMA reads 6 KB of operands per 32 tensor cycles (192 B/clk,
//     above the 128 B/clk of shared memory), so splitting Q K^T costs more tensor time than the shorter chain saves.
//
// Replaces flash_attn_func(q, k, v, causal=False) at RegionE/FluxKontext/inplace.py:796-801; K/V are read in place
// from the Region-Instruction KV cache instead of being re-normalised, re-rotated and re-concatenated every step
// (inplace.py:756-794).
#include <cstdlib>

#include "attention.cuh"
#include "ptx.cuh"
#include "tmap.cuh"
#include "tuning.cuh"

namespace rge {

namespace {

constexpr int kThreads = 384;
constexpr int kTile = 128;             // query rows per tile, kv rows per tile, head dim
constexpr int kHalfBytes = 128 * 128;  // one [128 rows x 64 bf16] swizzled half tile
constexpr int kTileBytes = 2 * kHalfBytes;
constexpr int kSlots = 5;
constexpr int kSmemBytes = 2 * kTileBytes + kSlots * kTileBytes + 256 + 1024;
constexpr int kDefaultKernel = 0;       // 0: this file, 1: attention64.cu (RGE_ATTN_KERNEL overrides)
constexpr int kDefaultPoly = 3;         // exponential pairs of every 8 on the FMA pipe (RGE_ATTN_POLY overrides)
constexpr uint32_t kColS = 0, kColO = 256;  // TMEM column bases: S_i at kColS + 128 i, O_i at kColO + 128 i

struct AttnDev {
  __nv_bfloat16* O;
  long ldo;
  int Sq, Skv;
  float sl2;  // softmax scale * log2(e)
  // Work units = (256-row query tile, head): the n_full_x full tiles of every head first (head-major), then the ragged
  // last tile of every head (Sq % 256 rows, if any). 1-D grid: CTAs [0, n_whole) run one unit over the whole K/V
  // sequence; the remaining units - the ones that would form a mostly empty last wave of CTAs - are cut along K/V:
  // CTA n_whole + u * n_split + s streams K/V tiles [s * tiles_per_split, ...) of unit n_whole + u and leaves
  // un-normalised fp32 partials (O, reference maximum, row sum) in `ws` for attention_combine_kernel.
  int H, n_full_x, n_whole, n_split, tiles_per_split;
  float* ws;  // [split unit][n_split][256 rows][kWsRow]
};

// (query tile index within the head, head) of a work unit
__device__ __forceinline__ void unit_decode(int unit, int n_full_x, int H, int& qx, int& head) {
  const int n_full = n_full_x * H;
  if (unit < n_full) {
    head = unit / n_full_x;
    qx = unit - head * n_full_x;
  } else {
    head = unit - n_full;
    qx = n_full_x;
  }
}
constexpr int kMaxSplit = 8;
constexpr int kWsRow = 128 + 4;   // floats per partial row (16-byte multiple): O[128], reference maximum (raw score units), row sum, pad

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for a PAIR of arguments on the FMA / ALU pipes: clamp, Cody-Waite split with the round-to-nearest magic constant,
// degree-3 minimax polynomial on [-0.5, 0.5] (relative error 7.5e-5, far below the bf16 rounding of P) evaluated with
// packed FFMA2, exponent re-inserted with one shift-add per element. The MUFU pipe (16 ex2 per clock per SM) needs
// exactly as long for a 128 x 128 score tile as the tensor pipe needs for its two MMAs, so every exponential moved here
// shortens the softmax leg of the ping-pong below the MMA leg.
__device__ __forceinline__ void ex2_poly2(uint64_t x, float& p0, float& p1) {
  float x0, x1;
  unpack2f(x, x0, x1);
  x0 = fmaxf(x0, -125.0f);
  x1 = fmaxf(x1, -125.0f);
  const uint64_t xc = pack2f(x0, x1);
  const uint64_t magic = pack2f(12582912.0f, 12582912.0f);   // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const uint64_t r = add2(xc, magic);
  const uint64_t f = sub2(xc, sub2(r, magic));               // fraction in [-0.5, 0.5]
  uint64_t p = fma2(pack2f(0.055171459913253784f, 0.055171459913253784f), f,
                    pack2f(0.2426108568906784f, 0.2426108568906784f));
  p = fma2(p, f, pack2f(0.6932609677314758f, 0.6932609677314758f));
  p = fma2(p, f, pack2f(0.9999281167984009f, 0.9999281167984009f));
  float r0, r1, q0, q1;
  unpack2f(r, r0, r1);
  unpack2f(p, q0, q1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(r0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(r1) << 23));
}

// which of the 8 pairs of a 16-score group take the polynomial path, for kPoly = 0, 2, 3, 4 pairs of 8
__device__ __forceinline__ constexpr bool poly_pair(int kPoly, int pair) {
  return kPoly == 4 ? (pair & 1) : kPoly == 3 ? (pair == 2 || pair == 5 || pair == 7)
       : kPoly == 2 ? (pair == 3 || pair == 7) : false;
}

// kPoly = how many of every eight score PAIRS use ex2_poly2 instead of MUFU.EX2 (0, 2, 3 or 4).
template <int kPoly>
__global__ void __launch_bounds__(kThreads, 1)
attention_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                 const __grid_constant__ CUtensorMap map_v, const AttnDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_q = smem;
  uint8_t* s_kv = smem + 2 * kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_kv + kSlots * kTileBytes);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + kSlots;
  uint64_t* s_full = kv_empty + kSlots;
  uint64_t* p_half = s_full + 2;   // [group][half]: P columns [64 half, 64 half + 64) of the group are in TMEM
  uint64_t* o_bar = p_half + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int bid = blockIdx.x;
  const bool split = bid >= p.n_whole;
  const int s_unit = split ? (bid - p.n_whole) / p.n_split : 0;                 // index among the split units
  const int part = split ? (bid - p.n_whole) - s_unit * p.n_split : 0;
  int qx, head;
  unit_decode(split ? p.n_whole + s_unit : bid, p.n_full_x, p.H, qx, head);
  const int q0 = qx * 2 * kTile;
  const int n_tiles_all = (p.Skv + kTile - 1) / kTile;
  const int tile0 = split ? part * p.tiles_per_split : 0;                       // first K/V tile of this CTA
  const int n_tiles = split ? min(p.tiles_per_split, n_tiles_all - tile0) : n_tiles_all;
  const bool two = q0 + kTile < p.Sq;   // the second query tile has rows: otherwise its softmax group and MMAs are skipped

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
      mbar_init(&p_half[2 * i], 4);
      mbar_init(&p_half[2 * i + 1], 4);
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

  // register re-distribution: warpgroup 0 (TMA / MMA / idle) keeps 88, the softmax warpgroups take 208 each
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (all lanes loop, one elected lane issues)
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
      for (int i = 0; i < 2; ++i)
        for (int half = 0; half < 2; ++half)
          tma_load_2d(s_q + i * kTileBytes + half * kHalfBytes, &map_q, q_full, head * 128 + half * 64,
                      q0 + i * kTile);
    }
    __syncwarp();
    for (int t = 0; t < 2 * n_tiles; ++t) {
      const int slot = t % kSlots;
      const uint32_t ph = (t / kSlots) & 1;
      mbar_wait(&kv_empty[slot], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&kv_full[slot], kTileBytes);
        const CUtensorMap* map = (t & 1) ? &map_v : &map_k;
        for (int half = 0; half < 2; ++half)
          tma_load_2d(s_kv + slot * kTileBytes + half * kHalfBytes, map, &kv_full[slot], head * 128 + half * 64,
                      (tile0 + (t >> 1)) * kTile);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer. The WHOLE warp runs the (warp-uniform)
    // control flow and one elected lane issues: inside a divergent `if (lane == 0)` region ptxas wraps every tcgen05
    // instruction in an ELECT / BRA.U.ANY loop and routes its operands through R2UR moves (~12 issue slots per MMA) -
    // ~400 issue slots per K/V tile sitting in the serial  P ready -> P V -> Q K^T -> S ready  chain.
    constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, 128, 0, 1);  // B (= V) is MN-major
    const uint32_t sq_addr = smem_u32(s_q);
    const uint32_t skv_addr = smem_u32(s_kv);
    const uint64_t q_desc[2] = {make_sdesc_sw128(sq_addr, 0, 1024), make_sdesc_sw128(sq_addr + kTileBytes, 0, 1024)};
    const uint64_t k_desc0 = make_sdesc_sw128(skv_addr, 0, 1024);            // + (slot offset >> 4)
    const uint64_t v_desc0 = make_sdesc_sw128(skv_addr, kHalfBytes, 1024);   // V: MN-major, d atoms 16 KB apart
    auto slot_off = [&](int t) { return (uint64_t)(((t % kSlots) * kTileBytes) >> 4); };
    auto wait_kv = [&](int t) {
      mbar_wait(&kv_full[t % kSlots], (t / kSlots) & 1);
      tc_fence_after();
    };
    // S_i = Q_i K^T : contraction over d, 8 steps of 16; both operands K-major, 128B-swizzled
    auto issue_qk = [&](int i, int t, uint64_t* release) {
      if (elect_one()) {
        const uint64_t kd = k_desc0 + slot_off(t);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = ((kk >> 2) * kHalfBytes + (kk & 3) * 32) >> 4;
          umma_ss(tmem + kColS + i * 128, q_desc[i] + off, kd + off, idesc_qk, kk != 0);
        }
        tc_commit(&s_full[i]);
        if (release) tc_commit(release);
      }
      __syncwarp();
    };
    // O_i += P_i V : contraction over kv, 8 steps of 16 rows (2048 B); V is [kv][d] = MN-major B with two
    // 64-wide d atoms 16 KB apart (LBO) and 8-row groups 1 KB apart (SBO)
    // issued in two halves of 64 kv rows: the first starts while the softmax group still exponentiates the second
    auto issue_pv = [&](int i, int t, uint32_t accumulate, int j, uint64_t* release) {
      const uint64_t vd = v_desc0 + slot_off(t);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        mbar_wait(&p_half[2 * i + half], j & 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int kk = 4 * half; kk < 4 * half + 4; ++kk) {
            umma_ts(tmem + kColO + i * 128, tmem + kColS + i * 128 + kk * 8, vd + ((kk * 2048) >> 4), idesc_pv,
                    accumulate | (kk != 0));
          }
          if (half == 1) {
            tc_commit(&o_bar[i]);
            if (release) tc_commit(release);
          }
        }
        __syncwarp();
      }
    };
    // A K/V slot is handed back by the last MMA that reads it: group 1's, or group 0's when the CTA has one query tile.
    mbar_wait(q_full, 0);
    wait_kv(0);
    issue_qk(0, 0, two ? nullptr : &kv_empty[0]);
    if (two) issue_qk(1, 0, &kv_empty[0]);
    for (int j = 0; j < n_tiles; ++j) {
      const int tv = 2 * j + 1, tk = 2 * j + 2;
      const bool more = j + 1 < n_tiles;
      wait_kv(tv);
      issue_pv(0, tv, j > 0, j, two ? nullptr : &kv_empty[tv % kSlots]);
      if (more) {
        wait_kv(tk);
        issue_qk(0, tk, two ? nullptr : &kv_empty[tk % kSlots]);
      }
      if (two) {
        issue_pv(1, tv, j > 0, j, &kv_empty[tv % kSlots]);
        if (more) issue_qk(1, tk, &kv_empty[tk % kSlots]);
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    // ------------------------------------------------------------ softmax groups
    const int grp = (warp - 4) >> 2;  // query tile handled by this warp group
    const int qtr = warp & 3;         // TMEM lane quarter this warp may access
    const int row = q0 + grp * kTile + qtr * 32 + lane;
    const uint32_t lane_base = uint32_t(qtr * 32) << 16;
    const uint32_t t_s = tmem + lane_base + kColS + grp * 128;
    const uint32_t t_o = tmem + lane_base + kColO + grp * 128;
    const float sl2 = p.sl2;
    float m_run = -INFINITY, l_run = 0.f;

    const int my_tiles = (grp == 0 || two) ? n_tiles : 0;   // a group without query rows has nothing to do
    for (int j = 0; j < my_tiles; ++j) {
      mbar_wait(&s_full[grp], j & 1);
      tc_fence_after();
      const int n_valid = min(kTile, p.Skv - (tile0 + j) * kTile);
      // the whole score row in registers: four 32-column TMEM loads in flight, one wait
      uint32_t v[128];
      tmem_ld32p(t_s, v);
      tmem_ld32p(t_s + 32, v + 32);
      tmem_ld32p(t_s + 64, v + 64);
      tmem_ld32p(t_s + 96, v + 96);
      tmem_ld_wait();
      if (n_valid < kTile) {  // KV tail (last tile only): masked columns behave as -inf
#pragma unroll
        for (int jj = 0; jj < 128; ++jj)
          if (jj >= n_valid) v[jj] = 0xff800000u;
      }
      // row maximum of the raw scores, four independent chains
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < 128; jj += 8) {   // FMNMX3: one instruction per two scores
        mx0 = max3f(mx0, __uint_as_float(v[jj + 0]), __uint_as_float(v[jj + 1]));
        mx1 = max3f(mx1, __uint_as_float(v[jj + 2]), __uint_as_float(v[jj + 3]));
        mx2 = max3f(mx2, __uint_as_float(v[jj + 4]), __uint_as_float(v[jj + 5]));
        mx3 = max3f(mx3, __uint_as_float(v[jj + 6]), __uint_as_float(v[jj + 7]));
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      if (j == 0) {
        m_run = mx;
      } else {
        const bool need = (mx - m_run) * sl2 > 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          // O_i must be quiescent: PV_i(j-1) complete, PV_i(j) not issued before our p_full arrive
          mbar_wait(&o_bar[grp], (j - 1) & 1);
          tc_fence_after();
          const float m_new = need ? mx : m_run;
          const float alpha = ex2((m_run - m_new) * sl2);
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
      // pass 2: P = exp2((s - m) * scale * log2 e), packed bf16 pairs into the first 64 columns of S_i. Scale-and-shift
      // and the row sum run as packed FFMA2 / FADD2 (one issue slot per two scores).
      const uint64_t sl2_2 = pack2f(sl2, sl2);
      const float neg_m = -m_run * sl2;
      const uint64_t neg_m2 = pack2f(neg_m, neg_m);
      uint64_t sum_a = pack2f(0.f, 0.f), sum_b = sum_a;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int jj = 64 * half; jj < 64 * half + 64; jj += 4) {
          const uint64_t xa = fma2(pack2u(v[jj + 0], v[jj + 1]), sl2_2, neg_m2);
          const uint64_t xb = fma2(pack2u(v[jj + 2], v[jj + 3]), sl2_2, neg_m2);
          float p0, p1, p2, p3;
          if (poly_pair(kPoly, (jj >> 1) & 7)) {
            ex2_poly2(xa, p0, p1);
          } else {
            unpack2f(xa, p0, p1);
            p0 = ex2(p0);
            p1 = ex2(p1);
          }
          if (poly_pair(kPoly, ((jj >> 1) + 1) & 7)) {
            ex2_poly2(xb, p2, p3);
          } else {
            unpack2f(xb, p2, p3);
            p2 = ex2(p2);
            p3 = ex2(p3);
          }
          sum_a = add2(sum_a, pack2f(p0, p1));
          sum_b = add2(sum_b, pack2f(p2, p3));
          v[(jj >> 1) + 0] = pack_bf16x2(p0, p1);   // in place: slot jj/2 <= jj has already been consumed
          v[(jj >> 1) + 1] = pack_bf16x2(p2, p3);
        }
        // publish this half of P (32 TMEM columns = 64 kv positions) so its PV MMAs can start
        tmem_st32p(t_s + 32 * half, v + 32 * half);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_half[2 * grp + half]);
      }
      float sum0, sum1, sum2, sum3;
      unpack2f(sum_a, sum0, sum1);
      unpack2f(sum_b, sum2, sum3);
      l_run += (sum0 + sum1) + (sum2 + sum3);
    }
    if (my_tiles > 0) {
    mbar_wait(&o_bar[grp], (n_tiles - 1) & 1);
    tc_fence_after();
    const bool valid = row < p.Sq;
    if (split) {
      // K/V-split CTA: un-normalised fp32 partials; attention_combine_kernel merges the n_split parts of a row
      float* wrow = p.ws + ((size_t)(s_unit * p.n_split + part) * 2 * kTile + (row - q0)) * kWsRow;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(t_o + c * 32, v);
        tmem_ld_wait();
        if (valid) {
          uint4* dst = reinterpret_cast<uint4*>(wrow + c * 32);
#pragma unroll
          for (int t = 0; t < 8; ++t) dst[t] = make_uint4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
        }
      }
      if (valid) {
        wrow[128] = m_run;
        wrow[129] = l_run;
      }
    } else {
    // epilogue: O_i / l -> bf16 -> global
    const float inv = 1.0f / l_run;
    __nv_bfloat16* orow = p.O + (long)(valid ? row : 0) * p.ldo + head * 128;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld32(t_o + c * 32, v);
      tmem_ld_wait();
      if (valid) {
        uint4* dst = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(v[8 * t + 0]) * inv, __uint_as_float(v[8 * t + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(v[8 * t + 2]) * inv, __uint_as_float(v[8 * t + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(v[8 * t + 4]) * inv, __uint_as_float(v[8 * t + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(v[8 * t + 6]) * inv, __uint_as_float(v[8 * t + 7]) * inv);
          dst[t] = u;
        }
      }
    }
    }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// Merges the K/V-split partials: one block per (row of the tile, split unit), one thread per head-dim column.
// O = sum_s w_s O_s / sum_s w_s l_s with w_s = 2^((m_s - max m) * scale * log2 e).
__global__ void __launch_bounds__(128)
attention_combine_kernel(const float* __restrict__ ws, __nv_bfloat16* __restrict__ O, long ldo, int Sq, int H,
                         int n_full_x, int n_whole, int n_split, float sl2) {
  const int r = blockIdx.x, s_unit = blockIdx.y, d = threadIdx.x;
  int qx, head;
  unit_decode(n_whole + s_unit, n_full_x, H, qx, head);
  const int row = qx * 2 * kTile + r;
  if (row >= Sq) return;
  const float* base = ws + ((size_t)s_unit * n_split * 2 * kTile + r) * kWsRow;
  const size_t stride = (size_t)2 * kTile * kWsRow;
  float m = -INFINITY;
  for (int s = 0; s < n_split; ++s) m = fmaxf(m, base[s * stride + 128]);
  float acc = 0.f, l = 0.f;
  for (int s = 0; s < n_split; ++s) {
    const float w = exp2f((base[s * stride + 128] - m) * sl2);
    acc = fmaf(w, base[s * stride + d], acc);
    l = fmaf(w, base[s * stride + 129], l);
  }
  O[(long)row * ldo + head * 128 + d] = __float2bfloat16_rn(acc / l);
}

}  // namespace

cudaError_t launch_attention128(const AttnArgs& a, cudaStream_t stream) {
  if (a.Sq <= 0 || a.H <= 0) return cudaSuccess;
  if (a.Skv <= 0 || (a.ldq % 8) || (a.ldk % 8) || (a.ldv % 8) || (a.ldo % 8)) return cudaErrorInvalidValue;
  // tuning knob attn_poly / RGE_ATTN_POLY: 0, 2, 3 (default) or 4 of every 8 exponential pairs on the FMA pipe instead
  // of MUFU (2 and 3 are within 1 % of each other, 3 ahead on the slower boxes: profiles/r02_attn_bench_final.log)
  int poly = tuning().attn_poly;
  if (poly < 0) poly = kDefaultPoly;
  if (poly != 2 && poly != 3 && poly != 4) poly = 0;
  typedef void (*Kernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttnDev);
  static const Kernel table[4] = {attention_kernel<0>, attention_kernel<2>, attention_kernel<3>, attention_kernel<4>};
  const int dev = current_device();
  static bool attr_set[kMaxDevices] = {};   // per device: a process may drive several GPUs
  if (!attr_set[dev]) {
    for (int k = 0; k < 4; ++k) {
      cudaError_t e = cudaFuncSetAttribute(table[k], cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
      if (e != cudaSuccess) return e;
    }
    attr_set[dev] = true;
  }
  CUtensorMap mq, mk, mv;
  if (!make_tmap_bf16_2d(&mq, a.Q, a.Sq, (uint64_t)a.H * 128, a.ldq, kTile)) return cudaErrorInvalidValue;
  if (!make_tmap_bf16_2d(&mk, a.K, a.Skv, (uint64_t)a.H * 128, a.ldk, kTile)) return cudaErrorInvalidValue;
  if (!make_tmap_bf16_2d(&mv, a.V, a.Skv, (uint64_t)a.H * 128, a.ldv, kTile)) return cudaErrorInvalidValue;
  AttnDev p;
  p.O = a.O;
  p.ldo = a.ldo;
  p.Sq = a.Sq;
  p.Skv = a.Skv;
  p.H = a.H;
  p.sl2 = a.scale * 1.4426950408889634f;
  const int n_x = (a.Sq + 2 * kTile - 1) / (2 * kTile);
  const int n_tiles = (a.Skv + kTile - 1) / kTile;
  const int n_units = n_x * a.H;
  p.n_full_x = a.Sq / (2 * kTile);
  p.n_whole = n_units;
  p.n_split = 0;
  p.tiles_per_split = 0;
  p.ws = nullptr;
  // K/V split of the units that would form a last, nearly empty wave of CTAs. REGION steps: ~1600 query rows = 7 tiles
  // per head x 24 heads = 168 CTAs on 148 SMs, i.e. TWO waves for 1.14 waves of work (228 us against 125 us for one
  // wave). The r = units % SMs trailing units (the ragged tiles come last in unit order: their partials are small) are
  // cut into s = SMs / r K/V ranges, so that they run as ONE extra wave of 1/s-length CTAs. Only for r <= SMs / 4:
  // measured (profiles/r02_attn_bench_session11_wide_split.log), a part costs ~6 us of prologue, partial stores and merge on top of
  // its share of the K/V sequence, so 1576 x 8704 gains 21 % (r = 20, s = 7) and 4864^2 7 % (r = 12, s = 8), while
  // the 3-way splits that r = 76 (8704^2: 5.51 waves) or r = 96 (872 rows) would need gain nothing or lose. Needs the
  // caller's workspace (attention_workspace_bytes).
  int num_sms = 0;
  cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  const int r = num_sms > 0 ? n_units % num_sms : 0;
  if (a.workspace && tuning().attn_split != 0 && r > 0 && 4 * r <= num_sms && n_tiles >= 8 &&
      a.workspace_bytes >= attention_workspace_bytes(a.H)) {
    int sp = num_sms / r;
    if (sp > kMaxSplit) sp = kMaxSplit;
    const int per = (n_tiles + sp - 1) / sp;
    const int n_split = (n_tiles + per - 1) / per;            // no empty part
    if (n_split >= 2) {
      p.n_whole = n_units - r;
      p.n_split = n_split;
      p.tiles_per_split = per;
      p.ws = static_cast<float*>(a.workspace);
    }
  }
  const int n_split_units = n_units - p.n_whole;
  const int grid = p.n_whole + n_split_units * p.n_split;
  table[poly == 0 ? 0 : poly - 1]<<<grid, kThreads, kSmemBytes, stream>>>(mq, mk, mv, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess || p.n_split == 0) return e;
  // rows to merge per split unit: the ragged tile's rows when only ragged tiles are split (REGION steps: 40 of 256)
  const int ragged_rows = a.Sq - p.n_full_x * 2 * kTile;
  const int rows = (ragged_rows > 0 && p.n_whole >= p.n_full_x * a.H) ? ragged_rows : 2 * kTile;
  attention_combine_kernel<<<dim3(rows, n_split_units), 128, 0, stream>>>(p.ws, a.O, a.ldo, a.Sq, a.H, p.n_full_x,
                                                                         p.n_whole, p.n_split, p.sl2);
  return cudaGetLastError();
}

// Upper bound of the scratch a launch may use: at most one wave of split CTAs (r s <= SMs), each with 256 partial rows.
size_t attention_workspace_bytes(int H) {
  (void)H;
  int dev = 0, num_sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (num_sms <= 0) num_sms = 148;
  return (size_t)num_sms * 2 * kTile * kWsRow * sizeof(float);
}

cudaError_t launch_attention(const AttnArgs& a, cudaStream_t stream) {
  int k = tuning().attn_kernel;
  if (k < 0) k = kDefaultKernel;
  return k == 1 ? launch_attention64(a, stream) : launch_attention128(a, stream);
}

}  // namespace rge
