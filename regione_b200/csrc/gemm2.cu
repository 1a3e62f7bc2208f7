// 2-CTA (cluster of two, tcgen05 cta_group::2) variant of the persistent bf16 GEMM for the large-M launches of a FULL
// step: one 256 x bn output tile per CTA pair (bn <= 256). Each CTA stages its own 128 rows of A and its own bn / 2
// rows (half of the tile's N) of W, the leader CTA's single MMA thread issues M=256 N=bn K=16 instructions that read
// both halves, and each CTA's TMEM receives the 128 x bn accumulator of its M half. Compared with the 1-CTA kernel this
// halves the shared-memory operand traffic per FLOP and leaves room for a 6-stage ring (32 KB per stage per CTA).
//
// The tile width is a launch parameter (any multiple of 16): with 74 pairs, 256-wide tiles leave the last wave of the
// hot shapes half empty (8704 x 3072: 408 tiles = 5.51 waves, 8192 x 3072: 384 tiles = 5.19 waves - and measured
// throughput is exactly the tail-free rate x waves / ceil(waves)); the host picks the width that minimises
// ceil(waves) x (width + per-tile overhead) (pick_bn2: 192 for the second shape). A narrower tile changes neither the k-order of
// any output element's sum nor its epilogue, so results stay bit-identical.
//
//   warp 0      TMA producer (both CTAs; transaction bytes are credited to the leader's `full` barrier)
//   warp 1      MMA issuer   (leader CTA only) / TMEM allocation (both CTAs, cta_group::2)
//   warps 2-5   epilogue     (both CTAs, shared with gemm.cu through gemm_epilogue.cuh)
#include "gemm_epilogue.cuh"
#include "tmap.cuh"

namespace rge {

namespace {

constexpr int BM = 128;   // rows per CTA (256 per pair)
constexpr int BN = 256;   // widest tile: columns per pair (128 rows of W staged per CTA), TMEM accumulator stride
constexpr int BK = 64;
constexpr int kThreads = 192;
constexpr int kStages = 6;
constexpr int kABytes = BM * BK * 2;
constexpr int kBBytes = (BN / 2) * BK * 2;
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kSmemBytes = kStages * kStageBytes + 256 + 1024;
constexpr int kTmemCols = 512;

// What a launch works on. Whole tiles: work items [0, end) are tiles of the tile order. Split tail (split > 1): work
// item w is K-range  w % split  of tile  tile_base + w / split ; its fp32 accumulators go to `ws` (one 256 x 256 slab
// per work item) instead of through the epilogue, and gemm2_finish_kernel sums the slabs and applies the epilogue.
struct WorkRange {
  int end;
  int split;
  int kb_per;      // k-blocks per K-range
  int tile_base;
  float* ws;
};
constexpr int kSlabFloats = 2 * BM * BN;

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GemmDev p,
             const int bn, const WorkRange wr) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 2);    // one arrive per CTA's producer; bytes of all four TMA loads
      mbar_init(&empty_bar[i], 1);   // multicast tcgen05.commit from the leader
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);   // multicast tcgen05.commit from the leader
      mbar_init(&tempty_bar[i], 8);  // 4 epilogue warps x 2 CTAs (used in the leader only)
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m = (p.M + 2 * BM - 1) / (2 * BM);
  const int num_n = (p.N + bn - 1) / bn;
  const int num_kb_all = (p.K + BK - 1) / BK;
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  // work item -> tile, first k-block, number of k-blocks
  constexpr bool kCanSplit = EPI != EPI_NORM_ROPE;     // that epilogue owns whole heads and needs every register
  auto decode = [&](int w, int& tile, int& kb0, int& nkb) {
    if (!kCanSplit || wr.split <= 1) {
      tile = w; kb0 = 0; nkb = num_kb_all;
    } else {
      tile = wr.tile_base + w / wr.split;
      kb0 = (w % wr.split) * wr.kb_per;
      nkb = min(wr.kb_per, num_kb_all - kb0);
    }
  };

  // Producer and issuer warps loop with all lanes and elect one lane per asynchronous instruction group (see gemm.cu:
  // a divergent `if (lane == 0)` region costs ~75 issue slots per k-block in ELECT / BRA.U.ANY loops and R2UR moves).
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    int stage = 0;
    uint32_t phase = 0;
    for (int w = pair; w < wr.end; w += num_pairs) {
      int tile, kb0, nkb, m_blk, n_blk;
      decode(w, tile, kb0, nkb);
      tile_decode(p, tile, num_m, num_n, m_blk, n_blk);
      for (int kb = kb0; kb < kb0 + nkb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * (kABytes + (bn / 2) * BK * 2));
          else mbar_arrive_cluster(&full_bar[stage], 0);
          tma_load_2d_pair(sa, &map_a, &full_bar[stage], kb * BK, m_blk * 2 * BM + (int)rank * BM);
          tma_load_2d_pair(sb, &map_b, &full_bar[stage], kb * BK, n_blk * bn + (int)rank * (bn / 2));
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA, one elected lane)
    if (leader) {
      const uint32_t idesc = make_idesc_bf16(2 * BM, bn, 0, 0);
      const uint64_t desc0 = make_sdesc_sw128(smem_u32(smem), 0, 1024);   // + (byte offset >> 4) per stage / operand
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = pair; w < wr.end; w += num_pairs, ++it) {
        int tile, kb0, num_kb;
        decode(w, tile, kb0, num_kb);
        const int acc = it & 1;
        const uint32_t use = (it >> 1) & 1;
        mbar_wait(&tempty_bar[acc], use ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t a_desc = desc0 + (uint64_t)((stage * kStageBytes) >> 4);
            const uint64_t b_desc = a_desc + (kABytes >> 4);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_ss_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
            tc_commit_pair(&empty_bar[stage], 3);
            if (kb == num_kb - 1) tc_commit_pair(&tfull_bar[acc], 3);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (both CTAs)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    int it = 0;
    for (int w = pair; w < wr.end; w += num_pairs, ++it) {
      int tile, kb0, nkb, m_blk, n_blk;
      decode(w, tile, kb0, nkb);
      tile_decode(p, tile, num_m, num_n, m_blk, n_blk);
      const int acc = it & 1;
      const uint32_t use = (it >> 1) & 1;
      mbar_wait(&tfull_bar[acc], use);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * BN;
      const TmemRelease release{&tempty_bar[acc], 0};
      if (!kCanSplit || wr.split <= 1) {
        gemm_epilogue_row<EPI, true>(p, taddr, m_blk * 2 * BM + (int)rank * BM + r, n_blk * bn, bn, release);
      } else if constexpr (kCanSplit) {
        // raw fp32 accumulators of this K-range into the work item's slab, row-major [256][BN]
        float* dst = wr.ws + (size_t)w * kSlabFloats + (size_t)((int)rank * BM + r) * BN;
#pragma unroll 1
        for (int c = 0; c < bn; c += 16) {
          uint32_t v[16];
          tmem_ld16p(taddr + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 4; ++t)
            reinterpret_cast<uint4*>(dst + c)[t] = make_uint4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
        }
        release();
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tmem_base, kTmemCols);
}

// Second half of a split tail: sums the K-range slabs of a tile in a fixed order and applies the epilogue with the
// same code the fused path runs (epilogue_chunk), 8 columns per thread, 8 rows per block.
template <int EPI>
__global__ void __launch_bounds__(256)
gemm2_finish_kernel(const GemmDev p, const int bn, const WorkRange wr) {
  const int num_m = (p.M + 2 * BM - 1) / (2 * BM);
  const int num_n = (p.N + bn - 1) / bn;
  const int t_idx = blockIdx.x / 32;                                   // tail tile
  const int row = (blockIdx.x % 32) * 8 + (threadIdx.x >> 5);         // row of the 256-row tile
  const int col = (threadIdx.x & 31) * 8;                              // first of this thread's 8 columns
  int m_blk, n_blk;
  tile_decode(p, wr.tile_base + t_idx, num_m, num_n, m_blk, n_blk);
  const int m = m_blk * 2 * BM + row, n0 = n_blk * bn + col;
  if (col >= bn || n0 >= p.N) return;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int s = 0; s < wr.split; ++s) {
    const float4* src = reinterpret_cast<const float4*>(wr.ws + (size_t)(t_idx * wr.split + s) * kSlabFloats +
                                                        (size_t)row * BN + col);
    const float4 x = src[0], y = src[1];
    acc[0] += x.x; acc[1] += x.y; acc[2] += x.z; acc[3] += x.w;
    acc[4] += y.x; acc[5] += y.y; acc[6] += y.z; acc[7] += y.w;
  }
  const bool valid = m < p.M;
  const long out_row = valid ? (long)((p.row_map ? __ldg(p.row_map + m) : m) + p.row_off) : 0;
  __nv_bfloat16* out_ptr = p.out + out_row * p.ldo + p.col_off;
  uint32_t v[32];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(acc[j]);
  uint4 rv[4];
  load_residual<EPI, 1>(p, rv, m, n0, valid);
  epilogue_chunk<EPI, 1>(p, v, rv, out_ptr, n0, valid);
}

// Split of the last, partial wave of tiles along K. With 74 pairs the K-heavy launches of a FULL step leave their last
// wave mostly empty (FF-down 8192 x 3072 x 12288: 384 tiles = 5.19 waves of 65-us tiles). The R = tiles mod pairs
// trailing tiles are cut into S K-ranges each (S <= 4, chosen like the attention split: ceil(R S / pairs) / S + 0.04 S
// full-tile times against 1), computed by a second launch of the same kernel on a library-owned side stream - its CTAs
// become resident as the whole-tile launch drains - and finished by gemm2_finish_kernel. No in-kernel spinning on
// other CTAs. Only for K >= 6144 (the slabs cost ~8 us) and not for EPI_NORM_ROPE (its epilogue owns whole heads).
struct SplitRes {
  cudaStream_t side = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  float* ws = nullptr;
  size_t ws_floats = 0;
};
SplitRes* split_resources(int max_pairs) {
  static SplitRes res[kMaxDevices];
  SplitRes& r = res[current_device()];
  if (!r.side) {
    const size_t n = (size_t)2 * max_pairs * kSlabFloats;
    if (cudaStreamCreateWithFlags(&r.side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&r.ev_in, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&r.ev_out, cudaEventDisableTiming) != cudaSuccess ||
        cudaMalloc(&r.ws, n * sizeof(float)) != cudaSuccess) {
      cudaGetLastError();
      r.side = nullptr;
      return nullptr;
    }
    r.ws_floats = n;
  }
  return &r;
}

// Cost, in full-tile times, of the trailing `tail` tiles when each is cut into the best number of K-ranges (<= 4, each
// at least 16 k-blocks, at most two waves of parts): ceil(tail s / pairs) / s + 0.04 s against 1 for the unsplit wave.
double tail_cost(int tail, int pairs, int num_kb, bool allow_split, int* split) {
  *split = 1;
  if (tail == 0) return 0.0;
  double best = 1.0;
  if (allow_split && num_kb >= 96) {
    for (int sp = 2; sp <= 4 && tail * sp <= 2 * pairs && num_kb / sp >= 16; ++sp) {
      const double cost = (double)((tail * sp + pairs - 1) / pairs) / sp + 0.04 * sp;
      if (cost < best - 0.15) { best = cost; *split = sp; }
    }
  }
  return best;
}

// Tile width: minimise  ceil(tiles / pairs) x (bn + 60) [x 1.04 for odd multiples of 16]  over the multiples of 16 (of
// 128 for EPI_NORM_ROPE, whose epilogue owns whole heads). The constants are fitted to the measured width sweep
// (profiles/r02_gemm2_width_sweep.log, tools/gemm2_width_sweep.py): a narrower tile re-reads the same 128 x 64 A block
// from shared memory for fewer FLOPs (the 256-wide tile sits exactly at the 128 B/clk the SS MMA may read), and widths
// that are not multiples of 32 pay the 16-column epilogue tail. RGE_GEMM2_BN / "gemm2_bn" forces a width.
// The K-split of the trailing wave (tail_cost) is chosen together with the width: waves = full waves + tail cost.
int pick_bn2(const GemmArgs& a, int pairs, int* split) {
  const bool heads = a.epilogue == EPI_NORM_ROPE;
  const bool allow_split = !heads && tuning().split_tail;
  const int num_kb = (a.K + BK - 1) / BK;
  const long num_m = (a.M + 2 * BM - 1) / (2 * BM);
  const int forced = tuning().gemm2_bn;
  int best = 0;
  double best_cost = 0;
  for (int bn = BN; bn >= (heads ? 128 : 64); bn -= heads ? 128 : 16) {
    if (forced >= 16 && forced <= BN && forced % (heads ? 128 : 16) == 0 && bn != forced) continue;
    const long tiles = num_m * ((a.N + bn - 1) / bn);
    int sp;
    const double waves = (double)(tiles / pairs) + tail_cost((int)(tiles % pairs), pairs, num_kb, allow_split, &sp);
    const double cost = waves * (bn + 60) * (bn % 32 ? 1.04 : 1.0);
    if (best == 0 || cost < best_cost) { best = bn; best_cost = cost; *split = sp; }
  }
  return best;
}

template <int EPI>
cudaError_t launch_t(const GemmArgs& a, int num_sms, cudaStream_t stream) {
  static bool attr_set[kMaxDevices] = {};   // per device
  const int dev = current_device();
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  const int max_pairs = num_sms / 2;
  int split = 1;
  const int bn = pick_bn2(a, max_pairs, &split);
  CUtensorMap map_a, map_b;
  if (!make_tmap_bf16_2d(&map_a, a.A, a.M, a.K, a.lda, BM)) return cudaErrorInvalidValue;
  if (!make_tmap_bf16_2d(&map_b, a.W, a.N, a.K, a.ldw, bn / 2)) return cudaErrorInvalidValue;
  GemmDev p = to_dev(a);
  p.n_fast = pick_n_fast(a);
  const int num_tiles = ((a.M + 2 * BM - 1) / (2 * BM)) * ((a.N + bn - 1) / bn);
  const int num_kb = (a.K + BK - 1) / BK;
  const int tail = num_tiles % max_pairs;
  SplitRes* res = split > 1 ? split_resources(max_pairs) : nullptr;
  if (!res) {
    const int pairs = num_tiles < max_pairs ? num_tiles : max_pairs;
    gemm2_kernel<EPI><<<2 * pairs, kThreads, kSmemBytes, stream>>>(map_a, map_b, p, bn,
                                                                   WorkRange{num_tiles, 1, 0, 0, nullptr});
    return cudaGetLastError();
  }
  const int full = num_tiles - tail;
  const int kb_per = (num_kb + split - 1) / split;
  const int parts = (num_kb + kb_per - 1) / kb_per;          // no empty K-range
  const WorkRange tail_wr{tail * parts, parts, kb_per, full, res->ws};
  cudaError_t e = cudaEventRecord(res->ev_in, stream);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(res->side, res->ev_in, 0);
  if (e != cudaSuccess) return e;
  if (full > 0)
    gemm2_kernel<EPI><<<2 * max_pairs, kThreads, kSmemBytes, stream>>>(map_a, map_b, p, bn,
                                                                       WorkRange{full, 1, 0, 0, nullptr});
  const int tail_pairs = tail_wr.end < max_pairs ? tail_wr.end : max_pairs;
  gemm2_kernel<EPI><<<2 * tail_pairs, kThreads, kSmemBytes, res->side>>>(map_a, map_b, p, bn, tail_wr);
  gemm2_finish_kernel<EPI><<<tail * 32, 256, 0, res->side>>>(p, bn, tail_wr);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaEventRecord(res->ev_out, res->side);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(stream, res->ev_out, 0);
  return e;
}

}  // namespace

cudaError_t launch_gemm_2cta(const GemmArgs& a, int num_sms, cudaStream_t stream) {
  if (a.N % 16) return cudaErrorNotSupported;
  switch (a.epilogue) {
    case EPI_STORE: return launch_t<EPI_STORE>(a, num_sms, stream);
    case EPI_GELU: return launch_t<EPI_GELU>(a, num_sms, stream);
    case EPI_GATE_RES: return launch_t<EPI_GATE_RES>(a, num_sms, stream);
    case EPI_NORM_ROPE: return launch_t<EPI_NORM_ROPE>(a, num_sms, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace rge
