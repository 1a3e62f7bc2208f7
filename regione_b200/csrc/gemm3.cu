// Grouped variant of the CTA-pair (cluster of two, tcgen05 cta_group::2) bf16 GEMM: up to kMaxGroup3 independent GEMMs -
// the q / k / v (/ text q / k / v) projections of a block, or the image- and text-stream halves of one stage - as ONE
// persistent launch whose tile list is the concatenation of all members' 256 x 256 tiles, cut into one contiguous,
// equally long (in k-blocks) range of whole tiles per CTA pair.
//
// Why: launched one by one, the hot path's GEMMs rarely fill a whole number of waves of the 74 SM pairs (REGION step,
// M ~ 1000-1600: 60-84 tiles = 0.8-1.1 waves; text stream, M = 512: 24 tiles = 0.3 waves), and as 128-row tiles of the
// 1-CTA kernel the small ones are bound by the L2 -> shared-memory fill. Same MMA shape, K order and epilogues as
// gemm2.cu, so results are bit-identical to the members launched one by one on that kernel.
//
// A stream-K version of this kernel (ranges cutting tiles, fp32 partials fixed up through a global workspace) was
// built and measured in round 2 (profiles/r02_gemm3_streamk_rejected.log): correct and deterministic, but the fix-up
// - 128 KB of partial accumulator per CTA through L2 per cut, read back by the tile's owner - cost more than the wave
// quantisation it removed on every shape of the path (e.g. 512 x 3072 x 3072: 168 vs 315 TFLOP/s), so it was removed.
//
//   warp 0      TMA producer (both CTAs; transaction bytes are credited to the leader's `full` barrier)
//   warp 1      MMA issuer   (leader CTA only) / TMEM allocation (both CTAs, cta_group::2)
//   warps 2-5   epilogue     (both CTAs; the member's epilogue is selected per tile)
#include "gemm_epilogue.cuh"
#include "tmap.cuh"

namespace rge {

namespace {

constexpr int BM = 128;   // rows per CTA (256 per pair)
constexpr int BN = 256;   // columns per pair (128 rows of W staged per CTA)
constexpr int BK = 64;
constexpr int kThreads = 192;
constexpr int kStages = 6;
constexpr int kABytes = BM * BK * 2;
constexpr int kBBytes = (BN / 2) * BK * 2;
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kSmemBytes = kStages * kStageBytes + 256 + 1024;
constexpr int kTmemCols = 512;
constexpr int kMaxGroup3 = 6;

struct Group2Params {
  CUtensorMap map_a[kMaxGroup3];
  CUtensorMap map_b[kMaxGroup3];
  GemmDev p[kMaxGroup3];
  long unit_end[kMaxGroup3];   // running end of each member's unit range (units = tiles x k-blocks)
  int num_m[kMaxGroup3];       // 256-row tile rows of the member
  int num_n[kMaxGroup3];
  int num_kb[kMaxGroup3];
  int epi[kMaxGroup3];
  int n_prob;
};

struct Segment {
  int prob, m_blk, n_blk, kb0, kb1;
  long tile_begin;             // unit index of the tile's first k-block
};

// unit -> (member, tile, k-block)
__device__ __forceinline__ Segment decode_unit(const Group2Params& g, long u) {
  Segment s;
  s.prob = 0;
  long base = 0;
  while (u >= g.unit_end[s.prob]) { base = g.unit_end[s.prob]; ++s.prob; }
  const int nkb = g.num_kb[s.prob];
  const long local = u - base;
  const int tile = (int)(local / nkb);
  s.kb0 = (int)(local - (long)tile * nkb);
  s.tile_begin = u - s.kb0;
  const int num_m = g.num_m[s.prob], num_n = g.num_n[s.prob];
  if (g.p[s.prob].n_fast) { s.m_blk = tile / num_n; s.n_blk = tile % num_n; }
  else { s.m_blk = tile % num_m; s.n_blk = tile / num_m; }
  s.kb1 = nkb;
  return s;
}

// first unit of pair `pair`'s range: an equal share of the k-blocks, snapped down to a tile boundary
__device__ __forceinline__ long range_begin(const Group2Params& g, long total, int pair, int num_pairs) {
  const long u = total * pair / num_pairs;
  return u < total ? decode_unit(g, u).tile_begin : u;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm3_kernel(const __grid_constant__ Group2Params g) {
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
    for (int i = 0; i < g.n_prob; ++i) {
      tma_prefetch_desc(&g.map_a[i]);
      tma_prefetch_desc(&g.map_b[i]);
    }
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 2);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);
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

  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const long total = g.unit_end[g.n_prob - 1];
  const long u0 = range_begin(g, total, pair, num_pairs);
  const long u1 = pair + 1 == num_pairs ? total : range_begin(g, total, pair + 1, num_pairs);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long u = u0; u < u1;) {
        const Segment s = decode_unit(g, u);
        const long seg_end = s.tile_begin + s.kb1;
        const CUtensorMap* ma = &g.map_a[s.prob];
        const CUtensorMap* mb = &g.map_b[s.prob];
        for (int kb = s.kb0; kb < s.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * kStageBytes);
          else mbar_arrive_cluster(&full_bar[stage], 0);
          tma_load_2d_pair(sa, ma, &full_bar[stage], kb * BK, s.m_blk * 2 * BM + (int)rank * BM);
          tma_load_2d_pair(sb, mb, &full_bar[stage], kb * BK, s.n_blk * BN + (int)rank * (BN / 2));
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        u = seg_end;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA, single thread)
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (long u = u0; u < u1; ++it) {
        const Segment s = decode_unit(g, u);
        const long seg_end = s.tile_begin + s.kb1;
        const int acc = it & 1;
        const uint32_t use = (it >> 1) & 1;
        mbar_wait(&tempty_bar[acc], use ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = s.kb0; kb < s.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint32_t sb = sa + kABytes;
          const uint64_t a_desc = make_sdesc_sw128(sa, 0, 1024);
          const uint64_t b_desc = make_sdesc_sw128(sb, 0, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_ss_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb != s.kb0 || k != 0) ? 1u : 0u);
          tc_commit_pair(&empty_bar[stage], 3);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        tc_commit_pair(&tfull_bar[acc], 3);
        u = seg_end;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (both CTAs)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    int it = 0;
    for (long u = u0; u < u1; ++it) {
      const Segment s = decode_unit(g, u);
      const int acc = it & 1;
      const uint32_t use = (it >> 1) & 1;
      mbar_wait(&tfull_bar[acc], use);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * BN;
      // the member's epilogue parameters BY VALUE: through a reference into the kernel-parameter block the compiler has
      // to re-read every field after each global store (the stores may alias it) - measured 30 % slower
      const GemmDev p = g.p[s.prob];
      const int m = s.m_blk * 2 * BM + (int)rank * BM + r, n0 = s.n_blk * BN;
      switch (g.epi[s.prob]) {
        case EPI_STORE: gemm_epilogue_row<EPI_STORE>(p, taddr, m, n0, BN); break;
        case EPI_GELU: gemm_epilogue_row<EPI_GELU>(p, taddr, m, n0, BN); break;
        case EPI_GATE_RES: gemm_epilogue_row<EPI_GATE_RES>(p, taddr, m, n0, BN); break;
        default: gemm_epilogue_row<EPI_NORM_ROPE>(p, taddr, m, n0, BN); break;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tempty_bar[acc], 0);
      u = s.tile_begin + s.kb1;
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tmem_base, kTmemCols);
}

}  // namespace

bool group2_eligible(const GemmArgs& a) {
  return a.M > 0 && a.N > 0 && a.N % BN == 0 && a.K > 0 && a.K % 8 == 0;
}

// Returns cudaErrorNotSupported if a member is outside the envelope (nothing launched).
cudaError_t launch_gemm_group2(const GemmArgs* args, int n, int num_sms, cudaStream_t stream) {
  const GemmArgs* live[kMaxGroup3];
  int n_live = 0;
  for (int i = 0; i < n; ++i) {
    if (args[i].M <= 0 || args[i].N <= 0) continue;
    if (!group2_eligible(args[i]) || n_live == kMaxGroup3) return cudaErrorNotSupported;
    live[n_live++] = &args[i];
  }
  if (n_live == 0) return cudaSuccess;
  static bool attr_set[kMaxDevices] = {};
  const int dev = current_device();
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  // members with the longest tiles (largest K) first, so that the short tiles even out the end of the launch
  int order[kMaxGroup3];
  for (int i = 0; i < n_live; ++i) order[i] = i;
  for (int i = 1; i < n_live; ++i)
    for (int j = i; j > 0 && live[order[j]]->K > live[order[j - 1]]->K; --j) {
      const int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t;
    }
  Group2Params g;
  long units = 0, tiles = 0;
  for (int k = 0; k < n_live; ++k) {
    const GemmArgs& a = *live[order[k]];
    if (!make_tmap_bf16_2d(&g.map_a[k], a.A, a.M, a.K, a.lda, BM)) return cudaErrorInvalidValue;
    if (!make_tmap_bf16_2d(&g.map_b[k], a.W, a.N, a.K, a.ldw, BN / 2)) return cudaErrorInvalidValue;
    g.p[k] = to_dev(a);
    g.p[k].n_fast = pick_n_fast(a);
    g.num_m[k] = (a.M + 2 * BM - 1) / (2 * BM);
    g.num_n[k] = a.N / BN;
    g.num_kb[k] = (a.K + BK - 1) / BK;
    g.epi[k] = a.epilogue;
    tiles += (long)g.num_m[k] * g.num_n[k];
    units += (long)g.num_m[k] * g.num_n[k] * g.num_kb[k];
    g.unit_end[k] = units;
  }
  for (int k = n_live; k < kMaxGroup3; ++k) g.unit_end[k] = units;
  g.n_prob = n_live;
  const long max_pairs = num_sms / 2;
  const long pairs = tiles < max_pairs ? tiles : max_pairs;
  gemm3_kernel<<<(unsigned)(2 * pairs), kThreads, kSmemBytes, stream>>>(g);
  return cudaGetLastError();
}

}  // namespace rge
