// Persistent, warp-specialised bf16 GEMM for sm_100a:  out = epilogue(A[M,K] @ W[N,K]^T).
//
//   warp 0      TMA producer   (cp.async.bulk.tensor, 128B-swizzled 128xBK / BNxBK tiles, kStages-deep ring)
//   warp 1      MMA issuer     (tcgen05.mma cta_group::1 kind::f16, M=128 N=bn K=16, accumulators in TMEM)
//   warps 2-5   epilogue       (tcgen05.ld -> registers -> fused epilogue -> global), one thread per output row
//
// TMEM holds two 256-column accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
// Replaces the reference's nn.Linear calls and its Triton scatter-GEMM `_partially_linear`
// (RegionE/FluxKontext/fused_kernels.py:9-101): the scatter is the `row_map` of the epilogue, and the
// per-head RMSNorm + RoPE the reference re-applies to the whole cache every step
// (RegionE/FluxKontext/inplace.py:756-763, 792-794) is fused here so the cache holds post-norm/post-RoPE rows.
// Large-M launches are routed to the CTA-pair kernel in gemm2.cu; the epilogues live in gemm_epilogue.cuh.
#include <cstdlib>

#include "gemm.cuh"
#include "gemm_epilogue.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace rge {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kThreads = 192;
constexpr int kMaxBN = 256;                 // widest tile; TMEM holds two accumulators of this width
constexpr int kMaxStages = 8;
constexpr int kABytes = BM * BK * 2;
constexpr int kRingBytes = 4 * (kABytes + kMaxBN * BK * 2);   // 192 KB operand ring, cut into stages at run time
constexpr int kBarBytes = 256;
constexpr int kSmemBytes = kRingBytes + kBarBytes + 1024;     // +1024: manual alignment slack
constexpr int kTmemCols = 2 * kMaxBN;

// Tile width `bn` and ring depth are launch parameters (not template parameters): region steps and the text stream
// have few rows (M = 512 ... ~2500), where a fixed 128 x 256 tile leaves the last wave of the 148 SMs half empty;
// the host picks the width that minimises the makespan (pick_bn below) and narrower tiles get a deeper ring.
struct TileCfg {
  int bn;       // multiple of 32 (of 128 for EPI_NORM_ROPE), 32 ... 256
  int stages;   // <= kMaxStages, stages * (kABytes + bn * 128) <= kRingBytes
};

template <int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GemmDev p,
            const TileCfg cfg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kRingBytes);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int bn = cfg.bn;
  const int n_stages = cfg.stages;
  const int stage_bytes = kABytes + bn * BK * 2;   // multiple of 1024: both operand tiles stay swizzle-aligned

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int i = 0; i < n_stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m = (p.M + BM - 1) / BM;
  const int num_n = (p.N + bn - 1) / bn;
  const int num_tiles = num_m * num_n;
  const int num_kb = (p.K + BK - 1) / BK;

  // The producer and issuer warps run their (warp-uniform) loops with ALL lanes and elect one lane per asynchronous
  // instruction group. Inside a divergent `if (lane == 0)` region ptxas wraps every UTMALDG / UTCHMMA / UTCBAR in an
  // ELECT + BRA.U.ANY loop and routes the operands through R2UR moves: ~75 issue slots per k-block, i.e. 250-350 cycles
  // for a single thread - more than the 128-256 cycles the tensor pipe needs for a k-block of a 128 x (128-256) tile.
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      tile_decode(p, tile, num_m, num_n, m_blk, n_blk);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * stage_bytes;
          uint8_t* sb = sa + kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
          tma_load_2d(sa, &map_a, &full_bar[stage], kb * BK, m_blk * BM);
          tma_load_2d(sb, &map_b, &full_bar[stage], kb * BK, n_blk * bn);
        }
        __syncwarp();
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (one elected lane)
    const uint32_t idesc = make_idesc_bf16(BM, bn, 0, 0);
    const uint64_t desc0 = make_sdesc_sw128(smem_u32(smem), 0, 1024);   // + (byte offset >> 4) per stage / operand
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t use = (it >> 1) & 1;
      mbar_wait(&tempty_bar[acc], use ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kMaxBN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t a_desc = desc0 + (uint64_t)((stage * stage_bytes) >> 4);
          const uint64_t b_desc = a_desc + (kABytes >> 4);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 B per K=16 step inside the 128 B swizzle row (encoded >>4)
            umma_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          }
          tc_commit(&empty_bar[stage]);
          if (kb == num_kb - 1) tc_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      int m_blk, n_blk;
      tile_decode(p, tile, num_m, num_n, m_blk, n_blk);
      const int acc = it & 1;
      const uint32_t use = (it >> 1) & 1;
      mbar_wait(&tfull_bar[acc], use);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * kMaxBN;
      gemm_epilogue_row<EPI>(p, taddr, m_blk * BM + r, n_blk * bn, bn, TmemRelease{&tempty_bar[acc], -1});
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------------------------- grouped launch
// Up to kMaxGroup independent GEMMs (own operands, shapes, tile widths and epilogues) as ONE persistent launch whose
// tile list is the concatenation of the members' tiles. The q / k / v (/ MLP-up) projections of a block, or the
// image- and text-stream halves of one stage of a double block, are such groups: launched one by one each of them
// fills a fraction of a wave of the 148 SMs and pays its own prologue; grouped, the CTAs stream through all tiles.
constexpr int kMaxGroup = 6;

struct GroupParams {
  CUtensorMap map_a[kMaxGroup];
  CUtensorMap map_b[kMaxGroup];
  GemmDev p[kMaxGroup];
  int tile_end[kMaxGroup];   // running end of each member's tile range (members sorted by decreasing tile cost)
  int num_m[kMaxGroup];
  int bn[kMaxGroup];
  int epi[kMaxGroup];
  int n_prob;
  int stages;
  int stage_bytes;           // kABytes + widest member's B tile
};

struct GroupTile { int prob, m_blk, n_blk; };

__device__ __forceinline__ GroupTile group_decode(const GroupParams& g, int tile) {
  GroupTile t;
  t.prob = 0;
  int start = 0;
  while (tile >= g.tile_end[t.prob]) { start = g.tile_end[t.prob]; ++t.prob; }
  const int local = tile - start;
  t.m_blk = local % g.num_m[t.prob];
  t.n_blk = local / g.num_m[t.prob];
  return t;
}

__global__ void __launch_bounds__(kThreads, 1) gemm_group_kernel(const __grid_constant__ GroupParams g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kRingBytes);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_stages = g.stages;
  const int stage_bytes = g.stage_bytes;
  const int num_tiles = g.tile_end[g.n_prob - 1];

  if (threadIdx.x == 0) {
    for (int i = 0; i < g.n_prob; ++i) {
      tma_prefetch_desc(&g.map_a[i]);
      tma_prefetch_desc(&g.map_b[i]);
    }
    for (int i = 0; i < n_stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (all lanes loop, one elected lane issues)
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const GroupTile t = group_decode(g, tile);
      const int bn = g.bn[t.prob];
      const int num_kb = (g.p[t.prob].K + BK - 1) / BK;
      const uint32_t bytes = kABytes + bn * BK * 2;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * stage_bytes;
          uint8_t* sb = sa + kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], bytes);
          tma_load_2d(sa, &g.map_a[t.prob], &full_bar[stage], kb * BK, t.m_blk * BM);
          tma_load_2d(sb, &g.map_b[t.prob], &full_bar[stage], kb * BK, t.n_blk * bn);
        }
        __syncwarp();
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (one elected lane)
    const uint64_t desc0 = make_sdesc_sw128(smem_u32(smem), 0, 1024);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const GroupTile t = group_decode(g, tile);
      const uint32_t idesc = make_idesc_bf16(BM, g.bn[t.prob], 0, 0);
      const int num_kb = (g.p[t.prob].K + BK - 1) / BK;
      const int acc = it & 1;
      const uint32_t use = (it >> 1) & 1;
      mbar_wait(&tempty_bar[acc], use ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kMaxBN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t a_desc = desc0 + (uint64_t)((stage * stage_bytes) >> 4);
          const uint64_t b_desc = a_desc + (kABytes >> 4);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          tc_commit(&empty_bar[stage]);
          if (kb == num_kb - 1) tc_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int q = warp & 3;
    const int r = q * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const GroupTile t = group_decode(g, tile);
      const int bn = g.bn[t.prob];
      const int acc = it & 1;
      const uint32_t use = (it >> 1) & 1;
      mbar_wait(&tfull_bar[acc], use);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * kMaxBN;
      const GemmDev& p = g.p[t.prob];
      const int m = t.m_blk * BM + r, n0 = t.n_blk * bn;
      const TmemRelease rel{&tempty_bar[acc], -1};
      switch (g.epi[t.prob]) {
        case EPI_STORE: gemm_epilogue_row<EPI_STORE>(p, taddr, m, n0, bn, rel); break;
        case EPI_GELU: gemm_epilogue_row<EPI_GELU>(p, taddr, m, n0, bn, rel); break;
        case EPI_GATE_RES: gemm_epilogue_row<EPI_GATE_RES>(p, taddr, m, n0, bn, rel); break;
        default: gemm_epilogue_row<EPI_NORM_ROPE, false, 0>(p, taddr, m, n0, bn, rel); break;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// Tile width for the 1-CTA kernel: minimise  waves x (bn + fixed per-tile cost)  over the widths the epilogue allows.
// The fixed cost (in output columns) stands for the prologue / epilogue drain of a tile and for the A-tile re-reads of
// narrow tiles; ties go to the wider tile. RGE_GEMM_BN=<n> forces a width (tuning / tests).
int pick_bn(const GemmArgs& a, int num_sms) {
  const int forced = tuning().gemm_bn;
  const bool heads = a.epilogue == EPI_NORM_ROPE;
  if (forced > 0 && forced <= kMaxBN && forced % (heads ? 128 : 32) == 0) return forced;
  const int num_m = (a.M + BM - 1) / BM;
  const int cand_all[] = {256, 224, 192, 160, 128, 96, 64};
  const int cand_heads[] = {256, 128};
  const int* cand = heads ? cand_heads : cand_all;
  const int n_cand = heads ? 2 : 7;
  int best = 0;
  long best_cost = 0;
  for (int i = 0; i < n_cand; ++i) {
    const int bn = cand[i];
    const long tiles = (long)num_m * ((a.N + bn - 1) / bn);
    const long waves = (tiles + num_sms - 1) / num_sms;
    const long cost = waves * (bn + 32);
    if (best == 0 || cost < best_cost) { best = bn; best_cost = cost; }
  }
  return best;
}

template <int EPI>
cudaError_t launch_t(const GemmArgs& a, int num_sms, cudaStream_t stream) {
  static bool attr_set[kMaxDevices] = {};   // per device: the attribute belongs to the device's copy of the function
  const int dev = current_device();
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  TileCfg cfg;
  cfg.bn = pick_bn(a, num_sms);
  cfg.stages = kRingBytes / (kABytes + cfg.bn * BK * 2);
  if (cfg.stages > kMaxStages) cfg.stages = kMaxStages;
  CUtensorMap map_a, map_b;
  if (!make_tmap_bf16_2d(&map_a, a.A, a.M, a.K, a.lda, BM)) return cudaErrorInvalidValue;
  if (!make_tmap_bf16_2d(&map_b, a.W, a.N, a.K, a.ldw, cfg.bn)) return cudaErrorInvalidValue;
  GemmDev p = to_dev(a);
  p.n_fast = pick_n_fast(a);
  const int num_tiles = ((a.M + BM - 1) / BM) * ((a.N + cfg.bn - 1) / cfg.bn);
  const int grid = num_tiles < num_sms ? num_tiles : num_sms;
  gemm_kernel<EPI><<<grid, kThreads, kSmemBytes, stream>>>(map_a, map_b, p, cfg);
  return cudaGetLastError();
}

cudaError_t launch_1cta(const GemmArgs& a, int num_sms, cudaStream_t stream) {
  switch (a.epilogue) {
    case EPI_STORE: return launch_t<EPI_STORE>(a, num_sms, stream);
    case EPI_GELU: return launch_t<EPI_GELU>(a, num_sms, stream);
    case EPI_GATE_RES: return launch_t<EPI_GATE_RES>(a, num_sms, stream);
    case EPI_NORM_ROPE: return launch_t<EPI_NORM_ROPE>(a, num_sms, stream);
  }
  return cudaErrorInvalidValue;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

bool gemm_args_ok(const GemmArgs& a) {
  if (a.K <= 0 || (a.K % 8) || (a.N % 32) || (a.lda % 8) || (a.ldw % 8) || (a.ldo % 8) || (a.col_off % 8)) return false;
  if (a.epilogue < EPI_STORE || a.epilogue > EPI_NORM_ROPE) return false;
  if (a.epilogue == EPI_NORM_ROPE && ((a.N % 128) || !a.norm_w || !a.rope_cs || !aligned16(a.norm_w))) return false;
  if (a.epilogue == EPI_GATE_RES && (!a.gate || !a.res || (a.ldr % 8) || !aligned16(a.gate) || !aligned16(a.res)))
    return false;
  // per-column vectors are read 16 bytes at a time in the epilogue; rows of out / res start 16-byte aligned
  if (!aligned16(a.bias) || !aligned16(a.out)) return false;
  return a.A && a.W && a.out;
}

}  // namespace

cudaError_t launch_gemm_group(const GemmArgs* args, int n, int num_sms, cudaStream_t stream) {
  const GemmArgs* live[kMaxGroup];
  int n_live = 0;
  for (int i = 0; i < n; ++i) {
    if (args[i].M <= 0 || args[i].N <= 0) continue;   // empty member (e.g. no edited token): nothing to do
    if (!gemm_args_ok(args[i]) || n_live == kMaxGroup) return cudaErrorInvalidValue;
    live[n_live++] = &args[i];
  }
  if (n_live == 0) return cudaSuccess;
  if (n_live == 1) return launch_gemm(*live[0], num_sms, stream);
  static bool attr_set[kMaxDevices] = {};
  const int dev = current_device();
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  // one width cap for the whole group: every member takes its widest allowed tile <= cap; the cap minimises
  //   (sum of tile costs) / SMs  +  half the longest tile   with tile cost = (bn + 32) * K
  // (the 32 stands for per-tile overhead and A re-reads of narrow tiles, the second term for the ragged last wave)
  const int caps[] = {256, 224, 192, 160, 128, 96, 64};
  auto width = [](const GemmArgs& a, int cap) {
    if (a.epilogue == EPI_NORM_ROPE) return cap >= 256 ? 256 : 128;
    return cap;
  };
  int best_cap = 256;
  double best_cost = 0;
  for (int c = 0; c < 7; ++c) {
    double sum = 0, longest = 0;
    for (int i = 0; i < n_live; ++i) {
      const GemmArgs& a = *live[i];
      const int bn = width(a, caps[c]);
      const double tile_cost = (double)(bn + 32) * a.K;
      sum += tile_cost * ((a.M + BM - 1) / BM) * ((a.N + bn - 1) / bn);
      if (tile_cost > longest) longest = tile_cost;
    }
    const double cost = sum / num_sms + 0.5 * longest;
    if (c == 0 || cost < best_cost) { best_cap = caps[c]; best_cost = cost; }
  }
  // longest tiles first, so that the short ones even out the end of the launch
  int order[kMaxGroup];
  for (int i = 0; i < n_live; ++i) order[i] = i;
  for (int i = 1; i < n_live; ++i)
    for (int j = i; j > 0; --j) {
      const GemmArgs &x = *live[order[j]], &y = *live[order[j - 1]];
      if ((double)(width(x, best_cap) + 32) * x.K > (double)(width(y, best_cap) + 32) * y.K) {
        const int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t;
      } else break;
    }
  GroupParams g;
  int max_bn = 0, tiles = 0;
  for (int k = 0; k < n_live; ++k) {
    const GemmArgs& a = *live[order[k]];
    const int bn = width(a, best_cap);
    if (!make_tmap_bf16_2d(&g.map_a[k], a.A, a.M, a.K, a.lda, BM)) return cudaErrorInvalidValue;
    if (!make_tmap_bf16_2d(&g.map_b[k], a.W, a.N, a.K, a.ldw, bn)) return cudaErrorInvalidValue;
    g.p[k] = to_dev(a);
    g.num_m[k] = (a.M + BM - 1) / BM;
    g.bn[k] = bn;
    g.epi[k] = a.epilogue;
    tiles += g.num_m[k] * ((a.N + bn - 1) / bn);
    g.tile_end[k] = tiles;
    if (bn > max_bn) max_bn = bn;
  }
  for (int k = n_live; k < kMaxGroup; ++k) g.tile_end[k] = tiles;
  g.n_prob = n_live;
  g.stage_bytes = kABytes + max_bn * BK * 2;
  g.stages = kRingBytes / g.stage_bytes;
  if (g.stages > kMaxStages) g.stages = kMaxStages;
  const int grid = tiles < num_sms ? tiles : num_sms;
  gemm_group_kernel<<<grid, kThreads, kSmemBytes, stream>>>(g);
  return cudaGetLastError();
}

bool gemm_args_valid(const GemmArgs& a) { return gemm_args_ok(a); }

void* get_tensor_map_encoder() { return reinterpret_cast<void*>(tensor_map_encoder()); }

cudaError_t launch_gemm(const GemmArgs& a, int num_sms, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0) return cudaSuccess;  // empty edited set: nothing to do
  if (!gemm_args_ok(a)) return cudaErrorInvalidValue;
  // Large-M launches (FULL steps) go to the CTA-pair kernel, and so do smaller ones whose rows fill 256-row tiles well
  // (<= 18 % padding: 512, 872, 1576 rows; not 1064 -> 1280): the 1-CTA kernel is bound by its L2 -> shared memory
  // operand traffic, which the pair halves (profiles/r02_gemm2_width_sweep.log: 1576 x 3072 x 3072 907 vs 805 TFLOP/s,
  // 512 x 12288 x 3072 935 vs 877). RGE_2CTA_MIN_M: -1 = this rule, 0 = never, n = from n rows on.
  const int min_m_2cta = tuning().min_m_2cta;
  const long padded = ((long)a.M + 255) / 256 * 256;
  const bool pair = min_m_2cta < 0 ? (a.M >= 2048 || padded * 100 <= (long)a.M * 118)
                                   : (min_m_2cta > 0 && a.M >= min_m_2cta);
  if (pair && a.N % 16 == 0) {
    cudaError_t e = launch_gemm_2cta(a, num_sms, stream);
    if (e != cudaErrorNotSupported) return e;
  }
  return launch_1cta(a, num_sms, stream);
}

}  // namespace rge
