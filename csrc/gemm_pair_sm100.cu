// Persistent 2-CTA ("CTA pair", tcgen05 cta_group::2) GEMM for the large layers.
//
//   D[M,N] = alpha * A[M,K] . B[N,K]^T     same operands / epilogue contract as gemm_sm100.cu
//
// Why a second kernel: at 8192^3 the one-tile-per-CTA kernel runs at ~50 % tensor-pipe utilisation
// (profiles/r1_ncu_gemm8192.md): 1 CTA/SM, so nothing overlaps a tile's epilogue and prologue, and a 128x256
// tile needs 48 KB of operands per 64-wide k-block, which leaves only 4 ring stages to cover TMA latency.
// Here
//   * a cluster of two CTAs on the two SMs of a TPC computes one 256 x 256 tile: each CTA stages its own 128
//     rows of A and HALF of the B tile (128 rows); the leader's single `tcgen05.mma.cta_group::2` reads both
//     CTAs' shared memory.  Per CTA a k-block is 32 KB -> 6 stages in flight and 1/3 less L2->SM traffic;
//   * the kernel is persistent (one pair per TPC, tiles walked in an L2-friendly grouped order);
//   * all 512 TMEM columns are used as TWO 256-column accumulators: the epilogue warps drain tile i while the
//     MMA thread is already accumulating tile i+1.
//
// Barriers (all at the same shared-memory offsets in both CTAs):
//   full[s]       leader only.  count 1 (leader's producer arrive.expect_tx of BOTH CTAs' bytes); both
//                 producers' TMA loads complete_tx on the LEADER's barrier.
//   empty[s]      each CTA.  count 1: the leader's tcgen05.commit multicast to both CTAs.
//   tmem_full[a]  each CTA.  count 1: commit multicast after the tile's last MMA.
//   tmem_empty[a] leader only.  count 8: one arrive per epilogue warp of both CTAs.
#include "sm100_ptx.cuh"
#include "sf_api.h"

namespace sf {

constexpr int kPairBN = 256;               // tile columns (B rows): 128 staged by each CTA
constexpr int kPairHalfN = kPairBN / 2;
constexpr int kPairABytes = kBM * kBK * 2;          // 16 KB
constexpr int kPairBBytes = kPairHalfN * kBK * 2;   // 16 KB
constexpr int kPairStageBytes = kPairABytes + kPairBBytes;
constexpr int kPairBarBytes = 256;
template <int kPairStages>
constexpr int pair_smem_bytes() { return kPairStages * kPairStageBytes + kPairBarBytes + kPairBN * 4 + 1024; }

__device__ __forceinline__ void pair_tile_coords(int t, int tiles_m, int tiles_n, int group_m, int& tm, int& tn) {
  const int per_group = group_m * tiles_n;
  const int g = t / per_group;
  const int first_m = g * group_m;
  const int gm = (tiles_m - first_m) < group_m ? (tiles_m - first_m) : group_m;
  const int r = t - g * per_group;
  tm = first_m + r % gm;
  tn = r / gm;
}

template <int kPairStages>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
sf_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const SfGemmEpilogue ep,
                    int M, int N, int K, int tiles_m, int tiles_n, int group_m) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kPairStages * kPairStageBytes);
  uint64_t* empty_bar = full_bar + kPairStages;
  uint64_t* tmem_full_bar = empty_bar + kPairStages;       // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* s_bias = reinterpret_cast<float*>(smem + kPairStages * kPairStageBytes + kPairBarBytes);

  TraceScope trace;
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = static_cast<int>(cluster_id_x());
  const int n_pairs = static_cast<int>(cluster_nctaid_x());
  const int num_kb = (K + kBK - 1) / kBK;
  const int num_tiles = tiles_m * tiles_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kPairStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 8);
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc_pair<512>(tmem_slot);
  tc_fence_before_sync();
  cluster_sync_all();                      // barriers + TMEM of BOTH CTAs exist before anybody touches the peer
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  trace.mark();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      const uint64_t hintA = ep.a_evict_first ? kEvictFirst : kEvictNormal;
      const uint64_t hintB = kEvictLast;
      uint32_t it = 0;
      for (int t = pair; t < num_tiles; t += n_pairs) {
        int tm, tn;
        pair_tile_coords(t, tiles_m, tiles_n, group_m, tm, tn);
        const int m_row = tm * 2 * kBM + static_cast<int>(rank) * kBM;
        const int n_row = tn * kPairBN + static_cast<int>(rank) * kPairHalfN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % kPairStages;
          const uint32_t ph = (it / kPairStages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1, 0x500 + s);
          uint8_t* a_dst = stage_base + s * kPairStageBytes;
          uint8_t* b_dst = a_dst + kPairABytes;
          if (leader) mbar_expect_tx(&full_bar[s], 2 * kPairStageBytes);
          tma_load_2d_pair(a_dst, &tmA, &full_bar[s], kb * kBK, m_row, hintA);
          tma_load_2d_pair(b_dst, &tmB, &full_bar[s], kb * kBK, n_row, hintB);
        }
      }
      // tail: every commit that targets this CTA's empty barriers has landed before the CTA may exit
      for (int k = 0; k < kPairStages; ++k, ++it) {
        const int s = it % kPairStages;
        const uint32_t ph = (it / kPairStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1, 0x540 + s);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = umma_idesc(1 /*bf16*/, 2 * kBM, kPairBN);
      uint32_t it = 0;
      int i = 0;
      for (int t = pair; t < num_tiles; t += n_pairs, ++i) {
        const int acc = i & 1;
        mbar_wait(&tmem_empty_bar[acc], ((i >> 1) & 1) ^ 1, 0x580 + acc);      // both CTAs' epilogues drained it
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kPairBN);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % kPairStages;
          const uint32_t ph = (it / kPairStages) & 1;
          mbar_wait(&full_bar[s], ph, 0x5a0 + s);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(stage_base + s * kPairStageBytes);
          const uint64_t a_desc = umma_desc_k_sw128(a_addr);
          const uint64_t b_desc = umma_desc_k_sw128(a_addr + kPairABytes);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k)
            umma_f16_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_pair(&empty_bar[s], 0b11);       // the slot is free in BOTH CTAs once these MMAs retire
        }
        umma_commit_pair(&tmem_full_bar[acc], 0b11);   // accumulator complete: wake both epilogues
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs; rows rank*128 .. +128 of the pair tile) =====================
    const int e = warp - 4;
    const int et = threadIdx.x - 128;
    int i = 0;
    for (int t = pair; t < num_tiles; t += n_pairs, ++i) {
      int tm, tn;
      pair_tile_coords(t, tiles_m, tiles_n, group_m, tm, tn);
      const int acc = i & 1;
      const int n0 = tn * kPairBN;
      const int row = tm * 2 * kBM + static_cast<int>(rank) * kBM + e * 32 + lane;
      const bool row_ok = row < M;
      asm volatile("bar.sync 1, 128;" ::: "memory");          // previous tile's bias reads are done
      for (int j = et; j < kPairBN; j += 128) s_bias[j] = (ep.bias != nullptr && n0 + j < N) ? ep.bias[n0 + j] : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      int nchunks = (ep.n_store_limit - n0 + 31) / 32;
      if (nchunks > kPairBN / 32) nchunks = kPairBN / 32;
      mbar_wait(&tmem_full_bar[acc], (i >> 1) & 1, 0x5c0 + acc);
      tc_fence_after_sync();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(e * 32) << 16) + static_cast<uint32_t>(acc * kPairBN);
#pragma unroll 1
      for (int c = 0; c < nchunks; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(t_row + c * 32, v);
        tmem_ld_wait();
        if (c == nchunks - 1) {
          // the accumulator is in registers: hand the TMEM half back before the (slow) global stores
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[acc]);
        }
        epi_chunk<false>(ep, v, s_bias + c * 32, row, row_ok, n0 + c * 32, M, N, lane, nullptr, nullptr);
      }
      if (nchunks <= 0) {
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[acc]);
      }
    }
  }

  __syncwarp();
  tc_fence_before_sync();
  cluster_sync_all();                      // nobody exits (or frees TMEM) while the peer may still signal it
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc_pair<512>(tmem_base);
  }
  trace.end(KID_GEMM);
}

}  // namespace sf

template <int kPairStages>
static int pair_launch(const SfGemm* g, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(sf::sf_gemm_pair_kernel<kPairStages>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         sf::pair_smem_bytes<kPairStages>());
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  const int tiles_m = (g->M + 2 * sf::kBM - 1) / (2 * sf::kBM);
  const int tiles_n = (g->N + sf::kPairBN - 1) / sf::kPairBN;
  const int tiles = tiles_m * tiles_n;
  int pairs = g->pair_ctas > 0 ? g->pair_ctas / 2 : 74;
  if (pairs > tiles) pairs = tiles;
  if (pairs < 1) pairs = 1;
  // a wave of `pairs` concurrent tiles should cover a near-square block of the output: group_m tile rows at a time
  static const int env_group = [] { const char* v = getenv("SPARKFLOW_PAIR_GROUP_M"); return v ? atoi(v) : 0; }();
  int group_m = env_group > 0 ? env_group : 8;
  if (group_m > tiles_m) group_m = tiles_m;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(sf::kGemmThreads);
  cfg.dynamicSmemBytes = sf::pair_smem_bytes<kPairStages>();
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = sf::pdl_enabled() ? 1 : 0;
  return static_cast<int>(cudaLaunchKernelEx(&cfg, sf::sf_gemm_pair_kernel<kPairStages>, g->tmA, g->tmB, g->ep, g->M, g->N, g->K, tiles_m, tiles_n,
                                             group_m));
}

extern "C" int sf_gemm_pair_launch(const SfGemm* g, cudaStream_t st) {
  static const int stages = [] { const char* v = getenv("SPARKFLOW_PAIR_STAGES"); return v ? atoi(v) : 6; }();
  return stages == 7 ? pair_launch<7>(g, st) : pair_launch<6>(g, st);
}
