// Whole-chain GEMM kernel: one persistent launch runs the forward / dgrad / wgrad GEMMs of a dense network.
//
// Why: the flagship step is 8 dependent GEMM launches deep; each boundary costs ~1.5 us of hardware dependency
// latency plus a kernel prologue (barrier init, TMEM allocation, descriptor prefetch), and those dependent launches
// are exactly what host->device traffic stretches in the end-to-end loop (profiles/r1_e2e_h2d_modes.md).  Here the
// chain is ONE launch: barriers and TMEM are set up once per CTA, a stage boundary is a release / acquire pair on a
// counter in L2.
//
// Scheduling: tiles of all GEMMs form one ticket sequence in a topological order (critical path first).  A CTA draws
// a ticket, waits until every GEMM its tile depends on is complete, runs the tile with the same warp roles as
// sf_gemm_kernel<32> (TMA producer / single-thread tcgen05 issuer / TMEM epilogue with the fused training epilogues),
// and publishes it.  A waiting CTA only ever waits for lower tickets, all of which are held by running CTAs, so the
// scheme cannot deadlock whatever the grid size.
#include "sm100_ptx.cuh"
#include "sf_api.h"

namespace sf {

constexpr int kMegaBN = 32;

__device__ __forceinline__ void red_release_gpu_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(kGemmThreads, 1)
sf_mega_kernel(const SfMegaArgs a) {
  using S = GemmSmem<kMegaBN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kStages * S::kStageBytes);
  uint64_t* empty_bar = full_bar + S::kStages;
  uint64_t* tmem_full_bar = empty_bar + S::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* s_bias = reinterpret_cast<float*>(smem + S::kStages * S::kStageBytes + S::kBarBytes);
  __shared__ int s_ticket;
  __shared__ int s_begin[SF_MEGA_MAX_GEMMS + 1];

  TraceScope trace;
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tid = threadIdx.x;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < S::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<S::kTmemCols>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // the chain's inputs (cast / pull) are complete; global memory is touched below
  trace.mark();
  if (tid <= a.n_gemms) s_begin[tid] = (tid < a.n_gemms) ? a.gemms[tid].tile_begin : a.total_tiles;

  uint32_t it = 0;          // k-blocks consumed so far by this CTA (every role advances it identically)
  uint32_t tiles_run = 0;   // tiles run so far by this CTA (parity of tmem_full)
  while (true) {
    __syncthreads();        // previous tile fully retired (TMEM drained, s_ticket / s_bias reusable); s_begin visible
    if (tid == 0) s_ticket = static_cast<int>(atomicAdd(a.ctr + 0, 1u));
    __syncthreads();
    const int t = s_ticket;
    if (t >= a.total_tiles) break;
    int g = 0;
    while (g + 1 < a.n_gemms && t >= s_begin[g + 1]) ++g;
    const SfMegaGemm& G = a.gemms[g];
    const int local = t - s_begin[g];
    const int n0 = (local % G.tiles_n) * kMegaBN;
    const int m0 = (local / G.tiles_n) * kBM;
    const int M = G.M, N = G.N;
    const int num_kb = (G.K + kBK - 1) / kBK;
    const SfGemmEpilogue& ep = G.ep;

    // ---- dependencies: every tile of the producing GEMMs has been published ----
    if (tid == 0) {
      for (int d = 0; d < G.n_deps; ++d) {
        const int dg = G.deps[d];
        const unsigned int need = static_cast<unsigned int>(s_begin[dg + 1] - s_begin[dg]);
        const unsigned long long t0 = sf_globaltimer();
        unsigned int spins = 0;
        while (true) {
          unsigned int v;
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.ctr + 2 + dg) : "memory");
          if (v >= need) break;
          if ((++spins & 0x3F) == 0 && sf_globaltimer() - t0 > 2000000000ull) sf_fail(0x600 + dg);
        }
      }
      // the producers wrote through the generic proxy; this thread's TMA loads go through the async proxy
      fence_proxy_async_global();
    }
    __syncthreads();

    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        tma_prefetch_desc(&G.tmA);
        tma_prefetch_desc(&G.tmB);
        const uint64_t hintA = ep.a_evict_first ? kEvictFirst : kEvictNormal;
        for (int i = 0; i < num_kb; ++i) {
          const uint32_t gi = it + i;
          const int s = gi % S::kStages;
          const uint32_t ph = (gi / S::kStages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1, 0x610 + s);
          uint8_t* a_dst = stage_base + s * S::kStageBytes;
          uint8_t* b_dst = a_dst + S::kABytes;
          mbar_expect_tx(&full_bar[s], S::kStageBytes);
          tma_load_2d(a_dst, &G.tmA, &full_bar[s], i * kBK, m0, hintA);
          tma_load_2d(b_dst, &G.tmB, &full_bar[s], i * kBK, n0, kEvictLast);
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      if (lane == 0) {
        constexpr uint32_t idesc = umma_idesc(1 /*bf16*/, kBM, kMegaBN);
        tc_fence_after_sync();
        for (int i = 0; i < num_kb; ++i) {
          const uint32_t gi = it + i;
          const int s = gi % S::kStages;
          const uint32_t ph = (gi / S::kStages) & 1;
          mbar_wait(&full_bar[s], ph, 0x620 + s);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(stage_base + s * S::kStageBytes);
          const uint64_t a_desc = umma_desc_k_sw128(a_addr);
          const uint64_t b_desc = umma_desc_k_sw128(a_addr + S::kABytes);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) umma_f16(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[s]);
        }
        umma_commit(tmem_full_bar);
      }
    } else if (warp >= 4) {
      // ===================== epilogue =====================
      const int e = warp - 4;
      const int row = m0 + e * 32 + lane;
      const bool row_ok = row < M;
      {
        const int et = tid - 128;
        if (et < kMegaBN) s_bias[et] = (ep.bias != nullptr && n0 + et < N) ? ep.bias[n0 + et] : 0.f;
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      uint4 aux0[4] = {};
      if (ep.aux != nullptr && row_ok) {
        const __nv_bfloat16* ap = ep.aux + static_cast<size_t>(row) * ep.ld_aux + n0;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (n0 + q * 8 < ep.ld_aux) aux0[q] = *reinterpret_cast<const uint4*>(ap + q * 8);
      }
      float tgt0[32];
      if (ep.loss_mode != SF_LOSS_NONE) {
        const float* tp = ep.target + static_cast<size_t>(row_ok ? row : 0) * ep.ld_target + n0;
#pragma unroll
        for (int j = 0; j < 32; ++j) tgt0[j] = (row_ok && n0 + j < N) ? tp[j] : 0.f;
      }
      mbar_wait(tmem_full_bar, tiles_run & 1, 0x630);
      tc_fence_after_sync();
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(e * 32) << 16), v);
      tmem_ld_wait();
      tc_fence_before_sync();
      if (n0 < ep.n_store_limit) epi_chunk<true>(ep, v, s_bias, row, row_ok, n0, M, N, lane, aux0, tgt0);
    }
    it += static_cast<uint32_t>(num_kb);
    tiles_run += 1;
    __syncthreads();        // every output store of this tile has been issued by its thread
    if (tid == 0) {
      __threadfence();
      red_release_gpu_add(a.ctr + 2 + g, 1u);
    }
  }

  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc<S::kTmemCols>(tmem_base);
  }
  if (tid == 0) {
    // the last CTA out re-arms the counters for the next launch (stream order separates the launches)
    const unsigned int prev = atomicAdd(a.ctr + 1, 1u);
    if (prev == gridDim.x - 1) {
      __threadfence();
      for (int i = 0; i < 2 + a.n_gemms; ++i) a.ctr[i] = 0u;
    }
  }
  trace.end(KID_GEMM);
}

}  // namespace sf

extern "C" int sf_mega_launch(const SfMegaArgs* a, int grid, cudaStream_t st) {
  using S = sf::GemmSmem<sf::kMegaBN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(sf::sf_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kBytes);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  if (a->n_gemms < 1 || a->n_gemms > SF_MEGA_MAX_GEMMS || a->total_tiles < 1) return -9;
  if (grid <= 0) grid = 64;           // the widest stage of the flagship is 56 tiles; leave SMs to the applier
  if (grid > a->total_tiles) grid = a->total_tiles;
  if (grid > 148) grid = 148;
  return static_cast<int>(sf::launch(sf::sf_mega_kernel, dim3(grid), dim3(sf::kGemmThreads), S::kBytes, st, *a));
}
