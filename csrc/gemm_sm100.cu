// tcgen05 / TMEM / TMA GEMM for sm_100a with fused training epilogues.
//
//   D[M,N] = alpha * A[M,K] . B[N,K]^T          (both operands K-major, bf16, fp32 accumulate)
//
// One CTA computes one 128 x BN output tile (optionally one K-split of it):
//   warp 0   : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx-count)
//   warp 1   : MMA issuer    (one elected lane issues tcgen05.mma, accumulator lives in TMEM)
//   warp 2   : TMEM allocator / deallocator
//   warps 4-7: epilogue      (tcgen05.ld TMEM -> registers -> fused math -> global)
//
// Fused epilogues cover everything a dense layer needs in training (reference ops K1/K2 of
// MatMul+BiasAdd+{Relu,Sigmoid,Tanh}, their gradients and BiasAddGrad):
//   * bias add + activation, bf16 output and its transpose (the transpose feeds the K-major
//     operand of the next wgrad GEMM, so no separate transpose kernel exists),
//   * dgrad: multiply by act'(previous activation) read from the forward buffer,
//   * column sums (bias gradient) via an in-register transpose-reduce + one red per column,
//   * fp32 store or atomic accumulate (split-K / straight into the flat gradient buffer that the
//     push kernel consumes).
//
// The B operand may live in *peer* GPU memory: the tensor map simply carries the peer-mapped
// address, and the TMA engine pulls weight tiles over NVLink while the MMA pipeline runs
// (the "pull-fused" first dense layer of the parameter-server design).
#include <cstdlib>
#include "sm100_ptx.cuh"
#include "sf_api.h"

namespace sf {

constexpr int kBM = 128;
constexpr int kBK = 64;           // 64 bf16 = 128 B = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 256;

// Fast forms: the results are rounded to bf16 (or feed a loss) - a slow-path division per element would triple the
// epilogue's code size, and the epilogue is instruction-fetch bound (it runs once per launch on a cold I-cache).
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) {
  const float t = __expf(-2.f * fabsf(x));
  return copysignf(__fdividef(1.f - t, 1.f + t), x);
}
__device__ __forceinline__ float act_fwd(float x, int act) {
  switch (act) {
    case SF_ACT_RELU: return fmaxf(x, 0.f);
    case SF_ACT_SIGMOID: return sigmoid_fast(x);
    case SF_ACT_TANH: return tanh_fast(x);
    default: return x;
  }
}
// derivative expressed through the activation *output* a
__device__ __forceinline__ float act_bwd_from_out(float a, int act) {
  switch (act) {
    case SF_ACT_RELU: return a > 0.f ? 1.f : 0.f;
    case SF_ACT_SIGMOID: return a * (1.f - a);
    case SF_ACT_TANH: return 1.f - a * a;
    default: return 1.f;
  }
}
// whole-chunk forms: ONE (uniform) branch on the activation kind instead of one per element
__device__ __forceinline__ void act_fwd32(float (&f)[32], int act) {
  if (act == SF_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
  } else if (act == SF_ACT_SIGMOID) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = sigmoid_fast(f[j]);
  } else if (act == SF_ACT_TANH) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = tanh_fast(f[j]);
  }
}
// f[j] *= act'(a[j]) (a = activation output); keep in (0, 1): a = act(z) * mask / keep (fused dropout), mask = a != 0
__device__ __forceinline__ void act_bwd32(float (&f)[32], const float (&a)[32], int act, float keep) {
  if (keep > 0.f && keep < 1.f) {
    const float ik = __fdividef(1.f, keep);
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] *= a[j] != 0.f ? act_bwd_from_out(a[j] * keep, act) * ik : 0.f;
  } else if (act == SF_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = a[j] > 0.f ? f[j] : 0.f;
  } else if (act == SF_ACT_SIGMOID) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] *= a[j] * (1.f - a[j]);
  } else if (act == SF_ACT_TANH) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] *= 1.f - a[j] * a[j];
  }
}

// One 32-column chunk of a thread's output row: everything between the accumulator (v, already in registers) and
// global memory.  `aux_pref` / `tgt_pref` are the act'(a) operand / label row of this chunk when the caller fetched
// them before the accumulator wait (nullptr: read here).  kLoss compiles the fused softmax-CE / MSE heads in.
template <bool kLoss>
__device__ __forceinline__ void epi_chunk(const SfGemmEpilogue& ep, const uint32_t (&v)[32], const float* s_bias32, int row, bool row_ok,
                                          int col0, int M, int N, int lane, const uint4* aux_pref, const float* tgt_pref,
                                          unsigned long long* tq = nullptr) {
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * ep.alpha + s_bias32[j];
    if (tq != nullptr) tq[0] = trace_now();
    act_fwd32(f, ep.act);
    if (tq != nullptr) tq[1] = trace_now();
    if (ep.drop_keep > 0.f && ep.drop_keep < 1.f) {
      // fused dropout: keep with probability drop_keep, scale the survivors by 1 / drop_keep
      const float inv_keep = __fdividef(1.f, ep.drop_keep);
      const uint32_t thr = static_cast<uint32_t>(fminf(ep.drop_keep * 4294967296.f, 4294967040.f));
      const uint32_t stepc = ep.drop_ctr != nullptr ? *ep.drop_ctr : 0u;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const uint4 rnd = philox4x32_10(make_uint4(static_cast<uint32_t>(row), static_cast<uint32_t>(col0 >> 2) + g, stepc, ep.drop_stream),
                                        make_uint2(ep.drop_seed, 0x5F3759DFu));
        f[4 * g + 0] = rnd.x < thr ? f[4 * g + 0] * inv_keep : 0.f;
        f[4 * g + 1] = rnd.y < thr ? f[4 * g + 1] * inv_keep : 0.f;
        f[4 * g + 2] = rnd.z < thr ? f[4 * g + 2] * inv_keep : 0.f;
        f[4 * g + 3] = rnd.w < thr ? f[4 * g + 3] * inv_keep : 0.f;
      }
    }
    if (tq != nullptr) tq[2] = trace_now();
    if (kLoss && ep.loss_mode == SF_LOSS_SOFTMAX_XENT) {
      // whole row lives in this thread (host guarantees N <= 32): softmax + CE + gradient in registers
      const float* yv = tgt_pref;                         // N <= 32: the single chunk was prefetched
      const int nvalid = row_ok ? (N - col0 < 32 ? N - col0 : 32) : 0;
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, j < nvalid ? f[j] : -INFINITY);
      float se = 0.f, ysum = 0.f, zy = 0.f;
      float ex[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const bool ok = j < nvalid;
        ex[j] = ok ? __expf(f[j] - mx) : 0.f;
        const float y = ok ? yv[j] : 0.f;
        se += ex[j];
        ysum += y;
        zy += ok ? y * f[j] : 0.f;
      }
      const float inv_b = __fdividef(1.f, static_cast<float>(M));
      float lrow = row_ok ? ((mx + __logf(se)) * ysum - zy) * inv_b : 0.f;
      const float inv_se = row_ok ? __fdividef(1.f, se) : 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = j < nvalid ? (ex[j] * inv_se * ysum - yv[j]) * inv_b : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lrow += __shfl_xor_sync(0xffffffffu, lrow, o);
      if (lane == 0 && lrow != 0.f) atomicAdd(ep.loss, lrow);
    } else if (kLoss && ep.loss_mode == SF_LOSS_MSE) {
      const float scale = 2.f / (static_cast<float>(M) * static_cast<float>(N));
      const float* tp = ep.target + static_cast<size_t>(row_ok ? row : 0) * ep.ld_target + col0;
      float lrow = 0.f;
      float a_out[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        a_out[j] = f[j];
        if (row_ok && col0 + j < N) {
          const float d = f[j] - (tgt_pref != nullptr ? tgt_pref[j] : tp[j]);
          lrow += d * d;
          f[j] = d * scale;
        } else {
          f[j] = 0.f;
        }
      }
      act_bwd32(f, a_out, ep.act, 0.f);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lrow += __shfl_xor_sync(0xffffffffu, lrow, o);
      if (lane == 0 && lrow != 0.f) atomicAdd(ep.loss, lrow * 0.5f * scale);
    }
    if (tq != nullptr) tq[3] = trace_now();
    if (ep.aux != nullptr && row_ok) {
      const __nv_bfloat16* ap = ep.aux + static_cast<size_t>(row) * ep.ld_aux + col0;
      float a32[32];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 q = make_uint4(0u, 0u, 0u, 0u);
        if (col0 + g * 8 < ep.ld_aux) q = (aux_pref != nullptr) ? aux_pref[g] : *reinterpret_cast<const uint4*>(ap + g * 8);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 a2 = __bfloat1622float2(h[t]);
          a32[g * 8 + 2 * t] = a2.x;
          a32[g * 8 + 2 * t + 1] = a2.y;
        }
      }
      act_bwd32(f, a32, ep.aux_act, ep.aux_keep);
    }
    if (tq != nullptr) tq[4] = trace_now();
    // padded columns / rows contribute exact zeros everywhere below
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (col0 + j >= N || !row_ok) f[j] = 0.f;

    if (ep.colsum != nullptr) {
      // transpose-reduce over the 32 rows held by this warp: 31 shuffles, lane j ends up with
      // the sum of column col0 + j.
      float r[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = f[j];
#pragma unroll
      for (int w = 16; w >= 1; w >>= 1) {
        const bool upper = (lane & w) != 0;
#pragma unroll
        for (int j = 0; j < w; ++j) {
          const float send = upper ? r[j] : r[j + w];
          const float keep = upper ? r[j + w] : r[j];
          r[j] = keep + __shfl_xor_sync(0xffffffffu, send, w);
        }
      }
      // after the butterfly lane l holds column bitrev-free index: position is l itself
      if (col0 + lane < N) atomicAdd(ep.colsum + col0 + lane, r[0]);
    }

    if (tq != nullptr) tq[5] = trace_now();          // tracer: math done, stores follow
    if ((ep.out_f32 != nullptr || ep.route != nullptr) && row_ok) {
      float* op;
      if (ep.route != nullptr) {
        // push fused into the wgrad epilogue: this 32-column chunk lies inside ONE 32 x 64 push tile; its gradient goes
        // straight into this worker's mailbox on the GPU that owns the tile (NVLink peer stores / reds)
        const int tile = ep.route_tile0 + (row >> 5) * ep.route_tiles_c + (col0 >> 6);
        int owner = 0;
#pragma unroll
        for (int r = 1; r < SF_MAX_SHARDS; ++r) owner += (r < ep.route->n_shards && tile >= ep.route->bounds[r]) ? 1 : 0;
        op = ep.route->mailbox[owner] + ep.route_off + static_cast<size_t>(row) * ep.ld_f32 + col0;
      } else {
        op = ep.out_f32 + static_cast<size_t>(row) * ep.ld_f32 + col0;
      }
      if (ep.accumulate) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < N) atomicAdd(op + j, f[j]);
      } else if ((ep.ld_f32 & 3) == 0 && col0 + 32 <= N &&
                 (reinterpret_cast<uintptr_t>(op) & 15) == 0) {
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *reinterpret_cast<float4*>(op + 4 * g) =
              make_float4(f[4 * g], f[4 * g + 1], f[4 * g + 2], f[4 * g + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < N) op[j] = f[j];
      }
    }
    if (tq != nullptr) tq[6] = trace_now();
    if (ep.out_bf16 != nullptr && row_ok) {
      __nv_bfloat16* op = ep.out_bf16 + static_cast<size_t>(row) * ep.ld_bf16 + col0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (col0 + g * 8 < ep.ld_bf16) {
          uint4 q;
          q.x = pack_bf16x2(f[g * 8 + 0], f[g * 8 + 1]);
          q.y = pack_bf16x2(f[g * 8 + 2], f[g * 8 + 3]);
          q.z = pack_bf16x2(f[g * 8 + 4], f[g * 8 + 5]);
          q.w = pack_bf16x2(f[g * 8 + 6], f[g * 8 + 7]);
          *reinterpret_cast<uint4*>(op + g * 8) = q;
        }
      }
    }
    if (tq != nullptr) tq[7] = trace_now();
    if (ep.outT_bf16 != nullptr && row_ok) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < N)
          ep.outT_bf16[static_cast<size_t>(col0 + j) * ep.ld_t + row] = __float2bfloat16(f[j]);
    }
}

template <int BN>
struct GemmSmem {
  // ~200 KB of operand ring per CTA: small-N tiles are latency bound, so they get the deepest ring
  // (all 13 k-blocks of the flagship's 784-wide layer are in flight at once)
  static constexpr int kStages = (BN >= 256) ? 4 : (BN >= 128) ? 6 : (BN >= 64) ? 8 : 10;
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;
  // + barriers (full, empty per stage, tmem_full) + tmem slot (256 B) + bias tile (BN floats) + 1024 alignment slack
  static constexpr int kBarBytes = 256;
  static constexpr int kBytes = kStages * kStageBytes + kBarBytes + BN * 4 + 1024;
};

// MN: both operands MN-major (D = a^T . b with a [K, M], b [K, N] row-major): the stage buffers have the same size, but a
// stage is filled with {64 MN elements, 64 K rows} boxes and described to the tensor core as MN-major.
template <int BN, bool MN = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
sf_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const SfGemmEpilogue ep, int M, int N, int K, int kblocks_per_split) {
  static_assert(!MN || BN % 64 == 0, "MN-major B needs whole 64-element swizzle atoms");
  using S = GemmSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle needs 1024-byte aligned stage bases
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kStages * S::kStageBytes);
  uint64_t* empty_bar = full_bar + S::kStages;
  uint64_t* tmem_full_bar = empty_bar + S::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* s_bias = reinterpret_cast<float*>(smem + S::kStages * S::kStageBytes + S::kBarBytes);

  TraceScope trace;
  // A GEMM that carries the fused read-your-writes wait may spin for a while: do NOT let its successors pre-launch behind
  // it (they would sit resident on SMs - one CTA per SM at this shared-memory size - that the shard applier whose
  // acknowledgement this kernel waits for needs for its own CTAs).
  if (ep.ryw == nullptr) pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * kBM;
  const int total_kb = (K + kBK - 1) / kBK;
  const int kb_begin = blockIdx.z * kblocks_per_split;
  int kb_end = kb_begin + kblocks_per_split;
  if (kb_end > total_kb) kb_end = total_kb;
  const int num_kb = kb_end - kb_begin;   // host guarantees >= 1

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
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
  pdl_wait();      // everything above overlapped the previous kernel; global memory is touched below
  trace.mark();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      if (ep.ryw != nullptr) {
        // pull fused into this GEMM: read-your-writes wait before the first weight tile is fetched (local spin)
        const SfRyw& w = *ep.ryw;
        const uint32_t posted = *w.my_posted;
        const unsigned long long t0 = sf_globaltimer();
        for (int r = 0; r < w.n_shards; ++r) {
          const uint32_t want = posted * static_cast<uint32_t>(w.ack_grid[r] ? w.ack_grid[r] : 1);
          while (static_cast<int32_t>(ld_relaxed_sys(w.applied + r * 16) - want) < 0) {
            if (sf_globaltimer() - t0 > 20000000000ull) sf_fail(0x405);
          }
        }
        if (w.stats != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
          const unsigned long long t1 = sf_globaltimer(), t_post = w.stats[0];
          if (t_post != 0 && t0 > t_post) {
            w.stats[8] += t0 - t_post;
            w.stats[9] += t1 - t0;
            w.stats[11] += 1ull;
          }
        }
      }
      // weights (B) are re-read by every M tile: keep them in L2; activations stream through.
      const uint64_t hintA = ep.a_evict_first ? kEvictFirst : kEvictNormal;
      const uint64_t hintB = kEvictLast;
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % S::kStages;
        const uint32_t ph = (i / S::kStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1, 0x100 + s);
        uint8_t* a_dst = stage_base + s * S::kStageBytes;
        uint8_t* b_dst = a_dst + S::kABytes;
        mbar_expect_tx(&full_bar[s], S::kStageBytes);
        const int kc = (kb_begin + i) * kBK;
        if constexpr (MN) {
#pragma unroll
          for (int h = 0; h < kBM / 64; ++h) tma_load_2d(a_dst + h * 8192, &tmA, &full_bar[s], m0 + h * 64, kc, hintA);
#pragma unroll
          for (int h = 0; h < BN / 64; ++h) tma_load_2d(b_dst + h * 8192, &tmB, &full_bar[s], n0 + h * 64, kc, hintB);
        } else {
          tma_load_2d(a_dst, &tmA, &full_bar[s], kc, m0, hintA);
          tma_load_2d(b_dst, &tmB, &full_bar[s], kc, n0, hintB);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(1 /*bf16*/, kBM, BN, MN);
      const bool tr = g_trace_buf != nullptr;
      unsigned long long tm0 = tr ? trace_now() : 0, tm1 = 0;
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % S::kStages;
        const uint32_t ph = (i / S::kStages) & 1;
        mbar_wait(&full_bar[s], ph, 0x200 + s);
        if (tr && i == 0) tm1 = trace_now();
        tc_fence_after_sync();
        const uint32_t a_addr = smem_u32(stage_base + s * S::kStageBytes);
        const uint32_t b_addr = a_addr + S::kABytes;
        if constexpr (MN) {
          // 64 K rows of 128 B per 64-element MN atom (8 KB); one MMA consumes 16 rows = 2 KB = +128 in 16-byte units
          const uint64_t a_desc = umma_desc_mn_sw128(a_addr, 8192);
          const uint64_t b_desc = umma_desc_mn_sw128(b_addr, 8192);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k)
            umma_f16(tmem_base, a_desc + 128 * k, b_desc + 128 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
        } else {
          const uint64_t a_desc = umma_desc_k_sw128(a_addr);
          const uint64_t b_desc = umma_desc_k_sw128(b_addr);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            // advance 16 elements (32 B) along K inside the swizzle atom: +2 in 16-byte units
            umma_f16(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[s]);      // smem slot reusable once these MMAs retire
      }
      umma_commit(tmem_full_bar);        // accumulator complete
      if (tr) {                           // phase record: role start / first operands landed / last MMA issued
        const unsigned int i = atomicAdd(&g_trace_n, 1u);
        if (i < g_trace_cap) g_trace_buf[i] = TraceRec{tm0, tm1, trace_now(), 101u, blockIdx.x + gridDim.x * blockIdx.y};
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int e = warp - 4;              // == warp % 4 -> TMEM lane quarter this warp may touch
    const bool tr = g_trace_buf != nullptr && e == 0 && lane == 0;
    const unsigned long long te0 = tr ? trace_now() : 0;
    const int row = m0 + e * 32 + lane;
    const bool row_ok = row < M;
    const bool is_split0 = (blockIdx.z == 0);
    // ---- everything that does not depend on the accumulator is fetched while the main loop runs ----
    {
      const int et = threadIdx.x - 128;                        // 0..127
      for (int j = et; j < BN; j += 128)
        s_bias[j] = (ep.bias != nullptr && is_split0 && n0 + j < N) ? ep.bias[n0 + j] : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");           // epilogue warps only
    }
    const SfGemmEpilogue& eps = ep;
    uint4 aux0[4] = {};
    if (eps.aux != nullptr && row_ok) {
      const __nv_bfloat16* ap = eps.aux + static_cast<size_t>(row) * eps.ld_aux + n0;
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (n0 + g * 8 < eps.ld_aux) aux0[g] = *reinterpret_cast<const uint4*>(ap + g * 8);
    }
    float tgt0[32];
    if (eps.loss_mode != SF_LOSS_NONE) {
      const float* tp = eps.target + static_cast<size_t>(row_ok ? row : 0) * eps.ld_target + n0;
#pragma unroll
      for (int j = 0; j < 32; ++j) tgt0[j] = (row_ok && n0 + j < N) ? tp[j] : 0.f;
    }
    // touch every line of the epilogue descriptor now: the kernel-parameter bank is cold at every launch, and each first
    // touch after the accumulator is ready would be a constant-cache miss on the critical path
    asm volatile("" ::"l"(eps.out_f32), "l"(eps.colsum), "r"(eps.n_store_limit), "l"(eps.target), "l"(eps.loss), "l"(eps.route),
                 "l"(eps.route_off), "l"(eps.drop_ctr), "f"(eps.aux_keep), "f"(eps.alpha), "r"(eps.accumulate));
    mbar_wait(tmem_full_bar, 0, 0x300);
    tc_fence_after_sync();
    const unsigned long long te1 = tr ? trace_now() : 0;
    unsigned long long te2 = 0, tq8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(e * 32) << 16) + c * 32, v);
      tmem_ld_wait();
      if (tr && c == 0) te2 = trace_now();
      const int col0 = n0 + c * 32;
      if (col0 >= eps.n_store_limit) break;   // whole chunk outside every output pitch
      epi_chunk<true>(eps, v, s_bias + c * 32, row, row_ok, col0, M, N, lane, c == 0 ? aux0 : nullptr, c == 0 ? tgt0 : nullptr,
                      (tr && c == 0) ? tq8 : nullptr);
    }
    if (tr) {                             // phase records: wait start / accumulator ready / epilogue done;
      const unsigned long long te4 = trace_now();                 // first chunk in registers / its math done / done
      const unsigned int i = atomicAdd(&g_trace_n, 4u);
      if (i + 3 < g_trace_cap) {
        const unsigned int bid = blockIdx.x + gridDim.x * blockIdx.y;
        g_trace_buf[i] = TraceRec{te0, te1, te4, 102u, bid};
        g_trace_buf[i + 1] = TraceRec{te2, tq8[0], tq8[1], 103u, bid};        // chunk 0 in registers / alpha+bias / activation
        g_trace_buf[i + 2] = TraceRec{tq8[2], tq8[3], tq8[4], 104u, bid};     // dropout / loss head / act'(aux)
        g_trace_buf[i + 3] = TraceRec{tq8[5], tq8[6], tq8[7], 105u, bid};     // padding + colsum / fp32 out / bf16 out; then outT -> 102.t2
      }
    }
    tc_fence_before_sync();
  }

  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc<S::kTmemCols>(tmem_base);
  }
  trace.end(KID_GEMM);
}

// ---------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

}  // namespace sf

extern "C" int sf_make_tmap_bf16_kmajor(CUtensorMap* out, const void* base, uint64_t rows,
                                        uint64_t cols, uint64_t ld_elems, uint32_t box_rows) {
  auto fn = sf::get_encode();
  if (!fn) return -1;
  cuuint64_t dims[2] = {cols, rows};                 // innermost first
  cuuint64_t strides[1] = {ld_elems * 2};            // bytes, dim 1
  cuuint32_t box[2] = {static_cast<cuuint32_t>(sf::kBK), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : static_cast<int>(r);
}

extern "C" int sf_gemm_pick_bn(int M, int N) {
  // Small problems are latency bound: prefer more CTAs (smaller BN) until the grid covers the
  // 148 SMs; big problems want the widest tile for operand reuse.
  const int m_tiles = (M + sf::kBM - 1) / sf::kBM;
  const int cands[4] = {256, 128, 64, 32};
  for (int i = 0; i < 4; ++i) {
    const int bn = cands[i];
    if (bn > 32 && bn / 2 >= N) continue;              // tile mostly padding
    const int n_tiles = (N + bn - 1) / bn;
    if (m_tiles * n_tiles >= 148 || bn == 32) return bn;
  }
  return 32;
}

extern "C" int sf_gemm_prepare(SfGemm* g) {
  if (g->M <= 0 || g->N <= 0 || g->K <= 0) return -2;
  if (g->bn == 0) g->bn = sf_gemm_pick_bn(g->M, g->N);
  if (g->split_k <= 0) g->split_k = 1;
  // large, un-split problems go to the persistent 2-CTA kernel (gemm_pair_sm100.cu): 256x256 tiles per CTA pair
  {
    const long long pair_tiles = static_cast<long long>((g->M + 255) / 256) * ((g->N + 255) / 256);
    const bool eligible = g->split_k == 1 && g->ep.loss_mode == SF_LOSS_NONE && g->M >= 256 && g->N >= 256;
    const char* env = getenv("SPARKFLOW_GEMM_PAIR");
    int want = g->pair;
    if (env != nullptr && want < 0) want = atoi(env);
    if (want < 0) want = (g->bn == 256 && pair_tiles >= 74 && g->K >= 512) ? 1 : 0;
    g->pair = (want > 0 && eligible && !g->mn_major) ? 1 : 0;
    if (g->pair) g->bn = 256;
  }
  int rc;
  if (g->mn_major) {
    // a is [K, lda] (M contiguous), b is [K, ldb] (N contiguous): boxes of {64 MN elements, 64 K rows}
    if (g->pair || g->ep.loss_mode != SF_LOSS_NONE) return -7;
    if (g->bn < 64) g->bn = 64;
    if (g->bn > 128) g->bn = 128;
    rc = sf_make_tmap_bf16_kmajor(&g->tmA, g->a, g->K, g->M, g->lda, 64);
    if (rc) return rc;
    rc = sf_make_tmap_bf16_kmajor(&g->tmB, g->b, g->K, g->N, g->ldb, 64);
    if (rc) return rc;
  } else {
    rc = sf_make_tmap_bf16_kmajor(&g->tmA, g->a, g->M, g->K, g->lda, sf::kBM);
    if (rc) return rc;
    rc = sf_make_tmap_bf16_kmajor(&g->tmB, g->b, g->N, g->K, g->ldb, g->pair ? 128 : g->bn);
    if (rc) return rc;
  }
  const int total_kb = (g->K + sf::kBK - 1) / sf::kBK;
  if (g->split_k > total_kb) g->split_k = total_kb;
  g->kblocks_per_split = (total_kb + g->split_k - 1) / g->split_k;
  g->split_k = (total_kb + g->kblocks_per_split - 1) / g->kblocks_per_split;
  if (g->split_k > 1) g->ep.accumulate = 1;
  if (g->ep.loss_mode == SF_LOSS_SOFTMAX_XENT) {
    if (g->N > 32 || g->split_k > 1) return -5;     // the row must fit one epilogue chunk
    g->bn = 32;
  }
  if (g->ep.loss_mode != SF_LOSS_NONE && (g->split_k > 1 || !g->ep.target || !g->ep.loss)) return -5;
  // largest column any output can hold
  int lim = g->N;
  if (g->ep.out_bf16 && g->ep.ld_bf16 > lim) lim = g->ep.ld_bf16;
  g->ep.n_store_limit = lim;
  return 0;
}

template <int BN, bool MN = false>
static cudaError_t launch_bn(const SfGemm* g, cudaStream_t st) {
  using S = sf::GemmSmem<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(sf::sf_gemm_kernel<BN, MN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, S::kBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid((g->N + BN - 1) / BN, (g->M + sf::kBM - 1) / sf::kBM, g->split_k);
  return sf::launch(sf::sf_gemm_kernel<BN, MN>, grid, dim3(sf::kGemmThreads), S::kBytes, st, g->tmA, g->tmB, g->ep,
                    g->M, g->N, g->K, g->kblocks_per_split);
}

extern "C" int sf_gemm_launch(const SfGemm* g, cudaStream_t st) {
  if (g->pair) return sf_gemm_pair_launch(g, st);
  cudaError_t e;
  if (g->mn_major) {
    switch (g->bn) {
      case 64: e = launch_bn<64, true>(g, st); break;
      case 128: e = launch_bn<128, true>(g, st); break;
      default: return -3;
    }
    return static_cast<int>(e);
  }
  switch (g->bn) {
    case 32: e = launch_bn<32>(g, st); break;
    case 64: e = launch_bn<64>(g, st); break;
    case 128: e = launch_bn<128>(g, st); break;
    case 256: e = launch_bn<256>(g, st); break;
    default: return -3;
  }
  return static_cast<int>(e);
}

extern "C" void sf_set_pdl(int enabled) { sf::pdl_enabled() = enabled ? 1 : 0; }

// tracing: `buf` is a device buffer of `cap` 32-byte records (or nullptr to disable)
extern "C" int sf_trace_enable(void* buf, unsigned int cap) {
  sf::TraceRec* b = static_cast<sf::TraceRec*>(buf);
  unsigned int zero = 0;
  cudaError_t e = cudaMemcpyToSymbol(sf::g_trace_buf, &b, sizeof(b));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(sf::g_trace_cap, &cap, sizeof(cap));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(sf::g_trace_n, &zero, sizeof(zero));
  return static_cast<int>(e);
}
extern "C" unsigned int sf_trace_count() {
  unsigned int n = 0;
  cudaMemcpyFromSymbol(&n, sf::g_trace_n, sizeof(n));
  return n;
}

static unsigned int* g_err_host_ptr = nullptr;

// allocate the host-visible error word for the current device and point the device symbol at it
extern "C" int sf_init_error_channel() {
  if (g_err_host_ptr == nullptr) {
    void* raw = nullptr;
    cudaError_t e = cudaHostAlloc(&raw, sizeof(unsigned int), cudaHostAllocMapped | cudaHostAllocPortable);
    if (e != cudaSuccess) return static_cast<int>(e);
    g_err_host_ptr = static_cast<unsigned int*>(raw);
    *g_err_host_ptr = 0;
  }
  unsigned int* dptr = nullptr;
  cudaError_t e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&dptr), g_err_host_ptr, 0);
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(sf::g_sf_error_host, &dptr, sizeof(dptr));
  return static_cast<int>(e);
}
extern "C" unsigned int sf_read_host_error_code() { return g_err_host_ptr ? *g_err_host_ptr : 0u; }

extern "C" unsigned int sf_read_error_code() {
  unsigned int v = 0;
  cudaMemcpyFromSymbol(&v, sf::g_sf_error_code, sizeof(v));
  return v;
}
