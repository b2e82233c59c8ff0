// Convolution building blocks for the compiled CNN plan (reference ops K3/K4 of SURVEY.md section 2.4:
// Conv2D + BiasAdd + Relu, MaxPool and their gradients; examples/cnn_example.py:10-22).
//
// Convolutions are lowered to the tcgen05 GEMM of gemm_sm100.cu:
//   forward   out[M, Cout]   = patches[M, K] . W^T[Cout, K]^T  (+bias, ReLU in the GEMM epilogue)
//   wgrad     dW[K, Cout]    = patches^T[K, M] . dz^T[Cout, M]^T      (split-K, atomic accumulate)
//   dgrad     dpatch[M, K]   = dz[M, Cout] . W[K, Cout]^T             -> col2im gather
// with M = batch * out_h * out_w and K = kh * kw * Cin in (kh, kw, cin) order, i.e. exactly the flattening
// of TensorFlow's HWIO filter layout.  The kernels here produce the K-major operands those GEMMs need:
// im2col writes patches AND patches^T; max-pool backward applies the ReLU derivative, scatters through
// the arg-max mask and writes dz, dz^T and the bias gradient in one pass.
#include "sm100_ptx.cuh"
#include "sf_api.h"

namespace sf {

// ------------------------------------------------------------------------------------------------
// im2col (NHWC, VALID, stride 1): out[m, k] = in[b, oh+kh, ow+kw, c];  32x32 tiles through smem so
// that the row-major store (along k) and the transposed store (along m) are both coalesced.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
im2col_kernel(const __nv_bfloat16* __restrict__ in, int n, int h, int w, int c, int kh, int kw,
              __nv_bfloat16* __restrict__ out, int ld_out, __nv_bfloat16* __restrict__ outT, int ld_t) {
  __shared__ __nv_bfloat16 tile[32][34];
  TraceScope trace;
  pdl_launch_dependents();
  pdl_wait();
  trace.mark();
  const int oh = h - kh + 1, ow = w - kw + 1;
  const int M = n * oh * ow, K = kh * kw * c;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int k0 = blockIdx.y * 32, m0 = blockIdx.x * 32;      // M tiles on x: a 4096-row evaluation chunk has > 65535 of them
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty + 8 * i, k = k0 + tx;
    __nv_bfloat16 v = __float2bfloat16(0.f);
    if (m < M && k < K) {
      const int b = m / (oh * ow), r = m - b * (oh * ow);
      const int y = r / ow, x = r - y * ow;
      const int cc = k % c, t = k / c;
      const int dx = t % kw, dy = t / kw;
      v = in[((static_cast<size_t>(b) * h + (y + dy)) * w + (x + dx)) * c + cc];
    }
    tile[ty + 8 * i][tx] = v;
    if (out != nullptr && m < M && k < ld_out) out[static_cast<size_t>(m) * ld_out + k] = v;
  }
  if (outT != nullptr) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + ty + 8 * i, m = m0 + tx;
      if (k < K && m < M) outT[static_cast<size_t>(k) * ld_t + m] = tile[tx][ty + 8 * i];
    }
  }
  trace.end(KID_IM2COL);
}

// ------------------------------------------------------------------------------------------------
// col2im as a gather (no atomics): din[b, y, x, c] = sum_{dy,dx} dcols[(b, y-dy, x-dx), (dy*kw+dx)*C + c]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
col2im_kernel(const __nv_bfloat16* __restrict__ dcols, int ld_cols, int n, int h, int w, int c, int kh, int kw,
              __nv_bfloat16* __restrict__ din) {
  TraceScope trace;
  pdl_launch_dependents();
  pdl_wait();
  trace.mark();
  const int oh = h - kh + 1, ow = w - kw + 1;
  const size_t total = static_cast<size_t>(n) * h * w * c;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cc = static_cast<int>(i % c);
    size_t t = i / c;
    const int x = static_cast<int>(t % w);
    t /= w;
    const int y = static_cast<int>(t % h);
    const int b = static_cast<int>(t / h);
    float acc = 0.f;
    for (int dy = 0; dy < kh; ++dy) {
      const int yy = y - dy;
      if (yy < 0 || yy >= oh) continue;
      for (int dx = 0; dx < kw; ++dx) {
        const int xx = x - dx;
        if (xx < 0 || xx >= ow) continue;
        const size_t m = (static_cast<size_t>(b) * oh + yy) * ow + xx;
        acc += __bfloat162float(dcols[m * ld_cols + (dy * kw + dx) * c + cc]);
      }
    }
    din[i] = __float2bfloat16(acc);
  }
  trace.end(KID_COL2IM);
}

// ------------------------------------------------------------------------------------------------
// 2x2 / stride-2 max pool (VALID).  argmax stores the winner's position inside the window (0..3).
// Optional outT[(y*ow + x)*C + c][b] is the transposed flattened activation (K-major operand of the
// following dense layer's wgrad).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ in, int n, int h, int w, int c, __nv_bfloat16* __restrict__ out,
                   uint8_t* __restrict__ argmax, __nv_bfloat16* __restrict__ outT, int ld_t) {
  TraceScope trace;
  pdl_launch_dependents();
  pdl_wait();
  trace.mark();
  const int oh = h / 2, ow = w / 2;
  const size_t total = static_cast<size_t>(n) * oh * ow * c;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cc = static_cast<int>(i % c);
    size_t t = i / c;
    const int x = static_cast<int>(t % ow);
    t /= ow;
    const int y = static_cast<int>(t % oh);
    const int b = static_cast<int>(t / oh);
    const __nv_bfloat16* p = in + ((static_cast<size_t>(b) * h + 2 * y) * w + 2 * x) * c + cc;
    float best = __bfloat162float(p[0]);
    int bi = 0;
    const float v1 = __bfloat162float(p[c]);
    const float v2 = __bfloat162float(p[static_cast<size_t>(w) * c]);
    const float v3 = __bfloat162float(p[static_cast<size_t>(w) * c + c]);
    if (v1 > best) { best = v1; bi = 1; }
    if (v2 > best) { best = v2; bi = 2; }
    if (v3 > best) { best = v3; bi = 3; }
    out[i] = __float2bfloat16(best);
    argmax[i] = static_cast<uint8_t>(bi);
    if (outT != nullptr) outT[static_cast<size_t>((y * ow + x) * c + cc) * ld_t + b] = __float2bfloat16(best);
  }
  trace.end(KID_POOL_FWD);
}

// ------------------------------------------------------------------------------------------------
// max-pool backward fused with the activation derivative of the conv output it pooled:
//   dz[b, y, x, c] = (argmax(window) == position) ? dout[window] * act'(a[b, y, x, c]) : 0
// written row-major [M, ld_dz] (dgrad A operand), transposed [C, ld_t] (wgrad B operand) and reduced
// over M into dbias[C].  One 32(m) x 32(c) tile per CTA.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float pool_dact(float a, int act) {
  switch (act) {
    case SF_ACT_RELU: return a > 0.f ? 1.f : 0.f;
    case SF_ACT_SIGMOID: return a * (1.f - a);
    case SF_ACT_TANH: return 1.f - a * a;
    default: return 1.f;
  }
}

__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const uint8_t* __restrict__ argmax, int n, int h, int w, int c,
                   const __nv_bfloat16* __restrict__ act_out, int act, __nv_bfloat16* __restrict__ dz, int ld_dz,
                   __nv_bfloat16* __restrict__ dzT, int ld_t, float* __restrict__ dbias) {
  __shared__ __nv_bfloat16 tile[32][34];
  __shared__ float s_col[32];
  TraceScope trace;
  pdl_launch_dependents();
  pdl_wait();
  trace.mark();
  const int oh = h / 2, ow = w / 2;
  const int M = n * h * w;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  if (threadIdx.x < 32) s_col[threadIdx.x] = 0.f;
  __syncthreads();
  float cacc = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty + 8 * i, cc = c0 + tx;
    float g = 0.f;
    if (m < M && cc < c) {
      const int b = m / (h * w), r = m - b * (h * w);
      const int y = r / w, x = r - y * w;
      const int py = y >> 1, px = x >> 1;
      if (py < oh && px < ow) {
        const size_t o = ((static_cast<size_t>(b) * oh + py) * ow + px) * c + cc;
        if (argmax[o] == ((y & 1) * 2 + (x & 1)))
          g = __bfloat162float(dout[o]) * pool_dact(__bfloat162float(act_out[static_cast<size_t>(m) * c + cc]), act);
      }
    }
    const __nv_bfloat16 gb = __float2bfloat16(g);
    tile[ty + 8 * i][tx] = gb;
    cacc += g;
    if (dz != nullptr && m < M && cc < ld_dz) dz[static_cast<size_t>(m) * ld_dz + cc] = gb;
  }
  if (dbias != nullptr) atomicAdd(&s_col[tx], cacc);
  __syncthreads();
  if (dbias != nullptr && threadIdx.x < 32 && c0 + threadIdx.x < c && s_col[threadIdx.x] != 0.f)
    atomicAdd(dbias + c0 + threadIdx.x, s_col[threadIdx.x]);
  if (dzT != nullptr) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int cc = c0 + ty + 8 * i, m = m0 + tx;
      if (cc < c && m < M) dzT[static_cast<size_t>(cc) * ld_t + m] = tile[tx][ty + 8 * i];
    }
  }
  trace.end(KID_POOL_BWD);
}

}  // namespace sf

static inline int grid_1d(size_t total) {
  size_t b = (total + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  return static_cast<int>(b < 1 ? 1 : b);
}

extern "C" int sf_im2col_nhwc(const __nv_bfloat16* in, int n, int h, int w, int c, int kh, int kw, __nv_bfloat16* out,
                              int ld_out, __nv_bfloat16* outT, int ld_t, cudaStream_t st) {
  const int M = n * (h - kh + 1) * (w - kw + 1), K = kh * kw * c;
  const int span_k = (out && ld_out > K) ? ld_out : K;
  dim3 grid((M + 31) / 32, (span_k + 31) / 32);
  return static_cast<int>(sf::launch(sf::im2col_kernel, grid, dim3(256), 0, st, in, n, h, w, c, kh, kw, out, ld_out, outT, ld_t));
}

extern "C" int sf_col2im_nhwc(const __nv_bfloat16* dcols, int ld_cols, int n, int h, int w, int c, int kh, int kw,
                              __nv_bfloat16* din, cudaStream_t st) {
  const size_t total = static_cast<size_t>(n) * h * w * c;
  return static_cast<int>(sf::launch(sf::col2im_kernel, dim3(grid_1d(total)), dim3(256), 0, st, dcols, ld_cols, n, h, w, c, kh, kw, din));
}

extern "C" int sf_maxpool2_fwd(const __nv_bfloat16* in, int n, int h, int w, int c, __nv_bfloat16* out, uint8_t* argmax,
                               __nv_bfloat16* outT, int ld_t, cudaStream_t st) {
  const size_t total = static_cast<size_t>(n) * (h / 2) * (w / 2) * c;
  return static_cast<int>(sf::launch(sf::maxpool_fwd_kernel, dim3(grid_1d(total)), dim3(256), 0, st, in, n, h, w, c, out, argmax, outT, ld_t));
}

extern "C" int sf_maxpool2_bwd(const __nv_bfloat16* dout, const uint8_t* argmax, int n, int h, int w, int c,
                               const __nv_bfloat16* act_out, int act, __nv_bfloat16* dz, int ld_dz, __nv_bfloat16* dzT, int ld_t,
                               float* dbias, cudaStream_t st) {
  const int M = n * h * w;
  const int span_c = (dz && ld_dz > c) ? ld_dz : c;
  dim3 grid((span_c + 31) / 32, (M + 31) / 32);
  return static_cast<int>(sf::launch(sf::maxpool_bwd_kernel, grid, dim3(256), 0, st, dout, argmax, n, h, w, c, act_out, act, dz,
                                     ld_dz, dzT, ld_t, dbias));
}
