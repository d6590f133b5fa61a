// K8: depthwise 3x3x3 convolution (groups = channels, stride 1, zero padding 1) -- the `conv_block` of every latent level,
// BasicConv(C, C, 3, padding=1, groups=C) (RA_HVED.py:406, buildingblocks.py:11-31), forward, input gradient and weight gradient.
// PyTorch runs these on its conv_depthwise3d kernels; with one volume per step and C = 4...32 their weight-gradient kernel takes
// 1.2 ms per call, 34 ms of a 117 ms training step once K6 / K7 are in (gpurun_out/r02x_model_kernels_patched_train.txt).
//
// HBM-bound (27 FMAs per 8 bytes moved): one CTA per (sample, channel, 8 x 8 x 32 tile), the (10, 10, 34) halo tile in shared
// memory, one thread per (h, w) column with its 8 outputs along d in registers.
//   forward   y = conv(x, w[c]) + b[c]
//   dgrad     dx = conv(dy, flip(w[c]))                                   -- the same kernel
//   wgrad     dw[c][tap] = sum_{n, v} dy[v] x[v + tap - 1], db[c] = sum dy  -- 27 (+1) accumulators per thread, one block reduction
//             per tile, per-CTA partial sums, then a reduction kernel over samples and tiles (deterministic, no atomics).
#include <cuda_runtime.h>
#include <stdint.h>

#include "xhved.h"

namespace {

constexpr int K = 3, TAPS = 27;
constexpr int TD = 8, TH = 8, TW = 32;
constexpr int HD = TD + 2, HH = TH + 2, HW = TW + 2;
constexpr int THREADS = TH * TW;
constexpr int PSTRIDE = 28;                     // 27 taps + the bias slot

struct Dims {
  int N, C, D, H, W;
  int tiles_d, tiles_h, tiles_w;
};

__device__ __forceinline__ void tile_origin(const Dims& s, int64_t& plane, int& d0, int& h0, int& w0) {
  int64_t t = blockIdx.x;
  w0 = static_cast<int>(t % s.tiles_w) * TW, t /= s.tiles_w;
  h0 = static_cast<int>(t % s.tiles_h) * TH, t /= s.tiles_h;
  d0 = static_cast<int>(t % s.tiles_d) * TD, plane = t / s.tiles_d;     // plane = n * C + c
}

// one thread per (h, w) column of the halo tile, walking along d: the h / w bounds and the address arithmetic are taken once per
// column, the HD loads of a column are independent (and consecutive lanes read consecutive w)
__device__ __forceinline__ void load_halo(float* __restrict__ tile, const float* __restrict__ src, const Dims& s, int d0, int h0, int w0) {
  const int64_t plane = static_cast<int64_t>(s.H) * s.W;
  for (int col = threadIdx.x; col < HH * HW; col += THREADS) {
    const int hy = col / HW, wx = col - hy * HW;
    const int h = h0 + hy - 1, w = w0 + wx - 1;
    const bool ok = h >= 0 && h < s.H && w >= 0 && w < s.W;
    const float* p = src + static_cast<int64_t>(ok ? h : 0) * s.W + (ok ? w : 0);
#pragma unroll
    for (int dz = 0; dz < HD; ++dz) {
      const int d = d0 + dz - 1;
      tile[dz * HH * HW + col] = (ok && d >= 0 && d < s.D) ? __ldg(p + d * plane) : 0.f;
    }
  }
}

// y = conv(x, w[c]) (+ bias[c]);  FLIP: correlation with the flipped kernel (the input gradient)
template <bool FLIP>
__global__ void __launch_bounds__(THREADS) dwconv3_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, Dims s, float* __restrict__ y) {
  __shared__ float tile[HD * HH * HW];
  __shared__ float wsm[TAPS];
  int64_t plane;
  int d0, h0, w0;
  tile_origin(s, plane, d0, h0, w0);
  const int c = static_cast<int>(plane % s.C);
  const int64_t vol = static_cast<int64_t>(s.D) * s.H * s.W;
  load_halo(tile, x + plane * vol, s, d0, h0, w0);
  if (threadIdx.x < TAPS) wsm[threadIdx.x] = w[c * TAPS + (FLIP ? TAPS - 1 - threadIdx.x : threadIdx.x)];
  __syncthreads();
  const int tx = threadIdx.x % TW, ty = threadIdx.x / TW;
  const float b = (!FLIP && bias) ? __ldg(bias + c) : 0.f;
  float acc[TD];
#pragma unroll
  for (int o = 0; o < TD; ++o) acc[o] = b;
#pragma unroll
  for (int kh = 0; kh < K; ++kh)
#pragma unroll
    for (int kw = 0; kw < K; ++kw) {
      const float* col = tile + (ty + kh) * HW + tx + kw;
      float in[HD];
#pragma unroll
      for (int i = 0; i < HD; ++i) in[i] = col[i * HH * HW];
#pragma unroll
      for (int kd = 0; kd < K; ++kd) {
        const float wk = wsm[(kd * K + kh) * K + kw];
#pragma unroll
        for (int o = 0; o < TD; ++o) acc[o] = fmaf(in[o + kd], wk, acc[o]);
      }
    }
  const int h = h0 + ty, wv = w0 + tx;
  if (h < s.H && wv < s.W) {
#pragma unroll
    for (int o = 0; o < TD; ++o)
      if (d0 + o < s.D) y[plane * vol + (static_cast<int64_t>(d0 + o) * s.H + h) * s.W + wv] = acc[o];
  }
}

// part[cta][28]: the tile's contribution to dw[c][0..26] and db[c]
__global__ void __launch_bounds__(THREADS) dwconv3_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, Dims s,
                                                               float* __restrict__ part) {
  __shared__ float tile[HD * HH * HW];
  __shared__ float red[THREADS / 32][PSTRIDE];
  int64_t plane;
  int d0, h0, w0;
  tile_origin(s, plane, d0, h0, w0);
  const int64_t vol = static_cast<int64_t>(s.D) * s.H * s.W;
  load_halo(tile, x + plane * vol, s, d0, h0, w0);
  const int tx = threadIdx.x % TW, ty = threadIdx.x / TW;
  const int h = h0 + ty, wv = w0 + tx;
  float g[TD], acc[PSTRIDE];
#pragma unroll
  for (int o = 0; o < TD; ++o)
    g[o] = (h < s.H && wv < s.W && d0 + o < s.D) ? __ldg(dy + plane * vol + (static_cast<int64_t>(d0 + o) * s.H + h) * s.W + wv) : 0.f;
  __syncthreads();
  acc[TAPS] = 0.f;
#pragma unroll
  for (int o = 0; o < TD; ++o) acc[TAPS] += g[o];
#pragma unroll
  for (int kh = 0; kh < K; ++kh)
#pragma unroll
    for (int kw = 0; kw < K; ++kw) {
      const float* col = tile + (ty + kh) * HW + tx + kw;
      float in[HD];
#pragma unroll
      for (int i = 0; i < HD; ++i) in[i] = col[i * HH * HW];
#pragma unroll
      for (int kd = 0; kd < K; ++kd) {
        float a = 0.f;
#pragma unroll
        for (int o = 0; o < TD; ++o) a = fmaf(g[o], in[o + kd], a);
        acc[(kd * K + kh) * K + kw] = a;
      }
    }
#pragma unroll
  for (int i = 0; i < PSTRIDE; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < PSTRIDE; ++i) red[warp][i] = acc[i];
  }
  __syncthreads();
  if (threadIdx.x < PSTRIDE) {
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < THREADS / 32; ++q) v += red[q][threadIdx.x];
    part[static_cast<int64_t>(blockIdx.x) * PSTRIDE + threadIdx.x] = v;
  }
}

// dw[c][i] (i < 27) and db[c] (i == 27) = sum over samples and tiles of part[((n * C + c) * tiles + t)][i]; grid = C, block = 28 x 8
__global__ void __launch_bounds__(PSTRIDE * 8) dwconv3_reduce_kernel(const float* __restrict__ part, int N, int C, int tiles,
                                                                    float* __restrict__ dw, float* __restrict__ db) {
  __shared__ double red[8][PSTRIDE];
  const int c = blockIdx.x, i = threadIdx.x % PSTRIDE, q = threadIdx.x / PSTRIDE;
  double v = 0.0;
  for (int n = 0; n < N; ++n) {
    const float* p = part + (static_cast<int64_t>(n) * C + c) * tiles * PSTRIDE;
    for (int t = q; t < tiles; t += 8) v += static_cast<double>(p[static_cast<int64_t>(t) * PSTRIDE + i]);
  }
  red[q][i] = v;
  __syncthreads();
  if (q == 0) {
    double t = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += red[r][i];
    if (i < TAPS) dw[c * TAPS + i] = static_cast<float>(t);
    else if (db) db[c] = static_cast<float>(t);
  }
}

int make_dims(int N, int C, int D, int H, int W, Dims& s, int64_t& ctas) {
  if (N <= 0 || C <= 0 || D <= 0 || H <= 0 || W <= 0) return XHVED_ERR_BAD_ARG;
  s.N = N, s.C = C, s.D = D, s.H = H, s.W = W;
  s.tiles_d = (D + TD - 1) / TD, s.tiles_h = (H + TH - 1) / TH, s.tiles_w = (W + TW - 1) / TW;
  ctas = static_cast<int64_t>(N) * C * s.tiles_d * s.tiles_h * s.tiles_w;
  return ctas > 0x7fffffffLL ? XHVED_ERR_BAD_SHAPE : 0;
}

}  // namespace

extern "C" int64_t xhved_dwconv3_workspace(int N, int C, int D, int H, int W) {
  Dims s;
  int64_t ctas;
  if (const int rc = make_dims(N, C, D, H, W, s, ctas)) return rc;
  return ctas * PSTRIDE * static_cast<int64_t>(sizeof(float));
}

extern "C" int xhved_dwconv3_fwd(const float* x, const float* w, const float* bias, int N, int C, int D, int H, int W, float* y,
                                 void* stream) {
  Dims s;
  int64_t ctas;
  if (const int rc = make_dims(N, C, D, H, W, s, ctas)) return rc;
  if (!x || !w || !y) return XHVED_ERR_BAD_ARG;
  dwconv3_kernel<false><<<static_cast<unsigned>(ctas), THREADS, 0, static_cast<cudaStream_t>(stream)>>>(x, w, bias, s, y);
  return (int)cudaGetLastError();
}

extern "C" int xhved_dwconv3_bwd(const float* x, const float* w, const float* dy, int N, int C, int D, int H, int W, void* partials,
                                 float* dx, float* dw, float* dbias, void* stream) {
  Dims s;
  int64_t ctas;
  if (const int rc = make_dims(N, C, D, H, W, s, ctas)) return rc;
  if (!dy || (dx && !w) || (dw && (!x || !partials))) return XHVED_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dx) dwconv3_kernel<true><<<static_cast<unsigned>(ctas), THREADS, 0, st>>>(dy, w, nullptr, s, dx);
  if (dw) {
    dwconv3_wgrad_kernel<<<static_cast<unsigned>(ctas), THREADS, 0, st>>>(x, dy, s, static_cast<float*>(partials));
    dwconv3_reduce_kernel<<<C, PSTRIDE * 8, 0, st>>>(static_cast<const float*>(partials), N, C, s.tiles_d * s.tiles_h * s.tiles_w, dw, dbias);
  }
  return (int)cudaGetLastError();
}
