// K7: the spatial gate of AttenModule2 (buildingblocks.py:259-301) -- a depthwise 7x7x7 convolution with channel expansion 4
// followed by a 1x1x1 convolution to ONE channel and a sigmoid (enc_spatial -> enc_spatial2, seg_spatial -> seg_spatial2,
// buildingblocks.py:271-274, 283-285, 294-296).  The two convolutions are linear maps applied back to back, so they compose into
// one dense G -> 1 convolution with W[g] = sum_j w2[4g+j] * W1[4g+j] (the host side builds W and the bias with two small torch
// ops and differentiates through them): a sixteenth of the multiplications and no 16-channel intermediate.  After K6 these two
// layers were the largest item of the whole model (PyTorch's conv_depthwise3d kernels: 6.3 ms forward and 35 ms backward for the
// (1, 4, 128^3) gate, cuDNN's dense 4 -> 1 form 12 ms: profiles/r02w_*).
//
// Direct convolution on the CUDA cores (K = 343 G taps per output, nothing a tensor core tile would fit): one CTA per 8 x 8 x 32
// output tile, the (14, 14, 38) halo tile of one channel in shared memory, one thread per (h, w) column with the 8 outputs along d
// in registers: per (kh, kw) it reads 14 inputs and 7 weights (two broadcast 128-bit loads) for 56 FMAs.
//   forward:  gate = sigmoid(conv(x, W) + b)
//   backward: dpre = dgate * gate * (1 - gate);  dx[g] = conv(dpre, flip(W[g])) -- the same kernel, one input and G outputs;
//             dW[g][tap] = sum_v dpre[v] x[g][v + tap - 3] -- per CTA tile: thread = (kd, kh) x row group, seven taps along kw
//             per thread, a sliding window over w; per-CTA partial sums, then one reduction kernel (deterministic, no atomics).
#include <cuda_runtime.h>
#include <stdint.h>

#include "xhved.h"

namespace {

constexpr int K = 7, R = 3, TAPS = K * K * K;
constexpr int TD = 8, TH = 8, TW = 32;
constexpr int HD = TD + K - 1, HH = TH + K - 1, HW = TW + K - 1;      // halo tile 14 x 14 x 38
constexpr int THREADS = TH * TW;
constexpr int WPAD = K * K * 8;                                       // weights of one channel as [kh][kw][kd padded to 8]

struct Dims {
  int N, G, D, H, W;
  int tiles_d, tiles_h, tiles_w;
};

__device__ __forceinline__ void tile_origin(const Dims& s, int& n, int& d0, int& h0, int& w0) {
  int t = blockIdx.x;
  w0 = (t % s.tiles_w) * TW, t /= s.tiles_w;
  h0 = (t % s.tiles_h) * TH, t /= s.tiles_h;
  d0 = (t % s.tiles_d) * TD, n = t / s.tiles_d;
}

// halo tile of one (sample, channel) volume, zero outside the volume.  DPRE: the value is dgate * gate * (1 - gate).
// One thread per (h, w) column walking along d: bounds and address arithmetic once per column, HD independent loads.
template <bool DPRE>
__device__ __forceinline__ void load_halo(float* __restrict__ tile, const float* __restrict__ a, const float* __restrict__ b, const Dims& s,
                                          int d0, int h0, int w0) {
  const int64_t plane = static_cast<int64_t>(s.H) * s.W;
  for (int col = threadIdx.x; col < HH * HW; col += THREADS) {
    const int hy = col / HW, wx = col - hy * HW;
    const int h = h0 + hy - R, w = w0 + wx - R;
    const bool ok = h >= 0 && h < s.H && w >= 0 && w < s.W;
    const int64_t base = static_cast<int64_t>(ok ? h : 0) * s.W + (ok ? w : 0);
#pragma unroll
    for (int dz = 0; dz < HD; ++dz) {
      const int d = d0 + dz - R;
      float v = 0.f;
      if (ok && d >= 0 && d < s.D) {
        if constexpr (DPRE) {
          const float g = __ldg(a + base + d * plane);
          v = __ldg(b + base + d * plane) * g * (1.f - g);
        } else {
          v = __ldg(a + base + d * plane);
        }
      }
      tile[dz * HH * HW + col] = v;
    }
  }
}

// weights of channel g into [kh][kw][kd(8)]; FLIP: the tap (6-kd, 6-kh, 6-kw) instead (correlation with the flipped kernel)
template <bool FLIP>
__device__ __forceinline__ void load_weights(float* __restrict__ wsm, const float* __restrict__ w, int g) {
  for (int i = threadIdx.x; i < WPAD; i += THREADS) {
    const int kd = i & 7, kk = i >> 3, kh = kk / K, kw = kk - kh * K;
    float v = 0.f;
    if (kd < K) {
      const int tap = FLIP ? ((K - 1 - kd) * K + (K - 1 - kh)) * K + (K - 1 - kw) : (kd * K + kh) * K + kw;
      v = w[g * TAPS + tap];
    }
    wsm[i] = v;
  }
}

__device__ __forceinline__ void accumulate(const float* __restrict__ tile, const float* __restrict__ wsm, int ty, int tx, float (&acc)[TD]) {
#pragma unroll 1
  for (int kh = 0; kh < K; ++kh) {
#pragma unroll
    for (int kw = 0; kw < K; ++kw) {
      const float* col = tile + (ty + kh) * HW + tx + kw;
      float in[HD];
#pragma unroll
      for (int i = 0; i < HD; ++i) in[i] = col[i * HH * HW];
      const float4 wa = *reinterpret_cast<const float4*>(wsm + (kh * K + kw) * 8);
      const float4 wb = *reinterpret_cast<const float4*>(wsm + (kh * K + kw) * 8 + 4);
      const float wk[K] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z};
#pragma unroll
      for (int o = 0; o < TD; ++o)
#pragma unroll
        for (int kd = 0; kd < K; ++kd) acc[o] = fmaf(in[o + kd], wk[kd], acc[o]);
    }
  }
}

// gate = sigmoid(sum_g conv(x[g], W[g]) + bias)
__global__ void __launch_bounds__(THREADS) gate7_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ bias, Dims s, float* __restrict__ gate) {
  __shared__ __align__(16) float tile[HD * HH * HW];
  __shared__ __align__(16) float wsm[WPAD];
  int n, d0, h0, w0;
  tile_origin(s, n, d0, h0, w0);
  const int tx = threadIdx.x % TW, ty = threadIdx.x / TW;
  const int64_t vol = static_cast<int64_t>(s.D) * s.H * s.W;
  float acc[TD];
#pragma unroll
  for (int o = 0; o < TD; ++o) acc[o] = 0.f;
  for (int g = 0; g < s.G; ++g) {
    __syncthreads();
    load_halo<false>(tile, x + (static_cast<int64_t>(n) * s.G + g) * vol, nullptr, s, d0, h0, w0);
    load_weights<false>(wsm, w, g);
    __syncthreads();
    accumulate(tile, wsm, ty, tx, acc);
  }
  const float b = bias ? __ldg(bias) : 0.f;
  const int h = h0 + ty, wv = w0 + tx;
  if (h < s.H && wv < s.W) {
#pragma unroll
    for (int o = 0; o < TD; ++o)
      if (d0 + o < s.D) gate[static_cast<int64_t>(n) * vol + (static_cast<int64_t>(d0 + o) * s.H + h) * s.W + wv] = 1.f / (1.f + __expf(-(acc[o] + b)));
  }
}

// dx[g] = conv(dpre, flip(W[g])), dpre = dgate * gate * (1 - gate)
__global__ void __launch_bounds__(THREADS) gate7_dgrad_kernel(const float* __restrict__ gate, const float* __restrict__ dgate,
                                                             const float* __restrict__ w, Dims s, float* __restrict__ dx) {
  __shared__ __align__(16) float tile[HD * HH * HW];
  __shared__ __align__(16) float wsm[WPAD];
  int n, d0, h0, w0;
  tile_origin(s, n, d0, h0, w0);
  const int tx = threadIdx.x % TW, ty = threadIdx.x / TW;
  const int64_t vol = static_cast<int64_t>(s.D) * s.H * s.W;
  load_halo<true>(tile, gate + n * vol, dgate + n * vol, s, d0, h0, w0);
  const int h = h0 + ty, wv = w0 + tx;
  for (int g = 0; g < s.G; ++g) {
    __syncthreads();                       // tile complete (first pass) / wsm no longer read (later passes)
    load_weights<true>(wsm, w, g);
    __syncthreads();
    float acc[TD];
#pragma unroll
    for (int o = 0; o < TD; ++o) acc[o] = 0.f;
    accumulate(tile, wsm, ty, tx, acc);
    if (h < s.H && wv < s.W) {
#pragma unroll
      for (int o = 0; o < TD; ++o)
        if (d0 + o < s.D) dx[(static_cast<int64_t>(n) * s.G + g) * vol + (static_cast<int64_t>(d0 + o) * s.H + h) * s.W + wv] = acc[o];
    }
  }
}

// per-CTA partial sums of dW[g][tap] and of dbias: part[cta][G * 343 + 1]
constexpr int PAIRS = K * K;                         // (kd, kh)
constexpr int GROUPS = THREADS / PAIRS;              // 5 row groups (245 of the 256 threads take part)
constexpr int ROWS = TD * TH;

__global__ void __launch_bounds__(THREADS) gate7_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ gate,
                                                             const float* __restrict__ dgate, Dims s, float* __restrict__ part) {
  __shared__ __align__(16) float tile[HD * HH * HW];
  __shared__ float dsm[ROWS * TW];
  __shared__ float red[GROUPS * PAIRS * K];
  __shared__ float bsum[THREADS / 32];
  int n, d0, h0, w0;
  tile_origin(s, n, d0, h0, w0);
  const int64_t vol = static_cast<int64_t>(s.D) * s.H * s.W;
  const int stride = s.G * TAPS + 1;
  float* out = part + static_cast<int64_t>(blockIdx.x) * stride;
  // dpre of the tile's own outputs (zero outside the volume) and its sum
  float mine = 0.f;
  for (int i = threadIdx.x; i < ROWS * TW; i += THREADS) {
    const int r = i / TW, wx = i - r * TW, d = d0 + r / TH, h = h0 + r % TH, wv = w0 + wx;
    float v = 0.f;
    if (d < s.D && h < s.H && wv < s.W) {
      const int64_t at = n * vol + (static_cast<int64_t>(d) * s.H + h) * s.W + wv;
      const float g = __ldg(gate + at);
      v = __ldg(dgate + at) * g * (1.f - g);
    }
    dsm[i] = v;
    mine += v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0) bsum[threadIdx.x >> 5] = mine;
  const int group = threadIdx.x / PAIRS, pair = threadIdx.x - group * PAIRS, kd = pair / K, kh = pair - kd * K;
  for (int g = 0; g < s.G; ++g) {
    __syncthreads();                       // dsm / bsum written; tile and red of the previous channel no longer read
    load_halo<false>(tile, x + (static_cast<int64_t>(n) * s.G + g) * vol, nullptr, s, d0, h0, w0);
    __syncthreads();
    float acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.f;
    if (group < GROUPS) {
      for (int r = group; r < ROWS; r += GROUPS) {
        const float* xr = tile + ((r / TH + kd) * HH + (r % TH) + kh) * HW;
        const float* dr = dsm + r * TW;
        float win[K];                      // x[w .. w + 6]
#pragma unroll
        for (int k = 0; k < K - 1; ++k) win[k + 1] = xr[k];
#pragma unroll
        for (int wx = 0; wx < TW; ++wx) {
#pragma unroll
          for (int k = 0; k < K - 1; ++k) win[k] = win[k + 1];
          win[K - 1] = xr[wx + K - 1];
          const float dv = dr[wx];
#pragma unroll
          for (int k = 0; k < K; ++k) acc[k] = fmaf(dv, win[k], acc[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < K; ++k) red[(group * PAIRS + pair) * K + k] = acc[k];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TAPS; i += THREADS) {          // i = (kd * 7 + kh) * 7 + kw = pair * 7 + kw
      float v = 0.f;
#pragma unroll
      for (int q = 0; q < GROUPS; ++q) v += red[q * PAIRS * K + i];
      out[g * TAPS + i] = v;
    }
  }
  if (threadIdx.x == 0) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < THREADS / 32; ++i) v += bsum[i];
    out[s.G * TAPS] = v;
  }
}

// dW[i] = sum over CTAs of part[cta][i]  (i < G * 343: weights, i == G * 343: bias); one CTA per 32 columns, fp64 accumulation
__global__ void __launch_bounds__(256) gate7_reduce_kernel(const float* __restrict__ part, int ctas, int stride, float* __restrict__ dw,
                                                          float* __restrict__ dbias) {
  __shared__ double red[8][32];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31), lane_row = threadIdx.x >> 5;
  double v = 0.0;
  if (col < stride)
    for (int r = lane_row; r < ctas; r += 8) v += static_cast<double>(part[static_cast<int64_t>(r) * stride + col]);
  red[lane_row][threadIdx.x & 31] = v;
  __syncthreads();
  if (threadIdx.x < 32 && col < stride) {
    double t = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += red[r][threadIdx.x];
    if (col < stride - 1) dw[col] = static_cast<float>(t);
    else if (dbias) *dbias = static_cast<float>(t);
  }
}

int make_dims(int N, int G, int D, int H, int W, Dims& s, int64_t& ctas) {
  if (N <= 0 || G <= 0 || D <= 0 || H <= 0 || W <= 0) return XHVED_ERR_BAD_ARG;
  s.N = N, s.G = G, s.D = D, s.H = H, s.W = W;
  s.tiles_d = (D + TD - 1) / TD, s.tiles_h = (H + TH - 1) / TH, s.tiles_w = (W + TW - 1) / TW;
  ctas = static_cast<int64_t>(N) * s.tiles_d * s.tiles_h * s.tiles_w;
  return ctas > 0x7fffffffLL ? XHVED_ERR_BAD_SHAPE : 0;
}

}  // namespace

extern "C" int64_t xhved_gate7_workspace(int N, int G, int D, int H, int W) {
  Dims s;
  int64_t ctas;
  if (const int rc = make_dims(N, G, D, H, W, s, ctas)) return rc;
  return ctas * (static_cast<int64_t>(G) * TAPS + 1) * static_cast<int64_t>(sizeof(float));
}

extern "C" int xhved_gate7_fwd(const float* x, const float* w, const float* bias, int N, int G, int D, int H, int W, float* gate,
                               void* stream) {
  Dims s;
  int64_t ctas;
  if (const int rc = make_dims(N, G, D, H, W, s, ctas)) return rc;
  if (!x || !w || !gate) return XHVED_ERR_BAD_ARG;
  gate7_fwd_kernel<<<static_cast<unsigned>(ctas), THREADS, 0, static_cast<cudaStream_t>(stream)>>>(x, w, bias, s, gate);
  return (int)cudaGetLastError();
}

extern "C" int xhved_gate7_bwd(const float* x, const float* w, const float* gate, const float* dgate, int N, int G, int D, int H, int W,
                               void* partials, float* dx, float* dw, float* dbias, void* stream) {
  Dims s;
  int64_t ctas;
  if (const int rc = make_dims(N, G, D, H, W, s, ctas)) return rc;
  if (!x || !w || !gate || !dgate) return XHVED_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dx) gate7_dgrad_kernel<<<static_cast<unsigned>(ctas), THREADS, 0, st>>>(gate, dgate, w, s, dx);
  if (dw) {
    if (!partials) return XHVED_ERR_BAD_ARG;
    const int stride = G * TAPS + 1;
    gate7_wgrad_kernel<<<static_cast<unsigned>(ctas), THREADS, 0, st>>>(x, gate, dgate, s, static_cast<float*>(partials));
    gate7_reduce_kernel<<<(stride + 31) / 32, 256, 0, st>>>(static_cast<const float*>(partials), static_cast<int>(ctas), stride, dw, dbias);
  }
  return (int)cudaGetLastError();
}
