// K10: dense 3x3x3 convolution with few channels (stride 1, zero padding 1, one group; Cin, Cout <= 64) -- the SingleConv /
// DoubleConv layers of the encoders and decoders (buildingblocks.py:444-507, RA_HVED.py:340-400) at 4 ... 48 channels.  cuDNN
// serves them with implicit-GEMM tensor-core kernels built for hundreds of channels: under the reference's fp16 autocast
// (train.py:207) the forward of a (1, 4, 128^3) 4 -> 4 layer takes 1.7 ms (fp32: 0.05 ms), plus NCHW <-> NHWC transposes around
// every call -- 100 of the 137 ms of GPU time left in the training step after K6-K9 (profiles/r02z_train_profile.txt).
//
// Direct convolution on the CUDA cores, the K8 tiling with a loop over input channels: CTA = (sample, 8 x 8 x 32 output tile,
// group of CO_T output channels); per input channel the (10, 10, 34) halo tile goes to shared memory as fp32 together with the
// 27 x CO_T weights of that channel; thread = (h, w) column, CO_T x 8 accumulators.  fp32 / fp16 / bf16 tensors, fp32 accumulation.
//   forward  y[o] = b[o] + sum_i conv(x[i], W[o][i])
//   dgrad    dx[i] = sum_o conv(dy[o], flip(W[o][i]))       -- the same kernel with the channel roles swapped
//   wgrad    dW[o][i][tap] = sum_v dy[o][v] x[i][v + tap - 1]: CTA = (sample, four tiles along d, input channel, four output
//            channels): 4 x 28 accumulators per thread over the four tiles, ONE block reduction, per-CTA partials and a reduction
//            kernel (deterministic); db[o] rides along in the CTAs of input channel 0.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "xhved.h"

namespace {

constexpr int K = 3, TAPS = 27;
constexpr int TD = 8, TH = 8, TW = 32;
constexpr int HD = TD + 2, HH = TH + 2, HW = TW + 2;
constexpr int THREADS = TH * TW;
constexpr int WG_CO = 2, WG_TILES = 4, PST = 28;

template <typename T> struct Cvt;
template <> struct Cvt<float> {
  static __device__ __forceinline__ float to_f(float v) { return v; }
  static __device__ __forceinline__ float from_f(float v) { return v; }
};
template <> struct Cvt<__half> {
  static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
};
template <> struct Cvt<__nv_bfloat16> {
  static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};

struct Dims {
  int N, Cin, Cout, D, H, W;
  int tiles_d, tiles_h, tiles_w;
};

// one thread per (h, w) column of the halo tile, walking along d: the h / w bounds and the address arithmetic are taken once per
// column, the HD loads of a column are independent (and consecutive lanes read consecutive w)
template <typename T>
__device__ __forceinline__ void load_halo(float* __restrict__ tile, const T* __restrict__ src, const Dims& s, int d0, int h0, int w0) {
  const int64_t plane = static_cast<int64_t>(s.H) * s.W;
  for (int col = threadIdx.x; col < HH * HW; col += THREADS) {
    const int hy = col / HW, wx = col - hy * HW;
    const int h = h0 + hy - 1, w = w0 + wx - 1;
    const bool ok = h >= 0 && h < s.H && w >= 0 && w < s.W;
    const T* p = src + static_cast<int64_t>(ok ? h : 0) * s.W + (ok ? w : 0);
#pragma unroll
    for (int dz = 0; dz < HD; ++dz) {
      const int d = d0 + dz - 1;
      tile[dz * HH * HW + col] = (ok && d >= 0 && d < s.D) ? Cvt<T>::to_f(p[d * plane]) : 0.f;
    }
  }
}

// out[n][o][v] = bias[o] + sum_i sum_tap in[n][i][v + tap - 1] w[(o * w_so + i * w_si) * 27 + (FLIP ? 26 - tap : tap)]
// `cin` / `cout` are the channel counts of `in` / `out` (swapped for the input gradient).
template <typename T, int CO_T, bool FLIP>
__global__ void __launch_bounds__(THREADS) conv3_kernel(const T* __restrict__ in, const float* __restrict__ w, int w_so, int w_si,
                                                       const float* __restrict__ bias, Dims s, int cin, int cout, T* __restrict__ out) {
  __shared__ float tile[HD * HH * HW];
  __shared__ __align__(16) float wsm[TAPS * CO_T];          // [tap][o]
  int t = blockIdx.x;
  const int w0 = (t % s.tiles_w) * TW;
  t /= s.tiles_w;
  const int h0 = (t % s.tiles_h) * TH;
  t /= s.tiles_h;
  const int d0 = (t % s.tiles_d) * TD, n = t / s.tiles_d;
  const int o0 = blockIdx.y * CO_T;
  const int64_t vol = static_cast<int64_t>(s.D) * s.H * s.W;
  const int tx = threadIdx.x % TW, ty = threadIdx.x / TW;
  float acc[CO_T][TD];
#pragma unroll
  for (int o = 0; o < CO_T; ++o) {
    const float b = (bias && o0 + o < cout) ? __ldg(bias + o0 + o) : 0.f;
#pragma unroll
    for (int k = 0; k < TD; ++k) acc[o][k] = b;
  }
  for (int i = 0; i < cin; ++i) {
    __syncthreads();
    load_halo<T>(tile, in + (static_cast<int64_t>(n) * cin + i) * vol, s, d0, h0, w0);
    for (int k = threadIdx.x; k < TAPS * CO_T; k += THREADS) {
      const int tap = k / CO_T, o = k - tap * CO_T;
      wsm[k] = o0 + o < cout ? w[(static_cast<int64_t>(o0 + o) * w_so + static_cast<int64_t>(i) * w_si) * TAPS + (FLIP ? TAPS - 1 - tap : tap)] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kh = 0; kh < K; ++kh)
#pragma unroll
      for (int kw = 0; kw < K; ++kw) {
        const float* col = tile + (ty + kh) * HW + tx + kw;
        float v[HD];
#pragma unroll
        for (int k = 0; k < HD; ++k) v[k] = col[k * HH * HW];
#pragma unroll
        for (int kd = 0; kd < K; ++kd) {
          const float* wt = wsm + ((kd * K + kh) * K + kw) * CO_T;
#pragma unroll
          for (int o4 = 0; o4 < CO_T; o4 += 4) {
            const float4 wk = *reinterpret_cast<const float4*>(wt + o4);
#pragma unroll
            for (int k = 0; k < TD; ++k) {
              acc[o4 + 0][k] = fmaf(v[k + kd], wk.x, acc[o4 + 0][k]);
              acc[o4 + 1][k] = fmaf(v[k + kd], wk.y, acc[o4 + 1][k]);
              acc[o4 + 2][k] = fmaf(v[k + kd], wk.z, acc[o4 + 2][k]);
              acc[o4 + 3][k] = fmaf(v[k + kd], wk.w, acc[o4 + 3][k]);
            }
          }
        }
      }
  }
  const int h = h0 + ty, wv = w0 + tx;
  if (h < s.H && wv < s.W) {
#pragma unroll
    for (int o = 0; o < CO_T; ++o)
      if (o0 + o < cout) {
        T* dst = out + (static_cast<int64_t>(n) * cout + o0 + o) * vol;
#pragma unroll
        for (int k = 0; k < TD; ++k)
          if (d0 + k < s.D) dst[(static_cast<int64_t>(d0 + k) * s.H + h) * s.W + wv] = Cvt<T>::from_f(acc[o][k]);
      }
  }
}

// part[((super * Cin + i) * co_groups + og) * WG_CO + o][28]; grid = (supers, Cin, co_groups), super = (n, d group of 4 tiles, h, w)
template <typename T>
__global__ void __launch_bounds__(THREADS) conv3_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, Dims s, int super_d,
                                                             float* __restrict__ part) {
  __shared__ float tile[HD * HH * HW];
  __shared__ float red[THREADS / 32][WG_CO * PST];
  int t = blockIdx.x;
  const int w0 = (t % s.tiles_w) * TW;
  t /= s.tiles_w;
  const int h0 = (t % s.tiles_h) * TH;
  t /= s.tiles_h;
  const int sd = t % super_d, n = t / super_d;
  const int i = blockIdx.y, o0 = blockIdx.z * WG_CO;
  const int64_t vol = static_cast<int64_t>(s.D) * s.H * s.W;
  const int tx = threadIdx.x % TW, ty = threadIdx.x / TW;
  const int h = h0 + ty, wv = w0 + tx;
  float acc[WG_CO][PST];
#pragma unroll
  for (int o = 0; o < WG_CO; ++o)
#pragma unroll
    for (int k = 0; k < PST; ++k) acc[o][k] = 0.f;
  for (int tt = 0; tt < WG_TILES; ++tt) {
    const int d0 = (sd * WG_TILES + tt) * TD;
    if (d0 >= s.D) break;                       // uniform over the CTA
    __syncthreads();
    load_halo<T>(tile, x + (static_cast<int64_t>(n) * s.Cin + i) * vol, s, d0, h0, w0);
    float g[WG_CO][TD];
#pragma unroll
    for (int o = 0; o < WG_CO; ++o)
#pragma unroll
      for (int k = 0; k < TD; ++k)
        g[o][k] = (o0 + o < s.Cout && h < s.H && wv < s.W && d0 + k < s.D)
                      ? Cvt<T>::to_f(dy[(static_cast<int64_t>(n) * s.Cout + o0 + o) * vol + (static_cast<int64_t>(d0 + k) * s.H + h) * s.W + wv])
                      : 0.f;
    __syncthreads();
#pragma unroll
    for (int o = 0; o < WG_CO; ++o)
#pragma unroll
      for (int k = 0; k < TD; ++k) acc[o][TAPS] += g[o][k];
#pragma unroll
    for (int kh = 0; kh < K; ++kh)
#pragma unroll
      for (int kw = 0; kw < K; ++kw) {
        const float* col = tile + (ty + kh) * HW + tx + kw;
        float v[HD];
#pragma unroll
        for (int k = 0; k < HD; ++k) v[k] = col[k * HH * HW];
#pragma unroll
        for (int kd = 0; kd < K; ++kd)
#pragma unroll
          for (int o = 0; o < WG_CO; ++o) {
            float a = acc[o][(kd * K + kh) * K + kw];
#pragma unroll
            for (int k = 0; k < TD; ++k) a = fmaf(g[o][k], v[k + kd], a);
            acc[o][(kd * K + kh) * K + kw] = a;
          }
      }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 0; o < WG_CO; ++o)
#pragma unroll
    for (int k = 0; k < PST; ++k) {
      float a = acc[o][k];
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) a += __shfl_xor_sync(0xffffffffu, a, sft);
      if (lane == 0) red[warp][o * PST + k] = a;
    }
  __syncthreads();
  if (threadIdx.x < WG_CO * PST) {
    float a = 0.f;
#pragma unroll
    for (int q = 0; q < THREADS / 32; ++q) a += red[q][threadIdx.x];
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * s.Cin + i) * gridDim.z + blockIdx.z;
    part[row * (WG_CO * PST) + threadIdx.x] = a;
  }
}

// dw[o][i][tap] = sum over supers of part[...]; db[o] from the rows of input channel 0.  grid = (Cout, Cin), block = 28 x 8
__global__ void __launch_bounds__(PST * 8) conv3_reduce_kernel(const float* __restrict__ part, int supers, int Cin, int Cout, int co_groups,
                                                              float* __restrict__ dw, float* __restrict__ db) {
  __shared__ double red[8][PST];
  const int o = blockIdx.x, i = blockIdx.y, k = threadIdx.x % PST, q = threadIdx.x / PST;
  const int og = o / WG_CO, ol = o % WG_CO;
  double v = 0.0;
  for (int sp = q; sp < supers; sp += 8)
    v += static_cast<double>(part[((static_cast<int64_t>(sp) * Cin + i) * co_groups + og) * (WG_CO * PST) + ol * PST + k]);
  red[q][k] = v;
  __syncthreads();
  if (q == 0) {
    double tsum = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r) tsum += red[r][k];
    if (k < TAPS) {
      if (dw) dw[(static_cast<int64_t>(o) * Cin + i) * TAPS + k] = static_cast<float>(tsum);
    } else if (db && i == 0) {
      db[o] = static_cast<float>(tsum);
    }
  }
}

int make_dims(int N, int Cin, int Cout, int D, int H, int W, Dims& s) {
  if (N <= 0 || Cin <= 0 || Cout <= 0 || D <= 0 || H <= 0 || W <= 0) return XHVED_ERR_BAD_ARG;
  if (Cin > 64 || Cout > 64) return XHVED_ERR_UNSUPPORTED_DIM;
  s.N = N, s.Cin = Cin, s.Cout = Cout, s.D = D, s.H = H, s.W = W;
  s.tiles_d = (D + TD - 1) / TD, s.tiles_h = (H + TH - 1) / TH, s.tiles_w = (W + TW - 1) / TW;
  return static_cast<int64_t>(N) * s.tiles_d * s.tiles_h * s.tiles_w > 0x7fffffffLL ? XHVED_ERR_BAD_SHAPE : 0;
}

template <typename T, bool FLIP>
int launch_conv(const T* in, const float* w, int w_so, int w_si, const float* bias, const Dims& s, int cin, int cout, T* out, cudaStream_t st) {
  const unsigned tiles = static_cast<unsigned>(s.N) * s.tiles_d * s.tiles_h * s.tiles_w;
  if (cout <= 4) conv3_kernel<T, 4, FLIP><<<dim3(tiles, 1), THREADS, 0, st>>>(in, w, w_so, w_si, bias, s, cin, cout, out);
  else conv3_kernel<T, 8, FLIP><<<dim3(tiles, (cout + 7) / 8), THREADS, 0, st>>>(in, w, w_so, w_si, bias, s, cin, cout, out);
  return (int)cudaGetLastError();
}

int supers_of(const Dims& s, int& super_d) {
  super_d = (s.tiles_d + WG_TILES - 1) / WG_TILES;
  return s.N * super_d * s.tiles_h * s.tiles_w;
}

template <typename T>
int bwd_t(const void* x, const float* w, const void* dy, const Dims& s, void* partials, void* dx, float* dw, float* db, cudaStream_t st) {
  if (dx) {
    const int rc = launch_conv<T, true>(static_cast<const T*>(dy), w, 1, s.Cin, nullptr, s, s.Cout, s.Cin, static_cast<T*>(dx), st);
    if (rc) return rc;
  }
  if (dw || db) {
    int super_d;
    const int supers = supers_of(s, super_d), co_groups = (s.Cout + WG_CO - 1) / WG_CO;
    float* part = static_cast<float*>(partials);
    conv3_wgrad_kernel<T><<<dim3(supers, s.Cin, co_groups), THREADS, 0, st>>>(static_cast<const T*>(x), static_cast<const T*>(dy), s, super_d, part);
    conv3_reduce_kernel<<<dim3(s.Cout, s.Cin), PST * 8, 0, st>>>(part, supers, s.Cin, s.Cout, co_groups, dw, db);
  }
  return (int)cudaGetLastError();
}

}  // namespace

extern "C" int64_t xhved_conv3_workspace(int N, int Cin, int Cout, int D, int H, int W) {
  Dims s;
  if (const int rc = make_dims(N, Cin, Cout, D, H, W, s)) return rc;
  int super_d;
  const int64_t supers = supers_of(s, super_d);
  return supers * Cin * ((Cout + WG_CO - 1) / WG_CO) * WG_CO * PST * static_cast<int64_t>(sizeof(float));
}

extern "C" int xhved_conv3_fwd(const void* x, const float* w, const float* bias, int N, int Cin, int Cout, int D, int H, int W, int dtype,
                               void* y, void* stream) {
  Dims s;
  if (const int rc = make_dims(N, Cin, Cout, D, H, W, s)) return rc;
  if (!x || !w || !y || dtype < 0 || dtype > 2) return XHVED_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case 0: return launch_conv<float, false>(static_cast<const float*>(x), w, Cin, 1, bias, s, Cin, Cout, static_cast<float*>(y), st);
    case 1: return launch_conv<__half, false>(static_cast<const __half*>(x), w, Cin, 1, bias, s, Cin, Cout, static_cast<__half*>(y), st);
    default:
      return launch_conv<__nv_bfloat16, false>(static_cast<const __nv_bfloat16*>(x), w, Cin, 1, bias, s, Cin, Cout,
                                               static_cast<__nv_bfloat16*>(y), st);
  }
}

extern "C" int xhved_conv3_bwd(const void* x, const float* w, const void* dy, int N, int Cin, int Cout, int D, int H, int W, int dtype,
                               void* partials, void* dx, float* dw, float* dbias, void* stream) {
  Dims s;
  if (const int rc = make_dims(N, Cin, Cout, D, H, W, s)) return rc;
  if (!dy || dtype < 0 || dtype > 2 || (dx && !w) || ((dw || dbias) && (!x || !partials))) return XHVED_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case 0: return bwd_t<float>(x, w, dy, s, partials, dx, dw, dbias, st);
    case 1: return bwd_t<__half>(x, w, dy, s, partials, dx, dw, dbias, st);
    default: return bwd_t<__nv_bfloat16>(x, w, dy, s, partials, dx, dw, dbias, st);
  }
}
