// K9: 1x1x1 convolution (channel mixing) of the conv path, Cin / Cout <= 32 -- the squeeze / excite / fuse layers of DuSEAttention
// (modules/DuSFE.py:100-104), the 1x1x1 layers of the VU blocks and final_conv (RA_HVED.py:567-605, 483).  Under the reference's
// fp16 autocast (train.py:207) cuDNN runs the BACKWARD of the (1, 4, 128^3) 4 -> 4 layers on a 32x32 wmma GEMM tile at 11.5 ms per
// call: 263 ms of a 431 ms training step (gpurun_out/r02z_train_profile.txt); in fp32 0.22 ms per call.
//
// HBM-bound: (Cin + Cout) * sizeof(T) bytes per voxel forward.  x, y: (N, C, vol) contiguous, T = fp32 / fp16 / bf16, weights,
// bias and accumulation in fp32.
//   forward / dgrad  one thread per 4 consecutive voxels, Cout accumulator vectors in registers, the weights (transposed for the
//                    input gradient) in shared memory;
//   wgrad            dW[co][ci] = sum_v dy[co][v] x[ci][v], db[co] = sum_v dy[co][v]: for Cin, Cout <= 8 every thread keeps the
//                    8 x 8 (+ 8) partial sums in registers over a grid-stride loop, one block reduction per CTA, per-CTA partials
//                    and a reduction kernel; wider layers (16 / 32 channels, at 32^3 and below) take one CTA per (co, ci) pair.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "xhved.h"

namespace {

constexpr int THREADS = 256;
constexpr int VEC = 4;
constexpr int MAXC = 32;

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

template <typename T> struct alignas(sizeof(T) * VEC) Pack { T v[VEC]; };

// VEC voxels starting at `at` of a plane of `vol` voxels; ALIGNED: vol % VEC == 0 and 16 / 8-byte aligned bases
template <typename T, bool ALIGNED>
__device__ __forceinline__ void load4(const T* __restrict__ p, int64_t at, int64_t vol, float (&v)[VEC]) {
  if constexpr (ALIGNED) {
    const Pack<T> raw = *reinterpret_cast<const Pack<T>*>(p + at);
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] = Cvt<T>::to_f(raw.v[i]);
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] = at + i < vol ? Cvt<T>::to_f(p[at + i]) : 0.f;
  }
}

template <typename T, bool ALIGNED>
__device__ __forceinline__ void store4(T* __restrict__ p, int64_t at, int64_t vol, const float (&v)[VEC]) {
  if constexpr (ALIGNED) {
    Pack<T> raw;
#pragma unroll
    for (int i = 0; i < VEC; ++i) raw.v[i] = Cvt<T>::from_f(v[i]);
    *reinterpret_cast<Pack<T>*>(p + at) = raw;
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i)
      if (at + i < vol) p[at + i] = Cvt<T>::from_f(v[i]);
  }
}

// out[n][o][v] = bias[o] + sum_i w[o * w_so + i * w_si] in[n][i][v];  CO_T = Cout rounded up to the register tile
template <typename T, int CO_T, bool ALIGNED>
__global__ void __launch_bounds__(THREADS) pw_kernel(const T* __restrict__ in, const float* __restrict__ w, int w_so, int w_si,
                                                    const float* __restrict__ bias, int N, int Cin, int Cout, int64_t vol,
                                                    int64_t vecs_per_plane, T* __restrict__ out) {
  __shared__ __align__(16) float wsm[MAXC * CO_T];        // [i][o]
  for (int k = threadIdx.x; k < Cin * CO_T; k += THREADS) {
    const int i = k / CO_T, o = k - i * CO_T;
    wsm[k] = o < Cout ? w[o * w_so + i * w_si] : 0.f;
  }
  __syncthreads();
  const int64_t t = static_cast<int64_t>(blockIdx.x) * THREADS + threadIdx.x;
  if (t >= vecs_per_plane * N) return;
  const int64_t n = t / vecs_per_plane, at = (t - n * vecs_per_plane) * VEC;
  float acc[CO_T][VEC];
#pragma unroll
  for (int o = 0; o < CO_T; ++o) {
    const float b = (bias && o < Cout) ? __ldg(bias + o) : 0.f;
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[o][k] = b;
  }
  const T* src = in + n * Cin * vol;
  for (int i = 0; i < Cin; ++i) {
    float xv[VEC];
    load4<T, ALIGNED>(src + i * vol, at, vol, xv);
#pragma unroll
    for (int o = 0; o < CO_T; ++o) {
      const float wk = wsm[i * CO_T + o];
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[o][k] = fmaf(wk, xv[k], acc[o][k]);
    }
  }
  T* dst = out + n * Cout * vol;
#pragma unroll
  for (int o = 0; o < CO_T; ++o)
    if (o < Cout) store4<T, ALIGNED>(dst + o * vol, at, vol, acc[o]);
}

// ---- wgrad, Cin, Cout <= 8: part[cta][72] = {dW[o][i] (o * 8 + i), db[o] (64 + o)}
constexpr int SMALL = 8, PSTRIDE = SMALL * SMALL + SMALL;

template <typename T, bool ALIGNED>
__global__ void __launch_bounds__(THREADS) pw_wgrad_small_kernel(const T* __restrict__ x, const T* __restrict__ dy, int N, int Cin, int Cout,
                                                                int64_t vol, int64_t vecs_per_plane, float* __restrict__ part) {
  __shared__ float red[THREADS / 32][PSTRIDE];
  float acc[SMALL][SMALL], accb[SMALL];
#pragma unroll
  for (int o = 0; o < SMALL; ++o) {
    accb[o] = 0.f;
#pragma unroll
    for (int i = 0; i < SMALL; ++i) acc[o][i] = 0.f;
  }
  const int64_t total = vecs_per_plane * N;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * THREADS + threadIdx.x; t < total; t += static_cast<int64_t>(gridDim.x) * THREADS) {
    const int64_t n = t / vecs_per_plane, at = (t - n * vecs_per_plane) * VEC;
    float xv[SMALL][VEC], gv[SMALL][VEC];
#pragma unroll
    for (int i = 0; i < SMALL; ++i) {
      if (i < Cin) load4<T, ALIGNED>(x + (n * Cin + i) * vol, at, vol, xv[i]);
      else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) xv[i][k] = 0.f;
      }
    }
#pragma unroll
    for (int o = 0; o < SMALL; ++o) {
      if (o < Cout) load4<T, ALIGNED>(dy + (n * Cout + o) * vol, at, vol, gv[o]);
      else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) gv[o][k] = 0.f;
      }
    }
#pragma unroll
    for (int o = 0; o < SMALL; ++o) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) accb[o] += gv[o][k];
#pragma unroll
      for (int i = 0; i < SMALL; ++i)
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[o][i] = fmaf(gv[o][k], xv[i][k], acc[o][i]);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 0; o < SMALL; ++o) {
#pragma unroll
    for (int i = 0; i < SMALL; ++i) {
      float v = acc[o][i];
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
      if (lane == 0) red[warp][o * SMALL + i] = v;
    }
    float v = accb[o];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if (lane == 0) red[warp][SMALL * SMALL + o] = v;
  }
  __syncthreads();
  if (threadIdx.x < PSTRIDE) {
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < THREADS / 32; ++q) v += red[q][threadIdx.x];
    part[static_cast<int64_t>(blockIdx.x) * PSTRIDE + threadIdx.x] = v;
  }
}

__global__ void __launch_bounds__(PSTRIDE) pw_wgrad_small_reduce_kernel(const float* __restrict__ part, int ctas, int Cin, int Cout,
                                                                       float* __restrict__ dw, float* __restrict__ db) {
  double v = 0.0;
  for (int r = 0; r < ctas; ++r) v += static_cast<double>(part[static_cast<int64_t>(r) * PSTRIDE + threadIdx.x]);
  if (threadIdx.x < SMALL * SMALL) {
    const int o = threadIdx.x / SMALL, i = threadIdx.x % SMALL;
    if (dw && o < Cout && i < Cin) dw[o * Cin + i] = static_cast<float>(v);
  } else if (db && threadIdx.x - SMALL * SMALL < Cout) {
    db[threadIdx.x - SMALL * SMALL] = static_cast<float>(v);
  }
}

// ---- wgrad, wider layers: CTA b < Cout * Cin -> dW[o][i]; CTA b >= Cout * Cin -> db[b - Cout * Cin]
template <typename T>
__global__ void __launch_bounds__(THREADS) pw_wgrad_wide_kernel(const T* __restrict__ x, const T* __restrict__ dy, int N, int Cin, int Cout,
                                                               int64_t vol, float* __restrict__ dw, float* __restrict__ db) {
  __shared__ double red[THREADS / 32];
  const int pairs = Cout * Cin;
  const bool is_bias = static_cast<int>(blockIdx.x) >= pairs;
  const int o = is_bias ? blockIdx.x - pairs : blockIdx.x / Cin, i = is_bias ? 0 : blockIdx.x % Cin;
  if (is_bias ? db == nullptr : dw == nullptr) return;
  double s = 0.0;
  for (int n = 0; n < N; ++n) {
    const T* g = dy + (static_cast<int64_t>(n) * Cout + o) * vol;
    const T* xs = x + (static_cast<int64_t>(n) * Cin + i) * vol;
    float a = 0.f;
    for (int64_t v = threadIdx.x; v < vol; v += THREADS) a = is_bias ? a + Cvt<T>::to_f(g[v]) : fmaf(Cvt<T>::to_f(g[v]), Cvt<T>::to_f(xs[v]), a);
    s += static_cast<double>(a);
  }
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < THREADS / 32; ++q) t += red[q];
    if (is_bias) db[o] = static_cast<float>(t);
    else dw[o * Cin + i] = static_cast<float>(t);
  }
}

template <typename T>
bool is_aligned(int64_t vol, const void* a, const void* b) {
  return vol % VEC == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) % (sizeof(T) * VEC) == 0);
}

template <typename T, bool ALIGNED>
int launch_pw(const T* in, const float* w, int w_so, int w_si, const float* bias, int N, int Cin, int Cout, int64_t vol, T* out,
              cudaStream_t st) {
  const int64_t vpp = (vol + VEC - 1) / VEC, threads = vpp * N, ctas = (threads + THREADS - 1) / THREADS;
  if (ctas > 0x7fffffffLL) return XHVED_ERR_BAD_SHAPE;
  const unsigned grid = static_cast<unsigned>(ctas);
  if (Cout <= 1) pw_kernel<T, 1, ALIGNED><<<grid, THREADS, 0, st>>>(in, w, w_so, w_si, bias, N, Cin, Cout, vol, vpp, out);
  else if (Cout <= 4) pw_kernel<T, 4, ALIGNED><<<grid, THREADS, 0, st>>>(in, w, w_so, w_si, bias, N, Cin, Cout, vol, vpp, out);
  else if (Cout <= 8) pw_kernel<T, 8, ALIGNED><<<grid, THREADS, 0, st>>>(in, w, w_so, w_si, bias, N, Cin, Cout, vol, vpp, out);
  else if (Cout <= 16) pw_kernel<T, 16, ALIGNED><<<grid, THREADS, 0, st>>>(in, w, w_so, w_si, bias, N, Cin, Cout, vol, vpp, out);
  else pw_kernel<T, 32, ALIGNED><<<grid, THREADS, 0, st>>>(in, w, w_so, w_si, bias, N, Cin, Cout, vol, vpp, out);
  return (int)cudaGetLastError();
}

template <typename T>
int fwd_t(const void* x, const float* w, const float* bias, int N, int Cin, int Cout, int64_t vol, void* y, cudaStream_t st) {
  const T* xi = static_cast<const T*>(x);
  T* yo = static_cast<T*>(y);
  return is_aligned<T>(vol, x, y) ? launch_pw<T, true>(xi, w, Cin, 1, bias, N, Cin, Cout, vol, yo, st)
                                  : launch_pw<T, false>(xi, w, Cin, 1, bias, N, Cin, Cout, vol, yo, st);
}

int small_ctas(int64_t vecs) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t need = (vecs + THREADS - 1) / THREADS, cap = static_cast<int64_t>(sms) * 4;
  return static_cast<int>(need < cap ? (need < 1 ? 1 : need) : cap);
}

template <typename T>
int bwd_t(const void* x, const float* w, const void* dy, int N, int Cin, int Cout, int64_t vol, void* partials, void* dx, float* dw,
          float* db, cudaStream_t st) {
  const T* xi = static_cast<const T*>(x);
  const T* g = static_cast<const T*>(dy);
  if (dx) {
    // dx[n][i][v] = sum_o w[o][i] dy[n][o][v]: the forward kernel with the roles of the channels swapped
    T* d = static_cast<T*>(dx);
    const int rc = is_aligned<T>(vol, dy, dx) ? launch_pw<T, true>(g, w, 1, Cin, nullptr, N, Cout, Cin, vol, d, st)
                                              : launch_pw<T, false>(g, w, 1, Cin, nullptr, N, Cout, Cin, vol, d, st);
    if (rc) return rc;
  }
  if (dw || db) {
    if (Cin <= SMALL && Cout <= SMALL) {
      if (!partials) return XHVED_ERR_BAD_ARG;
      const int64_t vpp = (vol + VEC - 1) / VEC;
      const int ctas = small_ctas(vpp * N);
      float* part = static_cast<float*>(partials);
      if (is_aligned<T>(vol, x, dy)) pw_wgrad_small_kernel<T, true><<<ctas, THREADS, 0, st>>>(xi, g, N, Cin, Cout, vol, vpp, part);
      else pw_wgrad_small_kernel<T, false><<<ctas, THREADS, 0, st>>>(xi, g, N, Cin, Cout, vol, vpp, part);
      pw_wgrad_small_reduce_kernel<<<1, PSTRIDE, 0, st>>>(part, ctas, Cin, Cout, dw, db);
    } else {
      pw_wgrad_wide_kernel<T><<<Cout * Cin + Cout, THREADS, 0, st>>>(xi, g, N, Cin, Cout, vol, dw, db);
    }
  }
  return (int)cudaGetLastError();
}

int check_args(int N, int Cin, int Cout, int64_t vol, int dtype) {
  if (N <= 0 || Cin <= 0 || Cout <= 0 || vol <= 0 || dtype < 0 || dtype > 2) return XHVED_ERR_BAD_ARG;
  if (Cin > MAXC || Cout > MAXC) return XHVED_ERR_UNSUPPORTED_DIM;
  return 0;
}

}  // namespace

extern "C" int64_t xhved_pwconv_workspace(int N, int Cin, int Cout, int64_t vol) {
  if (N <= 0 || Cin <= 0 || Cout <= 0 || vol <= 0) return XHVED_ERR_BAD_ARG;
  return static_cast<int64_t>(148 * 8) * PSTRIDE * static_cast<int64_t>(sizeof(float));      // >= small_ctas(...) partial rows
}

extern "C" int xhved_pwconv_fwd(const void* x, const float* w, const float* bias, int N, int Cin, int Cout, int64_t vol, int dtype, void* y,
                                void* stream) {
  if (const int rc = check_args(N, Cin, Cout, vol, dtype)) return rc;
  if (!x || !w || !y) return XHVED_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case 0: return fwd_t<float>(x, w, bias, N, Cin, Cout, vol, y, st);
    case 1: return fwd_t<__half>(x, w, bias, N, Cin, Cout, vol, y, st);
    default: return fwd_t<__nv_bfloat16>(x, w, bias, N, Cin, Cout, vol, y, st);
  }
}

extern "C" int xhved_pwconv_bwd(const void* x, const float* w, const void* dy, int N, int Cin, int Cout, int64_t vol, int dtype,
                                void* partials, void* dx, float* dw, float* dbias, void* stream) {
  if (const int rc = check_args(N, Cin, Cout, vol, dtype)) return rc;
  if (!dy || (dx && !w) || ((dw || dbias) && !x)) return XHVED_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case 0: return bwd_t<float>(x, w, dy, N, Cin, Cout, vol, partials, dx, dw, dbias, st);
    case 1: return bwd_t<__half>(x, w, dy, N, Cin, Cout, vol, partials, dx, dw, dbias, st);
    default: return bwd_t<__nv_bfloat16>(x, w, dy, N, Cin, Cout, vol, partials, dx, dw, dbias, st);
  }
}
