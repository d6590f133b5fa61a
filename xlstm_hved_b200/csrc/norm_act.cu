// K6: InstanceNorm3d / BatchNorm3d fused with the LeakyReLU that follows it -- the first row of SURVEY.md 8f ("conv path:
// fuse InstanceNorm + LeakyReLU"), the callers on either side of the ViL / S-MVAE path.
//
// Reference call sites: SingleConv order 'ilc' = InstanceNorm3d -> LeakyReLU(0.01) -> Conv3d (buildingblocks.py:400-462, used by
// every encoder / decoder block of XLSTM_HVED), BasicConv (buildingblocks.py:11-31), BatchNorm3d of DuSFE / AttenModule2
// (modules/DuSFE.py:17-36, 108-110, 187).  With one volume per step and 4-32 channels (train.py:50, f_maps = 4) PyTorch's
// batch_norm kernels launch ONE thread block per (sample, channel) plane: 3.1 ms for a (1, 4, 128^3) tensor of 33 MB, 64 of the
// 90 ms of a whole inference forward on a B200 (gpurun_out/r02v_model_kernels.txt).  Here a plane is cut into chunks of 16 KB, one
// CTA per chunk:
//
//   pass 1 (norm_stats):  chunk -> (mean_i, M2_i) around its own mean, both from registers (one read of x);
//   pass 2 (norm_apply):  every CTA merges the <= few hundred partials of its statistics group in fp64 (Chan's formula; they are
//                         L2-resident), then y = lrelu((x - mean) * rstd * gamma + beta) on its chunk (x again: an L2 hit for
//                         tensors below the 126 MB L2);
//   planes that fit one chunk take both passes in ONE launch (norm_small).
// Backward (the activation's mask is recomputed from x, nothing but mean / rstd is saved):
//   g = dy * (pre > 0 ? 1 : slope),  s1 = sum g,  s2 = sum g * xhat,  dx = gamma * rstd * (g - s1 / M - xhat * s2 / M),
//   dgamma_c += s2, dbeta_c += s1 -- the same two passes.
//
// HBM-bound: algorithmic bytes per element = 2 * sizeof(T) forward (x in, y out), 3 * sizeof(T) backward (x, dy in, dx out).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "prof.cuh"
#include "xhved.h"

namespace {

constexpr int THREADS = 256;
constexpr int ITEMS = 4;          // 16-byte vectors per thread and chunk

template <typename T> struct Io;
template <> struct Io<float> {
  static constexpr int VEC = 4;
  static __device__ __forceinline__ float to_f(float v) { return v; }
  static __device__ __forceinline__ float from_f(float v) { return v; }
};
template <> struct Io<__half> {
  static constexpr int VEC = 8;
  static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
};
template <> struct Io<__nv_bfloat16> {
  static constexpr int VEC = 8;
  static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};

template <typename T> constexpr int chunk_elems() { return THREADS * ITEMS * Io<T>::VEC; }

// VEC consecutive elements of a plane starting at element `at`; elements at or beyond `M` read as 0.
// ALIGNED: M % VEC == 0 and the plane base is 16-byte aligned, so a vector is either wholly inside or wholly outside.
template <typename T, bool ALIGNED>
__device__ __forceinline__ void load_vec(const T* __restrict__ plane, int64_t at, int64_t M, float (&v)[Io<T>::VEC]) {
  constexpr int VEC = Io<T>::VEC;
  if constexpr (ALIGNED) {
    if (at < M) {
      const uint4 raw = *reinterpret_cast<const uint4*>(plane + at);
      const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
      for (int i = 0; i < VEC; ++i) v[i] = Io<T>::to_f(e[i]);
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) v[i] = 0.f;
    }
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] = at + i < M ? Io<T>::to_f(plane[at + i]) : 0.f;
  }
}

template <typename T, bool ALIGNED>
__device__ __forceinline__ void store_vec(T* __restrict__ plane, int64_t at, int64_t M, const float (&v)[Io<T>::VEC]) {
  constexpr int VEC = Io<T>::VEC;
  if constexpr (ALIGNED) {
    if (at < M) {
      uint4 raw;
      T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
      for (int i = 0; i < VEC; ++i) e[i] = Io<T>::from_f(v[i]);
      *reinterpret_cast<uint4*>(plane + at) = raw;
    }
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i)
      if (at + i < M) plane[at + i] = Io<T>::from_f(v[i]);
  }
}

// Sum of a pair over the CTA; every thread receives the result.  `red` holds 2 * (THREADS / 32) values.
template <typename F>
__device__ __forceinline__ void block_sum2(F& a, F& b, F* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();                         // `red` may still be read from a previous call
  if (lane == 0) red[2 * warp] = a, red[2 * warp + 1] = b;
  __syncthreads();
  a = 0, b = 0;
#pragma unroll
  for (int w = 0; w < THREADS / 32; ++w) a += red[2 * w], b += red[2 * w + 1];
}

struct Geom {
  int64_t M;          // elements per (sample, channel) plane
  int nchunks;        // chunks per plane
  int C, N;
  int mode;           // 0 instance, 1 batch (statistics over N and the plane), 2 frozen (mean / rstd given per channel)
  float eps, slope;
};

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

// (mean_i, M2_i) of one chunk around its own mean.  The sum is taken of (x - pivot), pivot = the chunk's first element: a constant
// plane then yields mean == x exactly and the output is exactly beta (PyTorch's Welford kernel is exact there on power-of-two
// planes only; on the model's real tensors, where |mean| / std reaches 1e3 - 1e5, it is 8e-4 off the fp64 value and this
// kernel 2e-5: tools/diag_conv_norm.py) -- the encoders of a MISSING modality (input zeroed, evaluation.py:306-307) see constant planes, where one ulp of error in the mean is
// multiplied by rstd = 1 / sqrt(eps) = 316 and renormalised to unit variance by the next layer.
template <typename T, bool ALIGNED>
__device__ __forceinline__ void chunk_stats(const float (&v)[ITEMS][Io<T>::VEC], float pivot, int64_t first, int64_t M, float& mean,
                                            float& m2, float* red) {
  constexpr int VEC = Io<T>::VEC;
  float s = 0.f, unused = 0.f;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int64_t at = first + (static_cast<int64_t>(j) * THREADS + threadIdx.x) * VEC;
#pragma unroll
    for (int i = 0; i < VEC; ++i) s += (at + i < M) ? v[j][i] - pivot : 0.f;
  }
  block_sum2(s, unused, red);
  const int64_t left = M - first;
  const float cnt = static_cast<float>(left < chunk_elems<T>() ? left : chunk_elems<T>());
  mean = pivot + s / cnt;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int64_t at = first + (static_cast<int64_t>(j) * THREADS + threadIdx.x) * VEC;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float d = v[j][i] - mean;
      q += (at + i < M) ? d * d : 0.f;
    }
  }
  unused = 0.f;
  block_sum2(q, unused, red);
  m2 = q;
}

template <typename T, bool ALIGNED>
__device__ __forceinline__ void load_chunk(const T* __restrict__ plane, int64_t first, int64_t M, float (&v)[ITEMS][Io<T>::VEC]) {
#pragma unroll
  for (int j = 0; j < ITEMS; ++j)
    load_vec<T, ALIGNED>(plane, first + (static_cast<int64_t>(j) * THREADS + threadIdx.x) * Io<T>::VEC, M, v[j]);
}

// ---- pass 1 forward: partials[plane * nchunks + chunk] = (mean_i, M2_i)
template <typename T, bool ALIGNED>
__global__ void __launch_bounds__(THREADS) norm_stats_kernel(const T* __restrict__ x, Geom g, float2* __restrict__ partials) {
  __shared__ float red[2 * THREADS / 32];
  const int64_t plane = blockIdx.x / g.nchunks;
  const int chunk = static_cast<int>(blockIdx.x % g.nchunks);
  const int64_t first = static_cast<int64_t>(chunk) * chunk_elems<T>();
  float v[ITEMS][Io<T>::VEC];
  load_chunk<T, ALIGNED>(x + plane * g.M, first, g.M, v);
  float mean, m2;
  chunk_stats<T, ALIGNED>(v, Io<T>::to_f(x[plane * g.M + first]), first, g.M, mean, m2, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = make_float2(mean, m2);
}

// Merge the partials of the statistics group of (n, c): instance = the plane's own chunks, batch = the chunks of channel c over all
// samples.  Every thread returns (mean, rstd).
template <typename T>
__device__ __forceinline__ void merge_group(const float2* __restrict__ partials, const Geom& g, int n, int c, double* red, float& mean,
                                            float& rstd) {
  const int n_lo = g.mode == 0 ? n : 0, n_hi = g.mode == 0 ? n + 1 : g.N;
  const int per_plane = g.nchunks;
  const int total = (n_hi - n_lo) * per_plane;
  const int64_t last_cnt = g.M - static_cast<int64_t>(per_plane - 1) * chunk_elems<T>();
  double sw = 0.0, cnt_all = static_cast<double>(g.M) * (n_hi - n_lo), zero = 0.0;
  // around the first partial's mean, like the chunk sums: equal chunk means merge to exactly that value
  const double pivot = static_cast<double>(partials[(static_cast<int64_t>(n_lo) * g.C + c) * per_plane].x);
  for (int i = threadIdx.x; i < total; i += THREADS) {
    const int nn = n_lo + i / per_plane, ch = i % per_plane;
    const float2 p = partials[(static_cast<int64_t>(nn) * g.C + c) * per_plane + ch];
    sw += (static_cast<double>(p.x) - pivot) * static_cast<double>(ch == per_plane - 1 ? last_cnt : chunk_elems<T>());
  }
  block_sum2(sw, zero, red);
  const double mu = pivot + sw / cnt_all;
  double m2 = 0.0;
  zero = 0.0;
  for (int i = threadIdx.x; i < total; i += THREADS) {
    const int nn = n_lo + i / per_plane, ch = i % per_plane;
    const float2 p = partials[(static_cast<int64_t>(nn) * g.C + c) * per_plane + ch];
    const double d = static_cast<double>(p.x) - mu;
    m2 += static_cast<double>(p.y) + d * d * static_cast<double>(ch == per_plane - 1 ? last_cnt : chunk_elems<T>());
  }
  block_sum2(m2, zero, red);
  mean = static_cast<float>(mu);
  rstd = static_cast<float>(1.0 / sqrt(m2 / cnt_all + static_cast<double>(g.eps)));
}

// ---- pass 2 forward
template <typename T, bool ALIGNED>
__global__ void __launch_bounds__(THREADS) norm_apply_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, Geom g,
                                                             const float2* __restrict__ partials, float* __restrict__ mean_io,
                                                             float* __restrict__ rstd_io, T* __restrict__ y) {
  __shared__ double red[2 * THREADS / 32];
  const int64_t plane = blockIdx.x / g.nchunks;
  const int chunk = static_cast<int>(blockIdx.x % g.nchunks);
  const int n = static_cast<int>(plane / g.C), c = static_cast<int>(plane % g.C);
  const int64_t first = static_cast<int64_t>(chunk) * chunk_elems<T>();
  float v[ITEMS][Io<T>::VEC];
  load_chunk<T, ALIGNED>(x + plane * g.M, first, g.M, v);      // issued before the merge: the loads overlap it
  float mean, rstd;
  if (g.mode == 2) {
    mean = mean_io[c], rstd = rstd_io[c];
  } else {
    merge_group<T>(partials, g, n, c, red, mean, rstd);
    const int64_t slot = g.mode == 0 ? plane : c;
    if (threadIdx.x == 0 && chunk == 0 && (g.mode == 0 || n == 0)) mean_io[slot] = mean, rstd_io[slot] = rstd;
  }
  const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
  const float a = rstd * ga;               // (x - mean) first: x * a + (beta - mean * a) cancels badly where |mean| >> std
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
#pragma unroll
    for (int i = 0; i < Io<T>::VEC; ++i) v[j][i] = lrelu(fmaf(v[j][i] - mean, a, be), g.slope);
    store_vec<T, ALIGNED>(y + plane * g.M, first + (static_cast<int64_t>(j) * THREADS + threadIdx.x) * Io<T>::VEC, g.M, v[j]);
  }
}

// ---- both forward passes for planes of at most one chunk (instance mode): one CTA per plane
template <typename T, bool ALIGNED>
__global__ void __launch_bounds__(THREADS) norm_small_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, Geom g, float* __restrict__ mean_out,
                                                             float* __restrict__ rstd_out, T* __restrict__ y) {
  __shared__ float red[2 * THREADS / 32];
  const int64_t plane = blockIdx.x;
  const int c = static_cast<int>(plane % g.C);
  float v[ITEMS][Io<T>::VEC];
  load_chunk<T, ALIGNED>(x + plane * g.M, 0, g.M, v);
  float mean, m2;
  chunk_stats<T, ALIGNED>(v, Io<T>::to_f(x[plane * g.M]), 0, g.M, mean, m2, red);
  const float rstd = rsqrtf(m2 / static_cast<float>(g.M) + g.eps);
  if (threadIdx.x == 0) mean_out[plane] = mean, rstd_out[plane] = rstd;
  const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
  const float a = rstd * ga;               // (x - mean) first: x * a + (beta - mean * a) cancels badly where |mean| >> std
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
#pragma unroll
    for (int i = 0; i < Io<T>::VEC; ++i) v[j][i] = lrelu(fmaf(v[j][i] - mean, a, be), g.slope);
    store_vec<T, ALIGNED>(y + plane * g.M, (static_cast<int64_t>(j) * THREADS + threadIdx.x) * Io<T>::VEC, g.M, v[j]);
  }
}

// g = dy * act'(pre), xhat: shared by the two backward passes
template <typename T, bool ALIGNED>
__device__ __forceinline__ void load_grad_chunk(const T* __restrict__ xp, const T* __restrict__ dyp, int64_t first, int64_t M, float mean,
                                                float rstd, float ga, float be, float slope, float (&xh)[ITEMS][Io<T>::VEC],
                                                float (&gg)[ITEMS][Io<T>::VEC]) {
  load_chunk<T, ALIGNED>(xp, first, M, xh);
  load_chunk<T, ALIGNED>(dyp, first, M, gg);           // out-of-range elements: dy = 0, so g = 0 there
#pragma unroll
  for (int j = 0; j < ITEMS; ++j)
#pragma unroll
    for (int i = 0; i < Io<T>::VEC; ++i) {
      const float h = (xh[j][i] - mean) * rstd;
      const float pre = fmaf(h, ga, be);
      xh[j][i] = h;
      gg[j][i] = pre > 0.f ? gg[j][i] : gg[j][i] * slope;
    }
}

__device__ __forceinline__ void group_stats(const Geom& g, int64_t plane, int c, const float* mean, const float* rstd, float& m, float& r) {
  const int64_t slot = g.mode == 0 ? plane : c;
  m = mean[slot], r = rstd[slot];
}

// ---- pass 1 backward: partials[plane * nchunks + chunk] = (sum g, sum g * xhat)
template <typename T, bool ALIGNED>
__global__ void __launch_bounds__(THREADS) norm_bwd_stats_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                 const float* __restrict__ mean, const float* __restrict__ rstd, Geom g,
                                                                 float2* __restrict__ partials) {
  __shared__ float red[2 * THREADS / 32];
  const int64_t plane = blockIdx.x / g.nchunks;
  const int chunk = static_cast<int>(blockIdx.x % g.nchunks);
  const int c = static_cast<int>(plane % g.C);
  const int64_t first = static_cast<int64_t>(chunk) * chunk_elems<T>();
  float m, r;
  group_stats(g, plane, c, mean, rstd, m, r);
  float xh[ITEMS][Io<T>::VEC], gg[ITEMS][Io<T>::VEC];
  load_grad_chunk<T, ALIGNED>(x + plane * g.M, dy + plane * g.M, first, g.M, m, r, gamma ? gamma[c] : 1.f, beta ? beta[c] : 0.f, g.slope,
                              xh, gg);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j)
#pragma unroll
    for (int i = 0; i < Io<T>::VEC; ++i) s1 += gg[j][i], s2 = fmaf(gg[j][i], xh[j][i], s2);
  block_sum2(s1, s2, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = make_float2(s1, s2);
}

// ---- pass 2 backward
template <typename T, bool ALIGNED>
__global__ void __launch_bounds__(THREADS) norm_bwd_apply_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                 const float* __restrict__ mean, const float* __restrict__ rstd, Geom g,
                                                                 const float2* __restrict__ partials, T* __restrict__ dx,
                                                                 float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ double red[2 * THREADS / 32];
  const int64_t plane = blockIdx.x / g.nchunks;
  const int chunk = static_cast<int>(blockIdx.x % g.nchunks);
  const int n = static_cast<int>(plane / g.C), c = static_cast<int>(plane % g.C);
  const int64_t first = static_cast<int64_t>(chunk) * chunk_elems<T>();
  float m, r;
  group_stats(g, plane, c, mean, rstd, m, r);
  const float ga = gamma ? gamma[c] : 1.f;
  float xh[ITEMS][Io<T>::VEC], gg[ITEMS][Io<T>::VEC];
  load_grad_chunk<T, ALIGNED>(x + plane * g.M, dy + plane * g.M, first, g.M, m, r, ga, beta ? beta[c] : 0.f, g.slope, xh, gg);
  // sums of the statistics group (frozen statistics: the group for dgamma / dbeta is the channel, and dx takes no correction)
  const int n_lo = g.mode == 0 ? n : 0, n_hi = g.mode == 0 ? n + 1 : g.N;
  const int total = (n_hi - n_lo) * g.nchunks;
  double s1 = 0.0, s2 = 0.0;
  for (int i = threadIdx.x; i < total; i += THREADS) {
    const float2 p = partials[(static_cast<int64_t>(n_lo + i / g.nchunks) * g.C + c) * g.nchunks + i % g.nchunks];
    s1 += static_cast<double>(p.x), s2 += static_cast<double>(p.y);
  }
  block_sum2(s1, s2, red);
  if (threadIdx.x == 0 && chunk == 0 && (g.mode == 0 || n == 0)) {
    if (dgamma) atomicAdd(dgamma + c, static_cast<float>(s2));
    if (dbeta) atomicAdd(dbeta + c, static_cast<float>(s1));
  }
  const double cnt = static_cast<double>(g.M) * (n_hi - n_lo);
  const float k1 = g.mode == 2 ? 0.f : static_cast<float>(s1 / cnt), k2 = g.mode == 2 ? 0.f : static_cast<float>(s2 / cnt);
  const float a = ga * r;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
#pragma unroll
    for (int i = 0; i < Io<T>::VEC; ++i) gg[j][i] = a * (gg[j][i] - k1 - xh[j][i] * k2);
    store_vec<T, ALIGNED>(dx + plane * g.M, first + (static_cast<int64_t>(j) * THREADS + threadIdx.x) * Io<T>::VEC, g.M, gg[j]);
  }
}

template <typename T>
bool aligned(int64_t M, const void* a, const void* b, const void* c) {
  return M % Io<T>::VEC == 0 &&
         ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) % 16 == 0);
}

int check_shape(const xhved_norm_shape* sh) {
  if (!sh || sh->N <= 0 || sh->C <= 0 || sh->spatial <= 0 || sh->mode < 0 || sh->mode > 2 || sh->dtype < 0 || sh->dtype > 2)
    return XHVED_ERR_BAD_ARG;
  return 0;
}

template <typename T>
Geom geom(const xhved_norm_shape* sh) {
  Geom g;
  g.M = sh->spatial, g.C = sh->C, g.N = sh->N, g.mode = sh->mode, g.eps = sh->eps, g.slope = sh->slope;
  g.nchunks = static_cast<int>((sh->spatial + chunk_elems<T>() - 1) / chunk_elems<T>());
  return g;
}

template <typename T>
int fwd_t(const void* x_, const float* gamma, const float* beta, const xhved_norm_shape* sh, float* mean, float* rstd, void* partials,
          void* y_, cudaStream_t st) {
  const T* x = static_cast<const T*>(x_);
  T* y = static_cast<T*>(y_);
  const Geom g = geom<T>(sh);
  const int64_t planes = static_cast<int64_t>(g.N) * g.C, ctas = planes * g.nchunks;
  if (ctas > 0x7fffffffLL) return XHVED_ERR_BAD_SHAPE;
  const bool al = aligned<T>(g.M, x, y, nullptr);
  const unsigned grid = static_cast<unsigned>(ctas);
  if (g.mode == 0 && g.nchunks == 1) {
    if (al) norm_small_kernel<T, true><<<grid, THREADS, 0, st>>>(x, gamma, beta, g, mean, rstd, y);
    else norm_small_kernel<T, false><<<grid, THREADS, 0, st>>>(x, gamma, beta, g, mean, rstd, y);
    return (int)cudaGetLastError();
  }
  float2* part = static_cast<float2*>(partials);
  if (g.mode != 2) {
    if (!part) return XHVED_ERR_BAD_ARG;
    if (al) norm_stats_kernel<T, true><<<grid, THREADS, 0, st>>>(x, g, part);
    else norm_stats_kernel<T, false><<<grid, THREADS, 0, st>>>(x, g, part);
  }
  if (al) norm_apply_kernel<T, true><<<grid, THREADS, 0, st>>>(x, gamma, beta, g, part, mean, rstd, y);
  else norm_apply_kernel<T, false><<<grid, THREADS, 0, st>>>(x, gamma, beta, g, part, mean, rstd, y);
  return (int)cudaGetLastError();
}

template <typename T>
int bwd_t(const void* x_, const void* dy_, const float* gamma, const float* beta, const float* mean, const float* rstd,
          const xhved_norm_shape* sh, void* partials, void* dx_, float* dgamma, float* dbeta, cudaStream_t st) {
  const T* x = static_cast<const T*>(x_);
  const T* dy = static_cast<const T*>(dy_);
  T* dx = static_cast<T*>(dx_);
  const Geom g = geom<T>(sh);
  const int64_t planes = static_cast<int64_t>(g.N) * g.C, ctas = planes * g.nchunks;
  if (ctas > 0x7fffffffLL) return XHVED_ERR_BAD_SHAPE;
  const bool al = aligned<T>(g.M, x, dy, dx);
  const unsigned grid = static_cast<unsigned>(ctas);
  float2* part = static_cast<float2*>(partials);
  if (al) {
    norm_bwd_stats_kernel<T, true><<<grid, THREADS, 0, st>>>(x, dy, gamma, beta, mean, rstd, g, part);
    norm_bwd_apply_kernel<T, true><<<grid, THREADS, 0, st>>>(x, dy, gamma, beta, mean, rstd, g, part, dx, dgamma, dbeta);
  } else {
    norm_bwd_stats_kernel<T, false><<<grid, THREADS, 0, st>>>(x, dy, gamma, beta, mean, rstd, g, part);
    norm_bwd_apply_kernel<T, false><<<grid, THREADS, 0, st>>>(x, dy, gamma, beta, mean, rstd, g, part, dx, dgamma, dbeta);
  }
  return (int)cudaGetLastError();
}

}  // namespace

extern "C" int64_t xhved_norm_act_workspace(int N, int C, int64_t spatial, int dtype) {
  if (N <= 0 || C <= 0 || spatial <= 0 || dtype < 0 || dtype > 2) return XHVED_ERR_BAD_ARG;
  const int64_t chunk = dtype == 0 ? chunk_elems<float>() : chunk_elems<__half>();
  return static_cast<int64_t>(N) * C * ((spatial + chunk - 1) / chunk) * static_cast<int64_t>(sizeof(float2));
}

extern "C" int xhved_norm_act_fwd(const void* x, const float* gamma, const float* beta, const xhved_norm_shape* shape, float* mean,
                                  float* rstd, void* partials, void* y, void* stream) {
  if (const int rc = check_shape(shape)) return rc;
  if (!x || !y || !mean || !rstd) return XHVED_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  xhved::ProfScope ps(xhved::K_NORM_FWD, st);
  switch (shape->dtype) {
    case 0: return fwd_t<float>(x, gamma, beta, shape, mean, rstd, partials, y, st);
    case 1: return fwd_t<__half>(x, gamma, beta, shape, mean, rstd, partials, y, st);
    default: return fwd_t<__nv_bfloat16>(x, gamma, beta, shape, mean, rstd, partials, y, st);
  }
}

extern "C" int xhved_norm_act_bwd(const void* x, const void* dy, const float* gamma, const float* beta, const float* mean,
                                  const float* rstd, const xhved_norm_shape* shape, void* partials, void* dx, float* dgamma, float* dbeta,
                                  void* stream) {
  if (const int rc = check_shape(shape)) return rc;
  if (!x || !dy || !dx || !mean || !rstd || !partials) return XHVED_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  xhved::ProfScope ps(xhved::K_NORM_BWD, st);
  switch (shape->dtype) {
    case 0: return bwd_t<float>(x, dy, gamma, beta, mean, rstd, shape, partials, dx, dgamma, dbeta, st);
    case 1: return bwd_t<__half>(x, dy, gamma, beta, mean, rstd, shape, partials, dx, dgamma, dbeta, st);
    default: return bwd_t<__nv_bfloat16>(x, dy, gamma, beta, mean, rstd, shape, partials, dx, dgamma, dbeta, st);
  }
}
