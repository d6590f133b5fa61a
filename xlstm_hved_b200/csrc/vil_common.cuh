// Helpers shared by the ViL pre/post kernels (K2/K3): token geometry, parameter staging.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "umma.cuh"
#include "prof.cuh"
#include "xhved.h"

namespace xhved {

constexpr int kTok = 128;  // tokens per CTA == cell chunk length
__host__ __device__ constexpr uint32_t next_pow2_tmem(int n) { return n <= 32 ? 32u : n <= 64 ? 64u : n <= 128 ? 128u : n <= 256 ? 256u : 512u; }

struct VilGeom {
  int B, S, nc, Sp, NH, DH, DHP, reverse;
  int64_t xsb, xsn, xsc, ysb, ysn, ysc;
  int grep;          // gradient replicas (>= 1)
  int64_t gstride;   // elements between replicas
};

__host__ inline int vil_validate(const xhved_vil_shape* sh, VilGeom* g) {
  if (!sh || sh->B <= 0 || sh->S <= 0) return XHVED_ERR_BAD_SHAPE;
  const int C = sh->C, E = 2 * C;
  if (C != 16 && C != 32 && C != 64) return XHVED_ERR_UNSUPPORTED_DIM;
  if (sh->QB != 4 || sh->NH != 4) return XHVED_ERR_UNSUPPORTED_DIM;  // vision_lstm.py:357,402-405
  g->B = sh->B, g->S = sh->S, g->nc = (sh->S + kTok - 1) / kTok, g->Sp = g->nc * kTok;
  g->NH = sh->NH, g->DH = E / sh->NH, g->DHP = g->DH <= 16 ? 16 : g->DH, g->reverse = sh->reverse;
  g->xsb = sh->x_stride_b, g->xsn = sh->x_stride_n, g->xsc = sh->x_stride_c;
  g->ysb = sh->y_stride_b, g->ysn = sh->y_stride_n, g->ysc = sh->y_stride_c;
  g->grep = sh->grad_replicas > 1 ? sh->grad_replicas : 1;
  g->gstride = sh->grad_replica_stride;
  return 0;
}

// the gradient block of the replica this CTA accumulates into (see xhved_vil_shape::grad_replicas)
__device__ __forceinline__ xhved_vil_grads replica_of(xhved_vil_grads gr, const VilGeom& g) {
  if (g.grep > 1) {
    const int64_t off = static_cast<int64_t>(blockIdx.x % g.grep) * g.gstride;
    gr.norm_weight += off, gr.proj_up_weight += off, gr.conv_weight += off, gr.conv_bias += off, gr.q_weight += off;
    gr.k_weight += off, gr.v_weight += off, gr.igate_weight += off, gr.igate_bias += off, gr.fgate_weight += off;
    gr.fgate_bias += off, gr.outnorm_weight += off, gr.learnable_skip += off, gr.proj_down_weight += off;
  }
  return gr;
}

// sigmoid through the SFU (ex2 + rcp, ~2 ulp): an IEEE division costs ~15 issue slots and these kernels are issue-bound.
// The exponent is clamped so that the denominator stays inside __fdividef's exact range (< 2^126).
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(fminf(-x, 80.f))); }
__device__ __forceinline__ float silu(float x) { return x * fast_sigmoid(x); }
// d/dx silu(x) = s + x s (1 - s), s = sigmoid(x)
__device__ __forceinline__ float dsilu(float x) {
  const float s = fast_sigmoid(x);
  return s * (1.f + x * (1.f - s));
}
// silu and its derivative from one sigmoid
__device__ __forceinline__ void silu_both(float x, float& y, float& dy) {
  const float s = fast_sigmoid(x);
  y = x * s;
  dy = s * (1.f + x * (1.f - s));
}

// SM count of the current device (persistent kernels launch one or two CTAs per SM)
inline int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}
inline int persistent_grid(int ntiles, int ctas_per_sm) {
  const int cap = sm_count() * ctas_per_sm;
  return ntiles < cap ? ntiles : cap;
}

__device__ __forceinline__ void stage(float* dst, const float* __restrict__ src, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

}  // namespace xhved

namespace xhved {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// warp-reduce a per-token value and add it into a shared accumulator slot (one smem atomic per warp)
__device__ __forceinline__ void warp_acc(float* slot, float v) {
  v = warp_sum_f(v);
  if ((threadIdx.x & 31) == 0) atomicAdd(slot, v);
}

// Reduce 32 per-lane values across the warp at once: after the call lane i holds sum over lanes of v[i] (in v[0]).
// 31 shuffles instead of 32 x 5 for 32 separate butterfly reductions.
__device__ __forceinline__ float warp_transpose_sum32(float* v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int s = 16, n = 32; s >= 1; s >>= 1, n >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = up ? v[i] : v[i + n / 2];
      const float recv = __shfl_xor_sync(0xffffffffu, send, s);
      v[i] = (up ? v[i + n / 2] : v[i]) + recv;
    }
  }
  return v[0];
}
// N (power of two <= 32) per-lane values -> lane i holds the warp sum of v[i & (N-1)]
template <int N>
__device__ __forceinline__ float warp_transpose_sum(float* v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int s = N / 2, n = N; s >= 1; s >>= 1, n >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = up ? v[i] : v[i + n / 2];
      const float recv = __shfl_xor_sync(0xffffffffu, send, s);
      v[i] = (up ? v[i + n / 2] : v[i]) + recv;
    }
  }
  float r = v[0];
#pragma unroll
  for (int s = N; s < 32; s <<= 1) r += __shfl_xor_sync(0xffffffffu, r, s);
  return r;
}
// N per-token values -> N shared accumulators (slot i <- warp sum of v[i]); one conflict-free smem atomic per lane
template <int N>
__device__ __forceinline__ void warp_acc_vec(float* slots, float* v) {
  const float r = warp_transpose_sum<N>(v);
  if ((threadIdx.x & 31) < N) atomicAdd(slots + (threadIdx.x & 31), r);
}

}  // namespace xhved
