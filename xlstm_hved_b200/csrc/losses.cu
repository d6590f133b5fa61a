// Per-channel Dice coefficient of the training loss (loss.py:257-285, used through DiceLoss, loss.py:188-209) -- the tail of the
// training step SURVEY.md 8f lists next to the KL term: one pass over probabilities and targets instead of the reference's
// permute / contiguous / three element-wise products / three reductions.
//
//   dice_c = 2 I_c / max(D_c, eps),   I_c = sum p t,   D_c = sum p^2 + sum t^2      (sums over batch and voxels of channel c)
//   d dice_c / d p = 2 t / D_c - 4 I_c p / D_c^2          (0 where D_c <= eps: the clamp is flat there)
//
// HBM-bound: 8 B in per voxel forward (p, t), 8 B in + 4 B out backward.
#include <cuda_runtime.h>
#include <stdint.h>

#include "xhved.h"

namespace {

// sums[3 c + {0,1,2}] += {sum p t, sum p^2, sum t^2} over one contiguous slab of channel c; grid = (blocks per slab, N * C)
__global__ void __launch_bounds__(256) dice_sums_kernel(const float* __restrict__ p, const float* __restrict__ t, int C, int64_t spatial,
                                                        float* __restrict__ sums) {
  const int64_t slab = blockIdx.y;
  const int c = static_cast<int>(slab % C);
  const float* pp = p + slab * spatial;
  const float* tt = t + slab * spatial;
  float a = 0.f, b = 0.f, d = 0.f;
  const bool vec = (spatial % 4 == 0) && ((reinterpret_cast<uintptr_t>(pp) | reinterpret_cast<uintptr_t>(tt)) % 16 == 0);
  if (vec) {
    const int64_t n4 = spatial / 4;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(pp) + i), y = __ldg(reinterpret_cast<const float4*>(tt) + i);
      a += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
      b += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
      d += y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w;
    }
  } else {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < spatial; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
      const float x = __ldg(pp + i), y = __ldg(tt + i);
      a += x * y, b += x * x, d += y * y;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    d += __shfl_xor_sync(0xffffffffu, d, o);
  }
  __shared__ float red[3][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[0][warp] = a, red[1][warp] = b, red[2][warp] = d;
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
    atomicAdd(sums + 3 * c + threadIdx.x, s);
  }
}

// dp = g_c * (2 t / D_c - 4 I_c p / D_c^2), g_c = upstream gradient of dice_c
__global__ void __launch_bounds__(256) dice_bwd_kernel(const float* __restrict__ p, const float* __restrict__ t, const float* __restrict__ sums,
                                                       const float* __restrict__ g_dice, int C, int64_t spatial, float eps,
                                                       float* __restrict__ dp) {
  const int64_t slab = blockIdx.y;
  const int c = static_cast<int>(slab % C);
  const float I = sums[3 * c], D = sums[3 * c + 1] + sums[3 * c + 2];
  const float g = g_dice[c];
  const float ka = D > eps ? g * 2.f / D : 0.f, kb = D > eps ? g * 4.f * I / (D * D) : 0.f;
  const float* pp = p + slab * spatial;
  const float* tt = t + slab * spatial;
  float* dd = dp + slab * spatial;
  const bool vec = (spatial % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(pp) | reinterpret_cast<uintptr_t>(tt) | reinterpret_cast<uintptr_t>(dd)) % 16 == 0);
  if (vec) {
    const int64_t n4 = spatial / 4;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(pp) + i), y = __ldg(reinterpret_cast<const float4*>(tt) + i);
      reinterpret_cast<float4*>(dd)[i] = make_float4(ka * y.x - kb * x.x, ka * y.y - kb * x.y, ka * y.z - kb * x.z, ka * y.w - kb * x.w);
    }
  } else {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < spatial; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
      dd[i] = ka * __ldg(tt + i) - kb * __ldg(pp + i);
  }
}

int blocks_per_slab(int64_t spatial, int64_t slabs) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = (static_cast<int64_t>(sms) * 16 + slabs - 1) / slabs;       // ~16 blocks per SM over the whole launch
  const int64_t cap = (spatial / 4 + 255) / 256;
  const int64_t b = want < 1 ? 1 : (want > cap ? (cap < 1 ? 1 : cap) : want);
  return static_cast<int>(b);
}

}  // namespace

extern "C" int xhved_dice_sums(const float* p, const float* t, int N, int C, int64_t spatial, float* sums, void* stream) {
  if (!p || !t || !sums || N <= 0 || C <= 0 || spatial <= 0 || static_cast<int64_t>(N) * C > 65535) return XHVED_ERR_BAD_ARG;
  const dim3 grid(blocks_per_slab(spatial, static_cast<int64_t>(N) * C), N * C);
  dice_sums_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p, t, C, spatial, sums);
  return (int)cudaGetLastError();
}

extern "C" int xhved_dice_bwd(const float* p, const float* t, const float* sums, const float* g_dice, int N, int C, int64_t spatial,
                              float eps, float* dp, void* stream) {
  if (!p || !t || !sums || !g_dice || !dp || N <= 0 || C <= 0 || spatial <= 0 || static_cast<int64_t>(N) * C > 65535) return XHVED_ERR_BAD_ARG;
  const dim3 grid(blocks_per_slab(spatial, static_cast<int64_t>(N) * C), N * C);
  dice_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p, t, sums, g_dice, C, spatial, eps, dp);
  return (int)cudaGetLastError();
}
