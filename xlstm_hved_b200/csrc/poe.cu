// S-MVAE product-of-experts fusion + reparameterised sampling + KL-to-prior, one pass -- sm_100a.
//
// Restates buildingblocks.py:846-886 (ProductOfExperts / ProductOfExperts2), RA_HVED.py:741-747
// (reparametrize) and loss.py:29-40 (KL_divergence against the prior) of the reference as a single
// HBM-bound elementwise kernel: the five (mu, logvar) slabs are read ONCE (128-bit loads) and every
// requested missing-modality subset (up to all 15) is produced from registers.  The KL reduction uses
// warp-shuffle + one atomic per block and subset.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "prof.cuh"
#include "xhved.h"

namespace xhved {

struct PoeSubsets {
  uint32_t mask[15];
  float kld_scale[15];
  int n;
  uint32_t used;  // union of masks: which modality slabs must be loaded
};

template <int V>
struct Vec;
template <>
struct Vec<4> {
  using T = float4;
};
template <>
struct Vec<1> {
  using T = float;
};

template <int V>
__device__ __forceinline__ void ld(const float* p, float* r) {
  if constexpr (V == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    r[0] = t.x, r[1] = t.y, r[2] = t.z, r[3] = t.w;
  } else {
    r[0] = __ldg(p);
  }
}
template <int V>
__device__ __forceinline__ void st(float* p, const float* r) {
  if constexpr (V == 4) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(r[0], r[1], r[2], r[3]));
  } else {
    *p = r[0];
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// SP: expert 0 is the constant standard-normal prior and is not read (XHVED_POE_STANDARD_PRIOR)
template <int V, int NS, bool SP>
__global__ void __launch_bounds__(256) poe_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv, int64_t n,
                                                       int64_t stride, PoeSubsets ss, const uint8_t* __restrict__ drop,
                                                       int64_t per_sample, float eps, float* __restrict__ out_mu,
                                                       float* __restrict__ out_lv, const float* __restrict__ noise,
                                                       float* __restrict__ out_z, float* __restrict__ kld_out) {
  float kld_acc[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) kld_acc[s] = 0.f;
  const int64_t nvec = n / V;
  for (int64_t iv = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; iv < nvec;
       iv += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t i = iv * V;
    float T[5][V], M[5][V], Lv[5][V];
    uint32_t live = (ss.used << 1) | (SP ? 0u : 1u);       // bit e = expert e must be read
    if (drop) {
      const uint8_t* d = drop + (i / per_sample) * 4;
      const uint32_t dropbits = (d[0] ? 1u : 0u) | (d[1] ? 2u : 0u) | (d[2] ? 4u : 0u) | (d[3] ? 8u : 0u);
      live &= ~(dropbits << 1);
    }
    // all loads first (predicated, no control flow) so that every thread keeps ~10 x 16 B in flight
#pragma unroll
    for (int e = 0; e < 5; ++e) {
#pragma unroll
      for (int j = 0; j < V; ++j) M[e][j] = 0.f, Lv[e][j] = 0.f;
      if ((live >> e) & 1u) {
        ld<V>(mu + e * stride + i, M[e]);
        ld<V>(lv + e * stride + i, Lv[e]);
      }
    }
#pragma unroll
    for (int e = 0; e < 5; ++e) {
      const bool on = (live >> e) & 1u;
#pragma unroll
      for (int j = 0; j < V; ++j) T[e][j] = on ? __fdividef(1.0f, __expf(Lv[e][j]) + eps) : 0.f;
    }
    if (SP) {       // mu_0 = 0, logvar_0 = 0 without reading them
#pragma unroll
      for (int j = 0; j < V; ++j) T[0][j] = __fdividef(1.0f, 1.0f + eps);
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (s < ss.n) {
        const uint32_t mask = ss.mask[s];
        float om[V], ol[V];
#pragma unroll
        for (int j = 0; j < V; ++j) {
          float st_ = T[0][j], sm = M[0][j] * T[0][j];
#pragma unroll
          for (int e = 1; e < 5; ++e)
            if ((mask >> (e - 1)) & 1u) st_ += T[e][j], sm += M[e][j] * T[e][j];
          om[j] = __fdividef(sm, st_);
          ol[j] = -__logf(st_);
        }
        st<V>(out_mu + s * n + i, om);
        st<V>(out_lv + s * n + i, ol);
        if (out_z) {
          float nz[V], z[V];
          ld<V>(noise + s * n + i, nz);
#pragma unroll
          for (int j = 0; j < V; ++j) z[j] = om[j] + nz[j] * __expf(0.5f * ol[j]);
          st<V>(out_z + s * n + i, z);
        }
        if (kld_out) {
#pragma unroll
          for (int j = 0; j < V; ++j) kld_acc[s] += -1.0f - ol[j] + (__expf(ol[j]) + om[j] * om[j]);   // / (1 + 1e-8) == 1 in fp32
        }
      }
    }
  }
  if (kld_out) {
    __shared__ float red[8][15];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (s < ss.n) {
        const float v = warp_sum(kld_acc[s]);
        if (lane == 0) red[warp][s] = v;
      }
    }
    __syncthreads();
    if (threadIdx.x < ss.n) {
      float v = 0.f;
      for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
      atomicAdd(kld_out + threadIdx.x, v);
    }
  }
}

template <int V, int NS, bool SP>
__global__ void __launch_bounds__(256, (NS <= 4 ? 2 : 1)) poe_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv, int64_t n,
                                                       int64_t stride, PoeSubsets ss, const uint8_t* __restrict__ drop,
                                                       int64_t per_sample, float eps, const float* __restrict__ g_mu,
                                                       const float* __restrict__ g_lv, const float* __restrict__ noise,
                                                       const float* __restrict__ g_z, int use_kld, float* __restrict__ d_mu,
                                                       float* __restrict__ d_lv) {
  const int64_t nvec = n / V;
  for (int64_t iv = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; iv < nvec;
       iv += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t i = iv * V;
    float T[5][V], M[5][V], EL[5][V], dM[5][V], dT[5][V];
    uint32_t dropbits = 0;
    if (drop) {
      const uint8_t* d = drop + (i / per_sample) * 4;
      dropbits = (d[0] ? 1u : 0u) | (d[1] ? 2u : 0u) | (d[2] ? 4u : 0u) | (d[3] ? 8u : 0u);
    }
    // all loads first, then the arithmetic
#pragma unroll
    for (int e = 0; e < 5; ++e) {
      if (e == 0 && SP) {
#pragma unroll
        for (int j = 0; j < V; ++j) M[0][j] = 0.f, EL[0][j] = 0.f;
      } else {
        ld<V>(mu + e * stride + i, M[e]);
        ld<V>(lv + e * stride + i, EL[e]);
      }
    }
#pragma unroll
    for (int e = 0; e < 5; ++e) {
      const bool dropped = (e > 0) && ((dropbits >> (e - 1)) & 1u);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        EL[e][j] = __expf(EL[e][j]);
        T[e][j] = dropped ? 0.f : __fdividef(1.0f, EL[e][j] + eps);
        if (dropped) M[e][j] = 0.f;
        dM[e][j] = 0.f, dT[e][j] = 0.f;
      }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (s < ss.n) {
        const uint32_t mask = (ss.mask[s] << 1) | 1u;  // bit e = expert e, prior always in
        float gm[V], gl[V];
#pragma unroll
        for (int j = 0; j < V; ++j) gm[j] = 0.f, gl[j] = 0.f;
        if (g_mu) {
          float t[V];
          ld<V>(g_mu + s * n + i, t);
#pragma unroll
          for (int j = 0; j < V; ++j) gm[j] += t[j];
        }
        if (g_lv) {
          float t[V];
          ld<V>(g_lv + s * n + i, t);
#pragma unroll
          for (int j = 0; j < V; ++j) gl[j] += t[j];
        }
        float gz[V], nz[V];
        if (g_z) {
          ld<V>(g_z + s * n + i, gz);
          ld<V>(noise + s * n + i, nz);
        }
#pragma unroll
        for (int j = 0; j < V; ++j) {
          float S = 0.f, sm = 0.f;
#pragma unroll
          for (int e = 0; e < 5; ++e)
            if ((mask >> e) & 1u) S += T[e][j], sm += M[e][j] * T[e][j];
          const float rS = __fdividef(1.0f, S);
          const float mh = sm * rS;
          const float lh = -__logf(S);
          float a = gm[j], b = gl[j];
          if (g_z) {
            a += gz[j];
            b += gz[j] * nz[j] * 0.5f * __expf(0.5f * lh);
          }
          if (use_kld) {
            const float ks = ss.kld_scale[s];
            a += ks * 2.0f * mh;                       // / (1 + 1e-8) == 1 in fp32
            b += ks * (-1.0f + __expf(lh));
          }
#pragma unroll
          for (int e = 0; e < 5; ++e)
            if ((mask >> e) & 1u) {
              dM[e][j] += a * T[e][j] * rS;
              dT[e][j] += (a * (M[e][j] - mh) - b) * rS;
            }
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 5; ++e) {
      if (e == 0 && SP) continue;
      float dl[V];
#pragma unroll
      for (int j = 0; j < V; ++j) dl[j] = -dT[e][j] * T[e][j] * T[e][j] * EL[e][j];
      st<V>(d_mu + e * stride + i, dM[e]);
      st<V>(d_lv + e * stride + i, dl);
    }
  }
}

template <int V>
__global__ void __launch_bounds__(256) reparam_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                                                           const float* __restrict__ noise, int64_t n, float* __restrict__ z) {
  const int64_t nvec = n / V;
  for (int64_t iv = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; iv < nvec;
       iv += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float a[V], b[V], c[V], o[V];
    ld<V>(mu + iv * V, a);
    ld<V>(lv + iv * V, b);
    ld<V>(noise + iv * V, c);
#pragma unroll
    for (int j = 0; j < V; ++j) o[j] = a[j] + c[j] * __expf(0.5f * b[j]);
    st<V>(z + iv * V, o);
  }
}
template <int V>
__global__ void __launch_bounds__(256) reparam_bwd_kernel(const float* __restrict__ lv, const float* __restrict__ noise,
                                                           const float* __restrict__ g, int64_t n, float* __restrict__ d_mu,
                                                           float* __restrict__ d_lv) {
  const int64_t nvec = n / V;
  for (int64_t iv = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; iv < nvec;
       iv += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float b[V], c[V], gg[V], o[V];
    ld<V>(lv + iv * V, b);
    ld<V>(noise + iv * V, c);
    ld<V>(g + iv * V, gg);
#pragma unroll
    for (int j = 0; j < V; ++j) o[j] = gg[j] * c[j] * 0.5f * __expf(0.5f * b[j]);
    st<V>(d_mu + iv * V, gg);
    st<V>(d_lv + iv * V, o);
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int grid_for(int64_t nvec) {
  // grid sized in multiples of the SM count (148 SMs); XHVED_POE_WAVES (blocks per SM, default 8) is a tuning knob
  static int per_sm = [] {
    const char* e = getenv("XHVED_POE_WAVES");
    const int v = e ? atoi(e) : 8;
    return v > 0 ? v : 8;
  }();
  const int64_t want = (nvec + 255) / 256;
  const int64_t cap = 148LL * per_sm;
  return static_cast<int>(want < 1 ? 1 : (want > cap ? cap : want));
}

static int fill_subsets(PoeSubsets& ss, const uint32_t* masks, int n_subsets, const float* kld_scale) {
  if (n_subsets < 1 || n_subsets > 15 || !masks) return XHVED_ERR_BAD_ARG;
  ss.n = n_subsets;
  ss.used = 0;
  for (int s = 0; s < 15; ++s) {
    ss.mask[s] = s < n_subsets ? (masks[s] & 15u) : 0u;
    ss.kld_scale[s] = (s < n_subsets && kld_scale) ? kld_scale[s] : 0.f;
    ss.used |= ss.mask[s];
  }
  return 0;
}

}  // namespace xhved

using namespace xhved;

extern "C" int xhved_version(void) { return 100; }

extern "C" int xhved_poe_fwd(const float* mu, const float* logvar, int64_t n, int64_t expert_stride, const uint32_t* subset_masks,
                             int n_subsets, const uint8_t* drop, int64_t per_sample, float eps, float* out_mu, float* out_logvar,
                             const float* noise, float* out_z, float* kld_out, int flags, void* stream) {
  if (n <= 0 || expert_stride < n || !mu || !logvar || !out_mu || !out_logvar) return XHVED_ERR_BAD_ARG;
  const int std_prior = (flags & XHVED_POE_STANDARD_PRIOR) ? 1 : 0;
  if ((out_z != nullptr) != (noise != nullptr)) return XHVED_ERR_BAD_ARG;
  if (drop && (per_sample <= 0 || n % per_sample)) return XHVED_ERR_BAD_SHAPE;
  PoeSubsets ss;
  if (int rc = fill_subsets(ss, subset_masks, n_subsets, nullptr)) return rc;
  cudaStream_t st_ = static_cast<cudaStream_t>(stream);
  const bool v4 = (n % 4 == 0) && (expert_stride % 4 == 0) && (!drop || per_sample % 4 == 0) && aligned16(mu) && aligned16(logvar) &&
                  aligned16(out_mu) && aligned16(out_logvar) && (!noise || (aligned16(noise) && aligned16(out_z)));
  ProfScope ps(K_POE_FWD, st_);
#define XHVED_POE_FWD(V, NS, GRID) \
  do {                                                                                                                             \
    if (std_prior)                                                                                                                 \
      poe_fwd_kernel<V, NS, true><<<GRID, 256, 0, st_>>>(mu, logvar, n, expert_stride, ss, drop, per_sample, eps, out_mu, out_logvar, \
                                                         noise, out_z, kld_out);                                                  \
    else                                                                                                                           \
      poe_fwd_kernel<V, NS, false><<<GRID, 256, 0, st_>>>(mu, logvar, n, expert_stride, ss, drop, per_sample, eps, out_mu, out_logvar, \
                                                          noise, out_z, kld_out);                                                 \
  } while (0)
  if (v4) {
    if (n_subsets == 1) XHVED_POE_FWD(4, 1, grid_for(n / 4));
    else if (n_subsets <= 4) XHVED_POE_FWD(4, 4, grid_for(n / 4));
    else XHVED_POE_FWD(4, 15, grid_for(n / 4));
  } else {
    if (n_subsets == 1) XHVED_POE_FWD(1, 1, grid_for(n));
    else XHVED_POE_FWD(1, 15, grid_for(n));
  }
#undef XHVED_POE_FWD
  return (int)cudaGetLastError();
}

extern "C" int xhved_poe_bwd(const float* mu, const float* logvar, int64_t n, int64_t expert_stride, const uint32_t* subset_masks,
                             int n_subsets, const uint8_t* drop, int64_t per_sample, float eps, const float* g_mu, const float* g_logvar,
                             const float* noise, const float* g_z, const float* kld_scale, float* d_mu, float* d_logvar, int flags,
                             void* stream) {
  if (n <= 0 || expert_stride < n || !mu || !logvar || !d_mu || !d_logvar) return XHVED_ERR_BAD_ARG;
  const int std_prior = (flags & XHVED_POE_STANDARD_PRIOR) ? 1 : 0;
  if (g_z && !noise) return XHVED_ERR_BAD_ARG;
  if (drop && (per_sample <= 0 || n % per_sample)) return XHVED_ERR_BAD_SHAPE;
  PoeSubsets ss;
  if (int rc = fill_subsets(ss, subset_masks, n_subsets, kld_scale)) return rc;
  cudaStream_t st_ = static_cast<cudaStream_t>(stream);
  const bool v4 = (n % 4 == 0) && (expert_stride % 4 == 0) && (!drop || per_sample % 4 == 0) && aligned16(mu) && aligned16(logvar) &&
                  aligned16(d_mu) && aligned16(d_logvar) && (!g_mu || aligned16(g_mu)) && (!g_logvar || aligned16(g_logvar)) &&
                  (!g_z || (aligned16(g_z) && aligned16(noise)));
  ProfScope ps(K_POE_BWD, st_);
#define XHVED_POE_BWD(V, NS, GRID)                                                                                                  \
  do {                                                                                                                              \
    if (std_prior)                                                                                                                  \
      poe_bwd_kernel<V, NS, true><<<GRID, 256, 0, st_>>>(mu, logvar, n, expert_stride, ss, drop, per_sample, eps, g_mu, g_logvar, noise, \
                                                         g_z, kld_scale != nullptr, d_mu, d_logvar);                                \
    else                                                                                                                            \
      poe_bwd_kernel<V, NS, false><<<GRID, 256, 0, st_>>>(mu, logvar, n, expert_stride, ss, drop, per_sample, eps, g_mu, g_logvar, noise, \
                                                          g_z, kld_scale != nullptr, d_mu, d_logvar);                               \
  } while (0)
  if (v4) {
    if (n_subsets == 1) XHVED_POE_BWD(4, 1, grid_for(n / 4));
    else if (n_subsets <= 4) XHVED_POE_BWD(4, 4, grid_for(n / 4));
    else XHVED_POE_BWD(4, 15, grid_for(n / 4));
  } else {
    if (n_subsets == 1) XHVED_POE_BWD(1, 1, grid_for(n));
    else XHVED_POE_BWD(1, 15, grid_for(n));
  }
#undef XHVED_POE_BWD
  return (int)cudaGetLastError();
}

extern "C" int xhved_reparam_fwd(const float* mu, const float* logvar, const float* noise, int64_t n, float* z, void* stream) {
  if (n <= 0 || !mu || !logvar || !noise || !z) return XHVED_ERR_BAD_ARG;
  cudaStream_t st_ = static_cast<cudaStream_t>(stream);
  ProfScope ps(K_REPARAM_FWD, st_);
  if (n % 4 == 0 && aligned16(mu) && aligned16(logvar) && aligned16(noise) && aligned16(z))
    reparam_fwd_kernel<4><<<grid_for(n / 4), 256, 0, st_>>>(mu, logvar, noise, n, z);
  else
    reparam_fwd_kernel<1><<<grid_for(n), 256, 0, st_>>>(mu, logvar, noise, n, z);
  return (int)cudaGetLastError();
}
extern "C" int xhved_reparam_bwd(const float* logvar, const float* noise, const float* g_z, int64_t n, float* d_mu, float* d_logvar,
                                 void* stream) {
  if (n <= 0 || !logvar || !noise || !g_z || !d_mu || !d_logvar) return XHVED_ERR_BAD_ARG;
  cudaStream_t st_ = static_cast<cudaStream_t>(stream);
  ProfScope ps(K_REPARAM_BWD, st_);
  if (n % 4 == 0 && aligned16(logvar) && aligned16(noise) && aligned16(g_z) && aligned16(d_mu) && aligned16(d_logvar))
    reparam_bwd_kernel<4><<<grid_for(n / 4), 256, 0, st_>>>(logvar, noise, g_z, n, d_mu, d_logvar);
  else
    reparam_bwd_kernel<1><<<grid_for(n), 256, 0, st_>>>(logvar, noise, g_z, n, d_mu, d_logvar);
  return (int)cudaGetLastError();
}
