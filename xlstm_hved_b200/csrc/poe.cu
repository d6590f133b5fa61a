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
  // fused clip (RA_HVED.py:580, 749-753): the logvar slabs of the four modality experts are clamped to [clip_lo, clip_hi]
  // as they are loaded; the backward zeroes their gradient outside the interval (what torch.clamp's autograd does)
  int clip;
  float clip_lo, clip_hi;
};
constexpr float kKlEps = 1e-8f;   // KL_divergence's own eps (loss.py:29), independent of the fusion's

// One latent level of a launch.  A launch covers up to four levels (the four latent resolutions of a volume): the small
// levels are latency-bound on their own, so they ride along with the big one.  blk_end[l] = first block of level l+1.
struct PoeLevelF {
  const float* mu;
  const float* lv;
  int64_t n, stride;
  const uint8_t* drop;
  int64_t per_sample;
  float* out_mu;
  float* out_lv;
  const float* noise;
  float* out_z;
  float* kld_out;
};
struct PoeFwdArgs {
  PoeLevelF lev[4];
  int blk_end[4];
  int nlev;
};
struct PoeLevelB {
  const float* mu;
  const float* lv;
  int64_t n, stride;
  const uint8_t* drop;
  int64_t per_sample;
  const float* g_mu;
  const float* g_lv;
  const float* noise;
  const float* g_z;
  float* d_mu;
  float* d_lv;
  int use_kld;
  float kld_scale[15];
};
struct PoeBwdArgs {
  PoeLevelB lev[4];
  int blk_end[4];
  int nlev;
};
// level of this block (selects instead of dynamic indexing into the parameter space)
template <typename Args, typename Level>
__device__ __forceinline__ void select_level(const Args& a, Level& L, int& blk0, int& nblk) {
  L = a.lev[0];
  blk0 = 0;
  int end = a.blk_end[0];
#pragma unroll
  for (int l = 1; l < 4; ++l) {
    if (l < a.nlev && static_cast<int>(blockIdx.x) >= a.blk_end[l - 1]) {
      L = a.lev[l];
      blk0 = a.blk_end[l - 1];
      end = a.blk_end[l];
    }
  }
  nblk = end - blk0;
}

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
// STREAM: evict-first store.  Measured: it helps the backward (whose outputs compete with five input slabs for L2) by ~5 %
// and costs the forward up to 35 % when 45 output streams are written at once, so only the backward uses it.
template <int V, bool STREAM = false>
__device__ __forceinline__ void st(float* p, const float* r) {
  if constexpr (V == 4) {
    if constexpr (STREAM) __stcs(reinterpret_cast<float4*>(p), make_float4(r[0], r[1], r[2], r[3]));
    else *reinterpret_cast<float4*>(p) = make_float4(r[0], r[1], r[2], r[3]);
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
template <int V, int NS, bool SP, bool CLIP>
__global__ void __launch_bounds__(256) poe_fwd_kernel(const PoeFwdArgs args, const PoeSubsets ss, const float eps) {
  PoeLevelF L;
  int blk0, nblk;
  select_level(args, L, blk0, nblk);
  const float* __restrict__ mu = L.mu;
  const float* __restrict__ lv = L.lv;
  const int64_t n = L.n, stride = L.stride, per_sample = L.per_sample;
  const uint8_t* __restrict__ drop = L.drop;
  float* __restrict__ out_mu = L.out_mu;
  float* __restrict__ out_lv = L.out_lv;
  const float* __restrict__ noise = L.noise;
  float* __restrict__ out_z = L.out_z;
  float* __restrict__ kld_out = L.kld_out;
  float kld_acc[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) kld_acc[s] = 0.f;
  const int64_t nvec = n / V;
  for (int64_t iv = (blockIdx.x - blk0) * static_cast<int64_t>(blockDim.x) + threadIdx.x; iv < nvec;
       iv += static_cast<int64_t>(nblk) * blockDim.x) {
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
    if (CLIP) {
#pragma unroll
      for (int e = 1; e < 5; ++e)
#pragma unroll
        for (int j = 0; j < V; ++j) Lv[e][j] = fminf(fmaxf(Lv[e][j], ss.clip_lo), ss.clip_hi);
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
        // var^ = 1 / sum T: one reciprocal serves mu^ = (sum mu T) var^, exp(logvar^) = var^ (KL) and, through its square
        // root, the standard deviation of the sample
        float om[V], ol[V], var[V], sd[V];
#pragma unroll
        for (int j = 0; j < V; ++j) {
          float st_ = T[0][j], sm = M[0][j] * T[0][j];
#pragma unroll
          for (int e = 1; e < 5; ++e)
            if ((mask >> (e - 1)) & 1u) st_ += T[e][j], sm += M[e][j] * T[e][j];
          var[j] = __fdividef(1.0f, st_);
          sd[j] = rsqrtf(st_);
          om[j] = sm * var[j];
          ol[j] = -__logf(st_);
        }
        st<V>(out_mu + s * n + i, om);
        st<V>(out_lv + s * n + i, ol);
        if (out_z) {
          float nz[V], z[V];
          ld<V>(noise + s * n + i, nz);
#pragma unroll
          for (int j = 0; j < V; ++j) z[j] = om[j] + nz[j] * sd[j];
          st<V>(out_z + s * n + i, z);
        }
        if (kld_out) {
          // KL_divergence(sub_mu, sub_logvar, mu_prior, logvar_prior) (loss.py:29-40, 113): the prior is slab 0
#pragma unroll
          for (int j = 0; j < V; ++j) {
            if (SP) {
              kld_acc[s] += -1.0f - ol[j] + (var[j] + om[j] * om[j]);   // / (1 + 1e-8) == 1 in fp32
            } else {
              const float dm = om[j] - M[0][j];
              kld_acc[s] += -1.0f + Lv[0][j] - ol[j] + (var[j] + dm * dm) * __frcp_rn(__expf(Lv[0][j]) + kKlEps);
            }
          }
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

template <int V, int NS, bool SP, bool CLIP>
__global__ void __launch_bounds__(256, 2) poe_bwd_kernel(const PoeBwdArgs args, const PoeSubsets ss, const float eps) {
  PoeLevelB L;
  int blk0, nblk;
  select_level(args, L, blk0, nblk);
  const float* __restrict__ mu = L.mu;
  const float* __restrict__ lv = L.lv;
  const int64_t n = L.n, stride = L.stride, per_sample = L.per_sample;
  const uint8_t* __restrict__ drop = L.drop;
  const float* __restrict__ g_mu = L.g_mu;
  const float* __restrict__ g_lv = L.g_lv;
  const float* __restrict__ noise = L.noise;
  const float* __restrict__ g_z = L.g_z;
  float* __restrict__ d_mu = L.d_mu;
  float* __restrict__ d_lv = L.d_lv;
  const int use_kld = L.use_kld;
  const int64_t nvec = n / V;
  for (int64_t iv = (blockIdx.x - blk0) * static_cast<int64_t>(blockDim.x) + threadIdx.x; iv < nvec;
       iv += static_cast<int64_t>(nblk) * blockDim.x) {
    const int64_t i = iv * V;
    float T[5][V], M[5][V], EL[5][V], dM[5][V], dT[5][V];
    float dL0[V];                 // direct KL gradient w.r.t. the prior logvar (general prior only)
    uint32_t pass = 0xFFFFFFFFu;  // bit e*4+j: gradient of logvar[e][j] passes the fused clip
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
    if (CLIP) {
#pragma unroll
      for (int e = 1; e < 5; ++e)
#pragma unroll
        for (int j = 0; j < V; ++j) {
          if (EL[e][j] < ss.clip_lo || EL[e][j] > ss.clip_hi) pass &= ~(1u << (e * 4 + j));
          EL[e][j] = fminf(fmaxf(EL[e][j], ss.clip_lo), ss.clip_hi);
        }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) dL0[j] = 0.f;
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
          const float rS = __fdividef(1.0f, S);      // = exp(logvar^)
          const float mh = sm * rS;
          float a = gm[j], b = gl[j];
          if (g_z) {
            a += gz[j];
            b += gz[j] * nz[j] * 0.5f * rsqrtf(S);   // d z / d logvar^ = 0.5 eps exp(0.5 logvar^)
          }
          if (use_kld) {
            const float ks = L.kld_scale[s];
            if (SP) {
              a += ks * 2.0f * mh;                       // / (1 + 1e-8) == 1 in fp32
              b += ks * (-1.0f + rS);
            } else {
              // d/d(mu^, logvar^) of -1 + lv0 - lv^ + (exp(lv^) + (mu^ - mu0)^2) / (exp(lv0) + 1e-8), and the direct
              // terms w.r.t. the prior slab itself (on top of its gradient as expert 0 of the fusion)
              const float rv2 = __frcp_rn(EL[0][j] + kKlEps);   // exactly 1 for the standard prior
              const float dm = mh - M[0][j];
              a += ks * 2.0f * dm * rv2;
              b += ks * (-1.0f + rS * rv2);
              dM[0][j] -= ks * 2.0f * dm * rv2;
              dL0[j] += ks * (1.0f - (rS + dm * dm) * EL[0][j] * rv2 * rv2);
            }
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
      for (int j = 0; j < V; ++j) {
        dl[j] = -dT[e][j] * T[e][j] * T[e][j] * EL[e][j];
        if (e == 0) dl[j] += dL0[j];
        else if (CLIP && !((pass >> (e * 4 + j)) & 1u)) dl[j] = 0.f;
      }
      st<V, true>(d_mu + e * stride + i, dM[e]);
      st<V, true>(d_lv + e * stride + i, dl);
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
  // grid sized in multiples of the SM count (148 SMs); XHVED_POE_WAVES (blocks per SM, default 16) is a tuning knob:
  // measured on the bench step, 4 / 8 / 16 / 32 / 64 blocks per SM give 0.294 / 0.261 / 0.240 / 0.240 / 0.244 ms of PoE time
  static int per_sm = [] {
    const char* e = getenv("XHVED_POE_WAVES");
    const int v = e ? atoi(e) : 16;
    return v > 0 ? v : 16;
  }();
  const int64_t want = (nvec + 255) / 256;
  const int64_t cap = 148LL * per_sm;
  return static_cast<int>(want < 1 ? 1 : (want > cap ? cap : want));
}

static int fill_subsets(PoeSubsets& ss, const uint32_t* masks, int n_subsets, const xhved_poe_opts* o) {
  if (n_subsets < 1 || n_subsets > 15 || !masks) return XHVED_ERR_BAD_ARG;
  ss.clip = (o && (o->flags & XHVED_POE_CLIP)) ? 1 : 0;
  ss.clip_lo = o ? o->clip_lo : 0.f;
  ss.clip_hi = o ? o->clip_hi : 0.f;
  if (ss.clip && !(ss.clip_lo <= ss.clip_hi)) return XHVED_ERR_BAD_ARG;
  ss.n = n_subsets;
  ss.used = 0;
  for (int s = 0; s < 15; ++s) {
    ss.mask[s] = s < n_subsets ? (masks[s] & 15u) : 0u;
    ss.kld_scale[s] = 0.f;
    ss.used |= ss.mask[s];
  }
  return 0;
}

// blocks per level: the launch has 148 x XHVED_POE_WAVES blocks at most, shared out in proportion to the level sizes
template <typename Args>
static int assign_blocks(Args& a, int vec) {
  int64_t want[4], total = 0;
  for (int l = 0; l < a.nlev; ++l) {
    want[l] = (a.lev[l].n / vec + 255) / 256;
    if (want[l] < 1) want[l] = 1;
    total += want[l];
  }
  const int64_t cap = grid_for(INT64_MAX / 2);
  int end = 0;
  for (int l = 0; l < a.nlev; ++l) {
    int64_t blocks = total <= cap ? want[l] : (want[l] * cap + total - 1) / total;
    if (blocks < 1) blocks = 1;
    end += static_cast<int>(blocks);
    a.blk_end[l] = end;
  }
  for (int l = a.nlev; l < 4; ++l) a.blk_end[l] = end;
  return end;
}

static bool level_v4(const PoeLevelF& L) {
  return (L.n % 4 == 0) && (L.stride % 4 == 0) && (!L.drop || L.per_sample % 4 == 0) && aligned16(L.mu) && aligned16(L.lv) &&
         aligned16(L.out_mu) && aligned16(L.out_lv) && (!L.noise || (aligned16(L.noise) && aligned16(L.out_z)));
}
static bool level_v4(const PoeLevelB& L) {
  return (L.n % 4 == 0) && (L.stride % 4 == 0) && (!L.drop || L.per_sample % 4 == 0) && aligned16(L.mu) && aligned16(L.lv) &&
         aligned16(L.d_mu) && aligned16(L.d_lv) && (!L.g_mu || aligned16(L.g_mu)) && (!L.g_lv || aligned16(L.g_lv)) &&
         (!L.g_z || (aligned16(L.g_z) && aligned16(L.noise)));
}

template <int V, int NS>
static void launch_fwd(PoeFwdArgs& a, const PoeSubsets& ss, float eps, bool sp, cudaStream_t st) {
  const int grid = assign_blocks(a, V);
  if (ss.clip) {
    if (sp) poe_fwd_kernel<V, NS, true, true><<<grid, 256, 0, st>>>(a, ss, eps);
    else poe_fwd_kernel<V, NS, false, true><<<grid, 256, 0, st>>>(a, ss, eps);
  } else {
    if (sp) poe_fwd_kernel<V, NS, true, false><<<grid, 256, 0, st>>>(a, ss, eps);
    else poe_fwd_kernel<V, NS, false, false><<<grid, 256, 0, st>>>(a, ss, eps);
  }
}
template <int V, int NS>
static void launch_bwd(PoeBwdArgs& a, const PoeSubsets& ss, float eps, bool sp, cudaStream_t st) {
  const int grid = assign_blocks(a, V);
  // the fused clip is a compile-time variant: the un-clipped kernels keep the register budget they were tuned with
  if (ss.clip) {
    if (sp) poe_bwd_kernel<V, NS, true, true><<<grid, 256, 0, st>>>(a, ss, eps);
    else poe_bwd_kernel<V, NS, false, true><<<grid, 256, 0, st>>>(a, ss, eps);
  } else {
    if (sp) poe_bwd_kernel<V, NS, true, false><<<grid, 256, 0, st>>>(a, ss, eps);
    else poe_bwd_kernel<V, NS, false, false><<<grid, 256, 0, st>>>(a, ss, eps);
  }
}

static int run_fwd(PoeFwdArgs& a, const uint32_t* subset_masks, int n_subsets, float eps, int flags, cudaStream_t st,
                   const xhved_poe_opts* opts = nullptr) {
  PoeSubsets ss;
  if (int rc = fill_subsets(ss, subset_masks, n_subsets, opts)) return rc;
  bool v4 = true;
  for (int l = 0; l < a.nlev; ++l) {
    const PoeLevelF& L = a.lev[l];
    if (L.n <= 0 || L.stride < L.n || !L.mu || !L.lv || !L.out_mu || !L.out_lv) return XHVED_ERR_BAD_ARG;
    if ((L.out_z != nullptr) != (L.noise != nullptr)) return XHVED_ERR_BAD_ARG;
    if (L.drop && (L.per_sample <= 0 || L.n % L.per_sample)) return XHVED_ERR_BAD_SHAPE;
    v4 = v4 && level_v4(L);
  }
  const bool sp = (flags & XHVED_POE_STANDARD_PRIOR) != 0;
  ProfScope ps(K_POE_FWD, st);
  if (v4) {
    if (n_subsets == 1) launch_fwd<4, 1>(a, ss, eps, sp, st);
    else if (n_subsets <= 4) launch_fwd<4, 4>(a, ss, eps, sp, st);
    else launch_fwd<4, 15>(a, ss, eps, sp, st);
  } else {
    if (n_subsets == 1) launch_fwd<1, 1>(a, ss, eps, sp, st);
    else launch_fwd<1, 15>(a, ss, eps, sp, st);
  }
  return (int)cudaGetLastError();
}

static int run_bwd(PoeBwdArgs& a, const uint32_t* subset_masks, int n_subsets, float eps, int flags, cudaStream_t st,
                   const xhved_poe_opts* opts = nullptr) {
  PoeSubsets ss;
  if (int rc = fill_subsets(ss, subset_masks, n_subsets, opts)) return rc;
  bool v4 = true;
  for (int l = 0; l < a.nlev; ++l) {
    const PoeLevelB& L = a.lev[l];
    if (L.n <= 0 || L.stride < L.n || !L.mu || !L.lv || !L.d_mu || !L.d_lv) return XHVED_ERR_BAD_ARG;
    if (L.g_z && !L.noise) return XHVED_ERR_BAD_ARG;
    if (L.drop && (L.per_sample <= 0 || L.n % L.per_sample)) return XHVED_ERR_BAD_SHAPE;
    v4 = v4 && level_v4(L);
  }
  const bool sp = (flags & XHVED_POE_STANDARD_PRIOR) != 0;
  ProfScope ps(K_POE_BWD, st);
  if (v4) {
    if (n_subsets == 1) launch_bwd<4, 1>(a, ss, eps, sp, st);
    else if (n_subsets <= 4) launch_bwd<4, 4>(a, ss, eps, sp, st);
    else launch_bwd<4, 15>(a, ss, eps, sp, st);
  } else {
    if (n_subsets == 1) launch_bwd<1, 1>(a, ss, eps, sp, st);
    else launch_bwd<1, 15>(a, ss, eps, sp, st);
  }
  return (int)cudaGetLastError();
}

}  // namespace xhved

using namespace xhved;

extern "C" int xhved_version(void) { return 101; }

extern "C" int xhved_poe_fwd(const float* mu, const float* logvar, int64_t n, int64_t expert_stride, const uint32_t* subset_masks,
                             int n_subsets, const uint8_t* drop, int64_t per_sample, float eps, float* out_mu, float* out_logvar,
                             const float* noise, float* out_z, float* kld_out, int flags, void* stream) {
  PoeFwdArgs a = {};
  a.nlev = 1;
  a.lev[0] = PoeLevelF{mu, logvar, n, expert_stride, drop, per_sample, out_mu, out_logvar, noise, out_z, kld_out};
  return run_fwd(a, subset_masks, n_subsets, eps, flags, static_cast<cudaStream_t>(stream));
}

extern "C" int xhved_poe_bwd(const float* mu, const float* logvar, int64_t n, int64_t expert_stride, const uint32_t* subset_masks,
                             int n_subsets, const uint8_t* drop, int64_t per_sample, float eps, const float* g_mu, const float* g_logvar,
                             const float* noise, const float* g_z, const float* kld_scale, float* d_mu, float* d_logvar, int flags,
                             void* stream) {
  if (n_subsets < 1 || n_subsets > 15) return XHVED_ERR_BAD_ARG;
  PoeBwdArgs a = {};
  a.nlev = 1;
  PoeLevelB& L = a.lev[0];
  L.mu = mu, L.lv = logvar, L.n = n, L.stride = expert_stride, L.drop = drop, L.per_sample = per_sample;
  L.g_mu = g_mu, L.g_lv = g_logvar, L.noise = noise, L.g_z = g_z, L.d_mu = d_mu, L.d_lv = d_logvar;
  L.use_kld = kld_scale != nullptr;
  for (int s = 0; s < n_subsets; ++s) L.kld_scale[s] = kld_scale ? kld_scale[s] : 0.f;
  return run_bwd(a, subset_masks, n_subsets, eps, flags, static_cast<cudaStream_t>(stream));
}

extern "C" int xhved_poe_fwd_levels(const xhved_poe_level* levels, int n_levels, const uint32_t* subset_masks, int n_subsets, float eps,
                                    int flags, void* stream) {
  if (!levels || n_levels < 1 || n_levels > XHVED_POE_MAX_LEVELS) return XHVED_ERR_BAD_ARG;
  PoeFwdArgs a = {};
  a.nlev = n_levels;
  for (int l = 0; l < n_levels; ++l) {
    const xhved_poe_level& s = levels[l];
    a.lev[l] = PoeLevelF{s.mu, s.logvar, s.n, s.expert_stride, s.drop, s.per_sample, s.out_mu, s.out_logvar, s.noise, s.out_z, s.kld_out};
  }
  return run_fwd(a, subset_masks, n_subsets, eps, flags, static_cast<cudaStream_t>(stream));
}

extern "C" int xhved_poe_bwd_levels(const xhved_poe_level_grad* levels, int n_levels, const uint32_t* subset_masks, int n_subsets,
                                    float eps, int flags, void* stream) {
  if (!levels || n_levels < 1 || n_levels > XHVED_POE_MAX_LEVELS || n_subsets < 1 || n_subsets > 15) return XHVED_ERR_BAD_ARG;
  PoeBwdArgs a = {};
  a.nlev = n_levels;
  for (int l = 0; l < n_levels; ++l) {
    const xhved_poe_level_grad& s = levels[l];
    PoeLevelB& L = a.lev[l];
    L.mu = s.mu, L.lv = s.logvar, L.n = s.n, L.stride = s.expert_stride, L.drop = s.drop, L.per_sample = s.per_sample;
    L.g_mu = s.g_mu, L.g_lv = s.g_logvar, L.noise = s.noise, L.g_z = s.g_z, L.d_mu = s.d_mu, L.d_lv = s.d_logvar;
    L.use_kld = s.kld_scale != nullptr;
    for (int k = 0; k < n_subsets; ++k) L.kld_scale[k] = s.kld_scale ? s.kld_scale[k] : 0.f;
  }
  return run_bwd(a, subset_masks, n_subsets, eps, flags, static_cast<cudaStream_t>(stream));
}

static void fill_fwd_levels(PoeFwdArgs& a, const xhved_poe_level* levels, int n_levels) {
  a.nlev = n_levels;
  for (int l = 0; l < n_levels; ++l) {
    const xhved_poe_level& s = levels[l];
    a.lev[l] = PoeLevelF{s.mu, s.logvar, s.n, s.expert_stride, s.drop, s.per_sample, s.out_mu, s.out_logvar, s.noise, s.out_z, s.kld_out};
  }
}
static void fill_bwd_levels(PoeBwdArgs& a, const xhved_poe_level_grad* levels, int n_levels, int n_subsets) {
  a.nlev = n_levels;
  for (int l = 0; l < n_levels; ++l) {
    const xhved_poe_level_grad& s = levels[l];
    PoeLevelB& L = a.lev[l];
    L.mu = s.mu, L.lv = s.logvar, L.n = s.n, L.stride = s.expert_stride, L.drop = s.drop, L.per_sample = s.per_sample;
    L.g_mu = s.g_mu, L.g_lv = s.g_logvar, L.noise = s.noise, L.g_z = s.g_z, L.d_mu = s.d_mu, L.d_lv = s.d_logvar;
    L.use_kld = s.kld_scale != nullptr;
    for (int k = 0; k < n_subsets; ++k) L.kld_scale[k] = s.kld_scale ? s.kld_scale[k] : 0.f;
  }
}

extern "C" int xhved_poe_fwd_levels_opts(const xhved_poe_level* levels, int n_levels, const uint32_t* subset_masks, int n_subsets,
                                         const xhved_poe_opts* opts, void* stream) {
  if (!levels || !opts || n_levels < 1 || n_levels > XHVED_POE_MAX_LEVELS) return XHVED_ERR_BAD_ARG;
  PoeFwdArgs a = {};
  fill_fwd_levels(a, levels, n_levels);
  return run_fwd(a, subset_masks, n_subsets, opts->eps, opts->flags, static_cast<cudaStream_t>(stream), opts);
}
extern "C" int xhved_poe_bwd_levels_opts(const xhved_poe_level_grad* levels, int n_levels, const uint32_t* subset_masks, int n_subsets,
                                         const xhved_poe_opts* opts, void* stream) {
  if (!levels || !opts || n_levels < 1 || n_levels > XHVED_POE_MAX_LEVELS || n_subsets < 1 || n_subsets > 15) return XHVED_ERR_BAD_ARG;
  PoeBwdArgs a = {};
  fill_bwd_levels(a, levels, n_levels, n_subsets);
  return run_bwd(a, subset_masks, n_subsets, opts->eps, opts->flags, static_cast<cudaStream_t>(stream), opts);
}

// ---------------------------------------------------------------- stand-alone clip and ZeroLayerF
template <int V>
__global__ void __launch_bounds__(256) clip_fwd_kernel(const float* __restrict__ x, int64_t n, float lo, float hi, float* __restrict__ y) {
  const int64_t nvec = n / V;
  for (int64_t iv = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; iv < nvec;
       iv += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float a[V];
    ld<V>(x + iv * V, a);
#pragma unroll
    for (int j = 0; j < V; ++j) a[j] = (a[j] != a[j]) ? a[j] : fminf(fmaxf(a[j], lo), hi);   // NaN passes, like torch.clamp
    st<V>(y + iv * V, a);
  }
}
template <int V>
__global__ void __launch_bounds__(256) clip_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, int64_t n, float lo,
                                                        float hi, float* __restrict__ dx) {
  const int64_t nvec = n / V;
  for (int64_t iv = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; iv < nvec;
       iv += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float a[V], b[V];
    ld<V>(x + iv * V, a);
    ld<V>(g + iv * V, b);
#pragma unroll
    for (int j = 0; j < V; ++j) b[j] = (a[j] < lo || a[j] > hi) ? 0.f : b[j];
    st<V>(dx + iv * V, b);
  }
}
// y[b, :] = mask[b] ? 0 : x[b, :]   (ZeroLayerF forward AND backward, buildingblocks.py:308-323)
template <int V>
__global__ void __launch_bounds__(256) zero_rows_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask, int64_t n,
                                                         int64_t per_row, float* __restrict__ y) {
  const int64_t nvec = n / V;
  for (int64_t iv = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; iv < nvec;
       iv += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float a[V];
    const bool dead = mask[(iv * V) / per_row] != 0;
    if (!dead) {
      ld<V>(x + iv * V, a);
    } else {
#pragma unroll
      for (int j = 0; j < V; ++j) a[j] = 0.f;
    }
    st<V>(y + iv * V, a);
  }
}

extern "C" int xhved_clip_fwd(const float* x, int64_t n, float lo, float hi, float* y, void* stream) {
  if (n <= 0 || !x || !y || !(lo <= hi)) return XHVED_ERR_BAD_ARG;
  cudaStream_t st_ = static_cast<cudaStream_t>(stream);
  if (n % 4 == 0 && aligned16(x) && aligned16(y)) clip_fwd_kernel<4><<<grid_for(n / 4), 256, 0, st_>>>(x, n, lo, hi, y);
  else clip_fwd_kernel<1><<<grid_for(n), 256, 0, st_>>>(x, n, lo, hi, y);
  return (int)cudaGetLastError();
}
extern "C" int xhved_clip_bwd(const float* x, const float* g, int64_t n, float lo, float hi, float* dx, void* stream) {
  if (n <= 0 || !x || !g || !dx || !(lo <= hi)) return XHVED_ERR_BAD_ARG;
  cudaStream_t st_ = static_cast<cudaStream_t>(stream);
  if (n % 4 == 0 && aligned16(x) && aligned16(g) && aligned16(dx)) clip_bwd_kernel<4><<<grid_for(n / 4), 256, 0, st_>>>(x, g, n, lo, hi, dx);
  else clip_bwd_kernel<1><<<grid_for(n), 256, 0, st_>>>(x, g, n, lo, hi, dx);
  return (int)cudaGetLastError();
}
extern "C" int xhved_zero_rows(const float* x, const uint8_t* mask, int64_t rows, int64_t per_row, float* y, void* stream) {
  if (rows <= 0 || per_row <= 0 || !x || !mask || !y) return XHVED_ERR_BAD_ARG;
  cudaStream_t st_ = static_cast<cudaStream_t>(stream);
  const int64_t n = rows * per_row;
  if (per_row % 4 == 0 && aligned16(x) && aligned16(y)) zero_rows_kernel<4><<<grid_for(n / 4), 256, 0, st_>>>(x, mask, n, per_row, y);
  else zero_rows_kernel<1><<<grid_for(n), 256, 0, st_>>>(x, mask, n, per_row, y);
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
