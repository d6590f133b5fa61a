// K3: everything of the ViL block behind the mLSTM cell, fused -- sm_100a.
//
// Restates, per token, vision_lstm.py:271-287 (MultiHeadLayerNorm: per-head normalisation over DH, weight
// 1+w, eps 1e-5), :437 (learnable skip of the conv activation), :440 (SiLU(z) gate), :443 (proj_down),
// :446-451 (un-flip, folded into the index map), vision_lstm_util.py:171-175 (residual add) and the token ->
// NCDHW transpose of UxLSTMEnc_3d.py:61.  One CTA = 128 tokens; thread r owns token r.
#include "vil_common.cuh"

namespace xhved {

template <int C>
struct PostSmem {
  static constexpr int E = 2 * C;
  static constexpr int WDT = 0;            // proj_down transposed: (E, C)
  static constexpr int OW = WDT + E * C;   // outnorm weight (E)
  static constexpr int SK = OW + E;        // learnable skip (E)
  static constexpr int TOTAL = SK + E;
};

// load one head's row (DH values) of a tile-native bf16 tile
template <int DH>
__device__ __forceinline__ void load_h_row(const unsigned char* tile, int r, float* hv) {
#pragma unroll
  for (int cg = 0; cg < DH / 8; ++cg) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(tile + tile_off16(kTok, r, cg)));
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    hv[cg * 8 + 0] = a.x, hv[cg * 8 + 1] = a.y, hv[cg * 8 + 2] = b.x, hv[cg * 8 + 3] = b.y;
    hv[cg * 8 + 4] = c.x, hv[cg * 8 + 5] = c.y, hv[cg * 8 + 6] = d.x, hv[cg * 8 + 7] = d.y;
  }
}

template <int C>
__global__ void __launch_bounds__(kTok) vil_post_fwd_kernel(const float* __restrict__ x, const unsigned char* __restrict__ h_tiles,
                                                             const float* __restrict__ act, const float* __restrict__ z,
                                                             xhved_vil_params p, VilGeom g, float* __restrict__ y) {
  using L = PostSmem<C>;
  constexpr int E = L::E, DH = E / 4, DHP = DH < 16 ? 16 : DH;
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  const int b = blockIdx.x / g.nc, ch = blockIdx.x % g.nc;
  for (int i = tid; i < E * C; i += kTok) {
    const int c = i / E, e = i % E;
    sm[L::WDT + e * C + c] = __ldg(p.proj_down_weight + i);
  }
  stage(sm + L::OW, p.outnorm_weight, E);
  stage(sm + L::SK, p.learnable_skip, E);
  __syncthreads();
  const int tau = ch * kTok + tid;
  if (tau >= g.S) return;
  const int n = g.reverse ? g.S - 1 - tau : tau;
  const size_t tm_base = (static_cast<size_t>(b) * g.nc + ch) * E * kTok + tid;

  float out[C];
#pragma unroll
  for (int c = 0; c < C; ++c) out[c] = 0.f;
#pragma unroll 1
  for (int head = 0; head < 4; ++head) {
    const size_t tile = (static_cast<size_t>(b) * 4 + head) * g.nc + ch;
    float hv[DH];
    load_h_row<DH>(h_tiles + tile * (kTok * DHP * 2), tid, hv);
    float mean = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) mean += hv[d];
    mean *= (1.f / DH);
    float var = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) var += (hv[d] - mean) * (hv[d] - mean);
    const float rstd = rsqrtf(var * (1.f / DH) + 1e-5f);
#pragma unroll
    for (int d = 0; d < DH; ++d) {
      const int e = head * DH + d;
      const float hn = (hv[d] - mean) * rstd * (1.f + sm[L::OW + e]);
      const float a = __ldg(act + tm_base + static_cast<size_t>(e) * kTok);
      const float zz = __ldg(z + tm_base + static_cast<size_t>(e) * kTok);
      const float hg = (hn + sm[L::SK + e] * a) * silu(zz);
      const float* w = sm + L::WDT + e * C;
#pragma unroll
      for (int c = 0; c < C; c += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(w + c);
        out[c] += w4.x * hg, out[c + 1] += w4.y * hg, out[c + 2] += w4.z * hg, out[c + 3] += w4.w * hg;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < C; ++c) y[b * g.ysb + n * g.ysn + c * g.ysc] = __ldg(x + b * g.xsb + n * g.xsn + c * g.xsc) + out[c];
}

template <int C>
static int launch_post_fwd(const float* x, const void* h, const float* act, const float* z, const xhved_vil_params* p, const VilGeom& g,
                           float* y, cudaStream_t st) {
  const size_t smem = PostSmem<C>::TOTAL * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(vil_post_fwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  ProfScope ps(K_VIL_POST_FWD, st);
  vil_post_fwd_kernel<C><<<g.B * g.nc, kTok, smem, st>>>(x, (const unsigned char*)h, act, z, *p, g, y);
  return (int)cudaGetLastError();
}


// ------------------------------------------------------------------ backward
template <int C>
struct PostBwdSmem {
  static constexpr int E = 2 * C;
  static constexpr int WDT = 0;                 // proj_down transposed (E, C)
  static constexpr int OW = WDT + E * C;
  static constexpr int SK = OW + E;
  static constexpr int ACC_SK = SK + E;         // per-CTA partial of d learnable_skip
  static constexpr int ACC_OW = ACC_SK + E;     // per-CTA partial of d outnorm.weight
  static constexpr int HG = ACC_OW + E;         // (128, E+1) gated activations
  static constexpr int DY = HG + kTok * (E + 1);  // (128, C+1)
  static constexpr int TOTAL = DY + kTok * (C + 1);
};

template <int C>
__global__ void __launch_bounds__(kTok) vil_post_bwd_kernel(const float* __restrict__ dy, const unsigned char* __restrict__ h_tiles,
                                                             const float* __restrict__ act, const float* __restrict__ z,
                                                             xhved_vil_params p, VilGeom g, unsigned char* __restrict__ dh_tiles,
                                                             float* __restrict__ d_act, float* __restrict__ dz, xhved_vil_grads gr) {
  using L = PostBwdSmem<C>;
  constexpr int E = L::E, DH = E / 4, DHP = DH < 16 ? 16 : DH;
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  const int b = blockIdx.x / g.nc, ch = blockIdx.x % g.nc;
  for (int i = tid; i < E * C; i += kTok) {
    const int c = i / E, e = i % E;
    sm[L::WDT + e * C + c] = __ldg(p.proj_down_weight + i);
  }
  stage(sm + L::OW, p.outnorm_weight, E);
  stage(sm + L::SK, p.learnable_skip, E);
  for (int i = tid; i < 2 * E; i += kTok) sm[L::ACC_SK + i] = 0.f;
  const int tau = ch * kTok + tid;
  const bool valid = tau < g.S;
  const int n = g.reverse ? g.S - 1 - tau : tau;
  float dyr[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    dyr[c] = valid ? __ldg(dy + b * g.ysb + n * g.ysn + c * g.ysc) : 0.f;
    sm[L::DY + tid * (C + 1) + c] = dyr[c];
  }
  __syncthreads();
  const size_t tm_base = (static_cast<size_t>(b) * g.nc + ch) * E * kTok + tid;
#pragma unroll 1
  for (int head = 0; head < 4; ++head) {
    const size_t tile = (static_cast<size_t>(b) * 4 + head) * g.nc + ch;
    float hv[DH], gg[DH];
    load_h_row<DH>(h_tiles + tile * (kTok * DHP * 2), tid, hv);
    float mean = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) mean += hv[d];
    mean *= (1.f / DH);
    float var = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) var += (hv[d] - mean) * (hv[d] - mean);
    const float rstd = rsqrtf(var * (1.f / DH) + 1e-5f);
    float mean_g = 0.f, mean_gx = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) {
      const int e = head * DH + d;
      const float xhat = (hv[d] - mean) * rstd;
      const float ow1 = 1.f + sm[L::OW + e];
      const float a = __ldg(act + tm_base + static_cast<size_t>(e) * kTok);
      const float zz = __ldg(z + tm_base + static_cast<size_t>(e) * kTok);
      const float sz = silu(zz);
      const float hs = xhat * ow1 + sm[L::SK + e] * a;
      float dhg = 0.f;
      const float* w = sm + L::WDT + e * C;
#pragma unroll
      for (int c = 0; c < C; c += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(w + c);
        dhg += w4.x * dyr[c] + w4.y * dyr[c + 1] + w4.z * dyr[c + 2] + w4.w * dyr[c + 3];
      }
      const float dhs = dhg * sz;
      sm[L::HG + tid * (E + 1) + e] = valid ? hs * sz : 0.f;
      dz[tm_base + static_cast<size_t>(e) * kTok] = dhg * hs * dsilu(zz);
      d_act[tm_base + static_cast<size_t>(e) * kTok] = dhs * sm[L::SK + e];
      warp_acc(sm + L::ACC_SK + e, dhs * a);
      warp_acc(sm + L::ACC_OW + e, dhs * xhat);
      gg[d] = dhs * ow1;
      hv[d] = xhat;
      mean_g += gg[d];
      mean_gx += gg[d] * xhat;
    }
    mean_g *= (1.f / DH);
    mean_gx *= (1.f / DH);
    float o[DHP];
#pragma unroll
    for (int d = 0; d < DHP; ++d) o[d] = (d < DH && valid) ? rstd * (gg[d < DH ? d : 0] - mean_g - hv[d < DH ? d : 0] * mean_gx) : 0.f;
    unsigned char* dst = dh_tiles + tile * (kTok * DHP * 2);
#pragma unroll
    for (int cg = 0; cg < DHP / 8; ++cg)
      *reinterpret_cast<uint4*>(dst + tile_off16(kTok, tid, cg)) =
          make_uint4(pack_bf16x2(o[cg * 8], o[cg * 8 + 1]), pack_bf16x2(o[cg * 8 + 2], o[cg * 8 + 3]),
                     pack_bf16x2(o[cg * 8 + 4], o[cg * 8 + 5]), pack_bf16x2(o[cg * 8 + 6], o[cg * 8 + 7]));
  }
  __syncthreads();
  for (int e = tid; e < E; e += kTok) {
    atomicAdd(gr.learnable_skip + e, sm[L::ACC_SK + e]);
    atomicAdd(gr.outnorm_weight + e, sm[L::ACC_OW + e]);
  }
  // d proj_down[c][e] += sum_tok dy[tok][c] * hg[tok][e]
  outer_accumulate(sm + L::DY, C + 1, C, sm + L::HG, E + 1, E, kTok, gr.proj_down_weight);
}

template <int C>
static int launch_post_bwd(const float* dy, const void* h, const float* act, const float* z, const xhved_vil_params* p, const VilGeom& g,
                           void* dh, float* d_act, float* dz, const xhved_vil_grads* gr, cudaStream_t st) {
  const size_t smem = PostBwdSmem<C>::TOTAL * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(vil_post_bwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  ProfScope ps(K_VIL_POST_BWD, st);
  vil_post_bwd_kernel<C><<<g.B * g.nc, kTok, smem, st>>>(dy, (const unsigned char*)h, act, z, *p, g, (unsigned char*)dh, d_act, dz, *gr);
  return (int)cudaGetLastError();
}

}  // namespace xhved

using namespace xhved;

extern "C" int xhved_vil_post_fwd(const float* x, const void* h_tiles, const float* act, const float* z, const xhved_vil_params* p,
                                  const xhved_vil_shape* sh, float* y, void* stream) {
  VilGeom g;
  if (int rc = vil_validate(sh, &g)) return rc;
  if (!x || !h_tiles || !act || !z || !p || !y) return XHVED_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (sh->C) {
    case 16: return launch_post_fwd<16>(x, h_tiles, act, z, p, g, y, st);
    case 32: return launch_post_fwd<32>(x, h_tiles, act, z, p, g, y, st);
    case 64: return launch_post_fwd<64>(x, h_tiles, act, z, p, g, y, st);
    default: return XHVED_ERR_UNSUPPORTED_DIM;
  }
}

extern "C" int xhved_vil_post_bwd(const float* dy, const void* h_tiles, const float* act, const float* z, const xhved_vil_params* p,
                                  const xhved_vil_shape* sh, void* dh_tiles, float* d_act, float* dz, const xhved_vil_grads* g,
                                  void* stream) {
  VilGeom geo;
  if (int rc = vil_validate(sh, &geo)) return rc;
  if (!dy || !h_tiles || !act || !z || !p || !dh_tiles || !d_act || !dz || !g) return XHVED_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (sh->C) {
    case 16: return launch_post_bwd<16>(dy, h_tiles, act, z, p, geo, dh_tiles, d_act, dz, g, st);
    case 32: return launch_post_bwd<32>(dy, h_tiles, act, z, p, geo, dh_tiles, d_act, dz, g, st);
    case 64: return launch_post_bwd<64>(dy, h_tiles, act, z, p, geo, dh_tiles, d_act, dz, g, st);
    default: return XHVED_ERR_UNSUPPORTED_DIM;
  }
}
