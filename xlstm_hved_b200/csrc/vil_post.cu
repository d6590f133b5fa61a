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
  vil_post_fwd_kernel<C><<<g.B * g.nc, kTok, smem, st>>>(x, (const unsigned char*)h, act, z, *p, g, y);
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
