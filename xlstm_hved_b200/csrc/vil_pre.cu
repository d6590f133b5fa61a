// K2: everything of the ViL block in front of the mLSTM cell, fused -- sm_100a.
//
// Restates, per token, vision_lstm.py:252-259 (LayerNorm, weight 1+w, eps 1e-5), :427-428 (proj_up, split),
// :213-221 (causal depthwise conv1d k=4), :432 (SiLU), :158-168 (block-diagonal q/k/v projections; v from the
// pre-conv branch, :433-435), :305-318 (input / forget gate pre-activations from [q,k,v]).  The direction flip
// (:419-424) is folded into the token index map; tokens are read straight from the strided NCDHW feature
// (UxLSTMEnc_3d.py:59).  Outputs land in the cell's tile-native bf16 layout, so the cell kernels can bulk-copy
// MMA-ready operand tiles.
//
// One CTA = 128 consecutive tokens (in traversal order) = one cell chunk for all heads; 160 threads: thread r <
// 128 owns token r, threads 128..130 recompute the 3-token conv halo.
#include "vil_common.cuh"

namespace xhved {

template <int C>
struct PreSmem {
  static constexpr int E = 2 * C;
  static constexpr int XM_LD = E + 1;                 // odd row pitch: conflict-free column access
  static constexpr int XM_ROWS = kTok + 3;
  // float offsets
  static constexpr int W_UP = 0;                      // (2E, C)
  static constexpr int XM = W_UP + 2 * E * C;         // (131, E+1)
  static constexpr int CONV_W = XM + (XM_ROWS * XM_LD + 3) / 4 * 4; // (E, 4), 16-byte aligned
  static constexpr int CONV_B = CONV_W + E * 4;
  static constexpr int WQ = CONV_B + E;               // (E/4, 4, 4) = E*4
  static constexpr int WK = WQ + E * 4;
  static constexpr int WV = WK + E * 4;
  static constexpr int WI = WV + E * 4;               // (4, 3E)
  static constexpr int WF = WI + 4 * 3 * E;
  static constexpr int NW = WF + 4 * 3 * E;           // (C)
  static constexpr int TOTAL = NW + C;
};

// LayerNorm of one token held in registers; returns xhat*(1+w) in xn, optionally xhat / rstd
template <int C>
__device__ __forceinline__ void layernorm_token(const float* xin, const float* nw, float* xn, float* rstd_out) {
  float mean = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) mean += xin[c];
  mean *= (1.f / C);
  float var = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float d = xin[c] - mean;
    var += d * d;
  }
  const float rstd = rsqrtf(var * (1.f / C) + 1e-5f);
#pragma unroll
  for (int c = 0; c < C; ++c) xn[c] = (xin[c] - mean) * rstd * (1.f + nw[c]);
  if (rstd_out) *rstd_out = rstd;
}

template <int C>
__device__ __forceinline__ float dot_row(const float* xn, const float* wrow) {
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < C; c += 4) {
    const float4 w = *reinterpret_cast<const float4*>(wrow + c);
    acc += xn[c] * w.x + xn[c + 1] * w.y + xn[c + 2] * w.z + xn[c + 3] * w.w;
  }
  return acc;
}

template <int C>
__global__ void __launch_bounds__(160) vil_pre_fwd_kernel(const float* __restrict__ x, xhved_vil_params p, VilGeom g,
                                                           unsigned char* __restrict__ q_tiles, unsigned char* __restrict__ k_tiles,
                                                           unsigned char* __restrict__ v_tiles, float* __restrict__ igp,
                                                           float* __restrict__ fgp, float* __restrict__ act_out,
                                                           float* __restrict__ z_out) {
  using L = PreSmem<C>;
  constexpr int E = L::E;
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  const int b = blockIdx.x / g.nc, ch = blockIdx.x % g.nc;

  stage(sm + L::W_UP, p.proj_up_weight, 2 * E * C);
  stage(sm + L::CONV_W, p.conv_weight, E * 4);
  stage(sm + L::CONV_B, p.conv_bias, E);
  stage(sm + L::WQ, p.q_weight, E * 4);
  stage(sm + L::WK, p.k_weight, E * 4);
  stage(sm + L::WV, p.v_weight, E * 4);
  stage(sm + L::WI, p.igate_weight, 4 * 3 * E);
  stage(sm + L::WF, p.fgate_weight, 4 * 3 * E);
  stage(sm + L::NW, p.norm_weight, C);

  // token owned by this thread (traversal order tau); halo threads own tau0-3..tau0-1
  const bool is_main = tid < kTok, is_halo = tid >= kTok && tid < kTok + 3;
  const int tau = is_main ? ch * kTok + tid : ch * kTok - 3 + (tid - kTok);
  const bool valid = (is_main || is_halo) && tau >= 0 && tau < g.S;
  const int n = g.reverse ? g.S - 1 - tau : tau;
  float xin[C];
#pragma unroll
  for (int c = 0; c < C; ++c) xin[c] = valid ? __ldg(x + b * g.xsb + n * g.xsn + c * g.xsc) : 0.f;
  __syncthreads();

  float xn[C];
  layernorm_token<C>(xin, sm + L::NW, xn, nullptr);
  const int xm_row = is_main ? tid + 3 : tid - kTok;
  if (is_main || is_halo) {
    float* xm = sm + L::XM + xm_row * L::XM_LD;
#pragma unroll 1
    for (int e = 0; e < E; ++e) xm[e] = valid ? dot_row<C>(xn, sm + L::W_UP + e * C) : 0.f;
  }
  const size_t tm_base = (static_cast<size_t>(b) * g.nc + ch) * E * kTok;   // token-minor (B, nc, E, 128)
  if (is_main) {
#pragma unroll 1
    for (int e = 0; e < E; ++e) z_out[tm_base + static_cast<size_t>(e) * kTok + tid] = valid ? dot_row<C>(xn, sm + L::W_UP + (E + e) * C) : 0.f;
  }
  __syncthreads();
  if (!is_main) return;

  // conv + SiLU + block-diagonal q,k,v + gates, 8 channels at a time
  float ig_acc[4], fg_acc[4];
#pragma unroll
  for (int h = 0; h < 4; ++h) ig_acc[h] = __ldg(p.igate_bias + h), fg_acc[h] = __ldg(p.fgate_bias + h);
  const float* xm0 = sm + L::XM + tid * L::XM_LD;   // rows tid..tid+3 <-> tokens tau-3..tau
  const bool rowvalid = tau < g.S;
#pragma unroll 1
  for (int e8 = 0; e8 < E; e8 += 8) {
    float a8[8], xm8[8], q8[8], k8[8], v8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = e8 + j;
      const float4 w = *reinterpret_cast<const float4*>(sm + L::CONV_W + e * 4);
      const float conv = sm[L::CONV_B + e] + w.x * xm0[e] + w.y * xm0[L::XM_LD + e] + w.z * xm0[2 * L::XM_LD + e] +
                         w.w * xm0[3 * L::XM_LD + e];
      a8[j] = silu(conv);
      xm8[j] = xm0[3 * L::XM_LD + e];
      act_out[tm_base + static_cast<size_t>(e) * kTok + tid] = rowvalid ? a8[j] : 0.f;
    }
#pragma unroll
    for (int blk = 0; blk < 2; ++blk) {
      const int wb = ((e8 >> 2) + blk) * 16;   // (block, out, in) 4x4
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        float aq = 0.f, ak = 0.f, av = 0.f;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          aq += sm[L::WQ + wb + o * 4 + d] * a8[blk * 4 + d];
          ak += sm[L::WK + wb + o * 4 + d] * a8[blk * 4 + d];
          av += sm[L::WV + wb + o * 4 + d] * xm8[blk * 4 + d];
        }
        q8[blk * 4 + o] = aq, k8[blk * 4 + o] = ak, v8[blk * 4 + o] = av;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = e8 + j;
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        ig_acc[h] += sm[L::WI + h * 3 * E + e] * q8[j] + sm[L::WI + h * 3 * E + E + e] * k8[j] + sm[L::WI + h * 3 * E + 2 * E + e] * v8[j];
        fg_acc[h] += sm[L::WF + h * 3 * E + e] * q8[j] + sm[L::WF + h * 3 * E + E + e] * k8[j] + sm[L::WF + h * 3 * E + 2 * E + e] * v8[j];
      }
    }
    // 8 consecutive channels = one 16-byte group of one head's tile row
    const int head = e8 / g.DH, d0 = e8 % g.DH;
    const size_t tile = (static_cast<size_t>(b) * g.NH + head) * g.nc + ch;
    const size_t off = tile * (kTok * g.DHP * 2) + tile_off16(kTok, tid, d0 / 8);
    uint4 uq = make_uint4(0, 0, 0, 0), uk = uq, uv = uq;
    if (rowvalid) {
      uq = make_uint4(pack_bf16x2(q8[0], q8[1]), pack_bf16x2(q8[2], q8[3]), pack_bf16x2(q8[4], q8[5]), pack_bf16x2(q8[6], q8[7]));
      uk = make_uint4(pack_bf16x2(k8[0], k8[1]), pack_bf16x2(k8[2], k8[3]), pack_bf16x2(k8[4], k8[5]), pack_bf16x2(k8[6], k8[7]));
      uv = make_uint4(pack_bf16x2(v8[0], v8[1]), pack_bf16x2(v8[2], v8[3]), pack_bf16x2(v8[4], v8[5]), pack_bf16x2(v8[6], v8[7]));
    }
    *reinterpret_cast<uint4*>(q_tiles + off) = uq;
    *reinterpret_cast<uint4*>(k_tiles + off) = uk;
    *reinterpret_cast<uint4*>(v_tiles + off) = uv;
    if (g.DHP > g.DH && d0 + 8 == g.DH) {   // zero the padding column groups (DH = 8 padded to 16)
      const uint4 zz = make_uint4(0, 0, 0, 0);
      for (int cg = g.DH / 8; cg < g.DHP / 8; ++cg) {
        const size_t o2 = tile * (kTok * g.DHP * 2) + tile_off16(kTok, tid, cg);
        *reinterpret_cast<uint4*>(q_tiles + o2) = zz;
        *reinterpret_cast<uint4*>(k_tiles + o2) = zz;
        *reinterpret_cast<uint4*>(v_tiles + o2) = zz;
      }
    }
  }
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const size_t o = (static_cast<size_t>(b) * g.NH + h) * g.Sp + ch * kTok + tid;
    igp[o] = rowvalid ? ig_acc[h] : -1e30f;
    fgp[o] = rowvalid ? fg_acc[h] : 1e30f;
  }
}

template <int C>
static int launch_pre_fwd(const float* x, const xhved_vil_params* p, const VilGeom& g, void* q, void* k, void* v, float* ig, float* fg,
                          float* act, float* z, cudaStream_t st) {
  const size_t smem = PreSmem<C>::TOTAL * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(vil_pre_fwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  vil_pre_fwd_kernel<C><<<g.B * g.nc, 160, smem, st>>>(x, *p, g, (unsigned char*)q, (unsigned char*)k, (unsigned char*)v, ig, fg, act, z);
  return (int)cudaGetLastError();
}

}  // namespace xhved

using namespace xhved;

extern "C" int xhved_vil_pre_fwd(const float* x, const xhved_vil_params* p, const xhved_vil_shape* sh, void* q_tiles, void* k_tiles,
                                 void* v_tiles, float* ig_padded, float* fg_padded, float* act, float* z, void* stream) {
  VilGeom g;
  if (int rc = vil_validate(sh, &g)) return rc;
  if (!x || !p || !q_tiles || !k_tiles || !v_tiles || !ig_padded || !fg_padded || !act || !z) return XHVED_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (sh->C) {
    case 16: return launch_pre_fwd<16>(x, p, g, q_tiles, k_tiles, v_tiles, ig_padded, fg_padded, act, z, st);
    case 32: return launch_pre_fwd<32>(x, p, g, q_tiles, k_tiles, v_tiles, ig_padded, fg_padded, act, z, st);
    case 64: return launch_pre_fwd<64>(x, p, g, q_tiles, k_tiles, v_tiles, ig_padded, fg_padded, act, z, st);
    default: return XHVED_ERR_UNSUPPORTED_DIM;
  }
}
