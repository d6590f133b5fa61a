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
  ProfScope ps(K_VIL_PRE_FWD, st);
  vil_pre_fwd_kernel<C><<<g.B * g.nc, 160, smem, st>>>(x, *p, g, (unsigned char*)q, (unsigned char*)k, (unsigned char*)v, ig, fg, act, z);
  return (int)cudaGetLastError();
}


// ------------------------------------------------------------------ backward, kernel A
// Recomputes the forward up to q,k,v for 128 tokens (+3 halo), then pulls dq,dk,dv,dig,dfg and the skip-path
// d_act back to (a) dconv = d(conv pre-activation) and (b) dxm_v = d(x_mlstm) through the v projection, both
// written token-minor for kernel B; accumulates gate, q/k/v-projection and conv parameter gradients.
template <int C>
struct PreBwdASmem {
  using F = PreSmem<C>;
  static constexpr int E = 2 * C;
  static constexpr int ACC = F::TOTAL;                 // per-CTA partial sums, see offsets below
  static constexpr int A_WQ = 0, A_WK = E * 4, A_WV = 2 * E * 4, A_CW = 3 * E * 4, A_CB = 4 * E * 4, A_GB = 4 * E * 4 + E;
  static constexpr int ACC_N = A_GB + 8;
  static constexpr int DG = ACC + (ACC_N + 3) / 4 * 4;  // (128, 9): dig[4] | dfg[4] per token
  // (128, E) staging of q / k / v per token, column-skewed by the row index (conflict-free).  When it fits
  // (C >= 64) it ALIASES the proj_up weights, which are dead once x_mlstm has been recomputed.
  static constexpr bool ALIAS = 2 * E * C >= kTok * E;
  static constexpr int ST = ALIAS ? F::W_UP : DG + kTok * 9;
  static constexpr int TOTAL = ALIAS ? DG + kTok * 9 : ST + kTok * E;
};

template <int C>
__global__ void __launch_bounds__(160) vil_pre_bwd_a_kernel(const float* __restrict__ x, xhved_vil_params p, VilGeom g,
                                                             const float* __restrict__ dq, const float* __restrict__ dk,
                                                             const float* __restrict__ dv, const float* __restrict__ dig,
                                                             const float* __restrict__ dfg, const float* __restrict__ d_act,
                                                             float* __restrict__ dconv_out, float* __restrict__ dxmv_out,
                                                             xhved_vil_grads gr) {
  using L = PreSmem<C>;
  using LB = PreBwdASmem<C>;
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
  for (int i = tid; i < LB::ACC_N; i += blockDim.x) sm[LB::ACC + i] = 0.f;

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
  float dgi[4], dgf[4];
  const bool rowvalid = is_main && tau < g.S;
  if (is_main) {
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const size_t o = (static_cast<size_t>(b) * g.NH + h) * g.Sp + ch * kTok + tid;
      dgi[h] = rowvalid ? __ldg(dig + o) : 0.f;
      dgf[h] = rowvalid ? __ldg(dfg + o) : 0.f;
      sm[LB::DG + tid * 9 + h] = dgi[h];
      sm[LB::DG + tid * 9 + 4 + h] = dgf[h];
    }
  }
  __syncthreads();
  float* acc = sm + LB::ACC;
  float* stg = sm + LB::ST + (is_main ? tid : 0) * E;
  const size_t tm_base = (static_cast<size_t>(b) * g.nc + ch) * E * kTok + tid;
  const float* xm0 = sm + L::XM + (is_main ? tid : 0) * L::XM_LD;

  // three passes over the channels: part 0 stages q, 1 stages k, 2 stages v (for the gate-weight outer products);
  // pass 0 additionally does all the per-token backward work.
#pragma unroll 1
  for (int part = 0; part < 3; ++part) {
    if (is_main) {
#pragma unroll 1
      for (int e8 = 0; e8 < E; e8 += 8) {
        float a8[8], xm8[8], cv8[8], q8[8], k8[8], v8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int e = e8 + j;
          const float4 w = *reinterpret_cast<const float4*>(sm + L::CONV_W + e * 4);
          cv8[j] = sm[L::CONV_B + e] + w.x * xm0[e] + w.y * xm0[L::XM_LD + e] + w.z * xm0[2 * L::XM_LD + e] + w.w * xm0[3 * L::XM_LD + e];
          a8[j] = silu(cv8[j]);
          xm8[j] = xm0[3 * L::XM_LD + e];
        }
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          const int wb = ((e8 >> 2) + blk) * 16;
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
        for (int j = 0; j < 8; ++j) stg[(e8 + j + tid) & (E - 1)] = rowvalid ? (part == 0 ? q8[j] : part == 1 ? k8[j] : v8[j]) : 0.f;
        if (part == 0) {
          // upstream gradients of q,k,v for these 8 channels (+ the gate paths, vision_lstm.py:305-318)
          const int head = e8 / g.DH, d0 = e8 % g.DH;
          const size_t row = ((static_cast<size_t>(b) * g.NH + head) * g.Sp + ch * kTok + tid) * g.DHP + d0;
          float gq[8], gk[8], gv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int e = e8 + j;
            gq[j] = rowvalid ? __ldg(dq + row + j) : 0.f;
            gk[j] = rowvalid ? __ldg(dk + row + j) : 0.f;
            gv[j] = rowvalid ? __ldg(dv + row + j) : 0.f;
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              gq[j] += dgi[h] * sm[L::WI + h * 3 * E + e] + dgf[h] * sm[L::WF + h * 3 * E + e];
              gk[j] += dgi[h] * sm[L::WI + h * 3 * E + E + e] + dgf[h] * sm[L::WF + h * 3 * E + E + e];
              gv[j] += dgi[h] * sm[L::WI + h * 3 * E + 2 * E + e] + dgf[h] * sm[L::WF + h * 3 * E + 2 * E + e];
            }
          }
          float da8[8], dxv8[8];
#pragma unroll
          for (int blk = 0; blk < 2; ++blk) {
            const int wb = ((e8 >> 2) + blk) * 16;
#pragma unroll
            for (int d = 0; d < 4; ++d) {
              float sa = 0.f, sv = 0.f;
#pragma unroll
              for (int o = 0; o < 4; ++o) {
                sa += sm[L::WQ + wb + o * 4 + d] * gq[blk * 4 + o] + sm[L::WK + wb + o * 4 + d] * gk[blk * 4 + o];
                sv += sm[L::WV + wb + o * 4 + d] * gv[blk * 4 + o];
                // d q_proj[b][o][d] += gq[o] * act[d]  etc. (reduced over the CTA's tokens)
                warp_acc(acc + LB::A_WQ + wb + o * 4 + d, gq[blk * 4 + o] * a8[blk * 4 + d]);
                warp_acc(acc + LB::A_WK + wb + o * 4 + d, gk[blk * 4 + o] * a8[blk * 4 + d]);
                warp_acc(acc + LB::A_WV + wb + o * 4 + d, gv[blk * 4 + o] * xm8[blk * 4 + d]);
              }
              da8[blk * 4 + d] = sa, dxv8[blk * 4 + d] = sv;
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int e = e8 + j;
            const float dact = da8[j] + (rowvalid ? __ldg(d_act + tm_base + static_cast<size_t>(e) * kTok) : 0.f);
            const float dc = rowvalid ? dact * dsilu(cv8[j]) : 0.f;
            dconv_out[tm_base + static_cast<size_t>(e) * kTok] = dc;
            dxmv_out[tm_base + static_cast<size_t>(e) * kTok] = rowvalid ? dxv8[j] : 0.f;
            warp_acc(acc + LB::A_CB + e, dc);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) warp_acc(acc + LB::A_CW + e * 4 + jj, dc * xm0[jj * L::XM_LD + e]);
          }
        }
      }
      if (part == 0) {
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          warp_acc(acc + LB::A_GB + h, dgi[h]);
          warp_acc(acc + LB::A_GB + 4 + h, dgf[h]);
        }
      }
    }
    __syncthreads();
    // d igate.weight[h][part*E + e] += sum_tok dig[tok][h] * qkv_part[tok][e]   (same for fgate)
    for (int idx = tid; idx < 8 * E; idx += blockDim.x) {
      const int hh = idx / E, e = idx % E;
      float a = 0.f;
#pragma unroll 4
      for (int t = 0; t < kTok; ++t) a += sm[LB::DG + t * 9 + hh] * sm[LB::ST + t * E + ((e + t) & (E - 1))];
      float* dst = (hh < 4 ? gr.igate_weight + hh * 3 * E : gr.fgate_weight + (hh - 4) * 3 * E) + part * E + e;
      atomicAdd(dst, a);
    }
    __syncthreads();
  }
  for (int i = tid; i < E * 4; i += blockDim.x) {
    atomicAdd(gr.q_weight + i, acc[LB::A_WQ + i]);
    atomicAdd(gr.k_weight + i, acc[LB::A_WK + i]);
    atomicAdd(gr.v_weight + i, acc[LB::A_WV + i]);
    atomicAdd(gr.conv_weight + i, acc[LB::A_CW + i]);
  }
  for (int i = tid; i < E; i += blockDim.x) atomicAdd(gr.conv_bias + i, acc[LB::A_CB + i]);
  if (tid < 4) atomicAdd(gr.igate_bias + tid, acc[LB::A_GB + tid]);
  else if (tid < 8) atomicAdd(gr.fgate_bias + tid - 4, acc[LB::A_GB + tid]);
}

// ------------------------------------------------------------------ backward, kernel B
// d x_mlstm = dxm_v + transposed causal conv of dconv (tokens tau..tau+3); [d x_mlstm | dz] -> proj_up^T -> LayerNorm
// backward -> dx = dy + ...; accumulates proj_up and norm weight gradients.
template <int C>
struct PreBwdBSmem {
  static constexpr int E = 2 * C;
  static constexpr int W_UP = 0;                       // (2E, C)
  static constexpr int CONV_W = W_UP + 2 * E * C;      // (E, 4)
  static constexpr int NW = CONV_W + E * 4;
  static constexpr int ACC_NW = NW + C;
  static constexpr int DIN = ACC_NW + C;               // (128, E+1)   one half of d[x_mlstm | z] at a time
  static constexpr int XN = DIN + kTok * (E + 1);      // (128, C+1)   normalised input
  static constexpr int TOTAL = XN + kTok * (C + 1);
};

template <int C>
__global__ void __launch_bounds__(kTok) vil_pre_bwd_b_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                              xhved_vil_params p, VilGeom g, const float* __restrict__ dconv,
                                                              const float* __restrict__ dxmv, const float* __restrict__ dz,
                                                              float* __restrict__ dx, xhved_vil_grads gr) {
  using L = PreBwdBSmem<C>;
  constexpr int E = L::E;
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  const int b = blockIdx.x / g.nc, ch = blockIdx.x % g.nc;
  stage(sm + L::W_UP, p.proj_up_weight, 2 * E * C);
  stage(sm + L::CONV_W, p.conv_weight, E * 4);
  stage(sm + L::NW, p.norm_weight, C);
  for (int i = tid; i < C; i += kTok) sm[L::ACC_NW + i] = 0.f;
  const int tau = ch * kTok + tid;
  const bool valid = tau < g.S;
  const int n = g.reverse ? g.S - 1 - tau : tau;
  float xin[C];
#pragma unroll
  for (int c = 0; c < C; ++c) xin[c] = valid ? __ldg(x + b * g.xsb + n * g.xsn + c * g.xsc) : 0.f;
  __syncthreads();
  float xn[C], rstd;
  layernorm_token<C>(xin, sm + L::NW, xn, &rstd);
#pragma unroll
  for (int c = 0; c < C; ++c) sm[L::XN + tid * (C + 1) + c] = valid ? xn[c] : 0.f;

  float dxn[C];
#pragma unroll
  for (int c = 0; c < C; ++c) dxn[c] = 0.f;
  const size_t tm_chunk = (static_cast<size_t>(b) * g.nc + ch) * E * kTok;
  float* din = sm + L::DIN + tid * (E + 1);
  // two halves of the proj_up output: o in [0,E) = d x_mlstm, o in [E,2E) = dz; the staging buffer is reused
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
#pragma unroll 1
    for (int oo = 0; oo < E; ++oo) {
      const int o = half * E + oo;
      float d;
      if (half == 0) {
        // y_{t'} uses x_t with weight w[3 - (t' - t)], t' = t..t+3  (vision_lstm.py:213-221)
        d = __ldg(dxmv + tm_chunk + static_cast<size_t>(oo) * kTok + tid);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int tp = tau + k;
          if (tp < g.S) {
            const size_t off = (static_cast<size_t>(b) * g.nc + tp / kTok) * E * kTok + static_cast<size_t>(oo) * kTok + (tp % kTok);
            d += sm[L::CONV_W + oo * 4 + 3 - k] * __ldg(dconv + off);
          }
        }
      } else {
        d = __ldg(dz + tm_chunk + static_cast<size_t>(oo) * kTok + tid);
      }
      d = valid ? d : 0.f;
      din[oo] = d;
      const float* w = sm + L::W_UP + o * C;
#pragma unroll
      for (int c = 0; c < C; c += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(w + c);
        dxn[c] += w4.x * d, dxn[c + 1] += w4.y * d, dxn[c + 2] += w4.z * d, dxn[c + 3] += w4.w * d;
      }
    }
    __syncthreads();
    // d proj_up[o][c] += sum_tok din[tok][o] * xn[tok][c]
    outer_accumulate(sm + L::DIN, E + 1, E, sm + L::XN, C + 1, C, kTok, gr.proj_up_weight + half * E * C);
    __syncthreads();
  }
  // LayerNorm backward (weight 1+w, no bias): xn = xhat*(1+w)
  float mean_g = 0.f, mean_gx = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float w1 = 1.f + sm[L::NW + c];
    const float xhat = xn[c] / w1;
    warp_acc(sm + L::ACC_NW + c, valid ? dxn[c] * xhat : 0.f);
    dxn[c] *= w1;
    xn[c] = xhat;
    mean_g += dxn[c];
    mean_gx += dxn[c] * xhat;
  }
  mean_g *= (1.f / C);
  mean_gx *= (1.f / C);
  if (valid) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float v = rstd * (dxn[c] - mean_g - xn[c] * mean_gx);
      dx[b * g.ysb + n * g.ysn + c * g.ysc] = __ldg(dy + b * g.ysb + n * g.ysn + c * g.ysc) + v;
    }
  }
  __syncthreads();
  for (int c = tid; c < C; c += kTok) atomicAdd(gr.norm_weight + c, sm[L::ACC_NW + c]);
}

template <int C>
static int launch_pre_bwd(const float* x, const float* dy, const float* dq, const float* dk, const float* dv, const float* dig,
                          const float* dfg, const float* d_act, const float* dz, const xhved_vil_params* p, const VilGeom& g, float* dx,
                          const xhved_vil_grads* gr, float* ws_dconv, float* ws_dxmv, cudaStream_t st) {
  {
    const size_t smem = PreBwdASmem<C>::TOTAL * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(vil_pre_bwd_a_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ProfScope ps(K_VIL_PRE_BWD_A, st);
    vil_pre_bwd_a_kernel<C><<<g.B * g.nc, 160, smem, st>>>(x, *p, g, dq, dk, dv, dig, dfg, d_act, ws_dconv, ws_dxmv, *gr);
  }
  {
    const size_t smem = PreBwdBSmem<C>::TOTAL * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(vil_pre_bwd_b_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ProfScope ps(K_VIL_PRE_BWD_B, st);
    vil_pre_bwd_b_kernel<C><<<g.B * g.nc, kTok, smem, st>>>(x, dy, *p, g, ws_dconv, ws_dxmv, dz, dx, *gr);
  }
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

extern "C" int xhved_vil_pre_bwd(const float* x, const float* dy, const float* dq, const float* dk, const float* dv, const float* dig,
                                 const float* dfg, const float* d_act, const float* dz, const xhved_vil_params* p, const xhved_vil_shape* sh,
                                 float* dx, const xhved_vil_grads* g, float* ws_dconv, float* ws_dxmv, void* stream) {
  VilGeom geo;
  if (int rc = vil_validate(sh, &geo)) return rc;
  if (!x || !dy || !dq || !dk || !dv || !dig || !dfg || !d_act || !dz || !p || !dx || !g || !ws_dconv || !ws_dxmv) return XHVED_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (sh->C) {
    case 16: return launch_pre_bwd<16>(x, dy, dq, dk, dv, dig, dfg, d_act, dz, p, geo, dx, g, ws_dconv, ws_dxmv, st);
    case 32: return launch_pre_bwd<32>(x, dy, dq, dk, dv, dig, dfg, d_act, dz, p, geo, dx, g, ws_dconv, ws_dxmv, st);
    case 64: return launch_pre_bwd<64>(x, dy, dq, dk, dv, dig, dfg, d_act, dz, p, geo, dx, g, ws_dconv, ws_dxmv, st);
    default: return XHVED_ERR_UNSUPPORTED_DIM;
  }
}
