// ViL blocks WIDER than the fused K2 / K3 kernels (dim 128 / 256: E = 256 / 512, head dim 64 / 128):
// the per-token glue between the block's plain GEMMs and the mLSTM cell as four fused kernels -- sm_100a.
//
// At these widths the weights of proj_up / proj_down no longer fit one CTA's shared memory, so the three Linear layers
// (vision_lstm.py:427, 443 and the LayerNorm in front, 252-259) stay plain library GEMMs on the host side (ops.vil_block_wide),
// and everything between them runs here, per token, in one pass each:
//   pre_fwd   x_mlstm -> causal conv (213-221) -> SiLU (432) -> block-diagonal q, k (of the activation) and v (of x_mlstm)
//             (158-168, 433-435) -> the cell's bf16 operand tiles; gate pre-activations Linear(3E -> 4) x 2 with bias (305-318)
//             -> padded gate rows; the activation is kept for the skip
//   post_fwd  per-head norm of h (271-287), + skip * act (437), * SiLU(z) (440) -> the row proj_down multiplies
//   post_bwd / pre_bwd  their gradients, including every parameter gradient that is a token reduction (conv, 4x4 blocks,
//             outnorm, skip); the gate-weight gradient is taken through the projections as in vil_pre.cu:
//             [dig|dfg]^T q = ([dig|dfg]^T act) Wq^T, so the kernel only hands the gate gradients back in token order
// The direction flip (419-424, 446-451) is folded into the token index map.  Thread = (token of a 128-token chunk, one of 4
// channel parts = cell heads).
#include "vil_common.cuh"

namespace xhved {

namespace wideblk {

constexpr int NP = 4;                       // channel parts per token (blockDim.y): one cell head each

struct Geom {
  int B, S, nc, Sp, reverse;
};

__device__ __forceinline__ int tok_index(const Geom& g, int tau) { return g.reverse ? g.S - 1 - tau : tau; }

__device__ __forceinline__ void ld8(const float* p, float* v) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
  v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

struct PreParams {
  const float *conv_w, *conv_b, *qw, *kw, *vw, *igw, *igb, *fgw, *fgb;
};

// shared-memory layout of the pre kernels (floats): gate weights [8][3E] | conv w [E][4] | conv b [E] | q, k, v blocks [3][E*4]
template <int E>
struct PreSm {
  static constexpr int WG = 0, CW = 8 * 3 * E, CB = CW + 4 * E, WQ = CB + E, WK = WQ + 4 * E, WV = WK + 4 * E, END = WV + 4 * E;
};

template <int E>
__device__ __forceinline__ void stage_pre_params(float* sm, const PreParams& p) {
  using L = PreSm<E>;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
  for (int i = tid; i < 4 * 3 * E; i += nt) sm[L::WG + i] = __ldg(p.igw + i), sm[L::WG + 4 * 3 * E + i] = __ldg(p.fgw + i);
  for (int i = tid; i < 4 * E; i += nt) sm[L::CW + i] = __ldg(p.conv_w + i), sm[L::WQ + i] = __ldg(p.qw + i), sm[L::WK + i] = __ldg(p.kw + i), sm[L::WV + i] = __ldg(p.vw + i);
  for (int i = tid; i < E; i += nt) sm[L::CB + i] = __ldg(p.conv_b + i);
}

// x_mlstm of tokens tau-3 .. tau (zeros in front of the sequence), channels e8 .. e8+7, from the proj_up output (B, S, 2E)
template <int E>
__device__ __forceinline__ void load_taps(const float* up, const Geom& g, int b, int tau, int e8, float (*xr)[8]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int t = tau - 3 + k;
    if (t >= 0 && t < g.S) {
      ld8(up + (static_cast<size_t>(b) * g.S + tok_index(g, t)) * (2 * E) + e8, xr[k]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) xr[k][j] = 0.f;
    }
  }
}

template <int E>
__device__ __forceinline__ void conv8(const float* sm, int e8, const float (*xr)[8], float* cv) {
  using L = PreSm<E>;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 w = *reinterpret_cast<const float4*>(sm + L::CW + (e8 + j) * 4);
    cv[j] = sm[L::CB + e8 + j] + w.x * xr[0][j] + w.y * xr[1][j] + w.z * xr[2][j] + w.w * xr[3][j];
  }
}

// y[blk*4 + o] = sum_d W[blk][o][d] x[blk*4 + d] for the two 4x4 blocks of an 8-channel group (weights (E/4, 4, 4) in smem)
__device__ __forceinline__ void blockdiag8(const float* w, int e8, const float* x, float* y) {
#pragma unroll
  for (int blk = 0; blk < 2; ++blk) {
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const float4 r = *reinterpret_cast<const float4*>(w + ((e8 >> 2) + blk) * 16 + o * 4);
      y[blk * 4 + o] = r.x * x[blk * 4] + r.y * x[blk * 4 + 1] + r.z * x[blk * 4 + 2] + r.w * x[blk * 4 + 3];
    }
  }
}
// x_grad[blk*4 + d] (+)= sum_o W[blk][o][d] y_grad[blk*4 + o]
__device__ __forceinline__ void blockdiag8_t(const float* w, int e8, const float* gy, float* gx, bool accumulate) {
#pragma unroll
  for (int blk = 0; blk < 2; ++blk) {
    float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const float4 r = *reinterpret_cast<const float4*>(w + ((e8 >> 2) + blk) * 16 + o * 4);
      const float gg = gy[blk * 4 + o];
      a[0] += r.x * gg, a[1] += r.y * gg, a[2] += r.z * gg, a[3] += r.w * gg;
    }
#pragma unroll
    for (int d = 0; d < 4; ++d) gx[blk * 4 + d] = accumulate ? gx[blk * 4 + d] + a[d] : a[d];
  }
}

// ------------------------------------------------------------------ pre, forward
template <int E>
__global__ void __launch_bounds__(kTok* NP, 1) vil_wide_pre_fwd_kernel(const float* __restrict__ up, PreParams p, Geom g,
                                                                        unsigned char* __restrict__ q_tiles,
                                                                        unsigned char* __restrict__ k_tiles,
                                                                        unsigned char* __restrict__ v_tiles, float* __restrict__ igp,
                                                                        float* __restrict__ fgp, float* __restrict__ act_out) {
  using L = PreSm<E>;
  constexpr int DH = E / 4, CH = E / NP;
  constexpr uint32_t TILE = kTok * DH * 2;
  extern __shared__ __align__(16) float sm[];
  float* red = sm + L::END;                         // [NP][kTok][8] gate partial sums
  stage_pre_params<E>(sm, p);
  __syncthreads();
  const int tok = threadIdx.x, part = threadIdx.y;
  const int b = blockIdx.x / g.nc, ch = blockIdx.x % g.nc;
  const int tau = ch * kTok + tok;
  const bool valid = tau < g.S;
  const size_t row = static_cast<size_t>(b) * g.S + tok_index(g, valid ? tau : 0);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll 1
  for (int e8 = part * CH; e8 < (part + 1) * CH; e8 += 8) {
    float xr[4][8], cv[8], a8[8], q8[8], k8[8], v8[8];
    load_taps<E>(up, g, b, valid ? tau : -8, e8, xr);
    conv8<E>(sm, e8, xr, cv);
#pragma unroll
    for (int j = 0; j < 8; ++j) a8[j] = silu(cv[j]);
    blockdiag8(sm + L::WQ, e8, a8, q8);
    blockdiag8(sm + L::WK, e8, a8, k8);
    blockdiag8(sm + L::WV, e8, xr[3], v8);
    if (valid) st8(act_out + row * E + e8, a8);
    const int head = e8 / DH, d0 = e8 % DH;
    const size_t t2 = ((static_cast<size_t>(b) * 4 + head) * g.nc + ch) * TILE + tile_off16(kTok, tok, d0 / 8);
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(q_tiles + t2) = valid ? pack8_bf16(q8) : zero;
    *reinterpret_cast<uint4*>(k_tiles + t2) = valid ? pack8_bf16(k8) : zero;
    *reinterpret_cast<uint4*>(v_tiles + t2) = valid ? pack8_bf16(v8) : zero;
#pragma unroll
    for (int gi = 0; gi < 8; ++gi) {
      const float* w = sm + L::WG + gi * 3 * E + e8;
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += w[j] * q8[j] + w[E + j] * k8[j] + w[2 * E + j] * v8[j];
      acc[gi] += s;
    }
  }
#pragma unroll
  for (int gi = 0; gi < 8; ++gi) red[(part * kTok + tok) * 8 + gi] = acc[gi];
  __syncthreads();
  if (part < 2) {
    // part 0 finishes the input gates, part 1 the forget gates (padding rows: i = -1e30, f = +1e30, see xhved.h)
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const int gi = part * 4 + h;
      float s = __ldg((part == 0 ? p.igb : p.fgb) + h);
#pragma unroll
      for (int pp = 0; pp < NP; ++pp) s += red[(pp * kTok + tok) * 8 + gi];
      const size_t o = (static_cast<size_t>(b) * 4 + h) * g.Sp + ch * kTok + tok;
      if (part == 0) igp[o] = valid ? s : -1e30f;
      else fgp[o] = valid ? s : 1e30f;
    }
  }
}

// ------------------------------------------------------------------ post: per-head statistics shared by forward and backward
// This thread's CH channels of h (bf16 tile rows) and the head's mean / rstd (two parts per head exchange through `red`)
template <int E>
struct HeadNorm {
  static constexpr int DH = E / 4, CH = E / NP, PPH = DH / CH;
  uint4 hq[CH / 8];
  float mean, rstd;
  __device__ __forceinline__ void load(const unsigned char* h_tiles, const Geom& g, int b, int ch, int tok, int part, float* red) {
    const int head = part / PPH, d0 = (part % PPH) * CH;
    const unsigned char* t = h_tiles + ((static_cast<size_t>(b) * 4 + head) * g.nc + ch) * (kTok * DH * 2);
    float s = 0.f;
#pragma unroll
    for (int cg = 0; cg < CH / 8; ++cg) {
      hq[cg] = __ldg(reinterpret_cast<const uint4*>(t + tile_off16(kTok, tok, d0 / 8 + cg)));
      float v[8];
      unpack8_bf16(hq[cg], v);
#pragma unroll
      for (int i = 0; i < 8; ++i) s += v[i];
    }
    red[part * kTok + tok] = s;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int pp = 0; pp < PPH; ++pp) tot += red[(head * PPH + pp) * kTok + tok];
    mean = tot * (1.f / DH);
    float s2 = 0.f;
#pragma unroll
    for (int cg = 0; cg < CH / 8; ++cg) {
      float v[8];
      unpack8_bf16(hq[cg], v);
#pragma unroll
      for (int i = 0; i < 8; ++i) s2 += (v[i] - mean) * (v[i] - mean);
    }
    red[(NP + part) * kTok + tok] = s2;
    __syncthreads();
    float tot2 = 0.f;
#pragma unroll
    for (int pp = 0; pp < PPH; ++pp) tot2 += red[(NP + head * PPH + pp) * kTok + tok];
    rstd = rsqrtf(tot2 * (1.f / DH) + 1e-5f);
  }
};

template <int E>
__global__ void __launch_bounds__(kTok* NP, 1) vil_wide_post_fwd_kernel(const unsigned char* __restrict__ h_tiles,
                                                                         const float* __restrict__ act, const float* __restrict__ up,
                                                                         const float* __restrict__ ow, const float* __restrict__ sk,
                                                                         Geom g, float* __restrict__ hg) {
  constexpr int CH = E / NP;
  __shared__ float red[2 * NP * kTok];
  const int tok = threadIdx.x, part = threadIdx.y;
  const int b = blockIdx.x / g.nc, ch = blockIdx.x % g.nc;
  const int tau = ch * kTok + tok;
  const bool valid = tau < g.S;
  const size_t row = static_cast<size_t>(b) * g.S + tok_index(g, valid ? tau : 0);
  HeadNorm<E> hn;
  hn.load(h_tiles, g, b, ch, tok, part, red);
  if (!valid) return;
#pragma unroll
  for (int cg = 0; cg < CH / 8; ++cg) {
    const int e8 = part * CH + cg * 8;
    float hv[8], a8[8], z8[8], o8[8];
    unpack8_bf16(hn.hq[cg], hv);
    ld8(act + row * E + e8, a8);
    ld8(up + row * (2 * E) + E + e8, z8);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xh = (hv[i] - hn.mean) * hn.rstd;
      o8[i] = (xh * (1.f + __ldg(ow + e8 + i)) + __ldg(sk + e8 + i) * a8[i]) * silu(z8[i]);
    }
    st8(hg + row * E + e8, o8);
  }
}

// ------------------------------------------------------------------ post, backward
// dhg (B,S,E) -> dh tiles (bf16), d_act (skip path, (B,S,E)), dz into d_up[..., E:], d outnorm.weight / d learnable_skip
template <int E>
__global__ void __launch_bounds__(kTok* NP, 1) vil_wide_post_bwd_kernel(const float* __restrict__ dhg, const unsigned char* __restrict__ h_tiles,
                                                                         const float* __restrict__ act, const float* __restrict__ up,
                                                                         const float* __restrict__ ow, const float* __restrict__ sk,
                                                                         Geom g, unsigned char* __restrict__ dh_tiles,
                                                                         float* __restrict__ d_act, float* __restrict__ d_up,
                                                                         float* __restrict__ g_ow, float* __restrict__ g_sk) {
  constexpr int DH = E / 4, CH = E / NP, PPH = DH / CH;
  __shared__ float red[2 * NP * kTok];
  __shared__ float accs[2 * E];                      // d learnable_skip | d outnorm.weight of this block's tokens
  const int tok = threadIdx.x, part = threadIdx.y;
  const int tid = part * kTok + tok;
  for (int i = tid; i < 2 * E; i += kTok * NP) accs[i] = 0.f;
  const int b = blockIdx.x / g.nc, ch = blockIdx.x % g.nc;
  const int tau = ch * kTok + tok;
  const bool valid = tau < g.S;
  const size_t row = static_cast<size_t>(b) * g.S + tok_index(g, valid ? tau : 0);
  HeadNorm<E> hn;
  hn.load(h_tiles, g, b, ch, tok, part, red);        // (its barriers also publish the zeroed accumulators)
  // pass 1: everything that is local to a channel, and the two sums of the norm backward
  float sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int cg = 0; cg < CH / 8; ++cg) {
    const int e8 = part * CH + cg * 8;
    float hv[8], a8[8], z8[8], dy8[8], dz8[8], da8[8], r1[8], r2[8];
    unpack8_bf16(hn.hq[cg], hv);
    if (valid) {
      ld8(act + row * E + e8, a8);
      ld8(up + row * (2 * E) + E + e8, z8);
      ld8(dhg + row * E + e8, dy8);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) a8[i] = 0.f, z8[i] = 0.f, dy8[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float w1 = 1.f + __ldg(ow + e8 + i), s = __ldg(sk + e8 + i);
      const float xh = (hv[i] - hn.mean) * hn.rstd;
      float sz, dsz;
      silu_both(z8[i], sz, dsz);
      const float hs = xh * w1 + s * a8[i];
      const float dhs = dy8[i] * sz;
      dz8[i] = dy8[i] * hs * dsz;
      da8[i] = dhs * s;
      r1[i] = dhs * a8[i];
      r2[i] = dhs * xh;
      const float gg = dhs * w1;
      sg += gg;
      sgx += gg * xh;
    }
    if (valid) {
      st8(d_up + row * (2 * E) + E + e8, dz8);
      st8(d_act + row * E + e8, da8);
    }
    warp_acc_vec<8>(accs + e8, r1);
    warp_acc_vec<8>(accs + E + e8, r2);
  }
  __syncthreads();                                    // the statistics exchange buffers are free again
  red[part * kTok + tok] = sg;
  red[(NP + part) * kTok + tok] = sgx;
  __syncthreads();
  const int head = part / PPH;
  float mg = 0.f, mgx = 0.f;
#pragma unroll
  for (int pp = 0; pp < PPH; ++pp) mg += red[(head * PPH + pp) * kTok + tok], mgx += red[(NP + head * PPH + pp) * kTok + tok];
  mg *= (1.f / DH), mgx *= (1.f / DH);
  // pass 2: dh = rstd (g - mean(g) - xhat mean(g xhat)), g = dhg silu(z) (1 + w)
  const int d0 = (part % PPH) * CH;
  unsigned char* t = dh_tiles + ((static_cast<size_t>(b) * 4 + head) * g.nc + ch) * (kTok * DH * 2);
#pragma unroll
  for (int cg = 0; cg < CH / 8; ++cg) {
    const int e8 = part * CH + cg * 8;
    float hv[8], z8[8], dy8[8], o8[8];
    unpack8_bf16(hn.hq[cg], hv);
    if (valid) {
      ld8(up + row * (2 * E) + E + e8, z8);
      ld8(dhg + row * E + e8, dy8);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) z8[i] = 0.f, dy8[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xh = (hv[i] - hn.mean) * hn.rstd;
      const float gg = dy8[i] * silu(z8[i]) * (1.f + __ldg(ow + e8 + i));
      o8[i] = valid ? hn.rstd * (gg - mg - xh * mgx) : 0.f;
    }
    *reinterpret_cast<uint4*>(t + tile_off16(kTok, tok, d0 / 8 + cg)) = pack8_bf16(o8);
  }
  __syncthreads();
  for (int i = tid; i < E; i += kTok * NP) {
    atomicAdd(g_sk + i, accs[i]);
    atomicAdd(g_ow + i, accs[E + i]);
  }
}

// ------------------------------------------------------------------ pre, backward
// dq / dk / dv tiles, gate gradients, the skip path's d_act  ->  d x_mlstm into d_up[..., :E], the gate gradients in token
// order (dg_rm (B,S,8): the host takes [dig|dfg]^T act and [dig|dfg]^T x_mlstm as two plain GEMMs), and the gradients of
// conv weight / bias and of the three block-diagonal projections (token reductions: warp transpose-sums, one set of global
// atomics per block).  The transposed causal conv needs d conv of tokens tau .. tau+3: threads tok < 3 also evaluate the first
// three tokens of the NEXT chunk, everything travels through shared memory.
template <int E>
struct PreBwdSm {
  using P = PreSm<E>;
  static constexpr int ACC = P::END;                       // accumulators: conv w [E*4] | conv b [E] | dWq | dWk | dWv [E*4 each]
  static constexpr int A_CW = ACC, A_CB = A_CW + 4 * E, A_WQ = A_CB + E, A_WK = A_WQ + 4 * E, A_WV = A_WK + 4 * E;
  static constexpr int DCS = A_WV + 4 * E;                 // d conv of this 8-channel group: [NP][kTok + 3][8]
  static constexpr int END = DCS + NP * (kTok + 3) * 8;
};

template <int E>
__global__ void __launch_bounds__(kTok* NP, 1) vil_wide_pre_bwd_kernel(const float* __restrict__ up, PreParams p, Geom g,
                                                                        const unsigned char* __restrict__ dq_tiles,
                                                                        const unsigned char* __restrict__ dk_tiles,
                                                                        const unsigned char* __restrict__ dv_tiles,
                                                                        const float* __restrict__ dig, const float* __restrict__ dfg,
                                                                        const float* __restrict__ d_act, float* __restrict__ d_up,
                                                                        float* __restrict__ dg_rm, float* __restrict__ g_cw,
                                                                        float* __restrict__ g_cb, float* __restrict__ g_qw,
                                                                        float* __restrict__ g_kw, float* __restrict__ g_vw) {
  using L = PreSm<E>;
  using M = PreBwdSm<E>;
  constexpr int DH = E / 4, CH = E / NP;
  constexpr uint32_t TILE = kTok * DH * 2;
  extern __shared__ __align__(16) float sm[];
  const int tok = threadIdx.x, part = threadIdx.y;
  const int tid = part * kTok + tok;
  stage_pre_params<E>(sm, p);
  for (int i = tid; i < M::DCS - M::ACC; i += kTok * NP) sm[M::ACC + i] = 0.f;
  __syncthreads();
  const int b = blockIdx.x / g.nc, ch = blockIdx.x % g.nc;
  float* dcs = sm + M::DCS + part * (kTok + 3) * 8;

  // one token's gate gradients [dig | dfg] (zero outside the sequence)
  auto load_dg = [&](int tau, int chunk, int r, float* dg) {
    const bool on = tau < g.S;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const size_t o = (static_cast<size_t>(b) * 4 + h) * g.Sp + chunk * kTok + r;
      dg[h] = on ? __ldg(dig + o) : 0.f;
      dg[4 + h] = on ? __ldg(dfg + o) : 0.f;
    }
  };
  // everything up to d conv for channels e8 .. e8+7 of token (chunk, r); returns the intermediates the caller may need
  auto token_grads = [&](int chunk, int r, int e8, const float* dg, float (*xr)[8], float* a8, float* gq, float* gk, float* gv,
                         float* dxv, float* dc) {
    const int tau = chunk * kTok + r;
    const bool on = tau < g.S;
    load_taps<E>(up, g, b, on ? tau : -8, e8, xr);
    float cv[8], ds[8];
    conv8<E>(sm, e8, xr, cv);
#pragma unroll
    for (int j = 0; j < 8; ++j) silu_both(cv[j], a8[j], ds[j]);
    const int head = e8 / DH, d0 = e8 % DH;
    const size_t t2 = ((static_cast<size_t>(b) * 4 + head) * g.nc + chunk) * TILE + tile_off16(kTok, r, d0 / 8);
    unpack8_bf16(__ldg(reinterpret_cast<const uint4*>(dq_tiles + t2)), gq);
    unpack8_bf16(__ldg(reinterpret_cast<const uint4*>(dk_tiles + t2)), gk);
    unpack8_bf16(__ldg(reinterpret_cast<const uint4*>(dv_tiles + t2)), gv);
    // gate path: g_qkv[j] += sum_g [dig|dfg][g] Wg[g][j]
#pragma unroll
    for (int gi = 0; gi < 8; ++gi) {
      const float* w = sm + L::WG + gi * 3 * E + e8;
      const float d = dg[gi];
#pragma unroll
      for (int j = 0; j < 8; ++j) gq[j] += d * w[j], gk[j] += d * w[E + j], gv[j] += d * w[2 * E + j];
    }
    float da[8], dsk[8];
    blockdiag8_t(sm + L::WQ, e8, gq, da, false);
    blockdiag8_t(sm + L::WK, e8, gk, da, true);
    blockdiag8_t(sm + L::WV, e8, gv, dxv, false);
    if (on) {
      ld8(d_act + (static_cast<size_t>(b) * g.S + tok_index(g, tau)) * E + e8, dsk);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) dsk[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) dc[j] = on ? (da[j] + dsk[j]) * ds[j] : 0.f;
  };

  const int tau = ch * kTok + tok;
  const bool valid = tau < g.S;
  const size_t row = static_cast<size_t>(b) * g.S + tok_index(g, valid ? tau : 0);
  float dg[8], dgh[8];
  load_dg(tau, ch, tok, dg);
  const bool halo = tok < 3 && ch + 1 < g.nc;              // this thread also evaluates token `tok` of the next chunk
  if (halo) load_dg((ch + 1) * kTok + tok, ch + 1, tok, dgh);
  if (part == 0 && valid) st8(dg_rm + row * 8, dg);

#pragma unroll 1
  for (int e8 = part * CH; e8 < (part + 1) * CH; e8 += 8) {
    float xr[4][8], a8[8], gq[8], gk[8], gv[8], dxv[8], dc[8];
    token_grads(ch, tok, e8, dg, xr, a8, gq, gk, gv, dxv, dc);
    st8(dcs + tok * 8, dc);
    if (tok < 3) {
      float hx[4][8], ha[8], hq[8], hk[8], hv[8], hdx[8], hdc[8];
      if (halo) {
        token_grads(ch + 1, tok, e8, dgh, hx, ha, hq, hk, hv, hdx, hdc);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) hdc[j] = 0.f;
      }
      st8(dcs + (kTok + tok) * 8, hdc);
    }
    __syncthreads();
    // d x_mlstm[tau] = d(v path) + sum_k w[3-k] dconv[tau + k]   (transposed causal conv, vision_lstm.py:213-221)
    if (valid) {
      float dx8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 w = *reinterpret_cast<const float4*>(sm + L::CW + (e8 + j) * 4);
        dx8[j] = dxv[j] + w.w * dcs[tok * 8 + j] + w.z * dcs[(tok + 1) * 8 + j] + w.y * dcs[(tok + 2) * 8 + j] + w.x * dcs[(tok + 3) * 8 + j];
      }
      st8(d_up + row * (2 * E) + e8, dx8);
    }
    // token reductions of this group: conv weight [8][4], conv bias [8], the 4x4 blocks of q / k / v (2 blocks x 16 each)
    {
      float prod[32];
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) prod[j * 4 + k] = dc[j] * xr[k][j];
      warp_acc_vec<32>(sm + M::A_CW + e8 * 4, prod);
      warp_acc_vec<8>(sm + M::A_CB + e8, dc);
      if (!valid) {
#pragma unroll
        for (int j = 0; j < 8; ++j) gq[j] = 0.f, gk[j] = 0.f, gv[j] = 0.f;
      }
#pragma unroll
      for (int blk = 0; blk < 2; ++blk)
#pragma unroll
        for (int o = 0; o < 4; ++o)
#pragma unroll
          for (int d = 0; d < 4; ++d) prod[blk * 16 + o * 4 + d] = gq[blk * 4 + o] * a8[blk * 4 + d];
      warp_acc_vec<32>(sm + M::A_WQ + e8 * 4, prod);
#pragma unroll
      for (int blk = 0; blk < 2; ++blk)
#pragma unroll
        for (int o = 0; o < 4; ++o)
#pragma unroll
          for (int d = 0; d < 4; ++d) prod[blk * 16 + o * 4 + d] = gk[blk * 4 + o] * a8[blk * 4 + d];
      warp_acc_vec<32>(sm + M::A_WK + e8 * 4, prod);
#pragma unroll
      for (int blk = 0; blk < 2; ++blk)
#pragma unroll
        for (int o = 0; o < 4; ++o)
#pragma unroll
          for (int d = 0; d < 4; ++d) prod[blk * 16 + o * 4 + d] = gv[blk * 4 + o] * xr[3][blk * 4 + d];
      warp_acc_vec<32>(sm + M::A_WV + e8 * 4, prod);
    }
    __syncthreads();                                  // the d conv slab is rewritten by the next group
  }
  for (int i = tid; i < 4 * E; i += kTok * NP) {
    atomicAdd(g_cw + i, sm[M::A_CW + i]);
    atomicAdd(g_qw + i, sm[M::A_WQ + i]);
    atomicAdd(g_kw + i, sm[M::A_WK + i]);
    atomicAdd(g_vw + i, sm[M::A_WV + i]);
  }
  for (int i = tid; i < E; i += kTok * NP) atomicAdd(g_cb + i, sm[M::A_CB + i]);
}

// fp32 (rows, cols) -> bf16 (rows, 3 cols): [hi | lo | hi] (b_side = 0) or [hi | hi | lo] (b_side != 0), so that ONE bf16
// tensor-core GEMM over the tripled contraction dimension gives hi*hi + lo*hi + hi*lo with fp32 accumulation (~16 mantissa bits)
__global__ void split_hilo_cat_kernel(const float* __restrict__ x, int64_t rows, int cols, int b_side, unsigned char* __restrict__ out) {
  const int64_t groups = rows * (cols / 8);
  for (int64_t gi = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; gi < groups; gi += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = gi / (cols / 8);
    const int c = static_cast<int>(gi % (cols / 8)) * 8;
    float v[8];
    ld8(x + r * cols + c, v);
    uint4 hi, lo;
    split8_hilo(v, hi, lo);
    unsigned char* o = out + (r * 3 * cols + c) * 2;
    *reinterpret_cast<uint4*>(o) = hi;
    *reinterpret_cast<uint4*>(o + cols * 2) = b_side ? hi : lo;
    *reinterpret_cast<uint4*>(o + 2 * cols * 2) = b_side ? lo : hi;
  }
}

static int make_geom(int B, int S, int reverse, Geom* g) {
  if (B <= 0 || S <= 0) return XHVED_ERR_BAD_SHAPE;
  g->B = B, g->S = S, g->nc = (S + kTok - 1) / kTok, g->Sp = g->nc * kTok, g->reverse = reverse;
  return 0;
}

template <int E>
static int launch_pre_fwd(const float* up, const PreParams& p, const Geom& g, void* q, void* k, void* v, float* ig, float* fg, float* act,
                          cudaStream_t st) {
  const size_t smem = (PreSm<E>::END + NP * kTok * 8) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(vil_wide_pre_fwd_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  ProfScope ps(K_VIL_PRE_FWD, st);
  vil_wide_pre_fwd_kernel<E><<<g.B * g.nc, dim3(kTok, NP), smem, st>>>(up, p, g, (unsigned char*)q, (unsigned char*)k, (unsigned char*)v, ig,
                                                                       fg, act);
  return (int)cudaGetLastError();
}

template <int E>
static int launch_pre_bwd(const float* up, const PreParams& p, const Geom& g, const void* dq, const void* dk, const void* dv,
                          const float* dig, const float* dfg, const float* d_act, float* d_up, float* dg_rm, float* g_cw, float* g_cb,
                          float* g_qw, float* g_kw, float* g_vw, cudaStream_t st) {
  const size_t smem = PreBwdSm<E>::END * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(vil_wide_pre_bwd_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  ProfScope ps(K_VIL_PRE_BWD_A, st);
  vil_wide_pre_bwd_kernel<E><<<g.B * g.nc, dim3(kTok, NP), smem, st>>>(up, p, g, (const unsigned char*)dq, (const unsigned char*)dk,
                                                                       (const unsigned char*)dv, dig, dfg, d_act, d_up, dg_rm, g_cw, g_cb,
                                                                       g_qw, g_kw, g_vw);
  return (int)cudaGetLastError();
}

}  // namespace wideblk
}  // namespace xhved

using namespace xhved;
using namespace xhved::wideblk;

extern "C" int xhved_vil_wide_pre_fwd(const float* up, const float* conv_w, const float* conv_b, const float* qw, const float* kw,
                                      const float* vw, const float* igw, const float* igb, const float* fgw, const float* fgb, int B, int S,
                                      int E, int reverse, void* q_tiles, void* k_tiles, void* v_tiles, float* ig_padded, float* fg_padded,
                                      float* act, void* stream) {
  Geom g;
  if (int rc = make_geom(B, S, reverse, &g)) return rc;
  if (!up || !conv_w || !conv_b || !qw || !kw || !vw || !igw || !igb || !fgw || !fgb || !q_tiles || !k_tiles || !v_tiles || !ig_padded ||
      !fg_padded || !act)
    return XHVED_ERR_BAD_ARG;
  const PreParams p{conv_w, conv_b, qw, kw, vw, igw, igb, fgw, fgb};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (E) {
    case 256: return launch_pre_fwd<256>(up, p, g, q_tiles, k_tiles, v_tiles, ig_padded, fg_padded, act, st);
    case 512: return launch_pre_fwd<512>(up, p, g, q_tiles, k_tiles, v_tiles, ig_padded, fg_padded, act, st);
    default: return XHVED_ERR_UNSUPPORTED_DIM;
  }
}

extern "C" int xhved_vil_wide_post_fwd(const void* h_tiles, const float* act, const float* up, const float* outnorm_w, const float* skip,
                                       int B, int S, int E, int reverse, float* hg, void* stream) {
  Geom g;
  if (int rc = make_geom(B, S, reverse, &g)) return rc;
  if (!h_tiles || !act || !up || !outnorm_w || !skip || !hg) return XHVED_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope ps(K_VIL_POST_FWD, st);
  switch (E) {
    case 256: vil_wide_post_fwd_kernel<256><<<g.B * g.nc, dim3(kTok, NP), 0, st>>>((const unsigned char*)h_tiles, act, up, outnorm_w, skip, g, hg); break;
    case 512: vil_wide_post_fwd_kernel<512><<<g.B * g.nc, dim3(kTok, NP), 0, st>>>((const unsigned char*)h_tiles, act, up, outnorm_w, skip, g, hg); break;
    default: return XHVED_ERR_UNSUPPORTED_DIM;
  }
  return (int)cudaGetLastError();
}

extern "C" int xhved_vil_wide_post_bwd(const float* dhg, const void* h_tiles, const float* act, const float* up, const float* outnorm_w,
                                       const float* skip, int B, int S, int E, int reverse, void* dh_tiles, float* d_act, float* d_up,
                                       float* g_outnorm_w, float* g_skip, void* stream) {
  Geom g;
  if (int rc = make_geom(B, S, reverse, &g)) return rc;
  if (!dhg || !h_tiles || !act || !up || !outnorm_w || !skip || !dh_tiles || !d_act || !d_up || !g_outnorm_w || !g_skip) return XHVED_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope ps(K_VIL_POST_BWD, st);
  switch (E) {
    case 256:
      vil_wide_post_bwd_kernel<256><<<g.B * g.nc, dim3(kTok, NP), 0, st>>>(dhg, (const unsigned char*)h_tiles, act, up, outnorm_w, skip, g,
                                                                          (unsigned char*)dh_tiles, d_act, d_up, g_outnorm_w, g_skip);
      break;
    case 512:
      vil_wide_post_bwd_kernel<512><<<g.B * g.nc, dim3(kTok, NP), 0, st>>>(dhg, (const unsigned char*)h_tiles, act, up, outnorm_w, skip, g,
                                                                          (unsigned char*)dh_tiles, d_act, d_up, g_outnorm_w, g_skip);
      break;
    default: return XHVED_ERR_UNSUPPORTED_DIM;
  }
  return (int)cudaGetLastError();
}

extern "C" int xhved_vil_wide_pre_bwd(const float* up, const float* conv_w, const float* conv_b, const float* qw, const float* kw,
                                      const float* vw, const float* igw, const float* fgw, int B, int S, int E, int reverse,
                                      const void* dq_tiles, const void* dk_tiles, const void* dv_tiles, const float* dig, const float* dfg,
                                      const float* d_act, float* d_up, float* dg_rm, float* g_conv_w, float* g_conv_b, float* g_qw,
                                      float* g_kw, float* g_vw, void* stream) {
  Geom g;
  if (int rc = make_geom(B, S, reverse, &g)) return rc;
  if (!up || !conv_w || !conv_b || !qw || !kw || !vw || !igw || !fgw || !dq_tiles || !dk_tiles || !dv_tiles || !dig || !dfg || !d_act ||
      !d_up || !dg_rm || !g_conv_w || !g_conv_b || !g_qw || !g_kw || !g_vw)
    return XHVED_ERR_BAD_ARG;
  const PreParams p{conv_w, conv_b, qw, kw, vw, igw, nullptr, fgw, nullptr};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (E) {
    case 256: return launch_pre_bwd<256>(up, p, g, dq_tiles, dk_tiles, dv_tiles, dig, dfg, d_act, d_up, dg_rm, g_conv_w, g_conv_b, g_qw, g_kw, g_vw, st);
    case 512: return launch_pre_bwd<512>(up, p, g, dq_tiles, dk_tiles, dv_tiles, dig, dfg, d_act, d_up, dg_rm, g_conv_w, g_conv_b, g_qw, g_kw, g_vw, st);
    default: return XHVED_ERR_UNSUPPORTED_DIM;
  }
}

extern "C" int xhved_split_hilo_cat(const float* x, int64_t rows, int cols, int b_side, void* out, void* stream) {
  if (!x || !out || rows <= 0 || cols <= 0 || cols % 8) return XHVED_ERR_BAD_ARG;
  const int64_t groups = rows * (cols / 8);
  const int grid = static_cast<int>(groups / 256 + 1 < 148 * 16 ? groups / 256 + 1 : 148 * 16);
  split_hilo_cat_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, rows, cols, b_side, static_cast<unsigned char*>(out));
  return (int)cudaGetLastError();
}
