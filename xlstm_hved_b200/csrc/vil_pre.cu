// K2: everything of the ViL block in front of the mLSTM cell, fused -- sm_100a.
//
// Restates, per token, vision_lstm.py:252-259 (LayerNorm, weight 1+w, eps 1e-5), :427-428 (proj_up, split),
// :213-221 (causal depthwise conv1d k=4), :432 (SiLU), :158-168 (block-diagonal q/k/v projections; v from the
// pre-conv branch, :433-435), :305-318 (input / forget gate pre-activations from [q,k,v]).  The direction flip
// (:419-424) is folded into the token index map; tokens are read straight from the strided NCDHW feature
// (UxLSTMEnc_3d.py:59).  Outputs land in the cell's tile-native bf16 layout, so the cell kernels can bulk-copy
// MMA-ready operand tiles.
//
// One tile = 128 consecutive tokens (in traversal order) = one cell chunk for all heads; 512 threads = token x head group.
#include "vil_common.cuh"

namespace xhved {

// ------------------------------------------------------------------ forward (tcgen05 version)
// proj_up runs as a 3-product bf16 hi/lo UMMA (tokens x 2E, ~fp32 accuracy), the gate pre-activations as a UMMA over the
// bf16 q|k|v tile that is afterwards bulk-stored to the cell's operand tiles; conv / SiLU / the 4x4 block-diagonal
// projections stay on CUDA cores.
template <int C>
struct PreTC {
  static constexpr int E = 2 * C, DH = E / 4, DHP = DH < 16 ? 16 : DH, NQ = 12 * DHP;
  // C <= 32: persistent CTAs (two per SM) keep the proj_up / gate weight tiles resident while they walk their token tiles.
  // C = 64: one tile per CTA, and the q|k|v tile aliases ALL proj_up operands to fit into shared memory.
  static constexpr bool PERSIST = C <= 32;
  static constexpr uint32_t X_BYTES = kTok * C * 2, W_BYTES = 2 * E * C * 2, QKV_BYTES = kTok * NQ * 2;
  static constexpr uint32_t OPER = (2 * X_BYTES + 2 * W_BYTES) > QKV_BYTES ? (2 * X_BYTES + 2 * W_BYTES) : QKV_BYTES;
  static constexpr uint32_t WHI = PERSIST ? 0 : 2 * X_BYTES, WLO = WHI + W_BYTES;
  static constexpr uint32_t QKV = PERSIST ? 2 * W_BYTES : 0;      // the token tile's hi/lo rows alias the head of the q|k|v tile
  static constexpr uint32_t XHI = QKV, XLO = QKV + X_BYTES;
  // gate weights: [8 rows hh][NQ] bf16 tile read as a 16-row operand -- rows 8..15 of every column group fall onto the
  // next group's rows 0..7 and only produce accumulator columns 8..15, which nobody reads
  static constexpr uint32_t WG_BYTES = 8 * NQ * 2;
  static constexpr uint32_t WGHI = PERSIST ? QKV + QKV_BYTES : OPER, WGLO = WGHI + WG_BYTES;
  static constexpr uint32_t XM = WGLO + WG_BYTES;                 // fp32 (131, E+1)
  static constexpr int XM_LD = E + 1;
  static constexpr uint32_t PAR = XM + ((kTok + 3) * XM_LD * 4 + 15) / 16 * 16;   // fp32 small parameters
  static constexpr int P_CW = 0, P_CB = E * 4, P_WQ = P_CB + E, P_WK = P_WQ + E * 4, P_WV = P_WK + E * 4, P_NW = P_WV + E * 4,
                       P_HX = P_NW + C, P_GB = P_HX + 3 * C, P_LN = P_GB + 8, P_N = P_LN + 2 * 4 * kTok;   // LN exchange: 2 x [4][128]
  static constexpr uint32_t TOTAL = PAR + P_N * 4;
  static constexpr uint32_t TMEM_COLS = next_pow2_tmem(2 * E + 16);
  static_assert(2 * X_BYTES <= QKV_BYTES, "token rows must fit into the q|k|v tile they alias");
};

template <int C>
__global__ void __launch_bounds__(4 * kTok, (C <= 32 ? 2 : 1)) vil_pre_fwd_kernel(const float* __restrict__ x, xhved_vil_params p, VilGeom g,
                                                                unsigned char* __restrict__ q_tiles, unsigned char* __restrict__ k_tiles,
                                                                unsigned char* __restrict__ v_tiles, float* __restrict__ igp,
                                                                float* __restrict__ fgp, unsigned char* __restrict__ act_out,
                                                                unsigned char* __restrict__ z_out, unsigned char* __restrict__ xm_out,
                                                                int ntiles) {
  // 512 threads: thread = (token, head); head group 0 additionally owns the LayerNorm and the gate read-out of its token
  using L = PreTC<C>;
  constexpr int E = L::E, DH = L::DH, DHP = L::DHP, NQ = L::NQ;
  extern __shared__ __align__(128) unsigned char smem[];
  float* par = reinterpret_cast<float*>(smem + L::PAR);
  float* xm_s = reinterpret_cast<float*>(smem + L::XM);
  __shared__ __align__(8) uint64_t bar1, bar2;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int tok = tid & (kTok - 1), head = tid >> 7;

  if (tid == 0) {
    mbar_init(&bar1, 1);
    mbar_init(&bar2, 1);
    mbar_fence_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_slot, L::TMEM_COLS);
  // ---- stage parameters (once per CTA)
  stage(par + L::P_CW, p.conv_weight, E * 4);
  stage(par + L::P_CB, p.conv_bias, E);
  stage(par + L::P_WQ, p.q_weight, E * 4);
  stage(par + L::P_WK, p.k_weight, E * 4);
  stage(par + L::P_WV, p.v_weight, E * 4);
  stage(par + L::P_NW, p.norm_weight, C);
  if (tid < 8) par[L::P_GB + tid] = tid < 4 ? __ldg(p.igate_bias + tid) : __ldg(p.fgate_bias + tid - 4);
  stage_weight_tile(p.proj_up_weight, 2 * E, C, 2 * E, smem + L::WHI, smem + L::WLO);
  // gate weights as an [8][NQ] tile in the padded q|k|v column order: column (part*4 + head)*DHP + d
  for (int gi = tid; gi < 8 * (NQ / 8); gi += blockDim.x) {
    const int hh = gi % 8, cg = gi / 8;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = cg * 8 + i, part = j / (4 * DHP), hd = (j / DHP) % 4, d = j % DHP;
      const float* W = hh < 4 ? p.igate_weight + hh * 3 * E : p.fgate_weight + (hh - 4) * 3 * E;
      v[i] = d < DH ? __ldg(W + part * E + hd * DH + d) : 0.f;
    }
    uint4 h, l;
    split8_hilo(v, h, l);
    *reinterpret_cast<uint4*>(smem + L::WGHI + tile_off16(8, hh, cg)) = h;
    *reinterpret_cast<uint4*>(smem + L::WGLO + tile_off16(8, hh, cg)) = l;
  }
  tc_fence_before();
  __syncthreads();   // parameters staged, TMEM allocated
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  constexpr int CP = (C / 4) < 8 ? 8 : (C / 4), NPART = C / CP;   // channels per thread / head groups taking part in the LayerNorm
  const bool ln_on = head < NPART;
  const int c0 = head * CP;
  float* ln1 = par + L::P_LN;
  float* ln2 = ln1 + 4 * kTok;
  auto load_x = [&](int tile, float* xin) {
    const int b = tile / g.nc, tau = (tile % g.nc) * kTok + tok;
    const int n = g.reverse ? g.S - 1 - tau : tau;
#pragma unroll
    for (int i = 0; i < CP; ++i) xin[i] = (tau < g.S && ln_on) ? __ldg(x + b * g.xsb + n * g.xsn + (c0 + i) * g.xsc) : 0.f;
  };
  float xin[CP];
  load_x(blockIdx.x, xin);

  int it = 0;
#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int b = tile / g.nc, ch = tile % g.nc;
    // ---- LayerNorm, token rows staged as bf16 hi/lo; conv halo tokens (first 3 threads of group 1)
    const int tau = ch * kTok + tok;
    const bool valid = tau < g.S;
    {
      // LayerNorm of the CTA's tokens: every head group owns CP channels of its token, statistics go through shared memory
      // (mean first, then the centred second moment, as F.layer_norm does)
      float s1 = 0.f;
#pragma unroll
      for (int i = 0; i < CP; ++i) s1 += xin[i];
      ln1[head * kTok + tok] = ln_on ? s1 : 0.f;
      __syncthreads();
      const float mean = (ln1[tok] + ln1[kTok + tok] + ln1[2 * kTok + tok] + ln1[3 * kTok + tok]) * (1.f / C);
      float s2 = 0.f;
#pragma unroll
      for (int i = 0; i < CP; ++i) s2 += (xin[i] - mean) * (xin[i] - mean);
      ln2[head * kTok + tok] = ln_on ? s2 : 0.f;
      __syncthreads();
      const float rstd = rsqrtf((ln2[tok] + ln2[kTok + tok] + ln2[2 * kTok + tok] + ln2[3 * kTok + tok]) * (1.f / C) + 1e-5f);
      if (ln_on) {
#pragma unroll
        for (int cg = 0; cg < CP / 8; ++cg) {
          float v8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            v8[i] = valid ? (xin[cg * 8 + i] - mean) * rstd * (1.f + par[L::P_NW + c0 + cg * 8 + i]) : 0.f;
          uint4 h, l;
          split8_hilo(v8, h, l);
          *reinterpret_cast<uint4*>(smem + L::XHI + tile_off16(kTok, tok, c0 / 8 + cg)) = h;
          *reinterpret_cast<uint4*>(smem + L::XLO + tile_off16(kTok, tok, c0 / 8 + cg)) = l;
        }
      }
      if (tile + static_cast<int>(gridDim.x) < ntiles) load_x(tile + gridDim.x, xin);   // next tile's tokens
    }
    if (head == 2 && tok < 96) {
      // conv halo: the 3 tokens in front of the chunk, one warp per token, lane = channel (LayerNorm through shuffles)
      const int w = tok >> 5, lane = tok & 31;
      const int htau = ch * kTok - 3 + w;
      const int hn_ = g.reverse ? g.S - 1 - htau : htau;
      float hv[(C + 31) / 32], sum = 0.f;
#pragma unroll
      for (int i = 0; i < (C + 31) / 32; ++i) {
        const int c = lane + 32 * i;
        hv[i] = (htau >= 0 && c < C) ? __ldg(x + b * g.xsb + hn_ * g.xsn + c * g.xsc) : 0.f;
        sum += hv[i];
      }
      const float mean = warp_sum_f(sum) * (1.f / C);
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < (C + 31) / 32; ++i) sq += (lane + 32 * i < C) ? (hv[i] - mean) * (hv[i] - mean) : 0.f;
      const float rstd = rsqrtf(warp_sum_f(sq) * (1.f / C) + 1e-5f);
#pragma unroll
      for (int i = 0; i < (C + 31) / 32; ++i) {
        const int c = lane + 32 * i;
        if (c < C) par[L::P_HX + w * C + c] = htau >= 0 ? (hv[i] - mean) * rstd * (1.f + par[L::P_NW + c]) : 0.f;
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
      // D[tok][o] = sum_c xn[tok][c] W_up[o][c]
      umma_gemm_hilo(tmem, smem_u32(smem + L::XHI), smem_u32(smem + L::XLO), kTok * 16, 128, smem_u32(smem + L::WHI),
                     smem_u32(smem + L::WLO), 2 * E * 16, 128, umma_idesc(128, 2 * E, false, false), C);
      umma_commit(&bar1);
    }
    // ---- conv halo (3 previous tokens): x_mlstm only, spread over the CTA while the MMA runs
    for (int idx = tid; idx < 3 * E; idx += blockDim.x) {
      const int row = idx / E, e = idx % E;
      const float* w = p.proj_up_weight + static_cast<size_t>(e) * C;
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < C; c += 4) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + c));
        const float* hxn = par + L::P_HX + row * C + c;
        acc += hxn[0] * w4.x + hxn[1] * w4.y + hxn[2] * w4.z + hxn[3] * w4.w;
      }
      xm_s[row * L::XM_LD + e] = acc;
    }
    mbar_wait(&bar1, it & 1);
    tc_fence_after();
    const size_t tt_base = static_cast<size_t>(tile) * E * (kTok * 2);   // this chunk's [128][E] bf16 token tiles (act, z, xm)
    // this head's DH channels of x_mlstm (columns head*DH..) and of z (columns E + head*DH..)
#pragma unroll
    for (int c0 = 0; c0 < DH; c0 += 8) {
      float v[8];
      tmem_ld8(tmem + lane_base + head * DH + c0, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) xm_s[(tok + 3) * L::XM_LD + head * DH + c0 + i] = v[i];
      const size_t o = tt_base + tile_off16(kTok, tok, (head * DH + c0) / 8);
      *reinterpret_cast<uint4*>(xm_out + o) = pack8_bf16(v);
      tmem_ld8(tmem + lane_base + E + head * DH + c0, v);
      *reinterpret_cast<uint4*>(z_out + o) = pack8_bf16(v);
    }
    tc_fence_before();
    __syncthreads();   // x_mlstm of all tokens visible; the token rows are dead -> the QKV tile may overwrite them
    // ---- conv + SiLU + block-diagonal q,k,v, 8 channels at a time
    const float* xm0 = xm_s + tok * L::XM_LD;   // rows tok..tok+3 <-> tokens tau-3..tau
#pragma unroll 1
    for (int e8 = head * DH; e8 < (head + 1) * DH; e8 += 8) {
      float a8[8], xm8[8], q8[8], k8[8], v8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int e = e8 + j;
        const float4 w = *reinterpret_cast<const float4*>(par + L::P_CW + e * 4);
        const float conv = par[L::P_CB + e] + w.x * xm0[e] + w.y * xm0[L::XM_LD + e] + w.z * xm0[2 * L::XM_LD + e] +
                           w.w * xm0[3 * L::XM_LD + e];
        a8[j] = silu(conv);
        xm8[j] = xm0[3 * L::XM_LD + e];
      }
      {
        float ao[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) ao[j] = valid ? a8[j] : 0.f;
        *reinterpret_cast<uint4*>(act_out + tt_base + tile_off16(kTok, tok, e8 / 8)) = pack8_bf16(ao);
      }
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) {
        const int wb = ((e8 >> 2) + blk) * 16;   // (block, out, in) 4x4
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          const float4 wq = *reinterpret_cast<const float4*>(par + L::P_WQ + wb + o * 4);
          const float4 wk = *reinterpret_cast<const float4*>(par + L::P_WK + wb + o * 4);
          const float4 wv = *reinterpret_cast<const float4*>(par + L::P_WV + wb + o * 4);
          const float* a = a8 + blk * 4;
          const float* xv = xm8 + blk * 4;
          q8[blk * 4 + o] = wq.x * a[0] + wq.y * a[1] + wq.z * a[2] + wq.w * a[3];
          k8[blk * 4 + o] = wk.x * a[0] + wk.y * a[1] + wk.z * a[2] + wk.w * a[3];
          v8[blk * 4 + o] = wv.x * xv[0] + wv.y * xv[1] + wv.z * xv[2] + wv.w * xv[3];
        }
      }
      const int d0 = e8 % DH;
      const uint4 uq = valid ? pack8_bf16(q8) : zero, uk = valid ? pack8_bf16(k8) : zero, uv = valid ? pack8_bf16(v8) : zero;
      *reinterpret_cast<uint4*>(smem + L::QKV + tile_off16(kTok, tok, ((0 * 4 + head) * DHP + d0) / 8)) = uq;
      *reinterpret_cast<uint4*>(smem + L::QKV + tile_off16(kTok, tok, ((1 * 4 + head) * DHP + d0) / 8)) = uk;
      *reinterpret_cast<uint4*>(smem + L::QKV + tile_off16(kTok, tok, ((2 * 4 + head) * DHP + d0) / 8)) = uv;
      if (DHP > DH) {   // DH = 8 padded to 16: zero the second column group of every head
#pragma unroll
        for (int part = 0; part < 3; ++part)
          *reinterpret_cast<uint4*>(smem + L::QKV + tile_off16(kTok, tok, ((part * 4 + head) * DHP + 8) / 8)) = zero;
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
      // gates[tok][hh] = sum_j qkv[tok][j] Wg[hh][j]   (bf16 q,k,v exactly as the cell sees them; hi/lo weights)
      const uint32_t aq = smem_u32(smem + L::QKV);
      umma_gemm(tmem + 2 * E, aq, kTok * 16, 128, smem_u32(smem + L::WGHI), 8 * 16, 128, umma_idesc(128, 16, false, false), NQ, false);
      umma_gemm(tmem + 2 * E, aq, kTok * 16, 128, smem_u32(smem + L::WGLO), 8 * 16, 128, umma_idesc(128, 16, false, false), NQ, true);
      umma_commit(&bar2);
      // q/k/v head tiles are contiguous column blocks of the staged tile: bulk-store them to the cell's operand tiles
      constexpr uint32_t HT = kTok * DHP * 2;
#pragma unroll 1
      for (int hd = 0; hd < 4; ++hd) {
        const size_t t2 = (static_cast<size_t>(b) * 4 + hd) * g.nc + ch;
        bulk_s2g(q_tiles + t2 * HT, smem + L::QKV + (0 * 4 + hd) * HT, HT);
        bulk_s2g(k_tiles + t2 * HT, smem + L::QKV + (1 * 4 + hd) * HT, HT);
        bulk_s2g(v_tiles + t2 * HT, smem + L::QKV + (2 * 4 + hd) * HT, HT);
      }
      bulk_commit();
    }
    mbar_wait(&bar2, it & 1);     // everybody: the gate product reads the q|k|v tile that the next iteration overwrites
    tc_fence_after();
    if (head == 0) {
      float gt[8];
      tmem_ld8(tmem + lane_base + 2 * E, gt);
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const size_t o = (static_cast<size_t>(b) * 4 + h) * g.Sp + ch * kTok + tok;
        igp[o] = valid ? gt[h] + par[L::P_GB + h] : -1e30f;
        fgp[o] = valid ? gt[4 + h] + par[L::P_GB + 4 + h] : 1e30f;
      }
    }
    if (tid == 0) bulk_wait_read();
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) tmem_dealloc(tmem, L::TMEM_COLS);
}

template <int C>
static int launch_pre_fwd(const float* x, const xhved_vil_params* p, const VilGeom& g, void* q, void* k, void* v, float* ig, float* fg,
                          void* act, void* z, void* xm, cudaStream_t st) {
  const size_t smem = PreTC<C>::TOTAL;
  cudaError_t e = cudaFuncSetAttribute(vil_pre_fwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  ProfScope ps(K_VIL_PRE_FWD, st);
  const int ntiles = g.B * g.nc;
  const int grid = PreTC<C>::PERSIST ? persistent_grid(ntiles, 2) : ntiles;
  vil_pre_fwd_kernel<C><<<grid, 4 * kTok, smem, st>>>(x, *p, g, (unsigned char*)q, (unsigned char*)k, (unsigned char*)v, ig, fg,
                                                      (unsigned char*)act, (unsigned char*)z, (unsigned char*)xm, ntiles);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------ backward, tcgen05 versions
// Kernel B (tensor-core): d[x_mlstm | z] rows are staged as bf16 hi/lo tiles; dxn = din W_up (3-product UMMA) and
// d proj_up = din^T xn (UMMA over the CTA's tokens); LayerNorm backward and the residual add in the epilogue.
template <int C>
struct PreBwdBTC {
  static constexpr int E = 2 * C;
  static constexpr int CP = (C / 4) < 8 ? 8 : (C / 4);     // channels of one token handled by one thread
  static constexpr int NPART = C / CP;                      // head groups that take part in the LayerNorm (C=16: 2, else 4)
  static constexpr uint32_t DIN_BYTES = kTok * 2 * E * 2, W_BYTES = 2 * E * C * 2, XN_BYTES = kTok * C * 2;
  static constexpr uint32_t DINHI = 0, DINLO = DIN_BYTES, WHI = 2 * DIN_BYTES, WLO = WHI + W_BYTES, XNT = WLO + W_BYTES, PAR = XNT + XN_BYTES;
  static constexpr int P_CW = 0, P_NW = E * 4, P_ANW = P_NW + C, P_LN = P_ANW + C, P_N = P_LN + 2 * 4 * kTok;   // LN exchange: 2 x [4][128]
  static constexpr uint32_t TOTAL = PAR + P_N * 4;
  static constexpr int MT = (2 * E + 127) / 128;                       // M tiles of the weight-gradient GEMM
  static constexpr uint32_t TMEM_COLS = next_pow2_tmem(C + MT * C);
};

template <int C>
__global__ void __launch_bounds__(4 * kTok, (C <= 32 ? 2 : 1)) vil_pre_bwd_b_tc_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                                        xhved_vil_params p, VilGeom g,
                                                                                        const unsigned char* __restrict__ dconv,
                                                                                        const unsigned char* __restrict__ dxmv,
                                                                                        const unsigned char* __restrict__ dz, float* __restrict__ dx,
                                                                                        xhved_vil_grads gr_base, int ntiles) {
  const xhved_vil_grads gr = replica_of(gr_base, g);
  // 512 threads: thread = (token, part).  Every part owns CP channels of its token for the LayerNorm (statistics are
  // exchanged through shared memory) and a quarter of the 2E columns of d[x_mlstm | z].
  // Persistent: the CTA walks tiles blockIdx.x, +gridDim.x, ...; parameters are staged once, and d proj_up / d norm.weight
  // accumulate over all of the CTA's tiles (the UMMA keeps adding into the same TMEM columns) before ONE flush to global.
  using L = PreBwdBTC<C>;
  constexpr int E = L::E, CP = L::CP, NPART = L::NPART;
  extern __shared__ __align__(128) unsigned char smem[];
  float* par = reinterpret_cast<float*>(smem + L::PAR);
  float* ln1 = par + L::P_LN;
  float* ln2 = ln1 + 4 * kTok;
  __shared__ __align__(8) uint64_t bar1;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int tok = tid & (kTok - 1), part = tid >> 7;
  const bool ln_on = part < NPART;
  const int c0 = part * CP;
  auto load_x = [&](int tile, float* xin) {
    const int b = tile / g.nc, tau = (tile % g.nc) * kTok + tok;
    const int n = g.reverse ? g.S - 1 - tau : tau;
#pragma unroll
    for (int i = 0; i < CP; ++i) xin[i] = (tau < g.S && ln_on) ? __ldg(x + b * g.xsb + n * g.xsn + (c0 + i) * g.xsc) : 0.f;
  };
  float xin[CP];
  load_x(blockIdx.x, xin);
  if (tid == 0) {
    mbar_init(&bar1, 1);
    mbar_fence_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_slot, L::TMEM_COLS);
  stage(par + L::P_CW, p.conv_weight, E * 4);
  stage(par + L::P_NW, p.norm_weight, C);
  for (int i = tid; i < C; i += blockDim.x) par[L::P_ANW + i] = 0.f;
  stage_weight_tile(p.proj_up_weight, 2 * E, C, 2 * E, smem + L::WHI, smem + L::WLO);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;

  int it = 0;
#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int b = tile / g.nc, ch = tile % g.nc;
    constexpr size_t TT = static_cast<size_t>(E) * kTok * 2;      // one [128][E] bf16 token tile
    const int tau = ch * kTok + tok;
    const bool valid = tau < g.S;
    const int n = g.reverse ? g.S - 1 - tau : tau;
    // ---- LayerNorm statistics, two passes through shared memory (mean, then centred second moment)
    float s1 = 0.f;
#pragma unroll
    for (int i = 0; i < CP; ++i) s1 += xin[i];
    ln1[part * kTok + tok] = ln_on ? s1 : 0.f;
    __syncthreads();
    const float mean = (ln1[tok] + ln1[kTok + tok] + ln1[2 * kTok + tok] + ln1[3 * kTok + tok]) * (1.f / C);
    float s2 = 0.f;
#pragma unroll
    for (int i = 0; i < CP; ++i) s2 += (xin[i] - mean) * (xin[i] - mean);
    ln2[part * kTok + tok] = ln_on ? s2 : 0.f;
    __syncthreads();
    const float rstd = rsqrtf((ln2[tok] + ln2[kTok + tok] + ln2[2 * kTok + tok] + ln2[3 * kTok + tok]) * (1.f / C) + 1e-5f);
    float xhat[CP];
#pragma unroll
    for (int i = 0; i < CP; ++i) xhat[i] = (xin[i] - mean) * rstd;
    if (ln_on) {
#pragma unroll
      for (int cg = 0; cg < CP / 8; ++cg) {
        float v8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v8[i] = valid ? xhat[cg * 8 + i] * (1.f + par[L::P_NW + c0 + cg * 8 + i]) : 0.f;
        *reinterpret_cast<uint4*>(smem + L::XNT + tile_off16(kTok, tok, c0 / 8 + cg)) = pack8_bf16(v8);
      }
    }
    // ---- d[x_mlstm | z] row: transposed causal conv of dconv over tokens tau..tau+3 (vision_lstm.py:213-221) + the v path
    // the four taps read dconv of tokens tau..tau+3: this tile, or the first tokens of the next tile of the same sequence
    const unsigned char* pk[4];
    bool vk[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int t = tok + k;
      vk[k] = tau + k < g.S;
      pk[k] = dconv + static_cast<size_t>(tile + (vk[k] ? (t >> 7) : 0)) * TT + (t & (kTok - 1)) * 16;
    }
    const unsigned char* px = dxmv + static_cast<size_t>(tile) * TT + tok * 16;
    const unsigned char* pz = dz + static_cast<size_t>(tile) * TT + tok * 16;
    // every pass handles 8 conv channels and 8 z channels of this token, with all of their loads issued together
#pragma unroll 1
    for (int o8 = part * 8; o8 < E; o8 += 32) {
      float dc[4][8], dzv[8], dxv[8];
      {
        const uint32_t cgo = static_cast<uint32_t>(o8 / 8) * (kTok * 16);      // column group o8/8 of the token tiles
        uint4 u[6];
#pragma unroll
        for (int k = 0; k < 4; ++k) u[k] = vk[k] ? __ldg(reinterpret_cast<const uint4*>(pk[k] + cgo)) : make_uint4(0u, 0u, 0u, 0u);
        u[4] = __ldg(reinterpret_cast<const uint4*>(px + cgo));
        u[5] = __ldg(reinterpret_cast<const uint4*>(pz + cgo));
#pragma unroll
        for (int k = 0; k < 4; ++k) unpack8_bf16(u[k], dc[k]);
        unpack8_bf16(u[4], dxv);
        unpack8_bf16(u[5], dzv);
      }
      float d8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 w = *reinterpret_cast<const float4*>(par + L::P_CW + (o8 + i) * 4);      // taps of channel o8 + i
        const float d = dxv[i] + w.w * dc[0][i] + w.z * dc[1][i] + w.y * dc[2][i] + w.x * dc[3][i];
        d8[i] = valid ? d : 0.f;
      }
      uint4 hi, lo;
      split8_hilo(d8, hi, lo);
      *reinterpret_cast<uint4*>(smem + L::DINHI + tile_off16(kTok, tok, o8 / 8)) = hi;
      *reinterpret_cast<uint4*>(smem + L::DINLO + tile_off16(kTok, tok, o8 / 8)) = lo;
#pragma unroll
      for (int i = 0; i < 8; ++i) d8[i] = valid ? dzv[i] : 0.f;
      split8_hilo(d8, hi, lo);
      *reinterpret_cast<uint4*>(smem + L::DINHI + tile_off16(kTok, tok, (E + o8) / 8)) = hi;
      *reinterpret_cast<uint4*>(smem + L::DINLO + tile_off16(kTok, tok, (E + o8) / 8)) = lo;
    }
    // loads whose latency hides behind the MMA: the residual path's upstream gradient and the next tile's tokens
    float dyin[CP];
#pragma unroll
    for (int i = 0; i < CP; ++i) dyin[i] = (valid && ln_on) ? __ldg(dy + b * g.ysb + n * g.ysn + (c0 + i) * g.ysc) : 0.f;
    if (tile + static_cast<int>(gridDim.x) < ntiles) load_x(tile + gridDim.x, xin);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
      // dxn[tok][c] = sum_o din[tok][o] W_up[o][c]          (B = MN-major view of the [2E][C] weight tile)
      umma_gemm_hilo(tmem, smem_u32(smem + L::DINHI), smem_u32(smem + L::DINLO), kTok * 16, 128, smem_u32(smem + L::WHI),
                     smem_u32(smem + L::WLO), 128, 2 * E * 16, umma_idesc(128, C, false, true), 2 * E);
      // d proj_up[o][c] += sum_tok din[tok][o] xn[tok][c]    (both operands MN-major views of token-row tiles)
#pragma unroll
      for (int mt = 0; mt < L::MT; ++mt)
        umma_gemm(tmem + C + mt * C, smem_u32(smem + L::DINHI) + mt * 16 * kTok * 16, 128, kTok * 16, smem_u32(smem + L::XNT), 128, kTok * 16,
                  umma_idesc(128, C, true, true), kTok, it > 0);
      umma_commit(&bar1);
    }
    mbar_wait(&bar1, it & 1);
    tc_fence_after();
    // ---- LayerNorm backward (weight 1+w, no bias), channels c0..c0+CP of this token
    float dxn[CP], red[CP];
    if (ln_on) {
#pragma unroll
      for (int i = 0; i < CP; i += 8) tmem_ld8(tmem + lane_base + c0 + i, dxn + i);
    }
    float pg = 0.f, pgx = 0.f;
#pragma unroll
    for (int i = 0; i < CP; ++i) {
      if (!ln_on) dxn[i] = 0.f;
      red[i] = valid ? dxn[i] * xhat[i] : 0.f;
      dxn[i] *= 1.f + par[L::P_NW + (ln_on ? c0 + i : 0)];
      pg += dxn[i];
      pgx += dxn[i] * xhat[i];
    }
    ln1[part * kTok + tok] = ln_on ? pg : 0.f;
    ln2[part * kTok + tok] = ln_on ? pgx : 0.f;
    tc_fence_before();
    __syncthreads();
    const float mean_g = (ln1[tok] + ln1[kTok + tok] + ln1[2 * kTok + tok] + ln1[3 * kTok + tok]) * (1.f / C);
    const float mean_gx = (ln2[tok] + ln2[kTok + tok] + ln2[2 * kTok + tok] + ln2[3 * kTok + tok]) * (1.f / C);
    if (valid && ln_on) {
#pragma unroll
      for (int i = 0; i < CP; ++i) {
        const float v = rstd * (dxn[i] - mean_g - xhat[i] * mean_gx);
        dx[b * g.ysb + n * g.ysn + (c0 + i) * g.ysc] = dyin[i] + v;
      }
    }
    if (ln_on) warp_acc_vec<CP>(par + L::P_ANW + c0, red);
    __syncthreads();     // LayerNorm exchange buffers are rewritten by the next tile
  }
  tc_fence_after();
  // ---- weight-gradient rows: lane = output row o (second M tile: o + 128); the parts split the C columns
#pragma unroll
  for (int mt = 0; mt < L::MT; ++mt) {
    const int o = mt * 128 + tok;
#pragma unroll 1
    for (int cc = part * 8; cc < C; cc += 32) {
      float v[8];
      tmem_ld8(tmem + lane_base + C + mt * C + cc, v);
      if (o < 2 * E) {
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(gr.proj_up_weight + static_cast<size_t>(o) * C + cc + i, v[i]);
      }
    }
  }
  for (int c = tid; c < C; c += blockDim.x) atomicAdd(gr.norm_weight + c, par[L::P_ANW + c]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, L::TMEM_COLS);
}

// Kernel A (tensor-core).  Inputs per tile: the forward's saved x_mlstm (no recompute of proj_up), d_act, the cell's
// dq/dk/dv and the gate gradients.  Gate-path gradients (dig,dfg -> q,k,v) and every parameter gradient that is a token
// reduction run as UMMAs; conv / SiLU / 4x4 block backward on CUDA cores.
// The gate-weight gradient  d Wg = [dig|dfg]^T [q|k|v]  is taken THROUGH the block-diagonal projections:
//     [dig|dfg]^T q = ([dig|dfg]^T act) Wq^T   (k alike; v with x_mlstm and Wv)
// so the kernel only reduces [dig|dfg]^T act and [dig|dfg]^T x_mlstm over tokens (operand tiles it stages anyway) and
// applies the 4x4 blocks once per CTA at flush time -- the bf16 q|k|v tiles are not read at all.
// Everything behind proj_up is separable over channels (conv and the 4x4 projections are block-diagonal), so the E inner
// channels are processed in GROUPS of at most 64: dim 64 (E = 128) makes two sweeps over the CTA's tiles with the
// shared-memory / TMEM budget of dim 32.
template <int C>
struct PreBwdATC {
  static constexpr int E = 2 * C, DH = E / 4, DHP = DH < 16 ? 16 : DH;
  static constexpr int EG = E < 64 ? E : 64;          // channels per group
  static constexpr int NG = E / EG;                   // sweeps
  static constexpr int HG = EG / DH;                  // cell heads per group (4, or 2 for dim 64)
  static constexpr int NQ = 3 * HG * DHP;             // padded q|k|v columns of a group
  static constexpr int CPT = EG / 4;                  // channels per thread (thread = token x quarter of the group)
  static constexpr uint32_t DG_BYTES = kTok * 16 * 2, WG_BYTES = 16 * NQ * 2, T_BYTES = kTok * EG * 2;
  // [dig|dfg] (double-buffered hi/lo pairs) is also read as a 128-row MN-major A operand: 32 KB window that runs on over
  // the tiles behind it; rows >= 16 of those products are never read
  static constexpr uint32_t DG0 = 0, DG_STRIDE = 2 * DG_BYTES, WGHI = 2 * DG_STRIDE, WGLO = WGHI + WG_BYTES;
  // operands of the weight-gradient GEMMs: A_1 = GQK [128][2 EG]; A_2 = [GV | DGC] (the bf16 [dig|dfg] of this tile right behind
  // the g_v columns: ONE 128-row MN-major window whose rows [0, EG) are g_v and [EG, EG + 8) the gate gradients);
  // B = [ACT | XMT] as one N = 2 EG operand
  static constexpr uint32_t GQK = WGLO + WG_BYTES;
  static constexpr uint32_t GV = GQK + kTok * 2 * EG * 2, DGC = GV + T_BYTES, ACT = DGC + DG_BYTES, XMT = ACT + T_BYTES;
  static constexpr uint32_t END2 = XMT + T_BYTES, END3 = GV + 32768, END4 = DG0 + DG_STRIDE + 32768;
  static constexpr uint32_t PAR = END2 > END3 ? (END2 > END4 ? END2 : END4) : (END3 > END4 ? END3 : END4);
  // per-group parameter slices and accumulators (restaged / flushed every sweep); the gate-bias sums once
  static constexpr int P_CW = 0, P_CB = EG * 4, P_WQ = P_CB + EG, P_WK = P_WQ + EG * 4, P_WV = P_WK + EG * 4, A_CW = P_WV + EG * 4,
                       A_CB = A_CW + EG * 4, A_GB = A_CB + EG, P_N = A_GB + 8;
  // input blocks staged by bulk async copies (EG / 8 consecutive 2 KB column groups of a [128][E] bf16 token tile):
  // x_mlstm (two stages, each followed by the EG x 4 floats of the 3 tokens in front of the chunk) and d_act (one stage)
  static constexpr uint32_t BLK = EG * kTok * 2, HX_BYTES = EG * 4 * 4, XM_STRIDE = BLK + HX_BYTES, DA_BLK = EG * kTok * 2;
  // ... and the cell's dq / dk / dv tiles of the group's heads (one stage, refilled together with d_act)
  static constexpr uint32_t TILE_B = kTok * DHP * 2, DQ_BLK = 3 * HG * TILE_B;
  static constexpr uint32_t IN_XM = (PAR + P_N * 4 + 127) / 128 * 128, IN_DA = IN_XM + 2 * XM_STRIDE, IN_DQ = IN_DA + DA_BLK;
  static constexpr uint32_t TOTAL = IN_DQ + DQ_BLK;
  // accumulators: gate path | A_1^T B: columns [0, EG) = d[q_proj|k_proj] | A_2^T B: rows [0, EG) x columns [EG, 2 EG) = d v_proj,
  // rows [EG, EG + 8) = [dig|dfg]^T act (columns [0, EG)) and [dig|dfg]^T x_mlstm (columns [EG, 2 EG)); the other blocks are not read
  static constexpr uint32_t T_GQ = 0, T_W1 = NQ, T_W2 = NQ + 2 * EG;
  static constexpr uint32_t T_DWQK = T_W1, T_DWV = T_W2 + EG, T_DWGA = T_W2, T_DWGX = T_W2 + EG;
  static constexpr int DG_LANE = EG;                  // TMEM lane of gate row 0 in the second product
  static_assert(NQ + 4 * EG <= 512 && 2 * EG <= 128 && TOTAL <= 227 * 1024, "kernel A: a channel group must fit TMEM and shared memory");
};

template <int C>
__global__ void __launch_bounds__(4 * kTok, 1) vil_pre_bwd_a_tc_kernel(xhved_vil_params p, VilGeom g, const unsigned char* __restrict__ xm,
                                                                        const unsigned char* __restrict__ dq,
                                                                        const unsigned char* __restrict__ dk,
                                                                        const unsigned char* __restrict__ dv, const float* __restrict__ dig,
                                                                        const float* __restrict__ dfg,
                                                                        const unsigned char* __restrict__ d_act,
                                                                        unsigned char* __restrict__ dconv_out,
                                                                        unsigned char* __restrict__ dxmv_out, xhved_vil_grads gr_base,
                                                                        int ntiles) {
  const xhved_vil_grads gr = replica_of(gr_base, g);
  // 512 threads: thread = (token, quarter of the channel group); the four quarters of a token share its TMEM lane.
  // Persistent: one CTA per SM walks tiles blockIdx.x, +gridDim.x, ... (once per channel group).  All parameter gradients
  // accumulate over the CTA's tiles -- the weight-gradient UMMAs keep adding into their TMEM columns, conv / bias sums live
  // in shared memory -- and are flushed to global once per sweep.  x_mlstm of the next tile streams into the other stage
  // while this tile is processed, its gate gradients are fetched one tile ahead, and the gate-path UMMA of tile i+1 is
  // issued right behind the weight-gradient UMMAs of tile i: one __syncthreads per tile.
  using L = PreBwdATC<C>;
  constexpr int E = L::E, DH = L::DH, DHP = L::DHP, NQ = L::NQ, EG = L::EG, HG = L::HG, CPT = L::CPT;
  extern __shared__ __align__(128) unsigned char smem[];
  float* par = reinterpret_cast<float*>(smem + L::PAR);
  __shared__ __align__(8) uint64_t bar_xm[2], bar_da, bar1, bar2;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int tok = tid & (kTok - 1), quarter = tid >> 7;
  int ch0 = 0;                                         // first inner channel of the current group

  auto issue_xm = [&](int tile, int s) {
    mbar_expect_tx(&bar_xm[s], L::BLK);
    bulk_g2s(smem + L::IN_XM + s * L::XM_STRIDE, xm + (static_cast<size_t>(tile) * E + ch0) * (kTok * 2), L::BLK, &bar_xm[s]);
  };
  auto issue_da = [&](int tile) {
    // channels ch0 .. ch0+EG of the [128][E] bf16 token tile are EG/8 consecutive 2 KB column groups
    mbar_expect_tx(&bar_da, L::DA_BLK + L::DQ_BLK);
    bulk_g2s(smem + L::IN_DA, d_act + (static_cast<size_t>(tile) * E + ch0) * (kTok * 2), L::DA_BLK, &bar_da);
    const int b = tile / g.nc, ch = tile % g.nc, hd0 = ch0 / DH;
#pragma unroll 1
    for (int lh = 0; lh < HG; ++lh) {
      const size_t src = ((static_cast<size_t>(b) * 4 + hd0 + lh) * g.nc + ch) * L::TILE_B;
      bulk_g2s(smem + L::IN_DQ + (0 * HG + lh) * L::TILE_B, dq + src, L::TILE_B, &bar_da);
      bulk_g2s(smem + L::IN_DQ + (1 * HG + lh) * L::TILE_B, dk + src, L::TILE_B, &bar_da);
      bulk_g2s(smem + L::IN_DQ + (2 * HG + lh) * L::TILE_B, dv + src, L::TILE_B, &bar_da);
    }
  };
  // x_mlstm of the 3 tokens in front of chunk `tile` (zeros in front of the sequence) -> behind stage s
  auto load_halo = [&](int tile, int s) {
    float* hx = reinterpret_cast<float*>(smem + L::IN_XM + s * L::XM_STRIDE + L::BLK);
    const int ch = tile % g.nc;
    for (int i = tid; i < EG * 4; i += blockDim.x) {
      const int e = i >> 2, k = i & 3;
      // element (row 125 + k, column ch0 + e) of the previous chunk's token tile
      float hv = 0.f;
      if (k < 3 && ch > 0) {
        const unsigned short* src = reinterpret_cast<const unsigned short*>(xm + static_cast<size_t>(tile - 1) * E * (kTok * 2) +
                                                                             tile_off16(kTok, kTok - 3 + k, (ch0 + e) / 8)) + ((ch0 + e) & 7);
        hv = __uint_as_float(static_cast<uint32_t>(__ldg(src)) << 16);
      }
      hx[i] = hv;
    }
  };
  // [dig | dfg] of this token (quarter 0 only)
  auto load_dg = [&](int tile, float* dg) {
    const int b = tile / g.nc, ch = tile % g.nc;
    const bool valid = ch * kTok + tok < g.S;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const size_t o = (static_cast<size_t>(b) * 4 + h) * g.Sp + ch * kTok + tok;
      dg[h] = valid ? __ldg(dig + o) : 0.f;
      dg[4 + h] = valid ? __ldg(dfg + o) : 0.f;
    }
  };
  auto stage_dg = [&](float* dg, int s, bool bias_sums) {
    unsigned char* d0 = smem + L::DG0 + s * L::DG_STRIDE;
    uint4 hi, lo;
    split8_hilo(dg, hi, lo);
    *reinterpret_cast<uint4*>(d0 + tile_off16(kTok, tok, 0)) = hi;
    *reinterpret_cast<uint4*>(d0 + L::DG_BYTES + tile_off16(kTok, tok, 0)) = lo;
    *reinterpret_cast<uint4*>(d0 + tile_off16(kTok, tok, 1)) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(d0 + L::DG_BYTES + tile_off16(kTok, tok, 1)) = make_uint4(0, 0, 0, 0);
    if (bias_sums) warp_acc_vec<8>(par + L::A_GB, dg);       // gate bias gradients (first sweep only)
  };
  // gate path: g_qkv[tok][j] = sum_hh [dig|dfg][tok][hh] Wg[hh][j]     (B = MN-major view of the [16][NQ] weight tile)
  auto issue_mma1 = [&](uint32_t tmem, int s) {
    const uint32_t d0 = smem_u32(smem + L::DG0 + s * L::DG_STRIDE);
    umma_gemm_hilo(tmem + L::T_GQ, d0, d0 + L::DG_BYTES, kTok * 16, 128, smem_u32(smem + L::WGHI), smem_u32(smem + L::WGLO), 128, 16 * 16,
                   umma_idesc(128, NQ, false, true), 16);
    umma_commit(&bar1);
  };

  if (tid == 0) {
    mbar_init(&bar_xm[0], 1);
    mbar_init(&bar_xm[1], 1);
    mbar_init(&bar_da, 1);
    mbar_init(&bar1, 1);
    mbar_init(&bar2, 1);
    mbar_fence_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid < 8) par[L::A_GB + tid] = 0.f;
  if (quarter == 0) *reinterpret_cast<uint4*>(smem + L::DGC + tile_off16(kTok, tok, 1)) = make_uint4(0, 0, 0, 0);      // stays zero
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
  float* acc = par;
  float dgn[8];

  int it = 0;                    // tiles processed so far, over all sweeps: stage / barrier phases follow its parity
#pragma unroll 1
  for (int grp = 0; grp < L::NG; ++grp) {
    ch0 = grp * EG;
    const int hd0 = ch0 / DH;    // first cell head of the group
    // ---- per-sweep prologue: first tile's inputs, the group's parameter slices and gate-weight tile
    if (tid == 0) {
      issue_xm(blockIdx.x, it & 1);
      issue_da(blockIdx.x);
    }
    if (quarter == 0) load_dg(blockIdx.x, dgn);
    load_halo(blockIdx.x, it & 1);
    stage(par + L::P_CW, p.conv_weight + ch0 * 4, EG * 4);
    stage(par + L::P_CB, p.conv_bias + ch0, EG);
    stage(par + L::P_WQ, p.q_weight + ch0 * 4, EG * 4);
    stage(par + L::P_WK, p.k_weight + ch0 * 4, EG * 4);
    stage(par + L::P_WV, p.v_weight + ch0 * 4, EG * 4);
    for (int i = tid; i < EG * 4 + EG; i += blockDim.x) par[L::A_CW + i] = 0.f;
    for (int gi = tid; gi < 16 * (NQ / 8); gi += blockDim.x) {
      const int hh = gi % 16, cg = gi / 16;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = cg * 8 + i, part = j / (HG * DHP), hd = (j / DHP) % HG, d = j % DHP;
        const float* W = hh < 4 ? p.igate_weight + hh * 3 * E : p.fgate_weight + (hh - 4) * 3 * E;
        v[i] = (hh < 8 && d < DH) ? __ldg(W + part * E + (hd0 + hd) * DH + d) : 0.f;
      }
      uint4 h, l;
      split8_hilo(v, h, l);
      *reinterpret_cast<uint4*>(smem + L::WGHI + tile_off16(16, hh, cg)) = h;
      *reinterpret_cast<uint4*>(smem + L::WGLO + tile_off16(16, hh, cg)) = l;
    }
    if (quarter == 0) stage_dg(dgn, it & 1, grp == 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) issue_mma1(tmem, it & 1);

    bool first_tile = true;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      const int ch = tile % g.nc;
      const int nxt = tile + gridDim.x;
      const bool has_next = nxt < ntiles;
      // one tile ahead: x_mlstm stage, halo and gate gradients of the next tile
      if (has_next) {
        if (tid == 0) issue_xm(nxt, s ^ 1);
        load_halo(nxt, s ^ 1);
        if (quarter == 0) load_dg(nxt, dgn);
      }
      mbar_wait(&bar_xm[s], (it >> 1) & 1);
      if (!first_tile) {      // the previous tile's weight-gradient products still read the operand tiles this loop rewrites
        mbar_wait(&bar2, (it - 1) & 1);
        tc_fence_after();
      }
      const size_t tt_base = (static_cast<size_t>(tile) * E + ch0) * (kTok * 2);      // this group's columns of the token tile
      const unsigned char* s_xm = smem + L::IN_XM + s * L::XM_STRIDE;
      const float* s_hx = reinterpret_cast<const float*>(smem + L::IN_XM + s * L::XM_STRIDE + L::BLK);
      const unsigned char* s_da = smem + L::IN_DA;
      bool first = true;
#pragma unroll
      for (int e8 = quarter * CPT; e8 < (quarter + 1) * CPT; e8 += 8) {      // channel within the group
        const int lh = e8 / DH, d0 = e8 % DH;                                 // cell head within the group, offset inside it
        float a8[8], xm8[8], cv8[8], xr[4][8];
        // x_mlstm of tokens tau-3+k; only the first warp of a quarter reaches into the halo in front of the chunk
        if ((tok & ~31) != 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k) unpack8_bf16(*reinterpret_cast<const uint4*>(s_xm + tile_off16(kTok, tok - 3 + k, e8 / 8)), xr[k]);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (tok - 3 + k >= 0) {
              unpack8_bf16(*reinterpret_cast<const uint4*>(s_xm + tile_off16(kTok, tok - 3 + k, e8 / 8)), xr[k]);
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) xr[k][j] = s_hx[(e8 + j) * 4 + tok + k];
            }
          }
        }
        float gq[8], gk[8], gv[8], dsk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int e = e8 + j;
          const float4 w = *reinterpret_cast<const float4*>(par + L::P_CW + e * 4);
          cv8[j] = par[L::P_CB + e] + w.x * xr[0][j] + w.y * xr[1][j] + w.z * xr[2][j] + w.w * xr[3][j];
          xm8[j] = xr[3][j];
        }
        // everything above is independent of the gate-path UMMA and of d_act
        if (first) {
          mbar_wait(&bar1, it & 1);
          tc_fence_after();
          mbar_wait(&bar_da, it & 1);
          first = false;
        }
        unpack8_bf16(*reinterpret_cast<const uint4*>(s_da + tile_off16(kTok, tok, e8 / 8)), dsk);
        {
          // the cell's dq / dk / dv: bf16 tiles in the layout of q / k / v, staged together with d_act
          const unsigned char* sq = smem + L::IN_DQ + lh * L::TILE_B + tile_off16(kTok, tok, d0 / 8);
          unpack8_bf16(*reinterpret_cast<const uint4*>(sq), gq);
          unpack8_bf16(*reinterpret_cast<const uint4*>(sq + HG * L::TILE_B), gk);
          unpack8_bf16(*reinterpret_cast<const uint4*>(sq + 2 * HG * L::TILE_B), gv);
        }
        // Rows beyond the end of a sequence need no masking here: every upstream gradient is exactly zero there (the cell writes
        // zero dq / dk / dv / dig / dfg for rows whose gates are the padding values, vil_post_bwd zero d_act), x_mlstm is zero and
        // the activation finite, so dconv, dxmv and the token reductions vanish by themselves (xhved.h: xhved_vil_pre_bwd).
        float t8[8];
        tmem_ld8(tmem + lane_base + L::T_GQ + (0 * HG + lh) * DHP + d0, t8);
#pragma unroll
        for (int j = 0; j < 8; ++j) gq[j] += t8[j];
        tmem_ld8(tmem + lane_base + L::T_GQ + (1 * HG + lh) * DHP + d0, t8);
#pragma unroll
        for (int j = 0; j < 8; ++j) gk[j] += t8[j];
        tmem_ld8(tmem + lane_base + L::T_GQ + (2 * HG + lh) * DHP + d0, t8);
#pragma unroll
        for (int j = 0; j < 8; ++j) gv[j] += t8[j];
        // transposed 4x4 blocks: d act = Wq^T g_q + Wk^T g_k, d x_mlstm (v path) = Wv^T g_v; weight rows as 128-bit loads
        float da8[8], dxv8[8];
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          const int wb = ((e8 >> 2) + blk) * 16;
          float sa[4] = {0.f, 0.f, 0.f, 0.f}, sv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            const float4 wq = *reinterpret_cast<const float4*>(par + L::P_WQ + wb + o * 4);
            const float4 wk = *reinterpret_cast<const float4*>(par + L::P_WK + wb + o * 4);
            const float4 wv = *reinterpret_cast<const float4*>(par + L::P_WV + wb + o * 4);
            const float a = gq[blk * 4 + o], b2 = gk[blk * 4 + o], c2 = gv[blk * 4 + o];
            sa[0] += wq.x * a + wk.x * b2, sa[1] += wq.y * a + wk.y * b2, sa[2] += wq.z * a + wk.z * b2, sa[3] += wq.w * a + wk.w * b2;
            sv[0] += wv.x * c2, sv[1] += wv.y * c2, sv[2] += wv.z * c2, sv[3] += wv.w * c2;
          }
#pragma unroll
          for (int d = 0; d < 4; ++d) da8[blk * 4 + d] = sa[d], dxv8[blk * 4 + d] = sv[d];
        }
        float dc8[8], prod[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float ds;
          silu_both(cv8[j], a8[j], ds);
          dc8[j] = (da8[j] + dsk[j]) * ds;
#pragma unroll
          for (int k = 0; k < 4; ++k) prod[j * 4 + k] = dc8[j] * xr[k][j];
        }
        *reinterpret_cast<uint4*>(dconv_out + tt_base + tile_off16(kTok, tok, e8 / 8)) = pack8_bf16(dc8);
        *reinterpret_cast<uint4*>(dxmv_out + tt_base + tile_off16(kTok, tok, e8 / 8)) = pack8_bf16(dxv8);
        warp_acc_vec<32>(acc + L::A_CW + e8 * 4, prod);     // d conv.weight[e][k], 8 channels x 4 taps
        warp_acc_vec<8>(acc + L::A_CB + e8, dc8);           // d conv.bias
        // operands of the weight-gradient GEMMs (bf16)
        *reinterpret_cast<uint4*>(smem + L::GQK + tile_off16(kTok, tok, e8 / 8)) = pack8_bf16(gq);
        *reinterpret_cast<uint4*>(smem + L::GQK + tile_off16(kTok, tok, (EG + e8) / 8)) = pack8_bf16(gk);
        *reinterpret_cast<uint4*>(smem + L::GV + tile_off16(kTok, tok, e8 / 8)) = pack8_bf16(gv);
        *reinterpret_cast<uint4*>(smem + L::ACT + tile_off16(kTok, tok, e8 / 8)) = pack8_bf16(a8);
        *reinterpret_cast<uint4*>(smem + L::XMT + tile_off16(kTok, tok, e8 / 8)) = pack8_bf16(xm8);
      }
      if (quarter == 0)       // this tile's bf16 [dig|dfg] behind the g_v columns (same thread staged it one tile ago)
        *reinterpret_cast<uint4*>(smem + L::DGC + tile_off16(kTok, tok, 0)) =
            *reinterpret_cast<const uint4*>(smem + L::DG0 + s * L::DG_STRIDE + tile_off16(kTok, tok, 0));
      if (has_next && quarter == 0) stage_dg(dgn, s ^ 1, grp == 0);      // [dig|dfg] of the next tile (its buffer was last read two tiles ago)
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      if (tid == 0) {
        const uint32_t accu = first_tile ? 0u : 1u;
        // a tcgen05.mma costs ~80 cycles whatever its shape: TWO chains of eight against B = [act | x_mlstm] (N = 2 EG) give
        // all four token reductions -- d[q_proj|k_proj] = g_{q|k}^T act (dense 2 EG x EG, only its 4x4 diagonal blocks are read
        // back), d v_proj = g_v^T x_mlstm, [dig|dfg]^T act and [dig|dfg]^T x_mlstm -- and two blocks nobody reads
        umma_gemm(tmem + L::T_W1, smem_u32(smem + L::GQK), 128, kTok * 16, smem_u32(smem + L::ACT), 128, kTok * 16,
                  umma_idesc(128, 2 * EG, true, true), kTok, accu);
        umma_gemm(tmem + L::T_W2, smem_u32(smem + L::GV), 128, kTok * 16, smem_u32(smem + L::ACT), 128, kTok * 16,
                  umma_idesc(128, 2 * EG, true, true), kTok, accu);
        umma_commit(&bar2);
        if (has_next) {
          issue_mma1(tmem, s ^ 1);      // the gate-path accumulator has been drained by everybody (sync above)
          issue_da(nxt);                // ... and so has the d_act block
        }
      }
      first_tile = false;
    }
    mbar_wait(&bar2, (it - 1) & 1);
    tc_fence_after();
    // ---- flush the parameter gradients of this channel group
    // rows hh = 0..7 of the two gate reductions live in TMEM lanes EG..EG+7: the four warps of that quadrant share the columns,
    // 16 (= four 4x4 blocks) at a time; the block-diagonal projections are applied here
    if ((warp & 3) == L::DG_LANE / 32) {
      const int hh = tok & 31;
#pragma unroll 1
      for (int c0 = quarter * 16; c0 < EG; c0 += 64) {
        float ga[16], gx[16];
        tmem_ld16(tmem + lane_base + L::T_DWGA + c0, ga);
        tmem_ld16(tmem + lane_base + L::T_DWGX + c0, gx);
        if (hh < 8) {
          float* W = hh < 4 ? gr.igate_weight + hh * 3 * E : gr.fgate_weight + (hh - 4) * 3 * E;
#pragma unroll
          for (int blk = 0; blk < 4; ++blk) {
            const int wb = ((c0 >> 2) + blk) * 16;
#pragma unroll
            for (int o = 0; o < 4; ++o) {
              float sq = 0.f, sk = 0.f, sv = 0.f;
#pragma unroll
              for (int d = 0; d < 4; ++d) {
                sq += par[L::P_WQ + wb + o * 4 + d] * ga[blk * 4 + d];
                sk += par[L::P_WK + wb + o * 4 + d] * ga[blk * 4 + d];
                sv += par[L::P_WV + wb + o * 4 + d] * gx[blk * 4 + d];
              }
              const int e_out = ch0 + c0 + blk * 4 + o;
              atomicAdd(W + 0 * E + e_out, sq);
              atomicAdd(W + 1 * E + e_out, sk);
              atomicAdd(W + 2 * E + e_out, sv);
            }
          }
        }
      }
    }
    for (int i = tid; i < EG * 4; i += blockDim.x) atomicAdd(gr.conv_weight + ch0 * 4 + i, acc[L::A_CW + i]);
    for (int i = tid; i < EG; i += blockDim.x) atomicAdd(gr.conv_bias + ch0 + i, acc[L::A_CB + i]);
    {
      // row r of the (2 EG x EG) product: r < EG -> q_proj row e_out = r, r >= EG -> k_proj; its diagonal block = 4 columns.
      // TMEM loads take a warp-uniform column address: every quarter walks a quarter of the column groups.
      const int r = tok, e_out = r % EG, blk = e_out / 4;
      const int want = (4 * blk) & ~7;
      if ((warp & 3) * 32 < 2 * EG) {
#pragma unroll 1
        for (int c0 = quarter * 8; c0 < EG; c0 += 32) {
          float v[8];
          tmem_ld8(tmem + lane_base + L::T_DWQK + c0, v);
          if (want == c0 && r < 2 * EG) {
            float* W = (r < EG ? gr.q_weight : gr.k_weight) + ch0 * 4 + blk * 16 + (e_out % 4) * 4;
#pragma unroll
            for (int d = 0; d < 4; ++d) atomicAdd(W + d, ((4 * blk) & 7) ? v[4 + d] : v[d]);
          }
        }
      }
      if ((warp & 3) * 32 < EG) {
#pragma unroll 1
        for (int c0 = quarter * 8; c0 < EG; c0 += 32) {
          float v[8];
          tmem_ld8(tmem + lane_base + L::T_DWV + c0, v);
          if (want == c0 && r < EG) {
            float* W = gr.v_weight + ch0 * 4 + blk * 16 + (e_out % 4) * 4;
#pragma unroll
            for (int d = 0; d < 4; ++d) atomicAdd(W + d, ((4 * blk) & 7) ? v[4 + d] : v[d]);
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();      // accumulators, parameter slices and stages are free for the next channel group
    tc_fence_after();
  }
  if (tid < 4) atomicAdd(gr.igate_bias + tid, acc[L::A_GB + tid]);
  else if (tid < 8) atomicAdd(gr.fgate_bias + tid - 4, acc[L::A_GB + tid]);
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int C>
static int launch_pre_bwd(const float* x, const float* dy, const void* xm, const void* q, const void* k, const void* v, const void* dq,
                          const void* dk, const void* dv, const float* dig, const float* dfg, const void* d_act, const void* dz,
                          const xhved_vil_params* p, const VilGeom& g, float* dx, const xhved_vil_grads* gr, void* ws_dconv,
                          void* ws_dxmv, cudaStream_t st) {
  auto u8 = [](const void* v_) { return static_cast<const unsigned char*>(v_); };
  {
    const size_t smem = PreBwdATC<C>::TOTAL;
    cudaError_t e = cudaFuncSetAttribute(vil_pre_bwd_a_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ProfScope ps(K_VIL_PRE_BWD_A, st);
    const int ntiles = g.B * g.nc;
    vil_pre_bwd_a_tc_kernel<C><<<persistent_grid(ntiles, 1), 4 * kTok, smem, st>>>(*p, g, u8(xm), u8(dq), u8(dk), u8(dv), dig, dfg, u8(d_act),
                                                                                  static_cast<unsigned char*>(ws_dconv),
                                                                                  static_cast<unsigned char*>(ws_dxmv), *gr, ntiles);
  }
  {
    const size_t smem = PreBwdBTC<C>::TOTAL;
    cudaError_t e = cudaFuncSetAttribute(vil_pre_bwd_b_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ProfScope ps(K_VIL_PRE_BWD_B, st);
    const int ntiles = g.B * g.nc;
    vil_pre_bwd_b_tc_kernel<C><<<persistent_grid(ntiles, C <= 32 ? 2 : 1), 4 * kTok, smem, st>>>(x, dy, *p, g, u8(ws_dconv), u8(ws_dxmv), u8(dz), dx,
                                                                                                *gr, ntiles);
  }
  return (int)cudaGetLastError();
}

}  // namespace xhved

using namespace xhved;

extern "C" int xhved_vil_pre_fwd(const float* x, const xhved_vil_params* p, const xhved_vil_shape* sh, void* q_tiles, void* k_tiles,
                                 void* v_tiles, float* ig_padded, float* fg_padded, void* act, void* z, void* xm, void* stream) {
  VilGeom g;
  if (int rc = vil_validate(sh, &g)) return rc;
  if (!x || !p || !q_tiles || !k_tiles || !v_tiles || !ig_padded || !fg_padded || !act || !z || !xm) return XHVED_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (sh->C) {
    case 16: return launch_pre_fwd<16>(x, p, g, q_tiles, k_tiles, v_tiles, ig_padded, fg_padded, act, z, xm, st);
    case 32: return launch_pre_fwd<32>(x, p, g, q_tiles, k_tiles, v_tiles, ig_padded, fg_padded, act, z, xm, st);
    case 64: return launch_pre_fwd<64>(x, p, g, q_tiles, k_tiles, v_tiles, ig_padded, fg_padded, act, z, xm, st);
    default: return XHVED_ERR_UNSUPPORTED_DIM;
  }
}

extern "C" int xhved_vil_pre_bwd(const float* x, const float* dy, const void* xm, const void* q_tiles, const void* k_tiles,
                                 const void* v_tiles, const void* dq, const void* dk, const void* dv, const float* dig, const float* dfg,
                                 const void* d_act, const void* dz, const xhved_vil_params* p, const xhved_vil_shape* sh, float* dx,
                                 const xhved_vil_grads* g, void* ws_dconv, void* ws_dxmv, void* stream) {
  VilGeom geo;
  if (int rc = vil_validate(sh, &geo)) return rc;
  if (!x || !dy || !xm || !q_tiles || !k_tiles || !v_tiles || !dq || !dk || !dv || !dig || !dfg || !d_act || !dz || !p || !dx || !g ||
      !ws_dconv || !ws_dxmv)
    return XHVED_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (sh->C) {
    case 16: return launch_pre_bwd<16>(x, dy, xm, q_tiles, k_tiles, v_tiles, dq, dk, dv, dig, dfg, d_act, dz, p, geo, dx, g, ws_dconv, ws_dxmv, st);
    case 32: return launch_pre_bwd<32>(x, dy, xm, q_tiles, k_tiles, v_tiles, dq, dk, dv, dig, dfg, d_act, dz, p, geo, dx, g, ws_dconv, ws_dxmv, st);
    case 64: return launch_pre_bwd<64>(x, dy, xm, q_tiles, k_tiles, v_tiles, dq, dk, dv, dig, dfg, d_act, dz, p, geo, dx, g, ws_dconv, ws_dxmv, st);
    default: return XHVED_ERR_UNSUPPORTED_DIM;
  }
}
