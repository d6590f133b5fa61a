// K3: everything of the ViL block behind the mLSTM cell, fused -- sm_100a.
//
// Restates, per token, vision_lstm.py:271-287 (MultiHeadLayerNorm: per-head normalisation over DH, weight
// 1+w, eps 1e-5), :437 (learnable skip of the conv activation), :440 (SiLU(z) gate), :443 (proj_down),
// :446-451 (un-flip, folded into the index map), vision_lstm_util.py:171-175 (residual add) and the token ->
// NCDHW transpose of UxLSTMEnc_3d.py:61.  One tile = 128 tokens; 512 threads = token x head.
#include "vil_common.cuh"

namespace xhved {

// ------------------------------------------------------------------ forward (tcgen05 version)
// Per token: per-head normalisation, skip, SiLU(z) gate on CUDA cores -> the gated row is staged as a bf16 hi/lo tile and
// proj_down runs as a 3-product UMMA (tokens x C); the epilogue adds the residual and writes NCDHW.
template <int C>
struct PostTC {
  static constexpr int E = 2 * C;
  static constexpr uint32_t HG_BYTES = kTok * E * 2, WD_BYTES = C * E * 2;
  static constexpr uint32_t HGHI = 0, HGLO = HG_BYTES, WDHI = 2 * HG_BYTES, WDLO = WDHI + WD_BYTES, PAR = WDLO + WD_BYTES;
  static constexpr int P_OW = 0, P_SK = E, P_N = 2 * E;
  static constexpr uint32_t TOTAL = PAR + P_N * 4;
  static constexpr uint32_t TMEM_COLS = next_pow2_tmem(C);
};

// Per-(token, head) inputs of the gated activation, with the global loads separated out so that kernels can issue
// them before their first barrier.
template <int DH>
struct HeadInputs {
  uint4 h[DH / 8];
  float a[DH], z[DH];
  // act_tile / z_tile: the [128][E] bf16 token tiles of this chunk; cg0 = first 16-byte column group of this head
  __device__ __forceinline__ void load(const unsigned char* h_tile, int r, const unsigned char* act_tile, const unsigned char* z_tile,
                                       int cg0) {
    uint4 ua[DH / 8], uz[DH / 8];
#pragma unroll
    for (int cg = 0; cg < DH / 8; ++cg) {
      h[cg] = __ldg(reinterpret_cast<const uint4*>(h_tile + tile_off16(kTok, r, cg)));
      ua[cg] = __ldg(reinterpret_cast<const uint4*>(act_tile + tile_off16(kTok, r, cg0 + cg)));
      uz[cg] = __ldg(reinterpret_cast<const uint4*>(z_tile + tile_off16(kTok, r, cg0 + cg)));
    }
#pragma unroll
    for (int cg = 0; cg < DH / 8; ++cg) unpack8_bf16(ua[cg], a + cg * 8), unpack8_bf16(uz[cg], z + cg * 8);
  }
  // the same inputs from a TMA-staged tile: [4 heads][kTok x DHP] h tiles, [128][E] act and z token tiles
  __device__ __forceinline__ void load_smem(const unsigned char* h_tile, int r, const unsigned char* act_tile,
                                            const unsigned char* z_tile, int cg0) {
#pragma unroll
    for (int cg = 0; cg < DH / 8; ++cg) {
      h[cg] = *reinterpret_cast<const uint4*>(h_tile + tile_off16(kTok, r, cg));
      unpack8_bf16(*reinterpret_cast<const uint4*>(act_tile + tile_off16(kTok, r, cg0 + cg)), a + cg * 8);
      unpack8_bf16(*reinterpret_cast<const uint4*>(z_tile + tile_off16(kTok, r, cg0 + cg)), z + cg * 8);
    }
  }
  // hg = (norm(h)*(1+ow) + sk*a) * silu(z); optionally xhat and rstd (vision_lstm.py:271-287, 437, 440)
  __device__ __forceinline__ void gated(const float* ow, const float* sk, float* hg, float* xhat, float* rstd_out) const {
    float hv[DH];
#pragma unroll
    for (int cg = 0; cg < DH / 8; ++cg) unpack8_bf16(h[cg], hv + cg * 8);
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
      const float xh = (hv[d] - mean) * rstd;
      hg[d] = (xh * (1.f + ow[d]) + sk[d] * a[d]) * silu(z[d]);
      if (xhat) xhat[d] = xh;
    }
    if (rstd_out) *rstd_out = rstd;
  }
};

template <int C>
__global__ void __launch_bounds__(4 * kTok) vil_post_fwd_kernel(const float* __restrict__ x, const unsigned char* __restrict__ h_tiles,
                                                                 const unsigned char* __restrict__ act,
                                                                 const unsigned char* __restrict__ z,
                                                                 xhved_vil_params p, VilGeom g, float* __restrict__ y) {
  // 512 threads: thread = (token, head)
  using L = PostTC<C>;
  constexpr int E = L::E, DH = E / 4, DHP = DH < 16 ? 16 : DH;
  extern __shared__ __align__(128) unsigned char smem[];
  float* par = reinterpret_cast<float*>(smem + L::PAR);
  __shared__ __align__(8) uint64_t bar1;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int tok = tid & (kTok - 1), head = tid >> 7;
  const int b = blockIdx.x / g.nc, ch = blockIdx.x % g.nc;
  const int tau = ch * kTok + tok;
  const bool valid = tau < g.S;
  const int n = g.reverse ? g.S - 1 - tau : tau;
  const size_t tt_base = (static_cast<size_t>(b) * g.nc + ch) * E * (kTok * 2);      // this chunk's bf16 token tiles
  // all per-token global loads first: their latency overlaps the parameter staging below
  HeadInputs<DH> in;
  in.load(h_tiles + ((static_cast<size_t>(b) * 4 + head) * g.nc + ch) * (kTok * DHP * 2), tok, act + tt_base, z + tt_base,
          head * (DH / 8));
  float xres[8 * ((C + 31) / 32)];
#pragma unroll
  for (int it = 0; it < (C + 31) / 32; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = head * 8 + it * 32 + i;
      xres[it * 8 + i] = (valid && c < C) ? __ldg(x + b * g.xsb + n * g.xsn + c * g.xsc) : 0.f;
    }
  if (tid == 0) {
    mbar_init(&bar1, 1);
    mbar_fence_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_slot, L::TMEM_COLS);
  stage(par + L::P_OW, p.outnorm_weight, E);
  stage(par + L::P_SK, p.learnable_skip, E);
  stage_weight_tile(p.proj_down_weight, C, E, C, smem + L::WDHI, smem + L::WDLO);
  __syncthreads();
  {
    float hg[DH];
    in.gated(par + L::P_OW + head * DH, par + L::P_SK + head * DH, hg, nullptr, nullptr);
#pragma unroll
    for (int cg = 0; cg < DH / 8; ++cg) {
      float v8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v8[i] = valid ? hg[cg * 8 + i] : 0.f;
      uint4 hi, lo;
      split8_hilo(v8, hi, lo);
      *reinterpret_cast<uint4*>(smem + L::HGHI + tile_off16(kTok, tok, head * (DH / 8) + cg)) = hi;
      *reinterpret_cast<uint4*>(smem + L::HGLO + tile_off16(kTok, tok, head * (DH / 8) + cg)) = lo;
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    // out[tok][c] = sum_e hg[tok][e] W_down[c][e]
    umma_gemm_hilo(tmem, smem_u32(smem + L::HGHI), smem_u32(smem + L::HGLO), kTok * 16, 128, smem_u32(smem + L::WDHI),
                   smem_u32(smem + L::WDLO), C * 16, 128, umma_idesc(128, C, false, false), E);
    umma_commit(&bar1);
  }
  mbar_wait(&bar1, 0);
  tc_fence_after();
  const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
  // the four head groups split the C output channels of their token, 8 at a time
#pragma unroll
  for (int it = 0; it < (C + 31) / 32; ++it) {
    const int c0 = head * 8 + it * 32;
    if (c0 < C) {
      float o[8];
      tmem_ld8(tmem + lane_base + c0, o);
      if (valid) {
#pragma unroll
        for (int i = 0; i < 8; ++i) y[b * g.ysb + n * g.ysn + (c0 + i) * g.ysc] = xres[it * 8 + i] + o[i];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, L::TMEM_COLS);
}

template <int C>
static int launch_post_fwd(const float* x, const void* h, const void* act, const void* z, const xhved_vil_params* p, const VilGeom& g,
                           float* y, cudaStream_t st) {
  const size_t smem = PostTC<C>::TOTAL;
  cudaError_t e = cudaFuncSetAttribute(vil_post_fwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  ProfScope ps(K_VIL_POST_FWD, st);
  vil_post_fwd_kernel<C><<<g.B * g.nc, 4 * kTok, smem, st>>>(x, (const unsigned char*)h, (const unsigned char*)act, (const unsigned char*)z, *p,
                                                             g, y);
  return (int)cudaGetLastError();
}


// Persistent variant (C <= 32): one CTA per SM walks the token tiles; a tile's h / act / z blocks are contiguous in HBM, so
// the bulk-copy engine streams tile i+1 into the second smem stage while tile i is gated, multiplied and written back.
// Parameters, the proj_down tile and the TMEM allocation are set up once per CTA.
template <int C>
struct PostPersist {
  static constexpr int E = 2 * C, DH = E / 4, DHP = DH < 16 ? 16 : DH;
  static constexpr uint32_t ACT_BYTES = E * kTok * 2, H1_BYTES = kTok * DHP * 2, STAGE_BYTES = 2 * ACT_BYTES + 4 * H1_BYTES;
  static constexpr uint32_t S_ACT = 0, S_Z = ACT_BYTES, S_H = 2 * ACT_BYTES;
  static constexpr uint32_t HG_BYTES = kTok * E * 2, WD_BYTES = C * E * 2;
  static constexpr uint32_t HGHI = 2 * STAGE_BYTES, HGLO = HGHI + HG_BYTES, WDHI = HGLO + HG_BYTES, WDLO = WDHI + WD_BYTES,
                            PAR = WDLO + WD_BYTES;
  static constexpr int P_OW = 0, P_SK = E, P_N = 2 * E;
  static constexpr uint32_t TOTAL = PAR + P_N * 4;
  static constexpr uint32_t ACC_COLS = next_pow2_tmem(C), TMEM_COLS = 2 * ACC_COLS;
};

template <int C>
__global__ void __launch_bounds__(4 * kTok, 1) vil_post_fwd_persist_kernel(const float* __restrict__ x, const unsigned char* __restrict__ h_tiles,
                                                                            const unsigned char* __restrict__ act,
                                                                            const unsigned char* __restrict__ z,
                                                                            xhved_vil_params p, VilGeom g, float* __restrict__ y, int ntiles) {
  using L = PostPersist<C>;
  constexpr int E = L::E, DH = L::DH, NX = 8 * ((C + 31) / 32);
  extern __shared__ __align__(128) unsigned char smem[];
  float* par = reinterpret_cast<float*>(smem + L::PAR);
  __shared__ __align__(8) uint64_t bar_full[2], bar_mma;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int tok = tid & (kTok - 1), head = tid >> 7;

  auto issue = [&](int tile, int s) {
    unsigned char* st = smem + s * L::STAGE_BYTES;
    mbar_expect_tx(&bar_full[s], L::STAGE_BYTES);
    bulk_g2s(st + L::S_ACT, act + static_cast<size_t>(tile) * L::ACT_BYTES, L::ACT_BYTES, &bar_full[s]);
    bulk_g2s(st + L::S_Z, z + static_cast<size_t>(tile) * L::ACT_BYTES, L::ACT_BYTES, &bar_full[s]);
    const int b = tile / g.nc, ch = tile % g.nc;
#pragma unroll
    for (int hd = 0; hd < 4; ++hd)
      bulk_g2s(st + L::S_H + hd * L::H1_BYTES, h_tiles + ((static_cast<size_t>(b) * 4 + hd) * g.nc + ch) * L::H1_BYTES, L::H1_BYTES,
               &bar_full[s]);
  };
  auto load_x = [&](int tile, float* xr) {
    const int b = tile / g.nc, ch = tile % g.nc;
    const int tau = ch * kTok + tok;
    const int n = g.reverse ? g.S - 1 - tau : tau;
#pragma unroll
    for (int it = 0; it < (C + 31) / 32; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = head * 8 + it * 32 + i;
        xr[it * 8 + i] = (tau < g.S && c < C) ? __ldg(x + b * g.xsb + n * g.xsn + c * g.xsc) : 0.f;
      }
  };

  if (tid == 0) {
    mbar_init(&bar_full[0], 1);
    mbar_init(&bar_full[1], 1);
    mbar_init(&bar_mma, 1);
    mbar_fence_init();
    if (static_cast<int>(blockIdx.x) < ntiles) issue(blockIdx.x, 0);
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_slot, L::TMEM_COLS);
  float xres[NX];
  if (static_cast<int>(blockIdx.x) < ntiles) load_x(blockIdx.x, xres);
  stage(par + L::P_OW, p.outnorm_weight, E);
  stage(par + L::P_SK, p.learnable_skip, E);
  stage_weight_tile(p.proj_down_weight, C, E, C, smem + L::WDHI, smem + L::WDLO);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;

  // Software pipeline over the CTA's tiles: the UMMA of tile j runs while tile j-1's accumulator is drained and tile j+1's
  // inputs are gated; one __syncthreads per tile.  Accumulators alternate between two TMEM column ranges.
  auto epilogue = [&](int tile, uint32_t acc, const float* xr) {
    const int b = tile / g.nc, ch = tile % g.nc;
    const int tau = ch * kTok + tok;
    const int n = g.reverse ? g.S - 1 - tau : tau;
#pragma unroll
    for (int k = 0; k < (C + 31) / 32; ++k) {
      const int c0 = head * 8 + k * 32;
      if (c0 < C) {
        float o[8];
        tmem_ld8(acc + lane_base + c0, o);
        if (tau < g.S) {
#pragma unroll
          for (int i = 0; i < 8; ++i) y[b * g.ysb + n * g.ysn + (c0 + i) * g.ysc] = xr[k * 8 + i] + o[i];
        }
      }
    }
  };
  int it = 0, prev_tile = -1;
  float xprev[NX];
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int s = it & 1;
    const int nxt = tile + gridDim.x;
    if (tid == 0 && nxt < ntiles) issue(nxt, s ^ 1);
    float xnext[NX];
    if (nxt < ntiles) load_x(nxt, xnext);
    const bool valid = (tile % g.nc) * kTok + tok < g.S;
    mbar_wait(&bar_full[s], (it >> 1) & 1);
    {
      const unsigned char* st = smem + s * L::STAGE_BYTES;
      HeadInputs<DH> in;
      in.load_smem(st + L::S_H + head * L::H1_BYTES, tok, st + L::S_ACT, st + L::S_Z, head * (DH / 8));
      float hg[DH];
      in.gated(par + L::P_OW + head * DH, par + L::P_SK + head * DH, hg, nullptr, nullptr);
      if (it > 0) {       // the previous product has read the gated tile and filled its accumulator
        mbar_wait(&bar_mma, (it - 1) & 1);
        tc_fence_after();
      }
#pragma unroll
      for (int cg = 0; cg < DH / 8; ++cg) {
        float v8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v8[i] = valid ? hg[cg * 8 + i] : 0.f;
        uint4 hi, lo;
        split8_hilo(v8, hi, lo);
        *reinterpret_cast<uint4*>(smem + L::HGHI + tile_off16(kTok, tok, head * (DH / 8) + cg)) = hi;
        *reinterpret_cast<uint4*>(smem + L::HGLO + tile_off16(kTok, tok, head * (DH / 8) + cg)) = lo;
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
      umma_gemm_hilo(tmem + s * L::ACC_COLS, smem_u32(smem + L::HGHI), smem_u32(smem + L::HGLO), kTok * 16, 128, smem_u32(smem + L::WDHI),
                     smem_u32(smem + L::WDLO), C * 16, 128, umma_idesc(128, C, false, false), E);
      umma_commit(&bar_mma);
    }
    if (it > 0) epilogue(prev_tile, tmem + (s ^ 1) * L::ACC_COLS, xprev);
#pragma unroll
    for (int i = 0; i < NX; ++i) xprev[i] = xres[i], xres[i] = xnext[i];
    prev_tile = tile;
  }
  if (it > 0) {
    mbar_wait(&bar_mma, (it - 1) & 1);
    tc_fence_after();
    epilogue(prev_tile, tmem + ((it - 1) & 1) * L::ACC_COLS, xprev);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, L::TMEM_COLS);
}

template <int C>
static int launch_post_fwd_persist(const float* x, const void* h, const void* act, const void* z, const xhved_vil_params* p,
                                   const VilGeom& g, float* y, cudaStream_t st) {
  const size_t smem = PostPersist<C>::TOTAL;
  cudaError_t e = cudaFuncSetAttribute(vil_post_fwd_persist_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int ntiles = g.B * g.nc;
  const int grid = persistent_grid(ntiles, 1);
  ProfScope ps(K_VIL_POST_FWD, st);
  vil_post_fwd_persist_kernel<C><<<grid, 4 * kTok, smem, st>>>(x, (const unsigned char*)h, (const unsigned char*)act, (const unsigned char*)z,
                                                               *p, g, y, ntiles);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------ backward (tcgen05 version)
// d(gated) = dy W_down (3-product UMMA), elementwise backward of gate / skip / per-head norm on CUDA cores,
// d proj_down = gated^T dy over the CTA's tokens as a UMMA (bf16 operands), dh written as bf16 tiles by bulk store.
template <int C>
struct PostBwdTC {
  static constexpr int E = 2 * C, DH = E / 4, DHP = DH < 16 ? 16 : DH;
  static constexpr uint32_t DY_BYTES = kTok * C * 2, WD_BYTES = C * E * 2, DHT_BYTES = kTok * 4 * DHP * 2;
  static constexpr uint32_t HG = 0;                       // [128][E] bf16, read as a 128-row MN-major A operand: 32 KB window
  static constexpr uint32_t DYHI = 32768, DYLO = DYHI + DY_BYTES, WDHI = DYLO + DY_BYTES, WDLO = WDHI + WD_BYTES;
  static constexpr uint32_t DHT = WDLO + WD_BYTES, PAR = DHT + DHT_BYTES;
  static constexpr int P_OW = 0, P_SK = E, P_ASK = 2 * E, P_AOW = 3 * E, P_N = 4 * E;
  static constexpr uint32_t TOTAL = PAR + P_N * 4;
  static constexpr uint32_t TMEM_COLS = next_pow2_tmem(E + C);
};

template <int C>
__global__ void __launch_bounds__(4 * kTok) vil_post_bwd_kernel(const float* __restrict__ dy, const unsigned char* __restrict__ h_tiles,
                                                                 const unsigned char* __restrict__ act,
                                                                 const unsigned char* __restrict__ z,
                                                                 xhved_vil_params p, VilGeom g, unsigned char* __restrict__ dh_tiles,
                                                                 unsigned char* __restrict__ d_act, unsigned char* __restrict__ dz,
                                                                 xhved_vil_grads gr_base) {
  const xhved_vil_grads gr = replica_of(gr_base, g);
  // 512 threads: thread = (token, head)
  using L = PostBwdTC<C>;
  constexpr int E = L::E, DH = L::DH, DHP = L::DHP;
  extern __shared__ __align__(128) unsigned char smem[];
  float* par = reinterpret_cast<float*>(smem + L::PAR);
  __shared__ __align__(8) uint64_t bar1, bar2;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int tok = tid & (kTok - 1), head = tid >> 7;
  const int b = blockIdx.x / g.nc, ch = blockIdx.x % g.nc;
  const int tau = ch * kTok + tok;
  const bool valid = tau < g.S;
  const int n = g.reverse ? g.S - 1 - tau : tau;
  const size_t tt_base = (static_cast<size_t>(b) * g.nc + ch) * E * (kTok * 2);      // this tile's bf16 token tiles (act, z, dz, d_act)
  const size_t tile = (static_cast<size_t>(b) * 4 + head) * g.nc + ch;
  // all per-token global loads first: their latency overlaps the parameter staging below
  HeadInputs<DH> in;
  in.load(h_tiles + tile * (kTok * DHP * 2), tok, act + tt_base, z + tt_base, head * (DH / 8));
  if (tid == 0) {
    mbar_init(&bar1, 1);
    mbar_init(&bar2, 1);
    mbar_fence_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_slot, L::TMEM_COLS);
  stage(par + L::P_OW, p.outnorm_weight, E);
  stage(par + L::P_SK, p.learnable_skip, E);
  for (int i = tid; i < 2 * E; i += blockDim.x) par[L::P_ASK + i] = 0.f;
  stage_weight_tile(p.proj_down_weight, C, E, C, smem + L::WDHI, smem + L::WDLO);
  for (int cg = head; cg < C / 8; cg += 4) {
    float v8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v8[i] = valid ? __ldg(dy + b * g.ysb + n * g.ysn + (cg * 8 + i) * g.ysc) : 0.f;
    uint4 hi, lo;
    split8_hilo(v8, hi, lo);
    *reinterpret_cast<uint4*>(smem + L::DYHI + tile_off16(kTok, tok, cg)) = hi;
    *reinterpret_cast<uint4*>(smem + L::DYLO + tile_off16(kTok, tok, cg)) = lo;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    // dhg[tok][e] = sum_c dy[tok][c] W_down[c][e]     (B = MN-major view of the [C][E] weight tile)
    umma_gemm_hilo(tmem, smem_u32(smem + L::DYHI), smem_u32(smem + L::DYLO), kTok * 16, 128, smem_u32(smem + L::WDHI),
                   smem_u32(smem + L::WDLO), 128, C * 16, umma_idesc(128, E, false, true), C);
    umma_commit(&bar1);
  }
  // recompute the gated activation of this (token, head) while the MMA runs
  float hg[DH], xhat[DH], rstd;
  const float* ow = par + L::P_OW + head * DH;
  const float* sk = par + L::P_SK + head * DH;
  in.gated(ow, sk, hg, xhat, &rstd);
  mbar_wait(&bar1, 0);
  tc_fence_after();
  const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
  float dhg[DH];
  if (DH >= 16) {
#pragma unroll
    for (int c0 = 0; c0 < DH; c0 += 16) tmem_ld16(tmem + lane_base + head * DH + c0, dhg + c0);
  } else {
    tmem_ld8(tmem + lane_base + head * DH, dhg);
  }
  float gg[DH], r1[DH], r2[DH], mean_g = 0.f, mean_gx = 0.f;
  // dz and the skip-path d_act leave as bf16 token tiles [128][E] (their consumers stage them as bf16 operands anyway)
#pragma unroll
  for (int cg = 0; cg < DH / 8; ++cg) {
    float dz8[8], da8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int d = cg * 8 + i;
      const float a = in.a[d], zz = in.z[d];
      float sz, dsz;
      silu_both(zz, sz, dsz);
      const float hs = xhat[d] * (1.f + ow[d]) + sk[d] * a;
      const float dhs = valid ? dhg[d] * sz : 0.f;
      dz8[i] = valid ? dhg[d] * hs * dsz : 0.f;
      da8[i] = dhs * sk[d];
      r1[d] = dhs * a;
      r2[d] = dhs * xhat[d];
      gg[d] = dhs * (1.f + ow[d]);
      mean_g += gg[d];
      mean_gx += gg[d] * xhat[d];
    }
    const size_t o = tt_base + tile_off16(kTok, tok, head * (DH / 8) + cg);
    *reinterpret_cast<uint4*>(dz + o) = pack8_bf16(dz8);
    *reinterpret_cast<uint4*>(d_act + o) = pack8_bf16(da8);
  }
  mean_g *= (1.f / DH);
  mean_gx *= (1.f / DH);
  const uint4 zero = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int cg = 0; cg < DHP / 8; ++cg) {
    uint4 u = zero;
    if (cg * 8 < DH) {
      float o8[8], hg8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int d = cg * 8 + i;
        o8[i] = rstd * (gg[d] - mean_g - xhat[d] * mean_gx);
        hg8[i] = valid ? hg[d] : 0.f;
      }
      u = pack8_bf16(o8);
      *reinterpret_cast<uint4*>(smem + L::HG + tile_off16(kTok, tok, head * (DH / 8) + cg)) = pack8_bf16(hg8);
    }
    *reinterpret_cast<uint4*>(smem + L::DHT + tile_off16(kTok, tok, head * (DHP / 8) + cg)) = u;
  }
  // d learnable_skip / d outnorm.weight of this head's channels
#pragma unroll
  for (int d0 = 0; d0 < DH; d0 += (DH < 32 ? DH : 32)) {
    warp_acc_vec<(DH < 32 ? DH : 32)>(par + L::P_ASK + head * DH + d0, r1 + d0);
    warp_acc_vec<(DH < 32 ? DH : 32)>(par + L::P_AOW + head * DH + d0, r2 + d0);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    // d proj_down^T[e][c] = sum_tok hg[tok][e] dy[tok][c]   (both operands MN-major views of token-row tiles)
    umma_gemm(tmem + E, smem_u32(smem + L::HG), 128, kTok * 16, smem_u32(smem + L::DYHI), 128, kTok * 16, umma_idesc(128, C, true, true),
              kTok, false);
    umma_commit(&bar2);
    constexpr uint32_t HT = kTok * DHP * 2;
#pragma unroll 1
    for (int hd = 0; hd < 4; ++hd) {
      const size_t t2 = (static_cast<size_t>(b) * 4 + hd) * g.nc + ch;
      bulk_s2g(dh_tiles + t2 * HT, smem + L::DHT + hd * HT, HT);
    }
    bulk_commit();
  }
  for (int e = tid; e < E; e += blockDim.x) {
    atomicAdd(gr.learnable_skip + e, par[L::P_ASK + e]);
    atomicAdd(gr.outnorm_weight + e, par[L::P_AOW + e]);
  }
  mbar_wait(&bar2, 0);
  tc_fence_after();
  if ((warp & 3) * 32 < E) {
#pragma unroll 1
    for (int c0 = head * 16; c0 < C; c0 += 64) {
      float v[16];
      tmem_ld16(tmem + lane_base + E + c0, v);
      if (tok < E) {
#pragma unroll
        for (int i = 0; i < 16; ++i) atomicAdd(gr.proj_down_weight + static_cast<size_t>(c0 + i) * E + tok, v[i]);
      }
    }
  }
  if (tid == 0) bulk_wait_read();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, L::TMEM_COLS);
}

// Persistent variant (C <= 32): one CTA per SM walks the token tiles.  h / act / z tiles of tile i+1 stream into the second
// smem stage (bulk copies) while tile i is processed; d proj_down accumulates in TMEM and d learnable_skip / d outnorm.weight
// in shared memory over all of the CTA's tiles, and are flushed to global once.
template <int C>
struct PostBwdPersist {
  static constexpr int E = 2 * C, DH = E / 4, DHP = DH < 16 ? 16 : DH;
  static constexpr uint32_t ACT_BYTES = E * kTok * 2, H1_BYTES = kTok * DHP * 2, STAGE_BYTES = 2 * ACT_BYTES + 4 * H1_BYTES;
  static constexpr uint32_t S_ACT = 0, S_Z = ACT_BYTES, S_H = 2 * ACT_BYTES;
  static constexpr uint32_t HG_BYTES = kTok * E * 2, DY_BYTES = kTok * C * 2, WD_BYTES = C * E * 2, DHT_BYTES = 4 * H1_BYTES;
  // HG is read as a 128-row MN-major A operand through a 32 KB window that runs on over the tiles behind it
  static constexpr uint32_t HG = 2 * STAGE_BYTES, DYHI = HG + HG_BYTES, DYLO = DYHI + DY_BYTES, WDHI = DYLO + DY_BYTES,
                            WDLO = WDHI + WD_BYTES, DHT = WDLO + WD_BYTES, PAR = DHT + DHT_BYTES;
  static constexpr int P_OW = 0, P_SK = E, P_ASK = 2 * E, P_AOW = 3 * E, P_N = 4 * E;
  static constexpr uint32_t END = PAR + P_N * 4;
  static constexpr uint32_t TOTAL = END > HG + 32768 ? END : HG + 32768;
  static constexpr uint32_t TMEM_COLS = next_pow2_tmem(E + C);
};

template <int C>
__global__ void __launch_bounds__(4 * kTok, 1) vil_post_bwd_persist_kernel(const float* __restrict__ dy, const unsigned char* __restrict__ h_tiles,
                                                                            const unsigned char* __restrict__ act,
                                                                            const unsigned char* __restrict__ z,
                                                                            xhved_vil_params p, VilGeom g, unsigned char* __restrict__ dh_tiles,
                                                                            unsigned char* __restrict__ d_act,
                                                                            unsigned char* __restrict__ dz, xhved_vil_grads gr_base,
                                                                            int ntiles) {
  const xhved_vil_grads gr = replica_of(gr_base, g);
  using L = PostBwdPersist<C>;
  constexpr int E = L::E, DH = L::DH, DHP = L::DHP;
  extern __shared__ __align__(128) unsigned char smem[];
  float* par = reinterpret_cast<float*>(smem + L::PAR);
  __shared__ __align__(8) uint64_t bar_full[2], bar1, bar2;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int tok = tid & (kTok - 1), head = tid >> 7;

  auto issue = [&](int tile, int s) {
    unsigned char* st = smem + s * L::STAGE_BYTES;
    mbar_expect_tx(&bar_full[s], L::STAGE_BYTES);
    bulk_g2s(st + L::S_ACT, act + static_cast<size_t>(tile) * L::ACT_BYTES, L::ACT_BYTES, &bar_full[s]);
    bulk_g2s(st + L::S_Z, z + static_cast<size_t>(tile) * L::ACT_BYTES, L::ACT_BYTES, &bar_full[s]);
    const int b = tile / g.nc, ch = tile % g.nc;
#pragma unroll
    for (int hd = 0; hd < 4; ++hd)
      bulk_g2s(st + L::S_H + hd * L::H1_BYTES, h_tiles + ((static_cast<size_t>(b) * 4 + hd) * g.nc + ch) * L::H1_BYTES, L::H1_BYTES,
               &bar_full[s]);
  };
  // this thread's 8 channels of dy (column group `head` of the token; C <= 32 => at most one group per thread)
  auto load_dy = [&](int tile, float* v8) {
    const int b = tile / g.nc, tau = (tile % g.nc) * kTok + tok;
    const int n = g.reverse ? g.S - 1 - tau : tau;
#pragma unroll
    for (int i = 0; i < 8; ++i) v8[i] = (tau < g.S && head * 8 < C) ? __ldg(dy + b * g.ysb + n * g.ysn + (head * 8 + i) * g.ysc) : 0.f;
  };

  if (tid == 0) {
    mbar_init(&bar_full[0], 1);
    mbar_init(&bar_full[1], 1);
    mbar_init(&bar1, 1);
    mbar_init(&bar2, 1);
    mbar_fence_init();
    issue(blockIdx.x, 0);
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_slot, L::TMEM_COLS);
  float dy8[8];
  load_dy(blockIdx.x, dy8);
  stage(par + L::P_OW, p.outnorm_weight, E);
  stage(par + L::P_SK, p.learnable_skip, E);
  for (int i = tid; i < 2 * E; i += blockDim.x) par[L::P_ASK + i] = 0.f;
  stage_weight_tile(p.proj_down_weight, C, E, C, smem + L::WDHI, smem + L::WDLO);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const float* ow = par + L::P_OW + head * DH;
  const float* sk = par + L::P_SK + head * DH;

  int it = 0;
#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int s = it & 1;
    const int nxt = tile + gridDim.x;
    if (tid == 0 && nxt < ntiles) issue(nxt, s ^ 1);
    const int b = tile / g.nc, ch = tile % g.nc;
    const bool valid = ch * kTok + tok < g.S;
    const size_t tt_base = static_cast<size_t>(tile) * E * (kTok * 2);      // this tile's bf16 token tiles (dz, d_act)
    if (it > 0) {     // the previous tile's weight-gradient product still reads the dy and gated tiles
      mbar_wait(&bar2, (it - 1) & 1);
      tc_fence_after();
    }
    if (head * 8 < C) {
      uint4 hi, lo;
      split8_hilo(dy8, hi, lo);
      *reinterpret_cast<uint4*>(smem + L::DYHI + tile_off16(kTok, tok, head)) = hi;
      *reinterpret_cast<uint4*>(smem + L::DYLO + tile_off16(kTok, tok, head)) = lo;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
      // dhg[tok][e] = sum_c dy[tok][c] W_down[c][e]     (B = MN-major view of the [C][E] weight tile)
      umma_gemm_hilo(tmem, smem_u32(smem + L::DYHI), smem_u32(smem + L::DYLO), kTok * 16, 128, smem_u32(smem + L::WDHI),
                     smem_u32(smem + L::WDLO), 128, C * 16, umma_idesc(128, E, false, true), C);
      umma_commit(&bar1);
    }
    if (nxt < ntiles) load_dy(nxt, dy8);
    // recompute the gated activation of this (token, head) while the MMA runs
    mbar_wait(&bar_full[s], (it >> 1) & 1);
    const unsigned char* st = smem + s * L::STAGE_BYTES;
    HeadInputs<DH> in;
    in.load_smem(st + L::S_H + head * L::H1_BYTES, tok, st + L::S_ACT, st + L::S_Z, head * (DH / 8));
    float hg[DH], xhat[DH], rstd;
    in.gated(ow, sk, hg, xhat, &rstd);
    mbar_wait(&bar1, it & 1);
    tc_fence_after();
    float dhg[DH];
    if (DH >= 16) {
#pragma unroll
      for (int c0 = 0; c0 < DH; c0 += 16) tmem_ld16(tmem + lane_base + head * DH + c0, dhg + c0);
    } else {
      tmem_ld8(tmem + lane_base + head * DH, dhg);
    }
    float gg[DH], r1[DH], r2[DH], mean_g = 0.f, mean_gx = 0.f;
    // dz and the skip-path d_act leave as bf16 token tiles [128][E] (their consumers stage them as bf16 operands anyway)
#pragma unroll
    for (int cg = 0; cg < DH / 8; ++cg) {
      float dz8[8], da8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int d = cg * 8 + i;
        const float a = in.a[d], zz = in.z[d];
        float sz, dsz;
        silu_both(zz, sz, dsz);
        const float hs = xhat[d] * (1.f + ow[d]) + sk[d] * a;
        const float dhs = valid ? dhg[d] * sz : 0.f;
        dz8[i] = valid ? dhg[d] * hs * dsz : 0.f;
        da8[i] = dhs * sk[d];
        r1[d] = dhs * a;
        r2[d] = dhs * xhat[d];
        gg[d] = dhs * (1.f + ow[d]);
        mean_g += gg[d];
        mean_gx += gg[d] * xhat[d];
      }
      const size_t o = tt_base + tile_off16(kTok, tok, head * (DH / 8) + cg);
      *reinterpret_cast<uint4*>(dz + o) = pack8_bf16(dz8);
      *reinterpret_cast<uint4*>(d_act + o) = pack8_bf16(da8);
    }
    mean_g *= (1.f / DH);
    mean_gx *= (1.f / DH);
    const uint4 zero = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int cg = 0; cg < DHP / 8; ++cg) {
      uint4 u = zero;
      if (cg * 8 < DH) {
        float o8[8], hg8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int d = cg * 8 + i;
          o8[i] = rstd * (gg[d] - mean_g - xhat[d] * mean_gx);
          hg8[i] = valid ? hg[d] : 0.f;
        }
        u = pack8_bf16(o8);
        *reinterpret_cast<uint4*>(smem + L::HG + tile_off16(kTok, tok, head * (DH / 8) + cg)) = pack8_bf16(hg8);
      }
      *reinterpret_cast<uint4*>(smem + L::DHT + tile_off16(kTok, tok, head * (DHP / 8) + cg)) = u;
    }
    // d learnable_skip / d outnorm.weight of this head's channels (accumulated over the CTA's tiles)
#pragma unroll
    for (int d0 = 0; d0 < DH; d0 += (DH < 32 ? DH : 32)) {
      warp_acc_vec<(DH < 32 ? DH : 32)>(par + L::P_ASK + head * DH + d0, r1 + d0);
      warp_acc_vec<(DH < 32 ? DH : 32)>(par + L::P_AOW + head * DH + d0, r2 + d0);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
      // d proj_down^T[e][c] += sum_tok hg[tok][e] dy[tok][c]   (both operands MN-major views of token-row tiles)
      umma_gemm(tmem + E, smem_u32(smem + L::HG), 128, kTok * 16, smem_u32(smem + L::DYHI), 128, kTok * 16, umma_idesc(128, C, true, true),
                kTok, it > 0);
      umma_commit(&bar2);
#pragma unroll 1
      for (int hd = 0; hd < 4; ++hd) {
        const size_t t2 = (static_cast<size_t>(b) * 4 + hd) * g.nc + ch;
        bulk_s2g(dh_tiles + t2 * L::H1_BYTES, smem + L::DHT + hd * L::H1_BYTES, L::H1_BYTES);
      }
      bulk_commit();
      bulk_wait_read();
    }
    __syncthreads();      // stage s and the dh tile are free again
  }
  if (it > 0) {
    mbar_wait(&bar2, (it - 1) & 1);
    tc_fence_after();
  }
  for (int e = tid; e < E; e += blockDim.x) {
    atomicAdd(gr.learnable_skip + e, par[L::P_ASK + e]);
    atomicAdd(gr.outnorm_weight + e, par[L::P_AOW + e]);
  }
  if ((warp & 3) * 32 < E) {
#pragma unroll 1
    for (int c0 = head * 16; c0 < C; c0 += 64) {
      float v[16];
      tmem_ld16(tmem + lane_base + E + c0, v);
      if (tok < E) {
#pragma unroll
        for (int i = 0; i < 16; ++i) atomicAdd(gr.proj_down_weight + static_cast<size_t>(c0 + i) * E + tok, v[i]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, L::TMEM_COLS);
}

template <int C>
static int launch_post_bwd_persist(const float* dy, const void* h, const void* act, const void* z, const xhved_vil_params* p,
                                   const VilGeom& g, void* dh, void* d_act, void* dz, const xhved_vil_grads* gr, cudaStream_t st) {
  const size_t smem = PostBwdPersist<C>::TOTAL;
  cudaError_t e = cudaFuncSetAttribute(vil_post_bwd_persist_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int ntiles = g.B * g.nc;
  ProfScope ps(K_VIL_POST_BWD, st);
  vil_post_bwd_persist_kernel<C><<<persistent_grid(ntiles, 1), 4 * kTok, smem, st>>>(dy, (const unsigned char*)h, (const unsigned char*)act,
                                                                                    (const unsigned char*)z, *p, g,
                                                                                    (unsigned char*)dh, (unsigned char*)d_act, (unsigned char*)dz, *gr, ntiles);
  return (int)cudaGetLastError();
}

template <int C>
static int launch_post_bwd(const float* dy, const void* h, const void* act, const void* z, const xhved_vil_params* p, const VilGeom& g,
                           void* dh, void* d_act, void* dz, const xhved_vil_grads* gr, cudaStream_t st) {
  const size_t smem = PostBwdTC<C>::TOTAL;
  cudaError_t e = cudaFuncSetAttribute(vil_post_bwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  ProfScope ps(K_VIL_POST_BWD, st);
  vil_post_bwd_kernel<C><<<g.B * g.nc, 4 * kTok, smem, st>>>(dy, (const unsigned char*)h, (const unsigned char*)act, (const unsigned char*)z, *p, g,
                                                             (unsigned char*)dh,
                                                             (unsigned char*)d_act, (unsigned char*)dz, *gr);
  return (int)cudaGetLastError();
}

}  // namespace xhved

using namespace xhved;

extern "C" int xhved_vil_post_fwd(const float* x, const void* h_tiles, const void* act, const void* z, const xhved_vil_params* p,
                                  const xhved_vil_shape* sh, float* y, void* stream) {
  VilGeom g;
  if (int rc = vil_validate(sh, &g)) return rc;
  if (!x || !h_tiles || !act || !z || !p || !y) return XHVED_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (sh->C) {
    case 16: return launch_post_fwd_persist<16>(x, h_tiles, act, z, p, g, y, st);
    case 32: return launch_post_fwd_persist<32>(x, h_tiles, act, z, p, g, y, st);
    case 64: return launch_post_fwd<64>(x, h_tiles, act, z, p, g, y, st);
    default: return XHVED_ERR_UNSUPPORTED_DIM;
  }
}

extern "C" int xhved_vil_post_bwd(const float* dy, const void* h_tiles, const void* act, const void* z, const xhved_vil_params* p,
                                  const xhved_vil_shape* sh, void* dh_tiles, void* d_act, void* dz, const xhved_vil_grads* g,
                                  void* stream) {
  VilGeom geo;
  if (int rc = vil_validate(sh, &geo)) return rc;
  if (!dy || !h_tiles || !act || !z || !p || !dh_tiles || !d_act || !dz || !g) return XHVED_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (sh->C) {
    case 16: return launch_post_bwd_persist<16>(dy, h_tiles, act, z, p, geo, dh_tiles, d_act, dz, g, st);
    case 32: return launch_post_bwd_persist<32>(dy, h_tiles, act, z, p, geo, dh_tiles, d_act, dz, g, st);
    case 64: return launch_post_bwd<64>(dy, h_tiles, act, z, p, geo, dh_tiles, d_act, dz, g, st);
    default: return XHVED_ERR_UNSUPPORTED_DIM;
  }
}

namespace xhved {
__global__ void reduce_replicas_kernel(const float* __restrict__ src, int R, int64_t stride, int64_t n, float* __restrict__ dst) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  float acc = 0.f;
  for (int r = 0; r < R; ++r) acc += src[r * stride + i];
  dst[i] = acc;
}
}  // namespace xhved

extern "C" int xhved_reduce_replicas(const float* src, int replicas, int64_t stride, int64_t n, float* dst, void* stream) {
  if (!src || !dst || replicas < 1 || n <= 0 || stride < n) return XHVED_ERR_BAD_ARG;
  xhved::reduce_replicas_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, replicas, stride, n, dst);
  return (int)cudaGetLastError();
}
