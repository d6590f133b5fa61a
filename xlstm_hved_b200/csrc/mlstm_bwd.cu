// Chunkwise stabilised mLSTM cell, backward -- sm_100a (tcgen05 + TMEM + bulk async copies).
//
// Gradient of the reference's parallel_stabilized_simple (vision_lstm.py:48-130) in chunkwise form
// (formulas: SURVEY.md 8a-note, restated on the CPU in oracle/restate.py::mlstm_backward):
//
//   r_t = 1/N_t,  G_t = [dh_t r_t | db_t | 0]          (row-extended output gradient; V is extended by a ones column)
//   dP = G Vext^T,  D'' = scale*D',  P = S o D'',  dS = dP o D''
//   dQ = dS K + w (G [C|n]^T)            dK = dS^T Q + fac (Vext R^T)          dV = P^T G + fac (K R)
//   di_s = k_s.dK_s,  dc_u = q_u.dQ_u - k_u.dK_u,  dlogsigmoid(f) = reverse cumsum(dc),  df = dlf * sigmoid(-f)
//
//   phase B1  chunk_rstate  per chunk:  dR_c = sum_t exp(b_t - m_t - lambda_c) (q_t/sqrt(DH)) G_t^T
//   phase B2  state_scan (reverse)      R entering every chunk from the right, with its log-scale mu
//   phase B3  chunk_grad    per chunk:  the eight MMAs above, gate-gradient dot products
//   phase B4  gate_finish   per (b,head): reverse cumsum over the whole sequence
//
// The gradient through the row-max stabiliser m_t is dropped (relative effect ~1e-6, see tests).
#include <stdlib.h>

#include "mlstm_common.cuh"
#include "prof.cuh"
#include "xhved.h"

namespace xhved {

int launch_state_scan(int dhp, const float* dstate, const float* g, const float* amax, int BH, int nc, int reverse, void* states,
                      float* m_prev, cudaStream_t st);
int launch_chunk_grad_ws(int dhp, const void* q, const void* k, const void* v, const void* h, const void* dh_t, const float* ig,
                         const float* fg, const float* m, const float* den, const void* states, const float* m_prev, const void* rstates,
                         const float* mu_next, int BH, int nc, float scale, float eps, void* dq, void* dk, void* dv, float* dig,
                         float* dc, float* dc_tot, cudaStream_t st);

int launch_chunk_rstate_ws(int dhp, const void* q, const void* dh_t, const void* h, const float* fg, const float* m, const float* den,
                           int BH, int nc, float scale, float eps, float* dstate, float* g_out, float* lam_out, cudaStream_t st);
bool state_ws_enabled(int dhp);

int launch_chunk_grad_wide(const void* q, const void* k, const void* v, const void* h, const void* dh_t, const float* ig, const float* fg,
                           const float* m, const float* den, const void* states, const float* m_prev, const void* rstates,
                           const float* mu_next, int BH, int nc, float scale, float eps, void* dq, void* dk, void* dv, float* dig,
                           float* dc, float* dc_tot, cudaStream_t st);

// Which chunk_grad kernel runs: the persistent kernel with balanced roles and P^T in tensor memory (mlstm_bwd_ws.cu) where it
// is the faster of the two on B200 (measured, profiles/r02_cell_scaling.jsonl: dhp <= 32, two CTAs per SM), the
// one-tile-per-CTA kernel below for dhp = 64.  XHVED_GRAD_WS=0 / 1 forces one of them for A/B measurements.  Read once.
static bool grad_ws_enabled(int dhp) {
  static const int mode = [] {
    const char* e = getenv("XHVED_GRAD_WS");
    return !e ? -1 : (e[0] == '0' ? 0 : 1);
  }();
  return mode < 0 ? dhp <= 32 : mode != 0;
}

// Build the row-extended gradient G_t in place over the dH tile (sG: [128][NE] tile-native) using the H tile.
template <int DHP>
__device__ __forceinline__ void build_G_row(unsigned char* sG, const unsigned char* sH, int t, float m, float den, float eps) {
  const float flo = __expf(-m);
  const float nrm = fmaxf(fabsf(den), flo) + eps;
  const float r = 1.f / nrm;
  float dhh = 0.f;
#pragma unroll
  for (int cg = 0; cg < DHP / 8; ++cg) {
    uint4* pg = reinterpret_cast<uint4*>(sG + tile_off16(kL, t, cg));
    const uint4 ug = *pg;
    const uint4 uh = *reinterpret_cast<const uint4*>(sH + tile_off16(kL, t, cg));
    const float2 g0 = unpack_bf16x2(ug.x), g1 = unpack_bf16x2(ug.y), g2 = unpack_bf16x2(ug.z), g3 = unpack_bf16x2(ug.w);
    const float2 h0 = unpack_bf16x2(uh.x), h1 = unpack_bf16x2(uh.y), h2 = unpack_bf16x2(uh.z), h3 = unpack_bf16x2(uh.w);
    dhh += g0.x * h0.x + g0.y * h0.y + g1.x * h1.x + g1.y * h1.y + g2.x * h2.x + g2.y * h2.y + g3.x * h3.x + g3.y * h3.y;
    uint4 o;
    o.x = pack_bf16x2(g0.x * r, g0.y * r);
    o.y = pack_bf16x2(g1.x * r, g1.y * r);
    o.z = pack_bf16x2(g2.x * r, g2.y * r);
    o.w = pack_bf16x2(g3.x * r, g3.y * r);
    *pg = o;
  }
  const float dn = -dhh * r;
  const float db = (fabsf(den) > flo) ? (den >= 0.f ? dn : -dn) : 0.f;
  *reinterpret_cast<uint4*>(sG + tile_off16(kL, t, DHP / 8)) = make_uint4(pack_bf16x2(db, 0.f), 0u, 0u, 0u);
  *reinterpret_cast<uint4*>(sG + tile_off16(kL, t, DHP / 8 + 1)) = make_uint4(0u, 0u, 0u, 0u);
}

// ------------------------------------------------------------------ phase B1
template <int DHP>
__global__ void __launch_bounds__(kThreads) mlstm_chunk_rstate_kernel(const unsigned char* __restrict__ q_tiles,
                                                                       const unsigned char* __restrict__ dh_tiles,
                                                                       const unsigned char* __restrict__ h_tiles,
                                                                       const float* __restrict__ fg, const float* __restrict__ m_in,
                                                                       const float* __restrict__ den_in, float scale, float eps,
                                                                       float* __restrict__ dstate, float* __restrict__ g_out,
                                                                       float* __restrict__ lam_out) {
  constexpr int NE = ext_cols(DHP);
  constexpr uint32_t TILE = kL * DHP * 2;
  constexpr uint32_t TMEM_COLS = next_pow2_cols(NE);
  extern __shared__ __align__(128) unsigned char smem[];
  // Q~ hi / lo are read as 128-row MN-major A operands, i.e. through a 32 KB window each; rows >= DHP of the product are
  // never read, so the windows simply run on over the lo tile, G, H (and each other) instead of owning 32 KB of padding
  unsigned char* sQ = smem;
  unsigned char* sQlo = smem + TILE;
  unsigned char* sG = smem + 2 * TILE;          // [128][NE]
  unsigned char* sH = sG + kL * NE * 2;         // [128][DHP]
  __shared__ __align__(8) uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_slot;
  __shared__ float red[8];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int tile = blockIdx.x;
  const size_t grow = static_cast<size_t>(tile) * kL + tid;
  if (tid == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    mbar_fence_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_slot, TMEM_COLS);
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar_load, 3 * TILE);
    bulk_g2s(sQ, q_tiles + static_cast<size_t>(tile) * TILE, TILE, &bar_load);
    bulk_g2s(sG, dh_tiles + static_cast<size_t>(tile) * TILE, TILE, &bar_load);
    bulk_g2s(sH, h_tiles + static_cast<size_t>(tile) * TILE, TILE, &bar_load);
  }
  const float lf = log_sigmoid(fg[grow]);
  const float m = m_in[grow], den = den_in[grow];
  float g;
  const float b = block_cumsum128(lf, red, &g);
  float lam;
  block_cummax128(b - m, red, &lam);
  const float wgt = __expf(b - m - lam) * scale;
  mbar_wait(&bar_load, 0);
  build_G_row<DHP>(sG, sH, tid, m, den, eps);
  scale_row_hilo<DHP>(sQ, sQlo, tid, wgt);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  // dR[d][e'] = sum_t Q~[t][d] * G[t][e'].  As in chunk_state the lo tile sits inside the hi operand's 128-row window: one
  // pass yields hi^T G in rows [0, DHP) and lo^T G in rows [DHP, 2 DHP); the epilogue adds them (DHP <= 64).
  constexpr bool ONE_PASS = DHP <= 64;
  if (tid == 0) {
    umma_gemm(tmem, smem_u32(sQ), 128, kL * 16, smem_u32(sG), 128, kL * 16, umma_idesc(128, NE, true, true), kL, false);
    if (!ONE_PASS) umma_gemm(tmem, smem_u32(sQlo), 128, kL * 16, smem_u32(sG), 128, kL * 16, umma_idesc(128, NE, true, true), kL, true);
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  store_state_rows<DHP, ONE_PASS>(tmem, dstate + static_cast<size_t>(tile) * DHP * NE, reinterpret_cast<float*>(smem));
  if (tid == 0) {
    g_out[tile] = g;
    lam_out[tile] = lam;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

// ------------------------------------------------------------------ phase B3
// 16 columns of one row: D'' = exp2(u_row + v_col); P = S o D'' and dS = dP o D'' written as bf16 (two 16-byte groups each)
template <bool MASK>
__device__ __forceinline__ void decay_pair(const float* sv, const float* dp, float urow, const float* vcol, int s0, int row,
                                           unsigned char* dstP, unsigned char* dstS) {
#pragma unroll
  for (int j8 = 0; j8 < 2; ++j8) {
    float p[8], ds[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float d = fast_exp2(urow + vcol[j8 * 8 + j]);
      if (MASK) d = (s0 + j8 * 8 + j <= row) ? d : 0.f;
      p[j] = sv[j8 * 8 + j] * d;
      ds[j] = dp[j8 * 8 + j] * d;
    }
    *reinterpret_cast<uint4*>(dstP + j8 * (kL * 16)) = pack8_bf16(p);
    *reinterpret_cast<uint4*>(dstS + j8 * (kL * 16)) = pack8_bf16(ds);
  }
}

template <int DHP>
struct BwdSmem {
  static constexpr int NE = ext_cols(DHP);
  static constexpr uint32_t TILE = kL * DHP * 2;
  static constexpr uint32_t EXT = kL * NE * 2;
  static constexpr uint32_t SQ = 0;
  static constexpr uint32_t SK = SQ + TILE;
  static constexpr uint32_t SV = SK + TILE;          // Vext [128][NE]
  static constexpr uint32_t SG = SV + EXT;           // G    [128][NE]
  static constexpr uint32_t SP = SG + EXT;           // P    [128][128]   (first holds the H tile)
  static constexpr uint32_t SDS = SP + kL * kL * 2;  // dS   [128][128]
  static constexpr uint32_t ST = DHP * NE * 2;       // one state tile
  static constexpr uint32_t SC = SDS + kL * kL * 2;  // [C|n] entering the chunk   [DHP][NE]  hi, lo
  static constexpr uint32_t SR = SC + 2 * ST;        // R entering from the right  [DHP][NE]  hi, lo
  static constexpr uint32_t VCOL = SR + 2 * ST;
  static constexpr uint32_t TOTAL = VCOL + kL * 4;
};

// 256 threads: thread = (row, half).  Both halves know the gate quantities of their row; half 0 owns the row's operand
// staging and the dQ / dK epilogue, half 1 the dV epilogue; the S / dP -> P / dS conversion of a row (the causal part, up to
// 128 columns) is split between the two.
template <int DHP>
__global__ void __launch_bounds__(2 * kThreads) mlstm_chunk_grad_kernel(
    const unsigned char* __restrict__ q_tiles, const unsigned char* __restrict__ k_tiles, const unsigned char* __restrict__ v_tiles,
    const unsigned char* __restrict__ h_tiles, const unsigned char* __restrict__ dh_tiles, const float* __restrict__ ig,
    const float* __restrict__ fg, const float* __restrict__ m_in, const float* __restrict__ den_in,
    const unsigned char* __restrict__ states, const float* __restrict__ m_prev, const unsigned char* __restrict__ rstates,
    const float* __restrict__ mu_next, int nc, float scale, float eps, unsigned char* __restrict__ dq, unsigned char* __restrict__ dk,
    unsigned char* __restrict__ dv, float* __restrict__ dig, float* __restrict__ dc_out, float* __restrict__ dc_tot) {
  using L = BwdSmem<DHP>;
  constexpr int NE = L::NE;
  constexpr uint32_t TILE = L::TILE, ST1 = L::ST, ST_BYTES = 2 * L::ST;   // hi + lo tiles
  constexpr uint32_t TMEM_COLS = DHP <= 32 ? 256u : 512u;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *sQ = smem + L::SQ, *sK = smem + L::SK, *sV = smem + L::SV, *sG = smem + L::SG, *sP = smem + L::SP,
                *sdS = smem + L::SDS, *sC = smem + L::SC, *sR = smem + L::SR;
  float* vcol = reinterpret_cast<float*>(smem + L::VCOL);
  // MMA groups are issued by lane 0 of three different warps (each tracks its own tcgen05.commit), so the descriptor
  // arithmetic runs in parallel and the epilogue of dQ overlaps the dK / dV products
  __shared__ __align__(8) uint64_t bar_load, bar_s, bar_p, bar_q, bar_k, bar_v;
  __shared__ uint32_t tmem_slot;
  __shared__ float red[8];

  const int tid = threadIdx.x & (kL - 1), hsel = threadIdx.x >> 7, warp = tid >> 5;     // tid = row, warp = row group
  const bool lead = threadIdx.x == 0;
  const int tile = blockIdx.x;
  const int c = tile % nc;
  const size_t grow = static_cast<size_t>(tile) * kL + tid;
  const bool has_prev = c > 0, has_next = c < nc - 1;

  if (lead) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_s, 1);
    mbar_init(&bar_p, 1);
    mbar_init(&bar_q, 1);
    mbar_init(&bar_k, 1);
    mbar_init(&bar_v, 1);
    mbar_fence_init();
  }
  __syncwarp();
  if (threadIdx.x < 32) tmem_alloc(&tmem_slot, TMEM_COLS);
  if (hsel == 0) write_ext_ones(sV, DHP, tid);
  __syncthreads();
  if (lead) {
    mbar_expect_tx(&bar_load, 5 * TILE + (has_prev ? ST_BYTES : 0) + (has_next ? ST_BYTES : 0));
    const size_t to = static_cast<size_t>(tile) * TILE;
    bulk_g2s(sQ, q_tiles + to, TILE, &bar_load);
    bulk_g2s(sK, k_tiles + to, TILE, &bar_load);
    bulk_g2s(sV, v_tiles + to, TILE, &bar_load);
    bulk_g2s(sG, dh_tiles + to, TILE, &bar_load);
    bulk_g2s(sP, h_tiles + to, TILE, &bar_load);
    if (has_prev) bulk_g2s(sC, states + static_cast<size_t>(tile) * ST_BYTES, ST_BYTES, &bar_load);
    if (has_next) bulk_g2s(sR, rstates + static_cast<size_t>(tile) * ST_BYTES, ST_BYTES, &bar_load);
  }
  // ---- gate quantities ----
  const float iv = ig[grow];
  const float lf = log_sigmoid(fg[grow]);
  const float m = m_in[grow], den = den_in[grow];
  float g;
  const float b = block_cumsum128(lf, red, &g);
  vcol[tid] = (iv - b) * kLog2e;
  const float urow = (b - m) * kLog2e + log2f(scale);
  const float w = has_prev ? __expf(b + m_prev[tile] - m) : 0.f;
  const float fac = has_next ? __expf(g - b + iv + mu_next[tile]) : 0.f;

  mbar_wait(&bar_load, 0);
  if (hsel == 0) build_G_row<DHP>(sG, sP, tid, m, den, eps);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    // S[t][s] = Q K^T            -> cols [0,128)
    umma_gemm(tmem, smem_u32(sQ), kL * 16, 128, smem_u32(sK), kL * 16, 128, umma_idesc(128, kL, false, false), DHP, false);
    umma_commit(&bar_s);
  } else if (threadIdx.x == 32) {
    // dP[t][s] = G Vext^T        -> cols [128,256)
    umma_gemm(tmem + 128, smem_u32(sG), kL * 16, 128, smem_u32(sV), kL * 16, 128, umma_idesc(128, kL, false, false), NE, false);
    umma_commit(&bar_p);
  }
  mbar_wait(&bar_s, 0);
  mbar_wait(&bar_p, 0);
  tc_fence_after();
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
  // row group `warp` needs column blocks 0..warp (2 * (warp + 1) half-blocks of 16 columns): the two halves take warp + 1
  // half-blocks each; the all-zero blocks behind the diagonal are cleared by half 0 (P) and half 1 (dS)
#pragma unroll 1
  for (int hb = hsel * (warp + 1); hb < (hsel + 1) * (warp + 1); ++hb) {
    const int blk = hb >> 1, s0 = hb * 16;
    float sv[16], dp[16];
    tmem_ld16(tmem + lane_base + s0, sv);
    tmem_ld16(tmem + lane_base + 128 + s0, dp);
    // only the diagonal block needs the causal mask
    if (blk < warp)
      decay_pair<false>(sv, dp, urow, vcol + s0, s0, tid, sP + tile_off16(kL, tid, s0 / 8), sdS + tile_off16(kL, tid, s0 / 8));
    else
      decay_pair<true>(sv, dp, urow, vcol + s0, s0, tid, sP + tile_off16(kL, tid, s0 / 8), sdS + tile_off16(kL, tid, s0 / 8));
  }
  {
    unsigned char* zt = hsel == 0 ? sP : sdS;
    for (int cg = (warp + 1) * 4; cg < 16; ++cg) *reinterpret_cast<uint4*>(zt + tile_off16(kL, tid, cg)) = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  {
    const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aG = smem_u32(sG), aP = smem_u32(sP), aS = smem_u32(sdS),
                   aC = smem_u32(sC), aR = smem_u32(sR);
    if (threadIdx.x == 0) {
      // dQ_intra[t][d] = sum_s dS[t][s] K[s][d]
      umma_gemm(tmem + 0 * DHP, aS, kL * 16, 128, aK, 128, kL * 16, umma_idesc(128, DHP, false, true), kL, false);
      // dQ_inter[t][d] = sum_e' G[t][e'] Cn[d][e']
      if (has_prev) {
        umma_gemm(tmem + 1 * DHP, aG, kL * 16, 128, aC, DHP * 16, 128, umma_idesc(128, DHP, false, false), NE, false);
        umma_gemm(tmem + 1 * DHP, aG, kL * 16, 128, aC + ST1, DHP * 16, 128, umma_idesc(128, DHP, false, false), NE, true);
      }
      umma_commit(&bar_q);
    } else if (threadIdx.x == 32) {
      // dK_intra[s][d] = sum_t dS[t][s] Q[t][d]
      umma_gemm(tmem + 2 * DHP, aS, 128, kL * 16, aQ, 128, kL * 16, umma_idesc(128, DHP, true, true), kL, false);
      // dK_inter[s][d] = sum_e' Vext[s][e'] R[d][e']
      if (has_next) {
        umma_gemm(tmem + 3 * DHP, aV, kL * 16, 128, aR, DHP * 16, 128, umma_idesc(128, DHP, false, false), NE, false);
        umma_gemm(tmem + 3 * DHP, aV, kL * 16, 128, aR + ST1, DHP * 16, 128, umma_idesc(128, DHP, false, false), NE, true);
      }
      umma_commit(&bar_k);
    } else if (threadIdx.x == 64) {
      // dV_intra[s][e] = sum_t P[t][s] G[t][e]
      umma_gemm(tmem + 4 * DHP, aP, 128, kL * 16, aG, 128, kL * 16, umma_idesc(128, DHP, true, true), kL, false);
      // dV_inter[s][e] = sum_d K[s][d] R[d][e]
      if (has_next) {
        umma_gemm(tmem + 5 * DHP, aK, kL * 16, 128, aR, 128, DHP * 16, umma_idesc(128, DHP, false, true), DHP, false);
        umma_gemm(tmem + 5 * DHP, aK, kL * 16, 128, aR + ST1, 128, DHP * 16, umma_idesc(128, DHP, false, true), DHP, true);
      }
      umma_commit(&bar_v);
    }
    __syncwarp();
  }
  // ---- epilogue: combine intra/inter, write dq/dk/dv rows, gate-gradient dot products; one pass per product ----
  float q_dq = 0.f, k_dk = 0.f;
  // dq / dk / dv leave as bf16 tiles in the layout of q / k / v
  unsigned char* dq_t = dq + static_cast<size_t>(tile) * TILE;
  unsigned char* dk_t = dk + static_cast<size_t>(tile) * TILE;
  unsigned char* dv_t = dv + static_cast<size_t>(tile) * TILE;
  if (hsel == 0) {
  mbar_wait(&bar_q, 0);
  tc_fence_after();
#pragma unroll 1
  for (int c0 = 0; c0 < DHP; c0 += 16) {
    float a[16], bb[16];
    tmem_ld16(tmem + lane_base + 0 * DHP + c0, a);
    if (has_prev) {
      tmem_ld16(tmem + lane_base + 1 * DHP + c0, bb);
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] += w * bb[i];
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float f[8];
      unpack8_bf16(*reinterpret_cast<const uint4*>(sQ + tile_off16(kL, tid, c0 / 8 + half)), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) q_dq += f[i] * a[half * 8 + i];
    }
    store16_bf16_tile(dq_t, kL, tid, c0, a);
  }
  mbar_wait(&bar_k, 0);
  tc_fence_after();
#pragma unroll 1
  for (int c0 = 0; c0 < DHP; c0 += 16) {
    float a[16], bb[16];
    tmem_ld16(tmem + lane_base + 2 * DHP + c0, a);
    if (has_next) {
      tmem_ld16(tmem + lane_base + 3 * DHP + c0, bb);
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] += fac * bb[i];
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float f[8];
      unpack8_bf16(*reinterpret_cast<const uint4*>(sK + tile_off16(kL, tid, c0 / 8 + half)), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) k_dk += f[i] * a[half * 8 + i];
    }
    store16_bf16_tile(dk_t, kL, tid, c0, a);
  }
  dig[grow] = k_dk;
  } else {
  mbar_wait(&bar_v, 0);
  tc_fence_after();
#pragma unroll 1
  for (int c0 = 0; c0 < DHP; c0 += 16) {
    float a[16], bb[16];
    tmem_ld16(tmem + lane_base + 4 * DHP + c0, a);
    if (has_next) {
      tmem_ld16(tmem + lane_base + 5 * DHP + c0, bb);
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] += fac * bb[i];
    }
    store16_bf16_tile(dv_t, kL, tid, c0, a);
  }
  }
  {
    // d log f = reverse cumulative sum of dc over the whole sequence: the chunk-local suffix sum is taken here, the carry of
    // the later chunks (sum of their totals) is added by mlstm_gate_finish_kernel
    float tot;
    const float suffix = block_rcumsum128(q_dq - k_dk, red, &tot);     // half 1 scans zeros
    if (hsel == 0) {
      dc_out[grow] = suffix;
      if (tid == 0) dc_tot[tile] = tot;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, TMEM_COLS);
}

// ------------------------------------------------------------------ phase B4
// dlf_u = (chunk-local suffix sum) + (totals of all later chunks); dfg = dlf * sigmoid(-fg).  One CTA per (b, head, chunk).
__global__ void __launch_bounds__(kThreads) mlstm_gate_finish_kernel(const float* __restrict__ dc_suffix, const float* __restrict__ dc_tot,
                                                                      const float* __restrict__ fg, int nc, float* __restrict__ dfg) {
  __shared__ float red[4];
  const int tile = blockIdx.x, tid = threadIdx.x;
  const int bh = tile / nc, c = tile % nc;
  float part = 0.f;
  for (int cc = c + 1 + tid; cc < nc; cc += kThreads) part += dc_tot[static_cast<size_t>(bh) * nc + cc];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((tid & 31) == 0) red[tid >> 5] = part;
  __syncthreads();
  const float carry = red[0] + red[1] + red[2] + red[3];
  const size_t o = static_cast<size_t>(tile) * kL + tid;
  dfg[o] = (dc_suffix[o] + carry) * (1.f / (1.f + __expf(fg[o])));
}

template <int DHP>
static int launch_bwd(const void* q, const void* k, const void* v, const float* ig, const float* fg, const void* h, const void* dh_t,
                      const float* m, const float* den, const void* states, const float* m_prev, int BH, int nc, int dh, float eps,
                      void* dq, void* dk, void* dv, float* dig, float* dfg, float* ws_dstate, float* ws_g, float* ws_lam,
                      void* rstates, float* mu_next, float* ws_dc, cudaStream_t st) {
  constexpr int NE = ext_cols(DHP);
  const float scale = 1.0f / sqrtf(static_cast<float>(dh));
  const int ntiles = BH * nc;
  if (state_ws_enabled(DHP)) {
    if (int rc = launch_chunk_rstate_ws(DHP, q, dh_t, h, fg, m, den, BH, nc, scale, eps, ws_dstate, ws_g, ws_lam, st)) return rc;
  } else {
    const size_t used = 3 * kL * DHP * 2 + kL * NE * 2, window = kL * DHP * 2 + 32768;
    const size_t smem = used > window ? used : window;
    cudaError_t e = cudaFuncSetAttribute(mlstm_chunk_rstate_kernel<DHP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ProfScope ps(K_CHUNK_RSTATE, st);
    mlstm_chunk_rstate_kernel<DHP><<<ntiles, kThreads, smem, st>>>((const unsigned char*)q, (const unsigned char*)dh_t, (const unsigned char*)h,
                                                                   fg, m, den, scale, eps, ws_dstate, ws_g, ws_lam);
  }
  if (int rc = launch_state_scan(DHP, ws_dstate, ws_g, ws_lam, BH, nc, 1, rstates, mu_next, st)) return rc;
  if constexpr (DHP == 128) {
    // the widest head does not fit one fused kernel (346 KB of shared memory, 1024 TMEM columns): three part-kernels
    if (int rc = launch_chunk_grad_wide(q, k, v, h, dh_t, ig, fg, m, den, states, m_prev, rstates, mu_next, BH, nc, scale, eps, dq, dk, dv,
                                        dig, ws_dc, ws_lam, st))
      return rc;
  } else if (grad_ws_enabled(DHP)) {
    if (int rc = launch_chunk_grad_ws(DHP, q, k, v, h, dh_t, ig, fg, m, den, states, m_prev, rstates, mu_next, BH, nc, scale, eps, dq, dk,
                                      dv, dig, ws_dc, ws_lam, st))
      return rc;
  } else {
    const size_t smem = BwdSmem<DHP>::TOTAL;
    cudaError_t e = cudaFuncSetAttribute(mlstm_chunk_grad_kernel<DHP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ProfScope ps(K_CHUNK_GRAD, st);
    mlstm_chunk_grad_kernel<DHP><<<ntiles, 2 * kThreads, smem, st>>>(
        (const unsigned char*)q, (const unsigned char*)k, (const unsigned char*)v, (const unsigned char*)h, (const unsigned char*)dh_t, ig, fg,
        m, den, (const unsigned char*)states, m_prev, (const unsigned char*)rstates, mu_next, nc, scale, eps, (unsigned char*)dq,
        (unsigned char*)dk, (unsigned char*)dv, dig, ws_dc,
        ws_lam /* free again after the reverse scan: receives the per-chunk totals of dc */);
  }
  {
    ProfScope ps(K_GATE_FINISH, st);
    mlstm_gate_finish_kernel<<<ntiles, kThreads, 0, st>>>(ws_dc, ws_lam, fg, nc, dfg);
  }
  return (int)cudaGetLastError();
}

}  // namespace xhved

using namespace xhved;

extern "C" int xhved_mlstm_bwd(const void* q_tiles, const void* k_tiles, const void* v_tiles, const float* ig, const float* fg,
                               const void* h_tiles, const void* dh_tiles, const float* m, const float* den, const void* states,
                               const float* m_prev, int BH, int nc, int dh, int dhp, float eps, void* dq, void* dk, void* dv, float* dig,
                               float* dfg, float* ws_dstate, float* ws_g, float* ws_amax, void* rstates, float* mu_next, float* ws_dc,
                               void* stream) {
  if (BH <= 0 || nc <= 0 || dh <= 0 || dh > dhp) return XHVED_ERR_BAD_SHAPE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (dhp) {
    case 16: return launch_bwd<16>(q_tiles, k_tiles, v_tiles, ig, fg, h_tiles, dh_tiles, m, den, states, m_prev, BH, nc, dh, eps, dq, dk, dv, dig, dfg, ws_dstate, ws_g, ws_amax, rstates, mu_next, ws_dc, st);
    case 32: return launch_bwd<32>(q_tiles, k_tiles, v_tiles, ig, fg, h_tiles, dh_tiles, m, den, states, m_prev, BH, nc, dh, eps, dq, dk, dv, dig, dfg, ws_dstate, ws_g, ws_amax, rstates, mu_next, ws_dc, st);
    case 64: return launch_bwd<64>(q_tiles, k_tiles, v_tiles, ig, fg, h_tiles, dh_tiles, m, den, states, m_prev, BH, nc, dh, eps, dq, dk, dv, dig, dfg, ws_dstate, ws_g, ws_amax, rstates, mu_next, ws_dc, st);
    case 128: return launch_bwd<128>(q_tiles, k_tiles, v_tiles, ig, fg, h_tiles, dh_tiles, m, den, states, m_prev, BH, nc, dh, eps, dq, dk, dv, dig, dfg, ws_dstate, ws_g, ws_amax, rstates, mu_next, ws_dc, st);
    default: return XHVED_ERR_UNSUPPORTED_DH;
  }
}
