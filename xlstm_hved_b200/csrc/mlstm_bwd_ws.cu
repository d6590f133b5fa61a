// Chunkwise mLSTM backward, phase B3 (chunk_grad) as a PERSISTENT kernel with balanced, decoupled roles -- sm_100a.
//
// Same mathematics as mlstm_chunk_grad_kernel (mlstm_bwd.cu; gradient of vision_lstm.py:48-130 in the chunkwise form of
// SURVEY.md 8a-note), different machine mapping:
//
//   * a CTA walks tiles blockIdx.x, +gridDim.x, ...; barriers, tensor memory and the constant ext columns are set up once;
//     a control thread issues every bulk load and every tcgen05.mma;
//   * the two 128-thread groups split the L x L work by ORIENTATION, not by row halves: group 0 owns rows t and turns
//     dP[t][s] into dS (bf16, shared memory: it is the A operand of dQ = dS K and, transposed, of dK = dS^T Q); group 1
//     owns columns s of the TRANSPOSED product S^T = K Q^T and turns it into P^T in place in TENSOR MEMORY
//     (tcgen05.st), which is the A operand of dV = P^T G (an A operand in tensor memory cannot be transposed, hence the
//     transposed product).  Row group w of the causal tile needs w+1 column blocks, column group w needs 4-w row blocks, and
//     warp w of either group runs on SM sub-partition w: every sub-partition converts 5 blocks per tile instead of
//     2, 4, 6, 8 with the row-half split, and the 32 KB P tile no longer exists in shared memory;
//   * the groups are decoupled: dQ / dK are issued as soon as group 0 has written dS and are read back by group 0 (both are
//     indexed by this thread's lane, so q.dQ - k.dK needs no exchange); dV is issued when group 1 has written P^T and is read
//     back by group 1.  Each group has its own pair of mbarriers;
//   * the per-tile preparation of the NEXT tile (gate scans by group 0, the row-extended gradient G by group 1) runs while
//     the group waits for its own MMAs of the current tile; at dhp = 16 operands are double-buffered and S^T of the next
//     tile is issued as soon as dV has consumed P^T (the outputs live in the dead dP columns), i.e. during the epilogue;
//   * decay weights without one ex2 per (t, s) outside the diagonal blocks: D''_ts = exp2(u_t + vmax_j) * exp2(v_s - vmax_j)
//     with vmax_j the maximum of v over the 32-wide column block j (both factors <= 1 there, see mlstm_fwd_ws.cu).
#include <stdlib.h>

#include "mlstm_common.cuh"
#include "prof.cuh"
#include "xhved.h"

namespace xhved {

template <int DHP>
struct GradWs {
  static constexpr int NE = ext_cols(DHP);
  static constexpr uint32_t TILE = kL * DHP * 2, EXT = kL * NE * 2, ST1 = DHP * NE * 2;
  static constexpr bool PIPE = DHP <= 16;        // two operand stages, next tile prepared / started during this tile's tail
  static constexpr int NSTAGE = PIPE ? 2 : 1;
  static constexpr uint32_t OFF_Q = 0, OFF_K = TILE, OFF_V = 2 * TILE, OFF_G = 2 * TILE + EXT, OFF_H = 2 * TILE + 2 * EXT,
                            OFF_C = 3 * TILE + 2 * EXT, OFF_R = OFF_C + 2 * ST1;
  static constexpr uint32_t STAGE = OFF_R + 2 * ST1;
  static constexpr uint32_t OFF_DS = NSTAGE * STAGE;
  static constexpr uint32_t OFF_AUX = OFF_DS + kL * kL * 2;
  // fp32 arrays, double-buffered per tile parity: u[128] | vcol[128] | ev[128] | fac[128] | eu[3][128] | vmax[4] | pad
  static constexpr int A_U = 0, A_V = 128, A_EV = 256, A_FAC = 384, A_EU = 512, A_VMAX = 896, A_BUF = 904;
  static constexpr int A_RED = 2 * A_BUF;        // 8 floats of scan scratch (group 0)
  static constexpr uint32_t AUX = (2 * A_BUF + 8) * 4;
  static constexpr uint32_t SMEM = OFF_AUX + AUX;
  static constexpr uint32_t TMEM_COLS = DHP <= 32 ? 256u : 512u;
  static constexpr int NTHREADS = 288;
  static constexpr int CTAS_PER_SM = DHP <= 32 ? 2 : 1;
  // tensor-memory columns: S^T -> P^T at [0,128) (P^T packed into [0,64)), dP at [128,256); outputs reuse dead columns.
  // dhp = 16: every output lives in the dP columns, so [0,128) is free again as soon as dV has consumed P^T.
  static constexpr uint32_t T_ST = 0, T_DP = 128, T_DQI = 128;
  static constexpr uint32_t T_DQX = DHP == 64 ? 256 : 128 + DHP;
  static constexpr uint32_t T_DKI = DHP == 64 ? 192 : 128 + 2 * DHP;
  static constexpr uint32_t T_DKX = DHP == 64 ? 320 : 128 + 3 * DHP;
  static constexpr uint32_t T_DVI = DHP == 16 ? 192 : 64;
  static constexpr uint32_t T_DVX = DHP == 16 ? 208 : (DHP == 32 ? 96 : 384);
};

// Row-extended gradient G_t = [dh_t / N_t | db_t | 0] built in place over the dH tile (see build_G_row in mlstm_bwd.cu)
template <int DHP>
__device__ __forceinline__ void build_G_row_ws(unsigned char* sG, const unsigned char* sH, int t, float m, float den, float eps) {
  const float flo = __expf(-m);
  const float r = 1.f / (fmaxf(fabsf(den), flo) + eps);
  float dhh = 0.f;
#pragma unroll
  for (int cg = 0; cg < DHP / 8; ++cg) {
    uint4* pg = reinterpret_cast<uint4*>(sG + tile_off16(kL, t, cg));
    const uint4 ug = *pg;
    const uint4 uh = *reinterpret_cast<const uint4*>(sH + tile_off16(kL, t, cg));
    float g[8], h[8];
    unpack8_bf16(ug, g);
    unpack8_bf16(uh, h);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dhh += g[i] * h[i];
      g[i] *= r;
    }
    *pg = pack8_bf16(g);
  }
  const float dn = -dhh * r;
  const float db = (fabsf(den) > flo) ? (den >= 0.f ? dn : -dn) : 0.f;
  *reinterpret_cast<uint4*>(sG + tile_off16(kL, t, DHP / 8)) = make_uint4(pack_bf16x2(db, 0.f), 0u, 0u, 0u);
  *reinterpret_cast<uint4*>(sG + tile_off16(kL, t, DHP / 8 + 1)) = make_uint4(0u, 0u, 0u, 0u);
}

// out[c0 .. c0+16) = intra + wgt * inter for one row (lane) of two 16-column accumulator slices; both loads in flight together
__device__ __forceinline__ void load_combine16(uint32_t t_intra, uint32_t t_inter, bool has_inter, float wgt, float* f) {
  uint32_t a[16], b[16];
  tmem_ld16_nowait(t_intra, a);
  if (has_inter) {
    tmem_ld16_nowait(t_inter, b);
    tmem_wait_ld16(b);
  }
  tmem_wait_ld16(a);
#pragma unroll
  for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(a[i]);
  if (has_inter) {
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] += wgt * __uint_as_float(b[i]);
  }
}

template <int DHP>
__global__ void __launch_bounds__(GradWs<DHP>::NTHREADS, GradWs<DHP>::CTAS_PER_SM) mlstm_chunk_grad_ws_kernel(
    const unsigned char* __restrict__ q_tiles, const unsigned char* __restrict__ k_tiles, const unsigned char* __restrict__ v_tiles,
    const unsigned char* __restrict__ h_tiles, const unsigned char* __restrict__ dh_tiles, const float* __restrict__ ig,
    const float* __restrict__ fg, const float* __restrict__ m_in, const float* __restrict__ den_in,
    const unsigned char* __restrict__ states, const float* __restrict__ m_prev, const unsigned char* __restrict__ rstates,
    const float* __restrict__ mu_next, int nc, int ntiles, float scale, float eps, unsigned char* __restrict__ dq,
    unsigned char* __restrict__ dk, unsigned char* __restrict__ dv, float* __restrict__ dig, float* __restrict__ dc_out,
    float* __restrict__ dc_tot) {
  using C = GradWs<DHP>;
  constexpr int NE = C::NE, NSTAGE = C::NSTAGE;
  constexpr bool PIPE = C::PIPE;
  constexpr uint32_t TILE = C::TILE, ST1 = C::ST1;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sdS = smem + C::OFF_DS;
  float* aux = reinterpret_cast<float*>(smem + C::OFF_AUX);
  // bar_full: operands landed | bar_prep: scans published + G built (256) | bar_s / bar_p: S^T / dP in tensor memory
  // bar_c0 / bar_c1: dS / P^T written (128 each) | bar_ma / bar_mb: dQ,dK / dV accumulated | bar_tfree: tile left TMEM (256)
  __shared__ __align__(8) uint64_t bar_full[NSTAGE], bar_prep, bar_s, bar_p, bar_c0, bar_c1, bar_ma, bar_mb, bar_tfree;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_my = (ntiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) mbar_init(&bar_full[s], 1);
    mbar_init(&bar_prep, 2 * kL);
    mbar_init(&bar_s, 1);
    mbar_init(&bar_p, 1);
    mbar_init(&bar_c0, kL);
    mbar_init(&bar_c1, kL);
    mbar_init(&bar_ma, 1);
    mbar_init(&bar_mb, 1);
    mbar_init(&bar_tfree, 2 * kL);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, C::TMEM_COLS);
  // constant ext columns [1 | 0] of every stage's V buffer (bulk loads only ever overwrite the first DHP columns)
  for (int i = threadIdx.x; i < NSTAGE * kL; i += blockDim.x) write_ext_ones(smem + (i / kL) * C::STAGE + C::OFF_V, DHP, i % kL);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 8) {
    // ===================================================================== control: bulk loads + every tcgen05.mma
    // The WHOLE warp runs this loop in uniform control flow with warp-uniform operands (kernel parameters, blockIdx, loop
    // counters, the tensor-memory base through a shuffle); one lane is elected inside every issuing instruction (umma.cuh,
    // "elect" forms), so the tcgen05 / bulk-copy operands stay in uniform registers instead of going through a per-lane
    // waterfall of R2UR moves.
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    {
      auto issue_load = [&](int it) {
        const int s = it % NSTAGE;
        const int tile = blockIdx.x + it * gridDim.x;
        const int c = tile % nc;
        const bool has_prev = c > 0, has_next = c < nc - 1;
        unsigned char* st = smem + s * C::STAGE;
        mbar_expect_tx_e(&bar_full[s], 5 * TILE + (has_prev ? 2 * ST1 : 0) + (has_next ? 2 * ST1 : 0));
        const size_t to = static_cast<size_t>(tile) * TILE;
        bulk_g2s_e(st + C::OFF_Q, q_tiles + to, TILE, &bar_full[s]);
        bulk_g2s_e(st + C::OFF_K, k_tiles + to, TILE, &bar_full[s]);
        bulk_g2s_e(st + C::OFF_G, dh_tiles + to, TILE, &bar_full[s]);
        bulk_g2s_e(st + C::OFF_H, h_tiles + to, TILE, &bar_full[s]);
        bulk_g2s_e(st + C::OFF_V, v_tiles + to, TILE, &bar_full[s]);
        if (has_prev) bulk_g2s_e(st + C::OFF_C, states + static_cast<size_t>(tile) * (2 * ST1), 2 * ST1, &bar_full[s]);
        if (has_next) bulk_g2s_e(st + C::OFF_R, rstates + static_cast<size_t>(tile) * (2 * ST1), 2 * ST1, &bar_full[s]);
      };
      if (PIPE) issue_load(0);
      for (int it = 0; it < n_my; ++it) {
        const int s = it % NSTAGE;
        const int tile = blockIdx.x + it * gridDim.x;
        const int c = tile % nc;
        const bool has_prev = c > 0, has_next = c < nc - 1;
        const uint32_t st = smem_u32(smem + s * C::STAGE);
        const uint32_t aQ = st + C::OFF_Q, aK = st + C::OFF_K, aV = st + C::OFF_V, aG = st + C::OFF_G, aC = st + C::OFF_C,
                       aR = st + C::OFF_R, aS = smem_u32(sdS);
        // ---- S^T[s][t] = sum_d K[s][d] Q[t][d] -> cols [0,128): needs the raw operands and P^T of the previous tile consumed
        if (it > 0) mbar_wait(PIPE ? &bar_mb : &bar_tfree, (it - 1) & 1);
        if (!PIPE) issue_load(it);
        mbar_wait(&bar_full[s], (it / NSTAGE) & 1);
        tc_fence_after();
        umma_gemm_e(tmem_u + C::T_ST, aK, kL * 16, 128, aQ, kL * 16, 128, umma_idesc(128, kL, false, false), DHP, false);
        umma_commit_e(&bar_s);
        // ---- dP[t][s] = sum_e' G[t][e'] Vext[s][e'] -> cols [128,256): needs G and the previous tile's outputs read
        if (PIPE) {
          if (it > 0) mbar_wait(&bar_tfree, (it - 1) & 1);
          if (it + 1 < n_my) issue_load(it + 1);          // the stage of tile it-1 is free now
        }
        mbar_wait(&bar_prep, it & 1);
        tc_fence_after();
        umma_gemm_e(tmem_u + C::T_DP, aG, kL * 16, 128, aV, kL * 16, 128, umma_idesc(128, kL, false, false), NE, false);
        if (DHP > 32) {
          // wide heads: the inter-chunk products have their own columns and do not wait for the conversions
          if (has_prev) {
            umma_gemm_e(tmem_u + C::T_DQX, aG, kL * 16, 128, aC, DHP * 16, 128, umma_idesc(128, DHP, false, false), NE, false);
            umma_gemm_e(tmem_u + C::T_DQX, aG, kL * 16, 128, aC + ST1, DHP * 16, 128, umma_idesc(128, DHP, false, false), NE, true);
          }
          if (has_next) {
            umma_gemm_e(tmem_u + C::T_DKX, aV, kL * 16, 128, aR, DHP * 16, 128, umma_idesc(128, DHP, false, false), NE, false);
            umma_gemm_e(tmem_u + C::T_DKX, aV, kL * 16, 128, aR + ST1, DHP * 16, 128, umma_idesc(128, DHP, false, false), NE, true);
            umma_gemm_e(tmem_u + C::T_DVX, aK, kL * 16, 128, aR, 128, DHP * 16, umma_idesc(128, DHP, false, true), DHP, false);
            umma_gemm_e(tmem_u + C::T_DVX, aK, kL * 16, 128, aR + ST1, 128, DHP * 16, umma_idesc(128, DHP, false, true), DHP, true);
          }
        }
        umma_commit_e(&bar_p);
        // ---- dS written: dQ_intra[t][d] = sum_s dS[t][s] K[s][d],  dK_intra[s][d] = sum_t dS[t][s] Q[t][d] (MN-major A view)
        mbar_wait(&bar_c0, it & 1);
        tc_fence_after();
        umma_gemm_e(tmem_u + C::T_DQI, aS, kL * 16, 128, aK, 128, kL * 16, umma_idesc(128, DHP, false, true), kL, false);
        umma_gemm_e(tmem_u + C::T_DKI, aS, 128, kL * 16, aQ, 128, kL * 16, umma_idesc(128, DHP, true, true), kL, false);
        if (DHP <= 32) {
          if (has_prev) {
            // dQ_inter[t][d] = sum_e' G[t][e'] Cn[d][e']
            umma_gemm_e(tmem_u + C::T_DQX, aG, kL * 16, 128, aC, DHP * 16, 128, umma_idesc(128, DHP, false, false), NE, false);
            umma_gemm_e(tmem_u + C::T_DQX, aG, kL * 16, 128, aC + ST1, DHP * 16, 128, umma_idesc(128, DHP, false, false), NE, true);
          }
          if (has_next) {
            // dK_inter[s][d] = sum_e' Vext[s][e'] R[d][e']
            umma_gemm_e(tmem_u + C::T_DKX, aV, kL * 16, 128, aR, DHP * 16, 128, umma_idesc(128, DHP, false, false), NE, false);
            umma_gemm_e(tmem_u + C::T_DKX, aV, kL * 16, 128, aR + ST1, DHP * 16, 128, umma_idesc(128, DHP, false, false), NE, true);
          }
        }
        umma_commit_e(&bar_ma);
        // ---- P^T written: dV_intra[s][e] = sum_t P^T[s][t] G[t][e]  (A = bf16 P^T in tensor memory)
        mbar_wait(&bar_c1, it & 1);
        tc_fence_after();
        umma_gemm_ts_e(tmem_u + C::T_DVI, tmem_u + C::T_ST, aG, 128, kL * 16, umma_idesc(128, DHP, false, true), kL, false);
        if (DHP <= 32 && has_next) {
          // dV_inter[s][e] = sum_d K[s][d] R[d][e]
          umma_gemm_e(tmem_u + C::T_DVX, aK, kL * 16, 128, aR, 128, DHP * 16, umma_idesc(128, DHP, false, true), DHP, false);
          umma_gemm_e(tmem_u + C::T_DVX, aK, kL * 16, 128, aR + ST1, 128, DHP * 16, umma_idesc(128, DHP, false, true), DHP, true);
        }
        umma_commit_e(&bar_mb);
      }
    }
  } else if (warp < 4) {
    // ===================================================================== group 0: rows t.  scans, dP -> dS, dQ / dK epilogue
    const int w = warp, r = threadIdx.x;
    const uint32_t lane_base = static_cast<uint32_t>(w * 32) << 16;
    float* a_red = aux + C::A_RED;
    const float l2scale = log2f(scale);
    float urow = 0.f, wq = 0.f, fac = 0.f;
    // gate scans of tile `it` over the 128 rows: b_t (inclusive cumsum of log sigmoid f), block maxima of v_s, all weights
    auto scans = [&](int it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const size_t grow = static_cast<size_t>(tile) * kL + r;
      const int c = tile % nc;
      const bool has_prev = c > 0, has_next = c < nc - 1;
      const float iv = ig[grow], fv = fg[grow], mv = m_in[grow];
      const float mp = has_prev ? m_prev[tile] : 0.f, mun = has_next ? mu_next[tile] : 0.f;
      float* ab = aux + (it & 1) * C::A_BUF;
      float x = log_sigmoid(fv);
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      if (lane == 31) a_red[w] = x;
      named_bar_sync(2, kL);
      float off = 0.f, g = 0.f;
#pragma unroll
      for (int ww = 0; ww < 4; ++ww) {
        const float t = a_red[ww];
        if (ww < w) off += t;
        g += t;
      }
      const float b = x + off;
      const float v2 = (iv - b) * kLog2e;
      float bm = v2;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, o));
      if (lane == 0) ab[C::A_VMAX + w] = bm;
      urow = (b - mv) * kLog2e + l2scale;
      wq = has_prev ? __expf(b + mp - mv) : 0.f;
      fac = has_next ? __expf(g - b + iv + mun) : 0.f;
      ab[C::A_U + r] = urow;
      ab[C::A_V + r] = v2;
      ab[C::A_EV + r] = fast_exp2(v2 - bm);
      ab[C::A_FAC + r] = fac;
      named_bar_sync(2, kL);            // vmax[] complete (and a_red free again)
#pragma unroll
      for (int j = 0; j < 3; ++j)        // eu[j][t]: weight of row t towards column block j (j < row block of t)
        ab[C::A_EU + j * kL + r] = (j < w) ? fast_exp2(urow + ab[C::A_VMAX + j]) : 0.f;
    };
    scans(0);
    mbar_arrive(&bar_prep);
    for (int it = 0; it < n_my; ++it) {
      const int s = it % NSTAGE;
      const int tile = blockIdx.x + it * gridDim.x;
      const int c = tile % nc;
      const bool has_prev = c > 0, has_next = c < nc - 1;
      const size_t grow = static_cast<size_t>(tile) * kL + r;
      unsigned char* st = smem + s * C::STAGE;
      const float* ab = aux + (it & 1) * C::A_BUF;
      const float wq_cur = wq, fac_cur = fac, urow_cur = urow;
      mbar_wait(&bar_p, it & 1);
      tc_fence_after();
      // ---- dS[t][s] = dP[t][s] * D''[t][s] (causal), bf16, to shared memory ----
      {
        const uint32_t tD = tmem + C::T_DP + lane_base;
        unsigned char* drow = sdS + static_cast<uint32_t>(r) * 16u;
#pragma unroll 1
        for (int j = 0; j < w; ++j) {      // blocks left of the diagonal: separable weights
          uint32_t dp[32];
          tmem_ld32_nowait(tD + 32 * j, dp);
          const float eu = ab[C::A_EU + j * kL + r];
          tmem_wait_ld32(dp);
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            const float4 e0 = *reinterpret_cast<const float4*>(ab + C::A_EV + 32 * j + i);
            const float4 e1 = *reinterpret_cast<const float4*>(ab + C::A_EV + 32 * j + i + 4);
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(dp[i]) * (e0.x * eu), __uint_as_float(dp[i + 1]) * (e0.y * eu));
            o.y = pack_bf16x2(__uint_as_float(dp[i + 2]) * (e0.z * eu), __uint_as_float(dp[i + 3]) * (e0.w * eu));
            o.z = pack_bf16x2(__uint_as_float(dp[i + 4]) * (e1.x * eu), __uint_as_float(dp[i + 5]) * (e1.y * eu));
            o.w = pack_bf16x2(__uint_as_float(dp[i + 6]) * (e1.z * eu), __uint_as_float(dp[i + 7]) * (e1.w * eu));
            *reinterpret_cast<uint4*>(drow + (4 * j + i / 8) * (kL * 16)) = o;
          }
        }
        {                                   // the diagonal block: direct weights, causal mask
          uint32_t dp[32];
          tmem_ld32_nowait(tD + 32 * w, dp);
          tmem_wait_ld32(dp);
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            const float4 v0 = *reinterpret_cast<const float4*>(ab + C::A_V + 32 * w + i);
            const float4 v1 = *reinterpret_cast<const float4*>(ab + C::A_V + 32 * w + i + 4);
            const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float d = fast_exp2(urow_cur + vv[e]);
              o[e] = (i + e <= lane) ? __uint_as_float(dp[i + e]) * d : 0.f;
            }
            *reinterpret_cast<uint4*>(drow + (4 * w + i / 8) * (kL * 16)) = pack8_bf16(o);
          }
        }
#pragma unroll 1
        for (int cg = 4 * (w + 1); cg < 16; ++cg) *reinterpret_cast<uint4*>(drow + cg * (kL * 16)) = make_uint4(0u, 0u, 0u, 0u);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bar_c0);
      // ---- gate scans of the next tile while dQ / dK accumulate ----
      if (it + 1 < n_my) {
        scans(it + 1);
        mbar_arrive(&bar_prep);
      }
      // ---- epilogue: dQ and dK rows of this lane, gate-gradient dot products ----
      mbar_wait(&bar_full[s], (it / NSTAGE) & 1);      // (long complete) orders this thread's reads of the q / k rows
      mbar_wait(&bar_ma, it & 1);
      tc_fence_after();
      float q_dq = 0.f, k_dk = 0.f;
      // dq / dk / dv leave as bf16 tiles in the layout of q / k / v (the consumer, vil_pre_bwd_a, rounds them to bf16 operands anyway)
      unsigned char* dq_t = dq + static_cast<size_t>(tile) * TILE;
      unsigned char* dk_t = dk + static_cast<size_t>(tile) * TILE;
#pragma unroll 1
      for (int c0 = 0; c0 < DHP; c0 += 16) {
        float f[16];
        load_combine16(tmem + lane_base + C::T_DQI + c0, tmem + lane_base + C::T_DQX + c0, has_prev, wq_cur, f);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float qv[8];
          unpack8_bf16(*reinterpret_cast<const uint4*>(st + C::OFF_Q + tile_off16(kL, r, c0 / 8 + half)), qv);
#pragma unroll
          for (int i = 0; i < 8; ++i) q_dq += qv[i] * f[half * 8 + i];
        }
        store16_bf16_tile(dq_t, kL, r, c0, f);
        load_combine16(tmem + lane_base + C::T_DKI + c0, tmem + lane_base + C::T_DKX + c0, has_next, fac_cur, f);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float kv[8];
          unpack8_bf16(*reinterpret_cast<const uint4*>(st + C::OFF_K + tile_off16(kL, r, c0 / 8 + half)), kv);
#pragma unroll
          for (int i = 0; i < 8; ++i) k_dk += kv[i] * f[half * 8 + i];
        }
        store16_bf16_tile(dk_t, kL, r, c0, f);
      }
      tc_fence_before();
      mbar_arrive(&bar_tfree);
      dig[grow] = k_dk;
      // d log f = reverse cumulative sum of dc over the whole sequence: chunk-local suffix sum here, the carry of the later
      // chunks is added by mlstm_gate_finish_kernel
      float x = q_dq - k_dk;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float y = __shfl_down_sync(0xffffffffu, x, o);
        if (lane + o < 32) x += y;
      }
      if (lane == 0) a_red[4 + w] = x;
      named_bar_sync(2, kL);
      float off = 0.f, tot = 0.f;
#pragma unroll
      for (int ww = 0; ww < 4; ++ww) {
        const float t = a_red[4 + ww];
        if (ww > w) off += t;
        tot += t;
      }
      dc_out[grow] = x + off;
      if (r == 0) dc_tot[tile] = tot;
    }
  } else {
    // ===================================================================== group 1: columns s.  G, S^T -> P^T, dV epilogue
    const int w = warp & 3, r = threadIdx.x & (kL - 1);
    const uint32_t lane_base = static_cast<uint32_t>(w * 32) << 16;
    auto build_g = [&](int it) {
      const int s = it % NSTAGE;
      const int tile = blockIdx.x + it * gridDim.x;
      const size_t grow = static_cast<size_t>(tile) * kL + r;
      const float mv = m_in[grow], dn = den_in[grow];
      unsigned char* st = smem + s * C::STAGE;
      mbar_wait(&bar_full[s], (it / NSTAGE) & 1);
      build_G_row_ws<DHP>(st + C::OFF_G, st + C::OFF_H, r, mv, dn, eps);
      fence_proxy_async();
      mbar_arrive(&bar_prep);
    };
    if (PIPE) build_g(0);
    for (int it = 0; it < n_my; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int c = tile % nc;
      const bool has_next = c < nc - 1;
      const float* ab = aux + (it & 1) * C::A_BUF;
      if (!PIPE) build_g(it);
      mbar_wait(&bar_prep, it & 1);          // the scans of this tile are published
      const float vs = ab[C::A_V + r], evs = ab[C::A_EV + r], fac = ab[C::A_FAC + r];
      mbar_wait(&bar_s, it & 1);
      tc_fence_after();
      // ---- P^T[s][t] = S^T[s][t] * D''[t][s] (t >= s), bf16, back into tensor memory (block j -> cols [16j, 16j+16)) ----
      {
        const uint32_t tS = tmem + C::T_ST + lane_base;
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = 0u;
#pragma unroll 1
        for (int j = 0; j < w; ++j) tmem_st16(tS + 16 * j, pk);      // row blocks in front of this column block: zero
        {                                                            // the diagonal block: direct weights, causal mask
          uint32_t sv[32];
          tmem_ld32_nowait(tS + 32 * w, sv);
          tmem_wait_ld32(sv);
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 u4 = *reinterpret_cast<const float4*>(ab + C::A_U + 32 * w + i);
            float p0 = __uint_as_float(sv[i]) * fast_exp2(u4.x + vs), p1 = __uint_as_float(sv[i + 1]) * fast_exp2(u4.y + vs);
            float p2 = __uint_as_float(sv[i + 2]) * fast_exp2(u4.z + vs), p3 = __uint_as_float(sv[i + 3]) * fast_exp2(u4.w + vs);
            p0 = (i >= lane) ? p0 : 0.f;
            p1 = (i + 1 >= lane) ? p1 : 0.f;
            p2 = (i + 2 >= lane) ? p2 : 0.f;
            p3 = (i + 3 >= lane) ? p3 : 0.f;
            pk[i / 2] = pack_bf16x2(p0, p1);
            pk[i / 2 + 1] = pack_bf16x2(p2, p3);
          }
          tmem_st16(tS + 16 * w, pk);
        }
#pragma unroll 1
        for (int j = w + 1; j < 4; ++j) {   // rows t of block j lie behind every column of block w: eu[w][t] * ev_s
          uint32_t sv[32];
          tmem_ld32_nowait(tS + 32 * j, sv);
          tmem_wait_ld32(sv);
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 e4 = *reinterpret_cast<const float4*>(ab + C::A_EU + w * kL + 32 * j + i);
            pk[i / 2] = pack_bf16x2(__uint_as_float(sv[i]) * (e4.x * evs), __uint_as_float(sv[i + 1]) * (e4.y * evs));
            pk[i / 2 + 1] = pack_bf16x2(__uint_as_float(sv[i + 2]) * (e4.z * evs), __uint_as_float(sv[i + 3]) * (e4.w * evs));
          }
          tmem_st16(tS + 16 * j, pk);
        }
        tmem_wait_st();
      }
      tc_fence_before();
      mbar_arrive(&bar_c1);
      // ---- G of the next tile while dV accumulates ----
      if (PIPE && it + 1 < n_my) build_g(it + 1);
      // ---- epilogue: dV row of this lane ----
      mbar_wait(&bar_mb, it & 1);
      tc_fence_after();
      unsigned char* dv_t = dv + static_cast<size_t>(tile) * TILE;
#pragma unroll 1
      for (int c0 = 0; c0 < DHP; c0 += 16) {
        float f[16];
        load_combine16(tmem + lane_base + C::T_DVI + c0, tmem + lane_base + C::T_DVX + c0, has_next, fac, f);
        store16_bf16_tile(dv_t, kL, r, c0, f);
      }
      tc_fence_before();
      mbar_arrive(&bar_tfree);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, C::TMEM_COLS);
}

int sm_count_cached();

template <int DHP>
static int launch_grad_ws(const void* q, const void* k, const void* v, const void* h, const void* dh_t, const float* ig, const float* fg,
                          const float* m, const float* den, const void* states, const float* m_prev, const void* rstates,
                          const float* mu_next, int BH, int nc, float scale, float eps, void* dq, void* dk, void* dv, float* dig,
                          float* dc, float* dc_tot, cudaStream_t st) {
  using C = GradWs<DHP>;
  const int ntiles = BH * nc;
  cudaError_t e = cudaFuncSetAttribute(mlstm_chunk_grad_ws_kernel<DHP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
  if (e != cudaSuccess) return (int)e;
  const int cap = sm_count_cached() * C::CTAS_PER_SM;
  const int grid = ntiles < cap ? ntiles : cap;
  ProfScope ps(K_CHUNK_GRAD, st);
  mlstm_chunk_grad_ws_kernel<DHP><<<grid, C::NTHREADS, C::SMEM, st>>>(
      (const unsigned char*)q, (const unsigned char*)k, (const unsigned char*)v, (const unsigned char*)h, (const unsigned char*)dh_t, ig, fg,
      m, den, (const unsigned char*)states, m_prev, (const unsigned char*)rstates, mu_next, nc, ntiles, scale, eps, (unsigned char*)dq,
      (unsigned char*)dk, (unsigned char*)dv, dig, dc, dc_tot);
  return (int)cudaGetLastError();
}

// phase B3 of the backward on the persistent kernel; dhp in {16, 32, 64}
int launch_chunk_grad_ws(int dhp, const void* q, const void* k, const void* v, const void* h, const void* dh_t, const float* ig,
                         const float* fg, const float* m, const float* den, const void* states, const float* m_prev, const void* rstates,
                         const float* mu_next, int BH, int nc, float scale, float eps, void* dq, void* dk, void* dv, float* dig,
                         float* dc, float* dc_tot, cudaStream_t st) {
  switch (dhp) {
    case 16: return launch_grad_ws<16>(q, k, v, h, dh_t, ig, fg, m, den, states, m_prev, rstates, mu_next, BH, nc, scale, eps, dq, dk, dv, dig, dc, dc_tot, st);
    case 32: return launch_grad_ws<32>(q, k, v, h, dh_t, ig, fg, m, den, states, m_prev, rstates, mu_next, BH, nc, scale, eps, dq, dk, dv, dig, dc, dc_tot, st);
    case 64: return launch_grad_ws<64>(q, k, v, h, dh_t, ig, fg, m, den, states, m_prev, rstates, mu_next, BH, nc, scale, eps, dq, dk, dv, dig, dc, dc_tot, st);
    default: return XHVED_ERR_UNSUPPORTED_DH;
  }
}

}  // namespace xhved
