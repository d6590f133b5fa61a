// Chunkwise mLSTM backward, phase B3 (chunk_grad) for the WIDEST head (dhp = 128) -- sm_100a.
//
// At dhp = 128 one tile of the fused chunk_grad kernels (mlstm_bwd.cu / mlstm_bwd_ws.cu) needs 346 KB of shared memory
// (the carried states C and R alone are 2 x 73 KB as hi/lo pairs) and 1024 tensor-memory columns (six 128-wide
// accumulators next to S and dP).  The gradient is therefore produced by THREE persistent part-kernels per tile, each
// of which fits (215 KB, 384 columns) and recomputes the shared intermediates it needs:
//
//   part Q:  dP = G Vext^T -> dS -> dQ = dS K + w (G [C|n]^T);             writes dq and q.dQ
//   part K:  dP = G Vext^T -> dS -> dK = dS^T Q + fac (Vext R^T);          writes dk, di = k.dK and the suffix sums of q.dQ - k.dK
//   part V:  S^T = K Q^T -> P^T (tensor memory) -> dV = P^T G + fac (K R); writes dv
//
// Same mathematics and the same conversion code paths as mlstm_bwd_ws.cu (gradient of vision_lstm.py:48-130 in the chunkwise
// form of SURVEY.md 8a-note): separable decay weights outside the diagonal blocks, G built in place over the dH tile, states
// as bf16 hi/lo pairs.  128 worker threads (one per row / column of the tile) + one control warp that issues the bulk loads
// and every tcgen05.mma; one tile in flight per CTA, one CTA per SM.
#include <stdlib.h>

#include "mlstm_common.cuh"
#include "prof.cuh"
#include "xhved.h"

namespace xhved {

int sm_count_cached();

namespace wide {

constexpr int DHP = 128;
constexpr int NE = ext_cols(DHP);                    // 144
constexpr uint32_t TILE = kL * DHP * 2;              // 32 KB
constexpr uint32_t EXT = kL * NE * 2;                // 36 KB
constexpr uint32_t ST1 = DHP * NE * 2;               // 36 KB: one state tile (hi or lo)
// shared memory: G (dH in place) | H, later dS | Vext (parts Q, K) or Q (part V) | K (parts Q, V) or Q (part K) | state hi, lo
constexpr uint32_t OFF_G = 0, OFF_H = EXT, OFF_V = EXT + TILE, OFF_X = 2 * EXT + TILE, OFF_S = 2 * EXT + 2 * TILE;
constexpr uint32_t OFF_AUX = OFF_S + 2 * ST1;
// fp32 arrays: u[128] | vcol[128] | ev[128] | eu[3][128] | vmax[4] | red[8]
constexpr int A_U = 0, A_V = 128, A_EV = 256, A_EU = 384, A_VMAX = 768, A_RED = 772;
constexpr uint32_t AUX = (772 + 8) * 4;
constexpr uint32_t SMEM = OFF_AUX + AUX;
constexpr int NTHREADS = 160;
// tensor memory: dP or S^T (-> P^T packed into [0,64)) | intra accumulator | inter accumulator
constexpr uint32_t T_IN = 0, T_OI = 128, T_OX = 256, TMEM_COLS = 512;

enum { PART_Q = 0, PART_K = 1, PART_V = 2 };

// Row-extended gradient G_t = [dh_t / N_t | db_t | 0] built in place over the dH tile (see build_G_row in mlstm_bwd.cu)
__device__ __forceinline__ void build_G_row(unsigned char* sG, const unsigned char* sH, int t, float m, float den, float eps) {
  const float flo = __expf(-m);
  const float r = 1.f / (fmaxf(fabsf(den), flo) + eps);
  float dhh = 0.f;
#pragma unroll 4
  for (int cg = 0; cg < DHP / 8; ++cg) {
    uint4* pg = reinterpret_cast<uint4*>(sG + tile_off16(kL, t, cg));
    float g[8], h[8];
    unpack8_bf16(*pg, g);
    unpack8_bf16(*reinterpret_cast<const uint4*>(sH + tile_off16(kL, t, cg)), h);
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

template <int PART>
__global__ void __launch_bounds__(NTHREADS, 1) mlstm_chunk_grad_wide_kernel(
    const unsigned char* __restrict__ q_tiles, const unsigned char* __restrict__ k_tiles, const unsigned char* __restrict__ v_tiles,
    const unsigned char* __restrict__ h_tiles, const unsigned char* __restrict__ dh_tiles, const float* __restrict__ ig,
    const float* __restrict__ fg, const float* __restrict__ m_in, const float* __restrict__ den_in,
    const unsigned char* __restrict__ states, const float* __restrict__ m_prev, const unsigned char* __restrict__ rstates,
    const float* __restrict__ mu_next, int nc, int ntiles, float scale, float eps, unsigned char* __restrict__ dout, float* __restrict__ dig,
    float* __restrict__ dc_out, float* __restrict__ dc_tot) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* aux = reinterpret_cast<float*>(smem + OFF_AUX);
  // bar_full: operands landed | bar_prep: G built, scans published (128) | bar_m1: dP / S^T ready | bar_conv: dS / P^T written (128)
  // bar_m2: accumulators ready | bar_tfree: the tile has left shared and tensor memory (128)
  __shared__ __align__(8) uint64_t bar_full, bar_prep, bar_m1, bar_conv, bar_m2, bar_tfree;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_my = (ntiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (threadIdx.x == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_prep, kL);
    mbar_init(&bar_m1, 1);
    mbar_init(&bar_conv, kL);
    mbar_init(&bar_m2, 1);
    mbar_init(&bar_tfree, kL);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, TMEM_COLS);
  // constant ext columns [1 | 0] of the V buffer (its bulk load only ever overwrites the first DHP columns)
  if (PART != PART_V && threadIdx.x < kL) write_ext_ones(smem + OFF_V, DHP, threadIdx.x);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t aG = smem_u32(smem + OFF_G), aH = smem_u32(smem + OFF_H), aV = smem_u32(smem + OFF_V), aX = smem_u32(smem + OFF_X),
                 aS = smem_u32(smem + OFF_S);

  if (warp == 4) {
    // ===================================================================== control: bulk loads + every tcgen05.mma
    if (lane == 0) {
      for (int it = 0; it < n_my; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        const int c = tile % nc;
        const bool has_state = PART == PART_Q ? c > 0 : c < nc - 1;      // part Q uses C (from the left), K and V use R (from the right)
        if (it > 0) mbar_wait(&bar_tfree, (it - 1) & 1);
        {
          const size_t to = static_cast<size_t>(tile) * TILE;
          mbar_expect_tx(&bar_full, 4 * TILE + (has_state ? 2 * ST1 : 0));
          bulk_g2s(smem + OFF_G, dh_tiles + to, TILE, &bar_full);
          bulk_g2s(smem + OFF_H, h_tiles + to, TILE, &bar_full);
          if (PART == PART_V) {
            bulk_g2s(smem + OFF_V, q_tiles + to, TILE, &bar_full);
            bulk_g2s(smem + OFF_X, k_tiles + to, TILE, &bar_full);
          } else {
            bulk_g2s(smem + OFF_V, v_tiles + to, TILE, &bar_full);
            bulk_g2s(smem + OFF_X, (PART == PART_Q ? k_tiles : q_tiles) + to, TILE, &bar_full);
          }
          if (has_state)
            bulk_g2s(smem + OFF_S, (PART == PART_Q ? states : rstates) + static_cast<size_t>(tile) * (2 * ST1), 2 * ST1, &bar_full);
        }
        mbar_wait(&bar_prep, it & 1);          // G built (every worker waited for bar_full itself)
        tc_fence_after();
        if (PART == PART_V) {
          // S^T[s][t] = sum_d K[s][d] Q[t][d]
          umma_gemm(tmem + T_IN, aX, kL * 16, 128, aV, kL * 16, 128, umma_idesc(128, kL, false, false), DHP, false);
          // dV_inter[s][e] = sum_d K[s][d] R[d][e]   (does not wait for the conversion)
          if (has_state) {
            umma_gemm(tmem + T_OX, aX, kL * 16, 128, aS, 128, DHP * 16, umma_idesc(128, DHP, false, true), DHP, false);
            umma_gemm(tmem + T_OX, aX, kL * 16, 128, aS + ST1, 128, DHP * 16, umma_idesc(128, DHP, false, true), DHP, true);
          }
        } else {
          // dP[t][s] = sum_e' G[t][e'] Vext[s][e']
          umma_gemm(tmem + T_IN, aG, kL * 16, 128, aV, kL * 16, 128, umma_idesc(128, kL, false, false), NE, false);
          if (has_state) {
            if (PART == PART_Q) {
              // dQ_inter[t][d] = sum_e' G[t][e'] Cn[d][e']
              umma_gemm(tmem + T_OX, aG, kL * 16, 128, aS, DHP * 16, 128, umma_idesc(128, DHP, false, false), NE, false);
              umma_gemm(tmem + T_OX, aG, kL * 16, 128, aS + ST1, DHP * 16, 128, umma_idesc(128, DHP, false, false), NE, true);
            } else {
              // dK_inter[s][d] = sum_e' Vext[s][e'] R[d][e']
              umma_gemm(tmem + T_OX, aV, kL * 16, 128, aS, DHP * 16, 128, umma_idesc(128, DHP, false, false), NE, false);
              umma_gemm(tmem + T_OX, aV, kL * 16, 128, aS + ST1, DHP * 16, 128, umma_idesc(128, DHP, false, false), NE, true);
            }
          }
        }
        umma_commit(&bar_m1);
        mbar_wait(&bar_conv, it & 1);          // dS in shared memory (over the H tile) / P^T in tensor memory
        tc_fence_after();
        if (PART == PART_Q) {
          // dQ_intra[t][d] = sum_s dS[t][s] K[s][d]
          umma_gemm(tmem + T_OI, aH, kL * 16, 128, aX, 128, kL * 16, umma_idesc(128, DHP, false, true), kL, false);
        } else if (PART == PART_K) {
          // dK_intra[s][d] = sum_t dS[t][s] Q[t][d]     (A = MN-major view of dS)
          umma_gemm(tmem + T_OI, aH, 128, kL * 16, aX, 128, kL * 16, umma_idesc(128, DHP, true, true), kL, false);
        } else {
          // dV_intra[s][e] = sum_t P^T[s][t] G[t][e]    (A = bf16 P^T in tensor memory)
          umma_gemm_ts(tmem + T_OI, tmem + T_IN, aG, 128, kL * 16, umma_idesc(128, DHP, false, true), kL, false);
        }
        umma_commit(&bar_m2);
      }
    }
  } else {
    // ===================================================================== workers: thread r = row t (parts Q, K) / column s (part V)
    const int w = warp, r = threadIdx.x;
    const uint32_t lane_base = static_cast<uint32_t>(w * 32) << 16;
    float* a_u = aux + A_U;
    float* a_v = aux + A_V;
    float* a_ev = aux + A_EV;
    float* a_eu = aux + A_EU;
    float* a_vmax = aux + A_VMAX;
    float* a_red = aux + A_RED;
    const float l2scale = log2f(scale);
    for (int it = 0; it < n_my; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int c = tile % nc;
      const bool has_prev = c > 0, has_next = c < nc - 1;
      const bool has_state = PART == PART_Q ? has_prev : has_next;
      const size_t grow = static_cast<size_t>(tile) * kL + r;
      // ---- gate scans over the 128 rows: b_t (inclusive cumsum of log sigmoid f), block maxima of v_s, all weights ----
      const float iv = ig[grow], fv = fg[grow], mv = m_in[grow], dn = den_in[grow];
      float x = log_sigmoid(fv);
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      if (lane == 31) a_red[w] = x;
      named_bar_sync(1, kL);
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
      if (lane == 0) a_vmax[w] = bm;
      const float urow = (b - mv) * kLog2e + l2scale;
      const float evs = fast_exp2(v2 - bm);
      // row weight of the inter-chunk part: w_t = exp(b_t + m_prev - m_t) (part Q), fac_s = exp(g - b_s + i_s + mu_next) (K, V)
      const float wx = !has_state ? 0.f : (PART == PART_Q ? __expf(b + m_prev[tile] - mv) : __expf(g - b + iv + mu_next[tile]));
      a_u[r] = urow;
      a_v[r] = v2;
      a_ev[r] = evs;
      named_bar_sync(1, kL);            // vmax[] complete (and a_red free again)
#pragma unroll
      for (int j = 0; j < 3; ++j) a_eu[j * kL + r] = (j < w) ? fast_exp2(urow + a_vmax[j]) : 0.f;
      // ---- row-extended gradient G (needs the landed dH / H tiles) ----
      mbar_wait(&bar_full, it & 1);
      build_G_row(smem + OFF_G, smem + OFF_H, r, mv, dn, eps);
      fence_proxy_async();
      mbar_arrive(&bar_prep);

      mbar_wait(&bar_m1, it & 1);       // (arrives after every worker's bar_prep: a_u / a_eu of all rows are visible)
      tc_fence_after();
      const uint32_t tI = tmem + T_IN + lane_base;
      if (PART != PART_V) {
        // ---- dS[t][s] = dP[t][s] * D''[t][s] (causal), bf16, to shared memory (over the H tile, which G no longer needs) ----
        unsigned char* drow = smem + OFF_H + static_cast<uint32_t>(r) * 16u;
#pragma unroll 1
        for (int j = 0; j < w; ++j) {      // blocks left of the diagonal: separable weights
          uint32_t dp[32];
          tmem_ld32_nowait(tI + 32 * j, dp);
          const float eu = a_eu[j * kL + r];
          tmem_wait_ld32(dp);
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            const float4 e0 = *reinterpret_cast<const float4*>(a_ev + 32 * j + i);
            const float4 e1 = *reinterpret_cast<const float4*>(a_ev + 32 * j + i + 4);
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
          tmem_ld32_nowait(tI + 32 * w, dp);
          tmem_wait_ld32(dp);
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            const float4 v0 = *reinterpret_cast<const float4*>(a_v + 32 * w + i);
            const float4 v1 = *reinterpret_cast<const float4*>(a_v + 32 * w + i + 4);
            const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float d = fast_exp2(urow + vv[e]);
              o[e] = (i + e <= lane) ? __uint_as_float(dp[i + e]) * d : 0.f;
            }
            *reinterpret_cast<uint4*>(drow + (4 * w + i / 8) * (kL * 16)) = pack8_bf16(o);
          }
        }
#pragma unroll 1
        for (int cg = 4 * (w + 1); cg < 16; ++cg) *reinterpret_cast<uint4*>(drow + cg * (kL * 16)) = make_uint4(0u, 0u, 0u, 0u);
        fence_proxy_async();
      } else {
        // ---- P^T[s][t] = S^T[s][t] * D''[t][s] (t >= s), bf16, back into tensor memory (block j -> cols [16j, 16j+16)) ----
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = 0u;
#pragma unroll 1
        for (int j = 0; j < w; ++j) tmem_st16(tI + 16 * j, pk);      // row blocks in front of this column block: zero
        {                                                            // the diagonal block: direct weights, causal mask
          uint32_t sv[32];
          tmem_ld32_nowait(tI + 32 * w, sv);
          tmem_wait_ld32(sv);
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 u4 = *reinterpret_cast<const float4*>(a_u + 32 * w + i);
            float p0 = __uint_as_float(sv[i]) * fast_exp2(u4.x + v2), p1 = __uint_as_float(sv[i + 1]) * fast_exp2(u4.y + v2);
            float p2 = __uint_as_float(sv[i + 2]) * fast_exp2(u4.z + v2), p3 = __uint_as_float(sv[i + 3]) * fast_exp2(u4.w + v2);
            p0 = (i >= lane) ? p0 : 0.f;
            p1 = (i + 1 >= lane) ? p1 : 0.f;
            p2 = (i + 2 >= lane) ? p2 : 0.f;
            p3 = (i + 3 >= lane) ? p3 : 0.f;
            pk[i / 2] = pack_bf16x2(p0, p1);
            pk[i / 2 + 1] = pack_bf16x2(p2, p3);
          }
          tmem_st16(tI + 16 * w, pk);
        }
#pragma unroll 1
        for (int j = w + 1; j < 4; ++j) {   // rows t of block j lie behind every column of block w: eu[w][t] * ev_s
          uint32_t sv[32];
          tmem_ld32_nowait(tI + 32 * j, sv);
          tmem_wait_ld32(sv);
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 e4 = *reinterpret_cast<const float4*>(a_eu + w * kL + 32 * j + i);
            pk[i / 2] = pack_bf16x2(__uint_as_float(sv[i]) * (e4.x * evs), __uint_as_float(sv[i + 1]) * (e4.y * evs));
            pk[i / 2 + 1] = pack_bf16x2(__uint_as_float(sv[i + 2]) * (e4.z * evs), __uint_as_float(sv[i + 3]) * (e4.w * evs));
          }
          tmem_st16(tI + 16 * j, pk);
        }
        tmem_wait_st();
      }
      tc_fence_before();
      mbar_arrive(&bar_conv);

      // ---- epilogue: this lane's row of the result = intra + weight * inter; gate-gradient dot products (parts Q, K) ----
      mbar_wait(&bar_m2, it & 1);
      tc_fence_after();
      float dot = 0.f;
      unsigned char* otile = dout + static_cast<size_t>(tile) * TILE;      // bf16 tile in the layout of q / k / v
#pragma unroll 1
      for (int c0 = 0; c0 < DHP; c0 += 16) {
        uint32_t a[16], bb[16];
        tmem_ld16_nowait(tmem + lane_base + T_OI + c0, a);
        if (has_state) {
          tmem_ld16_nowait(tmem + lane_base + T_OX + c0, bb);
          tmem_wait_ld16(bb);
        }
        tmem_wait_ld16(a);
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(a[i]);
        if (has_state) {
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] += wx * __uint_as_float(bb[i]);
        }
        if (PART != PART_V) {
          // q.dQ (part Q: the X buffer holds K, so q comes from global memory) / k.dK (part K: X holds Q, k from global memory)
          const unsigned char* src = (PART == PART_Q ? q_tiles : k_tiles) + static_cast<size_t>(tile) * TILE;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float qv[8];
            unpack8_bf16(__ldg(reinterpret_cast<const uint4*>(src + tile_off16(kL, r, c0 / 8 + half))), qv);
#pragma unroll
            for (int i = 0; i < 8; ++i) dot += qv[i] * f[half * 8 + i];
          }
        }
        store16_bf16_tile(otile, kL, r, c0, f);
      }
      tc_fence_before();
      mbar_arrive(&bar_tfree);
      if (PART == PART_Q) {
        dc_out[grow] = dot;               // q.dQ, picked up by part K (same stream, later launch)
      } else if (PART == PART_K) {
        dig[grow] = dot;
        // d log f = reverse cumulative sum of dc = q.dQ - k.dK over the whole sequence: chunk-local suffix sum here, the carry of
        // the later chunks is added by mlstm_gate_finish_kernel
        float xs = dc_out[grow] - dot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float y = __shfl_down_sync(0xffffffffu, xs, o);
          if (lane + o < 32) xs += y;
        }
        if (lane == 0) a_red[4 + w] = xs;
        named_bar_sync(1, kL);
        float so = 0.f, tot = 0.f;
#pragma unroll
        for (int ww = 0; ww < 4; ++ww) {
          const float t = a_red[4 + ww];
          if (ww > w) so += t;
          tot += t;
        }
        dc_out[grow] = xs + so;
        if (r == 0) dc_tot[tile] = tot;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

template <int PART>
static int launch_part(const void* q, const void* k, const void* v, const void* h, const void* dh_t, const float* ig, const float* fg,
                       const float* m, const float* den, const void* states, const float* m_prev, const void* rstates,
                       const float* mu_next, int BH, int nc, float scale, float eps, void* dout, float* dig, float* dc, float* dc_tot,
                       cudaStream_t st) {
  const int ntiles = BH * nc;
  cudaError_t e = cudaFuncSetAttribute(mlstm_chunk_grad_wide_kernel<PART>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
  if (e != cudaSuccess) return (int)e;
  const int grid = ntiles < sm_count_cached() ? ntiles : sm_count_cached();
  mlstm_chunk_grad_wide_kernel<PART><<<grid, NTHREADS, SMEM, st>>>(
      (const unsigned char*)q, (const unsigned char*)k, (const unsigned char*)v, (const unsigned char*)h, (const unsigned char*)dh_t, ig, fg,
      m, den, (const unsigned char*)states, m_prev, (const unsigned char*)rstates, mu_next, nc, ntiles, scale, eps, (unsigned char*)dout, dig, dc,
      dc_tot);
  return (int)cudaGetLastError();
}

}  // namespace wide

// phase B3 of the backward at dhp = 128: three part-kernels (dQ, then dK -- which consumes q.dQ --, then dV)
int launch_chunk_grad_wide(const void* q, const void* k, const void* v, const void* h, const void* dh_t, const float* ig, const float* fg,
                           const float* m, const float* den, const void* states, const float* m_prev, const void* rstates,
                           const float* mu_next, int BH, int nc, float scale, float eps, void* dq, void* dk, void* dv, float* dig,
                           float* dc, float* dc_tot, cudaStream_t st) {
  ProfScope ps(K_CHUNK_GRAD, st);
  if (int rc = wide::launch_part<wide::PART_Q>(q, k, v, h, dh_t, ig, fg, m, den, states, m_prev, rstates, mu_next, BH, nc, scale, eps, dq,
                                               dig, dc, dc_tot, st))
    return rc;
  if (int rc = wide::launch_part<wide::PART_K>(q, k, v, h, dh_t, ig, fg, m, den, states, m_prev, rstates, mu_next, BH, nc, scale, eps, dk,
                                               dig, dc, dc_tot, st))
    return rc;
  return wide::launch_part<wide::PART_V>(q, k, v, h, dh_t, ig, fg, m, den, states, m_prev, rstates, mu_next, BH, nc, scale, eps, dv, dig,
                                         dc, dc_tot, st);
}

}  // namespace xhved
