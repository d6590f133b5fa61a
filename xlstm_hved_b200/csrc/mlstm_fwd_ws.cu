// Chunkwise mLSTM forward, phase 3 (chunk_out) as a PERSISTENT, WARP-SPECIALISED kernel -- sm_100a.
//
// Same mathematics as mlstm_chunk_out_kernel (mlstm_fwd.cu; reference: vision_lstm.py:99-128 in the chunkwise form of
// SURVEY.md 8a-note), different machine mapping:
//
//   * one CTA per SM walks tiles blockIdx.x, +gridDim.x, ...; NSLOT tiles are in flight in tensor memory at any time (S 128
//     + O dhp+16 columns each), and the producer runs NSTAGE >= NSLOT shared-memory stages (Q, K, [V|1], state hi/lo) ahead,
//     so that a tile's operands have landed long before its TMEM slot frees up;
//   * warp roles: NSLOT consumer warpgroups (128 threads = the 128 rows / TMEM lanes of a tile), one producer warp
//     (cp.async.bulk into the stage ring, full/empty mbarriers), TWO MMA warps (S = Q K^T / O = P V; each runs in uniform
//     control flow and elects a lane inside every tcgen05 instruction);
//   * P = S o D' never touches shared memory: the consumer converts its row in registers and stores bf16 P back into the
//     S columns with tcgen05.st; the second product O += P [V|1] takes its A operand from TENSOR MEMORY;
//   * the inter-chunk part w_t * (q_t [C|n]) is folded into the same accumulator: Q [C|n] is issued together with S, the
//     consumer scales its row by w_t in TMEM (tcgen05.ld / st) and P [V|1] accumulates on top.  The normaliser input
//     den_t = sum_s P_ts + w_t q_t.n is taken from the fp32 weights BEFORE they are rounded to bf16 (plus column dhp of the
//     scaled inter-chunk product): summing the rounded P instead costs up to 4x in gradient accuracy where |den| sits near
//     its floor (tests/test_gpu_cell.py, oracle emulation);
//   * decay weights without one ex2 per (t,s): left of the diagonal block D'_ts = exp2(u_t + vmax_j) * exp2(v_s - vmax_j)
//     with vmax_j the maximum over the 32-column block j (second factor once per column, first once per row and block;
//     u_t + vmax_j <= 0 there because m_t is the row maximum, so nothing overflows and an underflowing second factor
//     only drops terms below 2^-126); the diagonal block keeps the direct form with the causal mask.
#include <stdlib.h>

#include "mlstm_common.cuh"
#include "prof.cuh"
#include "xhved.h"

namespace xhved {

template <int DHP>
struct FwdWs {
  static constexpr int NE = ext_cols(DHP);
  // dhp = 16: a slot is 128 columns -- the accumulators live in the half of the S columns that the bf16 P leaves free
  // (O_intra = P V at [64, 80), O_inter = Q [C|n] at [80, 112), both issued once P is written) -- so FOUR tiles are in flight
  static constexpr bool COMPACT = DHP <= 16;
  static constexpr int NSLOT = DHP <= 16 ? 4 : (DHP <= 64 ? 2 : 1);    // tiles in flight in tensor memory (one consumer warpgroup each)
  static constexpr int NSTAGE = DHP <= 16 ? 8 : (DHP <= 32 ? 4 : (DHP <= 64 ? 2 : 1));   // operand stages the producer runs ahead by
  static constexpr uint32_t TILE = kL * DHP * 2;
  static constexpr uint32_t VEXT = kL * NE * 2;
  static constexpr uint32_t ST1 = DHP * NE * 2;            // one state tile (hi or lo)
  static constexpr uint32_t OFF_Q = 0, OFF_K = TILE, OFF_V = 2 * TILE, OFF_S = 2 * TILE + VEXT;
  static constexpr uint32_t STAGE = 2 * TILE + VEXT + 2 * ST1;
  static constexpr uint32_t TM_SLOT = COMPACT ? 128 : 128 + NE;   // TMEM columns per slot: S | O (COMPACT: O inside the S columns)
  static constexpr uint32_t T_OI = COMPACT ? 64 : 128;            // O (intra; COMPACT: intra only) relative to the slot
  static constexpr uint32_t T_OX = 80;                            // COMPACT: Q [C|n], NE columns
  static constexpr int NTHREADS = (4 * NSLOT + 3) * 32;      // consumers | producer | S issuer | PV issuer
  // per-slot fp32 arrays behind the stages: vcol[128], ev[128], vmax[4], red_sum[4], red_max[4]
  static constexpr uint32_t AUX = (128 + 128 + 4 + 4 + 4) * 4;
  static constexpr uint32_t SMEM_USED = NSTAGE * STAGE + NSLOT * AUX;
  static constexpr uint32_t SMEM = SMEM_USED > 120 * 1024 ? SMEM_USED : 120 * 1024;   // > half an SM: one CTA per SM
};

template <int DHP>
__global__ void __launch_bounds__(FwdWs<DHP>::NTHREADS, 1) mlstm_chunk_out_ws_kernel(
    const unsigned char* __restrict__ q_tiles, const unsigned char* __restrict__ k_tiles, const unsigned char* __restrict__ v_tiles,
    const float* __restrict__ ig, const float* __restrict__ fg, const unsigned char* __restrict__ states,
    const float* __restrict__ m_prev, int nc, int ntiles, float scale, float eps, unsigned char* __restrict__ h_tiles,
    float* __restrict__ m_out, float* __restrict__ den_out) {
  using C = FwdWs<DHP>;
  constexpr int NE = C::NE, NSLOT = C::NSLOT, NSTAGE = C::NSTAGE;
  constexpr uint32_t TILE = C::TILE, ST1 = C::ST1;
  extern __shared__ __align__(128) unsigned char smem[];
  float* aux = reinterpret_cast<float*>(smem + NSTAGE * C::STAGE);
  __shared__ __align__(8) uint64_t bar_full[NSTAGE], bar_empty[NSTAGE], bar_s[NSLOT], bar_p[NSLOT], bar_o[NSLOT], bar_free[NSLOT];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_my = (ntiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(&bar_s[s], 1);
      mbar_init(&bar_p[s], kL);
      mbar_init(&bar_o[s], 1);
      mbar_init(&bar_free[s], kL);
    }
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  // constant ext columns [1 | 0] of every stage's V buffer (bulk loads only ever overwrite the first DHP columns)
  for (int i = threadIdx.x; i < NSTAGE * kL; i += blockDim.x) write_ext_ones(smem + (i / kL) * C::STAGE + C::OFF_V, DHP, i % kL);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 4 * NSLOT) {
    // ===================================================================== producer
    if (lane == 0) {
      for (int it = 0; it < n_my; ++it) {
        const int s = it % NSTAGE, use = it / NSTAGE;
        const int tile = blockIdx.x + it * gridDim.x;
        const bool has_state = (tile % nc) > 0;
        unsigned char* st = smem + s * C::STAGE;
        mbar_wait(&bar_empty[s], (use & 1) ^ 1);
        mbar_expect_tx(&bar_full[s], 3 * TILE + (has_state ? 2 * ST1 : 0));
        const size_t to = static_cast<size_t>(tile) * TILE;
        bulk_g2s(st + C::OFF_Q, q_tiles + to, TILE, &bar_full[s]);
        bulk_g2s(st + C::OFF_K, k_tiles + to, TILE, &bar_full[s]);
        bulk_g2s(st + C::OFF_V, v_tiles + to, TILE, &bar_full[s]);
        if (has_state) bulk_g2s(st + C::OFF_S, states + static_cast<size_t>(tile) * (2 * ST1), 2 * ST1, &bar_full[s]);
      }
    }
  } else if (warp == 4 * NSLOT + 1) {
    // ===================================================================== MMA issuer 1: S = Q K^T (and, outside COMPACT, Q [C|n])
    // Two issuing warps: a single thread that waits for three barriers and issues ~13 tcgen05 instructions per tile (~100
    // cycles each) is itself the bottleneck of the kernel (2,700 cycles per tile measured, whatever the number of slots).
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    for (int it = 0; it < n_my; ++it) {
      const int s = it % NSLOT, use = it / NSLOT, sg = it % NSTAGE, use_sg = it / NSTAGE;
      const int tile = blockIdx.x + it * gridDim.x;
      const bool has_state = (tile % nc) > 0;
      const uint32_t st = smem_u32(smem + sg * C::STAGE);
      const uint32_t tS = tm + s * C::TM_SLOT, tO = tS + 128;
      mbar_wait(&bar_free[s], (use & 1) ^ 1);     // epilogue of the previous tile in this slot has drained TMEM
      mbar_wait(&bar_full[sg], use_sg & 1);
      tc_fence_after();
      // S[t][s'] = sum_d Q[t][d] K[s'][d]
      umma_gemm_e(tS, st + C::OFF_Q, kL * 16, 128, st + C::OFF_K, kL * 16, 128, umma_idesc(128, kL, false, false), DHP, false);
      // O[t][e'] = sum_d Q[t][d] [C|n][d][e']   (hi + lo state tiles; B = MN-major view)
      if (has_state && !C::COMPACT) {
        umma_gemm_e(tO, st + C::OFF_Q, kL * 16, 128, st + C::OFF_S, 128, DHP * 16, umma_idesc(128, NE, false, true), DHP, false);
        umma_gemm_e(tO, st + C::OFF_Q, kL * 16, 128, st + C::OFF_S + ST1, 128, DHP * 16, umma_idesc(128, NE, false, true), DHP, true);
      }
      umma_commit_e(&bar_s[s]);
    }
  } else if (warp == 4 * NSLOT + 2) {
    // ===================================================================== MMA issuer 2: O (+)= P [V|1] once P is written
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    for (int it = 0; it < n_my; ++it) {
      const int s = it % NSLOT, use = it / NSLOT, sg = it % NSTAGE;
      const int tile = blockIdx.x + it * gridDim.x;
      const bool has_state = (tile % nc) > 0;
      const uint32_t st = smem_u32(smem + sg * C::STAGE);
      const uint32_t tS = tm + s * C::TM_SLOT, tO = tS + 128;
      mbar_wait(&bar_p[s], use & 1);
      tc_fence_after();
      if (C::COMPACT) {
        // O_intra[t][e] = sum_s' P[t][s'] V[s'][e] and O_inter[t][e'] = sum_d Q[t][d] [C|n][d][e'] into the S columns P left free
        umma_gemm_ts_e(tS + C::T_OI, tS, st + C::OFF_V, 128, kL * 16, umma_idesc(128, DHP, false, true), kL, false);
        if (has_state) {
          umma_gemm_e(tS + C::T_OX, st + C::OFF_Q, kL * 16, 128, st + C::OFF_S, 128, DHP * 16, umma_idesc(128, NE, false, true), DHP, false);
          umma_gemm_e(tS + C::T_OX, st + C::OFF_Q, kL * 16, 128, st + C::OFF_S + ST1, 128, DHP * 16, umma_idesc(128, NE, false, true), DHP, true);
        }
      } else {
        // O[t][e'] += sum_s' P[t][s'] [V|1][s'][e']   (A = bf16 P in TMEM, B = MN-major view of the V stage)
        umma_gemm_ts_e(tO, tS, st + C::OFF_V, 128, kL * 16, umma_idesc(128, NE, false, true), kL, has_state);
      }
      umma_commit_e(&bar_o[s]);
      umma_commit_e(&bar_empty[sg]);                // every MMA reading this stage has completed
    }
  } else {
    // ===================================================================== consumers: warpgroup wg owns slot wg
    const int wg = warp >> 2, w = warp & 3;
    const int r = threadIdx.x & (kL - 1);
    const uint32_t lane_base = static_cast<uint32_t>(w * 32) << 16;
    float* vcol = aux + wg * (C::AUX / 4);
    float* ev = vcol + 128;
    float* vmax = ev + 128;
    float* red_sum = vmax + 4;
    float* red_max = red_sum + 4;
    const float l2scale = log2f(scale);
    float iv_n = 0.f, fv_n = 0.f, mp_n = 0.f;
    if (wg < n_my) {
      const int tile = blockIdx.x + wg * gridDim.x;
      iv_n = ig[static_cast<size_t>(tile) * kL + r];
      fv_n = fg[static_cast<size_t>(tile) * kL + r];
      mp_n = (tile % nc) > 0 ? m_prev[tile] : -INFINITY;
    }
    for (int it = wg; it < n_my; it += NSLOT) {
      const int use = it / NSLOT;
      const int tile = blockIdx.x + it * gridDim.x;
      const bool has_state = (tile % nc) > 0;
      const size_t grow = static_cast<size_t>(tile) * kL + r;
      const float iv = iv_n, fv = fv_n, mp = mp_n;
      if (it + NSLOT < n_my) {   // gate values of this warpgroup's next tile
        const int tn = tile + NSLOT * gridDim.x;
        iv_n = ig[static_cast<size_t>(tn) * kL + r];
        fv_n = fg[static_cast<size_t>(tn) * kL + r];
        mp_n = (tn % nc) > 0 ? m_prev[tn] : -INFINITY;
      }
      // ---- gate scans over the 128 rows (vision_lstm.py:82-111 as 1-D scans): b_t, m_t ----
      float x = log_sigmoid(fv);
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      float cm = iv - x;                      // (i_s - b_s) up to the warp's offset
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float y = __shfl_up_sync(0xffffffffu, cm, o);
        if (lane >= o) cm = fmaxf(cm, y);
      }
      if (lane == 31) red_sum[w] = x, red_max[w] = cm;
      named_bar_sync(1 + wg, kL);
      float off = 0.f, cmx = -INFINITY;
#pragma unroll
      for (int ww = 0; ww < 3; ++ww) {
        if (ww < w) {
          cmx = fmaxf(cmx, red_max[ww] - off);
          off += red_sum[ww];
        }
      }
      const float b = x + off;
      const float vc = iv - b;
      const float m_intra = b + fmaxf(cm - off, cmx);
      const float m_inter = b + mp;
      const float m = fmaxf(m_intra, m_inter);
      const float wgt = has_state ? __expf(m_inter - m) : 0.f;
      const float urow = (b - m) * kLog2e + l2scale;
      // per-column factors of the separable decay: block maximum and exp2(v_s - vmax)
      const float v2 = vc * kLog2e;
      float bm = v2;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, o));
      vcol[r] = v2;
      ev[r] = fast_exp2(v2 - bm);
      if (lane == 0) vmax[w] = bm;
      named_bar_sync(1 + wg, kL);

      const uint32_t tS = tmem + wg * C::TM_SLOT + lane_base, tO = tS + C::T_OI;
      mbar_wait(&bar_s[wg], use & 1);
      tc_fence_after();
      // ---- inter-chunk part: O <- w_t * (Q [C|n]) in place; column dhp of it is w_t q_t.n, the inter-chunk part of den ----
      float den = 0.f;
      if (has_state && !C::COMPACT) {
#pragma unroll
        for (int c0 = 0; c0 < NE; c0 += 16) {
          uint32_t o[16];
          tmem_ld16_nowait(tO + c0, o);
          tmem_wait_ld16(o);
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * wgt);
          if (c0 == DHP) den = __uint_as_float(o[0]);
          tmem_st16(tO + c0, o);
        }
      }
      // ---- P = S o D' (causal), bf16, back into the S columns (block j -> columns [16j, 16j+16)) ----
      {
        uint32_t pk[16];
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll 1
        for (int j = 0; j < w; ++j) {          // blocks left of the diagonal: separable weights
          uint32_t sv[32];
          tmem_ld32_nowait(tS + 32 * j, sv);
          const float eu = fast_exp2(urow + vmax[j]);
          tmem_wait_ld32(sv);
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 e4 = *reinterpret_cast<const float4*>(ev + 32 * j + i);
            const float p0 = __uint_as_float(sv[i]) * (e4.x * eu), p1 = __uint_as_float(sv[i + 1]) * (e4.y * eu);
            const float p2 = __uint_as_float(sv[i + 2]) * (e4.z * eu), p3 = __uint_as_float(sv[i + 3]) * (e4.w * eu);
            rs0 += p0 + p1;
            rs1 += p2 + p3;
            pk[i / 2] = pack_bf16x2(p0, p1);
            pk[i / 2 + 1] = pack_bf16x2(p2, p3);
          }
          tmem_st16(tS + 16 * j, pk);
        }
        {                                       // the diagonal block: direct weights, causal mask
          uint32_t sv[32];
          tmem_ld32_nowait(tS + 32 * w, sv);
          tmem_wait_ld32(sv);
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 v4 = *reinterpret_cast<const float4*>(vcol + 32 * w + i);
            float p0 = __uint_as_float(sv[i]) * fast_exp2(urow + v4.x), p1 = __uint_as_float(sv[i + 1]) * fast_exp2(urow + v4.y);
            float p2 = __uint_as_float(sv[i + 2]) * fast_exp2(urow + v4.z), p3 = __uint_as_float(sv[i + 3]) * fast_exp2(urow + v4.w);
            p0 = (i <= lane) ? p0 : 0.f;
            p1 = (i + 1 <= lane) ? p1 : 0.f;
            p2 = (i + 2 <= lane) ? p2 : 0.f;
            p3 = (i + 3 <= lane) ? p3 : 0.f;
            rs0 += p0 + p1;
            rs1 += p2 + p3;
            pk[i / 2] = pack_bf16x2(p0, p1);
            pk[i / 2 + 1] = pack_bf16x2(p2, p3);
          }
          tmem_st16(tS + 16 * w, pk);
        }
        den += rs0 + rs1;                       // normaliser input: fp32 sum of the UNROUNDED weights (vision_lstm.py:123)
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = 0u;
#pragma unroll 1
        for (int j = w + 1; j < 4; ++j) tmem_st16(tS + 16 * j, pk);      // behind the diagonal: zero
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&bar_p[wg]);

      // ---- epilogue: h = O / (max(|den|, exp(-m)) + eps)   (vision_lstm.py:123-128) ----
      mbar_wait(&bar_o[wg], use & 1);
      tc_fence_after();
      if (C::COMPACT && has_state) {               // the inter-chunk accumulator was not pre-scaled: fold it in here
        uint32_t ox[16], on[16];
        uint32_t o0[16];
        tmem_ld16_nowait(tS + C::T_OX, ox);
        tmem_ld16_nowait(tS + C::T_OX + 16, on);
        tmem_ld16_nowait(tO, o0);
        tmem_wait_ld16(ox);
        tmem_wait_ld16(on);
        tmem_wait_ld16(o0);
        den += wgt * __uint_as_float(on[0]);
#pragma unroll
        for (int i = 0; i < 16; ++i) o0[i] = __float_as_uint(__uint_as_float(o0[i]) + wgt * __uint_as_float(ox[i]));
        tmem_st16(tO, o0);                          // (this lane's own columns: read back below)
        tmem_wait_st();
      }
      const float rn = 1.f / (fmaxf(fabsf(den), __expf(-m)) + eps);
      unsigned char* hdst = h_tiles + static_cast<size_t>(tile) * TILE;
#pragma unroll
      for (int c0 = 0; c0 < DHP; c0 += 16) {
        uint32_t o[16];
        tmem_ld16_nowait(tO + c0, o);
        tmem_wait_ld16(o);
        uint4 u0, u1;
        u0.x = pack_bf16x2(__uint_as_float(o[0]) * rn, __uint_as_float(o[1]) * rn);
        u0.y = pack_bf16x2(__uint_as_float(o[2]) * rn, __uint_as_float(o[3]) * rn);
        u0.z = pack_bf16x2(__uint_as_float(o[4]) * rn, __uint_as_float(o[5]) * rn);
        u0.w = pack_bf16x2(__uint_as_float(o[6]) * rn, __uint_as_float(o[7]) * rn);
        u1.x = pack_bf16x2(__uint_as_float(o[8]) * rn, __uint_as_float(o[9]) * rn);
        u1.y = pack_bf16x2(__uint_as_float(o[10]) * rn, __uint_as_float(o[11]) * rn);
        u1.z = pack_bf16x2(__uint_as_float(o[12]) * rn, __uint_as_float(o[13]) * rn);
        u1.w = pack_bf16x2(__uint_as_float(o[14]) * rn, __uint_as_float(o[15]) * rn);
        *reinterpret_cast<uint4*>(hdst + tile_off16(kL, r, c0 / 8)) = u0;
        *reinterpret_cast<uint4*>(hdst + tile_off16(kL, r, c0 / 8 + 1)) = u1;
      }
      tc_fence_before();
      mbar_arrive(&bar_free[wg]);
      m_out[grow] = m;
      den_out[grow] = den;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int sm_count_cached() {
  static int cached[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) dev = 0;
  if (!cached[dev]) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached[dev] = n > 0 ? n : 148;
  }
  return cached[dev];
}

template <int DHP>
static int launch_out_ws(const void* q, const void* k, const void* v, const float* ig, const float* fg, const void* states,
                         const float* m_prev, int BH, int nc, float scale, float eps, void* h, float* m, float* den,
                         cudaStream_t st) {
  using C = FwdWs<DHP>;
  const int ntiles = BH * nc;
  cudaError_t e = cudaFuncSetAttribute(mlstm_chunk_out_ws_kernel<DHP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
  if (e != cudaSuccess) return (int)e;
  const int grid = ntiles < sm_count_cached() ? ntiles : sm_count_cached();
  ProfScope ps(K_CHUNK_OUT, st);
  mlstm_chunk_out_ws_kernel<DHP><<<grid, C::NTHREADS, C::SMEM, st>>>(
      (const unsigned char*)q, (const unsigned char*)k, (const unsigned char*)v, ig, fg, (const unsigned char*)states, m_prev, nc,
      ntiles, scale, eps, (unsigned char*)h, m, den);
  return (int)cudaGetLastError();
}

// phase 3 of the forward on the persistent warp-specialised kernel
int launch_chunk_out_ws(int dhp, const void* q, const void* k, const void* v, const float* ig, const float* fg, const void* states,
                        const float* m_prev, int BH, int nc, float scale, float eps, void* h, float* m, float* den,
                        cudaStream_t st) {
  switch (dhp) {
    case 16: return launch_out_ws<16>(q, k, v, ig, fg, states, m_prev, BH, nc, scale, eps, h, m, den, st);
    case 32: return launch_out_ws<32>(q, k, v, ig, fg, states, m_prev, BH, nc, scale, eps, h, m, den, st);
    case 64: return launch_out_ws<64>(q, k, v, ig, fg, states, m_prev, BH, nc, scale, eps, h, m, den, st);
    case 128: return launch_out_ws<128>(q, k, v, ig, fg, states, m_prev, BH, nc, scale, eps, h, m, den, st);
    default: return XHVED_ERR_UNSUPPORTED_DH;
  }
}

}  // namespace xhved
