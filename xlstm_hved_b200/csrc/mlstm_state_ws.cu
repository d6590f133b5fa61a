// Chunkwise mLSTM, phases 1 / B1 (chunk_state, chunk_rstate) as ONE persistent, warp-specialised kernel -- sm_100a.
//
// Same mathematics as mlstm_chunk_state_kernel (mlstm_fwd.cu) and mlstm_chunk_rstate_kernel (mlstm_bwd.cu; SURVEY.md 8a-note:
// the contribution of one chunk to the carried state  dC_c = sum_j exp(a_j - amax_c) (k_j/sqrt(DH)) [v_j | 1]^T  of
// vision_lstm.py:48-130 in chunkwise form, and its mirror image  dR_c = sum_t exp(b_t - m_t - lambda_c) (q_t/sqrt(DH)) G_t^T
// of the backward pass).  The one-tile-per-CTA kernels are a single dependent chain per tile -- bulk load -> gate scans -> row
// scaling -> eight tcgen05.mma -> tcgen05.ld -> store -- that six resident CTAs per SM only partly overlap (25 / 26 us for
// 4096 tiles, 18 % / 27 % of HBM).  Here a CTA walks tiles blockIdx.x, +gridDim.x, ... with one role per warp:
//
//   warp P   (producer)   streams the operand tiles of the next NSTAGE tiles into a ring of shared-memory stages (bulk copies)
//   warps 0-7 (consumers) two groups of 128 threads that take alternate tiles; thread = chunk row: gate scans (shuffles + ONE
//                         named barrier: cumulative sum and maximum share the exchange), row weights, the row-scaled A operand
//                         as a bf16 hi + lo pair in place (and, backward, the row-extended gradient G), gates of the group's
//                         NEXT tile prefetched into registers.  A group's tile is one dependent chain of ~2,000 cycles: two
//                         groups per CTA and two CTAs per SM keep four of them in flight
//   warp M   (MMA)        issues the K = 128 chain of a tile as soon as its stage is ready, into one of two TMEM slots
//   warps E  (epilogue)   drain a slot (rows d: hi part, rows d + DHP: lo part of the SAME product, see chunk_state) and store
//
// so the scans of tile i+1, the MMAs of tile i and the epilogue of tile i-1 run concurrently, barriers / tensor memory are set
// up once per CTA, and two CTAs share an SM.
#include "mlstm_common.cuh"
#include "prof.cuh"
#include "xhved.h"

namespace xhved {

template <int DHP, bool REV>
struct StateWs {
  static constexpr int NE = ext_cols(DHP);
  static constexpr uint32_t TILE = kL * DHP * 2, EXT = kL * NE * 2;
  // stage: [A hi][A lo][B ext = Vext or G][H (backward only)].  A hi / lo are read as ONE 128-row MN-major operand, i.e.
  // through a 32 KB window from A hi: rows [0, DHP) = hi, [DHP, 2 DHP) = lo, the rest runs on over the following bytes
  // (next stage / padding) and only produces accumulator lanes nobody reads
  static constexpr uint32_t OFF_A = 0, OFF_ALO = TILE, OFF_B = 2 * TILE, OFF_H = 2 * TILE + EXT;
  static constexpr uint32_t STAGE = 2 * TILE + EXT + (REV ? TILE : 0);
  static constexpr int NSTAGE = DHP <= 16 ? 4 : (REV ? 2 : 3);
  static constexpr uint32_t WINDOW = 32768;
  static constexpr uint32_t DATA = (NSTAGE * STAGE > (NSTAGE - 1) * STAGE + WINDOW) ? NSTAGE * STAGE : (NSTAGE - 1) * STAGE + WINDOW;
  static constexpr int EPI = (2 * DHP) / 32;                        // epilogue warps: TMEM lanes [0, 2 DHP)
  static constexpr uint32_t OFF_SCR = DATA;                          // fp32 [DHP][NE]: lo rows on their way to the hi warp (DHP = 32)
  static constexpr uint32_t OFF_AUX = OFF_SCR + (EPI > 1 ? DHP * NE * 4 : 0);
  static constexpr uint32_t SMEM = OFF_AUX + 2 * 2 * 8 * 4;          // scan scratch: per consumer group, double-buffered
  static constexpr int W_EPI = 8, W_PROD = 8 + EPI, W_MMA = 9 + EPI;
  static constexpr int NTHREADS = (10 + EPI) * 32;
  static constexpr uint32_t SLOT = NE;                               // TMEM columns of one accumulator slot
  static constexpr uint32_t TMEM_COLS = next_pow2_cols(2 * NE);
  static_assert(DHP == 16 || DHP == 32, "epilogue warps cover TMEM lanes [0, 64)");
};

// Row-extended gradient G_t = [dh_t / N_t | db_t | 0] built in place over the dH tile (see build_G_row in mlstm_bwd.cu)
template <int DHP>
__device__ __forceinline__ void state_build_G_row(unsigned char* sG, const unsigned char* sH, int t, float m, float den, float eps) {
  const float flo = __expf(-m);
  const float r = 1.f / (fmaxf(fabsf(den), flo) + eps);
  float dhh = 0.f;
#pragma unroll
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

// forward (REV = false): a = K tiles, b = V tiles, g0 = i gates, g1 = f gates.
// backward (REV = true): a = Q tiles, b = dH tiles, h = H tiles, g1 = f gates, m_in / den_in the saved stabiliser / normaliser.
template <int DHP, bool REV>
__global__ void __launch_bounds__(StateWs<DHP, REV>::NTHREADS, 2) mlstm_chunk_state_ws_kernel(
    const unsigned char* __restrict__ a_tiles, const unsigned char* __restrict__ b_tiles, const unsigned char* __restrict__ h_tiles,
    const float* __restrict__ g0, const float* __restrict__ g1, const float* __restrict__ m_in, const float* __restrict__ den_in,
    int ntiles, float scale, float eps, float* __restrict__ dstate, float* __restrict__ g_out, float* __restrict__ amax_out) {
  using C = StateWs<DHP, REV>;
  constexpr int NE = C::NE, NSTAGE = C::NSTAGE;
  constexpr uint32_t TILE = C::TILE;
  extern __shared__ __align__(128) unsigned char smem[];
  float* aux = reinterpret_cast<float*>(smem + C::OFF_AUX);
  // full: operands landed (tx) | ready: consumers done with the stage (128) | sfree: the MMAs have read the stage (commit)
  // mdone: accumulator slot complete (commit) | tfree: slot drained (32 per epilogue warp)
  __shared__ __align__(8) uint64_t bar_full[NSTAGE], bar_ready[NSTAGE], bar_sfree[NSTAGE], bar_mdone[2], bar_tfree[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_my = (ntiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_ready[s], kL);
      mbar_init(&bar_sfree[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_mdone[s], 1);
      mbar_init(&bar_tfree[s], 32 * C::EPI);
    }
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, C::TMEM_COLS);
  if (!REV) {
    // constant ext columns [1 | 0] behind the V tile of every stage (bulk loads only ever overwrite the first DHP columns)
    for (int i = threadIdx.x; i < NSTAGE * kL; i += blockDim.x) write_ext_ones(smem + (i / kL) * C::STAGE + C::OFF_B, DHP, i % kL);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == C::W_PROD) {
    // ===================================================================== producer: operand tiles through the stage ring
    for (int it = 0; it < n_my; ++it) {
      const int s = it % NSTAGE;
      const size_t to = static_cast<size_t>(blockIdx.x + it * gridDim.x) * TILE;
      if (it >= NSTAGE) mbar_wait(&bar_sfree[s], (it / NSTAGE - 1) & 1);
      unsigned char* st = smem + s * C::STAGE;
      mbar_expect_tx_e(&bar_full[s], (REV ? 3 : 2) * TILE);
      bulk_g2s_e(st + C::OFF_A, a_tiles + to, TILE, &bar_full[s]);
      bulk_g2s_e(st + C::OFF_B, b_tiles + to, TILE, &bar_full[s]);
      if (REV) bulk_g2s_e(st + C::OFF_H, h_tiles + to, TILE, &bar_full[s]);
    }
  } else if (warp == C::W_MMA) {
    // ===================================================================== MMA issuer (whole warp, elect forms)
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    for (int it = 0; it < n_my; ++it) {
      const int s = it % NSTAGE, slot = it & 1;
      const uint32_t st = smem_u32(smem + s * C::STAGE);
      mbar_wait(&bar_ready[s], (it / NSTAGE) & 1);
      if (it >= 2) mbar_wait(&bar_tfree[slot], ((it >> 1) - 1) & 1);
      tc_fence_after();
      // D[d][e'] = sum_j A~[j][d] * B[j][e']   (A, B both MN-major views of row-j tiles; rows [DHP, 2 DHP) = the lo half)
      umma_gemm_e(tmem_u + slot * C::SLOT, st + C::OFF_A, 128, kL * 16, st + C::OFF_B, 128, kL * 16, umma_idesc(128, NE, true, true), kL,
                  false);
      umma_commit_e(&bar_sfree[s]);
      umma_commit_e(&bar_mdone[slot]);
    }
  } else if (warp >= C::W_EPI) {
    // ===================================================================== epilogue: TMEM lanes [0, 2 DHP) -> dstate
    const int e = warp - C::W_EPI;                                 // TMEM quadrant of this warp
    const uint32_t lane_base = static_cast<uint32_t>(e * 32) << 16;
    float* scr = reinterpret_cast<float*>(smem + C::OFF_SCR);
    for (int it = 0; it < n_my; ++it) {
      const int slot = it & 1;
      const int tile = blockIdx.x + it * gridDim.x;
      float* out = dstate + static_cast<size_t>(tile) * DHP * NE;
      mbar_wait(&bar_mdone[slot], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < NE; c0 += 16) {
        float v[16];
        tmem_ld16(tmem + lane_base + slot * C::SLOT + c0, v);
        if (C::EPI == 1) {
          // DHP = 16: lanes d hold the hi rows, lanes d + 16 the lo rows
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += __shfl_down_sync(0xffffffffu, v[i], 16);
          if (lane < 16) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(out + lane * NE + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          }
        } else if (e == 1) {
#pragma unroll
          for (int i = 0; i < 16; ++i) scr[(c0 + i) * DHP + lane] = v[i];
        } else {
          if (c0 == 0) named_bar_sync(3, 64);                        // the lo rows of this tile are in the scratch
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += scr[(c0 + i) * DHP + lane];
#pragma unroll
          for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(out + lane * NE + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
      }
      if (C::EPI > 1 && e == 1) named_bar_sync(3, 64);
      tc_fence_before();
      mbar_arrive(&bar_tfree[slot]);
      if (C::EPI > 1) named_bar_sync(4, 64);                         // the scratch is free again
    }
  } else {
    // ===================================================================== consumers: thread = chunk row, group = tile parity
    const int grp = warp >> 2, w = warp & 3, r = threadIdx.x & (kL - 1);
    float x0 = 0.f, x1 = 0.f, x2 = 0.f;
    auto load_gates = [&](int it) {
      const size_t grow = static_cast<size_t>(blockIdx.x + it * gridDim.x) * kL + r;
      x1 = g1[grow];
      if (!REV) {
        x0 = g0[grow];
      } else {
        x0 = m_in[grow];
        x2 = den_in[grow];
      }
    };
    if (grp < n_my) load_gates(grp);
    for (int it = grp; it < n_my; it += 2) {
      const int s = it % NSTAGE;
      const int tile = blockIdx.x + it * gridDim.x;
      float* red = aux + grp * 16 + ((it >> 1) & 1) * 8;
      const float c0 = x0, fv = x1, c2 = x2;
      if (it + 2 < n_my) load_gates(it + 2);
      // x_j = warp-local inclusive cumsum of log sigmoid(f); b_j = x_j + (totals of the warps in front)
      float x = log_sigmoid(fv);
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      // forward: a_j = g - b_j + i_j, weight exp(a_j - max a);  backward: a_t = b_t - m_t, weight exp(a_t - max a).
      // Inside a warp the offset of b is a constant, so the warp maximum of the local part travels with the warp total
      // through ONE exchange:  max a = g + max_w(lmax_w - off_w)  (forward),  max_w(lmax_w + off_w)  (backward)
      const float loc = REV ? x - c0 : c0 - x;
      float lmax = loc;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
      if (lane == 31) red[w] = x;
      if (lane == 0) red[4 + w] = lmax;
      named_bar_sync(1 + grp, kL);
      float off = 0.f, g = 0.f, amax = -INFINITY;
#pragma unroll
      for (int ww = 0; ww < 4; ++ww) {
        const float t = red[ww], lm = red[4 + ww];
        amax = fmaxf(amax, REV ? lm + g : lm - g);                 // g = total of the warps in front of ww at this point
        if (ww < w) off += t;
        g += t;
      }
      if (!REV) amax += g;
      const float a = REV ? loc + off : g - off + loc;
      const float wgt = __expf(a - amax) * scale;
      unsigned char* st = smem + s * C::STAGE;
      mbar_wait(&bar_full[s], (it / NSTAGE) & 1);
      if (REV) state_build_G_row<DHP>(st + C::OFF_B, st + C::OFF_H, r, c0, c2, eps);
      // the scaled rows are kept as a bf16 hi + lo pair (~16 mantissa bits) so that the carried state stays consistent with
      // the intra-chunk products (gate gradients)
      scale_row_hilo<DHP>(st + C::OFF_A, st + C::OFF_ALO, r, wgt);
      if (r == 0) {
        g_out[tile] = g;
        amax_out[tile] = amax;
      }
      fence_proxy_async();
      mbar_arrive(&bar_ready[s]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, C::TMEM_COLS);
}

int sm_count_cached();

template <int DHP, bool REV>
static int launch_state_ws(const void* a, const void* b, const void* h, const float* g0, const float* g1, const float* m, const float* den,
                           int ntiles, float scale, float eps, float* dstate, float* g_out, float* amax_out, cudaStream_t st) {
  using C = StateWs<DHP, REV>;
  cudaError_t e = cudaFuncSetAttribute(mlstm_chunk_state_ws_kernel<DHP, REV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
  if (e != cudaSuccess) return (int)e;
  const int cap = sm_count_cached() * 2;
  const int grid = ntiles < cap ? ntiles : cap;
  ProfScope ps(REV ? K_CHUNK_RSTATE : K_CHUNK_STATE, st);
  mlstm_chunk_state_ws_kernel<DHP, REV><<<grid, C::NTHREADS, C::SMEM, st>>>(
      (const unsigned char*)a, (const unsigned char*)b, (const unsigned char*)h, g0, g1, m, den, ntiles, scale, eps, dstate, g_out, amax_out);
  return (int)cudaGetLastError();
}

// phase 1 of the forward on the persistent kernel; dhp in {16, 32}
int launch_chunk_state_ws(int dhp, const void* k, const void* v, const float* ig, const float* fg, int BH, int nc, float scale,
                          float* dstate, float* g_out, float* amax_out, cudaStream_t st) {
  switch (dhp) {
    case 16: return launch_state_ws<16, false>(k, v, nullptr, ig, fg, nullptr, nullptr, BH * nc, scale, 0.f, dstate, g_out, amax_out, st);
    case 32: return launch_state_ws<32, false>(k, v, nullptr, ig, fg, nullptr, nullptr, BH * nc, scale, 0.f, dstate, g_out, amax_out, st);
    default: return XHVED_ERR_UNSUPPORTED_DH;
  }
}

// phase B1 of the backward on the persistent kernel; dhp in {16, 32}
int launch_chunk_rstate_ws(int dhp, const void* q, const void* dh_t, const void* h, const float* fg, const float* m, const float* den,
                           int BH, int nc, float scale, float eps, float* dstate, float* g_out, float* lam_out, cudaStream_t st) {
  switch (dhp) {
    case 16: return launch_state_ws<16, true>(q, dh_t, h, nullptr, fg, m, den, BH * nc, scale, eps, dstate, g_out, lam_out, st);
    case 32: return launch_state_ws<32, true>(q, dh_t, h, nullptr, fg, m, den, BH * nc, scale, eps, dstate, g_out, lam_out, st);
    default: return XHVED_ERR_UNSUPPORTED_DH;
  }
}

}  // namespace xhved
