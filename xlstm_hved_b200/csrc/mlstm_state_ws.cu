// Chunkwise mLSTM, phases 1 / B1 (chunk_state, chunk_rstate) as ONE persistent, warp-specialised kernel -- sm_100a.
//
// Same mathematics as mlstm_chunk_state_kernel (mlstm_fwd.cu) and mlstm_chunk_rstate_kernel (mlstm_bwd.cu; SURVEY.md 8a-note:
// the contribution of one chunk to the carried state  dC_c = sum_j exp(a_j - amax_c) (k_j/sqrt(DH)) [v_j | 1]^T  of
// vision_lstm.py:48-130 in chunkwise form, and its mirror image  dR_c = sum_t exp(b_t - m_t - lambda_c) (q_t/sqrt(DH)) G_t^T
// of the backward pass).  The one-tile-per-CTA kernels are a single dependent chain per tile -- bulk load -> gate scans -> row
// scaling -> eight tcgen05.mma -> tcgen05.ld -> store -- that six resident CTAs per SM only partly overlap (25 / 26 us for
// 4096 tiles, 18 % / 27 % of HBM), and what bounds a pipelined version is the tensor front-end: a tcgen05.mma costs ~80 cycles
// whatever its shape (tools/bench_umma_issue.py), and the K = 128 contraction of ONE tile is eight of them that use 2 DHP of
// the 128 accumulator rows and NE of up to 256 columns.
//
// Here TPQ = 128 / (2 DHP) tiles ("a quad" at DHP = 16) share one chain of eight instructions.  The row-scaled A operands of
// the quad (bf16 hi + lo pairs, 2 DHP columns per tile) lie side by side and fill the 128-row MN-major window exactly; the
// B operands ([V | 1] or G, NE columns per tile) lie side by side as one N = TPQ NE operand.  The contraction index (the
// token inside the chunk) is common, so block (a, b) of the product is A_a^T B_b: the diagonal blocks are the TPQ results,
// the off-diagonal ones are never read.  Per tile that is 2 instructions instead of 8.  One role per warp:
//
//   warp P    (producer)   streams EVERYTHING the next quads need into a ring of shared-memory stages (bulk copies): operand
//                          tiles, the gate / stabiliser rows (512 B per tile and array) and, backward, the H tile -- into the
//                          slot of the lo half of A, whose rows their owner threads read before they overwrite them.  (With
//                          the gates fetched by the consumers themselves one quad ahead, a quad cost one DRAM latency.)
//   consumers              one group of 128 threads per tile of the quad; thread = chunk row: gate scans (shuffles + ONE
//                          named barrier: cumulative sum and maximum share the exchange), row weights, the row-scaled
//                          A operand as a hi + lo pair (and, backward, the row-extended gradient G in place)
//   warp M    (MMA)        issues the chain of a quad as soon as its stage is ready, into one of two TMEM slots
//   warps E   (epilogue)   one per TMEM quadrant: rows d hold the hi part, rows d + DHP the lo part of the same product
//
// so the scans of quad i+1, the MMAs of quad i and the epilogue of quad i-1 run concurrently; barriers and tensor memory are
// set up once per CTA (one CTA per SM).
#include "mlstm_common.cuh"
#include "prof.cuh"
#include "xhved.h"

namespace xhved {

#ifdef XHVED_TRACE
__device__ long long g_trace[8][32];
#define TRACE(role, it) do { if (blockIdx.x == 0 && (it) < 32 && (threadIdx.x & 31) == 0) g_trace[role][it] = clock64(); } while (0)
#else
#define TRACE(role, it)
#endif

template <int DHP>
struct StateWs {
  static constexpr int NE = ext_cols(DHP);
  static constexpr uint32_t TILE = kL * DHP * 2, EXT = kL * NE * 2;
  static constexpr int TPQ = 128 / (2 * DHP);                        // tiles per MMA chain: 4 (DHP 16), 2 (DHP 32)
  // stage: [A_0 hi | A_0 lo | A_1 hi | ...] = 32 KB = the 128-row MN-major window; [B_0 ext | B_1 ext | ...]
  static constexpr uint32_t A_BYTES = TPQ * 2 * TILE, B_BYTES = TPQ * EXT;
  static constexpr uint32_t G_BYTES = TPQ * 3 * kL * 4;              // per tile: [i | f | -] or [m | f | den] rows, fp32
  static constexpr uint32_t OFF_A = 0, OFF_B = A_BYTES, OFF_G = A_BYTES + B_BYTES, STAGE = A_BYTES + B_BYTES + G_BYTES;
  static constexpr int NSTAGE = 3;
  static constexpr uint32_t OFF_SCR = NSTAGE * STAGE;                // fp32 [TPQ][NE][DHP]: lo rows on their way to the hi warp (DHP = 32)
  static constexpr uint32_t OFF_AUX = OFF_SCR + (DHP > 16 ? TPQ * DHP * NE * 4 : 0);
  static constexpr uint32_t SMEM = OFF_AUX + TPQ * 2 * 8 * 4;        // scan scratch: per consumer group, double-buffered
  static constexpr int W_EPI = 4 * TPQ, W_PROD = W_EPI + 4, W_MMA = W_EPI + 5;
  static constexpr int NTHREADS = (W_EPI + 6) * 32;
  static constexpr uint32_t SLOT = TPQ * NE;                         // TMEM columns of one accumulator slot (128 / 96)
  static constexpr uint32_t TMEM_COLS = 256;
  static_assert(A_BYTES == 32768, "the A operands of a quad fill the 128-row window");
  static_assert(DHP == 16 || DHP == 32, "2 DHP rows per tile, whole TMEM quadrants");
  static_assert(SMEM <= 227 * 1024, "stage ring does not fit");
};

// Row-extended gradient G_t = [dh_t / N_t | db_t | 0] built in place over the dH tile (see build_G_row in mlstm_bwd.cu);
// the row of H comes in registers
template <int DHP>
__device__ __forceinline__ void state_build_G_row(unsigned char* sG, const uint4* hrow, int t, float m, float den, float eps) {
  const float flo = __expf(-m);
  const float r = 1.f / (fmaxf(fabsf(den), flo) + eps);
  float dhh = 0.f;
#pragma unroll
  for (int cg = 0; cg < DHP / 8; ++cg) {
    uint4* pg = reinterpret_cast<uint4*>(sG + tile_off16(kL, t, cg));
    float g[8], h[8];
    unpack8_bf16(*pg, g);
    unpack8_bf16(hrow[cg], h);
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
__global__ void __launch_bounds__(StateWs<DHP>::NTHREADS, 1) mlstm_chunk_state_ws_kernel(
    const unsigned char* __restrict__ a_tiles, const unsigned char* __restrict__ b_tiles, const unsigned char* __restrict__ h_tiles,
    const float* __restrict__ g0, const float* __restrict__ g1, const float* __restrict__ m_in, const float* __restrict__ den_in,
    int ntiles, float scale, float eps, float* __restrict__ dstate, float* __restrict__ g_out, float* __restrict__ amax_out) {
  using C = StateWs<DHP>;
  constexpr int NE = C::NE, NSTAGE = C::NSTAGE, TPQ = C::TPQ;
  constexpr uint32_t TILE = C::TILE;
  extern __shared__ __align__(128) unsigned char smem[];
  float* aux = reinterpret_cast<float*>(smem + C::OFF_AUX);
  // full: operands landed (tx) | ready: consumers done with the stage (128 TPQ) | sfree: the MMAs have read the stage (commit)
  // mdone: accumulator slot complete (commit) | tfree: slot drained (four epilogue warps)
  __shared__ __align__(8) uint64_t bar_full[NSTAGE], bar_ready[NSTAGE], bar_sfree[NSTAGE], bar_mdone[2], bar_tfree[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nquads = (ntiles + TPQ - 1) / TPQ;
  const int n_my = (nquads - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_ready[s], kL * TPQ);
      mbar_init(&bar_sfree[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_mdone[s], 1);
      mbar_init(&bar_tfree[s], 4 * 32);
    }
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, C::TMEM_COLS);
  if (!REV) {
    // constant ext columns [1 | 0] behind every V tile of every stage (bulk loads only ever overwrite the first DHP columns)
    for (int i = threadIdx.x; i < NSTAGE * TPQ * kL; i += blockDim.x)
      write_ext_ones(smem + (i / (TPQ * kL)) * C::STAGE + C::OFF_B + ((i / kL) % TPQ) * C::EXT, DHP, i % kL);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) TRACE(7, 0);

  if (warp == C::W_PROD) {
    // ===================================================================== producer: operand tiles through the stage ring
    for (int it = 0; it < n_my; ++it) {
      const int s = it % NSTAGE;
      const int tile0 = (blockIdx.x + it * gridDim.x) * TPQ;
      const int nt = ntiles - tile0 < TPQ ? ntiles - tile0 : TPQ;          // tiles present in this quad
      if (it >= NSTAGE) mbar_wait_spin(&bar_sfree[s], (it / NSTAGE - 1) & 1);
      TRACE(0, it);
      unsigned char* st = smem + s * C::STAGE;
      mbar_expect_tx_e(&bar_full[s], nt * ((REV ? 3 : 2) * TILE + (REV ? 3 : 2) * kL * 4));
      for (int a = 0; a < nt; ++a) {
        const size_t to = static_cast<size_t>(tile0 + a) * TILE, go = static_cast<size_t>(tile0 + a) * kL;
        unsigned char* gs = st + C::OFF_G + a * (3 * kL * 4);
        bulk_g2s_e(st + C::OFF_A + a * 2 * TILE, a_tiles + to, TILE, &bar_full[s]);
        bulk_g2s_e(st + C::OFF_B + a * C::EXT, b_tiles + to, TILE, &bar_full[s]);
        bulk_g2s_e(gs, (REV ? m_in : g0) + go, kL * 4, &bar_full[s]);
        bulk_g2s_e(gs + kL * 4, g1 + go, kL * 4, &bar_full[s]);
        if (REV) {
          bulk_g2s_e(gs + 2 * kL * 4, den_in + go, kL * 4, &bar_full[s]);
          bulk_g2s_e(st + C::OFF_A + a * 2 * TILE + TILE, h_tiles + to, TILE, &bar_full[s]);      // H rides in the lo slot
        }
      }
    }
  } else if (warp == C::W_MMA) {
    // ===================================================================== MMA issuer (whole warp, elect forms)
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    for (int it = 0; it < n_my; ++it) {
      const int s = it % NSTAGE, slot = it & 1;
      const uint32_t st = smem_u32(smem + s * C::STAGE);
      mbar_wait_spin(&bar_ready[s], (it / NSTAGE) & 1);
      TRACE(1, it);
      if (it >= 2) mbar_wait_spin(&bar_tfree[slot], ((it >> 1) - 1) & 1);
      tc_fence_after();
      // D[(a, d)][(b, e')] = sum_j A~_a[j][d] * B_b[j][e']   (A, B both MN-major views of row-j tiles); block a = b is tile a
      umma_gemm_e(tmem_u + slot * C::SLOT, st + C::OFF_A, 128, kL * 16, st + C::OFF_B, 128, kL * 16,
                  umma_idesc(128, TPQ * NE, true, true), kL, false);
      umma_commit_e(&bar_sfree[s]);
      umma_commit_e(&bar_mdone[slot]);
      TRACE(2, it);
    }
  } else if (warp >= C::W_EPI) {
    // ===================================================================== epilogue: one warp per TMEM quadrant
    const int e = warp - C::W_EPI;
    const uint32_t lane_base = static_cast<uint32_t>(e * 32) << 16;
    const int a = DHP == 16 ? e : (e >> 1);                        // tile of the quad this quadrant belongs to
    const bool lo_warp = DHP > 16 && (e & 1);                      // DHP = 32: quadrant 2a = hi rows, 2a + 1 = lo rows
    float* scr = reinterpret_cast<float*>(smem + C::OFF_SCR) + a * DHP * NE;
    for (int it = 0; it < n_my; ++it) {
      const int slot = it & 1;
      const int tile = (blockIdx.x + it * gridDim.x) * TPQ + a;
      float* out = dstate + static_cast<size_t>(tile) * DHP * NE;
      mbar_wait_spin(&bar_mdone[slot], (it >> 1) & 1);
      if (e == 0) TRACE(3, it);
      tc_fence_after();
      if (tile < ntiles) {
#pragma unroll
        for (int c0 = 0; c0 < NE; c0 += 16) {
          float v[16];
          tmem_ld16(tmem + lane_base + slot * C::SLOT + a * NE + c0, v);
          if (DHP == 16) {
            // lanes d hold the hi rows, lanes d + 16 the lo rows
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += __shfl_down_sync(0xffffffffu, v[i], 16);
            if (lane < 16) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(out + lane * NE + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
          } else if (lo_warp) {
#pragma unroll
            for (int i = 0; i < 16; ++i) scr[(c0 + i) * DHP + lane] = v[i];
          } else {
            if (c0 == 0) named_bar_sync(5 + a, 64);                  // the lo rows of this tile are in the scratch
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += scr[(c0 + i) * DHP + lane];
#pragma unroll
            for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(out + lane * NE + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          }
        }
        if (lo_warp) named_bar_sync(5 + a, 64);
        if (DHP > 16) named_bar_sync(7 + a, 64);                     // the scratch is free again
      }
      tc_fence_before();
      mbar_arrive(&bar_tfree[slot]);
      if (e == 0) TRACE(4, it);
    }
  } else {
    // ===================================================================== consumers: group = tile of the quad, thread = chunk row
    const int grp = warp >> 2, w = warp & 3, r = threadIdx.x & (kL - 1);
    for (int it = 0; it < n_my; ++it) {
      const int s = it % NSTAGE;
      const int tile = (blockIdx.x + it * gridDim.x) * TPQ + grp;
      const bool on = tile < ntiles;                                 // uniform over the group
      float* red = aux + grp * 16 + (it & 1) * 8;
      unsigned char* st = smem + s * C::STAGE;
      unsigned char* sA = st + C::OFF_A + grp * 2 * TILE;
      mbar_wait_spin(&bar_full[s], (it / NSTAGE) & 1);
      if (warp == 0) TRACE(5, it);
      if (on) {
        const float* gs = reinterpret_cast<const float*>(st + C::OFF_G + grp * (3 * kL * 4));
        const float c0 = gs[r], fv = gs[kL + r];
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
          amax = fmaxf(amax, REV ? lm + g : lm - g);               // g = total of the warps in front of ww at this point
          if (ww < w) off += t;
          g += t;
        }
        if (!REV) amax += g;
        const float a = REV ? loc + off : g - off + loc;
        const float wgt = __expf(a - amax) * scale;
        if (REV) {
          // this row of H sits where the lo half of the scaled row goes (only this thread touches either)
          uint4 hrow[DHP / 8];
#pragma unroll
          for (int cg = 0; cg < DHP / 8; ++cg) hrow[cg] = *reinterpret_cast<const uint4*>(sA + TILE + tile_off16(kL, r, cg));
          state_build_G_row<DHP>(st + C::OFF_B + grp * C::EXT, hrow, r, c0, gs[2 * kL + r], eps);
        }
        // the scaled rows are kept as a bf16 hi + lo pair (~16 mantissa bits) so that the carried state stays consistent with
        // the intra-chunk products (gate gradients)
        scale_row_hilo<DHP>(sA, sA + TILE, r, wgt);
        if (r == 0) {
          g_out[tile] = g;
          amax_out[tile] = amax;
        }
      }
      fence_proxy_async();
      mbar_arrive(&bar_ready[s]);
      if (warp == 0) TRACE(6, it);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) TRACE(7, 1);
  if (warp == 0) tmem_dealloc(tmem, C::TMEM_COLS);
}

int sm_count_cached();

template <int DHP, bool REV>
static int launch_state_ws(const void* a, const void* b, const void* h, const float* g0, const float* g1, const float* m, const float* den,
                           int ntiles, float scale, float eps, float* dstate, float* g_out, float* amax_out, cudaStream_t st) {
  using C = StateWs<DHP>;
  cudaError_t e = cudaFuncSetAttribute(mlstm_chunk_state_ws_kernel<DHP, REV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
  if (e != cudaSuccess) return (int)e;
  const int nquads = (ntiles + C::TPQ - 1) / C::TPQ;
  const int grid = nquads < sm_count_cached() ? nquads : sm_count_cached();
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

#ifdef XHVED_TRACE
extern "C" int xhved_debug_trace(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, xhved::g_trace, sizeof(long long) * 8 * 32);
}
#endif
