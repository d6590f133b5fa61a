// Chunkwise stabilised mLSTM cell, forward -- sm_100a (tcgen05 + TMEM + bulk async copies).
//
// Computes exactly what the reference's parallel_stabilized_simple does
// (UxLSTM/nnunetv2/nets/vision_lstm.py:48-130) in the chunkwise form of
// SURVEY.md 8a-note, chunk L = 128 = UMMA M:
//
//   phase 1  chunk_state   (b,head,chunk)-parallel: dC_c = sum_j exp(a_j - amax_c) (k_j/sqrt(DH)) [v_j | 1]^T
//   phase 2  state_scan    (b,head)-parallel, sequential over chunks: carried (C|n, m) entering every chunk
//   phase 3  chunk_out     (b,head,chunk)-parallel: S = Q K^T, P = S o D', O = P V + w (Q [C|n]),  h = O / N
//
// q/k/v/h tiles are bf16 in the tile-native layout (umma.cuh); gates, stabiliser m,
// normaliser input den and all accumulators are fp32.
#include <stdlib.h>

#include "mlstm_common.cuh"
#include "prof.cuh"
#include "xhved.h"

namespace xhved {

// ------------------------------------------------------------------ phase 1
template <int DHP>
__global__ void __launch_bounds__(kThreads) mlstm_chunk_state_kernel(const unsigned char* __restrict__ k_tiles,
                                                                      const unsigned char* __restrict__ v_tiles,
                                                                      const float* __restrict__ ig, const float* __restrict__ fg,
                                                                      int nc, float scale, float* __restrict__ dstate,
                                                                      float* __restrict__ g_out, float* __restrict__ amax_out) {
  constexpr int NE = ext_cols(DHP);
  constexpr uint32_t TILE = kL * DHP * 2;
  constexpr uint32_t TMEM_COLS = next_pow2_cols(NE);
  // smem: [K~ hi][K~ lo][Vext tile].  K~ hi / lo are read as 128-row MN-major A operands, i.e. through a 32 KB window each;
  // rows >= DHP of the product are never read, so the windows run on over the following tiles instead of owning padding
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sK = smem;
  unsigned char* sKlo = smem + TILE;
  unsigned char* sV = smem + 2 * TILE;
  __shared__ __align__(8) uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_slot;
  __shared__ float red[8];

  const int tid = threadIdx.x, warp = tid >> 5;
  const int tile = blockIdx.x;  // (b*NH + h) * nc + c
  const size_t grow = static_cast<size_t>(tile) * kL + tid;

  if (tid == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, TMEM_COLS);
  write_ext_ones(sV, DHP, tid);
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar_load, 2 * TILE);
    bulk_g2s(sK, k_tiles + static_cast<size_t>(tile) * TILE, TILE, &bar_load);
    bulk_g2s(sV, v_tiles + static_cast<size_t>(tile) * TILE, TILE, &bar_load);
  }
  // gates: b_j = chunk-local inclusive cumsum of log sigmoid(f), a_j = g - b_j + i_j
  const float iv = ig[grow];
  const float lf = log_sigmoid(fg[grow]);
  float g;
  const float b = block_cumsum128(lf, red, &g);
  const float a = g - b + iv;
  float amax;
  block_cummax128(a, red, &amax);
  const float wgt = __expf(a - amax) * scale;

  mbar_wait(&bar_load, 0);
  // scale K rows: k~_j = exp(a_j - amax) * k_j / sqrt(DH), kept as a bf16 hi + lo pair (~16 mantissa bits) so
  // that the carried state stays consistent with the intra-chunk products (gate gradients)
  scale_row_hilo<DHP>(sK, sKlo, tid, wgt);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  // D[d][e'] = sum_j K~[j][d] * Vext[j][e']   (A, B both MN-major views of row-j tiles).  The lo tile lies directly behind the
  // hi tile, i.e. inside the 128-row MN window of the hi operand: rows [DHP, 2 DHP) of the SAME product are lo^T Vext, so
  // for DHP <= 64 one pass of K = 128 yields both halves (a tcgen05.mma costs ~80 cycles whatever its size, measured with
  // tools/bench_umma_issue.py: the instruction count is what bounds this kernel) and the epilogue adds the two row groups.
  constexpr bool ONE_PASS = DHP <= 64;
  if (tid == 0) {
    umma_gemm(tmem, smem_u32(sK), /*lbo*/ 128, /*sbo*/ kL * 16, smem_u32(sV), 128, kL * 16,
              umma_idesc(128, NE, true, true), kL, false);
    if (!ONE_PASS)
      umma_gemm(tmem, smem_u32(sKlo), 128, kL * 16, smem_u32(sV), 128, kL * 16, umma_idesc(128, NE, true, true), kL, true);
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  store_state_rows<DHP, ONE_PASS>(tmem, dstate + static_cast<size_t>(tile) * DHP * NE, reinterpret_cast<float*>(smem));
  if (tid == 0) {
    g_out[tile] = g;
    amax_out[tile] = amax;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

// ------------------------------------------------------------------ phase 2
// One (b,head) sequence per blockIdx.x; a thread owns one 16-byte group (8 consecutive columns of one row) of the DHP x NE
// state, so every chunk costs it two 16-byte loads and two 16-byte stores (a warp covers 512 contiguous bytes of a tile).
// Writes, for every chunk c, the state ENTERING chunk c as a PAIR of bf16 tile-native tiles (hi, lo = residual)
// [DHP rows (key dim d)][NE cols (value dim e | n | 0)] plus its log-scale m_prev[c].
// `reverse` runs the same recurrence from the last chunk to the first (backward pass).
// Sequences of 64 chunks and more take mlstm_state_scan_par_kernel below.
template <int DHP>
__global__ void __launch_bounds__(256) mlstm_state_scan_kernel(const float* __restrict__ dstate, const float* __restrict__ g_in,
                                                                const float* __restrict__ amax_in, int nc, int reverse,
                                                                unsigned char* __restrict__ states, float* __restrict__ m_prev) {
  // grid = (B*NH, ceil(groups / blockDim)).  The scalar log-scale recurrence m' = max(g + m, a) is a scan over the maps
  // m -> max(G + m, A), which compose as (G1, A1) o (G2, A2) = (G1 + G2, max(A1 + G2, A2)): warp 0 runs it 32 chunks at a time
  // (shuffles) and leaves the per-chunk coefficients in shared memory; after that every state element follows
  // acc <- decay_c * acc + w_c * dstate_c, with the chunk contributions loaded eight chunks at a time so that the serial chain
  // sees ~nc/8 memory latencies instead of nc.
  constexpr int NE = ext_cols(DHP);
  constexpr int NEL = DHP * NE, NG = NEL / 8, CGS = NE / 8;
  extern __shared__ float scan_smem[];
  float* s_dec = scan_smem;          // [nc] in scan order
  float* s_w = s_dec + nc;
  const int bh = blockIdx.x, tid = threadIdx.x;
  const size_t tile0 = static_cast<size_t>(bh) * nc;
  if (tid < 32) {
    float cG = 0.f, cA = -INFINITY;                      // composition of all chunks in front of this block of 32
    for (int s0 = 0; s0 < nc; s0 += 32) {
      const int st = s0 + tid;
      const bool on = st < nc;
      const int c = reverse ? nc - 1 - st : st;
      const float g = on ? g_in[tile0 + c] : 0.f, a = on ? amax_in[tile0 + c] : -INFINITY;      // off: the identity map
      float G = g, A = a;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float Gp = __shfl_up_sync(0xffffffffu, G, o), Ap = __shfl_up_sync(0xffffffffu, A, o);
        if (tid >= o) {
          A = fmaxf(Ap + G, A);
          G += Gp;
        }
      }
      const float m_new = fmaxf(cA + G, A);               // log-scale after this chunk
      float m_in = __shfl_up_sync(0xffffffffu, m_new, 1);  // log-scale of the state ENTERING the chunk
      if (tid == 0) m_in = cA;
      if (on) {
        s_dec[st] = __expf(g + m_in - m_new);             // exp(-inf) = 0 on the first step
        s_w[st] = __expf(a - m_new);
        if (blockIdx.y == 0) m_prev[tile0 + c] = m_in;
      }
      cG += __shfl_sync(0xffffffffu, G, 31);
      cA = __shfl_sync(0xffffffffu, m_new, 31);
    }
  }
  __syncthreads();
  const int gi = blockIdx.y * blockDim.x + tid;
  if (gi >= NG) return;
  const int d = gi / CGS, cg = gi % CGS;
  const float* in = dstate + tile0 * NEL + d * NE + cg * 8;
  unsigned char* out = states + tile0 * (NEL * 4) + tile_off16(DHP, d, cg);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  constexpr int PF = 8;
  for (int s0 = 0; s0 < nc; s0 += PF) {
    float4 cur[PF][2];
#pragma unroll
    for (int j = 0; j < PF; ++j) {
      const int st = s0 + j;
      const int c = st < nc ? (reverse ? nc - 1 - st : st) : 0;
      const float4* sn = reinterpret_cast<const float4*>(in + static_cast<size_t>(c) * NEL);
      cur[j][0] = __ldg(sn);
      cur[j][1] = __ldg(sn + 1);
    }
#pragma unroll
    for (int j = 0; j < PF; ++j) {
      const int st = s0 + j;
      if (st < nc) {
        const int c = reverse ? nc - 1 - st : st;
        // emit the state entering this chunk as a bf16 hi/lo pair, then fold the chunk in
        uint4 hi, lo;
        split8_hilo(acc, hi, lo);
        unsigned char* o = out + static_cast<size_t>(c) * (NEL * 4);
        *reinterpret_cast<uint4*>(o) = hi;
        *reinterpret_cast<uint4*>(o + NEL * 2) = lo;
        const float dec = s_dec[st], w = s_w[st];
        const float v[8] = {cur[j][0].x, cur[j][0].y, cur[j][0].z, cur[j][0].w, cur[j][1].x, cur[j][1].y, cur[j][1].z, cur[j][1].w};
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = dec * acc[i] + w * v[i];
      }
    }
  }
}

// The same two recurrences for LONG sequences (nc >= 64: the 32^3 stage of SURVEY 8d config 5 has 256 chunks and few
// sequences, where the serial chain above is nc/8 memory latencies long and leaves most SMs idle).  Both are scans of affine
// maps, so a WARP runs them with lane = chunk:
//   log-scale:  m' = max(g + m, a)            maps m -> max(G + m, A) compose as (G1 + G2, max(A1 + G2, A2))
//   state:      acc' = dec * acc + w * dC     maps acc -> D acc + V  compose as (D1 D2, V1 D2 + V2)
// One warp owns one 16-byte group of the state: every lane loads the contribution of ITS chunk (one round of memory latency
// per 32 chunks), five shuffle steps give the inclusive scans, every lane writes the state entering its chunk; blocks of 32
// chunks follow each other with a carry.  (At nc = 32 the 16-byte stores 4 KB apart make it slower than the kernel above:
// 14.2 vs 10.9 us per launch, so it only runs for long sequences.)
template <int DHP>
__global__ void __launch_bounds__(256) mlstm_state_scan_par_kernel(const float* __restrict__ dstate, const float* __restrict__ g_in,
                                                                    const float* __restrict__ amax_in, int nc, int reverse,
                                                                    unsigned char* __restrict__ states, float* __restrict__ m_prev) {
  constexpr int NE = ext_cols(DHP);
  constexpr int NEL = DHP * NE, NG = NEL / 8, CGS = NE / 8;
  const int bh = blockIdx.x, lane = threadIdx.x & 31;
  const int gi = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);      // 16-byte group of this warp
  if (gi >= NG) return;
  const size_t tile0 = static_cast<size_t>(bh) * nc;
  const int d = gi / CGS, cg = gi % CGS;
  const float* in = dstate + tile0 * NEL + d * NE + cg * 8;
  unsigned char* out = states + tile0 * (NEL * 4) + tile_off16(DHP, d, cg);
  float cG = 0.f, cA = -INFINITY;                        // log-scale map of all chunks in front of this block of 32
  float cV[8];                                           // state after all chunks in front of this block
#pragma unroll
  for (int i = 0; i < 8; ++i) cV[i] = 0.f;
  for (int s0 = 0; s0 < nc; s0 += 32) {
    const int st = s0 + lane;
    const bool on = st < nc;
    const int c = on ? (reverse ? nc - 1 - st : st) : 0;
    const float4* sn = reinterpret_cast<const float4*>(in + static_cast<size_t>(c) * NEL);
    const float4 x0 = __ldg(sn), x1 = __ldg(sn + 1);
    const float g = on ? g_in[tile0 + c] : 0.f, a = on ? amax_in[tile0 + c] : -INFINITY;      // off: the identity map
    float G = g, A = a;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float Gp = __shfl_up_sync(0xffffffffu, G, o), Ap = __shfl_up_sync(0xffffffffu, A, o);
      if (lane >= o) {
        A = fmaxf(Ap + G, A);
        G += Gp;
      }
    }
    const float m_new = fmaxf(cA + G, A);                // log-scale after this chunk
    float m_in = __shfl_up_sync(0xffffffffu, m_new, 1);  // log-scale of the state ENTERING the chunk
    if (lane == 0) m_in = cA;
    float D = on ? __expf(g + m_in - m_new) : 1.f;       // exp(-inf) = 0 on the first step
    const float w = on ? __expf(a - m_new) : 0.f;
    float V[8] = {w * x0.x, w * x0.y, w * x0.z, w * x0.w, w * x1.x, w * x1.y, w * x1.z, w * x1.w};
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float Dp = __shfl_up_sync(0xffffffffu, D, o);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float Vp = __shfl_up_sync(0xffffffffu, V[i], o);
        if (lane >= o) V[i] = Vp * D + V[i];
      }
      if (lane >= o) D *= Dp;
    }
    float e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      V[i] = D * cV[i] + V[i];                            // state after this chunk
      e[i] = __shfl_up_sync(0xffffffffu, V[i], 1);        // ... and entering it
      if (lane == 0) e[i] = cV[i];
      cV[i] = __shfl_sync(0xffffffffu, V[i], 31);
    }
    if (on) {
      uint4 hi, lo;
      split8_hilo(e, hi, lo);
      unsigned char* o = out + static_cast<size_t>(c) * (NEL * 4);
      *reinterpret_cast<uint4*>(o) = hi;
      *reinterpret_cast<uint4*>(o + NEL * 2) = lo;
      if (gi == 0) m_prev[tile0 + c] = m_in;
    }
    cG += __shfl_sync(0xffffffffu, G, 31);
    cA = __shfl_sync(0xffffffffu, m_new, 31);
  }
}

// ------------------------------------------------------------------ phase 3
// p = S * exp2(u_row + v_col) for 16 columns of one row, bf16 P written to the row's two 16-byte groups (2 KB apart).
template <bool MASK>
__device__ __forceinline__ float decay_half(const float* sv, float urow, const float* vcol, int s0, int row, unsigned char* dst) {
  float rowsum = 0.f;
#pragma unroll
  for (int j8 = 0; j8 < 2; ++j8) {
    float p[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = fast_exp2(urow + vcol[j8 * 8 + j]);
      p[j] = sv[j8 * 8 + j] * d;
      if (MASK) p[j] = (s0 + j8 * 8 + j <= row) ? p[j] : 0.f;
      rowsum += p[j];
    }
    *reinterpret_cast<uint4*>(dst + j8 * (kL * 16)) = pack8_bf16(p);
  }
  return rowsum;
}

// 256 threads: thread = (row, half); the causal S -> P conversion of a row and (DHP >= 32) the output columns are split
// between the two halves, both of which know the gate quantities of their row.
template <int DHP>
__global__ void __launch_bounds__(2 * kThreads) mlstm_chunk_out_kernel(
    const unsigned char* __restrict__ q_tiles, const unsigned char* __restrict__ k_tiles, const unsigned char* __restrict__ v_tiles,
    const float* __restrict__ ig, const float* __restrict__ fg, const unsigned char* __restrict__ states,
    const float* __restrict__ m_prev, int nc, float scale, float eps, unsigned char* __restrict__ h_tiles,
    float* __restrict__ m_out, float* __restrict__ den_out) {
  constexpr int NE = ext_cols(DHP);
  constexpr uint32_t TILE = kL * DHP * 2;
  constexpr uint32_t ST_BYTES = 2 * DHP * NE * 2;   // hi + lo tiles
  constexpr uint32_t P_BYTES = kL * kL * 2;
  // TMEM columns: S at [0,128); afterwards O_intra at [0,DHP), O_inter at [DHP, DHP+NE)
  constexpr uint32_t TMEM_COLS = next_pow2_cols((2 * DHP + 16) > 128 ? (2 * DHP + 16) : 128);
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sQ = smem;
  unsigned char* sK = sQ + TILE;
  unsigned char* sV = sK + TILE;
  unsigned char* sP = sV + TILE;
  unsigned char* sS = sP + P_BYTES;
  float* vcol = reinterpret_cast<float*>(sS + ST_BYTES);  // (i_s - b_s) * log2e
  __shared__ __align__(8) uint64_t bar_load, bar_mma1, bar_mma2;
  __shared__ uint32_t tmem_slot;
  __shared__ float red[8];
  __shared__ float rs_part[2][kL];

  const int tid = threadIdx.x & (kL - 1), hsel = threadIdx.x >> 7, warp = tid >> 5;    // tid = row, warp = row group
  const bool lead = threadIdx.x == 0;
  const int tile = blockIdx.x;
  const int c = tile % nc;
  const size_t grow = static_cast<size_t>(tile) * kL + tid;
  const bool has_state = c > 0;

  if (lead) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma1, 1);
    mbar_init(&bar_mma2, 1);
    mbar_fence_init();
  }
  if (threadIdx.x < 32) tmem_alloc(&tmem_slot, TMEM_COLS);
  __syncthreads();
  if (lead) {
    mbar_expect_tx(&bar_load, 3 * TILE + (has_state ? ST_BYTES : 0));
    bulk_g2s(sQ, q_tiles + static_cast<size_t>(tile) * TILE, TILE, &bar_load);
    bulk_g2s(sK, k_tiles + static_cast<size_t>(tile) * TILE, TILE, &bar_load);
    bulk_g2s(sV, v_tiles + static_cast<size_t>(tile) * TILE, TILE, &bar_load);
    if (has_state) bulk_g2s(sS, states + static_cast<size_t>(tile) * ST_BYTES, ST_BYTES, &bar_load);
  }
  // ---- gate scans: b_t, m_t (vision_lstm.py:82-111 as a 1-D scan) ----
  const float iv = ig[grow];
  const float lf = log_sigmoid(fg[grow]);
  const float b = block_cumsum128(lf, red, nullptr);
  const float vc = iv - b;
  const float m_intra = b + block_cummax128(vc, red, nullptr);
  const float mp = has_state ? m_prev[tile] : -INFINITY;
  const float m_inter = b + mp;
  const float m = fmaxf(m_intra, m_inter);
  const float w = has_state ? __expf(m_inter - m) : 0.f;
  vcol[tid] = vc * kLog2e;
  const float urow = (b - m) * kLog2e + log2f(scale);

  mbar_wait(&bar_load, 0);
  tc_fence_before();
  __syncthreads();   // vcol + tmem_slot visible
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (lead) {
    // S[t][s] = sum_d Q[t][d] K[s][d]
    umma_gemm(tmem, smem_u32(sQ), kL * 16, 128, smem_u32(sK), kL * 16, 128, umma_idesc(128, kL, false, false), DHP, false);
    umma_commit(&bar_mma1);
  }
  mbar_wait(&bar_mma1, 0);
  tc_fence_after();
  // ---- P = S o D' (causal), row sums; P -> smem as the next A operand ----
  float rowsum = 0.f;
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
  // row group `warp` needs column blocks 0..warp, i.e. 2 * (warp + 1) half-blocks of 16 columns: warp + 1 per half
#pragma unroll 1
  for (int hb = hsel * (warp + 1); hb < (hsel + 1) * (warp + 1); ++hb) {
    const int blk = hb >> 1, s0 = hb * 16;
    float sv[16];
    tmem_ld16(tmem + lane_base + s0, sv);
    // only the diagonal 32x32 block needs the causal mask; blocks left of it are fully visible
    if (blk < warp)
      rowsum += decay_half<false>(sv, urow, vcol + s0, s0, tid, sP + tile_off16(kL, tid, s0 / 8));
    else
      rowsum += decay_half<true>(sv, urow, vcol + s0, s0, tid, sP + tile_off16(kL, tid, s0 / 8));
  }
  for (int cg = (warp + 1) * 4 + hsel; cg < 16; cg += 2) *reinterpret_cast<uint4*>(sP + tile_off16(kL, tid, cg)) = make_uint4(0, 0, 0, 0);
  rs_part[hsel][tid] = rowsum;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  rowsum = rs_part[0][tid] + rs_part[1][tid];
  if (lead) {
    // O_intra[t][e] = sum_s P[t][s] V[s][e]      (B = MN-major view of V)
    umma_gemm(tmem, smem_u32(sP), kL * 16, 128, smem_u32(sV), 128, kL * 16, umma_idesc(128, DHP, false, true), kL, false);
    // O_inter[t][e'] = sum_d Q[t][d] [C|n][d][e'] (B = MN-major view of the state tile)
    if (has_state) {
      umma_gemm(tmem + DHP, smem_u32(sQ), kL * 16, 128, smem_u32(sS), 128, DHP * 16, umma_idesc(128, NE, false, true), DHP, false);
      umma_gemm(tmem + DHP, smem_u32(sQ), kL * 16, 128, smem_u32(sS) + ST_BYTES / 2, 128, DHP * 16, umma_idesc(128, NE, false, true),
                DHP, true);
    }
    umma_commit(&bar_mma2);
  }
  mbar_wait(&bar_mma2, 0);
  tc_fence_after();
  // ---- epilogue: h = (O_intra + w O_inter) / (max(|den|, exp(-m)) + eps)   (vision_lstm.py:123-128) ----
  float den = rowsum;
  float inter_n = 0.f;
  if (has_state) {
    float t16[16];
    tmem_ld16(tmem + lane_base + 2 * DHP, t16);
    inter_n = t16[0];
    den += w * inter_n;
  }
  const float nrm = fmaxf(fabsf(den), __expf(-m)) + eps;
  const float rn = 1.f / nrm;
  unsigned char* hdst = h_tiles + static_cast<size_t>(tile) * TILE;
#pragma unroll
  for (int c0 = hsel * 16; c0 < DHP; c0 += 32) {
    float o[16];
    tmem_ld16(tmem + lane_base + c0, o);
    if (has_state) {
      float oi[16];
      tmem_ld16(tmem + lane_base + DHP + c0, oi);
#pragma unroll
      for (int i = 0; i < 16; ++i) o[i] += w * oi[i];
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] *= rn;
    uint4 u0, u1;
    u0.x = pack_bf16x2(o[0], o[1]);
    u0.y = pack_bf16x2(o[2], o[3]);
    u0.z = pack_bf16x2(o[4], o[5]);
    u0.w = pack_bf16x2(o[6], o[7]);
    u1.x = pack_bf16x2(o[8], o[9]);
    u1.y = pack_bf16x2(o[10], o[11]);
    u1.z = pack_bf16x2(o[12], o[13]);
    u1.w = pack_bf16x2(o[14], o[15]);
    *reinterpret_cast<uint4*>(hdst + tile_off16(kL, tid, c0 / 8)) = u0;
    *reinterpret_cast<uint4*>(hdst + tile_off16(kL, tid, c0 / 8 + 1)) = u1;
  }
  if (hsel == 0) {
    m_out[grow] = m;
    den_out[grow] = den;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, TMEM_COLS);
}

// ------------------------------------------------------------------ pack / unpack (standalone cell API)
// fp32 (BH, S, DH) row-major -> bf16 tile-native tiles (BH*nc tiles of 128 x DHP), zero padded.  The 128 x DH block of a tile
// is contiguous in the source and the tile is contiguous in the destination: both sides are moved with coalesced 16-byte
// accesses, the layout change happens in shared memory.
__global__ void __launch_bounds__(256) mlstm_pack_kernel(const float* __restrict__ src, int S, int DH, int DHP, int nc,
                                                         unsigned char* __restrict__ tiles) {
  extern __shared__ __align__(16) unsigned char pk_smem[];          // one tile: 128 * DHP * 2 bytes
  const int tile = blockIdx.x, tid = threadIdx.x;
  const int bh = tile / nc, c = tile % nc;
  const int t0 = c * kL, rows = min(kL, S - t0);
  const uint32_t tile_bytes = kL * DHP * 2;
  for (uint32_t i = tid; i < tile_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(pk_smem)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  const float* blk = src + (static_cast<size_t>(bh) * S + t0) * DH;
  const int n = rows * DH;
  if ((DH & 3) == 0 && (reinterpret_cast<uintptr_t>(blk) & 15) == 0) {
    for (int i = tid * 4; i < n; i += blockDim.x * 4) {
      const float4 f = __ldg(reinterpret_cast<const float4*>(blk + i));
      const int r = i / DH, d = i % DH;                              // DH % 4 == 0: the four values share a row
      unsigned char* dst = pk_smem + tile_off16(kL, r, d / 8) + (d % 8) * 2;
      *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(f.x, f.y);
      *reinterpret_cast<uint32_t*>(dst + 4) = pack_bf16x2(f.z, f.w);
    }
  } else {
    for (int i = tid; i < n; i += blockDim.x) {
      const int r = i / DH, d = i % DH;
      *reinterpret_cast<__nv_bfloat16*>(pk_smem + tile_off16(kL, r, d / 8) + (d % 8) * 2) = __float2bfloat16(__ldg(blk + i));
    }
  }
  __syncthreads();
  uint4* out = reinterpret_cast<uint4*>(tiles + static_cast<size_t>(tile) * tile_bytes);
  for (uint32_t i = tid; i < tile_bytes / 16; i += blockDim.x) out[i] = reinterpret_cast<const uint4*>(pk_smem)[i];
}
// gates (BH, S) -> (BH, nc*128); padding rows get i = -1e30 (no weight), f = +1e30 (log sigmoid = 0)
__global__ void mlstm_pack_gates_kernel(const float* __restrict__ ig, const float* __restrict__ fg, int S, int nc,
                                        float* __restrict__ igp, float* __restrict__ fgp) {
  const int tile = blockIdx.x, r = threadIdx.x;
  const int bh = tile / nc, c = tile % nc;
  const int t = c * kL + r;
  const size_t o = static_cast<size_t>(tile) * kL + r;
  igp[o] = t < S ? ig[static_cast<size_t>(bh) * S + t] : -1e30f;
  fgp[o] = t < S ? fg[static_cast<size_t>(bh) * S + t] : 1e30f;
}
// bf16 tiles -> fp32 (BH, S, DH): the same movement in the other direction
__global__ void __launch_bounds__(256) mlstm_unpack_kernel(const unsigned char* __restrict__ tiles, int S, int DH, int DHP, int nc,
                                                           float* __restrict__ dst) {
  extern __shared__ __align__(16) unsigned char pk_smem[];
  const int tile = blockIdx.x, tid = threadIdx.x;
  const int bh = tile / nc, c = tile % nc;
  const int t0 = c * kL, rows = min(kL, S - t0);
  const uint32_t tile_bytes = kL * DHP * 2;
  const uint4* in = reinterpret_cast<const uint4*>(tiles + static_cast<size_t>(tile) * tile_bytes);
  for (uint32_t i = tid; i < tile_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(pk_smem)[i] = __ldg(in + i);
  __syncthreads();
  float* blk = dst + (static_cast<size_t>(bh) * S + t0) * DH;
  const int n = rows * DH;
  if ((DH & 3) == 0 && (reinterpret_cast<uintptr_t>(blk) & 15) == 0) {
    for (int i = tid * 4; i < n; i += blockDim.x * 4) {
      const int r = i / DH, d = i % DH;
      const unsigned char* sp = pk_smem + tile_off16(kL, r, d / 8) + (d % 8) * 2;
      const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(sp)), b2 = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(sp + 4));
      *reinterpret_cast<float4*>(blk + i) = make_float4(a.x, a.y, b2.x, b2.y);
    }
  } else {
    for (int i = tid; i < n; i += blockDim.x) {
      const int r = i / DH, d = i % DH;
      blk[i] = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(pk_smem + tile_off16(kL, r, d / 8) + (d % 8) * 2));
    }
  }
}

// ------------------------------------------------------------------ host launchers
int launch_state_scan(int dhp, const float* dstate, const float* g, const float* amax, int BH, int nc, int reverse, void* states,
                      float* m_prev, cudaStream_t st);
int launch_chunk_out_ws(int dhp, const void* q, const void* k, const void* v, const float* ig, const float* fg, const void* states,
                        const float* m_prev, int BH, int nc, float scale, float eps, void* h, float* m, float* den,
                        cudaStream_t st);

int launch_chunk_state_ws(int dhp, const void* k, const void* v, const float* ig, const float* fg, int BH, int nc, float scale,
                          float* dstate, float* g_out, float* amax_out, cudaStream_t st);
// chunk_state / chunk_rstate on the persistent warp-specialised kernel (mlstm_state_ws.cu; dhp <= 32).  XHVED_STATE_WS=0 selects
// the one-tile-per-CTA kernels for A/B measurements.  Read once.
bool state_ws_enabled(int dhp) {
  static const int on = [] {
    const char* e = getenv("XHVED_STATE_WS");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  return on != 0 && dhp <= 32;
}

// Which chunk_out kernel runs: the persistent warp-specialised one with P kept in tensor memory (mlstm_fwd_ws.cu) -- the
// faster of the two at every head dim on B200 (profiles/r02_cell_scaling.jsonl).  XHVED_CELL_WS=0 selects the
// one-tile-per-CTA kernel above for A/B measurements.  Read once.
bool cell_ws_enabled(int) {
  static const int on = [] {
    const char* e = getenv("XHVED_CELL_WS");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  return on != 0;
}

template <int DHP>
static int launch_fwd(const void* q, const void* k, const void* v, const float* ig, const float* fg, int BH, int nc, int dh,
                      float eps, void* h, float* m, float* den, float* ws_dstate, float* ws_g, float* ws_amax, void* states,
                      float* m_prev, cudaStream_t st) {
  constexpr int NE = ext_cols(DHP);
  const float scale = 1.0f / sqrtf(static_cast<float>(dh));
  const int ntiles = BH * nc;
  // phase 1
  if (state_ws_enabled(DHP)) {
    if (int rc = launch_chunk_state_ws(DHP, k, v, ig, fg, BH, nc, scale, ws_dstate, ws_g, ws_amax, st)) return rc;
  } else {
    const size_t used = 2 * kL * DHP * 2 + kL * NE * 2, window = kL * DHP * 2 + 32768;
    const size_t smem = used > window ? used : window;
    cudaError_t e = cudaFuncSetAttribute(mlstm_chunk_state_kernel<DHP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ProfScope ps(K_CHUNK_STATE, st);
    mlstm_chunk_state_kernel<DHP><<<ntiles, kThreads, smem, st>>>((const unsigned char*)k, (const unsigned char*)v, ig, fg, nc, scale,
                                                                  ws_dstate, ws_g, ws_amax);
  }
  // phase 2
  if (int rc = launch_state_scan(DHP, ws_dstate, ws_g, ws_amax, BH, nc, 0, states, m_prev, st)) return rc;
  // phase 3
  if (cell_ws_enabled(DHP)) return launch_chunk_out_ws(DHP, q, k, v, ig, fg, states, m_prev, BH, nc, scale, eps, h, m, den, st);
  {
    const size_t smem = 3 * kL * DHP * 2 + kL * kL * 2 + 2 * DHP * NE * 2 + kL * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(mlstm_chunk_out_kernel<DHP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ProfScope ps(K_CHUNK_OUT, st);
    mlstm_chunk_out_kernel<DHP><<<ntiles, 2 * kThreads, smem, st>>>((const unsigned char*)q, (const unsigned char*)k, (const unsigned char*)v,
                                                                ig, fg, (const unsigned char*)states, m_prev, nc, scale, eps,
                                                                (unsigned char*)h, m, den);
  }
  return (int)cudaGetLastError();
}

// state scan launcher shared with the backward pass (reverse = 1 there)
int launch_state_scan(int dhp, const float* dstate, const float* g, const float* amax, int BH, int nc, int reverse, void* states,
                      float* m_prev, cudaStream_t st) {
  ProfScope ps(K_STATE_SCAN, st);
  const int groups = dhp * (dhp + 16) / 8;      // 16-byte groups of one state
  if (nc >= 64) {                               // long sequences: lane = chunk, one warp per group
    const int wpb = 8;
    const dim3 pgrid(BH, (groups + wpb - 1) / wpb);
    switch (dhp) {
      case 16: mlstm_state_scan_par_kernel<16><<<pgrid, 32 * wpb, 0, st>>>(dstate, g, amax, nc, reverse, (unsigned char*)states, m_prev); break;
      case 32: mlstm_state_scan_par_kernel<32><<<pgrid, 32 * wpb, 0, st>>>(dstate, g, amax, nc, reverse, (unsigned char*)states, m_prev); break;
      case 64: mlstm_state_scan_par_kernel<64><<<pgrid, 32 * wpb, 0, st>>>(dstate, g, amax, nc, reverse, (unsigned char*)states, m_prev); break;
      case 128: mlstm_state_scan_par_kernel<128><<<pgrid, 32 * wpb, 0, st>>>(dstate, g, amax, nc, reverse, (unsigned char*)states, m_prev); break;
      default: return XHVED_ERR_UNSUPPORTED_DH;
    }
    return (int)cudaGetLastError();
  }
  const int threads = groups < 256 ? groups : 256;
  const dim3 grid(BH, (groups + threads - 1) / threads);
  const size_t smem = static_cast<size_t>(nc) * 2 * sizeof(float);
  if (smem > 48 * 1024) return XHVED_ERR_BAD_SHAPE;
  switch (dhp) {
    case 16: mlstm_state_scan_kernel<16><<<grid, threads, smem, st>>>(dstate, g, amax, nc, reverse, (unsigned char*)states, m_prev); break;
    case 32: mlstm_state_scan_kernel<32><<<grid, threads, smem, st>>>(dstate, g, amax, nc, reverse, (unsigned char*)states, m_prev); break;
    case 64: mlstm_state_scan_kernel<64><<<grid, threads, smem, st>>>(dstate, g, amax, nc, reverse, (unsigned char*)states, m_prev); break;
    case 128: mlstm_state_scan_kernel<128><<<grid, threads, smem, st>>>(dstate, g, amax, nc, reverse, (unsigned char*)states, m_prev); break;
    default: return XHVED_ERR_UNSUPPORTED_DH;
  }
  return (int)cudaGetLastError();
}

}  // namespace xhved

using namespace xhved;

extern "C" int xhved_mlstm_fwd(const void* q_tiles, const void* k_tiles, const void* v_tiles, const float* ig, const float* fg,
                               int BH, int nc, int dh, int dhp, float eps, void* h_tiles, float* m, float* den, float* ws_dstate,
                               float* ws_g, float* ws_amax, void* states, float* m_prev, void* stream) {
  if (BH <= 0 || nc <= 0 || dh <= 0 || dh > dhp) return XHVED_ERR_BAD_SHAPE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (dhp) {
    case 16: return launch_fwd<16>(q_tiles, k_tiles, v_tiles, ig, fg, BH, nc, dh, eps, h_tiles, m, den, ws_dstate, ws_g, ws_amax, states, m_prev, st);
    case 32: return launch_fwd<32>(q_tiles, k_tiles, v_tiles, ig, fg, BH, nc, dh, eps, h_tiles, m, den, ws_dstate, ws_g, ws_amax, states, m_prev, st);
    case 64: return launch_fwd<64>(q_tiles, k_tiles, v_tiles, ig, fg, BH, nc, dh, eps, h_tiles, m, den, ws_dstate, ws_g, ws_amax, states, m_prev, st);
    case 128: return launch_fwd<128>(q_tiles, k_tiles, v_tiles, ig, fg, BH, nc, dh, eps, h_tiles, m, den, ws_dstate, ws_g, ws_amax, states, m_prev, st);
    default: return XHVED_ERR_UNSUPPORTED_DH;
  }
}

extern "C" int xhved_mlstm_pack(const float* src, int BH, int S, int dh, int dhp, void* tiles, void* stream) {
  if (BH <= 0 || S <= 0 || dh <= 0 || dh > dhp || dhp % 16) return XHVED_ERR_BAD_SHAPE;
  const int nc = (S + kL - 1) / kL;
  ProfScope ps(K_PACK, static_cast<cudaStream_t>(stream));
  mlstm_pack_kernel<<<BH * nc, 256, kL * dhp * 2, static_cast<cudaStream_t>(stream)>>>(src, S, dh, dhp, nc, (unsigned char*)tiles);
  return (int)cudaGetLastError();
}
extern "C" int xhved_mlstm_pack_gates(const float* ig, const float* fg, int BH, int S, float* ig_padded, float* fg_padded, void* stream) {
  if (BH <= 0 || S <= 0) return XHVED_ERR_BAD_SHAPE;
  const int nc = (S + kL - 1) / kL;
  ProfScope ps(K_PACK, static_cast<cudaStream_t>(stream));
  mlstm_pack_gates_kernel<<<BH * nc, kL, 0, static_cast<cudaStream_t>(stream)>>>(ig, fg, S, nc, ig_padded, fg_padded);
  return (int)cudaGetLastError();
}
extern "C" int xhved_mlstm_unpack(const void* tiles, int BH, int S, int dh, int dhp, float* dst, void* stream) {
  if (BH <= 0 || S <= 0 || dh <= 0 || dh > dhp || dhp % 16) return XHVED_ERR_BAD_SHAPE;
  const int nc = (S + kL - 1) / kL;
  ProfScope ps(K_UNPACK, static_cast<cudaStream_t>(stream));
  mlstm_unpack_kernel<<<BH * nc, 256, kL * dhp * 2, static_cast<cudaStream_t>(stream)>>>((const unsigned char*)tiles, S, dh, dhp, nc, dst);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------ UMMA self test (diagnostic entry point)
// D[128][N] = A * B^T with A, B given as tile-native tiles; a_mn / b_mn choose the MN-major view.
//   a_mn = 0: A tile is [128 rows (m)][K cols];   a_mn = 1: A tile is [K rows][128 cols (m)]
//   b_mn = 0: B tile is [N rows (n)][K cols];     b_mn = 1: B tile is [K rows][N cols (n)]
__global__ void __launch_bounds__(kThreads) umma_selftest_kernel(const unsigned char* __restrict__ a, const unsigned char* __restrict__ b, int N,
                                                                  int K, int a_mn, int b_mn, float* __restrict__ d) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t a_bytes = 128 * K * 2, b_bytes = N * K * 2;
  unsigned char* sA = smem;
  unsigned char* sB = smem + a_bytes;
  if (tid == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar_load, a_bytes + b_bytes);
    bulk_g2s(sA, a, a_bytes, &bar_load);
    bulk_g2s(sB, b, b_bytes, &bar_load);
  }
  mbar_wait(&bar_load, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t a_rows = a_mn ? K : 128, b_rows = b_mn ? K : N;
    const uint32_t a_lbo = a_mn ? 128 : a_rows * 16, a_sbo = a_mn ? a_rows * 16 : 128;
    const uint32_t b_lbo = b_mn ? 128 : b_rows * 16, b_sbo = b_mn ? b_rows * 16 : 128;
    umma_gemm(tmem, smem_u32(sA), a_lbo, a_sbo, smem_u32(sB), b_lbo, b_sbo, umma_idesc(128, N, a_mn != 0, b_mn != 0), K, false);
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    for (int i = 0; i < 16; ++i) d[tid * N + c0 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

extern "C" int xhved_umma_selftest(const void* a_tile, const void* b_tile, int N, int K, int a_mn, int b_mn, float* d, void* stream) {
  if (N % 16 || N < 16 || N > 256 || K % 16 || K < 16 || K > 128) return XHVED_ERR_BAD_SHAPE;
  const size_t smem = 128 * K * 2 + N * K * 2;
  cudaError_t e = cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  umma_selftest_kernel<<<1, kThreads, smem, static_cast<cudaStream_t>(stream)>>>((const unsigned char*)a_tile, (const unsigned char*)b_tile, N, K,
                                                                              a_mn, b_mn, d);
  return (int)cudaGetLastError();
}
