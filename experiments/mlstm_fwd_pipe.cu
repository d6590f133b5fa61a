// Chunkwise mLSTM forward, phase 3 (chunk_out) as a PERSISTENT, software-pipelined kernel -- sm_100a.
//
// Same arithmetic as mlstm_chunk_out_kernel (mlstm_fwd.cu; vision_lstm.py:99-128 in chunkwise form) but one CTA
// loops over many (b, head, chunk) tiles and keeps the TMA unit, the tensor core and the exp/epilogue warps busy at
// the same time:
//
//   control thread (warp 8, lane 0)   bulk loads of tile i+LA  ->  S_i = Q K^T (tcgen05)  ->  O_{i-1} = P V + Q [C|n]
//   softmax group 0 (warps 0-3)       even tiles:  E(i): gate scans, D' = exp2(u_t + v_s), P_i -> smem;  epi(i): O_i from TMEM,
//   softmax group 1 (warps 4-7)       odd tiles          normalise, store h      (the two groups ping-pong, so one group's exp
//                                                                                  phase overlaps the other's MMAs / epilogue)
//
// Operand stages (Q, K, V, state hi/lo, gates) form an NST-deep ring in shared memory; P tiles and TMEM accumulators
// are double buffered.  mbarriers: full[NST] (TMA bytes landed), sfree[NST] (MMAs finished reading a stage),
// s_ready[2] / o_ready[2] (tcgen05.commit), p_ready[2] / tfree[2] (128 thread arrivals).
#include "mlstm_common.cuh"
#include "prof.cuh"
#include "xhved.h"

namespace xhved {

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// named barrier 1 + g for softmax group g (128 threads each); barrier 0 stays the CTA-wide __syncthreads
__device__ __forceinline__ void softmax_group_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }

// 128-thread scans over the softmax group only (named barrier 1; the control warp does not take part)
__device__ __forceinline__ float group_cumsum128(float x, float* red, int g) {
  const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) & 3;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) red[warp] = x;
  softmax_group_sync(g);
  float off = 0.f;
#pragma unroll
  for (int w = 0; w < 4; ++w)
    if (w < warp) off += red[w];
  softmax_group_sync(g);
  return x + off;
}
__device__ __forceinline__ float group_cummax128(float x, float* red, int g) {
  const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) & 3;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x = fmaxf(x, y);
  }
  if (lane == 31) red[warp] = x;
  softmax_group_sync(g);
  float off = -INFINITY;
#pragma unroll
  for (int w = 0; w < 4; ++w)
    if (w < warp) off = fmaxf(off, red[w]);
  softmax_group_sync(g);
  return fmaxf(x, off);
}

template <int DHP, int NST>
struct PipeCfg {
  static constexpr int NE = ext_cols(DHP);
  static constexpr uint32_t TILE = kL * DHP * 2;
  static constexpr uint32_t ST_BYTES = 2 * DHP * NE * 2;        // state hi + lo
  static constexpr uint32_t GATE_BYTES = 2 * kL * 4;            // ig | fg of the chunk
  static constexpr uint32_t STAGE = 3 * TILE + ST_BYTES + GATE_BYTES;
  static constexpr uint32_t O_Q = 0, O_K = TILE, O_V = 2 * TILE, O_S = 3 * TILE, O_G = 3 * TILE + ST_BYTES;
  static constexpr uint32_t P_BYTES = kL * kL * 2;
  static constexpr uint32_t P0 = NST * STAGE;                   // two P tiles
  static constexpr uint32_t VCOL = P0 + 2 * P_BYTES;            // two vcol arrays
  static constexpr uint32_t TOTAL = VCOL + 2 * kL * 4;
  static constexpr uint32_t BUFCOLS = (2 * DHP + 16) > 128 ? (2 * DHP + 16) : 128;   // S, later O_intra | O_inter
  static constexpr uint32_t TMEM_COLS = next_pow2_cols(2 * BUFCOLS);
  static constexpr int LA = NST - 1;                            // tiles of load look-ahead
};

template <int DHP, int NST>
__global__ void __launch_bounds__(288) mlstm_chunk_out_pipe_kernel(
    const unsigned char* __restrict__ q_tiles, const unsigned char* __restrict__ k_tiles, const unsigned char* __restrict__ v_tiles,
    const float* __restrict__ ig, const float* __restrict__ fg, const unsigned char* __restrict__ states,
    const float* __restrict__ m_prev, int nc, int ntiles, float scale, float eps, unsigned char* __restrict__ h_tiles,
    float* __restrict__ m_out, float* __restrict__ den_out) {
  using L = PipeCfg<DHP, NST>;
  constexpr int NE = L::NE;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[NST], sfree[NST], s_ready[2], p_ready[2], o_ready[2], tfree[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float red2[2][4];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_my = (ntiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&sfree[s], 1);
    }
    for (int j = 0; j < 2; ++j) {
      mbar_init(&s_ready[j], 1);
      mbar_init(&p_ready[j], kL);
      mbar_init(&o_ready[j], 1);
      mbar_init(&tfree[j], kL);
    }
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(&tmem_slot, L::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 8) {
    // ======================================================= control thread: TMA producer + MMA issuer
    if ((tid & 31) == 0) {
      auto issue_loads = [&](int t) {
        const int s = t % NST;
        const int tile = blockIdx.x + t * gridDim.x;
        const bool has_state = (tile % nc) > 0;
        unsigned char* st = smem + s * L::STAGE;
        mbar_expect_tx(&full[s], 3 * L::TILE + L::GATE_BYTES + (has_state ? L::ST_BYTES : 0));
        bulk_g2s(st + L::O_Q, q_tiles + static_cast<size_t>(tile) * L::TILE, L::TILE, &full[s]);
        bulk_g2s(st + L::O_K, k_tiles + static_cast<size_t>(tile) * L::TILE, L::TILE, &full[s]);
        bulk_g2s(st + L::O_V, v_tiles + static_cast<size_t>(tile) * L::TILE, L::TILE, &full[s]);
        if (has_state) bulk_g2s(st + L::O_S, states + static_cast<size_t>(tile) * L::ST_BYTES, L::ST_BYTES, &full[s]);
        bulk_g2s(st + L::O_G, ig + static_cast<size_t>(tile) * kL, kL * 4, &full[s]);
        bulk_g2s(st + L::O_G + kL * 4, fg + static_cast<size_t>(tile) * kL, kL * 4, &full[s]);
      };
      auto issue_pv = [&](int pt) {
        const int jp = pt & 1, sp = pt % NST;
        const int tile = blockIdx.x + pt * gridDim.x;
        const bool has_state = (tile % nc) > 0;
        unsigned char* st = smem + sp * L::STAGE;
        const uint32_t tb = tmem + jp * L::BUFCOLS;
        mbar_wait(&p_ready[jp], (pt >> 1) & 1);
        tc_fence_after();
        // O_intra[t][e] = sum_s P[t][s] V[s][e]
        umma_gemm(tb, smem_u32(smem + L::P0 + jp * L::P_BYTES), kL * 16, 128, smem_u32(st + L::O_V), 128, kL * 16,
                  umma_idesc(128, DHP, false, true), kL, false);
        if (has_state) {
          // O_inter[t][e'] = sum_d Q[t][d] [C|n][d][e']   (hi + lo state tiles)
          umma_gemm(tb + DHP, smem_u32(st + L::O_Q), kL * 16, 128, smem_u32(st + L::O_S), 128, DHP * 16, umma_idesc(128, NE, false, true),
                    DHP, false);
          umma_gemm(tb + DHP, smem_u32(st + L::O_Q), kL * 16, 128, smem_u32(st + L::O_S) + L::ST_BYTES / 2, 128, DHP * 16,
                    umma_idesc(128, NE, false, true), DHP, true);
        }
        umma_commit(&o_ready[jp]);
        umma_commit(&sfree[sp]);
      };
      for (int t = 0; t < L::LA && t < n_my; ++t) issue_loads(t);
      for (int i = 0; i < n_my; ++i) {
        const int s = i % NST, j = i & 1;
        // B: S_i = Q K^T as soon as the operands have landed and the TMEM buffer has been drained
        mbar_wait(&full[s], (i / NST) & 1);
        if (i >= 2) mbar_wait(&tfree[j], ((i >> 1) - 1) & 1);
        tc_fence_after();
        unsigned char* st = smem + s * L::STAGE;
        umma_gemm(tmem + j * L::BUFCOLS, smem_u32(st + L::O_Q), kL * 16, 128, smem_u32(st + L::O_K), kL * 16, 128,
                  umma_idesc(128, kL, false, false), DHP, false);
        umma_commit(&s_ready[j]);
        // C: O_{i-1}
        if (i >= 1) issue_pv(i - 1);
        // A: refill the ring LA tiles ahead (its stage was last used by tile t - NST)
        const int t = i + L::LA;
        if (t < n_my) {
          if (t >= NST) mbar_wait(&sfree[t % NST], ((t / NST) - 1) & 1);
          issue_loads(t);
        }
      }
      if (n_my > 0) issue_pv(n_my - 1);
    }
    __syncwarp();      // lanes 1..31 wait here for the control lane before the CTA-wide barrier below
  } else {
    // ======================================================= softmax / epilogue groups (one thread per chunk row)
    const int grp = tid >> 7, row = tid & (kL - 1), gwarp = (tid >> 5) & 3;
    const uint32_t lane_base = static_cast<uint32_t>(gwarp * 32) << 16;
    float* red = red2[grp];
    const int j = grp;                                   // this group's TMEM / P buffer
    float* vcol = reinterpret_cast<float*>(smem + L::VCOL) + j * kL;
    unsigned char* sP = smem + L::P0 + j * L::P_BYTES;
    const uint32_t tb = tmem + j * L::BUFCOLS;
    for (int i = grp; i < n_my; i += 2) {
      // ---------------- E(i)
      const int s = i % NST;
      const int tile = blockIdx.x + i * gridDim.x;
      const bool has_state = (tile % nc) > 0;
      const float mp = has_state ? __ldg(m_prev + tile) : -INFINITY;
      const float* gates = reinterpret_cast<const float*>(smem + s * L::STAGE + L::O_G);
      mbar_wait(&full[s], (i / NST) & 1);
      const float iv = gates[row];
      const float lf = log_sigmoid(gates[kL + row]);
      const float b = group_cumsum128(lf, red, grp);
      const float vc = iv - b;
      const float m_intra = b + group_cummax128(vc, red, grp);
      const float m_inter = b + mp;
      const float m = fmaxf(m_intra, m_inter);
      const float w = has_state ? __expf(m_inter - m) : 0.f;
      vcol[row] = vc * kLog2e;
      const float urow = (b - m) * kLog2e + log2f(scale);
      softmax_group_sync(grp);      // vcol visible to the group
      mbar_wait(&s_ready[j], (i >> 1) & 1);
      tc_fence_after();
      float rowsum = 0.f;
#pragma unroll 1
      for (int blk = 0; blk < 4; ++blk) {
        if (blk <= gwarp) {
          float sv[32];
          tmem_ld32(tb + lane_base + blk * 32, sv);
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            float p[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const int sc = blk * 32 + j8 * 8 + jj;
              const float d = fast_exp2(urow + vcol[sc]);
              p[jj] = (sc <= row) ? sv[j8 * 8 + jj] * d : 0.f;
              rowsum += p[jj];
            }
            *reinterpret_cast<uint4*>(sP + tile_off16(kL, row, blk * 4 + j8)) = pack8_bf16(p);
          }
        } else {
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) *reinterpret_cast<uint4*>(sP + tile_off16(kL, row, blk * 4 + j8)) = make_uint4(0, 0, 0, 0);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&p_ready[j]);
      // ---------------- epi(i): h = (O_intra + w O_inter) / (max(|den|, exp(-m)) + eps)   (vision_lstm.py:123-128)
      mbar_wait(&o_ready[j], (i >> 1) & 1);
      tc_fence_after();
      float den = rowsum;
      if (has_state) {
        float t8[8];
        tmem_ld8(tb + lane_base + 2 * DHP, t8);
        den += w * t8[0];
      }
      const float rn = 1.f / (fmaxf(fabsf(den), __expf(-m)) + eps);
      unsigned char* hdst = h_tiles + static_cast<size_t>(tile) * L::TILE;
#pragma unroll
      for (int c0 = 0; c0 < DHP; c0 += 16) {
        float o[16];
        tmem_ld16(tb + lane_base + c0, o);
        if (has_state) {
          float oi[16];
          tmem_ld16(tb + lane_base + DHP + c0, oi);
#pragma unroll
          for (int k = 0; k < 16; ++k) o[k] += w * oi[k];
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) o[k] *= rn;
        *reinterpret_cast<uint4*>(hdst + tile_off16(kL, row, c0 / 8)) = pack8_bf16(o);
        *reinterpret_cast<uint4*>(hdst + tile_off16(kL, row, c0 / 8 + 1)) = pack8_bf16(o + 8);
      }
      const size_t grow = static_cast<size_t>(tile) * kL + row;
      m_out[grow] = m;
      den_out[grow] = den;
      tc_fence_before();
      mbar_arrive(&tfree[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, L::TMEM_COLS);
}

static int sm_count() {
  static int n = [] {
    int dev = 0, v = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v;
  }();
  return n;
}

template <int DHP, int NST>
static int launch_pipe(const void* q, const void* k, const void* v, const float* ig, const float* fg, const void* states,
                       const float* m_prev, int ntiles, int nc, float scale, float eps, void* h, float* m, float* den, cudaStream_t st) {
  using L = PipeCfg<DHP, NST>;
  auto kern = mlstm_chunk_out_pipe_kernel<DHP, NST>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::TOTAL);
  if (e != cudaSuccess) return (int)e;
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 288, L::TOTAL);
  const int by_tmem = 512 / L::TMEM_COLS;
  if (per_sm > by_tmem) per_sm = by_tmem;
  if (per_sm < 1) per_sm = 1;
  int grid = sm_count() * per_sm;
  if (grid > ntiles) grid = ntiles;
  ProfScope ps(K_CHUNK_OUT, st);
  kern<<<grid, 288, L::TOTAL, st>>>((const unsigned char*)q, (const unsigned char*)k, (const unsigned char*)v, ig, fg,
                                    (const unsigned char*)states, m_prev, nc, ntiles, scale, eps, (unsigned char*)h, m, den);
  return (int)cudaGetLastError();
}

// dispatcher used by mlstm_fwd.cu; returns -1000 when the pipelined kernel does not cover dhp
int launch_chunk_out_pipelined(int dhp, const void* q, const void* k, const void* v, const float* ig, const float* fg, const void* states,
                               const float* m_prev, int ntiles, int nc, float scale, float eps, void* h, float* m, float* den,
                               cudaStream_t st) {
  switch (dhp) {
    case 16: return launch_pipe<16, 3>(q, k, v, ig, fg, states, m_prev, ntiles, nc, scale, eps, h, m, den, st);
    case 32: return launch_pipe<32, 3>(q, k, v, ig, fg, states, m_prev, ntiles, nc, scale, eps, h, m, den, st);
    case 64: return launch_pipe<64, 2>(q, k, v, ig, fg, states, m_prev, ntiles, nc, scale, eps, h, m, den, st);
    default: return -1000;
  }
}

}  // namespace xhved
