// Shared device helpers for the chunkwise mLSTM kernels (forward and backward).
#pragma once
#include "umma.cuh"

namespace xhved {

constexpr int kL = 128;        // chunk length == UMMA M == TMEM lanes
constexpr int kThreads = 128;  // one thread per chunk row

// Extended ("ext") column count: a [value | 1 | 0...] / [C | n | 0...] block is DHP+16 wide.
__host__ __device__ constexpr int ext_cols(int dhp) { return dhp + 16; }
__host__ __device__ constexpr uint32_t next_pow2_cols(int n) { return n <= 32 ? 32u : n <= 64 ? 64u : n <= 128 ? 128u : n <= 256 ? 256u : 512u; }

// ---- 128-wide block scans (4 warps).  `red` is 8 floats of shared scratch.  A 256-thread CTA may call them too: its two
// 128-thread halves then scan the same kind of 128 values independently (threads 128.. use red[4..7]). ----
__device__ __forceinline__ float block_cumsum128(float x, float* red, float* total) {
  const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) & 3;
  red += (threadIdx.x >> 7) << 2;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) red[warp] = x;
  __syncthreads();
  float off = 0.f, tot = 0.f;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const float r = red[w];
    if (w < warp) off += r;
    tot += r;
  }
  __syncthreads();
  if (total) *total = tot;
  return x + off;
}
__device__ __forceinline__ float block_cummax128(float x, float* red, float* total) {
  const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) & 3;
  red += (threadIdx.x >> 7) << 2;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x = fmaxf(x, y);
  }
  if (lane == 31) red[warp] = x;
  __syncthreads();
  float off = -INFINITY, tot = -INFINITY;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const float r = red[w];
    if (w < warp) off = fmaxf(off, r);
    tot = fmaxf(tot, r);
  }
  __syncthreads();
  if (total) *total = tot;
  return fmaxf(x, off);
}
// reverse (suffix) inclusive cumulative sum over the 128 threads
__device__ __forceinline__ float block_rcumsum128(float x, float* red, float* total) {
  const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) & 3;
  red += (threadIdx.x >> 7) << 2;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float y = __shfl_down_sync(0xffffffffu, x, o);
    if (lane + o < 32) x += y;
  }
  if (lane == 0) red[warp] = x;
  __syncthreads();
  float off = 0.f, tot = 0.f;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const float r = red[w];
    if (w > warp) off += r;
    tot += r;
  }
  __syncthreads();
  if (total) *total = tot;
  return x + off;
}

// write the two "ext" column groups [1,0,...,0 | 0...0] of row r behind a 128-row tile of dhp columns
__device__ __forceinline__ void write_ext_ones(unsigned char* tile, int dhp, int r) {
  uint4 one = make_uint4(0x00003F80u, 0u, 0u, 0u);  // bf16(1.0) in element 0
  uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  *reinterpret_cast<uint4*>(tile + tile_off16(kL, r, dhp / 8)) = one;
  *reinterpret_cast<uint4*>(tile + tile_off16(kL, r, dhp / 8 + 1)) = zero;
}

// Scale row r of a [128][DHP] tile by w, writing the product as a bf16 hi + lo pair (hi in place, lo to `lo`).
template <int DHP>
__device__ __forceinline__ void scale_row_hilo(unsigned char* hi, unsigned char* lo, int r, float w) {
#pragma unroll
  for (int cg = 0; cg < DHP / 8; ++cg) {
    uint4* p = reinterpret_cast<uint4*>(hi + tile_off16(kL, r, cg));
    const uint4 u = *p;
    const float2 f[4] = {unpack_bf16x2(u.x), unpack_bf16x2(u.y), unpack_bf16x2(u.z), unpack_bf16x2(u.w)};
    uint32_t oh[4], ol[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float a = f[i].x * w, b = f[i].y * w;
      oh[i] = pack_bf16x2(a, b);
      const float2 back = unpack_bf16x2(oh[i]);
      ol[i] = pack_bf16x2(a - back.x, b - back.y);
    }
    *p = make_uint4(oh[0], oh[1], oh[2], oh[3]);
    *reinterpret_cast<uint4*>(lo + tile_off16(kL, r, cg)) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
  }
}

// Epilogue of chunk_state / chunk_rstate: rows d < DHP of the accumulator (plus, with ONE_PASS, rows d + DHP: the lo half of
// the hi/lo operand, see the callers) -> out[d][0 .. NE) fp32.  DHP = 16: both row groups live in warp 0 (lanes d, d + 16);
// DHP = 32 / 64: the lo rows belong to other warps and travel through `scratch` (>= DHP * NE floats of shared memory that
// the finished MMAs no longer read).
template <int DHP, bool ONE_PASS>
__device__ __forceinline__ void store_state_rows(uint32_t tmem, float* __restrict__ out, float* scratch) {
  constexpr int NE = ext_cols(DHP);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
  const int rows = ONE_PASS ? 2 * DHP : DHP;
  if (ONE_PASS && DHP >= 32) {
    if (tid >= DHP && tid < 2 * DHP) {
#pragma unroll
      for (int c0 = 0; c0 < NE; c0 += 16) {
        float v[16];
        tmem_ld16(tmem + lane_base + c0, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) scratch[(c0 + i) * DHP + (tid - DHP)] = v[i];
      }
    }
    __syncthreads();
  }
  if (warp * 32 < (DHP >= 32 ? DHP : rows)) {
#pragma unroll
    for (int c0 = 0; c0 < NE; c0 += 16) {
      float v[16];
      tmem_ld16(tmem + lane_base + c0, v);
      if (ONE_PASS && DHP == 16) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += __shfl_down_sync(0xffffffffu, v[i], 16);
      } else if (ONE_PASS) {
        if (tid < DHP) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += scratch[(c0 + i) * DHP + tid];
        }
      }
      if (tid < DHP) {
        float* o = out + static_cast<size_t>(tid) * NE + c0;
#pragma unroll
        for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
    }
  }
}

}  // namespace xhved
