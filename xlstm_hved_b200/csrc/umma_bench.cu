// Diagnostic: cost of issuing tcgen05.mma (M = 128, bf16) as a function of N, of the number of independent accumulators the
// K = 16 slices are spread over, and of where the A operand lives.  One CTA, one issuing thread; clock64 around the issue
// loop and around issue + completion (tcgen05.commit -> mbarrier).  Used to size the MMA phases of the cell kernels
// (DESIGN.md section 5); not part of the product path.
#include "mlstm_common.cuh"
#include "xhved.h"

namespace xhved {

__global__ void __launch_bounds__(kThreads) umma_issue_bench_kernel(int N, int reps, int n_acc, int a_in_tmem, int mn_major,
                                                                    long long* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];   // A: 128 x 128 bf16 (32 KB), B: 128 x 128 bf16 (32 KB), zero-filled
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 65536 / 16; i += kThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (mn_major == 3) {
    // whole-warp issue with provably uniform operands: tensor-memory base through __shfl_sync, election inside the asm
    if (warp == 0) {
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t aA = smem_u32(smem), aB = smem_u32(smem + 32768);
      const uint32_t idesc = umma_idesc(128, N, false, false);
      const uint32_t lbo = kL * 16, sbo = 128;
      const uint64_t ad0 = umma_desc(aA, lbo, sbo), bd0 = umma_desc(aB, lbo, sbo);
      const uint64_t step = (2u * lbo) >> 4;
      const int sh = n_acc == 1 ? 0 : (n_acc == 2 ? 1 : 2);
      const long long t0 = clock64();
#pragma unroll 8
      for (int r = 0; r < reps; ++r) {
        const int acc = r & (n_acc - 1), k = (r >> sh) & 7;
        const uint32_t d = tm + 256 + acc * (N < 32 ? 32 : N);
        if (a_in_tmem)
          umma_bf16_ts_e(d, tm + 8u * k, bd0 + step * k, idesc, r >= n_acc ? 1u : 0u);
        else
          umma_bf16_e(d, ad0 + step * k, bd0 + step * k, idesc, r >= n_acc ? 1u : 0u);
      }
      const long long t1 = clock64();
      umma_commit_e(&bar);
      mbar_wait(&bar, 0);
      const long long t2 = clock64();
      if (tid == 0) out[0] = t1 - t0, out[1] = t2 - t0;
    }
  } else if (mn_major >= 2) {
    // whole-warp issue: uniform control flow, the instruction itself predicated on one lane (umma_bf16_p)
    if (warp == 0) {
      const uint32_t leader = tid == 0 ? 1u : 0u;
      const uint32_t aA = smem_u32(smem), aB = smem_u32(smem + 32768);
      const uint32_t idesc = umma_idesc(128, N, false, false);
      const uint32_t lbo = kL * 16, sbo = 128;
      const uint64_t ad0 = umma_desc(aA, lbo, sbo), bd0 = umma_desc(aB, lbo, sbo);
      const uint64_t step = (2u * lbo) >> 4;
      const int sh = n_acc == 1 ? 0 : (n_acc == 2 ? 1 : 2);
      const long long t0 = clock64();
#pragma unroll 8
      for (int r = 0; r < reps; ++r) {
        const int acc = r & (n_acc - 1), k = (r >> sh) & 7;
        const uint32_t d = tmem + 256 + acc * (N < 32 ? 32 : N);
        if (a_in_tmem)
          umma_bf16_ts_p(d, tmem + 8u * k, bd0 + step * k, idesc, r >= n_acc ? 1u : 0u, leader);
        else
          umma_bf16_p(d, ad0 + step * k, bd0 + step * k, idesc, r >= n_acc ? 1u : 0u, leader);
      }
      const long long t1 = clock64();
      umma_commit_p(&bar, leader);
      mbar_wait(&bar, 0);
      const long long t2 = clock64();
      if (tid == 0) out[0] = t1 - t0, out[1] = t2 - t0;
    }
  } else if (tid == 0) {
    const uint32_t aA = smem_u32(smem), aB = smem_u32(smem + 32768);
    const uint32_t idesc = umma_idesc(128, N, mn_major != 0, mn_major != 0);
    // K-major: lbo = 128 rows * 16 B, sbo = 128;  MN-major: lbo = 128, sbo = 128 * 16
    const uint32_t lbo = mn_major ? 128 : kL * 16, sbo = mn_major ? kL * 16 : 128;
    const uint64_t ad0 = umma_desc(aA, lbo, sbo), bd0 = umma_desc(aB, lbo, sbo);
    const uint64_t step = (2u * lbo) >> 4;
    const long long t0 = clock64();
    const int sh = n_acc == 1 ? 0 : (n_acc == 2 ? 1 : 2);      // n_acc in {1, 2, 4}: no integer division in the timed loop
#pragma unroll 8
    for (int r = 0; r < reps; ++r) {
      const int acc = r & (n_acc - 1), k = (r >> sh) & 7;
      const uint32_t d = tmem + 256 + acc * (N < 32 ? 32 : N);      // accumulators behind the A region
      if (a_in_tmem)
        umma_bf16_ts(d, tmem + 8u * k, bd0 + step * k, umma_idesc(128, N, false, mn_major != 0), r >= n_acc ? 1u : 0u);
      else
        umma_bf16(d, ad0 + step * k, bd0 + step * k, idesc, r >= n_acc ? 1u : 0u);
    }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace xhved

extern "C" int xhved_umma_issue_bench(int N, int reps, int n_acc, int a_in_tmem, int mn_major, long long* out2, void* stream) {
  if (N % 16 || N < 16 || N > 128 || reps < 1 || (n_acc != 1 && n_acc != 2 && n_acc != 4) || n_acc * (N < 32 ? 32 : N) > 256) return XHVED_ERR_BAD_SHAPE;
  cudaError_t e = cudaFuncSetAttribute(xhved::umma_issue_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  if (e != cudaSuccess) return (int)e;
  xhved::umma_issue_bench_kernel<<<1, xhved::kThreads, 65536, static_cast<cudaStream_t>(stream)>>>(N, reps, n_acc, a_in_tmem, mn_major, out2);
  return (int)cudaGetLastError();
}
