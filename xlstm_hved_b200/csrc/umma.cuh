// Blackwell (sm_100a) primitives used by the mLSTM kernels: mbarrier, 1-D bulk
// async copies (TMA unit, UBLKCP), tcgen05 MMA with TMEM accumulators.
//
// Operand tiles live in shared memory in ONE canonical no-swizzle layout, the
// "tile-native" layout, which is also how q/k/v/dh tiles are stored in HBM so a
// single bulk copy lands them MMA-ready:
//
//     tile[R rows][Ccols] bf16  ->  element (r, c) at byte
//         ((c / 8) * R + r) * 16 + (c % 8) * 2
//
// i.e. 8x8 core matrices (8 rows x 16 B, 128 B contiguous), row groups 128 B
// apart, column groups R*16 B apart.  The same bytes serve as
//   * a K-major operand  (MN = row, K = col):  SBO = 128,    LBO = R*16
//   * an MN-major operand (MN = col, K = row): SBO = R*16,   LBO = 128
// (SBO = stride between 8-wide MN groups, LBO = stride between 8-wide K groups),
// so no transposed staging is ever needed.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace xhved {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// ---- "leader" forms.  tcgen05.mma / tcgen05.commit / cp.async.bulk take their operands from UNIFORM registers.  Issued from
// inside a divergent `if (lane == 0)` region every operand goes through R2UR (~100 cycles per instruction, measured); the
// forms below are meant to be executed by a WHOLE warp in uniform control flow with identical operands, only the
// instruction itself being predicated on `leader` (one lane), so the operands can live in uniform registers.
__device__ __forceinline__ void mbar_expect_tx_p(uint64_t* bar, uint32_t bytes, uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %2, 0;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
      "r"(bytes), "r"(leader)
      : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait with a suspend-time hint: the warp sleeps in hardware (up to `ns`) instead of spinning through issue slots
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
// Bounded wait: a producer that never arrives traps the kernel (reported as a
// launch failure through the C ABI) instead of hanging the device.  The clock is
// read once per 64 polls only: the poll loop itself stays three instructions.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = 0;
  for (uint32_t it = 1;; ++it) {
    if (mbar_try_wait_hint(bar, parity, 20000u)) return;
    if ((it & 63u) == 0u) {
      const long long t = clock64();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 4000000000LL) __trap();
    }
  }
}

// Bounded SPINNING wait (no suspend-time hint): for short hand-offs between the roles of a warp-specialised kernel, where the
// wake-up after a hinted sleep costs more than the phase being waited for.
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  long long t0 = 0;
  for (uint32_t it = 1;; ++it) {
    if (mbar_try_wait(bar, parity)) return;
    if ((it & 1023u) == 0u) {
      const long long t = clock64();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 4000000000LL) __trap();
    }
  }
}

// ---------------------------------------------------------------- bulk copy (TMA unit, no tensor map)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void bulk_g2s_p(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %4, 0;\n\t"
      "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}" ::"r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "r"(leader)
      : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma / bulk copies)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// one full warp; ncols power of two in [32, 512]; result written to *smem_slot
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// shared-memory matrix descriptor, no swizzle (SmemDescriptor of the sm_100 UMMA ISA):
// [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1, [61,64) layout=0
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

// instruction descriptor for kind::f16, bf16 x bf16 -> fp32:
// [4,6) c=F32(1), [7,10) a=BF16(1), [10,13) b=BF16(1), 15 a_major, 16 b_major (1 = MN-major),
// [17,23) N>>3, [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T, one K=16 slice.  Issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_p(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                            uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}
// "elect" forms: executed by a whole CONVERGED warp with warp-uniform operands; one lane is elected inside the asm.  With
// operands the compiler can prove uniform (kernel parameters, blockIdx, loop counters, __shfl_sync(.., 0) results) the
// tcgen05 / bulk-copy instruction is emitted straight from uniform registers -- no per-lane "waterfall" loop around it.
__device__ __forceinline__ void umma_bf16_e(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ts_e(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_e(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_e(uint64_t* bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
      "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s_e(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}" ::"r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// whole-warp forms of umma_gemm / umma_gemm_ts (uniform operands, see above)
__device__ __forceinline__ void umma_gemm_e(uint32_t d_tmem, uint32_t a_addr, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_addr,
                                            uint32_t b_lbo, uint32_t b_sbo, uint32_t idesc, int K, bool accumulate_first) {
  uint64_t ad = umma_desc(a_addr, a_lbo, a_sbo);
  uint64_t bd = umma_desc(b_addr, b_lbo, b_sbo);
  const uint64_t a_step = (2u * a_lbo) >> 4, b_step = (2u * b_lbo) >> 4;
  for (int k = 0; k < K / 16; ++k) {
    umma_bf16_e(d_tmem, ad, bd, idesc, (k > 0 || accumulate_first) ? 1u : 0u);
    ad += a_step;
    bd += b_step;
  }
}
__device__ __forceinline__ void umma_gemm_ts_e(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_addr, uint32_t b_lbo, uint32_t b_sbo,
                                               uint32_t idesc, int K, bool accumulate_first) {
  uint64_t bd = umma_desc(b_addr, b_lbo, b_sbo);
  const uint64_t b_step = (2u * b_lbo) >> 4;
  for (int k = 0; k < K / 16; ++k) {
    umma_bf16_ts_e(d_tmem, a_tmem + 8u * k, bd, idesc, (k > 0 || accumulate_first) ? 1u : 0u);
    bd += b_step;
  }
}
__device__ __forceinline__ void umma_commit_p(uint64_t* bar, uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)),
      "r"(leader)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread are done
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// A full GEMM slice sequence: D[128 x N] (+)= A[128 x K] * B[N x K]^T, K multiple of 16.
// a_addr/b_addr: smem byte addresses of tile-native tiles; *_lbo/_sbo as documented above.
__device__ __forceinline__ void umma_gemm(uint32_t d_tmem, uint32_t a_addr, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_addr,
                                          uint32_t b_lbo, uint32_t b_sbo, uint32_t idesc, int K, bool accumulate_first) {
  // descriptors are built once; each K=16 slice only advances the 14-bit start-address field (2 core matrices along K)
  uint64_t ad = umma_desc(a_addr, a_lbo, a_sbo);
  uint64_t bd = umma_desc(b_addr, b_lbo, b_sbo);
  const uint64_t a_step = (2u * a_lbo) >> 4, b_step = (2u * b_lbo) >> 4;
  for (int k = 0; k < K / 16; ++k) {
    umma_bf16(d_tmem, ad, bd, idesc, (k > 0 || accumulate_first) ? 1u : 0u);
    ad += a_step;
    bd += b_step;
  }
}

// One accumulation chain: D[128 x N] (+)= sum over K slices of A_k B_k^T.  The K = 16 slices of ONE chain are dependent
// (each reads the accumulator the previous one wrote): with small N they do not cover the tensor pipe's latency, so several
// independent chains are issued round-robin (umma_issue_interleaved) instead of one after the other.
struct UmmaChain {
  uint32_t d_tmem;      // accumulator
  uint64_t ad, bd;      // running descriptors (A unused when a_in_tmem)
  uint32_t a_tmem;      // A operand in tensor memory (packed bf16, 8 columns per slice) when a_in_tmem
  uint32_t a_step, b_step, idesc;
  int nk;               // remaining slices
  uint32_t acc;         // accumulate flag of the next slice
  bool a_in_tmem;
};
__device__ __forceinline__ UmmaChain umma_chain(uint32_t d_tmem, uint32_t a_addr, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_addr,
                                                uint32_t b_lbo, uint32_t b_sbo, uint32_t idesc, int K, bool accumulate_first) {
  UmmaChain c;
  c.d_tmem = d_tmem, c.ad = umma_desc(a_addr, a_lbo, a_sbo), c.bd = umma_desc(b_addr, b_lbo, b_sbo), c.a_tmem = 0;
  c.a_step = (2u * a_lbo) >> 4, c.b_step = (2u * b_lbo) >> 4, c.idesc = idesc, c.nk = K / 16, c.acc = accumulate_first ? 1u : 0u;
  c.a_in_tmem = false;
  return c;
}
__device__ __forceinline__ UmmaChain umma_chain_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_addr, uint32_t b_lbo, uint32_t b_sbo,
                                                   uint32_t idesc, int K, bool accumulate_first) {
  UmmaChain c;
  c.d_tmem = d_tmem, c.ad = 0, c.bd = umma_desc(b_addr, b_lbo, b_sbo), c.a_tmem = a_tmem;
  c.a_step = 8u, c.b_step = (2u * b_lbo) >> 4, c.idesc = idesc, c.nk = K / 16, c.acc = accumulate_first ? 1u : 0u;
  c.a_in_tmem = true;
  return c;
}
// continue a chain with a second operand pair into the same accumulator (e.g. the lo half of a hi/lo split)
__device__ __forceinline__ void umma_chain_rebase(UmmaChain& c, uint32_t a_addr, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_addr,
                                                  uint32_t b_lbo, uint32_t b_sbo, int K) {
  c.ad = umma_desc(a_addr, a_lbo, a_sbo), c.bd = umma_desc(b_addr, b_lbo, b_sbo);
  c.a_step = (2u * a_lbo) >> 4, c.b_step = (2u * b_lbo) >> 4, c.nk = K / 16;
}
__device__ __forceinline__ void umma_chain_step(UmmaChain& c, uint32_t leader = 1u);
template <int N>
__device__ __forceinline__ void umma_issue_interleaved(UmmaChain (&c)[N], uint32_t leader = 1u) {
  bool any = true;
  while (any) {
    any = false;
#pragma unroll
    for (int i = 0; i < N; ++i)
      if (c[i].nk > 0) {
        umma_chain_step(c[i], leader);
        any = true;
      }
  }
}

// whole-warp form of umma_gemm (see the "leader" forms above)
__device__ __forceinline__ void umma_gemm_p(uint32_t d_tmem, uint32_t a_addr, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_addr,
                                            uint32_t b_lbo, uint32_t b_sbo, uint32_t idesc, int K, bool accumulate_first, uint32_t leader) {
  uint64_t ad = umma_desc(a_addr, a_lbo, a_sbo);
  uint64_t bd = umma_desc(b_addr, b_lbo, b_sbo);
  const uint64_t a_step = (2u * a_lbo) >> 4, b_step = (2u * b_lbo) >> 4;
  for (int k = 0; k < K / 16; ++k) {
    umma_bf16_p(d_tmem, ad, bd, idesc, (k > 0 || accumulate_first) ? 1u : 0u, leader);
    ad += a_step;
    bd += b_step;
  }
}

// TMEM -> registers: this thread's lane (row), 16 / 32 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, "
      "[%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------- A operand in TMEM, TMEM stores, split load/wait
// D[tmem] (+)= A[tmem] * B[smem]^T, one K=16 slice.  A lives in tensor memory: row i of the M=128 tile in lane i, two bf16
// per 32-bit column (element 2j in the low half of column j), i.e. a K=16 slice spans 8 columns.  Issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ts_p(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                               uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}
// D[128 x N] (+)= A[128 x K] (TMEM, K-major packed bf16) * B[N x K]^T (smem tile-native view), K multiple of 16
__device__ __forceinline__ void umma_gemm_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_addr, uint32_t b_lbo, uint32_t b_sbo,
                                             uint32_t idesc, int K, bool accumulate_first) {
  uint64_t bd = umma_desc(b_addr, b_lbo, b_sbo);
  const uint64_t b_step = (2u * b_lbo) >> 4;
  for (int k = 0; k < K / 16; ++k) {
    umma_bf16_ts(d_tmem, a_tmem + 8u * k, bd, idesc, (k > 0 || accumulate_first) ? 1u : 0u);
    bd += b_step;
  }
}
__device__ __forceinline__ void umma_chain_step(UmmaChain& c, uint32_t leader) {
  if (c.a_in_tmem) {
    umma_bf16_ts_p(c.d_tmem, c.a_tmem, c.bd, c.idesc, c.acc, leader);
    c.a_tmem += c.a_step;
  } else {
    umma_bf16_p(c.d_tmem, c.ad, c.bd, c.idesc, c.acc, leader);
    c.ad += c.a_step;
  }
  c.bd += c.b_step;
  c.acc = 1u;
  --c.nk;
}
// registers -> TMEM: this thread's lane, 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// TMEM -> registers without the wait: pair with tmem_wait_ld32 / tmem_wait_ld16, which carry the registers as in/out
// operands so that no use can be scheduled in front of the wait.
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, "
      "[%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld32(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld16(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
// plain arrive (count 1) on a CTA-local mbarrier
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// named barrier over `nthreads` threads (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------- small helpers
// byte offset of the 16-byte group (row r, column group cg) in a tile-native tile with R rows
__device__ __forceinline__ uint32_t tile_off16(int R, int r, int cg) { return (static_cast<uint32_t>(cg) * R + r) * 16u; }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(t);
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// numerically safe log(sigmoid(x)) = min(x,0) - log1p(exp(-|x|))
__device__ __forceinline__ float log_sigmoid(float x) { return fminf(x, 0.f) - log1pf(__expf(-fabsf(x))); }

constexpr float kLog2e = 1.4426950408889634f;

}  // namespace xhved

namespace xhved {

// ---------------------------------------------------------------- bulk store smem -> global (TMA unit)
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the bulk stores of this thread have finished READING shared memory
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ---------------------------------------------------------------- fp32 -> bf16 hi/lo staging
// 8 consecutive fp32 values -> one 16-byte bf16 group (hi) and its residual group (lo)
__device__ __forceinline__ void split8_hilo(const float* v, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
    const float2 b = unpack_bf16x2(h[i]);
    l[i] = pack_bf16x2(v[2 * i] - b.x, v[2 * i + 1] - b.y);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ uint4 pack8_bf16(const float* v) {
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
__device__ __forceinline__ void unpack8_bf16(const uint4& u, float* v) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x, v[1] = a.y, v[2] = b.x, v[3] = b.y, v[4] = c.x, v[5] = c.y, v[6] = d.x, v[7] = d.y;
}
// 16 fp32 values of row r, columns c0 .. c0+15, rounded to bf16 into a tile-native tile of R rows in GLOBAL memory: two
// 16-byte groups; the 32 rows of a warp cover 512 contiguous bytes per group
__device__ __forceinline__ void store16_bf16_tile(unsigned char* tile, int R, int r, int c0, const float* f) {
  *reinterpret_cast<uint4*>(tile + tile_off16(R, r, c0 / 8)) = pack8_bf16(f);
  *reinterpret_cast<uint4*>(tile + tile_off16(R, r, c0 / 8 + 1)) = pack8_bf16(f + 8);
}

// Stage a row-major fp32 matrix W[rows][cols] (global) as tile-native bf16 hi (+ optional lo) tiles with R = rows_tile
// rows (rows beyond `rows` are zero-filled).  Cooperative over the CTA.  cols % 8 == 0.
__device__ __forceinline__ void stage_weight_tile(const float* __restrict__ W, int rows, int cols, int rows_tile, unsigned char* hi,
                                                  unsigned char* lo) {
  const int groups = rows_tile * (cols / 8);
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    const int r = g % rows_tile, cg = g / rows_tile;
    float v[8];
    if (r < rows) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(W + static_cast<size_t>(r) * cols + cg * 8));
      const float4 b = __ldg(reinterpret_cast<const float4*>(W + static_cast<size_t>(r) * cols + cg * 8 + 4));
      v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
    }
    uint4 h, l;
    split8_hilo(v, h, l);
    *reinterpret_cast<uint4*>(hi + tile_off16(rows_tile, r, cg)) = h;
    if (lo) *reinterpret_cast<uint4*>(lo + tile_off16(rows_tile, r, cg)) = l;
  }
}

// 3-product high-precision GEMM: D = (Ahi + Alo)(Bhi + Blo)^T without the lo*lo term (~16 mantissa bits)
__device__ __forceinline__ void umma_gemm_hilo(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_hi,
                                               uint32_t b_lo, uint32_t b_lbo, uint32_t b_sbo, uint32_t idesc, int K) {
  umma_gemm(d_tmem, a_hi, a_lbo, a_sbo, b_hi, b_lbo, b_sbo, idesc, K, false);
  umma_gemm(d_tmem, a_lo, a_lbo, a_sbo, b_hi, b_lbo, b_sbo, idesc, K, true);
  umma_gemm(d_tmem, a_hi, a_lbo, a_sbo, b_lo, b_lbo, b_sbo, idesc, K, true);
}

}  // namespace xhved
