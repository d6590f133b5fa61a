// Optional per-kernel timing (CUDA events on the launching stream).  Off by default: when off, XHVED_PROF
// costs one relaxed load.  bench.py switches it on for a separate pass to attribute time to kernels.
#pragma once
#include <cuda_runtime.h>

namespace xhved {

enum KernelId {
  K_POE_FWD = 0, K_POE_BWD, K_REPARAM_FWD, K_REPARAM_BWD,
  K_VIL_PRE_FWD, K_CHUNK_STATE, K_STATE_SCAN, K_CHUNK_OUT, K_VIL_POST_FWD,
  K_VIL_POST_BWD, K_CHUNK_RSTATE, K_CHUNK_GRAD, K_GATE_FINISH, K_VIL_PRE_BWD_A, K_VIL_PRE_BWD_B,
  K_PACK, K_UNPACK, K_NORM_FWD, K_NORM_BWD, K_COUNT
};

bool prof_enabled();
cudaEvent_t prof_begin(cudaStream_t st);
void prof_end(int id, cudaEvent_t begin, cudaStream_t st);

// The begin event lives in the scope object (not in a per-kernel-id global), so scopes opened concurrently from several
// host threads -- the reference runs replicas from several threads, train.py:148-151 -- cannot overwrite each other's.
struct ProfScope {
  int id;
  cudaStream_t st;
  bool on;
  cudaEvent_t begin;
  ProfScope(int id_, cudaStream_t st_) : id(id_), st(st_), on(prof_enabled()), begin(nullptr) {
    if (on) begin = prof_begin(st);
  }
  ~ProfScope() {
    if (on) prof_end(id, begin, st);
  }
};

}  // namespace xhved
