#include "prof.cuh"

#include <atomic>
#include <mutex>
#include <vector>

#include "xhved.h"

namespace xhved {

static std::atomic<bool> g_on{false};
static std::mutex g_mu;
struct Rec {
  int id;
  cudaEvent_t a, b;
};
static std::vector<Rec> g_recs;
static std::vector<cudaEvent_t> g_pool;

static cudaEvent_t get_event() {
  if (!g_pool.empty()) {
    cudaEvent_t e = g_pool.back();
    g_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

bool prof_enabled() { return g_on.load(std::memory_order_relaxed); }

cudaEvent_t prof_begin(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_mu);
  cudaEvent_t e = get_event();
  cudaEventRecord(e, st);
  return e;
}
void prof_end(int id, cudaEvent_t begin, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_mu);
  cudaEvent_t e = get_event();
  cudaEventRecord(e, st);
  g_recs.push_back({id, begin, e});
}

static const char* kNames[K_COUNT] = {
    "poe_fwd", "poe_bwd", "reparam_fwd", "reparam_bwd", "vil_pre_fwd", "mlstm_chunk_state", "mlstm_state_scan",
    "mlstm_chunk_out", "vil_post_fwd", "vil_post_bwd", "mlstm_chunk_rstate", "mlstm_chunk_grad", "mlstm_gate_finish",
    "vil_pre_bwd_a", "vil_pre_bwd_b", "pack", "unpack", "norm_act_fwd", "norm_act_bwd"};

}  // namespace xhved

using namespace xhved;

extern "C" int xhved_profile_enable(int on) {
  g_on.store(on != 0);
  return 0;
}
extern "C" int xhved_profile_kernel_count(void) { return K_COUNT; }
extern "C" const char* xhved_profile_kernel_name(int id) { return (id >= 0 && id < K_COUNT) ? kNames[id] : ""; }

// Synchronises the device, sums elapsed ms and launch counts per kernel id since the last read, and resets.
extern "C" int xhved_profile_read(float* ms, int* launches, int n) {
  if (!ms || !launches || n < K_COUNT) return XHVED_ERR_BAD_ARG;
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return (int)e;
  std::lock_guard<std::mutex> lk(g_mu);
  for (int i = 0; i < n; ++i) ms[i] = 0.f, launches[i] = 0;
  for (const Rec& r : g_recs) {
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    ms[r.id] += t;
    launches[r.id] += 1;
    g_pool.push_back(r.a);
    g_pool.push_back(r.b);
  }
  g_recs.clear();
  return 0;
}
