// One ViL block behind ONE C entry per direction (host code only): K2 -> cell -> K3 and their backward, enqueued on the
// caller's stream into caller-allocated blobs.  At the reference's real batch (one volume, train.py:50) a block is ~20
// launches of 5-30 us each: issued one by one through a foreign-function binding, with ~25 separate buffers, the host side
// costs several times what the kernels take.  Here the buffers are carved out of two blobs whose sizes
// xhved_vil_block_workspace reports, and the whole sequence is two calls.  Nothing is allocated, no state is kept: the
// entry points are re-entrant per stream (the reference runs replicas from several threads, train.py:148-151).
#include <cuda_runtime.h>
#include <stdint.h>

#include "xhved.h"

namespace {

constexpr int64_t kAlign = 256;
inline int64_t up(int64_t n) { return (n + kAlign - 1) / kAlign * kAlign; }

struct Carver {
  unsigned char* base;
  int64_t off = 0;
  explicit Carver(void* b) : base(static_cast<unsigned char*>(b)) {}
  template <typename T>
  T* take(int64_t bytes) {
    T* p = reinterpret_cast<T*>(base + off);     // with base == nullptr this only measures
    off += up(bytes);
    return p;
  }
};

// everything the backward needs from the forward, plus the forward's own scratch (dstate / g / amax are reused by the backward)
struct Saved {
  void *q, *k, *v, *h, *states, *act, *z, *xm;      // act / z / xm: bf16 token tiles
  float *ig, *fg, *m, *den, *ws_dstate, *ws_g, *ws_amax, *m_prev;
  int64_t bytes;
  Saved(void* blob, const xhved_vil_workspace& w) {
    Carver c(blob);
    const xhved_mlstm_workspace& cw = w.cell;
    q = c.take<void>(cw.tile_bytes), k = c.take<void>(cw.tile_bytes), v = c.take<void>(cw.tile_bytes), h = c.take<void>(cw.tile_bytes);
    states = c.take<void>(cw.states_bytes);
    ig = c.take<float>(cw.row_bytes), fg = c.take<float>(cw.row_bytes), m = c.take<float>(cw.row_bytes), den = c.take<float>(cw.row_bytes);
    ws_dstate = c.take<float>(cw.dstate_bytes), ws_g = c.take<float>(cw.chunk_bytes), ws_amax = c.take<float>(cw.chunk_bytes);
    m_prev = c.take<float>(cw.chunk_bytes);
    act = c.take<void>(w.token_tile_bytes), z = c.take<void>(w.token_tile_bytes), xm = c.take<void>(w.token_tile_bytes);
    bytes = c.off;
  }
};

struct Scratch {
  float *flat, *dig, *dfg, *mu_next, *ws_dc;
  void *dh_tiles, *rstates, *d_act, *dz, *dq, *dk, *dv, *ws_dconv, *ws_dxmv;      // bf16 tiles
  int64_t flat_bytes, bytes;
  Scratch(void* blob, const xhved_vil_workspace& w, int replicas) {
    Carver c(blob);
    const xhved_mlstm_workspace& cw = w.cell;
    flat_bytes = static_cast<int64_t>(replicas) * w.grad_replica_stride * 4;
    flat = c.take<float>(flat_bytes);
    dh_tiles = c.take<void>(cw.tile_bytes), rstates = c.take<void>(cw.states_bytes);
    d_act = c.take<void>(w.token_tile_bytes), dz = c.take<void>(w.token_tile_bytes);
    dq = c.take<void>(cw.tile_bytes), dk = c.take<void>(cw.tile_bytes), dv = c.take<void>(cw.tile_bytes);
    dig = c.take<float>(cw.row_bytes), dfg = c.take<float>(cw.row_bytes), mu_next = c.take<float>(cw.chunk_bytes);
    ws_dc = c.take<float>(cw.row_bytes);
    ws_dconv = c.take<void>(w.token_tile_bytes), ws_dxmv = c.take<void>(w.token_tile_bytes);
    bytes = c.off;
  }
};

// the 14 parameter gradients inside one replica, in xhved_vil_params order (= ops.VIL_PARAM_KEYS)
xhved_vil_grads grads_in(float* flat, int C) {
  const int64_t E = 2 * C;
  const int64_t sizes[14] = {C, 2 * E * C, E * 4, E, E * 4, E * 4, E * 4, 4 * 3 * E, 4, 4 * 3 * E, 4, E, E, C * E};
  float* p[14];
  int64_t off = 0;
  for (int i = 0; i < 14; ++i) {
    p[i] = flat + off;
    off += sizes[i];
  }
  xhved_vil_grads g;
  g.norm_weight = p[0], g.proj_up_weight = p[1], g.conv_weight = p[2], g.conv_bias = p[3], g.q_weight = p[4], g.k_weight = p[5];
  g.v_weight = p[6], g.igate_weight = p[7], g.igate_bias = p[8], g.fgate_weight = p[9], g.fgate_bias = p[10];
  g.outnorm_weight = p[11], g.learnable_skip = p[12], g.proj_down_weight = p[13];
  return g;
}
int64_t param_count(int C) {
  const int64_t E = 2 * C;
  return C + 2 * E * C + E * 4 + E + 3 * E * 4 + 2 * (4 * 3 * E + 4) + E + E + static_cast<int64_t>(C) * E;
}

}  // namespace

extern "C" int xhved_vil_block_workspace(int B, int S, int C, int grad_replicas, int64_t* saved_bytes, int64_t* scratch_bytes,
                                         int64_t* n_param_grads) {
  xhved_vil_workspace w;
  if (int rc = xhved_vil_workspace_query(B, S, C, &w)) return rc;
  if (grad_replicas < 1) return XHVED_ERR_BAD_ARG;
  if (saved_bytes) *saved_bytes = Saved(nullptr, w).bytes;
  if (scratch_bytes) *scratch_bytes = Scratch(nullptr, w, grad_replicas).bytes;
  if (n_param_grads) *n_param_grads = param_count(C);
  return 0;
}

extern "C" int xhved_vil_block_fwd(const float* x, const xhved_vil_params* p, const xhved_vil_shape* sh, float eps, void* saved,
                                   float* y, void* stream) {
  if (!x || !p || !sh || !saved || !y) return XHVED_ERR_BAD_ARG;
  xhved_vil_workspace w;
  if (int rc = xhved_vil_workspace_query(sh->B, sh->S, sh->C, &w)) return rc;
  Saved s(saved, w);
  if (int rc = xhved_vil_pre_fwd(x, p, sh, s.q, s.k, s.v, s.ig, s.fg, s.act, s.z, s.xm, stream)) return rc;
  if (int rc = xhved_mlstm_fwd(s.q, s.k, s.v, s.ig, s.fg, 4 * sh->B, w.cell.nc, sh->C / 2, w.cell.dhp, eps, s.h, s.m, s.den, s.ws_dstate,
                               s.ws_g, s.ws_amax, s.states, s.m_prev, stream))
    return rc;
  return xhved_vil_post_fwd(x, s.h, s.act, s.z, p, sh, y, stream);
}

extern "C" int xhved_vil_block_bwd(const float* x, const float* dy, const xhved_vil_params* p, const xhved_vil_shape* sh, float eps,
                                   void* saved, void* scratch, float* dx, float* param_grads, void* stream) {
  if (!x || !dy || !p || !sh || !saved || !scratch || !dx || !param_grads) return XHVED_ERR_BAD_ARG;
  xhved_vil_workspace w;
  if (int rc = xhved_vil_workspace_query(sh->B, sh->S, sh->C, &w)) return rc;
  const int R = sh->grad_replicas < 1 ? 1 : sh->grad_replicas;
  if (sh->grad_replica_stride != w.grad_replica_stride) return XHVED_ERR_BAD_ARG;
  Saved s(saved, w);
  Scratch t(scratch, w, R);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // the parameter gradients are accumulated with atomics into R zero-filled replicas (xhved_vil_shape)
  cudaError_t e = cudaMemsetAsync(t.flat, 0, t.flat_bytes, st);
  if (e != cudaSuccess) return (int)e;
  xhved_vil_grads g = grads_in(t.flat, sh->C);
  if (int rc = xhved_vil_post_bwd(dy, s.h, s.act, s.z, p, sh, t.dh_tiles, t.d_act, t.dz, &g, stream)) return rc;
  if (int rc = xhved_mlstm_bwd(s.q, s.k, s.v, s.ig, s.fg, s.h, t.dh_tiles, s.m, s.den, s.states, s.m_prev, 4 * sh->B, w.cell.nc, sh->C / 2,
                               w.cell.dhp, eps, t.dq, t.dk, t.dv, t.dig, t.dfg, s.ws_dstate, s.ws_g, s.ws_amax, t.rstates, t.mu_next, t.ws_dc,
                               stream))
    return rc;
  if (int rc = xhved_vil_pre_bwd(x, dy, s.xm, s.q, s.k, s.v, t.dq, t.dk, t.dv, t.dig, t.dfg, t.d_act, t.dz, p, sh, dx, &g, t.ws_dconv,
                                 t.ws_dxmv, stream))
    return rc;
  return xhved_reduce_replicas(t.flat, R, w.grad_replica_stride, param_count(sh->C), param_grads, stream);
}
