// Host-only size queries of the C ABI (include/xhved.h): how large the caller-allocated buffers must be.
#include <stdint.h>

#include "xhved.h"

namespace {
int padded_head_dim(int dh) {
  for (int p : {16, 32, 64, 128})
    if (dh <= p) return p;
  return 0;
}
}  // namespace

extern "C" int xhved_mlstm_workspace_query(int BH, int S, int dh, xhved_mlstm_workspace* out) {
  if (!out || BH <= 0 || S <= 0 || dh <= 0) return XHVED_ERR_BAD_SHAPE;
  const int dhp = padded_head_dim(dh);
  if (!dhp) return XHVED_ERR_UNSUPPORTED_DH;
  const int64_t nc = (S + 127) / 128, nt = static_cast<int64_t>(BH) * nc, ne = dhp + 16;
  out->nc = static_cast<int>(nc);
  out->dhp = dhp;
  out->tile_bytes = nt * 128 * dhp * 2;
  out->row_bytes = nt * 128 * 4;
  out->dstate_bytes = nt * dhp * ne * 4;
  out->chunk_bytes = nt * 4;
  out->states_bytes = nt * 2 * dhp * ne * 2;
  return 0;
}

extern "C" int xhved_vil_workspace_query(int B, int S, int C, xhved_vil_workspace* out) {
  if (!out || B <= 0 || S <= 0) return XHVED_ERR_BAD_SHAPE;
  if (C != 16 && C != 32 && C != 64) return XHVED_ERR_UNSUPPORTED_DIM;
  const int E = 2 * C;
  if (int rc = xhved_mlstm_workspace_query(4 * B, S, E / 4, &out->cell)) return rc;
  out->token_tile_bytes = static_cast<int64_t>(B) * out->cell.nc * E * 128 * 2;
  // norm, proj_up, conv w/b, q/k/v, igate w/b, fgate w/b, outnorm, skip, proj_down (ops.VIL_PARAM_KEYS order)
  const int64_t params = C + 2LL * E * C + E * 4 + E + 3LL * E * 4 + 2 * (4LL * 3 * E + 4) + E + E + static_cast<int64_t>(C) * E;
  out->grad_replica_stride = (params + 31) / 32 * 32;
  return 0;
}
