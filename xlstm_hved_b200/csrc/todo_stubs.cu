// TEMPORARY: entry points not implemented yet return -100 (removed as the kernels land).
#include "xhved.h"
extern "C" int xhved_mlstm_bwd(const void*, const void*, const void*, const float*, const float*, const void*, const void*, const float*,
                               const float*, const void*, const float*, int, int, int, int, float, float*, float*, float*, float*, float*,
                               float*, float*, float*, void*, float*, float*, void*) { return -100; }
extern "C" int xhved_mlstm_unpad_rows(const float*, int, int, int, int, float*, void*) { return -100; }
extern "C" int xhved_vil_pre_fwd(const float*, const xhved_vil_params*, const xhved_vil_shape*, void*, void*, void*, float*, float*, float*,
                                 float*, void*) { return -100; }
extern "C" int xhved_vil_post_fwd(const float*, const void*, const float*, const float*, const xhved_vil_params*, const xhved_vil_shape*,
                                  float*, void*) { return -100; }
extern "C" int xhved_vil_post_bwd(const float*, const void*, const float*, const float*, const xhved_vil_params*, const xhved_vil_shape*,
                                  void*, float*, float*, const xhved_vil_grads*, void*) { return -100; }
extern "C" int xhved_vil_pre_bwd(const float*, const float*, const float*, const float*, const float*, const float*, const float*,
                                 const float*, const float*, const xhved_vil_params*, const xhved_vil_shape*, float*,
                                 const xhved_vil_grads*, void*) { return -100; }
