// TEMPORARY: entry points not implemented yet return -100 (removed as the kernels land).
#include "xhved.h"
extern "C" int xhved_vil_post_bwd(const float*, const void*, const float*, const float*, const xhved_vil_params*, const xhved_vil_shape*,
                                  void*, float*, float*, const xhved_vil_grads*, void*) { return -100; }
extern "C" int xhved_vil_pre_bwd(const float*, const float*, const float*, const float*, const float*, const float*, const float*,
                                 const float*, const float*, const xhved_vil_params*, const xhved_vil_shape*, float*,
                                 const xhved_vil_grads*, void*) { return -100; }
