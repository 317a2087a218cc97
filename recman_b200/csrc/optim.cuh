// Stateless first-step optimizer update shared by optim.cu and the fused reduce+update kernels (scatter.cu).
#pragma once
#include <math.h>

#include "common.cuh"

namespace rm {

struct OptParams {
  int opt;
  float lr;    // plain learning rate (adagrad, gd)
  float lr_t;  // adam: lr*sqrt(1-b2)/(1-b1)
  float l2;
};

__device__ __forceinline__ float opt_update(float p, float g, const OptParams& o) {
  g += o.l2 * p;
  if (o.opt == RM_OPT_ADAM) {
    const float m = 0.1f * g;             // (1 - beta1) * g, beta1 = 0.9
    const float v = 0.001f * g * g;       // (1 - beta2) * g^2, beta2 = 0.999
    return p - o.lr_t * m / (sqrtf(v) + 1e-7f);
  } else if (o.opt == RM_OPT_ADAGRAD) {
    const float acc = 0.1f + g * g;       // initial_accumulator_value = 0.1
    return p - o.lr * g / (sqrtf(acc) + 1e-7f);
  }
  return p - o.lr * g;                    // gd / fresh momentum
}

static inline int make_params(int opt, float lr, float l2, OptParams* o) {
  if (opt != RM_OPT_ADAM && opt != RM_OPT_ADAGRAD && opt != RM_OPT_GD) {
    set_error("unknown optimizer kind %d", opt);
    return RM_E_INVALID;
  }
  o->opt = opt;
  o->lr = lr;
  o->lr_t = (float)((double)lr * sqrt(1.0 - 0.999) / (1.0 - 0.9));
  o->l2 = l2;
  return 0;
}


}  // namespace rm
