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

// p <- first step of a freshly created optimizer (zero slot variables) on gradient g (+ l2 * p).
// Adam (beta1 0.9, beta2 0.999, eps 1e-7): p - lr_t * m / (sqrt(v) + eps) with m = 0.1 g, v = 0.001 g^2, and
// sqrt(0.001 g^2) = sqrt(0.001) |g| - evaluated in that form (no underflow of g^2 for tiny g, as in the fp64 oracle) with
// one approximate division (2 ulp): the update is lr-sized, so its rounding is ~1e-10 absolute, far inside the 1e-5
// parity bound, and it keeps the fused backward kernels off the IEEE div / sqrt instruction sequences.
// Every operation is an explicit round-to-nearest intrinsic, so the compiler contracts nothing differently from one
// kernel to the next: the fused reduce+update kernels stay bit-identical to reduce followed by rm_sparse_opt_step.
__device__ __forceinline__ float opt_update(float p, float g, const OptParams& o) {
  g = __fmaf_rn(o.l2, p, g);
  if (o.opt == RM_OPT_ADAM) {
    const float q = __fdividef(__fmul_rn(0.1f, g), __fmaf_rn(0.0316227766f, fabsf(g), 1e-7f));
    return __fmaf_rn(-o.lr_t, q, p);
  } else if (o.opt == RM_OPT_ADAGRAD) {
    const float acc = __fmaf_rn(g, g, 0.1f);  // initial_accumulator_value = 0.1
    const float q = __fdividef(g, __fadd_rn(__fsqrt_rn(acc), 1e-7f));
    return __fmaf_rn(-o.lr, q, p);
  }
  return __fmaf_rn(-o.lr, g, p);  // gd / fresh momentum
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
