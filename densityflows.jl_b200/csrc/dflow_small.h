// Small-minibatch epoch kernel (dflow_small.cu): one persistent CTA runs every minibatch step of an epoch on-chip.
#pragma once
#include <cuda_runtime.h>

#include "dflow_internal.h"

namespace dflow {
struct SmallPlan;
int small_build_plan(dflow_chain* c);  // DFLOW_E_UNSUPPORTED when the chain does not fit the kernel
void small_free_plan(dflow_chain* c);
int small_train_epoch(dflow_chain* c, float* W, float* m, float* v, const float* x, const float* theta,
                      const int32_t* order, long long n, long long batchsize, float lr, float beta1, float beta2, float eps,
                      long long t0, int flags, float* loss2_out, cudaStream_t st);
}  // namespace dflow
