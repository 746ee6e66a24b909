// Wide-conditioner (tcgen05) path: plan structures shared by dflow_api.cu and dflow_wide.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "dflow_internal.h"

namespace dflow {

// One conditioner Dense(in,h,relu) -> Dense(h,h,relu) -> Dense(h,a) inside the wide weight image.
struct WideNet {
  int p_w[3], p_b[3];  // offsets into the packed parameter buffer
  long long img_off;   // start of this net's block in the wide image (floats)
  int b1, b2, b3;      // bias offsets inside the block
  int w1, w2, w3;      // chunked hi/lo operand images inside the block
};

struct WideLayer {
  int is_coupling;  // 0: NormalizationLayer
  int has_s;        // RNVP: 1, NICE: 0
  int h;            // hidden width (multiple of 32, <= 256 or a multiple of 256)
  int nin, kinp;    // conditioner inputs, padded to a multiple of 8
  int a, a16;       // transformed dims, padded to a multiple of 16
  int norm_off;     // NORM: offset of [x_min | x_max | alpha beta c] in the staged image
  unsigned char af[DMAX], id[DMAX];
  WideNet net[2];   // [0] = s_net, [1] = t_net
};

struct WidePlan {
  std::vector<WideLayer> layers;  // chain order
  WideLayer* d_layers = nullptr;
  float* d_img = nullptr;
  size_t img_floats = 0;
  // grow-only scratch for entry points that need a working copy (logpdf, gathers)
  float* d_scratch = nullptr;
  size_t scratch_floats = 0;
};

size_t wide_layer_smem_bytes(const WideLayer& Ld);
int wide_prepack(dflow_chain* c, const float* W, cudaStream_t st);
int wide_run_chain(dflow_chain* c, float* x, const float* theta, const float* theta_const, float* ldj, long long B,
                   int sampling, int flags, cudaStream_t st);
int wide_logpdf(dflow_chain* c, const float* z, const float* ldj, long long B, float* out, float* sum2, cudaStream_t st);
int wide_gather(dflow_chain* c, const float* src, const int32_t* idx, long long B, int rows, float* dst, cudaStream_t st);
int wide_philox(dflow_chain* c, float* z, long long B, unsigned long long seed, unsigned int offset,
                unsigned long long first, cudaStream_t st);

}  // namespace dflow
