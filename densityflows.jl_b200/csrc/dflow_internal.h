// Internal structures shared by the host API (dflow_api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/dflow.h"

namespace dflow {

constexpr int DMAX = 64;       // data dimensions handled by the narrow (CUDA-core) path
constexpr int NMAX = 32;       // conditions
constexpr int MAX_DENSE = 6;   // Dense layers per conditioner
constexpr int LMAX = 64;       // leaf elements per chain
constexpr int HP_MAX = 64;     // widest hidden layer of the narrow path

// One conditioner MLP.  Widths are the true Flux widths; `op` is the padded row stride of the staged image.
struct DevNet {
  int depth;
  int has_bias;
  int w[MAX_DENSE + 1];
  int act[MAX_DENSE];
  int op[MAX_DENSE];   // padded output width of dense j in the staged image (multiple of 4)
  int p_w[MAX_DENSE];  // offsets into the packed parameter / gradient buffer
  int p_b[MAX_DENSE];  // -1 when no bias
  int s_w[MAX_DENSE];  // offsets (floats) into the element's staged block: W as [in][op]
  int s_b[MAX_DENSE];  // offsets (floats) of the padded bias [op] (zeros when no bias)
};

struct DevElem {
  int kind;
  int a;          // |axis_af|
  int nid;        // |axis_id| = d - a
  int nin;        // conditioner input width = n + nid
  int stage_off;  // offset (floats, multiple of 4) of this element's block in the staged image
  int stage_len;  // floats, multiple of 4
  int ck_off;     // adjoint checkpoint offset (floats per sample) of this element's transformed coordinates
  int ck_len;     // a (coupling) or d (normalisation); 0 when nothing has to be restored
  unsigned char af[DMAX];  // 0-based, caller order (src/Axes.jl:91)
  unsigned char id[DMAX];  // 0-based ascending complement (src/Axes.jl:88)
  DevNet s, t;
  // NORM elements keep [x_min(d) | x_max(d) | alpha, beta, ldj_const, 0] in their staged block.
};

struct DevChainHdr {
  int d, n, L;
  int hp;            // padded hidden width template (8/16/32/64)
  int P;             // parameter count
  int stage_total;   // floats
  int stage_max;     // largest element block (floats)
  int resident;      // 1: whole staged image kept in shared memory
  int amax4;         // max padded output width
  int max_depth;
  int ck_total;      // checkpoint floats per sample (adjoint workspace)
  int relu_only;     // 1: every hidden activation is relu / identity (register-resident kernels apply)
  int has_theta_range;
  float logpdf_c0;   // -(d*log(2pi))/2
  float theta_min[NMAX];
  float theta_inv[NMAX];  // 1/(max-min), 0 for a zero range (src/Data.jl:213-218)
  float theta_rng[NMAX];  // max - min
};

struct DevChain {
  DevChainHdr h;
  DevElem e[1];  // L entries
};

enum FwdMode {
  MODE_NORMALIZE = 0,   // x -> z, ldj
  MODE_LOGPDF = 1,      // x -> logp
  MODE_LOGPDF_SUM = 2,  // x -> sum logp (+ non-finite count)
  MODE_SAMPLE = 3,      // z -> x (in place or out of place), no ldj
  MODE_FORWARD_LDJ = 4, // z -> x, ldj
  MODE_SAMPLE_RNG = 5   // philox -> x
};

struct FwdArgs {
  const DevChain* chain;   // device
  const float* staged;     // device, padded weight image
  const float* x_in;
  const float* theta;       // (n,B) or null
  const float* theta_const; // n floats or null
  const int32_t* idx;
  float* x_out;
  float* aux_out;  // ldj / logp / loss[2]
  long long B;
  int mode;
  int flags;
  int chain_bytes;
  unsigned long long seed;
  unsigned int rng_offset;
  unsigned long long first_sample;
  int cb_wofs;  // constant-bank kernels: float offset of the staged weights inside the bank
  // tensor-product grid input (dflow_logpdf_grid, src/Flows.jl:287-331): sample b has x_k = grid_vals[off_k + (b / stride_k) % len_k]
  const float* grid_vals;       // device: the d coordinate vectors back to back
  const long long* grid_meta;   // device: per dimension [len_k, stride_k, off_k]
};

struct GradArgs {
  const DevChain* chain;
  const float* staged;
  const float* x_in;
  const float* theta;
  const int32_t* idx;
  float* loss_out;  // [2]
  float* grad_out;  // [P]
  float* ws;        // checkpoint workspace: grid * ck_total * blockDim floats
  long long B;
  float inv_btot;
  int flags;
  int chain_bytes;
  int smem_grad;  // 1: accumulate the whole gradient in shared memory, flush once per CTA
  // vector-Jacobian product with caller cotangents (dflow_vjp; rrule of src/affine/RNVP.jl:99-147 through the chain):
  const float* zbar;   // (d, B) cotangent of z, or null: loss seeds z * inv_btot
  const float* jbar;   // (B) cotangent of ln_det_jac, or null: -inv_btot
  float* xbar_out;     // (d, B) cotangent of x, or null
  float* thbar_out;    // (n, B) cotangent of theta, or null (then the theta rows of the first Dense are dropped)
};

struct TcPlan;
struct SmallPlan;

struct PrepackArgs {
  const DevChain* chain;
  const float* W;
  float* staged;
};

}  // namespace dflow

struct dflow_chain {
  int device = 0;
  int sm_count = 148;
  int max_smem_optin = 0;
  std::vector<unsigned char> host_chain;  // DevChain image
  dflow::DevChain* d_chain = nullptr;
  float* d_staged = nullptr;
  int chain_bytes = 0;
  // tuning
  int fwd_spt = 0, fwd_threads = 0, grad_threads = 0, grad_spt = 0, ctas_per_sm = 0;
  int grad_smem = 0;   // -1: the narrow adjoint adds its weight gradients straight to global memory (RED) instead of a
                       // per-CTA shared-memory accumulator (shared float atomics are compare-and-swap loops)
  int fwd_const = 0;   // -1: never use the constant-bank forward kernel (default: use it when the chain is eligible)
  int cbank_ok = 0;    // relu chain, hidden <= 32, every conditioner >= 2 Dense, descriptor + staged image <= 60 KB
  long long launches = 0;
  // host pipeline scratch (dflow_*_host)
  void* pipe = nullptr;
  // caller-owned scratch of the forward-type calls on the tensor-core kernels (dflow_chain_set_scratch)
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  long long* d_gridmeta = nullptr;          // [DMAX][3] grid description of the current dflow_logpdf_grid call
  long long h_gridmeta[3 * dflow::DMAX] = {0};
  // tensor-core plan (dflow_tc.cu): warp-specialised tcgen05 forward + adjoint; built for every eligible chain
  dflow::TcPlan* tcp = nullptr;
  int must_wide = 0;   // some hidden width > 64: the CUDA-core kernels cannot run this chain
  // small-minibatch epoch kernel (dflow_small.cu): non-null when the chain fits it (every Dense output <= 32)
  dflow::SmallPlan* small = nullptr;
  int epoch_kernel = 0;  // tuning: -1 keeps dflow_train_epoch on the per-minibatch (loss_grad + Adam) launches
  int tc_ws_budget_mb = 0;  // adjoint workspace cap in MiB (0: 24 GiB); larger batches are processed in macro-batches
  int tc_ts = 0;       // -1: hidden 32 / 64 conditioners stay on the warp-specialised pipeline (tc_net_kernel)
  int tc_dw_groups = 0; // > 0: cap on the staging-warp groups of the weight-gradient kernel (1 = all warps on every stage)
  int tc_dw_ts = 0;    // -1: weight gradients always through tc_dw_kernel (all operands staged in shared memory)
  int tc_fuse = 1;     // hidden <= 128 RealNVP layers run their s and t conditioners as one block-diagonal conditioner:
                       // 0 never, 1 in the train step (default), 2 in forward-type calls as well
  int tc_debug = 0;    // timing experiments (dflow_tc.cu, only with -DDFLOW_TC_EXPERIMENTS)
  int tc_debug_cluster = 0;  // CTA-pair variants (experiment builds only): 1 multicast weight stream, 2 cta_group::2
  int tc_mode = 0;     // 0: automatic (tensor cores iff must_wide), 1: force tensor cores, -1: force CUDA cores
  bool use_tc() const { return tcp && (must_wide || tc_mode > 0); }
  // Hidden 32 / 64 chains are eligible for both paths; the automatic choice follows scripts/tc_thresholds.py (one B200,
  // C3-like chain of 8 blocks at hidden 64, 4 blocks at hidden 32; CUDA-core / tensor-core ms per call, final round-2 kernels):
  //   hidden 64 log-density   32 768: 0.20 / 0.25    65 536: 0.38 / 0.26    1 Mi: 4.90 / 2.00
  //   hidden 64 train step     8 192: 1.02 / 0.70    65 536: 3.99 / 0.99    1 Mi: 55.3 / 10.3
  //   hidden 32 log-density   always the CUDA-core kernel (131 072: 0.16 / 0.16, 1 Mi: 0.66 / 0.68)
  //   hidden 32 train step     8 192: 0.33 / 0.26    32 768: 0.33 / 0.34    65 536: 0.63 / 0.41    1 Mi: 8.47 / 3.95
  int hidden_max = 0;
  bool use_tc_fwd(long long B) const {
    if (use_tc()) return true;
    return tcp && tc_mode == 0 && hidden_max == 64 && B >= 65536;
  }
  bool use_tc_grad(long long B) const {
    if (use_tc()) return true;
    return tcp && tc_mode == 0 && ((hidden_max == 64 || hidden_max == 32) && B >= 8192);
  }
  const dflow::DevChain* hc() const { return reinterpret_cast<const dflow::DevChain*>(host_chain.data()); }
  dflow::DevChain* hc() { return reinterpret_cast<dflow::DevChain*>(host_chain.data()); }
};

namespace dflow {
void set_error(const char* fmt, ...);
int launch_prepack(dflow_chain* c, const float* W, cudaStream_t st);
int launch_fwd(dflow_chain* c, FwdArgs& a, cudaStream_t st);
int launch_grad(dflow_chain* c, GradArgs& a, cudaStream_t st);
int launch_adam(float* W, const float* g, float* m, float* v, long long P, float lr, float b1, float b2, float eps,
                long long t, cudaStream_t st);
int launch_axpy2(float* acc, const float* v, cudaStream_t st);
void adam_beta_powers(float b1, float b2, long long t, float* b1t, float* b2t);
int launch_adam_pw(float* W, const float* g, float* m, float* v, long long P, float lr, float b1, float b2, float eps,
                   float b1t, float b2t, cudaStream_t st);
int launch_minmax(const float* x, int rows, long long B, float* mn, float* mx, cudaStream_t st);
int launch_shuffle(unsigned long long seed, long long n, long long first, long long count, const int32_t* base, int32_t* out,
                   cudaStream_t st);
}  // namespace dflow
