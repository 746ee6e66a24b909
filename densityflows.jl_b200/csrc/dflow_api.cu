// C ABI of libdflow.so (include/dflow.h): handle construction, validation, launches, host-buffer pipelines.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "dflow_internal.h"
#include "dflow_small.h"
#include "dflow_tc.h"

namespace dflow {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

#define CKA(call)                                                                         \
  do {                                                                                    \
    cudaError_t _e = (call);                                                              \
    if (_e != cudaSuccess) {                                                              \
      set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DFLOW_E_CUDA;                                                                \
    }                                                                                     \
  } while (0)

static int round_hp(int w) {
  if (w <= 16) return 16;
  if (w <= 32) return 32;
  if (w <= 64) return 64;
  return -1;
}

static int round_out(int a) {  // padded width of the last Dense: 4, 8, 16, 32, 64
  int p = 4;
  while (p < a) p <<= 1;
  return p;
}

// Fill one DevNet from its descriptor; returns 0 or an error code.
static int fill_net(const dflow_net_desc& nd, int expect_in, int expect_out, DevNet& net, int& P, int& hidden_max,
                    const char* what, int elem) {
  memset(&net, 0, sizeof(net));
  if (nd.depth < 1 || nd.depth > MAX_DENSE) {
    set_error("element %d %s: depth %d outside [1,%d]", elem, what, nd.depth, MAX_DENSE);
    return nd.depth > MAX_DENSE ? DFLOW_E_UNSUPPORTED : DFLOW_E_INVALID_ARG;
  }
  if (!nd.widths || !nd.acts) {
    set_error("element %d %s: widths/acts missing", elem, what);
    return DFLOW_E_INVALID_ARG;
  }
  net.depth = nd.depth;
  net.has_bias = nd.has_bias ? 1 : 0;
  for (int j = 0; j <= nd.depth; ++j) {
    net.w[j] = nd.widths[j];
    if (net.w[j] < 1) {
      set_error("element %d %s: width[%d]=%d", elem, what, j, net.w[j]);
      return DFLOW_E_INVALID_ARG;
    }
  }
  if (net.w[0] != expect_in || net.w[nd.depth] != expect_out) {
    // CouplingLayer: input_dim = length(axis_nn), output_dim = length(axis_af) (src/Layers.jl:126-127)
    set_error("element %d %s: net maps %d->%d but the axes need %d->%d", elem, what, net.w[0], net.w[nd.depth],
              expect_in, expect_out);
    return DFLOW_E_INVALID_ARG;
  }
  for (int j = 0; j < nd.depth; ++j) {
    net.act[j] = nd.acts[j];
    if (net.act[j] < DFLOW_ACT_IDENTITY || net.act[j] > DFLOW_ACT_SIGMOID) {
      set_error("element %d %s: unsupported activation code %d", elem, what, net.act[j]);
      return DFLOW_E_UNSUPPORTED;
    }
    if (j < nd.depth - 1) hidden_max = std::max(hidden_max, net.w[j + 1]);
    net.p_w[j] = P;
    P += net.w[j] * net.w[j + 1];
    if (net.has_bias) {
      net.p_b[j] = P;
      P += net.w[j + 1];
    } else {
      net.p_b[j] = -1;
    }
  }
  return DFLOW_OK;
}

static void layout_net(DevNet& net, int hp, int& off) {
  for (int j = 0; j < net.depth; ++j) {
    net.op[j] = (j < net.depth - 1) ? hp : round_out(net.w[j + 1]);
    net.s_w[j] = off;
    off += (j == 0 ? net.w[j] : hp) * net.op[j];  // hidden Dense rows are zero-padded to hp (unrolled kernels)
    net.s_b[j] = off;
    off += net.op[j];
  }
}

}  // namespace dflow

using namespace dflow;

// Forward-type entry points of chains routed to the tensor-core kernels (dflow_tc.cu works on its own tile-blocked state).
static int wide_fwd(dflow_chain* c, const float* W, FwdArgs& a, void* stream) {
  if (a.B == 0) return DFLOW_OK;
  return tc_fwd(c, W, a, (cudaStream_t)stream);
}

extern "C" {

struct HostPipe;
static void pipe_free(HostPipe* p);

int dflow_version(void) { return DFLOW_VERSION; }
const char* dflow_last_error(void) { return g_err; }

int dflow_chain_create(const dflow_chain_desc* desc, dflow_chain** out) {
  if (!desc || !out) {
    set_error("null argument");
    return DFLOW_E_INVALID_ARG;
  }
  *out = nullptr;
  const int d = desc->d, n = desc->n, L = desc->n_elems;
  if (d < 1 || n < 0 || L < 1 || !desc->elems) {
    set_error("bad chain shape d=%d n=%d n_elems=%d", d, n, L);
    return DFLOW_E_INVALID_ARG;
  }
  if (d > DMAX || n > NMAX || L > LMAX) {
    set_error("chain shape d=%d n=%d n_elems=%d exceeds the narrow path limits (%d,%d,%d)", d, n, L, DMAX, NMAX, LMAX);
    return DFLOW_E_UNSUPPORTED;
  }
  std::vector<unsigned char> img(sizeof(DevChainHdr) + sizeof(DevElem) * (size_t)L, 0);
  DevChain* C = reinterpret_cast<DevChain*>(img.data());
  DevChainHdr& H = C->h;
  H.d = d;
  H.n = n;
  H.L = L;
  H.logpdf_c0 = -((float)d * 1.8378770664093453f) / 2.0f;  // -(d*log2π)/2 in Float32
  int P = 0, hidden_max = 1, amax4 = 4, max_depth = 1;
  bool relu_only = true;
  for (int ei = 0; ei < L; ++ei) {
    const dflow_elem_desc& ed = desc->elems[ei];
    DevElem& E = C->e[ei];
    E.kind = ed.kind;
    if (ed.kind == DFLOW_ELEM_NORM) {
      if (!ed.x_min || !ed.x_max) {
        set_error("element %d: NormalizationLayer needs x_min/x_max", ei);
        return DFLOW_E_INVALID_ARG;
      }
      if (!(ed.beta > ed.alpha)) {  // src/norm/Normalization.jl:55
        set_error("element %d: bounds of the normalisation need beta > alpha", ei);
        return DFLOW_E_INVALID_ARG;
      }
      continue;
    }
    if (ed.kind != DFLOW_ELEM_RNVP && ed.kind != DFLOW_ELEM_NICE) {
      set_error("element %d: unknown kind %d", ei, ed.kind);
      return DFLOW_E_INVALID_ARG;
    }
    const int a = ed.n_af;
    if (a < 1 || a > d || !ed.axis_af) {
      set_error("element %d: bad axis_af length %d", ei, a);
      return DFLOW_E_INVALID_ARG;
    }
    bool seen[DMAX] = {false};
    for (int j = 0; j < a; ++j) {
      const int k = ed.axis_af[j];
      if (k < 0 || k >= d) {  // src/Axes.jl:85
        set_error("element %d: the mask cannot contain values higher than the dimension (axis_af[%d]=%d)", ei, j, k);
        return DFLOW_E_INVALID_ARG;
      }
      if (seen[k]) {
        set_error("element %d: duplicate index %d in axis_af", ei, k);
        return DFLOW_E_INVALID_ARG;
      }
      seen[k] = true;
      E.af[j] = (unsigned char)k;
    }
    int nid = 0;
    if (ed.axis_id) {
      // explicit order (reverse(axes) keeps the former axis_af order, src/Axes.jl:129-135)
      nid = ed.n_id;
      if (nid != d - a) {
        set_error("element %d: axis_id has %d entries, expected %d", ei, nid, d - a);
        return DFLOW_E_INVALID_ARG;
      }
      bool seen_id[DMAX] = {false};
      for (int j = 0; j < nid; ++j) {
        const int k = ed.axis_id[j];
        if (k < 0 || k >= d || seen[k] || seen_id[k]) {
          set_error("element %d: axis_id[%d]=%d is out of range, duplicated or also in axis_af", ei, j, k);
          return DFLOW_E_INVALID_ARG;
        }
        seen_id[k] = true;
        E.id[j] = (unsigned char)k;
      }
    } else {
      // axis_id = findall(x -> !(x in mask), 1:d) (ascending), src/Axes.jl:88
      for (int k = 0; k < d; ++k)
        if (!seen[k]) E.id[nid++] = (unsigned char)k;
    }
    E.a = a;
    E.nid = nid;
    E.nin = n + nid;  // length(axis_nn), src/Axes.jl:98
    if (E.nin < 1) {
      set_error("element %d: conditioner has no inputs (n=0 and every dimension transformed)", ei);
      return DFLOW_E_UNSUPPORTED;
    }
    int rc;
    if (ed.kind == DFLOW_ELEM_RNVP) {
      rc = fill_net(ed.s_net, E.nin, a, E.s, P, hidden_max, "s_net", ei);
      if (rc) return rc;
      max_depth = std::max(max_depth, E.s.depth);
    }
    rc = fill_net(ed.t_net, E.nin, a, E.t, P, hidden_max, "t_net", ei);
    if (rc) return rc;
    max_depth = std::max(max_depth, E.t.depth);
    amax4 = std::max(amax4, round_out(a));
  }
  int hp = round_hp(hidden_max);
  const bool wide = hp < 0;  // hidden > 64: tcgen05 path (dflow_tc.cu)
  // the tensor-core kernels cover the reference's default conditioner only: Dense(in,h,relu), Dense(h,h,relu), Dense(h,a)
  bool tc_eligible = true;
  int has_coupling = 0;
  for (int ei = 0; ei < L; ++ei) {
    const DevElem& E = C->e[ei];
    if (E.kind == DFLOW_ELEM_NORM) continue;
    has_coupling = 1;
    for (int ni = (E.kind == DFLOW_ELEM_RNVP ? 0 : 1); ni < 2; ++ni) {
      const DevNet& net = ni == 0 ? E.s : E.t;
      const int h = net.w[1];
      // three Dense layers of equal hidden width, any of relu / tanh / sigmoid / identity on the hidden layers, identity on the
      // output (src/Layers.jl:33-50 with n_sublayers = 2), with or without bias
      const bool ok = net.depth == 3 && net.w[2] == h && net.act[2] == DFLOW_ACT_IDENTITY && h % 32 == 0 && h >= 32 &&
                      (h <= 256 || h == 512) && E.nin <= 64 && E.a <= 32 && h == E.t.w[1];
      if (!ok) {
        tc_eligible = false;
        if (wide) {
          set_error("element %d: wide conditioners (hidden > 64) must be Dense(in,h,σ)->Dense(h,h,σ)->Dense(h,a) (n_sublayers = 2, "
                    "identity output), h a multiple of 32 up to 256 or 512, <= 64 inputs and <= 32 outputs", ei);
          return DFLOW_E_UNSUPPORTED;
        }
      }
    }
  }
  if (!has_coupling) tc_eligible = false;
  if (wide) hp = 64;  // narrow-path fields stay consistent but are not used
  H.hp = hp;
  H.P = P;
  H.amax4 = amax4;
  H.max_depth = max_depth;
  for (int ei = 0; ei < L; ++ei) {
    const DevElem& E = C->e[ei];
    if (E.kind == DFLOW_ELEM_NORM) continue;
    for (int ni = (E.kind == DFLOW_ELEM_RNVP ? 0 : 1); ni < 2; ++ni) {
      const DevNet& net = ni == 0 ? E.s : E.t;
      for (int j = 0; j + 1 < net.depth; ++j)
        if (net.act[j] != DFLOW_ACT_RELU && net.act[j] != DFLOW_ACT_IDENTITY) relu_only = false;
    }
  }
  H.relu_only = relu_only ? 1 : 0;
  // staged image layout
  int total = 0, smax = 4;
  for (int ei = 0; ei < L; ++ei) {
    DevElem& E = C->e[ei];
    E.stage_off = total;
    int off = 0;
    if (E.kind == DFLOW_ELEM_NORM) {
      off = 2 * d + 4;
    } else if (!wide) {
      if (E.kind == DFLOW_ELEM_RNVP) layout_net(E.s, hp, off);
      layout_net(E.t, hp, off);
    }
    off = (off + 3) & ~3;
    E.stage_len = off;
    total += off;
    smax = std::max(smax, off);
  }
  H.stage_total = total;
  // adjoint checkpoints: the layer inputs' transformed coordinates are saved in the forward sweep and restored
  // bit-exactly in the reverse sweep (an inverse-map reconstruction drifts by ~1e-6 and flips ReLU kinks)
  int ck = 0;
  for (int ei = 0; ei < L; ++ei) {
    DevElem& E = C->e[ei];
    E.ck_off = ck;
    E.ck_len = (E.kind == DFLOW_ELEM_NORM) ? (ei == L - 1 ? 0 : d) : E.a;
    ck += E.ck_len;
  }
  H.ck_total = ck;
  H.stage_max = smax;

  dflow_chain* c = new (std::nothrow) dflow_chain();
  if (!c) {
    set_error("out of host memory");
    return DFLOW_E_NOMEM;
  }
  cudaDeviceProp prop;
  if (cudaGetDevice(&c->device) != cudaSuccess || cudaGetDeviceProperties(&prop, c->device) != cudaSuccess) {
    set_error("no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
    delete c;
    return DFLOW_E_CUDA;
  }
  c->sm_count = prop.multiProcessorCount;
  c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  // keep the whole staged image in shared memory when it leaves room for >= 2 CTAs of columns
  H.resident = (total * 4 <= 96 * 1024) ? 1 : 0;
  c->chain_bytes = (int)((img.size() + 15) & ~(size_t)15);
  img.resize(c->chain_bytes, 0);
  c->host_chain = img;
  {
    // constant-bank forward kernel (dflow_chain_kernels.cuh): relu conditioners of >= 2 Dense, hidden template 16 / 32,
    // descriptor + staged image inside the 60 KB bank
    bool ok = !wide && relu_only && (hp == 16 || hp == 32) && (size_t)c->chain_bytes + (size_t)total * 4 <= 60 * 1024;
    for (int ei = 0; ei < L && ok; ++ei) {
      const DevElem& E = C->e[ei];
      if (E.kind == DFLOW_ELEM_NORM) continue;
      if (E.kind == DFLOW_ELEM_RNVP && E.s.depth < 2) ok = false;
      if (E.t.depth < 2) ok = false;
    }
    c->cbank_ok = ok ? 1 : 0;
  }
  if (desc->theta_min && desc->theta_max) dflow_chain_set_theta_range(c, desc->theta_min, desc->theta_max);

  cudaError_t e1 = cudaMalloc(&c->d_chain, c->chain_bytes);
  cudaError_t e2 = cudaMalloc(&c->d_staged, std::max(total, 4) * sizeof(float));
  if (e1 == cudaSuccess) e1 = cudaMalloc(&c->d_gridmeta, sizeof(long long) * 3 * DMAX);
  if (e1 != cudaSuccess || e2 != cudaSuccess) {
    set_error("cudaMalloc failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    dflow_chain_destroy(c);
    return DFLOW_E_NOMEM;
  }
  // norm blocks of the staged image are static: [x_min | x_max | alpha beta ldj_const 0]
  std::vector<float> st(std::max(total, 4), 0.0f);
  for (int ei = 0; ei < L; ++ei) {
    const dflow_elem_desc& ed = desc->elems[ei];
    const DevElem& E = c->hc()->e[ei];
    if (E.kind != DFLOW_ELEM_NORM) continue;
    float* b = st.data() + E.stage_off;
    const float delta = ed.beta - ed.alpha;
    float csum = 0.0f;  // sum(log.(x_diff ./ δ)) in Float32, src/norm/Normalization.jl:73
    for (int k = 0; k < d; ++k) {
      b[k] = ed.x_min[k];
      b[d + k] = ed.x_max[k];
      csum += logf((ed.x_max[k] - ed.x_min[k]) / delta);
    }
    b[2 * d] = ed.alpha;
    b[2 * d + 1] = ed.beta;
    b[2 * d + 2] = csum;
  }
  if (cudaMemcpy(c->d_staged, st.data(), st.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(c->d_chain, c->host_chain.data(), c->chain_bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("cudaMemcpy failed: %s", cudaGetErrorString(cudaGetLastError()));
    dflow_chain_destroy(c);
    return DFLOW_E_CUDA;
  }
  c->must_wide = wide ? 1 : 0;
  c->hidden_max = hidden_max;
  if (tc_eligible) {
    int rc = tc_build_plan(c);
    if (rc && wide) {
      dflow_chain_destroy(c);
      return rc;
    }
    if (rc) tc_free_plan(c);  // a narrow chain simply stays on the CUDA-core path
  }
  if (!wide && small_build_plan(c) != DFLOW_OK) small_free_plan(c);  // optional fast path of dflow_train_epoch
  *out = c;
  return DFLOW_OK;
}

int dflow_chain_destroy(dflow_chain* c) {
  if (!c) return DFLOW_OK;
  if (c->d_chain) cudaFree(c->d_chain);
  if (c->d_staged) cudaFree(c->d_staged);
  if (c->d_gridmeta) cudaFree(c->d_gridmeta);
  tc_free_plan(c);
  small_free_plan(c);
  if (c->pipe) pipe_free((HostPipe*)c->pipe);
  delete c;
  return DFLOW_OK;
}

int64_t dflow_param_count(const dflow_chain* c) { return c ? c->hc()->h.P : DFLOW_E_INVALID_ARG; }

int dflow_chain_axes(const dflow_chain* c, int32_t elem, int32_t* axis_id, int32_t* n_id, int32_t* axis_nn,
                     int32_t* n_nn) {
  if (!c || elem < 0 || elem >= c->hc()->h.L) {
    set_error("bad element index");
    return DFLOW_E_INVALID_ARG;
  }
  const DevElem& E = c->hc()->e[elem];
  if (E.kind == DFLOW_ELEM_NORM) {
    set_error("element %d is not a coupling layer", elem);
    return DFLOW_E_INVALID_ARG;
  }
  const int n = c->hc()->h.n;
  if (n_id) *n_id = E.nid;
  if (n_nn) *n_nn = E.nin;
  if (axis_id)
    for (int k = 0; k < E.nid; ++k) axis_id[k] = E.id[k];
  if (axis_nn) {  // vcat(1:n, axis_id .+ n), src/Axes.jl:98 (0-based here)
    for (int k = 0; k < n; ++k) axis_nn[k] = k;
    for (int k = 0; k < E.nid; ++k) axis_nn[n + k] = E.id[k] + n;
  }
  return DFLOW_OK;
}

int dflow_param_offset(const dflow_chain* c, int32_t elem, int32_t net, int32_t dense, int64_t* w_off, int64_t* b_off) {
  if (!c || elem < 0 || elem >= c->hc()->h.L) {
    set_error("bad element index");
    return DFLOW_E_INVALID_ARG;
  }
  const DevElem& E = c->hc()->e[elem];
  if (E.kind == DFLOW_ELEM_NORM || (net != 0 && net != 1) || (net == 0 && E.kind != DFLOW_ELEM_RNVP)) {
    set_error("element %d has no net %d", elem, net);
    return DFLOW_E_INVALID_ARG;
  }
  const DevNet& N = net == 0 ? E.s : E.t;
  if (dense < 0 || dense >= N.depth) {
    set_error("bad dense index %d", dense);
    return DFLOW_E_INVALID_ARG;
  }
  if (w_off) *w_off = N.p_w[dense];
  if (b_off) *b_off = N.p_b[dense];
  return DFLOW_OK;
}

int dflow_chain_set_theta_range(dflow_chain* c, const float* tmin, const float* tmax) {
  if (!c || !tmin || !tmax) {
    set_error("null argument");
    return DFLOW_E_INVALID_ARG;
  }
  DevChainHdr& H = c->hc()->h;
  for (int k = 0; k < H.n; ++k) {
    H.theta_min[k] = tmin[k];
    H.theta_rng[k] = tmax[k] - tmin[k];
    H.theta_inv[k] = H.theta_rng[k] == 0.0f ? 0.0f : 1.0f / H.theta_rng[k];
  }
  H.has_theta_range = 1;
  if (c->d_chain) CKA(cudaMemcpy(c->d_chain, c->host_chain.data(), sizeof(DevChainHdr), cudaMemcpyHostToDevice));
  return DFLOW_OK;
}

static int check_common(dflow_chain* c, const float* W, const float* theta, const float* theta_const, int64_t B,
                        int32_t flags) {
  if (!c) {
    set_error("null chain");
    return DFLOW_E_INVALID_ARG;
  }
  const DevChainHdr& H = c->hc()->h;
  if (B < 0) {
    set_error("negative batch");
    return DFLOW_E_INVALID_ARG;
  }
  if (H.P > 0 && !W) {
    set_error("null parameter buffer");
    return DFLOW_E_INVALID_ARG;
  }
  if (H.n > 0 && B > 0 && !theta && !theta_const) {
    set_error("chain has %d conditions but no θ was given", H.n);
    return DFLOW_E_INVALID_ARG;
  }
  if (theta && theta_const) {
    set_error("pass either theta or theta_const, not both");
    return DFLOW_E_INVALID_ARG;
  }
  if ((flags & DFLOW_THETA_NORMALIZE) && !H.has_theta_range) {
    set_error("DFLOW_THETA_NORMALIZE needs θ_min/θ_max on the chain");
    return DFLOW_E_INVALID_ARG;
  }
  return DFLOW_OK;
}

static int run_fwd(dflow_chain* c, const float* W, FwdArgs& a, void* stream) {
  if (a.B == 0) return DFLOW_OK;
  if (c->use_tc_fwd(a.B)) return wide_fwd(c, W, a, stream);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = launch_prepack(c, W, st);
  if (rc) return rc;
  return launch_fwd(c, a, st);
}

int dflow_normalize(dflow_chain* c, const float* W, const float* x, const float* theta, int64_t B, int32_t flags,
                    float* z_out, float* ldj_out, void* stream) {
  int rc = check_common(c, W, theta, nullptr, B, flags);
  if (rc) return rc;
  if (B > 0 && (!x || !z_out || !ldj_out)) {
    set_error("null data pointer");
    return DFLOW_E_INVALID_ARG;
  }
  FwdArgs a{};
  a.x_in = x;
  a.theta = theta;
  a.x_out = z_out;
  a.aux_out = ldj_out;
  a.B = B;
  a.mode = MODE_NORMALIZE;
  a.flags = flags;
  return run_fwd(c, W, a, stream);
}

int dflow_logpdf(dflow_chain* c, const float* W, const float* x, const float* theta, int64_t B, const int32_t* idx,
                 int32_t flags, float* logp_out, void* stream) {
  int rc = check_common(c, W, theta, nullptr, B, flags);
  if (rc) return rc;
  if (B > 0 && (!x || !logp_out)) {
    set_error("null data pointer");
    return DFLOW_E_INVALID_ARG;
  }
  FwdArgs a{};
  a.x_in = x;
  a.theta = theta;
  a.idx = idx;
  a.aux_out = logp_out;
  a.B = B;
  a.mode = MODE_LOGPDF;
  a.flags = flags;
  return run_fwd(c, W, a, stream);
}

// logpdf(flow, x::NTuple{d, Vector}, θ::NTuple) on the tensor-product grid of the d coordinate vectors
// (src/Flows.jl:287-331): the reference materialises the (d, prod(lens)) array with Iterators.product and a (n, prod(lens))
// broadcast of θ; here the kernel derives every point's coordinates from its flat index and reads one constant θ.
int dflow_logpdf_grid(dflow_chain* c, const float* W, const float* grid_vals, const int64_t* lens, const float* theta_const,
                      int32_t flags, float* logp_out, void* stream) {
  if (!c || !lens) {
    set_error("null argument");
    return DFLOW_E_INVALID_ARG;
  }
  const DevChainHdr& H = c->hc()->h;
  long long B = 1, off = 0;
  for (int k = 0; k < H.d; ++k) {
    if (lens[k] < 0 || (lens[k] > 0 && B > (1LL << 40) / lens[k])) {
      set_error("bad grid length %lld for dimension %d", (long long)lens[k], k);
      return DFLOW_E_INVALID_ARG;
    }
    c->h_gridmeta[3 * k] = lens[k];
    c->h_gridmeta[3 * k + 1] = B;
    c->h_gridmeta[3 * k + 2] = off;
    B *= lens[k];
    off += lens[k];
  }
  int rc = check_common(c, W, nullptr, theta_const, B, flags);
  if (rc) return rc;
  if (B == 0) return DFLOW_OK;
  if (!grid_vals || !logp_out) {
    set_error("null data pointer");
    return DFLOW_E_INVALID_ARG;
  }
  CKA(cudaMemcpyAsync(c->d_gridmeta, c->h_gridmeta, sizeof(long long) * 3 * H.d, cudaMemcpyHostToDevice,
                      (cudaStream_t)stream));
  FwdArgs a{};
  a.theta_const = theta_const;
  a.aux_out = logp_out;
  a.B = B;
  a.mode = MODE_LOGPDF;
  a.flags = flags;
  a.grid_vals = grid_vals;
  a.grid_meta = c->d_gridmeta;
  return run_fwd(c, W, a, stream);
}

int dflow_logpdf_sum(dflow_chain* c, const float* W, const float* x, const float* theta, int64_t B, const int32_t* idx,
                     int32_t flags, float* loss_out, void* stream) {
  int rc = check_common(c, W, theta, nullptr, B, flags);
  if (rc) return rc;
  if (B > 0 && (!x || !loss_out)) {
    set_error("null data pointer");
    return DFLOW_E_INVALID_ARG;
  }
  FwdArgs a{};
  a.x_in = x;
  a.theta = theta;
  a.idx = idx;
  a.aux_out = loss_out;
  a.B = B;
  a.mode = MODE_LOGPDF_SUM;
  a.flags = flags;
  return run_fwd(c, W, a, stream);
}

int dflow_sample_inplace(dflow_chain* c, const float* W, float* z, const float* theta, const float* theta_const,
                         int64_t B, int32_t flags, void* stream) {
  int rc = check_common(c, W, theta, theta_const, B, flags);
  if (rc) return rc;
  if (B > 0 && !z) {
    set_error("null data pointer");
    return DFLOW_E_INVALID_ARG;
  }
  FwdArgs a{};
  a.x_in = z;
  a.theta = theta;
  a.theta_const = theta_const;
  a.x_out = z;
  a.B = B;
  a.mode = MODE_SAMPLE;
  a.flags = flags;
  return run_fwd(c, W, a, stream);
}

int dflow_forward_ldj(dflow_chain* c, const float* W, const float* z, const float* theta, int64_t B, int32_t flags,
                      float* x_out, float* ldj_out, void* stream) {
  int rc = check_common(c, W, theta, nullptr, B, flags);
  if (rc) return rc;
  if (B > 0 && (!z || !x_out || !ldj_out)) {
    set_error("null data pointer");
    return DFLOW_E_INVALID_ARG;
  }
  FwdArgs a{};
  a.x_in = z;
  a.theta = theta;
  a.x_out = x_out;
  a.aux_out = ldj_out;
  a.B = B;
  a.mode = MODE_FORWARD_LDJ;
  a.flags = flags;
  return run_fwd(c, W, a, stream);
}

int dflow_sample_rng(dflow_chain* c, const float* W, uint64_t seed, uint32_t offset, uint64_t first_sample,
                     const float* theta, const float* theta_const, int64_t B, int32_t flags, float* x_out,
                     void* stream) {
  int rc = check_common(c, W, theta, theta_const, B, flags);
  if (rc) return rc;
  if (B > 0 && !x_out) {
    set_error("null data pointer");
    return DFLOW_E_INVALID_ARG;
  }
  FwdArgs a{};
  a.theta = theta;
  a.theta_const = theta_const;
  a.x_out = x_out;
  a.B = B;
  a.mode = MODE_SAMPLE_RNG;
  a.flags = flags;
  a.seed = seed;
  a.rng_offset = offset;
  a.first_sample = first_sample;
  return run_fwd(c, W, a, stream);
}

// Scratch of the forward-type calls (normalize / logpdf / logpdf_sum / sample_* / forward_ldj) for a batch of B samples:
// 0 for chains that run on the CUDA-core kernels (whole chain in registers / shared memory), the tile-blocked working state
// for the tensor-core kernels.  The caller owns the buffer and attaches it to the handle; hot calls never allocate.
size_t dflow_scratch_bytes(const dflow_chain* c, int64_t B) {
  if (!c || B <= 0 || !c->tcp || !(c->use_tc_fwd(B))) return 0;
  return tc_scratch_bytes(c, B);
}

int dflow_chain_set_scratch(dflow_chain* c, void* scratch, size_t bytes) {
  if (!c || (bytes > 0 && !scratch)) {
    set_error("null argument");
    return DFLOW_E_INVALID_ARG;
  }
  c->scratch = scratch;
  c->scratch_bytes = scratch ? bytes : 0;
  return DFLOW_OK;
}

size_t dflow_workspace_bytes(const dflow_chain* c, int64_t B) {
  if (!c) return 0;
  if (c->use_tc_grad(B)) return tc_workspace_bytes(c, B);
  // one checkpoint slab per resident CTA (not per sample): the launcher caps the grid at sm_count * 4 CTAs, and the
  // largest tile of any instantiation is 768 sample slots (384 threads x 2, 192 threads x 4)
  return (size_t)c->sm_count * 4 * 768 * (size_t)c->hc()->h.ck_total * sizeof(float) + 256;
}

int dflow_loss_grad(dflow_chain* c, const float* W, const float* x, const float* theta, int64_t B, const int32_t* idx,
                    float inv_btot, int32_t flags, float* loss_out, float* grad_out, void* ws, size_t ws_bytes,
                    void* stream) {
  int rc = check_common(c, W, theta, nullptr, B, flags);
  if (rc) return rc;
  if (B > 0 && (!x || !loss_out || !grad_out)) {
    set_error("null data pointer");
    return DFLOW_E_INVALID_ARG;
  }
  if (B > 0 && (!ws || ws_bytes < dflow_workspace_bytes(c, B))) {
    set_error("workspace too small: need %zu bytes (dflow_workspace_bytes), got %zu", dflow_workspace_bytes(c, B),
              ws_bytes);
    return DFLOW_E_INVALID_ARG;
  }
  if (B == 0) return DFLOW_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (c->use_tc_grad(B)) return tc_loss_grad(c, W, x, theta, B, idx, inv_btot, flags, loss_out, grad_out, ws, ws_bytes, st);
  rc = launch_prepack(c, W, st);
  if (rc) return rc;
  GradArgs a{};
  a.x_in = x;
  a.theta = theta;
  a.idx = idx;
  a.loss_out = loss_out;
  a.grad_out = grad_out;
  a.ws = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  a.B = B;
  a.inv_btot = inv_btot;
  a.flags = flags;
  return launch_grad(c, a, st);
}

// Vector-Jacobian product of backward(chain, x, θ) -> (z, ln_det_jac) with caller cotangents (z̄, j̄): the pullback of
// ChainRulesCore.rrule(::typeof(backward), chain, x, θ), i.e. the rrule of src/affine/RNVP.jl:99-147 composed through the
// chain (src/Chains.jl:149-164) with the Dense pullbacks.
int dflow_vjp(dflow_chain* c, const float* W, const float* x, const float* theta, int64_t B, int32_t flags,
              const float* zbar, const float* jbar, float* grad_out, float* xbar_out, float* thetabar_out, void* ws,
              size_t ws_bytes, void* stream) {
  int rc = check_common(c, W, theta, nullptr, B, flags);
  if (rc) return rc;
  if (B > 0 && (!x || !grad_out || !zbar)) {
    set_error("null pointer (x, zbar and grad_out are required; jbar = NULL means a zero cotangent of ln_det_jac)");
    return DFLOW_E_INVALID_ARG;
  }
  if (B > 0 && (!ws || ws_bytes < dflow_workspace_bytes(c, B))) {
    set_error("workspace too small: need %zu bytes (dflow_workspace_bytes), got %zu", dflow_workspace_bytes(c, B),
              ws_bytes);
    return DFLOW_E_INVALID_ARG;
  }
  if (B == 0) return DFLOW_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (c->use_tc_grad(B)) {
    TcVjp v{};
    v.zbar = zbar;
    v.jbar = jbar;
    v.xbar_out = xbar_out;
    v.thbar_out = thetabar_out;
    // jbar == NULL: inv_btot = 0 is the (constant) minus-cotangent of ln_det_jac
    return tc_loss_grad(c, W, x, theta, B, nullptr, 0.0f, flags, nullptr, grad_out, ws, ws_bytes, st, &v);
  }
  rc = launch_prepack(c, W, st);
  if (rc) return rc;
  GradArgs a{};
  a.x_in = x;
  a.theta = theta;
  a.grad_out = grad_out;
  a.ws = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  a.B = B;
  a.inv_btot = 0.0f;
  a.flags = flags;
  a.zbar = zbar;
  a.jbar = jbar;
  a.xbar_out = xbar_out;
  a.thbar_out = thetabar_out;
  return launch_grad(c, a, st);
}

// One epoch of minibatch steps enqueued from C (src/Flows.jl:394-416): no host round trip between the steps.
int dflow_train_epoch(dflow_chain* c, float* W, float* m, float* v, const float* x, const float* theta,
                      const int32_t* order, int64_t n, int64_t batchsize, float lr, float beta1, float beta2, float eps,
                      int64_t* t_io, int32_t flags, float* grad_scratch, float* loss2_out, void* ws, size_t ws_bytes,
                      void* stream) {
  if (!c || !t_io || n < 0 || batchsize < 1 || *t_io < 0) {
    set_error("bad dflow_train_epoch arguments");
    return DFLOW_E_INVALID_ARG;
  }
  if (n == 0) return DFLOW_OK;
  const int64_t bmax = std::min<int64_t>(batchsize, n);
  int rc = check_common(c, W, theta, nullptr, bmax, flags);
  if (rc) return rc;
  const long long P = c->hc()->h.P;
  if (!x || !order || !m || !v || !grad_scratch) {
    set_error("null pointer");
    return DFLOW_E_INVALID_ARG;
  }
  if (!ws || ws_bytes < dflow_workspace_bytes(c, bmax)) {
    set_error("workspace too small: need %zu bytes (dflow_workspace_bytes at the batch size), got %zu",
              dflow_workspace_bytes(c, bmax), ws_bytes);
    return DFLOW_E_INVALID_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  // Minibatches of up to 512 samples (the reference default is 64): the whole epoch runs inside one persistent CTA with the
  // parameters, Adam moments and the minibatch's activations in shared memory (dflow_small.cu) -- one launch per epoch.
  // Larger minibatches fill the machine and go through the per-minibatch launches below.
  if (c->small && c->epoch_kernel >= 0 && batchsize <= 512 && !c->use_tc_grad(bmax)) {
    rc = small_train_epoch(c, W, m, v, x, theta, order, n, batchsize, lr, beta1, beta2, eps, *t_io, flags, loss2_out, st);
    if (rc) return rc;
    *t_io += (n + batchsize - 1) / batchsize;
    return DFLOW_OK;
  }
  float b1t, b2t;
  adam_beta_powers(beta1, beta2, *t_io, &b1t, &b2t);
  float* loss2 = grad_scratch + P;  // [grad (P) | sum logp, #non-finite] of the current minibatch
  for (int64_t b0 = 0; b0 < n; b0 += batchsize) {
    const int64_t nb = std::min<int64_t>(batchsize, n - b0);
    if (cudaMemsetAsync(grad_scratch, 0, sizeof(float) * (size_t)(P + 2), st) != cudaSuccess) {
      set_error("cudaMemsetAsync failed: %s", cudaGetErrorString(cudaGetLastError()));
      return DFLOW_E_CUDA;
    }
    rc = dflow_loss_grad(c, W, x, theta, nb, order + b0, (float)(1.0 / (double)nb), flags, loss2, grad_scratch, ws, ws_bytes,
                         stream);
    if (rc) return rc;
    if (loss2_out) {
      rc = launch_axpy2(loss2_out, loss2, st);
      if (rc) return rc;
    }
    b1t *= beta1;
    b2t *= beta2;
    ++*t_io;
    rc = launch_adam_pw(W, grad_scratch, m, v, P, lr, beta1, beta2, eps, b1t, b2t, st);
    if (rc) return rc;
  }
  return DFLOW_OK;
}

int dflow_adam_step(float* W, const float* g, float* m, float* v, int64_t P, float lr, float beta1, float beta2,
                    float eps, int64_t t, void* stream) {
  if (P < 0 || t < 1 || (P > 0 && (!W || !g || !m || !v))) {
    set_error("bad Adam arguments");
    return DFLOW_E_INVALID_ARG;
  }
  return launch_adam(W, g, m, v, P, lr, beta1, beta2, eps, t, (cudaStream_t)stream);
}

int dflow_minmax(const float* x, int32_t rows, int64_t B, float* min_out, float* max_out, void* stream) {
  if (rows < 1 || rows > 256 || B < 0 || !min_out || !max_out || (B > 0 && !x)) {
    set_error("bad minmax arguments");
    return DFLOW_E_INVALID_ARG;
  }
  return launch_minmax(x, rows, B, min_out, max_out, (cudaStream_t)stream);
}

// out[j] = base[perm(first + j)] (base == NULL: perm itself) for j in [0, count): perm = the seed's pseudo-random permutation of
// [0, n).  Stateless, so a data-parallel rank draws only its own slice of the epoch's order.
int dflow_shuffle_indices(uint64_t seed, int64_t n, int64_t first, int64_t count, const int32_t* base, int32_t* out,
                          void* stream) {
  if (n < 0 || first < 0 || count < 0 || first + count > n || n > 0x7FFFFFFFLL || (count > 0 && !out)) {
    set_error("bad dflow_shuffle_indices arguments (need 0 <= first, first + count <= n <= 2^31 - 1)");
    return DFLOW_E_INVALID_ARG;
  }
  if (count == 0) return DFLOW_OK;
  return launch_shuffle(seed, n, first, count, base, out, (cudaStream_t)stream);
}

// ---- host-buffer pipelines: chunked H2D -> kernel -> D2H on two streams ---------------------------------------
struct HostPipe {
  void* tc_scratch = nullptr;  // working state of the tensor-core kernels for one chunk (the _host entry points own their
  size_t tc_scratch_bytes = 0; // device staging: the caller only has host buffers)
  cudaStream_t st[2] = {nullptr, nullptr};
  float* dx[2] = {nullptr, nullptr};
  float* dt[2] = {nullptr, nullptr};
  float* dout[2] = {nullptr, nullptr};
  int64_t chunk = 0;
  int d = 0, n = 0;
};

static void pipe_free(HostPipe* p) {
  if (!p) return;
  if (p->tc_scratch) cudaFree(p->tc_scratch);
  for (int i = 0; i < 2; ++i) {
    if (p->dx[i]) cudaFree(p->dx[i]);
    if (p->dt[i]) cudaFree(p->dt[i]);
    if (p->dout[i]) cudaFree(p->dout[i]);
    if (p->st[i]) cudaStreamDestroy(p->st[i]);
  }
  delete p;
}

static int pipe_get(dflow_chain* c, int64_t chunk, HostPipe** out) {
  const DevChainHdr& H = c->hc()->h;
  HostPipe* p = (HostPipe*)c->pipe;
  if (p && p->chunk >= chunk) {
    *out = p;
    return DFLOW_OK;
  }
  pipe_free(p);
  c->pipe = nullptr;
  p = new (std::nothrow) HostPipe();
  if (!p) return DFLOW_E_NOMEM;
  p->chunk = chunk;
  p->d = H.d;
  p->n = H.n;
  const size_t outw = std::max(H.d, 1);
  for (int i = 0; i < 2; ++i) {
    if (cudaStreamCreateWithFlags(&p->st[i], cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&p->dx[i], sizeof(float) * H.d * chunk) != cudaSuccess ||
        cudaMalloc(&p->dt[i], sizeof(float) * std::max(H.n, 1) * chunk) != cudaSuccess ||
        cudaMalloc(&p->dout[i], sizeof(float) * outw * chunk) != cudaSuccess) {
      set_error("host pipeline allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
      pipe_free(p);
      return DFLOW_E_NOMEM;
    }
  }
  if (c->use_tc()) {
    p->tc_scratch_bytes = tc_scratch_bytes(c, chunk);
    if (cudaMalloc(&p->tc_scratch, p->tc_scratch_bytes) != cudaSuccess) {
      set_error("host pipeline allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
      pipe_free(p);
      return DFLOW_E_NOMEM;
    }
  }
  c->pipe = p;
  *out = p;
  return DFLOW_OK;
}

// the tensor-core kernels of a _host call work in the pipeline's own scratch, whatever the caller attached to the handle
struct ScratchSwap {
  dflow_chain* c;
  void* old;
  size_t old_bytes;
  ScratchSwap(dflow_chain* c_, HostPipe* p) : c(c_), old(c_->scratch), old_bytes(c_->scratch_bytes) {
    if (p->tc_scratch) {
      c->scratch = p->tc_scratch;
      c->scratch_bytes = p->tc_scratch_bytes;
    }
  }
  ~ScratchSwap() {
    c->scratch = old;
    c->scratch_bytes = old_bytes;
  }
};

int dflow_logpdf_host(dflow_chain* c, const float* W, const float* x_host, const float* theta_host, int64_t B,
                      int32_t flags, float* logp_host, int64_t chunk) {
  int rc = check_common(c, W, theta_host, nullptr, B, flags);
  if (rc) return rc;
  if (B == 0) return DFLOW_OK;
  if (!x_host || !logp_host) {
    set_error("null host pointer");
    return DFLOW_E_INVALID_ARG;
  }
  const DevChainHdr& H = c->hc()->h;
  if (chunk <= 0) chunk = 1 << 22;
  chunk = std::min<int64_t>(chunk, B);
  HostPipe* p;
  rc = pipe_get(c, chunk, &p);
  if (rc) return rc;
  ScratchSwap swap(c, p);
  // the staged image is shared by both streams: prepack once, make both streams wait for it
  cudaEvent_t ev;
  CKA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  rc = c->use_tc() ? DFLOW_OK : launch_prepack(c, W, p->st[0]);  // the wide path prepacks inside wide_fwd
  if (rc) return rc;
  CKA(cudaEventRecord(ev, p->st[0]));
  CKA(cudaStreamWaitEvent(p->st[1], ev, 0));
  int64_t done = 0;
  int slot = 0;
  while (done < B) {
    const int64_t nb = std::min(chunk, B - done);
    cudaStream_t st = p->st[slot];
    CKA(cudaMemcpyAsync(p->dx[slot], x_host + done * H.d, sizeof(float) * H.d * nb, cudaMemcpyHostToDevice, st));
    if (H.n > 0)
      CKA(cudaMemcpyAsync(p->dt[slot], theta_host + done * H.n, sizeof(float) * H.n * nb, cudaMemcpyHostToDevice, st));
    FwdArgs a{};
    a.x_in = p->dx[slot];
    a.theta = H.n > 0 ? p->dt[slot] : nullptr;
    a.aux_out = p->dout[slot];
    a.B = nb;
    a.mode = MODE_LOGPDF;
    a.flags = flags;
    rc = c->use_tc() ? wide_fwd(c, W, a, st) : launch_fwd(c, a, st);
    if (rc) return rc;
    CKA(cudaMemcpyAsync(logp_host + done, p->dout[slot], sizeof(float) * nb, cudaMemcpyDeviceToHost, st));
    done += nb;
    if (!c->use_tc()) slot ^= 1;  // the wide path owns one scratch buffer: keep it on a single stream
  }
  CKA(cudaStreamSynchronize(p->st[0]));
  CKA(cudaStreamSynchronize(p->st[1]));
  cudaEventDestroy(ev);
  return DFLOW_OK;
}

int dflow_sample_host(dflow_chain* c, const float* W, uint64_t seed, const float* theta_const_host, int64_t B,
                      int32_t flags, float* x_host, int64_t chunk) {
  if (!c) {
    set_error("null chain");
    return DFLOW_E_INVALID_ARG;
  }
  const DevChainHdr& H = c->hc()->h;
  if (B < 0 || (B > 0 && !x_host) || (H.n > 0 && !theta_const_host) || (H.P > 0 && !W)) {
    set_error("bad sample_host arguments");
    return DFLOW_E_INVALID_ARG;
  }
  if ((flags & DFLOW_THETA_NORMALIZE) && !H.has_theta_range) {
    set_error("DFLOW_THETA_NORMALIZE needs θ_min/θ_max on the chain");
    return DFLOW_E_INVALID_ARG;
  }
  if (B == 0) return DFLOW_OK;
  if (chunk <= 0) chunk = 1 << 22;
  chunk = std::min<int64_t>(chunk, B);
  HostPipe* p;
  int rc = pipe_get(c, chunk, &p);
  if (rc) return rc;
  ScratchSwap swap(c, p);
  if (H.n > 0) {
    CKA(cudaMemcpyAsync(p->dt[0], theta_const_host, sizeof(float) * H.n, cudaMemcpyHostToDevice, p->st[0]));
  }
  cudaEvent_t ev;
  CKA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  rc = c->use_tc() ? DFLOW_OK : launch_prepack(c, W, p->st[0]);  // the wide path prepacks inside wide_fwd
  if (rc) return rc;
  CKA(cudaEventRecord(ev, p->st[0]));
  CKA(cudaStreamWaitEvent(p->st[1], ev, 0));
  int64_t done = 0;
  int slot = 0;
  while (done < B) {
    const int64_t nb = std::min(chunk, B - done);
    cudaStream_t st = p->st[slot];
    FwdArgs a{};
    a.theta_const = H.n > 0 ? p->dt[0] : nullptr;
    a.x_out = p->dout[slot];
    a.B = nb;
    a.mode = MODE_SAMPLE_RNG;
    a.flags = flags;
    a.seed = seed;
    a.first_sample = (unsigned long long)done;
    rc = c->use_tc() ? wide_fwd(c, W, a, st) : launch_fwd(c, a, st);
    if (rc) return rc;
    CKA(cudaMemcpyAsync(x_host + done * H.d, p->dout[slot], sizeof(float) * H.d * nb, cudaMemcpyDeviceToHost, st));
    done += nb;
    if (!c->use_tc()) slot ^= 1;  // the wide path owns one scratch buffer: keep it on a single stream
  }
  CKA(cudaStreamSynchronize(p->st[0]));
  CKA(cudaStreamSynchronize(p->st[1]));
  cudaEventDestroy(ev);
  return DFLOW_OK;
}

int dflow_set_tuning(dflow_chain* c, const char* key, int32_t value) {
  if (!c || !key) {
    set_error("null argument");
    return DFLOW_E_INVALID_ARG;
  }
  if (!strcmp(key, "fwd_spt"))
    c->fwd_spt = value;
  else if (!strcmp(key, "grad_smem"))
    c->grad_smem = value;
  else if (!strcmp(key, "fwd_const"))
    c->fwd_const = value;
  else if (!strcmp(key, "fwd_threads"))
    c->fwd_threads = value;
  else if (!strcmp(key, "grad_threads"))
    c->grad_threads = value;
  else if (!strcmp(key, "grad_spt"))
    c->grad_spt = value;
  else if (!strcmp(key, "ctas_per_sm"))
    c->ctas_per_sm = value;
  else if (!strcmp(key, "epoch_kernel"))
    c->epoch_kernel = value;
  else if (!strcmp(key, "tc_fuse"))
    c->tc_fuse = value < 0 ? 0 : value > 2 ? 2 : value;
  else if (!strcmp(key, "tc_ts"))
    c->tc_ts = value < 0 ? -1 : 0;
  else if (!strcmp(key, "tc_dw_groups"))
    c->tc_dw_groups = value < 0 ? 0 : value;
  else if (!strcmp(key, "tc_dw_ts"))
    c->tc_dw_ts = value < 0 ? -1 : 0;
  else if (!strcmp(key, "tc_ws_budget_mb"))
    c->tc_ws_budget_mb = value;
#ifdef DFLOW_TC_EXPERIMENTS  // timing experiments (wrong results by construction): never part of a release build
  else if (!strcmp(key, "tc_debug"))
    c->tc_debug = value;
  else if (!strcmp(key, "tc_cluster"))
    c->tc_debug_cluster = value < 0 ? 0 : value > 2 ? 2 : value;
#endif
  else if (!strcmp(key, "tc_mode")) {
    if (value > 0 && !c->tcp) {
      set_error("tc_mode: this chain is not eligible for the tensor-core path (needs Dense(in,h,relu)->Dense(h,h,relu)->"
                "Dense(h,a) conditioners with bias, h a multiple of 32)");
      return DFLOW_E_UNSUPPORTED;
    }
    if (value < 0 && c->must_wide) {
      set_error("tc_mode: hidden widths above 64 only run on the tensor-core path");
      return DFLOW_E_UNSUPPORTED;
    }
    c->tc_mode = value;
  }
  else {
    set_error("unknown tuning key %s", key);
    return DFLOW_E_INVALID_ARG;
  }
  return DFLOW_OK;
}

int64_t dflow_launch_count(const dflow_chain* c) { return c ? c->launches : 0; }

}  // extern "C"
