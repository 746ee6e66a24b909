/* dflow.h -- C ABI of libdflow.so: B200 (sm_100a) kernels for the DensityFlows.jl coupling-chain hot path.
 *
 * The reference (pure Julia, /root/reference) has no FFI: its extension point is the FlowElement method
 * protocol `backward / forward / forward!(elem, array, θ) -> (array, ln_det_jac)` (src/Chains.jl:33-72,
 * docs/src/documentation.md:170-197).  This header is the boundary a Julia `ccall` layer binds
 * (julia/DensityFlowsB200.jl); every entry point cites the reference function it replaces.
 *
 * Conventions
 *  - All arrays are Float32, Julia column-major `(rows, B)` => sample-contiguous: element (k,b) at ptr[k + rows*b].
 *  - Index vectors are int32, 0-based (Julia side subtracts 1).
 *  - Data / weight / gradient / workspace buffers are DEVICE pointers owned by the caller unless the name ends
 *    in `_host`.  The library owns only the opaque handles.  Compute calls are asynchronous on `stream`
 *    (a cudaStream_t passed as void*); one stream at a time per chain handle.
 *  - Every function returns 0 on success or a negative DFLOW_E_* code; the message is in dflow_last_error()
 *    (thread-local).  Nothing throws or aborts across the ABI.  There is no CPU fallback.
 *  - Packed parameter buffer (length dflow_param_count): chain order; per coupling layer s_net then t_net
 *    (Flux.@layer trainable=(s_net,t_net), src/affine/RNVP.jl:51); per Dense `vec(weight)` (Flux (out,in)
 *    column-major: W[o + out*i]) then `bias`.  Gradients use the identical layout.
 */
#ifndef DFLOW_H
#define DFLOW_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFLOW_VERSION 100 /* 0.1.0 */

enum {
  DFLOW_OK = 0,
  DFLOW_E_INVALID_ARG = -1,
  DFLOW_E_UNSUPPORTED = -2,
  DFLOW_E_CUDA = -3,
  DFLOW_E_NCCL = -4,
  DFLOW_E_NOMEM = -5
};

/* element kinds: RNVPCouplingLayer (src/affine/RNVP.jl:41), NICECouplingLayer (src/affine/NICE.jl),
 * NormalizationLayer (src/norm/Normalization.jl:30).  CouplingBlock / nested FlowChain are flattened by the host
 * into chain order (block -> layer_1, layer_2; src/Blocks.jl:127-161). */
enum { DFLOW_ELEM_RNVP = 0, DFLOW_ELEM_NICE = 1, DFLOW_ELEM_NORM = 2 };

/* Dense activations (Flux.relu default, src/Layers.jl:37) */
enum { DFLOW_ACT_IDENTITY = 0, DFLOW_ACT_RELU = 1, DFLOW_ACT_TANH = 2, DFLOW_ACT_SIGMOID = 3 };

/* per-call flags */
enum {
  DFLOW_THETA_NORMALIZE = 1 /* apply normalize_input(θ, θ_min, θ_max) (src/Data.jl:213-218, src/Macros.jl:104-112) */
};

/* A conditioner MLP = Flux.Chain of Dense (src/Layers.jl:33-50). */
typedef struct dflow_net_desc {
  int32_t depth;         /* number of Dense layers (n_sublayers + 1); 0 = net absent */
  const int32_t* widths; /* depth+1 entries: in, ..., out */
  const int32_t* acts;   /* depth entries, DFLOW_ACT_* */
  int32_t has_bias;
} dflow_net_desc;

typedef struct dflow_elem_desc {
  int32_t kind;           /* DFLOW_ELEM_* */
  int32_t n_af;           /* |axis_af| (coupling layers) */
  const int32_t* axis_af; /* 0-based transformed dims in CALLER ORDER (src/Axes.jl:91) */
  int32_t n_id;           /* |axis_id|; 0 with axis_id == NULL => derive the ascending complement (src/Axes.jl:88) */
  const int32_t* axis_id; /* 0-based identity dims IN ORDER: they are the conditioner's x inputs (axis_nn = [0..n-1,
                             axis_id + n], src/Axes.jl:98).  reverse(axes) (src/Axes.jl:129-135) makes this the
                             un-sorted former axis_af, so it cannot always be derived. */
  dflow_net_desc s_net;   /* RNVP only */
  dflow_net_desc t_net;   /* RNVP and NICE */
  const float* x_min;     /* NORM only: host pointers, d entries (src/norm/Normalization.jl:51-57) */
  const float* x_max;
  float alpha, beta;
} dflow_elem_desc;

typedef struct dflow_chain_desc {
  int32_t d;       /* data dimensions */
  int32_t n;       /* conditions */
  int32_t n_elems; /* leaf elements in chain order */
  const dflow_elem_desc* elems;
  const float* theta_min; /* host, n entries, or NULL (MetaData, src/Data.jl:75-86) */
  const float* theta_max;
} dflow_chain_desc;

typedef struct dflow_chain dflow_chain; /* opaque */

int dflow_version(void);
const char* dflow_last_error(void);

/* Build a chain handle on the CURRENT CUDA device (replaces constructing FlowChain / CouplingLayer / CouplingAxes:
 * src/Chains.jl:99-101, src/Layers.jl:113-136, src/Axes.jl:79-102). */
int dflow_chain_create(const dflow_chain_desc* desc, dflow_chain** out);
int dflow_chain_destroy(dflow_chain* chain);
/* number of trainable Float32 parameters (= sum(length, Flux.trainables(chain))) */
int64_t dflow_param_count(const dflow_chain* chain);
/* Derived integer index vectors of coupling element `elem` (0-based): axis_id (ascending complement,
 * src/Axes.jl:88) and axis_nn (rows of vcat(θ,x), src/Axes.jl:98).  Buffers must hold d and n+d ints. */
int dflow_chain_axes(const dflow_chain* chain, int32_t elem, int32_t* axis_id, int32_t* n_id, int32_t* axis_nn,
                     int32_t* n_nn);
/* offset of Dense `dense` of net `net` (0 = s_net, 1 = t_net) of element `elem` inside the packed buffer;
 * *b_off = -1 if the Dense has no bias. */
int dflow_param_offset(const dflow_chain* chain, int32_t elem, int32_t net, int32_t dense, int64_t* w_off,
                       int64_t* b_off);
/* replace θ_min / θ_max (host pointers, n entries each) */
int dflow_chain_set_theta_range(dflow_chain* chain, const float* theta_min, const float* theta_max);

/* ---- scratch of the forward-type calls below ---------------------------------------------------------------------
 * Chains on the CUDA-core kernels need none (dflow_scratch_bytes returns 0: the whole chain runs in registers / shared
 * memory).  Chains routed to the tensor-core kernels (hidden > 64 always; hidden 64 from B >= 131072) work on a
 * tile-blocked copy of the state: the caller allocates dflow_scratch_bytes(chain, B_max) bytes once, attaches them with
 * dflow_chain_set_scratch and keeps them alive; a forward-type call whose batch needs more returns DFLOW_E_INVALID_ARG.
 * No entry point that takes device pointers allocates device memory (the *_host entry points own their staging buffers). */
size_t dflow_scratch_bytes(const dflow_chain* chain, int64_t B);
int dflow_chain_set_scratch(dflow_chain* chain, void* scratch, size_t bytes);

/* ---- normalising direction: backward(chain, x, θ) -> (z, ln_det_jac)  (src/Chains.jl:149-164) -------------- */
int dflow_normalize(dflow_chain* chain, const float* W, const float* x, const float* theta, int64_t B, int32_t flags,
                    float* z_out, float* ldj_out, void* stream);
/* logpdf(flow, x, θ) = logpdf(MvNormal(0,I), z) + ln_det_jac (src/Flows.jl:272-281).  idx (device int32[B], may be
 * NULL) gathers sample b from column idx[b] of x and θ (selectdim views of src/Data.jl:185-187 / DataLoader batches). */
int dflow_logpdf(dflow_chain* chain, const float* W, const float* x, const float* theta, int64_t B, const int32_t* idx,
                 int32_t flags, float* logp_out, void* stream);
/* logpdf(flow, x::NTuple{d,Vector}, θ::NTuple{n}) on the tensor-product grid (src/Flows.jl:287-331).  grid_vals (device):
 * the d coordinate vectors back to back; lens (host, d entries): their lengths.  logp_out (device, prod(lens) floats) is in
 * Julia's column-major order of the (lens...) array: the first vector varies fastest (Iterators.product).  theta_const:
 * device, n floats (NULL when n = 0).  The (d, prod(lens)) point array is never materialised. */
int dflow_logpdf_grid(dflow_chain* chain, const float* W, const float* grid_vals, const int64_t* lens,
                      const float* theta_const, int32_t flags, float* logp_out, void* stream);
/* Σ_b logpdf_b accumulated into loss_out[0] (device float[2], caller zeroes it); loss_out[1] counts non-finite
 * samples.  loss = -loss_out[0]/B (src/Flows.jl:352-359; epoch-end passes src/Flows.jl:419-430). */
int dflow_logpdf_sum(dflow_chain* chain, const float* W, const float* x, const float* theta, int64_t B,
                     const int32_t* idx, int32_t flags, float* loss_out, void* stream);

/* ---- sampling direction -------------------------------------------------------------------------------- */
/* forward!(chain, z, θ): in place, no ln_det_jac (src/Chains.jl:187-197).  Exactly one of theta (n x B) /
 * theta_const (device, n floats, same θ for every sample: the NTuple method src/Flows.jl:174-185) may be non-NULL
 * when n > 0. */
int dflow_sample_inplace(dflow_chain* chain, const float* W, float* z_inout, const float* theta,
                         const float* theta_const, int64_t B, int32_t flags, void* stream);
/* forward(chain, z, θ) -> (x, ln_det_jac) (src/Chains.jl:167-183) */
int dflow_forward_ldj(dflow_chain* chain, const float* W, const float* z, const float* theta, int64_t B, int32_t flags,
                      float* x_out, float* ldj_out, void* stream);
/* sample(flow, B, θ) with the base draw done in-kernel (Philox4x32-10 + Box-Muller, spec in oracle/philox.py);
 * replaces rand(rng, base, B) + forward! (src/Flows.jl:157-192).  Sample b uses counter first_sample + b. */
int dflow_sample_rng(dflow_chain* chain, const float* W, uint64_t seed, uint32_t offset, uint64_t first_sample,
                     const float* theta, const float* theta_const, int64_t B, int32_t flags, float* x_out,
                     void* stream);

/* ---- adjoint: gradient of loss = -inv_btot * Σ_b logpdf_b w.r.t. the packed parameters ------------------- */
/* Replaces Flux.gradient through backward(::FlowChain) + loss (src/Flows.jl:400-413) incl. the rrule of
 * src/affine/RNVP.jl:99-147 and the Dense pullbacks.  grad_out (P floats) is ACCUMULATED INTO (caller zeroes it);
 * loss_out[0] += Σ logpdf_b, loss_out[1] += #non-finite.  inv_btot = 1/B for a single device, 1/B_global for a
 * data-parallel shard (the all-reduce is then a pure sum).  ws / ws_bytes: workspace from dflow_workspace_bytes. */
/* The workspace holds the per-CTA checkpoints of each layer's transformed coordinates (Σ_l a_l floats per resident
 * sample slot, independent of B), so that the reverse sweep recomputes activations from bit-identical inputs. */
size_t dflow_workspace_bytes(const dflow_chain* chain, int64_t B);
int dflow_loss_grad(dflow_chain* chain, const float* W, const float* x, const float* theta, int64_t B,
                    const int32_t* idx, float inv_btot, int32_t flags, float* loss_out, float* grad_out, void* ws,
                    size_t ws_bytes, void* stream);

/* ---- pullback with caller cotangents: the ChainRulesCore.rrule of backward(chain, x, θ) -> (z, ln_det_jac) --------
 * Replaces rrule(RNVP_backward) (src/affine/RNVP.jl:99-147) composed through the chain (src/Chains.jl:149-164) with the
 * Dense pullbacks Zygote would add: given z̄ = zbar (d x B) and j̄ = jbar (B; NULL = zeros) it ACCUMULATES the parameter
 * cotangent into grad_out (P floats, packed layout = a Tangent of the Flux structs) and writes the input cotangents
 * x̄ (d x B, xbar_out, may be NULL) and θ̄ (n x B, thetabar_out, may be NULL; with DFLOW_THETA_NORMALIZE it is the
 * cotangent of the RAW θ).  The forward values are recomputed inside (no tape): call dflow_normalize for (z, ldj).
 * Any loss that is a function of (z, ln_det_jac) can therefore train on the GPU path; dflow_loss_grad is the special case
 * z̄ = z / B_tot, j̄ = -1 / B_tot with the loss reduction fused in. */
int dflow_vjp(dflow_chain* chain, const float* W, const float* x, const float* theta, int64_t B, int32_t flags,
              const float* zbar, const float* jbar, float* grad_out, float* xbar_out, float* thetabar_out, void* ws,
              size_t ws_bytes, void* stream);

/* ---- Optimisers.Adam update (call site src/Flows.jl:415), in place on packed buffers; t = step count >= 1 ---- */
int dflow_adam_step(float* W, const float* g, float* m, float* v, int64_t P, float lr, float beta1, float beta2,
                    float eps, int64_t t, void* stream);

/* ---- one epoch of minibatch steps, enqueued back to back from C (src/Flows.jl:394-416: `for (x, θ) in loader` with
 * Flux.gradient + Optimisers.update! per minibatch).  order: n 0-based sample indices (the epoch's shuffled training
 * partition); minibatch k is order[k*batchsize ..], the last one may be partial (Flux.DataLoader partial=true); each step is
 * dflow_loss_grad with inv_btot = 1/|minibatch| followed by dflow_adam_step with t = ++*t_io.  grad_scratch: P + 2 floats.
 * loss2_out (optional, 2 floats) accumulates [Σ logp, #non-finite] over the minibatches.  ws: dflow_workspace_bytes at
 * min(batchsize, n).  Single device; data-parallel training shards every minibatch instead (dflow_loss_grad + dflow_dp_*). */
int dflow_train_epoch(dflow_chain* chain, float* W, float* m, float* v, const float* x, const float* theta,
                      const int32_t* order, int64_t n, int64_t batchsize, float lr, float beta1, float beta2, float eps,
                      int64_t* t_io, int32_t flags, float* grad_scratch, float* loss2_out, void* ws, size_t ws_bytes,
                      void* stream);

/* ---- data-parallel train step: gradient all-reduce over NVLink peer memory fused with the Adam update ------------
 * (one process per GPU; replaces NCCL all-reduce + dflow_adam_step for src/Flows.jl:413-415).  Protocol per rank:
 *   dflow_dp_create(rank, nranks, P, &dp, handle64)     allocates this rank's communication buffer, returns its
 *                                                        64-byte CUDA IPC handle
 *   (exchange the handles, e.g. all_gather)  ->  dflow_dp_connect(dp, handles)    nranks x 64 bytes in rank order
 *   every step: buf = dflow_dp_grad_buffer(dp); zero P+2 floats; dflow_loss_grad(..., loss_out = buf + P,
 *               grad_out = buf, ...) with inv_btot = 1/B_global; dflow_dp_allreduce_adam(dp, W, m, v, ...).
 * The kernel waits until every peer has published its gradient of the same step, sums the ranks' buffers in rank order
 * over peer loads (bit-identical replicas, no broadcast) and applies Adam; loss2_out (device float[2], may be NULL)
 * receives the reduced [sum logp, #non-finite].  All ranks must call it the same number of times. */
typedef struct dflow_dp dflow_dp; /* opaque */
int dflow_dp_create(int32_t rank, int32_t nranks, int64_t P, dflow_dp** out, void* ipc_handle_out);
int dflow_dp_connect(dflow_dp* dp, const void* handles);
float* dflow_dp_grad_buffer(dflow_dp* dp);
int dflow_dp_allreduce_adam(dflow_dp* dp, float* W, float* m, float* v, float lr, float beta1, float beta2, float eps,
                            int64_t t, float* loss2_out, void* stream);
/* 0 = fine; 1 = a peer did not publish its gradient within the time-out: the reduction AND the Adam update of that step
 * were skipped on this rank (stale peer buffers are never summed), the replicas are no longer in step and the caller
 * must stop.  Synchronises `stream`. */
int dflow_dp_status(dflow_dp* dp, void* stream);
/* measurement: SM clocks this rank's kernels have spent waiting in the peer barrier, and the number of steps, since creation */
int dflow_dp_wait_stats(dflow_dp* dp, void* stream, int64_t* wait_clk, int64_t* steps);
int dflow_dp_set_timeout_ms(dflow_dp* dp, int64_t ms); /* barrier time-out of the fused kernel, default ~10 s */
int dflow_dp_destroy(dflow_dp* dp);

/* ---- the same, driven by ONE host process (the reference's train! is a single Julia process, src/Flows.jl:380-445) ---
 * dflow_dp_create_local builds one context per entry of `devs` (out: ndev handles; out[r] lives on devs[r] and is rank r)
 * and connects them with cudaDeviceEnablePeerAccess -- no IPC handles.  The per-rank calls above work on these handles
 * unchanged (with devs[r] current).  dflow_dp_train_step is the fan-out a train! loop calls once per minibatch: for every
 * rank it zeroes the step's accumulation buffer, runs dflow_loss_grad on the shard with the global seed
 * inv_btot = 1 / B_global and launches the fused all-reduce + Adam kernel, all asynchronously on shard.stream (NULL: the
 * context's own non-blocking stream); the devices' kernels meet in the peer barrier.  Every rank holds its own replica
 * (chain handle created on its device, W / m / v, resident data shard); replicas stay bit-identical.
 * dflow_dp_sync waits for every rank's stream and returns the worst dflow_dp_status. */
typedef struct dflow_dp_shard {
  dflow_chain* chain;      /* handle created on this rank's device */
  float* W;                /* this replica's packed parameters and Adam moments (device) */
  float* m;
  float* v;
  const float* x;          /* resident data of this rank: (d, *), (n, *) */
  const float* theta;
  int64_t B;               /* samples of this rank's share of the minibatch (0 allowed) */
  const int32_t* idx;      /* B column indices into x / theta, or NULL for the first B columns */
  void* ws;                /* dflow_workspace_bytes(chain, B) */
  size_t ws_bytes;
  float* loss2_out;        /* device float[2] or NULL: reduced [sum logp, #non-finite] of the minibatch */
  void* stream;            /* cudaStream_t on this rank's device, or NULL */
} dflow_dp_shard;
int dflow_dp_create_local(int32_t ndev, const int32_t* devs, int64_t P, dflow_dp** out);
int dflow_dp_train_step(dflow_dp* const* dps, int32_t ndev, const dflow_dp_shard* shards, float inv_btot, int32_t flags,
                        float lr, float beta1, float beta2, float eps, int64_t t);
int dflow_dp_sync(dflow_dp* const* dps, int32_t ndev, const dflow_dp_shard* shards);

/* ---- per-row min / max over B samples (NormalizationLayer ctor src/norm/Normalization.jl:52-53; minimum_θ /
 * maximum_θ src/Data.jl:182-183).  min_out / max_out: device float[rows]. */
int dflow_minmax(const float* x, int32_t rows, int64_t B, float* min_out, float* max_out, void* stream);

/* ---- pseudo-random permutations on the device (DataPartition's randperm src/Data.jl:112-128; the per-epoch shuffle of
 * Flux.DataLoader, src/Flows.jl:394).  perm = cycle-walked 6-round Feistel bijection of [0, n) keyed by `seed` (specification:
 * oracle/shuffle.py, bit-exact test).  out[j] = base[perm(first + j)], or perm(first + j) when base == NULL, for j < count:
 * stateless, so every data-parallel rank evaluates exactly its own slice of the epoch's order. */
int dflow_shuffle_indices(uint64_t seed, int64_t n, int64_t first, int64_t count, const int32_t* base, int32_t* out,
                          void* stream);

/* ---- host-buffer entry points (pinned or pageable host memory; chunked H2D -> kernel -> D2H pipeline) -------- */
int dflow_logpdf_host(dflow_chain* chain, const float* W, const float* x_host, const float* theta_host, int64_t B,
                      int32_t flags, float* logp_host, int64_t chunk);
int dflow_sample_host(dflow_chain* chain, const float* W, uint64_t seed, const float* theta_const_host, int64_t B,
                      int32_t flags, float* x_host, int64_t chunk);

/* ---- tuning / introspection ------------------------------------------------------------------------------ */
/* keys: "fwd_spt", "grad_spt" (samples per thread), "fwd_threads", "grad_threads", "ctas_per_sm"; value 0 = automatic.
 * "fwd_const" (-1: keep small relu chains off the constant-bank forward kernel), "grad_smem" (-1: weight gradients of the
 * narrow adjoint go straight to global memory), "tc_mode" (1 / 0 / -1: force eligible hidden <= 64 chains onto / automatic
 * / off the tensor-core kernels), "tc_fuse" (0 / 1 / 2: s and t conditioners of a layer as one block-diagonal conditioner
 * never / in the train step / everywhere), "tc_ts" (-1: hidden 32 / 64 conditioners stay on the warp-specialised
 * tensor-core pipeline instead of the TMEM-sourced four-chain kernel), "tc_dw_groups" (1: the staging warps of the
 * weight-gradient kernel never split into alternating groups), "tc_dw_ts" (-1: weight gradients never take their A operands
 * through TMEM), "tc_ws_budget_mb" (adjoint workspace cap), "epoch_kernel" (-1: dflow_train_epoch
 * never uses the persistent small-minibatch kernel).  Unknown keys return DFLOW_E_INVALID_ARG. */
int dflow_set_tuning(dflow_chain* chain, const char* key, int32_t value);
/* number of kernels the library has launched on behalf of this handle since creation */
int64_t dflow_launch_count(const dflow_chain* chain);

#ifdef __cplusplus
}
#endif
#endif /* DFLOW_H */
