// Data-parallel train step, collective part: gradient all-reduce over NVLink peer memory fused with the Adam update
// (src/Flows.jl:413-415: gradient -> Optimisers.update!).  One process per GPU; every rank owns one communication
// buffer (cudaMalloc + CUDA IPC handle, opened by its peers):
//
//   [ flags: uint32[64] | go | bad | pad | wait statistics: uint64[2] at word 68 | pad to 512 B ] [ half 0: P + 2 floats, padded ] [ half 1: P + 2 floats, padded ]
//
// Step e writes its local gradient sum [grad | sum logp | #non-finite] into half e & 1 (the adjoint kernels accumulate
// there directly), then ONE kernel per rank: (1) thread 0 of CTA 0 publishes flags[rank] = e in every peer's buffer
// (st.release.sys after a system fence) and waits until all peers have published e in its own buffer
// (ld.acquire.sys); (2) every thread sums element i over the ranks' halves in rank order 0..N-1 -- peer loads over
// NVLink, identical order on every rank => bit-identical replicas without a broadcast -- and applies Adam to its own
// replica.  No separate all-reduce launch, no reduced-gradient round trip through HBM.
// Halves ping-pong: a rank can only start step e+2 (which overwrites half e & 1) after the barrier of step e+1, which
// every peer reaches only after it has finished reading step e.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "dflow_internal.h"

namespace dflow {

constexpr int DP_MAX_RANKS = 16;
constexpr size_t DP_HDR_BYTES = 512;

struct DpArgs {
  const char* bufs[DP_MAX_RANKS];  // every rank's communication buffer (own buffer at index rank)
  int rank, nranks;
  uint32_t epoch;
  long long P;
  size_t half_floats;  // padded size of one half
  float* W;
  float* m;
  float* v;
  float lr, b1, b2, eps, c1, c2;
  float* loss2_out;   // [2] reduced sum logp / #non-finite, or null
  int* status;        // device int: set to 1 if the barrier timed out
  long long timeout_clk;  // barrier time-out in SM clocks
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer(const float* p) {  // bypass L1: peer data is not kept coherent there
  float v;
  asm volatile("ld.global.cv.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(256) dp_allreduce_adam_kernel(const DpArgs a) {
  uint32_t* my_flags = reinterpret_cast<uint32_t*>(const_cast<char*>(a.bufs[a.rank]));
  uint32_t* go = my_flags + 64;    // epoch released by CTA 0 once every peer has published it
  uint32_t* bad = my_flags + 65;   // epoch whose barrier timed out on this rank (poison: nobody reduces or updates)
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    __threadfence_system();  // the adjoint kernels' gradient (earlier launches on this stream) before the flag
    for (int j = 0; j < a.nranks; ++j)
      st_release_sys(reinterpret_cast<uint32_t*>(const_cast<char*>(a.bufs[j])) + a.rank, a.epoch);
    const long long t0 = clock64();
    bool ok = true;
    for (int j = 0; j < a.nranks && ok; ++j) {
      while ((int32_t)(ld_acquire_sys(my_flags + j) - a.epoch) < 0) {
        if (clock64() - t0 > a.timeout_clk) {  // ~10 s by default: a peer is gone
          ok = false;
          break;
        }
      }
    }
    // accumulated barrier wait (SM clocks) and step count: the measurement behind the scaling split in bench.py
    {
      unsigned long long* stats = reinterpret_cast<unsigned long long*>(my_flags + 68);
      stats[0] += (unsigned long long)(clock64() - t0);
      stats[1] += 1ull;
    }
    if (!ok) {
      // the peers' halves may be stale or half written: skip the reduction AND the update on this rank, and say so
      if (a.status) *a.status = 1;
      *bad = a.epoch;
    }
    __threadfence();
    atomicExch(go, a.epoch);
  }
  // all CTAs are co-resident (cooperative launch, grid <= SM count): wait for the release by CTA 0
  __shared__ uint32_t poisoned;
  if (threadIdx.x == 0) {
    while ((int32_t)(ld_acquire_gpu(go) - a.epoch) < 0) {
    }
    poisoned = (ld_acquire_gpu(bad) == a.epoch) ? 1u : 0u;
  }
  __syncthreads();
  if (poisoned) return;
  const size_t half_off = DP_HDR_BYTES / sizeof(float) + (size_t)(a.epoch & 1u) * a.half_floats;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.P + 2; i += (long long)gridDim.x * blockDim.x) {
    float g = 0.0f;
    for (int j = 0; j < a.nranks; ++j) {
      const float* h = reinterpret_cast<const float*>(a.bufs[j]) + half_off;
      g += (j == a.rank) ? h[i] : ld_peer(h + i);
    }
    if (i >= a.P) {
      if (a.loss2_out) a.loss2_out[i - a.P] = g;
      continue;
    }
    // Optimisers.Adam, same explicit round-to-nearest sequence as adam_kernel (dflow_kernels.cu)
    const float mi = __fadd_rn(__fmul_rn(a.b1, a.m[i]), __fmul_rn(1.0f - a.b1, g));
    const float vi = __fadd_rn(__fmul_rn(a.b2, a.v[i]), __fmul_rn(1.0f - a.b2, __fmul_rn(g, g)));
    a.m[i] = mi;
    a.v[i] = vi;
    const float den = __fadd_rn(__fsqrt_rn(__fdiv_rn(vi, a.c2)), a.eps);
    const float stepv = __fmul_rn(__fdiv_rn(__fdiv_rn(mi, a.c1), den), a.lr);
    a.W[i] = __fsub_rn(a.W[i], stepv);
  }
}

}  // namespace dflow

using namespace dflow;

struct dflow_dp {
  int rank = 0, nranks = 1;
  int device = 0;
  long long P = 0;
  size_t half_floats = 0, bytes = 0;
  char* own = nullptr;
  char* peers[DP_MAX_RANKS] = {nullptr};
  bool opened[DP_MAX_RANKS] = {false};  // peer buffers mapped through CUDA IPC (closed in dflow_dp_destroy)
  bool local_group = false;             // created by dflow_dp_create_local: peers are plain device pointers
  int* d_status = nullptr;
  uint32_t epoch = 0;
  int sm_count = 148;
  long long timeout_clk = 20000000000LL;  // ~10 s at 2 GHz
  cudaStream_t own_stream = nullptr;      // dflow_dp_train_step uses it when the shard gives no stream
};

static int dp_alloc(dflow_dp* d, int device) {
  d->device = device;
  d->half_floats = (size_t)((d->P + 2 + 63) & ~63LL);
  d->bytes = DP_HDR_BYTES + 2 * d->half_floats * sizeof(float);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    set_error("no CUDA device %d: %s", device, cudaGetErrorString(cudaGetLastError()));
    return DFLOW_E_CUDA;
  }
  d->sm_count = prop.multiProcessorCount;
  if (cudaMalloc(&d->own, d->bytes) != cudaSuccess || cudaMalloc(&d->d_status, sizeof(int)) != cudaSuccess ||
      cudaMemset(d->own, 0, d->bytes) != cudaSuccess || cudaMemset(d->d_status, 0, sizeof(int)) != cudaSuccess) {
    set_error("communication buffer setup failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (d->own) cudaFree(d->own);
    if (d->d_status) cudaFree(d->d_status);
    d->own = nullptr;
    d->d_status = nullptr;
    return DFLOW_E_NOMEM;
  }
  return DFLOW_OK;
}

extern "C" {

int dflow_dp_create(int32_t rank, int32_t nranks, int64_t P, dflow_dp** out, void* ipc_handle_out) {
  if (!out || !ipc_handle_out || rank < 0 || nranks < 1 || nranks > DP_MAX_RANKS || rank >= nranks || P < 0) {
    set_error("bad dflow_dp_create arguments (at most %d ranks)", DP_MAX_RANKS);
    return DFLOW_E_INVALID_ARG;
  }
  *out = nullptr;
  dflow_dp* d = new (std::nothrow) dflow_dp();
  if (!d) return DFLOW_E_NOMEM;
  d->rank = rank;
  d->nranks = nranks;
  d->P = P;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
    delete d;
    return DFLOW_E_CUDA;
  }
  int rc = dp_alloc(d, dev);
  if (rc) {
    delete d;
    return rc;
  }
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, d->own) != cudaSuccess) {
    set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d->own);
    cudaFree(d->d_status);
    delete d;
    return DFLOW_E_CUDA;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(ipc_handle_out, &h, 64);
  d->peers[rank] = d->own;
  *out = d;
  return DFLOW_OK;
}

/* handles: nranks x 64 bytes in rank order (the own entry is ignored) */
int dflow_dp_connect(dflow_dp* d, const void* handles) {
  if (!d || !handles) {
    set_error("null argument");
    return DFLOW_E_INVALID_ARG;
  }
  for (int j = 0; j < d->nranks; ++j) {
    if (j == d->rank || d->opened[j]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(handles) + 64 * (size_t)j, 64);
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      set_error("cudaIpcOpenMemHandle(rank %d) failed: %s", j, cudaGetErrorString(cudaGetLastError()));
      return DFLOW_E_CUDA;
    }
    d->peers[j] = static_cast<char*>(p);
    d->opened[j] = true;
  }
  return DFLOW_OK;
}

/* Single-process variant (the reference's train! is one host process, src/Flows.jl:380-445): one context per device
 * of `devs`, connected through cudaDeviceEnablePeerAccess instead of IPC handles.  out: ndev handles, out[r] lives on
 * devs[r] and plays rank r.  The same device may appear more than once (replicas sharing a GPU). */
int dflow_dp_create_local(int32_t ndev, const int32_t* devs, int64_t P, dflow_dp** out) {
  if (!out || !devs || ndev < 1 || ndev > DP_MAX_RANKS || P < 0) {
    set_error("bad dflow_dp_create_local arguments (1..%d devices)", DP_MAX_RANKS);
    return DFLOW_E_INVALID_ARG;
  }
  int prev = 0, count = 0;
  if (cudaGetDevice(&prev) != cudaSuccess || cudaGetDeviceCount(&count) != cudaSuccess) {
    set_error("no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
    return DFLOW_E_CUDA;
  }
  for (int r = 0; r < ndev; ++r) {
    out[r] = nullptr;
    if (devs[r] < 0 || devs[r] >= count) {
      set_error("device %d does not exist (%d visible)", devs[r], count);
      return DFLOW_E_INVALID_ARG;
    }
  }
  int rc = DFLOW_OK;
  for (int r = 0; r < ndev && !rc; ++r) {
    dflow_dp* d = new (std::nothrow) dflow_dp();
    if (!d) {
      rc = DFLOW_E_NOMEM;
      break;
    }
    out[r] = d;
    d->rank = r;
    d->nranks = ndev;
    d->P = P;
    d->local_group = true;
    if (cudaSetDevice(devs[r]) != cudaSuccess) {
      set_error("cudaSetDevice(%d) failed: %s", devs[r], cudaGetErrorString(cudaGetLastError()));
      rc = DFLOW_E_CUDA;
      break;
    }
    rc = dp_alloc(d, devs[r]);
    if (!rc && cudaStreamCreateWithFlags(&d->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
      set_error("cudaStreamCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
      rc = DFLOW_E_CUDA;
    }
    for (int j = 0; j < ndev && !rc; ++j) {
      if (devs[j] == devs[r]) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, devs[r], devs[j]);
      if (!can) {
        set_error("device %d cannot access device %d (no NVLink / PCIe peer path)", devs[r], devs[j]);
        rc = DFLOW_E_UNSUPPORTED;
        break;
      }
      const cudaError_t e = cudaDeviceEnablePeerAccess(devs[j], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        set_error("cudaDeviceEnablePeerAccess(%d -> %d) failed: %s", devs[r], devs[j], cudaGetErrorString(e));
        rc = DFLOW_E_CUDA;
      }
      cudaGetLastError();  // clear "already enabled"
    }
  }
  if (!rc)
    for (int r = 0; r < ndev; ++r)
      for (int j = 0; j < ndev; ++j) out[r]->peers[j] = out[j]->own;
  cudaSetDevice(prev);
  if (rc)
    for (int r = 0; r < ndev; ++r) {
      dflow_dp_destroy(out[r]);
      out[r] = nullptr;
    }
  return rc;
}

/* device pointer of the [grad (P) | sum logp | #non-finite] accumulation area of the NEXT step; the caller zeroes it and
 * passes grad / loss2 pointers into it to dflow_loss_grad */
float* dflow_dp_grad_buffer(dflow_dp* d) {
  if (!d) return nullptr;
  return reinterpret_cast<float*>(d->own + DP_HDR_BYTES) + (size_t)((d->epoch + 1) & 1u) * d->half_floats;
}

/* barrier time-out of the fused kernel in milliseconds (default ~10 s); a rank whose peers do not show up in time skips
 * the reduction and the update and raises its status flag */
int dflow_dp_set_timeout_ms(dflow_dp* d, int64_t ms) {
  if (!d || ms < 1) {
    set_error("bad time-out");
    return DFLOW_E_INVALID_ARG;
  }
  d->timeout_clk = ms * 2000000LL;  // SM clock ~2 GHz
  return DFLOW_OK;
}

int dflow_dp_allreduce_adam(dflow_dp* d, float* W, float* m, float* v, float lr, float beta1, float beta2, float eps,
                            int64_t t, float* loss2_out, void* stream) {
  if (!d || !W || !m || !v || t < 1) {
    set_error("bad dflow_dp_allreduce_adam arguments");
    return DFLOW_E_INVALID_ARG;
  }
  for (int j = 0; j < d->nranks; ++j)
    if (!d->peers[j]) {
      set_error("rank %d is not connected (dflow_dp_connect)", j);
      return DFLOW_E_INVALID_ARG;
    }
  d->epoch += 1;
  DpArgs a;
  memset(&a, 0, sizeof(a));
  for (int j = 0; j < d->nranks; ++j) a.bufs[j] = d->peers[j];
  a.rank = d->rank;
  a.nranks = d->nranks;
  a.epoch = d->epoch;
  a.P = d->P;
  a.half_floats = d->half_floats;
  a.W = W;
  a.m = m;
  a.v = v;
  a.lr = lr;
  a.b1 = beta1;
  a.b2 = beta2;
  a.eps = eps;
  float b1t, b2t;  // Float32 running products like Optimisers.jl (launch_adam)
  adam_beta_powers(beta1, beta2, t, &b1t, &b2t);
  a.c1 = 1.0f - b1t;
  a.c2 = 1.0f - b2t;
  a.loss2_out = loss2_out;
  a.status = d->d_status;
  a.timeout_clk = d->timeout_clk;
  long long blocks = (d->P + 2 + 1023) / 1024;
  if (blocks > d->sm_count) blocks = d->sm_count;
  if (blocks < 1) blocks = 1;
  // cooperative launch: every CTA of the grid is resident at once (they wait for CTA 0), or the launch fails
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof(lc));
  lc.gridDim = dim3((unsigned)blocks);
  lc.blockDim = dim3(256);
  lc.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  lc.attrs = attr;
  lc.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelEx(&lc, dp_allreduce_adam_kernel, a);
  if (e != cudaSuccess) {
    set_error("dp_allreduce_adam launch failed: %s", cudaGetErrorString(e));
    d->epoch -= 1;
    return DFLOW_E_CUDA;
  }
  return DFLOW_OK;
}

/* 0 = fine, 1 = a peer did not reach the barrier in time: that step's reduction and update were skipped on this rank
 * (synchronises the stream) */
int dflow_dp_status(dflow_dp* d, void* stream) {
  if (!d) return DFLOW_E_INVALID_ARG;
  int s = 0;
  if (cudaMemcpyAsync(&s, d->d_status, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess ||
      cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) {
    set_error("status read failed: %s", cudaGetErrorString(cudaGetLastError()));
    return DFLOW_E_CUDA;
  }
  return s;
}

/* ---- single-process fan-out: one minibatch step on every device of a local group ----------------------------------
 * For every rank r (its device made current): zero the step's accumulation buffer, dflow_loss_grad on the shard with the
 * GLOBAL seed inv_btot = 1 / B_global, then the fused all-reduce + Adam kernel -- all enqueued asynchronously on the
 * shard's stream (NULL: the context's own non-blocking stream), so the host returns while the devices work; the kernels
 * of different devices meet in the peer barrier.  dflow_dp_sync waits for all of them and returns the status. */
int dflow_dp_train_step(dflow_dp* const* dps, int32_t ndev, const dflow_dp_shard* shards, float inv_btot, int32_t flags,
                        float lr, float beta1, float beta2, float eps, int64_t t) {
  if (!dps || !shards || ndev < 1 || t < 1) {
    set_error("bad dflow_dp_train_step arguments");
    return DFLOW_E_INVALID_ARG;
  }
  int prev = 0;
  cudaGetDevice(&prev);
  int rc = DFLOW_OK;
  for (int r = 0; r < ndev && !rc; ++r) {
    dflow_dp* d = dps[r];
    const dflow_dp_shard& s = shards[r];
    if (!d || d->nranks != ndev || d->rank != r || !s.chain || dflow_param_count(s.chain) != d->P) {
      set_error("shard %d does not match its data-parallel context", r);
      rc = DFLOW_E_INVALID_ARG;
      break;
    }
    if (cudaSetDevice(d->device) != cudaSuccess) {
      set_error("cudaSetDevice(%d) failed: %s", d->device, cudaGetErrorString(cudaGetLastError()));
      rc = DFLOW_E_CUDA;
      break;
    }
    cudaStream_t st = s.stream ? (cudaStream_t)s.stream : d->own_stream;
    float* buf = dflow_dp_grad_buffer(d);
    if (cudaMemsetAsync(buf, 0, sizeof(float) * (size_t)(d->P + 2), st) != cudaSuccess) {
      set_error("cudaMemsetAsync failed: %s", cudaGetErrorString(cudaGetLastError()));
      rc = DFLOW_E_CUDA;
      break;
    }
    if (s.B > 0)
      rc = dflow_loss_grad(s.chain, s.W, s.x, s.theta, s.B, s.idx, inv_btot, flags, buf + d->P, buf, s.ws, s.ws_bytes, st);
    if (!rc) rc = dflow_dp_allreduce_adam(d, s.W, s.m, s.v, lr, beta1, beta2, eps, t, s.loss2_out, st);
  }
  cudaSetDevice(prev);
  return rc;
}

int dflow_dp_sync(dflow_dp* const* dps, int32_t ndev, const dflow_dp_shard* shards) {
  if (!dps || ndev < 1) {
    set_error("bad dflow_dp_sync arguments");
    return DFLOW_E_INVALID_ARG;
  }
  int prev = 0, worst = 0;
  cudaGetDevice(&prev);
  for (int r = 0; r < ndev; ++r) {
    if (!dps[r]) continue;
    cudaSetDevice(dps[r]->device);
    cudaStream_t st = (shards && shards[r].stream) ? (cudaStream_t)shards[r].stream : dps[r]->own_stream;
    const int s = dflow_dp_status(dps[r], st);
    if (s < 0) worst = s;
    else if (s > 0 && worst == 0) worst = s;
  }
  cudaSetDevice(prev);
  return worst;
}

/* total SM clocks rank-local CTA 0 spent waiting for its peers' gradients, and the number of steps, since creation
 * (synchronises the stream) */
int dflow_dp_wait_stats(dflow_dp* d, void* stream, int64_t* wait_clk, int64_t* steps) {
  if (!d || !wait_clk || !steps) {
    set_error("null argument");
    return DFLOW_E_INVALID_ARG;
  }
  unsigned long long h[2] = {0, 0};
  if (cudaMemcpyAsync(h, d->own + 68 * sizeof(uint32_t), sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream) !=
          cudaSuccess ||
      cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) {
    set_error("stats read failed: %s", cudaGetErrorString(cudaGetLastError()));
    return DFLOW_E_CUDA;
  }
  *wait_clk = (int64_t)h[0];
  *steps = (int64_t)h[1];
  return DFLOW_OK;
}

int dflow_dp_destroy(dflow_dp* d) {
  if (!d) return DFLOW_OK;
  int prev = 0;
  cudaGetDevice(&prev);
  if (d->local_group) cudaSetDevice(d->device);
  for (int j = 0; j < d->nranks; ++j)
    if (d->opened[j] && d->peers[j]) cudaIpcCloseMemHandle(d->peers[j]);
  if (d->own_stream) cudaStreamDestroy(d->own_stream);
  if (d->own) cudaFree(d->own);
  if (d->d_status) cudaFree(d->d_status);
  if (d->local_group) cudaSetDevice(prev);
  delete d;
  return DFLOW_OK;
}

}  // extern "C"
