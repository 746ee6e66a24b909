// Launchers and small kernels (prepack, Adam, min/max) of the narrow path; the chain kernels themselves are
// instantiated one (HP, S) pair per translation unit (dflow_inst.cu) so that they build in parallel.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>

#include "dflow_chain_kernels.cuh"
#include "dflow_grad_kernel.cuh"

namespace dflow {

// ------------------------------------------------------------------------------------------------------------
// prepack: packed Flux parameters -> padded staged image ([in][out4] per Dense, zero padded)
// ------------------------------------------------------------------------------------------------------------
__global__ void prepack_kernel(const PrepackArgs a) {
  const DevChain* C = a.chain;
  const DevElem& E = C->e[blockIdx.x];
  if (E.kind == DFLOW_ELEM_NORM) return;
  float* blk = a.staged + E.stage_off;
  for (int ni = 0; ni < 2; ++ni) {
    const DevNet& net = ni == 0 ? E.s : E.t;
    for (int j = 0; j < net.depth; ++j) {
      const int K = net.w[j], O = net.w[j + 1], op = net.op[j];
      const int Kp = (j == 0) ? K : C->h.hp;  // rows of hidden Dense layers are zero-padded to hp
      for (int i = threadIdx.x; i < Kp * op; i += blockDim.x) {
        const int k = i / op, o = i - k * op;
        blk[net.s_w[j] + i] = (o < O && k < K) ? a.W[net.p_w[j] + o + O * k] : 0.0f;
      }
      for (int o = threadIdx.x; o < op; o += blockDim.x)
        blk[net.s_b[j] + o] = (net.has_bias && o < O) ? a.W[net.p_b[j] + o] : 0.0f;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// K4: Optimisers.Adam (call site src/Flows.jl:415)
// ------------------------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ W, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long P, float lr, float b1, float b2, float eps, float c1,
                            float c2) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) {
    // explicit round-to-nearest ops (no FMA contraction): bit-identical to the un-fused Float32 broadcast of
    // Optimisers.jl  mt = β1*mt + (1-β1)*dx ; vt = β2*vt + (1-β2)*dx^2 ; dx' = mt/(1-β1^t)/(sqrt(vt/(1-β2^t))+ϵ)*η
    const float gi = g[i];
    const float mi = __fadd_rn(__fmul_rn(b1, m[i]), __fmul_rn(1.0f - b1, gi));
    const float vi = __fadd_rn(__fmul_rn(b2, v[i]), __fmul_rn(1.0f - b2, __fmul_rn(gi, gi)));
    m[i] = mi;
    v[i] = vi;
    const float den = __fadd_rn(__fsqrt_rn(__fdiv_rn(vi, c2)), eps);
    const float stepv = __fmul_rn(__fdiv_rn(__fdiv_rn(mi, c1), den), lr);
    W[i] = __fsub_rn(W[i], stepv);
  }
}

// ------------------------------------------------------------------------------------------------------------
// K5: per-row min / max
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_min_f(float* addr, float v) {
  if (v >= 0.0f)
    atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* addr, float v) {
  if (v >= 0.0f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void minmax_init_kernel(float* mn, float* mx, int rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows) {
    mn[i] = INFINITY;
    mx[i] = -INFINITY;
  }
}

__global__ void __launch_bounds__(256) minmax_kernel(const float* __restrict__ x, int rows, long long B, float* mn,
                                                     float* mx) {
  extern __shared__ float4 smem4[];
  float* smin = reinterpret_cast<float*>(smem4);
  float* smax = smin + rows * 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  // one warp per strip of samples; lane keeps per-row running values in shared memory [row][lane] of warp 0's
  // layout, reduced at the end.  Elements are read flat (coalesced).
  for (int i = tid; i < rows * 32; i += blockDim.x) {
    smin[i] = INFINITY;
    smax[i] = -INFINITY;
  }
  __syncthreads();
  const long long total = B * rows;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long e = blockIdx.x * (long long)blockDim.x + tid;
  int row = (int)(e % rows);
  const int rstep = (int)(stride % rows);
  (void)warp;
  (void)nw;
  for (; e < total; e += stride) {
    const float v = __ldg(x + e) + 0.0f;  // -0.0 -> +0.0 (keeps the integer-ordered atomics exact)
    // shared atomics keep this simple; the kernel is a set-up time reduction, HBM-bound
    atomic_min_f(&smin[row * 32 + lane], v);
    atomic_max_f(&smax[row * 32 + lane], v);
    row += rstep;
    if (row >= rows) row -= rows;
  }
  __syncthreads();
  for (int r = tid; r < rows; r += blockDim.x) {
    float a = INFINITY, b = -INFINITY;
    for (int l = 0; l < 32; ++l) {
      a = fminf(a, smin[r * 32 + l]);
      b = fmaxf(b, smax[r * 32 + l]);
    }
    atomic_min_f(mn + r, a);
    atomic_max_f(mx + r, b);
  }
}

// ------------------------------------------------------------------------------------------------------------
// K7: stateless pseudo-random permutation (DataPartition randperm src/Data.jl:112-128, DataLoader shuffle
// src/Flows.jl:394).  perm_i = cycle-walked balanced Feistel network over 2*hb >= log2(n) bits: a bijection of [0, n)
// that any thread (and any rank, for its own slice) evaluates independently from (seed, i) -- no sort, no state.
// Specification shared with the oracle: oracle/shuffle.py.
// ------------------------------------------------------------------------------------------------------------
__host__ __device__ inline uint32_t feistel_mix(uint32_t x) {  // murmur3 finaliser
  x ^= x >> 16;
  x *= 0x85EBCA6Bu;
  x ^= x >> 13;
  x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}
__host__ __device__ inline uint64_t feistel_perm(uint64_t i, uint64_t n, int hb, const uint32_t* key) {
  const uint32_t mask = hb >= 32 ? 0xFFFFFFFFu : ((1u << hb) - 1u);
  uint64_t x = i;
  do {
    uint32_t l = (uint32_t)(x >> hb) & mask, r = (uint32_t)x & mask;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      const uint32_t t = l ^ (feistel_mix(r ^ key[q]) & mask);
      l = r;
      r = t;
    }
    x = ((uint64_t)l << hb) | r;
  } while (x >= n);
  return x;
}

struct ShuffleArgs {
  uint32_t key[6];
  long long n, first, count;
  int hb;
  const int32_t* base;
  int32_t* out;
};

__global__ void shuffle_kernel(const ShuffleArgs a) {
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < a.count; j += (long long)gridDim.x * blockDim.x) {
    const long long p = (long long)feistel_perm((uint64_t)(a.first + j), (uint64_t)a.n, a.hb, a.key);
    a.out[j] = a.base ? a.base[p] : (int32_t)p;
  }
}

int launch_shuffle(unsigned long long seed, long long n, long long first, long long count, const int32_t* base, int32_t* out,
                   cudaStream_t st) {
  ShuffleArgs a;
  uint64_t z = seed;
  for (int q = 0; q < 6; ++q) {  // splitmix64 key schedule
    z += 0x9E3779B97F4A7C15ull;
    uint64_t y = z;
    y = (y ^ (y >> 30)) * 0xBF58476D1CE4E5B9ull;
    y = (y ^ (y >> 27)) * 0x94D049BB133111EBull;
    y ^= y >> 31;
    a.key[q] = (uint32_t)(y >> 32);
  }
  int bits = 1;
  while (bits < 62 && (1ull << bits) < (unsigned long long)n) ++bits;
  a.hb = (bits + 1) / 2;
  if (a.hb < 1) a.hb = 1;
  a.n = n;
  a.first = first;
  a.count = count;
  a.base = base;
  a.out = out;
  long long blocks = (count + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  shuffle_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
  if (cudaGetLastError() != cudaSuccess) {
    set_error("shuffle_kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return DFLOW_E_CUDA;
  }
  return DFLOW_OK;
}

// ------------------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------------------
#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t _e = (call);                                                              \
    if (_e != cudaSuccess) {                                                              \
      set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DFLOW_E_CUDA;                                                                \
    }                                                                                     \
  } while (0)

int launch_prepack(dflow_chain* c, const float* W, cudaStream_t st) {
  PrepackArgs pa{c->d_chain, W, c->d_staged};
  prepack_kernel<<<c->hc()->h.L, 256, 0, st>>>(pa);
  CK(cudaGetLastError());
  c->launches++;
  return DFLOW_OK;
}

template <int HP, int S, bool REG>
static int launch_fwd_t(dflow_chain* c, FwdArgs& a, cudaStream_t st, int nt) {
  const DevChainHdr& h = c->hc()->h;
  if (nt > fwd_max_threads<HP, S, REG>()) nt = fwd_max_threads<HP, S, REG>();
  SmemPlan p = plan_fwd(h, c->chain_bytes, nt * S, REG);
  while (p.bytes() > (size_t)c->max_smem_optin && nt > 32) {
    nt >>= 1;
    p = plan_fwd(h, c->chain_bytes, nt * S, REG);
  }
  if (p.bytes() > (size_t)c->max_smem_optin) {
    set_error("chain needs %zu bytes of shared memory (> %d)", p.bytes(), c->max_smem_optin);
    return DFLOW_E_UNSUPPORTED;
  }
  const long long ntiles = (a.B + (long long)nt * S - 1) / ((long long)nt * S);
  int per_sm = c->ctas_per_sm;
  if (per_sm <= 0) {
    per_sm = (int)((size_t)c->max_smem_optin / (p.bytes() + 1024));
    const int by_threads = (REG ? 512 : 2048) / nt;
    if (per_sm > by_threads) per_sm = by_threads;
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
  }
  long long grid = (long long)c->sm_count * per_sm;
  if (grid > ntiles) grid = ntiles;
  if (grid < 1) grid = 1;
  CK((launch_fwd_inst<HP, S, REG>(a, (unsigned)grid, nt, p.bytes(), st)));
  c->launches++;
  return DFLOW_OK;
}

// Constant-bank forward kernel (dflow_chain_kernels.cuh, DFLOW_CBANK): descriptor + staged weights in the 60 KB bank
template <int HP, int S>
static int launch_fwd_const_t(dflow_chain* c, FwdArgs& a, cudaStream_t st) {
  const DevChainHdr& h = c->hc()->h;
  int nt = fwd_max_threads<HP, S, true>();
  const SmemPlan p = plan_fwd(h, c->chain_bytes, nt * S, true, true);
  const long long ntiles = (a.B + (long long)nt * S - 1) / ((long long)nt * S);
  int per_sm = c->ctas_per_sm > 0 ? c->ctas_per_sm : 3;
  long long grid = (long long)c->sm_count * per_sm;
  if (grid > ntiles) grid = ntiles;
  if (grid < 1) grid = 1;
  a.cb_wofs = ((c->chain_bytes + 15) / 16) * 4;
  CK((launch_fwd_const_inst<HP, S>(a, (unsigned)grid, nt, p.bytes(), st, h.stage_total)));
  c->launches++;  // kernels only: the two bank uploads are device-to-device memcpy nodes, not kernels
  return DFLOW_OK;
}

// Kernel selection (fwd_spt tuning: 0 = automatic, else samples per thread of the shared-memory-column kernels)
int launch_fwd(dflow_chain* c, FwdArgs& a, cudaStream_t st) {
  const DevChainHdr& h = c->hc()->h;
  a.chain = c->d_chain;
  a.staged = c->d_staged;
  a.chain_bytes = c->chain_bytes;
  int spt = c->fwd_spt;
  // default for chains that fit the constant bank: weights as uniform-datapath operands, activations in registers
  if (spt == 0 && c->fwd_const >= 0 && c->cbank_ok) {
    if (h.hp == 16) return launch_fwd_const_t<16, 4>(c, a, st);
    if (h.hp == 32) return launch_fwd_const_t<32, 2>(c, a, st);
  }
  if (spt < 0) spt = -spt;
  int nt = c->fwd_threads > 0 ? c->fwd_threads : 256;  // clamped to the instantiation's fixed block size
  if (nt > 256) nt = 256;
  nt = (nt + 31) & ~31;
  // spt > 0: samples per thread of the shared-memory-column kernels; 0: automatic
  switch (h.hp) {
    case 16:
      if (spt == 0 || spt == 4) return launch_fwd_t<16, 4, false>(c, a, st, nt);
      if (spt == 2) return launch_fwd_t<16, 2, false>(c, a, st, nt);
      return launch_fwd_t<16, 1, false>(c, a, st, nt);
    case 32:
      if (spt == 0 || spt == 4) return launch_fwd_t<32, 4, false>(c, a, st, nt);
      if (spt == 2) return launch_fwd_t<32, 2, false>(c, a, st, nt);
      return launch_fwd_t<32, 1, false>(c, a, st, nt);
    case 64:
      if (spt == 0 || spt == 2) return launch_fwd_t<64, 2, false>(c, a, st, nt);
      return launch_fwd_t<64, 1, false>(c, a, st, nt);
  }
  set_error("hidden width template %d not built", h.hp);
  return DFLOW_E_UNSUPPORTED;
}

template <int HP, int S>
static int launch_grad2_t(dflow_chain* c, GradArgs& a, cudaStream_t st, int nt) {
  const DevChainHdr& h = c->hc()->h;
  if (nt > grad2_max_threads<HP, S>()) nt = grad2_max_threads<HP, S>();
  a.smem_grad = (h.P * 4 <= 64 * 1024 && c->grad_smem >= 0) ? 1 : 0;
  const int want_th = (a.thbar_out != nullptr && h.n > 0) ? 1 : 0;
  SmemPlan p = plan_grad2(h, c->chain_bytes, nt * S, a.smem_grad, want_th);
  while (p.bytes() > (size_t)c->max_smem_optin && nt > 32) {
    nt = (nt > 256) ? (nt > 384 ? 384 : 256) : nt >> 1;  // 448 -> 384 -> 256 -> 128 -> ...
    p = plan_grad2(h, c->chain_bytes, nt * S, a.smem_grad, want_th);
  }
  if (p.bytes() > (size_t)c->max_smem_optin) {
    set_error("adjoint needs %zu bytes of shared memory (> %d)", p.bytes(), c->max_smem_optin);
    return DFLOW_E_UNSUPPORTED;
  }
  const long long ntiles = (a.B + (long long)nt * S - 1) / ((long long)nt * S);
  int per_sm = c->ctas_per_sm;
  if (per_sm <= 0) {
    per_sm = (int)((size_t)c->max_smem_optin / (p.bytes() + 1024));
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
  }
  long long grid = (long long)c->sm_count * per_sm;
  if (grid > (long long)c->sm_count * 4) grid = (long long)c->sm_count * 4;  // workspace is sized for <= 4 CTAs/SM
  if (grid > ntiles) grid = ntiles;
  if (grid < 1) grid = 1;
  CK((launch_grad2_inst<HP, S>(a, (unsigned)grid, nt, p.bytes(), st)));
  c->launches++;
  return DFLOW_OK;
}

int launch_grad(dflow_chain* c, GradArgs& a, cudaStream_t st) {
  const DevChainHdr& h = c->hc()->h;
  a.chain = c->d_chain;
  a.staged = c->d_staged;
  a.chain_bytes = c->chain_bytes;
  int nt2 = c->grad_threads > 0 ? c->grad_threads : 1024;  // clamped to the instantiation's fixed block size
  nt2 = (nt2 + 31) & ~31;
  int s = c->grad_spt;
  // automatic: a minibatch that fits one wave of 256-thread CTAs runs one sample per thread -- a tile's latency is what
  // such a step costs (85 us against 134 us for the S = 2 / 384-thread configuration at the reference's batchsize 64)
  if (s == 0 && h.hp == 16 && a.B <= (long long)c->sm_count * 256) s = 1;
  switch (h.hp) {
    case 16:
      if (s == 1) return launch_grad2_t<16, 1>(c, a, st, nt2);
      if (s == 4) return launch_grad2_t<16, 4>(c, a, st, nt2);
      return launch_grad2_t<16, 2>(c, a, st, nt2);
    case 32:
      if (s == 1) return launch_grad2_t<32, 1>(c, a, st, nt2);
      return launch_grad2_t<32, 2>(c, a, st, nt2);
    case 64:
      if (s == 2) return launch_grad2_t<64, 2>(c, a, st, nt2);
      return launch_grad2_t<64, 1>(c, a, st, nt2);
  }
  set_error("hidden width template %d not built", h.hp);
  return DFLOW_E_UNSUPPORTED;
}

__global__ void axpy2_kernel(float* acc, const float* v) {
  if (threadIdx.x < 2) acc[threadIdx.x] += v[threadIdx.x];
}
int launch_axpy2(float* acc, const float* v, cudaStream_t st) {
  axpy2_kernel<<<1, 32, 0, st>>>(acc, v);
  CK(cudaGetLastError());
  return DFLOW_OK;
}

// Float32 running products beta^t of Optimisers.jl, as the host keeps them (bit-exact with t successive multiplications)
void adam_beta_powers(float b1, float b2, long long t, float* b1t, float* b2t) {
  float p1 = 1.0f, p2 = 1.0f;
  for (long long i = 0; i < t; ++i) {
    p1 *= b1;
    p2 *= b2;
    if (p1 == 0.0f && p2 == 0.0f) break;  // both underflowed: every further product is 0 as well
  }
  *b1t = p1;
  *b2t = p2;
}

int launch_adam(float* W, const float* g, float* m, float* v, long long P, float lr, float b1, float b2, float eps,
                long long t, cudaStream_t st) {
  float b1t, b2t;
  adam_beta_powers(b1, b2, t, &b1t, &b2t);
  return launch_adam_pw(W, g, m, v, P, lr, b1, b2, eps, b1t, b2t, st);
}

int launch_adam_pw(float* W, const float* g, float* m, float* v, long long P, float lr, float b1, float b2, float eps,
                   float b1t, float b2t, cudaStream_t st) {
  if (P <= 0) return DFLOW_OK;
  // bias corrections in Float32 like Optimisers.jl (βt is a Float32 running product)
  const float c1 = 1.0f - b1t, c2 = 1.0f - b2t;
  long long blocks = (P + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  adam_kernel<<<(unsigned)blocks, 256, 0, st>>>(W, g, m, v, P, lr, b1, b2, eps, c1, c2);
  CK(cudaGetLastError());
  return DFLOW_OK;
}

int launch_minmax(const float* x, int rows, long long B, float* mn, float* mx, cudaStream_t st) {
  minmax_init_kernel<<<(rows + 127) / 128, 128, 0, st>>>(mn, mx, rows);
  CK(cudaGetLastError());
  if (B <= 0) return DFLOW_OK;
  const long long total = B * rows;
  long long blocks = (total + 256 * 8 - 1) / (256 * 8);
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  const size_t smem = (size_t)rows * 32 * 2 * sizeof(float);
  minmax_kernel<<<(unsigned)blocks, 256, smem, st>>>(x, rows, B, mn, mx);
  CK(cudaGetLastError());
  return DFLOW_OK;
}

}  // namespace dflow
