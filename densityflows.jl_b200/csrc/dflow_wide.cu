// K6: wide conditioners (hidden 64 < h <= 512) on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a.
//
// One launch per coupling layer (x: 4d B/sample round trip per layer is negligible against ~1e5..1e6 FLOP/sample),
// CTA = 128 threads = one 128-sample MMA tile, thread t <-> sample row t <-> TMEM lane t.
//
// Float32 parity (rtol 1e-5) on a TF32 tensor core needs split accumulation: every operand is split into
// hi = cvt.rna.tf32(x) and lo = x - hi, and each product is issued as three MMAs  hi*hi + hi*lo + lo*hi
// (the dropped lo*lo term is ~2^-22 relative).  Accumulators stay in TMEM (fp32).
//
// Per conditioner (Dense(in,h,relu) -> Dense(h,h,relu) -> Dense(h,a)) and sample tile:
//   for each pass over <= 256 output columns of Dense 2 (h = 512 needs two: D2 would fill all 512 TMEM columns):
//     for each chunk c of 32 hidden-1 units:
//        D1c[128x32]   = A1[128xKin] * W1c^T            (tcgen05.mma, N = 32)          TMEM cols [0,32)
//        epilogue      : tcgen05.ld -> +b1, relu, split -> A2 (hi, lo) in shared memory (UMMA K-major core layout)
//        D2[128xNH]   += A2[128x32] * W2c^T             (N = NH <= 256)                TMEM cols [64,64+NH)
//     for each chunk cc of 32 hidden-2 units of this pass:
//        epilogue      : tcgen05.ld D2 chunk -> +b2, relu, split -> A2
//        D3[128xa16]  += A2 * W3cc^T                    (N = 16 or 32)                 TMEM cols [32,64)
//   s / t = D3 + b3
// Weights are pre-split (hi, lo) and pre-arranged per chunk in the UMMA no-swizzle K-major core-matrix layout by
// wide_prepack_kernel, so a CTA stages a chunk with a linear float4 copy.  This first version serialises
// copy -> MMA -> epilogue inside the CTA (no TMA / warp specialisation / 2-CTA pairs yet): see DESIGN.md §4.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "dflow_chain_kernels.cuh"
#include "dflow_wide.h"

namespace dflow {

constexpr int WKC = 32;   // hidden units per chunk (K of Dense 2 / Dense 3 MMAs, N of Dense 1 MMAs)
constexpr int WNH = 256;  // max Dense-2 output columns per pass

// float index of element (r, k) of a [R x Kc] K-major operand block in the no-swizzle UMMA core-matrix layout:
// 8 x 16-byte core matrices, LBO (next core along K) = 128 B, SBO (next 8 rows) = (Kc/4) * 128 B
__host__ __device__ inline int core_idx(int r, int k, int Kc) {
  return (r >> 3) * (Kc >> 2) * 32 + (k >> 2) * 32 + (r & 7) * 4 + (k & 3);
}

// ---- PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// shared-memory matrix descriptor: K-major, no swizzle (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout_type SWIZZLE_NONE
}
// instruction descriptor: kind::tf32, fp32 accumulate, A and B K-major, M = 128 (cute::UMMA::InstrDescriptor)
__device__ __forceinline__ uint32_t instr_desc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// D[128 x n] (+)= A[128 x K] * B[n x K]^T with the 3xTF32 split; operands in K-major core layout, K multiple of 8
__device__ __forceinline__ void gemm_3xtf32(uint32_t d_tmem, const float* a_hi, const float* a_lo, int a_kc,
                                            const float* b_hi, const float* b_lo, int b_kc, int K, int n, bool accumulate) {
  const uint32_t idesc = instr_desc_tf32(n);
  const uint32_t a_sbo = (uint32_t)(a_kc >> 2) * 128u, b_sbo = (uint32_t)(b_kc >> 2) * 128u;
  const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
  uint32_t acc = accumulate ? 1u : 0u;
  for (int ks = 0; ks < K; ks += 8) {
    const uint32_t off = (uint32_t)(ks >> 2) * 128u;  // two 128-byte core matrices per K step of 8
    const uint64_t dah = smem_desc(ah + off, 128u, a_sbo), dal = smem_desc(al + off, 128u, a_sbo);
    const uint64_t dbh = smem_desc(bh + off, 128u, b_sbo), dbl = smem_desc(bl + off, 128u, b_sbo);
    mma_tf32(d_tmem, dal, dbh, idesc, acc);  // small terms first
    mma_tf32(d_tmem, dah, dbl, idesc, 1u);
    mma_tf32(d_tmem, dah, dbh, idesc, 1u);
    acc = 1u;
  }
}

// ---- prepack: packed Flux parameters -> per-layer wide image (hi/lo split, core layout, chunked) -------------
__global__ void wide_prepack_kernel(const WideLayer* layers, int L, const float* __restrict__ W, float* __restrict__ img) {
  const WideLayer& Ld = layers[blockIdx.x];
  if (!Ld.is_coupling) return;
  const int h = Ld.h, KinP = Ld.kinp, a = Ld.a, a16 = Ld.a16, nin = Ld.nin;
  const int NH = h < WNH ? h : WNH, passes = h / NH, nch = h / WKC;
  for (int ni = (Ld.has_s ? 0 : 1); ni < 2; ++ni) {
    const WideNet& N = Ld.net[ni];
    float* base = img + N.img_off;
    // biases
    for (int i = threadIdx.x; i < h; i += blockDim.x) {
      base[N.b1 + i] = W[N.p_b[0] + i];
      base[N.b2 + i] = W[N.p_b[1] + i];
    }
    for (int i = threadIdx.x; i < a16; i += blockDim.x) base[N.b3 + i] = i < a ? W[N.p_b[2] + i] : 0.0f;
    // Dense 1: chunk c holds B = W1[c*32 .. c*32+31][0..KinP) as [32 x KinP]
    for (int i = threadIdx.x; i < h * KinP; i += blockDim.x) {
      const int o = i / KinP, k = i - o * KinP;
      const float w = k < nin ? W[N.p_w[0] + o + h * k] : 0.0f;  // Flux (out,in) column-major
      const float hi = to_tf32(w);
      const int c = o / WKC, r = o - c * WKC;
      float* blk = base + N.w1 + (size_t)c * (2 * WKC * KinP);
      blk[core_idx(r, k, KinP)] = hi;
      blk[WKC * KinP + core_idx(r, k, KinP)] = w - hi;
    }
    // Dense 2: (pass p, chunk c) holds B = W2[p*NH .. p*NH+NH-1][c*32 .. c*32+31] as [NH x 32]
    for (int i = threadIdx.x; i < h * h; i += blockDim.x) {
      const int o = i / h, k = i - o * h;
      const float w = W[N.p_w[1] + o + h * k];
      const float hi = to_tf32(w);
      const int p = o / NH, r = o - p * NH, c = k / WKC, kk = k - c * WKC;
      float* blk = base + N.w2 + ((size_t)p * nch + c) * (2 * NH * WKC);
      blk[core_idx(r, kk, WKC)] = hi;
      blk[NH * WKC + core_idx(r, kk, WKC)] = w - hi;
    }
    // Dense 3: chunk cc holds B = W3[0..a16)[cc*32 .. cc*32+31] as [a16 x 32]
    for (int i = threadIdx.x; i < a16 * h; i += blockDim.x) {
      const int o = i / h, k = i - o * h;
      const float w = o < a ? W[N.p_w[2] + o + a * k] : 0.0f;
      const float hi = to_tf32(w);
      const int c = k / WKC, kk = k - c * WKC;
      float* blk = base + N.w3 + (size_t)c * (2 * a16 * WKC);
      blk[core_idx(o, kk, WKC)] = hi;
      blk[a16 * WKC + core_idx(o, kk, WKC)] = w - hi;
    }
    (void)passes;
  }
}

// ---- the layer kernel --------------------------------------------------------------------------------------
struct WideArgs {
  WideLayer layer;
  const float* img;
  float* x;            // (d, B) in place
  const float* theta;  // (n, B) or null
  const float* theta_const;
  float* ldj;          // (B) accumulated, or null
  long long B;
  int d, n;
  int sampling;
  int flags;
  float theta_min[NMAX], theta_rng[NMAX];
};

__device__ __forceinline__ void copy_lin(float* dst, const float* __restrict__ src, int nfloats, int tid) {
  float4* d4 = reinterpret_cast<float4*>(dst);
  const float4* s4 = reinterpret_cast<const float4*>(src);
  for (int i = tid; i < (nfloats >> 2); i += 128) d4[i] = __ldg(s4 + i);
}

// write this thread's row of an activation chunk (32 values) as hi / lo operands (K-major core layout, Kc = 32)
__device__ __forceinline__ void store_a2_row(float* a_hi, float* a_lo, int row, const float (&v)[16], int k0) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float4 hi, lo;
    hi.x = to_tf32(v[4 * q + 0]); lo.x = v[4 * q + 0] - hi.x;
    hi.y = to_tf32(v[4 * q + 1]); lo.y = v[4 * q + 1] - hi.y;
    hi.z = to_tf32(v[4 * q + 2]); lo.z = v[4 * q + 2] - hi.z;
    hi.w = to_tf32(v[4 * q + 3]); lo.w = v[4 * q + 3] - hi.w;
    const int idx = core_idx(row, k0 + 4 * q, WKC);
    *reinterpret_cast<float4*>(a_hi + idx) = hi;
    *reinterpret_cast<float4*>(a_lo + idx) = lo;
  }
}

__global__ void __launch_bounds__(128, 1) wide_layer_kernel(const WideArgs a) {
  extern __shared__ float4 smem4[];
  float* smem = reinterpret_cast<float*>(smem4);
  const int tid = threadIdx.x, warp = tid >> 5;
  const WideLayer& Ld = a.layer;
  const int h = Ld.h, KinP = Ld.kinp, a16 = Ld.a16, d = a.d, n = a.n;
  const int NH = h < WNH ? h : WNH, passes = h / NH, nch = h / WKC, nch_pass = NH / WKC;

  // shared-memory carve-up (floats)
  float* A1h = smem;                     // [128 x KinP]
  float* A1l = A1h + 128 * KinP;
  float* A2h = A1l + 128 * KinP;         // [128 x 32]
  float* A2l = A2h + 128 * WKC;
  float* W1b = A2l + 128 * WKC;          // hi | lo  [32 x KinP] each
  float* W2b = W1b + 2 * WKC * KinP;     // hi | lo  [NH x 32] each
  float* W3b = W2b + 2 * NH * WKC;       // hi | lo  [a16 x 32] each
  float* outs = W3b + 2 * a16 * WKC;     // [2][a16][128] s / t values of the tile
  uint64_t* bar = reinterpret_cast<uint64_t*>(outs + 2 * a16 * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  if (warp == 0) tmem_alloc(tmem_slot, 512);
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;  // warp w owns TMEM lanes 32w .. 32w+31
  const uint32_t tD1 = tbase + 0, tD3 = tbase + 32, tD2 = tbase + 64;
  uint32_t phase = 0;

  const long long ntiles = (a.B + 127) / 128;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long gi = tile * 128 + tid;
    const bool valid = gi < a.B;
    // ---- A1: this sample's conditioner input row  [θ_0..θ_{n-1}, x[axis_id...], 0 pad]  (src/affine/RNVP.jl:157) ----
    for (int k0 = 0; k0 < KinP; k0 += 4) {
      float v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = k0 + q;
        float val = 0.0f;
        if (valid && k < Ld.nin) {
          if (k < n) {
            val = a.theta_const ? __ldg(a.theta_const + k) : __ldg(a.theta + gi * n + k);
            if (a.flags & DFLOW_THETA_NORMALIZE) val = (a.theta_rng[k] == 0.0f) ? 0.0f : (val - a.theta_min[k]) / a.theta_rng[k];
          } else {
            val = a.x[gi * d + Ld.id[k - n]];
          }
        }
        v[q] = val;
      }
      float4 hi, lo;
      hi.x = to_tf32(v[0]); lo.x = v[0] - hi.x;
      hi.y = to_tf32(v[1]); lo.y = v[1] - hi.y;
      hi.z = to_tf32(v[2]); lo.z = v[2] - hi.z;
      hi.w = to_tf32(v[3]); lo.w = v[3] - hi.w;
      const int idx = core_idx(tid, k0, KinP);
      *reinterpret_cast<float4*>(A1h + idx) = hi;
      *reinterpret_cast<float4*>(A1l + idx) = lo;
    }

    for (int ni = (Ld.has_s ? 0 : 1); ni < 2; ++ni) {
      const WideNet& N = Ld.net[ni];
      const float* base = a.img + N.img_off;
      for (int p = 0; p < passes; ++p) {
        for (int c = 0; c < nch; ++c) {
          // stage W1 chunk c and W2 chunk (p, c)
          copy_lin(W1b, base + N.w1 + (size_t)c * (2 * WKC * KinP), 2 * WKC * KinP, tid);
          copy_lin(W2b, base + N.w2 + ((size_t)p * nch + c) * (2 * NH * WKC), 2 * NH * WKC, tid);
          fence_async_smem();
          tc_fence_before();
          __syncthreads();
          if (tid == 0) {
            tc_fence_after();
            gemm_3xtf32(tD1, A1h, A1l, KinP, W1b, W1b + WKC * KinP, KinP, KinP, WKC, false);
            mma_commit(bar);
          }
          mbar_wait(bar, phase);
          phase ^= 1;
          tc_fence_after();
          // epilogue 1: h1 chunk = relu(D1c + b1) -> A2 (hi, lo)
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float v[16];
            tmem_ld16(tD1 + lane_off + half * 16, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j] + __ldg(base + N.b1 + c * WKC + half * 16 + j), 0.0f);
            store_a2_row(A2h, A2l, tid, v, half * 16);
          }
          fence_async_smem();
          tc_fence_before();
          __syncthreads();
          if (tid == 0) {
            tc_fence_after();
            gemm_3xtf32(tD2, A2h, A2l, WKC, W2b, W2b + NH * WKC, WKC, WKC, NH, c > 0);
            mma_commit(bar);
          }
          mbar_wait(bar, phase);
          phase ^= 1;
          tc_fence_after();
        }
        // Dense 3 over the hidden-2 units of this pass
        for (int cc = 0; cc < nch_pass; ++cc) {
          const int gc = p * nch_pass + cc;  // global hidden-2 chunk
          copy_lin(W3b, base + N.w3 + (size_t)gc * (2 * a16 * WKC), 2 * a16 * WKC, tid);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float v[16];
            tmem_ld16(tD2 + lane_off + cc * WKC + half * 16, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j] + __ldg(base + N.b2 + gc * WKC + half * 16 + j), 0.0f);
            store_a2_row(A2h, A2l, tid, v, half * 16);
          }
          fence_async_smem();
          tc_fence_before();
          __syncthreads();
          if (tid == 0) {
            tc_fence_after();
            gemm_3xtf32(tD3, A2h, A2l, WKC, W3b, W3b + a16 * WKC, WKC, WKC, a16, gc > 0);
            mma_commit(bar);
          }
          mbar_wait(bar, phase);
          phase ^= 1;
          tc_fence_after();
        }
      }
      // outputs of this conditioner: D3 + b3
      for (int o0 = 0; o0 < a16; o0 += 16) {
        float v[16];
        tmem_ld16(tD3 + lane_off + o0, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) outs[(ni * a16 + o0 + j) * 128 + tid] = v[j] + __ldg(base + N.b3 + o0 + j);
      }
      tc_fence_before();
      __syncthreads();  // D3 / D2 are about to be overwritten by the next conditioner
      tc_fence_after();
    }

    // ---- coupling transform on this sample (src/affine/RNVP.jl:92,184; NICE: s = 0) ----
    if (valid) {
      float lsum = 0.0f;
      for (int j = 0; j < Ld.a; ++j) {
        const int k = Ld.af[j];
        const float sv = Ld.has_s ? outs[(0 * a16 + j) * 128 + tid] : 0.0f;
        const float tv = outs[(1 * a16 + j) * 128 + tid];
        float* p = a.x + gi * d + k;
        const float xv = *p;
        *p = a.sampling ? xv * expf(sv) + tv : (xv - tv) * expf(-sv);
        lsum += sv;
      }
      if (a.ldj) a.ldj[gi] += a.sampling ? lsum : -lsum;
    }
    __syncthreads();  // outs / A1 reused by the next tile
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

// ---- small element-wise kernels ------------------------------------------------------------------------------
__global__ void wide_norm_kernel(float* x, float* ldj, long long B, int d, const float* __restrict__ blk, int sampling) {
  // blk = [x_min(d) | x_max(d) | alpha, beta, ldj_const]   (src/norm/Normalization.jl:64-103)
  const float alpha = blk[2 * d], beta = blk[2 * d + 1], c = blk[2 * d + 2];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < B * d; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % d);
    const float xmin = blk[k], xmax = blk[d + k], v = x[i];
    x[i] = sampling ? ((xmax - xmin) * v - alpha * xmax + beta * xmin) / (beta - alpha)
                    : (beta * (v - xmin) + alpha * (xmax - v)) / (xmax - xmin);
    if (k == 0 && ldj) ldj[i / d] += sampling ? c : -c;
  }
}

__global__ void wide_logpdf_kernel(const float* __restrict__ z, const float* __restrict__ ldj, long long B, int d, float c0,
                                   float* out, float* sum2) {
  float ls = 0.0f, bad = 0.0f;
  for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
    float q = 0.0f;
    for (int k = 0; k < d; ++k) {
      const float v = z[b * d + k];
      q = fmaf(v, v, q);
    }
    const float lp = c0 - 0.5f * q + ldj[b];  // src/Flows.jl:279
    if (out) out[b] = lp;
    if (isfinite(lp))
      ls += lp;
    else
      bad += 1.0f;
  }
  if (sum2) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ls += __shfl_xor_sync(0xffffffffu, ls, o);
      bad += __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(sum2, ls);
      if (bad != 0.0f) atomicAdd(sum2 + 1, bad);
    }
  }
}

__global__ void wide_gather_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, long long B, int rows,
                                   float* dst) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < B * rows; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / rows;
    dst[i] = src[(long long)idx[b] * rows + (i - b * rows)];
  }
}

__global__ void wide_philox_kernel(float* z, long long B, int d, unsigned long long seed, unsigned int offset,
                                   unsigned long long first) {
  for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
    const unsigned long long ctr = first + (unsigned long long)b;
    for (int g = 0; g < (d + 3) / 4; ++g) {
      unsigned int r[4];
      philox4x32_10((unsigned int)ctr, (unsigned int)(ctr >> 32), (unsigned int)g, offset, (unsigned int)seed,
                    (unsigned int)(seed >> 32), r);
      float v[4];
      box_muller(r[0], r[1], v[0], v[1]);
      box_muller(r[2], r[3], v[2], v[3]);
      for (int q = 0; q < 4; ++q)
        if (4 * g + q < d) z[b * d + 4 * g + q] = v[q];
    }
  }
}

// ---- host side -----------------------------------------------------------------------------------------------
#define CKW(call)                                                                          \
  do {                                                                                     \
    cudaError_t _e = (call);                                                               \
    if (_e != cudaSuccess) {                                                               \
      set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DFLOW_E_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

size_t wide_layer_smem_bytes(const WideLayer& Ld) {
  const int NH = Ld.h < WNH ? Ld.h : WNH;
  size_t f = 2 * 128 * (size_t)Ld.kinp + 2 * 128 * WKC + 2 * WKC * (size_t)Ld.kinp + 2 * (size_t)NH * WKC +
             2 * (size_t)Ld.a16 * WKC + 2 * (size_t)Ld.a16 * 128;
  return f * sizeof(float) + 64;
}

int wide_prepack(dflow_chain* c, const float* W, cudaStream_t st) {
  WidePlan* wp = c->wide;
  wide_prepack_kernel<<<(unsigned)wp->layers.size(), 256, 0, st>>>(wp->d_layers, (int)wp->layers.size(), W, wp->d_img);
  CKW(cudaGetLastError());
  c->launches++;
  return DFLOW_OK;
}

// Runs the whole chain on `x` in place (x already holds the input), accumulating ldj.
int wide_run_chain(dflow_chain* c, float* x, const float* theta, const float* theta_const, float* ldj, long long B,
                   int sampling, int flags, cudaStream_t st) {
  WidePlan* wp = c->wide;
  const DevChainHdr& H = c->hc()->h;
  const int L = (int)wp->layers.size();
  for (int step = 0; step < L; ++step) {
    const int ei = sampling ? step : (L - 1 - step);  // src/Chains.jl:155-161 vs :174-180
    const WideLayer& Ld = wp->layers[ei];
    if (!Ld.is_coupling) {
      long long blocks = (B * H.d + 255) / 256;
      if (blocks > 148 * 8) blocks = 148 * 8;
      wide_norm_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, ldj, B, H.d, c->d_staged + Ld.norm_off, sampling);
      CKW(cudaGetLastError());
      c->launches++;
      continue;
    }
    WideArgs a;
    a.layer = Ld;
    a.img = wp->d_img;
    a.x = x;
    a.theta = theta;
    a.theta_const = theta_const;
    a.ldj = ldj;
    a.B = B;
    a.d = H.d;
    a.n = H.n;
    a.sampling = sampling;
    a.flags = flags;
    for (int k = 0; k < NMAX; ++k) {
      a.theta_min[k] = H.theta_min[k];
      a.theta_rng[k] = H.theta_rng[k];
    }
    const size_t smem = wide_layer_smem_bytes(Ld);
    if (smem > (size_t)c->max_smem_optin) {
      set_error("wide layer needs %zu bytes of shared memory (> %d)", smem, c->max_smem_optin);
      return DFLOW_E_UNSUPPORTED;
    }
    CKW(cudaFuncSetAttribute(wide_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = (B + 127) / 128;
    if (grid > c->sm_count) grid = c->sm_count;
    wide_layer_kernel<<<(unsigned)grid, 128, smem, st>>>(a);
    CKW(cudaGetLastError());
    c->launches++;
  }
  return DFLOW_OK;
}

int wide_logpdf(dflow_chain* c, const float* z, const float* ldj, long long B, float* out, float* sum2, cudaStream_t st) {
  const DevChainHdr& H = c->hc()->h;
  long long blocks = (B + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  wide_logpdf_kernel<<<(unsigned)blocks, 256, 0, st>>>(z, ldj, B, H.d, H.logpdf_c0, out, sum2);
  CKW(cudaGetLastError());
  c->launches++;
  return DFLOW_OK;
}

int wide_gather(dflow_chain* c, const float* src, const int32_t* idx, long long B, int rows, float* dst, cudaStream_t st) {
  if (rows == 0 || B == 0) return DFLOW_OK;
  long long blocks = (B * rows + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  wide_gather_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, idx, B, rows, dst);
  CKW(cudaGetLastError());
  c->launches++;
  return DFLOW_OK;
}

int wide_philox(dflow_chain* c, float* z, long long B, unsigned long long seed, unsigned int offset,
                unsigned long long first, cudaStream_t st) {
  const DevChainHdr& H = c->hc()->h;
  long long blocks = (B + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  wide_philox_kernel<<<(unsigned)blocks, 256, 0, st>>>(z, B, H.d, seed, offset, first);
  CKW(cudaGetLastError());
  c->launches++;
  return DFLOW_OK;
}

}  // namespace dflow
