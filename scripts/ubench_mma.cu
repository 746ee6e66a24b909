// Micro-benchmark behind the "tensor cores for narrow conditioners?" decision:
//   (1) raw issue rate of mma.sync.m16n8k8 TF32 (the warp-level "legacy" tensor path, SASS HMMA.1688.F32.TF32) on sm_100a;
//   (2) a register-chained conditioner MLP  K0 -> H -> H -> 8  evaluated with 3xTF32 (hi*hi + hi*lo + lo*hi): the C
//       fragment of one Dense IS the A fragment of the next (the K order of every hidden layer is permuted so that
//       slot t <-> unit 2t, slot t+4 <-> unit 2t+1), bias enters as the accumulator init, relu + hi/lo split in
//       registers, weights pre-arranged in fragment order in shared memory (one conflict-free LDS.128 per B fragment
//       pair hi/lo), two 16-sample m-tiles per warp share every B fragment;
// The comparison point is the product's FFMA2 kernels (conditioner evaluations per second = samples/s x conditioners per sample).
// Output: MMA/clk/SM and conditioner evaluations per second for H = 16, 32, 64.
// Build (on the GPU box): nvcc -gencode arch=compute_100a,code=sm_100a -O3 --cudart shared -o ubench_mma ubench_mma.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x)                                                                     \
  do {                                                                            \
    cudaError_t e = (x);                                                          \
    if (e != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      return 1;                                                                   \
    }                                                                             \
  } while (0)

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t tf32_hi(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// ---- (1) raw rate: NACC independent accumulator tiles per warp ---------------------------------------------------
template <int NACC>
__global__ void k_mma_rate(float* out, int iters, long long* clocks) {
  float acc[NACC][4];
  uint32_t a[4], b0 = threadIdx.x * 7u + 1u, b1 = threadIdx.x * 13u + 5u;
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + 0.001f * (threadIdx.x + i)) & 0xFFFFE000u;
  b0 = __float_as_uint(0.5f + 0.001f * threadIdx.x) & 0xFFFFE000u;
  b1 = __float_as_uint(0.25f + 0.002f * threadIdx.x) & 0xFFFFE000u;
#pragma unroll
  for (int i = 0; i < NACC; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) mma_tf32(acc[i], a, b0, b1);
  }
  const long long t1 = clock64();
  float r = 0.0f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) r += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
}

// ---- (2) register-chained 3xTF32 conditioner -----------------------------------------------------------------------
// Weight image in shared memory, fragment order: for every (dense, kstep, ntile): 32 lanes x float4 {b0h, b1h, b0l, b1l};
// biases: per (dense, ntile): 32 lanes x float2 {bias[8j+2t], bias[8j+2t+1]}.
template <int H>
struct MmaNet {
  static constexpr int KS1 = 1, NT1 = H / 8;       // 8 -> H
  static constexpr int KS2 = H / 8, NT2 = H / 8;   // H -> H
  static constexpr int KS3 = H / 8, NT3 = 1;       // H -> 8
  static constexpr int W_F4 = 32 * (KS1 * NT1 + KS2 * NT2 + KS3 * NT3);
  static constexpr int B_F2 = 32 * (NT1 + NT2 + NT3);
};

template <int H, int MT>
__device__ __forceinline__ void split_frag(const float (&c)[MT][H / 8][4], int j, int act, uint32_t (&ah)[MT][4],
                                           uint32_t (&al)[MT][4]) {
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    // A(kstep j) = {a0 = c0, a1 = c2, a2 = c1, a3 = c3} of the n-tile j of the previous Dense
    const float v[4] = {c[m][j][0], c[m][j][2], c[m][j][1], c[m][j][3]};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float x = act ? fmaxf(v[q], 0.0f) : v[q];
      ah[m][q] = tf32_hi(x);
      al[m][q] = __float_as_uint(x - __uint_as_float(ah[m][q]));
    }
  }
}

template <int H, int MT>
__global__ void __launch_bounds__(256) k_mma_chain(float* out, int iters, long long* clocks, float seed) {
  using N = MmaNet<H>;
  extern __shared__ float4 sm4[];
  float4* wimg = sm4;
  float2* bimg = reinterpret_cast<float2*>(wimg + N::W_F4);
  for (int i = threadIdx.x; i < N::W_F4; i += blockDim.x) {
    const float w0 = 0.03f * ((i * 7) % 13 - 6) * seed, w1 = 0.02f * ((i * 5) % 11 - 5) * seed;
    const uint32_t h0 = tf32_hi(w0), h1 = tf32_hi(w1);
    wimg[i] = make_float4(__uint_as_float(h0), __uint_as_float(h1), w0 - __uint_as_float(h0), w1 - __uint_as_float(h1));
  }
  for (int i = threadIdx.x; i < N::B_F2; i += blockDim.x) bimg[i] = make_float2(0.01f * (i % 7), -0.01f * (i % 5));
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float accum = 0.0f;
  float xin = 0.1f * lane + seed;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    // input fragments of MT m-tiles (K0 = 8: theta + identity coordinates, zero padded), split hi/lo
    uint32_t ah[MT][4], al[MT][4];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float x = xin + 0.01f * (q + 4 * m);
        ah[m][q] = tf32_hi(x);
        al[m][q] = __float_as_uint(x - __uint_as_float(ah[m][q]));
      }
    xin += 0.001f;
    const float4* wp = wimg + lane;
    const float2* bp = bimg + lane;
    float c1[MT][H / 8][4], c2[MT][H / 8][4], c3[MT][4];
    // Dense 1: 8 -> H
#pragma unroll
    for (int j = 0; j < N::NT1; ++j) {
      const float2 b = bp[32 * j];
      const float4 w = wp[32 * j];
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        c1[m][j][0] = c1[m][j][2] = b.x;
        c1[m][j][1] = c1[m][j][3] = b.y;
        mma_tf32(c1[m][j], al[m], __float_as_uint(w.x), __float_as_uint(w.y));
        mma_tf32(c1[m][j], ah[m], __float_as_uint(w.z), __float_as_uint(w.w));
        mma_tf32(c1[m][j], ah[m], __float_as_uint(w.x), __float_as_uint(w.y));
      }
    }
    wp += 32 * N::KS1 * N::NT1;
    bp += 32 * N::NT1;
    // Dense 2: H -> H
#pragma unroll
    for (int j = 0; j < N::NT2; ++j) {
      const float2 b = bp[32 * j];
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        c2[m][j][0] = c2[m][j][2] = b.x;
        c2[m][j][1] = c2[m][j][3] = b.y;
      }
    }
#pragma unroll
    for (int k = 0; k < N::KS2; ++k) {
      split_frag<H, MT>(c1, k, 1, ah, al);
#pragma unroll
      for (int j = 0; j < N::NT2; ++j) {
        const float4 w = wp[32 * (k * N::NT2 + j)];
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          mma_tf32(c2[m][j], al[m], __float_as_uint(w.x), __float_as_uint(w.y));
          mma_tf32(c2[m][j], ah[m], __float_as_uint(w.z), __float_as_uint(w.w));
          mma_tf32(c2[m][j], ah[m], __float_as_uint(w.x), __float_as_uint(w.y));
        }
      }
    }
    wp += 32 * N::KS2 * N::NT2;
    bp += 32 * N::NT2;
    // Dense 3: H -> 8
    {
      const float2 b = bp[0];
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        c3[m][0] = c3[m][2] = b.x;
        c3[m][1] = c3[m][3] = b.y;
      }
    }
#pragma unroll
    for (int k = 0; k < N::KS3; ++k) {
      split_frag<H, MT>(c2, k, 1, ah, al);
      const float4 w = wp[32 * k];
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        mma_tf32(c3[m], al[m], __float_as_uint(w.x), __float_as_uint(w.y));
        mma_tf32(c3[m], ah[m], __float_as_uint(w.z), __float_as_uint(w.w));
        mma_tf32(c3[m], ah[m], __float_as_uint(w.x), __float_as_uint(w.y));
      }
    }
#pragma unroll
    for (int m = 0; m < MT; ++m) accum += c3[m][0] + c3[m][1] + c3[m][2] + c3[m][3];
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = accum;
  if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
}

static double max_clk(long long* d_clk, int n) {
  static long long h[4096];
  cudaMemcpy(h, d_clk, sizeof(long long) * n, cudaMemcpyDeviceToHost);
  long long m = 0;
  for (int i = 0; i < n; ++i) m = h[i] > m ? h[i] : m;
  return (double)m;
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  float* out;
  long long* clk;
  CK(cudaMalloc(&out, sizeof(float) * 4096 * 1024));
  CK(cudaMalloc(&clk, sizeof(long long) * 4096));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  printf("SMs: %d\n", sms);
  // (1) raw mma.sync rate, one CTA per SM, NW warps, 8 independent accumulators per warp
  for (int nw : {4, 8, 16, 32}) {
    const int iters = 4096;
    k_mma_rate<8><<<sms, nw * 32>>>(out, 64, clk);
    CK(cudaDeviceSynchronize());
    k_mma_rate<8><<<sms, nw * 32>>>(out, iters, clk);
    CK(cudaDeviceSynchronize());
    const double c = max_clk(clk, sms);
    const double mma_per_clk = (double)iters * 8 * nw / c;
    printf("mma.sync m16n8k8 tf32: %2d warps/SM: %.3f MMA/clk/SM = %.0f MAC/clk/SM (3xTF32 fp32-equivalent %.0f MAC/clk/SM; FFMA peak 128)\n",
           nw, mma_per_clk, mma_per_clk * 1024, mma_per_clk * 1024 / 3);
  }
  // (2) / (3) conditioner chains
  auto report = [&](const char* what, int H, double evals, float ms, double clocks) {
    const double macs = (8.0 * H + (double)H * H + H * 8.0) * evals;
    printf("%-28s H=%2d: %.3e conditioner evals/s, %.1f fp32-equivalent TFLOP/s, %.1f MAC/clk/SM\n", what, H,
           evals / (ms * 1e-3), 2 * macs / (ms * 1e-3) / 1e12, macs / clocks / sms);
  };
#define RUN_MMA(H, MT, NW, CPS)                                                                                    \
  {                                                                                                                \
    using N = MmaNet<H>;                                                                                           \
    const size_t smem = N::W_F4 * 16 + N::B_F2 * 8;                                                                \
    CK(cudaFuncSetAttribute(k_mma_chain<H, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));          \
    const int iters = 2000;                                                                                        \
    k_mma_chain<H, MT><<<sms * CPS, NW * 32, smem>>>(out, 10, clk, 1.0f);                                          \
    CK(cudaDeviceSynchronize());                                                                                   \
    cudaEventRecord(e0);                                                                                           \
    k_mma_chain<H, MT><<<sms * CPS, NW * 32, smem>>>(out, iters, clk, 1.0f);                                       \
    cudaEventRecord(e1);                                                                                           \
    CK(cudaDeviceSynchronize());                                                                                   \
    float ms;                                                                                                      \
    cudaEventElapsedTime(&ms, e0, e1);                                                                             \
    char nm[64];                                                                                                   \
    snprintf(nm, sizeof nm, "mma 3xTF32 MT=%d %dw x%d", MT, NW, CPS);                                              \
    report(nm, H, (double)iters * 16 * MT * NW * sms * CPS, ms, max_clk(clk, sms* CPS) * CPS);                     \
  }
  RUN_MMA(16, 1, 8, 2)
  RUN_MMA(16, 2, 8, 2)
  RUN_MMA(16, 2, 8, 4)
  RUN_MMA(16, 4, 8, 2)
  RUN_MMA(32, 1, 8, 2)
  RUN_MMA(32, 2, 8, 2)
  RUN_MMA(32, 2, 8, 3)
  RUN_MMA(64, 1, 8, 2)
  RUN_MMA(64, 1, 8, 3)
  RUN_MMA(64, 2, 8, 1)
  return 0;
}
