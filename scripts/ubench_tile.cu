// Micro-benchmark behind the "register-tiled FFMA2 conditioner for hidden 33..64" question:
// one CTA (256 threads) holds a 256-sample tile of activations and a 64 x 64 weight matrix in shared memory and evaluates
// H' = relu(W H + b) repeatedly.  Thread micro-tile 8 units x 8 samples; fma.rn.f32x2 pairs run along K (even / odd k partial
// sums in the two halves of the accumulator), so neither operand has to be duplicated; operands are stored as
// [k pair][quad j][thread group][4 floats] so that every LDS.128 of a warp is one 128-byte (or 64-byte) wavefront.
// Output: clocks per GEMM and MAC / clk / SM (FFMA2 peak = 256).
// Build (on the GPU box): nvcc -gencode arch=compute_100a,code=sm_100a -O3 --cudart shared -o ubench_tile ubench_tile.cu
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

typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ float2 unpack(u64 v) {
  float2 f;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(f.x), "=f"(f.y) : "l"(v));
  return f;
}

constexpr int H = 64, T = 256, KP = H / 2, NG = T / 8, NU = H / 8;

// swizzled position of sample group g in the row of k pair kp
__device__ __forceinline__ int gsw(int kp, int g) { return g ^ (((kp >> 2) & 7) << 2); }

__global__ void __launch_bounds__(256, 1) k_tile(const float* __restrict__ Wg, float* out, int iters, long long* clocks) {
  extern __shared__ float4 sm4[];
  ulonglong2* Wp = reinterpret_cast<ulonglong2*>(sm4);  // [KP][4][NU]
  ulonglong2* A = Wp + KP * 4 * NU;                     // [KP][4][NG]
  ulonglong2* B = A + KP * 4 * NG;
  float* bias = reinterpret_cast<float*>(B + KP * 4 * NG);
  const int tid = threadIdx.x, tu = tid & 7, g = tid >> 3;
  // weights: Wp[(kp*4 + j)*NU + tu] = (W[8tu+2j][2kp], W[8tu+2j][2kp+1], W[8tu+2j+1][2kp], W[8tu+2j+1][2kp+1])
  for (int i = tid; i < KP * 4 * NU; i += 256) {
    const int t = i % NU, j = (i / NU) & 3, kp = i / (4 * NU);
    const int u = 8 * t + 2 * j;
    float4 v = make_float4(Wg[u * H + 2 * kp], Wg[u * H + 2 * kp + 1], Wg[(u + 1) * H + 2 * kp], Wg[(u + 1) * H + 2 * kp + 1]);
    reinterpret_cast<float4*>(Wp)[i] = v;
  }
  for (int i = tid; i < KP * 4 * NG; i += 256) reinterpret_cast<float4*>(A)[i] = make_float4(0.01f * (i & 15), 0.02f, 0.03f, 0.01f);
  if (tid < H) bias[tid] = 0.001f * tid;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    u64 acc[8][8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int s = 0; s < 8; ++s) acc[u][s] = 0ull;
#pragma unroll 2
    for (int kp = 0; kp < KP; ++kp) {
      ulonglong2 a[4], w[4];
      const int gs = gsw(kp, g);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a[j] = A[(kp * 4 + j) * NG + gs];
        w[j] = Wp[(kp * 4 + j) * NU + tu];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const u64 wu = (u & 1) ? w[u >> 1].y : w[u >> 1].x;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          const u64 as = (s & 1) ? a[s >> 1].y : a[s >> 1].x;
          acc[u][s] = fma2(as, wu, acc[u][s]);
        }
      }
    }
    // epilogue: relu(sum of halves + bias) -> B as the next GEMM's operand: unit pair p of this thread = k pair 4tu + p
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const float b0 = bias[8 * tu + 2 * p], b1 = bias[8 * tu + 2 * p + 1];
      const int kp = 4 * tu + p, gs = gsw(kp, g);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 v00 = unpack(acc[2 * p][2 * j]), v10 = unpack(acc[2 * p + 1][2 * j]);
        float2 v01 = unpack(acc[2 * p][2 * j + 1]), v11 = unpack(acc[2 * p + 1][2 * j + 1]);
        float4 o;
        o.x = fmaxf(v00.x + v00.y + b0, 0.0f);
        o.y = fmaxf(v10.x + v10.y + b1, 0.0f);
        o.z = fmaxf(v01.x + v01.y + b0, 0.0f);
        o.w = fmaxf(v11.x + v11.y + b1, 0.0f);
        reinterpret_cast<float4*>(B)[(kp * 4 + j) * NG + gs] = o;
      }
    }
    __syncthreads();
    ulonglong2* tmp = A; A = B; B = tmp;
  }
  const long long t1 = clock64();
  float r = 0.0f;
  for (int i = tid; i < KP * 4 * NG; i += 256) {
    float4 v = reinterpret_cast<float4*>(A)[i];
    r += v.x + v.y + v.z + v.w;
  }
  out[blockIdx.x * 256 + tid] = r;
  if (tid == 0) clocks[blockIdx.x] = t1 - t0;
}

int main() {
  float *W, *out;
  long long* clk;
  const int nb = 148;
  CK(cudaMalloc(&W, H * H * 4));
  CK(cudaMalloc(&out, nb * 256 * 4));
  CK(cudaMalloc(&clk, nb * 8));
  float hw[H * H];
  for (int i = 0; i < H * H; ++i) hw[i] = 0.01f * ((i * 37) % 23 - 11);
  CK(cudaMemcpy(W, hw, sizeof(hw), cudaMemcpyHostToDevice));
  const size_t smem = (size_t)(KP * 4 * NU + 2 * KP * 4 * NG) * 16 + 256;
  CK(cudaFuncSetAttribute(k_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int rep = 0; rep < 2; ++rep) {
    const int iters = 2000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_tile<<<nb, 256, smem>>>(W, out, iters, clk);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[nb];
    CK(cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost));
    double c = 0;
    for (int i = 0; i < nb; ++i) c += (double)h[i];
    c /= nb;
    const double mac = (double)T * H * H;
    printf("tile GEMM 64x64 x 256 samples: %.0f clk per GEMM, %.1f MAC/clk/SM (FFMA2 peak 256), %.3f ms, %.2f TMAC/s\n", c / iters,
           mac / (c / iters), ms, mac * iters * nb / (ms * 1e-3) / 1e12);
  }
  return 0;
}
