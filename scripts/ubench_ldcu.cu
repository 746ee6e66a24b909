// Micro-benchmark: FFMA2 fed by (A) broadcast LDS.128 weights vs (B) uniform constant loads (LDCU), activations from
// shared-memory columns (LDS.128 per thread, 4 adjacent samples).  Prints FMA lanes per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>
__constant__ float4 CW[4096];
__device__ __forceinline__ void fma2_pair(float& r0, float& r1, float x0, float x1, float w, float c0, float c1) {
  unsigned long long x, ww, c, r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(x0), "f"(x1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w));
  asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(c0), "f"(c1));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(ww), "l"(c));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(r));
}
template <int MODE>  // 0: smem weights, 1: constant weights, 2: smem weights scalar FFMA
__global__ void __launch_bounds__(128, 4) k(float4* out, int K, int reps, int base) {
  extern __shared__ float4 sm[];
  float4* col = sm;                 // [K][128]
  float4* ws = sm + 16 * 128;       // [K][4]
  const int i = threadIdx.x;
  for (int j = i; j < 16 * 128; j += 128) col[j] = make_float4(1e-3f * j, 1.f, 2.f, 3.f);
  for (int j = i; j < 64 * 4; j += 128) ws[j] = make_float4(1e-3f, 2e-3f, 3e-3f, 1e-4f * j);
  __syncthreads();
  float acc[16][4];
  for (int o = 0; o < 16; ++o) for (int s = 0; s < 4; ++s) acc[o][s] = 0.f;
  for (int r = 0; r < reps; ++r) {
    const float4* wr = CW + base;
#pragma unroll 2
    for (int k = 0; k < K; ++k) {
      float4 v4 = col[k * 128 + i];
      float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float4 w = MODE == 1 ? wr[k * 4 + g] : ws[k * 4 + g];
        if (MODE == 2) {
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            acc[4 * g + 0][s] = fmaf(w.x, v[s], acc[4 * g + 0][s]);
            acc[4 * g + 1][s] = fmaf(w.y, v[s], acc[4 * g + 1][s]);
            acc[4 * g + 2][s] = fmaf(w.z, v[s], acc[4 * g + 2][s]);
            acc[4 * g + 3][s] = fmaf(w.w, v[s], acc[4 * g + 3][s]);
          }
        } else {
#pragma unroll
          for (int s = 0; s < 4; s += 2) {
            fma2_pair(acc[4 * g + 0][s], acc[4 * g + 0][s + 1], v[s], v[s + 1], w.x, acc[4 * g + 0][s], acc[4 * g + 0][s + 1]);
            fma2_pair(acc[4 * g + 1][s], acc[4 * g + 1][s + 1], v[s], v[s + 1], w.y, acc[4 * g + 1][s], acc[4 * g + 1][s + 1]);
            fma2_pair(acc[4 * g + 2][s], acc[4 * g + 2][s + 1], v[s], v[s + 1], w.z, acc[4 * g + 2][s], acc[4 * g + 2][s + 1]);
            fma2_pair(acc[4 * g + 3][s], acc[4 * g + 3][s + 1], v[s], v[s + 1], w.w, acc[4 * g + 3][s], acc[4 * g + 3][s + 1]);
          }
        }
      }
    }
  }
  for (int o = 0; o < 16; ++o) out[(blockIdx.x * 16 + o) * 128 + i] = make_float4(acc[o][0], acc[o][1], acc[o][2], acc[o][3]);
}
template <int MODE>
void run(const char* name, int ctas_per_sm, float4* out) {
  const int K = 16, reps = 4000;
  size_t smem = (16 * 128 + 64 * 4) * sizeof(float4);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (ctas_per_sm == 1) smem = 200 * 1024 > smem ? smem : smem;
  int grid = 148 * ctas_per_sm;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<grid, 128, smem>>>(out, K, 10, 0);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<grid, 128, smem>>>(out, K, reps, 0);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double fmas = (double)grid * 128 * 64.0 * K * reps;
  printf("%-28s ctas/SM %d: %.3f ms, %.1f FMA/clk/SM (at %d MHz nominal), err=%s\n", name, ctas_per_sm, ms,
         fmas / (ms * 1e-3) / (clk * 1e3) / 148.0, clk / 1000, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  float4* out; cudaMalloc(&out, 148 * 4 * 16 * 128 * sizeof(float4));
  static float4 h[4096];
  for (int i = 0; i < 4096; ++i) h[i] = make_float4(1e-3f, 2e-3f, 3e-3f, 1e-4f * i);
  cudaMemcpyToSymbol(CW, h, sizeof(h));
  for (int c = 1; c <= 4; ++c) {
    run<2>("smem weights, FFMA", c, out);
    run<0>("smem weights, FFMA2", c, out);
    run<1>("constant weights (LDCU), FFMA2", c, out);
  }
  return 0;
}
