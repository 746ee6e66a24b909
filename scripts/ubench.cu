// Micro-benchmarks that decide the narrow-kernel design: FFMA issue rate by operand form and the cost of
// broadcast shared-memory / constant loads on B200.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cuda_runtime.h>
#include <stdio.h>

__constant__ float cw[4096];

#define ITERS 2048

// 3-register FFMA: both multiplicands and the accumulator are vector registers
__global__ void k_ffma_rrr(float* out, float s) {
  float acc[16], x[4];
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
  for (int i = 0; i < 4; ++i) x[i] = s + i + threadIdx.x;
  float w0 = s * 0.5f, w1 = s * 0.25f, w2 = s * 0.125f, w3 = s * 0.0625f;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(x[i & 3], (i & 4) ? ((i & 8) ? w3 : w2) : ((i & 8) ? w1 : w0), acc[i]);
  }
  float r = 0;
  for (int i = 0; i < 16; ++i) r += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// FFMA with a constant-bank operand at a compile-time offset
__global__ void k_ffma_const(float* out, float s) {
  float acc[16], x[4];
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
  for (int i = 0; i < 4; ++i) x[i] = s + i + threadIdx.x;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(x[i & 3], cw[i + 16 * (it & 0)], acc[i]);
  }
  float r = 0;
  for (int i = 0; i < 16; ++i) r += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// FFMA with weights fetched from constant memory at a RUNTIME (warp-uniform) offset: ULDC + FFMA R,R,UR,R
__global__ void k_ffma_uldc(float* out, float s, int stride) {
  float acc[16], x[4];
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
  for (int i = 0; i < 4; ++i) x[i] = s + i + threadIdx.x;
  int off = 0;
  for (int it = 0; it < ITERS; ++it) {
    const float4* w4 = reinterpret_cast<const float4*>(cw + off);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float4 w = w4[g];
      acc[4 * g + 0] = fmaf(x[0], w.x, acc[4 * g + 0]);
      acc[4 * g + 1] = fmaf(x[1], w.y, acc[4 * g + 1]);
      acc[4 * g + 2] = fmaf(x[2], w.z, acc[4 * g + 2]);
      acc[4 * g + 3] = fmaf(x[3], w.w, acc[4 * g + 3]);
    }
    off = (off + stride) & 2047;
  }
  float r = 0;
  for (int i = 0; i < 16; ++i) r += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// same, weights from shared memory via broadcast LDS.128 (S samples per thread share each weight)
template <int S>
__global__ void k_ffma_lds(float* out, float s, int stride) {
  __shared__ float4 sw4[1024];
  float* sw = reinterpret_cast<float*>(sw4);
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sw[i] = s * i;
  __syncthreads();
  float acc[16][S], x[S];
  for (int i = 0; i < 16; ++i)
    for (int q = 0; q < S; ++q) acc[i][q] = threadIdx.x * 0.001f + i + q;
  for (int q = 0; q < S; ++q) x[q] = s + q + threadIdx.x;
  int off = 0;
  for (int it = 0; it < ITERS; ++it) {
    const float4* w4 = reinterpret_cast<const float4*>(sw + off);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float4 w = w4[g];
#pragma unroll
      for (int q = 0; q < S; ++q) {
        acc[4 * g + 0][q] = fmaf(x[q], w.x, acc[4 * g + 0][q]);
        acc[4 * g + 1][q] = fmaf(x[q], w.y, acc[4 * g + 1][q]);
        acc[4 * g + 2][q] = fmaf(x[q], w.z, acc[4 * g + 2][q]);
        acc[4 * g + 3][q] = fmaf(x[q], w.w, acc[4 * g + 3][q]);
      }
    }
    off = (off + stride) & 2047;
  }
  float r = 0;
  for (int i = 0; i < 16; ++i)
    for (int q = 0; q < S; ++q) r += acc[i][q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// pure broadcast LDS throughput (VEC = 1, 2, 4 floats per load)
template <int VEC>
__global__ void k_lds_bcast(float* out, int stride) {
  __shared__ float4 sw4[1024];
  float* sw = reinterpret_cast<float*>(sw4);
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sw[i] = i;
  __syncthreads();
  float r0 = 0, r1 = 0, r2 = 0, r3 = 0;
  int off = 0;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      if (VEC == 4) {
        const float4 w = *reinterpret_cast<const float4*>(sw + off + 4 * g);
        r0 += w.x; r1 += w.y; r2 += w.z; r3 += w.w;
      } else if (VEC == 2) {
        const float2 w = *reinterpret_cast<const float2*>(sw + off + 2 * g);
        r0 += w.x; r1 += w.y;
      } else {
        r0 += sw[off + g];
      }
    }
    off = (off + stride) & 2047;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r0 + r1 + r2 + r3;
}

// per-lane (non-broadcast) LDS.32 column reads
__global__ void k_lds_lane(float* out, int stride) {
  __shared__ float sw[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sw[i] = i;
  __syncthreads();
  float r0 = 0;
  int off = threadIdx.x;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int g = 0; g < 8; ++g) r0 += sw[(off + g * 256) & 8191];
    off = (off + stride) & 8191;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r0;
}

template <class F>
static float timeit(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / 5;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("SMs %d, max clock %.0f MHz\n", sms, clk_khz / 1e3);
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
  float hw[4096];
  for (int i = 0; i < 4096; ++i) hw[i] = 1e-3f * i;
  cudaMemcpyToSymbol(cw, hw, sizeof(hw));
  for (int bps = 2; bps <= 8; bps *= 2) {
    const int grid = sms * bps, nt = 256;
    const double warps = (double)grid * nt / 32;
    auto rep = [&](const char* name, float ms, double ffma_per_thread, double lds_per_thread) {
      const double clk = 1.9e9;  // nominal; compare ratios
      printf("bps=%d %-18s %8.3f ms  FFMA/clk/SM(@1.9GHz)=%7.1f  LDS-instr/clk/SM=%6.3f\n", bps, name, ms,
             ffma_per_thread * grid * nt / (ms * 1e-3) / clk / sms, lds_per_thread * warps / (ms * 1e-3) / clk / sms);
    };
    rep("ffma_rrr", timeit([&] { k_ffma_rrr<<<grid, nt>>>(out, 1.0f); }), 16.0 * ITERS, 0);
    rep("ffma_const", timeit([&] { k_ffma_const<<<grid, nt>>>(out, 1.0f); }), 16.0 * ITERS, 0);
    rep("ffma_uldc", timeit([&] { k_ffma_uldc<<<grid, nt>>>(out, 1.0f, 16); }), 16.0 * ITERS, 4.0 * ITERS);
    rep("ffma_lds S=1", timeit([&] { k_ffma_lds<1><<<grid, nt>>>(out, 1.0f, 16); }), 16.0 * ITERS, 4.0 * ITERS);
    rep("ffma_lds S=2", timeit([&] { k_ffma_lds<2><<<grid, nt>>>(out, 1.0f, 16); }), 32.0 * ITERS, 4.0 * ITERS);
    rep("ffma_lds S=4", timeit([&] { k_ffma_lds<4><<<grid, nt>>>(out, 1.0f, 16); }), 64.0 * ITERS, 4.0 * ITERS);
    rep("ffma_lds S=8", timeit([&] { k_ffma_lds<8><<<grid, nt>>>(out, 1.0f, 16); }), 128.0 * ITERS, 4.0 * ITERS);
    rep("lds_bcast.32", timeit([&] { k_lds_bcast<1><<<grid, nt>>>(out, 16); }), 0, 8.0 * ITERS);
    rep("lds_bcast.64", timeit([&] { k_lds_bcast<2><<<grid, nt>>>(out, 16); }), 0, 8.0 * ITERS);
    rep("lds_bcast.128", timeit([&] { k_lds_bcast<4><<<grid, nt>>>(out, 16); }), 0, 8.0 * ITERS);
    rep("lds_lane.32", timeit([&] { k_lds_lane<<<grid, nt>>>(out, 1); }), 0, 8.0 * ITERS);
  }
  printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
