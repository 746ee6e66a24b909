// Narrow-conditioner (hidden <= 64) fused coupling-chain kernels for sm_100a (templates).
//
// One thread owns S samples.  The whole FlowChain (all coupling layers + NormalizationLayer + Gaussian base term)
// runs inside one kernel, so x / z never round-trip HBM between layers (reference: one (d,B) allocation per
// element, src/Chains.jl:149-197).  Conditioner weights live in shared memory in a padded [in][out4] image
// (broadcast LDS.128 -> 4 FFMA per sample), per-sample state lives in bank-conflict-free shared-memory columns
// col[unit * CS + slot], and the FFMA accumulators live in registers.
//
// Reference math: src/affine/RNVP.jl:77-96 (normalising), :99-147 (adjoint), :168-205 (sampling);
// src/affine/NICE.jl:63-170; src/norm/Normalization.jl:64-103; src/Flows.jl:272-281,352-359; src/Data.jl:213-218.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include <type_traits>

#include "dflow_internal.h"

namespace dflow {

// ------------------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------------------

__device__ __forceinline__ float act_apply(int code, float v) {
  switch (code) {
    case DFLOW_ACT_RELU: return fmaxf(v, 0.0f);
    case DFLOW_ACT_TANH: return tanhf(v);
    case DFLOW_ACT_SIGMOID: return 1.0f / (1.0f + expf(-v));
    default: return v;
  }
}

// derivative w.r.t. the pre-activation, expressed through the OUTPUT y; relu'(0) = 0 (NNlib convention)
__device__ __forceinline__ float act_grad(int code, float y) {
  switch (code) {
    case DFLOW_ACT_RELU: return y > 0.0f ? 1.0f : 0.0f;
    case DFLOW_ACT_TANH: return 1.0f - y * y;
    case DFLOW_ACT_SIGMOID: return y * (1.0f - y);
    default: return 1.0f;
  }
}

// Selects the shared-memory column holding input row k of a Dense: for the first Dense the rows are
// vcat(θ, x)[axis_nn] = θ_0..θ_{n-1}, x[axis_id[0]], ... (src/affine/RNVP.jl:157); otherwise the hidden column.
struct InSel {
  const float* th;
  const float* xs;
  const float* hc;
  const unsigned char* id;
  int n;
  int CS;
  bool first;
  __device__ __forceinline__ const float* operator()(int k) const {
    if (!first) return hc + k * CS;
    return k < n ? th + k * CS : xs + (int)id[k - n] * CS;
  }
};

// The S samples of a thread sit next to each other in every column (slot = tid*S + s), so a thread moves its S
// values of one unit with a single 32/64/128-bit shared-memory access (conflict-free: consecutive lanes touch
// consecutive S*4-byte chunks) and the address arithmetic is one add per unit instead of one LEA per sample.
template <int S>
__device__ __forceinline__ void ld_samples(const float* p, float (&v)[S]) {
  if constexpr (S == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if constexpr (S == 2) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    v[0] = t.x; v[1] = t.y;
  } else {
#pragma unroll
    for (int s = 0; s < S; ++s) v[s] = p[s];
  }
}
template <int S>
__device__ __forceinline__ void st_samples(float* p, const float (&v)[S]) {
  if constexpr (S == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else if constexpr (S == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  } else {
#pragma unroll
    for (int s = 0; s < S; ++s) p[s] = v[s];
  }
}

// Packed FP32 FMA (sm_100 FFMA2): one issue slot performs two IEEE fmas, acc[s..s+1] = v[s..s+1] * w + c[s..s+1], on an
// even-aligned register pair; the scalar weight is a broadcast operand of the instruction (`R.F32`), so pairing over
// the thread's adjacent samples costs no packing moves (the mov.b64 below are register-allocation hints that ptxas
// folds away).  Each lane rounds exactly like fmaf: results are bit-identical to the scalar form.
__device__ __forceinline__ void fma2_pair(float& r0, float& r1, float x0, float x1, float w, float c0, float c1) {
  unsigned long long x, ww, c, r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(x0), "f"(x1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w));
  asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(c0), "f"(c1));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(ww), "l"(c));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(r));
}
// (r0, r1) += (x0 * y0, x1 * y1): both factors are register pairs
__device__ __forceinline__ void fma2_vv(float& r0, float& r1, float x0, float x1, float y0, float y1) {
  unsigned long long x, y, c, r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(x0), "f"(x1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(y0), "f"(y1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(r0), "f"(r1));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(y), "l"(c));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(r));
}
// acc[s] = v[s] * w + acc[s] for the S samples of a thread
template <int S>
__device__ __forceinline__ void fma_samples(float (&acc)[S], const float (&v)[S], float w) {
  if constexpr (S % 2 == 0) {
#pragma unroll
    for (int s = 0; s < S; s += 2) fma2_pair(acc[s], acc[s + 1], v[s], v[s + 1], w, acc[s], acc[s + 1]);
  } else {
#pragma unroll
    for (int s = 0; s < S; ++s) acc[s] = fmaf(w, v[s], acc[s]);
  }
}
// acc[s] = v[s] * w + b (first row of a Dense: the bias enters as the addend)
template <int S>
__device__ __forceinline__ void fma_samples_bias(float (&acc)[S], const float (&v)[S], float w, float b) {
  if constexpr (S % 2 == 0) {
#pragma unroll
    for (int s = 0; s < S; s += 2) fma2_pair(acc[s], acc[s + 1], v[s], v[s + 1], w, b, b);
  } else {
#pragma unroll
    for (int s = 0; s < S; ++s) acc[s] = fmaf(w, v[s], b);
  }
}

// acc[s] = b, written as 1.0 * b + 0 (exact) so that b stays a broadcast operand of the FFMA2: the constant-bank kernels
// must never need a bias or weight in a vector register, or ptxas switches the whole conditioner from uniform constant
// loads (LDCU) to per-thread indexed LDC (4x too slow; see profiles/r01_ncu_fwd_summary.md)
template <int S>
__device__ __forceinline__ void bias_samples(float (&acc)[S], float b) {
  static_assert(S % 2 == 0, "packed pairs");
#pragma unroll
  for (int s = 0; s < S; s += 2) fma2_pair(acc[s], acc[s + 1], 1.0f, 1.0f, b, 0.0f, 0.0f);
}

// acc[o][s] += W[k][o] * in_k[s] for one input row k (NG broadcast LDS.128 -> NG*4*S FMAs = NG*2*S FFMA2)
template <int NG, int S>
__device__ __forceinline__ void dense_row(const float* src, const float* __restrict__ wrow, int slot0, int NT,
                                          float (&acc)[NG * 4][S]) {
  float v[S];
  ld_samples<S>(src + slot0 * S, v);
  const float4* wr = reinterpret_cast<const float4*>(wrow);
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const float4 w = wr[g];
    fma_samples<S>(acc[4 * g + 0], v, w.x);
    fma_samples<S>(acc[4 * g + 1], v, w.y);
    fma_samples<S>(acc[4 * g + 2], v, w.z);
    fma_samples<S>(acc[4 * g + 3], v, w.w);
  }
}

// Outputs [o0, o0 + NG*4) of a Dense: acc = bias + sum_k W[k][o] * in_k, activation, store to outcol rows o0...
// W is staged as [K][ld] (ld = padded output width), so a chunk of the outputs is a column slice.
template <int NG, int S>
__device__ __forceinline__ void dense_to_col(const InSel& in, int K, const float* __restrict__ Wst, int ld,
                                             const float* __restrict__ bst, int act, float* outcol, int CS, int slot0,
                                             int NT) {
  float acc[NG * 4][S];
  // row 0 is peeled so that the bias enters as the FFMA addend (no per-sample register copies of the bias)
  {
    const float* src0 = in.first ? (in.n > 0 ? in.th : in.xs + (int)in.id[0] * in.CS) : in.hc;
    float v[S];
    ld_samples<S>(src0 + slot0 * S, v);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const float4 b = *reinterpret_cast<const float4*>(bst + 4 * g);
      const float4 w = *reinterpret_cast<const float4*>(Wst + 4 * g);
      fma_samples_bias<S>(acc[4 * g + 0], v, w.x, b.x);
      fma_samples_bias<S>(acc[4 * g + 1], v, w.y, b.y);
      fma_samples_bias<S>(acc[4 * g + 2], v, w.z, b.z);
      fma_samples_bias<S>(acc[4 * g + 3], v, w.w, b.w);
    }
  }
  if (in.first) {
    const int n = in.n;
    for (int k = 1; k < n; ++k) dense_row<NG, S>(in.th + k * in.CS, Wst + k * ld, slot0, NT, acc);
    for (int k = (n > 1 ? n : 1); k < K; ++k)
      dense_row<NG, S>(in.xs + (int)in.id[k - n] * in.CS, Wst + k * ld, slot0, NT, acc);
  } else {
    const float* src = in.hc + in.CS;
    const float* wr = Wst + ld;
#pragma unroll 2
    for (int k = 1; k < K; ++k) {
      dense_row<NG, S>(src, wr, slot0, NT, acc);
      src += in.CS;
      wr += ld;
    }
  }
  if (act == DFLOW_ACT_RELU) {
#pragma unroll
    for (int o = 0; o < NG * 4; ++o) {
#pragma unroll
      for (int s = 0; s < S; ++s) acc[o][s] = fmaxf(acc[o][s], 0.0f);
      st_samples<S>(outcol + o * CS + slot0 * S, acc[o]);
    }
  } else {
#pragma unroll
    for (int o = 0; o < NG * 4; ++o) st_samples<S>(outcol + o * CS + slot0 * S, acc[o]);
    if (act != DFLOW_ACT_IDENTITY) {
      // tanh / sigmoid: applied in a ROLLED loop over the stored column (keeps the transcendental code out of the
      // unrolled tile; an unrolled copy per element blew the kernel up to 270 KB of SASS and thrashed the I-cache)
#pragma unroll 1
      for (int o = 0; o < NG * 4; ++o)
#pragma unroll 1
        for (int s = 0; s < S; ++s) {
          float* p = outcol + o * CS + slot0 * S + s;
          *p = act_apply(act, *p);
        }
    }
  }
}

// Evaluate one conditioner (Flux.Chain of Dense, src/Layers.jl:33-50).  Hidden activations go to
// hc + j*hstride (hstride = 0: one reused column block; HP*CS: kept for the adjoint), the final output to outcol.
// The last Dense (padded width op in {4,8,16,32,64}) is produced in chunks of <= 16 outputs so that the
// accumulator tile never exceeds the hidden tile.
template <int HP, int S>
__device__ __forceinline__ void run_net(const DevNet& net, const float* __restrict__ wblk, InSel in, float* hc,
                                        int hstride, float* outcol, int CS, int slot0, int NT) {
  const int D = net.depth;
  for (int j = 0; j < D; ++j) {
    in.first = (j == 0);
    in.hc = hc + (j > 0 ? (j - 1) * hstride : 0);
    const int K = net.w[j];
    const float* Wst = wblk + net.s_w[j];
    const float* bst = wblk + net.s_b[j];
    const int act = net.act[j];
    if (j < D - 1) {
      dense_to_col<HP / 4, S>(in, K, Wst, HP, bst, act, hc + j * hstride, CS, slot0, NT);
    } else {
      const int op = net.op[j];
      if (op == 4) {
        dense_to_col<1, S>(in, K, Wst, 4, bst, act, outcol, CS, slot0, NT);
      } else if (op == 8) {
        dense_to_col<2, S>(in, K, Wst, 8, bst, act, outcol, CS, slot0, NT);
      } else {
        for (int o0 = 0; o0 < op; o0 += 16)
          dense_to_col<4, S>(in, K, Wst + o0, op, bst + o0, act, outcol + o0 * CS, CS, slot0, NT);
      }
    }
  }
}

// ---- register-resident activations (constant-bank kernels) ------------------------------------------------------
template <int HP, int S>
__device__ __forceinline__ void act_regs(float (&h)[HP][S], int act) {
  if (act == DFLOW_ACT_RELU) {
#pragma unroll
    for (int o = 0; o < HP; ++o)
#pragma unroll
      for (int s = 0; s < S; ++s) h[o][s] = fmaxf(h[o][s], 0.0f);
  }
  // tanh / sigmoid chains never reach the register-resident kernels (DevChainHdr::relu_only, launch_fwd)
}

// ---- constant-bank variant (weights never touch shared memory or vector registers) ----------------------------
// Chains whose descriptor + staged weight image fit the 64 KB constant bank (every hidden <= 32 README-class chain)
// run with the weights as *uniform-datapath* operands: the Dense loops are fully unrolled over register-resident
// activations, every weight is fetched by `LDCU` (uniform constant load, index provably warp-uniform) into a uniform
// register and enters `FFMA2 R, R.F32x2, UR.F32, R` as a broadcast operand.  Per hidden Dense: zero shared-memory
// wavefronts (the shared-memory-column kernel spends 12 per input row), zero weight registers.  Measured inner loop
// (scripts/ubench_ldcu.cu, B200): 109 FMA/clk/SM against 92 with broadcast LDS.128 weights.
// The bank image is [DevChain | pad to 16 B | staged weights] and is uploaded by the launcher right before the kernel.
#ifdef DFLOW_CBANK
constexpr int CBANK_BYTES = 60 * 1024;
__constant__ uint4 g_cbank[CBANK_BYTES / 16];
__device__ __forceinline__ float cbw(int off) { return reinterpret_cast<const float*>(g_cbank)[off]; }
// four consecutive weights (off is a multiple of 4 floats: staged blocks, rows and biases are 16-byte aligned) with one
// wide uniform load instead of four LDCU.32 (the 32-bit form made LDCU 24 % of all issued instructions)
// (offsets are passed in float4 units, base + compile-time immediate: with float offsets every load cost a uniform shift
// and add -- USHF + UIADD3 were 17 % of all issued instructions in the ncu capture)
__device__ __forceinline__ float4 cbw4(int off4) { return reinterpret_cast<const float4*>(g_cbank)[off4]; }
template <int S>
__device__ __forceinline__ void bias4_samples(float (&a0)[S], float (&a1)[S], float (&a2)[S], float (&a3)[S], int off) {
  const float4 b = cbw4(off);
  bias_samples<S>(a0, b.x);
  bias_samples<S>(a1, b.y);
  bias_samples<S>(a2, b.z);
  bias_samples<S>(a3, b.w);
}
template <int S>
__device__ __forceinline__ void fma4_samples(float (&a0)[S], float (&a1)[S], float (&a2)[S], float (&a3)[S],
                                             const float (&v)[S], int off) {
  const float4 w = cbw4(off);
  fma_samples<S>(a0, v, w.x);
  fma_samples<S>(a1, v, w.y);
  fma_samples<S>(a2, v, w.z);
  fma_samples<S>(a3, v, w.w);
}

template <int HP, int S>
__device__ __forceinline__ void run_net_const(const DevNet& net, int wofs, const InSel& in, float* outcol, int CS,
                                              int slot0) {
  const int D = net.depth;  // >= 2 (host-side eligibility)
  float h[HP][S];
  // first Dense: inputs are the gathered theta / x columns (runtime depth), outputs in registers; the bias is the
  // addend of the first row
  {
    const int w0 = (wofs + net.s_w[0]) >> 2, b0 = (wofs + net.s_b[0]) >> 2;  // float4 units (16-byte aligned blocks)
    const int n = in.n, K = net.w[0];
#pragma unroll
    for (int o = 0; o < HP; o += 4) bias4_samples<S>(h[o], h[o + 1], h[o + 2], h[o + 3], b0 + o / 4);
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
      const float* src = k < n ? in.th + k * CS : in.xs + (int)in.id[k - n] * CS;
      float v[S];
      ld_samples<S>(src + slot0 * S, v);
#pragma unroll
      for (int o = 0; o < HP; o += 4) fma4_samples<S>(h[o], h[o + 1], h[o + 2], h[o + 3], v, w0 + k * (HP / 4) + o / 4);
    }
  }
  act_regs<HP, S>(h, net.act[0]);
  // hidden Dense layers: registers -> registers (weight rows are zero-padded to HP)
#pragma unroll 1
  for (int j = 1; j < D - 1; ++j) {
    float o[HP][S];
    const int wj = (wofs + net.s_w[j]) >> 2, bj = (wofs + net.s_b[j]) >> 2;
#pragma unroll
    for (int oo = 0; oo < HP; oo += 4) bias4_samples<S>(o[oo], o[oo + 1], o[oo + 2], o[oo + 3], bj + oo / 4);
#pragma unroll
    for (int k = 0; k < HP; ++k)
#pragma unroll
      for (int oo = 0; oo < HP; oo += 4)
        fma4_samples<S>(o[oo], o[oo + 1], o[oo + 2], o[oo + 3], h[k], wj + k * (HP / 4) + oo / 4);
    act_regs<HP, S>(o, net.act[j]);
#pragma unroll
    for (int k = 0; k < HP; ++k)
#pragma unroll
      for (int s = 0; s < S; ++s) h[k][s] = o[k][s];
  }
  // last Dense (padded width op in {4, 8, 16, 32, 64}) in groups of 4 outputs straight from the registers; the row
  // stride is a compile-time constant for the common widths so that every weight address is base + immediate
  {
    const int jl = D - 1, op = net.op[jl], act = net.act[jl];
    const int wl = (wofs + net.s_w[jl]) >> 2, bl = (wofs + net.s_b[jl]) >> 2;
    auto group = [&](int o0, auto ldc) {
      constexpr int LD = decltype(ldc)::value;  // 0: runtime stride
      float acc[4][S];
      bias4_samples<S>(acc[0], acc[1], acc[2], acc[3], bl + (o0 >> 2));
#pragma unroll
      for (int k = 0; k < HP; ++k)
        fma4_samples<S>(acc[0], acc[1], acc[2], acc[3], h[k], wl + k * ((LD ? LD : op) >> 2) + (o0 >> 2));
#pragma unroll
      for (int oo = 0; oo < 4; ++oo) st_samples<S>(outcol + (o0 + oo) * CS + slot0 * S, acc[oo]);
    };
    if (op == 4) {
      group(0, std::integral_constant<int, 4>{});
    } else if (op == 8) {
      group(0, std::integral_constant<int, 8>{});
      group(4, std::integral_constant<int, 8>{});
    } else {
#pragma unroll 1
      for (int o0 = 0; o0 < op; o0 += 4) group(o0, std::integral_constant<int, 0>{});
    }
    if (act != DFLOW_ACT_IDENTITY) {
#pragma unroll 1
      for (int o = 0; o < op; ++o)
#pragma unroll 1
        for (int s = 0; s < S; ++s) {
          float* p = outcol + o * CS + slot0 * S + s;
          *p = act_apply(act, *p);
        }
    }
  }
}
#endif  // DFLOW_CBANK

// cooperative float4 copy global -> shared (len4 = number of float4)
__device__ __forceinline__ void copy_f4(float* dst, const float* __restrict__ src, int len4, int tid, int nt) {
  float4* d4 = reinterpret_cast<float4*>(dst);
  const float4* s4 = reinterpret_cast<const float4*>(src);
  for (int i = tid; i < len4; i += nt) d4[i] = __ldg(s4 + i);
}

// Apply one element in the normalising (`backward`, x -> z) or sampling direction to the S samples of this
// thread.  ldj is accumulated with the reference's signs.
template <int HP, int S, bool REG = false, bool CB = false>
__device__ __forceinline__ void elem_apply(const DevChainHdr& H, const DevElem& E, const float* __restrict__ wblk,
                                           bool sampling, float* xs, float* th, float* hc, int hstride, float* sb,
                                           float* tb, int CS, int slot0, int NT, float (&ldj)[S], int wofs = 0) {
#ifdef DFLOW_CBANK
#define wat(i) (CB ? cbw(wofs + (i)) : wblk[(i)])
#else
#define wat(i) (wblk[(i)])
#endif
  if (E.kind == DFLOW_ELEM_NORM) {
    // src/norm/Normalization.jl:64-103
    const int d = H.d;
    const float alpha = wat(2 * d), beta = wat(2 * d + 1), c = wat(2 * d + 2);
    for (int k = 0; k < d; ++k) {
      const float xmin = wat(k), xmax = wat(d + k);
#pragma unroll
      for (int s = 0; s < S; ++s) {
        float* p = xs + k * CS + slot0 * S + s;
        const float v = *p;
        if (!sampling)
          *p = (beta * (v - xmin) + alpha * (xmax - v)) / (xmax - xmin);
        else
          *p = ((xmax - xmin) * v - alpha * xmax + beta * xmin) / (beta - alpha);
      }
    }
#pragma unroll
    for (int s = 0; s < S; ++s) ldj[s] += sampling ? c : -c;
    return;
  }
#undef wat
  InSel in{th, xs, hc, E.id, H.n, CS, true};
  const bool rnvp = (E.kind == DFLOW_ELEM_RNVP);
  for (int ni = rnvp ? 0 : 1; ni < 2; ++ni) {
#ifdef DFLOW_CBANK
    if constexpr (CB) {
      run_net_const<HP, S>(ni == 0 ? E.s : E.t, wofs, in, ni == 0 ? sb : tb, CS, slot0);
      continue;
    }
#endif
    run_net<HP, S>(ni == 0 ? E.s : E.t, wblk, in, hc, hstride, ni == 0 ? sb : tb, CS, slot0, NT);
  }
  float lsum[S];
#pragma unroll
  for (int s = 0; s < S; ++s) lsum[s] = 0.0f;
  for (int j = 0; j < E.a; ++j) {
    const int k = E.af[j];
    float sv[S], tv[S], xv[S];
    if (rnvp) {
      ld_samples<S>(sb + j * CS + slot0 * S, sv);
    } else {
#pragma unroll
      for (int s = 0; s < S; ++s) sv[s] = 0.0f;
    }
    ld_samples<S>(tb + j * CS + slot0 * S, tv);
    ld_samples<S>(xs + k * CS + slot0 * S, xv);
#pragma unroll
    for (int s = 0; s < S; ++s) {
      if (!sampling)
        xv[s] = (xv[s] - tv[s]) * expf(-sv[s]);  // RNVP.jl:92
      else
        xv[s] = xv[s] * expf(sv[s]) + tv[s];  // RNVP.jl:184
      lsum[s] += sv[s];
    }
    st_samples<S>(xs + k * CS + slot0 * S, xv);
  }
#pragma unroll
  for (int s = 0; s < S; ++s) ldj[s] += sampling ? lsum[s] : -lsum[s];
}

// ------------------------------------------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller (spec: oracle/philox.py)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(unsigned int c0, unsigned int c1, unsigned int c2, unsigned int c3,
                                              unsigned int k0, unsigned int k1, unsigned int (&r)[4]) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const unsigned int n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  r[0] = c0;
  r[1] = c1;
  r[2] = c2;
  r[3] = c3;
}

__device__ __forceinline__ void box_muller(unsigned int a, unsigned int b, float& z0, float& z1) {
  const float u0 = ((float)(a >> 8) + 0.5f) * 5.9604644775390625e-08f;  // 2^-24
  const float u1 = ((float)(b >> 8) + 0.5f) * 5.9604644775390625e-08f;
  const float rad = sqrtf(-2.0f * logf(u0));
  float sn, cs;
  sincospif(2.0f * u1, &sn, &cs);
  z0 = rad * cs;
  z1 = rad * sn;
}

// ------------------------------------------------------------------------------------------------------------
// shared-memory carve-up (identical arithmetic on host and device)
// ------------------------------------------------------------------------------------------------------------
struct SmemPlan {
  int chain_f;  // floats reserved for the DevChain copy
  int w_f;      // floats for weights
  int cs;       // column stride (floats)
  int cols_f;   // floats for columns
  int grad_f;   // floats for the shared gradient accumulator (grad kernel)
  __host__ __device__ size_t bytes() const { return 4ull * ((size_t)chain_f + w_f + cols_f + grad_f); }
};

__host__ __device__ inline SmemPlan plan_fwd(const DevChainHdr& h, int chain_bytes, int nts, bool reg = false,
                                             bool cb = false) {
  SmemPlan p;
  p.chain_f = cb ? 0 : ((chain_bytes + 15) / 16) * 4;  // constant-bank variant: descriptor and weights stay out of shared
  p.w_f = cb ? 0 : (h.resident ? h.stage_total : h.stage_max);
  p.cs = nts;
  const int rows = h.d + h.n + (reg ? 0 : h.hp) + 2 * h.amax4;  // REG keeps hidden activations in registers
  p.cols_f = rows * p.cs;
  p.grad_f = 0;
  return p;
}

__host__ __device__ inline SmemPlan plan_grad(const DevChainHdr& h, int chain_bytes, int nt, int smem_grad) {
  SmemPlan p;
  p.chain_f = ((chain_bytes + 15) / 16) * 4;
  p.w_f = h.resident ? h.stage_total : h.stage_max;
  p.cs = nt + 4;
  const int hd = h.max_depth > 1 ? h.max_depth - 1 : 1;
  const int rows = 2 * h.d + h.n + hd * h.hp + h.hp + 4 * h.amax4;
  p.cols_f = rows * p.cs;
  p.grad_f = smem_grad ? ((h.P + 3) / 4) * 4 : 0;
  return p;
}

// ------------------------------------------------------------------------------------------------------------
// K1 / K2: fused chain, normalising or sampling direction
// ------------------------------------------------------------------------------------------------------------
// max threads per CTA of an instantiation (host clamps blockDim to this): big accumulator tiles get 128 threads
// and up to 255 registers, the others 256 threads x 2 CTAs (<= 128 registers)
template <int HP, int S, bool REG>
constexpr int fwd_max_threads() {
  return (REG || HP * S >= 64) ? 128 : 256;
}
template <int HP, int S, bool REG>
constexpr int fwd_min_ctas() {
  return (REG || HP * S > 64) ? 2 : (HP * S == 64 ? 3 : 2);  // 255 / 168 / 128 registers
}

// Tile input: global -> shared-memory columns (x, theta), or the in-kernel base draw.
template <int HP, int S>
__device__ __forceinline__ void fwd_load_tile(const FwdArgs& a, const DevChainHdr& H, float* xs, float* th, int CS, int tid,
                                              long long base, int NTS, bool io_aligned) {
  const int d = H.d, n = H.n;
    const bool vec_io = (S == 4) && a.idx == nullptr && (base + NTS <= a.B) && io_aligned;
    if (vec_io && a.mode != MODE_SAMPLE_RNG) {
      const float4* xp4 = reinterpret_cast<const float4*>(a.x_in + (base + (long long)tid * S) * d);
      int s = 0, k = 0;
      for (int c = 0; c < d; ++c) {
        const float4 t = __ldg(xp4 + c);
        const float e[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          xs[k * CS + tid * S + s] = e[q];
          if (++k == d) {
            k = 0;
            ++s;
          }
        }
      }
    }
    if (vec_io && a.theta != nullptr && n > 0) {
      const float4* tp4 = reinterpret_cast<const float4*>(a.theta + (base + (long long)tid * S) * n);
      int s = 0, k = 0;
      for (int c = 0; c < n; ++c) {
        const float4 t = __ldg(tp4 + c);
        const float e[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float v = e[q];
          if (a.flags & DFLOW_THETA_NORMALIZE) v = (H.theta_rng[k] == 0.0f) ? 0.0f : (v - H.theta_min[k]) / H.theta_rng[k];
          th[k * CS + tid * S + s] = v;
          if (++k == n) {
            k = 0;
            ++s;
          }
        }
      }
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int sl = tid * S + s;
      const long long gi = base + sl;
      const bool valid = gi < a.B;
      const long long src = (valid && a.idx) ? (long long)a.idx[gi] : gi;
      if (a.mode == MODE_SAMPLE_RNG) {
        const unsigned long long ctr = a.first_sample + (unsigned long long)gi;
        for (int g = 0; g < (d + 3) / 4; ++g) {
          unsigned int r[4];
          philox4x32_10((unsigned int)ctr, (unsigned int)(ctr >> 32), (unsigned int)g, a.rng_offset,
                        (unsigned int)a.seed, (unsigned int)(a.seed >> 32), r);
          float z[4];
          box_muller(r[0], r[1], z[0], z[1]);
          box_muller(r[2], r[3], z[2], z[3]);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (4 * g + q < d) xs[(4 * g + q) * CS + sl] = z[q];
        }
      } else if (!vec_io) {
        const float* xp = a.x_in + src * d;
        for (int k = 0; k < d; ++k) xs[k * CS + sl] = valid ? __ldg(xp + k) : 0.0f;
      }
      if (vec_io && a.theta != nullptr) continue;  // θ already loaded by the vector path
      for (int k = 0; k < n; ++k) {
        float v = 0.0f;
        if (a.theta_const)
          v = __ldg(a.theta_const + k);
        else if (valid && a.theta)
          v = __ldg(a.theta + src * n + k);
        if (a.flags & DFLOW_THETA_NORMALIZE) {
          // normalize_input, src/Data.jl:213-218
          v = (H.theta_rng[k] == 0.0f) ? 0.0f : (v - H.theta_min[k]) / H.theta_rng[k];
        }
        th[k * CS + sl] = v;
      }
    }
}

// Tile input of dflow_logpdf_grid: point gi of the tensor-product grid, first vector fastest (Iterators.product order,
// src/Flows.jl:301), one fixed condition.  A separate real call: the hot tile loader above stays exactly as profiled.
template <int S>
__device__ __noinline__ void fwd_load_grid_tile(const FwdArgs& a, const DevChainHdr& H, float* xs, float* th, int CS, int tid,
                                                long long base) {
  const int d = H.d, n = H.n;
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int sl = tid * S + s;
    const long long gi = base + sl;
    const bool valid = gi < a.B;
    for (int k = 0; k < d; ++k) {
      const long long len = a.grid_meta[3 * k], stride = a.grid_meta[3 * k + 1], off = a.grid_meta[3 * k + 2];
      xs[k * CS + sl] = valid ? __ldg(a.grid_vals + off + (gi / stride) % len) : 0.0f;
    }
    for (int k = 0; k < n; ++k) {
      float v = a.theta_const ? __ldg(a.theta_const + k) : 0.0f;
      if (a.flags & DFLOW_THETA_NORMALIZE) v = (H.theta_rng[k] == 0.0f) ? 0.0f : (v - H.theta_min[k]) / H.theta_rng[k];
      th[k * CS + sl] = v;
    }
  }
}

// Tile output: columns -> global (z / x, ldj or logp); returns this thread's (sum logp, #non-finite) contribution.
template <int HP, int S>
__device__ __forceinline__ float2 fwd_store_tile(const FwdArgs& a, const DevChainHdr& H, const float* xs, int CS, int tid,
                                                 long long base, int NTS, bool io_aligned, float4 ldj4) {
  const int d = H.d;
  const float ldj[4] = {ldj4.x, ldj4.y, ldj4.z, ldj4.w};
  float lsum_thread = 0.0f, nonfinite = 0.0f;
    // ---- store ----
    const bool vec_out = (S == 4) && (base + NTS <= a.B) && io_aligned;
    float aux[S];  // per-sample scalar result: logp or ldj
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int sl = tid * S + s;
      const long long gi = base + sl;
      aux[s] = ldj[s];
      if (a.mode == MODE_LOGPDF || a.mode == MODE_LOGPDF_SUM) {
        float q = 0.0f;
        for (int k = 0; k < d; ++k) {
          const float v = xs[k * CS + sl];
          q = fmaf(v, v, q);
        }
        const float lp = H.logpdf_c0 - 0.5f * q + ldj[s];  // src/Flows.jl:279
        aux[s] = lp;
        if (a.mode == MODE_LOGPDF_SUM && gi < a.B) {
          if (isfinite(lp))
            lsum_thread += lp;
          else
            nonfinite += 1.0f;
        }
      } else if (!vec_out && gi < a.B) {
        float* op = a.x_out + gi * d;
        for (int k = 0; k < d; ++k) op[k] = xs[k * CS + sl];
      }
      if (!vec_out && gi < a.B && a.mode != MODE_LOGPDF_SUM && a.mode != MODE_SAMPLE && a.mode != MODE_SAMPLE_RNG)
        a.aux_out[gi] = aux[s];
    }
    if (vec_out) {
      if (a.mode != MODE_LOGPDF && a.mode != MODE_LOGPDF_SUM) {
        // 4 consecutive samples = d float4 of contiguous output
        float4* op4 = reinterpret_cast<float4*>(a.x_out + (base + (long long)tid * S) * d);
        int s = 0, k = 0;
        for (int c = 0; c < d; ++c) {
          float e[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            e[q] = xs[k * CS + tid * S + s];
            if (++k == d) {
              k = 0;
              ++s;
            }
          }
          op4[c] = make_float4(e[0], e[1], e[2], e[3]);
        }
      }
      if (a.mode != MODE_LOGPDF_SUM && a.mode != MODE_SAMPLE && a.mode != MODE_SAMPLE_RNG)
        *reinterpret_cast<float4*>(a.aux_out + base + (long long)tid * S) = make_float4(aux[0], aux[1 % S], aux[2 % S], aux[3 % S]);
    }
  return make_float2(lsum_thread, nonfinite);
}

// Real calls for the constant-bank kernel: with the tile I/O inlined next to the chain, ptxas demotes the chain's
// weight loads from the uniform datapath (LDCU -> FFMA2 UR operand) to per-thread LDC, which is 4x too slow.
template <int HP, int S>
__device__ __noinline__ void fwd_load_tile_call(const FwdArgs& a, const DevChainHdr& H, float* xs, float* th, int CS, int tid,
                                                long long base, int NTS, bool io_aligned) {
  fwd_load_tile<HP, S>(a, H, xs, th, CS, tid, base, NTS, io_aligned);
}
template <int HP, int S>
__device__ __noinline__ float2 fwd_store_tile_call(const FwdArgs& a, const DevChainHdr& H, const float* xs, int CS, int tid,
                                                   long long base, int NTS, bool io_aligned, float4 ldj4) {
  return fwd_store_tile<HP, S>(a, H, xs, CS, tid, base, NTS, io_aligned, ldj4);
}

// FIXED: blockDim.x equals fwd_max_threads, so the column stride CS = NT*S is a compile-time constant and every
// column address `unit*CS + slot` of an unrolled loop folds into an immediate offset (the runtime-stride build spent
// ~15 % of its instructions on LEA/IMAD/IADD3 address arithmetic; profiles/r01_ncu_fwd_summary.md).
template <int HP, int S, bool REG, bool FIXED, bool CB = false>
__device__ __forceinline__ void chain_fwd_body(const FwdArgs& a) {
  extern __shared__ float4 smem4[];
  float* smem = reinterpret_cast<float*>(smem4);
  const int tid = threadIdx.x, NT = FIXED ? fwd_max_threads<HP, S, REG>() : (int)blockDim.x, NTS = NT * S;

  const DevChain* C;
#ifdef DFLOW_CBANK
  if constexpr (CB) {
    C = reinterpret_cast<const DevChain*>(g_cbank);  // descriptor and weights live in the constant bank
  } else
#endif
  {
    // DevChain -> shared
    copy_f4(smem, reinterpret_cast<const float*>(a.chain), (a.chain_bytes + 15) / 16, tid, NT);
    __syncthreads();
    C = reinterpret_cast<const DevChain*>(smem);
  }
  const DevChainHdr& H = C->h;
  const SmemPlan P = plan_fwd(H, a.chain_bytes, NTS, REG, CB);
  float* wsm = smem + P.chain_f;
  float* cols = wsm + P.w_f;
  const int CS = FIXED ? NTS : P.cs;
  float* xs = cols;
  float* th = xs + H.d * CS;
  float* hc = th + H.n * CS;
  float* sb = hc + (REG ? 0 : H.hp) * CS;
  float* tb = sb + H.amax4 * CS;

  if (!CB && H.resident) {
    copy_f4(wsm, a.staged, H.stage_total / 4, tid, NT);
    __syncthreads();
  }

  const int d = H.d, n = H.n, L = H.L;
  const bool sampling = (a.mode >= MODE_SAMPLE);
  // 128-bit global access needs 16-byte aligned array bases (cudaMalloc / CuArray / torch give >= 256)
  const bool io_aligned = ((reinterpret_cast<uintptr_t>(a.x_in) | reinterpret_cast<uintptr_t>(a.theta) |
                            reinterpret_cast<uintptr_t>(a.x_out) | reinterpret_cast<uintptr_t>(a.aux_out)) & 15) == 0;
  const long long ntiles = (a.B + NTS - 1) / NTS;
  float lsum_thread = 0.0f, nonfinite = 0.0f;

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long base = tile * NTS;
    // ---- load ----
    // fast path (S == 4, full tile, no gather): the thread's 4 consecutive samples are 4*d contiguous floats =
    // d aligned float4 -> coalesced 128-bit global loads, transposed into the columns on the fly
    if (a.grid_vals)
      fwd_load_grid_tile<S>(a, H, xs, th, CS, tid, base);
    else if constexpr (CB)
      fwd_load_tile_call<HP, S>(a, H, xs, th, CS, tid, base, NTS, io_aligned);
    else
      fwd_load_tile<HP, S>(a, H, xs, th, CS, tid, base, NTS, io_aligned);
    float ldj[S];
#pragma unroll
    for (int s = 0; s < S; ++s) ldj[s] = 0.0f;

    // ---- chain ----
    for (int step = 0; step < L; ++step) {
      const int ei = sampling ? step : (L - 1 - step);  // src/Chains.jl:155-161 vs :174-180
      const DevElem& E = C->e[ei];
      if constexpr (CB) {
        elem_apply<HP, S, REG, true>(H, E, nullptr, sampling, xs, th, hc, 0, sb, tb, CS, tid, NT, ldj,
                                     a.cb_wofs + E.stage_off);
        continue;
      }
      const float* wblk;
      if (H.resident) {
        wblk = wsm + E.stage_off;
      } else {
        __syncthreads();
        copy_f4(wsm, a.staged + E.stage_off, E.stage_len / 4, tid, NT);
        __syncthreads();
        wblk = wsm;
      }
      elem_apply<HP, S, REG>(H, E, wblk, sampling, xs, th, hc, 0, sb, tb, CS, tid, NT, ldj);
    }

    // ---- store ----
    {
      float2 acc2;
      if constexpr (CB)
        acc2 = fwd_store_tile_call<HP, S>(a, H, xs, CS, tid, base, NTS, io_aligned, make_float4(ldj[0], ldj[1 % S], ldj[2 % S], ldj[3 % S]));
      else
        acc2 = fwd_store_tile<HP, S>(a, H, xs, CS, tid, base, NTS, io_aligned, make_float4(ldj[0], ldj[1 % S], ldj[2 % S], ldj[3 % S]));
      lsum_thread += acc2.x;
      nonfinite += acc2.y;
    }
  }

  if (a.mode == MODE_LOGPDF_SUM) {
    // block reduce -> one atomic per CTA
    __syncthreads();
    float v0 = lsum_thread, v1 = nonfinite;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v0 += __shfl_xor_sync(0xffffffffu, v0, o);
      v1 += __shfl_xor_sync(0xffffffffu, v1, o);
    }
    float* red = cols;  // reuse
    if ((tid & 31) == 0) {
      red[(tid >> 5) * 2] = v0;
      red[(tid >> 5) * 2 + 1] = v1;
    }
    __syncthreads();
    if (tid == 0) {
      float t0 = 0.0f, t1 = 0.0f;
      for (int w = 0; w < (NT + 31) / 32; ++w) {
        t0 += red[2 * w];
        t1 += red[2 * w + 1];
      }
      atomicAdd(a.aux_out, t0);
      if (t1 != 0.0f) atomicAdd(a.aux_out + 1, t1);
    }
  }
}

template <int HP, int S, bool REG>
__global__ void __launch_bounds__((fwd_max_threads<HP, S, REG>()), (fwd_min_ctas<HP, S, REG>()))
    chain_fwd_kernel(const FwdArgs a) {
  if (blockDim.x == fwd_max_threads<HP, S, REG>())
    chain_fwd_body<HP, S, REG, true>(a);
  else
    chain_fwd_body<HP, S, REG, false>(a);
}

#ifdef DFLOW_CBANK
// 3 CTAs of 128 threads per SM: 142 registers, no spills.  A 128-register build (4 CTAs, 96 B of spills) measured slower
// (7.6-7.7e9 vs 8.1e9 samples/s on C2).
#ifndef DFLOW_CB_MINCTAS
#define DFLOW_CB_MINCTAS 3
#endif
template <int HP, int S>
__global__ void __launch_bounds__((fwd_max_threads<HP, S, true>()), DFLOW_CB_MINCTAS)
    chain_fwd_const_kernel(const __grid_constant__ FwdArgs a) {
  chain_fwd_body<HP, S, true, true, true>(a);  // always launched with fwd_max_threads (compile-time column stride)
}
#endif

template <int HP, int S, bool REG>
cudaError_t launch_fwd_inst(const FwdArgs& a, unsigned grid, int nt, size_t smem, cudaStream_t st);
// constant-bank forward kernel: uploads [DevChain | staged weights] into the instantiation's bank, then launches
template <int HP, int S>
cudaError_t launch_fwd_const_inst(const FwdArgs& a, unsigned grid, int nt, size_t smem, cudaStream_t st, int stage_floats);

}  // namespace dflow
