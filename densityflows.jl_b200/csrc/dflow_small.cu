// Small-minibatch training: a whole epoch of (gather, loss + gradient, Adam) steps inside ONE persistent CTA.
//
// The reference's default batchsize is 64 (src/Flows.jl:380): a train step is then 1e6 flops, and what it costs on a GPU
// is launch latency and the dependent-instruction latency of one thread walking a whole sample through the chain (the
// one-thread-per-sample adjoint kernel needs 85 us for one 64-sample tile, 96 us per step with the Adam launch).  Here
//   * one launch covers every minibatch of the epoch (src/Flows.jl:394-416: `for (x, θ) in loader` ... update!);
//   * the parameters, both Adam moments and the gradient live in shared memory for the whole epoch (C1: 4 x 2.6 K floats);
//   * a sample is spread over LPS lanes (LPS = 16 or 32 >= widest Dense output): lane u evaluates unit u of every Dense
//     (y_u = b_u + sum_k W[u,k] a_k, a_k broadcast from the sample's shared-memory tape), so the serial depth of a
//     conditioner is ~3 x K fmas instead of ~400, and 1024 / LPS samples are in flight per pass;
//   * nothing is recomputed: every activation of the minibatch stays on the tape (64 samples x ~250 floats);
//   * weight gradients are reduced in a fixed order (4 lanes per entry, 16 samples each, two shuffles, one owner) --
//     deterministic, no atomics; Adam (same explicit round-to-nearest sequence as adam_kernel) runs on the CTA.
// W is stored with an odd row pitch so that both W a (lanes over outputs) and W^T delta (lanes over inputs) are
// bank-conflict free.
//
// Reference math: src/affine/RNVP.jl:77-96,99-147; src/affine/NICE.jl; src/norm/Normalization.jl:64-103;
// src/Flows.jl:352-359,398-415; Optimisers.Adam.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "dflow_chain_kernels.cuh"
#include "dflow_small.h"

namespace dflow {

constexpr int SM_THREADS = 1024;
constexpr int SM_MAX_DENSE_PER_LAYER = 2 * MAX_DENSE;

struct SmNet {
  int depth, has_bias;
  int w[MAX_DENSE + 1];
  int act[MAX_DENSE];
  int pw[MAX_DENSE], pb[MAX_DENSE];   // packed offsets (global W / m / v)
  int sw[MAX_DENSE], sb[MAX_DENSE];   // offsets in the padded shared-memory parameter image: W[o + op k]; bias
  int op[MAX_DENSE];                  // row pitch of the image (multiple of 4, op / 4 odd): row k = the O outputs of input k
  int tw[MAX_DENSE], kp[MAX_DENSE];   // transposed copy for the forward pass: WT[k + kp u], row u = the K inputs of output u
  int ta[MAX_DENSE + 1];              // tape offsets: ta[0] = conditioner input (nin), ta[j] = output of Dense j-1
  int dl[MAX_DENSE];                  // delta offsets (per-sample delta row): dl[j] = cotangent of Dense j's pre-activation
};

struct SmElem {
  int kind, a, nid, nin;
  int t_ck;     // tape offset of the checkpoint: u_af (coupling, a floats) or the whole state (norm, d floats; -1: none)
  int t_em;     // tape offset of exp(-s) (RNVP)
  int nstage;   // norm: offset of [x_min | x_max | alpha beta c] in the constants block
  unsigned char af[DMAX], id[DMAX];
  SmNet s, t;
  // weight-gradient phase: entries of every Dense of the element
  int n_dense;
  int dw_start[SM_MAX_DENSE_PER_LAYER + 1];  // cumulative entry counts (weights + bias)
  int dw_O[SM_MAX_DENSE_PER_LAYER], dw_K[SM_MAX_DENSE_PER_LAYER], dw_op[SM_MAX_DENSE_PER_LAYER];
  int dw_in[SM_MAX_DENSE_PER_LAYER], dw_dl[SM_MAX_DENSE_PER_LAYER], dw_g[SM_MAX_DENSE_PER_LAYER],
      dw_gb[SM_MAX_DENSE_PER_LAYER];
};

struct SmallPlanDev {
  int d, n, L, lps;
  int P;          // packed parameter count
  int PS;         // padded parameter image size (floats, multiple of 4)
  int PT;         // size of the transposed weight copy (floats)
  int TP;         // tape pitch per sample (= 8 mod 32: float4 rows of neighbouring samples / sample quarters hit distinct banks)
  int DP;         // delta-row pitch per sample (= 8 mod 32)
  int t_x, t_th, t_gx;  // tape offsets of the state, the conditions, the state cotangent
  int nconst;     // floats of normalisation constants
  int has_theta_range;
  float logpdf_c0;
  float theta_min[NMAX], theta_rng[NMAX];
  SmElem e[1];  // L entries
};

struct SmallArgs {
  const SmallPlanDev* plan;
  int plan_bytes;
  const float* consts;  // normalisation constants (device)
  float* W;
  float* m;
  float* v;
  const float* x;
  const float* theta;
  const int32_t* order;
  long long n, batchsize;
  float lr, b1, b2, eps, b1t, b2t;  // running products beta^t BEFORE the first step of this launch
  int flags;
  float* loss2_out;
};

__device__ __forceinline__ float sm_act(int code, float v) {
  if (code == DFLOW_ACT_RELU) return fmaxf(v, 0.0f);
  if (code == DFLOW_ACT_IDENTITY) return v;
  return act_apply(code, v);
}
__device__ __forceinline__ float sm_act_grad(int code, float y) {
  if (code == DFLOW_ACT_RELU) return y > 0.0f ? 1.0f : 0.0f;
  return act_grad(code, y);
}

// y_u = act(b_u + sum_k W[u, k] a_k): lane u < O walks row u of the transposed copy and the sample's input vector with
// 128-bit shared-memory loads (rows and vectors are zero-padded to a multiple of 4)
__device__ __forceinline__ void sm_dense_fwd(const float* __restrict__ Wsm, const float* __restrict__ WT, const SmNet& net,
                                             int j, const float* in, float* out, int u) {
  const int K4 = (net.w[j] + 3) >> 2, O = net.w[j + 1];
  if (u < O) {
    const float4* wr = reinterpret_cast<const float4*>(WT + net.tw[j] + net.kp[j] * u);
    const float4* in4 = reinterpret_cast<const float4*>(in);
    float acc0 = net.has_bias ? Wsm[net.sb[j] + u] : 0.0f, acc1 = 0.0f, acc2 = 0.0f, acc3 = 0.0f;
#pragma unroll 4
    for (int k = 0; k < K4; ++k) {
      const float4 w = wr[k], x = in4[k];
      acc0 = fmaf(w.x, x.x, acc0);
      acc1 = fmaf(w.y, x.y, acc1);
      acc2 = fmaf(w.z, x.z, acc2);
      acc3 = fmaf(w.w, x.w, acc3);
    }
    out[u] = sm_act(net.act[j], (acc0 + acc1) + (acc2 + acc3));
  }
  __syncwarp();
}

template <int LPS>
__global__ void __launch_bounds__(SM_THREADS, 1) epoch_small_kernel(const SmallArgs a) {
  extern __shared__ float4 smem4[];
  float* smem = reinterpret_cast<float*>(smem4);
  const int tid = threadIdx.x;
  constexpr int NG = SM_THREADS / LPS;  // samples per pass
  const int grp = tid / LPS, u = tid % LPS;

  // ---- shared-memory carve-up: plan | consts | W | m | v | g | tape | delta rows | reduction scratch ----
  {
    const float4* src = reinterpret_cast<const float4*>(a.plan);
    float4* dst = smem4;
    for (int i = tid; i < (a.plan_bytes + 15) / 16; i += SM_THREADS) dst[i] = src[i];
  }
  __syncthreads();
  const SmallPlanDev& Pn = *reinterpret_cast<const SmallPlanDev*>(smem);
  const int d = Pn.d, n = Pn.n, L = Pn.L, PS = Pn.PS, TP = Pn.TP, DP = Pn.DP;
  float* cst = smem + ((a.plan_bytes + 15) / 16) * 4;
  float* Wsm = cst + ((Pn.nconst + 3) & ~3);
  float* Msm = Wsm + PS;
  float* Vsm = Msm + PS;
  float* Gsm = Vsm + PS;
  float* WT = Gsm + PS;  // transposed copy of the weights (forward pass), refreshed after every Adam update
  int* tmap = reinterpret_cast<int*>(WT + Pn.PT);  // image index -> index in the transposed copy (-1: bias / padding)
  float* tape = reinterpret_cast<float*>(tmap + PS);
  float* drow = tape + NG * TP;
  float* red = drow + NG * DP;  // [NG] logp values + 2 accumulators

  for (int i = tid; i < Pn.nconst; i += SM_THREADS) cst[i] = a.consts[i];
  // everything zero: image padding, tape / delta-row padding slots (read by the 128-bit loops, never written)
  for (int i = tid; i < 4 * PS + Pn.PT; i += SM_THREADS) Wsm[i] = 0.0f;
  for (int i = tid; i < NG * (TP + DP); i += SM_THREADS) tape[i] = 0.0f;
  for (int i = tid; i < PS; i += SM_THREADS) tmap[i] = -1;
  __syncthreads();
  // packed -> padded image of W, m, v
  for (int ei = 0; ei < L; ++ei) {
    const SmElem& E = Pn.e[ei];
    if (E.kind == DFLOW_ELEM_NORM) continue;
    for (int ni = (E.kind == DFLOW_ELEM_RNVP ? 0 : 1); ni < 2; ++ni) {
      const SmNet& net = ni == 0 ? E.s : E.t;
      for (int j = 0; j < net.depth; ++j) {
        const int K = net.w[j], O = net.w[j + 1], op = net.op[j];
        for (int i = tid; i < O * K; i += SM_THREADS) {
          const int k = i / O, o = i - k * O;
          const int si = net.sw[j] + o + op * k, gi = net.pw[j] + i;
          const float wv = a.W[gi];
          Wsm[si] = wv;
          WT[net.tw[j] + k + net.kp[j] * o] = wv;
          tmap[si] = net.tw[j] + k + net.kp[j] * o;
          Msm[si] = a.m[gi];
          Vsm[si] = a.v[gi];
        }
        if (net.has_bias)
          for (int o = tid; o < O; o += SM_THREADS) {
            Wsm[net.sb[j] + o] = a.W[net.pb[j] + o];
            Msm[net.sb[j] + o] = a.m[net.pb[j] + o];
            Vsm[net.sb[j] + o] = a.v[net.pb[j] + o];
          }
      }
    }
  }
  if (tid == 0) red[NG] = red[NG + 1] = 0.0f;
  __syncthreads();

  float* T = tape + grp * TP;   // this sample's tape
  float* Dr = drow + grp * DP;  // this sample's delta row
  float* xs = T + Pn.t_x;
  float* th = T + Pn.t_th;
  float* gx = T + Pn.t_gx;
  float b1t = a.b1t, b2t = a.b2t;

  for (long long b0 = 0; b0 < a.n; b0 += a.batchsize) {
    const long long nb = min(a.batchsize, a.n - b0);
    const float ib = (float)(1.0 / (double)nb);  // seed 1 / |minibatch| (mean over its true size, src/Flows.jl:358)
    for (long long p0 = 0; p0 < nb; p0 += NG) {
      const bool valid = p0 + grp < nb;
      // ---- gather the sample (src/Flows.jl:394 DataLoader batch through the index) ----
      const long long col = valid ? (long long)a.order[b0 + p0 + grp] : 0;
      for (int k = u; k < d; k += LPS) xs[k] = valid ? __ldg(a.x + col * d + k) : 0.0f;
      for (int k = u; k < n; k += LPS) {
        float v = valid ? __ldg(a.theta + col * n + k) : 0.0f;
        if (a.flags & DFLOW_THETA_NORMALIZE) v = (Pn.theta_rng[k] == 0.0f) ? 0.0f : (v - Pn.theta_min[k]) / Pn.theta_rng[k];
        th[k] = v;
      }
      __syncwarp();
      // ---- forward (normalising) sweep: last element first; everything stays on the tape ----
      float ldj = 0.0f;  // meaningful in lane 0 of the group
      for (int ei = L - 1; ei >= 0; --ei) {
        const SmElem& E = Pn.e[ei];
        if (E.kind == DFLOW_ELEM_NORM) {
          const float* cb = cst + E.nstage;
          const float alpha = cb[2 * d], beta = cb[2 * d + 1];
          for (int k = u; k < d; k += LPS) {
            const float xv = xs[k];
            if (E.t_ck >= 0) T[E.t_ck + k] = xv;
            xs[k] = (beta * (xv - cb[k]) + alpha * (cb[d + k] - xv)) / (cb[d + k] - cb[k]);
          }
          ldj -= cb[2 * d + 2];
          __syncwarp();
          continue;
        }
        const bool rnvp = E.kind == DFLOW_ELEM_RNVP;
        // conditioner input [theta ; x[axis_id]] (src/affine/RNVP.jl:157), shared by both nets
        float* in0 = T + E.t.ta[0];
        for (int k = u; k < E.nin; k += LPS) in0[k] = k < n ? th[k] : xs[E.id[k - n]];
        __syncwarp();
        for (int ni = rnvp ? 0 : 1; ni < 2; ++ni) {
          const SmNet& net = ni == 0 ? E.s : E.t;
          for (int j = 0; j < net.depth; ++j) sm_dense_fwd(Wsm, WT, net, j, T + net.ta[j], T + net.ta[j + 1], u);
        }
        const float* sv = T + E.s.ta[E.s.depth];
        const float* tv = T + E.t.ta[E.t.depth];
        if (u < E.a) {
          const int k = E.af[u];
          const float uv = xs[k];
          T[E.t_ck + u] = uv;
          const float em = rnvp ? expf(-sv[u]) : 1.0f;
          if (rnvp) T[E.t_em + u] = em;
          xs[k] = (uv - tv[u]) * em;  // RNVP.jl:92
        }
        __syncwarp();
        if (rnvp) {
          float ssum = 0.0f;
          for (int j = 0; j < E.a; ++j) ssum += sv[j];
          ldj -= ssum;
        }
      }
      // ---- loss and seeds: zbar = z / B, jbar = -1 / B ----
      {
        float q = 0.0f;
        for (int k = 0; k < d; ++k) q = fmaf(xs[k], xs[k], q);
        const float lp = Pn.logpdf_c0 - 0.5f * q + ldj;
        if (u == 0) red[grp] = valid ? lp : 0.0f;
        __syncwarp();
        for (int k = u; k < d; k += LPS) gx[k] = valid ? xs[k] * ib : 0.0f;
        __syncwarp();
      }
      const float ibv = valid ? ib : 0.0f;
      // ---- reverse sweep in chain order ----
      for (int ei = 0; ei < L; ++ei) {
        const SmElem& E = Pn.e[ei];
        if (E.kind == DFLOW_ELEM_NORM) {
          const float* cb = cst + E.nstage;
          const float alpha = cb[2 * d], beta = cb[2 * d + 1];
          for (int k = u; k < d; k += LPS) {
            gx[k] *= (beta - alpha) / (cb[d + k] - cb[k]);
            if (E.t_ck >= 0) xs[k] = T[E.t_ck + k];
          }
          __syncwarp();
          continue;
        }
        const bool rnvp = E.kind == DFLOW_ELEM_RNVP;
        // output cotangents of the conditioners (RNVP.jl:134-135): sbar = -zbar_af z_af - jbar, tbar = -zbar_af exp(-s)
        if (u < E.a) {
          const int k = E.af[u];
          const float zb = gx[k];
          if (rnvp) {
            Dr[E.s.dl[E.s.depth - 1] + u] = -zb * xs[k] + ibv;
            Dr[E.t.dl[E.t.depth - 1] + u] = -zb * T[E.t_em + u];
          } else {
            Dr[E.t.dl[E.t.depth - 1] + u] = -zb;
          }
        }
        __syncwarp();
        for (int ni = rnvp ? 0 : 1; ni < 2; ++ni) {
          const SmNet& net = ni == 0 ? E.s : E.t;
          const int D = net.depth;
          // the last Dense's activation (identity by default) acts on the net output
          if (net.act[D - 1] != DFLOW_ACT_IDENTITY) {
            if (u < net.w[D]) Dr[net.dl[D - 1] + u] *= act_grad(net.act[D - 1], T[net.ta[D] + u]);
            __syncwarp();
          }
          for (int j = D - 1; j >= 0; --j) {
            const int K = net.w[j], O4 = (net.w[j + 1] + 3) >> 2, op = net.op[j];
            const float4* dl4 = reinterpret_cast<const float4*>(Dr + net.dl[j]);
            // g_in[k] = sum_o W[o + op k] delta_j[o]: lane k walks row k of the image (128-bit loads, zero padded)
            for (int k = u; k < K; k += LPS) {
              const float4* wr = reinterpret_cast<const float4*>(Wsm + net.sw[j] + op * k);
              float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f, acc3 = 0.0f;
#pragma unroll 4
              for (int o = 0; o < O4; ++o) {
                const float4 w = wr[o], dv = dl4[o];
                acc0 = fmaf(w.x, dv.x, acc0);
                acc1 = fmaf(w.y, dv.y, acc1);
                acc2 = fmaf(w.z, dv.z, acc2);
                acc3 = fmaf(w.w, dv.w, acc3);
              }
              const float g = (acc0 + acc1) + (acc2 + acc3);
              if (j > 0) {
                Dr[net.dl[j - 1] + k] = g * sm_act_grad(net.act[j - 1], T[net.ta[j] + k]);
              } else if (k >= n) {
                gx[E.id[k - n]] += g;  // identity coordinates; the theta rows are dropped
              }
            }
            __syncwarp();
          }
        }
        // cotangent of the transformed coordinates and restore the layer input (RNVP.jl:137-139)
        if (u < E.a) {
          const int k = E.af[u];
          if (rnvp) gx[k] *= T[E.t_em + u];
          xs[k] = T[E.t_ck + u];
        }
        __syncthreads();
        // ---- weight gradients of this element.  A task = (output o, four consecutive inputs k) of one Dense, or one bias
        // entry; 8 tasks per warp iteration, each summed by 4 lanes over a quarter of the samples (lane = task + 8 quarter):
        // per sample one scalar delta load and one 128-bit activation load feed four fmas; the quarters meet in two shuffles
        // and one lane owns the entry -- fixed order, no atomics.
        {
          const int total = E.dw_start[E.n_dense];
          const int lane = tid & 31, warp = tid >> 5, q = lane >> 3;
          for (int base = 0; base < total; base += 8 * (SM_THREADS / 32)) {  // warp-uniform trip count
            const int e_raw = base + warp * 8 + (lane & 7);
            const bool live = e_raw < total;
            const int e = live ? e_raw : 0;
            int di = 0;
            while (e >= E.dw_start[di + 1]) ++di;
            const int r = e - E.dw_start[di];
            const int O = E.dw_O[di], K = E.dw_K[di], KQ = (K + 3) >> 2;
            float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
            const bool is_w = r < O * KQ;
            const int kq = is_w ? r / O : 0, o = is_w ? r - kq * O : r - O * KQ;
            const float* dp = drow + E.dw_dl[di] + o + q * DP;
            if (is_w) {
              const float* ap = tape + E.dw_in[di] + 4 * kq + q * TP;
#pragma unroll
              for (int i = 0; i < NG / 4; ++i) {
                const float dv = dp[4 * i * DP];
                const float4 av = *reinterpret_cast<const float4*>(ap + 4 * i * TP);
                a0 = fmaf(dv, av.x, a0);
                a1 = fmaf(dv, av.y, a1);
                a2 = fmaf(dv, av.z, a2);
                a3 = fmaf(dv, av.w, a3);
              }
            } else {
#pragma unroll
              for (int i = 0; i < NG / 4; ++i) a0 += dp[4 * i * DP];
            }
            a0 += __shfl_xor_sync(0xffffffffu, a0, 8);
            a1 += __shfl_xor_sync(0xffffffffu, a1, 8);
            a2 += __shfl_xor_sync(0xffffffffu, a2, 8);
            a3 += __shfl_xor_sync(0xffffffffu, a3, 8);
            a0 += __shfl_xor_sync(0xffffffffu, a0, 16);
            a1 += __shfl_xor_sync(0xffffffffu, a1, 16);
            a2 += __shfl_xor_sync(0xffffffffu, a2, 16);
            a3 += __shfl_xor_sync(0xffffffffu, a3, 16);
            if (live && q == 0) {
              if (is_w) {
                float* g = Gsm + E.dw_g[di] + o + E.dw_op[di] * 4 * kq;
                const int op = E.dw_op[di];
                g[0] += a0;
                if (4 * kq + 1 < K) g[op] += a1;
                if (4 * kq + 2 < K) g[2 * op] += a2;
                if (4 * kq + 3 < K) g[3 * op] += a3;
              } else {
                Gsm[E.dw_gb[di] + o] += a0;
              }
            }
          }
        }
        __syncthreads();
      }
      // loss of this pass (fixed order: deterministic)
      if (tid == 0) {
        float s = 0.0f, bad = 0.0f;
        for (int g = 0; g < NG && p0 + g < nb; ++g) {
          if (isfinite(red[g])) s += red[g];
          else bad += 1.0f;
        }
        red[NG] += s;
        red[NG + 1] += bad;
      }
      __syncthreads();
    }
    // ---- Optimisers.Adam on the padded images (padding entries stay zero) ----
    b1t *= a.b1;
    b2t *= a.b2;
    {
      const float c1 = 1.0f - b1t, c2 = 1.0f - b2t;
      for (int i = tid; i < PS; i += SM_THREADS) {
        const float gi = Gsm[i];
        const float mi = __fadd_rn(__fmul_rn(a.b1, Msm[i]), __fmul_rn(1.0f - a.b1, gi));
        const float vi = __fadd_rn(__fmul_rn(a.b2, Vsm[i]), __fmul_rn(1.0f - a.b2, __fmul_rn(gi, gi)));
        Msm[i] = mi;
        Vsm[i] = vi;
        const float den = __fadd_rn(__fsqrt_rn(__fdiv_rn(vi, c2)), a.eps);
        const float wn = __fsub_rn(Wsm[i], __fmul_rn(__fdiv_rn(__fdiv_rn(mi, c1), den), a.lr));
        Wsm[i] = wn;
        Gsm[i] = 0.0f;  // consumed: the next minibatch accumulates from zero
        const int ti = tmap[i];
        if (ti >= 0) WT[ti] = wn;  // transposed copy for the next step's forward pass
      }
    }
    __syncthreads();
  }

  // ---- padded image -> packed buffers ----
  for (int ei = 0; ei < L; ++ei) {
    const SmElem& E = Pn.e[ei];
    if (E.kind == DFLOW_ELEM_NORM) continue;
    for (int ni = (E.kind == DFLOW_ELEM_RNVP ? 0 : 1); ni < 2; ++ni) {
      const SmNet& net = ni == 0 ? E.s : E.t;
      for (int j = 0; j < net.depth; ++j) {
        const int K = net.w[j], O = net.w[j + 1], op = net.op[j];
        for (int i = tid; i < O * K; i += SM_THREADS) {
          const int k = i / O, o = i - k * O;
          const int si = net.sw[j] + o + op * k, gi = net.pw[j] + i;
          a.W[gi] = Wsm[si];
          a.m[gi] = Msm[si];
          a.v[gi] = Vsm[si];
        }
        if (net.has_bias)
          for (int o = tid; o < O; o += SM_THREADS) {
            a.W[net.pb[j] + o] = Wsm[net.sb[j] + o];
            a.m[net.pb[j] + o] = Msm[net.sb[j] + o];
            a.v[net.pb[j] + o] = Vsm[net.sb[j] + o];
          }
      }
    }
  }
  if (tid == 0 && a.loss2_out) {
    a.loss2_out[0] += red[NG];
    a.loss2_out[1] += red[NG + 1];
  }
}

// ---- host side -------------------------------------------------------------------------------------------------
struct SmallPlan {
  std::vector<unsigned char> host;
  SmallPlanDev* d_plan = nullptr;
  float* d_consts = nullptr;
  int lps = 0;
  size_t smem = 0;
};

static int pitch4(int w) {  // multiple of 4 with pitch / 4 odd: 128-bit rows, consecutive rows 4 (mod 8) banks apart
  int p = (w + 3) & ~3;
  if (((p >> 2) & 1) == 0) p += 4;
  return p;
}
static int pitch_mod32_8(int w) {  // smallest p >= w with p = 8 (mod 32)
  int p = (w + 7) & ~7;
  while ((p & 31) != 8) p += 8;
  return p;
}

static void layout_small_net(const DevNet& src, SmNet& dst, int& ps, int& pt, int& tp, int& dp, int in_tape) {
  memset(&dst, 0, sizeof(dst));
  dst.depth = src.depth;
  dst.has_bias = src.has_bias;
  for (int j = 0; j <= src.depth; ++j) dst.w[j] = src.w[j];
  dst.ta[0] = in_tape;
  for (int j = 0; j < src.depth; ++j) {
    dst.act[j] = src.act[j];
    dst.pw[j] = src.p_w[j];
    dst.pb[j] = src.p_b[j];
    dst.op[j] = pitch4(src.w[j + 1]);
    dst.sw[j] = ps;
    ps += dst.op[j] * src.w[j];
    dst.sb[j] = ps;
    ps += (src.w[j + 1] + 3) & ~3;
    dst.kp[j] = pitch4(src.w[j]);
    dst.tw[j] = pt;
    pt += dst.kp[j] * src.w[j + 1];
    dst.ta[j + 1] = tp;
    tp += (src.w[j + 1] + 3) & ~3;
    dst.dl[j] = dp;
    dp += (src.w[j + 1] + 3) & ~3;
  }
}

void small_free_plan(dflow_chain* c) {
  SmallPlan* sp = c->small;
  if (!sp) return;
  if (sp->d_plan) cudaFree(sp->d_plan);
  if (sp->d_consts) cudaFree(sp->d_consts);
  delete sp;
  c->small = nullptr;
}

// Builds the plan if the chain fits the kernel: every Dense output <= 32 lanes, everything inside 227 KB of shared memory.
int small_build_plan(dflow_chain* c) {
  const DevChain* C = c->hc();
  const DevChainHdr& H = C->h;
  if (c->must_wide) return DFLOW_E_UNSUPPORTED;
  int wmax = 1;
  for (int ei = 0; ei < H.L; ++ei) {
    const DevElem& E = C->e[ei];
    if (E.kind == DFLOW_ELEM_NORM) continue;
    if (E.a > 32) return DFLOW_E_UNSUPPORTED;
    for (int ni = (E.kind == DFLOW_ELEM_RNVP ? 0 : 1); ni < 2; ++ni) {
      const DevNet& net = ni == 0 ? E.s : E.t;
      for (int j = 1; j <= net.depth; ++j) wmax = std::max(wmax, net.w[j]);
    }
  }
  if (wmax > 32) return DFLOW_E_UNSUPPORTED;
  const int lps = wmax <= 16 ? 16 : 32;
  const size_t bytes = sizeof(SmallPlanDev) + sizeof(SmElem) * (size_t)(H.L > 0 ? H.L - 1 : 0);
  std::vector<unsigned char> img((bytes + 15) & ~(size_t)15, 0);
  SmallPlanDev* P = reinterpret_cast<SmallPlanDev*>(img.data());
  P->d = H.d;
  P->n = H.n;
  P->L = H.L;
  P->lps = lps;
  P->P = H.P;
  P->has_theta_range = H.has_theta_range;
  P->logpdf_c0 = H.logpdf_c0;
  int ps = 0, pt = 0, tp = 0, dp = 0, nconst = 0;
  auto r4 = [](int v) { return (v + 3) & ~3; };
  P->t_x = tp; tp += r4(H.d);
  P->t_th = tp; tp += r4(H.n);
  P->t_gx = tp; tp += r4(H.d);
  std::vector<float> consts;
  for (int ei = 0; ei < H.L; ++ei) {
    const DevElem& E = C->e[ei];
    SmElem& S = P->e[ei];
    S.kind = E.kind;
    S.a = E.a;
    S.nid = E.nid;
    S.nin = E.nin;
    memcpy(S.af, E.af, sizeof(S.af));
    memcpy(S.id, E.id, sizeof(S.id));
    if (E.kind == DFLOW_ELEM_NORM) {
      S.nstage = nconst;
      nconst += 2 * H.d + 4;
      S.t_ck = (ei == H.L - 1) ? -1 : tp;  // a trailing normalisation acts on the data itself: nothing to restore
      if (S.t_ck >= 0) tp += r4(H.d);
      continue;
    }
    const int in_tape = tp;
    tp += r4(E.nin);
    int dpe = 0;  // the delta rows are reused element after element (their weight-gradient phase ends before the next one)
    S.t_ck = tp;
    tp += r4(E.a);
    S.t_em = tp;
    tp += r4(E.a);
    S.n_dense = 0;
    S.dw_start[0] = 0;
    for (int ni = (E.kind == DFLOW_ELEM_RNVP ? 0 : 1); ni < 2; ++ni) {
      const DevNet& net = ni == 0 ? E.s : E.t;
      SmNet& sn = ni == 0 ? S.s : S.t;
      layout_small_net(net, sn, ps, pt, tp, dpe, in_tape);
      for (int j = 0; j < net.depth; ++j) {
        const int di = S.n_dense++;
        S.dw_O[di] = net.w[j + 1];
        S.dw_K[di] = net.w[j];
        S.dw_op[di] = sn.op[j];
        S.dw_in[di] = sn.ta[j];
        S.dw_dl[di] = sn.dl[j];
        S.dw_g[di] = sn.sw[j];
        S.dw_gb[di] = sn.sb[j];
        // tasks: (output, quad of inputs) pairs, then the bias entries
        S.dw_start[di + 1] = S.dw_start[di] + net.w[j + 1] * ((net.w[j] + 3) / 4) + (net.has_bias ? net.w[j + 1] : 0);
      }
    }
    dp = std::max(dp, dpe);
  }
  P->PS = (ps + 3) & ~3;
  P->PT = (pt + 3) & ~3;
  P->TP = pitch_mod32_8(tp);
  P->DP = pitch_mod32_8(std::max(dp, 4));
  P->nconst = nconst;
  const int ng = SM_THREADS / lps;
  const size_t smem = img.size() + 4 * (size_t)((nconst + 3) & ~3) +
                      4 * ((size_t)5 * P->PS + (size_t)P->PT + (size_t)ng * P->TP + (size_t)ng * P->DP + ng + 8);
  if (smem > (size_t)c->max_smem_optin) return DFLOW_E_UNSUPPORTED;
  SmallPlan* sp = new (std::nothrow) SmallPlan();
  if (!sp) return DFLOW_E_NOMEM;
  sp->host = img;
  sp->lps = lps;
  sp->smem = smem;
  c->small = sp;
  if (cudaMalloc(&sp->d_plan, img.size()) != cudaSuccess ||
      cudaMalloc(&sp->d_consts, sizeof(float) * std::max(nconst, 4)) != cudaSuccess) {
    small_free_plan(c);
    return DFLOW_E_NOMEM;
  }
  return DFLOW_OK;
}

// Theta range and normalisation constants can change after creation (dflow_chain_set_theta_range): refresh per launch.
int small_train_epoch(dflow_chain* c, float* W, float* m, float* v, const float* x, const float* theta,
                      const int32_t* order, long long n, long long batchsize, float lr, float beta1, float beta2, float eps,
                      long long t0, int flags, float* loss2_out, cudaStream_t st) {
  SmallPlan* sp = c->small;
  const DevChain* C = c->hc();
  const DevChainHdr& H = C->h;
  SmallPlanDev* P = reinterpret_cast<SmallPlanDev*>(sp->host.data());
  P->has_theta_range = H.has_theta_range;
  for (int k = 0; k < NMAX; ++k) {
    P->theta_min[k] = H.theta_min[k];
    P->theta_rng[k] = H.theta_rng[k];
  }
  if (cudaMemcpyAsync(sp->d_plan, sp->host.data(), sp->host.size(), cudaMemcpyHostToDevice, st) != cudaSuccess) {
    set_error("plan upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    return DFLOW_E_CUDA;
  }
  // normalisation constants live in the chain's staged image (static blocks written at creation)
  for (int ei = 0; ei < H.L; ++ei) {
    const DevElem& E = C->e[ei];
    if (E.kind != DFLOW_ELEM_NORM) continue;
    if (cudaMemcpyAsync(sp->d_consts + P->e[ei].nstage, c->d_staged + E.stage_off, sizeof(float) * (2 * H.d + 4),
                        cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
      set_error("constant upload failed: %s", cudaGetErrorString(cudaGetLastError()));
      return DFLOW_E_CUDA;
    }
  }
  SmallArgs a;
  memset(&a, 0, sizeof(a));
  a.plan = sp->d_plan;
  a.plan_bytes = (int)sp->host.size();
  a.consts = sp->d_consts;
  a.W = W;
  a.m = m;
  a.v = v;
  a.x = x;
  a.theta = theta;
  a.order = order;
  a.n = n;
  a.batchsize = batchsize;
  a.lr = lr;
  a.b1 = beta1;
  a.b2 = beta2;
  a.eps = eps;
  adam_beta_powers(beta1, beta2, t0, &a.b1t, &a.b2t);
  a.flags = flags;
  a.loss2_out = loss2_out;
  cudaError_t e;
  if (sp->lps == 16) {
    e = cudaFuncSetAttribute(epoch_small_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp->smem);
    if (e == cudaSuccess) epoch_small_kernel<16><<<1, SM_THREADS, sp->smem, st>>>(a);
  } else {
    e = cudaFuncSetAttribute(epoch_small_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp->smem);
    if (e == cudaSuccess) epoch_small_kernel<32><<<1, SM_THREADS, sp->smem, st>>>(a);
  }
  if (e != cudaSuccess || (e = cudaGetLastError()) != cudaSuccess) {
    set_error("epoch_small_kernel launch failed: %s", cudaGetErrorString(e));
    return DFLOW_E_CUDA;
  }
  c->launches++;
  return DFLOW_OK;
}

}  // namespace dflow
