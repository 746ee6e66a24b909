// K3 v2: adjoint of the fused chain (loss + gradient in one launch), S samples per thread, register-tiled dW.
//
// Same mathematics as chain_grad_kernel (dflow_chain_kernels.cuh; reference: rrule of src/affine/RNVP.jl:99-147 +
// Dense pullbacks + loss seeds of src/Flows.jl:352-359), restructured after the round-1 profile:
//   * S samples per thread in adjacent column slots (128/64-bit shared-memory access, weights reused S times);
//   * delta_{j-1} overwrites h_{j-1} in place (no separate delta column block) -> less shared memory per sample;
//   * the weight-gradient reduction over a warp's 32*S samples is a register-tiled mini-GEMM: a lane owns a
//     TO x TK tile of dW, operands come in as LDS.128 over 4 samples, lanes of a warp share rows (multicast), so one
//     sample-quad costs TO+TK wavefronts for 4*TO*TK FFMA (was 2 LDS.128 per 4 FFMA).
#pragma once
#include "dflow_chain_kernels.cuh"

namespace dflow {

__host__ __device__ inline SmemPlan plan_grad2(const DevChainHdr& h, int chain_bytes, int nts, int smem_grad,
                                               int want_thbar = 0) {
  SmemPlan p;
  p.chain_f = ((chain_bytes + 15) / 16) * 4;
  p.w_f = h.resident ? h.stage_total : h.stage_max;
  p.cs = nts + 4;  // +4: consecutive rows start 4 banks apart (conflict-free multi-row LDS.128 in the dW phase)
  const int hd = h.max_depth > 1 ? h.max_depth - 1 : 1;
  // x, xbar, theta (+ its cotangent for dflow_vjp), hidden activations, s, t, output cotangent
  const int rows = 2 * h.d + h.n * (want_thbar ? 2 : 1) + hd * h.hp + 3 * h.amax4;
  p.cols_f = rows * p.cs;
  p.grad_f = smem_grad ? ((h.P + 3) / 4) * 4 : 0;
  return p;
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  acc = fmaf(a.w, b.w, acc);
  return acc;
}

// One pass of the weight-gradient mini-GEMM over the warp's `nq` sample quads.
// Lane layout: lk = lane & (LK-1) indexes TK-wide column groups, lo = lane / LK indexes TO-tall row groups, so the
// warp covers a (32/LK*TO) x (LK*TK) block of dW starting at (ob, kb).
template <int TO, int TK, int LK>
__device__ __forceinline__ void dw_block(const float* __restrict__ dcol, const InSel& in, int O, int K, int ob, int kb,
                                         bool with_bias, int p_w, int p_b, float* __restrict__ gsm,
                                         float* __restrict__ ggl, int CS, int wbase, int nq, int lane) {
  const int lk = lane & (LK - 1), lo = lane / LK;
  const int o0 = ob + lo * TO, k0 = kb + lk * TK;
  const float4* dp[TO];
  const float4* hp[TK];
  bool ov[TO], kv[TK];
#pragma unroll
  for (int i = 0; i < TO; ++i) {
    ov[i] = (o0 + i) < O;
    dp[i] = reinterpret_cast<const float4*>(dcol + (ov[i] ? (o0 + i) : 0) * CS + wbase);
  }
#pragma unroll
  for (int i = 0; i < TK; ++i) {
    kv[i] = (k0 + i) < K;
    hp[i] = reinterpret_cast<const float4*>(in(kv[i] ? (k0 + i) : 0) + wbase);
  }
  // packed accumulators: (even-sample partial sum, odd-sample partial sum) per dW entry; a sample quad is two FFMA2
  // whose operands are the natural register pairs of the two LDS.128 results
  float acc[TO][TK][2], bs[TO];
#pragma unroll
  for (int i = 0; i < TO; ++i) {
    bs[i] = 0.0f;
#pragma unroll
    for (int j = 0; j < TK; ++j) acc[i][j][0] = acc[i][j][1] = 0.0f;
  }
  const bool do_bias = with_bias && lk == 0 && kb == 0;
#pragma unroll 2
  for (int q = 0; q < nq; ++q) {
    float4 dv[TO], hv[TK];
#pragma unroll
    for (int i = 0; i < TO; ++i) dv[i] = dp[i][q];
#pragma unroll
    for (int j = 0; j < TK; ++j) hv[j] = hp[j][q];
#pragma unroll
    for (int i = 0; i < TO; ++i) {
#pragma unroll
      for (int j = 0; j < TK; ++j) {
        fma2_vv(acc[i][j][0], acc[i][j][1], dv[i].x, dv[i].y, hv[j].x, hv[j].y);
        fma2_vv(acc[i][j][0], acc[i][j][1], dv[i].z, dv[i].w, hv[j].z, hv[j].w);
      }
      if (do_bias) bs[i] += (dv[i].x + dv[i].y) + (dv[i].z + dv[i].w);
    }
  }
#pragma unroll
  for (int i = 0; i < TO; ++i) {
    if (!ov[i]) continue;
#pragma unroll
    for (int j = 0; j < TK; ++j) {
      if (!kv[j]) continue;
      const int gi = p_w + (o0 + i) + O * (k0 + j);
      const float tot = acc[i][j][0] + acc[i][j][1];
      if (gsm)
        atomicAdd(gsm + gi, tot);
      else
        atomicAdd(ggl + gi, tot);
    }
    if (do_bias) {
      if (gsm)
        atomicAdd(gsm + p_b + o0 + i, bs[i]);
      else
        atomicAdd(ggl + p_b + o0 + i, bs[i]);
    }
  }
}

// db[o] += sum_samples delta[o][s]: lane handles row ob + (lane & 15) and every other sample quad
__device__ __forceinline__ void db_pass(const float* __restrict__ dcol, int O, int p_b, float* gsm, float* ggl, int CS,
                                        int wbase, int nq, int lane) {
  for (int ob = 0; ob < O; ob += 16) {
    const int o = ob + (lane & 15);
    const bool ok = o < O;
    const float4* dp = reinterpret_cast<const float4*>(dcol + (ok ? o : 0) * CS + wbase);
    float s0 = 0.0f, s1 = 0.0f;
    for (int q = lane >> 4; q < nq; q += 2) {
      const float4 v = dp[q];
      s0 += v.x + v.y;
      s1 += v.z + v.w;
    }
    float s = s0 + s1;
    s += __shfl_xor_sync(0xffffffffu, s, 16);
    if (ok && lane < 16) {
      if (gsm)
        atomicAdd(gsm + p_b + o, s);
      else
        atomicAdd(ggl + p_b + o, s);
    }
  }
}

// dW[o][k] += sum_samples delta[o][s] * in_k[s] ; db[o] += sum_samples delta[o][s]   over the warp's samples
__device__ __forceinline__ void dw_phase2(const float* __restrict__ dcol, const InSel& in, int O, int K, int has_bias,
                                          int p_w, int p_b, float* gsm, float* ggl, int CS, int wbase, int nq,
                                          int lane) {
  if (O <= 4) {  // skinny rows (last Dense of a small-a layer): 4 x 16 blocks
    for (int kb = 0; kb < K; kb += 16)
      dw_block<1, 2, 8>(dcol, in, O, K, 0, kb, false, p_w, p_b, gsm, ggl, CS, wbase, nq, lane);
  } else if (K <= 4) {  // skinny columns (first Dense with few inputs): 16 x 4 blocks
    for (int ob = 0; ob < O; ob += 16)
      dw_block<1, 2, 2>(dcol, in, O, K, ob, 0, false, p_w, p_b, gsm, ggl, CS, wbase, nq, lane);
  } else {  // 16 x 16 blocks, 2 x 4 register tile per lane
    for (int ob = 0; ob < O; ob += 16)
      for (int kb = 0; kb < K; kb += 16)
        dw_block<2, 4, 4>(dcol, in, O, K, ob, kb, false, p_w, p_b, gsm, ggl, CS, wbase, nq, lane);
  }
  if (has_bias) db_pass(dcol, O, p_b, gsm, ggl, CS, wbase, nq, lane);
}

// g_in[k] = sum_o W[k][o] * delta[o] for k in [k0, K) and the thread's S samples.
// hidden (first == false): delta_{j-1}[k] = g_in[k] * act'(h_{j-1}[k]) written IN PLACE over h_{j-1}[k];
// first Dense: gx[axis_id[k-n]] += g_in[k].
template <int NG, int S>
__device__ __forceinline__ void dense_T2(const float* dcol, int CS, int sb, const float* __restrict__ Wst, int k0,
                                         int K, bool first, float* hprev, int actp, float* gx,
                                         const unsigned char* id, int n, float* gth = nullptr) {
  float dl[NG * 4][S];
#pragma unroll
  for (int o = 0; o < NG * 4; ++o) ld_samples<S>(dcol + o * CS + sb, dl[o]);
  for (int k = k0; k < K; ++k) {
    const float4* wr = reinterpret_cast<const float4*>(Wst + k * (NG * 4));
    float a0[S], a1[S];
#pragma unroll
    for (int s = 0; s < S; ++s) a0[s] = a1[s] = 0.0f;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const float4 w = wr[g];
      fma_samples<S>(a0, dl[4 * g + 0], w.x);
      fma_samples<S>(a1, dl[4 * g + 1], w.y);
      fma_samples<S>(a0, dl[4 * g + 2], w.z);
      fma_samples<S>(a1, dl[4 * g + 3], w.w);
    }
    float* dst = first ? (k < n ? gth + k * CS + sb : gx + (int)id[k - n] * CS + sb) : hprev + k * CS + sb;
    float cur[S];
    ld_samples<S>(dst, cur);
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const float v = a0[s] + a1[s];
      cur[s] = first ? cur[s] + v : v * act_grad(actp, cur[s]);
    }
    st_samples<S>(dst, cur);
  }
}

// Back-propagate through one conditioner.  gb holds the output cotangent; hidden activations h_j live in
// hc + j*hstride and are overwritten by delta_j on the way down.
template <int HP, int S>
__device__ __forceinline__ void net_backward2(const DevChainHdr& H, const DevElem& E, const DevNet& net,
                                              const float* __restrict__ wblk, float* xs, float* gx, float* th,
                                              float* hc, int hstride, float* gb, const float* outvals, float* gsm,
                                              float* ggl, int CS, int sb, int wbase, int nq, int lane,
                                              float* gth = nullptr) {
  const int D = net.depth;
  {
    const int actL = net.act[D - 1];
    if (actL != DFLOW_ACT_IDENTITY) {
      const int O = net.w[D];
      for (int o = 0; o < O; ++o)
#pragma unroll
        for (int s = 0; s < S; ++s) gb[o * CS + sb + s] *= act_grad(actL, outvals[o * CS + sb + s]);
    }
  }
  for (int j = D - 1; j >= 0; --j) {
    const float* dcol = (j == D - 1) ? gb : hc + j * hstride;
    float* hprev = hc + (j > 0 ? (j - 1) * hstride : 0);
    InSel in{th, xs, hprev, E.id, H.n, CS, j == 0};
    const int O = net.w[j + 1], K = net.w[j];
    __syncwarp();
    dw_phase2(dcol, in, O, K, net.has_bias, net.p_w[j], net.p_b[j], gsm, ggl, CS, wbase, nq, lane);
    __syncwarp();
    const float* Wst = wblk + net.s_w[j];
    const bool first = (j == 0);
    const int actp = j > 0 ? net.act[j - 1] : 0;
    const int k0 = (first && !gth) ? H.n : 0;  // θ rows (k < n) of the first Dense are dropped unless θ̄ is wanted
    switch (net.op[j] >> 2) {
      case 1: dense_T2<1, S>(dcol, CS, sb, Wst, k0, K, first, hprev, actp, gx, E.id, H.n, gth); break;
      case 2: dense_T2<2, S>(dcol, CS, sb, Wst, k0, K, first, hprev, actp, gx, E.id, H.n, gth); break;
      case 4: dense_T2<4, S>(dcol, CS, sb, Wst, k0, K, first, hprev, actp, gx, E.id, H.n, gth); break;
      case 8: dense_T2<8, S>(dcol, CS, sb, Wst, k0, K, first, hprev, actp, gx, E.id, H.n, gth); break;
      default: dense_T2<16, S>(dcol, CS, sb, Wst, k0, K, first, hprev, actp, gx, E.id, H.n, gth); break;
    }
  }
}

template <int HP, int S>
constexpr int grad2_max_threads() {
  // hidden 16, S = 2: one 384-thread CTA per SM (12 warps; 448 threads measured slower: 224 KB of shared memory leave no L1) instead of one 256-thread CTA -- the kernel is latency-bound
  // and shared memory (240 B of columns per sample at C2) is what limits the resident warps
  return (HP == 16 && S == 4) ? 192 : HP * S >= 64 ? 128 : (HP == 16 && S == 2 ? 384 : 256);
}
template <int HP, int S>
constexpr int grad2_min_ctas() {
  return grad2_max_threads<HP, S>() > 256 ? 1 : 2;
}

// VJP = false is the loss-seeded train step exactly as profiled in round 1 (no cotangent inputs / outputs compiled in);
// VJP = true adds the caller's cotangents (zbar, jbar) and the input cotangents (xbar, thetabar) of dflow_vjp.
template <int HP, int S, bool FIXED, bool VJP>
__device__ __forceinline__ void chain_grad2_body(const GradArgs& a) {
  extern __shared__ float4 smem4[];
  float* smem = reinterpret_cast<float*>(smem4);
  const int tid = threadIdx.x, NT = FIXED ? grad2_max_threads<HP, S>() : (int)blockDim.x, NTS = NT * S, lane = tid & 31;
  const int sb = tid * S;              // this thread's first column slot
  const int wbase = (tid & ~31) * S;   // first slot of this warp
  constexpr int NQ = 32 * S / 4;       // sample quads per warp
  copy_f4(smem, reinterpret_cast<const float*>(a.chain), (a.chain_bytes + 15) / 16, tid, NT);
  __syncthreads();
  const DevChain* C = reinterpret_cast<const DevChain*>(smem);
  const DevChainHdr& H = C->h;
  const bool want_th = VJP && a.thbar_out != nullptr && H.n > 0;
  const SmemPlan P = plan_grad2(H, a.chain_bytes, NTS, a.smem_grad, want_th ? 1 : 0);
  float* wsm = smem + P.chain_f;
  float* cols = wsm + P.w_f;
  float* gsm = cols + P.cols_f;
  const int CS = FIXED ? NTS + 4 : P.cs;
  const int d = H.d, n = H.n, L = H.L;
  const int hd = H.max_depth > 1 ? H.max_depth - 1 : 1;
  const int hstride = HP * CS;
  float* xs = cols;
  float* gx = xs + d * CS;
  float* th = gx + d * CS;
  float* gth = want_th ? th + n * CS : nullptr;  // cotangent of the (normalised) conditions
  float* hc = th + (want_th ? 2 : 1) * n * CS;
  float* ob = hc + hd * hstride;  // s values
  float* tb = ob + H.amax4 * CS;  // t values
  float* gb = tb + H.amax4 * CS;  // output cotangent  (exp(-s) is recomputed from the kept s values: 4 fewer column
                                  // rows per sample buy two more resident warps)

  if (H.resident) copy_f4(wsm, a.staged, H.stage_total / 4, tid, NT);
  if (a.smem_grad)
    for (int i = tid; i < P.grad_f; i += NT) gsm[i] = 0.0f;
  __syncthreads();
  float* gacc = a.smem_grad ? gsm : nullptr;

  const long long ntiles = (a.B + NTS - 1) / NTS;
  float lsum_thread = 0.0f, nonfinite = 0.0f;
  float* ckb = a.ws + (size_t)blockIdx.x * H.ck_total * NTS + sb;  // this thread's checkpoint slots

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long base = tile * NTS;
    float ib[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const long long gi = base + sb + s;
      const bool valid = gi < a.B;
      const long long src = (valid && a.idx) ? (long long)a.idx[gi] : gi;
      const float* xp = a.x_in + src * d;
      for (int k = 0; k < d; ++k) xs[k * CS + sb + s] = valid ? __ldg(xp + k) : 0.0f;
      for (int k = 0; k < n; ++k) {
        float v = (valid && a.theta) ? __ldg(a.theta + src * n + k) : 0.0f;
        if (a.flags & DFLOW_THETA_NORMALIZE) v = (H.theta_rng[k] == 0.0f) ? 0.0f : (v - H.theta_min[k]) / H.theta_rng[k];
        th[k * CS + sb + s] = v;
      }
      // ib = -j̄ of this sample: the loss seed 1/B_tot (src/Flows.jl:352-359), or minus the caller's cotangent of ln_det_jac
      if constexpr (VJP) {
        ib[s] = valid ? (a.jbar ? -__ldg(a.jbar + gi) : a.inv_btot) : 0.0f;
        if (want_th)
          for (int k = 0; k < n; ++k) gth[k * CS + sb + s] = 0.0f;
      } else {
        ib[s] = valid ? a.inv_btot : 0.0f;
      }
    }
    // ---- forward (normalising) sweep: last element first; checkpoint what each element changes ----
    float ldj[S];
#pragma unroll
    for (int s = 0; s < S; ++s) ldj[s] = 0.0f;
    for (int step = 0; step < L; ++step) {
      const DevElem& E = C->e[L - 1 - step];
      const float* wblk;
      if (H.resident) {
        wblk = wsm + E.stage_off;
      } else {
        __syncthreads();
        copy_f4(wsm, a.staged + E.stage_off, E.stage_len / 4, tid, NT);
        __syncthreads();
        wblk = wsm;
      }
      if (E.ck_len > 0) {
        float* ck = ckb + (size_t)E.ck_off * NTS;
        float v[S];
        if (E.kind == DFLOW_ELEM_NORM)
          for (int k = 0; k < d; ++k) {
            ld_samples<S>(xs + k * CS + sb, v);
            st_samples<S>(ck + (size_t)k * NTS, v);
          }
        else
          for (int j = 0; j < E.a; ++j) {
            ld_samples<S>(xs + (int)E.af[j] * CS + sb, v);
            st_samples<S>(ck + (size_t)j * NTS, v);
          }
      }
      elem_apply<HP, S, false>(H, E, wblk, false, xs, th, hc, 0, ob, tb, CS, tid, NT, ldj);
    }
    // ---- loss and seeds: z̄ = z * inv_btot, j̄ = -inv_btot (src/Flows.jl:352-359), or the caller's cotangents ----
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const long long gi = base + sb + s;
      float q = 0.0f;
      for (int k = 0; k < d; ++k) {
        const float v = xs[k * CS + sb + s];
        q = fmaf(v, v, q);
        if constexpr (VJP)
          gx[k * CS + sb + s] = a.zbar ? (gi < a.B ? __ldg(a.zbar + gi * d + k) : 0.0f) : v * ib[s];
        else
          gx[k * CS + sb + s] = v * ib[s];
      }
      const float lp = H.logpdf_c0 - 0.5f * q + ldj[s];
      if (gi < a.B) {
        if (isfinite(lp))
          lsum_thread += lp;
        else
          nonfinite += 1.0f;
      }
    }

    // ---- reverse sweep in chain order ----
    for (int ei = 0; ei < L; ++ei) {
      const DevElem& E = C->e[ei];
      const float* wblk;
      if (H.resident) {
        wblk = wsm + E.stage_off;
      } else {
        __syncthreads();
        copy_f4(wsm, a.staged + E.stage_off, E.stage_len / 4, tid, NT);
        __syncthreads();
        wblk = wsm;
      }
      if (E.kind == DFLOW_ELEM_NORM) {
        const float alpha = wblk[2 * d], beta = wblk[2 * d + 1];
        for (int k = 0; k < d; ++k) {
          const float xmin = wblk[k], xmax = wblk[d + k];
          const float sc = (beta - alpha) / (xmax - xmin);
#pragma unroll
          for (int s = 0; s < S; ++s) {
            gx[k * CS + sb + s] *= sc;
            if (E.ck_len > 0) xs[k * CS + sb + s] = ckb[(size_t)(E.ck_off + k) * NTS + s];
          }
        }
        continue;
      }
      const bool rnvp = (E.kind == DFLOW_ELEM_RNVP);
      InSel in{th, xs, hc, E.id, n, CS, true};
      const int a4 = H.amax4;
      for (int ni = rnvp ? 0 : 1; ni < 2; ++ni) {
        const DevNet& net = ni == 0 ? E.s : E.t;
        float* outc = ni == 0 ? ob : tb;
        run_net<HP, S>(net, wblk, in, hc, hstride, outc, CS, tid, NT);  // recompute, keep hidden activations
        for (int j = 0; j < a4; ++j) {
#pragma unroll
          for (int s = 0; s < S; ++s) {
            float gout = 0.0f;
            if (j < E.a) {
              const int k = E.af[j];
              if (ni == 0) {
                gout = -gx[k * CS + sb + s] * xs[k * CS + sb + s] + ib[s];  // s̄ = -z̄_af z_af - j̄, RNVP.jl:134
              } else {
                const float em = rnvp ? expf(-ob[j * CS + sb + s]) : 1.0f;
                gout = -gx[k * CS + sb + s] * em;  // t̄, RNVP.jl:135
              }
            }
            gb[j * CS + sb + s] = gout;
          }
        }
        net_backward2<HP, S>(H, E, net, wblk, xs, gx, th, hc, hstride, gb, outc, gacc, a.grad_out, CS, sb, wbase, NQ,
                             lane, VJP ? gth : nullptr);
      }
      // restore the layer input from its checkpoint and finish ū (RNVP.jl:137-139)
      for (int j = 0; j < E.a; ++j) {
        const int k = E.af[j];
#pragma unroll
        for (int s = 0; s < S; ++s) {
          xs[k * CS + sb + s] = ckb[(size_t)(E.ck_off + j) * NTS + s];
          if (rnvp) gx[k * CS + sb + s] *= expf(-ob[j * CS + sb + s]);
        }
      }
    }
    // ---- cotangents of the inputs (dflow_vjp) ----
    if (VJP && (a.xbar_out || want_th)) {
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const long long gi = base + sb + s;
        if (gi >= a.B) continue;
        if (a.xbar_out)
          for (int k = 0; k < d; ++k) a.xbar_out[gi * d + k] = gx[k * CS + sb + s];
        if (want_th)
          for (int k = 0; k < n; ++k) {
            float v = gth[k * CS + sb + s];
            // chain rule through normalize_input (src/Data.jl:213-218): d θ̂ / d θ = 1 / (θ_max - θ_min), 0 for a zero range
            if (a.flags & DFLOW_THETA_NORMALIZE) v = (H.theta_rng[k] == 0.0f) ? 0.0f : v / H.theta_rng[k];
            a.thbar_out[gi * n + k] = v;
          }
      }
    }
    __syncwarp();
  }

  // ---- flush ----
  __syncthreads();
  if (a.smem_grad)
    for (int i = tid; i < H.P; i += NT) {
      const float v = gsm[i];
      if (v != 0.0f) atomicAdd(a.grad_out + i, v);
    }
  float v0 = lsum_thread, v1 = nonfinite;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v0 += __shfl_xor_sync(0xffffffffu, v0, o);
    v1 += __shfl_xor_sync(0xffffffffu, v1, o);
  }
  float* red = cols;
  if (lane == 0) {
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
    if (a.loss_out) {
      atomicAdd(a.loss_out, t0);
      if (t1 != 0.0f) atomicAdd(a.loss_out + 1, t1);
    }
  }
}

template <int HP, int S>
__global__ void __launch_bounds__((grad2_max_threads<HP, S>()), (grad2_min_ctas<HP, S>())) chain_grad2_kernel(const GradArgs a) {
  if (blockDim.x == grad2_max_threads<HP, S>())
    chain_grad2_body<HP, S, true, false>(a);
  else
    chain_grad2_body<HP, S, false, false>(a);
}

// the same sweep with caller cotangents and input-cotangent outputs (dflow_vjp): a kernel of its own, so that the train
// step's register allocation is not shaped by the pullback's extra state
template <int HP, int S>
__global__ void __launch_bounds__((grad2_max_threads<HP, S>()), (grad2_min_ctas<HP, S>())) chain_vjp2_kernel(const GradArgs a) {
  if (blockDim.x == grad2_max_threads<HP, S>())
    chain_grad2_body<HP, S, true, true>(a);
  else
    chain_grad2_body<HP, S, false, true>(a);
}

template <int HP, int S>
cudaError_t launch_grad2_inst(const GradArgs& a, unsigned grid, int nt, size_t smem, cudaStream_t st);

}  // namespace dflow
