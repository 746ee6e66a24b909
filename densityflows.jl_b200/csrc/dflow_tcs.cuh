// Tensor-core conditioner kernel for NARROW nets (hidden 32 / 64): the hi parts of every A operand live in TMEM
// (tcgen05.mma with the A matrix read from tensor memory), FOUR independent 128-sample tiles are in flight per SM, and the
// whole conditioner's weights stay resident in shared memory.  Included by dflow_tc.cu (same TcArgs, same pre-split weight
// image, same training buffers as tc_net_kernel, which keeps serving the wider nets).
//
// Why a second kernel: at hidden 64 the warp-specialised pipeline of tc_net_kernel is bound by its per-tile chain of
// hand-offs and memory round trips, not by the tensor pipe (ncu, profiles/r02_c3_tc.md: 12-23 % tensor-pipe active), and its
// TMEM / shared-memory footprint allows only two tiles in flight per SM.  Here a 512-thread CTA holds four warpgroups; each
// walks its own tile through
//   input row -> TMEM | D1 = b1 + in W1^T | act, hi/lo split in registers | D2 = b2 + h1 W2^T | ... | D3 = b3 + h2 W3^T | output
// (every accumulator is initialised with its bias by tcgen05.st) with one commit / wait and one 128-thread named barrier per
// GEMM; the four chains hide each other's latencies.
//
// Per chain, 128 TMEM columns (H = hidden width <= 64):
//   A = [0, H)    D1, then h1.hi in place, finally D3 (h1 is dead by then)
//   C = [64, 64+H) input row hi | lo, then D2, then h2.hi in place
// and one shared-memory buffer [128 x H] (K-major core layout) for the lo parts h1.lo / h2.lo: each product is
//   D += A.lo (shared memory) * B.hi  +  A.hi (TMEM) * B.lo  +  A.hi (TMEM) * B.hi.
// Reference math as in tc_net_kernel: src/affine/RNVP.jl:77-96 (normalising), :99-147 (adjoint), :150-205 (sampling).
#pragma once

namespace dflow {

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
          taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[128 x n] (+)= A[128 x 8] (TMEM, one tf32 element per column) * B[n x 8]^T (shared memory descriptor)
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

constexpr int TCS_CHAINS = 4;  // warpgroups = tiles in flight per CTA
constexpr int TCS_THREADS = 128 * TCS_CHAINS;
enum { TCS_BAR_W = 0, TCS_BAR_MMA = 1, TCS_BAR_SLOT = 1 + TCS_CHAINS, TCS_BAR_WORDS = 2 + TCS_CHAINS };

// a conditioner image this kernel can run: one D1 group, one pass, the input row (hi | lo) and D3 fit their regions
__host__ __device__ inline bool tcs_image_ok(const TcNetImg& im) {
  return im.halves <= 1 && im.passes == 1 && im.ng == 1 && im.NH == im.H && (im.H == 32 || im.H == 64) && 2 * im.K0p <= im.H &&
         im.N3p <= im.H && im.K0p <= 32 && im.N3p <= 32;
}
inline size_t tcs_smem_bytes(const TcNetImg& im) {
  return (size_t)(((2 * im.H + im.N3p + 3) & ~3) + im.blocks_floats + TCS_CHAINS * 128 * im.H) * 4 + TCS_BAR_WORDS * 8;
}

template <int MODE>
__global__ void __launch_bounds__(TCS_THREADS, 1) tcs_net_kernel(const __grid_constant__ TcArgs a) {
  extern __shared__ float4 smem4[];
  float* smem = reinterpret_cast<float*>(smem4);
  const TcNetImg& im = a.im;
  const int tid = threadIdx.x, warp = tid >> 5, wg = tid >> 7, row = tid & 127;
  const int K0p = im.K0p, H = im.H, N3p = im.N3p;
  const int d = a.d, n = a.n;
  const int nbias = (2 * H + N3p + 3) & ~3;
  float* biasS = smem;
  float* wblk = biasS + nbias;                             // [G1 | S2 | S3] blocks of the conditioner, resident
  float* lobuf = wblk + im.blocks_floats + wg * 128 * H;   // this chain's lo parts, [128 x H] K-major core layout
  uint64_t* bars = reinterpret_cast<uint64_t*>(wblk + im.blocks_floats + TCS_CHAINS * 128 * H);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + TCS_BAR_SLOT);
  const float* gimg = a.img + im.off;

  if (tid == 0) {
    mbar_init(bars + TCS_BAR_W, 1);
    for (int i = 0; i < TCS_CHAINS; ++i) mbar_init(bars + TCS_BAR_MMA + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512u);
  for (int i = tid; i < nbias; i += TCS_THREADS) biasS[i] = i < 2 * H + N3p ? __ldg(gimg + im.bias_off + i) : 0.0f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;
  if (tid == 0) {
    const uint32_t total = (uint32_t)im.blocks_floats * 4u;
    mbar_expect_tx(bars + TCS_BAR_W, total);
    uint32_t done = 0;
    while (done < total) {
      const uint32_t piece = min(total - done, 65536u);
      bulk_g2s(reinterpret_cast<char*>(wblk) + done, reinterpret_cast<const char*>(gimg + im.g1_off) + done, piece, bars + TCS_BAR_W);
      done += piece;
    }
  }
  const long long ntiles = (a.B + 127) / 128;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t tA = tbase + (uint32_t)(wg * 128), tC = tA + 64u;
  // weight-block descriptors (K-major core layout, see tc_prepack_kernel); offsets in 16-byte descriptor units
  const uint64_t dW1 = desc_at(desc_hi(K0p), smem_u32(wblk));
  const uint64_t dW16 = desc_at(desc_hi(WKC), smem_u32(wblk));
  const uint64_t dLo = desc_at(desc_hi(H), smem_u32(lobuf));  // A.lo, 8-column K step = 16 descriptor units
  const uint32_t g1_lo = ((uint32_t)(H * K0p) * 4u) >> 4, w2_lo = ((uint32_t)(H * WKC) * 4u) >> 4,
                 w3_lo = ((uint32_t)(N3p * WKC) * 4u) >> 4;
  const uint32_t s2_step = ((uint32_t)im.s2_floats * 4u) >> 4, s3_step = ((uint32_t)im.s3_floats * 4u) >> 4;
  const uint32_t res_s2 = ((uint32_t)(im.s2_off - im.g1_off) * 4u) >> 4, res_s3 = ((uint32_t)(im.s3_off - im.g1_off) * 4u) >> 4;
  const uint32_t id12 = instr_desc_tf32(H), id3 = instr_desc_tf32(N3p);
  const uint32_t bar_mma = smem_u32(bars + TCS_BAR_MMA + wg);
  uint32_t ph = 0;
  bool wready = false;
  const int nchunk = H >> 5;

  // hi / lo split of a 32-unit chunk: hi to the TMEM columns hi_col.. of this thread's lane, lo to the chain's
  // shared-memory operand (row = this sample, columns k0..k0+31)
  auto split_store = [&](const float (&v)[32], uint32_t hi_col, int k0) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t h[16];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        float4 lo;
        float hi;
        hi = tf32_hi(v[16 * half + 4 * q4 + 0]); h[4 * q4 + 0] = __float_as_uint(hi); lo.x = v[16 * half + 4 * q4 + 0] - hi;
        hi = tf32_hi(v[16 * half + 4 * q4 + 1]); h[4 * q4 + 1] = __float_as_uint(hi); lo.y = v[16 * half + 4 * q4 + 1] - hi;
        hi = tf32_hi(v[16 * half + 4 * q4 + 2]); h[4 * q4 + 2] = __float_as_uint(hi); lo.z = v[16 * half + 4 * q4 + 2] - hi;
        hi = tf32_hi(v[16 * half + 4 * q4 + 3]); h[4 * q4 + 3] = __float_as_uint(hi); lo.w = v[16 * half + 4 * q4 + 3] - hi;
        *reinterpret_cast<float4*>(lobuf + core_idx(row, k0 + 16 * half + 4 * q4, H)) = lo;
      }
      tmem_st16(hi_col + lane_off + 16u * half, h);
    }
  };
  // accumulator init: `ncols` (a multiple of 16) bias values from shared memory -> this thread's lane, columns col..
  // (the MMAs then accumulate on top of the bias: no bias add in the epilogue)
  auto bias_init = [&](uint32_t col, const float* b, int ncols) {
    for (int c0 = 0; c0 < ncols; c0 += 16) {
      uint32_t r[16];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const float4 bv = *reinterpret_cast<const float4*>(b + c0 + 4 * q4);
        r[4 * q4 + 0] = __float_as_uint(bv.x);
        r[4 * q4 + 1] = __float_as_uint(bv.y);
        r[4 * q4 + 2] = __float_as_uint(bv.z);
        r[4 * q4 + 3] = __float_as_uint(bv.w);
      }
      tmem_st16(col + lane_off + (uint32_t)c0, r);
    }
  };
  auto ld32 = [&](uint32_t taddr, float (&v)[32]) {
    uint32_t r0[16], r1[16];
    tmem_ld16_nowait(taddr, r0);
    tmem_ld16_nowait(taddr + 16, r1);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      v[j] = __uint_as_float(r0[j]);
      v[16 + j] = __uint_as_float(r1[j]);
    }
  };
  // this chain's TMEM / shared-memory operand stores are complete and ordered before the MMAs its elected lane issues
  auto publish = [&]() {
    tmem_st_wait();
    fence_async_smem();
    tc_fence_before();
    named_bar_sync(1 + wg, 128);
  };
  auto wait_mma = [&]() {
    mbar_wait(bars + TCS_BAR_MMA + wg, ph);
    ph ^= 1u;
    tc_fence_after();
  };
  // D (+)= A (K = 8 * ksteps: hi in TMEM at a_hi.., lo in the chain's shared-memory operand) * the 16-wide weight blocks
  // at b0, b0 + chunk_step, ... (lo part of a block lo_off further): lo*hi + hi*lo + hi*hi per K step of 8
  auto gemm_mixed = [&](uint32_t td, uint32_t a_hi, int ksteps, uint32_t idesc, uint64_t b0, uint32_t lo_off, uint32_t chunk_step) {
    uint32_t acc = MODE != TC_BWD ? 1u : 0u;  // forward modes: the accumulator was initialised with the bias
    for (int ks = 0; ks < ksteps; ++ks) {
      const uint64_t db = b0 + (uint64_t)((uint32_t)(ks >> 1) * chunk_step + (uint32_t)(ks & 1) * 16u);
      mma_tf32(td, dLo + (uint64_t)((uint32_t)ks * 16u), db, idesc, acc);
      mma_tf32_ts(td, a_hi + (uint32_t)ks * 8u, db + lo_off, idesc, 1u);
      mma_tf32_ts(td, a_hi + (uint32_t)ks * 8u, db, idesc, 1u);
      acc = 1u;
    }
  };

  for (long long tile = (long long)blockIdx.x * TCS_CHAINS + wg; tile < ntiles; tile += (long long)gridDim.x * TCS_CHAINS) {
    const long long gi = tile * 128 + row;
    const bool valid = gi < a.B;
    // this sample's column in the tile-blocked state / theta arrays: element k at base[k * 128] (one 64-bit base per
    // array and tile, 32-bit offsets per gathered coordinate)
    const size_t sd = (size_t)tile * (size_t)d * 128 + (size_t)row, sn = (size_t)tile * (size_t)n * 128 + (size_t)row;
    const float* const xin_t = a.x_in ? a.x_in + sd : nullptr;
    float* const xout_t = a.x_out ? a.x_out + sd : nullptr;
    float* const zbar_t = a.zbar ? a.zbar + sd : nullptr;
    const float* const zout_t = a.zout ? a.zout + sd : nullptr;
    const float* const th_t = a.theta ? a.theta + sn : nullptr;
    // relu masks of the adjoint chain, fetched before anything waits
    uint32_t mw2[2] = {0u, 0u}, mw1[2] = {0u, 0u};
    if constexpr (MODE == TC_BWD) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (c < nchunk && a.act2 == DFLOW_ACT_RELU) mw2[c] = a.m2buf[((size_t)tile * nchunk + c) * 128 + row];
        if (c < nchunk && a.act1 == DFLOW_ACT_RELU) mw1[c] = a.m1buf[((size_t)tile * nchunk + c) * 128 + row];
      }
    }
    // ---- 1. GEMM-1 operand row of this sample -> TMEM region C (hi at C, lo at C + K0p) ----
    for (int k0 = 0; k0 < K0p; k0 += 8) {
      float v[8];
      if constexpr (MODE == TC_BWD) {
        // delta3 of this conditioner (src/affine/RNVP.jl:118-127): s: -zbar_af * z_af - jbar, t: -zbar_af * exp(-s)
        const float njbar = (a.jbar && valid) ? -__ldg(a.jbar + gi) : a.inv_btot;
        float zb[8], zo[8], sv[8];
#pragma unroll
        for (int qq = 0; qq < 8; ++qq) {
          const int j = k0 + qq;
          zb[qq] = zo[qq] = sv[qq] = 0.0f;
          if (valid && j < a.a) {
            const int k = a.af[j];
            zb[qq] = zbar_t[k * 128];
            if (a.net_id == 0)
              zo[qq] = zout_t[k * 128];
            else if (a.has_s)
              sv[qq] = a.sbuf[((size_t)tile * a.a16 + j) * 128 + row];
          }
        }
#pragma unroll
        for (int qq = 0; qq < 8; ++qq) {
          const int j = k0 + qq;
          float val = 0.0f;
          if (valid && j < a.a) val = a.net_id == 0 ? -zb[qq] * zo[qq] + njbar : -zb[qq] * expf(-sv[qq]);
          v[qq] = val;
        }
#pragma unroll
        for (int qq = 0; qq < 8; ++qq) a.d3buf[tbuf_idx(tile, K0p, k0 + qq, row)] = v[qq];
      } else {
        // conditioner input row [theta_0..theta_{n-1}, x[axis_id...], 0 pad] (src/affine/RNVP.jl:157)
#pragma unroll
        for (int qq = 0; qq < 8; ++qq) {
          const int k = k0 + qq;
          float val = 0.0f;
          if (valid && k < a.nin) {
            if (k < n)
              val = a.theta_const ? __ldg(a.theta_const + k) : __ldg(th_t + k * 128);
            else
              val = xin_t[(int)a.id[k - n] * 128];
          }
          v[qq] = val;
        }
        if (a.flags & DFLOW_THETA_NORMALIZE) {
#pragma unroll
          for (int qq = 0; qq < 8; ++qq) {
            const int k = k0 + qq;
            if (valid && k < n) v[qq] = (a.theta_rng[k] == 0.0f) ? 0.0f : (v[qq] - a.theta_min[k]) / a.theta_rng[k];
          }
        }
        if constexpr (MODE == TC_FWD_STORE) {
          if (a.net_id >= 1) {
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) a.inbuf[tbuf_idx(tile, K0p, k0 + qq, row)] = v[qq];
          }
        }
      }
      uint32_t h[8], l[8];
#pragma unroll
      for (int qq = 0; qq < 8; ++qq) {
        const float hi = tf32_hi(v[qq]);
        h[qq] = __float_as_uint(hi);
        l[qq] = __float_as_uint(v[qq] - hi);
      }
      tmem_st8(tC + lane_off + (uint32_t)k0, h);
      tmem_st8(tC + lane_off + (uint32_t)(K0p + k0), l);
    }
    if constexpr (MODE != TC_BWD) bias_init(tA, biasS, H);  // D1 starts from b1
    publish();
    // ---- 2. D1 (region A) = input row * M1^T ----
    if ((warp & 3) == 0) {
      if (!wready) mbar_wait(bars + TCS_BAR_W, 0);
      tc_fence_after();
      if (elect_one()) {
        uint32_t acc = MODE != TC_BWD ? 1u : 0u;
        for (int ks = 0; ks < (K0p >> 3); ++ks) {
          const uint64_t db = dW1 + (uint64_t)((uint32_t)ks * 16u);
          mma_tf32_ts(tA, tC + (uint32_t)(K0p + ks * 8), db, id12, acc);
          mma_tf32_ts(tA, tC + (uint32_t)(ks * 8), db + g1_lo, id12, 1u);
          mma_tf32_ts(tA, tC + (uint32_t)(ks * 8), db, id12, 1u);
          acc = 1u;
        }
        mma_commit_a(bar_mma);
      }
      __syncwarp();
    }
    wready = true;
    if constexpr (MODE != TC_BWD) {
      if (a.x_out != a.x_in && a.net_id >= 1) {  // out-of-place (training sweep): carry the whole state
        for (int k0 = 0; k0 < d; k0 += 8) {
          float v[8];
#pragma unroll
          for (int qq = 0; qq < 8; ++qq) v[qq] = (k0 + qq < d) ? xin_t[(k0 + qq) * 128] : 0.0f;
#pragma unroll
          for (int qq = 0; qq < 8; ++qq)
            if (k0 + qq < d) xout_t[(k0 + qq) * 128] = v[qq];
        }
      }
    }
    wait_mma();
    // ---- 3. epilogue 1: h1 = act(D1 + b1) (adjoint: delta2 = D1 * act'(h2)) -> hi in place, lo to shared memory ----
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      if (c >= nchunk) break;
      float v[32];
      ld32(tA + lane_off + (uint32_t)(c * 32), v);
      if constexpr (MODE == TC_BWD) {
        if (a.act2 == DFLOW_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = ((mw2[c] >> j) & 1u) ? v[j] : 0.0f;
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = v[j] * act_grad(a.act2, a.h2buf[tbuf_idx(tile, H, c * 32 + j, row)]);
        }
      } else {
        if (a.act1 == DFLOW_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = act_apply(a.act1, v[j]);
        }
      }
      split_store(v, tA + (uint32_t)(c * 32), c * 32);
      if constexpr (MODE == TC_BWD) {
        float* grow = a.d2buf + tbuf_idx(tile, H, c * 32, row);
#pragma unroll
        for (int j = 0; j < 32; ++j) grow[j * 16] = v[j];
      } else if constexpr (MODE == TC_FWD_STORE) {
        float* grow = a.h1buf + tbuf_idx(tile, H, c * 32, row);
#pragma unroll
        for (int j = 0; j < 32; ++j) grow[j * 16] = v[j];
        if (a.act1 == DFLOW_ACT_RELU) a.m1buf[((size_t)tile * nchunk + c) * 128 + row] = positive_mask(v);
      }
    }
    if constexpr (MODE != TC_BWD) bias_init(tC, biasS + H, H);  // D2 starts from b2 (the input row in region C is dead)
    publish();
    // ---- 4. D2 (region C) = h1 * M2^T ----
    if ((warp & 3) == 0) {
      tc_fence_after();
      if (elect_one()) {
        gemm_mixed(tC, tA, H >> 3, id12, dW16 + res_s2, w2_lo, s2_step);
        mma_commit_a(bar_mma);
      }
      __syncwarp();
    }
    wait_mma();
    // ---- 5. epilogue 2: h2 = act(D2 + b2) (adjoint: delta1 = D2 * act'(h1)) -> hi in place, lo to shared memory ----
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      if (c >= nchunk) break;
      float v[32];
      ld32(tC + lane_off + (uint32_t)(c * 32), v);
      if constexpr (MODE == TC_BWD) {
        if (a.act1 == DFLOW_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = ((mw1[c] >> j) & 1u) ? v[j] : 0.0f;
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = v[j] * act_grad(a.act1, a.h1buf[tbuf_idx(tile, H, c * 32 + j, row)]);
        }
      } else {
        if (a.act2 == DFLOW_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = act_apply(a.act2, v[j]);
        }
      }
      split_store(v, tC + (uint32_t)(c * 32), c * 32);
      if constexpr (MODE == TC_BWD) {
        float* grow = a.d1buf + tbuf_idx(tile, H, c * 32, row);
#pragma unroll
        for (int j = 0; j < 32; ++j) grow[j * 16] = v[j];
      } else if constexpr (MODE == TC_FWD_STORE) {
        float* grow = a.h2buf + tbuf_idx(tile, H, c * 32, row);
#pragma unroll
        for (int j = 0; j < 32; ++j) grow[j * 16] = v[j];
        if (a.act2 == DFLOW_ACT_RELU) a.m2buf[((size_t)tile * nchunk + c) * 128 + row] = positive_mask(v);
      }
    }
    if constexpr (MODE != TC_BWD) bias_init(tA, biasS + 2 * H, N3p);  // D3 starts from b3 (h1.hi in region A is dead)
    publish();
    // ---- 6. D3 (region A, over the dead h1.hi) = h2 * M3^T ----
    if ((warp & 3) == 0) {
      tc_fence_after();
      if (elect_one()) {
        gemm_mixed(tA, tC, H >> 3, id3, dW16 + res_s3, w3_lo, s3_step);
        mma_commit_a(bar_mma);
      }
      __syncwarp();
    }
    wait_mma();
    // ---- 7. output: s values, coupling transform / log-det, or the cotangent of the conditioner input ----
    // (8 columns at a time, only the columns that exist: predicated-off gathers would still cost their issue slots)
    float lsum = 0.0f;
    const int nout = MODE == TC_BWD ? a.nin : a.a;
    for (int o0 = 0; o0 < nout; o0 += 8) {
      float v[8];
      {
        uint32_t r[8];
        tmem_ld8(tA + lane_off + (uint32_t)o0, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
      }
      if constexpr (MODE == TC_BWD) {
        // rows n.. go to the identity coordinates; rows 0..n-1 are the cotangent of the (normalised) conditions
        if (valid) {
          float zb[8];
          int off[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k = o0 + j;
            off[j] = (k >= n && k < a.nin) ? (int)a.id[k - n] * 128 : -1;
            zb[j] = off[j] >= 0 ? zbar_t[off[j]] : 0.0f;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (off[j] >= 0) zbar_t[off[j]] = zb[j] + v[j];
          if (a.thbar) {
            for (int j = 0; j < 8; ++j) {
              const int k = o0 + j;
              if (k < n) {
                // chain rule through normalize_input (src/Data.jl:213-218)
                const float sc = (a.flags & DFLOW_THETA_NORMALIZE) ? (a.theta_rng[k] == 0.0f ? 0.0f : 1.0f / a.theta_rng[k]) : 1.0f;
                a.thbar[sn + (size_t)k * 128] += v[j] * sc;
              }
            }
          }
        }
      } else if (a.net_id == 0) {
        float* srow = a.sbuf + ((size_t)tile * a.a16 + o0) * 128 + row;
#pragma unroll
        for (int j = 0; j < 8; ++j) srow[j * 128] = v[j];
      } else if (valid) {
        // coupling transform (src/affine/RNVP.jl:92,184; NICE: s = 0)
        float sv[8], xv[8];
        const float* srow = a.sbuf + ((size_t)tile * a.a16 + o0) * 128 + row;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int jj = o0 + j;
          sv[j] = (jj < a.a && a.has_s) ? srow[j * 128] : 0.0f;
          xv[j] = (jj < a.a) ? xin_t[(int)a.af[jj] * 128] : 0.0f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int jj = o0 + j;
          if (jj < a.a) {
            const float tv = v[j];
            xout_t[(int)a.af[jj] * 128] = a.sampling ? xv[j] * expf(sv[j]) + tv : (xv[j] - tv) * expf(-sv[j]);
            lsum += sv[j];
          }
        }
      }
    }
    if constexpr (MODE == TC_BWD) {
      // cotangent of the transformed coordinates: ubar_af = zbar_af * exp(-s) (src/affine/RNVP.jl:134)
      if (a.net_id >= 1 && a.has_s && valid)
        for (int j0 = 0; j0 < a.a; j0 += 8) {
          float sv[8], zb[8];
#pragma unroll
          for (int qq = 0; qq < 8; ++qq) {
            const int j = j0 + qq;
            sv[qq] = j < a.a ? a.sbuf[((size_t)tile * a.a16 + j) * 128 + row] : 0.0f;
            zb[qq] = j < a.a ? zbar_t[(int)a.af[j] * 128] : 0.0f;
          }
#pragma unroll
          for (int qq = 0; qq < 8; ++qq)
            if (j0 + qq < a.a) zbar_t[(int)a.af[j0 + qq] * 128] = zb[qq] * expf(-sv[qq]);
        }
    } else {
      if (a.net_id >= 1 && a.ldj && valid) a.ldj[gi] += a.sampling ? lsum : -lsum;
    }
    // (the next tile's publish() orders these TMEM reads before the chain's MMAs overwrite region A)
  }

  if (tid == 0 && !wready) mbar_wait(bars + TCS_BAR_W, 0);  // a CTA whose first chain has no tile still owns the bulk copy
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tbase, 512u);
  }
}

}  // namespace dflow
