// Weight gradients with the A operands in TMEM (included by dflow_tc.cu after tc_dw_kernel; same DwArgs, same buffers).
//
// tc_dw_kernel stages all six operand segments of a 16-sample stage through shared memory: 86 KB per stage at hidden 256,
// so its ring holds one stage pair and the CUDA-core staging (split + 172 KB of shared-memory stores per pair, each pair
// behind a fence.proxy.async that also waits for the staging thread's own global loads) runs in series with the MMAs
// (profiles/r02_c3_tc.md).  Here the three A segments (delta2, delta1, h2: the rows of this 128-row tile) never touch
// shared memory: the thread that owns hidden-unit row r loads the row's 16 samples (64 contiguous bytes), splits them and
// writes hi | lo into its TMEM lane with tcgen05.st; the MMAs read A from tensor memory.  No proxy fence on that path, so
// the next stage's loads stay in flight across the hand-off.  Only the B segments (h1, conditioner input, delta3) cross
// shared memory, in stages of (NB + K0p + a16) rows -- a 4-deep ring at hidden 256.
//
// Warps: 0-3 A stages of even index, 4-7 A stages of odd index (warp w owns TMEM lanes 32 (w % 4)..), 8-14 B stages,
// 15 allocates TMEM and issues the MMAs.  TMEM: dW2 [0, NB), dW1 [256, 288), dW3^T [288, 320), A ring 2 x 96 columns at 320.
#pragma once

namespace dflow {

constexpr int DWTS_BW = 7;        // B staging warps (8..14)
constexpr int DWTS_MAXB = 6;      // B row-blocks per staging warp and stage
constexpr uint32_t DWTS_W2 = 0, DWTS_W1 = 256, DWTS_W3 = 288, DWTS_A = 320;

__host__ __device__ inline bool dwts_shape_ok(int NB, int K0p, int a16) {
  const int K0n = (K0p + 15) & ~15;
  return NB <= 256 && K0n <= 32 && a16 <= 32 && (NB + K0p + a16) / 8 <= DWTS_MAXB * DWTS_BW;
}

__global__ void __launch_bounds__(DW_THREADS, 1) tc_dwts_kernel(const __grid_constant__ DwArgs a) {
  extern __shared__ float4 smem4[];
  float* smem = reinterpret_cast<float*>(smem4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = a.H, K0p = a.K0p, a16 = a.a16;
  const int unit = blockIdx.x % a.units, ks = blockIdx.x / a.units;
  const int nh = unit % a.nsplit, NB = a.NB;  // column half of dW2 handled here (nh > 0: dW2 only)
  const int net = a.first_net + unit / (a.mtiles * a.nsplit), mt = (unit / a.nsplit) % a.mtiles;
  const int rows_valid = min(128, H - mt * 128);
  // B segments of a stage in shared memory: h1 (NB rows) | in (K0p rows) | delta3 (a16 rows), each [hi | lo], K-major, K = 16
  const int brow[3] = {NB, nh ? 0 : K0p, nh ? 0 : a16};
  const int bdst[3] = {0, 2 * brow[0] * DW_KS, 2 * (brow[0] + brow[1]) * DW_KS};
  const int stage_fl = 2 * (brow[0] + brow[1] + brow[2]) * DW_KS;
  const int RB = (brow[0] + brow[1] + brow[2]) >> 3;
  const int NST = a.nstage;
  // B staging groups: narrow nets have few row-blocks per stage, three warps cover them -- then two groups take stage
  // pairs in turn (one group's loads fly while the other splits, stores and fences), as in tc_dw_kernel
  const int BG = (RB <= DWTS_MAXB * 3 && NST >= 4) ? 2 : 1, BGW = BG == 2 ? 3 : DWTS_BW;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)NST * stage_fl);
  uint64_t* fullB = bars;        // [NST <= 4]
  uint64_t* emptyB = bars + 4;   // [NST]
  uint64_t* fullA = bars + 8;    // [2]
  uint64_t* emptyA = bars + 10;  // [2]
  uint64_t* done = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(fullB + i, (uint32_t)BGW * 32u);
      mbar_init(emptyB + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(fullA + i, 128u);
      mbar_init(emptyA + i, 1);
    }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 15) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;

  const long long per = (a.ntiles + a.ksplit - 1) / a.ksplit;
  const long long t0 = ks * per, t1 = min(a.ntiles, t0 + per);
  const int nstages = (t1 > t0) ? (int)((t1 - t0) * (128 / DW_KS)) : 0;
  const size_t blk0 = (size_t)t0 * (128 / DW_KS);  // first [tile][16-sample block] of this CTA (tbuf_idx)

  if (warp < 8) {
    // ---- A stagers: hidden-unit row -> TMEM lane ----
    const int set = warp >> 2, q = warp & 3;
    const int row = 32 * q + lane;
    const bool rok = row < rows_valid;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int nseg = nh ? 1 : 3;
    const size_t roff = (size_t)(mt * 128 + row) * 16, bstride = (size_t)H * 16;
    const float* src[3] = {a.d2buf[net] + roff, a.d1buf[net] + roff, a.h2buf[net] + roff};
    float4 R[3][4];
    auto load = [&](int s) {
      const size_t o = (blk0 + (size_t)s) * bstride;
#pragma unroll
      for (int sg = 0; sg < 3; ++sg)
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4)
          R[sg][c4] = (rok && sg < nseg) ? __ldg(reinterpret_cast<const float4*>(src[sg] + o) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    float bsum2 = 0.0f, bsum1 = 0.0f;
    uint32_t par = 0;
    if (set < nstages) load(set);
    for (int s = set; s < nstages; s += 2) {
      mbar_wait(emptyA + set, par ^ 1u);  // the MMAs of this slot's previous stage are complete
      tc_fence_after();
#pragma unroll
      for (int sg = 0; sg < 3; ++sg) {
        if (sg < nseg) {
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const float v[4] = {R[sg][c4].x, R[sg][c4].y, R[sg][c4].z, R[sg][c4].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float h = tf32_hi(v[e]);
              hi[4 * c4 + e] = __float_as_uint(h);
              lo[4 * c4 + e] = __float_as_uint(v[e] - h);
            }
          }
          const uint32_t col = tbase + lane_off + DWTS_A + (uint32_t)(set * 96 + sg * 32);
          tmem_st16(col, hi);
          tmem_st16(col + 16u, lo);
          // bias gradients: sums over the samples of delta2 / delta1
          const float rs = ((R[sg][0].x + R[sg][0].y) + (R[sg][0].z + R[sg][0].w)) + ((R[sg][1].x + R[sg][1].y) + (R[sg][1].z + R[sg][1].w)) +
                           ((R[sg][2].x + R[sg][2].y) + (R[sg][2].z + R[sg][2].w)) + ((R[sg][3].x + R[sg][3].y) + (R[sg][3].z + R[sg][3].w));
          if (sg == 0) bsum2 += rs;
          if (sg == 1) bsum1 += rs;
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(fullA + set);
      if (s + 2 < nstages) load(s + 2);  // in flight while the MMAs of this stage (and the other set's stage) run
      if (rok && s + 4 < nstages) {
        const size_t o4 = (blk0 + (size_t)(s + 4)) * bstride;
#pragma unroll
        for (int sg = 0; sg < 3; ++sg)
          if (sg < nseg) asm volatile("prefetch.global.L2 [%0];" ::"l"(src[sg] + o4));
      }
      par ^= 1u;
    }
    if (rok && nh == 0 && nstages > 0) {
      const int hr = mt * 128 + row;
      const int bn = a.fused ? hr / a.hblk : net, hr2 = a.fused ? hr % a.hblk : hr;
      if (bn < 2) {
        if (a.p_b[bn][1] >= 0) atomicAdd(a.grad + a.p_b[bn][1] + hr2, bsum2);
        if (a.p_b[bn][0] >= 0) atomicAdd(a.grad + a.p_b[bn][0] + hr2, bsum1);
      }
    }
    // ---- flush: warps 0-3 own TMEM lanes 32w..32w+31 = hidden unit rows of this m-tile ----
    if (warp < 4 && nstages > 0) {
      mbar_wait(done, 0);
      tc_fence_after();
      const int og = mt * 128 + tid;  // hidden unit (row of delta2 / delta1 / h2) of the (possibly fused) conditioner
      const bool ok = tid < rows_valid;
      const int fn = a.fused ? og / a.hblk : net;           // net that owns this row
      const int o = a.fused ? og % a.hblk : og;             // row inside that net
      const int Hn = a.fused ? a.hblk : H;                  // width of that net
      const int i_lo = a.fused ? fn * a.hblk : 0;           // columns of dW2 that belong to it (diagonal block)
      const int j_lo = a.fused ? fn * a.ablk : 0, j_n = a.fused ? a.ablk : a16;
      float w[16];
      const int ncol = a.nsplit > 1 ? NB : Hn;  // columns of dW2 accumulated by this CTA, global column = nh * NB + i0 + j
      for (int i0 = 0; i0 < ncol; i0 += 16) {  // dW2[o][i] at p_w2 + o + Hn * i
        tmem_ld16(tbase + lane_off + DWTS_W2 + i_lo + i0, w);
        if (ok && fn < 2)
#pragma unroll
          for (int j = 0; j < 16; ++j) atomicAdd(a.grad + a.p_w[fn][1] + o + (size_t)Hn * (nh * NB + i0 + j), w[j]);
      }
      for (int k0 = 0; k0 < (nh ? 0 : K0p); k0 += 16) {  // dW1[o][k] at p_w1 + o + Hn * k
        tmem_ld16(tbase + lane_off + DWTS_W1 + k0, w);
        if (ok && fn < 2)
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (k0 + j < a.K0) atomicAdd(a.grad + a.p_w[fn][0] + o + (size_t)Hn * (k0 + j), w[j]);
      }
      for (int j0 = 0; j0 < (nh ? 0 : j_n); j0 += 16) {  // dW3[j][i=o] at p_w3 + j + a * o
        tmem_ld16(tbase + lane_off + DWTS_W3 + j_lo + j0, w);
        if (ok && fn < 2)
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j0 + j < a.a) atomicAdd(a.grad + a.p_w[fn][2] + (j0 + j) + (size_t)a.a * o, w[j]);
      }
    }
  } else if (warp < 15) {
    // ---- B stagers: global (fp32) -> hi/lo split -> shared operand layout, stage pairs behind one proxy fence ----
    const int grp = (warp - 8) / BGW, gw = (warp - 8) % BGW;
    const bool active = grp < BG;
    const float* ptr[DWTS_MAXB];
    int dst[DWTS_MAXB], lof[DWTS_MAXB], stride[DWTS_MAXB];
    uint32_t segmask = 0, biasmask = 0;
    const int lofs = (lane >> 3) * 32 + (lane & 7) * 4;
#pragma unroll
    for (int i = 0; i < DWTS_MAXB; ++i) {
      const int rb = gw + BGW * i;
      const int sg = rb < (brow[0] >> 3) ? 0 : rb < ((brow[0] + brow[1]) >> 3) ? 1 : 2;
      const int rb0 = sg == 0 ? 0 : sg == 1 ? (brow[0] >> 3) : ((brow[0] + brow[1]) >> 3);
      const int r = (rb - rb0) * 8 + (lane & 7);
      const float* base = sg == 0 ? a.h1buf[net] : sg == 1 ? a.inbuf : a.d3buf[net];
      const int srows = sg == 0 ? H : sg == 1 ? K0p : a16;  // rows of the source buffer per 16-sample block
      if (rb < RB) segmask |= 1u << i;
      if (rb < RB && sg == 2 && mt == 0 && nh == 0) biasmask |= 1u << i;  // bias gradient of the last Dense: sum of delta3
      ptr[i] = base + (size_t)((sg == 0 ? nh * NB : 0) + r) * 16 + (lane >> 3) * 4;
      stride[i] = srows * 16;
      dst[i] = (sg == 0 ? bdst[0] : sg == 1 ? bdst[1] : bdst[2]) + ((rb - rb0) * 8 >> 3) * 128 + lofs;
      lof[i] = (sg == 0 ? brow[0] : sg == 1 ? brow[1] : brow[2]) * DW_KS;
    }
    float4 vA[DWTS_MAXB], vB[DWTS_MAXB];
    float bacc[DWTS_MAXB];
#pragma unroll
    for (int i = 0; i < DWTS_MAXB; ++i) bacc[i] = 0.0f;
    auto load_stage = [&](int s, float4 (&v)[DWTS_MAXB]) {
      const size_t blk = blk0 + (size_t)s;
#pragma unroll
      for (int i = 0; i < DWTS_MAXB; ++i) {
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((segmask >> i) & 1u) v[i] = __ldg(reinterpret_cast<const float4*>(ptr[i] + blk * (size_t)stride[i]));
      }
    };
    auto prefetch_stage = [&](int s) {
      const size_t blk = blk0 + (size_t)s;
#pragma unroll
      for (int i = 0; i < DWTS_MAXB; ++i)
        if ((segmask >> i) & 1u) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr[i] + blk * (size_t)stride[i]));
    };
    auto put_stage = [&](float4 (&v)[DWTS_MAXB], uint32_t sl, uint32_t pr) {
      mbar_wait(emptyB + sl, pr ^ 1u);
      float* st = smem + (size_t)sl * stage_fl;
#pragma unroll
      for (int i = 0; i < DWTS_MAXB; ++i) {
        if ((segmask >> i) & 1u) {
          float4 hi, lo;
          hi.x = tf32_hi(v[i].x); lo.x = v[i].x - hi.x;
          hi.y = tf32_hi(v[i].y); lo.y = v[i].y - hi.y;
          hi.z = tf32_hi(v[i].z); lo.z = v[i].z - hi.z;
          hi.w = tf32_hi(v[i].w); lo.w = v[i].w - hi.w;
          *reinterpret_cast<float4*>(st + dst[i]) = hi;
          *reinterpret_cast<float4*>(st + dst[i] + lof[i]) = lo;
          if ((biasmask >> i) & 1u) bacc[i] += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
      }
    };
    // nstages is a multiple of 8 and NST is even: a pair occupies ring slots (s % NST, s % NST + 1)
    const int s_first = 2 * grp, s_step = 2 * BG;
    if (active && s_first < nstages) {
      load_stage(s_first, vA);
      load_stage(s_first + 1, vB);
    }
    uint32_t slA = (uint32_t)(s_first % NST), prA = 0;
    for (int s = s_first; active && s < nstages; s += s_step) {
      put_stage(vA, slA, prA);
      put_stage(vB, slA + 1, prA);
      fence_async_smem();
      mbar_arrive(fullB + slA);
      mbar_arrive(fullB + slA + 1);
      if (s + s_step < nstages) {
        load_stage(s + s_step, vA);
        load_stage(s + s_step + 1, vB);
      }
      if (s + 2 * s_step < nstages) {  // the pair after next: pulled into L2 now, so that its loads above find it there
        prefetch_stage(s + 2 * s_step);
        prefetch_stage(s + 2 * s_step + 1);
      }
      slA += (uint32_t)s_step;
      while (slA >= (uint32_t)NST) {
        slA -= (uint32_t)NST;
        prA ^= 1u;
      }
    }
    // bias gradient of the last Dense: reduce the four sample quads of a row, one atomic per row
#pragma unroll
    for (int i = 0; i < DWTS_MAXB; ++i) {
      float r = bacc[i];
      r += __shfl_xor_sync(0xffffffffu, r, 8);
      r += __shfl_xor_sync(0xffffffffu, r, 16);
      if (active && ((biasmask >> i) & 1u) && lane < 8 && nstages > 0) {
        const int rb = gw + BGW * i;
        const int row = (rb - ((brow[0] + brow[1]) >> 3)) * 8 + lane;
        const int bn = a.fused ? row / a.ablk : net, r2 = a.fused ? row % a.ablk : row;
        if (row < a16 && bn < 2 && r2 < a.a && a.p_b[bn][2] >= 0) atomicAdd(a.grad + a.p_b[bn][2] + r2, r);
      }
    }
  } else {
    // ---- MMA issuer: the warp runs the loop uniformly, one elected lane issues ----
    const int K0n = (K0p + 15) & ~15;  // N of the dW1 GEMM (M = 128 needs N % 16 == 0); extra rows read finite data
    const uint32_t idW2 = instr_desc_tf32(NB), idW1 = instr_desc_tf32(K0n), idW3 = instr_desc_tf32(a16);
    const uint64_t d0 = desc_at(desc_hi(DW_KS), smem_u32(smem));
    const uint32_t stage_step = ((uint32_t)stage_fl * 4u) >> 4;
    uint32_t so[3], sl[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      so[k] = ((uint32_t)bdst[k] * 4u) >> 4;
      sl[k] = ((uint32_t)(brow[k] * DW_KS) * 4u) >> 4;
    }
    const uint32_t fullB_u32 = smem_u32(fullB), emptyB_u32 = smem_u32(emptyB), fullA_u32 = smem_u32(fullA),
                   emptyA_u32 = smem_u32(emptyA);
    // D (+)= A (TMEM: hi at acol, lo at acol + 16; K = 16 samples = 2 steps) * B^T (shared memory: hi at bh, lo at bl)
    auto gemm_ts3 = [&](uint32_t td, uint32_t acol, uint64_t bh, uint64_t bl, uint32_t idesc, uint32_t acc) {
#pragma unroll
      for (int k2 = 0; k2 < DW_KS / 8; ++k2) {
        const uint64_t o = (uint64_t)(k2 * 16);
        mma_tf32_ts(td, acol + 16u + (uint32_t)k2 * 8u, bh + o, idesc, acc);
        mma_tf32_ts(td, acol + (uint32_t)k2 * 8u, bl + o, idesc, 1u);
        mma_tf32_ts(td, acol + (uint32_t)k2 * 8u, bh + o, idesc, 1u);
        acc = 1u;
      }
    };
    uint32_t slot = 0, parB = 0;
    for (int s = 0; s < nstages; ++s) {
      const uint32_t t = (uint32_t)s & 1u;
      mbar_wait_a(fullA_u32 + t * 8, ((uint32_t)s >> 1) & 1u);
      mbar_wait_a(fullB_u32 + slot * 8, parB);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t b = d0 + slot * stage_step;
        const uint32_t acc = s > 0 ? 1u : 0u;
        const uint32_t ac = tbase + DWTS_A + t * 96u;
        // dW2 += delta2 * h1^T ; dW1 += delta1 * in^T ; dW3^T += h2 * delta3^T
        gemm_ts3(tbase + DWTS_W2, ac, b + so[0], b + so[0] + sl[0], idW2, acc);
        if (nh == 0) {
          gemm_ts3(tbase + DWTS_W1, ac + 32u, b + so[1], b + so[1] + sl[1], idW1, acc);
          gemm_ts3(tbase + DWTS_W3, ac + 64u, b + so[2], b + so[2] + sl[2], idW3, acc);
        }
        mma_commit_a(emptyA_u32 + t * 8);
        mma_commit_a(emptyB_u32 + slot * 8);
        if (s == nstages - 1) mma_commit(done);
      }
      __syncwarp();
      if (++slot == (uint32_t)NST) {
        slot = 0;
        parB ^= 1u;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 15) {
    __syncwarp();
    tmem_dealloc(tbase, 512);
  }
}

}  // namespace dflow
