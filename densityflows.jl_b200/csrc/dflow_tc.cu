// Tensor-core path, generation 2 (sm_100a): warp-specialised tcgen05 / TMEM pipeline for conditioners with hidden
// width >= 32 (default for hidden > 64; hidden 64 at large batch), forward in both directions and the adjoint.
//
// One launch per (coupling layer, conditioner), persistent CTAs over 128-sample tiles.  Warps of a CTA:
//   chunk-epilogue warpgroup(s)  thread row <-> sample row of the tile <-> TMEM lane; two warpgroups alternate 32-unit
//                                chunks (448-thread CTA, one per SM), or one warpgroup in the 320-thread variant that
//                                runs two CTAs per SM for narrow nets (hidden <= 64)
//   loader / output warpgroup    builds the GEMM-1 operand of tile t+1 (gathers theta / x, or the adjoint seeds delta3)
//                                while the pipeline works on tile t; applies the coupling transform / cotangents of t
//   producer warp                streams pre-split weight blocks global -> shared with cp.async.bulk (TMA engine),
//                                completion on mbarriers; nets that fit stay resident for the life of the CTA
//   MMA warp                     runs uniformly, one elected lane issues tcgen05.mma kind::tf32 (accumulators in TMEM)
//
// Float32 parity on a TF32 tensor core: every operand is split into hi = cvt.rna.tf32(x), lo = x - hi and each
// product is issued as lo*hi + hi*lo + hi*hi (the dropped lo*lo term is ~2^-22 relative).  (Rounding lo to TF32 with
// cvt.rna as well was measured in round 2: no change of the chain-level error, which is ~10x the Float32 oracle's at
// 12-16 layers -- profiles/r02_parity.md -- and it would break hi + lo == x, which the training stores rely on.)
//
// Pipeline per tile and conditioner (hidden width H, NH = min(H, 256) output columns per pass):
//   D1 group (128 x GW)   = A1 (128 x K0) * M1_g^T        GW = 64 hidden units per group, NG-deep ring in TMEM,
//                                                          issued ahead of the epilogue
//   epilogue 1 (32 units) : tcgen05.ld, bias + relu (adjoint: relu mask), hi/lo split -> A2 slot in shared memory
//   D2 (128 x NH)        += A2 (128 x 32) * two 16-unit M2 blocks^T      (12 MMAs per issuer iteration)
//   epilogue 2 (32 units) : tcgen05.ld, bias + relu (or mask), split -> A2 slot
//   D3 (128 x N3)        += A2 * two 16-unit M3 blocks^T
//   output                : D3 + b3 -> s (kept for the t launch) or the coupling transform / log-det
// All hand-offs are mbarriers (tcgen05.commit on the MMA side, fence.proxy.async + one arrival per warp on the thread
// side).  The adjoint's input-gradient chain  delta3 -> (.W3) mask2 -> (.W2) mask1 -> (.W1)  is the same pipeline on
// the transposed matrices (dflow_tc.h); the forward sweep of a train step additionally stores the activations, their
// relu masks and the conditioner inputs; weight gradients are K = samples GEMMs in tc_dw_kernel.
// tc_cluster = 1 / 2 launch CTA pairs (weight stream shared by bulk-copy multicast / cta_group::2 with M = 256 MMAs):
// both parity-green and both slower than independent CTAs on B200 (profiles/r01_tc_summary.md), so off by default.
//
// Reference math: src/affine/RNVP.jl:77-96 (normalising), :99-147 (adjoint), :150-205 (gather, sampling);
// src/affine/NICE.jl:63-170; src/norm/Normalization.jl:64-103; src/Flows.jl:272-281,352-359.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <string.h>

#include <algorithm>
#include <new>

#include "dflow_chain_kernels.cuh"
#include "dflow_tc.cuh"
#include "dflow_tc.h"

namespace dflow {

using namespace tc;

constexpr int WKC = TC_WKC;  // hidden units per weight block (K of one block's MMAs)
constexpr int WKA = 32;      // hidden units per activation chunk (two weight blocks)
constexpr int NSMAX = TC_NSMAX;

enum { TC_FWD = 0, TC_FWD_STORE = 1, TC_BWD = 2 };

// Release builds compile the timing-experiment switches and the CTA-pair variants out (both measured, both slower or
// wrong by construction: profiles/r01_tc_summary.md); -DDFLOW_TC_EXPERIMENTS brings them back.
#ifdef DFLOW_TC_EXPERIMENTS
#define TC_DBG(a) ((a).debug)
#define TC_CLUSTER(a) ((a).cluster)
#else
#define TC_DBG(a) 0
#define TC_CLUSTER(a) 0
#endif

// TMEM column map of the net kernel (per launch, TcArgs): D1 ring of NG groups x GW columns at 0, then D3 (<= 64
// columns) and D2 (NH <= 256 columns)

// barrier slots (uint64 each) behind the ring
enum {
  BAR_W_FULL = 0,
  BAR_W_EMPTY = NSMAX,
  BAR_D1_FULL = 2 * NSMAX,       // [4] D1 group buffers
  BAR_D1_EMPTY = 2 * NSMAX + 4,  // [4]
  BAR_A2_FULL = 2 * NSMAX + 8,   // [4] A2 ring
  BAR_A2_EMPTY = 2 * NSMAX + 12, // [4]
  BAR_D2_FULL = 2 * NSMAX + 16,
  BAR_D2_EMPTY,
  BAR_D3_FULL,
  BAR_D3_EMPTY,
  BAR_A1_FULL,
  BAR_A1_EMPTY,
  BAR_W_PEER,                           // [NSMAX] CTA pair: the peer's half of a weight stage has landed
  BAR_TMEM_SLOT = BAR_W_PEER + NSMAX,
  BAR_COUNT
};

struct TcArgs {
  TcNetImg im;
  TcNetImg im2[2];  // cluster == 2: the two CTAs' half-row images
  const float* img;
  long long B;
  int net_id, has_s, d, n, a, a16, nin;
  int sampling, flags;
  int resident, NS, tmem_cols;
  int NG, NA, tm_d3, tm_d2;  // D1 group buffers, A2 ring depth, TMEM columns of D3 / D2
  int s3ps;                  // S3 blocks (W3 chunks) per ring stage when streamed
  int cluster;               // 1: CTA pairs share the weight stream (bulk-copy multicast); 2: cta_group::2 pairs (one
                             // issuer, M = 256, each CTA holds half of the weight rows)
  int debug;  // timing experiments only: bit 0 = weights loaded once per CTA (wrong results when streamed)
  unsigned char af[DMAX], id[DMAX];
  float theta_min[NMAX], theta_rng[NMAX];
  const float* x_in;
  float* x_out;
  const float* theta;
  const float* theta_const;
  float* ldj;
  float* sbuf;  // [tiles][a16][128] s values of this layer
  // training buffers, all [tiles][rows][128]
  float* inbuf;
  float* h1buf;
  float* h2buf;
  uint32_t* m1buf;  // relu masks, one word per (32-unit chunk, sample): [tiles][H/32][128]
  uint32_t* m2buf;
  float* d1buf;
  float* d2buf;
  float* d3buf;
  float* zbar;        // (d, B) cotangent of the layer output, updated in place to the cotangent of its input
  const float* zout;  // (d, B) layer output (normalising direction)
  int act1, act2;     // activations of the two hidden Dense layers (DFLOW_ACT_*; relu is the fast path)
  float inv_btot;
  const float* jbar;  // per-sample cotangent of ln_det_jac (dflow_vjp), or null: -inv_btot for every sample
  float* thbar;       // [tiles][n][128] cotangent of the conditions (dflow_vjp), or null: theta rows are dropped
  float* grad;
  int p_b3;
};

// ---- prepack: packed Flux parameters -> stage blocks (hi/lo split, core layout) ---------------------------------
__global__ void tc_prepack_kernel(const TcPackJob* jobs, const float* __restrict__ W, float* __restrict__ img) {
  const TcPackJob& J = jobs[blockIdx.x];
  const TcNetImg& im = J.im;
  float* base = img + im.off;
  const int H = im.H, K0p = im.K0p, N3p = im.N3p, NH = im.NH, nch = im.nch;
  const int part = blockIdx.y, nparts = gridDim.y;
  const int t0 = part * blockDim.x + threadIdx.x, ts = nparts * blockDim.x;
  // element (nn, k) of a matrix: plain affine map, or (fused pair) the diagonal block of the net it falls into
  auto getw = [&](int ba, int bb, int nn, int k, int sn, int sk, int rbs, int cbs, int vn, int vk) -> float {
    int net = 0;
    if (J.fuse) {
      const int rb = rbs ? nn / rbs : -1, cb = cbs ? k / cbs : -1;
      if (rb >= 0 && cb >= 0 && rb != cb) return 0.0f;
      net = rb >= 0 ? rb : cb;
      if (net > 1) return 0.0f;
      if (rbs) nn -= rb * rbs;
      if (cbs) k -= cb * cbs;
    }
    if (nn >= vn || k >= vk) return 0.0f;
    return W[(net ? bb : ba) + nn * sn + k * sk];
  };
  auto getb = [&](int pa, int pb, int i, int bs, int nvalid) -> float {
    int net = 0;
    if (J.fuse) {
      net = i / bs;
      i -= net * bs;
      if (net > 1) return 0.0f;
    }
    const int pp = net ? pb : pa;
    return (pp >= 0 && i < nvalid) ? W[pp + i] : 0.0f;
  };
  // biases
  for (int i = t0; i < H; i += ts) {
    base[im.bias_off + i] = getb(J.pb1, J.pb1b, i, J.bbs12, J.fuse ? J.bbs12 : H);
    base[im.bias_off + H + i] = getb(J.pb2, J.pb2b, i, J.bbs12, J.fuse ? J.bbs12 : H);
  }
  for (int i = t0; i < N3p; i += ts) base[im.bias_off + 2 * H + i] = getb(J.pb3, J.pb3b, i, J.bbs3, J.nb3);
  const int hv = im.halves > 1 ? im.halves : 1, rk = im.rank;
  const int gwr = im.GW / hv, nhr = NH / hv, n3r = N3p / hv;  // rows per block in this image
  // M1 [H x K0p]: group g = rows g*GW..
  for (int i = t0; i < H * K0p; i += ts) {
    const int u = i / K0p, k = i - u * K0p;
    const int g = u / im.GW, within = u - g * im.GW, sub = within / gwr, r = within - sub * gwr;
    if (sub != rk) continue;
    const float w = getw(J.base1, J.base1b, u, k, J.sn1, J.sk1, J.rbs1, J.cbs1, J.vn1, J.vk1);
    const float hi = to_tf32(w);
    float* blk = base + im.g1_off + (size_t)g * im.g1_floats;
    blk[core_idx(r, k, K0p)] = hi;
    blk[gwr * K0p + core_idx(r, k, K0p)] = w - hi;
  }
  // M2 [H x H]: (pass p, chunk c) holds rows p*NH.. (n index), columns c*WKC.. (k index)
  for (int i = t0; i < H * H; i += ts) {
    const int nn = i / H, k = i - nn * H;
    const int p = nn / NH, within = nn - p * NH, sub = within / nhr, r = within - sub * nhr;
    if (sub != rk) continue;
    const float w = getw(J.base2, J.base2b, nn, k, J.sn2, J.sk2, J.rbs2, J.cbs2, J.vn2, J.vk2);
    const float hi = to_tf32(w);
    const int c = k / WKC, kk = k - c * WKC;
    float* blk = base + im.s2_off + (size_t)(p * nch + c) * im.s2_floats;
    blk[core_idx(r, kk, WKC)] = hi;
    blk[nhr * WKC + core_idx(r, kk, WKC)] = w - hi;
  }
  // M3 [N3p x H]: chunk gc holds columns gc*WKC..
  for (int i = t0; i < N3p * H; i += ts) {
    const int nn = i / H, k = i - nn * H;
    const int sub = nn / n3r, r = nn - sub * n3r;
    if (sub != rk) continue;
    const float w = getw(J.base3, J.base3b, nn, k, J.sn3, J.sk3, J.rbs3, J.cbs3, J.vn3, J.vk3);
    const float hi = to_tf32(w);
    const int c = k / WKC, kk = k - c * WKC;
    float* blk = base + im.s3_off + (size_t)c * im.s3_floats;
    blk[core_idx(r, kk, WKC)] = hi;
    blk[n3r * WKC + core_idx(r, kk, WKC)] = w - hi;
  }
}

// ---- the net kernel ------------------------------------------------------------------------------------------
// Training buffers (h1, h2, delta1, delta2, delta3, conditioner inputs) are laid out for the weight-gradient kernel:
// [tile][block of 16 samples][row][16 samples], so that one K = 16 stage of a row range is one contiguous run.
__device__ __forceinline__ size_t tbuf_idx(long long tile, int rows, int r, int s) {
  return (((size_t)tile * 8 + (size_t)(s >> 4)) * (size_t)rows + (size_t)r) * 16 + (size_t)(s & 15);
}

// Working layout of the state (x / z / zbar) and of theta inside the tensor-core path: tile-blocked, dimension-major
// [tile][k][128 samples], so that the thread that owns a sample reads and writes coalesced 4-byte elements (a warp =
// one 128-byte line per dimension) instead of one 4d-byte row per thread.  Entry points transpose in / out.
__device__ __forceinline__ size_t tidx(long long tile, int rows, int k, int r) {
  return ((size_t)tile * (size_t)rows + (size_t)k) * 128 + (size_t)r;
}

// bit j <-> v[j] > 0 for values that are >= +0 (relu outputs): bits + 0x7fffffff carries into bit 31 iff the value is not
// zero, and a funnel shift collects that bit -- two instructions per value.  (Only read back when the activation is relu;
// other activations take their derivative from the stored values.)
__device__ __forceinline__ uint32_t positive_mask(const float (&v)[32]) {
  uint32_t m = 0;
#pragma unroll
  for (int j = 31; j >= 0; --j) m = __funnelshift_l(__float_as_uint(v[j]) + 0x7FFFFFFFu, m, 1);
  return m;
}

// ring cursor: slot index + phase parity, advanced without divisions
struct Ring {
  uint32_t slot, par, n;
  __device__ __forceinline__ void next() {
    if (++slot == n) {
      slot = 0;
      par ^= 1u;
    }
  }
};

// PAIR2 kernels contain cta_group::2 instructions and can only be launched as clusters of two CTAs.
// NWG = chunk-epilogue warpgroups: 2 (448 threads, one CTA per SM) or 1 (320 threads, two CTAs per SM for narrow nets).
template <int MODE, bool PAIR2, int NWG>
__global__ void __launch_bounds__((4 * NWG + 6) * 32, NWG == 1 ? 2 : 1) tc_net_kernel(const __grid_constant__ TcArgs a) {
  extern __shared__ float4 smem4[];
  float* smem = reinterpret_cast<float*>(smem4);
  constexpr bool pair2 = PAIR2;                             // cta_group::2: rank 0 of the pair issues for both CTAs
  const uint32_t crank = TC_CLUSTER(a) ? cluster_ctarank() : 0u;
  const TcNetImg& im = pair2 ? a.im2[crank] : a.im;         // (logical sizes are the same in all three images)
  const int hv = pair2 ? 2 : 1;                             // weight-block rows held by this CTA = logical rows / hv
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int EW = 4 * NWG;           // chunk-epilogue warps; then 4 loader/output warps, the producer, the MMA issuer
  constexpr int NTHREADS = (EW + 6) * 32;
  const int K0p = im.K0p, H = im.H, N3p = im.N3p, NH = im.NH, passes = im.passes, nch = im.nch,
            nch_pass = im.nch_pass, GW = im.GW, ng = im.ng;
  const int d = a.d, n = a.n, NG = a.NG, NA = a.NA;
  const uint32_t TM_D1 = 0, TM_D3 = (uint32_t)a.tm_d3, TM_D2 = (uint32_t)a.tm_d2;

  float* A1h = smem;
  float* A1l = A1h + 128 * K0p;
  float* A2 = A1l + 128 * K0p;  // NA slots of one 32-unit chunk: hi at A2 + s * 2*128*WKA, lo 128*WKA further
  float* biasS = A2 + NA * 2 * 128 * WKA;
  const int nbias = (2 * H + N3p + 3) & ~3;
  float* ring = biasS + nbias;
  const int ring_floats = a.resident ? im.blocks_floats : a.NS * im.slot_floats;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + ring_floats);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BAR_TMEM_SLOT);
  const float* gimg = a.img + im.off;

  if (tid == 0) {
    for (int i = 0; i < NSMAX; ++i) {
      mbar_init(bars + BAR_W_FULL + i, 1);
      mbar_init(bars + BAR_W_EMPTY + i, TC_CLUSTER(a) == 1 ? 2 : 1);  // multicast pair: released by both CTAs' MMA warps
      mbar_init(bars + BAR_W_PEER + i, 1);
    }
    const uint32_t two = pair2 ? 2u : 1u;  // cta_group::2: both CTAs' threads arrive on the leader's barriers
    for (int i = 0; i < 4; ++i) {
      mbar_init(bars + BAR_D1_FULL + i, 1);
      mbar_init(bars + BAR_D1_EMPTY + i, EW * two);  // warps
      mbar_init(bars + BAR_A2_FULL + i, 4 * two);
      mbar_init(bars + BAR_A2_EMPTY + i, 1);
    }
    mbar_init(bars + BAR_D2_FULL, 1);
    mbar_init(bars + BAR_D2_EMPTY, EW * two);
    mbar_init(bars + BAR_D3_FULL, 1);
    mbar_init(bars + BAR_D3_EMPTY, 4 * two);
    mbar_init(bars + BAR_A1_FULL, 4 * two);
    mbar_init(bars + BAR_A1_EMPTY, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == EW + 5) {
    if constexpr (pair2) tmem_alloc2(tmem_slot, (uint32_t)a.tmem_cols);
    else tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
  }
  for (int i = tid; i < nbias; i += NTHREADS) biasS[i] = i < 2 * H + N3p ? __ldg(gimg + im.bias_off + i) : 0.0f;
  tc_fence_before();
  __syncthreads();
  if (TC_CLUSTER(a)) cluster_sync_all();  // the peer's barriers exist before any multicast copy / commit can reach them
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;
  const long long ntiles = (a.B + 127) / 128;
  // every CTA walks the same number of tiles (a pair must issue identical stage sequences); tiles >= ntiles are dummies
  const long long iters = (ntiles + gridDim.x - 1) / gridDim.x;
  const uint32_t bars_u32g = smem_u32(bars);
  // arrival on a barrier the MMA issuer waits on: in a cta_group::2 pair that is always the leader's barrier
  // (one arrival per warp: __syncwarp orders the lanes' earlier shared-memory / TMEM accesses before lane 0's release)
  auto arrive_issuer = [&](int bar_index) {
    __syncwarp();
    if (lane == 0) {
      if (pair2) mbar_arrive_cluster(bars_u32g + (uint32_t)bar_index * 8u, 0u);
      else mbar_arrive(bars + bar_index);
    }
  };

  if (warp < EW) {
    // =========================== chunk-epilogue warps ===========================
    // Activation chunks are WKA = 32 hidden units wide (two 16-unit weight blocks per MMA-warp iteration: the issuer's
    // fixed cost per iteration is ~700 clk, so each iteration has to carry >= 12 full-width MMAs).  Two warpgroups
    // share the 128 sample rows (thread row = tid & 127 = TMEM lane); chunk q of the global chunk sequence belongs to
    // warpgroup q & 1 and to A2 slot q & 1.
    const int wg = warp >> 2, row = tid & 127;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const int cpg = GW / WKA, ncp = NH / WKA;
    uint32_t q = 0;      // global chunk counter (identical in both warpgroups and in the MMA warp)
    uint32_t uses = 0;   // hand-offs of this warpgroup = uses of its A2 slot
    Ring rD1{0, 0, (uint32_t)NG};
    uint32_t npass = 0;
    const uint32_t nown = NWG == 2 ? (uint32_t)NA >> 1 : (uint32_t)NA;  // A2 slots owned by this warpgroup
    // one 32-unit chunk of this thread's row -> A2 slot (hi / lo, K-major core layout with Kc = 32)
    auto handoff = [&](const float (&v)[32]) {
      const uint32_t slot = NWG == 2 ? (uint32_t)wg + 2u * (uses % nown) : uses % nown, par = (uses / nown) & 1u;
      float* a2h = A2 + slot * 2 * 128 * WKA;
      float* a2l = a2h + 128 * WKA;
      mbar_wait(bars + BAR_A2_EMPTY + slot, par ^ 1u);
      if (!(TC_DBG(a) & 16)) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 hi, lo;
          hi.x = tf32_hi(v[4 * j + 0]); lo.x = v[4 * j + 0] - hi.x;
          hi.y = tf32_hi(v[4 * j + 1]); lo.y = v[4 * j + 1] - hi.y;
          hi.z = tf32_hi(v[4 * j + 2]); lo.z = v[4 * j + 2] - hi.z;
          hi.w = tf32_hi(v[4 * j + 3]); lo.w = v[4 * j + 3] - hi.w;
          const int idx = core_idx(row, 4 * j, WKA);
          *reinterpret_cast<float4*>(a2h + idx) = hi;
          *reinterpret_cast<float4*>(a2l + idx) = lo;
        }
      }
      fence_async_smem();
      arrive_issuer(BAR_A2_FULL + (int)slot);
      ++uses;
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
    for (long long it = 0; it < iters; ++it) {
      const long long tile = blockIdx.x + it * gridDim.x;
      const bool live = tile < ntiles && !(TC_DBG(a) & (128 | 512));  // dummy tiles touch no global memory (512: timing
                                                                    // experiment without the training stores)
      for (int p = 0; p < passes; ++p) {
        // relu masks of the adjoint chain are fetched one chunk ahead (the global-load latency would otherwise sit
        // in the per-chunk chain); this warpgroup's chunks are first_c, first_c + NWG, ...
        uint32_t mnext = 0;
        if constexpr (MODE == TC_BWD) {
          const int first_c = NWG == 1 ? 0 : (int)((q ^ (uint32_t)wg) & 1u);
          if (live && first_c < (H >> 5)) mnext = a.m2buf[((size_t)tile * (H >> 5) + first_c) * 128 + row];
        }
        // ---- epilogue 1: hidden-1 units, one D1 group (GW columns) at a time ----
        for (int g = 0; g < ng; ++g) {
          mbar_wait(bars + BAR_D1_FULL + rD1.slot, rD1.par);
          tc_fence_after();
          // chunks of this group that belong to this warpgroup: with two warpgroups at most one (cpg <= 2)
          for (int cg = 0; cg < cpg; ++cg) {
            const bool mine = NWG == 1 || (((q + (uint32_t)cg) & 1u) == (uint32_t)wg);
            const bool last_ld = cg == cpg - 1;
            float v[32];
            if (mine) ld32(tbase + lane_off + TM_D1 + rD1.slot * (uint32_t)GW + (uint32_t)(cg * WKA), v);
            if (last_ld) {
              tc_fence_before();
              arrive_issuer(BAR_D1_EMPTY + (int)rD1.slot);
            }
            if (!mine) continue;
            const int c = g * cpg + cg;  // 32-unit chunk index inside the hidden layer
            if constexpr (MODE == TC_BWD) {
              const uint32_t mword = mnext;
              if (a.act2 == DFLOW_ACT_RELU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = ((mword >> j) & 1u) ? v[j] : 0.0f;
              } else {
                // non-relu hidden layer: derivative from the stored activation (output form: tanh 1 - y^2, sigmoid y (1 - y))
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  v[j] = live ? v[j] * act_grad(a.act2, a.h2buf[tbuf_idx(tile, H, c * WKA + j, row)]) : 0.0f;
              }
            } else {
              if (a.act1 == DFLOW_ACT_RELU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j] + biasS[c * WKA + j], 0.0f);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = act_apply(a.act1, v[j] + biasS[c * WKA + j]);
              }
            }
            handoff(v);
            // The training stores come AFTER the hand-off: its fence.proxy.async is a MEMBAR.ALL.CTA that waits for every
            // outstanding global store of the thread, which used to put 32 store round trips into the per-chunk chain.
            if constexpr (MODE == TC_BWD) {
              // (the next chunk's relu mask is fetched here, behind the fence, for the same reason)
              if (live && c + NWG < (H >> 5)) mnext = a.m2buf[((size_t)tile * (H >> 5) + c + NWG) * 128 + row];
              if (live) {
#pragma unroll
                for (int j = 0; j < 32; ++j) a.d2buf[tbuf_idx(tile, H, c * WKA + j, row)] = v[j];
              }
            } else if constexpr (MODE == TC_FWD_STORE) {
              if (p == 0 && live) {
#pragma unroll
                for (int j = 0; j < 32; ++j) a.h1buf[tbuf_idx(tile, H, c * WKA + j, row)] = v[j];
                a.m1buf[((size_t)tile * (H >> 5) + c) * 128 + row] = positive_mask(v);
              }
            }
          }
          q += (uint32_t)cpg;
          rD1.next();
        }
        // ---- epilogue 2: hidden-2 chunks of this pass ----
        if constexpr (MODE == TC_BWD) {
          const int first_cc = NWG == 1 ? 0 : (int)((q ^ (uint32_t)wg) & 1u);
          mnext = (live && first_cc < ncp) ? a.m1buf[((size_t)tile * (H >> 5) + p * ncp + first_cc) * 128 + row] : 0u;
        }
        mbar_wait(bars + BAR_D2_FULL, npass & 1);
        tc_fence_after();
        // this warpgroup's last chunk of the pass (its D2 reads end there); -1: it has none and releases D2 at once
        int my_last = (NWG == 1 || ((q + (uint32_t)(ncp - 1)) & 1u) == (uint32_t)wg) ? ncp - 1 : ncp - 2;
        if (my_last < 0) {
          tc_fence_before();
          arrive_issuer(BAR_D2_EMPTY);
        }
        for (int cc = 0; cc < ncp; ++cc, ++q) {
          if (NWG == 2 && (q & 1u) != (uint32_t)wg) continue;
          const int gc = p * ncp + cc;
          float v[32];
          ld32(tbase + lane_off + TM_D2 + cc * WKA, v);
          if (cc == my_last) {
            tc_fence_before();
            arrive_issuer(BAR_D2_EMPTY);
          }
          if constexpr (MODE == TC_BWD) {
            const uint32_t mword = mnext;
            if (a.act1 == DFLOW_ACT_RELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = ((mword >> j) & 1u) ? v[j] : 0.0f;
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                v[j] = live ? v[j] * act_grad(a.act1, a.h1buf[tbuf_idx(tile, H, gc * WKA + j, row)]) : 0.0f;
            }
          } else {
            if (a.act2 == DFLOW_ACT_RELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j] + biasS[H + gc * WKA + j], 0.0f);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = act_apply(a.act2, v[j] + biasS[H + gc * WKA + j]);
            }
          }
          handoff(v);
          if constexpr (MODE == TC_BWD) {  // (mask prefetch and stores after the hand-off's fence, see epilogue 1)
            if (live && cc + NWG < ncp) mnext = a.m1buf[((size_t)tile * (H >> 5) + gc + NWG) * 128 + row];
            if (live) {
#pragma unroll
              for (int j = 0; j < 32; ++j) a.d1buf[tbuf_idx(tile, H, gc * WKA + j, row)] = v[j];
            }
          } else if constexpr (MODE == TC_FWD_STORE) {
            if (live) {
#pragma unroll
              for (int j = 0; j < 32; ++j) a.h2buf[tbuf_idx(tile, H, gc * WKA + j, row)] = v[j];
              a.m2buf[((size_t)tile * (H >> 5) + gc) * 128 + row] = positive_mask(v);
            }
          }
        }
        ++npass;
      }
    }
  } else if (warp < EW + 4) {
    // =========================== loader / output warpgroup ===========================
    // Builds the GEMM-1 operand of tile t+1 (gathers from global memory) while the pipeline works on tile t, then
    // applies tile t's conditioner outputs; thread row = sample = TMEM lane.
    const int row = tid & 127;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    auto build_a1 = [&](long long it, uint32_t tc) {
      const long long tile = blockIdx.x + it * gridDim.x;
      const bool live = tile < ntiles && !(TC_DBG(a) & 128);
      const long long gi = tile * 128 + row;
      const bool valid = gi < a.B && !(TC_DBG(a) & 128);
      (void)live;
        mbar_wait(bars + BAR_A1_EMPTY, (tc & 1) ^ 1);
        // ---- A1: this sample's GEMM-1 input row ----
        if constexpr (MODE == TC_BWD) {
          // delta3 of this conditioner (src/affine/RNVP.jl:118-127): s: -zbar_af * z_af - jbar, t: -zbar_af * exp(-s)
          // (loads of a group of 8 are issued back to back before any store: one memory round trip per group)
          const float njbar = (a.jbar && valid) ? -__ldg(a.jbar + gi) : a.inv_btot;  // -jbar of this sample
          for (int k0 = 0; k0 < K0p; k0 += 8) {
            float zb[8], zo[8], sv[8], v[8];
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) {
              const int j = k0 + qq;
              zb[qq] = zo[qq] = sv[qq] = 0.0f;
              // fused pair (net_id 2): entries [0, a16) are sbar, [a16, 2 a16) are tbar
              const bool s_form = a.net_id == 0 || (a.net_id == 2 && j < a.a16);
              const int jj = (a.net_id == 2 && j >= a.a16) ? j - a.a16 : j;
              if (valid && jj < a.a) {
                const int k = a.af[jj];
                zb[qq] = a.zbar[tidx(tile, d, k, row)];
                if (s_form)
                  zo[qq] = a.zout[tidx(tile, d, k, row)];
                else if (a.has_s)
                  sv[qq] = a.sbuf[((size_t)tile * a.a16 + jj) * 128 + row];
              }
            }
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) {
              const int j = k0 + qq;
              const bool s_form = a.net_id == 0 || (a.net_id == 2 && j < a.a16);
              const int jj = (a.net_id == 2 && j >= a.a16) ? j - a.a16 : j;
              float val = 0.0f;
              if (valid && jj < a.a) val = s_form ? -zb[qq] * zo[qq] + njbar : -zb[qq] * expf(-sv[qq]);
              v[qq] = val;
            }
#pragma unroll
            for (int h4 = 0; h4 < 8; h4 += 4) {
              float4 hi, lo;
              hi.x = tf32_hi(v[h4 + 0]); lo.x = v[h4 + 0] - hi.x;
              hi.y = tf32_hi(v[h4 + 1]); lo.y = v[h4 + 1] - hi.y;
              hi.z = tf32_hi(v[h4 + 2]); lo.z = v[h4 + 2] - hi.z;
              hi.w = tf32_hi(v[h4 + 3]); lo.w = v[h4 + 3] - hi.w;
              const int idx = core_idx(row, k0 + h4, K0p);
              *reinterpret_cast<float4*>(A1h + idx) = hi;
              *reinterpret_cast<float4*>(A1l + idx) = lo;
            }
          }
        } else {
          // conditioner input row [theta_0..theta_{n-1}, x[axis_id...], 0 pad] (src/affine/RNVP.jl:157)
          for (int k0 = 0; k0 < K0p; k0 += 8) {
            float v[8];
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) {
              const int k = k0 + qq;
              float val = 0.0f;
              if (valid && k < a.nin) {
                if (k < n)
                  val = a.theta_const ? __ldg(a.theta_const + k) : __ldg(a.theta + tidx(tile, n, k, row));
                else
                  val = a.x_in[tidx(tile, d, a.id[k - n], row)];
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
#pragma unroll
            for (int h4 = 0; h4 < 8; h4 += 4) {
              float4 hi, lo;
              hi.x = tf32_hi(v[h4 + 0]); lo.x = v[h4 + 0] - hi.x;
              hi.y = tf32_hi(v[h4 + 1]); lo.y = v[h4 + 1] - hi.y;
              hi.z = tf32_hi(v[h4 + 2]); lo.z = v[h4 + 2] - hi.z;
              hi.w = tf32_hi(v[h4 + 3]); lo.w = v[h4 + 3] - hi.w;
              const int idx = core_idx(row, k0 + h4, K0p);
              *reinterpret_cast<float4*>(A1h + idx) = hi;
              *reinterpret_cast<float4*>(A1l + idx) = lo;
            }
          }
        }
        fence_async_smem();
        arrive_issuer(BAR_A1_FULL);
        // Global stores of the training modes come after the hand-off (its fence is a MEMBAR.ALL.CTA that would wait for
        // them).  The operand row is re-read from shared memory: hi + lo is the original Float32 value exactly, and the
        // MMA only reads the buffer until this warpgroup rewrites it in its next call.
        if constexpr (MODE == TC_BWD) {
          // (the bias gradient of the last Dense, sum_samples delta3, is accumulated by tc_dw_kernel while staging)
          if (live)
            for (int k0 = 0; k0 < K0p; k0 += 4) {
              const int idx = core_idx(row, k0, K0p);
              const float4 hi = *reinterpret_cast<const float4*>(A1h + idx), lo = *reinterpret_cast<const float4*>(A1l + idx);
              a.d3buf[tbuf_idx(tile, K0p, k0 + 0, row)] = hi.x + lo.x;
              a.d3buf[tbuf_idx(tile, K0p, k0 + 1, row)] = hi.y + lo.y;
              a.d3buf[tbuf_idx(tile, K0p, k0 + 2, row)] = hi.z + lo.z;
              a.d3buf[tbuf_idx(tile, K0p, k0 + 3, row)] = hi.w + lo.w;
            }
        } else {
          if constexpr (MODE == TC_FWD_STORE) {
            if (a.net_id >= 1 && live)
              for (int k0 = 0; k0 < K0p; k0 += 4) {
                const int idx = core_idx(row, k0, K0p);
                const float4 hi = *reinterpret_cast<const float4*>(A1h + idx), lo = *reinterpret_cast<const float4*>(A1l + idx);
                a.inbuf[tbuf_idx(tile, K0p, k0 + 0, row)] = hi.x + lo.x;
                a.inbuf[tbuf_idx(tile, K0p, k0 + 1, row)] = hi.y + lo.y;
                a.inbuf[tbuf_idx(tile, K0p, k0 + 2, row)] = hi.z + lo.z;
                a.inbuf[tbuf_idx(tile, K0p, k0 + 3, row)] = hi.w + lo.w;
              }
          }
          if (a.x_out != a.x_in && a.net_id >= 1 && live) {  // out-of-place (training sweep): carry the whole state
            for (int k0 = 0; k0 < d; k0 += 8) {
              float v[8];
#pragma unroll
              for (int qq = 0; qq < 8; ++qq) v[qq] = (k0 + qq < d) ? a.x_in[tidx(tile, d, k0 + qq, row)] : 0.0f;
#pragma unroll
              for (int qq = 0; qq < 8; ++qq)
                if (k0 + qq < d) a.x_out[tidx(tile, d, k0 + qq, row)] = v[qq];
            }
          }
        }
    };
    auto final_out = [&](long long it, uint32_t tc) {
      const long long tile = blockIdx.x + it * gridDim.x;
      const bool live = tile < ntiles && !(TC_DBG(a) & 128);
      const long long gi = tile * 128 + row;
      const bool valid = gi < a.B && !(TC_DBG(a) & 128);
      (void)live;
      mbar_wait(bars + BAR_D3_FULL, tc & 1);
      tc_fence_after();
      float lsum = 0.0f;
      for (int o0 = 0; o0 < N3p; o0 += 16) {
        float v[16];
        tmem_ld16(tbase + lane_off + TM_D3 + o0, v);
        if (o0 + 16 >= N3p) {
          tc_fence_before();
          arrive_issuer(BAR_D3_EMPTY);
        }
        if constexpr (MODE == TC_BWD) {
          // cotangent of the conditioner input: rows n.. go to the identity coordinates (theta rows are dropped)
          if (valid) {
            float zb[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int k = o0 + j;
              zb[j] = (k >= n && k < a.nin) ? a.zbar[tidx(tile, d, a.id[k - n], row)] : 0.0f;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int k = o0 + j;
              if (k >= n && k < a.nin) a.zbar[tidx(tile, d, a.id[k - n], row)] = zb[j] + v[j];
            }
            if (a.thbar) {  // rows 0..n-1: cotangent of the (normalised) conditions, summed over layers and conditioners
              for (int j = 0; j < 16; ++j) {
                const int k = o0 + j;
                if (k < n) {
                  // chain rule through normalize_input (src/Data.jl:213-218)
                  const float sc = (a.flags & DFLOW_THETA_NORMALIZE) ? (a.theta_rng[k] == 0.0f ? 0.0f : 1.0f / a.theta_rng[k]) : 1.0f;
                  a.thbar[tidx(tile, n, k, row)] += v[j] * sc;
                }
              }
            }
          }
        } else if (a.net_id == 0 || (a.net_id == 2 && o0 < a.a16)) {
          // s values (a fused pair delivers them in output columns [0, a16), the t values in [a16, 2 a16))
          if (live)
#pragma unroll
            for (int j = 0; j < 16; ++j) a.sbuf[((size_t)tile * a.a16 + o0 + j) * 128 + row] = v[j] + biasS[2 * H + o0 + j];
        } else if (valid) {
          // coupling transform (src/affine/RNVP.jl:92,184; NICE: s = 0)
          const int ob = a.net_id == 2 ? o0 - a.a16 : o0;
          float sv[16], xv[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int jj = ob + j;
            sv[j] = (jj < a.a && a.has_s) ? a.sbuf[((size_t)tile * a.a16 + jj) * 128 + row] : 0.0f;
            xv[j] = (jj < a.a) ? a.x_in[tidx(tile, d, a.af[jj], row)] : 0.0f;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int jj = ob + j;
            if (jj < a.a) {
              const float tv = v[j] + biasS[2 * H + o0 + j];
              a.x_out[tidx(tile, d, a.af[jj], row)] = a.sampling ? xv[j] * expf(sv[j]) + tv : (xv[j] - tv) * expf(-sv[j]);
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
              zb[qq] = j < a.a ? a.zbar[tidx(tile, d, a.af[j], row)] : 0.0f;
            }
#pragma unroll
            for (int qq = 0; qq < 8; ++qq)
              if (j0 + qq < a.a) a.zbar[tidx(tile, d, a.af[j0 + qq], row)] = zb[qq] * expf(-sv[qq]);
          }
      } else {
        if (a.net_id >= 1 && a.ldj && valid) a.ldj[gi] += a.sampling ? lsum : -lsum;
      }
    };
    if (iters > 0) build_a1(0, 0);
    for (long long it = 0; it < iters; ++it) {
      if (it + 1 < iters) build_a1(it + 1, (uint32_t)(it + 1));
      final_out(it, (uint32_t)it);
    }
  } else if (warp == EW + 4) {
    // =========================== producer ===========================
    // stage sequence per tile and pass (the MMA warp walks the same sequence):
    //   G1(0..NG-1), then per chunk c: S2(p,c) and, when c closes a group, G1(c/cpg + NG); then S3 chunks of the pass
    if (lane == 0) {
      if (a.resident) {
        // the whole block region [G1 | S2 | S3] of this conditioner stays in shared memory
        const uint32_t total = (uint32_t)im.blocks_floats * 4u;
        mbar_expect_tx(bars + BAR_W_FULL, total);
        uint32_t done = 0;
        while (done < total) {  // pieces of <= 64 KB
          const uint32_t piece = min(total - done, 65536u);
          bulk_g2s(reinterpret_cast<char*>(ring) + done, reinterpret_cast<const char*>(gimg + im.g1_off) + done, piece,
                   bars + BAR_W_FULL);
          done += piece;
        }
      } else {
        Ring rW{0, 0, (uint32_t)a.NS};
        const int cpg16 = GW / WKC;  // weight blocks per D1 group
        const uint32_t g1b = (uint32_t)im.g1_floats * 4u, s2b = (uint32_t)im.s2_floats * 4u, s3b = (uint32_t)im.s3_floats * 4u;
        auto put = [&](const float* src, uint32_t bytes) {
          if (TC_DBG(a) & 64) bytes = (bytes >> 3) & ~31u;  // timing experiment: an eighth of the weight stream
          mbar_wait(bars + BAR_W_EMPTY + rW.slot, rW.par ^ 1);
          mbar_expect_tx(bars + BAR_W_FULL + rW.slot, bytes);
          char* dst = reinterpret_cast<char*>(ring + (size_t)rW.slot * im.slot_floats);
          if (TC_CLUSTER(a) == 1) {
            // each CTA of the pair fetches one half of the block and delivers it to both
            const uint32_t half = bytes >> 1;
            bulk_g2s_multicast(dst + crank * half, reinterpret_cast<const char*>(src) + crank * half, half,
                               bars + BAR_W_FULL + rW.slot, (uint16_t)3);
          } else {
            bulk_g2s(dst, src, bytes, bars + BAR_W_FULL + rW.slot);
          }
          rW.next();
        };
        for (long long it = 0; it < iters; ++it) {
          for (int p = 0; p < passes; ++p) {
            for (int g = 0; g < NG && g < ng; ++g) put(gimg + im.g1_off + (size_t)g * im.g1_floats, g1b);
            for (int c = 0; c < nch; ++c) {
              put(gimg + im.s2_off + (size_t)(p * nch + c) * im.s2_floats, s2b);
              if ((c + 1) % cpg16 == 0) {
                const int gn = c / cpg16 + NG;
                if (gn < ng) put(gimg + im.g1_off + (size_t)gn * im.g1_floats, g1b);
              }
            }
            for (int cc = 0; cc < nch_pass; cc += a.s3ps)  // several W3 chunks per stage: one ring round trip for all of them
              put(gimg + im.s3_off + (size_t)(p * nch_pass + cc) * im.s3_floats, (uint32_t)min(a.s3ps, nch_pass - cc) * s3b);
          }
        }
      }
    }
  } else {
    // =========================== MMA issuer ===========================
    // The whole warp runs this loop uniformly; one elected lane issues.  The loop is the serial resource of the CTA
    // (one warp, dependent uniform-datapath arithmetic), so everything is precomputed in descriptor units (16 B):
    // a descriptor is `base + offset`, barriers are 32-bit shared addresses.
    const uint32_t hiK0 = desc_hi(K0p), hi16 = desc_hi(WKC);
    const uint64_t dA1h = desc_at(hiK0, smem_u32(A1h)), dA1l = desc_at(hiK0, smem_u32(A1l));
    const uint64_t dA2_0 = desc_at(desc_hi(WKA), smem_u32(A2));  // A2 slot s (128 x 32) hi: + s * a2_step, lo: + a2_lo
    const uint64_t dRing16 = desc_at(hi16, smem_u32(ring));    // weight stage read with a 16-float row pitch (S2, S3)
    const uint64_t dRingK0 = desc_at(hiK0, smem_u32(ring));    // ... with a K0p-float row pitch (G1)
    const uint32_t a2_lo = (128u * WKA * 4u) >> 4, a2_step = 2u * a2_lo;
    const int nch32 = H / WKA, ncp = NH / WKA, cpg = GW / WKA;
    uint32_t q = 0;  // global 32-unit chunk counter (A2 slot q % NA, use q / NA)
    const uint32_t na_mask = (uint32_t)NA - 1u, na_log = NA == 4 ? 2u : 1u;
    const uint32_t g1_lo = ((uint32_t)(GW / hv * K0p) * 4u) >> 4, w2_lo = ((uint32_t)(NH / hv * WKC) * 4u) >> 4,
                   w3_lo = ((uint32_t)(N3p / hv * WKC) * 4u) >> 4;  // hi -> lo distance inside a block (rows held here)
    const uint32_t slot_step = ((uint32_t)im.slot_floats * 4u) >> 4;
    const uint32_t g1_step = ((uint32_t)im.g1_floats * 4u) >> 4, s2_step = ((uint32_t)im.s2_floats * 4u) >> 4,
                   s3_step = ((uint32_t)im.s3_floats * 4u) >> 4;
    const uint32_t res_s2 = ((uint32_t)(im.s2_off - im.g1_off) * 4u) >> 4, res_s3 = ((uint32_t)(im.s3_off - im.g1_off) * 4u) >> 4;
    const uint32_t id1 = pair2 ? instr_desc_tf32_m256(GW) : instr_desc_tf32(GW),
                   id2 = pair2 ? instr_desc_tf32_m256(NH) : instr_desc_tf32(NH),
                   id3 = pair2 ? instr_desc_tf32_m256(N3p) : instr_desc_tf32(N3p);
    const int k1steps = K0p >> 3;
    const uint32_t bars_u32 = smem_u32(bars);
    const uint32_t bW_FULL = bars_u32 + BAR_W_FULL * 8, bW_EMPTY = bars_u32 + BAR_W_EMPTY * 8,
                   bD1_FULL = bars_u32 + BAR_D1_FULL * 8, bD1_EMPTY = bars_u32 + BAR_D1_EMPTY * 8,
                   bA2_FULL = bars_u32 + BAR_A2_FULL * 8, bA2_EMPTY = bars_u32 + BAR_A2_EMPTY * 8,
                   bW_PEER = bars_u32 + BAR_W_PEER * 8;
    const uint32_t tD1 = tbase + TM_D1, tD2 = tbase + TM_D2, tD3 = tbase + TM_D3;
    const bool resident = a.resident != 0, pair = TC_CLUSTER(a) == 1;
    // issue / commit variants: one CTA, or the cta_group::2 pair (commits are multicast to both CTAs' barriers)
    auto MMA = [&](uint32_t dt, uint64_t da_, uint64_t db_, uint32_t id_, uint32_t acc_) {
      if constexpr (pair2) mma_tf32_2(dt, da_, db_, id_, acc_);
      else mma_tf32(dt, da_, db_, id_, acc_);
    };
    auto COMMIT = [&](uint32_t addr) {
      if constexpr (pair2) mma_commit2_multicast_a(addr, (uint16_t)3);
      else mma_commit_a(addr);
    };
    auto release_stage = [&](uint32_t slot_) {
      if (resident) return;
      if (pair) mma_commit_multicast_a(bW_EMPTY + slot_ * 8, (uint16_t)3);
      else COMMIT(bW_EMPTY + slot_ * 8);
    };
    // wait until a ring stage has landed (cta_group::2: in both CTAs)
    auto wait_stage = [&](uint32_t slot_, uint32_t par_) {
      mbar_wait_a(bW_FULL + slot_ * 8, par_);
      if (pair2) mbar_wait_a(bW_PEER + slot_ * 8, par_);
    };
    const bool skip1 = (TC_DBG(a) & 2) != 0, skip2 = (TC_DBG(a) & 8) != 0, skip3 = (TC_DBG(a) & 4) != 0;
    Ring rW{0, 0, (uint32_t)(resident ? 1 : a.NS)}, rD1{0, 0, (uint32_t)NG};
    uint32_t npass = 0, tcount = 0;
    if (pair2 && crank != 0) {
      // ---- peer CTA of a cta_group::2 pair: no MMAs here; tell the leader when this CTA's half of a stage has landed ----
      const int cpg16 = GW / WKC;
      auto relay = [&]() {
        mbar_wait_a(bW_FULL + rW.slot * 8, rW.par);
        if (elect_one()) mbar_arrive_cluster(bW_PEER + rW.slot * 8, 0u);
        __syncwarp();
        rW.next();
      };
      for (long long it = 0; it < iters; ++it) {
        for (int p = 0; p < passes; ++p) {
          for (int g = 0; g < NG && g < ng; ++g) relay();
          for (int c = 0; c < nch; ++c) {
            relay();
            if ((c + 1) % cpg16 == 0 && c / cpg16 + NG < ng) relay();
          }
          for (int cc = 0; cc < nch_pass; cc += a.s3ps) relay();
        }
      }
    } else {
      if (resident) mbar_wait(bars + BAR_W_FULL, 0);
      for (long long it = 0; it < iters; ++it, ++tcount) {
        mbar_wait(bars + BAR_A1_FULL, tcount & 1);
        tc_fence_after();
        for (int p = 0; p < passes; ++p) {
          auto issue_d1 = [&](int g) {
            uint32_t woff;  // weight block position in descriptor units
            if (resident) {
              woff = (uint32_t)g * g1_step;
            } else {
              wait_stage(rW.slot, rW.par);
              woff = rW.slot * slot_step;
            }
            mbar_wait_a(bD1_EMPTY + rD1.slot * 8, rD1.par ^ 1);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t db = dRingK0 + woff;
              if (!skip1) {
                const uint32_t td = tD1 + rD1.slot * (uint32_t)GW;
                uint32_t acc = 0u;
                for (int ks = 0; ks < k1steps; ++ks) {
                  const uint64_t o = (uint64_t)(ks * 16);
                  MMA(td, dA1l + o, db + o, id1, acc);
                  MMA(td, dA1h + o, db + g1_lo + o, id1, 1u);
                  MMA(td, dA1h + o, db + o, id1, 1u);
                  acc = 1u;
                }
              }
              COMMIT(bD1_FULL + rD1.slot * 8);
              release_stage(rW.slot);
              if (g == ng - 1 && p == passes - 1) COMMIT(bars_u32 + BAR_A1_EMPTY * 8);
            }
            __syncwarp();
            if (!resident) rW.next();
            rD1.next();
          };
          for (int g = 0; g < NG && g < ng; ++g) issue_d1(g);
          mbar_wait(bars + BAR_D2_EMPTY, (npass & 1) ^ 1);
          tc_fence_after();
          // ---- D2 += h1 chunk (32 units) * two 16-unit M2 blocks ----
          int cg = 0, gnext = NG;
          for (int c = 0; c < nch32; ++c, ++q) {
            const uint32_t as = q & na_mask, apar = (q >> na_log) & 1u;
            const uint64_t da = dA2_0 + as * a2_step;
            // each 16-unit weight block is released as soon as its own six MMAs are queued, so the producer can refill
            // the slot while the second block of the chunk is still being multiplied
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              uint32_t woff, slot = 0;
              if (resident) {
                woff = res_s2 + (uint32_t)(p * nch + 2 * c + hf) * s2_step;
              } else {
                wait_stage(rW.slot, rW.par);
                slot = rW.slot;
                woff = slot * slot_step;
                rW.next();
              }
              if (hf == 0) mbar_wait_a(bA2_FULL + as * 8, apar);
              tc_fence_after();
              if (elect_one()) {
                const uint64_t db = dRing16 + woff;
                if (!skip2) {
                  uint32_t acc = (c > 0 || hf > 0) ? 1u : 0u;
#pragma unroll
                  for (int ks = 0; ks < 2; ++ks) {
                    const uint32_t ao = (uint32_t)(hf * 2 + ks) * 16u, bo = (uint32_t)ks * 16u;
                    MMA(tD2, da + a2_lo + ao, db + bo, id2, acc);
                    MMA(tD2, da + ao, db + w2_lo + bo, id2, 1u);
                    MMA(tD2, da + ao, db + bo, id2, 1u);
                    acc = 1u;
                  }
                }
                release_stage(slot);
                if (hf == 1) {
                  COMMIT(bA2_EMPTY + as * 8);
                  if (c == nch32 - 1) COMMIT(bars_u32 + BAR_D2_FULL * 8);
                }
              }
              __syncwarp();
            }
            if (++cg == cpg) {
              cg = 0;
              if (gnext < ng) issue_d1(gnext);
              ++gnext;
            }
          }
          // ---- D3 += h2 chunk (32 units) * two 16-unit M3 blocks ----
          uint32_t woff3 = 0;
          for (int cc = 0, ci = 0; cc < ncp; ++cc, ++q) {
            const int gc = p * ncp + cc;  // 32-unit chunk of hidden layer 2
            uint32_t woff;
            if (resident) {
              woff = res_s3 + (uint32_t)(2 * gc) * s3_step;
            } else {
              if (ci == 0) {
                wait_stage(rW.slot, rW.par);
                woff3 = rW.slot * slot_step;
              }
              woff = woff3 + (uint32_t)ci * s3_step;
            }
            const bool last_of_stage = (ci + 2 >= a.s3ps) || (cc + 1 == ncp);
            const uint32_t as = q & na_mask;
            mbar_wait_a(bA2_FULL + as * 8, (q >> na_log) & 1u);
            tc_fence_after();
            if (gc == 0) {
              mbar_wait(bars + BAR_D3_EMPTY, (tcount & 1) ^ 1);
              tc_fence_after();
            }
            if (elect_one()) {
              const uint64_t da = dA2_0 + as * a2_step;
              if (!skip3) {
                uint32_t acc = gc > 0 ? 1u : 0u;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                  const uint64_t db = dRing16 + woff + (uint32_t)hf * s3_step;
#pragma unroll
                  for (int ks = 0; ks < 2; ++ks) {
                    const uint32_t ao = (uint32_t)(hf * 2 + ks) * 16u, bo = (uint32_t)ks * 16u;
                    MMA(tD3, da + a2_lo + ao, db + bo, id3, acc);
                    MMA(tD3, da + ao, db + w3_lo + bo, id3, 1u);
                    MMA(tD3, da + ao, db + bo, id3, 1u);
                    acc = 1u;
                  }
                }
              }
              COMMIT(bA2_EMPTY + as * 8);
              if (last_of_stage) release_stage(rW.slot);
              if (gc == (H >> 5) - 1) COMMIT(bars_u32 + BAR_D3_FULL * 8);
            }
            __syncwarp();
            if (last_of_stage) {
              if (!resident) rW.next();
              ci = 0;
            } else {
              ci += 2;
            }
          }
          ++npass;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == EW + 5) {
    __syncwarp();
    if constexpr (pair2) tmem_dealloc2(tbase, (uint32_t)a.tmem_cols);
    else tmem_dealloc(tbase, (uint32_t)a.tmem_cols);
  }
  if (TC_CLUSTER(a)) cluster_sync_all();  // no CTA leaves while its peer may still signal its barriers
}

}  // namespace dflow
#include "dflow_tcs.cuh"
namespace dflow {

// ---- weight gradients: K = samples GEMMs ------------------------------------------------------------------------
// CTA = (conditioner, 128-row tile mt of the hidden units, a contiguous range of sample tiles).  Operands come from the
// [tile][row][128] buffers written by the forward / input-gradient sweeps: four consecutive samples of one row are
// one 16-byte core-matrix row of a K-major operand with K = samples.  Per stage of KS = 16 samples:
//   dW2[mt] (128 x H)    += delta2[mt] * h1^T          db2 += delta2[mt] * 1
//   dW1[mt] (128 x K0p)  += delta1[mt] * in^T          db1 += delta1[mt] * 1
//   dW3^T[mt] (128 x a16) += h2[mt] * delta3^T
// accumulated in TMEM over the CTA's whole sample range and flushed once with red.global.add.
constexpr int DW_KS = TC_DW_KS;
constexpr int DW_STAGE_WARPS = 15;  // + the MMA warp = 512 threads: 128 registers per thread
constexpr int DW_STAGE_THREADS = DW_STAGE_WARPS * 32;
constexpr int DW_THREADS = DW_STAGE_THREADS + 32;
constexpr int DW_MAXRB = 6;   // row-blocks (8 rows x 16 samples) per staging warp and stage
constexpr uint32_t DWT_W2 = 0, DWT_W1 = 256, DWT_W3 = 320;

struct DwArgs {
  int H, K0p, K0, a, a16, mtiles, units, ksplit, nstage;
  int ngroups;  // staging-warp groups that alternate stage pairs (narrow nets: one group's loads fly while another stores)
  int nsplit, NB;  // hidden 512: the H columns of dW2 are split over nsplit = 2 CTAs of NB = 256 columns each (TMEM holds 512);
                   // the second one carries only dW2 (no dW1 / dW3 / bias sums)
  long long ntiles;
  const float* h1buf[2];
  const float* h2buf[2];
  const float* d1buf[2];
  const float* d2buf[2];
  const float* d3buf[2];
  const float* inbuf;
  int first_net;
  int p_w[2][3], p_b[2][3];
  float* grad;
  // fused s + t pair: one conditioner of width H = 2 hblk whose rows [0, hblk) belong to the s net and [hblk, 2 hblk) to
  // the t net; delta3 rows [0, ablk) / [ablk, 2 ablk) likewise (a16 = 2 ablk); only the diagonal blocks are parameters
  int fused, hblk, ablk;
  int debug;  // timing experiments only (wrong results): 1 no MMAs, 2 no global loads, 4 no split / shared-memory stores
};

// Stage layout (floats), operands K-major with K = DW_KS samples, each [hi | lo]:
//   A segments (rows = RA = this m-tile's valid rows rounded to 8): delta2, delta1, h2
//   B segments: h1 (H rows), in (K0p rows), delta3 (a16 rows)
// The M = 128 MMAs read 128 rows from each A segment; rows beyond RA alias the following (finite) data and only
// feed accumulator rows that are never flushed.
__global__ void __launch_bounds__(DW_THREADS, 1) tc_dw_kernel(const __grid_constant__ DwArgs a) {
  extern __shared__ float4 smem4[];
  float* smem = reinterpret_cast<float*>(smem4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = a.H, K0p = a.K0p, a16 = a.a16;
  const int unit = blockIdx.x % a.units, ks = blockIdx.x / a.units;
  const int nh = unit % a.nsplit, NB = a.NB;  // column half of dW2 handled here
  const int net = a.first_net + unit / (a.mtiles * a.nsplit), mt = (unit / a.nsplit) % a.mtiles;
  const int rows_valid = min(128, H - mt * 128);
  const int RA = (rows_valid + 7) & ~7;
  // segments: delta2, delta1, h2 (A operands, this m-tile's rows) | h1 (this column half), in, delta3 (B operands)
  const int seg_rows[6] = {RA, nh ? 0 : RA, nh ? 0 : RA, NB, nh ? 0 : K0p, nh ? 0 : a16};
  int seg_dst[6], seg_rb0[7];
  {
    int o = 0, rb = 0;
    for (int sgi = 0; sgi < 6; ++sgi) {
      seg_dst[sgi] = o;
      seg_rb0[sgi] = rb;
      o += 2 * seg_rows[sgi] * DW_KS;
      rb += seg_rows[sgi] >> 3;
    }
    seg_rb0[6] = rb;
  }
  const int stage_fl = 2 * (seg_rows[0] + seg_rows[1] + seg_rows[2] + seg_rows[3] + seg_rows[4] + seg_rows[5]) * DW_KS;
  const int RB = seg_rb0[6];
  const int NST = a.nstage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)NST * stage_fl);
  uint64_t* full = bars;         // [NST]
  uint64_t* empty = bars + 4;    // [NST]
  uint64_t* done = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(full + i, (uint32_t)(DW_STAGE_WARPS / a.ngroups) * 32u);
      mbar_init(empty + i, 1);
    }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == DW_STAGE_WARPS) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;

  const long long per = (a.ntiles + a.ksplit - 1) / a.ksplit;
  const long long t0 = ks * per, t1 = min(a.ntiles, t0 + per);
  const int nstages = (t1 > t0) ? (int)((t1 - t0) * (128 / DW_KS)) : 0;  // (< 2^31: a CTA's share of the sample tiles)

  if (warp < DW_STAGE_WARPS) {
    // ---- staging warps: global (fp32) -> hi/lo split -> shared operand layout, loads issued one stage ahead ----
    // this warp's row-blocks: rb = warp + 8 i  (everything per row-block is resolved once, outside the stage loop)
    // Per row-block metadata is packed (bit masks + a 2-bit segment-shape code) so that two stages of loaded data fit
    // the 96-register budget next to it without spills.
    // Narrow nets need only a few row-blocks per stage: the staging warps then split into ngroups groups that take stage
    // pairs in turn (group g: pairs g, g + ngroups, ...), each warp covering more row-blocks of the pairs it handles.
    // The groups run out of phase, so one group's global loads are in flight while another splits and stores -- the
    // fence.proxy.async of a thread waits for that thread's own loads only.
    const int G = a.ngroups, GW = DW_STAGE_WARPS / G, grp = warp / GW, gw = warp % GW;
    const bool active = grp < G;
    const float* my_ptr[DW_MAXRB];
    int my_dst[DW_MAXRB];
    uint32_t okmask = 0, segmask = 0, biasmask = 0, codes = 0;  // code: 0 RA-row A segment, 1 h1, 2 in, 3 delta3
    const int lofs = (lane >> 3) * 32 + (lane & 7) * 4;
    auto seg_of = [&](int rb, int& rb0) {
      int sgi = 0;
#pragma unroll
      for (int k = 1; k < 6; ++k)
        if (rb >= seg_rb0[k]) sgi = k;
      rb0 = sgi == 0 ? seg_rb0[0] : sgi == 1 ? seg_rb0[1] : sgi == 2 ? seg_rb0[2] : sgi == 3 ? seg_rb0[3]
            : sgi == 4 ? seg_rb0[4] : seg_rb0[5];
      return sgi;
    };
#pragma unroll
    for (int i = 0; i < DW_MAXRB; ++i) {
      const int rb = gw + GW * i;
      int rb0;
      const int sgi = seg_of(rb, rb0);
      const int dst0 = sgi == 0 ? seg_dst[0] : sgi == 1 ? seg_dst[1] : sgi == 2 ? seg_dst[2] : sgi == 3 ? seg_dst[3]
                       : sgi == 4 ? seg_dst[4] : seg_dst[5];
      const int srows = sgi < 3 ? RA : sgi == 3 ? NB : sgi == 4 ? K0p : a16;         // rows of the segment in the stage
      const int valid = sgi < 3 ? rows_valid : srows;                                  // rows that exist in the source
      const float* base = sgi == 0 ? a.d2buf[net] : sgi == 1 ? a.d1buf[net] : sgi == 2 ? a.h2buf[net]
                          : sgi == 3 ? a.h1buf[net] : sgi == 4 ? a.inbuf : a.d3buf[net];
      const int r0 = (rb - rb0) * 8;
      const int row = r0 + (lane & 7);
      if (rb < RB) segmask |= 1u << i;
      if (rb < RB && row < valid) okmask |= 1u << i;
      // bias gradients: sums over the samples of delta2 / delta1 (this m-tile's rows) and delta3 (m-tile 0 only)
      if (rb < RB && nh == 0 && (sgi < 2 || (sgi == 5 && mt == 0))) biasmask |= 1u << i;
      codes |= (uint32_t)(sgi < 3 ? 0 : sgi - 2) << (2 * i);
      my_dst[i] = dst0 + (r0 >> 3) * 128 + lofs;
      my_ptr[i] = base + (size_t)((sgi < 3 ? mt * 128 : sgi == 3 ? nh * NB : 0) + row) * 16 + (lane >> 3) * 4;
    }
    // floats per [rows x 16 samples] block of the source buffer / offset of the lo half inside the stage, by code
    const int ts_by[4] = {H * 16, H * 16, K0p * 16, a16 * 16};
    const int lo_by[4] = {RA * DW_KS, NB * DW_KS, K0p * DW_KS, a16 * DW_KS};
    auto pick = [](const int (&t)[4], uint32_t c) { return c == 0 ? t[0] : c == 1 ? t[1] : c == 2 ? t[2] : t[3]; };
    // Stages are produced in PAIRS (register buffers vA / vB): both are loaded, split and stored, then ONE
    // fence.proxy.async + two barrier arrivals, then the loads of the next pair are issued.  The fence compiles to
    // MEMBAR.ALL.CTA, which also waits for the thread's outstanding global loads -- with one stage per fence every loop
    // iteration paid a full DRAM latency for 44 KB per CTA (the kernel sat at ~2.6 TB/s); a pair doubles the bytes in
    // flight per latency.  (Keeping two stages in flight ACROSS a fence does not work for the same reason.)
    float4 vA[DW_MAXRB], vB[DW_MAXRB];
    float bacc[DW_MAXRB];
#pragma unroll
    for (int i = 0; i < DW_MAXRB; ++i) bacc[i] = 0.0f;
    auto load_stage = [&](int s, float4 (&v)[DW_MAXRB]) {
      const size_t blk = (size_t)(t0 * (128 / DW_KS) + s);  // [tile][16-sample block] index (tbuf_idx)
#pragma unroll
      for (int i = 0; i < DW_MAXRB; ++i) {
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (((okmask >> i) & 1u) && !(TC_DBG(a) & 2))
          v[i] = __ldg(reinterpret_cast<const float4*>(my_ptr[i] + blk * (size_t)pick(ts_by, (codes >> (2 * i)) & 3u)));
      }
    };
    auto put_stage = [&](float4 (&v)[DW_MAXRB], uint32_t sl, uint32_t pr) {
      mbar_wait(empty + sl, pr ^ 1);
      float* st = smem + (size_t)sl * stage_fl;
#pragma unroll
      for (int i = 0; i < DW_MAXRB; ++i) {
        if (((segmask >> i) & 1u) && !(TC_DBG(a) & 4)) {
          float4 hi, lo;
          hi.x = tf32_hi(v[i].x); lo.x = v[i].x - hi.x;
          hi.y = tf32_hi(v[i].y); lo.y = v[i].y - hi.y;
          hi.z = tf32_hi(v[i].z); lo.z = v[i].z - hi.z;
          hi.w = tf32_hi(v[i].w); lo.w = v[i].w - hi.w;
          *reinterpret_cast<float4*>(st + my_dst[i]) = hi;
          *reinterpret_cast<float4*>(st + my_dst[i] + pick(lo_by, (codes >> (2 * i)) & 3u)) = lo;
          if ((biasmask >> i) & 1u) bacc[i] += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
      }
    };
    // nstages is a multiple of 128 / DW_KS = 8 and NST is even: a pair occupies ring slots (s % NST, s % NST + 1)
    const int s_first = 2 * grp, s_step = 2 * G;
    if (active && s_first < nstages) {
      load_stage(s_first, vA);
      load_stage(s_first + 1, vB);
    }
    for (int s = s_first; active && s < nstages; s += s_step) {
      const uint32_t slA = (uint32_t)(s % NST), prA = (uint32_t)((s / NST) & 1);
      put_stage(vA, slA, prA);
      put_stage(vB, slA + 1, prA);
      fence_async_smem();
      mbar_arrive(full + slA);
      mbar_arrive(full + slA + 1);
      if (s + s_step < nstages) {
        load_stage(s + s_step, vA);
        load_stage(s + s_step + 1, vB);
      }
    }
    // bias gradients: reduce the four sample quads of a row, one atomic per row
#pragma unroll
    for (int i = 0; i < DW_MAXRB; ++i) {
      float r = bacc[i];
      r += __shfl_xor_sync(0xffffffffu, r, 8);
      r += __shfl_xor_sync(0xffffffffu, r, 16);
      const int rb = gw + GW * i;
      int rb0;
      const int sgi = seg_of(rb, rb0);
      const int row = (rb - rb0) * 8 + (lane & 7);
      const bool seg_ok = active && ((segmask >> i) & 1u), ok = (okmask >> i) & 1u;
      if (seg_ok && sgi < 2 && lane < 8 && ok && nstages > 0) {
        const int hr = mt * 128 + row;
        const int bn = a.fused ? hr / a.hblk : net, hr2 = a.fused ? hr % a.hblk : hr;
        if (a.p_b[bn][sgi == 0 ? 1 : 0] >= 0) atomicAdd(a.grad + a.p_b[bn][sgi == 0 ? 1 : 0] + hr2, r);
      }
      if (seg_ok && sgi == 5 && mt == 0 && lane < 8 && nstages > 0) {
        const int bn = a.fused ? row / a.ablk : net, r2 = a.fused ? row % a.ablk : row;
        if (row < a16 && r2 < a.a && a.p_b[bn][2] >= 0) atomicAdd(a.grad + a.p_b[bn][2] + r2, r);
      }
    }
    // ---- flush: warps 0-3 own TMEM lanes 32w..32w+31 = hidden unit rows of this m-tile ----
    if (warp < 4 && nstages > 0) {
      mbar_wait(done, 0);
      tc_fence_after();
      const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
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
        tmem_ld16(tbase + lane_off + DWT_W2 + i_lo + i0, w);
        if (ok && fn < 2)
#pragma unroll
          for (int j = 0; j < 16; ++j) atomicAdd(a.grad + a.p_w[fn][1] + o + (size_t)Hn * (nh * NB + i0 + j), w[j]);
      }
      for (int k0 = 0; k0 < (nh ? 0 : K0p); k0 += 16) {  // dW1[o][k] at p_w1 + o + Hn * k
        tmem_ld16(tbase + lane_off + DWT_W1 + k0, w);
        if (ok && fn < 2)
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (k0 + j < a.K0) atomicAdd(a.grad + a.p_w[fn][0] + o + (size_t)Hn * (k0 + j), w[j]);
      }
      for (int j0 = 0; j0 < (nh ? 0 : j_n); j0 += 16) {  // dW3[j][i=o] at p_w3 + j + a * o
        tmem_ld16(tbase + lane_off + DWT_W3 + j_lo + j0, w);
        if (ok && fn < 2)
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j0 + j < a.a) atomicAdd(a.grad + a.p_w[fn][2] + (j0 + j) + (size_t)a.a * o, w[j]);
      }
    }
  } else {
    // ---- MMA issuer: the warp runs the loop uniformly, one elected lane issues ----
    const int K0n = (K0p + 15) & ~15;  // N of the dW1 GEMM (M = 128 needs N % 16 == 0); extra rows read finite data
    const uint32_t hi = desc_hi(DW_KS);
    const uint32_t idW2 = instr_desc_tf32(NB), idW1 = instr_desc_tf32(K0n), idW3 = instr_desc_tf32(a16);
    const uint64_t d0 = desc_at(hi, smem_u32(smem));
    const uint32_t stage_step = ((uint32_t)stage_fl * 4u) >> 4;
    uint32_t so[6], sl[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      so[k] = ((uint32_t)seg_dst[k] * 4u) >> 4;
      sl[k] = ((uint32_t)(seg_rows[k] * DW_KS) * 4u) >> 4;
    }
    const uint32_t full_u32 = smem_u32(full), empty_u32 = smem_u32(empty);
    uint32_t slot = 0, par = 0;
    for (int s = 0; s < nstages; ++s) {
      mbar_wait_a(full_u32 + slot * 8, par);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t b = d0 + slot * stage_step;
        const uint32_t acc = s > 0 ? 1u : 0u;
        // dW2 += delta2 * h1^T ; dW1 += delta1 * in^T ; dW3^T += h2 * delta3^T   (each: lo*hi + hi*lo + hi*hi, 2 K steps)
        if (!(TC_DBG(a) & 1)) {
          gemm3_desc(tbase + DWT_W2, b + so[0], b + so[0] + sl[0], b + so[3], b + so[3] + sl[3], DW_KS / 8, idW2, acc);
          if (nh == 0) {
            gemm3_desc(tbase + DWT_W1, b + so[1], b + so[1] + sl[1], b + so[4], b + so[4] + sl[4], DW_KS / 8, idW1, acc);
            gemm3_desc(tbase + DWT_W3, b + so[2], b + so[2] + sl[2], b + so[5], b + so[5] + sl[5], DW_KS / 8, idW3, acc);
          }
        }
        mma_commit_a(empty_u32 + slot * 8);
        if (s == nstages - 1) mma_commit(done);
      }
      __syncwarp();
      if (++slot == (uint32_t)NST) {
        slot = 0;
        par ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == DW_STAGE_WARPS) {
    __syncwarp();
    tmem_dealloc(tbase, 512);
  }
}

}  // namespace dflow
#include "dflow_tc_dwts.cuh"
namespace dflow {

// ---- small element-wise kernels ------------------------------------------------------------------------------
// All element-wise kernels below work on the tile-blocked layout (tidx): element i of a [tiles][rows][128] array.
__global__ void tc_norm_kernel(const float* x_in, float* x_out, float* ldj, long long B, int d,
                               const float* __restrict__ blk, int sampling) {
  // blk = [x_min(d) | x_max(d) | alpha, beta, ldj_const]   (src/norm/Normalization.jl:64-103)
  const float alpha = blk[2 * d], beta = blk[2 * d + 1], c = blk[2 * d + 2];
  const long long total = ((B + 127) / 128) * 128 * d;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i & 127), k = (int)((i >> 7) % d);
    const long long b = (i >> 7) / d * 128 + r;
    const float xmin = blk[k], xmax = blk[d + k], v = x_in[i];
    x_out[i] = sampling ? ((xmax - xmin) * v - alpha * xmax + beta * xmin) / (beta - alpha)
                        : (beta * (v - xmin) + alpha * (xmax - v)) / (xmax - xmin);
    if (k == 0 && ldj && b < B) ldj[b] += sampling ? c : -c;
  }
}

// cotangent through a NormalizationLayer in the normalising direction: dz/dx = (beta - alpha) / (x_max - x_min)
__global__ void tc_norm_bwd_kernel(float* zbar, long long B, int d, const float* __restrict__ blk) {
  const float alpha = blk[2 * d], beta = blk[2 * d + 1];
  const long long total = ((B + 127) / 128) * 128 * d;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)((i >> 7) % d);
    zbar[i] *= (beta - alpha) / (blk[d + k] - blk[k]);
  }
}

// log-density terms: logp_b = c0 - 0.5 |z_b|^2 + ldj_b (src/Flows.jl:279).  out: per-sample values; sum2: [sum, #non-finite];
// zbar (training): adjoint seeds z * inv_btot.
__global__ void tc_logpdf_kernel(const float* __restrict__ z, const float* __restrict__ ldj, long long B, int d, float c0,
                                 float inv_btot, float* zbar, float* out, float* sum2) {
  float ls = 0.0f, bad = 0.0f;
  for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
    const long long tile = b >> 7;
    const int r = (int)(b & 127);
    float qd = 0.0f;
    for (int k = 0; k < d; ++k) {
      const float v = z[tidx(tile, d, k, r)];
      qd = fmaf(v, v, qd);
      if (zbar) zbar[tidx(tile, d, k, r)] = v * inv_btot;
    }
    const float lp = c0 - 0.5f * qd + ldj[b];
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

// sample-major (rows, B) [optionally through an index] -> tile-blocked; padding samples are zero.  One 128-sample tile
// per CTA iteration, transposed through shared memory (row stride odd: conflict-free) so that both sides are coalesced.
__global__ void __launch_bounds__(256) tc_gather_t_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx,
                                                          long long first, long long B, int rows, float* dst) {
  extern __shared__ float tsm[];
  const int ld = rows | 1;
  const long long ntiles = (B + 127) / 128;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    for (int i = threadIdx.x; i < 128 * rows; i += 256) {
      const int r = i / rows, k = i - r * rows;
      const long long b = tile * 128 + r;
      float v = 0.0f;
      if (b < B) {
        const long long col = idx ? (long long)idx[first + b] : first + b;
        v = src[col * rows + k];
      }
      tsm[r * ld + k] = v;
    }
    __syncthreads();
    float* out = dst + (size_t)tile * rows * 128;
    for (int i = threadIdx.x; i < 128 * rows; i += 256) out[i] = tsm[(i & 127) * ld + (i >> 7)];
    __syncthreads();
  }
}

// tile-blocked -> sample-major (rows, B)
__global__ void __launch_bounds__(256) tc_scatter_t_kernel(const float* __restrict__ src, long long B, int rows, float* dst) {
  extern __shared__ float tsm[];
  const int ld = rows | 1;
  const long long ntiles = (B + 127) / 128;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const float* in = src + (size_t)tile * rows * 128;
    for (int i = threadIdx.x; i < 128 * rows; i += 256) tsm[(i & 127) * ld + (i >> 7)] = in[i];
    __syncthreads();
    for (int i = threadIdx.x; i < 128 * rows; i += 256) {
      const int r = i / rows, k = i - r * rows;
      const long long b = tile * 128 + r;
      if (b < B) dst[b * rows + k] = tsm[r * ld + k];
    }
    __syncthreads();
  }
}

// tensor-product grid points in the tile-blocked layout (dflow_logpdf_grid): x_k(b) = vals[off_k + (b / stride_k) % len_k]
__global__ void tc_grid_kernel(float* x, long long B, int d, const float* __restrict__ vals, const long long* __restrict__ meta) {
  const long long total = ((B + 127) / 128) * 128;
  for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < total; b += (long long)gridDim.x * blockDim.x) {
    const long long tile = b >> 7;
    const int r = (int)(b & 127);
    for (int k = 0; k < d; ++k)
      x[tidx(tile, d, k, r)] = b < B ? __ldg(vals + meta[3 * k + 2] + (b / meta[3 * k + 1]) % meta[3 * k]) : 0.0f;
  }
}

// base draw z ~ N(0, I) in the tile-blocked layout: Philox4x32-10 + Box-Muller, counter = global sample index
__global__ void tc_philox_kernel(float* z, long long B, int d, unsigned long long seed, unsigned int offset,
                                 unsigned long long first) {
  for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
    const unsigned long long ctr = first + (unsigned long long)b;
    const long long tile = b >> 7;
    const int r = (int)(b & 127);
    for (int g = 0; g < (d + 3) / 4; ++g) {
      unsigned int rr[4];
      philox4x32_10((unsigned int)ctr, (unsigned int)(ctr >> 32), (unsigned int)g, offset, (unsigned int)seed,
                    (unsigned int)(seed >> 32), rr);
      float v[4];
      box_muller(rr[0], rr[1], v[0], v[1]);
      box_muller(rr[2], rr[3], v[2], v[3]);
      for (int qq = 0; qq < 4; ++qq)
        if (4 * g + qq < d) z[tidx(tile, d, 4 * g + qq, r)] = v[qq];
    }
  }
}

// ---- host side -----------------------------------------------------------------------------------------------
#define CKT(call)                                                                          \
  do {                                                                                     \
    cudaError_t _e = (call);                                                               \
    if (_e != cudaSuccess) {                                                               \
      set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DFLOW_E_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

static void fill_img(TcNetImg& im, int K0, int H, int N3, long long& off, int k0_align = 8, int halves = 1,
                     int rank = 0) {
  im.off = off;
  im.K0 = K0;
  im.K0p = (K0 + k0_align - 1) & ~(k0_align - 1);
  im.H = H;
  im.N3 = N3;
  im.N3p = (N3 + 15) & ~15;
  im.NH = H < 256 ? H : 256;
  im.passes = H / im.NH;
  im.nch = H / WKC;
  im.nch_pass = im.NH / WKC;
  im.GW = (H % 64 == 0) ? 64 : 32;
  im.ng = H / im.GW;
  im.halves = halves;
  im.rank = rank;
  im.g1_floats = 2 * (im.GW / halves) * im.K0p;  // a CTA of a cta_group::2 pair holds half of every block's rows
  im.s2_floats = 2 * (im.NH / halves) * WKC;
  im.s3_floats = 2 * (im.N3p / halves) * WKC;
  im.slot_floats = std::max(im.g1_floats, std::max(im.s2_floats, im.s3_floats));
  int o = 0;
  im.bias_off = o;
  o += (2 * H + im.N3p + 3) & ~3;
  im.g1_off = o;
  o += im.ng * im.g1_floats;
  im.s2_off = o;
  o += im.passes * im.nch * im.s2_floats;
  im.s3_off = o;
  o += im.nch * im.s3_floats;
  im.blocks_floats = o - im.g1_off;
  im.total = o;
  off += o;
}

struct TcLaunchCfg {
  size_t smem;
  int resident, NS, NA, NG, tm_d3, tm_d2, tmem_cols, ctas_per_sm;
  int nwg;  // chunk-epilogue warpgroups: 2 (448 threads) or 1 (320 threads, two CTAs per SM)
};

static bool tc_launch_cfg(const dflow_chain* c, const TcNetImg& im, TcLaunchCfg& cfg, bool allow_resident = true) {
  const size_t bar_bytes = (size_t)BAR_COUNT * 8 + 16;
  const size_t cap = (size_t)c->max_smem_optin;
  auto base_bytes = [&](int na) {  // na = A2 slots of one 32-unit chunk (2 or 4)
    return (size_t)(2 * 128 * im.K0p + na * 2 * 128 * WKA + ((2 * im.H + im.N3p + 3) & ~3)) * 4;
  };
  // TMEM: compact map (256 columns, two CTAs per SM) for narrow nets, full map otherwise
  if (im.NH <= 64) {
    cfg.NG = 2;
    cfg.tm_d3 = 128;
    cfg.tm_d2 = 192;
    cfg.tmem_cols = 256;
  } else {
    cfg.NG = 3;
    cfg.tm_d3 = 192;
    cfg.tm_d2 = 256;
    cfg.tmem_cols = 512;
  }
  const size_t res = (size_t)im.blocks_floats * 4;
  cfg.resident = 0;
  cfg.nwg = 2;
  // narrow nets (hidden <= 64): the pipeline is latency bound, so run two small CTAs per SM (one chunk-epilogue
  // warpgroup each, 256 TMEM columns each) if the shared memory of both fits
  if (im.NH <= 64 && im.halves <= 1 && allow_resident) {
    const size_t half_cap = 233472 / 2 - 1024;
    const size_t b = base_bytes(2) + bar_bytes;
    if (im.passes == 1 && b + res <= half_cap) {
      cfg.resident = 1;
      cfg.NA = 2;
      cfg.NS = 1;
      cfg.smem = b + res;
      cfg.nwg = 1;
    } else if (b + 3 * (size_t)im.slot_floats * 4 <= half_cap) {
      int ns = (int)((half_cap - b) / ((size_t)im.slot_floats * 4));
      if (ns > 6) ns = 6;
      cfg.NA = 2;
      cfg.NS = ns;
      cfg.smem = b + (size_t)ns * im.slot_floats * 4;
      cfg.nwg = 1;
    }
    if (cfg.nwg == 1) {
      cfg.ctas_per_sm = 2;
      if (cfg.smem < 76900) cfg.smem = 76900;  // never three CTAs per SM: TMEM serves two (256 columns each)
      return true;
    }
  }
  for (int na = 2; na >= 2 && !cfg.resident; na -= 2) {
    if (allow_resident && im.passes == 1 && res < (1u << 20) && base_bytes(na) + res + bar_bytes <= cap) {
      cfg.resident = 1;
      cfg.NA = na;
      cfg.NS = 1;
      cfg.smem = base_bytes(na) + res + bar_bytes;
    }
  }
  if (!cfg.resident) {
    bool ok = false;
    // a cta_group::2 pair (half-size weight slots) can afford four A2 slots, which absorb the cross-CTA latencies
    for (int na = (im.halves == 2 ? 4 : 2); na >= 2 && !ok; na -= 2) {
      const size_t b = base_bytes(na) + bar_bytes;
      const size_t room = cap > b ? cap - b : 0;
      int ns = (int)(room / ((size_t)im.slot_floats * 4));
      if (ns > 8) ns = 8;
      if (ns >= (na == 4 ? 4 : 3)) {
        ok = true;
        cfg.NA = na;
        cfg.NS = ns;
        cfg.smem = b + (size_t)ns * im.slot_floats * 4;
      }
    }
    if (!ok) return false;
  }
  // co-residency: limited by shared memory (228 KB per SM, 1 KB reserved per CTA) and by TMEM columns
  const int by_tmem = 512 / cfg.tmem_cols;
  int by_smem = (int)(233472 / (cfg.smem + 1024));
  if (by_smem < 1) by_smem = 1;
  if (by_smem > by_tmem) {
    // pad the request so that no more CTAs than TMEM can serve become resident (they would spin in tcgen05.alloc)
    const size_t want = 233472 / (size_t)(by_tmem + 1) + 1;
    if (want > cfg.smem + 1024 && want - 1024 <= cap) cfg.smem = want - 1024;
    by_smem = by_tmem;
  }
  cfg.ctas_per_sm = std::min(std::min(by_smem, by_tmem), 2);
  return true;
}

int tc_build_plan(dflow_chain* c) {
  const DevChain* C = c->hc();
  const DevChainHdr& Hd = C->h;
  TcPlan* tp = new (std::nothrow) TcPlan();
  if (!tp) return DFLOW_E_NOMEM;
  c->tcp = tp;
  long long off = 0;
  tp->train_ok = true;
  for (int ei = 0; ei < Hd.L; ++ei) {
    const DevElem& E = C->e[ei];
    TcLayer Ld;
    memset(&Ld, 0, sizeof(Ld));
    if (E.kind == DFLOW_ELEM_NORM) {
      Ld.is_coupling = 0;
      Ld.norm_off = E.stage_off;
      tp->layers.push_back(Ld);
      continue;
    }
    Ld.is_coupling = 1;
    Ld.has_s = (E.kind == DFLOW_ELEM_RNVP) ? 1 : 0;
    Ld.h = E.t.w[1];
    Ld.nin = E.nin;
    Ld.a = E.a;
    Ld.a16 = (E.a + 15) & ~15;
    memcpy(Ld.af, E.af, sizeof(Ld.af));
    memcpy(Ld.id, E.id, sizeof(Ld.id));
    const int h = Ld.h, a = Ld.a;
    tp->hmax = std::max(tp->hmax, h);
    tp->a16max = std::max(tp->a16max, Ld.a16);
    for (int ni = (Ld.has_s ? 0 : 1); ni < 2; ++ni) {
      const DevNet& net = ni == 0 ? E.s : E.t;
      for (int j = 0; j < 3; ++j) {
        Ld.p_w[ni][j] = net.p_w[j];
        Ld.p_b[ni][j] = net.p_b[j];
      }
      Ld.act[ni][0] = net.act[0];
      Ld.act[ni][1] = net.act[1];
      // forward orientation: M1 = W1 (h x nin), M2 = W2, M3 = W3 (a x h); Flux (out,in) column-major W[o + out*i]
      fill_img(Ld.fwd[ni], Ld.nin, h, a, off);
      TcPackJob J;
      memset(&J, 0, sizeof(J));
      J.im = Ld.fwd[ni];
      J.base1 = net.p_w[0]; J.sn1 = 1; J.sk1 = h; J.vk1 = Ld.nin; J.vn1 = h;
      J.base2 = net.p_w[1]; J.sn2 = 1; J.sk2 = h; J.vn2 = h; J.vk2 = h;
      J.base3 = net.p_w[2]; J.sn3 = 1; J.sk3 = a; J.vn3 = a; J.vk3 = h;
      J.pb1 = net.p_b[0]; J.pb2 = net.p_b[1]; J.pb3 = net.p_b[2]; J.nb3 = a;
      tp->jobs_fwd.push_back(J);
      tp->k0pmax = std::max(tp->k0pmax, Ld.fwd[ni].K0p);
#ifdef DFLOW_TC_EXPERIMENTS
      for (int r = 0; r < 2; ++r) {  // half-row images of the two CTAs of a cta_group::2 pair
        fill_img(Ld.fwd2[ni][r], Ld.nin, h, a, off, 8, 2, r);
        J.im = Ld.fwd2[ni][r];
        tp->jobs_fwd.push_back(J);
      }
#endif
      {
        // adjoint orientation: M1[u][o] = W3[o][u], M2[i][o] = W2[o][i], M3[k][u] = W1[u][k]
        fill_img(Ld.bwd[ni], a, h, Ld.nin, off, 16);  // K0p = a16: delta3 rows double as the dW3 operand
        memset(&J, 0, sizeof(J));
        J.im = Ld.bwd[ni];
        J.base1 = net.p_w[2]; J.sn1 = a; J.sk1 = 1; J.vk1 = a; J.vn1 = h;
        J.base2 = net.p_w[1]; J.sn2 = h; J.sk2 = 1; J.vn2 = h; J.vk2 = h;
        J.base3 = net.p_w[0]; J.sn3 = h; J.sk3 = 1; J.vn3 = Ld.nin; J.vk3 = h;
        J.pb1 = J.pb2 = J.pb3 = -1;
        tp->jobs_bwd.push_back(J);
#ifdef DFLOW_TC_EXPERIMENTS
        for (int r = 0; r < 2; ++r) {
          fill_img(Ld.bwd2[ni][r], a, h, Ld.nin, off, 16, 2, r);
          J.im = Ld.bwd2[ni][r];
          tp->jobs_bwd.push_back(J);
        }
#endif
      }
    }
    tp->hu = std::max(tp->hu, h);
    // fused s + t pair: same input, same widths, 2h within one 256-column pass
    if (Ld.has_s && E.s.w[1] == h && E.s.w[2] == h && 2 * h <= 256 && E.s.act[0] == E.t.act[0] && E.s.act[1] == E.t.act[1] &&
        E.s.has_bias == E.t.has_bias) {
      const int a16 = Ld.a16;
      TcPackJob J;
      // forward: M1 = [W1_s ; W1_t] (2h x nin), M2 = diag(W2_s, W2_t), M3 = diag(W3_s, W3_t) with the t rows at a16
      fill_img(Ld.ffwd, Ld.nin, 2 * h, a16 + a, off);
      memset(&J, 0, sizeof(J));
      J.im = Ld.ffwd;
      J.fuse = 1;
      J.base1 = E.s.p_w[0]; J.base1b = E.t.p_w[0]; J.sn1 = 1; J.sk1 = h; J.rbs1 = h; J.cbs1 = 0; J.vn1 = h; J.vk1 = Ld.nin;
      J.base2 = E.s.p_w[1]; J.base2b = E.t.p_w[1]; J.sn2 = 1; J.sk2 = h; J.rbs2 = h; J.cbs2 = h; J.vn2 = h; J.vk2 = h;
      J.base3 = E.s.p_w[2]; J.base3b = E.t.p_w[2]; J.sn3 = 1; J.sk3 = a; J.rbs3 = a16; J.cbs3 = h; J.vn3 = a; J.vk3 = h;
      J.pb1 = E.s.p_b[0]; J.pb1b = E.t.p_b[0]; J.pb2 = E.s.p_b[1]; J.pb2b = E.t.p_b[1];
      J.pb3 = E.s.p_b[2]; J.pb3b = E.t.p_b[2]; J.nb3 = a; J.bbs12 = h; J.bbs3 = a16;
      tp->jobs_fwd.push_back(J);
      // adjoint: input [sbar (a16) | tbar (a16)], M1 = diag(W3_s^T, W3_t^T), M2 = diag(W2^T), M3 = [W1_s^T | W1_t^T]
      fill_img(Ld.fbwd, a16 + a, 2 * h, Ld.nin, off, 16);
      memset(&J, 0, sizeof(J));
      J.im = Ld.fbwd;
      J.fuse = 1;
      J.base1 = E.s.p_w[2]; J.base1b = E.t.p_w[2]; J.sn1 = a; J.sk1 = 1; J.rbs1 = h; J.cbs1 = a16; J.vn1 = h; J.vk1 = a;
      J.base2 = E.s.p_w[1]; J.base2b = E.t.p_w[1]; J.sn2 = h; J.sk2 = 1; J.rbs2 = h; J.cbs2 = h; J.vn2 = h; J.vk2 = h;
      J.base3 = E.s.p_w[0]; J.base3b = E.t.p_w[0]; J.sn3 = h; J.sk3 = 1; J.rbs3 = 0; J.cbs3 = h; J.vn3 = Ld.nin; J.vk3 = h;
      J.pb1 = J.pb2 = J.pb3 = J.pb1b = J.pb2b = J.pb3b = -1;
      J.bbs12 = h; J.bbs3 = 16;
      tp->jobs_bwd.push_back(J);
      TcLaunchCfg fc;
      Ld.fused = (tc_launch_cfg(c, Ld.ffwd, fc) && tc_launch_cfg(c, Ld.fbwd, fc)) ? 1 : 0;
      // the pair's weight-gradient launch (width 2h, 2 a16 delta3 rows) must fit one of the two kernels as well
      {
        const int H2 = 2 * h, RAf = std::min(128, H2), K0pf = Ld.ffwd.K0p, a16f = 2 * a16;
        const bool ts_ok = dwts_shape_ok(H2, K0pf, a16f);
        const int rows = 3 * RAf + H2 + K0pf + a16f;
        const size_t sb = 2 * (size_t)rows * DW_KS * 4, ov = (size_t)(128 - RAf) * DW_KS * 4 * 2;
        const bool smem_ok = rows / 8 <= DW_MAXRB * DW_STAGE_WARPS && (size_t)c->max_smem_optin >= 2 * sb + ov + 128;
        if (!ts_ok && !smem_ok) Ld.fused = 0;
      }
    }
    TcLaunchCfg cfg;
    Ld.tcs = 1;
    for (int ni = (Ld.has_s ? 0 : 1); ni < 2; ++ni)
      if (!tcs_image_ok(Ld.fwd[ni]) || !tcs_image_ok(Ld.bwd[ni]) || tcs_smem_bytes(Ld.fwd[ni]) > (size_t)c->max_smem_optin ||
          tcs_smem_bytes(Ld.bwd[ni]) > (size_t)c->max_smem_optin)
        Ld.tcs = 0;
    for (int ni = (Ld.has_s ? 0 : 1); ni < 2; ++ni) {
      if (!tc_launch_cfg(c, Ld.fwd[ni], cfg) || !tc_launch_cfg(c, Ld.bwd[ni], cfg)) {
        set_error("element %d: conditioner does not fit the tensor-core pipeline's shared memory", ei);
        return DFLOW_E_UNSUPPORTED;
      }
    }
    tp->layers.push_back(Ld);
  }
  tp->img_floats = (size_t)std::max<long long>(off, 4);
  if (cudaMalloc(&tp->d_img, tp->img_floats * sizeof(float)) != cudaSuccess) {
    set_error("cudaMalloc failed for the tensor-core weight image (%zu floats)", tp->img_floats);
    return DFLOW_E_NOMEM;
  }
  const size_t nf = tp->jobs_fwd.size(), nb = tp->jobs_bwd.size();
  if (cudaMalloc(&tp->d_jobs_fwd, std::max<size_t>(nf, 1) * sizeof(TcPackJob)) != cudaSuccess ||
      cudaMalloc(&tp->d_jobs_bwd, std::max<size_t>(nb, 1) * sizeof(TcPackJob)) != cudaSuccess) {
    set_error("cudaMalloc failed for the prepack job table");
    return DFLOW_E_NOMEM;
  }
  if ((nf && cudaMemcpy(tp->d_jobs_fwd, tp->jobs_fwd.data(), nf * sizeof(TcPackJob), cudaMemcpyHostToDevice) != cudaSuccess) ||
      (nb && cudaMemcpy(tp->d_jobs_bwd, tp->jobs_bwd.data(), nb * sizeof(TcPackJob), cudaMemcpyHostToDevice) != cudaSuccess)) {
    set_error("cudaMemcpy failed: %s", cudaGetErrorString(cudaGetLastError()));
    return DFLOW_E_CUDA;
  }
  return DFLOW_OK;
}

void tc_free_plan(dflow_chain* c) {
  TcPlan* tp = c->tcp;
  if (!tp) return;
  if (tp->d_img) cudaFree(tp->d_img);
  if (tp->d_jobs_fwd) cudaFree(tp->d_jobs_fwd);
  if (tp->d_jobs_bwd) cudaFree(tp->d_jobs_bwd);
  delete tp;
  c->tcp = nullptr;
}

int tc_prepack(dflow_chain* c, const float* W, bool with_bwd, cudaStream_t st) {
  TcPlan* tp = c->tcp;
  if (!tp->jobs_fwd.empty()) {
    tc_prepack_kernel<<<dim3((unsigned)tp->jobs_fwd.size(), 8), 256, 0, st>>>(tp->d_jobs_fwd, W, tp->d_img);
    CKT(cudaGetLastError());
    c->launches++;
  }
  if (with_bwd && !tp->jobs_bwd.empty()) {
    tc_prepack_kernel<<<dim3((unsigned)tp->jobs_bwd.size(), 8), 256, 0, st>>>(tp->d_jobs_bwd, W, tp->d_img);
    CKT(cudaGetLastError());
    c->launches++;
  }
  return DFLOW_OK;
}

static void fill_common(const dflow_chain* c, const TcLayer& Ld, TcArgs& a) {
  const DevChainHdr& Hd = c->hc()->h;
  a.img = c->tcp->d_img;
  a.has_s = Ld.has_s;
  a.d = Hd.d;
  a.n = Hd.n;
  a.a = Ld.a;
  a.a16 = Ld.a16;
  a.nin = Ld.nin;
  memcpy(a.af, Ld.af, sizeof(a.af));
  memcpy(a.id, Ld.id, sizeof(a.id));
  for (int k = 0; k < NMAX; ++k) {
    a.theta_min[k] = Hd.theta_min[k];
    a.theta_rng[k] = Hd.theta_rng[k];
  }
}

// narrow conditioners (hidden 32 / 64): the TMEM-sourced kernel of dflow_tcs.cuh, unless tc_ts = -1
static inline bool tcs_layer(const dflow_chain* c, const TcLayer& Ld) { return Ld.tcs && c->tc_ts >= 0; }

template <int MODE>
static int launch_tcs(dflow_chain* c, TcArgs& a, cudaStream_t st) {
  const size_t smem = tcs_smem_bytes(a.im);  // one 512-thread CTA per SM (it owns the SM's 512 TMEM columns)
  if (smem > (size_t)c->max_smem_optin) {
    set_error("conditioner does not fit the narrow tensor-core kernel's shared memory");
    return DFLOW_E_UNSUPPORTED;
  }
  a.tmem_cols = 512;
  void (*kern)(TcArgs) = tcs_net_kernel<MODE>;
  CKT(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long grid = ((a.B + 127) / 128 + TCS_CHAINS - 1) / TCS_CHAINS;
  if (grid > c->sm_count) grid = c->sm_count;
  // never two of these CTAs on one SM (the second would spin in tcgen05.alloc): request more than half of the shared memory
  const size_t smem_req = std::max(smem, (size_t)(233472 / 2 + 1 - 1024));
  CKT(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_req));
  kern<<<(unsigned)grid, TCS_THREADS, smem_req, st>>>(a);
  CKT(cudaGetLastError());
  c->launches++;
  return DFLOW_OK;
}

template <int MODE>
static int launch_net(dflow_chain* c, TcArgs& a, cudaStream_t st, const TcNetImg* half = nullptr) {
  if (c->tc_ts >= 0 && tcs_image_ok(a.im) && tcs_smem_bytes(a.im) <= (size_t)c->max_smem_optin) return launch_tcs<MODE>(c, a, st);
  TcLaunchCfg cfg;
  if (!tc_launch_cfg(c, a.im, cfg)) {
    set_error("conditioner does not fit the tensor-core pipeline's shared memory");
    return DFLOW_E_UNSUPPORTED;
  }
  long long grid = (a.B + 127) / 128;
  // streamed weights, CTA pairs: tc_cluster = 2 -> cta_group::2 (one issuer for two SMs, half of the weight rows per
  // CTA), tc_cluster = 1 -> independent CTAs that share the weight stream by bulk-copy multicast
  a.cluster = 0;
#ifdef DFLOW_TC_EXPERIMENTS
  if (!cfg.resident && grid >= 2 && c->tc_debug_cluster == 2 && half) {
    TcLaunchCfg c2;
    if (tc_launch_cfg(c, half[0], c2, false)) {
      cfg = c2;
      a.cluster = 2;
      a.im2[0] = half[0];
      a.im2[1] = half[1];
    }
  }
  if (!a.cluster && !cfg.resident && grid >= 2 && c->tc_debug_cluster == 1) a.cluster = 1;
#else
  (void)half;
#endif
  a.resident = cfg.resident;
  a.NS = cfg.NS;
  a.NA = cfg.NA;
  a.NG = cfg.NG;
  a.tm_d3 = cfg.tm_d3;
  a.tm_d2 = cfg.tm_d2;
  const TcNetImg& imu = a.cluster == 2 ? a.im2[0] : a.im;
  a.s3ps = std::max(2, (imu.slot_floats / imu.s3_floats) & ~1);  // an even number of 16-unit blocks
  a.debug = c->tc_debug;
  a.tmem_cols = cfg.tmem_cols;
#ifdef DFLOW_TC_EXPERIMENTS
  void (*kern)(TcArgs) = a.cluster == 2 ? tc_net_kernel<MODE, true, 2>
                         : cfg.nwg == 1 ? tc_net_kernel<MODE, false, 1>
                                        : tc_net_kernel<MODE, false, 2>;
#else
  void (*kern)(TcArgs) = cfg.nwg == 1 ? tc_net_kernel<MODE, false, 1> : tc_net_kernel<MODE, false, 2>;
#endif
  const unsigned nthreads = (unsigned)(4 * cfg.nwg + 6) * 32u;
  CKT(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem));
  const long long cap = (long long)c->sm_count * cfg.ctas_per_sm;
  if (grid > cap) grid = cap;
  if (a.cluster) {
    grid &= ~1LL;
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof(lc));
    lc.gridDim = dim3((unsigned)grid);
    lc.blockDim = dim3(nthreads);
    lc.dynamicSmemBytes = cfg.smem;
    lc.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    lc.attrs = attr;
    lc.numAttrs = 1;
    CKT(cudaLaunchKernelEx(&lc, kern, a));
  } else {
    kern<<<(unsigned)grid, nthreads, cfg.smem, st>>>(a);
  }
  CKT(cudaGetLastError());
  c->launches++;
  return DFLOW_OK;
}

static unsigned ew_blocks(long long work) {
  long long blocks = (work + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

static size_t tsm_bytes(int rows) { return (size_t)128 * (rows | 1) * sizeof(float); }
static unsigned tile_blocks(long long B) { return (unsigned)std::max<long long>(1, std::min<long long>((B + 127) / 128, 148 * 8)); }

static int gather_t(dflow_chain* c, const float* src, const int32_t* idx, long long first, long long B, int rows, float* dst,
                    cudaStream_t st) {
  tc_gather_t_kernel<<<tile_blocks(B), 256, tsm_bytes(rows), st>>>(src, idx, first, B, rows, dst);
  CKT(cudaGetLastError());
  c->launches++;
  return DFLOW_OK;
}

// Runs the whole chain on the tile-blocked state `x` in place (x already holds the input), accumulating ldj (per sample).
int tc_run_chain(dflow_chain* c, float* x, const float* theta, const float* theta_const, float* ldj, float* sbuf,
                 long long B, int sampling, int flags, cudaStream_t st) {
  TcPlan* tp = c->tcp;
  const DevChainHdr& Hd = c->hc()->h;
  const int L = (int)tp->layers.size();
  const long long Bp = ((B + 127) / 128) * 128;
  int rc = DFLOW_OK;
  for (int step = 0; step < L; ++step) {
    const int ei = sampling ? step : (L - 1 - step);  // src/Chains.jl:155-161 vs :174-180
    const TcLayer& Ld = tp->layers[ei];
    if (!Ld.is_coupling) {
      tc_norm_kernel<<<ew_blocks(Bp * Hd.d), 256, 0, st>>>(x, x, ldj, B, Hd.d, c->d_staged + Ld.norm_off, sampling);
      CKT(cudaGetLastError());
      c->launches++;
      continue;
    }
    // s and t conditioners as one block-diagonal conditioner: forward-type calls only on request (tc_fuse = 2) -- the
    // zero blocks double the weight image, which costs the hidden-64 nets their resident weights and second CTA per SM
    // (C3 log-density 6.9 ms fused vs 6.2 ms separate); the train step (tc_loss_grad) fuses by default (31.6 vs 36.8 ms)
    const bool fz = Ld.fused && c->tc_fuse >= 2 && !tcs_layer(c, Ld);
    for (int ni = (fz ? 2 : Ld.has_s ? 0 : 1); ni < (fz ? 3 : 2); ++ni) {
      TcArgs a;
      memset(&a, 0, sizeof(a));
      fill_common(c, Ld, a);
      a.im = fz ? Ld.ffwd : Ld.fwd[ni];
      a.net_id = ni;
      a.act1 = Ld.act[ni == 0 ? 0 : 1][0];
      a.act2 = Ld.act[ni == 0 ? 0 : 1][1];
      a.B = B;
      a.sampling = sampling;
      a.flags = flags;
      a.x_in = x;
      a.x_out = x;
      a.theta = theta;
      a.theta_const = theta_const;
      a.ldj = ldj;
      a.sbuf = sbuf;
      rc = launch_net<TC_FWD>(c, a, st, fz ? nullptr : Ld.fwd2[ni]);
      if (rc) return rc;
    }
  }
  return DFLOW_OK;
}

// Forward-type entry points (normalise / log-density / sample) on the tensor-core kernels: transpose the caller's
// sample-major arrays into the tile-blocked working layout, run the chain in place, transpose (or reduce) out.
// Working buffer (grow-only, owned by the plan): [x (d*Bp)] [ldj (Bp)] [theta (n*Bp)].
int tc_fwd(dflow_chain* c, const float* W, FwdArgs& a, cudaStream_t st) {
  if (a.B == 0) return DFLOW_OK;
  TcPlan* tp = c->tcp;
  const DevChainHdr& Hd = c->hc()->h;
  const long long B = a.B, Bp = ((B + 127) / 128) * 128;
  const int d = Hd.d, n = Hd.n;
  int rc = tc_prepack(c, W, false, st);
  if (rc) return rc;
  const bool sampling = a.mode >= MODE_SAMPLE;
  const bool own_ldj = (a.mode == MODE_LOGPDF || a.mode == MODE_LOGPDF_SUM);
  const bool want_ldj = !(a.mode == MODE_SAMPLE || a.mode == MODE_SAMPLE_RNG);
  const bool per_sample_theta = n > 0 && a.theta && !a.theta_const;
  // working state in the caller's scratch (dflow_scratch_bytes / dflow_chain_set_scratch): [x (d Bp) | ldj (Bp) | theta (n Bp) |
  // s values of the current layer (a16max Bp)] -- nothing is allocated here
  if (c->scratch_bytes < tc_scratch_bytes(c, B) || !c->scratch) {
    set_error("scratch too small for B = %lld: need %zu bytes (dflow_scratch_bytes), the handle has %zu "
              "(dflow_chain_set_scratch)", B, tc_scratch_bytes(c, B), c->scratch_bytes);
    return DFLOW_E_INVALID_ARG;
  }
  float* xw = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(c->scratch) + 255) & ~(uintptr_t)255);
  float* ldj = xw + (size_t)d * Bp;
  float* thw = ldj + Bp;
  float* sbuf = thw + (size_t)n * Bp;
  if (a.mode == MODE_SAMPLE_RNG) {
    tc_philox_kernel<<<ew_blocks(B), 256, 0, st>>>(xw, B, d, a.seed, a.rng_offset, a.first_sample);
    CKT(cudaGetLastError());
    c->launches++;
  } else if (a.grid_vals) {
    tc_grid_kernel<<<ew_blocks(B), 256, 0, st>>>(xw, B, d, a.grid_vals, a.grid_meta);
    CKT(cudaGetLastError());
    c->launches++;
  } else {
    rc = gather_t(c, a.x_in, a.idx, 0, B, d, xw, st);
    if (rc) return rc;
  }
  if (per_sample_theta) {
    rc = gather_t(c, a.theta, a.idx, 0, B, n, thw, st);
    if (rc) return rc;
  }
  float* ldj_dst = nullptr;
  if (want_ldj) {
    ldj_dst = own_ldj ? ldj : a.aux_out;
    CKT(cudaMemsetAsync(ldj_dst, 0, sizeof(float) * B, st));
  }
  rc = tc_run_chain(c, xw, per_sample_theta ? thw : nullptr, a.theta_const, ldj_dst, sbuf, B, sampling ? 1 : 0, a.flags, st);
  if (rc) return rc;
  if (own_ldj) {
    tc_logpdf_kernel<<<ew_blocks(B), 256, 0, st>>>(xw, ldj, B, d, Hd.logpdf_c0, 0.0f, nullptr,
                                                   a.mode == MODE_LOGPDF ? a.aux_out : nullptr,
                                                   a.mode == MODE_LOGPDF_SUM ? a.aux_out : nullptr);
  } else {
    tc_scatter_t_kernel<<<tile_blocks(B), 256, tsm_bytes(d), st>>>(xw, B, d, a.x_out);
  }
  CKT(cudaGetLastError());
  c->launches++;
  return DFLOW_OK;
}

size_t tc_scratch_bytes(const dflow_chain* c, long long B) {
  const TcPlan* tp = c->tcp;
  const DevChainHdr& Hd = c->hc()->h;
  const size_t Bp = (size_t)((B + 127) / 128) * 128;
  return ((size_t)(Hd.d + 1 + Hd.n + tp->a16max) * Bp + 128) * sizeof(float) + 256;
}

// ---- training -------------------------------------------------------------------------------------------------
static void train_layout(const dflow_chain* c, long long B, TcTrainLayout& T) {
  const TcPlan* tp = c->tcp;
  const DevChainHdr& Hd = c->hc()->h;
  const long long L = (long long)tp->layers.size();
  // per-sample floats of everything that scales with the macro-batch
  const long long per = (L + 1) * Hd.d + 1 + Hd.d + 2 * Hd.n + L * tp->a16max + L * tp->k0pmax + L * 4 * tp->hu +
                        L * 4 * (tp->hu / 32) + 4 * tp->hu + 2 * tp->a16max;
  long long MB = ((B + 127) / 128) * 128;
  const long long budget = c->tc_ws_budget_mb > 0 ? (long long)c->tc_ws_budget_mb << 20 : (long long)24 << 30;  // bytes
  long long cap = budget / (per * 4);
  cap = std::max<long long>(128, (cap / 128) * 128);
  if (MB > cap) MB = cap;
  T.MB = MB;
  size_t o = 64;
  auto take = [&](long long floats) {
    size_t r = o;
    o += (size_t)((floats + 63) & ~63LL);
    return r;
  };
  T.traj = take((L + 1) * Hd.d * MB);
  T.ldj = take(MB);
  T.zbar = take(Hd.d * MB);
  T.theta = take((long long)Hd.n * MB);
  T.sbuf = take(L * tp->a16max * MB);
  T.inbuf = take(L * tp->k0pmax * MB);
  T.hbuf = take(L * 4 * tp->hu * MB);
  T.mbuf = take(L * 4 * (tp->hu / 32) * MB);
  T.dbuf = take(4LL * tp->hu * MB);
  T.d3buf = take(2LL * tp->a16max * MB);
  T.thbar = take((long long)Hd.n * MB);
  T.total = o;
}

size_t tc_workspace_bytes(const dflow_chain* c, long long B) {
  TcTrainLayout T;
  train_layout(c, B, T);
  return T.total * sizeof(float) + 256;
}

int tc_loss_grad(dflow_chain* c, const float* W, const float* x, const float* theta, long long B, const int32_t* idx,
                 float inv_btot, int flags, float* loss_out, float* grad_out, void* ws, size_t ws_bytes,
                 cudaStream_t st, const TcVjp* vjp) {
  TcPlan* tp = c->tcp;
  const DevChainHdr& Hd = c->hc()->h;
  (void)ws_bytes;
  TcTrainLayout T;
  train_layout(c, B, T);
  float* wsf = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  const int L = (int)tp->layers.size(), d = Hd.d, n = Hd.n, hu = tp->hu, a16m = tp->a16max, k0m = tp->k0pmax;
  int rc = tc_prepack(c, W, true, st);
  if (rc) return rc;
  for (long long first = 0; first < B; first += T.MB) {
    const long long mb = std::min<long long>(T.MB, B - first);
    const long long ntiles = (mb + 127) / 128;
    float* traj = wsf + T.traj;
    float* ldj = wsf + T.ldj;
    float* zbar = wsf + T.zbar;
    float* thg = wsf + T.theta;
    const size_t st_d = (size_t)d * T.MB;
    auto slot = [&](int e) { return traj + (size_t)e * st_d; };
    auto sbuf_of = [&](int e) { return wsf + T.sbuf + (size_t)e * a16m * T.MB; };
    auto inbuf_of = [&](int e) { return wsf + T.inbuf + (size_t)e * k0m * T.MB; };
    // per layer four hu-wide slots: [net][h1 | h2]; a fused pair (width 2 hu) uses them as [h1 (2 slots) | h2 (2 slots)]
    auto hslot = [](int net, int which) { return net == 2 ? which * 2 : net * 2 + which; };
    auto hbuf_of = [&](int e, int net, int which) {
      return wsf + T.hbuf + ((size_t)e * 4 + hslot(net, which)) * (size_t)hu * T.MB;
    };
    auto mbuf_of = [&](int e, int net, int which) {
      return reinterpret_cast<uint32_t*>(wsf + T.mbuf) + ((size_t)e * 4 + hslot(net, which)) * (size_t)(hu / 32) * T.MB;
    };
    auto dbuf_of = [&](int net, int which) { return wsf + T.dbuf + (size_t)hslot(net, which) * (size_t)hu * T.MB; };
    auto d3buf_of = [&](int net) { return wsf + T.d3buf + (size_t)net * a16m * T.MB; };
    // input slot L: gather (or copy) this macro-batch
    rc = gather_t(c, x, idx, first, mb, d, slot(L), st);
    if (rc) return rc;
    const float* th = nullptr;
    if (n > 0) {
      rc = gather_t(c, theta, idx, first, mb, n, thg, st);
      if (rc) return rc;
      th = thg;
    }
    CKT(cudaMemsetAsync(ldj, 0, sizeof(float) * mb, st));
    // ---- forward (normalising) sweep with stored activations: elements L-1 .. 0 ----
    for (int ei = L - 1; ei >= 0; --ei) {
      const TcLayer& Ld = tp->layers[ei];
      if (!Ld.is_coupling) {
        tc_norm_kernel<<<ew_blocks(ntiles * 128 * d), 256, 0, st>>>(slot(ei + 1), slot(ei), ldj, mb, d, c->d_staged + Ld.norm_off, 0);
        CKT(cudaGetLastError());
        c->launches++;
        continue;
      }
      const bool fz = Ld.fused && c->tc_fuse > 0 && !tcs_layer(c, Ld);
      for (int ni = (fz ? 2 : Ld.has_s ? 0 : 1); ni < (fz ? 3 : 2); ++ni) {
        TcArgs a;
        memset(&a, 0, sizeof(a));
        fill_common(c, Ld, a);
        a.im = fz ? Ld.ffwd : Ld.fwd[ni];
        a.net_id = ni;
        a.act1 = Ld.act[ni == 0 ? 0 : 1][0];
        a.act2 = Ld.act[ni == 0 ? 0 : 1][1];
        a.B = mb;
        a.sampling = 0;
        a.flags = flags;
        a.x_in = slot(ei + 1);
        a.x_out = slot(ei);
        a.theta = th;
        a.ldj = ldj;
        a.sbuf = sbuf_of(ei);
        a.inbuf = inbuf_of(ei);
        a.h1buf = hbuf_of(ei, ni, 0);
        a.h2buf = hbuf_of(ei, ni, 1);
        a.m1buf = mbuf_of(ei, ni, 0);
        a.m2buf = mbuf_of(ei, ni, 1);
        rc = launch_net<TC_FWD_STORE>(c, a, st, fz ? nullptr : Ld.fwd2[ni]);
        if (rc) return rc;
      }
    }
    // ---- loss and seeds (or the caller's cotangents: dflow_vjp) ----
    const bool ext = vjp && vjp->zbar;
    if (loss_out || !ext) {
      tc_logpdf_kernel<<<ew_blocks(mb), 256, 0, st>>>(slot(0), ldj, mb, d, Hd.logpdf_c0, inv_btot, ext ? nullptr : zbar,
                                                     nullptr, loss_out);
      CKT(cudaGetLastError());
      c->launches++;
    }
    if (ext) {
      rc = gather_t(c, vjp->zbar, nullptr, first, mb, d, zbar, st);
      if (rc) return rc;
    }
    if (vjp && vjp->z_out) {
      tc_scatter_t_kernel<<<tile_blocks(mb), 256, tsm_bytes(d), st>>>(slot(0), mb, d, vjp->z_out + (size_t)first * d);
      CKT(cudaGetLastError());
      c->launches++;
    }
    if (vjp && vjp->ldj_out) CKT(cudaMemcpyAsync(vjp->ldj_out + first, ldj, sizeof(float) * mb, cudaMemcpyDeviceToDevice, st));
    float* thbar = nullptr;
    if (vjp && vjp->thbar_out && n > 0) {
      thbar = wsf + T.thbar;
      CKT(cudaMemsetAsync(thbar, 0, sizeof(float) * (size_t)n * ntiles * 128, st));
    }
    // ---- reverse sweep in chain order ----
    int last_coupling = -1;
    for (int ei = 0; ei < L; ++ei)
      if (tp->layers[ei].is_coupling) last_coupling = ei;
    if (vjp && vjp->xbar_out) last_coupling = L - 1;  // the cotangent of x also crosses a trailing NormalizationLayer
    for (int ei = 0; ei <= last_coupling; ++ei) {
      const TcLayer& Ld = tp->layers[ei];
      if (!Ld.is_coupling) {
        tc_norm_bwd_kernel<<<ew_blocks(ntiles * 128 * d), 256, 0, st>>>(zbar, mb, d, c->d_staged + Ld.norm_off);
        CKT(cudaGetLastError());
        c->launches++;
        continue;
      }
      const bool fz = Ld.fused && c->tc_fuse > 0 && !tcs_layer(c, Ld);
      for (int ni = (fz ? 2 : Ld.has_s ? 0 : 1); ni < (fz ? 3 : 2); ++ni) {
        TcArgs a;
        memset(&a, 0, sizeof(a));
        fill_common(c, Ld, a);
        a.im = fz ? Ld.fbwd : Ld.bwd[ni];
        a.net_id = ni;
        a.act1 = Ld.act[ni == 0 ? 0 : 1][0];
        a.act2 = Ld.act[ni == 0 ? 0 : 1][1];
        a.B = mb;
        a.flags = flags;
        a.sbuf = sbuf_of(ei);
        a.h1buf = hbuf_of(ei, ni, 0);  // non-relu hidden layers take their derivative from the stored activations
        a.h2buf = hbuf_of(ei, ni, 1);
        a.m1buf = mbuf_of(ei, ni, 0);
        a.m2buf = mbuf_of(ei, ni, 1);
        a.d1buf = dbuf_of(ni, 0);
        a.d2buf = dbuf_of(ni, 1);
        a.d3buf = d3buf_of(fz ? 0 : ni);
        a.zbar = zbar;
        a.zout = slot(ei);
        a.inv_btot = inv_btot;
        a.jbar = (vjp && vjp->jbar) ? vjp->jbar + first : nullptr;
        a.thbar = thbar;
        a.grad = grad_out;
        a.p_b3 = Ld.p_b[fz ? 0 : ni][2];
        rc = launch_net<TC_BWD>(c, a, st, fz ? nullptr : Ld.bwd2[ni]);
        if (rc) return rc;
      }
      // weight gradients of this layer
      DwArgs w;
      memset(&w, 0, sizeof(w));
      const bool fzw = Ld.fused && c->tc_fuse > 0 && !tcs_layer(c, Ld);
      w.H = fzw ? 2 * Ld.h : Ld.h;
      w.K0p = Ld.fwd[1].K0p;
      w.K0 = Ld.nin;
      w.a = Ld.a;
      w.a16 = fzw ? 2 * Ld.a16 : Ld.a16;
      w.fused = fzw ? 1 : 0;
      w.hblk = Ld.h;
      w.ablk = Ld.a16;
      w.mtiles = (w.H + 127) / 128;
      w.nsplit = w.H > 256 ? 2 : 1;
      // hidden 256 with many conditioner inputs and transformed coordinates: more row-blocks per stage than the staging
      // warps of tc_dw_kernel cover -> split the dW2 columns over two CTAs as at hidden 512
      if (w.nsplit == 1 && !fzw && !dwts_shape_ok(w.H, w.K0p, w.a16) &&
          (3 * std::min(128, w.H) + w.H + w.K0p + w.a16) / 8 > DW_MAXRB * DW_STAGE_WARPS && w.H % 32 == 0)
        w.nsplit = 2;
      w.NB = w.H / w.nsplit;
      w.first_net = (fzw || Ld.has_s) ? 0 : 1;
      w.units = (fzw ? 1 : Ld.has_s ? 2 : 1) * w.mtiles * w.nsplit;
      w.ksplit = (int)std::max<long long>(1, std::min<long long>(ntiles, c->sm_count / w.units));
      w.ntiles = ntiles;
      for (int ni = 0; ni < 2; ++ni) {
        const int bi = fzw ? 2 : ni;  // fused: one set of buffers (slot scheme of hbuf_of)
        w.h1buf[ni] = hbuf_of(ei, bi, 0);
        w.h2buf[ni] = hbuf_of(ei, bi, 1);
        w.d1buf[ni] = dbuf_of(bi, 0);
        w.d2buf[ni] = dbuf_of(bi, 1);
        w.d3buf[ni] = d3buf_of(fzw ? 0 : ni);
        for (int j = 0; j < 3; ++j) {
          w.p_w[ni][j] = Ld.p_w[ni][j];
          w.p_b[ni][j] = Ld.p_b[ni][j];
        }
      }
      w.inbuf = inbuf_of(ei);
      w.grad = grad_out;
      w.debug = (c->tc_debug >> 12) & 7;  // bits 12-14 of tc_debug: weight-gradient kernel timing experiments (experiment builds)
      if (c->tc_dw_ts >= 0 && dwts_shape_ok(w.NB, w.K0p, w.a16)) {
        // A operands through TMEM (dflow_tc_dwts.cuh): only the B segments are staged in shared memory
        const size_t stage_bytes = 2 * (size_t)(w.NB + w.K0p + w.a16) * DW_KS * 4;
        int nst = (int)(((size_t)c->max_smem_optin - 256) / stage_bytes);
        if (nst > 4) nst = 4;
        nst &= ~1;
        if (nst < 2) {
          set_error("weight-gradient stage does not fit (%zu bytes per stage)", stage_bytes);
          return DFLOW_E_UNSUPPORTED;
        }
        w.nstage = nst;
        w.ngroups = 1;
        const size_t smem = nst * stage_bytes + 256;
        CKT(cudaFuncSetAttribute(tc_dwts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_dwts_kernel<<<(unsigned)(w.units * w.ksplit), DW_THREADS, smem, st>>>(w);
        CKT(cudaGetLastError());
        c->launches++;
      } else {
        const int RA = std::min(128, w.H);  // rows of the A segments (a multiple of 8: H % 32 == 0)
        const size_t stage_bytes = 2 * (size_t)(3 * RA + w.NB + w.K0p + w.a16) * DW_KS * 4;  // the fullest CTA (column half 0)
        // the M = 128 MMAs read 128 rows of every A segment: keep that overrun inside the allocation
        const size_t overrun = (size_t)(128 - RA) * DW_KS * 4 * 2;
        int nst = (int)(((size_t)c->max_smem_optin - 128 - overrun) / stage_bytes);
        if (nst > 4) nst = 4;
        nst &= ~1;  // stage pairs
        const int nrb = (3 * RA + w.NB + w.K0p + w.a16) / 8;  // row-blocks per stage
        w.ngroups = 1;
        for (int g = 3; g >= 2; --g)
          if (w.ngroups == 1 && nst >= 2 * g && nrb <= DW_MAXRB * (DW_STAGE_WARPS / g)) w.ngroups = g;
        if (c->tc_dw_groups > 0 && c->tc_dw_groups <= w.ngroups) w.ngroups = c->tc_dw_groups;
        if (nst < 2 || nrb > DW_MAXRB * DW_STAGE_WARPS) {
          set_error("weight-gradient stage does not fit (%zu bytes per stage)", stage_bytes);
          return DFLOW_E_UNSUPPORTED;
        }
        w.nstage = nst;
        const size_t smem = nst * stage_bytes + overrun + 128;
        CKT(cudaFuncSetAttribute(tc_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_dw_kernel<<<(unsigned)(w.units * w.ksplit), DW_THREADS, smem, st>>>(w);
        CKT(cudaGetLastError());
        c->launches++;
      }
    }
    if (vjp && vjp->xbar_out) {
      tc_scatter_t_kernel<<<tile_blocks(mb), 256, tsm_bytes(d), st>>>(zbar, mb, d, vjp->xbar_out + (size_t)first * d);
      CKT(cudaGetLastError());
      c->launches++;
    }
    if (thbar) {
      tc_scatter_t_kernel<<<tile_blocks(mb), 256, tsm_bytes(n), st>>>(thbar, mb, n, vjp->thbar_out + (size_t)first * n);
      CKT(cudaGetLastError());
      c->launches++;
    }
  }
  return DFLOW_OK;
}

}  // namespace dflow
