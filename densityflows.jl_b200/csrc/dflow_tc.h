// Tensor-core (tcgen05) path, generation 2: plan structures shared by dflow_api.cu and dflow_tc.cu.
//
// A conditioner Dense(in,h,relu) -> Dense(h,h,relu) -> Dense(h,a) and its transposed (adjoint) chain
// delta3 -> (.W3) mask -> (.W2) mask -> (.W1) are the same three-GEMM pipeline with different matrices and epilogues,
// so both are described by one "net image": three matrices M1 [H x K0], M2 [H x H], M3 [N3 x H], pre-split into
// TF32 hi/lo parts and laid out as a sequence of stage blocks in consumption order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "dflow_internal.h"

namespace dflow {

constexpr int TC_WKC = 16;     // hidden units per pipeline chunk (N of GEMM 1, K of GEMMs 2 and 3)
constexpr int TC_NSMAX = 16;   // ring slots
constexpr int TC_THREADS = 448;  // 8 chunk-epilogue warps (two warpgroups) + 4 loader/output warps + producer + MMA warp
constexpr int TC_DW_KS = 16;   // samples per stage of the weight-gradient kernel

struct TcNetImg {
  long long off;  // floats, start of this image inside the plan's weight image
  int K0, K0p;    // GEMM-1 depth, padded to a multiple of 8
  int H;          // hidden width
  int N3, N3p;    // GEMM-3 outputs, padded to a multiple of 16
  int NH, passes, nch, nch_pass;
  int GW, ng;       // D1 is produced in groups of GW hidden units (64, or 32 when H % 64 != 0); ng = H / GW
  int g1_floats;    // G1 block: [M1 group hi | lo] (GW x K0p each)
  int s2_floats;    // S2 block: [M2 chunk hi | lo] (NH x WKC each)
  int s3_floats;    // S3 block: [M3 chunk hi | lo] (N3p x WKC each)
  int slot_floats;  // ring slot = largest block
  int bias_off;     // [b1 (H) | b2 (H) | b3 (N3p)] (forward orientation; zeros otherwise)
  int g1_off, s2_off, s3_off;
  int blocks_floats;  // size of the contiguous [G1 | S2 | S3] region
  int halves, rank;   // CTA-pair images: every block holds rows [rank * R/2, (rank+1) * R/2) of its R logical rows
  int total;      // floats
};

struct TcPackJob {
  TcNetImg im;
  // element (n, k) of matrix j is W[base + n * sn + k * sk]; M1 valid for k < vk1, M3 valid for n < vn3
  int base1, sn1, sk1, vk1;
  int base2, sn2, sk2;
  int base3, sn3, sk3, vn3;
  int pb1, pb2, pb3, nb3;  // bias offsets in the packed buffer (-1: zeros), number of real b3 entries
  // Fused s + t pair (one conditioner of width 2h with block-diagonal matrices): the `b` fields address the second net,
  // rbs / cbs are the row / column block sizes of the three matrices (0: that dimension is shared by both nets), vn / vk
  // the valid rows / columns inside a block.  A (row block, column block) pair off the diagonal is zero.
  int fuse;
  int base1b, base2b, base3b, pb1b, pb2b, pb3b;
  int rbs1, cbs1, rbs2, cbs2, rbs3, cbs3;
  int vn1, vn2, vk2, vk3;
  int bbs12, bbs3;  // block sizes of the bias vectors b1/b2 and b3
};

struct TcLayer {
  int is_coupling, has_s;
  int h, nin, a, a16;
  int norm_off;
  unsigned char af[DMAX], id[DMAX];
  int p_w[2][3], p_b[2][3];
  int act[2][2];  // [net][hidden layer]: DFLOW_ACT_* of the two hidden Dense layers
  TcNetImg fwd[2], bwd[2];
  TcNetImg fwd2[2][2], bwd2[2][2];  // [net][rank of the CTA pair]: half-row images for cta_group::2
  // hidden <= 128 RealNVP layers: the s and t conditioners (same input, same widths) run as ONE conditioner of width 2h
  // with block-diagonal W2 / W3 -- half the launches, MMA issues and pipeline hand-offs on an issue-bound shape
  int fused;
  int tcs;  // every conditioner of the layer (both orientations) fits the narrow TMEM-sourced kernel (dflow_tcs.cuh)
  TcNetImg ffwd, fbwd;
};

// per-layer training buffers (offsets in floats inside the caller's workspace)
struct TcTrainLayout {
  long long MB;        // macro-batch (samples, multiple of 128)
  size_t traj;         // [L+1][d * MB]   states: slot e = output of element e (normalising direction), slot L = input
  size_t ldj;          // [MB]
  size_t zbar;         // [d * MB]
  size_t theta;        // [n * MB] theta in the tile-blocked layout
  size_t sbuf;         // [L][a16max * MB]
  size_t inbuf;        // [L][K0pmax * MB]
  size_t hbuf;         // [L][2 nets][2][hmax * MB]   post-relu activations h1, h2
  size_t mbuf;         // [L][2 nets][2][(hmax/32) * MB]  relu masks (uint16 per 16-unit chunk)
  size_t dbuf;         // [2 nets][2][hmax * MB]      delta1, delta2 of the current layer
  size_t d3buf;        // [2 nets][a16max * MB]
  size_t thbar;        // [n * MB] cotangent of the conditions (dflow_vjp)
  size_t total;        // floats
};

struct TcPlan {
  std::vector<TcLayer> layers;  // chain order
  std::vector<TcPackJob> jobs_fwd, jobs_bwd;
  TcPackJob* d_jobs_fwd = nullptr;
  TcPackJob* d_jobs_bwd = nullptr;
  float* d_img = nullptr;
  size_t img_floats = 0;
  int hmax = 0, a16max = 0, k0pmax = 0;
  int hu = 0;  // widest single conditioner (training buffers hold 2 nets x 2 activations x hu per layer and sample)
  bool train_ok = false;  // adjoint supported (h <= 256)
};

int tc_build_plan(dflow_chain* c);
void tc_free_plan(dflow_chain* c);
int tc_prepack(dflow_chain* c, const float* W, bool with_bwd, cudaStream_t st);
int tc_fwd(dflow_chain* c, const float* W, FwdArgs& a, cudaStream_t st);
int tc_run_chain(dflow_chain* c, float* x, const float* theta, const float* theta_const, float* ldj, float* sbuf,
                 long long B, int sampling, int flags, cudaStream_t st);
size_t tc_scratch_bytes(const dflow_chain* c, long long B);  // working state of a forward-type call on B samples
size_t tc_workspace_bytes(const dflow_chain* c, long long B);
// caller cotangents / extra outputs of dflow_vjp (all sample-major device arrays; any pointer may be null)
struct TcVjp {
  const float* zbar;  // (d, B)
  const float* jbar;  // (B)
  float* xbar_out;    // (d, B)
  float* thbar_out;   // (n, B)
  float* z_out;       // (d, B) primal outputs of the normalising sweep
  float* ldj_out;     // (B)
};
int tc_loss_grad(dflow_chain* c, const float* W, const float* x, const float* theta, long long B, const int32_t* idx,
                 float inv_btot, int flags, float* loss_out, float* grad_out, void* ws, size_t ws_bytes,
                 cudaStream_t st, const TcVjp* vjp = nullptr);

}  // namespace dflow
