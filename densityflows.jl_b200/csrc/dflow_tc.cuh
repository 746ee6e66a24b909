// tcgen05 / TMEM / mbarrier / bulk-copy PTX wrappers shared by the tensor-core kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dflow {
namespace tc {

// float index of element (r, k) of an [R x Kc] K-major operand block in the no-swizzle UMMA core-matrix layout:
// 8 x 16-byte core matrices, LBO (next core along K) = 128 B, SBO (next 8 rows) = (Kc/4) * 128 B.
// The same bytes read as an MN-major operand give the transposed block (unit <-> sample), see dflow_tc_dw.
__host__ __device__ inline int core_idx(int r, int k, int Kc) {
  return (r >> 3) * (Kc >> 2) * 32 + (k >> 2) * 32 + (r & 7) * 4 + (k & 3);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// The same rounding (nearest, ties away from zero) for the hot epilogues: cvt.rna.tf32.f32 compiles to
// FSETP |x| >= inf, VIADD 0x1000, SEL, LOP3 -- the inf / nan guard is not needed here (inf stays inf, a nan stays a nan,
// a value that rounds past the largest float becomes inf as it should), so two of the four instructions go.
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

// ---- mbarrier ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}

// variants on 32-bit shared-window addresses (no generic -> shared conversion in hot loops)
__device__ __forceinline__ void mbar_wait_a(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mma_commit_a(uint32_t addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mma_commit_multicast_a(uint32_t addr, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   addr),
               "h"(cta_mask)
               : "memory");
}

// global -> shared bulk copy (TMA engine, no tensor map), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// the same copy delivered to the same shared-memory offset (and mbarrier offset) of every CTA in cta_mask
__device__ __forceinline__ void bulk_g2s_multicast(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                                   uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// 16 consecutive fp32 columns of this thread's TMEM lane (warp w of the CTA reads lanes 32*(w%4)..+31)
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  tmem_ld16_nowait(taddr, r);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- UMMA descriptors -----------------------------------------------------------------------------------------
// shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor, version 1).  K-major operands:
// lbo = byte distance of K-adjacent core matrices, sbo = byte distance of 8-row groups.  MN-major operands:
// sbo = byte distance of MN-adjacent 4-element groups, lbo = byte distance of 8-deep K groups.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout_type SWIZZLE_NONE
}
// instruction descriptor: kind::tf32, fp32 accumulate, M = 128 (cute::UMMA::InstrDescriptor); a_mn / b_mn select
// MN-major ("transposed") operands
__device__ __forceinline__ uint32_t instr_desc_tf32(int n, int a_mn = 0, int b_mn = 0) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// arrives on the mbarrier when every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// the same arrive on the mbarrier at this offset in every CTA of cta_mask
__device__ __forceinline__ void mma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// ---- CTA pair (cta_group::2): one issuer feeds the tensor cores of two SMs; D rows 0-127 live in the leader's TMEM,
// rows 128-255 in the peer's; every CTA holds its own A rows and one half of the B rows at the same shared offsets
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot_in_smem, uint32_t ncols) {  // one full warp in EACH CTA
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_tf32_2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_commit2_multicast_a(uint32_t addr, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   addr),
               "h"(cta_mask)
               : "memory");
}
// arrive on the barrier at the same shared offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t local_addr, uint32_t cta) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
// instruction descriptor with M = 256 (cta_group::2)
__device__ __forceinline__ uint32_t instr_desc_tf32_m256(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((256u >> 4) << 24);
}

// one lane of a converged warp (the warp keeps executing uniformly, so descriptors stay in uniform registers)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// K-major no-swizzle descriptor from its two halves: hi = (sbo >> 4) | version, lo = (addr >> 4) | (lbo = 128 B) << 16
__device__ __forceinline__ uint32_t desc_hi(int kc_floats) { return ((uint32_t)(kc_floats >> 2) * 128u >> 4) | (1u << 14); }
__device__ __forceinline__ uint64_t desc_at(uint32_t hi, uint32_t saddr) {
  return ((uint64_t)hi << 32) | (uint64_t)(((saddr >> 4) & 0x3FFFu) | (8u << 16));
}
// 3xTF32 product over `ksteps` K steps of 8 with prebuilt descriptors (advance = 256 B = 16 descriptor units per step)
__device__ __forceinline__ void gemm3_desc(uint32_t d_tmem, uint64_t dah, uint64_t dal, uint64_t dbh, uint64_t dbl,
                                           int ksteps, uint32_t idesc, uint32_t acc) {
  for (int ks = 0; ks < ksteps; ++ks) {
    const uint64_t o = (uint64_t)(ks * 16);
    mma_tf32(d_tmem, dal + o, dbh + o, idesc, acc);
    mma_tf32(d_tmem, dah + o, dbl + o, idesc, 1u);
    mma_tf32(d_tmem, dah + o, dbh + o, idesc, 1u);
    acc = 1u;
  }
}

// D[128 x n] (+)= A[128 x K] * B[n x K]^T with the 3xTF32 split (hi*hi + hi*lo + lo*hi, small terms first);
// operands K-major in the core layout with row pitches a_kc / b_kc (floats), K a multiple of 8.
__device__ __forceinline__ void gemm_3xtf32(uint32_t d_tmem, const float* a_hi, const float* a_lo, int a_kc,
                                            const float* b_hi, const float* b_lo, int b_kc, int K, int n,
                                            bool accumulate) {
  const uint32_t idesc = instr_desc_tf32(n);
  const uint32_t a_sbo = (uint32_t)(a_kc >> 2) * 128u, b_sbo = (uint32_t)(b_kc >> 2) * 128u;
  const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
  uint32_t acc = accumulate ? 1u : 0u;
  for (int ks = 0; ks < K; ks += 8) {
    const uint32_t off = (uint32_t)(ks >> 2) * 128u;  // two 128-byte core matrices per K step of 8
    const uint64_t dah = smem_desc(ah + off, 128u, a_sbo), dal = smem_desc(al + off, 128u, a_sbo);
    const uint64_t dbh = smem_desc(bh + off, 128u, b_sbo), dbl = smem_desc(bl + off, 128u, b_sbo);
    mma_tf32(d_tmem, dal, dbh, idesc, acc);
    mma_tf32(d_tmem, dah, dbl, idesc, 1u);
    mma_tf32(d_tmem, dah, dbh, idesc, 1u);
    acc = 1u;
  }
}

}  // namespace tc
}  // namespace dflow
