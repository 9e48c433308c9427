// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld / fences).  No CUTLASS dependency.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace semgate {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() {
  uint32_t l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

// One lane of a converged warp gets `true`.  Keeping role loops warp-uniform and electing
// only around the asynchronous instruction lets ptxas keep addresses/descriptors in uniform
// registers instead of a per-instruction R2UR waterfall.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred));
  return pred != 0;
}

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// arrive on the barrier at the same smem offset in CTA `cta` of the cluster.
// Default (cta-scope) semantics on purpose: a `.release.cluster` arrive compiles to
// MEMBAR.ALL.GPU + ERRBAR in the issuing thread, which drains the TMA pipeline every stage.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(cta) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

// non-blocking probe (try_wait may suspend the thread for a system-dependent time before it reports failure:
// measured ~2.5 us per failed probe, which is a k-block's worth of MMAs)
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

// Bounded wait: a protocol bug must surface as a launch failure, never as a hung
// GPU.  The timer is only consulted after many failed probes.
#ifndef SEMGATE_WAIT_TIMEOUT_NS
#define SEMGATE_WAIT_TIMEOUT_NS 4000000000ull
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xFFFu) == 0) {
      uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > SEMGATE_WAIT_TIMEOUT_NS) __trap();
    }
  }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tensormap(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(desc) : "memory");
}

// L2 eviction-priority policies for TMA loads (createpolicy.fractional encodings, fraction 1.0).
constexpr uint64_t kL2EvictNormal = 0x1000000000000000ull;
constexpr uint64_t kL2EvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kL2EvictLast = 0x14F0000000000000ull;

// 2D tile load global -> this CTA's smem, completion on this CTA's mbarrier.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(desc), "r"(bar), "r"(c0), "r"(c1), "l"(policy) : "memory");
}

// Same, issued from a CTA of a cta_group::2 pair.  `dst`/`bar` are shared::cluster
// addresses (the barrier may live in the peer CTA).
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(desc), "r"(bar), "r"(c0), "r"(c1), "l"(policy) : "memory");
}

// gpu-scope relaxed counter access for the pacing counters (no data is published through them)
__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_add_relaxed_gpu(uint32_t* p, uint32_t v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void nanosleep(uint32_t ns) { asm volatile("nanosleep.u32 %0;" ::"r"(ns)); }

// Multicast variant: the tile lands at the same smem offset in every CTA of `mask`; with
// cta_group::2 the complete_tx of each destination goes to the barrier at `bar`'s offset in the
// even (leader) CTA of that destination's pair when `bar` has its peer bit cleared (pair_leader_addr).
__device__ __forceinline__ void tma_load_2d_cg2_mc(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1,
                                                    uint16_t mask, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5, %6;"
      ::"r"(dst), "l"(desc), "r"(bar), "r"(c0), "r"(c1), "h"(mask), "l"(policy) : "memory");
}

// A CTA's own shared-window address names its smem inside the cluster; bit 24 selects the odd CTA
// of a cta_group::2 pair.  Clearing it addresses the same offset in the pair's leader.
__device__ __forceinline__ uint32_t pair_leader_addr(uint32_t addr) { return addr & 0xFEFFFFFFu; }

__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}

// 3-D tile load issued from a CTA of a cta_group::2 pair (`bar` may name the leader's barrier)
__device__ __forceinline__ void tma_load_3d_cg2(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(desc), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// ---------------------------------------------------------------- distributed shared memory (data, not only barriers)
// `addr` is a shared::cluster address (mapa).  The reds / stores are relaxed; the release-arrive below publishes them.
__device__ __forceinline__ void red_max_u32_cluster(uint32_t addr, uint32_t v) {
  asm volatile("red.relaxed.cluster.shared::cluster.max.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_f32_cluster(uint32_t addr, float v) {
  asm volatile("st.relaxed.cluster.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// release at cluster scope: everything this thread wrote before (remote reds / stores included) is visible to
// whoever acquires the phase.  Costs a fence; meant for once-per-work-item hand-offs, not per stage.
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(cta) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_acquire_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait_acquire_cluster(bar, parity)) {
    if ((++spins & 0xFFFu) == 0) {
      uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > SEMGATE_WAIT_TIMEOUT_NS) __trap();
    }
  }
}

// ---------------------------------------------------------------- cluster
__device__ __forceinline__ void cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_sync() {
  cluster_arrive();
  cluster_wait();
}

// ---------------------------------------------------------------- tcgen05
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  if constexpr (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}

template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32 (kind::f16)
template <int CG>
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  if constexpr (CG == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// The four K = 16 MMAs of one 64-wide k-block, cta_group::2, from the LOW words of the two shared-memory descriptors
// (make_smem_desc_sw128: the high word is the constant kSmemDescHi; stepping 16 bf16 inside the 128-byte swizzle row adds
// 2 to the address field, which never carries out of the low word).  32-bit arithmetic on uniform values instead of four
// 64-bit descriptor adds per operand: the issuing thread's loop is what bounds a kernel whose k-blocks are short.
constexpr uint32_t kSmemDescHi = 0x40004040u;       // SBO = 1024 >> 4 | version 1 << 14 | SWIZZLE_128B << 29
__device__ __forceinline__ uint32_t smem_desc_lo_sw128(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ void umma4_bf16_cg2(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate_first) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t.reg .b32 a1, b1;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.eq.b32 q, %4, %4;\n\t"
      "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t"
      "add.u32 a1, %1, 2;\n\tadd.u32 b1, %2, 2;\n\tmov.b64 da, {a1, %5};\n\tmov.b64 db, {b1, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, q;\n\t"
      "add.u32 a1, %1, 4;\n\tadd.u32 b1, %2, 4;\n\tmov.b64 da, {a1, %5};\n\tmov.b64 db, {b1, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, q;\n\t"
      "add.u32 a1, %1, 6;\n\tadd.u32 b1, %2, 6;\n\tmov.b64 da, {a1, %5};\n\tmov.b64 db, {b1, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, q;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate_first), "r"(kSmemDescHi) : "memory");
}

// All MMAs issued so far by this thread arrive on `bar` (this CTA) once complete.
__device__ __forceinline__ void umma_commit_cg1(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// cta_group::1 commit that arrives on the barrier at this offset in every CTA of `mask`
// (a stage that peers multicast into is free only when every consumer is done with it).
__device__ __forceinline__ void umma_commit_cg1_mc(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(mask) : "memory");
}
// cta_group::2: arrive on the barrier at this offset in every CTA of `mask`.
__device__ __forceinline__ void umma_commit_cg2_mc(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(mask) : "memory");
}

__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives lane
// (base_lane + t), columns [col, col+32).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle, rows of
// 64 bf16 (128 B) packed contiguously; 8-row groups are 1024 B apart (SBO).
// Fields: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) |
// layout_type=2 (SWIZZLE_128B) [61,64).
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;               // LBO (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;       // SBO
  d |= static_cast<uint64_t>(1) << 46;               // descriptor version (sm_100)
  d |= static_cast<uint64_t>(2) << 61;               // SWIZZLE_128B
  return d;
}

// Instruction descriptor for kind::f16: C=F32 [4,6)=1, A=BF16 [7,10)=1, B=BF16
// [10,13)=1, A/B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_bf16_f32(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

}  // namespace ptx
}  // namespace semgate
