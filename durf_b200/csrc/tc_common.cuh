// PTX wrappers shared by the tcgen05 kernels (mlp_tc.cu forward, mlp_tc_dgrad.cu, mlp_tc_wgrad.cu): mbarriers, bulk copies,
// tcgen05.mma / ld / st / commit, shared-memory matrix descriptors.  sm_100a only.
#pragma once

#include "common.cuh"

// Per-role cycle counters of the training kernels (DURF_TC_TRACE / DURF_WGRAD_TRACE=2) are compiled in only with
// `make EXTRA=-DDURF_TRACE_DETAIL=1`: their 64-bit accumulators cost registers in kernels that have none to spare
// (measured: +4 % on the forward-with-save, +8 % on the data-gradient kernel).
#ifndef DURF_TRACE_DETAIL
#define DURF_TRACE_DETAIL 0
#endif
// The forward kernel's own counters (DURF_TC_TRACE=1: where block 0's MMA thread and one epilogue thread spent their cycles)
// need `make EXTRA=-DDURF_TRACE=1`: even switched off at run time their ~25 predicated instructions per epilogue issue.
#ifndef DURF_TRACE
#define DURF_TRACE DURF_TRACE_DETAIL
#endif

namespace durf {

constexpr int kTileM = 128;
constexpr int kBlockBytes = 16384;    // one operand block: 128 rows x 64 bf16, 128-byte rows, SWIZZLE_128B (8-row groups of 1 KB)

// ---- PTX wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Spin on try_wait (which itself blocks for a bounded hardware interval).  A watchdog turns a protocol bug into a
// trapped launch (reported by the next CUDA call) instead of a hung GPU: no legitimate wait lasts 4e9 cycles.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = 0;
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    const long long now = clock64();
    if (t0 == 0) t0 = now;
    else if (now - t0 > 4000000000LL) {
      printf("durf tcgen05 kernel: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// ---- thread-block clusters: CTAs of a cluster share each weight fetch (one L2 read, delivered to every CTA) --------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Same copy delivered to the same CTA-relative offset of every CTA in cta_mask; each destination CTA's mbarrier (same
// offset) receives the complete_tx.
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t cta_mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// Shared-memory matrix descriptors are passed as (lo, hi) words: hi is constant (SBO, version, swizzle), lo holds the
// start address >> 4, so stepping through K blocks is a 32-bit add.
// Warp-converged variants: ALL lanes of the issuing warp execute the statement with identical operands, the instruction
// itself is predicated on elect.sync.  Keeping the control flow uniform lets ptxas hold descriptors in uniform registers
// instead of re-broadcasting them (R2UR) for every instruction.
__device__ __forceinline__ void umma_ss_conv(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, e;\n"
      ".reg .b64 da, db;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "setp.ne.b32 p, %5, 0;\n"
      "mov.b64 da, {%1, %3};\n"
      "mov.b64 db, {%2, %3};\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n"
      "}\n" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate) : "memory");
}
// One K block (4 x K=16) of MMAs by a converged warp, bracketed by a software-pipelined wait on the barriers the NEXT
// group of MMAs needs.  Why: an instruction of the issuing thread that returns a value through the memory pipe (mbarrier
// test/try_wait, any load) completes only after the tcgen05.mma's issued before it have left the queue, so a wait placed
// between two groups drains the tensor pipe (measured: 260-390 cycles for a wait on an already completed barrier).
// Here the barriers are probed (non-blocking test_wait) BEFORE this group's MMAs and the outcome is consumed AFTER them:
// if the data of the next group is already there - the common case - the thread never stalls with an empty queue
// behind it; only a failed probe falls into the blocking try_wait loop.  need_x == 0 disables barrier x (pass any valid
// barrier address).  do_commit: tcgen05.commit -> bar_commit right after the MMAs (before any blocking wait, so the
// consumer of the accumulator is never held up by this thread's wait).  do_release: a second commit that frees the ring
// stage this K block was the last reader of.  A_TMEM: A operand from tensor memory (a = TMEM address, +8 columns per K=16 step), else from a
// shared-memory descriptor (a = descriptor lo word, +2 per step); b_lo advances by 2 (32 bytes >> 4) per step.
template <bool A_TMEM>
__device__ __forceinline__ void umma_kblock_conv(uint32_t d_tmem, uint32_t a, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                                 uint32_t accumulate_first, uint32_t bar_a, uint32_t par_a, uint32_t need_a,
                                                 uint32_t bar_b, uint32_t par_b, uint32_t need_b, uint32_t bar_commit, uint32_t do_commit,
                                                 uint32_t bar_release, uint32_t do_release, uint32_t release_mask) {
#define DURF_KB_WAIT(Q, BAR, PAR, L)                                             \
      "@" Q " bra " L "_DONE;\n"                                                  \
      L "_WAIT:\n"                                                                \
      "mbarrier.try_wait.parity.shared::cta.b64 " Q ", [" BAR "], " PAR ";\n"     \
      "@" Q " bra " L "_DONE;\n"                                                  \
      "add.u32 cnt, cnt, 1;\n"                                                    \
      "setp.lt.u32 t, cnt, 0x1000000;\n"                                          \
      "@t bra " L "_WAIT;\n"                                                      \
      "trap;\n"                                                                   \
      L "_DONE:\n"
  // ring-stage release: the commit arrives on the stage's `empty` barrier once every MMA issued so far (in particular the
  // ones reading that stage) has completed; release_mask != 0 delivers it to the same barrier of every CTA in the mask
  // (CTA pairs fill their ring stages jointly by multicast)
#define DURF_KB_RELEASE                                                                                              \
      "setp.ne.u32 t, %15, 0;\n and.pred t, t, e;\n"                                                                 \
      "setp.ne.u32 m, %16, 0;\n and.pred u, t, m;\n"                                                                 \
      "@u tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%14], mk;\n"      \
      "not.pred m, m;\n and.pred u, t, m;\n"                                                                         \
      "@u tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%14];\n"
  if constexpr (A_TMEM) {
    asm volatile(
        "{\n"
        ".reg .pred e, pacc, ptrue, qa, qb, t, m, u;\n"
        ".reg .b64 db;\n"
        ".reg .b32 bl, al, cnt;\n"
        ".reg .b16 mk;\n"
        "cvt.u16.u32 mk, %16;\n"
        // a probe goes through the memory pipe of the issuing thread: only the barriers the next step really needs are probed
        // (need_x is a compile-time constant at most call sites, so the unused probe disappears altogether)
        "setp.ne.u32 t, %8, 0;\n setp.ne.u32 u, %11, 0;\n"
        "setp.ne.u32 qa, %8, %8;\n setp.ne.u32 qb, %8, %8;\n"
        "@t mbarrier.test_wait.parity.shared::cta.b64 qa, [%6], %7;\n"
        "@u mbarrier.test_wait.parity.shared::cta.b64 qb, [%9], %10;\n"
        "elect.sync _|e, 0xffffffff;\n"
        "setp.ne.b32 pacc, %5, 0;\n"
        "setp.eq.u32 ptrue, %5, %5;\n"
        "mov.b64 db, {%2, %3};\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, pacc;\n"
        "add.u32 bl, %2, 2;\n add.u32 al, %1, 8;\n mov.b64 db, {bl, %3};\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, ptrue;\n"
        "add.u32 bl, %2, 4;\n add.u32 al, %1, 16;\n mov.b64 db, {bl, %3};\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, ptrue;\n"
        "add.u32 bl, %2, 6;\n add.u32 al, %1, 24;\n mov.b64 db, {bl, %3};\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, ptrue;\n"
        "setp.ne.u32 t, %13, 0;\n and.pred t, t, e;\n"
        "@t tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%12];\n"
        DURF_KB_RELEASE
        "setp.eq.u32 t, %8, 0;\n or.pred qa, qa, t;\n"
        "setp.eq.u32 t, %11, 0;\n or.pred qb, qb, t;\n"
        "mov.u32 cnt, 0;\n"
        DURF_KB_WAIT("qa", "%6", "%7", "LA")
        DURF_KB_WAIT("qb", "%9", "%10", "LB")
        "}\n" ::"r"(d_tmem), "r"(a), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate_first), "r"(bar_a), "r"(par_a), "r"(need_a),
        "r"(bar_b), "r"(par_b), "r"(need_b), "r"(bar_commit), "r"(do_commit), "r"(bar_release), "r"(do_release), "r"(release_mask)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred e, pacc, ptrue, qa, qb, t, m, u;\n"
        ".reg .b64 da, db;\n"
        ".reg .b32 bl, al, cnt;\n"
        ".reg .b16 mk;\n"
        "cvt.u16.u32 mk, %16;\n"
        // a probe goes through the memory pipe of the issuing thread: only the barriers the next step really needs are probed
        // (need_x is a compile-time constant at most call sites, so the unused probe disappears altogether)
        "setp.ne.u32 t, %8, 0;\n setp.ne.u32 u, %11, 0;\n"
        "setp.ne.u32 qa, %8, %8;\n setp.ne.u32 qb, %8, %8;\n"
        "@t mbarrier.test_wait.parity.shared::cta.b64 qa, [%6], %7;\n"
        "@u mbarrier.test_wait.parity.shared::cta.b64 qb, [%9], %10;\n"
        "elect.sync _|e, 0xffffffff;\n"
        "setp.ne.b32 pacc, %5, 0;\n"
        "setp.eq.u32 ptrue, %5, %5;\n"
        "mov.b64 da, {%1, %3};\n mov.b64 db, {%2, %3};\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, pacc;\n"
        "add.u32 bl, %2, 2;\n add.u32 al, %1, 2;\n mov.b64 da, {al, %3};\n mov.b64 db, {bl, %3};\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, ptrue;\n"
        "add.u32 bl, %2, 4;\n add.u32 al, %1, 4;\n mov.b64 da, {al, %3};\n mov.b64 db, {bl, %3};\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, ptrue;\n"
        "add.u32 bl, %2, 6;\n add.u32 al, %1, 6;\n mov.b64 da, {al, %3};\n mov.b64 db, {bl, %3};\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, ptrue;\n"
        "setp.ne.u32 t, %13, 0;\n and.pred t, t, e;\n"
        "@t tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%12];\n"
        DURF_KB_RELEASE
        "setp.eq.u32 t, %8, 0;\n or.pred qa, qa, t;\n"
        "setp.eq.u32 t, %11, 0;\n or.pred qb, qb, t;\n"
        "mov.u32 cnt, 0;\n"
        DURF_KB_WAIT("qa", "%6", "%7", "LA")
        DURF_KB_WAIT("qb", "%9", "%10", "LB")
        "}\n" ::"r"(d_tmem), "r"(a), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate_first), "r"(bar_a), "r"(par_a), "r"(need_a),
        "r"(bar_b), "r"(par_b), "r"(need_b), "r"(bar_commit), "r"(do_commit), "r"(bar_release), "r"(do_release), "r"(release_mask)
        : "memory");
  }
#undef DURF_KB_WAIT
#undef DURF_KB_RELEASE
}
// One ring stage of the weight-gradient kernel (mlp_tc_wgrad.cu) by a converged warp: 4 (m_tiles == 1) or 8 K=16 MMAs whose
// MN-major operands advance by 2048 bytes per K step, the commit that frees the stage, and - software-pipelined as in
// umma_kblock_conv - the wait for the NEXT stage's operands: probed before the MMAs, consumed after them, so the issuing
// thread never sits in a barrier wait with an empty tensor queue behind it.  d1 / a1_lo: accumulator and A descriptor of the
// second M tile.
__device__ __forceinline__ void umma_stage_mn(uint32_t d0, uint32_t d1, uint32_t a0_lo, uint32_t a1_lo, uint32_t b_lo, uint32_t desc_hi,
                                              uint32_t idesc, uint32_t accumulate_first, uint32_t two_m_tiles, uint32_t bar_release,
                                              uint32_t bar_next, uint32_t par_next, uint32_t need_next) {
  asm volatile(
      "{\n"
      ".reg .pred e, pacc, ptrue, q, t, m2;\n"
      ".reg .b64 da, db;\n"
      ".reg .b32 al, bl, cnt;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 q, [%10], %11;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "setp.ne.b32 pacc, %7, 0;\n"
      "setp.eq.u32 ptrue, %7, %7;\n"
      "setp.ne.b32 m2, %8, 0;\n"
      "and.pred m2, m2, e;\n"
      "mov.b64 da, {%2, %5};\n mov.b64 db, {%4, %5};\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %6, pacc;\n"
      "add.u32 al, %2, 128;\n add.u32 bl, %4, 128;\n mov.b64 da, {al, %5};\n mov.b64 db, {bl, %5};\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %6, ptrue;\n"
      "add.u32 al, %2, 256;\n add.u32 bl, %4, 256;\n mov.b64 da, {al, %5};\n mov.b64 db, {bl, %5};\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %6, ptrue;\n"
      "add.u32 al, %2, 384;\n add.u32 bl, %4, 384;\n mov.b64 da, {al, %5};\n mov.b64 db, {bl, %5};\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %6, ptrue;\n"
      "mov.b64 da, {%3, %5};\n mov.b64 db, {%4, %5};\n"
      "@m2 tcgen05.mma.cta_group::1.kind::f16 [%1], da, db, %6, pacc;\n"
      "add.u32 al, %3, 128;\n add.u32 bl, %4, 128;\n mov.b64 da, {al, %5};\n mov.b64 db, {bl, %5};\n"
      "@m2 tcgen05.mma.cta_group::1.kind::f16 [%1], da, db, %6, ptrue;\n"
      "add.u32 al, %3, 256;\n add.u32 bl, %4, 256;\n mov.b64 da, {al, %5};\n mov.b64 db, {bl, %5};\n"
      "@m2 tcgen05.mma.cta_group::1.kind::f16 [%1], da, db, %6, ptrue;\n"
      "add.u32 al, %3, 384;\n add.u32 bl, %4, 384;\n mov.b64 da, {al, %5};\n mov.b64 db, {bl, %5};\n"
      "@m2 tcgen05.mma.cta_group::1.kind::f16 [%1], da, db, %6, ptrue;\n"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%9];\n"
      "setp.eq.u32 t, %12, 0;\n or.pred q, q, t;\n"
      "mov.u32 cnt, 0;\n"
      "@q bra LN_DONE;\n"
      "LN_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 q, [%10], %11;\n"
      "@q bra LN_DONE;\n"
      "add.u32 cnt, cnt, 1;\n"
      "setp.lt.u32 t, cnt, 0x4000000;\n"
      "@t bra LN_WAIT;\n"
      "trap;\n"
      "LN_DONE:\n"
      "}\n" ::"r"(d0), "r"(d1), "r"(a0_lo), "r"(a1_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate_first), "r"(two_m_tiles),
      "r"(bar_release), "r"(bar_next), "r"(par_next), "r"(need_next)
      : "memory");
}
__device__ __forceinline__ void tc_commit_conv(uint32_t bar) {
  asm volatile(
      "{\n"
      ".reg .pred e;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Ties the loaded registers to a point after tcgen05.wait::ld (volatile asms keep their order), so no consumer of
// the asynchronous load can be scheduled above the wait.  Emits no instruction.
__device__ __forceinline__ void tmem_ld_pin(uint32_t (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 32; i += 16)
    asm volatile("" : "+r"(v[i]), "+r"(v[i + 1]), "+r"(v[i + 2]), "+r"(v[i + 3]), "+r"(v[i + 4]), "+r"(v[i + 5]), "+r"(v[i + 6]),
                      "+r"(v[i + 7]), "+r"(v[i + 8]), "+r"(v[i + 9]), "+r"(v[i + 10]), "+r"(v[i + 11]), "+r"(v[i + 12]),
                      "+r"(v[i + 13]), "+r"(v[i + 14]), "+r"(v[i + 15]));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 16-byte load from shared memory by 32-bit shared address.  Not volatile and no memory clobber: the tables read
// with it (biases, head weights) are written once before the role dispatch, so the compiler may batch and hoist them.
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 r;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
  return r;
}

__device__ __forceinline__ float2 lds64(uint32_t addr) {
  float2 r;
  asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(addr));
  return r;
}
// Same load for data that changes during the kernel (the per-ray view bias): volatile keeps it ordered with the
// named barriers that publish it.
__device__ __forceinline__ float4 lds128_volatile(uint32_t addr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr) : "memory");
  return r;
}

// packed fp32 add (FADD2) and fp32 pair -> bf16x2 conversions (lo = first argument)
__device__ __forceinline__ float2 add2(float a0, float a1, float2 b) {
  float2 a = make_float2(a0, a1), r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long*>(&r))
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return r;
}
__device__ __forceinline__ uint32_t cvt_bf16x2_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// Shared-memory matrix descriptor, K-major, SWIZZLE_128B (cute::UMMA::SmemDescriptor): lo = start>>4 | LBO(=1, unused for
// swizzled K-major)<<16; hi = SBO (1024 B between 8-row groups)>>4 | version 1 <<14 | layout SWIZZLE_128B (2) <<29.
// Instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 (1<<4), a=b=BF16 (1<<7, 1<<10), K-major both,
// N>>3 at bit 17, M>>4 at bit 24.
__host__ __device__ constexpr uint32_t umma_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}


// MN-major variant (the 128-byte rows run along M or N, the 8-row groups along K): lo = start>>4 | LBO>>4 <<16 where
// LBO = byte distance between 64-element blocks along M/N; hi as above (SBO = 1024 B between 8-row K groups).
__host__ __device__ constexpr uint32_t umma_idesc_mn(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

}  // namespace durf
