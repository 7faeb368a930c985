// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05
// (MMA / TMEM alloc / ld / st / commit) and the fences between the generic and async proxies.
// Everything here is hand-written for Blackwell; there is no other backend.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace xf {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---------------------------------------------------------------- proxies / fences
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// TMA stores with batch coordinates (3-D / 4-D tensor maps)
__device__ __forceinline__ void tma_store_3d(const void* tmap, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(tmap)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void* tmap, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(reinterpret_cast<uint64_t>(tmap)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// ---------------------------------------------------------------- 2-CTA (cta_group::2) variants
// A CTA pair (cluster of 2 on one TPC) runs one M=256 MMA: each CTA stages its own 128 rows of A and
// half of the B tile; the even CTA issues the MMA and both CTAs' barriers are signalled by multicast.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> even CTA

// TMA store (shared -> global, bulk async group of the issuing thread).  The shared-memory writes that fill the
// box must be followed by fence_proxy_async_smem() (generic -> async proxy) before the store is issued.
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(tmap)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// waits until at most N of this thread's most recent bulk groups still READ their shared-memory source
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes are credited to the EVEN CTA's mbarrier (same smem offset)
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_cg2(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_even_cta(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_cg2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_cg2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier at this smem offset in every CTA of cta_mask once the prior MMAs completed
__device__ __forceinline__ void umma_commit_cg2_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask) : "memory");
}

// ---------------------------------------------------------------- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]; one thread issues for the whole CTA.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Issue-rate matters: one thread feeds the tensor pipe and a short MMA (N <= 64) retires every ~46 cycles
// (measured, tools/microbench/mma_bench.cu), so the helpers below issue 2 or 4 MMAs of one K-chain from a
// single asm block.  Descriptors are passed as (hi, lo) words; only the start-address field (lo) advances.
#define XF_UMMA_CHAIN_BODY(CG)                                                                   \
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\t"                        \
      "setp.ne.b32 p, %8, 0;\n\t"                                                                 \
      "setp.eq.b32 q, 0, 0;\n\t"                                                                  \
      "mov.b64 da, {%2, %1};\n\tmov.b64 db, {%5, %4};\n\t"                                        \
      "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %7, p;\n\t"                          \
      "add.u32 al, %2, %3;\n\tadd.u32 bl, %5, %6;\n\t"                                            \
      "mov.b64 da, {al, %1};\n\tmov.b64 db, {bl, %4};\n\t"                                        \
      "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %7, q;\n\t"
#define XF_UMMA_CHAIN_MORE(CG)                                                                   \
      "add.u32 al, al, %3;\n\tadd.u32 bl, bl, %6;\n\t"                                            \
      "mov.b64 da, {al, %1};\n\tmov.b64 db, {bl, %4};\n\t"                                        \
      "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %7, q;\n\t"

// D (+)= sum_{k<2} A_k B_k
__device__ __forceinline__ void umma_k2(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t a_step, uint32_t b_hi,
                                        uint32_t b_lo, uint32_t b_step, uint32_t idesc, uint32_t accumulate_first) {
  asm volatile(XF_UMMA_CHAIN_BODY("1") "}\n"
               ::"r"(d_tmem), "r"(a_hi), "r"(a_lo), "r"(a_step), "r"(b_hi), "r"(b_lo), "r"(b_step), "r"(idesc),
                 "r"(accumulate_first)
               : "memory");
}
// D (+)= sum_{k<4} A_k B_k
__device__ __forceinline__ void umma_k4(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t a_step, uint32_t b_hi,
                                        uint32_t b_lo, uint32_t b_step, uint32_t idesc, uint32_t accumulate_first) {
  asm volatile(XF_UMMA_CHAIN_BODY("1") XF_UMMA_CHAIN_MORE("1") XF_UMMA_CHAIN_MORE("1") "}\n"
               ::"r"(d_tmem), "r"(a_hi), "r"(a_lo), "r"(a_step), "r"(b_hi), "r"(b_lo), "r"(b_step), "r"(idesc),
                 "r"(accumulate_first)
               : "memory");
}
__device__ __forceinline__ void umma_k4_cg2(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t a_step, uint32_t b_hi,
                                            uint32_t b_lo, uint32_t b_step, uint32_t idesc, uint32_t accumulate_first) {
  asm volatile(XF_UMMA_CHAIN_BODY("2") XF_UMMA_CHAIN_MORE("2") XF_UMMA_CHAIN_MORE("2") "}\n"
               ::"r"(d_tmem), "r"(a_hi), "r"(a_lo), "r"(a_step), "r"(b_hi), "r"(b_lo), "r"(b_step), "r"(idesc),
                 "r"(accumulate_first)
               : "memory");
}
// A operand from TMEM (".ts" form): 128 lanes x 8 columns per K = 16 slice (bf16 packed two per column, element
// 2c in the low half; verified by tools/microbench/ts_mma_test.cu).  Saves the 4 KB shared-memory read of A per MMA.
#define XF_UMMA_TS_BODY                                                                           \
      "{\n\t.reg .pred p, q;\n\t.reg .b64 db;\n\t.reg .b32 at, bl;\n\t"                             \
      "setp.ne.b32 p, %6, 0;\n\t"                                                                 \
      "setp.eq.b32 q, 0, 0;\n\t"                                                                  \
      "mov.b64 db, {%3, %2};\n\t"                                                                 \
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %5, p;\n\t"                             \
      "add.u32 at, %1, 8;\n\tadd.u32 bl, %3, %4;\n\tmov.b64 db, {bl, %2};\n\t"                    \
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [at], db, %5, q;\n\t"
#define XF_UMMA_TS_MORE                                                                           \
      "add.u32 at, at, 8;\n\tadd.u32 bl, bl, %4;\n\tmov.b64 db, {bl, %2};\n\t"                    \
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [at], db, %5, q;\n\t"
__device__ __forceinline__ void umma_ts_k2(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_hi, uint32_t b_lo, uint32_t b_step,
                                           uint32_t idesc, uint32_t accumulate_first) {
  asm volatile(XF_UMMA_TS_BODY "}\n" ::"r"(d_tmem), "r"(a_tmem), "r"(b_hi), "r"(b_lo), "r"(b_step), "r"(idesc),
               "r"(accumulate_first) : "memory");
}
__device__ __forceinline__ void umma_ts_k4(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_hi, uint32_t b_lo, uint32_t b_step,
                                           uint32_t idesc, uint32_t accumulate_first) {
  asm volatile(XF_UMMA_TS_BODY XF_UMMA_TS_MORE XF_UMMA_TS_MORE "}\n" ::"r"(d_tmem), "r"(a_tmem), "r"(b_hi), "r"(b_lo),
               "r"(b_step), "r"(idesc), "r"(accumulate_first) : "memory");
}
// shared memory (UMMA descriptor, 128 rows x 32 B) -> TMEM (128 lanes x 8 columns); ordered with later tcgen05.mma
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

__device__ __forceinline__ uint32_t desc_hi(uint64_t d) { return static_cast<uint32_t>(d >> 32); }
__device__ __forceinline__ uint32_t desc_lo(uint64_t d) { return static_cast<uint32_t>(d); }

// Arrives (once) on the mbarrier when all previously issued tcgen05.mma of this thread completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (base+i), columns c..c+31.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor, SWIZZLE_128B, 1024-byte aligned atoms (cute::UMMA::SmemDescriptor
// bit layout: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout [61,64)).
//   K-major : rows of 128 B (64 bf16 along K); 8-row groups SBO apart; K advance = +32 B per UMMA_K.
//   MN-major: 64-element MN chunks (128 B) x 8 K-rows = 1024 B atoms; SBO = stride between 8-row K
//             groups, LBO = stride between 64-element MN chunks; K advance = +2*SBO per UMMA_K.
//   SWIZZLE_64B (layout code 4): same forms with 64-byte rows (32 bf16), 512-byte atoms.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout = 2) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(layout) << 61;  // 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
  return d;
}

// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32, M = 128 (cute::UMMA::InstrDescriptor).
__host__ __device__ __forceinline__ uint32_t make_idesc_bf16(uint32_t n, uint32_t a_mn_major, uint32_t b_mn_major,
                                                             uint32_t m = 128) {
  uint32_t d = 0;
  d |= 1u << 4;                  // D format F32
  d |= 1u << 7;                  // A format BF16
  d |= 1u << 10;                 // B format BF16
  d |= (a_mn_major & 1u) << 15;  // A major: 0 = K, 1 = MN
  d |= (b_mn_major & 1u) << 16;  // B major
  d |= ((n >> 3) & 0x3F) << 17;  // N >> 3
  d |= (m >> 4) << 24;           // M >> 4 (128, or 256 with cta_group::2)
  return d;
}

// ---------------------------------------------------------------- small math helpers
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

// erf via Abramowitz-Stegun 7.1.26 (|abs err| < 1.5e-7, far below bf16 resolution): one ex2 + one rcp
// instead of the ~30-instruction erff; e = exp(-z^2) is returned for reuse (GELU' needs the same term).
__device__ __forceinline__ float fast_erf_pos(float z, float& e) {  // z >= 0
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  e = __expf(-z * z);
  return 1.0f - poly * t * e;
}
__device__ __forceinline__ float gelu_erf(float x) {
  float e;
  const float er = fast_erf_pos(fabsf(x) * 0.70710678118654752f, e);
  return 0.5f * x * (1.0f + copysignf(er, x));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float e;  // e = exp(-x^2/2)
  const float er = fast_erf_pos(fabsf(x) * 0.70710678118654752f, e);
  const float cdf = 0.5f * (1.0f + copysignf(er, x));
  return fmaf(x * 0.39894228040143268f, e, cdf);
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// GELU(erf) on two values with packed fp32 math (fma.rn.f32x2): 0.5 erfc(|x| / sqrt 2) = exp2(P6(min(|x|, 6))),
// P6 fitted in the log2 domain (max exponent error 6.5e-5 -> relative error 4.5e-5 of the tail probability,
// |GELU error| <= 7e-6, |GELU' error| <= 2.3e-5; both far below bf16 resolution).  One exp2 per value, no divide.
//   GELU(x)  = max(x, 0) - |x| q(|x|)            GELU'(x) = Phi(x) + x phi(x),  Phi = x >= 0 ? 1 - q : q
__device__ __forceinline__ float2 gelu_tailprob2(float2 x, float2& a) {
  a = make_float2(fminf(fabsf(x.x), 6.0f), fminf(fabsf(x.y), 6.0f));
  float2 r = make_float2(2.29900633712532e-05f, 2.29900633712532e-05f);
  r = __ffma2_rn(r, a, make_float2(-0.0006111000548116863f, -0.0006111000548116863f));
  r = __ffma2_rn(r, a, make_float2(0.007195565849542618f, 0.007195565849542618f));
  r = __ffma2_rn(r, a, make_float2(-0.05118533596396446f, -0.05118533596396446f));
  r = __ffma2_rn(r, a, make_float2(-0.46127191185951233f, -0.46127191185951233f));
  r = __ffma2_rn(r, a, make_float2(-1.1501742601394653f, -1.1501742601394653f));
  r = __ffma2_rn(r, a, make_float2(-1.000064730644226f, -1.000064730644226f));
  return make_float2(fast_exp2(r.x), fast_exp2(r.y));
}
__device__ __forceinline__ float2 gelu2(float2 x) {
  float2 a;
  const float2 q = gelu_tailprob2(x, a);
  return __ffma2_rn(make_float2(-a.x, -a.y), q, make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)));
}
__device__ __forceinline__ float2 gelu_grad2(float2 x) {
  float2 a;
  const float2 q = gelu_tailprob2(x, a);
  const float2 h = __fadd2_rn(make_float2(0.5f, 0.5f), make_float2(-q.x, -q.y));       // 0.5 - q >= 0
  const float2 cdf = __fadd2_rn(make_float2(copysignf(h.x, x.x), copysignf(h.y, x.y)), make_float2(0.5f, 0.5f));
  // x phi(x) = x exp2(-x^2 log2(e)/2 - log2 sqrt(2 pi))
  const float2 w = __ffma2_rn(__fmul2_rn(x, x), make_float2(-0.72134752044448170f, -0.72134752044448170f),
                              make_float2(-1.32574806473616f, -1.32574806473616f));
  return __ffma2_rn(x, make_float2(fast_exp2(w.x), fast_exp2(w.y)), cdf);
}

// Counter-based dropout (no mask storage; the backward recomputes the same bits).
//   keep(row, col) = rowhash(key, row) * colhash(col) >= t32,   t32 = round(p * 65536) << 16
// rowhash is a full avalanche hash of (site key, row); colhash is an ODD avalanche hash of the column, so for a
// fixed column the product is a bijection of the row hash.  A thread that owns a row hashes it once; the column
// hashes are tile / table constants (shared-memory broadcasts in the tensor-core kernels, a small global table
// in the element-wise ones).  Per element: one IMAD, one unsigned compare, one select.
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t drop_key(uint32_t seed, uint32_t stream) {
  return mix32(seed ^ (stream * 0x9E3779B9U) ^ 0x5bd1e995U);
}
__host__ __device__ __forceinline__ uint32_t drop_thresh32(float p) {
  uint32_t t16 = static_cast<uint32_t>(static_cast<double>(p) * 65536.0 + 0.5);
  if (t16 > 65535u) t16 = 65535u;
  return t16 << 16;
}
__device__ __forceinline__ uint32_t drop_rowhash(uint32_t key, uint64_t row) {
  return mix32(key ^ static_cast<uint32_t>(row) ^ (static_cast<uint32_t>(row >> 32) * 0x85EBCA6BU));
}
// GEMM / LayerNorm sites: key-independent odd column hash (xf::drop_col_table() holds the first 16384 of them)
__host__ __device__ __forceinline__ uint32_t drop_colodd(uint32_t col) {
  return mix32(col * 0x9E3779B9U + 0x7F4A7C15U) | 1u;
}
// Attention probabilities: keep(q, k) = rowhash(q) * colhash(k) >= t32, symmetric in (query, key) so either
// orientation hoists the hashes: threads that own a query row hash the tile's keys once per warp (one per lane)
// and re-read them as shared-memory broadcasts; key-owning threads do the converse.
__device__ __forceinline__ uint32_t drop_colhash(uint32_t key, uint32_t col) {
  return mix32(key ^ 0xA511E9B3U ^ (col * 0x85EBCA6BU)) | 1u;
}
__device__ __forceinline__ bool drop_keep_rc(uint32_t rowhash, uint32_t colhash, uint32_t t32) {
  return rowhash * colhash >= t32;
}

}  // namespace xf
