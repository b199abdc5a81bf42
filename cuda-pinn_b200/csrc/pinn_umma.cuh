// pinn_umma.cuh — hand-written sm_100a tensor-core plumbing: tcgen05.mma (UMMA) with TMEM
// accumulators, mbarrier completion, shared-memory matrix descriptors.  Inline PTX only.
//
// Shared-memory operand layout used everywhere in this project (no swizzle, "interleaved"
// canonical UMMA layout): a logical matrix X[r, f] (r = row of the point tile, f = feature) is
// stored as 16-byte chunks of 8 consecutive bf16 features:
//
//        byte address(r, f) = (f / 8) * (R * 16) + r * 16 + (f % 8) * 2          ("chunk-major")
//
// so 8 consecutive rows of one chunk form one 128-byte UMMA core matrix.  The same bytes serve
//   * as a K-major operand   when the contraction runs over f  (Z = A W,  dA = Zbar W^T):
//         leading-dimension byte offset LBO = R*16  (next 16-byte chunk along K)
//         stride-dimension  byte offset SBO = 128   (next group of 8 rows along M/N)
//   * as an MN-major operand when the contraction runs over r  (dW = A^T Zbar):
//         SBO = R*16 (next chunk of 8 features along M/N),  LBO = 128 (next group of 8 rows along K)
// which is why one copy of A, one of Zbar and one of W feed all three contractions of a layer.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace pinn {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// Bounded: a phase that has not completed after 2^24 polls (>= 0.2 s even if a poll returned at once; normal waits are
// microseconds) is a protocol failure — trap, so the launch fails loudly (CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done, polls = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (!done && ++polls > (1u << 24)) __trap();
    } while (!done);
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// ---- TMA 1-D bulk copy global -> shared (cp.async.bulk, SASS UBLKCP), completes on an mbarrier ----
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- TMA 1-D bulk copy shared -> global (bulk async-group completion) ----
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }   // sources reusable
__device__ __forceinline__ void bulk_wait_all_but_last() { asm volatile("cp.async.bulk.wait_group 1;" ::: "memory"); }   // every group but the newest has landed
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }             // writes complete

// generic-proxy smem writes -> visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM -----------------------------------------------------------------------------
// One full warp allocates `ncols` (power of two >= 32) columns; the base address lands in *slot.
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 8 consecutive 32-bit columns: thread t of the warp receives lane (lane_base + t).
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    #pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors ----------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_NONE, Blackwell version field = 1.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);                 // [ 0,14) start address >> 4
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;        // [16,30) leading-dimension byte offset >> 4
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;        // [32,46) stride-dimension byte offset >> 4
    d |= (uint64_t)1 << 46;                                   // [46,48) descriptor version (sm_100)
    return d;                                                 // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
// Instruction descriptor for kind::f16 with BF16 operands and FP32 accumulation.
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// Make the mbarrier track completion of all MMAs issued so far by this thread.
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- lean issue path (tc3) ------------------------------------------------------------------
// One UMMA per ~16 SASS instructions (descriptor rebuilt from the address, an ELECT loop around every tcgen05.mma because the
// compiler cannot see that a `tid == k` branch holds one lane) makes a single issuing thread the bottleneck: ~94 cycles per UMMA
// against 69 of execution.  Here the descriptor words are built once per contraction and stepped by an add, and the issuing lane
// comes from elect.sync inside a warp-uniform branch.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14); }
__device__ __forceinline__ void mma_bf16_words(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
template <int KSTEPS>
__device__ __forceinline__ void gemm_issue_lean(uint32_t d_tmem, uint32_t a_addr, uint32_t a_lbo, uint32_t a_sbo, uint32_t a_kstep,
                                                uint32_t b_addr, uint32_t b_lbo, uint32_t b_sbo, uint32_t b_kstep,
                                                uint32_t idesc, uint32_t accumulate_first) {
    const uint32_t a_lo = desc_lo(a_addr, a_lbo), a_hi = desc_hi(a_sbo), b_lo = desc_lo(b_addr, b_lbo), b_hi = desc_hi(b_sbo);
    #pragma unroll
    for (int s = 0; s < KSTEPS; s++)
        mma_bf16_words(d_tmem, a_lo + (uint32_t)s * (a_kstep >> 4), a_hi, b_lo + (uint32_t)s * (b_kstep >> 4), b_hi, idesc,
                       (s > 0) ? 1u : accumulate_first);
}

// A contraction over `ksteps` UMMA K-steps (16 bf16 each): both operands advance by a fixed byte step (run-time step count; the
// descriptor words are built once and stepped by an add, as in gemm_issue_lean).
__device__ __forceinline__ void gemm_issue(uint32_t d_tmem, uint32_t a_addr, uint32_t a_lbo, uint32_t a_sbo, uint32_t a_kstep,
                                           uint32_t b_addr, uint32_t b_lbo, uint32_t b_sbo, uint32_t b_kstep,
                                           uint32_t idesc, int ksteps, uint32_t accumulate_first) {
    uint32_t a_lo = desc_lo(a_addr, a_lbo), b_lo = desc_lo(b_addr, b_lbo);
    const uint32_t a_hi = desc_hi(a_sbo), b_hi = desc_hi(b_sbo), a_inc = a_kstep >> 4, b_inc = b_kstep >> 4;
    for (int s = 0; s < ksteps; s++, a_lo += a_inc, b_lo += b_inc)
        mma_bf16_words(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, (s > 0) ? 1u : accumulate_first);
}

// byte offset of element (r, f) of an [R x F] bf16 matrix in the chunk-major layout above
__host__ __device__ __forceinline__ uint32_t cm_offset(int r, int f, int R) { return (uint32_t)((f >> 3) * (R * 16) + r * 16 + (f & 7) * 2); }

}  // namespace umma
}  // namespace pinn
