// pinn_fused_tc2.cuh — round-2 tensor-core train step for the wide networks (width 128 / 256, six Taylor
// channels: the Allen-Cahn / heat / Navier-Stokes nets of BASELINE C4 / C5), interior set, BF16 operands
// with FP32 accumulation in TMEM.  One persistent cooperative launch, ONE CTA per SM, two roles:
//
//   PRODUCER CTAs (blockIdx < nprod) walk 21-point tiles through the whole step.  The contractions are
//   "units as rows":            Z_l^T [j, n] = W_l^T [j, k] . A_{l-1}^T [k, n]        (forward)
//                               Abar_{l-1}^T [k, n] = W_l [k, j] . Zbar_l^T [j, n]    (reverse)
//   with n = 6 * point + channel (126 of the 128 operand columns carry data), M = 128 units per UMMA
//   (two M-tiles at width 256), N = 128, K = the width.  TMEM lane = unit, so ONE THREAD owns every Taylor
//   channel of every point of its unit: the tanh / Taylor forward, the reverse adjoint, the bias gradient,
//   dW_1 and dW_{L+1} are thread-local straight out of tcgen05.ld — no shared-memory transposition, no
//   shuffles, no block reduction.  A thread writes its results as 16-byte chunks of the NEXT operand
//   ([n/8][unit][8 n] bf16: MN-major B operand of the forward / reverse UMMAs, K-major operand of dW).
//   The stash for the reverse pass is that operand image itself: the reverse formulas are written in the
//   layer OUTPUTS (a, a', a''),
//        zbar''_i = s abar''_i          zbar'_i = s abar'_i - 4 a a'_i abar''_i
//        zbar     = s abar - sum_i 2 a a'_i abar'_i - sum_i (2 a a''_i + 2 a'_i^2) abar''_i
//   so the stash is a copy of the 64 KB image (plus s = 1 - a^2 in bf16, whose relative accuracy bf16(a) cannot
//   give): 12 + 2 bytes per (point, unit, layer) instead of 24 of FP32 (z', z'').  Every thread stores its own
//   16-byte chunks to shared memory (next UMMA operand) AND to global memory (stash / ring); the one-bulk-store-per-
//   image alternative is kept behind PINN_TC2_BULK=1 and measures slower.
//
//   dW SERVER CTAs (blockIdx >= nprod; q per hidden->hidden layer) never see a point: a 256 x 256 FP32 dW_l
//   is all of TMEM, so the producers cannot hold it next to their own accumulators, and flushing it per tile
//   costs 12 KB of L2 atomics per point-layer.  Instead server (l, i) streams the two dW_l operand images
//   (A_{l-1} from the stash, Zbar_l from a small ring) of its producers' tiles through a 3-stage
//   cp.async.bulk pipeline and accumulates dW_l = A_{l-1}^T Zbar_l in TMEM over the WHOLE step; it is
//   drained once, into the CTA's own partial row.  Hand-off is a pair of global flags per
//   (producer, layer) (release / acquire at gpu scope); servers take their producers in a fixed order, so
//   the summation order — and every bit of the result — is the same on every run.
//
// Reference: the reference has no network code (include/network.cuh is empty); the building blocks this
// replaces are tensor::matmul (include/tensor.cuh:380-409) + kernel::matmul (include/device/kernel.cuh:68-83).
#pragma once
#include "pinn_device.cuh"
#include "pinn_fused_fp32.cuh"
#include "pinn_umma.cuh"

namespace pinn {

struct Tc2Args {
    const float* params;
    PointSet set;                  // interior points; set.ntiles = ceil(n / 21)
    float* partial;                // [rows][PP]; this launch uses rows row0 + blockIdx.x
    int row0;
    const __nv_bfloat16* wf;       // forward mirror  [L-1][H/128][16 j-groups][H k][8 j]   (A operand, MN-major)
    const __nv_bfloat16* wb;       // reverse mirror  [L-1][H/128][H/8 j-groups][128 k][8 j] (A operand, K-major)
    uint8_t* stash_a;              // [nprod][L][H*256 B]   operand images A_1..A_L of the tile in flight
    uint8_t* stash_s;              // [nprod][L][6][H][4] bf16   s = 1 - a^2
    uint8_t* zring;                // [nprod][zdepth][H*256 B]  Zbar images on their way to the dW servers
    unsigned int* ready;           // [nprod][kFlagStride]  ready[p][l] = t+1: Zbar_l (and A_{l-1}) of tile t are in memory
    unsigned int* ready2;          // tc3 only (else null): the same flag for the second sub-tile's half of the images; the server waits for both
    unsigned int* done;            // [nprod][kFlagStride]  done[p][l]  = t+1: server has both operands of tile t on chip
    int nprod, q, zdepth;
    int bulk_store;                // tc2 (PINN_TC2_BULK=1): stash / ring images leave shared memory by ONE cp.async.bulk per layer instead of 8 x 16-byte st.global per thread
    int stagger;                   // tc3: sub-tile B issues a layer step once A has issued this many of its weight pieces (0 = free-running)
    long long* dbg;                // optional [8]: cycle counters of producer 0 (PINN_TC2_DBG=1), or null
    int* trap_site;                // mapped host memory [4]: {site, block, a, b} of a wait that timed out (see tc2::trap_at)
};
constexpr int kFlagStride = kMaxLayers + 2;

namespace tc2 {

constexpr int T = 21, G = 4, NGRP = 6;     // points per tile, per group; groups per tile (the last holds one point)
constexpr int OUTCOL = 256;                // TMEM columns of the output-layer accumulator
// single-thread roles next to the epilogue work every thread does (one per warp, so their waits do not serialise)
constexpr int kLoader = 128, kPublisher = 64, kWaiter = 96;   // (the loader sits in a warp of M-tile 1: it sees the end of the whole step first)

// tanh and s = 1 - tanh^2 from one exp2 and one reciprocal (absolute error ~2e-7; s keeps its RELATIVE accuracy
// for saturated units, which 1 - a*a cannot).
__device__ __forceinline__ void tanh_s(float z, float& a, float& s) {
    z = fminf(fmaxf(z, -15.0f), 15.0f);
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * 2.8853900817779268f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    a = fmaf(-2.0f, r, 1.0f);
    s = 4.0f * e * r * r;
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// A protocol failure (a hand-off that never happens) must fail loudly, and say where: the wait site goes to mapped host
// memory (it survives the trap), then the launch is killed.  pinn_last_error() reports it.
__device__ long long g_wait_bound = 1LL << 30;      // cycles a hand-off may take (~0.5 s; PINN_TC2_TIMEOUT_SHIFT raises it for profilers)
__device__ __noinline__ void trap_at(int* where, int site, int a, int b) {
    if (where) {
        where[1] = (int)blockIdx.x; where[2] = a; where[3] = b;
        __threadfence_system();
        where[0] = site;
        __threadfence_system();
    }
    __trap();
}
__device__ __forceinline__ void mbar_wait_at(uint64_t* bar, uint32_t parity, int* where, int site) {
    const uint32_t addr = umma::smem_u32(bar);
    uint32_t done, polls = 0;
    long long t0 = 0;
    do {
        // (suspend-time hint: the warp sleeps in hardware until the phase completes or ~10 us pass — far fewer issue slots taken
        //  from the warps that are computing than a tight poll)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity), "r"(10000u) : "memory");
        if (!done && (++polls & 63u) == 0u) {
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > g_wait_bound) trap_at(where, site, (int)parity, 0);
        }
    } while (!done);
}
__device__ __forceinline__ void wait_ge_at(const unsigned* p, unsigned v, int* where, int site) {
    if (ld_acquire(p) >= v) return;
    const long long t0 = clock64();
    while (ld_acquire(p) < v) {
        __nanosleep(64);
        if (clock64() - t0 > 4 * g_wait_bound) trap_at(where, site, (int)v, (int)ld_acquire(p));
    }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
// tcgen05.wait::ld that names the registers it guards, so no use of them can be scheduled above it
__device__ __forceinline__ void tmem_wait24(uint32_t* r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23])
                 :: "memory");
}
__device__ __forceinline__ void tmem_wait16(uint32_t* r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    #pragma unroll
    for (int i = 0; i < 4; i++) v[i] = __uint_as_float(r[i]);
}
// ---- small TMEM stores / loads: operand words parked in the accumulator columns their own thread has just consumed ----
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld4_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait12(uint32_t* r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]) :: "memory");
}
__device__ __forceinline__ void tmem_wait4(uint32_t* r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]) :: "memory");
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// tiles of producer p when `ntiles` tiles are dealt round-robin to `nprod` producers
__host__ __device__ __forceinline__ int tiles_of_producer(int ntiles, int p, int nprod) {
    return p < ntiles ? (ntiles - p + nprod - 1) / nprod : 0;
}

}  // namespace tc2

// Dynamic shared memory (bytes) the kernel needs for width H and depth L.
inline int tc2_smem_bytes(int H, int L) {
    const int slab = 128 * H * 2, opb = H * 256;
    const int producer = 2 * slab + opb + H * 32 + 2 * L * H * 4 + 2 * 128 * 16 + 32 * 16;
    const int server = 3 * opb;
    return producer > server ? producer : server;
}

// =====================================================================================
// dW server CTA (shared by the tc2 and tc3 producer kernels): dW_l = sum over its producers' tiles of A_{l-1}^T Zbar_l,
// accumulated in TMEM over the whole step, drained once into the CTA's own partial row.
// =====================================================================================
template <int HT>
__device__ __forceinline__ void dw_server(const NetDims& nd, const Tc2Args& ta, uint8_t* smem, uint64_t* full_bar, uint64_t* empty_bar,
                                          uint64_t* fin_bar_p, uint32_t tmem, float* part, uint32_t tmem_cols = 512u) {
    using namespace tc2;
    constexpr int H = HT, MT = H / 128;
    constexpr uint32_t OPB = (uint32_t)H * 256u, CHS = (uint32_t)H * 16u;
    const int L = nd.L;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q4 = warp & 3, mt = (warp >> 2) % MT, half = (warp >> 2) / MT;
    const int j = mt * 128 + q4 * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    const int ntiles = ta.set.ntiles, P = ta.nprod;
    uint64_t& fin_bar = *fin_bar_p;
    const int ci = (int)blockIdx.x - P, l = 2 + ci / ta.q, i0 = ci % ta.q;
    constexpr uint32_t HALF = OPB / 2;              // 8 n-chunks of both operands per pipeline stage
    bool any_work = false;
    for (int p = i0; p < P; p += ta.q) any_work |= tiles_of_producer(ntiles, p, P) > 0;
    if (warp == 0 && lane == 0) {                   // ---- loader ----
        unsigned it = 0;
        for (int t = 0;; t++) {
            bool any = false;
            for (int p = i0; p < P; p += ta.q) {
                if (t >= tiles_of_producer(ntiles, p, P)) continue;
                any = true;
                wait_ge_at(ta.ready + (size_t)p * kFlagStride + l, (unsigned)t + 1u, ta.trap_site, 111);
                if (ta.ready2) wait_ge_at(ta.ready2 + (size_t)p * kFlagStride + l, (unsigned)t + 1u, ta.trap_site, 122);
                fence_proxy_async_all();
                const unsigned item = (unsigned)t * (unsigned)(L - 1) + (unsigned)(L - l);
                // A_{l-1}; A_1 has two copies (tile parity; the second one in the unused slot of A_L), see the producer
                const uint8_t* srcA = ta.stash_a + ((size_t)p * L + ((l == 2 && (t & 1)) ? L - 1 : l - 2)) * OPB;
                const uint8_t* srcZ = ta.zring + ((size_t)p * ta.zdepth + item % (unsigned)ta.zdepth) * OPB;
                for (int h = 0; h < 2; h++, it++) {
                    const unsigned st = it % 3u;
                    if (it >= 3u) mbar_wait_at(&empty_bar[st], ((it / 3u) - 1u) & 1u, ta.trap_site, 101);
                    umma::mbar_expect_tx(&full_bar[st], 2 * HALF);
                    umma::bulk_g2s(smem + (size_t)st * OPB, srcA + (size_t)h * HALF, HALF, &full_bar[st]);
                    umma::bulk_g2s(smem + (size_t)st * OPB + HALF, srcZ + (size_t)h * HALF, HALF, &full_bar[st]);
                }
            }
            if (!any) break;
        }
    } else if (warp == 1 && lane == 0) {            // ---- MMA issuer ----
        unsigned it = 0;
        for (int t = 0;; t++) {
            bool any = false;
            for (int p = i0; p < P; p += ta.q) {
                if (t >= tiles_of_producer(ntiles, p, P)) continue;
                any = true;
                for (int h = 0; h < 2; h++, it++) {
                    const unsigned st = it % 3u;
                    mbar_wait_at(&full_bar[st], (it / 3u) & 1u, ta.trap_site, 102);
                    umma::tc_fence_after();
                    const uint32_t sa = umma::smem_u32(smem + (size_t)st * OPB), sz = sa + HALF;
                    // both operands K-major over n: 8-column chunks CHS apart, 8 rows (units) 128 B apart
                    #pragma unroll
                    for (int m = 0; m < MT; m++)
                        umma::gemm_issue(tmem + (uint32_t)(m * H), sa + (uint32_t)m * 2048u, CHS, 128, 2 * CHS,
                                         sz, CHS, 128, 2 * CHS, umma::idesc_bf16(128, H, 0, 0), 4, it > 0 ? 1u : 0u);
                    umma::mma_commit(&empty_bar[st]);
                }
                st_release(ta.done + (size_t)p * kFlagStride + l, (unsigned)t + 1u);   // both operand images are on chip
            }
            if (!any) break;
        }
        umma::mma_commit(&fin_bar);
    }
    mbar_wait_at(&fin_bar, 0, ta.trap_site, 103);
    umma::tc_fence_after();
    if (any_work && tid < 2 * HT) {                 // drain: lane = fan-in unit k, columns = units j (tc3 launches one more warp)
        float* dst = part + nd.offW[l] + (size_t)j * H;
        for (int c = half * (H / 2); c < (half + 1) * (H / 2); c += 16) {
            uint32_t r[16];
            tmem_ld16_nowait(tmem + lane_addr + (uint32_t)(mt * H + c), r);
            umma::tmem_ld_wait();
            #pragma unroll
            for (int v = 0; v < 4; v++)
                *reinterpret_cast<float4*>(dst + c + 4 * v) =
                    make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]), __uint_as_float(r[4 * v + 2]),
                                __uint_as_float(r[4 * v + 3]));
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, tmem_cols);
}

// Width 128: the kernel needs 107 KB of shared memory, 256 TMEM columns and — capped at 128 registers — half an SM's register file,
// so TWO CTAs share an SM (launch bounds below) and one's UMMAs run under the other's epilogue: the two-tiles-in-flight overlap that
// width 256 has no shared memory for, for free.
template <int DIN, int N2, int HT>
__global__ void __launch_bounds__(2 * HT, HT == 128 ? 2 : 1)
fused_step_tc2(const NetDims nd, const Tc2Args ta) {
    using namespace tc2;
    constexpr int NC = 1 + DIN + N2;
    static_assert(NC == 6 && DIN == 3 && N2 == 2, "built for the six-channel (x, y, t) networks");
    static_assert(HT == 128 || HT == 256, "width 128 or 256");
    constexpr int H = HT, MT = H / 128;
    constexpr uint32_t OPB = (uint32_t)H * 256u;        // operand image: [16 n-groups][H units][16 B]
    constexpr uint32_t SLAB = 128u * H * 2u;            // one weight slab: 128 rows of the A operand x K = H
    constexpr uint32_t CHS = (uint32_t)H * 16u;         // byte stride between n-chunks of an operand image
    extern __shared__ __align__(128) uint8_t smem2_raw[];
    uint8_t* const smem = smem2_raw;
    __shared__ uint64_t wbar[2], mma_bar, m0_bar, out_bar, slab_free, full_bar[3], empty_bar[3], fin_bar;
    __shared__ uint32_t tmem_slot;

    const int L = nd.L;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q4 = warp & 3, mt = (warp >> 2) % MT, half = (warp >> 2) / MT;
    const int j = mt * 128 + q4 * 32 + lane;            // the unit (TMEM lane q4*32 + lane of M-tile mt) this thread owns
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    // the UMMA issue blocks its thread for most of the step: it lives in a warp of the LAST M-tile, which waits for the end of the step anyway
    const int kIssuer = MT == 2 ? 160 : 0;
    // the issuer's WARP enters issue_step as a whole (warp-uniform branch) and one elected lane issues: with a `tid == k` branch the
    // compiler wraps every tcgen05.mma in an ELECT loop and rebuilds both descriptors — ~16 SASS instructions and ~94 cycles per UMMA
    // from a single thread, against 69 of execution
    const bool issuer_warp = __shfl_sync(0xffffffffu, warp, 0) == kIssuer / 32;

    constexpr uint32_t TMEM_COLS = HT == 128 ? 256u : 512u;      // accumulators: H columns; output layer: 16 more
    constexpr uint32_t OUTC = HT == 128 ? 128u : 256u;
    if (warp == 0) umma::tmem_alloc(&tmem_slot, TMEM_COLS);
    if (tid == 0) {
        umma::mbar_init(&wbar[0], 1); umma::mbar_init(&wbar[1], 1);
        umma::mbar_init(&mma_bar, 2); umma::mbar_init(&m0_bar, 2); umma::mbar_init(&out_bar, 1); umma::mbar_init(&fin_bar, 1); umma::mbar_init(&slab_free, 1);
        for (int i = 0; i < 3; i++) { umma::mbar_init(&full_bar[i], 1); umma::mbar_init(&empty_bar[i], 1); }
        umma::mbar_fence_init();
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    float* part = ta.partial + (size_t)(ta.row0 + blockIdx.x) * nd.PP;
    const int ntiles = ta.set.ntiles, P = ta.nprod;

    if ((int)blockIdx.x >= P) {
        dw_server<HT>(nd, ta, smem, full_bar, empty_bar, &fin_bar, tmem, part, TMEM_COLS);
        return;
    }

    // =========================================================================================
    // producer
    // =========================================================================================
    const int p = (int)blockIdx.x;
    const int my_tiles = tiles_of_producer(ntiles, p, P);
    uint8_t* Wb0 = smem;
    uint8_t* opA = smem + 2 * SLAB;
    uint8_t* wo16 = opA + OPB;                                   // output-layer B operand [2 column groups][H][8] bf16: hi | lo of W_{L+1}
    float* dbs = reinterpret_cast<float*>(wo16 + H * 32);        // [2 halves][L][H]: db_l sums of this CTA
    float* ex = reinterpret_cast<float*>(opA);                   // [2 halves][6][H]: dW_1 / dW_{L+1} exchange after the last tile
    float* ys = dbs + 2 * L * H;                                 // [128 n][4]: outputs y_o of column n
    float* yb = ys + 128 * 4;                                    // [128 n][4]: seeds ybar_o of column n
    float* xs = yb + 128 * 4;                                    // [32 points][4]
    uint8_t* stA = ta.stash_a + (size_t)p * L * OPB;
    uint2* stS = reinterpret_cast<uint2*>(ta.stash_s) + (size_t)p * L * NGRP * H;
    uint8_t* zr = ta.zring + (size_t)p * ta.zdepth * OPB;
    unsigned* rdy = ta.ready + (size_t)p * kFlagStride;
    const unsigned* dn = ta.done + (size_t)p * kFlagStride;
    const uint32_t op_addr = umma::smem_u32(opA), w_addr = umma::smem_u32(Wb0);

    for (int i = tid; i < 2 * L * H; i += blockDim.x) dbs[i] = 0.0f;
    for (int i = tid; i < 128 * 4; i += blockDim.x) { ys[i] = 0.0f; yb[i] = 0.0f; }
    for (int i = tid; i < H * 16; i += blockDim.x) {             // columns 0..2: bf16(W_{L+1}[k][o]); columns 8..10: its bf16 remainder
        const int og = i / (H * 8), k = (i / 8) % H, o = i & 7;
        const float w = o < nd.d_out ? __ldg(ta.params + nd.offW[L + 1] + k * nd.d_out + o) : 0.0f;
        const __nv_bfloat16 hi = __float2bfloat16(w);
        reinterpret_cast<__nv_bfloat16*>(wo16)[i] = og == 0 ? hi : __float2bfloat16(w - __bfloat162float(hi));
    }
    float dW1acc[DIN] = {0.0f, 0.0f, 0.0f}, dWoacc[3] = {0.0f, 0.0f, 0.0f};
    float loss_acc = 0.0f, dbo_acc[3] = {0.0f, 0.0f, 0.0f};

    // ---- weight slabs (thread 0): step sc of the per-tile schedule  fwd l = 2..L, then reverse l = L..2 ----
    const long long nsteps = (long long)my_tiles * 2 * (L - 1);
    auto slab_src = [&](long long s_, int m) -> const __nv_bfloat16* {
        const int idx = (int)((unsigned)s_ % (unsigned)(2 * (L - 1)));   // (32-bit: a 64-bit modulo here was 9 % of the kernel's instructions)
        const bool fwd = idx < L - 1;
        const int l = fwd ? 2 + idx : L - (idx - (L - 1));
        return (fwd ? ta.wf : ta.wb) + ((size_t)(l - 2) * MT + m) * (size_t)(128 * H);
    };
    auto slab_buf = [&](long long s_, int m) { return MT == 2 ? m : (int)(s_ & 1); };
    auto slab_phase = [&](long long s_) { return MT == 2 ? (uint32_t)(s_ & 1) : (uint32_t)((s_ >> 1) & 1); };
    auto request_step = [&](long long s_) {
        if (s_ >= nsteps) return;
        #pragma unroll
        for (int m = 0; m < MT; m++) {
            const int b = slab_buf(s_, m);
            umma::mbar_expect_tx(&wbar[b], SLAB);
            umma::bulk_g2s(Wb0 + (size_t)b * SLAB, slab_src(s_, m), SLAB, &wbar[b]);
        }
    };
    auto wait_step = [&](long long s_) {
        #pragma unroll
        for (int m = 0; m < MT; m++) mbar_wait_at(&wbar[slab_buf(s_, m)], slab_phase(s_), ta.trap_site, 104);
    };
    long long sc = 0;
    uint32_t mma_phase = 0, out_phase = 0;
    if (tid == kIssuer) request_step(0);
    __syncthreads();

    // This thread: unit j, column groups g0..g0+2 (4 points = 24 columns = 3 operand chunks each).  Half 1 ends with the
    // single-point tail group (columns 120..125): its three phantom points are computed and masked, never stored or summed.
    const int g0 = half * 3;
    uint8_t* const op_row = opA + (size_t)j * 16;
    const uint32_t tm_row = tmem + lane_addr + (uint32_t)(mt * 128);
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bool dbg = ta.dbg != nullptr && blockIdx.x == 0 && tid == 0;
    const bool bulk = ta.bulk_store != 0;
    const long long t_begin = clock64();

    auto point_ok = [&](int g, int pp) { return G * g + pp < T; };
    auto chunk_ok = [&](int g, int c) { return 3 * g + c < 16; };
    // 24 FP32 results of group g -> bf16 chunks of this unit's row of an operand image
    // (row: this unit's row of an image in shared memory; grow: the same row of its copy in global memory, or null)
    // early = the second M-tile's UMMA is still reading the operand buffer (this warp started as soon as ITS M-tile was done):
    // the chunks wait in TMEM — in this thread's own accumulator columns, already consumed — and move over afterwards
    const uint32_t tm_hold = tmem + lane_addr + (uint32_t)(mt * 128 + 72 * half);
    auto store_group = [&](uint8_t* row, uint8_t* grow, int g, const float* out, bool early = false) {
        #pragma unroll
        for (int c = 0; c < 3; c++)
            if (chunk_ok(g, c)) {
                const uint32_t w4[4] = {pack2(out[8 * c], out[8 * c + 1]), pack2(out[8 * c + 2], out[8 * c + 3]),
                                        pack2(out[8 * c + 4], out[8 * c + 5]), pack2(out[8 * c + 6], out[8 * c + 7])};
                const uint4 w = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                if (early) tmem_st4(tm_hold + (uint32_t)(12 * (g - g0) + 4 * c), w4);
                else *reinterpret_cast<uint4*>(row + (size_t)(3 * g + c) * CHS) = w;
                if (grow) __stcg(reinterpret_cast<uint4*>(grow + (size_t)(3 * g + c) * CHS), w);
            }
    };
    // parked chunks -> operand buffer (after the UMMA that was reading it has completed)
    auto unpark = [&]() {
        tmem_st_wait();
        #pragma unroll
        for (int gi = 0; gi < 3; gi++) {
            const int g = g0 + gi;
            uint32_t w[12];
            if (g < NGRP - 1) { tmem_ld8_nowait(tm_hold + (uint32_t)(12 * gi), w); tmem_ld4_nowait(tm_hold + (uint32_t)(12 * gi + 8), w + 8); tmem_wait12(w); }
            else { tmem_ld4_nowait(tm_hold + (uint32_t)(12 * gi), w); tmem_wait4(w); }
            #pragma unroll
            for (int c = 0; c < 3; c++)
                if (chunk_ok(g, c))
                    *reinterpret_cast<uint4*>(op_row + (size_t)(3 * g + c) * CHS) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
        }
    };
    auto unpack_chunk = [&](const uint4& w, float* v) {
        v[0] = bf_lo(w.x); v[1] = bf_hi(w.x); v[2] = bf_lo(w.y); v[3] = bf_hi(w.y);
        v[4] = bf_lo(w.z); v[5] = bf_hi(w.z); v[6] = bf_lo(w.w); v[7] = bf_hi(w.w);
    };
    // accumulator columns of group g of this unit's TMEM lane
    auto load_acc = [&](int g, float* v) {
        uint32_t r[24];
        tmem_ld16_nowait(tm_row + (uint32_t)(24 * g), r);
        tmem_ld8_nowait(tm_row + (uint32_t)(24 * g + 16), r + 16);
        tmem_wait24(r);
        #pragma unroll
        for (int i = 0; i < 24; i++) v[i] = __uint_as_float(r[i]);
    };
    // the same in two halves, so that the load of the next group is in flight while this one is worked on
    auto acc_issue = [&](int g, uint32_t* r) {
        tmem_ld16_nowait(tm_row + (uint32_t)(24 * g), r);
        tmem_ld8_nowait(tm_row + (uint32_t)(24 * g + 16), r + 16);
    };
    // forward tanh layer on one point: z (value, 3 first, 2 second tangents) -> a, a', a'' ; returns s
    auto fwd_point = [&](const float* z, float bias, float* out) -> float {
        float a, s;
        tanh_s(z[0] + bias, a, s);
        const float m2as = -2.0f * a * s;
        out[0] = a;
        #pragma unroll
        for (int i = 0; i < DIN; i++) out[1 + i] = s * z[1 + i];
        #pragma unroll
        for (int i = 0; i < N2; i++) out[1 + DIN + i] = fmaf(m2as, z[1 + i] * z[1 + i], s * z[1 + DIN + i]);
        return s;
    };
    // reverse through the tanh layer on one point, in the layer outputs av = (a, a', a''): ab -> zb ; returns zbar_0
    auto bwd_point = [&](const float* av, float s, const float* ab, float* zb) -> float {
        const float a = av[0], a2 = 2.0f * a;
        float z0 = s * ab[0];
        #pragma unroll
        for (int i = 0; i < DIN; i++) {
            float zp = s * ab[1 + i];
            z0 = fmaf(-a2 * av[1 + i], ab[1 + i], z0);
            if (i < N2) {
                const float abpp = ab[1 + DIN + i];
                zb[1 + DIN + i] = s * abpp;
                zp = fmaf(-2.0f * a2 * av[1 + i], abpp, zp);
                z0 = fmaf(-(a2 * av[1 + DIN + i] + 2.0f * av[1 + i] * av[1 + i]), abpp, z0);
            }
            zb[1 + i] = zp;
        }
        zb[0] = z0;
        return z0;
    };
    // forward epilogue of one group: pre-activations z (24) -> operand chunks (+ bf16 remainder image for layer L) + s stash
    auto fwd_group = [&](int l, int g, const float* z, float bias, uint8_t* lo_row, uint8_t* grow, bool early = false) {
        float out[24], sv[4];
        #pragma unroll
        for (int pp = 0; pp < G; pp++) sv[pp] = fwd_point(z + NC * pp, bias, out + NC * pp);
        if (g == NGRP - 1) { out[6] = 0.0f; out[7] = 0.0f; }                 // pad columns 126, 127 of the tile
        store_group(op_row, grow, g, out, early);
        if (lo_row) {
            float lo[24];
            #pragma unroll
            for (int i = 0; i < 24; i++) lo[i] = out[i] - __bfloat162float(__float2bfloat16(out[i]));
            store_group(lo_row, nullptr, g, lo);
        }
        stS[(size_t)((l - 1) * NGRP + g) * H + j] = make_uint2(pack2(sv[0], sv[1]), pack2(sv[2], sv[3]));
    };

    // Hand-off rules (thread 0 checks the server flags BEFORE it lets the CTA into the epilogue that overwrites the data):
    //   A_l   (stash image l, l < L; dW_{l+1} operand): rewritten by the forward epilogue of layer l of the NEXT tile ->
    //         done[l+1] >= t before that epilogue;
    //   Zbar_l (ring slot item % zdepth): written by the reverse epilogue of layer l -> the item zdepth earlier is done.
    // Every thread writes its chunks to shared memory (next UMMA operand) AND to global memory (stash / ring); thread 0
    // publishes ready[l] right after the block barrier that follows the epilogue (release at gpu scope is cumulative).
    auto ring_wait = [&](int t_, int l) {
        const unsigned item = (unsigned)t_ * (unsigned)(L - 1) + (unsigned)(L - l);
        if (item >= (unsigned)ta.zdepth) {
            const unsigned old = item - (unsigned)ta.zdepth;
            wait_ge_at(dn + (L - (int)(old % (unsigned)(L - 1))), old / (unsigned)(L - 1) + 1u, ta.trap_site, 112);
        }
    };
    auto ring_row = [&](int t_, int l) -> uint8_t* {
        const unsigned item = (unsigned)t_ * (unsigned)(L - 1) + (unsigned)(L - l);
        return zr + (size_t)(item % (unsigned)ta.zdepth) * OPB + (size_t)j * 16;
    };
    // forward / reverse UMMAs of step s_ (thread 0): M-tile m as soon as its weight slab has landed; after the last UMMA of
    // M-tile 0 a commit on slab_free lets the loader thread refill that buffer while M-tile 1 still runs.
    auto issue_step = [&](long long s_, bool fwd) {               // (all lanes of the issuer's warp)
        #pragma unroll
        for (int m = 0; m < MT; m++) {
            mbar_wait_at(&wbar[slab_buf(s_, m)], slab_phase(s_), ta.trap_site, 105);
            umma::tc_fence_after();
            if (ta.stagger < 0 ? lane == 0 : umma::elect_one()) {
                if (ta.stagger < 0) {      // (A/B switch for measurements, PINN_TC2_LEAN=0: the descriptor-per-UMMA issue of round 1)
                    if (fwd) umma::gemm_issue(tmem + (uint32_t)(m * 128), w_addr + (uint32_t)slab_buf(s_, m) * SLAB, 128, CHS, 256,
                                              op_addr, 128, CHS, 256, umma::idesc_bf16(128, 128, 1, 1), H / 16, 0u);
                    else umma::gemm_issue(tmem + (uint32_t)(m * 128), w_addr + (uint32_t)slab_buf(s_, m) * SLAB, 2048, 128, 4096,
                                          op_addr, 128, CHS, 256, umma::idesc_bf16(128, 128, 0, 1), H / 16, 0u);
                } else if (fwd)   // A = W_l^T slab (MN-major: 8-unit groups CHS apart, 8 k rows 128 B apart)
                    umma::gemm_issue_lean<H / 16>(tmem + (uint32_t)(m * 128), w_addr + (uint32_t)slab_buf(s_, m) * SLAB, 128, CHS, 256,
                                                  op_addr, 128, CHS, 256, umma::idesc_bf16(128, 128, 1, 1), 0u);
                else       // A = W_l rows (K-major: 8-unit chunks of j 2048 B apart, 8 k rows 128 B apart)
                    umma::gemm_issue_lean<H / 16>(tmem + (uint32_t)(m * 128), w_addr + (uint32_t)slab_buf(s_, m) * SLAB, 2048, 128, 4096,
                                                  op_addr, 128, CHS, 256, umma::idesc_bf16(128, 128, 0, 1), 0u);
                if (MT == 2 && m == 0) { umma::mma_commit(&slab_free); umma::mma_commit(&m0_bar); }
                if (m == MT - 1) {
                    umma::mma_commit(&mma_bar);
                    if (MT == 1) request_step(s_ + 1);
                }
            }
            __syncwarp();
        }
    };
    uint32_t free_phase = 0;                         // loader thread (tid 32), MT == 2
    // loader thread: next step's slab for buffer 0 as soon as M-tile 0 is done, for buffer 1 when the whole step is
    auto loader_first = [&](long long s_next) {
        mbar_wait_at(&slab_free, free_phase, ta.trap_site, 106); free_phase ^= 1;
        if (s_next < nsteps) {
            umma::mbar_expect_tx(&wbar[0], SLAB);
            umma::bulk_g2s(Wb0, slab_src(s_next, 0), SLAB, &wbar[0]);
        }
    };
    auto loader_second = [&](long long s_next) {
        if (s_next < nsteps) {
            umma::mbar_expect_tx(&wbar[1], SLAB);
            umma::bulk_g2s(Wb0 + SLAB, slab_src(s_next, 1), SLAB, &wbar[1]);
        }
    };

    for (int t = 0; t < my_tiles; t++) {
        const long long c_t0 = dbg ? clock64() : 0;
        const unsigned long long base = (unsigned long long)(p + (long long)t * P) * T;
        // A_1 is the first image a tile writes and the last one its dW server reads (dW_2 closes the reverse pass): with one copy
        // the next tile would wait for that server at once, so A_1 alternates between two copies (the slot of A_L is free: A_L is
        // never stashed) and a tile only waits for the tile before the previous one
        if (tid == kWaiter && t > 1) wait_ge_at(dn + 2, (unsigned)t - 1u, ta.trap_site, 113);
        const size_t a1_off = (t & 1) ? (size_t)(L - 1) * OPB : 0;
        if (tid < T * DIN) {
            const int pt = tid / DIN, k = tid - pt * DIN;
            xs[pt * 4 + k] = (base + pt < ta.set.n) ? __ldg(ta.set.x + (base + pt) * DIN + k) : 0.0f;
        }
        __syncthreads();

        // ---------------- layer 1 (K = d_in): CUDA cores ----------------
        {
            float w1[DIN];
            #pragma unroll
            for (int k = 0; k < DIN; k++) w1[k] = __ldg(ta.params + nd.offW[1] + k * H + j);
            const float b1 = __ldg(ta.params + nd.offB[1] + j);
            #pragma unroll
            for (int gi = 0; gi < 3; gi++) {
                const int g = g0 + gi;
                float z[24];
                #pragma unroll
                for (int pp = 0; pp < G; pp++) {
                    const float4 x = *reinterpret_cast<const float4*>(xs + ((G * g + pp) & 31) * 4);
                    z[NC * pp] = fmaf(x.x, w1[0], fmaf(x.y, w1[1], x.z * w1[2]));
                    #pragma unroll
                    for (int i = 0; i < DIN; i++) z[NC * pp + 1 + i] = w1[i];
                    #pragma unroll
                    for (int i = 0; i < N2; i++) z[NC * pp + 1 + DIN + i] = 0.0f;
                }
                fwd_group(1, g, z, b1, nullptr, bulk ? nullptr : stA + a1_off + (size_t)j * 16);
            }
        }
        umma::fence_async_smem();
        umma::tc_fence_before();
        __syncthreads();
        if (bulk && tid == kWaiter) { umma::bulk_s2g(stA + a1_off, opA, OPB); umma::bulk_commit(); }
        if (dbg) tacc[6] += clock64() - c_t0;

        // ---------------- hidden layers 2..L: tensor cores ----------------
        uint8_t* lo_img = Wb0;                       // bf16 remainder of A_L: in a weight buffer that is idle around the output layer
        for (int l = 2; l <= L; l++, sc++) {
            const long long c_a = dbg ? clock64() : 0;
            if (issuer_warp) issue_step(sc, true);
            if (tid == kWaiter) {
                if (bulk) umma::bulk_wait_read_all();    // the bulk store of the previous image has read the operand buffer
                if (l < L && t > 0) wait_ge_at(dn + (l + 1), (unsigned)t, ta.trap_site, 114);   // A_l of the previous tile may be overwritten
                if (MT == 2) mbar_arrive(&m0_bar);
                mbar_arrive(&mma_bar);
            }
            if (MT == 2 && tid == kLoader && l < L) loader_first(sc + 1);
            if (MT == 2 && tid == kLoader && l == L) { mbar_wait_at(&slab_free, free_phase, ta.trap_site, 107); free_phase ^= 1; }
            const float bl = __ldg(ta.params + nd.offB[l] + j);
            // the warps of M-tile 0 start their epilogue as soon as ITS UMMAs are done, under the UMMAs of M-tile 1
            const bool early = MT == 2 && mt == 0;
            if (early) { mbar_wait_at(&m0_bar, mma_phase, ta.trap_site, 115); }
            else {
                mbar_wait_at(&mma_bar, mma_phase, ta.trap_site, 108);
                if (MT == 2 && tid == kLoader && l < L) loader_second(sc + 1);
            }
            umma::tc_fence_after();
            if (l == L) lo_img = Wb0 + (MT == 2 ? 0 : (size_t)(sc & 1) * SLAB);
            const long long c_b = dbg ? clock64() : 0;
            {
                uint32_t rz[2][24];
                acc_issue(g0, rz[0]);
                #pragma unroll
                for (int gi = 0; gi < 3; gi++) {
                    tmem_wait24(rz[gi & 1]);
                    if (gi < 2) acc_issue(g0 + gi + 1, rz[(gi + 1) & 1]);
                    float z[24];
                    #pragma unroll
                    for (int i = 0; i < 24; i++) z[i] = __uint_as_float(rz[gi & 1][i]);
                    fwd_group(l, g0 + gi, z, bl, l == L ? lo_img + (size_t)j * 16 : nullptr,
                              (l < L && !bulk) ? stA + (size_t)(l - 1) * OPB + (size_t)j * 16 : nullptr, early);
                }
            }
            if (early) {
                mbar_wait_at(&mma_bar, mma_phase, ta.trap_site, 116);
                umma::tc_fence_after();
                unpark();
            }
            mma_phase ^= 1;
            umma::fence_async_smem();
            umma::tc_fence_before();
            __syncthreads();
            if (bulk && l < L && tid == kWaiter) { umma::bulk_s2g(stA + (size_t)(l - 1) * OPB, opA, OPB); umma::bulk_commit(); }
            if (dbg) { const long long c_c = clock64(); tacc[0] += c_b - c_a; tacc[1] += c_c - c_b; }
        }

        // ---------------- output layer on the tensor cores: y[n, :] = (A_L hi + lo)^T[n, k] . [W_{L+1} hi | lo][k, :] ----------------
        const long long c_o = dbg ? clock64() : 0;
        if (issuer_warp) {
            umma::tc_fence_after();
            if (umma::elect_one()) {
                umma::gemm_issue_lean<H / 16>(tmem + OUTC, op_addr, 128, CHS, 256, umma::smem_u32(wo16), 128, CHS, 256,
                                              umma::idesc_bf16(128, 16, 1, 1), 0u);
                umma::gemm_issue_lean<H / 16>(tmem + OUTC, umma::smem_u32(lo_img), 128, CHS, 256, umma::smem_u32(wo16), 128, CHS, 256,
                                              umma::idesc_bf16(128, 16, 1, 1), 1u);
                umma::mma_commit(&out_bar);
            }
            __syncwarp();
        }
        mbar_wait_at(&out_bar, out_phase, ta.trap_site, 109); out_phase ^= 1;
        umma::tc_fence_after();
        if (MT == 2 && tid == kLoader) { loader_second(sc); if (sc < nsteps) { umma::mbar_expect_tx(&wbar[0], SLAB); umma::bulk_g2s(Wb0, slab_src(sc, 0), SLAB, &wbar[0]); } }
        if (warp < 4) {                              // lane n = column n of the tile
            uint32_t r[16];
            tmem_ld16_nowait(tmem + lane_addr + OUTC, r);
            tmem_wait16(r);
            *reinterpret_cast<float4*>(ys + tid * 4) =
                make_float4(__uint_as_float(r[0]) + __uint_as_float(r[8]), __uint_as_float(r[1]) + __uint_as_float(r[9]),
                            __uint_as_float(r[2]) + __uint_as_float(r[10]), 0.0f);
        }
        umma::tc_fence_before();
        __syncthreads();
        if (tid < T) {                               // residual + seeds, one thread per point
            float y[3][NC], ybar[3][NC], r[3], x[3];
            #pragma unroll
            for (int o = 0; o < 3; o++)
                #pragma unroll
                for (int c = 0; c < NC; c++)
                    y[o][c] = (o < nd.d_out) ? ys[(NC * tid + c) * 4 + o] + (c == 0 ? __ldg(ta.params + nd.offB[L + 1] + o) : 0.0f) : 0.0f;
            #pragma unroll
            for (int k = 0; k < DIN; k++) x[k] = xs[tid * 4 + k];
            float ss = pde_residual<NC>(nd, x, y, ta.set.scale, ybar, r);
            const bool valid = base + tid < ta.set.n;
            if (!valid) ss = 0.0f;
            #pragma unroll
            for (int c = 0; c < NC; c++)
                *reinterpret_cast<float4*>(yb + (NC * tid + c) * 4) =
                    valid ? make_float4(ybar[0][c], ybar[1][c], ybar[2][c], 0.0f) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            loss_acc += ss;
            if (valid) {
                #pragma unroll
                for (int o = 0; o < 3; o++) dbo_acc[o] += ybar[o][0];
            }
        }
        if (tid == kWaiter) { if (bulk) umma::bulk_wait_read_all(); ring_wait(t, L); }   // ring slot of Zbar_L
        __syncthreads();
        if (dbg) tacc[2] += clock64() - c_o;

        // ---------------- reverse pass ----------------
        float wo[3];
        #pragma unroll
        for (int o = 0; o < 3; o++) wo[o] = (o < nd.d_out) ? __ldg(ta.params + nd.offW[L + 1] + j * nd.d_out + o) : 0.0f;
        // Stash of reverse step l (s, and the layer outputs A_l): fetched one step ahead — it does not depend on the UMMA in
        // flight, so its L2 latency hides behind the block barrier and the UMMA.
        uint4 pa[3][3]; uint2 ps[3];
        auto prefetch_stash = [&](int l) {
            #pragma unroll
            for (int gi = 0; gi < 3; gi++) {
                const int g = g0 + gi;
                ps[gi] = __ldcg(stS + (size_t)((l - 1) * NGRP + g) * H + j);
                #pragma unroll
                for (int c = 0; c < 3; c++) {
                    pa[gi][c] = make_uint4(0u, 0u, 0u, 0u);
                    if (chunk_ok(g, c)) {
                        if (l < L) pa[gi][c] = __ldcg(reinterpret_cast<const uint4*>(stA + (l == 1 ? a1_off : (size_t)(l - 1) * OPB) + (size_t)(3 * g + c) * CHS + (size_t)j * 16));
                        else pa[gi][c] = *reinterpret_cast<const uint4*>(op_row + (size_t)(3 * g + c) * CHS);   // A_L is still the operand image
                    }
                }
            }
        };
        prefetch_stash(L);
        for (int l = L; l >= 1; l--) {
            const long long c_d = dbg ? clock64() : 0;
            const bool early = MT == 2 && mt == 0 && l < L;      // (see the forward pass)
            if (l < L) {
                if (MT == 2 && tid == kLoader) loader_first(sc);
                if (early) { mbar_wait_at(&m0_bar, mma_phase, ta.trap_site, 117); }
                else {
                    mbar_wait_at(&mma_bar, mma_phase, ta.trap_site, 110);
                    if (MT == 2 && tid == kLoader) loader_second(sc);
                }
                umma::tc_fence_after();
            }
            const long long c_e = dbg ? clock64() : 0;
            float dbsum = 0.0f, dw1[DIN] = {0.0f, 0.0f, 0.0f};
            uint8_t* const zrow = (l > 1 && !bulk) ? ring_row(t, l) : nullptr;
            #pragma unroll
            for (int gi = 0; gi < 3; gi++) {
                const int g = g0 + gi;
                float av[24], ab[24], zb[24];
                #pragma unroll
                for (int c = 0; c < 3; c++) unpack_chunk(pa[gi][c], av + 8 * c);
                if (l == L) {
                    // Abar_L = ybar W_{L+1}^T and dW_{L+1} on the CUDA cores
                    #pragma unroll
                    for (int i = 0; i < 24; i++) {
                        const float4 s4 = *reinterpret_cast<const float4*>(yb + ((24 * g + i) & 127) * 4);
                        ab[i] = fmaf(wo[0], s4.x, fmaf(wo[1], s4.y, wo[2] * s4.z));
                        const float a_ok = point_ok(g, i / NC) ? av[i] : 0.0f;
                        dWoacc[0] = fmaf(a_ok, s4.x, dWoacc[0]);
                        dWoacc[1] = fmaf(a_ok, s4.y, dWoacc[1]);
                        dWoacc[2] = fmaf(a_ok, s4.z, dWoacc[2]);
                    }
                } else {
                    load_acc(g, ab);
                }
                const float sv[4] = {bf_lo(ps[gi].x), bf_hi(ps[gi].x), bf_lo(ps[gi].y), bf_hi(ps[gi].y)};
                #pragma unroll
                for (int pp = 0; pp < G; pp++) {
                    const float z0 = bwd_point(av + NC * pp, sv[pp], ab + NC * pp, zb + NC * pp);
                    if (point_ok(g, pp)) {
                        dbsum += z0;
                        if (l == 1) {
                            const float4 x = *reinterpret_cast<const float4*>(xs + ((G * g + pp) & 31) * 4);
                            dw1[0] += fmaf(x.x, z0, zb[NC * pp + 1]);
                            dw1[1] += fmaf(x.y, z0, zb[NC * pp + 2]);
                            dw1[2] += fmaf(x.z, z0, zb[NC * pp + 3]);
                        }
                    }
                }
                if (g == NGRP - 1) { zb[6] = 0.0f; zb[7] = 0.0f; }
                if (l > 1) store_group(op_row, zrow, g, zb, early);
            }
            if (l < L) {
                if (early) {
                    mbar_wait_at(&mma_bar, mma_phase, ta.trap_site, 118);
                    umma::tc_fence_after();
                    if (l > 1) unpark();
                }
                mma_phase ^= 1;
            }
            dbs[(size_t)(half * L + (l - 1)) * H + j] += dbsum;
            if (l == 1) {
                #pragma unroll
                for (int k = 0; k < DIN; k++) dW1acc[k] += dw1[k];
                if (dbg) { const long long c_f = clock64(); tacc[3] += c_e - c_d; tacc[4] += c_f - c_e; }
                break;
            }
            prefetch_stash(l - 1);
            umma::fence_async_smem();
            umma::tc_fence_before();
            __syncthreads();
            const long long c_f = dbg ? clock64() : 0;
            if (issuer_warp) issue_step(sc, false);
            if (tid == kPublisher && !bulk) {            // Zbar_l (and, since the forward pass, A_{l-1}) are in memory
                __threadfence();
                st_release(rdy + l, (unsigned)t + 1u);
            }
            if (tid == kWaiter) {
                if (bulk) {
                    // the image leaves by one bulk store.  Its completion is not waited for here (this thread's arrival starts every
                    // epilogue warp): the flag of layer l + 1 is released now that only the newest store can still be in flight, and
                    // the last two of the tile together
                    umma::bulk_s2g(ring_row(t, l) - (size_t)j * 16, opA, OPB);
                    umma::bulk_commit();
                    if (l == 2) umma::bulk_wait_all(); else umma::bulk_wait_all_but_last();
                    __threadfence();
                    if (l < L) st_release(rdy + (l + 1), (unsigned)t + 1u);
                    if (l == 2) st_release(rdy + 2, (unsigned)t + 1u);
                }
                if (l > 2) ring_wait(t, l - 1);          // ring slot of Zbar_{l-1}, written by the next epilogue
                if (MT == 2) mbar_arrive(&m0_bar);
                mbar_arrive(&mma_bar);
            }
            sc++;
            if (dbg) { const long long c_g = clock64(); tacc[3] += c_e - c_d; tacc[4] += c_f - c_e; tacc[5] += c_g - c_f; }
        }
        __syncthreads();                             // xs / yb are rewritten by the next tile
    }
    if (dbg) {
        tacc[7] = clock64() - t_begin;
        for (int i = 0; i < 8; i++) ta.dbg[i] = tacc[i];
    }

    // ---------------- flush the per-CTA sums into this CTA's partial row (single owner per element) ----------------
    __syncthreads();
    for (int i = tid; i < L * H; i += blockDim.x) {
        const int l = i / H + 1, jj = i - (l - 1) * H;
        part[nd.offB[l] + jj] = dbs[(size_t)(l - 1) * H + jj] + dbs[(size_t)(L + (l - 1)) * H + jj];
    }
    // dW_1 [k][j] and dW_{L+1} [j][o]: the two column halves meet in shared memory, half 0 then half 1
    #pragma unroll
    for (int k = 0; k < DIN; k++) ex[(size_t)(half * 6 + k) * H + j] = dW1acc[k];
    #pragma unroll
    for (int o = 0; o < 3; o++) ex[(size_t)(half * 6 + 3 + o) * H + j] = dWoacc[o];
    __syncthreads();
    if (half == 0) {
        #pragma unroll
        for (int k = 0; k < DIN; k++) part[nd.offW[1] + k * H + j] = ex[(size_t)k * H + j] + ex[(size_t)(6 + k) * H + j];
        for (int o = 0; o < nd.d_out; o++) part[nd.offW[L + 1] + j * nd.d_out + o] = ex[(size_t)(3 + o) * H + j] + ex[(size_t)(9 + o) * H + j];
    }
    if (warp == 0) {
        float v[4] = {loss_acc, dbo_acc[0], dbo_acc[1], dbo_acc[2]};
        #pragma unroll
        for (int i = 0; i < 4; i++)
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
        if (lane == 0) {
            part[nd.P + ta.set.set] = v[0];
            for (int o = 0; o < nd.d_out; o++) part[nd.offB[L + 1] + o] = v[1 + o];
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, TMEM_COLS);
}

// BF16 mirrors of the hidden->hidden weights in the two slab layouts of Tc2Args (wf, wb).
// khalf = 1 (tc3): the forward mirror keeps each 128-wide half of K contiguous, [j/128][k/128][(j%128)/8][k%128][j%8], so that
// a 32 KB piece (M-tile x K-half) is one bulk copy; the reverse mirror already is (its K runs over the j-chunks).
__global__ void pack_w2_kernel(const NetDims nd, const float* __restrict__ params, __nv_bfloat16* __restrict__ wf,
                               __nv_bfloat16* __restrict__ wb, int khalf) {
    const int H = nd.H, MT = H / 128;
    const long long n = (long long)(nd.L - 1) * H * H;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int l = (int)(i / ((long long)H * H)) + 2;
    const int r = (int)(i % ((long long)H * H)), k = r / H, j = r - k * H;
    const __nv_bfloat16 w = __float2bfloat16(params[nd.offW[l] + r]);
    const size_t lbase = (size_t)(l - 2) * H * H;
    // forward: [j/128][ (j%128)/8 ][k][j%8]
    if (khalf) wf[lbase + (size_t)(j >> 7) * (128 * H) + (size_t)(k >> 7) * (128 * 128) + (size_t)((j & 127) >> 3) * (128 * 8) + (size_t)(k & 127) * 8 + (j & 7)] = w;
    else wf[lbase + (size_t)(j >> 7) * (128 * H) + (size_t)((j & 127) >> 3) * (H * 8) + (size_t)k * 8 + (j & 7)] = w;
    // reverse: [k/128][ j/8 ][k%128][j%8]
    wb[lbase + (size_t)(k >> 7) * (128 * H) + (size_t)(j >> 3) * (128 * 8) + (size_t)(k & 127) * 8 + (j & 7)] = w;
    (void)MT;
}

}  // namespace pinn
