// pinn_fused_tc3.cuh — producer CTA of the wide-network tensor-core train step, third form: TWO SUB-TILES IN FLIGHT,
// warp-specialised.
//
// The tc2 producer (pinn_fused_tc2.cuh) walks one 21-point tile through the layers: UMMA -> epilogue -> UMMA ..., so
// the tensor pipe idles during every tanh / Taylor epilogue and the CUDA cores idle during every UMMA (ncu: tensor pipe
// 28 % active, 42 % of the stall samples on the UMMA-done barrier).  Shared memory has no room for a second tile next to
// the 128 KB of layer weights at width 256, so one 128-column operand image is run as two independent 64-column
// SUB-TILES A and B (10 points x 6 Taylor channels = 60 columns + 4 zero columns each) that share the resident weights:
//
//   epilogue warps (all 16):   E(A, step s)   E(B, s)        E(A, s+1)      E(B, s+1) ...
//   tensor pipe:                              [A blocks s+1] [B blocks s+1] [A blocks s+2] ...
//
// * EPILOGUE WARPS (2 H threads): thread = (unit j, half h of the sub-tile's points): TMEM lane = unit, so the tanh /
//   Taylor forward, the reverse adjoint, bias gradient, dW_1 and dW_{L+1} are thread-local straight out of tcgen05.ld.
//   ALL of them work on ONE sub-tile at a time (4 + 1 points per thread) and alternate A, B, A, B: each sub-tile's
//   epilogue gets the whole SM (4 warps per scheduler instead of 2), and while it runs the other sub-tile's UMMAs do.
// * FOUR ISSUER WARPS (sub-tile x M-tile): wait for the sub-tile's "ready" mbarrier (one arrival per epilogue warp), issue
//   their UMMA block with a lean path (descriptors stepped by an add, elect.sync), commit to the sub-tile's "done"
//   mbarrier.  A single thread issues one UMMA per ~69 cycles whatever N is; two threads interleave to ~37 cycles per
//   N = 64 UMMA, which is what makes 64-column sub-tiles as cheap as one 128-column tile (measured).
// * LOADER WARP: weights travel as 32 KB pieces (M-tile x K-half) through a ring; a piece is refilled as soon as both
//   sub-tiles' UMMAs on it are done (mbarrier of count two).
// * HELPER WARP: everything that talks to the dW servers — stash / ring back-pressure (done flags) before a sub-tile's
//   UMMAs may complete, and the release of ready flags (one per sub-tile; the server waits for both halves of an image).
// Registers (width 256): 768 threads, 80 at launch; setmaxnreg gives the epilogue warps 104, the issuer warps 40, the loader /
// helper warps 24.
// STATUS: parity-green and bit-reproducible (tests/test_gpu_parity.py, PINN_TC3=1) but NOT the default: 29.6 Mpts/s against 32.4
// for tc2 on C5 — see DESIGN.md section 5 for what the five variants of this file measured.
// The output layer runs on the tensor cores from an FP16 image of A_L (11-bit significands, saturating conversion; the
// hidden layers keep BF16) against W_{L+1} split hi | lo in FP16.
// The math, the operand / stash / ring images and the dW-server CTAs are those of tc2 (a tile = the pair A|B = one
// 128-column image).
//
// Reference: the reference has no network code (include/network.cuh is empty); the building blocks this replaces are
// tensor::matmul (include/tensor.cuh:380-409) + kernel::matmul (include/device/kernel.cuh:68-83).
#pragma once
#include <cuda_fp16.h>
#include "pinn_fused_tc2.cuh"

namespace pinn {

namespace tc3 {
constexpr int NP = 10, T = 2 * NP;              // points per sub-tile, per tile (pair of sub-tiles)
constexpr int kHelperThreads = 256;             // four issuer warps (sub-tile x M-tile), loader, helper (+ two idle: register allocation is per four warps)
// Instruction descriptor for kind::f16 with FP16 operands and FP32 accumulation (the output layer).
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t pack2h(float lo, float hi) {          // saturating: a huge tangent becomes +-65504, not inf
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
}  // namespace tc3

template <int DIN, int N2, int HT>
__global__ void __launch_bounds__(2 * HT + tc3::kHelperThreads, 1)
fused_step_tc3(const NetDims nd, const Tc2Args ta) {
    using namespace tc2;
    constexpr int NC = 1 + DIN + N2;
    static_assert(NC == 6 && DIN == 3 && N2 == 2, "built for the six-channel (x, y, t) networks");
    static_assert(HT == 128 || HT == 256, "width 128 or 256");
    constexpr int H = HT, MT = H / 128, NP = tc3::NP, T = tc3::T;
    constexpr int NEPI = 2 * HT;                        // epilogue threads
    constexpr uint32_t OPB = (uint32_t)H * 256u;        // image of the pair: [16 n-chunks][H units][16 B]
    constexpr uint32_t SLAB = 128u * H * 2u;
    constexpr uint32_t CHS = (uint32_t)H * 16u;
    constexpr uint32_t SUB = 8u * CHS;                  // one sub-tile's half of an image
    constexpr int NKH = H / 128, PPS = MT * NKH;        // K-halves per M-tile block; weight pieces per layer step (4 at width 256, 1 at 128)
    constexpr uint32_t PIECE = 128u * 128u * 2u;
    constexpr int NBUF = (int)(2 * SLAB / PIECE);       // ring: 4 buffers at width 256, 2 at width 128
    constexpr int AHEAD = NBUF / PPS;                   // layer steps a refill runs ahead (1 / 2)
    extern __shared__ __align__(128) uint8_t smem3_raw[];
    uint8_t* const smem = smem3_raw;
    __shared__ uint64_t wbar[4], free_bar[4], ud[2], rdy[2], out_bar[2], full_bar[3], empty_bar[3], fin_bar;
    __shared__ uint32_t tmem_slot;

    const int L = nd.L;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q4 = warp & 3, mt = (warp >> 2) % MT, h = ((warp >> 2) / MT) & 1;
    const int j = mt * 128 + q4 * 32 + lane;            // (epilogue threads) the unit (TMEM lane q4*32 + lane of M-tile mt) this thread owns
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;

    if (warp == 0) umma::tmem_alloc(&tmem_slot, 512u);
    if (tid == 0) {
        for (int i = 0; i < 4; i++) { umma::mbar_init(&wbar[i], 1); umma::mbar_init(&free_bar[i], 2); }
        umma::mbar_init(&ud[0], MT + 1); umma::mbar_init(&ud[1], MT + 1);           // one commit per M-tile issuer + the helper's arrival
        umma::mbar_init(&rdy[0], NEPI / 32); umma::mbar_init(&rdy[1], NEPI / 32);   // one arrival per epilogue warp
        umma::mbar_init(&out_bar[0], 1); umma::mbar_init(&out_bar[1], 1);
        umma::mbar_init(&fin_bar, 1);
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
        dw_server<HT>(nd, ta, smem, full_bar, empty_bar, &fin_bar, tmem, part);
        return;
    }

    // =========================================================================================
    // producer
    // =========================================================================================
    const int p = (int)blockIdx.x;
    const int my_tiles = tiles_of_producer(ntiles, p, P);
    uint8_t* Wb0 = smem;
    uint8_t* img = smem + 2 * SLAB;                               // [A | B] x [8 chunks][H][16 B]
    uint8_t* wo16 = img + OPB;                                    // output-layer B operand [2 column groups][H][8] fp16: hi | lo of W_{L+1}
    float* dbs = reinterpret_cast<float*>(wo16 + H * 32);         // [2 halves][L][H]: db_l sums of this CTA
    float* ex = reinterpret_cast<float*>(img);                    // [2][6][H]: dW_1 / dW_{L+1} exchange after the last tile
    float* ys_base = dbs + 2 * L * H;                             // [2 sub-tiles][64 n][4]: outputs y_o of column n
    float* yb_base = ys_base + 128 * 4;                           // [2][64 n][4]: seeds ybar_o
    float* xs_base = yb_base + 128 * 4;                           // [2][16 points][4]
    uint8_t* stA = ta.stash_a + (size_t)p * L * OPB;
    uint2* stS = reinterpret_cast<uint2*>(ta.stash_s) + (size_t)p * L * 6 * H;
    uint8_t* zr = ta.zring + (size_t)p * ta.zdepth * OPB;
    const unsigned* dn = ta.done + (size_t)p * kFlagStride;
    const uint32_t img_addr = umma::smem_u32(img), w_addr = umma::smem_u32(Wb0);
    const int spt = 2 * (L - 1);                                  // layer steps per tile
    const int nsteps = my_tiles * spt;
    auto piece_slot = [&](int s_, int i) { return PPS == 4 ? i : ((s_ * PPS + i) % NBUF); };
    auto piece_phase = [&](int s_) { return (uint32_t)((s_ * PPS / NBUF) & 1); };
    auto ring_item = [&](int t_, int l) { return (unsigned)t_ * (unsigned)(L - 1) + (unsigned)(L - l); };
    auto ring_wait = [&](int t_, int l) {
        const unsigned item = ring_item(t_, l);
        if (item >= (unsigned)ta.zdepth) {
            const unsigned old = item - (unsigned)ta.zdepth;
            wait_ge_at(dn + (L - (int)(old % (unsigned)(L - 1))), old / (unsigned)(L - 1) + 1u, ta.trap_site, 112);
        }
    };
    const bool dbg = ta.dbg != nullptr && blockIdx.x == 0;
    constexpr int kTraceTile = 5;                                 // (debug) the tile whose timeline is recorded

    for (int i = tid; i < 2 * L * H; i += blockDim.x) dbs[i] = 0.0f;
    for (int i = tid; i < H * 16; i += blockDim.x) {              // columns 0..2: fp16(W_{L+1}[k][o]); columns 8..10: its fp16 remainder
        const int og = i / (H * 8), k = (i / 8) % H, o = i & 7;
        const float w = o < nd.d_out ? __ldg(ta.params + nd.offW[L + 1] + k * nd.d_out + o) : 0.0f;
        const __half hi = __float2half_rn(w);
        reinterpret_cast<__half*>(wo16)[i] = og == 0 ? hi : __float2half_rn(w - __half2float(hi));
    }
    __syncthreads();

    if (warp >= NEPI / 32) {
        // =====================================================================================
        // the four single-thread warps
        // =====================================================================================
        // register budget at width 256 (768 threads, 80 at launch): epilogue warps 104, issuer warps 40, loader / helper warps 24
        const int role = warp - NEPI / 32;
        if (MT == 2) { if (role < 4) tc3::setmaxnreg_dec<40>(); else tc3::setmaxnreg_dec<24>(); }
        if (role < 4) {
            // ---------------- issuer of (sub-tile X, M-tile m): whole warp walks the schedule, one elected lane issues ----------------
            // (one thread issues a UMMA per ~69 cycles at best, whatever N: the two M-tile blocks of a step go out from two threads)
            const int X = role >> 1, m = role & 1;
            if (m < MT) {
                uint32_t rdy_par = 0;
                for (int t = 0, sc = 0; t < my_tiles; t++) {
                    for (int e = 0; e < 2 * L - 1; e++) {
                        umma::mbar_wait(&rdy[X], rdy_par);
                        rdy_par ^= 1;
                        umma::tc_fence_after();
                        if (e == L - 1) {
                            // output layer: y[n, :] = A_L^T[n, k] . [W_{L+1} hi | lo][k, :]  (FP16).  M = 128 rows = the whole pair image; the
                            // other sub-tile's rows are whatever its image holds and land in TMEM lanes nobody reads
                            if (m == 0) {
                                if (umma::elect_one()) {
                                    umma::gemm_issue_lean<H / 16>(tmem + (uint32_t)(OUTCOL + 16 * X), img_addr, 128, CHS, 256, umma::smem_u32(wo16), 128, CHS, 256,
                                                                  tc3::idesc_f16(128, 16, 1, 1), 0u);
                                    umma::mma_commit(&out_bar[X]);
                                }
                                __syncwarp();
                            }
                        } else {
                            const bool fwd = e < L - 1;
                            #pragma unroll
                            for (int kh = 0; kh < NKH; kh++) {
                                const int i = m * NKH + kh, slot = piece_slot(sc, i);
                                umma::mbar_wait(&wbar[slot], piece_phase(sc));
                                const uint32_t d = tmem + (uint32_t)((X * MT + m) * 64);
                                const uint32_t a = w_addr + (uint32_t)slot * PIECE, b = img_addr + (uint32_t)X * SUB + (uint32_t)kh * 2048u;
                                if (umma::elect_one()) {
                                    if (fwd)   // A = W_l^T piece (MN-major: 8-unit groups 2048 B apart, 8 k rows 128 B apart)
                                        umma::gemm_issue_lean<8>(d, a, 128, 2048, 256, b, 128, CHS, 256, umma::idesc_bf16(128, 64, 1, 1), kh > 0 ? 1u : 0u);
                                    else       // A = W_l rows (K-major: 8-unit chunks of j 2048 B apart, 8 k rows 128 B apart)
                                        umma::gemm_issue_lean<8>(d, a, 2048, 128, 4096, b, 128, CHS, 256, umma::idesc_bf16(128, 64, 0, 1), kh > 0 ? 1u : 0u);
                                    umma::mma_commit(&free_bar[slot]);            // (count 2: BOTH sub-tiles are done with this piece)
                                    if (kh == NKH - 1) umma::mma_commit(&ud[X]);
                                }
                                __syncwarp();
                            }
                            sc++;
                        }
                    }
                }
            }
        } else if (role == 4) {
            // ---------------- loader: as each piece of step s falls free (both sub-tiles' UMMAs on it done), the same piece of the
            // step AHEAD goes into its buffer ----------------
            if (lane == 0) {
                auto request = [&](int s_, int i) {
                    if (s_ >= nsteps) return;
                    const int idx = s_ % spt;
                    const bool fwd = idx < L - 1;
                    const int l = fwd ? 2 + idx : L - (idx - (L - 1));
                    // both mirrors keep a piece contiguous: wf [l][m][kh][16 j-groups][128 k][8 j], wb [l][m][16 kh + j-chunk][128 k][8 j]
                    const __nv_bfloat16* src = (fwd ? ta.wf : ta.wb) + (size_t)(l - 2) * H * H + (size_t)i * (128 * 128);
                    const int b = piece_slot(s_, i);
                    umma::mbar_expect_tx(&wbar[b], PIECE);
                    umma::bulk_g2s(Wb0 + (size_t)b * PIECE, src, PIECE, &wbar[b]);
                };
                for (int s_ = 0; s_ < AHEAD; s_++)
                    for (int i = 0; i < PPS; i++) request(s_, i);
                for (int s_ = 0; s_ < nsteps; s_++)
                    for (int i = 0; i < PPS; i++) {
                        mbar_wait_at(&free_bar[piece_slot(s_, i)], piece_phase(s_), ta.trap_site, 108);
                        request(s_ + AHEAD, i);
                    }
            }
        } else if (role == 5) {
            // ---------------- helper: the hand-off with the dW servers, in the order the sub-tiles' steps are issued ----------------
            // Before the UMMAs of a step may "complete" (second arrival on ud[X]), the memory its epilogue will overwrite must be free:
            //   A_l   (stash image l, l < L; dW_{l+1} operand): rewritten by the forward epilogue of layer l of the NEXT tile -> done[l+1] >= t;
            //   Zbar_l (ring slot item % zdepth): written by the reverse epilogue of layer l -> the item zdepth earlier is done.
            // After the reverse epilogue of layer l of sub-tile X (ready[X] complete): its half of Zbar_l (and, since the forward pass, of
            // A_{l-1}) is in memory -> release ready_X[l] (st.release is cumulative over the stores the mbarrier made visible).
            if (lane == 0) {
                // (an arrival must not land in a phase of ud[X] that is still waiting for its UMMA commit: the helper paces itself on
                //  the completion of the previous phase)
                uint32_t hp[2] = {0, 0}; bool first[2] = {true, true};
                auto arrive_ud = [&](int X) {
                    if (!first[X]) mbar_wait_at(&ud[X], hp[X] ^ 1u, ta.trap_site, 123);
                    first[X] = false;
                    mbar_arrive(&ud[X]);
                    hp[X] ^= 1u;
                };
                for (int t = 0; t < my_tiles; t++)
                    for (int e = 0; e < 2 * L - 1; e++)
                        for (int X = 0; X < 2; X++) {
                            unsigned* my_ready = (X == 0 ? ta.ready : ta.ready2) + (size_t)p * kFlagStride;
                            // follow ready[X] phase by phase, like the issuer (this tile's phases: layer 1, forward 2..L, then reverse
                            // L..2 -> event e waits for phase e): a waiter that skipped phases would read an old phase of the same
                            // parity as "complete"
                            mbar_wait_at(&rdy[X], (uint32_t)((t * (2 * L - 1) + e) & 1), ta.trap_site, 120);
                            if (e < L - 1) {                                  // forward step of layer l = e + 2
                                const int l = e + 2;
                                if (l < L && t > 0) wait_ge_at(dn + (l + 1), (unsigned)t, ta.trap_site, 114);
                                arrive_ud(X);
                            } else if (e >= L) {                              // reverse step after the epilogue of layer l = 2L - e
                                const int l = 2 * L - e;
                                st_release(my_ready + l, (unsigned)t + 1u);
                                if (l > 2) ring_wait(t, l - 1);               // ring slot of Zbar_{l-1}, written by the next epilogue
                                arrive_ud(X);
                            }
                        }
            }
        }
        __syncwarp();
    } else {
        // =====================================================================================
        // epilogue warps
        // =====================================================================================
        if (MT == 2) tc3::setmaxnreg_inc<104>();
        // This thread: unit j; of a sub-tile's 10 points the four of group h (columns 24 h .. 24 h + 23: three whole 16-byte
        // chunks) and the tail point 8 + h (columns 48 + 6 h ..: 12 B of chunk 6 | 4 B of chunk 6 + 8 B of chunk 7 + zero pad).
        float dW1acc[DIN] = {0.0f, 0.0f, 0.0f}, dWoacc[3] = {0.0f, 0.0f, 0.0f};
        float loss_acc = 0.0f, dbo_acc[3] = {0.0f, 0.0f, 0.0f};
        long long tacc[6] = {0, 0, 0, 0, 0, 0}, facc[3] = {0, 0, 0};
        const long long t_begin = clock64();
        const bool tracer = dbg && tid == 0;
        auto epi_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(NEPI) : "memory"); };
        // 2 H single arrivals on one mbarrier serialise (an atomic on one shared-memory word each): one arrival per warp, after the
        // warp's lanes have fenced their own writes and met
        auto warp_arrive = [&](uint64_t* bar) { __syncwarp(); if (lane == 0) mbar_arrive(bar); };

        // ---- one unit's row of an image (shared or global): main group + tail point, 16-bit values ----
        // w[0..11]: the 24 values of the main group; w[12..14]: the 6 values of the tail point
        auto store_row = [&](uint8_t* row, bool global, const uint32_t* w) {
            uint8_t* m = row + (size_t)(3 * h) * CHS;
            uint8_t* c6 = row + (size_t)6 * CHS, *c7 = row + (size_t)7 * CHS;
            if (global) {
                #pragma unroll
                for (int c = 0; c < 3; c++) __stcg(reinterpret_cast<uint4*>(m + (size_t)c * CHS), make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]));
                if (h == 0) { __stcg(reinterpret_cast<uint2*>(c6), make_uint2(w[12], w[13])); __stcg(reinterpret_cast<uint32_t*>(c6 + 8), w[14]); }
                else { __stcg(reinterpret_cast<uint32_t*>(c6 + 12), w[12]); __stcg(reinterpret_cast<uint4*>(c7), make_uint4(w[13], w[14], 0u, 0u)); }
            } else {
                #pragma unroll
                for (int c = 0; c < 3; c++) *reinterpret_cast<uint4*>(m + (size_t)c * CHS) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
                if (h == 0) { *reinterpret_cast<uint2*>(c6) = make_uint2(w[12], w[13]); *reinterpret_cast<uint32_t*>(c6 + 8) = w[14]; }
                else { *reinterpret_cast<uint32_t*>(c6 + 12) = w[12]; *reinterpret_cast<uint4*>(c7) = make_uint4(w[13], w[14], 0u, 0u); }
            }
        };
        auto load_row = [&](const uint8_t* row, bool global, uint32_t* w) {
            const uint8_t* m = row + (size_t)(3 * h) * CHS;
            const uint8_t* c6 = row + (size_t)6 * CHS, *c7 = row + (size_t)7 * CHS;
            #pragma unroll
            for (int c = 0; c < 3; c++) {
                const uint4 v = global ? __ldcg(reinterpret_cast<const uint4*>(m + (size_t)c * CHS)) : *reinterpret_cast<const uint4*>(m + (size_t)c * CHS);
                w[4 * c] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
            }
            if (h == 0) {
                const uint2 v = global ? __ldcg(reinterpret_cast<const uint2*>(c6)) : *reinterpret_cast<const uint2*>(c6);
                w[12] = v.x; w[13] = v.y;
                w[14] = global ? __ldcg(reinterpret_cast<const uint32_t*>(c6 + 8)) : *reinterpret_cast<const uint32_t*>(c6 + 8);
            } else {
                w[12] = global ? __ldcg(reinterpret_cast<const uint32_t*>(c6 + 12)) : *reinterpret_cast<const uint32_t*>(c6 + 12);
                const uint2 v = global ? __ldcg(reinterpret_cast<const uint2*>(c7)) : *reinterpret_cast<const uint2*>(c7);
                w[13] = v.x; w[14] = v.y;
            }
        };
        // this thread's 30 accumulator columns of sub-tile X: main group 24 h .., tail point 48 + 6 h ..
        auto acc_issue = [&](int X, uint32_t* r) {
            const uint32_t base = tmem + lane_addr + (uint32_t)((X * MT + mt) * 64);
            tmem_ld16_nowait(base + (uint32_t)(24 * h), r);
            tmem_ld8_nowait(base + (uint32_t)(24 * h + 16), r + 16);
            tmem_ld8_nowait(base + (uint32_t)(48 + 6 * h), r + 24);          // (6 values + 2 columns of the neighbour / the zero pad)
        };
        auto acc_wait = [&](uint32_t* r) {
            asm volatile("tcgen05.wait::ld.sync.aligned;"
                         : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                           "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                           "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                           "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                         :: "memory");
        };
        auto pt_of = [&](int i) { return i < 4 ? 4 * h + i : 8 + h; };       // point (of the sub-tile) of this thread's i-th point
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
        // forward epilogue of sub-tile X, layer l: pre-activations z[30] -> image row (FP16 for layer L, whose only readers are the
        // output layer and this thread's own reverse step) + stash copy + s stash
        auto fwd_store = [&](int X, int l, const float* z, float bias, uint8_t* grow) {
            uint32_t w[15];
            float sv[5];
            #pragma unroll
            for (int i = 0; i < 5; i++) {                                    // point by point: few values live at a time
                float out[NC];
                sv[i] = fwd_point(z + NC * i, bias, out);
                #pragma unroll
                for (int c = 0; c < 3; c++) w[3 * i + c] = l == L ? tc3::pack2h(out[2 * c], out[2 * c + 1]) : pack2(out[2 * c], out[2 * c + 1]);
            }
            store_row(img + (size_t)X * SUB + (size_t)j * 16, false, w);
            if (grow) store_row(grow, true, w);
            uint2* sp = stS + (size_t)((l - 1) * 6 + X * 3) * H + j;
            __stcg(sp + (size_t)h * H, make_uint2(pack2(sv[0], sv[1]), pack2(sv[2], sv[3])));
            __stcg(reinterpret_cast<uint32_t*>(sp + (size_t)2 * H) + h, __float_as_uint(sv[4]));
        };
        // stash of (sub-tile X, layer l) for the reverse step: 15 words of layer outputs + s (3 words); fetched while the OTHER
        // sub-tile's epilogue runs
        struct Stash { uint32_t a[15]; uint2 s4; uint32_t s1; };
        auto load_stash = [&](int X, int l, size_t a1_off, Stash& st) {
            const uint2* sp = stS + (size_t)((l - 1) * 6 + X * 3) * H + j;
            st.s4 = __ldcg(sp + (size_t)h * H);
            st.s1 = __ldcg(reinterpret_cast<const uint32_t*>(sp + (size_t)2 * H) + h);
            if (l < L) load_row(stA + (l == 1 ? a1_off : (size_t)(l - 1) * OPB) + (size_t)X * SUB + (size_t)j * 16, true, st.a);
            else load_row(img + (size_t)X * SUB + (size_t)j * 16, false, st.a);      // A_L is still the (FP16) operand image
        };

        uint32_t udbits = 0, outbits = 0;                                 // bit X: phase of ud[X] / out_bar[X] this thread waits for next
        Stash held;                                                       // prefetched stash of the NEXT reverse epilogue
        for (int t = 0; t < my_tiles; t++) {
            const size_t a1_off = (t & 1) ? (size_t)(L - 1) * OPB : 0;    // A_1 alternates between two stash copies (the slot of A_L is free)
            for (int ph = 0; ph < 2 * L; ph++) {
                #pragma unroll 1
                for (int X = 0; X < 2; X++) {
                    float* xs = xs_base + X * 64;
                    float* ys = ys_base + X * 256;
                    float* yb = yb_base + X * 256;
                    const unsigned long long base = (unsigned long long)(p + (long long)t * P) * T + (unsigned long long)(X * NP);
                    const size_t grow_off = (size_t)X * SUB + (size_t)j * 16;
                    const long long c_a = tracer ? clock64() : 0;
                    long long c_b = c_a;
                    if (ph == 0) {
                        // ---------------- layer 1 (K = d_in): CUDA cores ----------------
                        // (a tile only waits for the dW_2 server to have taken the tile before the previous one: two copies of A_1)
                        if (tid == 0 && t > 1) wait_ge_at(dn + 2, (unsigned)t - 1u, ta.trap_site, 113);
                        epi_sync();                                        // (the previous tile's layer-1 reverse step has read xs)
                        if (tid < NP * DIN) {
                            const int pt = tid / DIN, k = tid - pt * DIN;
                            xs[pt * 4 + k] = (base + pt < ta.set.n) ? __ldg(ta.set.x + (base + pt) * DIN + k) : 0.0f;
                        }
                        epi_sync();
                        float w1[DIN];
                        #pragma unroll
                        for (int k = 0; k < DIN; k++) w1[k] = __ldg(ta.params + nd.offW[1] + k * H + j);
                        const float b1 = __ldg(ta.params + nd.offB[1] + j);
                        float z[30];
                        #pragma unroll
                        for (int i = 0; i < 5; i++) {
                            const float4 x = *reinterpret_cast<const float4*>(xs + pt_of(i) * 4);
                            z[NC * i] = fmaf(x.x, w1[0], fmaf(x.y, w1[1], x.z * w1[2]));
                            #pragma unroll
                            for (int k = 0; k < DIN; k++) z[NC * i + 1 + k] = w1[k];
                            #pragma unroll
                            for (int k = 0; k < N2; k++) z[NC * i + 1 + DIN + k] = 0.0f;
                        }
                        fwd_store(X, 1, z, b1, stA + a1_off + grow_off);
                        umma::fence_async_smem();
                        warp_arrive(&rdy[X]);
                    } else if (ph <= L - 1) {
                        // ---------------- forward epilogue of hidden layer l ----------------
                        const int l = ph + 1;
                        const float bl = __ldg(ta.params + nd.offB[l] + j);
                        mbar_wait_at(&ud[X], (udbits >> X) & 1u, ta.trap_site, 109); udbits ^= 1u << X;
                        umma::tc_fence_after();
                        c_b = tracer ? clock64() : 0;
                        uint32_t r[32];
                        acc_issue(X, r);
                        acc_wait(r);
                        const long long c_1 = tracer ? clock64() : 0;
                        float z[30];
                        #pragma unroll
                        for (int i = 0; i < 30; i++) z[i] = __uint_as_float(r[i]);
                        fwd_store(X, l, z, bl, l < L ? stA + (size_t)(l - 1) * OPB + grow_off : nullptr);
                        const long long c_2 = tracer ? clock64() : 0;
                        umma::fence_async_smem();
                        umma::tc_fence_before();
                        warp_arrive(&rdy[X]);
                        if (tracer) { const long long c_3 = clock64(); facc[0] += c_1 - c_b; facc[1] += c_2 - c_1; facc[2] += c_3 - c_2; }
                    } else {
                        // ---------------- reverse epilogue of layer l = L .. 1 ----------------
                        const int l = 2 * L - ph;
                        float ab[30];
                        if (l == L) {
                            // output layer first: y out of TMEM (lane n = column n of the pair), residual + seeds per point
                            mbar_wait_at(&out_bar[X], (outbits >> X) & 1u, ta.trap_site, 110); outbits ^= 1u << X;
                            umma::tc_fence_after();
                            c_b = tracer ? clock64() : 0;
                            if (mt == 0 && h == 0 && (q4 >> 1) == X) {
                                uint32_t r[16];
                                tmem_ld16_nowait(tmem + lane_addr + (uint32_t)(OUTCOL + 16 * X), r);
                                tmem_wait16(r);
                                *reinterpret_cast<float4*>(ys + ((q4 & 1) * 32 + lane) * 4) =
                                    make_float4(__uint_as_float(r[0]) + __uint_as_float(r[8]), __uint_as_float(r[1]) + __uint_as_float(r[9]),
                                                __uint_as_float(r[2]) + __uint_as_float(r[10]), 0.0f);
                            }
                            umma::tc_fence_before();
                            load_stash(X, L, a1_off, held);
                            epi_sync();
                            if (tid < NP) {                                  // residual + seeds, one thread per point
                                const int col0 = tid * NC;
                                float y[3][NC], ybar[3][NC], rr[3], x[3];
                                #pragma unroll
                                for (int o = 0; o < 3; o++)
                                    #pragma unroll
                                    for (int c = 0; c < NC; c++)
                                        y[o][c] = (o < nd.d_out) ? ys[(col0 + c) * 4 + o] + (c == 0 ? __ldg(ta.params + nd.offB[L + 1] + o) : 0.0f) : 0.0f;
                                #pragma unroll
                                for (int k = 0; k < DIN; k++) x[k] = xs[tid * 4 + k];
                                float ss = pde_residual<NC>(nd, x, y, ta.set.scale, ybar, rr);
                                const bool valid = base + tid < ta.set.n;
                                if (!valid) ss = 0.0f;
                                #pragma unroll
                                for (int c = 0; c < NC; c++)
                                    *reinterpret_cast<float4*>(yb + (col0 + c) * 4) =
                                        valid ? make_float4(ybar[0][c], ybar[1][c], ybar[2][c], 0.0f) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                                loss_acc += ss;
                                if (valid) {
                                    #pragma unroll
                                    for (int o = 0; o < 3; o++) dbo_acc[o] += ybar[o][0];
                                }
                            }
                            if (tid == 32) ring_wait(t, L);                  // ring slot of Zbar_L
                            epi_sync();
                        } else {
                            mbar_wait_at(&ud[X], (udbits >> X) & 1u, ta.trap_site, 117); udbits ^= 1u << X;
                            umma::tc_fence_after();
                            c_b = tracer ? clock64() : 0;
                            uint32_t r[32];
                            acc_issue(X, r);
                            acc_wait(r);
                            #pragma unroll
                            for (int i = 0; i < 30; i++) ab[i] = __uint_as_float(r[i]);
                        }
                        // point by point (few values live at a time): stash words -> layer outputs, adjoint -> Zbar words
                        float wo[3];
                        if (l == L) {
                            #pragma unroll
                            for (int o = 0; o < 3; o++) wo[o] = (o < nd.d_out) ? __ldg(ta.params + nd.offW[L + 1] + j * nd.d_out + o) : 0.0f;
                        }
                        const float sv[5] = {bf_lo(held.s4.x), bf_hi(held.s4.x), bf_lo(held.s4.y), bf_hi(held.s4.y), __uint_as_float(held.s1)};
                        uint32_t w[15];
                        float dbsum = 0.0f;
                        #pragma unroll
                        for (int i = 0; i < 5; i++) {
                            float av[NC], abp[NC], zb[NC];
                            #pragma unroll
                            for (int c = 0; c < 3; c++) {
                                const uint32_t ww = held.a[3 * i + c];
                                if (l == L) { const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&ww)); av[2 * c] = f.x; av[2 * c + 1] = f.y; }
                                else { av[2 * c] = bf_lo(ww); av[2 * c + 1] = bf_hi(ww); }
                            }
                            if (l == L) {
                                // Abar_L = ybar W_{L+1}^T and dW_{L+1} on the CUDA cores
                                #pragma unroll
                                for (int c = 0; c < NC; c++) {
                                    const float4 s4 = *reinterpret_cast<const float4*>(yb + (pt_of(i) * NC + c) * 4);
                                    abp[c] = fmaf(wo[0], s4.x, fmaf(wo[1], s4.y, wo[2] * s4.z));
                                    dWoacc[0] = fmaf(av[c], s4.x, dWoacc[0]);
                                    dWoacc[1] = fmaf(av[c], s4.y, dWoacc[1]);
                                    dWoacc[2] = fmaf(av[c], s4.z, dWoacc[2]);
                                }
                            } else {
                                #pragma unroll
                                for (int c = 0; c < NC; c++) abp[c] = ab[NC * i + c];
                            }
                            const float z0 = bwd_point(av, sv[i], abp, zb);
                            dbsum += z0;
                            if (l == 1) {
                                const float4 x = *reinterpret_cast<const float4*>(xs + pt_of(i) * 4);
                                dW1acc[0] += fmaf(x.x, z0, zb[1]);
                                dW1acc[1] += fmaf(x.y, z0, zb[2]);
                                dW1acc[2] += fmaf(x.z, z0, zb[3]);
                            }
                            #pragma unroll
                            for (int c = 0; c < 3; c++) w[3 * i + c] = pack2(zb[2 * c], zb[2 * c + 1]);
                        }
                        {
                            // the stash registers are free: fetch for the next reverse epilogue in program order — (X = 0 -> sub-tile B, same
                            // layer) | (X = 1 -> sub-tile A, layer l - 1) — which runs after the other sub-tile's epilogue
                            const int nX = X ^ 1, nl = X == 0 ? l : l - 1;
                            if (nl >= 1 && nl < L) load_stash(nX, nl, a1_off, held);       // (layer L: fetched after the output layer, above)
                        }
                        dbs[(size_t)(h * L + (l - 1)) * H + j] += dbsum;
                        if (l > 1) {
                            store_row(img + grow_off, false, w);
                            store_row(zr + (size_t)(ring_item(t, l) % (unsigned)ta.zdepth) * OPB + grow_off, true, w);
                            umma::fence_async_smem();
                            umma::tc_fence_before();
                            warp_arrive(&rdy[X]);
                        }
                    }
                    if (tracer) {
                        const long long c_c = clock64();
                        tacc[0] += c_b - c_a;
                        tacc[ph == 0 ? 1 : (ph <= L - 1 ? 2 : (ph == L ? 3 : 4))] += c_c - c_b;
                        if (t == kTraceTile) { ta.dbg[160 + (ph * 2 + X) * 3] = c_a; ta.dbg[161 + (ph * 2 + X) * 3] = c_b; ta.dbg[162 + (ph * 2 + X) * 3] = c_c; }
                    }
                }
            }
        }
        if (tracer) { for (int i = 0; i < 5; i++) ta.dbg[i] = tacc[i]; for (int i = 0; i < 3; i++) ta.dbg[8 + i] = facc[i]; ta.dbg[7] = clock64() - t_begin; }

        // ---------------- the per-thread sums meet in shared memory (all UMMAs are complete: the last epilogue waited for them) ----------------
        epi_sync();
        #pragma unroll
        for (int k = 0; k < DIN; k++) ex[(size_t)(h * 6 + k) * H + j] = dW1acc[k];
        #pragma unroll
        for (int o = 0; o < 3; o++) ex[(size_t)(h * 6 + 3 + o) * H + j] = dWoacc[o];
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
    }

    // ---------------- flush the per-CTA sums into this CTA's partial row (single owner per element) ----------------
    __syncthreads();
    for (int i = tid; i < L * H; i += blockDim.x) {
        const int l = i / H + 1, jj = i - (l - 1) * H;
        part[nd.offB[l] + jj] = dbs[(size_t)(l - 1) * H + jj] + dbs[(size_t)(L + (l - 1)) * H + jj];
    }
    if (tid < NEPI && h == 0) {
        // dW_1 [k][j] and dW_{L+1} [j][o]: the two point halves, 0 then 1
        #pragma unroll
        for (int k = 0; k < DIN; k++) part[nd.offW[1] + k * H + j] = ex[(size_t)k * H + j] + ex[(size_t)(6 + k) * H + j];
        for (int o = 0; o < nd.d_out; o++) part[nd.offW[L + 1] + j * nd.d_out + o] = ex[(size_t)(3 + o) * H + j] + ex[(size_t)(9 + o) * H + j];
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 512u);
}

}  // namespace pinn
