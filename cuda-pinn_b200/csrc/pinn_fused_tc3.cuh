// pinn_fused_tc3.cuh — producer CTA of the wide-network tensor-core train step, third form: TWO SUB-TILES IN FLIGHT.
//
// The tc2 producer (pinn_fused_tc2.cuh) walks one 21-point tile through the layers: UMMA -> epilogue -> UMMA ..., so
// the tensor pipe idles during every tanh / Taylor epilogue and the CUDA cores idle during every UMMA (ncu: tensor pipe
// 28 % active, 42 % of the stall samples on the UMMA-done barrier).  Shared memory has no room for a second tile next to
// the 128 KB of layer weights at width 256 — but the UMMA dispatch floor has: a K = 16 step costs ~69 cycles for any
// N <= 128 (profiles/r02_ubench_b200.txt), so one 128-column operand image can be run as two independent 64-column
// SUB-TILES A and B (10 points x 6 Taylor channels = 60 columns + 4 zero columns each) that share the resident weights
// and are otherwise two separate dataflow loops, half a layer step out of phase:
//
//   tensor pipe:   [A m0][A m1]            [B m0][B m1]            [A m0][A m1] ...      (m = 128-unit M-tile of the weight slab)
//   CUDA cores:               epilogue(A) ............  epilogue(B) ............
//
// Nothing is a block barrier: the 256 threads of a sub-tile (one per unit, TMEM lane = unit, all 10 points of the
// sub-tile thread-local) wait on the mbarrier their UMMA blocks commit to, write the next operand image and arrive on the
// sub-tile's "ready" mbarrier; the sub-tile's own issuer thread waits for it and issues the next two blocks.  An issuer
// sits in the last piece of its own sub-tile: tcgen05.mma issue is throttled to the execution rate of the tensor pipe (a
// 16-UMMA block holds its thread ~1,000 cycles), so it returns just when its warp's epilogue can start.  The only
// coupling between A and B is the weight ring: buffer 0 is refilled once [A m0] AND [B m0] are done (mbarrier of count
// two), buffer 1 after both [m1] — every slab is fetched once per layer step for both sub-tiles.
// The output layer runs on the tensor cores from an FP16 image of A_L (11-bit significands, saturating conversion; the
// hidden layers keep BF16) against W_{L+1} split hi | lo in FP16.
// The math, the operand / stash / ring images and the dW-server CTAs are those of tc2 (a tile = the pair A|B = one
// 128-column image); each sub-tile publishes its half of an image through its own ready flag and the server waits for both.
//
// Reference: the reference has no network code (include/network.cuh is empty); the building blocks this replaces are
// tensor::matmul (include/tensor.cuh:380-409) + kernel::matmul (include/device/kernel.cuh:68-83).
#pragma once
#include <cuda_fp16.h>
#include "pinn_fused_tc2.cuh"

namespace pinn {

namespace tc3 {
constexpr int NP = 10, T = 2 * NP, NG = 3;      // points per sub-tile, per tile (pair of sub-tiles); 4-point groups per sub-tile
// Instruction descriptor for kind::f16 with FP16 operands and FP32 accumulation (the output layer).
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t pack2h(float lo, float hi) {          // saturating: a huge tangent becomes +-65504, not inf
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
}  // namespace tc3

template <int DIN, int N2, int HT>
__global__ void __launch_bounds__(2 * HT, 1)
fused_step_tc3(const NetDims nd, const Tc2Args ta) {
    using namespace tc2;
    constexpr int NC = 1 + DIN + N2;
    static_assert(NC == 6 && DIN == 3 && N2 == 2, "built for the six-channel (x, y, t) networks");
    static_assert(HT == 128 || HT == 256, "width 128 or 256");
    constexpr int H = HT, MT = H / 128, NP = tc3::NP, T = tc3::T, NG = tc3::NG, G = 4;
    constexpr uint32_t OPB = (uint32_t)H * 256u;        // image of the pair: [16 n-chunks][H units][16 B]
    constexpr uint32_t SLAB = 128u * H * 2u;
    constexpr uint32_t CHS = (uint32_t)H * 16u;
    constexpr uint32_t SUB = 8u * CHS;                  // one sub-tile's half of an image
    constexpr int NSUB = MT * 128;                      // threads of one sub-tile
    // The weights travel as 32 KB PIECES (one M-tile x one 128-wide half of K) through a ring of NBUF buffers: a piece is
    // refilled as soon as BOTH sub-tiles' UMMAs on it are done, half a block earlier than a whole slab would be — the slack
    // that lets the two sub-tiles stay half a step apart instead of waiting for one another's weights.
    constexpr int NKH = H / 128, PPS = MT * NKH;        // K-halves per M-tile block; pieces per layer step (4 at width 256, 1 at 128)
    constexpr uint32_t PIECE = 128u * 128u * 2u;
    constexpr int NBUF = (int)(2 * SLAB / PIECE);       // 4 at width 256, 2 at width 128
    constexpr int AHEAD = NBUF / PPS;                   // layer steps a refill runs ahead (1 / 2)
    extern __shared__ __align__(128) uint8_t smem3_raw[];
    uint8_t* const smem = smem3_raw;
    __shared__ uint64_t wbar[4], free_bar[4], ud[2], rdy[2], out_bar[2], full_bar[3], empty_bar[3], fin_bar;
    __shared__ uint32_t tmem_slot;
    __shared__ volatile int a_issued;                   // weight pieces sub-tile A's issuer has issued UMMAs on so far

    const int L = nd.L;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q4 = warp & 3, piece = warp >> 2, mt = piece % MT, X = piece / MT;
    const int lt = tid - X * NSUB;                      // index inside the sub-tile
    const int j = mt * 128 + q4 * 32 + lane;            // the unit (TMEM lane q4*32 + lane of M-tile mt) this thread owns
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;

    if (warp == 0) umma::tmem_alloc(&tmem_slot, 512u);
    if (tid == 0) {
        for (int i = 0; i < 4; i++) { umma::mbar_init(&wbar[i], 1); umma::mbar_init(&free_bar[i], 2); }
        umma::mbar_init(&ud[0], 2); umma::mbar_init(&ud[1], 2);
        umma::mbar_init(&rdy[0], NSUB); umma::mbar_init(&rdy[1], NSUB);
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
    float* dbs = reinterpret_cast<float*>(wo16 + H * 32);         // [2 sub-tiles][L][H]: db_l sums of this CTA
    float* ex = reinterpret_cast<float*>(img);                    // [2][6][H]: dW_1 / dW_{L+1} exchange after the last tile
    float* ys = dbs + 2 * L * H + X * 64 * 4;                     // [64 n][4]: outputs y_o of column n of this sub-tile
    float* yb = dbs + 2 * L * H + 128 * 4 + X * 64 * 4;           // [64 n][4]: seeds ybar_o
    float* xs = dbs + 2 * L * H + 256 * 4 + X * 16 * 4;           // [16 points][4]
    uint8_t* stA = ta.stash_a + (size_t)p * L * OPB;
    uint2* stS = reinterpret_cast<uint2*>(ta.stash_s) + (size_t)p * L * (2 * NG) * H;
    uint8_t* zr = ta.zring + (size_t)p * ta.zdepth * OPB;
    unsigned* my_ready = (X == 0 ? ta.ready : ta.ready2) + (size_t)p * kFlagStride;
    const unsigned* dn = ta.done + (size_t)p * kFlagStride;
    const uint32_t img_addr = umma::smem_u32(img), w_addr = umma::smem_u32(Wb0);

    // single-thread roles, each in a different warp, at a point where that warp would be waiting anyway
    // the issuer's WARP (first warp of the sub-tile's last piece) enters issue_step as a whole — a warp-uniform branch — and one
    // elected lane issues; its phase counters are kept by all 32 lanes alike
    const bool issuer_warp = __shfl_sync(0xffffffffu, lt >> 5, 0) == (MT - 1) * 4;
    const bool is_issuer = lt == (MT - 1) * 128;                  // (debug stamps only)
    const bool is_waiter = lt == 32;                              // server back-pressure; second arrival on ud[X]
    const bool is_loader = tid == NSUB + 64;                      // in sub-tile B (the one that runs behind): refills the weight ring

    for (int i = tid; i < 2 * L * H; i += blockDim.x) dbs[i] = 0.0f;
    for (int i = tid; i < H * 16; i += blockDim.x) {              // columns 0..2: fp16(W_{L+1}[k][o]); columns 8..10: its fp16 remainder
        const int og = i / (H * 8), k = (i / 8) % H, o = i & 7;
        const float w = o < nd.d_out ? __ldg(ta.params + nd.offW[L + 1] + k * nd.d_out + o) : 0.0f;
        const __half hi = __float2half_rn(w);
        reinterpret_cast<__half*>(wo16)[i] = og == 0 ? hi : __float2half_rn(w - __half2float(hi));
    }
    float dW1acc[DIN] = {0.0f, 0.0f, 0.0f}, dWoacc[3] = {0.0f, 0.0f, 0.0f};
    float loss_acc = 0.0f, dbo_acc[3] = {0.0f, 0.0f, 0.0f};

    // ---- weight pieces: step s of the per-tile schedule  fwd l = 2..L, then reverse l = L..2;  piece i = (M-tile, K-half) ----
    const int spt = 2 * (L - 1);                                  // layer steps per tile
    const int nsteps = my_tiles * spt;
    auto piece_src = [&](int s_, int i) -> const __nv_bfloat16* {
        const int idx = s_ % spt;
        const bool fwd = idx < L - 1;
        const int l = fwd ? 2 + idx : L - (idx - (L - 1));
        // both mirrors keep a piece contiguous: wf [l][m][kh][16 j-groups][128 k][8 j], wb [l][m][16 kh + j-chunk][128 k][8 j]
        return (fwd ? ta.wf : ta.wb) + (size_t)(l - 2) * H * H + (size_t)i * (128 * 128);
    };
    auto piece_slot = [&](int s_, int i) { return PPS == 4 ? i : ((s_ * PPS + i) % NBUF); };
    auto piece_phase = [&](int s_) { return (uint32_t)((s_ * PPS / NBUF) & 1); };
    auto request = [&](int s_, int i) {
        if (s_ >= nsteps) return;
        const int b = piece_slot(s_, i);
        umma::mbar_expect_tx(&wbar[b], PIECE);
        umma::bulk_g2s(Wb0 + (size_t)b * PIECE, piece_src(s_, i), PIECE, &wbar[b]);
    };
    if (tid == 0) a_issued = 0;
    if (tid == 0)
        for (int s_ = 0; s_ < AHEAD; s_++)
            for (int i = 0; i < PPS; i++) request(s_, i);
    __syncthreads();
    auto sub_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + X), "r"(NSUB) : "memory"); };   // the threads of this sub-tile

    // This thread: unit j, the NP points of sub-tile X (columns 6 pt + c of the sub-tile's 64), in groups of 4 / 4 / 2 points.
    uint8_t* const img_row = img + (size_t)X * SUB + (size_t)j * 16;
    const uint32_t tm_acc = tmem + lane_addr + (uint32_t)((X * MT + mt) * 64);
    const uint32_t out_col = (uint32_t)(OUTCOL + 16 * X);
    const size_t grow_off = (size_t)X * SUB + (size_t)j * 16;     // the same row of a global image (stash / ring)
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, racc[4] = {0, 0, 0, 0};
    const bool dbg = ta.dbg != nullptr && blockIdx.x == 0;
    constexpr int kTraceTile = 5;                                 // (debug) the tile whose timeline is recorded
    const bool tracer = dbg && (tid & 127) == 0;                  // first thread of each piece
    auto trace = [&](int t_, int k, int which) { if (tracer && t_ == kTraceTile) ta.dbg[272 + (piece * 16 + k) * 2 + which] = clock64(); };
    const long long t_begin = clock64();

    auto point_ok = [&](int g, int pp) { return G * g + pp < NP; };
    auto chunk_ok = [&](int g, int c) { return 3 * g + c < 8; };
    // 24 FP32 results of group g -> 16-bit chunks of this unit's row of an image in shared memory and / or its copy in global memory
    auto store_group = [&](uint8_t* row, uint8_t* grow, int g, const float* out, bool f16 = false) {
        #pragma unroll
        for (int c = 0; c < 3; c++)
            if (chunk_ok(g, c)) {
                uint4 w;
                if (f16) w = make_uint4(tc3::pack2h(out[8 * c], out[8 * c + 1]), tc3::pack2h(out[8 * c + 2], out[8 * c + 3]),
                                        tc3::pack2h(out[8 * c + 4], out[8 * c + 5]), tc3::pack2h(out[8 * c + 6], out[8 * c + 7]));
                else w = make_uint4(pack2(out[8 * c], out[8 * c + 1]), pack2(out[8 * c + 2], out[8 * c + 3]),
                                    pack2(out[8 * c + 4], out[8 * c + 5]), pack2(out[8 * c + 6], out[8 * c + 7]));
                if (row) *reinterpret_cast<uint4*>(row + (size_t)(3 * g + c) * CHS) = w;
                if (grow) __stcg(reinterpret_cast<uint4*>(grow + (size_t)(3 * g + c) * CHS), w);
            }
    };
    auto unpack_chunk = [&](const uint4& w, float* v, bool f16) {
        if (f16) {
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
            #pragma unroll
            for (int i = 0; i < 4; i++) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&ww[i]));
                v[2 * i] = f.x; v[2 * i + 1] = f.y;
            }
        } else {
            v[0] = bf_lo(w.x); v[1] = bf_hi(w.x); v[2] = bf_lo(w.y); v[3] = bf_hi(w.y);
            v[4] = bf_lo(w.z); v[5] = bf_hi(w.z); v[6] = bf_lo(w.w); v[7] = bf_hi(w.w);
        }
    };
    // accumulator columns of group g of this unit's TMEM lane (the last group: 12 values + 4 zero columns)
    auto acc_issue = [&](int g, uint32_t* r) {
        tmem_ld16_nowait(tm_acc + (uint32_t)(24 * g), r);
        if (g < NG - 1) tmem_ld8_nowait(tm_acc + (uint32_t)(24 * g + 16), r + 16);
    };
    auto acc_wait = [&](int g, uint32_t* r) {
        if (g < NG - 1) tmem_wait24(r); else tmem_wait16(r);
    };
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
    // forward epilogue of one group: pre-activations z -> operand chunks (FP16 for layer L, whose only readers are the output
    // layer and this thread's own reverse step) + s stash
    auto fwd_group = [&](int l, int g, const float* z, float bias, bool f16, uint8_t* grow) {
        float out[24], sv[4];
        #pragma unroll
        for (int pp = 0; pp < G; pp++) {
            if (point_ok(g, pp)) sv[pp] = fwd_point(z + NC * pp, bias, out + NC * pp);
            else {
                sv[pp] = 0.0f;
                #pragma unroll
                for (int c = 0; c < NC; c++) out[NC * pp + c] = 0.0f;
            }
        }
        store_group(img_row, grow, g, out, f16);
        stS[(size_t)((l - 1) * (2 * NG) + X * NG + g) * H + j] = make_uint2(pack2(sv[0], sv[1]), pack2(sv[2], sv[3]));
    };
    auto ring_item = [&](int t_, int l) { return (unsigned)t_ * (unsigned)(L - 1) + (unsigned)(L - l); };
    auto ring_wait = [&](int t_, int l) {
        const unsigned item = ring_item(t_, l);
        if (item >= (unsigned)ta.zdepth) {
            const unsigned old = item - (unsigned)ta.zdepth;
            wait_ge_at(dn + (L - (int)(old % (unsigned)(L - 1))), old / (unsigned)(L - 1) + 1u, ta.trap_site, 112);
        }
    };
    auto ring_row = [&](int t_, int l) -> uint8_t* {
        return zr + (size_t)(ring_item(t_, l) % (unsigned)ta.zdepth) * OPB + grow_off;
    };

    // ---- issuer of sub-tile X: the UMMA blocks [X m0] [X m1] of layer step s_ ----
    uint32_t rdy_par = 0;                                         // (issuer) phase of ready[X] it waits for next
    auto issue_step = [&](int s_, bool fwd) {               // (all lanes of the issuer's warp)
        const long long c0 = dbg ? clock64() : 0;
        mbar_wait_at(&rdy[X], rdy_par, ta.trap_site, 105);
        rdy_par ^= 1;
        // B issues a step only after A has issued it: B runs (at least) A's two blocks behind, so that each sub-tile's epilogue
        // runs under the other's UMMA blocks instead of both epilogues running while the tensor pipe idles
        if (X == 1 && ta.stagger > 0) {
            const long long t0 = clock64();
            while (a_issued < s_ * PPS + (ta.stagger < PPS ? ta.stagger : PPS)) {
                __nanosleep(32);
                if (clock64() - t0 > g_wait_bound) trap_at(ta.trap_site, 121, s_, a_issued);
            }
        }
        const long long c1 = dbg ? clock64() : 0;
        long long cwait = 0;
        umma::tc_fence_after();
        #pragma unroll
        for (int i = 0; i < PPS; i++) {
            const int m = i / NKH, kh = i % NKH, slot = piece_slot(s_, i);
            const long long cw0 = dbg ? clock64() : 0;
            mbar_wait_at(&wbar[slot], piece_phase(s_), ta.trap_site, 104);
            if (dbg) cwait += clock64() - cw0;
            const uint32_t d = tmem + (uint32_t)((X * MT + m) * 64);
            const uint32_t a = w_addr + (uint32_t)slot * PIECE, b = img_addr + (uint32_t)X * SUB + (uint32_t)kh * 2048u;
            if (umma::elect_one()) {
                if (fwd)   // A = W_l^T piece (MN-major: 8-unit groups 2048 B apart, 8 k rows 128 B apart)
                    umma::gemm_issue_lean<8>(d, a, 128, 2048, 256, b, 128, CHS, 256, umma::idesc_bf16(128, 64, 1, 1), kh > 0 ? 1u : 0u);
                else       // A = W_l rows (K-major: 8-unit chunks of j 2048 B apart, 8 k rows 128 B apart)
                    umma::gemm_issue_lean<8>(d, a, 2048, 128, 4096, b, 128, CHS, 256, umma::idesc_bf16(128, 64, 0, 1), kh > 0 ? 1u : 0u);
                umma::mma_commit(&free_bar[slot]);                // (count 2: BOTH sub-tiles are done with this piece)
                if (i == PPS - 1) umma::mma_commit(&ud[X]);
            }
            if (X == 0 && lane == 0) a_issued = s_ * PPS + i + 1;
            __syncwarp();
        }
        if (dbg && is_issuer) {
            const long long c4 = clock64(); tacc[0] += c1 - c0; tacc[1] += cwait; tacc[2] += (c4 - c1) - cwait;
            if (s_ / spt == kTraceTile) { long long* e = ta.dbg + 16 + ((s_ % spt) * 2 + X) * 8; e[0] = c0; e[1] = c1; e[2] = c1 + cwait; e[3] = c4; e[4] = c4; e[5] = c4; }
        }
    };
    uint32_t step_par = 0, out_par = 0, pub_par = 0;              // step_par: phase of ud[X]; pub_par: phase of ready[X] this thread arrives on next
    // the loader, before it joins its own sub-tile's wait for step s_: as each piece of step s_ falls free (both sub-tiles' UMMAs
    // on it done — the last one about when the loader's own blocks complete), the same piece of the step AHEAD goes into its buffer
    auto refill = [&](int s_) {
        if (!is_loader) return;
        #pragma unroll
        for (int i = 0; i < PPS; i++) {
            mbar_wait_at(&free_bar[piece_slot(s_, i)], piece_phase(s_), ta.trap_site, 108);
            request(s_ + AHEAD, i);
        }
    };

    int sc = 0;
    for (int t = 0; t < my_tiles; t++) {
        const unsigned long long base = (unsigned long long)(p + (long long)t * P) * T + (unsigned long long)(X * NP);
        // A_1 alternates between two stash copies (the slot of A_L is free: A_L is never stashed), so a tile only waits
        // for the dW_2 server to have taken the tile before the previous one
        if (is_waiter && t > 1) wait_ge_at(dn + 2, (unsigned)t - 1u, ta.trap_site, 113);
        const size_t a1_off = (t & 1) ? (size_t)(L - 1) * OPB : 0;
        if (lt < NP * DIN) {
            const int pt = lt / DIN, k = lt - pt * DIN;
            xs[pt * 4 + k] = (base + pt < ta.set.n) ? __ldg(ta.set.x + (base + pt) * DIN + k) : 0.0f;
        }
        sub_sync();

        // ---------------- layer 1 (K = d_in): CUDA cores ----------------
        trace(t, 2 * L - 1, 0);
        {
            float w1[DIN];
            #pragma unroll
            for (int k = 0; k < DIN; k++) w1[k] = __ldg(ta.params + nd.offW[1] + k * H + j);
            const float b1 = __ldg(ta.params + nd.offB[1] + j);
            #pragma unroll
            for (int g = 0; g < NG; g++) {
                float z[24];
                #pragma unroll
                for (int pp = 0; pp < G; pp++) {
                    const int pt = point_ok(g, pp) ? G * g + pp : NP - 1;
                    const float4 x = *reinterpret_cast<const float4*>(xs + pt * 4);
                    z[NC * pp] = fmaf(x.x, w1[0], fmaf(x.y, w1[1], x.z * w1[2]));
                    #pragma unroll
                    for (int i = 0; i < DIN; i++) z[NC * pp + 1 + i] = w1[i];
                    #pragma unroll
                    for (int i = 0; i < N2; i++) z[NC * pp + 1 + DIN + i] = 0.0f;
                }
                fwd_group(1, g, z, b1, false, stA + a1_off + grow_off);
            }
        }
        umma::fence_async_smem();
        mbar_arrive(&rdy[X]); pub_par ^= 1;
        trace(t, 2 * L - 1, 1);

        // ---------------- hidden layers 2..L: tensor cores ----------------
        for (int l = 2; l <= L; l++, sc++) {
            if (issuer_warp) issue_step(sc, true);
            if (is_waiter) {
                if (l < L && t > 0) wait_ge_at(dn + (l + 1), (unsigned)t, ta.trap_site, 114);   // A_l of the previous tile may be overwritten
                mbar_arrive(&ud[X]);
            }
            const float bl = __ldg(ta.params + nd.offB[l] + j);
            const long long c_a = dbg ? clock64() : 0;
            refill(sc);
            mbar_wait_at(&ud[X], step_par, ta.trap_site, 109);
            umma::tc_fence_after();
            const long long c_b = dbg ? clock64() : 0;
            trace(t, l - 2, 0);
            {
                uint32_t rz[2][24];
                acc_issue(0, rz[0]);
                #pragma unroll
                for (int g = 0; g < NG; g++) {
                    acc_wait(g, rz[g & 1]);
                    if (g < NG - 1) acc_issue(g + 1, rz[(g + 1) & 1]);
                    float z[24];
                    #pragma unroll
                    for (int i = 0; i < 24; i++) z[i] = __uint_as_float(rz[g & 1][i]);
                    fwd_group(l, g, z, bl, l == L, l < L ? stA + (size_t)(l - 1) * OPB + grow_off : nullptr);
                }
            }
            step_par ^= 1;
            umma::fence_async_smem();
            umma::tc_fence_before();
            mbar_arrive(&rdy[X]); pub_par ^= 1;
            trace(t, l - 2, 1);
            if (dbg && tid == 0) { const long long c_c = clock64(); tacc[3] += c_b - c_a; tacc[4] += c_c - c_b; }
        }

        // ---------------- output layer on the tensor cores: y[n, :] = A_L^T[n, k] . [W_{L+1} hi | lo][k, :]  (FP16) ----------------
        // (M = 128 rows = the whole pair image; the other sub-tile's rows are whatever its image holds and land in TMEM lanes
        //  nobody reads)
        const long long c_o = dbg ? clock64() : 0;
        if (issuer_warp) {
            mbar_wait_at(&rdy[X], rdy_par, ta.trap_site, 115);
            rdy_par ^= 1;
            umma::tc_fence_after();
            if (umma::elect_one()) {
                umma::gemm_issue_lean<H / 16>(tmem + out_col, img_addr, 128, CHS, 256, umma::smem_u32(wo16), 128, CHS, 256,
                                              tc3::idesc_f16(128, 16, 1, 1), 0u);
                umma::mma_commit(&out_bar[X]);
            }
            __syncwarp();
        }
        mbar_wait_at(&out_bar[X], out_par, ta.trap_site, 110); out_par ^= 1;
        umma::tc_fence_after();
        if (mt == 0 && (q4 >> 1) == X) {             // TMEM lane n = column n of the pair: this sub-tile's 64 lanes
            uint32_t r[16];
            tmem_ld16_nowait(tmem + lane_addr + out_col, r);
            tmem_wait16(r);
            *reinterpret_cast<float4*>(ys + ((q4 & 1) * 32 + lane) * 4) =
                make_float4(__uint_as_float(r[0]) + __uint_as_float(r[8]), __uint_as_float(r[1]) + __uint_as_float(r[9]),
                            __uint_as_float(r[2]) + __uint_as_float(r[10]), 0.0f);
        }
        umma::tc_fence_before();
        sub_sync();
        if (lt < NP) {                               // residual + seeds, one thread per point
            const int col0 = lt * NC;
            float y[3][NC], ybar[3][NC], r[3], x[3];
            #pragma unroll
            for (int o = 0; o < 3; o++)
                #pragma unroll
                for (int c = 0; c < NC; c++)
                    y[o][c] = (o < nd.d_out) ? ys[(col0 + c) * 4 + o] + (c == 0 ? __ldg(ta.params + nd.offB[L + 1] + o) : 0.0f) : 0.0f;
            #pragma unroll
            for (int k = 0; k < DIN; k++) x[k] = xs[lt * 4 + k];
            float ss = pde_residual<NC>(nd, x, y, ta.set.scale, ybar, r);
            const bool valid = base + lt < ta.set.n;
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
        if (is_waiter) ring_wait(t, L);              // ring slot of Zbar_L
        sub_sync();
        if (dbg && tid == 0) tacc[6] += clock64() - c_o;

        // ---------------- reverse pass ----------------
        float wo[3];
        #pragma unroll
        for (int o = 0; o < 3; o++) wo[o] = (o < nd.d_out) ? __ldg(ta.params + nd.offW[L + 1] + j * nd.d_out + o) : 0.0f;
        // Stash of a reverse step (s, and the layer outputs A_l): only GROUP 0 is fetched ahead and held across the wait for the
        // UMMAs (14 registers); inside the epilogue the stash of group g+1 — at the end, group 0 of the next step — is in flight
        // while group g is worked on.  (Holding all three groups across the wait cost 42 registers and spilled.)
        auto load_stash = [&](int l, int g, uint4* pa_, uint2& ps_) {
            ps_ = __ldcg(stS + (size_t)((l - 1) * (2 * NG) + X * NG + g) * H + j);
            #pragma unroll
            for (int c = 0; c < 3; c++) {
                pa_[c] = make_uint4(0u, 0u, 0u, 0u);
                if (chunk_ok(g, c)) {
                    if (l < L) pa_[c] = __ldcg(reinterpret_cast<const uint4*>(stA + (l == 1 ? a1_off : (size_t)(l - 1) * OPB) + grow_off + (size_t)(3 * g + c) * CHS));
                    else pa_[c] = *reinterpret_cast<const uint4*>(img_row + (size_t)(3 * g + c) * CHS);   // A_L is still the (FP16) operand image
                }
            }
        };
        uint4 pc[3]; uint2 psc;
        load_stash(L, 0, pc, psc);
        for (int l = L; l >= 1; l--) {
            const long long c_d = dbg ? clock64() : 0;
            if (l < L) {
                refill(sc - 1);
                mbar_wait_at(&ud[X], step_par, ta.trap_site, 117);
                step_par ^= 1;
                umma::tc_fence_after();
            }
            const long long c_e = dbg ? clock64() : 0;
            trace(t, (L - 1) + (L - l), 0);
            float dbsum = 0.0f, dw1[DIN] = {0.0f, 0.0f, 0.0f};
            uint8_t* const zrow = l > 1 ? ring_row(t, l) : nullptr;
            #pragma unroll
            for (int g = 0; g < NG; g++) {
                float av[24], ab[24], zb[24];
                uint4 pn[3]; uint2 psn;
                if (g < NG - 1) load_stash(l, g + 1, pn, psn);
                else if (l > 1) load_stash(l - 1, 0, pn, psn);
                #pragma unroll
                for (int c = 0; c < 3; c++) unpack_chunk(pc[c], av + 8 * c, l == L);
                const float sv[4] = {bf_lo(psc.x), bf_hi(psc.x), bf_lo(psc.y), bf_hi(psc.y)};
                if (l == L) {
                    // Abar_L = ybar W_{L+1}^T and dW_{L+1} on the CUDA cores
                    #pragma unroll
                    for (int i = 0; i < 24; i++) {
                        if (point_ok(g, i / NC)) {
                            const float4 s4 = *reinterpret_cast<const float4*>(yb + (24 * g + i) * 4);
                            ab[i] = fmaf(wo[0], s4.x, fmaf(wo[1], s4.y, wo[2] * s4.z));
                            dWoacc[0] = fmaf(av[i], s4.x, dWoacc[0]);
                            dWoacc[1] = fmaf(av[i], s4.y, dWoacc[1]);
                            dWoacc[2] = fmaf(av[i], s4.z, dWoacc[2]);
                        } else ab[i] = 0.0f;
                    }
                } else {
                    const long long c_x = (dbg && tid == 0) ? clock64() : 0;
                    uint32_t r[24];
                    acc_issue(g, r);
                    acc_wait(g, r);
                    #pragma unroll
                    for (int i = 0; i < 24; i++) ab[i] = (g < NG - 1 || i < 16) ? __uint_as_float(r[i]) : 0.0f;
                    if (dbg && tid == 0) racc[0] += clock64() - c_x;
                }
                #pragma unroll
                for (int pp = 0; pp < G; pp++) {
                    if (point_ok(g, pp)) {
                        const float z0 = bwd_point(av + NC * pp, sv[pp], ab + NC * pp, zb + NC * pp);
                        dbsum += z0;
                        if (l == 1) {
                            const float4 x = *reinterpret_cast<const float4*>(xs + (G * g + pp) * 4);
                            dw1[0] += fmaf(x.x, z0, zb[NC * pp + 1]);
                            dw1[1] += fmaf(x.y, z0, zb[NC * pp + 2]);
                            dw1[2] += fmaf(x.z, z0, zb[NC * pp + 3]);
                        }
                    } else {
                        #pragma unroll
                        for (int c = 0; c < NC; c++) zb[NC * pp + c] = 0.0f;
                    }
                }
                if (l > 1) store_group(img_row, zrow, g, zb);
                #pragma unroll
                for (int c = 0; c < 3; c++) pc[c] = pn[c];
                psc = psn;
            }
            dbs[(size_t)(X * L + (l - 1)) * H + j] += dbsum;
            if (l == 1) {
                trace(t, (L - 1) + (L - l), 1);
                #pragma unroll
                for (int k = 0; k < DIN; k++) dW1acc[k] += dw1[k];
                if (dbg && tid == 0) { const long long c_f = clock64(); tacc[3] += c_e - c_d; tacc[5] += c_f - c_e; }
                break;
            }
            const long long c_q = (dbg && tid == 0) ? clock64() : 0;
            umma::fence_async_smem();
            umma::tc_fence_before();
            mbar_arrive(&rdy[X]); pub_par ^= 1;
            if (dbg && tid == 0) { const long long c_r = clock64(); racc[2] += c_r - c_q; racc[3] += c_q - c_e; }
            trace(t, (L - 1) + (L - l), 1);
            if (dbg && tid == 0) { const long long c_f = clock64(); tacc[3] += c_e - c_d; tacc[5] += c_f - c_e; }
            if (issuer_warp) issue_step(sc, false);
            if (is_waiter) {
                // Every thread of the sub-tile has stored its part of Zbar_l (and, in the forward pass, of A_{l-1}): release this
                // half of the images to the dW server.  (A second waiter on ready[X] next to the issuer; the barrier cannot run
                // ahead, its next phase needs this thread.  st.release is cumulative over the stores the mbarrier made visible;
                // its cost — draining the SM's outstanding stores — stays off the issuer.)
                mbar_wait_at(&rdy[X], pub_par ^ 1u, ta.trap_site, 120);
                st_release(my_ready + l, (unsigned)t + 1u);
                if (l > 2) ring_wait(t, l - 1);          // ring slot of Zbar_{l-1}, written by the next epilogue
                mbar_arrive(&ud[X]);
            }
            sc++;
        }
        sub_sync();                                  // xs / ys / yb are rewritten by the next tile
    }
    if (dbg) {
        if (is_issuer && X == 0) { ta.dbg[0] = tacc[0]; ta.dbg[1] = tacc[1]; ta.dbg[2] = tacc[2]; }
        if (tid == 0) { for (int i = 0; i < 4; i++) ta.dbg[8 + i] = racc[i]; ta.dbg[3] = tacc[3]; ta.dbg[4] = tacc[4]; ta.dbg[5] = tacc[5]; ta.dbg[6] = tacc[6]; ta.dbg[7] = clock64() - t_begin; }
    }

    // ---------------- flush the per-CTA sums into this CTA's partial row (single owner per element) ----------------
    __syncthreads();
    for (int i = tid; i < L * H; i += blockDim.x) {
        const int l = i / H + 1, jj = i - (l - 1) * H;
        part[nd.offB[l] + jj] = dbs[(size_t)(l - 1) * H + jj] + dbs[(size_t)(L + (l - 1)) * H + jj];
    }
    // dW_1 [k][j] and dW_{L+1} [j][o]: the two sub-tiles meet in shared memory, A then B
    #pragma unroll
    for (int k = 0; k < DIN; k++) ex[(size_t)(X * 6 + k) * H + j] = dW1acc[k];
    #pragma unroll
    for (int o = 0; o < 3; o++) ex[(size_t)(X * 6 + 3 + o) * H + j] = dWoacc[o];
    // loss and output-bias sums: sub-tile A's first warp, then sub-tile B's, through two slots behind the exchange rows
    float* red = ex + 12 * H;
    if (lt < 32) {
        float v[4] = {loss_acc, dbo_acc[0], dbo_acc[1], dbo_acc[2]};
        #pragma unroll
        for (int i = 0; i < 4; i++)
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
        if (lane == 0)
            for (int i = 0; i < 4; i++) red[X * 4 + i] = v[i];
    }
    __syncthreads();
    if (X == 0) {
        #pragma unroll
        for (int k = 0; k < DIN; k++) part[nd.offW[1] + k * H + j] = ex[(size_t)k * H + j] + ex[(size_t)(6 + k) * H + j];
        for (int o = 0; o < nd.d_out; o++) part[nd.offW[L + 1] + j * nd.d_out + o] = ex[(size_t)(3 + o) * H + j] + ex[(size_t)(9 + o) * H + j];
    }
    if (tid == 0) {
        part[nd.P + ta.set.set] = red[0] + red[4];
        for (int o = 0; o < nd.d_out; o++) part[nd.offB[L + 1] + o] = red[1 + o] + red[5 + o];
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 512u);
}

}  // namespace pinn
