// pinn_fused_tc.cuh — fused PINN train step with the hidden-layer contractions on the 5th-gen
// tensor cores (tcgen05.mma, BF16 operands, FP32 accumulation in TMEM).  Widths 64 / 128 / 256.
//
// Same per-tile dataflow as the FP32 kernel (pinn_fused_fp32.cuh): a CTA owns a tile of TP = 16
// collocation points; row m = c*16 + p of the 128-row UMMA tile is Taylor channel c of point p
// (rows >= NC*16 are zero padding).  Per hidden->hidden layer l the three contractions
//      Z_l      = A_{l-1} W_l            (M = rows, N = units, K = units;  A K-major, W MN-major)
//      Abar_l-1 = Zbar_l  W_l^T          (A K-major, W K-major)
//      dW_l     = A_{l-1}^T Zbar_l       (both MN-major, K = the 128 rows)
// are single-thread tcgen05.mma sequences on operands that sit in shared memory ONCE, in the
// chunk-major no-swizzle layout of pinn_umma.cuh.  Weights arrive by cp.async.bulk (TMA 1-D) from a
// BF16 chunk-major mirror in global memory (maintained next to the FP32 master copy); width 256
// streams them in two 64 KiB halves.  The tanh-Taylor elementwise work, the first (K = d_in) and
// last (N = d_out) layers, the loss and all reductions stay on the CUDA cores in FP32, thread
// (p, g) owning all NC channels of 8 units — accumulator rows come back from TMEM through an FP32
// staging area that aliases the operand buffers once the MMAs that read them have completed.
//
// NTERMS = 1: plain BF16 operands.  NTERMS = 2: each operand is split x = hi + mid (two BF16
// terms, 16 mantissa bits) and every contraction issues hi*hi + hi*mid + mid*hi — near-FP32
// products at 3x the (cheap) tensor work.
#pragma once
#include "pinn_device.cuh"
#include "pinn_fused_fp32.cuh"
#include "pinn_umma.cuh"

namespace pinn {

struct TcArgs {
    FusedArgs f;
    const __nv_bfloat16* wq;     // [NTERMS][L-1][H*H] chunk-major BF16 mirror of W_2..W_L
    long long wq_term_stride;    // elements between the hi and mid mirrors
    int tmem_cols;               // TMEM columns to allocate (power of two)
    // Deferred dW (width 256): the two BF16 dW operands of every tile-layer go to a per-CTA ring by
    // TMA bulk store; every `defer_g` tiles a catch-up pass streams them back and accumulates dW_l in
    // TMEM over all of them, so the 256 KB flush happens once per defer_g tiles instead of every tile.
    uint8_t* op_ring;            // [grid][defer_g][L-1][2][ (H/8) * RV * 16 bytes ]
    long long ring_per_cta;      // bytes
    int defer_g;                 // 0 = flush dW every tile
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}

// UT = units per thread (8: one 16-byte operand chunk per channel, 2H threads; 4: half a chunk,
// 4H threads at <= 64 registers — twice the resident warps to hide the serialized phase latencies).
// HT: hidden width fixed at compile time (64 / 128 / 256): offsets fold to immediates.
// VAL: value-channel-only tiles for the boundary / initial sets.  The NC "channels" of a tile are then 8
// independent groups of 16 points (row m = c*16 + p is point base + c*16 + p): 128 points fill the UMMA tile,
// the elementwise work is a plain tanh per row, and the loss is the Dirichlet / initial target misfit.
template <int DIN, int N2, int NTERMS, int UT, int HT, bool VAL = false>
__global__ void __launch_bounds__(16 * HT / UT, 512 / (16 * HT / UT))
fused_step_tc(const NetDims nd, const TcArgs ta) {
    constexpr int NC = VAL ? 8 : 1 + DIN + N2, TP = 16, R = 128;
    static_assert(UT == 8 || UT == 4, "unit group of 4 or 8");
    constexpr int LD = VAL ? NC * TP : NC * TP + 2;   // staging leading dimension (floats); VAL: 128 rows fill the
                                                      // aliased operand buffers exactly (H * 512 B), so no padding there
    constexpr int RV = NC * TP;                       // valid rows of the 128-row tile
    const FusedArgs& args = ta.f;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t mma_bar, w_bar, ld_bar[2], done_bar[2];
    __shared__ uint32_t tmem_slot;

    constexpr int H = HT;
    const int L = nd.L;
    constexpr int HW = H > 128 ? 128 : H, nhalf = H / HW, mt = (H + 127) >> 7;
    constexpr uint32_t opBytes = (uint32_t)H * 256u;  // one operand term: [H/8][128][16 B]
    constexpr uint32_t wBytes = (uint32_t)HW * H * 2u; // one weight term (one half for H = 256)
    uint8_t* bufA = smem_raw;
    uint8_t* bufZ = bufA + NTERMS * opBytes;
    uint8_t* bufW = bufZ + NTERMS * opBytes;
    float* Zs = reinterpret_cast<float*>(smem_raw);   // FP32 staging [unit][LD], aliases bufA|bufZ
    float* xbuf = reinterpret_cast<float*>(smem_raw); // output-layer exchange [H/UT][3][NC][16] + seeds, aliases the dead operands

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int p = tid % TP, g = tid / TP, j0 = g * UT;
    constexpr int nwarps = (16 * H / UT) >> 5, nsub = nwarps >> 2;
    const int lg = warp & 3, wsub = warp >> 2;
    float* stash = args.stash + (size_t)(args.row0 + blockIdx.x) * args.stash_per_cta;
    float* part = args.partial + (size_t)(args.row0 + blockIdx.x) * nd.PP;
    constexpr int lstride = NC * H * TP;

    if (warp == 0) umma::tmem_alloc(&tmem_slot, (uint32_t)ta.tmem_cols);
    if (tid == 0) {
        umma::mbar_init(&mma_bar, 1); umma::mbar_init(&w_bar, 1);
        umma::mbar_init(&ld_bar[0], 1); umma::mbar_init(&ld_bar[1], 1); umma::mbar_init(&done_bar[0], 1); umma::mbar_init(&done_bar[1], 1);
        umma::mbar_fence_init();
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t a_addr = umma::smem_u32(bufA), z_addr = umma::smem_u32(bufZ), w_addr = umma::smem_u32(bufW);
    const uint32_t da_off = (mt * H + H <= ta.tmem_cols) ? (uint32_t)(mt * H) : 0u;   // dA accumulator columns
    const bool da_separate = da_off != 0u;
    uint32_t mma_phase = 0, w_phase = 0;
    int w_cur = -1; bool w_pending = false;           // meaningful in thread 0 only

    // ---- weight staging (thread 0): cp.async.bulk of layer l (2..L), half h, all terms ----
    auto w_request = [&](int l, int h) {
        const int key = l * 2 + h;
        if (key == w_cur) return;
        if (w_pending) { umma::mbar_wait(&w_bar, w_phase); w_phase ^= 1; w_pending = false; }   // one copy in flight at a time
        umma::mbar_expect_tx(&w_bar, wBytes * NTERMS);
        #pragma unroll
        for (int t = 0; t < NTERMS; t++)
            umma::bulk_g2s(bufW + t * wBytes,
                           ta.wq + t * ta.wq_term_stride + (size_t)(l - 2) * H * H + (size_t)h * HW * H, wBytes, &w_bar);
        w_cur = key; w_pending = true;
    };
    auto w_ready = [&]() {
        if (w_pending) { umma::mbar_wait(&w_bar, w_phase); w_phase ^= 1; w_pending = false; }
    };
    // ---- operand write: thread (p, g) stores its 8 units of every channel as one 16-byte chunk ----
    auto write_operand = [&](uint8_t* buf, const float (&v)[NC][UT], bool zero_pad) {
        // chunk of 8 features = 16 bytes; this thread owns UT of them starting at feature j0
        const size_t coff = (size_t)(j0 >> 3) * (R * 16) + (size_t)(j0 & 7) * 2;
        #pragma unroll
        for (int c = 0; c < NC; c++) {
            uint8_t* dst = buf + coff + (size_t)(c * TP + p) * 16;
            uint32_t hi[UT / 2];
            #pragma unroll
            for (int h = 0; h < UT / 2; h++) hi[h] = pack_bf16x2(v[c][2 * h], v[c][2 * h + 1]);
            if constexpr (UT == 8) *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            else *reinterpret_cast<uint2*>(dst) = make_uint2(hi[0], hi[1]);
            if constexpr (NTERMS == 2) {
                uint32_t mid[UT / 2];
                #pragma unroll
                for (int h = 0; h < UT / 2; h++) {
                    const float r0 = v[c][2 * h] - __bfloat162float(__float2bfloat16(v[c][2 * h]));
                    const float r1 = v[c][2 * h + 1] - __bfloat162float(__float2bfloat16(v[c][2 * h + 1]));
                    mid[h] = pack_bf16x2(r0, r1);
                }
                if constexpr (UT == 8) *reinterpret_cast<uint4*>(dst + opBytes) = make_uint4(mid[0], mid[1], mid[2], mid[3]);
                else *reinterpret_cast<uint2*>(dst + opBytes) = make_uint2(mid[0], mid[1]);
            }
        }
        if (zero_pad) {                               // rows >= NC*16 are the K-padding of the dW contraction
            for (int rr = RV + p; rr < R; rr += TP)
                #pragma unroll
                for (int t = 0; t < NTERMS; t++) {
                    uint8_t* dst = buf + t * opBytes + coff + (size_t)rr * 16;
                    if constexpr (UT == 8) *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
                    else *reinterpret_cast<uint2*>(dst) = make_uint2(0, 0);
                }
        }
    };
    // ---- accumulator rows TMEM -> FP32 staging -> thread (p, g) registers ----
    auto stage_to_regs = [&](uint32_t dcol, float (&acc)[NC][UT]) {
        umma::tc_fence_after();
        if (lg * 32 < RV) {
            const int m = lg * 32 + lane;
            const int c0 = wsub * (H / nsub);
            for (int c = c0; c < c0 + H / nsub; c += 16) {
                uint32_t r[16];
                umma::tmem_ld16_nowait(tmem + ((uint32_t)(lg * 32) << 16) + dcol + c, r);
                umma::tmem_ld_wait();
                if (m < RV) {
                    #pragma unroll
                    for (int i = 0; i < 16; i++) Zs[(c + i) * LD + m] = __uint_as_float(r[i]);
                }
            }
        }
        umma::tc_fence_before();
        __syncthreads();
        #pragma unroll
        for (int u = 0; u < UT; u++)
            #pragma unroll
            for (int c = 0; c < NC; c++) acc[c][u] = Zs[(j0 + u) * LD + c * TP + p];
        __syncthreads();                              // staging is free again (operands may be rewritten)
    };
    // ---- MMA issue helpers (thread 0) ----
    auto issue_terms = [&](uint32_t d, uint32_t a0, uint32_t a_lbo, uint32_t a_sbo, uint32_t a_k,
                           uint32_t b0, uint32_t b_lbo, uint32_t b_sbo, uint32_t b_k, uint32_t b_term_bytes,
                           uint32_t idesc, int ksteps, uint32_t acc_first) {
        umma::gemm_issue(d, a0, a_lbo, a_sbo, a_k, b0, b_lbo, b_sbo, b_k, idesc, ksteps, acc_first);
        if constexpr (NTERMS == 2) {
            umma::gemm_issue(d, a0, a_lbo, a_sbo, a_k, b0 + b_term_bytes, b_lbo, b_sbo, b_k, idesc, ksteps, 1u);
            umma::gemm_issue(d, a0 + opBytes, a_lbo, a_sbo, a_k, b0, b_lbo, b_sbo, b_k, idesc, ksteps, 1u);
        }
    };
    auto wait_mma = [&]() { umma::mbar_wait(&mma_bar, mma_phase); mma_phase ^= 1; };

    // ---- deferred dW (see TcArgs) ----
    constexpr uint32_t OPV = (uint32_t)(H / 8) * RV * 16;          // one stored operand: valid rows only
    const bool defer = (!da_separate) && ta.defer_g > 0 && NTERMS == 1 && !VAL;   // VAL tiles have 128 valid rows: stages would not fit
    uint8_t* ring = ta.op_ring + (size_t)blockIdx.x * ta.ring_per_cta;
    int gcount = 0;                                                 // tiles stored since the last catch-up
    uint32_t ld_ph = 0, done_ph = 0;                                // phase bits of ld_bar[s] / done_bar[s] (thread 0)
    auto drain_dw_to = [&](float* dst) {                           // TMEM [fan-in k][unit j] -> partial row, single owner
        umma::tc_fence_after();
        for (int tt = 0; tt < mt; tt++) {
            if (tt * 128 + lg * 32 < H) {
                const int k = tt * 128 + lg * 32 + lane;
                const int c0 = wsub * (H / nsub);
                for (int c = c0; c < c0 + H / nsub; c += 16) {
                    uint32_t r[16];
                    umma::tmem_ld16_nowait(tmem + ((uint32_t)(lg * 32) << 16) + tt * H + c, r);
                    umma::tmem_ld_wait();
                    float4* d4 = reinterpret_cast<float4*>(dst + (size_t)k * H + c);
                    if (k < H) {
                        #pragma unroll
                        for (int q = 0; q < 4; q++)
                            atomicAdd(d4 + q, make_float4(__uint_as_float(r[4 * q + 0]), __uint_as_float(r[4 * q + 1]),
                                                          __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3])));
                    }
                }
            }
        }
        umma::tc_fence_before();
    };
    auto catch_up = [&](int cnt) {
        if (tid == 0) umma::bulk_wait_all();                        // every operand store of these tiles has landed
        __syncthreads();
        for (int l = 2; l <= L; l++) {
            if (tid == 0) {
                umma::tc_fence_after();
                auto load_item = [&](int i) {
                    const int s = i & 1;
                    uint8_t* stg = smem_raw + (size_t)s * 2 * OPV;
                    const uint8_t* src = ring + ((size_t)(i * (L - 1) + (l - 2)) * 2) * OPV;
                    umma::mbar_expect_tx(&ld_bar[s], 2 * OPV);
                    umma::bulk_g2s(stg, src, 2 * OPV, &ld_bar[s]);  // A operand then Zbar operand, contiguous in the ring
                };
                load_item(0);
                for (int i = 0; i < cnt; i++) {
                    const int s = i & 1;
                    if (i + 1 < cnt) {
                        if (i >= 1) { umma::mbar_wait(&done_bar[s ^ 1], (done_ph >> (s ^ 1)) & 1u); done_ph ^= 1u << (s ^ 1); }
                        load_item(i + 1);
                    }
                    umma::mbar_wait(&ld_bar[s], (ld_ph >> s) & 1u); ld_ph ^= 1u << s;
                    const uint32_t sa = umma::smem_u32(smem_raw + (size_t)s * 2 * OPV), sz = sa + OPV;
                    for (int tt = 0; tt < mt; tt++)
                        umma::gemm_issue(tmem + tt * H, sa + tt * 16 * RV * 16, 128, RV * 16, 256, sz, 128, RV * 16, 256,
                                         umma::idesc_bf16(128, H, 1, 1), RV / 16, i > 0 ? 1u : 0u);
                    umma::mma_commit(&done_bar[s]);
                }
                // drain the done barriers so their phases stay in step for the next layer
                if (cnt >= 2) { const int s2 = (cnt - 2) & 1; umma::mbar_wait(&done_bar[s2], (done_ph >> s2) & 1u); done_ph ^= 1u << s2; }
                { const int s1 = (cnt - 1) & 1; umma::mbar_wait(&done_bar[s1], (done_ph >> s1) & 1u); done_ph ^= 1u << s1; }
                umma::mma_commit(&mma_bar);
            }
            wait_mma();
            drain_dw_to(part + nd.offW[l]);
            __syncthreads();
        }
        if (tid == 0) { w_cur = -1; w_pending = false; }            // the stages overwrote the weight buffer
    };

    const int total_tiles = args.sets[0].ntiles + ((VAL && args.nsets > 1) ? args.sets[1].ntiles : 0);
    constexpr int PPT = VAL ? NC * TP : TP;               // points per tile
    float acc[NC][UT];

    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int si = (VAL && t >= args.sets[0].ntiles) ? 1 : 0;
        const PointSet ps = args.sets[si];
        const unsigned long long base = (unsigned long long)(t - (si ? args.sets[0].ntiles : 0)) * PPT;
        const bool valid = base + p < ps.n;               // (VAL: validity is per point group, see vpt())
        auto vpt = [&](int c) { return base + (unsigned long long)(c * TP + p) < ps.n; };
        float x[3] = {0.0f, 0.0f, 0.0f};
        if (!VAL && valid) {
            #pragma unroll
            for (int k = 0; k < DIN; k++) x[k] = __ldg(ps.x + (base + p) * DIN + k);
        }
        auto xval = [&](int c, int k) -> float {          // VAL: coordinate k of this thread's point in group c
            return vpt(c) ? __ldg(ps.x + (base + c * TP + p) * DIN + k) : 0.0f;
        };
        if (tid == 0 && L >= 2) w_request(2, 0);

        // ---------------- layer 1 (K = d_in): CUDA cores ----------------
        if constexpr (VAL) {
            const float* W1 = args.params + nd.offW[1];
            float* st = stash + (size_t)j0 * TP + p;
            float xc[NC][DIN];
            #pragma unroll
            for (int c = 0; c < NC; c++)
                #pragma unroll
                for (int k = 0; k < DIN; k++) xc[c][k] = xval(c, k);
            #pragma unroll
            for (int u = 0; u < UT; u++) {
                const float b1 = __ldg(args.params + nd.offB[1] + j0 + u);
                float w1[DIN];
                #pragma unroll
                for (int k = 0; k < DIN; k++) w1[k] = __ldg(W1 + k * H + j0 + u);
                #pragma unroll
                for (int c = 0; c < NC; c++) {
                    float z0 = b1;
                    #pragma unroll
                    for (int k = 0; k < DIN; k++) z0 = fmaf(xc[c][k], w1[k], z0);
                    const float a = tanhf(z0);
                    st[(size_t)(c * H + u) * TP] = a; acc[c][u] = a;
                }
            }
        } else {
            const float* W1 = args.params + nd.offW[1];
            float* st = stash + (size_t)j0 * TP + p;
            #pragma unroll
            for (int u = 0; u < UT; u++) {
                float z0 = __ldg(args.params + nd.offB[1] + j0 + u);
                float w1[DIN];
                #pragma unroll
                for (int k = 0; k < DIN; k++) { w1[k] = __ldg(W1 + k * H + j0 + u); z0 = fmaf(x[k], w1[k], z0); }
                const float a = tanhf(z0), s = 1.0f - a * a;
                st[(size_t)u * TP] = a; acc[0][u] = a;
                #pragma unroll
                for (int i = 0; i < DIN; i++) {
                    const float zp = w1[i];               // tangent of the input is the unit vector e_i
                    st[(size_t)((1 + i) * H + u) * TP] = zp;
                    acc[1 + i][u] = s * zp;
                    if (i < N2) {
                        st[(size_t)((1 + DIN + i) * H + u) * TP] = 0.0f;
                        acc[1 + DIN + i][u] = -2.0f * a * s * zp * zp;
                    }
                }
            }
        }
        __syncthreads();                                  // previous tile's last staging reads are done
        write_operand(bufA, acc, false);

        // ---------------- hidden layers 2..L: tensor cores ----------------
        for (int l = 2; l <= L; l++) {
            umma::fence_async_smem();
            umma::tc_fence_before();
            __syncthreads();
            for (int h = 0; h < nhalf; h++) {
                if (tid == 0) {
                    umma::tc_fence_after();
                    w_request(l, h); w_ready();
                    issue_terms(tmem + h * HW, a_addr, R * 16, 128, 2 * R * 16, w_addr, 128, H * 16, 256, wBytes,
                                umma::idesc_bf16(128, HW, 0, 1), H / 16, 0u);
                    umma::mma_commit(&mma_bar);
                }
                wait_mma();
                // width 256 runs two halves per layer: without a block barrier here thread 0 could commit the second half
                // before a delayed warp has polled the first — mma_bar would then be TWO phases ahead of that warp, which
                // sees its old parity again and waits forever (rare: observed as a hang only at C5's full 8.4 M points)
                if (nhalf > 1) __syncthreads();
                if (tid == 0) {                           // prefetch the next weights under the epilogue
                    if (h + 1 < nhalf) w_request(l, h + 1);
                    else if (l < L) w_request(l + 1, 0);
                    else if (nhalf > 1) w_request(L, 0);  // first dA contraction of the reverse pass
                }
            }
            stage_to_regs(0u, acc);
            float* st = stash + (size_t)(l - 1) * lstride + (size_t)j0 * TP + p;
            if constexpr (VAL) {
                #pragma unroll
                for (int u = 0; u < UT; u++) {
                    const float bl = __ldg(args.params + nd.offB[l] + j0 + u);
                    #pragma unroll
                    for (int c = 0; c < NC; c++) {
                        const float a = tanhf(acc[c][u] + bl);
                        st[(size_t)(c * H + u) * TP] = a; acc[c][u] = a;
                    }
                }
            } else
            #pragma unroll
            for (int u = 0; u < UT; u++) {
                const float a = tanhf(acc[0][u] + __ldg(args.params + nd.offB[l] + j0 + u));
                const float s = 1.0f - a * a;
                st[(size_t)u * TP] = a; acc[0][u] = a;
                #pragma unroll
                for (int i = 0; i < DIN; i++) {
                    const float zp = acc[1 + i][u];
                    st[(size_t)((1 + i) * H + u) * TP] = zp;
                    acc[1 + i][u] = s * zp;
                    if (i < N2) {
                        const float zpp = acc[1 + DIN + i][u];
                        st[(size_t)((1 + DIN + i) * H + u) * TP] = zpp;
                        acc[1 + DIN + i][u] = s * zpp - 2.0f * a * s * zp * zp;
                    }
                }
            }
            if (l < L) write_operand(bufA, acc, false);
        }

        // ---------------- output layer + residual + seeds: CUDA cores ----------------
        // acc = A_L (FP32).  Partial outputs over this thread's 8 units, summed over groups through smem.
        constexpr int G = H / UT;
        const float* Wo = args.params + nd.offW[L + 1];
        float wo[UT][3];
        #pragma unroll
        for (int u = 0; u < UT; u++)
            #pragma unroll
            for (int o = 0; o < 3; o++) wo[u][o] = (o < nd.d_out) ? __ldg(Wo + (j0 + u) * nd.d_out + o) : 0.0f;
        #pragma unroll
        for (int o = 0; o < 3; o++)
            #pragma unroll
            for (int c = 0; c < NC; c++) {
                float s = 0.0f;
                #pragma unroll
                for (int u = 0; u < UT; u++) s = fmaf(acc[c][u], wo[u][o], s);
                xbuf[((g * 3 + o) * NC + c) * TP + p] = s;
            }
        __syncthreads();
        float* ybuf = xbuf + (size_t)G * 3 * NC * TP;     // [o][c][p] seeds
        if (g == 0) {
            float y[3][NC], ybar[3][NC], r[3] = {0.0f, 0.0f, 0.0f};
            #pragma unroll
            for (int o = 0; o < 3; o++)
                #pragma unroll
                for (int c = 0; c < NC; c++) {
                    float s = 0.0f;
                    for (int gg = 0; gg < G; gg++) s += xbuf[((gg * 3 + o) * NC + c) * TP + p];
                    if ((VAL || c == 0) && o < nd.d_out) s += __ldg(args.params + nd.offB[L + 1] + o);
                    y[o][c] = (o < nd.d_out) ? s : 0.0f;
                }
            float ss = 0.0f;
            if constexpr (VAL) {                          // every "channel" is a point: target misfit per point
                #pragma unroll
                for (int c = 0; c < NC; c++) {
                    float xc[3] = {0.0f, 0.0f, 0.0f}, gt[3] = {0.0f, 0.0f, 0.0f};
                    #pragma unroll
                    for (int k = 0; k < DIN; k++) xc[k] = xval(c, k);
                    pde_target(nd, ps.set, xc, gt);
                    const bool ok = vpt(c);
                    #pragma unroll
                    for (int o = 0; o < 3; o++) {
                        const float e = (ok && o < nd.d_out) ? y[o][c] - gt[o] : 0.0f;
                        ss += e * e;
                        ybar[o][c] = ps.scale * e;
                    }
                }
            } else {
                ss = pde_residual<NC>(nd, x, y, ps.scale, ybar, r);
            }
            if (!VAL && args.mode == 1) {
                if (valid) {
                    for (int o = 0; o < nd.d_out; o++)
                        #pragma unroll
                        for (int c = 0; c < NC; c++) args.y_out[((base + p) * nd.d_out + o) * NC + c] = y[o][c];
                    if (args.r_out)
                        for (int q = 0; q < nd.n_res; q++) args.r_out[(base + p) * nd.n_res + q] = r[q];
                }
            } else {
                if (!VAL && !valid) {
                    ss = 0.0f;
                    #pragma unroll
                    for (int o = 0; o < 3; o++)
                        #pragma unroll
                        for (int c = 0; c < NC; c++) ybar[o][c] = 0.0f;
                }
                #pragma unroll
                for (int o = 0; o < 3; o++) {
                    #pragma unroll
                    for (int c = 0; c < NC; c++) ybuf[(o * NC + c) * TP + p] = ybar[o][c];
                    float bsum = ybar[o][0];              // bias gradient: value channel (VAL: every row is one)
                    if constexpr (VAL) {
                        #pragma unroll
                        for (int c = 1; c < NC; c++) bsum += ybar[o][c];
                    }
                    const float sb = lane_group_sum<TP>(bsum);
                    if (p == 0 && o < nd.d_out) red_add(part + nd.offB[L + 1] + o, sb);
                }
                const float sl = lane_group_sum<TP>(ss);
                if (p == 0) red_add(part + nd.P + ps.set, sl);
            }
        }
        __syncthreads();
        if (args.mode == 1) continue;

        // dW_{L+1}, Abar_L on the CUDA cores
        {
            float yb[3][NC];
            #pragma unroll
            for (int o = 0; o < 3; o++)
                #pragma unroll
                for (int c = 0; c < NC; c++) yb[o][c] = ybuf[(o * NC + c) * TP + p];
            #pragma unroll
            for (int u = 0; u < UT; u++)
                #pragma unroll
                for (int o = 0; o < 3; o++) {
                    float s = 0.0f;
                    #pragma unroll
                    for (int c = 0; c < NC; c++) s = fmaf(acc[c][u], yb[o][c], s);
                    s = lane_group_sum<TP>(s);
                    if (p == 0 && o < nd.d_out) red_add(part + nd.offW[L + 1] + (j0 + u) * nd.d_out + o, s);
                }
            #pragma unroll
            for (int u = 0; u < UT; u++)
                #pragma unroll
                for (int c = 0; c < NC; c++) {
                    float s = 0.0f;
                    #pragma unroll
                    for (int o = 0; o < 3; o++) s = fmaf(yb[o][c], wo[u][o], s);
                    acc[c][u] = s;
                }
        }
        __syncthreads();                                  // xbuf / ybuf (alias bufA | bufZ) are free

        // ---------------- reverse pass ----------------
        for (int l = L; l >= 1; l--) {
            // adjoint of the tanh layer, in place: acc = Abar_l -> Zbar_l
            const float* st = stash + (size_t)(l - 1) * lstride + (size_t)j0 * TP + p;
            if constexpr (VAL) {
                #pragma unroll
                for (int u = 0; u < UT; u++) {
                    float bsum = 0.0f;
                    #pragma unroll
                    for (int c = 0; c < NC; c++) {
                        const float a = st[(size_t)(c * H + u) * TP];
                        const float zb = (1.0f - a * a) * acc[c][u];
                        acc[c][u] = zb; bsum += zb;
                    }
                    const float sb = lane_group_sum<TP>(bsum);
                    if (p == 0) red_add(part + nd.offB[l] + j0 + u, sb);
                }
            } else
            #pragma unroll
            for (int u = 0; u < UT; u++) {
                const float a = st[(size_t)u * TP];
                const float s = 1.0f - a * a;
                float zb0 = s * acc[0][u];
                #pragma unroll
                for (int i = 0; i < DIN; i++) {
                    const float zp = st[(size_t)((1 + i) * H + u) * TP];
                    const float abp = acc[1 + i][u];
                    float zbp = s * abp;
                    zb0 += (-2.0f * a * s * zp) * abp;
                    if (i < N2) {
                        const float zpp = st[(size_t)((1 + DIN + i) * H + u) * TP];
                        const float abpp = acc[1 + DIN + i][u];
                        acc[1 + DIN + i][u] = s * abpp;
                        zbp -= 4.0f * a * s * zp * abpp;
                        zb0 += (-2.0f * a * s * zpp - 2.0f * s * (1.0f - 3.0f * a * a) * zp * zp) * abpp;
                    }
                    acc[1 + i][u] = zbp;
                }
                acc[0][u] = zb0;
                const float sb = lane_group_sum<TP>(zb0);
                if (p == 0) red_add(part + nd.offB[l] + j0 + u, sb);
            }
            if (l == 1) {
                // dW_1[k][j] = sum_p ( x_k * zbar_0 + zbar_{1+k} )   (inputs: value x, unit tangents)
                if constexpr (VAL) {
                    #pragma unroll
                    for (int k = 0; k < DIN; k++) {
                        float xk[NC];
                        #pragma unroll
                        for (int c = 0; c < NC; c++) xk[c] = xval(c, k);
                        #pragma unroll
                        for (int u = 0; u < UT; u++) {
                            float sx = 0.0f;
                            #pragma unroll
                            for (int c = 0; c < NC; c++) sx = fmaf(xk[c], acc[c][u], sx);
                            const float s = lane_group_sum<TP>(sx);
                            if (p == 0) red_add(part + nd.offW[1] + k * H + j0 + u, s);
                        }
                    }
                } else {
                    #pragma unroll
                    for (int u = 0; u < UT; u++)
                        #pragma unroll
                        for (int k = 0; k < DIN; k++) {
                            const float s = lane_group_sum<TP>(fmaf(x[k], acc[0][u], acc[1 + k][u]));
                            if (p == 0) red_add(part + nd.offW[1] + k * H + j0 + u, s);
                        }
                }
                break;
            }
            write_operand(bufZ, acc, !defer);
            // A_{l-1} (dW operand) recomputed from the stash
            {
                const float* sp = stash + (size_t)(l - 2) * lstride + (size_t)j0 * TP + p;
                if constexpr (VAL) {
                    #pragma unroll
                    for (int u = 0; u < UT; u++)
                        #pragma unroll
                        for (int c = 0; c < NC; c++) acc[c][u] = sp[(size_t)(c * H + u) * TP];
                } else
                #pragma unroll
                for (int u = 0; u < UT; u++) {
                    const float a = sp[(size_t)u * TP];
                    const float s = 1.0f - a * a;
                    acc[0][u] = a;
                    #pragma unroll
                    for (int i = 0; i < DIN; i++) {
                        const float zp = sp[(size_t)((1 + i) * H + u) * TP];
                        acc[1 + i][u] = s * zp;
                        if (i < N2) {
                            const float zpp = sp[(size_t)((1 + DIN + i) * H + u) * TP];
                            acc[1 + DIN + i][u] = s * zpp - 2.0f * a * s * zp * zp;
                        }
                    }
                }
            }
            write_operand(bufA, acc, !defer);
            umma::fence_async_smem();
            umma::tc_fence_before();
            __syncthreads();

            auto issue_dw = [&]() {
                for (int tt = 0; tt < mt; tt++)
                    issue_terms(tmem + tt * H, a_addr + tt * 16 * R * 16, 128, R * 16, 256, z_addr, 128, R * 16, 256, opBytes,
                                umma::idesc_bf16(128, H, 1, 1), R / 16, 0u);
            };
            auto issue_da = [&](int h) {
                issue_terms(tmem + da_off, z_addr + h * (HW / 8) * R * 16, R * 16, 128, 2 * R * 16, w_addr, H * 16, 128, 2 * H * 16,
                            wBytes, umma::idesc_bf16(128, H, 0, 0), HW / 16, h > 0 ? 1u : 0u);
            };
            auto drain_dw = [&]() { drain_dw_to(part + nd.offW[l]); };

            if (da_separate) {
                if (tid == 0) {
                    umma::tc_fence_after();
                    issue_dw();
                    w_request(l, 0); w_ready();
                    issue_da(0);
                    umma::mma_commit(&mma_bar);
                }
                wait_mma();
                if (tid == 0 && l > 2) w_request(l - 1, 0);
                drain_dw();
            } else {
                if (defer) {
                    // dW of this tile-layer is deferred: its two operands (valid rows only) go to the ring by TMA bulk store
                    if (tid == 0) {
                        uint8_t* slot = ring + ((size_t)(gcount * (L - 1) + (l - 2)) * 2) * OPV;
                        for (int h8 = 0; h8 < H / 8; h8++) {
                            umma::bulk_s2g(slot + (size_t)h8 * RV * 16, bufA + (size_t)h8 * R * 16, RV * 16);
                            umma::bulk_s2g(slot + OPV + (size_t)h8 * RV * 16, bufZ + (size_t)h8 * R * 16, RV * 16);
                        }
                        umma::bulk_commit();
                    }
                } else {
                    if (tid == 0) { umma::tc_fence_after(); issue_dw(); umma::mma_commit(&mma_bar); }
                    wait_mma();
                    drain_dw();
                    __syncthreads();
                }
                for (int h = 0; h < nhalf; h++) {
                    if (tid == 0) {
                        umma::tc_fence_after();
                        w_request(l, h); w_ready();
                        issue_da(h);
                        umma::mma_commit(&mma_bar);
                    }
                    wait_mma();
                    if (nhalf > 1) __syncthreads();       // (same two-commit hazard as in the forward loop)
                    if (tid == 0) {
                        if (h + 1 < nhalf) w_request(l, h + 1);
                        else if (l > 2) w_request(l - 1, 0);
                    }
                }
            }
            if (defer) {                                  // the staging area aliases the operands the bulk stores read
                if (tid == 0) umma::bulk_wait_read_all();
                __syncthreads();
            }
            stage_to_regs(da_off, acc);                   // acc = Abar_{l-1}
        }
        if (defer && args.mode == 0) {
            if (++gcount == ta.defer_g) { catch_up(gcount); gcount = 0; }
        }
    }

    if (defer && gcount > 0) catch_up(gcount);
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, (uint32_t)ta.tmem_cols);
}

// BF16 chunk-major mirror(s) of the hidden->hidden weights (hi term, and hi + mid when nterms = 2).
__global__ void pack_wq_kernel(const NetDims nd, const float* __restrict__ params, __nv_bfloat16* __restrict__ wq,
                               long long term_stride, int nterms) {
    const int H = nd.H;
    const long long n = (long long)(nd.L - 1) * H * H;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int l = (int)(i / ((long long)H * H)) + 2;
    const int r = (int)(i % ((long long)H * H)), k = r / H, j = r - k * H;
    const float w = params[nd.offW[l] + r];
    const __nv_bfloat16 hi = __float2bfloat16(w);
    const size_t dst = (size_t)(l - 2) * H * H + (size_t)(j >> 3) * (H * 8) + (size_t)k * 8 + (j & 7);
    wq[dst] = hi;
    if (nterms == 2) wq[term_stride + dst] = __float2bfloat16(w - __bfloat162float(hi));
}

}  // namespace pinn
