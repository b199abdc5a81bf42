// pinn_fused_mma.cuh — register-chained tensor-core train step for NARROW nets (width 20, the headline
// Burgers network 2 -> [20 x L] -> 1), PINN_PREC_BF16X3 (two BF16 terms per operand, 3 MMAs per product) and
// PINN_PREC_BF16X6 (three terms = an exact 24-bit split of the FP32 value, 6 MMAs: FP32-equivalent products).
//
// Why a third kernel: at width 20 the FP32 block kernel is bound by shared-memory operand traffic (5 LDS per
// 16 FFMA, profiles/r01_c2_fp32_block_kernel_ncu.txt) and the tcgen05 kernel cannot pay for its TMEM / shared
// round trips with 20x20 matrices.  The warp-level mma.sync path (SASS HMMA; 550 TFLOP/s BF16 measured on
// B200, tests/cuda/mma_sync_peak.cu) has the one property that matters here: the m16n8 FP32 accumulator
// fragment of one layer has exactly the register layout of the m16k16 A-operand fragment of the next, so the
// activations of a 16-point tile travel through the whole network IN REGISTERS:
//
//   warp  = 16 collocation points; Taylor channel c = one m16 row tile (row = point, NC tiles per warp)
//   lane  = (g = lane/4, t = lane%4) holds, for n-tile nt, points g and g+8, units 8nt+2t and 8nt+2t+1
//           -> the tanh / Taylor elementwise work (all channels of one (point, unit)) is thread-local
//   Z = A W       A-fragments = previous activations split into BF16 terms (hi, mid[, lo]), B-fragments =
//                 pre-split weights read from a fragment-ordered mirror (global, L1-resident); the cross terms
//                 of total order <= NTERMS-1 are issued smallest first (FP32 accumulate); channels are
//                 independent in the contraction, so each is split, multiplied and overwritten in place
//   dA = Zbar W^T same chain with the transposed weight fragments
//   dW = A^T Zbar contraction over the points: both operands go through a 12 (18) KB per-warp shared buffer and
//                 come back transposed by ldmatrix.trans (the only shared-memory traffic of the kernel)
//   K = 20 is covered by one k16 and one k8 MMA (no padding to 32); N = 20 by three n8 tiles.
//
// No block-level synchronisation exists: warps are independent (only __syncwarp around the dW buffer).
// Gradients go to the WARP's private partial row (plain stores for its first tile, fire-and-forget RED after; single
// owner per element -> bit-reproducible), the stash (a, z', z'' per layer) lives in the warp's private ring in fragment order
// (float4 per lane: perfectly coalesced, L2-resident).  Math: SURVEY.md Appendix B, same as the FP32 kernel.
#pragma once
#include <type_traits>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "pinn_device.cuh"
#include "pinn_fused_fp32.cuh"

namespace pinn {

struct MmaArgs {
    const float* params;        // flat [P] (biases, first and last layer are read in FP32)
    const uint32_t* wfrag;      // [(L-1)][dir 2][9*NTERMS][32 lanes]: pre-split BF16 B-fragments of W_2..W_L (pack_wfrag_kernel)
    PointSet set;               // the interior set (ntiles counts 16-point tiles)
    PointSet vsets[2];          // boundary / initial sets: value-only 16-point tiles of the same launch
    int nvsets;
    float* partial;             // [rows][PP], one row per WARP
    float* stash;               // [rows][stash_per_row]
    int row0;
    long long stash_per_row;    // floats
    float* x_copy;              // non-null: set.x is mapped HOST memory; interior tiles also store their points here (HBM)
};

constexpr int kMmaWarps = 4;                      // warps per CTA

__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma1688(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(b0));
}
// (x0, x1) -> NTERMS packed BF16 pairs: f[0] = rn(x), f[1] = rn(x - f[0]), f[2] = rn(x - f[0] - f[1]);
// x0 goes to the low half (lower column / k index).  Three terms reproduce the 24-bit significand exactly.
template <int NTERMS>
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t (&f)[NTERMS]) {
    #pragma unroll
    for (int k = 0; k < NTERMS; k++) {
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(f[k]) : "f"(x1), "f"(x0));
        if (k + 1 < NTERMS) { x0 -= __uint_as_float(f[k] << 16); x1 -= __uint_as_float(f[k] & 0xffff0000u); }
    }
}
// FP16 flavour (PINN_PREC_FP16X3): two FP16 terms carry 22 significand bits, so three MMAs give near-FP32 products at the
// cost of the two-term BF16 split.  FP16 has a 5-bit exponent, so every operand fragment is first multiplied by an exact
// power of two chosen per warp and channel from the fragment's largest magnitude (results are scaled back exactly).
constexpr int kF16WeightShift = 4;                // the weight mirror holds W * 2^4
constexpr int kF16TargetExp = 13;                 // scaled fragment maximum lies in [2^13, 2^14)
__device__ __forceinline__ void mma16816_f16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma1688_f16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ void split_pair_f16(float x0, float x1, uint32_t (&f)[2]) {
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(f[0]) : "f"(x1), "f"(x0));
    const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&f[0]));
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(f[1]) : "f"(x1 - h.y), "f"(x0 - h.x));
}
// power-of-two shift k (scale 2^k) that brings a fragment whose largest magnitude is `m` to [2^13, 2^14); clamped to
// [-17, 43] (magnitudes in [2^-30, 2^30] keep full precision) so that sums of two shifts stay representable
__device__ __forceinline__ int f16_shift(float m) {
    int e = (int)((__float_as_uint(m) >> 23) & 0xffu);
    e = e < 97 ? 97 : (e > 157 ? 157 : e);
    return kF16TargetExp + 127 - e;
}
__device__ __forceinline__ float pow2i(int k) { return __uint_as_float((uint32_t)(127 + k) << 23); }
__device__ __forceinline__ float warp_max(float m) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    return m;
}

// cross terms (A term, B term) kept by the NTERMS-term product, smallest first
template <int NTERMS> struct TermPairs {
    static constexpr int n = NTERMS == 3 ? 6 : 3;
    // NTERMS = 2: (mid,hi) (hi,mid) (hi,hi);  NTERMS = 3: (mid,mid) (hi,lo) (lo,hi) (hi,mid) (mid,hi) (hi,hi)
    __host__ __device__ static constexpr int a(int q) { return NTERMS == 3 ? (q == 0 ? 1 : q == 2 ? 2 : q == 4 ? 1 : 0) : (q == 0 ? 1 : 0); }
    __host__ __device__ static constexpr int b(int q) { return NTERMS == 3 ? (q == 0 ? 1 : q == 1 ? 2 : q == 3 ? 1 : 0) : (q == 1 ? 1 : 0); }
};

__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr) : "memory");
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t (&r)[2], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(saddr) : "memory");
}
// sum over the 8 lanes that share t (all g): the column sums of a fragment
__device__ __forceinline__ float sum_over_g(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    return v;
}

// Fragment-ordered, pre-split BF16 mirror of the hidden->hidden weights.  Word (li, dir, idx, lane) with
// idx = (nt*nterms + term)*3 + r holds B[k0][n], B[k0+1][n] (low, high half) for n = 8nt + g and
// k0 = 2t (r=0, k16 b0) | 2t+8 (r=1, k16 b1) | 16+2t (r=2, k8 b0);  dir 0: B[k][n] = W[k][n] (forward),
// dir 1: B[k][n] = W[n][k] (dA = Zbar W^T).  Entries outside the H x H matrix are zero.
__global__ void pack_wfrag_kernel(const NetDims nd, const float* __restrict__ params, uint32_t* __restrict__ wfrag, int nterms, int f16) {
    const int words = 9 * nterms;
    const int total = (nd.L - 1) * 2 * words * 32;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= total) return;
    const int lane = w & 31, idx = (w >> 5) % words, dir = ((w >> 5) / words) & 1, li = (w >> 5) / (2 * words);
    const int g = lane >> 2, t = lane & 3, nt = idx / (3 * nterms), term = (idx / 3) % nterms, r = idx % 3;
    const int H = nd.H, n = 8 * nt + g, k0 = (r == 0) ? 2 * t : (r == 1 ? 2 * t + 8 : 16 + 2 * t);
    const float* W = params + nd.offW[li + 2];
    float x[2];
    #pragma unroll
    for (int e = 0; e < 2; e++) {
        const int k = k0 + e;
        x[e] = (k < H && n < H) ? (dir == 0 ? W[k * H + n] : W[n * H + k]) : 0.0f;
    }
    uint32_t f[3];
    if (f16) { uint32_t h[2]; split_pair_f16(x[0] * (float)(1 << kF16WeightShift), x[1] * (float)(1 << kF16WeightShift), h); f[0] = h[0]; f[1] = h[1]; f[2] = 0u; }
    else split_pair<3>(x[0], x[1], f);
    wfrag[w] = f[term];
}

// One 16-point tile on one warp.  VAL = false: interior tile, all NC Taylor channels, PDE residual.
// VAL = true: boundary / initial tile, value channel only, Dirichlet / initial target (same fragment schedule with NC = 1).
template <int DIN, int N2, int HT, int NTERMS, bool VAL, bool F16>
__device__ __forceinline__ void mma_tile(const NetDims& nd, const MmaArgs& args, const PointSet& ps, int tile, uint32_t* sbuf,
                                         float* part, float4* stash, bool first) {
    constexpr int ND = VAL ? 0 : DIN, NS = VAL ? 0 : N2;   // tangent channels carried
    constexpr int NC = 1 + ND + NS, NT = 3;
    constexpr int NCS = 1 + DIN + N2;                     // channel stride of the shared buffer (sized for interior tiles)
    constexpr int WORDS = 9 * NTERMS;                     // fragment words per (layer, direction)
    using TP = TermPairs<NTERMS>;
    static_assert(!F16 || NTERMS == 2, "the FP16 flavour is a two-term split");
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int L = nd.L;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sbuf);
    // gradient output: the warp's FIRST tile stores (its row needs no zeroing), later tiles add by fire-and-forget RED
    auto emit = [&](float* p, float val) { if (first) *p = val; else red_add(p, val); };
    auto mma_k16 = [](float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
        if constexpr (F16) mma16816_f16(d, a0, a1, a2, a3, b0, b1); else mma16816(d, a0, a1, a2, a3, b0, b1);
    };
    auto mma_k8 = [](float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
        if constexpr (F16) mma1688_f16(d, a0, a1, b0); else mma1688(d, a0, a1, b0);
    };
    auto split = [](float x0, float x1, uint32_t (&f)[NTERMS]) {
        if constexpr (F16) split_pair_f16(x0, x1, f); else split_pair<NTERMS>(x0, x1, f);
    };
    // FP16 flavour: power-of-two shift of the last chain_gemm's input fragment per channel (dW needs the Zbar ones), and
    // the per-layer word of forward shifts kept in the stash (byte c = shift of A_{l-1,c} + 64)
    int kin[NC];
    uint32_t* stash_exp = reinterpret_cast<uint32_t*>(stash) + (size_t)nd.L * NCS * NT * 128;

    // unit index of fragment slot (nt, e) and its validity
    int ju[NT][2]; bool jv[NT][2];
    #pragma unroll
    for (int nt = 0; nt < NT; nt++)
        #pragma unroll
        for (int e = 0; e < 2; e++) { ju[nt][e] = 8 * nt + 2 * t + e; jv[nt][e] = ju[nt][e] < HT; }

    float v[NC][NT][4];                                   // the travelling fragment: Z -> A -> ... -> Abar -> Zbar

    // bias + tanh-Taylor of layer l in place (v: pre-activations without bias -> activations), stash (a, z', z'')
    auto activate = [&](int l) {
        #pragma unroll
        for (int nt = 0; nt < NT; nt++) {
            float sv[NC][4];
            #pragma unroll
            for (int q = 0; q < 4; q++) {
                const int e = q & 1;
                const float b = jv[nt][e] ? __ldg(args.params + nd.offB[l] + ju[nt][e]) : 0.0f;
                const float a = tanhf(v[0][nt][q] + b);
                const float s = 1.0f - a * a;
                sv[0][q] = a;
                v[0][nt][q] = a;
                #pragma unroll
                for (int i = 0; i < ND; i++) {
                    const float zp = v[1 + i][nt][q];
                    sv[1 + i][q] = zp;
                    v[1 + i][nt][q] = s * zp;
                    if (i < NS) {
                        const float zpp = v[1 + ND + i][nt][q];
                        sv[1 + ND + i][q] = zpp;
                        v[1 + ND + i][nt][q] = s * zpp - 2.0f * a * s * zp * zp;
                    }
                }
            }
            #pragma unroll
            for (int c = 0; c < NC; c++)
                stash[((size_t)(l - 1) * NC * NT + c * NT + nt) * 32 + lane] = make_float4(sv[c][0], sv[c][1], sv[c][2], sv[c][3]);
        }
    };

    // v_c <- v_c x B for every channel c, in place; B = fragment mirror of (layer l, dir).
    // STORE: also leave the BF16 terms of the INPUT fragments in the warp's shared buffer 1 ([c][term][point][24]).
    auto chain_gemm = [&](int l, int dir, auto store_tag) {
        constexpr bool STORE = decltype(store_tag)::value;
        const uint32_t* wf = args.wfrag + ((size_t)((l - 2) * 2 + dir) * WORDS) * 32 + lane;
        uint32_t b[NT][NTERMS][3];
        #pragma unroll
        for (int nt = 0; nt < NT; nt++)
            #pragma unroll
            for (int k = 0; k < NTERMS; k++)
                #pragma unroll
                for (int r = 0; r < 3; r++) b[nt][k][r] = __ldg(wf + ((nt * NTERMS + k) * 3 + r) * 32);
        #pragma unroll
        for (int c = 0; c < NC; c++) {
            float sc = 1.0f;
            if constexpr (F16) {                            // exact power-of-two scaling of this channel's fragment
                float mx = 0.0f;
                #pragma unroll
                for (int nt = 0; nt < NT; nt++)
                    #pragma unroll
                    for (int q = 0; q < 4; q++) mx = fmaxf(mx, fabsf(v[c][nt][q]));
                kin[c] = f16_shift(warp_max(mx));
                sc = pow2i(kin[c]);
            }
            uint32_t f[6][NTERMS];                          // slot nt*2+h = (rows g+8h, columns of n-tile nt)
            #pragma unroll
            for (int nt = 0; nt < NT; nt++)
                #pragma unroll
                for (int h = 0; h < 2; h++) {
                    split(v[c][nt][2 * h] * sc, v[c][nt][2 * h + 1] * sc, f[nt * 2 + h]);
                    if constexpr (STORE) {
                        const int word = (g + 8 * h) * 12 + 4 * nt + t;
                        #pragma unroll
                        for (int k = 0; k < NTERMS; k++) sbuf[((1 * NCS + c) * NTERMS + k) * 192 + word] = f[nt * 2 + h][k];
                    }
                }
            #pragma unroll
            for (int nt = 0; nt < NT; nt++) v[c][nt][0] = v[c][nt][1] = v[c][nt][2] = v[c][nt][3] = 0.0f;
            // units 16..19 (an m16n8k8 costs a full tensor slot): the remainders of TWO cross terms share one k16 MMA
            #pragma unroll
            for (int q = 0; q + 1 < TP::n; q += 2)
                #pragma unroll
                for (int nt = 0; nt < NT; nt++)
                    mma_k16(v[c][nt], f[4][TP::a(q)], f[5][TP::a(q)], f[4][TP::a(q + 1)], f[5][TP::a(q + 1)],
                            b[nt][TP::b(q)][2], b[nt][TP::b(q + 1)][2]);
            if constexpr (TP::n & 1) {
                #pragma unroll
                for (int nt = 0; nt < NT; nt++) mma_k8(v[c][nt], f[4][TP::a(TP::n - 1)], f[5][TP::a(TP::n - 1)], b[nt][TP::b(TP::n - 1)][2]);
            }
            #pragma unroll
            for (int q = 0; q < TP::n; q++)                 // units 0..15; the three n-tiles are independent accumulators
                #pragma unroll
                for (int nt = 0; nt < NT; nt++)
                    mma_k16(v[c][nt], f[0][TP::a(q)], f[1][TP::a(q)], f[2][TP::a(q)], f[3][TP::a(q)], b[nt][TP::b(q)][0], b[nt][TP::b(q)][1]);
            if constexpr (F16) {
                const float us = pow2i(-(kin[c] + kF16WeightShift));
                #pragma unroll
                for (int nt = 0; nt < NT; nt++)
                    #pragma unroll
                    for (int q = 0; q < 4; q++) v[c][nt][q] *= us;
            }
        }
        if constexpr (F16 && !STORE) {                      // forward: the shifts of A_{l-1}, needed again for dW_l
            static_assert(NC <= 4, "one byte per channel");
            uint32_t w = 0;
            #pragma unroll
            for (int c = 0; c < NC; c++) w |= (uint32_t)(kin[c] + 64) << (8 * c);
            stash_exp[(size_t)(l - 1) * 32 + lane] = w;
        }
    };

    const unsigned long long base = (unsigned long long)tile * 16;
    float x[2][3]; bool valid[2];
    #pragma unroll
    for (int h = 0; h < 2; h++) {
        const unsigned long long p = base + g + 8 * h;
        valid[h] = p < ps.n;
        #pragma unroll
        for (int k = 0; k < 3; k++) x[h][k] = (k < DIN && valid[h]) ? __ldg(ps.x + p * DIN + k) : 0.0f;
        if constexpr (!VAL) {
            if (args.x_copy != nullptr && t == 0 && valid[h]) {
                #pragma unroll
                for (int k = 0; k < DIN; k++) args.x_copy[p * DIN + k] = x[h][k];
            }
        }
    }

    // ------------------------------ forward ------------------------------
    // layer 1 (fan-in DIN) in FP32: z = x W_1, z'_i = W_1[i], z''_i = 0
    #pragma unroll
    for (int nt = 0; nt < NT; nt++)
        #pragma unroll
        for (int q = 0; q < 4; q++) {
            const int e = q & 1, h = q >> 1;
            float w[DIN];
            #pragma unroll
            for (int k = 0; k < DIN; k++) w[k] = jv[nt][e] ? __ldg(args.params + nd.offW[1] + k * HT + ju[nt][e]) : 0.0f;
            float z = 0.0f;
            #pragma unroll
            for (int k = 0; k < DIN; k++) z = fmaf(x[h][k], w[k], z);
            v[0][nt][q] = z;
            #pragma unroll
            for (int i = 0; i < ND; i++) {
                v[1 + i][nt][q] = w[i];
                if (i < NS) v[1 + ND + i][nt][q] = 0.0f;
            }
        }
    activate(1);
    for (int l = 2; l <= L; l++) {
        chain_gemm(l, 0, std::false_type{});
        activate(l);
    }

    // ---- linear output layer (d_out = 1), residual / target, loss seeds ----
    float wo[NT][2];
    #pragma unroll
    for (int nt = 0; nt < NT; nt++)
        #pragma unroll
        for (int e = 0; e < 2; e++) wo[nt][e] = jv[nt][e] ? __ldg(args.params + nd.offW[L + 1] + ju[nt][e]) : 0.0f;
    float ybar[2][NC];
    float ss_sum = 0.0f, yb0_sum = 0.0f;
    #pragma unroll
    for (int h = 0; h < 2; h++) {
        float y[3][NC], yb[3][NC];
        #pragma unroll
        for (int c = 0; c < NC; c++) {
            float s = 0.0f;
            #pragma unroll
            for (int nt = 0; nt < NT; nt++)
                #pragma unroll
                for (int e = 0; e < 2; e++) s = fmaf(v[c][nt][2 * h + e], wo[nt][e], s);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            y[0][c] = (c == 0) ? s + __ldg(args.params + nd.offB[L + 1]) : s;
            y[1][c] = 0.0f; y[2][c] = 0.0f;
        }
        float ss;
        if constexpr (VAL) {
            float gt[3] = {0.0f, 0.0f, 0.0f};
            pde_target(nd, ps.set, x[h], gt);
            const float e = y[0][0] - gt[0];
            ss = e * e;
            yb[0][0] = ps.scale * e;
        } else {
            float r[3] = {0.0f, 0.0f, 0.0f};
            ss = pde_residual<NC>(nd, x[h], y, ps.scale, yb, r);
        }
        #pragma unroll
        for (int c = 0; c < NC; c++) ybar[h][c] = valid[h] ? yb[0][c] : 0.0f;
        if (!valid[h]) ss = 0.0f;
        if (t == 0) { ss_sum += ss; yb0_sum += ybar[h][0]; }
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ss_sum += __shfl_xor_sync(0xffffffffu, ss_sum, o);
        yb0_sum += __shfl_xor_sync(0xffffffffu, yb0_sum, o);
    }
    if (lane == 0) {
        if (first) {                                        // the first tile initialises all three loss slots of the row
            #pragma unroll
            for (int q = 0; q < 3; q++) part[nd.P + q] = (q == ps.set) ? ss_sum : 0.0f;
        } else {
            red_add(part + nd.P + ps.set, ss_sum);
        }
        emit(part + nd.offB[L + 1], yb0_sum);
    }
    // dW_out[j] = sum_p sum_c a_c[p][j] ybar_c[p];  Abar_c[p][j] = ybar_c[p] W_out[j]
    #pragma unroll
    for (int nt = 0; nt < NT; nt++)
        #pragma unroll
        for (int e = 0; e < 2; e++) {
            float s = 0.0f;
            #pragma unroll
            for (int h = 0; h < 2; h++)
                #pragma unroll
                for (int c = 0; c < NC; c++) s = fmaf(v[c][nt][2 * h + e], ybar[h][c], s);
            s = sum_over_g(s);
            if (g == 0 && jv[nt][e]) emit(part + nd.offW[L + 1] + ju[nt][e], s);
            #pragma unroll
            for (int h = 0; h < 2; h++)
                #pragma unroll
                for (int c = 0; c < NC; c++) v[c][nt][2 * h + e] = ybar[h][c] * wo[nt][e];
        }

    // ------------------------------ reverse ------------------------------
    for (int l = L; l >= 1; l--) {
        // elementwise adjoint of the tanh layer: v = Abar_l -> Zbar_l
        #pragma unroll
        for (int nt = 0; nt < NT; nt++) {
            float sv[NC][4];
            #pragma unroll
            for (int c = 0; c < NC; c++) {
                const float4 t4 = stash[((size_t)(l - 1) * NC * NT + c * NT + nt) * 32 + lane];
                sv[c][0] = t4.x; sv[c][1] = t4.y; sv[c][2] = t4.z; sv[c][3] = t4.w;
            }
            #pragma unroll
            for (int q = 0; q < 4; q++) {
                const float a = sv[0][q];
                const float s = 1.0f - a * a;
                float zb0 = s * v[0][nt][q];
                #pragma unroll
                for (int i = 0; i < ND; i++) {
                    const float zp = sv[1 + i][q];
                    const float abp = v[1 + i][nt][q];
                    float zbp = s * abp;
                    zb0 += (-2.0f * a * s * zp) * abp;
                    if (i < NS) {
                        const float zpp = sv[1 + ND + i][q];
                        const float abpp = v[1 + ND + i][nt][q];
                        v[1 + ND + i][nt][q] = s * abpp;
                        zbp -= 4.0f * a * s * zp * abpp;
                        zb0 += (-2.0f * a * s * zpp - 2.0f * s * (1.0f - 3.0f * a * a) * zp * zp) * abpp;
                    }
                    v[1 + i][nt][q] = zbp;
                }
                v[0][nt][q] = zb0;
            }
            // bias gradient: column sums of Zbar_0
            #pragma unroll
            for (int e = 0; e < 2; e++) {
                const float s = sum_over_g(v[0][nt][e] + v[0][nt][2 + e]);
                if (g == 0 && jv[nt][e]) emit(part + nd.offB[l] + ju[nt][e], s);
            }
        }
        if (l == 1) {
            // dW_1[k][j] = sum_p x_k Zbar_0 + Zbar_{1+k}   (A_0: value = coordinates, tangents = unit vectors)
            #pragma unroll
            for (int nt = 0; nt < NT; nt++)
                #pragma unroll
                for (int e = 0; e < 2; e++)
                    #pragma unroll
                    for (int k = 0; k < DIN; k++) {
                        float s = 0.0f;
                        #pragma unroll
                        for (int h = 0; h < 2; h++) {
                            if constexpr (VAL) s = fmaf(x[h][k], v[0][nt][2 * h + e], s);
                            else s += fmaf(x[h][k], v[0][nt][2 * h + e], v[1 + k][nt][2 * h + e]);
                        }
                        s = sum_over_g(s);
                        if (g == 0 && jv[nt][e]) emit(part + nd.offW[1] + k * HT + ju[nt][e], s);
                    }
            break;
        }
        chain_gemm(l, 1, std::true_type{});                // Zbar_l terms -> shared buffer 1;  v = Abar_{l-1} = Zbar_l W_l^T
        // FP16 flavour: give A_{l-1,c} the shift K - shift(Zbar_c), K = min_c(own best shift + shift(Zbar_c)): every channel's
        // products then carry 2^K and share one dW accumulator; the channel that dominates dW keeps full precision
        float sa[NC];
        int K = 0;
        if constexpr (F16) {
            const uint32_t w = stash_exp[(size_t)(l - 1) * 32 + lane];
            K = 1 << 20;
            #pragma unroll
            for (int c = 0; c < NC; c++) { const int kopt = (int)((w >> (8 * c)) & 0xffu) - 64; K = min(K, kopt + kin[c]); }
            #pragma unroll
            for (int c = 0; c < NC; c++) sa[c] = pow2i(K - kin[c]);
        } else {
            #pragma unroll
            for (int c = 0; c < NC; c++) sa[c] = 1.0f;
        }
        // A_{l-1} (the other dW operand) recomputed from the stash of layer l-1, straight into shared buffer 0
        #pragma unroll
        for (int nt = 0; nt < NT; nt++) {
            float sv[NC][4];
            #pragma unroll
            for (int c = 0; c < NC; c++) {
                const float4 t4 = stash[((size_t)(l - 2) * NC * NT + c * NT + nt) * 32 + lane];
                sv[c][0] = t4.x; sv[c][1] = t4.y; sv[c][2] = t4.z; sv[c][3] = t4.w;
            }
            float ac[NC][4];
            #pragma unroll
            for (int q = 0; q < 4; q++) {
                const float a = sv[0][q], s = 1.0f - a * a;
                ac[0][q] = a;
                #pragma unroll
                for (int i = 0; i < ND; i++) {
                    const float zp = sv[1 + i][q];
                    ac[1 + i][q] = s * zp;
                    if (i < NS) ac[1 + ND + i][q] = s * sv[1 + ND + i][q] - 2.0f * a * s * zp * zp;
                }
            }
            #pragma unroll
            for (int c = 0; c < NC; c++)
                #pragma unroll
                for (int h = 0; h < 2; h++) {
                    uint32_t f[NTERMS];
                    if constexpr (F16) split(ac[c][2 * h] * sa[c], ac[c][2 * h + 1] * sa[c], f);
                    else split(ac[c][2 * h], ac[c][2 * h + 1], f);
                    const int word = (g + 8 * h) * 12 + 4 * nt + t;
                    #pragma unroll
                    for (int k = 0; k < NTERMS; k++) sbuf[((0 * NCS + c) * NTERMS + k) * 192 + word] = f[k];
                }
        }
        __syncwarp();
        // dW_l[i][j] = sum_c sum_p A_{l-1,c}[p][i] Zbar_c[p][j]: M = i (two m16 tiles), N = j, K = the 16 points
        float dw[2][NT][4];
        #pragma unroll
        for (int mt = 0; mt < 2; mt++)
            #pragma unroll
            for (int nt = 0; nt < NT; nt++) dw[mt][nt][0] = dw[mt][nt][1] = dw[mt][nt][2] = dw[mt][nt][3] = 0.0f;
        // ldmatrix row addresses (bytes inside one [16][24] bf16 matrix): lane -> matrix m = lane/8, row r = lane%8
        const int lm = lane >> 3, lr = lane & 7;
        const uint32_t offA0 = (uint32_t)((((lm >> 1) * 8 + lr) * 24 + (lm & 1) * 8) * 2);      // tiles (p0,i0) = (0,0) (0,8) (8,0) (8,8)
        const uint32_t offA1 = (uint32_t)((((lm & 1) * 8 + lr) * 24 + 16) * 2);                 // tiles (0,16) (8,16)
        const uint32_t offB0 = (uint32_t)((((lm & 1) * 8 + lr) * 24 + (lm >> 1) * 8) * 2);      // tiles (0,0) (8,0) (0,8) (8,8)
        const uint32_t offB1 = offA1;
        #pragma unroll
        for (int c = 0; c < NC; c++) {
            uint32_t a0[NTERMS][4], a1[NTERMS][2], b0[NTERMS][4], b1[NTERMS][2];
            #pragma unroll
            for (int k = 0; k < NTERMS; k++) {
                const uint32_t mA = sbase + (uint32_t)(((0 * NCS + c) * NTERMS + k) * 768);
                const uint32_t mZ = sbase + (uint32_t)(((1 * NCS + c) * NTERMS + k) * 768);
                ldsm_x4_trans(a0[k], mA + offA0);
                ldsm_x2_trans(a1[k], mA + offA1);
                ldsm_x4_trans(b0[k], mZ + offB0);
                ldsm_x2_trans(b1[k], mZ + offB1);
            }
            auto bfrag = [&](int nt, int kb, int w) -> uint32_t { return nt < 2 ? b0[kb][2 * (nt & 1) + w] : b1[kb][w]; };
            #pragma unroll
            for (int q = 0; q < TP::n; q++)                 // units i = 0..15
                #pragma unroll
                for (int nt = 0; nt < NT; nt++)
                    mma_k16(dw[0][nt], a0[TP::a(q)][0], a0[TP::a(q)][1], a0[TP::a(q)][2], a0[TP::a(q)][3],
                            bfrag(nt, TP::b(q), 0), bfrag(nt, TP::b(q), 1));
            // units i = 16..19 fill only rows 0..7 of their m16 tile: rows 8..15 take a SECOND A term that meets the same
            // B term (A terms ka with ka + kb <= NTERMS - 1), and are folded into rows 0..7 after the channel loop
            #pragma unroll
            for (int kb = 0; kb < NTERMS; kb++)
                #pragma unroll
                for (int ka = 0; ka + kb < NTERMS; ka += 2)
                    #pragma unroll
                    for (int nt = 0; nt < NT; nt++) {
                        const bool two = ka + 1 + kb < NTERMS;
                        mma_k16(dw[1][nt], a1[ka][0], two ? a1[ka + 1][0] : 0u, a1[ka][1], two ? a1[ka + 1][1] : 0u,
                                bfrag(nt, kb, 0), bfrag(nt, kb, 1));
                    }
        }
        __syncwarp();                                       // buffer reads done before the next layer overwrites it
        const float usw = F16 ? pow2i(-K) : 1.0f;
        #pragma unroll
        for (int nt = 0; nt < NT; nt++)
            #pragma unroll
            for (int e = 0; e < 2; e++) {
                if (!jv[nt][e]) continue;
                float* col = part + nd.offW[l] + ju[nt][e];
                emit(col + (g) * HT, dw[0][nt][e] * usw);
                emit(col + (g + 8) * HT, dw[0][nt][2 + e] * usw);
                if (g < HT - 16) emit(col + (16 + g) * HT, (dw[1][nt][e] + dw[1][nt][2 + e]) * usw);
            }
    }
}

template <int DIN, int N2, int HT, int NTERMS, bool F16 = false>
__global__ void __launch_bounds__(kMmaWarps * 32, 3)
fused_step_mma(const NetDims nd, const MmaArgs args) {
    static_assert(HT == 20, "fragment schedule: one k16 step + an 8-wide remainder, three n8 tiles");
    constexpr int NC = 1 + DIN + N2;
    extern __shared__ __align__(16) uint32_t smem_u[];
    const int warp = threadIdx.x >> 5;
    // per-warp transposition buffer: [which: A | Zbar][c][term][16 points][12 words = 24 bf16]
    uint32_t* sbuf = smem_u + warp * (2 * NC * NTERMS * 192);
    const int row = args.row0 + blockIdx.x * kMmaWarps + warp;
    float* part = args.partial + (size_t)row * nd.PP;
    float4* stash = reinterpret_cast<float4*>(args.stash + (size_t)row * args.stash_per_row);
    // tile list of the launch: interior tiles, then the value-only tiles of the boundary / initial sets
    const int n_full = args.set.ntiles;
    const int n_v0 = args.nvsets > 0 ? args.vsets[0].ntiles : 0, n_v1 = args.nvsets > 1 ? args.vsets[1].ntiles : 0;
    const int total = n_full + n_v0 + n_v1;
    bool first = true;
    for (int tile = blockIdx.x * kMmaWarps + warp; tile < total; tile += gridDim.x * kMmaWarps) {
        if (tile < n_full) {
            mma_tile<DIN, N2, HT, NTERMS, false, F16>(nd, args, args.set, tile, sbuf, part, stash, first);
        } else {
            const bool second = tile >= n_full + n_v0;
            const PointSet ps = second ? args.vsets[1] : args.vsets[0];
            mma_tile<DIN, N2, HT, NTERMS, true, F16>(nd, args, ps, tile - n_full - (second ? n_v0 : 0), sbuf, part, stash, first);
        }
        first = false;
    }
}

}  // namespace pinn
