// pinn_fused_warp.cuh — warp-autonomous fused FP32 train step for narrow networks (width HT <= 32,
// fixed at compile time): the latency-oriented kernel of BASELINE configs C1 / C2.
//
// One WARP owns a tile of 32 collocation points (lane = point) and walks the whole network for it
// with no block-level barrier at all:
//   * all parameters (W and the transposed mirror) sit in shared memory once per CTA; a contraction
//     step reads one weight row as broadcast LDS.128 and feeds NC*HT independent FFMAs per thread;
//   * the activations of the current layer — every Taylor channel of every unit of the lane's point —
//     live in REGISTERS (NC*HT accumulators), so the forward / dA contractions touch no activation
//     memory; the contraction loops are fully unrolled (offsets are immediates);
//   * dW = A^T Zbar needs a cross-lane reduction: the warp writes A_{l-1} and Zbar_l to its private
//     shared-memory panels ([unit][c*32 + lane]) and 30 lanes run a 4x4-register-tile micro-GEMM over
//     the 128 columns.  A constant row of ones (value-channel columns) appended to the A panel makes
//     the bias gradient row HT of the same product — no shuffle reductions;
//   * gradient tiles are added to the warp's private partial row with fire-and-forget RED (single
//     owner per element => bit-reproducible);
//   * the (a, z', z'') stash of the tile is a per-warp ring in global memory (L2-resident).  A lone
//     warp cannot hide L2 latency, so the reverse pass never reads it directly: two shared-memory
//     panels R0/R1 alternate per layer — cp.async (LDGSTS) prefetches the stash of layer l-1 into one
//     while the adjoint of layer l consumes the other and overwrites it in place with Zbar_l (every
//     panel column is private to its lane, so no barrier is involved).
#pragma once
#include "pinn_device.cuh"
#include "pinn_fused_fp32.cuh"

namespace pinn {

// dW tile product for one warp.  bufA: [FI1 rows][LD] (row FI1-1 = ones row), bufZ: [FO rows][LD],
// M = NC*32 columns.  Row k < FI1-1 of the result goes to gw[k*FO + j], the ones row to gb[j].
template <int NC, int FI1, int FO>
__device__ __forceinline__ void warp_dw(const float* __restrict__ bufA, const float* __restrict__ bufZ,
                                        float* __restrict__ gw, float* __restrict__ gb, int lane) {
    constexpr int LD = NC * 32 + 4, M4 = NC * 8;
    constexpr int NKB = (FI1 + 3) / 4, NJB = (FO + 3) / 4, NT = NKB * NJB;
    for (int tile = lane; tile < NT; tile += 32) {
        const int kb = tile / NJB, jb = tile - kb * NJB;
        const float* ar[4]; const float* zr[4];
        #pragma unroll
        for (int i = 0; i < 4; i++) {
            int k = kb + i * NKB; k = k < FI1 ? k : FI1 - 1;
            int j = jb + i * NJB; j = j < FO ? j : FO - 1;
            ar[i] = bufA + k * LD; zr[i] = bufZ + j * LD;
        }
        float acc[4][4];
        #pragma unroll
        for (int i = 0; i < 4; i++)
            #pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = 0.0f;
        #pragma unroll 2
        for (int m = 0; m < M4 * 4; m += 4) {
            float4 a[4], z[4];
            #pragma unroll
            for (int i = 0; i < 4; i++) {
                a[i] = *reinterpret_cast<const float4*>(ar[i] + m);
                z[i] = *reinterpret_cast<const float4*>(zr[i] + m);
            }
            #pragma unroll
            for (int i = 0; i < 4; i++)
                #pragma unroll
                for (int j = 0; j < 4; j++) {
                    acc[i][j] = fmaf(a[i].x, z[j].x, acc[i][j]);
                    acc[i][j] = fmaf(a[i].y, z[j].y, acc[i][j]);
                    acc[i][j] = fmaf(a[i].z, z[j].z, acc[i][j]);
                    acc[i][j] = fmaf(a[i].w, z[j].w, acc[i][j]);
                }
        }
        #pragma unroll
        for (int i = 0; i < 4; i++) {
            const int k = kb + i * NKB;
            #pragma unroll
            for (int j = 0; j < 4; j++) {
                const int jj = jb + j * NJB;
                if (jj < FO) {
                    if (k < FI1 - 1) red_add(gw + k * FO + jj, acc[i][j]);
                    else if (k == FI1 - 1) red_add(gb + jj, acc[i][j]);
                }
            }
        }
    }
}

template <int DIN, int N2, int HT, int WPC>
__global__ void __launch_bounds__(WPC * 32, 1)
fused_step_warp(const NetDims nd, const FusedArgs args) {
    constexpr int NC = 1 + DIN + N2, LD = NC * 32 + 4;
    constexpr int RA = HT + 1;                       // A panel rows: HT units + the ones row
    static_assert(HT % 4 == 0 && HT <= 32, "narrow widths only");
    extern __shared__ __align__(16) float smem[];
    const int L = nd.L, P4 = (nd.P + 3) & ~3;
    float* Ws = smem;                                // flat parameters (same offsets as the global vector)
    float* WTs = Ws + P4;                            // transposed mirror
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* bufA = WTs + P4 + (size_t)warp * (RA + 2 * HT) * LD;   // A panel (+ ones row)
    float* R0 = bufA + (size_t)RA * LD;                            // stash / Zbar panels, alternating by layer parity
    float* R1 = R0 + (size_t)HT * LD;
    auto Rof = [&](int l) -> float* { return (l & 1) ? R1 : R0; };

    for (int i = tid * 4; i < P4; i += WPC * 32 * 4) {
        *reinterpret_cast<float4*>(Ws + i) = __ldg(reinterpret_cast<const float4*>(args.params + i));
        *reinterpret_cast<float4*>(WTs + i) = __ldg(reinterpret_cast<const float4*>(args.paramsT + i));
    }
    #pragma unroll
    for (int c = 0; c < NC; c++) bufA[HT * LD + c * 32 + lane] = (c == 0) ? 1.0f : 0.0f;   // ones row (hidden layers)
    __syncthreads();                                 // the only block barrier: weights are in place

    const int wrow = args.row0 + blockIdx.x * WPC + warp;
    float* stash = args.stash + (size_t)wrow * args.stash_per_cta;
    float* part = args.partial + (size_t)wrow * nd.PP;
    const PointSet ps = args.sets[0];
    const int ntiles = ps.ntiles;
    constexpr size_t lstride = (size_t)NC * HT * 32;
    // Forward activations ping-pong between the A panel and the free R panel so that A_L lands in
    // the A panel (it is the dW_{L+1} operand) while stash_L sits in R(L).
    auto Fof = [&](int l) -> float* { return ((L - l) & 1) ? Rof(L + 1) : bufA; };

    float z[NC][HT];                                 // the NC x HT accumulators of this lane's point

    // contraction z[c][j] = sum_k P[k][c] * W[k*HT + j]: activations from a panel (lane-private LDS),
    // weight row as broadcast LDS.128; rolled over k (code stays inside the instruction cache)
    auto contract = [&](const float* __restrict__ Pn, const float* __restrict__ W) {
        #pragma unroll 4
        for (int k = 0; k < HT; k++) {
            float av[NC], w[HT];
            #pragma unroll
            for (int c = 0; c < NC; c++) av[c] = Pn[k * LD + c * 32 + lane];
            #pragma unroll
            for (int q = 0; q < HT / 4; q++) {
                const float4 tq = *reinterpret_cast<const float4*>(W + k * HT + 4 * q);
                w[4 * q] = tq.x; w[4 * q + 1] = tq.y; w[4 * q + 2] = tq.z; w[4 * q + 3] = tq.w;
            }
            #pragma unroll
            for (int c = 0; c < NC; c++)
                #pragma unroll
                for (int j = 0; j < HT; j++) z[c][j] = fmaf(av[c], w[j], z[c][j]);
        }
    };
    auto dump = [&](float* __restrict__ Pn) {        // accumulators -> panel (lane-private columns)
        #pragma unroll
        for (int j = 0; j < HT; j++)
            #pragma unroll
            for (int c = 0; c < NC; c++) Pn[j * LD + c * 32 + lane] = z[c][j];
    };
    // tanh + Taylor channels in place on a panel holding the pre-activations; stash to `st`
    auto act_forward = [&](float* __restrict__ Pn, float* __restrict__ st, int us, int cs) {
        #pragma unroll 2
        for (int j = 0; j < HT; j++) {
            float* f = Pn + j * LD + lane;
            const float a0 = tanhf(f[0]), s = 1.0f - a0 * a0;
            st[j * us] = a0; f[0] = a0;
            #pragma unroll
            for (int i = 0; i < DIN; i++) {
                const float zp = f[(1 + i) * 32];
                st[(1 + i) * cs + j * us] = zp;
                f[(1 + i) * 32] = s * zp;
                if (i < N2) {
                    const float zpp = f[(1 + DIN + i) * 32];
                    st[(1 + DIN + i) * cs + j * us] = zpp;
                    f[(1 + DIN + i) * 32] = s * zpp - 2.0f * a0 * s * zp * zp;
                }
            }
        }
    };

    for (int t = blockIdx.x + gridDim.x * warp; t < ntiles; t += gridDim.x * WPC) {
        const unsigned long long base = (unsigned long long)t * 32;
        const bool valid = base + lane < ps.n;
        float x[DIN];
        #pragma unroll
        for (int k = 0; k < DIN; k++) x[k] = valid ? __ldg(ps.x + (base + lane) * DIN + k) : 0.0f;

        // ---------------- forward ----------------
        // layer 1: value = b + x W1, first tangents = rows of W1, second tangents = 0
        {
            const float* W1 = Ws + nd.offW[1];
            const float* b1 = Ws + nd.offB[1];
            float* F1 = Fof(1);
            #pragma unroll 2
            for (int j = 0; j < HT; j++) {
                float* f = F1 + j * LD + lane;
                float z0 = b1[j];
                #pragma unroll
                for (int k = 0; k < DIN; k++) z0 = fmaf(x[k], W1[k * HT + j], z0);
                f[0] = z0;
                #pragma unroll
                for (int i = 0; i < DIN; i++) {
                    f[(1 + i) * 32] = W1[i * HT + j];
                    if (i < N2) f[(1 + DIN + i) * 32] = 0.0f;
                }
            }
            if (L == 1) act_forward(F1, Rof(1) + lane, LD, 32);
            else act_forward(F1, stash + lane, 32, HT * 32);
        }
        for (int l = 2; l <= L; l++) {
            const float* b = Ws + nd.offB[l];
            #pragma unroll
            for (int j = 0; j < HT; j++) {
                z[0][j] = b[j];
                #pragma unroll
                for (int c = 1; c < NC; c++) z[c][j] = 0.0f;
            }
            contract(Fof(l - 1), Ws + nd.offW[l]);
            float* Fl = Fof(l);
            dump(Fl);
            if (l == L) act_forward(Fl, Rof(L) + lane, LD, 32);      // the last stash goes straight to its panel
            else act_forward(Fl, stash + (size_t)(l - 1) * lstride + lane, 32, HT * 32);
        }

        // ---------------- output layer, residual, seeds (A_L is in the A panel) ----------------
        float y[3][NC], ybar[3][NC], r[3] = {0.0f, 0.0f, 0.0f};
        {
            const float* Wo = Ws + nd.offW[L + 1];
            #pragma unroll
            for (int o = 0; o < 3; o++)
                #pragma unroll
                for (int c = 0; c < NC; c++) y[o][c] = (c == 0 && o < nd.d_out) ? Ws[nd.offB[L + 1] + o] : 0.0f;
            #pragma unroll 4
            for (int k = 0; k < HT; k++) {
                float av[NC];
                #pragma unroll
                for (int c = 0; c < NC; c++) av[c] = bufA[k * LD + c * 32 + lane];
                #pragma unroll
                for (int o = 0; o < 3; o++)
                    if (o < nd.d_out) {
                        const float wo = Wo[k * nd.d_out + o];
                        #pragma unroll
                        for (int c = 0; c < NC; c++) y[o][c] = fmaf(av[c], wo, y[o][c]);
                    }
            }
        }
        float ss = pde_residual<NC>(nd, x, y, ps.scale, ybar, r);
        if (!valid) {
            ss = 0.0f;
            #pragma unroll
            for (int o = 0; o < 3; o++)
                #pragma unroll
                for (int c = 0; c < NC; c++) ybar[o][c] = 0.0f;
        }
        ss = lane_group_sum<32>(ss);
        if (lane == 0) red_add(part + nd.P + ps.set, ss);

        // ---------------- reverse: output layer ----------------
        float* Rs = Rof(L + 1);                           // the panel not holding stash_L: seeds
        #pragma unroll
        for (int o = 0; o < 3; o++)
            if (o < nd.d_out) {
                #pragma unroll
                for (int c = 0; c < NC; c++) Rs[o * LD + c * 32 + lane] = ybar[o][c];
            }
        __syncwarp();
        if (nd.d_out == 1) warp_dw<NC, RA, 1>(bufA, Rs, part + nd.offW[L + 1], part + nd.offB[L + 1], lane);
        else warp_dw<NC, RA, 3>(bufA, Rs, part + nd.offW[L + 1], part + nd.offB[L + 1], lane);
        __syncwarp();
        {   // Abar_L = seeds x W_{L+1}^T, written to the A panel (the adjoint reads it from there)
            const float* Wo = Ws + nd.offW[L + 1];
            #pragma unroll 4
            for (int k = 0; k < HT; k++) {
                #pragma unroll
                for (int c = 0; c < NC; c++) {
                    float s = 0.0f;
                    #pragma unroll
                    for (int o = 0; o < 3; o++)
                        if (o < nd.d_out) s = fmaf(ybar[o][c], Wo[k * nd.d_out + o], s);
                    bufA[k * LD + c * 32 + lane] = s;
                }
            }
        }

        // ---------------- reverse: hidden layers ----------------
        for (int l = L; l >= 1; l--) {
            float* Rc = Rof(l);                           // holds stash_l; becomes the Zbar_l panel
            float* Rn = Rof(l - 1);                       // receives stash_{l-1}
            if (l > 1) {
                // prefetch stash_{l-1} (global ring) -> Rn, one 4-byte LDGSTS per element, lane-private columns
                const float* src = stash + (size_t)(l - 2) * lstride + lane;
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(Rn + lane);
                #pragma unroll 4
                for (int j = 0; j < HT; j++)
                    #pragma unroll
                    for (int c = 0; c < NC; c++)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;"
                                     ::"r"(dst + (uint32_t)((j * LD + c * 32) * 4)), "l"(src + (size_t)(c * HT + j) * 32) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            // adjoint of the tanh layer: Abar_l (A panel) + stash_l (Rc) -> Zbar_l, Rc overwritten in place
            #pragma unroll 2
            for (int j = 0; j < HT; j++) {
                float* rc = Rc + j * LD + lane;
                const float* ab = bufA + j * LD + lane;
                const float a0 = rc[0];
                const float s = 1.0f - a0 * a0;
                float zb0 = s * ab[0];
                #pragma unroll
                for (int i = 0; i < DIN; i++) {
                    const float zp = rc[(1 + i) * 32];
                    const float abp = ab[(1 + i) * 32];
                    float zbp = s * abp;
                    zb0 += (-2.0f * a0 * s * zp) * abp;
                    if (i < N2) {
                        const float zpp = rc[(1 + DIN + i) * 32];
                        const float abpp = ab[(1 + DIN + i) * 32];
                        rc[(1 + DIN + i) * 32] = s * abpp;
                        zbp -= 4.0f * a0 * s * zp * abpp;
                        zb0 += (-2.0f * a0 * s * zpp - 2.0f * s * (1.0f - 3.0f * a0 * a0) * zp * zp) * abpp;
                    }
                    rc[(1 + i) * 32] = zbp;
                }
                rc[0] = zb0;
            }
            if (l > 1) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                // A_{l-1} from the prefetched stash into the A panel (the raw stash stays in Rn for the next adjoint)
                #pragma unroll 2
                for (int k = 0; k < HT; k++) {
                    const float* rn = Rn + k * LD + lane;
                    float* pa = bufA + k * LD + lane;
                    const float a0 = rn[0];
                    const float s = 1.0f - a0 * a0;
                    pa[0] = a0;
                    #pragma unroll
                    for (int i = 0; i < DIN; i++) {
                        const float zp = rn[(1 + i) * 32];
                        pa[(1 + i) * 32] = s * zp;
                        if (i < N2) {
                            const float zpp = rn[(1 + DIN + i) * 32];
                            pa[(1 + DIN + i) * 32] = s * zpp - 2.0f * a0 * s * zp * zp;
                        }
                    }
                }
                __syncwarp();
                warp_dw<NC, RA, HT>(bufA, Rc, part + nd.offW[l], part + nd.offB[l], lane);
                // Abar_{l-1} = Zbar_l W_l^T into the accumulators, then to the A panel for the next adjoint
                #pragma unroll
                for (int c = 0; c < NC; c++)
                    #pragma unroll
                    for (int k = 0; k < HT; k++) z[c][k] = 0.0f;
                contract(Rc, WTs + nd.offW[l]);
                __syncwarp();                         // every lane is done with the panels of this layer
                dump(bufA);
            } else {
                // layer 1: A_0 rows = inputs (value x_k, unit first tangents, zero second tangents), then a ones row
                #pragma unroll
                for (int k = 0; k < DIN; k++)
                    #pragma unroll
                    for (int c = 0; c < NC; c++) bufA[k * LD + c * 32 + lane] = (c == 0) ? x[k] : ((c == 1 + k) ? 1.0f : 0.0f);
                #pragma unroll
                for (int c = 0; c < NC; c++) bufA[DIN * LD + c * 32 + lane] = (c == 0) ? 1.0f : 0.0f;
                __syncwarp();
                warp_dw<NC, DIN + 1, HT>(bufA, Rc, part + nd.offW[1], part + nd.offB[1], lane);
                __syncwarp();
            }
        }
    }
}

}  // namespace pinn
