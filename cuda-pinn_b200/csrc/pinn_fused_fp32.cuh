// pinn_fused_fp32.cuh — the fused FP32 train-step kernel (CUDA-core path, any width H % 4 == 0).
//
// One persistent CTA walks point tiles of TP points.  For a tile, the whole network runs inside
// the CTA: forward with all NC Taylor channels, residual / loss seeds, reverse pass with the
// three contractions per layer (forward Z = A W, dA = Zbar W^T, dW = A^T Zbar).  Activations never
// leave the SM except the per-layer stash (a, z', z'') the reverse pass re-reads, which lives in a
// per-CTA scratch ring that stays L2-resident.
//
// Data layout inside the CTA (SURVEY.md Appendix B math, row-major batch-as-rows like the
// reference matmul kernel.cuh:68-83, i.e. W[l] is [fan_in, fan_out]):
//   bufA / bufZ : shared, [row = unit k][col = c*TP + p] floats, leading dimension LD = NC*TP+4
//                 (unit-major so that lanes = points read conflict-free; +4 keeps float4 row reads
//                 of the dW contraction on distinct bank groups)
//   thread (p, g): p = tid % TP is the point, g = tid / TP owns units [g*UT, g*UT+UT)
//                  → the tanh-Taylor elementwise work is thread-local (all channels in registers)
//   weights      : read through the read-only path (L1-resident, broadcast across the warp)
//   gradients    : every CTA owns a private partial row partial[cta][slot][PP]; each element has
//                  exactly one owning thread (no atomics → bit-reproducible); the reduce kernel
//                  sums rows in fixed order.
#pragma once
#include "pinn_device.cuh"

namespace pinn {

struct PointSet {
    const float* x;             // [n, d_in] row-major (reference tensor layout)
    unsigned long long n;       // local count
    float scale;                // 2 * lambda / N_global
    int set;                    // pinn_point_set
    int ntiles;
};

struct FusedArgs {
    const float* params;        // flat [P]
    const float* paramsT;       // same offsets, W^T per layer ([fan_out, fan_in])
    PointSet sets[2];
    int nsets;
    float* partial;             // [grid][nslot][PP]
    float* stash;               // [grid][stash_per_cta]
    int nslot;
    long long stash_per_cta;    // floats
    int mode;                   // 0 = train (loss + gradients), 1 = evaluate (y, r out)
    float* y_out;               // eval: [n, d_out, NC]
    float* r_out;               // eval: [n, n_res] (interior only)
};

// acc[c][u] = sum_k bufIn[k][c*TP+p] * Wg[k*N + j0+u]      (k ascending, zero-initialised)
template <int NC, int UT, int TP>
__device__ __forceinline__ void gemm_rows(const float* __restrict__ bufIn, int LD,
                                          const float* __restrict__ Wg, int K, int N, int j0, int p,
                                          float (&acc)[NC][UT]) {
    #pragma unroll
    for (int c = 0; c < NC; c++)
        #pragma unroll
        for (int u = 0; u < UT; u++) acc[c][u] = 0.0f;
    if (j0 >= N) return;
    const float* a_ptr = bufIn + p;
    if ((j0 + UT <= N) && ((N & 3) == 0)) {
        const float* w_ptr = Wg + j0;
        #pragma unroll 2
        for (int k = 0; k < K; k++) {
            float a[NC], w[UT];
            #pragma unroll
            for (int c = 0; c < NC; c++) a[c] = a_ptr[k * LD + c * TP];
            #pragma unroll
            for (int q = 0; q < UT / 4; q++) {
                float4 t = __ldg(reinterpret_cast<const float4*>(w_ptr + (size_t)k * N) + q);
                w[4 * q + 0] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
            }
            #pragma unroll
            for (int c = 0; c < NC; c++)
                #pragma unroll
                for (int u = 0; u < UT; u++) acc[c][u] = fmaf(a[c], w[u], acc[c][u]);
        }
    } else {
        for (int k = 0; k < K; k++) {
            float a[NC];
            #pragma unroll
            for (int c = 0; c < NC; c++) a[c] = a_ptr[k * LD + c * TP];
            #pragma unroll
            for (int u = 0; u < UT; u++) {
                float w = (j0 + u < N) ? __ldg(Wg + (size_t)k * N + j0 + u) : 0.0f;
                #pragma unroll
                for (int c = 0; c < NC; c++) acc[c][u] = fmaf(a[c], w, acc[c][u]);
            }
        }
    }
}

// part[k*fo + j] += sum_m bufA[k][m] * bufZ[j][m]  over the tile's M = NC*TP columns.
// 4x4 register micro-tiles with strided row ownership; when there are fewer micro-tiles than
// threads the column range is split over `nslot` partial rows (summed later in fixed order).
template <int NC, int TP>
__device__ __forceinline__ void dw_tile(const float* __restrict__ bufA, const float* __restrict__ bufZ,
                                        int LD, int fi, int fo, float* __restrict__ part, int PP, int nslot) {
    constexpr int M4 = NC * TP / 4;
    const int NKB = (fi + 3) >> 2, NJB = (fo + 3) >> 2, ntiles = NKB * NJB;
    const int nthreads = blockDim.x;
    int nsplit = nthreads / ntiles;
    nsplit = nsplit < 1 ? 1 : (nsplit > nslot ? nslot : nsplit);
    if (nsplit > M4) nsplit = M4;
    for (int w = threadIdx.x; w < ntiles * nsplit; w += nthreads) {
        const int tile = w % ntiles, sp = w / ntiles;
        const int kb = tile / NJB, jb = tile - kb * NJB;
        const int m0 = (M4 * sp / nsplit) * 4, m1 = (M4 * (sp + 1) / nsplit) * 4;
        const float* ar[4]; const float* zr[4];
        #pragma unroll
        for (int i = 0; i < 4; i++) {
            int k = kb + i * NKB; k = k < fi ? k : fi - 1;
            int j = jb + i * NJB; j = j < fo ? j : fo - 1;
            ar[i] = bufA + k * LD; zr[i] = bufZ + j * LD;
        }
        float acc[4][4];
        #pragma unroll
        for (int i = 0; i < 4; i++)
            #pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = 0.0f;
        for (int m = m0; m < m1; m += 4) {
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
        float* dst = part + (size_t)sp * PP;
        #pragma unroll
        for (int i = 0; i < 4; i++) {
            const int k = kb + i * NKB;
            #pragma unroll
            for (int j = 0; j < 4; j++) {
                const int jj = jb + j * NJB;
                if (k < fi && jj < fo) dst[k * fo + jj] += acc[i][j];
            }
        }
    }
}

template <int DIN, int N2, bool FULL, int UT, int TP>
__global__ void __launch_bounds__(512)
fused_step_fp32(const NetDims nd, const FusedArgs args) {
    constexpr int NC = FULL ? 1 + DIN + N2 : 1;
    constexpr int LD = NC * TP + 4;
    extern __shared__ __align__(16) float smem[];
    const int H = nd.H, L = nd.L;
    const int rows = (H + 3) & ~3;
    float* bufA = smem;
    float* bufZ = smem + (size_t)rows * LD;

    const int tid = threadIdx.x;
    const int p = tid % TP, g = tid / TP, j0 = g * UT;
    float* stash = args.stash + (size_t)blockIdx.x * args.stash_per_cta;
    float* part = args.partial + (size_t)blockIdx.x * args.nslot * nd.PP;
    const int total_tiles = args.sets[0].ntiles + (args.nsets > 1 ? args.sets[1].ntiles : 0);
    const size_t lstride = (size_t)NC * H * TP;          // stash floats per layer

    float acc[NC][UT];

    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int si = (t < args.sets[0].ntiles) ? 0 : 1;
        const PointSet ps = args.sets[si];
        const unsigned long long base = (unsigned long long)(t - (si ? args.sets[0].ntiles : 0)) * TP;
        const bool valid = base + p < ps.n;
        float x[3] = {0.0f, 0.0f, 0.0f};
        if (valid) {
            #pragma unroll
            for (int k = 0; k < DIN; k++) x[k] = __ldg(ps.x + (base + p) * DIN + k);
        }
        __syncthreads();                                  // previous tile is done with bufA / bufZ
        // ---- network inputs: value = coordinates, first tangents = unit vectors, second = 0 ----
        if (g == 0) {
            #pragma unroll
            for (int k = 0; k < DIN; k++)
                #pragma unroll
                for (int c = 0; c < NC; c++)
                    bufA[k * LD + c * TP + p] = (c == 0) ? x[k] : ((c == 1 + k) ? 1.0f : 0.0f);
        }
        __syncthreads();

        // ---------------------------------- forward ----------------------------------
        for (int l = 1; l <= L; l++) {
            gemm_rows<NC, UT, TP>(bufA, LD, args.params + nd.offW[l], l == 1 ? DIN : H, H, j0, p, acc);
            float* st = stash + (size_t)(l - 1) * lstride + (size_t)j0 * TP + p;
            #pragma unroll
            for (int u = 0; u < UT; u++) {
                const float a = tanhf(acc[0][u] + __ldg(args.params + nd.offB[l] + j0 + u));
                const float s = 1.0f - a * a;
                st[(size_t)u * TP] = a;
                acc[0][u] = a;
                if constexpr (FULL) {
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
            }
            __syncthreads();                              // every warp finished reading bufA
            #pragma unroll
            for (int u = 0; u < UT; u++)
                #pragma unroll
                for (int c = 0; c < NC; c++) bufA[(j0 + u) * LD + c * TP + p] = acc[c][u];
            __syncthreads();
        }

        // ---- linear output layer, residual / target, loss seeds ----
        gemm_rows<NC, UT, TP>(bufA, LD, args.params + nd.offW[L + 1], H, nd.d_out, j0, p, acc);
        if (g == 0) {
            float y[3][NC], ybar[3][NC], r[3] = {0.0f, 0.0f, 0.0f};
            #pragma unroll
            for (int o = 0; o < 3; o++)
                #pragma unroll
                for (int c = 0; c < NC; c++) {
                    float v = (o < UT) ? acc[c][o < UT ? o : 0] : 0.0f;
                    if (c == 0 && o < nd.d_out) v += __ldg(args.params + nd.offB[L + 1] + o);
                    y[o][c] = (o < nd.d_out) ? v : 0.0f;
                }
            float ss = 0.0f;
            if constexpr (FULL) {
                ss = pde_residual<NC>(nd, x, y, ps.scale, ybar, r);
            } else {
                float gt[3] = {0.0f, 0.0f, 0.0f};
                pde_target(nd, ps.set, x, gt);
                #pragma unroll
                for (int o = 0; o < 3; o++) {
                    const float e = (o < nd.d_out) ? y[o][0] - gt[o] : 0.0f;
                    ss += e * e;
                    ybar[o][0] = ps.scale * e;
                }
            }
            if (args.mode == 1) {                         // evaluate: write channels and residuals
                if (valid) {
                    for (int o = 0; o < nd.d_out; o++)
                        #pragma unroll
                        for (int c = 0; c < NC; c++)
                            args.y_out[((base + p) * nd.d_out + o) * NC + c] = y[o][c];
                    if (FULL && args.r_out)
                        for (int q = 0; q < nd.n_res; q++) args.r_out[(base + p) * nd.n_res + q] = r[q];
                }
            } else {
                if (!valid) {
                    ss = 0.0f;
                    #pragma unroll
                    for (int o = 0; o < 3; o++)
                        #pragma unroll
                        for (int c = 0; c < NC; c++) ybar[o][c] = 0.0f;
                }
                for (int o = 0; o < nd.d_out; o++) {
                    #pragma unroll
                    for (int c = 0; c < NC; c++) bufZ[o * LD + c * TP + p] = ybar[o][c];
                    const float sb = lane_group_sum<TP>(ybar[o][0]);
                    if (p == 0) part[nd.offB[L + 1] + o] += sb;
                }
                const float sl = lane_group_sum<TP>(ss);
                if (p == 0) part[nd.P + ps.set] += sl;
            }
        }
        if (args.mode == 1) continue;
        __syncthreads();

        // ---------------------------------- reverse ----------------------------------
        dw_tile<NC, TP>(bufA, bufZ, LD, H, nd.d_out, part + nd.offW[L + 1], nd.PP, args.nslot);
        gemm_rows<NC, UT, TP>(bufZ, LD, args.paramsT + nd.offW[L + 1], nd.d_out, H, j0, p, acc);
        __syncthreads();

        for (int l = L; l >= 1; l--) {
            // elementwise adjoint of the tanh layer: acc = Abar_l  →  Zbar_l into bufZ
            const float* st = stash + (size_t)(l - 1) * lstride + (size_t)j0 * TP + p;
            #pragma unroll
            for (int u = 0; u < UT; u++) {
                const float a = st[(size_t)u * TP];
                const float s = 1.0f - a * a;
                float zb0 = s * acc[0][u];
                float* zrow = bufZ + (j0 + u) * LD + p;
                if constexpr (FULL) {
                    #pragma unroll
                    for (int i = 0; i < DIN; i++) {
                        const float zp = st[(size_t)((1 + i) * H + u) * TP];
                        const float abp = acc[1 + i][u];
                        float zbp = s * abp;
                        zb0 += (-2.0f * a * s * zp) * abp;
                        if (i < N2) {
                            const float zpp = st[(size_t)((1 + DIN + i) * H + u) * TP];
                            const float abpp = acc[1 + DIN + i][u];
                            zrow[(1 + DIN + i) * TP] = s * abpp;
                            zbp -= 4.0f * a * s * zp * abpp;
                            zb0 += (-2.0f * a * s * zpp - 2.0f * s * (1.0f - 3.0f * a * a) * zp * zp) * abpp;
                        }
                        zrow[(1 + i) * TP] = zbp;
                    }
                }
                zrow[0] = zb0;
                const float sb = lane_group_sum<TP>(zb0);
                if (p == 0) part[nd.offB[l] + j0 + u] += sb;
            }
            // A_{l-1} (the dW operand) back into bufA: recomputed from the stash, or the inputs
            if (l > 1) {
                const float* sp = stash + (size_t)(l - 2) * lstride + (size_t)j0 * TP + p;
                #pragma unroll
                for (int u = 0; u < UT; u++) {
                    const float a = sp[(size_t)u * TP];
                    float* arow = bufA + (j0 + u) * LD + p;
                    arow[0] = a;
                    if constexpr (FULL) {
                        const float s = 1.0f - a * a;
                        #pragma unroll
                        for (int i = 0; i < DIN; i++) {
                            const float zp = sp[(size_t)((1 + i) * H + u) * TP];
                            arow[(1 + i) * TP] = s * zp;
                            if (i < N2) {
                                const float zpp = sp[(size_t)((1 + DIN + i) * H + u) * TP];
                                arow[(1 + DIN + i) * TP] = s * zpp - 2.0f * a * s * zp * zp;
                            }
                        }
                    }
                }
            } else if (g == 0) {
                #pragma unroll
                for (int k = 0; k < DIN; k++)
                    #pragma unroll
                    for (int c = 0; c < NC; c++)
                        bufA[k * LD + c * TP + p] = (c == 0) ? x[k] : ((c == 1 + k) ? 1.0f : 0.0f);
            }
            __syncthreads();
            dw_tile<NC, TP>(bufA, bufZ, LD, l == 1 ? DIN : H, H, part + nd.offW[l], nd.PP, args.nslot);
            if (l > 1) gemm_rows<NC, UT, TP>(bufZ, LD, args.paramsT + nd.offW[l], H, H, j0, p, acc);
            __syncthreads();
        }
    }
}

}  // namespace pinn
