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
    float* partial;             // [rows][PP]: one row per CTA of every launch of the step
    float* stash;               // [rows][stash_per_cta]
    int row0;                   // first partial / stash row of this launch (launches of one step run concurrently)
    int nsplit_cap;             // > 1: dW micro-tiles may split the tile columns, partials meet in smem scratch
    long long stash_per_cta;    // floats
    int mode;                   // 0 = train (loss + gradients), 1 = evaluate (y, r out)
    float* y_out;               // eval: [n, d_out, NC]
    float* r_out;               // eval: [n, n_res] (interior only)
};

// acc[c][u] = sum_k bufIn[k][c*TP+p] * Wg[k*N + j0+u]      (k ascending, zero-initialised)
// WS = true: Wg is a shared-memory copy staged by cp.async (narrow nets); false: global, read-only path.
template <int NC, int UT, int TP, bool WS = false, int KU = 2>
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
        #pragma unroll KU
        for (int k = 0; k < K; k++) {
            float a[NC], w[UT];
            #pragma unroll
            for (int c = 0; c < NC; c++) a[c] = a_ptr[k * LD + c * TP];
            #pragma unroll
            for (int q = 0; q < UT / 4; q++) {
                float4 t = WS ? *(reinterpret_cast<const float4*>(w_ptr + (size_t)k * N) + q)
                              : __ldg(reinterpret_cast<const float4*>(w_ptr + (size_t)k * N) + q);
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
                float w = (j0 + u < N) ? (WS ? Wg[(size_t)k * N + j0 + u] : __ldg(Wg + (size_t)k * N + j0 + u)) : 0.0f;
                #pragma unroll
                for (int c = 0; c < NC; c++) acc[c][u] = fmaf(a[c], w, acc[c][u]);
            }
        }
    }
}

// Gradient accumulation into this CTA's private partial row: every element has exactly one owning
// thread, so a fire-and-forget reduction (RED.ADD, no return value => no stall on the L2 round trip)
// is still applied in program order per address => bit-reproducible.
__device__ __forceinline__ void red_add(float* addr, float v) { atomicAdd(addr, v); }

__device__ __forceinline__ int dw_nsplit(int ntiles, int nsplit_cap, int M4, int nthreads) {
    int nsplit = nthreads / ntiles;
    nsplit = nsplit < 1 ? 1 : (nsplit > nsplit_cap ? nsplit_cap : nsplit);
    return nsplit > M4 ? M4 : nsplit;
}

// dW[k][j] = sum_m bufA[k][m] * bufZ[j][m] over the tile's M = NC*TP columns: 4x4 register
// micro-tiles with strided row ownership.  With fewer micro-tiles than threads the columns are split
// over `nsplit` thread groups whose partial tiles go to `scratch` (summed by dw_flush after the next
// barrier); otherwise the tile is added straight into the partial row.
template <int NC, int TP>
__device__ __forceinline__ void dw_compute(const float* __restrict__ bufA, const float* __restrict__ bufZ,
                                           int LD, int fi, int fo, float* __restrict__ part,
                                           float* __restrict__ scratch, int nsplit_cap, int nthreads) {
    constexpr int M4 = NC * TP / 4;
    const int NKB = (fi + 3) >> 2, NJB = (fo + 3) >> 2, ntiles = NKB * NJB;
    const int nsplit = dw_nsplit(ntiles, nsplit_cap, M4, nthreads);
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
        if (nsplit == 1) {
            #pragma unroll
            for (int i = 0; i < 4; i++) {
                const int k = kb + i * NKB;
                #pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int jj = jb + j * NJB;
                    if (k < fi && jj < fo) red_add(part + k * fo + jj, acc[i][j]);
                }
            }
        } else {
            float4* dst = reinterpret_cast<float4*>(scratch + (size_t)w * 16);
            #pragma unroll
            for (int i = 0; i < 4; i++) dst[i] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
    }
}

// Second half of a split dW: after a barrier, element e of micro-tile `tile` is summed over the
// splits in fixed order by one thread and added to the partial row.
template <int NC, int TP>
__device__ __forceinline__ void dw_flush(int fi, int fo, float* __restrict__ part,
                                         const float* __restrict__ scratch, int nsplit_cap, int nthreads) {
    constexpr int M4 = NC * TP / 4;
    const int NKB = (fi + 3) >> 2, NJB = (fo + 3) >> 2, ntiles = NKB * NJB;
    const int nsplit = dw_nsplit(ntiles, nsplit_cap, M4, nthreads);
    if (nsplit == 1) return;
    for (int idx = threadIdx.x; idx < ntiles * 16; idx += nthreads) {
        const int tile = idx >> 4, e = idx & 15, i = e >> 2, j = e & 3;
        const int kb = tile / NJB, jb = tile - kb * NJB;
        const int k = kb + i * NKB, jj = jb + j * NJB;
        if (k < fi && jj < fo) {
            float sum = 0.0f;
            for (int sp = 0; sp < nsplit; sp++) sum += scratch[(size_t)(sp * ntiles + tile) * 16 + e];
            red_add(part + k * fo + jj, sum);
        }
    }
}

// ---- weight staging for narrow nets: one contraction ahead, cp.async (LDGSTS) double buffer ----
// Schedule of the matrices a tile consumes, in order: W_1..W_L, W_{L+1}, W^T_{L+1}, W^T_L..W^T_2.
__device__ __forceinline__ const float* wsched(const NetDims& nd, const FusedArgs& args, int q, int* nfloat) {
    const int L = nd.L;
    int l; const float* base;
    if (q <= L) { l = q + 1; base = args.params; }                // forward: W_{q+1}
    else { l = 2 * L + 2 - q; base = args.paramsT; }              // reverse: W^T_{L+1}, W^T_L, ...
    *nfloat = nd.fan_in(l) * nd.fan_out(l);
    return base + nd.offW[l];
}
__device__ __forceinline__ void wstage_issue(float* dst, const float* __restrict__ src, int nfloat, int nthreads) {
    const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(dst);
    for (int i = threadIdx.x * 4; i < nfloat; i += nthreads * 4)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + i * 4), "l"(src + i) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void wstage_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// HT > 0 fixes the hidden width at compile time (contraction loops fully unrolled, offsets become
// immediates, no runtime integer division in the dW helpers); HT = 0 reads it from NetDims.
template <int DIN, int N2, bool FULL, int UT, int TP, int MAXT, int MINB, int HT = 0>
__global__ void __launch_bounds__(MAXT, MINB)
fused_step_fp32(const NetDims nd, const FusedArgs args) {
    constexpr int NC = FULL ? 1 + DIN + N2 : 1;
    constexpr int LD = NC * TP + 4;
    extern __shared__ __align__(16) float smem[];
    const int H = HT > 0 ? HT : nd.H, L = nd.L;
    const int NTHR = HT > 0 ? TP * (HT / UT) : (int)blockDim.x;
    const int nsplit_cap = HT > 0 ? 8 : args.nsplit_cap;     // specialised narrow variants always carry the scratch
    const int rows = (H + 3) & ~3;
    float* bufA = smem;
    float* bufZ = smem + (size_t)rows * LD;
    float* scratch = bufZ + (size_t)rows * LD;          // [blockDim.x][16] when nsplit_cap > 1
    constexpr bool WS = (UT == 4);
    constexpr int KU = HT > 0 ? (HT % 5 == 0 ? 5 : 4) : 2;   // unroll of the contraction loops                      // narrow nets: weights staged in smem, one contraction ahead
    float* wbuf = scratch + (nsplit_cap > 1 ? (size_t)NTHR * 16 : 0);   // [2][rows*rows] when WS
    const int wslot = rows * rows;
    const int nq = 2 * L + 1;                           // matrices per tile: q = 0 .. 2L
    int q = 0;                                          // next matrix to consume (block-uniform)
    // consume(): wait for matrix q, make it visible, prefetch matrix q+1 (next tile wraps to W_1)
    auto w_consume = [&]() -> const float* {
        if constexpr (!WS) return nullptr;
        wstage_wait();
        __syncthreads();                                // copy of q visible to all; buffer (q+1)&1 no longer read
        int nf; const int qn = (q + 1 > nq - 1) ? 0 : q + 1;
        const float* src = wsched(nd, args, qn, &nf);
        wstage_issue(wbuf + ((q + 1) & 1) * wslot, src, nf, NTHR);
        const float* cur = wbuf + (q & 1) * wslot;
        q++;
        return cur;
    };
    if constexpr (WS) { int nf; const float* src = wsched(nd, args, 0, &nf); wstage_issue(wbuf, src, nf, NTHR); }

    const int tid = threadIdx.x;
    const int p = tid % TP, g = tid / TP, j0 = g * UT;
    float* stash = args.stash + (size_t)(args.row0 + blockIdx.x) * args.stash_per_cta;
    float* part = args.partial + (size_t)(args.row0 + blockIdx.x) * nd.PP;
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
            if constexpr (WS) { const float* wl = w_consume(); gemm_rows<NC, UT, TP, true, KU>(bufA, LD, wl, l == 1 ? DIN : H, H, j0, p, acc); }
            else gemm_rows<NC, UT, TP>(bufA, LD, args.params + nd.offW[l], l == 1 ? DIN : H, H, j0, p, acc);
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
        if constexpr (WS) { const float* wl = w_consume(); gemm_rows<NC, UT, TP, true, KU>(bufA, LD, wl, H, nd.d_out, j0, p, acc); }
        else gemm_rows<NC, UT, TP>(bufA, LD, args.params + nd.offW[L + 1], H, nd.d_out, j0, p, acc);
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
                    if (p == 0) red_add(part + nd.offB[L + 1] + o, sb);
                }
                const float sl = lane_group_sum<TP>(ss);
                if (p == 0) red_add(part + nd.P + ps.set, sl);
            }
        }
        if (args.mode == 1) {                           // evaluation: skip the reverse matrices of the schedule
            if constexpr (WS) { wstage_wait(); __syncthreads(); int nf; const float* src = wsched(nd, args, 0, &nf); wstage_issue(wbuf, src, nf, NTHR); q = 0; }
            continue;
        }
        __syncthreads();

        // ---------------------------------- reverse ----------------------------------
        dw_compute<NC, TP>(bufA, bufZ, LD, H, nd.d_out, part + nd.offW[L + 1], scratch, nsplit_cap, NTHR);
        if constexpr (WS) { const float* wl = w_consume(); gemm_rows<NC, UT, TP, true, KU>(bufZ, LD, wl, nd.d_out, H, j0, p, acc); }
        else gemm_rows<NC, UT, TP>(bufZ, LD, args.paramsT + nd.offW[L + 1], nd.d_out, H, j0, p, acc);
        __syncthreads();
        dw_flush<NC, TP>(H, nd.d_out, part + nd.offW[L + 1], scratch, nsplit_cap, NTHR);

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
                if (p == 0) red_add(part + nd.offB[l] + j0 + u, sb);
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
            dw_compute<NC, TP>(bufA, bufZ, LD, l == 1 ? DIN : H, H, part + nd.offW[l], scratch, nsplit_cap, NTHR);
            if (l > 1) {
                if constexpr (WS) { const float* wl = w_consume(); gemm_rows<NC, UT, TP, true, KU>(bufZ, LD, wl, H, H, j0, p, acc); }
                else gemm_rows<NC, UT, TP>(bufZ, LD, args.paramsT + nd.offW[l], H, H, j0, p, acc);
            }
            __syncthreads();
            dw_flush<NC, TP>(l == 1 ? DIN : H, H, part + nd.offW[l], scratch, nsplit_cap, NTHR);
        }
        q = 0;                                          // the last prefetch of the tile was W_1 of the next one
    }
    if constexpr (WS) wstage_wait();
}

}  // namespace pinn
