// pinn_device.cuh — device-side building blocks shared by every kernel of the PINN train step:
// network dimension block, counter-based sampler, PDE residuals / targets.
// sm_100a only.  Nothing here falls back to the host.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/pinn_abi.h"

namespace pinn {

constexpr int kMaxLayers = 32;                 // hidden layers
constexpr float kPi        = 3.14159265358979323846f;
constexpr float kBurgersNu = 0.003183098861837907f;   // 0.01 / pi
constexpr float kAcGamma   = 1.0e-4f;
constexpr float kHeatAlpha = 0.05f;
constexpr float kNsNu      = 0.01f;
constexpr float kTwoPi     = 6.28318530717958647692f;

// Everything a kernel needs to know about the network and the PDE; passed by value
// (lives in the constant bank), like the reference passes variables<T> by value (kernel.cuh:11).
struct NetDims {
    int d_in, d_out, H, L, pde;
    int C;          // Taylor channels of the interior set: 1 + d_in + n2
    int n2;         // coordinates that carry a second derivative
    int n_space, has_t, n_res;
    int P;          // parameter count
    int PP;         // padded partial-row length: P + 3 loss sums, rounded up to 32
    int offW[kMaxLayers + 2], offB[kMaxLayers + 2];   // 1-based layer index
    float lo[3], span[3], hi[3];
    float k2;       // Helmholtz k^2
    __host__ __device__ int fan_in(int l) const { return l == 1 ? d_in : H; }
    __host__ __device__ int fan_out(int l) const { return l == L + 1 ? d_out : H; }
};

// ---- sampler: SplitMix64 finaliser over (seed, counter); top 24 bits → [0,1) exactly ----
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t seed, uint64_t ctr) {
    uint64_t z = seed + (ctr + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ float u01(uint64_t seed, uint64_t ctr) {
    return (float)(mix64(seed, ctr) >> 40) * 5.9604644775390625e-08f;   // 2^-24
}

// Point `i` (GLOBAL index) of point set `set`: bit-identical on any rank count.
__device__ __forceinline__ void sample_point(const NetDims& nd, int set, uint64_t seed, uint64_t i, float* out) {
    const int nfaces = 2 * nd.n_space;
    const int face = (int)(i % (uint64_t)nfaces), axis = face >> 1, side = face & 1;
    #pragma unroll
    for (int k = 0; k < 3; k++) {
        if (k < nd.d_in) {
            float u = u01(seed, i * (uint64_t)nd.d_in + (uint64_t)k);
            float v = fmaf(nd.span[k], u, nd.lo[k]);
            if (set == PINN_SET_BOUNDARY && k == axis) v = side ? nd.hi[k] : nd.lo[k];
            if (set == PINN_SET_INITIAL && nd.has_t && k == nd.d_in - 1) v = nd.lo[k];
            out[k] = v;
        }
    }
}

// Dirichlet / initial targets g_o(x).
__device__ __forceinline__ void pde_target(const NetDims& nd, int set, const float* x, float* g) {
    switch (nd.pde) {
    case PINN_PDE_BURGERS:
        g[0] = (set == PINN_SET_INITIAL) ? -sinf(kPi * x[0]) : 0.0f; break;
    case PINN_PDE_HELMHOLTZ:
        g[0] = 0.0f; break;
    case PINN_PDE_ALLEN_CAHN:
        g[0] = x[0] * x[0] * cosf(kPi * x[0]) * (x[1] * x[1] * cosf(kPi * x[1])); break;
    case PINN_PDE_HEAT:
        g[0] = (set == PINN_SET_INITIAL) ? sinf(kPi * x[0]) * sinf(kPi * x[1]) : 0.0f; break;
    case PINN_PDE_NAVIER_STOKES: {
        float e2 = expf(-2.0f * kNsNu * x[2]), e4 = expf(-4.0f * kNsNu * x[2]);
        g[0] = -cosf(x[0]) * sinf(x[1]) * e2;
        g[1] = sinf(x[0]) * cosf(x[1]) * e2;
        g[2] = -0.25f * (cosf(2.0f * x[0]) + cosf(2.0f * x[1])) * e4;
        break; }
    default: g[0] = 0.0f;
    }
}

// Residuals r_q and seeds ybar[o][c] = scale * sum_q r_q * d r_q / d y[o][c].
// y / ybar are [3][C] register arrays (only [d_out][C] used).  Returns sum_q r_q^2.
template <int C>
__device__ __forceinline__ float pde_residual(const NetDims& nd, const float* x, const float (&y)[3][C],
                                              float scale, float (&ybar)[3][C], float* r_out) {
    #pragma unroll
    for (int o = 0; o < 3; o++)
        #pragma unroll
        for (int c = 0; c < C; c++) ybar[o][c] = 0.0f;
    float ss = 0.0f;
    if constexpr (C == 4) {             // Burgers (x,t): u, u_x, u_t, u_xx
        float u = y[0][0], ux = y[0][1], ut = y[0][2], uxx = y[0][3];
        float r = ut + u * ux - kBurgersNu * uxx;
        float s = scale * r;
        ybar[0][0] = s * ux; ybar[0][1] = s * u; ybar[0][2] = s; ybar[0][3] = -(s * kBurgersNu);
        r_out[0] = r; ss = r * r;
    } else if constexpr (C == 5) {      // Helmholtz (x,y): u, u_x, u_y, u_xx, u_yy
        float q = (nd.k2 - kPi * kPi - 16.0f * kPi * kPi) * (sinf(kPi * x[0]) * sinf(4.0f * kPi * x[1]));
        float r = y[0][3] + y[0][4] + nd.k2 * y[0][0] - q;
        float s = scale * r;
        ybar[0][0] = s * nd.k2; ybar[0][3] = s; ybar[0][4] = s;
        r_out[0] = r; ss = r * r;
    } else if constexpr (C == 6) {      // (x,y,t): val, x, y, t, xx, yy
        if (nd.pde == PINN_PDE_ALLEN_CAHN) {
            float u = y[0][0];
            float r = y[0][3] - kAcGamma * (y[0][4] + y[0][5]) + 5.0f * u * u * u - 5.0f * u;
            float s = scale * r;
            ybar[0][0] = s * (15.0f * u * u - 5.0f); ybar[0][3] = s;
            ybar[0][4] = -(s * kAcGamma); ybar[0][5] = -(s * kAcGamma);
            r_out[0] = r; ss = r * r;
        } else if (nd.pde == PINN_PDE_HEAT) {
            float r = y[0][3] - kHeatAlpha * (y[0][4] + y[0][5]);
            float s = scale * r;
            ybar[0][3] = s; ybar[0][4] = -(s * kHeatAlpha); ybar[0][5] = -(s * kHeatAlpha);
            r_out[0] = r; ss = r * r;
        } else {                        // Navier-Stokes: outputs u, v, p
            const float (&U)[C] = y[0]; const float (&V)[C] = y[1]; const float (&Pp)[C] = y[2];
            float r0 = U[3] + U[0] * U[1] + V[0] * U[2] + Pp[1] - kNsNu * (U[4] + U[5]);
            float r1 = V[3] + U[0] * V[1] + V[0] * V[2] + Pp[2] - kNsNu * (V[4] + V[5]);
            float r2 = U[1] + V[2];
            float s0 = scale * r0, s1 = scale * r1, s2 = scale * r2;
            ybar[0][0] = s0 * U[1] + s1 * V[1];
            ybar[0][1] = s0 * U[0] + s2;
            ybar[0][2] = s0 * V[0];
            ybar[0][3] = s0;
            ybar[0][4] = -(s0 * kNsNu); ybar[0][5] = -(s0 * kNsNu);
            ybar[1][0] = s0 * U[2] + s1 * V[2];
            ybar[1][1] = s1 * U[0];
            ybar[1][2] = s1 * V[0] + s2;
            ybar[1][3] = s1;
            ybar[1][4] = -(s1 * kNsNu); ybar[1][5] = -(s1 * kNsNu);
            ybar[2][1] = s0; ybar[2][2] = s1;
            r_out[0] = r0; r_out[1] = r1; r_out[2] = r2;
            ss = r0 * r0 + r1 * r1 + r2 * r2;
        }
    }
    return ss;
}

// Sum over the `W` consecutive lanes that hold one point tile (W = 16 or 32).
// Only the lanes of that group need to be converged (a half-warp may call it alone when W = 16).
template <int W>
__device__ __forceinline__ float lane_group_sum(float v) {
    const unsigned mask = (W == 32) ? 0xffffffffu : (0xffffu << (16 * ((threadIdx.x & 31) >> 4)));
    #pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}

}  // namespace pinn
