// fma_peak2.cu — FFMA rate of the exact register-tile patterns of the fused kernels (NC x NU accumulators,
// z[c][j] += a[c] * w[j]) for both loop orders, with loop-invariant operands.
#include <cstdio>
#include <cuda_runtime.h>

template <int NC, int NU, int ORDER>
__global__ void __launch_bounds__(256) tile_kernel(float* out, int iters, float x, float y) {
    float z[NC][NU], a[NC], w[NU];
    for (int c = 0; c < NC; c++) { a[c] = x + c * 1e-3f; for (int j = 0; j < NU; j++) z[c][j] = threadIdx.x * 1e-3f + j; }
    for (int j = 0; j < NU; j++) w[j] = y + j * 1e-3f;
    for (int it = 0; it < iters; it++) {
        if (ORDER == 0) {
            #pragma unroll
            for (int c = 0; c < NC; c++)
                #pragma unroll
                for (int j = 0; j < NU; j++) z[c][j] = fmaf(a[c], w[j], z[c][j]);
        } else {
            #pragma unroll
            for (int j = 0; j < NU; j++)
                #pragma unroll
                for (int c = 0; c < NC; c++) z[c][j] = fmaf(a[c], w[j], z[c][j]);
        }
        // perturb operands so nothing is hoisted (NC + NU cheap ops per NC*NU FFMAs)
        #pragma unroll
        for (int c = 0; c < NC; c++) a[c] = a[c] * 0.999999f;
        #pragma unroll
        for (int j = 0; j < NU; j++) w[j] = w[j] * 1.000001f;
    }
    float s = 0;
    for (int c = 0; c < NC; c++) for (int j = 0; j < NU; j++) s += z[c][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NC, int NU, int ORDER>
void run(const char* name, int blocks_per_sm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * blocks_per_sm, iters = 2048;
    float* out; cudaMalloc(&out, (size_t)blocks * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    tile_kernel<NC, NU, ORDER><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    float best = 1e9f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        tile_kernel<NC, NU, ORDER><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    const double flops = 2.0 * blocks * 256.0 * iters * (NC * NU);   // FFMA only (the NC+NU FMULs are overhead)
    printf("%-40s warps/SM %2d  %.3f ms  %.1f TFLOP/s (FFMA only)\n", name, blocks_per_sm * 8, best, flops / best / 1e9);
    cudaFree(out);
}

int main() {
    run<4, 4, 0>("4x4   c-outer (block kernel gemm)", 4);
    run<4, 4, 1>("4x4   j-outer", 4);
    run<4, 20, 0>("4x20  c-outer (warp kernel contract)", 1);
    run<4, 20, 1>("4x20  j-outer", 1);
    run<4, 20, 0>("4x20  c-outer", 2);
    run<6, 8, 0>("6x8   c-outer (wide nets)", 2);
    run<6, 8, 1>("6x8   j-outer", 2);
    run<8, 8, 0>("8x8   c-outer", 2);
    return 0;
}
