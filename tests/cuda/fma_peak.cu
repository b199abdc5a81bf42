// fma_peak.cu — measures the attainable FP32 FFMA rate of the CUDA cores on this GPU (the roofline
// denominator of the FP32 path; BASELINE.md section 2 asks for this confirmation of the derived
// 148 x 128 x 2 x clock figure).  Register-only kernels with NACC independent accumulators per thread:
//   variant A: acc[i] = fma(acc[i], x, y)        (2 shared operands: best case for the reuse cache)
//   variant B: acc[i] = fma(a[i%4], w[i%8], acc[i])  (the GEMM-like pattern of the fused kernels)
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC, int VAR>
__global__ void __launch_bounds__(256) fma_kernel(float* out, int iters, float x, float y) {
    float acc[NACC], a[4], w[8];
    for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x * 1e-3f + i;
    for (int i = 0; i < 4; i++) a[i] = x + i * 1e-3f;
    for (int i = 0; i < 8; i++) w[i] = y + i * 1e-3f;
    for (int it = 0; it < iters; it++) {
        #pragma unroll
        for (int i = 0; i < NACC; i++) {
            if (VAR == 0) acc[i] = fmaf(acc[i], x, y);
            else acc[i] = fmaf(a[i & 3], w[(i >> 2) & 7], acc[i]);
        }
        if (VAR == 1) { a[it & 3] += 1e-7f; w[it & 7] -= 1e-7f; }
    }
    float s = 0;
    for (int i = 0; i < NACC; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC, int VAR>
void run(const char* name, int blocks_per_sm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * blocks_per_sm, iters = 4096;
    float* out; cudaMalloc(&out, (size_t)blocks * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    fma_kernel<NACC, VAR><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    float best = 1e9f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        fma_kernel<NACC, VAR><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    const double flops = 2.0 * blocks * 256.0 * iters * NACC;
    printf("%-34s warps/SM %2d  %.3f ms  %.1f TFLOP/s\n", name, blocks_per_sm * 8, best, flops / best / 1e9);
    cudaFree(out);
}

int main() {
    run<32, 0>("A: fma(acc,x,y)      32 acc", 8);
    run<32, 0>("A: fma(acc,x,y)      32 acc", 2);
    run<32, 0>("A: fma(acc,x,y)      32 acc", 1);
    run<32, 1>("B: fma(a[c],w[j],acc) 32 acc", 8);
    run<32, 1>("B: fma(a[c],w[j],acc) 32 acc", 2);
    run<32, 1>("B: fma(a[c],w[j],acc) 32 acc", 1);
    run<80, 1>("B: fma(a[c],w[j],acc) 80 acc", 1);
    return 0;
}
