// mma_sync_peak.cu — throughput of the warp-level mma.sync path on B200 (SASS HMMA), the candidate engine
// for the narrow-net (H = 20) contractions where operands can stay in registers from layer to layer.
//   m16n8k16 BF16 (FP32 accumulate) and m16n8k8 TF32, NACC independent accumulator tiles per warp,
//   W warps per SM.  Prints cycles per MMA per SM sub-partition and dense TFLOP/s.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int NACC, bool TF32>
__global__ void mma_loop(float* out, int iters) {
    float c[NACC][4];
    #pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f; }
    uint32_t a0 = threadIdx.x * 0x3c003c00u, a1 = a0 ^ 0x01010101u, a2 = a0 + 7u, a3 = a1 + 9u, b0 = a0 ^ 0x00100010u, b1 = b0 + 3u;
    for (int it = 0; it < iters; it++) {
        #pragma unroll
        for (int i = 0; i < NACC; i++) {
            if constexpr (TF32)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    float s = 0.f;
    #pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456f) out[threadIdx.x] = s;
}

template <int NACC, bool TF32>
static void run(int warps_per_sm, int sms, double ghz) {
    const int iters = 4096;
    float* d; cudaMalloc(&d, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        mma_loop<NACC, TF32><<<sms, warps_per_sm * 32>>>(d, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    const double mmas_per_smsp = (double)iters * NACC * warps_per_sm / 4.0;
    const double clk = best * 1e-3 * ghz * 1e9;
    const double flop = 2.0 * 16 * 8 * (TF32 ? 8 : 16) * (double)iters * NACC * warps_per_sm * sms;
    printf("%s nacc=%d warps/SM=%2d: %.2f clk per MMA per SMSP, %.1f TFLOP/s dense\n", TF32 ? "tf32 m16n8k8 " : "bf16 m16n8k16", NACC,
           warps_per_sm, clk / mmas_per_smsp, flop / (best * 1e-3) / 1e12);
    cudaFree(d);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    printf("%s, %d SMs, %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
    for (int w : {4, 8, 12, 16}) { run<4, false>(w, p.multiProcessorCount, ghz); run<8, false>(w, p.multiProcessorCount, ghz); }
    for (int w : {4, 8, 12, 16}) { run<4, true>(w, p.multiProcessorCount, ghz); run<8, true>(w, p.multiProcessorCount, ghz); }
    return 0;
}
