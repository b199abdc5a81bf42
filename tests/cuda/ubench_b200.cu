// ubench_b200.cu — design microbenchmarks for the round-2 tensor-core pipeline (measured numbers go to
// profiles/r02_ubench.txt).  One process, one GPU; every figure is CUDA-event time over the whole grid.
//   l2read   : aggregate cp.async.bulk global->shared bandwidth out of an L2-resident working set
//   l2write  : aggregate cp.async.bulk shared->global bandwidth into an L2-resident working set
//   red      : red.global.add.v4.f32 / cp.reduce.async.bulk add.f32 into a 1.8 MB (L2-resident) buffer
//   mma      : tcgen05.mma SS-mode cycles per K=16 dispatch at M=128, N in {64..256}, with/without smem store noise
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../cuda-pinn_b200/csrc/pinn_umma.cuh"

using namespace pinn::umma;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// ---------------------------------------------------------------- L2 read (bulk g2s) -------------
__global__ void __launch_bounds__(128) l2read_kernel(const uint8_t* src, size_t per_cta, int chunk, int reps, int shared_src) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar[4];
    if (threadIdx.x == 0) { for (int i = 0; i < 4; i++) mbar_init(&bar[i], 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const uint8_t* base = src + (shared_src ? 0 : (size_t)blockIdx.x * per_cta);
    const int nchunk = (int)(per_cta / chunk);
    const long total = (long)reps * nchunk;
    uint32_t ph[4] = {0, 0, 0, 0};
    for (long i = 0; i < total + 4; i++) {
        const int s = (int)(i & 3);
        if (i >= 4) { mbar_wait(&bar[s], ph[s]); ph[s] ^= 1; }
        if (i < total) {
            mbar_expect_tx(&bar[s], chunk);
            bulk_g2s(smem + (size_t)s * chunk, base + (size_t)(i % nchunk) * chunk, chunk, &bar[s]);
        }
    }
}

// ---------------------------------------------------------------- L2 write (bulk s2g) ------------
__global__ void __launch_bounds__(128) l2write_kernel(uint8_t* dst, size_t per_cta, int chunk, int reps) {
    extern __shared__ __align__(128) uint8_t smem[];
    for (int i = threadIdx.x; i < chunk / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i;
    fence_async_smem();
    __syncthreads();
    if (threadIdx.x != 0) return;
    uint8_t* base = dst + (size_t)blockIdx.x * per_cta;
    const int nchunk = (int)(per_cta / chunk);
    for (long i = 0; i < (long)reps * nchunk; i++) {
        bulk_s2g(base + (size_t)(i % nchunk) * chunk, smem, chunk);
        bulk_commit();
        asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
    }
    bulk_wait_all();
}

// ---------------------------------------------------------------- RED into a shared buffer -------
// mode 0: red.global.add.f32 (scalar)  1: red.global.add.v4.f32  2: cp.reduce.async.bulk add.f32 (16 KB chunks)
__global__ void __launch_bounds__(256) red_kernel(float* acc, int nlayers, int layer_floats, int reps, int mode) {
    extern __shared__ __align__(128) uint8_t smem[];
    float* sf = reinterpret_cast<float*>(smem);
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sf[i] = 1.0f;
    fence_async_smem();
    __syncthreads();
    for (int r = 0; r < reps; r++) {
        float* dst = acc + (size_t)((blockIdx.x + r) % nlayers) * layer_floats;
        if (mode == 0) {
            for (int i = threadIdx.x; i < layer_floats; i += blockDim.x)
                asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + i), "f"(1.0f) : "memory");
        } else if (mode == 1) {
            for (int i = threadIdx.x * 4; i < layer_floats; i += blockDim.x * 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + i), "f"(1.0f), "f"(1.0f), "f"(1.0f), "f"(1.0f) : "memory");
        } else {
            if (threadIdx.x == 0) {
                for (int i = 0; i < layer_floats; i += 4096) {
                    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                                 ::"l"(dst + i), "r"(smem_u32(sf)), "r"(16384) : "memory");
                    bulk_commit();
                    asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
                }
            }
        }
    }
    if (mode == 2 && threadIdx.x == 0) bulk_wait_all();
}

// ---------------------------------------------------------------- tcgen05.mma rate ---------------
// layout 0: formulation "U" forward (A = W^T MN-major chunk-major, B = activations MN-major)
// layout 1: A K-major (dA: W rows), B MN-major
// layout 2: dW (A K-major over n, B K-major over n)
__global__ void __launch_bounds__(256) mma_kernel(int N, int layout, int iters, int noise, long long* cycles_out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint8_t* As = smem;                 // 64 KB: [128 x 256] bf16
    uint8_t* Bs = smem + 65536;         // up to 128 KB: [256 x N]
    uint8_t* Ns = smem + 65536 + 131072;// 16 KB noise area
    for (int i = tid; i < (65536 + 131072) / 4; i += blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u;
        reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u ^ ((h >> 7) & 0x007f007fu) ^ ((h & 1u) << 31) ^ ((h & 2u) << 14);
    }
    if (warp == 0) tmem_alloc(&slot, 512);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); *reinterpret_cast<volatile int*>(Ns + 16380) = 0; }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    const uint32_t a = smem_u32(As), b = smem_u32(Bs);
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        t0 = clock64();
        uint32_t ph = 0;
        for (int it = 0; it < iters; it++) {
            const uint32_t d = tmem + (uint32_t)((it & 1) * 256);
            if (layout == 0)       // A MN-major: SBO = 4096 (next 8 m), LBO = 128 (next 8 k), kstep 256;  B MN-major same form
                gemm_issue(d, a, 128, 4096, 256, b, 128, 4096, 256, idesc_bf16(128, N, 1, 1), 16, 0);
            else if (layout == 1)  // A K-major: LBO = 2048 (next k chunk; 128 rows*16), SBO = 128, kstep 4096
                gemm_issue(d, a, 2048, 128, 4096, b, 128, 4096, 256, idesc_bf16(128, N, 0, 1), 16, 0);
            else                   // both K-major over 128 "n": A [128 m][K], B [N][K]; K = 128 -> 8 ksteps, issue twice
                for (int h = 0; h < 2; h++)
                    gemm_issue(d, a, 2048, 128, 4096, b, (uint32_t)N * 16, 128, (uint32_t)N * 32, idesc_bf16(128, N, 0, 0), 8, h);
            if ((it & 3) == 3 || it == iters - 1) { mma_commit(&bar); mbar_wait(&bar, ph); ph ^= 1; }
        }
        t1 = clock64();
        *reinterpret_cast<volatile int*>(Ns + 16380) = 1;   // stop flag for the noise warps
    } else if (noise && warp >= 4) {
        // shared-memory store noise: 4 warps write 16-byte chunks continuously while the MMAs run
        volatile int* stop = reinterpret_cast<volatile int*>(Ns + 16380);
        uint4 v = make_uint4(tid, tid, tid, tid);
        int k = 0;
        while (!*stop) {
            #pragma unroll
            for (int u = 0; u < noise; u++)
                *reinterpret_cast<uint4*>(Ns + ((tid - 128) * 16 + ((k + u) & 3) * 2048)) = v;
            k++;
            __nanosleep(0);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) cycles_out[blockIdx.x] = t1 - t0;
    if (warp == 0) tmem_dealloc(tmem, 512);
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main(int argc, char** argv) {
    int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
    int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    printf("SMs %d, max clock %d kHz\n", nsm, clk);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    uint8_t* buf; const size_t BUF = (size_t)512 << 20; CK(cudaMalloc(&buf, BUF)); CK(cudaMemset(buf, 1, BUF));

    // ---- L2 read ----
    CK(cudaFuncSetAttribute(l2read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768));
    for (int shared_src = 0; shared_src < 2; shared_src++)
        for (size_t per : {(size_t)128 << 10, (size_t)512 << 10, (size_t)2 << 20}) {
            for (int chunk : {8192, 32768}) {
                const int reps = (int)(((size_t)64 << 20) / per);
                l2read_kernel<<<nsm, 128, 4 * chunk>>>(buf, per, chunk, 2, shared_src);
                CK(cudaEventRecord(e0));
                l2read_kernel<<<nsm, 128, 4 * chunk>>>(buf, per, chunk, reps, shared_src);
                CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
                const double bytes = (double)nsm * per * reps;
                printf("l2read  %s per_cta %7zu KB (set %6.1f MB) chunk %5d : %8.1f GB/s\n", shared_src ? "same-src" : "own-src ",
                       per >> 10, shared_src ? per / 1048576.0 : nsm * per / 1048576.0, chunk, bytes / time_ms(e0, e1) * 1e-6);
            }
        }
    // ---- L2 write ----
    CK(cudaFuncSetAttribute(l2write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    for (size_t per : {(size_t)128 << 10, (size_t)512 << 10, (size_t)2 << 20}) {
        const int reps = (int)(((size_t)64 << 20) / per);
        l2write_kernel<<<nsm, 128, 32768>>>(buf, per, 32768, 2);
        CK(cudaEventRecord(e0));
        l2write_kernel<<<nsm, 128, 32768>>>(buf, per, 32768, reps);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        printf("l2write per_cta %7zu KB (set %6.1f MB) : %8.1f GB/s\n", per >> 10, nsm * per / 1048576.0,
               (double)nsm * per * reps / time_ms(e0, e1) * 1e-6);
    }
    // ---- RED ----
    {
        float* acc; const int nl = 7, lf = 65536; CK(cudaMalloc(&acc, (size_t)nl * lf * 4)); CK(cudaMemset(acc, 0, (size_t)nl * lf * 4));
        for (int mode = 0; mode < 3; mode++) {
            const int reps = 20;
            red_kernel<<<nsm, 256, 16384>>>(acc, nl, lf, 2, mode);
            CK(cudaEventRecord(e0));
            red_kernel<<<nsm, 256, 16384>>>(acc, nl, lf, reps, mode);
            CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
            const double bytes = (double)nsm * reps * lf * 4;
            const float ms = time_ms(e0, e1);
            printf("red mode %d (%s): %8.1f GB/s aggregate, %.1f us per 256 KB flush per CTA\n", mode,
                   mode == 0 ? "red.f32" : mode == 1 ? "red.v4.f32" : "cp.reduce.async.bulk", bytes / ms * 1e-6, ms * 1e3 / reps);
        }
        std::vector<float> h(8); CK(cudaMemcpy(h.data(), acc, 32, cudaMemcpyDeviceToHost));
        printf("red check: acc[0] = %.1f (expected %d adds per element across modes)\n", h[0], 0);
        CK(cudaFree(acc));
    }
    // ---- MMA ----
    {
        long long* cyc; CK(cudaMalloc(&cyc, nsm * 8));
        const int smem_bytes = 65536 + 131072 + 16384;
        CK(cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        std::vector<long long> h(nsm);
        for (int layout = 0; layout < 3; layout++)
            for (int N : {64, 96, 128, 192, 256})
                for (int noise : {0, 4}) {
                    const int iters = 400;
                    mma_kernel<<<nsm, 256, smem_bytes>>>(N, layout, 20, noise, cyc);
                    CK(cudaEventRecord(e0));
                    mma_kernel<<<nsm, 256, smem_bytes>>>(N, layout, iters, noise, cyc);
                    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
                    CK(cudaMemcpy(h.data(), cyc, nsm * 8, cudaMemcpyDeviceToHost));
                    double avg = 0; for (auto c : h) avg += (double)c; avg /= nsm;
                    const float ms = time_ms(e0, e1);
                    const double flops = 2.0 * 128 * N * 256 * iters * nsm;
                    printf("mma layout %d N %3d noise %d: %7.1f cycles/dispatch (ideal %5.1f), %7.1f TFLOP/s, kernel %.3f ms\n", layout, N,
                           noise, avg / (iters * 16.0), 128.0 * N / 256.0, flops / ms * 1e-9, ms);
                }
        CK(cudaFree(cyc));
    }
    printf("done\n");
    return 0;
}
