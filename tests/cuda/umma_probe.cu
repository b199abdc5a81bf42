// umma_probe.cu — standalone check of the tcgen05 plumbing in cuda-pinn_b200/csrc/pinn_umma.cuh:
// the three contractions of one PINN layer (Z = A W, dA = Zbar W^T, dW = A^T Zbar) on a 128-row
// tile, operands in the chunk-major no-swizzle layout, against a CPU float reference computed
// from the same bf16-rounded inputs.  Usage: umma_probe   (prints PASS/FAIL per case)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../cuda-pinn_b200/csrc/pinn_umma.cuh"

using namespace pinn::umma;

static inline float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

// A, Z: [128, H] fp32 row-major; W: [H, H] row-major (in x out).  D1 = A W, D2 = Z W^T, D3 = A^T Z.
__global__ void __launch_bounds__(128) probe(const float* A, const float* Z, const float* W, int H,
                                             float* D1, float* D2, float* D3) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int R = 128;
    const int HW = H > 128 ? 128 : H;                   // W is staged in j-halves of at most 128 columns
    uint8_t* As = smem;                                  // [H/8][128][16B]
    uint8_t* Zs = As + (size_t)H * 256;
    uint8_t* Ws = Zs + (size_t)H * 256;                  // [HW/8][H][16B]
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < R * H; i += blockDim.x) {
        const int r = i / H, f = i % H;
        *reinterpret_cast<__nv_bfloat16*>(As + cm_offset(r, f, R)) = __float2bfloat16(A[i]);
        *reinterpret_cast<__nv_bfloat16*>(Zs + cm_offset(r, f, R)) = __float2bfloat16(Z[i]);
    }
    auto stage_w = [&](int half) {                       // columns j in [half*HW, half*HW+HW)
        for (int i = tid; i < H * HW; i += blockDim.x) {
            const int k = i / HW, jj = i % HW;
            *reinterpret_cast<__nv_bfloat16*>(Ws + cm_offset(k, jj, H)) = __float2bfloat16(W[k * H + half * HW + jj]);
        }
    };
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    uint32_t phase = 0;
    const uint32_t a_addr = smem_u32(As), z_addr = smem_u32(Zs), w_addr = smem_u32(Ws);
    const int nhalf = H / HW;

    auto drain = [&](float* D, int ld, int rows_valid, int col0, int ncols, uint32_t tcol0) {
        tc_fence_after();
        const int row = warp * 32 + lane;
        for (int c = 0; c < ncols; c += 8) {
            float v[8];
            tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + tcol0 + c, v);
            if (row < rows_valid)
                for (int i = 0; i < 8; i++) D[(size_t)row * ld + col0 + c + i] = v[i];
        }
        tc_fence_before();
        __syncthreads();
    };

    // ---- case 1: D1[128,H] = A (K-major over f) x W (N = j, MN-major B), one j-half at a time ----
    for (int h = 0; h < nhalf; h++) {
        stage_w(h);
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            gemm_issue(tmem + h * HW, a_addr, R * 16, 128, 2 * R * 16, w_addr, 128, H * 16, 256,
                       idesc_bf16(128, HW, 0, 1), H / 16, 0);
            mma_commit(&bar);
        }
        mbar_wait(&bar, phase); phase ^= 1;
    }
    drain(D1, H, 128, 0, H, 0);

    // ---- case 2: D2[128,H] = Z (K-major over j) x W^T (N = k, K-major B), accumulating over j-halves ----
    for (int h = 0; h < nhalf; h++) {
        stage_w(h);
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            // A operand: the j-half h of Z starts h*HW/8 chunks in
            gemm_issue(tmem, z_addr + h * (HW / 8) * R * 16, R * 16, 128, 2 * R * 16, w_addr, H * 16, 128, 2 * H * 16,
                       idesc_bf16(128, H, 0, 0), HW / 16, h > 0);
            mma_commit(&bar);
        }
        mbar_wait(&bar, phase); phase ^= 1;
    }
    drain(D2, H, 128, 0, H, 0);

    // ---- case 3: D3[H,H] = A^T (M = f, MN-major A) x Z (N = f, MN-major B), contraction over the 128 rows ----
    const int mt = (H + 127) / 128;
    if (tid == 0) {
        tc_fence_after();
        for (int t = 0; t < mt; t++)
            gemm_issue(tmem + t * H, a_addr + t * 16 * R * 16, 128, R * 16, 256, z_addr, 128, R * 16, 256,
                       idesc_bf16(128, H, 1, 1), R / 16, 0);
        mma_commit(&bar);
    }
    mbar_wait(&bar, phase); phase ^= 1;
    for (int t = 0; t < mt; t++) drain(D3 + (size_t)t * 128 * H, H, H - t * 128 < 128 ? H - t * 128 : 128, 0, H, t * H);

    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    int fails = 0;
    for (int H : {64, 128, 256}) {
        const int R = 128;
        std::vector<float> A(R * H), Z(R * H), W(H * H);
        srand(1234 + H);
        auto rnd = [] { return (float)rand() / RAND_MAX * 2.0f - 1.0f; };
        for (auto& v : A) v = bf16_round(rnd());
        for (auto& v : Z) v = bf16_round(rnd());
        for (auto& v : W) v = bf16_round(rnd() * 0.25f);
        float *dA, *dZ, *dW, *d1, *d2, *d3;
        cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dZ, Z.size() * 4); cudaMalloc(&dW, W.size() * 4);
        cudaMalloc(&d1, R * H * 4); cudaMalloc(&d2, R * H * 4); cudaMalloc(&d3, (size_t)256 * H * 4);
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dZ, Z.data(), Z.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
        const int HW = H > 128 ? 128 : H;
        size_t smem = (size_t)H * 256 * 2 + (size_t)HW * H * 2 + 4096;
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        probe<<<1, 128, smem>>>(dA, dZ, dW, H, d1, d2, d3);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("H=%d CUDA error: %s\n", H, cudaGetErrorString(e)); return 2; }
        std::vector<float> D1(R * H), D2(R * H), D3(H * H);
        cudaMemcpy(D1.data(), d1, D1.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(D2.data(), d2, D2.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(D3.data(), d3, D3.size() * 4, cudaMemcpyDeviceToHost);
        double e1 = 0, e2 = 0, e3 = 0, m1 = 0, m2 = 0, m3 = 0;
        for (int r = 0; r < R; r++)
            for (int j = 0; j < H; j++) {
                double s1 = 0, s2 = 0;
                for (int k = 0; k < H; k++) { s1 += (double)A[r * H + k] * W[k * H + j]; s2 += (double)Z[r * H + k] * W[j * H + k]; }
                e1 = fmax(e1, fabs(s1 - D1[r * H + j])); m1 = fmax(m1, fabs(s1));
                e2 = fmax(e2, fabs(s2 - D2[r * H + j])); m2 = fmax(m2, fabs(s2));
            }
        for (int k = 0; k < H; k++)
            for (int j = 0; j < H; j++) {
                double s = 0;
                for (int r = 0; r < R; r++) s += (double)A[r * H + k] * Z[r * H + j];
                e3 = fmax(e3, fabs(s - D3[k * H + j])); m3 = fmax(m3, fabs(s));
            }
        const bool ok = e1 < 1e-3 * m1 && e2 < 1e-3 * m2 && e3 < 1e-3 * m3;
        printf("H=%3d  fwd err %.2e/%.2f  dA err %.2e/%.2f  dW err %.2e/%.2f  %s\n", H, e1, m1, e2, m2, e3, m3, ok ? "PASS" : "FAIL");
        fails += !ok;
        cudaFree(dA); cudaFree(dZ); cudaFree(dW); cudaFree(d1); cudaFree(d2); cudaFree(d3);
    }
    return fails ? 1 : 0;
}
