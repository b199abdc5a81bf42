// pinn_tensor_ops.cu — B200-native device arithmetic of the reference's tensor<T> (include/pinn_tensor_abi.h): elementwise
// add / sub / mul / div with the reference's flat-index broadcast, row-major matmul, dot, variadic add; device-resident,
// stream-ordered.  These are HBM-bound (elementwise, dot: 8-12 B per element) or small FP32 contractions whose results must
// equal the reference's bit for bit, so: coalesced 16-byte accesses, grids sized to the SM count, a shared-memory tiled
// CUDA-core matmul that keeps the reference's per-element operation order (one accumulator, k ascending, fused
// multiply-add — kernel.cuh:77-80 as nvcc compiles it) — no tensor cores, which would change the rounding.
// Replaces: include/device/kernel.cuh:10-99 and the host sequences around them in include/tensor.cuh:84-140, 347-508.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include "../../include/pinn_tensor_abi.h"

namespace {

int g_sms = 0;
int sm_count() {
    if (g_sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) g_sms = 148;
    }
    return g_sms;
}

template <typename T, int OP>
__device__ __forceinline__ T apply(T a, T b) {
    if (OP == PINN_OP_ADD) return a + b;
    if (OP == PINN_OP_SUB) return a - b;
    if (OP == PINN_OP_MUL) return a * b;
    return a / b;
}

// same-size operands (the common case): 16-byte accesses, grid-stride
template <typename T, int OP>
__global__ void elementwise_same(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ r, size_t n, int* __restrict__ flag) {
    constexpr int V = 16 / sizeof(T);
    struct alignas(16) Vec { T v[V]; };
    const size_t nv = n / V, stride = (size_t)gridDim.x * blockDim.x;
    bool zero = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        const Vec x = reinterpret_cast<const Vec*>(a)[i], y = reinterpret_cast<const Vec*>(b)[i];
        Vec z;
        if (OP == PINN_OP_DIV) z = reinterpret_cast<Vec*>(r)[i];               // (div only: a zero denominator leaves the element as it was)
        #pragma unroll
        for (int k = 0; k < V; k++) {
            if (OP == PINN_OP_DIV && y.v[k] == T(0)) zero = true;
            else z.v[k] = apply<T, OP>(x.v[k], y.v[k]);
        }
        reinterpret_cast<Vec*>(r)[i] = z;
    }
    for (size_t i = nv * V + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (OP == PINN_OP_DIV && b[i] == T(0)) zero = true;
        else r[i] = apply<T, OP>(a[i], b[i]);
    }
    if (OP == PINN_OP_DIV && zero) atomicExch(flag, 1);
}

// the reference's broadcast: flat index modulo the operand size (kernel.cuh:16)
template <typename T, int OP>
__global__ void elementwise_bcast(const T* __restrict__ a, size_t na, const T* __restrict__ b, size_t nb, T* __restrict__ r, size_t bound,
                                  int* __restrict__ flag) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    bool zero = false;
    const bool small = na < (1ull << 32) && nb < (1ull << 32) && bound < (1ull << 32);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < bound; i += stride) {
        const size_t ia = small ? (size_t)((uint32_t)i % (uint32_t)na) : i % na, ib = small ? (size_t)((uint32_t)i % (uint32_t)nb) : i % nb;
        const T y = b[ib];
        if (OP == PINN_OP_DIV && y == T(0)) zero = true;
        else r[i] = apply<T, OP>(a[ia], y);
    }
    if (OP == PINN_OP_DIV && zero) atomicExch(flag, 1);
}

// the common broadcast — one operand has the result's size, the other is small (a bias): 16-byte accesses on the big one, the small
// one through the cache.  AFULL: the first operand is the full-size one.
template <typename T, int OP, bool AFULL>
__global__ void elementwise_bcast_vec(const T* __restrict__ a, size_t na, const T* __restrict__ b, size_t nb, T* __restrict__ r, size_t bound,
                                      int* __restrict__ flag) {
    constexpr int V = 16 / sizeof(T);
    struct alignas(16) Vec { T v[V]; };
    const size_t nv = bound / V, stride = (size_t)gridDim.x * blockDim.x;
    const uint32_t ns = (uint32_t)(AFULL ? nb : na);                            // (small operand: < 2^32 elements, checked by the caller)
    bool zero = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        const Vec big = reinterpret_cast<const Vec*>(AFULL ? a : b)[i];
        Vec z;
        if (OP == PINN_OP_DIV) z = reinterpret_cast<Vec*>(r)[i];
        uint32_t m = (uint32_t)((i * V) % ns);
        #pragma unroll
        for (int k = 0; k < V; k++) {
            const T sm = (AFULL ? b : a)[m];
            m = m + 1 == ns ? 0 : m + 1;
            const T x = AFULL ? big.v[k] : sm, y = AFULL ? sm : big.v[k];
            if (OP == PINN_OP_DIV && y == T(0)) zero = true;
            else z.v[k] = apply<T, OP>(x, y);
        }
        reinterpret_cast<Vec*>(r)[i] = z;
    }
    for (size_t i = nv * V + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < bound; i += stride) {
        const T x = a[i % na], y = b[i % nb];
        if (OP == PINN_OP_DIV && y == T(0)) zero = true;
        else r[i] = apply<T, OP>(x, y);
    }
    if (OP == PINN_OP_DIV && zero) atomicExch(flag, 1);
}

// R[I, J] = A[I, K] B[K, J]: (16 Q) x (16 Q) output tile per CTA of 256 threads, Q x Q register tile per thread (Q = 4: 64 x 64, for
// narrow results and double; Q = 8: 128 x 128 as 2 x 2 blocks of 4 x 4 so that every shared-memory read is a conflict-free 16 bytes),
// K in slices of 16 through shared memory.  Every output keeps ONE accumulator and takes its products in ascending k with fma —
// the reference's order, whatever the tiling.
template <typename T, bool TA, bool TB, int Q>
__global__ void __launch_bounds__(256) matmul_tiled(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ R, size_t I, size_t K, size_t J) {
    constexpr int TM = 16 * Q, TN = 16 * Q, TK = 16, NB = Q / 4;                  // NB x NB blocks of 4 x 4 per thread, 64 apart
    __shared__ T As[TK][TM + 4];
    __shared__ T Bs[TK][TN + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const size_t i0 = (size_t)blockIdx.y * TM, j0 = (size_t)blockIdx.x * TN;
    T acc[Q][Q];
    #pragma unroll
    for (int a = 0; a < Q; a++)
        #pragma unroll
        for (int b = 0; b < Q; b++) acc[a][b] = T(0);
    for (size_t k0 = 0; k0 < K; k0 += TK) {
        const int kmax = (int)((K - k0) < (size_t)TK ? (K - k0) : (size_t)TK);
        // A slice -> As[k][i] (coalesced along the stored-contiguous index)
        #pragma unroll
        for (int e = threadIdx.x; e < TM * TK; e += 256) {
            int k, i;
            if (TA) { i = e % TM; k = e / TM; } else { k = e % TK; i = e / TK; }
            const size_t gi = i0 + i, gk = k0 + k;
            T v = T(0);
            if (gi < I && gk < K) v = TA ? A[gk * I + gi] : A[gi * K + gk];
            As[k][i] = v;
        }
        #pragma unroll
        for (int e = threadIdx.x; e < TN * TK; e += 256) {
            int k, j;
            if (TB) { k = e % TK; j = e / TK; } else { j = e % TN; k = e / TN; }
            const size_t gj = j0 + j, gk = k0 + k;
            T v = T(0);
            if (gj < J && gk < K) v = TB ? B[gj * K + gk] : B[gk * J + gj];
            Bs[k][j] = v;
        }
        __syncthreads();
        for (int k = 0; k < kmax; k++) {                                       // (exact bound: no padded products, the sign of a zero sum is the reference's)
            T a[Q], b[Q];
            #pragma unroll
            for (int nb = 0; nb < NB; nb++)
                #pragma unroll
                for (int u = 0; u < 4; u++) { a[nb * 4 + u] = As[k][nb * 64 + ty * 4 + u]; b[nb * 4 + u] = Bs[k][nb * 64 + tx * 4 + u]; }
            #pragma unroll
            for (int u = 0; u < Q; u++)
                #pragma unroll
                for (int v = 0; v < Q; v++) acc[u][v] = fma(a[u], b[v], acc[u][v]);
        }
        __syncthreads();
    }
    #pragma unroll
    for (int u = 0; u < Q; u++)
        #pragma unroll
        for (int v = 0; v < Q; v++) {
            const size_t gi = i0 + (u >> 2) * 64 + ty * 4 + (u & 3), gj = j0 + (v >> 2) * 64 + tx * 4 + (v & 3);
            if (gi < I && gj < J) R[gi * J + gj] = acc[u][v];
        }
}

// dot, stage 1: fixed assignment of elements to threads, shuffle tree, shared-memory tree -> one partial per CTA
template <typename T>
__global__ void __launch_bounds__(256) dot_partial(const T* __restrict__ a, const T* __restrict__ b, size_t n, T* __restrict__ partial) {
    __shared__ T warp_sum[8];
    constexpr int V = 16 / sizeof(T);
    struct alignas(16) Vec { T v[V]; };
    T acc[V];
    #pragma unroll
    for (int k = 0; k < V; k++) acc[k] = T(0);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const bool aligned = (((uintptr_t)a | (uintptr_t)b) & 15) == 0;
    const size_t nv = aligned ? n / V : 0;                                      // 16-byte loads, V independent accumulators per thread
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        const Vec x = reinterpret_cast<const Vec*>(a)[i], y = reinterpret_cast<const Vec*>(b)[i];
        #pragma unroll
        for (int k = 0; k < V; k++) acc[k] = fma(x.v[k], y.v[k], acc[k]);
    }
    for (size_t i = nv * V + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc[0] = fma(a[i], b[i], acc[0]);
    T s = acc[0];
    #pragma unroll
    for (int k = 1; k < V; k++) s += acc[k];                                    // fixed order
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        T t = T(0);
        #pragma unroll
        for (int w = 0; w < 8; w++) t += warp_sum[w];
        partial[blockIdx.x] = t;
    }
}
template <typename T>
__global__ void __launch_bounds__(256) dot_final(const T* __restrict__ partial, int count, T* __restrict__ r) {
    __shared__ T warp_sum[8];
    T s = T(0);
    for (int i = threadIdx.x; i < count; i += 256) s += partial[i];
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        T t = T(0);
        #pragma unroll
        for (int w = 0; w < 8; w++) t += warp_sum[w];
        r[0] = t;
    }
}

struct SrcList { const void* p[16]; };
template <typename T>
__global__ void add_multiple_kernel(const SrcList srcs, int count, size_t n, T* __restrict__ r) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        T sum = T(0);
        for (int k = 0; k < count; k++) sum += static_cast<const T*>(srcs.p[k])[i];     // (the reference's order, kernel.cuh:61-63)
        r[i] = sum;
    }
}

// NumPy broadcasting: per-dimension element strides of the operands (0 where the dimension stretches), row-major result
struct BcastDims { int dim; size_t shape[8]; size_t sa[8]; size_t sb[8]; };
template <typename T, int OP>
__global__ void broadcast_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ r, size_t n, const BcastDims bd) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        size_t rem = i, ia = 0, ib = 0;
        #pragma unroll 1
        for (int k = bd.dim - 1; k >= 0; k--) {
            const size_t c = rem % bd.shape[k]; rem /= bd.shape[k];
            ia += c * bd.sa[k]; ib += c * bd.sb[k];
        }
        r[i] = apply<T, OP>(a[ia], b[ib]);
    }
}

// per-column minimum / maximum of x[n, d], two fixed-order stages (d <= 8)
__global__ void __launch_bounds__(256) colminmax_partial(const float* __restrict__ x, size_t n, int d, float* __restrict__ part) {
    __shared__ float smn[8][8], smx[8][8];
    float mn[8], mx[8];
    for (int k = 0; k < 8; k++) { mn[k] = INFINITY; mx[k] = -INFINITY; }
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        for (int k = 0; k < d; k++) { const float v = x[i * d + k]; mn[k] = fminf(mn[k], v); mx[k] = fmaxf(mx[k], v); }
    for (int k = 0; k < d; k++) {
        for (int o = 16; o > 0; o >>= 1) { mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o)); mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o)); }
        if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5][k] = mn[k]; smx[threadIdx.x >> 5][k] = mx[k]; }
    }
    __syncthreads();
    if (threadIdx.x < d) {
        float a = INFINITY, b = -INFINITY;
        for (int w = 0; w < 8; w++) { a = fminf(a, smn[w][threadIdx.x]); b = fmaxf(b, smx[w][threadIdx.x]); }
        part[(size_t)blockIdx.x * 16 + threadIdx.x] = a; part[(size_t)blockIdx.x * 16 + 8 + threadIdx.x] = b;
    }
}
__global__ void colminmax_final(const float* __restrict__ part, int blocks, int d, float* __restrict__ out) {
    if (threadIdx.x < d) {
        float a = INFINITY, b = -INFINITY;
        for (int i = 0; i < blocks; i++) { a = fminf(a, part[(size_t)i * 16 + threadIdx.x]); b = fmaxf(b, part[(size_t)i * 16 + 8 + threadIdx.x]); }
        out[threadIdx.x] = a; out[8 + threadIdx.x] = b;
    }
}
struct ColBounds { float lo[8], hi[8]; };
__global__ void normalize_kernel(float* __restrict__ x, size_t total, int d, const ColBounds cb, const float* __restrict__ dev_bounds) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int k = (int)(i % (size_t)d);
        const float lo = dev_bounds ? dev_bounds[k] : cb.lo[k], hi = dev_bounds ? dev_bounds[8 + k] : cb.hi[k];
        // utils::normalize (src/utils.cu): 2.f * (x - lo) / (hi - lo) - 1.f, without contraction into an fma
        x[i] = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, __fsub_rn(x[i], lo)), __fsub_rn(hi, lo)), 1.f);
    }
}

int grid_for(size_t n, int per_thread) {
    const size_t want = (n + (size_t)256 * per_thread - 1) / ((size_t)256 * per_thread);
    const size_t cap = (size_t)sm_count() * 8;                                  // a multiple of the SM count; grid-stride loops take the rest
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

template <typename T, int OP>
int run_elementwise(const T* a, size_t na, const T* b, size_t nb, T* r, size_t nr, int exact_bounds, cudaStream_t st) {
    int* flag = nullptr;
    if (OP == PINN_OP_DIV) {
        if (cudaMallocAsync(&flag, sizeof(int), st) != cudaSuccess || cudaMemsetAsync(flag, 0, sizeof(int), st) != cudaSuccess) return PINN_ERR_CUDA;
    }
    const size_t bound = ((OP == PINN_OP_MUL || OP == PINN_OP_DIV) && exact_bounds) ? (na < nr ? na : nr) : nr;
    const bool aligned = (((uintptr_t)a | (uintptr_t)b | (uintptr_t)r) & 15) == 0;
    if (bound > 0) {
        if (na == nb && na == nr && aligned) elementwise_same<T, OP><<<grid_for(nr, 16 / sizeof(T) * 2), 256, 0, st>>>(a, b, r, nr, flag);
        else if (na == nr && bound == nr && nb < (1ull << 32) && aligned)
            elementwise_bcast_vec<T, OP, true><<<grid_for(nr, 16 / sizeof(T) * 2), 256, 0, st>>>(a, na, b, nb, r, bound, flag);
        else if (nb == nr && bound == nr && na < (1ull << 32) && aligned)
            elementwise_bcast_vec<T, OP, false><<<grid_for(nr, 16 / sizeof(T) * 2), 256, 0, st>>>(a, na, b, nb, r, bound, flag);
        else elementwise_bcast<T, OP><<<grid_for(bound, 4), 256, 0, st>>>(a, na, b, nb, r, bound, flag);
    }
    if (cudaGetLastError() != cudaSuccess) return PINN_ERR_CUDA;
    if (OP == PINN_OP_DIV) {
        int h = 0;
        if (cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) return PINN_ERR_CUDA;
        cudaFreeAsync(flag, st);
        if (h) return PINN_ERR_DIVISION_BY_ZERO;
    }
    return PINN_OK;
}

template <typename T>
int dispatch_elementwise(int op, const void* a, size_t na, const void* b, size_t nb, void* r, size_t nr, int eb, cudaStream_t st) {
    const T* pa = static_cast<const T*>(a); const T* pb = static_cast<const T*>(b); T* pr = static_cast<T*>(r);
    switch (op) {
        case PINN_OP_ADD: return run_elementwise<T, PINN_OP_ADD>(pa, na, pb, nb, pr, nr, eb, st);
        case PINN_OP_SUB: return run_elementwise<T, PINN_OP_SUB>(pa, na, pb, nb, pr, nr, eb, st);
        case PINN_OP_MUL: return run_elementwise<T, PINN_OP_MUL>(pa, na, pb, nb, pr, nr, eb, st);
        case PINN_OP_DIV: return run_elementwise<T, PINN_OP_DIV>(pa, na, pb, nb, pr, nr, eb, st);
    }
    return PINN_ERR_INVALID_ARGUMENT;
}

template <typename T, int Q>
int launch_matmul(const T* a, int ta, const T* b, int tb, T* r, size_t I, size_t K, size_t J, cudaStream_t st) {
    const dim3 grid((unsigned)((J + 16 * Q - 1) / (16 * Q)), (unsigned)((I + 16 * Q - 1) / (16 * Q)));
    if (ta && tb) matmul_tiled<T, true, true, Q><<<grid, 256, 0, st>>>(a, b, r, I, K, J);
    else if (ta) matmul_tiled<T, true, false, Q><<<grid, 256, 0, st>>>(a, b, r, I, K, J);
    else if (tb) matmul_tiled<T, false, true, Q><<<grid, 256, 0, st>>>(a, b, r, I, K, J);
    else matmul_tiled<T, false, false, Q><<<grid, 256, 0, st>>>(a, b, r, I, K, J);
    return cudaGetLastError() == cudaSuccess ? PINN_OK : PINN_ERR_CUDA;
}
template <typename T>
int run_matmul(const T* a, int ta, const T* b, int tb, T* r, size_t I, size_t K, size_t J, cudaStream_t st) {
    // 128 x 128 tiles (8 x 8 per thread) for float results wide and tall enough to fill them, 64 x 64 otherwise
    if (sizeof(T) == 4 && I > 64 && J > 64) return launch_matmul<T, 8>(a, ta, b, tb, r, I, K, J, st);
    return launch_matmul<T, 4>(a, ta, b, tb, r, I, K, J, st);
}

template <typename T>
int run_dot(const T* a, const T* b, size_t n, T* r, T* scratch, cudaStream_t st) {
    const int blocks = (int)pinn_tensor_dot_scratch(n);
    T* part = scratch;
    if (!part && cudaMallocAsync(&part, (size_t)blocks * sizeof(T), st) != cudaSuccess) return PINN_ERR_CUDA;
    dot_partial<T><<<blocks, 256, 0, st>>>(a, b, n, part);
    dot_final<T><<<1, 256, 0, st>>>(part, blocks, r);
    const bool ok = cudaGetLastError() == cudaSuccess;
    if (!scratch) cudaFreeAsync(part, st);
    return ok ? PINN_OK : PINN_ERR_CUDA;
}

}  // namespace

extern "C" {

int pinn_tensor_elementwise(int op, int dtype, const void* a, size_t na, const void* b, size_t nb, void* r, size_t nr, int exact_bounds, void* stream) {
    if (!a || !b || !r || na == 0 || nb == 0 || nr != (na > nb ? na : nb) || op < 0 || op > 3) return PINN_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == 0) return dispatch_elementwise<float>(op, a, na, b, nb, r, nr, exact_bounds, st);
    if (dtype == 1) return dispatch_elementwise<double>(op, a, na, b, nb, r, nr, exact_bounds, st);
    return PINN_ERR_INVALID_ARGUMENT;
}

int pinn_tensor_matmul(int dtype, const void* a, int a_transposed, const void* b, int b_transposed, void* r, size_t I, size_t K, size_t J, void* stream) {
    if (!a || !b || !r || I == 0 || K == 0 || J == 0) return PINN_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == 0) return run_matmul<float>(static_cast<const float*>(a), a_transposed, static_cast<const float*>(b), b_transposed, static_cast<float*>(r), I, K, J, st);
    if (dtype == 1) return run_matmul<double>(static_cast<const double*>(a), a_transposed, static_cast<const double*>(b), b_transposed, static_cast<double*>(r), I, K, J, st);
    return PINN_ERR_INVALID_ARGUMENT;
}

size_t pinn_tensor_dot_scratch(size_t n) {
    const size_t want = (n + 256 * 16 - 1) / (256 * 16);
    const size_t cap = (size_t)sm_count() * 8;
    return want < 1 ? 1 : (want > cap ? cap : want);
}

int pinn_tensor_dot(int dtype, const void* a, const void* b, size_t n, void* r, void* scratch, void* stream) {
    if (!a || !b || !r || n == 0) return PINN_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == 0) return run_dot<float>(static_cast<const float*>(a), static_cast<const float*>(b), n, static_cast<float*>(r), static_cast<float*>(scratch), st);
    if (dtype == 1) return run_dot<double>(static_cast<const double*>(a), static_cast<const double*>(b), n, static_cast<double*>(r), static_cast<double*>(scratch), st);
    return PINN_ERR_INVALID_ARGUMENT;
}

int pinn_tensor_add_multiple(int dtype, const void* const* srcs, size_t count, size_t n, void* r, void* stream) {
    if (!srcs || !r || count == 0 || count > 16 || n == 0) return PINN_ERR_INVALID_ARGUMENT;
    SrcList l{};
    for (size_t k = 0; k < count; k++) { if (!srcs[k]) return PINN_ERR_INVALID_ARGUMENT; l.p[k] = srcs[k]; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == 0) add_multiple_kernel<float><<<grid_for(n, 4), 256, 0, st>>>(l, (int)count, n, static_cast<float*>(r));
    else if (dtype == 1) add_multiple_kernel<double><<<grid_for(n, 4), 256, 0, st>>>(l, (int)count, n, static_cast<double*>(r));
    else return PINN_ERR_INVALID_ARGUMENT;
    return cudaGetLastError() == cudaSuccess ? PINN_OK : PINN_ERR_CUDA;
}

int pinn_tensor_broadcast(int op, int dtype, const void* a, const size_t* a_shape, int a_dim, const void* b, const size_t* b_shape, int b_dim,
                          void* r, const size_t* r_shape, int r_dim, void* stream) {
    if (!a || !b || !r || !r_shape || r_dim < 1 || r_dim > 8 || a_dim > r_dim || b_dim > r_dim || a_dim < 0 || b_dim < 0 || op < 0 || op > 3 ||
        (a_dim > 0 && !a_shape) || (b_dim > 0 && !b_shape))
        return PINN_ERR_INVALID_ARGUMENT;
    BcastDims bd{};
    bd.dim = r_dim;
    size_t n = 1, stra = 1, strb = 1;
    for (int k = r_dim - 1; k >= 0; k--) {
        const int ka = k - (r_dim - a_dim), kb = k - (r_dim - b_dim);
        const size_t da = ka >= 0 ? a_shape[ka] : 1, db = kb >= 0 ? b_shape[kb] : 1, dr = r_shape[k];
        if (dr == 0 || (da != dr && da != 1) || (db != dr && db != 1) || dr != (da > db ? da : db)) return PINN_ERR_INVALID_ARGUMENT;
        bd.shape[k] = dr; bd.sa[k] = da == 1 ? 0 : stra; bd.sb[k] = db == 1 ? 0 : strb;
        stra *= da; strb *= db; n *= dr;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = grid_for(n, 4);
#define PINN_BCAST(T, OPC) broadcast_kernel<T, OPC><<<grid, 256, 0, st>>>(static_cast<const T*>(a), static_cast<const T*>(b), static_cast<T*>(r), n, bd)
    if (dtype == 0) { switch (op) { case 0: PINN_BCAST(float, 0); break; case 1: PINN_BCAST(float, 1); break; case 2: PINN_BCAST(float, 2); break; default: PINN_BCAST(float, 3); } }
    else if (dtype == 1) { switch (op) { case 0: PINN_BCAST(double, 0); break; case 1: PINN_BCAST(double, 1); break; case 2: PINN_BCAST(double, 2); break; default: PINN_BCAST(double, 3); } }
    else return PINN_ERR_INVALID_ARGUMENT;
#undef PINN_BCAST
    return cudaGetLastError() == cudaSuccess ? PINN_OK : PINN_ERR_CUDA;
}

int pinn_tensor_einsum(int dtype, const char* spec, const void* a, const size_t* a_shape, int a_dim, const void* b, const size_t* b_shape, int b_dim,
                       void* r, void* stream) {
    if (!spec || !a || !b || !r || !a_shape || !b_shape) return PINN_ERR_INVALID_ARGUMENT;
    // parse "<A>,<B>-><R>" with 1 or 2 letters per operand and 0..2 for the result
    char A[3] = {0, 0, 0}, B[3] = {0, 0, 0}, R[3] = {0, 0, 0};
    int na = 0, nb = 0, nr = 0, part = 0;
    for (const char* p = spec; *p; p++) {
        if (*p == ' ') continue;
        if (*p == ',') { if (part != 0) return PINN_ERR_INVALID_ARGUMENT; part = 1; continue; }
        if (*p == '-' && p[1] == '>') { if (part != 1) return PINN_ERR_INVALID_ARGUMENT; part = 2; p++; continue; }
        const bool letter = (*p >= 'a' && *p <= 'z') || (*p >= 'A' && *p <= 'Z');
        if (!letter) return PINN_ERR_INVALID_ARGUMENT;
        if (part == 0) { if (na == 2) return PINN_ERR_UNSUPPORTED; A[na++] = *p; }
        else if (part == 1) { if (nb == 2) return PINN_ERR_UNSUPPORTED; B[nb++] = *p; }
        else { if (nr == 2) return PINN_ERR_UNSUPPORTED; R[nr++] = *p; }
    }
    if (part != 2 || na == 0 || nb == 0 || na != a_dim || nb != b_dim) return PINN_ERR_INVALID_ARGUMENT;
    auto size_of = [&](char c) -> size_t { for (int i = 0; i < na; i++) if (A[i] == c) return a_shape[i]; for (int i = 0; i < nb; i++) if (B[i] == c) return b_shape[i]; return 0; };
    for (int i = 0; i < na; i++) for (int k = 0; k < nb; k++) if (A[i] == B[k] && a_shape[i] != b_shape[k]) return PINN_ERR_INVALID_ARGUMENT;
    if ((na == 2 && A[0] == A[1]) || (nb == 2 && B[0] == B[1])) return PINN_ERR_UNSUPPORTED;
    if (na == 1 && nb == 1) {
        if (A[0] == B[0] && nr == 0) return pinn_tensor_dot(dtype, a, b, a_shape[0], r, nullptr, stream);                       // i,i->
        if (A[0] != B[0] && nr == 2 && R[0] == A[0] && R[1] == B[0]) return pinn_tensor_matmul(dtype, a, 0, b, 0, r, a_shape[0], 1, b_shape[0], stream);   // i,j->ij
        return PINN_ERR_UNSUPPORTED;
    }
    if (na == 2 && nb == 2 && nr == 2) {
        // the contracted letter: in both operands, not in the result
        char c = 0;
        for (int i = 0; i < 2; i++) for (int k = 0; k < 2; k++) if (A[i] == B[k]) c = A[i];
        if (!c || R[0] == c || R[1] == c) return PINN_ERR_UNSUPPORTED;
        const char fa = A[0] == c ? A[1] : A[0], fb = B[0] == c ? B[1] : B[0];
        if (fa == fb || R[0] != fa || R[1] != fb) return PINN_ERR_UNSUPPORTED;
        return pinn_tensor_matmul(dtype, a, A[0] == c, b, B[1] == c, r, size_of(fa), size_of(c), size_of(fb), stream);
    }
    if (na == 2 && nb == 1 && nr == 1) {                                                                                                   // ij,j->i   ji,j->i
        const char c = B[0];
        if (A[0] != c && A[1] != c) return PINN_ERR_UNSUPPORTED;
        const char fa = A[0] == c ? A[1] : A[0];
        if (R[0] != fa) return PINN_ERR_UNSUPPORTED;
        return pinn_tensor_matmul(dtype, a, A[0] == c, b, 0, r, size_of(fa), size_of(c), 1, stream);
    }
    if (na == 1 && nb == 2 && nr == 1) {                                                                                                   // i,ij->j   i,ji->j
        const char c = A[0];
        if (B[0] != c && B[1] != c) return PINN_ERR_UNSUPPORTED;
        const char fb = B[0] == c ? B[1] : B[0];
        if (R[0] != fb) return PINN_ERR_UNSUPPORTED;
        return pinn_tensor_matmul(dtype, a, 0, b, B[1] == c, r, 1, size_of(c), size_of(fb), stream);
    }
    return PINN_ERR_UNSUPPORTED;
}

int pinn_tensor_normalize_columns(float* x, size_t n, size_t d, const float* lo, const float* hi, float* stats_out, void* stream) {
    if (!x || n == 0 || d == 0 || d > 8 || ((lo == nullptr) != (hi == nullptr))) return PINN_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ColBounds cb{};
    float* dev_bounds = nullptr;
    float* part = nullptr;
    if (lo) {
        for (size_t k = 0; k < d; k++) { cb.lo[k] = lo[k]; cb.hi[k] = hi[k]; }
    } else {
        const int blocks = grid_for(n, 8);
        if (cudaMallocAsync(&part, (size_t)blocks * 16 * sizeof(float), st) != cudaSuccess || cudaMallocAsync(&dev_bounds, 16 * sizeof(float), st) != cudaSuccess) return PINN_ERR_CUDA;
        colminmax_partial<<<blocks, 256, 0, st>>>(x, n, (int)d, part);
        colminmax_final<<<1, 32, 0, st>>>(part, blocks, (int)d, dev_bounds);
    }
    normalize_kernel<<<grid_for(n * d, 4), 256, 0, st>>>(x, n * d, (int)d, cb, dev_bounds);
    int rc = cudaGetLastError() == cudaSuccess ? PINN_OK : PINN_ERR_CUDA;
    if (dev_bounds) {
        float h[16];
        if (cudaMemcpyAsync(h, dev_bounds, sizeof(h), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) rc = PINN_ERR_CUDA;
        else for (size_t k = 0; k < d; k++) { cb.lo[k] = h[k]; cb.hi[k] = h[8 + k]; }
        cudaFreeAsync(dev_bounds, st); cudaFreeAsync(part, st);
    }
    if (stats_out && rc == PINN_OK) for (size_t k = 0; k < d; k++) { stats_out[k] = cb.lo[k]; stats_out[d + k] = cb.hi[k]; }
    return rc;
}

int pinn_device_alloc(void** ptr, size_t bytes, void* stream) {
    if (!ptr || bytes == 0) return PINN_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (cudaMallocAsync(ptr, bytes, st) != cudaSuccess) return PINN_ERR_CUDA;
    return cudaMemsetAsync(*ptr, 0, bytes, st) == cudaSuccess ? PINN_OK : PINN_ERR_CUDA;   // (the reference zero-fills new tensors, interface.cuh:99-103)
}
int pinn_device_free(void* ptr, void* stream) {
    if (!ptr) return PINN_OK;
    return cudaFreeAsync(ptr, static_cast<cudaStream_t>(stream)) == cudaSuccess ? PINN_OK : PINN_ERR_CUDA;
}
int pinn_device_upload(void* dst, const void* host_src, size_t bytes, void* stream) {
    if (!dst || !host_src) return PINN_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (cudaMemcpyAsync(dst, host_src, bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) return PINN_ERR_CUDA;
    return cudaStreamSynchronize(st) == cudaSuccess ? PINN_OK : PINN_ERR_CUDA;
}
int pinn_device_download(void* host_dst, const void* src, size_t bytes, void* stream) {
    if (!host_dst || !src) return PINN_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (cudaMemcpyAsync(host_dst, src, bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess) return PINN_ERR_CUDA;
    return cudaStreamSynchronize(st) == cudaSuccess ? PINN_OK : PINN_ERR_CUDA;
}

int pinn_device_copy(void* dst, const void* src, size_t bytes, void* stream) {
    if (!dst || !src) return PINN_ERR_INVALID_ARGUMENT;
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)) == cudaSuccess ? PINN_OK : PINN_ERR_CUDA;
}

}  // extern "C"
