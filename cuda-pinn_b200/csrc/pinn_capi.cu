// pinn_capi.cu — host side of the B200 PINN train step behind the C ABI of include/pinn_abi.h.
// Owns device memory, the stream, kernel selection and (N>1) the NCCL communicator.
// There is no CPU arithmetic path for the step: every compute entry point runs CUDA kernels or fails
// (host arithmetic exists only for the analytic comparison solutions of pinn_exact_solution).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <memory>
#include <string>
#include <vector>
#include <cuda_runtime.h>

#include "../../include/pinn_abi.h"
#include "pinn_device.cuh"
#include "pinn_fused_fp32.cuh"
#include "pinn_fused_tc.cuh"
#include "pinn_fused_tc2.cuh"
#include "pinn_fused_tc3.cuh"
#include "pinn_fused_mma.cuh"

using namespace pinn;

// ------------------------------------------------------------------------------------------
// small kernels around the fused step
// ------------------------------------------------------------------------------------------

// Collocation generation (reference src/utils.cu is empty): one thread per point of the shard.
__global__ void sample_points_kernel(const NetDims nd, int set, uint64_t seed, uint64_t begin, uint64_t count,
                                     float* __restrict__ x) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float p[3];
    sample_point(nd, set, seed, begin + i, p);
    for (int k = 0; k < nd.d_in; k++) x[i * nd.d_in + k] = p[k];
}

// gbuf[col] = sum over the CTA partial rows of both launches of the step, in a fixed order
// (bit-reproducible); rows are zeroed for the next step.  Block = 32 columns x 32 row lanes: lane y
// adds rows y, y+32, ... of range A then of range B; the 32 lane sums are combined in lane order.
__global__ void __launch_bounds__(1024)
reduce_partials_kernel(float* __restrict__ part_a, int rows_a, float* __restrict__ part_b, int rows_b, int PP, int n_out,
                       float* __restrict__ gbuf, unsigned long long* step, int bump, int zero_a) {
    __shared__ float part[32][33];
    constexpr int RU = 8;                                // independent loads in flight per thread
    const int col = blockIdx.x * 32 + threadIdx.x;
    if (blockIdx.x == 0 && threadIdx.x == 0 && threadIdx.y == 0 && bump) *step += 1ULL;
    float s = 0.0f;
    if (col < n_out) {
        auto sum_rows = [&](const float* __restrict__ base, int rows) {
            for (int r0 = threadIdx.y; r0 < rows; r0 += 32 * RU) {
                float v[RU];
                #pragma unroll
                for (int u = 0; u < RU; u++) {
                    const int r = r0 + 32 * u;
                    v[u] = (r < rows) ? __ldcg(base + (size_t)r * PP + col) : 0.0f;
                }
                #pragma unroll
                for (int u = 0; u < RU; u++) s += v[u];  // rows in increasing order: fixed summation order
            }
        };
        sum_rows(part_a, rows_a);
        sum_rows(part_b, rows_b);
        if (zero_a) {                                     // (rows written by plain stores — the mma path — need none)
            for (int r = threadIdx.y; r < rows_a; r += 32) part_a[(size_t)r * PP + col] = 0.0f;
            for (int r = threadIdx.y; r < rows_b; r += 32) part_b[(size_t)r * PP + col] = 0.0f;
        }
    }
    part[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && col < n_out) {
        float t = 0.0f;
        #pragma unroll
        for (int y = 0; y < 32; y++) t += part[y][threadIdx.x];
        gbuf[col] = t;
    }
}

struct AdamArgs {
    float lr, beta1, beta2, eps;
    float lambda_r, lambda_b, lambda_0;
    double inv_nf, inv_nb, inv_n0;
};

// Fragment-ordered BF16 mirror of the width-20 mma.sync kernel (pinn_fused_mma.cuh, layout of pack_wfrag_kernel),
// viewed as 16-bit halves; w16 == nullptr: the handle has none.
struct FragMirror { unsigned short* w16; int nterms; int f16; };

// the BF16 terms of hidden->hidden weight W_l[k][j] into both directions of the mirror
__device__ __forceinline__ void frag_mirror_store(const FragMirror& fm, int l, int k, int j, float w) {
    uint32_t f[3];
    if (fm.f16) { uint32_t h[2]; split_pair_f16(w * (float)(1 << kF16WeightShift), 0.0f, h); f[0] = h[0]; f[1] = h[1]; f[2] = 0u; }
    else split_pair<3>(w, 0.0f, f);                         // low halves: hi, mid, lo of w
    const int words = 9 * fm.nterms;
    #pragma unroll
    for (int dir = 0; dir < 2; dir++) {
        const int kk = dir ? j : k, n = dir ? k : j;        // dir 0: B[k][n] = W[k][n];  dir 1: B[k][n] = W[n][k]
        const int r = kk < 8 ? 0 : (kk < 16 ? 1 : 2), kl = kk - 8 * r;
        const int lane = (n & 7) * 4 + (kl >> 1);
        for (int term = 0; term < fm.nterms; term++) {
            const size_t word = ((size_t)((l - 2) * 2 + dir) * words + ((n >> 3) * fm.nterms + term) * 3 + r) * 32 + lane;
            fm.w16[word * 2 + (kl & 1)] = (unsigned short)(f[term] & 0xffffu);
        }
    }
}

// Adam for parameter i + refresh of the transposed mirror W^T[l] ([fan_out, fan_in]) and, where the
// handle has one, of the BF16 fragment mirror.
__device__ __forceinline__ void adam_core(const NetDims& nd, const AdamArgs& a, float* __restrict__ params,
                                          float* __restrict__ paramsT, float* __restrict__ m, float* __restrict__ v,
                                          float g, float c1, float c2, int i, const FragMirror& fm) {
    const float mi = a.beta1 * m[i] + (1.0f - a.beta1) * g;
    const float vi = a.beta2 * v[i] + (1.0f - a.beta2) * (g * g);
    m[i] = mi; v[i] = vi;
    const float mh = mi * c1, vh = vi * c2;
    const float w = params[i] - a.lr * (mh / (sqrtf(vh) + a.eps));
    params[i] = w;
    int l = 1;
    while (l <= nd.L && i >= nd.offW[l + 1]) l++;
    const int fi = nd.fan_in(l), fo = nd.fan_out(l);
    if (i < nd.offB[l]) {
        const int r = i - nd.offW[l], k = r / fo, j = r - k * fo;
        paramsT[nd.offW[l] + j * fi + k] = w;
        if (fm.w16 && l >= 2 && l <= nd.L) frag_mirror_store(fm, l, k, j, w);
    } else {
        paramsT[i] = w;
    }
}
__device__ __forceinline__ void adam_one(const NetDims& nd, const AdamArgs& a, float* __restrict__ params,
                                         float* __restrict__ paramsT, float* __restrict__ m, float* __restrict__ v,
                                         float g, double t, int i, const FragMirror& fm) {
    const float c1 = (float)(1.0 / (1.0 - pow((double)a.beta1, t)));
    const float c2 = (float)(1.0 / (1.0 - pow((double)a.beta2, t)));
    adam_core(nd, a, params, paramsT, m, v, g, c1, c2, i, fm);
}

// hrec (may be null): the handle's MAPPED pinned host record, two 16-byte halves {loss, mse_r, mse_b, seq} {mse_0, flag, seq, 0}
// (seq = low 32 bits of the step number).  The step tail writes it itself — two vector stores over PCIe, each carrying the
// sequence number, so no system fence is needed — and the host reads a finished step by polling its own memory instead of
// queueing a D2H copy and synchronising the stream.
__device__ __forceinline__ void loss_record(const NetDims& nd, const AdamArgs& a, float sr, float sb, float s0,
                                            float* __restrict__ loss_out, int* __restrict__ flag, float* hrec,
                                            unsigned long long seq) {
    const float mr = (float)((double)sr * a.inv_nf);
    const float mb = (float)((double)sb * a.inv_nb);
    const float m0 = (float)((double)s0 * a.inv_n0);
    const float loss = a.lambda_r * mr + a.lambda_b * mb + a.lambda_0 * m0;
    loss_out[0] = loss; loss_out[1] = mr; loss_out[2] = mb; loss_out[3] = m0;
    if (!isfinite(loss)) *flag = PINN_ERR_DEVICE_OVERFLOW;       // reference runtime.cuh error channel
    if (hrec) {
        const float fs = __uint_as_float((unsigned int)seq);
        reinterpret_cast<float4*>(hrec)[0] = make_float4(loss, mr, mb, fs);
        reinterpret_cast<float4*>(hrec)[1] = make_float4(m0, __int_as_float(*flag), fs, 0.0f);
    }
}

__global__ void adam_kernel(const NetDims nd, const AdamArgs a, float* __restrict__ params,
                            float* __restrict__ paramsT, float* __restrict__ m, float* __restrict__ v,
                            const float* __restrict__ gbuf, const unsigned long long* __restrict__ step,
                            float* __restrict__ loss_out, int* __restrict__ flag, int apply, const FragMirror fm, float* hrec,
                            double* __restrict__ beta_pow) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) loss_record(nd, a, gbuf[nd.P + 0], gbuf[nd.P + 1], gbuf[nd.P + 2], loss_out, flag, apply ? hrec : nullptr, *step);
    if (i == 0 && apply) { beta_pow[0] *= (double)a.beta1; beta_pow[1] *= (double)a.beta2; }   // keep the single-GPU tail's running powers in step
    if (i >= nd.P || !apply) return;
    adam_one(nd, a, params, paramsT, m, v, gbuf[i], (double)(*step), i, fm);
}

// Single-GPU step tail in ONE launch: fixed-order sum of the partial rows, Adam on the block's own 32 columns, mirrors; the
// last block to finish records the loss and bumps the step counter.  Latency-oriented shape (the step is ~80 us, this kernel
// must not be 15 of them): block = 8 column quads x 128 row lanes, float4 loads, so a thread walks rows/128 rows (13 for the
// 1,700 warp rows of C2) with 4 loads in flight; lane sums meet in shared memory and are added in lane order (deterministic).
constexpr int kTailLanes = 128;
__global__ void __launch_bounds__(8 * kTailLanes)
reduce_adam_kernel(const NetDims nd, const AdamArgs a, float* __restrict__ part_a, int rows_a, int zero_a,
                   float* __restrict__ part_b, int rows_b, float* __restrict__ params, float* __restrict__ paramsT,
                   float* __restrict__ m, float* __restrict__ v, float* __restrict__ gbuf, unsigned long long* __restrict__ step,
                   float* __restrict__ loss_out, int* __restrict__ flag, unsigned int* __restrict__ done_counter, const FragMirror fm,
                   float* hrec, double* __restrict__ beta_pow) {
    __shared__ float part[kTailLanes][33];
    constexpr int RU = 4;
    const int PP = nd.PP, n_out = nd.P + 3;
    const int col0 = blockIdx.x * 32 + threadIdx.x * 4;    // this thread's column quad (PP is a multiple of 32: always inside a row)
    const unsigned long long tstep = *step + 1ULL;         // every block reads it before the LAST block writes it back
    // bias corrections from the running powers beta^t (one multiply per step instead of two double pow() in the only warp
    // of the block that runs Adam — they were the longest dependent chain of this kernel)
    const double b1t = beta_pow[0] * (double)a.beta1, b2t = beta_pow[1] * (double)a.beta2;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    auto sum_rows = [&](float* __restrict__ base, int rows, bool zero) {
        for (int r0 = threadIdx.y; r0 < rows; r0 += kTailLanes * RU) {
            float4 vv[RU];
            #pragma unroll
            for (int u = 0; u < RU; u++) {
                const int r = r0 + kTailLanes * u;
                vv[u] = (r < rows) ? __ldcg(reinterpret_cast<const float4*>(base + (size_t)r * PP + col0)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            #pragma unroll
            for (int u = 0; u < RU; u++) {                 // rows in increasing order: fixed summation order
                s.x += vv[u].x; s.y += vv[u].y; s.z += vv[u].z; s.w += vv[u].w;
                const int r = r0 + kTailLanes * u;
                if (zero && r < rows) *reinterpret_cast<float4*>(base + (size_t)r * PP + col0) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    };
    sum_rows(part_a, rows_a, zero_a != 0);
    sum_rows(part_b, rows_b, zero_a != 0);
    part[threadIdx.y][threadIdx.x * 4 + 0] = s.x; part[threadIdx.y][threadIdx.x * 4 + 1] = s.y;
    part[threadIdx.y][threadIdx.x * 4 + 2] = s.z; part[threadIdx.y][threadIdx.x * 4 + 3] = s.w;
    __syncthreads();
    const int tid = threadIdx.y * 8 + threadIdx.x;
    if (tid < 32) {
        const int col = blockIdx.x * 32 + tid;
        if (col < n_out) {
            float t = 0.0f;
            #pragma unroll 16
            for (int y = 0; y < kTailLanes; y++) t += part[y][tid];
            gbuf[col] = t;
            if (col < nd.P) adam_core(nd, a, params, paramsT, m, v, t, (float)(1.0 / (1.0 - b1t)), (float)(1.0 / (1.0 - b2t)), col, fm);
        }
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(done_counter, 1u) == gridDim.x - 1) {
            *done_counter = 0u;
            __threadfence();
            *step = tstep;
            beta_pow[0] = b1t; beta_pow[1] = b2t;
            loss_record(nd, a, __ldcg(gbuf + nd.P), __ldcg(gbuf + nd.P + 1), __ldcg(gbuf + nd.P + 2), loss_out, flag, hrec, tstep);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Fused gradient reduction + all-reduce + Adam over NVLink peer memory (one kernel per rank per step).
// PUSH protocol — every rank owns an exchange buffer (IPC-mapped into all peers) with one slot per
// SOURCE rank and parity, and one flag word per source rank:
//   1. fixed-order sum of this GPU's CTA partial rows (as reduce_partials_kernel); each column sum is
//      STORED into slot[this rank][parity] of EVERY rank's buffer (NVLink stores, fire and forget);
//   2. the last block to finish does threadfence_system and release-stores the sequence number into
//      flag[this rank] of every rank's buffer;
//   3. every block polls only its OWN buffer's flags (local memory) until all sources have reached the
//      sequence number (bounded spin), then sums the R local slots in RANK ORDER — every rank adds the
//      same numbers in the same order, so the weights stay bit-identical;
//   4. Adam on the summed gradient, loss record.
// Replaces reduce kernel + ncclAllReduce + Adam kernel; the payload is P+3 floats (12 KB .. 1.8 MB).
// Parity double buffering + the flags make a slot reusable: a rank can write parity q of a peer again only
// after that peer has published the step in between, i.e. after it finished reading q.
// ------------------------------------------------------------------------------------------
constexpr int kMaxP2PRanks = 16;
constexpr int kXchgHeader = 32;                         // floats in front of the slots: one 64-bit flag per source rank
struct P2PArgs {
    float* peers[kMaxP2PRanks];                          // exchange buffers in rank order (own included)
    int nranks, rank;
    unsigned long long seq;                              // monotonic over the life of the communicator
    unsigned int* done_counter;                          // local, zero between launches
    int* err_flag;                                       // local: set when a peer never showed up
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(1024)
reduce_exchange_adam_kernel(const NetDims nd, const AdamArgs a, float* __restrict__ part_a, int rows_a,
                            float* __restrict__ part_b, int rows_b, float* __restrict__ params,
                            float* __restrict__ paramsT, float* __restrict__ m, float* __restrict__ v,
                            float* __restrict__ gbuf, unsigned long long* __restrict__ step, unsigned long long tstep,
                            float* __restrict__ loss_out, int* __restrict__ flag, const P2PArgs x, const FragMirror fm, float* hrec,
                            double* __restrict__ beta_pow) {
    __shared__ float part[32][33];
    __shared__ int peer_missing;
    constexpr int RU = 8;
    const int PP = nd.PP, n_out = nd.P + 3;
    const int col = blockIdx.x * 32 + threadIdx.x;
    if (threadIdx.x == 0 && threadIdx.y == 0) peer_missing = 0;
    float s = 0.0f;
    if (col < n_out) {
        auto sum_rows = [&](const float* __restrict__ base, int rows) {
            for (int r0 = threadIdx.y; r0 < rows; r0 += 32 * RU) {
                float vv[RU];
                #pragma unroll
                for (int u = 0; u < RU; u++) {
                    const int r = r0 + 32 * u;
                    vv[u] = (r < rows) ? __ldcg(base + (size_t)r * PP + col) : 0.0f;
                }
                #pragma unroll
                for (int u = 0; u < RU; u++) s += vv[u];
            }
        };
        sum_rows(part_a, rows_a);
        sum_rows(part_b, rows_b);
        for (int r = threadIdx.y; r < rows_a; r += 32) part_a[(size_t)r * PP + col] = 0.0f;
        for (int r = threadIdx.y; r < rows_b; r += 32) part_b[(size_t)r * PP + col] = 0.0f;
    }
    part[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    const int par = (int)(x.seq & 1ULL);
    if (threadIdx.y == 0 && col < n_out) {
        float t = 0.0f;
        #pragma unroll
        for (int y = 0; y < 32; y++) t += part[y][threadIdx.x];
        const size_t slot = kXchgHeader + ((size_t)(x.rank * 2 + par)) * PP + col;
        for (int r = 0; r < x.nranks; r++) x.peers[r][slot] = t;        // push this rank's local sum to every rank (own included)
    }
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        __threadfence();                                               // this block's pushes ordered before its arrival (GPU scope)
        const unsigned int old = atomicAdd(x.done_counter, 1u);
        if (old == gridDim.x - 1) {                                    // every block of this rank has pushed its columns
            *x.done_counter = 0u;
            __threadfence_system();                                    // ONE system fence (cumulative over the arrivals observed above) ...
            for (int r = 0; r < x.nranks; r++)                         // ... then relaxed flag stores: all my columns have landed in rank r
                st_relaxed_sys(reinterpret_cast<unsigned long long*>(x.peers[r]) + x.rank, x.seq);
            *step = tstep;
        }
        const long long t0 = clock64();
        const unsigned long long* my_flags = reinterpret_cast<const unsigned long long*>(x.peers[x.rank]);
        for (int r = 0; r < x.nranks; r++) {                           // poll LOCAL memory only
            while (ld_acquire_sys(my_flags + r) < x.seq) {
                if (clock64() - t0 > 4000000000LL) { *x.err_flag = PINN_ERR_NCCL; peer_missing = 1; break; }   // ~2 s: a peer never arrived
                __nanosleep(100);
            }
        }
    }
    __syncthreads();
    // a peer that never arrived: its slot holds an old step — do NOT apply Adam to an incomplete sum (the ranks would drift
    // apart silently); the flag reaches the caller through the step record (pinn_train_step returns PINN_ERR_NCCL)
    if (peer_missing) {
        if (threadIdx.y == 0 && col == nd.P && hrec) {
            const float fs = __uint_as_float((unsigned int)tstep);
            reinterpret_cast<float4*>(hrec)[0] = make_float4(0.0f, 0.0f, 0.0f, fs);
            reinterpret_cast<float4*>(hrec)[1] = make_float4(0.0f, __int_as_float(PINN_ERR_NCCL), fs, 0.0f);
        }
        return;
    }
    if (threadIdx.y == 0 && col < n_out) {
        const float* mine = x.peers[x.rank] + kXchgHeader;
        float g = 0.0f;
        for (int r = 0; r < x.nranks; r++) g += __ldcv(mine + ((size_t)(r * 2 + par)) * PP + col);   // rank order, local reads
        gbuf[col] = g;
        if (col < nd.P) {
            adam_one(nd, a, params, paramsT, m, v, g, (double)tstep, col, fm);
        } else if (col == nd.P) {
            float s3[3];
            #pragma unroll
            for (int q = 0; q < 3; q++) {
                float t = 0.0f;
                for (int r = 0; r < x.nranks; r++) t += __ldcv(mine + ((size_t)(r * 2 + par)) * PP + nd.P + q);
                s3[q] = t;
            }
            loss_record(nd, a, s3[0], s3[1], s3[2], loss_out, flag, hrec, tstep);
            beta_pow[0] *= (double)a.beta1; beta_pow[1] *= (double)a.beta2;
        }
    }
}

__global__ void transpose_params_kernel(const NetDims nd, const float* __restrict__ params,
                                        float* __restrict__ paramsT) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nd.P) return;
    int l = 1;
    while (l <= nd.L && i >= nd.offW[l + 1]) l++;
    const int fi = nd.fan_in(l), fo = nd.fan_out(l);
    if (i < nd.offB[l]) {
        const int r = i - nd.offW[l], k = r / fo, j = r - k * fo;
        paramsT[nd.offW[l] + j * fi + k] = params[i];
    } else {
        paramsT[i] = params[i];
    }
}

// ------------------------------------------------------------------------------------------
// NCCL through dlopen (torch's bundled libnccl.so.2 when loaded in a torch process, else the
// system one); no link-time dependency so single-GPU users need no NCCL at all.
// ------------------------------------------------------------------------------------------
namespace {
struct NcclId { char b[128]; };          // ncclUniqueId (nccl.h: 128 opaque bytes, passed by value)
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;
bool nccl_load(std::string& err) {
    if (g_nccl.ok) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) { err = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return false; }
    auto sym = [&](const char* s) { return dlsym(g_nccl.handle, s); };
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))sym("ncclAllReduce");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
        err = "libnccl is missing a required symbol"; return false;
    }
    g_nccl.ok = true;
    return true;
}
constexpr int kNcclFloat32 = 7, kNcclSum = 0;   // ncclDataType_t / ncclRedOp_t values (nccl.h)
std::string g_create_error;
}  // namespace

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
typedef void (*fused_fn)(const NetDims, const FusedArgs);

struct pinn_ctx {
    pinn_config cfg{};
    NetDims nd{};
    cudaStream_t stream = nullptr, stream2 = nullptr;   // stream2: boundary/initial launch, forked per step
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaGraph_t graph = nullptr; cudaGraphExec_t graph_exec = nullptr;
    uint64_t graph_key[4] = {0, 0, 0, 0}; bool graph_ok = true, capturing = false;
    std::string graph_note;                            // why the step is not a graph replay (empty: it is, or profiling is on)
    uint64_t step_launches = 0, launches_per_graph = 0;
    float *d_params = nullptr, *d_paramsT = nullptr, *d_m = nullptr, *d_v = nullptr;
    float *d_gbuf = nullptr, *d_partial = nullptr, *d_stash = nullptr, *d_loss = nullptr;
    unsigned long long* d_step = nullptr;
    int* d_flag = nullptr;
    float* d_x[3] = {nullptr, nullptr, nullptr};
    uint64_t begin[3] = {0, 0, 0}, cap[3] = {0, 0, 0}, n_local[3] = {0, 0, 0};
    uint64_t seeds[3] = {0, 0, 0};
    float scale[3] = {0, 0, 0};
    float* h_pinned = nullptr;          // [8]: loss record + flag (+ step sequence word), mapped: the step tail writes it over PCIe
    float* d_hrec = nullptr;            // device address of h_pinned
    const float* x_override = nullptr;  // pinn_train_step_host: device-mapped address of the caller's pinned interior points
    // launch configuration of the fused kernels
    fused_fn kern_full = nullptr, kern_val = nullptr;
    int TP = 32, UT = 4, threads = 0, smem_full = 0, smem_val = 0, grid_cap_full = 0, grid_cap_val = 0;
    int val_TP = 32, val_UT = 4, val_threads = 0;   // value-only (boundary / initial / pinn_forward) launch shape
    int rows_cap_a = 0, rows_cap_b = 0, nsplit_cap = 1, kernel_id = 0, num_sms = 148;
    long long stash_per_cta = 0;
    // tensor-core path (precision != FP32): interior set only
    typedef void (*tc_fn)(const NetDims, const TcArgs);
    tc_fn kern_tc = nullptr, kern_tc_val = nullptr;     // kern_tc_val: value-only tiles (boundary / initial sets)
    __nv_bfloat16* d_wq = nullptr;
    long long wq_term_stride = 0;
    uint8_t* d_ring = nullptr; long long ring_per_cta = 0; int tc_defer_g = 0;   // deferred dW (width 256)
    int tc_ut = 8, nterms = 0, tc_threads = 0, tc_smem = 0, tc_tmem_cols = 0, grid_cap_tc = 0;
    // producer / dW-server tensor-core kernel (pinn_fused_tc2.cuh): training launches of the interior set, widths 128 / 256
    typedef void (*tc2_fn)(const NetDims, const Tc2Args);
    tc2_fn kern_tc2 = nullptr;
    int tc2_q = 0, tc2_nprod = 0, tc2_zdepth = 4, tc2_smem = 0, tc2_T = tc2::T, tc2_threads = 0, tc2_grid = 0;   // tc2_grid: CTAs of the launch (num_sms, or twice that at width 128)   // tc2_T: points per producer tile (tc3: 20)
    __nv_bfloat16 *d_wf = nullptr, *d_wb = nullptr;
    uint8_t *d_st_a = nullptr, *d_st_s = nullptr, *d_zring = nullptr;
    unsigned int* d_tc2_flags = nullptr;
    long long* d_tc2_dbg = nullptr;                    // PINN_TC2_DBG=1: phase cycle counters of producer 0, printed at destroy
    // register-chained mma.sync kernel (width 20, PINN_PREC_BF16X3): interior set only
    typedef void (*mma_fn)(const NetDims, const MmaArgs);
    mma_fn kern_mma = nullptr;
    uint32_t* d_wfrag = nullptr;
    int grid_cap_mma = 0, mma_smem = 0, mma_terms = 0, mma_f16 = 0;
    AdamArgs adam{};
    void* comm = nullptr;
    // peer-memory exchange (fused reduce + all-reduce + Adam)
    float* d_xchg = nullptr; float* peer_ptr[kMaxP2PRanks] = {nullptr}; bool p2p_ready = false;
    unsigned long long xseq = 0; unsigned int* d_done = nullptr;
    unsigned int* d_tail = nullptr;     // arrival counter of reduce_adam_kernel
    double* d_beta_pow = nullptr;       // {beta1^t, beta2^t} at the device step count t (single-GPU tail only)
    std::string err;
    // profiling
    bool prof = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
    double kernel_ms = 0.0; uint64_t kernel_count = 0, launches = 0;
    int last_grid = 0, rows_a = 0, rows_b = 0;        // partial rows the last interior / boundary launches used
    uint64_t host_step = 0;
    unsigned long long h2d_point_bytes = 0;            // host->device bytes this handle has copied for point sets
};

static int fail(pinn_ctx* c, int code, const std::string& msg) {
    if (c) {
        c->err = msg;
        // a wait inside the producer / dW-server kernel that timed out left its site in the mapped record before trapping
        const volatile int* ts = c->h_pinned ? reinterpret_cast<const volatile int*>(c->h_pinned + 16) : nullptr;
        if (ts && ts[0] != 0)
            c->err += " [tensor-core kernel hand-off timed out: wait site " + std::to_string(ts[0]) + ", block " + std::to_string(ts[1]) +
                      ", a " + std::to_string(ts[2]) + ", b " + std::to_string(ts[3]) + "]";
    } else g_create_error = msg;
    return code;
}
// A handle created with nranks > 1 holds a SHARD: without a communicator its sums are local yet scaled by the global counts.
static int need_comm(pinn_ctx* c, const char* fn) {
    if (c->cfg.nranks > 1 && !c->comm)
        return fail(c, PINN_ERR_INVALID_ARGUMENT, std::string(fn) + ": nranks > 1 but no communicator (pinn_comm_init), or drive the "
                    "collective yourself with pinn_step_local / pinn_grad_buffer / pinn_step_apply");
    return PINN_OK;
}

#define CUDA_TRY(c, fn, call)                                                                         \
    do { cudaError_t e_ = (call);                                                                     \
         if (e_ != cudaSuccess) return fail((c), PINN_ERR_CUDA, std::string(fn) + ": " + #call + " -> " + \
                                            cudaGetErrorString(e_)); } while (0)

static int pde_shape(int pde, int* n_space, int* n2, int* n_res, int* d_out) {
    switch (pde) {
    case PINN_PDE_BURGERS:       *n_space = 1; *n2 = 1; *n_res = 1; *d_out = 1; return 1;
    case PINN_PDE_HELMHOLTZ:     *n_space = 2; *n2 = 2; *n_res = 1; *d_out = 1; return 1;
    case PINN_PDE_ALLEN_CAHN:    *n_space = 2; *n2 = 2; *n_res = 1; *d_out = 1; return 1;
    case PINN_PDE_HEAT:          *n_space = 2; *n2 = 2; *n_res = 1; *d_out = 1; return 1;
    case PINN_PDE_NAVIER_STOKES: *n_space = 2; *n2 = 2; *n_res = 3; *d_out = 3; return 1;
    }
    return 0;
}

static bool make_dims(const pinn_config* cfg, NetDims* nd, std::string& why) {
    memset(nd, 0, sizeof(*nd));
    int n_space, n2, n_res, d_out;
    if (!pde_shape(cfg->pde, &n_space, &n2, &n_res, &d_out)) { why = "unknown pde"; return false; }
    const int want_din = n_space + (cfg->pde == PINN_PDE_HELMHOLTZ ? 0 : 1);
    if (cfg->d_in != want_din) { why = "d_in does not match the pde"; return false; }
    if (cfg->d_out != d_out) { why = "d_out does not match the pde"; return false; }
    if (cfg->depth < 1 || cfg->depth > kMaxLayers) { why = "depth out of range [1,32]"; return false; }
    if (cfg->width < 4 || cfg->width > 1024 || (cfg->width & 3)) { why = "width must be a multiple of 4 in [4,1024]"; return false; }
    nd->d_in = cfg->d_in; nd->d_out = cfg->d_out; nd->H = cfg->width; nd->L = cfg->depth; nd->pde = cfg->pde;
    nd->n2 = n2; nd->n_space = n_space; nd->n_res = n_res; nd->has_t = cfg->d_in > n_space;
    nd->C = 1 + cfg->d_in + n2;
    for (int k = 0; k < cfg->d_in; k++) {
        const bool is_t = nd->has_t && k == cfg->d_in - 1;
        if (is_t) { nd->lo[k] = 0.0f; nd->hi[k] = 1.0f; }
        else if (cfg->pde == PINN_PDE_NAVIER_STOKES) { nd->lo[k] = 0.0f; nd->hi[k] = kTwoPi; }
        else { nd->lo[k] = -1.0f; nd->hi[k] = 1.0f; }
        nd->span[k] = nd->hi[k] - nd->lo[k];
    }
    size_t off = 0;
    for (int l = 1; l <= nd->L + 1; l++) {
        nd->offW[l] = (int)off; off += (size_t)nd->fan_in(l) * nd->fan_out(l);
        nd->offB[l] = (int)off; off += (size_t)nd->fan_out(l);
    }
    nd->P = (int)off;
    nd->PP = (int)((off + 3 + 31) & ~(size_t)31);
    const float k = cfg->pde_param != 0.0f ? cfg->pde_param : 1.0f;
    nd->k2 = k * k;
    return true;
}

extern "C" {

int pinn_config_default(pinn_config* cfg, int which) {
    if (!cfg || which < 1 || which > 5) return PINN_ERR_INVALID_ARGUMENT;
    memset(cfg, 0, sizeof(*cfg));
    cfg->abi_version = PINN_ABI_VERSION;
    static const int pde[6] = {0, PINN_PDE_BURGERS, PINN_PDE_BURGERS, PINN_PDE_HELMHOLTZ, PINN_PDE_ALLEN_CAHN, PINN_PDE_NAVIER_STOKES};
    static const int din[6] = {0, 2, 2, 2, 3, 3}, dout[6] = {0, 1, 1, 1, 1, 3};
    static const int width[6] = {0, 20, 20, 64, 128, 256}, depth[6] = {0, 8, 8, 4, 6, 8};
    static const uint64_t nf[6] = {0, 4096, 25600, 1ull << 20, 1ull << 24, 1ull << 26};
    cfg->pde = pde[which]; cfg->d_in = din[which]; cfg->d_out = dout[which];
    cfg->width = width[which]; cfg->depth = depth[which];
    cfg->precision = PINN_PREC_FP32; cfg->device = 0; cfg->rank = 0; cfg->nranks = 1;
    cfg->n_interior = nf[which];
    const uint64_t nb = nf[which] / 32 > 256 ? nf[which] / 32 : 256;
    cfg->n_boundary = nb;
    cfg->n_initial = (cfg->pde == PINN_PDE_HELMHOLTZ) ? 0 : nb;
    cfg->seed_interior = 0x5EED0001ull; cfg->seed_boundary = 0x5EED0002ull; cfg->seed_initial = 0x5EED0003ull;
    cfg->seed_weights = 0x5EED0100ull;
    cfg->lr = 1e-3f; cfg->beta1 = 0.9f; cfg->beta2 = 0.999f; cfg->eps = 1e-8f;
    cfg->lambda_r = cfg->lambda_b = cfg->lambda_0 = 1.0f;
    cfg->pde_param = 0.0f;
    return PINN_OK;
}

int pinn_num_channels(const pinn_config* cfg) {
    NetDims nd; std::string why;
    if (!cfg || !make_dims(cfg, &nd, why)) return -1;
    return nd.C;
}

size_t pinn_num_params(const pinn_config* cfg) {
    NetDims nd; std::string why;
    if (!cfg || !make_dims(cfg, &nd, why)) return 0;
    return (size_t)nd.P;
}

void pinn_shard_range(uint64_t n, int32_t rank, int32_t nranks, uint64_t* begin, uint64_t* end) {
    *begin = (uint64_t)(((unsigned __int128)n * (unsigned)rank) / (unsigned)nranks);
    *end = (uint64_t)(((unsigned __int128)n * (unsigned)(rank + 1)) / (unsigned)nranks);
}

const char* pinn_last_error(const pinn_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

}  // extern "C"

// ---- kernel selection -------------------------------------------------------------------
// Launch bounds per variant: narrow nets (UT = 4, width <= 28, <= 224 threads) are compiled for
// <= 64 registers so that 6 CTAs of 160 threads are resident per SM (C2: all 800 tiles in one wave).
template <int DIN, int N2, bool FULL>
static fused_fn pick_ut_tp(int UT, int TP) {
    if (UT == 4 && TP == 32) return fused_step_fp32<DIN, N2, FULL, 4, 32, 256, 4>;
    if (UT == 8 && TP == 32) return fused_step_fp32<DIN, N2, FULL, 8, 32, 512, 1>;
    if (UT == 8 && TP == 16) return fused_step_fp32<DIN, N2, FULL, 8, 16, 512, 1>;
    if constexpr (!FULL) {      // value-only kernel of wide nets: 32 units per thread (32 FFMAs per 1 LDS + 8 LDG.128)
        if (UT == 32 && TP == 32) return fused_step_fp32<DIN, N2, FULL, 32, 32, 256, 1>;
    }
    return nullptr;
}
static fused_fn pick_kernel(int d_in, int n2, bool full, int UT, int TP, int H) {
    // width-specialised variants of the headline network (2 -> [20 x L] -> 1, BASELINE C1 / C2)
    if (H == 20 && UT == 4 && TP == 32 && d_in == 2) {
        if (full && n2 == 1) return fused_step_fp32<2, 1, true, 4, 32, 160, 6, 20>;
        if (!full) return fused_step_fp32<2, 0, false, 4, 32, 160, 6, 20>;
    }
    if (full) {
        if (d_in == 2 && n2 == 1) return pick_ut_tp<2, 1, true>(UT, TP);
        if (d_in == 2 && n2 == 2) return pick_ut_tp<2, 2, true>(UT, TP);
        if (d_in == 3 && n2 == 2) return pick_ut_tp<3, 2, true>(UT, TP);
    } else {
        if (d_in == 2) return pick_ut_tp<2, 0, false>(UT, TP);
        if (d_in == 3) return pick_ut_tp<3, 0, false>(UT, TP);
    }
    return nullptr;
}

static int configure_kernels(pinn_ctx* c) {
    const NetDims& nd = c->nd;
    cudaDeviceProp prop;
    CUDA_TRY(c, "pinn_create", cudaGetDeviceProperties(&prop, c->cfg.device));
    c->num_sms = prop.multiProcessorCount;
    const int max_smem = (int)prop.sharedMemPerBlockOptin;
    const int H = nd.H, rows = (H + 3) & ~3;
    c->UT = (H % 8 == 0 && H >= 32) ? 8 : 4;
    if (c->UT == 4 && H > 28) return fail(c, PINN_ERR_UNSUPPORTED, "pinn_create: widths above 28 must be multiples of 8");
    auto smem_for = [&](int nc, int tp) { return (int)(2 * (size_t)rows * (nc * tp + 4) * sizeof(float)); };
    c->TP = 32;
    if (smem_for(nd.C, 32) > max_smem || 32 * (H / c->UT) > 512) {
        c->TP = 16;
        if (c->UT != 8) return fail(c, PINN_ERR_UNSUPPORTED, "pinn_create: width needs the 16-point tile, which requires width % 8 == 0");
        if (smem_for(nd.C, 16) > max_smem || 16 * (H / c->UT) > 512)
            return fail(c, PINN_ERR_UNSUPPORTED, "pinn_create: width too large for the FP32 fused kernel");
    }
    c->threads = c->TP * (H / c->UT);
    // split-dW scratch ([threads][16] floats) only where it is cheap: narrow nets, whose 4x4 dW
    // micro-tiles are fewer than the threads of the CTA
    c->nsplit_cap = (H <= 64) ? 8 : 1;
    const int scratch = c->nsplit_cap > 1 ? c->threads * 16 * (int)sizeof(float) : 0;
    const int wstage = c->UT == 4 ? 2 * rows * rows * (int)sizeof(float) : 0;     // cp.async weight double buffer
    c->smem_full = smem_for(nd.C, c->TP) + scratch + wstage;
    c->val_UT = c->UT; c->val_TP = c->TP; c->val_threads = c->threads;
    int scratch_val = scratch;
    if (H >= 64 && H % 32 == 0) {                         // wide nets: one channel only, so take 32 units per thread
        c->val_UT = 32; c->val_TP = 32; c->val_threads = 32 * (H / 32);
        scratch_val = c->nsplit_cap > 1 ? c->val_threads * 16 * (int)sizeof(float) : 0;
    }
    c->smem_val = smem_for(1, c->val_TP) + scratch_val + wstage;
    c->kern_full = pick_kernel(nd.d_in, nd.n2, true, c->UT, c->TP, H);
    c->kern_val = pick_kernel(nd.d_in, nd.n2, false, c->val_UT, c->val_TP, H);
    if (!c->kern_full || !c->kern_val) return fail(c, PINN_ERR_UNSUPPORTED, "pinn_create: no kernel for this configuration");
    CUDA_TRY(c, "pinn_create", cudaFuncSetAttribute(c->kern_full, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_full));
    CUDA_TRY(c, "pinn_create", cudaFuncSetAttribute(c->kern_val, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_val));
    int occ_f = 0, occ_v = 0;
    CUDA_TRY(c, "pinn_create", cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_f, c->kern_full, c->threads, c->smem_full));
    CUDA_TRY(c, "pinn_create", cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_v, c->kern_val, c->val_threads, c->smem_val));
    if (occ_f < 1 || occ_v < 1) return fail(c, PINN_ERR_UNSUPPORTED, "pinn_create: fused kernel does not fit on an SM");
    if (occ_v > 4) occ_v = 4;
    c->grid_cap_full = occ_f * c->num_sms;
    c->grid_cap_val = occ_v * c->num_sms;
    c->kernel_id = (c->UT == 8 ? 10 : 0) + (c->TP == 16 ? 1 : 0);
    // kernels of one step run concurrently (interior || boundary/initial): give them the SAME shared-memory carve-out, or an
    // SM configured for one cannot take CTAs of the other until it drains (measured: the value launch serialised, +30 us)
    CUDA_TRY(c, "pinn_create", cudaFuncSetAttribute(c->kern_full, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CUDA_TRY(c, "pinn_create", cudaFuncSetAttribute(c->kern_val, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    if (c->cfg.precision != PINN_PREC_FP32 && H == 20) {
        // narrow net: register-chained mma.sync kernel for the interior set (boundary / initial points keep the FP32 kernel)
        if (c->cfg.precision == PINN_PREC_BF16 || nd.d_in != 2 || nd.n2 != 1 || nd.d_out != 1)
            return fail(c, PINN_ERR_UNSUPPORTED, "pinn_create: at width 20 the tensor-core path is PINN_PREC_BF16X3, PINN_PREC_BF16X6 or PINN_PREC_FP16X3 for the 2 -> [20 x L] -> 1 Burgers network");
        c->mma_terms = c->cfg.precision == PINN_PREC_BF16X6 ? 3 : 2;
        c->mma_f16 = c->cfg.precision == PINN_PREC_FP16X3 ? 1 : 0;
        c->kern_mma = c->mma_f16 ? fused_step_mma<2, 1, 20, 2, true> : (c->mma_terms == 3 ? fused_step_mma<2, 1, 20, 3> : fused_step_mma<2, 1, 20, 2>);
        c->mma_smem = kMmaWarps * 2 * nd.C * c->mma_terms * 768;
        CUDA_TRY(c, "pinn_create", cudaFuncSetAttribute(c->kern_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, c->mma_smem));
        CUDA_TRY(c, "pinn_create", cudaFuncSetAttribute(c->kern_mma, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        int occ = 0;
        CUDA_TRY(c, "pinn_create", cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, c->kern_mma, kMmaWarps * 32, c->mma_smem));
        if (occ < 1) return fail(c, PINN_ERR_UNSUPPORTED, "pinn_create: mma kernel does not fit on an SM");
        c->grid_cap_mma = occ * c->num_sms;
        c->kernel_id = c->mma_f16 ? 204 : 200 + c->mma_terms;
        return PINN_OK;
    }
    if (c->cfg.precision == PINN_PREC_BF16X6 || c->cfg.precision == PINN_PREC_FP16X3)
        return fail(c, PINN_ERR_UNSUPPORTED, "pinn_create: PINN_PREC_BF16X6 / PINN_PREC_FP16X3 are built for width 20 (the 2 -> [20 x L] -> 1 Burgers network)");
    if (c->cfg.precision != PINN_PREC_FP32) {
        // tensor-core kernel for the interior set; boundary / initial points keep the FP32 kernel
        const int nterms = c->cfg.precision == PINN_PREC_BF16X3 ? 2 : 1;
        if (!(H == 64 || H == 128 || H == 256))
            return fail(c, PINN_ERR_UNSUPPORTED, "pinn_create: the tensor-core path is built for width 20 (BF16X3), 64, 128 or 256 (use PINN_PREC_FP32)");
        const int ut = 8;
        c->tc_ut = ut;
#define PICK_TC_H(D, N, HH) (nterms == 2 ? fused_step_tc<D, N, 2, 8, HH> : fused_step_tc<D, N, 1, 8, HH>)
#define PICK_TC(D, N) (H == 64 ? PICK_TC_H(D, N, 64) : (H == 128 ? PICK_TC_H(D, N, 128) : fused_step_tc<D, N, 1, 8, 256>))
        if (nterms == 2 && H == 256)
            return fail(c, PINN_ERR_UNSUPPORTED, "pinn_create: PINN_PREC_BF16X3 needs width <= 128 (operand terms do not fit in shared memory)");
        if (nd.d_in == 2 && nd.n2 == 1) c->kern_tc = PICK_TC(2, 1);
        else if (nd.d_in == 2 && nd.n2 == 2) c->kern_tc = PICK_TC(2, 2);
        else c->kern_tc = PICK_TC(3, 2);
#undef PICK_TC
#undef PICK_TC_H
        // boundary / initial sets on the tensor cores too (128-point value-only tiles); PINN_TC_VAL=0 keeps them on FP32
        const char* ev = getenv("PINN_TC_VAL");
        if (!(ev && atoi(ev) == 0)) {
#define PICK_TCV(D) (H == 64 ? (nterms == 2 ? fused_step_tc<D, 0, 2, 8, 64, true> : fused_step_tc<D, 0, 1, 8, 64, true>) \
                   : (H == 128 ? (nterms == 2 ? fused_step_tc<D, 0, 2, 8, 128, true> : fused_step_tc<D, 0, 1, 8, 128, true>) \
                               : fused_step_tc<D, 0, 1, 8, 256, true>))
            c->kern_tc_val = nd.d_in == 2 ? PICK_TCV(2) : PICK_TCV(3);
#undef PICK_TCV
        }
        const int HW = H > 128 ? 128 : H;
        c->nterms = nterms;
        c->tc_threads = 16 * H / c->tc_ut;
        c->tc_smem = nterms * (2 * H * 256 + HW * H * 2);
        if (c->tc_smem > max_smem)
            return fail(c, PINN_ERR_UNSUPPORTED, "pinn_create: PINN_PREC_BF16X3 needs width <= 128 (operand terms do not fit in shared memory)");
        c->tc_tmem_cols = (H == 256) ? 512 : 2 * H;
        CUDA_TRY(c, "pinn_create", cudaFuncSetAttribute(c->kern_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, c->tc_smem));
        if (c->kern_tc_val)
            CUDA_TRY(c, "pinn_create", cudaFuncSetAttribute(c->kern_tc_val, cudaFuncAttributeMaxDynamicSharedMemorySize, c->tc_smem));
        // CTAs per SM.  cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for the tcgen05 kernels whatever they need (measured on
        // B200, CUDA 12.9: 128 threads x 128 registers and 40 KB still give 1), which left this kernel at a half or a quarter of its
        // residency through round 1 and most of round 2 (C3: 72.8 -> 116 Mpts/s with the two CTAs per SM it has room for).  So the
        // limits are taken from the function and device attributes: registers, shared memory (+ the per-block reservation), threads,
        // and the TMEM columns each CTA allocates.
        int occ = 0;
        {
            auto fit = [&](pinn_ctx::tc_fn k, int* out) -> int {
                cudaFuncAttributes fa{};
                if (cudaFuncGetAttributes(&fa, reinterpret_cast<const void*>(k)) != cudaSuccess) return PINN_ERR_CUDA;
                int regs_sm = 0, smem_sm = 0, smem_res = 0, thr_sm = 0;
                cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, c->cfg.device);
                cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, c->cfg.device);
                cudaDeviceGetAttribute(&smem_res, cudaDevAttrReservedSharedMemoryPerBlock, c->cfg.device);
                cudaDeviceGetAttribute(&thr_sm, cudaDevAttrMaxThreadsPerMultiProcessor, c->cfg.device);
                const int warps = (c->tc_threads + 31) / 32;
                const int regs_cta = ((fa.numRegs * 32 + 255) / 256 * 256) * warps;          // allocated per warp in units of 256
                const int smem_cta = c->tc_smem + (int)fa.sharedSizeBytes + smem_res;
                int o = regs_cta > 0 ? regs_sm / regs_cta : 0;
                if (smem_cta > 0 && smem_sm / smem_cta < o) o = smem_sm / smem_cta;
                if (thr_sm / c->tc_threads < o) o = thr_sm / c->tc_threads;
                *out = o;
                return PINN_OK;
            };
            int o1 = 0, o2 = 1 << 30;
            if (fit(c->kern_tc, &o1)) return fail(c, PINN_ERR_CUDA, "pinn_create: cudaFuncGetAttributes failed");
            if (c->kern_tc_val && fit(c->kern_tc_val, &o2)) return fail(c, PINN_ERR_CUDA, "pinn_create: cudaFuncGetAttributes failed");
            occ = o1 < o2 ? o1 : o2;
        }
        if (occ < 1) return fail(c, PINN_ERR_UNSUPPORTED, "pinn_create: tensor-core kernel does not fit on an SM");
        const int tmem_limit = 512 / c->tc_tmem_cols;
        if (getenv("PINN_DEBUG_OCC")) fprintf(stderr, "[pinn] tc kernel: occupancy %d by registers / shared memory (%d threads, %d B), TMEM allows %d\n", occ, c->tc_threads, c->tc_smem, tmem_limit);
        if (occ > tmem_limit) occ = tmem_limit;
        if (const char* eo = getenv("PINN_TC_OCC")) { const int o = atoi(eo); if (o >= 1 && o <= tmem_limit) occ = o; }
        c->grid_cap_tc = occ * c->num_sms;
        c->kernel_id = 100 + nterms;
        // Training launches of the six-channel nets at width 128 / 256 (BASELINE C4 / C5): producer + dW-server kernel,
        // one cooperative launch with one CTA per SM.  PINN_TC2=0 keeps the round-1 kernel (it still serves evaluation).
        {
            const char* e2 = getenv("PINN_TC2");
            const bool want = !(e2 && atoi(e2) == 0);
            if (want && nterms == 1 && (H == 128 || H == 256) && nd.d_in == 3 && nd.n2 == 2 && nd.L >= 2) {
                // Width 128 (tc2 only): two CTAs per SM — 2 x 107 KB of shared memory, 2 x 256 TMEM columns, 2 x 256 threads x 128 registers —
                // so that one CTA's UMMAs run under the other's epilogue.  (A plain launch then: the cooperative-launch check uses the
                // occupancy calculator that answers 1 for tcgen05 kernels.  All CTAs are resident because nothing else runs on the device
                // while a step does — and every hand-off wait is bounded and reports its site instead of hanging.)
                const char* e3a = getenv("PINN_TC3");
                const char* e2c = getenv("PINN_TC2_CTAS");
                int ctas = (H == 128 && !(e3a && atoi(e3a) != 0)) ? 2 : 1;
                if (e2c && atoi(e2c) == 1) ctas = 1;
                const int grid2 = ctas * c->num_sms;
                int q = H == 256 ? 3 : (ctas == 2 ? 10 : 7);    // dW servers per hidden->hidden layer (measured best; C5 full size, same box: q = 3 -> 30.7, 4 -> 28.7 Mpts/s, 2 saturates the servers: 21.1; C4: q = 8..12 -> 92, 97, 101-104, 102, 100)
                if (const char* eq = getenv("PINN_TC2_Q")) q = atoi(eq);
                if (q < 1) q = 1;
                while (q > 1 && grid2 - (nd.L - 1) * q < grid2 / 2) q--;
                const int nprod = grid2 - (nd.L - 1) * q;
                const int smem2 = tc2_smem_bytes(H, nd.L);
                int coop = 0;
                CUDA_TRY(c, "pinn_create", cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->cfg.device));
                if (coop && nprod >= 8 && smem2 <= max_smem) {
                    // PINN_TC3=1: the experimental producer with two sub-tiles in flight (pinn_fused_tc3.cuh; measured no faster)
                    const char* e3 = getenv("PINN_TC3");
                    const bool tc3on = e3 && atoi(e3) != 0;
                    if (tc3on) { c->kern_tc2 = H == 128 ? fused_step_tc3<3, 2, 128> : fused_step_tc3<3, 2, 256>; c->tc2_T = tc3::T; c->tc2_threads = 2 * H + tc3::kHelperThreads; }
                    else { c->kern_tc2 = H == 128 ? fused_step_tc2<3, 2, 128> : fused_step_tc2<3, 2, 256>; c->tc2_T = tc2::T; c->tc2_threads = 2 * H; }
                    CUDA_TRY(c, "pinn_create", cudaFuncSetAttribute(c->kern_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
                    int occ2 = 0;
                    CUDA_TRY(c, "pinn_create", cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, c->kern_tc2, c->tc2_threads, smem2));
                    if (occ2 < 1) c->kern_tc2 = nullptr;
                    else {
                        c->tc2_q = q; c->tc2_nprod = nprod; c->tc2_smem = smem2; c->tc2_grid = grid2;
                        if (const char* ez = getenv("PINN_TC2_ZD")) c->tc2_zdepth = atoi(ez);
                        if (c->tc2_zdepth < 2) c->tc2_zdepth = 2;
                        c->kernel_id = (tc3on ? 120 : 110) + nterms;
                    }
                }
            }
        }
        if (H == 256 && nterms == 1 && nd.L >= 2 && !c->kern_tc2) {
            c->tc_defer_g = 16;
            if (const char* e = getenv("PINN_TC_DEFER")) c->tc_defer_g = atoi(e);
            if (c->tc_defer_g < 0 || c->tc_defer_g > 64) c->tc_defer_g = 0;
            c->ring_per_cta = (long long)c->tc_defer_g * (nd.L - 1) * 2 * (H / 8) * (nd.C * 16) * 16;
        }
    }
    return PINN_OK;
}

// {beta1^t, beta2^t} as the ITERATED product the tail kernel forms (bit-identical, so a resumed run matches an uninterrupted one)
static int upload_beta_pow(pinn_ctx* c, uint64_t t) {
    double p[2] = {1.0, 1.0};
    const double b1 = (double)c->cfg.beta1, b2 = (double)c->cfg.beta2;
    for (uint64_t k = 0; k < t; k++) { p[0] *= b1; p[1] *= b2; }
    if (cudaMemcpyAsync(c->d_beta_pow, p, sizeof(p), cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess)
        return fail(c, PINN_ERR_CUDA, "pinn_set_optimizer: upload of the Adam bias powers failed");
    return PINN_OK;
}

static int tiles_of(uint64_t n, int tp) { return (int)((n + tp - 1) / tp); }

static void host_init_params(const NetDims& nd, uint64_t seed, std::vector<float>& w) {
    w.assign(nd.P, 0.0f);
    for (int l = 1; l <= nd.L + 1; l++) {
        const int fi = nd.fan_in(l), fo = nd.fan_out(l);
        const float a = (float)std::sqrt(6.0 / (double)(fi + fo));
        const float two_a = 2.0f * a;
        for (size_t k = 0; k < (size_t)fi * fo; k++) w[nd.offW[l] + k] = fmaf(two_a, u01(seed + (uint64_t)l, k), -a);
    }
}

// One fused launch.  `row0` is the first partial/stash row it may use; returns the grid in *grid_out.
static int launch_fused(pinn_ctx* c, cudaStream_t stream, bool full, const PointSet* sets, int nsets, int mode,
                        float* y_out, float* r_out, int row0, int* grid_out) {
    if (grid_out) *grid_out = 0;
    int total = 0;
    for (int i = 0; i < nsets; i++) total += sets[i].ntiles;
    if (total == 0) return PINN_OK;
    FusedArgs a{};
    a.params = c->d_params; a.paramsT = c->d_paramsT;
    a.nsets = nsets;
    for (int i = 0; i < nsets; i++) a.sets[i] = sets[i];
    a.partial = c->d_partial; a.stash = c->d_stash; a.row0 = row0; a.nsplit_cap = c->nsplit_cap;
    a.stash_per_cta = c->stash_per_cta;
    a.mode = mode; a.y_out = y_out; a.r_out = r_out;
    const bool tcv = !full && mode == 0 && c->kern_tc_val != nullptr;      // value-only tiles of 128 points
    const bool tc = (full && c->kern_tc != nullptr) || tcv;
    if (tc) {                                             // 16-point (value-only: 128-point) tiles on the tensor-core kernel
        total = 0;
        for (int i = 0; i < nsets; i++) { a.sets[i].ntiles = tiles_of(a.sets[i].n, tcv ? 128 : 16); total += a.sets[i].ntiles; }
    }
    const int cap = tc ? c->grid_cap_tc : (full ? c->grid_cap_full : c->grid_cap_val);
    (void)tcv;
    int grid = total < cap ? total : cap;
    const bool timed = c->prof && full && mode == 0 && !c->capturing;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (timed) {
        CUDA_TRY(c, "pinn_train_step", cudaEventCreate(&e0)); CUDA_TRY(c, "pinn_train_step", cudaEventCreate(&e1));
        CUDA_TRY(c, "pinn_train_step", cudaEventRecord(e0, stream));
    }
    if (tc) {
        TcArgs ta{};
        ta.f = a; ta.wq = c->d_wq; ta.wq_term_stride = c->wq_term_stride; ta.tmem_cols = c->tc_tmem_cols;
        ta.op_ring = c->d_ring; ta.ring_per_cta = c->ring_per_cta; ta.defer_g = (mode == 0 && !tcv) ? c->tc_defer_g : 0;
        (tcv ? c->kern_tc_val : c->kern_tc)<<<grid, c->tc_threads, c->tc_smem, stream>>>(c->nd, ta);
    } else {
        if (full) c->kern_full<<<grid, c->threads, c->smem_full, stream>>>(c->nd, a);
        else c->kern_val<<<grid, c->val_threads, c->smem_val, stream>>>(c->nd, a);
    }
    if (timed) { cudaEventRecord(e1, stream); c->pending.emplace_back(e0, e1); }
    if (full) c->last_grid = grid;
    if (grid_out) *grid_out = grid;                       // partial rows this launch used
    c->step_launches++;
    CUDA_TRY(c, "pinn_train_step", cudaGetLastError());
    return PINN_OK;
}

// Width-20 mma.sync kernel: ONE launch walks the 16-point tiles of all three sets (interior tiles with every Taylor channel,
// boundary / initial tiles value-only).  A second concurrent kernel would take resident-CTA slots away from it: its 400 CTAs
// (C2) need 400 of the 444 register-limited slots, so any co-resident CTA pushes some of them into a second wave (+30 us).
static int launch_mma(pinn_ctx* c, const PointSet& interior, const PointSet* vsets, int nv, int* rows_out) {
    MmaArgs ma{};
    ma.params = c->d_params; ma.wfrag = c->d_wfrag; ma.set = interior; ma.set.ntiles = tiles_of(interior.n, 16);
    ma.nvsets = nv;
    int total = ma.set.ntiles;
    for (int i = 0; i < nv; i++) { ma.vsets[i] = vsets[i]; ma.vsets[i].ntiles = tiles_of(vsets[i].n, 16); total += ma.vsets[i].ntiles; }
    ma.partial = c->d_partial; ma.stash = c->d_stash; ma.row0 = 0; ma.stash_per_row = c->stash_per_cta;
    ma.x_copy = (interior.x != c->d_x[0]) ? c->d_x[0] : nullptr;      // points come from mapped host memory: keep the HBM copy current
    *rows_out = 0;
    if (total == 0) return PINN_OK;
    const int ctas = (total + kMmaWarps - 1) / kMmaWarps;
    const int grid = ctas < c->grid_cap_mma ? ctas : c->grid_cap_mma;
    const bool timed = c->prof && !c->capturing;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (timed) {
        CUDA_TRY(c, "pinn_train_step", cudaEventCreate(&e0)); CUDA_TRY(c, "pinn_train_step", cudaEventCreate(&e1));
        CUDA_TRY(c, "pinn_train_step", cudaEventRecord(e0, c->stream));
    }
    c->kern_mma<<<grid, kMmaWarps * 32, c->mma_smem, c->stream>>>(c->nd, ma);
    if (timed) { cudaEventRecord(e1, c->stream); c->pending.emplace_back(e0, e1); }
    c->last_grid = grid;
    *rows_out = grid * kMmaWarps < total ? grid * kMmaWarps : total;   // one row per warp that owns at least one tile
    c->step_launches++;
    CUDA_TRY(c, "pinn_train_step", cudaGetLastError());
    return PINN_OK;
}

// Producer / dW-server tensor-core kernel: one cooperative launch, one CTA per SM (the CTAs hand tiles to one another
// through global flags, so all of them must be resident).  Uses partial rows [0, num_sms).
static int launch_tc2(pinn_ctx* c, const PointSet& interior, int* rows_out) {
    *rows_out = 0;
    if (interior.n == 0) return PINN_OK;
    Tc2Args a{};
    a.params = c->d_params; a.set = interior; a.set.ntiles = tiles_of(interior.n, c->tc2_T);
    a.partial = c->d_partial; a.row0 = 0;
    a.wf = c->d_wf; a.wb = c->d_wb; a.stash_a = c->d_st_a; a.stash_s = c->d_st_s; a.zring = c->d_zring;
    a.ready = c->d_tc2_flags; a.done = c->d_tc2_flags + (size_t)c->tc2_nprod * kFlagStride;
    a.stagger = 4;
    { const char* eb = getenv("PINN_TC2_BULK"); a.bulk_store = (eb && atoi(eb) != 0) ? 1 : 0; }
    if (const char* es = getenv("PINN_TC3_STAGGER")) a.stagger = atoi(es);
    if (c->tc2_T != tc3::T) { const char* el = getenv("PINN_TC2_LEAN"); a.stagger = (el && atoi(el) == 0) ? -1 : 0; }
    a.ready2 = c->tc2_T == tc3::T ? c->d_tc2_flags + 2 * (size_t)c->tc2_nprod * kFlagStride : nullptr;
    a.nprod = c->tc2_nprod; a.q = c->tc2_q; a.zdepth = c->tc2_zdepth; a.dbg = c->d_tc2_dbg;
    a.trap_site = reinterpret_cast<int*>(c->d_hrec + 16);
    CUDA_TRY(c, "pinn_train_step", cudaMemsetAsync(c->d_tc2_flags, 0, 3 * (size_t)c->tc2_nprod * kFlagStride * sizeof(unsigned int), c->stream));
    const bool timed = c->prof && !c->capturing;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (timed) {
        CUDA_TRY(c, "pinn_train_step", cudaEventCreate(&e0)); CUDA_TRY(c, "pinn_train_step", cudaEventCreate(&e1));
        CUDA_TRY(c, "pinn_train_step", cudaEventRecord(e0, c->stream));
    }
    NetDims nd = c->nd;
    void* kargs[] = {&nd, &a};
    if (c->tc2_grid == c->num_sms)
        CUDA_TRY(c, "pinn_train_step", cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(c->kern_tc2), dim3(c->tc2_grid),
                                                                   dim3(c->tc2_threads), kargs, (size_t)c->tc2_smem, c->stream));
    else
        CUDA_TRY(c, "pinn_train_step", cudaLaunchKernel(reinterpret_cast<const void*>(c->kern_tc2), dim3(c->tc2_grid), dim3(c->tc2_threads), kargs,
                                                        (size_t)c->tc2_smem, c->stream));
    if (timed) { CUDA_TRY(c, "pinn_train_step", cudaEventRecord(e1, c->stream)); c->pending.emplace_back(e0, e1); }
    c->last_grid = c->tc2_grid;
    *rows_out = c->tc2_grid;
    c->step_launches++;
    return PINN_OK;
}

// Local part of a step: interior launch on the main stream, boundary + initial launch concurrently
// on the auxiliary stream (disjoint partial rows), then the fixed-order reduction of all rows.
static int run_local(pinn_ctx* c, int bump_step, bool do_reduce = true) {
    PointSet v[2]; int nv = 0;
    for (int k = 1; k <= 2; k++) {
        if (c->n_local[k] == 0) continue;
        v[nv].x = c->d_x[k]; v[nv].n = c->n_local[k]; v[nv].scale = c->scale[k]; v[nv].set = k;
        v[nv].ntiles = tiles_of(c->n_local[k], c->val_TP); nv++;
    }
    const bool mk = c->kern_mma != nullptr;
    const bool t2 = c->kern_tc2 != nullptr;              // cooperative launch on every SM: the value-only launch follows it on the same stream
    if (nv && !mk && !t2) {
        CUDA_TRY(c, "pinn_train_step", cudaEventRecord(c->ev_fork, c->stream));
        CUDA_TRY(c, "pinn_train_step", cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
    }
    PointSet s{};
    s.x = (mk && c->x_override) ? c->x_override : c->d_x[0];
    s.n = c->n_local[0]; s.scale = c->scale[0]; s.set = PINN_SET_INTERIOR; s.ntiles = tiles_of(s.n, c->TP);
    int grid_a = 0, grid_b = 0;
    int rc = mk ? launch_mma(c, s, v, nv, &grid_a)
                : (t2 ? launch_tc2(c, s, &grid_a) : launch_fused(c, c->stream, true, &s, 1, 0, nullptr, nullptr, 0, &grid_a));
    if (rc) return rc;
    if (nv && !mk) {
        const int row0 = c->rows_cap_a;                   // rows [0, rows_cap_a) belong to the interior launch
        rc = launch_fused(c, t2 ? c->stream : c->stream2, false, v, nv, 0, nullptr, nullptr, row0, &grid_b);
        if (rc) return rc;
        if (!t2) {
            CUDA_TRY(c, "pinn_train_step", cudaEventRecord(c->ev_join, c->stream2));
            CUDA_TRY(c, "pinn_train_step", cudaStreamWaitEvent(c->stream, c->ev_join, 0));
        }
    }
    c->rows_a = grid_a; c->rows_b = grid_b;
    if (!do_reduce) return PINN_OK;
    const int n_out = c->nd.P + 3;
    reduce_partials_kernel<<<(n_out + 31) / 32, dim3(32, 32), 0, c->stream>>>(
        c->d_partial, grid_a, c->d_partial + (size_t)c->rows_cap_a * c->nd.PP, grid_b, c->nd.PP, n_out, c->d_gbuf, c->d_step, bump_step,
        c->kern_mma ? 0 : 1);
    c->step_launches++;
    CUDA_TRY(c, "pinn_train_step", cudaGetLastError());
    return PINN_OK;
}

static void refresh_wq(pinn_ctx* c, bool after_adam = false);
static FragMirror frag_mirror(const pinn_ctx* c) {
    FragMirror fm{};
    if (c->d_wfrag && c->nd.L >= 2) { fm.w16 = reinterpret_cast<unsigned short*>(c->d_wfrag); fm.nterms = c->mma_terms; fm.f16 = c->mma_f16; }
    return fm;
}

// Single-GPU tail of the step: reduce + Adam + mirrors + loss record in one launch.
static int run_tail(pinn_ctx* c) {
    const int n_out = c->nd.P + 3;
    reduce_adam_kernel<<<(n_out + 31) / 32, dim3(8, kTailLanes), 0, c->stream>>>(
        c->nd, c->adam, c->d_partial, c->rows_a, c->kern_mma ? 0 : 1, c->d_partial + (size_t)c->rows_cap_a * c->nd.PP, c->rows_b,
        c->d_params, c->d_paramsT, c->d_m, c->d_v, c->d_gbuf, c->d_step, c->d_loss, c->d_flag, c->d_tail, frag_mirror(c), c->d_hrec,
        c->d_beta_pow);
    c->step_launches++;
    refresh_wq(c, true);
    CUDA_TRY(c, "pinn_train_step", cudaGetLastError());
    return PINN_OK;
}

// Fused reduce + peer-memory all-reduce + Adam (see reduce_exchange_adam_kernel).
static int run_exchange_apply(pinn_ctx* c) {
    P2PArgs x{};
    for (int r = 0; r < c->cfg.nranks; r++) x.peers[r] = c->peer_ptr[r];
    x.nranks = c->cfg.nranks; x.rank = c->cfg.rank; x.seq = ++c->xseq;
    x.done_counter = c->d_done; x.err_flag = c->d_flag;
    const int n_out = c->nd.P + 3;
    reduce_exchange_adam_kernel<<<(n_out + 31) / 32, dim3(32, 32), 0, c->stream>>>(
        c->nd, c->adam, c->d_partial, c->rows_a, c->d_partial + (size_t)c->rows_cap_a * c->nd.PP, c->rows_b,
        c->d_params, c->d_paramsT, c->d_m, c->d_v, c->d_gbuf, c->d_step, (unsigned long long)(c->host_step + 1),
        c->d_loss, c->d_flag, x, frag_mirror(c), c->d_hrec, c->d_beta_pow);
    c->step_launches++;
    refresh_wq(c, true);
    CUDA_TRY(c, "pinn_train_step", cudaGetLastError());
    return PINN_OK;
}

static void refresh_wq(pinn_ctx* c, bool after_adam) {
    if (c->d_wfrag && c->nd.L >= 2 && !after_adam) {       // (the Adam kernels refresh this mirror themselves)
        const int n = (c->nd.L - 1) * 2 * 9 * c->mma_terms * 32;
        pack_wfrag_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->nd, c->d_params, c->d_wfrag, c->mma_terms, c->mma_f16);
        c->step_launches++;
    }
    if (c->d_wf) {
        const long long n2 = (long long)(c->nd.L - 1) * c->nd.H * c->nd.H;
        pack_w2_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, c->stream>>>(c->nd, c->d_params, c->d_wf, c->d_wb, c->tc2_T == tc3::T ? 1 : 0);
        c->step_launches++;
    }
    if (!c->d_wq) return;
    const long long n = c->wq_term_stride;
    pack_wq_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->nd, c->d_params, c->d_wq, c->wq_term_stride, c->nterms);
    c->step_launches++;
}

static int run_allreduce(pinn_ctx* c) {
    if (c->cfg.nranks <= 1 || !c->comm) return PINN_OK;
    const int rc = g_nccl.AllReduce(c->d_gbuf, c->d_gbuf, (size_t)c->nd.P + 3, kNcclFloat32, kNcclSum, c->comm, c->stream);
    if (rc != 0) return fail(c, PINN_ERR_NCCL, std::string("pinn_train_step: ncclAllReduce -> ") +
                                                   (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"));
    return PINN_OK;
}

static int run_apply(pinn_ctx* c, int apply) {
    adam_kernel<<<(c->nd.P + 255) / 256, 256, 0, c->stream>>>(c->nd, c->adam, c->d_params, c->d_paramsT, c->d_m, c->d_v,
                                                            c->d_gbuf, c->d_step, c->d_loss, c->d_flag, apply, frag_mirror(c), c->d_hrec,
                                                            c->d_beta_pow);
    c->step_launches++;
    if (apply) refresh_wq(c, true);
    CUDA_TRY(c, "pinn_train_step", cudaGetLastError());
    return PINN_OK;
}

// All device work of one training step, issued on the handle's streams (directly, or under capture).
static int issue_step(pinn_ctx* c) {
    const bool dist = c->cfg.nranks > 1 && c->comm != nullptr;
    const bool p2p = dist && c->p2p_ready;
    if (!dist) { const int rc0 = run_local(c, 0, false); return rc0 ? rc0 : run_tail(c); }
    int rc = run_local(c, 1, !p2p);
    if (p2p) return rc ? rc : run_exchange_apply(c);
    if (!rc) rc = run_allreduce(c);
    if (!rc) rc = run_apply(c, 1);
    return rc;
}

static void drop_graph(pinn_ctx* c) {
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    if (c->graph) cudaGraphDestroy(c->graph);
    c->graph_exec = nullptr; c->graph = nullptr;
}

// One training step.  The device work that is local to this GPU is replayed from a captured CUDA graph
// (single GPU: interior || boundary/initial -> reduce -> Adam; with a communicator: interior ||
// boundary/initial -> reduce only).  NCCL is never captured: its lazy first-call setup is not
// capturable and a failed capture would leave the ranks out of step, so the all-reduce and the Adam
// kernel that follows it are launched directly.  Per-kernel profiling disables the graph.
static int step_once(pinn_ctx* c) {
    const bool want_graph = c->graph_ok && !c->prof;
    const bool dist = c->cfg.nranks > 1 && c->comm != nullptr;
    if (want_graph) {
        const bool p2p = dist && c->p2p_ready;
        const uint64_t key[4] = {c->n_local[0], c->n_local[1], c->n_local[2], (uint64_t)dist + 2 * (uint64_t)p2p};
        if (!c->graph_exec || memcmp(key, c->graph_key, sizeof(key)) != 0) {
            drop_graph(c);
            c->step_launches = 0; c->capturing = true;
            cudaError_t e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
            int rc = PINN_ERR_CUDA;
            if (e == cudaSuccess) {
                if (!dist) { rc = run_local(c, 0, false); if (!rc) rc = run_tail(c); }     // single GPU: the whole step
                else rc = run_local(c, 1, !p2p);          // with the peer exchange the reduction is part of the fused kernel
            }
            cudaGraph_t g = nullptr;
            if (e == cudaSuccess) e = cudaStreamEndCapture(c->stream, &g);
            c->capturing = false;
            if (e == cudaSuccess && rc == PINN_OK && g) e = cudaGraphInstantiate(&c->graph_exec, g, 0);
            if (e != cudaSuccess || rc != PINN_OK || !c->graph_exec) {
                if (g) cudaGraphDestroy(g);
                const cudaError_t why = e != cudaSuccess ? e : cudaGetLastError();
                cudaGetLastError();
                c->graph_exec = nullptr; c->graph_ok = false;        // direct launches for this handle from now on — and say so:
                c->graph_note = std::string("CUDA graph capture of the step failed (") + (rc != PINN_OK ? c->err : std::string(cudaGetErrorString(why))) +
                                "): steps are launched kernel by kernel";   // pinn_graph_state() reports it
            } else {
                c->graph = g; memcpy(c->graph_key, key, sizeof(key));
                c->launches_per_graph = c->step_launches;
            }
        }
        if (c->graph_exec) {
            CUDA_TRY(c, "pinn_train_step", cudaGraphLaunch(c->graph_exec, c->stream));
            c->launches += c->launches_per_graph;
            if (dist) {
                c->step_launches = 0;
                int rc;
                if (p2p) rc = run_exchange_apply(c);
                else { rc = run_allreduce(c); if (!rc) rc = run_apply(c, 1); }
                c->launches += c->step_launches;
                if (rc) return rc;
            }
            c->host_step++;
            return PINN_OK;
        }
    }
    c->step_launches = 0;
    const int rc = issue_step(c);
    c->launches += c->step_launches;
    if (!rc) c->host_step++;
    return rc;
}

// Device error channel (reference include/device/runtime.cuh:8-48: a device-side code read back by the host and thrown as
// device_exception{code, name}): ANY non-zero flag the step tail left is returned and cleared — not only the overflow code.
static int device_flag_error(pinn_ctx* c, int flag) {
    cudaMemsetAsync(c->d_flag, 0, sizeof(int), c->stream);
    if (flag == PINN_ERR_DEVICE_OVERFLOW)
        return fail(c, PINN_ERR_DEVICE_OVERFLOW, "network::train_step: overflow");   // reference runtime.cuh:17-38 text
    if (flag == PINN_ERR_NCCL)
        return fail(c, PINN_ERR_NCCL, "network::train_step: a peer rank never arrived at the gradient exchange (step not applied)");
    return fail(c, flag, "network::train_step: device error " + std::to_string(flag));
}

static int read_out(pinn_ctx* c, pinn_step_out* out) {
    if (!out) return PINN_OK;
    float* hp = c->h_pinned + 8;                          // copy-path scratch (the first 8 floats are the mapped step record)
    CUDA_TRY(c, "pinn_train_step", cudaMemcpyAsync(hp, c->d_loss, 5 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, "pinn_train_step", cudaStreamSynchronize(c->stream));
    out->loss = hp[0]; out->mse_residual = hp[1];
    out->mse_boundary = hp[2]; out->mse_initial = hp[3];
    out->step = c->host_step;
    int flag; memcpy(&flag, hp + 4, sizeof(int));
    return flag ? device_flag_error(c, flag) : PINN_OK;
}

// Result of the latest TRAINING step without a copy or a stream synchronisation: the step tail wrote the loss record and,
// last, the step number into the mapped pinned record (loss_record); poll that for a bounded time (steps of the narrow
// nets take < 0.1 ms), then fall back to a stream synchronisation so long steps and failed launches behave as usual.
static int read_step(pinn_ctx* c, pinn_step_out* out) {
    if (!out) return PINN_OK;
    volatile unsigned int* w = reinterpret_cast<volatile unsigned int*>(c->h_pinned);
    const unsigned int want = (unsigned int)c->host_step;
    // short steps: poll (bounded, ~0.3 ms); long steps: block in the driver like any other caller, then read the record
    bool ready = false;
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned spins = 0;; spins++) {
        if (w[3] == want && w[6] == want) { ready = true; break; }   // both halves of the record carry this step's number
        if ((spins & 63u) == 63u && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(300)) break;
    }
    if (!ready) {
        CUDA_TRY(c, "pinn_train_step", cudaStreamSynchronize(c->stream));
        if (!(w[3] == want && w[6] == want)) return read_out(c, out);                        // (defensive) take the copy path
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    volatile float* h = c->h_pinned;
    out->loss = h[0]; out->mse_residual = h[1]; out->mse_boundary = h[2]; out->mse_initial = h[4];
    out->step = c->host_step;
    const int flag = (int)w[5];
    return flag ? device_flag_error(c, flag) : PINN_OK;
}

static void drain_profile(pinn_ctx* c) {
    for (auto& pr : c->pending) {
        float ms = 0.0f;
        if (cudaEventSynchronize(pr.second) == cudaSuccess && cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) {
            c->kernel_ms += ms; c->kernel_count++;
        }
        cudaEventDestroy(pr.first); cudaEventDestroy(pr.second);
    }
    c->pending.clear();
}

extern "C" {

int pinn_create(pinn_ctx** out, const pinn_config* cfg) {
    if (!out || !cfg) return fail(nullptr, PINN_ERR_INVALID_ARGUMENT, "pinn_create: null argument");
    *out = nullptr;
    if (cfg->abi_version != PINN_ABI_VERSION) return fail(nullptr, PINN_ERR_INVALID_ARGUMENT, "pinn_create: abi_version mismatch");
    if (cfg->nranks < 1 || cfg->rank < 0 || cfg->rank >= cfg->nranks) return fail(nullptr, PINN_ERR_INVALID_ARGUMENT, "pinn_create: invalid rank / nranks");
    if (cfg->precision < PINN_PREC_FP32 || cfg->precision > PINN_PREC_FP16X3) return fail(nullptr, PINN_ERR_INVALID_ARGUMENT, "pinn_create: unknown precision");
    NetDims nd; std::string why;
    if (!make_dims(cfg, &nd, why)) return fail(nullptr, PINN_ERR_INVALID_ARGUMENT, "pinn_create: " + why);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, PINN_ERR_NO_DEVICE, "pinn_create: no CUDA device (this library has no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, PINN_ERR_INVALID_ARGUMENT, "pinn_create: device ordinal out of range");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess || prop.major != 10)
        return fail(nullptr, PINN_ERR_NO_DEVICE, "pinn_create: device is not sm_100 (Blackwell B200); kernels are built for sm_100a only");

    pinn_ctx* c = new pinn_ctx();
    c->cfg = *cfg; c->nd = nd;
    auto bail = [&](int rc) { g_create_error = c->err; pinn_destroy(c); return rc; };
#define TRY_C(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { c->err = std::string("pinn_create: ") + #call + " -> " + cudaGetErrorString(e_); return bail(PINN_ERR_CUDA); } } while (0)
    TRY_C(cudaSetDevice(cfg->device));
    TRY_C(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    TRY_C(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    TRY_C(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    TRY_C(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    int rc = configure_kernels(c);
    if (rc) return bail(rc);

    const uint64_t counts[3] = {cfg->n_interior, cfg->n_boundary, cfg->n_initial};
    const float lambdas[3] = {cfg->lambda_r, cfg->lambda_b, cfg->lambda_0};
    c->seeds[0] = cfg->seed_interior; c->seeds[1] = cfg->seed_boundary; c->seeds[2] = cfg->seed_initial;
    for (int k = 0; k < 3; k++) {
        uint64_t b, e;
        pinn_shard_range(counts[k], cfg->rank, cfg->nranks, &b, &e);
        c->begin[k] = b; c->cap[k] = e - b; c->n_local[k] = e - b;
        c->scale[k] = counts[k] ? (float)(2.0 * (double)lambdas[k] / (double)counts[k]) : 0.0f;
        if (c->cap[k]) TRY_C(cudaMalloc(&c->d_x[k], c->cap[k] * nd.d_in * sizeof(float)));
    }
    c->adam.lr = cfg->lr; c->adam.beta1 = cfg->beta1; c->adam.beta2 = cfg->beta2; c->adam.eps = cfg->eps;
    c->adam.lambda_r = cfg->lambda_r; c->adam.lambda_b = cfg->lambda_b; c->adam.lambda_0 = cfg->lambda_0;
    c->adam.inv_nf = counts[0] ? 1.0 / (double)counts[0] : 0.0;
    c->adam.inv_nb = counts[1] ? 1.0 / (double)counts[1] : 0.0;
    c->adam.inv_n0 = counts[2] ? 1.0 / (double)counts[2] : 0.0;

    // partial / stash rows: [0, rows_cap_a) interior launch (FP32 or tensor-core kernel, also the
    // evaluation calls), [rows_cap_a, rows_cap_a + rows_cap_b) the concurrent boundary/initial launch
    c->rows_cap_a = c->grid_cap_full > c->grid_cap_tc ? c->grid_cap_full : c->grid_cap_tc;
    if (c->grid_cap_val > c->rows_cap_a) c->rows_cap_a = c->grid_cap_val;      // pinn_forward uses the value kernel on rows A
    if (c->kern_mma && c->grid_cap_mma * kMmaWarps > c->rows_cap_a) c->rows_cap_a = c->grid_cap_mma * kMmaWarps;
    if (c->kern_tc2 && c->tc2_grid > c->rows_cap_a) c->rows_cap_a = c->tc2_grid;
    c->rows_cap_b = c->grid_cap_val > c->grid_cap_tc ? c->grid_cap_val : c->grid_cap_tc;
    const size_t nrows = (size_t)c->rows_cap_a + c->rows_cap_b;
    c->stash_per_cta = (long long)nd.L * nd.C * nd.H * c->TP;   // the tensor-core kernel (16-point tiles) needs no more
    if ((long long)nd.L * nd.H * c->val_TP > c->stash_per_cta) c->stash_per_cta = (long long)nd.L * nd.H * c->val_TP;
    if (c->kern_tc_val && (long long)nd.L * 8 * nd.H * 16 > c->stash_per_cta) c->stash_per_cta = (long long)nd.L * 8 * nd.H * 16;
    if (c->kern_mma && (long long)nd.L * (nd.C * 3 * 128 + 32) > c->stash_per_cta) c->stash_per_cta = (long long)nd.L * (nd.C * 3 * 128 + 32);
    if (c->kern_mma) TRY_C(cudaMalloc(&c->d_wfrag, (size_t)(nd.L > 1 ? nd.L - 1 : 1) * 2 * 27 * 32 * sizeof(uint32_t)));
    const size_t P = nd.P;
    TRY_C(cudaMalloc(&c->d_params, (P + 4) * sizeof(float)));     // +4: kernels copy them in float4 units
    TRY_C(cudaMalloc(&c->d_paramsT, (P + 4) * sizeof(float)));
    TRY_C(cudaMemsetAsync(c->d_params, 0, (P + 4) * sizeof(float), c->stream));
    TRY_C(cudaMemsetAsync(c->d_paramsT, 0, (P + 4) * sizeof(float), c->stream));
    TRY_C(cudaMalloc(&c->d_m, P * sizeof(float)));
    TRY_C(cudaMalloc(&c->d_v, P * sizeof(float)));
    TRY_C(cudaMalloc(&c->d_gbuf, (size_t)nd.PP * sizeof(float)));
    TRY_C(cudaMalloc(&c->d_partial, nrows * nd.PP * sizeof(float)));
    TRY_C(cudaMalloc(&c->d_stash, nrows * c->stash_per_cta * sizeof(float)));
    if (c->kern_tc && nd.L >= 2) {
        c->wq_term_stride = (long long)(nd.L - 1) * nd.H * nd.H;
        TRY_C(cudaMalloc(&c->d_wq, (size_t)c->nterms * c->wq_term_stride * sizeof(__nv_bfloat16)));
    }
    if (c->tc_defer_g > 0) TRY_C(cudaMalloc(&c->d_ring, (size_t)c->grid_cap_tc * c->ring_per_cta));
    if (c->kern_tc2) {
        const size_t opb = (size_t)nd.H * 256, np = (size_t)c->tc2_nprod;
        TRY_C(cudaMalloc(&c->d_wf, (size_t)(nd.L - 1) * nd.H * nd.H * sizeof(__nv_bfloat16)));
        TRY_C(cudaMalloc(&c->d_wb, (size_t)(nd.L - 1) * nd.H * nd.H * sizeof(__nv_bfloat16)));
        TRY_C(cudaMalloc(&c->d_st_a, np * nd.L * opb));
        TRY_C(cudaMalloc(&c->d_st_s, np * nd.L * tc2::NGRP * nd.H * 8));
        TRY_C(cudaMalloc(&c->d_zring, np * c->tc2_zdepth * opb));
        TRY_C(cudaMalloc(&c->d_tc2_flags, 3 * np * kFlagStride * sizeof(unsigned int)));
        if (const char* ts = getenv("PINN_TC2_TIMEOUT_SHIFT")) {   // under a profiler every hand-off is slower: relax the wait bounds
            const long long bound = 1LL << (30 + atoi(ts));
            TRY_C(cudaMemcpyToSymbol(tc2::g_wait_bound, &bound, sizeof(bound)));
        }
        if (getenv("PINN_TC2_DBG")) { TRY_C(cudaMalloc(&c->d_tc2_dbg, 512 * sizeof(long long))); TRY_C(cudaMemsetAsync(c->d_tc2_dbg, 0, 512 * sizeof(long long), c->stream)); }
    }
    TRY_C(cudaMalloc(&c->d_loss, 8 * sizeof(float)));          // [0..3] loss record, [4] overflow flag word: one D2H copy per read-back
    TRY_C(cudaMalloc(&c->d_step, sizeof(unsigned long long)));
    TRY_C(cudaMalloc(&c->d_tail, sizeof(unsigned int)));
    TRY_C(cudaMalloc(&c->d_beta_pow, 2 * sizeof(double)));
    TRY_C(cudaMemsetAsync(c->d_tail, 0, sizeof(unsigned int), c->stream));
    c->d_flag = reinterpret_cast<int*>(c->d_loss + 4);
    TRY_C(cudaHostAlloc(&c->h_pinned, 32 * sizeof(float), cudaHostAllocMapped));   // [0,8) step record, [8,13) copy-path scratch, [16,20) trap site
    memset(c->h_pinned, 0, 32 * sizeof(float));
    TRY_C(cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->d_hrec), c->h_pinned, 0));
    TRY_C(cudaMemsetAsync(c->d_m, 0, P * sizeof(float), c->stream));
    TRY_C(cudaMemsetAsync(c->d_v, 0, P * sizeof(float), c->stream));
    TRY_C(cudaMemsetAsync(c->d_gbuf, 0, (size_t)nd.PP * sizeof(float), c->stream));
    TRY_C(cudaMemsetAsync(c->d_partial, 0, nrows * nd.PP * sizeof(float), c->stream));
    TRY_C(cudaMemsetAsync(c->d_loss, 0, 8 * sizeof(float), c->stream));
    TRY_C(cudaMemsetAsync(c->d_step, 0, sizeof(unsigned long long), c->stream));
#undef TRY_C
    std::vector<float> w;
    host_init_params(nd, cfg->seed_weights, w);
    *out = c;
    rc = pinn_set_params(c, w.data(), w.size(), 1);
    if (rc) { *out = nullptr; return bail(rc); }
    rc = pinn_sample_points(c);
    if (rc) { *out = nullptr; return bail(rc); }
    return PINN_OK;
}

void pinn_destroy(pinn_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    if (c->stream2) cudaStreamSynchronize(c->stream2);
    if (c->stream) cudaStreamSynchronize(c->stream);
    drain_profile(c);
    if (c->d_tc2_dbg) {
        long long h[512] = {0};
        if (cudaMemcpy(h, c->d_tc2_dbg, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess && c->tc2_T == tc3::T) {
            fprintf(stderr, "[tc3 dbg] cycles of producer 0, last launch | epilogue thread 0: waits %lld  layer-1 %lld  fwd epi %lld  out + rev L %lld  rev epi %lld  total %lld\n",
                    h[0], h[1], h[2], h[3], h[4], h[7]);
            fprintf(stderr, "[tc3 dbg] forward epilogues of thread 0: TMEM load %lld  math + stores %lld  fences + arrive %lld\n", h[8], h[9], h[10]);
            // timeline of one tile (PINN_TC2_DBG=2 prints it): cycles since the first stamp of the tile
            if (atoi(getenv("PINN_TC2_DBG") ? getenv("PINN_TC2_DBG") : "0") >= 2) {
                const int L = c->nd.L;
                const long long t0 = h[160];
                for (int e = 0; e < 2 * L - 1; e++)
                    for (int x = 0; x < 2; x++) {
                        const long long* r = h + 16 + (e * 2 + x) * 4;
                        fprintf(stderr, "  event %2d issuer %c: start %6lld  ready %6lld  (weights %5lld)  issued %6lld\n", e, 'A' + x, r[0] - t0, r[1] - t0, r[2] - r[1], r[3] - t0);
                    }
                for (int ph = 0; ph < 2 * L; ph++)
                    for (int x = 0; x < 2; x++) {
                        const long long* r = h + 160 + (ph * 2 + x) * 3;
                        fprintf(stderr, "  phase %2d %c epilogue: enter %6lld  work %6lld .. %6lld\n", ph, 'A' + x, r[0] - t0, r[1] - t0, r[2] - t0);
                    }
            }
        } else
            fprintf(stderr, "[tc2 dbg] cycles of producer 0, last launch: fwd wait %lld  fwd epi %lld  out %lld  rev wait %lld  rev epi %lld  rev issue %lld  layer1 %lld  total %lld\n",
                    h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
        cudaFree(c->d_tc2_dbg);
    }
    if (c->comm && g_nccl.ok) g_nccl.CommDestroy(c->comm);
    for (int k = 0; k < 3; k++) cudaFree(c->d_x[k]);
    cudaFree(c->d_params); cudaFree(c->d_paramsT); cudaFree(c->d_m); cudaFree(c->d_v);
    cudaFree(c->d_gbuf); cudaFree(c->d_partial); cudaFree(c->d_stash); cudaFree(c->d_loss);
    cudaFree(c->d_step); cudaFree(c->d_wq); cudaFree(c->d_ring); cudaFree(c->d_wfrag);
    cudaFree(c->d_wf); cudaFree(c->d_wb); cudaFree(c->d_st_a); cudaFree(c->d_st_s); cudaFree(c->d_zring); cudaFree(c->d_tc2_flags);
    for (int r = 0; r < kMaxP2PRanks; r++)
        if (c->peer_ptr[r] && c->peer_ptr[r] != c->d_xchg) cudaIpcCloseMemHandle(c->peer_ptr[r]);
    cudaFree(c->d_xchg); cudaFree(c->d_done); cudaFree(c->d_tail); cudaFree(c->d_beta_pow);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    drop_graph(c);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int pinn_set_params(pinn_ctx* c, const float* flat, size_t n, int reset_optimizer) {
    if (!c || !flat) return fail(c, PINN_ERR_INVALID_ARGUMENT, "pinn_set_params: null argument");
    if (n != (size_t)c->nd.P) return fail(c, PINN_ERR_INVALID_ARGUMENT, "pinn_set_params: size mismatch - " + std::to_string(n) + " != " + std::to_string(c->nd.P));
    CUDA_TRY(c, "pinn_set_params", cudaSetDevice(c->cfg.device));
    CUDA_TRY(c, "pinn_set_params", cudaMemcpyAsync(c->d_params, flat, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    transpose_params_kernel<<<(c->nd.P + 255) / 256, 256, 0, c->stream>>>(c->nd, c->d_params, c->d_paramsT);
    c->launches++;
    refresh_wq(c);
    if (reset_optimizer) {
        CUDA_TRY(c, "pinn_set_params", cudaMemsetAsync(c->d_m, 0, n * sizeof(float), c->stream));
        CUDA_TRY(c, "pinn_set_params", cudaMemsetAsync(c->d_v, 0, n * sizeof(float), c->stream));
        CUDA_TRY(c, "pinn_set_params", cudaMemsetAsync(c->d_step, 0, sizeof(unsigned long long), c->stream));
        c->host_step = 0;
    }
    CUDA_TRY(c, "pinn_set_params", cudaStreamSynchronize(c->stream));
    if (reset_optimizer && c->h_pinned) { c->h_pinned[3] = 0.0f; c->h_pinned[6] = 0.0f; }         // sequence words of read_step
    if (reset_optimizer) { const int rb = upload_beta_pow(c, 0); if (rb) return rb; }
    return PINN_OK;
}

static int copy_out(pinn_ctx* c, const char* fn, const float* src, float* dst, size_t n) {
    if (!c || !dst) return fail(c, PINN_ERR_INVALID_ARGUMENT, std::string(fn) + ": null argument");
    if (n != (size_t)c->nd.P) return fail(c, PINN_ERR_INVALID_ARGUMENT, std::string(fn) + ": size mismatch");
    CUDA_TRY(c, fn, cudaSetDevice(c->cfg.device));
    CUDA_TRY(c, fn, cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, fn, cudaStreamSynchronize(c->stream));
    return PINN_OK;
}
int pinn_get_params(pinn_ctx* c, float* flat, size_t n) { return copy_out(c, "pinn_get_params", c ? c->d_params : nullptr, flat, n); }
int pinn_get_grads(pinn_ctx* c, float* flat, size_t n) { return copy_out(c, "pinn_get_grads", c ? c->d_gbuf : nullptr, flat, n); }

int pinn_get_optimizer(pinn_ctx* c, float* m, float* v, size_t n, uint64_t* step) {
    int rc = copy_out(c, "pinn_get_optimizer", c ? c->d_m : nullptr, m, n);
    if (rc) return rc;
    rc = copy_out(c, "pinn_get_optimizer", c->d_v, v, n);
    if (rc) return rc;
    if (step) *step = c->host_step;
    return PINN_OK;
}

int pinn_set_optimizer(pinn_ctx* c, const float* m, const float* v, size_t n, uint64_t step) {
    if (!c || !m || !v) return fail(c, PINN_ERR_INVALID_ARGUMENT, "pinn_set_optimizer: null argument");
    if (n != (size_t)c->nd.P) return fail(c, PINN_ERR_INVALID_ARGUMENT, "pinn_set_optimizer: size mismatch");
    CUDA_TRY(c, "pinn_set_optimizer", cudaSetDevice(c->cfg.device));
    unsigned long long s = step;
    CUDA_TRY(c, "pinn_set_optimizer", cudaMemcpyAsync(c->d_m, m, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, "pinn_set_optimizer", cudaMemcpyAsync(c->d_v, v, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, "pinn_set_optimizer", cudaMemcpyAsync(c->d_step, &s, sizeof(s), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, "pinn_set_optimizer", cudaStreamSynchronize(c->stream));
    c->host_step = step;
    c->h_pinned[3] = 0.0f; c->h_pinned[6] = 0.0f;                                  // sequence words of read_step
    { const int rb = upload_beta_pow(c, step); if (rb) return rb; }
    return PINN_OK;
}

int pinn_sample_points(pinn_ctx* c) {
    if (!c) return PINN_ERR_INVALID_ARGUMENT;
    CUDA_TRY(c, "pinn_sample_points", cudaSetDevice(c->cfg.device));
    for (int k = 0; k < 3; k++) {
        c->n_local[k] = c->cap[k];
        if (!c->cap[k]) continue;
        const uint64_t n = c->cap[k];
        sample_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->nd, k, c->seeds[k], c->begin[k], n, c->d_x[k]);
        c->launches++;
    }
    CUDA_TRY(c, "pinn_sample_points", cudaGetLastError());
    return PINN_OK;
}

int pinn_resample_points(pinn_ctx* c, uint64_t si, uint64_t sb, uint64_t s0) {
    if (!c) return PINN_ERR_INVALID_ARGUMENT;
    c->seeds[0] = si; c->seeds[1] = sb; c->seeds[2] = s0;
    return pinn_sample_points(c);
}

int pinn_set_points(pinn_ctx* c, int set, const float* x, size_t n) {
    if (!c || set < 0 || set > 2 || (!x && n)) return fail(c, PINN_ERR_INVALID_ARGUMENT, "pinn_set_points: invalid argument");
    if (n > c->cap[set]) return fail(c, PINN_ERR_INVALID_ARGUMENT, "pinn_set_points: more points than the shard fixed at create time");
    CUDA_TRY(c, "pinn_set_points", cudaSetDevice(c->cfg.device));
    if (n) CUDA_TRY(c, "pinn_set_points", cudaMemcpyAsync(c->d_x[set], x, n * c->nd.d_in * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    c->h2d_point_bytes += n * c->nd.d_in * sizeof(float);
    c->n_local[set] = n;
    return PINN_OK;
}

// Points that already live on this device — the device half of a reference tensor<float> (interface<T>::d_var().data,
// include/device/interface.cuh:248-249) or any other device buffer: device-to-device, no host round trip.
int pinn_set_points_device(pinn_ctx* c, int set, const float* d_x, size_t n) {
    if (!c || set < 0 || set > 2 || (!d_x && n)) return fail(c, PINN_ERR_INVALID_ARGUMENT, "pinn_set_points_device: invalid argument");
    if (n > c->cap[set]) return fail(c, PINN_ERR_INVALID_ARGUMENT, "pinn_set_points_device: more points than the shard fixed at create time");
    CUDA_TRY(c, "pinn_set_points_device", cudaSetDevice(c->cfg.device));
    if (n) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, d_x) != cudaSuccess || at.type != cudaMemoryTypeDevice || at.device != c->cfg.device) {
            cudaGetLastError();
            return fail(c, PINN_ERR_INVALID_ARGUMENT, "pinn_set_points_device: not a device pointer of this handle's GPU (host memory goes through pinn_set_points)");
        }
        CUDA_TRY(c, "pinn_set_points_device", cudaMemcpyAsync(c->d_x[set], d_x, n * c->nd.d_in * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    }
    c->n_local[set] = n;
    return PINN_OK;
}

// The same from the reference's kernel-argument view device::tensor::variables<float> (include/device/variables.cuh:6-52):
// a 48-byte POD passed BY VALUE to every reference kernel — {T* data @0, size_t* shape @8, size_t* stride @16, size_t dim @24,
// size_t n @32, bool transposed @40}.  Accepts a [n_points, d_in] or flat view whose n is a multiple of d_in.
int pinn_set_points_view(pinn_ctx* c, int set, const void* variables48) {
    if (!c || !variables48) return fail(c, PINN_ERR_INVALID_ARGUMENT, "pinn_set_points_view: null argument");
    struct View { const float* data; const size_t* shape; const size_t* stride; size_t dim, n; bool transposed; };
    static_assert(sizeof(View) == 48, "device::tensor::variables<float> is 48 bytes");
    View v; memcpy(&v, variables48, sizeof(v));
    if (v.transposed) return fail(c, PINN_ERR_INVALID_ARGUMENT, "pinn_set_points_view: transposed views are not contiguous point lists");
    if (v.n % (size_t)c->nd.d_in) return fail(c, PINN_ERR_INVALID_ARGUMENT, "pinn_set_points_view: size mismatch - " + std::to_string(v.n) + " is not a multiple of d_in");
    return pinn_set_points_device(c, set, v.data, v.n / (size_t)c->nd.d_in);
}

int pinn_train_step_device(pinn_ctx* c, const float* d_x_interior, size_t n, pinn_step_out* out) {
    if (!c) return PINN_ERR_INVALID_ARGUMENT;
    if (int rcn = need_comm(c, "pinn_train_step_device")) return rcn;
    const int rc = pinn_set_points_device(c, PINN_SET_INTERIOR, d_x_interior, n);
    return rc ? rc : pinn_train_step(c, out);
}

unsigned long long pinn_h2d_point_bytes(const pinn_ctx* c) { return c ? c->h2d_point_bytes : 0ull; }

int pinn_get_points(pinn_ctx* c, int set, float* x, size_t n) {
    if (!c || set < 0 || set > 2 || (!x && n)) return fail(c, PINN_ERR_INVALID_ARGUMENT, "pinn_get_points: invalid argument");
    if (n > c->n_local[set]) return fail(c, PINN_ERR_INVALID_ARGUMENT, "pinn_get_points: more points than resident");
    CUDA_TRY(c, "pinn_get_points", cudaSetDevice(c->cfg.device));
    if (n) CUDA_TRY(c, "pinn_get_points", cudaMemcpyAsync(x, c->d_x[set], n * c->nd.d_in * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, "pinn_get_points", cudaStreamSynchronize(c->stream));
    return PINN_OK;
}

size_t pinn_local_count(const pinn_ctx* c, int set) { return (c && set >= 0 && set <= 2) ? (size_t)c->n_local[set] : 0; }

int pinn_step_local(pinn_ctx* c) {
    if (!c) return PINN_ERR_INVALID_ARGUMENT;
    CUDA_TRY(c, "pinn_step_local", cudaSetDevice(c->cfg.device));
    c->step_launches = 0;
    const int rc = run_local(c, 1);
    c->launches += c->step_launches;
    return rc;
}

int pinn_step_apply(pinn_ctx* c, pinn_step_out* out) {
    if (!c) return PINN_ERR_INVALID_ARGUMENT;
    CUDA_TRY(c, "pinn_step_apply", cudaSetDevice(c->cfg.device));
    c->step_launches = 0;
    int rc = run_apply(c, 1);
    c->launches += c->step_launches;
    if (rc) return rc;
    c->host_step++;
    return read_out(c, out);
}

int pinn_train_step(pinn_ctx* c, pinn_step_out* out) {
    if (!c) return PINN_ERR_INVALID_ARGUMENT;
    if (int rcn = need_comm(c, "pinn_train_step")) return rcn;
    CUDA_TRY(c, "pinn_train_step", cudaSetDevice(c->cfg.device));
    int rc = step_once(c);
    if (!rc) rc = read_step(c, out);
    return rc;
}

int pinn_train_steps(pinn_ctx* c, int k, pinn_step_out* out) {
    if (!c || k < 1) return fail(c, PINN_ERR_INVALID_ARGUMENT, "pinn_train_steps: k must be >= 1");
    if (int rcn = need_comm(c, "pinn_train_steps")) return rcn;
    CUDA_TRY(c, "pinn_train_steps", cudaSetDevice(c->cfg.device));
    for (int s = 0; s < k; s++) {
        const int rc = step_once(c);
        if (rc) return rc;
    }
    return read_step(c, out);
}

int pinn_train_step_host(pinn_ctx* c, const float* x_interior, size_t n, pinn_step_out* out) {
    if (!c) return PINN_ERR_INVALID_ARGUMENT;
    if (int rcn = need_comm(c, "pinn_train_step_host")) return rcn;
    // Width-20 mma kernel + pinned caller memory + a caller that waits for the result: no copy operation at all — the kernel
    // reads the points straight from the mapped host buffer (one 128-byte PCIe read per 16-point tile, hidden behind the
    // tile's own work) and leaves the resident copy in HBM behind for pinn_get_points / later steps.
    if (c->kern_mma && out && x_interior && n > 0 && n <= c->cap[PINN_SET_INTERIOR]) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, x_interior) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
            CUDA_TRY(c, "pinn_train_step_host", cudaSetDevice(c->cfg.device));
            c->n_local[PINN_SET_INTERIOR] = n;
            c->x_override = static_cast<const float*>(at.devicePointer);
            c->step_launches = 0;
            int rc = issue_step(c);
            c->launches += c->step_launches;
            c->x_override = nullptr;
            if (!rc) { c->host_step++; rc = read_step(c, out); }
            return rc;
        }
        cudaGetLastError();                               // pageable memory: not an error, take the copy path
    }
    int rc = pinn_set_points(c, PINN_SET_INTERIOR, x_interior, n);
    if (rc) return rc;
    return pinn_train_step(c, out);
}

int pinn_loss_grad(pinn_ctx* c, pinn_step_out* out) {
    if (!c) return PINN_ERR_INVALID_ARGUMENT;
    CUDA_TRY(c, "pinn_loss_grad", cudaSetDevice(c->cfg.device));
    c->step_launches = 0;
    int rc = run_local(c, 0);
    if (!rc) rc = run_allreduce(c);
    if (!rc) rc = run_apply(c, 0);
    c->launches += c->step_launches;
    if (!rc) rc = read_out(c, out);
    return rc;
}

static int eval_common(pinn_ctx* c, const char* fn, bool full, const float* x, float* y, float* r, size_t n) {
    if (!c || (!x && n)) return fail(c, PINN_ERR_INVALID_ARGUMENT, std::string(fn) + ": null argument");
    if (n == 0) return PINN_OK;
    CUDA_TRY(c, fn, cudaSetDevice(c->cfg.device));
    const NetDims& nd = c->nd;
    const int nc = full ? nd.C : 1;
    float *dx = nullptr, *dy = nullptr, *dr = nullptr;
    CUDA_TRY(c, fn, cudaMallocAsync(&dx, n * nd.d_in * sizeof(float), c->stream));
    CUDA_TRY(c, fn, cudaMallocAsync(&dy, n * nd.d_out * nc * sizeof(float), c->stream));
    if (full) CUDA_TRY(c, fn, cudaMallocAsync(&dr, n * nd.n_res * sizeof(float), c->stream));
    CUDA_TRY(c, fn, cudaMemcpyAsync(dx, x, n * nd.d_in * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    PointSet s{};
    s.x = dx; s.n = n; s.scale = 0.0f; s.set = full ? PINN_SET_INTERIOR : PINN_SET_BOUNDARY; s.ntiles = tiles_of(n, full ? c->TP : c->val_TP);
    c->step_launches = 0;
    int rc = launch_fused(c, c->stream, full, &s, 1, 1, dy, dr, 0, nullptr);
    c->launches += c->step_launches;
    if (rc) return rc;
    if (y) CUDA_TRY(c, fn, cudaMemcpyAsync(y, dy, n * nd.d_out * nc * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (r && full) CUDA_TRY(c, fn, cudaMemcpyAsync(r, dr, n * nd.n_res * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    cudaFreeAsync(dx, c->stream); cudaFreeAsync(dy, c->stream);
    if (dr) cudaFreeAsync(dr, c->stream);
    CUDA_TRY(c, fn, cudaStreamSynchronize(c->stream));
    return PINN_OK;
}

int pinn_forward(pinn_ctx* c, const float* x, float* y, size_t n) { return eval_common(c, "pinn_forward", false, x, y, nullptr, n); }
int pinn_residual(pinn_ctx* c, const float* x, float* y, float* r, size_t n) { return eval_common(c, "pinn_residual", true, x, y, r, n); }

// ---- checkpoint + evaluation ----------------------------------------------------
namespace {
struct CkptHeader {                     // 64 bytes, little-endian (pinn_abi.h: pinn_save)
    char     magic[8];
    uint32_t version;
    int32_t  pde, d_in, d_out, width, depth;
    uint32_t reserved0[2];
    uint64_t P, step;
    uint8_t  reserved1[8];
};
static_assert(sizeof(CkptHeader) == 64, "checkpoint header is 64 bytes");
const char kCkptMagic[8] = {'P', 'I', 'N', 'N', 'C', 'K', 'P', 'T'};
struct FileCloser { void operator()(FILE* f) const { if (f) fclose(f); } };
}  // namespace

int pinn_save(pinn_ctx* c, const char* path) {
    if (!c || !path) return fail(c, PINN_ERR_INVALID_ARGUMENT, "pinn_save: null argument");
    const size_t P = (size_t)c->nd.P;
    std::vector<float> w(P), m(P), v(P);
    uint64_t step = 0;
    int rc = pinn_get_params(c, w.data(), P);
    if (rc) return rc;
    rc = pinn_get_optimizer(c, m.data(), v.data(), P, &step);
    if (rc) return rc;
    CkptHeader h{};
    memcpy(h.magic, kCkptMagic, 8);
    h.version = 1; h.pde = c->cfg.pde; h.d_in = c->cfg.d_in; h.d_out = c->cfg.d_out; h.width = c->cfg.width; h.depth = c->cfg.depth;
    h.P = P; h.step = step;
    std::unique_ptr<FILE, FileCloser> f(fopen(path, "wb"));
    if (!f) return fail(c, PINN_ERR_INVALID_ARGUMENT, std::string("pinn_save: cannot open ") + path);
    const bool ok = fwrite(&h, sizeof(h), 1, f.get()) == 1 && fwrite(w.data(), sizeof(float), P, f.get()) == P &&
                    fwrite(m.data(), sizeof(float), P, f.get()) == P && fwrite(v.data(), sizeof(float), P, f.get()) == P &&
                    fflush(f.get()) == 0;
    if (!ok) return fail(c, PINN_ERR_INVALID_ARGUMENT, std::string("pinn_save: short write to ") + path);
    return PINN_OK;
}

int pinn_load(pinn_ctx* c, const char* path) {
    if (!c || !path) return fail(c, PINN_ERR_INVALID_ARGUMENT, "pinn_load: null argument");
    std::unique_ptr<FILE, FileCloser> f(fopen(path, "rb"));
    if (!f) return fail(c, PINN_ERR_INVALID_ARGUMENT, std::string("pinn_load: cannot open ") + path);
    CkptHeader h{};
    if (fread(&h, sizeof(h), 1, f.get()) != 1 || memcmp(h.magic, kCkptMagic, 8) != 0)
        return fail(c, PINN_ERR_INVALID_ARGUMENT, std::string("pinn_load: not a checkpoint - ") + path);
    if (h.version != 1) return fail(c, PINN_ERR_INVALID_ARGUMENT, "pinn_load: unknown checkpoint version " + std::to_string(h.version));
    if (h.pde != c->cfg.pde || h.d_in != c->cfg.d_in || h.d_out != c->cfg.d_out || h.width != c->cfg.width ||
        h.depth != c->cfg.depth || h.P != (uint64_t)c->nd.P)
        return fail(c, PINN_ERR_INVALID_ARGUMENT, "pinn_load: network mismatch - file has pde " + std::to_string(h.pde) + " " +
                    std::to_string(h.d_in) + "->[" + std::to_string(h.width) + "x" + std::to_string(h.depth) + "]->" + std::to_string(h.d_out));
    const size_t P = (size_t)h.P;
    std::vector<float> w(P), m(P), v(P);
    if (fread(w.data(), sizeof(float), P, f.get()) != P || fread(m.data(), sizeof(float), P, f.get()) != P ||
        fread(v.data(), sizeof(float), P, f.get()) != P)
        return fail(c, PINN_ERR_INVALID_ARGUMENT, std::string("pinn_load: truncated checkpoint - ") + path);
    // both uploads are validated before either mutates the handle ("left untouched" on failure): sizes match by the header
    // check above; what can still fail is a CUDA call, after which the handle is unusable anyway
    std::vector<float> w_old(P);
    int rc = pinn_get_params(c, w_old.data(), P);
    if (rc) return rc;
    rc = pinn_set_params(c, w.data(), P, 0);
    if (rc) return rc;
    rc = pinn_set_optimizer(c, m.data(), v.data(), P, h.step);
    if (rc) pinn_set_params(c, w_old.data(), P, 0);      // roll the parameters back: the handle is as it was
    return rc;
}

int pinn_exact_solution(const pinn_config* cfg, const float* x, float* u, size_t n) {
    if (!cfg || (n && (!x || !u))) return PINN_ERR_INVALID_ARGUMENT;
    int n_space, n2, n_res, d_out;
    if (!pde_shape(cfg->pde, &n_space, &n2, &n_res, &d_out)) return PINN_ERR_INVALID_ARGUMENT;
    const double pi = 3.14159265358979323846;
    const int d = cfg->d_in;
    switch (cfg->pde) {
    case PINN_PDE_HELMHOLTZ:
        for (size_t i = 0; i < n; i++) u[i] = (float)(sin(pi * x[i * d]) * sin(4.0 * pi * x[i * d + 1]));
        return PINN_OK;
    case PINN_PDE_HEAT:
        for (size_t i = 0; i < n; i++)
            u[i] = (float)(sin(pi * x[i * d]) * sin(pi * x[i * d + 1]) * exp(-2.0 * (double)kHeatAlpha * pi * pi * x[i * d + 2]));
        return PINN_OK;
    case PINN_PDE_NAVIER_STOKES:
        for (size_t i = 0; i < n; i++) {
            const double X = x[i * d], Y = x[i * d + 1], T = x[i * d + 2];
            const double e2 = exp(-2.0 * (double)kNsNu * T), e4 = exp(-4.0 * (double)kNsNu * T);
            u[i * 3 + 0] = (float)(-cos(X) * sin(Y) * e2);
            u[i * 3 + 1] = (float)(sin(X) * cos(Y) * e2);
            u[i * 3 + 2] = (float)(-0.25 * (cos(2.0 * X) + cos(2.0 * Y)) * e4);
        }
        return PINN_OK;
    default:
        return PINN_ERR_UNSUPPORTED;     // Burgers, Allen-Cahn: no closed form
    }
}

int pinn_evaluate(pinn_ctx* c, int per_axis, pinn_eval_out* out) {
    if (!c || !out) return fail(c, PINN_ERR_INVALID_ARGUMENT, "pinn_evaluate: null argument");
    if (per_axis < 2 || per_axis > 4096) return fail(c, PINN_ERR_INVALID_ARGUMENT, "pinn_evaluate: per_axis must be in [2, 4096]");
    const NetDims& nd = c->nd;
    const int d = nd.d_in, C = nd.C, O = nd.d_out, R = nd.n_res;
    uint64_t total = 1;
    for (int k = 0; k < d; k++) total *= (uint64_t)per_axis;
    if (total > ((uint64_t)1 << 32)) return fail(c, PINN_ERR_INVALID_ARGUMENT, "pinn_evaluate: grid larger than 2^32 nodes");
    memset(out, 0, sizeof(*out));
    out->n = total;
    const size_t chunk = 1 << 18;
    std::vector<float> x(chunk * d), y(chunk * O * C), r(chunk * R), u(chunk * O);
    double num[3] = {0, 0, 0}, den[3] = {0, 0, 0}, mx[3] = {0, 0, 0}, rs = 0.0;
    bool exact = true;
    for (uint64_t i0 = 0; i0 < total; i0 += chunk) {
        const size_t m = (size_t)std::min<uint64_t>(chunk, total - i0);
        for (size_t i = 0; i < m; i++) {
            uint64_t q = i0 + i;
            for (int k = d; k-- > 0;) {
                const uint64_t idx = q % (uint64_t)per_axis; q /= (uint64_t)per_axis;
                x[i * d + k] = nd.lo[k] + nd.span[k] * (float)idx / (float)(per_axis - 1);
            }
        }
        int rc = eval_common(c, "pinn_evaluate", true, x.data(), y.data(), r.data(), m);
        if (rc) return rc;
        for (size_t i = 0; i < m * (size_t)R; i++) rs += (double)r[i] * (double)r[i];
        if (exact && pinn_exact_solution(&c->cfg, x.data(), u.data(), m) != PINN_OK) exact = false;
        if (exact)
            for (size_t i = 0; i < m; i++)
                for (int o = 0; o < O; o++) {
                    const double t = u[i * O + o], e = (double)y[(i * O + o) * C] - t;
                    num[o] += e * e; den[o] += t * t; mx[o] = std::max(mx[o], fabs(e));
                }
    }
    out->mse_residual = rs / (double)total;
    for (int o = 0; o < 3; o++) {
        out->rel_l2[o] = (exact && o < O) ? sqrt(num[o] / (den[o] > 0.0 ? den[o] : 1.0)) : NAN;
        out->max_abs[o] = (exact && o < O) ? mx[o] : NAN;
    }
    return PINN_OK;
}

int pinn_synchronize(pinn_ctx* c) {
    if (!c) return PINN_ERR_INVALID_ARGUMENT;
    CUDA_TRY(c, "pinn_synchronize", cudaSetDevice(c->cfg.device));
    CUDA_TRY(c, "pinn_synchronize", cudaStreamSynchronize(c->stream));
    return PINN_OK;
}

int pinn_comm_unique_id(void* id128) {
    std::string err;
    if (!id128) return PINN_ERR_INVALID_ARGUMENT;
    if (!nccl_load(err)) return fail(nullptr, PINN_ERR_NCCL, "pinn_comm_unique_id: " + err);
    const int rc = g_nccl.GetUniqueId(id128);
    return rc == 0 ? PINN_OK : fail(nullptr, PINN_ERR_NCCL, "pinn_comm_unique_id: ncclGetUniqueId failed");
}

int pinn_comm_init(pinn_ctx* c, const void* id128) {
    if (!c || !id128) return fail(c, PINN_ERR_INVALID_ARGUMENT, "pinn_comm_init: null argument");
    if (c->cfg.nranks <= 1) return PINN_OK;
    std::string err;
    if (!nccl_load(err)) return fail(c, PINN_ERR_NCCL, "pinn_comm_init: " + err);
    CUDA_TRY(c, "pinn_comm_init", cudaSetDevice(c->cfg.device));
    NcclId id; memcpy(&id, id128, sizeof(id));
    const int rc = g_nccl.CommInitRank(&c->comm, c->cfg.nranks, id, c->cfg.rank);
    if (rc != 0) return fail(c, PINN_ERR_NCCL, std::string("pinn_comm_init: ncclCommInitRank -> ") +
                                                   (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"));
    return PINN_OK;
}

int pinn_comm_p2p_export(pinn_ctx* c, void* handle64) {
    if (!c || !handle64) return fail(c, PINN_ERR_INVALID_ARGUMENT, "pinn_comm_p2p_export: null argument");
    if (c->cfg.nranks > kMaxP2PRanks) return fail(c, PINN_ERR_UNSUPPORTED, "pinn_comm_p2p_export: more than 16 ranks");
    CUDA_TRY(c, "pinn_comm_p2p_export", cudaSetDevice(c->cfg.device));
    if (!c->d_xchg) {
        const size_t bytes = (kXchgHeader + 2 * (size_t)c->cfg.nranks * (size_t)c->nd.PP) * sizeof(float);   // [source rank][parity][PP]
        CUDA_TRY(c, "pinn_comm_p2p_export", cudaMalloc(&c->d_xchg, bytes));
        CUDA_TRY(c, "pinn_comm_p2p_export", cudaMemset(c->d_xchg, 0, bytes));
        CUDA_TRY(c, "pinn_comm_p2p_export", cudaMalloc(&c->d_done, sizeof(unsigned int)));
        CUDA_TRY(c, "pinn_comm_p2p_export", cudaMemset(c->d_done, 0, sizeof(unsigned int)));
        CUDA_TRY(c, "pinn_comm_p2p_export", cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t h;
    CUDA_TRY(c, "pinn_comm_p2p_export", cudaIpcGetMemHandle(&h, c->d_xchg));
    static_assert(sizeof(h) == PINN_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, sizeof(h));
    return PINN_OK;
}

int pinn_comm_p2p_open(pinn_ctx* c, const void* handles, int count) {
    if (!c || !handles) return fail(c, PINN_ERR_INVALID_ARGUMENT, "pinn_comm_p2p_open: null argument");
    if (count != c->cfg.nranks || !c->d_xchg) return fail(c, PINN_ERR_INVALID_ARGUMENT, "pinn_comm_p2p_open: need one handle per rank, after pinn_comm_p2p_export");
    CUDA_TRY(c, "pinn_comm_p2p_open", cudaSetDevice(c->cfg.device));
    {   // every block of the exchange kernel waits for the last block of its own rank: all of them must be resident at once
        int occ = 0;
        CUDA_TRY(c, "pinn_comm_p2p_open", cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, reduce_exchange_adam_kernel, 1024, 0));
        const int blocks = (c->nd.P + 3 + 31) / 32;
        if (blocks > occ * c->num_sms)
            return fail(c, PINN_ERR_UNSUPPORTED, "pinn_comm_p2p_open: the fused exchange kernel needs " + std::to_string(blocks) +
                        " co-resident blocks, this GPU holds " + std::to_string(occ * c->num_sms) + " (use the NCCL collective for this network)");
    }
    for (int r = 0; r < count; r++) {
        if (r == c->cfg.rank) { c->peer_ptr[r] = c->d_xchg; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(handles) + (size_t)r * PINN_IPC_HANDLE_BYTES, sizeof(h));
        void* ptr = nullptr;
        CUDA_TRY(c, "pinn_comm_p2p_open", cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer_ptr[r] = static_cast<float*>(ptr);
    }
    c->p2p_ready = true;
    return PINN_OK;
}

void* pinn_grad_buffer(pinn_ctx* c, size_t* n_floats) {
    if (!c) return nullptr;
    if (n_floats) *n_floats = (size_t)c->nd.P + 3;
    return c->d_gbuf;
}
void* pinn_stream(pinn_ctx* c) { return c ? (void*)c->stream : nullptr; }

int pinn_graph_state(const pinn_ctx* c, const char** note) {
    if (note) *note = c ? c->graph_note.c_str() : "";
    if (!c) return -1;
    return c->graph_exec != nullptr ? 1 : 0;
}

int pinn_profile_enable(pinn_ctx* c, int enable) {
    if (!c) return PINN_ERR_INVALID_ARGUMENT;
    c->prof = enable != 0;
    return PINN_OK;
}

int pinn_profile_get(pinn_ctx* c, pinn_profile* out, int reset) {
    if (!c || !out) return PINN_ERR_INVALID_ARGUMENT;
    cudaSetDevice(c->cfg.device);
    drain_profile(c);
    out->launches = c->launches; out->step_kernel_ms = c->kernel_ms; out->step_kernel_count = c->kernel_count;
    out->grid = c->last_grid;
    // launch shape of the kernel the interior training launch runs (the mma.sync kernel used to report the FP32 kernel's)
    out->block = c->kern_mma ? kMmaWarps * 32 : (c->kern_tc2 ? c->tc2_threads : (c->kern_tc ? c->tc_threads : c->threads));
    out->smem_bytes = c->kern_mma ? c->mma_smem : (c->kern_tc2 ? c->tc2_smem : (c->kern_tc ? c->tc_smem : c->smem_full));
    out->kernel_id = c->kernel_id;
    if (reset) { c->launches = 0; c->kernel_ms = 0.0; c->kernel_count = 0; }
    return PINN_OK;
}

}  // extern "C"
