/*
 * pinn_abi.h — C ABI of the B200-native PINN train-step hot path.
 *
 * This is the drop-in boundary for the path BASELINE.json:north_star names: one training
 * step over a batch of collocation points (dense tanh-MLP forward, forward-mode Taylor
 * tangents u, u_x, u_xx, u_t, residual + boundary + initial MSE loss, reverse-mode weight
 * gradients, Adam).  The reference (Deadman621/CUDA-PINN) has NO plugin / FFI interface and
 * its network layer is an empty placeholder:
 *
 *   reference include/network.cuh      (0 bytes)  — the `network` class this ABI sits under
 *   reference src/network.cu           (0 bytes)  — "layers, forward and backward passes, and
 *                                                    loss calculations" (README.md:53)
 *   reference src/utils.cu             (0 bytes)  — "data handling and normalization"
 *                                                    (README.md:54) → collocation sampling here
 *
 * so every entry point below cites the reference building block it replaces (the op a
 * reference-style `network` would have had to compose) rather than an existing FFI symbol.
 * The C++ `network<T>` (include/network.cuh in this repo, written in the reference's idiom
 * over its `tensor<T>`, tensor.cuh:17-595) forwards to these functions; INTEGRATION.md shows
 * the binding a maintainer of the reference adds.
 *
 * Conventions
 *   - plain C: POD structs, pointers and sizes only; no C++ or torch types.
 *   - every function returns a pinn_status (0 = PINN_OK); the reason text for the last
 *     failure on a handle is pinn_last_error(ctx) ("<fn>: <reason>", the reference's
 *     "<class>::<fn>: <reason>" convention, tensor.cuh:88, init_tensor.h:67).  Codes 0..2
 *     are the reference's device::constants::error (constants.h:27-31).
 *   - nothing throws across the ABI; there is NO CPU fallback: without a CUDA device
 *     pinn_create fails with PINN_ERR_NO_DEVICE.
 *   - one handle = one GPU = one host thread at a time; the handle owns all device memory,
 *     its stream, its CUDA graph and (N>1) its NCCL communicator.
 *   - matrices are row-major, batch as rows: points [N, d_in], weights W[l] are
 *     [fan_in, fan_out] (in x out), biases [fan_out] — the layout of the reference matmul
 *     kernel (kernel.cuh:68-83) and its suffix-broadcast bias add (kernel.cuh:10-17).
 *   - flattened parameter order: W1 (row-major), b1, W2, b2, ..., W_{L+1}, b_{L+1}.
 */
#ifndef PINN_ABI_H
#define PINN_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PINN_ABI_VERSION 1

/* ---- status codes: 0..2 mirror reference device::constants::error (constants.h:27-31) ---- */
typedef enum pinn_status {
    PINN_OK                   = 0,  /* NORMAL */
    PINN_ERR_DIVISION_BY_ZERO = 1,  /* DIVISION_BY_ZERO (kept for enum parity; unused here) */
    PINN_ERR_DEVICE_OVERFLOW  = 2,  /* DEVICE_OVERFLOW: non-finite loss seen on device */
    PINN_ERR_INVALID_ARGUMENT = 3,  /* std::invalid_argument in the reference's host checks */
    PINN_ERR_NO_DEVICE        = 4,  /* no CUDA device / wrong architecture: no CPU fallback */
    PINN_ERR_CUDA             = 5,  /* a CUDA runtime call failed (the reference ignores these) */
    PINN_ERR_NCCL             = 6,  /* an NCCL call failed or libnccl is not loadable */
    PINN_ERR_UNSUPPORTED      = 7   /* configuration outside the compiled kernel set */
} pinn_status;

/* ---- PDE families (builder-defined, the reference ships none) ---- */
typedef enum pinn_pde {
    PINN_PDE_BURGERS       = 0, /* 1-D viscous Burgers (x,t):   r = u_t + u u_x - (0.01/pi) u_xx */
    PINN_PDE_HELMHOLTZ     = 1, /* 2-D Helmholtz (x,y):         r = u_xx + u_yy + k^2 u - q(x,y) */
    PINN_PDE_ALLEN_CAHN    = 2, /* 2-D Allen-Cahn (x,y,t):      r = u_t - 1e-4 (u_xx+u_yy) + 5u^3 - 5u */
    PINN_PDE_NAVIER_STOKES = 3, /* 2-D incompressible NS (x,y,t)->(u,v,p), nu = 0.01 */
    PINN_PDE_HEAT          = 4  /* 2-D heat (x,y,t):            r = u_t - 0.05 (u_xx+u_yy) */
} pinn_pde;

/* ---- arithmetic path of the dense contractions ---- */
typedef enum pinn_precision {
    PINN_PREC_FP32   = 0, /* FP32 CUDA-core FMA (parity path; all widths) */
    PINN_PREC_BF16X3 = 1, /* tcgen05 kind::f16, each operand split into two BF16 terms (hi*hi + hi*mid +
                             mid*hi), FP32 accumulate in TMEM: near-FP32 products; width 64 or 128 */
    PINN_PREC_BF16   = 2, /* tcgen05 kind::f16, BF16 operands, FP32 accumulate in TMEM; width 64/128/256 */
    PINN_PREC_BF16X6 = 3, /* width 20 (2 -> [20 x L] -> 1): warp-level mma.sync, each operand split into THREE BF16 terms
                             (an exact 24-bit split), the six cross terms of order <= 2, FP32 accumulate: FP32-equivalent
                             products with the activations chained from layer to layer in registers.  PINN_PREC_BF16X3 at
                             width 20 selects the same kernel with two terms (three MMAs). */
    PINN_PREC_FP16X3 = 4  /* width 20: the same kernel with each operand split into TWO FP16 terms (22 significand bits, three
                             MMAs) after an exact per-warp, per-channel power-of-two scaling into the FP16 range: near-FP32
                             products at the cost of BF16X3. */
} pinn_precision;

/* Point sets of one training batch. */
typedef enum pinn_point_set {
    PINN_SET_INTERIOR = 0,  /* collocation points: all Taylor channels, PDE residual */
    PINN_SET_BOUNDARY = 1,  /* value channel only, Dirichlet target */
    PINN_SET_INITIAL  = 2   /* value channel only, t = 0 target */
} pinn_point_set;

/*
 * Problem + run configuration.  Zero-initialise, then fill; pinn_config_default() fills the
 * BASELINE configs.  All counts are GLOBAL (over all ranks); rank r of nranks owns indices
 * [floor(r*n/nranks), floor((r+1)*n/nranks)) of every set.
 */
typedef struct pinn_config {
    int32_t  abi_version;   /* PINN_ABI_VERSION */
    int32_t  pde;           /* pinn_pde */
    int32_t  d_in;          /* network inputs: 2 (x,t | x,y) or 3 (x,y,t) */
    int32_t  d_out;         /* network outputs: 1, or 3 for Navier-Stokes (u,v,p) */
    int32_t  width;         /* hidden width H */
    int32_t  depth;         /* number of hidden tanh layers L (parameters: L+1 affine maps) */
    int32_t  precision;     /* pinn_precision */
    int32_t  device;        /* CUDA device ordinal */
    int32_t  rank;          /* data-parallel rank, 0 <= rank < nranks */
    int32_t  nranks;        /* data-parallel world size (1 = no collective) */
    uint64_t n_interior;    /* N_f, global */
    uint64_t n_boundary;    /* N_b, global */
    uint64_t n_initial;     /* N_0, global (0 for steady problems) */
    uint64_t seed_interior; /* counter-based sampler seeds */
    uint64_t seed_boundary;
    uint64_t seed_initial;
    uint64_t seed_weights;  /* layer l uses seed_weights + l */
    float    lr, beta1, beta2, eps;         /* Adam */
    float    lambda_r, lambda_b, lambda_0;  /* loss weights */
    float    pde_param;     /* Helmholtz k (default 1); others: 0 = canonical constant */
} pinn_config;

/* Per-step result: the three un-weighted MSE terms and the weighted total. */
typedef struct pinn_step_out {
    float    loss;          /* lambda_r*mse_r + lambda_b*mse_b + lambda_0*mse_0 */
    float    mse_residual;  /* mean over N_f of sum_k r_k^2 */
    float    mse_boundary;  /* mean over N_b of sum_o (y_o - g_o)^2 */
    float    mse_initial;   /* mean over N_0 of sum_o (y_o - g_o)^2 */
    uint64_t step;          /* Adam step counter t after this step (from 1) */
} pinn_step_out;

typedef struct pinn_ctx pinn_ctx;   /* opaque */

/* ---------------------------------------------------------------------------------------- */
/* configuration helpers (host only)                                                        */
/* ---------------------------------------------------------------------------------------- */

/* Fill *cfg with BASELINE config `which` in 1..5 (C1..C5 of BASELINE.md; N_b = N_0 =
 * max(256, N_f/32), seeds 0x5EED0001/2/3 and 0x5EED0100, Adam 1e-3/(0.9,0.999)/1e-8). */
int pinn_config_default(pinn_config* cfg, int which);

/* Number of Taylor channels C = 1 + d_in + n_second the PDE carries (4, 5 or 6). */
int pinn_num_channels(const pinn_config* cfg);

/* Total parameter count P = sum_l (fan_in*fan_out + fan_out). */
size_t pinn_num_params(const pinn_config* cfg);

/* Shard [begin,end) of a global count n owned by `rank` of `nranks` (bit-exact contract). */
void pinn_shard_range(uint64_t n, int32_t rank, int32_t nranks, uint64_t* begin, uint64_t* end);

/* ---------------------------------------------------------------------------------------- */
/* lifetime                                                                                 */
/* ---------------------------------------------------------------------------------------- */

/* Replaces: the `network` constructor the reference never wrote (network.cuh, 0 bytes) — what
 * would have been one tensor<T> per W/b, each with 3 cudaMalloc + 3 H2D (tensor.cuh:43-56,
 * interface.cuh:94-118).  Allocates every device buffer once, Glorot-initialises the
 * parameters from seed_weights on the host (bit-exact with the oracle) and uploads them. */
int pinn_create(pinn_ctx** out, const pinn_config* cfg);
void pinn_destroy(pinn_ctx* ctx);
const char* pinn_last_error(const pinn_ctx* ctx);   /* ctx may be NULL: last create error */

/* ---------------------------------------------------------------------------------------- */
/* parameters (host <-> device): the tensor<T> host mirror of the reference, made lazy      */
/* ---------------------------------------------------------------------------------------- */

/* Replaces tensor<T>::assign / host_dirty flush (tensor.cuh:72-77,271-290). n must be P.
 * Resets the Adam moments and the step counter when reset_optimizer != 0. */
int pinn_set_params(pinn_ctx* ctx, const float* flat, size_t n, int reset_optimizer);
/* Replaces the D2H copy after every op (tensor.cuh:107,406). n must be P. */
int pinn_get_params(pinn_ctx* ctx, float* flat, size_t n);
/* Gradient of the last pinn_loss_grad call (global mean gradient, after the all-reduce). */
int pinn_get_grads(pinn_ctx* ctx, float* flat, size_t n);
/* Adam moments + step (checkpoint / "identical initialisation" parity runs). */
int pinn_get_optimizer(pinn_ctx* ctx, float* m, float* v, size_t n, uint64_t* step);
int pinn_set_optimizer(pinn_ctx* ctx, const float* m, const float* v, size_t n, uint64_t step);

/* ---------------------------------------------------------------------------------------- */
/* collocation points (reference src/utils.cu, 0 bytes)                                     */
/* ---------------------------------------------------------------------------------------- */

/* Generate this rank's shard of all three point sets ON DEVICE from the config seeds.
 * Point i, coordinate d of a set is a pure function of (seed, i*d_in + d): bit-identical on
 * the CPU oracle, on the GPU and for any nranks. */
int pinn_sample_points(pinn_ctx* ctx);
/* Re-seed (e.g. resampling every k steps) and regenerate. */
int pinn_resample_points(pinn_ctx* ctx, uint64_t seed_interior, uint64_t seed_boundary,
                         uint64_t seed_initial);
/* Upload caller points for one set: x is HOST memory, row-major [n, d_in], the LOCAL shard.
 * Replaces tensor<float> points = {...} + H2D (tensor.cuh:146-151).  n may not exceed the
 * shard size fixed at create time. */
int pinn_set_points(pinn_ctx* ctx, int set, const float* x, size_t n);
/* The same for points that already live on this handle's GPU: d_x is DEVICE memory — e.g. the device half of a
 * reference tensor<float>, device::tensor::interface<T>::d_var().data (include/device/interface.cuh:248-249), which
 * tensor::sync_device keeps current (tensor.cuh:72-82).  Device-to-device: no host round trip, no H2D bytes.
 * A host pointer is refused with PINN_ERR_INVALID_ARGUMENT. */
int pinn_set_points_device(pinn_ctx* ctx, int set, const float* d_x, size_t n);
/* ... and from the reference's kernel-argument view itself, device::tensor::variables<float>
 * (include/device/variables.cuh:6-52): the 48-byte POD every reference kernel takes by value (kernel.cuh:11) —
 * {T* data @0, size_t* shape @8, size_t* stride @16, size_t dim @24, size_t n @32, bool transposed @40}.
 * `variables48` points at such a struct in HOST memory (as d_var() returns it); its data pointer is device memory
 * holding n = n_points * d_in floats, row-major; transposed views are refused. */
int pinn_set_points_view(pinn_ctx* ctx, int set, const void* variables48);
/* Host->device bytes this handle has copied for point sets so far (pinn_set_points, pinn_train_step_host's copy
 * path): lets a caller — and the tests — verify that the device entry points move nothing over PCIe. */
unsigned long long pinn_h2d_point_bytes(const pinn_ctx* ctx);
/* Copy this rank's shard of a set back to the host, row-major [n, d_in]. */
int pinn_get_points(pinn_ctx* ctx, int set, float* x, size_t n);
/* Local shard size of a set on this rank. */
size_t pinn_local_count(const pinn_ctx* ctx, int set);

/* ---------------------------------------------------------------------------------------- */
/* the hot path                                                                             */
/* ---------------------------------------------------------------------------------------- */

/* One full training step on the resident points: forward (all Taylor channels), residual +
 * BC + IC loss, reverse-mode gradients, (N>1: one all-reduce of the flat [P+3] buffer),
 * Adam.  Replaces what the reference would compose from tensor::matmul (tensor.cuh:380-409)
 * + add (331-333) + mul/sub (335-341) + dot (439-465) + transpose (510-592) with a blocking
 * D2H after every op.  `out` (HOST, may be NULL) receives the loss terms of THIS step's
 * pre-update parameters; passing NULL skips the device->host read (fully asynchronous). */
int pinn_train_step(pinn_ctx* ctx, pinn_step_out* out);

/* Same step, end to end from HOST buffers: uploads the caller's interior points (row-major
 * [n, d_in], local shard; pinned memory makes the copy asynchronous), runs the step and
 * reads the loss back.  This is the call bench.py's `e2e` times. */
int pinn_train_step_host(pinn_ctx* ctx, const float* x_interior, size_t n, pinn_step_out* out);
/* The same step from interior points that are already on the device (see pinn_set_points_device). */
int pinn_train_step_device(pinn_ctx* ctx, const float* d_x_interior, size_t n, pinn_step_out* out);

/* k steps back to back without host synchronisation (CUDA-graph replay); losses of the last
 * step are returned. */
int pinn_train_steps(pinn_ctx* ctx, int k, pinn_step_out* out);

/* Loss + gradient only (no Adam): fills the gradient buffer (pinn_get_grads) and *out. */
int pinn_loss_grad(pinn_ctx* ctx, pinn_step_out* out);

/* Inference: y[n, d_out] = network(x[n, d_in]), HOST buffers, value channel only.
 * Replaces the chain of tensor::matmul + add the reference would use (tensor.cuh:380-409). */
int pinn_forward(pinn_ctx* ctx, const float* x, float* y, size_t n);

/* Forward with all Taylor channels and the PDE residual for caller points (HOST buffers):
 * y is [n, d_out, C] (value, first derivatives, second derivatives), r is [n, n_res]
 * (n_res = 3 for Navier-Stokes, else 1).  Either may be NULL. */
int pinn_residual(pinn_ctx* ctx, const float* x, float* y, float* r, size_t n);

/* Block until all work queued on the handle's stream is complete. */
int pinn_synchronize(pinn_ctx* ctx);

/* ---------------------------------------------------------------------------------------- */
/* multi-GPU (new layer: the reference is single-GPU)                          */
/* ---------------------------------------------------------------------------------------- */

#define PINN_NCCL_ID_BYTES 128
/* Rank 0 creates the id; the caller distributes the 128 bytes (torch.distributed / MPI /
 * file); every rank then calls pinn_comm_init with its own ctx (cfg.rank / cfg.nranks). */
#define PINN_COMM_ID_BYTES 128     /* ncclUniqueId */
int pinn_comm_unique_id(void* id128);
int pinn_comm_init(pinn_ctx* ctx, const void* id128);

/* Optional fused collective (NVLink peer memory, one process per GPU on one box): every rank exports the
 * IPC handle of its exchange buffer, the caller gathers the handles in rank order (64 bytes each) and
 * every rank opens them.  From then on pinn_train_step replaces "reduce kernel + ncclAllReduce + Adam
 * kernel" by ONE kernel per rank that sums its partial gradient rows, publishes the 4*(P+3)-byte sum,
 * reads every peer's sum over NVLink in rank order (bit-identical result on all ranks) and applies
 * Adam.  pinn_comm_init must have succeeded first (it validates the world). */
#define PINN_IPC_HANDLE_BYTES 64
int pinn_comm_p2p_export(pinn_ctx* ctx, void* handle64);
int pinn_comm_p2p_open(pinn_ctx* ctx, const void* handles, int count);

/* Alternative when the caller owns the collective (e.g. torch.distributed): split the step.
 * pinn_step_local computes local gradient/loss sums into the device buffer returned by
 * pinn_grad_buffer ([P+3] floats: gradients in parameter order, then sum r^2, sum bc^2,
 * sum ic^2); the caller all-reduces (sum) that buffer on `stream`-ordered work, then
 * pinn_step_apply runs Adam. */
int pinn_step_local(pinn_ctx* ctx);
int pinn_step_apply(pinn_ctx* ctx, pinn_step_out* out);
void* pinn_grad_buffer(pinn_ctx* ctx, size_t* n_floats);   /* DEVICE pointer */
void* pinn_stream(pinn_ctx* ctx);                           /* cudaStream_t */

/* ---------------------------------------------------------------------------------------- */
/* checkpoint + evaluation (the reference has neither)                    */
/* ---------------------------------------------------------------------------------------- */

/* Flat little-endian binary checkpoint: 64-byte header {"PINNCKPT", u32 version = 1, i32 pde, d_in,
 * d_out, width, depth, u32 reserved[2], u64 P, u64 step, 8 reserved bytes} followed by params[P], m[P],
 * v[P] as float32.  What tensor<T>::raw() + fwrite of every W/b would be in the reference
 * (init_tensor.h:303), as one file.  pinn_load checks the header against the handle's network
 * (PINN_ERR_INVALID_ARGUMENT on any mismatch, the handle is left untouched) and restores parameters,
 * Adam moments and the step counter: a resumed run continues bit-exactly. */
int pinn_save(pinn_ctx* ctx, const char* path);
int pinn_load(pinn_ctx* ctx, const char* path);

/* Analytic solution u*(x) of the PDE where Appendix B has one (HOST buffers, host arithmetic in double,
 * x is [n, d_in], u is [n, d_out]): Helmholtz sin(pi x) sin(4 pi y); heat sin(pi x) sin(pi y)
 * exp(-2 alpha pi^2 t); Navier-Stokes Taylor-Green (u, v, p).  Burgers and Allen-Cahn have no closed
 * form here: PINN_ERR_UNSUPPORTED. */
int pinn_exact_solution(const pinn_config* cfg, const float* x, float* u, size_t n);

typedef struct pinn_eval_out {
    uint64_t n;              /* grid nodes evaluated (per_axis ^ d_in) */
    double   rel_l2[3];      /* per output: ||y - u*||_2 / ||u*||_2 over the grid */
    double   max_abs[3];     /* per output: max |y - u*| */
    double   mse_residual;   /* mean over the grid of sum_k r_k^2 (PDE residual of the network) */
} pinn_eval_out;

/* Evaluate the network on the regular grid of per_axis nodes per input axis over the PDE's box
 * (row-major, last axis fastest, end points included): forward + residual on the device (the same
 * kernels as pinn_residual), error norms against pinn_exact_solution accumulated in double on the
 * host.  For PDEs without an analytic solution rel_l2/max_abs are NaN and only mse_residual is set. */
int pinn_evaluate(pinn_ctx* ctx, int per_axis, pinn_eval_out* out);

/* ---------------------------------------------------------------------------------------- */
/* measurement hooks (bench.py)                                                             */
/* ---------------------------------------------------------------------------------------- */

typedef struct pinn_profile {
    uint64_t launches;          /* kernels launched by this handle since the last reset */
    double   step_kernel_ms;    /* summed CUDA-event time of the fused interior kernel */
    uint64_t step_kernel_count; /* how many launches that sum covers */
    int32_t  grid, block;       /* launch shape of the fused interior kernel */
    int32_t  smem_bytes;        /* dynamic shared memory per CTA */
    int32_t  kernel_id;         /* which kernel variant was selected */
} pinn_profile;

/* 1: the last train step replayed the handle's CUDA graph; 0: steps are launched kernel by kernel (before the first step, while
 * per-kernel profiling is on, or because capture failed — then *note (valid until the handle is destroyed; may be NULL) says
 * why instead of the downgrade being silent); -1: null handle. */
int pinn_graph_state(const pinn_ctx* ctx, const char** note);

/* enable != 0: bracket each fused-kernel launch with CUDA events on the launch stream. */
int pinn_profile_enable(pinn_ctx* ctx, int enable);
int pinn_profile_get(pinn_ctx* ctx, pinn_profile* out, int reset);

#ifdef __cplusplus
}
#endif
#endif /* PINN_ABI_H */
