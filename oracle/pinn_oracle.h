/*
 * pinn_oracle.h — CPU oracle for the PINN train-step hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library.  The product (cuda-pinn_b200/, include/, src/) never links or calls it.
 *
 * PARITY UNPINNED for the train step: the reference (Deadman621/CUDA-PINN) contains no
 * network, autodiff, loss, optimizer or sampler (include/network.cuh, src/network.cu and
 * src/utils.cu are 0-byte files), so there is no reference result, golden
 * vector or test to pin this restatement on.  What it restates is the builder-defined math
 * of SURVEY.md Appendix B, following the reference only for LAYOUT conventions:
 *   row-major [N,in]x[in,out] products, k ascending   reference include/device/kernel.cuh:68-83
 *   suffix-broadcast bias add                          reference include/device/kernel.cuh:10-17
 *   zero-initialised sum reductions                    reference include/device/kernel.cuh:85-99
 *   C-order strides of the flattened parameters        reference include/host/init_tensor.h:429-441
 * It is cross-validated against torch.autograd (float64, double backward) through the
 * committed fixtures under tests/golden/ (generator: tests/golden/make_golden.py) and against
 * central finite differences (tests/test_oracle.py).  What the reference CAN pin — the
 * tensor<T> host API vectors G1 and the main.cu output G0 — is pinned
 * by oracle/_ref (built from the reference's own headers by oracle/Makefile).
 */
#ifndef PINN_ORACLE_H
#define PINN_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include "../include/pinn_abi.h"   /* pinn_config / enums only (types, no code) */

#ifdef __cplusplus
extern "C" {
#endif

/* ---- counter-based sampler: bit-exact contract with the GPU ---- */
uint64_t pinn_oracle_mix(uint64_t seed, uint64_t ctr);
float    pinn_oracle_u01(uint64_t seed, uint64_t ctr);
/* Fill x[(end-begin), d_in] with points [begin,end) of `set` (pinn_point_set). */
void pinn_oracle_sample(const pinn_config* cfg, int set, uint64_t begin, uint64_t end, float* x);
/* Glorot-uniform weights from seed_weights + layer, zero biases; flat[P]. */
void pinn_oracle_init_params(const pinn_config* cfg, float* flat);
size_t pinn_oracle_num_params(const pinn_config* cfg);
int pinn_oracle_num_channels(const pinn_config* cfg);
void pinn_oracle_shard_range(uint64_t n, int32_t rank, int32_t nranks, uint64_t* b, uint64_t* e);

/*
 * Loss + gradient over explicit point arrays (row-major [n, d_in] floats).
 *   params  flat[P]
 *   n_*_global  the GLOBAL counts the means are taken over (seed scale 2*lambda/N_global);
 *               with local arrays of a shard this yields that shard's contribution.
 *   grad    flat[P]   (d loss / d params contribution of these points), may be NULL
 *   sums    [3]       sum r^2, sum bc^2, sum ic^2 over these points (un-normalised)
 * _f64 computes in double, _f32 in float (explicit single-precision libm calls).
 * Threads: OpenMP over points, per-thread partial sums combined in thread order.
 */
void pinn_oracle_loss_grad_f64(const pinn_config* cfg, const double* params,
                               const float* xf, size_t nf, const float* xb, size_t nb,
                               const float* x0, size_t n0, double* grad, double* sums);
void pinn_oracle_loss_grad_f32(const pinn_config* cfg, const float* params,
                               const float* xf, size_t nf, const float* xb, size_t nb,
                               const float* x0, size_t n0, float* grad, float* sums);

/* Adam: t is the step index AFTER increment (from 1). */
void pinn_oracle_adam_f64(const pinn_config* cfg, double* params, double* m, double* v,
                          const double* grad, size_t n, uint64_t t);
void pinn_oracle_adam_f32(const pinn_config* cfg, float* params, float* m, float* v,
                          const float* grad, size_t n, uint64_t t);

/* Forward with all Taylor channels: y[n, d_out, C], r[n, n_res]; either may be NULL. */
void pinn_oracle_residual_f64(const pinn_config* cfg, const double* params, const float* x,
                              size_t n, double* y, double* r);
void pinn_oracle_residual_f32(const pinn_config* cfg, const float* params, const float* x,
                              size_t n, float* y, float* r);

/* Target g_o(x) of a boundary / initial point (value channel of each output). */
void pinn_oracle_target_f64(const pinn_config* cfg, int set, const float* x, double* g);

/* Full training loop on the oracle's own sampled points (global sets, single "rank"):
 * runs `steps` Adam steps from params/m/v in place; losses[steps*4] = loss, mse_r, mse_b,
 * mse_0 per step (pre-update parameters).  Returns wall seconds spent in the steps. */
double pinn_oracle_train_f32(const pinn_config* cfg, float* params, float* m, float* v,
                             uint64_t t0, int steps, float* losses);
double pinn_oracle_train_f64(const pinn_config* cfg, double* params, double* m, double* v,
                             uint64_t t0, int steps, double* losses);

int pinn_oracle_threads(void);
void pinn_oracle_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
