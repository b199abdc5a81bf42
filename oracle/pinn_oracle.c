/*
 * pinn_oracle.c — CPU oracle of the PINN train step.  TEST INFRASTRUCTURE ONLY
 * (see pinn_oracle.h: "parity unpinned" for the train step; reference lines followed for
 * layout only).  Plain C99 + OpenMP; built by oracle/Makefile into oracle/liboracle.so.
 */
#include "pinn_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

/* Constants of SURVEY.md Appendix B (single-precision literals are what the GPU uses). */
#define PINN_PI          3.14159265358979323846
#define PINN_BURGERS_NU  (0.01 / PINN_PI)
#define PINN_AC_GAMMA    1.0e-4
#define PINN_HEAT_ALPHA  0.05
#define PINN_NS_NU       0.01
#define PINN_TWO_PI_F    6.28318530717958647692f

#define OR_MAXL 32

typedef struct odims {
    int d_in, d_out, H, L, C, n1, n2, n_res, n_space, has_t, Hm;
    float lo[3], span[3], hi[3];
    int fi[OR_MAXL + 2], fo[OR_MAXL + 2];
    size_t offW[OR_MAXL + 2], offB[OR_MAXL + 2];
    size_t P;
} odims;

static void odims_make(odims* d, const pinn_config* cfg) {
    memset(d, 0, sizeof(*d));
    d->d_in = cfg->d_in; d->d_out = cfg->d_out; d->H = cfg->width; d->L = cfg->depth;
    d->n1 = cfg->d_in;
    switch (cfg->pde) {
    case PINN_PDE_BURGERS:       d->n_space = 1; d->n2 = 1; d->n_res = 1; break;
    case PINN_PDE_HELMHOLTZ:     d->n_space = 2; d->n2 = 2; d->n_res = 1; break;
    case PINN_PDE_ALLEN_CAHN:    d->n_space = 2; d->n2 = 2; d->n_res = 1; break;
    case PINN_PDE_HEAT:          d->n_space = 2; d->n2 = 2; d->n_res = 1; break;
    case PINN_PDE_NAVIER_STOKES: d->n_space = 2; d->n2 = 2; d->n_res = 3; break;
    default:                     d->n_space = 1; d->n2 = 1; d->n_res = 1; break;
    }
    d->has_t = d->d_in > d->n_space;
    d->C = 1 + d->n1 + d->n2;
    d->Hm = d->H > 8 ? d->H : 8;
    /* domains: Burgers x in [-1,1], t in [0,1]; Helmholtz / Allen-Cahn / heat [-1,1]^2 (x [0,1]);
     * Navier-Stokes [0,2pi]^2 x [0,1] */
    for (int k = 0; k < d->d_in; k++) {
        int is_t = d->has_t && k == d->d_in - 1;
        if (is_t) { d->lo[k] = 0.0f; d->hi[k] = 1.0f; }
        else if (cfg->pde == PINN_PDE_NAVIER_STOKES) { d->lo[k] = 0.0f; d->hi[k] = PINN_TWO_PI_F; }
        else { d->lo[k] = -1.0f; d->hi[k] = 1.0f; }
        d->span[k] = d->hi[k] - d->lo[k];
    }
    size_t off = 0;
    for (int l = 1; l <= d->L + 1; l++) {
        d->fi[l] = (l == 1) ? d->d_in : d->H;
        d->fo[l] = (l == d->L + 1) ? d->d_out : d->H;
        d->offW[l] = off; off += (size_t)d->fi[l] * d->fo[l];
        d->offB[l] = off; off += (size_t)d->fo[l];
    }
    d->P = off;
}

/* ---- sampler: SplitMix64 finaliser over (seed, counter); 24-bit mantissa uniform ---- */
uint64_t pinn_oracle_mix(uint64_t seed, uint64_t ctr) {
    uint64_t z = seed + (ctr + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
float pinn_oracle_u01(uint64_t seed, uint64_t ctr) {
    return (float)(pinn_oracle_mix(seed, ctr) >> 40) * 0x1.0p-24f;
}

void pinn_oracle_sample(const pinn_config* cfg, int set, uint64_t begin, uint64_t end, float* x) {
    odims d; odims_make(&d, cfg);
    const uint64_t seed = set == PINN_SET_INTERIOR ? cfg->seed_interior
                        : set == PINN_SET_BOUNDARY ? cfg->seed_boundary : cfg->seed_initial;
    const int nfaces = 2 * d.n_space;
    #pragma omp parallel for schedule(static)
    for (long long ii = (long long)begin; ii < (long long)end; ii++) {
        uint64_t i = (uint64_t)ii;
        float* p = x + (size_t)(i - begin) * d.d_in;
        int face = (int)(i % (uint64_t)nfaces), axis = face / 2, side = face % 2;
        for (int k = 0; k < d.d_in; k++) {
            float u = pinn_oracle_u01(seed, i * (uint64_t)d.d_in + (uint64_t)k);
            float v = fmaf(d.span[k], u, d.lo[k]);
            if (set == PINN_SET_BOUNDARY && k == axis) v = side ? d.hi[k] : d.lo[k];
            if (set == PINN_SET_INITIAL && d.has_t && k == d.d_in - 1) v = d.lo[k];
            p[k] = v;
        }
    }
}

size_t pinn_oracle_num_params(const pinn_config* cfg) { odims d; odims_make(&d, cfg); return d.P; }
int pinn_oracle_num_channels(const pinn_config* cfg) { odims d; odims_make(&d, cfg); return d.C; }

void pinn_oracle_shard_range(uint64_t n, int32_t rank, int32_t nranks, uint64_t* b, uint64_t* e) {
    /* floor(r*n/R): 128-bit product so 64-bit counts cannot overflow */
    *b = (uint64_t)(((unsigned __int128)n * (unsigned)rank) / (unsigned)nranks);
    *e = (uint64_t)(((unsigned __int128)n * (unsigned)(rank + 1)) / (unsigned)nranks);
}

void pinn_oracle_init_params(const pinn_config* cfg, float* flat) {
    odims d; odims_make(&d, cfg);
    for (int l = 1; l <= d.L + 1; l++) {
        float a = (float)sqrt(6.0 / (double)(d.fi[l] + d.fo[l]));
        float two_a = 2.0f * a;
        size_t nw = (size_t)d.fi[l] * d.fo[l];
        for (size_t k = 0; k < nw; k++)
            flat[d.offW[l] + k] = fmaf(two_a, pinn_oracle_u01(cfg->seed_weights + (uint64_t)l, k), -a);
        for (int j = 0; j < d.fo[l]; j++) flat[d.offB[l] + j] = 0.0f;
    }
}

int pinn_oracle_threads(void) { return omp_get_max_threads(); }
void pinn_oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

/* ---- double instantiation ---- */
#define REAL double
#define SFX _f64
#define RTANH tanh
#define RSIN sin
#define RCOS cos
#define REXP exp
#define RSQRT sqrt
#include "pinn_oracle_body.h"
#undef REAL
#undef SFX
#undef RTANH
#undef RSIN
#undef RCOS
#undef REXP
#undef RSQRT

/* ---- float instantiation ---- */
#define REAL float
#define SFX _f32
#define RTANH tanhf
#define RSIN sinf
#define RCOS cosf
#define REXP expf
#define RSQRT sqrtf
#include "pinn_oracle_body.h"
#undef REAL
#undef SFX

void pinn_oracle_target_f64(const pinn_config* cfg, int set, const float* x, double* g) {
    odims d; odims_make(&d, cfg);
    target_f64(&d, cfg, set, x, g);
}
