/*
 * pinn_oracle_body.h — arithmetic of the oracle, compiled twice by pinn_oracle.c:
 * REAL=double (SFX=_f64) and REAL=float (SFX=_f32).  TEST INFRASTRUCTURE ONLY; see
 * pinn_oracle.h for the "parity unpinned" statement and the reference lines followed.
 *
 * Math: SURVEY.md Appendix B.  Layout: batch as rows, W[l] is [fan_in, fan_out] row-major
 * (reference kernel.cuh:68-83: R[i*J+j] = sum_k A[i*K+k]*B[k*J+j], k ascending from a
 * zero-initialised accumulator), bias added by suffix broadcast (kernel.cuh:10-17).
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)

typedef struct FN(ws) {
    REAL* A;     /* [(L+1)][C][Hm]  layer outputs; slot 0 = network inputs */
    REAL* Z;     /* [(L+1)][C][Hm]  pre-activations of hidden layers 1..L (slot 0 unused) */
    REAL* Abar;  /* [C][Hm] */
    REAL* Zbar;  /* [C][Hm] */
    REAL  y[3 * 6], ybar[3 * 6], r[3];
} FN(ws);

static FN(ws)* FN(ws_new)(const odims* dv) {
    FN(ws)* w = (FN(ws)*)calloc(1, sizeof(FN(ws)));
    size_t slab = (size_t)dv->C * dv->Hm;
    w->A = (REAL*)calloc((size_t)(dv->L + 1) * slab, sizeof(REAL));
    w->Z = (REAL*)calloc((size_t)(dv->L + 1) * slab, sizeof(REAL));
    w->Abar = (REAL*)calloc(slab, sizeof(REAL));
    w->Zbar = (REAL*)calloc(slab, sizeof(REAL));
    return w;
}
static void FN(ws_free)(FN(ws)* w) {
    free(w->A); free(w->Z); free(w->Abar); free(w->Zbar); free(w);
}

/* Dirichlet / initial targets g_o(x): SURVEY.md Appendix B "PDEs" (analytic / manufactured). */
static void FN(target)(const odims* dv, const pinn_config* cfg, int set, const float* xf, REAL* g) {
    const REAL pi = (REAL)PINN_PI;
    REAL x = (REAL)xf[0], y = dv->d_in > 1 ? (REAL)xf[1] : 0, t = dv->d_in > 2 ? (REAL)xf[2] : 0;
    switch (cfg->pde) {
    case PINN_PDE_BURGERS:      /* x = (x, t): IC u(x,0) = -sin(pi x); BC u(+-1,t) = 0 */
        g[0] = (set == PINN_SET_INITIAL) ? -RSIN(pi * x) : (REAL)0;
        break;
    case PINN_PDE_HELMHOLTZ:    /* homogeneous Dirichlet */
        g[0] = (REAL)0;
        break;
    case PINN_PDE_ALLEN_CAHN:   /* u0 = x^2 cos(pi x) y^2 cos(pi y); Dirichlet data = u0 on the edge */
        g[0] = x * x * RCOS(pi * x) * (y * y * RCOS(pi * y));
        break;
    case PINN_PDE_HEAT:         /* u = exp(-2 pi^2 alpha t) sin(pi x) sin(pi y): IC at t=0, BC 0 */
        g[0] = (set == PINN_SET_INITIAL) ? RSIN(pi * x) * RSIN(pi * y) : (REAL)0;
        break;
    case PINN_PDE_NAVIER_STOKES: { /* Taylor-Green vortex, nu = 0.01 */
        const REAL nu = (REAL)PINN_NS_NU;
        REAL e2 = REXP((REAL)-2 * nu * t), e4 = REXP((REAL)-4 * nu * t);
        g[0] = -RCOS(x) * RSIN(y) * e2;
        g[1] = RSIN(x) * RCOS(y) * e2;
        g[2] = (REAL)-0.25 * (RCOS((REAL)2 * x) + RCOS((REAL)2 * y)) * e4;
        break; }
    default: g[0] = 0;
    }
}

/* PDE residuals r_q from the output Taylor channels y[o*C + c], and the seeds
 * ybar[o*C+c] = scale * sum_q r_q * d r_q / d y[o][c].  Channel order: 0 value, 1+i d/dx_i
 * (i < d_in), 1+d_in+i d2/dx_i^2 (i < n2). */
static void FN(residual)(const odims* dv, const pinn_config* cfg, const float* xf,
                         const REAL* y, REAL* r, REAL scale, REAL* ybar) {
    const int C = dv->C;
    const REAL pi = (REAL)PINN_PI;
    for (int i = 0; i < dv->d_out * C; i++) ybar[i] = 0;
    switch (cfg->pde) {
    case PINN_PDE_BURGERS: {   /* (x,t): u, u_x, u_t, u_xx */
        const REAL nu = (REAL)PINN_BURGERS_NU;
        REAL u = y[0], ux = y[1], ut = y[2], uxx = y[3];
        r[0] = ut + u * ux - nu * uxx;
        REAL s = scale * r[0];
        ybar[0] = s * ux; ybar[1] = s * u; ybar[2] = s; ybar[3] = -(s * nu);
        break; }
    case PINN_PDE_HELMHOLTZ: { /* (x,y): u, u_x, u_y, u_xx, u_yy */
        REAL k = cfg->pde_param != 0.0f ? (REAL)cfg->pde_param : (REAL)1;
        REAL k2 = k * k;
        REAL q = (k2 - pi * pi - (REAL)16 * pi * pi) *
                 (RSIN(pi * (REAL)xf[0]) * RSIN((REAL)4 * pi * (REAL)xf[1]));
        r[0] = y[3] + y[4] + k2 * y[0] - q;
        REAL s = scale * r[0];
        ybar[0] = s * k2; ybar[3] = s; ybar[4] = s;
        break; }
    case PINN_PDE_ALLEN_CAHN: { /* (x,y,t): u, u_x, u_y, u_t, u_xx, u_yy */
        const REAL gam = (REAL)PINN_AC_GAMMA;
        REAL u = y[0];
        r[0] = y[3] - gam * (y[4] + y[5]) + (REAL)5 * u * u * u - (REAL)5 * u;
        REAL s = scale * r[0];
        ybar[0] = s * ((REAL)15 * u * u - (REAL)5); ybar[3] = s;
        ybar[4] = -(s * gam); ybar[5] = -(s * gam);
        break; }
    case PINN_PDE_HEAT: {
        const REAL al = (REAL)PINN_HEAT_ALPHA;
        r[0] = y[3] - al * (y[4] + y[5]);
        REAL s = scale * r[0];
        ybar[3] = s; ybar[4] = -(s * al); ybar[5] = -(s * al);
        break; }
    case PINN_PDE_NAVIER_STOKES: { /* outputs u,v,p; channels val,x,y,t,xx,yy */
        const REAL nu = (REAL)PINN_NS_NU;
        const REAL *U = y, *V = y + C, *Pp = y + 2 * C;
        r[0] = U[3] + U[0] * U[1] + V[0] * U[2] + Pp[1] - nu * (U[4] + U[5]);
        r[1] = V[3] + U[0] * V[1] + V[0] * V[2] + Pp[2] - nu * (V[4] + V[5]);
        r[2] = U[1] + V[2];
        REAL s0 = scale * r[0], s1 = scale * r[1], s2 = scale * r[2];
        REAL *Ub = ybar, *Vb = ybar + C, *Pb = ybar + 2 * C;
        Ub[0] = s0 * U[1] + s1 * V[1];
        Ub[1] = s0 * U[0] + s2;
        Ub[2] = s0 * V[0];
        Ub[3] = s0;
        Ub[4] = -(s0 * nu); Ub[5] = -(s0 * nu);
        Vb[0] = s0 * U[2] + s1 * V[2];
        Vb[1] = s1 * U[0];
        Vb[2] = s1 * V[0] + s2;
        Vb[3] = s1;
        Vb[4] = -(s1 * nu); Vb[5] = -(s1 * nu);
        Pb[1] = s0; Pb[2] = s1;
        break; }
    }
}

/*
 * One point: forward with nc Taylor channels, loss contribution, reverse pass.
 *   set == INTERIOR: nc = C, residual loss.  else nc = 1, (y - g)^2 loss.
 *   grad (may be NULL): += d/dparams of  scale/2 * sum(err^2)  (scale = 2*lambda/N)
 *   *sum_sq += sum of squared errors of this point
 */
static void FN(point)(const odims* dv, const pinn_config* cfg, const REAL* params,
                      const float* xf, int set, REAL scale, REAL* grad, REAL* sum_sq,
                      REAL* y_out, REAL* r_out, FN(ws)* w) {
    const int C = dv->C, Hm = dv->Hm, L = dv->L, H = dv->H, n1 = dv->n1, n2 = dv->n2;
    const int nc = (set == PINN_SET_INTERIOR) ? C : 1;
    const size_t slab = (size_t)C * Hm;

    /* inputs: value channel = coordinates, first tangents = unit vectors, second = 0 */
    REAL* A0 = w->A;
    for (int c = 0; c < nc; c++)
        for (int k = 0; k < dv->d_in; k++)
            A0[c * Hm + k] = (c == 0) ? (REAL)xf[k] : ((c == 1 + k) ? (REAL)1 : (REAL)0);

    /* hidden layers */
    for (int l = 1; l <= L; l++) {
        const int fi = dv->fi[l], fo = dv->fo[l];
        const REAL* W = params + dv->offW[l];
        const REAL* b = params + dv->offB[l];
        const REAL* Ain = w->A + (size_t)(l - 1) * slab;
        REAL* Aout = w->A + (size_t)l * slab;
        REAL* Zl = w->Z + (size_t)l * slab;
        for (int c = 0; c < nc; c++)
            for (int j = 0; j < fo; j++) {
                REAL acc = 0;                                   /* kernel.cuh:77 "T sum = 0" */
                for (int k = 0; k < fi; k++) acc += Ain[c * Hm + k] * W[(size_t)k * fo + j];
                Zl[c * Hm + j] = (c == 0) ? acc + b[j] : acc;    /* bias on the value channel only */
            }
        for (int j = 0; j < fo; j++) {
            REAL a = RTANH(Zl[j]);
            REAL s = (REAL)1 - a * a;
            Aout[j] = a;
            if (nc > 1) {
                for (int i = 0; i < n1; i++) Aout[(1 + i) * Hm + j] = s * Zl[(1 + i) * Hm + j];
                for (int i = 0; i < n2; i++) {
                    REAL zp = Zl[(1 + i) * Hm + j], zpp = Zl[(1 + n1 + i) * Hm + j];
                    Aout[(1 + n1 + i) * Hm + j] = s * zpp - (REAL)2 * a * s * zp * zp;
                }
            }
        }
    }
    /* linear output layer */
    {
        const int fo = dv->d_out;
        const REAL* W = params + dv->offW[L + 1];
        const REAL* b = params + dv->offB[L + 1];
        const REAL* Ain = w->A + (size_t)L * slab;
        for (int o = 0; o < fo; o++)
            for (int c = 0; c < nc; c++) {
                REAL acc = 0;
                for (int k = 0; k < H; k++) acc += Ain[c * Hm + k] * W[(size_t)k * fo + o];
                w->y[o * C + c] = (c == 0) ? acc + b[o] : acc;
            }
    }
    if (y_out) for (int o = 0; o < dv->d_out; o++) for (int c = 0; c < nc; c++) y_out[o * nc + c] = w->y[o * C + c];

    /* loss term and seeds */
    if (set == PINN_SET_INTERIOR) {
        FN(residual)(dv, cfg, xf, w->y, w->r, scale, w->ybar);
        for (int q = 0; q < dv->n_res; q++) { *sum_sq += w->r[q] * w->r[q]; if (r_out) r_out[q] = w->r[q]; }
    } else {
        REAL g[3];
        FN(target)(dv, cfg, set, xf, g);
        for (int o = 0; o < dv->d_out; o++) {
            REAL e = w->y[o * C] - g[o];
            *sum_sq += e * e;
            w->ybar[o * C] = scale * e;
        }
    }
    if (!grad) return;

    /* reverse: output layer */
    {
        const int fo = dv->d_out;
        const REAL* W = params + dv->offW[L + 1];
        REAL* gW = grad + dv->offW[L + 1];
        REAL* gb = grad + dv->offB[L + 1];
        const REAL* Ain = w->A + (size_t)L * slab;
        for (int k = 0; k < H; k++)
            for (int o = 0; o < fo; o++) {
                REAL acc = 0;
                for (int c = 0; c < nc; c++) acc += Ain[c * Hm + k] * w->ybar[o * C + c];
                gW[(size_t)k * fo + o] += acc;
            }
        for (int o = 0; o < fo; o++) gb[o] += w->ybar[o * C];
        for (int c = 0; c < nc; c++)
            for (int k = 0; k < H; k++) {
                REAL acc = 0;
                for (int o = 0; o < fo; o++) acc += w->ybar[o * C + c] * W[(size_t)k * fo + o];
                w->Abar[c * Hm + k] = acc;
            }
    }
    /* reverse: hidden layers L..1 */
    for (int l = L; l >= 1; l--) {
        const int fi = dv->fi[l], fo = dv->fo[l];
        const REAL* W = params + dv->offW[l];
        REAL* gW = grad + dv->offW[l];
        REAL* gb = grad + dv->offB[l];
        const REAL* Ain = w->A + (size_t)(l - 1) * slab;
        const REAL* Al = w->A + (size_t)l * slab;
        const REAL* Zl = w->Z + (size_t)l * slab;
        REAL* Zb = w->Zbar;
        const REAL* Ab = w->Abar;
        for (int j = 0; j < fo; j++) {
            REAL a = Al[j], s = (REAL)1 - a * a;
            REAL zb0 = s * Ab[j];
            if (nc > 1) {
                for (int i = 0; i < n1; i++) {
                    REAL zp = Zl[(1 + i) * Hm + j];
                    REAL abp = Ab[(1 + i) * Hm + j];
                    REAL zbp = s * abp;
                    zb0 += ((REAL)-2 * a * s * zp) * abp;
                    if (i < n2) {
                        REAL zpp = Zl[(1 + n1 + i) * Hm + j];
                        REAL abpp = Ab[(1 + n1 + i) * Hm + j];
                        Zb[(1 + n1 + i) * Hm + j] = s * abpp;
                        zbp -= (REAL)4 * a * s * zp * abpp;
                        zb0 += ((REAL)-2 * a * s * zpp - (REAL)2 * s * ((REAL)1 - (REAL)3 * a * a) * zp * zp) * abpp;
                    }
                    Zb[(1 + i) * Hm + j] = zbp;
                }
            }
            Zb[j] = zb0;
        }
        for (int k = 0; k < fi; k++)
            for (int j = 0; j < fo; j++) {
                REAL acc = 0;
                for (int c = 0; c < nc; c++) acc += Ain[c * Hm + k] * Zb[c * Hm + j];
                gW[(size_t)k * fo + j] += acc;
            }
        for (int j = 0; j < fo; j++) gb[j] += Zb[j];
        if (l > 1)
            for (int c = 0; c < nc; c++)
                for (int k = 0; k < fi; k++) {
                    REAL acc = 0;
                    for (int j = 0; j < fo; j++) acc += Zb[c * Hm + j] * W[(size_t)k * fo + j];
                    w->Abar[c * Hm + k] = acc;
                }
    }
}

/* Accumulate one point set into per-thread partials (static schedule → deterministic). */
static void FN(accumulate_set)(const odims* dv, const pinn_config* cfg, const REAL* params,
                               const float* x, size_t n, int set, REAL scale,
                               REAL* tgrads /* [T][P] or NULL */, REAL* tsums /* [T][3] */) {
    if (n == 0) return;
    #pragma omp parallel
    {
        int tid = omp_get_thread_num();
        FN(ws)* w = FN(ws_new)(dv);
        REAL* g = tgrads ? tgrads + (size_t)tid * dv->P : NULL;
        REAL ss = 0;
        #pragma omp for schedule(static)
        for (long long i = 0; i < (long long)n; i++)
            FN(point)(dv, cfg, params, x + (size_t)i * dv->d_in, set, scale, g, &ss, NULL, NULL, w);
        tsums[(size_t)tid * 3 + set] += ss;
        FN(ws_free)(w);
    }
}

void FN(pinn_oracle_loss_grad)(const pinn_config* cfg, const REAL* params,
                               const float* xf, size_t nf, const float* xb, size_t nb,
                               const float* x0, size_t n0, REAL* grad, REAL* sums) {
    odims dv; odims_make(&dv, cfg);
    int T = omp_get_max_threads();
    REAL* tg = grad ? (REAL*)calloc((size_t)T * dv.P, sizeof(REAL)) : NULL;
    REAL* ts = (REAL*)calloc((size_t)T * 3, sizeof(REAL));
    /* seed scales 2*lambda/N_global, formed in double and rounded once (same on the GPU host side) */
    REAL sr = cfg->n_interior ? (REAL)(2.0 * (double)cfg->lambda_r / (double)cfg->n_interior) : 0;
    REAL sb = cfg->n_boundary ? (REAL)(2.0 * (double)cfg->lambda_b / (double)cfg->n_boundary) : 0;
    REAL s0 = cfg->n_initial ? (REAL)(2.0 * (double)cfg->lambda_0 / (double)cfg->n_initial) : 0;
    FN(accumulate_set)(&dv, cfg, params, xf, nf, PINN_SET_INTERIOR, sr, tg, ts);
    FN(accumulate_set)(&dv, cfg, params, xb, nb, PINN_SET_BOUNDARY, sb, tg, ts);
    FN(accumulate_set)(&dv, cfg, params, x0, n0, PINN_SET_INITIAL, s0, tg, ts);
    for (int q = 0; q < 3; q++) { REAL a = 0; for (int t = 0; t < T; t++) a += ts[(size_t)t * 3 + q]; sums[q] = a; }
    if (grad) {
        #pragma omp parallel for schedule(static)
        for (long long i = 0; i < (long long)dv.P; i++) {
            REAL a = 0;
            for (int t = 0; t < T; t++) a += tg[(size_t)t * dv.P + i];
            grad[i] = a;
        }
    }
    free(tg); free(ts);
}

/* Adam, SURVEY.md Appendix B: bias corrections formed in double from t, then rounded. */
void FN(pinn_oracle_adam)(const pinn_config* cfg, REAL* params, REAL* m, REAL* v,
                          const REAL* grad, size_t n, uint64_t t) {
    const REAL b1 = (REAL)cfg->beta1, b2 = (REAL)cfg->beta2, lr = (REAL)cfg->lr, eps = (REAL)cfg->eps;
    const REAL c1 = (REAL)(1.0 / (1.0 - pow((double)cfg->beta1, (double)t)));
    const REAL c2 = (REAL)(1.0 / (1.0 - pow((double)cfg->beta2, (double)t)));
    for (size_t i = 0; i < n; i++) {
        REAL g = grad[i];
        REAL mi = b1 * m[i] + ((REAL)1 - b1) * g;
        REAL vi = b2 * v[i] + ((REAL)1 - b2) * (g * g);
        m[i] = mi; v[i] = vi;
        REAL mh = mi * c1, vh = vi * c2;
        params[i] = params[i] - lr * (mh / (RSQRT(vh) + eps));
    }
}

void FN(pinn_oracle_residual)(const pinn_config* cfg, const REAL* params, const float* x,
                              size_t n, REAL* y, REAL* r) {
    odims dv; odims_make(&dv, cfg);
    #pragma omp parallel
    {
        FN(ws)* w = FN(ws_new)(&dv);
        #pragma omp for schedule(static)
        for (long long i = 0; i < (long long)n; i++) {
            REAL ss = 0;
            FN(point)(&dv, cfg, params, x + (size_t)i * dv.d_in, PINN_SET_INTERIOR, (REAL)0, NULL, &ss,
                      y ? y + (size_t)i * dv.d_out * dv.C : NULL, r ? r + (size_t)i * dv.n_res : NULL, w);
        }
        FN(ws_free)(w);
    }
}

double FN(pinn_oracle_train)(const pinn_config* cfg, REAL* params, REAL* m, REAL* v,
                             uint64_t t0, int steps, REAL* losses) {
    odims dv; odims_make(&dv, cfg);
    size_t nf = cfg->n_interior, nb = cfg->n_boundary, n0 = cfg->n_initial;
    float* xf = (float*)malloc((nf ? nf : 1) * dv.d_in * sizeof(float));
    float* xb = (float*)malloc((nb ? nb : 1) * dv.d_in * sizeof(float));
    float* x0 = (float*)malloc((n0 ? n0 : 1) * dv.d_in * sizeof(float));
    pinn_oracle_sample(cfg, PINN_SET_INTERIOR, 0, nf, xf);
    pinn_oracle_sample(cfg, PINN_SET_BOUNDARY, 0, nb, xb);
    pinn_oracle_sample(cfg, PINN_SET_INITIAL, 0, n0, x0);
    REAL* g = (REAL*)malloc(dv.P * sizeof(REAL));
    double t_begin = omp_get_wtime();
    for (int s = 0; s < steps; s++) {
        REAL sums[3];
        FN(pinn_oracle_loss_grad)(cfg, params, xf, nf, xb, nb, x0, n0, g, sums);
        REAL mr = nf ? sums[0] / (REAL)nf : 0, mb = nb ? sums[1] / (REAL)nb : 0, m0 = n0 ? sums[2] / (REAL)n0 : 0;
        if (losses) {
            losses[4 * s + 0] = (REAL)cfg->lambda_r * mr + (REAL)cfg->lambda_b * mb + (REAL)cfg->lambda_0 * m0;
            losses[4 * s + 1] = mr; losses[4 * s + 2] = mb; losses[4 * s + 3] = m0;
        }
        FN(pinn_oracle_adam)(cfg, params, m, v, g, dv.P, t0 + (uint64_t)s + 1);
    }
    double dt = omp_get_wtime() - t_begin;
    free(g); free(xf); free(xb); free(x0);
    return dt;
}

#undef FN
#undef CAT
#undef CAT_
