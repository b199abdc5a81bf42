// ref_step.cu — "reference tensor ops + shim": one PINN train step (Burgers, 2 -> [H x L] -> 1, interior
// collocation points) composed op-by-op from the REFERENCE's own, unmodified tensor<float> operations
// (include/tensor.cuh: matmul / add / sub / mul / div / dot / transpose, each with its own device
// allocation, launch and blocking D2H copy — tensor.cuh:84-110, 380-409).  This is the timing baseline
// of SURVEY.md 8(d) / BASELINE.md B-ref-cuda: what a `network` written in the reference's architecture
// costs.  It is NOT the reference's train step — the reference has none (network.cu is 0 bytes).
//
// SHIM (not reference code): the reference has no unary op at all, so tanh() and sqrt() are applied
// on the HOST mirror between ops (the reference keeps the host copy authoritative after every op,
// tensor.cuh:107, so this adds no transfer it would not do anyway).  Boundary / initial terms are
// omitted (1.5 % of the work): the printed rate is an upper bound for the reference-style step.
//
//   usage: ref_step [N=4096] [H=20] [L=8] [steps=3]     prints loss and points/s
// Built by oracle/Makefile from /root/reference/include where it lies.  Test infrastructure only.
#include<tensor.cuh>
#include<chrono>
#include<cmath>
#include<cstdint>
#include<cstdlib>
#include<iostream>
#include<iomanip>
#include<memory>
#include<vector>

using T = tensor<float>;
using P = std::unique_ptr<T>;
template<typename... A> static P mk(A&&... a) { return std::make_unique<T>(std::forward<A>(a)...); }

static uint64_t mix(uint64_t seed, uint64_t ctr) {
    uint64_t z = seed + (ctr + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static float u01(uint64_t seed, uint64_t ctr) { return (float)(mix(seed, ctr) >> 40) * 0x1.0p-24f; }

// ---- shim: unary ops on the host mirror (the reference cannot express them) ----
static T shim_tanh(const T& x) { T y = x; for (size_t i = 0; i < y.size(); i++) y[i] = std::tanh(x.raw()[i]); return y; }
static T shim_sqrt(const T& x) { T y = x; for (size_t i = 0; i < y.size(); i++) y[i] = std::sqrt(x.raw()[i]); return y; }
static T filled(const std::vector<size_t>& shape, float v) { T t(as_shape, shape); for (size_t i = 0; i < t.size(); i++) t[i] = v; return t; }

int main(int argc, char** argv) {
    const size_t N = argc > 1 ? std::strtoul(argv[1], nullptr, 10) : 4096;
    const size_t H = argc > 2 ? std::strtoul(argv[2], nullptr, 10) : 20;
    const size_t L = argc > 3 ? std::strtoul(argv[3], nullptr, 10) : 8;
    const int steps = argc > 4 ? std::atoi(argv[4]) : 3;
    const float nu = 0.003183098861837907f, lr = 1e-3f, b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
    constexpr int C = 4;                                   // u, d/dx, d/dt, d2/dx2

    // parameters (Glorot from the project's counter-based generator), Adam state
    std::vector<P> W, B, mW, vW, mB, vB;
    for (size_t l = 1; l <= L + 1; l++) {
        const size_t fi = l == 1 ? 2 : H, fo = l == L + 1 ? 1 : H;
        const float a = (float)std::sqrt(6.0 / (double)(fi + fo));
        P w = mk(as_shape, std::vector<size_t>{fi, fo});
        for (size_t k = 0; k < fi * fo; k++) (*w)[k] = fmaf(2.0f * a, u01(0x5EED0100ULL + l, k), -a);
        W.push_back(std::move(w)); B.push_back(mk(as_shape, std::vector<size_t>{fo}));
        mW.push_back(mk(as_shape, std::vector<size_t>{fi, fo})); vW.push_back(mk(as_shape, std::vector<size_t>{fi, fo}));
        mB.push_back(mk(as_shape, std::vector<size_t>{fo})); vB.push_back(mk(as_shape, std::vector<size_t>{fo}));
    }
    // collocation points and the input channels: value = (x,t), d/dx = (1,0), d/dt = (0,1), d2/dx2 = 0
    std::vector<P> A0;
    for (int c = 0; c < C; c++) A0.push_back(mk(as_shape, std::vector<size_t>{N, 2}));
    for (size_t i = 0; i < N; i++) {
        (*A0[0])[2 * i] = fmaf(2.0f, u01(0x5EED0001ULL, 2 * i), -1.0f);
        (*A0[0])[2 * i + 1] = u01(0x5EED0001ULL, 2 * i + 1);
        (*A0[1])[2 * i] = 1.0f; (*A0[2])[2 * i + 1] = 1.0f;
    }
    const T ones_row = filled({1, N}, 1.0f);               // column sums as a reference matmul
    const T zero(0.0f), two(2.0f), three(3.0f), nu_t(nu), neg_nu(-nu), seed_scale(2.0f / (float)N);

    double loss = 0.0, seconds = 0.0;
    for (int step = 0; step <= steps; step++) {            // step 0 = warm-up (not timed)
        const auto t0 = std::chrono::steady_clock::now();
        // ---------------- forward ----------------
        std::vector<std::vector<P>> A(L + 1), Z(L + 1);    // A[l][c]: layer outputs; Z[l][c]: pre-activations (c >= 1)
        std::vector<P> Sg(L + 1);                          // s = 1 - a^2
        // (a reference op, not a copy: tensor's copy constructor drops the device mirror of a host-dirty source, tensor.cuh:153-157)
        for (int c = 0; c < C; c++) A[0].push_back(mk(T::add(*A0[c], zero)));
        for (size_t l = 1; l <= L; l++) {
            std::vector<P> z;
            for (int c = 0; c < C; c++) z.push_back(mk(T::matmul(*A[l - 1][c], *W[l - 1])));
            z[0] = mk(T::add(*z[0], *B[l - 1]));           // suffix-broadcast bias (kernel.cuh:10-17)
            P a = mk(shim_tanh(*z[0]));
            P s = mk(T::sub(filled({N, H}, 1.0f), T::mul(*a, *a)));
            A[l].push_back(mk(*a));
            A[l].push_back(mk(T::mul(*s, *z[1])));
            A[l].push_back(mk(T::mul(*s, *z[2])));
            A[l].push_back(mk(T::sub(T::mul(*s, *z[3]), T::mul(T::mul(T::mul(*a, two), *s), T::mul(*z[1], *z[1])))));
            Sg[l] = std::move(s);
            Z[l] = std::move(z);
        }
        std::vector<P> Y;
        for (int c = 0; c < C; c++) Y.push_back(mk(T::matmul(*A[L][c], *W[L])));
        Y[0] = mk(T::add(*Y[0], *B[L]));
        // residual r = u_t + u u_x - nu u_xx and the loss mean(r^2) through the reference dot
        T r = T::sub(T::add(*Y[2], T::mul(*Y[0], *Y[1])), T::mul(*Y[3], nu_t));
        T rflat = r; rflat.reshape({N});
        loss = (double)T::dot(rflat, rflat).raw()[0] / (double)N;
        // ---------------- reverse ----------------
        T sr = T::mul(r, seed_scale);                      // 2 r / N
        std::vector<P> Yb;
        Yb.push_back(mk(T::mul(sr, *Y[1]))); Yb.push_back(mk(T::mul(sr, *Y[0]))); Yb.push_back(mk(sr)); Yb.push_back(mk(T::mul(sr, neg_nu)));
        std::vector<P> gW(L + 1), gB(L + 1), Ab;
        {
            // (accumulate through fresh tensors: tensor's move-assignment leaks / mis-flags the device mirror, interface.cuh:287-305)
            P g = mk(T::matmul(T::transpose(*A[L][0]), *Yb[0]));
            for (int c = 1; c < C; c++) g = mk(T::add(*g, T::matmul(T::transpose(*A[L][c]), *Yb[c])));
            gW[L] = std::move(g);
            T gb = T::matmul(ones_row, *Yb[0]); gb.reshape({1});
            gB[L] = mk(gb);
            for (int c = 0; c < C; c++) Ab.push_back(mk(T::matmul(*Yb[c], T::transpose(*W[L]))));
        }
        for (size_t l = L; l >= 1; l--) {
            const T& a = *A[l][0]; const T& s = *Sg[l];
            const T as2 = T::mul(T::mul(a, two), s);       // 2 a s
            std::vector<P> zb(C);
            zb[3] = mk(T::mul(s, *Ab[3]));
            zb[1] = mk(T::sub(T::mul(s, *Ab[1]), T::mul(T::mul(T::mul(as2, two), *Z[l][1]), *Ab[3])));
            zb[2] = mk(T::mul(s, *Ab[2]));
            // zb0 = s ab0 - 2as z1 ab1 - 2as z2 ab2 - (2as z3 + 2 s (1 - 3 a^2) z1^2) ab3
            T curv = T::mul(T::mul(T::mul(s, two), T::sub(filled({N, H}, 1.0f), T::mul(T::mul(a, a), three))), T::mul(*Z[l][1], *Z[l][1]));
            T z0 = T::sub(T::sub(T::sub(T::mul(s, *Ab[0]), T::mul(T::mul(as2, *Z[l][1]), *Ab[1])), T::mul(T::mul(as2, *Z[l][2]), *Ab[2])),
                          T::mul(T::add(T::mul(as2, *Z[l][3]), curv), *Ab[3]));
            zb[0] = mk(z0);
            P g = mk(T::matmul(T::transpose(*A[l - 1][0]), *zb[0]));
            for (int c = 1; c < C; c++) g = mk(T::add(*g, T::matmul(T::transpose(*A[l - 1][c]), *zb[c])));
            gW[l - 1] = std::move(g);
            T gb = T::matmul(ones_row, *zb[0]); gb.reshape({H});
            gB[l - 1] = mk(gb);
            if (l > 1) for (int c = 0; c < C; c++) Ab[c] = mk(T::matmul(*zb[c], T::transpose(*W[l - 1])));
        }
        // ---------------- Adam (reference elementwise ops + sqrt shim) ----------------
        const float c1 = 1.0f / (1.0f - std::pow(b1, (float)(step + 1))), c2 = 1.0f / (1.0f - std::pow(b2, (float)(step + 1)));
        auto adam = [&](P& w, P& m, P& v, const T& g) {
            m = mk(T::add(T::mul(*m, T(b1)), T::mul(g, T(1.0f - b1))));
            v = mk(T::add(T::mul(*v, T(b2)), T::mul(T::mul(g, g), T(1.0f - b2))));
            T den = T::add(shim_sqrt(T::mul(*v, T(c2))), T(eps));
            w = mk(T::sub(*w, T::mul(T::div(T::mul(*m, T(c1)), den), T(lr))));
        };
        for (size_t l = 0; l <= L; l++) { adam(W[l], mW[l], vW[l], *gW[l]); adam(B[l], mB[l], vB[l], *gB[l]); }
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (step > 0) seconds += dt;
        if (step == 0) std::cout << "loss0=" << std::setprecision(9) << loss << std::endl;
    }
    std::cout << "steps=" << steps << " N=" << N << " H=" << H << " L=" << L << std::endl;
    std::cout << "seconds_per_step=" << seconds / steps << std::endl;
    std::cout << "points_per_second=" << (double)N * steps / seconds << std::endl;
    std::cout << "loss_last=" << std::setprecision(9) << loss << std::endl;
    return 0;
}
