// tensor_ops_test.cu — the reference's tensor<T> arithmetic and this repo's device_tensor<T> (include/tensor_ops.cuh, kernels in
// libpinn_b200.so) IN ONE PROCESS on the same inputs: compiled against the reference's unmodified headers
// (-I/root/reference/include) by oracle/Makefile.  Prints key=value lines; tests/test_gpu_tensor_ops.py reads them.
#include<iostream>
#include<cstring>
#include<cstdint>
#include<tensor_ops.cuh>

using std::cout;
using std::endl;

static uint64_t mix(uint64_t seed, uint64_t ctr) {
    uint64_t z = seed + (ctr + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
template<typename T> static T u(uint64_t seed, uint64_t i) { return static_cast<T>(2.0 * (double)(mix(seed, i) >> 11) * 0x1.0p-53 - 1.0); }

template<typename T>
static size_t mismatches(const tensor<T> &a, const tensor<T> &b) {
    if (a.size() != b.size() || a.get_shape() != b.get_shape()) return static_cast<size_t>(-1);
    size_t bad = 0;
    for (size_t i = 0; i < a.size(); i++) bad += std::memcmp(&a.raw()[i], &b.raw()[i], sizeof(T)) != 0;
    return bad;
}

template<typename T>
static void run(const char *name, size_t I, size_t K, size_t J) {
    tensor<T> A(as_shape, {I, K}), B(as_shape, {K, J}), bias(as_shape, {J}), C(as_shape, {I, J}), D(as_shape, {I, J});
    for (size_t i = 0; i < I * K; i++) A[i] = u<T>(1, i);
    for (size_t i = 0; i < K * J; i++) B[i] = u<T>(2, i);
    for (size_t i = 0; i < J; i++) bias[i] = u<T>(3, i) + T(2);
    for (size_t i = 0; i < I * J; i++) { C[i] = u<T>(4, i); D[i] = u<T>(5, i); }
    const device_tensor<T> dA(A), dB(B), dbias(bias), dC(C), dD(D);

    const tensor<T> Z = tensor<T>::matmul(A, B);
    const device_tensor<T> dZ = device_tensor<T>::matmul(dA, dB);
    cout << name << ".matmul=" << mismatches(Z, dZ.to_tensor()) << endl;
    cout << name << ".add_bias=" << mismatches(tensor<T>::add(Z, bias), (dZ + dbias).to_tensor()) << endl;
    cout << name << ".sub=" << mismatches(tensor<T>::sub(C, D), (dC - dD).to_tensor()) << endl;
    cout << name << ".mul=" << mismatches(tensor<T>::mul(C, D), (dC * dD).to_tensor()) << endl;
    cout << name << ".div_bias=" << mismatches(tensor<T>::div(Z, bias), (dZ / dbias).to_tensor()) << endl;
    // the bound of mul on the FIRST operand (kernel.cuh:31): smaller first operand -> only its length is written
    cout << name << ".mul_small_first=" << mismatches(tensor<T>::mul(bias, Z), (dbias * dZ).to_tensor()) << endl;
    cout << name << ".add3=" << mismatches(tensor<T>::add(Z, C, D), device_tensor<T>::add(dZ, dC, dD).to_tensor()) << endl;
    // A^T Z through the transposed view (tensor.cuh:510-592)
    const tensor<T> G = tensor<T>::matmul(tensor<T>::transpose(A), Z);
    const device_tensor<T> dG = device_tensor<T>::matmul(device_tensor<T>::transpose(dA), dZ);
    cout << name << ".matmul_At=" << mismatches(G, dG.to_tensor()) << endl;
    cout << name << ".transpose=" << mismatches(tensor<T>::matmul(tensor<T>::transpose(A), tensor<T>::matmul(A, B)), dG.to_tensor()) << endl;
    // in-place forms
    tensor<T> E = C; E.add(D); device_tensor<T> dE(dC); dE += dD;
    cout << name << ".inplace_add=" << mismatches(E, dE.to_tensor()) << endl;
    // dot: within rounding (the reference combines its blocks with atomics)
    tensor<T> v(as_shape, {I * K}), w(as_shape, {I * K});
    for (size_t i = 0; i < I * K; i++) { v[i] = A.raw()[i]; w[i] = u<T>(6, i); }
    const T ref = tensor<T>::dot(v, w).raw()[0], got = device_tensor<T>::dot(device_tensor<T>(v), device_tensor<T>(w)).to_tensor().raw()[0];
    double scale = 0.0; for (size_t i = 0; i < I * K; i++) scale += std::abs((double)v.raw()[i] * (double)w.raw()[i]);
    cout << name << ".dot_rel=" << std::abs((double)ref - (double)got) / scale << endl;
}

int main(void) {
    run<float>("f32_small", 7, 5, 6);
    run<float>("f32_ragged", 130, 37, 65);
    run<float>("f32_layer", 256, 20, 20);
    run<double>("f64_ragged", 67, 129, 33);
    // row (f).4 through the header: einsum forms = matmul over views, NumPy broadcasting, column normalization
    {
        tensor<float> A(as_shape, {9, 5}), B(as_shape, {5, 4}), col(as_shape, {9, 1}), row(as_shape, {1, 4});
        for (size_t i = 0; i < 45; i++) A[i] = u<float>(11, i);
        for (size_t i = 0; i < 20; i++) B[i] = u<float>(12, i);
        for (size_t i = 0; i < 9; i++) col[i] = u<float>(13, i);
        for (size_t i = 0; i < 4; i++) row[i] = u<float>(14, i);
        const device_tensor<float> dA(A), dB(B), dcol(col), drow(row);
        const tensor<float> Z = tensor<float>::matmul(A, B);
        cout << "f4.einsum_ik=" << mismatches(Z, device_tensor<float>::einsum("ij,jk->ik", dA, dB).to_tensor()) << endl;
        const tensor<float> G = tensor<float>::matmul(tensor<float>::transpose(A), Z);
        cout << "f4.einsum_At=" << mismatches(G, device_tensor<float>::einsum("ji,jk->ik", dA, device_tensor<float>(Z)).to_tensor()) << endl;
        const tensor<float> outer = device_tensor<float>::broadcast_add(dcol, drow).to_tensor();      // (9,1) + (1,4): NOT the reference's flat modulo
        size_t bad = outer.get_shape() != std::vector<size_t>{9, 4};
        for (size_t i = 0; i < 9 && !bad; i++) for (size_t k = 0; k < 4; k++) bad += outer.raw()[i * 4 + k] != col.raw()[i] + row.raw()[k];
        cout << "f4.broadcast_outer=" << bad << endl;
        tensor<float> pts(as_shape, {6, 2});
        for (size_t i = 0; i < 12; i++) pts[i] = u<float>(15, i);
        device_tensor<float> dp(pts); dp.normalize({-1.f, -1.f}, {1.f, 1.f});
        const tensor<float> back = dp.to_tensor();
        size_t nb = 0; for (size_t i = 0; i < 12; i++) nb += back.raw()[i] != 2.f * (pts.raw()[i] + 1.f) / 2.f - 1.f;
        cout << "f4.normalize=" << nb << endl;
        try { (void)device_tensor<float>::einsum("ijk,kl->ijl", dA, dB); cout << "f4.throws=none" << endl; }
        catch (const std::invalid_argument &e) { cout << "f4.throws=" << e.what() << endl; }
    }
    // validation texts and the error channel are the reference's
    const tensor<float> a(as_shape, {2, 3}), b(as_shape, {2, 3});
    try { (void)device_tensor<float>::matmul(device_tensor<float>(a), device_tensor<float>(b)); cout << "throws.matmul=none" << endl; }
    catch (const std::invalid_argument &e) { cout << "throws.matmul=" << e.what() << endl; }
    tensor<float> z = {1.f, 0.f, 2.f}, o = {1.f, 1.f, 1.f};
    try { (void)(device_tensor<float>(o) / device_tensor<float>(z)); cout << "throws.div=none" << endl; }
    catch (const device::runtime::device_exception &e) { cout << "throws.div=" << e.what() << endl; }
    return 0;
}
