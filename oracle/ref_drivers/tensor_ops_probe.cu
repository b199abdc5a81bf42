// tensor_ops_probe.cu — drives the REFERENCE's own tensor<T> device ops (include/tensor.cuh,
// include/device/kernel.cuh, compiled from /root/reference where they lie) on seeded inputs
// and prints results, so the layout conventions the oracle follows (row-major matmul
// kernel.cuh:68-83, suffix-broadcast add kernel.cuh:10-17, dot kernel.cuh:85-99) are pinned
// against the reference itself on a GPU.  Test infrastructure only (oracle/_ref).
//   usage: ref_tensor_ops I K J seed   → prints A, B, bias, matmul(A,B), matmul(A,B)+bias, dot
#include<tensor.cuh>
#include<cstdint>
#include<cstdlib>
#include<iostream>
#include<iomanip>
#include<vector>

static uint64_t mix(uint64_t seed, uint64_t ctr) {
    uint64_t z = seed + (ctr + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static float u01(uint64_t seed, uint64_t ctr) { return (float)(mix(seed, ctr) >> 40) * 0x1.0p-24f; }

static void dump(const char* name, const tensor<float>& t) {
    std::cout << name << "=";
    for (size_t i = 0; i < t.size(); i++) std::cout << std::setprecision(9) << t.raw()[i] << (i + 1 == t.size() ? "" : " ");
    std::cout << "\n";
}

int main(int argc, char** argv) {
    size_t I = argc > 1 ? std::strtoul(argv[1], nullptr, 10) : 5;
    size_t K = argc > 2 ? std::strtoul(argv[2], nullptr, 10) : 3;
    size_t J = argc > 3 ? std::strtoul(argv[3], nullptr, 10) : 4;
    uint64_t seed = argc > 4 ? std::strtoull(argv[4], nullptr, 10) : 7;
    tensor<float> A(as_shape, {I, K}), B(as_shape, {K, J}), bias(as_shape, {J}), v(as_shape, {I * K}), w(as_shape, {I * K});
    for (size_t i = 0; i < I * K; i++) { A[i] = 2.0f * u01(seed, i) - 1.0f; v[i] = A.raw()[i]; w[i] = u01(seed + 3, i); }
    for (size_t i = 0; i < K * J; i++) B[i] = 2.0f * u01(seed + 1, i) - 1.0f;
    for (size_t i = 0; i < J; i++) bias[i] = u01(seed + 2, i);
    tensor<float> Z = tensor<float>::matmul(A, B);
    tensor<float> Zb = tensor<float>::add(Z, bias);
    tensor<float> d = tensor<float>::dot(v, w);
    tensor<float> At = tensor<float>::transpose(A);
    tensor<float> G = tensor<float>::matmul(At, Z);   // A^T Z : the dW = A^T Zbar product of a reference-style backward
    dump("A", A); dump("B", B); dump("bias", bias); dump("matmul", Z); dump("matmul_bias", Zb);
    dump("dot", d); dump("AtZ", G);
    std::cout << "shape_Z=" << vec_to_str(Z.get_shape()) << "\n";
    std::cout << "shape_AtZ=" << vec_to_str(G.get_shape()) << "\n";
    return 0;
}
