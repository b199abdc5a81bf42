/* pinn_tensor_abi.h — C ABI of the B200-native tensor operations: the device arithmetic of
 * the reference's tensor<T> — add / sub / mul / div with its flat-index broadcast, row-major matmul, dot, variadic add —
 * on DEVICE-RESIDENT buffers, stream-ordered, with no host mirror and no copy a caller did not ask for.
 *
 * The reference (header-only C++, no FFI) runs every one of these as: allocate the result on host and device, memset,
 * upload, launch a scalar kernel with 64-bit div/mod per access, blocking D2H of the whole result, free
 * (include/tensor.cuh:84-140, 347-465).  These entry points are what a tensor<T> whose device half is authoritative
 * (the dormant `device_dirty`, tensor.cuh:22) would call instead; include/tensor_ops.cuh is that type.
 *
 * Conventions (same as include/pinn_abi.h): plain pointers and sizes, every function returns a pinn_status, nothing throws.
 * All pointers are DEVICE pointers; `stream` is a cudaStream_t passed as void* (NULL = the default stream); calls are
 * asynchronous on that stream unless stated.  dtype: 0 = float, 1 = double (the reference's demo type).
 *
 * Results are those of the reference on the same inputs:
 *   - elementwise and matmul: bit-identical (same operation order per element: the matmul keeps one accumulator per output and
 *     walks k upwards with fused multiply-adds, as nvcc compiles kernel.cuh:77-80);
 *   - mul / div keep the reference's bound (kernel.cuh:31,40: the guard is the FIRST operand's size, so a smaller first operand
 *     leaves the tail of the result at its initial zero); pass exact_bounds = 0 for NumPy-like full-length results instead;
 *   - div by zero leaves the element unwritten and reports PINN_ERR_DIVISION_BY_ZERO (kernel.cuh:43-47, runtime.cuh:40-47);
 *   - dot: the reference sums 256-element blocks serially and combines them with atomicAdd in arrival order
 *     (kernel.cuh:85-99: not reproducible run to run); here a fixed-order tree — equal to the reference within rounding
 *     (tests: 1e-6 relative), bit-reproducible.
 */
#ifndef PINN_TENSOR_ABI_H
#define PINN_TENSOR_ABI_H
#include <stddef.h>
#include "pinn_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef enum pinn_tensor_op { PINN_OP_ADD = 0, PINN_OP_SUB = 1, PINN_OP_MUL = 2, PINN_OP_DIV = 3 } pinn_tensor_op;

/* r[i] = a[i % na] (op) b[i % nb] for i < bound; replaces tensor::add/sub/mul/div (tensor.cuh:315-345) and
 * kernel::add/sub/mul/div (kernel.cuh:10-48).  nr must be max(na, nb) (tensor.cuh:95: the result takes the bigger operand's
 * shape).  bound = nr for add / sub; for mul / div it is na when exact_bounds != 0 (the reference), else nr.
 * r must not alias a or b unless the sizes are equal (the in-place forms of the reference, tensor.cuh:315-329).
 * For PINN_OP_DIV the call SYNCHRONISES the stream and returns PINN_ERR_DIVISION_BY_ZERO if any denominator was zero. */
int pinn_tensor_elementwise(int op, int dtype, const void* a, size_t na, const void* b, size_t nb, void* r, size_t nr,
                            int exact_bounds, void* stream);

/* r[I, J] = A[I, K] . B[K, J], row-major (tensor::matmul tensor.cuh:347-409, kernel::matmul kernel.cuh:68-83).
 * a_transposed / b_transposed: the operand is stored as its transpose ([K, I] / [J, K]) — the reference's transpose is
 * metadata over the same buffer (tensor.cuh:510-592, variables.cuh:15-41), so a transposed view costs no copy here either. */
int pinn_tensor_matmul(int dtype, const void* a, int a_transposed, const void* b, int b_transposed, void* r,
                       size_t I, size_t K, size_t J, void* stream);

/* r[0] = sum_i a[i] b[i] (tensor::dot tensor.cuh:411-465, kernel::dot kernel.cuh:85-99).  scratch: device buffer of
 * pinn_tensor_dot_scratch(n) elements of the dtype, or NULL to let the call allocate and free one (stream-ordered). */
size_t pinn_tensor_dot_scratch(size_t n);
int pinn_tensor_dot(int dtype, const void* a, const void* b, size_t n, void* r, void* scratch, void* stream);

/* r[i] = srcs[0][i] + srcs[1][i] + ... in that order (variadic tensor::add tensor.cuh:471-508, kernel::add_multiple
 * kernel.cuh:55-66).  srcs: HOST array of `count` device pointers (count <= 16), each of n elements. */
int pinn_tensor_add_multiple(int dtype, const void* const* srcs, size_t count, size_t n, void* r, void* stream);

/* ---- row (f).4: what the reference names but does not define (README.md:54, tensor.cuh:292-313, 467-469) --------------------
 * The reference gives these no behaviour to be identical to (einsum is an empty stub, broadcast_order is suffix-only, the
 * normalization of README.md:54 was never written); the semantics here are NumPy's, stated per entry and tested against numpy. */

/* General broadcasting: r = a (op) b with NumPy rules (shapes aligned at the trailing dimension, a dimension of size 1
 * stretches).  Shapes are HOST arrays; r_shape must be the broadcast of the two (at most 8 dimensions).  Division by zero
 * follows IEEE (inf / nan): the error channel of pinn_tensor_elementwise belongs to the reference's flat-index semantics. */
int pinn_tensor_broadcast(int op, int dtype, const void* a, const size_t* a_shape, int a_dim, const void* b, const size_t* b_shape, int b_dim,
                          void* r, const size_t* r_shape, int r_dim, void* stream);

/* Two-operand einsum for the forms that are ONE contraction over at most one index of vectors / matrices — exactly the products a
 * dense layer and its backward pass need: "ij,jk->ik" "ji,jk->ik" "ij,kj->ik" "ji,kj->ik" (matmul over transposed views),
 * "ij,j->i" "i,ij->j" (matrix-vector), "i,i->" (dot), "i,j->ij" (outer product); any index letters.  Shapes as for numpy.einsum;
 * r has the output's size.  Anything else: PINN_ERR_UNSUPPORTED.  Same kernels and the same per-element order as
 * pinn_tensor_matmul / pinn_tensor_dot. */
int pinn_tensor_einsum(int dtype, const char* spec, const void* a, const size_t* a_shape, int a_dim, const void* b, const size_t* b_shape, int b_dim,
                       void* r, void* stream);

/* Column normalization of a point set x[n, d] in place (README.md:54 "data handling and normalization"): x <- 2 (x - lo) / (hi - lo) - 1
 * per column, the arithmetic of utils::normalize (src/utils.cu) bit for bit.  lo / hi: HOST arrays of d floats, or both NULL to
 * use the columns' own minimum / maximum (computed on the device, fixed-order reduction); stats_out (HOST, 2 d floats, may be
 * NULL) receives the bounds used.  Synchronises when the bounds are computed or returned. */
int pinn_tensor_normalize_columns(float* x, size_t n, size_t d, const float* lo, const float* hi, float* stats_out, void* stream);

/* Stream-ordered device memory for callers without a CUDA toolchain (ctypes, the tests): cudaMallocAsync / cudaFreeAsync /
 * cudaMemcpyAsync + synchronise.  The reference allocates three buffers per tensor with synchronous cudaMalloc
 * (interface.cuh:86-115). */
int pinn_device_alloc(void** ptr, size_t bytes, void* stream);
int pinn_device_free(void* ptr, void* stream);
int pinn_device_upload(void* dst, const void* host_src, size_t bytes, void* stream);     /* synchronises */
int pinn_device_download(void* host_dst, const void* src, size_t bytes, void* stream);   /* synchronises */
int pinn_device_copy(void* dst, const void* src, size_t bytes, void* stream);            /* device to device, asynchronous */

#ifdef __cplusplus
}
#endif
#endif
