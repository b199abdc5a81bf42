"""GPU tests of the B200-native tensor operations (include/pinn_tensor_abi.h) against the REFERENCE
ITSELF: oracle/_ref/ref_tensor_ops is the reference's own tensor<T> (include/tensor.cuh, include/device/kernel.cuh) compiled from
/root/reference by oracle/Makefile and run here on the same seeded inputs.  Bars: matmul, broadcast add and the transposed-view
matmul BIT-EXACT (same per-element operation order); dot 1e-6 relative (the reference combines its blocks with atomics in arrival
order, so its own last bits vary).  The remaining semantics the reference's kernels define — flat-index modulo broadcast, the
mul / div bound on the FIRST operand's size, division by zero — are checked against a numpy restatement of kernel.cuh:10-66."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    from __graft_entry__ import load
    return load()


def _ref_ops(oracle, I, K, J, seed):
    exe = os.path.join(oracle.ref_dir(), "ref_tensor_ops")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_tensor_ops not built")
    out = subprocess.run([exe, str(I), str(K), str(J), str(seed)], capture_output=True, text=True, check=True, timeout=120).stdout
    kv = dict(line.split("=", 1) for line in out.strip().splitlines())
    return {k: np.array(v.split(), dtype=np.float32) for k, v in kv.items() if not k.startswith("shape")}


@pytest.mark.parametrize("I,K,J,seed", [(7, 5, 6, 11), (64, 64, 64, 3), (130, 37, 65, 5), (1, 300, 1, 9), (257, 20, 20, 2),
                                          (1, 1, 1, 4), (3, 4096, 2, 6), (129, 1, 65, 8), (512, 256, 256, 1)])
def test_matmul_add_dot_equal_the_reference_binary(P, oracle, I, K, J, seed):
    ref = _ref_ops(oracle, I, K, J, seed)
    A, B, bias = P.DeviceArray(ref["A"]), P.DeviceArray(ref["B"]), P.DeviceArray(ref["bias"])
    Z = P.tensor_matmul(A, B, I, K, J)
    assert np.array_equal(Z.numpy(), ref["matmul"])                                   # bit-exact: tensor::matmul (tensor.cuh:380-409)
    rc, Zb = P.tensor_elementwise(P.OP_ADD, Z, bias)
    assert rc == 0 and np.array_equal(Zb.numpy(), ref["matmul_bias"])                 # suffix broadcast: kernel.cuh:10-17
    G = P.tensor_matmul(A, Z, K, I, J, a_transposed=True)                             # A^T Z over the transposed VIEW (tensor.cuh:510-592)
    assert np.array_equal(G.numpy(), ref["AtZ"])
    v = P.DeviceArray(ref["A"])                                                       # the probe's dot operands: v = A flattened, w = u01(seed + 3)
    w = P.DeviceArray(np.array([float(oracle.lib().pinn_oracle_u01(seed + 3, i)) for i in range(I * K)], dtype=np.float32))
    d = P.tensor_dot(v, w).numpy()[0]
    # tolerance stated: 1e-6 of the sum of |terms| (FP32 accumulation in a different order; the reference combines its 256-element
    # blocks with atomics in arrival order, so its own last bits vary from run to run)
    wv = np.array([float(oracle.lib().pinn_oracle_u01(seed + 3, i)) for i in range(I * K)], dtype=np.float64)
    scale = float(np.abs(ref["A"].astype(np.float64) * wv).sum())
    assert abs(float(d) - float(ref["dot"][0])) <= 1e-6 * scale + 1e-30
    d2 = P.tensor_dot(v, w).numpy()[0]
    assert d == d2                                                                    # ours is bit-reproducible


def test_elementwise_semantics_of_the_reference_kernels(P):
    rng = np.random.default_rng(0)
    for dt in (np.float32, np.float64):
        a = rng.standard_normal(1000).astype(dt); b = rng.standard_normal(1000).astype(dt)
        small = rng.standard_normal(8).astype(dt) + dt(3)
        A, B, S = P.DeviceArray(a), P.DeviceArray(b), P.DeviceArray(small)
        for op, f in ((P.OP_ADD, np.add), (P.OP_SUB, np.subtract), (P.OP_MUL, np.multiply), (P.OP_DIV, np.divide)):
            rc, R = P.tensor_elementwise(op, A, B)
            assert rc == 0 and np.array_equal(R.numpy(), f(a, b))
        # flat-index modulo broadcast of the smaller SECOND operand (kernel.cuh:16): r[i] = a[i] op s[i % 8]
        idx = np.arange(1000) % 8
        rc, R = P.tensor_elementwise(P.OP_SUB, A, S)
        assert rc == 0 and np.array_equal(R.numpy(), a - small[idx])
        rc, R = P.tensor_elementwise(P.OP_DIV, A, S)
        assert rc == 0 and np.array_equal(R.numpy(), a / small[idx])
        # smaller FIRST operand: add covers the whole result, mul stops at the first operand's size (kernel.cuh:31) ...
        rc, R = P.tensor_elementwise(P.OP_ADD, S, A)
        assert rc == 0 and np.array_equal(R.numpy(), small[idx] + a)
        rc, R = P.tensor_elementwise(P.OP_MUL, S, A)
        want = np.zeros(1000, dtype=dt); want[:8] = small * a[:8]
        assert rc == 0 and np.array_equal(R.numpy(), want)
        # ... unless the caller asks for full-length results
        rc, R = P.tensor_elementwise(P.OP_MUL, S, A, exact_bounds=False)
        assert rc == 0 and np.array_equal(R.numpy(), small[idx] * a)


def test_division_by_zero_is_reported_and_leaves_the_element_unwritten(P):
    a = np.arange(1, 33, dtype=np.float32); b = np.ones(32, dtype=np.float32); b[5] = 0.0; b[17] = 0.0
    rc, R = P.tensor_elementwise(P.OP_DIV, P.DeviceArray(a), P.DeviceArray(b))
    assert rc == 1                                                                    # PINN_ERR_DIVISION_BY_ZERO == constants::error::DIVISION_BY_ZERO
    want = a.copy(); want[5] = 0.0; want[17] = 0.0                                    # result buffers start zero-filled, like the reference's
    assert np.array_equal(R.numpy(), want)


def test_variadic_add_and_double_matmul(P):
    rng = np.random.default_rng(1)
    xs = [rng.standard_normal(5000).astype(np.float32) for _ in range(5)]
    R = P.tensor_add_multiple([P.DeviceArray(x) for x in xs]).numpy()
    want = np.zeros(5000, dtype=np.float32)
    for x in xs:
        want = want + x                                                               # the reference's order (kernel.cuh:61-63), float32 at every step
    assert np.array_equal(R, want)
    I, K, J = 33, 70, 19
    a = rng.standard_normal((I, K)); b = rng.standard_normal((K, J))
    Z = P.tensor_matmul(P.DeviceArray(a), P.DeviceArray(b), I, K, J).numpy().reshape(I, J)
    assert np.allclose(Z, a @ b, rtol=1e-13, atol=1e-13)
    Zt = P.tensor_matmul(P.DeviceArray(a), P.DeviceArray(np.ascontiguousarray(b.T)), I, K, J, b_transposed=True).numpy().reshape(I, J)
    assert np.array_equal(Z, Zt)


def test_large_elementwise_and_dot_sizes(P):
    n = (1 << 22) + 3                                                                 # not a multiple of the vector width
    rng = np.random.default_rng(2)
    a = rng.standard_normal(n).astype(np.float32); b = rng.standard_normal(n).astype(np.float32)
    A, B = P.DeviceArray(a), P.DeviceArray(b)
    rc, R = P.tensor_elementwise(P.OP_ADD, A, B)
    assert rc == 0 and np.array_equal(R.numpy(), a + b)
    d = P.tensor_dot(A, B).numpy()[0]
    assert d == pytest.approx(float(np.dot(a.astype(np.float64), b.astype(np.float64))), rel=1e-5, abs=1e-2)


def test_device_tensor_header_equals_the_reference_tensor_in_one_process(oracle):
    """include/tensor_ops.cuh (device_tensor<T>) and the reference's tensor<T>, same process, same inputs: every elementwise result,
    matmul and transposed-view matmul bit-identical (0 mismatching elements); dot within 1e-6 of the sum of |terms|; the reference's
    exception texts."""
    exe = os.path.join(oracle.ref_dir(), "tensor_ops_test")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/tensor_ops_test not built")
    out = subprocess.run([exe], capture_output=True, text=True, check=True, timeout=300).stdout
    kv = dict(line.split("=", 1) for line in out.strip().splitlines())
    cases = ["f32_small", "f32_ragged", "f32_layer", "f64_ragged"]
    for c in cases:
        for op in ("matmul", "add_bias", "sub", "mul", "div_bias", "mul_small_first", "add3", "matmul_At", "transpose", "inplace_add"):
            assert kv[f"{c}.{op}"] == "0", (c, op, kv[f"{c}.{op}"])
        assert float(kv[f"{c}.dot_rel"]) < 1e-6
    for k in ("f4.einsum_ik", "f4.einsum_At", "f4.broadcast_outer", "f4.normalize"):
        assert kv[k] == "0", (k, kv[k])
    assert kv["f4.throws"].startswith("tensor::einsum:")
    assert kv["throws.matmul"] == "tensor::matmul: invalid shapes"                      # tensor.cuh:354
    assert kv["throws.div"] == "tensor::div: division by zero"                         # runtime.cuh:17-38


# ---- row (f).4: NumPy-rule broadcasting, the einsum forms of a dense layer, column normalization (semantics: numpy's; the reference
# ---- defines none — einsum is an empty stub, tensor.cuh:467-469) -------------------------------------------------------------
@pytest.mark.parametrize("sa,sb", [((3, 1), (1, 4)), ((5, 4, 3), (3,)), ((2, 1, 6), (7, 1)), ((1,), (4, 5)), ((4, 5), (4, 5)), ((3, 1, 1, 2), (1, 4, 5, 1))])
def test_general_broadcasting_matches_numpy(P, sa, sb):
    rng = np.random.default_rng(3)
    for dt in (np.float32, np.float64):
        a = rng.standard_normal(sa).astype(dt); b = (rng.standard_normal(sb) + 3).astype(dt)
        for op, f in ((P.OP_ADD, np.add), (P.OP_SUB, np.subtract), (P.OP_MUL, np.multiply), (P.OP_DIV, np.divide)):
            R, shape = P.tensor_broadcast(op, P.DeviceArray(a), sa, P.DeviceArray(b), sb)
            want = f(a, b)
            assert shape == want.shape and np.array_equal(R.numpy().reshape(shape), want)          # bit-exact: one IEEE operation per element


def test_einsum_forms_of_a_dense_layer_match_numpy_and_matmul(P):
    rng = np.random.default_rng(4)
    I, K, J = 33, 20, 17
    a = rng.standard_normal((I, K)).astype(np.float32); b = rng.standard_normal((K, J)).astype(np.float32)
    v = rng.standard_normal(K).astype(np.float32); u = rng.standard_normal(I).astype(np.float32)
    A, B, At, Bt, V, U = (P.DeviceArray(x) for x in (a, b, np.ascontiguousarray(a.T), np.ascontiguousarray(b.T), v, u))
    Z = P.tensor_matmul(A, B, I, K, J).numpy()
    for spec, x, sx, y, sy in (("ij,jk->ik", A, (I, K), B, (K, J)), ("ji,jk->ik", At, (K, I), B, (K, J)),
                               ("ij,kj->ik", A, (I, K), Bt, (J, K)), ("ab,cb->ac", A, (I, K), Bt, (J, K)), ("ji,kj->ik", At, (K, I), Bt, (J, K))):
        rc, R = P.tensor_einsum(spec, x, sx, y, sy, I * J)
        assert rc == 0 and np.array_equal(R.numpy(), Z), spec                                    # same kernel, same order: bit-identical to matmul
    rc, R = P.tensor_einsum("ij,j->i", A, (I, K), V, (K,), I)
    assert rc == 0 and np.allclose(R.numpy(), np.einsum("ij,j->i", a.astype(np.float64), v.astype(np.float64)), rtol=1e-5, atol=1e-5)
    rc, R = P.tensor_einsum("i,ij->j", U, (I,), A, (I, K), K)
    assert rc == 0 and np.allclose(R.numpy(), np.einsum("i,ij->j", u.astype(np.float64), a.astype(np.float64)), rtol=1e-5, atol=1e-5)
    rc, R = P.tensor_einsum("i,i->", V, (K,), V, (K,), 1)
    assert rc == 0 and R.numpy()[0] == pytest.approx(float(np.dot(v.astype(np.float64), v.astype(np.float64))), rel=1e-6)
    rc, R = P.tensor_einsum("i,j->ij", U, (I,), V, (K,), I * K)
    assert rc == 0 and np.array_equal(R.numpy().reshape(I, K), np.outer(u, v))
    for bad in ("ij,jk->ki", "ii,ij->j", "ijk,kl->ijl", "ij,jk"):
        rc, _ = P.tensor_einsum(bad, A, (I, K), B, (K, J), I * J)
        assert rc in (3, 7), bad                                                                 # PINN_ERR_INVALID_ARGUMENT / PINN_ERR_UNSUPPORTED
    rc, _ = P.tensor_einsum("ij,jk->ik", A, (I, K), A, (I, K), I * K)
    assert rc == 3                                                                               # contracted sizes differ


def test_column_normalization_matches_the_host_helper_arithmetic(P):
    rng = np.random.default_rng(5)
    n, d = 100003, 3
    x = (rng.random((n, d)) * np.array([2.0, 1.0, 6.28]) + np.array([-1.0, 0.0, 0.0])).astype(np.float32)
    lo = np.array([-1.0, 0.0, 0.0], dtype=np.float32); hi = np.array([1.0, 1.0, 6.28], dtype=np.float32)
    X = P.DeviceArray(x)
    P.tensor_normalize_columns(X, n, d, lo, hi)
    want = (np.float32(2.0) * (x - lo) / (hi - lo) - np.float32(1.0)).astype(np.float32)          # utils::normalize (src/utils.cu), float32 at every step
    assert np.array_equal(X.numpy().reshape(n, d), want)
    X = P.DeviceArray(x)
    glo, ghi = P.tensor_normalize_columns(X, n, d)                                               # bounds from the data itself
    assert np.array_equal(glo, x.min(axis=0)) and np.array_equal(ghi, x.max(axis=0))
    y = X.numpy().reshape(n, d)
    assert y.min() == -1.0 and y.max() == 1.0


def test_scalar_and_single_element_operands(P):
    """Edge cases of the reference's flat-index broadcast: a one-element operand on either side (kernel.cuh:16), one-element dot."""
    a = np.arange(1, 10, dtype=np.float32); one = np.array([4.0], dtype=np.float32)
    A, O = P.DeviceArray(a), P.DeviceArray(one)
    rc, R = P.tensor_elementwise(P.OP_ADD, A, O); assert rc == 0 and np.array_equal(R.numpy(), a + 4.0)
    rc, R = P.tensor_elementwise(P.OP_SUB, O, A); assert rc == 0 and np.array_equal(R.numpy(), 4.0 - a)
    rc, R = P.tensor_elementwise(P.OP_DIV, A, O); assert rc == 0 and np.array_equal(R.numpy(), a / np.float32(4.0))
    rc, R = P.tensor_elementwise(P.OP_MUL, O, A)                                     # first operand smaller: only its one element is written
    want = np.zeros(9, dtype=np.float32); want[0] = 4.0
    assert rc == 0 and np.array_equal(R.numpy(), want)
    assert P.tensor_dot(O, O).numpy()[0] == 16.0
    rc, R = P.tensor_elementwise(P.OP_DIV, A, P.DeviceArray(np.zeros(1, dtype=np.float32)))
    assert rc == 1 and np.array_equal(R.numpy(), np.zeros(9, dtype=np.float32))      # every denominator zero: nothing written, error reported
