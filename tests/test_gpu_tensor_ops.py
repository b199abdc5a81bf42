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


@pytest.mark.parametrize("I,K,J,seed", [(7, 5, 6, 11), (64, 64, 64, 3), (130, 37, 65, 5), (1, 300, 1, 9), (257, 20, 20, 2)])
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
    assert d == pytest.approx(float(ref["dot"][0]), rel=1e-6, abs=1e-5)               # tolerance stated: 1e-6 relative (+1e-5 absolute near zero) (atomics order in the reference)
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
    assert kv["throws.matmul"] == "tensor::matmul: invalid shapes"                      # tensor.cuh:354
    assert kv["throws.div"] == "tensor::div: division by zero"                         # runtime.cuh:17-38
