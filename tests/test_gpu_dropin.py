"""GPU tests of the drop-in boundary against the REFERENCE itself (binaries under oracle/_ref are
built from /root/reference's own sources by oracle/Makefile; they travel to the GPU box prebuilt).

 * G0: the literal reference program (src/main.cu, unmodified) prints the expected line.
 * reference tensor<T> ops (matmul / suffix-broadcast add / dot / transpose-matmul) pin the layout
   conventions the oracle and the CUDA path follow (kernel.cuh:68-83, 10-17, 85-99).
 * include/network.cuh + src/network.cu + src/utils.cu compiled against the reference's unmodified
   tensor.cuh give the same numbers as the ctypes path and the oracle.
"""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ref(oracle, name):
    exe = os.path.join(oracle.ref_dir(), name)
    if not os.path.exists(exe):
        pytest.skip(f"oracle/_ref/{name} not built")
    return exe


def test_reference_main_prints_g0(oracle):
    out = subprocess.run([_ref(oracle, "ref_main")], capture_output=True, text=True, check=True, timeout=120).stdout
    assert "result = [0.2, 0.333333, 0.428571, 0.5]" in out.splitlines()[1]
    assert out.startswith("\n") and out.endswith("\n\n")


def test_reference_tensor_ops_pin_the_layout_conventions(oracle):
    I, K, J, seed = 7, 5, 6, 11
    out = subprocess.run([_ref(oracle, "ref_tensor_ops"), str(I), str(K), str(J), str(seed)],
                         capture_output=True, text=True, check=True, timeout=120).stdout
    kv = dict(line.split("=", 1) for line in out.strip().splitlines())
    f = lambda k: np.array(kv[k].split(), dtype=np.float32)
    A, B, bias = f("A").reshape(I, K), f("B").reshape(K, J), f("bias")
    # inputs come from the same counter-based generator the sampler uses
    assert A[0, 0] == np.float32(2.0) * oracle.lib().pinn_oracle_u01(seed, 0) - np.float32(1.0)
    Z = A.astype(np.float64) @ B.astype(np.float64)
    assert np.allclose(f("matmul").reshape(I, J), Z, rtol=1e-5, atol=1e-6)          # row-major [I,K]x[K,J]
    assert np.allclose(f("matmul_bias").reshape(I, J), Z + bias, rtol=1e-5, atol=1e-6)  # suffix-broadcast bias
    assert kv["shape_Z"] == f"({I}, {J})" and kv["shape_AtZ"] == f"({K}, {J})"
    assert np.allclose(f("AtZ").reshape(K, J), A.T.astype(np.float64) @ f("matmul").reshape(I, J), rtol=1e-4, atol=1e-5)  # dW = A^T Zbar


def test_network_header_against_reference_tensor(oracle):
    from __graft_entry__ import load
    P = load()
    out = subprocess.run([_ref(oracle, "network_dropin_test")], capture_output=True, text=True, check=True, timeout=300).stdout
    kv = {}
    for line in out.strip().splitlines():
        k, v = line.split("=", 1)
        kv[k] = v
    assert kv["g0"] == "[0.2, 0.333333, 0.428571, 0.5]"
    assert kv["params"] == "3021 layers=9 channels=4"
    assert kv["w0.shape"] == "(2, 20) b8.shape=(1)"
    # same numbers as the ctypes path on the same configuration (same library, same kernels)
    cfg = P.default_config(1)
    with P.Network(cfg) as net:
        w = net.get_params()
        l0 = net.loss_grad()
        g = net.get_grads()
        steps = [net.train_step()["loss"] for _ in range(3)]
        w_after = net.get_params()
        x = np.array([[a, b] for a in (-1.0, 0.0, 1.0) for b in (0.0, 0.5, 1.0)], dtype=np.float32)
        y = net.forward(x)
        xb = net.get_points(1)
    assert float(kv["w0[0]"]) == pytest.approx(float(w[0]), rel=1e-7)
    assert float(kv["loss0"].split()[0]) == pytest.approx(l0["loss"], rel=1e-6)
    assert float(kv["grad_sq"]) == pytest.approx(float((g.astype(np.float64) ** 2).sum()), rel=1e-5)
    for i in range(3):
        assert float(kv[f"step{i + 1}"]) == pytest.approx(steps[i], rel=1e-6)
    assert float(kv["w0[0]_after"]) == pytest.approx(float(w_after[0]), rel=1e-6)
    ys = kv["y.shape"].split(" y=")[1].replace("[", "").replace("]", "").split(",")
    assert kv["y.shape"].startswith("(9, 1)")
    assert np.allclose(np.array(ys, dtype=np.float32), y[:, 0], rtol=2e-5, atol=1e-6)
    assert kv["r.shape"] == "(9, 1)"
    assert kv["xb.shape"].startswith("(256, 2)")
    assert [float(t) for t in kv["xb.shape"].split("xb0=")[1].split(",")] == pytest.approx([float(xb[0, 0]), float(xb[0, 1])], rel=1e-6)
    assert kv["shard"] == "9600,12800"
    # layer interface, device-resident points, device error channel, rank queries (include/network.cuh)
    assert kv["layer0"] == "2x20 tanh size=60 layer8=20x1 identity output=1"
    assert float(kv["layer0.w[0]"]) == float(kv["w0[0]_after"])
    d = dict(t.split("=") for t in ("dev.step1=" + kv["dev.step1"]).split())
    assert float(d["dev.step1"]) == pytest.approx(steps[0], rel=1e-6) and float(d["dev.step2"]) == pytest.approx(steps[1], rel=1e-6)
    assert int(d["dev.h2d_bytes"]) == 0                               # trained from a device pointer: nothing crossed PCIe
    assert int(kv["dev.h2d_bytes_after_tensor"]) == 4096 * 2 * 4      # reference tensor<T>: device half private -> host upload
    assert "not a device pointer" in kv["e.hostptr"]
    assert kv["e.overflow"] == "network::train_step: overflow code=2"
    assert kv["rank"] == "0/1"
    assert kv["e.assign"] == "network::assign: size mismatch - 7 != 3021"
    assert "pinn_create: width must be a multiple of 4" in kv["e.width"]
    # and the oracle agrees with what the C++ path printed
    ocfg = oracle.baseline_config(1)
    _, s_ref = oracle.loss_grad(ocfg, w.astype(np.float64), *[oracle.sample(ocfg, k) for k in range(3)], want_grad=False)
    assert float(kv["loss0"].split()[0]) == pytest.approx(oracle.loss_from_sums(ocfg, s_ref)[0], rel=1e-5)


def test_reference_composed_step_matches_oracle_loss(oracle):
    """oracle/_ref/ref_step: the reference's own tensor<float> ops (+ labelled host shim for tanh/sqrt) composing the
    interior part of the train step.  Its loss must equal the oracle's residual MSE on the same points and weights —
    this pins the oracle's math on the reference's matmul / broadcast / dot / transpose semantics end to end."""
    out = subprocess.run([_ref(oracle, "ref_step"), "512", "20", "8", "1"], capture_output=True, text=True, check=True, timeout=600).stdout
    kv = dict(line.split("=", 1) for line in out.strip().splitlines() if "=" in line and not line.startswith("steps"))
    ocfg = oracle.make_config(oracle.PDE_BURGERS, 20, 8, 512, 1, 1)
    xf = oracle.sample(ocfg, 0)
    w = oracle.init_params(ocfg).astype(np.float64)
    _, s = oracle.loss_grad(ocfg, w, xf, np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32), want_grad=False)
    assert float(kv["loss0"]) == pytest.approx(s[0] / 512, rel=2e-5)
    # one Adam step on the same gradient lowers the interior loss the same way the oracle does
    g, _ = oracle.loss_grad(ocfg, w, xf, np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32))
    # the oracle's gradient includes the 2/N seed with N = n_interior, like ref_step
    m = np.zeros_like(w); v = np.zeros_like(w); w1 = w.copy()
    oracle.adam(ocfg, w1, m, v, g, 1)
    _, s1 = oracle.loss_grad(ocfg, w1, xf, np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32), want_grad=False)
    assert float(kv["loss_last"]) == pytest.approx(s1[0] / 512, rel=1e-3)
    assert float(kv["points_per_second"]) > 0
