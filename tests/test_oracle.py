"""CPU tests of the oracle (no GPU): pin it on everything available.

 * G1 — the reference's own host tensor storage (oracle/_ref/ref_host_probe, built from
   /root/reference/include/host/init_tensor.h): the only reference-produced vectors runnable on CPU.
 * tests/golden/*.npz — independent torch.autograd float64 implementation (make_golden.py).
 * central finite differences of the oracle's own loss.
 The train-step oracle itself is PARITY UNPINNED against the reference (it has no train step).
"""
import os
import subprocess

import numpy as np
import pytest

from helpers import GOLDEN_CASES, cfg_from_dict, load_golden, rel_err

G1 = {  # SURVEY.md section 4, golden vectors G1 (observed from the reference's init_tensor.h)
    "v1": "[1, 2, 3, 4]", "v1.shape": "(4)", "v1.stride": "(1)",
    "m": "[[1, 2, 3], [4, 5, 6]]", "m.shape": "(2, 3)", "m.stride": "(3, 1)", "c.stride": "(4, 2, 1)",
    "s": "3.5", "s.shape": "(scalar)", "z": "[[0, 0, 0], [0, 0, 0]]", "r": "[[1, 2], [3, 4], [5, 6]]",
    "f": "[0.1, 0.333333, 1e-07, 1.23457e+08]", "m(1,2)": "6",
    "e.at": "tensor::calculate_stride: out of bounds access",
    "e.reshape": "init_tensor::reshape: cannot reshape, total size doesn't match",
    "e.ragged": "init_tensor::init_tensor: inconsistent shape",
    "e.assign": "init_tensor::assign: shape mismatch - (2, 3) != (2, 2)",
    "e.cast": "init_tensor::operator T: invalid cast",
}


def test_reference_host_probe_g1(oracle):
    exe = os.path.join(oracle.ref_dir(), "ref_host_probe")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref not built (reference checkout absent)")
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    got = dict(line.split("=", 1) for line in out.strip().splitlines())
    assert got == G1


def test_sampler_bit_exact_vs_independent(oracle, golden_dir):
    z = np.load(os.path.join(golden_dir, "sampler.npz"))
    mixes = np.array([oracle.lib().pinn_oracle_mix(1, c) for c in range(16)], dtype=np.uint64)
    assert np.array_equal(mixes, z["mix_seed1_ctr0_15"])
    for pde in range(5):
        cfg = oracle.make_config(pde, 20, 8, 4096)
        for which in range(3):
            x = oracle.sample(cfg, which, 1000, 1048)
            assert np.array_equal(x.view(np.uint32), z[f"pde{pde}_set{which}"]), (pde, which)
    w = oracle.init_params(oracle.baseline_config(1))
    assert np.array_equal(w[:512].view(np.uint32), z["burgers_20x8_params_head"])


def test_sampler_ranges_and_faces(oracle):
    cfg = oracle.baseline_config(5, n_interior=4096)
    x = oracle.sample(cfg, 0)
    assert x[:, :2].min() >= 0 and x[:, :2].max() <= np.float32(2 * np.pi) and x[:, 2].min() >= 0 and x[:, 2].max() < 1
    xb = oracle.sample(cfg, 1)
    lo, hi = np.float32(0), np.float32(6.28318530717958647692)
    for i in range(16):
        axis, side = (i % 4) // 2, i % 2
        assert xb[i, axis] == (hi if side else lo)
    x0 = oracle.sample(cfg, 2)
    assert np.all(x0[:, 2] == 0)


@pytest.mark.parametrize("R", [1, 2, 3, 4, 8])
def test_shards_tile_the_global_set_bit_exact(oracle, R):
    cfg = oracle.make_config(oracle.PDE_ALLEN_CAHN, 16, 2, 1003, 101, 77)
    for which in range(3):
        n = oracle.set_count(cfg, which)
        whole = oracle.sample(cfg, which)
        parts, prev_e = [], 0
        for r in range(R):
            b, e = oracle.shard_range(n, r, R)
            assert b == prev_e and b == (r * n) // R and e == ((r + 1) * n) // R
            prev_e = e
            parts.append(oracle.sample(cfg, which, b, e))
        assert prev_e == n
        assert np.array_equal(np.concatenate(parts).view(np.uint32), whole.view(np.uint32))


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_torch_autograd(oracle, name):
    d, z = load_golden(name)
    cfg = cfg_from_dict(oracle, d)
    # the oracle's own sampler reproduces the fixture's points bit for bit
    for which, key in ((0, "xf"), (1, "xb"), (2, "x0")):
        assert np.array_equal(oracle.sample(cfg, which).view(np.uint32), z[key].view(np.uint32))
    p = z["params"].astype(np.float64)
    g, s = oracle.loss_grad(cfg, p, z["xf"], z["xb"], z["x0"])
    assert rel_err(s, z["sums"]) < 1e-12
    assert abs(oracle.loss_from_sums(cfg, s)[0] - float(z["loss"])) < 1e-12 * abs(float(z["loss"]))
    assert rel_err(g, z["grad"]) < 1e-11
    Y, r = oracle.residual(cfg, p, z["xf"])
    assert rel_err(Y, z["Y"]) < 1e-12 and rel_err(r, z["r"]) < 1e-11
    if len(z["xb"]):
        assert rel_err(oracle.target(cfg, 1, z["xb"]), z["target_b"]) < 1e-14 or np.abs(z["target_b"]).max() == 0
    if len(z["x0"]):
        assert rel_err(oracle.target(cfg, 2, z["x0"]), z["target_0"]) < 1e-14
    # float32 build of the same code stays within FP32 round-off of the float64 one
    g32, s32 = oracle.loss_grad(cfg, z["params"], z["xf"], z["xb"], z["x0"], dtype=np.float32)
    assert rel_err(s32, z["sums"]) < 2e-5 and rel_err(g32, z["grad"]) < 2e-4


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_adam_trajectory(oracle, name):
    d, z = load_golden(name)
    cfg = cfg_from_dict(oracle, d)
    steps = len(z["traj_loss"])
    p, m, v, losses, _ = oracle.train(cfg, z["params"].astype(np.float64), steps, dtype=np.float64)
    assert rel_err(losses[:, 0], z["traj_loss"]) < 1e-11
    assert rel_err(p, z["params_after"]) < 1e-11
    assert rel_err(m, z["adam_m"]) < 1e-10 and rel_err(v, z["adam_v"]) < 1e-10


def test_oracle_gradient_vs_finite_differences(oracle):
    cfg = oracle.make_config(oracle.PDE_NAVIER_STOKES, 8, 2, 24, 8, 8)
    xf, xb, x0 = (oracle.sample(cfg, k) for k in range(3))
    rng = np.random.default_rng(0)
    p = (oracle.init_params(cfg) + 0.05 * rng.standard_normal(oracle.num_params(cfg))).astype(np.float64)
    g, _ = oracle.loss_grad(cfg, p, xf, xb, x0)

    def loss(q):
        _, s = oracle.loss_grad(cfg, q, xf, xb, x0, want_grad=False)
        return oracle.loss_from_sums(cfg, s)[0]

    idx = rng.choice(p.size, 24, replace=False)
    for i in idx:
        h = 1e-6
        e = np.zeros_like(p); e[i] = h
        fd = (loss(p + e) - loss(p - e)) / (2 * h)
        assert abs(fd - g[i]) < 1e-6 * max(1.0, abs(g[i])), (i, fd, g[i])


def test_oracle_thread_count_independent_to_roundoff(oracle):
    cfg = oracle.baseline_config(1)
    xf, xb, x0 = (oracle.sample(cfg, k) for k in range(3))
    p = oracle.init_params(cfg).astype(np.float64)
    nt = oracle.threads()
    g_all, s_all = oracle.loss_grad(cfg, p, xf, xb, x0)
    oracle.lib().pinn_oracle_set_threads(1)
    g_1, s_1 = oracle.loss_grad(cfg, p, xf, xb, x0)
    oracle.lib().pinn_oracle_set_threads(nt)
    assert rel_err(g_all, g_1) < 1e-12 and rel_err(s_all, s_1) < 1e-13


def test_empty_and_tiny_sets(oracle):
    cfg = oracle.make_config(oracle.PDE_HELMHOLTZ, 8, 2, 1, 1, 0)
    xf, xb, x0 = (oracle.sample(cfg, k) for k in range(3))
    assert xf.shape == (1, 2) and x0.shape == (0, 2)
    g, s = oracle.loss_grad(cfg, oracle.init_params(cfg).astype(np.float64), xf, xb, x0)
    assert np.isfinite(g).all() and s[2] == 0


@pytest.mark.parametrize("R", [2, 3])
def test_shard_gradients_add_up_to_the_global_gradient(oracle, R):
    """The data-parallel contract: seeds are scaled by 1/N_global, so the per-shard gradients and loss sums of
    any partition ADD UP to the single-rank result."""
    cfg = oracle.make_config(oracle.PDE_ALLEN_CAHN, 12, 3, 1001, 97, 103)
    xs = [oracle.sample(cfg, k) for k in range(3)]
    rng = np.random.default_rng(2)
    p = (oracle.init_params(cfg) + 0.05 * rng.standard_normal(oracle.num_params(cfg))).astype(np.float64)
    g_all, s_all = oracle.loss_grad(cfg, p, *xs)
    g_sum, s_sum = np.zeros_like(g_all), np.zeros(3)
    for r in range(R):
        parts = []
        for k in range(3):
            b, e = oracle.shard_range(len(xs[k]), r, R)
            parts.append(xs[k][b:e])
        g, s = oracle.loss_grad(cfg, p, *parts)
        g_sum += g; s_sum += s
    assert rel_err(g_sum, g_all) < 1e-12 and rel_err(s_sum, s_all) < 1e-12


def test_gradient_is_linear_in_the_loss_weights(oracle):
    cfg = oracle.make_config(oracle.PDE_BURGERS, 20, 4, 512, 64, 64)
    xs = [oracle.sample(cfg, k) for k in range(3)]
    p = oracle.init_params(cfg).astype(np.float64)
    parts = []
    for lam in ((1, 0, 0), (0, 1, 0), (0, 0, 1)):
        c = cfg.copy(); c.lambda_r, c.lambda_b, c.lambda_0 = lam
        parts.append(oracle.loss_grad(c, p, *xs)[0])
    c = cfg.copy(); c.lambda_r, c.lambda_b, c.lambda_0 = 0.25, 3.0, 7.5
    g, _ = oracle.loss_grad(c, p, *xs)
    assert rel_err(g, 0.25 * parts[0] + 3.0 * parts[1] + 7.5 * parts[2]) < 1e-12


def test_closed_form_solutions_satisfy_their_pdes():
    """The analytic solutions the product reports errors against (pinn_exact_solution; tests/test_abi_cpu.py pins the library
    to these closed forms) really solve the PDEs of SURVEY.md Appendix B: residuals by central differences in float64."""
    rng = np.random.default_rng(4)
    h = 1e-4
    d1 = lambda f, x, k: (f(*[xi + (h if i == k else 0) for i, xi in enumerate(x)]) - f(*[xi - (h if i == k else 0) for i, xi in enumerate(x)])) / (2 * h)
    d2 = lambda f, x, k: (f(*[xi + (h if i == k else 0) for i, xi in enumerate(x)]) - 2 * f(*x) + f(*[xi - (h if i == k else 0) for i, xi in enumerate(x)])) / h ** 2
    pi = np.pi
    # Helmholtz, k = 1: u_xx + u_yy + k^2 u = q = (k^2 - pi^2 - 16 pi^2) u*
    u = lambda x, y: np.sin(pi * x) * np.sin(4 * pi * y)
    pt = (rng.uniform(-1, 1, 50), rng.uniform(-1, 1, 50))
    r = d2(u, pt, 0) + d2(u, pt, 1) + u(*pt) - (1 - pi ** 2 - 16 * pi ** 2) * u(*pt)
    assert np.abs(r).max() < 1e-4
    # heat, alpha = 0.05: u_t = alpha (u_xx + u_yy)
    alpha = 0.05
    u = lambda x, y, t: np.sin(pi * x) * np.sin(pi * y) * np.exp(-2 * alpha * pi ** 2 * t)
    pt = (rng.uniform(-1, 1, 50), rng.uniform(-1, 1, 50), rng.uniform(0.1, 0.9, 50))
    assert np.abs(d1(u, pt, 2) - alpha * (d2(u, pt, 0) + d2(u, pt, 1))).max() < 1e-5
    # Navier-Stokes, nu = 0.01: Taylor-Green vortex
    nu = 0.01
    U = lambda x, y, t: -np.cos(x) * np.sin(y) * np.exp(-2 * nu * t)
    V = lambda x, y, t: np.sin(x) * np.cos(y) * np.exp(-2 * nu * t)
    Pp = lambda x, y, t: -0.25 * (np.cos(2 * x) + np.cos(2 * y)) * np.exp(-4 * nu * t)
    pt = (rng.uniform(0, 2 * pi, 50), rng.uniform(0, 2 * pi, 50), rng.uniform(0.1, 0.9, 50))
    r1 = d1(U, pt, 2) + U(*pt) * d1(U, pt, 0) + V(*pt) * d1(U, pt, 1) + d1(Pp, pt, 0) - nu * (d2(U, pt, 0) + d2(U, pt, 1))
    r2 = d1(V, pt, 2) + U(*pt) * d1(V, pt, 0) + V(*pt) * d1(V, pt, 1) + d1(Pp, pt, 1) - nu * (d2(V, pt, 0) + d2(V, pt, 1))
    r3 = d1(U, pt, 0) + d1(V, pt, 1)
    assert max(np.abs(r1).max(), np.abs(r2).max(), np.abs(r3).max()) < 1e-6
