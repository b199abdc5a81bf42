"""CPU tests of the C-ABI boundary: the shared library loads, exports every symbol the header
declares, its host-only helpers agree with the oracle, and it FAILS LOUDLY without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import ROOT


@pytest.fixture(scope="module")
def P():
    from __graft_entry__ import build, load
    if not os.path.exists(os.path.join(ROOT, "cuda-pinn_b200", "libpinn_b200.so")):
        build()
    return load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "pinn_abi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pinn_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(P):
    L = P.lib()
    declared = header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L, name), f"libpinn_b200.so does not export {name}"
    assert sorted(P.ABI_SYMBOLS) == declared, "python binding and header disagree on the ABI surface"


def test_library_exports_the_tensor_abi(P):
    """include/pinn_tensor_abi.h: every declared entry point is exported, and argument checks that need no GPU answer
    PINN_ERR_INVALID_ARGUMENT (no compute call is made here)."""
    src = open(os.path.join(ROOT, "include", "pinn_tensor_abi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    declared = sorted(set(re.findall(r"\b(pinn_[a-z_0-9]+)\s*\(", src)))
    assert declared == ["pinn_device_alloc", "pinn_device_copy", "pinn_device_download", "pinn_device_free", "pinn_device_upload",
                        "pinn_tensor_add_multiple", "pinn_tensor_broadcast", "pinn_tensor_dot", "pinn_tensor_dot_scratch", "pinn_tensor_einsum",
                        "pinn_tensor_elementwise", "pinn_tensor_matmul", "pinn_tensor_normalize_columns"]
    L = P.lib()
    for name in declared:
        assert hasattr(L, name), f"libpinn_b200.so does not export {name}"
    assert L.pinn_tensor_elementwise(0, 0, None, 4, None, 4, None, 4, 1, None) == 3       # null pointers
    assert L.pinn_tensor_matmul(0, None, 0, None, 0, None, 2, 2, 2, None) == 3
    assert L.pinn_tensor_dot(0, None, None, 8, None, None, None) == 3
    assert L.pinn_device_alloc(None, 16, None) == 3


def test_config_struct_layout_matches_oracle_mirror(P, oracle):
    assert C.sizeof(P.PinnConfig) == C.sizeof(oracle.PinnConfig) == 128
    assert [f[0] for f in P.PinnConfig._fields_] == [f[0] for f in oracle.PinnConfig._fields_]


@pytest.mark.parametrize("which", [1, 2, 3, 4, 5])
def test_default_configs_match_baseline(P, oracle, which):
    cfg = P.default_config(which)
    ocfg = oracle.baseline_config(which)
    for f, _ in P.PinnConfig._fields_:
        a, b = getattr(cfg, f), getattr(ocfg, f)
        assert a == pytest.approx(b), f
    assert P.lib().pinn_num_params(C.byref(cfg)) == oracle.num_params(ocfg) == [0, 3021, 3021, 12737, 83201, 462339][which]
    assert P.lib().pinn_num_channels(C.byref(cfg)) == oracle.num_channels(ocfg) == [0, 4, 4, 5, 6, 6][which]


def test_shard_ranges_bit_exact_with_oracle(P, oracle):
    for n in (0, 1, 31, 25600, (1 << 26) + 7, (1 << 40) + 12345):
        for R in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(R):
                b, e = P.shard_range(n, r, R)
                assert (b, e) == oracle.shard_range(n, r, R) == ((r * n) // R, ((r + 1) * n) // R)
                assert b == prev
                prev = e
            assert prev == n


def test_no_gpu_means_loud_failure_not_fallback(P):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(P.PinnError) as ei:
        P.Network(P.default_config(1))
    assert ei.value.code == 4 and "no CPU fallback" in str(ei.value)


def test_invalid_configs_are_rejected_with_reference_style_messages(P):
    import torch
    cfg = P.default_config(1, width=18)
    with pytest.raises(P.PinnError) as ei:
        P.Network(cfg)
    # width check precedes the device probe, so this also holds on the CPU box
    assert ei.value.code == 3 and "pinn_create: width must be a multiple of 4" in str(ei.value)
    with pytest.raises(P.PinnError) as ei:
        P.Network(P.default_config(1, d_in=3))
    assert "d_in does not match the pde" in str(ei.value)
    with pytest.raises(P.PinnError):
        P.Network(P.default_config(1, rank=2, nranks=2))
    assert P.lib().pinn_num_params(C.byref(P.default_config(1, depth=0))) == 0


def test_product_sources_never_touch_the_oracle():
    """The product path may not import, link or execute anything under oracle/."""
    bad = []
    for base in ("cuda-pinn_b200", "include", "src"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"import\s+oracle|from\s+oracle|liboracle|oracle/|oracle\.|include.*oracle|dlopen.*oracle", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_exact_solutions_match_closed_forms_and_oracle_targets(P, oracle):
    """pinn_exact_solution is host arithmetic: checked here against numpy closed forms and, for Navier-Stokes
    (whose Dirichlet/initial target IS the Taylor-Green solution), against the oracle's target function."""
    rng = np.random.default_rng(7)
    x2 = rng.uniform(-1, 1, (257, 2)).astype(np.float32)
    x3 = np.concatenate([rng.uniform(-1, 1, (257, 2)), rng.uniform(0, 1, (257, 1))], axis=1).astype(np.float32)
    xn = np.concatenate([rng.uniform(0, 2 * np.pi, (257, 2)), rng.uniform(0, 1, (257, 1))], axis=1).astype(np.float32)
    d = x2.astype(np.float64)
    u = P.exact_solution(P.default_config(3), x2)
    assert np.allclose(u[:, 0], np.sin(np.pi * d[:, 0]) * np.sin(4 * np.pi * d[:, 1]), atol=1e-6)
    heat = P.default_config(4, pde=P.PDE_HEAT)
    d = x3.astype(np.float64)
    u = P.exact_solution(heat, x3)
    assert np.allclose(u[:, 0], np.sin(np.pi * d[:, 0]) * np.sin(np.pi * d[:, 1]) * np.exp(-2 * 0.05 * np.pi ** 2 * d[:, 2]), atol=1e-6)
    u = P.exact_solution(P.default_config(5), xn)
    assert u.shape == (257, 3)
    assert np.allclose(u, oracle.target(oracle.baseline_config(5), 1, xn), atol=2e-6)
    # initial target of the heat problem is the exact solution at t = 0
    x0 = x3.copy(); x0[:, 2] = 0
    ocfg_heat = oracle.make_config(oracle.PDE_HEAT, 16, 2, 64, 64, 64)
    assert np.allclose(P.exact_solution(heat, x0), oracle.target(ocfg_heat, 2, x0), atol=2e-6)
    for which in (1, 4):                       # Burgers, Allen-Cahn: no closed form
        with pytest.raises(P.PinnError) as e:
            P.exact_solution(P.default_config(which), x3[:, :P.default_config(which).d_in])
        assert e.value.code == 7
    assert P.exact_solution(P.default_config(3), np.zeros((0, 2), np.float32)).shape == (0, 1)


def test_training_driver_argument_errors_without_a_gpu():
    exe = os.path.join(ROOT, "bin", "pinn_train")
    if not os.path.exists(exe):
        pytest.skip("bin/pinn_train is built only where the reference headers are present")
    import subprocess
    r = subprocess.run([exe, "--bogus", "1"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 2 and "unknown option --bogus" in r.stderr
    import torch
    if not torch.cuda.is_available():          # no CPU fallback: the driver reports the library's error and exits 1
        r = subprocess.run([exe, "--config", "1", "--steps", "1"], capture_output=True, text=True, timeout=120)
        assert r.returncode == 1 and "pinn_train: " in r.stderr


def test_python_binding_enums_match_the_header(P):
    """pinn_precision / pinn_pde / pinn_point_set values of include/pinn_abi.h against the ctypes binding's constants."""
    src = open(os.path.join(ROOT, "include", "pinn_abi.h")).read()
    vals = {k: int(v) for k, v in re.findall(r"\b(PINN_(?:PREC|PDE|SET)_[A-Z0-9_]+)\s*=\s*(\d+)", src)}
    assert vals["PINN_PREC_FP32"] == P.PREC_FP32 and vals["PINN_PREC_BF16X3"] == P.PREC_BF16X3 and vals["PINN_PREC_BF16"] == P.PREC_BF16
    assert vals["PINN_PREC_BF16X6"] == P.PREC_BF16X6 and vals["PINN_PREC_FP16X3"] == P.PREC_FP16X3
    assert [vals[k] for k in ("PINN_PDE_BURGERS", "PINN_PDE_HELMHOLTZ", "PINN_PDE_ALLEN_CAHN", "PINN_PDE_NAVIER_STOKES", "PINN_PDE_HEAT")] == \
        [P.PDE_BURGERS, P.PDE_HELMHOLTZ, P.PDE_ALLEN_CAHN, P.PDE_NAVIER_STOKES, P.PDE_HEAT]
    assert [vals[k] for k in ("PINN_SET_INTERIOR", "PINN_SET_BOUNDARY", "PINN_SET_INITIAL")] == [P.SET_INTERIOR, P.SET_BOUNDARY, P.SET_INITIAL]
    assert sorted(v for k, v in vals.items() if k.startswith("PINN_PREC_")) == [0, 1, 2, 3, 4]
