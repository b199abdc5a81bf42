"""GPU tests of SURVEY.md 8(f).2/3: checkpoint save/load (bit-exact resume), grid evaluation against the
analytic solutions (checked against the float64 oracle's forward pass + numpy formulas), and the
training driver bin/pinn_train (src/main.cu over network<T> and the reference's tensor headers)."""
import csv
import os
import struct
import subprocess

import numpy as np
import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    from __graft_entry__ import load
    return load()


def pcfg(P, ocfg, **kw):
    c = P.PinnConfig()
    for f, _ in P.PinnConfig._fields_:
        setattr(c, f, getattr(ocfg, f))
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def test_checkpoint_resume_is_bit_exact_and_file_layout_is_flat(P, oracle, tmp_path):
    ocfg = oracle.make_config(oracle.PDE_BURGERS, 20, 8, 4099, 257, 257)
    path = str(tmp_path / "net.ckpt")
    with P.Network(pcfg(P, ocfg)) as a:
        a.train_steps(7)
        a.save(path)
        w7 = a.get_params(); m7, v7, t7 = a.get_optimizer()
        la = [a.train_step() for _ in range(5)]
        wa = a.get_params()
    # file layout: 64-byte header + params + m + v (float32, little-endian)
    raw = open(path, "rb").read()
    n = w7.size
    assert len(raw) == 64 + 12 * n
    magic, version, pde, d_in, d_out, width, depth, _, _, Pn, step = struct.unpack("<8sI5i2I2Q", raw[:56])
    assert (magic, version, pde, d_in, d_out, width, depth, Pn, step) == (b"PINNCKPT", 1, 0, 2, 1, 20, 8, n, 7) and t7 == 7
    body = np.frombuffer(raw[64:], dtype="<f4").reshape(3, n)
    assert np.array_equal(body[0].view(np.uint32), w7.view(np.uint32))
    assert np.array_equal(body[1].view(np.uint32), m7.view(np.uint32))
    assert np.array_equal(body[2].view(np.uint32), v7.view(np.uint32))
    # a fresh handle (different initial weights) resumed from the file continues bit-exactly
    with P.Network(pcfg(P, ocfg, seed_weights=12345)) as b:
        assert not np.array_equal(b.get_params(), w7)
        b.load(path)
        lb = [b.train_step() for _ in range(5)]
        wb = b.get_params()
    assert [x["step"] for x in lb] == [8, 9, 10, 11, 12]
    assert [np.float32(x["loss"]).view(np.uint32) for x in la] == [np.float32(x["loss"]).view(np.uint32) for x in lb]
    assert np.array_equal(wa.view(np.uint32), wb.view(np.uint32))


def test_checkpoint_rejects_mismatch_truncation_and_garbage(P, oracle, tmp_path):
    path = str(tmp_path / "net.ckpt")
    with P.Network(pcfg(P, oracle.make_config(oracle.PDE_BURGERS, 20, 8, 512, 64, 64))) as a:
        a.train_steps(2)
        a.save(path)
        with pytest.raises(P.PinnError) as e:
            a.save(str(tmp_path / "no_such_dir" / "x.ckpt"))
        assert e.value.code == 3 and "pinn_save: cannot open" in str(e.value)
    with P.Network(pcfg(P, oracle.make_config(oracle.PDE_BURGERS, 16, 8, 512, 64, 64))) as b:
        w0 = b.get_params()
        with pytest.raises(P.PinnError) as e:
            b.load(path)
        assert e.value.code == 3 and "network mismatch" in str(e.value)
        assert np.array_equal(b.get_params(), w0) and b.get_optimizer()[2] == 0     # untouched
        with pytest.raises(P.PinnError) as e:
            b.load(str(tmp_path / "missing.ckpt"))
        assert "cannot open" in str(e.value)
    raw = open(path, "rb").read()
    with P.Network(pcfg(P, oracle.make_config(oracle.PDE_BURGERS, 20, 8, 512, 64, 64))) as c:
        w0 = c.get_params()
        open(path, "wb").write(raw[:len(raw) // 2])
        with pytest.raises(P.PinnError) as e:
            c.load(path)
        assert "truncated" in str(e.value) and np.array_equal(c.get_params(), w0)
        open(path, "wb").write(b"not a checkpoint at all" * 10)
        with pytest.raises(P.PinnError) as e:
            c.load(path)
        assert "not a checkpoint" in str(e.value)


def grid_points(lo, hi, per_axis):
    axes = [lo[k] + (hi[k] - lo[k]) * (np.arange(per_axis, dtype=np.float32) / np.float32(per_axis - 1)) for k in range(len(lo))]
    return np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, len(lo)).astype(np.float32)


@pytest.mark.parametrize("pde,per_axis", [(1, 65), (4, 17), (3, 17)])
def test_grid_evaluation_matches_oracle_forward_and_analytic_solution(P, oracle, pde, per_axis):
    ocfg = oracle.make_config(pde, 16, 3, 2048, 256, 0 if pde == 1 else 256)
    lo = [0.0, 0.0, 0.0] if pde == 3 else [-1.0, -1.0, 0.0]
    hi = [2 * np.pi, 2 * np.pi, 1.0] if pde == 3 else [1.0, 1.0, 1.0]
    x = grid_points(np.float32(lo[:ocfg.d_in]), np.float32(hi[:ocfg.d_in]), per_axis)
    with P.Network(pcfg(P, ocfg)) as net:
        net.train_steps(20)
        w = net.get_params()
        ev = net.evaluate(per_axis)
        xs = net.get_points(0)[:4]
    assert ev["n"] == len(x) == per_axis ** ocfg.d_in
    y, r = oracle.residual(ocfg, w.astype(np.float64), x)                 # float64 oracle forward
    u = P.exact_solution(pcfg(P, ocfg), x).astype(np.float64)
    assert np.isfinite(xs).all()
    for o in range(ocfg.d_out):
        e = y[:, o, 0] - u[:, o]
        assert ev["rel_l2"][o] == pytest.approx(np.sqrt((e * e).sum() / (u[:, o] ** 2).sum()), rel=1e-4)
        assert ev["max_abs"][o] == pytest.approx(np.abs(e).max(), rel=1e-4, abs=1e-6)
    assert ev["mse_residual"] == pytest.approx((r * r).sum() / len(x), rel=1e-4)


def test_evaluation_without_analytic_solution_reports_residual_only(P, oracle):
    ocfg = oracle.make_config(oracle.PDE_BURGERS, 20, 8, 1024, 64, 64)
    with P.Network(pcfg(P, ocfg)) as net:
        ev = net.evaluate(33)
        assert ev["n"] == 33 * 33 and np.isnan(ev["rel_l2"][0]) and np.isnan(ev["max_abs"][0]) and ev["mse_residual"] > 0
        for bad in (1, 0, 5000):
            with pytest.raises(P.PinnError) as e:
                net.evaluate(bad)
            assert e.value.code == 3


def test_training_improves_helmholtz_accuracy(P, oracle):
    """Accuracy check of SURVEY.md Appendix B: the relative L2 error vs sin(pi x) sin(4 pi y) falls while training."""
    ocfg = oracle.make_config(oracle.PDE_HELMHOLTZ, 64, 4, 8192, 1024, 0)
    with P.Network(pcfg(P, ocfg, lambda_b=100.0)) as net:
        e0 = net.evaluate(65)
        l0 = net.train_step()["loss"]
        out = net.train_steps(3000)
        e1 = net.evaluate(65)
    assert out["loss"] < 0.05 * l0
    assert e1["mse_residual"] < 0.05 * e0["mse_residual"]
    assert e1["rel_l2"][0] < e0["rel_l2"][0]


# ------------------------------------------------------------------ training driver (src/main.cu)
TRAIN = os.path.join(ROOT, "bin", "pinn_train")


@pytest.mark.skipif(not os.path.exists(TRAIN), reason="bin/pinn_train is built only where the reference headers are present")
def test_training_driver_logs_checkpoints_resumes_and_matches_the_library(P, oracle, tmp_path):
    log, ck = str(tmp_path / "loss.csv"), str(tmp_path / "c.ckpt")
    r = subprocess.run([TRAIN, "--config", "1", "--steps", "60", "--log-every", "20", "--csv", log, "--save", ck, "--eval", "33"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "3021 parameters, 4 Taylor channels, 4096 interior / 256 boundary / 256 initial points" in r.stdout
    assert "grid 1089 nodes: residual mse" in r.stdout and "saved " + ck in r.stdout
    rows = list(csv.DictReader(open(log)))
    assert [int(x["step"]) for x in rows] == [20, 40, 60]
    with P.Network(P.default_config(1)) as net:                       # same steps driven through the binding
        want = [net.train_steps(20) for _ in range(3)]
        w60 = net.get_params()
    for row, w in zip(rows, want):
        assert np.float32(row["loss"]) == np.float32(w["loss"]) and np.float32(row["mse_residual"]) == np.float32(w["mse_residual"])
    body = np.frombuffer(open(ck, "rb").read()[64:], dtype="<f4").reshape(3, -1)
    assert np.array_equal(body[0].view(np.uint32), w60.view(np.uint32))
    # resume: 20 more steps from the file == steps 61..80 of an uninterrupted run
    r2 = subprocess.run([TRAIN, "--config", "1", "--steps", "20", "--log-every", "20", "--resume", ck, "--csv", log],
                        capture_output=True, text=True, timeout=300)
    assert r2.returncode == 0, r2.stdout + r2.stderr
    rows = list(csv.DictReader(open(log)))
    with P.Network(P.default_config(1)) as net:
        w80 = net.train_steps(80)
    assert int(rows[0]["step"]) == 80 and np.float32(rows[0]["loss"]) == np.float32(w80["loss"])


@pytest.mark.skipif(not os.path.exists(TRAIN), reason="bin/pinn_train is built only where the reference headers are present")
def test_training_driver_precision_option_selects_the_mma_kernel(P, tmp_path):
    log = str(tmp_path / "loss.csv")
    r = subprocess.run([TRAIN, "--config", "1", "--precision", "fp16x3", "--steps", "40", "--log-every", "20", "--csv", log],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = list(csv.DictReader(open(log)))
    with P.Network(P.default_config(1, precision=P.PREC_FP16X3)) as net:
        want = [net.train_steps(20) for _ in range(2)]
        assert net.profile_get()["kernel_id"] == 204
    for row, w in zip(rows, want):
        assert np.float32(row["loss"]) == np.float32(w["loss"])
    r = subprocess.run([TRAIN, "--precision", "fp8"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 2 and "--precision must be one of" in r.stderr


@pytest.mark.skipif(not os.path.exists(TRAIN), reason="bin/pinn_train is built only where the reference headers are present")
def test_training_driver_rejects_bad_arguments():
    for args in (["--steps"], ["--bogus", "1"], ["--steps", "abc"], ["steps", "3"], ["--log-every", "0"]):
        r = subprocess.run([TRAIN] + args, capture_output=True, text=True, timeout=60)
        assert r.returncode == 2 and "usage: pinn_train" in r.stderr, args
    r = subprocess.run([TRAIN, "--config", "9", "--steps", "1"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "network::baseline: config index must be in 1..5" in r.stderr
    r = subprocess.run([TRAIN, "--config", "1", "--steps", "0", "--resume", "/nonexistent.ckpt"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "pinn_load: cannot open" in r.stderr


def test_train_steps_replay_a_cuda_graph_and_say_so(P):
    """pinn_graph_state: after a step the handle replays its captured graph (C1: mma kernel + fused tail; C5-like net: the cooperative
    producer / dW-server launch + value launch + tail); with per-kernel profiling on it does not, and the note stays empty (no
    capture failure to report)."""
    for which, n in ((1, 4096), (4, 2100), (5, 2100)):      # mma kernel | two plain-launched CTAs per SM | cooperative launch
        cfg = P.default_config(which)
        cfg.n_interior, cfg.n_boundary, cfg.n_initial = n, 300, 300
        with P.Network(cfg) as net:
            assert net.graph_state() == (False, "")
            net.train_step()
            on, note = net.graph_state()
            assert on and note == "", (which, on, note)
