"""GPU probe for the producer / dW-server tensor-core kernel (pinn_fused_tc2.cuh): loss / gradient error against the
float64 oracle per parameter block, next to the round-1 kernel (PINN_TC2=0), plus kernel timings at larger sizes.
Usage: tc2_probe.py [check|time|all] [q] [zdepth]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load
from oracle import oracle as O
P = load()
what = sys.argv[1] if len(sys.argv) > 1 else "all"
if len(sys.argv) > 2: os.environ["PINN_TC2_Q"] = sys.argv[2]
if len(sys.argv) > 3: os.environ["PINN_TC2_ZD"] = sys.argv[3]


def pcfg(ocfg, prec):
    c = P.PinnConfig()
    for f, _ in P.PinnConfig._fields_:
        setattr(c, f, getattr(ocfg, f))
    c.precision = prec
    return c


def blocks(d_in, H, L, d_out):
    off, out = 0, []
    for l in range(1, L + 2):
        fi = d_in if l == 1 else H
        fo = d_out if l == L + 1 else H
        out.append((f"W{l}", off, off + fi * fo)); off += fi * fo
        out.append((f"b{l}", off, off + fo)); off += fo
    return out


if what in ("check", "all"):
    cases = [(O.PDE_NAVIER_STOKES, 128, 2, 17), (O.PDE_ALLEN_CAHN, 128, 6, 2000), (O.PDE_HEAT, 128, 3, 1500),
             (O.PDE_NAVIER_STOKES, 256, 8, 1100), (O.PDE_NAVIER_STOKES, 256, 3, 9000), (O.PDE_ALLEN_CAHN, 128, 4, 30011)]
    for pde, H, L, n in cases:
        ocfg = O.make_config(pde, H, L, n, 300, 300)
        xs = [O.sample(ocfg, k) for k in range(3)]
        rng = np.random.default_rng(1)
        w = O.init_params(ocfg)
        w = (w + 0.02 * rng.standard_normal(w.size)).astype(np.float32)
        g_ref, s_ref = O.loss_grad(ocfg, w.astype(np.float64), *xs)
        l_ref = O.loss_from_sums(ocfg, s_ref)[0]
        for tc2 in ("1", "0"):
            os.environ["PINN_TC2"] = tc2
            try:
                with P.Network(pcfg(ocfg, 2)) as net:
                    net.set_params(w)
                    out = net.loss_grad()
                    g = net.get_grads()
                    out2 = net.loss_grad()
                    g2 = net.get_grads()
                    kid = net.profile_get()["kernel_id"]
                gm = np.abs(g_ref).max()
                print(f"pde{pde} H{H} L{L} n{n} tc2={tc2} kid {kid}: loss rel {abs(out['loss'] / l_ref - 1):.2e}  grad rel {np.abs(g - g_ref).max() / gm:.2e}"
                      f"  rerun identical {bool(np.array_equal(g, g2) and out['loss'] == out2['loss'])}", flush=True)
                if tc2 == "1":
                    d_out = 3 if pde == O.PDE_NAVIER_STOKES else 1
                    bad = [(nm, np.abs(g[a:b] - g_ref[a:b]).max() / gm) for nm, a, b in blocks(3, H, L, d_out)]
                    print("    per block: " + " ".join(f"{nm}:{e:.1e}" for nm, e in bad), flush=True)
            except P.PinnError as e:
                print(f"pde{pde} H{H} L{L} n{n} tc2={tc2}: {e}", flush=True)

if what in ("time", "all"):
    for which, n in ((5, 262144), (5, 1048576), (4, 1048576)):
        for tc2 in ("1", "0"):
            os.environ["PINN_TC2"] = tc2
            ocfg = O.baseline_config(which, n_interior=n, n_boundary=n // 32, n_initial=n // 32)
            with P.Network(pcfg(ocfg, 2)) as net:
                for _ in range(2):
                    net.train_step()
                net.synchronize()
                t0 = time.perf_counter()
                K = 3
                for _ in range(K):
                    o = net.train_step()
                net.synchronize()
                dt = (time.perf_counter() - t0) / K
                net.profile_enable(1)
                for _ in range(2):
                    net.train_step()
                net.synchronize()
                pr = net.profile_get()
                kms = pr["step_kernel_ms"] / max(pr["step_kernel_count"], 1)
                print(f"C{which} n={n} tc2={tc2} kid {pr['kernel_id']}: step {dt * 1e3:.2f} ms = {n / dt * 1e-6:.2f} Mpts/s; interior kernel {kms:.2f} ms"
                      f" = {n / kms * 1e-3:.2f} Mpts/s; loss {o['loss']:.5f}", flush=True)
