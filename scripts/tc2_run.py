"""Run K training steps of BASELINE config `which` with n interior points on the BF16 tensor path (for ncu / timing).
Usage: tc2_run.py which n steps [q] [zdepth]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
which, n, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
if len(sys.argv) > 4: os.environ["PINN_TC2_Q"] = sys.argv[4]
if len(sys.argv) > 5: os.environ["PINN_TC2_ZD"] = sys.argv[5]
from __graft_entry__ import load
from oracle import oracle as O
P = load()
ocfg = O.baseline_config(which, n_interior=n, n_boundary=max(n // 32, 256), n_initial=max(n // 32, 256))
c = P.PinnConfig()
for f, _ in P.PinnConfig._fields_:
    setattr(c, f, getattr(ocfg, f))
c.precision = 2
with P.Network(c) as net:
    net.train_step(); net.synchronize()
    net.profile_enable(1)
    t0 = time.perf_counter()
    for _ in range(K):
        o = net.train_step()
    net.synchronize()
    dt = (time.perf_counter() - t0) / K
    pr = net.profile_get()
    kms = pr["step_kernel_ms"] / max(pr["step_kernel_count"], 1)
    print(f"C{which} n={n} q={os.environ.get('PINN_TC2_Q', '-')} zd={os.environ.get('PINN_TC2_ZD', '-')} kid {pr['kernel_id']}: step {dt * 1e3:.2f} ms;"
          f" interior kernel {kms:.3f} ms = {n / kms * 1e-3:.2f} Mpts/s; loss {o['loss']:.5f}", flush=True)
