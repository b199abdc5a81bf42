"""C3 (2->[64x4]->1 Helmholtz, 1,048,576 points) on the round-1 tcgen05 kernel: launch shape and kernel time per precision.
Usage: c3_probe.py [n]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load
from oracle import oracle as O
P = load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
for prec, name in ((1, "bf16x3"), (2, "bf16"), (0, "fp32")):
    ocfg = O.baseline_config(3, n_interior=n, n_boundary=n // 32)
    c = P.PinnConfig()
    for f, _ in P.PinnConfig._fields_:
        setattr(c, f, getattr(ocfg, f))
    c.precision = prec
    with P.Network(c) as net:
        for _ in range(2):
            net.train_step()
        net.synchronize()
        net.profile_enable(1)
        for _ in range(3):
            net.train_step()
        net.synchronize()
        pr = net.profile_get()
        kms = pr["step_kernel_ms"] / max(pr["step_kernel_count"], 1)
        print(f"C3 {name}: kernel_id {pr['kernel_id']} grid {pr['grid']} block {pr['block']} smem {pr['smem_bytes']}: interior kernel {kms:.3f} ms = {n / kms * 1e-3:.1f} Mpts/s", flush=True)
