"""Where does a C2 step go?  Step time (CUDA-graph replay, back to back) with and without the boundary/initial sets."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from __graft_entry__ import load
P = load()
for prec in (0, 1, 3):
    for nb in (800, 0):
        cfg = P.default_config(2, precision=prec, n_boundary=nb, n_initial=nb)
        with P.Network(cfg) as net:
            net.train_steps(50); net.synchronize()
            t0 = time.perf_counter(); net.train_steps(500); net.synchronize(); dt = (time.perf_counter() - t0) / 500
            print(f"prec={prec} nb=n0={nb}: step {dt*1e6:.1f} us")
