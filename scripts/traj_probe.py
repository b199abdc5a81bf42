"""GPU probe: 100-step Adam trajectories — GPU FP32 vs oracle FP64 vs oracle FP32 (to state the tolerance)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load
from oracle import oracle as O
P = load()
for which, n in ((1, 4096), (3, 4096)):
    ocfg = O.baseline_config(which, n_interior=n, n_boundary=256, n_initial=0 if which == 3 else 256)
    c = P.PinnConfig()
    for f, _ in P.PinnConfig._fields_:
        setattr(c, f, getattr(ocfg, f))
    with P.Network(c) as net:
        w0 = net.get_params()
        lg = np.array([net.train_step()["loss"] for _ in range(100)])
        wg = net.get_params()
    p64, _, _, l64, _ = O.train(ocfg, w0.astype(np.float64), 100, dtype=np.float64)
    p32, _, _, l32, _ = O.train(ocfg, w0, 100, dtype=np.float32)
    for k in (1, 10, 30, 100):
        print(f"C{which} step {k:3d}: loss rel gpu-vs-f64 {abs(lg[k-1]/l64[k-1,0]-1):.2e}  cpu32-vs-f64 {abs(l32[k-1,0]/l64[k-1,0]-1):.2e}")
    print(f"C{which} max loss rel over 100: gpu {np.abs(lg/l64[:,0]-1).max():.2e} cpu32 {np.abs(l32[:,0]/l64[:,0]-1).max():.2e}")
    print(f"C{which} params max abs diff after 100: gpu-vs-f64 {np.abs(wg-p64).max():.2e} cpu32-vs-f64 {np.abs(p32-p64).max():.2e} gpu-vs-cpu32 {np.abs(wg-p32).max():.2e}")
