"""GPU probe: loss / gradient error of the FP32, BF16X3 and BF16 paths against the float64 oracle, plus timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load
from oracle import oracle as O
P = load()
cases = [(O.PDE_HELMHOLTZ, 64, 4, 4000), (O.PDE_BURGERS, 64, 3, 3000), (O.PDE_ALLEN_CAHN, 128, 6, 2000),
         (O.PDE_NAVIER_STOKES, 128, 3, 1500), (O.PDE_NAVIER_STOKES, 256, 8, 1100), (O.PDE_HEAT, 64, 1, 700)]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
for pde, H, L, n in cases:
    ocfg = O.make_config(pde, H, L, n, 300, 0 if pde == O.PDE_HELMHOLTZ else 300)
    xs = [O.sample(ocfg, k) for k in range(3)]
    rng = np.random.default_rng(1)
    w = O.init_params(ocfg)
    w = (w + 0.02 * rng.standard_normal(w.size)).astype(np.float32)
    g_ref, s_ref = O.loss_grad(ocfg, w.astype(np.float64), *xs)
    l_ref = O.loss_from_sums(ocfg, s_ref)[0]
    for prec, name in ((0, "fp32"), (1, "bf16x3"), (2, "bf16")):
        c = P.PinnConfig()
        for f, _ in P.PinnConfig._fields_:
            setattr(c, f, getattr(ocfg, f))
        c.precision = prec
        try:
            with P.Network(c) as net:
                net.set_params(w)
                out = net.loss_grad()
                g = net.get_grads()
                Y, r = net.residual(xs[0][:64])
                net.synchronize(); t0 = time.perf_counter()
                for _ in range(5):
                    net.loss_grad()
                net.synchronize(); dt = (time.perf_counter() - t0) / 5
            Y_ref, r_ref = O.residual(ocfg, w.astype(np.float64), xs[0][:64])
            print(f"pde{pde} H{H} L{L} {name:7s}: loss rel {abs(out['loss']/l_ref-1):.2e}  grad rel {np.abs(g-g_ref).max()/np.abs(g_ref).max():.2e}"
                  f"  Y rel {np.abs(Y-Y_ref).max()/np.abs(Y_ref).max():.2e}  {dt*1e3:.2f} ms", flush=True)
        except P.PinnError as e:
            print(f"pde{pde} H{H} L{L} {name:7s}: {e}", flush=True)
