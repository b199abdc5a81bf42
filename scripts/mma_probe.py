"""Quick B200 probe of the width-20 mma.sync kernel: parity vs the float64 oracle and step time vs the FP32 kernel."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from __graft_entry__ import load
sys.path.insert(0, "oracle")
import oracle as O
P = load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 25600
ocfg = O.baseline_config(2, n_interior=n)
def pcfg(**kw):
    c = P.PinnConfig()
    for f, _ in P.PinnConfig._fields_: setattr(c, f, getattr(ocfg, f))
    for k, v in kw.items(): setattr(c, k, v)
    return c
xs = [O.sample(ocfg, k) for k in range(3)]
w = O.init_params(ocfg).astype(np.float32)
g_ref, s_ref = O.loss_grad(ocfg, w.astype(np.float64), *xs)
l_ref = O.loss_from_sums(ocfg, s_ref)[0]
for prec in (0, 1, 3, 4):
    with P.Network(pcfg(precision=prec)) as net:
        out = net.loss_grad(); g = net.get_grads()
        err = np.abs(g - g_ref).max() / np.abs(g_ref).max()
        # per-layer error profile
        net.train_steps(20); net.synchronize()
        t0 = time.perf_counter(); net.train_steps(200); net.synchronize(); dt = (time.perf_counter() - t0) / 200
        net.profile_enable(True); [net.train_step(read_loss=False) for _ in range(20)]; net.synchronize()
        pr = net.profile_get()
        print(f"prec={prec} kernel_id={pr['kernel_id']} loss_rel={abs(out['loss']/l_ref-1):.2e} grad_rel={err:.2e} "
              f"step={dt*1e6:.1f} us ({n/dt/1e6:.1f} Mpts/s) kernel={pr['step_kernel_ms']/max(pr['step_kernel_count'],1)*1e3:.1f} us grid={pr['grid']}")

# 100-step drift against the float64 oracle on the C1 problem
ocfg = O.baseline_config(1)
w0 = O.init_params(ocfg).astype(np.float32)
p_ref, _, _, l_ref, _ = O.train(ocfg, w0.astype(np.float64), 100, dtype=np.float64)
p32, _, _, l32, _ = O.train(ocfg, w0, 100, dtype=np.float32)
print(f"cpu f32 oracle: loss30={np.abs(l32[:30,0]/l_ref[:30,0]-1).max():.2e} loss100={np.abs(l32[:,0]/l_ref[:,0]-1).max():.2e} params={np.abs(p32-p_ref).max():.2e}")
for prec in (0, 1, 3, 4):
    with P.Network(pcfg(precision=prec)) as net:
        losses = np.array([net.train_step()["loss"] for _ in range(100)])
        w = net.get_params()
    print(f"prec={prec} 100 steps: loss30={np.abs(losses[:30]/l_ref[:30,0]-1).max():.2e} loss100={np.abs(losses/l_ref[:,0]-1).max():.2e} params={np.abs(w-p_ref).max():.2e}")
