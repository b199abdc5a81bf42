"""One loss_grad call of a small six-channel net on the BF16 tensor path (for compute-sanitizer). Usage: tc2_one.py pde H L n"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load
from oracle import oracle as O
P = load()
pde, H, L, n = (int(a) for a in sys.argv[1:5])
ocfg = O.make_config(pde, H, L, n, 300, 300)
c = P.PinnConfig()
for f, _ in P.PinnConfig._fields_:
    setattr(c, f, getattr(ocfg, f))
c.precision = 2
with P.Network(c) as net:
    out = net.loss_grad()
    g = net.get_grads()
    print("kernel", net.profile_get()["kernel_id"], "loss", out["loss"], "|g|max", float(np.abs(g).max()))
