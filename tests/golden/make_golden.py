#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ (run once, in the build container).

The reference (Deadman621/CUDA-PINN) has no train step to import or run, so the
fixtures are produced by an INDEPENDENT implementation of SURVEY.md Appendix B:

  * sampler / Glorot init: numpy uint64 + float64 arithmetic (exact for these operand sizes);
  * network, Taylor channels, residuals, loss: plain torch float64, derivatives by
    torch.autograd (double backward) — no hand-written tangents or adjoints;
  * Adam: torch float64, written from the formula.

tests/test_oracle.py checks the C oracle (oracle/pinn_oracle.c) against these files; the GPU
parity tests check the CUDA path against them as well.  Usage:  python tests/golden/make_golden.py
"""
import json
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
M64 = (1 << 64) - 1

PDE_BURGERS, PDE_HELMHOLTZ, PDE_ALLEN_CAHN, PDE_NAVIER_STOKES, PDE_HEAT = 0, 1, 2, 3, 4
PDE_SHAPE = {0: (2, 1, 1, 1), 1: (2, 1, 2, 2), 2: (3, 1, 2, 2), 3: (3, 3, 2, 2), 4: (3, 1, 2, 2)}  # d_in,d_out,n_space,n2


def f32(v):
    """hyper-parameters cross the C ABI as float: use the float32-rounded values here too"""
    return float(np.float32(v))


def mix(seed, ctr):
    z = (seed + (ctr + 1) * 0x9E3779B97F4A7C15) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def u01(seed, ctr):
    return (mix(seed, ctr) >> 40) / float(1 << 24)   # exact in float64 (and in float32)


def domain(pde, d_in, n_space):
    lo, hi = [], []
    for k in range(d_in):
        is_t = d_in > n_space and k == d_in - 1
        if is_t:
            lo.append(np.float32(0.0)); hi.append(np.float32(1.0))
        elif pde == PDE_NAVIER_STOKES:
            lo.append(np.float32(0.0)); hi.append(np.float32(6.28318530717958647692))
        else:
            lo.append(np.float32(-1.0)); hi.append(np.float32(1.0))
    return lo, hi


def sample(cfg, which, begin, end):
    pde, d_in = cfg["pde"], cfg["d_in"]
    n_space = PDE_SHAPE[pde][2]
    lo, hi = domain(pde, d_in, n_space)
    seed = (cfg["seed_interior"], cfg["seed_boundary"], cfg["seed_initial"])[which]
    has_t = d_in > n_space
    x = np.zeros((end - begin, d_in), dtype=np.float32)
    for i in range(begin, end):
        face = i % (2 * n_space)
        axis, side = face // 2, face % 2
        for k in range(d_in):
            span = np.float32(hi[k] - lo[k])
            # fmaf(span, u, lo): product and sum are exact in float64 for these operands → one rounding
            v = np.float32(float(span) * u01(seed, i * d_in + k) + float(lo[k]))
            if which == 1 and k == axis:
                v = hi[k] if side else lo[k]
            if which == 2 and has_t and k == d_in - 1:
                v = lo[k]
            x[i - begin, k] = v
    return x


def layer_dims(cfg):
    L, H = cfg["depth"], cfg["width"]
    return [((cfg["d_in"] if l == 1 else H), (cfg["d_out"] if l == L + 1 else H)) for l in range(1, L + 2)]


def init_params(cfg):
    out = []
    for l, (fi, fo) in enumerate(layer_dims(cfg), start=1):
        a = np.float32(math.sqrt(6.0 / (fi + fo)))
        two_a = np.float32(2.0) * a
        w = np.array([np.float32(float(two_a) * u01(cfg["seed_weights"] + l, k) - float(a)) for k in range(fi * fo)],
                     dtype=np.float32)
        out.append(w)
        out.append(np.zeros(fo, dtype=np.float32))
    return np.concatenate(out)


def unflatten(cfg, flat):
    Ws, bs, off = [], [], 0
    for fi, fo in layer_dims(cfg):
        Ws.append(flat[off:off + fi * fo].reshape(fi, fo)); off += fi * fo
        bs.append(flat[off:off + fo]); off += fo
    return Ws, bs


def net(cfg, flat, x):
    Ws, bs = unflatten(cfg, flat)
    a = x
    for l in range(cfg["depth"]):
        a = torch.tanh(a @ Ws[l] + bs[l])
    return a @ Ws[-1] + bs[-1]


def taylor_channels(cfg, flat, x):
    """y[n, d_out, C] via autograd: value, d/dx_i (i<d_in), d2/dx_i^2 (i<n2)."""
    d_in, d_out, _, n2 = PDE_SHAPE[cfg["pde"]]
    x = x.clone().requires_grad_(True)
    y = net(cfg, flat, x)
    chans = []
    for o in range(d_out):
        g = torch.autograd.grad(y[:, o].sum(), x, create_graph=True)[0]
        cols = [y[:, o]] + [g[:, i] for i in range(d_in)]
        for i in range(n2):
            g2 = torch.autograd.grad(g[:, i].sum(), x, create_graph=True)[0]
            cols.append(g2[:, i])
        chans.append(torch.stack(cols, dim=1))
    return torch.stack(chans, dim=1)


def residual(cfg, x, Y):
    pde = cfg["pde"]
    pi = math.pi
    if pde == PDE_BURGERS:
        u, ux, ut, uxx = Y[:, 0, 0], Y[:, 0, 1], Y[:, 0, 2], Y[:, 0, 3]
        return (ut + u * ux - (0.01 / pi) * uxx)[:, None]
    if pde == PDE_HELMHOLTZ:
        k = cfg["pde_param"] if cfg["pde_param"] != 0 else 1.0
        q = (k * k - pi * pi - 16 * pi * pi) * torch.sin(pi * x[:, 0]) * torch.sin(4 * pi * x[:, 1])
        return (Y[:, 0, 3] + Y[:, 0, 4] + k * k * Y[:, 0, 0] - q)[:, None]
    if pde == PDE_ALLEN_CAHN:
        u = Y[:, 0, 0]
        return (Y[:, 0, 3] - 1e-4 * (Y[:, 0, 4] + Y[:, 0, 5]) + 5 * u ** 3 - 5 * u)[:, None]
    if pde == PDE_HEAT:
        return (Y[:, 0, 3] - 0.05 * (Y[:, 0, 4] + Y[:, 0, 5]))[:, None]
    if pde == PDE_NAVIER_STOKES:
        nu = 0.01
        U, V, P = Y[:, 0], Y[:, 1], Y[:, 2]
        r1 = U[:, 3] + U[:, 0] * U[:, 1] + V[:, 0] * U[:, 2] + P[:, 1] - nu * (U[:, 4] + U[:, 5])
        r2 = V[:, 3] + U[:, 0] * V[:, 1] + V[:, 0] * V[:, 2] + P[:, 2] - nu * (V[:, 4] + V[:, 5])
        r3 = U[:, 1] + V[:, 2]
        return torch.stack([r1, r2, r3], dim=1)
    raise ValueError(pde)


def target(cfg, which, x):
    pde, pi = cfg["pde"], math.pi
    X = x[:, 0]
    Yc = x[:, 1] if x.shape[1] > 1 else None
    T = x[:, 2] if x.shape[1] > 2 else None
    z = torch.zeros_like(X)
    if pde == PDE_BURGERS:
        return (-torch.sin(pi * X) if which == 2 else z)[:, None]
    if pde == PDE_HELMHOLTZ:
        return z[:, None]
    if pde == PDE_ALLEN_CAHN:
        return (X * X * torch.cos(pi * X) * (Yc * Yc * torch.cos(pi * Yc)))[:, None]
    if pde == PDE_HEAT:
        return (torch.sin(pi * X) * torch.sin(pi * Yc) if which == 2 else z)[:, None]
    if pde == PDE_NAVIER_STOKES:
        nu = 0.01
        e2, e4 = torch.exp(-2 * nu * T), torch.exp(-4 * nu * T)
        return torch.stack([-torch.cos(X) * torch.sin(Yc) * e2, torch.sin(X) * torch.cos(Yc) * e2,
                            -0.25 * (torch.cos(2 * X) + torch.cos(2 * Yc)) * e4], dim=1)
    raise ValueError(pde)


def loss_terms(cfg, flat, xf, xb, x0):
    Y = taylor_channels(cfg, flat, xf)
    r = residual(cfg, xf, Y)
    sums = [(r ** 2).sum()]
    for which, x in ((1, xb), (2, x0)):
        if len(x):
            sums.append(((net(cfg, flat, x) - target(cfg, which, x)) ** 2).sum())
        else:
            sums.append(torch.zeros((), dtype=torch.float64))
    return sums, Y, r


def total_loss(cfg, sums):
    t = cfg["lambda_r"] * sums[0] / cfg["n_interior"] + cfg["lambda_b"] * sums[1] / cfg["n_boundary"]
    if cfg["n_initial"]:
        t = t + cfg["lambda_0"] * sums[2] / cfg["n_initial"]
    return t


def make_case(name, pde, width, depth, nf, nb, n0, adam_steps=3, pde_param=0.0, seed_shift=0):
    d_in, d_out, n_space, n2 = PDE_SHAPE[pde]
    cfg = dict(pde=pde, d_in=d_in, d_out=d_out, width=width, depth=depth,
               n_interior=nf, n_boundary=nb, n_initial=n0,
               seed_interior=0x5EED0001 + seed_shift, seed_boundary=0x5EED0002 + seed_shift,
               seed_initial=0x5EED0003 + seed_shift, seed_weights=0x5EED0100 + seed_shift,
               lr=f32(1e-3), beta1=f32(0.9), beta2=f32(0.999), eps=f32(1e-8), lambda_r=1.0, lambda_b=1.0, lambda_0=1.0,
               pde_param=pde_param)
    xf, xb, x0 = sample(cfg, 0, 0, nf), sample(cfg, 1, 0, nb), sample(cfg, 2, 0, n0)
    p32 = init_params(cfg)
    # biases are zero at init; perturb them deterministically so bias gradients/paths are exercised
    Ws_dims = layer_dims(cfg)
    off = 0
    for l, (fi, fo) in enumerate(Ws_dims, start=1):
        off += fi * fo
        p32[off:off + fo] = np.array([np.float32(0.2 * u01(cfg["seed_weights"] + 100 + l, j) - 0.1) for j in range(fo)],
                                     dtype=np.float32)
        off += fo
    txf, txb, tx0 = (torch.tensor(a, dtype=torch.float64) for a in (xf, xb, x0))
    flat = torch.tensor(p32, dtype=torch.float64, requires_grad=True)
    sums, Y, r = loss_terms(cfg, flat, txf, txb, tx0)
    loss = total_loss(cfg, sums)
    grad = torch.autograd.grad(loss, flat)[0]

    # Adam trajectory (float64), formula of SURVEY.md Appendix B
    p = flat.detach().clone()
    m = torch.zeros_like(p); v = torch.zeros_like(p)
    traj_loss = []
    for t in range(1, adam_steps + 1):
        pr = p.clone().requires_grad_(True)
        s_t, _, _ = loss_terms(cfg, pr, txf, txb, tx0)
        l_t = total_loss(cfg, s_t)
        g_t = torch.autograd.grad(l_t, pr)[0]
        traj_loss.append(float(l_t.detach()))
        m = cfg["beta1"] * m + (1 - cfg["beta1"]) * g_t
        v = cfg["beta2"] * v + (1 - cfg["beta2"]) * g_t * g_t
        mh = m / (1 - cfg["beta1"] ** t); vh = v / (1 - cfg["beta2"] ** t)
        p = p - cfg["lr"] * mh / (torch.sqrt(vh) + cfg["eps"])

    out = dict(cfg_json=np.array(json.dumps(cfg)), params=p32, xf=xf, xb=xb, x0=x0,
               sums=np.array([float(s) for s in sums]), loss=np.array(float(loss)),
               grad=grad.numpy(), Y=Y.detach().numpy(), r=r.detach().numpy(),
               target_b=target(cfg, 1, txb).numpy() if nb else np.zeros((0, d_out)),
               target_0=target(cfg, 2, tx0).numpy() if n0 else np.zeros((0, d_out)),
               traj_loss=np.array(traj_loss), params_after=p.numpy(), adam_m=m.numpy(), adam_v=v.numpy())
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: P={p32.size} loss={float(loss):.6e} |grad|max={float(grad.abs().max()):.3e} -> {os.path.getsize(path)} B")


def make_sampler_fixture():
    """First 48 points of every set for every PDE + the C1 Glorot init head, as raw uint32 bit patterns."""
    out = {}
    for pde in range(5):
        d_in, d_out, _, _ = PDE_SHAPE[pde]
        cfg = dict(pde=pde, d_in=d_in, d_out=d_out, width=20, depth=8, seed_interior=0x5EED0001,
                   seed_boundary=0x5EED0002, seed_initial=0x5EED0003, seed_weights=0x5EED0100)
        for which in range(3):
            out[f"pde{pde}_set{which}"] = sample(cfg, which, 1000, 1048).view(np.uint32)
        if pde == 0:
            out["burgers_20x8_params_head"] = init_params(cfg)[:512].view(np.uint32)
    out["mix_seed1_ctr0_15"] = np.array([mix(1, c) for c in range(16)], dtype=np.uint64)
    np.savez_compressed(os.path.join(HERE, "sampler.npz"), **out)
    print("sampler.npz written")


if __name__ == "__main__":
    torch.set_default_dtype(torch.float64)
    torch.manual_seed(0)
    make_sampler_fixture()
    make_case("burgers_w8_d2", PDE_BURGERS, 8, 2, 96, 40, 40)
    make_case("burgers_w20_d8", PDE_BURGERS, 20, 8, 200, 64, 64, seed_shift=7)       # the C1/C2 network
    make_case("helmholtz_w16_d3", PDE_HELMHOLTZ, 16, 3, 128, 64, 0, pde_param=1.0)
    make_case("allencahn_w16_d2", PDE_ALLEN_CAHN, 16, 2, 100, 48, 48)
    make_case("heat_w8_d3", PDE_HEAT, 8, 3, 70, 33, 31)                             # ragged counts
    make_case("navierstokes_w16_d3", PDE_NAVIER_STOKES, 16, 3, 96, 40, 40)
    make_case("helmholtz_w64_d4", PDE_HELMHOLTZ, 64, 4, 160, 64, 0, adam_steps=2)    # the C3 network
