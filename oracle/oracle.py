"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module; the product path (cuda-pinn_b200/) never does.  The train-step oracle is
PARITY UNPINNED (the reference has no train step); see oracle/pinn_oracle.h.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


class PinnConfig(C.Structure):
    """Mirror of `pinn_config` (include/pinn_abi.h)."""
    _fields_ = [
        ("abi_version", C.c_int32), ("pde", C.c_int32), ("d_in", C.c_int32), ("d_out", C.c_int32),
        ("width", C.c_int32), ("depth", C.c_int32), ("precision", C.c_int32), ("device", C.c_int32),
        ("rank", C.c_int32), ("nranks", C.c_int32),
        ("n_interior", C.c_uint64), ("n_boundary", C.c_uint64), ("n_initial", C.c_uint64),
        ("seed_interior", C.c_uint64), ("seed_boundary", C.c_uint64), ("seed_initial", C.c_uint64),
        ("seed_weights", C.c_uint64),
        ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
        ("lambda_r", C.c_float), ("lambda_b", C.c_float), ("lambda_0", C.c_float),
        ("pde_param", C.c_float),
    ]

    def copy(self):
        c = PinnConfig()
        C.memmove(C.byref(c), C.byref(self), C.sizeof(PinnConfig))
        return c


PDE_BURGERS, PDE_HELMHOLTZ, PDE_ALLEN_CAHN, PDE_NAVIER_STOKES, PDE_HEAT = 0, 1, 2, 3, 4
SET_INTERIOR, SET_BOUNDARY, SET_INITIAL = 0, 1, 2

_PDE_SHAPE = {  # pde -> (d_in, d_out, has_initial)
    PDE_BURGERS: (2, 1, True), PDE_HELMHOLTZ: (2, 1, False), PDE_ALLEN_CAHN: (3, 1, True),
    PDE_NAVIER_STOKES: (3, 3, True), PDE_HEAT: (3, 1, True),
}


def make_config(pde, width, depth, n_interior, n_boundary=None, n_initial=None, **kw):
    """Builder-defined defaults of SURVEY.md 8(d): N_b = N_0 = max(256, N_f/32), seeds 0x5EED...."""
    d_in, d_out, has_ic = _PDE_SHAPE[pde]
    c = PinnConfig()
    c.abi_version = 1
    c.pde, c.d_in, c.d_out, c.width, c.depth = pde, d_in, d_out, width, depth
    c.precision, c.device, c.rank, c.nranks = 0, 0, 0, 1
    nb = max(256, n_interior // 32) if n_boundary is None else n_boundary
    n0 = (max(256, n_interior // 32) if has_ic else 0) if n_initial is None else n_initial
    c.n_interior, c.n_boundary, c.n_initial = n_interior, nb, n0
    c.seed_interior, c.seed_boundary, c.seed_initial = 0x5EED0001, 0x5EED0002, 0x5EED0003
    c.seed_weights = 0x5EED0100
    c.lr, c.beta1, c.beta2, c.eps = 1e-3, 0.9, 0.999, 1e-8
    c.lambda_r = c.lambda_b = c.lambda_0 = 1.0
    c.pde_param = 0.0
    for k, v in kw.items():
        setattr(c, k, v)
    return c


BASELINE_CONFIGS = {
    1: dict(pde=PDE_BURGERS, width=20, depth=8, n_interior=4096),
    2: dict(pde=PDE_BURGERS, width=20, depth=8, n_interior=25600),
    3: dict(pde=PDE_HELMHOLTZ, width=64, depth=4, n_interior=1 << 20),
    4: dict(pde=PDE_ALLEN_CAHN, width=128, depth=6, n_interior=1 << 24),
    5: dict(pde=PDE_NAVIER_STOKES, width=256, depth=8, n_interior=1 << 26),
}


def baseline_config(which, **kw):
    d = dict(BASELINE_CONFIGS[which])
    d.update(kw)
    return make_config(**d)


def build(force=False):
    """Compile liboracle.so (and oracle/_ref when the reference checkout is present)."""
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
            for f in ("pinn_oracle.c", "pinn_oracle_body.h", "pinn_oracle.h")):
        subprocess.run(["make", "-C", _HERE, "-s", os.path.join(_HERE, "liboracle.so")], check=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        P = C.POINTER
        cfgp = P(PinnConfig)
        L.pinn_oracle_mix.restype = C.c_uint64
        L.pinn_oracle_mix.argtypes = [C.c_uint64, C.c_uint64]
        L.pinn_oracle_u01.restype = C.c_float
        L.pinn_oracle_u01.argtypes = [C.c_uint64, C.c_uint64]
        L.pinn_oracle_sample.argtypes = [cfgp, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p]
        L.pinn_oracle_init_params.argtypes = [cfgp, C.c_void_p]
        L.pinn_oracle_num_params.restype = C.c_size_t
        L.pinn_oracle_num_params.argtypes = [cfgp]
        L.pinn_oracle_num_channels.argtypes = [cfgp]
        L.pinn_oracle_shard_range.argtypes = [C.c_uint64, C.c_int32, C.c_int32, P(C.c_uint64), P(C.c_uint64)]
        for sfx in ("f64", "f32"):
            getattr(L, "pinn_oracle_loss_grad_" + sfx).argtypes = [
                cfgp, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                C.c_void_p, C.c_void_p]
            getattr(L, "pinn_oracle_adam_" + sfx).argtypes = [
                cfgp, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64]
            getattr(L, "pinn_oracle_residual_" + sfx).argtypes = [
                cfgp, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
            f = getattr(L, "pinn_oracle_train_" + sfx)
            f.restype = C.c_double
            f.argtypes = [cfgp, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
        L.pinn_oracle_target_f64.argtypes = [cfgp, C.c_int, C.c_void_p, C.c_void_p]
        L.pinn_oracle_threads.restype = C.c_int
        L.pinn_oracle_set_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def num_params(cfg):
    return int(lib().pinn_oracle_num_params(C.byref(cfg)))


def num_channels(cfg):
    return int(lib().pinn_oracle_num_channels(C.byref(cfg)))


def n_res(cfg):
    return 3 if cfg.pde == PDE_NAVIER_STOKES else 1


def shard_range(n, rank, nranks):
    b, e = C.c_uint64(), C.c_uint64()
    lib().pinn_oracle_shard_range(n, rank, nranks, C.byref(b), C.byref(e))
    return int(b.value), int(e.value)


def set_count(cfg, which):
    return int((cfg.n_interior, cfg.n_boundary, cfg.n_initial)[which])


def sample(cfg, which, begin=0, end=None):
    end = set_count(cfg, which) if end is None else end
    x = np.zeros((max(end - begin, 0), cfg.d_in), dtype=np.float32)
    if end > begin:
        lib().pinn_oracle_sample(C.byref(cfg), which, begin, end, _p(x))
    return x


def init_params(cfg):
    w = np.zeros(num_params(cfg), dtype=np.float32)
    lib().pinn_oracle_init_params(C.byref(cfg), _p(w))
    return w


def loss_grad(cfg, params, xf, xb, x0, dtype=np.float64, want_grad=True):
    """Returns (grad[P] or None, sums[3]) for the given point arrays (means over the GLOBAL
    counts in cfg)."""
    dt = np.dtype(dtype)
    sfx = "f64" if dt == np.float64 else "f32"
    p = np.ascontiguousarray(params, dtype=dt)
    xf = np.ascontiguousarray(xf, dtype=np.float32)
    xb = np.ascontiguousarray(xb, dtype=np.float32)
    x0 = np.ascontiguousarray(x0, dtype=np.float32)
    g = np.zeros_like(p) if want_grad else None
    s = np.zeros(3, dtype=dt)
    getattr(lib(), "pinn_oracle_loss_grad_" + sfx)(
        C.byref(cfg), _p(p), _p(xf), len(xf), _p(xb), len(xb), _p(x0), len(x0), _p(g), _p(s))
    return g, s


def loss_from_sums(cfg, sums):
    mr = sums[0] / cfg.n_interior if cfg.n_interior else 0.0
    mb = sums[1] / cfg.n_boundary if cfg.n_boundary else 0.0
    m0 = sums[2] / cfg.n_initial if cfg.n_initial else 0.0
    return cfg.lambda_r * mr + cfg.lambda_b * mb + cfg.lambda_0 * m0, mr, mb, m0


def adam(cfg, params, m, v, grad, t):
    dt = params.dtype
    sfx = "f64" if dt == np.float64 else "f32"
    getattr(lib(), "pinn_oracle_adam_" + sfx)(C.byref(cfg), _p(params), _p(m), _p(v), _p(grad), params.size, t)


def residual(cfg, params, x, dtype=np.float64):
    dt = np.dtype(dtype)
    sfx = "f64" if dt == np.float64 else "f32"
    p = np.ascontiguousarray(params, dtype=dt)
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.zeros((len(x), cfg.d_out, num_channels(cfg)), dtype=dt)
    r = np.zeros((len(x), n_res(cfg)), dtype=dt)
    getattr(lib(), "pinn_oracle_residual_" + sfx)(C.byref(cfg), _p(p), _p(x), len(x), _p(y), _p(r))
    return y, r


def target(cfg, which, x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    g = np.zeros((len(x), cfg.d_out), dtype=np.float64)
    for i in range(len(x)):
        lib().pinn_oracle_target_f64(C.byref(cfg), which, _p(x[i]), _p(g[i]))
    return g


def train(cfg, params, steps, dtype=np.float32, m=None, v=None, t0=0):
    """Run `steps` Adam steps on the oracle's own sampled global point sets.
    Returns (params, m, v, losses[steps,4], seconds)."""
    dt = np.dtype(dtype)
    sfx = "f64" if dt == np.float64 else "f32"
    p = np.array(params, dtype=dt)
    m = np.zeros_like(p) if m is None else np.array(m, dtype=dt)
    v = np.zeros_like(p) if v is None else np.array(v, dtype=dt)
    losses = np.zeros((steps, 4), dtype=dt)
    sec = getattr(lib(), "pinn_oracle_train_" + sfx)(C.byref(cfg), _p(p), _p(m), _p(v), t0, steps, _p(losses))
    return p, m, v, losses, float(sec)


def threads():
    return int(lib().pinn_oracle_threads())


def ref_dir():
    return os.path.join(_HERE, "_ref")
