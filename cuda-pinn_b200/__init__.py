"""cuda-pinn_b200 — host-side mirror of the reference's intended `network` interface over the C ABI.

The product is the CUDA library (csrc/ → libpinn_b200.so, entry points in include/pinn_abi.h) and
the C++ `network<T>` header (include/network.cuh).  This module is the thin ctypes binding the
tests and bench.py drive it through; it contains no arithmetic and has NO fallback: if the
shared library is missing or there is no B200, construction raises.

(The directory name has a hyphen, so import it with importlib — see `load()` in
`__graft_entry__.py`, which registers it as `cuda_pinn_b200`.)
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PINN_B200_LIB") or os.path.join(_HERE, "libpinn_b200.so")   # (the override is for A/B measurements of two builds on one box)

PDE_BURGERS, PDE_HELMHOLTZ, PDE_ALLEN_CAHN, PDE_NAVIER_STOKES, PDE_HEAT = 0, 1, 2, 3, 4
SET_INTERIOR, SET_BOUNDARY, SET_INITIAL = 0, 1, 2
PREC_FP32, PREC_BF16X3, PREC_BF16, PREC_BF16X6, PREC_FP16X3 = 0, 1, 2, 3, 4
ABI_VERSION = 1


class PinnConfig(C.Structure):
    """`pinn_config` of include/pinn_abi.h."""
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


class PinnStepOut(C.Structure):
    _fields_ = [("loss", C.c_float), ("mse_residual", C.c_float), ("mse_boundary", C.c_float),
                ("mse_initial", C.c_float), ("step", C.c_uint64)]


class PinnEvalOut(C.Structure):
    _fields_ = [("n", C.c_uint64), ("rel_l2", C.c_double * 3), ("max_abs", C.c_double * 3), ("mse_residual", C.c_double)]


class PinnProfile(C.Structure):
    _fields_ = [("launches", C.c_uint64), ("step_kernel_ms", C.c_double), ("step_kernel_count", C.c_uint64),
                ("grid", C.c_int32), ("block", C.c_int32), ("smem_bytes", C.c_int32), ("kernel_id", C.c_int32)]


# every symbol include/pinn_abi.h declares (tests check the library exports exactly these)
ABI_SYMBOLS = [
    "pinn_config_default", "pinn_num_channels", "pinn_num_params", "pinn_shard_range",
    "pinn_create", "pinn_destroy", "pinn_last_error",
    "pinn_set_params", "pinn_get_params", "pinn_get_grads", "pinn_get_optimizer", "pinn_set_optimizer",
    "pinn_sample_points", "pinn_resample_points", "pinn_set_points", "pinn_set_points_device", "pinn_set_points_view",
    "pinn_h2d_point_bytes", "pinn_get_points", "pinn_local_count",
    "pinn_train_step", "pinn_train_step_host", "pinn_train_step_device", "pinn_train_steps", "pinn_loss_grad",
    "pinn_forward", "pinn_residual", "pinn_synchronize",
    "pinn_save", "pinn_load", "pinn_exact_solution", "pinn_evaluate",
    "pinn_comm_unique_id", "pinn_comm_init", "pinn_comm_p2p_export", "pinn_comm_p2p_open", "pinn_step_local", "pinn_step_apply",
    "pinn_grad_buffer", "pinn_stream", "pinn_graph_state", "pinn_profile_enable", "pinn_profile_get",
]

_lib = None


class PinnError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"[pinn status {code}] {message}")
        self.code = code


def lib():
    """Load libpinn_b200.so; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python __graft_entry__.py build` (nvcc, sm_100a). "
                          "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    P = C.POINTER
    vp, cfgp = C.c_void_p, P(PinnConfig)
    L.pinn_config_default.argtypes = [cfgp, C.c_int]
    L.pinn_num_channels.argtypes = [cfgp]
    L.pinn_num_params.argtypes = [cfgp]; L.pinn_num_params.restype = C.c_size_t
    L.pinn_shard_range.argtypes = [C.c_uint64, C.c_int32, C.c_int32, P(C.c_uint64), P(C.c_uint64)]
    L.pinn_shard_range.restype = None
    # include/pinn_tensor_abi.h: device-resident tensor operations
    sz = C.c_size_t
    L.pinn_tensor_elementwise.argtypes = [C.c_int, C.c_int, vp, sz, vp, sz, vp, sz, C.c_int, vp]
    L.pinn_tensor_matmul.argtypes = [C.c_int, vp, C.c_int, vp, C.c_int, vp, sz, sz, sz, vp]
    L.pinn_tensor_dot_scratch.argtypes = [sz]; L.pinn_tensor_dot_scratch.restype = sz
    L.pinn_tensor_dot.argtypes = [C.c_int, vp, vp, sz, vp, vp, vp]
    L.pinn_tensor_add_multiple.argtypes = [C.c_int, P(vp), sz, sz, vp, vp]
    L.pinn_tensor_broadcast.argtypes = [C.c_int, C.c_int, vp, P(sz), C.c_int, vp, P(sz), C.c_int, vp, P(sz), C.c_int, vp]
    L.pinn_tensor_einsum.argtypes = [C.c_int, C.c_char_p, vp, P(sz), C.c_int, vp, P(sz), C.c_int, vp, vp]
    L.pinn_tensor_normalize_columns.argtypes = [vp, sz, sz, P(C.c_float), P(C.c_float), P(C.c_float), vp]
    L.pinn_device_alloc.argtypes = [P(vp), sz, vp]
    L.pinn_device_free.argtypes = [vp, vp]
    L.pinn_device_upload.argtypes = [vp, vp, sz, vp]
    L.pinn_device_download.argtypes = [vp, vp, sz, vp]
    L.pinn_graph_state.argtypes = [vp, P(C.c_char_p)]
    L.pinn_create.argtypes = [P(vp), cfgp]
    L.pinn_destroy.argtypes = [vp]; L.pinn_destroy.restype = None
    L.pinn_last_error.argtypes = [vp]; L.pinn_last_error.restype = C.c_char_p
    L.pinn_set_params.argtypes = [vp, vp, C.c_size_t, C.c_int]
    L.pinn_get_params.argtypes = [vp, vp, C.c_size_t]
    L.pinn_get_grads.argtypes = [vp, vp, C.c_size_t]
    L.pinn_get_optimizer.argtypes = [vp, vp, vp, C.c_size_t, P(C.c_uint64)]
    L.pinn_set_optimizer.argtypes = [vp, vp, vp, C.c_size_t, C.c_uint64]
    L.pinn_sample_points.argtypes = [vp]
    L.pinn_resample_points.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint64]
    L.pinn_set_points.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.pinn_set_points_device.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.pinn_set_points_view.argtypes = [vp, C.c_int, vp]
    L.pinn_h2d_point_bytes.argtypes = [vp]; L.pinn_h2d_point_bytes.restype = C.c_ulonglong
    L.pinn_train_step_device.argtypes = [vp, vp, C.c_size_t, P(PinnStepOut)]
    L.pinn_get_points.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.pinn_local_count.argtypes = [vp, C.c_int]; L.pinn_local_count.restype = C.c_size_t
    L.pinn_train_step.argtypes = [vp, P(PinnStepOut)]
    L.pinn_train_step_host.argtypes = [vp, vp, C.c_size_t, P(PinnStepOut)]
    L.pinn_train_steps.argtypes = [vp, C.c_int, P(PinnStepOut)]
    L.pinn_loss_grad.argtypes = [vp, P(PinnStepOut)]
    L.pinn_forward.argtypes = [vp, vp, vp, C.c_size_t]
    L.pinn_residual.argtypes = [vp, vp, vp, vp, C.c_size_t]
    L.pinn_synchronize.argtypes = [vp]
    L.pinn_save.argtypes = [vp, C.c_char_p]
    L.pinn_load.argtypes = [vp, C.c_char_p]
    L.pinn_exact_solution.argtypes = [cfgp, vp, vp, C.c_size_t]
    L.pinn_evaluate.argtypes = [vp, C.c_int, P(PinnEvalOut)]
    L.pinn_comm_unique_id.argtypes = [vp]
    L.pinn_comm_init.argtypes = [vp, vp]
    L.pinn_comm_p2p_export.argtypes = [vp, vp]
    L.pinn_comm_p2p_open.argtypes = [vp, vp, C.c_int]
    L.pinn_step_local.argtypes = [vp]
    L.pinn_step_apply.argtypes = [vp, P(PinnStepOut)]
    L.pinn_grad_buffer.argtypes = [vp, P(C.c_size_t)]; L.pinn_grad_buffer.restype = vp
    L.pinn_stream.argtypes = [vp]; L.pinn_stream.restype = vp
    L.pinn_profile_enable.argtypes = [vp, C.c_int]
    L.pinn_profile_get.argtypes = [vp, P(PinnProfile), C.c_int]
    _lib = L
    return L


def default_config(which, **overrides):
    """BASELINE config C1..C5 (pinn_config_default) with field overrides."""
    cfg = PinnConfig()
    rc = lib().pinn_config_default(C.byref(cfg), which)
    if rc:
        raise PinnError(rc, "pinn_config_default: invalid config index")
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg


def shard_range(n, rank, nranks):
    b, e = C.c_uint64(), C.c_uint64()
    lib().pinn_shard_range(n, rank, nranks, C.byref(b), C.byref(e))
    return int(b.value), int(e.value)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def exact_solution(cfg, x):
    """Analytic solution u*(x) [n, d_out] where the PDE has one (pinn_exact_solution); PinnError(7) otherwise."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = np.empty((len(x), cfg.d_out), dtype=np.float32)
    rc = lib().pinn_exact_solution(C.byref(cfg), _ptr(x), _ptr(u), len(x))
    if rc:
        raise PinnError(rc, "pinn_exact_solution: no analytic solution for this pde" if rc == 7 else "pinn_exact_solution: invalid argument")
    return u


class Network:
    """Handle on one GPU's PINN (mirrors the `network` class of include/network.cuh)."""

    def __init__(self, cfg):
        self._L = lib()
        self.cfg = cfg
        self._h = C.c_void_p()
        rc = self._L.pinn_create(C.byref(self._h), C.byref(cfg))
        if rc:
            raise PinnError(rc, self._L.pinn_last_error(None).decode())
        self.num_params = int(self._L.pinn_num_params(C.byref(cfg)))
        self.num_channels = int(self._L.pinn_num_channels(C.byref(cfg)))
        self.n_res = 3 if cfg.pde == PDE_NAVIER_STOKES else 1

    # -- lifetime
    def close(self):
        if self._h:
            self._L.pinn_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise PinnError(rc, self._L.pinn_last_error(self._h).decode())

    # -- parameters
    def get_params(self):
        w = np.empty(self.num_params, dtype=np.float32)
        self._check(self._L.pinn_get_params(self._h, _ptr(w), w.size))
        return w

    def set_params(self, w, reset_optimizer=True):
        w = np.ascontiguousarray(w, dtype=np.float32)
        self._check(self._L.pinn_set_params(self._h, _ptr(w), w.size, int(reset_optimizer)))

    def get_grads(self):
        g = np.empty(self.num_params, dtype=np.float32)
        self._check(self._L.pinn_get_grads(self._h, _ptr(g), g.size))
        return g

    def get_optimizer(self):
        m = np.empty(self.num_params, dtype=np.float32)
        v = np.empty(self.num_params, dtype=np.float32)
        t = C.c_uint64()
        self._check(self._L.pinn_get_optimizer(self._h, _ptr(m), _ptr(v), m.size, C.byref(t)))
        return m, v, int(t.value)

    def set_optimizer(self, m, v, step):
        m = np.ascontiguousarray(m, dtype=np.float32); v = np.ascontiguousarray(v, dtype=np.float32)
        self._check(self._L.pinn_set_optimizer(self._h, _ptr(m), _ptr(v), m.size, step))

    # -- points
    def sample_points(self):
        self._check(self._L.pinn_sample_points(self._h))

    def resample_points(self, si, sb, s0):
        self._check(self._L.pinn_resample_points(self._h, si, sb, s0))

    def local_count(self, which):
        return int(self._L.pinn_local_count(self._h, which))

    def get_points(self, which):
        n = self.local_count(which)
        x = np.empty((n, self.cfg.d_in), dtype=np.float32)
        self._check(self._L.pinn_get_points(self._h, which, _ptr(x), n))
        return x

    def set_points(self, which, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        self._check(self._L.pinn_set_points(self._h, which, _ptr(x), len(x)))

    def set_points_device(self, which, d_ptr, n):
        """d_ptr: integer DEVICE address of [n, d_in] float32 points on this handle's GPU (device-to-device, no H2D)."""
        self._check(self._L.pinn_set_points_device(self._h, which, C.c_void_p(d_ptr), n))

    def set_points_view(self, which, d_ptr, n_floats, transposed=False):
        """Builds the reference's 48-byte kernel-argument view device::tensor::variables<float> around a device buffer
        (variables.cuh:6-52: data @0, shape @8, stride @16, dim @24, n @32, transposed @40) and hands it over."""
        import struct
        v = struct.pack("<QQQQQ?7x", d_ptr, 0, 0, 2, n_floats, bool(transposed))
        assert len(v) == 48
        buf = C.create_string_buffer(v, 48)
        self._check(self._L.pinn_set_points_view(self._h, which, buf))

    def h2d_point_bytes(self):
        return int(self._L.pinn_h2d_point_bytes(self._h))

    def train_step_device(self, d_ptr, n):
        o = PinnStepOut()
        self._check(self._L.pinn_train_step_device(self._h, C.c_void_p(d_ptr), n, C.byref(o)))
        return self._out(o)

    # -- hot path
    @staticmethod
    def _out(o):
        return dict(loss=o.loss, mse_residual=o.mse_residual, mse_boundary=o.mse_boundary,
                    mse_initial=o.mse_initial, step=int(o.step))

    def train_step(self, read_loss=True):
        o = PinnStepOut()
        self._check(self._L.pinn_train_step(self._h, C.byref(o) if read_loss else None))
        return self._out(o) if read_loss else None

    def train_step_host(self, x_ptr, n):
        """x_ptr: integer address (or numpy array) of HOST points [n, d_in] float32."""
        o = PinnStepOut()
        p = _ptr(x_ptr) if isinstance(x_ptr, np.ndarray) else C.c_void_p(x_ptr)
        self._check(self._L.pinn_train_step_host(self._h, p, n, C.byref(o)))
        return self._out(o)

    def train_steps(self, k, read_loss=True):
        o = PinnStepOut()
        self._check(self._L.pinn_train_steps(self._h, k, C.byref(o) if read_loss else None))
        return self._out(o) if read_loss else None

    def loss_grad(self):
        o = PinnStepOut()
        self._check(self._L.pinn_loss_grad(self._h, C.byref(o)))
        return self._out(o)

    def step_local(self):
        self._check(self._L.pinn_step_local(self._h))

    def step_apply(self, read_loss=True):
        o = PinnStepOut()
        self._check(self._L.pinn_step_apply(self._h, C.byref(o) if read_loss else None))
        return self._out(o) if read_loss else None

    def forward(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.empty((len(x), self.cfg.d_out), dtype=np.float32)
        self._check(self._L.pinn_forward(self._h, _ptr(x), _ptr(y), len(x)))
        return y

    def residual(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.empty((len(x), self.cfg.d_out, self.num_channels), dtype=np.float32)
        r = np.empty((len(x), self.n_res), dtype=np.float32)
        self._check(self._L.pinn_residual(self._h, _ptr(x), _ptr(y), _ptr(r), len(x)))
        return y, r

    def synchronize(self):
        self._check(self._L.pinn_synchronize(self._h))

    # -- checkpoint + evaluation
    def save(self, path):
        self._check(self._L.pinn_save(self._h, os.fsencode(path)))

    def load(self, path):
        self._check(self._L.pinn_load(self._h, os.fsencode(path)))

    def evaluate(self, per_axis):
        """Grid evaluation: relative L2 / max error vs the analytic solution (NaN if none) and residual MSE."""
        o = PinnEvalOut()
        self._check(self._L.pinn_evaluate(self._h, per_axis, C.byref(o)))
        k = self.cfg.d_out
        return dict(n=int(o.n), rel_l2=list(o.rel_l2)[:k], max_abs=list(o.max_abs)[:k], mse_residual=o.mse_residual)

    # -- multi-GPU
    @staticmethod
    def comm_unique_id():
        buf = (C.c_char * 128)()
        rc = lib().pinn_comm_unique_id(buf)
        if rc:
            raise PinnError(rc, lib().pinn_last_error(None).decode())
        return bytes(buf)

    def comm_init(self, id_bytes):
        buf = (C.c_char * 128).from_buffer_copy(id_bytes)
        self._check(self._L.pinn_comm_init(self._h, buf))

    def p2p_export(self):
        """IPC handle (64 bytes) of this rank's exchange buffer for the fused peer-memory all-reduce."""
        buf = (C.c_char * 64)()
        self._check(self._L.pinn_comm_p2p_export(self._h, buf))
        return bytes(buf)

    def p2p_open(self, handles):
        """handles: list of the 64-byte IPC handles of ALL ranks in rank order."""
        blob = b"".join(handles)
        buf = (C.c_char * len(blob)).from_buffer_copy(blob)
        self._check(self._L.pinn_comm_p2p_open(self._h, buf, len(handles)))

    def grad_buffer(self):
        n = C.c_size_t()
        p = self._L.pinn_grad_buffer(self._h, C.byref(n))
        return int(p), int(n.value)

    def stream(self):
        return int(self._L.pinn_stream(self._h) or 0)

    # -- measurement
    def profile_enable(self, on=True):
        self._check(self._L.pinn_profile_enable(self._h, int(on)))

    def graph_state(self):
        """(replaying a CUDA graph?, note): the note says why not when capture failed (never silent)."""
        note = C.c_char_p()
        st = self._L.pinn_graph_state(self._h, C.byref(note))
        return st == 1, (note.value or b"").decode()

    def profile_get(self, reset=False):
        p = PinnProfile()
        self._check(self._L.pinn_profile_get(self._h, C.byref(p), int(reset)))
        return dict(launches=int(p.launches), step_kernel_ms=p.step_kernel_ms, step_kernel_count=int(p.step_kernel_count),
                    grid=p.grid, block=p.block, smem_bytes=p.smem_bytes, kernel_id=p.kernel_id)


# ---------------------------------------------------------------------------------------------
# include/pinn_tensor_abi.h — device-resident tensor operations (pass-through; no arithmetic here)
# ---------------------------------------------------------------------------------------------
OP_ADD, OP_SUB, OP_MUL, OP_DIV = 0, 1, 2, 3


class DeviceArray:
    """A flat device buffer of float32 / float64 owned through pinn_device_alloc / pinn_device_free."""

    def __init__(self, data=None, n=None, dtype=np.float32):
        self._L = lib()
        if data is not None:
            data = np.ascontiguousarray(data)
            dtype, n = data.dtype, data.size
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise ValueError("float32 or float64")
        self.n = int(n)
        self.ptr = C.c_void_p()
        rc = self._L.pinn_device_alloc(C.byref(self.ptr), self.n * self.dtype.itemsize, None)
        if rc:
            raise PinnError(rc, "pinn_device_alloc failed")
        if data is not None:
            rc = self._L.pinn_device_upload(self.ptr, data.ctypes.data_as(C.c_void_p), data.nbytes, None)
            if rc:
                raise PinnError(rc, "pinn_device_upload failed")

    @property
    def code(self):
        return 0 if self.dtype == np.dtype(np.float32) else 1

    def numpy(self):
        out = np.empty(self.n, dtype=self.dtype)
        rc = self._L.pinn_device_download(out.ctypes.data_as(C.c_void_p), self.ptr, out.nbytes, None)
        if rc:
            raise PinnError(rc, "pinn_device_download failed")
        return out

    def free(self):
        if self.ptr:
            self._L.pinn_device_free(self.ptr, None)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def tensor_elementwise(op, a, b, exact_bounds=True):
    """r = a (op) b with the reference's flat-index broadcast; returns (status, DeviceArray)."""
    r = DeviceArray(n=max(a.n, b.n), dtype=a.dtype)
    rc = lib().pinn_tensor_elementwise(op, a.code, a.ptr, a.n, b.ptr, b.n, r.ptr, r.n, 1 if exact_bounds else 0, None)
    return rc, r


def tensor_matmul(a, b, I, K, J, a_transposed=False, b_transposed=False):
    r = DeviceArray(n=I * J, dtype=a.dtype)
    rc = lib().pinn_tensor_matmul(a.code, a.ptr, int(a_transposed), b.ptr, int(b_transposed), r.ptr, I, K, J, None)
    if rc:
        raise PinnError(rc, "pinn_tensor_matmul failed")
    return r


def tensor_dot(a, b):
    r = DeviceArray(n=1, dtype=a.dtype)
    rc = lib().pinn_tensor_dot(a.code, a.ptr, b.ptr, a.n, r.ptr, None, None)
    if rc:
        raise PinnError(rc, "pinn_tensor_dot failed")
    return r


def tensor_add_multiple(srcs):
    r = DeviceArray(n=srcs[0].n, dtype=srcs[0].dtype)
    arr = (C.c_void_p * len(srcs))(*[s.ptr for s in srcs])
    rc = lib().pinn_tensor_add_multiple(srcs[0].code, arr, len(srcs), srcs[0].n, r.ptr, None)
    if rc:
        raise PinnError(rc, "pinn_tensor_add_multiple failed")
    return r


def _shape_arr(shape):
    return (C.c_size_t * max(len(shape), 1))(*shape)


def tensor_broadcast(op, a, a_shape, b, b_shape):
    """NumPy-rule broadcasting (pinn_tensor_broadcast); returns (DeviceArray, result shape)."""
    r_shape = tuple(np.broadcast_shapes(tuple(a_shape), tuple(b_shape)))
    r = DeviceArray(n=int(np.prod(r_shape)), dtype=a.dtype)
    rc = lib().pinn_tensor_broadcast(op, a.code, a.ptr, _shape_arr(a_shape), len(a_shape), b.ptr, _shape_arr(b_shape), len(b_shape),
                                     r.ptr, _shape_arr(r_shape), len(r_shape), None)
    if rc:
        raise PinnError(rc, "pinn_tensor_broadcast failed")
    return r, r_shape


def tensor_einsum(spec, a, a_shape, b, b_shape, out_size):
    r = DeviceArray(n=max(int(out_size), 1), dtype=a.dtype)
    rc = lib().pinn_tensor_einsum(a.code, spec.encode(), a.ptr, _shape_arr(a_shape), len(a_shape), b.ptr, _shape_arr(b_shape), len(b_shape), r.ptr, None)
    return rc, r


def tensor_normalize_columns(x, n, d, lo=None, hi=None):
    """In place; returns the (lo, hi) bounds used."""
    stats = (C.c_float * (2 * d))()
    plo = (C.c_float * d)(*lo) if lo is not None else None
    phi = (C.c_float * d)(*hi) if hi is not None else None
    rc = lib().pinn_tensor_normalize_columns(x.ptr, n, d, plo, phi, stats, None)
    if rc:
        raise PinnError(rc, "pinn_tensor_normalize_columns failed")
    return np.array(stats[:d], dtype=np.float32), np.array(stats[d:], dtype=np.float32)
