"""Throughput of the device-resident tensor operations (include/pinn_tensor_abi.h) against their rooflines on one B200:
elementwise / dot are HBM-bound (12 / 8 B per element, peak = MEASURED_PEAKS.json hbm), the bit-exact FP32 matmul is CUDA-core
FMA-bound (peak = profiles/fp32_peak.json).  Host wall clock around K back-to-back asynchronous calls + one synchronising download
(calls are 0.1 - 10 ms each, launch overhead ~5 us).  Usage: tensor_ops_bench.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load, ROOT
P = load()
L = P.lib()
hbm = 6536.0
try:
    hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbps_sustained", hbm))
except Exception:
    pass
fma = float(json.load(open(os.path.join(ROOT, "profiles", "fp32_peak.json"))).get("tflops", 70.8)) if os.path.exists(os.path.join(ROOT, "profiles", "fp32_peak.json")) else 70.8


def sync(arr):
    tmp = np.empty(1, dtype=arr.dtype)
    L.pinn_device_download(tmp.ctypes.data, arr.ptr, tmp.nbytes, None)


out = {}
n = 1 << 27
a = P.DeviceArray(np.ones(n, dtype=np.float32)); b = P.DeviceArray(np.full(n, 2.0, dtype=np.float32)); r = P.DeviceArray(n=n, dtype=np.float32)
for name, op in (("add", P.OP_ADD), ("mul", P.OP_MUL)):
    for _ in range(3):
        L.pinn_tensor_elementwise(op, 0, a.ptr, n, b.ptr, n, r.ptr, n, 1, None)
    sync(r); K = 20; t0 = time.perf_counter()
    for _ in range(K):
        L.pinn_tensor_elementwise(op, 0, a.ptr, n, b.ptr, n, r.ptr, n, 1, None)
    sync(r); dt = (time.perf_counter() - t0) / K
    out[name] = {"elements": n, "ms": dt * 1e3, "GB/s": 12 * n / dt * 1e-9, "frac_of_hbm": 12 * n / dt * 1e-9 / hbm}
small = P.DeviceArray(np.arange(256, dtype=np.float32))
for _ in range(3):
    L.pinn_tensor_elementwise(P.OP_ADD, 0, a.ptr, n, small.ptr, 256, r.ptr, n, 1, None)
sync(r); t0 = time.perf_counter()
for _ in range(20):
    L.pinn_tensor_elementwise(P.OP_ADD, 0, a.ptr, n, small.ptr, 256, r.ptr, n, 1, None)
sync(r); dt = (time.perf_counter() - t0) / 20
out["add_broadcast_bias"] = {"elements": n, "ms": dt * 1e3, "GB/s": 8 * n / dt * 1e-9, "frac_of_hbm": 8 * n / dt * 1e-9 / hbm}
d = P.DeviceArray(n=1, dtype=np.float32)
scratch = P.DeviceArray(n=int(L.pinn_tensor_dot_scratch(n)), dtype=np.float32)
for _ in range(3):
    L.pinn_tensor_dot(0, a.ptr, b.ptr, n, d.ptr, scratch.ptr, None)
sync(d); t0 = time.perf_counter()
for _ in range(20):
    L.pinn_tensor_dot(0, a.ptr, b.ptr, n, d.ptr, scratch.ptr, None)
sync(d); dt = (time.perf_counter() - t0) / 20
out["dot"] = {"elements": n, "ms": dt * 1e3, "GB/s": 8 * n / dt * 1e-9, "frac_of_hbm": 8 * n / dt * 1e-9 / hbm, "value_ok": bool(d.numpy()[0] == 2.0 * n)}
del a, b, r
for I, K_, J in ((4096, 4096, 4096), (25600, 20, 20), (1048576, 64, 64)):
    A = P.DeviceArray(np.random.default_rng(0).standard_normal(I * K_).astype(np.float32))
    B = P.DeviceArray(np.random.default_rng(1).standard_normal(K_ * J).astype(np.float32))
    R = P.DeviceArray(n=I * J, dtype=np.float32)
    for _ in range(2):
        L.pinn_tensor_matmul(0, A.ptr, 0, B.ptr, 0, R.ptr, I, K_, J, None)
    sync(R); reps = 5; t0 = time.perf_counter()
    for _ in range(reps):
        L.pinn_tensor_matmul(0, A.ptr, 0, B.ptr, 0, R.ptr, I, K_, J, None)
    sync(R); dt = (time.perf_counter() - t0) / reps
    fl = 2.0 * I * K_ * J
    by = 4.0 * (I * K_ + K_ * J + I * J)
    out[f"matmul_{I}x{K_}x{J}"] = {"ms": dt * 1e3, "TFLOP/s": fl / dt * 1e-12, "frac_of_fp32_fma": fl / dt * 1e-12 / fma, "GB/s": by / dt * 1e-9, "frac_of_hbm": by / dt * 1e-9 / hbm}
print(json.dumps(out, indent=1))
