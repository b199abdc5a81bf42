#!/usr/bin/env python
"""bench.py — collocation points/s per PINN train step (fwd + residual + bwd + Adam) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1..5] [--points n] [--impl reference]

Default workload (every N): the per-GPU share of BASELINE config 5 — Navier-Stokes 3->[256x8]->3, 8,388,608 interior
points per GPU (64M on 8 GPUs), the configuration the north star's roofline and weak-scaling targets are stated on.  At
N = 1 the same JSON line carries `also: {c2: ..., c3: ...}`: BASELINE config 2 (the latency-bound Burgers net) and config 3
(1M-point Helmholtz, the largest single-GPU configuration), each with its own value / roofline / e2e.

One process per GPU (torchrun for N>1, NCCL); a "step" is one pass of the hot path over this rank's
shard of the collocation set.  Prints ONE JSON line on rank 0 (contract in the task statement):
`value` = whole-job points/s with inputs resident in HBM, `e2e` = the same through the public
C-ABI call with HOST point buffers (pinned H2D of the step's points + D2H of the loss inside the
timed region), `roofline` for the fused kernel from CUDA-event launch times measured live,
`cpu_baseline` = the CPU oracle (port; the reference has no CPU path) on a bounded sample.

--impl reference: the reference has no implementation of this path, so the arm
times the oracle port (FP32, OpenMP, all host threads) on a bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIG_NAMES = {
    1: "C1 builder-default Burgers 2->[20x8]->1, 4,096 pts",
    2: "C2 1-D viscous Burgers 2->[20x8]->1, 25,600 pts",
    3: "C3 2-D Helmholtz 2->[64x4]->1, 1,048,576 pts",
    4: "C4 2-D Allen-Cahn 3->[128x6]->1, 16,777,216 pts",
    5: "C5 2-D Navier-Stokes 3->[256x8]->3, 67,108,864 pts",
}
FP32_PEAK_DERIVED = 148 * 128 * 2 * 1.965e9 / 1e12    # CUDA-core FMA peak at max SM clock (derived)


def fp32_peak():
    """Measured FP32 FFMA peak of this pool's B200 (tests/cuda/fma_peak.cu → profiles/fp32_peak.json);
    MEASURED_PEAKS.json has no FP32 entry, so this repo measured it itself.  Falls back to the derived figure."""
    try:
        with open(os.path.join(ROOT, "profiles", "fp32_peak.json")) as f:
            return float(json.load(f)["fp32_ffma_tflops"]), "measured (tests/cuda/fma_peak.cu, profiles/fp32_peak.json): register-only FFMA, 64 warps/SM"
    except Exception:
        return FP32_PEAK_DERIVED, "derived 148 SM x 128 FMA x 2 x 1.965 GHz"


def traffic_for(which, prec_name):
    """dram__bytes_read.sum + dram__bytes_write.sum of the fused kernel per launch from the committed ncu capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(f"c{which}_{prec_name}")
    except Exception:
        return None


def mma_sync_line(prec_name, points, k_ms, depth):
    """Issued tensor work of the width-20 mma.sync kernel against the measured mma.sync BF16 peak (profiles/mma_sync_peak.json)."""
    pairs = 6 if prec_name == "bf16x6" else 3
    # per 16-point warp tile and hidden->hidden layer: forward and dA = C x 3 n-tiles x pairs x (m16n8k16 + m16n8k8),
    # dW = C x 3 n-tiles x pairs x 2 m-tiles x m16n8k16
    per_tile_layer = 4 * 3 * pairs * (2 * (4096 + 2048) + 2 * 4096)
    issued = points / 16.0 * (depth - 1) * per_tile_layer
    peak = 549.8
    try:
        with open(os.path.join(ROOT, "profiles", "mma_sync_peak.json")) as f:
            peak = float(json.load(f)["bf16_m16n8k16_tflops"])
    except Exception:
        pass
    tf = issued / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
    return {"issued_tflops": tf, "peak": peak, "frac": tf / peak, "of": "measured mma.sync m16n8k16 BF16 peak, tests/cuda/mma_sync_peak.cu"}


def traffic_bytes(which, prec_name, points):
    t = traffic_for(which, prec_name)
    return None if not t else t["bytes_per_point"] * points


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0, "bf16_tflops": 1590.0, "source": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        p.update({k: m[k] for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained") if k in m})
        p["source"] = "measured"
    except Exception:
        pass
    return p


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe), as a SUBPROCESS polling every
    20 ms (in-process NVML polling from a thread stalled long power-capped runs: C5 at full size never finished).  It is
    started before the warm-up (nvidia-smi needs ~0.2 s to produce its first row); rows are stamped on arrival and only
    those inside the load window [mark_begin, mark_end] are used."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t_begin, self.t_end = None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lo = self.t_begin if self.t_begin is not None else 0.0
        hi = (self.t_end if self.t_end is not None else float("inf")) + 0.02      # a row describes the 20 ms before it
        for ts, r in self.rows:
            if not (lo <= ts <= hi):
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def global_counts(which, gpus, points):
    base = {1: 4096, 2: 25600, 3: 1 << 20, 4: 1 << 24, 5: 1 << 26}[which]
    if points:
        per_gpu = points
    elif which in (4, 5):
        per_gpu = base // 8          # the 16M / 64M-point configs are stated for 8 GPUs: 2M / 8M per GPU
    else:
        per_gpu = base
    nf = per_gpu * gpus
    nb = max(256, nf // 32)
    n0 = 0 if which == 3 else nb
    return nf, nb, n0, per_gpu


def algorithmic(which):
    """FLOPs/pt = 6*C*W (matmul only), HBM bytes/pt = 4*d_in   (BASELINE.md section 3)."""
    C = {1: 4, 2: 4, 3: 5, 4: 6, 5: 6}[which]
    W = {1: 2860, 2: 2860, 3: 12480, 4: 82432, 5: 460288}[which]
    d_in = {1: 2, 2: 2, 3: 2, 4: 3, 5: 3}[which]
    return 6 * C * W, 4 * d_in


def oracle_all_cores():
    """The oracle with every host core: torchrun exports OMP_NUM_THREADS=1 to its workers, which would otherwise cap the port
    at one thread (the N > 1 reference lines ran on 1 core)."""
    from oracle import oracle as O
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    O.lib().pinn_oracle_set_threads(int(n))
    return O, O.threads()


def reference_sample_points(which, cores):
    """Bounded sample of the per-GPU workload, the same at every N: ~0.6 s of CPU work per step at ~1.5 GFLOP/s/core."""
    flops_pt, _ = algorithmic(which)
    _, _, _, per_gpu = global_counts(which, 1, 0)
    return int(max(256, min(per_gpu, 0.6 * cores * 1.5e9 / flops_pt)))


def run_reference(args):
    """Reference arm: the oracle port (FP32, OpenMP over all host threads) on a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    O, cores = oracle_all_cores()
    which = args.config
    _, _, _, per_gpu = global_counts(which, args.gpus, args.points)
    budget_pts = reference_sample_points(which, cores)
    cfg = O.baseline_config(which, n_interior=budget_pts, n_boundary=max(256, budget_pts // 32),
                            n_initial=0 if which == 3 else max(256, budget_pts // 32))
    w = O.init_params(cfg)
    _, m, v, _, _ = O.train(cfg, w, max(1, args.warmup), dtype="float32")
    t0 = time.perf_counter()
    _, _, _, losses, sec = O.train(cfg, w, args.steps, dtype="float32")
    wall = time.perf_counter() - t0
    pts_s = budget_pts * args.steps / sec
    line = {
        "impl": "reference", "metric": "collocation points/sec per train step (fwd+residual+bwd+Adam)",
        "value": pts_s, "unit": "points/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": CONFIG_NAMES[which], "points_per_gpu": per_gpu},
        "cpu_baseline": {"value": pts_s, "unit": "points/s", "cores": cores, "kind": "port",
                         "sample": f"{budget_pts} interior points x {args.steps} steps (oracle FP32, OpenMP, {cores} threads; the same "
                                   "sample at every N); the reference has no implementation of this path (SURVEY.md F1)"},
        "e2e": {"value": pts_s, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": wall, "final_loss": float(losses[-1, 0]),
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(which, seconds=12.0):
    O, cores = oracle_all_cores()
    flops_pt, _ = algorithmic(which)
    pts = int(max(256, min({1: 4096, 2: 25600}.get(which, 1 << 30), 0.5 * cores * 1.5e9 / flops_pt)))
    cfg = O.baseline_config(which, n_interior=pts, n_boundary=max(256, pts // 32),
                            n_initial=0 if which == 3 else max(256, pts // 32))
    w = O.init_params(cfg)
    _, _, _, _, sec1 = O.train(cfg, w, 1, dtype="float32")
    steps = int(max(2, min(200, seconds / max(sec1, 1e-4))))
    _, _, _, _, sec = O.train(cfg, w, steps, dtype="float32")
    return {"value": pts * steps / sec, "unit": "points/s", "cores": cores, "kind": "port",
            "sample": f"{pts} interior points x {steps} steps, oracle FP32 + OpenMP ({sec:.1f} s)"}


def reference_cuda_composed(which):
    """B-ref-cuda: the reference's own tensor<float> ops composing the interior part of the step
    op by op (+ labelled host shim for tanh/sqrt), oracle/_ref/ref_step built from /root/reference.  Burgers nets only."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_step")
    if which not in (1, 2) or not os.path.exists(exe):
        return None
    out = {}
    for name, path, n, steps in (("O3", exe, 4096, 2), ("as_shipped_G", exe + "_dbg", 2048, 1)):
        if not os.path.exists(path):
            continue
        try:
            r = subprocess.run([path, str(n), "20", "8", str(steps)], capture_output=True, text=True, timeout=240)
            kv = dict(l.split("=", 1) for l in r.stdout.splitlines() if "=" in l and not l.startswith("steps"))
            out[name] = {"value": float(kv["points_per_second"]), "unit": "points/s",
                         "sample": f"{n} interior points x {steps} steps (chunked: the reference mirrors every tensor on the host)"}
        except Exception as e:   # noqa: BLE001
            out[name] = {"error": str(e)[:200]}
    out["what"] = "reference tensor<float> ops + host shim for tanh/sqrt; NOT the reference's train step (it has none)"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=int(os.environ.get("PINN_BENCH_CONFIG", "5")), choices=[1, 2, 3, 4, 5])
    ap.add_argument("--points", type=int, default=0, help="interior points per GPU (default: the config's)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("PINN_BENCH_PRECISION", "auto"),
                    choices=["auto", "fp32", "bf16x3", "bf16", "bf16x6", "fp16x3"],
                    help="contraction arithmetic: fp32 CUDA cores | fp16x3 / bf16x6 mma.sync (width 20) | bf16x3 / bf16 tcgen05 "
                         "(width 64/128/256); auto = fp16x3 for the width-20 Burgers net, bf16x3 for width 64, bf16 for 128 / 256")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the C2 / C3 sub-records of the default N = 1 run")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    from __graft_entry__ import load
    P = load()

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if rank == 0:
            print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch with torchrun", file=sys.stderr)
        sys.exit(2)
    if not torch.cuda.is_available():
        print("bench.py: no CUDA device; the product path has no CPU fallback", file=sys.stderr)
        sys.exit(3)
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def run_config(which, steps, warmup, points, with_cpu):
        """One bench record: BASELINE config `which` on this job's ranks."""
        nf, nb, n0, per_gpu = global_counts(which, world, points)
        width = {1: 20, 2: 20, 3: 64, 4: 128, 5: 256}[which]
        prec_name = args.precision
        if prec_name == "auto":
            # width 20 Burgers (C1, C2): FP32-grade scaled 2-term FP16 split on mma.sync (meets the FP32 path's bars); other narrow nets: FP32 CUDA cores
            prec_name = ("fp16x3" if which in (1, 2) and width == 20 else "fp32") if width < 64 else ("bf16x3" if width < 128 else "bf16")
        prec = {"fp32": P.PREC_FP32, "bf16x3": P.PREC_BF16X3, "bf16": P.PREC_BF16, "bf16x6": P.PREC_BF16X6, "fp16x3": P.PREC_FP16X3}[prec_name]
        cfg = P.default_config(which, n_interior=nf, n_boundary=nb, n_initial=n0, rank=rank, nranks=world,
                               device=local_rank, precision=prec)
        net = P.Network(cfg)
        if world > 1:
            box = [P.Network.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            net.comm_init(box[0])
            collective = "nccl all-reduce of [P+3] floats"
            if os.environ.get("PINN_P2P", "0") == "1":
                hs = [None] * world
                dist.all_gather_object(hs, net.p2p_export())
                net.p2p_open(hs)
                collective = "fused reduce + peer-memory all-reduce + Adam kernel (NVLink P2P, IPC-mapped exchange buffers)"
        stream = torch.cuda.ExternalStream(net.stream(), device=local_rank)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

        def barrier():
            net.synchronize(); torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()

        def timed_steps(k, fn):
            """k steps, each bracketed by CUDA events on the handle's stream, L2 flushed between them."""
            evs = []
            with torch.cuda.stream(stream):
                for _ in range(k):
                    flush.zero_()
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    e0.record(stream); fn(); e1.record(stream)
                    evs.append((e0, e1))
            net.synchronize(); torch.cuda.synchronize()
            return sum(a.elapsed_time(b) for a, b in evs)

        # ---- device-resident throughput (CUDA-graph step replay) ----
        sampler = ClockSampler(local_rank); sampler.start()
        for _ in range(warmup):
            net.train_step(read_loss=False)
        barrier()
        net.profile_get(reset=True)
        sampler.mark_begin()
        t_wall0 = time.perf_counter()
        ms_total = timed_steps(steps, lambda: net.train_step(read_loss=False))
        barrier()
        wall = time.perf_counter() - t_wall0
        launches = net.profile_get(reset=True)["launches"]
        # the narrow-net timed region lasts milliseconds, below nvidia-smi's 20 ms period: keep the SAME load (identical untimed
        # steps, L2 flushed) running until the clock samples cover about 0.3 s that contains the timed region.  The step count
        # comes from the max-over-ranks step time, so every rank runs the same number of (collective) steps.
        tl = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tl, op=dist.ReduceOp.MAX)
        est_step_s = float(tl.item()) / steps * 1e-3 + 1.0e-4          # + the L2 flush memset
        load_steps = max(0, min(5000, int(0.3 / est_step_s) - steps))
        if load_steps:
            timed_steps(load_steps, lambda: net.train_step(read_loss=False))
        sampler.mark_end()
        clocks = sampler.stop()
        clocks["window"] = f"timed region + {load_steps} identical untimed steps (about 0.3 s under the same load)" if load_steps else "timed region"
        net.profile_get(reset=True)
        # ---- per-kernel CUDA-event pass for the roofline (same steps, direct launches bracketed by events) ----
        net.profile_enable(True)
        timed_steps(min(steps, 10), lambda: net.train_step(read_loss=False))
        prof = net.profile_get(reset=True)
        prof["launches"] = launches
        net.profile_enable(False)
        last = net.loss_grad()
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        ms_step = ms_total / steps
        value = nf / (ms_step * 1e-3)

        # ---- end to end through the public call with host buffers ----
        e2e = None
        if not args.no_e2e:
            x_host = torch.from_numpy(net.get_points(0)).pin_memory()
            n_loc = x_host.shape[0]
            for _ in range(3):
                net.train_step_host(x_host.data_ptr(), n_loc)
            barrier()
            ms_e2e = timed_steps(steps, lambda: net.train_step_host(x_host.data_ptr(), n_loc))
            t = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda")
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e2e = float(t.item()) / steps
            e2e = {"value": nf / (ms_e2e * 1e-3), "unit": "points/s", "ms_per_step": ms_e2e,
                   "h2d_bytes_per_step": int(n_loc * cfg.d_in * 4), "d2h_bytes_per_step": 20,
                   "api": "pinn_train_step_host (C ABI), pinned host points, loss read back every step"}

        if rank == 0:
            pk = peaks()
            flops_pt, bytes_pt = algorithmic(which)
            k_ms = prof["step_kernel_ms"] / max(1, prof["step_kernel_count"])
            n_loc0 = net.local_count(0)
            ach_tf = n_loc0 * flops_pt / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
            ach_gb = n_loc0 * bytes_pt / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
            narrow_mma = prof["kernel_id"] in (202, 203, 204)      # register-chained mma.sync kernel (width 20)
            tensor_path = prec_name != "fp32" and not narrow_mma
            mma_mult = 3 if prec_name == "bf16x3" else 1     # issued tensor FLOPs per algorithmic FLOP
            if narrow_mma:
                # FP32-equivalent throughput against the FP32 FMA peak (the pipe this kernel replaces, same denominator as the
                # FP32 kernel's line); the issued tensor work against the measured mma.sync peak is reported beside it
                rf_bound = "fp32_fma"
                rf_peak, rf_src = fp32_peak()
                rf_src += "; FP32-equivalent algorithmic FLOPs 6CW/pt executed as split-BF16 mma.sync products"
                rf_kernel = f"fused_step_mma (interior, mma.sync {prec_name}, activations chained in registers)"
            elif tensor_path:
                rf_bound, rf_peak = "tensor", pk["bf16_tflops_sustained"]
                rf_src = f"{pk['source']} bf16_tflops_sustained (MEASURED_PEAKS.json); algorithmic FLOPs 6CW/pt, issued tensor FLOPs x{mma_mult}"
                rf_kernel = ("fused_step_tc2 (interior: producer + dW-server CTAs, one cooperative launch, tcgen05 " if prof["kernel_id"] // 10 == 11
                             else "fused_step_tc (interior, tcgen05 ") + prec_name + ")"
            else:
                rf_bound = "fp32_fma"
                rf_peak, rf_src = fp32_peak()
                rf_kernel = "fused_step_fp32 (interior)"
            roofline = {
                "bound": rf_bound, "kernel": rf_kernel, "achieved": ach_tf, "peak": rf_peak,
                "unit": "TFLOP/s", "frac": ach_tf / rf_peak, "traffic": traffic_bytes(which, prec_name, n_loc0), "traffic_source": (traffic_for(which, prec_name) or {}).get("source"),
                "peak_source": rf_src, "issued_tensor_tflops": ach_tf * mma_mult if tensor_path else 0.0,
                "mma_sync": mma_sync_line(prec_name, n_loc0, k_ms, cfg.depth) if narrow_mma else None,
                "kernel_ms": k_ms, "kernel_share_of_step": k_ms / ms_step if ms_step else None,
                "hbm": {"achieved": ach_gb, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach_gb / pk["hbm_gbs"], "of": pk["source"]},
                "tensor": {"achieved": ach_tf, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                           "frac": ach_tf / pk["bf16_tflops_sustained"], "of": pk["source"]},
                "flops_per_point": flops_pt, "bytes_per_point": bytes_pt, "points_per_launch": n_loc0,
            }
            line = {
                "metric": "collocation points/sec per train step (fwd+residual+bwd+Adam)",
                "value": value, "unit": "points/s", "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"fp32": "f32", "fp16x3": "f32-equivalent: operands scaled by an exact power of two and split into 2 fp16 terms (22 bits), "
                                                   "3 mma.sync products, f32 accumulate + f32 elementwise",
                          "bf16x6": "f32-equivalent: operands split into 3 bf16 terms (exact 24-bit split), 6 mma.sync products, "
                                                   "f32 accumulate + f32 elementwise",
                          "bf16x3": "bf16x3 operands (2-term split), f32 accumulate + f32 elementwise",
                          "bf16": "bf16 operands, f32 accumulate + f32 elementwise"}[prec_name], "data": "synthetic",
                "config": {"workload": CONFIG_NAMES[which], "precision": prec_name, "points_per_gpu": per_gpu, "n_interior": nf, "n_boundary": nb,
                           "n_initial": n0, "parallelism": f"dp{world}", "collective": collective if world > 1 else None, "l2": "flushed (256 MiB memset) between timed steps",
                           "kernel_id": prof["kernel_id"], "grid": prof["grid"], "block": prof["block"], "smem": prof["smem_bytes"]},
                "roofline": roofline, "e2e": e2e, "gpu_launches": prof["launches"], "clocks": clocks,
                "final_loss": last["loss"], "wall_s": wall,
                # stated tolerances of this precision against the float64 oracle (tests/test_gpu_parity.py holds them)
                "parity_tolerance": {
                    "fp32": "loss 1e-5, gradient 1e-4; 100 Adam steps: parameters 1e-4, loss 1e-5 (first 30 steps) / 5e-4 (100 steps; the CPU float32 oracle itself drifts 1.2e-5)",
                    "fp16x3": "as fp32 (loss 1e-5, gradient 1e-4; 100 steps: parameters 1e-4, loss 1e-5 / 5e-4)",
                    "bf16x6": "as fp32 (loss 1e-5, gradient 1e-4; 100 steps: parameters 1e-4, loss 1e-5 / 5e-4)",
                    "bf16x3": "loss 1e-5, gradient 5e-5; 30 Adam steps: loss 1e-4, parameters 1e-4",
                    "bf16": "loss 5e-3, gradient 1e-2 (2e-2 on the eight-layer width-256 net); 100 Adam steps: loss 1e-1 (first 20 steps) / 1e-2, parameters 1e-2",
                }[prec_name],
            }
            if with_cpu and world == 1:
                line["cpu_baseline"] = cpu_baseline(which)
                rc = reference_cuda_composed(which)
                if rc:
                    line["reference_cuda_composed"] = rc
        # cross-rank correctness travels with the scaling record: every rank must hold bit-identical parameters after the
        # same steps (the data-parallel tests need more than one GPU, which the test box does not have)
        import zlib
        crc = zlib.crc32(net.get_params().tobytes())
        ranks_identical = None
        if dist is not None:
            cs = torch.tensor([crc], dtype=torch.int64, device="cuda")
            gathered = [torch.zeros_like(cs) for _ in range(world)]
            dist.all_gather(gathered, cs)
            ranks_identical = bool(all(int(g.item()) == crc for g in gathered))
        if rank == 0:
            line["ranks_identical"] = ranks_identical
            line["params_crc32"] = crc
        net.close()
        del flush
        torch.cuda.empty_cache()
        return line if rank == 0 else None

    line = run_config(args.config, args.steps, args.warmup, args.points, not args.no_cpu_baseline)
    if world == 1 and args.config == 5 and not args.points and not args.no_also:
        also = {}
        for key, which, st in (("c2", 2, 20), ("c3", 3, 10)):
            sub = run_config(which, st, max(3, args.warmup), 0, False)
            also[key] = {k: sub[k] for k in ("value", "unit", "steps", "warmup", "ms_per_step", "dtype", "config", "roofline", "e2e",
                                             "gpu_launches", "clocks", "final_loss")}
        line["also"] = also
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
