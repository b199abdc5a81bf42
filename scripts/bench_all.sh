#!/bin/bash
# Round-end measurement pass on one B200: every BASELINE config at its default precision, the C2 precision variants,
# the reference arm, and the ncu launch list of the default command (taken after it ran plainly).  Output: gpurun_out/.
set -u
mkdir -p gpurun_out
for c in 1 3 4 5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_c$c.err | tail -1 > gpurun_out/bench_c$c.json
done
for p in fp32 bf16x6 bf16x3; do
  timeout 300 python bench.py --config 2 --precision $p --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_c2_$p.json
done
timeout 600 python bench.py 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01_launches_c2_default.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_*.json")):
    try:
        d = json.loads(open(f).read())
        r = d.get("roofline") or {}
        e = d.get("e2e") or {}
        print(f.split("/")[-1], d.get("config", {}).get("precision"), "Mpts/s %.2f" % (d["value"] / 1e6), "ms %.4f" % d["ms_per_step"],
              "kernel_ms %.4f" % r.get("kernel_ms", 0), "frac %.3f" % r.get("frac", 0), "e2e %.2f" % ((e.get("value") or 0) / 1e6), d.get("gpu_launches"))
    except Exception as ex:
        print(f, "unreadable", ex)
PY
