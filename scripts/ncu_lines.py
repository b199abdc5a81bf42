"""Rank source lines of one kernel by executed instructions / stall samples.
usage: python scripts/ncu_lines.py <report.ncu-rep> <mangled-kernel-substring> [top]
Joins ncu's SASS page (per-instruction counters) with nvdisasm -g line info of the built library."""
import csv, io, os, re, subprocess, sys, tempfile
from collections import defaultdict
rep, sub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "cuda-pinn_b200", "libpinn_b200.so")], cwd=tmp, capture_output=True)
sass = []
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):      # one cubin per source file of the library
    text = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
    if any(l.startswith(".text.") and sub in l for l in text):
        sass = text
        break
start = [i for i, l in enumerate(sass) if l.startswith(".text.") and sub in l][0]
cur, ins = None, []
for l in sass[start + 1:]:
    if l.startswith(".text.") and ins:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((m.group(2), cur))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if "Source" in r and "Address" in r][0]
hdr = rows[hi]; ii = hdr.index("Instructions Executed"); wi = hdr.index("Warp Stall Sampling (All Samples)")
data = rows[hi + 1:]
agg = defaultdict(lambda: [0, 0])
ops = defaultdict(lambda: [0, 0])
for k in range(min(len(data), len(ins))):
    try:
        ie, ws = int(data[k][ii]), int(data[k][wi])
    except Exception:
        continue
    agg[ins[k][1]][0] += ie; agg[ins[k][1]][1] += ws
    op = ins[k][0].split()[0] if not ins[k][0].startswith("@") else ins[k][0].split()[1]
    ops[op.split(".")[0]][0] += ie; ops[op.split(".")[0]][1] += ws
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
src = {}
print(f"sass instructions {len(ins)} (ncu rows {len(data)}), executed {tot}, stall samples {tots}")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    f, ln = key if key else ("?", 0)
    path = os.path.join(root, "cuda-pinn_b200", "csrc", f)
    if f not in src and os.path.exists(path):
        src[f] = open(path).read().split("\n")
    text = src[f][ln - 1].strip()[:95] if f in src and ln - 1 < len(src[f]) else ""
    print(f"{v[0] / tot * 100:5.1f}% inst {v[1] / tots * 100:5.1f}% stall | {f}:{ln} {text}")
print("--- by opcode")
for op, v in sorted(ops.items(), key=lambda kv: -kv[1][0])[:22]:
    print(f"{v[0] / tot * 100:5.1f}% inst {v[1] / tots * 100:5.1f}% stall | {op}")
