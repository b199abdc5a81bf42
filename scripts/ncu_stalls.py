"""Per-SASS-instruction stall reasons of one kernel (ncu source page): top instructions by long-scoreboard / barrier / wait samples.
usage: python scripts/ncu_stalls.py <report.ncu-rep> [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if "Source" in r and "Address" in r][0]
hdr = rows[hi]; data = rows[hi + 1:]
col = {h: i for i, h in enumerate(hdr)}
def num(r, h):
    try: return int(r[col[h]])
    except Exception: return 0
tot = sum(num(r, "# Samples") for r in data)
for reason in ("stall_long_sb", "stall_barrier", "stall_wait", "stall_short_sb", "stall_branch_resolving"):
    t = sum(num(r, reason) for r in data)
    print(f"== {reason}: {t} samples of {tot} ({100.0 * t / max(tot, 1):.1f}%)")
    for r in sorted(data, key=lambda r: -num(r, reason))[:top if reason == "stall_long_sb" else 6]:
        print(f"   {num(r, reason):7d}  {r[col['Source']][:110]}")
