"""Write a text summary of an .ncu-rep (raw page, selected metrics) — the files committed under profiles/.
usage: python scripts/ncu_summary.py <report.ncu-rep> <out.txt>"""
import csv, io, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units = rows[0], rows[1]
keep = ("gpu__time_duration", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput", "sm__throughput.avg.pct",
        "sm__warps_active.avg.pct", "launch__registers", "launch__occupancy_limit", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_fma.avg.pct", "sm__pipe_fma_cycles_active.avg.pct",
        "sm__pipe_tensor_cycles_active.avg.pct", "sm__inst_executed_pipe_tmem.avg.pct", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct",
        "l1tex__throughput.avg.pct", "smsp__average_warps_issue_stalled", "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg")
with open(out, "w") as f:
    for r in rows[2:]:
        f.write(f"kernel: {r[hdr.index('Kernel Name')][:120]}\n  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n")
        for i, h in enumerate(hdr):
            if h.startswith(keep) and not h.endswith(("_not_issued_per_issue_active.ratio",)):
                f.write(f"  {h:90s} {r[i]:>18s} {units[i]}\n")
print(open(out).read()[:400])
