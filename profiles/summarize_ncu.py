#!/usr/bin/env python3
"""Turns an `ncu --set full` report into the per-kernel CSV kept under profiles/ (the .ncu-rep itself is tens of MB).

    python profiles/summarize_ncu.py gpurun_out/r1b_full_256mib.ncu-rep profiles/r1b_ncu_kernels.csv

One row per kernel NAME: launches, summed/mean duration, and the launch-weighted means of the metrics named in
/opt/skills/guides/B200_PROFILING.md (DRAM bytes and throughput, warps active, issue active, registers, stall mix).
Needs only `ncu -i` (no GPU)."""
import csv
import subprocess
import sys
from collections import OrderedDict

METRICS = OrderedDict([
    ("gpu__time_duration.sum", "time_ms"),
    ("dram__bytes_read.sum", "dram_read_GB"),
    ("dram__bytes_write.sum", "dram_write_GB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("launch__registers_per_thread", "regs"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
])
UNIT_SCALE = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0,
              "Tbyte": 1e3}


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    kcol = col["Kernel Name"]
    agg = OrderedDict()
    for r in data:
        name = r[kcol].split("(")[0].replace("bzb::", "")
        a = agg.setdefault(name, {"launches": 0, **{v: 0.0 for v in METRICS.values()}})
        a["launches"] += 1
        for m, short in METRICS.items():
            if m not in col:
                continue
            try:
                v = float(r[col[m]].replace(",", ""))
            except ValueError:
                continue
            v *= UNIT_SCALE.get(units[col[m]], 1.0)
            a[short] += v
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "time_ms_sum", "dram_read_GB_sum", "dram_write_GB_sum", "warp_inst_sum"] +
                   [s + "_mean" for s in METRICS.values() if s not in ("time_ms", "dram_read_GB", "dram_write_GB", "warp_inst")])
        for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["time_ms"]):
            n = a["launches"]
            w.writerow([name, n, f"{a['time_ms']:.4f}", f"{a['dram_read_GB']:.4f}", f"{a['dram_write_GB']:.4f}",
                        f"{a['warp_inst']:.0f}"] +
                       [f"{a[s] / n:.3f}" for s in METRICS.values()
                        if s not in ("time_ms", "dram_read_GB", "dram_write_GB", "warp_inst")])
    print(f"{out}: {len(agg)} kernels, {sum(a['launches'] for a in agg.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
