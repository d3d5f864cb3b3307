#!/usr/bin/env python3
"""Per-kernel DRAM traffic next to the algorithmic bytes of the SAME launch -> profiles/r2_ncu_kernels.csv, the table
bench.py reads for `roofline.traffic` (no constants in bench.py).

    python profiles/make_traffic_table.py <ncu raw page csv> <log of the profiled run> <out csv>

Inputs (both written by profiles/ncu_encoder.sh / tools/gpu_ab.sh on the GPU box):
  * the raw page (`ncu -i rep --page raw --csv`) of an `ncu --set full --kernel-name-base demangled --kernel-id :::1`
    capture: the FIRST launch of every kernel of one device-resident compression of the 256 MiB text corpus, level 9;
  * the log of that run, which carries the per-round work of the sort (BZB200_TRACE_ROUNDS=1) and the path statistics
    printed by tools/gpu_enc_once.py — the units the algorithmic bytes are computed from.

Algorithmic bytes per unit are the compulsory HBM bytes stated in DESIGN.md section 4 (one line per kernel below).
"""
import ast
import csv
import re
import sys


def units_from_log(path):
    txt = open(path).read()
    u = {}
    m = re.search(r"k2 initial: elems (\d+) passes (\d+) .*?-> unresolved (\d+) max_radix (\d+) radix (\d+)", txt)
    u["n_rle"], u["unres0"], u["radix1"] = int(m.group(1)), int(m.group(3)), int(m.group(5))
    m = re.search(r"k2 round 1 \(h \d+\): local (\d+) radix (\d+)", txt)
    u["local1"] = int(m.group(1)) if m else 0
    m = re.search(r"k2 round 2 \(h \d+\): local (\d+) radix (\d+)", txt)
    u["local2"] = int(m.group(1)) if m else 0
    m = re.search(r"compressed (\d+) -> (\d+) (\{.*\})", txt)
    u["n_in"], u["n_out"] = int(m.group(1)), int(m.group(2))
    u.update(ast.literal_eval(m.group(3)))
    return u


# kernel-name regex -> (bench.py name, algorithmic bytes of the captured launch as a function of the units, what it is)
MODEL = [
    (r"k2_os_scatter<0, 0, 9>", "k2_rs_scatter", lambda u: 16 * u["n_rle"], "8 B read + 8 B written per element"),
    (r"k2_os_scatter<[1-9], \d, \d>", "k2_rs_scatter_pass0", lambda u: 9 * u["n_rle"],
     "pass 0: 1 B of text read + 8 B written per element"),
    (r"k2_os_scatter<0, 0, 8>", "k2_rs_scatter_rounds", lambda u: 16 * u["radix1"], "8 B + 8 B per radix-path element"),
    (r"k2_local_sort_rx", "k2_local_sort_rx", lambda u: 16 * u["local2"],
     "round 2 (the first sparse round): 8 B entry read + 4 B SA + 4 B rank per entry"),
    (r"k2_local_sort", "k2_local_sort", lambda u: 16 * u["local1"], "round 1: 8 B entry read + 4 B SA + 4 B rank per entry"),
    (r"k2_gather", "k2_gather", lambda u: 4 * u["n_rle"] + 12 * u["unres0"],
     "4 B SA per slot + 4 B key gathered + 8 B entry written per unresolved slot"),
    (r"k2_rg_apply<\(bool\)1>|k2_rg_apply<1>", "k2_rg_apply", lambda u: 16 * u["n_rle"],
     "8 B element read + 4 B SA + 4 B rank per element"),
    (r"k2_rg_flags", "k2_rg_flags", lambda u: 8 * u["n_rle"], "8 B per element"),
    (r"k2_finish", "k2_finish", lambda u: 6 * u["n_rle"], "4 B SA + 1 B gathered + 1 B written per slot"),
    (r"k2_pair_hist|k2_os_hist_txt", "k2_pair_hist", lambda u: u["n_rle"], "1 B per element"),
    (r"k3_apply", "k3_apply", lambda u: u["n_rle"] + 2 * u["mtf_symbols"], "1 B read + 2 B per emitted symbol"),
    (r"k3_chunk_scan_a", "k3_chunk_scan_a", lambda u: u["n_rle"], "1 B per last-column byte"),
    (r"k1_tile_heads", "k1_tile_heads", lambda u: u["n_in"], "1 B per input byte"),
    (r"k1_tile_counts", "k1_tile_counts", lambda u: u["n_in"], "1 B per input byte"),
    (r"k1_scatter", "k1_scatter", lambda u: u["n_in"] + u["n_rle"], "1 B read + 1 B written"),
    (r"k5_crc_blocks", "k5_crc_blocks", lambda u: u["n_in"], "1 B per input byte"),
    (r"k1_inuse", "k1_inuse", lambda u: u["n_rle"], "1 B per RLE1 byte"),
    (r"k4_cost_select", "k4_cost_select", lambda u: 2 * u["mtf_symbols"], "2 B per symbol"),
    (r"k6_pack_symbols", "k6_pack_symbols", lambda u: 2 * u["mtf_symbols"] + u["n_out"], "2 B per symbol + output bytes"),
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def main(raw, log, out):
    u = units_from_log(log)
    rows = list(csv.reader(open(raw)))
    hdr, unit, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, m):
        return float(r[col[m]].replace(",", "")) * SCALE.get(unit[col[m]], 1.0)

    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "bench_name", "time_ms", "dram_read_bytes", "dram_write_bytes", "alg_bytes", "traffic_ratio",
                    "alg_GBps", "dram_pct_of_peak", "warps_active_pct", "issue_active_pct", "warp_inst", "alg_model"])
        for r in data:
            name = r[col["Kernel Name"]].split("(")[0].replace("bzb::", "").replace("void ", "")
            hit = next((m for m in MODEL if re.search(m[0], name)), None)
            rd, wr, t = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum"), val(r, "gpu__time_duration.sum")
            alg = float(hit[2](u)) if hit else 0.0
            w.writerow([name, hit[1] if hit else "", f"{t:.4f}", f"{rd:.0f}", f"{wr:.0f}", f"{alg:.0f}",
                        f"{(rd + wr) / alg:.3f}" if alg else "", f"{alg / t / 1e6:.1f}" if alg else "",
                        r[col["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]],
                        r[col["sm__warps_active.avg.pct_of_peak_sustained_active"]],
                        r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]],
                        r[col["smsp__inst_executed.sum"]], hit[3] if hit else ""])
    print(out, "units:", {k: u[k] for k in ("n_in", "n_rle", "unres0", "local1", "radix1", "mtf_symbols", "n_out")})


if __name__ == "__main__":
    main(*sys.argv[1:4])
