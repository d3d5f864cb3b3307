#!/usr/bin/env python3
"""SASS opcode histogram of every kernel in libbzb200.so (no GPU needed).

    python profiles/sass_histogram.py rust-compression_b200/libbzb200.so profiles/r2_sass_histogram.csv

One row per kernel: instruction count and the counts of the opcodes that say how a kernel touches memory and
synchronises (LDG/STG with width, LDS/STS, ATOMS/ATOMG/RED, MATCH, SHFL, VOTE, REDUX, BAR, UBLKCP = the 1-D TMA bulk
copy, SYNCS = mbarrier operations), then the ten most frequent opcodes.  tcgen05/UTMALDG do not appear by design:
nothing on this path is a dense contraction or a multi-dimensional tile."""
import collections
import csv
import re
import subprocess
import sys

TRACK = ["LDG.E.128", "LDG.E.64", "LDG.E", "LDG.E.U8", "STG.E.128", "STG.E.64", "STG.E", "STG.E.U8", "STG.E.U16", "LDS", "STS",
         "ATOMS", "ATOMG", "ATOM", "RED", "MATCH", "SHFL", "VOTE", "REDUX", "BAR", "UBLKCP", "SYNCS", "LDL", "STL"]


def main(lib, out):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = kernels.setdefault(re.sub(r"\(.*", "", name).replace("bzb::", "").replace("void ", ""),
                                     collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "instructions"] + TRACK + ["top10"])
        for name, c in kernels.items():
            def fam(op):
                if op in ("LDS", "STS", "ATOMS", "ATOMG", "ATOM", "RED", "MATCH", "SHFL", "VOTE", "REDUX", "BAR", "UBLKCP",
                          "SYNCS", "LDL", "STL"):
                    return sum(v for k, v in c.items() if k.split(".")[0] == op)
                # loads/stores by width: exact prefix up to the width suffix, cache hints ignored
                tot = 0
                for k, v in c.items():
                    parts = k.split(".")
                    if parts[0] != op.split(".")[0]:
                        continue
                    width = next((p for p in parts if p in ("128", "64", "U8", "U16", "S8", "S16")), "32")
                    want = next((p for p in op.split(".") if p in ("128", "64", "U8", "U16")), "32")
                    tot += v if width == want else 0
                return tot
            total = sum(c.values())
            top = " ".join(f"{k}:{v}" for k, v in c.most_common(10))
            w.writerow([name, total] + [fam(op) for op in TRACK] + [top])
    print(out, len(kernels), "kernels")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
