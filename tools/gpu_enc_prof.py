#!/usr/bin/env python3
"""Per-kernel device times (CUDA events around every launch) of one device-resident compression of a synthetic corpus,
after a warm-up run — the quick A/B tool for kernel experiments (env switches are read by the library).

    python tools/gpu_enc_prof.py <MiB> <level> <text|mixed>
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

import gen
import rust_compression_b200  # noqa: F401
from rust_compression_b200 import device as dv

mib, level, kind = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
seg = 64 << 20
n = mib << 20
parts = [(gen.text if kind == "text" else gen.mixed)(1 + k, min(seg, n - k * seg)) for k in range((n + seg - 1) // seg)]
d_in = torch.from_numpy(np.frombuffer(b"".join(parts), dtype=np.uint8).copy()).cuda()
ctx = dv.Context()
d_out = torch.zeros(dv.max_output_bytes(level, n), dtype=torch.uint8, device="cuda")


def run():
    d_out.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nbytes = ctx.compress_device(level, d_in, d_out)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), nbytes


run()
ms = min(run()[0] for _ in range(3))
ctx.profile(True)
run()
ctx.profile(False)
recs = sorted(ctx.profile_records().items(), key=lambda kv: -kv[1][1])
print("step %.2f ms  %.0f MB/s" % (ms, n / ms / 1e3), ctx.sort_stats())
print("  ".join("%s %.2f(%d)" % (k, v[1], v[0]) for k, v in recs[:12]))
