#!/usr/bin/env python3
"""One device-resident compression of a synthetic corpus — the workload ncu profiles (profiles/ncu_encoder.sh).

    python tools/gpu_enc_once.py <MiB> <level> <text|mixed> [repeats]
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
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
seg = 64 << 20
n = mib << 20
parts = [(gen.text if kind == "text" else gen.mixed)(1 + k, min(seg, n - k * seg)) for k in range((n + seg - 1) // seg)]
d_in = torch.from_numpy(np.frombuffer(b"".join(parts), dtype=np.uint8).copy()).cuda()
ctx = dv.Context()
d_out = torch.zeros(dv.max_output_bytes(level, n), dtype=torch.uint8, device="cuda")
for _ in range(reps):
    d_out.zero_()
    nbytes = ctx.compress_device(level, d_in, d_out)
torch.cuda.synchronize()
print("compressed", n, "->", nbytes, ctx.sort_stats())
