"""A/B timing of bzb200_compress_host (pinned host in/out): one-shot copy vs segmented pipeline.  Diagnostic script,
not a test:  python tests/gpu_e2e_ab.py [MiB]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gen  # noqa: E402
from rust_compression_b200 import device as dv  # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
data = gen.text(1, mib << 20)
h_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).pin_memory()
h_out = torch.empty(dv.max_output_bytes(9, len(data)), dtype=torch.uint8).pin_memory()
ref = None
for seg in ("0", str(1 << 40), str(512 << 20), str(256 << 20), str(128 << 20), str(64 << 20), str(1 << 40)):
    if seg == "0":
        os.environ.pop("BZB200_HOST_SEGMENT", None)
    else:
        os.environ["BZB200_HOST_SEGMENT"] = seg
    ctx = dv.Context()
    ts = []
    for it in range(6):
        torch.cuda.synchronize()
        t = time.perf_counter()
        n = ctx.compress_host(9, h_in, h_out)
        ts.append(time.perf_counter() - t)
    out = h_out[:n].numpy().tobytes()
    if ref is None:
        ref = out
    ts = sorted(ts[1:])
    print(f"segment {seg:>14}: median {1e3 * ts[len(ts) // 2]:.1f} ms  min {1e3 * ts[0]:.1f} ms  "
          f"{len(data) / ts[len(ts) // 2] / 1e6:.0f} MB/s  same={out == ref}", flush=True)
    ctx.close()
