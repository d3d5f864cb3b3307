"""Quick decoder timing on the GPU box (not a pytest): python tests/gpu_dec_bench.py [MiB] [level] [gen]."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

import gen
import rust_compression_b200  # noqa: F401
from rust_compression_b200 import device as dv

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
level = int(sys.argv[2]) if len(sys.argv) > 2 else 9
kind = sys.argv[3] if len(sys.argv) > 3 else "text"
n = mib << 20
data = gen.text(1, n) if kind == "text" else gen.mixed(1, n)
d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
ctx = dv.Context()
comp = dv.compress_tensor(ctx, level, d_in).clone()
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
res = {"mib": mib, "level": level, "gen": kind, "compressed_bytes": comp.numel()}
times = []
for it in range(4):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    got, k = ctx.decompress_device(comp, d_out)
    b.record()
    torch.cuda.synchronize()
    times.append(a.elapsed_time(b))
    assert (got, k) == (n, 0)
assert torch.equal(d_out, d_in)
res["ms"] = times
res["stats"] = ctx.dec_stats()
ctx.profile(True)
ctx.decompress_device(comp, d_out)
ctx.profile(False)
res["kernels_ms"] = {k: round(v[1], 3) for k, v in sorted(ctx.profile_records().items(), key=lambda kv: -kv[1][1])}
h_comp = comp.cpu().pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
e2e = []
for it in range(3):
    t = time.perf_counter()
    got, k = ctx.decompress_host(h_comp, h_out)
    e2e.append((time.perf_counter() - t) * 1e3)
res["e2e_ms"] = e2e
res["decode_MBps_device"] = round(n / 1e6 / (min(times[1:]) / 1e3), 1)
res["decode_MBps_e2e"] = round(n / 1e6 / (min(e2e[1:]) / 1e3), 1)
print(json.dumps(res))
