#!/usr/bin/env python3
"""bench.py — bzip2 block-compression throughput on B200 (BASELINE.json metric: uncompressed MB/s, bit-exact).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N>1 is launched by torchrun (one rank per GPU). A "step" is one pass of the whole hot path (K1..K7) over the
synthetic corpus: every GPU holds the corpus in HBM and compresses its contiguous range of blocks; the compressed
bit strings are gathered to rank 0 over NCCL and joined at bit granularity into ONE .bz2 stream.
Weak scaling: 1 GiB of text per GPU (N GiB corpus at N GPUs).  One JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# The headline workload is BASELINE.json configs[2] (level 9, text).  BZB200_BENCH_LEVEL / BZB200_BENCH_GEN=mixed
# switch to the other throughput case of BASELINE.json (configs[3]: level 1 on mixed binary/text) for side runs.
LEVEL = int(os.environ.get("BZB200_BENCH_LEVEL", "9"))
GEN = os.environ.get("BZB200_BENCH_GEN", "text")
METRIC = "bzip2 compress MB/s (uncompressed)"
BYTES_PER_GPU = int(os.environ.get("BZB200_BENCH_BYTES", str(1 << 30)))
CPU_SAMPLE_BYTES = int(os.environ.get("BZB200_CPU_SAMPLE_BYTES", str(128 << 20)))
# Algorithmic HBM bytes per unit of work for the kernels that can top the step (DESIGN.md "Kernels").
#   k2_rs_scatter : one radix pass over a sort element = 8 B read + 8 B written; pass 0 of the initial sort reads
#                   the text instead (1 B per element, the 5-byte windows overlap)
#   k2_local_sort : one work-list entry = 8 B entry read + 4 B SA slot written + 4 B rank written
#   k3_apply      : one last-column byte read + 2 B per emitted symbol
ALG_BYTES = {
    "k2_rs_scatter": lambda st: 16.0 * st["elems_sorted_radix"] - 7.0 * st["n_rle"],
    "k2_local_sort": lambda st: 16.0 * st["elems_local"],
    "k3_apply": lambda st: 1.0 * st["n_rle"] + 2.0 * st["mtf_symbols"],
}


# DRAM traffic per algorithmic byte, from the committed `ncu --set full` capture (profiles/r1b_ncu_kernels.csv:
# (dram__bytes_read.sum + dram__bytes_write.sum) of the kernel's full-size launch on the 256 MiB slice of the same corpus
# / the algorithmic bytes of that launch).  roofline.traffic = ratio x algorithmic bytes per launch.
#   k2_rs_scatter : (2.155 + 2.175) GB / (268.5 M elements x 16 B)
#   k2_local_sort : (2.911 + 2.146) GB / (230.7 M entries x 16 B)   (round 1)
#   k3_apply      : (0.380 + 0.317) GB / (268.5 M bytes + 2 B x 164.3 M symbols)
NCU_TRAFFIC_RATIO = {"k2_rs_scatter": 1.008, "k2_local_sort": 1.370, "k3_apply": 1.168}


def workload_name(n_gpus):
    gib = BYTES_PER_GPU / float(1 << 30)
    kind = "English-like text" if GEN == "text" else "mixed binary/text (64 KiB segments)"
    return (f"{gib:g} GiB synthetic {kind} per GPU, level {LEVEL} (~{int(BYTES_PER_GPU / (LEVEL * 100000 - 19))} blocks "
            f"of {LEVEL}00 kB per GPU), one .bz2 stream sharded block-wise over {n_gpus} GPU(s)")


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md).  In-process NVML
    (nvidia_ml_py) every 300 ms: a looping `nvidia-smi --query-gpu` was measured to stretch the timed steps by up to
    15 % on some boxes (each query stalls launches and synchronisations); nvidia-smi remains the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []      # nvidia-smi fallback rows
        self.samples = []   # (sm_mhz, sm_max_mhz, reason_bits)
        self.p = None
        self.t = None
        self.stop_flag = False
        self.how = None

    def _nvml_handle(self):
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.gpu)

    def start(self):
        try:
            nv, h = self._nvml_handle()
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                nv.nvmlDeviceGetCurrentClocksThrottleReasons
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)   # first calls initialise NVML state: keep that out of
            get_reasons(h)                                    # the timed region

            def loop():
                # every NVML query takes driver locks that launches and synchronisations also need: on some boxes a
                # 100 ms period stretched the 460-launch step by 13 %, so the period is 300 ms (first sample at 50 ms)
                time.sleep(0.05)
                while not self.stop_flag:
                    try:
                        self.samples.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), mx, int(get_reasons(h))))
                    except Exception:
                        pass
                    for _ in range(6):
                        if self.stop_flag:
                            break
                        time.sleep(0.05)
            self.how = "nvml"
            self.t = threading.Thread(target=loop, daemon=True)
            self.t.start()
            return
        except Exception:
            pass
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                       "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                      text=True)
            self.how = "nvidia-smi"
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 9:
                self.rows.append(f)

    def stop(self):
        if self.how == "nvml":
            self.stop_flag = True
            self.t.join(timeout=1)
            sm = [s[0] for s in self.samples]
            bits = 0
            for s in self.samples:
                bits |= s[2]
            return {"sm_mhz": statistics.median(sm) if sm else None,
                    "sm_max_mhz": self.samples[0][1] if self.samples else None,
                    "reasons": sorted(n for b, n in self.REASONS.items() if bits & b), "samples": len(self.samples),
                    "source": "nvml, 300 ms period, during the timed region"}
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows), "source": "nvidia-smi -lms 200"}


def gen_slice(rank, nbytes):
    import gen
    return gen.text(1 + rank, nbytes) if GEN == "text" else gen.mixed(1 + rank, nbytes)


def run_reference(args, rank):
    """Reference arm: the reference's CPU implementation of the path. The Rust crate cannot be built here (no
    rustc/cargo), so this times the C++ oracle port (oracle/), block-parallel over all host cores: each worker
    compresses its own contiguous slice as an independent stream (the reference itself is single-threaded)."""
    if rank != 0:
        return
    import multiprocessing as mp
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    per = int(os.environ.get("BZB200_REF_BYTES_PER_CORE", str(16 << 20)))
    decode = os.environ.get("BZB200_BENCH_MODE") == "decode"
    global _REF_DECODE
    _REF_DECODE = decode
    with mp.Pool(cores, initializer=_ref_init, initargs=(decode,)) as pool:
        # every worker keeps its slice of the synthetic corpus between steps: the timed region is compression only
        pool.map(_ref_worker, [(i, per, True) for i in range(cores)])

        def step():
            t = time.perf_counter()
            outs = pool.map(_ref_worker, [(i, per, False) for i in range(cores)], chunksize=1)
            return time.perf_counter() - t, sum(outs)
        for w in range(args.warmup):
            step()
        tot = 0.0
        for k in range(args.steps):
            dt, _ = step()
            tot += dt
    nbytes = per * cores
    ms = tot / args.steps * 1e3
    val = nbytes / (ms / 1e3) / 1e6
    line = {
        "impl": "reference", "metric": DEC_METRIC if decode else METRIC, "value": val, "unit": "MB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(args.gpus), "level": LEVEL},
        "cpu_baseline": {"value": val, "unit": "MB/s", "cores": cores, "kind": "port",
                         "sample": f"{cores} workers x {per >> 20} MiB of the same synthetic text per step (generated outside "
                                   f"the timed region), each worker one independent level-{LEVEL} stream (C++ oracle port "
                                   "of the reference " + ("decoder, streams compressed outside the timed region"
                                                          if decode else "encoder") + "; rustc unavailable)"},
        "e2e": {"value": val, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


_REF_DATA = {}
_REF_DECODE = False


def _ref_init(decode):
    global _REF_DECODE
    _REF_DECODE = decode


def _ref_worker(a):
    """Pool worker of the reference arm: slice `idx` of the synthetic corpus, generated once per process (and, for
    the decode side case, compressed once per process: the timed region is the decoder alone)."""
    idx, n, prepare = a
    import gen
    from oracle import orc
    key = (os.getpid(), n)
    if key not in _REF_DATA:
        raw = gen.text(1000 + idx, n) if GEN == "text" else gen.mixed(1000 + idx, n)
        _REF_DATA[key] = orc.compress(raw, LEVEL) if _REF_DECODE else raw
    if prepare:
        orc.lib()
        return 0
    if _REF_DECODE:
        return len(orc.decode(_REF_DATA[key]))
    return len(orc.compress(_REF_DATA[key], LEVEL))


def cpu_baseline(sample):
    from oracle import orc
    t = time.perf_counter()
    out = orc.compress(sample, LEVEL)
    dt = time.perf_counter() - t
    return {"value": len(sample) / dt / 1e6, "unit": "MB/s", "cores": 1, "kind": "port",
            "sample": f"first {len(sample) >> 20} MiB of rank 0's corpus, level {LEVEL}, single thread (the reference is "
                      f"single-threaded); C++ oracle port, ratio {len(sample) / max(1, len(out)):.3f}"}


def check_prefix(stream, corpus_prefix):
    """Bit-exact parity of the stream's leading blocks against the oracle run on a prefix of the corpus (block cuts
    depend only on preceding input, so all but the oracle's last block must match bit for bit)."""
    from oracle import orc
    r = orc.Run(corpus_prefix, LEVEL)
    if r.nblocks < 2:
        return "skipped"
    bits = r.info(r.nblocks - 2)["bit_end"]
    nb = bits // 8
    ok = stream[:nb] == r.out[:nb]
    if ok and bits % 8:
        m = (0xFF << (8 - bits % 8)) & 0xFF
        ok = (stream[nb] & m) == (r.out[nb] & m)
    return f"{r.nblocks - 1} leading blocks ({bits} bits) bit-exact vs oracle" if ok else "MISMATCH"


DEC_METRIC = "bzip2 decompress MB/s (uncompressed)"
# Algorithmic HBM bytes of the decoder's kernels (DESIGN.md "Decoder"): st = bzb200_dec_stats + compressed size
DEC_ALG_BYTES = {
    "d2_decode": lambda st: st["comp_bytes"] + 4.0 * st["pre_rle_bytes"],   # fused path: bits in; byte ‖ occ word out
    "d2_huff": lambda st: st["comp_bytes"] + 2.0 * st["symbols"],           # bits in; one u16 symbol out
    "d2_mtf_a": lambda st: 4.25 * st["symbols"],   # u16 symbol in, index byte out, 1.25 KB of chunk tables per 1 024
    "d2_mtf_c": lambda st: 4.25 * st["symbols"] + 4.0 * st["pre_rle_bytes"],  # symbol + index + tables in; word out
    "d3_scatter": lambda st: 9.0 * st["pre_rle_bytes"],                     # L + occ in, V (4 B) scattered
    "d4_walk_a": lambda st: 4.0 * st["pre_rle_bytes"],                      # one V entry per step
    "d4_walk_c": lambda st: 5.0 * st["pre_rle_bytes"],                      # one V entry per step + the byte out
    "d5_count": lambda st: 1.0 * st["pre_rle_bytes"],
    "d5_expand": lambda st: 1.0 * st["pre_rle_bytes"] + st["out_bytes"],
    "k5_crc_blocks": lambda st: 1.0 * st["out_bytes"],
}


def run_decode(args):
    """Side case, BZB200_BENCH_MODE=decode: the block-parallel GPU decoder on the stream the
    encoder produces for the headline corpus.  Same JSON contract; one GPU (replicas only at N > 1, not launched)."""
    import torch

    import rust_compression_b200  # noqa: F401
    from oracle import orc
    from rust_compression_b200 import device as dv

    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    raw = gen_slice(0, BYTES_PER_GPU)
    h_raw = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
    d_raw = h_raw.to(dev)
    ctx = dv.Context()
    d_comp = dv.compress_tensor(ctx, LEVEL, d_raw).clone()
    h_comp = d_comp.cpu().pin_memory()
    n, nc = d_raw.numel(), d_comp.numel()
    d_out = torch.empty(n, dtype=torch.uint8, device=dev)
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()

    def step_device():
        return ctx.decompress_device(d_comp, d_out)

    def step_e2e():
        return ctx.decompress_host(h_comp, h_out)

    def timed(fn, steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            res = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), res

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(0)
    sampler.start()
    l0 = ctx.launch_count()
    ms_total, res = timed(step_device, args.steps)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop()
    assert res == (n, 0), res
    same = bool(torch.equal(d_out, d_raw))
    ms_step = ms_total / args.steps
    st = ctx.dec_stats()
    st["comp_bytes"] = nc
    ctx.profile(True)
    ms_prof, _ = timed(step_device, 1)
    ctx.profile(False)
    recs = ctx.profile_records()
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    e2e_steps = max(2, args.steps // 2)
    ms_e2e, res = timed(step_e2e, e2e_steps)
    same_e2e = res == (n, 0) and bool(torch.equal(h_out, h_raw))
    # CPU beside it: the restated reference decoder, one thread, on the stream of a prefix of the corpus
    sample = d_raw[:min(n, CPU_SAMPLE_BYTES)]
    s_comp = dv.compress_tensor(ctx, LEVEL, sample).cpu().numpy().tobytes()
    t = time.perf_counter()
    back = orc.decode(s_comp)
    dt = time.perf_counter() - t
    ok_cpu = back == raw[:sample.numel()]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    top = sorted(recs.items(), key=lambda kv: -kv[1][1])
    name, (nl, kms) = top[0]
    alg = DEC_ALG_BYTES.get(name, lambda s_: nc + n)(st)
    achieved = alg / (kms / 1e3) / 1e9
    line = {
        "metric": DEC_METRIC, "value": n / (ms_step / 1e3) / 1e6, "unit": "MB/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": "decode of the level-%d stream of: %s" % (LEVEL, workload_name(1)), "level": LEVEL,
                   "compressed_bytes": nc, "stats": st,
                   "l2": "compressed input (0.3 GB) and all per-block arrays (10 GB) exceed the 126 MB L2; no flush",
                   "verified": ("device output == corpus; " if same else "DEVICE OUTPUT DIFFERS; ") +
                               ("e2e output == corpus; " if same_e2e else "E2E OUTPUT DIFFERS; ") +
                               ("block and stream CRCs verified by the decoder; oracle decoder agrees on the sample"
                                if ok_cpu else "ORACLE DECODER DISAGREES")},
        "e2e": {"value": n / (ms_e2e / e2e_steps / 1e3) / 1e6, "unit": "MB/s", "h2d_bytes_per_step": nc,
                "d2h_bytes_per_step": n, "steps": e2e_steps, "api": "bzb200_decompress_host (C ABI, pinned host in/out)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "launches_per_step": nl, "avg_launch_ms": kms / max(1, nl),
                     "kernel_share_of_step": kms / ms_prof, "algorithmic_bytes_per_launch": alg / max(1, nl),
                     "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback",
                     "path": {"achieved": (n + nc) / (ms_step / 1e3) / 1e9, "unit": "GB/s",
                              "frac": (n + nc) / (ms_step / 1e3) / 1e9 / peak,
                              "note": "whole path: (compressed in + bytes out) / device time"}},
        "kernels_ms_per_step": {k: round(v[1], 3) for k, v in top[:12]},
        "profiled_step_ms": round(ms_prof, 3),
        "cpu_baseline": {"value": sample.numel() / dt / 1e6, "unit": "MB/s", "cores": 1, "kind": "port",
                         "sample": f"stream of the first {sample.numel() >> 20} MiB of the corpus, single thread, restated "
                                   "reference BZip2Decoder (oracle/bz2_decoder_oracle.cpp)"},
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if os.environ.get("BZB200_BENCH_MODE") == "decode":
        if rank == 0:
            run_decode(args)
        return

    import torch
    import torch.distributed as dist

    import rust_compression_b200  # noqa: F401  (fails loudly if libbzb200.so is missing)
    from rust_compression_b200 import device as dv
    from rust_compression_b200 import sharded

    # rank 0 prints exactly one line on stdout: keep NCCL's version banner (NCCL_DEBUG=VERSION/INFO) off it
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO") and not os.environ.get("BZB200_KEEP_NCCL_DEBUG"):
        os.environ["NCCL_DEBUG"] = "WARN"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # ---- synthetic corpus: each rank generates its slice; slices are all-gathered so every GPU holds the corpus
    t0 = time.time()
    raw = gen_slice(rank, BYTES_PER_GPU)
    h_slice = torch.frombuffer(bytearray(raw), dtype=torch.uint8).pin_memory()
    d_slice = h_slice.to(dev)
    if world > 1:
        d_full = torch.empty(world * BYTES_PER_GPU, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(d_full, d_slice)
    else:
        d_full = d_slice
    gen_s = time.time() - t0
    h_out = torch.empty(dv.max_output_bytes(LEVEL, world * BYTES_PER_GPU), dtype=torch.uint8).pin_memory() \
        if rank == 0 else None

    ctx = dv.Context()
    total_bytes = world * BYTES_PER_GPU

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = None
        for _ in range(steps):
            tw = time.perf_counter()
            res = fn()
            if os.environ.get("BZB200_BENCH_DEBUG"):
                torch.cuda.synchronize()
                sys.stderr.write(f"[rank {rank}] {fn.__name__} wall {1e3 * (time.perf_counter() - tw):.1f} ms\n")
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), res

    def step_device():
        return sharded.compress_sharded(ctx, LEVEL, d_full)

    def step_e2e():
        if world == 1:
            n = ctx.compress_host(LEVEL, h_slice, h_out)
            return h_out[:n], {"h2d_bytes": h_slice.numel(), "d2h_bytes": n}
        return sharded.compress_host_sharded(ctx, LEVEL, h_slice, h_out)

    # ---- device-resident throughput
    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    ms_total, (d_stream, info) = timed(step_device, args.steps)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    sstats = ctx.sort_stats()
    ms_step = ms_total / args.steps
    value = total_bytes / (ms_step / 1e3) / 1e6
    # one more step with a CUDA-event pair around every launch (same stream) for the per-kernel breakdown; it is
    # not part of the timed region because the event traffic slows the host side of the step down
    ctx.profile(True)
    ms_prof, _ = timed(step_device, 1)
    ctx.profile(False)
    recs = ctx.profile_records()

    # ---- end to end with host buffers
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    e2e_steps = max(2, args.steps // 2)
    ms_e2e, (h_stream, einfo) = timed(step_e2e, e2e_steps)
    e2e_value = total_bytes / (ms_e2e / e2e_steps / 1e3) / 1e6

    lsum = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(lsum)

    if rank == 0:
        stream = h_stream.numpy().tobytes()
        out_bytes = len(stream)
        dev_stream = d_stream.cpu().numpy().tobytes()
        verified = "device and e2e streams identical; " if dev_stream == stream else "DEVICE/E2E STREAMS DIFFER; "
        verified += check_prefix(stream, raw[: min(len(raw), 6 << 20)])
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        # dominant kernel by device time inside the timed region
        top = sorted(recs.items(), key=lambda kv: -kv[1][1])
        name, (nl, kms) = top[0] if top else ("none", (0, 0.0))
        step_ms_kernels = sum(v[1] for v in recs.values())
        if name in ALG_BYTES:
            alg_bytes = ALG_BYTES[name](sstats)
        else:
            alg_bytes = total_bytes / world + out_bytes / world
        achieved = alg_bytes / (kms / 1e3) / 1e9 if kms > 0 else 0.0
        roofline = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak,
                    "traffic": (NCU_TRAFFIC_RATIO[name] * alg_bytes / max(1, nl)) if name in NCU_TRAFFIC_RATIO else None,
                    "traffic_source": "profiles/r1b_ncu_kernels.csv (ncu --set full, dram bytes / algorithmic bytes of "
                                      "the full-size launch) x algorithmic bytes per launch",
                    "peak_source": peak_src,
                    "launches_per_step": nl, "avg_launch_ms": kms / max(1, nl),
                    "kernel_share_of_step": kms / ms_prof,
                    "algorithmic_bytes_per_launch": alg_bytes / max(1, nl),
                    "note": "achieved = algorithmic bytes of the kernel's launches in one step / their summed CUDA-event "
                            "time (profiled step, events on the launching stream)",
                    "path": {"achieved": (total_bytes + out_bytes) / (ms_step / 1e3) / 1e9 / world, "unit": "GB/s per GPU",
                             "frac": (total_bytes + out_bytes) / (ms_step / 1e3) / 1e9 / world / peak,
                             "note": "whole path: (input + output bytes) / device time (SURVEY.md 8(d))"}}
        line = {
            "metric": METRIC, "value": value, "unit": "MB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(world), "level": LEVEL, "bytes_per_gpu": BYTES_PER_GPU,
                       "blocks": info["nblocks"], "compressed_bytes": out_bytes,
                       "ratio": total_bytes / max(1, out_bytes),
                       "l2": "inputs (1 GiB per GPU) are larger than the 126 MB L2; no explicit flush",
                       "sort": sstats, "verified": verified, "corpus_gen_s": round(gen_s, 1)},
            "e2e": {"value": e2e_value, "unit": "MB/s", "h2d_bytes_per_step": int(total_bytes),
                    "d2h_bytes_per_step": int(out_bytes), "steps": e2e_steps,
                    "api": "bzb200_compress_host (C ABI, pinned host in/out)" if world == 1 else
                           "sharded.compress_host_sharded (pinned slices H2D + NCCL all-gather + C ABI + D2H)"},
            "gpu_launches": int(lsum.item()),
            "roofline": roofline,
            "kernels_ms_per_step": {k: round(v[1], 3) for k, v in top[:14]},
            "profiled_step_ms": round(ms_prof, 3),
            "kernel_time_ms_per_step": round(step_ms_kernels, 3),
            "cpu_baseline": cpu_baseline(raw[:CPU_SAMPLE_BYTES]),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
