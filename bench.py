#!/usr/bin/env python3
"""bench.py — bzip2 block-compression throughput on B200 (BASELINE.json metric: uncompressed MB/s, bit-exact).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload = BASELINE.json configs[2] as written: ONE 1 GiB synthetic English-like text corpus, level 9 (~1200 blocks of
900 kB), sharded block-wise over the N GPUs of the box — strong scaling: the corpus does not grow with N.  N > 1 is
launched by torchrun (one rank per GPU).  A "step" is one pass of the whole hot path (K1..K7) over the corpus.
  value  device-resident: every rank holds ITS SLICE of the corpus in HBM (N = 1: the whole corpus), compresses the
         blocks that start in it and the bit strings are gathered to rank 0 over NCCL into one .bz2 stream in HBM
         (rust-compression_b200/sharded.py; the sliced plan of include/bzb200.h section 2b).
  e2e    host buffers through the C ABI: N = 1 bzb200_compress_host; N > 1 the in-library multi-GPU engine
         (bzb200_pool_compress_host, include/bzb200.h section 2c) driven from rank 0's process over all N GPUs with the
         corpus and the stream in pinned host memory — what `BZip2Encoder::new(9)` binds on a multi-GPU box.
  e2e_stream  the same through the drop-in object: bzb200_enc_write in 1 MiB pieces, finish, read (the Rust shim's
         pattern, rust/bzip2_b200.rs).
Every stream produced is checked against the oracle: ALL blocks and the trailer, block-parallel on the host cores
(oracle/verify.py), outside the timed regions.  Side cases (BASELINE.json configs[1], [3], [4] and the decoder) run after
the headline on rank 0 and are reported under config.side_cases.  One JSON line on rank 0.
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
# switch the headline to the other throughput case of BASELINE.json (configs[3]) for experiments.
LEVEL = int(os.environ.get("BZB200_BENCH_LEVEL", "9"))
GEN = os.environ.get("BZB200_BENCH_GEN", "text")
METRIC = "bzip2 compress MB/s (uncompressed)"
TOTAL_BYTES = int(os.environ.get("BZB200_BENCH_BYTES", str(1 << 30)))
CPU_SAMPLE_BYTES = int(os.environ.get("BZB200_CPU_SAMPLE_BYTES", str(128 << 20)))
SEG = 64 << 20   # the corpus is a sequence of 64 MiB segments, segment k = gen.text(1 + k, SEG) / gen.mixed(1 + k, SEG)
SIDE = os.environ.get("BZB200_BENCH_SIDE", "1") != "0"
# Algorithmic HBM bytes per unit of work for the kernels that can top the step (DESIGN.md "Kernels").
#   k2_rs_scatter : one radix pass over a sort element = 8 B read + 8 B written; pass 0 of the initial sort reads
#                   the text instead (1 B per element, the key windows overlap)
#   k2_local_sort : one work-list entry = 8 B entry read + 4 B SA slot written + 4 B rank written
#   k3_apply      : one last-column byte read + 2 B per emitted symbol
ALG_BYTES = {
    "k2_rs_scatter": lambda st: 16.0 * st["elems_sorted_radix"] - 7.0 * st["n_rle"],
    "k2_local_sort": lambda st: 16.0 * st["elems_local"],
    "k3_apply": lambda st: 1.0 * st["n_rle"] + 2.0 * st["mtf_symbols"],
}


def ncu_traffic_ratio(kernel):
    """DRAM traffic per algorithmic byte of `kernel`, read from the committed ncu table of THIS round
    (profiles/r2_ncu_kernels.csv, written by profiles/make_traffic_table.py from one `ncu --set full` record per kernel
    of the shipped library: columns kernel, bench_name, dram_read_bytes, dram_write_bytes, alg_bytes of the SAME
    launch).  Returns (ratio, kernel name in the table, path); ratio None when the table has no row."""
    path = os.path.join(ROOT, "profiles", "r2_ncu_kernels.csv")
    try:
        import csv
        for row in csv.DictReader(open(path)):
            if row.get("bench_name") == kernel:
                alg = float(row.get("alg_bytes") or 0)
                if alg > 0:
                    return (float(row["dram_read_bytes"]) + float(row["dram_write_bytes"])) / alg, row["kernel"], path
    except Exception:
        pass
    return None, None, path


def workload_name(n_gpus, level=None, gen_name=None, total=None):
    level = LEVEL if level is None else level
    gen_name = GEN if gen_name is None else gen_name
    total = TOTAL_BYTES if total is None else total
    gib = total / float(1 << 30)
    kind = "English-like text" if gen_name == "text" else "mixed binary/text (64 KiB segments)"
    return (f"{gib:g} GiB synthetic {kind} corpus, level {level} (~{int(total / (level * 100000 - 19))} blocks of "
            f"{level}00 kB), one .bz2 stream sharded block-wise over {n_gpus} GPU(s)")


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


# ------------------------------------------------------------------------------------------------ corpus
def _gen_segment(a):
    path, k, lo, n, gen_name = a
    import numpy as np
    import gen
    raw = gen.text(1 + k, n) if gen_name == "text" else gen.mixed(1 + k, n)
    m = np.memmap(path, dtype=np.uint8, mode="r+")
    m[lo:lo + n] = np.frombuffer(raw, dtype=np.uint8)
    m.flush()
    return k


def _wait_for(paths, timeout=900):
    t0 = time.time()
    while not all(os.path.exists(p) for p in paths):
        if time.time() - t0 > timeout:
            raise RuntimeError("timed out waiting for the other ranks' corpus segments")
        time.sleep(0.05)


def make_corpus(gen_name, total, rank, world):
    """The corpus as a read-only numpy memmap shared by the ranks of the box (/dev/shm): segment k is generated by rank
    k mod world with a few forked worker processes.  Runs BEFORE CUDA / NCCL are touched (fork safety); the ranks meet
    through marker files (same parent pid under torchrun)."""
    import multiprocessing as mp
    import numpy as np
    tag = os.getppid() if "RANK" in os.environ else os.getpid()
    path = f"/dev/shm/bzb200_bench_{os.getuid()}_{tag}_{gen_name}_{total}.bin"
    if rank == 0:
        with open(path, "wb") as f:
            f.truncate(total)
        open(path + ".ready", "w").close()
    _wait_for([path + ".ready"])
    segs = [(path, k, k * SEG, min(SEG, total - k * SEG), gen_name) for k in range((total + SEG - 1) // SEG)]
    mine = [s for s in segs if s[1] % world == rank]
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    procs = max(1, min(len(mine), max(1, cores // max(1, world)), 16))
    if mine:
        with mp.get_context("fork").Pool(procs) as pool:
            pool.map(_gen_segment, mine)
    open(path + f".done.{rank}", "w").close()
    _wait_for([path + f".done.{r}" for r in range(world)])
    return np.memmap(path, dtype=np.uint8, mode="r"), path


def drop_corpus(path, world):
    for p in [path, path + ".ready"] + [path + f".done.{r}" for r in range(world)]:
        try:
            os.unlink(p)
        except OSError:
            pass


# ------------------------------------------------------------------------------------------------ CPU arms
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, rank):
    """Reference arm: the reference's CPU implementation of the path.  The Rust crate cannot be built here (no
    rustc/cargo), so this times the C++ oracle port (oracle/): the block cuts in one sequential pass
    (oracle orc_cut_table, the reference's EncoderInner::next bookkeeping), then every block encoded block-parallel on
    all host cores — the reference itself is single-threaded, which is the second figure of the line.  Each step is a
    bounded sample: the leading cores x 16 MiB of the SAME corpus the GPU arm compresses."""
    if rank != 0:
        return
    decode = os.environ.get("BZB200_BENCH_MODE") == "decode"
    from oracle import orc, verify
    cores = host_cores()
    per = int(os.environ.get("BZB200_REF_BYTES_PER_CORE", str(16 << 20)))
    sample_n = min(TOTAL_BYTES, per * cores)
    corpus, path = make_corpus(GEN, sample_n, 0, 1)
    import numpy as np
    sample = np.ascontiguousarray(corpus[:sample_n])
    drop_corpus(path, 1)
    if decode:
        table = verify.cut_table(sample, LEVEL)
        # independent streams of ~16 MiB each, decoded by the restated reference decoder on all cores
        parts = [orc.compress(sample[lo:lo + per], LEVEL) for lo in range(0, sample_n, per)]
        from concurrent.futures import ThreadPoolExecutor

        def step():
            t = time.perf_counter()
            with ThreadPoolExecutor(max_workers=cores) as ex:
                n = sum(len(x) for x in ex.map(orc.decode, parts))
            return time.perf_counter() - t, n
    else:
        def step():
            t = time.perf_counter()
            table = verify.cut_table(sample, LEVEL)
            bits, _ = verify.encode_blocks_parallel(sample, LEVEL, table, threads=cores)
            return time.perf_counter() - t, bits
    for _ in range(args.warmup):
        step()
    tot = 0.0
    for _ in range(args.steps):
        dt, _ = step()
        tot += dt
    ms = tot / args.steps * 1e3
    val = sample_n / (ms / 1e3) / 1e6
    # the faithful figure: one thread (the reference is single-threaded), on a smaller sample
    one_n = min(sample_n, 16 << 20)
    t = time.perf_counter()
    if decode:
        orc.decode(parts[0])
        one_n = min(per, sample_n)
    else:
        orc.compress(sample[:one_n], LEVEL)
    one = one_n / (time.perf_counter() - t) / 1e6
    line = {
        "impl": "reference", "metric": DEC_METRIC if decode else METRIC, "value": val, "unit": "MB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(args.gpus), "level": LEVEL},
        "cpu_baseline": {"value": val, "unit": "MB/s", "cores": cores, "kind": "port",
                         "single_thread_value": one,
                         "sample": f"the leading {sample_n >> 20} MiB of the same corpus per step ({cores} cores x "
                                   f"{per >> 20} MiB): " +
                                   ("independent 16 MiB streams decoded by the restated reference decoder on all cores"
                                    if decode else
                                    "block cuts in one sequential pass, then every block encoded block-parallel on all "
                                    "cores (C++ oracle port of the reference encoder; rustc unavailable)") +
                                   f"; single thread on {one_n >> 20} MiB: {one:.2f} MB/s — the reference itself is "
                                   "single-threaded"},
        "e2e": {"value": val, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(sample):
    from oracle import orc
    t = time.perf_counter()
    out = orc.compress(sample, LEVEL)
    dt = time.perf_counter() - t
    return {"value": len(sample) / dt / 1e6, "unit": "MB/s", "cores": 1, "kind": "port",
            "sample": f"first {len(sample) >> 20} MiB of the corpus, level {LEVEL}, single thread (the reference is "
                      f"single-threaded); C++ oracle port, ratio {len(sample) / max(1, len(out)):.3f}"}


def verify_full(corpus, level, stream, gpu_in_off=None):
    """ALL blocks and the trailer of `stream` against the oracle (block cuts from the oracle's own sequential pass, every
    block encoded by the oracle block-parallel on the host cores).  Returns a one-line verdict."""
    from oracle import verify
    import numpy as np
    t = time.perf_counter()
    table = verify.cut_table(corpus, level)
    if gpu_in_off is not None and (len(gpu_in_off) != len(table) or (np.asarray(gpu_in_off, dtype=np.int64) != table).any()):
        return "MISMATCH: block table differs from the oracle's cuts"
    ok, msg, st = verify.verify_stream(corpus, level, stream, table)
    return (msg if ok else "MISMATCH: " + msg) + f" ({time.perf_counter() - t:.1f} s on {host_cores()} cores)"


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
    import numpy as np
    corpus, path = make_corpus(GEN, TOTAL_BYTES, 0, 1)
    raw = np.ascontiguousarray(corpus)
    drop_corpus(path, 1)
    h_raw = torch.from_numpy(raw)
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
    ok_cpu = back == raw[:sample.numel()].tobytes()
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
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
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


# ------------------------------------------------------------------------------------------------ side cases
def timed_ms(fn, steps, torch):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = None
    for _ in range(steps):
        res = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, res


def side_cases(ctx, torch, dv):
    """BASELINE.json configs[1], [3], [4] and the decoder, device-resident on one GPU, every stream fully verified."""
    import numpy as np
    from oracle import orc
    out = {}

    def run(name, data, level, steps=3, warm=1):
        a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        d_in = torch.from_numpy(np.ascontiguousarray(a)).cuda()
        cap = dv.max_output_bytes(level, a.size)
        d_out = torch.zeros(cap, dtype=torch.uint8, device="cuda")

        def step():
            d_out.zero_()
            return ctx.compress_device(level, d_in, d_out)
        for _ in range(warm):
            step()
        ms, n = timed_ms(step, steps, torch)
        st = ctx.sort_stats()
        stream = d_out[:n].cpu().numpy()
        in_off = ctx.block_table(with_crc=False)[0]
        rec = {"level": level, "bytes": int(a.size), "blocks": int(len(in_off) - 1), "ms": round(ms, 3),
               "MB_per_s": round(a.size / ms / 1e3, 1), "compressed_bytes": int(n), "sort_rounds": st["rounds"],
               "radix_passes": st["radix_passes"], "verified": verify_full(a, level, stream, in_off)}
        out[name] = rec
        del d_in, d_out
        return rec

    # configs[1]: a single 900 kB block of text, level 9 (latency of one block: it cannot fill 148 SMs)
    import gen
    run("single_block_text_level9", gen.text(1, 899_876), 9, steps=5, warm=2)
    # configs[4]: adversarial periodic / near-periodic full blocks (deep doubling, equal-rotation tie-break, RLE1 edges)
    run("all_a_one_full_block", np.full(45_899_235, ord("a"), dtype=np.uint8), 9, steps=3)
    run("period2_ab_full_block", b"ab" * 449_990, 9)
    run("period4_aabb_full_block", b"aabb" * 224_995, 9)
    run("near_periodic_abcd_x", b"abcd" * 224_995 + b"x", 9)
    run("near_periodic_ab_c_level1", b"ab" * 49_990 + b"c", 1)
    # configs[3]: level 1 on 1 GiB mixed binary/text (many small blocks)
    mixed, path = make_corpus("mixed", TOTAL_BYTES, 0, 1)
    run("mixed_1gib_level1", mixed, 1, steps=3)
    drop_corpus(path, 1)
    del mixed
    return out


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

    # ---- synthetic corpus (before CUDA / NCCL: the generator forks worker processes)
    t0 = time.time()
    corpus, corpus_file = make_corpus(GEN, TOTAL_BYTES, rank, world)
    gen_s = time.time() - t0

    import numpy as np
    import torch
    import torch.distributed as dist

    import rust_compression_b200  # noqa: F401  (fails loudly if libbzb200.so is missing)
    from rust_compression_b200 import device as dv
    from rust_compression_b200 import sharded

    # rank 0 prints exactly one line on stdout: NCCL's own log (NCCL_DEBUG=INFO/VERSION) goes to stderr, untouched
    if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
        os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    idle = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        idle = dist.new_group(backend="gloo")  # ranks waiting for rank 0's single-process legs sleep on the CPU

    total_bytes = TOTAL_BYTES
    ctx = dv.Context()
    if world == 1:
        d_in = torch.from_numpy(np.array(corpus)).to(dev)
        shard = None
    else:
        shard = sharded.Shard(LEVEL, total_bytes, rank, world, dev)
        if shard.hi > shard.lo:
            shard.slice_view().copy_(torch.from_numpy(np.array(corpus[shard.lo:shard.hi])))
    cap_out = dv.max_output_bytes(LEVEL, total_bytes)
    d_out1 = torch.zeros(cap_out, dtype=torch.uint8, device=dev) if world == 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    walls = []   # host wall time of every step of the most recent timed() call (this rank)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = None
        walls.clear()
        for _ in range(steps):
            tw = time.perf_counter()
            res = fn()   # synchronous on the host side: the wall time of a call is the time of its step on this rank
            walls.append(round(1e3 * (time.perf_counter() - tw), 2))
            if os.environ.get("BZB200_BENCH_DEBUG"):
                sys.stderr.write(f"[rank {rank}] {fn.__name__} wall {walls[-1]:.1f} ms\n")
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), res

    def step_device():
        if world == 1:
            d_out1.zero_()
            n = ctx.compress_device(LEVEL, d_in, d_out1)
            return d_out1[:n], {"nblocks": len(ctx.block_table(with_crc=False)[0]) - 1}
        return sharded.compress_sharded(ctx, shard)

    # ---- device-resident throughput
    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    ms_total, (d_stream, info) = timed(step_device, args.steps)
    launches = ctx.launch_count() - l0
    step_walls = list(walls)
    clocks = sampler.stop() if rank == 0 else None
    sstats = ctx.sort_stats()
    ms_step = ms_total / args.steps
    value = total_bytes / (ms_step / 1e3) / 1e6
    dev_stream = d_stream.cpu().numpy().copy() if rank == 0 else None
    # one more step with a CUDA-event pair around every launch (same stream) for the per-kernel breakdown; it is
    # not part of the timed region because the event traffic slows the host side of the step down
    ctx.profile(True)
    ms_prof, _ = timed(step_device, 1)
    ctx.profile(False)
    recs = ctx.profile_records()
    lsum = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(lsum)

    # ---- end to end with host buffers through the C ABI: rank 0's process, all N GPUs (the other ranks sleep)
    e2e = stream_leg = None
    e2e_stream_bytes = stream_bytes = None
    if world > 1:
        barrier()
    if rank == 0:
        h_in = torch.from_numpy(np.array(corpus)).pin_memory()
        h_out = torch.empty(cap_out, dtype=torch.uint8).pin_memory()
        e2e_steps = max(2, args.steps // 2)
        devices = list(range(world))
        if world == 1:
            def step_e2e():
                return ctx.compress_host(LEVEL, h_in, h_out)
            api = "bzb200_compress_host (C ABI, pinned host in/out, one GPU)"
            pool = None
        else:
            pool = dv.Pool(devices)

            def step_e2e():
                return pool.compress_host(LEVEL, h_in, h_out)
            api = (f"bzb200_pool_compress_host (C ABI, in-library engine: one process, {world} GPUs, one worker thread "
                   "per GPU, pinned host in/out)")
        for _ in range(max(1, args.warmup // 2)):
            step_e2e()
        t = time.perf_counter()
        for _ in range(e2e_steps):
            n_e2e = step_e2e()
        ms_e2e = (time.perf_counter() - t) * 1e3 / e2e_steps   # the call is synchronous: host clock around it
        e2e_stream_bytes = h_out[:n_e2e].numpy().copy()
        e2e = {"value": total_bytes / (ms_e2e / 1e3) / 1e6, "unit": "MB/s", "h2d_bytes_per_step": int(total_bytes),
               "d2h_bytes_per_step": int(n_e2e), "steps": e2e_steps, "ms_per_step": ms_e2e, "api": api,
               "timing": "host clock around the synchronous call (the copies and every GPU's work are inside it)"}
        if pool is not None:
            e2e["engine"] = pool.stats()
            pool.close()
        # the drop-in object: write in 1 MiB pieces (the shim's pattern), finish, read
        from rust_compression_b200 import _lib as lib_mod
        import ctypes as C
        L = lib_mod.lib()

        def step_stream():
            h = C.c_void_p()
            if world == 1:
                rc = L.bzb200_enc_create(LEVEL, local_rank, C.byref(h))
            else:
                arr = (C.c_int * world)(*devices)
                rc = L.bzb200_enc_create_multi(LEVEL, world, arr, C.byref(h))
            assert rc == 0
            return h
        base = h_in.data_ptr()
        piece = 1 << 20

        phase_s = {"write": 0.0, "finish": 0.0, "read": 0.0}

        def drive(h):
            got = 0
            t0 = time.perf_counter()
            for lo in range(0, total_bytes, piece):
                rc = L.bzb200_enc_write(h, C.c_void_p(base + lo), min(piece, total_bytes - lo))
                assert rc == 0, L.bzb200_enc_last_error(h)
                if L.bzb200_enc_output_size(h):  # blocks that have closed are taken out while the input is still coming
                    got += L.bzb200_enc_read(h, C.c_void_p(h_out.data_ptr() + got), h_out.numel() - got)
            t1 = time.perf_counter()
            rc = L.bzb200_enc_finish(h)
            assert rc == 0, L.bzb200_enc_last_error(h)
            t2 = time.perf_counter()
            while True:
                k = L.bzb200_enc_read(h, C.c_void_p(h_out.data_ptr() + got), h_out.numel() - got)
                if k == 0:
                    break
                got += k
            t3 = time.perf_counter()
            phase_s["write"], phase_s["finish"], phase_s["read"] = t1 - t0, t2 - t1, t3 - t2
            L.bzb200_enc_reset(h)
            return got
        h = step_stream()
        drive(h)  # warm-up: allocates the pinned windows and the pool
        t = time.perf_counter()
        s_steps = max(2, args.steps // 2)
        for _ in range(s_steps):
            n_s = drive(h)
        ms_s = (time.perf_counter() - t) * 1e3 / s_steps
        stream_bytes = h_out[:n_s].numpy().copy()
        st4 = (C.c_uint64 * 4)()
        L.bzb200_enc_destroy(h)
        stream_leg = {"value": total_bytes / (ms_s / 1e3) / 1e6, "unit": "MB/s", "ms_per_step": ms_s, "steps": s_steps,
                      "api": "bzb200_enc_write x %d (1 MiB pieces) + bzb200_enc_finish + bzb200_enc_read, %d GPU(s)"
                             % ((total_bytes + piece - 1) // piece, world),
                      "window_bytes": int(os.environ.get("BZB200_ENC_WINDOW", str((256 << 20) * world))),
                      "last_step_ms": {k: round(v * 1e3, 1) for k, v in phase_s.items()},
                      "pattern": "write 1 MiB, take out whatever has become readable, ...; finish; read the rest"}
    if world > 1:
        dist.barrier(group=idle)

    if rank == 0:
        out_bytes = int(dev_stream.size)
        verified = verify_full(corpus, LEVEL, dev_stream)
        same_e2e = bool(e2e_stream_bytes.size == dev_stream.size and (e2e_stream_bytes == dev_stream).all())
        same_stream = bool(stream_bytes.size == dev_stream.size and (stream_bytes == dev_stream).all())
        verified = ("device stream: " + verified + "; e2e stream " + ("identical" if same_e2e else "DIFFERS (MISMATCH)") +
                    "; e2e_stream (drop-in object) stream " + ("identical" if same_stream else "DIFFERS (MISMATCH)"))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        # dominant kernel by device time inside the profiled step (rank 0's blocks)
        top = sorted(recs.items(), key=lambda kv: -kv[1][1])
        name, (nl, kms) = top[0] if top else ("none", (0, 0.0))
        step_ms_kernels = sum(v[1] for v in recs.values())
        if name in ALG_BYTES:
            alg_bytes = ALG_BYTES[name](sstats)
        else:
            alg_bytes = total_bytes / world + out_bytes / world
        achieved = alg_bytes / (kms / 1e3) / 1e9 if kms > 0 else 0.0
        ratio, ncu_name, ratio_src = ncu_traffic_ratio(name)
        roofline = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak,
                    "frac_is": "per-kernel model: algorithmic bytes of the dominant kernel's launches / their summed "
                               "CUDA-event time / measured copy peak (NOT the SURVEY 8(d) whole-path figure: that is "
                               "roofline.path)",
                    "traffic": (ratio * alg_bytes / max(1, nl)) if ratio else None,
                    "traffic_source": (f"profiles/r2_ncu_kernels.csv row {ncu_name}: measured (dram_read_bytes + "
                                       f"dram_write_bytes) / alg_bytes of that launch = {ratio:.3f}, x the algorithmic "
                                       "bytes per launch of this step") if ratio else
                                      "no ncu row for this kernel in " + ratio_src,
                    "peak_source": peak_src,
                    "launches_per_step": nl, "avg_launch_ms": kms / max(1, nl),
                    "kernel_share_of_step": kms / ms_prof,
                    "algorithmic_bytes_per_launch": alg_bytes / max(1, nl),
                    "note": "achieved = algorithmic bytes of the kernel's launches in one step / their summed CUDA-event "
                            "time (profiled step, events on the launching stream)",
                    "path": {"achieved": (total_bytes + out_bytes) / (ms_step / 1e3) / 1e9 / world, "unit": "GB/s per GPU",
                             "frac": (total_bytes + out_bytes) / (ms_step / 1e3) / 1e9 / world / peak,
                             "note": "whole path: (input + output bytes) / device time (SURVEY.md 8(d))"}}
        side = {}
        if SIDE and world == 1:
            try:
                side = side_cases(ctx, torch, dv)
            except Exception as ex:  # a side case must never cost the headline line
                side = {"error": repr(ex)}
        base1 = cpu_baseline(np.ascontiguousarray(corpus[:CPU_SAMPLE_BYTES]))
        line = {
            "metric": METRIC, "value": value, "unit": "MB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(world), "level": LEVEL, "total_bytes": total_bytes,
                       "bytes_per_gpu": total_bytes // world, "blocks": info["nblocks"], "compressed_bytes": out_bytes,
                       "ratio": total_bytes / max(1, out_bytes),
                       "l2": f"every rank's slice ({total_bytes // world >> 20} MiB) and its sort arrays (~30 B per "
                             "input byte) exceed the 126 MB L2; no explicit flush",
                       "sort": sstats, "verified": verified, "corpus_gen_s": round(gen_s, 1),
                       "side_cases": side},
            "e2e": e2e,
            "e2e_stream": stream_leg,
            "gpu_launches": int(lsum.item()),
            "roofline": roofline,
            "kernels_ms_per_step": {k: round(v[1], 3) for k, v in top[:14]},
            "profiled_step_ms": round(ms_prof, 3),
            "step_wall_ms_rank0": step_walls,
            "kernel_time_ms_per_step": round(step_ms_kernels, 3),
            "cpu_baseline": base1,
            "speedup_vs_single_thread_reference": {"device": value / base1["value"], "e2e": e2e["value"] / base1["value"]},
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        drop_corpus(corpus_file, world)


if __name__ == "__main__":
    main()
