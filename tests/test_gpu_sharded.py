"""GPU tests of the multi-GPU data plane: the sliced K1 plan (include/bzb200.h 2b), the
in-library engine (2c, bzb200_pool_* and bzb200_enc_create_multi) and the one-process-per-GPU path over real NCCL
(rust-compression_b200/sharded.py).  Every result is compared with the oracle's one-pass stream bit for bit.

On a one-GPU box the engine runs with several contexts on the same GPU (BZB200_MG_CTX_PER_GPU): every context holds
only its slice, so slice bounds, carried run heads, emitted offsets, halo, window rows, block tails and the host join
are exercised exactly as on several GPUs.  The NCCL test needs two GPUs and is skipped otherwise."""
import os
import sys

import numpy as np
import pytest

import gen
from oracle import orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cases():
    return [
        ("mixed level 1", gen.mixed(4, 3_000_000), 1),
        ("runs across slice boundaries", b"x" * 2_500_000 + gen.g2(3, 900_000) + b"y" * 5_000_000 + gen.text(2, 700_000), 1),
        ("aaaab: several plan phases", b"aaaab" * 2_400_000, 1),
        ("text level 9", gen.text(5, 8_000_000), 9),
        ("fewer blocks than workers", gen.text(9, 150_000), 2),
        ("one byte", b"q", 9),
        ("all a, one full level-9 block and a bit", b"a" * 46_000_000, 9),
    ]


@pytest.mark.parametrize("per_gpu", [1, 2, 3, 5])
def test_pool_slices_on_one_gpu(monkeypatch, per_gpu):
    import torch
    from rust_compression_b200 import device as dv
    monkeypatch.setenv("BZB200_MG_CTX_PER_GPU", str(per_gpu))
    pool = dv.Pool([0])
    assert pool.size() == per_gpu
    for name, data, level in _cases():
        want = orc.compress(data, level)
        h_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).pin_memory()
        h_out = torch.empty(dv.max_output_bytes(level, len(data)), dtype=torch.uint8).pin_memory()
        for rep in range(2):  # the second run reuses every buffer of the pool
            n = pool.compress_host(level, h_in, h_out)
            assert h_out[:n].numpy().tobytes() == want, f"{name}: {per_gpu} contexts, run {rep}"
    st = pool.stats()
    assert st["spans"] == 2 * len(_cases())
    if per_gpu > 1:
        assert st["phases"] > st["spans"]  # aaaab needs more than one phase of the sliced cut chain
    # empty input and a too-small output buffer
    h_out = torch.empty(64, dtype=torch.uint8)
    assert pool.compress_host(9, torch.empty(0, dtype=torch.uint8), h_out) == 14
    assert h_out[:14].numpy().tobytes() == orc.compress(b"", 9)
    with pytest.raises(Exception):
        pool.compress_host(1, torch.frombuffer(bytearray(gen.text(1, 500_000)), dtype=torch.uint8), h_out)
    # ... and the pool still works afterwards
    d = gen.text(1, 500_000)
    big = torch.empty(dv.max_output_bytes(1, len(d)), dtype=torch.uint8)
    n = pool.compress_host(1, torch.frombuffer(bytearray(d), dtype=torch.uint8), big)
    assert big[:n].numpy().tobytes() == orc.compress(d, 1)
    pool.close()


def test_multi_gpu_encoder_object_streams(monkeypatch):
    """bzb200_enc_create_multi: the drop-in object over the engine, windows smaller than the input, pieces that do not
    line up with anything; the worker thread compresses a window while the next one is written."""
    import rust_compression_b200 as rc
    monkeypatch.setenv("BZB200_MG_CTX_PER_GPU", "2")
    data = gen.mixed(5, 2_200_000) + b"q" * 400_000 + gen.text(6, 1_500_000)
    for level, window, piece in ((1, 600_000, 170_001), (2, 1, 333_333), (9, 2_000_000, 1 << 19)):
        monkeypatch.setenv("BZB200_ENC_WINDOW", str(window))
        enc = rc.BZip2Encoder(level, devices=[0])
        got = bytearray()
        for lo in range(0, len(data), piece):
            enc.write(data[lo:lo + piece])
            got += enc.read_available()
        st = enc.stats()
        got += enc.finish()
        assert bytes(got) == orc.compress(data, level), (level, window, piece)
        assert st["windows"] >= 1
        again = bytes(rc.encode(b"second", enc, rc.Action.Finish))
        assert again == orc.compress(b"second", level)
    monkeypatch.delenv("BZB200_ENC_WINDOW")
    enc = rc.BZip2Encoder(9, devices=[0])
    assert bytes(rc.encode(b"", enc, rc.Action.Finish)) == orc.compress(b"", 9)
    assert bytes(rc.encode(b"a\n", enc, rc.Action.Finish)) == orc.compress(b"a\n", 9)


def test_streaming_encoder_overlaps_and_matches_one_shot(monkeypatch):
    """SURVEY.md section 8(f).2 on the single-GPU object: many 1 MiB writes (the Rust shim's pattern) with 8 MiB windows;
    blocks become readable while input is still being written, and the bytes equal the one-shot stream."""
    import rust_compression_b200 as rc
    monkeypatch.setenv("BZB200_ENC_WINDOW", str(8 << 20))
    data = gen.text(3, 40 << 20)
    want = orc.compress(data, 9)
    enc = rc.BZip2Encoder(9)
    got = bytearray()
    early = 0
    for lo in range(0, len(data), 1 << 20):
        enc.write(data[lo:lo + (1 << 20)])
        got += enc.read_available()
        early = len(got)
    got += enc.finish()
    assert bytes(got) == want
    assert early > len(want) // 2


def _nccl_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import rust_compression_b200  # noqa: F401
        from rust_compression_b200 import device as dv
        from rust_compression_b200 import sharded
        ctx = dv.Context()
        res = []
        for name, data, level in _cases():
            if len(data) < 2:
                continue
            sh = sharded.Shard(level, len(data), rank, world, dev)
            if sh.hi > sh.lo:
                sh.slice_view().copy_(torch.frombuffer(bytearray(data[sh.lo:sh.hi]), dtype=torch.uint8))
            stream, info = sharded.compress_sharded(ctx, sh)
            res.append((name, stream.cpu().numpy().tobytes() if stream is not None else None, info))
        q.put((rank, res))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharded_over_nccl_two_ranks():
    """The one-process-per-GPU path with the real device.Context on two GPUs: block ranges meet at a rank boundary,
    the tail of rank 0's last block comes over NVLink, payloads are joined at an unaligned bit."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() * 11) % 2000
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        rank, res = q.get(timeout=900)
        got[rank] = res
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    wants = {name: orc.compress(data, level) for name, data, level in _cases() if len(data) >= 2}
    for (name, stream, info), (_, s1, i1) in zip(got[0], got[1]):
        assert stream == wants[name], f"{name}: joined stream differs from the oracle"
        assert s1 is None and info["b1"] == i1["b0"] and i1["b1"] == info["nblocks"]


def test_pool_over_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from rust_compression_b200 import device as dv
    # kernel attributes are per device: the process has used GPU 0 alone before the pool drives GPU 1 as well
    warm = dv.Context(0)
    d = gen.text(4, 2_000_000)
    assert dv.compress_tensor(warm, 9, torch.frombuffer(bytearray(d), dtype=torch.uint8).cuda(0)).cpu().numpy().tobytes() \
        == orc.compress(d, 9)
    warm.close()
    pool = dv.Pool([0, 1])
    for name, data, level in _cases():
        want = orc.compress(data, level)
        h_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).pin_memory()
        h_out = torch.empty(dv.max_output_bytes(level, len(data)), dtype=torch.uint8).pin_memory()
        n = pool.compress_host(level, h_in, h_out)
        assert h_out[:n].numpy().tobytes() == want, name
    pool.close()
