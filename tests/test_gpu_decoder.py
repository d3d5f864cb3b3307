"""GPU parity tests of the decoder (run with -m gpu): the CUDA decoder, called through the C ABI, against the restated
reference BZip2Decoder (oracle/bz2_decoder_oracle.cpp) — same bytes, and for malformed buffers the same bytes before
the error and the same BZip2Error kind — plus the reference's own decoder tests (bzip2/mod.rs:84-172) restated on the
mirrored API.  Nothing here reads /root/reference."""
import bz2
import os

import numpy as np
import pytest

import dec_cases
import gen
from oracle import orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rc():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import rust_compression_b200 as m
    return m


def gpu_decode(rc, buf):
    try:
        return 0, rc.decompress(buf)
    except rc.BZip2Error as e:
        code = {v: k for k, v in rc.BZip2Error.KINDS.items()}[e.kind]
        return code, e.partial


# ---- the reference's own decoder tests, restated ----

@pytest.mark.parametrize("idx", [1, 2, 3, 4])
def test_check_unzip_samples(rc, sample_data, idx):
    """bzip2/mod.rs:72-148 check_unzip(sampleN.bz2, sampleN.ref) through DecodeExt::decode."""
    with open(os.path.join(dec_cases.DATA, f"sample{idx}.bz2"), "rb") as f:
        actual = f.read()
    ret = bytes(rc.decode(actual, rc.BZip2Decoder()))
    assert ret == sample_data[idx]


def test_unit_and_long_round_trip(rc):
    """bzip2/mod.rs:41-70,150-172: encode on the GPU, decode on the GPU."""
    for data in (b"a\n", b"a" * 1000, b"aabbaabbaabbaabb\n"):
        comp = bytes(rc.encode(data, rc.BZip2Encoder(9), rc.Action.Finish))
        assert bytes(rc.decode(comp, rc.BZip2Decoder())) == data


def test_decoder_object_reuse_and_error_after_bytes(rc):
    dec = rc.BZip2Decoder()
    s = orc.compress(gen.text(3, 120000), 1)
    assert dec.decode_all(s) == gen.text(3, 120000)
    assert dec.decode_all(orc.compress(b"", 9)) == b""
    bad = s[:len(s) // 2]
    it = rc.decode(bad, dec)
    got = bytearray()
    with pytest.raises(rc.BZip2Error) as ei:
        for b in it:
            got.append(b)
    want = dec_cases.expected(bad)
    assert bytes(got) == want[1] and ei.value.kind == orc.DecodeError.KINDS[want[0]]
    assert ei.value.to_compression_error().kind == "DataError"
    assert dec.decode_all(s) == gen.text(3, 120000)  # usable again after an error


# ---- parity with the restated reference decoder ----

@pytest.mark.parametrize("split", ["0", "1"])   # 0: fused d2_decode, 1: d2_huff + chunk-parallel d2_mtf_a/b/c
def test_valid_streams(rc, monkeypatch, split):
    monkeypatch.setenv("BZB200_DEC_SPLIT", split)
    for name, buf in dec_cases.valid_cases(big=True):
        want = dec_cases.expected(buf)
        assert want[0] == 0, name
        got = gpu_decode(rc, buf)
        assert got[0] == 0, f"{name}: error {got[0]}"
        assert got[1] == want[1], f"{name}: bytes differ (gpu {len(got[1])}, reference {len(want[1])})"


@pytest.mark.parametrize("split", ["0", "1"])
def test_malformed_streams_report_what_the_reference_reports(rc, monkeypatch, split):
    monkeypatch.setenv("BZB200_DEC_SPLIT", split)
    for name, buf in dec_cases.malformed_cases():
        want = dec_cases.expected(buf)
        got = gpu_decode(rc, buf)
        assert dec_cases.same_result(got, want), \
            f"{name}: kind {got[0]}, {len(got[1])} bytes before the error; reference kind {want[0]}, {len(want[1])} bytes"


def test_fuzzed_streams(rc):
    """Random mutations of valid streams (the generator of the CPU emulation test): bytes and error kind as the
    restated reference decoder reports them."""
    for name, buf in dec_cases.fuzz_cases(120, 7):
        want = dec_cases.expected(buf)
        got = gpu_decode(rc, buf)
        assert dec_cases.same_result(got, want), f"{name}: got {got[0]}/{len(got[1])} bytes, reference {want[0]}/{len(want[1])}"


def test_device_api_batches_and_small_output(rc, monkeypatch):
    import torch
    from rust_compression_b200 import device as dv
    data = gen.mixed(1, 1_500_000)
    buf = orc.compress(data, 1)
    d_in = torch.frombuffer(bytearray(buf), dtype=torch.uint8).cuda()
    ctx = dv.Context()
    d_small = torch.empty(1000, dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError) as ei:
        ctx.decompress_device(d_in, d_small)
    assert int(ei.value.args[0]) == len(data)
    d_out = torch.empty(len(data), dtype=torch.uint8, device="cuda")
    n, kind = ctx.decompress_device(d_in, d_out)
    assert (n, kind) == (len(data), 0) and d_out.cpu().numpy().tobytes() == data
    st = ctx.dec_stats()
    assert st["streams"] == 1 and st["blocks"] == 14 and st["batches"] == 1 and st["out_bytes"] == len(data)
    monkeypatch.setenv("BZB200_DEC_BATCH_BYTES", "2500000")   # a few blocks per batch
    d_out.zero_()
    n, kind = ctx.decompress_device(d_in, d_out)
    assert (n, kind) == (len(data), 0) and d_out.cpu().numpy().tobytes() == data
    assert ctx.dec_stats()["batches"] > 3
    # an input view at an odd device address (byte-wise loads instead of aligned words)
    d_odd = torch.empty(len(buf) + 1, dtype=torch.uint8, device="cuda")
    d_odd[1:] = d_in
    d_out.zero_()
    n, kind = ctx.decompress_device(d_odd[1:], d_out)
    assert (n, kind) == (len(data), 0) and d_out.cpu().numpy().tobytes() == data
    ctx.close()


@pytest.mark.parametrize("split", ["0", "1"])
def test_gpu_encoder_streams_decode_on_gpu(rc, monkeypatch, split):
    monkeypatch.setenv("BZB200_DEC_SPLIT", split)
    _round_trips(rc)


def _round_trips(rc):
    """Encoder -> decoder round trips at sizes the oracle does not need to see: levels 1 and 9, multi-block, plus the
    adversarial periodic inputs (cycle walks of the inverse BWT)."""
    cases = [(gen.text(21, 5_000_000), 9), (gen.mixed(4, 3_000_000), 1), (b"ab" * 1_000_000, 9),
             (b"aabb" * 500_000 + b"z", 5), (b"a" * 10_000_000, 9), (gen.g2(9, 2_000_000), 9)]
    for data, level in cases:
        comp = rc.compress(data, level)
        assert bz2.decompress(comp) == data
        assert rc.decompress(comp) == data


def test_large_corpus_round_trip_host_api(rc):
    """256 MiB of text at level 9 (about 300 blocks) through bzb200_compress_host and bzb200_decompress_host; the
    decoder's output is compared with the input on the device and through the stream's own CRCs."""
    import torch
    from rust_compression_b200 import device as dv
    n = 256 << 20
    data = gen.text(1, n)
    h_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).pin_memory()
    ctx = dv.Context()
    h_comp = torch.empty(dv.max_output_bytes(9, n), dtype=torch.uint8).pin_memory()
    m = ctx.compress_host(9, h_in, h_comp)
    h_back = torch.empty(n, dtype=torch.uint8).pin_memory()
    got, kind = ctx.decompress_host(h_comp[:m], h_back)
    assert (got, kind) == (n, 0)
    assert torch.equal(h_back, h_in)
    st = ctx.dec_stats()
    assert st["streams"] == 1 and st["blocks"] >= 290 and st["candidates"] == st["blocks"] + 1
    ctx.close()
