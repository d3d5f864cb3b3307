"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI, against the oracle
on the same inputs — bit-exact — plus the reference's own bzip2 tests restated against the mirrored API.
Nothing here reads /root/reference; inputs come from tests/golden/ and the deterministic generators."""
import bz2
import hashlib
import os

import numpy as np
import pytest

import gen
import parity
from oracle import orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rc():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import rust_compression_b200 as m
    return m


# ---- the reference's own tests (src/bzip2/mod.rs:41-172, src/lib.rs:13-33), restated on the mirrored API ----

def test_unit(rc, golden):
    """bzip2/mod.rs:41-70 test_unit."""
    ret = bytes(rc.encode(b"a\n", rc.BZip2Encoder(9), rc.Action.Finish))
    assert ret == bytes.fromhex(golden["test_unit"]["output"])
    assert bz2.decompress(ret) == b"a\n"


@pytest.mark.parametrize("idx,level", [(1, 1), (2, 2), (3, 3)])
def test_sample_levels(rc, sample_data, idx, level):
    """bzip2/mod.rs:84-139 test_sample1/2/3: encode at level 1/2/3, decode == input; here also bit-exact."""
    encoder = rc.BZip2Encoder(level)
    encoder.write(sample_data[idx])
    ret = encoder.finish()
    assert bz2.decompress(ret) == sample_data[idx]
    assert ret == orc.compress(sample_data[idx], level)


def test_long(rc):
    """bzip2/mod.rs:150-172 test_long."""
    data = b"a" * 1000
    compressed = bytes(rc.encode(data, rc.BZip2Encoder(9), rc.Action.Finish))
    assert bz2.decompress(compressed) == data
    assert compressed == orc.compress(data, 9)


def test_readme_doctest(rc):
    """lib.rs:13-33."""
    data = b"aabbaabbaabbaabb\n"
    compressed = bytes(rc.encode(data, rc.BZip2Encoder(9), rc.Action.Finish))
    assert compressed.hex() == ("425a68393141592653597e6ce699000002410000103000200030934c154da91a231e2ee48a70a120"
                                "fcd9cd32")


# ---- API semantics ----

def test_invalid_level(rc):
    for lv in (0, 10, -1):
        with pytest.raises(ValueError):
            rc.BZip2Encoder(lv)


def test_run_then_finish_and_reuse(rc):
    enc = rc.BZip2Encoder(9)
    data = gen.text(4, 50000)
    it1 = iter(data[:20000])
    assert enc.next(it1, rc.Action.Run) is None          # input drained, nothing emitted yet
    it2 = iter(data[20000:])
    out = []
    while True:
        b = enc.next(it2, rc.Action.Finish)
        if b is None:
            break
        out.append(b)
    assert bytes(out) == orc.compress(data, 9)
    # the encoder re-arms after returning None (encoder.rs:87-90,130-133)
    again = bytes(rc.encode(b"a\n", enc, rc.Action.Finish))
    assert again == orc.compress(b"a\n", 9)


def test_run_streams_closed_blocks_before_finish(rc, monkeypatch):
    """with Action.Run the encoder compresses the blocks that have closed whenever a window of
    input has accumulated and hands their bytes out early; the concatenation is the oracle's stream bit for bit —
    including the partial byte carried from window to window, a run that straddles a window edge, and multi-stream
    reuse of the same encoder object."""
    for level, window, piece in ((1, 250_000, 70_001), (1, 99_981, 33_333), (2, 1, 500_000), (9, 1_000_000, 1 << 18)):
        monkeypatch.setenv("BZB200_ENC_WINDOW", str(window))
        enc = rc.BZip2Encoder(level)
        monkeypatch.delenv("BZB200_ENC_WINDOW")
        data = gen.mixed(5, 1_200_000) + b"q" * 300_000 + gen.text(6, 900_000)
        got = bytearray()
        early = 0
        for lo in range(0, len(data), piece):
            enc.write(data[lo:lo + piece])
            got += enc.read_available()
            early = len(got)
        st = enc.stats()
        got += enc.finish()
        assert bytes(got) == orc.compress(data, level), (level, window, piece)
        assert early > 0 and st["windows"] >= 2, (level, window, st)
        # the same object starts a second, independent stream (multi-stream container: SURVEY.md section 8(f).3)
        second = bytes(rc.encode(b"second stream", enc, rc.Action.Finish))
        assert second == orc.compress(b"second stream", level)
        assert rc.decompress(bytes(got) + second) == data + b"second stream"
    # through the iterator adapter: bytes appear under Run, None means "feed more"
    monkeypatch.setenv("BZB200_ENC_WINDOW", "150000")
    enc = rc.BZip2Encoder(1)
    data = gen.text(8, 400_000)
    out = bytearray()
    for lo in range(0, len(data), 100_000):
        it = iter(data[lo:lo + 100_000])
        while (b := enc.next(it, rc.Action.Run)) is not None:
            out.append(b)
    assert len(out) > 0
    it = iter(b"")
    while (b := enc.next(it, rc.Action.Finish)) is not None:
        out.append(b)
    assert bytes(out) == orc.compress(data, 1)


def test_empty_input(rc):
    assert rc.compress(b"", 9).hex() == "425a683917724538509000000000"
    assert bytes(rc.encode(b"", rc.BZip2Encoder(1), rc.Action.Finish)) == orc.compress(b"", 1)


def test_one_shot_c_abi(rc):
    d = gen.mixed(3, 300000)
    assert rc.compress(d, 5) == orc.compress(d, 5)


# ---- stage-by-stage parity on the reference fixtures and generators ----

@pytest.mark.parametrize("idx", [1, 2, 3, 4, 5, 6, 7])
def test_samples_level9_stages(sample_data, idx):
    parity.assert_parity(sample_data[idx], 9)


APPB = [b"a", b"aa", b"a" * 4, b"a" * 5, b"aaaa\x00", b"a" * 255 + b"b", b"a" * 256, b"ab" * 500, b"aabb" * 300,
        b"abcd" * 64 + b"e", bytes(range(256)), bytes(range(256)) * 3, b"\x00", b"\xff" * 7]


@pytest.mark.parametrize("i", range(len(APPB)))
def test_small_cases_stages(i):
    parity.assert_parity(APPB[i], 9)


def test_generators_multiblock():
    parity.assert_parity(gen.g1(1, 250000), 1)      # 3 blocks, App. B
    parity.assert_parity(gen.g2(2, 4000000), 1)     # 255-splits, count bytes, T..T+4 slack
    parity.assert_parity(gen.mixed(1, 1500000), 1)  # mixed binary/text, many small blocks


def test_config2_single_text_block():
    """BASELINE config 2: a single 900 kB block of synthetic text at level 9.  SURVEY.md section 8(d) writes it as
    TEXT(seed=1, n=899 900) expecting RLE1 to add fewer than 81 bytes; with this generator RLE1 adds 105 (the block
    closes after 899 876 input bytes), so the single-block input is the first 899 876 bytes — the whole first block of
    TEXT(1, 899 900) — and the 899 900-byte input itself is the two-block case (a full block plus a 24-byte one)."""
    data = gen.text(1, 899_900)
    r = orc.Run(data, 9)
    assert r.nblocks == 2 and r.info(0)["in_end"] == 899_876 and r.info(0)["nblock"] == 899_981
    r.close()
    r = orc.Run(data[:899_876], 9)
    assert r.nblocks == 1
    r.close()
    parity.assert_parity(data[:899_876], 9)
    parity.assert_parity(data, 9)


@pytest.mark.parametrize("level", [1, 9])
def test_ragged_sizes_around_block_limit(level):
    T = level * 100000 - 19
    base = gen.text(21, T + 200)
    for n in (T - 1, T, T + 1, T + 5):
        parity.assert_parity(base[:n], level, keep_sa=False)


@pytest.mark.parametrize("unit,tail", [(b"a", b""), (b"ab", b""), (b"aabb", b""), (b"abcd", b""), (b"ab", b"c"),
                                       (b"aabb", b"a"), (b"abcd", b"x"), (b"a", b"b")])
@pytest.mark.parametrize("level", [1, 9])
def test_adversarial_periodic(unit, tail, level):
    """BASELINE config 5: periodic and near-periodic blocks (deep doubling, equal-rotation tie-break)."""
    n = level * 100000 + 5000  # a full block plus a little
    reps = 460 * n // 1000 if unit == b"a" and False else n // len(unit)
    data = unit * reps + tail
    parity.assert_parity(data, level, keep_sa=(level == 1))


def test_all_a_fills_exactly_one_block():
    # 'a' x 45 899 235 = 179 997 runs of 255 -> one level-9 block of 899 985 bytes
    data = b"a" * 45_899_235
    res = parity.compare(data, 9, orc, keep_sa=False)
    assert not {k: v for k, v in res.items() if v}, res


def test_cut_chain_drift_leaves_window():
    """'aaaab' repeated: every cut lands inside a 5-byte piece, so every block is 3 bytes longer than T and the drift
    of the cut positions leaves k1_cut_windows' 256-offset window several times (multi-phase cut chain)."""
    data = b"aaaab" * 3_000_000
    res = parity.compare(data, 1, orc, keep_sa=False, check_blocks=[0, 1, 84, 85, 86, 87, 170, 179, 180])
    assert not {k: v for k, v in res.items() if v}, res


def test_cut_chain_long_runs_and_last_piece():
    """Runs far longer than a tile (cut candidates many tiles apart) and an input whose last piece is the one that
    reaches T (no cut there: encoder.rs:729-739)."""
    parity.assert_parity(b"x" * 3_000_000 + gen.text(5, 150000) + b"y" * 2_000_000, 1, keep_sa=False)
    T = 99981
    d = gen.text(6, T - 3) + b"zzzz"          # the run of 4 'z' ends the input and crosses T
    parity.assert_parity(d, 1, keep_sa=False)


def test_doubling_rounds_group_sizes():
    """Group-size classes of the rotation sort: enumeration (<= 16), bucket split (17..1535), BIG groups (radix
    path), sparse blocks, and a block that needs many rounds."""
    rng = np.random.default_rng(7)
    words = [bytes(rng.integers(97, 101, size=int(rng.integers(3, 9))).tolist()) for _ in range(40)]
    d1 = b" ".join(words[int(i)] for i in rng.integers(0, 40, size=40000))            # few distinct words: BIG + buckets
    parity.assert_parity(d1, 9)
    d2 = (gen.text(9, 30000) * 12)[:350000]                                            # long repeats: many rounds
    parity.assert_parity(d2, 9)
    d3 = bytes(rng.integers(0, 4, size=300000, dtype=np.uint8))                        # 4-letter alphabet
    parity.assert_parity(d3, 9)


def test_sliced_plan_equals_plan():
    """include/bzb200.h section 2b driven by hand for three slices held by three contexts on this GPU (what
    sharded.py does per rank and the engine of 2c per GPU): same block table as bzb200_plan, and the blocks encode
    identically from a slice."""
    import torch
    from rust_compression_b200 import device as dv
    from rust_compression_b200 import sharded
    data = gen.g2(5, 3_000_000) + gen.text(8, 700_000)
    n, level = len(data), 1
    d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    a = dv.Context()
    nb = a.plan(level, d_in)
    want = a.block_table(with_crc=False)
    halo, W, T = dv.slice_halo_bytes(), dv.cut_window(), level * 100000 - 19
    bounds = sharded.slice_bounds(n, 3)
    ctxs, bufs = [], []
    heads = []
    for lo, hi in bounds:
        c = dv.Context()
        reserve = min(n, hi + sharded.tail_reserve(level))
        buf = torch.zeros(sharded.LEFT + reserve - lo + 64, dtype=torch.uint8, device="cuda")
        left = 16 if lo else 0
        avail = min(n, hi + halo)
        buf[sharded.LEFT - left:sharded.LEFT + avail - lo] = d_in[lo - left:avail]
        heads.append(c.slice_begin(level, n, lo, hi, buf, sharded.LEFT, avail, reserve))
        ctxs.append(c)
        bufs.append(buf)
    emitted = [c.slice_counts(max([-1] + heads[:r])) for r, c in enumerate(ctxs)]
    for r, c in enumerate(ctxs):
        c.slice_prefix(sum(emitted[:r]), sum(emitted))
    e_tot = sum(emitted)
    max_blocks = (n + n // 4 + 64) // T + 2
    state = np.zeros(4, dtype=np.uint64)
    in_off = np.zeros(max_blocks + 1, dtype=np.uint64)
    rle_off = np.zeros(max_blocks + 1, dtype=np.uint64)
    while not state[2]:
        x0 = int(state[1])
        K = (e_tot - x0) // T if e_tot >= x0 + T else 0
        F = np.zeros((K, W), dtype=np.uint64)
        rows = torch.zeros((max(K, 1), W), dtype=torch.int64, device="cuda")
        for c in ctxs:
            j0, nj = c.slice_windows(x0, rows)
            c.sync()
            if nj:
                F[j0:j0 + nj] = rows[:nj].cpu().numpy().view(np.uint64)
        got_nb, ml = dv.cut_walk(F, T, e_tot, n, state, in_off, rle_off)
    assert got_nb == nb
    assert (in_off[:nb + 1] == want[0]).all() and (rle_off[:nb + 1] == want[1]).all()
    # the middle slice encodes its blocks from its resident bytes (+ the tail of its last block)
    c, (lo, hi), buf = ctxs[1], bounds[1], bufs[1]
    c.slice_set_blocks(in_off[:nb + 1], rle_off[:nb + 1], ml)
    b0, b1, need = c.slice_blocks()
    assert b1 > b0 and int(in_off[b0]) >= lo and int(in_off[b1 - 1]) < hi
    avail = min(n, hi + halo)
    if need > avail:
        buf[sharded.LEFT + avail - lo:sharded.LEFT + need - lo] = d_in[avail:need]
        c.slice_extend(need)
    cap = dv.max_output_bytes(level, n)
    o1 = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    o2 = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    e1 = a.encode_blocks(b0, b1, o1, 5)
    e2 = c.encode_blocks(b0, b1, o2, 5)
    assert e1 == e2 and torch.equal(o1, o2)
    assert (a.block_table(with_crc=False)[2][b0:b1] == c.block_table(with_crc=False)[2][b0:b1]).all()
    for x in ctxs + [a]:
        x.close()


def test_host_path_segmented_pipeline(monkeypatch):
    """bzb200_compress_host with the input copied in segments and every segment planned from the previous block cut
    (H2D / compute / D2H overlap): the stream must not depend on the segment size.  g2 has long runs, so cuts fall
    inside maximal runs and segments start at unaligned offsets."""
    import torch
    from rust_compression_b200 import device as dv
    data = gen.g2(11, 5_000_000) + gen.text(12, 3_000_000) + b"q" * 700_000 + gen.mixed(13, 2_000_000)
    want = orc.compress(data, 1)
    h_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).pin_memory()
    h_out = torch.empty(dv.max_output_bytes(1, len(data)), dtype=torch.uint8).pin_memory()
    for seg in (1 << 20, 3_333_333):
        monkeypatch.setenv("BZB200_HOST_SEGMENT", str(seg))
        ctx = dv.Context()
        n = ctx.compress_host(1, h_in, h_out)
        assert h_out[:n].numpy().tobytes() == want, f"segment size {seg}"
        ctx.close()
    monkeypatch.delenv("BZB200_HOST_SEGMENT")


def test_multi_batch_equals_single_batch(monkeypatch):
    d = gen.mixed(5, 1200000)
    monkeypatch.setenv("BZB200_BATCH_ELEMS", "250000")
    parity.assert_parity(d, 1, keep_sa=False)   # last batch's stages + whole stream
    monkeypatch.delenv("BZB200_BATCH_ELEMS")


def test_package_merge_fallback_is_exercised():
    """A skewed block drives plain Huffman depths past 17 so the reverse package-merge path runs on the GPU."""
    data = gen.geometric(1, 880000)
    r = orc.Run(data, 9)
    used = r.info(0)["lm_used"]
    r.close()
    res = parity.compare(data, 9, orc)
    assert not {k: v for k, v in res.items() if v}, res
    assert used > 0, "input no longer triggers the package-merge path; pick a more skewed one"


# ---- full-size properties (BASELINE config 3 shape, size-independent checks) ----

def test_large_corpus_properties():
    import torch
    from rust_compression_b200 import device as dv
    n = int(os.environ.get("BZB200_TEST_LARGE_BYTES", str(256 << 20)))
    data = gen.text(2, n)
    ctx = dv.Context()
    d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    stream = dv.compress_tensor(ctx, 9, d_in).cpu().numpy().tobytes()
    in_off, rle_off, crc = ctx.block_table()
    # (1) checksum of checksums: trailer CRC == fold of per-block CRCs recomputed on the CPU from the input ranges
    want = 0
    for b in range(len(crc)):
        c = orc.crc32_bzip2(data[int(in_off[b]):int(in_off[b + 1])])
        assert c == int(crc[b])
        want = (((want << 1) | (want >> 31)) & 0xFFFFFFFF) ^ c
    tail = int.from_bytes(stream[-6:], "big")  # the 32-bit CRC ends within the last 5 bytes; search the bit offset
    found = any(((int.from_bytes(stream[-12:], "big") >> s) & 0xFFFFFFFF) == want for s in range(8))
    assert found, "combined CRC not found at the end of the stream"
    # (2) block sizes: every block but the last holds between T and T+4 bytes (encoder.rs:692)
    sizes = np.diff(rle_off.astype(np.int64))
    assert ((sizes[:-1] >= 899981) & (sizes[:-1] <= 899985)).all() and 0 < sizes[-1] <= 899985
    # (3) EVERY block section and the trailer bit-exact vs the oracle (block-parallel on the host cores), and a libbz2
    # round trip of the first 32 MiB
    from oracle import verify
    ok, msg, st = verify.verify_stream(data, 9, stream, in_off)
    assert ok and st["blocks_checked"] == len(crc), msg
    dec = bz2.BZ2Decompressor()
    got = bytearray()
    pos = 0
    while len(got) < (32 << 20) and pos < len(stream):
        got += dec.decompress(stream[pos:pos + (1 << 20)])
        pos += 1 << 20
    m = min(len(got), 32 << 20)
    assert bytes(got[:m]) == data[:m]
    ctx.close()


# ---- K7: the bit-granular join of shard bit strings (BitWriter<Left> across shard boundaries, bitio/writer.rs:186-224) ----

def test_bit_append_against_numpy():
    """bzb200_bit_append alone: random source bits OR-ed at random destination bit offsets (all 32 word phases, lengths
    that end inside a word, zero length) must equal the same concatenation done with numpy on the host."""
    import torch
    from rust_compression_b200 import device as dv
    rng = np.random.default_rng(3)
    ctx = dv.Context()
    for trial in range(40):
        nbits = int(rng.integers(0, 200_000)) if trial else 0
        dst_bit = int(rng.integers(0, 4096)) if trial % 2 else trial
        src = rng.integers(0, 256, size=(nbits + 7) // 8 + 8, dtype=np.uint8)
        head = rng.integers(0, 256, size=(dst_bit + 7) // 8, dtype=np.uint8)
        hb = np.unpackbits(head)[:dst_bit]
        want = np.concatenate([hb, np.unpackbits(src)[:nbits]])
        cap = ((dst_bit + nbits + 31) // 32) * 4 + 8
        dst = np.zeros(cap, dtype=np.uint8)
        dst[:(dst_bit + 7) // 8] = np.packbits(hb)  # bits below dst_bit set, everything above zero
        d_dst = torch.from_numpy(dst).cuda()
        d_src = torch.from_numpy(np.concatenate([src, np.zeros((-src.size) % 4, dtype=np.uint8)])).cuda()
        ctx.bit_append(d_dst, dst_bit, d_src, nbits)
        ctx.sync()
        got = np.unpackbits(d_dst.cpu().numpy())
        assert (got[:want.size] == want).all(), (trial, dst_bit, nbits)
        assert not got[want.size:].any(), (trial, "bits written past the end")
    ctx.close()


@pytest.mark.parametrize("level,make", [(1, lambda: gen.mixed(7, 1_300_000) + gen.g2(3, 500_000)),
                                        (9, lambda: gen.text(3, 4_000_000))])
def test_shard_streams_joined_by_bit_append(level, make):
    """What a sharded run does at a rank boundary: blocks [0,k) and [k,nb) are encoded into separate buffers (the
    second one from bit 0, as a rank other than 0 does), joined with bzb200_bit_append at the — unaligned — bit where
    the first part ends, the trailer is appended, and the result must be the oracle's stream bit for bit; every
    split point k is tried, and a three-way split."""
    import torch
    from rust_compression_b200 import device as dv
    data = make()
    want = orc.compress(data, level)
    d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    ctx = dv.Context()
    nb = ctx.plan(level, d_in)
    assert nb >= 4
    cap = (dv.max_output_bytes(level, len(data)) + 64 + 3) & ~3
    splits = [(k,) for k in range(1, nb)] + [(1, nb - 1), (nb // 3, 2 * nb // 3)]
    for cuts in splits:
        edges = [0] + list(cuts) + [nb]
        d_out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
        ctx.write_stream_header(level, d_out)
        cur = ctx.encode_blocks(edges[0], edges[1], d_out, 32)
        crcs = [ctx.block_table(with_crc=False)[2][edges[0]:edges[1]].copy()]
        for lo, hi in zip(edges[1:-1], edges[2:]):
            part = torch.zeros(cap, dtype=torch.uint8, device="cuda")
            nbits = ctx.encode_blocks(lo, hi, part, 0)
            crcs.append(ctx.block_table(with_crc=False)[2][lo:hi].copy())
            ctx.bit_append(d_out, cur, part, nbits)
            cur += nbits
        total = ctx.write_stream_trailer(d_out, cur, ctx.combine_crc(np.concatenate(crcs)))
        ctx.sync()
        got = d_out[:total].cpu().numpy().tobytes()
        assert got == want, f"split {cuts}: joined stream differs from the oracle"
    ctx.close()
