"""Pins the CPU oracle (oracle/bz2_oracle.cpp) against every golden vector the
reference's own tests hold for the bzip2 encode path, against SURVEY.md App. B
known answers (an independent transliteration), and against libbz2 as decoder.
CPU-only; runs in seconds.
"""
import bz2
import hashlib
import os
import random

import numpy as np
import pytest

import gen
from oracle import orc


# ---- reference goldens ----------------------------------------------------

def test_unit_golden(golden):
    """src/bzip2/mod.rs:41-58 — the only encoder-bit golden in the reference."""
    g = golden["test_unit"]
    out = orc.compress(bytes.fromhex(g["input"]), g["level"])
    assert out.hex() == g["output"]
    assert bz2.decompress(out) == b"a\n"


def test_bwt_strings(golden):
    """src/suffix_array/sais.rs:294-345 test_bwt1-8 (BWT last column)."""
    assert len(golden["bwt_str"]) == 15
    for v in golden["bwt_str"]:
        src = bytes.fromhex(v["src"])
        for mode in (0, 1, 2):
            sa, _ = orc.bwt(src, mode)
            last = bytes(src[(int(s) - 1) % len(src)] for s in sa)
            assert last == bytes.fromhex(v["bwt"])


def test_bwt_positions(golden):
    """src/suffix_array/sais.rs:348-556 test_bwtpos (exact rotation-index arrays)."""
    assert len(golden["bwt_pos"]) == 42
    for v in golden["bwt_pos"]:
        src = bytes.fromhex(v["src"])
        for mode in (0, 1, 2):
            sa, _ = orc.bwt(src, mode)
            assert sa.tolist() == v["pos"]


def test_huffman_goldens(golden):
    h = golden["huffman"]
    # cano_huff_table.rs:252-264
    w = [f << h["with_fn"]["shift"] for f in h["with_fn"]["freq"]]
    lens, _ = orc.huffman(w, h["with_fn"]["lim"], 1)
    assert lens.tolist() == h["with_fn"]["lens"]
    # :238-250
    lens, _ = orc.huffman(h["cost80"]["freq"], h["cost80"]["lim"], 0)
    assert sum(int(l) * f for l, f in zip(lens, h["cost80"]["freq"])) == h["cost80"]["total_cost"]
    # :266-286 (length limit 8 on 63 symbols: bound + max only)
    freq = h["lim_len"]["freq"]
    lens, lm = orc.huffman([f << 8 for f in freq], h["lim_len"]["lim"], 1)
    assert lm, "package-merge path expected"
    assert max(lens) <= 8
    assert sum(int(l) * f for l, f in zip(lens, freq)) < sum(freq) * 6
    # :288-294
    lens, _ = orc.huffman(h["unit"]["freq"], h["unit"]["lim"], 0)
    assert lens.tolist() == h["unit"]["lens"]


def test_canonical_codes_golden(golden):
    """src/huffman/encoder.rs:64-79."""
    c = golden["canonical"]
    codes = orc.canonical_codes(c["lens"])
    for got, want, l in zip(codes, c["codes"], c["lens"]):
        if l:
            assert int(got) == want


def test_bitwriter_goldens(golden):
    """src/bitio/writer.rs:253-322 (Left cases)."""
    for v in golden["bitwriter"]:
        assert list(orc.pack_bits([tuple(f) for f in v["fields"]])) == v["bytes"]


def test_crc():
    # CRC-32/BZIP2 check value of "123456789"
    assert orc.crc32_bzip2(b"123456789") == 0xFC891918
    # the CRC field of test_unit (src/bzip2/mod.rs:53-54)
    assert orc.crc32_bzip2(b"a\n") == 0x633ED6E2


# ---- SURVEY.md App. B (second opinion) ---------------------------------------

APPB_HEX = [
    (b"", 9, "425a683917724538509000000000"),
    (b"a", 9, "425a683931415926535919939b6b00000001002000200021184682ee48a70a120332736d60"),
    (b"aabbaabbaabbaabb\n", 9,
     "425a68393141592653597e6ce699000002410000103000200030934c154da91a231e2ee48a70a120fcd9cd32"),
    (b"a" * 4, 9, "425a6839314159265359881233a600000241004000200020002100820b177245385090881233a6"),
    (b"a" * 5, 9, "425a683931415926535944a4303d00000241002000200020002100820b17724538509044a4303d"),
    (b"aaaa\x00", 9, "425a6839314159265359ecc0eb1d000002c100400020002000308049ea0ce2ee48a70a121d981d63a0"),
    (b"a" * 255 + b"b", 9, "425a6839314159265359e2ef08ee0000000100b0000008200030934c33d41738bb9229c2848717784770"),
    (b"a" * 256, 9, "425a6839314159265359efac2e370000008100a0000008200021008293177245385090efac2e37"),
    (b"a" * 1000, 9,
     "425a683931415926535949dc4f630000018101a00000800008200020aa6d41269aea0f17724538509049dc4f63"),
    (b"ab" * 500, 9, "425a6839314159265359fc30145d0000f981003000200030804d46a41a907177245385090fc30145d0"),
    (b"aabb" * 300, 9,
     "425a6839314159265359422bc47c00009581003000200030802918a462918a46291c5dc914e1424108af11f0"),
    (b"abcd" * 64 + b"e", 9,
     "425a6839314159265359bbd7b10600000001003e00200030c1a0548696c924d0380a0778bb9229c28485debd8830"),
    (b"a" * 100000, 1,
     "425a68313141592653594351d9f50000c61100840020000008200030934c14a69c610b6042f1772453850904351d9f50"),
]


@pytest.mark.parametrize("data,level,want", APPB_HEX, ids=[f"appb{i}" for i in range(len(APPB_HEX))])
def test_appendix_b_hex(data, level, want):
    out = orc.compress(data, level)
    assert out.hex() == want
    assert bz2.decompress(out) == data


APPB_SAMPLES = [
    (1, 9, 32352, [(98170, 97613, 59500, 258, 6, 1190)], "435c67f98520df57c33d1d057fdaa4b6f305987075df1e148e5ceb6f04293305"),
    (2, 9, 72618, [(211468, 210026, 134484, 258, 6, 2690)], "01549ee6bd261c1ce6e5b9f394c861ab0f1d9604ba2925a7381acc0a11e7b013"),
    (3, 9, 234, [(120244, 30061, 292, 34, 3, 6)], "e4946445c7f425d84332bdc4a0d06ddfb4a7a60e9fbbe7547587f8dc168ea239"),
    (4, 9, 40488, [(196340, 195226, 100444, 258, 6, 2009)], "1ffbb3bd07e573f8d7054724ebb65c57d0772cef9fa4c7bdafd21cd037cd4dc1"),
    (5, 9, 325566, [(323742, 173552, 323738, 258, 6, 6475)], "f3394877007534b1e4428d55d29fb2df1a7d1ea895238931504a27c1c487f963"),
    (6, 9, 325645, [(323743, 173551, 323739, 258, 6, 6475)], "8f9701e057b95c780c568ab580efdaae359a63280b240e5916bf06bd3a3f022e"),
    (7, 9, 66179, [(65539, 4936, 65538, 258, 6, 1311)], "07c97bf8d76e79a8fd1a87d0369d942aa9b6ead3d395b76e0dfb5ec3b439218f"),
    (1, 1, 32352, None, "75404d78acd14546952feb6961043f2e6bcb1c5ed939a2577693a916d44cc60b"),
    (2, 2, 73730, [(199981, 198650, 126791, 258, 6, 2536), (11487, 8635, 8245, 247, 6, 165)],
     "9eb5acceee8506dd0123e8e6a3e897381a5c0624851e59e5c189a5fcab7ce6c2"),
    (3, 3, 234, None, "74d531bb44a4d42d2d38403f6cd99eaa2b854c8fdbf04590a88cf40c41f78528"),
]


@pytest.mark.parametrize("idx,level,size,blocks,sha", APPB_SAMPLES, ids=[f"sample{a[0]}@{a[1]}" for a in APPB_SAMPLES])
def test_appendix_b_samples(sample_data, idx, level, size, blocks, sha):
    data = sample_data[idx]
    r = orc.Run(data, level)
    assert len(r.out) == size
    assert hashlib.sha256(r.out).hexdigest() == sha
    assert bz2.decompress(r.out) == data
    if blocks:
        assert r.nblocks == len(blocks)
        for b, want in enumerate(blocks):
            i = r.info(b)
            assert (i["nblock"], i["orig_ptr"], i["mtf_count"], i["alpha"], i["ngroups"], i["nselectors"]) == want


def test_appendix_b_generators():
    out = orc.Run(gen.g1(1, 250000), 1)
    assert len(out.out) == 147477
    assert hashlib.sha256(out.out).hexdigest() == "2d66d9a4d1e1f6a6aa6cb893e7643971c33e18c91104fdec0324ed1f7178c7ec"
    assert [out.info(b)["nblock"] for b in range(out.nblocks)] == [99981, 99981, 50072]
    out = orc.Run(gen.g1(3, 200000), 9)
    assert hashlib.sha256(out.out).hexdigest() == "5152906437cac48a5d15a516a1b9005ef7e8ed6c4576de86d42b4464283953f9"
    d = gen.g2(2, 4000000)
    assert hashlib.sha256(d).hexdigest() == "626811ac515f30961553b4fea88f852a4f0f3d555b72bce40e4e4d663156d80f"
    out = orc.Run(d, 1)
    assert hashlib.sha256(out.out).hexdigest() == "b835d779dda11bee33b26d4401426b6745dae6b64b0c10d3773198114ac49e71"
    assert [out.info(b)["nblock"] for b in range(3)] == [99983, 99985, 50812]


def test_appendix_b_huffman_limit():
    fib = [1, 1]
    while len(fib) < 30:
        fib.append(fib[-1] + fib[-2])
    lens, lm = orc.huffman(fib, 17, 2)
    assert lm
    assert lens.tolist() == [17, 17, 16, 15, 14, 13, 13, 13, 12, 12, 11, 11, 10, 10, 9, 9, 8, 8, 7, 7, 6, 6, 5, 5, 4, 4,
                             3, 3, 2, 2]
    lens, lm = orc.huffman([0] * 5 + fib[:25] + [0, 7, 900000], 17, 2)
    assert lens.tolist() == [15, 15, 15, 15, 15, 15, 15, 14, 14, 13, 13, 12, 11, 11, 10, 10, 9, 9, 8, 8, 7, 7, 6, 6, 5,
                             5, 4, 4, 3, 3, 15, 13, 1]


# ---- structural properties -----------------------------------------------------

def _model_order(s):
    """suffix order of T' = rot(s, shift) + sentinel, mapped back; shift = smallest
    index of a minimal rotation."""
    n = len(s)
    rots = [s[i:] + s[:i] for i in range(n)]
    m = min(rots)
    shift = rots.index(m)
    t = s[shift:] + s[:shift]
    order = sorted(range(n), key=lambda p: t[p:])  # python compares prefixes as smaller == implicit sentinel
    return [(p + shift) % n for p in order], shift


def test_bwt_matches_rotation_model():
    rnd = random.Random(1234)
    for trial in range(1500):
        kind = trial % 5
        if kind == 0:
            s = bytes(rnd.choice(b"ab") for _ in range(rnd.randint(1, 40)))
        elif kind == 1:
            u = bytes(rnd.choice(b"abc") for _ in range(rnd.randint(1, 6)))
            s = u * rnd.randint(1, 8)
        elif kind == 2:
            u = bytes(rnd.choice(b"abc") for _ in range(rnd.randint(1, 6)))
            s = u * rnd.randint(2, 8)
            s = s[:-1] + bytes([rnd.choice(b"abcd")])
        elif kind == 3:
            s = bytes(rnd.randrange(256) for _ in range(rnd.randint(1, 64)))
        else:
            s = bytes(rnd.choice(b"aab") for _ in range(rnd.randint(1, 200)))
        want, shift = _model_order(s)
        for mode in (0, 1):
            sa, sh = orc.bwt(s, mode)
            assert sh == shift, (s, mode)
            assert sa.tolist() == want, (s, mode)


def test_least_rotation_fast_equals_literal():
    rnd = random.Random(7)
    for trial in range(3000):
        n = rnd.randint(1, 300)
        if trial % 3 == 0:
            u = bytes(rnd.choice(b"ab") for _ in range(rnd.randint(1, 7)))
            s = (u * (n // len(u) + 1))[:n]
        elif trial % 3 == 1:
            s = bytes(rnd.choice(b"abc") for _ in range(n))
        else:
            s = bytes([rnd.choice(b"ab")]) * n
        assert orc.least_rotation(s, True) == orc.least_rotation(s, False), s


def test_mtf_distinct_count_formulation():
    """MTF position = #distinct symbols since the previous occurrence, with the
    virtual prefix k-1..0 — the formulation the CUDA kernel uses — equals mtf.rs:22-38."""
    rnd = random.Random(5)
    for _ in range(200):
        k = rnd.randint(1, 40)
        seq = [rnd.randrange(k) for _ in range(rnd.randint(1, 300))]
        want = orc.mtf_positions(bytes(seq), k).tolist()
        last = {s: -1 - s for s in range(k)}
        got = []
        for i, c in enumerate(seq):
            got.append(sum(1 for s in range(k) if s != c and last[s] > last[c]))
            last[c] = i
        assert got == want


@pytest.mark.parametrize("level", [1, 9])
def test_roundtrip_libbz2(level):
    rnd = random.Random(level)
    cases = [b"", b"x", b"ab" * 70000, b"aabb" * 30000 + b"c", bytes(rnd.randrange(256) for _ in range(30000)),
             gen.g2(5, 300000), gen.text(3, 250000), gen.mixed(2, 300000), b"\x00" * 70000 + b"\xff" * 70000]
    for d in cases:
        out = orc.compress(d, level)
        assert bz2.decompress(out) == d


def test_stage_agreement_with_libbz2_on_text():
    """RLE1/cuts/CRC/BWT/MTF stages equal libbz2's on non-periodic input; check the
    fields that are visible in a libbz2 stream header: block CRC, origPtr, and stream CRC."""
    d = gen.text(9, 200000)
    r = orc.Run(d, 1)
    ref = bz2.compress(d, 1)

    def header(b):
        # first block: magic(32) blockmagic(48) crc(32) rand(1) origptr(24)
        bits = int.from_bytes(b[:20], "big")
        total = 160
        crc = (bits >> (total - 32 - 48 - 32)) & 0xFFFFFFFF
        op = (bits >> (total - 32 - 48 - 32 - 1 - 24)) & 0xFFFFFF
        return crc, op

    assert header(r.out) == header(ref)
    assert r.out[-4:] != b"" and r.info(0)["crc"] == header(ref)[0]


# ---- the restated reference DEcoder (oracle/bz2_decoder_oracle.cpp), pinned on the reference's decoder fixtures ----

@pytest.mark.parametrize("idx", [1, 2, 3, 4])
def test_decoder_reference_fixtures(sample_data, idx):
    """bzip2/mod.rs:84-148 test_sample1..4 (decode half): data/sampleN.bz2 -> sampleN.ref; sample4.bz2 is two
    concatenated streams (multi-stream restart, decoder.rs:510-520)."""
    z = open(os.path.join(os.path.dirname(__file__), "golden", "data", f"sample{idx}.bz2"), "rb").read()
    assert orc.decode(z) == sample_data[idx]
    assert bz2.decompress(z) == sample_data[idx]


def test_decoder_accepts_encoder_output_at_the_block_limits(sample_data):
    """The encoder's largest block (level*100000-15 bytes) is inside the decoder's limits (decoder.rs:238,399,427):
    round trip of a full block and of the reference's own round-trip inputs (mod.rs:84-172, lib.rs:13-33)."""
    for data, level in ((gen.text(31, 100200), 1), (b"a" * 1000, 9), (b"aabbaabbaabbaabb\n", 9), (b"", 9),
                        (sample_data[2], 2), (gen.g2(4, 400000), 1)):
        assert orc.decode(orc.compress(data, level)) == data


def test_decoder_error_kinds():
    """BZip2Error mapping (bzip2/error.rs:13-19; decoder.rs:170-188,197,470-475)."""
    z = orc.compress(gen.text(32, 60000), 9)
    # 'B','Z','h' are read with check_u8 and the comparison is discarded (decoder.rs:155-161,177-182): any three bytes
    # pass, only the level byte is checked — "XZh9" decodes, "BZx:" after a stream is DataErrorMagic because of ':'
    assert orc.decode(b"XZh9" + z[4:]) == gen.text(32, 60000)
    for bad, kind in ((b"XZh0" + z[4:], "DataErrorMagicFirst"), (z[:3] + b"0" + z[4:], "DataErrorMagicFirst"),
                      (z[:2], "DataErrorMagicFirst"), (z[:3], "UnexpectedEof"),
                      (z[:-3], "UnexpectedEof"), (z + b"BZx:", "DataErrorMagic"), (z + b"BZ", "DataErrorMagic"),
                      (z[:len(z) // 2] + bytes([z[len(z) // 2] ^ 0x10]) + z[len(z) // 2 + 1:], "DataError"),
                      (z[:-3] + bytes([z[-3] ^ 0x80]) + z[-2:], "DataError")):   # a bit of the combined CRC
        with pytest.raises(orc.DecodeError) as e:
            orc.decode(bad)
        assert e.value.kind == kind, (kind, e.value.kind)
    assert orc.decode(z + z) == gen.text(32, 60000) * 2   # multi-stream


def test_invalid_level():
    with pytest.raises(ValueError):
        orc.compress(b"abc", 0)
    with pytest.raises(ValueError):
        orc.compress(b"abc", 10)
