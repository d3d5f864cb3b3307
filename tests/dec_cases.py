"""Decoder test cases shared by the CPU emulation test (test_dec_emu.py) and the GPU parity test
(test_gpu_decoder.py): (name, .bz2 buffer) pairs, valid and malformed, all deterministic.  The expected result of a
case is whatever the restated reference decoder (oracle.orc.decode) does with it: bytes, or bytes + BZip2Error kind."""
import bz2
import os

import numpy as np

import gen
from oracle import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "tests", "golden", "data")


def expected(buf):
    """(error code 0..5, bytes yielded before the error) from the restated reference decoder."""
    try:
        return 0, orc.decode(buf)
    except orc.DecodeError as ex:
        code = {v: k for k, v in orc.DecodeError.KINDS.items()}[ex.kind]
        return code, ex.partial


def valid_cases(big=False):
    cases = []
    for i in range(1, 5):  # the reference's decoder fixtures (bzip2/mod.rs:84-148); sample4 = two streams
        with open(os.path.join(DATA, f"sample{i}.bz2"), "rb") as f:
            cases.append((f"sample{i}.bz2", f.read()))
    txt = gen.text(3, 250000)
    for lvl in (1, 2, 9):
        cases.append((f"oracle-encoded text level {lvl}", orc.compress(txt, lvl)))
        cases.append((f"libbz2-encoded text level {lvl}", bz2.compress(txt, lvl)))
    cases.append(("empty stream", orc.compress(b"", 9)))
    cases.append(("test_unit a\\n", orc.compress(b"a\n", 9)))
    cases.append(("test_long a*1000", orc.compress(b"a" * 1000, 9)))
    cases.append(("all a, 3 blocks", orc.compress(b"a" * 300000, 1)))
    cases.append(("period 2", orc.compress(b"ab" * 60000, 1)))
    cases.append(("period 4 + tail", orc.compress(b"aabb" * 30000 + b"x", 1)))
    cases.append(("runs g2", orc.compress(gen.g2(5, 200000), 1)))
    cases.append(("random bytes", orc.compress(np.random.default_rng(1).integers(0, 256, 150000, dtype=np.uint8)
                                               .tobytes(), 1)))
    cases.append(("run lengths 4,5,255,259 and every byte value",
                  orc.compress(bytes(range(256)) * 40 + b"\xff" * 1000 + b"\x00" * 259 + b"\x00" * 4 + b"\x01" * 5, 9)))
    cases.append(("three streams, levels 1/3/5, one empty",
                  orc.compress(txt[:50000], 1) + orc.compress(b"", 3) + bz2.compress(txt[50000:120000], 5)))
    cases.append(("mixed level 1, 14 blocks", orc.compress(gen.mixed(1, 1_500_000), 1)))
    cases += randomised_cases()
    cases += damaged_magic_cases(valid=True)
    # symbol counts of 16 384, 16 385 and 16 386: EOB as the last symbol of a full 1 024-symbol chunk of the split D2,
    # alone in the last chunk, and second in it
    for extra in (615, 616, 618):
        cases.append((f"symbol count at a chunk boundary (+{extra})",
                      orc.compress(gen.text(12, 20000) + gen.text(99, extra), 9)))
    if big:
        t2 = gen.text(7, 2_000_000)
        cases.append(("text 2 MB level 9", orc.compress(t2, 9)))
        cases.append(("all-a exactly one full block", orc.compress(b"a" * 45_899_235, 9)))
        cases.append(("period 2, full block", orc.compress(b"ab" * 500_000, 9)))
        cases.append(("period 4 + tail, full block", orc.compress(b"aabb" * 230_000 + b"q", 9)))
    return cases


def randomised_cases():
    """Streams with the `randomised` bit set, as bzip2 <= 0.9.0 wrote them (decoder.rs:94-116,230-234,478-480,537-539).
    No such encoder exists any more, so the fixtures come from oracle.orc.compress_randomised (the format's rule: block
    bytes XOR the BZ2_rNums mask before the BWT) and are pinned on libbz2, which still decodes them."""
    txt = gen.text(3, 250000)
    out = []
    for name, data, lvl in (("text level 1, 3 blocks", txt, 1), ("text level 9", gen.text(5, 1_200_000), 9),
                            ("all a", b"a" * 300000, 1), ("period 2", b"ab" * 70000, 2),
                            ("mixed level 1", gen.mixed(2, 350000), 1)):
        s = orc.compress_randomised(data, lvl)
        assert bz2.decompress(s) == data, name
        out.append((f"randomised blocks: {name}", s))
    # a randomised stream between two ordinary ones (multi-stream), and a randomised block next to an ordinary one
    out.append(("randomised stream between ordinary streams",
                orc.compress(txt[:40000], 2) + orc.compress_randomised(txt[40000:90000], 1) + bz2.compress(txt[90000:], 3)))
    return out


def _xor_bits(stream, bitpos, nbytes, mask=0xFF):
    """XORs `nbytes` bytes starting at bit offset bitpos (unaligned) with mask."""
    b = bytearray(stream)
    for k in range(nbytes):
        for i in range(8):
            if (mask >> (7 - i)) & 1:
                p = bitpos + 8 * k + i
                b[p >> 3] ^= 0x80 >> (p & 7)
    return bytes(b)


def damaged_magic_cases(valid):
    """The reference reads every magic byte except the first of each 48-bit magic with check_u8 and discards the
    comparison (decoder.rs:155-161,177-182,211-224,495-508): 'B','Z','h' and bytes 2..6 of both magics may hold
    anything (valid=True: these decode normally); the level byte and the first magic byte are checked
    (valid=False: DataErrorMagicFirst / DataErrorMagic / DataError)."""
    txt = gen.text(3, 250000)
    s = orc.compress(txt, 1)                       # three blocks
    r = orc.Run(txt, 1)
    starts = [r.info(b)["bit_start"] for b in range(r.nblocks)]
    end = r.info(r.nblocks - 1)["bit_end"]
    r.close()
    two = orc.compress(txt[:30000], 1)
    if valid:
        out = [("stream magic 'XZh'", b"X" + s[1:]), ("stream magic all zero", b"\0\0\0" + s[3:]),
               ("block magic bytes 2..6 damaged, first block", _xor_bits(s, starts[0] + 8, 5)),
               ("block magic byte 4 damaged, second block (unaligned)", _xor_bits(s, starts[1] + 24, 1, 0x10)),
               ("block magic bytes 2 and 6 damaged, last block", _xor_bits(_xor_bits(s, starts[2] + 8, 1), starts[2] + 40, 1)),
               ("every block magic damaged", _xor_bits(_xor_bits(_xor_bits(s, starts[0] + 16, 2), starts[1] + 8, 5),
                                                       starts[2] + 32, 1, 0x01)),
               ("end magic bytes 2..6 damaged", _xor_bits(s, end + 8, 5)),
               ("end magic damaged, a level-9 stream follows", _xor_bits(two, len(two) * 8 - 80 + 16, 1)[:len(two)] +
                orc.compress(gen.text(8, 300000), 9)),
               ("second stream magic 'QQQ'", two + b"QQQ" + orc.compress(txt[:20000], 3)[3:])]
        return out
    return [("first block magic, first byte damaged", _xor_bits(s, starts[0], 1, 0x01)),
            ("second block magic, first byte damaged", _xor_bits(s, starts[1], 1, 0x80)),
            ("end magic, first byte damaged", _xor_bits(s, end, 1, 0x02)),
            ("block head byte 0x17 instead of 0x31", _xor_bits(s, starts[1], 1, 0x31 ^ 0x17)),
            ("end head byte 0x31 instead of 0x17", _xor_bits(s, end, 1, 0x31 ^ 0x17)),
            ("level byte '0' with magic 'XZh'", b"XZh0" + s[4:]),
            ("second stream level byte ':'", two + b"BZh:" + orc.compress(txt[:20000], 3)[4:]),
            ("damaged block magic cut off after 3 bytes", _xor_bits(s, starts[2] + 8, 2)[:(starts[2] + 24 + 7) // 8])]


def malformed_cases():
    """Truncations, bit flips and trailing bytes of one two-block stream.  Positions were chosen to hit every header
    field, the selector/table area, the symbol area of both blocks, the end magic and the combined CRC; each was run
    through the restated reference decoder when the list was written (none sends it into its endless-read case, a block
    that ends in four equal bytes without a count)."""
    s = orc.compress(gen.text(3, 120000), 1)
    n = len(s)
    cases = [("empty input", b"")]
    for cut in (1, 2, 3, 4, 5, 9, 10, 13, 14, 17, 18, 20, 40, 100, n // 2, n - 11, n - 10, n - 5, n - 4, n - 1):
        cases.append((f"truncated to {cut}", s[:cut]))
    for pos in (0, 1, 2, 3, 4, 5, 10, 11, 14, 15, 17, 18, 19, 20, 25, 30, 60, 200, n // 2, n - 12, n - 8, n - 3, n - 1):
        b = bytearray(s)
        b[pos] ^= 0x10
        cases.append((f"bit flip in byte {pos}", bytes(b)))
    for tail in (b"xyz", b"B", b"BZh", b"BZh9", b"BZh0"):
        cases.append((f"trailing {tail!r}", s + tail))
    # header fields patched in otherwise valid streams: the stream's level against the block sizes it holds
    # (decoder.rs:399,427), origPtr against its bounds (:238, :446) and — inside the bounds — a walk that starts at
    # the wrong rotation (CRC error after the block's bytes; a periodic block decodes to a rotation of a periodic text)
    t = gen.text(3, 250000)
    s9 = orc.compress(t, 9)
    for lv in b"12348":
        cases.append((f"level byte patched to {chr(lv)}", s9[:3] + bytes([lv]) + s9[4:]))
    s1 = orc.compress(t[:60000], 1)
    for v in (0, 1, 59999, 60000, 100010, 100011, 0xFFFFFF):
        cases.append((f"origPtr patched to {v}", _set_orig(s1, v)))
    per = orc.compress(b"abcabc" * 9000, 1)
    for v in (0, 3, 100, 53999):
        cases.append((f"periodic block, origPtr patched to {v}", _set_orig(per, v)))
    alla = orc.compress(b"a" * 200000, 1)
    for v in (0, 1, 4, 100):  # 0: the block then ends in four equal bytes without a count (see same_result)
        cases.append((f"all-a block, origPtr patched to {v}", _set_orig(alla, v)))
    cases += damaged_magic_cases(valid=False)
    # a randomised stream that is damaged: truncated inside the block, and with the randomised bit cleared (CRC error)
    rs = orc.compress_randomised(t[:90000], 1)
    cases.append(("randomised stream truncated", rs[:len(rs) // 2]))
    b = bytearray(rs)
    b[(32 + 48 + 32) >> 3] ^= 0x80 >> ((32 + 48 + 32) & 7)
    cases.append(("randomised bit cleared", bytes(b)))
    b = bytearray(orc.compress(t[:90000], 1))
    b[(32 + 48 + 32) >> 3] ^= 0x80 >> ((32 + 48 + 32) & 7)
    cases.append(("randomised bit set on an ordinary block", bytes(b)))
    return cases


def _set_orig(stream, v):
    """Overwrites the 24-bit origPtr of the first block (stream header 32 bits, magic 48, CRC 32, randomised 1)."""
    b = bytearray(stream)
    pos = 32 + 48 + 32 + 1
    for i in range(24):
        p = pos + i
        m = 0x80 >> (p & 7)
        b[p >> 3] = (b[p >> 3] & ~m) | (m if (v >> (23 - i)) & 1 else 0)
    return bytes(b)


def fuzz_cases(n, seed):
    """n random mutations (bit flips, truncation, overwrite, splice, delete, insert) of a few small valid streams."""
    import random
    rnd = random.Random(seed)
    bases = [orc.compress(gen.text(3, 30000), 1), bz2.compress(gen.mixed(2, 40000), 1), orc.compress(gen.g2(5, 20000), 1),
             orc.compress(b"ab" * 3000 + b"c" * 5000 + bytes(range(256)) * 4, 2), orc.compress(gen.text(9, 130000), 1),
             orc.compress(gen.text(4, 5000), 9) + orc.compress(gen.text(5, 7000), 3)]
    for it in range(n):
        b = bytearray(rnd.choice(bases))
        op = rnd.randrange(6)
        if op == 0:
            for _ in range(rnd.choice([1, 1, 1, 2, 3])):
                b[rnd.randrange(len(b))] ^= 1 << rnd.randrange(8)
        elif op == 1:
            b = b[:rnd.randrange(len(b))]
        elif op == 2:
            b[rnd.randrange(len(b))] = rnd.randrange(256)
        elif op == 3:
            p, q = rnd.randrange(len(b)), rnd.randrange(len(b))
            b[p:p + 8] = b[q:q + 8]
        elif op == 4:
            p = rnd.randrange(len(b))
            del b[p:p + rnd.randrange(1, 6)]
        else:
            p = rnd.randrange(len(b))
            b[p:p] = bytes(rnd.randrange(256) for _ in range(rnd.randrange(1, 5)))
        yield f"fuzz {seed}/{it} op {op}", bytes(b)


def same_result(got, want):
    """got / want = (error code, bytes).  Code 6 of the restated reference decoder means "the reference never stops on
    this input" (a block ending in four equal bytes with no count byte): this build reports DataError before that block,
    so its bytes are a prefix of what the reference had yielded by then."""
    if want[0] == 6:
        return got[0] == 1 and want[1].startswith(got[1])
    return got == want
