"""The full-stream verifier (oracle/verify.py) against the oracle's own sequential run: the per-block factorisation it
relies on must reproduce the one-pass stream bit for bit, and it must notice every kind of damage it is there to find
(bench.py and the GPU tests use it to check ALL blocks of the large streams)."""
import numpy as np
import pytest

import gen
from oracle import orc, verify


def _table(r, n):
    return [0] + [r.info(b)["in_end"] for b in range(r.nblocks)] if n else [0]


CASES = [
    ("g2", lambda: gen.g2(2, 1_500_000), 1),
    ("mixed", lambda: gen.mixed(1, 700_000), 1),
    ("text", lambda: gen.text(1, 2_000_000), 9),
    ("aaaab", lambda: b"aaaab" * 200_000, 1),
    ("a", lambda: b"a" * 3_000_000, 1),
    ("empty", lambda: b"", 5),
    ("one", lambda: b"x", 9),
]


@pytest.mark.parametrize("name,make,level", CASES, ids=[c[0] for c in CASES])
def test_factorisation_equals_one_pass_stream(name, make, level):
    data = make()
    r = orc.Run(data, level)
    offs = _table(r, len(data))
    ok, msg, st = verify.verify_stream(data, level, r.out, offs)
    assert ok, msg
    assert st["blocks"] == r.nblocks and st["bits"] + (-st["bits"]) % 8 == len(r.out) * 8
    if r.nblocks:
        bad = bytearray(r.out)
        bad[len(bad) // 2] ^= 0x10                      # one bit inside some block section
        assert not verify.verify_stream(data, level, bytes(bad), offs)[0]
        bad = bytearray(r.out)
        bad[-2] ^= 1                                    # combined CRC
        assert not verify.verify_stream(data, level, bytes(bad), offs)[0]
        assert not verify.verify_stream(data, level, r.out + b"\0", offs)[0]   # trailing byte
        assert not verify.verify_stream(data, level, r.out[:-1], offs)[0]      # truncated
    if r.nblocks >= 3:
        early = list(offs)
        early[1] -= 1                                   # cut one byte early: < T bytes or inside a piece
        assert not verify.verify_stream(data, level, r.out, early)[0]
        late = list(offs)
        late[1] += 1                                    # cut one byte late: the oracle cuts the range in two
        ok, msg, _ = verify.verify_stream(data, level, r.out, late)
        assert not ok
    r.close()


def test_prefix_subset_and_wrong_level():
    data = gen.mixed(3, 600_000)
    r = orc.Run(data, 1)
    offs = _table(r, len(data))
    ok, msg, st = verify.verify_stream(data, 1, r.out, offs, blocks=range(3))
    assert ok and st["blocks_checked"] == 3
    assert not verify.verify_stream(data, 2, r.out, offs)[0]
    assert not verify.verify_stream(data, 1, r.out, offs, blocks=[1, 2])[0]
    r.close()
