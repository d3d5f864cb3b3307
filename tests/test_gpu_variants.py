"""The kernel variants behind the library's environment switches produce the same bits as the defaults.

The switches are read once per process, so every setting runs in a child process: the radix group sort in every round /
never (BZB200_LS_RX), the persistent TMA-fed radix pass for every pass / never (BZB200_OS_PF), the byte-key initial sort
(BZB200_KEY_BITS=8).  Same check as the parity tests: the stream of the device path equals the oracle's."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import torch
import gen
import rust_compression_b200  # noqa: F401
from rust_compression_b200 import device as dv
from oracle import orc
cases = [("text level 9", gen.text(21, 2_700_000), 9),
         ("mixed level 1", gen.mixed(22, 1_500_000), 1),
         ("runs + text level 2", b"z" * 300_000 + gen.text(23, 500_000) + b"ab" * 200_000 + b"q", 2),
         ("near periodic", b"abcd" * 120_000 + b"x", 9),
         ("small alphabet", gen.g2(24, 900_000), 5)]
ctx = dv.Context()
for name, data, level in cases:
    d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    got = dv.compress_tensor(ctx, level, d_in).cpu().numpy().tobytes()
    assert got == orc.compress(data, level), name
print("ok")
"""

VARIANTS = [{"BZB200_LS_RX": "0"}, {"BZB200_LS_RX": "1"}, {"BZB200_OS_PF": "0"}, {"BZB200_OS_PF": "1"},
            {"BZB200_KEY_BITS": "8"}, {"BZB200_LS_RX": "1", "BZB200_OS_PF": "1", "BZB200_LS_TPC": "5"}]


@pytest.mark.gpu
@pytest.mark.parametrize("env", VARIANTS, ids=[" ".join(f"{k}={v}" for k, v in e.items()) for e in VARIANTS])
def test_variant_matches_oracle(env):
    e = dict(os.environ)
    e.update(env)
    p = subprocess.run([sys.executable, "-c", CHILD, ROOT], env=e, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and p.stdout.strip().endswith("ok"), p.stdout[-2000:] + p.stderr[-4000:]
