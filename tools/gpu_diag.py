#!/usr/bin/env python3
"""Runs a battery of inputs through the CUDA path and the oracle and prints, per case, which stages agree.
Does not stop at the first failure — one GPU call gives the whole picture.  python tests/gpu_diag.py [filter]"""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import gen  # noqa: E402
import parity  # noqa: E402


def cases():
    yield "a_nl", b"a\n", 9
    yield "a", b"a", 9
    yield "aa", b"aa", 9
    yield "readme", b"aabbaabbaabbaabb\n", 9
    yield "a4", b"a" * 4, 9
    yield "a5", b"a" * 5, 9
    yield "a256", b"a" * 256, 9
    yield "a1000", b"a" * 1000, 9
    yield "ab500", b"ab" * 500, 9
    yield "aabb300", b"aabb" * 300, 9
    yield "abcd64e", b"abcd" * 64 + b"e", 9
    yield "text5k", gen.text(5, 5000), 9
    yield "text100k", gen.text(7, 100000), 9
    yield "g1_250k_l1", gen.g1(1, 250000), 1
    yield "g2_1m_l1", gen.g2(2, 1000000), 1
    for i in (1, 3, 4, 7):
        with open(os.path.join(ROOT, "tests", "golden", "data", f"sample{i}.ref"), "rb") as f:
            yield f"sample{i}", f.read(), 9
    yield "a100000_l1", b"a" * 100000, 1
    yield "text900k", gen.text(1, 899900), 9
    yield "rand300k", bytes(gen.splitmix64(3, 300000 // 8).view("uint8")), 9


def main():
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    from oracle import orc
    nfail = 0
    for name, data, level in cases():
        if flt and flt not in name:
            continue
        t = time.time()
        try:
            res = parity.compare(data, level, orc, keep_sa=len(data) <= 2_000_000)
            bad = {k: v for k, v in res.items() if v}
            status = "OK  " if not bad else "FAIL"
            nfail += bool(bad)
            print(f"{status} {name:14s} level={level} n={len(data):8d} {time.time()-t:6.2f}s", flush=True)
            for k in parity.STAGES:
                if res[k]:
                    print(f"       {k:8s} {res[k]}", flush=True)
        except Exception:
            nfail += 1
            print(f"EXC  {name}: {traceback.format_exc()}", flush=True)
    print(f"failures: {nfail}")
    return 1 if nfail else 0


if __name__ == "__main__":
    sys.exit(main())
