"""bzb200_cut_walk — the host part of the sliced K1 plan (include/bzb200.h section 2b) — against the oracle's block
table.  The window rows the GPU kernels tabulate (k1_cut_windows: F[j][w] = f(centre_j + w) - centre_j, f(x) = emitted
offset of the first run-piece end at or after x) are rebuilt here with numpy from the definition of RLE1
(encoder.rs:671-716), so the walk, its phase handling (a drift that leaves the 256-offset window) and the end-of-input
rule (encoder.rs:729-739) run without a GPU."""
import ctypes as C

import numpy as np
import pytest

import gen
from oracle import orc


def rle_model(a):
    """Per input byte: emitted offset E(i) (inclusive) and piece-end flag, from the parallel formulation of RLE1."""
    n = a.size
    head = np.ones(n, dtype=bool)
    head[1:] = a[1:] != a[:-1]
    idx = np.arange(n)
    s = np.maximum.accumulate(np.where(head, idx, 0))
    q = (idx - s) % 255
    tail = np.ones(n, dtype=bool)
    tail[:-1] = a[1:] != a[:-1]
    pe = tail | (q == 254)
    emit = (q < 4).astype(np.int64) + (pe & (q >= 3)).astype(np.int64)
    return np.cumsum(emit), pe


def rows(E, pe, x0, T, K, W):
    """F rows 0..K-1 of the phase that starts at x0."""
    ends = np.nonzero(pe)[0]
    Ee = E[ends]
    n = E.size
    F = np.zeros((K, W), dtype=np.uint64)
    for j in range(K):
        c = x0 + (j + 1) * T
        xs = c + np.arange(W)
        k = np.searchsorted(Ee, xs, side="left")
        ok = k < ends.size
        kk = np.minimum(k, ends.size - 1)
        i = ends[kk]
        v = (Ee[kk] - c).astype(np.uint64) | ((i + 1).astype(np.uint64) << np.uint64(16))
        v |= np.where(i == n - 1, np.uint64(1) << np.uint64(63), np.uint64(0))
        F[j] = np.where(ok, v, np.uint64(0))
    return F


CASES = [("g2", lambda: gen.g2(2, 1_200_000), 1), ("aaaab: drift leaves the window", lambda: b"aaaab" * 2_400_000, 1),
         ("text", lambda: gen.text(3, 700_000), 1), ("last piece reaches T", lambda: gen.text(6, 99_978) + b"zzzz", 1),
         ("one block", lambda: gen.text(1, 5000), 9), ("long runs", lambda: b"x" * 2_000_000 + gen.text(5, 150_000), 1)]


@pytest.mark.parametrize("name,make,level", CASES, ids=[c[0] for c in CASES])
def test_cut_walk_reproduces_the_reference_cuts(name, make, level):
    from rust_compression_b200 import _lib
    L = _lib.lib()
    data = make()
    a = np.frombuffer(data, dtype=np.uint8)
    E, pe = rle_model(a)
    Etot, N, T, W = int(E[-1]), a.size, level * 100000 - 19, int(L.bzb200_cut_window())
    max_blocks = (N + N // 4 + 64) // T + 2
    state = np.zeros(4, dtype=np.uint64)
    in_off = np.zeros(max_blocks + 1, dtype=np.uint64)
    rle_off = np.zeros(max_blocks + 1, dtype=np.uint64)
    nb, ml = C.c_uint32(0), C.c_uint32(0)
    phases = 0
    while not state[2]:
        x0 = int(state[1])
        K = (Etot - x0) // T if Etot >= x0 + T else 0
        F = rows(E, pe, x0, T, K, W)
        rc = L.bzb200_cut_walk(F.ctypes.data if K else None, K, T, Etot, N, max_blocks, state.ctypes.data,
                               in_off.ctypes.data, rle_off.ctypes.data, C.byref(nb), C.byref(ml))
        assert rc == 0
        phases += 1
        assert phases < 1000
    r = orc.Run(data, level)
    assert nb.value == r.nblocks
    want_in = [0] + [r.info(b)["in_end"] for b in range(r.nblocks)]
    want_n = [r.info(b)["nblock"] for b in range(r.nblocks)]
    assert list(map(int, in_off[:nb.value + 1])) == want_in
    assert list(np.diff(rle_off[:nb.value + 1].astype(np.int64))) == want_n
    assert ml.value == max(want_n)
    if name.startswith("aaaab"):
        assert phases > 1
    r.close()


def test_cut_walk_rejects_corrupt_rows():
    from rust_compression_b200 import _lib
    L = _lib.lib()
    a = np.frombuffer(gen.text(3, 400_000), dtype=np.uint8)
    E, pe = rle_model(a)
    Etot, N, T, W = int(E[-1]), a.size, 99981, int(L.bzb200_cut_window())
    K = Etot // T
    F = rows(E, pe, 0, T, K, W)
    F[1, :] = 0  # a row nobody wrote
    state = np.zeros(4, dtype=np.uint64)
    in_off = np.zeros(64, dtype=np.uint64)
    rle_off = np.zeros(64, dtype=np.uint64)
    rc = L.bzb200_cut_walk(F.ctypes.data, K, T, Etot, N, 60, state.ctypes.data, in_off.ctypes.data, rle_off.ctypes.data,
                           None, None)
    assert rc == _lib.E_INTERNAL
