"""Full-stream verifier: every block of a .bz2 stream against the oracle, block-parallel over the host cores.

TEST INFRASTRUCTURE ONLY (see oracle/orc.py): used by tests/ and by bench.py as the checker of the streams the GPU
path produced — never on the product path, never inside a timed region.

The reference encoder (src/bzip2/encoder.rs) is a sequential state machine, but its output factors per block once
the block cuts are known: a cut is taken right after a flushed run piece (encoder.rs:692-696), the RLE1 run counter
starts afresh in the new block, and nothing but the combined CRC carries over (write_block, :224-291).  So

  stream == "BZh" level  ++  for each block b: section_b  ++  end magic  ++  combined CRC  ++  zero padding

where section_b is what the oracle emits for in[in_off[b]:in_off[b+1]] encoded as a stream of its own, PROVIDED the
cuts are the reference's cuts.  That is checked separately, from the input alone:
  * the range of every block but the last emits >= T bytes after RLE1, and the oracle does not cut it any earlier
    (it comes out as exactly one block);
  * every cut is a piece end: the byte after the cut differs from the byte before it, or the equal bytes before the
    cut, counted back to the block start / the previous different byte, are a multiple of 255.
"""
import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import orc

END_MAGIC = 0x177245385090


def _threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def _lib():
    L = orc.lib()
    if not getattr(L, "_verify_ready", False):
        L.orc_encode_block.restype = C.c_longlong
        L.orc_encode_block.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_cut_table.restype = C.c_longlong
        L.orc_cut_table.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_bits_diff.restype = C.c_uint64
        L.orc_bits_diff.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p, C.c_uint64]
        L._verify_ready = True
    return L


def _get_bits(a, bit, n):
    """n <= 64 bits of the uint8 array a starting at bit `bit`, MSB first."""
    v = 0
    for k in range(bit, bit + n):
        v = (v << 1) | ((int(a[k >> 3]) >> (7 - (k & 7))) & 1)
    return v


def cut_table(data, level):
    """in_off[nb + 1] of the reference's block cuts (sequential pass, oracle/bz2_oracle.cpp orc_cut_table)."""
    L = _lib()
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data)
    cap = a.size // (level * 100000 - 19) * 2 + 8
    out = np.zeros(cap, dtype=np.uint64)
    r = L.orc_cut_table(level, a.ctypes.data, a.size, out.ctypes.data, cap)
    if r < -1:
        out = np.zeros(-r, dtype=np.uint64)
        r = L.orc_cut_table(level, a.ctypes.data, a.size, out.ctypes.data, out.size)
    assert r >= 0
    return out[:r + 1].astype(np.int64)


def encode_blocks_parallel(data, level, in_off, threads=None):
    """Every block of the table encoded by the oracle, block-parallel on `threads` host threads (ctypes releases the
    GIL).  Returns (total compressed bits, list of block CRCs)."""
    L = _lib()
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data)
    cap = level * 100000 + level * 100000 // 4 + 8192

    def enc(b):
        lo, hi = int(in_off[b]), int(in_off[b + 1])
        out = np.empty(cap, dtype=np.uint8)
        info = np.zeros(4, dtype=np.uint64)
        L.orc_encode_block(level, a.ctypes.data + lo, hi - lo, out.ctypes.data, cap, info.ctypes.data)
        return int(info[3]), int(info[1])

    with ThreadPoolExecutor(max_workers=threads or _threads()) as ex:
        res = list(ex.map(enc, range(len(in_off) - 1)))
    return sum(r[0] for r in res), [r[1] for r in res]


def verify_stream(data, level, stream, in_off, threads=None, blocks=None):
    """data: the input (bytes / uint8 array); stream: the .bz2 bytes; in_off[nb+1]: the block table (input offsets).
    blocks: optional iterable of block indices to check bit for bit (default all; the structure checks always cover
    every block, and without all sections the positions of the checked ones come from the preceding sections, so a
    subset must be a prefix range).  Returns (ok, message, stats)."""
    L = _lib()
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    s = np.frombuffer(stream, dtype=np.uint8) if not isinstance(stream, np.ndarray) else stream
    a = np.ascontiguousarray(a)
    s = np.ascontiguousarray(s)
    in_off = np.asarray(in_off, dtype=np.int64)
    nb = in_off.size - 1
    T = level * 100000 - 19
    if s.size < 14 or bytes(s[:4]) != b"BZh" + bytes([0x30 + level]):
        return False, "stream header differs", {}
    if in_off[0] != 0 or in_off[-1] != a.size or (nb > 0 and (np.diff(in_off) <= 0).any()):
        return False, "block table does not tile the input", {}
    if a.size == 0 and nb != 0:
        return False, "blocks for an empty input", {}
    # cuts are piece ends
    for b in range(1, nb):
        c = int(in_off[b])
        if a[c] == a[c - 1]:
            lo = int(in_off[b - 1])
            seg = a[lo:c]
            d = np.nonzero(seg != seg[-1])[0]
            run = seg.size - (int(d[-1]) + 1 if d.size else 0)
            if run % 255 != 0:
                return False, f"cut before block {b} (input offset {c}) is inside a run piece (run of {run})", {}
    todo = list(range(nb)) if blocks is None else sorted(blocks)
    if todo != list(range(len(todo))):
        return False, "blocks must be a prefix range", {}
    cap = level * 100000 + level * 100000 // 4 + 8192
    res = [None] * len(todo)

    def enc(b):
        lo, hi = int(in_off[b]), int(in_off[b + 1])
        out = np.empty(cap, dtype=np.uint8)
        info = np.zeros(4, dtype=np.uint64)
        r = L.orc_encode_block(level, a.ctypes.data + lo, hi - lo, out.ctypes.data, cap, info.ctypes.data)
        if r < -1:
            out = np.empty(-r, dtype=np.uint8)
            r = L.orc_encode_block(level, a.ctypes.data + lo, hi - lo, out.ctypes.data, out.size, info.ctypes.data)
        return out[:max(r, 0)].copy(), [int(x) for x in info], int(r)

    with ThreadPoolExecutor(max_workers=threads or _threads()) as ex:
        for b, r in zip(todo, ex.map(enc, todo)):
            res[b] = r
    bit = 32
    combined = 0
    for b in todo:
        sect, (nblk, crc, nrle, nbits), r = res[b]
        if r < 0 or nblk != 1:
            return False, f"block {b}: the oracle cuts the range [{in_off[b]}, {in_off[b + 1]}) into {nblk} blocks", {}
        if b + 1 < nb and nrle < T:
            return False, f"block {b}: only {nrle} bytes after RLE1, the reference cuts at >= {T}", {}
        d = int(L.orc_bits_diff(s.ctypes.data, s.size, bit, sect.ctypes.data, nbits))
        if d != nbits:
            where = "stream too short" if d == (1 << 64) - 1 else f"first differing bit {d} of {nbits}"
            return False, f"block {b} (stream bit {bit}): {where}", {}
        bit += nbits
        combined = (((combined << 1) | (combined >> 31)) & 0xFFFFFFFF) ^ crc
    stats = {"blocks": nb, "blocks_checked": len(todo), "bits": bit}
    if len(todo) == nb:
        total = (bit + 80 + 7) // 8
        if s.size != total:
            return False, f"stream is {s.size} bytes, expected {total}", stats
        if _get_bits(s, bit, 48) != END_MAGIC:
            return False, "end-of-stream magic differs", stats
        if _get_bits(s, bit + 48, 32) != combined:
            return False, "combined CRC differs", stats
        pad = total * 8 - (bit + 80)
        if pad and _get_bits(s, bit + 80, pad) != 0:
            return False, "padding bits are not zero", stats
        stats["bits"] = bit + 80
    return True, f"all {len(todo)} blocks" + (" + trailer" if len(todo) == nb else f" of {nb}") + " bit-exact vs oracle", stats
