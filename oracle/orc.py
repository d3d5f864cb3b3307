"""ctypes binding of the CPU oracle (oracle/bz2_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(rust-compression_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    srcs = [os.path.join(_HERE, "bz2_oracle.cpp"), os.path.join(_HERE, "bz2_decoder_oracle.cpp")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liborc.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
        L.orc_compress.restype = C.c_longlong
        L.orc_compress.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_compress_randomised.restype = C.c_longlong
        L.orc_compress_randomised.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_run.restype = C.c_void_p
        L.orc_run.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_int]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_out_size.restype = C.c_size_t
        L.orc_out_size.argtypes = [C.c_void_p]
        L.orc_out_copy.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_nblocks.restype = C.c_size_t
        L.orc_nblocks.argtypes = [C.c_void_p]
        L.orc_block_info.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_block_inuse.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_block_field.restype = C.c_size_t
        L.orc_block_field.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t]
        L.orc_bwt.restype = C.c_size_t
        L.orc_bwt.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
        L.orc_least_rotation.restype = C.c_size_t
        L.orc_least_rotation.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
        L.orc_huffman.restype = C.c_size_t
        L.orc_huffman.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        L.orc_canonical_codes.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_pack_bits.restype = C.c_size_t
        L.orc_pack_bits.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_crc32_bzip2.restype = C.c_uint32
        L.orc_crc32_bzip2.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_mtf_positions.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
        L.orc_decode.restype = C.c_int
        L.orc_decode.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.orc_decode_free.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def _buf(data):
    a = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, dtype=np.uint8)
    return a


def compress(data, level=9):
    """BZip2Encoder::new(level) + encode(Action::Finish) — whole stream bytes."""
    if level < 1 or level > 9:
        raise ValueError("invalid level")  # bzip2/encoder.rs:59-61
    a = _buf(data)
    cap = int(a.size * 1.3) + 4096
    out = np.empty(cap, dtype=np.uint8)
    r = lib().orc_compress(level, a.ctypes.data, a.size, out.ctypes.data, cap)
    if r < -1:
        cap = -r
        out = np.empty(cap, dtype=np.uint8)
        r = lib().orc_compress(level, a.ctypes.data, a.size, out.ctypes.data, cap)
    assert r >= 0
    return out[:r].tobytes()


def compress_randomised(data, level=9):
    """FIXTURE GENERATOR, not reference behaviour: the stream bzip2 <= 0.9.0 wrote when it randomised a block (block
    bytes XOR-ed with the BZ2_rNums mask before the BWT, randomised bit set).  Valid .bz2 that the reference decoder —
    and libbz2 — un-randomise (decoder.rs:94-116,537-539); the reference ENcoder never produces it."""
    a = _buf(data)
    cap = int(a.size * 1.3) + 4096
    out = np.empty(cap, dtype=np.uint8)
    r = lib().orc_compress_randomised(level, a.ctypes.data, a.size, out.ctypes.data, cap)
    if r < -1:
        out = np.empty(-r, dtype=np.uint8)
        r = lib().orc_compress_randomised(level, a.ctypes.data, a.size, out.ctypes.data, out.size)
    assert r >= 0
    return out[:r].tobytes()


_FIELDS = {
    "rle": (0, np.uint8), "sa": (1, np.uint32), "last": (2, np.uint8), "mtf": (3, np.uint16), "freq": (4, np.uint32),
    "sel1": (5, np.uint8), "sel2": (6, np.uint8), "sel3": (7, np.uint8), "sel4": (8, np.uint8),
    "len0": (9, np.uint8), "len1": (10, np.uint8), "len2": (11, np.uint8), "len3": (12, np.uint8), "len4": (13, np.uint8),
}
_INFO = ["in_start", "in_end", "nblock", "crc", "orig_ptr", "mtf_count", "alpha", "ngroups", "nselectors",
         "bit_start", "bit_end", "shift", "lm_used"]


class Run:
    """Staged oracle run: whole output plus per-block stage dumps."""

    def __init__(self, data, level=9, keep_sa=False):
        self._a = _buf(data)
        self._h = lib().orc_run(level, self._a.ctypes.data, self._a.size, 1 if keep_sa else 0)
        if not self._h:
            raise ValueError("invalid level")
        n = lib().orc_out_size(self._h)
        o = np.empty(n, dtype=np.uint8)
        lib().orc_out_copy(self._h, o.ctypes.data)
        self.out = o.tobytes()
        self.nblocks = lib().orc_nblocks(self._h)

    def info(self, b):
        v = np.zeros(16, dtype=np.uint64)
        lib().orc_block_info(self._h, b, v.ctypes.data)
        return {k: int(v[i]) for i, k in enumerate(_INFO)}

    def inuse(self, b):
        v = np.zeros(8, dtype=np.uint32)
        lib().orc_block_inuse(self._h, b, v.ctypes.data)
        return v

    def field(self, b, name):
        fid, dt = _FIELDS[name]
        n = lib().orc_block_field(self._h, b, fid, None, 0)
        o = np.empty(n, dtype=dt)
        if n:
            lib().orc_block_field(self._h, b, fid, o.ctypes.data, n)
        return o

    def close(self):
        if self._h:
            lib().orc_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def bwt(data, mode=2):
    a = _buf(data)
    sa = np.zeros(max(a.size, 1), dtype=np.uint32)
    shift = lib().orc_bwt(a.ctypes.data, a.size, sa.ctypes.data, mode)
    return sa[:a.size], int(shift)


def least_rotation(data, fast):
    a = _buf(data)
    return int(lib().orc_least_rotation(a.ctypes.data, a.size, 1 if fast else 0))


def huffman(freq, lim, kind):
    f = np.ascontiguousarray(freq, dtype=np.uint64)
    out = np.zeros(max(f.size, 1), dtype=np.uint8)
    lm = C.c_int(0)
    n = lib().orc_huffman(f.ctypes.data, f.size, lim, kind, out.ctypes.data, C.byref(lm))
    return out[:n].copy(), bool(lm.value)


def canonical_codes(lens):
    l = np.ascontiguousarray(lens, dtype=np.uint8)
    c = np.zeros(l.size, dtype=np.uint32)
    lib().orc_canonical_codes(l.ctypes.data, l.size, c.ctypes.data)
    return c


def pack_bits(fields):
    v = np.array([f[0] for f in fields], dtype=np.uint32)
    l = np.array([f[1] for f in fields], dtype=np.uint32)
    out = np.zeros(len(fields) * 4 + 8, dtype=np.uint8)
    n = lib().orc_pack_bits(v.ctypes.data, l.ctypes.data, len(fields), out.ctypes.data, out.size)
    return out[:n].tobytes()


def crc32_bzip2(data):
    a = _buf(data)
    return int(lib().orc_crc32_bzip2(a.ctypes.data, a.size))


def mtf_positions(syms, k):
    a = _buf(syms)
    o = np.zeros(a.size, dtype=np.uint8)
    lib().orc_mtf_positions(a.ctypes.data, a.size, k, o.ctypes.data)
    return o


class DecodeError(Exception):
    """BZip2Error of the restated reference decoder (bzip2/error.rs:13-19)."""
    KINDS = {1: "DataError", 2: "DataErrorMagicFirst", 3: "DataErrorMagic", 4: "UnexpectedEof", 5: "Unexpected",
             6: "NonTerminating"}  # 6: not a BZip2Error — the reference never stops on this input (see the .cpp)

    def __init__(self, code, partial):
        super().__init__(self.KINDS.get(code, str(code)))
        self.kind = self.KINDS.get(code, str(code))
        self.partial = partial


def decode(data):
    """Restatement of BZip2Decoder (oracle/bz2_decoder_oracle.cpp): multi-stream .bz2 bytes -> original bytes."""
    b = _buf(data)
    out = C.c_void_p()
    n = C.c_size_t(0)
    rc = lib().orc_decode(b.ctypes.data if b.size else None, b.size, C.byref(out), C.byref(n))
    res = C.string_at(out, n.value) if n.value else b""
    lib().orc_decode_free(out)
    if rc != 0:
        raise DecodeError(rc, res)
    return res
