"""Host-side mirror of the reference's public API for the bzip2 DEcode path.

Reference (chalharu/rust-compression):
  BZip2Decoder::new / Default   src/bzip2/decoder.rs:584-601
  Decoder::next                 src/bzip2/decoder.rs:603-615   (one output byte per call, None = drained)
  DecodeExt::decode             src/traits/decoder.rs:14-43    (iterator adapter)
  BZip2Error                    src/bzip2/error.rs:4-52        (kinds, messages, mapping onto CompressionError)

Semantics kept: the byte sequence the reference yields, and — for a malformed buffer — the bytes it yields before it
fails followed by the same BZip2Error kind.  Timing differs: the reference decodes while it pulls input, this mirror
drains the input iterator first and decodes all blocks at once on the GPU (libbzb200.so; there is no CPU path).
"""
import ctypes as C

from . import _lib
from .encoder import CompressionError


class BZip2Error(Exception):
    """bzip2/error.rs:4-11; `kind` is the variant name, str() is description_in() (error.rs:31-41)."""
    KINDS = {1: "DataError", 2: "DataErrorMagicFirst", 3: "DataErrorMagic", 4: "UnexpectedEof", 5: "Unexpected"}
    MESSAGES = {
        "DataError": "data integrity (CRC) error in data",
        "DataErrorMagicFirst": "bad magic number (file not created by bzip2)",
        "DataErrorMagic": "trailing garbage after EOF ignored",
        "UnexpectedEof": "file ends unexpectedly",
        "Unexpected": "unexpected error",
    }

    def __init__(self, kind, partial=b"", detail=""):
        if isinstance(kind, int):
            kind = self.KINDS.get(kind, "Unexpected")
        self.kind = kind
        self.partial = partial  # the bytes the reference yields before the error
        self.detail = detail
        super().__init__(self.MESSAGES[kind] + (f" ({detail})" if detail else ""))

    def to_compression_error(self):
        """impl From<BZip2Error> for CompressionError (error.rs:44-52)."""
        if self.kind == "UnexpectedEof":
            return CompressionError("UnexpectedEof")
        if self.kind == "Unexpected":
            return CompressionError("Unexpected")
        return CompressionError("DataError")


class BZip2Decoder:
    def __init__(self, device=-1):
        self._h = C.c_void_p()
        rc = _lib.lib().bzb200_dec_create(device, C.byref(self._h))
        if rc != _lib.OK:
            raise BZip2Error("Unexpected", detail=f"bzb200_dec_create rc={rc}")
        self._chunk = b""
        self._pos = 0
        self._decoded = False
        self._pending_error = None

    def __del__(self):
        try:
            if self._h:
                _lib.lib().bzb200_dec_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _run(self):
        L = _lib.lib()
        rc = L.bzb200_dec_finish(self._h)
        if rc not in (_lib.OK, _lib.E_DATA):
            msg = L.bzb200_dec_last_error(self._h)
            raise BZip2Error("Unexpected", detail=f"bzb200_dec_finish rc={rc}: {msg.decode() if msg else ''}")
        n = L.bzb200_dec_output_size(self._h)
        buf = C.create_string_buffer(max(n, 1))
        got = L.bzb200_dec_read(self._h, buf, n) if n else 0
        kind = L.bzb200_dec_error_kind(self._h)
        L.bzb200_dec_reset(self._h)
        return buf.raw[:got], kind

    def decode_all(self, data):
        """Bulk form: the whole .bz2 buffer in, the original bytes out; raises BZip2Error (with .partial)."""
        L = _lib.lib()
        b = bytes(data)
        rc = L.bzb200_dec_write(self._h, b, len(b))
        if rc != _lib.OK:
            raise BZip2Error("Unexpected", detail=f"bzb200_dec_write rc={rc}")
        out, kind = self._run()
        if kind:
            raise BZip2Error(kind, partial=out)
        return out

    def next(self, it):
        """Decoder::next (decoder.rs:607-614): the next output byte (int), None when drained; raises BZip2Error at
        the point where the reference returns Some(Err(..))."""
        if self._pos < len(self._chunk):
            b = self._chunk[self._pos]
            self._pos += 1
            return b
        if self._decoded:
            err, self._pending_error = self._pending_error, None
            self._decoded = False
            self._chunk, self._pos = b"", 0
            if err is not None:
                raise err
            return None
        L = _lib.lib()
        pending = bytearray()

        def push():
            rc = L.bzb200_dec_write(self._h, bytes(pending), len(pending))
            if rc != _lib.OK:  # e.g. E_STATE after a failed finish: input must not be dropped silently
                msg = L.bzb200_dec_last_error(self._h)
                L.bzb200_dec_reset(self._h)
                raise BZip2Error("Unexpected", detail=f"bzb200_dec_write rc={rc}: {msg.decode() if msg else ''}")
            pending.clear()

        for x in it:
            pending.append(x)
            if len(pending) >= (1 << 20):
                push()
        if pending:
            push()
        out, kind = self._run()
        self._chunk, self._pos = out, 0
        self._decoded = True
        self._pending_error = BZip2Error(kind, partial=out) if kind else None
        return self.next(it)


class DecodeIterator:
    """traits/decoder.rs:45-99."""

    def __init__(self, inner, decoder):
        self.inner = iter(inner)
        self.decoder = decoder

    def __iter__(self):
        return self

    def __next__(self):
        b = self.decoder.next(self.inner)
        if b is None:
            raise StopIteration
        return b


def decode(iterable, decoder):
    """DecodeExt::decode (traits/decoder.rs:27-43): `iterable.decode(&mut decoder)`."""
    return DecodeIterator(iterable, decoder)


def decompress(data, device=-1):
    """One-shot host->host through bzb200_decompress; raises BZip2Error (with .partial) like the reference fails."""
    L = _lib.lib()
    b = bytes(data)
    out = C.c_void_p()
    out_n = C.c_size_t(0)
    kind = C.c_int(0)
    rc = L.bzb200_decompress(device, b, len(b), C.byref(out), C.byref(out_n), C.byref(kind))
    if rc not in (_lib.OK, _lib.E_DATA):
        raise BZip2Error("Unexpected", detail=f"bzb200_decompress rc={rc}")
    try:
        res = C.string_at(out, out_n.value) if out_n.value else b""
    finally:
        L.bzb200_free(out)
    if rc == _lib.E_DATA:
        raise BZip2Error(kind.value, partial=res)
    return res
