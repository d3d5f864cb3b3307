"""Host-side mirror of the reference's public API for the bzip2 encode path.

Reference (chalharu/rust-compression):
  Action                 src/action.rs:8-13
  CompressionError       src/error.rs:10-42
  BZip2Encoder::new      src/bzip2/encoder.rs:58-72      (panics "invalid level" unless 1 <= level <= 9)
  Encoder::next          src/bzip2/encoder.rs:116-158    (one output byte per call, None = drained)
  EncodeExt::encode      src/traits/encoder.rs:12-39     (iterator adapter)

Semantics kept: with Action.Finish, `next` yields every byte of the .bz2 stream and then None, after which the
encoder is re-armed (encoder.rs:87-90,130-133); with Action.Run, None means "input drained, feed more".
Action.Flush is out of contract and behaves like Run.  Only the concatenated byte sequence is
guaranteed, not the call at which each byte appears: closed blocks are compressed on the GPU whenever a window of
input (256 MiB, env BZB200_ENC_WINDOW) has accumulated, the rest when Finish arrives.
All compute happens in libbzb200.so (CUDA, sm_100a); there is no CPU path.
"""
import ctypes as C
import enum

from . import _lib


class Action(enum.Enum):
    Run = 0
    Flush = 1
    Finish = 2


class CompressionError(Exception):
    """error.rs:10-15. `kind` is one of DataError, UnexpectedEof, Unexpected."""
    DataError = "data integrity error in data"
    UnexpectedEof = "file ends unexpectedly"
    Unexpected = "unexpected error"

    def __init__(self, kind="Unexpected", detail=""):
        self.kind = kind
        self.detail = detail
        super().__init__(getattr(CompressionError, kind) + (f" ({detail})" if detail else ""))


class BZip2Encoder:
    def __init__(self, level=9, device=-1, devices=None):
        """devices: list of CUDA ordinals — the same object sharding every window of input block-wise over several GPUs
        of the box (bzb200_enc_create_multi); default: one GPU (`device`, -1 = the current one)."""
        if level < 1 or level > 9:
            raise ValueError("invalid level")  # the reference panics (encoder.rs:59-61)
        self.level = level
        self._h = C.c_void_p()
        L = _lib.lib()
        if devices:
            arr = (C.c_int * len(devices))(*devices)
            rc = L.bzb200_enc_create_multi(level, len(devices), arr, C.byref(self._h))
        else:
            rc = L.bzb200_enc_create(level, device, C.byref(self._h))
        if rc != _lib.OK:
            raise CompressionError("Unexpected", f"bzb200_enc_create rc={rc}")
        self._finished = False   # encoder.rs:45 `finished`
        self._chunk = b""
        self._pos = 0

    def __del__(self):
        try:
            if self._h:
                _lib.lib().bzb200_enc_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _err(self, what, rc):
        msg = _lib.lib().bzb200_enc_last_error(self._h)
        return CompressionError("Unexpected", f"{what} rc={rc}: {msg.decode() if msg else ''}")

    def write(self, data):
        """Bulk form of feeding input under Action.Run."""
        L = _lib.lib()
        b = bytes(data)
        rc = L.bzb200_enc_write(self._h, b, len(b))
        if rc != _lib.OK:
            raise self._err("bzb200_enc_write", rc)

    def finish(self):
        """Bulk form of Action.Finish; returns the bytes of the .bz2 stream that have not been read yet (the whole
        stream unless read_available()/next() already handed out the blocks of earlier windows)."""
        L = _lib.lib()
        rc = L.bzb200_enc_finish(self._h)
        if rc != _lib.OK:
            raise self._err("bzb200_enc_finish", rc)
        n = L.bzb200_enc_output_size(self._h)
        buf = C.create_string_buffer(n)
        got = L.bzb200_enc_read(self._h, buf, n)
        L.bzb200_enc_reset(self._h)
        return buf.raw[:got]

    def stats(self):
        v = (C.c_uint64 * 4)()
        _lib.lib().bzb200_enc_stats(self._h, v, 4)
        return {"blocks": int(v[0]), "windows": int(v[1]), "buffered_in": int(v[2]), "ready_out": int(v[3])}

    def read_available(self):
        """Output bytes that are ready now (closed blocks of full windows under Action.Run; everything after finish)."""
        L = _lib.lib()
        n = L.bzb200_enc_output_size(self._h)
        if not n:
            return b""
        buf = C.create_string_buffer(n)
        got = L.bzb200_enc_read(self._h, buf, n)
        return buf.raw[:got]

    def next(self, it, action):
        """Encoder::next (encoder.rs:120-158): returns the next output byte (int) or None."""
        L = _lib.lib()
        if self._pos < len(self._chunk):
            b = self._chunk[self._pos]
            self._pos += 1
            return b
        if not self._finished:
            # drain the input iterator (the reference pulls one byte at a time, encoder.rs:79-85)
            pending = bytearray()
            for x in it:
                pending.append(x)
                if len(pending) >= (1 << 20):
                    self.write(pending)
                    pending.clear()
            if pending:
                self.write(pending)
            if action is Action.Finish:
                rc = L.bzb200_enc_finish(self._h)
                if rc != _lib.OK:
                    raise self._err("bzb200_enc_finish", rc)
                self._finished = True
        chunk = self.read_available()  # under Run: the blocks that closed so far; after Finish: the rest
        if chunk:
            self._chunk, self._pos = chunk, 0
            return self.next(it, action)
        self._chunk, self._pos = b"", 0
        if self._finished:
            # stream fully drained: reset the latches like the reference (encoder.rs:87-90,130-133)
            self._finished = False
            L.bzb200_enc_reset(self._h)
        return None  # Run/Flush: input drained, nothing (more) to hand out yet


class EncodeIterator:
    """traits/encoder.rs:41-79."""

    def __init__(self, inner, encoder, action):
        self.inner = iter(inner)
        self.encoder = encoder
        self.action = action

    def __iter__(self):
        return self

    def __next__(self):
        b = self.encoder.next(self.inner, self.action)
        if b is None:
            raise StopIteration
        return b


def encode(iterable, encoder, action):
    """EncodeExt::encode (traits/encoder.rs:29-38): `iterable.encode(&mut encoder, action)`."""
    return EncodeIterator(iterable, encoder, action)


def compress(data, level=9, device=-1):
    """One-shot host->host through bzb200_compress (lib.rs:13-33 doctest shape)."""
    if level < 1 or level > 9:
        raise ValueError("invalid level")
    L = _lib.lib()
    b = bytes(data)
    out = C.c_void_p()
    out_n = C.c_size_t(0)
    rc = L.bzb200_compress(level, device, b, len(b), C.byref(out), C.byref(out_n))
    if rc != _lib.OK:
        raise CompressionError("Unexpected", f"bzb200_compress rc={rc}")
    try:
        return C.string_at(out, out_n.value)
    finally:
        L.bzb200_free(out)
