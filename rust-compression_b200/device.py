"""Device-resident job API (section 2/3 of include/bzb200.h) over torch tensors.

torch is plumbing here: it owns device memory, the CUDA stream and (in sharded.py) torch.distributed. Every
stage runs in libbzb200.so.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .encoder import CompressionError

FIELDS = {
    "rle": (0, np.uint8), "rank": (1, np.uint32), "last": (2, np.uint8), "mtf": (3, np.uint16), "freq": (4, np.uint32),
    "sel": (5, np.uint8), "len0": (6, np.uint8), "len1": (7, np.uint8), "len2": (8, np.uint8), "len3": (9, np.uint8),
    "len4": (10, np.uint8), "info": (11, np.uint64),
}
INFO_KEYS = ["in_start", "in_end", "nblock", "crc", "orig_ptr", "mtf_count", "alpha", "ngroups", "nselectors",
             "bit_start", "bit_end", "sort_rounds", "periodic", "lm_used", "dev_error"]


def max_output_bytes(level, n):
    return int(_lib.lib().bzb200_max_output_bytes(level, n))


class Context:
    """One context = one GPU + one stream (defaults to torch's current stream on the current device)."""

    def __init__(self, device=None, stream=None):
        L = _lib.lib()
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index)
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        self.stream = stream
        self._h = C.c_void_p()
        rc = L.bzb200_ctx_create(self.device.index, C.c_void_p(stream.cuda_stream), C.byref(self._h))
        if rc != _lib.OK:
            msg = L.bzb200_last_error(self._h).decode() if self._h else ""
            if self._h:
                L.bzb200_ctx_destroy(self._h)
                self._h = None
            raise CompressionError("Unexpected", f"bzb200_ctx_create rc={rc}: {msg}")
        self.level = None
        self.nblocks = 0

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().bzb200_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != _lib.OK:
            msg = _lib.lib().bzb200_last_error(self._h)
            if rc == _lib.E_LEVEL:
                raise ValueError("invalid level")
            raise CompressionError("Unexpected", f"{what} rc={rc}: {msg.decode() if msg else ''}")

    @staticmethod
    def _u8(t):
        assert t.is_cuda and t.dtype == torch.uint8 and t.is_contiguous()
        return t

    def sync(self):
        self._check(_lib.lib().bzb200_sync(self._h), "bzb200_sync")

    def plan(self, level, d_in):
        self._u8(d_in)
        nb = C.c_uint32(0)
        self._check(_lib.lib().bzb200_plan(self._h, level, C.c_void_p(d_in.data_ptr()), d_in.numel(), C.byref(nb)),
                    "bzb200_plan")
        self._keep_in = d_in
        self.level = level
        self.nblocks = nb.value
        return nb.value

    # ---- sliced plan (include/bzb200.h section 2b): this context holds one slice of the stream
    def slice_begin(self, level, n_total, lo, hi, d_buf, buf_lo_offset, avail_hi, reserve_hi):
        """d_buf: device uint8 tensor that holds input byte `lo` at index buf_lo_offset (16-byte aligned address), the 16
        bytes in front of it when lo > 0, and the bytes up to avail_hi behind it.  Returns the slice's last run head."""
        self._u8(d_buf)
        lh = C.c_int64(-1)
        self._check(_lib.lib().bzb200_slice_begin(self._h, level, n_total, lo, hi,
                                                  C.c_void_p(d_buf.data_ptr() + buf_lo_offset), avail_hi, reserve_hi,
                                                  C.byref(lh)), "bzb200_slice_begin")
        self._keep_in = d_buf
        self.level = level
        return lh.value

    def slice_counts(self, carry_in):
        em = C.c_uint64(0)
        self._check(_lib.lib().bzb200_slice_counts(self._h, carry_in, C.byref(em)), "bzb200_slice_counts")
        return em.value

    def slice_prefix(self, e_lo, e_tot):
        self._check(_lib.lib().bzb200_slice_prefix(self._h, e_lo, e_tot), "bzb200_slice_prefix")

    def slice_windows(self, x0, out_rows):
        """Tabulates the slice's cut windows of the phase starting at x0 into the leading rows of out_rows (device int64
        tensor [rows, bzb200_cut_window()], stream ordered).  Returns (j0, nj)."""
        L = _lib.lib()
        j0, nj, d_f = C.c_uint64(0), C.c_uint32(0), C.c_void_p()
        self._check(L.bzb200_slice_windows(self._h, x0, C.byref(j0), C.byref(nj), C.byref(d_f)), "bzb200_slice_windows")
        if nj.value:
            assert out_rows.is_cuda and out_rows.dtype == torch.int64 and out_rows.is_contiguous()
            assert out_rows.shape[0] >= nj.value and out_rows.shape[1] == cut_window()
            with torch.cuda.stream(self.stream):
                src = _as_tensor(d_f.value, nj.value * cut_window(), self.device)
                out_rows.view(-1)[:nj.value * cut_window()].copy_(src, non_blocking=True)
        return int(j0.value), int(nj.value)

    def slice_set_blocks(self, in_off, rle_off, max_block_len):
        a = np.ascontiguousarray(in_off, dtype=np.uint64)
        b = np.ascontiguousarray(rle_off, dtype=np.uint64)
        self._check(_lib.lib().bzb200_slice_set_blocks(self._h, a.size - 1, a.ctypes.data, b.ctypes.data, max_block_len),
                    "bzb200_slice_set_blocks")
        self.nblocks = a.size - 1

    def slice_blocks(self):
        b0, b1, need = C.c_uint32(0), C.c_uint32(0), C.c_uint64(0)
        self._check(_lib.lib().bzb200_slice_blocks(self._h, C.byref(b0), C.byref(b1), C.byref(need)), "bzb200_slice_blocks")
        return b0.value, b1.value, need.value

    def slice_extend(self, avail_hi):
        self._check(_lib.lib().bzb200_slice_extend(self._h, avail_hi), "bzb200_slice_extend")

    def block_table(self, with_crc=True):
        """(in_off, rle_off, crc) of the planned blocks.  CRCs are computed for the blocks this context encodes;
        with_crc=True asks for all of them (the missing ones are computed on demand), with_crc=False returns the
        array as it stands (zeros for blocks not encoded here) without touching the device."""
        nb = self.nblocks = int(_lib.lib().bzb200_num_blocks(self._h))
        in_off = np.zeros(nb + 1, dtype=np.uint64)
        rle_off = np.zeros(nb + 1, dtype=np.uint64)
        crc = np.zeros(max(nb, 1), dtype=np.uint32)
        self._check(_lib.lib().bzb200_block_table(self._h, in_off.ctypes.data, rle_off.ctypes.data,
                                                  crc.ctypes.data if with_crc else None), "bzb200_block_table")
        if not with_crc and nb:
            self._check(_lib.lib().bzb200_block_crcs(self._h, crc.ctypes.data, nb), "bzb200_block_crcs")
        return in_off, rle_off, crc[:nb]

    def encode_blocks(self, b0, b1, d_out, start_bit):
        self._u8(d_out)
        end = C.c_uint64(0)
        self._check(_lib.lib().bzb200_encode_blocks(self._h, b0, b1, C.c_void_p(d_out.data_ptr()), d_out.numel(),
                                                    start_bit, C.byref(end)), "bzb200_encode_blocks")
        return end.value

    def bit_append(self, d_dst, dst_bit, d_src, nbits):
        self._u8(d_dst)
        self._u8(d_src)
        self._check(_lib.lib().bzb200_bit_append(self._h, C.c_void_p(d_dst.data_ptr()), d_dst.numel(), dst_bit,
                                                 C.c_void_p(d_src.data_ptr()), nbits), "bzb200_bit_append")

    def write_stream_header(self, level, d_out):
        self._check(_lib.lib().bzb200_write_stream_header(self._h, level, C.c_void_p(d_out.data_ptr()), d_out.numel()),
                    "bzb200_write_stream_header")

    def write_stream_trailer(self, d_out, at_bit, combined_crc):
        total = C.c_size_t(0)
        self._check(_lib.lib().bzb200_write_stream_trailer(self._h, C.c_void_p(d_out.data_ptr()), d_out.numel(), at_bit,
                                                           combined_crc, C.byref(total)), "bzb200_write_stream_trailer")
        return total.value

    @staticmethod
    def combine_crc(crcs, seed=0):
        a = np.ascontiguousarray(crcs, dtype=np.uint32)
        return int(_lib.lib().bzb200_combine_crc(seed, a.ctypes.data, a.size))

    def compress_device(self, level, d_in, d_out):
        """Whole stream, device in -> device out; d_out must be zero-filled. Returns the stream length in bytes."""
        self._u8(d_in)
        self._u8(d_out)
        n = C.c_size_t(0)
        self._check(_lib.lib().bzb200_compress_device(self._h, level, C.c_void_p(d_in.data_ptr()), d_in.numel(),
                                                      C.c_void_p(d_out.data_ptr()), d_out.numel(), C.byref(n)),
                    "bzb200_compress_device")
        self.level = level
        return n.value

    def compress_host(self, level, h_in, h_out):
        """Whole stream, host tensor in -> host tensor out (pinned tensors give async DMA). Returns bytes written."""
        assert (not h_in.is_cuda) and (not h_out.is_cuda) and h_in.dtype == torch.uint8 and h_out.dtype == torch.uint8
        n = C.c_size_t(0)
        self._check(_lib.lib().bzb200_compress_host(self._h, level, C.c_void_p(h_in.data_ptr()), h_in.numel(),
                                                    C.c_void_p(h_out.data_ptr()), h_out.numel(), C.byref(n)),
                    "bzb200_compress_host")
        return n.value

    # ---- decoder (include/bzb200.h section 3) ----
    def decompress_device(self, d_in, d_out):
        """Device tensor holding a .bz2 buffer -> original bytes in d_out.  Returns (bytes written, BZip2Error code
        or 0); raises ValueError(needed) when d_out is too small."""
        self._u8(d_in)
        self._u8(d_out)
        n, kind = C.c_size_t(0), C.c_int(0)
        rc = _lib.lib().bzb200_decompress_device(self._h, C.c_void_p(d_in.data_ptr()), d_in.numel(),
                                                 C.c_void_p(d_out.data_ptr()), d_out.numel(), C.byref(n), C.byref(kind))
        if rc == _lib.E_ARG and n.value > d_out.numel():
            raise ValueError(n.value)
        if rc not in (_lib.OK, _lib.E_DATA):
            self._check(rc, "bzb200_decompress_device")
        return n.value, kind.value

    def decompress_host(self, h_in, h_out):
        """Host tensor in -> host tensor out through bzb200_decompress_host; same return convention."""
        assert (not h_in.is_cuda) and (not h_out.is_cuda) and h_in.dtype == torch.uint8 and h_out.dtype == torch.uint8
        n, kind = C.c_size_t(0), C.c_int(0)
        rc = _lib.lib().bzb200_decompress_host(self._h, C.c_void_p(h_in.data_ptr()), h_in.numel(),
                                               C.c_void_p(h_out.data_ptr()), h_out.numel(), C.byref(n), C.byref(kind))
        if rc == _lib.E_ARG and n.value > h_out.numel():
            raise ValueError(n.value)
        if rc not in (_lib.OK, _lib.E_DATA):
            self._check(rc, "bzb200_decompress_host")
        return n.value, kind.value

    def dec_stats(self):
        v = (C.c_uint64 * 8)()
        _lib.lib().bzb200_dec_stats(self._h, v, 8)
        keys = ["streams", "blocks", "candidates", "batches", "symbols", "pre_rle_bytes", "out_bytes"]
        return {k: int(v[i]) for i, k in enumerate(keys)}

    # ---- instrumentation ----
    def debug_stage(self, block, name):
        fid, dt = FIELDS[name]
        cnt = C.c_size_t(0)
        self._check(_lib.lib().bzb200_debug_stage(self._h, block, fid, None, 0, C.byref(cnt)), "bzb200_debug_stage")
        out = np.zeros(max(cnt.value, 1), dtype=dt)
        self._check(_lib.lib().bzb200_debug_stage(self._h, block, fid, out.ctypes.data, out.size, C.byref(cnt)),
                    "bzb200_debug_stage")
        out = out[:cnt.value]
        if name == "info":
            d = {k: int(out[i]) for i, k in enumerate(INFO_KEYS)}
            d["in_use"] = out[16:24].astype(np.uint32)
            return d
        return out

    def profile(self, on):
        self._check(_lib.lib().bzb200_profile(self._h, 1 if on else 0), "bzb200_profile")

    def profile_records(self):
        L = _lib.lib()
        recs = {}
        for i in range(L.bzb200_profile_count(self._h)):
            name = C.c_char_p()
            n = C.c_uint64(0)
            ms = C.c_double(0)
            L.bzb200_profile_get(self._h, i, C.byref(name), C.byref(n), C.byref(ms))
            recs[name.value.decode()] = (int(n.value), float(ms.value))
        return recs

    def launch_count(self):
        return int(_lib.lib().bzb200_launch_count(self._h))

    def sort_stats(self):
        r, p, e = C.c_uint32(0), C.c_uint32(0), C.c_uint64(0)
        _lib.lib().bzb200_sort_stats(self._h, C.byref(r), C.byref(p), C.byref(e))
        v = (C.c_uint64 * 8)()
        _lib.lib().bzb200_path_stats(self._h, v, 8)
        return {"rounds": r.value, "radix_passes": p.value, "elems_sorted": e.value,
                "elems_sorted_radix": int(v[3]), "elems_local": int(v[4]), "n_rle": int(v[5]),
                "mtf_symbols": int(v[6])}


def cut_window():
    return int(_lib.lib().bzb200_cut_window())


def slice_halo_bytes():
    return int(_lib.lib().bzb200_slice_halo_bytes())


def plan_tile_bytes():
    return int(_lib.lib().bzb200_plan_tile_bytes())


def cut_walk(F, T, e_tot, n, state, in_off, rle_off):
    """bzb200_cut_walk (host): one phase of the cut chain through the window rows F (numpy uint64 [K, cut_window()]).
    state / in_off / rle_off are numpy uint64 arrays updated in place.  Returns (nblocks, max_block_len) once
    state[2] == 1."""
    nb, ml = C.c_uint32(0), C.c_uint32(0)
    F = np.ascontiguousarray(F, dtype=np.uint64)
    rc = _lib.lib().bzb200_cut_walk(F.ctypes.data if F.size else None, F.shape[0] if F.ndim == 2 else 0, T, e_tot, n,
                                    in_off.size - 1, state.ctypes.data, in_off.ctypes.data, rle_off.ctypes.data,
                                    C.byref(nb), C.byref(ml))
    if rc != _lib.OK:
        raise CompressionError("Unexpected", f"bzb200_cut_walk rc={rc}")
    return nb.value, ml.value


def _as_tensor(ptr, n_int64, device):
    """A torch view (int64) of device memory owned by the library (no copy)."""
    class _Ext:
        pass
    e = _Ext()
    e.__cuda_array_interface__ = {"shape": (n_int64,), "typestr": "<i8", "data": (ptr, False), "version": 2}
    return torch.as_tensor(e, device=device)


class Pool:
    """The in-library multi-GPU engine (include/bzb200.h section 2c): host buffers in, host buffers out, one process."""

    def __init__(self, devices):
        L = _lib.lib()
        devs = (C.c_int * len(devices))(*devices)
        self._h = C.c_void_p()
        rc = L.bzb200_pool_create(len(devices), devs, C.byref(self._h))
        if rc != _lib.OK:
            self._h = None
            raise CompressionError("Unexpected", f"bzb200_pool_create rc={rc}")
        self.devices = list(devices)

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().bzb200_pool_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def size(self):
        return int(_lib.lib().bzb200_pool_size(self._h))

    def compress_host(self, level, h_in, h_out):
        assert (not h_in.is_cuda) and (not h_out.is_cuda) and h_in.dtype == torch.uint8 and h_out.dtype == torch.uint8
        n = C.c_size_t(0)
        rc = _lib.lib().bzb200_pool_compress_host(self._h, level, C.c_void_p(h_in.data_ptr()), h_in.numel(),
                                                  C.c_void_p(h_out.data_ptr()), h_out.numel(), C.byref(n))
        if rc == _lib.E_LEVEL:
            raise ValueError("invalid level")
        if rc != _lib.OK:
            msg = _lib.lib().bzb200_pool_last_error(self._h)
            raise CompressionError("Unexpected", f"bzb200_pool_compress_host rc={rc}: {msg.decode() if msg else ''}")
        return n.value

    def stats(self):
        v = (C.c_uint64 * 4)()
        _lib.lib().bzb200_pool_stats(self._h, v, 4)
        return {"spans": int(v[0]), "blocks": int(v[1]), "phases": int(v[2]), "launches": int(v[3])}


def compress_tensor(ctx, level, d_in):
    """Convenience: device tensor in -> device tensor holding exactly the .bz2 stream."""
    cap = max_output_bytes(level, d_in.numel())
    d_out = torch.zeros(cap, dtype=torch.uint8, device=d_in.device)
    n = ctx.compress_device(level, d_in, d_out)
    return d_out[:n]
