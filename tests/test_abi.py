"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/bzb200.h
declares, and its host-only entry points behave (no compute calls without a GPU)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    h = open(os.path.join(ROOT, "include", "bzb200.h")).read()
    return sorted(set(re.findall(r"BZB200_API[^;(]*?\b(bzb200_\w+)\s*\(", h)))


def test_library_exports_every_declared_symbol():
    import rust_compression_b200 as rc
    from rust_compression_b200 import _lib
    syms = _header_symbols()
    assert len(syms) >= 30
    L = C.CDLL(rc.lib_path())
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/bzb200.h but not exported"
    assert set(syms) == set(_lib.SIGNATURES), "ctypes table and header disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", rc.lib_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (bzb200_\w+)", out))
    assert exported == set(syms), f"exported but undeclared: {exported - set(syms)}; missing: {set(syms) - exported}"


def test_level_validation_needs_no_gpu():
    from rust_compression_b200 import _lib
    import rust_compression_b200 as rc
    L = _lib.lib()
    h = C.c_void_p()
    assert L.bzb200_enc_create(0, -1, C.byref(h)) == _lib.E_LEVEL
    assert L.bzb200_enc_create(10, -1, C.byref(h)) == _lib.E_LEVEL
    with pytest.raises(ValueError):
        rc.BZip2Encoder(0)
    with pytest.raises(ValueError):
        rc.compress(b"x", 11)


def test_host_helpers():
    from rust_compression_b200 import _lib
    L = _lib.lib()
    crcs = np.array([0x633ED6E2, 0x12345678, 0xFFFFFFFF], dtype=np.uint32)
    want = 0
    for c in crcs:
        want = (((want << 1) | (want >> 31)) & 0xFFFFFFFF) ^ int(c)
    assert L.bzb200_combine_crc(0, crcs.ctypes.data, 3) == want
    assert L.bzb200_max_output_bytes(9, 1 << 20) > (1 << 20)
    assert b"sm_100a" in L.bzb200_version()


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product path must fail loudly, never fall back to a CPU implementation."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import rust_compression_b200 as rc
    with pytest.raises(rc.CompressionError):
        rc.compress(b"hello", 9)
    enc = rc.BZip2Encoder(9)
    with pytest.raises(rc.CompressionError):
        list(rc.encode(b"hello", enc, rc.Action.Finish))
    # the decoder as well: a CUDA failure surfaces as BZip2Error::Unexpected, not as decoded bytes
    stream = bytes.fromhex("425a683917724538509000000000")
    with pytest.raises(rc.BZip2Error) as ei:
        rc.decompress(stream)
    assert ei.value.kind == "Unexpected"
    with pytest.raises(rc.BZip2Error) as ei:
        list(rc.decode(stream, rc.BZip2Decoder()))
    assert ei.value.kind == "Unexpected"


def test_decoder_emulation_is_test_infrastructure_only():
    """tests/cpp/dec_emu.cpp compiles the decoder's kernel bodies for the host (-DBZB_EMU) to check the algorithm without
    a GPU.  None of that may be in the product: the library is built without BZB_EMU and exports no emulation entry."""
    import rust_compression_b200 as rc
    build_py = open(os.path.join(ROOT, "rust-compression_b200", "build.py")).read()
    assert "BZB_EMU" not in build_py
    out = subprocess.run(["nm", "-D", "--defined-only", rc.lib_path()], capture_output=True, text=True).stdout
    assert "emu_" not in out
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rust-compression_b200")):
        for f in files:
            if f.endswith(".py"):
                assert "libdecemu" not in open(os.path.join(dirpath, f)).read(), f


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing in the product package may import, link or execute it."""
    pkg = os.path.join(ROOT, "rust-compression_b200")
    pat = re.compile(r"(from|import)\s+oracle|liborc|\borc_\w+\s*\(|oracle/|bz2_oracle")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".rs")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                m = pat.search(txt)
                assert not m, f"{f} references the oracle: {m.group(0)}"
    out = subprocess.run(["ldd", os.path.join(pkg, "libbzb200.so")], capture_output=True, text=True).stdout
    assert "liborc" not in out
