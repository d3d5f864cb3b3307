"""rust-compression_b200 — B200-native drop-in for the bzip2 encode path of chalharu/rust-compression.

Public surface mirrors the reference's prelude for this path (src/lib.rs:70-117): BZip2Encoder, Action,
CompressionError, encode() (= EncodeExt::encode), and — SURVEY.md section 8(f).1 — BZip2Decoder, BZip2Error,
decode() (= DecodeExt::decode).  `device`/`sharded` expose the device-resident job API used for
block-wise multi-GPU sharding and benchmarking.  All compute is in libbzb200.so (hand-written sm_100a CUDA behind
the C ABI of include/bzb200.h); importing this package without that library fails.
"""
from . import _lib
from .decoder import BZip2Decoder, BZip2Error, DecodeIterator, decode, decompress
from .encoder import Action, BZip2Encoder, CompressionError, EncodeIterator, compress, encode

__all__ = ["Action", "BZip2Decoder", "BZip2Encoder", "BZip2Error", "CompressionError", "DecodeIterator",
           "EncodeIterator", "compress", "decode", "decompress", "encode", "lib_path", "version"]


def lib_path():
    return _lib.LIB_PATH


def version():
    return _lib.lib().bzb200_version().decode()


_lib.lib()  # fail loudly at import time if the CUDA library is missing
