"""ctypes binding of libbzb200.so (include/bzb200.h). No fallback: if the library is missing, importing fails."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbzb200.so")

OK, E_LEVEL, E_CUDA, E_ARG, E_STATE, E_INTERNAL, E_DATA = 0, -1, -2, -3, -4, -5, -6

# every symbol include/bzb200.h declares: (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "bzb200_enc_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(_P)]),
    "bzb200_enc_write": (C.c_int, [_P, _P, C.c_size_t]),
    "bzb200_enc_finish": (C.c_int, [_P]),
    "bzb200_enc_read": (C.c_size_t, [_P, _P, C.c_size_t]),
    "bzb200_enc_output_size": (C.c_size_t, [_P]),
    "bzb200_enc_reset": (C.c_int, [_P]),
    "bzb200_enc_stats": (C.c_int, [_P, _P, C.c_size_t]),
    "bzb200_enc_destroy": (None, [_P]),
    "bzb200_enc_last_error": (C.c_char_p, [_P]),
    "bzb200_compress": (C.c_int, [C.c_int, C.c_int, _P, C.c_size_t, C.POINTER(_P), C.POINTER(C.c_size_t)]),
    "bzb200_free": (None, [_P]),
    "bzb200_ctx_create": (C.c_int, [C.c_int, _P, C.POINTER(_P)]),
    "bzb200_ctx_destroy": (None, [_P]),
    "bzb200_last_error": (C.c_char_p, [_P]),
    "bzb200_sync": (C.c_int, [_P]),
    "bzb200_plan": (C.c_int, [_P, C.c_int, _P, C.c_size_t, C.POINTER(C.c_uint32)]),
    "bzb200_num_blocks": (C.c_uint32, [_P]),
    "bzb200_block_table": (C.c_int, [_P, _P, _P, _P]),
    "bzb200_encode_blocks": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P, C.c_size_t, C.c_uint64, C.POINTER(C.c_uint64)]),
    "bzb200_bit_append": (C.c_int, [_P, _P, C.c_size_t, C.c_uint64, _P, C.c_uint64]),
    "bzb200_combine_crc": (C.c_uint32, [C.c_uint32, _P, C.c_size_t]),
    "bzb200_write_stream_header": (C.c_int, [_P, C.c_int, _P, C.c_size_t]),
    "bzb200_write_stream_trailer": (C.c_int, [_P, _P, C.c_size_t, C.c_uint64, C.c_uint32, C.POINTER(C.c_size_t)]),
    "bzb200_max_output_bytes": (C.c_size_t, [C.c_int, C.c_size_t]),
    "bzb200_compress_device": (C.c_int, [_P, C.c_int, _P, C.c_size_t, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "bzb200_compress_host": (C.c_int, [_P, C.c_int, _P, C.c_size_t, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "bzb200_debug_stage": (C.c_int, [_P, C.c_uint32, C.c_int, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "bzb200_profile": (C.c_int, [_P, C.c_int]),
    "bzb200_profile_count": (C.c_int, [_P]),
    "bzb200_profile_get": (C.c_int, [_P, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_uint64), C.POINTER(C.c_double)]),
    "bzb200_launch_count": (C.c_uint64, [_P]),
    "bzb200_sort_stats": (C.c_int, [_P, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]),
    "bzb200_path_stats": (C.c_int, [_P, C.POINTER(C.c_uint64), C.c_size_t]),
    "bzb200_block_crcs": (C.c_int, [_P, C.c_void_p, C.c_size_t]),
    "bzb200_plan_tile_bytes": (C.c_size_t, []),
    "bzb200_enc_create_multi": (C.c_int, [C.c_int, C.c_int, _P, C.POINTER(_P)]),
    "bzb200_slice_halo_bytes": (C.c_size_t, []),
    "bzb200_slice_begin": (C.c_int, [_P, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, _P, C.c_uint64, C.c_uint64,
                                     C.POINTER(C.c_int64)]),
    "bzb200_slice_counts": (C.c_int, [_P, C.c_int64, C.POINTER(C.c_uint64)]),
    "bzb200_slice_prefix": (C.c_int, [_P, C.c_uint64, C.c_uint64]),
    "bzb200_slice_windows": (C.c_int, [_P, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(_P)]),
    "bzb200_cut_window": (C.c_uint32, []),
    "bzb200_cut_walk": (C.c_int, [_P, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, _P, _P, _P,
                                  C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "bzb200_slice_set_blocks": (C.c_int, [_P, C.c_uint32, _P, _P, C.c_uint32]),
    "bzb200_slice_blocks": (C.c_int, [_P, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]),
    "bzb200_slice_extend": (C.c_int, [_P, C.c_uint64]),
    "bzb200_pool_create": (C.c_int, [C.c_int, _P, C.POINTER(_P)]),
    "bzb200_pool_destroy": (None, [_P]),
    "bzb200_pool_size": (C.c_int, [_P]),
    "bzb200_pool_last_error": (C.c_char_p, [_P]),
    "bzb200_pool_compress_host": (C.c_int, [_P, C.c_int, _P, C.c_size_t, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "bzb200_pool_stats": (C.c_int, [_P, _P, C.c_size_t]),
    "bzb200_version": (C.c_char_p, []),
    "bzb200_decompress_device": (C.c_int, [_P, _P, C.c_size_t, _P, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_int)]),
    "bzb200_decompress_host": (C.c_int, [_P, _P, C.c_size_t, _P, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_int)]),
    "bzb200_dec_stats": (C.c_int, [_P, _P, C.c_size_t]),
    "bzb200_dec_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "bzb200_dec_write": (C.c_int, [_P, _P, C.c_size_t]),
    "bzb200_dec_finish": (C.c_int, [_P]),
    "bzb200_dec_error_kind": (C.c_int, [_P]),
    "bzb200_dec_read": (C.c_size_t, [_P, _P, C.c_size_t]),
    "bzb200_dec_output_size": (C.c_size_t, [_P]),
    "bzb200_dec_reset": (C.c_int, [_P]),
    "bzb200_dec_destroy": (None, [_P]),
    "bzb200_dec_last_error": (C.c_char_p, [_P]),
    "bzb200_decompress": (C.c_int, [C.c_int, _P, C.c_size_t, C.POINTER(_P), C.POINTER(C.c_size_t), C.POINTER(C.c_int)]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python rust-compression_b200/build.py` "
                "(there is no CPU fallback for this path)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
