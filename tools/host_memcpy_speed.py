#!/usr/bin/env python3
"""Host memcpy bandwidth of one thread (what bounds a copying write/read API): 1 GiB in 1 MiB pieces, pageable -> pinned."""
import ctypes, time, torch
n = 1 << 30
src = torch.empty(n, dtype=torch.uint8).pin_memory(); src.fill_(7)
dst = torch.empty(n, dtype=torch.uint8).pin_memory(); dst.fill_(1)
libc = ctypes.CDLL("libc.so.6")
libc.memcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
for rep in range(3):
    t = time.perf_counter()
    for lo in range(0, n, 1 << 20):
        libc.memcpy(dst.data_ptr() + lo, src.data_ptr() + lo, 1 << 20)
    dt = time.perf_counter() - t
    print("memcpy 1 GiB in 1 MiB pieces: %.1f ms  %.1f GB/s" % (dt * 1e3, n / dt / 1e9))
