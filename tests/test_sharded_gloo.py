"""world_size 2 and 3 gloo tests (CPU) of rust-compression_b200/sharded.py — the one-process-per-GPU sharding of one
.bz2 stream: slice bounds, halo and tail exchange (batched P2P), the three all-gathers of the sliced K1 plan, the host
cut walk, block ownership, the (bits, first byte, CRC) exchange, the whole-byte ownership join and the trailer.

The CUDA context is replaced by a CPU stand-in with the same surface that computes every per-slice quantity FROM THE
RANK'S RESIDENT BYTES ONLY (numpy restatement of the K1 definitions; block bit strings from the oracle run on the block's
resident input) — so a wrong carry, emitted offset, halo, tail range or window row produces a different stream.  The
result on rank 0 must equal the oracle's one-pass stream bit for bit.  (The CUDA kernels behind the same calls are
checked on the GPU by tests/test_gpu_sharded.py.)"""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TILE = 4096


class CpuSliceContext:
    """device.Context's sliced-plan surface on CPU tensors."""

    def __init__(self):
        self.device = torch.device("cpu")

    def sync(self):
        pass

    # ---- plan
    def slice_begin(self, level, n, lo, hi, d_buf, off, avail, reserve):
        self.level, self.n, self.lo, self.hi, self.buf, self.off, self.avail = level, n, lo, hi, d_buf, off, avail
        assert lo % TILE == 0 and (hi % TILE == 0 or hi == n) and off % 16 == 0
        self._model(-1, 0)
        heads = np.nonzero(self.head[:hi - lo])[0]
        return int(lo + heads[-1]) if heads.size else -1

    def _bytes(self, a, b):
        return self.buf[self.off + (a - self.lo):self.off + (b - self.lo)].numpy()

    def _model(self, carry_in, e_lo):
        """K1 definitions (k1_rle.cu header) over the usable resident bytes [lo, tn*TILE)."""
        lo, n = self.lo, self.n
        tn = (n + TILE - 1) // TILE if self.avail >= n else (self.avail - 1) // TILE
        end = min(n, tn * TILE)
        a = self._bytes(lo, end).astype(np.int16)
        prev = int(self._bytes(lo - 1, lo)[0]) if lo > 0 else -1
        nxt = int(self._bytes(end, end + 1)[0]) if end < n else -1
        m = a.size
        head = np.empty(m, dtype=bool)
        head[0] = (lo == 0) or a[0] != prev
        head[1:] = a[1:] != a[:-1]
        idx = np.arange(m, dtype=np.int64) + lo
        s = np.maximum.accumulate(np.where(head, idx, -1))
        s = np.maximum(s, carry_in)
        q = (idx - s) % 255
        tail = np.empty(m, dtype=bool)
        tail[:-1] = a[1:] != a[:-1]
        tail[-1] = (end == n) or a[-1] != nxt
        pe = tail | (q == 254)
        emit = (q < 4).astype(np.int64) + (pe & (q >= 3)).astype(np.int64)
        self.head, self.pe, self.E, self.end = head, pe, e_lo + np.cumsum(emit), end
        self.emit = emit

    def slice_counts(self, carry_in):
        self.carry_in = carry_in
        self._model(carry_in, 0)
        return int(self.emit[:self.hi - self.lo].sum())

    def slice_prefix(self, e_lo, e_tot):
        self.e_lo, self.e_tot = e_lo, e_tot
        self._model(self.carry_in, e_lo)
        self.e_hi = int(self.E[self.hi - self.lo - 1])

    def slice_windows(self, x0, out_rows):
        T, W = self.level * 100000 - 19, out_rows.shape[1]
        lo_j = (self.e_lo - x0) // T if self.e_lo >= x0 else 0
        cnt = 0
        if self.e_hi >= x0 + T:
            hi_j = (self.e_hi - x0) // T - 1
            if hi_j >= lo_j:
                cnt = hi_j - lo_j + 1
        ends = np.nonzero(self.pe)[0]
        Ee = self.E[ends]
        for k in range(cnt):
            c = x0 + (lo_j + k + 1) * T
            xs = c + np.arange(W)
            kk = np.searchsorted(Ee, xs, side="left")
            ok = (kk < ends.size) & (xs <= self.e_tot)
            assert ok[xs <= min(self.e_tot, c + W - 1)].all(), "the halo does not cover the window"
            kk = np.minimum(kk, ends.size - 1)
            i = ends[kk] + self.lo
            v = (Ee[kk] - c).astype(np.uint64) | ((i + 1).astype(np.uint64) << np.uint64(16))
            v |= np.where(i == self.n - 1, np.uint64(1) << np.uint64(63), np.uint64(0))
            out_rows[k] = torch.from_numpy(np.where(ok, v, np.uint64(0)).view(np.int64))
        return lo_j, cnt

    def slice_set_blocks(self, in_off, rle_off, ml):
        self.in_off, self.rle_off = np.asarray(in_off), np.asarray(rle_off)
        self.nblocks = self.in_off.size - 1
        self.crc = np.zeros(self.nblocks, dtype=np.uint32)

    def slice_extend(self, avail):
        assert avail > self.avail
        self.avail = avail

    # ---- encode: every block from the rank's RESIDENT bytes
    def encode_blocks(self, b0, b1, d_out, start_bit):
        from oracle import orc, verify
        L = verify._lib()
        bits = np.unpackbits(d_out.numpy())
        pos = start_bit
        for b in range(b0, b1):
            lo, hi = int(self.in_off[b]), int(self.in_off[b + 1])
            assert lo >= self.lo and hi <= self.avail, "block input is not resident on this rank"
            data = np.ascontiguousarray(self._bytes(lo, hi))
            out = np.empty(self.level * 125000 + 8192, dtype=np.uint8)
            info = np.zeros(4, dtype=np.uint64)
            r = L.orc_encode_block(self.level, data.ctypes.data, data.size, out.ctypes.data, out.size, info.ctypes.data)
            assert r > 0 and info[0] == 1 and int(info[2]) == int(self.rle_off[b + 1] - self.rle_off[b])
            nb = int(info[3])
            bits[pos:pos + nb] = np.unpackbits(out[:r])[:nb]
            pos += nb
            self.crc[b] = int(info[1])
        d_out.copy_(torch.from_numpy(np.packbits(bits)))
        return pos

    def block_table(self, with_crc=True):
        assert not with_crc
        return self.in_off, self.rle_off, self.crc

    @staticmethod
    def combine_crc(crcs, seed=0):
        c = seed
        for x in crcs:
            c = (((c << 1) | (c >> 31)) & 0xFFFFFFFF) ^ int(x)
        return c

    def write_stream_header(self, level, d_out):
        d_out[:4] = torch.tensor(list(b"BZh" + bytes([0x30 + level])), dtype=torch.uint8)

    def bit_append(self, d_dst, dst_bit, d_src, nbits):
        dst = np.unpackbits(d_dst.numpy())
        dst[dst_bit:dst_bit + nbits] |= np.unpackbits(d_src.numpy())[:nbits]
        d_dst.copy_(torch.from_numpy(np.packbits(dst)))

    def write_stream_trailer(self, d_out, at_bit, combined):
        bits = np.unpackbits(d_out.numpy())
        val = (0x177245385090 << 32) | combined
        bits[at_bit:at_bit + 80] |= np.array([(val >> (79 - i)) & 1 for i in range(80)], dtype=np.uint8)
        d_out.copy_(torch.from_numpy(np.packbits(bits)))
        return (at_bit + 80 + 7) // 8


def _worker(rank, world, port, case, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import gen
        import rust_compression_b200  # noqa: F401
        from rust_compression_b200 import sharded
        name, level = case
        data = _make(gen, name)
        sharded.MIN_SLICE = 8192  # small inputs must still be sliced over every rank in this test
        sh = sharded.Shard(level, len(data), rank, world, torch.device("cpu"))
        sh.slice_view().copy_(torch.frombuffer(bytearray(data[sh.lo:sh.hi]), dtype=torch.uint8))
        ctx = CpuSliceContext()
        stream, info = sharded.compress_sharded(ctx, sh)
        out = stream.numpy().tobytes() if stream is not None else None
        q.put((rank, out, info))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _make(gen, name):
    if name == "mixed":
        return gen.mixed(4, 900_000)
    if name == "runs":   # long runs across slice boundaries (carried run heads, 255-piece anchoring), block tails > halo
        return b"x" * 700_000 + gen.g2(3, 400_000) + b"y" * 3_000_000 + gen.text(2, 300_000)
    if name == "aaaab":  # the drift of the cut positions leaves the window: several plan phases
        return b"aaaab" * 2_400_000
    if name == "small":  # fewer blocks than ranks
        return gen.text(9, 150_000)
    raise KeyError(name)


@pytest.mark.parametrize("world,name,level", [(2, "mixed", 1), (3, "mixed", 1), (3, "runs", 1), (2, "aaaab", 1),
                                              (3, "small", 2)])
def test_sharded_stream_equals_oracle(world, name, level):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import gen
    from oracle import orc
    want = orc.compress(_make(gen, name), level)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() * 7 + world * 13 + len(name)) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, (name, level), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        rank, out, info = q.get(timeout=600)
        res[rank] = (out, info)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res[0][0] == want, f"joined stream differs from the oracle ({name}, world {world})"
    assert all(res[r][0] is None for r in range(1, world))
    nb = res[0][1]["nblocks"]
    covered = sorted((res[r][1]["b0"], res[r][1]["b1"]) for r in range(world))
    assert covered[0][0] == 0 and covered[-1][1] == nb and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    if name == "aaaab":
        assert res[0][1]["plan_phases"] > 1
