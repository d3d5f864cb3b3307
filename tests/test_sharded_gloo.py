"""world_size-2 gloo test (CPU) of the block-wise sharding logic in rust-compression_b200/sharded.py: block
ranges, bit-length exchange, payload gather, bit-granular join and trailer CRC fold.  The CUDA context is replaced
by a CPU stand-in that serves each rank's block bit strings from the oracle's stage dump, so the test exercises the
product's orchestration code (not its kernels) and the result must equal the oracle's single-stream output."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackedContext:
    """Same surface as device.Context, CPU tensors, bits taken from the oracle."""

    def __init__(self):
        self.device = torch.device("cpu")

    def plan(self, level, d_in):
        from oracle import orc
        self.level = level
        self.run = orc.Run(d_in.numpy().tobytes(), level)
        self.nblocks = self.run.nblocks
        self.bits = np.unpackbits(np.frombuffer(self.run.out, dtype=np.uint8))
        self.encoded = (0, 0)
        return self.nblocks

    # the plan in four steps: the stand-in checks that sharded.py hands every rank the COMPLETE tile arrays
    def plan_begin(self, level, d_in):
        self._level, self._d_in = level, d_in
        self.nt = (d_in.numel() + 4095) // 4096
        return self.nt

    def plan_heads(self, t0, t1, t_head):
        t_head[t0:t1] = torch.arange(t0, t1, dtype=torch.int64) * 3 + 1

    def plan_counts(self, t_head, t0, t1, t_cnt):
        assert torch.equal(t_head[:self.nt], torch.arange(self.nt, dtype=torch.int64) * 3 + 1), "tile heads incomplete"
        t_cnt[t0:t1] = torch.arange(t0, t1, dtype=torch.int32) + 7

    def plan_finish(self, t_cnt):
        assert torch.equal(t_cnt[:self.nt], torch.arange(self.nt, dtype=torch.int32) + 7), "tile counts incomplete"
        return self.plan(self._level, self._d_in)

    def block_table(self, with_crc=True):
        """Like device.Context.block_table: with_crc=False returns CRCs only for the blocks this rank encoded
        (zeros elsewhere), so the test exercises the CRC exchange between ranks."""
        nb = self.nblocks
        infos = [self.run.info(b) for b in range(nb)]
        in_off = np.array([i["in_start"] for i in infos] + [infos[-1]["in_end"] if nb else 0], dtype=np.uint64)
        rle = np.zeros(nb + 1, dtype=np.uint64)
        crc = np.array([i["crc"] for i in infos], dtype=np.uint32)
        if not with_crc:
            keep = np.zeros(nb, dtype=bool)
            keep[self.encoded[0]:self.encoded[1]] = True
            crc = np.where(keep, crc, 0).astype(np.uint32)
        return in_off, rle, crc

    @staticmethod
    def _or_bits(dst, dst_bit, bits):
        nbytes = (dst_bit + len(bits) + 7) // 8
        cur = np.unpackbits(dst[:nbytes].numpy())
        cur[dst_bit:dst_bit + len(bits)] |= bits
        dst[:nbytes] = torch.from_numpy(np.packbits(cur))

    def encode_blocks(self, b0, b1, d_out, start_bit):
        self.encoded = (b0, b1)
        s = self.run.info(b0)["bit_start"]
        e = self.run.info(b1 - 1)["bit_end"]
        self._or_bits(d_out, start_bit, self.bits[s:e])
        return start_bit + (e - s)

    def bit_append(self, d_dst, dst_bit, d_src, nbits):
        self._or_bits(d_dst, dst_bit, np.unpackbits(d_src.numpy())[:nbits])

    def write_stream_header(self, level, d_out):
        d_out[:4] = torch.tensor([0x42, 0x5A, 0x68, 0x30 + level], dtype=torch.uint8)

    def write_stream_trailer(self, d_out, at_bit, combined_crc):
        v = (0x177245385090 << 32) | combined_crc
        bits = np.array([(v >> (79 - i)) & 1 for i in range(80)], dtype=np.uint8)
        self._or_bits(d_out, at_bit, bits)
        return (at_bit + 80 + 7) // 8

    @staticmethod
    def combine_crc(crcs, seed=0):
        c = seed
        for x in crcs:
            c = (((c << 1) | (c >> 31)) & 0xFFFFFFFF) ^ int(x)
        return c

    def sync(self):
        pass


def _worker(rank, world, port, level, nbytes, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import gen
    from oracle import orc
    from rust_compression_b200 import sharded
    data = gen.g2(7, nbytes) if level == 1 else gen.text(3, nbytes)
    t_in = torch.frombuffer(bytearray(data), dtype=torch.uint8)
    ctx = OracleBackedContext()
    stream, info = sharded.compress_sharded(ctx, level, t_in)
    if rank == 0:
        got = stream.numpy().tobytes()
        q.put((got == orc.compress(data, level), info))
    else:
        assert stream is None
        q.put((True, info))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("level,nbytes", [(1, 700000), (9, 120000)])
def test_sharded_two_ranks_gloo(level, nbytes):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, level, nbytes, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for ok, _ in res)
    infos = sorted((i for _, i in res), key=lambda i: i["rank"])
    assert infos[0]["b0"] == 0 and infos[0]["b1"] == infos[1]["b0"] and infos[1]["b1"] == infos[1]["nblocks"]


def test_block_range_partition():
    sys.path.insert(0, ROOT)
    from rust_compression_b200 import sharded
    for nb in (0, 1, 2, 7, 8, 1194, 10740):
        for w in (1, 2, 4, 8):
            r = [sharded.block_range(nb, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == nb
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
