"""Block-wise sharding of one .bz2 stream across the GPUs of a box — one process per GPU (torch.distributed).

bzip2 blocks are independent once the RLE1 cut points are known (the reference cuts them in one sequential pass,
src/bzip2/encoder.rs:671-716, cut test :692-696).  Every rank keeps ONLY ITS SLICE of the input in HBM:

  1. plan  (include/bzb200.h section 2b)  per-slice K1 work + three tiny all-gathers: the last run head of every slice
           (1 word per rank), the bytes every slice emits (1 word per rank), the cut-window rows (2 KB per block);
           every rank then walks the cut chain on the host (bzb200_cut_walk) and holds the same block table.  No
           O(input) array is exchanged and nobody scans more than its own tiles.
  2. tail  a rank encodes the blocks that START in its slice; the last of them ends in the next rank's bytes (usually
           < 1 MB): those bytes come over NVLink with one batched P2P exchange.
  3. encode  K2-K6 for the rank's blocks, from bit 0 of its own buffer.
  4. join  one all-gather of (bit count, first byte, block CRCs); every rank shifts its bit string to the bit phase it
           has in the joined stream on its own GPU (K7) and sends the WHOLE BYTES it owns — the byte two neighbours share
           belongs to the earlier one, which ORs in the later one's leading bits — so rank 0 receives every payload
           straight at its final byte offset (one batched NCCL P2P exchange) and only appends the trailer.
No collective touches a block's data path.
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

from . import device as dv
from .device import max_output_bytes

TRACE = bool(os.environ.get("BZB200_SHARD_TRACE"))   # per-phase wall times of compress_sharded on stderr (adds syncs)


class _Phases:
    """Wall time per phase of one compress_sharded call (only when BZB200_SHARD_TRACE is set: each mark synchronises)."""

    def __init__(self, rank):
        self.rank, self.t, self.rows = rank, time.perf_counter(), []

    def mark(self, name):
        if not TRACE:
            return
        torch.cuda.synchronize()
        now = time.perf_counter()
        self.rows.append((name, 1e3 * (now - self.t)))
        self.t = now

    def done(self):
        if TRACE:
            sys.stderr.write("[rank %d] " % self.rank + "  ".join("%s %.2f" % r for r in self.rows) + "\n")


LEFT = 256             # bytes in front of the slice inside a rank's buffer (16 of them hold the left neighbour's bytes)
MIN_SLICE = 1 << 20    # slices shorter than this are not worth a rank: the leading ranks take the whole (small) input


def slice_bounds(n, world):
    """[(lo, hi)] per rank: tile-aligned equal slices (trailing ranks may be empty for a small input)."""
    tile = dv.plan_tile_bytes()
    per = (n + world - 1) // world
    per = max(per, MIN_SLICE)
    per = (per + tile - 1) // tile * tile
    return [(min(n, per * r), min(n, per * (r + 1))) for r in range(world)]


def tail_reserve(level):
    """Input bytes one block can span beyond the slice end (runs: 255 input bytes per 5 emitted) plus the halo."""
    return dv.slice_halo_bytes() + level * 100000 * 51 + 4 * dv.plan_tile_bytes()


class Shard:
    """A rank's part of the stream: a buffer with the slice at offset LEFT and room behind it for the halo / tail."""

    def __init__(self, level, n_total, rank, world, device):
        self.level, self.n, self.rank, self.world = level, n_total, rank, world
        self.bounds = slice_bounds(n_total, world)
        self.lo, self.hi = self.bounds[rank]
        self.reserve_hi = min(n_total, self.hi + tail_reserve(level)) if self.hi > self.lo else self.lo
        self.buf = torch.zeros(LEFT + max(0, self.reserve_hi - self.lo) + 64, dtype=torch.uint8, device=device)
        self.avail = self.hi  # how far the input is resident

    def slice_view(self):
        """Where the caller puts input bytes [lo, hi)."""
        return self.buf[LEFT:LEFT + (self.hi - self.lo)]

    def owner_ranges(self, g_lo, g_hi):
        """[(rank, lo, hi)] pieces of the global byte range held by each rank's slice."""
        out = []
        for r, (lo, hi) in enumerate(self.bounds):
            a, b = max(lo, g_lo), min(hi, g_hi)
            if b > a:
                out.append((r, a, b))
        return out


def _fetch(shard, wants, group):
    """wants[r] = list of global byte ranges rank r needs from other ranks' slices.  One batched P2P exchange."""
    ops = []
    me = shard.rank
    for dst in range(shard.world):
        for (g_lo, g_hi) in wants[dst]:
            for (src, a, b) in shard.owner_ranges(g_lo, g_hi):
                if src == dst:
                    continue
                if me == src:
                    t = shard.buf[LEFT + (a - shard.lo):LEFT + (b - shard.lo)]
                    ops.append(dist.P2POp(dist.isend, t, dist.get_global_rank(group, dst) if group else dst, group))
                elif me == dst:
                    off = LEFT + a - shard.lo  # may be < LEFT: the bytes in front of the slice
                    t = shard.buf[off:off + (b - a)]
                    ops.append(dist.P2POp(dist.irecv, t, dist.get_global_rank(group, src) if group else src, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def _all_gather_i64(vals, device, group):
    t = torch.tensor(vals, dtype=torch.int64, device=device)
    world = dist.get_world_size(group)
    out = torch.empty(world * t.numel(), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, t, group=group)
    return out.cpu().numpy().reshape(world, -1)


def plan_sharded(ctx, shard, group=None):
    """The K1 plan over the ranks' slices.  Returns (in_off, rle_off, max_block_len) — identical on every rank."""
    level, n, rank, world = shard.level, shard.n, shard.rank, shard.world
    dev = shard.buf.device
    halo = dv.slice_halo_bytes()
    have = shard.hi > shard.lo
    # halo: 16 bytes in front of the slice, up to `halo` bytes behind it
    wants = []
    for (lo, hi) in shard.bounds:
        w = []
        if hi > lo:
            if lo > 0:
                w.append((lo - 16, lo))
            if hi < n:
                w.append((hi, min(n, hi + halo)))
        wants.append(w)
    _fetch(shard, wants, group)
    if have:
        shard.avail = min(n, shard.hi + halo)
    lh = ctx.slice_begin(level, n, shard.lo, shard.hi, shard.buf, LEFT, shard.avail, shard.reserve_hi) if have else -1
    heads = _all_gather_i64([lh], dev, group)[:, 0]
    em = ctx.slice_counts(int(max([-1] + [int(h) for h in heads[:rank]]))) if have else 0
    emitted = [int(x) for x in _all_gather_i64([em], dev, group)[:, 0]]
    e_lo = [sum(emitted[:r]) for r in range(world)]
    e_tot = sum(emitted)
    if have:
        ctx.slice_prefix(e_lo[rank], e_tot)
    T = level * 100000 - 19
    W = dv.cut_window()
    max_blocks = (n + n // 4 + 64) // T + 2
    state = np.zeros(4, dtype=np.uint64)
    in_off = np.zeros(max_blocks + 1, dtype=np.uint64)
    rle_off = np.zeros(max_blocks + 1, dtype=np.uint64)
    nb = ml = 0
    phases = 0
    while not state[2]:
        x0 = int(state[1])
        K = (e_tot - x0) // T if e_tot >= x0 + T else 0
        # rows every slice tabulates (centres x0 + (j+1) T inside (E_lo, E_hi]): known to everybody
        spans = []
        for r in range(world):
            lo_j = (e_lo[r] - x0) // T if e_lo[r] >= x0 else 0
            cnt = 0
            e_hi = e_lo[r] + emitted[r]
            if emitted[r] and e_hi >= x0 + T:
                hi_j = (e_hi - x0) // T - 1
                if hi_j >= lo_j:
                    cnt = hi_j - lo_j + 1
            spans.append((lo_j, cnt))
        rows_max = max(1, max(c for _, c in spans))
        mine = torch.zeros((rows_max, W), dtype=torch.int64, device=dev)
        if have and spans[rank][1]:
            j0, nj = ctx.slice_windows(x0, mine)
            assert (j0, nj) == spans[rank], ((j0, nj), spans[rank])
        allrows = torch.empty((world * rows_max, W), dtype=torch.int64, device=dev)
        if hasattr(ctx, "stream") and mine.is_cuda:
            torch.cuda.current_stream(dev).wait_stream(ctx.stream)
        dist.all_gather_into_tensor(allrows, mine, group=group)
        h = allrows.cpu().numpy().view(np.uint64).reshape(world, rows_max, W)
        F = np.zeros((K, W), dtype=np.uint64)
        for r, (lo_j, cnt) in enumerate(spans):
            if cnt:
                F[lo_j:lo_j + cnt] = h[r, :cnt]
        nb, ml = dv.cut_walk(F, T, e_tot, n, state, in_off, rle_off)
        phases += 1
        if phases > (1 << 20):
            raise RuntimeError("cut chain does not terminate")
    in_off = in_off[:nb + 1].copy()
    rle_off = rle_off[:nb + 1].copy()
    if have:
        ctx.slice_set_blocks(in_off, rle_off, ml)
    shard.plan_phases = phases
    return in_off, rle_off, ml


def block_ranges(shard, in_off):
    """[(b0, b1)] per rank: the blocks that start inside the rank's slice."""
    starts = in_off[:-1]
    out = []
    for (lo, hi) in shard.bounds:
        b0 = int(np.searchsorted(starts, lo, side="left"))
        b1 = int(np.searchsorted(starts, hi, side="left"))
        out.append((b0, b1) if hi > lo else (b0, b0))
    return out


def compress_sharded(ctx, shard, group=None):
    """shard.slice_view() holds this rank's input bytes.  Returns (d_stream or None, info): on rank 0 d_stream is the
    complete .bz2 stream (device uint8 tensor)."""
    level, n, rank, world = shard.level, shard.n, shard.rank, shard.world
    dev = shard.buf.device
    tile = dv.plan_tile_bytes()
    shard.avail = shard.hi
    ph = _Phases(rank)
    in_off, rle_off, _ = plan_sharded(ctx, shard, group)
    ph.mark("plan")
    nb = in_off.size - 1
    ranges = block_ranges(shard, in_off)
    # tails: the input of a rank's last block beyond what is resident
    wants, needs = [], []
    halo = dv.slice_halo_bytes()
    for r, (b0, b1) in enumerate(ranges):
        lo, hi = shard.bounds[r]
        need = 0
        w = []
        if b1 > b0:
            end = int(in_off[b1])
            need = min(n, (end + tile - 1) // tile * tile + (1 if end < n else 0))
            avail = min(n, hi + halo)
            if need > avail:
                w.append((avail, need))
        wants.append(w)
        needs.append(need)
    _fetch(shard, wants, group)
    b0, b1 = ranges[rank]
    if b1 > b0 and needs[rank] > shard.avail:
        ctx.slice_extend(needs[rank])
        shard.avail = needs[rank]
    ph.mark("tail")
    # encode
    my_in = int(in_off[b1] - in_off[b0]) if b1 > b0 else 0
    cap = (max_output_bytes(level, my_in) + 64 + 3) & ~3
    d_out = torch.zeros(cap, dtype=torch.uint8, device=dev)
    bits = 0
    if rank == 0:
        ctx.write_stream_header(level, d_out)
        bits = 32
    ph.mark("alloc")
    if b1 > b0:
        bits = ctx.encode_blocks(b0, b1, d_out, bits)
    ph.mark("encode")
    info = {"nblocks": nb, "b0": b0, "b1": b1, "bits": bits, "rank": rank, "world": world,
            "plan_phases": getattr(shard, "plan_phases", 0)}
    crc_mine = ctx.block_table(with_crc=False)[2][b0:b1] if b1 > b0 else np.zeros(0, dtype=np.uint32)
    if world == 1:
        total = ctx.write_stream_trailer(d_out, bits, ctx.combine_crc(crc_mine))
        ctx.sync()
        return d_out[:total], info
    # join: (bits, first byte, CRCs) of every rank
    per_max = max(1, max(e - s for s, e in ranges))
    first = int(d_out[0].item()) if bits else 0
    meta = [bits, first] + [int(c) for c in crc_mine] + [0] * (per_max - len(crc_mine))
    allm = _all_gather_i64(meta, dev, group)
    ph.mark("meta")
    all_bits = [int(allm[r, 0]) for r in range(world)]
    crc = np.zeros(nb, dtype=np.uint32)
    for r, (s, e) in enumerate(ranges):
        crc[s:e] = allm[r, 2:2 + (e - s)].astype(np.uint32)
    P = [sum(all_bits[:r]) for r in range(world)]
    total_bits = sum(all_bits)
    live = [r for r in range(world) if all_bits[r]]

    def owned(r):  # whole bytes of the joined stream rank r owns
        return (P[r] + 7) // 8, (P[r] + all_bits[r] + 7) // 8

    src = None
    if bits:
        bph = P[rank] & 7
        if bph:  # K7 on this GPU: the bit string at the bit phase it has in the joined stream
            sh = torch.zeros(((bph + bits + 31) // 32) * 4 + 64, dtype=torch.uint8, device=dev)
            ctx.bit_append(sh, bph, d_out, bits)
            src = sh[1:(bph + bits + 7) // 8]  # the first byte belongs to the rank in front
        else:
            src = d_out[:(bits + 7) // 8]
        nxt = [r for r in live if r > rank]
        e = (P[rank] + bits) & 7
        if nxt and e:  # the later rank's leading bits go into the byte both share
            fb = int(allm[nxt[0], 1]) >> e
            if fb:
                ctx.sync()
                src[-1:] |= torch.tensor([fb], dtype=torch.uint8, device=dev)
        ctx.sync()
    ph.mark("shift")
    if rank == 0:
        need = ((total_bits + 80 + 31) // 32) * 4 + 64
        d_final = torch.zeros(need, dtype=torch.uint8, device=dev)
        ops = []
        for r in live:
            a, b = owned(r)
            if b <= a:
                continue
            if r == 0:
                d_final[a:b] = src[:b - a]
            else:
                ops.append(dist.P2POp(dist.irecv, d_final[a:b], dist.get_global_rank(group, r) if group else r, group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        ph.mark("gather")
        total = ctx.write_stream_trailer(d_final, total_bits, ctx.combine_crc(crc))
        ctx.sync()
        ph.mark("trailer")
        ph.done()
        return d_final[:total], info
    a, b = owned(rank)
    if bits and b > a:
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, src[:b - a].contiguous(),
                                                    dist.get_global_rank(group, 0) if group else 0, group)]):
            w.wait()
    ph.mark("send")
    ph.done()
    return None, info


def compress_host_sharded(ctx, shard, h_slice, h_out=None, group=None):
    """End to end with HOST buffers, one process per GPU: every rank copies its own slice (pinned host uint8 tensor)
    H2D over its own PCIe link, the ranks compress (compress_sharded) and rank 0 copies the stream D2H (into h_out if
    given).  Returns (host tensor view or None, info)."""
    dev = shard.buf.device
    shard.slice_view().copy_(h_slice, non_blocking=True)
    d_stream, info = compress_sharded(ctx, shard, group=group)
    if d_stream is None:
        return None, info
    nbytes = d_stream.numel()
    if h_out is None:
        h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_out[:nbytes].copy_(d_stream, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    info["h2d_bytes"] = int(h_slice.numel())
    info["d2h_bytes"] = int(nbytes)
    return h_out[:nbytes], info
