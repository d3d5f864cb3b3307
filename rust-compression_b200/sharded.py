"""Block-wise sharding of one .bz2 stream across the GPUs of a box (one process per GPU, torch.distributed).

bzip2 blocks are independent once the RLE1 cut points are known, so the path partitions:
  1. plan: every rank evaluates K1's per-tile summaries (last run head, emitted bytes; 12 bytes per 4 KiB tile) for
     its 1/W of the tiles; two small all-gathers make them complete everywhere and every rank runs the cheap global
     part (prefix sum + cut chain), so all ranks hold the same block table;
  2. every rank compresses only its contiguous range of blocks (RLE1 scatter, CRC, K2-K6 for those blocks);
  3. the block CRCs (all-gather) and the compressed bit strings (NCCL P2P send/recv over NVLink) go to rank 0, where
     K7 joins them at bit granularity and the stream trailer with the folded CRC is appended.
No collective touches the data path of a block.
"""
import numpy as np
import torch
import torch.distributed as dist

from .device import Context, max_output_bytes


def block_range(nblocks, rank, world):
    """Contiguous ranges: block b belongs to rank floor(b*world/nblocks)."""
    lo = (nblocks * rank) // world
    hi = (nblocks * (rank + 1)) // world
    return lo, hi


def plan_sharded(ctx, level, d_in, rank, world, group=None):
    """The K1 plan with the per-tile work split over the ranks (see the module docstring). Returns nblocks."""
    if world == 1 or not hasattr(ctx, "plan_begin"):
        return ctx.plan(level, d_in)
    dev = d_in.device
    nt = ctx.plan_begin(level, d_in)
    per = max(1, (nt + world - 1) // world)
    t0, t1 = min(nt, rank * per), min(nt, (rank + 1) * per)
    t_head = torch.full((world * per,), -1, dtype=torch.int64, device=dev)
    ctx.plan_heads(t0, t1, t_head)
    mine = t_head[rank * per:(rank + 1) * per].clone()
    dist.all_gather_into_tensor(t_head, mine, group=group)
    t_cnt = torch.zeros(world * per, dtype=torch.int32, device=dev)
    ctx.plan_counts(t_head, t0, t1, t_cnt)
    mine = t_cnt[rank * per:(rank + 1) * per].clone()
    dist.all_gather_into_tensor(t_cnt, mine, group=group)
    return ctx.plan_finish(t_cnt)


def compress_sharded(ctx, level, d_in, group=None, gather=True):
    """Returns (d_stream or None, info). On rank 0 d_stream holds the complete .bz2 stream (device uint8 tensor).

    d_in: the WHOLE input, resident on this rank's GPU (every rank holds the same bytes).
    """
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    dev = d_in.device
    nb = plan_sharded(ctx, level, d_in, rank, world, group)
    in_off, rle_off, _ = ctx.block_table(with_crc=False)
    b0, b1 = block_range(nb, rank, world)
    my_in = int(in_off[b1] - in_off[b0]) if nb else 0
    cap = (max_output_bytes(level, my_in) + 64 + 3) & ~3
    d_out = torch.zeros(cap, dtype=torch.uint8, device=dev)
    start = 32 if rank == 0 else 0
    if rank == 0:
        ctx.write_stream_header(level, d_out)
    end = ctx.encode_blocks(b0, b1, d_out, start) if b1 > b0 else start
    info = {"nblocks": nb, "b0": b0, "b1": b1, "bits": end - start, "rank": rank, "world": world}
    crc = ctx.block_table(with_crc=False)[2]  # valid for [b0, b1)
    if world == 1:
        total = ctx.write_stream_trailer(d_out, end, ctx.combine_crc(crc))
        return d_out[:total], info
    if not gather:
        return None, info

    # all-gather (nbits, block CRCs of the rank's range), then payloads to rank 0
    per = (nb + world - 1) // world + 1
    meta = np.zeros(1 + per, dtype=np.int64)
    meta[0] = end - start
    meta[1:1 + (b1 - b0)] = crc[b0:b1]
    t_meta = torch.from_numpy(meta).to(dev)
    t_all = torch.empty(world * (1 + per), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(t_all, t_meta, group=group)
    h_all = t_all.cpu().numpy().reshape(world, 1 + per)
    bits = [int(h_all[r, 0]) for r in range(world)]
    crc = np.zeros(nb, dtype=np.uint32)
    for r in range(world):
        lo, hi = block_range(nb, r, world)
        crc[lo:hi] = h_all[r, 1:1 + (hi - lo)].astype(np.uint32)
    if rank == 0:
        total_bits = 32 + sum(bits)
        need = ((total_bits + 80 + 31) // 32) * 4 + 64
        if need > d_out.numel():
            big = torch.zeros(need, dtype=torch.uint8, device=dev)
            big[: (end + 7) // 8] = d_out[: (end + 7) // 8]
            d_out = big
        recv = []
        reqs = []
        for r in range(1, world):
            nbytes = ((bits[r] + 31) // 32) * 4
            buf = torch.empty(max(nbytes, 4), dtype=torch.uint8, device=dev)
            recv.append(buf)
            if nbytes:
                reqs.append(dist.irecv(buf, src=dist.get_global_rank(group, r) if group else r, group=group))
        for q in reqs:
            q.wait()
        cur = end
        for r in range(1, world):
            if bits[r]:
                ctx.bit_append(d_out, cur, recv[r - 1], bits[r])
                cur += bits[r]
        total = ctx.write_stream_trailer(d_out, cur, ctx.combine_crc(crc))
        ctx.sync()
        return d_out[:total], info
    else:
        nbytes = ((bits[rank] + 31) // 32) * 4
        if nbytes:
            dist.send(d_out[:nbytes], dst=dist.get_global_rank(group, 0) if group else 0, group=group)
        return None, info


def compress_host_sharded(ctx: Context, level, h_slice, h_out=None, group=None):
    """End-to-end call with HOST buffers: every rank passes its own slice of the input (pinned host uint8 tensor);
    slices are copied H2D, all-gathered over NCCL/NVLink into the whole input on every GPU, compressed block-wise
    (compress_sharded) and the finished stream is copied D2H on rank 0 (into h_out if given).
    Returns (host tensor view holding the stream or None, info)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    dev = ctx.device
    d_slice = h_slice.to(dev, non_blocking=True)
    if world > 1:
        d_full = torch.empty(world * d_slice.numel(), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(d_full, d_slice, group=group)
    else:
        d_full = d_slice
    d_stream, info = compress_sharded(ctx, level, d_full, group=group)
    if d_stream is None:
        return None, info
    n = d_stream.numel()
    if h_out is None:
        h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_out[:n].copy_(d_stream, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    info["h2d_bytes"] = int(h_slice.numel())
    info["d2h_bytes"] = int(n)
    return h_out[:n], info
