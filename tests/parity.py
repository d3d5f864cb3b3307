"""Shared helpers for the GPU parity tests: run the CUDA path through the C ABI, run the oracle, compare stage
by stage so a failure names the first stage that diverges."""
import bz2

import numpy as np


def gpu_run(data, level, batch_elems=None):
    """Returns (stream_bytes, ctx, nblocks). Uses the device-resident job API so that stage dumps stay available."""
    import torch
    from rust_compression_b200 import device as dv

    ctx = dv.Context()
    a = np.frombuffer(bytes(data), dtype=np.uint8)
    d_in = torch.from_numpy(a.copy()).cuda() if a.size else torch.zeros(0, dtype=torch.uint8, device="cuda")
    cap = dv.max_output_bytes(level, a.size)
    d_out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    n = ctx.compress_device(level, d_in, d_out)
    out = d_out[:n].cpu().numpy().tobytes()
    ctx.nblocks_seen = len(ctx.block_table()[0]) - 1 if True else 0
    return out, ctx


STAGES = ["blocks", "crc", "rle", "inuse", "order", "last", "origptr", "mtf", "freq", "len0", "len1", "len2", "len3",
          "len4", "sel", "bits", "stream", "libbz2", "refdec"]


def compare(data, level, orc, keep_sa=True, check_blocks=None):
    """Runs both sides; returns dict stage -> None (ok) or a message describing the first mismatch."""
    from oracle import orc as O

    res = {s: None for s in STAGES}
    ref = O.Run(data, level, keep_sa=keep_sa)
    out, ctx = gpu_run(data, level)
    try:
        in_off, rle_off, crc = ctx.block_table()
        nb = len(in_off) - 1
        if nb != ref.nblocks:
            res["blocks"] = f"nblocks gpu={nb} oracle={ref.nblocks}"
        blocks = range(min(nb, ref.nblocks)) if check_blocks is None else check_blocks

        def first_diff(a, b):
            a = np.asarray(a)
            b = np.asarray(b)
            if a.shape != b.shape:
                return f"shape gpu={a.shape} oracle={b.shape}"
            d = np.nonzero(a != b)[0]
            if d.size:
                i = int(d[0])
                return f"first diff at {i}: gpu={a[i]} oracle={b[i]} ({d.size} diffs)"
            return None

        for b in blocks:
            ri = ref.info(b)

            def setres(stage, msg):
                if msg and res[stage] is None:
                    res[stage] = f"block {b}: {msg}"

            if (int(in_off[b]), int(in_off[b + 1])) != (ri["in_start"], ri["in_end"]) or \
                    int(rle_off[b + 1] - rle_off[b]) != ri["nblock"]:
                setres("blocks", f"range gpu=({in_off[b]},{in_off[b+1]},n={rle_off[b+1]-rle_off[b]}) "
                       f"oracle=({ri['in_start']},{ri['in_end']},n={ri['nblock']})")
                continue
            if int(crc[b]) != ri["crc"]:
                setres("crc", f"gpu={int(crc[b]):08x} oracle={ri['crc']:08x}")
            try:
                gi = ctx.debug_stage(b, "info")
            except Exception as e:  # block not in the last batch
                continue
            setres("rle", first_diff(ctx.debug_stage(b, "rle"), ref.field(b, "rle")))
            setres("inuse", first_diff(gi["in_use"], ref.inuse(b)))
            if keep_sa:
                rank = ctx.debug_stage(b, "rank") & 0xFFFFF
                sa = ref.field(b, "sa")
                want_rank = np.empty_like(sa)
                want_rank[sa] = np.arange(sa.size, dtype=sa.dtype)
                setres("order", first_diff(rank, want_rank))
            setres("last", first_diff(ctx.debug_stage(b, "last"), ref.field(b, "last")))
            if gi["orig_ptr"] != ri["orig_ptr"]:
                setres("origptr", f"gpu={gi['orig_ptr']} oracle={ri['orig_ptr']} (rounds={gi['sort_rounds']}, "
                       f"periodic={gi['periodic']})")
            setres("mtf", first_diff(ctx.debug_stage(b, "mtf"), ref.field(b, "mtf")))
            setres("freq", first_diff(ctx.debug_stage(b, "freq"), ref.field(b, "freq")))
            for k in range(5):
                setres(f"len{k}", first_diff(ctx.debug_stage(b, f"len{k}"), ref.field(b, f"len{k}")))
            setres("sel", first_diff(ctx.debug_stage(b, "sel"), ref.field(b, "sel4")))
            if (gi["bit_start"], gi["bit_end"]) != (ri["bit_start"], ri["bit_end"]):
                setres("bits", f"gpu=({gi['bit_start']},{gi['bit_end']}) oracle=({ri['bit_start']},{ri['bit_end']})")
            if gi["dev_error"]:
                setres("len4", f"device error flag {gi['dev_error']}")
        if out != ref.out:
            a = np.frombuffer(out, dtype=np.uint8)
            r = np.frombuffer(ref.out, dtype=np.uint8)
            m = min(a.size, r.size)
            d = np.nonzero(a[:m] != r[:m])[0]
            res["stream"] = f"len gpu={a.size} oracle={r.size}, first diff byte {int(d[0]) if d.size else m}"
        try:
            if bz2.decompress(out) != bytes(data):
                res["libbz2"] = "decodes to different bytes"
        except Exception as e:
            res["libbz2"] = f"libbz2 rejects the stream: {e}"
        # ... and through the restated reference BZip2Decoder with its own acceptance limits (origPtr bound,
        # tt.len() < 100000*level; oracle/bz2_decoder_oracle.cpp)
        try:
            if O.decode(out) != bytes(data):
                res["refdec"] = "reference decoder restatement decodes to different bytes"
        except O.DecodeError as e:
            res["refdec"] = f"reference decoder restatement rejects the stream: {e.kind}"
    finally:
        ref.close()
        ctx.close()
    return res


def assert_parity(data, level, keep_sa=True):
    from oracle import orc
    res = compare(data, level, orc, keep_sa=keep_sa)
    bad = {k: v for k, v in res.items() if v}
    assert not bad, f"first diverging stage: {next(iter(bad))}: {bad}"
