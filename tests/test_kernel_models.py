"""CPU restatements of the arithmetic a few kernels rest on (no GPU, no library): the identities are small enough to
check exhaustively or on many random cases, and a slip in one of them shows up on the GPU only as a wrong bit somewhere
in a 300 MB stream.

  * k3_mtf.cu  mtf_lane_access: zero-byte flags of (word ^ cccc) and the masked funnel shift of the 32-entry register
               list, including the 0xFF filler behind the in-use bytes;
  * k1_rle.cu  k5_crc_blocks: slicing-by-4 tables of CRC-32/BZIP2 (crc32.rs:82-84);
  * k2_bwt.cu  k2_os_scatter_pf: ticket g -> (block g mod nb, tile g div nb) keeps every tile behind the tiles in front
               of it in its block; k2_local_sort_rx: the greedy window packing never exceeds the window capacity and
               takes every list entry exactly once.
"""
import random

M32 = 0xFFFFFFFF


# ------------------------------------------------------------------ K3: the register part of the MTF list
def _flags(word, c):
    z = word ^ (c * 0x01010101)
    return ((z - 0x01010101) & ~z & 0x80808080) & M32


def _shr_sat(v, s):  # PTX shr.b32: amounts >= 32 give 0
    s = max(s, 0)
    return 0 if s >= 32 else (v >> s)


def mtf_front_model(f, c):
    """The kernel's register-list update on eight 32-bit words (entry j = byte j&3 of word j>>2).
    Returns (new words, position or 255, carry out)."""
    H = wsel = 0
    for k in range(7, -1, -1):
        h = _flags(f[k], c)
        if h:
            H, wsel = h, k
    if H:
        low = (H & -H).bit_length() - 1
        pos = 4 * wsel + (low >> 3)
    else:
        pos = 255
    A = 32 - 8 * (pos + 1)
    prev = (c << 24) & M32
    carry_out = f[7] >> 24
    out = []
    for k in range(8):
        shifted = ((f[k] << 8) | (prev >> 24)) & M32
        mask = _shr_sat(M32, min(A + 32 * k, 32))
        out.append((shifted & mask) | (f[k] & ~mask & M32))
        prev = f[k]
    return out, pos, carry_out


def _pack(entries):
    return [entries[4 * k] | entries[4 * k + 1] << 8 | entries[4 * k + 2] << 16 | entries[4 * k + 3] << 24 for k in range(8)]


def _unpack(words):
    return [(w >> (8 * i)) & 255 for w in words for i in range(4)]


def test_zero_byte_flags_lowest_flag_is_exact():
    rng = random.Random(1)
    for _ in range(20000):
        w, c = rng.getrandbits(32), rng.randrange(256)
        h = _flags(w, c)
        bytes_ = [(w >> (8 * i)) & 255 for i in range(4)]
        assert (h != 0) == (c in bytes_)          # non-zero exactly when the word holds c
        if h:
            assert ((h & -h).bit_length() - 1) >> 3 == bytes_.index(c)   # and its lowest flag is the first hit


def test_mtf_register_list_matches_move_to_front():
    rng = random.Random(2)
    for case in range(4000):
        alpha = rng.choice([2, 5, 17, 31, 32])
        inuse = rng.sample(range(256), alpha) if case % 3 else rng.sample(range(255), alpha - 1) + [255]
        lst = inuse + [255] * (32 - alpha)        # filler behind the in-use bytes (a real 0xFF sits in front of it)
        for _ in range(12):
            c = rng.choice(inuse)
            if lst[0] == c:
                continue                           # the kernel handles position 0 as a zero run
            want_pos = lst.index(c)
            want = [c] + lst[:want_pos] + lst[want_pos + 1:]
            got, pos, _ = mtf_front_model(_pack(lst), c)
            assert pos == want_pos and _unpack(got) == want
            lst = want


def test_mtf_register_list_miss_shifts_everything():
    rng = random.Random(3)
    for _ in range(500):
        lst = rng.sample(range(1, 200), 32)
        c = 250
        got, pos, carry = mtf_front_model(_pack(lst), c)
        assert pos == 255 and carry == lst[31] and _unpack(got) == [c] + lst[:31]


# ------------------------------------------------------------------ K5: slicing-by-4
def test_crc_slicing_by_4_equals_bytewise():
    poly = 0x04C11DB7
    t0 = []
    for i in range(256):
        v = i << 24
        for _ in range(8):
            v = ((v << 1) ^ poly) & M32 if v & 0x80000000 else (v << 1) & M32
        t0.append(v)
    tabs = [t0]
    for k in range(1, 4):
        tabs.append([((tabs[k - 1][i] << 8) & M32) ^ t0[tabs[k - 1][i] >> 24] for i in range(256)])
    rng = random.Random(4)
    for _ in range(300):
        data = [rng.randrange(256) for _ in range(4 * rng.randrange(1, 30))]
        r0 = rng.getrandbits(32)
        a = r0
        for b in data:
            a = t0[((a >> 24) ^ b) & 255] ^ ((a << 8) & M32)
        r = r0
        for i in range(0, len(data), 4):
            x = r ^ (data[i] << 24 | data[i + 1] << 16 | data[i + 2] << 8 | data[i + 3])
            r = tabs[3][x >> 24] ^ tabs[2][(x >> 16) & 255] ^ tabs[1][(x >> 8) & 255] ^ tabs[0][x & 255]
        assert a == r


# ------------------------------------------------------------------ K2: ticket order, window packing
def test_ticket_order_keeps_tiles_of_a_block_in_order():
    for nb, tiles in [(1, 7), (3, 5), (1194, 440), (9809, 3)]:
        seen = {}
        for g in range(min(nb * tiles, 20000)):
            b, t = g % nb, g // nb
            assert seen.get(b, -1) == t - 1        # the tile in front of (b, t) has a smaller ticket
            seen[b] = t


def test_rx_window_packing_takes_every_entry_once():
    LS_T, LOCAL_MAX, CAP = 2048, 1535, 2048 + 1535
    rng = random.Random(5)
    for _ in range(2000):
        nt = rng.randrange(1, 33)
        tcnt = [rng.randrange(0, LS_T + 1) for _ in range(nt)] + [rng.randrange(0, LS_T + 1)]
        tlead = [min(c, rng.randrange(0, LOCAL_MAX)) for c in tcnt]
        taken = 0
        a = 0
        while a < nt:
            e = a + 1
            count = tcnt[a] - tlead[a]
            while e < nt and count + tcnt[e] + tlead[e + 1] <= CAP:
                count += tcnt[e]
                e += 1
            count += tlead[e]
            assert count <= CAP                    # a window never exceeds the staging buffer
            taken += count
            a = e
        # every entry of the CTA's tiles except the first tile's lead entries, plus the next tile's lead entries
        assert taken == sum(tcnt[:nt]) - tlead[0] + tlead[nt]
