"""Deterministic input generators shared by tests and bench.

Integer-only so that any language reproduces them. numpy-vectorised where the
stream is long.
"""
import numpy as np

MASK = (1 << 64) - 1
ALPH = b" etaoinshrdlcumwfgypbvkjxqz"


def _lcg_stream(seed, n):
    """u64 LCG s = s*6364136223846793005 + 1442695040888963407; yields r = s >> 33 (App. B)."""
    out = np.empty(n, dtype=np.uint64)
    s = seed & MASK
    for i in range(n):
        s = (s * 6364136223846793005 + 1442695040888963407) & MASK
        out[i] = s >> 33
    return out


def g1(seed, n):
    """App. B G1: n times { step; emit ALPH[min(r % 27, (r / 27) % 27)] }."""
    r = _lcg_stream(seed, n)
    a = (r % 27).astype(np.int64)
    b = ((r // 27) % 27).astype(np.int64)
    idx = np.minimum(a, b)
    return np.frombuffer(ALPH, dtype=np.uint8)[idx].tobytes()


def g2(seed, n):
    """App. B G2: runs of 1..7 (or 0..599) copies of a random letter, truncated to n bytes."""
    out = bytearray()
    s = seed & MASK
    while len(out) < n:
        s = (s * 6364136223846793005 + 1442695040888963407) & MASK
        r = s >> 33
        ch = ALPH[r % 27]
        if ((r >> 16) & 3) != 0:
            L = 1 + (r >> 5) % 7
        else:
            L = (r >> 18) % 600
        out += bytes([ch]) * L
    return bytes(out[:n])


def splitmix64(seed, n):
    """n outputs of splitmix64 as a numpy uint64 array (vectorised)."""
    with np.errstate(over="ignore"):
        i = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed) + i * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


_LET = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)


def _vocab(seed):
    """8192 words of 1..8 letters; letters skewed to frequent English letters."""
    r = splitmix64(seed ^ 0x766F636162, 8192 * 10).reshape(8192, 10)
    words = []
    for k in range(8192):
        L = 1 + int(r[k, 0] % 4) + int(r[k, 1] % 4) + (1 if k >= 64 else 0)
        idx = np.minimum((r[k, 2:2 + L] % 26).astype(np.int64), ((r[k, 2:2 + L] >> np.uint64(8)) % 26).astype(np.int64)) \
            if L <= 8 else None
        words.append(_LET[idx].tobytes())
    return words


def text(seed, n):
    """English-like synthetic text (Zipf-ish word ranks over an 8192-word vocabulary, sentences,
    ~72-column lines). Deterministic; vectorised with numpy so 1 GiB is practical.

    The stream is produced in chunks of words; each word draw `a` picks j = a % 13 and rank
    k = 2^j - 1 + ((a >> 8) % 2^j); separator by (a >> 40) % 16: 0 -> '. ', 1 -> ', ', else ' ';
    the word after a sentence end is capitalised; every 12th separator space becomes '\\n'.
    """
    words = _vocab(seed)
    wl = np.array([len(w) for w in words], dtype=np.int64)
    maxw = 9
    wtab = np.zeros((8192, maxw), dtype=np.uint8)
    for k, w in enumerate(words):
        wtab[k, :len(w)] = np.frombuffer(w, dtype=np.uint8)
    out = np.empty(n + 64, dtype=np.uint8)
    filled = 0
    chunk = 1 << 20
    ctr = 0
    prev_end_sentence = True
    while filled < n:
        a = splitmix64(seed + 0x1000003 * ctr, chunk)
        ctr += 1
        j = (a % np.uint64(13)).astype(np.int64)
        k = ((np.int64(1) << j) - 1) + ((a >> np.uint64(8)).astype(np.int64) & ((np.int64(1) << j) - 1))
        sepc = ((a >> np.uint64(40)) % np.uint64(16)).astype(np.int64)
        seplen = np.where(sepc <= 1, 2, 1)
        L = wl[k] + seplen
        ends = np.cumsum(L)
        starts = ends - L
        total = int(ends[-1])
        buf = np.full(total, 32, dtype=np.uint8)
        # letters
        for c in range(maxw):
            m = wl[k] > c
            buf[starts[m] + c] = wtab[k[m], c]
        # punctuation
        m0 = sepc == 0
        buf[starts[m0] + wl[k[m0]]] = ord(".")
        m1 = sepc == 1
        buf[starts[m1] + wl[k[m1]]] = ord(",")
        # capitalise the word after a sentence end
        cap = np.empty(chunk, dtype=bool)
        cap[0] = prev_end_sentence
        cap[1:] = m0[:-1]
        prev_end_sentence = bool(m0[-1])
        buf[starts[cap]] -= 32
        # newline instead of the final space of every 12th word
        nl = np.arange(chunk) % 12 == 11
        buf[ends[nl] - 1] = 10
        take = min(total, n - filled)
        out[filled:filled + take] = buf[:take]
        filled += take
    return out[:n].tobytes()


def mixed(seed, n):
    """64 KiB segments cycling {text slice, uniform random bytes, 32-byte records, 16-symbol low-entropy bytes}."""
    seg = 65536
    nseg = (n + seg - 1) // seg
    out = np.empty(nseg * seg, dtype=np.uint8)
    t = np.frombuffer(text(seed, ((nseg + 3) // 4) * seg), dtype=np.uint8)
    ti = 0
    for s in range(nseg):
        kind = s % 4
        o = out[s * seg:(s + 1) * seg]
        if kind == 0:
            o[:] = t[ti * seg:(ti + 1) * seg]
            ti += 1
        elif kind == 1:
            o[:] = (splitmix64(seed * 7919 + s, seg // 8).view(np.uint8))
        elif kind == 2:
            rec = np.zeros((seg // 32, 32), dtype=np.uint8)
            cnt = np.arange(seg // 32, dtype=np.uint32) + np.uint32(s * (seg // 32))
            rec[:, 0:4] = cnt.view(np.uint8).reshape(-1, 4)
            r = splitmix64(seed * 104729 + s, seg // 32)
            rec[:, 16:24] = r.view(np.uint8).reshape(-1, 8)
            nib = (r >> np.uint64(60)).astype(np.uint8)
            rec[:, 24:32] = (nib * 17)[:, None]
            o[:] = rec.reshape(-1)
        else:
            r = splitmix64(seed * 31337 + s, seg // 8).view(np.uint8)
            o[:] = (r & 15) + 65
    return out[:n].tobytes()


def geometric(seed, n, ratio=1.7, nsym=60):
    """i.i.d. bytes with P(40+k) ~ ratio^-(k+1): skewed enough that plain Huffman depths exceed 17, which drives
    the reference's reverse package-merge fallback (cano_huff_table.rs:58-151)."""
    p = np.array([ratio ** -(i + 1) for i in range(nsym)], dtype=np.float64)
    cdf = np.floor(np.cumsum(p / p.sum()) * float(1 << 32)).astype(np.uint64)
    u = splitmix64(seed, n) >> np.uint64(32)
    k = np.minimum(np.searchsorted(cdf, u, side="right"), nsym - 1)
    return (k.astype(np.uint8) + np.uint8(40)).tobytes()
