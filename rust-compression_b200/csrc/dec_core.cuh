// dec_core.cuh — per-thread bodies of the block-parallel bzip2 DEcoder kernels.
//
// replaces, block by block and in parallel, the reference's BZip2Decoder:
//   init_block             /root/reference/src/bzip2/decoder.rs:163-525  (block header, mapping table, selectors,
//                                                                         coding tables, MTF/RUNA/RUNB, T^-1 vector)
//   get_next_lfm           decoder.rs:527-542                            (inverse-BWT step)
//   BitDecodeService::next decoder.rs:545-581                            (RLE1 undo, block CRC)
//   MtfPositionDecoder     bzip2/mtf.rs:41-65
//   canonical codes        huffman/mod.rs:22-67 (codes by (length, symbol)); a code that is not in the table is a
//                          DataError (decoder.rs:376-379)
//
// Every kernel body is a plain function of (thread coordinates, pointers).  decoder.cu wraps each body in a
// __global__ kernel; tests/cpp/dec_emu.cpp compiles the SAME bodies for the host (BZB_EMU) and runs them thread by
// thread, so the algorithm can be checked against the restated reference decoder without a GPU.  The emulation is
// test infrastructure; libbzb200.so contains device code only.
//
// Stages (one .bz2 buffer, all of its blocks at once):
//   D1 d1_scan      every bit offset is tested for the 48-bit block magic 0x314159265359 and the end-of-stream magic
//                   0x177245385090 -> candidate list (the host later keeps the candidates that lie on the chain
//                   "block i ends where block i+1 starts", which is what a sequential parse would have visited)
//   D2 d2_decode    one warp per candidate, lane 0 decodes: header, tables, Huffman symbols, MTF, RUNA/RUNB ->
//                   last column byte L[i] and occurrence index occ[i] = #{j < i : L[j] == L[i]} packed in one word,
//                   byte counts -> cftab
//   D3 d3_scatter   V[cftab[L[i]] + occ[i]] = i << 8 | L[i]   (the reference's tt after decoder.rs:479-484, with the
//                   first-column byte in the low bits so that one load per step yields the output byte)
//   D4 d4_walk_a    list ranking of the walk p -> V[p] >> 8: every SEG-th slot (and origPtr) is a splitter; a thread
//                   walks from its splitter to the next one (length, successor)
//      d4_schedule  one thread per block follows the splitters from origPtr and assigns output offsets; a periodic
//                   block revisits its cycle (cycle length recorded)
//      d4_walk_c    every scheduled splitter copies the bytes its first walk kept (walking on only when the segment is
//                   longer than SEG_KEEP) to its output offset (the block before RLE1 undo)
//   D5 d5_count     RLE1 undo as a 5-state machine (k equal bytes seen so far, k = 4: the next byte is a count): per
//                   chunk and entry state -> exit state and expanded length
//      d5_compose   one thread per block composes the chunk maps -> entry state and output offset of every chunk
//      d5_expand    every chunk re-runs the machine from its entry state and writes the original bytes
//   CRC             k5_crc_blocks (the encoder's kernel) over the output ranges of the blocks, compared on the host
#pragma once
#include <stdint.h>

#ifdef BZB_EMU
#define BZB_DEV inline
#define BZB_HD inline
#else
#define BZB_DEV __device__ __forceinline__
#define BZB_HD __host__ __device__ __forceinline__
#endif

namespace bzb {
namespace dec {

constexpr uint32_t SEG = 1024;          // splitter spacing of the inverse-BWT walks (slots)
constexpr uint32_t SEG_KEEP = 4096;     // bytes of a walk kept by d4_walk_a for d4_walk_c (longer walks are resumed)
constexpr uint32_t RLE_CHUNK = 1024;    // bytes of the pre-RLE1 block per d5 thread
constexpr uint32_t LUT_BITS = 10;       // primary Huffman lookup width
constexpr uint32_t LUT2_BITS = 9;       // window of the two-symbol lookup of the split path (same 2 KB per table)
constexpr uint32_t MAX_SEL = 32768;     // n_selectors is a 15-bit field (decoder.rs:285)
constexpr uint32_t MTF_CHUNK = 1024;    // symbols per thread of the chunk-parallel MTF stage (split path)
constexpr uint64_t KIND_END = 1ull << 63;
constexpr uint64_t MAGIC_BLOCK = 0x314159265359ull;
constexpr uint64_t MAGIC_END = 0x177245385090ull;

// BZip2Error ordinals + 1 (bzip2/error.rs:4-11), 0 = no error
enum : uint32_t { E_OK = 0, E_DATA = 1, E_MAGIC_FIRST = 2, E_MAGIC = 3, E_EOF = 4, E_UNEXPECTED = 5 };

struct CandInfo {       // one per candidate, written by d2_decode / d4_schedule / d5_compose, read by the host
  uint64_t start_bit;   // position of the 48-bit magic
  uint64_t end_bit;     // block: first bit after the EOB code; end of stream: first bit after the stored CRC
  uint32_t kind;        // 0 block, 1 end of stream
  uint32_t err;         // E_* raised while parsing (0: parsed through)
  uint32_t err_early;   // 1: err was raised before origPtr had been read (decoder.rs:231-237)
  uint32_t stored_crc;  // block CRC / combined CRC as stored in the stream
  uint32_t orig_pos;
  uint32_t randomised;
  uint32_t nblock;      // tt.len(): bytes of the block before RLE1 undo
  uint32_t need_max;    // smallest 100000*level that passes decoder.rs:399,427: nblock, +1 if a run was pushed last
  uint32_t cyc;         // 0, or the length of the walk's cycle when it is shorter than nblock (periodic block)
  uint32_t rle_len;     // bytes after RLE1 undo
  uint32_t rle_dangling;  // 1: the block ends with four equal bytes and no count byte
  uint32_t nsym;        // Huffman symbols decoded (incl. EOB)
  uint32_t nsyms;       // bytes in use in the block (the MTF alphabet; EOB = nsyms + 1)
  uint8_t tail[4];      // end of stream: the bytes after the padding (next stream's "BZh" + level, if any)
  uint32_t tail_n;      // how many of them exist
};

// ---------------------------------------------------------------------------------------------------------------
// Bit input, MSB first (bitio/reader.rs with direction Left).  Bytes beyond n read as zero; `read` refuses to go
// beyond n*8 like the reference's reader does.
BZB_DEV uint32_t load_be32(const uint8_t* p, uint64_t n, uint64_t w) {
  const uint64_t o = w * 4;
  if (o + 4 <= n && (((uintptr_t)p) & 3u) == 0) {
    const uint32_t v = *reinterpret_cast<const uint32_t*>(p + o);
    return (v >> 24) | ((v >> 8) & 0xFF00u) | ((v << 8) & 0xFF0000u) | (v << 24);
  }
  uint32_t r = 0;
  for (int k = 0; k < 4; ++k) r = (r << 8) | (o + k < n ? (uint32_t)p[o + k] : 0u);
  return r;
}

struct Reader {
  const uint8_t* p;
  uint64_t n, nbits;
  uint64_t pos;    // next unread bit
  uint64_t buf;    // unread bits, left aligned
  uint32_t cnt;    // valid bits in buf (kept > 32 after refill)
  uint64_t w;      // next 32-bit word to append
  uint32_t ahead;  // word w, loaded one refill early so that its latency overlaps decoding

  BZB_DEV void init(const uint8_t* p_, uint64_t n_, uint64_t pos_) {
    p = p_;
    n = n_;
    nbits = n_ * 8;
    pos = pos_;
    w = pos_ >> 5;
    const uint32_t sh = (uint32_t)(pos_ & 31);
    buf = ((uint64_t)load_be32(p, n, w)) << 32;
    buf <<= sh;
    cnt = 32 - sh;
    ++w;
    ahead = load_be32(p, n, w);
    refill();
  }
  BZB_DEV void refill() {
    if (cnt <= 32) {
      buf |= ((uint64_t)ahead) << (32 - cnt);
      cnt += 32;
      ++w;
      ahead = load_be32(p, n, w);
    }
  }
  // the next k (1..32) bits without consuming them; bits beyond the end of the input are zero
  BZB_DEV uint32_t peek(uint32_t k) const { return (uint32_t)(buf >> (64 - k)); }
  BZB_DEV void skip(uint32_t k) {
    buf <<= k;
    cnt -= k;
    pos += k;
    refill();
  }
  BZB_DEV bool read(uint32_t k, uint32_t& v) {
    if (pos + k > nbits) return false;
    v = peek(k);
    skip(k);
    return true;
  }
};

// ---------------------------------------------------------------------------------------------------------------
// D1: magic scan.  Thread x tests the 32 bit offsets of input word x.
BZB_DEV void d1_scan_body(uint64_t x, const uint8_t* in, uint64_t n, uint64_t* cand, uint32_t* cand_count, uint32_t cap) {
  const uint64_t nbits = n * 8;
  if (x * 32 + 48 > nbits) return;
  const uint64_t hi = ((uint64_t)load_be32(in, n, x) << 32) | load_be32(in, n, x + 1);
  const uint32_t lo = load_be32(in, n, x + 2);
  for (uint32_t s = 0; s < 32; ++s) {
    const uint64_t win = s ? ((hi << s) | ((uint64_t)lo >> (32 - s))) : hi;
    const uint64_t v = win >> 16;
    if (v != MAGIC_BLOCK && v != MAGIC_END) continue;
    const uint64_t bit = x * 32 + s;
    if (bit + 48 > nbits) continue;
#ifdef BZB_EMU
    const uint32_t idx = (*cand_count)++;
#else
    const uint32_t idx = atomicAdd(cand_count, 1u);
#endif
    if (idx < cap) cand[idx] = bit | (v == MAGIC_END ? KIND_END : 0ull);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// D2: one candidate -> header, coding tables, symbols, MTF, RUNA/RUNB.
struct D2Scratch {
  uint16_t lut[6][1u << LUT_BITS];  // len << 9 | sym for codes of at most LUT_BITS bits, 0 = longer or invalid
  uint32_t first_code[6][32];       // per length 1..31 (huffman/mod.rs:22-67)
  uint16_t count[6][32];
  uint16_t offs[6][32];             // start of the length's symbols inside perm
  uint16_t perm[6][258];            // symbols by (length, symbol)
  uint8_t max_len[6];
  uint8_t len[258];
  uint8_t mtf[256];                 // MTF start list holding the block's bytes (seq2unseq applied)
  uint32_t cnt[256];                // occurrences per byte (filled from the lanes' registers after the symbol loop)
  CandInfo info;                    // lane 0's result record
  uint64_t pos;                     // bit position after the header
  uint32_t nsyms, n_sel, go;        // header results handed to all lanes
};

// Canonical codes of one table into scratch; false: over-subscribed (rejected as DataError, as the oracle does).
BZB_DEV bool d2_build_table(D2Scratch* s, uint32_t t, uint32_t alpha, bool pair_lut) {
  for (uint32_t l = 0; l < 32; ++l) {
    s->count[t][l] = 0;
    s->first_code[t][l] = 0;
    s->offs[t][l] = 0;
  }
  uint32_t maxl = 0;
  for (uint32_t i = 0; i < alpha; ++i) {
    const uint32_t l = s->len[i];
    if (l) s->count[t][l]++;
    maxl = l > maxl ? l : maxl;
  }
  s->max_len[t] = (uint8_t)maxl;
  uint64_t code = 0;
  uint32_t prev = 0, off = 0;
  for (uint32_t l = 1; l <= maxl; ++l) {
    if (!s->count[t][l]) continue;
    code <<= (l - prev);
    prev = l;
    s->first_code[t][l] = (uint32_t)code;
    s->offs[t][l] = (uint16_t)off;
    off += s->count[t][l];
    code += s->count[t][l];
    if (code > (1ull << l)) return false;
  }
  // symbols by (length, symbol): a counting sort, stable in the symbol (bucket_sort.rs:43-75)
  {
    uint16_t next[32];
    for (uint32_t l = 0; l < 32; ++l) next[l] = s->offs[t][l];
    for (uint32_t i = 0; i < alpha; ++i) {
      const uint32_t l = s->len[i];
      if (l) s->perm[t][next[l]++] = (uint16_t)i;
    }
  }
  if (pair_lut) {
    // split path: the table's 2 KB hold 512 words instead — for a 9-bit window the first symbol and, when a second
    // code fits into the same window, the second one too:
    //   bits 0-8 sym1, 9-17 sym2, 18-21 len1, 22-25 len1 + len2, bit 26: two symbols;  0 = first code longer than 9 bits
    uint32_t* l2 = reinterpret_cast<uint32_t*>(s->lut[t]);
    for (uint32_t w = 0; w < (1u << LUT2_BITS); ++w) {
      uint32_t sym[2] = {0, 0}, len[2] = {0, 0}, have = 0, used = 0;
      for (uint32_t k = 0; k < 2; ++k) {
        uint32_t found = 0;
        for (uint32_t l = 1; l <= maxl && used + l <= LUT2_BITS; ++l) {
          if (!s->count[t][l]) continue;
          const uint32_t cbits = (w >> (LUT2_BITS - used - l)) & ((1u << l) - 1u);
          const uint32_t f = s->first_code[t][l];
          if (cbits >= f && cbits - f < s->count[t][l]) {
            sym[k] = s->perm[t][s->offs[t][l] + (cbits - f)];
            len[k] = l;
            found = 1;
            break;
          }
        }
        if (!found) break;
        used += len[k];
        have = k + 1;
      }
      uint32_t e = 0;
      if (have >= 1) e = sym[0] | (len[0] << 18) | (len[0] << 22);
      if (have == 2) e = sym[0] | (sym[1] << 9) | (len[0] << 18) | ((len[0] + len[1]) << 22) | (1u << 26);
      l2[w] = e;
    }
    return true;
  }
  for (uint32_t e = 0; e < (1u << LUT_BITS); ++e) s->lut[t][e] = 0;
  for (uint32_t l = 1; l <= maxl && l <= LUT_BITS; ++l) {
    for (uint32_t j = 0; j < s->count[t][l]; ++j) {
      const uint32_t c = s->first_code[t][l] + j;
      const uint16_t v = (uint16_t)((l << 9) | s->perm[t][s->offs[t][l] + j]);
      const uint32_t lo = c << (LUT_BITS - l), hi = (c + 1) << (LUT_BITS - l);
      for (uint32_t e = lo; e < hi; ++e) s->lut[t][e] = v;
    }
  }
  return true;
}

// One Huffman symbol of table t: returns the symbol, or -1 (no such code / input exhausted -> DataError).
BZB_DEV int d2_symbol(const D2Scratch* s, uint32_t t, Reader& r) {
  const uint32_t e = s->lut[t][r.peek(LUT_BITS)];
  if (e) {
    const uint32_t l = e >> 9;
    if (r.pos + l > r.nbits) return -1;
    r.skip(l);
    return (int)(e & 511u);
  }
  const uint32_t maxl = s->max_len[t];
  for (uint32_t l = LUT_BITS + 1; l <= maxl; ++l) {
    if (!s->count[t][l]) continue;
    const uint32_t c = r.peek(l);
    const uint32_t f = s->first_code[t][l];
    if (c >= f && c - f < s->count[t][l]) {
      if (r.pos + l > r.nbits) return -1;
      r.skip(l);
      return (int)s->perm[t][s->offs[t][l] + (c - f)];
    }
  }
  return -1;
}

// ---- warp-resident MTF list: entry 32*q + lane lives in register q of `lane`; every lane runs the (uniform)
// symbol loop, so v below is the same in all lanes.  The host emulation keeps the list in a plain array (one "lane").
#ifdef BZB_EMU
struct LaneState {
  uint8_t list[256];
  void init(const uint8_t* mtf0, uint32_t) {
    for (int i = 0; i < 256; ++i) list[i] = mtf0[i];
  }
  uint32_t front() const { return list[0]; }
  uint32_t pop(uint32_t v) {  // MtfPositionDecoder::pop (mtf.rs:52-64)
    const uint8_t t = list[v];
    for (uint32_t q = v; q > 0; --q) list[q] = list[q - 1];
    list[0] = t;
    return t;
  }
};
BZB_DEV void warp_sync() {}
#else
struct LaneState {
  uint32_t row[8];
  __device__ __forceinline__ void init(const uint8_t* mtf0, uint32_t lane) {
#pragma unroll
    for (int q = 0; q < 8; ++q) row[q] = mtf0[32 * q + lane];
  }
  __device__ __forceinline__ uint32_t front() const { return __shfl_sync(0xffffffffu, row[0], 0); }
  __device__ __forceinline__ uint32_t pop(uint32_t v) {
    const uint32_t lane = threadIdx.x & 31u;
    if (v < 32) {  // the usual case: two independent shuffles of row 0
      const uint32_t t = __shfl_sync(0xffffffffu, row[0], v);
      const uint32_t up = __shfl_up_sync(0xffffffffu, row[0], 1);
      row[0] = lane == 0 ? t : (lane <= v ? up : row[0]);
      return t;
    }
    const uint32_t q = v >> 5, l = v & 31u;
    uint32_t pick = row[1];
#pragma unroll
    for (int k = 2; k < 8; ++k) pick = ((uint32_t)k == q) ? row[k] : pick;
    const uint32_t t = __shfl_sync(0xffffffffu, pick, l);
    uint32_t carry = t;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if ((uint32_t)k <= q) {  // uniform
        const uint32_t up = __shfl_up_sync(0xffffffffu, row[k], 1);
        const uint32_t last = __shfl_sync(0xffffffffu, row[k], 31);
        const uint32_t limit = ((uint32_t)k < q) ? 31u : l;
        row[k] = lane == 0 ? carry : (lane <= limit ? up : row[k]);
        carry = last;
      }
    }
    return t;
  }
};
BZB_DEV void warp_sync() { __syncwarp(); }
#endif

// Bit input of the symbol loop: like Reader, but the position is derived on demand and running past the end of the
// input is only detected after the loop (every error of the symbol phase is a DataError, so the kind is the same
// wherever it is raised; bits beyond the end read as zero and the loop is bounded by n_selectors * 50 symbols).
struct FastBits {
  const uint8_t* p;
  uint64_t n;
  uint64_t buf;
  uint32_t bits;
  uint64_t w;      // index of the word held in `ahead`
  uint32_t ahead;
  BZB_DEV void init(const uint8_t* p_, uint64_t n_, uint64_t pos_) {
    p = p_;
    n = n_;
    w = pos_ >> 5;
    const uint32_t sh = (uint32_t)(pos_ & 31);
    buf = (((uint64_t)load_be32(p, n, w)) << 32) << sh;
    bits = 32 - sh;
    ++w;
    ahead = load_be32(p, n, w);
    drop(0);
  }
  BZB_DEV uint32_t peek(uint32_t k) const { return (uint32_t)(buf >> (64 - k)); }
  BZB_DEV void drop(uint32_t l) {
    buf <<= l;
    bits -= l;
    if (bits <= 32) {
      buf |= ((uint64_t)ahead) << (32 - bits);
      bits += 32;
      ++w;
      ahead = load_be32(p, n, w);
    }
  }
  BZB_DEV uint64_t pos() const { return w * 32 - bits; }
};

// Header phase, one lane: stream position, block header, mapping table, selectors, coding tables.  Leaves the MTF
// start list in s->mtf and the tables in s; s->go = 1 when the symbol phase should run.
BZB_DEV void d2_header(D2Scratch* s, const uint8_t* in, uint64_t n, uint64_t cand, uint8_t* sel, bool pair_lut) {
  CandInfo& I = s->info;
  I.start_bit = cand & ~KIND_END;
  I.end_bit = 0;
  I.kind = (cand & KIND_END) ? 1u : 0u;
  I.err = E_OK;
  I.err_early = 0;
  I.stored_crc = 0;
  I.orig_pos = 0;
  I.randomised = 0;
  I.nblock = 0;
  I.need_max = 0;
  I.cyc = 0;
  I.rle_len = 0;
  I.rle_dangling = 0;
  I.nsym = 0;
  I.nsyms = 0;
  I.tail_n = 0;
  for (int k = 0; k < 4; ++k) I.tail[k] = 0;
  s->go = 0;

  Reader r;
  r.init(in, n, I.start_bit + 48);
  if (I.kind == 1) {  // end of stream (decoder.rs:494-520)
    if (!r.read(32, I.stored_crc)) { I.err = E_EOF; return; }
    I.end_bit = r.pos;
    const uint64_t nb = ((r.pos + 7) >> 3);
    for (uint32_t k = 0; k < 4 && nb + k < n; ++k) {
      I.tail[k] = in[nb + k];
      I.tail_n = k + 1;
    }
    return;
  }
  // ---- block header (decoder.rs:226-241)
  I.err_early = 1;
  if (!r.read(32, I.stored_crc)) { I.err = E_EOF; return; }
  if (!r.read(1, I.randomised)) { I.err = E_EOF; return; }
  if (!r.read(24, I.orig_pos)) { I.err = E_EOF; return; }
  I.err_early = 0;
  // (origPtr > 10 + 100000*level is checked by the host, which knows the stream's level)
  // (a randomised block — bzip2 <= 0.9.0 — is decoded like any other; its bytes are un-randomised after the inverse BWT,
  //  d4_derand_body, as BlockRandomise does in get_next_lfm, decoder.rs:94-116,537-539)
  // ---- mapping table (decoder.rs:243-281)
  uint32_t in_use16;
  if (!r.read(16, in_use16)) { I.err = E_EOF; return; }
  uint32_t nsyms = 0;
  for (uint32_t i = 0; i < 16; ++i) {
    if (!((in_use16 >> (15 - i)) & 1u)) continue;
    uint32_t m;
    if (!r.read(16, m)) { I.err = E_EOF; return; }
    for (uint32_t j = 0; j < 16; ++j)
      if ((m >> (15 - j)) & 1u) s->mtf[nsyms++] = (uint8_t)(i * 16 + j);
  }
  if (nsyms == 0) { I.err = E_DATA; return; }
  const uint32_t alpha = nsyms + 2;
  // ---- selectors (decoder.rs:283-318)
  uint32_t n_groups, n_sel;
  if (!r.read(3, n_groups)) { I.err = E_EOF; return; }
  if (n_groups < 2 || n_groups > 6) { I.err = E_DATA; return; }
  if (!r.read(15, n_sel)) { I.err = E_EOF; return; }
  if (n_sel < 1) { I.err = E_DATA; return; }
  {
    uint8_t sm[6] = {0, 1, 2, 3, 4, 5};
    for (uint32_t k = 0; k < n_sel; ++k) {
      uint32_t j = 0;
      for (;;) {
        uint32_t bit;
        if (!r.read(1, bit)) { I.err = E_EOF; return; }
        if (!bit) break;
        if (++j >= n_groups) { I.err = E_DATA; return; }
      }
      const uint8_t t = sm[j];
      for (uint32_t q = j; q > 0; --q) sm[q] = sm[q - 1];
      sm[0] = t;
      sel[k] = t;
    }
  }
  // ---- coding tables (decoder.rs:320-358)
  for (uint32_t t = 0; t < n_groups; ++t) {
    uint32_t curr;
    if (!r.read(5, curr)) { I.err = E_EOF; return; }
    for (uint32_t i = 0; i < alpha; ++i) {
      for (;;) {
        uint32_t bit;
        if (!r.read(1, bit)) { I.err = E_EOF; return; }
        if (!bit) break;
        if (curr < 1 || curr > 20) { I.err = E_DATA; return; }
        if (!r.read(1, bit)) { I.err = E_EOF; return; }
        if (bit == 0) curr += 1; else curr -= 1;
      }
      s->len[i] = (uint8_t)curr;
    }
    if (!d2_build_table(s, t, alpha, pair_lut)) { I.err = E_DATA; return; }
  }
  s->pos = r.pos;
  s->nsyms = nsyms;
  s->n_sel = n_sel;
  s->go = 1;
  I.nsyms = nsyms;
}

// c: candidate index inside the batch.  cap = 100000 * (largest level of any stream in the buffer): the most a block
// of this buffer may hold; stride >= cap is the per-candidate pitch of L / occ.  Called by all 32 lanes of the
// candidate's warp (the emulation calls it once with lane 0): lane 0 parses the header, then every lane runs the
// symbol loop in lock step — the bit reader and Huffman lookups are replicated, the MTF list and the occurrence
// counters are spread over the lanes' registers, runs are stored by all lanes.
BZB_DEV void d2_decode_body(uint32_t c, uint32_t lane, D2Scratch* s, const uint8_t* in, uint64_t n, const uint64_t* cand,
                            uint32_t cap, uint64_t stride, uint32_t* occbuf, uint8_t* selbuf, uint32_t* cftab,
                            CandInfo* infos) {
  uint32_t* occ = occbuf + (uint64_t)c * stride;  // per position: byte << 24 | occurrence index (< 2^20)
  uint8_t* sel = selbuf + (uint64_t)c * MAX_SEL;
  uint32_t* cf = cftab + (uint64_t)c * 257;
  if (lane == 0) d2_header(s, in, n, cand[c], sel, false);
  warp_sync();
  if (!s->go) {
    if (lane == 0) infos[c] = s->info;
    return;
  }
  const uint32_t nsyms = s->nsyms, n_sel = s->n_sel;
  const uint32_t eob = nsyms + 1;
  FastBits r;
  r.init(in, n, s->pos);
  LaneState st;
  st.init(s->mtf, lane);
#ifdef BZB_EMU
  for (uint32_t k = 0; k < 256; ++k) s->cnt[k] = 0;
#else
  for (uint32_t k = lane; k < 256; k += 32) s->cnt[k] = 0;
#endif
  warp_sync();
  // ---- symbols (decoder.rs:360-444).  Occurrence counters: shared memory, written by lane 0 only.
  uint32_t err = 0;
  uint32_t group_no = 0, group_pos = 0;
  uint32_t tbl_next = sel[0];  // selectors are fetched one group ahead
  const uint16_t* lut = s->lut[0];
  uint32_t tbl = 0;
  uint32_t nn = 1, es = 0;     // nn < 2^21 is enforced below, so es < 2^23
  uint32_t size = 0, need = 0, nsym = 0;
  for (;;) {
    if (group_pos == 0) {
      group_no += 1;
      if (group_no > n_sel) { err = E_DATA; break; }
      group_pos = 50;
      tbl = tbl_next;
      lut = s->lut[tbl];
      tbl_next = group_no < n_sel ? sel[group_no] : 0u;
    }
    group_pos -= 1;
    uint32_t next_sym;
    {
      const uint32_t e = lut[r.peek(LUT_BITS)];
      if (e) {
        r.drop(e >> 9);
        next_sym = e & 511u;
      } else {  // a code longer than LUT_BITS, or no code at all
        const uint32_t maxl = s->max_len[tbl];
        uint32_t found = 0xFFFFFFFFu;
        for (uint32_t l = LUT_BITS + 1; l <= maxl; ++l) {
          if (!s->count[tbl][l]) continue;
          const uint32_t cbits = r.peek(l);
          const uint32_t f = s->first_code[tbl][l];
          if (cbits >= f && cbits - f < s->count[tbl][l]) {
            r.drop(l);
            found = s->perm[tbl][s->offs[tbl][l] + (cbits - f)];
            break;
          }
        }
        if (found == 0xFFFFFFFFu) { err = E_DATA; break; }
        next_sym = found;
      }
    }
    ++nsym;
    if (next_sym > 1) {
      if (es > 0) {  // flush the zero run: es copies of the list front
        const uint32_t uc = st.front();
        if ((uint64_t)size + es >= (uint64_t)cap) { err = E_DATA; break; }  // decoder.rs:399 at the largest level
        warp_sync();
        const uint32_t base = s->cnt[uc];
        warp_sync();
        if (lane == 0) s->cnt[uc] = base + es;
#ifdef BZB_EMU
        for (uint32_t k = 0; k < es; ++k) {
#else
        for (uint32_t k = lane; k < es; k += 32) {
#endif
          occ[size + k] = (uc << 24) | (base + k);
        }
        size += es;
        need = size + 1;
        nn = 1;
        es = 0;
      }
      if (next_sym == eob) break;
      if (size >= cap) { err = E_DATA; break; }              // decoder.rs:427 at the largest level
      const uint32_t v = next_sym - 1;
      if (v >= nsyms) { err = E_DATA; break; }
      const uint32_t uc = st.pop(v);
      const uint32_t o = s->cnt[uc];  // every lane reads, lane 0's value is the one that counts
      if (lane == 0) {
        s->cnt[uc] = o + 1;
        occ[size] = (uc << 24) | o;
      }
      size += 1;
      need = size;
    } else {
      if (nn >= 2u * 1024u * 1024u) { err = E_DATA; break; }  // decoder.rs:416
      if (next_sym == 0) {
        es += nn;
        nn <<= 1;
      } else {
        nn <<= 1;
        es += nn;
      }
    }
  }
  if (!err && r.pos() > n * 8) err = E_DATA;  // the symbols ran past the end of the input (decoder.rs:376-379)
  warp_sync();
  if (lane != 0) return;
  CandInfo& I = s->info;
  I.nsym = nsym;
  if (err) {
    I.err = err;
  } else {
    I.end_bit = r.pos();
    I.nblock = size;
    I.need_max = need;
    if (I.orig_pos >= size) {  // decoder.rs:446-450
      I.err = E_DATA;
    } else {
      uint32_t acc = 0;
      for (uint32_t k = 0; k < 256; ++k) {  // cftab (decoder.rs:452-476)
        cf[k] = acc;
        acc += s->cnt[k];
      }
      cf[256] = acc;
    }
  }
  infos[c] = I;
}

// ---------------------------------------------------------------------------------------------------------------
// D2, split path: the serial chain of a block is cut down to the Huffman symbols alone (d2_huff); the MTF list and the
// RUNA/RUNB expansion become parallel over chunks of MTF_CHUNK symbols by list composition — the inverse of the
// encoder's K3:
//   d2_huff   one lane per candidate: header, tables, symbols -> sym[] (u16), the block's start list mtf0[]
//   d2_mtf_a  one thread per chunk: MTF on the identity list of POSITIONS -> for every symbol the index p into the
//             list the chunk starts with (P[]), the chunk's permutation, bytes emitted per index, and the run digits
//             that lead/trail the chunk (a run may straddle chunks; digit j of a run weighs 2^j)
//   d2_mtf_b  one warp per candidate walks the chunks: start list, per-byte counts and output offset of every chunk;
//             block totals, cftab, and the checks of decoder.rs:399,416,427,446
//   d2_mtf_c  one thread per chunk: byte = start_list[P], occurrence index = count before the chunk + count inside ->
//             the same packed words the fused d2_decode writes
struct ChunkMeta {      // written by d2_mtf_a
  uint32_t leadval;     // sum over the leading run digits j of digit << j  (digit: RUNA 1, RUNB 2)
  uint32_t leadcnt;     // number of leading run digits (the whole chunk if it holds nothing else)
  uint32_t trail;       // run digits at the end of the chunk (index the next chunk's first digit continues from)
  uint32_t rest;        // bytes emitted by everything after the leading run
  uint32_t err;         // 1: a run has more than 21 digits (decoder.rs:416)
  uint32_t has_lit;     // 1: the chunk contains a literal (a chunk without one extends the run it was entered with)
};

BZB_HD uint32_t d2_nchunks(uint32_t nsym) { return (nsym + MTF_CHUNK - 1) / MTF_CHUNK; }

// c: candidate.  Lane 0 only.  symstride >= cap + 4 symbols per candidate.
BZB_DEV void d2_huff_body(uint32_t c, D2Scratch* s, const uint8_t* in, uint64_t n, const uint64_t* cand, uint32_t cap,
                          uint64_t symstride, uint16_t* symbuf, uint8_t* selbuf, uint8_t* mtf0buf, CandInfo* infos) {
  uint8_t* sel = selbuf + (uint64_t)c * MAX_SEL;
  uint16_t* sym = symbuf + (uint64_t)c * symstride;
  d2_header(s, in, n, cand[c], sel, true);
  if (!s->go) {
    infos[c] = s->info;
    return;
  }
  const uint32_t nsyms = s->nsyms, n_sel = s->n_sel;
  const uint32_t eob = nsyms + 1;
  for (uint32_t k = 0; k < 256; ++k) mtf0buf[(uint64_t)c * 256 + k] = k < nsyms ? s->mtf[k] : 0;
  FastBits r;
  r.init(in, n, s->pos);
  uint32_t err = 0, group_no = 0, group_pos = 0, tbl = 0, nsym = 0;
  uint32_t tbl_next = sel[0];
  const uint32_t* lut2 = reinterpret_cast<const uint32_t*>(s->lut[0]);
  for (;;) {
    if (group_pos == 0) {
      group_no += 1;
      if (group_no > n_sel) { err = E_DATA; break; }
      group_pos = 50;
      tbl = tbl_next;
      lut2 = reinterpret_cast<const uint32_t*>(s->lut[tbl]);
      tbl_next = group_no < n_sel ? sel[group_no] : 0u;
    }
    // every symbol but EOB yields at least one byte, so more than cap + 1 symbols cannot pass decoder.rs:399,427
    if (nsym > cap) { err = E_DATA; break; }   // (a step stores at most two symbols: nsym <= cap + 2 <= symstride - 2)
    const uint32_t e = lut2[r.peek(LUT2_BITS)];
    if (e) {
      const uint32_t s1 = e & 511u;
      if ((e >> 26) && group_pos >= 2 && s1 != eob) {  // two symbols of the same group in one step
        const uint32_t s2 = (e >> 9) & 511u;
        r.drop((e >> 22) & 15u);
        sym[nsym] = (uint16_t)s1;
        sym[nsym + 1] = (uint16_t)s2;
        nsym += 2;
        group_pos -= 2;
        if (s2 == eob) break;
      } else {
        r.drop((e >> 18) & 15u);
        sym[nsym++] = (uint16_t)s1;
        group_pos -= 1;
        if (s1 == eob) break;
      }
    } else {  // a code longer than the window, or no code at all
      const uint32_t maxl = s->max_len[tbl];
      uint32_t found = 0xFFFFFFFFu;
      for (uint32_t l = LUT2_BITS + 1; l <= maxl; ++l) {
        if (!s->count[tbl][l]) continue;
        const uint32_t cbits = r.peek(l);
        const uint32_t f = s->first_code[tbl][l];
        if (cbits >= f && cbits - f < s->count[tbl][l]) {
          r.drop(l);
          found = s->perm[tbl][s->offs[tbl][l] + (cbits - f)];
          break;
        }
      }
      if (found == 0xFFFFFFFFu) { err = E_DATA; break; }
      sym[nsym++] = (uint16_t)found;
      group_pos -= 1;
      if (found == eob) break;
    }
  }
  if (!err && r.pos() > n * 8) err = E_DATA;  // the symbols ran past the end of the input (decoder.rs:376-379)
  CandInfo& I = s->info;
  I.nsym = nsym;
  if (err) I.err = err;
  else I.end_bit = r.pos();
  infos[c] = I;
}

// x: chunk, y: candidate.  pl / cnt: 256 entries each of per-thread scratch, entry k at [k * st] (shared memory with
// st = threads per CTA on the GPU — a per-thread local array would thrash L1 at full occupancy).
BZB_DEV void d2_mtf_a_body(uint32_t x, uint32_t y, const CandInfo* infos, uint64_t symstride, const uint16_t* symbuf,
                           uint32_t chunks_pitch, uint8_t* Pbuf, uint8_t* permbuf, uint32_t* cntpbuf, ChunkMeta* metabuf,
                           uint8_t* pl, uint32_t* cnt, uint32_t st) {
  const CandInfo& I = infos[y];
  if (I.kind != 0 || I.err != 0) return;
  const uint32_t nsym = I.nsym;
  if (x >= d2_nchunks(nsym)) return;
  const uint32_t lo = x * MTF_CHUNK, hi = lo + MTF_CHUNK < nsym ? lo + MTF_CHUNK : nsym;
  const uint32_t eob = I.nsyms + 1;
  const uint16_t* sym = symbuf + (uint64_t)y * symstride;
  uint8_t* P = Pbuf + (uint64_t)y * symstride;
  for (uint32_t k = 0; k < 256; ++k) {
    pl[k * st] = (uint8_t)k;
    cnt[k * st] = 0;
  }
  ChunkMeta m;
  m.leadval = 0;
  m.leadcnt = 0;
  m.trail = 0;
  m.rest = 0;
  m.err = 0;
  m.has_lit = 0;
  uint32_t j = 0;  // index of the next digit inside the current (non-leading) run
  for (uint32_t i = lo; i < hi; ++i) {
    const uint32_t sy = sym[i];
    if (sy <= 1) {
      const uint32_t d = sy + 1;
      if (!m.has_lit) {
        if (m.leadcnt > 20) m.err = 1; else m.leadval += d << m.leadcnt;
        m.leadcnt += 1;
        P[i] = 0;
      } else {
        const uint32_t front = pl[0];
        if (j > 20) {
          m.err = 1;
        } else {
          const uint32_t v = d << j;
          m.rest += v;
          cnt[front * st] += v;
        }
        j += 1;
        P[i] = (uint8_t)front;
      }
      m.trail += 1;
    } else if (sy == eob) {
      m.trail = 0;  // EOB flushes whatever run precedes it; nothing continues into a next chunk
      break;
    } else {
      m.has_lit = 1;
      j = 0;
      m.trail = 0;
      const uint32_t v = sy - 1;  // < nsyms: the alphabet has nsyms + 2 symbols
      const uint32_t p = pl[v * st];
      for (uint32_t q = v; q > 0; --q) pl[q * st] = pl[(q - 1) * st];
      pl[0] = (uint8_t)p;
      P[i] = (uint8_t)p;
      cnt[p * st] += 1;
      m.rest += 1;
    }
  }
  const uint64_t co = (uint64_t)y * chunks_pitch + x;
  for (uint32_t k = 0; k < 256; ++k) {
    permbuf[co * 256 + k] = pl[k * st];
    cntpbuf[co * 256 + k] = cnt[k * st];
  }
  metabuf[co] = m;
}

// One warp per candidate (the emulation runs it as one lane that covers every index).
#ifdef BZB_EMU
#define BZB_LANE_LOOP(i, n) for (uint32_t i = 0; i < (n); ++i)
#else
#define BZB_LANE_LOOP(i, n) for (uint32_t i = lane; i < (n); i += 32)
#endif
struct MtfBScratch {
  uint8_t cur[256], nxt[256];
  uint32_t cntb[256];
  uint64_t off;
  uint32_t d0, err, add0;
};

BZB_DEV void d2_mtf_b_body(uint32_t y, uint32_t lane, MtfBScratch* s, CandInfo* infos, uint32_t cap, uint64_t symstride,
                           const uint16_t* symbuf, const uint8_t* mtf0buf, uint32_t chunks_pitch, const uint8_t* permbuf,
                           const uint32_t* cntpbuf, const ChunkMeta* metabuf, uint8_t* initlbuf, uint32_t* basebuf,
                           uint32_t* coffbuf, uint32_t* cd0buf, uint32_t* cftab) {
  (void)lane;
  CandInfo& I = infos[y];
  if (I.kind != 0 || I.err != 0) return;
  const uint32_t nsym = I.nsym, nsyms = I.nsyms;
  const uint32_t nch = d2_nchunks(nsym);
  BZB_LANE_LOOP(i, 256) {
    s->cur[i] = mtf0buf[(uint64_t)y * 256 + i];
    s->cntb[i] = 0;
  }
  if (lane == 0) {
    s->off = 0;
    s->d0 = 0;
    s->err = 0;
  }
  warp_sync();
  for (uint32_t k = 0; k < nch; ++k) {
    const uint64_t co = (uint64_t)y * chunks_pitch + k;
    BZB_LANE_LOOP(i, 256) {
      initlbuf[co * 256 + i] = s->cur[i];
      basebuf[co * 256 + i] = s->cntb[i];
    }
    if (lane == 0) {
      const ChunkMeta m = metabuf[co];
      uint32_t e = m.err;
      const uint32_t d0 = s->d0;
      if (m.leadcnt && d0 + m.leadcnt > 21) e = 1;  // decoder.rs:416: a run of more than 21 digits
      uint64_t lead = e ? 0 : ((uint64_t)m.leadval << d0);
      if (lead > cap) {  // decoder.rs:399
        e = 1;
        lead = 0;
      }
      coffbuf[co] = (uint32_t)(s->off > cap ? cap : s->off);
      cd0buf[co] = d0;
      s->off += lead + m.rest;
      s->add0 = (uint32_t)lead;
      s->d0 = m.has_lit ? m.trail : d0 + m.leadcnt;
      if (e) s->err = 1;
    }
    warp_sync();
    BZB_LANE_LOOP(p, 256) {  // cur is a permutation: every p adds to a different byte's counter
      const uint32_t add = cntpbuf[co * 256 + p] + (p == 0 ? s->add0 : 0u);
      if (add) s->cntb[s->cur[p]] += add;
    }
    BZB_LANE_LOOP(i, 256) s->nxt[i] = s->cur[permbuf[co * 256 + i]];
    warp_sync();
    BZB_LANE_LOOP(i, 256) s->cur[i] = s->nxt[i];
    // read before the barrier: lane 0 changes off/err again only after it (next iteration), so every lane sees the
    // same values and the loop exit is uniform
    const bool stop = s->err || s->off > cap;
    warp_sync();
    if (stop) break;
  }
  if (lane != 0) return;
  // was the last byte pushed by a run? (decides whether decoder.rs:399 or :427 was the last size check)
  bool last_is_run = false;
  if (nsym >= 2) last_is_run = symbuf[(uint64_t)y * symstride + nsym - 2] <= 1;
  const uint64_t size = s->off;
  if (s->err || size + (last_is_run ? 1 : 0) > cap) {
    I.err = E_DATA;
    return;
  }
  I.nblock = (uint32_t)size;
  I.need_max = (uint32_t)size + (last_is_run ? 1u : 0u);
  if (I.orig_pos >= size) {  // decoder.rs:446-450
    I.err = E_DATA;
    return;
  }
  (void)nsyms;
  uint32_t acc = 0;
  uint32_t* cf = cftab + (uint64_t)y * 257;
  for (uint32_t k = 0; k < 256; ++k) {
    cf[k] = acc;
    acc += s->cntb[k];
  }
  cf[256] = acc;
}

// il / loc: per-thread scratch like in d2_mtf_a (start list of the chunk; running occurrence index per byte, which
// starts at the count before the chunk)
BZB_DEV void d2_mtf_c_body(uint32_t x, uint32_t y, const CandInfo* infos, uint64_t symstride, const uint16_t* symbuf,
                           const uint8_t* Pbuf, uint32_t chunks_pitch, const uint8_t* initlbuf, const uint32_t* basebuf,
                           const uint32_t* coffbuf, const uint32_t* cd0buf, uint64_t stride, uint32_t* occbuf,
                           uint8_t* il, uint32_t* loc, uint32_t st) {
  const CandInfo& I = infos[y];
  if (I.kind != 0 || I.err != 0) return;
  const uint32_t nsym = I.nsym;
  if (x >= d2_nchunks(nsym)) return;
  const uint32_t lo = x * MTF_CHUNK, hi = lo + MTF_CHUNK < nsym ? lo + MTF_CHUNK : nsym;
  const uint32_t eob = I.nsyms + 1;
  const uint16_t* sym = symbuf + (uint64_t)y * symstride;
  const uint8_t* P = Pbuf + (uint64_t)y * symstride;
  const uint64_t co = (uint64_t)y * chunks_pitch + x;
  uint32_t* occ = occbuf + (uint64_t)y * stride + coffbuf[co];
  const uint32_t d0 = cd0buf[co];
  for (uint32_t k = 0; k < 256; ++k) {
    il[k * st] = initlbuf[co * 256 + k];
    loc[k * st] = basebuf[co * 256 + k];
  }
  bool has_lit = false;
  uint32_t j = 0, jl = 0;
  for (uint32_t i = lo; i < hi; ++i) {
    const uint32_t sy = sym[i];
    if (sy == eob) break;
    const uint32_t b = il[(uint32_t)P[i] * st];
    uint32_t v = 1;
    if (sy <= 1) {
      if (!has_lit) v = (sy + 1) << (d0 + jl++);
      else v = (sy + 1) << j++;
    } else {
      has_lit = true;
      j = 0;
    }
    const uint32_t o = loc[b * st];
    loc[b * st] = o + v;
    const uint32_t w = (b << 24) | o;
    for (uint32_t t = 0; t < v; ++t) occ[t] = w + t;
    occ += v;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// D3: x = position inside the block, y = candidate.
BZB_DEV void d3_scatter_body(uint32_t x, uint32_t y, const CandInfo* infos, uint64_t stride, const uint32_t* occbuf,
                             const uint32_t* cftab, uint32_t* Vbuf) {
  const CandInfo& I = infos[y];
  if (I.kind != 0 || I.err != 0 || x >= I.nblock) return;
  const uint64_t base = (uint64_t)y * stride;
  const uint32_t w = occbuf[base + x];
  const uint32_t b = w >> 24;
  const uint32_t slot = cftab[(uint64_t)y * 257 + b] + (w & 0xFFFFFFu);
  Vbuf[base + slot] = (x << 8) | b;
}

// ---------------------------------------------------------------------------------------------------------------
// D4: list ranking.  Splitter ids: x < nseg0 = ceil(nblock / SEG) is slot x*SEG; id nseg0 is slot origPtr (unused
// when origPtr is a multiple of SEG).  segs_pitch = splitters per candidate the arrays were sized for.
BZB_HD uint32_t d4_nseg0(uint32_t nblock) { return (nblock + SEG - 1) / SEG; }

// The walk also keeps the first SEG_KEEP bytes it passes (4 to a word, in the segment's own SEG_KEEP-byte slot of
// Tbuf) and the slot it has reached by then, so that d4_walk_c copies instead of walking a second time.
BZB_DEV void d4_walk_a_body(uint32_t x, uint32_t y, const CandInfo* infos, uint64_t stride, const uint32_t* Vbuf,
                            uint32_t segs_pitch, uint32_t* seg_len, uint32_t* seg_next, uint32_t* seg_resume,
                            uint8_t* Tbuf) {
  const CandInfo& I = infos[y];
  if (I.kind != 0 || I.err != 0) return;
  const uint32_t n0 = d4_nseg0(I.nblock);
  const uint32_t orig = I.orig_pos;
  const bool orig_on_grid = (orig % SEG) == 0;
  uint32_t pos;
  if (x < n0) pos = x * SEG;
  else if (x == n0 && !orig_on_grid) pos = orig;
  else return;
  const uint32_t* V = Vbuf + (uint64_t)y * stride;
  const uint64_t o = (uint64_t)y * segs_pitch + x;
  uint32_t* T = reinterpret_cast<uint32_t*>(Tbuf + o * SEG_KEEP);
  uint32_t len = 0, acc = 0, resume = 0;
  do {
    const uint32_t v = V[pos];
    pos = v >> 8;
    if (len < SEG_KEEP) {
      acc |= (v & 255u) << (8 * (len & 3u));
      if ((len & 3u) == 3u) {
        T[len >> 2] = acc;
        acc = 0;
      }
      if (len + 1 == SEG_KEEP) resume = pos;
    }
    ++len;
  } while ((pos % SEG) != 0 && pos != orig);
  if (len < SEG_KEEP && (len & 3u)) T[len >> 2] = acc;  // the last, partial word
  seg_len[o] = len;
  seg_next[o] = (pos == orig && !orig_on_grid) ? n0 : pos / SEG;
  seg_resume[o] = resume;
}

// one thread per candidate; seg_off must be filled with 0xFFFFFFFF beforehand
BZB_DEV void d4_schedule_body(uint32_t y, CandInfo* infos, uint32_t segs_pitch, const uint32_t* seg_len,
                              const uint32_t* seg_next, uint32_t* seg_off) {
  CandInfo& I = infos[y];
  if (I.kind != 0 || I.err != 0) return;
  const uint32_t n0 = d4_nseg0(I.nblock);
  const uint64_t base = (uint64_t)y * segs_pitch;
  uint32_t cur = (I.orig_pos % SEG) == 0 ? I.orig_pos / SEG : n0;
  uint32_t off = 0;
  uint32_t cyc = 0;
  while (off < I.nblock) {
    if (seg_off[base + cur] != 0xFFFFFFFFu) {  // back at the start: the walk is a cycle shorter than the block
      cyc = off;
      break;
    }
    seg_off[base + cur] = off;
    off += seg_len[base + cur];
    cur = seg_next[base + cur];
  }
  I.cyc = cyc;
}

BZB_DEV void d4_walk_c_body(uint32_t x, uint32_t y, const CandInfo* infos, uint64_t stride, const uint32_t* Vbuf,
                            uint32_t segs_pitch, const uint32_t* seg_len, const uint32_t* seg_off,
                            const uint32_t* seg_resume, const uint8_t* Tbuf, uint8_t* Wbuf) {
  const CandInfo& I = infos[y];
  if (I.kind != 0 || I.err != 0) return;
  const uint32_t n0 = d4_nseg0(I.nblock);
  if (x > n0) return;
  const uint64_t so = (uint64_t)y * segs_pitch + x;
  const uint32_t off = seg_off[so];
  if (off == 0xFFFFFFFFu) return;
  const uint32_t len = seg_len[so];
  const uint32_t n = I.nblock;
  const uint32_t* V = Vbuf + (uint64_t)y * stride;
  const uint32_t* T = reinterpret_cast<const uint32_t*>(Tbuf + so * SEG_KEEP);
  uint8_t* W = Wbuf + (uint64_t)y * stride;
  if (I.cyc == 0) {
    // bytes off .. off+end-1: the first SEG_KEEP come from the kept copy, the rest (long segments) from walking on;
    // stores are byte-wise up to a 4-byte boundary of W, then whole words
    const uint32_t end = off + len < n ? len : n - off;
    const uint32_t kept = end < SEG_KEEP ? end : SEG_KEEP;
    uint32_t pos = seg_resume[so];
    uint32_t acc = 0, have = 0, src = 0;
    for (uint32_t k = 0; k < end; ++k) {
      uint32_t byte;
      if (k < kept) {
        if ((k & 3u) == 0) src = T[k >> 2];
        byte = (src >> (8 * (k & 3u))) & 255u;
      } else {
        const uint32_t v = V[pos];
        pos = v >> 8;
        byte = v & 255u;
      }
      const uint32_t o = off + k;
      if (have == 0 && ((o & 3u) != 0 || end - k < 4)) {
        W[o] = (uint8_t)byte;
        continue;
      }
      acc |= byte << (8 * have);
      if (++have == 4) {
        *reinterpret_cast<uint32_t*>(W + (o - 3)) = acc;
        acc = 0;
        have = 0;
      }
    }
  } else {  // periodic block: the cycle is replayed until the block is full
    const uint32_t cyc = I.cyc;
    uint32_t pos = x < n0 ? x * SEG : I.orig_pos;
    for (uint32_t k = 0; k < len; ++k) {
      const uint32_t v = V[pos];
      pos = v >> 8;
      for (uint64_t o = (uint64_t)off + k; o < n; o += cyc) W[o] = (uint8_t)v;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// D5: RLE1 undo.  State k = number of equal bytes seen so far in the current run (0 fresh, 1..3; 4: the next byte
// is a count).  For k in 1..3 the run's byte is the byte before the chunk.
struct RleStep {
  uint32_t k;
  uint32_t prev;
};

BZB_HD uint32_t d5_nchunks(uint32_t nblock) { return (nblock + RLE_CHUNK - 1) / RLE_CHUNK; }

// Un-randomise (blocks written by bzip2 <= 0.9.0 with the `randomised` bit set; BlockRandomise, decoder.rs:94-116, applied
// in get_next_lfm, :537-539): byte p of the block in inverse-BWT order — count bytes included — is XOR-ed with 1 on the
// calls that leave n2go == 1, i.e. p = k * BZ_RAND_PERIOD + cum[m + 1] - 2 (cum = prefix sums of BZ2_rNums,
// bz_rand_table.h).  Thread x handles toggle number x = 512 k + m of candidate y; the launch is skipped when no block of
// the batch is randomised.
BZB_DEV void d4_derand_body(uint32_t x, uint32_t y, const CandInfo* infos, uint64_t stride, const uint32_t* cum,
                            uint32_t period, uint8_t* Wbuf) {
  const CandInfo& I = infos[y];
  if (I.kind != 0 || I.err || !I.randomised) return;
  const uint64_t p = (uint64_t)(x >> 9) * period + cum[(x & 511u) + 1] - 2u;
  if (p < I.nblock) Wbuf[(uint64_t)y * stride + p] ^= 1u;
}

// map entry: exit state in bits 29..31, expanded length in bits 0..28.  All five entry states are advanced in one
// pass over the chunk (states that cannot occur at this chunk boundary are computed too and never used).
BZB_DEV void d5_count_body(uint32_t x, uint32_t y, const CandInfo* infos, uint64_t stride, const uint8_t* Wbuf,
                           uint32_t chunks_pitch, uint32_t* rle_map /*[cand][chunk][5]*/) {
  const CandInfo& I = infos[y];
  if (I.kind != 0 || I.err != 0) return;
  const uint32_t n = I.nblock;
  if (x >= d5_nchunks(n)) return;
  const uint8_t* W = Wbuf + (uint64_t)y * stride;
  const uint32_t lo = x * RLE_CHUNK, hi = lo + RLE_CHUNK < n ? lo + RLE_CHUNK : n;
  uint32_t prev = lo ? W[lo - 1] : 0x100u;
  uint32_t k[5] = {0, 1, 2, 3, 4}, len[5] = {0, 0, 0, 0, 0};
  for (uint32_t i = lo; i < hi; ++i) {
    const uint32_t b = W[i];
#pragma unroll
    for (int s = 0; s < 5; ++s) {
      if (k[s] == 4) {
        len[s] += b;
        k[s] = 0;
      } else if (k[s] > 0 && b == prev) {
        ++k[s];
        ++len[s];
      } else {
        k[s] = 1;
        ++len[s];
      }
    }
    prev = b;
  }
  uint32_t* m = rle_map + ((uint64_t)y * chunks_pitch + x) * 5;
#pragma unroll
  for (int s = 0; s < 5; ++s) m[s] = (k[s] << 29) | len[s];
}

BZB_DEV void d5_compose_body(uint32_t y, CandInfo* infos, uint32_t chunks_pitch, const uint32_t* rle_map,
                             uint32_t* chunk_entry /*[cand][chunk]: state << 29 ... */, uint64_t* chunk_off) {
  CandInfo& I = infos[y];
  if (I.kind != 0 || I.err != 0) return;
  const uint32_t nc = d5_nchunks(I.nblock);
  uint32_t k = 0;
  uint64_t off = 0;
  for (uint32_t ch = 0; ch < nc; ++ch) {
    const uint64_t o = (uint64_t)y * chunks_pitch + ch;
    chunk_entry[o] = k;
    chunk_off[o] = off;
    const uint32_t m = rle_map[o * 5 + k];
    k = m >> 29;
    off += m & 0x1FFFFFFFu;
  }
  I.rle_len = (uint32_t)(off > 0xFFFFFFFFull ? 0xFFFFFFFFull : off);
  I.rle_dangling = k == 4 ? 1u : 0u;
}

// out_off[y] = byte offset of the block's original bytes in the output, or ~0 when the block is not on the chain
BZB_DEV void d5_expand_body(uint32_t x, uint32_t y, const CandInfo* infos, uint64_t stride, const uint8_t* Wbuf,
                            uint32_t chunks_pitch, const uint32_t* chunk_entry, const uint64_t* chunk_off,
                            const uint64_t* out_off, uint8_t* out) {
  const CandInfo& I = infos[y];
  if (I.kind != 0 || I.err != 0) return;
  if (out_off[y] == ~0ull) return;
  const uint32_t n = I.nblock;
  if (x >= d5_nchunks(n)) return;
  const uint8_t* W = Wbuf + (uint64_t)y * stride;
  const uint32_t lo = x * RLE_CHUNK, hi = lo + RLE_CHUNK < n ? lo + RLE_CHUNK : n;
  const uint64_t co = (uint64_t)y * chunks_pitch + x;
  uint32_t k = chunk_entry[co];
  uint8_t* o = out + out_off[y] + chunk_off[co];
  uint32_t prev = lo ? W[lo - 1] : 0x100u;
  for (uint32_t i = lo; i < hi; ++i) {
    const uint32_t b = W[i];
    if (k == 4) {
      for (uint32_t q = 0; q < b; ++q) o[q] = (uint8_t)prev;
      o += b;
      k = 0;
      // `prev` keeps the run's byte, but the next byte starts a new run whatever it is (result_count >= 4)
      continue;
    }
    if (k > 0 && b == prev) ++k;
    else k = 1;
    *o++ = (uint8_t)b;
    prev = b;
  }
}

}  // namespace dec
}  // namespace bzb
