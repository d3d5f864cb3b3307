// k6_pack.cu — K6 (block header + symbol bit packing) and K7 (bit-granular concatenation), stream framing.
//
// Replaces, by result: the per-block section of write_block (src/bzip2/encoder.rs:253-277), the origPtr push
// (:332-334), mapping table / selectors / coding tables / block data emission (:526-629), HuffmanEncoder::enc
// (src/huffman/encoder.rs:42-54) and BitWriter<Left>::write_bits/flush (src/bitio/writer.rs:186-242): an MSB-first
// concatenation of (value,len) fields.  Bit lengths are known before anything is written (k4 computed the header
// size and the per-group data sizes), so every block and every 50-symbol group is packed directly at its final bit
// offset; only words that straddle two writers are touched with atomicOr.
#include "common.cuh"
#include "kernels.h"

namespace bzb {

constexpr int META = 8;

// Sequential MSB-first writer with a 64-bit accumulator. Interior (fully owned) words may be stored plainly.
template <bool kPlainInterior>
struct SeqBits {
  uint32_t* words;
  uint64_t w;
  unsigned long long acc;
  int nacc;
  bool first;
  __device__ __forceinline__ void init(uint32_t* out_words, uint64_t pos) {
    words = out_words;
    w = pos >> 5;
    nacc = (int)(pos & 31);
    acc = 0;
    first = true;
  }
  __device__ __forceinline__ void put(uint32_t v, int len) {
    if (len == 0) return;
    if (len < 32) v &= (1u << len) - 1u;
    acc = (acc << len) | v;
    nacc += len;
    if (nacc >= 32) {
      uint32_t word = (uint32_t)(acc >> (nacc - 32));
      if (kPlainInterior && !first) words[w] = bswap32(word);
      else if (word) atomicOr(&words[w], bswap32(word));
      first = false;
      ++w;
      nacc -= 32;
      acc &= (1ull << nacc) - 1ull;
    }
  }
  __device__ __forceinline__ void finish() {
    if (nacc > 0) {
      uint32_t word = (uint32_t)(acc << (32 - nacc));
      if (word) atomicOr(&words[w], bswap32(word));
    }
  }
};

// ---- bit offsets of the blocks of this batch ----
constexpr int BO_NT = 1024;
__global__ void __launch_bounds__(BO_NT) k6_block_offsets(uint32_t nb, const uint32_t* __restrict__ meta,
                                                          uint64_t* __restrict__ bitcursor,
                                                          uint64_t* __restrict__ blockbit) {
  __shared__ unsigned long long ws[BO_NT / 32 + 1];
  const uint32_t per = (nb + BO_NT - 1) / BO_NT;
  const uint32_t lo = min(nb, per * threadIdx.x), hi = min(nb, lo + per);
  unsigned long long s = 0;
  for (uint32_t b = lo; b < hi; ++b) s += (unsigned long long)meta[(size_t)b * META + 3] + meta[(size_t)b * META + 4];
  unsigned long long inc = s;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d);
    if ((int)lane_id() >= d) inc += t;
  }
  const int w = threadIdx.x >> 5;
  if (lane_id() == 31) ws[w] = inc;
  __syncthreads();
  if (w == 0) {
    unsigned long long x = ws[lane_id()], xi = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      unsigned long long t = __shfl_up_sync(0xffffffffu, xi, d);
      if ((int)lane_id() >= d) xi += t;
    }
    ws[lane_id()] = xi - x;
    if (lane_id() == 31) ws[32] = xi;
  }
  __syncthreads();
  const unsigned long long start = *bitcursor;
  unsigned long long run = start + inc - s + ws[w];
  for (uint32_t b = lo; b < hi; ++b) {
    blockbit[b] = run;
    run += (unsigned long long)meta[(size_t)b * META + 3] + meta[(size_t)b * META + 4];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    blockbit[nb] = start + ws[32];
    *bitcursor = start + ws[32];
  }
}

// ---- block header: one CTA per block ----
// Thread 0 writes the fixed part; the selector codes are cut into one chunk per thread (bit offsets by a CTA scan of
// the code lengths); the coding tables are written by one thread each.  Every writer ORs into the stream with
// atomics, so chunks may share boundary words.
constexpr int PH_NT = 256;
__global__ void __launch_bounds__(PH_NT) k6_pack_header(uint32_t nb, const uint32_t* __restrict__ meta,
                                                        const uint32_t* __restrict__ inuse,
                                                        const uint32_t* __restrict__ crc,
                                                        const uint32_t* __restrict__ origptr,
                                                        const uint8_t* __restrict__ selmtf,
                                                        const uint8_t* __restrict__ lens_final_base,
                                                        size_t lens_block_stride, size_t lens_final_off,
                                                        const uint64_t* __restrict__ blockbit,
                                                        uint32_t* __restrict__ out_words) {
  __shared__ uint32_t ws[PH_NT / 32 + 1];
  __shared__ uint32_t s_tab[MAX_GROUPS + 1];
  const uint32_t b = blockIdx.x;
  const uint32_t* m = meta + (size_t)b * META;
  const int alpha = (int)m[0], ng = (int)m[1];
  const uint32_t nsel = m[2];
  const uint64_t base = blockbit[b];
  uint32_t used16 = 0, nused = 0;
  for (int r = 0; r < 16; ++r) {
    uint32_t wv = inuse[b * 8 + (r >> 1)];
    uint32_t h = (r & 1) ? (wv >> 16) : (wv & 0xFFFFu);
    used16 = (used16 << 1) | (h ? 1u : 0u);
    nused += h ? 1u : 0u;
  }
  const uint32_t fixed_bits = 48 + 32 + 1 + 24 + 16 + 16 * nused + 3 + 15;
  if (threadIdx.x == 0) {
    SeqBits<false> bw;
    bw.init(out_words, base);
    bw.put(0x314159u, 24);  // block magic 0x314159265359 (encoder.rs:254-259)
    bw.put(0x265359u, 24);
    bw.put(crc[b], 32);     // (:262)
    bw.put(0, 1);           // randomised bit (:273)
    bw.put(origptr[b], 24); // (:333)
    bw.put(used16, 16);     // mapping table (:528-554; bitset.rs:186-199)
    for (int r = 0; r < 16; ++r) {
      uint32_t wv = inuse[b * 8 + (r >> 1)];
      uint32_t h = (r & 1) ? (wv >> 16) : (wv & 0xFFFFu);
      if (h) bw.put(__brev(h) >> 16, 16);  // symbol 16r+j is bit j of h; it is written j-th, i.e. MSB first
    }
    bw.put((uint32_t)ng, 3);   // (:569)
    bw.put(nsel, 15);          // (:570)
    bw.finish();
  }
  // selector MTF values j as j ones and a zero (:572-574)
  const uint8_t* sm = selmtf + (size_t)b * MAX_SELECTORS;
  const uint32_t per = (nsel + PH_NT - 1) / PH_NT;
  const uint32_t lo = min(nsel, per * threadIdx.x), hi = min(nsel, lo + per);
  uint32_t mybits = 0;
  for (uint32_t i = lo; i < hi; ++i) mybits += (uint32_t)sm[i] + 1u;
  uint32_t selbits;
  const uint32_t ex = cta_excl_scan_add<PH_NT>(mybits, ws, &selbits);
  if (hi > lo) {
    SeqBits<false> bw;
    bw.init(out_words, base + fixed_bits + ex);
    for (uint32_t i = lo; i < hi; ++i) {
      const uint32_t j = sm[i];
      bw.put((1u << (j + 1)) - 2u, (int)j + 1);
    }
    bw.finish();
  }
  // coding tables (:585-601): 5-bit start length, then per symbol (10 | 11)* 0
  if (threadIdx.x < (uint32_t)ng) {
    const uint8_t* l = lens_final_base + (size_t)b * lens_block_stride + lens_final_off + (size_t)threadIdx.x * MAX_ALPHA;
    uint32_t bits = 5;
    int curr = l[0];
    for (int sy = 0; sy < alpha; ++sy) {
      const int d = (int)l[sy] - curr;
      bits += 1 + 2 * (uint32_t)(d < 0 ? -d : d);
      curr = l[sy];
    }
    s_tab[threadIdx.x] = bits;
  }
  __syncthreads();
  if (threadIdx.x < (uint32_t)ng) {
    uint32_t off = 0;
    for (uint32_t t = 0; t < threadIdx.x; ++t) off += s_tab[t];
    const uint8_t* l = lens_final_base + (size_t)b * lens_block_stride + lens_final_off + (size_t)threadIdx.x * MAX_ALPHA;
    SeqBits<false> bw;
    bw.init(out_words, base + fixed_bits + selbits + off);
    int curr = l[0];
    bw.put((uint32_t)curr, 5);
    for (int sy = 0; sy < alpha; ++sy) {
      const int li = l[sy];
      while (curr < li) { bw.put(2, 2); ++curr; }
      while (curr > li) { bw.put(3, 2); --curr; }
      bw.put(0, 1);
    }
    bw.finish();
  }
}

// ---- block data: one thread per 50-symbol group ----
constexpr int PS_NT = 256;
__global__ void __launch_bounds__(PS_NT) k6_pack_symbols(const uint16_t* __restrict__ sym,
                                                         const BlockDesc* __restrict__ desc,
                                                         const uint32_t* __restrict__ mtf_count,
                                                         const uint32_t* __restrict__ meta,
                                                         const uint8_t* __restrict__ sel,
                                                         const uint32_t* __restrict__ codes,
                                                         const uint32_t* __restrict__ gbits,
                                                         const uint64_t* __restrict__ blockbit,
                                                         uint32_t* __restrict__ out_words) {
  __shared__ uint32_t s_codes[MAX_GROUPS * MAX_ALPHA];
  const uint32_t b = blockIdx.y;
  const uint32_t* m = meta + (size_t)b * META;
  const uint32_t nsel = m[2];
  if (blockIdx.x * PS_NT >= nsel) return;
  const int ng = (int)m[1];
  const uint32_t* cb = codes + (size_t)b * MAX_GROUPS * MAX_ALPHA;
  for (int i = threadIdx.x; i < ng * MAX_ALPHA; i += PS_NT) s_codes[i] = cb[i];
  __syncthreads();
  const uint32_t g = blockIdx.x * PS_NT + threadIdx.x;
  if (g >= nsel) return;
  const uint32_t mc = mtf_count[b];
  const uint32_t gs = g * G_SIZE, ge = min(gs + G_SIZE, mc);
  const uint16_t* sp = sym + desc[b].symoff;
  const uint32_t* ct = s_codes + (int)sel[(size_t)b * MAX_SELECTORS + g] * MAX_ALPHA;
  SeqBits<true> bw;
  bw.init(out_words, blockbit[b] + m[3] + gbits[(size_t)b * MAX_SELECTORS + g]);
  for (uint32_t i = gs; i < ge; ++i) {
    uint32_t c = ct[sp[i]];
    bw.put(c & 0xFFFFFFu, (int)(c >> 24));
  }
  bw.finish();
}

// ---- K7: OR nbits of src (from bit 0) into dst at bit offset dst_bit ----
__global__ void __launch_bounds__(256) k7_bit_append(uint32_t* __restrict__ dst_words, uint64_t dst_bit,
                                                     const uint32_t* __restrict__ src_words, uint64_t nbits) {
  const uint64_t w0 = dst_bit >> 5;
  const uint32_t sh = (uint32_t)(dst_bit & 31);
  const uint64_t ndw = (sh + nbits + 31) >> 5;  // destination words touched
  const uint64_t nsw = (nbits + 31) >> 5;       // source words holding data
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < ndw; j += (uint64_t)gridDim.x * blockDim.x) {
    auto srcw = [&](long long k) -> uint32_t {
      if (k < 0 || (uint64_t)k >= nsw) return 0u;
      uint32_t v = bswap32(src_words[k]);
      uint64_t rem = nbits - (uint64_t)k * 32;
      if (rem < 32) v &= ~((1u << (32 - (uint32_t)rem)) - 1u);
      return v;
    };
    uint32_t v;
    if (sh == 0) v = srcw((long long)j);
    else v = (srcw((long long)j - 1) << (32 - sh)) | (srcw((long long)j) >> sh);
    if (j == 0 || j == ndw - 1) { if (v) atomicOr(&dst_words[w0 + j], bswap32(v)); }
    else dst_words[w0 + j] = bswap32(v);
  }
}

__global__ void k6_write_stream_header(int level, uint32_t* __restrict__ out_words) {
  // 'B','Z','h','0'+level (encoder.rs:245-251)
  uint32_t v = (0x42u << 24) | (0x5Au << 16) | (0x68u << 8) | (uint32_t)(0x30 + level);
  atomicOr(&out_words[0], bswap32(v));
}

__global__ void k6_write_trailer(uint32_t* __restrict__ out_words, uint64_t at_bit, const uint64_t* __restrict__ at_bit_dev,
                                 uint32_t crc, const uint32_t* __restrict__ crc_dev) {
  // end magic 0x177245385090 + combined CRC (encoder.rs:279-289)
  uint64_t pos = at_bit_dev ? *at_bit_dev : at_bit;
  uint32_t c = crc_dev ? *crc_dev : crc;
  put_bits_atomic(out_words, pos, 0x177245u, 24);
  put_bits_atomic(out_words, pos + 24, 0x385090u, 24);
  put_bits_atomic(out_words, pos + 48, c, 32);
}

// combined = rotl1(combined) ^ crc[i] (encoder.rs:237-238), seeded with *combined
__global__ void k6_combine_crc(const uint32_t* __restrict__ crc, uint32_t n, uint32_t* __restrict__ combined) {
  uint32_t c = *combined;
  for (uint32_t i = 0; i < n; ++i) c = ((c << 1) | (c >> 31)) ^ crc[i];
  *combined = c;
}

void launch_pack(Launcher& L, const uint16_t* d_sym, const BlockDesc* d_desc, const uint32_t* d_mtf_count,
                 const uint32_t* d_inuse, const uint32_t* d_crc, const uint32_t* d_origptr, uint32_t nb,
                 uint32_t max_groups_per_block, HuffBuffers& H, uint64_t* d_blockbit, uint64_t* d_bitcursor,
                 uint8_t* d_out) {
  // Two phases so that the host can check the output capacity in between: d_bitcursor != nullptr computes the
  // bit offsets (and advances the cursor); d_out != nullptr writes the bits.
  if (d_bitcursor)
    L.launch("k6_block_offsets", k6_block_offsets, dim3(1), dim3(BO_NT), nb, (const uint32_t*)H.meta, d_bitcursor,
             d_blockbit);
  if (!d_out) return;
  uint32_t* words = reinterpret_cast<uint32_t*>(d_out);
  const size_t lens_block_stride = (size_t)5 * MAX_GROUPS * MAX_ALPHA;
  const size_t lens_final_off = (size_t)4 * MAX_GROUPS * MAX_ALPHA;
  L.launch("k6_pack_header", k6_pack_header, dim3(nb), dim3(PH_NT), nb, (const uint32_t*)H.meta, d_inuse, d_crc,
           d_origptr, (const uint8_t*)H.selmtf, (const uint8_t*)H.lens, lens_block_stride, lens_final_off,
           (const uint64_t*)d_blockbit, words);
  L.launch("k6_pack_symbols", k6_pack_symbols, dim3((max_groups_per_block + PS_NT - 1) / PS_NT, nb), dim3(PS_NT), d_sym,
           d_desc, d_mtf_count, (const uint32_t*)H.meta, (const uint8_t*)H.sel, (const uint32_t*)H.codes,
           (const uint32_t*)H.gbits, (const uint64_t*)d_blockbit, words);
}

void launch_bit_append(Launcher& L, uint8_t* d_dst, uint64_t dst_bit, const uint8_t* d_src, uint64_t nbits) {
  if (nbits == 0) return;
  uint64_t ndw = ((dst_bit & 31) + nbits + 31) >> 5;
  uint64_t want = (ndw + 255) / 256;
  unsigned grid = (unsigned)(want < 148ull * 16 ? want : 148ull * 16);
  L.launch("k7_bit_append", k7_bit_append, dim3(grid), dim3(256), reinterpret_cast<uint32_t*>(d_dst), dst_bit,
           reinterpret_cast<const uint32_t*>(d_src), nbits);
}

void launch_write_header(Launcher& L, int level, uint8_t* d_out) {
  L.launch("k6_write_stream_header", k6_write_stream_header, dim3(1), dim3(1), level, reinterpret_cast<uint32_t*>(d_out));
}

void launch_write_trailer(Launcher& L, uint8_t* d_out, uint64_t at_bit, uint32_t combined_crc) {
  L.launch("k6_write_trailer", k6_write_trailer, dim3(1), dim3(1), reinterpret_cast<uint32_t*>(d_out), at_bit,
           (const uint64_t*)nullptr, combined_crc, (const uint32_t*)nullptr);
}

void launch_combine_crc(Launcher& L, const uint32_t* d_crc, uint32_t n, uint32_t* d_combined) {
  L.launch("k6_combine_crc", k6_combine_crc, dim3(1), dim3(1), d_crc, n, d_combined);
}

void launch_write_trailer_dev(Launcher& L, uint8_t* d_out, const uint64_t* d_bitcursor, const uint32_t* d_combined) {
  L.launch("k6_write_trailer", k6_write_trailer, dim3(1), dim3(1), reinterpret_cast<uint32_t*>(d_out), (uint64_t)0,
           d_bitcursor, (uint32_t)0, d_combined);
}

}  // namespace bzb
