// dec_emu.cpp — host emulation of the GPU decoder's kernel bodies (TEST INFRASTRUCTURE, never linked into
// libbzb200.so).  Compiles rust-compression_b200/csrc/decoder.cu with -DBZB_EMU: every kernel "launch" becomes a loop
// over the same per-thread bodies (dec_core.cuh) and device memory becomes malloc'ed memory, so tests/test_dec_emu.py
// can compare the decoder's algorithm — candidate scan, chain validation, Huffman/MTF decode, list-ranking inverse
// BWT, RLE1 undo, error kinds — with the restated reference decoder (oracle/bz2_decoder_oracle.cpp) without a GPU.
#define BZB_EMU 1
#include "../../rust-compression_b200/csrc/decoder.cu"

namespace {

struct EmuMem : bzb::DecMem {
  void* p[bzb::DS_NSLOTS] = {};
  size_t cap[bzb::DS_NSLOTS] = {};
  ~EmuMem() override {
    for (auto q : p) free(q);
  }
  void* buf(int s, size_t bytes) override {
    if (bytes == 0) bytes = 16;
    if (cap[s] < bytes) {
      free(p[s]);
      p[s] = malloc(bytes);
      cap[s] = bytes;
      memset(p[s], 0xCD, bytes);  // device memory is not zeroed either
    }
    return p[s];
  }
  int fill(void* q, int byte, size_t bytes) override {
    memset(q, byte, bytes);
    return 0;
  }
  int to_host(void* dst, const void* src, size_t bytes) override {
    memcpy(dst, src, bytes);
    return 0;
  }
  int to_dev(void* dst, const void* src, size_t bytes) override {
    memcpy(dst, src, bytes);
    return 0;
  }
  int crc_blocks(const uint8_t* d, const uint64_t* off, uint32_t nb, uint32_t* crc) override {
    static uint32_t tab[256];
    if (!tab[1])
      for (uint32_t i = 0; i < 256; ++i) {
        uint32_t v = i << 24;
        for (int k = 0; k < 8; ++k) v = (v & 0x80000000u) ? (v << 1) ^ 0x04C11DB7u : (v << 1);
        tab[i] = v;
      }
    for (uint32_t b = 0; b < nb; ++b) {
      uint32_t r = 0xFFFFFFFFu;
      for (uint64_t i = off[b]; i < off[b + 1]; ++i) r = tab[((r >> 24) ^ d[i]) & 0xFF] ^ (r << 8);
      crc[b] = ~r;
    }
    return 0;
  }
  int check() override { return 0; }
  std::string err() override { return ""; }
};

}  // namespace

extern "C" {

// Returns dec_run's code; *bz_error = BZip2Error ordinal + 1 or 0; *out is malloc'ed (free with emu_free).
// info[0..7] = streams, blocks, candidates, batches, launches, retried(0/1), 0, 0.
int emu_decode(const uint8_t* in, size_t n, size_t first_cap, size_t batch_bytes, uint32_t flags, uint8_t** out,
               size_t* out_n, uint32_t* bz_error, uint64_t* info) {
  EmuMem M;
  bzb::Launcher L;
  bzb::DecResult R;
  size_t cap = first_cap;
  uint8_t* o = (uint8_t*)malloc(cap + 64);
  uint8_t* padded = (uint8_t*)malloc(n + 4);  // exact-size copy: reads beyond n would be caught by ASan builds
  if (n) memcpy(padded, in, n);
  int rc = bzb::dec_run(L, M, padded, n, o, cap, batch_bytes, flags, &R);
  int retried = 0;
  if (rc == 0 && R.too_small) {
    free(o);
    cap = R.needed;
    o = (uint8_t*)malloc(cap + 64);
    rc = bzb::dec_run(L, M, padded, n, o, cap, batch_bytes, flags, &R);
    retried = 1;
  }
  free(padded);
  *out = o;
  *out_n = R.out_n;
  *bz_error = R.bz_error;
  if (info) {
    info[0] = R.streams;
    info[1] = R.blocks;
    info[2] = R.candidates;
    info[3] = R.batches;
    info[4] = L.launches;
    info[5] = retried;
    info[6] = R.syms;
    info[7] = R.pre_rle;
  }
  return rc;
}
void emu_free(uint8_t* p) { free(p); }
}
