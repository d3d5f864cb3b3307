// decoder.cu — block-parallel bzip2 decoder: kernels (thin wrappers around the bodies in dec_core.cuh) and the
// stage orchestration with the host-side chain validation.  SURVEY.md §8(f).1; replaces BZip2Decoder
// (/root/reference/src/bzip2/decoder.rs:163-581) for whole buffers.
//
// Compiled twice: by nvcc into libbzb200.so (kernels), and by g++ with -DBZB_EMU into tests/cpp/libdecemu.so, where
// every "launch" is a loop over the same bodies — test infrastructure for checking the algorithm without a GPU.
#include "decoder.h"

#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#ifndef BZB_EMU
#include "kernels.h"
#endif
#include "bz_rand_table.h"
#include "dec_core.cuh"

namespace bzb {

using namespace dec;

// ------------------------------------------------------------------------------------------------ launches
#ifndef BZB_EMU

__global__ void __launch_bounds__(256) d1_scan(const uint8_t* __restrict__ in, uint64_t n, uint64_t nwords,
                                               uint64_t* cand, uint32_t* cand_count, uint32_t cap) {
  const uint64_t x = (uint64_t)blockIdx.x * 256u + threadIdx.x;
  if (x < nwords) d1_scan_body(x, in, n, cand, cand_count, cap);
}

__global__ void __launch_bounds__(32) d2_decode(const uint8_t* __restrict__ in, uint64_t n, const uint64_t* cand,
                                                uint32_t cap, uint64_t stride, uint32_t* occbuf, uint8_t* selbuf,
                                                uint32_t* cftab, CandInfo* infos) {
  __shared__ D2Scratch s;
  d2_decode_body(blockIdx.x, threadIdx.x, &s, in, n, cand, cap, stride, occbuf, selbuf, cftab, infos);
}

__global__ void __launch_bounds__(32) d2_huff(const uint8_t* __restrict__ in, uint64_t n, const uint64_t* cand,
                                              uint32_t cap, uint64_t symstride, uint16_t* symbuf, uint8_t* selbuf,
                                              uint8_t* mtf0buf, CandInfo* infos) {
  __shared__ D2Scratch s;
  if (threadIdx.x == 0) d2_huff_body(blockIdx.x, &s, in, n, cand, cap, symstride, symbuf, selbuf, mtf0buf, infos);
}

__global__ void __launch_bounds__(32) d2_mtf_a(const CandInfo* __restrict__ infos, uint64_t symstride,
                                               const uint16_t* __restrict__ symbuf, uint32_t chunks_pitch, uint8_t* Pbuf,
                                               uint8_t* permbuf, uint32_t* cntpbuf, ChunkMeta* metabuf) {
  __shared__ uint32_t cnt[256 * 32];  // entry k of thread t at [k * 32 + t]: conflict-free
  __shared__ uint8_t pl[256 * 32];
  d2_mtf_a_body(blockIdx.x * 32u + threadIdx.x, blockIdx.y, infos, symstride, symbuf, chunks_pitch, Pbuf, permbuf,
                cntpbuf, metabuf, pl + threadIdx.x, cnt + threadIdx.x, 32);
}

__global__ void __launch_bounds__(32) d2_mtf_b(CandInfo* infos, uint32_t cap, uint64_t symstride,
                                               const uint16_t* __restrict__ symbuf, const uint8_t* __restrict__ mtf0buf,
                                               uint32_t chunks_pitch, const uint8_t* __restrict__ permbuf,
                                               const uint32_t* __restrict__ cntpbuf,
                                               const ChunkMeta* __restrict__ metabuf, uint8_t* initlbuf,
                                               uint32_t* basebuf, uint32_t* coffbuf, uint32_t* cd0buf, uint32_t* cftab) {
  __shared__ MtfBScratch s;
  d2_mtf_b_body(blockIdx.x, threadIdx.x, &s, infos, cap, symstride, symbuf, mtf0buf, chunks_pitch, permbuf, cntpbuf,
                metabuf, initlbuf, basebuf, coffbuf, cd0buf, cftab);
}

__global__ void __launch_bounds__(32) d2_mtf_c(const CandInfo* __restrict__ infos, uint64_t symstride,
                                               const uint16_t* __restrict__ symbuf, const uint8_t* __restrict__ Pbuf,
                                               uint32_t chunks_pitch, const uint8_t* __restrict__ initlbuf,
                                               const uint32_t* __restrict__ basebuf, const uint32_t* __restrict__ coffbuf,
                                               const uint32_t* __restrict__ cd0buf, uint64_t stride, uint32_t* occbuf) {
  __shared__ uint32_t loc[256 * 32];
  __shared__ uint8_t il[256 * 32];
  d2_mtf_c_body(blockIdx.x * 32u + threadIdx.x, blockIdx.y, infos, symstride, symbuf, Pbuf, chunks_pitch, initlbuf,
                basebuf, coffbuf, cd0buf, stride, occbuf, il + threadIdx.x, loc + threadIdx.x, 32);
}

__global__ void __launch_bounds__(256) d3_scatter(const CandInfo* __restrict__ infos, uint64_t stride,
                                                  const uint32_t* __restrict__ occbuf,
                                                  const uint32_t* __restrict__ cftab, uint32_t* Vbuf) {
  d3_scatter_body(blockIdx.x * 256u + threadIdx.x, blockIdx.y, infos, stride, occbuf, cftab, Vbuf);
}

__global__ void __launch_bounds__(128) d4_walk_a(const CandInfo* __restrict__ infos, uint64_t stride,
                                                 const uint32_t* __restrict__ Vbuf, uint32_t segs_pitch,
                                                 uint32_t* seg_len, uint32_t* seg_next, uint32_t* seg_resume,
                                                 uint8_t* Tbuf) {
  d4_walk_a_body(blockIdx.x * 128u + threadIdx.x, blockIdx.y, infos, stride, Vbuf, segs_pitch, seg_len, seg_next,
                 seg_resume, Tbuf);
}

__global__ void __launch_bounds__(64) d4_schedule(uint32_t nc, CandInfo* infos, uint32_t segs_pitch,
                                                  const uint32_t* seg_len, const uint32_t* seg_next, uint32_t* seg_off) {
  const uint32_t y = blockIdx.x * 64u + threadIdx.x;
  if (y < nc) d4_schedule_body(y, infos, segs_pitch, seg_len, seg_next, seg_off);
}

__global__ void __launch_bounds__(128) d4_walk_c(const CandInfo* __restrict__ infos, uint64_t stride,
                                                 const uint32_t* __restrict__ Vbuf, uint32_t segs_pitch,
                                                 const uint32_t* __restrict__ seg_len,
                                                 const uint32_t* __restrict__ seg_off,
                                                 const uint32_t* __restrict__ seg_resume,
                                                 const uint8_t* __restrict__ Tbuf, uint8_t* Wbuf) {
  d4_walk_c_body(blockIdx.x * 128u + threadIdx.x, blockIdx.y, infos, stride, Vbuf, segs_pitch, seg_len, seg_off,
                 seg_resume, Tbuf, Wbuf);
}

__global__ void __launch_bounds__(256) d4_derand(const CandInfo* __restrict__ infos, uint64_t stride,
                                                 const uint32_t* __restrict__ cum, uint32_t period, uint8_t* Wbuf) {
  d4_derand_body(blockIdx.x * 256u + threadIdx.x, blockIdx.y, infos, stride, cum, period, Wbuf);
}

__global__ void __launch_bounds__(128) d5_count(const CandInfo* __restrict__ infos, uint64_t stride,
                                                const uint8_t* __restrict__ Wbuf, uint32_t chunks_pitch,
                                                uint32_t* rle_map) {
  d5_count_body(blockIdx.x * 128u + threadIdx.x, blockIdx.y, infos, stride, Wbuf, chunks_pitch, rle_map);
}

__global__ void __launch_bounds__(64) d5_compose(uint32_t nc, CandInfo* infos, uint32_t chunks_pitch,
                                                 const uint32_t* __restrict__ rle_map, uint32_t* chunk_entry,
                                                 uint64_t* chunk_off) {
  const uint32_t y = blockIdx.x * 64u + threadIdx.x;
  if (y < nc) d5_compose_body(y, infos, chunks_pitch, rle_map, chunk_entry, chunk_off);
}

__global__ void __launch_bounds__(128) d5_expand(const CandInfo* __restrict__ infos, uint64_t stride,
                                                 const uint8_t* __restrict__ Wbuf, uint32_t chunks_pitch,
                                                 const uint32_t* __restrict__ chunk_entry,
                                                 const uint64_t* __restrict__ chunk_off,
                                                 const uint64_t* __restrict__ out_off, uint8_t* out) {
  d5_expand_body(blockIdx.x * 128u + threadIdx.x, blockIdx.y, infos, stride, Wbuf, chunks_pitch, chunk_entry, chunk_off,
                 out_off, out);
}

#define GRID1(n, t) dim3((unsigned)(((n) + (t) - 1) / (t)))
#define GRID2(n, t, y) dim3((unsigned)(((n) + (t) - 1) / (t)), (unsigned)(y))

static void run_d1(Launcher& L, const uint8_t* in, uint64_t n, uint64_t* cand, uint32_t* count, uint32_t cap) {
  const uint64_t nwords = (n + 3) / 4;
  L.launch("d1_scan", d1_scan, GRID1(nwords, 256), dim3(256), in, n, nwords, cand, count, cap);
}
static void run_d2(Launcher& L, uint32_t nc, const uint8_t* in, uint64_t n, const uint64_t* cand, uint32_t cap,
                   uint64_t stride, uint32_t* occ, uint8_t* sel, uint32_t* cftab, CandInfo* infos) {
  L.launch("d2_decode", d2_decode, dim3(nc), dim3(32), in, n, cand, cap, stride, occ, sel, cftab, infos);
}
struct SplitBufs {
  uint64_t symstride;
  uint32_t chunks_pitch;
  uint16_t* sym;
  uint8_t* P;
  uint8_t* mtf0;
  uint8_t* perm;
  uint32_t* cntp;
  ChunkMeta* meta;
  uint8_t* initl;
  uint32_t* base;
  uint32_t* coff;
  uint32_t* cd0;
};
static void run_d2_split(Launcher& L, uint32_t nc, const uint8_t* in, uint64_t n, const uint64_t* cand, uint32_t cap,
                         uint64_t stride, uint32_t* occ, uint8_t* sel, uint32_t* cftab, CandInfo* infos, SplitBufs& B) {
  L.launch("d2_huff", d2_huff, dim3(nc), dim3(32), in, n, cand, cap, B.symstride, B.sym, sel, B.mtf0, infos);
  L.launch("d2_mtf_a", d2_mtf_a, GRID2(B.chunks_pitch, 32, nc), dim3(32), infos, B.symstride, B.sym, B.chunks_pitch, B.P,
           B.perm, B.cntp, B.meta);
  L.launch("d2_mtf_b", d2_mtf_b, dim3(nc), dim3(32), infos, cap, B.symstride, B.sym, B.mtf0, B.chunks_pitch, B.perm,
           B.cntp, B.meta, B.initl, B.base, B.coff, B.cd0, cftab);
  L.launch("d2_mtf_c", d2_mtf_c, GRID2(B.chunks_pitch, 32, nc), dim3(32), infos, B.symstride, B.sym, B.P, B.chunks_pitch,
           B.initl, B.base, B.coff, B.cd0, stride, occ);
}
static void run_d3(Launcher& L, uint32_t nc, uint32_t nmax, const CandInfo* infos, uint64_t stride,
                   const uint32_t* occ, const uint32_t* cftab, uint32_t* V) {
  L.launch("d3_scatter", d3_scatter, GRID2(nmax, 256, nc), dim3(256), infos, stride, occ, cftab, V);
}
static void run_d4a(Launcher& L, uint32_t nc, uint32_t segs, const CandInfo* infos, uint64_t stride, const uint32_t* V,
                    uint32_t pitch, uint32_t* seg_len, uint32_t* seg_next, uint32_t* seg_resume, uint8_t* T) {
  L.launch("d4_walk_a", d4_walk_a, GRID2(segs, 128, nc), dim3(128), infos, stride, V, pitch, seg_len, seg_next,
           seg_resume, T);
}
static void run_d4s(Launcher& L, uint32_t nc, CandInfo* infos, uint32_t pitch, const uint32_t* seg_len,
                    const uint32_t* seg_next, uint32_t* seg_off) {
  L.launch("d4_schedule", d4_schedule, GRID1(nc, 64), dim3(64), nc, infos, pitch, seg_len, seg_next, seg_off);
}
static void run_d4c(Launcher& L, uint32_t nc, uint32_t segs, const CandInfo* infos, uint64_t stride, const uint32_t* V,
                    uint32_t pitch, const uint32_t* seg_len, const uint32_t* seg_off, const uint32_t* seg_resume,
                    const uint8_t* T, uint8_t* W) {
  L.launch("d4_walk_c", d4_walk_c, GRID2(segs, 128, nc), dim3(128), infos, stride, V, pitch, seg_len, seg_off,
           seg_resume, T, W);
}
static void run_d4r(Launcher& L, uint32_t nc, uint32_t toggles, const CandInfo* infos, uint64_t stride,
                    const uint32_t* cum, uint8_t* W) {
  L.launch("d4_derand", d4_derand, GRID2(toggles, 256, nc), dim3(256), infos, stride, cum, (uint32_t)BZ_RAND_PERIOD, W);
}
static void run_d5a(Launcher& L, uint32_t nc, uint32_t chunks, const CandInfo* infos, uint64_t stride, const uint8_t* W,
                    uint32_t pitch, uint32_t* rle_map) {
  L.launch("d5_count", d5_count, GRID2(chunks, 128, nc), dim3(128), infos, stride, W, pitch, rle_map);
}
static void run_d5b(Launcher& L, uint32_t nc, CandInfo* infos, uint32_t pitch, const uint32_t* rle_map,
                    uint32_t* chunk_entry, uint64_t* chunk_off) {
  L.launch("d5_compose", d5_compose, GRID1(nc, 64), dim3(64), nc, infos, pitch, rle_map, chunk_entry, chunk_off);
}
static void run_d5c(Launcher& L, uint32_t nc, uint32_t chunks, const CandInfo* infos, uint64_t stride, const uint8_t* W,
                    uint32_t pitch, const uint32_t* chunk_entry, const uint64_t* chunk_off, const uint64_t* out_off,
                    uint8_t* out) {
  L.launch("d5_expand", d5_expand, GRID2(chunks, 128, nc), dim3(128), infos, stride, W, pitch, chunk_entry, chunk_off,
           out_off, out);
}

#else  // ---------------------------------------------------------------- host emulation: one loop per kernel

static void run_d1(Launcher& L, const uint8_t* in, uint64_t n, uint64_t* cand, uint32_t* count, uint32_t cap) {
  ++L.launches;
  for (uint64_t x = 0; x < (n + 3) / 4; ++x) d1_scan_body(x, in, n, cand, count, cap);
}
static void run_d2(Launcher& L, uint32_t nc, const uint8_t* in, uint64_t n, const uint64_t* cand, uint32_t cap,
                   uint64_t stride, uint32_t* occ, uint8_t* sel, uint32_t* cftab, CandInfo* infos) {
  ++L.launches;
  D2Scratch* s = new D2Scratch();
  for (uint32_t c = 0; c < nc; ++c) {
    memset(s, 0xA5, sizeof(*s));  // shared memory is not zeroed between CTAs either
    d2_decode_body(c, 0, s, in, n, cand, cap, stride, occ, sel, cftab, infos);
  }
  delete s;
}
struct SplitBufs {
  uint64_t symstride;
  uint32_t chunks_pitch;
  uint16_t* sym;
  uint8_t* P;
  uint8_t* mtf0;
  uint8_t* perm;
  uint32_t* cntp;
  ChunkMeta* meta;
  uint8_t* initl;
  uint32_t* base;
  uint32_t* coff;
  uint32_t* cd0;
};
static void run_d2_split(Launcher& L, uint32_t nc, const uint8_t* in, uint64_t n, const uint64_t* cand, uint32_t cap,
                         uint64_t stride, uint32_t* occ, uint8_t* sel, uint32_t* cftab, CandInfo* infos, SplitBufs& B) {
  L.launches += 4;
  D2Scratch* s = new D2Scratch();
  for (uint32_t c = 0; c < nc; ++c) {
    memset(s, 0xA5, sizeof(*s));
    d2_huff_body(c, s, in, n, cand, cap, B.symstride, B.sym, sel, B.mtf0, infos);
  }
  delete s;
  const uint32_t gx = (B.chunks_pitch + 63) / 64 * 64;
  uint8_t sc8[256];
  uint32_t sc32[256];
  for (uint32_t y = 0; y < nc; ++y)
    for (uint32_t x = 0; x < gx; ++x) {
      memset(sc8, 0xA5, sizeof(sc8));
      memset(sc32, 0xA5, sizeof(sc32));
      d2_mtf_a_body(x, y, infos, B.symstride, B.sym, B.chunks_pitch, B.P, B.perm, B.cntp, B.meta, sc8, sc32, 1);
    }
  MtfBScratch* sb = new MtfBScratch();
  for (uint32_t y = 0; y < nc; ++y) {
    memset(sb, 0xA5, sizeof(*sb));
    d2_mtf_b_body(y, 0, sb, infos, cap, B.symstride, B.sym, B.mtf0, B.chunks_pitch, B.perm, B.cntp, B.meta, B.initl,
                  B.base, B.coff, B.cd0, cftab);
  }
  delete sb;
  for (uint32_t y = 0; y < nc; ++y)
    for (uint32_t x = 0; x < gx; ++x) {
      memset(sc8, 0xA5, sizeof(sc8));
      memset(sc32, 0xA5, sizeof(sc32));
      d2_mtf_c_body(x, y, infos, B.symstride, B.sym, B.P, B.chunks_pitch, B.initl, B.base, B.coff, B.cd0, stride, occ,
                    sc8, sc32, 1);
    }
}
static void run_d3(Launcher& L, uint32_t nc, uint32_t nmax, const CandInfo* infos, uint64_t stride,
                   const uint32_t* occ, const uint32_t* cftab, uint32_t* V) {
  ++L.launches;
  for (uint32_t y = 0; y < nc; ++y)
    for (uint32_t x = 0; x < (nmax + 255) / 256 * 256; ++x) d3_scatter_body(x, y, infos, stride, occ, cftab, V);
}
static void run_d4a(Launcher& L, uint32_t nc, uint32_t segs, const CandInfo* infos, uint64_t stride, const uint32_t* V,
                    uint32_t pitch, uint32_t* seg_len, uint32_t* seg_next, uint32_t* seg_resume, uint8_t* T) {
  ++L.launches;
  for (uint32_t y = 0; y < nc; ++y)
    for (uint32_t x = 0; x < (segs + 127) / 128 * 128; ++x)
      d4_walk_a_body(x, y, infos, stride, V, pitch, seg_len, seg_next, seg_resume, T);
}
static void run_d4s(Launcher& L, uint32_t nc, CandInfo* infos, uint32_t pitch, const uint32_t* seg_len,
                    const uint32_t* seg_next, uint32_t* seg_off) {
  ++L.launches;
  for (uint32_t y = 0; y < nc; ++y) d4_schedule_body(y, infos, pitch, seg_len, seg_next, seg_off);
}
static void run_d4c(Launcher& L, uint32_t nc, uint32_t segs, const CandInfo* infos, uint64_t stride, const uint32_t* V,
                    uint32_t pitch, const uint32_t* seg_len, const uint32_t* seg_off, const uint32_t* seg_resume,
                    const uint8_t* T, uint8_t* W) {
  ++L.launches;
  for (uint32_t y = 0; y < nc; ++y)
    for (uint32_t x = 0; x < (segs + 127) / 128 * 128; ++x)
      d4_walk_c_body(x, y, infos, stride, V, pitch, seg_len, seg_off, seg_resume, T, W);
}
static void run_d4r(Launcher& L, uint32_t nc, uint32_t toggles, const CandInfo* infos, uint64_t stride,
                    const uint32_t* cum, uint8_t* W) {
  ++L.launches;
  for (uint32_t y = 0; y < nc; ++y)
    for (uint32_t x = 0; x < (toggles + 255) / 256 * 256; ++x)
      d4_derand_body(x, y, infos, stride, cum, (uint32_t)BZ_RAND_PERIOD, W);
}
static void run_d5a(Launcher& L, uint32_t nc, uint32_t chunks, const CandInfo* infos, uint64_t stride, const uint8_t* W,
                    uint32_t pitch, uint32_t* rle_map) {
  ++L.launches;
  for (uint32_t y = 0; y < nc; ++y)
    for (uint32_t x = 0; x < (chunks + 127) / 128 * 128; ++x) d5_count_body(x, y, infos, stride, W, pitch, rle_map);
}
static void run_d5b(Launcher& L, uint32_t nc, CandInfo* infos, uint32_t pitch, const uint32_t* rle_map,
                    uint32_t* chunk_entry, uint64_t* chunk_off) {
  ++L.launches;
  for (uint32_t y = 0; y < nc; ++y) d5_compose_body(y, infos, pitch, rle_map, chunk_entry, chunk_off);
}
static void run_d5c(Launcher& L, uint32_t nc, uint32_t chunks, const CandInfo* infos, uint64_t stride, const uint8_t* W,
                    uint32_t pitch, const uint32_t* chunk_entry, const uint64_t* chunk_off, const uint64_t* out_off,
                    uint8_t* out) {
  ++L.launches;
  for (uint32_t y = 0; y < nc; ++y)
    for (uint32_t x = 0; x < (chunks + 127) / 128 * 128; ++x)
      d5_expand_body(x, y, infos, stride, W, pitch, chunk_entry, chunk_off, out_off, out);
}

#endif

// ------------------------------------------------------------------------------------------------ chain validation
namespace {

// What a sequential parse (decoder.rs:163-525) would have visited: stream header, then magics back to back.  The
// scan found every magic in the buffer; the chain keeps the candidates that start exactly where their predecessor
// ended and turns everything else into the error the reference reports at that point.
struct Chain {
  const std::vector<uint64_t>& cand;  // sorted; bit 63 = end-of-stream kind
  uint64_t n, nbits;
  uint8_t head4[4];
  uint32_t head_n;

  uint64_t pos = 0;
  uint32_t level = 0, stream_no = 1, block_no = 0, combined = 0;
  uint64_t out_total = 0;   // bytes of the accepted blocks
  uint32_t err = 0;         // terminal error (0 with done == true: clean end)
  bool done = false;
  uint32_t streams = 0, blocks = 0;
  uint64_t syms = 0, pre_rle = 0;

  Chain(const std::vector<uint64_t>& c, uint64_t n_, const uint8_t* h4, uint32_t hn) : cand(c), n(n_), nbits(n_ * 8) {
    memcpy(head4, h4, 4);
    head_n = hn;
  }

  // decoder.rs:171-196.  The reference reads 'B','Z','h' with check_u8 and DISCARDS the comparison
  // (`let _ = Self::check_u8(..).map_err(|_| magic_err)?`): only a missing byte is an error; the level byte is checked.
  bool stream_header(const uint8_t* b, uint32_t have) {
    const uint32_t magic_err = stream_no == 1 ? E_MAGIC_FIRST : E_MAGIC;
    if (have < 3) return fail(magic_err);
    if (have < 4) return fail(E_EOF);
    if (b[3] < '1' || b[3] > '9') return fail(magic_err);
    level = b[3] - '0';
    pos += 32;
    ++streams;
    return true;
  }
  bool fail(uint32_t e) {
    err = e;
    done = true;
    return false;
  }
  size_t find(uint64_t p) const {
    size_t lo = 0, hi = cand.size();
    while (lo < hi) {
      const size_t m = (lo + hi) / 2;
      if ((cand[m] & ~KIND_END) < p) lo = m + 1; else hi = m;
    }
    return (lo < cand.size() && (cand[lo] & ~KIND_END) == p) ? lo : (size_t)-1;
  }

  // Walks while the next candidate lies in [c0, c1) (the batch whose infos are given).  accepted: (index inside the
  // batch, output offset) of every block taken.  Returns DONE (verdict reached), LATER (the chain needs a later
  // batch) or MISS: no candidate starts at `pos` — the scan lists only complete 48-bit magics, while the reference
  // looks at the FIRST byte alone (0x31 / 0x17, decoder.rs:206-210,494) and skips the other five without comparing
  // them; the caller reads that byte and either injects a candidate there or reports the reference's error.
  enum { DONE = 0, LATER = 1, MISS = 2 };
  int advance(size_t c0, size_t c1, const CandInfo* infos, std::vector<std::pair<uint32_t, uint64_t>>& accepted) {
    while (!done) {
      if (block_no == 0 && level == 0) {
        if (!stream_header(head4, head_n)) break;
      }
      const size_t ci = find(pos);
      if (ci == (size_t)-1) return MISS;
      if (ci >= c1) return LATER;
      if (ci < c0) {  // cannot happen: positions only grow
        fail(E_UNEXPECTED);
        break;
      }
      const CandInfo& I = infos[ci - c0];
      if (I.kind == 0) {
        block_no += 1;
        uint32_t e = 0;
        if (I.err && I.err_early) e = I.err;
        else if (I.orig_pos > 10u + 100000u * level) e = E_DATA;  // decoder.rs:238
        else if (I.err) e = I.err;
        else if (I.need_max > 100000u * level) e = E_DATA;       // decoder.rs:399,427
        else if (I.rle_dangling) e = E_DATA;  // four equal bytes with no count at the block end (never produced by
                                              // an encoder; the reference would read past the block here)
        if (e) {
          fail(e);
          break;
        }
        accepted.push_back({(uint32_t)(ci - c0), out_total});
        out_total += I.rle_len;
        combined = ((combined << 1) | (combined >> 31)) ^ I.stored_crc;
        pos = I.end_bit;
        ++blocks;
        syms += I.nsym;
        pre_rle += I.nblock;
      } else {
        if (I.err) {
          fail(I.err);
          break;
        }
        if (I.stored_crc != combined) {
          fail(E_DATA);
          break;
        }
        pos = (I.end_bit + 7) & ~7ull;  // skip_to_next_byte
        if (nbits - pos >= 8) {          // another stream follows (decoder.rs:510-517)
          block_no = 0;
          combined = 0;
          stream_no += 1;
          level = 0;
          if (!stream_header(I.tail, I.tail_n)) break;
        } else {
          done = true;
        }
      }
    }
    return DONE;
  }
};

template <class T>
T* slot(DecMem& M, int s, size_t count) {
  return reinterpret_cast<T*>(M.buf(s, count * sizeof(T)));
}

}  // namespace

#define DTRY(x)            \
  do {                     \
    if ((x) != 0) return -2; /* BZB200_E_CUDA */ \
  } while (0)

int dec_run(Launcher& L, DecMem& M, const uint8_t* d_in, uint64_t n, uint8_t* d_out, uint64_t cap_out,
            uint64_t batch_bytes, uint32_t flags, DecResult* res) {
  *res = DecResult();
  if (n == 0) {  // the very first read fails (decoder.rs:176-181)
    res->bz_error = E_MAGIC_FIRST;
    return 0;
  }
  // ---- D1: candidates
  std::vector<uint64_t> cand;
  {
    uint32_t cap = 1u << 16;
    for (;;) {
      uint64_t* d_cand = slot<uint64_t>(M, DS_CAND, cap);
      uint32_t* d_count = slot<uint32_t>(M, DS_COUNT, 4);
      if (!d_cand || !d_count) return -2;
      DTRY(M.fill(d_count, 0, 16));
      run_d1(L, d_in, n, d_cand, d_count, cap);
      DTRY(M.check());
      uint32_t cnt = 0;
      DTRY(M.to_host(&cnt, d_count, 4));
      if (cnt > cap) {
        cap = cnt + 16;
        continue;
      }
      cand.resize(cnt);
      if (cnt) DTRY(M.to_host(cand.data(), d_cand, (size_t)cnt * 8));
      break;
    }
    std::sort(cand.begin(), cand.end(),
              [](uint64_t a, uint64_t b) { return (a & ~KIND_END) < (b & ~KIND_END); });
  }
  res->candidates = (uint32_t)cand.size();
  // ---- stream headers: the first one, and whatever follows each end-of-stream magic, bound the block size
  uint8_t head4[4] = {0, 0, 0, 0};
  const uint32_t head_n = (uint32_t)std::min<uint64_t>(4, n);
  DTRY(M.to_host(head4, d_in, head_n));
  uint32_t maxlevel = (head_n == 4 && head4[3] >= '1' && head4[3] <= '9') ? head4[3] - '0' : 0;
  {
    uint32_t ends = 0;
    for (size_t i = 0; i < cand.size(); ++i) {
      if (!(cand[i] & KIND_END)) continue;
      if (++ends > 64) {  // many streams: stop probing, size for the largest block
        maxlevel = 9;
        break;
      }
      const uint64_t after = (((cand[i] & ~KIND_END) + 80 + 7) >> 3) + 3;  // the level byte of a following header
      if (after < n) {
        uint8_t b = 0;
        DTRY(M.to_host(&b, d_in + after, 1));
        if (b >= '1' && b <= '9') maxlevel = std::max<uint32_t>(maxlevel, b - '0');
      }
    }
  }
  if (maxlevel == 0) {  // bad first header: the chain alone yields the error
    Chain chain(cand, n, head4, head_n);
    std::vector<std::pair<uint32_t, uint64_t>> none;
    chain.advance(0, 0, nullptr, none);
    if (!chain.done) chain.fail(E_UNEXPECTED);
    res->bz_error = chain.err;
    res->streams = chain.streams;
    return 0;
  }
  const uint64_t nbits = n * 8;
  // The pass below runs again from the start in one rare case: an end-of-stream magic with damaged bytes 2..6 (found
  // only by the chain, see Chain::advance) in front of a stream of a higher level than the buffers were sized for.
  for (int attempt = 0;; ++attempt) {
    bool restart = false;
    Chain chain(cand, n, head4, head_n);
    std::vector<std::pair<uint32_t, uint64_t>> accepted;
    const uint32_t cap = 100000u * maxlevel;
    const uint64_t stride = ((uint64_t)cap + 63) & ~63ull;
    const uint32_t segs_pitch = d4_nseg0(cap) + 1;
    const uint32_t chunks_pitch = d5_nchunks(cap);
    const bool split = (flags & DEC_SPLIT_D2) != 0;
    SplitBufs SB;
    SB.symstride = ((uint64_t)cap + 4 + 63) & ~63ull;
    SB.chunks_pitch = d2_nchunks(cap + 2);
    const uint64_t per_cand = stride * 9 + MAX_SEL + 257 * 4 + (uint64_t)segs_pitch * (16 + SEG_KEEP) +
                              (uint64_t)chunks_pitch * 32 + sizeof(CandInfo) + 64 +
                              (split ? SB.symstride * 3 + 256 + (uint64_t)SB.chunks_pitch * (256 * 10 + 32 + 8) : 0);
    size_t batch = (size_t)std::max<uint64_t>(1, batch_bytes / per_cand);
    batch = std::min<size_t>(batch, 32768);
    batch = std::max<size_t>(1, std::min<size_t>(batch, cand.size()));

    uint64_t* d_cand = slot<uint64_t>(M, DS_CAND, cand.size() + 1);
    if (!d_cand) return -2;
    if (!cand.empty()) DTRY(M.to_dev(d_cand, cand.data(), cand.size() * 8));
    CandInfo* d_info = slot<CandInfo>(M, DS_INFO, batch);
    uint32_t* d_occ = slot<uint32_t>(M, DS_OCC, batch * stride);
    uint32_t* d_V = slot<uint32_t>(M, DS_V, batch * stride);
    uint8_t* d_W = slot<uint8_t>(M, DS_W, batch * stride);
    uint8_t* d_sel = slot<uint8_t>(M, DS_SEL, batch * (size_t)MAX_SEL);
    uint32_t* d_cftab = slot<uint32_t>(M, DS_CFTAB, batch * 257);
    uint32_t* d_seglen = slot<uint32_t>(M, DS_SEGLEN, batch * segs_pitch);
    uint32_t* d_segnext = slot<uint32_t>(M, DS_SEGNEXT, batch * segs_pitch);
    uint32_t* d_segoff = slot<uint32_t>(M, DS_SEGOFF, batch * segs_pitch);
    uint32_t* d_segres = slot<uint32_t>(M, DS_SEGRES, batch * segs_pitch);
    uint8_t* d_T = slot<uint8_t>(M, DS_T, batch * (size_t)segs_pitch * SEG_KEEP);
    uint32_t* d_rlemap = slot<uint32_t>(M, DS_RLEMAP, batch * (size_t)chunks_pitch * 5);
    uint32_t* d_chentry = slot<uint32_t>(M, DS_CHENTRY, batch * (size_t)chunks_pitch);
    uint64_t* d_choff = slot<uint64_t>(M, DS_CHOFF, batch * (size_t)chunks_pitch);
    uint64_t* d_outoff = slot<uint64_t>(M, DS_OUTOFF, batch);
    uint64_t* d_crcoff = slot<uint64_t>(M, DS_CRCOFF, batch + 1);
    uint32_t* d_crc = slot<uint32_t>(M, DS_CRC, batch);
    uint32_t* d_randcum = nullptr;  // uploaded when the first randomised block shows up
    if (split) {
      const size_t ch = batch * (size_t)SB.chunks_pitch;
      SB.sym = slot<uint16_t>(M, DS_SYM, batch * SB.symstride);
      SB.P = slot<uint8_t>(M, DS_P, batch * SB.symstride);
      SB.mtf0 = slot<uint8_t>(M, DS_MTF0, batch * 256);
      SB.perm = slot<uint8_t>(M, DS_PERM, ch * 256);
      SB.cntp = slot<uint32_t>(M, DS_CNTP, ch * 256);
      SB.meta = slot<ChunkMeta>(M, DS_CMETA, ch);
      SB.initl = slot<uint8_t>(M, DS_INITL, ch * 256);
      SB.base = slot<uint32_t>(M, DS_BASE, ch * 256);
      SB.coff = slot<uint32_t>(M, DS_COFF, ch);
      SB.cd0 = slot<uint32_t>(M, DS_CD0, ch);
      if (!SB.sym || !SB.P || !SB.mtf0 || !SB.perm || !SB.cntp || !SB.meta || !SB.initl || !SB.base || !SB.coff || !SB.cd0)
        return -2;
    }
    if (!d_info || !d_occ || !d_V || !d_W || !d_sel || !d_cftab || !d_seglen || !d_segnext || !d_segoff || !d_segres || !d_T ||
        !d_rlemap || !d_chentry || !d_choff || !d_outoff || !d_crcoff || !d_crc)
      return -2;

    std::vector<CandInfo> infos(batch);
    std::vector<uint64_t> h_outoff(batch), h_crcoff(batch + 1);
    std::vector<uint32_t> h_crc(batch);
    bool dry = false;         // the output buffer is too small: finish the chain for the exact size, write nothing
    uint32_t crc_err = 0;
    uint64_t crc_err_out = 0;
    res->batches = 0;

    size_t c0 = 0;
    while (!chain.done && !crc_err) {
      const size_t c1 = std::min(cand.size(), c0 + batch);
      const uint32_t nc = (uint32_t)(c1 - c0);
      uint32_t nmax = 0;
      if (nc) {
        ++res->batches;
        // ---- D2: header, tables, symbols, MTF, runs
        if (split) run_d2_split(L, nc, d_in, n, d_cand + c0, cap, stride, d_occ, d_sel, d_cftab, d_info, SB);
        else run_d2(L, nc, d_in, n, d_cand + c0, cap, stride, d_occ, d_sel, d_cftab, d_info);
        DTRY(M.check());
        DTRY(M.to_host(infos.data(), d_info, (size_t)nc * sizeof(CandInfo)));
        bool any_rand = false;
        for (uint32_t i = 0; i < nc; ++i)
          if (infos[i].kind == 0 && infos[i].err == 0) {
            nmax = std::max(nmax, infos[i].nblock);
            any_rand |= infos[i].randomised != 0;
          }
        if (nmax) {
          const uint32_t segs = d4_nseg0(nmax) + 1, chunks = d5_nchunks(nmax);
          // ---- D3/D4: inverse BWT
          run_d3(L, nc, nmax, d_info, stride, d_occ, d_cftab, d_V);
          DTRY(M.fill(d_segoff, 0xFF, (size_t)nc * segs_pitch * 4));
          run_d4a(L, nc, segs, d_info, stride, d_V, segs_pitch, d_seglen, d_segnext, d_segres, d_T);
          run_d4s(L, nc, d_info, segs_pitch, d_seglen, d_segnext, d_segoff);
          run_d4c(L, nc, segs, d_info, stride, d_V, segs_pitch, d_seglen, d_segoff, d_segres, d_T, d_W);
          if (any_rand) {  // blocks written by bzip2 <= 0.9.0 (decoder.rs:94-116,537-539)
            if (!d_randcum) {
              d_randcum = slot<uint32_t>(M, DS_RAND, 513);
              if (!d_randcum) return -2;
              DTRY(M.to_dev(d_randcum, BZ_RAND_CUM, sizeof(BZ_RAND_CUM)));
            }
            const uint32_t toggles = (nmax / BZ_RAND_PERIOD + 1) * 512u;
            run_d4r(L, nc, toggles, d_info, stride, d_randcum, d_W);
          }
          // ---- D5: RLE1 undo, sizes
          run_d5a(L, nc, chunks, d_info, stride, d_W, chunks_pitch, d_rlemap);
          run_d5b(L, nc, d_info, chunks_pitch, d_rlemap, d_chentry, d_choff);
          DTRY(M.check());
          DTRY(M.to_host(infos.data(), d_info, (size_t)nc * sizeof(CandInfo)));
        }
      }
      // ---- chain: which candidates a sequential parse visits, and where their bytes go
      accepted.clear();
      const int verdict = chain.advance(c0, c1, infos.data(), accepted);
      if (!accepted.empty()) {
        if (chain.out_total > cap_out) dry = true;
        if (!dry) {
          for (uint32_t i = 0; i < nc; ++i) h_outoff[i] = ~0ull;
          for (size_t k = 0; k < accepted.size(); ++k) {
            h_outoff[accepted[k].first] = accepted[k].second;
            h_crcoff[k] = accepted[k].second;
          }
          h_crcoff[accepted.size()] = chain.out_total;
          DTRY(M.to_dev(d_outoff, h_outoff.data(), (size_t)nc * 8));
          DTRY(M.to_dev(d_crcoff, h_crcoff.data(), (accepted.size() + 1) * 8));
          run_d5c(L, nc, d5_nchunks(nmax), d_info, stride, d_W, chunks_pitch, d_chentry, d_choff, d_outoff, d_out);
          DTRY(M.crc_blocks(d_out, d_crcoff, (uint32_t)accepted.size(), d_crc));
          DTRY(M.check());
          DTRY(M.to_host(h_crc.data(), d_crc, accepted.size() * 4));
          for (size_t k = 0; k < accepted.size(); ++k) {
            if (h_crc[k] != infos[accepted[k].first].stored_crc) {  // found when the next header is read (decoder.rs:198-204)
              crc_err = E_DATA;
              crc_err_out = h_crcoff[k + 1];
              break;
            }
          }
        }
      }
      if (crc_err || chain.done) break;
      if (verdict == Chain::LATER) {
        c0 = c1;
        continue;
      }
      // verdict == MISS: no complete magic starts at chain.pos.  What the reference does there (decoder.rs:206-224,
      // 494-508, 521-523): the head byte must exist (else UnexpectedEof) and be 0x31 or 0x17 (else DataError); the
      // five bytes after it are read and NOT compared (a missing one is DataError).  So a block or an end-of-stream
      // marker with damaged magic bytes 2..6 is decoded all the same: inject it as a candidate and go on from there.
      const uint64_t pos = chain.pos;
      if (nbits - pos < 8) {
        chain.fail(E_EOF);
        break;
      }
      uint8_t two[2] = {0, 0};
      DTRY(M.to_host(two, d_in + (pos >> 3), (pos >> 3) + 2 <= n ? 2 : 1));
      const uint32_t head = (((uint32_t)two[0] << 8 | two[1]) >> (8 - (pos & 7))) & 0xFFu;
      if ((head != 0x31 && head != 0x17) || nbits - pos < 48) {
        chain.fail(E_DATA);
        break;
      }
      const uint64_t entry = pos | (head == 0x17 ? KIND_END : 0ull);
      size_t at = 0;
      while (at < cand.size() && (cand[at] & ~KIND_END) < pos) ++at;
      cand.insert(cand.begin() + at, entry);
      if (head == 0x17 && maxlevel < 9) {  // a stream of unknown level may follow: size for the largest and start over
        maxlevel = 9;
        restart = true;
        break;
      }
      d_cand = slot<uint64_t>(M, DS_CAND, cand.size() + 1);
      if (!d_cand) return -2;
      DTRY(M.to_dev(d_cand, cand.data(), cand.size() * 8));
      c0 = at;
    }
    if (restart && attempt < 64) continue;
    if (!chain.done && !crc_err) chain.fail(E_UNEXPECTED);  // cannot happen
    res->streams = chain.streams;
    res->blocks = chain.blocks;
    res->syms = chain.syms;
    res->pre_rle = chain.pre_rle;
    if (dry) {
      res->too_small = 1;
      res->needed = chain.out_total;
      res->out_n = 0;
      res->bz_error = chain.err;
      return 0;
    }
    if (crc_err) {
      res->bz_error = crc_err;
      res->out_n = res->needed = crc_err_out;
      return 0;
    }
    res->bz_error = chain.err;
    res->out_n = res->needed = chain.out_total;
    return 0;
  }
}

}  // namespace bzb
