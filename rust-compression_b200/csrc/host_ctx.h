// host_ctx.h — the context object behind the C ABI and the small host helpers shared by pipeline.cu (encode path),
// enc_stream.cu (streaming encoder object) and dec_abi.cu (decoder entry points).  Internal; not installed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/bzb200.h"
#include "common.cuh"
#include "decoder.h"
#include "kernels.h"

using namespace bzb;

// ============================================================== context
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

constexpr uint32_t MAX_BATCH_BLOCKS = 32768;

struct bzb200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  Launcher L;
  std::string err;

  // ---- plan ----
  int level = 0;
  uint32_t T = 0;
  const uint8_t* d_in = nullptr;
  uint64_t n_in = 0;
  uint32_t nblocks = 0;
  uint32_t max_block_len = 0;
  bool planned = false;
  bool plan_open = false;  // between plan_begin and plan_finish
  DevBuf tile_head, tile_carry, tile_cnt, tile_E, in_off, rle_off, txt, crc, inuse, scal, cut_state, cut_F;
  uint32_t prep_lo = 0, prep_hi = 0;  // blocks whose RLE1 bytes, CRC and in-use map are on the device
  bool crc_all = false;
  // ---- sliced plan (include/bzb200.h section 2b): this context holds input bytes [sl_lo, sl_avail) of an n_in-byte
  // stream; d_in / tile arrays / txt are then addressed through the v_*() accessors below, which shift the pointers so
  // that GLOBAL byte, tile and emitted-offset indices work unchanged in every kernel
  bool sliced = false;
  const uint8_t* sl_d_lo = nullptr;            // device address of input byte sl_lo
  uint64_t sl_lo = 0, sl_hi = 0, sl_avail = 0;  // own bytes [sl_lo, sl_hi); resident up to sl_avail (halo / block tail)
  uint64_t sl_t0 = 0, sl_t1 = 0, sl_tn = 0;    // own tiles [t0, t1); tiles [t0, tn) have summaries (resident + 1 byte)
  long long sl_carry_in = -1;                  // last run head before the slice
  uint64_t sl_E_lo = 0, sl_E_hi = 0, sl_E_tot = 0;  // emitted offsets at sl_lo / sl_hi / end of the stream
  uint64_t txt_origin = 0;                     // emitted offset of txt.p[0]
  int sl_stage = 0;                            // 1 begun, 2 counted, 3 prefixed
  DevBuf sl_F, sl_sum;
  std::vector<uint64_t> h_in_off, h_rle_off;
  std::vector<uint32_t> h_crc;
  std::vector<uint32_t> h_inuse;  // [nblocks][8] in-use byte maps of the prepared blocks (alphabet sizes for K2 / K3)

  // ---- batch scratch ----
  DevBuf desc, A, B, rank, sa, tile_meta, cnt, hist, oshist, pairhist, ticket, tsum, state, shift, sparse, stats, rounds, global, last, origptr;
  DevBuf chunk_state, chunk_zle, chunk_base, sym, freq, mtf_count;
  DevBuf lens, rfreq, sel, selmtf, codes, gbits, meta, lm_scratch, lm_list, lm_count, blockbit, bitcursor, combined;
  DevBuf stage_in, stage_out;  // bzb200_compress_host staging
  DevBuf dec_bufs[DS_NSLOTS];  // decoder scratch (decoder.cu), one buffer per role
  DevBuf dec_in, dec_out;      // bzb200_decompress_host / bzb200_dec staging
  DecResult dec_last;
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  std::vector<cudaEvent_t> seg_events;
  uint64_t batch_elems_cap = (uint64_t)1400 * 1000 * 1000;

  // ---- last batch (debug) ----
  uint32_t batch_b0 = 0, batch_nb = 0;
  std::vector<BlockDesc> h_desc;
  std::vector<uint64_t> h_blockbit;
  std::vector<uint32_t> h_mtf_count;
  uint8_t* last_out = nullptr;
  BwtStats bstats;
  uint64_t stat_rle = 0, stat_mtf = 0;

  std::vector<DevBuf*> all;
};

#define CK(ctx, call)                                                                                  \
  do {                                                                                                 \
    cudaError_t e__ = (call);                                                                          \
    if (e__ != cudaSuccess) {                                                                          \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                                \
      return BZB200_E_CUDA;                                                                            \
    }                                                                                                  \
  } while (0)

inline int ensure(bzb200_ctx* c, DevBuf& b, size_t bytes) {
  if (bytes == 0) bytes = 16;
  if (b.cap >= bytes) return BZB200_OK;
  if (b.p) {
    CK(c, cudaStreamSynchronize(c->stream));
    CK(c, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
  }
  size_t want = bytes + bytes / 16 + 256;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    want = bytes;
    e = cudaMalloc(&b.p, want);
  }
  if (e != cudaSuccess) {
    c->err = std::string("cudaMalloc(") + std::to_string(want) + "): " + cudaGetErrorString(e);
    b.p = nullptr;
    return BZB200_E_CUDA;
  }
  b.cap = want;
  return BZB200_OK;
}

inline int check_launch(bzb200_ctx* c) {
  if (c->L.err != cudaSuccess) {
    c->err = std::string("kernel launch ") + (c->L.err_kernel ? c->L.err_kernel : "?") + ": " +
             cudaGetErrorString(c->L.err);
    c->L.err = cudaSuccess;
    return BZB200_E_CUDA;
  }
  return BZB200_OK;
}

template <class T>
T* ptr(DevBuf& b) {
  return reinterpret_cast<T*>(b.p);
}

#define TRY(x)                \
  do {                        \
    int r__ = (x);            \
    if (r__ != BZB200_OK) return r__; \
  } while (0)

inline int level_of(const bzb200_ctx* c) { return c->level; }

// Pointers addressed by global indices (identity unless the plan is sliced).
inline const uint8_t* v_in(const bzb200_ctx* c) { return c->sliced ? c->sl_d_lo - c->sl_lo : c->d_in; }
inline long long* v_carry(bzb200_ctx* c) { return ptr<long long>(c->tile_carry) - (c->sliced ? c->sl_t0 : 0); }
inline long long* v_head(bzb200_ctx* c) { return ptr<long long>(c->tile_head) - (c->sliced ? c->sl_t0 : 0); }
inline uint32_t* v_cnt(bzb200_ctx* c) { return ptr<uint32_t>(c->tile_cnt) - (c->sliced ? c->sl_t0 : 0); }
inline uint64_t* v_E(bzb200_ctx* c) { return ptr<uint64_t>(c->tile_E) - (c->sliced ? c->sl_t0 : 0); }
inline uint8_t* v_txt(bzb200_ctx* c) { return ptr<uint8_t>(c->txt) - c->txt_origin; }

inline int set_device(bzb200_ctx* c) {
  CK(c, cudaSetDevice(c->device));
  return BZB200_OK;
}

// Creates a context; own_stream: the context creates (and later destroys) a non-blocking stream of its own.
// Defined in pipeline.cu; hidden visibility (only BZB200_API symbols are exported).
int bzb200_ctx_create_impl(int device, void* stream, bool own_stream, bzb200_ctx** out);

// slice_plan.cu: sizes the RLE1 buffer of a sliced context for blocks [b0, b1) and sets its origin.
int slice_reserve_txt(bzb200_ctx* c, uint32_t b0, uint32_t b1);
