// pipeline.cu — host side of libbzb200: context, device memory, stage orchestration of the encode path and its part
// of the C ABI (include/bzb200.h sections 2 and 4; the streaming encoder object is in enc_stream.cu, the decoder entry
// points in dec_abi.cu).  No CPU implementation of any stage lives here: if CUDA is unavailable every compute entry
// point returns BZB200_E_CUDA.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
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

// ============================================================== Launcher
cudaEvent_t Launcher::get_event() {
  if (!pool.empty()) {
    cudaEvent_t e = pool.back();
    pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

void Launcher::raw_launch(const char* name, const void* fn, dim3 grid, dim3 block, size_t smem, void** args) {
  if (err != cudaSuccess) return;
  Pending p{name, nullptr, nullptr};
  if (profiling) {
    p.a = get_event();
    p.b = get_event();
    cudaEventRecord(p.a, stream);
  }
  cudaError_t e = cudaLaunchKernel(fn, grid, block, args, smem, stream);
  if (e != cudaSuccess) {
    err = e;
    err_kernel = name;
    return;
  }
  ++launches;
  if (profiling) {
    cudaEventRecord(p.b, stream);
    pending.push_back(p);
  }
}

void Launcher::resolve() {
  if (pending.empty()) return;
  cudaStreamSynchronize(stream);
  for (auto& p : pending) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, p.a, p.b);
    bool found = false;
    for (auto& r : recs)
      if (strcmp(r.name, p.name) == 0) {
        r.launches += 1;
        r.ms += ms;
        found = true;
        break;
      }
    if (!found) recs.push_back(Rec{p.name, 1, (double)ms});
    pool.push_back(p.a);
    pool.push_back(p.b);
  }
  pending.clear();
}

void Launcher::clear_profile() {
  resolve();
  recs.clear();
}

Launcher::~Launcher() {
  for (auto& p : pending) {
    cudaEventDestroy(p.a);
    cudaEventDestroy(p.b);
  }
  for (auto e : pool) cudaEventDestroy(e);
}

#include "host_ctx.h"

extern "C" {

const char* bzb200_version(void) { return "bzb200 0.1 (sm_100a)"; }

}  // extern "C"

int bzb200_ctx_create_impl(int device, void* stream, bool own_stream, bzb200_ctx** out) {
  if (!out) return BZB200_E_ARG;
  *out = nullptr;
  bzb200_ctx* c = new bzb200_ctx();
  if (device < 0) {
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) {
      c->err = std::string("cudaGetDevice: ") + cudaGetErrorString(e);
      *out = c;  // returned so that the caller can read the message
      return BZB200_E_CUDA;
    }
  }
  c->device = device;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    c->err = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
    *out = c;
    return BZB200_E_CUDA;
  }
  if (own_stream) {
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      c->err = std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
      *out = c;
      return BZB200_E_CUDA;
    }
    c->own_stream = true;
  } else {
    c->stream = reinterpret_cast<cudaStream_t>(stream);  // NULL = the device's default stream
  }
  c->L.stream = c->stream;
  const char* be = getenv("BZB200_BATCH_ELEMS");
  if (be) {
    unsigned long long v = strtoull(be, nullptr, 10);
    if (v >= 1000) c->batch_elems_cap = v;
  }
  c->all = {&c->tile_head, &c->tile_carry, &c->tile_cnt, &c->tile_E, &c->in_off, &c->rle_off, &c->txt, &c->crc,
            &c->inuse, &c->scal, &c->cut_state, &c->cut_F, &c->desc, &c->A, &c->B, &c->rank, &c->sa,
            &c->tile_meta, &c->cnt, &c->hist, &c->oshist, &c->ticket, &c->tsum, &c->state, &c->shift, &c->sparse,
            &c->stats, &c->rounds, &c->global, &c->last, &c->origptr, &c->chunk_state, &c->chunk_zle, &c->chunk_base,
            &c->sym, &c->freq, &c->mtf_count, &c->lens, &c->rfreq, &c->sel, &c->selmtf, &c->codes, &c->gbits, &c->meta,
            &c->lm_scratch, &c->lm_list, &c->lm_count, &c->blockbit, &c->bitcursor, &c->combined, &c->stage_in,
            &c->stage_out, &c->dec_in, &c->dec_out, &c->sl_F, &c->sl_sum, &c->pairhist};
  for (DevBuf& b : c->dec_bufs) c->all.push_back(&b);
  *out = c;
  return BZB200_OK;
}

extern "C" {

int bzb200_ctx_create(int device, void* stream, bzb200_ctx** out) { return bzb200_ctx_create_impl(device, stream, false, out); }

void bzb200_ctx_destroy(bzb200_ctx* c) {
  if (!c) return;
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (DevBuf* b : c->all)
    if (b->p) cudaFree(b->p);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  if (c->h2d_stream) cudaStreamDestroy(c->h2d_stream);
  if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
  for (cudaEvent_t ev : c->seg_events) cudaEventDestroy(ev);
  delete c;
}

const char* bzb200_last_error(const bzb200_ctx* c) { return c ? c->err.c_str() : "null context"; }

int bzb200_sync(bzb200_ctx* c) {
  if (!c) return BZB200_E_ARG;
  CK(c, cudaStreamSynchronize(c->stream));
  return BZB200_OK;
}

size_t bzb200_max_output_bytes(int level, size_t n) {
  if (level < 1 || level > 9) level = 1;
  size_t T = (size_t)level * 100000 - 19;
  size_t nblocks = (n + n / 4) / T + 2;
  size_t bytes = n + n / 2 + nblocks * 4096 + 1024;
  return (bytes + 7) & ~(size_t)7;
}

// ------------------------------------------------------------------ plan (K1: run pieces, emitted-length prefix, cuts)
size_t bzb200_plan_tile_bytes(void) { return k1_tile_bytes(); }

int bzb200_plan_begin(bzb200_ctx* c, int level, const uint8_t* d_in, size_t n, uint64_t* ntiles) {
  if (!c) return BZB200_E_ARG;
  if (level < 1 || level > 9) {
    c->err = "invalid level";
    return BZB200_E_LEVEL;
  }
  if (!d_in && n) return BZB200_E_ARG;
  TRY(set_device(c));
  c->planned = false;
  c->plan_open = false;
  c->sliced = false;
  c->sl_stage = 0;
  c->txt_origin = 0;
  c->level = level;
  c->T = (uint32_t)level * 100000u - 19u;  // encoder.rs:186
  c->d_in = d_in;
  c->n_in = n;
  c->nblocks = 0;
  c->max_block_len = 0;
  c->h_in_off.assign(1, 0);
  c->h_rle_off.assign(1, 0);
  c->h_crc.clear();
  c->batch_nb = 0;
  c->prep_lo = c->prep_hi = 0;
  c->crc_all = false;
  const uint64_t nt = n ? k1_num_tiles(n) : 0;
  if (ntiles) *ntiles = nt;
  c->plan_open = true;
  if (n == 0) return BZB200_OK;
  const uint64_t emax = (uint64_t)n + n / 4 + 64;
  const uint64_t max_blocks64 = emax / c->T + 2;
  if (max_blocks64 > 0x7FFFFFF0ull) return BZB200_E_ARG;
  const uint32_t max_blocks = (uint32_t)max_blocks64;
  TRY(ensure(c, c->tile_head, nt * 8));
  TRY(ensure(c, c->tile_carry, nt * 8));
  TRY(ensure(c, c->tile_cnt, nt * 4));
  TRY(ensure(c, c->tile_E, (nt + 1) * 8));
  TRY(ensure(c, c->in_off, ((size_t)max_blocks + 1) * 8));
  TRY(ensure(c, c->rle_off, ((size_t)max_blocks + 1) * 8));
  TRY(ensure(c, c->scal, 64));
  TRY(ensure(c, c->txt, emax));
  return BZB200_OK;
}

int bzb200_plan_heads(bzb200_ctx* c, uint64_t t0, uint64_t t1, int64_t* d_tile_head) {
  if (!c || !c->plan_open) return BZB200_E_STATE;
  const uint64_t nt = c->n_in ? k1_num_tiles(c->n_in) : 0;
  if (t0 > t1 || t1 > nt || (!d_tile_head && nt)) return BZB200_E_ARG;
  TRY(set_device(c));
  launch_k1_heads(c->L, c->d_in, c->n_in, t0, t1, reinterpret_cast<long long*>(d_tile_head));
  return check_launch(c);
}

int bzb200_plan_counts(bzb200_ctx* c, const int64_t* d_tile_head, uint64_t t0, uint64_t t1, uint32_t* d_tile_cnt) {
  if (!c || !c->plan_open) return BZB200_E_STATE;
  const uint64_t nt = c->n_in ? k1_num_tiles(c->n_in) : 0;
  if (t0 > t1 || t1 > nt || ((!d_tile_head || !d_tile_cnt) && nt)) return BZB200_E_ARG;
  if (nt == 0) return BZB200_OK;
  TRY(set_device(c));
  launch_k1_counts(c->L, c->d_in, c->n_in, t0, t1, reinterpret_cast<const long long*>(d_tile_head),
                   ptr<long long>(c->tile_carry), d_tile_cnt);
  return check_launch(c);
}

int bzb200_plan_finish(bzb200_ctx* c, const uint32_t* d_tile_cnt, uint32_t* nblocks) {
  if (!c || !c->plan_open) return BZB200_E_STATE;
  c->plan_open = false;
  const size_t n = c->n_in;
  if (n == 0) {
    c->planned = true;
    if (nblocks) *nblocks = 0;
    return BZB200_OK;
  }
  if (!d_tile_cnt) return BZB200_E_ARG;
  TRY(set_device(c));
  const uint8_t* d_in = c->d_in;
  const uint64_t emax = (uint64_t)n + n / 4 + 64;
  const uint32_t max_blocks = (uint32_t)(emax / c->T + 2);
  launch_k1_prefix(c->L, n, d_tile_cnt, ptr<uint64_t>(c->tile_E));
  // cut chain: phases of (windows, walk) until the walk reports done (one phase unless the drift leaves a window)
  TRY(ensure(c, c->cut_state, 64));
  CK(c, cudaMemsetAsync(c->cut_state.p, 0, 64, c->stream));
  uint32_t sc[2] = {0, 0};
  uint64_t st[4] = {0, 0, 0, 0};
  for (int phase = 0; phase < 1 << 20; ++phase) {
    const uint64_t left = emax - std::min<uint64_t>(emax, st[1]);
    const uint32_t K = (uint32_t)std::min<uint64_t>(left / c->T + 2, 1u << 20);
    TRY(ensure(c, c->cut_F, (size_t)K * k1_cut_window() * 8));
    launch_k1_cut_phase(c->L, d_in, n, c->T, ptr<long long>(c->tile_carry), ptr<uint64_t>(c->tile_E), K,
                        ptr<uint64_t>(c->cut_F), ptr<uint64_t>(c->cut_state), ptr<uint64_t>(c->in_off),
                        ptr<uint64_t>(c->rle_off), max_blocks, ptr<uint32_t>(c->scal), ptr<uint32_t>(c->scal) + 1);
    TRY(check_launch(c));
    CK(c, cudaMemcpyAsync(st, c->cut_state.p, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaMemcpyAsync(sc, c->scal.p, sizeof(sc), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    if (st[2]) break;
  }
  const uint32_t nb = sc[0];
  if (!st[2] || nb == 0 || nb > max_blocks) {
    c->err = "cut chain produced an invalid block count";
    return BZB200_E_INTERNAL;
  }
  c->nblocks = nb;
  c->max_block_len = sc[1];
  c->h_in_off.resize((size_t)nb + 1);
  c->h_rle_off.resize((size_t)nb + 1);
  c->h_crc.assign(nb, 0);
  TRY(ensure(c, c->crc, (size_t)nb * 4));
  TRY(ensure(c, c->inuse, (size_t)nb * 32));
  CK(c, cudaMemcpyAsync(c->h_in_off.data(), c->in_off.p, ((size_t)nb + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaMemcpyAsync(c->h_rle_off.data(), c->rle_off.p, ((size_t)nb + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  if (c->max_block_len > (uint32_t)level_of(c) * 100000u || c->max_block_len > MAX_BLOCK) {
    c->err = "block longer than level*100000";
    return BZB200_E_INTERNAL;
  }
  c->planned = true;
  if (nblocks) *nblocks = nb;
  return BZB200_OK;
}

int bzb200_plan(bzb200_ctx* c, int level, const uint8_t* d_in, size_t n, uint32_t* nblocks) {
  uint64_t nt = 0;
  TRY(bzb200_plan_begin(c, level, d_in, n, &nt));
  if (nt) {
    TRY(bzb200_plan_heads(c, 0, nt, ptr<int64_t>(c->tile_head)));
    TRY(bzb200_plan_counts(c, ptr<int64_t>(c->tile_head), 0, nt, ptr<uint32_t>(c->tile_cnt)));
  }
  return bzb200_plan_finish(c, ptr<uint32_t>(c->tile_cnt), nblocks);
}

uint32_t bzb200_num_blocks(const bzb200_ctx* c) { return (c && c->planned) ? c->nblocks : 0; }

int bzb200_block_table(bzb200_ctx* c, uint64_t* in_off, uint64_t* rle_off, uint32_t* crc) {
  if (!c || !c->planned) return BZB200_E_STATE;
  if (in_off) memcpy(in_off, c->h_in_off.data(), c->h_in_off.size() * 8);
  if (rle_off) memcpy(rle_off, c->h_rle_off.data(), c->h_rle_off.size() * 8);
  if (crc && c->nblocks) {
    if (!c->crc_all && !(c->prep_lo == 0 && c->prep_hi == c->nblocks)) {  // CRCs of blocks this context did not encode
      if (c->sliced) {
        c->err = "block_table: a sliced context holds the CRCs of its own blocks only (bzb200_block_crcs)";
        return BZB200_E_STATE;
      }
      TRY(set_device(c));
      launch_k5_crc(c->L, c->d_in, ptr<uint64_t>(c->in_off), c->nblocks, ptr<uint32_t>(c->crc));
      TRY(check_launch(c));
      CK(c, cudaMemcpyAsync(c->h_crc.data(), c->crc.p, (size_t)c->nblocks * 4, cudaMemcpyDeviceToHost, c->stream));
      CK(c, cudaStreamSynchronize(c->stream));
      c->crc_all = true;
    }
    memcpy(crc, c->h_crc.data(), (size_t)c->nblocks * 4);
  }
  return BZB200_OK;
}

int bzb200_block_crcs(const bzb200_ctx* c, uint32_t* crc, size_t cap) {
  if (!c || !c->planned || !crc) return BZB200_E_STATE;
  for (size_t i = 0; i < cap && i < c->h_crc.size(); ++i) crc[i] = c->h_crc[i];
  return BZB200_OK;
}

// RLE1 bytes, CRCs and in-use maps of blocks [b0, b1) (K1 scatter, K5, in-use) — only what this context encodes.
static int prepare_blocks(bzb200_ctx* c, uint32_t b0, uint32_t b1) {
  if (b0 >= b1 || (b0 >= c->prep_lo && b1 <= c->prep_hi)) return BZB200_OK;
  if (c->sliced) {
    const uint64_t need = std::min<uint64_t>(c->n_in, (c->h_in_off[b1] + k1_tile_bytes() - 1) / k1_tile_bytes() * k1_tile_bytes());
    const uint64_t tn_need = (need + k1_tile_bytes() - 1) / k1_tile_bytes();
    if (c->h_in_off[b0] < c->sl_lo || tn_need > c->sl_tn) {
      c->err = "encode_blocks: blocks reach outside the resident part of the slice (bzb200_slice_extend)";
      return BZB200_E_STATE;
    }
  }
  launch_k1_scatter(c->L, v_in(c), c->n_in, c->h_in_off[b0], c->h_in_off[b1], v_carry(c), v_E(c), v_txt(c));
  launch_k5_crc(c->L, v_in(c), ptr<uint64_t>(c->in_off) + b0, b1 - b0, ptr<uint32_t>(c->crc) + b0);
  launch_k1_inuse(c->L, v_txt(c), ptr<uint64_t>(c->rle_off) + b0, b1 - b0,
                  ptr<uint32_t>(c->inuse) + (size_t)b0 * 8);
  TRY(check_launch(c));
  CK(c, cudaMemcpyAsync(c->h_crc.data() + b0, ptr<uint32_t>(c->crc) + b0, (size_t)(b1 - b0) * 4, cudaMemcpyDeviceToHost,
                        c->stream));
  if (c->h_inuse.size() < (size_t)c->nblocks * 8) c->h_inuse.resize((size_t)c->nblocks * 8);
  CK(c, cudaMemcpyAsync(c->h_inuse.data() + (size_t)b0 * 8, ptr<uint32_t>(c->inuse) + (size_t)b0 * 8,
                        (size_t)(b1 - b0) * 32, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  c->prep_lo = b0;
  c->prep_hi = b1;
  return BZB200_OK;
}

// ------------------------------------------------------------------ encode (K2..K6)
static int encode_batch(bzb200_ctx* c, uint32_t b0, uint32_t nb, uint8_t* d_out, size_t cap_bytes) {
  const uint64_t base = c->h_rle_off[b0];
  const uint64_t M = c->h_rle_off[b0 + nb] - base;
  // descriptors
  c->h_desc.resize(nb);
  uint32_t nmax = 0;
  uint64_t symoff = 0;
  for (uint32_t i = 0; i < nb; ++i) {
    BlockDesc& d = c->h_desc[i];
    d.off = (uint32_t)(c->h_rle_off[b0 + i] - base);
    d.n = (uint32_t)(c->h_rle_off[b0 + i + 1] - c->h_rle_off[b0 + i]);
    d.symoff = (uint32_t)symoff;
    d.pad = 0;
    symoff += ((uint64_t)d.n + 1 + 7) & ~7ull;
    nmax = std::max(nmax, d.n);
  }
  if (symoff >= 0xFFFFFFF0ull || M >= 0xFFFFFFF0ull) return BZB200_E_INTERNAL;
  const uint32_t tile = bwt_tile_elems();
  const uint32_t tiles = (nmax + tile - 1) / tile;
  const uint32_t chunk = mtf_chunk_elems(M);
  const uint32_t chunks = (nmax + chunk - 1) / chunk;
  const uint32_t max_groups = (nmax + 1 + G_SIZE - 1) / G_SIZE;

  TRY(ensure(c, c->desc, (size_t)nb * sizeof(BlockDesc)));
  TRY(ensure(c, c->A, M * 8 + 16));  // + one element pair: a bulk copy of a tile may read one element past it
  TRY(ensure(c, c->B, M * 8 + 16));
  TRY(ensure(c, c->rank, M * 4));
  TRY(ensure(c, c->sa, M * 4));
  const uint32_t ls_tile = bwt_ls_tile_elems();
  const uint32_t ls_tiles = (nmax + ls_tile - 1) / ls_tile + 1;
  TRY(ensure(c, c->tile_meta, (size_t)nb * ls_tiles * 8));
  TRY(ensure(c, c->cnt, (size_t)nb * 4));
  TRY(ensure(c, c->hist, (size_t)nb * tiles * 512 * 4));
  TRY(ensure(c, c->oshist, (size_t)nb * 5 * 512 * 4));
  uint32_t max_alpha = 0;  // most in-use byte values of any block of the batch
  for (uint32_t i = 0; i < nb; ++i) {
    uint32_t a = 0;
    for (int k = 0; k < 8; ++k) a += (uint32_t)__builtin_popcount(c->h_inuse[(size_t)(b0 + i) * 8 + k]);
    max_alpha = std::max(max_alpha, a);
  }
  TRY(ensure(c, c->pairhist, bwt_pairhist_bytes(nb, max_alpha)));
  TRY(ensure(c, c->ticket, ((size_t)nb + 1) * 4));  // per-block tile tickets + the global counter of k2_os_scatter_pf
  TRY(ensure(c, c->tsum, (size_t)nb * tiles * sizeof(int4)));
  TRY(ensure(c, c->state, (size_t)nb * 4));
  TRY(ensure(c, c->shift, (size_t)nb * 4));
  TRY(ensure(c, c->sparse, (size_t)nb * 4));
  TRY(ensure(c, c->stats, (size_t)nb * 16));
  TRY(ensure(c, c->rounds, (size_t)nb * 4));
  TRY(ensure(c, c->global, 64));
  TRY(ensure(c, c->last, M + 16));  // K3 reads the last column through aligned 32-bit words
  TRY(ensure(c, c->origptr, (size_t)nb * 4));
  TRY(ensure(c, c->chunk_state, (size_t)nb * chunks * 256 * 4));
  TRY(ensure(c, c->chunk_zle, (size_t)nb * chunks * sizeof(uint4)));
  TRY(ensure(c, c->chunk_base, (size_t)nb * chunks * sizeof(uint2)));
  TRY(ensure(c, c->sym, symoff * 2));
  TRY(ensure(c, c->freq, (size_t)nb * MAX_ALPHA * 4));
  TRY(ensure(c, c->mtf_count, (size_t)nb * 4));
  TRY(ensure(c, c->lens, (size_t)nb * 5 * MAX_GROUPS * MAX_ALPHA));
  TRY(ensure(c, c->rfreq, (size_t)nb * MAX_GROUPS * MAX_ALPHA * 4));
  TRY(ensure(c, c->sel, (size_t)nb * MAX_SELECTORS));
  TRY(ensure(c, c->selmtf, (size_t)nb * MAX_SELECTORS));
  TRY(ensure(c, c->codes, (size_t)nb * MAX_GROUPS * MAX_ALPHA * 4));
  TRY(ensure(c, c->gbits, (size_t)nb * MAX_SELECTORS * 4));
  TRY(ensure(c, c->meta, (size_t)nb * 8 * 4));
  const uint32_t lm_slots = 512;
  TRY(ensure(c, c->lm_scratch, (size_t)lm_slots * huff_lm_scratch_bytes()));
  TRY(ensure(c, c->lm_list, (size_t)nb * MAX_GROUPS * 4));
  TRY(ensure(c, c->lm_count, 16));
  TRY(ensure(c, c->blockbit, ((size_t)nb + 1) * 8));

  CK(c, cudaMemcpyAsync(c->desc.p, c->h_desc.data(), (size_t)nb * sizeof(BlockDesc), cudaMemcpyHostToDevice, c->stream));
  const uint8_t* d_txt = v_txt(c) + base;
  const BlockDesc* d_desc = ptr<BlockDesc>(c->desc);
  const uint32_t* d_inuse = ptr<uint32_t>(c->inuse) + (size_t)b0 * 8;
  const uint32_t* d_crc = ptr<uint32_t>(c->crc) + b0;

  BwtScratch S;
  S.A = ptr<uint64_t>(c->A);
  S.B = ptr<uint64_t>(c->B);
  S.rank = ptr<uint32_t>(c->rank);
  S.sa = ptr<uint32_t>(c->sa);
  S.tile_meta = ptr<uint2>(c->tile_meta);
  S.ls_tiles_cap = ls_tiles;
  S.cnt = ptr<uint32_t>(c->cnt);
  S.hist = ptr<uint32_t>(c->hist);
  S.oshist = ptr<uint32_t>(c->oshist);
  S.pairhist = ptr<uint32_t>(c->pairhist);
  S.ticket = ptr<uint32_t>(c->ticket);
  S.tsum = ptr<int4>(c->tsum);
  S.state = ptr<uint32_t>(c->state);
  S.shift = ptr<uint32_t>(c->shift);
  S.sparse = ptr<uint32_t>(c->sparse);
  S.stats = ptr<uint32_t>(c->stats);
  S.rounds = ptr<uint32_t>(c->rounds);
  S.global = ptr<uint32_t>(c->global);
  S.tiles_cap = tiles;
  int r = run_bwt(c->L, d_txt, d_desc, d_inuse, max_alpha, nb, nmax, M, S, ptr<uint8_t>(c->last), ptr<uint32_t>(c->origptr),
                  &c->bstats);
  c->stat_rle += M;
  if (r != 0) {
    TRY(check_launch(c));
    cudaError_t e = cudaGetLastError();
    c->err = r == -5 ? "rotation sort did not converge (internal)"
                     : std::string("rotation sort: ") + cudaGetErrorString(e);
    return r == -5 ? BZB200_E_INTERNAL : BZB200_E_CUDA;
  }

  launch_mtf(c->L, ptr<uint8_t>(c->last), d_desc, d_inuse, nb, nmax, max_alpha, chunk, ptr<int>(c->chunk_state),
             ptr<uint4>(c->chunk_zle), ptr<uint2>(c->chunk_base), chunks, ptr<uint16_t>(c->sym),
             ptr<uint32_t>(c->freq), ptr<uint32_t>(c->mtf_count));

  HuffBuffers H;
  H.lens = ptr<uint8_t>(c->lens);
  H.rfreq = ptr<uint32_t>(c->rfreq);
  H.sel = ptr<uint8_t>(c->sel);
  H.selmtf = ptr<uint8_t>(c->selmtf);
  H.codes = ptr<uint32_t>(c->codes);
  H.gbits = ptr<uint32_t>(c->gbits);
  H.meta = ptr<uint32_t>(c->meta);
  H.lm_scratch = ptr<uint8_t>(c->lm_scratch);
  H.lm_list = ptr<uint32_t>(c->lm_list);
  H.lm_count = ptr<uint32_t>(c->lm_count);
  H.lm_slots = lm_slots;
  launch_huffman(c->L, ptr<uint16_t>(c->sym), d_desc, ptr<uint32_t>(c->mtf_count), ptr<uint32_t>(c->freq), d_inuse, nb,
                 max_groups, H);
  TRY(check_launch(c));

  // bit offsets first, so that the capacity can be checked before a single bit is written
  uint64_t cursor_before = 0, cursor_after = 0;
  CK(c, cudaMemcpyAsync(&cursor_before, c->bitcursor.p, 8, cudaMemcpyDeviceToHost, c->stream));
  launch_pack(c->L, ptr<uint16_t>(c->sym), d_desc, ptr<uint32_t>(c->mtf_count), d_inuse, d_crc,
              ptr<uint32_t>(c->origptr), nb, max_groups, H, ptr<uint64_t>(c->blockbit), ptr<uint64_t>(c->bitcursor),
              nullptr);  // offsets only (d_out == nullptr)
  TRY(check_launch(c));
  CK(c, cudaMemcpyAsync(&cursor_after, c->bitcursor.p, 8, cudaMemcpyDeviceToHost, c->stream));
  c->h_mtf_count.resize(nb);
  CK(c, cudaMemcpyAsync(c->h_mtf_count.data(), c->mtf_count.p, (size_t)nb * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  for (uint32_t i = 0; i < nb; ++i) c->stat_mtf += c->h_mtf_count[i];
  (void)cursor_before;
  if ((cursor_after + 7) / 8 + 16 > cap_bytes) {
    c->err = "output buffer too small: need " + std::to_string((cursor_after + 7) / 8 + 16) + " bytes";
    return BZB200_E_ARG;
  }
  launch_pack(c->L, ptr<uint16_t>(c->sym), d_desc, ptr<uint32_t>(c->mtf_count), d_inuse, d_crc,
              ptr<uint32_t>(c->origptr), nb, max_groups, H, ptr<uint64_t>(c->blockbit), nullptr, d_out);
  TRY(check_launch(c));
  // device-side invariants
  c->batch_b0 = b0;
  c->batch_nb = nb;
  c->last_out = d_out;
  return BZB200_OK;
}

int bzb200_encode_blocks(bzb200_ctx* c, uint32_t b0, uint32_t b1, uint8_t* d_out, size_t cap_bytes,
                         uint64_t start_bit, uint64_t* end_bit) {
  if (!c) return BZB200_E_ARG;
  if (!c->planned) {
    c->err = "bzb200_encode_blocks before bzb200_plan";
    return BZB200_E_STATE;
  }
  if (b0 > b1 || b1 > c->nblocks || !d_out || (reinterpret_cast<uintptr_t>(d_out) & 3)) {
    c->err = "bad block range or unaligned output";
    return BZB200_E_ARG;
  }
  TRY(set_device(c));
  if (c->sliced && !(b0 >= c->prep_lo && b1 <= c->prep_hi)) TRY(slice_reserve_txt(c, b0, b1));
  TRY(prepare_blocks(c, b0, b1));
  TRY(ensure(c, c->bitcursor, 16));
  CK(c, cudaMemcpyAsync(c->bitcursor.p, &start_bit, 8, cudaMemcpyHostToDevice, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  c->bstats = BwtStats();
  c->stat_rle = 0;
  c->stat_mtf = 0;
  uint32_t b = b0;
  while (b < b1) {
    uint32_t e = b;
    uint64_t elems = 0;
    while (e < b1 && e - b < MAX_BATCH_BLOCKS) {
      uint64_t n = c->h_rle_off[e + 1] - c->h_rle_off[e];
      if (e > b && elems + n > c->batch_elems_cap) break;
      elems += n;
      ++e;
    }
    TRY(encode_batch(c, b, e - b, d_out, cap_bytes));
    b = e;
  }
  uint64_t cur = start_bit;
  CK(c, cudaMemcpyAsync(&cur, c->bitcursor.p, 8, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  TRY(check_launch(c));
  if (end_bit) *end_bit = cur;
  return BZB200_OK;
}

int bzb200_bit_append(bzb200_ctx* c, uint8_t* d_dst, size_t dst_cap_bytes, uint64_t dst_bit, const uint8_t* d_src,
                      uint64_t nbits) {
  if (!c || !d_dst || (!d_src && nbits)) return BZB200_E_ARG;
  if ((reinterpret_cast<uintptr_t>(d_dst) & 3) || (reinterpret_cast<uintptr_t>(d_src) & 3)) return BZB200_E_ARG;
  if (((dst_bit + nbits + 31) / 32) * 4 > dst_cap_bytes) {
    c->err = "bit_append: destination too small";
    return BZB200_E_ARG;
  }
  TRY(set_device(c));
  launch_bit_append(c->L, d_dst, dst_bit, d_src, nbits);
  return check_launch(c);
}

uint32_t bzb200_combine_crc(uint32_t seed, const uint32_t* crc, size_t n) {
  uint32_t c = seed;
  for (size_t i = 0; i < n; ++i) c = ((c << 1) | (c >> 31)) ^ crc[i];
  return c;
}

int bzb200_write_stream_header(bzb200_ctx* c, int level, uint8_t* d_out, size_t cap_bytes) {
  if (!c || !d_out || cap_bytes < 4 || (reinterpret_cast<uintptr_t>(d_out) & 3)) return BZB200_E_ARG;
  if (level < 1 || level > 9) return BZB200_E_LEVEL;
  TRY(set_device(c));
  launch_write_header(c->L, level, d_out);
  return check_launch(c);
}

int bzb200_write_stream_trailer(bzb200_ctx* c, uint8_t* d_out, size_t cap_bytes, uint64_t at_bit, uint32_t combined_crc,
                                size_t* total_bytes) {
  if (!c || !d_out || (reinterpret_cast<uintptr_t>(d_out) & 3)) return BZB200_E_ARG;
  const size_t total = (size_t)((at_bit + 80 + 7) / 8);
  if (((at_bit + 80 + 31) / 32) * 4 > cap_bytes) {
    c->err = "trailer: output buffer too small";
    return BZB200_E_ARG;
  }
  TRY(set_device(c));
  launch_write_trailer(c->L, d_out, at_bit, combined_crc);
  if (total_bytes) *total_bytes = total;
  return check_launch(c);
}

int bzb200_compress_device(bzb200_ctx* c, int level, const uint8_t* d_in, size_t n, uint8_t* d_out, size_t cap_bytes,
                           size_t* out_n) {
  if (!c || !d_out || !out_n) return BZB200_E_ARG;
  uint32_t nb = 0;
  TRY(bzb200_plan(c, level, d_in, n, &nb));
  TRY(bzb200_write_stream_header(c, level, d_out, cap_bytes));
  uint64_t end_bit = 32;
  if (nb) TRY(bzb200_encode_blocks(c, 0, nb, d_out, cap_bytes, 32, &end_bit));
  const uint32_t combined = bzb200_combine_crc(0, c->h_crc.data(), nb);
  TRY(bzb200_write_stream_trailer(c, d_out, cap_bytes, end_bit, combined, out_n));
  CK(c, cudaStreamSynchronize(c->stream));
  return BZB200_OK;
}

// Host buffers in, host buffers out.  The input is copied in segments on a copy stream; as soon as a segment has
// landed, the bytes from the last block cut up to the end of that segment are planned as a stream of their own and
// all of their blocks but the (still open) last one are encoded, while the next segments are still in flight —
// a block cut is a piece boundary, so RLE1 restarts there exactly as in the one-pass plan.  Finished output bytes
// go back to the host on a second copy stream while the next segment is compressed.
int bzb200_compress_host(bzb200_ctx* c, int level, const uint8_t* h_in, size_t n, uint8_t* h_out, size_t cap_bytes,
                         size_t* out_n) {
  if (!c || !h_out || !out_n || (!h_in && n)) return BZB200_E_ARG;
  if (level < 1 || level > 9) {
    c->err = "invalid level";
    return BZB200_E_LEVEL;
  }
  TRY(set_device(c));
  const size_t cap = bzb200_max_output_bytes(level, n);
  TRY(ensure(c, c->stage_in, n + 16));
  TRY(ensure(c, c->stage_out, cap));
  // Segment schedule: a small first segment (compute starts after ~2.5 ms of copying), then 1 GiB segments — every
  // extra segment costs ~5 ms of fixed per-batch latency on a B200, more than the copy time it hides.
  size_t seg = (size_t)1 << 30, first = (size_t)128 << 20;
  if (const char* e = getenv("BZB200_HOST_SEGMENT")) {
    unsigned long long v = strtoull(e, nullptr, 10);
    if (v >= (1u << 20)) seg = (size_t)v;
  }
  first = std::min(first, seg / 2);
  if (n <= 2 * first) {  // small input: one copy, one plan
    if (n) CK(c, cudaMemcpyAsync(c->stage_in.p, h_in, n, cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaMemsetAsync(c->stage_out.p, 0, cap, c->stream));
    size_t got = 0;
    TRY(bzb200_compress_device(c, level, ptr<uint8_t>(c->stage_in), n, ptr<uint8_t>(c->stage_out), cap, &got));
    if (got > cap_bytes) {
      c->err = "compress_host: output buffer too small: need " + std::to_string(got) + " bytes";
      return BZB200_E_ARG;
    }
    CK(c, cudaMemcpyAsync(h_out, c->stage_out.p, got, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    *out_n = got;
    return BZB200_OK;
  }
  if (!c->h2d_stream) CK(c, cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
  if (!c->d2h_stream) CK(c, cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
  // Every exit below — also the user-triggerable "output buffer too small" ones — leaves with the copy streams
  // drained: the caller may free h_in / h_out as soon as the call returns.
  const int rc = [&]() -> int {
  std::vector<size_t> ends;
  for (size_t e = first; e < n; e += seg) ends.push_back(e);
  if (ends.size() && n - ends.back() < seg / 4) ends.pop_back();  // no tiny tail segment
  ends.push_back(n);
  const size_t nseg = ends.size();
  while (c->seg_events.size() < nseg + 1) {
    cudaEvent_t ev;
    CK(c, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    c->seg_events.push_back(ev);
  }
  uint8_t* d_in = ptr<uint8_t>(c->stage_in);
  uint8_t* d_out = ptr<uint8_t>(c->stage_out);
  for (size_t s = 0; s < nseg; ++s) {
    const size_t lo = s ? ends[s - 1] : 0, len = ends[s] - lo;
    CK(c, cudaMemcpyAsync(d_in + lo, h_in + lo, len, cudaMemcpyHostToDevice, c->h2d_stream));
    CK(c, cudaEventRecord(c->seg_events[s], c->h2d_stream));
  }
  {  // size the RLE1 buffer for the longest sub-stream up front (a segment plus the open block carried into it)
    const size_t longest = std::min(n, seg + ((size_t)64 << 20));
    TRY(ensure(c, c->txt, longest + longest / 4 + 64));
  }
  CK(c, cudaMemsetAsync(d_out, 0, cap, c->stream));
  TRY(bzb200_write_stream_header(c, level, d_out, cap));
  size_t pos = 0, copied = 0;
  uint64_t bit = 32;
  uint32_t combined = 0;
  for (size_t s = 0; s < nseg; ++s) {
    const size_t avail = ends[s];
    const bool last = s + 1 == nseg;
    CK(c, cudaStreamWaitEvent(c->stream, c->seg_events[s], 0));
    uint32_t nb = 0;
    TRY(bzb200_plan(c, level, d_in + pos, avail - pos, &nb));
    const uint32_t nenc = last ? nb : (nb ? nb - 1 : 0);  // the last block of a segment is still open
    if (nenc) {
      TRY(bzb200_encode_blocks(c, 0, nenc, d_out, cap, bit, &bit));
      combined = bzb200_combine_crc(combined, c->h_crc.data(), nenc);
      pos += (size_t)c->h_in_off[nenc];
    }
    const size_t fin = (size_t)(bit / 8);  // bytes below `bit` are final (later blocks only OR at >= bit)
    if (fin > cap_bytes) {
      c->err = "compress_host: output buffer too small";
      return BZB200_E_ARG;
    }
    if (!last && fin > copied) {  // encode_blocks has synchronised: those bytes are in place
      CK(c, cudaMemcpyAsync(h_out + copied, d_out + copied, fin - copied, cudaMemcpyDeviceToHost, c->d2h_stream));
      copied = fin;
    }
  }
  size_t got = 0;
  TRY(bzb200_write_stream_trailer(c, d_out, cap, bit, combined, &got));
  if (got > cap_bytes) {
    c->err = "compress_host: output buffer too small: need " + std::to_string(got) + " bytes";
    return BZB200_E_ARG;
  }
  CK(c, cudaMemcpyAsync(h_out + copied, d_out + copied, got - copied, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  CK(c, cudaStreamSynchronize(c->d2h_stream));
  *out_n = got;
  return BZB200_OK;
  }();
  if (rc != BZB200_OK) {
    cudaStreamSynchronize(c->h2d_stream);
    cudaStreamSynchronize(c->d2h_stream);
    cudaStreamSynchronize(c->stream);
  }
  return rc;
}

// ------------------------------------------------------------------ instrumentation
int bzb200_debug_stage(bzb200_ctx* c, uint32_t block, int field, void* host_dst, size_t cap_elems, size_t* count) {
  if (!c || !c->planned) return BZB200_E_STATE;
  if (block < c->batch_b0 || block >= c->batch_b0 + c->batch_nb) {
    c->err = "debug_stage: block is not in the most recent batch";
    return BZB200_E_ARG;
  }
  TRY(set_device(c));
  CK(c, cudaStreamSynchronize(c->stream));
  const uint32_t i = block - c->batch_b0;
  const BlockDesc& d = c->h_desc[i];
  uint32_t meta[8];
  CK(c, cudaMemcpy(meta, ptr<uint32_t>(c->meta) + (size_t)i * 8, sizeof(meta), cudaMemcpyDeviceToHost));
  uint32_t mc = 0;
  CK(c, cudaMemcpy(&mc, ptr<uint32_t>(c->mtf_count) + i, 4, cudaMemcpyDeviceToHost));
  const void* src = nullptr;
  size_t n = 0, esz = 1;
  switch (field) {
    case BZB200_F_RLE: src = v_txt(c) + c->h_rle_off[block]; n = d.n; esz = 1; break;
    case BZB200_F_RANK: src = ptr<uint32_t>(c->rank) + d.off; n = d.n; esz = 4; break;
    case BZB200_F_LAST: src = ptr<uint8_t>(c->last) + d.off; n = d.n; esz = 1; break;
    case BZB200_F_MTF: src = ptr<uint16_t>(c->sym) + d.symoff; n = mc; esz = 2; break;
    case BZB200_F_FREQ: src = ptr<uint32_t>(c->freq) + (size_t)i * MAX_ALPHA; n = meta[0]; esz = 4; break;
    case BZB200_F_SEL: src = ptr<uint8_t>(c->sel) + (size_t)i * MAX_SELECTORS; n = meta[2]; esz = 1; break;
    case BZB200_F_LEN0: case BZB200_F_LEN1: case BZB200_F_LEN2: case BZB200_F_LEN3: case BZB200_F_LEN4: {
      // gather [ngroups][alpha] out of the padded [6][258] layout
      const int slot = field - BZB200_F_LEN0;
      const size_t alpha = meta[0], ng = meta[1];
      std::vector<uint8_t> tmp((size_t)MAX_GROUPS * MAX_ALPHA);
      CK(c, cudaMemcpy(tmp.data(), ptr<uint8_t>(c->lens) + (((size_t)i * 5 + slot) * MAX_GROUPS) * MAX_ALPHA, tmp.size(),
                       cudaMemcpyDeviceToHost));
      n = ng * alpha;
      if (count) *count = n;
      uint8_t* o = (uint8_t*)host_dst;
      size_t k = 0;
      for (size_t t = 0; t < ng; ++t)
        for (size_t s = 0; s < alpha; ++s, ++k)
          if (o && k < cap_elems) o[k] = tmp[t * MAX_ALPHA + s];
      return BZB200_OK;
    }
    case BZB200_F_INFO: {
      uint64_t info[16] = {0};
      uint32_t op = 0, rounds = 0, st[4] = {0, 0, 0, 0}, iu[8];
      CK(c, cudaMemcpy(&op, ptr<uint32_t>(c->origptr) + i, 4, cudaMemcpyDeviceToHost));
      CK(c, cudaMemcpy(&rounds, ptr<uint32_t>(c->rounds) + i, 4, cudaMemcpyDeviceToHost));
      CK(c, cudaMemcpy(st, ptr<uint32_t>(c->stats) + (size_t)i * 4, 16, cudaMemcpyDeviceToHost));
      CK(c, cudaMemcpy(iu, ptr<uint32_t>(c->inuse) + (size_t)block * 8, 32, cudaMemcpyDeviceToHost));
      uint64_t bb[2] = {0, 0};
      CK(c, cudaMemcpy(bb, ptr<uint64_t>(c->blockbit) + i, 16, cudaMemcpyDeviceToHost));
      info[0] = c->h_in_off[block]; info[1] = c->h_in_off[block + 1]; info[2] = d.n; info[3] = c->h_crc[block];
      info[4] = op; info[5] = mc; info[6] = meta[0]; info[7] = meta[1]; info[8] = meta[2];
      info[9] = bb[0]; info[10] = bb[1]; info[11] = rounds; info[12] = st[3];
      info[13] = meta[5];  // package-merge fallbacks taken
      info[14] = meta[6];  // device-side error flags
      n = 16;
      if (count) *count = 16 + 8;
      uint64_t* o = (uint64_t*)host_dst;
      for (size_t k = 0; k < 16 && o && k < cap_elems; ++k) o[k] = info[k];
      for (size_t k = 0; k < 8 && o && 16 + k < cap_elems; ++k) o[16 + k] = iu[k];
      return BZB200_OK;
    }
    default: return BZB200_E_ARG;
  }
  if (count) *count = n;
  size_t ncopy = std::min(n, cap_elems);
  if (host_dst && ncopy) CK(c, cudaMemcpy(host_dst, src, ncopy * esz, cudaMemcpyDeviceToHost));
  return BZB200_OK;
}

int bzb200_profile(bzb200_ctx* c, int on) {
  if (!c) return BZB200_E_ARG;
  if (on) {
    c->L.clear_profile();
    c->L.profiling = true;
  } else {
    c->L.resolve();
    c->L.profiling = false;
  }
  return BZB200_OK;
}
int bzb200_profile_count(bzb200_ctx* c) {
  if (!c) return 0;
  c->L.resolve();
  return (int)c->L.recs.size();
}
int bzb200_profile_get(bzb200_ctx* c, int i, const char** name, uint64_t* launches, double* total_ms) {
  if (!c || i < 0 || i >= (int)c->L.recs.size()) return BZB200_E_ARG;
  if (name) *name = c->L.recs[i].name;
  if (launches) *launches = c->L.recs[i].launches;
  if (total_ms) *total_ms = c->L.recs[i].ms;
  return BZB200_OK;
}
uint64_t bzb200_launch_count(const bzb200_ctx* c) { return c ? c->L.launches : 0; }
int bzb200_sort_stats(const bzb200_ctx* c, uint32_t* rounds, uint32_t* radix_passes, uint64_t* elems_sorted) {
  if (!c) return BZB200_E_ARG;
  if (rounds) *rounds = c->bstats.rounds;
  if (radix_passes) *radix_passes = c->bstats.radix_passes;
  if (elems_sorted) *elems_sorted = c->bstats.elems_sorted;
  return BZB200_OK;
}
int bzb200_path_stats(const bzb200_ctx* c, uint64_t* out, size_t cap) {
  if (!c || !out) return BZB200_E_ARG;
  const uint64_t v[8] = {c->bstats.rounds, c->bstats.radix_passes, c->bstats.elems_sorted, c->bstats.radix_elem_passes,
                         c->bstats.local_elems, c->stat_rle, c->stat_mtf, 0};
  for (size_t i = 0; i < cap && i < 8; ++i) out[i] = v[i];
  return BZB200_OK;
}

}  // extern "C"
