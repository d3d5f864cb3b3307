// slice_plan.cu — the K1 plan for ONE SLICE of a stream (include/bzb200.h section 2b): what a rank of a sharded run, or
// one GPU of the in-library multi-GPU engine (mgpu.cu), executes when it holds only its part of the input.
//
// The reference cuts blocks in one sequential pass (EncoderInner::next / write_rle, src/bzip2/encoder.rs:671-716, cut
// test :692-696).  Sharded, the same cuts come out of per-slice work plus three tiny exchanges between the slices
// (done by the caller: NCCL all-gathers in rust-compression_b200/sharded.py, host memory in mgpu.cu):
//   1. slice_begin   last run head of every own tile            -> exchange: last run head inside the slice (1 word)
//   2. slice_counts  RLE1 bytes every own tile emits             -> exchange: bytes the slice emits       (1 word)
//   3. slice_prefix  emitted offset of every own tile
//   4. slice_windows the cut windows (k1_cut_windows) whose centre falls into the slice, for the phase that starts at
//                    emitted offset x0                            -> exchange: the window rows (2 KB each)
//      bzb200_cut_walk (host) follows the chain through the rows exactly like k1_cut_walk does on one GPU; a drift
//      that leaves the 256-offset window starts another phase at that cut
//   5. slice_set_blocks: every slice receives the same block table; a slice then encodes the blocks that START inside
//      it.  Its last block ends in the neighbour's bytes: slice_need says how far, the caller fetches that tail
//      (P2P / host copy) behind the slice and slice_extend brings the tile summaries up to it.
// No O(input) array is exchanged and no rank scans more than its own tiles (+ the tail of one block).
#include "host_ctx.h"

namespace {

constexpr uint64_t HALO = 65536;  // bytes behind the slice that must be resident before slice_windows (a window reaches
                                  // 260 emitted bytes past its centre = at most 13.3 kB of input: 51 input bytes/emitted)

inline uint64_t tile_of(uint64_t byte) { return byte / k1_tile_bytes(); }
inline uint64_t tiles_up(uint64_t byte) { return (byte + k1_tile_bytes() - 1) / k1_tile_bytes(); }

// tiles whose bytes AND the byte after them are resident (a tile's last thread reads in[end]); all tiles at the end of
// the stream
inline uint64_t usable_tiles(uint64_t avail, uint64_t N) { return avail >= N ? tiles_up(N) : tile_of(avail ? avail - 1 : 0); }

int ensure_tiles(bzb200_ctx* c, uint64_t ntl) {
  TRY(ensure(c, c->tile_head, (ntl + 1) * 8));
  TRY(ensure(c, c->tile_carry, (ntl + 1) * 8));
  TRY(ensure(c, c->tile_cnt, (ntl + 1) * 4));
  TRY(ensure(c, c->tile_E, (ntl + 2) * 8));
  return BZB200_OK;
}

// Tile summaries of the new tiles [told, tb) behind the ones the slice already has: the per-tile kernels run on the new
// tiles only; the two single-CTA scans run again over [t0, tb) (they rewrite the old entries with the same values).
int summarise_more(bzb200_ctx* c, uint64_t told, uint64_t tb) {
  const uint64_t t0 = c->sl_t0;
  launch_k1_slice_heads(c->L, v_in(c), c->n_in, told, tb, v_head(c));
  launch_k1_slice_scan_carry(c->L, t0, tb, c->sl_carry_in, v_head(c), v_carry(c));
  launch_k1_slice_tile_counts(c->L, v_in(c), c->n_in, told, tb, v_carry(c), v_cnt(c));
  launch_k1_slice_prefix(c->L, t0, tb, c->sl_E_lo, v_cnt(c), v_E(c));
  return check_launch(c);
}

}  // namespace

extern "C" {

size_t bzb200_slice_halo_bytes(void) { return (size_t)HALO; }
uint32_t bzb200_cut_window(void) { return k1_cut_window(); }

int bzb200_slice_begin(bzb200_ctx* c, int level, uint64_t N, uint64_t lo, uint64_t hi, const uint8_t* d_lo,
                       uint64_t avail_hi, uint64_t reserve_hi, int64_t* last_head) {
  if (!c || !last_head) return BZB200_E_ARG;
  if (level < 1 || level > 9) {
    c->err = "invalid level";
    return BZB200_E_LEVEL;
  }
  if (lo >= hi || hi > N || lo % k1_tile_bytes() || (hi % k1_tile_bytes() && hi != N) || !d_lo ||
      (reinterpret_cast<uintptr_t>(d_lo) & 15) || avail_hi < std::min<uint64_t>(N, hi + HALO) || avail_hi > N) {
    c->err = "slice_begin: bad slice (tile-aligned [lo,hi), 16-byte aligned d_lo, halo resident)";
    return BZB200_E_ARG;
  }
  TRY(set_device(c));
  c->planned = false;
  c->plan_open = false;
  c->sliced = true;
  c->level = level;
  c->T = (uint32_t)level * 100000u - 19u;  // encoder.rs:186
  c->d_in = nullptr;
  c->n_in = N;
  c->sl_d_lo = d_lo;
  c->sl_lo = lo;
  c->sl_hi = hi;
  c->sl_avail = avail_hi;
  c->sl_t0 = tile_of(lo);
  c->sl_t1 = tiles_up(hi);
  c->sl_tn = usable_tiles(avail_hi, N);
  c->sl_carry_in = -1;
  c->sl_E_lo = c->sl_E_hi = c->sl_E_tot = 0;
  c->txt_origin = 0;
  c->nblocks = 0;
  c->max_block_len = 0;
  c->h_in_off.assign(1, 0);
  c->h_rle_off.assign(1, 0);
  c->h_crc.clear();
  c->batch_nb = 0;
  c->prep_lo = c->prep_hi = 0;
  c->crc_all = false;
  reserve_hi = std::min<uint64_t>(N, std::max(reserve_hi, avail_hi));
  TRY(ensure_tiles(c, tiles_up(reserve_hi) - c->sl_t0));
  TRY(ensure(c, c->sl_sum, 64));
  launch_k1_slice_heads(c->L, v_in(c), N, c->sl_t0, c->sl_tn, v_head(c));
  launch_k1_slice_summary(c->L, ptr<long long>(c->tile_head), nullptr, c->sl_t1 - c->sl_t0, ptr<uint64_t>(c->sl_sum));
  TRY(check_launch(c));
  uint64_t v = 0;
  CK(c, cudaMemcpyAsync(&v, c->sl_sum.p, 8, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  *last_head = (int64_t)v;
  c->sl_stage = 1;
  return BZB200_OK;
}

int bzb200_slice_counts(bzb200_ctx* c, int64_t carry_in, uint64_t* emitted) {
  if (!c || !emitted) return BZB200_E_ARG;
  if (!c->sliced || c->sl_stage < 1) return BZB200_E_STATE;
  TRY(set_device(c));
  c->sl_carry_in = carry_in;
  launch_k1_slice_counts(c->L, v_in(c), c->n_in, c->sl_t0, c->sl_tn, carry_in, v_head(c), v_carry(c), v_cnt(c));
  launch_k1_slice_summary(c->L, nullptr, ptr<uint32_t>(c->tile_cnt), c->sl_t1 - c->sl_t0, ptr<uint64_t>(c->sl_sum));
  TRY(check_launch(c));
  uint64_t v[2] = {0, 0};
  CK(c, cudaMemcpyAsync(v, c->sl_sum.p, 16, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  *emitted = v[1];
  c->sl_E_hi = v[1];  // relative until slice_prefix
  c->sl_stage = 2;
  return BZB200_OK;
}

int bzb200_slice_prefix(bzb200_ctx* c, uint64_t E_lo, uint64_t E_tot) {
  if (!c) return BZB200_E_ARG;
  if (!c->sliced || c->sl_stage < 2) return BZB200_E_STATE;
  TRY(set_device(c));
  const uint64_t emitted = c->sl_stage == 2 ? c->sl_E_hi : c->sl_E_hi - c->sl_E_lo;
  c->sl_E_lo = E_lo;
  c->sl_E_hi = E_lo + emitted;
  c->sl_E_tot = E_tot;
  if (c->sl_E_hi > E_tot) {
    c->err = "slice_prefix: inconsistent emitted totals";
    return BZB200_E_ARG;
  }
  launch_k1_slice_prefix(c->L, c->sl_t0, c->sl_tn, E_lo, v_cnt(c), v_E(c));
  c->sl_stage = 3;
  return check_launch(c);
}

int bzb200_slice_windows(bzb200_ctx* c, uint64_t x0, uint64_t* j0, uint32_t* nj, const uint64_t** d_F) {
  if (!c || !j0 || !nj || !d_F) return BZB200_E_ARG;
  if (!c->sliced || c->sl_stage < 3) return BZB200_E_STATE;
  TRY(set_device(c));
  const uint64_t T = c->T;
  // centres x0 + (j+1) T that fall into (E_lo, E_hi]
  uint64_t lo_j = c->sl_E_lo >= x0 ? (c->sl_E_lo - x0) / T : 0;
  uint64_t cnt = 0;
  if (c->sl_E_hi >= x0 + T) {
    const uint64_t hi_j = (c->sl_E_hi - x0) / T - 1;
    if (hi_j >= lo_j) cnt = hi_j - lo_j + 1;
  }
  if (cnt > (1u << 22)) return BZB200_E_ARG;
  *j0 = lo_j;
  *nj = (uint32_t)cnt;
  *d_F = nullptr;
  if (cnt == 0) return BZB200_OK;
  TRY(ensure(c, c->sl_F, cnt * k1_cut_window() * 8));
  CK(c, cudaMemsetAsync(c->sl_F.p, 0, cnt * k1_cut_window() * 8, c->stream));  // a row entry nobody wrote reads as 0: the walk rejects it
  launch_k1_slice_windows(c->L, v_in(c), c->n_in, c->T, v_carry(c), v_E(c), c->sl_t0, c->sl_t1 - 1, c->sl_tn, c->sl_E_tot, x0,
                          lo_j, (uint32_t)cnt, ptr<uint64_t>(c->sl_F));
  *d_F = ptr<uint64_t>(c->sl_F);
  return check_launch(c);
}

// Host: the chain through the window rows of one phase.  The restatement of k1_cut_walk (k1_rle.cu) in scalar form;
// F holds rows j = 0 .. K-1 of the phase that starts at x0 = state[1].  state[4] = {blocks cut so far, emitted offset
// where the open block starts, done, longest block}; in_off / rle_off have room for max_blocks + 1 entries.
int bzb200_cut_walk(const uint64_t* F, uint64_t K, uint32_t T, uint64_t Etot, uint64_t N, uint32_t max_blocks,
                    uint64_t* state, uint64_t* in_off, uint64_t* rle_off, uint32_t* nblocks, uint32_t* max_block_len) {
  if (!state || !in_off || !rle_off || (!F && K)) return BZB200_E_ARG;
  if (state[2]) return BZB200_OK;
  constexpr uint64_t LAST = 1ull << 63;
  const uint32_t W = k1_cut_window();
  const uint64_t x0 = state[1];
  uint64_t k = state[0], S = x0, maxlen = state[3], d = 0;
  bool done = false;
  if (k == 0) { in_off[0] = 0; rle_off[0] = 0; }
  for (uint64_t j = 0;; ++j) {
    const uint64_t center = x0 + (j + 1) * (uint64_t)T;
    const bool reach = center + d <= Etot && k + 2 <= max_blocks;  // otherwise the rest is the last block
    if (!reach) { done = true; break; }
    if (j >= K) break;  // table exhausted: another phase starts at S
    const uint64_t v = F[j * W + d];
    if (v & LAST) { done = true; break; }  // the piece that reaches T is the last piece of the input (encoder.rs:729-739)
    const uint64_t rel = v & 0xFFFFull;
    if (rel < d || rel > d + 4) return BZB200_E_INTERNAL;  // f(x) - x <= 4: anything else is a corrupt row
    k += 1;
    in_off[k] = (v >> 16) & 0x7FFFFFFFFFFFull;
    if (in_off[k] <= in_off[k - 1] || in_off[k] > N) return BZB200_E_INTERNAL;
    rle_off[k] = center + rel;
    maxlen = std::max(maxlen, center + rel - S);
    S = center + rel;
    if (rel >= W) { ++j; break; }  // drift left the window: the next phase starts at S
    d = rel;
  }
  state[0] = k;
  state[1] = S;
  state[3] = maxlen;
  if (done) {
    state[2] = 1;
    const uint64_t nb = k + 1;
    in_off[nb] = N;
    rle_off[nb] = Etot;
    maxlen = std::max(maxlen, Etot - S);
    if (nblocks) *nblocks = (uint32_t)nb;
    if (max_block_len) *max_block_len = (uint32_t)std::min<uint64_t>(0xFFFFFFFFu, maxlen);
  }
  return BZB200_OK;
}

int bzb200_slice_set_blocks(bzb200_ctx* c, uint32_t nblocks, const uint64_t* in_off, const uint64_t* rle_off,
                            uint32_t max_block_len) {
  if (!c || !in_off || !rle_off || nblocks == 0) return BZB200_E_ARG;
  if (!c->sliced || c->sl_stage < 3) return BZB200_E_STATE;
  if (in_off[0] != 0 || in_off[nblocks] != c->n_in || rle_off[nblocks] != c->sl_E_tot) {
    c->err = "slice_set_blocks: the block table does not cover the stream";
    return BZB200_E_ARG;
  }
  if (max_block_len > (uint32_t)c->level * 100000u || max_block_len > MAX_BLOCK) {
    c->err = "block longer than level*100000";
    return BZB200_E_INTERNAL;
  }
  TRY(set_device(c));
  c->nblocks = nblocks;
  c->max_block_len = max_block_len;
  c->h_in_off.assign(in_off, in_off + nblocks + 1);
  c->h_rle_off.assign(rle_off, rle_off + nblocks + 1);
  c->h_crc.assign(nblocks, 0);
  TRY(ensure(c, c->in_off, ((size_t)nblocks + 1) * 8));
  TRY(ensure(c, c->rle_off, ((size_t)nblocks + 1) * 8));
  TRY(ensure(c, c->crc, (size_t)nblocks * 4));
  TRY(ensure(c, c->inuse, (size_t)nblocks * 32));
  CK(c, cudaMemcpyAsync(c->in_off.p, c->h_in_off.data(), ((size_t)nblocks + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  CK(c, cudaMemcpyAsync(c->rle_off.p, c->h_rle_off.data(), ((size_t)nblocks + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));  // the host vectors may be reassigned by the next plan
  c->planned = true;
  return BZB200_OK;
}

// The blocks that START inside the slice, and how far the input must be resident to encode them (whole tiles, so the
// answer is a tile boundary or N).
int bzb200_slice_blocks(const bzb200_ctx* c, uint32_t* b0, uint32_t* b1, uint64_t* need_hi) {
  if (!c || !b0 || !b1 || !need_hi) return BZB200_E_ARG;
  if (!c->sliced || !c->planned) return BZB200_E_STATE;
  const auto& off = c->h_in_off;
  const uint32_t lo = (uint32_t)(std::lower_bound(off.begin(), off.begin() + c->nblocks, c->sl_lo) - off.begin());
  const uint32_t hi = (uint32_t)(std::lower_bound(off.begin(), off.begin() + c->nblocks, c->sl_hi) - off.begin());
  *b0 = lo;
  *b1 = hi;
  *need_hi = hi > lo ? std::min<uint64_t>(c->n_in, tiles_up(off[hi]) * k1_tile_bytes() + (off[hi] < c->n_in ? 1 : 0)) : c->sl_lo;
  return BZB200_OK;
}

int bzb200_slice_extend(bzb200_ctx* c, uint64_t avail_hi) {
  if (!c) return BZB200_E_ARG;
  if (!c->sliced || c->sl_stage < 3) return BZB200_E_STATE;
  if (avail_hi > c->n_in) return BZB200_E_ARG;
  if (avail_hi <= c->sl_avail) return BZB200_OK;
  TRY(set_device(c));
  c->sl_avail = avail_hi;
  const uint64_t tn = usable_tiles(avail_hi, c->n_in);
  if (tn <= c->sl_tn) return BZB200_OK;
  {  // growing the tile arrays would drop their contents: the caller reserved the room in slice_begin (reserve_hi)
    const uint64_t ntl = tn - c->sl_t0;
    if (c->tile_head.cap < (ntl + 1) * 8 || c->tile_carry.cap < (ntl + 1) * 8 || c->tile_cnt.cap < (ntl + 1) * 4 ||
        c->tile_E.cap < (ntl + 2) * 8) {
      c->err = "slice_extend: beyond the range reserved in slice_begin";
      return BZB200_E_ARG;
    }
  }
  const uint64_t told = c->sl_tn;
  c->sl_tn = tn;
  return summarise_more(c, told, tn);
}

}  // extern "C"

// RLE1 buffer of a sliced context: room for the emitted bytes of blocks [b0, b1) plus the rest of their first and last
// tiles (called by bzb200_encode_blocks).
int slice_reserve_txt(bzb200_ctx* c, uint32_t b0, uint32_t b1) {
  if (b0 >= b1) return BZB200_OK;
  // k1_scatter writes whole tiles: from the emitted offset of the tile holding in_off[b0] to that of the tile after
  // in_off[b1]; a tile emits at most 1.25 bytes per input byte
  const uint64_t span = c->h_rle_off[b1] - c->h_rle_off[b0];
  const uint64_t slack = 2 * (uint64_t)k1_tile_bytes() * 5 / 4 + 256;
  CK(c, cudaStreamSynchronize(c->stream));
  TRY(ensure(c, c->txt, span + 2 * slack + 256));
  // a multiple of 256: kernels that read txt with 128-bit loads align on the INDEX (k1_inuse), so the shifted base must
  // keep the alignment of the allocation
  c->txt_origin = (c->h_rle_off[b0] >= slack ? c->h_rle_off[b0] - slack : 0) & ~(uint64_t)255;
  c->prep_lo = c->prep_hi = 0;
  return BZB200_OK;
}
