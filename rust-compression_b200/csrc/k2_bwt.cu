// k2_bwt.cu — K2: the BWT as a batched GPU prefix-doubling sort of the cyclic rotations of every block.
//
// Replaces suffix_array::sais::bwt (src/suffix_array/sais.rs:266-272 = least-rotation pre-pass :12-68 + cyclic
// SA-IS :127-264) by result, not by method.  Order to reproduce (pinned by
// tests/test_oracle.py::test_bwt_matches_rotation_model): rotations in lexicographic order; equal rotations
// (periodic block) by DESCENDING (pos - shift) mod n, shift = smallest start of a minimal rotation;
// origPtr = rank of rotation 0 (src/bzip2/encoder.rs:332-334).
//
// State per block:  SA[slot] = pos | HEAD | BIG | SINGLE   (rotations in the order established so far; a *group* is
//                                a maximal run of slots whose rotations are still tied; HEAD marks its first slot)
//                   rank[pos] = slot of the head of pos's group (| RESOLVED once the group is a singleton)
// Round 0  : 64-bit elements [55:20] initial key | [19:0] pos; the key is the first S symbols of the rotation, a symbol
//            being the rank of the byte among the block's in-use bytes (text: 6 symbols x 6 bits, 4 one-sweep passes of
//            9-bit digits; byte alphabets: 5 bytes, 5 passes of 8 bits — see key_mode).  Pass 0 builds the elements
//            from the text; the digit histograms of all passes come from one histogram of symbol pairs.  Then
//            regroup -> SA, rank.
// Round r  : step h = S, 2S, 4S, ...  key of a tied rotation = rank[(pos + h) mod n]:
//   k2_gather         every unresolved slot fetches its key (slot order, coalesced SA read, L2-resident rank gather),
//                     reads its group head off the HEAD flags of its tile and is appended to the tile's dense work
//                     list; members of BIG groups (> LOCAL_MAX slots) and of sparse blocks are emitted as
//                     [59:40] g | [39:20] key | [19:0] pos
//   k2_local_sort     dense lists: one CTA per 2048-slot tile sorts every group it owns inside shared memory (key-range
//                     bucket split for larger groups, then enumeration sort on (group, key)), SA and rank in place
//   k2_local_sort_rx  sparse lists: a CTA packs several tiles into one window and sorts it by (group, key) with four
//                     stable 8-bit counting passes in shared memory
//   BIG groups        LSD radix sort of the emitted elements + regroup (the round-0 machinery on a shorter list; the
//                     8-bit passes run persistent and TMA-fed: k2_os_scatter_pf)
//   k2_round_finalize per-block state machine: done / periodic fix-up pending / active / sparse
// Fix-up   : a round that splits nothing means the block is periodic and the groups are the sets of equal
//            rotations; one more round with key = n-1-((pos-shift) mod n) applies the reference's tie-break.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace bzb {

constexpr int RS_NT = 256;
constexpr int RS_IPT = 16;
constexpr int RS_TILE = RS_NT * RS_IPT;  // 4096 elements per CTA
constexpr int RS_WARPS = RS_NT / 32;
constexpr int KEY_LO = 20;                  // key field = bits 20..59
constexpr uint64_t POS_MASK = 0xFFFFFull;

uint32_t bwt_tile_elems() { return RS_TILE / 2; }  // sizing unit of the per-tile arrays (smallest tile of any pass variant)

// ------------------------------------------------------------------ one-sweep LSD radix pass (per-block segments)
// Digit histograms of all passes are taken once (k2_os_hist* / k2_pair_hist + k2_digit_offsets), turned into bucket
// offsets, and each pass is then ONE kernel: a CTA ranks its tile (warp match + per-warp counters), publishes the
// tile's digit counts, obtains the counts of the preceding tiles of its block by decoupled look-back over the status
// words of those tiles, and scatters through shared memory.  Tiles of a block are numbered by a ticket, so a tile only
// ever waits for tiles that are already running.  The grid is (block, tile): CTAs in flight belong to different
// blocks, which keeps the look-back chains short.
//
// Initial sort keys.  The key of rotation pos is the first S symbols of the rotation, BITS bits each, where a symbol is
// the rank of the byte among the block's in-use bytes (order preserving).  The batch's largest alphabet picks the mode:
//   > 128 bytes in use : BITS 8, S 5, 5 passes of 8-bit digits  (a digit is a byte: every pass has the byte histogram)
//   65..128            : BITS 7, S 5, 4 passes of 9-bit digits
//   17..64 (text)      : BITS 6, S 6, 4 passes of 9-bit digits  (one pass less AND one symbol more than bytes give)
//   <= 16              : BITS 4, S 8, 4 passes of 8-bit digits
// In the compact modes a digit straddles at most two neighbouring symbols of the key, and every pair of neighbouring
// text symbols is digit p of exactly one rotation, so all digit histograms follow from ONE histogram of symbol pairs
// (k2_pair_hist, 1 B read per element) — no pass over the keys.
constexpr uint32_t ST_AGG = 1u << 20, ST_PREFIX = 2u << 20, ST_VAL = 0xFFFFFu;
constexpr int OS_NT = 256, OS_IPT = 8, OS_MINB = 4;  // 2 048-element tiles, 4 CTAs per SM: the fastest geometry measured
constexpr int OS_TILE = OS_NT * OS_IPT;
constexpr int OS_BINS_MAX = 512;   // row pitch of the histogram / status arrays
constexpr int OS_PASSES_MAX = 5;

template <int BINS>
struct OsSmemT {
  static constexpr int WARPS = OS_NT / 32;
  uint64_t stage[OS_TILE];
  uint32_t wcnt[WARPS][BINS];
  uint32_t dstart[BINS];
  int goff[BINS];
  uint32_t ws[OS_NT / 32 + 1];
  uint32_t tile;
  uint8_t lut[256];
};

struct KeyMode {
  int bits;    // bits per symbol
  int syms;    // symbols in the initial key (= step h of the first doubling round)
  int wbits;   // digit width of a pass
  int passes;
};
static KeyMode key_mode(uint32_t max_alpha) {
  if (max_alpha == 0 || max_alpha > 128) return {8, 5, 8, 5};
  if (max_alpha > 64) return {7, 5, 9, 4};
  if (max_alpha > 16) return {6, 6, 9, 4};
  return {4, 8, 8, 4};
}

// byte -> rank among the block's in-use bytes
__device__ __forceinline__ void build_sym_lut(const uint32_t* __restrict__ inuse8, uint8_t* lut) {
  for (int v = threadIdx.x; v < 256; v += blockDim.x) {
    uint32_t r = 0;
#pragma unroll
    for (int wd = 0; wd < 8; ++wd) {
      const uint32_t m = inuse8[wd];
      if (wd < (v >> 5)) r += __popc(m);
      else if (wd == (v >> 5)) r += __popc(m & ((1u << (v & 31)) - 1u));
    }
    lut[v] = (uint8_t)r;
  }
}

// Byte histogram of a block's text = digit histogram of every pass of the byte-key sort (every byte of the block is
// digit p of exactly one rotation).  grid (chunks, nb); hist[b][p][d] accumulated with global atomics.
__global__ void __launch_bounds__(256) k2_os_hist_txt(const uint8_t* __restrict__ txt, const BlockDesc* __restrict__ desc,
                                                      uint32_t* __restrict__ hist, uint32_t* __restrict__ cnt,
                                                      uint32_t chunk) {
  __shared__ uint32_t h[8][256];
  const BlockDesc d = desc[blockIdx.y];
  const uint32_t lo = blockIdx.x * chunk;
  if (lo >= d.n) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) cnt[blockIdx.y] = d.n;  // list length of the initial sort
  const uint32_t hi = min(d.n, lo + chunk);
  for (int i = threadIdx.x; i < 8 * 256; i += 256) (&h[0][0])[i] = 0;
  __syncthreads();
  const uint8_t* t = txt + d.off;
  uint32_t* hw = h[(threadIdx.x >> 5) & 7];
  for (uint32_t i = lo + threadIdx.x; i < hi; i += 256) atomicAdd(&hw[t[i]], 1u);
  __syncthreads();
  uint32_t v = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) v += h[k][threadIdx.x];
  if (v) {
    uint32_t* o = hist + (uint64_t)blockIdx.y * OS_PASSES_MAX * OS_BINS_MAX + threadIdx.x;
#pragma unroll
    for (int p = 0; p < 5; ++p) atomicAdd(o + p * OS_BINS_MAX, v);
  }
}

// Histogram of the pairs (symbol at i, symbol at i+1 cyclic) of a block, symbols = ranks among the in-use bytes.
// grid (chunks, nb); pair[b][x << BITS | y] accumulated with global atomics (the table must be zero on entry).
template <int BITS>
__global__ void __launch_bounds__(256) k2_pair_hist(const uint8_t* __restrict__ txt, const BlockDesc* __restrict__ desc,
                                                    const uint32_t* __restrict__ inuse, uint32_t* __restrict__ pair,
                                                    uint32_t* __restrict__ cnt, uint32_t chunk) {
  extern __shared__ __align__(16) uint32_t ph_raw[];  // [1 << 2 BITS] counters, then the 256-byte symbol table
  constexpr int NP = 1 << (2 * BITS);
  uint8_t* lut = reinterpret_cast<uint8_t*>(ph_raw + NP);
  const BlockDesc d = desc[blockIdx.y];
  const uint32_t lo = blockIdx.x * chunk;
  if (lo >= d.n) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) cnt[blockIdx.y] = d.n;  // list length of the initial sort
  const uint32_t hi = min(d.n, lo + chunk);
  for (int i = threadIdx.x; i < NP; i += 256) ph_raw[i] = 0;
  build_sym_lut(inuse + blockIdx.y * 8, lut);
  __syncthreads();
  const uint8_t* t = txt + d.off;
  for (uint32_t i = lo + threadIdx.x; i < hi; i += 256) {
    const uint32_t x = lut[t[i]], y = lut[t[i + 1 == d.n ? 0 : i + 1]];
    atomicAdd(&ph_raw[(x << BITS) | y], 1u);
  }
  __syncthreads();
  uint32_t* o = pair + (uint64_t)blockIdx.y * NP;
  for (int i = threadIdx.x; i < NP; i += 256) {
    const uint32_t v = ph_raw[i];
    if (v) atomicAdd(o + i, v);
  }
}

// Bucket offsets of every pass of the compact-key sort from the pair histogram: digit p covers key bits
// [p W, p W + W); its lowest bit lies in key symbol k0 = p W / BITS (counted from the LAST symbol of the key) at bit
// o = p W - k0 BITS, so digit = ((x << BITS | y) >> o) & (2^W - 1) with (x, y) = key symbols (k0 + 1, k0) — two
// neighbouring text symbols, x first; above the first symbol of the key (k0 + 1 == S) x is 0.  One CTA per block.
template <int BITS, int S, int W>
__global__ void __launch_bounds__(512) k2_digit_offsets(const uint32_t* __restrict__ pair, uint32_t* __restrict__ hist) {
  constexpr int NP = 1 << (2 * BITS), BINS = 1 << W, P = (BITS * S + W - 1) / W;
  static_assert(BINS <= 512 && P <= OS_PASSES_MAX, "histogram row pitch");
  __shared__ uint32_t h[BINS];
  __shared__ uint32_t ws[512 / 32 + 1];
  const uint32_t* pr = pair + (uint64_t)blockIdx.x * NP;
  uint32_t* out = hist + (uint64_t)blockIdx.x * OS_PASSES_MAX * OS_BINS_MAX;
  for (int p = 0; p < P; ++p) {
    for (int i = threadIdx.x; i < BINS; i += 512) h[i] = 0;
    __syncthreads();
    const int k0 = p * W / BITS, o = p * W - k0 * BITS;
    const bool top = k0 + 1 >= S;
    for (int i = threadIdx.x; i < NP; i += 512) {
      const uint32_t c = pr[i];
      if (c) {
        const uint32_t pv = top ? (uint32_t)(i & ((1 << BITS) - 1)) : (uint32_t)i;
        atomicAdd(&h[(pv >> o) & (BINS - 1)], c);
      }
    }
    __syncthreads();
    const uint32_t v = threadIdx.x < BINS ? h[threadIdx.x] : 0u;
    const uint32_t ex = cta_excl_scan_add<512>(v, ws, nullptr);
    if (threadIdx.x < BINS) out[p * OS_BINS_MAX + threadIdx.x] = ex;
    __syncthreads();
  }
}

// Digit histograms of the five passes from the 64-bit elements themselves (rounds).  grid (chunks, nb).
__global__ void __launch_bounds__(256) k2_os_hist(const uint64_t* __restrict__ src, const BlockDesc* __restrict__ desc,
                                                  const uint32_t* __restrict__ cnt, uint32_t* __restrict__ hist,
                                                  uint32_t chunk) {
  __shared__ uint32_t h[4][5][256];
  const uint32_t c = cnt[blockIdx.y];
  const uint32_t lo = blockIdx.x * chunk;
  if (lo >= c) return;
  const uint32_t hi = min(c, lo + chunk);
  for (int i = threadIdx.x; i < 4 * 5 * 256; i += 256) (&h[0][0][0])[i] = 0;
  __syncthreads();
  const uint64_t* s = src + desc[blockIdx.y].off;
  uint32_t(*hw)[256] = h[(threadIdx.x >> 5) & 3];
  for (uint32_t i = lo + threadIdx.x; i < hi; i += 256) {
    const uint64_t k = s[i] >> KEY_LO;
#pragma unroll
    for (int p = 0; p < 5; ++p) atomicAdd(&hw[p][(uint32_t)(k >> (8 * p)) & 255u], 1u);
  }
  __syncthreads();
  uint32_t* o = hist + (uint64_t)blockIdx.y * OS_PASSES_MAX * OS_BINS_MAX + threadIdx.x;
#pragma unroll
  for (int p = 0; p < 5; ++p) {
    const uint32_t v = h[0][p][threadIdx.x] + h[1][p][threadIdx.x] + h[2][p][threadIdx.x] + h[3][p][threadIdx.x];
    if (v) atomicAdd(o + p * OS_BINS_MAX, v);
  }
}

// counts -> exclusive bucket offsets, per block and pass (8-bit digits, five passes).  grid (nb), 256 threads.
__global__ void __launch_bounds__(256) k2_os_offsets(uint32_t* __restrict__ hist) {
  __shared__ uint32_t ws[256 / 32 + 1];
  uint32_t* h = hist + (uint64_t)blockIdx.x * OS_PASSES_MAX * OS_BINS_MAX;
  for (int p = 0; p < 5; ++p) {
    const uint32_t v = h[p * OS_BINS_MAX + threadIdx.x];
    h[p * OS_BINS_MAX + threadIdx.x] = cta_excl_scan_add<256>(v, ws, nullptr);
  }
}

__device__ __forceinline__ uint32_t ld_status(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// KBITS: 0 = the list holds 64-bit elements; 8 / 7 / 6 / 4 = pass 0 of the initial sort builds its elements (the first
// KSYMS symbols of the rotation, KBITS bits each | pos) from the block's text instead of reading a key array.
// WBITS = digit width of the pass.
// SM: shared-memory block with wcnt / dstart / goff / ws (/ lut for pass 0); `stage` = the tile-sized staging buffer of the
// scatter (pass 0 also stages the tile's symbols there); `tin` != nullptr: the tile's elements are already in shared
// memory (k2_os_scatter_pf brings them in with a bulk copy), read from there instead of from `src`.
template <int KBITS, int KSYMS, int WBITS, bool FULL, class SM>
__device__ __forceinline__ void os_tile(SM& sm, uint64_t* stage, const uint64_t* tin, const uint64_t* __restrict__ src,
                                        uint64_t* __restrict__ dst, const BlockDesc* __restrict__ desc,
                                        uint32_t* __restrict__ status, uint32_t tiles_cap, uint32_t epoch, int pass,
                                        uint32_t b, uint32_t c, uint32_t tile, uint32_t base, uint32_t tcount,
                                        const uint32_t (&boff)[(1 << WBITS) / OS_NT]) {
  constexpr int BINS = 1 << WBITS;
  constexpr int DPT = BINS / OS_NT;
  constexpr int OS_WARPS = OS_NT / 32, OS_WCH = OS_TILE / OS_WARPS;
  const int w = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const uint32_t off = desc[b].off;
  const uint64_t* s = src + off + base;
  const int shift = KEY_LO + WBITS * pass;
  constexpr uint32_t DMASK = BINS - 1;

  // ---- rank inside the warp's 256-element chunk; memory order == (warp, row, lane), so the pass is stable
  // (pass 0 of the initial sort need not be stable — nothing is ordered yet — and assigns OS_IPT CONSECUTIVE rotations
  // to a thread: their keys are one sliding window over OS_IPT + KSYMS - 1 symbols of the text)
  uint64_t e[OS_IPT];
  uint32_t rk[OS_IPT];
  if (KBITS != 0) {
    // the tile's symbols (text bytes through the alphabet table), staged once: symbol i of the tile at tx[i]
    uint8_t* tx = reinterpret_cast<uint8_t*>(stage);  // free until the scatter phase
    const uint8_t* t = reinterpret_cast<const uint8_t*>(src) + off;  // src carries the text pointer; c == block length
    const uint32_t need = tcount + KSYMS - 1;
    constexpr int NLD = (OS_TILE + KSYMS - 1 + OS_NT - 1) / OS_NT;
    uint32_t by[NLD];
    if (c >= need) {  // all byte loads of the thread are issued before any of them is used
#pragma unroll
      for (int k = 0; k < NLD; ++k) {
        const uint32_t i = threadIdx.x + k * OS_NT;
        uint32_t p = base + i;
        p = p >= c ? p - c : p;  // cyclic: only the last tile of a block wraps
        by[k] = i < need ? t[p] : 0u;
      }
    } else {  // a block shorter than a key window wraps more than once
#pragma unroll
      for (int k = 0; k < NLD; ++k) {
        const uint32_t i = threadIdx.x + k * OS_NT;
        by[k] = i < need ? t[(base + i) % c] : 0u;
      }
    }
#pragma unroll
    for (int k = 0; k < NLD; ++k) {
      const uint32_t i = threadIdx.x + k * OS_NT;
      if (i < need) tx[i] = KBITS == 8 ? (uint8_t)by[k] : sm.lut[by[k]];
    }
    __syncthreads();
    static_assert(OS_IPT == 8 && KSYMS <= 9, "a thread's window is two 8-byte words of the staged symbols");
    const uint2* tw = reinterpret_cast<const uint2*>(tx) + threadIdx.x;  // symbols 8t .. 8t+15
    const uint2 w0 = tw[0], w1 = tw[1];
    const uint64_t lo = (uint64_t)w0.x | ((uint64_t)w0.y << 32), hi = (uint64_t)w1.x | ((uint64_t)w1.y << 32);
    constexpr uint64_t KMASK = (1ull << (KBITS * KSYMS)) - 1ull;
    uint64_t key = 0;
#pragma unroll
    for (int j = 0; j < KSYMS - 1; ++j) key = (key << KBITS) | ((lo >> (8 * j)) & 0xFFull);
#pragma unroll
    for (int it = 0; it < OS_IPT; ++it) {
      const int q = it + KSYMS - 1;  // symbol entering the window
      const uint64_t sy = q < 8 ? (lo >> (8 * q)) & 0xFFull : (hi >> (8 * (q - 8))) & 0xFFull;
      key = ((key << KBITS) | sy) & KMASK;
      const uint32_t li = threadIdx.x * OS_IPT + it;
      e[it] = (FULL || li < tcount) ? (key << KEY_LO) | (uint64_t)(base + li) : ~0ull;
    }
  } else {
#pragma unroll
    for (int it = 0; it < OS_IPT; ++it) {
      const uint32_t li = w * OS_WCH + it * 32 + lane;
      e[it] = (FULL || li < tcount) ? (tin ? tin[li] : s[li]) : ~0ull;
    }
  }
  if (KBITS != 0) {
    // pass 0 need not be stable (nothing is ordered yet; rotations with equal keys end up in one group whatever their
    // order), so an element's rank inside (warp, digit) is simply what the counter returns.  The digits of 32
    // neighbouring rotations are almost all different, which is the worst case for match.any and the best for atomics.
#pragma unroll
    for (int it = 0; it < OS_IPT; ++it) {
      const uint32_t li = threadIdx.x * OS_IPT + it;
      const uint32_t dgt = (uint32_t)(e[it] >> shift) & DMASK;
      rk[it] = (FULL || li < tcount) ? atomicAdd(&sm.wcnt[w][dgt], 1u) : 0u;
    }
  } else {
    // three separate sweeps so that the eight matches, the eight leader atomics and the eight broadcasts of a thread
    // are independent instructions the scheduler can overlap
    uint32_t peers[OS_IPT];
#pragma unroll
    for (int it = 0; it < OS_IPT; ++it) {
      const uint32_t li = w * OS_WCH + it * 32 + lane;
      const uint32_t dgt = (FULL || li < tcount) ? (uint32_t)(e[it] >> shift) & DMASK : (uint32_t)BINS;  // invalid lanes: own class
      peers[it] = __match_any_sync(0xffffffffu, dgt);
    }
#pragma unroll
    for (int it = 0; it < OS_IPT; ++it) {
      const uint32_t li = w * OS_WCH + it * 32 + lane;
      const uint32_t dgt = (uint32_t)(e[it] >> shift) & DMASK;
      rk[it] = 0;
      if ((FULL || li < tcount) && (peers[it] & lanemask_lt()) == 0)  // lowest lane of the peer group
        rk[it] = atomicAdd(&sm.wcnt[w][dgt], (uint32_t)__popc(peers[it]));
    }
#pragma unroll
    for (int it = 0; it < OS_IPT; ++it) {
      const int leader = __ffs(peers[it]) - 1;
      rk[it] = __shfl_sync(0xffffffffu, rk[it], leader) + __popc(peers[it] & lanemask_lt());
    }
  }
  __syncthreads();

  // ---- per digit (thread t owns digits t*DPT .. t*DPT+DPT-1): exclusive prefix over the warps, tile count, start
  // inside the tile; publish + look back
  uint32_t run[DPT];
  uint32_t tsum = 0;
  if (DPT == 2) {  // both digits of the thread with one 64-bit shared access per warp row
    uint32_t r0 = 0, r1 = 0;
#pragma unroll
    for (int ww = 0; ww < OS_WARPS; ++ww) {
      uint2* cell = reinterpret_cast<uint2*>(&sm.wcnt[ww][threadIdx.x * 2]);
      const uint2 t = *cell;
      *cell = make_uint2(r0, r1);
      r0 += t.x;
      r1 += t.y;
    }
    run[0] = r0;
    run[DPT - 1] = r1;
    tsum = r0 + r1;
  } else {
#pragma unroll
    for (int q = 0; q < DPT; ++q) {
      const uint32_t dgt = threadIdx.x * DPT + q;
      uint32_t r = 0;
#pragma unroll
      for (int ww = 0; ww < OS_WARPS; ++ww) {
        const uint32_t t = sm.wcnt[ww][dgt];
        sm.wcnt[ww][dgt] = r;
        r += t;
      }
      run[q] = r;
      tsum += r;
    }
  }
  uint32_t* st = status + ((uint64_t)b * tiles_cap) * OS_BINS_MAX + threadIdx.x * DPT;
  const uint32_t tag = epoch << 22;
  if (tile != 0) {  // the tile's own counts are known: let the tiles behind it make progress while this one scans
#pragma unroll
    for (int q = 0; q < DPT; ++q) st_status(st + (uint64_t)tile * OS_BINS_MAX + q, tag | ST_AGG | run[q]);
  }
  // exclusive scan of the per-thread digit totals over the CTA (one barrier: every thread sums the warp totals before it)
  uint32_t ds;
  {
    const uint32_t inc = warp_incl_scan_add(tsum);
    if (lane == 31) sm.ws[w] = inc;
    __syncthreads();
    uint32_t wb = 0;
#pragma unroll
    for (int ww = 0; ww < OS_WARPS; ++ww) wb += ww < w ? sm.ws[ww] : 0u;
    ds = wb + inc - tsum;
  }
  {
    // the thread's DPT look-back chains advance in lock step: the status words of one earlier tile are fetched for
    // all of them before any is waited for
    uint32_t excl[DPT];
#pragma unroll
    for (int q = 0; q < DPT; ++q) excl[q] = 0;
    if (tile == 0) {
#pragma unroll
      for (int q = 0; q < DPT; ++q) st_status(st + q, tag | ST_PREFIX | run[q]);
    } else {
      uint32_t open_mask = (1u << DPT) - 1u;
      for (int t = (int)tile - 1; t >= 0 && open_mask; --t) {
        uint32_t v[DPT];
        bool ready;
        do {
          ready = true;
#pragma unroll
          for (int q = 0; q < DPT; ++q) v[q] = ld_status(st + (uint64_t)t * OS_BINS_MAX + q);
#pragma unroll
          for (int q = 0; q < DPT; ++q) ready &= (v[q] >> 22) == epoch;
        } while (!ready);
#pragma unroll
        for (int q = 0; q < DPT; ++q) {
          if (open_mask & (1u << q)) {
            excl[q] += v[q] & ST_VAL;
            if (v[q] & ST_PREFIX) open_mask &= ~(1u << q);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < DPT; ++q) st_status(st + (uint64_t)tile * OS_BINS_MAX + q, tag | ST_PREFIX | (excl[q] + run[q]));
    }
#pragma unroll
    for (int q = 0; q < DPT; ++q) {
      const uint32_t dgt = threadIdx.x * DPT + q;
      sm.dstart[dgt] = ds;
      sm.goff[dgt] = (int)(boff[q] + excl[q]) - (int)ds;
      ds += run[q];
    }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < OS_IPT; ++it) {
    const uint32_t li = KBITS != 0 ? threadIdx.x * OS_IPT + it : w * OS_WCH + it * 32 + lane;
    if ((FULL || li < tcount)) {
      const uint32_t dgt = (uint32_t)(e[it] >> shift) & DMASK;
      stage[sm.dstart[dgt] + sm.wcnt[w][dgt] + rk[it]] = e[it];
    }
  }
  __syncthreads();
  uint64_t* o = dst + off;
#pragma unroll
  for (int k = 0; k < OS_IPT; ++k) {
    const uint32_t i = threadIdx.x + k * OS_NT;
    if (FULL || i < tcount) {
      const uint64_t v = stage[i];
      const uint32_t dgt = (uint32_t)(v >> shift) & DMASK;
      o[sm.goff[dgt] + (int)i] = v;
    }
  }
}

template <int KBITS, int KSYMS, int WBITS>
__global__ void __launch_bounds__(OS_NT, OS_MINB) k2_os_scatter(const uint64_t* __restrict__ src, uint64_t* __restrict__ dst,
                                                          const BlockDesc* __restrict__ desc,
                                                          const uint32_t* __restrict__ cnt,
                                                          const uint32_t* __restrict__ bucket_off,
                                                          uint32_t* __restrict__ status, uint32_t* __restrict__ ticket,
                                                          uint32_t ticket_base, uint32_t tiles_cap, uint32_t epoch,
                                                          int pass, const uint32_t* __restrict__ inuse) {
  extern __shared__ __align__(16) uint8_t os_raw[];
  constexpr int BINS = 1 << WBITS;
  constexpr int DPT = BINS / OS_NT;  // digits per thread in the per-digit phases (1 or 2)
  static_assert(DPT >= 1 && BINS % OS_NT == 0, "digit count must be a multiple of the CTA size");
  using OsSmem = OsSmemT<BINS>;
  OsSmem& sm = *reinterpret_cast<OsSmem*>(os_raw);
  const uint32_t b = blockIdx.x;
  const uint32_t c = cnt[b];
  if (threadIdx.x == 0) sm.tile = atomicAdd(&ticket[b], 1u) - ticket_base;
  const int w = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  uint32_t boff[DPT];  // bucket offsets of the thread's digits: fetched now, used after the look-back
#pragma unroll
  for (int q = 0; q < DPT; ++q)
    boff[q] = bucket_off[((uint64_t)b * OS_PASSES_MAX + pass) * OS_BINS_MAX + threadIdx.x * DPT + q];
#pragma unroll
  for (int i = lane; i < BINS / 4; i += 32) reinterpret_cast<uint4*>(sm.wcnt[w])[i] = make_uint4(0u, 0u, 0u, 0u);
  if (KBITS != 0 && KBITS != 8) build_sym_lut(inuse + b * 8, sm.lut);
  __syncthreads();
  const uint32_t tile = sm.tile;
  const uint32_t base = tile * OS_TILE;
  if (base >= c) return;
  const uint32_t tcount = min((uint32_t)OS_TILE, c - base);
  // full tiles (all but the last of a list) run without per-element bounds checks
  if (tcount == (uint32_t)OS_TILE)
    os_tile<KBITS, KSYMS, WBITS, true>(sm, sm.stage, nullptr, src, dst, desc, status, tiles_cap, epoch, pass, b, c, tile,
                                       base, tcount, boff);
  else
    os_tile<KBITS, KSYMS, WBITS, false>(sm, sm.stage, nullptr, src, dst, desc, status, tiles_cap, epoch, pass, b, c, tile,
                                        base, tcount, boff);
}

// ------------------------------------------------------------------ the same pass, persistent, tiles brought in by TMA
// k2_os_scatter_pf: passes over 64-bit elements (KBITS == 0).  The grid is one wave of CTAs; a CTA takes tiles from a
// global ticket counter (ticket g -> block g % nb, tile g / nb: a tile's ticket is always larger than the tickets of
// the tiles in front of it in its block, which is all the look-back needs) and keeps TWO tiles in shared memory: while
// it ranks and scatters tile i, the bulk-copy engine (cp.async.bulk, 1-D TMA, completion on an mbarrier) brings in
// tile i+1, and thread 0 already holds the ticket of tile i+2 — so neither the ticket's atomic round trip nor the
// tile's load latency is on a CTA's critical path.  The buffer a tile arrived in doubles as the staging buffer of its
// scatter.  A bulk copy needs 16-byte alignment and a tile starts at an arbitrary 8-byte element, so the copy starts at
// the even element at or below it (`a` = 0 or 1 elements of skew).
template <int BINS>
struct OsPfSmem {
  static constexpr int WARPS = OS_NT / 32;
  uint64_t buf[2][OS_TILE + 2];
  uint32_t wcnt[WARPS][BINS];
  uint32_t dstart[BINS];
  int goff[BINS];
  uint32_t ws[OS_NT / 32 + 1];
  uint32_t g[4];            // ticket ring
  uint64_t mbar[2];
  uint8_t lut[4];           // (unused: os_tile's pass-0 branch is not instantiated here)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n"
      " bra WAIT_%=;\n DONE_%=:\n}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int WBITS>
__global__ void __launch_bounds__(OS_NT, OS_MINB) k2_os_scatter_pf(const uint64_t* __restrict__ src, uint64_t* __restrict__ dst,
                                                             const BlockDesc* __restrict__ desc,
                                                             const uint32_t* __restrict__ cnt,
                                                             const uint32_t* __restrict__ bucket_off,
                                                             uint32_t* __restrict__ status, uint32_t* __restrict__ gticket,
                                                             uint32_t nb, uint32_t tiles, uint32_t tiles_cap,
                                                             uint32_t epoch, int pass) {
  extern __shared__ __align__(128) uint8_t os_pf_raw[];
  constexpr int BINS = 1 << WBITS;
  constexpr int DPT = BINS / OS_NT;
  using Sm = OsPfSmem<BINS>;
  Sm& sm = *reinterpret_cast<Sm*>(os_pf_raw);
  const int w = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const uint32_t total = nb * tiles;

  // tile of a ticket: list length c, first element, elements, skew of the bulk copy
  struct Tile { uint32_t b, tile, c, base, tcount, a; const uint64_t* gsrc; uint32_t bytes; };
  auto decode = [&](uint32_t g) {
    Tile t;
    t.b = g % nb;
    t.tile = g / nb;
    t.c = cnt[t.b];
    t.base = t.tile * OS_TILE;
    t.tcount = t.base < t.c ? min((uint32_t)OS_TILE, t.c - t.base) : 0u;
    const uint64_t first = (uint64_t)desc[t.b].off + t.base;
    t.a = (uint32_t)(first & 1u);
    t.gsrc = src + (first - t.a);
    t.bytes = ((t.a + t.tcount + 1u) & ~1u) * 8u;
    return t;
  };
  if (threadIdx.x == 0) {
    mbar_init(&sm.mbar[0], 1);
    mbar_init(&sm.mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t g0 = atomicAdd(gticket, 1u), g1 = atomicAdd(gticket, 1u);
    sm.g[0] = g0;
    sm.g[1] = g1;
  }
#pragma unroll
  for (int i = lane; i < BINS / 4; i += 32) reinterpret_cast<uint4*>(sm.wcnt[w])[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  if (threadIdx.x == 0 && sm.g[0] < total) {  // tile 0 on its way
    const Tile t = decode(sm.g[0]);
    if (t.tcount) {
      mbar_expect_tx(&sm.mbar[0], t.bytes);
      bulk_g2s(sm.buf[0], t.gsrc, t.bytes, &sm.mbar[0]);
    }
  }
  uint32_t phases = 0;  // bit k: parity of the next completion of mbar[k]
  uint32_t boff[DPT], boff_next[DPT];
  {
    const uint32_t g0 = sm.g[0];
#pragma unroll
    for (int q = 0; q < DPT; ++q)
      boff[q] = g0 < total ? bucket_off[((uint64_t)(g0 % nb) * OS_PASSES_MAX + pass) * OS_BINS_MAX + threadIdx.x * DPT + q] : 0u;
  }
  for (uint32_t it = 0;; ++it) {
    const uint32_t cur = it & 1u;
    const uint32_t g = sm.g[it % 3u];
    if (g >= total) break;  // tickets only grow: nothing follows either
    const uint32_t gn = sm.g[(it + 1u) % 3u];
    uint32_t g2 = total;
    if (threadIdx.x == 0) {
      // the ticket of tile it+2: asked for now, stored at the end of the iteration (its slot of the ring was read by
      // everybody before the last barrier of tile it-1), so the atomic's round trip hides behind tile it
      if (gn < total) g2 = atomicAdd(gticket, 1u);
      if (gn < total) {  // tile it+1 into the buffer tile it-1 has left
        const Tile t = decode(gn);
        if (t.tcount) {
          mbar_expect_tx(&sm.mbar[cur ^ 1u], t.bytes);
          bulk_g2s(sm.buf[cur ^ 1u], t.gsrc, t.bytes, &sm.mbar[cur ^ 1u]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < DPT; ++q)
      boff_next[q] = gn < total ? bucket_off[((uint64_t)(gn % nb) * OS_PASSES_MAX + pass) * OS_BINS_MAX + threadIdx.x * DPT + q] : 0u;
    const Tile t = decode(g);
    if (t.tcount) {
      mbar_wait(&sm.mbar[cur], (phases >> cur) & 1u);
      phases ^= 1u << cur;
      if (t.tcount == (uint32_t)OS_TILE)
        os_tile<0, 0, WBITS, true>(sm, sm.buf[cur], sm.buf[cur] + t.a, src, dst, desc, status, tiles_cap, epoch, pass, t.b,
                                   t.c, t.tile, t.base, t.tcount, boff);
      else
        os_tile<0, 0, WBITS, false>(sm, sm.buf[cur], sm.buf[cur] + t.a, src, dst, desc, status, tiles_cap, epoch, pass,
                                    t.b, t.c, t.tile, t.base, t.tcount, boff);
      __syncwarp();
#pragma unroll
      for (int i = lane; i < BINS / 4; i += 32)  // the warp's own counters, for the next tile
        reinterpret_cast<uint4*>(sm.wcnt[w])[i] = make_uint4(0u, 0u, 0u, 0u);
      fence_proxy_async();  // this thread's reads and writes of buf[cur] come before the bulk copy that reuses it
    }
#pragma unroll
    for (int q = 0; q < DPT; ++q) boff[q] = boff_next[q];
    if (threadIdx.x == 0) sm.g[(it + 2u) % 3u] = g2;
    __syncthreads();
  }
}

// ------------------------------------------------------------------ SA entry layout / local-sort geometry
constexpr uint32_t SA_HEAD = 0x80000000u;    // first slot of a group
constexpr uint32_t SA_BIG = 0x40000000u;     // member of a group with more than LOCAL_MAX slots (radix path)
constexpr uint32_t SA_SINGLE = 0x20000000u;  // group of one: final position
constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr int G_NT = 256;        // k2_gather / k2_finish threads
constexpr int LS_NT = 512;       // k2_local_sort threads
constexpr int LS_T = 2048;       // slots per tile (k2_gather, k2_local_sort)
constexpr int LOCAL_MAX = 1535;  // largest group sorted inside a CTA
constexpr int ENUM_MAX_DEFAULT = 48;  // largest group sorted by plain enumeration (larger ones are bucketed first)
constexpr int ENUM_MAX_CAP = 64;
constexpr int LS_CAP = LS_T + LOCAL_MAX;
constexpr int LS_IPT = (LS_CAP + LS_NT - 1) / LS_NT;  // 7 (odd: blocked shared-memory access is conflict-free)
static_assert(LS_CAP < 4095, "group start must fit the 12 bits above the 20-bit key (4095 = sentinel)");
static_assert(LS_CAP + 1 <= LS_NT * LS_IPT, "the prefix sum covers slot `count` too");
static_assert(LOCAL_MAX <= LS_T, "a group owned by tile t must end inside tile t+1");

uint32_t bwt_ls_tile_elems() { return LS_T; }

// ------------------------------------------------------------------ regroup (after a radix sort)
// Tile summaries: last (g,k2)-run head, last g-run head, first (g,k2)-run head inside the tile.
__global__ void __launch_bounds__(RS_NT) k2_rg_flags(const uint64_t* __restrict__ srt, const BlockDesc* __restrict__ desc,
                                                     const uint32_t* __restrict__ cnt, int4* __restrict__ tsum,
                                                     uint32_t tiles_cap, int initial) {
  __shared__ int sh[RS_WARPS], sg[RS_WARPS], sf[RS_WARPS];
  const uint32_t c = cnt[blockIdx.y];
  const uint32_t base = blockIdx.x * RS_TILE;
  if (base >= c) return;
  const uint64_t* s = srt + desc[blockIdx.y].off;
  int lh = -1, lg = -1, fh = 0x7FFFFFFF;
#pragma unroll 4
  for (int k = 0; k < RS_IPT; ++k) {
    uint32_t a = base + threadIdx.x + k * RS_NT;
    if (a < c) {
      uint64_t cur = s[a] >> KEY_LO;
      uint64_t prv = a > 0 ? (s[a - 1] >> KEY_LO) : ~0ull;
      if (cur != prv) { lh = max(lh, (int)a); fh = min(fh, (int)a); }
      bool hg = initial ? (a == 0) : ((cur >> 20) != (prv >> 20));
      if (hg) lg = (int)a;
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    lh = max(lh, __shfl_xor_sync(0xffffffffu, lh, d));
    lg = max(lg, __shfl_xor_sync(0xffffffffu, lg, d));
    fh = min(fh, __shfl_xor_sync(0xffffffffu, fh, d));
  }
  if (lane_id() == 0) { sh[threadIdx.x >> 5] = lh; sg[threadIdx.x >> 5] = lg; sf[threadIdx.x >> 5] = fh; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < RS_WARPS; ++w) { lh = max(lh, sh[w]); lg = max(lg, sg[w]); fh = min(fh, sf[w]); }
    tsum[(uint64_t)blockIdx.y * tiles_cap + blockIdx.x] = make_int4(lh, lg, fh, 0);
  }
}

__device__ __forceinline__ int warp_incl_scan_min(int v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, d);
    if ((int)lane_id() >= d) v = min(v, t);
  }
  return v;
}

// One warp per block: exclusive max-scan of (x, y) over the tiles, exclusive reverse min-scan of z (in place).
__global__ void __launch_bounds__(32) k2_rg_scan(const uint32_t* __restrict__ cnt, int4* __restrict__ tsum,
                                                 uint32_t tiles_cap) {
  const uint32_t c = cnt[blockIdx.x];
  if (c == 0) return;
  const uint32_t nt = (c + RS_TILE - 1) / RS_TILE;
  int4* ts = tsum + (uint64_t)blockIdx.x * tiles_cap;
  int ch = -1, cg = -1;
  for (uint32_t t0 = 0; t0 < nt; t0 += 32) {
    uint32_t t = t0 + lane_id();
    int4 v = t < nt ? ts[t] : make_int4(-1, -1, 0, 0);
    int ih = warp_incl_scan_max(v.x), ig = warp_incl_scan_max(v.y);
    int eh = __shfl_up_sync(0xffffffffu, ih, 1), eg = __shfl_up_sync(0xffffffffu, ig, 1);
    if (lane_id() == 0) { eh = -1; eg = -1; }
    eh = max(eh, ch); eg = max(eg, cg);
    if (t < nt) { ts[t].x = eh; ts[t].y = eg; }
    ch = max(ch, __shfl_sync(0xffffffffu, ih, 31));
    cg = max(cg, __shfl_sync(0xffffffffu, ig, 31));
  }
  int cn = (int)c;  // first head after the tiles handled so far (walking backwards); c = end of the list
  for (uint32_t k0 = 0; k0 < nt; k0 += 32) {
    uint32_t k = k0 + lane_id();
    uint32_t t = nt - 1 - k;
    int v = k < nt ? ts[t].z : 0x7FFFFFFF;
    int im = warp_incl_scan_min(v);
    int em = __shfl_up_sync(0xffffffffu, im, 1);
    if (lane_id() == 0) em = 0x7FFFFFFF;
    em = min(em, cn);
    if (k < nt) ts[t].z = em;
    cn = min(cn, __shfl_sync(0xffffffffu, im, 31));
  }
}

// New ranks, SA slots with flags, per-block statistics.  For the element at index a of the sorted list:
//   fa = start of its g-run, ha = start of its (g,k2)-run, he = end of that run
//   slot = g + (a - fa)        new rank = g + (ha - fa)        run size = he - ha
template <bool initial>
__global__ void __launch_bounds__(RS_NT) k2_rg_apply(const uint64_t* __restrict__ srt, const BlockDesc* __restrict__ desc,
                                                     const uint32_t* __restrict__ cnt, const int4* __restrict__ tsum,
                                                     uint32_t tiles_cap, uint32_t* __restrict__ rank,
                                                     uint32_t* __restrict__ sa, uint32_t* __restrict__ stats) {
  __shared__ int wsh[RS_WARPS], wsg[RS_WARPS], wsf[RS_WARPS];
  __shared__ uint32_t red[3][RS_WARPS];
  const uint32_t c = cnt[blockIdx.y];
  const uint32_t base = blockIdx.x * RS_TILE;
  if (base >= c) return;
  const uint32_t off = desc[blockIdx.y].off;
  const uint64_t* s = srt + off;
  uint32_t* rk = rank + off;
  uint32_t* so = sa + off;
  const int4 carry = tsum[(uint64_t)blockIdx.y * tiles_cap + blockIdx.x];

  // The tile is loaded coalesced into shared memory and read back in a blocked arrangement (thread t owns elements
  // base + t*16 .. +15; index i sits at i + i/16, which keeps both access patterns conflict-free).
  extern __shared__ __align__(16) uint8_t rg_raw[];
  uint64_t* stg = reinterpret_cast<uint64_t*>(rg_raw);
#pragma unroll 4
  for (int k = 0; k < RS_IPT; ++k) {
    const uint32_t i = threadIdx.x + k * RS_NT;
    stg[i + (i >> 4)] = (base + i < c) ? s[base + i] : ~0ull;
  }
  __syncthreads();
  const uint32_t a0 = base + threadIdx.x * RS_IPT;
  uint64_t e[RS_IPT];
  uint64_t prv = ~0ull;
  if (a0 < c) {
    if (threadIdx.x > 0) prv = stg[threadIdx.x * 17 - 2] >> KEY_LO;  // element 16t-1 sits at 17t-2
    else if (a0 > 0) prv = s[a0 - 1] >> KEY_LO;
#pragma unroll
    for (int j = 0; j < RS_IPT; ++j) e[j] = stg[threadIdx.x * 17 + j];
  }
  __syncthreads();
  uint32_t* sa_stage = reinterpret_cast<uint32_t*>(rg_raw);  // initial round: SA slots leave through shared memory
  int lh = -1, lg = -1, fh = 0x7FFFFFFF;
  uint32_t hmask = 0, gmask = 0;
  if (a0 < c) {
    uint64_t p = prv;
#pragma unroll
    for (int j = 0; j < RS_IPT; ++j) {
      if (a0 + j < c) {
        uint64_t cur = e[j] >> KEY_LO;
        if (cur != p) { lh = (int)(a0 + j); hmask |= 1u << j; }
        bool hg = initial ? (a0 + j == 0) : ((cur >> 20) != (p >> 20));
        if (hg) { lg = (int)(a0 + j); gmask |= 1u << j; }
        p = cur;
      }
    }
    if (hmask) fh = (int)a0 + __ffs(hmask) - 1;
  }
  // CTA exclusive max-scan of (lh, lg); exclusive reverse min-scan of fh  (initial round: one g-run, lg scan skipped)
  int ih = warp_incl_scan_max(lh), ig = initial ? 0 : warp_incl_scan_max(lg);
  int rf = fh;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_down_sync(0xffffffffu, rf, d);
    if ((int)lane_id() + d < 32) rf = min(rf, t);
  }
  const int w = threadIdx.x >> 5;
  if (lane_id() == 31) { wsh[w] = ih; wsg[w] = ig; }
  if (lane_id() == 0) wsf[w] = rf;
  __syncthreads();
  int eh = __shfl_up_sync(0xffffffffu, ih, 1), eg = initial ? 0 : __shfl_up_sync(0xffffffffu, ig, 1);
  if (lane_id() == 0) { eh = -1; if (!initial) eg = -1; }
  for (int ww = 0; ww < w; ++ww) { eh = max(eh, wsh[ww]); if (!initial) eg = max(eg, wsg[ww]); }
  eh = max(eh, carry.x);
  if (!initial) eg = max(eg, carry.y);
  int nh = __shfl_down_sync(0xffffffffu, rf, 1);
  if (lane_id() == 31) nh = 0x7FFFFFFF;
  for (int ww = w + 1; ww < RS_WARPS; ++ww) nh = min(nh, wsf[ww]);
  nh = min(nh, carry.z);  // first head after this thread's elements (c when none)

  uint32_t n_new = 0, n_big = 0, n_unres = 0;
  if (a0 < c) {
    int ha = eh, fa = eg;
#pragma unroll
    for (int j = 0; j < RS_IPT; ++j) {
      if (a0 + j < c) {
        const int a = (int)(a0 + j);
        const bool is_h = hmask & (1u << j);
        const bool is_g = gmask & (1u << j);
        if (is_h) ha = a;
        if (is_g) fa = a;
        if (is_h && !is_g) ++n_new;
        const uint32_t later = j < RS_IPT - 1 ? (hmask >> (j + 1)) : 0u;
        const int he = later ? a + __ffs(later) : nh;
        const uint32_t size = (uint32_t)(he - ha);
        const uint32_t pos = (uint32_t)(e[j] & POS_MASK);
        const uint32_t g = initial ? 0u : (uint32_t)(e[j] >> 40) & RANK_MASK;
        const uint32_t nr = g + (uint32_t)(ha - fa);
        const uint32_t slot = g + (uint32_t)(a - fa);
        uint32_t fl = is_h ? SA_HEAD : 0u;
        if (size == 1) fl |= SA_SINGLE;
        else {
          ++n_unres;
          if (size > (uint32_t)LOCAL_MAX) { fl |= SA_BIG; ++n_big; }
        }
        rk[pos] = nr | (size == 1 ? RANK_RESOLVED : 0u);
        if (initial) sa_stage[threadIdx.x * 17 + j] = pos | fl;  // slot == a
        else so[slot] = pos | fl;
      }
    }
  }
  if (initial) {
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < RS_IPT; ++k) {
      const uint32_t i = threadIdx.x + k * RS_NT;
      if (base + i < c) so[base + i] = sa_stage[i + (i >> 4)];
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    n_new += __shfl_xor_sync(0xffffffffu, n_new, d);
    n_big += __shfl_xor_sync(0xffffffffu, n_big, d);
    n_unres += __shfl_xor_sync(0xffffffffu, n_unres, d);
  }
  if (lane_id() == 0) { red[0][w] = n_new; red[1][w] = n_big; red[2][w] = n_unres; }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t a = 0, b = 0, u = 0;
    for (int ww = 0; ww < RS_WARPS; ++ww) { a += red[0][ww]; b += red[1][ww]; u += red[2][ww]; }
    uint32_t* st = stats + blockIdx.y * 4;
    if (a) atomicAdd(&st[0], a);
    if (b) atomicAdd(&st[1], b);
    if (u) atomicAdd(&st[2], u);
  }
}

// Per-block state machine after a round. state: 0 active, 1 fix-up pending, 2 done.
// stats[b] = {groups created by splitting, members of BIG groups, unresolved rotations, periodic flag}
__global__ void k2_round_finalize(uint32_t nb, uint32_t round_no, const BlockDesc* __restrict__ desc,
                                  uint32_t* __restrict__ cnt, uint32_t* __restrict__ stats,
                                  uint32_t* __restrict__ state, uint32_t* __restrict__ sparse,
                                  uint32_t* __restrict__ rounds, uint32_t* __restrict__ global) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  uint32_t* st = stats + b * 4;
  const uint32_t created = st[0], big = st[1], unres = st[2];
  uint32_t s = state[b];
  if (s != 2) {
    if (unres == 0) {
      s = 2;
      rounds[b] = round_no;
    } else if (s == 1) {
      atomicOr(&global[3], 1u);  // the fix-up key makes every rotation distinct; anything else is a bug
      s = 2;
    } else if (created == 0) {
      s = 1;  // nothing split: the block is periodic, the groups are the sets of equal rotations
      st[3] = 1;
    }
    state[b] = s;
    if (s != 2) {
      // few unresolved rotations left: the whole block takes the radix path next round (the per-tile cost of the
      // local sort would dominate)
      const bool sp = unres < (desc[b].n >> 5);
      sparse[b] = sp ? 1u : 0u;
      atomicAdd(&global[0], unres);
      atomicMax(&global[1], sp ? unres : big);
      atomicAdd(&global[2], sp ? unres : big);  // rotations that take the radix path next round
    }
  }
  st[0] = 0; st[1] = 0; st[2] = 0;
  cnt[b] = 0;
}

// ------------------------------------------------------------------ periodic blocks: shift = min pos of group 0
// One CTA per block; does work only for a block that has just entered the fix-up state.
__global__ void __launch_bounds__(256) k2_periodic_shift(const BlockDesc* __restrict__ desc,
                                                         const uint32_t* __restrict__ state,
                                                         const uint32_t* __restrict__ sa, uint32_t* __restrict__ shift) {
  __shared__ uint32_t s_min, s_end;
  const uint32_t b = blockIdx.x;
  if (state[b] != 1 || shift[b] != NONE) return;
  const BlockDesc d = desc[b];
  const uint32_t* s = sa + d.off;
  if (threadIdx.x == 0) { s_min = NONE; s_end = 0; }
  __syncthreads();
  uint32_t mine = NONE;
  for (uint32_t base = 0; base < d.n; base += 256) {
    const uint32_t i = base + threadIdx.x;
    uint32_t e = i < d.n ? s[i] : SA_HEAD;
    const bool stop = (i > 0) && (e & SA_HEAD);  // head of the second group (or past the end)
    if (stop) atomicMin(&s_min, i), atomicOr(&s_end, 1u);
    __syncthreads();
    const uint32_t lim = s_min;  // first slot not in group 0 (NONE while not seen)
    if (i < lim && i < d.n) mine = min(mine, e & RANK_MASK);
    const bool done = s_end != 0;
    __syncthreads();
    if (done) break;
  }
#pragma unroll
  for (int dlt = 16; dlt > 0; dlt >>= 1) mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, dlt));
  if (threadIdx.x == 0) s_min = NONE;
  __syncthreads();
  if (lane_id() == 0) atomicMin(&s_min, mine);
  __syncthreads();
  if (threadIdx.x == 0) shift[b] = s_min;
}

// ------------------------------------------------------------------ k2_gather: work lists of the next round
// Slot order, one CTA per 2048-slot tile.  Every unresolved slot fetches its key rank[(pos+h) mod n] and the head
// slot g = rank[pos] of its group.
//   small groups : a 64-bit entry  [19:0] pos | [39:20] key | [50:40] slot - g | [61:51] slot - tile base  is appended
//                  to the tile's dense list (slot order kept: warp w owns slots w*256.., rows of 32) in `lst`;
//                  tile_meta[b][t] = (entries, entries before the tile's first HEAD slot) lets the CTA that owns a
//                  group spanning two tiles pick up its tail.
//   BIG groups / sparse blocks : [59:40] g | [39:20] key | [19:0] pos appended to A (any order), counted in cnt[b].
__global__ void __launch_bounds__(G_NT, 5) k2_gather(const BlockDesc* __restrict__ desc, const uint32_t* __restrict__ rank,
                                                  const uint32_t* __restrict__ sa, const uint32_t* __restrict__ state,
                                                  const uint32_t* __restrict__ shiftv,
                                                  const uint32_t* __restrict__ sparse, uint32_t h,
                                                  uint64_t* __restrict__ lst, uint64_t* __restrict__ A,
                                                  uint32_t* __restrict__ cnt, uint2* __restrict__ tile_meta,
                                                  uint32_t ls_tiles_cap) {
  __shared__ uint32_t ws[G_NT / 32 + 1];
  __shared__ uint32_t s_wtot[G_NT / 32];
  __shared__ int s_whead[G_NT / 32];
  __shared__ uint32_t s_base, s_lead;
  constexpr int ROWS = LS_T / G_NT;  // 8 rows of 32 slots per warp
  const BlockDesc d = desc[blockIdx.y];
  const uint32_t n = d.n;
  const uint32_t base = blockIdx.x * LS_T;
  if (base >= n) return;
  const uint32_t st = state[blockIdx.y];
  if (st == 2) return;
  const uint64_t ti = (uint64_t)blockIdx.y * ls_tiles_cap + blockIdx.x;
  if (threadIdx.x == 0) s_lead = 0;
  const uint32_t* rk = rank + d.off;
  const uint32_t* s = sa + d.off;
  const uint32_t hm = h % n;
  const uint32_t sh = st == 1 ? shiftv[blockIdx.y] : 0u;
  const bool sp = sparse[blockIdx.y] != 0;  // sparse block: every unresolved slot is emitted to the radix path
  const int w = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  uint32_t ev[ROWS];
#pragma unroll
  for (int k = 0; k < ROWS; ++k) {
    const uint32_t i = base + w * (ROWS * 32) + k * 32 + lane;
    ev[k] = i < n ? s[i] : SA_SINGLE;
  }
  // The head slot of a small group is read off the HEAD flags of the tile (ballots inside a row, carried down the
  // warp's rows and across the warps); only entries whose group starts in the previous tile (the tile's lead entries)
  // and the elements of the radix path fetch rank[pos] for it.
  uint32_t nbig = 0, wrun = 0;
  uint32_t kk[ROWS];  // [19:0] key | [21:20] 0 nothing, 1 small-group entry, 2 radix-path element | [29:22] index of the
                      // small-group entry inside the warp's part of the list
  int hg[ROWS];       // small entry: head slot relative to the tile base (-1: not in an earlier row of this warp);
                      // radix path: head slot from rank[pos]
  int carry = -1;
#pragma unroll
  for (int k = 0; k < ROWS; ++k) {
    const uint32_t rel = w * (ROWS * 32) + k * 32 + lane;
    const uint32_t i = base + rel;
    const uint32_t e = ev[k];
    const uint32_t hb = __ballot_sync(0xffffffffu, i < n && (e & SA_HEAD));
    const uint32_t mine = hb & (lanemask_lt() | (1u << lane));
    hg[k] = mine ? (int)(rel - lane) + 31 - __clz(mine) : carry;
    if (hb) carry = (int)(rel - lane) + 31 - __clz(hb);
    kk[k] = 0;
    bool small = false;
    if (!(e & SA_SINGLE)) {
      const uint32_t pos = e & RANK_MASK;
      uint32_t k2;
      if (st == 0) {
        uint32_t p2 = pos + hm;
        if (p2 >= n) p2 -= n;
        k2 = rk[p2] & RANK_MASK;
      } else {
        const uint32_t rl = pos >= sh ? pos - sh : pos + n - sh;  // (pos - shift) mod n
        k2 = n - 1 - rl;
      }
      if (sp || (e & SA_BIG)) {
        hg[k] = (int)(rk[pos] & RANK_MASK);
        kk[k] = k2 | (2u << 20);
        ++nbig;
      } else {
        kk[k] = k2 | (1u << 20);
        small = true;
      }
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, small);
    kk[k] |= (wrun + __popc(bal & lanemask_lt())) << 22;
    wrun += __popc(bal);
  }
  __syncthreads();
  if (lane == 0) {
    s_wtot[w] = wrun;
    s_whead[w] = carry;
  }
  const int anybig = __syncthreads_or((int)nbig);
  uint32_t wbase = 0, acnt = 0;
  int cin = -1;  // last head slot in the warps before this one
#pragma unroll
  for (int ww = 0; ww < G_NT / 32; ++ww) {
    if (ww < w) {
      wbase += s_wtot[ww];
      cin = max(cin, s_whead[ww]);
    }
    acnt += s_wtot[ww];
  }
  // entries before the tile's first HEAD slot belong to a group owned by the previous tile
  uint32_t lead = 0;
  uint64_t* lo = lst + d.off + base;
#pragma unroll
  for (int k = 0; k < ROWS; ++k) {
    if (((kk[k] >> 20) & 3u) == 1u) {
      const uint32_t rel = w * (ROWS * 32) + k * 32 + lane;
      const uint32_t pos = ev[k] & RANK_MASK;
      const int h = hg[k] < 0 ? cin : hg[k];
      uint32_t dist;
      if (h < 0) {  // the group's head lies in the previous tile
        dist = base + rel - (rk[pos] & RANK_MASK);
        ++lead;
      } else {
        dist = rel - (uint32_t)h;
      }
      lo[wbase + ((kk[k] >> 22) & 0xFFu)] =
          (uint64_t)pos | ((uint64_t)(kk[k] & RANK_MASK) << 20) | ((uint64_t)dist << 40) | ((uint64_t)rel << 51);
    }
  }
#pragma unroll
  for (int dlt = 16; dlt > 0; dlt >>= 1) lead += __shfl_xor_sync(0xffffffffu, lead, dlt);
  if (lane == 0 && lead) atomicAdd(&s_lead, lead);
  __syncthreads();
  if (threadIdx.x == 0) tile_meta[ti] = make_uint2(acnt, s_lead);
  if (!anybig) return;
  uint32_t total;
  const uint32_t ex = cta_excl_scan_add<G_NT>(nbig, ws, &total);
  if (threadIdx.x == 0) s_base = atomicAdd(&cnt[blockIdx.y], total);
  __syncthreads();
  uint64_t* o = A + d.off + s_base + ex;
#pragma unroll
  for (int k = 0; k < ROWS; ++k)
    if (((kk[k] >> 20) & 3u) == 2u)
      *o++ = ((uint64_t)(uint32_t)hg[k] << 40) | ((uint64_t)(kk[k] & RANK_MASK) << 20) | (ev[k] & RANK_MASK);
}

// ------------------------------------------------------------------ k2_local_sort: groups up to LOCAL_MAX slots
// CTA (b, t) owns the groups whose HEAD lies in tile t: its window is the tile's list minus the leading entries
// that continue a group of tile t-1, plus the leading entries of tile t+1's list — at most LS_T + LOCAL_MAX
// entries, dense (resolved slots are not in the lists).  Elements live in registers; shared memory holds, per
// window index, the composite (group start << 20 | key) of the element currently placed there.
//   * split levels: a group larger than ENUM_MAX whose keys are not all equal is partitioned into key-range
//     buckets (range [min,max] of its keys, ~1-2 entries per bucket, counting pass with shared-memory atomics);
//     the buckets are groups of their own from then on; repeated until every group is small or flat (all keys
//     equal — nothing to sort, which is what heavy duplicates end as);
//   * enumeration: final index of an element = group start + #{smaller keys} + #{equal keys placed earlier};
//     equal keys stay one (smaller) group.
// A group occupies consecutive slots, so window index i of a group maps to slot (head slot + i - group start);
// SA and rank are written in place.
constexpr uint32_t GS_FLAT = 0x8000u;
constexpr int LS_PAD = ENUM_MAX_CAP + 8;
struct LsSmem {
  uint32_t ck[LS_CAP + LS_PAD];
  uint32_t cnt[LS_CAP + 8];
  uint32_t gmin[LS_CAP + 8];
  uint32_t gmax[LS_CAP + 8];
  uint32_t gsz[LS_CAP + 8];
  uint32_t wsc[LS_NT / 32 + 1];
  uint32_t red[3][LS_NT / 32];
};
size_t bwt_ls_smem_bytes() { return sizeof(LsSmem); }

template <int ENUM_MAX>
__global__ void __launch_bounds__(LS_NT, 2) k2_local_sort(const BlockDesc* __restrict__ desc,
                                                          const uint32_t* __restrict__ state,
                                                          const uint32_t* __restrict__ sparse,
                                                          uint32_t* __restrict__ sa, const uint64_t* __restrict__ lst,
                                                          uint32_t* __restrict__ rank,
                                                          const uint2* __restrict__ tile_meta,
                                                          uint32_t ls_tiles_cap, uint32_t* __restrict__ stats) {
  extern __shared__ __align__(16) uint8_t ls_raw[];
  LsSmem& sm = *reinterpret_cast<LsSmem*>(ls_raw);
  const uint32_t b = blockIdx.y, t = blockIdx.x;
  const BlockDesc d = desc[b];
  const uint32_t n = d.n;
  const uint32_t tbase = t * LS_T;
  if (tbase >= n) return;
  if (state[b] == 2 || sparse[b]) return;
  const uint64_t ti = (uint64_t)b * ls_tiles_cap + t;
  const uint2 m0 = tile_meta[ti];
  const uint2 m1 = (tbase + LS_T < n) ? tile_meta[ti + 1] : make_uint2(0u, 0u);
  const uint32_t own = m0.x - m0.y, over = m1.y;
  const uint32_t count = own + over;  // <= LS_T + LOCAL_MAX
  if (count == 0) return;
  const uint64_t* l0 = lst + d.off + tbase + m0.y;
  const uint64_t* l1 = lst + d.off + tbase + LS_T;
  uint32_t* s = sa + d.off;
  const int w = threadIdx.x >> 5;

  // ---- load. Per element: pv = pos, kv = key, hv = head slot of its group, gp = group start | index << 12 | pending << 31
  uint32_t pv[LS_IPT], kv[LS_IPT], hv[LS_IPT], gp[LS_IPT];
  uint32_t heads_before = 0;
#pragma unroll
  for (int k = 0; k < LS_IPT; ++k) {
    const uint32_t r = threadIdx.x + k * LS_NT;
    gp[k] = NONE;
    if (r < count) {
      const uint64_t en = r < own ? l0[r] : l1[r - own];
      const uint32_t dist = (uint32_t)(en >> 40) & 0x7FFu;
      const uint32_t slot = (r < own ? tbase : tbase + LS_T) + ((uint32_t)(en >> 51) & 0x7FFu);
      pv[k] = (uint32_t)en & RANK_MASK;
      kv[k] = (uint32_t)(en >> 20) & RANK_MASK;
      hv[k] = slot - dist;
      const uint32_t gs = r - dist;
      gp[k] = gs | (r << 12);
      sm.ck[r] = (gs << 20) | kv[k];
      if (dist == 0) {
        ++heads_before;
        sm.gsz[r] = 0;
        sm.gmin[r] = NONE;
        sm.gmax[r] = 0;
      }
    }
  }
  for (uint32_t r = count + threadIdx.x; r < count + LS_PAD; r += LS_NT) sm.ck[r] = NONE;
  __syncthreads();
  bool any_pend = false;
#pragma unroll
  for (int k = 0; k < LS_IPT; ++k) {
    if (gp[k] != NONE) {
      const uint32_t gs = gp[k] & 0xFFFu;
      const bool pend = (sm.ck[gs + ENUM_MAX] >> 20) == gs;  // more than ENUM_MAX entries
      if (pend) gp[k] |= 0x80000000u;
      any_pend |= pend;
    }
  }

  // ---- split levels
  for (int level = 0; level < 24; ++level) {
    if (!__syncthreads_or((int)any_pend)) break;
    for (uint32_t r = threadIdx.x; r <= count; r += LS_NT) sm.cnt[r] = 0;
#pragma unroll
    for (int k = 0; k < LS_IPT; ++k)
      if (gp[k] != NONE && (gp[k] >> 31)) {
        const uint32_t gs = gp[k] & 0xFFFu;
        atomicMin(&sm.gmin[gs], kv[k]);
        atomicMax(&sm.gmax[gs], kv[k]);
        if (level == 0) atomicAdd(&sm.gsz[gs], 1u);
      }
    __syncthreads();
    uint32_t bi[LS_IPT];  // bucket index | arrival index << 12
    uint32_t flat_heads = 0;
#pragma unroll
    for (int k = 0; k < LS_IPT; ++k) {
      bi[k] = NONE;
      if (gp[k] != NONE && (gp[k] >> 31)) {
        const uint32_t gs = gp[k] & 0xFFFu;
        const uint32_t mn = sm.gmin[gs], mx = sm.gmax[gs];
        const uint32_t size = sm.gsz[gs] & 0xFFFu;
        if (mn == mx) {  // all keys equal: nothing to sort
          gp[k] &= 0x7FFFFFFFu;
          if (((gp[k] >> 12) & 0xFFFu) == gs) flat_heads |= 1u << k;  // flag written after the barrier (others read gsz)
        } else {
          const int lnb = 31 - __clz(size);            // buckets = largest power of two <= size
          const int rb = 32 - __clz(mx - mn);           // bits of the key range
          const int sh = max(rb - lnb, 0);
          const uint32_t bs = gs + ((kv[k] - mn) >> sh);
          bi[k] = bs | (atomicAdd(&sm.cnt[bs], 1u) << 12);
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < LS_IPT; ++k)
      if (flat_heads & (1u << k)) sm.gsz[gp[k] & 0xFFFu] |= GS_FLAT;
    {  // exclusive prefix sum of the bucket counters over indices 0..count (blocked), in place
      const uint32_t r0 = threadIdx.x * LS_IPT;
      uint32_t loc[LS_IPT];
      uint32_t sum = 0;
#pragma unroll
      for (int j = 0; j < LS_IPT; ++j) {
        loc[j] = (r0 + j <= count) ? sm.cnt[r0 + j] : 0u;
        sum += loc[j];
      }
      uint32_t ex = cta_excl_scan_add<LS_NT>(sum, sm.wsc, nullptr);
#pragma unroll
      for (int j = 0; j < LS_IPT; ++j) {
        if (r0 + j <= count) sm.cnt[r0 + j] = ex;
        ex += loc[j];
      }
    }
    __syncthreads();
    any_pend = false;
#pragma unroll
    for (int k = 0; k < LS_IPT; ++k) {
      if (bi[k] != NONE) {
        const uint32_t gs = gp[k] & 0xFFFu;
        const uint32_t bs = bi[k] & 0xFFFu, idx = bi[k] >> 12;
        const uint32_t pb = sm.cnt[bs];
        const uint32_t bstart = gs + (pb - sm.cnt[gs]);
        const uint32_t bcount = sm.cnt[bs + 1] - pb;
        const uint32_t np = bstart + idx;
        sm.ck[np] = (bstart << 20) | kv[k];
        if (idx == 0) {
          sm.gsz[bstart] = bcount;
          sm.gmin[bstart] = NONE;
          sm.gmax[bstart] = 0;
        }
        hv[k] += bstart - gs;  // slot of the (new) group's first entry
        const bool pend = bcount > (uint32_t)ENUM_MAX;
        gp[k] = bstart | (np << 12) | (pend ? 0x80000000u : 0u);
        any_pend |= pend;
      }
    }
  }
  __syncthreads();

  // ---- enumeration inside every group / bucket: lt = #{keys below mine}, eq = #{keys equal}, eqb = #{equal keys
  // at an earlier index};  new index = start + lt + eqb, the head of the (sub)group sits at start + lt.  The scan
  // runs over 128-bit vectors from the aligned start of the group until a vector ends outside the group: entries
  // of earlier groups in the first vector compare below (subtracted again), entries of later groups compare above.
  uint32_t n_heads = 0, n_unres = 0;
  uint32_t* rk = rank + d.off;
#pragma unroll
  for (int k = 0; k < LS_IPT; ++k) {
    if (gp[k] != NONE) {
      const uint32_t gs = gp[k] & 0xFFFu, r = (gp[k] >> 12) & 0xFFFu;
      uint32_t lt = 0, eq, eqb;
      const uint32_t gz = sm.gsz[gs];
      if (gz & GS_FLAT) {
        eq = gz & 0xFFFu;
        eqb = r - gs;
      } else {
        const uint32_t mine = (gs << 20) | kv[k];
        const uint32_t m1 = mine + 1;
        uint32_t jb = gs & ~3u;
        const uint32_t rb = r & ~3u;
        uint32_t le = 0;
        for (; jb < rb; jb += 4) {
          const uint4 c = *reinterpret_cast<const uint4*>(&sm.ck[jb]);
          lt += (c.x < mine) + (c.y < mine) + (c.z < mine) + (c.w < mine);
          le += (c.x < m1) + (c.y < m1) + (c.z < m1) + (c.w < m1);
        }
        eqb = le - lt;
        uint32_t lastw;
        {
          const uint4 c = *reinterpret_cast<const uint4*>(&sm.ck[jb]);
          const uint32_t dd = r & 3u;
          lt += (c.x < mine) + (c.y < mine) + (c.z < mine) + (c.w < mine);
          le += (c.x < m1) + (c.y < m1) + (c.z < m1) + (c.w < m1);
          eqb += (dd > 0 && c.x == mine) + (dd > 1 && c.y == mine) + (dd > 2 && c.z == mine);
          lastw = c.w;
          jb += 4;
        }
        while ((lastw >> 20) == gs) {
          const uint4 c = *reinterpret_cast<const uint4*>(&sm.ck[jb]);
          lt += (c.x < mine) + (c.y < mine) + (c.z < mine) + (c.w < mine);
          le += (c.x < m1) + (c.y < m1) + (c.z < m1) + (c.w < m1);
          lastw = c.w;
          jb += 4;
        }
        eq = le - lt;
        lt -= gs & 3u;  // the entries of earlier groups in the first vector
      }
      uint32_t fl = 0;
      if (eqb == 0) { fl = SA_HEAD; ++n_heads; }
      if (eq == 1) fl |= SA_SINGLE; else ++n_unres;
      s[hv[k] + lt + eqb] = pv[k] | fl;
      rk[pv[k]] = (hv[k] + lt) | (eq == 1 ? RANK_RESOLVED : 0u);
    }
  }
#pragma unroll
  for (int dlt = 16; dlt > 0; dlt >>= 1) {
    n_heads += __shfl_xor_sync(0xffffffffu, n_heads, dlt);
    n_unres += __shfl_xor_sync(0xffffffffu, n_unres, dlt);
    heads_before += __shfl_xor_sync(0xffffffffu, heads_before, dlt);
  }
  if (lane_id() == 0) { sm.red[0][w] = n_heads; sm.red[1][w] = n_unres; sm.red[2][w] = heads_before; }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t a = 0, u = 0, hb = 0;
    for (int ww = 0; ww < LS_NT / 32; ++ww) { a += sm.red[0][ww]; u += sm.red[1][ww]; hb += sm.red[2][ww]; }
    uint32_t* stp = stats + b * 4;
    if (a > hb) atomicAdd(&stp[0], a - hb);
    if (u) atomicAdd(&stp[2], u);
  }
}

// ------------------------------------------------------------------ k2_local_sort_rx: the group sort as an LSD radix sort
// Same contract as k2_local_sort (groups up to LOCAL_MAX slots, owned by the tile that holds their HEAD), different
// method: the CTA's window is sorted as a whole, in shared memory, by the 32-bit composite (group start | key) with
// four stable 8-bit counting passes (match.any ranking + per-warp counters, like k2_os_scatter but without leaving the
// SM).  The work per entry no longer depends on the group sizes (the enumeration sort scans the whole group for every
// entry, and the lanes of a warp wait for the largest group among them), and equal (group, key) runs — the new groups —
// are read off the sorted window with ballots.
//   * a CTA owns `tpc` consecutive tiles and packs as many of them as fit (<= LS_CAP entries) into one window, so the
//     sparse lists of the later rounds are sorted a few thousand entries at a time instead of a few dozen;
//   * element = (composite [31:20] group start (window index) | [19:0] key, pos); window index i of a group maps to
//     slot ghead[group start] + i - group start, so SA and rank are written in place as before;
//   * the window is padded with all-ones elements to a whole number of rows per warp (they sort to the end), and the
//     passes are compiled once per row count, so the inner loops carry no bounds checks.
constexpr int RX_NT = 512, RX_WARPS = RX_NT / 32;
constexpr int RX_ROWS = (LS_CAP + 1 + RX_NT - 1) / RX_NT;  // rows of 32 entries per warp: at most 7
constexpr int RX_TPC_MAX = 32;                              // tiles per CTA
static_assert(RX_ROWS * RX_NT == LS_CAP + 1, "the padded window fills the stage exactly");
struct RxSmem {
  uint2 stage[LS_CAP + 1];  // .x = pos, .y = composite
  uint32_t wcnt[RX_WARPS][256];
  uint32_t dstart[256];
  uint32_t ghead[LS_CAP + 1];
  uint32_t ws[RX_WARPS];
  int whead[RX_WARPS];
  uint32_t tcnt[RX_TPC_MAX + 2], tlead[RX_TPC_MAX + 2];
  uint32_t red[3][RX_WARPS];
};
size_t bwt_rx_smem_bytes() { return sizeof(RxSmem); }

// shared-memory atomicAdd executed by the lanes with p set (a predicated ATOMS, no divergent region)
__device__ __forceinline__ uint32_t atoms_add_if(bool p, uint32_t* addr, uint32_t v) {
  uint32_t r = 0;
  asm volatile(
      "{\n .reg .pred q;\n setp.ne.u32 q, %3, 0;\n @q atom.shared.add.u32 %0, [%1], %2;\n}"
      : "+r"(r)
      : "r"((uint32_t)__cvta_generic_to_shared(addr)), "r"(v), "r"((uint32_t)p)
      : "memory");
  return r;
}
__device__ __forceinline__ uint32_t lanemask_gt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_gt;" : "=r"(m));
  return m;
}

// Sorts the padded window (RPW rows per warp) and writes SA / rank.  Returns (new heads, unresolved) of this thread.
template <int RPW>
__device__ __forceinline__ void rx_sort_window(RxSmem& sm, uint32_t count, uint32_t* __restrict__ s,
                                               uint32_t* __restrict__ rk, uint32_t& n_heads, uint32_t& n_unres) {
  const int w = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const uint32_t wbase = w * RPW * 32;
  const uint32_t lt = lanemask_lt(), gt = lanemask_gt();
  uint32_t* wc = sm.wcnt[w];
#pragma unroll 1
  for (uint32_t pass = 0; pass < 4; ++pass) {
    const uint32_t sel = 0x4440u | pass;  // byte `pass` of the composite
    uint2 ev[RPW];
    uint32_t rkv[RPW], peers[RPW];
#pragma unroll
    for (int it = 0; it < RPW; ++it) ev[it] = sm.stage[wbase + it * 32 + lane];
#pragma unroll
    for (int it = 0; it < RPW; ++it) peers[it] = __match_any_sync(0xffffffffu, __byte_perm(ev[it].y, 0u, sel));
#pragma unroll
    for (int it = 0; it < RPW; ++it)  // the highest lane of a peer group counts for the group
      rkv[it] = atoms_add_if((peers[it] & gt) == 0u, wc + __byte_perm(ev[it].y, 0u, sel), (uint32_t)__popc(peers[it]));
#pragma unroll
    for (int it = 0; it < RPW; ++it)
      rkv[it] = __shfl_sync(0xffffffffu, rkv[it], 31 - __clz(peers[it])) + __popc(peers[it] & lt);
    __syncthreads();  // every entry is in registers, every counter final
    uint32_t run = 0, inc = 0;
    if (threadIdx.x < 256) {
#pragma unroll
      for (int ww = 0; ww < RX_WARPS; ++ww) {
        const uint32_t t = sm.wcnt[ww][threadIdx.x];
        sm.wcnt[ww][threadIdx.x] = run;
        run += t;
      }
      inc = warp_incl_scan_add(run);
      if (lane == 31) sm.ws[w] = inc;
    }
    __syncthreads();
    if (threadIdx.x < 256) {
      uint32_t base = 0;
#pragma unroll
      for (int ww = 0; ww < 8; ++ww) base += ww < w ? sm.ws[ww] : 0u;
      sm.dstart[threadIdx.x] = base + inc - run;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < RPW; ++it) {
      const uint32_t dgt = __byte_perm(ev[it].y, 0u, sel);
      sm.stage[sm.dstart[dgt] + wc[dgt] + rkv[it]] = ev[it];
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) wc[i * 32 + lane] = 0;  // the warp's own row, for the next pass / window
    __syncthreads();
  }

  // ---- the sorted window: runs of equal composites are the new groups.  hd = index of the run's first entry: last
  // run start at or before li (ballots inside a row, carried down the warp's rows, then across the warps)
  uint2 cur[RPW];
  uint32_t flg[RPW];  // bit 0: first of its run, bit 1: last of its run
  int hd[RPW];
  int carry = -1;
#pragma unroll
  for (int it = 0; it < RPW; ++it) {
    const uint32_t li = wbase + it * 32 + lane;
    cur[it] = sm.stage[li];
    const uint32_t pv = li > 0 ? sm.stage[li - 1].y : ~cur[it].y;
    const uint32_t nx = li + 1 < count ? sm.stage[li + 1].y : ~cur[it].y;
    const bool valid = li < count;
    const bool first = valid && cur[it].y != pv;
    flg[it] = (first ? 1u : 0u) | ((valid && cur[it].y != nx) ? 2u : 0u);
    const uint32_t bal = __ballot_sync(0xffffffffu, first);
    const uint32_t mine = bal & ~gt;
    hd[it] = mine ? (int)(wbase + it * 32) + 31 - __clz(mine) : carry;
    if (bal) carry = (int)(wbase + it * 32) + 31 - __clz(bal);
  }
  if (lane == 0) sm.whead[w] = carry;
  __syncthreads();
  int cin = -1;
#pragma unroll
  for (int ww = 0; ww < RX_WARPS; ++ww) cin = max(cin, ww < w ? sm.whead[ww] : -1);
#pragma unroll
  for (int it = 0; it < RPW; ++it) {
    const uint32_t li = wbase + it * 32 + lane;
    if (li < count) {
      const uint32_t h0 = (uint32_t)(hd[it] < 0 ? cin : hd[it]);
      const uint32_t gs = cur[it].y >> 20, pos = cur[it].x;
      const uint32_t hs = sm.ghead[gs];
      const bool single = flg[it] == 3u;
      uint32_t fl = (flg[it] & 1u) ? SA_HEAD : 0u;
      if (single) fl |= SA_SINGLE; else ++n_unres;
      n_heads += flg[it] & 1u;
      s[hs + (li - gs)] = pos | fl;
      rk[pos] = (hs + (h0 - gs)) | (single ? RANK_RESOLVED : 0u);
    }
  }
}

__global__ void __launch_bounds__(RX_NT, 2) k2_local_sort_rx(const BlockDesc* __restrict__ desc,
                                                              const uint32_t* __restrict__ state,
                                                              const uint32_t* __restrict__ sparse,
                                                              uint32_t* __restrict__ sa, const uint64_t* __restrict__ lst,
                                                              uint32_t* __restrict__ rank,
                                                              const uint2* __restrict__ tile_meta,
                                                              uint32_t ls_tiles_cap, uint32_t* __restrict__ stats,
                                                              uint32_t tpc) {
  extern __shared__ __align__(16) uint8_t rx_raw[];
  RxSmem& sm = *reinterpret_cast<RxSmem*>(rx_raw);
  const uint32_t b = blockIdx.y;
  const BlockDesc d = desc[b];
  const uint32_t n = d.n;
  const uint32_t ntiles = (n + LS_T - 1) / LS_T;
  const uint32_t T0 = blockIdx.x * tpc;
  if (T0 >= ntiles) return;
  if (state[b] == 2 || sparse[b]) return;
  const uint32_t nt = min(tpc, ntiles - T0);  // tiles of this CTA: T0 .. T0+nt-1 (tile T0+nt only lends its lead entries)
  if (threadIdx.x <= nt) {
    const uint32_t t = T0 + threadIdx.x;
    const uint2 m = t < ntiles ? tile_meta[(uint64_t)b * ls_tiles_cap + t] : make_uint2(0u, 0u);
    sm.tcnt[threadIdx.x] = m.x;
    sm.tlead[threadIdx.x] = m.y;
  }
  const int w = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  for (int i = lane; i < 256; i += 32) sm.wcnt[w][i] = 0;
  __syncthreads();
  uint32_t* s = sa + d.off;
  uint32_t* rk = rank + d.off;
  const uint64_t* lb = lst + d.off;
  uint32_t n_heads = 0, n_unres = 0, heads_before = 0;

  uint32_t a = 0;
  while (a < nt) {
    // ---- the next window: tiles a .. e-1 (+ the lead entries of tile e), as many as fit
    uint32_t e = a + 1;
    uint32_t count = sm.tcnt[a] - sm.tlead[a];
    while (e < nt && count + sm.tcnt[e] + sm.tlead[e + 1] <= (uint32_t)LS_CAP) {
      count += sm.tcnt[e];
      ++e;
    }
    count += sm.tlead[e];
    if (count == 0) {
      a = e;
      continue;
    }
    const uint32_t rpw = (((count + 31) >> 5) + RX_WARPS - 1) / RX_WARPS;  // rows per warp, <= RX_ROWS
    // ---- load: list entry [19:0] pos | [39:20] key | [50:40] slot - head slot | [61:51] slot - tile base
    {
      uint32_t wb = 0;  // window index of the first entry taken from tile j
      for (uint32_t j = a; j <= e; ++j) {
        const uint32_t lo = j == a ? sm.tlead[a] : 0u;
        const uint32_t hi = j == e ? sm.tlead[e] : sm.tcnt[j];
        const uint64_t* lt = lb + (uint64_t)(T0 + j) * LS_T;
        const uint32_t tb = (T0 + j) * LS_T;
        for (uint32_t i = lo + threadIdx.x; i < hi; i += RX_NT) {
          const uint64_t en = lt[i];
          const uint32_t r = wb + (i - lo);
          const uint32_t dist = (uint32_t)(en >> 40) & 0x7FFu;
          const uint32_t gs = r - dist;
          sm.stage[r] = make_uint2((uint32_t)en & RANK_MASK, (gs << 20) | ((uint32_t)(en >> 20) & RANK_MASK));
          if (dist == 0) {
            sm.ghead[r] = tb + ((uint32_t)(en >> 51) & 0x7FFu);
            ++heads_before;
          }
        }
        wb += hi - lo;
      }
      for (uint32_t r = count + threadIdx.x; r < rpw * RX_NT; r += RX_NT) sm.stage[r] = make_uint2(~0u, ~0u);
    }
    __syncthreads();
    switch (rpw) {
      case 1: rx_sort_window<1>(sm, count, s, rk, n_heads, n_unres); break;
      case 2: rx_sort_window<2>(sm, count, s, rk, n_heads, n_unres); break;
      case 3: rx_sort_window<3>(sm, count, s, rk, n_heads, n_unres); break;
      case 4: rx_sort_window<4>(sm, count, s, rk, n_heads, n_unres); break;
      case 5: rx_sort_window<5>(sm, count, s, rk, n_heads, n_unres); break;
      case 6: rx_sort_window<6>(sm, count, s, rk, n_heads, n_unres); break;
      default: rx_sort_window<7>(sm, count, s, rk, n_heads, n_unres); break;
    }
    __syncthreads();  // stage, ghead and whead are reused by the next window
    a = e;
  }
#pragma unroll
  for (int dlt = 16; dlt > 0; dlt >>= 1) {
    n_heads += __shfl_xor_sync(0xffffffffu, n_heads, dlt);
    n_unres += __shfl_xor_sync(0xffffffffu, n_unres, dlt);
    heads_before += __shfl_xor_sync(0xffffffffu, heads_before, dlt);
  }
  if (lane == 0) { sm.red[0][w] = n_heads; sm.red[1][w] = n_unres; sm.red[2][w] = heads_before; }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t hh = 0, u = 0, hb = 0;
    for (int ww = 0; ww < RX_WARPS; ++ww) { hh += sm.red[0][ww]; u += sm.red[1][ww]; hb += sm.red[2][ww]; }
    uint32_t* stp = stats + b * 4;
    if (hh > hb) atomicAdd(&stp[0], hh - hb);
    if (u) atomicAdd(&stp[2], u);
  }
}

// ------------------------------------------------------------------ last column + origPtr (slot order)
__global__ void __launch_bounds__(G_NT) k2_finish(const uint8_t* __restrict__ txt, const BlockDesc* __restrict__ desc,
                                                   const uint32_t* __restrict__ sa, uint8_t* __restrict__ last,
                                                   uint32_t* __restrict__ origptr) {
  const BlockDesc d = desc[blockIdx.y];
  const uint32_t n = d.n;
  const uint32_t base = blockIdx.x * (G_NT * 16);
  if (base >= n) return;
  const uint8_t* t = txt + d.off;
  const uint32_t* s = sa + d.off;
  // two sweeps (all SA loads, then all text gathers) keep 16 independent loads per thread in flight: the kernel is
  // bound by the latency of the gather
  uint32_t pos[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const uint32_t i = base + threadIdx.x + k * G_NT;
    pos[k] = i < n ? (s[i] & RANK_MASK) : 1u;
  }
  uint8_t c[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) c[k] = t[pos[k] == 0 ? n - 1 : pos[k] - 1];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const uint32_t i = base + threadIdx.x + k * G_NT;
    if (i < n) {
      last[d.off + i] = c[k];
      if (pos[k] == 0) origptr[blockIdx.y] = i;
    }
  }
}

// ------------------------------------------------------------------ host driver
struct OsState {
  uint32_t epoch = 0;        // status-word tag of the most recent pass (10 bits; status is cleared at 0)
  uint32_t ticket_base = 0;  // tickets handed out per block so far
};

template <int KBITS, int KSYMS, int WBITS>
static void launch_pass(Launcher& L, const uint64_t* src, uint64_t* dst, const BlockDesc* d_desc, uint32_t nb,
                        uint32_t tiles, BwtScratch& S, OsState& os, int pass, const uint32_t* d_inuse) {
  static PerDeviceOnce once;
  once.run([] {
    cudaFuncSetAttribute((const void*)k2_os_scatter<KBITS, KSYMS, WBITS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)sizeof(OsSmemT<1 << WBITS>));
  });
  if (os.epoch == 1023) {  // tag space exhausted: start over with a clean status array
    cudaMemsetAsync(S.hist, 0, (size_t)nb * S.tiles_cap * OS_BINS_MAX * sizeof(uint32_t), L.stream);
    os.epoch = 0;
  }
  ++os.epoch;
  L.launch_smem("k2_rs_scatter", k2_os_scatter<KBITS, KSYMS, WBITS>, dim3(nb, tiles), dim3(OS_NT),
                sizeof(OsSmemT<1 << WBITS>), src, dst, d_desc, S.cnt, S.oshist, S.hist, S.ticket, os.ticket_base,
                S.tiles_cap, os.epoch, pass, d_inuse);
  os.ticket_base += tiles;
}

template <int WBITS>
static void launch_pass_pf(Launcher& L, const uint64_t* src, uint64_t* dst, const BlockDesc* d_desc, uint32_t nb,
                           uint32_t tiles, BwtScratch& S, OsState& os, int pass) {
  static PerDeviceOnce once;
  once.run([] {
    cudaFuncSetAttribute((const void*)k2_os_scatter_pf<WBITS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)sizeof(OsPfSmem<1 << WBITS>));
  });
  if (os.epoch == 1023) {
    cudaMemsetAsync(S.hist, 0, (size_t)nb * S.tiles_cap * OS_BINS_MAX * sizeof(uint32_t), L.stream);
    os.epoch = 0;
  }
  ++os.epoch;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const uint64_t total = (uint64_t)nb * tiles;
  const uint32_t grid = (uint32_t)std::min<uint64_t>(total, (uint64_t)sms * OS_MINB);
  cudaMemsetAsync(S.ticket + nb, 0, sizeof(uint32_t), L.stream);  // the global ticket counter sits behind the per-block ones
  L.launch_smem("k2_rs_scatter", k2_os_scatter_pf<WBITS>, dim3(grid), dim3(OS_NT), sizeof(OsPfSmem<1 << WBITS>), src, dst,
                d_desc, S.cnt, S.oshist, S.hist, S.ticket + nb, nb, tiles, S.tiles_cap, os.epoch, pass);
}

// Passes over 64-bit elements.  BZB200_OS_PF: 0 = one tile per CTA (k2_os_scatter), 1 = the persistent TMA-fed kernel
// (k2_os_scatter_pf), 2 (default) = by measurement (profiles/README.md, round 2): the 9-bit passes are bound by the
// shared-memory pipe (MIO throttle), where reading the tile back out of shared memory costs more than the hidden load
// latency gains (1.96 ms against 1.73 ms per 256 MiB pass); the 8-bit passes (256 counters per warp) have that headroom
// and run 5 % faster persistent.
template <int WBITS>
static void launch_pass_elems(Launcher& L, const uint64_t* src, uint64_t* dst, const BlockDesc* d_desc, uint32_t nb,
                              uint32_t tiles, BwtScratch& S, OsState& os, int pass) {
  static const int pf = [] {
    const char* e = getenv("BZB200_OS_PF");
    return e ? atoi(e) : 2;
  }();
  // (a persistent CTA works through its tiles one after the other: with fewer than a few tiles per CTA — single
  // blocks, the short lists of late rounds — one tile per CTA finishes sooner)
  const uint64_t total = (uint64_t)nb * tiles;
  if ((pf == 1 || (pf == 2 && WBITS == 8 && total >= 8ull * 148 * OS_MINB)) && total < (1ull << 31))
    launch_pass_pf<WBITS>(L, src, dst, d_desc, nb, tiles, S, os, pass);
  else
    launch_pass<0, 0, WBITS>(L, src, dst, d_desc, nb, tiles, S, os, pass, nullptr);
}

// LSD sort of every block's rotations by their initial key (built from the text in pass 0).  Returns the number of
// passes; the sorted list ends up in `src`.
template <int KBITS, int KSYMS, int WBITS>
static int initial_sort_mode(Launcher& L, uint64_t*& src, uint64_t*& dst, const uint8_t* d_txt, const BlockDesc* d_desc,
                             const uint32_t* d_inuse, uint32_t nb, uint32_t nmax, BwtScratch& S, OsState& os) {
  constexpr int P = (KBITS * KSYMS + WBITS - 1) / WBITS;
  const uint32_t tiles = (nmax + OS_TILE - 1) / OS_TILE;
  const uint32_t chunk = 16 * 4096;
  const uint32_t chunks = (nmax + chunk - 1) / chunk;
  if (KBITS == 8) {
    cudaMemsetAsync(S.oshist, 0, (size_t)nb * OS_PASSES_MAX * OS_BINS_MAX * sizeof(uint32_t), L.stream);
    L.launch("k2_os_hist_txt", k2_os_hist_txt, dim3(chunks, nb), dim3(256), d_txt, d_desc, S.oshist, S.cnt, chunk);
    L.launch("k2_os_offsets", k2_os_offsets, dim3(nb), dim3(256), S.oshist);
  } else {
    constexpr int BITS = KBITS == 8 ? 4 : KBITS;  // (the byte mode never gets here; keeps the table small when instantiated)
    constexpr size_t NP = (size_t)1 << (2 * BITS);
    static PerDeviceOnce once;
    once.run([] {
      cudaFuncSetAttribute((const void*)k2_pair_hist<BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(NP * 4 + 256));
    });
    cudaMemsetAsync(S.pairhist, 0, (size_t)nb * NP * sizeof(uint32_t), L.stream);
    L.launch_smem("k2_pair_hist", k2_pair_hist<BITS>, dim3(chunks, nb), dim3(256), NP * 4 + 256, d_txt, d_desc, d_inuse,
                  S.pairhist, S.cnt, chunk);
    L.launch("k2_digit_offsets", k2_digit_offsets<BITS, KSYMS, WBITS>, dim3(nb), dim3(512), (const uint32_t*)S.pairhist,
             S.oshist);
  }
  for (int p = 0; p < P; ++p) {
    if (p == 0)
      launch_pass<KBITS, KSYMS, WBITS>(L, reinterpret_cast<const uint64_t*>(d_txt), dst, d_desc, nb, tiles, S, os, p,
                                       d_inuse);
    else
      launch_pass_elems<WBITS>(L, src, dst, d_desc, nb, tiles, S, os, p);
    uint64_t* t = src; src = dst; dst = t;
  }
  return P;
}

size_t bwt_pairhist_bytes(uint32_t nb, uint32_t max_alpha) {
  const KeyMode m = key_mode(max_alpha);
  return m.bits == 8 ? 16 : (size_t)nb * ((size_t)1 << (2 * m.bits)) * sizeof(uint32_t);
}

// One 40-bit LSD sort (five 8-bit passes) of every block's list of (group, key, pos) elements (cnt[b] elements at
// src + desc[b].off): the BIG-group path of a doubling round.
static void radix_sort40(Launcher& L, uint64_t*& src, uint64_t*& dst, const BlockDesc* d_desc, uint32_t nb,
                         uint32_t maxcnt, BwtScratch& S, OsState& os) {
  const uint32_t tiles = (maxcnt + OS_TILE - 1) / OS_TILE;
  if (tiles == 0) return;
  cudaMemsetAsync(S.oshist, 0, (size_t)nb * OS_PASSES_MAX * OS_BINS_MAX * sizeof(uint32_t), L.stream);
  const uint32_t chunk = 16 * 4096;
  const uint32_t chunks = (maxcnt + chunk - 1) / chunk;
  L.launch("k2_os_hist", k2_os_hist, dim3(chunks, nb), dim3(256), src, d_desc, S.cnt, S.oshist, chunk);
  L.launch("k2_os_offsets", k2_os_offsets, dim3(nb), dim3(256), S.oshist);
  for (int p = 0; p < 5; ++p) {
    launch_pass_elems<8>(L, src, dst, d_desc, nb, tiles, S, os, p);
    uint64_t* t = src; src = dst; dst = t;
  }
}

static void regroup(Launcher& L, const uint64_t* srt, const BlockDesc* d_desc, uint32_t nb, uint32_t maxcnt,
                    BwtScratch& S, int initial) {
  const uint32_t tiles = (maxcnt + RS_TILE - 1) / RS_TILE;
  L.launch("k2_rg_flags", k2_rg_flags, dim3(tiles, nb), dim3(RS_NT), srt, d_desc, S.cnt, S.tsum, S.tiles_cap, initial);
  L.launch("k2_rg_scan", k2_rg_scan, dim3(nb), dim3(32), S.cnt, S.tsum, S.tiles_cap);
  if (initial)
    L.launch_smem("k2_rg_apply", k2_rg_apply<true>, dim3(tiles, nb), dim3(RS_NT), (size_t)(RS_TILE + RS_TILE / 16) * 8,
                  srt, d_desc, S.cnt, S.tsum, S.tiles_cap, S.rank, S.sa, S.stats);
  else
    L.launch_smem("k2_rg_apply", k2_rg_apply<false>, dim3(tiles, nb), dim3(RS_NT), (size_t)(RS_TILE + RS_TILE / 16) * 8,
                  srt, d_desc, S.cnt, S.tsum, S.tiles_cap, S.rank, S.sa, S.stats);
}

int run_bwt(Launcher& L, const uint8_t* d_txt, const BlockDesc* d_desc, const uint32_t* d_inuse, uint32_t max_alpha,
            uint32_t nb, uint32_t nmax, uint64_t M, BwtScratch& S, uint8_t* d_last, uint32_t* d_origptr, BwtStats* stats) {
  cudaStream_t st = L.stream;
  const uint32_t ls_tiles = (nmax + LS_T - 1) / LS_T;
  cudaMemsetAsync(S.state, 0, nb * sizeof(uint32_t), st);
  cudaMemsetAsync(S.shift, 0xFF, nb * sizeof(uint32_t), st);
  cudaMemsetAsync(S.stats, 0, nb * 4 * sizeof(uint32_t), st);
  cudaMemsetAsync(S.rounds, 0, nb * sizeof(uint32_t), st);
  cudaMemsetAsync(S.sparse, 0, nb * sizeof(uint32_t), st);
  cudaMemsetAsync(S.global, 0, 4 * sizeof(uint32_t), st);

  static const int ls_enum = [] {
    const char* e = getenv("BZB200_LS_ENUM");
    return e ? atoi(e) : ENUM_MAX_DEFAULT;
  }();
  static const int ls_rx = [] {
    // 0: the enumeration group sort (k2_local_sort) in every round, 1: the radix group sort (k2_local_sort_rx) in
    // every round, 2 (default): enumeration while the lists are dense, radix windows once they are sparse
    const char* e = getenv("BZB200_LS_RX");
    return e ? atoi(e) : 2;
  }();
  static const int ls_tpc = [] {
    const char* e = getenv("BZB200_LS_TPC");  // experiments: tiles per CTA of k2_local_sort_rx (0 = by list density)
    return e ? atoi(e) : 0;
  }();
  static PerDeviceOnce once_ls;
  once_ls.run([] {
    cudaFuncSetAttribute((const void*)k2_local_sort_rx, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RxSmem));
    cudaFuncSetAttribute((const void*)k2_local_sort<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LsSmem));
    cudaFuncSetAttribute((const void*)k2_local_sort<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LsSmem));
    cudaFuncSetAttribute((const void*)k2_local_sort<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LsSmem));
  });
  uint32_t rounds = 0, passes = 0;
  uint64_t elems = 0, local_elems = 0;
  uint64_t *src = S.A, *dst = S.B;
  OsState os;
  cudaMemsetAsync(S.hist, 0, (size_t)nb * S.tiles_cap * OS_BINS_MAX * sizeof(uint32_t), st);
  cudaMemsetAsync(S.ticket, 0, nb * sizeof(uint32_t), st);
  static const int force_mode = [] {
    const char* e = getenv("BZB200_KEY_BITS");  // experiments: 8 forces the byte-key sort
    return e ? atoi(e) : 0;
  }();
  const KeyMode km = key_mode(force_mode == 8 ? 0u : max_alpha);
  int p0;
  switch (km.bits) {
    case 7: p0 = initial_sort_mode<7, 5, 9>(L, src, dst, d_txt, d_desc, d_inuse, nb, nmax, S, os); break;
    case 6: p0 = initial_sort_mode<6, 6, 9>(L, src, dst, d_txt, d_desc, d_inuse, nb, nmax, S, os); break;
    case 4: p0 = initial_sort_mode<4, 8, 8>(L, src, dst, d_txt, d_desc, d_inuse, nb, nmax, S, os); break;
    default: p0 = initial_sort_mode<8, 5, 8>(L, src, dst, d_txt, d_desc, d_inuse, nb, nmax, S, os); break;
  }
  passes += p0;
  uint64_t radix_elem_passes = (uint64_t)p0 * M;
  elems += M;
  regroup(L, src, d_desc, nb, nmax, S, 1);
  L.launch("k2_round_finalize", k2_round_finalize, dim3((nb + 255) / 256), dim3(256), nb, 0u, d_desc, S.cnt, S.stats,
           S.state, S.sparse, S.rounds, S.global);

  uint32_t g[4];
  if (L.err != cudaSuccess) return -2;
  if (cudaMemcpyAsync(g, S.global, sizeof(g), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -2;
  if (cudaStreamSynchronize(st) != cudaSuccess) return -2;

  static const bool trace = getenv("BZB200_TRACE_ROUNDS") != nullptr;  // per-round work, for the ncu tables
  if (trace)
    fprintf(stderr, "k2 initial: elems %llu passes %d key %dx%d bits -> unresolved %u max_radix %u radix %u\n",
            (unsigned long long)M, p0, km.syms, km.bits, g[0], g[1], g[2]);
  uint32_t h = (uint32_t)km.syms;  // the initial key covers the first h symbols of every rotation
  while (g[0] > 0) {
    if (g[3]) return -5;
    ++rounds;
    if (rounds > 64) return -5;
    cudaMemsetAsync(S.global, 0, 3 * sizeof(uint32_t), st);
    L.launch("k2_periodic_shift", k2_periodic_shift, dim3(nb), dim3(256), d_desc, S.state, S.sa, S.shift);
    // BIG-group elements go to S.A (both radix buffers are free between rounds)
    // small groups -> per-tile lists in S.B (free until the BIG path's radix sort, which runs after the local sort);
    // BIG-group / sparse-block elements -> S.A
    L.launch("k2_gather", k2_gather, dim3(ls_tiles, nb), dim3(G_NT), d_desc, S.rank, S.sa, S.state, S.shift, S.sparse, h,
             S.B, S.A, S.cnt, S.tile_meta, S.ls_tiles_cap);
    if (g[0] == g[2]) {
      // every unresolved rotation takes the radix path: the lists are empty
    } else if (ls_rx == 1 || (ls_rx == 2 && (g[0] - g[2]) < (uint64_t)nb * ls_tiles * (LS_T / 4) &&
                              (uint64_t)nb * ls_tiles >= 4096)) {  // (a handful of blocks: one CTA per tile finishes sooner)
      // tiles per CTA: enough of them that a window holds a few thousand entries (g[0] - g[2] entries in all lists)
      const uint64_t local_total = g[0] - g[2];
      const uint64_t avg = std::max<uint64_t>(1, local_total / std::max<uint64_t>(1, (uint64_t)nb * ls_tiles));
      uint32_t tpc = ls_tpc > 0 ? (uint32_t)ls_tpc : (uint32_t)std::min<uint64_t>(RX_TPC_MAX, (3000 + avg - 1) / avg);
      // ... but never so many that the grid no longer fills the GPU (single blocks: one tile per CTA)
      if (ls_tpc <= 0) tpc = (uint32_t)std::min<uint64_t>(tpc, std::max<uint64_t>(1, (uint64_t)nb * ls_tiles / (4 * 148)));
      tpc = std::max(1u, std::min(tpc, (uint32_t)RX_TPC_MAX));
      L.launch_smem("k2_local_sort", k2_local_sort_rx, dim3((ls_tiles + tpc - 1) / tpc, nb), dim3(RX_NT), sizeof(RxSmem),
                    d_desc, S.state, S.sparse, S.sa, S.B, S.rank, S.tile_meta, S.ls_tiles_cap, S.stats, tpc);
    } else {
      switch (ls_enum) {
        case 16:
          L.launch_smem("k2_local_sort", k2_local_sort<16>, dim3(ls_tiles, nb), dim3(LS_NT), sizeof(LsSmem), d_desc,
                        S.state, S.sparse, S.sa, S.B, S.rank, S.tile_meta, S.ls_tiles_cap, S.stats);
          break;
        case 32:
          L.launch_smem("k2_local_sort", k2_local_sort<32>, dim3(ls_tiles, nb), dim3(LS_NT), sizeof(LsSmem), d_desc,
                        S.state, S.sparse, S.sa, S.B, S.rank, S.tile_meta, S.ls_tiles_cap, S.stats);
          break;
        default:
          L.launch_smem("k2_local_sort", k2_local_sort<ENUM_MAX_DEFAULT>, dim3(ls_tiles, nb), dim3(LS_NT), sizeof(LsSmem),
                        d_desc, S.state, S.sparse, S.sa, S.B, S.rank, S.tile_meta, S.ls_tiles_cap, S.stats);
          break;
      }
    }
    const uint32_t maxbig = g[1];
    if (maxbig > 0) {
      uint64_t *s2 = S.A, *d2 = S.B;
      radix_sort40(L, s2, d2, d_desc, nb, maxbig, S, os);
      passes += 5;
      regroup(L, s2, d_desc, nb, maxbig, S, 0);
    }
    const uint32_t elems_round = g[0], radix_round = g[2];
    elems += g[0];
    radix_elem_passes += 5ull * g[2];
    local_elems += g[0] - g[2];
    L.launch("k2_round_finalize", k2_round_finalize, dim3((nb + 255) / 256), dim3(256), nb, rounds, d_desc, S.cnt,
             S.stats, S.state, S.sparse, S.rounds, S.global);
    if (L.err != cudaSuccess) return -2;
    if (cudaMemcpyAsync(g, S.global, sizeof(g), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -2;
    if (cudaStreamSynchronize(st) != cudaSuccess) return -2;
    if (trace)
      fprintf(stderr, "k2 round %u (h %u): local %u radix %u -> unresolved %u max_radix %u radix %u\n", rounds, h,
              (uint32_t)(elems_round - radix_round), radix_round, g[0], g[1], g[2]);
    if (h < (1u << 21)) h *= 2;
  }
  if (g[3]) return -5;
  const uint32_t ftiles = (nmax + G_NT * 16 - 1) / (G_NT * 16);
  L.launch("k2_finish", k2_finish, dim3(ftiles, nb), dim3(G_NT), d_txt, d_desc, S.sa, d_last, d_origptr);
  if (stats) {
    stats->rounds = std::max(stats->rounds, rounds);
    stats->radix_passes += passes;
    stats->elems_sorted += elems;
    stats->radix_elem_passes += radix_elem_passes;
    stats->local_elems += local_elems;
  }
  return L.err == cudaSuccess ? 0 : -2;
}

}  // namespace bzb
