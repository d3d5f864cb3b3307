// k2_bwt.cu — K2: the BWT as a batched GPU prefix-doubling sort of the cyclic rotations of every block.
//
// Replaces suffix_array::sais::bwt (src/suffix_array/sais.rs:266-272 = least-rotation pre-pass :12-68 + cyclic
// SA-IS :127-264) by result, not by method.  Order to reproduce (pinned by
// tests/test_oracle.py::test_bwt_matches_rotation_model): rotations in lexicographic order; equal rotations
// (periodic block) by DESCENDING (pos - shift) mod n, shift = smallest start of a minimal rotation;
// origPtr = rank of rotation 0 (src/bzip2/encoder.rs:332-334).
//
// Method: one 64-bit sort element per unresolved rotation,  [59:40] g | [39:20] k2 | [19:0] pos
//   round 0 : g|k2 = the rotation's first 5 bytes (40 bits)           -> classes by 5-byte prefix
//   round r : g = rank_h[pos] (slot of the class head), k2 = rank_h[(pos+h) mod n]   -> classes by 2h prefix
//   fix-up  : k2 = n-1-((pos-shift) mod n) once a block's partition stops refining (periodic block)
// Each round: compact unresolved rotations in position order (coalesced), LSD radix sort on bits 20..59
// (5 passes of 8 bits, per-block segments, stable), then regroup: new rank = g + (start of (g,k2) run - start of
// g run); singletons are marked resolved and leave the active set.  rank[] is the only persistent state; the
// last column is scattered from it at the end.  All blocks of the batch advance together.
#include "common.cuh"
#include "kernels.h"

namespace bzb {

constexpr int RS_NT = 256;
constexpr int RS_IPT = 16;
constexpr int RS_TILE = RS_NT * RS_IPT;  // 4096 elements per CTA
constexpr int RS_WARPS = RS_NT / 32;
constexpr int RS_WCH = RS_TILE / RS_WARPS;  // 512 elements per warp
constexpr int KEY_LO = 20;                  // key field = bits 20..59
constexpr uint64_t POS_MASK = 0xFFFFFull;

uint32_t bwt_tile_elems() { return RS_TILE; }

// ------------------------------------------------------------------ round 0 keys
__global__ void __launch_bounds__(RS_NT) k2_init_keys(const uint8_t* __restrict__ txt, const BlockDesc* __restrict__ desc,
                                                      uint64_t* __restrict__ A, uint32_t* __restrict__ cnt) {
  const BlockDesc d = desc[blockIdx.y];
  const uint32_t n = d.n;
  const uint32_t base = blockIdx.x * RS_TILE;
  if (base >= n) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) cnt[blockIdx.y] = n;
  const uint8_t* t = txt + d.off;
#pragma unroll 4
  for (int k = 0; k < RS_IPT; ++k) {
    uint32_t pos = base + threadIdx.x + k * RS_NT;
    if (pos < n) {
      uint64_t key = 0;
      uint32_t idx = pos;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        key = (key << 8) | t[idx];
        ++idx;
        if (idx == n) idx = 0;
      }
      A[(uint64_t)d.off + pos] = (key << KEY_LO) | pos;
    }
  }
}

// ------------------------------------------------------------------ LSD radix pass (per-block segments)
__device__ __forceinline__ uint32_t match_digit(uint32_t dgt, bool valid) {
  uint32_t m = __ballot_sync(0xffffffffu, valid);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    uint32_t bit = (dgt >> k) & 1u;
    uint32_t bk = __ballot_sync(0xffffffffu, bit);
    m &= bit ? bk : ~bk;
  }
  return m;
}

__global__ void __launch_bounds__(RS_NT) k2_rs_hist(const uint64_t* __restrict__ src, const BlockDesc* __restrict__ desc,
                                                    const uint32_t* __restrict__ cnt, uint32_t* __restrict__ hist,
                                                    uint32_t tiles_cap, int shift) {
  __shared__ uint32_t h[256];
  const uint32_t c = cnt[blockIdx.y];
  const uint32_t base = blockIdx.x * RS_TILE;
  if (base >= c) return;
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t* s = src + desc[blockIdx.y].off;
#pragma unroll 4
  for (int k = 0; k < RS_IPT; ++k) {
    uint32_t i = base + threadIdx.x + k * RS_NT;
    bool valid = i < c;
    uint32_t dgt = valid ? (uint32_t)(s[i] >> shift) & 255u : 0u;
    uint32_t peers = match_digit(dgt, valid);
    if (valid && (peers & lanemask_lt()) == 0) atomicAdd(&h[dgt], __popc(peers));
  }
  __syncthreads();
  hist[((uint64_t)blockIdx.y * tiles_cap + blockIdx.x) * 256 + threadIdx.x] = h[threadIdx.x];
}

// One CTA per block: turns per-(tile,digit) counts into exclusive offsets in (digit, tile) order.
__global__ void __launch_bounds__(256) k2_rs_scan(const uint32_t* __restrict__ cnt, uint32_t* __restrict__ hist,
                                                  uint32_t tiles_cap) {
  __shared__ uint32_t ws[256 / 32 + 1];
  const uint32_t c = cnt[blockIdx.x];
  if (c == 0) return;
  const uint32_t nt = (c + RS_TILE - 1) / RS_TILE;
  uint32_t* hb = hist + (uint64_t)blockIdx.x * tiles_cap * 256;
  uint32_t tot = 0;
  for (uint32_t t = 0; t < nt; ++t) tot += hb[t * 256 + threadIdx.x];
  uint32_t run = cta_excl_scan_add<256>(tot, ws, nullptr);
  for (uint32_t t = 0; t < nt; ++t) {
    uint32_t v = hb[t * 256 + threadIdx.x];
    hb[t * 256 + threadIdx.x] = run;
    run += v;
  }
}

__global__ void __launch_bounds__(RS_NT) k2_rs_scatter(const uint64_t* __restrict__ src, uint64_t* __restrict__ dst,
                                                       const BlockDesc* __restrict__ desc,
                                                       const uint32_t* __restrict__ cnt,
                                                       const uint32_t* __restrict__ hist, uint32_t tiles_cap, int shift) {
  __shared__ uint32_t wcnt[RS_WARPS][256];
  __shared__ uint64_t stage[RS_TILE];
  __shared__ uint32_t dstart[256];
  __shared__ int goff[256];
  __shared__ uint32_t ws[RS_NT / 32 + 1];
  const uint32_t c = cnt[blockIdx.y];
  const uint32_t base = blockIdx.x * RS_TILE;
  if (base >= c) return;
  const uint32_t tcount = min((uint32_t)RS_TILE, c - base);
  const uint32_t off = desc[blockIdx.y].off;
  const uint64_t* s = src + off;
  const int w = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  for (int i = lane; i < 256; i += 32) wcnt[w][i] = 0;
  __syncwarp();

  uint64_t e[RS_IPT];
  uint16_t rk[RS_IPT];
#pragma unroll
  for (int it = 0; it < RS_IPT; ++it) {
    uint32_t li = w * RS_WCH + it * 32 + lane;  // index inside the tile; memory order == (warp, it, lane)
    bool valid = li < tcount;
    e[it] = valid ? s[base + li] : 0ull;
    uint32_t dgt = (uint32_t)(e[it] >> shift) & 255u;
    uint32_t peers = match_digit(dgt, valid);
    uint32_t old = valid ? wcnt[w][dgt] : 0u;
    __syncwarp();
    if (valid && (peers & lanemask_lt()) == 0) wcnt[w][dgt] = old + __popc(peers);
    __syncwarp();
    rk[it] = (uint16_t)(old + __popc(peers & lanemask_lt()));
  }
  __syncthreads();
  // thread d: exclusive prefix over warps for digit d, and the tile total
  {
    const int dgt = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int ww = 0; ww < RS_WARPS; ++ww) {
      uint32_t t = wcnt[ww][dgt];
      wcnt[ww][dgt] = run;
      run += t;
    }
    uint32_t ds = cta_excl_scan_add<RS_NT>(run, ws, nullptr);
    dstart[dgt] = ds;
    goff[dgt] = (int)hist[((uint64_t)blockIdx.y * tiles_cap + blockIdx.x) * 256 + dgt] - (int)ds;
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < RS_IPT; ++it) {
    uint32_t li = w * RS_WCH + it * 32 + lane;
    if (li < tcount) {
      uint32_t dgt = (uint32_t)(e[it] >> shift) & 255u;
      stage[dstart[dgt] + wcnt[w][dgt] + rk[it]] = e[it];
    }
  }
  __syncthreads();
  uint64_t* o = dst + off;
#pragma unroll 4
  for (int k = 0; k < RS_IPT; ++k) {
    uint32_t i = threadIdx.x + k * RS_NT;
    if (i < tcount) {
      uint64_t v = stage[i];
      uint32_t dgt = (uint32_t)(v >> shift) & 255u;
      o[goff[dgt] + (int)i] = v;
    }
  }
}

// ------------------------------------------------------------------ regroup
// Tile summaries: index of the last (g,k2)-run head and of the last g-run head inside the tile (or -1).
__global__ void __launch_bounds__(RS_NT) k2_rg_flags(const uint64_t* __restrict__ srt, const BlockDesc* __restrict__ desc,
                                                     const uint32_t* __restrict__ cnt, int2* __restrict__ tsum,
                                                     uint32_t tiles_cap, int initial) {
  __shared__ int sh[RS_WARPS], sg[RS_WARPS];
  const uint32_t c = cnt[blockIdx.y];
  const uint32_t base = blockIdx.x * RS_TILE;
  if (base >= c) return;
  const uint64_t* s = srt + desc[blockIdx.y].off;
  int lh = -1, lg = -1;
#pragma unroll 4
  for (int k = 0; k < RS_IPT; ++k) {
    uint32_t a = base + threadIdx.x + k * RS_NT;
    if (a < c) {
      uint64_t cur = s[a] >> KEY_LO;
      uint64_t prv = a > 0 ? (s[a - 1] >> KEY_LO) : ~0ull;
      if (cur != prv) lh = (int)a;
      bool hg = initial ? (a == 0) : ((cur >> 20) != (prv >> 20));
      if (hg) lg = (int)a;
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    lh = max(lh, __shfl_xor_sync(0xffffffffu, lh, d));
    lg = max(lg, __shfl_xor_sync(0xffffffffu, lg, d));
  }
  if (lane_id() == 0) { sh[threadIdx.x >> 5] = lh; sg[threadIdx.x >> 5] = lg; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < RS_WARPS; ++w) { lh = max(lh, sh[w]); lg = max(lg, sg[w]); }
    tsum[(uint64_t)blockIdx.y * tiles_cap + blockIdx.x] = make_int2(lh, lg);
  }
}

// One warp per block: exclusive max-scan of the tile summaries (in place).
__global__ void __launch_bounds__(32) k2_rg_scan(const uint32_t* __restrict__ cnt, int2* __restrict__ tsum,
                                                 uint32_t tiles_cap) {
  const uint32_t c = cnt[blockIdx.x];
  if (c == 0) return;
  const uint32_t nt = (c + RS_TILE - 1) / RS_TILE;
  int2* ts = tsum + (uint64_t)blockIdx.x * tiles_cap;
  int ch = -1, cg = -1;
  for (uint32_t t0 = 0; t0 < nt; t0 += 32) {
    uint32_t t = t0 + lane_id();
    int2 v = t < nt ? ts[t] : make_int2(-1, -1);
    int ih = warp_incl_scan_max(v.x), ig = warp_incl_scan_max(v.y);
    int eh = __shfl_up_sync(0xffffffffu, ih, 1), eg = __shfl_up_sync(0xffffffffu, ig, 1);
    if (lane_id() == 0) { eh = -1; eg = -1; }
    eh = max(eh, ch); eg = max(eg, cg);
    if (t < nt) ts[t] = make_int2(eh, eg);
    ch = max(ch, __shfl_sync(0xffffffffu, ih, 31));
    cg = max(cg, __shfl_sync(0xffffffffu, ig, 31));
  }
}

// New ranks, resolved flags, per-block statistics.
__global__ void __launch_bounds__(RS_NT) k2_rg_apply(const uint64_t* __restrict__ srt, const BlockDesc* __restrict__ desc,
                                                     const uint32_t* __restrict__ cnt, const int2* __restrict__ tsum,
                                                     uint32_t tiles_cap, int initial, uint32_t* __restrict__ rank,
                                                     uint32_t* __restrict__ stats, uint32_t* __restrict__ shift) {
  __shared__ int wsh[RS_WARPS], wsg[RS_WARPS];
  __shared__ uint32_t red[4][RS_WARPS];
  const uint32_t c = cnt[blockIdx.y];
  const uint32_t base = blockIdx.x * RS_TILE;
  if (base >= c) return;
  const uint32_t off = desc[blockIdx.y].off;
  const uint64_t* s = srt + off;
  uint32_t* rk = rank + off;
  const int2 carry = tsum[(uint64_t)blockIdx.y * tiles_cap + blockIdx.x];

  // blocked arrangement: thread t owns elements base + t*16 .. +15
  const uint32_t a0 = base + threadIdx.x * RS_IPT;
  uint64_t e[RS_IPT + 1];
  uint64_t prv = ~0ull;
  if (a0 < c) {
    if (a0 > 0) prv = s[a0 - 1] >> KEY_LO;
#pragma unroll
    for (int j = 0; j <= RS_IPT; ++j) e[j] = (a0 + j < c) ? s[a0 + j] : ~0ull;
  }
  int lh = -1, lg = -1;
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
  }
  // CTA exclusive max-scan of (lh, lg)
  int ih = warp_incl_scan_max(lh), ig = warp_incl_scan_max(lg);
  const int w = threadIdx.x >> 5;
  if (lane_id() == 31) { wsh[w] = ih; wsg[w] = ig; }
  __syncthreads();
  int eh = __shfl_up_sync(0xffffffffu, ih, 1), eg = __shfl_up_sync(0xffffffffu, ig, 1);
  if (lane_id() == 0) { eh = -1; eg = -1; }
  for (int ww = 0; ww < w; ++ww) { eh = max(eh, wsh[ww]); eg = max(eg, wsg[ww]); }
  eh = max(eh, carry.x);
  eg = max(eg, carry.y);

  uint32_t n_h = 0, n_g = 0, n_unres = 0, min_pos0 = 0xFFFFFFFFu;
  if (a0 < c) {
    int ha = eh, fa = eg;
#pragma unroll
    for (int j = 0; j < RS_IPT; ++j) {
      if (a0 + j < c) {
        const int a = (int)(a0 + j);
        if (hmask & (1u << j)) { ha = a; ++n_h; }
        if (gmask & (1u << j)) { fa = a; ++n_g; }
        const uint64_t cur = e[j] >> KEY_LO;
        const uint32_t pos = (uint32_t)(e[j] & POS_MASK);
        const uint32_t g = initial ? 0u : (uint32_t)(cur >> 20);
        const uint32_t nr = g + (uint32_t)(ha - fa);
        // singleton <=> this element heads its run and the next element (if any) heads another
        const bool next_head = (a0 + j + 1 >= c) || ((e[j + 1] >> KEY_LO) != cur);
        const bool single = (ha == a) && next_head;
        rk[pos] = nr | (single ? RANK_RESOLVED : 0u);
        if (!single) ++n_unres;
        if (nr == 0) min_pos0 = min(min_pos0, pos);
      }
    }
  }
  // CTA reductions -> one atomic per statistic
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    n_h += __shfl_xor_sync(0xffffffffu, n_h, d);
    n_g += __shfl_xor_sync(0xffffffffu, n_g, d);
    n_unres += __shfl_xor_sync(0xffffffffu, n_unres, d);
    min_pos0 = min(min_pos0, __shfl_xor_sync(0xffffffffu, min_pos0, d));
  }
  if (lane_id() == 0) { red[0][w] = n_h; red[1][w] = n_g; red[2][w] = n_unres; red[3][w] = min_pos0; }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t a = 0, b = 0, u = 0, m = 0xFFFFFFFFu;
    for (int ww = 0; ww < RS_WARPS; ++ww) { a += red[0][ww]; b += red[1][ww]; u += red[2][ww]; m = min(m, red[3][ww]); }
    uint32_t* st = stats + blockIdx.y * 4;
    if (a) atomicAdd(&st[0], a);
    if (b) atomicAdd(&st[1], b);
    if (u) atomicAdd(&st[2], u);
    if (m != 0xFFFFFFFFu) atomicMin(&shift[blockIdx.y], m);
  }
}

// Per-block state machine after a round. state: 0 active, 1 fix-up pending, 2 done.
__global__ void k2_round_finalize(uint32_t nb, uint32_t round_no, uint32_t* __restrict__ cnt,
                                  uint32_t* __restrict__ stats, uint32_t* __restrict__ state,
                                  uint32_t* __restrict__ shift, uint32_t* __restrict__ rounds,
                                  uint32_t* __restrict__ global) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  uint32_t* st = stats + b * 4;
  const uint32_t heads_h = st[0], heads_g = st[1], unres = st[2];
  uint32_t s = state[b];
  if (s != 2) {
    if (unres == 0) {
      s = 2;
      rounds[b] = round_no;
    } else if (s == 1) {
      atomicOr(&global[2], 1u);  // the fix-up key makes every rotation distinct; anything else is a bug
      s = 2;
    } else if (heads_h == heads_g) {
      s = 1;  // the partition did not refine: the block is periodic, classes are sets of equal rotations
      st[3] = 1;
    } else {
      shift[b] = 0xFFFFFFFFu;
    }
    state[b] = s;
    if (s != 2) {
      atomicAdd(&global[0], unres);
      atomicMax(&global[1], unres);
    }
  }
  st[0] = 0; st[1] = 0; st[2] = 0;
  cnt[b] = 0;
}

// ------------------------------------------------------------------ next round's elements
__global__ void __launch_bounds__(RS_NT) k2_build_active(const BlockDesc* __restrict__ desc,
                                                         const uint32_t* __restrict__ rank,
                                                         const uint32_t* __restrict__ state,
                                                         const uint32_t* __restrict__ shiftv, uint32_t h,
                                                         uint64_t* __restrict__ A, uint32_t* __restrict__ cnt) {
  __shared__ uint32_t ws[RS_NT / 32 + 1];
  __shared__ uint32_t s_base;
  const BlockDesc d = desc[blockIdx.y];
  const uint32_t n = d.n;
  const uint32_t base = blockIdx.x * RS_TILE;
  if (base >= n) return;
  const uint32_t st = state[blockIdx.y];
  if (st == 2) return;
  const uint32_t* rk = rank + d.off;
  const uint32_t hm = h % n;
  const uint32_t sh = st == 1 ? shiftv[blockIdx.y] : 0u;
  // blocked: thread t owns positions base + t*16 .. +15 (keeps the compaction order-preserving inside the CTA)
  const uint32_t p0 = base + threadIdx.x * RS_IPT;
  uint64_t e[RS_IPT];
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < RS_IPT; ++j) {
    uint32_t pos = p0 + j;
    if (pos < n) {
      uint32_t r = rk[pos];
      if (!(r & RANK_RESOLVED)) {
        uint32_t k2;
        if (st == 0) {
          uint32_t p2 = pos + hm;
          if (p2 >= n) p2 -= n;
          k2 = rk[p2] & RANK_MASK;
        } else {
          uint32_t rel = pos >= sh ? pos - sh : pos + n - sh;  // (pos - shift) mod n
          k2 = n - 1 - rel;
        }
        e[j] = ((uint64_t)(r & RANK_MASK) << 40) | ((uint64_t)k2 << 20) | pos;
        m |= 1u << j;
      }
    }
  }
  uint32_t total;
  uint32_t ex = cta_excl_scan_add<RS_NT>(__popc(m), ws, &total);
  if (threadIdx.x == 0) s_base = total ? atomicAdd(&cnt[blockIdx.y], total) : 0u;
  __syncthreads();
  uint64_t* o = A + d.off + s_base + ex;
#pragma unroll
  for (int j = 0; j < RS_IPT; ++j)
    if (m & (1u << j)) *o++ = e[j];
}

// ------------------------------------------------------------------ last column + origPtr
__global__ void __launch_bounds__(RS_NT) k2_finish(const uint8_t* __restrict__ txt, const BlockDesc* __restrict__ desc,
                                                   const uint32_t* __restrict__ rank, uint8_t* __restrict__ last,
                                                   uint32_t* __restrict__ origptr) {
  const BlockDesc d = desc[blockIdx.y];
  const uint32_t n = d.n;
  const uint32_t base = blockIdx.x * RS_TILE;
  if (base >= n) return;
  const uint8_t* t = txt + d.off;
#pragma unroll 4
  for (int k = 0; k < RS_IPT; ++k) {
    uint32_t pos = base + threadIdx.x + k * RS_NT;
    if (pos < n) {
      uint32_t r = rank[d.off + pos] & RANK_MASK;
      last[d.off + r] = t[pos == 0 ? n - 1 : pos - 1];
      if (pos == 0) origptr[blockIdx.y] = r;
    }
  }
}

// ------------------------------------------------------------------ host driver
static void radix_sort40(Launcher& L, uint64_t*& src, uint64_t*& dst, const BlockDesc* d_desc, uint32_t nb,
                         uint32_t maxcnt, BwtScratch& S) {
  const uint32_t tiles = (maxcnt + RS_TILE - 1) / RS_TILE;
  for (int p = 0; p < 5; ++p) {
    const int shift = KEY_LO + 8 * p;
    L.launch("k2_rs_hist", k2_rs_hist, dim3(tiles, nb), dim3(RS_NT), src, d_desc, S.cnt, S.hist, S.tiles_cap, shift);
    L.launch("k2_rs_scan", k2_rs_scan, dim3(nb), dim3(256), S.cnt, S.hist, S.tiles_cap);
    L.launch("k2_rs_scatter", k2_rs_scatter, dim3(tiles, nb), dim3(RS_NT), src, dst, d_desc, S.cnt, S.hist,
             S.tiles_cap, shift);
    uint64_t* t = src; src = dst; dst = t;
  }
}

static void regroup(Launcher& L, const uint64_t* srt, const BlockDesc* d_desc, uint32_t nb, uint32_t maxcnt,
                    BwtScratch& S, int initial, uint32_t round_no) {
  const uint32_t tiles = (maxcnt + RS_TILE - 1) / RS_TILE;
  L.launch("k2_rg_flags", k2_rg_flags, dim3(tiles, nb), dim3(RS_NT), srt, d_desc, S.cnt, S.tsum, S.tiles_cap, initial);
  L.launch("k2_rg_scan", k2_rg_scan, dim3(nb), dim3(32), S.cnt, S.tsum, S.tiles_cap);
  L.launch("k2_rg_apply", k2_rg_apply, dim3(tiles, nb), dim3(RS_NT), srt, d_desc, S.cnt, S.tsum, S.tiles_cap, initial,
           S.rank, S.stats, S.shift);
  L.launch("k2_round_finalize", k2_round_finalize, dim3((nb + 255) / 256), dim3(256), nb, round_no, S.cnt, S.stats,
           S.state, S.shift, S.rounds, S.global);
}

int run_bwt(Launcher& L, const uint8_t* d_txt, const BlockDesc* d_desc, uint32_t nb, uint32_t nmax, uint64_t M,
            BwtScratch& S, uint8_t* d_last, uint32_t* d_origptr, uint32_t* h_rounds, uint32_t* h_passes,
            uint64_t* h_elems) {
  cudaStream_t st = L.stream;
  const uint32_t tiles_n = (nmax + RS_TILE - 1) / RS_TILE;
  cudaMemsetAsync(S.state, 0, nb * sizeof(uint32_t), st);
  cudaMemsetAsync(S.shift, 0xFF, nb * sizeof(uint32_t), st);
  cudaMemsetAsync(S.stats, 0, nb * 4 * sizeof(uint32_t), st);
  cudaMemsetAsync(S.rounds, 0, nb * sizeof(uint32_t), st);
  cudaMemsetAsync(S.global, 0, 4 * sizeof(uint32_t), st);

  uint32_t rounds = 0, passes = 0;
  uint64_t elems = 0;
  uint64_t *src = S.A, *dst = S.B;
  L.launch("k2_init_keys", k2_init_keys, dim3(tiles_n, nb), dim3(RS_NT), d_txt, d_desc, S.A, S.cnt);
  radix_sort40(L, src, dst, d_desc, nb, nmax, S);
  passes += 5;
  elems += M;
  regroup(L, src, d_desc, nb, nmax, S, 1, 0);

  uint32_t g[4];
  if (L.err != cudaSuccess) return -2;
  if (cudaMemcpyAsync(g, S.global, sizeof(g), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -2;
  if (cudaStreamSynchronize(st) != cudaSuccess) return -2;

  uint32_t h = 5;
  while (g[0] > 0) {
    if (g[2]) return -5;
    ++rounds;
    if (rounds > 64) return -5;
    const uint32_t maxact = g[1];
    elems += g[0];
    cudaMemsetAsync(S.global, 0, 2 * sizeof(uint32_t), st);
    // the sorted result of the previous round lives in `src`; build the new list into the other buffer
    uint64_t* build = dst;
    L.launch("k2_build_active", k2_build_active, dim3(tiles_n, nb), dim3(RS_NT), d_desc, S.rank, S.state, S.shift, h,
             build, S.cnt);
    uint64_t *s2 = build, *d2 = src;
    radix_sort40(L, s2, d2, d_desc, nb, maxact, S);
    passes += 5;
    src = s2;
    dst = d2;
    regroup(L, src, d_desc, nb, maxact, S, 0, rounds);
    if (L.err != cudaSuccess) return -2;
    if (cudaMemcpyAsync(g, S.global, sizeof(g), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -2;
    if (cudaStreamSynchronize(st) != cudaSuccess) return -2;
    if (h < (1u << 21)) h *= 2;
  }
  if (g[2]) return -5;
  L.launch("k2_finish", k2_finish, dim3(tiles_n, nb), dim3(RS_NT), d_txt, d_desc, S.rank, d_last, d_origptr);
  if (h_rounds) *h_rounds = rounds;
  if (h_passes) *h_passes = passes;
  if (h_elems) *h_elems = elems;
  return L.err == cudaSuccess ? 0 : -2;
}

}  // namespace bzb
