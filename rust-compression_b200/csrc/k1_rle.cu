// k1_rle.cu — K1 (RLE1 run collapsing + greedy block cutting as scans) and K5 (per-block CRC-32/BZIP2),
// plus the per-block in-use byte map.
//
// Semantics reproduced (not code): EncoderInner::next / write_rle, src/bzip2/encoder.rs:671-716; the block
// cut test :692-696 with T = level*100000-19 (:186); crc32.rs:58-72,82-84,129-131.
//
// Parallel formulation:
//   head(i)  = i==0 || in[i]!=in[i-1]              s(i) = last head <= i      q(i) = (i-s(i)) mod 255
//   tail(i)  = i==N-1 || in[i+1]!=in[i]            pe(i) = tail(i) || q(i)==254          (piece end)
//   byte i emits a literal iff q<4, and the count byte q-3 iff pe(i) && q>=3.
//   E(i) = inclusive prefix sum of emitted bytes.  A block that starts at emitted offset S is closed after the
//   first piece end i with E(i)-S >= T, unless i is the last byte of the input.
#include "common.cuh"
#include "kernels.h"

namespace bzb {

constexpr int K1_NT = 256;
constexpr int K1_BPT = 16;
constexpr int K1_TILE = K1_NT * K1_BPT;  // 4096 input bytes per CTA

struct ThreadBytes {
  uint8_t b[K1_BPT];
  uint8_t prev;   // in[i0-1] (undefined if i0==0)
  uint8_t next;   // in[i0+cnt] (undefined if i0+cnt==N)
  int cnt;        // valid bytes (0..16)
  uint64_t i0;
};

__device__ __forceinline__ void load_thread_bytes(const uint8_t* __restrict__ in, uint64_t N, uint64_t tile_base,
                                                  ThreadBytes& tb) {
  uint64_t i0 = tile_base + (uint64_t)threadIdx.x * K1_BPT;
  tb.i0 = i0;
  if (i0 >= N) { tb.cnt = 0; tb.prev = 0; tb.next = 0; return; }
  int cnt = (int)min((uint64_t)K1_BPT, N - i0);
  tb.cnt = cnt;
  const uint8_t* p = in + i0;
  const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 15);
  if (cnt == K1_BPT && mis == 0) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < K1_BPT; ++j) tb.b[j] = (uint8_t)(w[j >> 2] >> ((j & 3) * 8));
  } else if (cnt == K1_BPT && i0 + 2 * K1_BPT <= N) {
    // unaligned input (a plan over a sub-range of a buffer): two aligned 128-bit loads + funnel shifts; the second
    // load stays inside the input because at least 16 more bytes follow
    const uint4* q = reinterpret_cast<const uint4*>(p - mis);
    const uint4 v0 = __ldg(q), v1 = __ldg(q + 1);
    const uint32_t bs = (mis & 3) * 8;
    uint32_t w[4];
    switch (mis >> 2) {  // the same for every thread of the grid
      case 0: w[0] = __funnelshift_r(v0.x, v0.y, bs); w[1] = __funnelshift_r(v0.y, v0.z, bs);
              w[2] = __funnelshift_r(v0.z, v0.w, bs); w[3] = __funnelshift_r(v0.w, v1.x, bs); break;
      case 1: w[0] = __funnelshift_r(v0.y, v0.z, bs); w[1] = __funnelshift_r(v0.z, v0.w, bs);
              w[2] = __funnelshift_r(v0.w, v1.x, bs); w[3] = __funnelshift_r(v1.x, v1.y, bs); break;
      case 2: w[0] = __funnelshift_r(v0.z, v0.w, bs); w[1] = __funnelshift_r(v0.w, v1.x, bs);
              w[2] = __funnelshift_r(v1.x, v1.y, bs); w[3] = __funnelshift_r(v1.y, v1.z, bs); break;
      default: w[0] = __funnelshift_r(v0.w, v1.x, bs); w[1] = __funnelshift_r(v1.x, v1.y, bs);
               w[2] = __funnelshift_r(v1.y, v1.z, bs); w[3] = __funnelshift_r(v1.z, v1.w, bs); break;
    }
#pragma unroll
    for (int j = 0; j < K1_BPT; ++j) tb.b[j] = (uint8_t)(w[j >> 2] >> ((j & 3) * 8));
  } else {
#pragma unroll
    for (int j = 0; j < K1_BPT; ++j) tb.b[j] = j < cnt ? __ldg(p + j) : 0;
  }
  tb.prev = i0 > 0 ? __ldg(p - 1) : 0;
  tb.next = (i0 + cnt < N) ? __ldg(p + cnt) : 0;
}

// Global index of the last head inside this thread's bytes, or -1.
__device__ __forceinline__ long long thread_last_head(const ThreadBytes& tb) {
  long long lh = -1;
#pragma unroll
  for (int j = 0; j < K1_BPT; ++j) {
    if (j < tb.cnt) {
      uint8_t pb = j == 0 ? tb.prev : tb.b[j - 1];
      bool head = (tb.i0 + j == 0) || (tb.b[j] != pb);
      if (head) lh = (long long)(tb.i0 + j);
    }
  }
  return lh;
}

struct ThreadEval {
  uint32_t lit_mask;  // bit j: byte j emits its literal
  uint32_t cb_mask;   // bit j: byte j also emits a count byte
  uint32_t q_first;   // q of byte 0 (for recomputing count-byte values)
  uint32_t count;     // emitted bytes of this thread
};

// S = last head strictly before i0 (only used when byte 0 is not a head).
__device__ __forceinline__ void eval_thread(const ThreadBytes& tb, uint64_t N, long long S, ThreadEval& ev,
                                            uint32_t* q_out /*[16] or nullptr*/) {
  ev.lit_mask = 0; ev.cb_mask = 0; ev.count = 0; ev.q_first = 0;
  if (tb.cnt == 0) return;
  uint32_t q = 0;
  if (tb.i0 > 0) q = (uint32_t)((tb.i0 - 1 - (uint64_t)S) % 255ull);  // q of byte i0-1 (valid if byte 0 is no head)
#pragma unroll
  for (int j = 0; j < K1_BPT; ++j) {
    if (j < tb.cnt) {
      uint64_t i = tb.i0 + j;
      uint8_t pb = j == 0 ? tb.prev : tb.b[j - 1];
      bool head = (i == 0) || (tb.b[j] != pb);
      q = head ? 0u : (q == 254u ? 0u : q + 1u);
      uint8_t nb = (j + 1 < tb.cnt) ? tb.b[j + 1] : tb.next;
      bool tail = (i == N - 1) || (nb != tb.b[j]);
      bool pe = tail || q == 254u;
      if (j == 0) ev.q_first = q;
      if (q_out) q_out[j] = q;
      if (q < 4u) ev.lit_mask |= 1u << j;
      if (pe && q >= 3u) ev.cb_mask |= 1u << j;
    }
  }
  ev.count = __popc(ev.lit_mask) + __popc(ev.cb_mask);
}

// ---- kernel A: last head per tile ----
__global__ void __launch_bounds__(K1_NT) k1_tile_heads(const uint8_t* __restrict__ in, uint64_t N, uint64_t tile0,
                                                       long long* __restrict__ tile_last_head) {
  __shared__ long long ws[K1_NT / 32];
  ThreadBytes tb;
  const uint64_t tile = tile0 + blockIdx.x;
  load_thread_bytes(in, N, tile * K1_TILE, tb);
  long long lh = thread_last_head(tb);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) lh = max(lh, __shfl_xor_sync(0xffffffffu, lh, d));
  if (lane_id() == 0) ws[threadIdx.x >> 5] = lh;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long m = -1;
    for (int w = 0; w < K1_NT / 32; ++w) m = max(m, ws[w]);
    tile_last_head[tile] = m;
  }
}

// ---- generic single-CTA scans over per-tile arrays ----
constexpr int SC_NT = 1024;
// `init` = value carried into element 0 (a slice of a longer array: the maximum of everything before the slice).
__global__ void __launch_bounds__(SC_NT) k_scan_max64_excl(const long long* __restrict__ in, long long* __restrict__ out,
                                                           uint64_t n, long long init) {
  __shared__ long long ws[SC_NT / 32];
  uint64_t per = (n + SC_NT - 1) / SC_NT;
  uint64_t lo = min(n, per * threadIdx.x), hi = min(n, lo + per);
  long long m = -1;
  for (uint64_t i = lo; i < hi; ++i) m = max(m, in[i]);
  long long carry = max(cta_excl_scan_max64<SC_NT>(m, ws), init);
  for (uint64_t i = lo; i < hi; ++i) {
    long long v = in[i];
    out[i] = carry;
    carry = max(carry, v);
  }
}

// out has n+1 entries: exclusive prefix, out[n] = total; `init` is added to all of them (a slice of a longer array).
__global__ void __launch_bounds__(SC_NT) k_scan_add64_excl(const uint32_t* __restrict__ in, uint64_t* __restrict__ out,
                                                           uint64_t n, unsigned long long init) {
  __shared__ unsigned long long ws[SC_NT / 32 + 1];
  uint64_t per = (n + SC_NT - 1) / SC_NT;
  uint64_t lo = min(n, per * threadIdx.x), hi = min(n, lo + per);
  unsigned long long s = 0;
  for (uint64_t i = lo; i < hi; ++i) s += in[i];
  // CTA exclusive scan (64-bit)
  unsigned long long inc = s;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d);
    if ((int)lane_id() >= d) inc += t;
  }
  int w = threadIdx.x >> 5;
  if (lane_id() == 31) ws[w] = inc;
  __syncthreads();
  if (w == 0) {
    unsigned long long x = ws[lane_id()];
    unsigned long long xi = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      unsigned long long t = __shfl_up_sync(0xffffffffu, xi, d);
      if ((int)lane_id() >= d) xi += t;
    }
    ws[lane_id()] = xi - x;
    if (lane_id() == 31) ws[32] = xi;
  }
  __syncthreads();
  unsigned long long carry = inc - s + ws[w] + init;
  for (uint64_t i = lo; i < hi; ++i) {
    uint32_t v = in[i];
    out[i] = carry;
    carry += v;
  }
  if (threadIdx.x == 0) out[n] = ws[32] + init;
}

// Slice summaries for a sharded plan: out[0] = max of head[0..n) (-1: none), out[1] = sum of cnt[0..n).  One CTA.
__global__ void __launch_bounds__(SC_NT) k1_slice_summary(const long long* __restrict__ head, const uint32_t* __restrict__ cnt,
                                                          uint64_t n, unsigned long long* __restrict__ out) {
  __shared__ long long wm[SC_NT / 32];
  __shared__ unsigned long long wsum[SC_NT / 32];
  long long m = -1;
  unsigned long long s = 0;
  for (uint64_t i = threadIdx.x; i < n; i += SC_NT) {
    if (head) m = max(m, head[i]);
    if (cnt) s += cnt[i];
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
    s += __shfl_xor_sync(0xffffffffu, s, d);
  }
  if (lane_id() == 0) { wm[threadIdx.x >> 5] = m; wsum[threadIdx.x >> 5] = s; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < SC_NT / 32; ++w) { m = max(m, wm[w]); s += wsum[w]; }
    if (head) out[0] = (unsigned long long)m;
    if (cnt) out[1] = s;
  }
}

// Evaluate one tile with the whole CTA: per-thread ThreadEval + exclusive emitted offset inside the tile.
__device__ __forceinline__ uint32_t tile_eval_cta(const uint8_t* __restrict__ in, uint64_t N, uint64_t tile,
                                                  long long tile_carry, ThreadBytes& tb, ThreadEval& ev,
                                                  uint32_t* q_out, long long* ws64, uint32_t* ws32, uint32_t* total) {
  load_thread_bytes(in, N, tile * K1_TILE, tb);
  long long lh = thread_last_head(tb);
  long long S = cta_excl_scan_max64<K1_NT>(lh, ws64);
  S = max(S, tile_carry);
  eval_thread(tb, N, S, ev, q_out);
  return cta_excl_scan_add<K1_NT>(ev.count, ws32, total);
}

// ---- kernel C: emitted bytes per tile ----
__global__ void __launch_bounds__(K1_NT) k1_tile_counts(const uint8_t* __restrict__ in, uint64_t N, uint64_t tile0,
                                                        const long long* __restrict__ tile_carry,
                                                        uint32_t* __restrict__ tile_cnt) {
  __shared__ long long ws64[K1_NT / 32];
  __shared__ uint32_t ws32[K1_NT / 32 + 1];
  ThreadBytes tb;
  ThreadEval ev;
  uint32_t total;
  const uint64_t tile = tile0 + blockIdx.x;
  tile_eval_cta(in, N, tile, tile_carry[tile], tb, ev, nullptr, ws64, ws32, &total);
  if (threadIdx.x == 0) tile_cnt[tile] = total;
}

// ---- kernel F: scatter the RLE1 byte stream ----
// A thread's bytes go to shared memory at the offset they have inside the tile's output (skewed so that shared index and
// global address agree modulo 16); the CTA then copies the tile's output range with 128-bit stores — byte stores only
// for the unaligned ends.  (A tile emits at most 5 bytes per 4 input bytes: 4 literals + a count.)
__global__ void __launch_bounds__(K1_NT) k1_scatter(const uint8_t* __restrict__ in, uint64_t N,
                                                    const long long* __restrict__ tile_carry,
                                                    const uint64_t* __restrict__ tile_E, uint64_t tile0,
                                                    uint8_t* __restrict__ out) {
  __shared__ long long ws64[K1_NT / 32];
  __shared__ uint32_t ws32[K1_NT / 32 + 1];
  __shared__ __align__(16) uint8_t sbuf[K1_TILE + K1_TILE / 4 + 48];
  ThreadBytes tb;
  ThreadEval ev;
  uint32_t q[K1_BPT];
  uint32_t total;
  const uint64_t tile = tile0 + blockIdx.x;
  const uint32_t ex = tile_eval_cta(in, N, tile, tile_carry[tile], tb, ev, q, ws64, ws32, &total);
  uint8_t* g = out + tile_E[tile];
  const uint32_t skew = (uint32_t)(reinterpret_cast<uintptr_t>(g) & 15u);
  uint32_t o = skew + ex;
#pragma unroll
  for (int j = 0; j < K1_BPT; ++j) {
    if (j < tb.cnt) {
      if (ev.lit_mask & (1u << j)) sbuf[o++] = tb.b[j];
      if (ev.cb_mask & (1u << j)) sbuf[o++] = (uint8_t)(q[j] - 3u);
    }
  }
  __syncthreads();
  const uint32_t head = min(total, (16u - skew) & 15u);  // bytes up to the first 16-byte boundary
  if (threadIdx.x < head) g[threadIdx.x] = sbuf[skew + threadIdx.x];
  const uint32_t nvec = (total - head) >> 4;
  const uint4* sv = reinterpret_cast<const uint4*>(sbuf + skew + head);  // skew + head is 0 or 16
  uint4* gv = reinterpret_cast<uint4*>(g + head);
  for (uint32_t v = threadIdx.x; v < nvec; v += K1_NT) gv[v] = sv[v];
  const uint32_t done = head + (nvec << 4);
  if (threadIdx.x < total - done) g[done + threadIdx.x] = sbuf[skew + done + threadIdx.x];
}

// ---- kernel E': the cut chain as parallel windows + a pointer walk ----
// A block that starts at emitted offset S ends at f(S + T), f(x) = emitted offset of the first piece end at or
// after x; f(x) - x <= 4.  With center_j = x0 + (j+1) T and drift d_j = S_j - (x0 + j T) the chain is
// d_{j+1} = f(center_j + d_j) - center_j.  k1_cut_windows tabulates F[j][w] = f(center_j + w) - center_j for
// w < CW_W for every j at once (one CTA per window: locate the tile, evaluate it, every piece end fills the <= 5
// offsets it covers); k1_cut_walk then follows the chain through the table, 32 windows per step as long as the drift
// does not change.  A drift that leaves the window ends the phase; the host starts the next one at that cut.
constexpr int CW_W = 256;
constexpr uint64_t CW_LAST = 1ull << 63;  // the piece is the last piece of the input: no cut (encoder.rs:729-739)

// One window: CTA-wide.  The first tile t with tile_E[t+1] >= center is searched in [t_lo, t_hi] (the whole input, or —
// sharded plan — the tiles of the slice the center lies in); tiles are evaluated up to t_limit (exclusive), which must
// reach the tile where the emitted offset passes center + CW_W + 4 (the slice's halo guarantees it).
__device__ __forceinline__ void cut_window_body(const uint8_t* __restrict__ in, uint64_t N,
                                                const long long* __restrict__ tile_carry,
                                                const uint64_t* __restrict__ tile_E, uint64_t t_lo, uint64_t t_hi,
                                                uint64_t t_limit, uint64_t center, uint64_t* __restrict__ Fj) {
  __shared__ long long ws64[K1_NT / 32];
  __shared__ uint32_t ws32[K1_NT / 32 + 1];
  __shared__ uint64_t s_lo, s_hi;
  __shared__ uint32_t s_first;
  // first tile t with tile_E[t+1] >= center (256-ary search)
  if (threadIdx.x == 0) { s_lo = t_lo; s_hi = t_hi; }
  __syncthreads();
  while (true) {
    const uint64_t lo = s_lo, hi = s_hi;
    if (lo >= hi) break;
    const uint64_t span = hi - lo + 1;
    const uint64_t step = (span + K1_NT - 1) / K1_NT;
    const uint64_t p = min(lo + (uint64_t)threadIdx.x * step, hi);
    const bool ok = tile_E[p + 1] >= center;
    if (threadIdx.x == 0) s_first = K1_NT;
    __syncthreads();
    if (ok) atomicMin(&s_first, (uint32_t)threadIdx.x);
    __syncthreads();
    const uint32_t f = s_first;
    __syncthreads();
    if (threadIdx.x == 0) {
      s_hi = min(lo + (uint64_t)f * step, hi);
      s_lo = f > 0 ? min(lo + (uint64_t)(f - 1) * step, hi) + 1 : lo;
    }
    __syncthreads();
  }
  for (uint64_t t = s_lo; t < t_limit; ++t) {
    ThreadBytes tb;
    ThreadEval ev;
    uint32_t q[K1_BPT];
    const uint32_t ex = tile_eval_cta(in, N, t, tile_carry[t], tb, ev, q, ws64, ws32, nullptr);
    uint64_t E = tile_E[t] + ex;
#pragma unroll
    for (int j = 0; j < K1_BPT; ++j) {
      if (j < tb.cnt) {
        E += ((ev.lit_mask >> j) & 1u) + ((ev.cb_mask >> j) & 1u);
        const uint64_t i = tb.i0 + j;
        const uint8_t nb = (j + 1 < tb.cnt) ? tb.b[j + 1] : tb.next;
        const bool pe = (i == N - 1) || (nb != tb.b[j]) || q[j] == 254u;
        if (pe && E >= center) {
          const uint32_t piece = min(q[j] + 1u, 4u) + (q[j] >= 3u ? 1u : 0u);  // bytes this piece emits
          const uint64_t eprev = E - piece;                                    // previous piece end
          if (eprev + 1 < center + CW_W) {
            const uint64_t wlo = eprev + 1 > center ? eprev + 1 - center : 0;
            const uint64_t whi = min((uint64_t)CW_W - 1, E - center);
            const uint64_t v = (E - center) | ((i + 1) << 16) | (i == N - 1 ? CW_LAST : 0ull);
            for (uint64_t w = wlo; w <= whi; ++w) Fj[w] = v;
          }
        }
      }
    }
    if (tile_E[t + 1] >= center + CW_W + 4) break;  // every piece that covers an offset of the window ends by here
    __syncthreads();
  }
}

__global__ void __launch_bounds__(K1_NT) k1_cut_windows(const uint8_t* __restrict__ in, uint64_t N,
                                                        const long long* __restrict__ tile_carry,
                                                        const uint64_t* __restrict__ tile_E, uint64_t ntiles, uint32_t T,
                                                        const uint64_t* __restrict__ state, uint64_t* __restrict__ F) {
  if (state[2]) return;  // chain already finished
  const uint64_t x0 = state[1];
  const uint64_t Etot = tile_E[ntiles];
  const uint64_t center = x0 + (uint64_t)(blockIdx.x + 1) * T;
  if (center > Etot) return;
  cut_window_body(in, N, tile_carry, tile_E, 0, ntiles - 1, ntiles, center, F + (uint64_t)blockIdx.x * CW_W);
}

// Sharded plan: windows j0 .. j0 + gridDim.x - 1 of a phase that starts at emitted offset x0, tabulated by the rank whose
// slice [t_lo, t_hi] holds the tile the center falls into; in / tile_carry / tile_E are addressed by GLOBAL byte and
// tile indices (the caller passes pointers shifted by the slice origin), valid for tiles [t_lo, t_limit).
__global__ void __launch_bounds__(K1_NT) k1_cut_windows_slice(const uint8_t* __restrict__ in, uint64_t N,
                                                              const long long* __restrict__ tile_carry,
                                                              const uint64_t* __restrict__ tile_E, uint64_t t_lo,
                                                              uint64_t t_hi, uint64_t t_limit, uint64_t Etot, uint32_t T,
                                                              uint64_t x0, uint64_t j0, uint64_t* __restrict__ F) {
  const uint64_t center = x0 + (j0 + blockIdx.x + 1) * (uint64_t)T;
  if (center > Etot) return;
  cut_window_body(in, N, tile_carry, tile_E, t_lo, t_hi, t_limit, center, F + (uint64_t)blockIdx.x * CW_W);
}

// state[0] = blocks cut so far (k), [1] = emitted offset where the current block starts (S), [2] = done,
// [3] = longest block so far.  One warp.
__global__ void __launch_bounds__(32) k1_cut_walk(const uint64_t* __restrict__ F, uint32_t K, uint32_t T,
                                                  const uint64_t* __restrict__ tile_E, uint64_t ntiles, uint64_t N,
                                                  uint64_t* __restrict__ in_off, uint64_t* __restrict__ rle_off,
                                                  uint32_t max_blocks, uint64_t* __restrict__ state,
                                                  uint32_t* __restrict__ nblocks_out, uint32_t* __restrict__ max_block_len) {
  if (state[2]) return;
  const uint32_t lane = lane_id();
  const uint64_t Etot = tile_E[ntiles];
  const uint64_t x0 = state[1];
  uint64_t k = state[0], S = x0, maxlen = state[3];
  uint64_t d = 0;
  uint32_t j = 0;
  bool done = false;
  if (k == 0 && lane == 0) { in_off[0] = 0; rle_off[0] = 0; }
  while (true) {
    const uint32_t jj = j + lane;
    const uint64_t center = x0 + (uint64_t)(jj + 1) * T;
    const bool in_table = jj < K;
    const bool reach = center + d <= Etot && k + lane + 2 <= max_blocks;  // otherwise the rest is the last block
    const uint64_t v = (in_table && reach) ? F[(uint64_t)jj * CW_W + d] : 0ull;
    const uint64_t rel = v & 0xFFFFull;
    const bool good = in_table && reach && !(v & CW_LAST) && rel == d;
    const uint32_t bad = __ballot_sync(0xffffffffu, !good);
    const uint32_t nb_good = bad ? (uint32_t)(__ffs(bad) - 1) : 32u;
    if (lane < nb_good) {  // cuts with unchanged drift: every block is exactly T long
      in_off[k + 1 + lane] = (v >> 16) & 0x7FFFFFFFFFFFull;
      rle_off[k + 1 + lane] = center + rel;
    }
    if (nb_good) {
      maxlen = max(maxlen, (uint64_t)T);
      k += nb_good;
      j += nb_good;
      S = x0 + (uint64_t)j * T + d;
    }
    if (nb_good == 32) continue;
    // the first lane that breaks the pattern decides what happens next (all lanes follow it)
    const uint32_t src = nb_good;
    const uint64_t bv = __shfl_sync(0xffffffffu, v, src);
    const bool b_in = __shfl_sync(0xffffffffu, (int)in_table, src);
    const bool b_reach = __shfl_sync(0xffffffffu, (int)reach, src);
    const uint64_t b_center = x0 + (uint64_t)(j + 1) * T;
    if (!b_in) break;                                   // table exhausted: the host starts another phase at S
    if (!b_reach || (bv & CW_LAST)) { done = true; break; }
    const uint64_t brel = bv & 0xFFFFull;               // a cut that changes the drift
    k += 1;
    if (lane == 0) {
      in_off[k] = (bv >> 16) & 0x7FFFFFFFFFFFull;
      rle_off[k] = b_center + brel;
    }
    maxlen = max(maxlen, b_center + brel - S);
    S = b_center + brel;
    j += 1;
    if (brel >= (uint64_t)CW_W) break;                  // drift left the window: next phase starts at S
    d = brel;
  }
  if (lane == 0) {
    state[0] = k;
    state[1] = S;
    state[3] = maxlen;
    if (done) {
      state[2] = 1;
      const uint32_t nb = (uint32_t)k + 1;
      in_off[nb] = N;
      rle_off[nb] = Etot;
      maxlen = max(maxlen, Etot - S);
      *nblocks_out = nb;
      *max_block_len = (uint32_t)min((uint64_t)0xFFFFFFFFu, maxlen);
    }
  }
}

// ---- K5: CRC-32/BZIP2 per block over its input range ----
__device__ __forceinline__ uint32_t gf2_mulmod(uint32_t a, uint32_t b) {
  // a*b mod P, P = x^32 + 0x04C11DB7, bit 31 = x^31
  uint32_t r = 0;
#pragma unroll 4
  for (int i = 31; i >= 0; --i) {
    r = (r << 1) ^ ((r & 0x80000000u) ? 0x04C11DB7u : 0u);
    if ((b >> i) & 1u) r ^= a;
  }
  return r;
}
// x^(8*L) mod P
__device__ __forceinline__ uint32_t gf2_xpow8(uint64_t L) {
  uint32_t result = 1u;       // x^0
  uint32_t base = 0x100u;     // x^8
  while (L) {
    if (L & 1ull) result = gf2_mulmod(result, base);
    base = gf2_mulmod(base, base);
    L >>= 1;
  }
  return result;
}

constexpr int CRC_NT = 1024;
__global__ void __launch_bounds__(CRC_NT) k5_crc_blocks(const uint8_t* __restrict__ in,
                                                        const uint64_t* __restrict__ in_off,
                                                        uint32_t* __restrict__ crc_out) {
  // slicing-by-4: tab[k][i] = register after byte i followed by k zero bytes, so four bytes cost four INDEPENDENT lookups
  // instead of a chain of four (MSB-first CRC-32/BZIP2, crc32.rs:82-84)
  __shared__ uint32_t tab[4][256];
  __shared__ uint32_t red[CRC_NT / 32];
  for (int i = threadIdx.x; i < 256; i += CRC_NT) {
    uint32_t v = (uint32_t)i << 24;
#pragma unroll
    for (int k = 0; k < 8; ++k) v = (v & 0x80000000u) ? (v << 1) ^ 0x04C11DB7u : (v << 1);
    tab[0][i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += CRC_NT) {
    uint32_t v = tab[0][i];
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      v = (v << 8) ^ tab[0][v >> 24];
      tab[k][i] = v;
    }
  }
  __syncthreads();
  const uint64_t lo = in_off[blockIdx.x], hi = in_off[blockIdx.x + 1];
  const uint64_t L = hi - lo;
  const uint64_t per = (L + CRC_NT - 1) / CRC_NT;
  const uint64_t a = min(L, per * threadIdx.x), b = min(L, a + per);
  uint32_t r = 0;
  const uint8_t* p = in + lo;
  uint64_t i = a;
  for (; i < b && ((reinterpret_cast<uintptr_t>(p + i)) & 3u); ++i) r = tab[0][((r >> 24) ^ __ldg(p + i)) & 0xFF] ^ (r << 8);
  for (; i + 4 <= b; i += 4) {
    const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p + i));
    const uint32_t x = r ^ __byte_perm(w, 0u, 0x0123);  // the first byte of the four on top
    r = tab[3][x >> 24] ^ tab[2][(x >> 16) & 0xFF] ^ tab[1][(x >> 8) & 0xFF] ^ tab[0][x & 0xFF];
  }
  for (; i < b; ++i) r = tab[0][((r >> 24) ^ __ldg(p + i)) & 0xFF] ^ (r << 8);
  // contribution of this span to the register at the end of the block: raw * x^(8*(L-b))
  uint32_t contrib = (b > a) ? gf2_mulmod(r, gf2_xpow8(L - b)) : 0u;
  if (threadIdx.x == 0) contrib ^= gf2_mulmod(0xFFFFFFFFu, gf2_xpow8(L));  // initial register value
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) contrib ^= __shfl_xor_sync(0xffffffffu, contrib, d);
  if (lane_id() == 0) red[threadIdx.x >> 5] = contrib;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t x = 0;
    for (int w = 0; w < CRC_NT / 32; ++w) x ^= red[w];
    crc_out[blockIdx.x] = ~x;
  }
}

// ---- per-block in-use map over the RLE1 bytes (EncoderInner::in_use, encoder.rs:707,713) ----
__global__ void __launch_bounds__(256) k1_inuse(const uint8_t* __restrict__ txt, const uint64_t* __restrict__ rle_off,
                                                uint32_t* __restrict__ inuse /*[nb][8]*/) {
  // one flag byte per byte value and warp (plain stores: setting a flag twice is harmless), folded into the 256-bit map
  // with ballots at the end — a register bitmap needs an 8-way select per input byte
  __shared__ uint8_t flag[8][256];
  __shared__ uint32_t m[8];
  for (int i = threadIdx.x; i < 8 * 256 / 4; i += 256) reinterpret_cast<uint32_t*>(&flag[0][0])[i] = 0;
  if (threadIdx.x < 8) m[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t lo = rle_off[blockIdx.x], hi = rle_off[blockIdx.x + 1];
  uint8_t* fw = flag[threadIdx.x >> 5];
  auto mark = [&](uint32_t c) { fw[c] = 1; };
  // unaligned head and tail byte-wise, the body as 128-bit loads
  const uint64_t alo = min(hi, (uint64_t)((lo + 15) & ~15ull)), ahi = max(alo, (uint64_t)(hi & ~15ull));
  for (uint64_t i = lo + threadIdx.x; i < alo; i += blockDim.x) mark(__ldg(txt + i));
  for (uint64_t i = alo + (uint64_t)threadIdx.x * 16; i < ahi; i += (uint64_t)blockDim.x * 16) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(txt + i));
    const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      mark(wv[k] & 255u);
      mark((wv[k] >> 8) & 255u);
      mark((wv[k] >> 16) & 255u);
      mark(wv[k] >> 24);
    }
  }
  for (uint64_t i = ahi + threadIdx.x; i < hi; i += blockDim.x) mark(__ldg(txt + i));
  __syncthreads();
  {  // thread c folds the eight warps' flags of byte value c; a ballot per warp gives 32 bits of the map
    uint32_t any = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) any |= flag[w][threadIdx.x];
    const uint32_t bits = __ballot_sync(0xffffffffu, any != 0);
    if (lane_id() == 0) m[threadIdx.x >> 5] = bits;
  }
  __syncthreads();
  if (threadIdx.x < 8) inuse[blockIdx.x * 8 + threadIdx.x] = m[threadIdx.x];
}

// =========================== host launchers ===========================
uint64_t k1_num_tiles(uint64_t N) { return (N + K1_TILE - 1) / K1_TILE; }

// K1 in four steps so that a sharded caller can compute the per-tile summaries of its own tile range and exchange
// them between ranks: heads [t0,t1) -> (exchange) -> carry scan + counts [t0,t1) -> (exchange) -> prefix sum.
uint32_t k1_tile_bytes() { return K1_TILE; }
void launch_k1_heads(Launcher& L, const uint8_t* d_in, uint64_t N, uint64_t t0, uint64_t t1, long long* d_tile_head) {
  if (t1 > t0)
    L.launch("k1_tile_heads", k1_tile_heads, dim3((unsigned)(t1 - t0)), dim3(K1_NT), d_in, N, t0, d_tile_head);
}
void launch_k1_counts(Launcher& L, const uint8_t* d_in, uint64_t N, uint64_t t0, uint64_t t1,
                      const long long* d_tile_head, long long* d_tile_carry, uint32_t* d_tile_cnt) {
  uint64_t nt = k1_num_tiles(N);
  L.launch("k_scan_max64_excl", k_scan_max64_excl, dim3(1), dim3(SC_NT), d_tile_head, d_tile_carry, nt, -1ll);
  if (t1 > t0)
    L.launch("k1_tile_counts", k1_tile_counts, dim3((unsigned)(t1 - t0)), dim3(K1_NT), d_in, N, t0,
             (const long long*)d_tile_carry, d_tile_cnt);
}
void launch_k1_prefix(Launcher& L, uint64_t N, const uint32_t* d_tile_cnt, uint64_t* d_tile_E) {
  uint64_t nt = k1_num_tiles(N);
  L.launch("k_scan_add64_excl", k_scan_add64_excl, dim3(1), dim3(SC_NT), d_tile_cnt, d_tile_E, nt, 0ull);
}

uint32_t k1_cut_window() { return CW_W; }

// ---- sharded plan (one slice of the input per context; see pipeline.cu "sliced plan") ----
// All pointers are shifted so that GLOBAL byte / tile indices address them; only tiles [t_a, t_b) are touched.
void launch_k1_slice_heads(Launcher& L, const uint8_t* v_in, uint64_t N, uint64_t t_a, uint64_t t_b, long long* v_tile_head) {
  if (t_b > t_a)
    L.launch("k1_tile_heads", k1_tile_heads, dim3((unsigned)(t_b - t_a)), dim3(K1_NT), v_in, N, t_a, v_tile_head);
}
void launch_k1_slice_summary(Launcher& L, const long long* head, const uint32_t* cnt, uint64_t n, uint64_t* d_out2) {
  L.launch("k1_slice_summary", k1_slice_summary, dim3(1), dim3(SC_NT), head, cnt, n,
           reinterpret_cast<unsigned long long*>(d_out2));
}
void launch_k1_slice_counts(Launcher& L, const uint8_t* v_in, uint64_t N, uint64_t t_a, uint64_t t_b, long long carry_in,
                            const long long* v_tile_head, long long* v_tile_carry, uint32_t* v_tile_cnt) {
  if (t_b <= t_a) return;
  L.launch("k_scan_max64_excl", k_scan_max64_excl, dim3(1), dim3(SC_NT), v_tile_head + t_a, v_tile_carry + t_a, t_b - t_a,
           carry_in);
  L.launch("k1_tile_counts", k1_tile_counts, dim3((unsigned)(t_b - t_a)), dim3(K1_NT), v_in, N, t_a,
           (const long long*)v_tile_carry, v_tile_cnt);
}
void launch_k1_slice_scan_carry(Launcher& L, uint64_t t_a, uint64_t t_b, long long carry_in, const long long* v_tile_head,
                                long long* v_tile_carry) {
  if (t_b > t_a)
    L.launch("k_scan_max64_excl", k_scan_max64_excl, dim3(1), dim3(SC_NT), v_tile_head + t_a, v_tile_carry + t_a, t_b - t_a,
             carry_in);
}
void launch_k1_slice_tile_counts(Launcher& L, const uint8_t* v_in, uint64_t N, uint64_t t_a, uint64_t t_b,
                                 const long long* v_tile_carry, uint32_t* v_tile_cnt) {
  if (t_b > t_a)
    L.launch("k1_tile_counts", k1_tile_counts, dim3((unsigned)(t_b - t_a)), dim3(K1_NT), v_in, N, t_a, v_tile_carry,
             v_tile_cnt);
}
void launch_k1_slice_prefix(Launcher& L, uint64_t t_a, uint64_t t_b, uint64_t E_in, const uint32_t* v_tile_cnt,
                            uint64_t* v_tile_E) {
  if (t_b <= t_a) return;
  L.launch("k_scan_add64_excl", k_scan_add64_excl, dim3(1), dim3(SC_NT), v_tile_cnt + t_a, v_tile_E + t_a, t_b - t_a,
           (unsigned long long)E_in);
}
void launch_k1_slice_windows(Launcher& L, const uint8_t* v_in, uint64_t N, uint32_t T, const long long* v_tile_carry,
                             const uint64_t* v_tile_E, uint64_t t_lo, uint64_t t_hi, uint64_t t_limit, uint64_t Etot,
                             uint64_t x0, uint64_t j0, uint32_t nj, uint64_t* d_F) {
  if (nj)
    L.launch("k1_cut_windows", k1_cut_windows_slice, dim3(nj), dim3(K1_NT), v_in, N, v_tile_carry, v_tile_E, t_lo, t_hi,
             t_limit, Etot, T, x0, j0, d_F);
}

// One phase of the cut chain: K windows from the chain state in d_state, then the walk.
void launch_k1_cut_phase(Launcher& L, const uint8_t* d_in, uint64_t N, uint32_t T, const long long* d_tile_carry,
                         const uint64_t* d_tile_E, uint32_t K, uint64_t* d_F, uint64_t* d_state, uint64_t* d_in_off,
                         uint64_t* d_rle_off, uint32_t max_blocks, uint32_t* d_nblocks, uint32_t* d_maxlen) {
  uint64_t nt = k1_num_tiles(N);
  if (K)
    L.launch("k1_cut_windows", k1_cut_windows, dim3(K), dim3(K1_NT), d_in, N, d_tile_carry, d_tile_E, nt, T,
             (const uint64_t*)d_state, d_F);
  L.launch("k1_cut_walk", k1_cut_walk, dim3(1), dim3(32), (const uint64_t*)d_F, K, T, d_tile_E, nt, N, d_in_off,
           d_rle_off, max_blocks, d_state, d_nblocks, d_maxlen);
}

// RLE1 bytes of the input range [in_lo, in_hi) (whole tiles; the bytes land at their global emitted offsets).
void launch_k1_scatter(Launcher& L, const uint8_t* d_in, uint64_t N, uint64_t in_lo, uint64_t in_hi,
                       const long long* d_tile_carry, const uint64_t* d_tile_E, uint8_t* d_txt) {
  if (in_hi <= in_lo) return;
  const uint64_t t0 = in_lo / K1_TILE, t1 = (in_hi + K1_TILE - 1) / K1_TILE;
  L.launch("k1_scatter", k1_scatter, dim3((unsigned)(t1 - t0)), dim3(K1_NT), d_in, N, d_tile_carry, d_tile_E, t0,
           d_txt);
}

void launch_k5_crc(Launcher& L, const uint8_t* d_in, const uint64_t* d_in_off, uint32_t nblocks, uint32_t* d_crc) {
  L.launch("k5_crc_blocks", k5_crc_blocks, dim3(nblocks), dim3(CRC_NT), d_in, d_in_off, d_crc);
}

void launch_k1_inuse(Launcher& L, const uint8_t* d_txt, const uint64_t* d_rle_off, uint32_t nblocks, uint32_t* d_inuse) {
  L.launch("k1_inuse", k1_inuse, dim3(nblocks), dim3(256), d_txt, d_rle_off, d_inuse);
}

}  // namespace bzb
