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
  if (cnt == K1_BPT && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
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
__global__ void __launch_bounds__(K1_NT) k1_tile_heads(const uint8_t* __restrict__ in, uint64_t N,
                                                       long long* __restrict__ tile_last_head) {
  __shared__ long long ws[K1_NT / 32];
  ThreadBytes tb;
  load_thread_bytes(in, N, (uint64_t)blockIdx.x * K1_TILE, tb);
  long long lh = thread_last_head(tb);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) lh = max(lh, __shfl_xor_sync(0xffffffffu, lh, d));
  if (lane_id() == 0) ws[threadIdx.x >> 5] = lh;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long m = -1;
    for (int w = 0; w < K1_NT / 32; ++w) m = max(m, ws[w]);
    tile_last_head[blockIdx.x] = m;
  }
}

// ---- generic single-CTA scans over per-tile arrays ----
constexpr int SC_NT = 1024;
__global__ void __launch_bounds__(SC_NT) k_scan_max64_excl(const long long* __restrict__ in, long long* __restrict__ out,
                                                           uint64_t n) {
  __shared__ long long ws[SC_NT / 32];
  uint64_t per = (n + SC_NT - 1) / SC_NT;
  uint64_t lo = min(n, per * threadIdx.x), hi = min(n, lo + per);
  long long m = -1;
  for (uint64_t i = lo; i < hi; ++i) m = max(m, in[i]);
  long long carry = cta_excl_scan_max64<SC_NT>(m, ws);
  for (uint64_t i = lo; i < hi; ++i) {
    long long v = in[i];
    out[i] = carry;
    carry = max(carry, v);
  }
}

// out has n+1 entries: exclusive prefix, out[n] = total.
__global__ void __launch_bounds__(SC_NT) k_scan_add64_excl(const uint32_t* __restrict__ in, uint64_t* __restrict__ out,
                                                           uint64_t n) {
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
  unsigned long long carry = inc - s + ws[w];
  for (uint64_t i = lo; i < hi; ++i) {
    uint32_t v = in[i];
    out[i] = carry;
    carry += v;
  }
  if (threadIdx.x == 0) out[n] = ws[32];
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
__global__ void __launch_bounds__(K1_NT) k1_tile_counts(const uint8_t* __restrict__ in, uint64_t N,
                                                        const long long* __restrict__ tile_carry,
                                                        uint32_t* __restrict__ tile_cnt) {
  __shared__ long long ws64[K1_NT / 32];
  __shared__ uint32_t ws32[K1_NT / 32 + 1];
  ThreadBytes tb;
  ThreadEval ev;
  uint32_t total;
  tile_eval_cta(in, N, blockIdx.x, tile_carry[blockIdx.x], tb, ev, nullptr, ws64, ws32, &total);
  if (threadIdx.x == 0) tile_cnt[blockIdx.x] = total;
}

// ---- kernel F: scatter the RLE1 byte stream ----
__global__ void __launch_bounds__(K1_NT) k1_scatter(const uint8_t* __restrict__ in, uint64_t N,
                                                    const long long* __restrict__ tile_carry,
                                                    const uint64_t* __restrict__ tile_E, uint8_t* __restrict__ out) {
  __shared__ long long ws64[K1_NT / 32];
  __shared__ uint32_t ws32[K1_NT / 32 + 1];
  ThreadBytes tb;
  ThreadEval ev;
  uint32_t q[K1_BPT];
  uint32_t ex = tile_eval_cta(in, N, blockIdx.x, tile_carry[blockIdx.x], tb, ev, q, ws64, ws32, nullptr);
  uint64_t o = tile_E[blockIdx.x] + ex;
#pragma unroll
  for (int j = 0; j < K1_BPT; ++j) {
    if (j < tb.cnt) {
      if (ev.lit_mask & (1u << j)) out[o++] = tb.b[j];
      if (ev.cb_mask & (1u << j)) out[o++] = (uint8_t)(q[j] - 3u);
    }
  }
}

// ---- kernel E: greedy cut chain (one CTA) ----
// cuts: in_off[k], rle_off[k] for k=0..nblocks; nblocks_out. max_blocks bounds the arrays (nblocks+1 <= max_blocks).
__global__ void __launch_bounds__(K1_NT) k1_cut_chain(const uint8_t* __restrict__ in, uint64_t N,
                                                      const long long* __restrict__ tile_carry,
                                                      const uint64_t* __restrict__ tile_E, uint64_t ntiles, uint32_t T,
                                                      uint64_t* __restrict__ in_off, uint64_t* __restrict__ rle_off,
                                                      uint32_t max_blocks, uint32_t* __restrict__ nblocks_out,
                                                      uint32_t* __restrict__ max_block_len) {
  __shared__ long long ws64[K1_NT / 32];
  __shared__ uint32_t ws32[K1_NT / 32 + 1];
  __shared__ uint64_t s_lo, s_hi, s_S, s_cut_i, s_cut_E;
  __shared__ uint32_t s_first, s_k, s_done, s_maxlen;
  const uint64_t Etot = tile_E[ntiles];
  if (threadIdx.x == 0) {
    s_S = 0; s_k = 0; s_done = 0; s_lo = 0; s_maxlen = 0;
    in_off[0] = 0; rle_off[0] = 0;
  }
  __syncthreads();
  while (true) {
    const uint64_t S = s_S;
    const uint64_t target = S + T;
    if (target > Etot || s_k + 2 > max_blocks) break;  // remaining bytes fit in the last block
    // -- 256-ary search: smallest tile t in [lo, ntiles-1] with tile_E[t+1] >= target
    if (threadIdx.x == 0) s_hi = ntiles - 1;
    __syncthreads();
    while (true) {
      uint64_t lo = s_lo, hi = s_hi;
      if (lo >= hi) break;
      uint64_t span = hi - lo + 1;
      uint64_t step = (span + K1_NT - 1) / K1_NT;
      uint64_t p = min(lo + (uint64_t)threadIdx.x * step, hi);
      bool ok = tile_E[p + 1] >= target;
      if (threadIdx.x == 0) s_first = K1_NT;
      __syncthreads();
      if (ok) atomicMin(&s_first, (uint32_t)threadIdx.x);
      __syncthreads();
      uint32_t f = s_first;
      __syncthreads();
      if (threadIdx.x == 0) {
        uint64_t pf = min(lo + (uint64_t)f * step, hi);
        uint64_t pl = f > 0 ? min(lo + (uint64_t)(f - 1) * step, hi) + 1 : lo;
        s_hi = pf;
        s_lo = pl;
      }
      __syncthreads();
    }
    const uint64_t t = s_lo;
    // -- evaluate tile t, find first byte with E(i) >= target
    ThreadBytes tb;
    ThreadEval ev;
    uint32_t q[K1_BPT];
    uint32_t ex = tile_eval_cta(in, N, t, tile_carry[t], tb, ev, q, ws64, ws32, nullptr);
    uint64_t E = tile_E[t] + ex;
    int cand = -1;
    uint64_t candE = 0;
#pragma unroll
    for (int j = 0; j < K1_BPT; ++j) {
      if (j < tb.cnt) {
        E += ((ev.lit_mask >> j) & 1u) + ((ev.cb_mask >> j) & 1u);
        if (cand < 0 && E >= target) { cand = j; candE = E; }
      }
    }
    if (threadIdx.x == 0) s_first = 0xFFFFFFFFu;
    __syncthreads();
    if (cand >= 0) atomicMin(&s_first, (uint32_t)(threadIdx.x * K1_BPT + cand));
    __syncthreads();
    if (cand >= 0 && s_first == (uint32_t)(threadIdx.x * K1_BPT + cand)) {
      // walk to the end of the piece that contains byte i
      uint64_t i = tb.i0 + cand;
      uint32_t qq = q[cand];
      uint8_t c = tb.b[cand];
      bool tail = (i == N - 1) || (((cand + 1 < tb.cnt) ? tb.b[cand + 1] : tb.next) != c);
      bool pe = tail || qq == 254u;
      uint64_t EE = candE;
      while (!pe) {
        ++i;
        ++qq;  // same run, qq <= 254
        tail = (i == N - 1) || (in[i + 1] != c);
        pe = tail || qq == 254u;
        EE += (qq < 4u ? 1u : 0u) + ((pe && qq >= 3u) ? 1u : 0u);
      }
      s_cut_i = i;
      s_cut_E = EE;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      if (s_cut_i == N - 1) {
        s_done = 1;  // the piece that reaches T is the last piece of the input: no cut (encoder.rs:729-739)
      } else {
        uint32_t k = s_k + 1;
        in_off[k] = s_cut_i + 1;
        rle_off[k] = s_cut_E;
        s_maxlen = max(s_maxlen, (uint32_t)(s_cut_E - S));
        s_k = k;
        s_S = s_cut_E;
        s_lo = t;  // next search starts at this tile (E is monotone)
      }
    }
    __syncthreads();
    if (s_done) break;
  }
  if (threadIdx.x == 0) {
    uint32_t nb = s_k + 1;
    in_off[nb] = N;
    rle_off[nb] = Etot;
    s_maxlen = max(s_maxlen, (uint32_t)min((uint64_t)0xFFFFFFFFu, Etot - s_S));
    *nblocks_out = nb;
    *max_block_len = s_maxlen;
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
  __shared__ uint32_t tab[256];
  __shared__ uint32_t red[CRC_NT / 32];
  for (int i = threadIdx.x; i < 256; i += CRC_NT) {
    uint32_t v = (uint32_t)i << 24;
#pragma unroll
    for (int k = 0; k < 8; ++k) v = (v & 0x80000000u) ? (v << 1) ^ 0x04C11DB7u : (v << 1);
    tab[i] = v;
  }
  __syncthreads();
  const uint64_t lo = in_off[blockIdx.x], hi = in_off[blockIdx.x + 1];
  const uint64_t L = hi - lo;
  const uint64_t per = (L + CRC_NT - 1) / CRC_NT;
  const uint64_t a = min(L, per * threadIdx.x), b = min(L, a + per);
  uint32_t r = 0;
  const uint8_t* p = in + lo;
  for (uint64_t i = a; i < b; ++i) r = tab[((r >> 24) ^ __ldg(p + i)) & 0xFF] ^ (r << 8);
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
  __shared__ uint32_t m[8];
  if (threadIdx.x < 8) m[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t lo = rle_off[blockIdx.x], hi = rle_off[blockIdx.x + 1];
  uint32_t loc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    uint8_t c = __ldg(txt + i);
#pragma unroll
    for (int w = 0; w < 8; ++w) loc[w] |= ((c >> 5) == w) ? (1u << (c & 31)) : 0u;
  }
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    uint32_t v = loc[w];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, d);
    if (lane_id() == 0 && v) atomicOr(&m[w], v);
  }
  __syncthreads();
  if (threadIdx.x < 8) inuse[blockIdx.x * 8 + threadIdx.x] = m[threadIdx.x];
}

// =========================== host launchers ===========================
uint64_t k1_num_tiles(uint64_t N) { return (N + K1_TILE - 1) / K1_TILE; }

void launch_k1_plan(Launcher& L, const uint8_t* d_in, uint64_t N, uint32_t T, long long* d_tile_head,
                    long long* d_tile_carry, uint32_t* d_tile_cnt, uint64_t* d_tile_E, uint64_t* d_in_off,
                    uint64_t* d_rle_off, uint32_t max_blocks, uint32_t* d_nblocks, uint32_t* d_maxlen) {
  uint64_t nt = k1_num_tiles(N);
  L.launch("k1_tile_heads", k1_tile_heads, dim3((unsigned)nt), dim3(K1_NT), d_in, N, d_tile_head);
  L.launch("k_scan_max64_excl", k_scan_max64_excl, dim3(1), dim3(SC_NT), (const long long*)d_tile_head,
           d_tile_carry, nt);
  L.launch("k1_tile_counts", k1_tile_counts, dim3((unsigned)nt), dim3(K1_NT), d_in, N,
           (const long long*)d_tile_carry, d_tile_cnt);
  L.launch("k_scan_add64_excl", k_scan_add64_excl, dim3(1), dim3(SC_NT), (const uint32_t*)d_tile_cnt,
           d_tile_E, nt);
  L.launch("k1_cut_chain", k1_cut_chain, dim3(1), dim3(K1_NT), d_in, N, (const long long*)d_tile_carry,
           (const uint64_t*)d_tile_E, nt, T, d_in_off, d_rle_off, max_blocks, d_nblocks, d_maxlen);
}

void launch_k1_scatter(Launcher& L, const uint8_t* d_in, uint64_t N, const long long* d_tile_carry,
                       const uint64_t* d_tile_E, uint8_t* d_txt) {
  uint64_t nt = k1_num_tiles(N);
  L.launch("k1_scatter", k1_scatter, dim3((unsigned)nt), dim3(K1_NT), d_in, N, d_tile_carry, d_tile_E,
           d_txt);
}

void launch_k5_crc(Launcher& L, const uint8_t* d_in, const uint64_t* d_in_off, uint32_t nblocks, uint32_t* d_crc) {
  L.launch("k5_crc_blocks", k5_crc_blocks, dim3(nblocks), dim3(CRC_NT), d_in, d_in_off, d_crc);
}

void launch_k1_inuse(Launcher& L, const uint8_t* d_txt, const uint64_t* d_rle_off, uint32_t nblocks, uint32_t* d_inuse) {
  L.launch("k1_inuse", k1_inuse, dim3(nblocks), dim3(256), d_txt, d_rle_off, d_inuse);
}

}  // namespace bzb
