// k3_mtf.cu — K3: move-to-front + RUNA/RUNB zero-run coding of the BWT last column.
//
// Replaces the MTF/ZLE loop of write_blockdata (src/bzip2/encoder.rs:318-358), MtfPosition::pop
// (src/bzip2/mtf.rs:22-38) and zle_write (encoder.rs:653-669) by result.
//
// Formulation (checked in tests/test_oracle.py::test_mtf_distinct_count_formulation):
//   MTF position of byte c at index i = #{ s : last[s] > last[c] }, last[s] = index of the previous occurrence
//   of s, or the virtual index -1-dense(s) (dense = rank of s among the in-use bytes) if there is none.
// A position is 0 exactly when L[i] == L[i-1] (or, at i = 0, L[0] is the smallest in-use byte), so the zero-run
// structure is known from L alone and positions are only evaluated at run heads.
// The block is cut into chunks of MTF_CHUNK bytes, one warp per chunk:
//   pass A  per chunk: last occurrence of every byte + zero-run summary (lead, trail, nonzeros, interior digits)
//   pass B  per block: running max over chunks -> state at each chunk start; output offsets of each chunk
//   pass C  one lane per chunk: the list at the chunk start is rebuilt from `last` (rank of every byte's previous
//           occurrence); its first 32 entries are packed into eight registers of the lane (SIMD byte compare +
//           byte permute), the rest sits in shared memory; symbols are written straight to the offsets pass B
//           assigned; symbol histogram in shared memory.
#include <limits.h>

#include "common.cuh"
#include "kernels.h"

namespace bzb {

// Bytes per chunk: 4 096 when the batch is large enough to fill the GPU with chunks of that size (a chunk costs 1 KB of
// state and a list rebuild: 4 096 is the fastest at 1 GiB), 2 048 / 1 024 for smaller batches — one GPU's share of a
// sharded stream, a single block — where k3_apply's duration is ONE lane's walk through its chunk.
constexpr int MTF_CHUNK_MAX = 4096;
constexpr int MTF_WARPS = 4;  // warps (= chunks) per CTA
constexpr int NEG_UNUSED = -(1 << 30);

uint32_t mtf_chunk_elems(uint64_t batch_bytes) {
  // warp tasks of k3_apply = batch / (32 lanes x chunk); ~5 300 warps are resident on a B200 (36 per SM)
  for (uint32_t c = MTF_CHUNK_MAX; c > 1024; c >>= 1)
    if (batch_bytes / (32ull * c) >= 4000) return c;
  return 1024;
}

__device__ __forceinline__ uint32_t zle_digits(uint32_t z) {  // number of RUNA/RUNB symbols for a run of z zeros
  return z ? (31u - __clz(z + 1u)) : 0u;
}

__device__ __forceinline__ uint8_t smallest_inuse(const uint32_t* __restrict__ iu) {
  for (int w = 0; w < 8; ++w) {
    uint32_t v = iu[w];
    if (v) return (uint8_t)(w * 32 + __ffs(v) - 1);
  }
  return 0;
}

// ---- pass A ----
// zle summary per chunk: x = lead zeros (== len if all zero), y = trail zeros, z = nonzeros + interior digits,
// w = 1 if the chunk has at least one nonzero.
template <int MTF_CHUNK>
__global__ void __launch_bounds__(MTF_WARPS * 32) k3_chunk_scan_a(const uint8_t* __restrict__ last,
                                                                   const BlockDesc* __restrict__ desc,
                                                                   const uint32_t* __restrict__ inuse,
                                                                   int* __restrict__ chunk_state,
                                                                   uint4* __restrict__ chunk_zle, uint32_t chunks_cap) {
  __shared__ int s_last[MTF_WARPS][256];
  const int w = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const BlockDesc d = desc[blockIdx.y];
  const uint32_t chunk = blockIdx.x * MTF_WARPS + w;
  const uint32_t c0 = chunk * MTF_CHUNK;
  if (c0 >= d.n) return;
  const uint32_t len = min((uint32_t)MTF_CHUNK, d.n - c0);
  const uint8_t* L = last + d.off;
  for (int i = lane; i < 256; i += 32) s_last[w][i] = INT_MIN;  // absent: below every virtual index
  __syncwarp();
  uint8_t prev = c0 > 0 ? L[c0 - 1] : smallest_inuse(inuse + blockIdx.y * 8);
  uint32_t lead = 0, zrun = 0, emitted = 0;
  bool seen = false;
  for (uint32_t i0 = 0; i0 < len; i0 += 32) {
    const uint32_t i = i0 + lane;
    const bool valid = i < len;
    const uint32_t c = valid ? L[c0 + i] : 0x100u + lane;
    uint32_t p = __shfl_up_sync(0xffffffffu, c, 1);
    if (lane == 0) p = prev;
    prev = (uint8_t)__shfl_sync(0xffffffffu, c, 31);
    // last occurrence inside the row: no higher lane holds the same byte; rows are visited in order
    const uint32_t peers = __match_any_sync(0xffffffffu, c);
    if (valid && (peers >> lane) == 1u) s_last[w][c] = (int)(c0 + i);
    __syncwarp();  // the next row may store to the same byte's slot from another lane: keep the stores ordered
    const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
    const uint32_t nz = __ballot_sync(0xffffffffu, valid && c != p);
    const bool is_nz = (nz >> lane) & 1u;
    const uint32_t below = nz & lanemask_lt();
    uint32_t z = below ? lane - (32u - __clz(below)) : zrun + lane;  // zeros in front of this run head
    uint32_t e = 0;
    if (is_nz) {
      if (!seen && !below) lead = z;  // first head of the chunk: its zeros join the previous chunk's run
      else e = zle_digits(z);
      e += 1;
    }
    emitted += e;
    if (nz) seen = true;
    const uint32_t nvalid = __popc(vmask);
    if (nz) zrun = nvalid - (32u - __clz(nz));
    else zrun += nvalid;
  }
#pragma unroll
  for (int dlt = 16; dlt > 0; dlt >>= 1) {
    emitted += __shfl_xor_sync(0xffffffffu, emitted, dlt);
    lead = max(lead, __shfl_xor_sync(0xffffffffu, lead, dlt));  // exactly one lane holds it
  }
  __syncwarp();
  int* cs = chunk_state + ((uint64_t)blockIdx.y * chunks_cap + chunk) * 256;
  for (int i = lane; i < 256; i += 32) cs[i] = s_last[w][i];
  if (lane == 0) {
    uint4 z;
    z.x = seen ? lead : len;
    z.y = seen ? zrun : 0u;
    z.z = emitted;
    z.w = seen ? 1u : 0u;
    chunk_zle[(uint64_t)blockIdx.y * chunks_cap + chunk] = z;
  }
}

// ---- pass B ---- one CTA (256 threads) per block
template <int MTF_CHUNK>
__global__ void __launch_bounds__(256) k3_chunk_scan_b(const BlockDesc* __restrict__ desc,
                                                       const uint32_t* __restrict__ inuse, int* __restrict__ chunk_state,
                                                       const uint4* __restrict__ chunk_zle,
                                                       uint2* __restrict__ chunk_base, uint32_t chunks_cap,
                                                       uint32_t* __restrict__ mtf_count) {
  const BlockDesc d = desc[blockIdx.x];
  const uint32_t nch = (d.n + MTF_CHUNK - 1) / MTF_CHUNK;
  const uint32_t* iu = inuse + blockIdx.x * 8;
  const int s = threadIdx.x;
  // virtual previous occurrence: -1 - dense(s) for in-use bytes
  int cur = NEG_UNUSED;
  if ((iu[s >> 5] >> (s & 31)) & 1u) {
    int dense = 0;
    for (int w = 0; w < (s >> 5); ++w) dense += __popc(iu[w]);
    dense += __popc(iu[s >> 5] & ((1u << (s & 31)) - 1u));
    cur = -1 - dense;
  }
  int* cs = chunk_state + (uint64_t)blockIdx.x * chunks_cap * 256;
  for (uint32_t c = 0; c < nch; ++c) {
    int v = cs[c * 256 + s];
    cs[c * 256 + s] = cur;
    cur = max(cur, v);
  }
  if (threadIdx.x == 0) {
    const uint4* cz = chunk_zle + (uint64_t)blockIdx.x * chunks_cap;
    uint2* cb = chunk_base + (uint64_t)blockIdx.x * chunks_cap;
    uint32_t carry = 0, out = 0;
    for (uint32_t c = 0; c < nch; ++c) {
      uint4 z = cz[c];
      cb[c] = make_uint2(out, carry);
      if (z.w) {
        out += zle_digits(carry + z.x) + z.z;
        carry = z.y;
      } else {
        carry += z.x;
      }
    }
    out += zle_digits(carry) + 1;  // final run + EOB (encoder.rs:355-358)
    mtf_count[blockIdx.x] = out;
  }
}

// ---- pass C ----
// One LANE per chunk: a warp runs 32 chunks side by side, each lane a scalar move-to-front coder of its own.
// The first 32 list entries of a lane live in eight packed registers (lookup = one SIMD byte compare per word,
// move-to-front = one byte permute per word); entries 32..255 live as packed words in a lane-interleaved
// shared-memory column and are only touched when a byte is found that deep (rare on BWT output, bounded work
// otherwise).  The list of each
// chunk at its start is rebuilt by the whole warp from the chunk's `last occurrence` state (rank of every byte's
// previous occurrence).
constexpr int MTF_FRONT = 32;                      // list entries held in registers (8 packed words)
constexpr int MTF_DEEP_WORDS = (256 - MTF_FRONT) / 4;  // the rest: packed words in shared memory, one column per lane;
                                                       // a launch keeps only as many as the batch's largest alphabet needs

struct MtfLane {
  uint32_t f[8];
};

// hit in word `f` at byte p: [carry, b0..b(p-1), b(p+1)..b3]
__device__ __forceinline__ uint32_t mtf_word_hit(uint32_t f, uint32_t carry, uint32_t p) {
  const uint32_t sel = (uint32_t)(0x2104310432043214ull >> (16 * p)) & 0xFFFFu;
  return __byte_perm(f, carry, sel);
}

__device__ __forceinline__ uint32_t shr_sat(uint32_t v, int s) {  // v >> s with s clamped to [0, 32] (32 gives 0)
  uint32_t r;
  asm("shr.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"((uint32_t)max(s, 0)));
  return r;
}

// Moves byte c (not at the front) to the front of the lane's list, returns its previous position.  The register
// part is branch-free per lane (lanes of a warp sit at different depths; a branch per word would serialise them):
//   find    per word the zero-byte flags of (word ^ c c c c): (z - 0x01010101) & ~z & 0x80808080 — a false flag can only
//           sit above a true one, so the lowest flag of the first word that has flags is c;
//   update  entry j takes the value of entry j-1 for j <= pos (a funnel shift by one byte across the words) and keeps
//           its own behind that: one byte mask per word, derived from pos.
// Words wholly behind the deepest hit among the lanes that are here are left alone (warp-uniform bound).
__device__ __forceinline__ uint32_t mtf_lane_access(MtfLane& l, uint32_t c, uint2* deep /* column, stride 32 */,
                                                    uint32_t deep_words /* 64-bit words: 8 entries each */) {
  const uint32_t x = c * 0x01010101u;
  uint32_t H = 0, wsel = 0;  // flags and index of the FIRST word with a hit (0xFF also matches the filler behind it)
#pragma unroll
  for (int k = 7; k >= 0; --k) {
    const uint32_t z = l.f[k] ^ x;
    const uint32_t h = (z - 0x01010101u) & ~z & 0x80808080u;
    H = h ? h : H;
    wsel = h ? (uint32_t)k : wsel;
  }
  const uint32_t pos = H ? 4u * wsel + ((uint32_t)(__ffs(H) - 1) >> 3) : 255u;  // 255: behind the register part
  const uint32_t wmax = __reduce_max_sync(__activemask(), min(pos >> 2, 7u));      // deepest word any lane touches
  const int A = 32 - 8 * (int)(pos + 1u);  // word k takes the shifted bytes in its low clamp(pos + 1 - 4k, 0, 4) bytes
  uint32_t prev = c << 24;
  const uint32_t carry_out = l.f[7] >> 24;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if ((uint32_t)k <= wmax) {
      const uint32_t f = l.f[k];
      const uint32_t shifted = __funnelshift_l(prev, f, 8);
      const uint32_t mask = shr_sat(0xFFFFFFFFu, min(A + 32 * k, 32));
      l.f[k] = (shifted & mask) | (f & ~mask);
      prev = f;
    }
  }
  uint32_t carry = carry_out;
  const uint32_t before = H ? 0u : 1u;
  const bool done = before == 0;
  if (done) return pos;
  for (uint32_t k = 0; k < deep_words; ++k) {  // every entry in front of c moves down by one
    const uint2 f = deep[k * 32];
    const uint32_t zx = f.x ^ x, zy = f.y ^ x;  // zero-byte flags: the lowest one of a word is exact
    const uint32_t m0 = (zx - 0x01010101u) & ~zx & 0x80808080u, m1 = (zy - 0x01010101u) & ~zy & 0x80808080u;
    if (m0) {
      const uint32_t p = (__ffs(m0) - 1) >> 3;
      deep[k * 32].x = mtf_word_hit(f.x, carry, p);  // the high half sits behind the hit: unchanged
      return MTF_FRONT + 8 * k + p;
    }
    if (m1) {
      const uint32_t p = (__ffs(m1) - 1) >> 3;
      deep[k * 32] = make_uint2((f.x << 8) | carry, mtf_word_hit(f.y, f.x >> 24, p));
      return MTF_FRONT + 8 * k + 4 + p;
    }
    deep[k * 32] = make_uint2((f.x << 8) | carry, (f.y << 8) | (f.x >> 24));
    carry = f.y >> 24;
  }
  return 255;
}

struct MtfOut {
  uint16_t* out;
  uint32_t o;        // next symbol index
  uint32_t pend;     // symbol waiting for its pair (stores are 32-bit where aligned)
  bool have;
  unsigned long long hot;  // counts of symbols 0..3, 16 bits each
  uint32_t* freq;          // shared-memory histogram
  __device__ __forceinline__ void put(uint32_t sv) {
    if (sv < 4) hot += 1ull << (16 * sv);
    else atomicAdd(&freq[sv], 1u);
    if (have) {
      *reinterpret_cast<uint32_t*>(out + o - 1) = pend | (sv << 16);
      have = false;
    } else if ((o & 1u) == 0) {
      pend = sv;
      have = true;
    } else {
      out[o] = (uint16_t)sv;
    }
    ++o;
  }
  __device__ __forceinline__ void zero_run(uint32_t z) {  // RUNA/RUNB digits of a run of z zeros (encoder.rs:653-669)
    uint32_t zz = z + 1;
    while (zz > 1) {
      put(zz & 1u);
      zz >>= 1;
    }
  }
  __device__ __forceinline__ void flush() {
    if (have) out[o - 1] = (uint16_t)pend;
    have = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t v = (uint32_t)(hot >> (16 * k)) & 0xFFFFu;
      if (v) atomicAdd(&freq[k], v);
    }
    hot = 0;
  }
};

template <int MTF_CHUNK>
__global__ void __launch_bounds__(MTF_WARPS * 32) k3_apply(const uint8_t* __restrict__ last,
                                                           const BlockDesc* __restrict__ desc,
                                                           const uint32_t* __restrict__ inuse,
                                                           const int* __restrict__ chunk_state,
                                                           const uint2* __restrict__ chunk_base, uint32_t chunks_cap,
                                                           uint32_t nb, uint32_t groups_cap, uint32_t deep_words,
                                                           uint16_t* __restrict__ sym, uint32_t* __restrict__ freq) {
  // deep_words = 64-bit words of the shared-memory tail every lane keeps: ceil((largest alphabet of the batch - 32) / 8);
  // text (~70 symbols) needs 5 of the 28 a full byte alphabet takes, which is what lets 2-3x more warps live on an SM
  __shared__ uint8_t s_front[MTF_WARPS][MTF_FRONT];
  __shared__ uint32_t s_freqw[MTF_WARPS][MAX_ALPHA + 2];
  extern __shared__ __align__(16) uint2 s_deep_raw[];  // [MTF_WARPS][deep_words][32], 8 list entries per 64-bit word
  const int w = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  uint2* s_deepw = s_deep_raw + (size_t)w * deep_words * 32;
  // warp task = (block, group of 32 chunks); the warps of a CTA may belong to different blocks
  const uint32_t task = blockIdx.x * MTF_WARPS + w;
  const uint32_t blk = task / groups_cap;
  if (blk >= nb) return;
  const BlockDesc d = desc[blk];
  uint32_t* s_freq = s_freqw[w];
  for (int i = lane; i < MAX_ALPHA + 2; i += 32) s_freq[i] = 0;
  __syncwarp();
  const uint32_t nch = (d.n + MTF_CHUNK - 1) / MTF_CHUNK;
  const uint32_t cbase = (task % groups_cap) * 32;  // first chunk of this warp
  if (cbase < nch) {
    // ---- the list of every lane's chunk at the chunk start
    MtfLane ml;
#pragma unroll
    for (int k = 0; k < 8; ++k) ml.f[k] = 0;
    const uint32_t nmine = min(32u, nch - cbase);
    for (uint32_t j = 0; j < nmine; ++j) {
      const int* cs = chunk_state + ((uint64_t)blk * chunks_cap + cbase + j) * 256;
      int mine[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) mine[k] = cs[k * 32 + lane];
      // filler behind the in-use bytes (a real 0xFF always sits in front of it)
      s_front[w][lane] = 0xFFu;
      for (uint32_t q = lane; q < deep_words; q += 32) s_deepw[q * 32 + j] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
      // rank of every byte's previous occurrence among the in-use bytes = its list position; the 256 values are
      // broadcast from the registers that hold them
      uint32_t rank[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int k2 = 0; k2 < 8; ++k2) {
        for (int l = 0; l < 32; ++l) {
          const int v = __shfl_sync(0xffffffffu, mine[k2], l);
          if (v == NEG_UNUSED) continue;  // warp-uniform
#pragma unroll
          for (int k = 0; k < 8; ++k) rank[k] += v > mine[k];
        }
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (mine[k] != NEG_UNUSED) {
          const uint32_t p = rank[k];
          if (p < (uint32_t)MTF_FRONT) s_front[w][p] = (uint8_t)(k * 32 + lane);
          else reinterpret_cast<uint8_t*>(&s_deepw[((p - MTF_FRONT) >> 3) * 32 + j])[(p - MTF_FRONT) & 7u] = (uint8_t)(k * 32 + lane);
        }
      }
      __syncwarp();
      if (lane == j) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          ml.f[k] = s_front[w][4 * k] | (s_front[w][4 * k + 1] << 8) | (s_front[w][4 * k + 2] << 16) |
                    ((uint32_t)s_front[w][4 * k + 3] << 24);
      }
      __syncwarp();
    }

    // ---- every lane codes its own chunk
    const uint32_t chunk = cbase + lane;
    if (chunk < nch) {
      const uint32_t c0 = chunk * MTF_CHUNK;
      const uint32_t len = min((uint32_t)MTF_CHUNK, d.n - c0);
      const uint8_t* L = last + d.off + c0;
      const uint2 cb = chunk_base[(uint64_t)blk * chunks_cap + chunk];
      MtfOut mo;
      mo.out = sym + d.symoff;
      mo.o = cb.x;
      mo.pend = 0;
      mo.have = false;
      mo.hot = 0;
      mo.freq = s_freq;
      uint32_t zrun = cb.y;
      uint32_t prev = c0 > 0 ? L[-1] : smallest_inuse(inuse + blk * 8);
      uint2* deep = s_deepw + lane;
      auto step = [&](uint32_t c) {
        if (c == prev) {
          ++zrun;
        } else {
          mo.zero_run(zrun);
          zrun = 0;
          mo.put(mtf_lane_access(ml, c, deep, deep_words) + 1u);  // position p > 0 is written as p+1 (encoder.rs:340)
          prev = c;
        }
      };
      uint32_t i = 0;
      const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(L) & 15u);
      const uint32_t headn = min(len, mis ? 16u - mis : 0u);
      for (; i < headn; ++i) step(L[i]);
      for (; i + 16 <= len; i += 16) {
        const uint4 q = *reinterpret_cast<const uint4*>(L + i);
        unsigned long long lo = q.x | ((unsigned long long)q.y << 32), hi = q.z | ((unsigned long long)q.w << 32);
#pragma unroll 1
        for (int k = 0; k < 8; ++k) { step((uint32_t)lo & 255u); lo >>= 8; }
#pragma unroll 1
        for (int k = 0; k < 8; ++k) { step((uint32_t)hi & 255u); hi >>= 8; }
      }
      for (; i < len; ++i) step(L[i]);
      if (c0 + len == d.n) {
        // final zero run + EOB (encoder.rs:355-358); EOB = in_use_count + 1
        mo.zero_run(zrun);
        uint32_t k = 0;
        for (int ww = 0; ww < 8; ++ww) k += __popc(inuse[blk * 8 + ww]);
        mo.put(k + 1);
      }
      mo.flush();
    }
  }
  __syncwarp();
  uint32_t* f = freq + (uint64_t)blk * MAX_ALPHA;
  for (int i = lane; i < MAX_ALPHA; i += 32) {
    uint32_t v = s_freq[i];
    if (v) atomicAdd(&f[i], v);
  }
}

template <int MTF_CHUNK>
static void launch_mtf_t(Launcher& L, const uint8_t* d_last, const BlockDesc* d_desc, const uint32_t* d_inuse, uint32_t nb,
                         uint32_t nmax, uint32_t max_alpha_bytes, int* d_chunk_state, uint4* d_chunk_zle,
                         uint2* d_chunk_base, uint32_t chunks_cap, uint16_t* d_sym, uint32_t* d_freq,
                         uint32_t* d_mtf_count) {
  const uint32_t nch = (nmax + MTF_CHUNK - 1) / MTF_CHUNK;
  const uint32_t gx = (nch + MTF_WARPS - 1) / MTF_WARPS;
  cudaMemsetAsync(d_freq, 0, (size_t)nb * MAX_ALPHA * sizeof(uint32_t), L.stream);
  L.launch("k3_chunk_scan_a", k3_chunk_scan_a<MTF_CHUNK>, dim3(gx, nb), dim3(MTF_WARPS * 32), d_last, d_desc, d_inuse,
           d_chunk_state, d_chunk_zle, chunks_cap);
  L.launch("k3_chunk_scan_b", k3_chunk_scan_b<MTF_CHUNK>, dim3(nb), dim3(256), d_desc, d_inuse, d_chunk_state,
           (const uint4*)d_chunk_zle, d_chunk_base, chunks_cap, d_mtf_count);
  const uint32_t groups_cap = (nch + 31) / 32;  // warp tasks per block: one lane per chunk
  const uint32_t ntask = groups_cap * nb;
  // in-use bytes of the batch's largest alphabet beyond the 32 register entries, in packed words (at least one)
  const uint32_t a = max_alpha_bytes < 1 || max_alpha_bytes > 256 ? 256u : max_alpha_bytes;
  const uint32_t deep_words = a > (uint32_t)MTF_FRONT ? (a - MTF_FRONT + 7) / 8 : 1u;  // 64-bit words
  L.launch_smem("k3_apply", k3_apply<MTF_CHUNK>, dim3((ntask + MTF_WARPS - 1) / MTF_WARPS), dim3(MTF_WARPS * 32),
                (size_t)MTF_WARPS * deep_words * 32 * sizeof(uint2), d_last, d_desc, d_inuse, (const int*)d_chunk_state,
                (const uint2*)d_chunk_base, chunks_cap, nb, groups_cap, deep_words, d_sym, d_freq);
}

// `chunk` = mtf_chunk_elems(batch bytes), the value the chunk arrays were sized with (chunks_cap chunks per block)
void launch_mtf(Launcher& L, const uint8_t* d_last, const BlockDesc* d_desc, const uint32_t* d_inuse, uint32_t nb,
                uint32_t nmax, uint32_t max_alpha_bytes, uint32_t chunk, int* d_chunk_state, uint4* d_chunk_zle,
                uint2* d_chunk_base, uint32_t chunks_cap, uint16_t* d_sym, uint32_t* d_freq, uint32_t* d_mtf_count) {
  switch (chunk) {
    case 1024:
      launch_mtf_t<1024>(L, d_last, d_desc, d_inuse, nb, nmax, max_alpha_bytes, d_chunk_state, d_chunk_zle, d_chunk_base,
                         chunks_cap, d_sym, d_freq, d_mtf_count);
      break;
    case 2048:
      launch_mtf_t<2048>(L, d_last, d_desc, d_inuse, nb, nmax, max_alpha_bytes, d_chunk_state, d_chunk_zle, d_chunk_base,
                         chunks_cap, d_sym, d_freq, d_mtf_count);
      break;
    default:
      launch_mtf_t<4096>(L, d_last, d_desc, d_inuse, nb, nmax, max_alpha_bytes, d_chunk_state, d_chunk_zle, d_chunk_base,
                         chunks_cap, d_sym, d_freq, d_mtf_count);
      break;
  }
}

}  // namespace bzb
