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
//   pass C  per chunk: the list at the chunk start is rebuilt from `last` (rank of every byte's previous occurrence)
//           and kept in registers (8 per lane); run heads are looked up by ballot and moved to the front by
//           shuffles; RUNA/RUNB expansion by a warp scan of the digit counts, symbol histogram in shared memory.
#include <limits.h>

#include "common.cuh"
#include "kernels.h"

namespace bzb {

constexpr int MTF_CHUNK = 4096;
constexpr int MTF_WARPS = 4;  // warps (= chunks) per CTA
constexpr int NEG_UNUSED = -(1 << 30);

uint32_t mtf_chunk_elems() { return MTF_CHUNK; }

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
// The warp keeps the whole MTF list in registers: list position p lives in lane p % 32, register p / 32.  A run head
// looks its byte up with one ballot per row of 32 positions (row 0 first — BWT output mostly hits the front of the
// list) and the entries in front of it move down by one lane (shfl_up); nothing touches memory.
__device__ __forceinline__ uint32_t mtf_access(uint32_t (&lst)[8], uint32_t c, uint32_t lane) {
  uint32_t hit = __ballot_sync(0xffffffffu, lst[0] == c);
  if (hit) {  // front row
    const uint32_t q = __ffs(hit) - 1;
    const uint32_t up = __shfl_up_sync(0xffffffffu, lst[0], 1);
    if (lane <= q) lst[0] = lane ? up : c;
    return q;
  }
  uint32_t carry = c;  // value entering lane 0 of the current row
  uint32_t pos = 0;
  bool done = false;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    if (!done) {  // warp-uniform
      if (r) hit = __ballot_sync(0xffffffffu, lst[r] == c);
      const uint32_t q = hit ? __ffs(hit) - 1 : 32u;  // entries 0..q-1 of this row move; q == 32: the whole row
      const uint32_t up = __shfl_up_sync(0xffffffffu, lst[r], 1);
      const uint32_t out = __shfl_sync(0xffffffffu, lst[r], 31);
      if (lane <= q) lst[r] = lane ? up : carry;
      carry = out;
      if (hit) { pos = r * 32 + q; done = true; }
    }
  }
  return pos;
}

__global__ void __launch_bounds__(MTF_WARPS * 32) k3_apply(const uint8_t* __restrict__ last,
                                                           const BlockDesc* __restrict__ desc,
                                                           const uint32_t* __restrict__ inuse,
                                                           const int* __restrict__ chunk_state,
                                                           const uint2* __restrict__ chunk_base, uint32_t chunks_cap,
                                                           uint16_t* __restrict__ sym, uint32_t* __restrict__ freq) {
  __shared__ int s_last[MTF_WARPS][256];
  __shared__ uint32_t s_list[MTF_WARPS][256];
  __shared__ uint32_t s_freq[MAX_ALPHA + 2];
  const int w = threadIdx.x >> 5;
  const uint32_t lane = lane_id();
  const BlockDesc d = desc[blockIdx.y];
  for (int i = threadIdx.x; i < MAX_ALPHA + 2; i += MTF_WARPS * 32) s_freq[i] = 0;
  __syncthreads();
  const uint32_t chunk = blockIdx.x * MTF_WARPS + w;
  const uint32_t c0 = chunk * MTF_CHUNK;
  if (c0 < d.n) {
    const uint32_t len = min((uint32_t)MTF_CHUNK, d.n - c0);
    const bool last_chunk = c0 + len == d.n;
    const uint8_t* L = last + d.off;
    const int* cs = chunk_state + ((uint64_t)blockIdx.y * chunks_cap + chunk) * 256;
    // list at the chunk start: bytes by decreasing previous occurrence (virtual occurrences for unseen in-use bytes)
    int mine[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      mine[k] = cs[k * 32 + lane];
      s_last[w][k * 32 + lane] = mine[k];
      s_list[w][k * 32 + lane] = 0x1FFu;  // never matches a byte
    }
    __syncwarp();
    {
      uint32_t rank[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int j = 0; j < 256; ++j) {
        const int v = s_last[w][j];
        if (v == NEG_UNUSED) continue;  // warp-uniform
#pragma unroll
        for (int k = 0; k < 8; ++k) rank[k] += v > mine[k];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (mine[k] != NEG_UNUSED) s_list[w][rank[k]] = k * 32 + lane;
    }
    __syncwarp();
    uint32_t lst[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) lst[k] = s_list[w][k * 32 + lane];

    const uint2 cb = chunk_base[(uint64_t)blockIdx.y * chunks_cap + chunk];
    uint16_t* out = sym + d.symoff;
    uint32_t obase = cb.x;
    uint32_t zrun = cb.y;
    uint8_t prev = c0 > 0 ? L[c0 - 1] : smallest_inuse(inuse + blockIdx.y * 8);
    for (uint32_t i0 = 0; i0 < len; i0 += 32) {
      const uint32_t i = i0 + lane;
      const bool valid = i < len;
      const uint32_t c = valid ? L[c0 + i] : 0u;
      uint32_t p = __shfl_up_sync(0xffffffffu, c, 1);
      if (lane == 0) p = prev;
      prev = (uint8_t)__shfl_sync(0xffffffffu, c, 31);
      const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
      const uint32_t nz = __ballot_sync(0xffffffffu, valid && c != p);
      // MTF positions of the run heads, in order
      uint32_t mypos = 0;
      uint32_t m = nz;
      while (m) {
        const uint32_t l = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t hc = __shfl_sync(0xffffffffu, c, l);
        const uint32_t pos = mtf_access(lst, hc, lane);
        if (lane == l) mypos = pos;
      }
      // zero-run bookkeeping and emission
      const bool is_nz = (nz >> lane) & 1u;
      uint32_t z = 0;
      if (is_nz) {
        uint32_t below = nz & lanemask_lt();
        if (below) z = lane - (32u - __clz(below));   // zeros between the previous head in this row and me
        else z = zrun + lane;                         // reaches back into previous rows/chunks
      }
      uint32_t dg = is_nz ? zle_digits(z) : 0u;
      uint32_t ecount = is_nz ? dg + 1u : 0u;
      uint32_t inc = warp_incl_scan_add(ecount);
      uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
      if (is_nz) {
        uint32_t o = obase + inc - ecount;
        uint32_t zz = z + 1;
        for (uint32_t k = 0; k < dg; ++k) {
          uint32_t bit = zz & 1u;
          out[o++] = (uint16_t)bit;
          atomicAdd(&s_freq[bit], 1u);
          zz >>= 1;
        }
        uint32_t sv = mypos + 1u;  // position p>0 is written as p+1 (encoder.rs:340)
        out[o] = (uint16_t)sv;
        atomicAdd(&s_freq[sv], 1u);
      }
      obase += total;
      const uint32_t nvalid = __popc(vmask);
      if (nz) zrun = nvalid - (32u - __clz(nz));
      else zrun += nvalid;
    }
    if (last_chunk && lane == 0) {
      // final zero run + EOB (encoder.rs:355-358); EOB = in_use_count + 1
      uint32_t o = obase;
      uint32_t zz = zrun + 1;
      for (uint32_t k = 0, dg = zle_digits(zrun); k < dg; ++k) {
        uint32_t bit = zz & 1u;
        out[o++] = (uint16_t)bit;
        atomicAdd(&s_freq[bit], 1u);
        zz >>= 1;
      }
      uint32_t k = 0;
      for (int ww = 0; ww < 8; ++ww) k += __popc(inuse[blockIdx.y * 8 + ww]);
      out[o] = (uint16_t)(k + 1);
      atomicAdd(&s_freq[k + 1], 1u);
    }
  }
  __syncthreads();
  uint32_t* f = freq + (uint64_t)blockIdx.y * MAX_ALPHA;
  for (int i = threadIdx.x; i < MAX_ALPHA; i += MTF_WARPS * 32) {
    uint32_t v = s_freq[i];
    if (v) atomicAdd(&f[i], v);
  }
}

void launch_mtf(Launcher& L, const uint8_t* d_last, const BlockDesc* d_desc, const uint32_t* d_inuse, uint32_t nb,
                uint32_t nmax, int* d_chunk_state, uint4* d_chunk_zle, uint2* d_chunk_base, uint32_t chunks_cap,
                uint16_t* d_sym, uint32_t* d_freq, uint32_t* d_mtf_count) {
  const uint32_t nch = (nmax + MTF_CHUNK - 1) / MTF_CHUNK;
  const uint32_t gx = (nch + MTF_WARPS - 1) / MTF_WARPS;
  cudaMemsetAsync(d_freq, 0, (size_t)nb * MAX_ALPHA * sizeof(uint32_t), L.stream);
  L.launch("k3_chunk_scan_a", k3_chunk_scan_a, dim3(gx, nb), dim3(MTF_WARPS * 32), d_last, d_desc, d_inuse,
           d_chunk_state, d_chunk_zle, chunks_cap);
  L.launch("k3_chunk_scan_b", k3_chunk_scan_b, dim3(nb), dim3(256), d_desc, d_inuse, d_chunk_state,
           (const uint4*)d_chunk_zle, d_chunk_base, chunks_cap, d_mtf_count);
  L.launch("k3_apply", k3_apply, dim3(gx, nb), dim3(MTF_WARPS * 32), d_last, d_desc, d_inuse,
           (const int*)d_chunk_state, (const uint2*)d_chunk_base, chunks_cap, d_sym, d_freq);
}

}  // namespace bzb
