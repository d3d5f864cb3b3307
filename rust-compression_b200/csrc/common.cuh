// common.cuh — shared device helpers and the internal launcher interface.
// sm_100a only; no multi-arch dispatch.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

namespace bzb {

// Kernel attributes (cudaFuncSetAttribute) are per DEVICE: run(f) executes f once per device (the calling thread's
// current one) and makes concurrent callers on the same device wait until it is done, so that a process that drives
// several GPUs and several contexts per GPU — the multi-GPU engine, mgpu.cu — configures its kernels on each of them.
struct PerDeviceOnce {
  std::mutex mu;
  unsigned long long done[2] = {0, 0};
  template <class F>
  void run(F f) {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d > 127) d = 0;
    std::lock_guard<std::mutex> g(mu);
    const unsigned long long bit = 1ull << (d & 63);
    if (done[d >> 6] & bit) return;
    f();
    done[d >> 6] |= bit;
  }
};

// ---- constants of the format (bzip2/mod.rs:20, encoder.rs:186,294-298) ----
constexpr int G_SIZE = 50;          // BZ_G_SIZE
constexpr int N_ITERS = 4;          // BZ_N_ITERS
constexpr int MAX_ALPHA = 258;      // 256 symbols + RUNB shift + EOB
constexpr int MAX_GROUPS = 6;
constexpr int MAX_SELECTORS = 18002;  // 2 + 900000/50
constexpr int MAX_CODE_LEN = 17;
constexpr uint32_t MAX_BLOCK = 900000;  // level*100000 upper bound (block_buf capacity)

// ---- rank word layout used by the rotation sort ----
constexpr uint32_t RANK_RESOLVED = 0x80000000u;
constexpr uint32_t RANK_MASK = 0x000FFFFFu;  // 20 bits: n <= 900000 < 2^20

// Per-block descriptor of a batch (device memory, one entry per block in the batch).
struct BlockDesc {
  uint32_t off;     // offset of the block's bytes inside the batch's RLE1 slice
  uint32_t n;       // block length after RLE1 (1 .. level*100000-15)
  uint32_t symoff;  // offset of the block's MTF/ZLE symbol region (u16 units) inside the batch
  uint32_t pad;
};

// ---- small device utilities ----
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

__device__ __forceinline__ uint32_t warp_incl_scan_add(uint32_t v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
    if ((int)lane_id() >= d) v += t;
  }
  return v;
}
__device__ __forceinline__ int warp_incl_scan_max(int v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, d);
    if ((int)lane_id() >= d) v = max(v, t);
  }
  return v;
}
__device__ __forceinline__ long long warp_incl_scan_max64(long long v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    long long t = __shfl_up_sync(0xffffffffu, v, d);
    if ((int)lane_id() >= d) v = max(v, t);
  }
  return v;
}

// CTA-wide exclusive sum scan of one u32 per thread. NT = threads per CTA (multiple of 32, <= 1024).
// Returns the exclusive prefix; *total receives the CTA total. `ws` = shared scratch of NT/32 words.
template <int NT>
__device__ __forceinline__ uint32_t cta_excl_scan_add(uint32_t v, uint32_t* ws, uint32_t* total) {
  uint32_t inc = warp_incl_scan_add(v);
  const int w = threadIdx.x >> 5;
  if (lane_id() == 31) ws[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t x = lane_id() < NT / 32 ? ws[lane_id()] : 0u;
    uint32_t xi = warp_incl_scan_add(x);
    if (lane_id() < NT / 32) ws[lane_id()] = xi - x;
    if (lane_id() == 31) ws[NT / 32] = xi;  // total (ws needs NT/32+1 words)
  }
  __syncthreads();
  uint32_t r = inc - v + ws[w];
  if (total) *total = ws[NT / 32];
  __syncthreads();
  return r;
}

// CTA-wide exclusive max scan of one long long per thread (identity = -1).
template <int NT>
__device__ __forceinline__ long long cta_excl_scan_max64(long long v, long long* ws) {
  long long inc = warp_incl_scan_max64(v);
  const int w = threadIdx.x >> 5;
  if (lane_id() == 31) ws[w] = inc;
  __syncthreads();
  if (w == 0) {
    long long x = lane_id() < NT / 32 ? ws[lane_id()] : -1ll;
    long long xi = warp_incl_scan_max64(x);
    long long xe = __shfl_up_sync(0xffffffffu, xi, 1);
    if (lane_id() == 0) xe = -1ll;
    if (lane_id() < NT / 32) ws[lane_id()] = xe;
  }
  __syncthreads();
  long long up = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane_id() == 0) up = -1ll;
  long long r = max(up, ws[w]);
  __syncthreads();
  return r;
}

// MSB-first bit position helpers: the output stream is addressed in bits; bit p lives in byte p>>3 at
// mask 0x80>>(p&7).  We OR 32-bit big-endian words: logical word w (bits 32w..32w+31, MSB first) is stored
// byte-swapped in little-endian memory.
__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// OR `len` (1..32) low bits of `value` into the stream at bit position `pos`. words = (uint32_t*)out (4-byte aligned).
__device__ __forceinline__ void put_bits_atomic(uint32_t* words, uint64_t pos, uint32_t value, uint32_t len) {
  if (len == 0) return;
  if (len < 32) value &= (1u << len) - 1u;
  uint64_t w = pos >> 5;
  uint32_t sh = (uint32_t)(pos & 31);  // bits already used in word w
  uint64_t v64 = ((uint64_t)value) << (64 - len);  // left-aligned in 64 bits
  v64 >>= sh;
  uint32_t hi = (uint32_t)(v64 >> 32), lo = (uint32_t)v64;
  if (hi) atomicOr(&words[w], bswap32(hi));
  if (lo) atomicOr(&words[w + 1], bswap32(lo));
}

}  // namespace bzb
