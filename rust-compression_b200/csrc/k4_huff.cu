// k4_huff.cu — K4: the 2..6 Huffman-table selection / refinement loop, code-length construction and canonical codes.
//
// Replaces, by result: table count + initial tables (src/bzip2/encoder.rs:367-426), the 4 refinement passes
// (:428-509), create_huffman (:641-651) -> make_tab_with_fn / gen_code / gen_code_lm / down_heap / create_heap /
// take_package (src/huffman/cano_huff_table.rs:14-225), selector MTF (:511-517) and canonical code assignment
// (src/huffman/mod.rs:22-67 with src/bucket_sort.rs:43-75).
//
// The heap construction is inherently sequential and its tie-breaks decide the output bits, so every table is
// built by ONE thread that replays the reference's array-heap operation for operation (same index arithmetic,
// same strict/non-strict comparisons); parallelism comes from the (blocks x tables) in flight.  The per-group
// cost/selection step is data parallel: one thread per 50-symbol group with the 6 table lengths of a symbol
// packed into one 64-bit word (6 x 10-bit lanes; 50 x 17 = 850 < 1024, so lanes never carry).
#include "common.cuh"
#include "kernels.h"

namespace bzb {

constexpr int LENS_SLOTS = 5;  // initial + after each of the 4 passes
__host__ __device__ __forceinline__ size_t lens_index(uint32_t b, int slot, int t) {
  return (((size_t)b * LENS_SLOTS + slot) * MAX_GROUPS + t) * MAX_ALPHA;
}

// meta[b][8]: 0 alpha, 1 ngroups, 2 nselectors, 3 hdr_bits, 4 data_bits, 5 lm_count, 6 error, 7 spare
constexpr int META = 8;

// ---------------------------------------------------------------- initial tables (encoder.rs:367-426)
__global__ void k4_init(uint32_t nb, const uint32_t* __restrict__ mtf_count, const uint32_t* __restrict__ freq,
                        const uint32_t* __restrict__ inuse, uint8_t* __restrict__ lens, uint32_t* __restrict__ meta,
                        uint32_t* __restrict__ rfreq) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  uint32_t k = 0;
  for (int w = 0; w < 8; ++w) k += __popc(inuse[b * 8 + w]);
  const int alpha = (int)k + 2;
  const uint32_t mc = mtf_count[b];
  const int ng = mc < 200 ? 2 : mc < 600 ? 3 : mc < 1200 ? 4 : mc < 2400 ? 5 : 6;
  const uint32_t* f = freq + (size_t)b * MAX_ALPHA;
  uint32_t rem = mc;
  int gs = 0;
  for (int n_part = ng; n_part >= 1; --n_part) {
    uint32_t t_freq = rem / (uint32_t)n_part;
    int ge = gs - 1;
    uint32_t a_freq = 0;
    while (a_freq < t_freq && ge < alpha - 1) {
      ge += 1;
      a_freq += f[ge];
    }
    if (ge > gs && n_part != ng && n_part != 1 && (((ng - n_part) & 1) == 1)) {
      a_freq -= f[ge];
      ge -= 1;
    }
    uint8_t* l = lens + lens_index(b, 0, n_part - 1);  // libbzip2 table id = n_part-1
    for (int i = 0; i < alpha; ++i) l[i] = (i >= gs && i <= ge) ? 0 : 15;
    rem -= a_freq;
    gs = ge + 1;
  }
  uint32_t* m = meta + (size_t)b * META;
  m[0] = (uint32_t)alpha;
  m[1] = (uint32_t)ng;
  m[2] = (mc + G_SIZE - 1) / G_SIZE;
  m[3] = 0; m[4] = 0; m[5] = 0; m[6] = 0; m[7] = 0;
  uint32_t* rf = rfreq + (size_t)b * MAX_GROUPS * MAX_ALPHA;
  for (int i = 0; i < MAX_GROUPS * MAX_ALPHA; ++i) rf[i] = 0;
}

// ---------------------------------------------------------------- cost + selection (encoder.rs:440-481)
// mode 0: choose the first-minimum table per group, write selector, accumulate rfreq.
// mode 1: selectors fixed; write the bit length of each group under its table (final tables) into gbits.
constexpr int CS_NT = 256;
constexpr int CS_SYMS = CS_NT * G_SIZE;  // symbols of a CTA's 256 groups: 12 800 (25.6 KB)
static_assert((CS_SYMS * 2) % 16 == 0, "a CTA's symbols start on a 16-byte boundary of the block's (16-byte aligned) region");
__global__ void __launch_bounds__(CS_NT) k4_cost_select(const uint16_t* __restrict__ sym,
                                                        const BlockDesc* __restrict__ desc,
                                                        const uint32_t* __restrict__ mtf_count,
                                                        const uint8_t* __restrict__ lens, int slot,
                                                        const uint32_t* __restrict__ meta, uint8_t* __restrict__ sel,
                                                        uint32_t* __restrict__ rfreq, uint32_t* __restrict__ gbits,
                                                        int mode) {
  __shared__ unsigned long long cost[MAX_ALPHA];
  __shared__ uint32_t hist[MAX_GROUPS][MAX_ALPHA];
  __shared__ __align__(16) uint16_t ssym[CS_SYMS];
  const uint32_t b = blockIdx.y;
  const uint32_t* m = meta + (size_t)b * META;
  const int alpha = (int)m[0], ng = (int)m[1];
  const uint32_t nsel = m[2];
  if (blockIdx.x * CS_NT >= nsel) return;
  // the CTA's symbols, coalesced (a thread's 50 symbols are 100 bytes apart from its neighbour's: read straight from
  // global memory they cost 50 two-byte loads per thread); the region of a block is padded to whole 16-byte vectors
  const uint32_t mc = mtf_count[b];
  const uint32_t s0 = blockIdx.x * CS_SYMS;
  {
    const uint4* src = reinterpret_cast<const uint4*>(sym + desc[b].symoff + s0);
    const uint32_t nvec = (min((uint32_t)CS_SYMS, mc - s0) + 7u) / 8u;
    for (uint32_t i = threadIdx.x; i < nvec; i += CS_NT) reinterpret_cast<uint4*>(ssym)[i] = src[i];
  }
  for (int s = threadIdx.x; s < alpha; s += CS_NT) {
    unsigned long long c = 0;
    for (int t = 0; t < ng; ++t) c |= (unsigned long long)lens[lens_index(b, slot, t) + s] << (10 * t);
    cost[s] = c;
  }
  if (mode == 0)
    for (int i = threadIdx.x; i < MAX_GROUPS * MAX_ALPHA; i += CS_NT) (&hist[0][0])[i] = 0;
  __syncthreads();
  const uint32_t g = blockIdx.x * CS_NT + threadIdx.x;
  if (g < nsel) {
    const uint32_t gs = g * G_SIZE, ge = min(gs + G_SIZE, mc);
    const uint32_t n = ge - gs;
    const uint16_t* v = ssym + threadIdx.x * G_SIZE;  // 25 words per thread: conflict-free
    unsigned long long sum = 0;
#pragma unroll
    for (int i = 0; i < G_SIZE; ++i)
      if ((uint32_t)i < n) sum += cost[v[i]];
    if (mode == 0) {
      int bt = 0;
      uint32_t bc = (uint32_t)(sum & 1023ull);
      for (int t = 1; t < ng; ++t) {
        uint32_t c = (uint32_t)((sum >> (10 * t)) & 1023ull);
        if (c < bc) { bc = c; bt = t; }  // min_by keeps the FIRST minimum (encoder.rs:452-467)
      }
      sel[(size_t)b * MAX_SELECTORS + g] = (uint8_t)bt;
#pragma unroll
      for (int i = 0; i < G_SIZE; ++i)
        if ((uint32_t)i < n) atomicAdd(&hist[bt][v[i]], 1u);
    } else {
      const int t = sel[(size_t)b * MAX_SELECTORS + g];
      gbits[(size_t)b * MAX_SELECTORS + g] = (uint32_t)((sum >> (10 * t)) & 1023ull);
    }
  }
  if (mode == 0) {
    __syncthreads();
    uint32_t* rf = rfreq + (size_t)b * MAX_GROUPS * MAX_ALPHA;
    for (int i = threadIdx.x; i < ng * MAX_ALPHA; i += CS_NT) {
      uint32_t c = (&hist[0][0])[i];
      if (c) atomicAdd(&rf[i], c);
    }
  }
}

// ---------------------------------------------------------------- code lengths (cano_huff_table.rs)
// buf layout as in gen_code (:153-196): buf[0..n) heap of node ids, buf[n..2n) leaf weights; internal node i
// reuses slot i once the heap has shrunk below it.
__device__ __forceinline__ uint32_t bz_weight_add(uint32_t x, uint32_t y) {  // encoder.rs:647-650
  return ((x & 0xFFFFFF00u) + (y & 0xFFFFFF00u)) | (1u + max(x & 0xFFu, y & 0xFFu));
}

__device__ void down_heap(uint32_t* buf, uint32_t n, uint32_t len) {  // cano_huff_table.rs:14-31
  const uint32_t tmp = buf[n];
  uint32_t leaf = (n << 1) + 1;
  while (leaf < len) {
    if (leaf + 1 < len && buf[buf[leaf]] > buf[buf[leaf + 1]]) leaf += 1;
    if (buf[tmp] < buf[buf[leaf]]) break;
    buf[n] = buf[leaf];
    n = leaf;
    leaf = (n << 1) + 1;
  }
  buf[n] = tmp;
}

constexpr int HB_NT = 32;
constexpr int LM_LEVELS = MAX_CODE_LEN;         // lim = 17
constexpr int LM_ROW = 2 * MAX_ALPHA;           // max_elem[j] <= 2*len
constexpr size_t LM_BYTES = (size_t)LM_LEVELS * LM_ROW * (sizeof(uint32_t) + sizeof(uint16_t));
size_t huff_lm_scratch_bytes() { return LM_BYTES; }

// Reverse package merge (cano_huff_table.rs:58-151), lim = 17, weights = bzip2 weights. Returns 0 or error.
__device__ int gen_code_lm_dev(const uint32_t* __restrict__ wt, int len, uint8_t* __restrict__ out_len,
                               uint32_t* __restrict__ val /*[17][LM_ROW]*/, uint16_t* __restrict__ ty /*[17][LM_ROW]*/) {
  const int lim = LM_LEVELS;
  uint32_t sfreq[MAX_ALPHA];
  uint16_t map[MAX_ALPHA];
  uint8_t c[MAX_ALPHA];
  // stable descending sort by weight (sort_by(|x,y| y.1.cmp(&x.1)), :69) — insertion sort is stable
  for (int i = 0; i < len; ++i) {
    uint32_t w = wt[i];
    int j = i;
    while (j > 0 && sfreq[j - 1] < w) { sfreq[j] = sfreq[j - 1]; map[j] = map[j - 1]; --j; }
    sfreq[j] = w;
    map[j] = (uint16_t)i;
  }
  uint32_t max_elem[LM_LEVELS], bb[LM_LEVELS], cur[LM_LEVELS];
  for (int j = 0; j < lim; ++j) { max_elem[j] = 0; bb[j] = 0; cur[j] = 0; }
  uint32_t excess = (1u << lim) - (uint32_t)len;
  const uint32_t half = 1u << (lim - 1);
  max_elem[lim - 1] = (uint32_t)len;
  for (int j = 0; j < lim; ++j) {
    if (excess >= half) { bb[j] = 1; excess -= half; }
    excess <<= 1;
    if (lim >= 2 + j) max_elem[lim - 2 - j] = max_elem[lim - 1 - j] / 2 + (uint32_t)len;
  }
  max_elem[0] = bb[0];
  for (int j = 1; j < lim; ++j)
    if (max_elem[j] > 2 * max_elem[j - 1] + bb[j]) max_elem[j] = 2 * max_elem[j - 1] + bb[j];
  for (int j = 0; j < lim; ++j) {
    if (max_elem[j] > (uint32_t)LM_ROW) return 1;
    for (uint32_t t = 0; t < max_elem[j]; ++t) { val[j * LM_ROW + t] = 0; ty[j * LM_ROW + t] = 0; }
  }
  for (int i = 0; i < len; ++i) c[i] = (uint8_t)lim;
  for (uint32_t t = 0; t < (uint32_t)len && t < max_elem[lim - 1]; ++t) {
    val[(lim - 1) * LM_ROW + t] = sfreq[t];
    ty[(lim - 1) * LM_ROW + t] = (uint16_t)t;
  }
  if (bb[lim - 1] == 1) { c[0] -= 1; cur[lim - 1] += 1; }

  int j = lim - 1;
  while (j > 0) {
    int i = 0;
    uint32_t next = cur[j];
    for (uint32_t t = 0; t < max_elem[j - 1]; ++t) {
      uint32_t weight =
          (next + 1 < max_elem[j]) ? bz_weight_add(val[j * LM_ROW + next], val[j * LM_ROW + next + 1]) : 0u;
      if (weight > sfreq[i]) {
        val[(j - 1) * LM_ROW + t] = weight;
        ty[(j - 1) * LM_ROW + t] = (uint16_t)len;
        next += 2;
      } else {
        val[(j - 1) * LM_ROW + t] = sfreq[i];
        ty[(j - 1) * LM_ROW + t] = (uint16_t)i;
        i += 1;
        if (i >= len) break;
      }
    }
    j -= 1;
    cur[j] = 0;
    if (bb[j] == 1) {
      // take_package(ty, c, cur, j) (:40-55), recursion unrolled onto an explicit stack
      int stk_lvl[LM_LEVELS + 1];
      int stk_ph[LM_LEVELS + 1];
      int sp = 0;
      stk_lvl[0] = j; stk_ph[0] = 0;
      while (sp >= 0) {
        const int lvl = stk_lvl[sp];
        if (stk_ph[sp] == 0) {
          if (lvl >= lim || cur[lvl] >= max_elem[lvl]) return 2;  // the reference would panic (index out of bounds)
          const uint32_t x = ty[lvl * LM_ROW + cur[lvl]];
          if (x == (uint32_t)len) {
            stk_ph[sp] = 1;
            ++sp; stk_lvl[sp] = lvl + 1; stk_ph[sp] = 0;
          } else {
            c[x] -= 1;
            cur[lvl] += 1;
            --sp;
          }
        } else if (stk_ph[sp] == 1) {
          stk_ph[sp] = 2;
          ++sp; stk_lvl[sp] = lvl + 1; stk_ph[sp] = 0;
        } else {
          cur[lvl] += 1;
          --sp;
        }
      }
    }
  }
  for (int i = 0; i < len; ++i) out_len[map[i]] = c[i];
  return 0;
}

// One thread per (block, table): lens[slot_out][t] = create_huffman(rfreq[t], 17).  Tables whose plain Huffman
// depth exceeds 17 are queued for k4_lm_fallback.
__global__ void __launch_bounds__(HB_NT) k4_build_tables(uint32_t nb, uint32_t* __restrict__ rfreq,
                                                         uint8_t* __restrict__ lens, int slot_out,
                                                         uint32_t* __restrict__ meta, uint32_t* __restrict__ lm_list,
                                                         uint32_t* __restrict__ lm_count) {
  const uint32_t id = blockIdx.x * HB_NT + threadIdx.x;
  if (id >= nb * MAX_GROUPS) return;
  const uint32_t b = id / MAX_GROUPS;
  const int t = (int)(id % MAX_GROUPS);
  const uint32_t* m = meta + (size_t)b * META;
  const int n = (int)m[0];
  if (t >= (int)m[1]) return;
  uint32_t* rf = rfreq + ((size_t)b * MAX_GROUPS + t) * MAX_ALPHA;
  uint32_t buf[2 * MAX_ALPHA];
  for (int i = 0; i < n; ++i) {
    buf[i] = (uint32_t)(n + i);
    buf[n + i] = max(1u, rf[i]) << 8;  // encoder.rs:642-645
  }
  // create_heap (:33-38): s = buf.len()>>1 = n
  for (int i = (n >> 1) - 1; i >= 0; --i) down_heap(buf, (uint32_t)i, (uint32_t)n);
  for (int i = n - 1; i >= 1; --i) {  // (:164-178)
    const uint32_t m1 = buf[0];
    buf[0] = buf[i];
    down_heap(buf, 0, (uint32_t)i);
    const uint32_t m2 = buf[0];
    buf[i] = bz_weight_add(buf[m1], buf[m2]);
    buf[0] = (uint32_t)i;
    buf[m1] = (uint32_t)i;
    buf[m2] = (uint32_t)i;
    down_heap(buf, 0, (uint32_t)i);
  }
  buf[1] = 0;  // (:180-183)
  for (int i = 2; i < n; ++i) buf[i] = buf[buf[i]] + 1;
  uint8_t* out = lens + lens_index(b, slot_out, t);
  bool over = false;
  for (int i = 0; i < n; ++i) {
    uint32_t l = buf[buf[i + n]] + 1;
    out[i] = (uint8_t)l;
    over |= l > (uint32_t)MAX_CODE_LEN;
  }
  if (over) {
    uint32_t k = atomicAdd(lm_count, 1u);
    lm_list[k] = id;  // rfreq row is kept for the fallback kernel, which clears it afterwards
  } else {
    for (int i = 0; i < n; ++i) rf[i] = 0;
  }
}

// Length-limited fallback for the queued tables: CTA s works on items s, s+nslots, ...  The package-merge lists
// (17 levels x <= 516 entries) live in shared memory — the algorithm is a serial chain of dependent reads, so it
// runs at shared-memory latency on one thread; parallelism comes from the tables in flight.
__global__ void __launch_bounds__(32) k4_lm_fallback(uint32_t nslots, uint32_t* __restrict__ rfreq,
                                                     uint8_t* __restrict__ lens, int slot_out,
                                                     uint32_t* __restrict__ meta, const uint32_t* __restrict__ lm_list,
                                                     const uint32_t* __restrict__ lm_count,
                                                     uint8_t* __restrict__ scratch) {
  extern __shared__ __align__(16) uint8_t lm_raw[];
  (void)scratch;
  if (threadIdx.x != 0) return;
  const uint32_t s = blockIdx.x;
  const uint32_t cnt = *lm_count;
  uint32_t* val = reinterpret_cast<uint32_t*>(lm_raw);
  uint16_t* ty = reinterpret_cast<uint16_t*>(lm_raw + (size_t)LM_LEVELS * LM_ROW * 4);
  for (uint32_t k = s; k < cnt; k += nslots) {
    const uint32_t id = lm_list[k];
    const uint32_t b = id / MAX_GROUPS;
    const int t = (int)(id % MAX_GROUPS);
    uint32_t* m = meta + (size_t)b * META;
    const int n = (int)m[0];
    uint32_t* rf = rfreq + ((size_t)b * MAX_GROUPS + t) * MAX_ALPHA;
    uint32_t wt[MAX_ALPHA];
    for (int i = 0; i < n; ++i) wt[i] = max(1u, rf[i]) << 8;
    int e = gen_code_lm_dev(wt, n, lens + lens_index(b, slot_out, t), val, ty);
    if (e) atomicOr(&m[6], (uint32_t)e);
    atomicAdd(&m[5], 1u);
    for (int i = 0; i < n; ++i) rf[i] = 0;
  }
}

// ---------------------------------------------------------------- canonical codes (huffman/mod.rs:22-67)
__global__ void k4_codes(uint32_t nb, const uint8_t* __restrict__ lens, const uint32_t* __restrict__ meta,
                         uint32_t* __restrict__ codes) {
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= nb * MAX_GROUPS) return;
  const uint32_t b = id / MAX_GROUPS;
  const int t = (int)(id % MAX_GROUPS);
  const uint32_t* m = meta + (size_t)b * META;
  const int n = (int)m[0];
  if (t >= (int)m[1]) return;
  const uint8_t* l = lens + lens_index(b, LENS_SLOTS - 1, t);
  uint32_t* cd = codes + ((size_t)b * MAX_GROUPS + t) * MAX_ALPHA;
  uint32_t cc = 0;
  int cl = 0;
  // stable sort by length == for each length ascending, symbols in index order
  for (int len = 1; len <= 32; ++len) {
    for (int s = 0; s < n; ++s) {
      if (l[s] == len) {
        uint32_t code = cc << (cl < len ? len - cl : 0);
        cl = len;
        cc = code + 1;
        cd[s] = code | ((uint32_t)len << 24);
      }
    }
  }
}

// ---------------------------------------------------------------- selector MTF + header size (encoder.rs:511-601)
// One CTA per block.  The selector MTF (mtf.rs:22-38 on a list of <= 6 tables) is chunked: a chunk's effect on the
// list is "the tables it touches, most recent first, then the others in their previous order", so every thread
// (1) runs its chunk from the identity list to get that summary, (2) thread 0 composes the summaries into the list
// at each chunk start, (3) every thread re-runs its chunk from the true list and writes the MTF values.
// Lists are packed 4 bits per entry.
constexpr int HS_NT = 320;
__device__ __forceinline__ uint32_t selmtf_step(uint32_t& list, uint32_t v) {  // returns the position of v
  uint32_t j = 0;
#pragma unroll
  for (int q = 1; q < MAX_GROUPS; ++q)
    if (((list >> (4 * q)) & 15u) == v) j = q;
  const uint32_t low = (1u << (4 * j)) - 1u;           // entries in front of v
  const uint32_t keep = ~((1u << (4 * (j + 1))) - 1u);  // entries behind v
  list = (list & keep) | ((list & low) << 4) | v;
  return j;
}

__global__ void __launch_bounds__(HS_NT) k4_header_size(uint32_t nb, const uint8_t* __restrict__ lens,
                                                        const uint8_t* __restrict__ sel, uint8_t* __restrict__ selmtf,
                                                        const uint32_t* __restrict__ inuse, uint32_t* __restrict__ meta) {
  __shared__ uint32_t s_sum[HS_NT];   // chunk summary: list after the chunk run from identity | touched mask << 24
  __shared__ uint32_t s_start[HS_NT];
  __shared__ uint32_t s_bits;
  const uint32_t b = blockIdx.x;
  uint32_t* m = meta + (size_t)b * META;
  const int alpha = (int)m[0], ng = (int)m[1];
  const uint32_t nsel = m[2];
  const uint8_t* sp = sel + (size_t)b * MAX_SELECTORS;
  uint8_t* so = selmtf + (size_t)b * MAX_SELECTORS;
  const uint32_t per = (nsel + HS_NT - 1) / HS_NT;
  const uint32_t lo = min(nsel, per * threadIdx.x), hi = min(nsel, lo + per);
  const uint32_t ident = 0x543210u;
  if (threadIdx.x == 0) s_bits = 0;
  {
    uint32_t list = ident, touched = 0;
    for (uint32_t i = lo; i < hi; ++i) {
      const uint32_t v = sp[i];
      selmtf_step(list, v);
      touched |= 1u << v;
    }
    s_sum[threadIdx.x] = list | (touched << 24);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t cur = ident;
    for (int c = 0; c < HS_NT; ++c) {
      s_start[c] = cur;
      const uint32_t sm = s_sum[c];
      const uint32_t touched = sm >> 24;
      const int k = __popc(touched);
      // new list = first k entries of the summary, then the untouched entries of cur in order
      uint32_t nl = sm & ((1u << (4 * k)) - 1u);
      int o = k;
      for (int q = 0; q < MAX_GROUPS; ++q) {
        const uint32_t v = (cur >> (4 * q)) & 15u;
        if (!((touched >> v) & 1u)) { nl |= v << (4 * o); ++o; }
      }
      cur = nl;
    }
  }
  __syncthreads();
  uint32_t bits = 0;
  {
    uint32_t list = s_start[threadIdx.x];
    for (uint32_t i = lo; i < hi; ++i) {
      const uint32_t j = selmtf_step(list, sp[i]);
      so[i] = (uint8_t)j;
      bits += j + 1;
    }
  }
  // coding tables: 5-bit start length, then per symbol (10|11)* 0
  for (int i = threadIdx.x; i < ng * alpha; i += HS_NT) {
    const int t = i / alpha, sy = i - t * alpha;
    const uint8_t* l = lens + lens_index(b, LENS_SLOTS - 1, t);
    const int d = sy ? (int)l[sy] - (int)l[sy - 1] : 0;
    bits += 1 + 2 * (uint32_t)(d < 0 ? -d : d) + (sy == 0 ? 5u : 0u);
  }
  if (threadIdx.x < 16) {
    const uint32_t wv = inuse[b * 8 + (threadIdx.x >> 1)];
    const uint32_t hh = (threadIdx.x & 1) ? (wv >> 16) : (wv & 0xFFFFu);
    if (hh) bits += 16;
  }
  if (threadIdx.x == 0) bits += 48 + 32 + 1 + 24 + 16 + 3 + 15;  // magic, crc, randomised, origPtr, range map, nGroups, nSelectors
#pragma unroll
  for (int dlt = 16; dlt > 0; dlt >>= 1) bits += __shfl_xor_sync(0xffffffffu, bits, dlt);
  if (lane_id() == 0) atomicAdd(&s_bits, bits);
  __syncthreads();
  if (threadIdx.x == 0) m[3] = s_bits;
}

// Per-block exclusive scan of the group bit lengths (in place) + data_bits.
constexpr int GS_NT = 1024;
__global__ void __launch_bounds__(GS_NT) k4_group_scan(uint32_t* __restrict__ gbits, uint32_t* __restrict__ meta) {
  __shared__ uint32_t ws[GS_NT / 32 + 1];
  const uint32_t b = blockIdx.x;
  uint32_t* m = meta + (size_t)b * META;
  const uint32_t nsel = m[2];
  uint32_t* g = gbits + (size_t)b * MAX_SELECTORS;
  const uint32_t per = (nsel + GS_NT - 1) / GS_NT;
  const uint32_t lo = min(nsel, per * threadIdx.x), hi = min(nsel, lo + per);
  uint32_t s = 0;
  for (uint32_t i = lo; i < hi; ++i) s += g[i];
  uint32_t total;
  uint32_t run = cta_excl_scan_add<GS_NT>(s, ws, &total);
  for (uint32_t i = lo; i < hi; ++i) {
    uint32_t v = g[i];
    g[i] = run;
    run += v;
  }
  if (threadIdx.x == 0) m[4] = total;
}

void launch_huffman(Launcher& L, const uint16_t* d_sym, const BlockDesc* d_desc, const uint32_t* d_mtf_count,
                    const uint32_t* d_freq, const uint32_t* d_inuse, uint32_t nb, uint32_t max_groups_per_block,
                    HuffBuffers& H) {
  uint32_t* lm_list = reinterpret_cast<uint32_t*>(H.lm_list);
  static PerDeviceOnce once_lm;
  once_lm.run([] {
    cudaFuncSetAttribute((const void*)k4_lm_fallback, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LM_BYTES);
  });
  L.launch("k4_init", k4_init, dim3((nb + 63) / 64), dim3(64), nb, d_mtf_count, d_freq, d_inuse, H.lens, H.meta,
           H.rfreq);
  const dim3 cgrid((max_groups_per_block + CS_NT - 1) / CS_NT, nb);
  const uint32_t ntab = nb * MAX_GROUPS;
  for (int it = 0; it < N_ITERS; ++it) {
    cudaMemsetAsync(H.lm_count, 0, sizeof(uint32_t), L.stream);
    L.launch("k4_cost_select", k4_cost_select, cgrid, dim3(CS_NT), d_sym, d_desc, d_mtf_count,
             (const uint8_t*)H.lens, it, (const uint32_t*)H.meta, H.sel, H.rfreq, H.gbits, 0);
    L.launch("k4_build_tables", k4_build_tables, dim3((ntab + HB_NT - 1) / HB_NT), dim3(HB_NT), nb, H.rfreq, H.lens,
             it + 1, H.meta, lm_list, H.lm_count);
    L.launch_smem("k4_lm_fallback", k4_lm_fallback, dim3(H.lm_slots), dim3(32), LM_BYTES, H.lm_slots, H.rfreq, H.lens,
                  it + 1, H.meta, (const uint32_t*)lm_list, (const uint32_t*)H.lm_count, H.lm_scratch);
  }
  L.launch("k4_codes", k4_codes, dim3((ntab + 63) / 64), dim3(64), nb, (const uint8_t*)H.lens,
           (const uint32_t*)H.meta, H.codes);
  L.launch("k4_header_size", k4_header_size, dim3(nb), dim3(HS_NT), nb, (const uint8_t*)H.lens,
           (const uint8_t*)H.sel, H.selmtf, d_inuse, H.meta);
  L.launch("k4_cost_select", k4_cost_select, cgrid, dim3(CS_NT), d_sym, d_desc, d_mtf_count, (const uint8_t*)H.lens,
           LENS_SLOTS - 1, (const uint32_t*)H.meta, H.sel, H.rfreq, H.gbits, 1);
  L.launch("k4_group_scan", k4_group_scan, dim3(nb), dim3(GS_NT), H.gbits, H.meta);
}

}  // namespace bzb
