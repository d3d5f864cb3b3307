// bz2_oracle.cpp — CPU restatement of chalharu/rust-compression's BZip2Encoder path.
//
// TEST INFRASTRUCTURE ONLY.  This file is the bit-exact *checker* for the CUDA
// path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load it.  Nothing under rust-compression_b200/
// links, imports or executes it; the product path has no CPU fallback.
//
// The Rust crate cannot be built in this image (no rustc/cargo), so this is a
// statement-by-statement restatement in C++17.  Every function cites the
// reference file:line it follows (paths relative to /root/reference/src).
// Pinning: tests/test_oracle.py checks it against every golden vector the
// reference's own tests hold for this path (bzip2/mod.rs:41-58 test_unit,
// suffix_array/sais.rs:294-556 test_bwt1-12, huffman/cano_huff_table.rs:238-294,
// huffman/encoder.rs:64-79, bitio/writer.rs:253-322) plus SURVEY.md App. B.
//
// Build: see oracle/Makefile  (g++ -O3 -shared -fPIC).

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <vector>

#include "bz_rand_table.h"

namespace {

using usize = size_t;
static const usize NONE = SIZE_MAX;  // usize::max_value()

// ---------------------------------------------------------------------------
// crc32.rs:58-72 (make_table_normal), :82-84 (update_normal), :129-131 (finish)
// CRC-32/BZIP2: poly 0x04C11DB7, MSB first, init 0xFFFFFFFF, final NOT.
// ---------------------------------------------------------------------------
struct Crc {
  uint32_t table[256];
  Crc() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t v = i << 24;
      for (int k = 0; k < 8; ++k) v = (v & 0x80000000u) ? (v << 1) ^ 0x04C11DB7u : (v << 1);
      table[i] = v;
    }
  }
  inline uint32_t update(uint32_t value, uint8_t byte) const {
    return table[((value >> 24) ^ byte) & 0xFF] ^ (value << 8);
  }
};
static const Crc g_crc;

// ---------------------------------------------------------------------------
// bitio/writer.rs:186-242 + bitio/direction/left.rs:17-62 + small_bit_vec.rs:
// the Left writer is a plain MSB-first concatenation of (value,len<=32) fields
// (only the low `len` bits of value survive the convert() shift); flush pads
// the last byte with zero bits.
// ---------------------------------------------------------------------------
struct BitSink {
  std::vector<uint8_t> bytes;
  uint64_t nbits = 0;
  uint64_t acc = 0;
  unsigned accn = 0;
  void put(uint32_t value, unsigned len) {
    if (!len) return;
    uint64_t v = len >= 32 ? (uint64_t)value : (uint64_t)(value & ((1u << len) - 1u));
    acc = (acc << len) | v;
    accn += len;
    nbits += len;
    while (accn >= 8) {
      bytes.push_back((uint8_t)(acc >> (accn - 8)));
      accn -= 8;
    }
    acc &= ((uint64_t)1 << accn) - 1;
  }
  void flush() {  // writer.rs:226-242: zero-pad to a byte boundary
    if (accn) { bytes.push_back((uint8_t)(acc << (8 - accn))); accn = 0; acc = 0; }
  }
};

// ---------------------------------------------------------------------------
// bitset.rs:21-83 BitArray (only get/set/len are needed by the suffix sorter)
// ---------------------------------------------------------------------------
struct BitArray {
  std::vector<uint64_t> data;
  usize len_ = 0;
  BitArray() {}
  explicit BitArray(usize len) : data((len + 63) >> 6, 0), len_(len) {}
  bool get(usize idx) const { return (data[idx >> 6] >> (idx & 63)) & 1u; }
  void set(usize idx, bool v) {
    uint64_t m = (uint64_t)1 << (idx & 63);
    if (v) data[idx >> 6] |= m; else data[idx >> 6] &= ~m;
  }
  usize len() const { return len_; }
};

// ---------------------------------------------------------------------------
// suffix_array/ls_type.rs:10-92  LSTypeArray::with_shift
// bitmap[i] = true  <=> position i is S-type.
// ---------------------------------------------------------------------------
struct LSTypeArray {
  BitArray bitmap, lms;
  template <class T>
  static LSTypeArray with_shift(const T* array, usize count, usize shift) {
    LSTypeArray r;
    r.bitmap = BitArray(count);
    usize start = shift == 0 ? count : shift;
    // ls_type.rs:21-31
    for (usize i = start; i-- > 1;) {  // i in (1..start).rev()
      bool b = r.bitmap.get(i);
      r.bitmap.set(i - 1, array[i] == array[i - 1] ? b : (array[i - 1] < array[i]));
    }
    // ls_type.rs:33-53
    if (shift != 0) {
      bool b = r.bitmap.get(0);
      r.bitmap.set(count - 1, array[0] == array[count - 1] ? b : (array[count - 1] < array[0]));
      for (usize i = count; i-- > shift + 1;) {  // i in (shift+1..count).rev()
        bool bb = r.bitmap.get(i);
        r.bitmap.set(i - 1, array[i] == array[i - 1] ? bb : (array[i - 1] < array[i]));
      }
    }
    // ls_type.rs:55-77
    r.lms = BitArray(count);
    if (shift == 0) {
      bool old = true;
      for (usize i = 0; i < count; ++i) {
        bool b = r.bitmap.get(i);
        r.lms.set(i, b && !old);
        old = b;
      }
    } else {
      bool old = r.bitmap.get(count - 1);
      for (usize i = 0; i < count; ++i) {
        bool b = r.bitmap.get(i);
        r.lms.set(i, i != shift && b && !old);
        old = b;
      }
    }
    return r;
  }
  bool get(usize i) const { return bitmap.get(i); }
  bool is_lms(usize i) const { return lms.get(i); }
  usize len() const { return bitmap.len(); }
};

// ---------------------------------------------------------------------------
// suffix_array/bucket.rs:8-85  BucketBuilder / Bucket
// ---------------------------------------------------------------------------
template <class T>
struct Bucket {
  std::vector<usize> data;
  const T* array;
  usize min;
  usize& operator[](usize idx) { return data[(usize)array[idx] - min]; }
};

template <class T>
struct BucketBuilder {
  std::vector<usize> data;
  const T* array;
  usize min;
  BucketBuilder(const T* a, usize count, usize mn, usize mx) : data(mx - mn + 2, 0), array(a), min(mn) {
    for (usize i = 0; i < count; ++i) data[(usize)a[i] - mn] += 1;  // bucket.rs:21-30
    usize sum = 0;
    for (auto& d : data) { usize v = d; d = sum; sum += v; }          // bucket.rs:32-37
  }
  Bucket<T> build(bool has_end) const {  // bucket.rs:41-53
    Bucket<T> b;
    b.data.assign(data.size() - 1, 0);
    for (usize i = 0; i + 1 < data.size(); ++i) b.data[i] = has_end ? data[i + 1] : data[i];
    b.array = array;
    b.min = min;
    return b;
  }
};

// ---------------------------------------------------------------------------
// suffix_array/sais.rs:12-68  array_rotate_for_non_sentinel_bwt  (literal)
// Returns NONE if `budget` inner steps are exceeded (the caller then uses the
// O(n) equivalent below; the reference itself has no budget — it is Θ(n²) on
// periodic blocks).
// ---------------------------------------------------------------------------
static usize least_rotation_literal(const uint8_t* array, usize count, usize* sarray, usize bucket_max,
                                    uint64_t budget) {
  usize n1 = 0, val = bucket_max + 1, prev_pos = 0;
  for (usize i = 0; i < count; ++i) {
    usize j = array[i];
    if (val > j) {
      sarray[0] = i; val = j; n1 = 1; prev_pos = i;
    } else if (val == j) {
      prev_pos += 1;
      if (prev_pos != i) { sarray[n1] = i; n1 += 1; }
    }
  }
  uint64_t steps = 0;
  for (usize i = 0; i < count; ++i) {
    usize n2 = 0;
    val = bucket_max + 1;
    steps += n1;
    if (steps > budget) return NONE;
    for (usize j = 0; j < n1; ++j) {
      usize k = sarray[j] + 1;
      if (k >= count) k -= count;
      usize l = array[k];
      if (val == l) { sarray[n2] = k; n2 += 1; }
      else if (val > l) { sarray[0] = k; val = l; n2 = 1; }
    }
    if (n2 == 1) return sarray[0] <= i ? sarray[0] + count - i - 1 : sarray[0] - i - 1;
    n1 = n2;
  }
  return sarray[0];
}

// O(n) equivalent of sais.rs:12-68: smallest index at which a lexicographically
// minimal rotation starts.  Booth's least-rotation gives
// one minimal start k; all minimal starts are k + j*p for the cyclic period p,
// so the smallest is k mod p.  Cross-checked against the literal version in
// tests/test_oracle.py.
static usize least_rotation_fast(const uint8_t* s, usize n) {
  if (n == 0) return 0;
  // Booth
  std::vector<long> f(2 * n, -1);
  usize k = 0;
  for (usize j = 1; j < 2 * n; ++j) {
    uint8_t sj = s[j % n];
    long i = f[j - k - 1];
    while (i != -1 && sj != s[(k + i + 1) % n]) {
      if (sj < s[(k + i + 1) % n]) k = j - i - 1;
      i = f[i];
    }
    if (sj != s[(k + i + 1) % n]) {  // i == -1
      if (sj < s[k % n]) k = j;
      f[j - k] = -1;
    } else {
      f[j - k] = i + 1;
    }
  }
  k %= n;
  // cyclic period via KMP failure function
  std::vector<usize> fail(n + 1, 0);
  usize q = 0;
  for (usize i = 1; i < n; ++i) {
    while (q > 0 && s[i] != s[q]) q = fail[q];
    if (s[i] == s[q]) ++q;
    fail[i + 1] = q;
  }
  usize p = n - fail[n];
  if (n % p != 0) p = n;
  return k % p;
}

// ---------------------------------------------------------------------------
// suffix_array/sais.rs:76-121  induce_sa
// ---------------------------------------------------------------------------
template <class T>
static void induce_sa(const BucketBuilder<T>& bb, const LSTypeArray& ty, usize* sa, usize shift) {
  usize n = ty.len();
  {  // compute SAl  (sais.rs:85-104)
    Bucket<T> bucket = bb.build(false);
    usize k = (shift == 0 ? n : shift) - 1;
    usize bk = bucket[k];
    sa[bk] = k;
    bucket[k] = bk + 1;
    for (usize i = 0; i < n; ++i) {
      usize j = sa[i];
      if (j != NONE && j != shift) {
        j = (j == 0 ? n : j) - 1;
        if (!ty.get(j)) {
          usize bj = bucket[j];
          sa[bj] = j;
          bucket[j] = bj + 1;
        }
      }
    }
  }
  {  // compute SAs  (sais.rs:106-120)
    Bucket<T> bucket = bb.build(true);
    for (usize i = n; i-- > 0;) {
      usize j = sa[i];
      if (j != NONE && j != shift) {
        j = (j == 0 ? n : j) - 1;
        if (ty.get(j)) {
          usize bj = bucket[j] - 1;
          bucket[j] = bj;
          sa[bj] = j;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// suffix_array/sais.rs:127-264  sa_is  (cyclic, sentinel-free SA-IS)
// ---------------------------------------------------------------------------
template <class T>
static void sa_is(const T* array, usize count, usize* sa, usize bucket_min, usize bucket_max, usize shift) {
  LSTypeArray ty = LSTypeArray::with_shift(array, count, shift);

  // stage 1 (sais.rs:139-154)
  BucketBuilder<T> bb(array, count, bucket_min, bucket_max);
  {
    Bucket<T> bucket = bb.build(true);
    for (usize i = 0; i < count; ++i) sa[i] = NONE;
    auto place = [&](usize i) {
      if (ty.is_lms(i)) {
        usize bi = bucket[i] - 1;
        bucket[i] = bi;
        sa[bi] = i;
      }
    };
    for (usize i = shift + 1; i < count; ++i) place(i);
    for (usize i = 0; i < shift; ++i) place(i);
  }
  induce_sa(bb, ty, sa, shift);

  // compact sorted LMS substrings (sais.rs:156-165)
  usize n1 = 0;
  for (usize i = 0; i < count; ++i) {
    if (sa[i] != NONE && ty.is_lms(sa[i])) {
      sa[n1] = sa[i];
      n1 += 1;
    }
  }

  // name the substrings (sais.rs:167-207)
  for (usize i = n1; i < count; ++i) sa[i] = NONE;
  usize name = 0;
  usize prev_store = NONE;
  for (usize i = 0; i < n1; ++i) {
    usize prev = prev_store;
    usize pos = sa[i];
    usize now = pos;
    bool diff = false;
    for (;;) {
      if (prev == NONE || now == shift || prev == shift || array[now] != array[prev] ||
          ty.get(now) != ty.get(prev)) {
        diff = true;
        break;
      } else if (now != pos && (ty.is_lms(now) || ty.is_lms(prev))) {
        break;
      }
      now = now == count - 1 ? 0 : now + 1;
      prev = prev == count - 1 ? 0 : prev + 1;
    }
    if (diff) {
      name += 1;
      prev_store = pos;
    }
    pos = (pos > shift ? pos - shift : pos + count - shift) >> 1;
    sa[n1 + pos] = name - 1;
  }
  {  // sais.rs:208-216
    usize j = count - 1;
    for (usize i = count; i-- > n1;) {  // i in (n1..=count-1).rev()
      if (sa[i] != NONE) {
        sa[j] = sa[i];
        j -= 1;  // may wrap after the last store; never read again in that case
      }
    }
  }

  // stage 2 (sais.rs:218-231)
  usize* s1 = sa + (count - n1);
  if (name < n1) {
    sa_is<usize>(s1, n1, sa, 0, name - 1, 0);
  } else {
    for (usize i = 0; i < n1; ++i) sa[s1[i]] = i;
  }

  // stage 3 (sais.rs:233-263)
  Bucket<T> bucket2 = bb.build(true);
  {
    usize j = 0;
    for (usize i = shift + 1; i < count; ++i) if (ty.is_lms(i)) s1[j++] = i;
    for (usize i = 0; i < shift; ++i) if (ty.is_lms(i)) s1[j++] = i;
  }
  for (usize i = 0; i < n1; ++i) sa[i] = s1[sa[i]];
  for (usize i = n1; i < count; ++i) sa[i] = NONE;
  for (usize i = n1; i-- > 0;) {
    usize j = sa[i];
    sa[i] = NONE;
    usize b2j = bucket2[j] - 1;
    bucket2[j] = b2j;
    sa[b2j] = j;
  }
  induce_sa(bb, ty, sa, shift);
}

// suffix_array/sais.rs:266-272  bwt()
// literal_budget: inner-step budget for the literal least-rotation pre-pass
// before switching to the O(n) equivalent (0 = always fast path).
static std::vector<usize> bwt(const uint8_t* array, usize count, usize max_value, usize* shift_out,
                              uint64_t literal_budget = 20000000ull) {
  std::vector<usize> sa(count, 0);
  if (count == 0) { if (shift_out) *shift_out = 0; return sa; }
  usize shift = NONE;
  if (literal_budget) shift = least_rotation_literal(array, count, sa.data(), max_value, literal_budget);
  if (shift == NONE) shift = least_rotation_fast(array, count);
  sa_is<uint8_t>(array, count, sa.data(), 0, max_value, shift);
  if (shift_out) *shift_out = shift;
  return sa;
}

// ---------------------------------------------------------------------------
// huffman/cano_huff_table.rs:14-31 down_heap, :33-38 create_heap
// ---------------------------------------------------------------------------
static void down_heap(std::vector<usize>& buf, usize n, usize len) {
  usize tmp = buf[n];
  usize leaf = (n << 1) + 1;
  while (leaf < len) {
    if (leaf + 1 < len && buf[buf[leaf]] > buf[buf[leaf + 1]]) leaf += 1;
    if (buf[tmp] < buf[buf[leaf]]) break;
    buf[n] = buf[leaf];
    n = leaf;
    leaf = (n << 1) + 1;
  }
  buf[n] = tmp;
}
static void create_heap(std::vector<usize>& buf) {
  usize s = buf.size() >> 1;
  for (usize i = s >> 1; i-- > 0;) down_heap(buf, i, s);
}

using AddFn = std::function<usize(usize, usize)>;

// cano_huff_table.rs:40-55 take_package
static void take_package(std::vector<std::vector<usize>>& ty, std::vector<usize>& len, std::vector<usize>& cur,
                         usize i) {
  usize x = ty[i][cur[i]];
  if (x == len.size()) {
    take_package(ty, len, cur, i + 1);
    take_package(ty, len, cur, i + 1);
  } else {
    len[x] -= 1;
  }
  cur[i] += 1;
}

// cano_huff_table.rs:58-151 gen_code_lm (reverse package merge)
static std::vector<uint8_t> gen_code_lm(const std::vector<usize>& freq, usize lim, const AddFn& add) {
  usize len = freq.size();
  std::vector<std::pair<usize, usize>> freqmap(len);
  for (usize i = 0; i < len; ++i) freqmap[i] = {i, freq[i]};
  std::stable_sort(freqmap.begin(), freqmap.end(),
                   [](const auto& x, const auto& y) { return y.second < x.second; });  // sort_by(|x,y| y.1.cmp(&x.1))
  std::vector<usize> map(len), sfreq(len);
  for (usize i = 0; i < len; ++i) { map[i] = freqmap[i].first; sfreq[i] = freqmap[i].second; }

  std::vector<usize> max_elem(lim, 0), b(lim, 0);
  usize excess = ((usize)1 << lim) - len;
  usize half = (usize)1 << (lim - 1);
  max_elem[lim - 1] = len;
  for (usize j = 0; j < lim; ++j) {
    if (excess >= half) { b[j] = 1; excess -= half; }
    excess <<= 1;
    if (lim >= 2 + j) max_elem[lim - 2 - j] = max_elem[lim - 1 - j] / 2 + len;
  }
  max_elem[0] = b[0];
  for (usize j = 1; j < lim; ++j)
    if (max_elem[j] > 2 * max_elem[j - 1] + b[j]) max_elem[j] = 2 * max_elem[j - 1] + b[j];

  std::vector<std::vector<usize>> val(lim), ty(lim);
  for (usize i = 0; i < lim; ++i) { val[i].assign(max_elem[i], 0); ty[i].assign(max_elem[i], 0); }
  std::vector<usize> c(len, lim);

  for (usize t = 0; t < len && t < max_elem[lim - 1]; ++t) {
    val[lim - 1][t] = sfreq[t];
    ty[lim - 1][t] = t;
  }

  std::vector<usize> cur(lim, 0);
  if (b[lim - 1] == 1) { c[0] -= 1; cur[lim - 1] += 1; }

  usize j = lim - 1;
  while (j > 0) {
    usize i = 0;
    usize next = cur[j];
    for (usize t = 0; t < max_elem[j - 1]; ++t) {
      usize weight = (next + 1 < max_elem[j]) ? add(val[j][next], val[j][next + 1]) : 0;
      if (weight > sfreq[i]) {
        val[j - 1][t] = weight;
        ty[j - 1][t] = len;
        next += 2;
      } else {
        val[j - 1][t] = sfreq[i];
        ty[j - 1][t] = i;
        i += 1;
        if (i >= len) break;
      }
    }
    j -= 1;
    cur[j] = 0;
    if (b[j] == 1) take_package(ty, c, cur, j);
  }

  std::vector<uint8_t> r(len);
  for (usize i = 0; i < len; ++i) r[map[i]] = (uint8_t)c[i];  // zip(map) + sort by original index
  return r;
}

// cano_huff_table.rs:153-196 gen_code
static std::vector<uint8_t> gen_code(const std::vector<usize>& freq, usize lim, const AddFn& add, bool* used_lm) {
  usize n = freq.size();
  if (used_lm) *used_lm = false;
  if (n == 1) return std::vector<uint8_t>{1};
  std::vector<usize> buf(2 * n);
  for (usize i = 0; i < n; ++i) buf[i] = n + i;
  for (usize i = 0; i < n; ++i) buf[n + i] = freq[i];
  create_heap(buf);
  for (usize i = n; i-- > 1;) {  // i in (1..n).rev()
    usize m1 = buf[0];
    buf[0] = buf[i];
    down_heap(buf, 0, i);
    usize m2 = buf[0];
    buf[i] = add(buf[m1], buf[m2]);
    buf[0] = i;
    buf[m1] = i;
    buf[m2] = i;
    down_heap(buf, 0, i);
  }
  buf[1] = 0;
  for (usize i = 2; i < n; ++i) buf[i] = buf[buf[i]] + 1;
  std::vector<uint8_t> ret(n);
  bool over = false;
  for (usize i = 0; i < n; ++i) {
    ret[i] = (uint8_t)(buf[buf[i + n]] + 1);
    if ((usize)ret[i] > lim) over = true;
  }
  if (over) {
    if (used_lm) *used_lm = true;
    return gen_code_lm(freq, lim, add);
  }
  return ret;
}

// cano_huff_table.rs:198-225 make_tab_with_fn
static std::vector<uint8_t> make_tab_with_fn(const std::vector<usize>& freq, usize lim, const AddFn& add,
                                             bool* used_lm = nullptr) {
  std::vector<usize> s, l;
  for (usize i = 0; i < freq.size(); ++i) if (freq[i] != 0) { s.push_back(i); l.push_back(freq[i]); }
  std::vector<uint8_t> out;
  if (s.empty()) return out;
  std::vector<uint8_t> g = gen_code(l, lim, add, used_lm);
  usize c = 0;
  for (usize k = 0; k < s.size(); ++k) {
    for (; c < s[k]; ++c) out.push_back(0);
    out.push_back(g[k]);
    c = s[k] + 1;
  }
  return out;
}

// bzip2/encoder.rs:641-651 create_huffman
static usize bz_weight_add(usize x, usize y) {
  return ((x & 0xFFFFFF00u) + (y & 0xFFFFFF00u)) | (1 + std::max(x & 0xFF, y & 0xFF));
}
static std::vector<uint8_t> create_huffman(const std::vector<usize>& freq, usize lim, bool* used_lm = nullptr) {
  std::vector<usize> weight(freq.size());
  for (usize i = 0; i < freq.size(); ++i) weight[i] = std::max<usize>(1, freq[i]) << 8;
  return make_tab_with_fn(weight, lim, bz_weight_add, used_lm);
}

// ---------------------------------------------------------------------------
// huffman/mod.rs:22-67 create_huffman_table (Left => not reversed) with
// bucket_sort.rs:43-75: symbols stably sorted by length; code = prev << dlen.
// Returns code[sym] (valid where len != 0).
// ---------------------------------------------------------------------------
static std::vector<uint32_t> canonical_codes(const std::vector<uint8_t>& symb_len) {
  std::vector<uint32_t> code(symb_len.size(), 0);
  std::vector<std::pair<usize, uint8_t>> symbs;
  for (usize i = 0; i < symb_len.size(); ++i) if (symb_len[i] != 0) symbs.push_back({i, symb_len[i]});
  std::stable_sort(symbs.begin(), symbs.end(), [](const auto& a, const auto& b) { return a.second < b.second; });
  uint8_t cl = 0;
  uint32_t cc = 0;
  for (auto& sl : symbs) {
    uint32_t cd = cc << (cl < sl.second ? sl.second - cl : 0);
    cl = sl.second;
    cc = cd + 1;
    code[sl.first] = cd;
  }
  return code;
}

// ---------------------------------------------------------------------------
// bzip2/mtf.rs:12-39 MtfPosition
// ---------------------------------------------------------------------------
struct MtfPosition {
  std::vector<usize> data;
  explicit MtfPosition(usize count) : data(count) { for (usize i = 0; i < count; ++i) data[i] = i; }
  usize pop(usize value) {
    if (value == data[0]) return 0;
    usize t = data[0];
    data[0] = value;
    for (usize i = 1; i < data.size(); ++i) {
      std::swap(data[i], t);
      if (t == value) return i;
    }
    return NONE;  // unreachable!()
  }
};

// ---------------------------------------------------------------------------
// Per-block stage dump (test instrumentation; not part of the reference).
// ---------------------------------------------------------------------------
struct BlockDump {
  uint64_t in_start = 0, in_end = 0;   // input byte range whose pieces are in this block
  std::vector<uint8_t> rle;            // block_buf (encoder.rs:699-716)
  uint32_t crc = 0;
  uint32_t in_use[8] = {0};            // bit s of word s>>5 (LSB first)
  uint64_t shift = 0;
  uint32_t orig_ptr = 0;
  std::vector<uint32_t> sa;            // rotation starts in sorted order (sais.rs:266)
  std::vector<uint8_t> last;           // last column
  std::vector<uint16_t> mtf;           // mtf_buffer[..mtf_count] (RUNA/RUNB/sym+1/EOB)
  std::vector<uint32_t> freq;          // mtf_freq[alpha]
  uint32_t alpha = 0, ngroups = 0, nselectors = 0;
  std::vector<uint8_t> selectors[4];   // selector[] after each of the 4 passes (libbz2 table ids)
  std::vector<uint8_t> lens[5];        // [0] initial, [k] after pass k; layout [table 0..ng-1][alpha]
  uint32_t lm_used = 0;                // how many create_huffman calls took the package-merge path
  uint64_t bit_start = 0, bit_end = 0; // position of this block's section in the output bit stream
};

// ---------------------------------------------------------------------------
// bzip2/encoder.rs:162-740 EncoderInner
// ---------------------------------------------------------------------------
struct EncoderInner {
  static constexpr usize BZ_G_SIZE = 50;       // bzip2/mod.rs:20
  static constexpr usize BZ_N_ITERS = 4;       // encoder.rs:294
  static constexpr uint8_t BZ_LESSER_ICOST = 0, BZ_GREATER_ICOST = 15;  // :297-298

  std::vector<uint8_t> block_buf;
  bool finished = false;
  usize block_size_100k;
  usize block_max_len;
  uint32_t combined_crc = 0;
  usize block_no = 1;
  uint32_t block_crc = 0xFFFFFFFFu;
  uint8_t rle_buffer = 0;
  usize rle_count = 0;
  bool in_use[256];
  std::vector<uint16_t> mtf_buffer;

  BitSink* sink;
  std::vector<BlockDump>* dumps;   // optional
  bool keep_sa;
  uint64_t in_pos = 0;             // number of input bytes consumed by write_rle so far
  uint64_t blk_in_start = 0;
  // FIXTURE GENERATOR ONLY (not reference behaviour: the reference encoder always writes a 0 bit, encoder.rs:273):
  // emit blocks the way bzip2 <= 0.9.0 did when its sorter gave up — block bytes XOR-ed with the BZ2_rNums mask
  // before the BWT, in-use map taken from the masked bytes, randomised bit = 1 — so that the decoders' un-randomise
  // path (decoder.rs:27-115,478-480,537-539) has valid streams to decode.  libbz2 decodes these streams too.
  bool fixture_randomise = false;

  EncoderInner(usize level, BitSink* s, std::vector<BlockDump>* d, bool ksa)
      : block_size_100k(level), block_max_len(level * 100000 - 19), sink(s), dumps(d), keep_sa(ksa) {
    block_buf.reserve(level * 100000);
    mtf_buffer.assign(level * 100000 + 1, 0);
    std::fill(in_use, in_use + 256, false);
  }

  void prepare_new_block() {  // encoder.rs:178-183
    block_no += 1;
    block_crc = 0xFFFFFFFFu;
    block_buf.clear();
    std::fill(in_use, in_use + 256, false);
  }

  void write(uint32_t v, unsigned len) { sink->put(v, len); }
  void write_u8(uint8_t v) { write(v, 8); }

  // encoder.rs:699-716
  void write_rle() {
    for (usize i = 0; i < rle_count; ++i) block_crc = g_crc.update(block_crc, rle_buffer);
    in_pos += rle_count;
    usize ret_count = std::min<usize>(rle_count, 4);
    for (usize i = 0; i < ret_count; ++i) {
      in_use[rle_buffer] = true;
      block_buf.push_back(rle_buffer);
    }
    if (ret_count == 4) {
      uint8_t v = (uint8_t)(rle_count - 4);
      in_use[v] = true;
      block_buf.push_back(v);
    }
  }

  // encoder.rs:671-697
  void next(uint8_t buf) {
    if (rle_count == 0) { rle_buffer = buf; rle_count = 1; return; }
    if (rle_buffer == buf && rle_count < 255) { rle_count += 1; return; }
    write_rle();
    rle_count = 1;
    rle_buffer = buf;
    if (block_buf.size() >= block_max_len) write_block(false);
  }

  // encoder.rs:729-739
  void finish() {
    if (!finished) { finished = true; write_block(true); }
  }

  // encoder.rs:224-291
  void write_block(bool is_final) {
    if (is_final) { write_rle(); rle_count = 0; }
    usize nblock = block_buf.size();
    uint32_t bcrc = ~block_crc;
    combined_crc = ((combined_crc << 1) | (combined_crc >> 31)) ^ bcrc;
    if (block_no == 1) {
      write_u8(0x42); write_u8(0x5a); write_u8(0x68); write_u8((uint8_t)(0x30 + block_size_100k));
    }
    if (nblock > 0) {
      BlockDump* d = nullptr;
      if (dumps) {
        dumps->emplace_back();
        d = &dumps->back();
        d->in_start = blk_in_start;
        d->in_end = in_pos;
        d->rle = block_buf;
        d->crc = bcrc;
        for (int s = 0; s < 256; ++s) if (in_use[s]) d->in_use[s >> 5] |= 1u << (s & 31);
        d->bit_start = sink->nbits;
      }
      blk_in_start = in_pos;
      write_u8(0x31); write_u8(0x41); write_u8(0x59); write_u8(0x26); write_u8(0x53); write_u8(0x59);
      write(bcrc, 32);
      if (fixture_randomise) {
        usize n2go = 0, tpos = 0;
        std::fill(in_use, in_use + 256, false);
        for (usize i = 0; i < block_buf.size(); ++i) {
          if (n2go == 0) { n2go = BZ_RAND_NUMS[tpos]; tpos = (tpos + 1) % 512; }
          n2go -= 1;
          if (n2go == 1) block_buf[i] ^= 1;
          in_use[block_buf[i]] = true;
        }
        if (d) {
          d->rle = block_buf;
          std::fill(d->in_use, d->in_use + 8, 0u);
          for (int s2 = 0; s2 < 256; ++s2) if (in_use[s2]) d->in_use[s2 >> 5] |= 1u << (s2 & 31);
        }
      }
      write(fixture_randomise ? 1 : 0, 1);
      write_blockdata(d);
      if (d) d->bit_end = sink->nbits;
      prepare_new_block();
    }
    if (is_final) {
      write_u8(0x17); write_u8(0x72); write_u8(0x45); write_u8(0x38); write_u8(0x50); write_u8(0x90);
      write(combined_crc, 32);
    }
  }

  // encoder.rs:653-669
  void zle_write(usize zero_count, std::vector<uint32_t>& mtf_freq, usize& mtf_count) {
    if (zero_count != 0) {
      zero_count += 1;
      while (zero_count > 1) {
        uint16_t run = (uint16_t)(zero_count & 1);
        mtf_buffer[mtf_count] = run;
        mtf_count += 1;
        mtf_freq[run] += 1;
        zero_count >>= 1;
      }
    }
  }

  // encoder.rs:300-639
  void write_blockdata(BlockDump* d) {
    usize in_use_count = 0;
    uint8_t unseq2seq[256] = {0};
    for (int i = 0; i < 256; ++i) if (in_use[i]) { unseq2seq[i] = (uint8_t)in_use_count; in_use_count += 1; }
    usize eob = in_use_count + 1;

    MtfPosition mtf_table(in_use_count);
    usize zero_count = 0;
    std::vector<uint32_t> mtf_freq(in_use_count + 2, 0);
    usize mtf_count = 0;

    usize shift = 0;
    std::vector<usize> sa = bwt(block_buf.data(), block_buf.size(), 255, &shift);
    if (d) {
      d->shift = shift;
      d->last.resize(sa.size());
      if (keep_sa) d->sa.assign(sa.begin(), sa.end());
    }
    for (usize i = 0; i < sa.size(); ++i) {  // encoder.rs:324-353
      usize s = sa[i];
      usize j;
      if (s == 0) {
        write((uint32_t)i, 24);
        if (d) d->orig_ptr = (uint32_t)i;
        j = block_buf.size() - 1;
      } else {
        j = s - 1;
      }
      if (d) d->last[i] = block_buf[j];
      usize val = unseq2seq[block_buf[j]];
      uint16_t c = (uint16_t)(mtf_table.pop(val) + 1);
      if (c == 1) {
        zero_count += 1;
      } else {
        zle_write(zero_count, mtf_freq, mtf_count);
        zero_count = 0;
        mtf_buffer[mtf_count] = c;
        mtf_count += 1;
        mtf_freq[c] += 1;
      }
    }
    zle_write(zero_count, mtf_freq, mtf_count);
    mtf_buffer[mtf_count] = (uint16_t)eob;
    mtf_count += 1;
    mtf_freq[eob] += 1;

    usize alpha_size = in_use_count + 2;

    // encoder.rs:369-376
    usize group_num = mtf_count < 200 ? 2 : mtf_count < 600 ? 3 : mtf_count < 1200 ? 4 : mtf_count < 2400 ? 5 : 6;

    // encoder.rs:378-426 — `len` is held in REVERSE table order: len[0] is table group_num-1.
    std::vector<std::vector<uint8_t>> len;
    {
      uint32_t rem = (uint32_t)mtf_count;
      long gs_prev = 0;
      for (usize n_part = group_num; n_part >= 1; --n_part) {
        uint32_t t_freq = rem / (uint32_t)n_part;
        long ge = gs_prev - 1;
        uint32_t a_freq = 0;
        while (a_freq < t_freq && ge < (long)alpha_size - 1) {
          ge += 1;
          a_freq += mtf_freq[ge];
        }
        if (ge > gs_prev && n_part != group_num && n_part != 1 && (((group_num - n_part) & 1) == 1)) {
          a_freq -= mtf_freq[ge];
          ge -= 1;
        }
        long gs = gs_prev;
        rem -= a_freq;
        gs_prev = ge + 1;
        std::vector<uint8_t> l(alpha_size);
        for (long i = 0; i < (long)alpha_size; ++i) l[i] = (i >= gs && i <= ge) ? BZ_LESSER_ICOST : BZ_GREATER_ICOST;
        len.push_back(std::move(l));
      }
    }
    auto dump_lens = [&](int slot) {
      if (!d) return;
      d->lens[slot].clear();
      for (usize t = 0; t < group_num; ++t) {  // libbz2 order: table t is len[group_num-1-t]
        const auto& l = len[group_num - 1 - t];
        d->lens[slot].insert(d->lens[slot].end(), l.begin(), l.end());
      }
    };
    dump_lens(0);

    usize n_selectors = 0;
    std::vector<usize> selector(2 + 900000 / BZ_G_SIZE, 0);
    // encoder.rs:433-509
    for (usize iter = 0; iter < BZ_N_ITERS; ++iter) {
      std::vector<std::vector<usize>> rfreq(group_num, std::vector<usize>(alpha_size, 0));
      n_selectors = 0;
      usize gs = 0;
      while (gs < mtf_count) {
        usize ge = std::min(gs + BZ_G_SIZE, mtf_count);
        // len.iter().rev() => libbz2 table 0..group_num-1; min_by keeps the FIRST minimum
        usize bt = 0;
        uint16_t bc = 0;
        for (usize t = 0; t < group_num; ++t) {
          const auto& li = len[group_num - 1 - t];
          uint16_t sum = 0;
          for (usize k = gs; k < ge; ++k) sum = (uint16_t)(sum + li[mtf_buffer[k]]);
          if (t == 0 || sum < bc) { bt = t; bc = sum; }
        }
        selector[n_selectors] = bt;
        n_selectors += 1;
        for (usize k = gs; k < ge; ++k) rfreq[bt][mtf_buffer[k]] += 1;
        gs = ge;
      }
      // encoder.rs:504-508: len = rfreq.iter().rev().map(create_huffman)
      len.clear();
      for (usize t = group_num; t-- > 0;) {
        bool lm = false;
        len.push_back(create_huffman(rfreq[t], 17, &lm));
        if (d && lm) d->lm_used += 1;
      }
      if (d) {
        d->selectors[iter].resize(n_selectors);
        for (usize i = 0; i < n_selectors; ++i) d->selectors[iter][i] = (uint8_t)selector[i];
      }
      dump_lens((int)iter + 1);
    }

    // encoder.rs:511-517 selector MTF
    MtfPosition selector_mtf_tab(group_num);
    std::vector<usize> selector_mtf(n_selectors);
    for (usize i = 0; i < n_selectors; ++i) selector_mtf[i] = selector_mtf_tab.pop(selector[i]);

    // encoder.rs:519-524 codes (libbz2 order)
    std::vector<std::vector<uint32_t>> code(group_num);
    for (usize t = 0; t < group_num; ++t) code[t] = canonical_codes(len[group_num - 1 - t]);

    // encoder.rs:527-565 mapping table (bitset.rs:186-199: 16-bit group k = symbols 16k..16k+15)
    {
      bool in_use16[16];
      for (int k = 0; k < 16; ++k) {
        in_use16[k] = false;
        for (int j = 0; j < 16; ++j) if (in_use[k * 16 + j]) in_use16[k] = true;
      }
      uint32_t m = 0;
      for (int k = 0; k < 16; ++k) m = (m << 1) + (in_use16[k] ? 1 : 0);
      write(m, 16);
      for (int k = 0; k < 16; ++k)
        if (in_use16[k])
          for (int j = 0; j < 16; ++j) write(in_use[k * 16 + j] ? 1 : 0, 1);
    }

    // encoder.rs:567-574 selectors
    write((uint32_t)group_num, 3);
    write((uint32_t)n_selectors, 15);
    for (usize s : selector_mtf) write((1u << (s + 1)) - 2, (unsigned)(s + 1));

    // encoder.rs:583-601 coding tables
    for (usize t = 0; t < group_num; ++t) {
      const auto& l = len[group_num - 1 - t];
      uint8_t curr = l[0];
      write(curr, 5);
      for (uint8_t li : l) {
        while (curr < li) { write(2, 2); curr += 1; }
        while (curr > li) { write(3, 2); curr -= 1; }
        write(0, 1);
      }
    }

    // encoder.rs:609-629 block data
    {
      usize sel_ctr = 0, gs = 0;
      while (gs < mtf_count) {
        usize ge = std::min(gs + BZ_G_SIZE, mtf_count);
        usize t = selector[sel_ctr];
        const auto& l = len[group_num - 1 - t];
        for (usize i = gs; i < ge; ++i) {
          uint16_t b = mtf_buffer[i];
          write(code[t][b], l[b]);
        }
        gs = ge;
        sel_ctr += 1;
      }
    }

    if (d) {
      d->mtf.assign(mtf_buffer.begin(), mtf_buffer.begin() + mtf_count);
      d->freq = mtf_freq;
      d->alpha = (uint32_t)alpha_size;
      d->ngroups = (uint32_t)group_num;
      d->nselectors = (uint32_t)n_selectors;
    }
  }
};

struct Run {
  BitSink sink;
  std::vector<BlockDump> dumps;
};

// BZip2Encoder::new + encode(.., Action::Finish) (encoder.rs:58-158): feed every
// byte, finish, zero-pad to a byte (bitio/writer.rs:226-242).
static void run_encoder(int level, const uint8_t* in, size_t n, BitSink& sink, std::vector<BlockDump>* dumps,
                        bool keep_sa, bool fixture_randomise = false) {
  EncoderInner enc((usize)level, &sink, dumps, keep_sa);
  enc.fixture_randomise = fixture_randomise;
  for (size_t i = 0; i < n; ++i) enc.next(in[i]);
  enc.finish();
  sink.flush();
}

}  // namespace

// ===========================================================================
// C ABI for tests / bench (ctypes)
// ===========================================================================
extern "C" {

// Whole-stream compress. Returns number of bytes written, or -(needed) if cap is too small, -1 on bad level.
long long orc_compress(int level, const uint8_t* in, size_t n, uint8_t* out, size_t cap) {
  if (level < 1 || level > 9) return -1;  // encoder.rs:59-61 panics "invalid level"
  BitSink sink;
  run_encoder(level, in, n, sink, nullptr, false);
  if (sink.bytes.size() > cap) return -(long long)sink.bytes.size();
  memcpy(out, sink.bytes.data(), sink.bytes.size());
  return (long long)sink.bytes.size();
}

// Fixture generator (see EncoderInner::fixture_randomise): the stream bzip2 <= 0.9.0 would have written with every
// block randomised.  NOT something the reference encoder can produce.
long long orc_compress_randomised(int level, const uint8_t* in, size_t n, uint8_t* out, size_t cap) {
  if (level < 1 || level > 9) return -1;
  BitSink sink;
  run_encoder(level, in, n, sink, nullptr, false, true);
  if (sink.bytes.size() > cap) return -(long long)sink.bytes.size();
  memcpy(out, sink.bytes.data(), sink.bytes.size());
  return (long long)sink.bytes.size();
}

// Staged run: keeps all per-block dumps.
void* orc_run(int level, const uint8_t* in, size_t n, int keep_sa) {
  if (level < 1 || level > 9) return nullptr;
  Run* r = new Run();
  run_encoder(level, in, n, r->sink, &r->dumps, keep_sa != 0);
  return r;
}
void orc_free(void* h) { delete (Run*)h; }
size_t orc_out_size(void* h) { return ((Run*)h)->sink.bytes.size(); }
void orc_out_copy(void* h, uint8_t* dst) { Run* r = (Run*)h; memcpy(dst, r->sink.bytes.data(), r->sink.bytes.size()); }
size_t orc_nblocks(void* h) { return ((Run*)h)->dumps.size(); }

// info[16]: in_start,in_end,nblock,crc,orig_ptr,mtf_count,alpha,ngroups,nselectors,bit_start,bit_end,shift,lm_used, in_use[0..]
void orc_block_info(void* h, size_t b, uint64_t* info) {
  const BlockDump& d = ((Run*)h)->dumps[b];
  info[0] = d.in_start; info[1] = d.in_end; info[2] = d.rle.size(); info[3] = d.crc; info[4] = d.orig_ptr;
  info[5] = d.mtf.size(); info[6] = d.alpha; info[7] = d.ngroups; info[8] = d.nselectors;
  info[9] = d.bit_start; info[10] = d.bit_end; info[11] = d.shift; info[12] = d.lm_used;
  info[13] = 0; info[14] = 0; info[15] = 0;
}
void orc_block_inuse(void* h, size_t b, uint32_t* dst) { memcpy(dst, ((Run*)h)->dumps[b].in_use, 32); }

// field: 0 rle(u8) 1 sa(u32) 2 last(u8) 3 mtf(u16) 4 freq(u32) 5..8 selectors pass1..4 (u8) 9..13 lens initial, pass1..4 (u8)
// returns element count; copies min(count, cap_elems) elements.
size_t orc_block_field(void* h, size_t b, int field, void* dst, size_t cap_elems) {
  const BlockDump& d = ((Run*)h)->dumps[b];
  auto cp = [&](const void* src, size_t cnt, size_t esz) {
    if (dst) memcpy(dst, src, std::min(cnt, cap_elems) * esz);
    return cnt;
  };
  switch (field) {
    case 0: return cp(d.rle.data(), d.rle.size(), 1);
    case 1: return cp(d.sa.data(), d.sa.size(), 4);
    case 2: return cp(d.last.data(), d.last.size(), 1);
    case 3: return cp(d.mtf.data(), d.mtf.size(), 2);
    case 4: return cp(d.freq.data(), d.freq.size(), 4);
    case 5: case 6: case 7: case 8: return cp(d.selectors[field - 5].data(), d.selectors[field - 5].size(), 1);
    case 9: case 10: case 11: case 12: case 13: return cp(d.lens[field - 9].data(), d.lens[field - 9].size(), 1);
  }
  return 0;
}

// The reference's block cuts alone (EncoderInner::next / write_rle, encoder.rs:671-716; cut test :692-696; finish
// :729-739): the same run counter and block_buf length bookkeeping as EncoderInner above with everything but the
// lengths left out.  in_off[k] = input offset where block k starts; returns the number of blocks (in_off has nb + 1
// entries, the last one n), or -(needed) when cap is too small.  Used by bench.py --impl reference to hand the blocks
// of one stream to all host cores.
long long orc_cut_table(int level, const uint8_t* in, size_t n, uint64_t* in_off, size_t cap) {
  if (level < 1 || level > 9) return -1;
  const size_t T = (size_t)level * 100000 - 19;
  std::vector<uint64_t> cuts{0};
  size_t blk = 0;          // block_buf.len()
  size_t rle_count = 0;
  uint8_t rle_buffer = 0;
  uint64_t consumed = 0;   // input bytes whose runs have been flushed
  for (size_t i = 0; i < n; ++i) {
    const uint8_t b = in[i];
    if (rle_count == 0) { rle_buffer = b; rle_count = 1; continue; }
    if (rle_buffer == b && rle_count < 255) { rle_count += 1; continue; }
    blk += rle_count < 4 ? rle_count : 5;  // write_rle: up to 4 literals + a count byte
    consumed += rle_count;
    rle_count = 1;
    rle_buffer = b;
    if (blk >= T) { cuts.push_back(consumed); blk = 0; }
  }
  if (n) cuts.push_back(n);  // finish(): the pending run goes into the last block
  const size_t nb = cuts.size() - 1;
  if (cuts.size() > cap) return -(long long)cuts.size();
  for (size_t k = 0; k < cuts.size(); ++k) in_off[k] = cuts[k];
  return (long long)nb;
}

// Full-stream verifier (bench.py / tests hand every block of the GPU's block table to this, in parallel over the
// host cores).  in[0..n) is the input range the GPU assigned to one block; it is encoded as a stream of its own — a
// block cut is a piece boundary, so RLE1 restarts there exactly as in the one-pass encoder — and must come out as
// exactly ONE block.  Its bit section (block magic .. last code) starts at bit 32 of that stream, i.e. at byte 4; the
// bytes from there on are copied to out.
// info[4]: 0 blocks the range produced, 1 block CRC, 2 bytes after RLE1, 3 bits of the block section.
// Returns the bytes written, or -(needed) when cap is too small.
long long orc_encode_block(int level, const uint8_t* in, size_t n, uint8_t* out, size_t cap, uint64_t* info) {
  for (int i = 0; i < 4; ++i) info[i] = 0;
  if (level < 1 || level > 9) return -1;
  Run r;
  run_encoder(level, in, n, r.sink, &r.dumps, false);
  info[0] = r.dumps.size();
  if (r.dumps.empty()) return 0;
  const BlockDump& d = r.dumps[0];
  info[1] = d.crc;
  info[2] = d.rle.size();
  info[3] = d.bit_end - d.bit_start;
  if (d.bit_start != 32) return -1;
  const size_t nbytes = (size_t)((info[3] + 7) / 8);
  if (nbytes > cap) return -(long long)nbytes;
  memcpy(out, r.sink.bytes.data() + 4, nbytes);
  return (long long)nbytes;
}

// First bit (0-based) at which bits [0, nbits) of `sect` differ from bits [at_bit, at_bit+nbits) of `stream`
// (MSB first), nbits when they are equal, or ~0 when the stream is too short to hold them.
uint64_t orc_bits_diff(const uint8_t* stream, size_t stream_len, uint64_t at_bit, const uint8_t* sect, uint64_t nbits) {
  if (at_bit + nbits > (uint64_t)stream_len * 8) return ~0ull;
  auto get = [](const uint8_t* p, uint64_t bit) -> uint64_t {  // 48 bits starting at `bit` (reads 8 bytes)
    uint64_t v = 0;
    const uint64_t by = bit >> 3;
    for (unsigned k = 0; k < 8; ++k) v = (v << 8) | p[by + k];
    return (v << (bit & 7)) >> 16;
  };
  const uint64_t sect_len = (nbits + 7) / 8;
  uint64_t done = 0;
  while (done + 48 <= nbits && (done >> 3) + 8 <= sect_len && ((at_bit + done) >> 3) + 8 <= stream_len) {
    if (get(sect, done) != get(stream, at_bit + done)) break;
    done += 48;
  }
  for (; done < nbits; ++done) {
    const uint64_t pb = at_bit + done;
    const int ba = (sect[done >> 3] >> (7 - (done & 7))) & 1, bb = (stream[pb >> 3] >> (7 - (pb & 7))) & 1;
    if (ba != bb) return done;
  }
  return nbits;
}

// suffix_array::sais::bwt (sais.rs:266-272). mode 0: literal pre-pass only (unbounded), 1: fast pre-pass only,
// 2: default (literal with budget, then fast). Returns shift.
size_t orc_bwt(const uint8_t* s, size_t n, uint32_t* sa_out, int mode) {
  usize shift = 0;
  std::vector<usize> sa = bwt(s, n, 255, &shift, mode == 0 ? UINT64_MAX : mode == 1 ? 0 : 20000000ull);
  for (size_t i = 0; i < n; ++i) sa_out[i] = (uint32_t)sa[i];
  return shift;
}
size_t orc_least_rotation(const uint8_t* s, size_t n, int fast) {
  if (n == 0) return 0;
  if (fast) return least_rotation_fast(s, n);
  std::vector<usize> tmp(n);
  return least_rotation_literal(s, n, tmp.data(), 255, UINT64_MAX);
}

// make_tab_with_fn with kind 0: |x,y| x+y (make_table, cano_huff_table.rs:228-230)
//                         kind 1: the test closure ((x&!0xFF)+(y&!0xFF)) | (max(x&0xFF,y&0xFF)+1) on raw weights
//                         kind 2: bzip2 create_huffman (encoder.rs:641-651) on frequencies
// returns number of lengths written; *used_lm set if the package-merge path was taken.
size_t orc_huffman(const uint64_t* freq, size_t n, size_t lim, int kind, uint8_t* out, int* used_lm) {
  std::vector<usize> f(freq, freq + n);
  bool lm = false;
  std::vector<uint8_t> r;
  if (kind == 0) r = make_tab_with_fn(f, lim, [](usize x, usize y) { return x + y; }, &lm);
  else if (kind == 1)
    r = make_tab_with_fn(f, lim, [](usize x, usize y) {
      return ((x & ~(usize)0xFF) + (y & ~(usize)0xFF)) | (std::max(x & 0xFF, y & 0xFF) + 1); }, &lm);
  else r = create_huffman(f, lim, &lm);
  if (used_lm) *used_lm = lm ? 1 : 0;
  memcpy(out, r.data(), r.size());
  return r.size();
}

void orc_canonical_codes(const uint8_t* lens, size_t n, uint32_t* codes) {
  std::vector<uint8_t> l(lens, lens + n);
  std::vector<uint32_t> c = canonical_codes(l);
  memcpy(codes, c.data(), n * 4);
}

// MSB-first writer (bitio/writer.rs Left): pack (value,len) fields, flush-pad. Returns bytes written.
size_t orc_pack_bits(const uint32_t* values, const uint32_t* lens, size_t n, uint8_t* out, size_t cap) {
  BitSink s;
  for (size_t i = 0; i < n; ++i) s.put(values[i], lens[i]);
  s.flush();
  memcpy(out, s.bytes.data(), std::min(cap, s.bytes.size()));
  return s.bytes.size();
}

uint32_t orc_crc32_bzip2(const uint8_t* p, size_t n) {
  uint32_t v = 0xFFFFFFFFu;
  for (size_t i = 0; i < n; ++i) v = g_crc.update(v, p[i]);
  return ~v;
}

// bzip2/mtf.rs:22-38: MTF positions of a dense-symbol sequence over an alphabet of k symbols.
void orc_mtf_positions(const uint8_t* syms, size_t n, size_t k, uint8_t* out) {
  MtfPosition m(k);
  for (size_t i = 0; i < n; ++i) out[i] = (uint8_t)m.pop(syms[i]);
}

}  // extern "C"
