// The reference's own bzip2 encoder and decoder tests (src/bzip2/mod.rs:41-172, src/lib.rs:13-33), restated against the C++
// mirror of its API (rust-compression_b200/csrc/bzb200.hpp).  Expected streams come from the CPU oracle (linked
// here as the CHECKER only) and from the reference's golden vector.  Needs a CUDA device.
//   usage: test_bzip2 <dir with sample1-3.ref>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "../../rust-compression_b200/csrc/bzb200.hpp"

extern "C" long long orc_compress(int level, const uint8_t* in, size_t n, uint8_t* out, size_t cap);
extern "C" int orc_decode(const uint8_t* in, size_t n, uint8_t** out, size_t* out_n);
extern "C" void orc_decode_free(uint8_t* p);

using namespace compression;

static int failures = 0;
#define CHECK(c)                                                   \
  do {                                                             \
    if (!(c)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); ++failures; } \
  } while (0)

static std::vector<uint8_t> oracle(const std::vector<uint8_t>& in, int level) {
  std::vector<uint8_t> out(in.size() * 2 + 4096);
  long long n = orc_compress(level, in.data(), in.size(), out.data(), out.size());
  out.resize(n > 0 ? (size_t)n : 0);
  return out;
}

static std::vector<uint8_t> read_file(const std::string& p) {
  std::ifstream f(p, std::ios::binary);
  return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

static void test_unit() {  // bzip2/mod.rs:41-58
  std::vector<uint8_t> in = {'a', '\n'};
  BZip2Encoder enc(9);
  std::vector<uint8_t> ret;
  CHECK(encode_collect(in, enc, Action::Finish, ret));
  const std::vector<uint8_t> want = {0x42, 0x5A, 0x68, 0x39, 0x31, 0x41, 0x59, 0x26, 0x53, 0x59, 0x63, 0x3E, 0xD6,
                                     0xE2, 0x00, 0x00, 0x00, 0xC1, 0x00, 0x00, 0x10, 0x20, 0x00, 0x20, 0x00, 0x21,
                                     0x00, 0x82, 0xB1, 0x77, 0x24, 0x53, 0x85, 0x09, 0x06, 0x33, 0xED, 0x6E, 0x20};
  CHECK(ret == want);
}

static void test_sample(const std::string& dir, int idx, int level) {  // bzip2/mod.rs:84-139
  auto in = read_file(dir + "/sample" + std::to_string(idx) + ".ref");
  CHECK(!in.empty());
  BZip2Encoder encoder(level);
  std::vector<uint8_t> ret;
  CHECK(encode_collect(in, encoder, Action::Finish, ret));
  CHECK(ret == oracle(in, level));
}

static void test_long() {  // bzip2/mod.rs:150-172
  std::vector<uint8_t> data(1000, 'a');
  BZip2Encoder enc(9);
  std::vector<uint8_t> ret;
  CHECK(encode_collect(data, enc, Action::Finish, ret));
  CHECK(ret == oracle(data, 9));
}

static void test_doc() {  // lib.rs:13-33
  std::string s = "aabbaabbaabbaabb\n";
  std::vector<uint8_t> in(s.begin(), s.end());
  BZip2Encoder enc(9);
  std::vector<uint8_t> ret;
  CHECK(encode_collect(in, enc, Action::Finish, ret));
  CHECK(ret == oracle(in, 9));
}

static void test_invalid_level() {  // encoder.rs:59-61
  bool threw = false;
  try { BZip2Encoder e(0); } catch (const std::invalid_argument&) { threw = true; }
  CHECK(threw);
  threw = false;
  try { BZip2Encoder e(10); } catch (const std::invalid_argument&) { threw = true; }
  CHECK(threw);
}

static void test_run_then_finish() {  // Action::Run semantics, encoder.rs:91-107,142-144
  std::vector<uint8_t> a(30000, 'x'), b;
  for (int i = 0; i < 50000; ++i) b.push_back((uint8_t)("etaoin shrdlu"[i % 13]));
  BZip2Encoder enc(9);
  auto it = a.begin();
  CHECK(!enc.next(it, a.end(), Action::Run).has_value());
  std::vector<uint8_t> ret;
  CHECK(encode_collect(b, enc, Action::Finish, ret));
  std::vector<uint8_t> all(a);
  all.insert(all.end(), b.begin(), b.end());
  CHECK(ret == oracle(all, 9));
  // encoder is re-armed after None
  std::vector<uint8_t> in = {'a', '\n'}, again;
  CHECK(encode_collect(in, enc, Action::Finish, again));
  CHECK(again == oracle(in, 9));
}

// bzip2/mod.rs:72-82 check_unzip: actual.decode(&mut BZip2Decoder::new()).collect() == Ok(expected)
static void check_unzip(const std::vector<uint8_t>& actual, const std::vector<uint8_t>& expected) {
  BZip2Decoder dec;
  std::vector<uint8_t> ret;
  BZip2Error e = BZip2Error::Unexpected;
  CHECK(decode_collect(actual, dec, ret, &e));
  CHECK(ret == expected);
}

static void test_decode_samples(const std::string& dir) {  // bzip2/mod.rs:84-148 (the .bz2 halves), :150-172
  for (int idx = 1; idx <= 4; ++idx) {
    auto ref = read_file(dir + "/sample" + std::to_string(idx) + ".ref");
    auto bz = read_file(dir + "/sample" + std::to_string(idx) + ".bz2");
    CHECK(!ref.empty() && !bz.empty());
    check_unzip(bz, ref);
    if (idx <= 3) {  // encode at level idx, then unzip (the first halves of test_sample1..3)
      BZip2Encoder encoder(idx);
      std::vector<uint8_t> ret;
      CHECK(encode_collect(ref, encoder, Action::Finish, ret));
      check_unzip(ret, ref);
    }
  }
  std::vector<uint8_t> data(1000, 'a'), comp;
  BZip2Encoder enc(9);
  CHECK(encode_collect(data, enc, Action::Finish, comp));
  check_unzip(comp, data);
}

static void test_decode_errors(const std::string& dir) {  // same bytes and BZip2Error kind as the restated reference
  auto bz = read_file(dir + "/sample1.bz2");
  std::vector<std::vector<uint8_t>> bad;
  bad.push_back({});                                                              // DataErrorMagicFirst
  bad.push_back(std::vector<uint8_t>(bz.begin(), bz.begin() + bz.size() / 2));    // cut inside the block
  bad.push_back(std::vector<uint8_t>(bz.begin(), bz.end() - 2));                  // cut inside the combined CRC
  { auto t = bz; t[bz.size() / 2] ^= 0x04; bad.push_back(t); }                    // flipped bit
  { auto t = bz; t.push_back('x'); t.push_back('y'); bad.push_back(t); }          // trailing garbage: DataErrorMagic
  for (const auto& b : bad) {
    uint8_t* o = nullptr;
    size_t on = 0;
    const int want = orc_decode(b.data(), b.size(), &o, &on);
    std::vector<uint8_t> want_bytes(o, o + on);
    orc_decode_free(o);
    BZip2Decoder dec;
    std::vector<uint8_t> ret;
    BZip2Error e = BZip2Error::Unexpected;
    const bool ok = decode_collect(b, dec, ret, &e);
    CHECK(ok == (want == 0));
    CHECK(ok || (int)e == want);
    CHECK(ret == want_bytes);
  }
}

static void test_run_streams_blocks_early() {  // closed blocks become readable under Action::Run
  setenv("BZB200_ENC_WINDOW", "300000", 1);
  std::vector<uint8_t> all;
  for (int i = 0; i < 1000000; ++i) all.push_back((uint8_t)("the quick brown fox jumps over the lazy dog. "[(i * 7 + i / 13) % 45]));
  BZip2Encoder enc(1);
  unsetenv("BZB200_ENC_WINDOW");
  std::vector<uint8_t> ret;
  size_t early = 0;
  for (size_t lo = 0; lo < all.size(); lo += 250000) {
    std::vector<uint8_t> part(all.begin() + lo, all.begin() + std::min(all.size(), lo + 250000));
    auto it = part.begin();
    while (auto r = enc.next(it, part.end(), Action::Run)) {
      CHECK(r->ok);
      ret.push_back(r->value);
    }
    early = ret.size();
  }
  CHECK(early > 0);  // bytes were handed out before Finish
  std::vector<uint8_t> none, rest;
  CHECK(encode_collect(none, enc, Action::Finish, rest));
  ret.insert(ret.end(), rest.begin(), rest.end());
  CHECK(ret == oracle(all, 1));
}

int main(int argc, char** argv) {
  std::string dir = argc > 1 ? argv[1] : "tests/golden/data";
  test_invalid_level();
  test_unit();
  test_sample(dir, 1, 1);
  test_sample(dir, 2, 2);
  test_sample(dir, 3, 3);
  test_long();
  test_doc();
  test_run_then_finish();
  test_run_streams_blocks_early();
  test_decode_samples(dir);
  test_decode_errors(dir);
  printf(failures ? "FAILED (%d)\n" : "all C++ mirror tests passed\n", failures);
  return failures ? 1 : 0;
}
