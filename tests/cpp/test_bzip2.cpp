// The reference's own bzip2 encoder tests (src/bzip2/mod.rs:41-172, src/lib.rs:13-33), restated against the C++
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
  printf(failures ? "FAILED (%d)\n" : "all C++ mirror tests passed\n", failures);
  return failures ? 1 : 0;
}
