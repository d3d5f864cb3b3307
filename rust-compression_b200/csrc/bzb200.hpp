// bzb200.hpp — C++ host-side mirror of the reference's operator interface for the bzip2 encode path, on top of
// the C ABI (include/bzb200.h).  Header-only.  Same names, argument meaning and error behaviour as
// chalharu/rust-compression:
//   Action            src/action.rs:8-13
//   CompressionError  src/error.rs:10-42
//   BZip2Encoder      src/bzip2/encoder.rs:39-159  (new(level) panics "invalid level" -> std::invalid_argument)
//   encode(...)       src/traits/encoder.rs:12-79  (EncodeExt::encode / EncodeIterator)
// The Rust toolchain is absent from the build image, so this is the compiled-language host layer the parity tests
// in tests/cpp/ are written against; the Rust shim itself ships as source (rust/bzip2_b200.rs).
#pragma once
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/bzb200.h"

namespace compression {

enum class Action { Run, Flush, Finish };

enum class CompressionError { DataError, UnexpectedEof, Unexpected };

inline const char* description(CompressionError e) {  // error.rs:35-41
  switch (e) {
    case CompressionError::DataError: return "data integrity error in data";
    case CompressionError::UnexpectedEof: return "file ends unexpectedly";
    default: return "unexpected error";
  }
}

// Result<u8, CompressionError>
struct ByteResult {
  bool ok;
  uint8_t value;
  CompressionError error;
};

class BZip2Encoder {
 public:
  explicit BZip2Encoder(size_t level = 9, int device = -1) {
    if (level < 1 || level > 9) throw std::invalid_argument("invalid level");  // encoder.rs:59-61 panics
    if (bzb200_enc_create((int)level, device, &h_) != BZB200_OK || !h_) throw std::runtime_error("bzb200_enc_create");
  }
  ~BZip2Encoder() {
    if (h_) bzb200_enc_destroy(h_);
  }
  BZip2Encoder(const BZip2Encoder&) = delete;
  BZip2Encoder& operator=(const BZip2Encoder&) = delete;

  // Encoder::next (encoder.rs:120-158): the next output byte, or nullopt when drained.  Under Action::Run the bytes
  // of blocks that have already closed are handed out as they become ready (encoder.rs:91-107).
  template <class It>
  std::optional<ByteResult> next(It& it, const It& end, Action action) {
    for (;;) {
      if (pos_ < out_.size()) return ByteResult{true, out_[pos_++], CompressionError::Unexpected};
      if (!finished_) {
        std::vector<uint8_t> in;
        for (; it != end; ++it) {
          in.push_back((uint8_t)*it);
          if (in.size() == kChunk) {
            if (bzb200_enc_write(h_, in.data(), in.size()) != BZB200_OK) return err();
            in.clear();
          }
        }
        if (!in.empty() && bzb200_enc_write(h_, in.data(), in.size()) != BZB200_OK) return err();
        if (action == Action::Finish) {  // Flush is out of contract and behaves like Run
          if (bzb200_enc_finish(h_) != BZB200_OK) return err();
          finished_ = true;
        }
      }
      out_.resize(kChunk);
      const size_t n = bzb200_enc_read(h_, out_.data(), kChunk);
      out_.resize(n);
      pos_ = 0;
      if (n) continue;
      if (finished_) {
        finished_ = false;  // re-arm (encoder.rs:87-90,130-133)
        bzb200_enc_reset(h_);
      }
      return std::nullopt;
    }
  }

  std::string last_error() const { return bzb200_enc_last_error(h_); }

 private:
  static constexpr size_t kChunk = 1 << 20;
  std::optional<ByteResult> err() { return ByteResult{false, 0, CompressionError::Unexpected}; }
  bzb200_enc* h_ = nullptr;
  std::vector<uint8_t> out_;
  size_t pos_ = 0;
  bool finished_ = false;
};

// iter.encode(&mut encoder, action).collect::<Result<Vec<_>, _>>()
template <class Container>
inline bool encode_collect(const Container& input, BZip2Encoder& enc, Action action, std::vector<uint8_t>& out,
                           CompressionError* e = nullptr) {
  auto it = input.begin();
  const auto end = input.end();
  out.clear();
  while (auto r = enc.next(it, end, action)) {
    if (!r->ok) {
      if (e) *e = r->error;
      return false;
    }
    out.push_back(r->value);
  }
  return true;
}

// ---- decode side ----
//   BZip2Error     src/bzip2/error.rs:4-52
//   BZip2Decoder   src/bzip2/decoder.rs:584-615
//   decode(...)    src/traits/decoder.rs:14-99 (DecodeExt::decode / DecodeIterator)
enum class BZip2Error { DataError = 1, DataErrorMagicFirst, DataErrorMagic, UnexpectedEof, Unexpected };

inline const char* description(BZip2Error e) {  // error.rs:31-41
  switch (e) {
    case BZip2Error::DataError: return "data integrity (CRC) error in data";
    case BZip2Error::DataErrorMagicFirst: return "bad magic number (file not created by bzip2)";
    case BZip2Error::DataErrorMagic: return "trailing garbage after EOF ignored";
    case BZip2Error::UnexpectedEof: return "file ends unexpectedly";
    default: return "unexpected error";
  }
}

inline CompressionError to_compression_error(BZip2Error e) {  // error.rs:44-52
  if (e == BZip2Error::UnexpectedEof) return CompressionError::UnexpectedEof;
  if (e == BZip2Error::Unexpected) return CompressionError::Unexpected;
  return CompressionError::DataError;
}

struct DecodedByte {  // Result<u8, BZip2Error>
  bool ok;
  uint8_t value;
  BZip2Error error;
};

class BZip2Decoder {
 public:
  explicit BZip2Decoder(int device = -1) {
    if (bzb200_dec_create(device, &h_) != BZB200_OK || !h_) throw std::runtime_error("bzb200_dec_create");
  }
  ~BZip2Decoder() {
    if (h_) bzb200_dec_destroy(h_);
  }
  BZip2Decoder(const BZip2Decoder&) = delete;
  BZip2Decoder& operator=(const BZip2Decoder&) = delete;

  // Decoder::next (decoder.rs:607-614): the next output byte; Some(Err(kind)) after the bytes that precede an error;
  // nullopt when drained.
  template <class It>
  std::optional<DecodedByte> next(It& it, const It& end) {
    for (;;) {
      if (pos_ < out_.size()) return DecodedByte{true, out_[pos_++], BZip2Error::Unexpected};
      if (decoded_) {
        out_.resize(kChunk);
        const size_t n = bzb200_dec_read(h_, out_.data(), kChunk);
        out_.resize(n);
        pos_ = 0;
        if (n) continue;
        const int kind = bzb200_dec_error_kind(h_);
        decoded_ = false;
        bzb200_dec_reset(h_);
        if (kind) return DecodedByte{false, 0, (BZip2Error)kind};
        return std::nullopt;
      }
      std::vector<uint8_t> in;
      for (; it != end; ++it) {
        in.push_back((uint8_t)*it);
        if (in.size() == kChunk) {
          if (bzb200_dec_write(h_, in.data(), in.size()) != BZB200_OK) return DecodedByte{false, 0, BZip2Error::Unexpected};
          in.clear();
        }
      }
      if (!in.empty() && bzb200_dec_write(h_, in.data(), in.size()) != BZB200_OK)
        return DecodedByte{false, 0, BZip2Error::Unexpected};
      const int rc = bzb200_dec_finish(h_);
      if (rc != BZB200_OK && rc != BZB200_E_DATA) {
        bzb200_dec_reset(h_);
        return DecodedByte{false, 0, BZip2Error::Unexpected};
      }
      decoded_ = true;
      out_.clear();
      pos_ = 0;
    }
  }

 private:
  static constexpr size_t kChunk = 1 << 20;
  bzb200_dec* h_ = nullptr;
  std::vector<uint8_t> out_;
  size_t pos_ = 0;
  bool decoded_ = false;
};

// bytes.decode(&mut decoder).collect::<Result<Vec<_>, _>>(): false + *e on error (out keeps the bytes before it)
template <class Container>
inline bool decode_collect(const Container& input, BZip2Decoder& dec, std::vector<uint8_t>& out, BZip2Error* e = nullptr) {
  auto it = input.begin();
  const auto end = input.end();
  out.clear();
  while (auto r = dec.next(it, end)) {
    if (!r->ok) {
      if (e) *e = r->error;
      return false;
    }
    out.push_back(r->value);
  }
  return true;
}

}  // namespace compression
