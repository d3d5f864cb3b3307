// enc_stream.cu — the streaming encoder object of the C ABI (include/bzb200.h section 1): what a shim's
// `BZip2Encoder` binds.  replaces BZip2Encoder::new / Encoder::next (/root/reference/src/bzip2/encoder.rs:58-158) on the
// host side; every stage runs through the context API of pipeline.cu.
#include "host_ctx.h"

extern "C" {

// ------------------------------------------------------------------ streaming encoder (BZip2Encoder)
// Action::Run pipelining: input accumulates in a window; when the window is full, the bytes from
// the last block cut onwards are planned as a stream of their own, every block that is already closed is encoded and
// its bytes become readable at once (the reference also yields a block as soon as it closes, encoder.rs:91-107), and
// the still-open last block stays buffered.  A block cut is a piece boundary, so RLE1 restarts there exactly as in
// the one-pass plan; the partial last byte of the bit stream is carried into the next window.
struct bzb200_enc {
  int level = 9;
  int device = -1;
  bzb200_ctx* ctx = nullptr;
  std::vector<uint8_t> in;   // input from the last block cut onwards
  std::vector<uint8_t> out;  // finished output bytes
  size_t rd = 0;
  bool finished = false;
  bool started = false;      // the stream header has been written
  uint32_t carry_bits = 0;   // valid bits (0..7) of the partial last byte
  uint8_t carry = 0;
  uint32_t combined = 0;     // combined CRC of the blocks encoded so far (encoder.rs:237-238)
  uint64_t blocks = 0, windows = 0;
  size_t window = (size_t)256 << 20;
  std::string err;
};

int bzb200_enc_create(int level, int device, bzb200_enc** out) {
  if (!out) return BZB200_E_ARG;
  *out = nullptr;
  if (level < 1 || level > 9) return BZB200_E_LEVEL;  // BZip2Encoder::new panics "invalid level"
  bzb200_enc* e = new bzb200_enc();
  e->level = level;
  e->device = device;
  if (const char* w = getenv("BZB200_ENC_WINDOW")) {
    unsigned long long v = strtoull(w, nullptr, 10);
    if (v >= 1) e->window = (size_t)v;
  }
  *out = e;
  return BZB200_OK;
}

// Compresses the closed blocks of the buffered input (all blocks and the trailer when `final`).
static int enc_pump(bzb200_enc* e, bool final) {
  if (!e->ctx) {
    int r = bzb200_ctx_create_impl(e->device, nullptr, true, &e->ctx);
    if (r != BZB200_OK) {
      e->err = e->ctx ? e->ctx->err : "context creation failed";
      if (e->ctx) { bzb200_ctx_destroy(e->ctx); e->ctx = nullptr; }
      return r;
    }
  }
  bzb200_ctx* c = e->ctx;
  int r = BZB200_OK;
  auto fail = [&](int code) {
    if (code == BZB200_E_CUDA && c->err.empty()) c->err = std::string("cuda: ") + cudaGetErrorString(cudaGetLastError());
    e->err = c->err;
    return code;
  };
  if ((r = set_device(c)) != BZB200_OK) return fail(r);
  const size_t n = e->in.size();
  const size_t cap = bzb200_max_output_bytes(e->level, n) + 16;
  if ((r = ensure(c, c->stage_in, n + 16)) != BZB200_OK) return fail(r);
  if ((r = ensure(c, c->stage_out, cap)) != BZB200_OK) return fail(r);
  uint8_t* d_in = ptr<uint8_t>(c->stage_in);
  uint8_t* d_out = ptr<uint8_t>(c->stage_out);
  if (n && cudaMemcpyAsync(d_in, e->in.data(), n, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return fail(BZB200_E_CUDA);
  if (cudaMemsetAsync(d_out, 0, cap, c->stream) != cudaSuccess) return fail(BZB200_E_CUDA);
  uint64_t bit = 0;
  if (!e->started) {
    if ((r = bzb200_write_stream_header(c, e->level, d_out, cap)) != BZB200_OK) return fail(r);
    bit = 32;
    e->started = true;
  } else if (e->carry_bits) {
    if (cudaMemcpyAsync(d_out, &e->carry, 1, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return fail(BZB200_E_CUDA);
    bit = e->carry_bits;
  }
  uint32_t nb = 0;
  if ((r = bzb200_plan(c, e->level, d_in, n, &nb)) != BZB200_OK) return fail(r);
  const uint32_t nenc = final ? nb : (nb ? nb - 1 : 0);  // the last block of a window is still open
  size_t consumed = 0;
  if (nenc) {
    if ((r = bzb200_encode_blocks(c, 0, nenc, d_out, cap, bit, &bit)) != BZB200_OK) return fail(r);
    e->combined = bzb200_combine_crc(e->combined, c->h_crc.data(), nenc);
    consumed = (size_t)c->h_in_off[nenc];
    e->blocks += nenc;
  }
  size_t take = (size_t)(bit / 8);  // whole bytes that are final
  if (final) {
    if ((r = bzb200_write_stream_trailer(c, d_out, cap, bit, e->combined, &take)) != BZB200_OK) return fail(r);
    e->carry_bits = 0;
  } else {
    e->carry_bits = (uint32_t)(bit & 7);
  }
  const size_t fetch = take + ((!final && e->carry_bits) ? 1 : 0);
  const size_t at = e->out.size();
  e->out.resize(at + fetch);
  if (fetch && cudaMemcpyAsync(e->out.data() + at, d_out, fetch, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
    return fail(BZB200_E_CUDA);
  if (cudaStreamSynchronize(c->stream) != cudaSuccess) return fail(BZB200_E_CUDA);
  if (fetch > take) {
    e->carry = e->out.back();
    e->out.pop_back();
  }
  e->in.erase(e->in.begin(), e->in.begin() + consumed);
  ++e->windows;
  return BZB200_OK;
}

int bzb200_enc_write(bzb200_enc* e, const uint8_t* p, size_t n) {
  if (!e || (!p && n)) return BZB200_E_ARG;
  if (e->finished) {
    e->err = "write after finish";
    return BZB200_E_STATE;
  }
  while (n) {
    const size_t room = e->in.size() < e->window ? e->window - e->in.size() : 0;
    const size_t take = std::min(n, std::max<size_t>(room, 1));
    e->in.insert(e->in.end(), p, p + take);
    p += take;
    n -= take;
    if (e->in.size() >= e->window) {
      const size_t before = e->in.size();
      int r = enc_pump(e, false);
      if (r != BZB200_OK) return r;
      if (e->in.size() == before) {
        // not even one closed block in a full window (window smaller than a block): let the window grow
        e->in.insert(e->in.end(), p, p + n);
        e->window = std::max(e->window * 2, e->in.size() + 1);
        n = 0;
      }
    }
  }
  return BZB200_OK;
}

int bzb200_enc_finish(bzb200_enc* e) {
  if (!e) return BZB200_E_ARG;
  if (e->finished) return BZB200_OK;
  int r = enc_pump(e, true);
  if (r != BZB200_OK) return r;
  e->in.clear();
  e->in.shrink_to_fit();
  e->finished = true;
  return BZB200_OK;
}

size_t bzb200_enc_read(bzb200_enc* e, uint8_t* dst, size_t cap) {
  if (!e || !dst) return 0;
  size_t n = std::min(cap, e->out.size() - e->rd);
  if (n) memcpy(dst, e->out.data() + e->rd, n);
  e->rd += n;
  if (e->rd == e->out.size() && !e->finished) {  // drained mid-stream: drop the bytes already handed out
    e->out.clear();
    e->rd = 0;
  }
  return n;
}

size_t bzb200_enc_output_size(const bzb200_enc* e) { return e ? e->out.size() - e->rd : 0; }

int bzb200_enc_reset(bzb200_enc* e) {
  if (!e) return BZB200_E_ARG;
  e->in.clear();
  e->out.clear();
  e->rd = 0;
  e->finished = false;
  e->started = false;
  e->carry_bits = 0;
  e->carry = 0;
  e->combined = 0;
  e->blocks = 0;
  e->windows = 0;
  e->err.clear();
  return BZB200_OK;
}

void bzb200_enc_destroy(bzb200_enc* e) {
  if (!e) return;
  if (e->ctx) bzb200_ctx_destroy(e->ctx);
  delete e;
}

const char* bzb200_enc_last_error(const bzb200_enc* e) { return e ? e->err.c_str() : "null encoder"; }

int bzb200_enc_stats(const bzb200_enc* e, uint64_t* out, size_t cap) {
  if (!e || !out) return BZB200_E_ARG;
  const uint64_t v[4] = {e->blocks, e->windows, (uint64_t)e->in.size(), (uint64_t)(e->out.size() - e->rd)};
  for (size_t i = 0; i < cap && i < 4; ++i) out[i] = v[i];
  return BZB200_OK;
}

static int compress_host_with_ctx(bzb200_ctx* c, int level, const uint8_t* in, size_t n, std::vector<uint8_t>& out) {
  DevBuf din, dout;
  const size_t cap = bzb200_max_output_bytes(level, n);
  int r = BZB200_OK;
  do {
    if ((r = ensure(c, din, n + 16)) != BZB200_OK) break;
    if ((r = ensure(c, dout, cap)) != BZB200_OK) break;
    if (n && cudaMemcpyAsync(din.p, in, n, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { r = BZB200_E_CUDA; break; }
    if (cudaMemsetAsync(dout.p, 0, cap, c->stream) != cudaSuccess) { r = BZB200_E_CUDA; break; }
    size_t out_n = 0;
    if ((r = bzb200_compress_device(c, level, (const uint8_t*)din.p, n, (uint8_t*)dout.p, cap, &out_n)) != BZB200_OK) break;
    out.resize(out_n);
    if (cudaMemcpyAsync(out.data(), dout.p, out_n, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { r = BZB200_E_CUDA; break; }
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) { r = BZB200_E_CUDA; break; }
  } while (0);
  if (r == BZB200_E_CUDA && c->err.empty()) c->err = std::string("cuda: ") + cudaGetErrorString(cudaGetLastError());
  if (din.p) cudaFree(din.p);
  if (dout.p) cudaFree(dout.p);
  return r;
}

int bzb200_compress(int level, int device, const uint8_t* in, size_t n, uint8_t** out, size_t* out_n) {
  if (!out || !out_n || (!in && n)) return BZB200_E_ARG;
  *out = nullptr;
  *out_n = 0;
  if (level < 1 || level > 9) return BZB200_E_LEVEL;
  bzb200_ctx* c = nullptr;
  int r = bzb200_ctx_create_impl(device, nullptr, true, &c);
  if (r != BZB200_OK) {
    if (c) bzb200_ctx_destroy(c);
    return r;
  }
  std::vector<uint8_t> o;
  r = compress_host_with_ctx(c, level, in, n, o);
  if (r == BZB200_OK) {
    *out = (uint8_t*)malloc(o.size() ? o.size() : 1);
    if (!*out) r = BZB200_E_ARG;
    else {
      memcpy(*out, o.data(), o.size());
      *out_n = o.size();
    }
  }
  bzb200_ctx_destroy(c);
  return r;
}

void bzb200_free(void* p) { free(p); }

}  // extern "C"
