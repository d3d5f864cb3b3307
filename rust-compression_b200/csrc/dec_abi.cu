// dec_abi.cu — the decoder entry points of the C ABI (include/bzb200.h section 3) over decoder.cu: device memory
// through the context's buffers, result mapping onto the ABI's return convention, the host→host path and the decoder
// object a shim's `BZip2Decoder` binds (/root/reference/src/bzip2/decoder.rs:584-615).
#include "host_ctx.h"

namespace {

struct CtxDecMem : DecMem {
  bzb200_ctx* c;
  explicit CtxDecMem(bzb200_ctx* c_) : c(c_) {}
  void* buf(int slot, size_t bytes) override {
    if (slot < 0 || slot >= DS_NSLOTS) return nullptr;
    if (ensure(c, c->dec_bufs[slot], bytes) != BZB200_OK) return nullptr;
    return c->dec_bufs[slot].p;
  }
  int cu(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    c->err = std::string(what) + ": " + cudaGetErrorString(e);
    return 1;
  }
  int fill(void* p, int byte, size_t bytes) override {
    return cu(cudaMemsetAsync(p, byte, bytes, c->stream), "decoder cudaMemsetAsync");
  }
  int to_host(void* dst, const void* src, size_t bytes) override {
    if (cu(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream), "decoder D2H")) return 1;
    return cu(cudaStreamSynchronize(c->stream), "decoder sync");
  }
  int to_dev(void* dst, const void* src, size_t bytes) override {
    // pageable source: the runtime stages it before returning, so the caller may reuse src
    return cu(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream), "decoder H2D");
  }
  int crc_blocks(const uint8_t* d_data, const uint64_t* d_off, uint32_t nb, uint32_t* d_crc) override {
    launch_k5_crc(c->L, d_data, d_off, nb, d_crc);
    return check();
  }
  int check() override { return check_launch(c) == BZB200_OK ? 0 : 1; }
  std::string err() override { return c->err; }
};

// Scratch per batch of blocks (about 13 MB per 900 kB block on the split path): BZB200_DEC_BATCH_BYTES, else 24 GB but
// never more than 60 % of the memory that is free on the device right now — the block count of a buffer is untrusted
// input, the batch size must not be.
uint64_t dec_batch_bytes() {
  if (const char* e = getenv("BZB200_DEC_BATCH_BYTES")) {
    unsigned long long x = strtoull(e, nullptr, 10);
    if (x >= 1) return x;
  }
  uint64_t v = (uint64_t)24 << 30;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) v = std::min<uint64_t>(v, (uint64_t)free_b / 10 * 6);
  return std::max<uint64_t>(v, (uint64_t)64 << 20);
}

// Largest output the host->host entry points will allocate for (BZB200_DEC_MAX_OUTPUT; default: half of the free device
// memory).  A few kilobytes of highly compressible blocks can ask for terabytes; that is an argument error, not a
// reason to bring the process down.
uint64_t dec_max_output() {
  if (const char* e = getenv("BZB200_DEC_MAX_OUTPUT")) {
    unsigned long long x = strtoull(e, nullptr, 10);
    if (x >= 1) return x;
  }
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return (uint64_t)1 << 32;
  return std::max<uint64_t>((uint64_t)free_b / 2, (uint64_t)1 << 20);
}

// decode on the device; maps the result onto the ABI's return convention
int decode_device(bzb200_ctx* c, const uint8_t* d_in, size_t n, uint8_t* d_out, size_t cap, size_t* out_n, int* bz_error) {
  CtxDecMem M(c);
  DecResult R;
  // default: split D2 (d2_huff + chunk-parallel d2_mtf_a/b/c, 100 ms per GiB of text); BZB200_DEC_SPLIT=0 selects the
  // fused d2_decode (139 ms), kept as the second implementation the parity tests run as well
  uint32_t flags = DEC_SPLIT_D2;
  if (const char* e = getenv("BZB200_DEC_SPLIT")) flags = (atoi(e) != 0) ? DEC_SPLIT_D2 : 0u;
  const int rc = dec_run(c->L, M, d_in, n, d_out, cap, dec_batch_bytes(), flags, &R);
  c->dec_last = R;
  if (rc != 0) {
    if (c->err.empty()) c->err = "decoder: device memory or launch failure";
    return rc;
  }
  CK(c, cudaStreamSynchronize(c->stream));
  if (R.too_small) {
    *out_n = (size_t)R.needed;
    *bz_error = 0;
    c->err = "decompress: output buffer too small: need " + std::to_string(R.needed) + " bytes";
    return BZB200_E_ARG;
  }
  *out_n = (size_t)R.out_n;
  *bz_error = (int)R.bz_error;
  if (R.bz_error) {
    static const char* kinds[] = {"", "DataError", "DataErrorMagicFirst", "DataErrorMagic", "UnexpectedEof", "Unexpected"};
    c->err = std::string("bzip2 stream error: ") + kinds[R.bz_error <= 5 ? R.bz_error : 5];
    return BZB200_E_DATA;
  }
  return BZB200_OK;
}

}  // namespace

extern "C" {

int bzb200_decompress_device(bzb200_ctx* c, const uint8_t* d_in, size_t n, uint8_t* d_out, size_t cap_bytes,
                             size_t* out_n, int* bz_error) {
  if (!c || !out_n || !bz_error || (!d_in && n) || (!d_out && cap_bytes)) return BZB200_E_ARG;
  *out_n = 0;
  *bz_error = 0;
  TRY(set_device(c));
  return decode_device(c, d_in, n, d_out, cap_bytes, out_n, bz_error);
}

int bzb200_decompress_host(bzb200_ctx* c, const uint8_t* h_in, size_t n, uint8_t* h_out, size_t cap_bytes,
                           size_t* out_n, int* bz_error) {
  if (!c || !out_n || !bz_error || (!h_in && n) || (!h_out && cap_bytes)) return BZB200_E_ARG;
  *out_n = 0;
  *bz_error = 0;
  TRY(set_device(c));
  TRY(ensure(c, c->dec_in, n + 16));
  TRY(ensure(c, c->dec_out, cap_bytes + 16));
  if (n) CK(c, cudaMemcpyAsync(c->dec_in.p, h_in, n, cudaMemcpyHostToDevice, c->stream));
  const int rc = decode_device(c, ptr<uint8_t>(c->dec_in), n, ptr<uint8_t>(c->dec_out), cap_bytes, out_n, bz_error);
  if (rc == BZB200_OK || rc == BZB200_E_DATA) {
    if (*out_n) CK(c, cudaMemcpyAsync(h_out, c->dec_out.p, *out_n, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
  }
  return rc;
}

int bzb200_dec_stats(const bzb200_ctx* c, uint64_t* out, size_t cap) {
  if (!c || !out) return BZB200_E_ARG;
  const DecResult& R = c->dec_last;
  const uint64_t v[8] = {R.streams, R.blocks, R.candidates, R.batches, R.syms, R.pre_rle, R.out_n, 0};
  for (size_t i = 0; i < cap && i < 8; ++i) out[i] = v[i];
  return BZB200_OK;
}

// ---- streaming decoder object (BZip2Decoder)
struct bzb200_dec {
  int device = -1;
  bzb200_ctx* ctx = nullptr;
  std::vector<uint8_t> in;
  std::vector<uint8_t> out;
  size_t rd = 0;
  bool finished = false;
  int kind = 0;
  std::string err;
};

int bzb200_dec_create(int device, bzb200_dec** out) {
  if (!out) return BZB200_E_ARG;
  bzb200_dec* d = new bzb200_dec();
  d->device = device;
  *out = d;
  return BZB200_OK;
}

int bzb200_dec_write(bzb200_dec* d, const uint8_t* p, size_t n) {
  if (!d || (!p && n)) return BZB200_E_ARG;
  if (d->finished) {
    d->err = "write after finish";
    return BZB200_E_STATE;
  }
  try {
    d->in.insert(d->in.end(), p, p + n);
  } catch (const std::exception&) {  // std::bad_alloc must not cross the C boundary
    d->err = "out of host memory";
    return BZB200_E_INTERNAL;
  }
  return BZB200_OK;
}

// host -> host with an output buffer that grows to the size the stream needs
static int decompress_host_with_ctx(bzb200_ctx* c, const uint8_t* in, size_t n, std::vector<uint8_t>& out, int* kind) {
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) free_b = (size_t)1 << 30;
  size_t cap = std::max<size_t>((size_t)1 << 20, n * 6);
  cap = std::min(cap, std::max<size_t>((size_t)1 << 20, free_b / 4));
  TRY(ensure(c, c->dec_in, n + 16));
  if (n) CK(c, cudaMemcpyAsync(c->dec_in.p, in, n, cudaMemcpyHostToDevice, c->stream));
  size_t out_n = 0;
  int rc = BZB200_OK;
  for (int attempt = 0; attempt < 2; ++attempt) {
    TRY(ensure(c, c->dec_out, cap + 16));
    rc = decode_device(c, ptr<uint8_t>(c->dec_in), n, ptr<uint8_t>(c->dec_out), cap, &out_n, kind);
    if (rc != BZB200_E_ARG) break;
    cap = out_n;  // exact size reported by the dry pass
    if (cap > dec_max_output()) {
      c->err = "decompress: the stream expands to " + std::to_string(cap) + " bytes, above the limit (BZB200_DEC_MAX_OUTPUT)";
      return BZB200_E_ARG;
    }
  }
  if (rc != BZB200_OK && rc != BZB200_E_DATA) return rc;
  out.resize(out_n);
  if (out_n) CK(c, cudaMemcpyAsync(out.data(), c->dec_out.p, out_n, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return rc;
}

int bzb200_dec_finish(bzb200_dec* d) {
  if (!d) return BZB200_E_ARG;
  if (d->finished) return d->kind ? BZB200_E_DATA : BZB200_OK;
  if (!d->ctx) {
    int r = bzb200_ctx_create_impl(d->device, nullptr, true, &d->ctx);
    if (r != BZB200_OK) {
      d->err = d->ctx ? d->ctx->err : "context creation failed";
      if (d->ctx) { bzb200_ctx_destroy(d->ctx); d->ctx = nullptr; }
      return r;
    }
  }
  int r = set_device(d->ctx);
  try {
    if (r == BZB200_OK) r = decompress_host_with_ctx(d->ctx, d->in.data(), d->in.size(), d->out, &d->kind);
  } catch (const std::exception&) {
    d->ctx->err = "out of host memory";
    r = BZB200_E_INTERNAL;
  }
  if (r != BZB200_OK && r != BZB200_E_DATA) {
    d->err = d->ctx->err;
    return r;
  }
  if (r == BZB200_E_DATA) d->err = d->ctx->err;
  d->in.clear();
  d->in.shrink_to_fit();
  d->rd = 0;
  d->finished = true;
  return r;
}

int bzb200_dec_error_kind(const bzb200_dec* d) { return (d && d->finished) ? d->kind : 0; }

size_t bzb200_dec_read(bzb200_dec* d, uint8_t* dst, size_t cap) {
  if (!d || !d->finished || !dst) return 0;
  size_t n = std::min(cap, d->out.size() - d->rd);
  if (n) memcpy(dst, d->out.data() + d->rd, n);
  d->rd += n;
  return n;
}

size_t bzb200_dec_output_size(const bzb200_dec* d) { return (d && d->finished) ? d->out.size() : 0; }

int bzb200_dec_reset(bzb200_dec* d) {
  if (!d) return BZB200_E_ARG;
  d->in.clear();
  d->out.clear();
  d->rd = 0;
  d->finished = false;
  d->kind = 0;
  d->err.clear();
  return BZB200_OK;
}

void bzb200_dec_destroy(bzb200_dec* d) {
  if (!d) return;
  if (d->ctx) bzb200_ctx_destroy(d->ctx);
  delete d;
}

const char* bzb200_dec_last_error(const bzb200_dec* d) { return d ? d->err.c_str() : "null decoder"; }

int bzb200_decompress(int device, const uint8_t* in, size_t n, uint8_t** out, size_t* out_n, int* bz_error) {
  if (!out || !out_n || !bz_error || (!in && n)) return BZB200_E_ARG;
  *out = nullptr;
  *out_n = 0;
  *bz_error = 0;
  bzb200_ctx* c = nullptr;
  int r = bzb200_ctx_create_impl(device, nullptr, true, &c);
  if (r != BZB200_OK) {
    if (c) bzb200_ctx_destroy(c);
    return r;
  }
  std::vector<uint8_t> o;
  try {
    r = decompress_host_with_ctx(c, in, n, o, bz_error);
  } catch (const std::exception&) {
    r = BZB200_E_INTERNAL;
  }
  if (r == BZB200_OK || r == BZB200_E_DATA) {
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

}  // extern "C"
