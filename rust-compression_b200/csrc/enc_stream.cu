// enc_stream.cu — the streaming encoder object of the C ABI (include/bzb200.h section 1): what a shim's
// `BZip2Encoder` binds.  replaces BZip2Encoder::new / Encoder::next (/root/reference/src/bzip2/encoder.rs:58-158) on the
// host side; every stage runs through the context API of pipeline.cu.
#include "host_ctx.h"

#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

#include "mgpu.h"

extern "C" {
int bzb200_pool_create(int ngpus, const int* devices, bzb200_pool** out);
void bzb200_pool_destroy(bzb200_pool* p);
}

// ------------------------------------------------------------------ streaming encoder (BZip2Encoder)
// Action::Run pipelining (the reference hands a block out as soon as it closes, encoder.rs:91-107).
// Input is collected in one of two window buffers; a full window is handed to the object's worker thread, which runs
// it through the GPU engine (mgpu.cu) while the caller keeps filling the other buffer.  A window is planned as a stream
// of its own from the last block cut onwards — a cut is a piece boundary, so RLE1 restarts there exactly as in the
// one-pass plan — every block but the still-open last one is encoded, their bytes become readable at once, and the
// open block's input is carried in front of the next window together with the partial last byte of the bit stream.
//   one GPU   : the windows live in DEVICE memory.  bzb200_enc_write copies the caller's bytes straight there (one DMA
//               per call when the caller's buffer is pinned, the driver's staged copy otherwise) and returns when the
//               copy is done, so the input never takes a second trip through host memory and is already in HBM when
//               its window is submitted; finished bytes stay on the device until bzb200_enc_read copies them out.
//   n GPUs    : the windows are pinned host buffers (the engine's workers copy their slices from there).
namespace {

struct DeviceScope {  // switches the calling thread to `dev` and back
  int prev = -1;
  bool ok = false;
  explicit DeviceScope(int dev) {
    if (dev < 0) return;
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); return; }
    ok = cudaSetDevice(dev) == cudaSuccess;
    if (!ok) cudaGetLastError();
  }
  ~DeviceScope() {
    if (ok && prev >= 0) cudaSetDevice(prev);
  }
};

// A byte buffer in pinned host memory (pageable if pinning fails) or in device memory.
struct Buf {
  uint8_t* p = nullptr;
  size_t cap = 0;
  int dev = -1;          // >= 0: device memory on that GPU
  bool pinned = false;
  cudaStream_t cs = nullptr;  // device buffers: stream of the copies
  void release() {
    if (p) {
      if (dev >= 0) {
        DeviceScope ds(dev);
        cudaFree(p);
      } else if (pinned) {
        cudaFreeHost(p);
      } else {
        free(p);
      }
    }
    p = nullptr;
    cap = 0;
  }
  // at least `bytes`; keeps [keep_lo, keep_hi), moved up by `shift`
  bool reserve(size_t bytes, size_t keep_lo = 0, size_t keep_hi = 0, size_t shift = 0) {
    if (cap >= bytes && shift == 0) return true;
    uint8_t* q = nullptr;
    bool pin = false;
    if (dev >= 0) {
      DeviceScope ds(dev);
      if (!ds.ok || cudaMalloc((void**)&q, bytes) != cudaSuccess) { cudaGetLastError(); return false; }
      if (p && keep_hi > keep_lo) {
        if (cudaMemcpyAsync(q + keep_lo + shift, p + keep_lo, keep_hi - keep_lo, cudaMemcpyDeviceToDevice, cs) != cudaSuccess ||
            cudaStreamSynchronize(cs) != cudaSuccess) {
          cudaGetLastError();
          cudaFree(q);
          return false;
        }
      }
    } else {
      pin = cudaHostAlloc((void**)&q, bytes, cudaHostAllocPortable) == cudaSuccess;
      if (!pin) {
        cudaGetLastError();
        q = (uint8_t*)malloc(bytes);
        if (!q) return false;
      }
      if (p && keep_hi > keep_lo) memcpy(q + keep_lo + shift, p + keep_lo, keep_hi - keep_lo);
    }
    release();
    p = q;
    cap = bytes;
    pinned = pin;
    return true;
  }
  // n bytes of host memory -> [off, off + n); done when the call returns (copy semantics of bzb200_enc_write)
  bool put(size_t off, const uint8_t* src, size_t n) {
    if (dev < 0) {
      memcpy(p + off, src, n);
      return true;
    }
    DeviceScope ds(dev);
    if (!ds.ok || cudaMemcpyAsync(p + off, src, n, cudaMemcpyDefault, cs) != cudaSuccess ||
        cudaStreamSynchronize(cs) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    return true;
  }
  // bytes [a, b) of `o` (same kind of memory) -> [at, at + b - a)
  bool take(const Buf& o, size_t a, size_t b, size_t at) {
    if (dev < 0) {
      memcpy(p + at, o.p + a, b - a);
      return true;
    }
    DeviceScope ds(dev);
    if (!ds.ok || cudaMemcpyAsync(p + at, o.p + a, b - a, cudaMemcpyDeviceToDevice, cs) != cudaSuccess ||
        cudaStreamSynchronize(cs) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    return true;
  }
  // [off, off + n) -> host memory
  bool get(size_t off, uint8_t* dst, size_t n) const {
    if (dev < 0) {
      memcpy(dst, p + off, n);
      return true;
    }
    DeviceScope ds(dev);
    if (!ds.ok || cudaMemcpy(dst, p + off, n, cudaMemcpyDefault) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    return true;
  }
};

struct OutChunk {
  Buf buf;
  size_t len = 0, rd = 0;
};

}  // namespace

struct bzb200_enc {
  int level = 9;
  std::vector<int> devs;     // one entry per GPU; {-1} = the current device
  bzb200_pool* pool = nullptr;
  size_t window = (size_t)256 << 20;
  size_t first_window = (size_t)64 << 20;  // the first window of a stream is smaller: the GPU starts sooner
  uint64_t submitted = 0;                  // windows handed to the worker in this stream
  int mem_dev = -2;                        // -2 undecided, -1 host windows, >= 0 device windows on that GPU
  cudaStream_t copy_stream = nullptr;
  // the window being filled: data in win[cur][lo, hi); room in front of lo for the carried open block
  Buf win[2];
  int cur = 0;
  size_t lo = 0, hi = 0;
  // worker
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  bool has_task = false, busy = false, quit = false, th_started = false;
  int t_buf = 0;
  size_t t_lo = 0, t_hi = 0;
  bool t_final = false;
  size_t tail_lo = 0, tail_hi = 0;  // result of the last task: the open block's input inside its buffer
  int tail_buf = -1;
  bool last_closed_none = false;
  // stream state (touched by the worker while busy, by the caller otherwise)
  bool started = false, finished = false;
  uint8_t carry = 0;
  uint32_t carry_bits = 0, combined = 0;
  uint64_t blocks = 0, windows = 0, total_in = 0;
  std::deque<OutChunk*> outq;
  std::vector<OutChunk*> spare;
  int rc = BZB200_OK;
  std::string err;
};

namespace {

// Decides where the windows live (first write / finish, caller's thread).
void memory_setup(bzb200_enc* e) {
  if (e->mem_dev != -2) return;
  e->mem_dev = -1;
  if (e->devs.size() != 1 || pool_contexts_per_gpu() != 1) return;  // d_in/d_out are for one-worker pools
  if (const char* v = getenv("BZB200_ENC_DEVICE_WINDOWS")) {
    if (atoi(v) == 0) return;
  }
  int dev = e->devs[0];
  if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return; }
  DeviceScope ds(dev);
  if (!ds.ok) return;
  if (cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return; }
  e->devs[0] = dev;  // the worker's engine uses the same device
  e->mem_dev = dev;
  for (Buf& b : e->win) {
    b.dev = dev;
    b.cs = e->copy_stream;
  }
}

OutChunk* get_chunk(bzb200_enc* e, size_t bytes) {
  OutChunk* c = nullptr;
  {
    std::lock_guard<std::mutex> g(e->mu);
    for (size_t i = 0; i < e->spare.size(); ++i)
      if (e->spare[i]->buf.cap >= bytes) {
        c = e->spare[i];
        e->spare.erase(e->spare.begin() + i);
        break;
      }
    if (!c && !e->spare.empty()) {
      c = e->spare.back();
      e->spare.pop_back();
    }
  }
  if (!c) {
    c = new OutChunk();
    c->buf.dev = e->mem_dev >= 0 ? e->mem_dev : -1;
    c->buf.cs = e->copy_stream;
  }
  if (c->buf.cap < bytes && !c->buf.reserve(bytes + bytes / 8)) {
    c->buf.release();
    delete c;
    return nullptr;
  }
  c->len = c->rd = 0;
  return c;
}

void put_bits_msb(uint8_t* out, uint64_t pos, uint64_t value, int len) {
  for (int i = len - 1; i >= 0; --i, ++pos)
    if ((value >> i) & 1) out[pos >> 3] |= (uint8_t)(0x80u >> (pos & 7));
}

// Worker side: one window through the engine.
int run_window(bzb200_enc* e, int buf, size_t lo, size_t hi, bool final) {
  if (!e->pool) {
    std::vector<int> devs = e->devs;
    if (devs.size() == 1 && devs[0] < 0) {
      int d = 0;
      if (cudaGetDevice(&d) != cudaSuccess) {
        e->err = std::string("cudaGetDevice: ") + cudaGetErrorString(cudaGetLastError());
        return BZB200_E_CUDA;
      }
      devs[0] = d;
    }
    const int r = bzb200_pool_create((int)devs.size(), devs.data(), &e->pool);
    if (r != BZB200_OK) {
      e->err = r == BZB200_E_CUDA ? std::string("no usable CUDA device: ") + cudaGetErrorString(cudaGetLastError())
                                  : "pool creation failed";
      e->pool = nullptr;
      return r;
    }
  }
  const bool on_dev = e->mem_dev >= 0;
  const size_t n = hi - lo;
  if (n == 0) {  // only possible for an empty stream: header + trailer (encoder.rs:224-291 with nblock == 0)
    uint8_t tmp[16];
    memset(tmp, 0, sizeof(tmp));
    uint64_t bit = 0;
    if (!e->started) {
      tmp[0] = 'B'; tmp[1] = 'Z'; tmp[2] = 'h'; tmp[3] = (uint8_t)('0' + e->level);
      bit = 32;
      e->started = true;
    } else if (e->carry_bits) {
      tmp[0] = e->carry;
      bit = e->carry_bits;
    }
    put_bits_msb(tmp, bit, 0x177245385090ull, 48);
    put_bits_msb(tmp, bit + 48, e->combined, 32);
    OutChunk* c = get_chunk(e, 64);
    if (!c || !c->buf.put(0, tmp, 16)) return BZB200_E_INTERNAL;
    c->len = (size_t)((bit + 80 + 7) / 8);
    e->carry_bits = 0;
    std::lock_guard<std::mutex> g(e->mu);
    e->outq.push_back(c);
    return BZB200_OK;
  }
  const size_t cap = bzb200_max_output_bytes(e->level, n) + 64;
  OutChunk* c = get_chunk(e, cap);
  if (!c) {
    e->err = "out of memory for the output of a window";
    return BZB200_E_INTERNAL;
  }
  SpanJob J;
  J.level = e->level;
  J.n = n;
  J.first = !e->started;
  J.final = final;
  J.carry = e->carry;
  J.carry_bits = e->started ? e->carry_bits : 0;
  J.combined = e->combined;
  J.cap = c->buf.cap;
  if (on_dev) {
    J.d_in = e->win[buf].p + lo;
    J.d_out = c->buf.p;
  } else {
    J.h_in = e->win[buf].p + lo;
    J.h_out = c->buf.p;
  }
  const int r = pool_run_span(e->pool, &J);
  if (r != BZB200_OK) {
    e->err = pool_error(e->pool);
    std::lock_guard<std::mutex> g(e->mu);
    e->spare.push_back(c);
    return r;
  }
  e->started = true;
  e->combined = J.combined;
  e->blocks += J.blocks;
  e->windows += 1;
  if (final) {
    c->len = (size_t)((J.end_bits + 7) / 8);
    e->carry_bits = 0;
    e->tail_buf = -1;
  } else {
    c->len = (size_t)(J.end_bits / 8);
    e->carry_bits = (uint32_t)(J.end_bits & 7);
    e->carry = e->carry_bits ? J.last_byte : 0;
    e->tail_buf = buf;
    e->tail_lo = lo + (size_t)J.consumed;
    e->tail_hi = hi;
    e->last_closed_none = J.blocks == 0;
  }
  std::lock_guard<std::mutex> g(e->mu);
  if (c->len) e->outq.push_back(c);
  else e->spare.push_back(c);
  return BZB200_OK;
}

void enc_worker(bzb200_enc* e) {
  for (;;) {
    int buf;
    size_t lo, hi;
    bool final;
    {
      std::unique_lock<std::mutex> lk(e->mu);
      e->cv.wait(lk, [&] { return e->quit || e->has_task; });
      if (e->quit) return;
      buf = e->t_buf; lo = e->t_lo; hi = e->t_hi; final = e->t_final;
      e->has_task = false;
    }
    int r;
    try {
      r = run_window(e, buf, lo, hi, final);
    } catch (const std::exception&) {
      r = BZB200_E_INTERNAL;
      e->err = "out of host memory";
    }
    {
      std::lock_guard<std::mutex> lk(e->mu);
      if (r != BZB200_OK && e->rc == BZB200_OK) e->rc = r;
      e->busy = false;
    }
    e->cv.notify_all();
  }
}

void wait_idle(bzb200_enc* e) {
  std::unique_lock<std::mutex> lk(e->mu);
  e->cv.wait(lk, [&] { return !e->busy; });
}

// Room for the carried open block in front of a window: a level-9 block of text spans ~0.9 MB of input; longer spans
// (runs) are handled by moving the data (submit).
size_t front_room(const bzb200_enc* e) { return std::min<size_t>(e->window, (size_t)e->level * 200000) + 4096; }

// Caller side: hands win[cur][lo,hi) to the worker (after the previous window is done and its open block has been
// put in front of this one) and switches to the other buffer.
int submit(bzb200_enc* e, bool final) {
  wait_idle(e);
  if (e->rc != BZB200_OK) return e->rc;
  if (e->tail_buf >= 0 && e->tail_hi > e->tail_lo) {  // carried open block: goes in front of the current data
    const size_t tail = e->tail_hi - e->tail_lo;
    Buf& B = e->win[e->cur];
    if (e->lo < tail) {  // not enough room in front: move the data up (rare: a window smaller than a block)
      const size_t shift = tail - e->lo + (tail >> 1);
      if (!B.reserve(B.cap + shift + 64, e->lo, e->hi, shift)) return BZB200_E_INTERNAL;
      e->lo += shift;
      e->hi += shift;
    }
    if (!B.take(e->win[e->tail_buf], e->tail_lo, e->tail_hi, e->lo - tail)) return BZB200_E_CUDA;
    e->lo -= tail;
    if (e->last_closed_none) e->window = std::max(e->window * 2, (e->hi - e->lo) + 1);  // let the window grow
  }
  e->tail_buf = -1;
  if (!e->th_started) {
    e->th = std::thread(enc_worker, e);
    e->th_started = true;
  }
  {
    std::lock_guard<std::mutex> lk(e->mu);
    e->t_buf = e->cur;
    e->t_lo = e->lo;
    e->t_hi = e->hi;
    e->t_final = final;
    e->has_task = true;
    e->busy = true;
  }
  e->cv.notify_all();
  ++e->submitted;
  e->cur ^= 1;
  e->lo = e->hi = 0;  // set up by the next write
  return BZB200_OK;
}

}  // namespace

extern "C" {

static int enc_create(int level, const int* devices, int ngpus, bzb200_enc** out) {
  if (!out) return BZB200_E_ARG;
  *out = nullptr;
  if (level < 1 || level > 9) return BZB200_E_LEVEL;  // BZip2Encoder::new panics "invalid level"
  if (ngpus < 1 || ngpus > 64) return BZB200_E_ARG;
  bzb200_enc* e = new bzb200_enc();
  e->level = level;
  for (int g = 0; g < ngpus; ++g) e->devs.push_back(devices ? devices[g] : g);
  e->window = ((size_t)256 << 20) * (size_t)ngpus;
  if (const char* w = getenv("BZB200_ENC_WINDOW")) {
    unsigned long long v = strtoull(w, nullptr, 10);
    if (v >= 1) e->window = (size_t)v;
  }
  e->first_window = ((size_t)64 << 20) * (size_t)ngpus;
  if (const char* w = getenv("BZB200_ENC_FIRST_WINDOW")) {
    unsigned long long v = strtoull(w, nullptr, 10);
    if (v >= 1) e->first_window = (size_t)v;
  }
  *out = e;
  return BZB200_OK;
}

int bzb200_enc_create(int level, int device, bzb200_enc** out) { return enc_create(level, &device, 1, out); }

int bzb200_enc_create_multi(int level, int ngpus, const int* devices, bzb200_enc** out) {
  return enc_create(level, devices, ngpus, out);
}

int bzb200_enc_write(bzb200_enc* e, const uint8_t* p, size_t n) {
  if (!e || (!p && n)) return BZB200_E_ARG;
  if (e->finished) {
    e->err = "write after finish";
    return BZB200_E_STATE;
  }
  if (e->rc != BZB200_OK) return e->rc;
  try {
    memory_setup(e);
    while (n) {
      Buf& B = e->win[e->cur];
      if (e->hi == e->lo && e->lo == 0) e->lo = e->hi = front_room(e);  // a fresh window
      const size_t wnd = e->submitted == 0 ? std::min(e->window, e->first_window) : e->window;
      const size_t want = e->lo + wnd;
      if (B.cap < e->lo + e->window + 64) {
        // first use (or a grown window): the buffer is sized once for a full window, not per write
        if (!B.reserve(e->lo + e->window + 64, e->lo, e->hi)) {
          e->err = B.dev >= 0 ? "out of device memory for the input window" : "out of host memory";
          return BZB200_E_INTERNAL;
        }
      }
      const size_t room = want > e->hi ? want - e->hi : 0;
      const size_t take = std::min(n, room);
      if (take) {
        if (!B.put(e->hi, p, take)) {
          e->err = std::string("copying input to the device: ") + cudaGetErrorString(cudaGetLastError());
          return BZB200_E_CUDA;
        }
        e->hi += take;
        e->total_in += take;
        p += take;
        n -= take;
      }
      if (e->hi >= want) {
        const int r = submit(e, false);
        if (r != BZB200_OK) return r;
      }
    }
  } catch (const std::exception&) {
    e->err = "out of host memory";
    return BZB200_E_INTERNAL;
  }
  return BZB200_OK;
}

int bzb200_enc_finish(bzb200_enc* e) {
  if (!e) return BZB200_E_ARG;
  if (e->finished) return BZB200_OK;
  if (e->rc != BZB200_OK) return e->rc;
  try {
    memory_setup(e);
    if (e->hi == e->lo && e->lo == 0) {  // nothing in the current buffer: it still has to carry the open block
      if (e->win[e->cur].cap < front_room(e) + 64 && !e->win[e->cur].reserve(front_room(e) + 64)) {
        if (e->win[e->cur].dev >= 0) {  // no usable device: fall back to a host buffer, the engine reports the error
          for (Buf& b : e->win) b.dev = -1;
          e->mem_dev = -1;
          if (!e->win[e->cur].reserve(front_room(e) + 64)) return BZB200_E_INTERNAL;
        } else {
          return BZB200_E_INTERNAL;
        }
      }
      e->lo = e->hi = front_room(e);
    }
    const int r = submit(e, true);
    if (r != BZB200_OK) return r;
  } catch (const std::exception&) {
    e->err = "out of host memory";
    return BZB200_E_INTERNAL;
  }
  wait_idle(e);
  if (e->rc != BZB200_OK) return e->rc;
  e->finished = true;
  return BZB200_OK;
}

size_t bzb200_enc_read(bzb200_enc* e, uint8_t* dst, size_t cap) {
  if (!e || !dst) return 0;
  size_t got = 0;
  for (;;) {
    OutChunk* c = nullptr;
    {
      std::lock_guard<std::mutex> g(e->mu);
      if (got >= cap || e->outq.empty()) break;
      c = e->outq.front();  // only this thread removes chunks, the worker only appends
    }
    const size_t n = std::min(cap - got, c->len - c->rd);
    if (n && !c->buf.get(c->rd, dst + got, n)) break;
    c->rd += n;
    got += n;
    if (c->rd == c->len) {
      std::lock_guard<std::mutex> g(e->mu);
      e->outq.pop_front();
      e->spare.push_back(c);
    }
  }
  return got;
}

size_t bzb200_enc_output_size(const bzb200_enc* e) {
  if (!e) return 0;
  bzb200_enc* m = const_cast<bzb200_enc*>(e);
  std::lock_guard<std::mutex> g(m->mu);
  size_t n = 0;
  for (const OutChunk* c : e->outq) n += c->len - c->rd;
  return n;
}

int bzb200_enc_reset(bzb200_enc* e) {
  if (!e) return BZB200_E_ARG;
  wait_idle(e);
  {
    std::lock_guard<std::mutex> g(e->mu);
    for (OutChunk* c : e->outq) e->spare.push_back(c);
    e->outq.clear();
  }
  e->lo = e->hi = 0;
  e->cur = 0;
  e->submitted = 0;
  e->tail_buf = -1;
  e->tail_lo = e->tail_hi = 0;
  e->last_closed_none = false;
  e->finished = false;
  e->started = false;
  e->carry_bits = 0;
  e->carry = 0;
  e->combined = 0;
  e->blocks = 0;
  e->windows = 0;
  e->total_in = 0;
  e->rc = BZB200_OK;
  e->err.clear();
  return BZB200_OK;
}

void bzb200_enc_destroy(bzb200_enc* e) {
  if (!e) return;
  if (e->th_started) {
    wait_idle(e);
    {
      std::lock_guard<std::mutex> lk(e->mu);
      e->quit = true;
    }
    e->cv.notify_all();
    e->th.join();
  }
  if (e->pool) bzb200_pool_destroy(e->pool);
  for (OutChunk* c : e->outq) { c->buf.release(); delete c; }
  for (OutChunk* c : e->spare) { c->buf.release(); delete c; }
  e->win[0].release();
  e->win[1].release();
  if (e->copy_stream) {
    DeviceScope ds(e->mem_dev);
    cudaStreamDestroy(e->copy_stream);
  }
  delete e;
}

const char* bzb200_enc_last_error(const bzb200_enc* e) { return e ? e->err.c_str() : "null encoder"; }

int bzb200_enc_stats(const bzb200_enc* e, uint64_t* out, size_t cap) {
  if (!e || !out) return BZB200_E_ARG;
  wait_idle(const_cast<bzb200_enc*>(e));  // the counters belong to the worker while a window is in flight
  const uint64_t v[4] = {e->blocks, e->windows, (uint64_t)(e->hi - e->lo) + (e->tail_buf >= 0 ? e->tail_hi - e->tail_lo : 0),
                         (uint64_t)bzb200_enc_output_size(e)};
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
