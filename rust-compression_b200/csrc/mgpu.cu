// mgpu.cu — the multi-GPU engine behind the C ABI (include/bzb200.h section 2c): one .bz2 stream compressed block-wise
// on several GPUs of one box from ONE process — what `BZip2Encoder::new(level)` binds when the shim is built for a
// multi-GPU box (bzb200_enc_create_multi), and the one-shot bzb200_pool_compress_host.
//
// replaces the sequential block loop of the reference (EncoderInner::next -> write_block, src/bzip2/encoder.rs:671-697,
// 224-291) by: one worker thread + context per GPU (BZB200_MG_CTX_PER_GPU contexts per GPU, default 1); every worker
//   1. copies ITS slice of the host input to its GPU (all PCIe links in parallel; nothing is all-gathered),
//   2. runs the sliced K1 plan (slice_plan.cu); the three exchanges between slices go through host memory,
//   3. pulls the tail of its last block from the host input, encodes the blocks that start in its slice (K2-K6),
//   4. shifts its bit string to the bit phase it has in the joined stream (K7 on its own GPU) and copies it D2H
//      straight to its byte offset in the caller's output buffer (again all links in parallel);
// the coordinating worker then ORs the few bytes two neighbours share, folds the block CRCs and writes the trailer with
// host bit operations: the bitstreams are concatenated at bit granularity on the host (BitWriter<Left>,
// src/bitio/writer.rs:186-242).  No collective is needed: the only data that crosses GPUs is a handful of words.
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "host_ctx.h"
#include "mgpu.h"

extern "C" {
int bzb200_slice_begin(bzb200_ctx* c, int level, uint64_t N, uint64_t lo, uint64_t hi, const uint8_t* d_lo,
                       uint64_t avail_hi, uint64_t reserve_hi, int64_t* last_head);
int bzb200_slice_counts(bzb200_ctx* c, int64_t carry_in, uint64_t* emitted);
int bzb200_slice_prefix(bzb200_ctx* c, uint64_t E_lo, uint64_t E_tot);
int bzb200_slice_windows(bzb200_ctx* c, uint64_t x0, uint64_t* j0, uint32_t* nj, const uint64_t** d_F);
int bzb200_cut_walk(const uint64_t* F, uint64_t K, uint32_t T, uint64_t Etot, uint64_t N, uint32_t max_blocks,
                    uint64_t* state, uint64_t* in_off, uint64_t* rle_off, uint32_t* nblocks, uint32_t* max_block_len);
int bzb200_slice_set_blocks(bzb200_ctx* c, uint32_t nblocks, const uint64_t* in_off, const uint64_t* rle_off,
                            uint32_t max_block_len);
int bzb200_slice_blocks(const bzb200_ctx* c, uint32_t* b0, uint32_t* b1, uint64_t* need_hi);
int bzb200_slice_extend(bzb200_ctx* c, uint64_t avail_hi);
size_t bzb200_slice_halo_bytes(void);
uint32_t bzb200_cut_window(void);
}

namespace {

// Spinning barrier: the workers are dedicated threads and the waits are microseconds long.
struct SpinBarrier {
  std::atomic<uint32_t> count{0}, gen{0};
  uint32_t n = 1;
  void wait() {
    const uint32_t g = gen.load(std::memory_order_acquire);
    if (count.fetch_add(1, std::memory_order_acq_rel) + 1 == n) {
      count.store(0, std::memory_order_relaxed);
      gen.fetch_add(1, std::memory_order_release);
    } else {
      uint32_t spins = 0;
      while (gen.load(std::memory_order_acquire) == g)
        if (++spins > 2000) std::this_thread::yield();
    }
  }
};

constexpr size_t LEFT = 256;  // margin in front of the slice in the staging buffer (16 bytes of it are filled)

}  // namespace

struct bzb200_pool {
  int n = 0;
  std::vector<int> dev;
  std::vector<bzb200_ctx*> ctx;
  std::vector<std::thread> th;
  std::mutex mu;
  std::condition_variable cv_job, cv_done;
  uint64_t job_seq = 0;
  int running = 0;
  bool quit = false;
  SpanJob* job = nullptr;
  SpinBarrier bar;
  std::atomic<int> failed{0};
  std::mutex err_mu;
  std::string err;
  int rc = BZB200_OK;
  // shared between the workers during a job
  std::vector<int64_t> last_head;
  std::vector<uint64_t> emitted, lo, hi, bits;
  std::vector<uint32_t> b0, b1;
  std::vector<uint64_t> F, in_off, rle_off;
  std::vector<uint32_t> crc;
  uint64_t state[4] = {0, 0, 0, 0};
  uint64_t Etot = 0, phase_K = 0;
  uint32_t nblocks = 0, max_block_len = 0, max_blocks = 0, phases = 0;
  uint8_t* edge = nullptr;  // pinned: first and last byte of every worker's shifted bit string
  std::vector<DevBuf> shift;  // per worker: bit string moved to its phase in the joined stream
  uint64_t stat_spans = 0, stat_blocks = 0, stat_phases = 0;
};

namespace {

void pool_fail(bzb200_pool* p, int w, int code, const std::string& what) {
  std::lock_guard<std::mutex> g(p->err_mu);
  if (!p->failed.load()) {
    p->rc = code;
    p->err = "gpu worker " + std::to_string(w) + ": " + what;
    p->failed.store(1);
  }
}

// One phase of a job: runs fn unless an earlier phase failed anywhere, then meets the other workers.  Returns false
// when the job is to be abandoned (every worker sees the same answer after the barrier).
template <class Fn>
bool phase(bzb200_pool* p, int w, Fn fn) {
  if (!p->failed.load()) {
    bzb200_ctx* c = p->ctx[w];
    c->err.clear();
    const int r = fn();
    if (r != BZB200_OK) {
      if (r == BZB200_E_CUDA && c->err.empty()) c->err = std::string("cuda: ") + cudaGetErrorString(cudaGetLastError());
      pool_fail(p, w, r, c->err);
    }
  }
  p->bar.wait();
  return !p->failed.load();
}

void put_bits_host(uint8_t* out, uint64_t pos, uint64_t value, int len) {  // MSB first, ORs into zeroed bits
  for (int i = len - 1; i >= 0; --i, ++pos)
    if ((value >> i) & 1) out[pos >> 3] |= (uint8_t)(0x80u >> (pos & 7));
}

// One worker: no slicing — the plain one-context path (plan, encode, framing on the device), with the input either
// copied here or already resident (a streaming encoder copies every piece as it is written).
int run_span_single(bzb200_pool* p) {
  SpanJob& J = *p->job;
  bzb200_ctx* c = p->ctx[0];
  c->err.clear();
  TRY(set_device(c));
  const uint8_t* d_in = J.d_in;
  if (!d_in) {
    TRY(ensure(c, c->stage_in, J.n + 64));
    CK(c, cudaMemcpyAsync(c->stage_in.p, J.h_in, J.n, cudaMemcpyHostToDevice, c->stream));
    d_in = ptr<uint8_t>(c->stage_in);
  }
  uint32_t nb = 0;
  TRY(bzb200_plan(c, J.level, d_in, J.n, &nb));
  const uint32_t nenc = J.final ? nb : nb - 1;
  size_t cap = bzb200_max_output_bytes(J.level, nenc ? (size_t)c->h_in_off[nenc] : 0) + 64;
  uint8_t* d_out = J.d_out;
  if (d_out) {
    cap = std::min(cap, J.cap) & ~(size_t)3;
  } else {
    TRY(ensure(c, c->stage_out, cap));
    d_out = ptr<uint8_t>(c->stage_out);
  }
  CK(c, cudaMemsetAsync(d_out, 0, cap, c->stream));
  uint64_t bit = 0;
  if (J.first) {
    TRY(bzb200_write_stream_header(c, J.level, d_out, cap));
    bit = 32;
  } else if (J.carry_bits) {
    CK(c, cudaMemcpyAsync(d_out, &J.carry, 1, cudaMemcpyHostToDevice, c->stream));
    bit = J.carry_bits;
  }
  if (nenc) TRY(bzb200_encode_blocks(c, 0, nenc, d_out, cap, bit, &bit));
  J.combined = bzb200_combine_crc(J.combined, c->h_crc.data(), nenc);
  size_t bytes = (size_t)((bit + 7) / 8);
  if (J.final) {
    TRY(bzb200_write_stream_trailer(c, d_out, cap, bit, J.combined, &bytes));
    bit += 80;
  }
  if (bytes > J.cap) {
    c->err = "output buffer too small: need " + std::to_string(bytes) + " bytes";
    return BZB200_E_ARG;
  }
  if (!J.d_out && bytes) CK(c, cudaMemcpyAsync(J.h_out, d_out, bytes, cudaMemcpyDeviceToHost, c->stream));
  J.last_byte = 0;
  if (bytes) CK(c, cudaMemcpyAsync(&J.last_byte, d_out + bytes - 1, 1, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  J.end_bits = bit;
  J.consumed = nenc ? c->h_in_off[nenc] : 0;
  J.blocks = nenc;
  p->stat_spans += 1;
  p->stat_blocks += nenc;
  p->stat_phases += 1;
  return BZB200_OK;
}

void run_span(bzb200_pool* p, int w) {
  SpanJob& J = *p->job;
  bzb200_ctx* c = p->ctx[w];
  const int W = p->n;
  if (W == 1) {
    const int r = run_span_single(p);
    if (r != BZB200_OK) {
      if (r == BZB200_E_CUDA && c->err.empty()) c->err = std::string("cuda: ") + cudaGetErrorString(cudaGetLastError());
      pool_fail(p, 0, r, c->err);
    }
    return;
  }
  if (J.d_in || J.d_out || !J.h_in || !J.h_out) {  // the sliced path copies its slices from / to host memory
    pool_fail(p, w, BZB200_E_ARG, "device-resident spans need a pool of one worker");
    return;
  }
  const uint64_t N = J.n;
  const uint64_t tile = k1_tile_bytes();
  const uint64_t halo = bzb200_slice_halo_bytes();
  uint8_t* d_lo = nullptr;
  uint64_t lo = 0, hi = 0, avail = 0, reserve = 0;
  bool have = false;  // this worker owns a non-empty slice

  // ---- 1. slice -> device, last run heads
  if (!phase(p, w, [&]() -> int {
        TRY(set_device(c));
        const uint64_t per = ((N + W - 1) / W + tile - 1) / tile * tile;
        lo = std::min<uint64_t>(N, per * w);
        hi = std::min<uint64_t>(N, per * (w + 1));
        p->lo[w] = lo;
        p->hi[w] = hi;
        p->last_head[w] = -1;
        p->emitted[w] = 0;
        p->bits[w] = 0;
        p->b0[w] = p->b1[w] = 0;
        have = hi > lo;
        if (!have) return BZB200_OK;
        const uint64_t tail_max = (uint64_t)J.level * 100000ull * 51ull + 4 * tile;  // input bytes one block can span
        avail = std::min<uint64_t>(N, hi + halo);
        reserve = std::min<uint64_t>(N, hi + halo + tail_max);
        TRY(ensure(c, c->stage_in, LEFT + (reserve - lo) + 64));
        d_lo = ptr<uint8_t>(c->stage_in) + LEFT;
        const uint64_t left = lo ? 16 : 0;
        CK(c, cudaMemcpyAsync(d_lo - left, J.h_in + lo - left, (size_t)(avail - lo + left), cudaMemcpyHostToDevice, c->stream));
        int64_t lh = -1;
        TRY(bzb200_slice_begin(c, J.level, N, lo, hi, d_lo, avail, reserve, &lh));
        p->last_head[w] = lh;
        return BZB200_OK;
      }))
    return;
  // ---- 2. bytes every tile emits
  if (!phase(p, w, [&]() -> int {
        if (!have) return BZB200_OK;
        int64_t carry = -1;
        for (int v = 0; v < w; ++v) carry = std::max(carry, p->last_head[v]);
        uint64_t em = 0;
        TRY(bzb200_slice_counts(c, carry, &em));
        p->emitted[w] = em;
        return BZB200_OK;
      }))
    return;
  // ---- 3. emitted offsets; the coordinator prepares the chain
  if (!phase(p, w, [&]() -> int {
        uint64_t E_lo = 0, tot = 0;
        for (int v = 0; v < W; ++v) {
          if (v < w) E_lo += p->emitted[v];
          tot += p->emitted[v];
        }
        if (w == 0) {
          p->Etot = tot;
          const uint64_t emax = N + N / 4 + 64;
          const uint64_t mb = emax / ((uint64_t)J.level * 100000u - 19u) + 2;
          if (mb > 0x7FFFFFF0ull) return BZB200_E_ARG;
          p->max_blocks = (uint32_t)mb;
          p->in_off.assign(mb + 1, 0);
          p->rle_off.assign(mb + 1, 0);
          p->state[0] = p->state[1] = p->state[2] = p->state[3] = 0;
          p->nblocks = 0;
          p->phases = 0;
        }
        if (!have) return BZB200_OK;
        return bzb200_slice_prefix(c, E_lo, tot);
      }))
    return;
  // ---- 4. cut chain: window rows per slice -> host, the coordinator walks them; one pass unless the drift of the cut
  // positions leaves a window
  const uint32_t CW = bzb200_cut_window();
  for (;;) {
    if (!phase(p, w, [&]() -> int {  // size the shared table for this phase
          if (w != 0) return BZB200_OK;
          const uint64_t T = (uint64_t)J.level * 100000u - 19u;
          const uint64_t x0 = p->state[1];
          p->phase_K = p->Etot >= x0 + T ? (p->Etot - x0) / T : 0;
          if (p->F.size() < p->phase_K * CW) p->F.resize(p->phase_K * CW);
          ++p->phases;
          return BZB200_OK;
        }))
      return;
    if (!phase(p, w, [&]() -> int {
          if (!have) return BZB200_OK;
          uint64_t j0 = 0;
          uint32_t nj = 0;
          const uint64_t* d_F = nullptr;
          TRY(bzb200_slice_windows(c, p->state[1], &j0, &nj, &d_F));
          if (nj == 0) return BZB200_OK;
          if (j0 + nj > p->phase_K) return BZB200_E_INTERNAL;
          CK(c, cudaMemcpyAsync(p->F.data() + j0 * CW, d_F, (size_t)nj * CW * 8, cudaMemcpyDeviceToHost, c->stream));
          CK(c, cudaStreamSynchronize(c->stream));
          return BZB200_OK;
        }))
      return;
    if (!phase(p, w, [&]() -> int {
          if (w != 0) return BZB200_OK;
          const int r = bzb200_cut_walk(p->F.data(), p->phase_K, (uint32_t)J.level * 100000u - 19u, p->Etot, N, p->max_blocks,
                                        p->state, p->in_off.data(), p->rle_off.data(), &p->nblocks, &p->max_block_len);
          if (r != BZB200_OK) c->err = "cut chain: inconsistent window rows";
          if (r == BZB200_OK && !p->state[2] && p->phases > (1u << 20)) return BZB200_E_INTERNAL;
          return r;
        }))
      return;
    if (p->state[2]) break;
  }
  const uint32_t nb = p->nblocks;
  const uint32_t nenc = J.final ? nb : nb - 1;  // a span that is not the last one leaves its last block open
  // ---- 5. block table -> every slice; tail of the slice's last block; encode
  if (!phase(p, w, [&]() -> int {
        if (w == 0) p->crc.assign(nb, 0);
        if (!have) return BZB200_OK;
        TRY(bzb200_slice_set_blocks(c, nb, p->in_off.data(), p->rle_off.data(), p->max_block_len));
        uint32_t b0 = 0, b1 = 0;
        uint64_t need = 0;
        TRY(bzb200_slice_blocks(c, &b0, &b1, &need));
        b1 = std::min(b1, nenc);
        b0 = std::min(b0, b1);
        p->b0[w] = b0;
        p->b1[w] = b1;
        if (b1 > b0) {
          need = std::min<uint64_t>(N, (p->in_off[b1] + tile - 1) / tile * tile + (p->in_off[b1] < N ? 1 : 0));
          if (need > reserve) return BZB200_E_INTERNAL;
          if (need > avail) {
            CK(c, cudaMemcpyAsync(d_lo + (avail - lo), J.h_in + avail, (size_t)(need - avail), cudaMemcpyHostToDevice, c->stream));
            TRY(bzb200_slice_extend(c, need));
            avail = need;
          }
        }
        return BZB200_OK;
      }))
    return;
  if (!phase(p, w, [&]() -> int {
        if (!have) return BZB200_OK;
        const uint32_t b0 = p->b0[w], b1 = p->b1[w];
        const bool lead = b0 == 0;  // the worker whose bit string opens the span (header / carried bits in front)
        if (b1 == b0 && !(lead && w == 0)) return BZB200_OK;
        const uint64_t in_bytes = b1 > b0 ? p->in_off[b1] - p->in_off[b0] : 0;
        const size_t cap = bzb200_max_output_bytes(J.level, in_bytes) + 64;
        TRY(ensure(c, c->stage_out, cap));
        uint8_t* d_out = ptr<uint8_t>(c->stage_out);
        CK(c, cudaMemsetAsync(d_out, 0, cap, c->stream));
        uint64_t bit = 0;
        if (lead && w == 0) {
          if (J.first) {
            TRY(bzb200_write_stream_header(c, J.level, d_out, cap));
            bit = 32;
          } else if (J.carry_bits) {
            CK(c, cudaMemcpyAsync(d_out, &J.carry, 1, cudaMemcpyHostToDevice, c->stream));
            bit = J.carry_bits;
          }
        }
        if (b1 > b0) {
          TRY(bzb200_encode_blocks(c, b0, b1, d_out, cap, bit, &bit));
          for (uint32_t b = b0; b < b1; ++b) p->crc[b] = c->h_crc[b];
        }
        p->bits[w] = bit;
        return BZB200_OK;
      }))
    return;
  // ---- 6. every bit string to its place in the caller's buffer
  uint64_t P = 0, total_bits = 0;
  for (int v = 0; v < W; ++v) {
    if (v < w) P += p->bits[v];
    total_bits += p->bits[v];
  }
  const uint64_t out_bytes = (total_bits + (J.final ? 80 : 0) + 7) / 8;
  if (!phase(p, w, [&]() -> int {
        if (w == 0 && out_bytes > J.cap) {
          c->err = "output buffer too small: need " + std::to_string(out_bytes) + " bytes";
          return BZB200_E_ARG;
        }
        return BZB200_OK;
      }))
    return;
  if (!phase(p, w, [&]() -> int {
        const uint64_t nbits = p->bits[w];
        p->edge[2 * w] = p->edge[2 * w + 1] = 0;
        if (nbits == 0) return BZB200_OK;
        const uint32_t ph = (uint32_t)(P & 7);
        const uint64_t nbytes = (ph + nbits + 7) / 8;
        const uint8_t* src = ptr<uint8_t>(c->stage_out);
        if (ph) {  // K7 on this GPU: the bit string at the phase it has in the joined stream
          const size_t scap = ((ph + nbits + 31) / 32) * 4 + 64;
          TRY(ensure(c, p->shift[w], scap));
          CK(c, cudaMemsetAsync(p->shift[w].p, 0, scap, c->stream));
          TRY(bzb200_bit_append(c, ptr<uint8_t>(p->shift[w]), scap, ph, src, nbits));
          src = ptr<uint8_t>(p->shift[w]);
        }
        uint8_t* dst = J.h_out + (P >> 3);
        // the first and the last byte may be shared with a neighbour: they travel separately and are OR-ed on the host
        CK(c, cudaMemcpyAsync(p->edge + 2 * w, src, 1, cudaMemcpyDeviceToHost, c->stream));
        if (nbytes > 1) CK(c, cudaMemcpyAsync(p->edge + 2 * w + 1, src + nbytes - 1, 1, cudaMemcpyDeviceToHost, c->stream));
        if (nbytes > 2) CK(c, cudaMemcpyAsync(dst + 1, src + 1, (size_t)(nbytes - 2), cudaMemcpyDeviceToHost, c->stream));
        CK(c, cudaStreamSynchronize(c->stream));
        return BZB200_OK;
      }))
    return;
  // ---- 7. the coordinator joins: shared bytes, combined CRC, trailer
  phase(p, w, [&]() -> int {
    if (w != 0) return BZB200_OK;
    uint64_t pos = 0;
    uint64_t written = 0;  // bytes of h_out that hold final or partial data so far
    for (int v = 0; v < W; ++v) {
      const uint64_t nbits = p->bits[v];
      if (!nbits) continue;
      const uint32_t ph = (uint32_t)(pos & 7);
      const uint64_t nbytes = (ph + nbits + 7) / 8;
      const uint64_t at = pos >> 3;
      if (ph && at < written) J.h_out[at] |= p->edge[2 * v];
      else J.h_out[at] = p->edge[2 * v];
      if (nbytes > 1) J.h_out[at + nbytes - 1] = p->edge[2 * v + 1];
      written = at + nbytes;
      pos += nbits;
    }
    uint32_t comb = J.combined;
    for (uint32_t b = 0; b < nenc; ++b) comb = ((comb << 1) | (comb >> 31)) ^ p->crc[b];
    J.combined = comb;
    if (J.final) {
      for (uint64_t i = written; i < out_bytes; ++i) J.h_out[i] = 0;
      if (pos & 7) J.h_out[pos >> 3] &= (uint8_t)(0xFF00u >> (pos & 7));  // bits above pos are zero already (K7 masks)
      put_bits_host(J.h_out, pos, 0x177245385090ull, 48);                    // encoder.rs:279-286
      put_bits_host(J.h_out, pos + 48, comb, 32);                            // :287-289
      pos += 80;
    }
    J.end_bits = pos;
    J.last_byte = pos ? J.h_out[(pos - 1) >> 3] : 0;
    J.consumed = nenc ? p->in_off[nenc] : 0;
    J.blocks = nenc;
    p->stat_spans += 1;
    p->stat_blocks += nenc;
    p->stat_phases += p->phases;
    return BZB200_OK;
  });
}

void worker_main(bzb200_pool* p, int w) {
  cudaSetDevice(p->dev[w]);
  uint64_t seen = 0;
  for (;;) {
    {
      std::unique_lock<std::mutex> lk(p->mu);
      p->cv_job.wait(lk, [&] { return p->quit || p->job_seq != seen; });
      if (p->quit) return;
      seen = p->job_seq;
    }
    run_span(p, w);
    {
      std::lock_guard<std::mutex> lk(p->mu);
      if (--p->running == 0) p->cv_done.notify_all();
    }
  }
}

}  // namespace

// Runs one span on the pool (blocking).  Empty input: handled by the caller (no block, header + trailer only).
int pool_run_span(bzb200_pool* p, SpanJob* J) {
  if (!p || !J) return BZB200_E_ARG;
  p->failed.store(0);
  p->rc = BZB200_OK;
  p->err.clear();
  {
    std::lock_guard<std::mutex> lk(p->mu);
    p->job = J;
    p->running = p->n;
    ++p->job_seq;
  }
  p->cv_job.notify_all();
  {
    std::unique_lock<std::mutex> lk(p->mu);
    p->cv_done.wait(lk, [&] { return p->running == 0; });
  }
  return p->failed.load() ? p->rc : BZB200_OK;
}

const std::string& pool_error(const bzb200_pool* p) { return p->err; }

int pool_contexts_per_gpu() {
  int per = 1;
  if (const char* e = getenv("BZB200_MG_CTX_PER_GPU")) per = std::max(1, std::min(8, atoi(e)));
  return per;
}

extern "C" {

int bzb200_pool_create(int ngpus, const int* devices, bzb200_pool** out) {
  if (!out || ngpus < 1 || ngpus > 64) return BZB200_E_ARG;
  *out = nullptr;
  const int per = pool_contexts_per_gpu();
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return BZB200_E_CUDA;
  bzb200_pool* p = new bzb200_pool();
  for (int g = 0; g < ngpus; ++g) {
    const int d = devices ? devices[g] : g;
    if (d < 0 || d >= ndev) {
      delete p;
      return BZB200_E_ARG;
    }
    for (int k = 0; k < per; ++k) p->dev.push_back(d);
  }
  p->n = (int)p->dev.size();
  // contexts of the same GPU sit next to each other, so their slices are neighbours
  for (int w = 0; w < p->n; ++w) {
    bzb200_ctx* c = nullptr;
    const int r = bzb200_ctx_create_impl(p->dev[w], nullptr, true, &c);
    if (r != BZB200_OK) {
      if (c) bzb200_ctx_destroy(c);
      for (bzb200_ctx* x : p->ctx) bzb200_ctx_destroy(x);
      delete p;
      return r;
    }
    p->ctx.push_back(c);
  }
  if (cudaHostAlloc((void**)&p->edge, (size_t)p->n * 2 + 16, cudaHostAllocPortable) != cudaSuccess) {
    for (bzb200_ctx* x : p->ctx) bzb200_ctx_destroy(x);
    delete p;
    return BZB200_E_CUDA;
  }
  p->last_head.assign(p->n, -1);
  p->emitted.assign(p->n, 0);
  p->lo.assign(p->n, 0);
  p->hi.assign(p->n, 0);
  p->bits.assign(p->n, 0);
  p->b0.assign(p->n, 0);
  p->b1.assign(p->n, 0);
  p->shift.resize(p->n);
  p->bar.n = (uint32_t)p->n;
  for (int w = 0; w < p->n; ++w) p->th.emplace_back(worker_main, p, w);
  *out = p;
  return BZB200_OK;
}

void bzb200_pool_destroy(bzb200_pool* p) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> lk(p->mu);
    p->quit = true;
  }
  p->cv_job.notify_all();
  for (auto& t : p->th) t.join();
  for (int w = 0; w < p->n; ++w) {
    cudaSetDevice(p->dev[w]);
    if (p->shift[w].p) cudaFree(p->shift[w].p);
    bzb200_ctx_destroy(p->ctx[w]);
  }
  if (p->edge) cudaFreeHost(p->edge);
  delete p;
}

int bzb200_pool_size(const bzb200_pool* p) { return p ? p->n : 0; }
}  // extern "C"

int pool_device(const bzb200_pool* p, int w) { return p->dev[w]; }

extern "C" {
const char* bzb200_pool_last_error(const bzb200_pool* p) { return p ? p->err.c_str() : "null pool"; }

int bzb200_pool_stats(const bzb200_pool* p, uint64_t* out, size_t cap) {
  if (!p || !out) return BZB200_E_ARG;
  uint64_t launches = 0;
  for (bzb200_ctx* c : p->ctx) launches += c->L.launches;
  const uint64_t v[4] = {p->stat_spans, p->stat_blocks, p->stat_phases, launches};
  for (size_t i = 0; i < cap && i < 4; ++i) out[i] = v[i];
  return BZB200_OK;
}

int bzb200_pool_compress_host(bzb200_pool* p, int level, const uint8_t* h_in, size_t n, uint8_t* h_out, size_t cap_bytes,
                              size_t* out_n) {
  if (!p || !h_out || !out_n || (!h_in && n)) return BZB200_E_ARG;
  *out_n = 0;
  if (level < 1 || level > 9) {
    p->err = "invalid level";
    return BZB200_E_LEVEL;
  }
  if (n == 0) {  // header + trailer, no block (encoder.rs:224-291 with nblock == 0)
    if (cap_bytes < 14) return BZB200_E_ARG;
    memset(h_out, 0, 14);
    h_out[0] = 'B'; h_out[1] = 'Z'; h_out[2] = 'h'; h_out[3] = (uint8_t)('0' + level);
    put_bits_host(h_out, 32, 0x177245385090ull, 48);
    *out_n = 14;
    return BZB200_OK;
  }
  SpanJob J;
  J.level = level;
  J.h_in = h_in;
  J.n = n;
  J.h_out = h_out;
  J.cap = cap_bytes;
  const int r = pool_run_span(p, &J);
  if (r != BZB200_OK) return r;
  *out_n = (size_t)((J.end_bits + 7) / 8);
  return BZB200_OK;
}

}  // extern "C"
