/* bzb200.h — C ABI of the B200-native bzip2 block-compression path.
 *
 * Drop-in boundary for chalharu/rust-compression's `BZip2Encoder`
 * (reference: src/bzip2/encoder.rs).  The reference has no FFI layer; its
 * boundary is the Rust API `BZip2Encoder::new(level)` + `Encoder::next(iter,
 * action)` (encoder.rs:58-72,116-158; traits/encoder.rs:81-93).  A Rust shim
 * that keeps that API and binds the functions below is shown in INTEGRATION.md
 * and shipped as source in rust-compression_b200/rust/.
 *
 * All entry points are plain C: pointers and sizes only, no CUDA/torch types
 * (a CUDA stream is passed as an opaque void*).  Every function returns
 * BZB200_OK (0) or a negative error code; nothing aborts the process.  There is
 * no CPU fallback: without a CUDA device every compute entry point fails with
 * BZB200_E_CUDA and bzb200_*last_error() carries the CUDA message.
 *
 * One object = one host thread.  Different objects may be used from different
 * threads concurrently.
 */
#ifndef BZB200_H
#define BZB200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define BZB200_API __attribute__((visibility("default")))
#else
#define BZB200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define BZB200_OK 0
#define BZB200_E_LEVEL (-1)    /* level outside 1..9 — the reference panics "invalid level" (encoder.rs:59-61) */
#define BZB200_E_CUDA (-2)     /* CUDA runtime/driver error; maps to CompressionError::Unexpected (error.rs:14) */
#define BZB200_E_ARG (-3)      /* null pointer / bad range / output capacity too small */
#define BZB200_E_STATE (-4)    /* call order violated (e.g. write after finish) */
#define BZB200_E_INTERNAL (-5) /* device-side invariant failed (reported, never silently ignored) */

/* ------------------------------------------------------------------------
 * 1. Streaming encoder — what the Rust shim's `BZip2Encoder` binds.
 *    replaces: BZip2Encoder::new            src/bzip2/encoder.rs:58-72
 *              Encoder::next + Action::Run  src/bzip2/encoder.rs:74-114,120-158 (input side)
 *              Action::Finish               src/bzip2/encoder.rs:99-105,729-739
 *              output byte iterator         src/bzip2/encoder.rs:153-157
 *    Action::Flush is out of contract: the shim maps it to Run.
 * ------------------------------------------------------------------------ */
typedef struct bzb200_enc bzb200_enc;

/* level 1..9 (else BZB200_E_LEVEL); device = CUDA ordinal, -1 = current device. */
BZB200_API int bzb200_enc_create(int level, int device, bzb200_enc** out);
/* Action::Run: append n input bytes (copied into pinned host staging). */
BZB200_API int bzb200_enc_write(bzb200_enc* e, const uint8_t* p, size_t n);
/* Action::Finish: compress everything written so far into one .bz2 stream. */
BZB200_API int bzb200_enc_finish(bzb200_enc* e);
/* Drain output bytes; returns the number copied, 0 after the last byte (-> `None`). */
BZB200_API size_t bzb200_enc_read(bzb200_enc* e, uint8_t* dst, size_t cap);
/* Total size of the finished stream (valid after finish). */
BZB200_API size_t bzb200_enc_output_size(const bzb200_enc* e);
/* Re-arm for a new stream at the same level (the reference resets its latches when it returns None,
 * encoder.rs:87-90,130-133). */
BZB200_API int bzb200_enc_reset(bzb200_enc* e);
BZB200_API void bzb200_enc_destroy(bzb200_enc* e);
BZB200_API const char* bzb200_enc_last_error(const bzb200_enc* e);

/* One-shot host->host: `iter.encode(&mut BZip2Encoder::new(level), Action::Finish).collect()`
 * (src/lib.rs:13-33 doctest).  *out is malloc'ed; free with bzb200_free. */
BZB200_API int bzb200_compress(int level, int device, const uint8_t* in, size_t n, uint8_t** out, size_t* out_n);
BZB200_API void bzb200_free(void* p);

/* ------------------------------------------------------------------------
 * 2. Device-resident job API — block-wise sharding across GPUs and the
 *    HBM-resident benchmark.  All d_* pointers are device pointers on the
 *    context's device; work is enqueued on the context's stream.
 * ------------------------------------------------------------------------ */
typedef struct bzb200_ctx bzb200_ctx;

/* stream: a cudaStream_t passed as void*; NULL = the device's default stream (what torch calls the default current
 * stream).  All work of the context is enqueued on that stream, so the caller's earlier work on it is ordered before. */
BZB200_API int bzb200_ctx_create(int device, void* stream, bzb200_ctx** out);
BZB200_API void bzb200_ctx_destroy(bzb200_ctx* c);
BZB200_API const char* bzb200_last_error(const bzb200_ctx* c);
/* cudaStreamSynchronize on the context's stream. */
BZB200_API int bzb200_sync(bzb200_ctx* c);

/* K1+K5: RLE1 run collapsing, greedy block cutting (T = level*100000-19), per-block CRC and in-use maps
 * over the whole input.  replaces EncoderInner::next/write_rle (encoder.rs:671-716), the cut test
 * (:692-696) and crc32::Digest (crc32.rs:82-84,129-131).  Synchronises once (block count -> host). */
BZB200_API int bzb200_plan(bzb200_ctx* c, int level, const uint8_t* d_in, size_t n, uint32_t* nblocks);
/* The same plan in four steps, for a sharded caller: K1's per-tile summaries (last run head, emitted bytes) are
 * computed by the rank that owns the tile range [t0,t1) into CALLER-owned device arrays of `ntiles` entries, the
 * caller exchanges the ranges between ranks (rust-compression_b200/sharded.py: NCCL all-gather), and every rank
 * finishes with the cheap global part (prefix sum + cut chain).  A tile is bzb200_plan_tile_bytes() input bytes.
 *   begin  : binds level/input, sizes the buffers, returns ntiles
 *   heads  : d_tile_head[t0..t1) = index of the last run head inside the tile (-1: none)
 *   counts : needs d_tile_head complete; d_tile_cnt[t0..t1) = RLE1 bytes the tile emits
 *   finish : needs d_tile_cnt complete; block cuts -> host; afterwards identical to bzb200_plan */
BZB200_API size_t bzb200_plan_tile_bytes(void);
BZB200_API int bzb200_plan_begin(bzb200_ctx* c, int level, const uint8_t* d_in, size_t n, uint64_t* ntiles);
BZB200_API int bzb200_plan_heads(bzb200_ctx* c, uint64_t t0, uint64_t t1, int64_t* d_tile_head);
BZB200_API int bzb200_plan_counts(bzb200_ctx* c, const int64_t* d_tile_head, uint64_t t0, uint64_t t1, uint32_t* d_tile_cnt);
BZB200_API int bzb200_plan_finish(bzb200_ctx* c, const uint32_t* d_tile_cnt, uint32_t* nblocks);
/* Number of blocks of the current plan (0 before any plan). */
BZB200_API uint32_t bzb200_num_blocks(const bzb200_ctx* c);
/* Block table of the current plan: in_off[nblocks+1] (input byte offsets), rle_off[nblocks+1] (offsets into the
 * RLE1 stream), crc[nblocks].  Any pointer may be NULL.  The RLE1 bytes, CRCs and in-use maps of a block are
 * produced when the block is encoded (bzb200_encode_blocks); asking for crc here computes the CRCs of the blocks
 * this context has not encoded as well (one K5 launch over the whole input). */
BZB200_API int bzb200_block_table(bzb200_ctx* c, uint64_t* in_off, uint64_t* rle_off, uint32_t* crc);
/* CRCs as they stand: valid for the blocks encoded by this context, 0 elsewhere; no device work.  A sharded caller
 * exchanges these slices between ranks instead of recomputing them (rust-compression_b200/sharded.py). */
BZB200_API int bzb200_block_crcs(const bzb200_ctx* c, uint32_t* crc, size_t cap);

/* K2-K6 for blocks [b0,b1) of the current plan: BWT (prefix-doubling rotation sort), MTF + RUNA/RUNB,
 * Huffman table selection/refinement, header + symbol bit packing.  replaces write_blockdata
 * (encoder.rs:300-639) and the per-block part of write_block (:253-277).  The blocks' bit strings are
 * written back to back, MSB first, into d_out starting at bit `start_bit`; bytes d_out[start_bit/8 ..] must be
 * zero on entry (the library ORs into them).  *end_bit = bit position after the last block. Synchronises. */
BZB200_API int bzb200_encode_blocks(bzb200_ctx* c, uint32_t b0, uint32_t b1, uint8_t* d_out, size_t cap_bytes,
                         uint64_t start_bit, uint64_t* end_bit);

/* K7: OR the first nbits bits of d_src (MSB first, starting at bit 0) into d_dst at bit offset dst_bit.
 * replaces BitWriter<Left>::write_bits across block/shard boundaries (bitio/writer.rs:186-224). */
BZB200_API int bzb200_bit_append(bzb200_ctx* c, uint8_t* d_dst, size_t dst_cap_bytes, uint64_t dst_bit, const uint8_t* d_src,
                      uint64_t nbits);

/* Stream framing (write_block, encoder.rs:237-251,279-289). */
/* combined = rotl1(combined) ^ crc[i], folded left to right starting from `seed` (0 for a new stream). */
BZB200_API uint32_t bzb200_combine_crc(uint32_t seed, const uint32_t* crc, size_t n);
/* Writes 'B','Z','h','0'+level at bit 0 of d_out (32 bits). */
BZB200_API int bzb200_write_stream_header(bzb200_ctx* c, int level, uint8_t* d_out, size_t cap_bytes);
/* ORs the 48-bit end magic + 32-bit combined CRC at bit `at_bit`; *total_bytes = stream length after zero padding. */
BZB200_API int bzb200_write_stream_trailer(bzb200_ctx* c, uint8_t* d_out, size_t cap_bytes, uint64_t at_bit, uint32_t combined_crc,
                                size_t* total_bytes);

/* Upper bound on the output bytes for n input bytes at `level` (header + per-block worst case + trailer). */
BZB200_API size_t bzb200_max_output_bytes(int level, size_t n);

/* Whole stream on one device, device in -> device out (plan + encode all blocks + framing).
 * d_out must be zero-filled for at least bzb200_max_output_bytes(level, n) bytes. */
BZB200_API int bzb200_compress_device(bzb200_ctx* c, int level, const uint8_t* d_in, size_t n, uint8_t* d_out, size_t cap_bytes,
                           size_t* out_n);

/* Whole stream, HOST in -> HOST out on the context's device and stream: H2D copy, bzb200_compress_device, D2H
 * copy of exactly *out_n bytes, synchronised on return.  Device staging buffers live in the context and are reused
 * across calls.  Pinned (page-locked) host buffers make both copies asynchronous DMA; pageable buffers work too. */
BZB200_API int bzb200_compress_host(bzb200_ctx* c, int level, const uint8_t* h_in, size_t n, uint8_t* h_out, size_t cap_bytes,
                         size_t* out_n);

/* ------------------------------------------------------------------------
 * 3. Instrumentation (parity tests, bench roofline).  Not needed by a shim.
 * ------------------------------------------------------------------------ */
/* Stage dump of block b (absolute index in the current plan; must be inside the most recent
 * bzb200_encode_blocks batch).  field: */
enum {
  BZB200_F_RLE = 0,    /* u8  block bytes after RLE1                       (block_buf, encoder.rs:699-716) */
  BZB200_F_RANK = 1,   /* u32 rank[pos] = position of rotation pos in the sorted order (inverse of sais.rs:266 result) */
  BZB200_F_LAST = 2,   /* u8  BWT last column                              (encoder.rs:332-338) */
  BZB200_F_MTF = 3,    /* u16 RUNA/RUNB/sym+1/EOB stream                   (mtf_buffer, encoder.rs:344-358) */
  BZB200_F_FREQ = 4,   /* u32 mtf_freq[alpha]                              (encoder.rs:321,351,358) */
  BZB200_F_SEL = 5,    /* u8  selector[] of the 4th pass                   (encoder.rs:471) */
  BZB200_F_LEN0 = 6,   /* u8  [ngroups][alpha] initial tables, libbzip2 table order (encoder.rs:379-426) */
  BZB200_F_LEN1 = 7,   /*     ... after refinement pass 1..4               (encoder.rs:504-508) */
  BZB200_F_LEN2 = 8,
  BZB200_F_LEN3 = 9,
  BZB200_F_LEN4 = 10,
  BZB200_F_INFO = 11   /* u64[16]: in_start,in_end,nblock,crc,orig_ptr,mtf_count,alpha,ngroups,nselectors,
                          bit_start,bit_end,sort_rounds,periodic,in_use words packed in [12..15] */
};
/* Copies up to cap_elems elements to host memory; *count = number of elements available. */
BZB200_API int bzb200_debug_stage(bzb200_ctx* c, uint32_t block, int field, void* host_dst, size_t cap_elems, size_t* count);

/* Per-kernel device timing with CUDA events on the context's stream. on=1 starts/clears, on=0 stops. */
BZB200_API int bzb200_profile(bzb200_ctx* c, int on);
/* Number of distinct kernels recorded since profiling was enabled. */
BZB200_API int bzb200_profile_count(bzb200_ctx* c);
/* i-th record: kernel name, launches, total milliseconds. */
BZB200_API int bzb200_profile_get(bzb200_ctx* c, int i, const char** name, uint64_t* launches, double* total_ms);
/* Kernels launched by this context since creation (for bench.py's gpu_launches). */
BZB200_API uint64_t bzb200_launch_count(const bzb200_ctx* c);
/* Sort statistics of the last encode_blocks call: doubling rounds run, radix passes run, elements sorted. */
BZB200_API int bzb200_sort_stats(const bzb200_ctx* c, uint32_t* rounds, uint32_t* radix_passes, uint64_t* elems_sorted);
/* Work counters of the last encode_blocks call (bench.py turns them into algorithmic bytes), out[0..7]:
 * 0 doubling rounds, 1 radix pass launches, 2 rotations sorted (initial + unresolved entering each round),
 * 3 elements moved by radix passes (list length x 5, summed), 4 work-list entries of the in-CTA group sort,
 * 5 RLE1 bytes (= BWT elements), 6 MTF/RUNA/RUNB symbols emitted incl. EOB, 7 reserved. */
BZB200_API int bzb200_path_stats(const bzb200_ctx* c, uint64_t* out, size_t cap);

BZB200_API const char* bzb200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* BZB200_H */
