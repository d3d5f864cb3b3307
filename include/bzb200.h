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
#define BZB200_E_DATA (-6)     /* decoder: the input is not a valid .bz2 stream; the BZip2Error kind is reported next to it */

/* BZip2Error of the reference decoder (src/bzip2/error.rs:4-11), as reported by the decoder entry points. */
#define BZB200_BZ_OK 0
#define BZB200_BZ_DATA_ERROR 1             /* -> CompressionError::DataError   (error.rs:44-52) */
#define BZB200_BZ_DATA_ERROR_MAGIC_FIRST 2 /* -> CompressionError::DataError */
#define BZB200_BZ_DATA_ERROR_MAGIC 3       /* -> CompressionError::DataError */
#define BZB200_BZ_UNEXPECTED_EOF 4         /* -> CompressionError::UnexpectedEof */
#define BZB200_BZ_UNEXPECTED 5             /* -> CompressionError::Unexpected */

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
/* The same object over several GPUs of the box: every window of input is sharded block-wise over `ngpus` devices
 * (devices[ngpus], NULL = 0..ngpus-1) by the in-library engine of section 2c — one host thread per GPU inside the
 * library, slices copied over all PCIe links in parallel, bit strings joined at bit granularity on the host.  This is
 * what `BZip2Encoder::new(level)` binds on a multi-GPU box (SURVEY.md section 8(b),(e)); the stream is bit-identical
 * to the single-GPU one. */
BZB200_API int bzb200_enc_create_multi(int level, int ngpus, const int* devices, bzb200_enc** out);
/* Action::Run: append n input bytes (copied before the call returns) to the window being filled (256 MiB per GPU, the
 * first window of a stream 64 MiB per GPU; env BZB200_ENC_WINDOW, BZB200_ENC_FIRST_WINDOW).  A full window is handed
 * to the object's worker thread, which compresses every block that has already closed while the caller keeps writing
 * into the second window; their bytes become readable as soon as the window is done, and the still-open last block is
 * carried in front of the next window (the reference also yields a block as soon as it
 * closes, encoder.rs:91-107).  Only the concatenation of all bytes read is defined, not which call yields which.
 *   One GPU: the windows live in device memory — the call is one DMA from p into HBM when p is pinned memory (the
 *   driver's staged copy otherwise), the finished bytes stay on the device until bzb200_enc_read copies them into dst
 *   (BZB200_ENC_DEVICE_WINDOWS=0: host windows as below).
 *   Several GPUs: the windows are pinned host buffers (memcpy), from which every GPU's worker copies its slice. */
BZB200_API int bzb200_enc_write(bzb200_enc* e, const uint8_t* p, size_t n);
/* Action::Finish: compress the remaining blocks and append the stream trailer. */
BZB200_API int bzb200_enc_finish(bzb200_enc* e);
/* Drain output bytes; returns the number copied.  Before finish, 0 means "nothing ready yet, feed more input"
 * (-> `None` under Action::Run); after finish, 0 means the stream is complete (-> `None` under Action::Finish). */
BZB200_API size_t bzb200_enc_read(bzb200_enc* e, uint8_t* dst, size_t cap);
/* Output bytes ready to be read right now (after finish and before any read: the size of the whole remainder). */
BZB200_API size_t bzb200_enc_output_size(const bzb200_enc* e);
/* Re-arm for a new stream at the same level (the reference resets its latches when it returns None,
 * encoder.rs:87-90,130-133). */
BZB200_API int bzb200_enc_reset(bzb200_enc* e);
/* out[0..3]: blocks encoded so far, windows compressed, input bytes still buffered, output bytes ready.  Waits for the
 * window that is in flight, if any. */
BZB200_API int bzb200_enc_stats(const bzb200_enc* e, uint64_t* out, size_t cap);
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
 * 2b. Sliced plan — the K1 plan when a context holds only ONE SLICE of the stream (a rank of a sharded run, one GPU of
 *     the engine of 2c).  The reference cuts blocks in one sequential pass (encoder.rs:671-716, cut test :692-696);
 *     here every slice works on its own bytes and three tiny exchanges between the slices — done by the caller, NCCL
 *     all-gathers in rust-compression_b200/sharded.py, host memory in the engine — reproduce the same cuts:
 *       begin   -> exchange the last run head of every slice (1 word)   -> counts
 *       counts  -> exchange the bytes every slice emits (1 word)        -> prefix
 *       windows -> exchange the cut-window rows (2 KB per block)        -> bzb200_cut_walk on the host (repeat from
 *                  `windows` with the new start while the walk is not done: only when the drift of the cut positions
 *                  leaves a window) -> set_blocks
 *     A slice then encodes (bzb200_encode_blocks) the blocks that START inside it; the last of them ends in the next
 *     slice's bytes: bzb200_slice_blocks says how far the input must be resident, the caller copies that tail behind
 *     the slice and calls bzb200_slice_extend.  No O(input) array is exchanged.
 * ------------------------------------------------------------------------ */
/* Bytes behind the slice end that must be resident at bzb200_slice_begin (a cut window reaches that far at most). */
BZB200_API size_t bzb200_slice_halo_bytes(void);
/* Slice granularity: lo must be a multiple, hi too unless hi == N. */
BZB200_API size_t bzb200_plan_tile_bytes(void);
/* Binds the slice [lo, hi) of an N-byte stream.  d_lo = device address of input byte lo, 16-byte aligned; the bytes
 * [lo - 16 (if lo > 0), avail_hi) must be in device memory around it, avail_hi >= min(N, hi + halo); reserve_hi = how
 * far the resident part may grow later (bzb200_slice_extend; at most one block's input span past hi + halo).
 * *last_head = input index of the last run head inside [lo, hi) (-1: none).  Synchronises. */
BZB200_API int bzb200_slice_begin(bzb200_ctx* c, int level, uint64_t N, uint64_t lo, uint64_t hi, const uint8_t* d_lo,
                                  uint64_t avail_hi, uint64_t reserve_hi, int64_t* last_head);
/* carry_in = max of the last_head values of all slices in front of this one (-1: none).  *emitted = RLE1 bytes the
 * slice emits.  Synchronises. */
BZB200_API int bzb200_slice_counts(bzb200_ctx* c, int64_t carry_in, uint64_t* emitted);
/* E_lo = bytes emitted by all slices in front, E_tot = by the whole stream. */
BZB200_API int bzb200_slice_prefix(bzb200_ctx* c, uint64_t E_lo, uint64_t E_tot);
/* Cut windows of the phase that starts at emitted offset x0 whose centre x0 + (j+1) T falls into the slice: rows
 * j0 .. j0 + nj - 1, bzb200_cut_window() 64-bit entries each, in device memory owned by the context (*d_F, valid until
 * the next call; stream ordered). */
BZB200_API int bzb200_slice_windows(bzb200_ctx* c, uint64_t x0, uint64_t* j0, uint32_t* nj, const uint64_t** d_F);
BZB200_API uint32_t bzb200_cut_window(void);
/* HOST: follows the greedy cut chain (encoder.rs:692-696) through the rows F[0 .. K) of one phase.  state[4] = {blocks
 * cut so far, emitted offset where the open block starts, done, longest block} — all zero before the first phase; the
 * phase's x0 is state[1] on entry.  in_off / rle_off need max_blocks + 1 entries (max_blocks = (N + N/4 + 64)/T + 2).
 * When state[2] becomes 1 the table is complete: *nblocks, *max_block_len, in_off[nblocks] = N, rle_off[nblocks] = Etot. */
BZB200_API int bzb200_cut_walk(const uint64_t* F, uint64_t K, uint32_t T, uint64_t Etot, uint64_t N, uint32_t max_blocks,
                               uint64_t* state, uint64_t* in_off, uint64_t* rle_off, uint32_t* nblocks,
                               uint32_t* max_block_len);
/* Gives the context the block table every slice agreed on; afterwards bzb200_block_table / bzb200_encode_blocks /
 * bzb200_debug_stage work as after bzb200_plan, for blocks whose input is resident. */
BZB200_API int bzb200_slice_set_blocks(bzb200_ctx* c, uint32_t nblocks, const uint64_t* in_off, const uint64_t* rle_off,
                                       uint32_t max_block_len);
/* [*b0, *b1) = the blocks that start inside the slice; *need_hi = how far the input must be resident to encode them. */
BZB200_API int bzb200_slice_blocks(const bzb200_ctx* c, uint32_t* b0, uint32_t* b1, uint64_t* need_hi);
/* More input bytes have been copied behind the slice: the resident part now ends at avail_hi (<= reserve_hi). */
BZB200_API int bzb200_slice_extend(bzb200_ctx* c, uint64_t avail_hi);

/* ------------------------------------------------------------------------
 * 2c. Multi-GPU engine, one process: a pool of contexts (one worker thread per GPU inside the library) that compresses
 *     HOST buffers block-wise over several GPUs with the sliced plan of 2b.  Every GPU copies its own slice in and its
 *     own bit string out (all PCIe links in parallel); the bit strings are shifted to their bit phase on their GPU
 *     (K7), the bytes two neighbours share are OR-ed and the trailer written on the host — the bit-granular
 *     concatenation of BitWriter<Left> (bitio/writer.rs:186-242).  Bit-identical to bzb200_compress_host.
 *     Environment: BZB200_MG_CTX_PER_GPU (contexts and worker threads per GPU, default 1).
 * ------------------------------------------------------------------------ */
typedef struct bzb200_pool bzb200_pool;
/* devices[ngpus] = CUDA ordinals, NULL = 0 .. ngpus-1. */
BZB200_API int bzb200_pool_create(int ngpus, const int* devices, bzb200_pool** out);
BZB200_API void bzb200_pool_destroy(bzb200_pool* p);
/* Workers (contexts) of the pool. */
BZB200_API int bzb200_pool_size(const bzb200_pool* p);
BZB200_API const char* bzb200_pool_last_error(const bzb200_pool* p);
/* Whole stream, HOST in -> HOST out (pinned buffers make every copy asynchronous DMA); synchronised on return. */
BZB200_API int bzb200_pool_compress_host(bzb200_pool* p, int level, const uint8_t* h_in, size_t n, uint8_t* h_out,
                                         size_t cap_bytes, size_t* out_n);
/* out[0..3]: spans compressed, blocks encoded, cut-chain phases run, kernels launched by all workers. */
BZB200_API int bzb200_pool_stats(const bzb200_pool* p, uint64_t* out, size_t cap);

/* ------------------------------------------------------------------------
 * 3. Decoder — block-parallel bzip2 decompression of whole buffers.
 *    replaces: BZip2Decoder::new / Decoder::next  src/bzip2/decoder.rs:584-615
 *              BZip2DecoderBase::init_block       src/bzip2/decoder.rs:163-525
 *              get_next_lfm, BitDecodeService     src/bzip2/decoder.rs:527-581
 *              DecodeExt::decode                  src/traits/decoder.rs:14-43
 *    Accepts what the reference accepts (multi-stream buffers, any level, its limits decoder.rs:238,285,292,399,416,427;
 *    blocks with the `randomised` bit of bzip2 <= 0.9.0, decoder.rs:94-116,537-539; magics of which only the first byte
 *    is right — the reference reads the other bytes without comparing them, decoder.rs:155-161,177-182,211-224,495-508)
 *    and reports what it reports: the bytes the reference would have yielded before an error, then the BZip2Error kind.
 *    Deviations, malformed input only: over-subscribed coding tables are DataError (as in the restated reference
 *    decoder under oracle/); a block that ends in four equal bytes with no count byte is DataError (the reference reads
 *    past the block there and does not terminate).
 *    Environment: BZB200_DEC_BATCH_BYTES (scratch memory per batch of blocks, default 24 GB and at most 60 % of the free
 *    device memory), BZB200_DEC_MAX_OUTPUT (largest output the host->host entry points allocate for; default half of
 *    the free device memory), BZB200_DEC_SPLIT=0 (fused Huffman+MTF kernel instead of the default split pipeline).
 * ------------------------------------------------------------------------ */
/* Device in -> device out on the context's stream; synchronises.  Returns BZB200_OK (*bz_error = 0, *out_n bytes
 * written), BZB200_E_DATA (*bz_error = kind, the first *out_n bytes are what the reference yields before the error), or
 * BZB200_E_ARG when cap_bytes is too small (*out_n = bytes required; nothing useful was written). */
BZB200_API int bzb200_decompress_device(bzb200_ctx* c, const uint8_t* d_in, size_t n, uint8_t* d_out, size_t cap_bytes,
                             size_t* out_n, int* bz_error);
/* HOST in -> HOST out: H2D copy, decode, D2H copy of *out_n bytes; same return convention.  A device staging buffer
 * of cap_bytes is kept in the context, so pass the expected size (BZB200_E_ARG reports the exact one), not a huge bound. */
BZB200_API int bzb200_decompress_host(bzb200_ctx* c, const uint8_t* h_in, size_t n, uint8_t* h_out, size_t cap_bytes,
                           size_t* out_n, int* bz_error);
/* Counters of the last decode, out[0..7]: 0 streams, 1 blocks on the chain, 2 magic candidates found, 3 batches,
 * 4 Huffman symbols decoded, 5 bytes before RLE1 undo (inverse-BWT elements), 6 output bytes, 7 reserved. */
BZB200_API int bzb200_dec_stats(const bzb200_ctx* c, uint64_t* out, size_t cap);

/* Streaming decoder object — what a Rust shim's `BZip2Decoder` binds (input is buffered; the blocks are decoded on
 * the GPU when the input iterator is exhausted). */
typedef struct bzb200_dec bzb200_dec;
BZB200_API int bzb200_dec_create(int device, bzb200_dec** out);
BZB200_API int bzb200_dec_write(bzb200_dec* d, const uint8_t* p, size_t n);
/* Decodes everything written so far.  BZB200_OK, or BZB200_E_DATA: the bytes before the error can still be read, and
 * bzb200_dec_error_kind() gives the BZip2Error the shim returns after them. */
BZB200_API int bzb200_dec_finish(bzb200_dec* d);
BZB200_API int bzb200_dec_error_kind(const bzb200_dec* d);
BZB200_API size_t bzb200_dec_read(bzb200_dec* d, uint8_t* dst, size_t cap);
BZB200_API size_t bzb200_dec_output_size(const bzb200_dec* d);
BZB200_API int bzb200_dec_reset(bzb200_dec* d);
BZB200_API void bzb200_dec_destroy(bzb200_dec* d);
BZB200_API const char* bzb200_dec_last_error(const bzb200_dec* d);
/* One-shot host->host: `bytes.decode(&mut BZip2Decoder::new()).collect()`.  *out is malloc'ed (also on
 * BZB200_E_DATA: the bytes before the error); free with bzb200_free. */
BZB200_API int bzb200_decompress(int device, const uint8_t* in, size_t n, uint8_t** out, size_t* out_n, int* bz_error);

/* ------------------------------------------------------------------------
 * 4. Instrumentation (parity tests, bench roofline).  Not needed by a shim.
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
