// decoder.h — internal interface of the block-parallel bzip2 decoder (decoder.cu) towards its host
// (dec_abi.cu for the GPU build, tests/cpp/dec_emu.cpp for the host emulation of the same kernel bodies).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string>

namespace bzb {

#ifdef BZB_EMU
struct Launcher {  // stand-in: the emulation runs the kernel bodies in loops
  uint64_t launches = 0;
};
#else
struct Launcher;
#endif

// Device memory and the few services the decoder needs from its host side.
struct DecMem {
  virtual ~DecMem() {}
  // a device buffer of at least `bytes` for role `slot`, contents undefined, kept until the next request for the slot
  virtual void* buf(int slot, size_t bytes) = 0;
  virtual int fill(void* p, int byte, size_t bytes) = 0;                     // stream ordered
  virtual int to_host(void* dst, const void* src, size_t bytes) = 0;         // synchronises
  virtual int to_dev(void* dst, const void* src, size_t bytes) = 0;          // stream ordered, src may be reused on return
  // CRC-32/BZIP2 of data[off[i] .. off[i+1]) for i < nb (k5_crc_blocks, the encoder's kernel)
  virtual int crc_blocks(const uint8_t* d_data, const uint64_t* d_off, uint32_t nb, uint32_t* d_crc) = 0;
  virtual int check() = 0;  // 0, or nonzero when a launch failed (message in err())
  virtual std::string err() = 0;
};

enum DecSlot {
  DS_CAND = 0, DS_COUNT, DS_INFO, DS_OCC, DS_V, DS_W, DS_SEL, DS_CFTAB, DS_SEGLEN, DS_SEGNEXT, DS_SEGOFF,
  DS_SEGRES, DS_T, DS_RLEMAP, DS_CHENTRY, DS_CHOFF, DS_OUTOFF, DS_CRCOFF, DS_CRC, DS_SYM, DS_P, DS_MTF0, DS_PERM, DS_CNTP,
  DS_CMETA, DS_INITL, DS_BASE, DS_COFF, DS_CD0, DS_RAND, DS_NSLOTS
};

struct DecResult {
  uint64_t out_n = 0;      // bytes the reference decoder yields before it stops (all of them when bz_error == 0)
  uint32_t bz_error = 0;   // 0, or BZip2Error ordinal + 1: 1 DataError, 2 DataErrorMagicFirst, 3 DataErrorMagic,
                           // 4 UnexpectedEof, 5 Unexpected (bzip2/error.rs:4-11)
  uint64_t needed = 0;     // output bytes required (== out_n unless the caller's buffer was too small)
  uint32_t too_small = 0;  // 1: nothing beyond `cap` was written and `needed` says how much room a retry needs
  uint32_t streams = 0, blocks = 0, candidates = 0;
  uint64_t syms = 0;       // Huffman symbols decoded on the chain
  uint64_t pre_rle = 0;    // bytes of the chain's blocks before RLE1 undo (= inverse-BWT elements)
  uint32_t batches = 0;
};

// Decodes the (possibly multi-stream) .bz2 buffer d_in[0..n) into d_out[0..cap).  Returns 0 when the pipeline ran
// (data errors of the stream are reported through res->bz_error exactly like the reference reports them), or a
// negative BZB200_E_* code for CUDA / argument failures.  batch_bytes bounds the scratch memory per batch of blocks.
// flags: DEC_SPLIT_D2 selects the split D2 (d2_huff + chunk-parallel d2_mtf_a/b/c) instead of the fused d2_decode.
enum : uint32_t { DEC_SPLIT_D2 = 1u };
int dec_run(Launcher& L, DecMem& M, const uint8_t* d_in, uint64_t n, uint8_t* d_out, uint64_t cap, uint64_t batch_bytes,
            uint32_t flags, DecResult* res);

}  // namespace bzb
