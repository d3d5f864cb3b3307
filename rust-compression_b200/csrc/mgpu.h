// mgpu.h — internal interface of the multi-GPU engine (mgpu.cu) towards the streaming encoder object (enc_stream.cu).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string>

struct bzb200_pool;

// What one span of input turns into (a whole stream, or one window of a streaming encoder).
struct SpanJob {
  int level = 9;
  const uint8_t* h_in = nullptr;
  uint64_t n = 0;
  bool first = true;        // the stream header goes in front
  bool final = true;        // false: the last block stays open (its input is handed back through `consumed`)
  uint8_t carry = 0;        // partial byte carried in from the previous span ...
  uint32_t carry_bits = 0;  // ... and its valid bits (0..7)
  uint32_t combined = 0;    // combined CRC so far (in), after this span (out)
  uint8_t* h_out = nullptr;
  size_t cap = 0;
  // single-GPU pools only: the span is already in the memory of the pool's device (d_in, instead of h_in) and / or the
  // output stays there (d_out, instead of h_out; zero-filled by the engine); last_byte = the byte that holds bit
  // end_bits - 1 (the partial byte a streaming encoder carries into its next span)
  const uint8_t* d_in = nullptr;
  uint8_t* d_out = nullptr;
  uint8_t last_byte = 0;
  // results
  uint64_t end_bits = 0;  // bits in h_out (incl. the carried ones); when !final the last byte may be partial
  uint64_t consumed = 0;  // input bytes whose blocks were encoded
  uint64_t blocks = 0;
};

// Runs one span on the pool (blocking).  n must be > 0.
int pool_run_span(bzb200_pool* p, SpanJob* J);
const std::string& pool_error(const bzb200_pool* p);
int pool_device(const bzb200_pool* p, int w);  // CUDA ordinal of worker w
int pool_contexts_per_gpu();                   // workers a pool creates per GPU (env BZB200_MG_CTX_PER_GPU, default 1)
