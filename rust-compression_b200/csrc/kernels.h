// kernels.h — internal host-side interface between pipeline.cu and the kernel files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <tuple>
#include <utility>
#include <vector>

namespace bzb {

struct BlockDesc;

// Typed kernel launcher: converts arguments to the kernel's exact parameter types, counts launches and
// (optionally) brackets every launch with CUDA events on the launching stream.
struct Launcher {
  cudaStream_t stream = nullptr;
  uint64_t launches = 0;
  cudaError_t err = cudaSuccess;
  const char* err_kernel = nullptr;

  bool profiling = false;
  struct Rec {
    const char* name;
    uint64_t launches;
    double ms;
  };
  std::vector<Rec> recs;
  struct Pending {
    const char* name;
    cudaEvent_t a, b;
  };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> pool;

  cudaEvent_t get_event();
  void resolve();  // synchronises pending events into recs
  void clear_profile();
  ~Launcher();

  template <class... KArgs, class... Args>
  void launch(const char* name, void (*fn)(KArgs...), dim3 grid, dim3 block, Args&&... args) {
    launch_smem(name, fn, grid, block, 0, std::forward<Args>(args)...);
  }

  template <class... KArgs, class... Args>
  void launch_smem(const char* name, void (*fn)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
    static_assert(sizeof...(KArgs) == sizeof...(Args), "kernel argument count mismatch");
    if (err != cudaSuccess) return;
    if (grid.x == 0 || grid.y == 0 || grid.z == 0) return;
    std::tuple<KArgs...> vals{static_cast<KArgs>(args)...};
    launch_impl(name, (const void*)fn, grid, block, smem, vals, std::index_sequence_for<KArgs...>{});
  }

 private:
  template <class Tuple, size_t... I>
  void launch_impl(const char* name, const void* fn, dim3 grid, dim3 block, size_t smem, Tuple& vals,
                   std::index_sequence<I...>) {
    void* ptrs[] = {(void*)&std::get<I>(vals)..., nullptr};
    raw_launch(name, fn, grid, block, smem, ptrs);
  }
  void raw_launch(const char* name, const void* fn, dim3 grid, dim3 block, size_t smem, void** args);
};

// ---- k1_rle.cu ----
uint64_t k1_num_tiles(uint64_t N);
uint32_t k1_cut_window();
uint32_t k1_tile_bytes();
void launch_k1_heads(Launcher& L, const uint8_t* d_in, uint64_t N, uint64_t t0, uint64_t t1, long long* d_tile_head);
void launch_k1_counts(Launcher& L, const uint8_t* d_in, uint64_t N, uint64_t t0, uint64_t t1,
                      const long long* d_tile_head, long long* d_tile_carry, uint32_t* d_tile_cnt);
void launch_k1_prefix(Launcher& L, uint64_t N, const uint32_t* d_tile_cnt, uint64_t* d_tile_E);
void launch_k1_cut_phase(Launcher& L, const uint8_t* d_in, uint64_t N, uint32_t T, const long long* d_tile_carry,
                         const uint64_t* d_tile_E, uint32_t K, uint64_t* d_F, uint64_t* d_state, uint64_t* d_in_off,
                         uint64_t* d_rle_off, uint32_t max_blocks, uint32_t* d_nblocks, uint32_t* d_maxlen);
// sharded plan: pointers shifted to global indices (v_*), tiles [t_a, t_b) only
void launch_k1_slice_heads(Launcher& L, const uint8_t* v_in, uint64_t N, uint64_t t_a, uint64_t t_b, long long* v_tile_head);
void launch_k1_slice_summary(Launcher& L, const long long* head, const uint32_t* cnt, uint64_t n, uint64_t* d_out2);
void launch_k1_slice_counts(Launcher& L, const uint8_t* v_in, uint64_t N, uint64_t t_a, uint64_t t_b, long long carry_in,
                            const long long* v_tile_head, long long* v_tile_carry, uint32_t* v_tile_cnt);
void launch_k1_slice_scan_carry(Launcher& L, uint64_t t_a, uint64_t t_b, long long carry_in, const long long* v_tile_head,
                                long long* v_tile_carry);
void launch_k1_slice_tile_counts(Launcher& L, const uint8_t* v_in, uint64_t N, uint64_t t_a, uint64_t t_b,
                                 const long long* v_tile_carry, uint32_t* v_tile_cnt);
void launch_k1_slice_prefix(Launcher& L, uint64_t t_a, uint64_t t_b, uint64_t E_in, const uint32_t* v_tile_cnt,
                            uint64_t* v_tile_E);
void launch_k1_slice_windows(Launcher& L, const uint8_t* v_in, uint64_t N, uint32_t T, const long long* v_tile_carry,
                             const uint64_t* v_tile_E, uint64_t t_lo, uint64_t t_hi, uint64_t t_limit, uint64_t Etot,
                             uint64_t x0, uint64_t j0, uint32_t nj, uint64_t* d_F);
void launch_k1_scatter(Launcher& L, const uint8_t* d_in, uint64_t N, uint64_t in_lo, uint64_t in_hi,
                       const long long* d_tile_carry, const uint64_t* d_tile_E, uint8_t* d_txt);
void launch_k5_crc(Launcher& L, const uint8_t* d_in, const uint64_t* d_in_off, uint32_t nblocks, uint32_t* d_crc);
void launch_k1_inuse(Launcher& L, const uint8_t* d_txt, const uint64_t* d_rle_off, uint32_t nblocks, uint32_t* d_inuse);

// ---- k2_bwt.cu ----
struct BwtScratch {
  uint64_t* A;          // [M] sort elements (ping)
  uint64_t* B;          // [M] sort elements (pong)
  uint32_t* rank;       // [M] rank[pos] = slot of the head of pos's group
  uint32_t* sa;         // [M] sa[slot] = pos | flags (rotations in the order established so far)
  uint2* tile_meta;     // [nb][ls_tiles] (entries in the tile's work list, entries before its first group head)
  uint32_t ls_tiles_cap;
  uint32_t* cnt;        // [nb] active elements per block
  uint32_t* hist;       // [nb][tiles][512] radix pass: per-tile status words (decoupled look-back)
  uint32_t* oshist;     // [nb][5][512] digit histograms -> bucket offsets of the passes (row pitch 512)
  uint32_t* pairhist;   // [nb][2^(2 bits)] symbol-pair histogram of the compact-key initial sort
  uint32_t* ticket;     // [nb] tile tickets
  int4* tsum;           // [nb][tiles] regroup tile summaries
  uint32_t* state;      // [nb] 0 active, 1 fix-up pending, 2 done
  uint32_t* shift;      // [nb]
  uint32_t* sparse;     // [nb] 1: few unresolved rotations left, the block takes the radix path
  uint32_t* stats;      // [nb][4]: groups created, BIG members, unresolved, periodic flag
  uint32_t* rounds;     // [nb] rounds until done (instrumentation)
  uint32_t* global;     // [4]: total unresolved, max active per block, error flag, spare
  uint32_t tiles_cap;   // tiles per block the hist/tsum arrays were sized for
};
uint32_t bwt_tile_elems();
uint32_t bwt_ls_tile_elems();
// Runs the whole rotation sort for a batch. txt = batch slice of the RLE1 stream, desc[nb] on device.
// Outputs: last column L[M] (same layout as txt), origptr[nb]. Returns 0 or a negative internal error.
struct BwtStats {
  uint32_t rounds = 0;             // doubling rounds after the initial sort (max over batches)
  uint32_t radix_passes = 0;       // radix pass launches
  uint64_t elems_sorted = 0;       // rotations in the initial sort + unresolved rotations entering each round
  uint64_t radix_elem_passes = 0;  // elements moved by radix passes (list length x 5, summed)
  uint64_t local_elems = 0;        // work-list entries handled by k2_local_sort
};
// d_inuse[nb][8] = in-use byte maps of the blocks, max_alpha = most in-use byte values of any block (0: unknown): the
// initial sort packs its keys with the bit width that alphabet needs (k2_bwt.cu, "Initial sort keys").
size_t bwt_pairhist_bytes(uint32_t nb, uint32_t max_alpha);
int run_bwt(Launcher& L, const uint8_t* d_txt, const BlockDesc* d_desc, const uint32_t* d_inuse, uint32_t max_alpha,
            uint32_t nb, uint32_t nmax, uint64_t M, BwtScratch& S, uint8_t* d_last, uint32_t* d_origptr, BwtStats* stats);

// ---- k3_mtf.cu ----
uint32_t mtf_chunk_elems(uint64_t batch_bytes);  // bytes per MTF chunk for a batch of that many RLE1 bytes
void launch_mtf(Launcher& L, const uint8_t* d_last, const BlockDesc* d_desc, const uint32_t* d_inuse, uint32_t nb,
                uint32_t nmax, uint32_t max_alpha_bytes /* most in-use byte values of any block, 0 = unknown */,
                uint32_t chunk /* mtf_chunk_elems of the batch */, int* d_chunk_state /*[nb][chunks][256]*/, uint4* d_chunk_zle /*[nb][chunks]*/,
                uint2* d_chunk_base /*[nb][chunks]*/, uint32_t chunks_cap, uint16_t* d_sym, uint32_t* d_freq /*[nb][258]*/,
                uint32_t* d_mtf_count /*[nb]*/);

// ---- k4_huff.cu ----
struct HuffBuffers {
  uint8_t* lens;       // [nb][5][6][258]  tables after init / pass 1..4 (libbzip2 table order)
  uint32_t* rfreq;     // [nb][6][258]
  uint8_t* sel;        // [nb][MAX_SELECTORS]
  uint8_t* selmtf;     // [nb][MAX_SELECTORS]
  uint32_t* codes;     // [nb][6][258]   code | len<<24
  uint32_t* gbits;     // [nb][MAX_SELECTORS] exclusive bit offsets of each 50-symbol group inside the block's data
  uint32_t* meta;      // [nb][8]: alpha, ngroups, nsel, hdr_bits, data_bits, lm_count, spare, spare
  uint8_t* lm_scratch; // package-merge scratch, [lm_slots][LM_BYTES]
  uint32_t* lm_list;   // [nb*6] queue of (block*6+table) ids that need the length-limited fallback
  uint32_t* lm_count;  // [1]
  uint32_t lm_slots;   // concurrent fallback workers
};
size_t huff_lm_scratch_bytes();
void launch_huffman(Launcher& L, const uint16_t* d_sym, const BlockDesc* d_desc, const uint32_t* d_mtf_count,
                    const uint32_t* d_freq, const uint32_t* d_inuse, uint32_t nb, uint32_t max_groups_per_block,
                    HuffBuffers& H);

// ---- k6_pack.cu ----
void launch_pack(Launcher& L, const uint16_t* d_sym, const BlockDesc* d_desc, const uint32_t* d_mtf_count,
                 const uint32_t* d_inuse, const uint32_t* d_crc, const uint32_t* d_origptr, uint32_t nb,
                 uint32_t max_groups_per_block, HuffBuffers& H, uint64_t* d_blockbit /*[nb+1]*/, uint64_t* d_bitcursor,
                 uint8_t* d_out);
void launch_bit_append(Launcher& L, uint8_t* d_dst, uint64_t dst_bit, const uint8_t* d_src, uint64_t nbits);
void launch_write_header(Launcher& L, int level, uint8_t* d_out);
void launch_write_trailer(Launcher& L, uint8_t* d_out, uint64_t at_bit, uint32_t combined_crc);
void launch_combine_crc(Launcher& L, const uint32_t* d_crc, uint32_t n, uint32_t* d_combined);
void launch_write_trailer_dev(Launcher& L, uint8_t* d_out, const uint64_t* d_bitcursor, const uint32_t* d_combined);

}  // namespace bzb
