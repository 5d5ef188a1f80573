// internal.h — host-side declarations shared by the .cu files of libniqki_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "niqki_b200.h"

struct nq_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;  // compute stream (owned or borrowed)
  bool own_stream = false;
  cudaStream_t copy_stream = nullptr;  // H2D pipeline stream (owned)
  cudaStream_t d2h_stream = nullptr;   // D2H pipeline stream (owned)
  // persistent staging of nq_sketch_batch (host-buffer form): two slots, grown on demand
  struct Slot {
    char* d_bases = nullptr; size_t cap_bases = 0;
    int32_t* d_sk = nullptr; size_t cap_cells = 0;
    uint32_t* d_flags = nullptr; size_t cap_entries = 0;
    uint32_t* h_flags = nullptr;               // pinned landing buffer of the flags (the caller's array may be pageable:
    uint64_t flags_e0 = 0, flags_n = 0;        // a direct copy would block the host until the batch is done)
    cudaEvent_t h2d = nullptr, done = nullptr, d2h = nullptr;
    // 2-bit packed form of the batch (pack.cpp): pinned host staging + device copies
    uint32_t* h_codes = nullptr; uint32_t* d_codes = nullptr; size_t cap_words = 0;
    uint32_t* h_blk = nullptr; uint32_t* d_blk = nullptr;
    uint16_t* h_pool = nullptr; uint16_t* d_pool = nullptr;
  } slot[3];
  // pinned ring for small host -> device uploads (span tables, offsets) that must not stall the host:
  // a copy from pageable memory would wait for the stream
  char* h_ring = nullptr; size_t ring_cap = 0, ring_at = 0;
  cudaEvent_t ring_wrap = nullptr;  // recorded when the write position wraps; awaited before the ring is reused
  unsigned host_threads = 0;  // packer threads (0 = hardware concurrency, capped)
  int pack_mode = -1;         // -1 auto (long entries travel packed), 0 never, 1 always
  int sm_count = 148;
  size_t smem_optin = 0;  // max dynamic shared memory per block (opt-in)
  uint64_t launches = 0;  // kernels launched through this context
  // optional per-kernel device timing (nq_ctx_set_timing): event pairs resolved at report time
  bool timing = false;
  struct Timed { int kind; cudaEvent_t a, b; };
  std::vector<Timed> timed;
  double kind_ms[8] = {0};
  uint64_t kind_n[8] = {0};
  uint64_t last_query_gathered = 0;  // gids gathered by the last query call (roofline numerator)
  uint64_t h2d_bytes = 0;            // bytes the host-buffer sketch calls have copied to the device so far
};

enum NqKernelKind { NQK_SCAN = 0, NQK_DENSIFY = 1, NQK_TRANSPOSE = 2, NQK_CELLSORT = 3, NQK_QUERY = 4, NQK_MATRIX = 5, NQK_SLAB = 6 };

// RAII bracket: records an event pair around a kernel launch when timing is on
struct NqTimer {
  nq_ctx* ctx; int kind; cudaEvent_t a = nullptr, b = nullptr;
  NqTimer(nq_ctx* c, int k) : ctx(c), kind(k) {
    if (ctx->timing) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, ctx->stream); }
  }
  ~NqTimer() {
    if (a) { cudaEventRecord(b, ctx->stream); ctx->timed.push_back({kind, a, b}); }
  }
};

int nq_set_error(int code, const char* fmt, ...);

#define NQ_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return nq_set_error(NQ_ERR_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,    \
                          cudaGetErrorString(_e));                                             \
  } while (0)

#define NQ_TRY(expr)            \
  do {                          \
    int _s = (expr);            \
    if (_s != NQ_OK) return _s; \
  } while (0)

#define NQ_CHECK_LAUNCH(ctx)       \
  do {                             \
    (ctx)->launches++;             \
    NQ_CUDA(cudaPeekAtLastError()); \
  } while (0)

// stream-ordered scratch allocation
inline int nq_dmalloc(nq_ctx* ctx, void** p, size_t bytes) {
  NQ_CUDA(cudaMallocAsync(p, bytes ? bytes : 16, ctx->stream));
  return NQ_OK;
}
inline void nq_dfree(nq_ctx* ctx, void* p) {
  if (p) cudaFreeAsync(p, ctx->stream);
}

// NVTX range around a C-ABI entry point (tracing): shows up in Nsight Systems / ncu --nvtx
// timelines under the domain "niqki_b200"; a no-op when no tool is attached.
#include <nvtx3/nvToolsExt.h>
struct NqRange {
  explicit NqRange(const char* name) { nvtxRangePushA(name); }
  ~NqRange() { nvtxRangePop(); }
  NqRange(const NqRange&) = delete;
  NqRange& operator=(const NqRange&) = delete;
};
#define NQ_RANGE() NqRange _nq_range(__func__)

// stream-ordered scratch that is returned on every path out of a function
struct NqScratch {
  nq_ctx* ctx;
  void* p = nullptr;
  explicit NqScratch(nq_ctx* c) : ctx(c) {}
  NqScratch(const NqScratch&) = delete;
  NqScratch& operator=(const NqScratch&) = delete;
  ~NqScratch() { nq_dfree(ctx, p); }
  template <typename T>
  T* as() const { return static_cast<T*>(p); }
};

int nq_params_check(const nq_params* p);
// stream-ordered upload of a small host array through the context's pinned ring (capi.cu)
int nq_upload_small(nq_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);

// Measurement knobs (NQ_* environment variables) exist only in builds with -DNQ_TUNING; the
// product library never lets the environment choose a kernel.
inline const char* nq_tuning_env(const char* name) {
#ifdef NQ_TUNING
  return getenv(name);
#else
  (void)name;
  return nullptr;
#endif
}

// ---- sketch.cu
int nq_launch_sketch(nq_ctx* ctx, const nq_params* p, const char* d_bases, uint64_t bases_capacity,
                     const uint64_t* h_offsets, uint64_t n, const uint32_t* h_rec_entry, uint64_t n_entries,
                     int32_t* d_sketches, uint32_t* d_flags);
int nq_launch_densify(nq_ctx* ctx, const nq_params* p, int32_t* d_sketches, uint64_t n, uint32_t* d_flags);
// sketch_packed.cu: the same on the 2-bit packed wire format of pack.cpp
int nq_launch_sketch_packed(nq_ctx* ctx, const nq_params* p, const uint32_t* d_codes, const uint32_t* d_blk, const uint16_t* d_pool,
                            const uint64_t* h_offsets, uint64_t n, const uint32_t* h_rec_entry, uint64_t n_entries,
                            int32_t* d_sketches, uint32_t* d_flags);

// ---- index.cu / query.cu / matrix.cu
struct nq_index {
  nq_ctx* ctx = nullptr;
  nq_params p{};
  uint32_t n = 0;         // genomes in this shard
  uint32_t gid_base = 0;  // gid of local genome 0
  uint64_t n_postings = 0;
  // Cell-major CSR with fixed strides, element type u16 when n <= 65535 ("compact", halves the
  // bytes every probe drags through L2/HBM), u32 otherwise.  Postings hold LOCAL ids (gid-gid_base).
  //   {b,e} = dir[cell*row_stride + fp];  list (cell, fp) = gids[cell*gid_stride + b .. + e)
  uint32_t elem = 4;        // bytes per offset / gid element: 2 or 4
  uint32_t row_stride = 0;  // directory entries per cell (= range)
  uint32_t gid_stride = 0;  // elements per cell in d_gids: n rounded up to 32 bytes
  void* d_row = nullptr;    // directory [F][row_stride] of packed {begin,end}: u32 (elem 2) or uint2 (elem 4)
  void* d_gids = nullptr;   // [F][gid_stride] + kQuerySlack elements
  // split16 side arrays (65400 < n <= 131072 only): the same postings as u16 (gid & 0xFFFF; lists are
  // gid-sorted, so each list is a run of ids < 65536 followed by a run of ids >= 65536) and a directory
  // {begin, mid, end, 0} with mid = first posting >= 65536.  Halves the bytes the query kernel gathers.
  uint4* d_dir3 = nullptr;       // [F][row_stride]
  uint16_t* d_gids16 = nullptr;  // [F][gid_stride] + kQuerySlack
  // slab side structure (compact shards, 2^W >= 32; slab.cu): the query kernel's own copy of the
  // postings.  A list occupies 1..3 granules of G ids (16 B * G/8, granule-aligned, padded with ids
  // >= n that land in spare counter words); meta[cell][fp/32] = {b0, b1, b2, base} gives the list of
  // fp its granule count (bit fp%32 of b0 + 2*b1) and its first granule (base + weighted popcount of
  // the bits below); b2 marks lists longer than 3 granules, whose tail stays in d_gids behind a
  // descriptor granule.  A probe is one 16-byte meta read + one or two sectors of ids.
  uint4* d_meta = nullptr;        // [F][range/32]
  uint2* d_slab = nullptr;        // granules, 8-byte units; granule 0 is the all-padding dummy
  uint32_t* d_cell_gran = nullptr;  // [F+1] first granule of each cell
  uint32_t slab_G = 0;            // ids per granule (8, 16, 32, 64); 0 = no slab
  uint32_t slab_nrf = 0;          // G = 8: fixed gather rounds per group of 32 probes (0 = batches)
  uint64_t slab_granules = 0;
  // device-resident results of the last nq_query_batch_device(out == NULL)
  uint64_t* d_pool = nullptr;
  uint64_t pool_cap = 0;
};

struct nq_hits {
  std::vector<uint64_t> ptr;
  std::vector<uint32_t> counts, gids;
};

int nq_index_build_impl(nq_ctx* ctx, const nq_params* p, const int32_t* d_sketches, uint64_t n,
                        uint32_t gid_base, nq_index** out);
constexpr uint32_t kMaxCompact = 65400;  // genomes per shard in the u16 form (leaves room for the query kernel's dummy ids)
constexpr uint32_t kQuerySlack = 64;     // elements reserved behind d_gids[F][gid_stride]
int nq_query_prepare(nq_index* ix);      // fills that slack; call once the index arrays exist
int nq_slab_build(nq_index* ix);            // slab.cu: builds the slab side structure when the shard calls for it
void nq_slab_free(nq_index* ix);
int nq_index_make_split16(nq_index* ix);  // index.cu: builds d_dir3 / d_gids16 when the shard size calls for them
int nq_query_impl(nq_index* ix, const int32_t* d_sketches, uint64_t nq, uint32_t min_score, nq_hits** out);
int nq_query_dense_impl(nq_index* ix, const int32_t* d_sketches, uint64_t nq, uint32_t wrap_mask, uint32_t* d_out);
int nq_matrix_impl(nq_index* ix, uint32_t row_begin, uint32_t row_end, int wrap16, uint32_t* h_counts);
int nq_index_sketches_impl(nq_index* ix, uint32_t row_begin, uint32_t row_end, int32_t* d_sketches);
int nq_matrix_tile_impl(nq_index* ix, const int32_t* d_rows, uint32_t nrows, int wrap16, uint32_t* h_counts);
