/*
 * niqki_b200.h — C ABI of the B200-native NIQKI hot path (libniqki_b200.so).
 *
 * The reference has no FFI seam: its hot path is the L1 methods of `class Index`
 * (/root/reference/src/niqki_index.h:79-108, 142, 206, 211), called once per sequence from
 * OpenMP file loops.  This header re-expresses those methods as batch-oriented entry points
 * (SURVEY.md §8b).  Every function cites the reference member it replaces (file:line relative to
 * /root/reference/).  Conventions:
 *   - plain C, opaque handles, `int` status (0 = NQ_OK); message via nq_last_error();
 *   - no exceptions cross the ABI; no torch / CUDA types in signatures (a CUDA stream travels as
 *     void*);
 *   - "host" pointers may be pageable or pinned (pinned = faster copies; see nq_host_alloc);
 *     "device" pointers are plain CUDA device pointers on the context's device;
 *   - one host thread per context; calls on one handle are not concurrent;
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with
 *     NQ_ERR_CUDA.
 */
#ifndef NIQKI_B200_H
#define NIQKI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NQ_OK 0
#define NQ_ERR_INVALID 1     /* bad argument / parameter outside the reference's limits (B14) */
#define NQ_ERR_CUDA 2        /* CUDA runtime error (including "no device") */
#define NQ_ERR_UNSUPPORTED 3 /* valid for the reference, not built here yet (e.g. W > 15 index) */
#define NQ_ERR_OVERFLOW 4    /* caller-provided capacity too small */

/* per-entry flags written by the sketch calls */
#define NQ_ENTRY_SKIPPED 1u         /* len <= K: callers of the reference drop it (:395,:450) */
#define NQ_ENTRY_DENSIFY_STALLED 2u /* the reference's densification loop would never end */

/* Scalar fields of class Index (src/niqki_index.h:38-50).  M, mask_M and maxrem are three
 * independent inputs because `-G` leaves the last two stale. */
typedef struct nq_params {
  uint32_t K;         /* k-mer length, <= 31                       */
  uint32_t S;         /* lF = log2(#cells)                         */
  uint32_t W;         /* fingerprint width, W+S < 32               */
  uint32_t H;         /* HyperLogLog bits                          */
  uint32_t M;         /* shift of the HLL part                     */
  uint32_t F;         /* 1 << S                                    */
  uint32_t mask_M;    /* mask of the MinHash part                  */
  uint32_t maxrem;    /* saturation of the HLL part                */
  int32_t range;      /* fingerprint_range = 1 << W                */
  uint32_t min_score; /* (uint32)(min_fract * F), truncating       */
} nq_params;

typedef struct nq_ctx nq_ctx;     /* one CUDA device + stream + scratch */
typedef struct nq_index nq_index; /* one index shard resident in HBM    */
typedef struct nq_hits nq_hits;   /* host-side result of a query batch  */

const char* nq_last_error(void);
const char* nq_version(void);

/* ---- parameters ------------------------------------------------------------------------- */
/* Index::Index(lF,K,W,H,filename,min_fract) — src/niqki_index.cpp:13-29 */
int nq_params_init(nq_params* p, uint32_t K, uint32_t S, uint32_t W, uint32_t H, double min_fract);
/* Index::select_best_H / score_H — src/niqki_index.cpp:126-164 (host scalar code, runs once) */
int nq_params_select_best_H(nq_params* p, double genome_size);

/* ---- context ---------------------------------------------------------------------------- */
int nq_device_count(int* count);
/* `cuda_stream` = an existing cudaStream_t to launch on (e.g. torch's current stream) or NULL to
 * let the context own one. */
int nq_ctx_create(int device, void* cuda_stream, nq_ctx** out);
int nq_ctx_destroy(nq_ctx* ctx);
int nq_ctx_sync(nq_ctx* ctx);
/* number of kernel launches issued through this context so far */
uint64_t nq_ctx_launch_count(const nq_ctx* ctx);
/* Profiling support: when on, every kernel family launched through the context is bracketed by
 * CUDA events on the launching stream.  nq_ctx_timing() synchronises and returns the accumulated
 * device milliseconds and launch count of one family since the last reset, kind = 0 scan
 * (hash + bucket-min), 1 densify, 2 transpose, 3 cell sort (CSR build), 4 query count, 5 matrix,
 * 6 granule copy of the postings for the query kernel (slab). */
int nq_ctx_set_timing(nq_ctx* ctx, int on);
int nq_ctx_timing(nq_ctx* ctx, int kind, double* ms, uint64_t* launches);
int nq_ctx_timing_reset(nq_ctx* ctx);
/* posting-list entries gathered by the most recent query call on this context (its algorithmic
 * HBM traffic is 4 B per entry + 8 B per probed cell + 4 B per sketch cell + 8 B per hit) */
uint64_t nq_ctx_last_query_gathered(const nq_ctx* ctx);
/* bytes of sequence (packed or not) that the host-buffer sketch calls have copied to the device so far */
uint64_t nq_ctx_h2d_bytes(const nq_ctx* ctx);
/* pinned host memory for fast host<->device copies */
void* nq_host_alloc(size_t bytes);
void nq_host_free(void* p);
/* plain device memory / synchronous copies for hosts that do not link the CUDA runtime
 * (kind: 0 host->device, 1 device->host, 2 device->device) */
int nq_device_alloc(nq_ctx* ctx, size_t bytes, void** out);
int nq_device_free(nq_ctx* ctx, void* p);
int nq_device_copy(nq_ctx* ctx, void* dst, const void* src, size_t bytes, int kind);
/* device -> device across two contexts (NVLink peer copy, or through the host where peers are not
 * enabled); synchronous.  nq_device_fill: stream-ordered memset on the context's stream. */
int nq_device_copy_peer(nq_ctx* dst_ctx, void* dst, nq_ctx* src_ctx, const void* src, size_t bytes);
int nq_device_fill(nq_ctx* ctx, void* p, int byte, size_t bytes);

/* ---- sketching: Index::compute_sketch + sketch_densification ---------------------------- */
/* src/niqki_index.cpp:335-358, 313-331 (and 114-123, 211-236, 240-273, 277-310 underneath).
 * `bases` = the entries' characters concatenated (raw ASCII as read from FASTA/FASTQ, any byte
 * allowed), `offsets[n+1]` their boundaries.  One sketch per entry, fully densified, int32[n][F]
 * with -1 for cells that stayed empty (only possible with NQ_ENTRY_* flags set).
 * `flags` (nullable) receives NQ_ENTRY_* per entry. */
int nq_sketch_batch(nq_ctx* ctx, const nq_params* p, const char* bases, const uint64_t* offsets,
                    uint64_t n, int32_t* sketches, uint32_t* flags);
/* Record form.  The reference's whole-file loops (insert_file_whole :442-457, query_file_whole
 * :505-519) call compute_sketch once per FASTA/FASTQ record on ONE sketch; with more than one
 * record > K that loop never terminates, so the defined behaviour here is: every
 * record is scanned with its own seed (:340-341), records of one entry are min-merged, and the
 * entry is densified once.  `rec_entry[n_records]` (non-decreasing; NULL = one entry per record)
 * maps records to sketch rows 0..n_entries-1; entries without any record > K come back all -1
 * with NQ_ENTRY_SKIPPED.  `sketches_on_device` != 0: `sketches` is a device array (flags stay a
 * host array either way) — the CLI host keeps its sketch store in HBM. */
int nq_sketch_records(nq_ctx* ctx, const nq_params* p, const char* bases, const uint64_t* rec_offsets,
                      uint64_t n_records, const uint32_t* rec_entry, uint64_t n_entries, int32_t* sketches,
                      uint32_t* flags, int sketches_on_device);
/* K1, sequence packing to 2 bits per base (src/niqki_index.cpp:114-123, 211-221, 255-273 decide the
 * codes): the host-buffer sketch calls above pack long entries on the host (AVX2, `threads` helper
 * threads; 0 = all cores, at most 32) so that a base crosses PCIe as 2 bits, plus one bit per base
 * for the 512-base blocks that hold anything but upper-case ACGT; the device cuts every k-mer out
 * of the packed stream.  mode: -1 automatic (entries of >= 4096 characters on average), 0 never
 * (characters travel as they are), 1 always.  Results are identical either way. */
int nq_ctx_set_host_packing(nq_ctx* ctx, int mode, unsigned threads);
/* The packer and the packed kernel as entry points of their own (a host that keeps its genomes
 * packed).  Wire format of a batch of records (rec_offsets[n_records+1] into `bases`):
 *   codes[words]  u32 per 16 bases, base i at bits 2*(i%16), forward code {A:0,C:1,G:2,T:3}, 0 otherwise;
 *                 the stream starts with 512 unused bases and ends with 4 zero words;
 *   blk[blocks]   slot of 512-base block b in `pool`, or 0xFFFFFFFF when all its bytes are upper-case ACGT;
 *   pool[slots*32] u16 per 16 bases: bit i set = byte i is not upper-case ACGT (both strands' code 0, B3);
 *   the first K-1 characters of every record are written with the digits of str2numstrand (:255-273).
 * nq_pack_sizes gives words/blocks for nbytes characters; `pool` needs at most `blocks` slots. */
int nq_pack_sizes(uint64_t nbytes, uint64_t* words, uint64_t* blocks);
int nq_pack_sequences(const char* bases, const uint64_t* rec_offsets, uint64_t n_records, uint32_t K, uint32_t* codes,
                      uint32_t* blk, uint16_t* pool, uint64_t pool_slots, uint64_t* pool_used, unsigned threads);
/* Index::compute_sketch over a packed batch resident in HBM; offsets[n+1] (host) are the records'
 * boundaries in the original characters, offsets[0] == 0 */
int nq_sketch_batch_packed_device(nq_ctx* ctx, const nq_params* p, const uint32_t* d_codes, const uint32_t* d_blk,
                                  const uint16_t* d_pool, const uint64_t* offsets, uint64_t n, int32_t* d_sketches,
                                  uint32_t* d_flags);
/* Same with the characters already in HBM.  `d_bases` must be 16-byte aligned and its allocation
 * must extend to `bases_capacity` >= offsets[n] rounded up to 16.  `offsets` stays a host array.
 * `d_sketches` / `d_flags` are device buffers (d_flags nullable). */
int nq_sketch_batch_device(nq_ctx* ctx, const nq_params* p, const char* d_bases,
                           uint64_t bases_capacity, const uint64_t* offsets, uint64_t n,
                           int32_t* d_sketches, uint32_t* d_flags);
/* Index::sketch_densification alone (src/niqki_index.cpp:313-331) on n device sketches in place */
int nq_densify_device(nq_ctx* ctx, const nq_params* p, int32_t* d_sketches, uint64_t n,
                      uint32_t* d_flags);

/* ---- inverted index: Index::insert_sketch over a batch ---------------------------------- */
/* src/niqki_index.cpp:362-370; storage src/niqki_index.h:55.  Builds, in HBM, the posting lists
 * (cell, fp) -> [gid] of n sketches with gids gid_base .. gid_base+n-1, lists gid-ascending
 * (== the reference at OMP_NUM_THREADS=1).  Only cells with 0 <= fp < range are posted. */
int nq_index_build(nq_ctx* ctx, const nq_params* p, const int32_t* sketches, uint64_t n,
                   uint32_t gid_base, nq_index** out);
int nq_index_build_device(nq_ctx* ctx, const nq_params* p, const int32_t* d_sketches, uint64_t n,
                          uint32_t gid_base, nq_index** out);
int nq_index_free(nq_index* ix);
int nq_index_info(const nq_index* ix, uint64_t* n_postings, uint32_t* n_genomes, uint32_t* gid_base,
                  uint64_t* device_bytes);
/* Glue to the dump layout (Index::dump_index_disk, src/niqki_index.cpp:51-55): per-list sizes for
 * all range*F lists in list-id order, and the lists' gids concatenated in the same order. */
int nq_index_export(nq_index* ix, uint32_t* list_sizes, uint32_t* gids, uint64_t gids_capacity);
/* load-ctor (src/niqki_index.cpp:84-91): gids inside [gid_base, gid_base+n_genomes) are kept */
int nq_index_import(nq_ctx* ctx, const nq_params* p, const uint32_t* list_sizes, const uint32_t* gids,
                    uint32_t n_genomes, uint32_t gid_base, nq_index** out);

/* ---- query: Index::query_sketch over a batch -------------------------------------------- */
/* src/niqki_index.cpp:633-687.  For each query sketch: per-genome hit counts over the probed
 * lists, keep count >= min_score, sort by (count, gid) descending.  Counter width follows the
 * reference (u8 for S<=7, u16 for S<=15, u32 above). */
int nq_query_batch(nq_index* ix, const int32_t* sketches, uint64_t nq, uint32_t min_score,
                   nq_hits** out);
/* sketches already in HBM; `out` may be NULL to leave the (unsorted) hits on the device — used to
 * time the kernels alone */
int nq_query_batch_device(nq_index* ix, const int32_t* d_sketches, uint64_t nq, uint32_t min_score,
                          nq_hits** out);
uint64_t nq_hits_total(const nq_hits* h);
const uint64_t* nq_hits_ptr(const nq_hits* h);    /* nq+1 offsets into counts/gids */
const uint32_t* nq_hits_counts(const nq_hits* h);
const uint32_t* nq_hits_gids(const nq_hits* h);
void nq_hits_free(nq_hits* h);

/* ---- all-vs-all: Index::query_range ----------------------------------------------------- */
/* src/niqki_index.cpp:570-598.  counts[(q-row_begin)*n + j] = number of lists holding both q and
 * j, for q in [row_begin,row_end) (gids relative to gid_base), j over all n genomes.  wrap16 != 0
 * reproduces the reference's uint16_t counters (value mod 65536); thresholding and
 * formatting stay on the host (:600-608, :747-763). */
int nq_matrix_rows(nq_index* ix, uint32_t row_begin, uint32_t row_end, int wrap16, uint32_t* counts);

/* One tile of the genome x genome grid (--matrix over a gid-sharded index).  The
 * reference derives the matrix from the posting lists alone (src/niqki_index.cpp:570-598); here the
 * sketches of a row block are rebuilt from the shard that owns them (nq_index_sketches_device, rows
 * local to the shard), handed to every shard (nq_bcast_sketches), and each shard counts them against
 * its own columns: counts[r*n_shard + j], same wrap16 rule as nq_matrix_rows. */
int nq_index_sketches_device(nq_index* ix, uint32_t row_begin, uint32_t row_end, int32_t* d_sketches);
int nq_matrix_tile(nq_index* ix, const int32_t* d_row_sketches, uint32_t nrows, int wrap16, uint32_t* counts);

/* ---- multi-GPU: the sharded index ------------------------------------------ */
/* The reference is one address space (its callers insert_file_of_file_whole :461-500,
 * query_file_of_file_whole :523-540 and query_matrix :614-628 become multi-GPU here).  The index
 * shards by genome id in contiguous blocks (nq_shard_range), every shard is an ordinary nq_index
 * with its own gid_base on its own device; the one exchange step is an all-gather of the query
 * sketches over NCCL (u16 on the wire when W <= 15: cells outside [0, 2^W) are never probed, :655,
 * and come back as -1), after which every shard counts every query and the host merges the
 * per-shard hit lists — disjoint in gid, so concatenate + sort == the reference's order (:685).
 * A communicator belongs to one context; collective calls are made by every rank, each from the
 * thread that drives its context, and run on the context's stream.  NCCL is loaded at run time. */
typedef struct nq_comm nq_comm;
int nq_comm_unique_id(void* id128);  /* 128 bytes, made on rank 0 and handed to the other ranks */
int nq_comm_init_rank(nq_ctx* ctx, const void* id128, int nranks, int rank, nq_comm** out);
/* one process driving n devices: out[n] communicators, rank i on ctxs[i] */
int nq_comm_init_all(nq_ctx* const* ctxs, int n, nq_comm** out);
int nq_comm_destroy(nq_comm* c);
int nq_comm_info(const nq_comm* c, int* rank, int* nranks, int* nccl_version);
int nq_shard_range(uint64_t n, int nranks, int rank, uint64_t* begin, uint64_t* end);
/* every rank contributes n_local sketches int32[n_local][F] (device); all receive
 * int32[nranks*n_local][F] in rank order.  n_local must be the same on every rank (a rank with
 * fewer queries pads its block with rows of -1, which probe nothing) */
int nq_allgather_sketches(nq_comm* c, const nq_params* p, const int32_t* d_local, uint64_t n_local, int32_t* d_all);
int nq_bcast_sketches(nq_comm* c, const nq_params* p, int32_t* d_sketches, uint64_t n, int root);
/* per-shard results of the same query batch -> one result, each query sorted (count, gid) descending */
int nq_hits_merge(const nq_hits* const* parts, int nparts, nq_hits** out);
int nq_hits_from_arrays(const uint64_t* ptr, const uint32_t* counts, const uint32_t* gids, uint64_t nq, nq_hits** out);

/* ---- synthetic inputs, generated directly in HBM for the benchmarks ------ */
int nq_synth_genomes_device(nq_ctx* ctx, uint64_t seed, uint64_t first_genome, uint64_t n,
                            uint64_t len, char* d_out);
/* entry i = mutated copy q[i] of genome g[i] with substitution threshold thr[i] = floor(d*2^64) */
int nq_synth_mutants_device(nq_ctx* ctx, uint64_t seed, const uint64_t* g, const uint64_t* q,
                            const uint64_t* thr, uint64_t n, uint64_t len, char* d_out);
int nq_synth_reads_device(nq_ctx* ctx, uint64_t seed, uint64_t first_read, uint64_t n,
                          uint64_t genome_len, uint32_t read_len, char* d_out);

#ifdef __cplusplus
}
#endif
#endif
