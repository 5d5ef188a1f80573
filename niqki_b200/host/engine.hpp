// engine.hpp — the L2 layer of the niqki_b200 host: file-level operations of the reference's
// `class Index` (/root/reference/src/niqki_index.h:154-206) re-written on top of the C ABI of
// libniqki_b200.so.  Where the reference calls compute_sketch / insert_sketch / query_sketch once
// per sequence from OpenMP workers, this host fills pinned batches (files are inflated and parsed
// by a pool of reader threads) and hands whole batches to the GPU.  Same method names, same
// observable behaviour (genome ids in file/record order == the reference at OMP_NUM_THREADS=1).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "gz_writer.hpp"
#include "niqki_b200.h"

namespace nqh {

struct EngineOptions {
  int device = 0;
  bool binary_output = false;   // additive --binary: the reference's unreachable binary writer (B9)
  bool matrix_nowrap = false;   // additive --nowrap: 32-bit matrix counters instead of uint16 (B6)
  unsigned reader_threads = 0;  // 0 = hardware concurrency (capped)
  bool verbose = false;
};

class Engine {
 public:
  // Index::Index(lF,K,W,H,filename,min_fract) — niqki_index.cpp:13-38
  Engine(uint32_t S, uint32_t K, uint32_t W, uint32_t H, const std::string& out_path, double min_fract,
         const EngineOptions& opt);
  // Index::Index(filestr, pretty, filename) — the load constructor, niqki_index.cpp:63-102
  Engine(const std::string& dump_path, const std::string& out_path, const EngineOptions& opt);
  ~Engine();
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;

  void select_best_H(double genome_size);                     // :126-138
  void insert_file_of_file_whole(const std::string& fof);    // :461-500
  void insert_file_lines(const std::string& file);           // :383-408
  void query_file_of_file_whole(const std::string& fof);     // :523-540
  void query_file_lines(const std::string& file);            // :412-430
  void query_matrix();                                       // :614-628
  void dump_index_disk(const std::string& path);             // :42-59
  void close_output();

  uint32_t getNbGenomes() const { return genome_numbers_; }
  const nq_params& params() const { return p_; }
  uint64_t kernel_launches() const;

 private:
  struct Batch;
  void init_ctx();
  void flush_insert(Batch& b);
  void flush_query(Batch& b);
  void ensure_store(uint64_t extra_entries);
  void build_index();  // posting lists of everything inserted so far (no-op when up to date)
  void write_hits(const std::string& name, const uint32_t* counts, const uint32_t* gids, uint64_t n);
  const std::string& frac_text(uint32_t count);

  EngineOptions opt_;
  nq_params p_{};
  nq_ctx* ctx_ = nullptr;
  nq_index* ix_ = nullptr;          // index over gids [0, indexed_)
  uint32_t indexed_ = 0;            // genomes covered by ix_
  uint32_t genome_numbers_ = 0;     // ids handed out so far
  std::vector<std::string> filenames_;
  // sketches of genomes [store_base_, store_base_+store_n_) waiting in HBM for the next build
  int32_t* d_store_ = nullptr;
  uint64_t store_cap_ = 0, store_n_ = 0;
  uint32_t store_base_ = 0;
  int32_t* d_query_ = nullptr;  // sketches of the current query batch
  uint64_t query_cap_ = 0;
  GzWriter out_;
  std::vector<std::string> frac_cache_;
};

}  // namespace nqh
