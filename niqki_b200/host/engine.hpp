// engine.hpp — the L2 layer of the niqki_b200 host: file-level operations of the reference's
// `class Index` (/root/reference/src/niqki_index.h:154-206) re-written on top of the C ABI of
// libniqki_b200.so.  Where the reference calls compute_sketch / insert_sketch / query_sketch once
// per sequence from OpenMP workers, this host fills pinned batches (files are inflated and parsed
// by a pool of reader threads) and hands whole batches to the GPU.  Same method names, same
// observable behaviour (genome ids in file/record order == the reference at OMP_NUM_THREADS=1).
#pragma once
#include <cstdint>
#include <future>
#include <memory>
#include <string>
#include <vector>

#include "gz_writer.hpp"
#include "niqki_b200.h"

namespace nqh {

struct EngineOptions {
  int device = 0;               // first CUDA device
  int gpus = 1;                 // additive --gpus N: the index is sharded by genome id over devices device..device+N-1
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
  // One device of the box.  With --gpus N the index is sharded by genome id in contiguous blocks
  // (nq_shard_range), shard r on device r; sketching work goes round-robin over the devices and
  // the sketches move to their owner when the index is built.
  struct Segment { uint32_t gid0, count; uint64_t row; };  // sketches of gids [gid0, gid0+count) at store row `row`
  struct Shard {
    nq_ctx* ctx = nullptr;
    nq_comm* comm = nullptr;
    nq_index* ix = nullptr;        // posting lists of gids [gid0, gid0+n)
    uint32_t gid0 = 0, n = 0;
    int32_t* d_store = nullptr;    // sketches produced on this device since the last build
    uint64_t store_cap = 0, store_n = 0;
    std::vector<Segment> segs;
    int32_t* d_query = nullptr;    // this device's slice of the current query batch
    uint64_t query_cap = 0;
    int32_t* d_all = nullptr;      // all-gathered query sketches / broadcast matrix rows
    uint64_t all_cap = 0;
  };
  void init_ctx();
  template <typename Fn>
  void on_all_devices(Fn&& fn);   // fn(r) on one host thread per device; rethrows the first failure
  void submit_insert(Batch& b);   // sketch a batch on the next device (asynchronous with --gpus > 1)
  void sketch_into_store(int r, Batch& b);
  void wait_inserts();
  Batch& insert_batch();          // the batch buffer to fill next
  void flush_query(Batch& b);
  void reserve_rows(int r, int32_t*& buf, uint64_t& cap, uint64_t rows);
  void build_index();  // posting lists of everything inserted so far (no-op when up to date)
  void export_all(std::vector<uint32_t>& sizes, std::vector<uint32_t>& gids);  // dump layout over all shards
  void import_all(const std::vector<uint32_t>& sizes, const std::vector<uint32_t>& gids);
  uint64_t pending_sketches() const;
  void write_hits(const std::string& name, const uint32_t* counts, const uint32_t* gids, uint64_t n);
  const std::string& frac_text(uint32_t count);

  EngineOptions opt_;
  nq_params p_{};
  std::vector<Shard> sh_;           // one per device
  bool have_index_ = false;         // shards hold posting lists of gids [0, indexed_)
  uint32_t indexed_ = 0;            // genomes covered by the index
  uint32_t genome_numbers_ = 0;     // ids handed out so far
  std::vector<std::string> filenames_;
  // insert pipeline: one pinned batch buffer per device, filled by the reader while the others are sketched
  std::vector<std::unique_ptr<Batch>> ins_batches_;
  std::vector<std::future<void>> ins_pending_;
  int ins_cur_ = 0;
  GzWriter out_;
  std::vector<std::string> frac_cache_;
};

}  // namespace nqh
