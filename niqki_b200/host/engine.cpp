// engine.cpp — see engine.hpp.  All sequence work goes through libniqki_b200.so; there is no CPU
// implementation of sketching, indexing or counting in this host.
#include "engine.hpp"

#include <sys/stat.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <deque>
#include <fstream>
#include <future>
#include <iostream>
#include <stdexcept>
#include <thread>

#include "bio_reader.hpp"

namespace nqh {

namespace {

constexpr uint32_t kBufferSize = 10000;           // rows per query_range call (niqki_index.cpp:9)
constexpr uint64_t kBatchBases = 256ull << 20;    // flush a batch at 256 MiB of sequence ...
constexpr uint64_t kBatchCells = 64ull << 20;     // ... or 64 Mi sketch cells (256 MiB of int32)

void check(int status, const char* what) {
  if (status != NQ_OK) throw std::runtime_error(std::string(what) + ": " + nq_last_error());
}

bool exists_test(const std::string& name) {  // niqki_index.h:161-164
  struct stat st;
  return stat(name.c_str(), &st) == 0;
}

// one sequence file, parsed: characters of the records with len > K back to back
struct FileData {
  std::vector<char> bases;
  std::vector<uint64_t> rec_len;
  std::string error;
};

FileData read_whole_file(const std::string& path, uint32_t K) {
  FileData fd;
  try {
    LineReader in(path);
    const char type = data_type_of(path);
    while (!in.eof()) {
      const uint64_t len = read_record(in, type, K, fd.bases, nullptr);
      if (len > K) fd.rec_len.push_back(len);          // `if(ref.size()>K)` (:450, :512)
      else fd.bases.resize(fd.bases.size() - len);     // len == K: dropped
    }
  } catch (const std::exception& e) {
    fd.error = e.what();
  }
  return fd;
}

}  // namespace

// A batch of entries on their way to the GPU: characters in pinned memory, record boundaries, the
// record -> entry map and the entries' output names.
struct Engine::Batch {
  char* bases = nullptr;
  uint64_t cap = 0, used = 0;
  std::vector<uint64_t> rec_off{0};
  std::vector<uint32_t> rec_entry;
  std::vector<std::string> names;
  uint64_t n_entries = 0;
  uint32_t gid0 = 0;  // id of the batch's first entry (insert batches)

  ~Batch() { nq_host_free(bases); }
  void reserve(uint64_t need) {
    if (used + need <= cap) return;
    uint64_t ncap = std::max<uint64_t>(cap ? cap * 2 : kBatchBases + (64ull << 20), used + need);
    char* nb = static_cast<char*>(nq_host_alloc(ncap));
    if (!nb) throw std::runtime_error(std::string("pinned allocation failed: ") + nq_last_error());
    if (used) memcpy(nb, bases, used);
    nq_host_free(bases);
    bases = nb;
    cap = ncap;
  }
  // appends one entry made of the given records (possibly none)
  void add_entry(const char* data, const std::vector<uint64_t>& rec_len, std::string name) {
    uint64_t total = 0;
    for (uint64_t l : rec_len) total += l;
    reserve(total);
    if (total) memcpy(bases + used, data, total);
    for (uint64_t l : rec_len) {
      used += l;
      rec_off.push_back(used);
      rec_entry.push_back((uint32_t)n_entries);
    }
    names.push_back(std::move(name));
    ++n_entries;
  }
  bool full(uint64_t F) const { return used >= kBatchBases || n_entries * F >= kBatchCells; }
  void clear() {
    used = 0;
    rec_off.assign(1, 0);
    rec_entry.clear();
    names.clear();
    n_entries = 0;
  }
};

// ------------------------------------------------------------------------------------------------
void Engine::init_ctx() {
  int ndev = 0;
  check(nq_device_count(&ndev), "nq_device_count");
  const int n = std::max(1, opt_.gpus);
  if (opt_.device < 0 || opt_.device + n > ndev)
    throw std::runtime_error("--device " + std::to_string(opt_.device) + " --gpus " + std::to_string(n) + ": only " +
                             std::to_string(ndev) + " CUDA device(s) visible");
  sh_.resize(n);
  for (int r = 0; r < n; ++r) check(nq_ctx_create(opt_.device + r, nullptr, &sh_[r].ctx), "nq_ctx_create");
  if (n > 1) {
    std::vector<nq_ctx*> ctxs(n);
    std::vector<nq_comm*> comms(n, nullptr);
    for (int r = 0; r < n; ++r) ctxs[r] = sh_[r].ctx;
    check(nq_comm_init_all(ctxs.data(), n, comms.data()), "nq_comm_init_all");
    for (int r = 0; r < n; ++r) sh_[r].comm = comms[r];
  }
  ins_batches_.resize(n);
  ins_pending_.resize(n);
  for (auto& b : ins_batches_) b.reset(new Batch());
}

// fn(r) for every device, each on its own host thread (a context is driven by one thread at a time;
// the NCCL calls of one collective are made concurrently).  One device: the calling thread.
template <typename Fn>
void Engine::on_all_devices(Fn&& fn) {
  const int n = (int)sh_.size();
  if (n == 1) {
    fn(0);
    return;
  }
  std::vector<std::future<void>> fs;
  fs.reserve(n);
  for (int r = 0; r < n; ++r) fs.push_back(std::async(std::launch::async, [&fn, r] { fn(r); }));
  std::exception_ptr first;
  for (auto& f : fs) {
    try {
      f.get();
    } catch (...) {
      if (!first) first = std::current_exception();
    }
  }
  if (first) std::rethrow_exception(first);
}

Engine::Engine(uint32_t S, uint32_t K, uint32_t W, uint32_t H, const std::string& out_path, double min_fract,
               const EngineOptions& opt)
    : opt_(opt) {
  check(nq_params_init(&p_, K, S, W, H, min_fract), "bad parameters");
  init_ctx();
  out_.open(out_path);  // the reference truncates its output file at construction (:33-37)
}

Engine::~Engine() {
  close_output();
  for (auto& f : ins_pending_)
    if (f.valid()) {
      try { f.get(); } catch (...) {}
    }
  for (Shard& s : sh_) {
    if (s.ix) nq_index_free(s.ix);
    if (s.comm) nq_comm_destroy(s.comm);
    if (s.ctx) {
      if (s.d_store) nq_device_free(s.ctx, s.d_store);
      if (s.d_query) nq_device_free(s.ctx, s.d_query);
      if (s.d_all) nq_device_free(s.ctx, s.d_all);
    }
  }
  ins_batches_.clear();  // pinned buffers go before the contexts
  for (Shard& s : sh_)
    if (s.ctx) nq_ctx_destroy(s.ctx);
}

void Engine::close_output() { out_.close(); }

uint64_t Engine::kernel_launches() const {
  uint64_t t = 0;
  for (const Shard& s : sh_) t += nq_ctx_launch_count(s.ctx);
  return t;
}

void Engine::select_best_H(double genome_size) {
  check(nq_params_select_best_H(&p_, genome_size), "select_best_H");
  std::cout << "I chosed H=" << p_.H << std::endl;  // sic (:137)
}

void Engine::reserve_rows(int r, int32_t*& buf, uint64_t& cap, uint64_t rows) {
  if (rows <= cap) return;
  Shard& s = sh_[r];
  if (buf) nq_device_free(s.ctx, buf);
  buf = nullptr;
  cap = 0;
  void* nd = nullptr;
  const uint64_t ncap = std::max<uint64_t>(rows, 64);
  check(nq_device_alloc(s.ctx, ncap * p_.F * sizeof(int32_t), &nd), "sketch buffer allocation");
  buf = static_cast<int32_t*>(nd);
  cap = ncap;
}

// sketches of batch `b` (gids [gid0, gid0 + n_entries)) appended to device r's store
void Engine::sketch_into_store(int r, Batch& b) {
  Shard& s = sh_[r];
  if (s.store_n + b.n_entries > s.store_cap) {
    const uint64_t ncap = std::max<uint64_t>(s.store_cap * 2, s.store_n + b.n_entries);
    void* nd = nullptr;
    check(nq_device_alloc(s.ctx, ncap * p_.F * sizeof(int32_t), &nd), "sketch store allocation");
    if (s.store_n) check(nq_device_copy(s.ctx, nd, s.d_store, s.store_n * p_.F * sizeof(int32_t), 2), "sketch store copy");
    if (s.d_store) nq_device_free(s.ctx, s.d_store);
    s.d_store = static_cast<int32_t*>(nd);
    s.store_cap = ncap;
  }
  std::vector<uint32_t> flags(b.n_entries, 0);
  check(nq_sketch_records(s.ctx, &p_, b.bases, b.rec_off.data(), b.rec_entry.size(), b.rec_entry.data(), b.n_entries,
                          s.d_store + s.store_n * p_.F, flags.data(), 1),
        "nq_sketch_records");
  for (uint64_t i = 0; i < b.n_entries; ++i)
    if (flags[i] & NQ_ENTRY_DENSIFY_STALLED)
      std::cerr << "warning: densification cannot complete for '" << b.names[i]
                << "' (the reference would not terminate on this entry); empty cells left empty\n";
  s.segs.push_back({b.gid0, (uint32_t)b.n_entries, s.store_n});
  s.store_n += b.n_entries;
  b.clear();
}

Engine::Batch& Engine::insert_batch() {
  if (ins_pending_[ins_cur_].valid()) ins_pending_[ins_cur_].get();  // its previous content has been sketched
  return *ins_batches_[ins_cur_];
}

// The batch that insert_batch() handed out is full: sketch it on its device.  With one device the
// call is synchronous; with several the reader goes on filling the next device's buffer meanwhile.
void Engine::submit_insert(Batch& b) {
  if (b.n_entries == 0) return;
  b.gid0 = genome_numbers_ - (uint32_t)b.n_entries;  // ids are handed out in entry order
  const int r = ins_cur_;
  if (sh_.size() == 1) {
    sketch_into_store(r, b);
    return;
  }
  Batch* bp = &b;
  ins_pending_[r] = std::async(std::launch::async, [this, r, bp] { sketch_into_store(r, *bp); });
  ins_cur_ = (ins_cur_ + 1) % (int)sh_.size();
}

void Engine::wait_inserts() {
  std::exception_ptr first;
  for (auto& f : ins_pending_)
    if (f.valid()) {
      try {
        f.get();
      } catch (...) {
        if (!first) first = std::current_exception();
      }
    }
  if (first) std::rethrow_exception(first);
}

uint64_t Engine::pending_sketches() const {
  uint64_t t = 0;
  for (const Shard& s : sh_) t += s.store_n;
  return t;
}

// Posting lists of all shards in the dump layout (A7): per-list sizes for the F*2^W lists, and the
// lists' gids back to back — shard order == gid order inside every list.
void Engine::export_all(std::vector<uint32_t>& sizes, std::vector<uint32_t>& gids) {
  const uint64_t nlists = (uint64_t)p_.F * (uint64_t)p_.range;
  sizes.assign(nlists, 0);
  gids.clear();
  std::vector<std::vector<uint32_t>> ssz(sh_.size()), sg(sh_.size());
  uint64_t total = 0;
  for (size_t r = 0; r < sh_.size(); ++r) {
    if (!sh_[r].ix) continue;
    uint64_t npost = 0;
    check(nq_index_info(sh_[r].ix, &npost, nullptr, nullptr, nullptr), "nq_index_info");
    ssz[r].resize(nlists);
    sg[r].resize(std::max<uint64_t>(npost, 1));
    check(nq_index_export(sh_[r].ix, ssz[r].data(), sg[r].data(), sg[r].size()), "nq_index_export");
    total += npost;
  }
  size_t live = 0, only = 0;
  for (size_t r = 0; r < sh_.size(); ++r)
    if (sh_[r].ix) { ++live; only = r; }
  if (live == 1) {
    sizes.swap(ssz[only]);
    gids.swap(sg[only]);
    gids.resize(total);
    return;
  }
  gids.reserve(total);
  std::vector<uint64_t> at(sh_.size(), 0);
  for (uint64_t l = 0; l < nlists; ++l)
    for (size_t r = 0; r < sh_.size(); ++r) {
      if (!sh_[r].ix) continue;
      const uint32_t sz = ssz[r][l];
      if (sz) {
        gids.insert(gids.end(), sg[r].begin() + at[r], sg[r].begin() + at[r] + sz);
        at[r] += sz;
        sizes[l] += sz;
      }
    }
}

// (sizes, gids) over gids [0, genome_numbers_) -> one shard per device (nq_index_import keeps the
// gids of its own block)
void Engine::import_all(const std::vector<uint32_t>& sizes, const std::vector<uint32_t>& gids) {
  static const uint32_t zero = 0;
  const uint32_t* gp = gids.empty() ? &zero : gids.data();
  const int n = (int)sh_.size();
  on_all_devices([&](int r) {
    Shard& s = sh_[r];
    if (s.ix) {
      nq_index_free(s.ix);
      s.ix = nullptr;
    }
    uint64_t b = 0, e = 0;
    check(nq_shard_range(genome_numbers_, n, r, &b, &e), "nq_shard_range");
    s.gid0 = (uint32_t)b;
    s.n = (uint32_t)(e - b);
    if (s.n) check(nq_index_import(s.ctx, &p_, sizes.data(), gp, s.n, s.gid0, &s.ix), "nq_index_import");
  });
  have_index_ = genome_numbers_ != 0;
}

// Posting lists over every genome id handed out so far, sharded over the devices by contiguous
// gid blocks.  Fresh index: every shard's sketches are collected on its device (they were
// sketched round-robin) and built there.  Index that already exists (--load, or an earlier query)
// and then received more entries: the old lists are merged with the new sketches' postings on the
// host and re-imported (new gids are larger than every old one, so appending keeps the lists
// gid-ascending like push_back does, :366).
void Engine::build_index() {
  wait_inserts();
  const uint64_t fresh = pending_sketches();
  if (fresh == 0) return;
  const int n = (int)sh_.size();
  const uint64_t F = p_.F;
  if (!have_index_) {
    // all pending sketches cover gids [0, genome_numbers_)
    on_all_devices([&](int r) {
      Shard& s = sh_[r];
      uint64_t b = 0, e = 0;
      check(nq_shard_range(genome_numbers_, n, r, &b, &e), "nq_shard_range");
      s.gid0 = (uint32_t)b;
      s.n = (uint32_t)(e - b);
      if (!s.n) return;
      const int32_t* src = nullptr;
      int32_t* own = nullptr;
      if (n == 1 && s.segs.size() >= 1 && s.store_n == s.n) {
        src = s.d_store;  // one device: the store already is the shard, in gid order
      } else {
        void* nd = nullptr;
        check(nq_device_alloc(s.ctx, (uint64_t)s.n * F * sizeof(int32_t), &nd), "shard sketch allocation");
        own = static_cast<int32_t*>(nd);
        for (int t = 0; t < n; ++t)
          for (const Segment& g : sh_[t].segs) {
            const uint64_t lo = std::max<uint64_t>(g.gid0, b), hi = std::min<uint64_t>((uint64_t)g.gid0 + g.count, e);
            if (lo >= hi) continue;
            const int32_t* from = sh_[t].d_store + (g.row + (lo - g.gid0)) * F;
            int32_t* to = own + (lo - b) * F;
            const size_t bytes = (hi - lo) * F * sizeof(int32_t);
            if (t == r) check(nq_device_copy(s.ctx, to, from, bytes, 2), "sketch copy");
            else check(nq_device_copy_peer(s.ctx, to, sh_[t].ctx, from, bytes), "sketch peer copy");
          }
        src = own;
      }
      check(nq_index_build_device(s.ctx, &p_, src, s.n, s.gid0, &s.ix), "nq_index_build_device");
      if (own) nq_device_free(s.ctx, own);
    });
    have_index_ = true;
  } else {
    const uint64_t range = (uint64_t)p_.range, nlists = F * range;
    std::vector<uint32_t> sizes, gids;
    export_all(sizes, gids);
    // the new sketches, in gid order, on the host
    struct Seg { uint32_t gid0, count; int dev; uint64_t row; };
    std::vector<Seg> segs;
    for (int t = 0; t < n; ++t)
      for (const Segment& g : sh_[t].segs) segs.push_back({g.gid0, g.count, t, g.row});
    std::sort(segs.begin(), segs.end(), [](const Seg& x, const Seg& y) { return x.gid0 < y.gid0; });
    const uint32_t base = segs.front().gid0;
    std::vector<int32_t> sk(fresh * F);
    for (const Seg& g : segs)
      check(nq_device_copy(sh_[g.dev].ctx, sk.data() + (uint64_t)(g.gid0 - base) * F, sh_[g.dev].d_store + g.row * F,
                           (uint64_t)g.count * F * sizeof(int32_t), 1),
            "sketch download");
    std::vector<uint32_t> nsizes(nlists), ngids;
    ngids.reserve(gids.size() + fresh * F);
    std::vector<std::pair<uint32_t, uint32_t>> add;  // (fp, gid) of the new entries in one cell
    uint64_t rpos = 0;
    for (uint64_t c = 0; c < F; ++c) {
      add.clear();
      for (uint64_t g = 0; g < fresh; ++g) {
        const int32_t fp = sk[g * F + c];
        if (fp >= 0 && fp < p_.range) add.emplace_back((uint32_t)fp, base + (uint32_t)g);  // :364
      }
      std::stable_sort(add.begin(), add.end(), [](const auto& a, const auto& b2) { return a.first < b2.first; });
      size_t ai = 0;
      for (uint64_t f = 0; f < range; ++f) {
        const uint32_t sz = sizes[c * range + f];
        ngids.insert(ngids.end(), gids.begin() + rpos, gids.begin() + rpos + sz);
        rpos += sz;
        uint32_t extra = 0;
        while (ai < add.size() && add[ai].first == f) {
          ngids.push_back(add[ai].second);
          ++ai;
          ++extra;
        }
        nsizes[c * range + f] = sz + extra;
      }
    }
    import_all(nsizes, ngids);
  }
  indexed_ = genome_numbers_;
  // the stores have been consumed: later insertions start new ones
  for (Shard& s : sh_) {
    s.segs.clear();
    s.store_n = 0;
    if (s.d_store) {
      nq_device_free(s.ctx, s.d_store);
      s.d_store = nullptr;
      s.store_cap = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// insert_file_of_file_whole (:461-500): one entry per listed file.  The list is read with a plain
// ifstream in the reference; lines of size <= 2 or naming a missing file are skipped (:481-482).
void Engine::insert_file_of_file_whole(const std::string& fof) {
  std::ifstream in(fof);
  if (!in) {
    std::cout << "Unable to open the file '" << fof << "'" << std::endl;
    exit(0);  // as the reference does (:464-467)
  }
  std::vector<std::string> files;
  std::string line;
  while (std::getline(in, line))
    if (line.size() > 2 && exists_test(line)) files.push_back(line);

  unsigned nthreads = opt_.reader_threads ? opt_.reader_threads : std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
  const size_t window = std::max<size_t>(2, (size_t)nthreads * 2);
  std::deque<std::future<FileData>> inflight;
  size_t next = 0;
  Batch* batch = &insert_batch();
  const uint32_t K = p_.K;
  for (size_t i = 0; i < files.size(); ++i) {
    while (next < files.size() && inflight.size() < window) {
      inflight.push_back(std::async(std::launch::async, read_whole_file, files[next], K));
      ++next;
    }
    FileData fd = inflight.front().get();
    inflight.pop_front();
    if (!fd.error.empty()) throw std::runtime_error(files[i] + ": " + fd.error);  // zstr would throw as well
    if (fd.rec_len.size() > 1 && opt_.verbose)
      std::cerr << "note: '" << files[i] << "' has " << fd.rec_len.size()
                << " records; they are min-merged into one sketch (the reference does not terminate on such files)\n";
    filenames_.push_back(files[i]);
    ++genome_numbers_;
    batch->add_entry(fd.bases.data(), fd.rec_len, files[i]);
    if (batch->full(p_.F)) {
      submit_insert(*batch);
      batch = &insert_batch();
    }
  }
  submit_insert(*batch);
  wait_inserts();
}

// insert_file_lines (:383-408): one entry per record with len > K; name = header line as read.
void Engine::insert_file_lines(const std::string& file) {
  const char type = data_type_of(file);
  LineReader in(file);
  Batch* batch = &insert_batch();
  std::vector<char> rec;
  std::string header;
  std::vector<uint64_t> one(1);
  while (!in.eof()) {
    rec.clear();
    const uint64_t len = read_record(in, type, p_.K, rec, &header);
    if (len > p_.K) {
      one[0] = len;
      filenames_.push_back(header);
      ++genome_numbers_;
      batch->add_entry(rec.data(), one, header);
      if (batch->full(p_.F)) {
        submit_insert(*batch);
        batch = &insert_batch();
      }
    }
  }
  submit_insert(*batch);
  wait_inserts();
}

// ------------------------------------------------------------------------------------------------
const std::string& Engine::frac_text(uint32_t count) {
  // `(double)count/F` through ostream's default formatting == "%g" (6 significant digits)
  if (frac_cache_.empty()) frac_cache_.resize((size_t)p_.F + 1);
  static thread_local std::string overflow;
  char buf[64];
  if (count > p_.F) {  // cannot happen for real counts; keeps the function total
    snprintf(buf, sizeof buf, "%g", (double)count / (double)p_.F);
    overflow = buf;
    return overflow;
  }
  std::string& s = frac_cache_[count];
  if (s.empty()) {
    snprintf(buf, sizeof buf, "%g", (double)count / (double)p_.F);
    s = buf;
  }
  return s;
}

// output_query (:544-566)
void Engine::write_hits(const std::string& name, const uint32_t* counts, const uint32_t* gids, uint64_t n) {
  if (!opt_.binary_output) {
    std::string line = name;
    line += ' ';
    for (uint64_t i = 0; i < n; ++i) {
      line += filenames_[gids[i]];
      line += ':';
      line += frac_text(counts[i]);
      line += ' ';
    }
    line += '\n';
    out_.write(line);
  } else {
    out_.write(name);
    out_.write("\n", 1);
    out_.put_u32((uint32_t)n);
    for (uint64_t i = 0; i < n; ++i) {
      out_.put_u32(gids[i]);
      out_.put_u32(counts[i]);
    }
  }
}

// One query batch: its entries are cut into one contiguous slice per device, every device sketches
// its slice, the sketches are all-gathered (NCCL), every shard counts all of them, and the per-shard
// hit lists are merged on the host (disjoint gids: concatenate + sort, :685).
void Engine::flush_query(Batch& b) {
  if (b.n_entries == 0) return;
  build_index();
  const int n = (int)sh_.size();
  const uint64_t ne = b.n_entries, m = (ne + n - 1) / n, F = p_.F;  // rows per device (the last slices may be short)
  if (!have_index_) {
    // nothing indexed: the reference scans an empty table and reports nothing (count >= min_score
    // holds for no genome because there are none).  The sketches are still computed (and flagged).
    for (uint64_t i = 0; i < ne; ++i) write_hits(b.names[i], nullptr, nullptr, 0);
    b.clear();
    return;
  }
  // record range of every entry (records of one entry are contiguous, entries without records are legal)
  std::vector<uint64_t> first_rec(ne + 1, b.rec_entry.size());
  for (uint64_t r = b.rec_entry.size(); r-- > 0;) first_rec[b.rec_entry[r]] = r;
  for (uint64_t e = ne; e-- > 0;)
    if (first_rec[e] > first_rec[e + 1]) first_rec[e] = first_rec[e + 1];
  std::vector<nq_hits*> part(n, nullptr);
  on_all_devices([&](int r) {
    Shard& s = sh_[r];
    const uint64_t e0 = std::min<uint64_t>(ne, (uint64_t)r * m), e1 = std::min<uint64_t>(ne, e0 + m), cnt = e1 - e0;
    reserve_rows(r, s.d_query, s.query_cap, m);
    if (cnt < m) check(nq_device_fill(s.ctx, s.d_query + cnt * F, 0xFF, (m - cnt) * F * sizeof(int32_t)), "nq_device_fill");
    if (cnt) {
      const uint64_t r0 = first_rec[e0], r1 = first_rec[e1];
      std::vector<uint32_t> ent(r1 - r0), flags(cnt, 0);
      for (uint64_t i = r0; i < r1; ++i) ent[i - r0] = b.rec_entry[i] - (uint32_t)e0;
      check(nq_sketch_records(s.ctx, &p_, b.bases, b.rec_off.data() + r0, r1 - r0, ent.data(), cnt, s.d_query, flags.data(), 1),
            "nq_sketch_records");
    }
    const int32_t* all = s.d_query;
    if (n > 1) {
      reserve_rows(r, s.d_all, s.all_cap, m * n);
      check(nq_allgather_sketches(s.comm, &p_, s.d_query, m, s.d_all), "nq_allgather_sketches");
      all = s.d_all;
    }
    if (s.ix) {
      check(nq_query_batch_device(s.ix, all, m * n, p_.min_score, &part[r]), "nq_query_batch_device");
    } else {  // a shard without genomes (fewer genomes than devices)
      std::vector<uint64_t> zp(m * n + 1, 0);
      check(nq_hits_from_arrays(zp.data(), nullptr, nullptr, m * n, &part[r]), "nq_hits_from_arrays");
    }
  });
  nq_hits* hits = part[0];
  if (n > 1) {
    hits = nullptr;
    check(nq_hits_merge(part.data(), n, &hits), "nq_hits_merge");
    for (nq_hits* h : part) nq_hits_free(h);
  }
  const uint64_t* ptr = nq_hits_ptr(hits);
  const uint32_t* counts = nq_hits_counts(hits);
  const uint32_t* gids = nq_hits_gids(hits);
  for (uint64_t i = 0; i < ne; ++i) write_hits(b.names[i], counts + ptr[i], gids + ptr[i], ptr[i + 1] - ptr[i]);
  nq_hits_free(hits);
  b.clear();
}

// query_file_of_file_whole (:523-540): the list itself may be gzip-compressed (it is read through
// zstr); every line naming an existing file is one query, reported under the line's text (:518).
void Engine::query_file_of_file_whole(const std::string& fof) {
  LineReader in(fof);
  std::vector<std::string> files;
  std::string line;
  while (!in.eof()) {
    in.getline(line);
    if (exists_test(line)) files.push_back(line);
  }
  unsigned nthreads = opt_.reader_threads ? opt_.reader_threads : std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
  const size_t window = std::max<size_t>(2, (size_t)nthreads * 2);
  std::deque<std::future<FileData>> inflight;
  size_t next = 0;
  Batch batch;
  for (size_t i = 0; i < files.size(); ++i) {
    while (next < files.size() && inflight.size() < window) {
      inflight.push_back(std::async(std::launch::async, read_whole_file, files[next], p_.K));
      ++next;
    }
    FileData fd = inflight.front().get();
    inflight.pop_front();
    if (!fd.error.empty()) throw std::runtime_error(files[i] + ": " + fd.error);
    batch.add_entry(fd.bases.data(), fd.rec_len, files[i]);
    if (batch.full(p_.F)) flush_query(batch);
  }
  flush_query(batch);
}

// query_file_lines (:412-430)
void Engine::query_file_lines(const std::string& file) {
  const char type = data_type_of(file);
  LineReader in(file);
  Batch batch;
  std::vector<char> rec;
  std::string header;
  std::vector<uint64_t> one(1);
  while (!in.eof()) {
    rec.clear();
    const uint64_t len = read_record(in, type, p_.K, rec, &header);
    if (len > p_.K) {
      one[0] = len;
      batch.add_entry(rec.data(), one, header);
      if (batch.full(p_.F)) flush_query(batch);
    }
  }
  flush_query(batch);
}

// query_matrix + query_range + output_matrix (:570-628, :747-763).  Counters are uint16_t in the
// reference for every S, i.e. values are taken mod 65536 before the threshold;
// --nowrap keeps all 32 bits.
void Engine::query_matrix() {
  build_index();
  std::string line = "##Names\t";
  for (const std::string& nm : filenames_) {
    line += nm;
    line += '\t';
  }
  line += '\n';
  out_.write(line);
  const uint32_t n = genome_numbers_;
  if (n == 0 || !have_index_) return;
  const int nd = (int)sh_.size();
  const int wrap = opt_.matrix_nowrap ? 0 : 1;
  // rows travel in blocks small enough for the host (the reference's 10 000-row batches only bound
  // its own memory; results do not depend on the batch size)
  const uint32_t block = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(kBufferSize, (256ull << 20) / ((uint64_t)n * 4)));
  auto write_rows = [&](uint32_t q0, uint32_t rows, const std::vector<std::vector<uint32_t>>& tiles) {
    for (uint32_t q = 0; q < rows; ++q) {
      line = filenames_[q0 + q];
      line += '\t';
      for (int r = 0; r < nd; ++r) {
        const uint32_t nr = sh_[r].n;
        const uint32_t* row = tiles[r].data() + (size_t)q * nr;
        for (uint32_t j = 0; j < nr; ++j) {
          if (row[j] >= p_.min_score && row[j] != 0) line += frac_text(row[j]);
          else line += '0';
          line += '\t';
        }
      }
      line += '\n';
      out_.write(line);
    }
  };
  std::vector<std::vector<uint32_t>> tiles(nd);
  if (nd == 1) {
    tiles[0].resize((size_t)block * n);
    for (uint32_t b = 0; b < n; b += block) {
      const uint32_t e = std::min(n, b + block);
      check(nq_matrix_rows(sh_[0].ix, b, e, wrap, tiles[0].data()), "nq_matrix_rows");
      write_rows(b, e - b, tiles);
    }
    return;
  }
  // genome x genome grid tiled over the devices: the rows of shard `owner` are rebuilt from its
  // posting lists block by block, broadcast, and every device counts them against its own columns
  for (int r = 0; r < nd; ++r) tiles[r].resize((size_t)block * std::max<uint32_t>(sh_[r].n, 1));
  for (int owner = 0; owner < nd; ++owner)
    for (uint32_t b = 0; b < sh_[owner].n; b += block) {
      const uint32_t rows = std::min(sh_[owner].n - b, block);
      on_all_devices([&](int r) {
        Shard& s = sh_[r];
        reserve_rows(r, s.d_all, s.all_cap, rows);
        if (r == owner) check(nq_index_sketches_device(s.ix, b, b + rows, s.d_all), "nq_index_sketches_device");
        check(nq_bcast_sketches(s.comm, &p_, s.d_all, rows, owner), "nq_bcast_sketches");
        if (s.ix) check(nq_matrix_tile(s.ix, s.d_all, rows, wrap, tiles[r].data()), "nq_matrix_tile");
      });
      write_rows(sh_[owner].gid0 + b, rows, tiles);
    }
}

// ------------------------------------------------------------------------------------------------
// dump_index_disk (:42-59): u32 lF,K,H,W,min_score,genome_numbers; per list u32 size + gids;
// names, one per line.  gzip-compressed like every zstr::ofstream.
void Engine::dump_index_disk(const std::string& path) {
  build_index();
  GzWriter dump;
  dump.open(path);
  const uint32_t hdr[6] = {p_.S, p_.K, p_.H, p_.W, p_.min_score, genome_numbers_};
  dump.write(hdr, sizeof hdr);
  const uint64_t nlists = (uint64_t)p_.F * (uint64_t)p_.range;
  if (have_index_) {
    std::vector<uint32_t> sizes, gids;
    export_all(sizes, gids);
    std::vector<uint32_t> buf;
    buf.reserve((1u << 20) + 16);
    uint64_t r = 0;
    for (uint64_t l = 0; l < nlists; ++l) {
      const uint32_t sz = sizes[l];
      buf.push_back(sz);
      buf.insert(buf.end(), gids.begin() + r, gids.begin() + r + sz);
      r += sz;
      if (buf.size() >= (1u << 20)) {
        dump.write(buf.data(), buf.size() * 4);
        buf.clear();
      }
    }
    dump.write(buf.data(), buf.size() * 4);
  } else {
    std::vector<uint32_t> zeros(1u << 20, 0);
    for (uint64_t l = 0; l < nlists; l += zeros.size())
      dump.write(zeros.data(), std::min<uint64_t>(zeros.size(), nlists - l) * 4);
  }
  for (const std::string& n : filenames_) {
    dump.write(n);
    dump.write("\n", 1);
  }
  dump.close();
}

namespace {
// sequential binary reads over gzread()
class BinReader {
 public:
  explicit BinReader(const std::string& path) : buf_(4 << 20) {
    gz_ = gzopen(path.c_str(), "rb");
    if (!gz_) throw std::runtime_error("cannot open '" + path + "'");
    gzbuffer(gz_, 1 << 20);
  }
  ~BinReader() { gzclose(gz_); }
  void read(void* dst, size_t n) {
    char* d = static_cast<char*>(dst);
    while (n) {
      if (pos_ >= len_) {
        const int got = gzread(gz_, buf_.data(), (unsigned)buf_.size());
        if (got <= 0) throw std::runtime_error("index file is truncated");
        pos_ = 0;
        len_ = (size_t)got;
      }
      const size_t take = std::min(n, len_ - pos_);
      memcpy(d, buf_.data() + pos_, take);
      d += take;
      pos_ += take;
      n -= take;
    }
  }
  bool getline(std::string& out) {
    out.clear();
    bool any = false;
    for (;;) {
      if (pos_ >= len_) {
        const int got = gzread(gz_, buf_.data(), (unsigned)buf_.size());
        if (got <= 0) return any;
        pos_ = 0;
        len_ = (size_t)got;
      }
      any = true;
      const char c = buf_[pos_++];
      if (c == '\n') return true;
      out.push_back(c);
    }
  }

 private:
  gzFile gz_;
  std::vector<char> buf_;
  size_t pos_ = 0, len_ = 0;
};
}  // namespace

// load constructor (:63-102): masks are re-derived from the stored H, min_score comes from the
// file (so --minjac is ignored with --load, as in the reference).
Engine::Engine(const std::string& dump_path, const std::string& out_path, const EngineOptions& opt) : opt_(opt) {
  BinReader in(dump_path);
  uint32_t hdr[6];
  in.read(hdr, sizeof hdr);
  check(nq_params_init(&p_, hdr[1], hdr[0], hdr[3], hdr[2], 0.0), "bad parameters in index file");
  p_.min_score = hdr[4];
  genome_numbers_ = hdr[5];
  init_ctx();
  const uint64_t nlists = (uint64_t)p_.F * (uint64_t)p_.range;
  std::vector<uint32_t> sizes(nlists), gids;
  for (uint64_t l = 0; l < nlists; ++l) {
    uint32_t sz;
    in.read(&sz, 4);
    sizes[l] = sz;
    if (sz) {
      const size_t at = gids.size();
      gids.resize(at + sz);
      in.read(gids.data() + at, (size_t)sz * 4);
    }
  }
  std::string name;
  for (uint32_t g = 0; g < genome_numbers_; ++g) {
    in.getline(name);
    filenames_.push_back(name);
  }
  if (genome_numbers_) import_all(sizes, gids);
  indexed_ = genome_numbers_;
  out_.open(out_path);
}

}  // namespace nqh
