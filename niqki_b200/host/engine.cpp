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
  check(nq_ctx_create(opt_.device, nullptr, &ctx_), "nq_ctx_create");
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
  if (ix_) nq_index_free(ix_);
  if (ctx_) {
    if (d_store_) nq_device_free(ctx_, d_store_);
    if (d_query_) nq_device_free(ctx_, d_query_);
    nq_ctx_destroy(ctx_);
  }
}

void Engine::close_output() { out_.close(); }

uint64_t Engine::kernel_launches() const { return nq_ctx_launch_count(ctx_); }

void Engine::select_best_H(double genome_size) {
  check(nq_params_select_best_H(&p_, genome_size), "select_best_H");
  std::cout << "I chosed H=" << p_.H << std::endl;  // sic (:137)
}

void Engine::ensure_store(uint64_t extra) {
  if (store_n_ + extra <= store_cap_) return;
  const uint64_t ncap = std::max<uint64_t>(store_cap_ * 2, store_n_ + extra);
  void* nd = nullptr;
  check(nq_device_alloc(ctx_, ncap * p_.F * sizeof(int32_t), &nd), "sketch store allocation");
  if (store_n_) check(nq_device_copy(ctx_, nd, d_store_, store_n_ * p_.F * sizeof(int32_t), 2), "sketch store copy");
  if (d_store_) nq_device_free(ctx_, d_store_);
  d_store_ = static_cast<int32_t*>(nd);
  store_cap_ = ncap;
}

void Engine::flush_insert(Batch& b) {
  if (b.n_entries == 0) return;
  ensure_store(b.n_entries);
  std::vector<uint32_t> flags(b.n_entries, 0);
  check(nq_sketch_records(ctx_, &p_, b.bases, b.rec_off.data(), b.rec_entry.size(), b.rec_entry.data(), b.n_entries,
                          d_store_ + store_n_ * p_.F, flags.data(), 1),
        "nq_sketch_records");
  for (uint64_t i = 0; i < b.n_entries; ++i)
    if (flags[i] & NQ_ENTRY_DENSIFY_STALLED)
      std::cerr << "warning: densification cannot complete for '" << b.names[i]
                << "' (the reference would not terminate on this entry); empty cells left empty\n";
  store_n_ += b.n_entries;
  b.clear();
}

// Posting lists over every genome id handed out so far.  Fresh index: one build from the sketch
// store.  Index that came from --load and then received more entries: the old lists are merged
// with the new sketches' postings on the host and re-imported (new gids are larger than every
// loaded one, so appending keeps lists gid-ascending like push_back does, :366).
void Engine::build_index() {
  if (store_n_ == 0 && (ix_ || genome_numbers_ == 0)) return;
  if (!ix_) {
    if (store_n_ == 0) return;
    check(nq_index_build_device(ctx_, &p_, d_store_, store_n_, store_base_, &ix_), "nq_index_build_device");
  } else {
    const uint64_t F = p_.F, range = (uint64_t)p_.range, nlists = F * range;
    uint64_t old_post = 0;
    check(nq_index_info(ix_, &old_post, nullptr, nullptr, nullptr), "nq_index_info");
    std::vector<uint32_t> sizes(nlists), gids(std::max<uint64_t>(old_post, 1));
    check(nq_index_export(ix_, sizes.data(), gids.data(), gids.size()), "nq_index_export");
    std::vector<int32_t> sk(store_n_ * F);
    check(nq_device_copy(ctx_, sk.data(), d_store_, sk.size() * sizeof(int32_t), 1), "sketch download");
    std::vector<uint32_t> nsizes(nlists), ngids;
    ngids.reserve(old_post + store_n_ * F);
    std::vector<std::pair<uint32_t, uint32_t>> add;  // (fp, gid) of the new entries in one cell
    uint64_t r = 0;
    for (uint64_t c = 0; c < F; ++c) {
      add.clear();
      for (uint64_t g = 0; g < store_n_; ++g) {
        const int32_t fp = sk[g * F + c];
        if (fp >= 0 && fp < p_.range) add.emplace_back((uint32_t)fp, store_base_ + (uint32_t)g);  // :364
      }
      std::stable_sort(add.begin(), add.end(), [](const auto& a, const auto& b2) { return a.first < b2.first; });
      size_t ai = 0;
      for (uint64_t f = 0; f < range; ++f) {
        const uint32_t sz = sizes[c * range + f];
        ngids.insert(ngids.end(), gids.begin() + r, gids.begin() + r + sz);
        r += sz;
        uint32_t extra = 0;
        while (ai < add.size() && add[ai].first == f) {
          ngids.push_back(add[ai].second);
          ++ai;
          ++extra;
        }
        nsizes[c * range + f] = sz + extra;
      }
    }
    nq_index* merged = nullptr;
    if (ngids.empty()) ngids.push_back(0);
    check(nq_index_import(ctx_, &p_, nsizes.data(), ngids.data(), genome_numbers_, 0, &merged), "nq_index_import");
    nq_index_free(ix_);
    ix_ = merged;
  }
  indexed_ = genome_numbers_;
  // the store has been consumed: later insertions start a new one
  store_base_ = genome_numbers_;
  store_n_ = 0;
  if (d_store_) {
    nq_device_free(ctx_, d_store_);
    d_store_ = nullptr;
    store_cap_ = 0;
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
  if (ix_ && store_n_ == 0) store_base_ = genome_numbers_;
  std::vector<std::string> files;
  std::string line;
  while (std::getline(in, line))
    if (line.size() > 2 && exists_test(line)) files.push_back(line);

  unsigned nthreads = opt_.reader_threads ? opt_.reader_threads : std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
  const size_t window = std::max<size_t>(2, (size_t)nthreads * 2);
  std::deque<std::future<FileData>> inflight;
  size_t next = 0;
  Batch batch;
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
    batch.add_entry(fd.bases.data(), fd.rec_len, files[i]);
    if (batch.full(p_.F)) flush_insert(batch);
  }
  flush_insert(batch);
}

// insert_file_lines (:383-408): one entry per record with len > K; name = header line as read.
void Engine::insert_file_lines(const std::string& file) {
  const char type = data_type_of(file);
  LineReader in(file);
  if (ix_ && store_n_ == 0) store_base_ = genome_numbers_;
  Batch batch;
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
      batch.add_entry(rec.data(), one, header);
      if (batch.full(p_.F)) flush_insert(batch);
    }
  }
  flush_insert(batch);
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

void Engine::flush_query(Batch& b) {
  if (b.n_entries == 0) return;
  build_index();
  if (b.n_entries > query_cap_) {
    if (d_query_) nq_device_free(ctx_, d_query_);
    d_query_ = nullptr;
    void* nd = nullptr;
    const uint64_t cap = std::max<uint64_t>(b.n_entries, 64);
    check(nq_device_alloc(ctx_, cap * p_.F * sizeof(int32_t), &nd), "query sketch allocation");
    d_query_ = static_cast<int32_t*>(nd);
    query_cap_ = cap;
  }
  std::vector<uint32_t> flags(b.n_entries, 0);
  check(nq_sketch_records(ctx_, &p_, b.bases, b.rec_off.data(), b.rec_entry.size(), b.rec_entry.data(), b.n_entries,
                          d_query_, flags.data(), 1),
        "nq_sketch_records");
  if (!ix_) {
    // nothing indexed: the reference scans an empty table and reports nothing (count >= min_score
    // holds for no genome because there are none)
    for (uint64_t i = 0; i < b.n_entries; ++i) write_hits(b.names[i], nullptr, nullptr, 0);
    b.clear();
    return;
  }
  nq_hits* hits = nullptr;
  check(nq_query_batch_device(ix_, d_query_, b.n_entries, p_.min_score, &hits), "nq_query_batch_device");
  const uint64_t* ptr = nq_hits_ptr(hits);
  const uint32_t* counts = nq_hits_counts(hits);
  const uint32_t* gids = nq_hits_gids(hits);
  for (uint64_t i = 0; i < b.n_entries; ++i) write_hits(b.names[i], counts + ptr[i], gids + ptr[i], ptr[i + 1] - ptr[i]);
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
  for (const std::string& n : filenames_) {
    line += n;
    line += '\t';
  }
  line += '\n';
  out_.write(line);
  const uint32_t n = genome_numbers_;
  if (n == 0 || !ix_) return;
  // rows travel in slabs small enough for the host (the reference's 10 000-row batches only bound
  // its own memory; results do not depend on the batch size)
  const uint32_t slab = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(kBufferSize, (256ull << 20) / ((uint64_t)n * 4)));
  std::vector<uint32_t> counts((size_t)slab * n);
  for (uint32_t b = 0; b < n; b += slab) {
    const uint32_t e = std::min(n, b + slab);
    check(nq_matrix_rows(ix_, b, e, opt_.matrix_nowrap ? 0 : 1, counts.data()), "nq_matrix_rows");
    for (uint32_t q = b; q < e; ++q) {
      line = filenames_[q];
      line += '\t';
      const uint32_t* row = counts.data() + (size_t)(q - b) * n;
      for (uint32_t j = 0; j < n; ++j) {
        if (row[j] >= p_.min_score && row[j] != 0) line += frac_text(row[j]);
        else line += '0';
        line += '\t';
      }
      line += '\n';
      out_.write(line);
    }
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
  if (ix_) {
    uint64_t npost = 0;
    check(nq_index_info(ix_, &npost, nullptr, nullptr, nullptr), "nq_index_info");
    std::vector<uint32_t> sizes(nlists), gids(std::max<uint64_t>(npost, 1));
    check(nq_index_export(ix_, sizes.data(), gids.data(), gids.size()), "nq_index_export");
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
  if (genome_numbers_) {
    if (gids.empty()) gids.push_back(0);
    check(nq_index_import(ctx_, &p_, sizes.data(), gids.data(), genome_numbers_, 0, &ix_), "nq_index_import");
  }
  indexed_ = genome_numbers_;
  store_base_ = genome_numbers_;
  out_.open(out_path);
}

}  // namespace nqh
