// ref_shim.cpp — TEST INFRASTRUCTURE ONLY.
//
// A C-ABI window onto the UNMODIFIED reference: this file includes the reference's own header and
// is compiled together with /root/reference/src/niqki_index.cpp where it lies (recipe:
// oracle/Makefile, target `ref`; outputs only under oracle/_ref/, which is git-ignored).  Nothing
// from the reference tree is copied into this repository.  All members of class Index are public
// (src/niqki_index.h:36), so the shim simply forwards to them.
//
// Used by tests/ to pin the C restatement in niqki_oracle.c, by tests/golden/make_golden.py to
// generate the committed fixtures, and by bench.py as the "reference" CPU baseline.
#include "niqki_index.h"

#include <chrono>
#include <cstring>

extern "C" {

// Index::Index(lF,K,W,H,filename,min_fract) — src/niqki_index.cpp:13-38.  `out_path` receives
// whatever the reference writes to its output stream.
void* ref_index_new(uint32_t S, uint32_t K, uint32_t W, uint32_t H, const char* out_path, double J) {
  return new Index(S, K, W, H, std::string(out_path), J);
}
void ref_index_free(void* h) { delete static_cast<Index*>(h); }

void ref_select_best_H(void* h, double genome_size) { static_cast<Index*>(h)->select_best_H(genome_size); }

// K,S,W,H,M,F,mask_M,maxrem,range,min_score — same order as nqo_params
void ref_get_params(void* h, uint32_t* out) {
  Index* ix = static_cast<Index*>(h);
  out[0] = ix->K; out[1] = ix->lF; out[2] = ix->W; out[3] = ix->H; out[4] = ix->M; out[5] = ix->F;
  out[6] = ix->mask_M; out[7] = ix->maximal_remainder; out[8] = (uint32_t)ix->fingerprint_range;
  out[9] = ix->min_score;
}

uint64_t ref_revhash64(void* h, uint64_t x) { return static_cast<Index*>(h)->revhash64(x); }
uint64_t ref_unrevhash64(void* h, uint64_t x) { return static_cast<Index*>(h)->unrevhash64(x); }
uint64_t ref_hash_family(void* h, uint64_t x, uint32_t f) { return static_cast<Index*>(h)->hash_family(x, f); }
int32_t ref_get_fingerprint(void* h, uint64_t x) { return static_cast<Index*>(h)->get_fingerprint(x); }
uint64_t ref_str2numstrand(void* h, const char* s, size_t n) { return static_cast<Index*>(h)->str2numstrand(std::string(s, n)); }
uint64_t ref_rcb(void* h, uint64_t x) { return static_cast<Index*>(h)->rcb(x); }

// Index::compute_sketch — src/niqki_index.cpp:335-358; sketch is F cells in/out
void ref_compute_sketch(void* h, const char* seq, size_t len, int32_t* sketch) {
  Index* ix = static_cast<Index*>(h);
  std::string s(seq, len);
  std::vector<int32_t> sk(sketch, sketch + ix->F);
  ix->compute_sketch(s, sk);
  std::memcpy(sketch, sk.data(), sizeof(int32_t) * ix->F);
}

// Sketch n entries with the reference's compute_sketch under OpenMP (one entry per task, like
// src/niqki_index.cpp:386-407); returns wall seconds of the sketching only.
double ref_sketch_batch(void* h, const char* bases, const uint64_t* offsets, size_t n, int32_t* out,
                        int nthreads) {
  Index* ix = static_cast<Index*>(h);
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  std::vector<std::string> seqs(n);
  for (size_t e = 0; e < n; ++e) seqs[e].assign(bases + offsets[e], bases + offsets[e + 1]);
  auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (long e = 0; e < (long)n; ++e) {
    std::vector<int32_t> sk;
    if (seqs[e].size() > ix->K) ix->compute_sketch(seqs[e], sk);
    else sk.assign(ix->F, -1);
    if (out) std::memcpy(out + (size_t)e * ix->F, sk.data(), sizeof(int32_t) * ix->F);
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// Index::insert_sketch (:362-370) plus the bookkeeping its callers do (:398-400, :483-486)
void ref_insert_sketch(void* h, const int32_t* sketch, uint32_t gid, const char* name) {
  Index* ix = static_cast<Index*>(h);
  std::vector<int32_t> sk(sketch, sketch + ix->F);
  ix->insert_sketch(sk, gid);
  if (gid + 1 > ix->genome_numbers) ix->genome_numbers = gid + 1;
  if (ix->filenames.size() < ix->genome_numbers) ix->filenames.resize(ix->genome_numbers);
  ix->filenames[gid] = name ? name : "";
}

uint32_t ref_num_genomes(void* h) { return static_cast<Index*>(h)->genome_numbers; }

// list sizes for every (cell,fp) row and the concatenated gids, read straight from Buckets
uint64_t ref_export_postings(void* h, uint32_t* sizes, uint32_t* gids, uint64_t cap) {
  Index* ix = static_cast<Index*>(h);
  const uint64_t rows = (uint64_t)ix->fingerprint_range * ix->F;
  uint64_t n = 0;
  for (uint64_t r = 0; r < rows; ++r) {
    const std::vector<gid>& b = ix->Buckets[r];
    if (sizes) sizes[r] = (uint32_t)b.size();
    for (size_t j = 0; j < b.size(); ++j, ++n)
      if (gids && n < cap) gids[n] = b[j];
  }
  return n;
}

// Index::query_sketch — src/niqki_index.cpp:633-687
size_t ref_query_sketch(void* h, const int32_t* sketch, uint32_t* out_count, uint32_t* out_gid, size_t cap) {
  Index* ix = static_cast<Index*>(h);
  std::vector<int32_t> sk(sketch, sketch + ix->F);
  query_output r = ix->query_sketch(sk);
  for (size_t i = 0; i < r.size() && i < cap; ++i) {
    out_count[i] = r[i].first;
    out_gid[i] = r[i].second;
  }
  return r.size();
}

// nq query_sketch calls under OpenMP; returns wall seconds; total hits in *total_hits
double ref_query_batch(void* h, const int32_t* sketches, size_t nq, uint64_t* total_hits, int nthreads) {
  Index* ix = static_cast<Index*>(h);
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  uint64_t hits = 0;
  auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) reduction(+ : hits)
  for (long q = 0; q < (long)nq; ++q) {
    std::vector<int32_t> sk(sketches + (size_t)q * ix->F, sketches + (size_t)(q + 1) * ix->F);
    hits += ix->query_sketch(sk).size();
  }
  auto t1 = std::chrono::steady_clock::now();
  if (total_hits) *total_hits = hits;
  return std::chrono::duration<double>(t1 - t0).count();
}

// Index::query_matrix — src/niqki_index.cpp:614-628; text goes to the index's output stream
void ref_query_matrix(void* h) { static_cast<Index*>(h)->query_matrix(); }
// Index::output_query — src/niqki_index.cpp:544-566 (pretty != 0 -> text, else binary records)
void ref_output_query(void* h, const uint32_t* counts, const uint32_t* gids, size_t n, const char* name, int pretty) {
  Index* ix = static_cast<Index*>(h);
  query_output q;
  for (size_t i = 0; i < n; ++i) q.push_back({counts[i], gids[i]});
  bool old = ix->pretty_printing;
  ix->pretty_printing = pretty != 0;
  ix->output_query(q, name);
  ix->pretty_printing = old;
}
void ref_flush_output(void* h) { static_cast<Index*>(h)->outfile->flush(); }
// Index::dump_index_disk — src/niqki_index.cpp:42-59
void ref_dump(void* h, const char* path) { static_cast<Index*>(h)->dump_index_disk(path); }

}  // extern "C"
