/*
 * niqki_oracle.c — TEST INFRASTRUCTURE ONLY (see niqki_oracle.h for the rules and the parity
 * status: PINNED against SURVEY.md App. C, tests/golden/ and oracle/_ref).
 *
 * CPU restatement of the NIQKI hot path written from SURVEY.md Appendix A.  Citations are
 * file:line relative to /root/reference/.
 */
#include "niqki_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ parameters */

/* src/niqki_index.cpp:13-29: F=1<<lF, M=W-H, min_score=min_fract*F (double -> u32 truncation),
 * fingerprint_range=1<<W, mask_M=(1<<M)-1, maximal_remainder=(1<<H)-1. */
void nqo_params_init(nqo_params* p, uint32_t K, uint32_t S, uint32_t W, uint32_t H, double J) {
  p->K = K;
  p->S = S;
  p->W = W;
  p->H = H;
  p->F = 1u << S;
  p->M = W - H;
  p->min_score = (uint32_t)(J * (double)p->F);
  p->range = (int32_t)(1u << W);
  p->mask_M = (1u << p->M) - 1u;
  p->maxrem = (1u << H) - 1u;
}

/* src/niqki_index.cpp:142-164 — closed-form width of the usable Jaccard interval for a given H.
 * Kept in double with the same operation order so the argmax is the same. */
static double score_H(const nqo_params* p, double x, int try_h) {
  const double epsilon = 0.02;
  const double W = (double)p->W;
  const double try_m = W - try_h;
  const double two_h = pow(2, try_h);
  double k[2];
  for (int side = 0; side < 2; ++side) {
    const double base = side == 0 ? 1 - epsilon : epsilon;
    const double u = ((double)1 - pow(base, 1 / x)) * pow(2, 64);
    const double i = log2(u) + two_h - 64;
    const double j = u * pow(2, try_m - 64 - i + two_h);
    if (u < pow(2, 64 - two_h + 1)) {
      k[side] = u * pow(2, two_h - 64 - (W - try_h) - 1);
    } else {
      k[side] = i * pow(2, try_m) + j;
    }
  }
  return k[1] - k[0];
}

/* src/niqki_index.cpp:126-138 — picks H in [2,6], updates H and M only; mask_M and maxrem keep
 * their construction-time values. */
void nqo_select_best_H(nqo_params* p, double genome_size) {
  const double x = genome_size / (double)p->F;
  double best = 0;
  for (uint32_t h = 2; h < 7; ++h) {
    const double s = score_H(p, x, (int)h);
    if (s > best) {
      best = s;
      p->H = h;
    }
  }
  p->M = p->W - p->H;
}

/* ------------------------------------------------------------------ hashes */

static inline uint64_t fold_mul2(uint64_t x, uint64_t c) {
  x = ((x >> 32) ^ x) * c;
  x = ((x >> 32) ^ x) * c;
  return (x >> 32) ^ x;
}
/* src/niqki_index.cpp:291-296 */
uint64_t nqo_revhash64(uint64_t x) { return fold_mul2(x, 0xD6E8FEB86659FD93ull); }
/* src/niqki_index.cpp:300-305 */
uint64_t nqo_unrevhash64(uint64_t x) { return fold_mul2(x, 0xCFEE444D8B59A89Bull); }
/* src/niqki_index.cpp:308-310 */
uint64_t nqo_hash_family(uint64_t x, uint32_t factor) {
  return nqo_unrevhash64(x) + (uint64_t)factor * nqo_revhash64(x);
}

/* src/niqki_index.cpp:277-287: low M bits of the hash, plus max(0, maxrem - clz64) << M.  The
 * reference takes clz from x86 `bsr` (:199-206), undefined for 0; observed result is fp=0, which
 * clz64(0)=64 reproduces. */
int32_t nqo_get_fingerprint(const nqo_params* p, uint64_t hashed) {
  const int lz = hashed ? __builtin_clzll(hashed) : 64;
  int rem = (int)p->maxrem - lz;
  if (rem < 0) rem = 0;
  return (int32_t)((uint32_t)(hashed & p->mask_M) + ((uint32_t)rem << p->M));
}

/* ------------------------------------------------------------------ k-mer encoding */

/* src/niqki_index.cpp:114-123: forward code, exact upper-case match only */
static inline uint64_t fw_code(char c) { return c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 0; }
/* src/niqki_index.cpp:211-221: complement code, exact upper-case match only */
static inline uint64_t rv_code(char c) { return c == 'A' ? 3 : c == 'C' ? 2 : c == 'G' ? 1 : 0; }

/* src/niqki_index.cpp:255-273: big-endian base-4 value, case-insensitive; 0 on any other byte */
uint64_t nqo_str2numstrand(const char* s, size_t n) {
  uint64_t v = 0;
  for (size_t i = 0; i < n; ++i) {
    uint64_t d;
    switch (s[i]) {
      case 'A': case 'a': d = 0; break;
      case 'C': case 'c': d = 1; break;
      case 'G': case 'g': d = 2; break;
      case 'T': case 't': d = 3; break;
      default: return 0; /* one foreign byte zeroes the whole seed */
    }
    v = (v << 2) + d;
  }
  return v;
}

/* src/niqki_index.cpp:240-250: reverse the K base-4 digits and complement each */
uint64_t nqo_rcb(const nqo_params* p, uint64_t x) {
  uint64_t r = 0;
  for (uint32_t i = 0; i < p->K; ++i) {
    r = (r << 2) | (3 - (x & 3));
    x >>= 2;
  }
  return r;
}

/* ------------------------------------------------------------------ sketch */

/* src/niqki_index.cpp:339-356.  Seed from the first K-1 characters, then one canonical k-mer per
 * position i with i+K < len (the k-mer starting at len-K is never hashed). */
uint32_t nqo_sketch_scan(const nqo_params* p, const char* seq, size_t len, int32_t* sketch) {
  const uint32_t K = p->K;
  uint32_t filled = 0;
  if (len <= K) return 0;
  const uint64_t kmask = (K < 32) ? ((1ull << (2 * K)) - 1) : ~0ull; /* % offsetUpdatekmer, :228 */
  uint64_t f = nqo_str2numstrand(seq, K - 1);
  uint64_t r = nqo_rcb(p, f);
  for (size_t i = 0; i + K < len; ++i) {
    const char c = seq[i + K - 1];
    f = ((f << 2) + fw_code(c)) & kmask;                 /* :225-229 */
    r = (r >> 2) + (rv_code(c) << (2 * K - 2));          /* :233-236 */
    const uint64_t canon = f < r ? f : r;                /* :345 */
    const uint64_t bucket = nqo_unrevhash64(canon) >> (64 - p->S); /* :347 */
    const int32_t fp = nqo_get_fingerprint(p, nqo_revhash64(canon));
    if (sketch[bucket] == -1) {                          /* :350-355 */
      sketch[bucket] = fp;
      ++filled;
    } else if (sketch[bucket] > fp) {
      sketch[bucket] = fp;
    }
  }
  return filled;
}

/* src/niqki_index.cpp:313-331.  Sequential, order dependent: cells filled earlier in a pass are
 * sources later in the same pass. */
long nqo_sketch_densification(const nqo_params* p, int32_t* sketch, uint32_t empty_cell,
                              long max_passes) {
  const uint32_t F = p->F;
  uint32_t step = 0;
  long passes = 0;
  while (empty_cell != 0) {
    if (max_passes && passes >= max_passes) return -1;
    for (uint32_t i = 0; i < F; ++i) {
      if (sketch[i] == -1) continue;
      const uint64_t t = nqo_hash_family((uint64_t)sketch[i], step) % F;
      if (sketch[t] == -1) {
        sketch[t] = sketch[i];
        if (--empty_cell == 0) return passes + 1;
      }
    }
    ++step;
    ++passes;
  }
  return passes;
}

/* src/niqki_index.cpp:335-358 */
long nqo_compute_sketch(const nqo_params* p, const char* seq, size_t len, int32_t* sketch,
                        long max_passes) {
  if (len <= p->K) return 0; /* callers gate on size()>K (:395,:423,:450,:512) */
  const uint32_t filled = nqo_sketch_scan(p, seq, len, sketch);
  return nqo_sketch_densification(p, sketch, p->F - filled, max_passes);
}

void nqo_sketch_batch(const nqo_params* p, const char* bases, const uint64_t* offsets, size_t n,
                      int32_t* out, int nthreads) {
  (void)nthreads;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
#endif
  for (long e = 0; e < (long)n; ++e) {
    int32_t* sk = out + (size_t)e * p->F;
    for (uint32_t i = 0; i < p->F; ++i) sk[i] = -1;
    nqo_compute_sketch(p, bases + offsets[e], (size_t)(offsets[e + 1] - offsets[e]), sk, 0);
  }
}

/* ------------------------------------------------------------------ inverted index */

struct nqo_index {
  nqo_params p;
  uint32_t n_genomes;
  /* append log, in insert order */
  uint32_t* keys;
  uint32_t* vals;
  size_t n, cap;
  /* CSR, valid when frozen */
  int frozen;
  uint64_t* row_ptr;
  uint32_t* gids;
};

nqo_index* nqo_index_new(const nqo_params* p) {
  nqo_index* ix = (nqo_index*)calloc(1, sizeof(*ix));
  ix->p = *p;
  return ix;
}

void nqo_index_free(nqo_index* ix) {
  if (!ix) return;
  free(ix->keys);
  free(ix->vals);
  free(ix->row_ptr);
  free(ix->gids);
  free(ix);
}

/* src/niqki_index.cpp:362-370: only cells with 0 <= fp < range are posted, list id = fp+cell*range */
void nqo_index_insert(nqo_index* ix, const int32_t* sketch, uint32_t gid) {
  const uint32_t F = ix->p.F;
  if (ix->n + F > ix->cap) {
    size_t nc = ix->cap ? ix->cap * 2 : (size_t)F * 16;
    while (nc < ix->n + F) nc *= 2;
    ix->keys = (uint32_t*)realloc(ix->keys, nc * sizeof(uint32_t));
    ix->vals = (uint32_t*)realloc(ix->vals, nc * sizeof(uint32_t));
    ix->cap = nc;
  }
  for (uint32_t i = 0; i < F; ++i) {
    if (sketch[i] < ix->p.range && sketch[i] >= 0) {
      ix->keys[ix->n] = (uint32_t)sketch[i] + i * (uint32_t)ix->p.range;
      ix->vals[ix->n] = gid;
      ++ix->n;
    }
  }
  if (gid + 1 > ix->n_genomes) ix->n_genomes = gid + 1;
  ix->frozen = 0;
}

void nqo_index_finalize(nqo_index* ix) {
  if (ix->frozen) return;
  const size_t nrows = (size_t)ix->p.range * ix->p.F;
  free(ix->row_ptr);
  free(ix->gids);
  ix->row_ptr = (uint64_t*)calloc(nrows + 1, sizeof(uint64_t));
  ix->gids = (uint32_t*)malloc((ix->n ? ix->n : 1) * sizeof(uint32_t));
  for (size_t i = 0; i < ix->n; ++i) ix->row_ptr[ix->keys[i] + 1]++;
  for (size_t r = 0; r < nrows; ++r) ix->row_ptr[r + 1] += ix->row_ptr[r];
  uint64_t* cur = (uint64_t*)malloc(nrows * sizeof(uint64_t));
  memcpy(cur, ix->row_ptr, nrows * sizeof(uint64_t));
  for (size_t i = 0; i < ix->n; ++i) ix->gids[cur[ix->keys[i]]++] = ix->vals[i];
  free(cur);
  ix->frozen = 1;
}

uint32_t nqo_index_num_genomes(const nqo_index* ix) { return ix->n_genomes; }
uint64_t nqo_index_num_postings(nqo_index* ix) { return ix->n; }
const uint64_t* nqo_index_row_ptr(nqo_index* ix) {
  nqo_index_finalize(ix);
  return ix->row_ptr;
}
const uint32_t* nqo_index_gids(nqo_index* ix) {
  nqo_index_finalize(ix);
  return ix->gids;
}

/* ------------------------------------------------------------------ query */

static int cmp_hit_desc(const void* a, const void* b) {
  /* std::greater<pair<u32 count, u32 gid>> — src/niqki_index.cpp:685 */
  const uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return x < y ? 1 : x > y ? -1 : 0;
}

/* src/niqki_index.cpp:633-687.  Counter width u8 (S<=7), u16 (S<=15), u32 otherwise; modelled as
 * u32 counts reduced mod the width at the end (identical for ++ only). */
static size_t query_one(nqo_index* ix, const int32_t* sketch, uint32_t* counts, uint64_t* packed) {
  const nqo_params* p = &ix->p;
  const uint32_t N = ix->n_genomes;
  memset(counts, 0, (size_t)N * sizeof(uint32_t));
  for (uint32_t i = 0; i < p->F; ++i) {
    if (sketch[i] < p->range && sketch[i] >= 0) {
      const size_t key = (size_t)sketch[i] + (size_t)i * (size_t)p->range;
      for (uint64_t j = ix->row_ptr[key]; j < ix->row_ptr[key + 1]; ++j) counts[ix->gids[j]]++;
    }
  }
  const uint32_t wrap = p->S <= 7 ? 0xFFu : p->S <= 15 ? 0xFFFFu : 0xFFFFFFFFu;
  size_t nh = 0;
  for (uint32_t g = 0; g < N; ++g) {
    const uint32_t c = counts[g] & wrap;
    if (c >= p->min_score) packed[nh++] = ((uint64_t)c << 32) | g;
  }
  qsort(packed, nh, sizeof(uint64_t), cmp_hit_desc);
  return nh;
}

size_t nqo_query_sketch(nqo_index* ix, const int32_t* sketch, uint32_t* out_count,
                        uint32_t* out_gid, size_t cap) {
  nqo_index_finalize(ix);
  const uint32_t N = ix->n_genomes;
  uint32_t* counts = (uint32_t*)malloc((N ? N : 1) * sizeof(uint32_t));
  uint64_t* packed = (uint64_t*)malloc((N ? N : 1) * sizeof(uint64_t));
  const size_t nh = query_one(ix, sketch, counts, packed);
  for (size_t i = 0; i < nh && i < cap; ++i) {
    out_count[i] = (uint32_t)(packed[i] >> 32);
    out_gid[i] = (uint32_t)packed[i];
  }
  free(counts);
  free(packed);
  return nh;
}

size_t nqo_query_batch(nqo_index* ix, const int32_t* sketches, size_t nq, uint64_t* hit_ptr,
                       uint32_t* out_count, uint32_t* out_gid, size_t cap, int nthreads) {
  nqo_index_finalize(ix);
  const uint32_t N = ix->n_genomes;
  const uint32_t F = ix->p.F;
  (void)nthreads;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#endif
  uint64_t** per_q = (uint64_t**)calloc(nq ? nq : 1, sizeof(uint64_t*));
  size_t* per_n = (size_t*)calloc(nq ? nq : 1, sizeof(size_t));
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
  {
    uint32_t* counts = (uint32_t*)malloc((N ? N : 1) * sizeof(uint32_t));
    uint64_t* packed = (uint64_t*)malloc((N ? N : 1) * sizeof(uint64_t));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
    for (long q = 0; q < (long)nq; ++q) {
      const size_t nh = query_one(ix, sketches + (size_t)q * F, counts, packed);
      per_n[q] = nh;
      per_q[q] = (uint64_t*)malloc((nh ? nh : 1) * sizeof(uint64_t));
      memcpy(per_q[q], packed, nh * sizeof(uint64_t));
    }
    free(counts);
    free(packed);
  }
  size_t total = 0;
  for (size_t q = 0; q < nq; ++q) {
    if (hit_ptr) hit_ptr[q] = total;
    for (size_t i = 0; i < per_n[q]; ++i, ++total) {
      if (out_count && total < cap) {
        out_count[total] = (uint32_t)(per_q[q][i] >> 32);
        out_gid[total] = (uint32_t)per_q[q][i];
      }
    }
    free(per_q[q]);
  }
  if (hit_ptr) hit_ptr[nq] = total;
  free(per_q);
  free(per_n);
  return total;
}

/* src/niqki_index.cpp:570-598.  For every list: T = members inside [begin,end); every member a
 * and every t in T bump counts[a*batch + t] (uint16_t, wraps).  Returned transposed into the row
 * order the reference prints (:600-607): out[(q-begin)*N + j]. */
void nqo_matrix_counts(nqo_index* ix, uint32_t begin, uint32_t end, uint16_t* out, int nthreads) {
  nqo_index_finalize(ix);
  const uint32_t N = ix->n_genomes;
  const size_t batch = end - begin;
  const size_t nrows = (size_t)ix->p.range * ix->p.F;
  uint16_t* counts = (uint16_t*)calloc((batch * N) != 0 ? batch * N : 1, sizeof(uint16_t));
  (void)nthreads;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
  for (long long r = 0; r < (long long)nrows; ++r) {
    const uint64_t b = ix->row_ptr[r], e = ix->row_ptr[r + 1];
    for (uint64_t k = b; k < e; ++k) {
      const uint32_t t = ix->gids[k];
      if (t < begin || t >= end) continue;
      for (uint64_t j = b; j < e; ++j) {
        uint16_t* c = &counts[(size_t)ix->gids[j] * batch + (t - begin)];
#ifdef _OPENMP
#pragma omp atomic
#endif
        (*c)++;
      }
    }
  }
  for (size_t q = 0; q < batch; ++q)
    for (uint32_t j = 0; j < N; ++j) out[q * N + j] = counts[(size_t)j * batch + q];
  free(counts);
}

/* ------------------------------------------------------------------ synthetic inputs (§8d) */

uint64_t nqo_mix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

static inline unsigned synth_base_code(uint64_t seed, uint64_t g, uint64_t i) {
  return (unsigned)(nqo_mix(seed + g * 0xD1B54A32D192ED03ull + i) >> 62);
}

void nqo_synth_genome(uint64_t seed, uint64_t g, uint64_t len, char* out) {
  for (uint64_t i = 0; i < len; ++i) out[i] = "ACGT"[synth_base_code(seed, g, i)];
}

/* d*2^64 as an integer threshold: floor(d * 2^64) computed in long double, d in [0,1) */
static inline uint64_t rate_threshold(double d) {
  if (d <= 0) return 0;
  if (d >= 1) return ~0ull;
  return (uint64_t)ldexpl((long double)d, 64);
}

void nqo_synth_mutant(uint64_t seed, uint64_t g, uint64_t q, double d, uint64_t len, char* out) {
  const uint64_t thr = rate_threshold(d);
  for (uint64_t i = 0; i < len; ++i) {
    unsigned b = synth_base_code(seed, g, i);
    const uint64_t r = nqo_mix((seed ^ 0xA5A5A5A5ull) + q * 0x9E3779B97F4A7C15ull + 2 * i + 1);
    if (r < thr) b = (b + 1 + (unsigned)(nqo_mix(r) % 3)) & 3;
    out[i] = "ACGT"[b];
  }
}

void nqo_synth_read(uint64_t seed, uint64_t r, uint64_t genome_len, uint32_t read_len, char* out) {
  const uint64_t g = r % 64;
  const uint64_t start = nqo_mix(r) % (genome_len - read_len);
  for (uint32_t i = 0; i < read_len; ++i) out[i] = "ACGT"[synth_base_code(seed, g, start + i)];
}
