// pack.cpp — K1, host half: sequence packing to 2 bits per base for the host-buffer entry points.
// A base travels over PCIe as 2 bits instead of 8, which is what bounds nq_sketch_batch end to end.
//
// The reference's per-character rules (/root/reference/src/niqki_index.cpp:114-123, 211-221,
// 255-273) are not a plain 2-bit alphabet, so the wire format is
//   codes : u32 per 16 bases, base i at bits 2*(i%16): nuc2int's forward code {A:0,C:1,G:2,T:3},
//           0 for every other byte;
//   other : one bit per base, set where the byte is not upper-case ACGT — there the reference's
//           complement code (nuc2intrc) is 0 as well, instead of 3 - code.  Kept sparse: blk[b] is
//           the slot of 512-base block b in `pool` (32 x u16), or 0xFFFFFFFF when the block holds
//           no such byte (almost all blocks of an assembled genome);
//   seed  : the first K-1 characters of a record follow str2numstrand (:255-273) instead: both
//           cases accepted, and ANY other byte zeroes the whole seed.  The packer writes those K-1
//           positions with exactly the digits the rolling update would then see — the
//           case-insensitive codes, or K-1 'A's — and clears their `other` bits, so the device
//           needs no per-record special case: every k-mer is a window of the packed stream.
// The stream starts with kPackLead unused bases (the window of a record's first k-mer begins one
// base early).
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "pack.h"

namespace {

inline uint32_t fw_code_of(uint8_t c) { return c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 0u; }
inline bool is_acgt(uint8_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
inline uint32_t seed_code_of(uint8_t c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
  }
}

// 16 bases -> code word + other mask (scalar)
inline void pack16_scalar(const uint8_t* s, uint32_t n, uint32_t& codes, uint16_t& other) {
  uint32_t c = 0, o = 0;
  for (uint32_t i = 0; i < n; ++i) {
    c |= fw_code_of(s[i]) << (2 * i);
    o |= (is_acgt(s[i]) ? 0u : 1u) << i;
  }
  codes = c;
  other = (uint16_t)o;
}

#if defined(__x86_64__)
// 32 bases -> two code words + 32 other bits.  (c >> 1) & 3 maps A,C,G,T to 0,1,3,2; x ^ (x >> 1)
// turns that into 0,1,2,3.  Validity: a 16-entry table of the four letters indexed by the low
// nibble must reproduce the byte.
__attribute__((target("avx2,bmi2"))) inline void pack32_avx2(const uint8_t* s, uint64_t& codes, uint32_t& other) {
  const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s));
  // (entry 0 is 0xFF: a byte with low nibble 0 — NUL included — must never match)
  const __m256i tab = _mm256_setr_epi8(-1, 'A', 0, 'C', 'T', 0, 0, 'G', 0, 0, 0, 0, 0, 0, 0, 0,
                                       -1, 'A', 0, 'C', 'T', 0, 0, 'G', 0, 0, 0, 0, 0, 0, 0, 0);
  const __m256i lo = _mm256_and_si256(v, _mm256_set1_epi8(0x0F));
  const __m256i ok = _mm256_cmpeq_epi8(_mm256_shuffle_epi8(tab, lo), v);  // high bit set bytes shuffle to 0 != v
  const uint32_t valid = (uint32_t)_mm256_movemask_epi8(ok);
  // bit 1 and bit 2 of the byte -> planes (c>>1)&1 and (c>>2)&1
  const uint32_t p0 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 6));  // bit 1 -> bit 7
  const uint32_t p1 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 5));  // bit 2 -> bit 7
  // x = p1:p0 per base in {0,1,3,2}; code = x ^ (x >> 1): bit1 = p1, bit0 = p0 ^ p1
  const uint32_t b1 = p1 & valid, b0 = (p0 ^ p1) & valid;
  codes = _pdep_u64(b0, 0x5555555555555555ull) | _pdep_u64(b1, 0xAAAAAAAAAAAAAAAAull);
  other = ~valid;
}
#endif

bool have_avx2() {
#if defined(__x86_64__)
  static const bool ok = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2");
  return ok;
#else
  return false;
#endif
}

#if defined(__x86_64__)
__attribute__((target("avx2,bmi2"))) void pack_range_avx2(const uint8_t* src, uint64_t w0, uint64_t w1, uint64_t nbytes,
                                                          uint32_t* codes, uint16_t* other) {
  // words w0..w1 of the SOURCE stream (word w = bytes 16w .. 16w+15); full 32-byte pairs with AVX2
  uint64_t w = w0;
  for (; w + 2 <= w1 && (w + 2) * 16 <= nbytes; w += 2) {
    uint64_t c;
    uint32_t o;
    pack32_avx2(src + w * 16, c, o);
    codes[w] = (uint32_t)c;
    codes[w + 1] = (uint32_t)(c >> 32);
    other[w] = (uint16_t)o;
    other[w + 1] = (uint16_t)(o >> 16);
  }
  for (; w < w1; ++w) {
    const uint64_t at = w * 16;
    const uint32_t n = at >= nbytes ? 0u : (uint32_t)(nbytes - at < 16 ? nbytes - at : 16);
    pack16_scalar(src + at, n, codes[w], other[w]);
  }
}
#endif

void pack_range_scalar(const uint8_t* src, uint64_t w0, uint64_t w1, uint64_t nbytes, uint32_t* codes, uint16_t* other) {
  for (uint64_t w = w0; w < w1; ++w) {
    const uint64_t at = w * 16;
    const uint32_t n = at >= nbytes ? 0u : (uint32_t)(nbytes - at < 16 ? nbytes - at : 16);
    pack16_scalar(src + at, n, codes[w], other[w]);
  }
}

}  // namespace

uint64_t nq_pack_words(uint64_t nbytes) { return (kPackLead + nbytes + 15) / 16 + kPackTailWords; }
uint64_t nq_pack_blocks(uint64_t nbytes) { return (nq_pack_words(nbytes) + 31) / 32; }

// codes[words], blk[blocks], pool[(pool_cap) * 32], dense[words] = caller-provided scratch (u16 per word).
// Returns the number of pool slots used, or ~0ull when pool_cap is too small.
uint64_t nq_pack_host(const char* bases, uint64_t nbytes, const uint64_t* rec_offsets, uint64_t n_rec, uint32_t K,
                      uint32_t* codes, uint32_t* blk, uint16_t* pool, uint64_t pool_cap, uint16_t* dense, unsigned threads) {
  const uint8_t* src = reinterpret_cast<const uint8_t*>(bases);
  const uint64_t words = nq_pack_words(nbytes), blocks = nq_pack_blocks(nbytes);
  constexpr uint64_t lead_w = kPackLead / 16;  // a whole block: source word w lands at packed word w + lead_w
  for (uint64_t w = 0; w < lead_w; ++w) { codes[w] = 0; dense[w] = 0; }
  blk[0] = 0xFFFFFFFFu;
  const uint64_t src_words = words - lead_w;
  if (threads == 0) threads = 1;
  const uint64_t chunk = 1 << 16;  // source words per task (1 MiB of sequence, 2048 blocks)
  const uint64_t tasks = (src_words + chunk - 1) / chunk;
  std::atomic<uint64_t> next{0};
  auto work = [&]() {
    for (;;) {
      const uint64_t t = next.fetch_add(1);
      if (t >= tasks) break;
      const uint64_t w0 = t * chunk, w1 = w0 + chunk < src_words ? w0 + chunk : src_words;
#if defined(__x86_64__)
      if (have_avx2()) pack_range_avx2(src, w0, w1, nbytes, codes + lead_w, dense + lead_w);
      else
#endif
        pack_range_scalar(src, w0, w1, nbytes, codes + lead_w, dense + lead_w);
      // which 512-base blocks hold a byte that is not upper-case ACGT
      for (uint64_t b0 = w0; b0 < w1; b0 += 32) {
        const uint64_t e = b0 + 32 < w1 ? b0 + 32 : w1;
        uint16_t any = 0;
        for (uint64_t w = b0; w < e; ++w) any |= dense[lead_w + w];
        blk[(lead_w + b0) / 32] = any ? 1u : 0xFFFFFFFFu;
      }
    }
  };
  const unsigned nt = (unsigned)(tasks < threads ? (tasks ? tasks : 1) : threads);
  if (nt <= 1) {
    work();
  } else {
    std::vector<std::thread> pool_threads;
    pool_threads.reserve(nt - 1);
    for (unsigned i = 1; i < nt; ++i) pool_threads.emplace_back(work);
    work();
    for (auto& th : pool_threads) th.join();
  }
  // seeds: the first K-1 characters of every record (:255-273, :340-341)
  for (uint64_t r = 0; r < n_rec; ++r) {
    const uint64_t e0 = rec_offsets[r] - rec_offsets[0], len = rec_offsets[r + 1] - rec_offsets[r];
    const uint32_t ns = (uint32_t)(len < K - 1 ? len : K - 1);
    bool ok = true;
    for (uint32_t j = 0; j < ns; ++j) ok = ok && seed_code_of(src[e0 + j]) < 4;
    for (uint32_t j = 0; j < ns; ++j) {
      const uint64_t p = kPackLead + e0 + j;
      const uint32_t code = ok ? seed_code_of(src[e0 + j]) : 0u, sh = 2 * (uint32_t)(p & 15);
      codes[p >> 4] = (codes[p >> 4] & ~(3u << sh)) | (code << sh);
      dense[p >> 4] &= (uint16_t)~(1u << (p & 15));
    }
  }
  // sparse `other` masks: flagged blocks get a pool slot (a block whose only such bytes sat in a seed
  // keeps a slot of zeros)
  uint64_t used = 0;
  for (uint64_t b = 0; b < blocks; ++b) {
    if (blk[b] == 0xFFFFFFFFu) continue;
    if (used >= pool_cap) return ~0ull;
    const uint64_t w0 = b * 32, w1 = w0 + 32 < words ? w0 + 32 : words;
    uint16_t* slot = pool + used * 32;
    for (uint64_t w = w0; w < w0 + 32; ++w) slot[w - w0] = w < w1 ? dense[w] : 0;
    blk[b] = (uint32_t)used++;
  }
  return used;
}
