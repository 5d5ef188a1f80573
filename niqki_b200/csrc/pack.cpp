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
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
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

// Helper threads that outlive a call: a batch is packed in ~2 ms, starting 15 threads for it costs as much.
class Workers {
 public:
  static Workers& get() {
    static Workers w;
    return w;
  }
  // runs fn() on `extra` helper threads and on the caller; returns when all are done
  void run(unsigned extra, const std::function<void()>& fn) {
    std::lock_guard<std::mutex> one_call(call_);
    {
      std::unique_lock<std::mutex> lk(m_);
      while (threads_.size() < extra) threads_.emplace_back([this, i = threads_.size()] { loop(i); });
      fn_ = &fn;
      want_ = extra;
      running_ = extra;
      ++epoch_;
    }
    cv_.notify_all();
    fn();
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [this] { return running_ == 0; });
    fn_ = nullptr;
  }
  ~Workers() {
    {
      std::unique_lock<std::mutex> lk(m_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }

 private:
  void loop(size_t me) {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void()>* fn = nullptr;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return stop_ || (epoch_ != seen && me < want_); });
        if (stop_) return;
        seen = epoch_;
        fn = fn_;
      }
      (*fn)();
      std::unique_lock<std::mutex> lk(m_);
      if (--running_ == 0) done_.notify_all();
    }
  }
  std::mutex m_, call_;
  std::condition_variable cv_, done_;
  std::vector<std::thread> threads_;
  const std::function<void()>* fn_ = nullptr;
  uint64_t epoch_ = 0;
  unsigned want_ = 0, running_ = 0;
  bool stop_ = false;
};

int simd_level() {  // 2: AVX-512BW + BMI2, 1: AVX2 + BMI2, 0: scalar
#if defined(__x86_64__)
  static const int lvl = (__builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("bmi2")) ? 2
                         : (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2")) ? 1 : 0;
  return lvl;
#else
  return 0;
#endif
}

#if defined(__x86_64__)
// 64 bases -> four code words + 64 other bits
__attribute__((target("avx512f,avx512bw,bmi2"))) inline void pack64_avx512(const uint8_t* s, uint64_t& c_lo, uint64_t& c_hi, uint64_t& other) {
  const __m512i v = _mm512_loadu_si512(reinterpret_cast<const void*>(s));
  const __mmask64 valid = _mm512_cmpeq_epi8_mask(v, _mm512_set1_epi8('A')) | _mm512_cmpeq_epi8_mask(v, _mm512_set1_epi8('C')) |
                          _mm512_cmpeq_epi8_mask(v, _mm512_set1_epi8('G')) | _mm512_cmpeq_epi8_mask(v, _mm512_set1_epi8('T'));
  const __mmask64 p0 = _mm512_test_epi8_mask(v, _mm512_set1_epi8(0x02)), p1 = _mm512_test_epi8_mask(v, _mm512_set1_epi8(0x04));
  const uint64_t b1 = (uint64_t)p1 & (uint64_t)valid, b0 = ((uint64_t)p0 ^ (uint64_t)p1) & (uint64_t)valid;
  c_lo = _pdep_u64(b0 & 0xFFFFFFFFull, 0x5555555555555555ull) | _pdep_u64(b1 & 0xFFFFFFFFull, 0xAAAAAAAAAAAAAAAAull);
  c_hi = _pdep_u64(b0 >> 32, 0x5555555555555555ull) | _pdep_u64(b1 >> 32, 0xAAAAAAAAAAAAAAAAull);
  other = ~(uint64_t)valid;
}
#endif

// One 512-base block (32 source words from byte offset `at`): codes out, the block's `other` masks into oth[32].
// Returns the OR of the masks.  `nbytes` bounds the reads (bytes past it pack as code 0, other 0).
template <int LEVEL>
#if defined(__x86_64__)
__attribute__((target("avx512f,avx512bw,avx2,bmi2")))
#endif
inline uint32_t pack_block(const uint8_t* src, uint64_t at, uint64_t nbytes, uint32_t* codes, uint16_t* oth) {
  uint32_t any = 0;
  if (at + 512 <= nbytes) {
#if defined(__x86_64__)
    if (LEVEL == 2) {
      alignas(64) uint64_t cw[16];  // the block's 128 bytes of codes, written with streaming stores (no read-for-ownership)
      for (int k = 0; k < 8; ++k) {
        uint64_t o;
        pack64_avx512(src + at + 64 * k, cw[2 * k], cw[2 * k + 1], o);
        oth[4 * k] = (uint16_t)o; oth[4 * k + 1] = (uint16_t)(o >> 16); oth[4 * k + 2] = (uint16_t)(o >> 32); oth[4 * k + 3] = (uint16_t)(o >> 48);
        any |= (o != 0);
      }
      if ((reinterpret_cast<uintptr_t>(codes) & 63) == 0) {
        _mm512_stream_si512(reinterpret_cast<__m512i*>(codes), _mm512_load_si512(cw));
        _mm512_stream_si512(reinterpret_cast<__m512i*>(codes + 16), _mm512_load_si512(cw + 8));
      } else {
        memcpy(codes, cw, 128);
      }
      return any;
    }
    if (LEVEL == 1) {
      for (int k = 0; k < 16; ++k) {
        uint64_t c;
        uint32_t o;
        pack32_avx2(src + at + 32 * k, c, o);
        codes[2 * k] = (uint32_t)c; codes[2 * k + 1] = (uint32_t)(c >> 32);
        oth[2 * k] = (uint16_t)o; oth[2 * k + 1] = (uint16_t)(o >> 16);
        any |= o;
      }
      return any;
    }
#endif
  }
  for (int w = 0; w < 32; ++w) {
    const uint64_t a = at + 16 * (uint64_t)w;
    const uint32_t n = a >= nbytes ? 0u : (uint32_t)(nbytes - a < 16 ? nbytes - a : 16);
    pack16_scalar(src + a, n, codes[w], oth[w]);
    any |= oth[w];
  }
  return any;
}

}  // namespace

uint64_t nq_pack_words(uint64_t nbytes) { return ((kPackLead + nbytes + 15) / 16 + kPackTailWords + 31) / 32 * 32; }  // whole blocks
uint64_t nq_pack_blocks(uint64_t nbytes) { return nq_pack_words(nbytes) / 32; }

// codes[words], blk[blocks], pool[pool_cap * 32].  Returns the pool slots used, or ~0ull when pool_cap is
// too small.  Blocks are packed by `threads` workers (tasks of 1 MiB of sequence); a block with a byte
// that is not upper-case ACGT leaves its 32 mask words in the task's own list, and the lists are laid
// into the pool in task order afterwards, so the result does not depend on the thread count.
uint64_t nq_pack_host(const char* bases, uint64_t nbytes, const uint64_t* rec_offsets, uint64_t n_rec, uint32_t K,
                      uint32_t* codes, uint32_t* blk, uint16_t* pool, uint64_t pool_cap, unsigned threads) {
  const uint8_t* src = reinterpret_cast<const uint8_t*>(bases);
  const uint64_t blocks = nq_pack_blocks(nbytes);
  constexpr uint64_t lead_b = kPackLead / 512;  // source block b lands at packed block b + lead_b
  for (uint64_t w = 0; w < lead_b * 32; ++w) codes[w] = 0;
  for (uint64_t b = 0; b < lead_b; ++b) blk[b] = 0xFFFFFFFFu;
  const uint64_t src_blocks = blocks - lead_b;
  if (threads == 0) threads = 1;
  const uint64_t chunk = 2048;  // source blocks per task (1 MiB of sequence)
  const uint64_t tasks = (src_blocks + chunk - 1) / chunk;
  struct TaskOut { std::vector<uint16_t> masks; std::vector<uint64_t> which; };
  std::vector<TaskOut> outs(tasks);
  std::atomic<uint64_t> next{0};
  const int level = simd_level();
  auto work = [&]() {
    for (;;) {
      const uint64_t t = next.fetch_add(1);
      if (t >= tasks) break;
      const uint64_t b0 = t * chunk, b1 = b0 + chunk < src_blocks ? b0 + chunk : src_blocks;
      TaskOut& out = outs[t];
      uint16_t oth[32];
      for (uint64_t b = b0; b < b1; ++b) {
        uint32_t* c = codes + (lead_b + b) * 32;
        const uint32_t any = level == 2 ? pack_block<2>(src, b * 512, nbytes, c, oth)
                             : level == 1 ? pack_block<1>(src, b * 512, nbytes, c, oth) : pack_block<0>(src, b * 512, nbytes, c, oth);
        if (any) {
          out.which.push_back(lead_b + b);
          out.masks.insert(out.masks.end(), oth, oth + 32);
        } else {
          blk[lead_b + b] = 0xFFFFFFFFu;
        }
      }
    }
#if defined(__x86_64__)
    _mm_sfence();  // streaming stores of this worker are visible before it reports done
#endif
  };
  const unsigned nt = (unsigned)(tasks < threads ? (tasks ? tasks : 1) : threads);
  if (nt <= 1) work();
  else Workers::get().run(nt - 1, work);
#if defined(__x86_64__)
  _mm_sfence();  // (each worker's streaming stores are ordered by its own exit from run(); this covers the caller's)
#endif
  // sparse `other` masks: the flagged blocks, in stream order
  uint64_t used = 0;
  for (uint64_t t = 0; t < tasks; ++t) {
    const TaskOut& out = outs[t];
    if (used + out.which.size() > pool_cap) return ~0ull;
    for (size_t i = 0; i < out.which.size(); ++i) {
      memcpy(pool + (used + i) * 32, out.masks.data() + i * 32, 64);
      blk[out.which[i]] = (uint32_t)(used + i);
    }
    used += out.which.size();
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
      const uint32_t slot = blk[p >> 9];
      if (slot != 0xFFFFFFFFu) pool[(uint64_t)slot * 32 + ((p >> 4) & 31)] &= (uint16_t)~(1u << (p & 15));
    }
  }
  return used;
}
