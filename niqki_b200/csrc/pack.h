// pack.h — host-side 2-bit packing of sequence batches (pack.cpp) and the device kernel that reads
// it (sketch_packed.cu).  Internal to libniqki_b200.so.
#pragma once
#include <cstdint>

constexpr uint64_t kPackLead = 512;      // unused bases in front of the stream (one whole mask block)
constexpr uint64_t kPackTailWords = 4;   // zero words behind the last base (k-mer windows read ahead)

uint64_t nq_pack_words(uint64_t nbytes);   // u32 code words of a batch of nbytes characters
uint64_t nq_pack_blocks(uint64_t nbytes);  // 512-base mask blocks
// codes[words], blk[blocks], pool[pool_cap * 32].  Returns the pool slots used, or ~0ull when pool_cap is
// too small.
uint64_t nq_pack_host(const char* bases, uint64_t nbytes, const uint64_t* rec_offsets, uint64_t n_rec, uint32_t K,
                      uint32_t* codes, uint32_t* blk, uint16_t* pool, uint64_t pool_cap, unsigned threads);
