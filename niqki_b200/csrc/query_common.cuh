// query_common.cuh — pieces shared by the query kernels (query.cu: CSR gather forms; slab.cu: granule form).
#pragma once
#include "device_common.cuh"
#include "internal.h"

namespace nq {

struct QueryArgs {
  const int32_t* qsk;  // [nq][F]
  const void* dir;     // [F][row_stride] packed {begin,end} (u16 pair in a u32, or uint2)
  const void* gids;    // [F][gid_stride] local ids (u16 or u32)
  uint32_t F, range, n, row_stride, gid_stride, gid_base, min_score, wrap_mask;
  uint64_t* pool;  // count<<32 | gid
  uint64_t pool_cap;
  unsigned long long* cursor;  // [0] pool cursor, [1] posting entries gathered (statistics)
  uint64_t* hit_begin;  // [nq]
  uint32_t* hit_n;      // [nq]
  uint32_t* gcounts;    // global counters [gridDim.x][n] (GLOBAL mode only)
  uint32_t* slice_hits;           // GLOBAL mode finish scratch: [queries per launch][slices]
  unsigned long long* slice_base;
  uint32_t slices;
  uint32_t parts;       // GLOBAL mode, segment-table form: gridDim.y CTAs share one query (counters zeroed before, finished after)
  uint32_t prefetch;    // cooperative L2 prefetch of upcoming cells (only when a few chunks of cells fit in L2)
  const uint4* dir3;       // split16 side arrays of the index (internal.h), or null
  const uint16_t* gids16;
  uint64_t nq_total;    // queries of the whole call (CTA size of the small-shard form)
  uint64_t q_end;       // one past the last query of the launch (kDual16: a CTA's second query may not exist)
  const uint4* meta;       // slab form (slab.cu): [F][mgroups] {b0, b1, b2, base}
  const uint2* slab;       //   granules in 8-byte units
  const uint32_t* cell_gran;  // [F+1] first granule of each cell (prefetch windows)
  uint32_t mgroups;        //   range / 32
  uint32_t* desc;          //   scratch [queries per launch][F]: resolved probes (slab.cu step 1)
  uint32_t* pf_claim;      //   scratch [F / 256 + 16]: prefetch slices handed out per chunk of cells
  uint32_t exp;            // measurement builds only (NQ_TUNING): experiment bits
  uint32_t* dense;      // when set: row q of [nq][n] takes every genome's count instead of the thresholded hit list (--matrix)
};

enum CountMode { kPack16 = 0, kSmem32 = 1, kGlobal32 = 2, kDual16 = 3 };

// clamped shift: PTX shl.b32 yields 0 for shift amounts >= 32 (C++ leaves that undefined)
__device__ __forceinline__ uint32_t shl_clamp(uint32_t x, uint32_t s) {
  uint32_t r;
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(s));
  return r;
}
// `if (s < total) red.shared.add(addr, v)` as ONE predicated instruction (no branch, no reconvergence)
__device__ __forceinline__ void red_shared_add_if_lt(uint32_t saddr, uint32_t v, uint32_t s, uint32_t total) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.lt.u32 q, %2, %3;\n\t@q red.shared.add.u32 [%0], %1;\n\t}"
               :: "r"(saddr), "r"(v), "r"(s), "r"(total) : "memory");
}
// Pull a contiguous region into L2 through the bulk-copy engine (no LSU/L1TEX work, no registers).
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
}

// ---- threshold (:661-665) + compaction, shared by both gather forms: count the hits, reserve a pool
// segment with one atomicAdd, then write (count, gid) in gid order
// SUM: the posting entries gathered are not handed in but read off the counters (every gathered
// posting incremented exactly one genome's counter; padding ids land behind the n counters)
template <int MODE, int NT, bool SUM = false>
__device__ __forceinline__ void query_finish(const QueryArgs& a, uint64_t q, const uint32_t* cnt, uint32_t gathered,
                                             uint32_t shift) {
  __shared__ uint32_t s_warp[NT / 32];
  __shared__ unsigned long long s_base;
  __shared__ uint32_t s_total;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr unsigned kFull = 0xFFFFFFFFu;
  auto count_raw = [&](uint32_t g) -> uint32_t {
    uint32_t c;
    if (MODE == kPack16) c = (cnt[g >> 1] >> ((g & 1) * 16)) & 0xFFFFu;
    else if (MODE == kDual16) c = (cnt[g] >> shift) & 0xFFFFu;
    else if (MODE == kGlobal32) c = __ldcg(&cnt[g]);
    else c = cnt[g];
    return c;
  };
  auto count_of = [&](uint32_t g) -> uint32_t { return count_raw(g) & a.wrap_mask; };

  if (a.dense) {  // all-vs-all rows (:570-598): the whole counter array, coalesced
    uint32_t* row = a.dense + q * a.n;
    for (uint32_t g = tid; g < a.n; g += NT) row[g] = count_of(g);
    return;
  }
  uint32_t mine = 0;
  for (uint32_t g = tid; g < a.n; g += NT) {
    const uint32_t c = count_raw(g);
    if (SUM) gathered += c;
    mine += (c & a.wrap_mask) >= a.min_score;
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    mine += __shfl_xor_sync(kFull, mine, d);
    gathered += __shfl_xor_sync(kFull, gathered, d);
  }
  if (lane == 0) {
    s_warp[warp] = mine;
    if (gathered) atomicAdd(a.cursor + 1, (unsigned long long)gathered);
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t t = 0;
    for (int w = 0; w < NT / 32; ++w) t += s_warp[w];
    s_total = t;
    s_base = atomicAdd(a.cursor, (unsigned long long)t);
    a.hit_begin[q] = s_base;
    a.hit_n[q] = t;
  }
  __syncthreads();
  const uint32_t total = s_total;
  const unsigned long long base = s_base;
  if (total == 0 || base + total > a.pool_cap) return;  // overflow: host re-runs with a larger pool

  uint32_t done = 0;
  for (uint32_t g0 = 0; g0 < a.n; g0 += NT) {
    const uint32_t g = g0 + tid;
    uint32_t c = 0;
    bool hit = false;
    if (g < a.n) {
      c = count_of(g);
      hit = c >= a.min_score;
    }
    if (__syncthreads_count(hit) == 0) continue;  // hits are sparse: most chunks of NT genomes hold none
    const unsigned bal = __ballot_sync(kFull, hit);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    uint32_t before = 0, chunk = 0;
    for (int w = 0; w < NT / 32; ++w) {
      const uint32_t x = s_warp[w];
      before += w < (int)warp ? x : 0;
      chunk += x;
    }
    if (hit) a.pool[base + done + before + __popc(bal & ((1u << lane) - 1))] = ((uint64_t)c << 32) | (a.gid_base + g);
    done += chunk;
    __syncthreads();  // s_warp is rewritten by the next chunk that holds a hit
  }
}

}  // namespace nq
