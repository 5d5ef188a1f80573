// index.cu — K3: inverted-index build in HBM.  Replaces Index::insert_sketch and the
// vector<gid>[2^(W+S)] storage (/root/reference/src/niqki_index.cpp:362-370, niqki_index.h:55-56).
//
// A posting's list id is fp + cell*2^W and its cell is its position in the sketch, so the build is
// not a general sort: for every cell it is a stable counting sort of the n genomes by their W-bit
// fingerprint.  Layout in HBM (fixed cell strides, no global prefix needed; see internal.h):
//   dir [F][2^W]        : packed {begin,end} of list (cell,fp) inside the cell — one aligned word per
//                         probe (two u16 in a u32 when the shard has <= 65535 genomes, else uint2)
//   gids[F][gid_stride] : the cell's genomes (local ids, u16 or u32) ordered by (fp, gid) —
//                         gid-ascending inside a list, i.e. the reference's push_back order at
//                         OMP_NUM_THREADS=1.
// Step 1 transposes the int32 sketches [n][F] into u16 fingerprints [F][n_pad] (0xFFFF = not
// posted: the reference only posts 0 <= fp < range, :364) so that step 2 reads each cell's column
// coalesced.  Step 2 gives one warp per cell: histogram in shared memory (two 16-bit counters per
// word in the compact form), warp scan -> row[], then an ordered sweep that ranks equal
// fingerprints inside each 32-genome chunk with __match_any_sync so the scatter is stable.  The
// column is streamed with several 64-byte loads in flight per warp; 16+ warps per SM keep the
// serial cursor chain of the ordered sweep covered.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "device_common.cuh"
#include "internal.h"

namespace nq {

constexpr uint16_t kNoPost = 0xFFFFu;

// int32 [n][F] -> u16 [F][n_pad]
__global__ void __launch_bounds__(256) transpose_fp_kernel(const int32_t* __restrict__ sk, uint16_t* __restrict__ fpT,
                                                           uint32_t n, uint32_t F, uint32_t n_pad, uint32_t range) {
  __shared__ uint16_t tile[32][34];
  const uint32_t c0 = blockIdx.y * 32, g0 = blockIdx.x * 32;  // genomes along x: no 65535-block limit
  const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t g = g0 + ty + k * 8, c = c0 + tx;
    uint16_t v = kNoPost;
    if (g < n && c < F) {
      const uint32_t x = (uint32_t)sk[(size_t)g * F + c];
      if (x < range) v = (uint16_t)x;
    }
    tile[ty + k * 8][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t c = c0 + ty + k * 8, g = g0 + tx;
    if (c < F && g < n_pad) fpT[(size_t)c * n_pad + g] = tile[tx][ty + k * 8];
  }
}

// Histogram counter access: IT = uint16_t packs two counters per shared word (counts <= n < 65536).
template <typename IT>
struct Hist;
template <>
struct Hist<uint16_t> {
  static __device__ __forceinline__ size_t bytes(uint32_t range) { return (size_t)range * 2; }
  static __device__ __forceinline__ void add(uint32_t* h, uint32_t bin) { atomicAdd(&h[bin >> 1], 1u << ((bin & 1) * 16)); }
  static __device__ __forceinline__ uint32_t get(const uint32_t* h, uint32_t bin) {
    return reinterpret_cast<const uint16_t*>(h)[bin];
  }
  static __device__ __forceinline__ void set(uint32_t* h, uint32_t bin, uint32_t v) {
    reinterpret_cast<uint16_t*>(h)[bin] = (uint16_t)v;
  }
};
template <>
struct Hist<uint32_t> {
  static __device__ __forceinline__ size_t bytes(uint32_t range) { return (size_t)range * 4; }
  static __device__ __forceinline__ void add(uint32_t* h, uint32_t bin) { atomicAdd(&h[bin], 1u); }
  static __device__ __forceinline__ uint32_t get(const uint32_t* h, uint32_t bin) { return h[bin]; }
  static __device__ __forceinline__ void set(uint32_t* h, uint32_t bin, uint32_t v) { h[bin] = v; }
};

template <typename IT>
struct DirEntry;
template <>
struct DirEntry<uint16_t> {
  typedef uint32_t type;
  static __host__ __device__ __forceinline__ type make(uint32_t b, uint32_t e) { return b | (e << 16); }
  static __host__ __device__ __forceinline__ uint32_t begin(type w) { return w & 0xFFFFu; }
  static __host__ __device__ __forceinline__ uint32_t end(type w) { return w >> 16; }
};
template <>
struct DirEntry<uint32_t> {
  typedef uint2 type;
  static __host__ __device__ __forceinline__ type make(uint32_t b, uint32_t e) { return make_uint2(b, e); }
  static __host__ __device__ __forceinline__ uint32_t begin(type w) { return w.x; }
  static __host__ __device__ __forceinline__ uint32_t end(type w) { return w.y; }
};

// one warp per cell (warps loop over cells): stable counting sort of the cell's genomes by fingerprint
template <typename IT>
__global__ void __launch_bounds__(256) cell_sort_kernel(const uint16_t* __restrict__ fpT, uint32_t n, uint32_t n_pad,
                                                        uint32_t range, uint32_t F,
                                                        typename DirEntry<IT>::type* __restrict__ dir,
                                                        uint32_t row_stride, IT* __restrict__ gids, uint32_t gid_stride,
                                                        unsigned long long* __restrict__ total_postings) {
  extern __shared__ __align__(16) uint32_t smem[];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  uint32_t* cnt = smem + (size_t)warp * (Hist<IT>::bytes(range) / 4);
  constexpr unsigned kFull = 0xFFFFFFFFu;
  const unsigned lt = (1u << lane) - 1;
  constexpr int U = 8;  // chunks of 32 genomes in flight per warp
  unsigned long long posted = 0;

  for (uint32_t cell = blockIdx.x * nw + warp; cell < F; cell += gridDim.x * nw) {
    const uint16_t* col = fpT + (size_t)cell * n_pad;
    for (uint32_t i = lane; i < Hist<IT>::bytes(range) / 4; i += 32) cnt[i] = 0;
    __syncwarp();
    // ---- pass 1: histogram (n_pad is a multiple of 32; the padding holds kNoPost)
    for (uint32_t g0 = 0; g0 < n_pad; g0 += 32 * U) {
      uint16_t fp[U];
#pragma unroll
      for (int u = 0; u < U; ++u) fp[u] = g0 + u * 32 < n_pad ? col[g0 + u * 32 + lane] : kNoPost;
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (fp[u] != kNoPost) Hist<IT>::add(cnt, fp[u]);
    }
    __syncwarp();
    // ---- exclusive scan of the histogram, 32 bins per round; cnt[] becomes the write cursor
    typename DirEntry<IT>::type* myrow = dir + (size_t)cell * row_stride;
    uint32_t carry = 0;
    for (uint32_t b0 = 0; b0 < range; b0 += 32) {
      const uint32_t bin = b0 + lane;
      const uint32_t c = bin < range ? Hist<IT>::get(cnt, bin) : 0;
      uint32_t incl = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, incl, d);
        if (lane >= d) incl += t;
      }
      const uint32_t excl = carry + incl - c;
      if (bin < range) {
        Hist<IT>::set(cnt, bin, excl);
        myrow[bin] = DirEntry<IT>::make(excl, excl + c);
      }
      carry += __shfl_sync(kFull, incl, 31);
    }
    if (lane == 0) posted += carry;
    __syncwarp();
    // ---- pass 2: ordered scatter
    IT* out = gids + (size_t)cell * gid_stride;
    for (uint32_t g0 = 0; g0 < n_pad; g0 += 32 * U) {
      uint16_t fp[U];
#pragma unroll
      for (int u = 0; u < U; ++u) fp[u] = g0 + u * 32 < n_pad ? col[g0 + u * 32 + lane] : kNoPost;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (g0 + u * 32 >= n_pad) break;
        const bool valid = fp[u] != kNoPost;
        const unsigned same = __match_any_sync(kFull, fp[u]);
        const uint32_t rank = __popc(same & lt);
        const uint32_t base = valid ? Hist<IT>::get(cnt, fp[u]) : 0;
        __syncwarp();
        if (valid && rank == 0) Hist<IT>::set(cnt, fp[u], base + __popc(same));
        __syncwarp();
        if (valid) out[base + rank] = (IT)(g0 + u * 32 + lane);
      }
    }
    __syncwarp();
  }
  if (lane == 0 && posted) atomicAdd(total_postings, posted);  // `posted` is only maintained by lane 0
}


// ------------------------------------------------------------------------------------------
// Few cells, many entries (reads as entries: --indexlines at S = 8 has 256 cells and millions of
// ids): one warp per cell leaves the machine empty, so the cell's column is cut into chunks of
// kChunk entries and the counting sort becomes three kernels over (cell, chunk) tasks —
//   A. chunk_hist_kernel:    per-task histogram of the W-bit keys           -> H[task][bin]
//   B. chunk_scan_kernel:    per cell, bin-major / chunk-minor exclusive scan -> directory row,
//                            H[task][bin] becomes the task's first slot of list `bin`
//   C. chunk_scatter_kernel: per task, the ordered scatter of cell_sort_kernel from those cursors
// (lists stay gid-ascending because chunks are gid ranges and the scatter inside one is stable).
// ------------------------------------------------------------------------------------------
constexpr uint32_t kChunk = 65536;

__global__ void __launch_bounds__(256) chunk_hist_kernel(const uint16_t* __restrict__ fpT, uint32_t n_pad, uint32_t range,
                                                         uint32_t chunks, uint64_t tasks, uint32_t* __restrict__ H) {
  extern __shared__ __align__(16) uint32_t smem[];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  uint32_t* cnt = smem + (size_t)warp * range;
  constexpr int U = 8;
  for (uint64_t task = (uint64_t)blockIdx.x * nw + warp; task < tasks; task += (uint64_t)gridDim.x * nw) {
    const uint32_t cell = (uint32_t)(task / chunks), chunk = (uint32_t)(task % chunks);
    const uint16_t* col = fpT + (size_t)cell * n_pad;
    const uint32_t g_begin = chunk * kChunk, g_end = min(n_pad, g_begin + kChunk);
    for (uint32_t i = lane; i < range; i += 32) cnt[i] = 0;
    __syncwarp();
    for (uint32_t g0 = g_begin; g0 < g_end; g0 += 32 * U) {
      uint16_t fp[U];
#pragma unroll
      for (int u = 0; u < U; ++u) fp[u] = g0 + u * 32 < g_end ? col[g0 + u * 32 + lane] : kNoPost;
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (fp[u] != kNoPost) atomicAdd(&cnt[fp[u]], 1u);
    }
    __syncwarp();
    uint32_t* h = H + task * range;
    for (uint32_t i = lane; i < range; i += 32) h[i] = cnt[i];
    __syncwarp();
  }
}

template <int NT>
__global__ void __launch_bounds__(NT) chunk_scan_kernel(uint32_t* __restrict__ H, uint32_t range, uint32_t chunks,
                                                        uint2* __restrict__ dir, uint32_t row_stride,
                                                        unsigned long long* __restrict__ total_postings) {
  extern __shared__ __align__(16) uint32_t s_tot[];  // [range] list sizes, then list begins
  __shared__ uint32_t s_warp[NT / 32];
  const uint32_t cell = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t* h = H + (size_t)cell * chunks * range;
  for (uint32_t bin = tid; bin < range; bin += NT) {  // chunk-minor running sums, coalesced over bins
    uint32_t run = 0;
    for (uint32_t c = 0; c < chunks; ++c) {
      const uint32_t v = h[(size_t)c * range + bin];
      h[(size_t)c * range + bin] = run;
      run += v;
    }
    s_tot[bin] = run;
  }
  __syncthreads();
  // exclusive scan over the bins: every thread owns `per` consecutive bins
  const uint32_t per = (range + NT - 1) / NT, b0 = tid * per, b1 = min(range, b0 + per);
  uint32_t mine = 0;
  for (uint32_t b = b0; b < b1; ++b) mine += s_tot[b];
  uint32_t incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  uint32_t before = incl - mine;
  for (uint32_t w = 0; w < warp; ++w) before += s_warp[w];
  uint2* myrow = dir + (size_t)cell * row_stride;
  for (uint32_t b = b0; b < b1; ++b) {
    const uint32_t c = s_tot[b];
    myrow[b] = make_uint2(before, before + c);
    s_tot[b] = before;
    before += c;
  }
  if (tid == NT - 1 && before) atomicAdd(total_postings, (unsigned long long)before);
  __syncthreads();
  for (uint32_t bin = tid; bin < range; bin += NT) {
    const uint32_t base = s_tot[bin];
    for (uint32_t c = 0; c < chunks; ++c) h[(size_t)c * range + bin] += base;
  }
}

__global__ void __launch_bounds__(256) chunk_scatter_kernel(const uint16_t* __restrict__ fpT, uint32_t n_pad, uint32_t range,
                                                            uint32_t chunks, uint64_t tasks, const uint32_t* __restrict__ H,
                                                            uint32_t* __restrict__ gids, uint32_t gid_stride) {
  extern __shared__ __align__(16) uint32_t smem[];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  uint32_t* cur = smem + (size_t)warp * range;
  constexpr unsigned kFull = 0xFFFFFFFFu;
  const unsigned lt = (1u << lane) - 1;
  constexpr int U = 8;
  for (uint64_t task = (uint64_t)blockIdx.x * nw + warp; task < tasks; task += (uint64_t)gridDim.x * nw) {
    const uint32_t cell = (uint32_t)(task / chunks), chunk = (uint32_t)(task % chunks);
    const uint16_t* col = fpT + (size_t)cell * n_pad;
    const uint32_t g_begin = chunk * kChunk, g_end = min(n_pad, g_begin + kChunk);
    const uint32_t* h = H + task * range;
    for (uint32_t i = lane; i < range; i += 32) cur[i] = h[i];
    __syncwarp();
    uint32_t* out = gids + (size_t)cell * gid_stride;
    for (uint32_t g0 = g_begin; g0 < g_end; g0 += 32 * U) {
      uint16_t fp[U];
#pragma unroll
      for (int u = 0; u < U; ++u) fp[u] = g0 + u * 32 < g_end ? col[g0 + u * 32 + lane] : kNoPost;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (g0 + u * 32 >= g_end) break;
        const bool valid = fp[u] != kNoPost;
        const unsigned same = __match_any_sync(kFull, fp[u]);
        const uint32_t rank = __popc(same & lt);
        const uint32_t base = valid ? cur[fp[u]] : 0;
        __syncwarp();
        if (valid && rank == 0) cur[fp[u]] = base + __popc(same);
        __syncwarp();
        if (valid) out[base + rank] = g0 + u * 32 + lane;
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// Compact form (n <= kMaxCompact, u16 ids): one CTA per cell, everything in shared memory.
//   1. histogram of the cell's column (packed u16 counters, shared atomics)
//   2. block scan -> per-list begin; the directory row {begin,end} goes out coalesced
//   3. scatter with a RETURNING shared atomic on the list cursor, genomes taken in blocks of 2*NT
//      with a barrier in between: every genome gets a slot of its list, out of gid order only
//      against genomes of its own block
//   4. restore gid order inside every list: lists are tiny (2.4 genomes on average at n = 10k), so a
//      thread insertion-sorts the (almost sorted) lists of its bins; a list longer than kSerialSort is rebuilt by
//      the whole CTA from a bitmap of its members (ids are distinct, so enumerating the set bits IS
//      the sorted list) - near-duplicate genomes make such lists, random ones do not
//   5. the cell's posting array goes out with 16-byte stores
// The column is read twice (the second pass hits L2) instead of being staged, which keeps the
// footprint at 2 B per genome + 8 KB and 8 CTAs on an SM at n = 10k.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kSerialSort = 32;

template <int NT>
__global__ void __launch_bounds__(NT) cell_build_kernel(const uint16_t* __restrict__ fpT, uint32_t n, uint32_t n_pad,
                                                        uint32_t range, uint32_t* __restrict__ dir,
                                                        uint32_t row_stride, uint16_t* __restrict__ gids,
                                                        uint32_t gid_stride, uint32_t max_long,
                                                        unsigned long long* __restrict__ total_postings) {
  extern __shared__ __align__(16) uint32_t smem[];
  __shared__ uint32_t s_warp[NT / 32];
  __shared__ uint32_t s_nlong;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t cell = blockIdx.x;
  constexpr unsigned kFull = 0xFFFFFFFFu;
  // shared layout: cursors (packed u16) | out ids (u16) | bitmap | long-list bins (u16)
  const uint32_t cur_words = (max(range / 2, 1u) + 3) & ~3u;  // keeps `out` 16-byte aligned
  uint32_t* cur = smem;
  uint16_t* out = reinterpret_cast<uint16_t*>(cur + cur_words);
  uint32_t* bitmap = reinterpret_cast<uint32_t*>(out + n_pad);  // n_pad is a multiple of 32: stays 4-byte aligned
  uint16_t* long_bins = reinterpret_cast<uint16_t*>(bitmap + n_pad / 32);
  const uint16_t* col = fpT + (size_t)cell * n_pad;
  const uint4* col4 = reinterpret_cast<const uint4*>(col);  // n_pad * 2 B is a multiple of 64 B
  const uint32_t nvec = n_pad / 8;

  for (uint32_t i = tid; i < cur_words; i += NT) cur[i] = 0;
  if (tid == 0) s_nlong = 0;
  __syncthreads();

  // ---- 1. histogram
  auto for_each_fp = [&](auto&& f) {
    for (uint32_t v = tid; v < nvec; v += NT) {
      const uint4 q = __ldg(col4 + v);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t lo = w[j] & 0xFFFFu, hi = w[j] >> 16;
        if (lo != kNoPost) f(lo, v * 8 + 2 * j);
        if (hi != kNoPost) f(hi, v * 8 + 2 * j + 1);
      }
    }
  };
  for_each_fp([&](uint32_t fp, uint32_t) { atomicAdd(&cur[fp >> 1], 1u << ((fp & 1) * 16)); });
  __syncthreads();

  // ---- 2. exclusive scan over the bins; thread t owns bins [t*bpt, (t+1)*bpt), bpt even
  const uint32_t bpt = max(2u, range / NT);
  const uint32_t bin0 = tid * bpt;
  const bool owner = bin0 < range;
  uint32_t mine = 0;
  if (owner)
    for (uint32_t k = 0; k < bpt / 2; ++k) {
      const uint32_t w = cur[bin0 / 2 + k];
      mine += (w & 0xFFFFu) + (w >> 16);
    }
  uint32_t incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(kFull, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  uint32_t base = incl - mine;
  for (uint32_t w = 0; w < warp; ++w) base += s_warp[w];
  uint32_t posted = 0;
  for (uint32_t w = 0; w < NT / 32; ++w) posted += s_warp[w];
  if (owner) {
    uint32_t* myrow = dir + (size_t)cell * row_stride + bin0;
    uint32_t run = base;
    for (uint32_t k = 0; k < bpt / 2; ++k) {
      const uint32_t w = cur[bin0 / 2 + k];
      const uint32_t c0 = w & 0xFFFFu, c1 = w >> 16;
      const uint32_t b0 = run, b1 = run + c0;
      run = b1 + c1;
      cur[bin0 / 2 + k] = b0 | (b1 << 16);  // cursors start at the list begins
      if (range >= 2) {
        myrow[2 * k] = b0 | (b1 << 16);       // {begin, end} of bin 2k
        myrow[2 * k + 1] = b1 | (run << 16);  // and of bin 2k+1
      } else {
        myrow[0] = b0 | (b1 << 16);
      }
    }
  }
  __syncthreads();

  // ---- 3. scatter (second pass over the column): slot = cursor++ of the genome's list.  The CTA
  // walks the genomes in blocks of 2*NT with a barrier per block, so a list can only be out of gid
  // order between genomes of the same block: step 4 then sees almost sorted lists (~L compares).
  {
    const uint32_t* col32 = reinterpret_cast<const uint32_t*>(col);
    const uint32_t npairs = n_pad / 2, rounds = (npairs + NT - 1) / NT;
    uint32_t w_next = tid < npairs ? __ldg(col32 + tid) : 0xFFFFFFFFu;
    for (uint32_t r = 0; r < rounds; ++r) {
      const uint32_t w = w_next, p = r * NT + tid;
      w_next = p + NT < npairs ? __ldg(col32 + p + NT) : 0xFFFFFFFFu;
      const uint32_t fps[2] = {w & 0xFFFFu, w >> 16};
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (fps[j] != kNoPost) {
          const uint32_t sh = (fps[j] & 1) * 16;
          const uint32_t old = atomicAdd(&cur[fps[j] >> 1], 1u << sh);
          out[(old >> sh) & 0xFFFFu] = (uint16_t)(2 * p + j);
        }
      }
      __syncthreads();
    }
  }

  // ---- 4. gid order inside every list.  After the scatter cursor[bin] == end of bin == begin of bin+1.
  // Cursor words (two bins each) are dealt round-robin: fingerprints are heavily skewed (a few
  // hundred neighbouring bins hold most postings) and neighbouring lanes then get lists of similar
  // length.  Lists arrive almost sorted, so the common case per element is one load and one compare.
  for (uint32_t wi = tid; wi < range / 2; wi += NT) {
    const uint32_t w = cur[wi];
    uint32_t b = wi ? cur[wi - 1] >> 16 : 0u;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint32_t e = h ? w >> 16 : w & 0xFFFFu;
      const uint32_t len = e - b;
      if (len > kSerialSort) {
        const uint32_t at = atomicAdd(&s_nlong, 1u);
        if (at < max_long) long_bins[at] = (uint16_t)(2 * wi + h);
      } else if (len > 1) {
        uint32_t prev = out[b];
        for (uint32_t i = b + 1; i < e; ++i) {
          const uint32_t x = out[i];
          if (x >= prev) {
            prev = x;
            continue;
          }
          uint32_t j = i;  // insert x into the sorted prefix; its maximum (prev) ends up at i
          do {
            out[j] = out[j - 1];
            --j;
          } while (j > b && out[j - 1] > x);
          out[j] = (uint16_t)x;
        }
      }
      b = e;
    }
  }
  __syncthreads();
  const uint32_t nlong = min(s_nlong, max_long);  // max_long >= n / (kSerialSort+1) + 1: never exceeded
  const uint32_t nwords = n_pad / 32;
  const uint32_t wpt = (nwords + NT - 1) / NT;  // bitmap words per thread, contiguous
  for (uint32_t li = 0; li < nlong; ++li) {
    const uint32_t bin = long_bins[li];
    const uint32_t e = (cur[bin >> 1] >> ((bin & 1) * 16)) & 0xFFFFu;
    const uint32_t b = bin ? (cur[(bin - 1) >> 1] >> (((bin - 1) & 1) * 16)) & 0xFFFFu : 0u;
    for (uint32_t i = tid; i < nwords; i += NT) bitmap[i] = 0;
    __syncthreads();
    for (uint32_t i = b + tid; i < e; i += NT) {
      const uint32_t g = out[i];
      atomicOr(&bitmap[g >> 5], 1u << (g & 31));
    }
    __syncthreads();
    uint32_t cnt = 0;
    for (uint32_t k = 0; k < wpt; ++k) {
      const uint32_t w = tid * wpt + k;
      if (w < nwords) cnt += __popc(bitmap[w]);
    }
    uint32_t inc2 = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(kFull, inc2, d);
      if (lane >= d) inc2 += t;
    }
    __syncthreads();  // s_warp reuse
    if (lane == 31) s_warp[warp] = inc2;
    __syncthreads();
    uint32_t pos = b + inc2 - cnt;
    for (uint32_t w = 0; w < warp; ++w) pos += s_warp[w];
    for (uint32_t k = 0; k < wpt; ++k) {
      const uint32_t w = tid * wpt + k;
      if (w >= nwords) break;
      uint32_t bits = bitmap[w];
      while (bits) {
        const uint32_t bit = __ffs(bits) - 1;
        bits &= bits - 1;
        out[pos++] = (uint16_t)(w * 32 + bit);
      }
    }
    __syncthreads();
  }

  // ---- 5. posting array of the cell (gid_stride * 2 B is a multiple of 32 B)
  uint4* dst = reinterpret_cast<uint4*>(gids + (size_t)cell * gid_stride);
  const uint4* src = reinterpret_cast<const uint4*>(out);
  for (uint32_t v = tid; v < (posted + 7) / 8; v += NT) dst[v] = src[v];
  if (tid == 0 && posted) atomicAdd(total_postings, (unsigned long long)posted);
}


// ---- split16 side arrays (internal.h): u16 copy of the postings + {begin, mid, end} directory -------
__global__ void split16_gids_kernel(const uint32_t* __restrict__ g32, uint16_t* __restrict__ g16, size_t count) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
    g16[i] = (uint16_t)g32[i];
}
__global__ void split16_dir_kernel(const uint2* __restrict__ dir, const uint32_t* __restrict__ g32, uint4* __restrict__ dir3,
                                   uint32_t row_stride, uint32_t gid_stride, size_t entries) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < entries; i += (size_t)gridDim.x * blockDim.x) {
    const uint2 w = dir[i];
    const uint32_t* g = g32 + (i / row_stride) * gid_stride;
    uint32_t lo = w.x, hi = w.y;  // first posting >= 65536 in the gid-sorted list [w.x, w.y)
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (g[mid] < 65536u) lo = mid + 1; else hi = mid;
    }
    dir3[i] = make_uint4(w.x, lo, w.y, 0u);
  }
}

}  // namespace nq

using namespace nq;

int nq_index_make_split16(nq_index* ix) {
  static const char* env = nq_tuning_env("NQ_SPLIT16");  // "0": keep the u32 gather (measurement only)
  // below 65600 genomes the dummy ids of the query kernel (just above n) would not be >= 2^16
  if (ix->elem != 4 || ix->n < 65600u || ix->n > 131072u || ix->p.S > 15 || (env && env[0] == '0')) return NQ_OK;
  nq_ctx* ctx = ix->ctx;
  const size_t entries = (size_t)ix->p.F * ix->row_stride, count = (size_t)ix->p.F * ix->gid_stride;
  NQ_TRY(nq_dmalloc(ctx, (void**)&ix->d_dir3, entries * sizeof(uint4)));
  NQ_TRY(nq_dmalloc(ctx, (void**)&ix->d_gids16, (count + kQuerySlack) * sizeof(uint16_t)));
  split16_gids_kernel<<<ctx->sm_count * 16, 256, 0, ctx->stream>>>(static_cast<const uint32_t*>(ix->d_gids), ix->d_gids16, count);
  NQ_CHECK_LAUNCH(ctx);
  split16_dir_kernel<<<ctx->sm_count * 16, 256, 0, ctx->stream>>>(static_cast<const uint2*>(ix->d_row),
                                                                  static_cast<const uint32_t*>(ix->d_gids), ix->d_dir3,
                                                                  ix->row_stride, ix->gid_stride, entries);
  NQ_CHECK_LAUNCH(ctx);
  return NQ_OK;
}

template <typename IT>
static cudaError_t launch_cell_sort(nq_ctx* ctx, const uint16_t* d_fpT, uint32_t n, uint32_t n_pad, nq_index* ix,
                                    unsigned long long* d_total) {
  const uint32_t range = (uint32_t)ix->p.range, F = ix->p.F;
  const size_t per_warp = (size_t)range * sizeof(IT);
  // warps per CTA: as many as their 2^W counters fit (8 at W <= 12, 1 at W = 15 with u32 ids)
  uint32_t nw = 8;
  while (nw > 1 && per_warp * nw + 1024 > ctx->smem_optin) nw >>= 1;
  if (per_warp * nw + 1024 > ctx->smem_optin) return cudaErrorInvalidValue;
  const size_t smem = per_warp * nw;
  cudaError_t e = cudaFuncSetAttribute(cell_sort_kernel<IT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  // resident CTAs per SM by shared memory (8 warps each), capped by the 2048-thread limit
  const uint32_t per_sm = (uint32_t)std::max<size_t>(1, std::min<size_t>(8, (ctx->smem_optin + 1024) / (smem + 1024)));
  const uint32_t grid = std::min<uint32_t>((F + nw - 1) / nw, (uint32_t)ctx->sm_count * per_sm);
  NqTimer timer(ctx, NQK_CELLSORT);
  cell_sort_kernel<IT><<<grid, nw * 32, smem, ctx->stream>>>(d_fpT, n, n_pad, range, F,
                                                            static_cast<typename DirEntry<IT>::type*>(ix->d_row),
                                                            ix->row_stride, static_cast<IT*>(ix->d_gids), ix->gid_stride,
                                                            d_total);
  ctx->launches++;
  return cudaPeekAtLastError();
}

// u32 ids with few cells: the (cell, chunk) counting sort.  Returns cudaErrorInvalidValue when it does not apply.
static cudaError_t launch_chunked_sort(nq_ctx* ctx, const uint16_t* d_fpT, uint32_t n_pad, nq_index* ix,
                                       unsigned long long* d_total) {
  const uint32_t range = (uint32_t)ix->p.range, F = ix->p.F;
  const uint32_t chunks = (n_pad + kChunk - 1) / kChunk;
  const uint64_t tasks = (uint64_t)F * chunks;
  // worth it only when one warp per cell cannot fill the SMs and there is more than one chunk
  if (ix->elem != 4 || chunks < 2 || F >= (uint32_t)ctx->sm_count * 16 || tasks * range * 4 > (8ull << 30))
    return cudaErrorInvalidValue;
  uint32_t nw = 8;
  while (nw > 1 && (size_t)range * 4 * nw + 1024 > ctx->smem_optin) nw >>= 1;
  if ((size_t)range * 4 * nw + 1024 > ctx->smem_optin) return cudaErrorInvalidValue;
  const size_t smem = (size_t)range * 4 * nw;
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(chunk_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess ||
      (e = cudaFuncSetAttribute(chunk_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess ||
      (e = cudaFuncSetAttribute(chunk_scan_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)range * 4)) != cudaSuccess)
    return e;
  uint32_t* d_H = nullptr;
  if (nq_dmalloc(ctx, (void**)&d_H, tasks * range * 4) != NQ_OK) return cudaErrorMemoryAllocation;
  const uint32_t per_sm = (uint32_t)std::max<size_t>(1, std::min<size_t>(8, (ctx->smem_optin + 1024) / (smem + 1024)));
  const uint32_t grid = (uint32_t)std::min<uint64_t>((tasks + nw - 1) / nw, (uint64_t)ctx->sm_count * per_sm);
  {
    NqTimer timer(ctx, NQK_CELLSORT);
    chunk_hist_kernel<<<grid, nw * 32, smem, ctx->stream>>>(d_fpT, n_pad, range, chunks, tasks, d_H);
    chunk_scan_kernel<1024><<<F, 1024, (size_t)range * 4, ctx->stream>>>(d_H, range, chunks, static_cast<uint2*>(ix->d_row),
                                                                        ix->row_stride, d_total);
    chunk_scatter_kernel<<<grid, nw * 32, smem, ctx->stream>>>(d_fpT, n_pad, range, chunks, tasks, d_H,
                                                               static_cast<uint32_t*>(ix->d_gids), ix->gid_stride);
  }
  ctx->launches += 3;
  nq_dfree(ctx, d_H);
  return cudaPeekAtLastError();
}

// compact form: one CTA per cell.  Returns cudaErrorInvalidValue when the cell does not fit in shared memory.
static cudaError_t launch_cell_build(nq_ctx* ctx, const uint16_t* d_fpT, uint32_t n, uint32_t n_pad, nq_index* ix,
                                     unsigned long long* d_total) {
  const uint32_t range = (uint32_t)ix->p.range, F = ix->p.F;
  const uint32_t max_long = n / (kSerialSort + 1) + 1;
  const size_t smem = (size_t)((std::max(range / 2, 1u) + 3) & ~3u) * 4 + (size_t)n_pad * 2 + n_pad / 8 + (((size_t)max_long * 2 + 15) & ~(size_t)15);
  if (smem + 2048 > ctx->smem_optin || range < 2) return cudaErrorInvalidValue;
  const bool big = smem > 56 * 1024;
  cudaError_t e = big ? cudaFuncSetAttribute(cell_build_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                      : cudaFuncSetAttribute(cell_build_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  NqTimer timer(ctx, NQK_CELLSORT);
  if (big)
    cell_build_kernel<1024><<<F, 1024, smem, ctx->stream>>>(d_fpT, n, n_pad, range, static_cast<uint32_t*>(ix->d_row), ix->row_stride,
                                                           static_cast<uint16_t*>(ix->d_gids), ix->gid_stride, max_long, d_total);
  else
    cell_build_kernel<256><<<F, 256, smem, ctx->stream>>>(d_fpT, n, n_pad, range, static_cast<uint32_t*>(ix->d_row), ix->row_stride,
                                                         static_cast<uint16_t*>(ix->d_gids), ix->gid_stride, max_long, d_total);
  ctx->launches++;
  return cudaPeekAtLastError();
}

static void set_layout(nq_index* ix) {
  ix->elem = ix->n <= kMaxCompact ? 2u : 4u;
  const uint32_t per32 = 32 / ix->elem;
  ix->row_stride = (uint32_t)ix->p.range;  // directory entries per cell (2*elem bytes each)
  ix->gid_stride = (std::max<uint32_t>(ix->n, 1) + per32 - 1) / per32 * per32;
}

int nq_index_build_impl(nq_ctx* ctx, const nq_params* p, const int32_t* d_sketches, uint64_t n64,
                        uint32_t gid_base, nq_index** out) {
  NQ_TRY(nq_params_check(p));
  if (p->W > 15)  // u16 fingerprint transpose and per-cell shared-memory counters; the reference takes any W + S < 32
    return nq_set_error(NQ_ERR_UNSUPPORTED, "index build supports W <= 15 (got W=%u)", p->W);
  if (n64 == 0 || n64 > 0xFFFFFFF0ull || n64 + gid_base > 0xFFFFFFFFull)
    return nq_set_error(NQ_ERR_INVALID, "bad genome count %llu (gid_base %u)", (unsigned long long)n64, gid_base);
  const uint32_t n = (uint32_t)n64, F = p->F, range = (uint32_t)p->range;

  nq_index* ix = new nq_index();
  ix->ctx = ctx; ix->p = *p; ix->n = n; ix->gid_base = gid_base;
  set_layout(ix);
  const uint32_t n_pad = (n + 31) & ~31u;
  uint16_t* d_fpT = nullptr;
  unsigned long long* d_total = nullptr;
  auto fail = [&](int st) {
    nq_dfree(ctx, d_fpT); nq_dfree(ctx, d_total);
    nq_index_free(ix);
    return st;
  };
  cudaError_t e;
  int st;
  if ((st = nq_dmalloc(ctx, &ix->d_row, (size_t)F * ix->row_stride * ix->elem * 2)) != NQ_OK ||
      (st = nq_dmalloc(ctx, &ix->d_gids, ((size_t)F * ix->gid_stride + kQuerySlack) * ix->elem)) != NQ_OK)
    return fail(st);
  if ((st = nq_dmalloc(ctx, (void**)&d_fpT, (size_t)F * n_pad * 2)) != NQ_OK) return fail(st);
  if ((st = nq_dmalloc(ctx, (void**)&d_total, 8)) != NQ_OK) return fail(st);
  if (cudaMemsetAsync(d_total, 0, 8, ctx->stream) != cudaSuccess) return fail(nq_set_error(NQ_ERR_CUDA, "memset failed"));

  dim3 tg((n_pad + 31) / 32, (F + 31) / 32);
  if (tg.y > 65535) return fail(nq_set_error(NQ_ERR_UNSUPPORTED, "S > 21 is not supported by the index build"));
  {
    NqTimer timer(ctx, NQK_TRANSPOSE);
    transpose_fp_kernel<<<tg, 256, 0, ctx->stream>>>(d_sketches, d_fpT, n, F, n_pad, range);
  }
  ctx->launches++;
  e = cudaErrorInvalidValue;
  if (ix->elem == 2) e = launch_cell_build(ctx, d_fpT, n, n_pad, ix, d_total);
  if (e == cudaErrorInvalidValue) e = launch_chunked_sort(ctx, d_fpT, n_pad, ix, d_total);  // few cells, many entries
  if (e == cudaErrorInvalidValue)  // u32 ids, or a cell that does not fit in shared memory: one warp per cell
    e = ix->elem == 2 ? launch_cell_sort<uint16_t>(ctx, d_fpT, n, n_pad, ix, d_total)
                      : launch_cell_sort<uint32_t>(ctx, d_fpT, n, n_pad, ix, d_total);
  if (e == cudaErrorInvalidValue)
    return fail(nq_set_error(NQ_ERR_UNSUPPORTED, "2^W = %u fingerprint counters of one cell do not fit in shared memory", range));
  if (e != cudaSuccess) return fail(nq_set_error(NQ_ERR_CUDA, "cell_sort launch failed: %s", cudaGetErrorString(e)));
  unsigned long long total = 0;
  if ((e = cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess ||
      (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess)
    return fail(nq_set_error(NQ_ERR_CUDA, "index build failed: %s", cudaGetErrorString(e)));
  ix->n_postings = total;
  nq_dfree(ctx, d_fpT);
  nq_dfree(ctx, d_total);
  d_fpT = nullptr; d_total = nullptr;
  if ((st = nq_index_make_split16(ix)) != NQ_OK || (st = nq_slab_build(ix)) != NQ_OK || (st = nq_query_prepare(ix)) != NQ_OK)
    return fail(st);
  *out = ix;
  return NQ_OK;
}

extern "C" int nq_index_free(nq_index* ix) {
  if (!ix) return NQ_OK;
  if (ix->ctx) {
    cudaSetDevice(ix->ctx->device);
    nq_dfree(ix->ctx, ix->d_row);
    nq_dfree(ix->ctx, ix->d_gids);
    nq_dfree(ix->ctx, ix->d_dir3);
    nq_dfree(ix->ctx, ix->d_gids16);
    nq_dfree(ix->ctx, ix->d_pool);
    nq_slab_free(ix);
  }
  delete ix;
  return NQ_OK;
}

extern "C" int nq_index_info(const nq_index* ix, uint64_t* n_postings, uint32_t* n_genomes, uint32_t* gid_base,
                             uint64_t* device_bytes) {
  if (!ix) return nq_set_error(NQ_ERR_INVALID, "null index");
  if (n_postings) *n_postings = ix->n_postings;
  if (n_genomes) *n_genomes = ix->n;
  if (gid_base) *gid_base = ix->gid_base;
  if (device_bytes)
    *device_bytes = (uint64_t)ix->p.F * (2ull * ix->row_stride + ix->gid_stride) * ix->elem +
                    (ix->d_dir3 ? (uint64_t)ix->p.F * (16ull * ix->row_stride + 2ull * ix->gid_stride) : 0) +
                    (ix->slab_G ? ix->slab_granules * ix->slab_G * 2 + (uint64_t)ix->p.F * (ix->p.range / 32) * 16 : 0);
  return NQ_OK;
}

template <typename IT>
static int export_t(nq_index* ix, uint32_t* list_sizes, uint32_t* gids) {
  typedef typename DirEntry<IT>::type DT;
  nq_ctx* ctx = ix->ctx;
  const uint32_t F = ix->p.F, range = (uint32_t)ix->p.range;
  // stream cells through a bounded host staging buffer
  const size_t per_cell = (size_t)ix->row_stride * sizeof(DT) + (size_t)ix->gid_stride * sizeof(IT);
  const uint32_t cells_per_chunk = (uint32_t)std::max<size_t>(1, (64u << 20) / per_cell);
  std::vector<DT> hrow((size_t)cells_per_chunk * ix->row_stride);
  std::vector<IT> hg((size_t)cells_per_chunk * ix->gid_stride);
  uint64_t w = 0;
  for (uint32_t c0 = 0; c0 < F; c0 += cells_per_chunk) {
    const uint32_t nc = std::min(cells_per_chunk, F - c0);
    NQ_CUDA(cudaMemcpyAsync(hrow.data(), static_cast<DT*>(ix->d_row) + (size_t)c0 * ix->row_stride,
                            (size_t)nc * ix->row_stride * sizeof(DT), cudaMemcpyDeviceToHost, ctx->stream));
    if (gids)
      NQ_CUDA(cudaMemcpyAsync(hg.data(), static_cast<IT*>(ix->d_gids) + (size_t)c0 * ix->gid_stride,
                              (size_t)nc * ix->gid_stride * sizeof(IT), cudaMemcpyDeviceToHost, ctx->stream));
    NQ_CUDA(cudaStreamSynchronize(ctx->stream));
    for (uint32_t c = 0; c < nc; ++c) {
      const DT* r = &hrow[(size_t)c * ix->row_stride];
      if (list_sizes)
        for (uint32_t f = 0; f < range; ++f)
          list_sizes[(size_t)(c0 + c) * range + f] = DirEntry<IT>::end(r[f]) - DirEntry<IT>::begin(r[f]);
      if (gids) {
        const IT* g = &hg[(size_t)c * ix->gid_stride];
        const uint32_t cnt = DirEntry<IT>::end(r[range - 1]);  // lists are laid out back to back
        for (uint32_t i = 0; i < cnt; ++i) gids[w + i] = ix->gid_base + g[i];
        w += cnt;
      }
    }
  }
  return NQ_OK;
}

extern "C" int nq_index_export(nq_index* ix, uint32_t* list_sizes, uint32_t* gids, uint64_t gids_capacity) {
  NQ_RANGE();
  if (!ix) return nq_set_error(NQ_ERR_INVALID, "null index");
  if (gids && gids_capacity < ix->n_postings)
    return nq_set_error(NQ_ERR_OVERFLOW, "gids capacity %llu < %llu postings", (unsigned long long)gids_capacity,
                        (unsigned long long)ix->n_postings);
  NQ_CUDA(cudaSetDevice(ix->ctx->device));
  return ix->elem == 2 ? export_t<uint16_t>(ix, list_sizes, gids) : export_t<uint32_t>(ix, list_sizes, gids);
}

template <typename IT>
static int import_t(nq_index* ix, const uint32_t* list_sizes, const uint32_t* gids) {
  typedef typename DirEntry<IT>::type DT;
  nq_ctx* ctx = ix->ctx;
  const uint32_t F = ix->p.F, range = (uint32_t)ix->p.range;
  std::vector<DT> hrow((size_t)F * ix->row_stride);
  std::vector<IT> hg((size_t)F * ix->gid_stride, 0);
  uint64_t r = 0, total = 0;
  for (uint32_t c = 0; c < F; ++c) {
    uint32_t run = 0;
    DT* rowp = &hrow[(size_t)c * ix->row_stride];
    IT* gp = &hg[(size_t)c * ix->gid_stride];
    for (uint32_t f = 0; f < range; ++f) {
      const uint32_t begin = run;
      const uint32_t sz = list_sizes[(size_t)c * range + f];
      for (uint32_t j = 0; j < sz; ++j) {
        const uint32_t g = gids[r + j];
        if (g >= ix->gid_base && g - ix->gid_base < ix->n) {
          if (run >= ix->n)
            return nq_set_error(NQ_ERR_INVALID, "cell %u holds more than %u postings of this shard", c, ix->n);
          gp[run++] = (IT)(g - ix->gid_base);
        }
      }
      // the reference's own dumps hold a list in the order its OpenMP threads appended; the
      // split16 directory and the granule form rely on gid order, and no result depends on list order
      if (run - begin > 1 && !std::is_sorted(gp + begin, gp + run)) std::sort(gp + begin, gp + run);
      r += sz;
      rowp[f] = DirEntry<IT>::make(begin, run);
    }
    total += run;
  }
  ix->n_postings = total;
  cudaError_t e = cudaSuccess;
  if (nq_dmalloc(ctx, &ix->d_row, hrow.size() * sizeof(DT)) != NQ_OK ||
      nq_dmalloc(ctx, &ix->d_gids, (hg.size() + kQuerySlack) * sizeof(IT)) != NQ_OK ||
      (e = cudaMemcpyAsync(ix->d_row, hrow.data(), hrow.size() * sizeof(DT), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
      (e = cudaMemcpyAsync(ix->d_gids, hg.data(), hg.size() * sizeof(IT), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
      (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess)
    return nq_set_error(NQ_ERR_CUDA, "index import failed: %s", cudaGetErrorString(e));
  NQ_TRY(nq_index_make_split16(ix));
  NQ_TRY(nq_slab_build(ix));
  return nq_query_prepare(ix);
}

extern "C" int nq_index_import(nq_ctx* ctx, const nq_params* p, const uint32_t* list_sizes, const uint32_t* gids,
                               uint32_t n_genomes, uint32_t gid_base, nq_index** out) {
  NQ_RANGE();
  NQ_TRY(nq_params_check(p));
  if (!ctx || !list_sizes || !gids || !out) return nq_set_error(NQ_ERR_INVALID, "null argument");
  if (n_genomes == 0) return nq_set_error(NQ_ERR_INVALID, "empty shard");
  NQ_CUDA(cudaSetDevice(ctx->device));
  nq_index* ix = new nq_index();
  ix->ctx = ctx; ix->p = *p; ix->n = n_genomes; ix->gid_base = gid_base;
  set_layout(ix);
  const int st = ix->elem == 2 ? import_t<uint16_t>(ix, list_sizes, gids) : import_t<uint32_t>(ix, list_sizes, gids);
  if (st != NQ_OK) {
    nq_index_free(ix);
    return st;
  }
  *out = ix;
  return NQ_OK;
}
