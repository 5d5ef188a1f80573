// query.cu — K4a: posting-list gather + per-genome hit counting + fused min_score threshold and
// compaction.  Replaces Index::query_sketch (/root/reference/src/niqki_index.cpp:633-687).
//
// One CTA per query sketch; all queries of a batch are resident at once where shared memory allows
// (256-thread CTAs, 8 per SM), so the wave sweeps the cells roughly in step and a cell's directory
// and lists are served from L2 after their first touch.  The per-genome counters of the shard live
// in shared memory (two 16-bit counters per word when S <= 15, because a count can never exceed
// F = 2^S <= 32768 — the same widths the reference uses, :635-683).
//
// The gather is bound by the L1TEX tag stage (one cache line per cycle per SM for divergent
// loads), not by bytes, so it is organised to touch as few (line, instruction) pairs as possible:
// a warp takes 32 consecutive cells, every lane reads its cell's fingerprint (coalesced) and ONE
// packed directory word {begin,end}; the 32 lists are then treated as one concatenated stream
// that the whole warp walks 32 postings at a time — lanes reading neighbouring postings of one list
// share a cache line, no lane idles on a short list, and every shared-memory atomicAdd carries 32
// useful lanes.  The owner of stream slot s is found without a search: the lanes whose list starts
// inside the current 32-slot round set one bit each (REDUX.OR), and a population count of the
// bits at or below a slot gives the owner's rank among the non-empty lists.
// The threshold pass then compacts (count, gid) pairs in gid order into a global pool (one
// atomicAdd per query reserves the segment).  Sorting by (count, gid) descending (:685) is done by
// the host when it merges shards.
#include <algorithm>
#include <vector>

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
};

enum CountMode { kPack16 = 0, kSmem32 = 1, kGlobal32 = 2 };

template <typename IT>
__device__ __forceinline__ void load_dir(const void* dir, size_t at, uint32_t& b, uint32_t& e);
template <>
__device__ __forceinline__ void load_dir<uint16_t>(const void* dir, size_t at, uint32_t& b, uint32_t& e) {
  const uint32_t w = __ldg(static_cast<const uint32_t*>(dir) + at);
  b = w & 0xFFFFu;
  e = w >> 16;
}
template <>
__device__ __forceinline__ void load_dir<uint32_t>(const void* dir, size_t at, uint32_t& b, uint32_t& e) {
  const uint2 w = __ldg(static_cast<const uint2*>(dir) + at);
  b = w.x;
  e = w.y;
}

// IDX = uint32_t when every posting index F*gid_stride fits 32 bits (always for S <= 15 in the compact form)
template <typename IT, int MODE, int NT, typename IDX>
__global__ void __launch_bounds__(NT) query_count_kernel(QueryArgs a, uint64_t q0) {
  extern __shared__ __align__(16) uint32_t smem[];
  __shared__ IDX s_src[NT / 32][32];  // per warp: stream base of the rank-th non-empty list
  __shared__ uint32_t s_warp[NT / 32];
  __shared__ unsigned long long s_base;
  __shared__ uint32_t s_total;
  const uint64_t q = q0 + blockIdx.x;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t* cnt = MODE == kGlobal32 ? a.gcounts + (size_t)blockIdx.x * a.n : smem;
  const uint32_t words = MODE == kPack16 ? (a.n + 1) / 2 : a.n;
  constexpr unsigned kFull = 0xFFFFFFFFu;

  for (uint32_t i = tid; i < words; i += NT) cnt[i] = 0;
  __syncthreads();

  // ---- gather + count (:654-660)
  const int32_t* sk = a.qsk + q * a.F;
  const IT* gids = static_cast<const IT*>(a.gids);
  const unsigned le = 0xFFFFFFFFu >> (31 - lane);  // bits at or below this lane
  uint32_t gathered = 0;
  auto count = [&](uint32_t l) {
    if (MODE == kPack16)
      atomicAdd(&cnt[l >> 1], (l & 1) ? 0x10000u : 1u);
    else
      atomicAdd(&cnt[l], 1u);
  };
  // software pipeline over the warp's groups of 32 cells: fingerprints are fetched two groups
  // ahead, directory words one group ahead, so neither latency sits in front of the stream walk
  const uint32_t step = NT;
  uint32_t c_cur = warp * 32;
  uint32_t fp_next = 0xFFFFFFFFu, fp_next2 = 0xFFFFFFFFu;
  if (c_cur + step + lane < a.F) fp_next = (uint32_t)__ldg(&sk[c_cur + step + lane]);
  uint32_t b = 0, e = 0;
  if (c_cur + lane < a.F) {
    const uint32_t fp = (uint32_t)__ldg(&sk[c_cur + lane]);
    if (fp < a.range) load_dir<IT>(a.dir, (size_t)(c_cur + lane) * a.row_stride + fp, b, e);  // 0 <= fp < range (:655)
  }
  for (; c_cur < a.F; c_cur += step) {
    const uint32_t cell = c_cur + lane;
    // prefetch: fingerprint of group +2, directory word of group +1
    fp_next2 = 0xFFFFFFFFu;
    if (c_cur + 2 * step + lane < a.F) fp_next2 = (uint32_t)__ldg(&sk[c_cur + 2 * step + lane]);
    uint32_t nb = 0, ne = 0;
    if (fp_next < a.range) load_dir<IT>(a.dir, (size_t)(cell + step) * a.row_stride + fp_next, nb, ne);

    const uint32_t len = e - b;
    uint32_t incl = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(kFull, incl, d);
      if (lane >= d) incl += t;
    }
    const uint32_t excl = incl - len;
    const uint32_t total = __shfl_sync(kFull, incl, 31);
    const unsigned nonempty = __ballot_sync(kFull, len != 0);
    if (len) s_src[warp][__popc(nonempty & (le >> 1))] = (IDX)((IDX)cell * a.gid_stride + b - excl);
    __syncwarp();
    gathered += len;
    uint32_t starts_before = 0;
    constexpr int R = 4;  // rounds of 32 postings whose loads are in flight together
    for (uint32_t r0 = 0; r0 < total; r0 += 32 * R) {
      uint32_t l[R];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const uint32_t rel = excl - (r0 + 32 * k);  // wraps to a huge value when the list starts before this round
        const unsigned m = __reduce_or_sync(kFull, (len && rel < 32u) ? 1u << rel : 0u);
        const uint32_t s = r0 + 32 * k + lane;
        l[k] = 0xFFFFFFFFu;
        if (s < total) {
          const uint32_t owner = starts_before + __popc(m & le) - 1;
          l[k] = gids[s_src[warp][owner] + s];
        }
        starts_before += __popc(m);
      }
#pragma unroll
      for (int k = 0; k < R; ++k)
        if (l[k] != 0xFFFFFFFFu) count(l[k]);
    }
    __syncwarp();
    b = nb; e = ne;
    fp_next = fp_next2;
  }
  __syncthreads();

  auto count_of = [&](uint32_t g) -> uint32_t {
    uint32_t c;
    if (MODE == kPack16) c = (cnt[g >> 1] >> ((g & 1) * 16)) & 0xFFFFu;
    else if (MODE == kGlobal32) c = __ldcg(&cnt[g]);
    else c = cnt[g];
    return c & a.wrap_mask;
  };

  // ---- threshold (:661-665): count hits, reserve a pool segment, then write them in gid order
  uint32_t mine = 0;
  for (uint32_t g = tid; g < a.n; g += NT) mine += count_of(g) >= a.min_score;
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
    const unsigned bal = __ballot_sync(kFull, hit);
    __syncthreads();  // s_warp reuse
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
  }
}

}  // namespace nq

using namespace nq;

template <typename IT, int MODE, int NT>
static cudaError_t launch_query_t(size_t smem, unsigned nb, const QueryArgs& a, uint64_t q0, cudaStream_t st) {
  const bool idx32 = (uint64_t)a.F * a.gid_stride < (1ull << 32);
  cudaError_t e;
  if (idx32) {
    e = cudaFuncSetAttribute(query_count_kernel<IT, MODE, NT, uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    query_count_kernel<IT, MODE, NT, uint32_t><<<nb, NT, smem, st>>>(a, q0);
  } else {
    e = cudaFuncSetAttribute(query_count_kernel<IT, MODE, NT, uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    query_count_kernel<IT, MODE, NT, uint64_t><<<nb, NT, smem, st>>>(a, q0);
  }
  return cudaSuccess;
}
template <typename IT>
static cudaError_t launch_query_it(int mode, size_t smem, unsigned nb, const QueryArgs& a, uint64_t q0, cudaStream_t st) {
  // small counter arrays: 256-thread CTAs so that 8 queries share an SM; large ones: one big CTA
  const bool small = smem <= 24 * 1024;
  if (mode == kPack16) return small ? launch_query_t<IT, kPack16, 256>(smem, nb, a, q0, st) : launch_query_t<IT, kPack16, 1024>(smem, nb, a, q0, st);
  if (mode == kSmem32) return small ? launch_query_t<IT, kSmem32, 256>(smem, nb, a, q0, st) : launch_query_t<IT, kSmem32, 1024>(smem, nb, a, q0, st);
  return launch_query_t<IT, kGlobal32, 512>(0, nb, a, q0, st);
}
static cudaError_t launch_query(uint32_t elem, int mode, size_t smem, unsigned nb, const QueryArgs& a, uint64_t q0,
                                cudaStream_t st) {
  return elem == 2 ? launch_query_it<uint16_t>(mode, smem, nb, a, q0, st) : launch_query_it<uint32_t>(mode, smem, nb, a, q0, st);
}

int nq_query_impl(nq_index* ix, const int32_t* d_sketches, uint64_t nq, uint32_t min_score, nq_hits** out) {
  if (!ix) return nq_set_error(NQ_ERR_INVALID, "null index");
  nq_ctx* ctx = ix->ctx;
  if (out) *out = nullptr;
  nq_hits* hits = out ? new nq_hits() : nullptr;
  if (hits) hits->ptr.assign(nq + 1, 0);
  if (nq == 0) {
    if (out) *out = hits;
    return NQ_OK;
  }
  const nq_params& p = ix->p;
  QueryArgs a{};
  a.qsk = d_sketches; a.dir = ix->d_row; a.gids = ix->d_gids;
  a.F = p.F; a.range = (uint32_t)p.range; a.n = ix->n; a.row_stride = ix->row_stride; a.gid_stride = ix->gid_stride;
  a.gid_base = ix->gid_base;
  a.min_score = min_score;
  a.wrap_mask = p.S <= 7 ? 0xFFu : p.S <= 15 ? 0xFFFFu : 0xFFFFFFFFu;  // counter widths of :635/:651/:667

  int mode;
  size_t smem;
  if (p.S <= 15 && (size_t)((ix->n + 1) / 2) * 4 <= ctx->smem_optin) { mode = kPack16; smem = (size_t)((ix->n + 1) / 2) * 4; }
  else if ((size_t)ix->n * 4 <= ctx->smem_optin) { mode = kSmem32; smem = (size_t)ix->n * 4; }
  else { mode = kGlobal32; smem = 0; }

  // queries per launch: everything at once unless global counters would be too large
  uint64_t q_per_launch = nq;
  if (mode == kGlobal32) q_per_launch = std::max<uint64_t>(1, std::min<uint64_t>(nq, (1ull << 30) / ((uint64_t)ix->n * 4)));

  unsigned long long* d_cursor = nullptr;
  uint64_t* d_begin = nullptr;
  uint32_t* d_n = nullptr;
  uint32_t* d_gcounts = nullptr;
  int st = NQ_OK;
  auto cleanup = [&]() {
    nq_dfree(ctx, d_cursor); nq_dfree(ctx, d_begin); nq_dfree(ctx, d_n); nq_dfree(ctx, d_gcounts);
  };
  auto fail = [&](int s) {
    cleanup();
    delete hits;
    return s;
  };
  if ((st = nq_dmalloc(ctx, (void**)&d_cursor, 16)) || (st = nq_dmalloc(ctx, (void**)&d_begin, nq * 8)) ||
      (st = nq_dmalloc(ctx, (void**)&d_n, nq * 4)))
    return fail(st);
  if (mode == kGlobal32 && (st = nq_dmalloc(ctx, (void**)&d_gcounts, q_per_launch * ix->n * 4))) return fail(st);

  // pool capacity: exact when every genome is reported (min_score == 0, the CLI default), otherwise a
  // guess that is corrected from the cursor after the first run
  uint64_t want = min_score == 0 ? nq * (uint64_t)ix->n : std::max<uint64_t>(1u << 20, nq * 256);
  want = std::min<uint64_t>(want, nq * (uint64_t)ix->n);
  unsigned long long total = 0, stats[2] = {0, 0};
  for (int attempt = 0; attempt < 2; ++attempt) {
    if (ix->pool_cap < want) {
      nq_dfree(ctx, ix->d_pool);
      ix->d_pool = nullptr;
      ix->pool_cap = 0;
      if ((st = nq_dmalloc(ctx, (void**)&ix->d_pool, want * 8)) != NQ_OK) return fail(st);
      ix->pool_cap = want;
    }
    a.pool = ix->d_pool; a.pool_cap = ix->pool_cap; a.cursor = d_cursor; a.hit_begin = d_begin; a.hit_n = d_n;
    a.gcounts = d_gcounts;
    if (cudaMemsetAsync(d_cursor, 0, 16, ctx->stream) != cudaSuccess) return fail(nq_set_error(NQ_ERR_CUDA, "memset failed"));
    NqTimer* timer = new NqTimer(ctx, NQK_QUERY);
    for (uint64_t q0 = 0; q0 < nq; q0 += q_per_launch) {
      const unsigned nb = (unsigned)std::min<uint64_t>(q_per_launch, nq - q0);
      const cudaError_t e0 = launch_query(ix->elem, mode, smem, nb, a, q0, ctx->stream);
      cudaError_t e = e0;
      ctx->launches++;
      if (e != cudaSuccess || (e = cudaPeekAtLastError()) != cudaSuccess) {
        delete timer;
        return fail(nq_set_error(NQ_ERR_CUDA, "query kernel launch failed: %s", cudaGetErrorString(e)));
      }
    }
    delete timer;
    if (!out && min_score == 0) break;  // capacity is exact; leave everything on the device
    cudaError_t e;
    if ((e = cudaMemcpyAsync(stats, d_cursor, 16, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess)
      return fail(nq_set_error(NQ_ERR_CUDA, "query failed: %s", cudaGetErrorString(e)));
    total = stats[0];
    ctx->last_query_gathered = stats[1];
    if (total <= ix->pool_cap) break;
    want = total;  // second attempt with the exact size
  }

  if (out) {
    std::vector<uint64_t> hbegin(nq);
    std::vector<uint32_t> hn(nq);
    std::vector<uint64_t> pool(total);
    cudaError_t e;
    if ((e = cudaMemcpyAsync(hbegin.data(), d_begin, nq * 8, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess ||
        (e = cudaMemcpyAsync(hn.data(), d_n, nq * 4, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess ||
        (total && (e = cudaMemcpyAsync(pool.data(), ix->d_pool, total * 8, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) ||
        (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess)
      return fail(nq_set_error(NQ_ERR_CUDA, "query result copy failed: %s", cudaGetErrorString(e)));
    hits->counts.resize(total);
    hits->gids.resize(total);
    uint64_t w = 0;
    for (uint64_t q = 0; q < nq; ++q) {
      uint64_t* seg = pool.data() + hbegin[q];
      // std::greater<pair<count,gid>> (:685) == descending order of count<<32|gid
      std::sort(seg, seg + hn[q], std::greater<uint64_t>());
      hits->ptr[q] = w;
      for (uint32_t i = 0; i < hn[q]; ++i, ++w) {
        hits->counts[w] = (uint32_t)(seg[i] >> 32);
        hits->gids[w] = (uint32_t)seg[i];
      }
    }
    hits->ptr[nq] = w;
    *out = hits;
  }
  cleanup();
  return NQ_OK;
}
