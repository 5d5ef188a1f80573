// query.cu — K4a: posting-list gather + per-genome hit counting + fused min_score threshold and
// compaction.  Replaces Index::query_sketch (/root/reference/src/niqki_index.cpp:633-687).
//
// One CTA per query sketch.  The per-genome counters of the shard live in shared memory (two
// 16-bit counters per word when S <= 15, because a count can never exceed F = 2^S <= 32768 — the
// same widths the reference uses, :635-683), every thread probes a strided subset of the F
// cells: row[cell][fp] / row[cell][fp+1] delimit the list, whose gids are gathered and counted
// with shared-memory atomics.  The threshold pass then compacts (count, gid) pairs in gid order
// into a global pool (one atomicAdd per query reserves the segment).  Sorting by (count, gid)
// descending (:685) is done by the host when it merges shards.
#include <algorithm>
#include <vector>

#include "device_common.cuh"
#include "internal.h"

namespace nq {

struct QueryArgs {
  const int32_t* qsk;    // [nq][F]
  const uint32_t* row;   // [F][range+1]
  const uint32_t* gids;  // [F][n_stride]
  uint32_t F, range, n, n_stride, gid_base, min_score, wrap_mask;
  uint64_t* pool;  // count<<32 | gid
  uint64_t pool_cap;
  unsigned long long* cursor;  // [0] pool cursor, [1] posting entries gathered (statistics)
  uint64_t* hit_begin;  // [nq]
  uint32_t* hit_n;      // [nq]
  uint32_t* gcounts;    // global counters [gridDim.x][n] (GLOBAL mode only)
};

enum CountMode { kPack16 = 0, kSmem32 = 1, kGlobal32 = 2 };

template <int MODE, int NT>
__global__ void __launch_bounds__(NT) query_count_kernel(QueryArgs a, uint64_t q0) {
  extern __shared__ __align__(16) uint32_t smem[];
  __shared__ uint32_t s_warp[NT / 32];
  __shared__ unsigned long long s_base;
  __shared__ uint32_t s_total;
  const uint64_t q = q0 + blockIdx.x;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t* cnt = MODE == kGlobal32 ? a.gcounts + (size_t)blockIdx.x * a.n : smem;
  const uint32_t words = MODE == kPack16 ? (a.n + 1) / 2 : a.n;

  for (uint32_t i = tid; i < words; i += NT) cnt[i] = 0;
  __syncthreads();

  // ---- gather + count (:654-660)
  const int32_t* sk = a.qsk + q * a.F;
  uint32_t gathered = 0;
  for (uint32_t cell = tid; cell < a.F; cell += NT) {
    const uint32_t fp = (uint32_t)sk[cell];
    if (fp < a.range) {  // 0 <= fp < fingerprint_range (:655)
      const uint32_t* r = a.row + (size_t)cell * (a.range + 1) + fp;
      const uint32_t b = r[0], e = r[1];
      const uint32_t* g = a.gids + (size_t)cell * a.n_stride;
      gathered += e - b;
      for (uint32_t j = b; j < e; ++j) {
        const uint32_t l = g[j] - a.gid_base;
        if (MODE == kPack16)
          atomicAdd(&cnt[l >> 1], 1u << ((l & 1) * 16));
        else
          atomicAdd(&cnt[l], 1u);
      }
    }
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
    mine += __shfl_xor_sync(0xFFFFFFFFu, mine, d);
    gathered += __shfl_xor_sync(0xFFFFFFFFu, gathered, d);
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
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, hit);
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
  constexpr int NT = 512;
  QueryArgs a{};
  a.qsk = d_sketches; a.row = ix->d_row; a.gids = ix->d_gids;
  a.F = p.F; a.range = (uint32_t)p.range; a.n = ix->n; a.n_stride = ix->n_stride; a.gid_base = ix->gid_base;
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
      cudaError_t e = cudaSuccess;
      if (mode == kPack16) {
        e = cudaFuncSetAttribute(query_count_kernel<kPack16, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        query_count_kernel<kPack16, NT><<<nb, NT, smem, ctx->stream>>>(a, q0);
      } else if (mode == kSmem32) {
        e = cudaFuncSetAttribute(query_count_kernel<kSmem32, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        query_count_kernel<kSmem32, NT><<<nb, NT, smem, ctx->stream>>>(a, q0);
      } else {
        query_count_kernel<kGlobal32, NT><<<nb, NT, 0, ctx->stream>>>(a, q0);
      }
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
