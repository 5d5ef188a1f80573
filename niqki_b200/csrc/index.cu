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
#include <vector>

#include "device_common.cuh"
#include "internal.h"

namespace nq {

constexpr uint16_t kNoPost = 0xFFFFu;

// int32 [n][F] -> u16 [F][n_pad]
__global__ void __launch_bounds__(256) transpose_fp_kernel(const int32_t* __restrict__ sk, uint16_t* __restrict__ fpT,
                                                           uint32_t n, uint32_t F, uint32_t n_pad, uint32_t range) {
  __shared__ uint16_t tile[32][34];
  const uint32_t c0 = blockIdx.x * 32, g0 = blockIdx.y * 32;
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

}  // namespace nq

using namespace nq;

template <typename IT>
static cudaError_t launch_cell_sort(nq_ctx* ctx, const uint16_t* d_fpT, uint32_t n, uint32_t n_pad, nq_index* ix,
                                    unsigned long long* d_total) {
  const uint32_t range = (uint32_t)ix->p.range, F = ix->p.F;
  const size_t per_warp = (size_t)range * sizeof(IT);
  const uint32_t nw = 8;
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

static void set_layout(nq_index* ix) {
  ix->elem = ix->n <= kMaxCompact ? 2u : 4u;
  const uint32_t per32 = 32 / ix->elem;
  ix->row_stride = (uint32_t)ix->p.range;  // directory entries per cell (2*elem bytes each)
  ix->gid_stride = (std::max<uint32_t>(ix->n, 1) + per32 - 1) / per32 * per32;
}

int nq_index_build_impl(nq_ctx* ctx, const nq_params* p, const int32_t* d_sketches, uint64_t n64,
                        uint32_t gid_base, nq_index** out) {
  NQ_TRY(nq_params_check(p));
  if (p->W > 15) return nq_set_error(NQ_ERR_UNSUPPORTED, "index build supports W <= 15 (got W=%u)", p->W);
  if (n64 == 0 || n64 > 0xFFFFFFF0ull || n64 + gid_base > 0xFFFFFFFFull)
    return nq_set_error(NQ_ERR_INVALID, "bad genome count %llu (gid_base %u)", (unsigned long long)n64, gid_base);
  const uint32_t n = (uint32_t)n64, F = p->F, range = (uint32_t)p->range;
  if ((size_t)range * 4 * 8 > ctx->smem_optin)
    return nq_set_error(NQ_ERR_UNSUPPORTED, "2^W counters do not fit in shared memory");

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

  dim3 tg((F + 31) / 32, (n_pad + 31) / 32);
  if (tg.y > 65535) return fail(nq_set_error(NQ_ERR_UNSUPPORTED, "more than 2M genomes per build call"));
  {
    NqTimer timer(ctx, NQK_TRANSPOSE);
    transpose_fp_kernel<<<tg, 256, 0, ctx->stream>>>(d_sketches, d_fpT, n, F, n_pad, range);
  }
  ctx->launches++;
  e = ix->elem == 2 ? launch_cell_sort<uint16_t>(ctx, d_fpT, n, n_pad, ix, d_total)
                    : launch_cell_sort<uint32_t>(ctx, d_fpT, n, n_pad, ix, d_total);
  if (e != cudaSuccess) return fail(nq_set_error(NQ_ERR_CUDA, "cell_sort launch failed: %s", cudaGetErrorString(e)));
  unsigned long long total = 0;
  if ((e = cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess ||
      (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess)
    return fail(nq_set_error(NQ_ERR_CUDA, "index build failed: %s", cudaGetErrorString(e)));
  ix->n_postings = total;
  nq_dfree(ctx, d_fpT);
  nq_dfree(ctx, d_total);
  d_fpT = nullptr; d_total = nullptr;
  if ((st = nq_query_prepare(ix)) != NQ_OK) return fail(st);
  *out = ix;
  return NQ_OK;
}

extern "C" int nq_index_free(nq_index* ix) {
  if (!ix) return NQ_OK;
  if (ix->ctx) {
    cudaSetDevice(ix->ctx->device);
    nq_dfree(ix->ctx, ix->d_row);
    nq_dfree(ix->ctx, ix->d_gids);
    nq_dfree(ix->ctx, ix->d_pool);
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
  if (device_bytes) *device_bytes = (uint64_t)ix->p.F * (2ull * ix->row_stride + ix->gid_stride) * ix->elem;
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
  return nq_query_prepare(ix);
}

extern "C" int nq_index_import(nq_ctx* ctx, const nq_params* p, const uint32_t* list_sizes, const uint32_t* gids,
                               uint32_t n_genomes, uint32_t gid_base, nq_index** out) {
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
