// index.cu — K3: inverted-index build in HBM.  Replaces Index::insert_sketch and the
// vector<gid>[2^(W+S)] storage (/root/reference/src/niqki_index.cpp:362-370, niqki_index.h:55-56).
//
// A posting's list id is fp + cell*2^W and its cell is its position in the sketch, so the build is
// not a general sort: for every cell it is a stable counting sort of the n genomes by their W-bit
// fingerprint.  Layout in HBM (fixed cell stride, no global prefix needed):
//   row [F][range+1] u32 : row[cell][fp] .. row[cell][fp+1] delimit list (cell,fp) inside the cell
//   gids[F][n_stride] u32: the cell's genomes ordered by (fp, gid)  — gid-ascending inside a list,
//                          i.e. the reference's push_back order at OMP_NUM_THREADS=1.
// Step 1 transposes the int32 sketches [n][F] into u16 fingerprints [F][n_pad] (0xFFFF = not
// posted: the reference only posts 0 <= fp < range, :364) so that step 2 reads each cell's column
// coalesced.  Step 2 gives one warp per cell: histogram in shared memory, warp scan -> row[],
// then a second sweep that ranks equal fingerprints inside each 32-genome chunk with
// __match_any_sync so the scatter is stable.
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

// one warp per cell: stable counting sort of the cell's genomes by fingerprint
__global__ void cell_sort_kernel(const uint16_t* __restrict__ fpT, uint32_t n, uint32_t n_pad, uint32_t range,
                                 uint32_t F, uint32_t gid_base, uint32_t* __restrict__ row,
                                 uint32_t* __restrict__ gids, uint32_t n_stride,
                                 unsigned long long* __restrict__ total_postings) {
  extern __shared__ uint32_t smem[];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const uint32_t cell = blockIdx.x * nw + warp;
  if (cell >= F) return;
  uint32_t* cnt = smem + (size_t)warp * range;
  const uint16_t* col = fpT + (size_t)cell * n_pad;
  constexpr unsigned kFull = 0xFFFFFFFFu;

  for (uint32_t i = lane; i < range; i += 32) cnt[i] = 0;
  __syncwarp();
  for (uint32_t g0 = 0; g0 < n; g0 += 32) {
    const uint32_t g = g0 + lane;
    const uint16_t fp = g < n ? col[g] : kNoPost;
    if (fp != kNoPost) atomicAdd(&cnt[fp], 1u);
  }
  __syncwarp();

  // exclusive scan of the histogram, 32 bins per round; cnt[] becomes the write cursor
  uint32_t* myrow = row + (size_t)cell * (range + 1);
  uint32_t carry = 0;
  for (uint32_t b0 = 0; b0 < range; b0 += 32) {
    const uint32_t bin = b0 + lane;
    const uint32_t c = bin < range ? cnt[bin] : 0;
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(kFull, incl, d);
      if (lane >= d) incl += t;
    }
    const uint32_t excl = carry + incl - c;
    if (bin < range) {
      cnt[bin] = excl;
      myrow[bin] = excl;
    }
    carry += __shfl_sync(kFull, incl, 31);
  }
  if (lane == 0) {
    myrow[range] = carry;
    atomicAdd(total_postings, (unsigned long long)carry);
  }
  __syncwarp();

  uint32_t* out = gids + (size_t)cell * n_stride;
  const unsigned lt = (1u << lane) - 1;
  for (uint32_t g0 = 0; g0 < n; g0 += 32) {
    const uint32_t g = g0 + lane;
    const uint16_t fp = g < n ? col[g] : kNoPost;
    const bool valid = fp != kNoPost;
    const unsigned same = __match_any_sync(kFull, fp);
    const uint32_t rank = __popc(same & lt);
    const uint32_t base = valid ? cnt[fp] : 0;
    __syncwarp();
    if (valid && rank == 0) cnt[fp] = base + __popc(same);
    __syncwarp();
    if (valid) out[base + rank] = gid_base + g;
  }
}

}  // namespace nq

using namespace nq;

int nq_index_build_impl(nq_ctx* ctx, const nq_params* p, const int32_t* d_sketches, uint64_t n64,
                        uint32_t gid_base, nq_index** out) {
  NQ_TRY(nq_params_check(p));
  if (p->W > 15) return nq_set_error(NQ_ERR_UNSUPPORTED, "index build supports W <= 15 (got W=%u)", p->W);
  if (n64 == 0 || n64 > 0xFFFFFFF0ull || n64 + gid_base > 0xFFFFFFFFull)
    return nq_set_error(NQ_ERR_INVALID, "bad genome count %llu (gid_base %u)", (unsigned long long)n64, gid_base);
  const uint32_t n = (uint32_t)n64, F = p->F, range = (uint32_t)p->range;
  const size_t per_warp = (size_t)range * 4;
  if (per_warp > ctx->smem_optin) return nq_set_error(NQ_ERR_UNSUPPORTED, "2^W counters do not fit in shared memory");

  nq_index* ix = new nq_index();
  ix->ctx = ctx; ix->p = *p; ix->n = n; ix->gid_base = gid_base; ix->n_stride = n;
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
  if ((st = nq_dmalloc(ctx, (void**)&ix->d_row, (size_t)F * (range + 1) * 4)) != NQ_OK ||
      (st = nq_dmalloc(ctx, (void**)&ix->d_gids, (size_t)F * n * 4)) != NQ_OK)
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

  uint32_t nw = (uint32_t)std::min<size_t>(8, ctx->smem_optin / per_warp);
  while (nw > 1 && (F % nw)) --nw;
  const size_t smem = per_warp * nw;
  if (cudaFuncSetAttribute(cell_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return fail(nq_set_error(NQ_ERR_CUDA, "cudaFuncSetAttribute(cell_sort_kernel) failed"));
  {
    NqTimer timer(ctx, NQK_CELLSORT);
    cell_sort_kernel<<<(F + nw - 1) / nw, nw * 32, smem, ctx->stream>>>(d_fpT, n, n_pad, range, F, gid_base, ix->d_row,
                                                                        ix->d_gids, ix->n_stride, d_total);
  }
  ctx->launches++;
  unsigned long long total = 0;
  if ((e = cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess ||
      (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess)
    return fail(nq_set_error(NQ_ERR_CUDA, "index build failed: %s", cudaGetErrorString(e)));
  ix->n_postings = total;
  nq_dfree(ctx, d_fpT);
  nq_dfree(ctx, d_total);
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
  if (device_bytes)
    *device_bytes = (uint64_t)ix->p.F * ((uint64_t)ix->p.range + 1) * 4 + (uint64_t)ix->p.F * ix->n_stride * 4;
  return NQ_OK;
}

extern "C" int nq_index_export(nq_index* ix, uint32_t* list_sizes, uint32_t* gids, uint64_t gids_capacity) {
  if (!ix) return nq_set_error(NQ_ERR_INVALID, "null index");
  if (gids && gids_capacity < ix->n_postings)
    return nq_set_error(NQ_ERR_OVERFLOW, "gids capacity %llu < %llu postings", (unsigned long long)gids_capacity,
                        (unsigned long long)ix->n_postings);
  nq_ctx* ctx = ix->ctx;
  const uint32_t F = ix->p.F, range = (uint32_t)ix->p.range;
  // stream cells through a bounded host staging buffer
  const uint32_t cells_per_chunk = std::max<uint32_t>(1, (64u << 20) / std::max<uint32_t>(1, (range + 1 + ix->n_stride) * 4));
  std::vector<uint32_t> hrow((size_t)cells_per_chunk * (range + 1)), hg((size_t)cells_per_chunk * ix->n_stride);
  uint64_t w = 0;
  for (uint32_t c0 = 0; c0 < F; c0 += cells_per_chunk) {
    const uint32_t nc = std::min(cells_per_chunk, F - c0);
    NQ_CUDA(cudaMemcpyAsync(hrow.data(), ix->d_row + (size_t)c0 * (range + 1), (size_t)nc * (range + 1) * 4,
                            cudaMemcpyDeviceToHost, ctx->stream));
    if (gids)
      NQ_CUDA(cudaMemcpyAsync(hg.data(), ix->d_gids + (size_t)c0 * ix->n_stride, (size_t)nc * ix->n_stride * 4,
                              cudaMemcpyDeviceToHost, ctx->stream));
    NQ_CUDA(cudaStreamSynchronize(ctx->stream));
    for (uint32_t c = 0; c < nc; ++c) {
      const uint32_t* r = &hrow[(size_t)c * (range + 1)];
      if (list_sizes)
        for (uint32_t f = 0; f < range; ++f) list_sizes[(size_t)(c0 + c) * range + f] = r[f + 1] - r[f];
      if (gids) {
        std::copy(&hg[(size_t)c * ix->n_stride], &hg[(size_t)c * ix->n_stride] + r[range], gids + w);
        w += r[range];
      }
    }
  }
  return NQ_OK;
}

extern "C" int nq_index_import(nq_ctx* ctx, const nq_params* p, const uint32_t* list_sizes, const uint32_t* gids,
                               uint32_t n_genomes, uint32_t gid_base, nq_index** out) {
  NQ_TRY(nq_params_check(p));
  if (!ctx || !list_sizes || !gids || !out) return nq_set_error(NQ_ERR_INVALID, "null argument");
  const uint32_t F = p->F, range = (uint32_t)p->range;
  // pass 1: per-cell totals of the gids that belong to this shard
  std::vector<uint32_t> hrow((size_t)F * (range + 1));
  uint64_t r = 0, total = 0;
  uint32_t stride = 1;
  for (uint32_t c = 0; c < F; ++c) {
    uint32_t run = 0;
    for (uint32_t f = 0; f < range; ++f) {
      hrow[(size_t)c * (range + 1) + f] = run;
      const uint32_t sz = list_sizes[(size_t)c * range + f];
      for (uint32_t j = 0; j < sz; ++j) {
        const uint32_t g = gids[r + j];
        if (g >= gid_base && g - gid_base < n_genomes) ++run;
      }
      r += sz;
    }
    hrow[(size_t)c * (range + 1) + range] = run;
    stride = std::max(stride, run);
    total += run;
  }
  std::vector<uint32_t> hg((size_t)F * stride, 0);
  r = 0;
  for (uint32_t c = 0; c < F; ++c) {
    uint32_t w = 0;
    for (uint32_t f = 0; f < range; ++f) {
      const uint32_t sz = list_sizes[(size_t)c * range + f];
      for (uint32_t j = 0; j < sz; ++j) {
        const uint32_t g = gids[r + j];
        if (g >= gid_base && g - gid_base < n_genomes) hg[(size_t)c * stride + w++] = g;
      }
      r += sz;
    }
  }
  nq_index* ix = new nq_index();
  ix->ctx = ctx; ix->p = *p; ix->n = n_genomes; ix->gid_base = gid_base; ix->n_stride = stride; ix->n_postings = total;
  cudaError_t e = cudaSuccess;
  if (nq_dmalloc(ctx, (void**)&ix->d_row, hrow.size() * 4) != NQ_OK ||
      nq_dmalloc(ctx, (void**)&ix->d_gids, hg.size() * 4) != NQ_OK ||
      (e = cudaMemcpyAsync(ix->d_row, hrow.data(), hrow.size() * 4, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
      (e = cudaMemcpyAsync(ix->d_gids, hg.data(), hg.size() * 4, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
      (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) {
    nq_index_free(ix);
    return nq_set_error(NQ_ERR_CUDA, "index import failed: %s", cudaGetErrorString(e));
  }
  *out = ix;
  return NQ_OK;
}
