// matrix.cu — K4b: all-vs-all hit counting straight from the posting lists.  Replaces
// Index::query_range (/root/reference/src/niqki_index.cpp:570-598): for every list L, every member
// a inside the row block and every member b of L: ++counts[a][b].  (The reference indexes its
// array [b*batch + a]; the relation is symmetric, only the layout differs.)  The reference's
// counters are uint16_t for every S, so values wrap mod 65536 when S >= 16; wrap16
// reproduces that after counting in 32 bits.
#include <algorithm>
#include <vector>

#include "device_common.cuh"
#include "internal.h"

namespace nq {

template <typename IT>
__device__ __forceinline__ void load_dir_entry(const void* dir, size_t at, uint32_t& b, uint32_t& e);
template <>
__device__ __forceinline__ void load_dir_entry<uint16_t>(const void* dir, size_t at, uint32_t& b, uint32_t& e) {
  const uint32_t w = static_cast<const uint32_t*>(dir)[at];
  b = w & 0xFFFFu;
  e = w >> 16;
}
template <>
__device__ __forceinline__ void load_dir_entry<uint32_t>(const void* dir, size_t at, uint32_t& b, uint32_t& e) {
  const uint2 w = static_cast<const uint2*>(dir)[at];
  b = w.x;
  e = w.y;
}

// One warp walks 32 consecutive lists of one cell at a time; non-empty lists are then expanded by
// the whole warp: rows (members inside [rb,re)) sequentially, columns across lanes.
template <typename IT>
__global__ void __launch_bounds__(256) matrix_count_kernel(const void* __restrict__ dir, const IT* __restrict__ gids,
                                                           uint32_t F, uint32_t range, uint32_t n, uint32_t row_stride,
                                                           uint32_t gid_stride, uint32_t rb, uint32_t re,
                                                           uint32_t* __restrict__ counts) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint32_t groups_per_cell = (range + 31) / 32;
  const uint64_t ngroups = (uint64_t)F * groups_per_cell;
  for (uint64_t grp = warp; grp < ngroups; grp += nwarps) {
    const uint32_t cell = (uint32_t)(grp / groups_per_cell);
    const uint32_t fp = (uint32_t)(grp % groups_per_cell) * 32 + lane;
    uint32_t b = 0, e = 0;
    if (fp < range) load_dir_entry<IT>(dir, (size_t)cell * row_stride + fp, b, e);
    unsigned live = __ballot_sync(0xFFFFFFFFu, e > b);
    const IT* g = gids + (size_t)cell * gid_stride;
    while (live) {
      const int src = __ffs(live) - 1;
      live &= live - 1;
      const uint32_t lb = __shfl_sync(0xFFFFFFFFu, b, src), le = __shfl_sync(0xFFFFFFFFu, e, src);
      for (uint32_t i = lb; i < le; ++i) {
        const uint32_t a = g[i];
        if (a < rb || a >= re) continue;
        uint32_t* crow = counts + (size_t)(a - rb) * n;
        for (uint32_t j = lb + lane; j < le; j += 32) atomicAdd(&crow[g[j]], 1u);
      }
    }
  }
}

__global__ void wrap16_kernel(uint32_t* counts, size_t cells) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (size_t)gridDim.x * blockDim.x)
    counts[i] &= 0xFFFFu;
}

}  // namespace nq

using namespace nq;

int nq_matrix_impl(nq_index* ix, uint32_t row_begin, uint32_t row_end, int wrap16, uint32_t* h_counts) {
  if (!ix || !h_counts) return nq_set_error(NQ_ERR_INVALID, "null argument");
  if (row_begin > row_end || row_end > ix->n)
    return nq_set_error(NQ_ERR_INVALID, "row range [%u,%u) outside [0,%u)", row_begin, row_end, ix->n);
  nq_ctx* ctx = ix->ctx;
  const uint32_t n = ix->n;
  // row blocks sized so the counters stay resident in the 126 MB L2 while the lists stream by
  const uint32_t rows_per_pass = std::max<uint32_t>(1, std::min<uint32_t>(row_end - row_begin ? row_end - row_begin : 1,
                                                                          (uint32_t)((64ull << 20) / ((uint64_t)n * 4) + 1)));
  uint32_t* d_counts = nullptr;
  NQ_TRY(nq_dmalloc(ctx, (void**)&d_counts, (size_t)rows_per_pass * n * 4));
  for (uint32_t rb = row_begin; rb < row_end; rb += rows_per_pass) {
    const uint32_t re = std::min(row_end, rb + rows_per_pass);
    const size_t cells = (size_t)(re - rb) * n;
    NQ_CUDA(cudaMemsetAsync(d_counts, 0, cells * 4, ctx->stream));
    NqTimer timer(ctx, NQK_MATRIX);
    if (ix->elem == 2)
      matrix_count_kernel<uint16_t><<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(
          ix->d_row, static_cast<const uint16_t*>(ix->d_gids), ix->p.F, (uint32_t)ix->p.range, n, ix->row_stride,
          ix->gid_stride, rb, re, d_counts);
    else
      matrix_count_kernel<uint32_t><<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(
          ix->d_row, static_cast<const uint32_t*>(ix->d_gids), ix->p.F, (uint32_t)ix->p.range, n, ix->row_stride,
          ix->gid_stride, rb, re, d_counts);
    NQ_CHECK_LAUNCH(ctx);
    if (wrap16) {
      wrap16_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(d_counts, cells);
      NQ_CHECK_LAUNCH(ctx);
    }
    NQ_CUDA(cudaMemcpyAsync(h_counts + (size_t)(rb - row_begin) * n, d_counts, cells * 4, cudaMemcpyDeviceToHost, ctx->stream));
    NQ_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  nq_dfree(ctx, d_counts);
  return NQ_OK;
}
