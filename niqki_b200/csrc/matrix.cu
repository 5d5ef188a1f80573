// matrix.cu — K4b: all-vs-all hit counting.  Replaces Index::query_range
// (/root/reference/src/niqki_index.cpp:570-598): for every list L, every member a inside the row
// block and every member b of L: ++counts[a][b].  (The reference indexes its array [b*batch + a];
// the relation is symmetric, only the layout differs.)
//
// counts[a][b] is the number of cells in which a and b sit in the same list, i.e. the number of
// cells where both carry the same valid fingerprint — exactly what query_sketch (:633-687) counts
// when it is handed genome a's own sketch.  So a row block is computed as a batch of queries:
//   1. `extract_rows_kernel` rebuilds the sketches of the block's genomes from the index itself
//      (posting (cell, fp, gid) says sketch[gid][cell] = fp; cells without a posting stay -1), one
//      CTA per cell with the cell's directory row staged in shared memory;
//   2. the query kernels (query.cu) count them with the per-genome counters in shared memory and
//      write every counter out as a dense row instead of thresholding.
// That replaces one L2 atomic per pair-increment (193 G/s measured with an L2-resident tile) by a
// shared-memory atomic and reads a list only for the rows that need it.  The reference's
// counters are uint16_t for every S, so values wrap mod 65536 when S >= 16; wrap16
// reproduces that by masking the 32-bit counts on the way out.
#include <algorithm>
#include <vector>

#include "device_common.cuh"
#include "internal.h"

namespace nq {

// One CTA per cell (grid-stride).  begin[] of the cell's lists goes to shared memory; a posting at
// position pos belongs to the LAST list whose begin is <= pos: the lists of a cell are contiguous
// and ordered by fingerprint, and every builder (cell_build_kernel, cell_sort_kernel, import_t)
// gives an empty list begin == end == the running prefix, so begin[] is non-decreasing and an
// empty list is always followed by an entry with the same begin.
template <typename IT>
__device__ __forceinline__ void load_dir(const void* dir, size_t at, uint32_t& b, uint32_t& e) {
  if (sizeof(IT) == 2) {
    const uint32_t w = static_cast<const uint32_t*>(dir)[at];
    b = w & 0xFFFFu; e = w >> 16;
  } else {
    const uint2 w = static_cast<const uint2*>(dir)[at];
    b = w.x; e = w.y;
  }
}

// STAGED = the begin[] column of the cell's directory row fits shared memory (W <= 15); otherwise the
// search reads the row in place.
template <typename IT, bool STAGED>
__global__ void __launch_bounds__(256) extract_rows_kernel(const void* __restrict__ dir, const IT* __restrict__ gids,
                                                           uint32_t F, uint32_t range, uint32_t row_stride,
                                                           uint32_t gid_stride, uint32_t rb, uint32_t re,
                                                           int32_t* __restrict__ sk) {
  extern __shared__ uint32_t s_begin[];  // [range] when STAGED
  __shared__ uint32_t s_used;
  for (uint32_t cell = blockIdx.x; cell < F; cell += gridDim.x) {
    const size_t row = (size_t)cell * row_stride;
    __syncthreads();  // s_begin / s_used of the previous cell are no longer read
    if (threadIdx.x == 0) load_dir<IT>(dir, row + range - 1, s_used, s_used);  // postings of the cell = end of its last list
    if (STAGED)
      for (uint32_t fp = threadIdx.x; fp < range; fp += blockDim.x) {
        uint32_t b, e;
        load_dir<IT>(dir, row + fp, b, e);
        s_begin[fp] = b;
      }
    __syncthreads();
    const uint32_t total = s_used;
    const IT* g = gids + (size_t)cell * gid_stride;
    for (uint32_t pos = threadIdx.x; pos < total; pos += blockDim.x) {
      const uint32_t gid = g[pos];
      if (gid < rb || gid >= re) continue;
      uint32_t lo = 0, hi = range;  // begin[lo] <= pos holds throughout (begin[0] == 0); answer in [lo, hi)
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        uint32_t b, e;
        if (STAGED) b = s_begin[mid];
        else load_dir<IT>(dir, row + mid, b, e);
        if (b <= pos) lo = mid; else hi = mid;
      }
      sk[(size_t)(gid - rb) * F + cell] = (int32_t)lo;
    }
  }
}

}  // namespace nq

using namespace nq;

// sketches of local genomes [rb, re) rebuilt from the posting lists: d_sk[(g-rb)*F + cell] = fp or -1
static int extract_rows(nq_index* ix, uint32_t rb, uint32_t re, int32_t* d_sk) {
  nq_ctx* ctx = ix->ctx;
  const uint32_t F = ix->p.F, range = (uint32_t)ix->p.range, rows = re - rb;
  const bool staged = (size_t)range * 4 <= 64 * 1024;
  const size_t smem = staged ? (size_t)range * 4 : 0;
  if (smem > 48 * 1024)
    NQ_CUDA(ix->elem == 2 ? cudaFuncSetAttribute(extract_rows_kernel<uint16_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                          : cudaFuncSetAttribute(extract_rows_kernel<uint32_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NQ_CUDA(cudaMemsetAsync(d_sk, 0xFF, (size_t)rows * F * 4, ctx->stream));
  NqTimer timer(ctx, NQK_MATRIX);
  const unsigned grid = (unsigned)std::min<uint64_t>(F, (uint64_t)ctx->sm_count * 16);
#define NQ_EXTRACT(IT, ST)                                                                                          \
  extract_rows_kernel<IT, ST><<<grid, 256, smem, ctx->stream>>>(ix->d_row, static_cast<const IT*>(ix->d_gids), F, range, \
                                                                ix->row_stride, ix->gid_stride, rb, re, d_sk)
  if (ix->elem == 2) { if (staged) NQ_EXTRACT(uint16_t, true); else NQ_EXTRACT(uint16_t, false); }
  else { if (staged) NQ_EXTRACT(uint32_t, true); else NQ_EXTRACT(uint32_t, false); }
#undef NQ_EXTRACT
  NQ_CHECK_LAUNCH(ctx);
  return NQ_OK;
}

int nq_index_sketches_impl(nq_index* ix, uint32_t row_begin, uint32_t row_end, int32_t* d_sketches) {
  if (!ix || (row_begin < row_end && !d_sketches)) return nq_set_error(NQ_ERR_INVALID, "null argument");
  if (row_begin > row_end || row_end > ix->n)
    return nq_set_error(NQ_ERR_INVALID, "row range [%u,%u) outside [0,%u)", row_begin, row_end, ix->n);
  if (row_begin == row_end) return NQ_OK;
  return extract_rows(ix, row_begin, row_end, d_sketches);
}

// counts[r*n + j] = cells in which row sketch r and local genome j carry the same valid fingerprint
// (a dense query): one tile of the genome x genome grid, rows from anywhere, columns = this shard.
int nq_matrix_tile_impl(nq_index* ix, const int32_t* d_rows, uint32_t nrows, int wrap16, uint32_t* h_counts) {
  if (!ix || (nrows && (!d_rows || !h_counts))) return nq_set_error(NQ_ERR_INVALID, "null argument");
  if (nrows == 0) return NQ_OK;
  nq_ctx* ctx = ix->ctx;
  const uint32_t n = ix->n, F = ix->p.F;
  const uint32_t per_pass = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(nrows, (2ull << 30) / ((uint64_t)n * 4)));
  uint32_t* d_counts = nullptr;
  NQ_TRY(nq_dmalloc(ctx, (void**)&d_counts, (size_t)per_pass * n * 4));
  int st = NQ_OK;
  cudaError_t e = cudaSuccess;
  for (uint32_t r0 = 0; r0 < nrows && st == NQ_OK; r0 += per_pass) {
    const uint32_t rows = std::min(per_pass, nrows - r0);
    st = nq_query_dense_impl(ix, d_rows + (size_t)r0 * F, rows, wrap16 ? 0xFFFFu : 0xFFFFFFFFu, d_counts);
    if (st != NQ_OK) break;
    if ((e = cudaMemcpyAsync(h_counts + (size_t)r0 * n, d_counts, (size_t)rows * n * 4, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess)
      break;
  }
  nq_dfree(ctx, d_counts);
  if (e != cudaSuccess) return nq_set_error(NQ_ERR_CUDA, "matrix tile failed: %s", cudaGetErrorString(e));
  return st;
}

int nq_matrix_impl(nq_index* ix, uint32_t row_begin, uint32_t row_end, int wrap16, uint32_t* h_counts) {
  if (!ix || !h_counts) return nq_set_error(NQ_ERR_INVALID, "null argument");
  if (row_begin > row_end || row_end > ix->n)
    return nq_set_error(NQ_ERR_INVALID, "row range [%u,%u) outside [0,%u)", row_begin, row_end, ix->n);
  if (row_begin == row_end) return NQ_OK;
  nq_ctx* ctx = ix->ctx;
  const uint32_t n = ix->n, F = ix->p.F;
  // rows per pass: rebuilt sketches + dense counts of one pass within ~6 GB of scratch
  const uint64_t per_row = ((uint64_t)F + n) * 4;
  const uint32_t rows_per_pass =
      (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(row_end - row_begin, (6ull << 30) / per_row));
  int32_t* d_sk = nullptr;
  uint32_t* d_counts = nullptr;
  NQ_TRY(nq_dmalloc(ctx, (void**)&d_sk, (size_t)rows_per_pass * F * 4));
  int st = nq_dmalloc(ctx, (void**)&d_counts, (size_t)rows_per_pass * n * 4);
  if (st != NQ_OK) { nq_dfree(ctx, d_sk); return st; }
  const uint32_t wrap_mask = wrap16 ? 0xFFFFu : 0xFFFFFFFFu;
  cudaError_t e = cudaSuccess;
  for (uint32_t rb = row_begin; rb < row_end && st == NQ_OK; rb += rows_per_pass) {
    const uint32_t re = std::min(row_end, rb + rows_per_pass);
    const uint32_t rows = re - rb;
    // local ids of the block (postings hold gid - gid_base; matrix rows are local to the shard too)
    if ((st = extract_rows(ix, rb, re, d_sk)) != NQ_OK) break;
    st = nq_query_dense_impl(ix, d_sk, rows, wrap_mask, d_counts);
    if (st != NQ_OK) break;
    if ((e = cudaMemcpyAsync(h_counts + (size_t)(rb - row_begin) * n, d_counts, (size_t)rows * n * 4, cudaMemcpyDeviceToHost,
                             ctx->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess)
      break;
  }
  nq_dfree(ctx, d_sk);
  nq_dfree(ctx, d_counts);
  if (e != cudaSuccess) return nq_set_error(NQ_ERR_CUDA, "matrix rows failed: %s", cudaGetErrorString(e));
  return st;
}
