// synth.cu — counter-based synthetic inputs generated directly in HBM.  Bench /
// test support: lets the throughput runs start with the sequences already resident, and lets the
// CPU oracle regenerate the very same bytes (oracle/niqki_oracle.c nqo_synth_*).
#include "device_common.cuh"
#include "internal.h"

namespace nq {

__device__ __forceinline__ uint32_t synth_code(uint64_t seed, uint64_t g, uint64_t i) {
  return (uint32_t)(mix64(seed + g * 0xD1B54A32D192ED03ull + i) >> 62);
}
__device__ __constant__ char kACGT[4] = {'A', 'C', 'G', 'T'};

// entry e = genome first+e; 16 bases per thread, one 16-byte store
__global__ void synth_genomes_kernel(uint64_t seed, uint64_t first, uint64_t n, uint64_t len, uint8_t* out) {
  const uint64_t chunks_per = (len + 15) / 16;
  const uint64_t total = n * chunks_per;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t e = t / chunks_per, i0 = (t % chunks_per) * 16;
    uint8_t* dst = out + e * len + i0;
    for (uint32_t j = 0; j < 16 && i0 + j < len; ++j) dst[j] = kACGT[synth_code(seed, first + e, i0 + j)];
  }
}

__global__ void synth_mutants_kernel(uint64_t seed, const uint64_t* g, const uint64_t* q, const uint64_t* thr,
                                     uint64_t n, uint64_t len, uint8_t* out) {
  const uint64_t total = n * len;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t e = t / len, i = t % len;
    uint32_t b = synth_code(seed, g[e], i);
    const uint64_t r = mix64((seed ^ 0xA5A5A5A5ull) + q[e] * 0x9E3779B97F4A7C15ull + 2 * i + 1);
    if (r < thr[e]) b = (b + 1 + (uint32_t)(mix64(r) % 3)) & 3;
    out[t] = kACGT[b];
  }
}

__global__ void synth_reads_kernel(uint64_t seed, uint64_t first, uint64_t n, uint64_t genome_len, uint32_t read_len,
                                   uint8_t* out) {
  const uint64_t total = n * read_len;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = first + t / read_len;
    const uint32_t i = (uint32_t)(t % read_len);
    const uint64_t start = mix64(r) % (genome_len - read_len);
    out[t] = kACGT[synth_code(seed, r % 64, start + i)];
  }
}

}  // namespace nq

using namespace nq;

extern "C" int nq_synth_genomes_device(nq_ctx* ctx, uint64_t seed, uint64_t first_genome, uint64_t n, uint64_t len,
                                       char* d_out) {
  if (!ctx || !d_out) return nq_set_error(NQ_ERR_INVALID, "null argument");
  if (n == 0 || len == 0) return NQ_OK;
  synth_genomes_kernel<<<ctx->sm_count * 16, 256, 0, ctx->stream>>>(seed, first_genome, n, len, (uint8_t*)d_out);
  NQ_CHECK_LAUNCH(ctx);
  return NQ_OK;
}

extern "C" int nq_synth_mutants_device(nq_ctx* ctx, uint64_t seed, const uint64_t* g, const uint64_t* q,
                                       const uint64_t* thr, uint64_t n, uint64_t len, char* d_out) {
  if (!ctx || !d_out || !g || !q || !thr) return nq_set_error(NQ_ERR_INVALID, "null argument");
  if (n == 0 || len == 0) return NQ_OK;
  uint64_t* d = nullptr;
  NQ_TRY(nq_dmalloc(ctx, (void**)&d, 3 * n * 8));
  NQ_CUDA(cudaMemcpyAsync(d, g, n * 8, cudaMemcpyHostToDevice, ctx->stream));
  NQ_CUDA(cudaMemcpyAsync(d + n, q, n * 8, cudaMemcpyHostToDevice, ctx->stream));
  NQ_CUDA(cudaMemcpyAsync(d + 2 * n, thr, n * 8, cudaMemcpyHostToDevice, ctx->stream));
  synth_mutants_kernel<<<ctx->sm_count * 16, 256, 0, ctx->stream>>>(seed, d, d + n, d + 2 * n, n, len, (uint8_t*)d_out);
  NQ_CHECK_LAUNCH(ctx);
  nq_dfree(ctx, d);
  return NQ_OK;
}

extern "C" int nq_synth_reads_device(nq_ctx* ctx, uint64_t seed, uint64_t first_read, uint64_t n, uint64_t genome_len,
                                     uint32_t read_len, char* d_out) {
  if (!ctx || !d_out) return nq_set_error(NQ_ERR_INVALID, "null argument");
  if (genome_len <= read_len) return nq_set_error(NQ_ERR_INVALID, "genome_len must exceed read_len");
  if (n == 0 || read_len == 0) return NQ_OK;
  synth_reads_kernel<<<ctx->sm_count * 16, 256, 0, ctx->stream>>>(seed, first_read, n, genome_len, read_len, (uint8_t*)d_out);
  NQ_CHECK_LAUNCH(ctx);
  return NQ_OK;
}
