// capi.cu — the C ABI of libniqki_b200.so (include/niqki_b200.h): parameter block, context,
// host-buffer entry points (copies pipelined against the kernels) and thin forwards to the
// device-pointer implementations in sketch.cu / index.cu / query.cu / matrix.cu.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include <thread>

#include "internal.h"
#include "pack.h"

static thread_local char g_err[1024] = "";

int nq_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

extern "C" const char* nq_last_error(void) { return g_err; }
extern "C" const char* nq_version(void) { return "niqki_b200 0.1 (sm_100a)"; }

// ---------------------------------------------------------------- parameters (host scalar code)

// Index::Index — /root/reference/src/niqki_index.cpp:13-29
extern "C" int nq_params_init(nq_params* p, uint32_t K, uint32_t S, uint32_t W, uint32_t H, double min_fract) {
  if (!p) return nq_set_error(NQ_ERR_INVALID, "null params");
  // limits of the reference: 1<<2K in 64 bits, 32-bit fingerprint_range*F
  if (K < 2 || K > 31) return nq_set_error(NQ_ERR_INVALID, "K=%u outside [2,31]", K);
  if (S < 1 || W < 1 || W + S >= 32) return nq_set_error(NQ_ERR_INVALID, "need S>=1, W>=1, W+S<32 (S=%u W=%u)", S, W);
  if (H > W) return nq_set_error(NQ_ERR_INVALID, "H=%u > W=%u", H, W);
  p->K = K; p->S = S; p->W = W; p->H = H;
  p->F = 1u << S;
  p->M = W - H;
  p->min_score = (uint32_t)(min_fract * (double)p->F);
  p->range = (int32_t)(1u << W);
  p->mask_M = (1u << p->M) - 1u;
  p->maxrem = (1u << H) - 1u;
  return NQ_OK;
}

// Index::score_H — src/niqki_index.cpp:142-164 (double arithmetic, evaluated once on the host)
static double score_h(uint32_t W, double x, int h) {
  const double eps = 0.02, m = (double)W - h, two_h = std::pow(2, h);
  auto bound = [&](double base) {
    const double u = ((double)1 - std::pow(base, 1 / x)) * std::pow(2, 64);
    const double i = std::log2(u) + two_h - 64;
    const double j = u * std::pow(2, m - 64 - i + two_h);
    if (u < std::pow(2, 64 - two_h + 1)) return u * std::pow(2, two_h - 64 - ((double)W - h) - 1);
    return i * std::pow(2, m) + j;
  };
  return bound(eps) - bound(1 - eps);
}

// Index::select_best_H — src/niqki_index.cpp:126-138.  Only H and M change; mask_M and maxrem keep
// their construction-time values, exactly like the reference.
extern "C" int nq_params_select_best_H(nq_params* p, double genome_size) {
  if (!p) return nq_set_error(NQ_ERR_INVALID, "null params");
  const double x = genome_size / (double)p->F;
  double best = 0;
  for (uint32_t h = 2; h < 7; ++h) {
    const double s = score_h(p->W, x, (int)h);
    if (s > best) {
      best = s;
      p->H = h;
    }
  }
  p->M = p->W - p->H;
  return NQ_OK;
}

int nq_params_check(const nq_params* p) {
  if (!p) return nq_set_error(NQ_ERR_INVALID, "null params");
  if (p->K < 2 || p->K > 31 || p->S < 1 || p->W < 1 || p->W + p->S >= 32 || p->F != (1u << p->S) ||
      p->range != (int32_t)(1u << p->W) || p->M >= 32)
    return nq_set_error(NQ_ERR_INVALID, "inconsistent parameter block (K=%u S=%u W=%u M=%u F=%u)", p->K, p->S, p->W,
                        p->M, p->F);
  return NQ_OK;
}

// ---------------------------------------------------------------- context

static void nq_timing_resolve(nq_ctx* ctx);

extern "C" int nq_device_count(int* count) {
  if (!count) return nq_set_error(NQ_ERR_INVALID, "null argument");
  *count = 0;
  NQ_CUDA(cudaGetDeviceCount(count));
  return NQ_OK;
}

extern "C" int nq_ctx_create(int device, void* cuda_stream, nq_ctx** out) {
  if (!out) return nq_set_error(NQ_ERR_INVALID, "null argument");
  *out = nullptr;
  int count = 0;
  NQ_CUDA(cudaGetDeviceCount(&count));
  if (count == 0) return nq_set_error(NQ_ERR_CUDA, "no CUDA device: libniqki_b200 has no CPU fallback");
  if (device < 0 || device >= count) return nq_set_error(NQ_ERR_INVALID, "device %d outside [0,%d)", device, count);
  NQ_CUDA(cudaSetDevice(device));
  nq_ctx* ctx = new nq_ctx();
  ctx->device = device;
  if (cuda_stream) {
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);
  } else {
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete ctx;
      return nq_set_error(NQ_ERR_CUDA, "cudaStreamCreate failed");
    }
    ctx->own_stream = true;
  }
  cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
  }
  if (const char* fg = nq_tuning_env("NQ_L2_FETCH")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(fg));
  // keep freed scratch in the pool instead of returning it to the driver after every call
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  *out = ctx;
  return NQ_OK;
}

extern "C" int nq_ctx_destroy(nq_ctx* ctx) {
  if (!ctx) return NQ_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  nq_timing_resolve(ctx);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
  if (ctx->h_ring) cudaFreeHost(ctx->h_ring);
  if (ctx->ring_wrap) cudaEventDestroy(ctx->ring_wrap);
  for (auto& s : ctx->slot) {
    cudaFree(s.d_bases); cudaFree(s.d_sk); cudaFree(s.d_flags);
    cudaFreeHost(s.h_flags);
    cudaFree(s.d_codes); cudaFree(s.d_blk); cudaFree(s.d_pool);
    cudaFreeHost(s.h_codes); cudaFreeHost(s.h_blk); cudaFreeHost(s.h_pool);
    if (s.h2d) cudaEventDestroy(s.h2d);
    if (s.done) cudaEventDestroy(s.done);
    if (s.d2h) cudaEventDestroy(s.d2h);
  }
  delete ctx;
  return NQ_OK;
}

int nq_upload_small(nq_ctx* ctx, void* d_dst, const void* h_src, size_t bytes) {
  if (bytes == 0) return NQ_OK;
  constexpr size_t kRing = 16u << 20;
  if (bytes > kRing / 4) {  // large tables: plain copy (synchronises the stream when the source is pageable)
    NQ_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    NQ_CUDA(cudaStreamSynchronize(ctx->stream));
    return NQ_OK;
  }
  if (!ctx->h_ring) {
    NQ_CUDA(cudaMallocHost((void**)&ctx->h_ring, kRing));
    ctx->ring_cap = kRing;
    NQ_CUDA(cudaEventCreateWithFlags(&ctx->ring_wrap, cudaEventDisableTiming));
    NQ_CUDA(cudaEventRecord(ctx->ring_wrap, ctx->stream));
  }
  const size_t need = (bytes + 255) & ~(size_t)255;
  if (ctx->ring_at + need > ctx->ring_cap) {
    // wrap: everything queued from the ring's previous lap must have been copied before it is overwritten.
    // The half-way mark is awaited, so the host only ever waits for uploads queued half a ring ago.
    NQ_CUDA(cudaEventRecord(ctx->ring_wrap, ctx->stream));
    NQ_CUDA(cudaEventSynchronize(ctx->ring_wrap));
    ctx->ring_at = 0;
  }
  char* slot = ctx->h_ring + ctx->ring_at;
  memcpy(slot, h_src, bytes);
  ctx->ring_at += need;
  NQ_CUDA(cudaMemcpyAsync(d_dst, slot, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return NQ_OK;
}

extern "C" int nq_ctx_sync(nq_ctx* ctx) {
  if (!ctx) return nq_set_error(NQ_ERR_INVALID, "null context");
  NQ_CUDA(cudaStreamSynchronize(ctx->stream));
  return NQ_OK;
}

extern "C" uint64_t nq_ctx_launch_count(const nq_ctx* ctx) { return ctx ? ctx->launches : 0; }

static void nq_timing_resolve(nq_ctx* ctx) {
  for (auto& t : ctx->timed) {
    float ms = 0;
    if (cudaEventSynchronize(t.b) == cudaSuccess && cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) {
      ctx->kind_ms[t.kind] += ms;
      ctx->kind_n[t.kind]++;
    }
    cudaEventDestroy(t.a);
    cudaEventDestroy(t.b);
  }
  ctx->timed.clear();
}

extern "C" int nq_ctx_set_timing(nq_ctx* ctx, int on) {
  if (!ctx) return nq_set_error(NQ_ERR_INVALID, "null context");
  ctx->timing = on != 0;
  return NQ_OK;
}
extern "C" int nq_ctx_timing(nq_ctx* ctx, int kind, double* ms, uint64_t* launches) {
  if (!ctx || kind < 0 || kind >= 8) return nq_set_error(NQ_ERR_INVALID, "bad timing query");
  nq_timing_resolve(ctx);
  if (ms) *ms = ctx->kind_ms[kind];
  if (launches) *launches = ctx->kind_n[kind];
  return NQ_OK;
}
extern "C" int nq_ctx_timing_reset(nq_ctx* ctx) {
  if (!ctx) return nq_set_error(NQ_ERR_INVALID, "null context");
  nq_timing_resolve(ctx);
  for (int k = 0; k < 8; ++k) { ctx->kind_ms[k] = 0; ctx->kind_n[k] = 0; }
  return NQ_OK;
}
extern "C" uint64_t nq_ctx_last_query_gathered(const nq_ctx* ctx) { return ctx ? ctx->last_query_gathered : 0; }
extern "C" uint64_t nq_ctx_h2d_bytes(const nq_ctx* ctx) { return ctx ? ctx->h2d_bytes : 0; }

extern "C" void* nq_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
    nq_set_error(NQ_ERR_CUDA, "cudaMallocHost(%zu) failed", bytes);
    return nullptr;
  }
  return p;
}
extern "C" void nq_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

// ---------------------------------------------------------------- sketching

extern "C" int nq_sketch_batch_device(nq_ctx* ctx, const nq_params* p, const char* d_bases, uint64_t bases_capacity,
                                      const uint64_t* offsets, uint64_t n, int32_t* d_sketches, uint32_t* d_flags) {
  NQ_RANGE();
  if (!ctx || !offsets || (n && (!d_bases || !d_sketches))) return nq_set_error(NQ_ERR_INVALID, "null argument");
  NQ_CUDA(cudaSetDevice(ctx->device));
  return nq_launch_sketch(ctx, p, d_bases, bases_capacity, offsets, n, nullptr, n, d_sketches, d_flags);
}

extern "C" int nq_densify_device(nq_ctx* ctx, const nq_params* p, int32_t* d_sketches, uint64_t n, uint32_t* d_flags) {
  NQ_RANGE();
  if (!ctx || (n && !d_sketches)) return nq_set_error(NQ_ERR_INVALID, "null argument");
  NQ_CUDA(cudaSetDevice(ctx->device));
  return nq_launch_densify(ctx, p, d_sketches, n, d_flags);
}

// Host-buffer form: entries are grouped into device batches; batch i+1's characters are copied
// (copy stream) while batch i is sketched (compute stream), and batch i's sketches travel back
// (or are placed in the caller's device array) while batch i+1 runs.  `rec_entry` (nullable)
// maps records to sketch rows, non-decreasing, so an entry's records are contiguous.
static int sketch_records_impl(nq_ctx* ctx, const nq_params* p, const char* bases, const uint64_t* offsets,
                               uint64_t n_rec, const uint32_t* rec_entry, uint64_t n_entries, int32_t* sketches,
                               uint32_t* flags, bool out_on_device) {
  if (!ctx || !offsets || (n_rec && !bases) || (n_entries && !sketches)) return nq_set_error(NQ_ERR_INVALID, "null argument");
  NQ_TRY(nq_params_check(p));
  NQ_CUDA(cudaSetDevice(ctx->device));
  if (!rec_entry) n_entries = n_rec;
  if (n_entries == 0) return NQ_OK;
  const uint64_t F = p->F;
  // long entries may travel packed (below): smaller batches, so that packing one overlaps the copy of another
  const bool long_entries = n_rec && (offsets[n_rec] - offsets[0]) / n_rec >= 4096;
  const bool packed = ctx->pack_mode == 1 || (ctx->pack_mode != 0 && long_entries);
  const uint64_t max_bases = packed ? 192ull << 20 : 512ull << 20, max_cells = (1ull << 30) / 4;  // per batch; <= 1 GB of sketches out
  // first record of every entry (entries without records are legal: they stay empty and are flagged)
  std::vector<uint64_t> first(n_entries + 1, n_rec);
  if (rec_entry) {
    uint32_t prev = 0;
    for (uint64_t r = n_rec; r-- > 0;) {
      if (rec_entry[r] >= n_entries) return nq_set_error(NQ_ERR_INVALID, "record %llu maps to entry %u >= %llu", (unsigned long long)r, rec_entry[r], (unsigned long long)n_entries);
      first[rec_entry[r]] = r;
    }
    for (uint64_t r = 0; r < n_rec; ++r) {
      if (rec_entry[r] < prev) return nq_set_error(NQ_ERR_INVALID, "rec_entry must be non-decreasing");
      prev = rec_entry[r];
    }
    for (uint64_t e = n_entries; e-- > 0;)
      if (first[e] == n_rec || first[e] > first[e + 1]) first[e] = first[e + 1];
  } else {
    for (uint64_t e = 0; e <= n_entries; ++e) first[e] = e;
  }
  // batch boundaries (in entries)
  std::vector<uint64_t> cut{0};
  {
    uint64_t b0 = 0;
    for (uint64_t e = 0; e < n_entries; ++e) {
      const bool full = (offsets[first[e + 1]] - offsets[first[b0]] > max_bases || (e + 1 - b0) * F > max_cells) && e > b0;
      if (full) {
        cut.push_back(e);
        b0 = e;
      }
    }
    cut.push_back(n_entries);
  }
  uint64_t cap_bases = 0, cap_entries = 0;
  for (size_t b = 0; b + 1 < cut.size(); ++b) {
    cap_bases = std::max(cap_bases, offsets[first[cut[b + 1]]] - offsets[first[cut[b]]]);
    cap_entries = std::max(cap_entries, cut[b + 1] - cut[b]);
  }
  // K1: long entries travel as 2 bits per base, packed on the host (pack.cpp) while earlier batches are on the
  // link and the device.  Short reads keep the character form: their kernel fuses densification, and the
  // per-record seed fix-up of the packer would dominate.  (Sending a share of the batches as characters, so
  // that the link works while the cores pack, was built and measured on a 16-core host: 68-73 Gbases/s
  // against 83 with everything packed — the DMA reads and the packer's reads share the host's memory
  // bandwidth — and removed.)
  const uint64_t cap_words = nq_pack_words(cap_bases), cap_blocks = nq_pack_blocks(cap_bases);
  cap_bases = (cap_bases + 15 + 16) & ~15ull;
  int st = NQ_OK;
  cudaError_t e = cudaSuccess;
  const int nslots = cut.size() > 3 && packed ? 3 : cut.size() > 2 ? 2 : 1;
  for (int i = 0; i < nslots && e == cudaSuccess; ++i) {
    nq_ctx::Slot& s = ctx->slot[i];
    if (!s.h2d) {
      cudaEventCreateWithFlags(&s.h2d, cudaEventDisableTiming);
      cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming);
      cudaEventCreateWithFlags(&s.d2h, cudaEventDisableTiming);
    }
    if (packed && s.cap_words < cap_words) {
      cudaFree(s.d_codes); cudaFree(s.d_blk); cudaFree(s.d_pool);
      cudaFreeHost(s.h_codes); cudaFreeHost(s.h_blk); cudaFreeHost(s.h_pool);
      s.d_codes = s.d_blk = nullptr; s.d_pool = nullptr; s.h_codes = s.h_blk = nullptr; s.h_pool = nullptr; s.cap_words = 0;
      if ((e = cudaMalloc((void**)&s.d_codes, cap_words * 4)) != cudaSuccess || (e = cudaMalloc((void**)&s.d_blk, cap_blocks * 4)) != cudaSuccess ||
          (e = cudaMalloc((void**)&s.d_pool, cap_blocks * 64)) != cudaSuccess || (e = cudaMallocHost((void**)&s.h_codes, cap_words * 4)) != cudaSuccess ||
          (e = cudaMallocHost((void**)&s.h_blk, cap_blocks * 4)) != cudaSuccess || (e = cudaMallocHost((void**)&s.h_pool, cap_blocks * 64)) != cudaSuccess)
        break;
      s.cap_words = cap_words;
    }
    if (!packed && s.cap_bases < cap_bases) {
      cudaFree(s.d_bases); s.d_bases = nullptr; s.cap_bases = 0;
      if ((e = cudaMalloc((void**)&s.d_bases, cap_bases)) != cudaSuccess) break;
      s.cap_bases = cap_bases;
    }
    if (!out_on_device && s.cap_cells < cap_entries * F) {
      cudaFree(s.d_sk); s.d_sk = nullptr; s.cap_cells = 0;
      if ((e = cudaMalloc((void**)&s.d_sk, cap_entries * F * 4)) != cudaSuccess) break;
      s.cap_cells = cap_entries * F;
    }
    if (s.cap_entries < cap_entries) {
      cudaFree(s.d_flags); s.d_flags = nullptr; s.cap_entries = 0;
      cudaFreeHost(s.h_flags); s.h_flags = nullptr;
      if ((e = cudaMalloc((void**)&s.d_flags, cap_entries * 4)) != cudaSuccess ||
          (e = cudaMallocHost((void**)&s.h_flags, cap_entries * 4)) != cudaSuccess)
        break;
      s.cap_entries = cap_entries;
    }
    s.flags_n = 0;
  }
  if (e != cudaSuccess) st = nq_set_error(NQ_ERR_CUDA, "sketch batch allocation failed: %s", cudaGetErrorString(e));
  std::vector<uint64_t> local_off[3];
  std::vector<uint32_t> local_ent[3];
  for (size_t b = 0; st == NQ_OK && b + 1 < cut.size(); ++b) {
    nq_ctx::Slot& s = ctx->slot[b % nslots];
    std::vector<uint64_t>& loff = local_off[b % nslots];
    std::vector<uint32_t>& lent = local_ent[b % nslots];
    const uint64_t e0 = cut[b], e1 = cut[b + 1], nb = e1 - e0;
    const uint64_t r0 = first[e0], r1 = first[e1], nr = r1 - r0, base0 = offsets[r0], nbytes = offsets[r1] - base0;
    if (b >= (size_t)nslots) {
      cudaEventSynchronize(s.d2h);  // the slot's previous results are out
      if (flags && s.flags_n) memcpy(flags + s.flags_e0, s.h_flags, s.flags_n * 4);
      s.flags_n = 0;
    }
    loff.resize(nr + 1);
    for (uint64_t i = 0; i <= nr; ++i) loff[i] = offsets[r0 + i] - base0;
    if (rec_entry) {
      lent.resize(nr);
      for (uint64_t i = 0; i < nr; ++i) lent[i] = rec_entry[r0 + i] - (uint32_t)e0;
    }
    int32_t* d_out = out_on_device ? sketches + e0 * F : s.d_sk;
    const unsigned nt = ctx->host_threads ? ctx->host_threads : std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
    const bool as_packed = packed;
    if (as_packed) {
      if (b >= (size_t)nslots) cudaEventSynchronize(s.h2d);  // the staging buffers have left for the device
      const uint64_t words = nq_pack_words(nbytes), blocks = nq_pack_blocks(nbytes);
      const uint64_t used = nq_pack_host(bases + base0, nbytes, loff.data(), nr, p->K, s.h_codes, s.h_blk, s.h_pool, cap_blocks, nt);
      if (used == ~0ull) { st = nq_set_error(NQ_ERR_INVALID, "packer: mask pool overflow"); break; }
      if ((e = cudaMemcpyAsync(s.d_codes, s.h_codes, words * 4, cudaMemcpyHostToDevice, ctx->copy_stream)) != cudaSuccess ||
          (e = cudaMemcpyAsync(s.d_blk, s.h_blk, blocks * 4, cudaMemcpyHostToDevice, ctx->copy_stream)) != cudaSuccess ||
          (used && (e = cudaMemcpyAsync(s.d_pool, s.h_pool, used * 64, cudaMemcpyHostToDevice, ctx->copy_stream)) != cudaSuccess)) {
        st = nq_set_error(NQ_ERR_CUDA, "H2D copy failed: %s", cudaGetErrorString(e));
        break;
      }
      ctx->h2d_bytes += words * 4 + blocks * 4 + used * 64;
      cudaEventRecord(s.h2d, ctx->copy_stream);
      cudaStreamWaitEvent(ctx->stream, s.h2d, 0);
      st = nq_launch_sketch_packed(ctx, p, s.d_codes, s.d_blk, s.d_pool, loff.data(), nr, rec_entry ? lent.data() : nullptr, nb,
                                   d_out, s.d_flags);
    } else {
      if (b >= (size_t)nslots) cudaEventSynchronize(s.done);  // the slot's character buffer is no longer read
      if (nbytes) e = cudaMemcpyAsync(s.d_bases, bases + base0, nbytes, cudaMemcpyHostToDevice, ctx->copy_stream);
      if (e != cudaSuccess) { st = nq_set_error(NQ_ERR_CUDA, "H2D copy failed: %s", cudaGetErrorString(e)); break; }
      ctx->h2d_bytes += nbytes;
      cudaEventRecord(s.h2d, ctx->copy_stream);
      cudaStreamWaitEvent(ctx->stream, s.h2d, 0);
      st = nq_launch_sketch(ctx, p, s.d_bases, s.cap_bases, loff.data(), nr, rec_entry ? lent.data() : nullptr, nb, d_out,
                            s.d_flags);
    }
    if (st != NQ_OK) break;
    cudaEventRecord(s.done, ctx->stream);
    cudaStreamWaitEvent(ctx->d2h_stream, s.done, 0);
    if (!out_on_device) e = cudaMemcpyAsync(sketches + e0 * F, s.d_sk, nb * F * 4, cudaMemcpyDeviceToHost, ctx->d2h_stream);
    if (e == cudaSuccess && flags) {
      e = cudaMemcpyAsync(s.h_flags, s.d_flags, nb * 4, cudaMemcpyDeviceToHost, ctx->d2h_stream);
      s.flags_e0 = e0; s.flags_n = nb;
    }
    if (e != cudaSuccess) { st = nq_set_error(NQ_ERR_CUDA, "D2H copy failed: %s", cudaGetErrorString(e)); break; }
    cudaEventRecord(s.d2h, ctx->d2h_stream);
  }
  if ((e = cudaStreamSynchronize(ctx->copy_stream)) != cudaSuccess && st == NQ_OK)
    st = nq_set_error(NQ_ERR_CUDA, "sketch batch failed: %s", cudaGetErrorString(e));
  if ((e = cudaStreamSynchronize(ctx->d2h_stream)) != cudaSuccess && st == NQ_OK)
    st = nq_set_error(NQ_ERR_CUDA, "sketch batch failed: %s", cudaGetErrorString(e));
  if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess && st == NQ_OK)
    st = nq_set_error(NQ_ERR_CUDA, "sketch batch failed: %s", cudaGetErrorString(e));
  for (int i = 0; i < nslots; ++i) {
    nq_ctx::Slot& s = ctx->slot[i];
    if (flags && s.flags_n && st == NQ_OK) memcpy(flags + s.flags_e0, s.h_flags, s.flags_n * 4);
    s.flags_n = 0;
  }
  return st;
}

extern "C" int nq_sketch_batch(nq_ctx* ctx, const nq_params* p, const char* bases, const uint64_t* offsets, uint64_t n,
                               int32_t* sketches, uint32_t* flags) {
  NQ_RANGE();
  return sketch_records_impl(ctx, p, bases, offsets, n, nullptr, n, sketches, flags, false);
}

extern "C" int nq_sketch_records(nq_ctx* ctx, const nq_params* p, const char* bases, const uint64_t* rec_offsets,
                                 uint64_t n_records, const uint32_t* rec_entry, uint64_t n_entries, int32_t* sketches,
                                 uint32_t* flags, int sketches_on_device) {
  NQ_RANGE();
  return sketch_records_impl(ctx, p, bases, rec_offsets, n_records, rec_entry, n_entries, sketches, flags,
                             sketches_on_device != 0);
}

// ---------------------------------------------------------------- plain device memory for hosts
// that do not link the CUDA runtime themselves (the C++ CLI host keeps its sketch store in HBM)
extern "C" int nq_device_alloc(nq_ctx* ctx, size_t bytes, void** out) {
  if (!ctx || !out) return nq_set_error(NQ_ERR_INVALID, "null argument");
  NQ_CUDA(cudaSetDevice(ctx->device));
  NQ_CUDA(cudaMalloc(out, bytes ? bytes : 16));
  return NQ_OK;
}
extern "C" int nq_device_free(nq_ctx* ctx, void* p) {
  if (!ctx) return nq_set_error(NQ_ERR_INVALID, "null argument");
  NQ_CUDA(cudaSetDevice(ctx->device));
  NQ_CUDA(cudaStreamSynchronize(ctx->stream));
  NQ_CUDA(cudaFree(p));
  return NQ_OK;
}
extern "C" int nq_device_copy(nq_ctx* ctx, void* dst, const void* src, size_t bytes, int kind) {
  if (!ctx || (bytes && (!dst || !src))) return nq_set_error(NQ_ERR_INVALID, "null argument");
  if (kind < 0 || kind > 2) return nq_set_error(NQ_ERR_INVALID, "copy kind %d (0 H2D, 1 D2H, 2 D2D)", kind);
  NQ_CUDA(cudaSetDevice(ctx->device));
  const cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  NQ_CUDA(cudaMemcpyAsync(dst, src, bytes, k, ctx->stream));
  NQ_CUDA(cudaStreamSynchronize(ctx->stream));
  return NQ_OK;
}

// ---------------------------------------------------------------- index / query / matrix

extern "C" int nq_index_build_device(nq_ctx* ctx, const nq_params* p, const int32_t* d_sketches, uint64_t n,
                                     uint32_t gid_base, nq_index** out) {
  NQ_RANGE();
  if (!ctx || !d_sketches || !out) return nq_set_error(NQ_ERR_INVALID, "null argument");
  NQ_CUDA(cudaSetDevice(ctx->device));
  return nq_index_build_impl(ctx, p, d_sketches, n, gid_base, out);
}

extern "C" int nq_index_build(nq_ctx* ctx, const nq_params* p, const int32_t* sketches, uint64_t n, uint32_t gid_base,
                              nq_index** out) {
  NQ_RANGE();
  if (!ctx || !sketches || !out) return nq_set_error(NQ_ERR_INVALID, "null argument");
  NQ_TRY(nq_params_check(p));
  NQ_CUDA(cudaSetDevice(ctx->device));
  int32_t* d = nullptr;
  const size_t bytes = (size_t)n * p->F * 4;
  NQ_TRY(nq_dmalloc(ctx, (void**)&d, bytes));
  cudaError_t e = cudaMemcpyAsync(d, sketches, bytes, cudaMemcpyHostToDevice, ctx->stream);
  int st = e == cudaSuccess ? nq_index_build_impl(ctx, p, d, n, gid_base, out)
                            : nq_set_error(NQ_ERR_CUDA, "H2D copy failed: %s", cudaGetErrorString(e));
  nq_dfree(ctx, d);
  return st;
}

extern "C" int nq_query_batch_device(nq_index* ix, const int32_t* d_sketches, uint64_t nq, uint32_t min_score,
                                     nq_hits** out) {
  NQ_RANGE();
  if (!ix || (nq && !d_sketches)) return nq_set_error(NQ_ERR_INVALID, "null argument");
  NQ_CUDA(cudaSetDevice(ix->ctx->device));
  return nq_query_impl(ix, d_sketches, nq, min_score, out);
}

extern "C" int nq_query_batch(nq_index* ix, const int32_t* sketches, uint64_t nq, uint32_t min_score, nq_hits** out) {
  NQ_RANGE();
  if (!ix || !out || (nq && !sketches)) return nq_set_error(NQ_ERR_INVALID, "null argument");
  nq_ctx* ctx = ix->ctx;
  NQ_CUDA(cudaSetDevice(ctx->device));
  int32_t* d = nullptr;
  const size_t bytes = (size_t)nq * ix->p.F * 4;
  NQ_TRY(nq_dmalloc(ctx, (void**)&d, bytes));
  cudaError_t e = bytes ? cudaMemcpyAsync(d, sketches, bytes, cudaMemcpyHostToDevice, ctx->stream) : cudaSuccess;
  int st = e == cudaSuccess ? nq_query_impl(ix, d, nq, min_score, out)
                            : nq_set_error(NQ_ERR_CUDA, "H2D copy failed: %s", cudaGetErrorString(e));
  nq_dfree(ctx, d);
  return st;
}

extern "C" uint64_t nq_hits_total(const nq_hits* h) { return h ? h->counts.size() : 0; }
extern "C" const uint64_t* nq_hits_ptr(const nq_hits* h) { return h ? h->ptr.data() : nullptr; }
extern "C" const uint32_t* nq_hits_counts(const nq_hits* h) { return h ? h->counts.data() : nullptr; }
extern "C" const uint32_t* nq_hits_gids(const nq_hits* h) { return h ? h->gids.data() : nullptr; }
extern "C" void nq_hits_free(nq_hits* h) { delete h; }

extern "C" int nq_matrix_rows(nq_index* ix, uint32_t row_begin, uint32_t row_end, int wrap16, uint32_t* counts) {
  NQ_RANGE();
  if (!ix) return nq_set_error(NQ_ERR_INVALID, "null index");
  NQ_CUDA(cudaSetDevice(ix->ctx->device));
  return nq_matrix_impl(ix, row_begin, row_end, wrap16, counts);
}

extern "C" int nq_index_sketches_device(nq_index* ix, uint32_t row_begin, uint32_t row_end, int32_t* d_sketches) {
  NQ_RANGE();
  if (!ix) return nq_set_error(NQ_ERR_INVALID, "null index");
  NQ_CUDA(cudaSetDevice(ix->ctx->device));
  return nq_index_sketches_impl(ix, row_begin, row_end, d_sketches);
}

extern "C" int nq_matrix_tile(nq_index* ix, const int32_t* d_row_sketches, uint32_t nrows, int wrap16, uint32_t* counts) {
  NQ_RANGE();
  if (!ix) return nq_set_error(NQ_ERR_INVALID, "null index");
  NQ_CUDA(cudaSetDevice(ix->ctx->device));
  return nq_matrix_tile_impl(ix, d_row_sketches, nrows, wrap16, counts);
}

extern "C" int nq_device_copy_peer(nq_ctx* dst_ctx, void* dst, nq_ctx* src_ctx, const void* src, size_t bytes) {
  if (!dst_ctx || !src_ctx || (bytes && (!dst || !src))) return nq_set_error(NQ_ERR_INVALID, "null argument");
  if (bytes == 0) return NQ_OK;
  NQ_CUDA(cudaSetDevice(src_ctx->device));
  NQ_CUDA(cudaStreamSynchronize(src_ctx->stream));  // the source is complete
  NQ_CUDA(cudaSetDevice(dst_ctx->device));
  NQ_CUDA(cudaMemcpyPeerAsync(dst, dst_ctx->device, src, src_ctx->device, bytes, dst_ctx->stream));
  NQ_CUDA(cudaStreamSynchronize(dst_ctx->stream));
  return NQ_OK;
}

extern "C" int nq_device_fill(nq_ctx* ctx, void* p, int byte, size_t bytes) {
  if (!ctx || (bytes && !p)) return nq_set_error(NQ_ERR_INVALID, "null argument");
  NQ_CUDA(cudaSetDevice(ctx->device));
  NQ_CUDA(cudaMemsetAsync(p, byte, bytes, ctx->stream));
  return NQ_OK;
}

// host threads of the packer (0 = hardware concurrency, at most 32) and whether host sequences travel
// packed: mode -1 auto (entries of >= 4096 characters on average), 0 never, 1 always
extern "C" int nq_ctx_set_host_packing(nq_ctx* ctx, int mode, unsigned threads) {
  if (!ctx || mode < -1 || mode > 1) return nq_set_error(NQ_ERR_INVALID, "bad packing mode");
  ctx->pack_mode = mode;
  ctx->host_threads = threads;
  return NQ_OK;
}

// ---------------------------------------------------------------- K1 as entry points of its own
extern "C" int nq_pack_sizes(uint64_t nbytes, uint64_t* words, uint64_t* blocks) {
  if (!words || !blocks) return nq_set_error(NQ_ERR_INVALID, "null argument");
  *words = nq_pack_words(nbytes);
  *blocks = nq_pack_blocks(nbytes);
  return NQ_OK;
}

extern "C" int nq_pack_sequences(const char* bases, const uint64_t* rec_offsets, uint64_t n_records, uint32_t K, uint32_t* codes,
                                 uint32_t* blk, uint16_t* pool, uint64_t pool_slots, uint64_t* pool_used, unsigned threads) {
  NQ_RANGE();
  if (!rec_offsets || !codes || !blk || (n_records && !bases)) return nq_set_error(NQ_ERR_INVALID, "null argument");
  if (K < 2 || K > 31) return nq_set_error(NQ_ERR_INVALID, "K=%u outside [2,31]", K);
  const uint64_t nbytes = rec_offsets[n_records] - rec_offsets[0];
  if (threads == 0) threads = std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
  const uint64_t used = nq_pack_host(bases + rec_offsets[0], nbytes, rec_offsets, n_records, K, codes, blk, pool, pool ? pool_slots : 0,
                                     threads);
  if (used == ~0ull) return nq_set_error(NQ_ERR_OVERFLOW, "mask pool too small (%llu slots)", (unsigned long long)pool_slots);
  if (pool_used) *pool_used = used;
  return NQ_OK;
}

extern "C" int nq_sketch_batch_packed_device(nq_ctx* ctx, const nq_params* p, const uint32_t* d_codes, const uint32_t* d_blk,
                                             const uint16_t* d_pool, const uint64_t* offsets, uint64_t n, int32_t* d_sketches,
                                             uint32_t* d_flags) {
  NQ_RANGE();
  if (!ctx || !offsets || (n && (!d_codes || !d_blk || !d_sketches))) return nq_set_error(NQ_ERR_INVALID, "null argument");
  if (n && offsets[0] != 0) return nq_set_error(NQ_ERR_INVALID, "offsets of a packed batch start at 0");
  NQ_CUDA(cudaSetDevice(ctx->device));
  return nq_launch_sketch_packed(ctx, p, d_codes, d_blk, d_pool, offsets, n, nullptr, n, d_sketches, d_flags);
}
