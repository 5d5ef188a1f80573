// comm.cu — the one exchange step of the sharded index: NCCL over NVLink behind
// the C ABI.  The index shards by genome id with no communication; queries are sketched where
// their sequences are and ALL-GATHERED so that every shard counts every query; --matrix rows are
// BROADCAST from the shard that owns them.  Per-shard hit lists are disjoint in gid, so the merge
// is concatenate + sort (nq_hits_merge) and equals std::sort(greater) of the reference
// (/root/reference/src/niqki_index.cpp:685).
//
// NCCL is resolved at run time (dlopen of libnccl.so.2: the system copy, or the one a host process
// such as PyTorch has already loaded), so the library has no link-time dependency on it and a
// single-GPU user never touches it.
#include <dlfcn.h>

#include <algorithm>
#include <mutex>

#include "device_common.cuh"
#include "internal.h"

namespace {

// minimal NCCL surface (nccl.h: ncclUniqueId is 128 opaque bytes passed by value; ncclUint8 == 1)
struct nq_nccl_id { char internal[128]; };
typedef struct ncclComm* nq_nccl_comm;
constexpr int kNcclUint8 = 1;

struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(nq_nccl_id*) = nullptr;
  int (*CommInitRank)(nq_nccl_comm*, int, nq_nccl_id, int) = nullptr;
  int (*CommInitAll)(nq_nccl_comm*, int, const int*) = nullptr;
  int (*CommDestroy)(nq_nccl_comm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, nq_nccl_comm, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, nq_nccl_comm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};

NcclApi g_nccl;
std::once_flag g_nccl_once;
const char* g_nccl_err = nullptr;

void nccl_load() {
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names)
    if ((g_nccl.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break;
  if (!g_nccl.h) {
    g_nccl_err = "libnccl.so.2 not found";
    return;
  }
#define NQ_SYM(field, name)                                                    \
  *reinterpret_cast<void**>(&g_nccl.field) = dlsym(g_nccl.h, name);            \
  if (!g_nccl.field) { g_nccl_err = "NCCL symbol missing: " name; return; }
  NQ_SYM(GetUniqueId, "ncclGetUniqueId")
  NQ_SYM(CommInitRank, "ncclCommInitRank")
  NQ_SYM(CommInitAll, "ncclCommInitAll")
  NQ_SYM(CommDestroy, "ncclCommDestroy")
  NQ_SYM(AllGather, "ncclAllGather")
  NQ_SYM(Broadcast, "ncclBroadcast")
  NQ_SYM(GroupStart, "ncclGroupStart")
  NQ_SYM(GroupEnd, "ncclGroupEnd")
  NQ_SYM(GetErrorString, "ncclGetErrorString")
  NQ_SYM(GetVersion, "ncclGetVersion")
#undef NQ_SYM
}

int nccl_ready() {
  std::call_once(g_nccl_once, nccl_load);
  if (g_nccl_err) return nq_set_error(NQ_ERR_UNSUPPORTED, "NCCL unavailable: %s", g_nccl_err);
  return NQ_OK;
}

#define NQ_NCCL(expr)                                                                                       \
  do {                                                                                                      \
    const int _r = (expr);                                                                                  \
    if (_r != 0) return nq_set_error(NQ_ERR_CUDA, "%s failed: %s", #expr, g_nccl.GetErrorString(_r));      \
  } while (0)

}  // namespace

struct nq_comm {
  nq_ctx* ctx = nullptr;
  nq_nccl_comm comm = nullptr;
  int rank = 0, nranks = 1;
  uint16_t* d_wire = nullptr;  // u16 staging: [1 + nranks] blocks of wire_cap cells
  size_t wire_cap = 0;
};

namespace nq {

// sketches on the wire: a query only ever uses cells with 0 <= fp < 2^W (:655), so with W <= 15 a
// cell travels as u16 and everything else (empty, out of range) as 0xFFFF — half the bytes
__global__ void sketch_pack16_kernel(const int32_t* __restrict__ in, uint16_t* __restrict__ out, size_t cells, uint32_t range) {
  const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < cells; i += stride) {
    if (i + 4 <= cells) {
      const int4 v = *reinterpret_cast<const int4*>(in + i);
      const uint32_t a = (uint32_t)v.x < range ? (uint32_t)v.x : 0xFFFFu, b = (uint32_t)v.y < range ? (uint32_t)v.y : 0xFFFFu;
      const uint32_t c = (uint32_t)v.z < range ? (uint32_t)v.z : 0xFFFFu, d = (uint32_t)v.w < range ? (uint32_t)v.w : 0xFFFFu;
      *reinterpret_cast<uint2*>(out + i) = make_uint2(a | (b << 16), c | (d << 16));
    } else {
      for (size_t j = i; j < cells; ++j) out[j] = (uint32_t)in[j] < range ? (uint16_t)in[j] : (uint16_t)0xFFFFu;
    }
  }
}
__global__ void sketch_unpack16_kernel(const uint16_t* __restrict__ in, int32_t* __restrict__ out, size_t cells) {
  const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < cells; i += stride) {
    if (i + 4 <= cells) {
      const uint2 v = *reinterpret_cast<const uint2*>(in + i);
      const uint32_t w[4] = {v.x & 0xFFFFu, v.x >> 16, v.y & 0xFFFFu, v.y >> 16};
      *reinterpret_cast<int4*>(out + i) = make_int4(w[0] == 0xFFFFu ? -1 : (int)w[0], w[1] == 0xFFFFu ? -1 : (int)w[1],
                                                    w[2] == 0xFFFFu ? -1 : (int)w[2], w[3] == 0xFFFFu ? -1 : (int)w[3]);
    } else {
      for (size_t j = i; j < cells; ++j) out[j] = in[j] == 0xFFFFu ? -1 : (int)in[j];
    }
  }
}

}  // namespace nq

extern "C" int nq_comm_unique_id(void* id128) {
  if (!id128) return nq_set_error(NQ_ERR_INVALID, "null argument");
  NQ_TRY(nccl_ready());
  NQ_NCCL(g_nccl.GetUniqueId(static_cast<nq_nccl_id*>(id128)));
  return NQ_OK;
}

extern "C" int nq_comm_init_rank(nq_ctx* ctx, const void* id128, int nranks, int rank, nq_comm** out) {
  if (!ctx || !id128 || !out || nranks < 1 || rank < 0 || rank >= nranks) return nq_set_error(NQ_ERR_INVALID, "bad communicator arguments");
  *out = nullptr;
  NQ_TRY(nccl_ready());
  NQ_CUDA(cudaSetDevice(ctx->device));
  nq_comm* c = new nq_comm();
  c->ctx = ctx; c->rank = rank; c->nranks = nranks;
  nq_nccl_id id;
  memcpy(&id, id128, sizeof id);
  const int r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
  if (r != 0) {
    delete c;
    return nq_set_error(NQ_ERR_CUDA, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
  }
  *out = c;
  return NQ_OK;
}

extern "C" int nq_comm_init_all(nq_ctx* const* ctxs, int n, nq_comm** out) {
  if (!ctxs || !out || n < 1 || n > 64) return nq_set_error(NQ_ERR_INVALID, "bad communicator arguments");
  NQ_TRY(nccl_ready());
  int devs[64];
  nq_nccl_comm comms[64];
  for (int i = 0; i < n; ++i) {
    if (!ctxs[i]) return nq_set_error(NQ_ERR_INVALID, "null context");
    devs[i] = ctxs[i]->device;
    out[i] = nullptr;
  }
  NQ_NCCL(g_nccl.CommInitAll(comms, n, devs));
  for (int i = 0; i < n; ++i) {
    nq_comm* c = new nq_comm();
    c->ctx = ctxs[i]; c->comm = comms[i]; c->rank = i; c->nranks = n;
    out[i] = c;
  }
  return NQ_OK;
}

extern "C" int nq_comm_destroy(nq_comm* c) {
  if (!c) return NQ_OK;
  if (c->ctx) {
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    if (c->d_wire) cudaFree(c->d_wire);
  }
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  delete c;
  return NQ_OK;
}

extern "C" int nq_comm_info(const nq_comm* c, int* rank, int* nranks, int* nccl_version) {
  if (!c) return nq_set_error(NQ_ERR_INVALID, "null communicator");
  if (rank) *rank = c->rank;
  if (nranks) *nranks = c->nranks;
  if (nccl_version) {
    *nccl_version = 0;
    if (g_nccl.GetVersion) g_nccl.GetVersion(nccl_version);
  }
  return NQ_OK;
}

// contiguous gid blocks: rank r of R owns [r*ceil(n/R), min(n, (r+1)*ceil(n/R)))
extern "C" int nq_shard_range(uint64_t n, int nranks, int rank, uint64_t* begin, uint64_t* end) {
  if (nranks < 1 || rank < 0 || rank >= nranks || !begin || !end) return nq_set_error(NQ_ERR_INVALID, "bad shard arguments");
  const uint64_t per = (n + (uint64_t)nranks - 1) / (uint64_t)nranks;
  *begin = std::min(n, per * (uint64_t)rank);
  *end = std::min(n, *begin + per);
  return NQ_OK;
}

static int wire_reserve(nq_comm* c, size_t cells) {
  if (c->wire_cap >= cells) return NQ_OK;
  NQ_CUDA(cudaStreamSynchronize(c->ctx->stream));
  if (c->d_wire) cudaFree(c->d_wire);
  c->d_wire = nullptr; c->wire_cap = 0;
  NQ_CUDA(cudaMalloc((void**)&c->d_wire, (size_t)(1 + c->nranks) * cells * sizeof(uint16_t)));
  c->wire_cap = cells;
  return NQ_OK;
}

extern "C" int nq_allgather_sketches(nq_comm* c, const nq_params* p, const int32_t* d_local, uint64_t n_local, int32_t* d_all) {
  NQ_RANGE();
  if (!c || !p || (n_local && (!d_local || !d_all))) return nq_set_error(NQ_ERR_INVALID, "null argument");
  if (n_local == 0) return NQ_OK;
  nq_ctx* ctx = c->ctx;
  NQ_CUDA(cudaSetDevice(ctx->device));
  const size_t cells = (size_t)n_local * p->F;
  if (c->nranks == 1) {
    if (d_all != d_local) NQ_CUDA(cudaMemcpyAsync(d_all, d_local, cells * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    return NQ_OK;
  }
  if (p->W <= 15) {
    NQ_TRY(wire_reserve(c, cells));
    uint16_t* mine = c->d_wire;
    uint16_t* all = c->d_wire + c->wire_cap;
    const unsigned grid = (unsigned)std::min<size_t>((cells / 4 + 255) / 256 + 1, (size_t)ctx->sm_count * 16);
    nq::sketch_pack16_kernel<<<grid, 256, 0, ctx->stream>>>(d_local, mine, cells, (uint32_t)p->range);
    NQ_CHECK_LAUNCH(ctx);
    NQ_NCCL(g_nccl.AllGather(mine, all, cells * 2, kNcclUint8, c->comm, ctx->stream));
    const size_t total = cells * (size_t)c->nranks;
    const unsigned grid2 = (unsigned)std::min<size_t>((total / 4 + 255) / 256 + 1, (size_t)ctx->sm_count * 16);
    nq::sketch_unpack16_kernel<<<grid2, 256, 0, ctx->stream>>>(all, d_all, total);  // NCCL wrote the blocks back to back
    NQ_CHECK_LAUNCH(ctx);
  } else {
    NQ_NCCL(g_nccl.AllGather(d_local, d_all, cells * 4, kNcclUint8, c->comm, ctx->stream));
  }
  return NQ_OK;
}

extern "C" int nq_bcast_sketches(nq_comm* c, const nq_params* p, int32_t* d_sketches, uint64_t n, int root) {
  NQ_RANGE();
  if (!c || !p || (n && !d_sketches) || root < 0 || root >= c->nranks) return nq_set_error(NQ_ERR_INVALID, "bad broadcast arguments");
  if (n == 0 || c->nranks == 1) return NQ_OK;
  nq_ctx* ctx = c->ctx;
  NQ_CUDA(cudaSetDevice(ctx->device));
  const size_t cells = (size_t)n * p->F;
  if (p->W <= 15) {
    NQ_TRY(wire_reserve(c, cells));
    const unsigned grid = (unsigned)std::min<size_t>((cells / 4 + 255) / 256 + 1, (size_t)ctx->sm_count * 16);
    if (c->rank == root) {
      nq::sketch_pack16_kernel<<<grid, 256, 0, ctx->stream>>>(d_sketches, c->d_wire, cells, (uint32_t)p->range);
      NQ_CHECK_LAUNCH(ctx);
    }
    NQ_NCCL(g_nccl.Broadcast(c->d_wire, c->d_wire, cells * 2, kNcclUint8, root, c->comm, ctx->stream));
    if (c->rank != root) {
      nq::sketch_unpack16_kernel<<<grid, 256, 0, ctx->stream>>>(c->d_wire, d_sketches, cells);
      NQ_CHECK_LAUNCH(ctx);
    }
  } else {
    NQ_NCCL(g_nccl.Broadcast(d_sketches, d_sketches, cells * 4, kNcclUint8, root, c->comm, ctx->stream));
  }
  return NQ_OK;
}

// Host merge of per-shard results of the SAME query batch (:685 ordering).
extern "C" int nq_hits_merge(const nq_hits* const* parts, int nparts, nq_hits** out) {
  NQ_RANGE();
  if (!parts || !out || nparts < 1) return nq_set_error(NQ_ERR_INVALID, "bad merge arguments");
  for (int s = 0; s < nparts; ++s)
    if (!parts[s] || parts[s]->ptr.size() != parts[0]->ptr.size())
      return nq_set_error(NQ_ERR_INVALID, "hit lists of different query batches");
  const size_t nq = parts[0]->ptr.size() - 1;
  nq_hits* h = new nq_hits();
  h->ptr.assign(nq + 1, 0);
  size_t total = 0;
  for (int s = 0; s < nparts; ++s) total += parts[s]->counts.size();
  h->counts.resize(total);
  h->gids.resize(total);
  std::vector<uint64_t> keys;
  size_t w = 0;
  for (size_t q = 0; q < nq; ++q) {
    keys.clear();
    for (int s = 0; s < nparts; ++s)
      for (uint64_t i = parts[s]->ptr[q]; i < parts[s]->ptr[q + 1]; ++i)
        keys.push_back(((uint64_t)parts[s]->counts[i] << 32) | parts[s]->gids[i]);
    std::sort(keys.begin(), keys.end(), std::greater<uint64_t>());
    h->ptr[q] = w;
    for (uint64_t k : keys) {
      h->counts[w] = (uint32_t)(k >> 32);
      h->gids[w] = (uint32_t)k;
      ++w;
    }
  }
  h->ptr[nq] = w;
  *out = h;
  return NQ_OK;
}

// The same from raw arrays (a host that gathered the per-rank results over its own transport)
extern "C" int nq_hits_from_arrays(const uint64_t* ptr, const uint32_t* counts, const uint32_t* gids, uint64_t nq, nq_hits** out) {
  if (!ptr || !out) return nq_set_error(NQ_ERR_INVALID, "null argument");
  nq_hits* h = new nq_hits();
  h->ptr.assign(ptr, ptr + nq + 1);
  const uint64_t total = ptr[nq];
  if (total && (!counts || !gids)) {
    delete h;
    return nq_set_error(NQ_ERR_INVALID, "null argument");
  }
  h->counts.assign(counts, counts + total);
  h->gids.assign(gids, gids + total);
  *out = h;
  return NQ_OK;
}
