// slab.cu — the granule ("slab") form of the posting lists and the query kernel that reads it.
// Replaces the gather of Index::query_sketch (/root/reference/src/niqki_index.cpp:654-660) for
// shards whose ids fit 16 bits; the cell-major CSR of index.cu stays the canonical form (export,
// dump, --matrix row extraction), the slab is derived from it once per build.
//
// Why: with the CSR a probe of list (cell, fp) is a directory word, then 1-2 unaligned posting
// sectors behind it — 3.3 L2 sectors and two dependent loads for ~38 algorithmic bytes, and the
// r01 kernel sat at half of the L2 sector rate whatever its instruction count.  Lists are short and
// tightly distributed (a probed list holds 8 +- 3 ids at 10k genomes, 97.8 % hold <= 16), so:
//   * every list occupies 1..3 GRANULES of G ids (G*2 bytes, granule-aligned, padded with ids
//     n..n+31, which are counted into 32 spare counter words instead of being tested);
//   * the directory shrinks to 16 bytes per 32 fingerprints: meta[cell][fp/32] = {b0, b1, b2, base};
//     bit fp%32 of (b0 + 2*b1) is the list's granule count, its first granule is
//     base + popc(b0 & below) + 2*popc(b1 & below) + popc(b2 & below).  2 KB per cell at W=12, so the
//     CTAs that sweep the cells together mostly find it in L1;
//   * b2 marks the rare list longer than 3 granules: one more granule follows the three inline ones
//     and holds {begin, count} of the tail, which stays in the CSR posting array.
// A warp takes 32 cells: one 16-byte meta read per lane, two ballots give every lane the position
// of its granules in the group's work list (a per-warp table in shared memory), then rounds of 32
// lanes x 8 bytes gather whole granules — 4 ids per lane and load, no per-posting bookkeeping.
// Same register ring (batches of R rounds, D in flight), cooperative L2 prefetch of upcoming cells
// and wave-sized launches as the CSR kernels of query.cu.
#include <algorithm>
#include <vector>

#include "query_common.cuh"

namespace nq {

constexpr uint32_t kSlabPads = 32;  // padding ids n .. n+31

__host__ __device__ __forceinline__ uint32_t slab_pad(uint32_t n, uint32_t fp, uint32_t k) { return n + ((fp + k) & 31u); }

// ---- build ---------------------------------------------------------------------------------------
// size-biased mean list length over a sample of cells: sums[0] += len, sums[1] += len^2
__global__ void __launch_bounds__(256) slab_stat_kernel(const uint32_t* __restrict__ dir, uint32_t row_stride, uint32_t range,
                                                        uint32_t F, uint32_t cells, unsigned long long* __restrict__ sums) {
  const uint32_t cell = (uint32_t)((uint64_t)blockIdx.x * F / cells);
  const uint32_t* row = dir + (size_t)cell * row_stride;
  unsigned long long s1 = 0, s2 = 0;
  for (uint32_t f = threadIdx.x; f < range; f += blockDim.x) {
    const uint32_t w = row[f], len = (w >> 16) - (w & 0xFFFFu);
    s1 += len; s2 += (unsigned long long)len * len;
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, d);
    s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, d);
  }
  if ((threadIdx.x & 31) == 0 && s1) { atomicAdd(sums, s1); atomicAdd(sums + 1, s2); }
}

struct SlabGroup { uint32_t b0, b1, b2, gran; };
// the 32 lists of directory group `grp` of a cell: granule-count bit planes and granules used
__device__ __forceinline__ SlabGroup slab_group(const uint32_t* __restrict__ row, uint32_t grp, uint32_t G) {
  SlabGroup s{0, 0, 0, 0};
  const uint4* r4 = reinterpret_cast<const uint4*>(row + grp * 32);
#pragma unroll
  for (int v = 0; v < 8; ++v) {
    const uint4 q = __ldg(r4 + v);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t len = (w[j] >> 16) - (w[j] & 0xFFFFu);
      const uint32_t cls = min(3u, (len + G - 1) / G), ext = len > 3 * G ? 1u : 0u;
      s.b0 |= (cls & 1u) << (v * 4 + j);
      s.b1 |= (cls >> 1) << (v * 4 + j);
      s.b2 |= ext << (v * 4 + j);
      s.gran += cls + ext;
    }
  }
  return s;
}

// one CTA per cell, one thread per directory group (blockDim.x >= mgroups): granules of the cell
__global__ void slab_count_kernel(const uint32_t* __restrict__ dir, uint32_t row_stride, uint32_t mgroups, uint32_t G,
                                  uint32_t* __restrict__ sizes) {
  __shared__ uint32_t s_warp[32];
  const uint32_t cell = blockIdx.x, tid = threadIdx.x;
  uint32_t g = tid < mgroups ? slab_group(dir + (size_t)cell * row_stride, tid, G).gran : 0u;
  g = __reduce_add_sync(0xFFFFFFFFu, g);
  if ((tid & 31) == 0) s_warp[tid >> 5] = g;
  __syncthreads();
  if (tid == 0) {
    uint32_t t = 0;
    for (uint32_t w = 0; w < (blockDim.x + 31) / 32; ++w) t += s_warp[w];
    sizes[cell] = t;
  }
}

// sizes[c] = granules of cell c  ->  cell_gran[c] = first granule of cell c (granule 0 is the dummy),
// cell_gran[F] = total.  One CTA.
__global__ void __launch_bounds__(1024) slab_scan_kernel(const uint32_t* __restrict__ sizes, uint32_t* __restrict__ cell_gran,
                                                         uint32_t F, unsigned long long* __restrict__ total) {
  __shared__ unsigned long long s_warp[32];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t per = (F + 1023) / 1024, c0 = min(F, tid * per), c1 = min(F, c0 + per);
  unsigned long long mine = 0;
  for (uint32_t c = c0; c < c1; ++c) mine += sizes[c];
  unsigned long long incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  unsigned long long run = incl - mine + 1;  // granule 0 = dummy
  for (uint32_t w = 0; w < warp; ++w) run += s_warp[w];
  for (uint32_t c = c0; c < c1; ++c) {
    cell_gran[c] = (uint32_t)run;
    run += sizes[c];
  }
  if (tid == 1023) {
    cell_gran[F] = (uint32_t)run;
    *total = run;
  }
}

// one CTA per cell: meta row + the cell's granules.  Pass A, one thread per directory group: bit planes,
// the group's first granule, and every list's first granule into shared memory.  Pass B, threads
// strided over the LISTS (popular fingerprints are neighbours, so the copy work spreads over all
// threads): ids, padding and tail descriptors.  Cells of at most `img_granules` granules are assembled
// in shared memory and leave with 16-byte stores.
__global__ void slab_fill_kernel(const uint32_t* __restrict__ dir, uint32_t row_stride, uint32_t range, uint32_t mgroups, uint32_t G,
                                 const uint16_t* __restrict__ gids, uint32_t gid_stride, uint32_t n,
                                 const uint32_t* __restrict__ cell_gran, uint4* __restrict__ meta, uint16_t* __restrict__ slab,
                                 uint32_t img_granules) {
  extern __shared__ __align__(16) uint16_t s_dyn[];  // [range] u32 list offsets, then the granule image
  __shared__ uint32_t s_warp[32];
  uint32_t* s_goff = reinterpret_cast<uint32_t*>(s_dyn);
  uint16_t* s_img = s_dyn + (size_t)range * 2;
  const uint32_t cell = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t* row = dir + (size_t)cell * row_stride;
  SlabGroup s{0, 0, 0, 0};
  if (tid < mgroups) s = slab_group(row, tid, G);
  uint32_t incl = s.gran;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  uint32_t rel = incl - s.gran;  // first granule of this group inside the cell
  for (uint32_t w = 0; w < warp; ++w) rel += s_warp[w];
  const uint32_t cell_first = cell_gran[cell], cell_total = cell_gran[cell + 1] - cell_first;
  if (tid < mgroups) {
    meta[(size_t)cell * mgroups + tid] = make_uint4(s.b0, s.b1, s.b2, cell_first + rel);
    uint32_t g = rel;
#pragma unroll 8
    for (uint32_t j = 0; j < 32; ++j) {
      const uint32_t cls = ((s.b0 >> j) & 1u) + 2u * ((s.b1 >> j) & 1u), ext = (s.b2 >> j) & 1u;
      s_goff[tid * 32 + j] = g | (cls << 28) | (ext << 30);
      g += cls + ext;
    }
  }
  __syncthreads();
  const bool staged = cell_total <= img_granules;
  uint16_t* out = staged ? s_img : slab + (size_t)cell_first * G;  // granule g of the cell at out + g*G
  const uint16_t* src = gids + (size_t)cell * gid_stride;
  for (uint32_t fp = tid; fp < range; fp += blockDim.x) {
    const uint32_t go = s_goff[fp], cls = (go >> 28) & 3u;
    if (!cls) continue;
    const uint32_t g = go & 0x0FFFFFFFu, w = __ldg(row + fp), b = w & 0xFFFFu, len = (w >> 16) - b;
    const uint32_t inl = min(len, cls * G);
    uint16_t* o = out + (size_t)g * G;
    for (uint32_t k = 0; k < inl; ++k) o[k] = src[b + k];
    for (uint32_t k = inl; k < cls * G; ++k) o[k] = (uint16_t)slab_pad(n, fp, k);
    if (go >> 30) {  // descriptor granule: {begin inside the cell's CSR row, count} of the tail
      uint16_t* d = o + (size_t)cls * G;
      const uint32_t tb = b + 3 * G, tc = len - 3 * G;
      d[0] = (uint16_t)tb; d[1] = (uint16_t)(tb >> 16); d[2] = (uint16_t)tc; d[3] = (uint16_t)(tc >> 16);
      for (uint32_t k = 4; k < G; ++k) d[k] = (uint16_t)slab_pad(n, fp, k);
    }
  }
  if (staged) {
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(slab + (size_t)cell_first * G);
    const uint4* s4 = reinterpret_cast<const uint4*>(s_img);
    const uint32_t nvec = cell_total * G / 8;
    for (uint32_t v = tid; v < nvec; v += blockDim.x) dst[v] = s4[v];
  }
}

__global__ void slab_dummy_kernel(uint16_t* __restrict__ slab, uint32_t G, uint32_t n) {
  if (threadIdx.x < G) slab[threadIdx.x] = (uint16_t)(n + (threadIdx.x & 31u));
}

// ---- query ---------------------------------------------------------------------------------------
constexpr uint32_t kSlabPfCells = 256;  // cells per prefetch chunk

// this CTA's slice of a region, cut into <= 32 bulk requests (one per lane of warp 0)
__device__ __forceinline__ void slab_prefetch_region(const char* base, uint64_t bytes, uint32_t lane) {
  const uint64_t per_cta = ((bytes + gridDim.x - 1) / gridDim.x + 15) & ~15ull;
  const uint64_t lo = (uint64_t)blockIdx.x * per_cta;
  if (lo >= bytes) return;
  const uint32_t mine = (uint32_t)min(per_cta, bytes - lo);
  const uint32_t piece = max(2048u, ((mine + 31) / 32 + 15) & ~15u);
  const uint32_t o = lane * piece;
  if (o < mine) l2_prefetch_bulk(base + lo + o, (min(piece, mine - o) + 15) & ~15u);
}
template <int G>
__device__ __forceinline__ void slab_prefetch_chunk(const QueryArgs& a, uint32_t chunk, uint32_t lane) {
  const uint32_t c0 = chunk * kSlabPfCells;
  if (c0 >= a.F) return;
  const uint32_t c1 = min(a.F, c0 + kSlabPfCells);
  slab_prefetch_region(reinterpret_cast<const char*>(a.meta + (size_t)c0 * a.mgroups), (uint64_t)(c1 - c0) * a.mgroups * 16, lane);
  const uint32_t g0 = __ldg(a.cell_gran + c0), g1 = __ldg(a.cell_gran + c1);
  slab_prefetch_region(reinterpret_cast<const char*>(a.slab) + (size_t)g0 * (G * 2), (uint64_t)(g1 - g0) * (G * 2), lane);
}

template <int MODE, int NT, int G>
__global__ void __launch_bounds__(NT, NT == 128 ? 8 : NT == 256 ? 4 : NT == 512 ? 2 : 1) query_slab_kernel(QueryArgs a, uint64_t q0) {
  static_assert(MODE == kPack16 || MODE == kSmem32, "shared-memory counters");
  constexpr int LPG = G / 4;     // lanes per granule: every lane takes 8 bytes = 4 ids
  constexpr int GPR = 32 / LPG;  // granules per round
  constexpr int R = 4, D = 3;    // rounds per batch, batches in the register ring
  constexpr int TAB = 96 + GPR;  // <= 3 granules per cell + the dead tail of the last round
  extern __shared__ __align__(16) uint32_t smem[];
  __shared__ uint32_t s_tab[NT / 32][TAB];
  const uint64_t q = q0 + blockIdx.x;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t words = MODE == kPack16 ? (a.n + kSlabPads + 1) / 2 : a.n + kSlabPads;
  constexpr unsigned kFull = 0xFFFFFFFFu;

  if (warp == 0 && a.prefetch)
    for (uint32_t ch = 0; ch < a.prefetch; ++ch) slab_prefetch_chunk<G>(a, ch, lane);
  for (uint32_t i = tid; i < words; i += NT) smem[i] = 0;
  __syncthreads();

  const int32_t* sk = a.qsk + q * a.F;
  uint32_t* tab = s_tab[warp];
  const uint32_t sub = lane % LPG, grp = lane / LPG;
  const unsigned below = (1u << lane) - 1;
#ifdef NQ_TUNING
  uint32_t sink = 0;
#endif
  auto count2 = [&](uint32_t w) {  // two ids in one word
#ifdef NQ_TUNING
    if (a.exp & 1u) { sink ^= w; return; }  // experiment: gathers without counting
#endif
    if (MODE == kPack16) {
      atomicAdd(&smem[(w & 0xFFFFu) >> 1], (w & 1u) ? 0x10000u : 1u);
      atomicAdd(&smem[w >> 17], (w & 0x10000u) ? 0x10000u : 1u);
    } else {
      atomicAdd(&smem[w & 0xFFFFu], 1u);
      atomicAdd(&smem[w >> 16], 1u);
    }
  };
  auto count1 = [&](uint32_t id) {
    if (MODE == kPack16) atomicAdd(&smem[id >> 1], (id & 1u) ? 0x10000u : 1u);
    else atomicAdd(&smem[id], 1u);
  };

  // software pipeline over the warp's groups of 32 cells: fingerprints two groups ahead, meta words
  // one group ahead, granule gathers D-1 batches ahead of their counting
  const uint32_t step = NT;
  uint32_t c_cur = warp * 32;
  auto probe = [&](uint32_t cell, uint32_t fp) {  // clamped into the table; fp >= range is masked at decode
    return __ldg(a.meta + (size_t)min(cell, a.F - 1) * a.mgroups + (min(fp, a.range - 1) >> 5));
  };
  uint32_t fp_cur = 0xFFFFFFFFu, fp_next = 0xFFFFFFFFu, fp_next2;
  if (c_cur + lane < a.F) fp_cur = (uint32_t)__ldg(&sk[c_cur + lane]);
  if (c_cur + step + lane < a.F) fp_next = (uint32_t)__ldg(&sk[c_cur + step + lane]);
  uint4 mw = probe(c_cur + lane, fp_cur), mw_next;

  uint2 lbuf[D][R];
  uint32_t live[D];
#pragma unroll
  for (int d = 0; d < D; ++d) live[d] = 0;
  uint32_t phase = 0;
  auto drain = [&](uint2 (&l)[R], uint32_t nl) {
    if (nl >= (uint32_t)R) {
#pragma unroll
      for (int k = 0; k < R; ++k) { count2(l[k].x); count2(l[k].y); }
    } else {
#pragma unroll
      for (int k = 0; k < R - 1; ++k)
        if ((uint32_t)k < nl) { count2(l[k].x); count2(l[k].y); }
    }
  };
  auto gather = [&](uint2 (&l)[R], const uint32_t* t0, uint32_t nb) {
#ifdef NQ_TUNING
    if (a.exp & 2u) {  // experiment: counting without gathers (ids made up from the table)
#pragma unroll
      for (int k = 0; k < R; ++k) l[k] = make_uint2((t0[k * GPR + grp] * 0x10003u) % a.n * 0x10001u, (t0[k * GPR + grp] * 7u) % a.n * 0x10001u);
      return;
    }
#endif
    if (nb >= (uint32_t)R) {
#pragma unroll
      for (int k = 0; k < R; ++k) l[k] = __ldg(a.slab + (t0[k * GPR + grp] * (uint32_t)LPG + sub));
    } else {
#pragma unroll
      for (int k = 0; k < R - 1; ++k)
        if ((uint32_t)k < nb) l[k] = __ldg(a.slab + (t0[k * GPR + grp] * (uint32_t)LPG + sub));
    }
  };
  auto batch = [&](const uint32_t* t0, uint32_t nb) {
#pragma unroll
    for (int p = 0; p < D; ++p)
      if (phase == (uint32_t)p) {
        gather(lbuf[p], t0, nb);
        drain(lbuf[(p + 1) % D], live[(p + 1) % D]);
        live[(p + 1) % D] = 0;
        live[p] = nb;
      }
    phase = phase + 1 == (uint32_t)D ? 0 : phase + 1;
  };

  for (; c_cur < a.F; c_cur += step) {
    if (warp == 0 && a.prefetch && c_cur % kSlabPfCells == 0) slab_prefetch_chunk<G>(a, c_cur / kSlabPfCells + a.prefetch, lane);
    fp_next2 = 0xFFFFFFFFu;
    if (c_cur + 2 * step + lane < a.F) fp_next2 = (uint32_t)__ldg(&sk[c_cur + 2 * step + lane]);
    mw_next = probe(c_cur + step + lane, fp_next);

    // decode: granule count (0..3) and first granule of this lane's list
    const uint32_t sh = fp_cur & 31u, lt = (1u << sh) - 1u;
    const bool valid = fp_cur < a.range;
    const uint32_t c0 = valid ? (mw.x >> sh) & 1u : 0u, c1 = valid ? (mw.y >> sh) & 1u : 0u;
    const uint32_t ce = valid ? (mw.z >> sh) & 1u : 0u;
    const uint32_t off = mw.w + __popc(mw.x & lt) + 2 * __popc(mw.y & lt) + __popc(mw.z & lt);
    const unsigned bal0 = __ballot_sync(kFull, c0), bal1 = __ballot_sync(kFull, c1), bale = __ballot_sync(kFull, ce);
    const uint32_t excl = __popc(bal0 & below) + 2 * __popc(bal1 & below);
    const uint32_t total = __popc(bal0) + 2 * __popc(bal1);
    if (c0 | c1) tab[excl] = off;
    if (c1) tab[excl + 1] = off + 1;
    if (c0 & c1) tab[excl + 2] = off + 2;
    if (lane < GPR) tab[total + lane] = 0;  // dead slots of the last round gather the dummy granule
    __syncwarp();
    const uint32_t rounds = (total + GPR - 1) / GPR;
    for (uint32_t r0 = 0; r0 < rounds; r0 += R) batch(tab + r0 * GPR, min((uint32_t)R, rounds - r0));
    if (bale) {  // tails of lists longer than 3 granules (rare): straight from the CSR posting array
      unsigned m = bale;
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t dg = __shfl_sync(kFull, off + 3, src);
        const uint2 d = __ldg(a.slab + (size_t)dg * LPG);
        const uint16_t* g = static_cast<const uint16_t*>(a.gids) + (size_t)(c_cur + src) * a.gid_stride + d.x;
        for (uint32_t i = lane; i < d.y; i += 32) count1(g[i]);
      }
    }
    __syncwarp();  // the table is rewritten by the next group
    mw = mw_next;
    fp_cur = fp_next;
    fp_next = fp_next2;
  }
#pragma unroll
  for (int i = 1; i < D; ++i) {
#pragma unroll
    for (int p = 0; p < D; ++p)
      if (phase == (uint32_t)p) drain(lbuf[(p + i) % D], live[(p + i) % D]);
  }
  __syncthreads();
#ifdef NQ_TUNING
  if (sink == 0x12345u) smem[0] = sink;
#endif
  query_finish<MODE, NT, true>(a, q, smem, 0, 0);
}

}  // namespace nq

using namespace nq;

void nq_slab_free(nq_index* ix) {
  if (!ix || !ix->ctx) return;
  nq_dfree(ix->ctx, ix->d_meta);
  nq_dfree(ix->ctx, ix->d_slab);
  nq_dfree(ix->ctx, ix->d_cell_gran);
  ix->d_meta = nullptr; ix->d_slab = nullptr; ix->d_cell_gran = nullptr;
  ix->slab_G = 0; ix->slab_granules = 0;
}

// shared-memory counters of the slab kernel: n + 32 padding ids, packed u16 when S <= 15
bool nq_slab_layout(const nq_index* ix, int& mode, size_t& smem) {
  const size_t fixed = 8 * 1024, optin = ix->ctx->smem_optin;
  const size_t pack = (size_t)((ix->n + kSlabPads + 1) / 2) * 4, full = (size_t)(ix->n + kSlabPads) * 4;
  if (ix->p.S <= 15 && pack + fixed <= optin) { mode = kPack16; smem = pack; return true; }
  if (full + fixed <= optin) { mode = kSmem32; smem = full; return true; }
  return false;
}

int nq_slab_build(nq_index* ix) {
  nq_ctx* ctx = ix->ctx;
  const uint32_t range = (uint32_t)ix->p.range, F = ix->p.F;
  const char* env = nq_tuning_env("NQ_SLAB");  // "0": CSR kernels only; "8".."64": force the granule size
  int mode;
  size_t smem;
  if (ix->elem != 2 || range < 32 || range > 32768 || ix->n + kSlabPads > 65536 || !nq_slab_layout(ix, mode, smem) ||
      (env && env[0] == '0'))
    return NQ_OK;
  const uint32_t mgroups = range / 32;
  const uint32_t* dir = static_cast<const uint32_t*>(ix->d_row);
  unsigned long long* d_sums = nullptr;
  NQ_TRY(nq_dmalloc(ctx, (void**)&d_sums, 32));
  uint32_t* d_sizes = nullptr;
  auto fail = [&](int st) {
    nq_dfree(ctx, d_sums);
    nq_dfree(ctx, d_sizes);
    nq_slab_free(ix);
    return st;
  };
  cudaError_t e;
  unsigned long long sums[3] = {0, 0, 0};
  const uint32_t sample = std::min<uint32_t>(F, 512);
  if ((e = cudaMemsetAsync(d_sums, 0, 32, ctx->stream)) != cudaSuccess) return fail(nq_set_error(NQ_ERR_CUDA, "memset failed"));
  slab_stat_kernel<<<sample, 256, 0, ctx->stream>>>(dir, ix->row_stride, range, F, sample, d_sums);
  ctx->launches++;
  if ((e = cudaMemcpyAsync(sums, d_sums, 16, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess ||
      (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess)
    return fail(nq_set_error(NQ_ERR_CUDA, "slab statistics failed: %s", cudaGetErrorString(e)));
  // ids per granule ~ the length of a probed list (size-biased mean): 1-2 granules per probe
  const double m = sums[0] ? (double)sums[1] / (double)sums[0] : 1.0;
  uint32_t G = m <= 12.0 ? 8u : m <= 24.0 ? 16u : m <= 48.0 ? 32u : 64u;
  if (env && atoi(env) >= 8) G = (uint32_t)atoi(env);
  if (G != 8 && G != 16 && G != 32 && G != 64) G = 8;

  int st;
  if ((st = nq_dmalloc(ctx, (void**)&d_sizes, (size_t)F * 4)) != NQ_OK ||
      (st = nq_dmalloc(ctx, (void**)&ix->d_cell_gran, ((size_t)F + 2) * 4)) != NQ_OK ||
      (st = nq_dmalloc(ctx, (void**)&ix->d_meta, (size_t)F * mgroups * sizeof(uint4))) != NQ_OK)
    return fail(st);
  const uint32_t nt = std::max(32u, (mgroups + 31) & ~31u);
  {
    NqTimer timer(ctx, NQK_SLAB);
    slab_count_kernel<<<F, nt, 0, ctx->stream>>>(dir, ix->row_stride, mgroups, G, d_sizes);
    slab_scan_kernel<<<1, 1024, 0, ctx->stream>>>(d_sizes, ix->d_cell_gran, F, d_sums + 2);
  }
  ctx->launches += 2;
  if ((e = cudaMemcpyAsync(&sums[2], d_sums + 2, 8, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess ||
      (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess)
    return fail(nq_set_error(NQ_ERR_CUDA, "slab scan failed: %s", cudaGetErrorString(e)));
  const uint64_t granules = sums[2];
  nq_dfree(ctx, d_sizes);
  d_sizes = nullptr;
  if (granules * (G / 4) >= (1ull << 32)) {  // 32-bit granule arithmetic in the query kernel: fall back to the CSR kernels
    nq_dfree(ctx, d_sums);
    nq_slab_free(ix);
    return NQ_OK;
  }
  if ((st = nq_dmalloc(ctx, (void**)&ix->d_slab, granules * G * 2)) != NQ_OK) return fail(st);
  // cells are assembled in shared memory when they fit beside 3 more CTAs of the same kind
  const size_t goff_bytes = (size_t)range * 4;
  const size_t img_cap = std::max<size_t>(goff_bytes, std::min<size_t>(ctx->smem_optin - 2048, 72 * 1024));
  const uint32_t img_granules = (uint32_t)((img_cap - goff_bytes) / (G * 2));
  if ((e = cudaFuncSetAttribute(slab_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)img_cap)) != cudaSuccess)
    return fail(nq_set_error(NQ_ERR_CUDA, "slab fill attribute: %s", cudaGetErrorString(e)));
  {
    NqTimer timer(ctx, NQK_SLAB);
    slab_dummy_kernel<<<1, 64, 0, ctx->stream>>>(reinterpret_cast<uint16_t*>(ix->d_slab), G, ix->n);
    slab_fill_kernel<<<F, std::max(nt, 128u), img_cap, ctx->stream>>>(dir, ix->row_stride, range, mgroups, G, static_cast<const uint16_t*>(ix->d_gids),
                                                      ix->gid_stride, ix->n, ix->d_cell_gran, ix->d_meta,
                                                      reinterpret_cast<uint16_t*>(ix->d_slab), img_granules);
  }
  ctx->launches += 2;
  if ((e = cudaPeekAtLastError()) != cudaSuccess) return fail(nq_set_error(NQ_ERR_CUDA, "slab fill launch failed: %s", cudaGetErrorString(e)));
  nq_dfree(ctx, d_sums);
  ix->slab_G = G;
  ix->slab_granules = granules;
  return NQ_OK;
}

// ---- launch: CTA size from the shared memory the counters take (the L1 that is left tracks the gathers)
template <int MODE, int NT, int G>
static cudaError_t launch_slab_t(size_t smem, unsigned nb, const QueryArgs& a, uint64_t q0, cudaStream_t st, int* occ) {
  auto k = query_slab_kernel<MODE, NT, G>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (occ) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k, NT, smem);
  k<<<nb, NT, smem, st>>>(a, q0);
  return cudaSuccess;
}
template <int MODE, int G>
static cudaError_t launch_slab_nt(int nt, size_t smem, unsigned nb, const QueryArgs& a, uint64_t q0, cudaStream_t st, int* occ) {
  switch (nt) {
    case 128: return launch_slab_t<MODE, 128, G>(smem, nb, a, q0, st, occ);
    case 256: return launch_slab_t<MODE, 256, G>(smem, nb, a, q0, st, occ);
    case 512: return launch_slab_t<MODE, 512, G>(smem, nb, a, q0, st, occ);
    default: return launch_slab_t<MODE, 1024, G>(smem, nb, a, q0, st, occ);
  }
}
template <int MODE>
static cudaError_t launch_slab_g(uint32_t G, int nt, size_t smem, unsigned nb, const QueryArgs& a, uint64_t q0, cudaStream_t st, int* occ) {
  switch (G) {
    case 8: return launch_slab_nt<MODE, 8>(nt, smem, nb, a, q0, st, occ);
    case 16: return launch_slab_nt<MODE, 16>(nt, smem, nb, a, q0, st, occ);
    case 32: return launch_slab_nt<MODE, 32>(nt, smem, nb, a, q0, st, occ);
    default: return launch_slab_nt<MODE, 64>(nt, smem, nb, a, q0, st, occ);
  }
}

// CTA size: as many queries per SM as leave >= ~40 KB of L1 beside their counters (8 x 128 threads,
// 4 x 256, 2 x 512, else one 1024-thread CTA); long batches of small shards prefer 256 threads.
int nq_slab_cta_threads(const nq_index* ix, size_t smem, uint64_t nq_total) {
  const char* env = nq_tuning_env("NQ_QUERY_NT");
  if (env && atoi(env) >= 128) return atoi(env);
  const size_t per = smem + 1024 + 2048, l1 = 40 * 1024, sm = 228 * 1024;
  if (8 * per + l1 <= sm && nq_total < (uint64_t)ix->ctx->sm_count * 36) return 128;
  if (4 * per + l1 <= sm) return 256;
  if (2 * per + l1 <= sm) return 512;
  return 1024;
}

cudaError_t nq_slab_launch(const nq_index* ix, int mode, size_t smem, unsigned nb, const QueryArgs& a, uint64_t q0,
                           cudaStream_t st, int* occ) {
  const int nt = nq_slab_cta_threads(ix, smem, a.nq_total);
  return mode == kPack16 ? launch_slab_g<kPack16>(ix->slab_G, nt, smem, nb, a, q0, st, occ)
                         : launch_slab_g<kSmem32>(ix->slab_G, nt, smem, nb, a, q0, st, occ);
}
