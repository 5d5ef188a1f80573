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
// size-biased list statistics over a sample of cells: sums[0] += len, sums[1] += len^2, sums[3] += len * granules(G=8)
__global__ void __launch_bounds__(256) slab_stat_kernel(const uint32_t* __restrict__ dir, uint32_t row_stride, uint32_t range,
                                                        uint32_t F, uint32_t cells, unsigned long long* __restrict__ sums) {
  const uint32_t cell = (uint32_t)((uint64_t)blockIdx.x * F / cells);
  const uint32_t* row = dir + (size_t)cell * row_stride;
  unsigned long long s1 = 0, s2 = 0, s3 = 0;
  for (uint32_t f = threadIdx.x; f < range; f += blockDim.x) {
    const uint32_t w = row[f], len = (w >> 16) - (w & 0xFFFFu);
    s1 += len; s2 += (unsigned long long)len * len;
    s3 += (unsigned long long)len * min(3u, (len + 7) / 8);  // granules a probe of this list gathers at G = 8
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, d);
    s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, d);
    s3 += __shfl_xor_sync(0xFFFFFFFFu, s3, d);
  }
  if ((threadIdx.x & 31) == 0 && s1) { atomicAdd(sums, s1); atomicAdd(sums + 1, s2); atomicAdd(sums + 3, s3); }
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

// Cooperative L2 prefetch of the granules of upcoming cells, driven by whoever gets there first:
// the CTAs of a launch sweep the cells in roughly the same order, and the first kSlabPfSlices of
// them to enter chunk c (kSlabPfCells cells) each pull one slice of chunk c + lead into L2 through
// the bulk-copy engine.  DRAM then sees long sequential reads (the slab is read about once per
// launch) and the 16-byte gathers of every query hit L2, whatever the spread between the fastest and
// the slowest CTA.  claim[] is zeroed before the launch.  Purely a hint.
constexpr uint32_t kSlabPfSlices = 64;
template <int G>
__device__ __forceinline__ void slab_prefetch_claim(const QueryArgs& a, uint32_t chunk, uint32_t lane) {
  const uint32_t c0 = chunk * kSlabPfCells;
  if (c0 >= a.F) return;
  uint32_t t = 0;
  if (lane == 0) t = atomicAdd(a.pf_claim + chunk, 1u);
  t = __shfl_sync(0xFFFFFFFFu, t, 0);
  if (t >= kSlabPfSlices) return;
  const uint32_t c1 = min(a.F, c0 + kSlabPfCells);
  const uint64_t g0 = __ldg(a.cell_gran + c0), g1 = __ldg(a.cell_gran + c1);
  const uint64_t bytes = (g1 - g0) * (G * 2);
  const uint64_t per = ((bytes + kSlabPfSlices - 1) / kSlabPfSlices + 15) & ~15ull, lo = (uint64_t)t * per;
  if (lo >= bytes) return;
  const uint32_t mine = (uint32_t)min(per, bytes - lo);
  const uint32_t piece = ((mine + 31) / 32 + 15) & ~15u, o = lane * piece;
  if (o < mine)
    l2_prefetch_bulk(reinterpret_cast<const char*>(a.slab) + g0 * (G * 2) + lo + o, (min(piece, mine - o) + 15) & ~15u);
}

// ---- step 1: resolve.  desc[q][cell] = first granule (29 bits) | granules inline (2 bits) | tail flag of
// the list that query q probes in `cell`.  One lane per cell, four probes in flight per lane and no
// shared memory, so the SMs are full of warps that do nothing but cover the latency of the meta reads
// (inside the counting kernel the same reads sat in front of every group with one load in flight).
constexpr uint32_t kDescOffBits = 29;
constexpr uint32_t kDescOffMask = (1u << kDescOffBits) - 1u;

__global__ void __launch_bounds__(256) slab_resolve_kernel(QueryArgs a, uint64_t q0, uint32_t nqb, uint32_t* __restrict__ desc) {
  constexpr int U = 4;
  const uint32_t lane = threadIdx.x & 31;
  // warps in cell-major order: neighbouring warps resolve the same 128 cells for different queries, so
  // the meta rows of those cells are read from DRAM once and then found in L2
  const uint64_t gw = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t wpq = (a.F + 32 * U - 1) / (32 * U);  // warps per query
  if (gw >= (uint64_t)wpq * nqb) return;
  const uint64_t ql = gw % nqb;
  const uint32_t c0 = (uint32_t)(gw / nqb) * (32 * U);
  const int32_t* sk = a.qsk + (q0 + ql) * a.F;
  uint32_t fp[U];
  uint4 mw[U];
#pragma unroll
  for (int k = 0; k < U; ++k) {
    const uint32_t cell = c0 + 32 * k + lane;
    fp[k] = cell < a.F ? (uint32_t)__ldcs(&sk[cell]) : 0xFFFFFFFFu;  // read once: streaming
  }
#pragma unroll
  for (int k = 0; k < U; ++k) {
    const uint32_t cell = min(c0 + 32 * k + lane, a.F - 1);
    mw[k] = __ldg(a.meta + (size_t)cell * a.mgroups + (min(fp[k], a.range - 1) >> 5));
  }
#pragma unroll
  for (int k = 0; k < U; ++k) {
    const uint32_t cell = c0 + 32 * k + lane;
    const uint32_t sh = fp[k] & 31u, lt = (1u << sh) - 1u;
    uint32_t d = 0;
    if (fp[k] < a.range) {
      const uint32_t cls = ((mw[k].x >> sh) & 1u) + 2u * ((mw[k].y >> sh) & 1u), ext = (mw[k].z >> sh) & 1u;
      const uint32_t off = mw[k].w + __popc(mw[k].x & lt) + 2 * __popc(mw[k].y & lt) + __popc(mw[k].z & lt);
      d = cls ? off | (cls << kDescOffBits) | (ext << 31) : 0u;
    }
    if (cell < a.F) desc[ql * a.F + cell] = d;
  }
}

// ---- step 2: gather + count.  A warp takes 32 cells of its query per group: descriptors (coalesced,
// fetched ahead), two ballots place every list's granules in the group's work list, rounds of 32 lanes x
// 8 bytes gather whole granules, a register ring keeps D batches in flight, ids are counted with shared-
// memory atomics.  MODE kDual16 (S <= 15): the CTA counts TWO queries, warps [0, NW/2) the first and
// [NW/2, NW) the second, into one u32 word per genome — low half-word = first query (a count never
// exceeds F <= 32768, so it cannot carry) — which makes counting an id `min, shift, ATOMS` with a
// per-warp constant increment instead of picking the half-word of a packed pair.  Padding ids (>= n) are
// clamped to n + lane: 32 spare words, one bank each.
// NRF > 0 (G = 8, short lists): a group's granules are gathered in exactly NRF rounds of straight-line
// code (16 granules each; slots past the group's total read the dummy granule), kept in a ring of D
// groups; the rare group with more granules takes its extra rounds on the spot.  NRF = 0: batches of R
// rounds as the lists need them.
template <int MODE, int NT, int G, int NRF>
__global__ void __launch_bounds__(NT, NT == 128 ? 8 : NT == 256 ? 4 : NT == 512 ? 2 : 1)
query_slab_kernel(QueryArgs a, uint64_t q0, uint32_t nqb, const uint32_t* __restrict__ desc) {
  static_assert(MODE == kDual16 || MODE == kSmem32, "shared-memory counters, one word per genome");
  constexpr bool DUAL = MODE == kDual16;
  constexpr int NW = NT / 32, NWQ = DUAL ? NW / 2 : NW;  // warps per query
  constexpr int LPG = G / 4;     // lanes per granule: every lane takes 8 bytes = 4 ids
  constexpr int GPR = 32 / LPG;  // granules per round
  constexpr int R = NRF ? NRF : 4, D = 3;  // rounds per ring slot, slots in the register ring
  constexpr int TAB = 96 + GPR * (NRF ? NRF : 1);  // <= 3 granules per cell + the dead tail
  constexpr int AHEAD = 3;       // descriptor groups in flight
  extern __shared__ __align__(16) uint32_t smem[];
  __shared__ uint32_t s_tab[NW][TAB];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t half = DUAL ? warp / NWQ : 0, qwarp = DUAL ? warp % NWQ : warp;
  const uint32_t ql = blockIdx.x * (DUAL ? 2 : 1) + half;  // query inside the launch
  const bool have_q = ql < nqb;
  const uint32_t n = a.n;
  constexpr unsigned kFull = 0xFFFFFFFFu;

  if (warp == 0 && a.prefetch)
    for (uint32_t ch = 0; ch < a.prefetch; ++ch) slab_prefetch_claim<G>(a, ch, lane);
  for (uint32_t i = tid; i < n + kSlabPads; i += NT) smem[i] = 0;
  __syncthreads();

  const uint32_t F = have_q ? a.F : 0;  // a CTA's missing second query walks no cells
  const uint32_t* dq = desc + (size_t)(have_q ? ql : 0) * a.F;
  uint32_t* tab = s_tab[warp];
  const uint32_t sub = lane % LPG, grp = lane / LPG;
  const unsigned below = (1u << lane) - 1;
  const uint32_t inc = half ? 0x10000u : 1u;
  const uint32_t pad = n + lane;  // padding ids (>= n) are clamped to one spare word per lane: min, shift, ATOMS
#ifdef NQ_TUNING
  uint32_t sink = 0;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
#endif
  auto count1 = [&](uint32_t id) {
#ifdef NQ_TUNING
    if (a.exp & 4u) { red_shared_add_if_lt(sbase + id * 4u, inc, id, n); return; }  // experiment: predicated instead of clamped
#endif
    atomicAdd(&smem[min(id, pad)], inc);
  };
  auto count2 = [&](uint32_t w) {  // two ids in one word
#ifdef NQ_TUNING
    if (a.exp & 1u) { sink ^= w; return; }  // experiment: gathers without counting
#endif
    count1(w & 0xFFFFu);
    count1(w >> 16);
  };
  auto load_gran = [&](uint32_t g) -> uint2 {
#ifdef NQ_TUNING
    if (a.exp & 2u) return make_uint2((g * 0x10003u) % n * 0x10001u, (g * 7u) % n * 0x10001u);  // experiment: no gathers
#endif
    return __ldg(a.slab + (g * (uint32_t)LPG + sub));
  };

  const uint32_t step = NWQ * 32;
  uint32_t c_cur = qwarp * 32;
  uint32_t dn[AHEAD];  // descriptors of the next AHEAD groups
#pragma unroll
  for (int k = 0; k < AHEAD; ++k) dn[k] = c_cur + k * step + lane < F ? __ldcs(&dq[c_cur + k * step + lane]) : 0u;

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
    if (nb >= (uint32_t)R) {
#pragma unroll
      for (int k = 0; k < R; ++k) l[k] = load_gran(t0[k * GPR + grp]);
    } else {
#pragma unroll
      for (int k = 0; k < R - 1; ++k)
        if ((uint32_t)k < nb) l[k] = load_gran(t0[k * GPR + grp]);
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

  // one group of 32 cells: its granules go into ring slot `gb` (NRF) or through batch(); the slot filled
  // D-1 groups ago (`cb`) is counted while these gathers fly
  auto group = [&](uint2 (&gb)[R], uint2 (&cb)[R], bool& cb_live, bool& gb_live) {
    if (warp == 0 && a.prefetch && c_cur % kSlabPfCells == 0) slab_prefetch_claim<G>(a, c_cur / kSlabPfCells + a.prefetch, lane);
    const uint32_t d = dn[0];
#pragma unroll
    for (int k = 0; k + 1 < AHEAD; ++k) dn[k] = dn[k + 1];
    dn[AHEAD - 1] = c_cur + AHEAD * step + lane < F ? __ldcs(&dq[c_cur + AHEAD * step + lane]) : 0u;

    const uint32_t off = d & kDescOffMask, c0 = (d >> kDescOffBits) & 1u, c1 = (d >> (kDescOffBits + 1)) & 1u;
    const unsigned bal0 = __ballot_sync(kFull, c0), bal1 = __ballot_sync(kFull, c1), bale = __ballot_sync(kFull, d >> 31);
    const uint32_t excl = __popc(bal0 & below) + 2 * __popc(bal1 & below);
    const uint32_t total = __popc(bal0) + 2 * __popc(bal1);
    if (c0 | c1) tab[excl] = off;
    if (c1) tab[excl + 1] = off + 1;
    if (c0 & c1) tab[excl + 2] = off + 2;
    if (NRF) {
      // dead slots up to the fixed rounds (and of the last extra round) gather the dummy granule
      for (uint32_t i = total + lane; i < max((uint32_t)(NRF * GPR), (total + GPR - 1) / GPR * GPR); i += 32) tab[i] = 0;
    } else if (lane < GPR) {
      tab[total + lane] = 0;
    }
    __syncwarp();
    if (NRF) {
#pragma unroll
      for (int k = 0; k < R; ++k) gb[k] = load_gran(tab[k * GPR + grp]);
      gb_live = true;
      if (cb_live) {
#pragma unroll
        for (int k = 0; k < R; ++k) { count2(cb[k].x); count2(cb[k].y); }
        cb_live = false;
      }
      for (uint32_t r = NRF * GPR; r < total; r += GPR) {  // rare: more granules than the fixed rounds hold
        const uint2 v = load_gran(tab[r + grp]);
        count2(v.x); count2(v.y);
      }
    } else {
      const uint32_t rounds = (total + GPR - 1) / GPR;
      for (uint32_t r0 = 0; r0 < rounds; r0 += R) batch(tab + r0 * GPR, min((uint32_t)R, rounds - r0));
    }
    if (bale) {  // tails of lists longer than 3 granules (rare): straight from the CSR posting array
      unsigned m = bale;
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t dg = __shfl_sync(kFull, off + 3, src);
        const uint2 td = __ldg(a.slab + (size_t)dg * LPG);
        const uint16_t* g = static_cast<const uint16_t*>(a.gids) + (size_t)(c_cur + src) * a.gid_stride + td.x;
        for (uint32_t i = lane; i < td.y; i += 32) count1(g[i]);
      }
    }
    __syncwarp();  // the table is rewritten by the next group
    c_cur += step;
  };

  if (NRF) {
    bool lv[D] = {false, false, false};
    while (c_cur < F) {
      group(lbuf[0], lbuf[1], lv[1], lv[0]);
      if (c_cur >= F) break;
      group(lbuf[1], lbuf[2], lv[2], lv[1]);
      if (c_cur >= F) break;
      group(lbuf[2], lbuf[0], lv[0], lv[2]);
    }
#pragma unroll
    for (int p = 0; p < D; ++p)
      if (lv[p]) {
#pragma unroll
        for (int k = 0; k < R; ++k) { count2(lbuf[p][k].x); count2(lbuf[p][k].y); }
      }
  } else {
    bool unused0 = false, unused1 = false;
    while (c_cur < F) group(lbuf[0], lbuf[0], unused0, unused1);
#pragma unroll
    for (int i = 1; i < D; ++i) {
#pragma unroll
      for (int p = 0; p < D; ++p)
        if (phase == (uint32_t)p) drain(lbuf[(p + i) % D], live[(p + i) % D]);
    }
  }
  __syncthreads();
#ifdef NQ_TUNING
  if (sink == 0x12345u) smem[0] = sink;
#endif
  const uint64_t qa = q0 + (uint64_t)blockIdx.x * (DUAL ? 2 : 1);
  query_finish<MODE, NT, true>(a, qa, smem, 0, 0);
  if (DUAL && qa + 1 < q0 + nqb) {
    __syncthreads();
    query_finish<MODE, NT, true>(a, qa + 1, smem, 0, 16);
  }
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
  const size_t full = (size_t)(ix->n + kSlabPads) * 4;  // one u32 word per genome (+ 32 spare): two queries (S <= 15) or one
  if (full + fixed > optin) return false;
  mode = ix->p.S <= 15 ? kDual16 : kSmem32;
  smem = full;
  return true;
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
  unsigned long long sums[4] = {0, 0, 0, 0};
  const uint32_t sample = std::min<uint32_t>(F, 512);
  if ((e = cudaMemsetAsync(d_sums, 0, 32, ctx->stream)) != cudaSuccess) return fail(nq_set_error(NQ_ERR_CUDA, "memset failed"));
  slab_stat_kernel<<<sample, 256, 0, ctx->stream>>>(dir, ix->row_stride, range, F, sample, d_sums);
  ctx->launches++;
  if ((e = cudaMemcpyAsync(sums, d_sums, 32, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess ||
      (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess)
    return fail(nq_set_error(NQ_ERR_CUDA, "slab statistics failed: %s", cudaGetErrorString(e)));
  // ids per granule ~ the length of a probed list (size-biased mean): 1-2 granules per probe
  const double m = sums[0] ? (double)sums[1] / (double)sums[0] : 1.0;
  // Short lists (m <= 12: shards of up to ~15k bacterial genomes at the default W) stay on the CSR kernels of
  // query.cu: measured on B200 the 8-id granule form loses to them there (1000 queries vs 10k genomes: 0.77 vs
  // 0.64 ms; 10k vs 12.5k: 7.2 vs 6.6 ms — both forms sit at the L1TEX wavefront rate of an SM, and the CSR
  // form's sequential L2 prefetch survives the drift between CTAs better), and wins from 16-id granules on
  // (25k genomes: 8.6 vs 13.3 ms; 50k: 14.2 vs 17.4 ms).
  uint32_t G = m <= 12.0 ? 0u : m <= 24.0 ? 16u : m <= 48.0 ? 32u : 64u;
  if (env && atoi(env) >= 8) G = (uint32_t)atoi(env);
#ifdef NQ_TUNING
  if (G != 0 && G != 8 && G != 16 && G != 32 && G != 64) G = 0;
#else
  if (G != 0 && G != 16 && G != 32 && G != 64) G = 0;
#endif
  if (G == 0) {
    nq_dfree(ctx, d_sums);
    return NQ_OK;
  }

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
  if (granules + 4 >= (1ull << kDescOffBits)) {  // granule numbers travel in 29 bits of a descriptor: else the CSR kernels
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
  // G = 8: rounds of 16 granules a group of 32 probes takes in the count kernel's straight-line path
  // (mean + ~2 sigma of the granules of 32 probes must fit; groups beyond it take extra rounds)
  const double gpp = sums[0] ? (double)sums[3] / (double)sums[0] : 1.0;
  ix->slab_nrf = G == 8 ? (32.0 * gpp + 4.0 <= 32.0 ? 2u : 32.0 * gpp + 4.0 <= 48.0 ? 3u : 4u) : 0u;
  if (const char* nr = nq_tuning_env("NQ_SLAB_NRF")) ix->slab_nrf = G == 8 ? (uint32_t)atoi(nr) : 0u;
  return NQ_OK;
}

// ---- launch.  Queries per CTA: 2 (kDual16) or 1.  CTA size and CTAs per SM follow the shared memory the
// counters take: as many CTAs as leave the SM >= ~56 KB of L1 (outstanding gathers allocate L1 lines: with
// less, the SM cannot keep enough of them in flight), enforced through the shared-memory carve-out.
struct SlabCfg { int nt, per_sm, carve; };
static SlabCfg slab_cfg(const nq_index* ix, int mode, size_t smem, uint64_t nq_total) {
  const size_t per = smem + 1024 + 2048, sm = 228 * 1024, l1 = 56 * 1024;
  const int fit = (int)std::max<size_t>(1, std::min<size_t>(8, (sm - l1) / per));
  SlabCfg c;
  // threads: 2048 per SM over the resident CTAs (64-register kernels), at least 2 warps per query
  c.nt = fit >= 8 && mode != kDual16 ? 128 : fit >= 4 ? 256 : fit >= 2 ? 512 : 1024;
  const char* env = nq_tuning_env("NQ_QUERY_NT");
  if (env && atoi(env) >= 128) c.nt = atoi(env);
  c.per_sm = std::max(1, std::min(fit, 2048 / c.nt));
  c.carve = (int)std::min<size_t>(100, (c.per_sm * per * 100 + sm - 1) / sm + 1);
  (void)nq_total;
  return c;
}

template <int MODE, int NT, int G, int NRF>
static cudaError_t launch_slab_t(const SlabCfg& cfg, size_t smem, unsigned nb, const QueryArgs& a, uint64_t q0, cudaStream_t st, int* occ) {
  auto k = query_slab_kernel<MODE, NT, G, NRF>;
  constexpr unsigned QPC = MODE == kDual16 ? 2 : 1;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cfg.carve)) != cudaSuccess) return e;
  if (occ) {  // resident QUERIES per SM (wave sizing), no launch
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k, NT, smem);
    *occ = std::min(*occ, cfg.per_sm) * (int)QPC;
    return e;
  }
  // step 1 (descriptors of the launch's queries), then step 2
  if (a.prefetch && (e = cudaMemsetAsync(a.pf_claim, 0, (a.F / kSlabPfCells + 16) * sizeof(uint32_t), st)) != cudaSuccess) return e;
  const uint64_t warps = (uint64_t)nb * ((a.F + 127) / 128);
  slab_resolve_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(a, q0, nb, a.desc);
  k<<<(nb + QPC - 1) / QPC, NT, smem, st>>>(a, q0, nb, a.desc);
  return cudaSuccess;
}
template <int MODE, int G, int NRF>
static cudaError_t launch_slab_nt(const SlabCfg& cfg, size_t smem, unsigned nb, const QueryArgs& a, uint64_t q0, cudaStream_t st, int* occ) {
  switch (cfg.nt) {
    case 128: if (MODE != kDual16) return launch_slab_t<MODE, MODE == kDual16 ? 256 : 128, G, NRF>(cfg, smem, nb, a, q0, st, occ);  // fallthrough
    case 256: return launch_slab_t<MODE, 256, G, NRF>(cfg, smem, nb, a, q0, st, occ);
    case 512: return launch_slab_t<MODE, 512, G, NRF>(cfg, smem, nb, a, q0, st, occ);
    default: return launch_slab_t<MODE, 1024, G, NRF>(cfg, smem, nb, a, q0, st, occ);
  }
}
template <int MODE>
static cudaError_t launch_slab_g(uint32_t G, uint32_t nrf, const SlabCfg& cfg, size_t smem, unsigned nb, const QueryArgs& a, uint64_t q0, cudaStream_t st, int* occ) {
  (void)nrf;
  switch (G) {
#ifdef NQ_TUNING  // 8-id granules only exist in measurement builds (nq_slab_build never picks them)
    case 8:
      if (nrf == 2) return launch_slab_nt<MODE, 8, 2>(cfg, smem, nb, a, q0, st, occ);
      if (nrf == 3) return launch_slab_nt<MODE, 8, 3>(cfg, smem, nb, a, q0, st, occ);
      if (nrf == 4) return launch_slab_nt<MODE, 8, 4>(cfg, smem, nb, a, q0, st, occ);
      return launch_slab_nt<MODE, 8, 0>(cfg, smem, nb, a, q0, st, occ);
#endif
    case 16: return launch_slab_nt<MODE, 16, 0>(cfg, smem, nb, a, q0, st, occ);
    case 32: return launch_slab_nt<MODE, 32, 0>(cfg, smem, nb, a, q0, st, occ);
    default: return launch_slab_nt<MODE, 64, 0>(cfg, smem, nb, a, q0, st, occ);
  }
}

cudaError_t nq_slab_launch(const nq_index* ix, int mode, size_t smem, unsigned nb, const QueryArgs& a, uint64_t q0,
                           cudaStream_t st, int* occ) {
  const SlabCfg cfg = slab_cfg(ix, mode, smem, a.nq_total);
  return mode == kDual16 ? launch_slab_g<kDual16>(ix->slab_G, ix->slab_nrf, cfg, smem, nb, a, q0, st, occ)
                         : launch_slab_g<kSmem32>(ix->slab_G, ix->slab_nrf, cfg, smem, nb, a, q0, st, occ);
}
