// query.cu — K4a: posting-list gather + per-genome hit counting + fused min_score threshold and
// compaction.  Replaces Index::query_sketch (/root/reference/src/niqki_index.cpp:633-687).
//
// One CTA per query sketch.  The per-genome counters of the shard live in shared memory (two 16-bit
// counters per word when S <= 15, because a count can never exceed F = 2^S <= 32768 — the same
// widths the reference uses, :635-683), or in HBM/L2 when the shard is too large for that.  A warp
// takes 32 consecutive cells: every lane reads its cell's fingerprint (coalesced, two groups ahead)
// and ONE packed directory word {begin,end} (one group ahead); the 32 lists are then gathered 32
// postings per round through a register ring (the batch gathered before is counted while the next
// one's loads fly).  Two gather forms share that frame:
//   * query_count_kernel ("stream"): the 32 lists are one concatenated stream; the owner of a
//     stream slot is found per round without a search (REDUX.OR of the lists that start inside the
//     round + POPC of the bits at or below the slot);
//   * query_count_seg_kernel ("segment table"): the lists are cut into segments of 8 / 16 / 32
//     postings whose descriptors each cell's lane writes into a per-warp shared-memory table; a
//     round is then one broadcast LDS + a few ALU instructions + the load.  Variants: split16 (u16
//     copy of the postings for shards of 65.6k..131k genomes), many CTAs per query with global
//     counters, two queries per CTA (measured, not default).
// Which form, CTA size and segment size run is decided in launch_query_it from the counter footprint
// — the L1 that the shared-memory counters leave is what tracks the gathers, DESIGN.md 4 — and the
// batch size; co-resident CTAs of a small shard additionally pull upcoming cells into L2 with bulk
// prefetches and are launched one resident wave at a time so that they sweep the cells in step.
// The threshold pass compacts (count, gid) pairs in gid order into a global pool (one atomicAdd per
// query reserves the segment), or writes the whole counter row (--matrix).  Sorting by (count, gid)
// descending (:685) is done by the host when it merges shards.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "query_common.cuh"

namespace nq {


// id carried by stream slots past the end of a group's stream: lands in spare counter word `lane`
// (shared-memory modes) or is skipped (global mode).  Must fit IT: n <= kMaxCompact in the u16 form.
template <typename IT>
__host__ __device__ __forceinline__ uint32_t query_dummy_id(int mode, uint32_t words, uint32_t lane) {
  return mode == kPack16 ? 2 * (words + lane) : (mode == kSmem32 || mode == kDual16) ? words + lane : (uint32_t)(IT)0xFFFFFFFFu;
}

// Every load of the gather loop is UNCONDITIONAL (dead lanes read a harmless location and their
// result is neutralised where it is consumed, one loop iteration later): ptxas implements a
// predicated load as load-to-temporary + predicated MOV right behind it, which stalls the warp
// for the full memory latency and defeats the software pipeline.
// Directory word of list (cell, fp): {begin,end} as two u16 in a u32 (compact form) or a uint2.
// Kept RAW in registers until the group is processed, one loop iteration after the load.
template <typename IT>
struct DirWord;
template <>
struct DirWord<uint16_t> {
  uint32_t w;
  __device__ __forceinline__ void clear() { w = 0; }
  __device__ __forceinline__ void load(const void* dir, size_t at) { w = __ldg(static_cast<const uint32_t*>(dir) + at); }
  __device__ __forceinline__ void load32(const void* dir, uint32_t at) { w = __ldg(static_cast<const uint32_t*>(dir) + at); }
  __device__ __forceinline__ uint32_t mid() const { return 0; }
  __device__ __forceinline__ uint32_t begin() const { return w & 0xFFFFu; }
  __device__ __forceinline__ uint32_t end() const { return w >> 16; }
};
template <>
struct DirWord<uint32_t> {
  uint32_t x, y;
  __device__ __forceinline__ void clear() { x = y = 0; }
  __device__ __forceinline__ void load(const void* dir, size_t at) {
    const uint2 v = __ldg(static_cast<const uint2*>(dir) + at);
    x = v.x; y = v.y;
  }
  __device__ __forceinline__ void load32(const void* dir, uint32_t at) {
    const uint2 v = __ldg(static_cast<const uint2*>(dir) + at);
    x = v.x; y = v.y;
  }
  __device__ __forceinline__ uint32_t mid() const { return 0; }
  __device__ __forceinline__ uint32_t begin() const { return x; }
  __device__ __forceinline__ uint32_t end() const { return y; }
};

// split16 directory entry {begin, mid, end, 0}
struct DirWord3 {
  uint32_t x, y, z;
  __device__ __forceinline__ void load32(const void* dir, uint32_t at) {
    const uint4 v = __ldg(static_cast<const uint4*>(dir) + at);
    x = v.x; y = v.y; z = v.z;
  }
  __device__ __forceinline__ uint32_t begin() const { return x; }
  __device__ __forceinline__ uint32_t mid() const { return y; }
  __device__ __forceinline__ uint32_t end() const { return z; }
};


// The co-resident query CTAs sweep the cells roughly in step, so each one pulls ITS slice of the
// directory rows and posting arrays of a chunk of cells two chunks ahead of where it is: DRAM then
// sees long sequential bursts (the whole index is read about once per launch) and the random
// probes of every query hit L2.  Purely a hint: a late slice only costs its readers an L2 miss.
constexpr uint32_t kPfCells = 256;  // cells per prefetch chunk
constexpr uint32_t kPfAhead = 2;    // chunks of lead

struct PfSlice {  // this lane's piece of every chunk of the two regions: offset inside the chunk, bytes, chunk pitch
  uint32_t off[2], len[2], pitch[2];
};
__device__ __forceinline__ PfSlice make_pf_slice(const QueryArgs& a, uint32_t elem, uint32_t lane) {
  PfSlice p;
  const uint32_t cells = min(kPfCells, a.F);
  const uint32_t cell_bytes[2] = {a.row_stride * 2 * elem, a.gid_stride * elem};
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const uint32_t size = cells * cell_bytes[r];  // <= 256 cells * 512 KB
    p.pitch[r] = size;
    // the CTA's slice, cut into <= 32 pieces of >= 2 KB (one bulk request per lane)
    const uint32_t per_cta = ((size + gridDim.x - 1) / gridDim.x + 15) & ~15u;
    const uint32_t lo = blockIdx.x * per_cta;
    const uint32_t mine = lo < size ? min(per_cta, size - lo) : 0;
    const uint32_t piece = max(2048u, ((mine + 31) / 32 + 15) & ~15u);
    const uint32_t o = lane * piece;
    p.off[r] = lo + o;
    p.len[r] = o < mine ? (min(piece, mine - o) + 15) & ~15u : 0;
  }
  return p;
}
__device__ __forceinline__ void prefetch_chunk(const QueryArgs& a, const PfSlice& p, uint32_t chunk) {
  if (chunk * kPfCells >= a.F) return;
  if (p.len[0]) l2_prefetch_bulk(static_cast<const char*>(a.dir) + (size_t)chunk * p.pitch[0] + p.off[0], p.len[0]);
  if (p.len[1]) l2_prefetch_bulk(static_cast<const char*>(a.gids) + (size_t)chunk * p.pitch[1] + p.off[1], p.len[1]);
}


// IDX = uint32_t when every posting index F*gid_stride fits 32 bits (always for S <= 15 in the compact form)
template <typename IT, int MODE, int NT, typename IDX, int R = 4, int D = 2>
__global__ void __launch_bounds__(NT, NT == 128 ? 9 : 1) query_count_kernel(QueryArgs a, uint64_t q0) {
  extern __shared__ __align__(16) uint32_t smem[];
  __shared__ IDX s_src[NT / 32][32];  // per warp: stream base of the rank-th non-empty list
  const uint64_t q = q0 + blockIdx.x;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t* cnt = MODE == kGlobal32 ? a.gcounts + (size_t)blockIdx.x * a.n : smem;
  const uint32_t words = MODE == kPack16 ? (a.n + 1) / 2 : a.n;
  constexpr unsigned kFull = 0xFFFFFFFFu;

  PfSlice pf{};
  if (warp == 0 && a.prefetch) {
    pf = make_pf_slice(a, sizeof(IT), lane);
    for (uint32_t ch = 0; ch < a.prefetch; ++ch) prefetch_chunk(a, pf, ch);
  }
  for (uint32_t i = tid; i < words; i += NT) cnt[i] = 0;
  __syncthreads();

  // ---- gather + count (:654-660)
  const int32_t* sk = a.qsk + q * a.F;
  const IT* gids = static_cast<const IT*>(a.gids);
  const unsigned le = 0xFFFFFFFFu >> (31 - lane);  // bits at or below this lane
  uint32_t gathered = 0;
  // count one gathered posting l (stream slot s of `total`)
  // Slots past the end of the stream (last round of a group only) carry a per-lane dummy id that
  // lands in 32 spare words behind the counters, so the shared-memory path needs no branch.
  // (the same ids sit in the 32 elements behind the posting arrays, nq_query_prepare, so that a
  // dead lane's gather is an ordinary load)
  const uint32_t dummy = query_dummy_id<IT>(MODE, words, lane);
  const IDX dead_at = (IDX)a.F * a.gid_stride + lane;
  auto count = [&](uint32_t l) {
    if (MODE == kPack16) atomicAdd(&smem[l >> 1], (l & 1) * 0xFFFFu + 1u);
    else if (MODE == kSmem32) atomicAdd(&smem[l], 1u);
    else if (l != dummy) atomicAdd(&cnt[l], 1u);  // dummy = all ones in IT, never a local id
  };
  // software pipeline over the warp's groups of 32 cells: fingerprints are fetched two groups
  // ahead, directory words one group ahead, so neither latency sits in front of the stream walk
  const uint32_t step = NT;
  uint32_t c_cur = warp * 32;
  uint32_t fp_next = 0xFFFFFFFFu, fp_next2 = 0xFFFFFFFFu;
  if (c_cur + step + lane < a.F) fp_next = (uint32_t)__ldg(&sk[c_cur + step + lane]);
  // directory probe of (cell, fp): cell and fp are clamped into the table; fp >= range (empty or
  // out-of-range fingerprint, :655) is turned into an empty list where the word is unpacked
  auto probe = [&](DirWord<IT>& d, uint32_t cell, uint32_t fp) {
    d.load(a.dir, (size_t)min(cell, a.F - 1) * a.row_stride + min(fp, a.range - 1));
  };
  DirWord<IT> dw, dw_next;
  uint32_t fp_cur = 0xFFFFFFFFu;
  if (c_cur + lane < a.F) fp_cur = (uint32_t)__ldg(&sk[c_cur + lane]);
  probe(dw, c_cur + lane, fp_cur);
  // R = rounds of 32 postings whose gathers issue together (one batch); D = batches in flight
  static_assert(D == 2 || D == 3, "ring depth");
  uint32_t lbuf[D][R];  // ring of gathered batches
  uint32_t live[D];     // rounds each one holds (warp-uniform)
#pragma unroll
  for (int d = 0; d < D; ++d) live[d] = 0;
  uint32_t phase = 0;
  auto drain = [&](uint32_t (&l)[R], uint32_t nl) {
    if (nl >= (uint32_t)R) {
#pragma unroll
      for (int k = 0; k < R; ++k) count(l[k]);
    } else {
#pragma unroll
      for (int k = 0; k < R - 1; ++k)
        if ((uint32_t)k < nl) count(l[k]);
    }
  };
  for (; c_cur < a.F; c_cur += step) {
    const uint32_t cell = c_cur + lane;
    if (warp == 0 && a.prefetch && c_cur % kPfCells == 0) prefetch_chunk(a, pf, c_cur / kPfCells + a.prefetch);
    // prefetch: fingerprint of group +2, directory word of group +1
    fp_next2 = 0xFFFFFFFFu;
    if (c_cur + 2 * step + lane < a.F) fp_next2 = (uint32_t)__ldg(&sk[c_cur + 2 * step + lane]);
    probe(dw_next, cell + step, fp_next);

    const uint32_t b = dw.begin(), len = fp_cur < a.range ? dw.end() - b : 0u;
    uint32_t incl = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(kFull, incl, d);
      if (lane >= d) incl += t;
    }
    const uint32_t excl = incl - len;
    const uint32_t total = __shfl_sync(kFull, incl, 31);
    const unsigned nonempty = __ballot_sync(kFull, len != 0);
    if (len) s_src[warp][__popc(nonempty & (le >> 1))] = (IDX)((IDX)cell * a.gid_stride + b - excl);
    __syncwarp();
    gathered += len;
    // The 32 lists are walked as one concatenated stream, 32 slots per round.  rel = distance from
    // the round's first slot to this lane's list start: a list starts inside the round iff
    // rel < 32 (it wraps to a huge value once the start is behind; empty lists never start).
    uint32_t rel = len ? excl : 0x80000000u;
    const IDX* rank_ptr = &s_src[warp][0] - 1;  // &s_src[warp][lists started before the round - 1]
    uint32_t s = lane;
    // Batches of R rounds: the R owner look-ups and gathers of a batch are independent and issue
    // back to back; the batch gathered before it (possibly by the previous group) is counted while
    // they are in flight.  D register buffers rotate under a warp-uniform branch so that no
    // register copy ever has to wait for a gather.
    // gather NB <= R rounds into l[0..NB) (straight-line code per NB, so the partial last batch of a
    // group costs exactly its rounds)
    auto gather_n = [&](uint32_t (&l)[R], auto nb_tag) {
      constexpr int NB = decltype(nb_tag)::value;
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        const unsigned m = __reduce_or_sync(kFull, shl_clamp(1u, rel));  // list starts inside this round
        const IDX at = rank_ptr[__popc(m & le)] + s + 32 * k;  // owner = last list started at or before s
        l[k] = gids[s + 32 * k < total ? at : dead_at];       // past the end: this lane's dummy id
        rank_ptr += __popc(m);
        rel -= 32;
      }
      s += 32 * NB;
    };
    auto gather = [&](uint32_t (&l)[R], uint32_t nb) {  // nb warp-uniform, 1..R
      if (nb >= (uint32_t)R) gather_n(l, std::integral_constant<int, R>());
      else if (R == 4 && nb == 3) gather_n(l, std::integral_constant<int, (R == 4 ? 3 : 1)>());
      else if (R == 4 && nb == 2) gather_n(l, std::integral_constant<int, (R == 4 ? 2 : 1)>());
      else if (R == 4) gather_n(l, std::integral_constant<int, 1>());
      else {  // deep batches (long lists): the one partial batch of a group goes round by round
#pragma unroll
        for (int k = 0; k < R - 1; ++k)
          if ((uint32_t)k < nb) {
            const unsigned m = __reduce_or_sync(kFull, shl_clamp(1u, rel));
            const IDX at = rank_ptr[__popc(m & le)] + s;
            l[k] = gids[s < total ? at : dead_at];
            rank_ptr += __popc(m);
            rel -= 32;
            s += 32;
          }
      }
    };
    for (uint32_t r0 = 0; r0 < total; r0 += 32 * R) {
      const uint32_t nb = min((uint32_t)R, (total - r0 + 31) >> 5);
      // fill buffer `phase`, count the oldest one ((phase+1) % D, filled D-1 batches ago)
      if (D == 2) {
        if (phase == 0) { gather(lbuf[0], nb); drain(lbuf[1], live[1]); live[0] = nb; }
        else { gather(lbuf[1], nb); drain(lbuf[0], live[0]); live[1] = nb; }
      } else {
        if (phase == 0) { gather(lbuf[0], nb); drain(lbuf[1], live[1]); live[0] = nb; }
        else if (phase == 1) { gather(lbuf[1], nb); drain(lbuf[2], live[2]); live[1] = nb; }
        else { gather(lbuf[2], nb); drain(lbuf[0], live[0]); live[2] = nb; }
      }
      phase = phase + 1 == D ? 0 : phase + 1;
    }
    __syncwarp();
    dw = dw_next;
    fp_cur = fp_next;
    fp_next = fp_next2;
  }
  // the D-1 batches still in flight: every buffer but `phase`, which was drained by the last step
#pragma unroll
  for (int d = 0; d < D; ++d)
    if ((uint32_t)d != phase) drain(lbuf[d], live[d]);
  __syncthreads();

  query_finish<MODE, NT>(a, q, cnt, gathered, 0);
}

// ---- segment-table form --------------------------------------------------------------------------
// The concatenated-stream walk above finds the owner of every stream slot in every round (REDUX +
// 2 POPC + LDS per 32 postings).  Here the lists of a warp's 32 cells are cut into SEGMENTS of SEG
// consecutive postings (SEG = 32: a whole round per segment, for shards whose lists are tens of
// postings long; SEG = 8: four lists side by side per round, for short lists) and the descriptors
// {first posting, postings left in the list} of the group's segments are written, in list order,
// into a small per-warp table in shared memory: each cell's lane writes the segments of its own
// list at the position a warp scan of the segment counts gives it.  The gather of a round is then
// one broadcast LDS.64 of the descriptor + add, compare, select, address, load — no look-up — and
// rounds are independent of each other, so the R gathers of a batch issue back to back.  Lists
// longer than the table are taken in windows of T segments.  Same register ring of two R-round
// batches as the stream form (the previous batch is counted while this one's loads fly), same
// dummy-id trick for the dead lanes of a list's last segment.
template <typename IDX>
struct SegRef {
  IDX at;       // index of the segment's first posting in gids[]
  int32_t rem;  // postings of the list from there on (lanes at or past it are dead)
};

// MODE kDual16 (small shards, S <= 15): the CTA counts TWO queries, warps [0, NW/2) the first and
// [NW/2, NW) the second, into one u32 word per genome — the first query owns the low half-word, the
// second the high one (a count never exceeds F <= 32768, so the low half cannot carry).  Counting a
// posting is then `atomicAdd(&cnt[id], inc)` with a per-warp constant: shift + ATOMS instead of the
// five ALU instructions that pick the half-word of a packed pair.
// D = batches in the register ring: the batch gathered D-1 batches ago is counted while the newer
// ones fly (D = 3 for the 128/256-thread forms: with D = 2 only one batch's issue time, less than
// an L2 hit, lies between a gather and its use — 31 % of the stall samples in r01 ncu).
// SPLIT (shards of 65.6k..131k genomes, SEG = 32, packed counters): postings come from the u16 copy
// `gids16` and the directory from `dir3` {begin, mid, end}: ids from position `mid` of a list on
// are >= 65536 and get the 2^16 back when they are counted.  Which lanes of a round lie past `mid`
// is known at gather time only, so one bit per round is kept beside each ring slot (`hmask`).  A
// segment descriptor carries {postings left, min(postings left, ids below 2^16 left)} as two
// half-words: lanes at or past the second number are "high", which includes the dead lanes of a
// list's last round — their dummy id is stored minus 2^16.
template <typename IT, int MODE, int NT, typename IDX, int SEG, int T, int R, int D, bool SPLIT = false>
__global__ void __launch_bounds__(NT, NT == 128 ? 8 : NT == 256 ? 4 : (NT == 512 && MODE != kGlobal32) ? 2 : 1)
query_count_seg_kernel(QueryArgs a, uint64_t q0) {
  static_assert(!SPLIT || (SEG == 32 && MODE == kPack16 && sizeof(IT) == 2 && sizeof(IDX) == 4), "split16 form");
  constexpr int SPR = 32 / SEG;  // segments per round
  constexpr bool DUAL = MODE == kDual16;
  constexpr int NW = NT / 32, NWQ = DUAL ? NW / 2 : NW;  // warps per query
  static_assert(SEG == 8 || SEG == 16 || SEG == 32, "segment size");
  static_assert(T % SPR == 0, "table = whole rounds");
  static_assert(D >= 2 && D <= 4, "ring depth");
  extern __shared__ __align__(16) uint32_t smem[];
  __shared__ __align__(16) SegRef<IDX> s_tab[NW][T + SPR];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t half = DUAL ? warp / NWQ : 0, qwarp = DUAL ? warp % NWQ : warp;
  const uint64_t qbase = q0 + (uint64_t)blockIdx.x * (DUAL ? 2 : 1);
  const uint64_t q = qbase + half;
  const bool have_q = !DUAL || q < a.q_end;
  uint32_t* cnt = MODE == kGlobal32 ? a.gcounts + (size_t)blockIdx.x * a.n : smem;
  const uint32_t words = MODE == kPack16 ? (a.n + 1) / 2 : a.n;
  constexpr unsigned kFull = 0xFFFFFFFFu;

  PfSlice pf{};
  if (warp == 0 && a.prefetch) {
    pf = make_pf_slice(a, sizeof(IT), lane);
    for (uint32_t ch = 0; ch < a.prefetch; ++ch) prefetch_chunk(a, pf, ch);
  }
  const bool shared_q = MODE == kGlobal32 && a.parts != 0;  // counters zeroed by the host, finished by query_finish_kernel
  if (!shared_q)
    for (uint32_t i = tid; i < words; i += NT) cnt[i] = 0;
  __syncthreads();

  const int32_t* sk = a.qsk + (have_q ? q : qbase) * a.F;
  const IT* gids = SPLIT ? reinterpret_cast<const IT*>(a.gids16) : static_cast<const IT*>(a.gids);
  SegRef<IDX>* tab = s_tab[warp];
  const uint32_t sub = lane & (SEG - 1), grp = lane / SEG;
  uint32_t gathered = 0;
  const uint32_t dummy = query_dummy_id<IT>(MODE, words, lane);
  const IDX dead_at = (IDX)a.F * a.gid_stride + (DUAL ? 32 : 0) + lane;
  const uint32_t inc = half ? 0x10000u : 1u;
  auto count = [&](uint32_t l) {
    if (DUAL) atomicAdd(&smem[l], inc);
    else if (MODE == kPack16) atomicAdd(&smem[l >> 1], (l & 1) * 0xFFFFu + 1u);
    else if (MODE == kSmem32) atomicAdd(&smem[l], 1u);
    else if (l != dummy) atomicAdd(&cnt[l], 1u);
  };
  // Cells per sweep.  GLOBAL mode with `parts`: the gridDim.y CTAs of a query form one pool of
  // virtual warps; with more warps than groups of 32 cells (reads as entries: 256 cells, lists of
  // 10^5 postings) every list is cut into `pieces` and a warp takes one piece of its group's lists.
  uint32_t step = NWQ * 32, c_cur = qwarp * 32, piece = 0, pieces = 1;
  uint32_t F = have_q ? a.F : 0;  // a CTA's missing second query walks no cells
  if (MODE == kGlobal32 && a.parts) {
    const uint32_t vw = blockIdx.y * NW + warp, VW = gridDim.y * NW, ngroups = (a.F + 31) / 32;
    if (VW <= ngroups) {
      c_cur = vw * 32; step = VW * 32;
    } else {
      pieces = VW / ngroups; piece = vw / ngroups;
      c_cur = (vw % ngroups) * 32; step = ngroups * 32;  // one group per warp
      if (piece >= pieces) F = 0;
    }
  }
  uint32_t fp_next = 0xFFFFFFFFu, fp_next2 = 0xFFFFFFFFu;
  if (c_cur + step + lane < F) fp_next = (uint32_t)__ldg(&sk[c_cur + step + lane]);
  // directory probe with 32-bit element indexes (the host checks F * row_stride < 2^32)
  typedef typename std::conditional<SPLIT, DirWord3, DirWord<IT>>::type DW;
  auto probe = [&](DW& d, uint32_t cell, uint32_t fp) {
    d.load32(SPLIT ? static_cast<const void*>(a.dir3) : a.dir, min(cell, a.F - 1) * a.row_stride + min(fp, a.range - 1));
  };
  DW dw, dw_next;
  uint32_t fp_cur = 0xFFFFFFFFu;
  if (c_cur + lane < F) fp_cur = (uint32_t)__ldg(&sk[c_cur + lane]);
  probe(dw, c_cur + lane, fp_cur);

  uint32_t lbuf[D][R];
  uint32_t live[D], hmask[D];
#pragma unroll
  for (int d = 0; d < D; ++d) live[d] = hmask[d] = 0;
  uint32_t phase = 0;
  // (split16: AND + multiply-add forms of this count on the FMA pipe, and a validity test on the
  // packed descriptor word, measured 6 % slower than the plain form — 34.6 vs 32.6 ms)
  auto drain = [&](uint32_t (&l)[R], uint32_t hm, uint32_t nl) {
    if (nl >= (uint32_t)R) {
#pragma unroll
      for (int k = 0; k < R; ++k) count(SPLIT ? l[k] + (((hm >> k) & 1u) << 16) : l[k]);
    } else {
#pragma unroll
      for (int k = 0; k < R - 1; ++k)
        if ((uint32_t)k < nl) count(SPLIT ? l[k] + (((hm >> k) & 1u) << 16) : l[k]);
    }
  };
  auto gather_round = [&](uint32_t& dst, uint32_t& hm, const SegRef<IDX>* t, int k) {
    const SegRef<IDX> d = t[k * SPR + grp];
    if (SPLIT) {
      const uint32_t rem = (uint32_t)d.rem & 0xFFFFu, below = (uint32_t)d.rem >> 16;
      dst = gids[sub < rem ? d.at + sub : dead_at];
      hm |= (sub >= below ? 1u : 0u) << k;
    } else {
      dst = gids[(int32_t)sub < d.rem ? d.at + sub : dead_at];
    }
  };
  // gather nb (1..R, warp-uniform) rounds starting at table position t0 (a multiple of SPR)
  auto gather = [&](uint32_t (&l)[R], uint32_t& hm, const SegRef<IDX>* t0, uint32_t nb) {
    hm = 0;
    if (nb >= (uint32_t)R) {
#pragma unroll
      for (int k = 0; k < R; ++k) gather_round(l[k], hm, t0, k);
    } else {
#pragma unroll
      for (int k = 0; k < R - 1; ++k)
        if ((uint32_t)k < nb) gather_round(l[k], hm, t0, k);
    }
  };
  // one batch: fill ring slot `phase`, count the oldest slot (phase+1 mod D, filled D-1 batches ago)
  auto batch = [&](const SegRef<IDX>* t0, uint32_t nb) {
#pragma unroll
    for (int p = 0; p < D; ++p)
      if (phase == (uint32_t)p) {
        gather(lbuf[p], hmask[p], t0, nb);
        drain(lbuf[(p + 1) % D], hmask[(p + 1) % D], live[(p + 1) % D]);
        live[(p + 1) % D] = 0;
        live[p] = nb;
      }
    phase = phase + 1 == (uint32_t)D ? 0 : phase + 1;
  };
  for (; c_cur < F; c_cur += step) {
    const uint32_t cell = c_cur + lane;
    if (warp == 0 && a.prefetch && c_cur % kPfCells == 0) prefetch_chunk(a, pf, c_cur / kPfCells + a.prefetch);
    fp_next2 = 0xFFFFFFFFu;
    if (c_cur + 2 * step + lane < F) fp_next2 = (uint32_t)__ldg(&sk[c_cur + 2 * step + lane]);
    probe(dw_next, cell + step, fp_next);

    uint32_t b = dw.begin(), len = fp_cur < a.range ? dw.end() - b : 0u;
    if (MODE == kGlobal32 && pieces > 1) {  // this warp's piece of the list
      const uint32_t lo = (uint32_t)((uint64_t)len * piece / pieces), hi = (uint32_t)((uint64_t)len * (piece + 1) / pieces);
      b += lo; len = hi - lo;
    }
    const uint32_t low = SPLIT ? dw.mid() - b : 0u;  // ids below 2^16 at the head of the list
    // descriptor of the segment that starts k postings into the list
    auto seg_ref = [&](IDX at, uint32_t k) {
      if constexpr (!SPLIT) {
        return SegRef<IDX>{at, (int32_t)(len - k)};
      } else {
        const uint32_t rem = min(len - k, 0xFFFFu), below = low > k ? min(low - k, rem) : 0u;
        return SegRef<IDX>{at, (int32_t)(rem | (below << 16))};
      }
    };
    const uint32_t nseg = (len + SEG - 1) / SEG;
    uint32_t incl = nseg;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(kFull, incl, d);
      if (lane >= d) incl += t;
    }
    const uint32_t excl = incl - nseg;
    const uint32_t total = __shfl_sync(kFull, incl, 31);
    const IDX first = (IDX)cell * a.gid_stride + b;
    gathered += len;

    if (total <= (uint32_t)T) {
      // the whole group fits the table (the usual case): straight-line stores for the first two
      // segments of a list, a loop only for longer ones
      if (nseg) tab[excl] = seg_ref(first, 0);
      if (nseg > 1) tab[excl + 1] = seg_ref((IDX)(first + SEG), SEG);
#pragma unroll 1
      for (uint32_t sg = 2; sg < nseg; ++sg) tab[excl + sg] = seg_ref((IDX)(first + sg * SEG), sg * SEG);
      if (SPR > 1 && lane < SPR) tab[total + lane] = SegRef<IDX>{0, 0};  // dead slots of the last round
      __syncwarp();
      const uint32_t rounds = (total + SPR - 1) / SPR;
      uint32_t r0 = 0;
      for (; r0 + R <= rounds; r0 += R) batch(tab + r0 * SPR, R);
      if (r0 < rounds) batch(tab + r0 * SPR, rounds - r0);
      __syncwarp();  // the table is rewritten by the next group
    } else {
      for (uint32_t w0 = 0; w0 < total; w0 += T) {  // windows of T segments
        const uint32_t nwin = min((uint32_t)T, total - w0);
        const uint32_t lo = max(excl, w0), hi = min(incl, w0 + T);  // this lane's segments inside the window
#pragma unroll 1
        for (uint32_t sg = lo; sg < hi; ++sg) {
          const uint32_t k = (sg - excl) * SEG;
          tab[sg - w0] = seg_ref((IDX)(first + k), k);
        }
        if (SPR > 1 && lane < SPR) tab[nwin + lane] = SegRef<IDX>{0, 0};
        __syncwarp();
        const uint32_t rounds = (nwin + SPR - 1) / SPR;
        for (uint32_t r0 = 0; r0 < rounds; r0 += R) batch(tab + r0 * SPR, min((uint32_t)R, rounds - r0));
        __syncwarp();
      }
    }
    dw = dw_next;
    fp_cur = fp_next;
    fp_next = fp_next2;
  }
  // the D-1 batches still in flight, oldest first
#pragma unroll
  for (int i = 1; i < D; ++i) {
#pragma unroll
    for (int p = 0; p < D; ++p)
      if (phase == (uint32_t)p) drain(lbuf[(p + i) % D], hmask[(p + i) % D], live[(p + i) % D]);
  }
  __syncthreads();
  if (shared_q) {  // only the gather statistics; the counters are complete when the whole grid is
#pragma unroll
    for (int d = 16; d; d >>= 1) gathered += __shfl_xor_sync(kFull, gathered, d);
    if (lane == 0 && gathered) atomicAdd(a.cursor + 1, (unsigned long long)gathered);
    return;
  }
  if (DUAL) {
    query_finish<MODE, NT>(a, qbase, cnt, gathered, 0);
    if (qbase + 1 < a.q_end) {
      __syncthreads();
      query_finish<MODE, NT>(a, qbase + 1, cnt, 0, 16);
    }
  } else {
    query_finish<MODE, NT>(a, q, cnt, gathered, 0);
  }
}

// ---- threshold + compaction (or the dense row) of queries whose counters were filled by a whole
// grid.  n is in the millions here, so the genomes of a query are cut into gridDim.y slices:
//   global_hits_count_kernel   hits per (query, slice)
//   global_hits_reserve_kernel per query: exclusive offsets of its slices + one pool reservation
//   global_hits_write_kernel   (count, gid) in gid order inside each slice, slices in order
// (or global_dense_kernel alone when the whole counter row is wanted).
__device__ __forceinline__ void slice_of(const QueryArgs& a, uint32_t& g_lo, uint32_t& g_hi) {
  const uint32_t per = (a.n + gridDim.y - 1) / gridDim.y;
  g_lo = min(a.n, blockIdx.y * per);
  g_hi = min(a.n, g_lo + per);
}
template <int NT>
__global__ void __launch_bounds__(NT) global_hits_count_kernel(QueryArgs a, uint32_t* __restrict__ slice_hits) {
  __shared__ uint32_t s_warp[NT / 32];
  const uint32_t* cnt = a.gcounts + (size_t)blockIdx.x * a.n;
  uint32_t g_lo, g_hi;
  slice_of(a, g_lo, g_hi);
  uint32_t mine = 0;
  for (uint32_t g = g_lo + threadIdx.x; g < g_hi; g += NT) mine += (__ldcg(&cnt[g]) & a.wrap_mask) >= a.min_score;
  mine = __reduce_add_sync(0xFFFFFFFFu, mine);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < NT / 32; ++w) t += s_warp[w];
    slice_hits[(size_t)blockIdx.x * gridDim.y + blockIdx.y] = t;
  }
}
__global__ void global_hits_reserve_kernel(QueryArgs a, uint64_t q0, uint32_t nb, uint32_t slices, uint32_t* __restrict__ slice_hits,
                                           unsigned long long* __restrict__ slice_base) {
  const uint32_t ql = blockIdx.x * blockDim.x + threadIdx.x;
  if (ql >= nb) return;
  uint32_t total = 0;
  for (uint32_t s = 0; s < slices; ++s) total += slice_hits[(size_t)ql * slices + s];
  unsigned long long base = atomicAdd(a.cursor, (unsigned long long)total);
  a.hit_begin[q0 + ql] = base;
  a.hit_n[q0 + ql] = total;
  const bool fits = base + total <= a.pool_cap;  // overflow: nothing is written, the host re-runs with a larger pool
  for (uint32_t s = 0; s < slices; ++s) {
    slice_base[(size_t)ql * slices + s] = fits ? base : ~0ull;
    base += slice_hits[(size_t)ql * slices + s];
  }
}
template <int NT>
__global__ void __launch_bounds__(NT) global_hits_write_kernel(QueryArgs a, const uint32_t* __restrict__ slice_hits,
                                                               const unsigned long long* __restrict__ slice_base) {
  __shared__ uint32_t s_warp[NT / 32];
  const size_t sl = (size_t)blockIdx.x * gridDim.y + blockIdx.y;
  const unsigned long long base = slice_base[sl];
  if (slice_hits[sl] == 0 || base == ~0ull) return;
  const uint32_t* cnt = a.gcounts + (size_t)blockIdx.x * a.n;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t g_lo, g_hi;
  slice_of(a, g_lo, g_hi);
  uint32_t done = 0;
  for (uint32_t g0 = g_lo; g0 < g_hi; g0 += NT) {
    const uint32_t g = g0 + tid;
    const uint32_t c = g < g_hi ? __ldcg(&cnt[g]) & a.wrap_mask : 0u;
    const bool hit = g < g_hi && c >= a.min_score;
    if (__syncthreads_count(hit) == 0) continue;
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, hit);
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
    __syncthreads();
  }
}
template <int NT>
__global__ void __launch_bounds__(NT) global_dense_kernel(QueryArgs a, uint64_t q0) {
  const uint32_t* cnt = a.gcounts + (size_t)blockIdx.x * a.n;
  uint32_t* row = a.dense + (q0 + blockIdx.x) * a.n;
  uint32_t g_lo, g_hi;
  slice_of(a, g_lo, g_hi);
  for (uint32_t g = g_lo + threadIdx.x; g < g_hi; g += NT) row[g] = __ldcg(&cnt[g]) & a.wrap_mask;
}

}  // namespace nq

using namespace nq;

template <typename IT, int MODE, int NT>
static cudaError_t launch_query_t(size_t smem, unsigned nb, const QueryArgs& a, uint64_t q0, cudaStream_t st, int* occ) {
  const bool idx32 = (uint64_t)a.F * a.gid_stride + kQuerySlack < (1ull << 32);
  // one big CTA per SM (many genomes, long lists, HBM-bound): 8 gathers per batch for bytes in flight
  constexpr int R = NT == 1024 ? 8 : 4;
  auto k32 = query_count_kernel<IT, MODE, NT, uint32_t, R>;
  auto k64 = query_count_kernel<IT, MODE, NT, uint64_t, R>;
  cudaError_t e = cudaFuncSetAttribute(idx32 ? k32 : k64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (occ)  // resident CTAs per SM of the kernel that would run (wave sizing), no launch
    return idx32 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k32, NT, smem)
                 : cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k64, NT, smem);
  if (idx32) k32<<<nb, NT, smem, st>>>(a, q0);
  else k64<<<nb, NT, smem, st>>>(a, q0);
  return cudaSuccess;
}
template <typename IT, int MODE, int NT, int SEG, int T, bool SPLIT = false>
static cudaError_t launch_query_seg_t(size_t smem, unsigned nb, const QueryArgs& a, uint64_t q0, cudaStream_t st, int* occ) {
  const bool idx32 = (uint64_t)a.F * a.gid_stride + kQuerySlack < (1ull << 32);
  constexpr int R = NT <= 256 ? 4 : 8;
  constexpr int D = NT <= 256 ? 3 : 2;  // registers: 64 per thread at 1024 and at 2 x 512 threads per SM
  constexpr unsigned QPC = MODE == kDual16 ? 2 : 1;  // queries per CTA
  // 64-bit posting indexes only occur with global counters (n > 131k at S=15), where shared memory is free
  auto k32 = query_count_seg_kernel<IT, MODE, NT, uint32_t, SEG, T, R, D, SPLIT>;
  // 16-byte descriptors: half the window (the split16 form exists with 32-bit posting indexes only)
  auto k64 = query_count_seg_kernel<IT, MODE, NT, typename std::conditional<SPLIT, uint32_t, uint64_t>::type, SEG, SPLIT ? T : T / 2, R, D, SPLIT>;
  // the segment table is static shared memory on top of the counters: cudaErrorInvalidValue here
  // (counters + table over the per-block limit) sends the caller back to the stream form
  cudaError_t e = cudaFuncSetAttribute(idx32 ? k32 : k64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (occ) {  // resident QUERIES per SM of the kernel that would run (wave sizing), no launch
    e = idx32 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k32, NT, smem)
              : cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k64, NT, smem);
    *occ *= QPC;
    return e;
  }
  QueryArgs b = a;
  b.q_end = q0 + nb;
  const dim3 grid((nb + QPC - 1) / QPC, MODE == kGlobal32 && a.parts ? a.parts : 1);
  if (MODE == kGlobal32 && a.parts &&
      (e = cudaMemsetAsync(a.gcounts, 0, (size_t)nb * a.n * sizeof(uint32_t), st)) != cudaSuccess)
    return e;
  if (idx32) k32<<<grid, NT, smem, st>>>(b, q0);
  else k64<<<grid, NT, smem, st>>>(b, q0);
  if (MODE == kGlobal32 && a.parts) {
    const dim3 fg(nb, a.slices);
    if (a.dense) {
      global_dense_kernel<1024><<<fg, 1024, 0, st>>>(b, q0);
    } else {
      global_hits_count_kernel<1024><<<fg, 1024, 0, st>>>(b, a.slice_hits);
      global_hits_reserve_kernel<<<(nb + 127) / 128, 128, 0, st>>>(b, q0, nb, a.slices, a.slice_hits, a.slice_base);
      global_hits_write_kernel<1024><<<fg, 1024, 0, st>>>(b, a.slice_hits, a.slice_base);
    }
  }
  return cudaSuccess;
}
// Gather form.  Lists of a shard of n genomes hold ~n * 6.8e-4 postings on bacterial sketches at the
// default W.  NQ_QUERY_FORM = stream | seg8 | seg16 | seg32 | dual8 | dual16 overrides
// (measurement only).
enum QueryForm { kFormStream = 0, kFormSeg8 = 8, kFormSeg16 = 16, kFormSeg32 = 32, kFormDual8 = 108, kFormDual16 = 116 };
// the two-queries-per-CTA counters: S <= 15 (half-word counts), u16 ids that leave room for the
// dummy ids, and four 256-thread CTAs per SM
static bool query_dual_ok(const nq_index* ix) {
  return ix->p.S <= 15 && ix->elem == 2 && (size_t)ix->n * 4 + 128 <= 54 * 1024 && ix->n + 64 < 65536;
}
static int query_form(const nq_index* ix, bool small) {
  static const char* env = nq_tuning_env("NQ_QUERY_FORM");
  const bool seg_ok = (uint64_t)ix->p.F * ix->row_stride < (1ull << 32);  // 32-bit directory indexes
  if (env && seg_ok) {
    if (!strcmp(env, "stream")) return kFormStream;
    if (!strcmp(env, "seg8")) return kFormSeg8;
    if (!strcmp(env, "seg16")) return kFormSeg16;
    if (!strcmp(env, "seg32")) return kFormSeg32;
    if (!strcmp(env, "dual8") && query_dual_ok(ix)) return kFormDual8;
    if (!strcmp(env, "dual16") && query_dual_ok(ix)) return kFormDual16;
  }
  if (!seg_ok) return kFormStream;
  // small shards: seg8 with the three-deep ring.  stream / dual8 / seg16 measure within 3 % of it on a box
  // where the L2 prefetch works as intended (0.63-0.65 ms at configs[1]); the deeper ring is the safer choice
  // when it does not (one B200 of the pool ran the stream form at 1.04 ms, its lone-CTA time)
  if (small) return kFormSeg8;
  return ix->n >= 40000 ? kFormSeg32 : kFormStream;
}
template <typename IT>
static cudaError_t launch_query_it(const nq_index* ix, int mode, size_t smem, unsigned nb, const QueryArgs& a, uint64_t q0,
                                   cudaStream_t st, int* occ) {
  // small counter arrays: 128-thread CTAs (4 warps with ~56 registers each carry a deep gather
  // pipeline, and ~9 queries share an SM); large ones: one big CTA per SM
  const bool small = mode != kGlobal32 && smem <= 26 * 1024;  // at least 8 such CTAs per SM
  const int form = query_form(ix, small);
  if (mode == kGlobal32) {
    if (form == kFormStream || !a.parts) return launch_query_t<IT, kGlobal32, 512>(0, nb, a, q0, st, occ);
    return launch_query_seg_t<IT, kGlobal32, 512, 32, 96>(0, nb, a, q0, st, occ);
  }
#ifdef NQ_TUNING  // two-queries-per-CTA CSR forms: measured within 2 % of seg8 (r01), measurement builds only
  if (form == kFormDual8 || form == kFormDual16) {
    const size_t dsmem = (size_t)ix->n * 4 + 128;  // one u32 per genome + the dummy words
    if (form == kFormDual8) return launch_query_seg_t<IT, kDual16, 256, 8, 64>(dsmem, nb, a, q0, st, occ);
    return launch_query_seg_t<IT, kDual16, 256, 16, 64>(dsmem, nb, a, q0, st, occ);
  }
#endif
  // shards between "small" and "one CTA per SM": four 256-thread CTAs per SM while their counters + tables
  // leave >= 40 KB of L1 (n <= ~20k packed), two 512-thread CTAs while they leave >= 48 KB (n <= ~38k).  10k
  // queries vs 16k / 20k / 25k / 35k genomes: 12.9 / ~14 / 15.5 / 18.3 ms as one 1024-thread CTA per SM ->
  // 8.1 / 10.5 / 13.3 / 16.6 ms.  Sized WITHOUT regard for L1 (256 threads up to 51 KB, 512 up to 100 KB) the
  // same kernels were slower than the 1024-thread form (16.6 vs 15.4 ms at 25k, 21.6 vs 17.2 ms at 50k).
  static const char* mid_env = nq_tuning_env("NQ_QUERY_MID");  // "0": the 1024-thread forms (measurement only)
  const bool mid_on = !(mid_env && mid_env[0] == '0');
  const bool mid256 = mid_on && !small && 4 * (smem + 1024 + 4400) + 40 * 1024 <= 228 * 1024;
  const bool mid512 = mid_on && !small && !mid256 && 2 * (smem + 1024 + 12600) + 48 * 1024 <= 228 * 1024;
  // small shards: 128-thread CTAs, 8-9 queries per SM — unless their counters leave the SM less than ~24 KB of
  // L1 to track the gathers (12.5k genomes: 8 x 26.9 KB of shared memory), or the batch is many waves long;
  // then 256-thread CTAs (4 queries per SM, 8 warps each: the same 32 warps on half the shared memory).
  // 12.5k genomes x 10k queries: 11.8 -> 6.65 ms; 10k x 10k: 6.63 -> 5.94 ms; but 10k x 2k: 1.29 -> 1.40 ms and
  // 10k x 1k: 0.64 -> 0.70 ms (half-empty waves), which stay on 128 threads.  NQ_QUERY_NT overrides.
  static const char* nt_env = nq_tuning_env("NQ_QUERY_NT");
  const size_t cta128 = smem + 1024 + 800;  // dynamic + per-CTA reserve + static tables
  const size_t occ128 = std::min<size_t>(8, (228 * 1024) / cta128);
  const bool l1_starved = 228 * 1024 - occ128 * cta128 < 24 * 1024;
  const bool nt256 = nt_env ? atoi(nt_env) == 256 : (l1_starved || a.nq_total >= (uint64_t)ix->ctx->sm_count * 36);
#ifdef NQ_TUNING  // seg16 / seg32 / stream at 128 threads: selectable in measurement builds only
#define NQ_SMALL_OTHER_FORMS(MODE)                                                                                 \
    if (form == kFormSeg16) return launch_query_seg_t<IT, MODE, 128, 16, 64>(smem, nb, a, q0, st, occ);            \
    if (form == kFormSeg32) return launch_query_seg_t<IT, MODE, 128, 32, 64>(smem, nb, a, q0, st, occ);            \
    if (form == kFormStream && seg_ok_) return launch_query_t<IT, MODE, 128>(smem, nb, a, q0, st, occ);
#else
#define NQ_SMALL_OTHER_FORMS(MODE)
#endif
  const bool seg_ok_ = (uint64_t)ix->p.F * ix->row_stride < (1ull << 32);
  (void)seg_ok_;
#define NQ_SEG_DISPATCH(MODE)                                                                                      \
  if (small) {                                                                                                     \
    NQ_SMALL_OTHER_FORMS(MODE)                                                                                     \
    if (form != kFormStream && nt256) return launch_query_seg_t<IT, MODE, 256, 8, 64>(smem, nb, a, q0, st, occ);   \
    if (form != kFormStream) return launch_query_seg_t<IT, MODE, 128, 8, 64>(smem, nb, a, q0, st, occ);            \
    return launch_query_t<IT, MODE, 128>(smem, nb, a, q0, st, occ); /* 64-bit directory indexes */                 \
  }                                                                                                                \
  if (mid256) return launch_query_seg_t<IT, MODE, 256, 16, 64>(smem, nb, a, q0, st, occ);                          \
  if (mid512) return launch_query_seg_t<IT, MODE, 512, 16, 96>(smem, nb, a, q0, st, occ);                          \
  if (form == kFormSeg32) return launch_query_seg_t<IT, MODE, 1024, 32, 96>(smem, nb, a, q0, st, occ);             \
  return launch_query_t<IT, MODE, 1024>(smem, nb, a, q0, st, occ);
  // 65.6k..131k genomes with packed counters: gather from the u16 copy of the postings
  if (mode == kPack16 && form == kFormSeg32 && a.gids16 && (uint64_t)a.F * a.gid_stride + kQuerySlack < (1ull << 32))
    return launch_query_seg_t<uint16_t, kPack16, 1024, 32, 96, true>(smem, nb, a, q0, st, occ);
  if (mode == kPack16) { NQ_SEG_DISPATCH(kPack16) }
  NQ_SEG_DISPATCH(kSmem32)
#undef NQ_SEG_DISPATCH
#undef NQ_SMALL_OTHER_FORMS
}
// slab.cu
bool nq_slab_layout(const nq_index* ix, int& mode, size_t& smem);
cudaError_t nq_slab_launch(const nq_index* ix, int mode, size_t smem, unsigned nb, const QueryArgs& a, uint64_t q0,
                           cudaStream_t st, int* occ);

static cudaError_t launch_query(const nq_index* ix, int mode, size_t smem, unsigned nb, const QueryArgs& a, uint64_t q0,
                                cudaStream_t st, int* occ = nullptr) {
  if (a.slab) return nq_slab_launch(ix, mode, smem, nb, a, q0, st, occ);
  return ix->elem == 2 ? launch_query_it<uint16_t>(ix, mode, smem, nb, a, q0, st, occ)
                       : launch_query_it<uint32_t>(ix, mode, smem, nb, a, q0, st, occ);
}

// Chunks of lead of the cooperative L2 prefetch (0 = off): the window of lead + 1 chunks of kPfCells cells
// (directory rows + posting arrays) must sit in L2.  NQ_QUERY_PF_AHEAD overrides the lead (measurement only).
static uint32_t query_prefetch_lead(const nq_index* ix, bool slab) {
  static const char* env = nq_tuning_env("NQ_QUERY_PF_AHEAD");
  // (slab form: off by default — the CTAs of a launch drift apart by more cells than L2 holds, and
  // the gathers of the laggards then pay for the prefetch AND their own misses: 0.77 vs 0.68 ms)
  const uint32_t lead = env ? (uint32_t)atoi(env) : slab ? 0u : kPfAhead;
  const uint64_t chunk = slab ? (uint64_t)kPfCells * (ix->p.range / 32) * 16 + ix->slab_granules * ix->slab_G * 2 / ix->p.F * kPfCells
                              : (uint64_t)kPfCells * ((uint64_t)ix->row_stride * 2 + ix->gid_stride) * ix->elem;
  return (lead + 1) * chunk <= (64ull << 20) ? lead : 0u;
}

// Queries per launch.  Global counters bound it by memory.  When the L2 prefetch is on (a few
// chunks of cells fit in L2: the co-resident CTAs share what each of them pulls in, as long as
// they sweep the cells in step), more queries than CTA slots are cut into equal launches of at most
// one resident wave: CTAs of one launch start together and stay in step, whereas a single grid
// refills slots one by one and ends up with every CTA at a different cell, i.e. random DRAM
// access for two dependent sectors per probe (measured at 12.5k genomes x 10k queries: 13.2 ms
// as one grid).
static uint64_t query_wave(const nq_index* ix, int mode, size_t smem, const QueryArgs& a, uint64_t nq) {
  // global counters: as many queries per launch as keep their counters in L2 (the count kernel's
  // atomics run ~6x faster there than in HBM: 193 vs 33 G/s measured); the grid is filled by cutting
  // every query over `parts` CTAs instead (set_query_parts)
  if (mode == kGlobal32) return std::max<uint64_t>(1, std::min<uint64_t>(nq, (64ull << 20) / ((uint64_t)ix->n * 4)));
  static const char* env = nq_tuning_env("NQ_QUERY_WAVES");  // "0": one grid (measurement only)
  // (the slab form is always cut into waves: its descriptor scratch is sized by the launch and stays in L2)
  if (!a.slab && (!a.prefetch || (env && env[0] == '0'))) return nq;
  int occ = 0;
  if (launch_query(ix, mode, smem, 1, a, 0, nullptr, &occ) != cudaSuccess || occ <= 0) return nq;
  static const char* cap_env = nq_tuning_env("NQ_QUERY_WAVE_SM");  // cap of resident queries per SM in a wave (measurement only)
  if (cap_env && atoi(cap_env) > 0) occ = std::min(occ, atoi(cap_env));
  const uint64_t slots = (uint64_t)occ * ix->ctx->sm_count;
  if (nq <= slots) return nq;
  const uint64_t waves = (nq + slots - 1) / slots;
  return (nq + waves - 1) / waves;
}

// CTAs per query of the global-counter form: two resident waves of 512-thread CTAs over the launch
static int set_query_parts(const nq_index* ix, int mode, uint64_t q_per_launch, QueryArgs& a) {
  a.parts = 0;
  if (mode != kGlobal32 || query_form(ix, false) == kFormStream) return NQ_OK;
  const uint64_t ctas = (uint64_t)ix->ctx->sm_count * 4 * 2;
  a.parts = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(2048, ctas / std::max<uint64_t>(1, q_per_launch)));
  a.slices = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(1024, ((uint64_t)ix->n + 65535) / 65536));
  NQ_TRY(nq_dmalloc(ix->ctx, (void**)&a.slice_hits, q_per_launch * a.slices * sizeof(uint32_t)));
  NQ_TRY(nq_dmalloc(ix->ctx, (void**)&a.slice_base, q_per_launch * a.slices * sizeof(unsigned long long)));
  return NQ_OK;
}

// counter mode and dynamic shared memory of the query kernel for this index
static void query_layout(const nq_index* ix, int& mode, size_t& smem) {
  // counters + 32 spare words (dummy targets of the branch-free count) + the kernel's static shared memory
  const size_t spare = 128, fixed = 26 * 1024, optin = ix->ctx->smem_optin;
  const size_t pack = (size_t)((ix->n + 1) / 2) * 4, full = (size_t)ix->n * 4;
  if (ix->p.S <= 15 && pack + spare + fixed <= optin) { mode = kPack16; smem = pack + spare; }
  else if (full + spare + fixed <= optin) { mode = kSmem32; smem = full + spare; }
  else { mode = kGlobal32; smem = 0; }
}

// Writes the 32 dummy ids behind the posting arrays (kQuerySlack elements are reserved there).
int nq_query_prepare(nq_index* ix) {
  int mode;
  size_t smem;
  query_layout(ix, mode, smem);
  const uint32_t words = mode == kPack16 ? (ix->n + 1) / 2 : ix->n;
  uint16_t h16[64];
  uint32_t h32[64];
  for (uint32_t l = 0; l < 32; ++l) {
    h32[l] = query_dummy_id<uint32_t>(mode, words, l);
    h16[l] = (uint16_t)query_dummy_id<uint16_t>(mode, words, l);
    if (ix->elem == 2 && mode != kGlobal32 && h32[l] > 0xFFFFu) return nq_set_error(NQ_ERR_INVALID, "dummy id overflow");
    // second set (slack elements 32..63): the two-queries-per-CTA form counts into word `id`
    h32[32 + l] = query_dummy_id<uint32_t>(kDual16, ix->n, l);
    h16[32 + l] = (uint16_t)h32[32 + l];  // only used when query_dual_ok() (n + 64 < 65536)
  }
  char* end = static_cast<char*>(ix->d_gids) + (size_t)ix->p.F * ix->gid_stride * ix->elem;
  NQ_CUDA(cudaMemcpyAsync(end, ix->elem == 2 ? (const void*)h16 : (const void*)h32, 64 * ix->elem, cudaMemcpyHostToDevice,
                          ix->ctx->stream));
  if (ix->d_gids16) {  // split16 copy: dead lanes count as "high", so their dummy id is stored minus 2^16
    for (uint32_t l = 0; l < 32; ++l) h16[l] = (uint16_t)(h32[l] - 65536u);
    NQ_CUDA(cudaMemcpyAsync(ix->d_gids16 + (size_t)ix->p.F * ix->gid_stride, h16, 32 * sizeof(uint16_t), cudaMemcpyHostToDevice,
                            ix->ctx->stream));
  }
  NQ_CUDA(cudaStreamSynchronize(ix->ctx->stream));  // the staging arrays live on this stack frame
  return NQ_OK;
}

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
  QueryArgs a{};
  a.qsk = d_sketches; a.dir = ix->d_row; a.gids = ix->d_gids;
  a.F = p.F; a.range = (uint32_t)p.range; a.n = ix->n; a.row_stride = ix->row_stride; a.gid_stride = ix->gid_stride;
  a.gid_base = ix->gid_base;
  a.dir3 = ix->d_dir3; a.gids16 = ix->d_gids16;
  a.nq_total = nq;
  a.min_score = min_score;
  a.wrap_mask = p.S <= 7 ? 0xFFu : p.S <= 15 ? 0xFFFFu : 0xFFFFFFFFu;  // counter widths of :635/:651/:667

  int mode;
  size_t smem;
  const bool slab = ix->slab_G && nq_slab_layout(ix, mode, smem);  // granule form of the postings (slab.cu)
  if (slab) {
    a.meta = ix->d_meta; a.slab = ix->d_slab; a.cell_gran = ix->d_cell_gran; a.mgroups = a.range / 32;
  } else {
    query_layout(ix, mode, smem);
  }
  // the prefetch window (kPfAhead + 1 chunks of kPfCells cells: directory rows + posting arrays) must sit in L2
  a.prefetch = query_prefetch_lead(ix, slab);
  if (const char* ex = nq_tuning_env("NQ_QUERY_EXP")) a.exp = (uint32_t)atoi(ex);

  const uint64_t q_per_launch = query_wave(ix, mode, smem, a, nq);
  if (set_query_parts(ix, mode, q_per_launch, a) != NQ_OK) { nq_dfree(ctx, a.slice_hits); delete hits; return NQ_ERR_CUDA; }
  if (slab) {  // descriptor scratch of one launch + the prefetch claims behind it
    if (nq_dmalloc(ctx, (void**)&a.desc, (q_per_launch * p.F + p.F / 256 + 16) * sizeof(uint32_t)) != NQ_OK) { delete hits; return NQ_ERR_CUDA; }
    a.pf_claim = a.desc + q_per_launch * p.F;
  }

  unsigned long long* d_cursor = nullptr;
  uint64_t* d_begin = nullptr;
  uint32_t* d_n = nullptr;
  uint32_t* d_gcounts = nullptr;
  int st = NQ_OK;
  auto cleanup = [&]() {
    nq_dfree(ctx, d_cursor); nq_dfree(ctx, d_begin); nq_dfree(ctx, d_n); nq_dfree(ctx, d_gcounts);
    nq_dfree(ctx, a.slice_hits); nq_dfree(ctx, a.slice_base); nq_dfree(ctx, a.desc);
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
      const cudaError_t e0 = launch_query(ix, mode, smem, nb, a, q0, ctx->stream);
      cudaError_t e = e0;
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

// Dense rows for --matrix: d_out[q][g] = (cells where query q and genome g carry the same valid
// fingerprint) & wrap_mask, for every genome of the shard.  Same kernels as nq_query_impl.
int nq_query_dense_impl(nq_index* ix, const int32_t* d_sketches, uint64_t nq, uint32_t wrap_mask, uint32_t* d_out) {
  if (!ix || !d_out) return nq_set_error(NQ_ERR_INVALID, "null argument");
  if (nq == 0) return NQ_OK;
  nq_ctx* ctx = ix->ctx;
  const nq_params& p = ix->p;
  QueryArgs a{};
  a.qsk = d_sketches; a.dir = ix->d_row; a.gids = ix->d_gids;
  a.F = p.F; a.range = (uint32_t)p.range; a.n = ix->n; a.row_stride = ix->row_stride; a.gid_stride = ix->gid_stride;
  a.gid_base = ix->gid_base;
  a.dir3 = ix->d_dir3; a.gids16 = ix->d_gids16;
  a.nq_total = nq;
  a.wrap_mask = wrap_mask;
  a.dense = d_out;
  int mode;
  size_t smem;
  const bool slab = ix->slab_G && nq_slab_layout(ix, mode, smem);
  if (slab) {
    a.meta = ix->d_meta; a.slab = ix->d_slab; a.cell_gran = ix->d_cell_gran; a.mgroups = a.range / 32;
  } else {
    query_layout(ix, mode, smem);
  }
  a.prefetch = query_prefetch_lead(ix, slab);
  if (const char* ex = nq_tuning_env("NQ_QUERY_EXP")) a.exp = (uint32_t)atoi(ex);
  const uint64_t q_per_launch = query_wave(ix, mode, smem, a, nq);
  if (set_query_parts(ix, mode, q_per_launch, a) != NQ_OK) { nq_dfree(ctx, a.slice_hits); return NQ_ERR_CUDA; }
  if (slab) {
    NQ_TRY(nq_dmalloc(ctx, (void**)&a.desc, (q_per_launch * p.F + p.F / 256 + 16) * sizeof(uint32_t)));
    a.pf_claim = a.desc + q_per_launch * p.F;
  }
  unsigned long long* d_cursor = nullptr;  // [1] = gather statistics
  NQ_TRY(nq_dmalloc(ctx, (void**)&d_cursor, 16));
  if (mode == kGlobal32) {
    const int st = nq_dmalloc(ctx, (void**)&a.gcounts, q_per_launch * ix->n * 4);
    if (st != NQ_OK) { nq_dfree(ctx, d_cursor); return st; }
  }
  a.cursor = d_cursor;
  cudaError_t e = cudaMemsetAsync(d_cursor, 0, 16, ctx->stream);
  {
    NqTimer timer(ctx, NQK_MATRIX);
    for (uint64_t q0 = 0; q0 < nq && e == cudaSuccess; q0 += q_per_launch) {
      const unsigned nb = (unsigned)std::min<uint64_t>(q_per_launch, nq - q0);
      e = launch_query(ix, mode, smem, nb, a, q0, ctx->stream);
      ctx->launches++;
      if (e == cudaSuccess) e = cudaPeekAtLastError();
    }
  }
  nq_dfree(ctx, d_cursor);
  nq_dfree(ctx, a.gcounts);
  nq_dfree(ctx, a.slice_hits);
  nq_dfree(ctx, a.slice_base);
  nq_dfree(ctx, a.desc);
  if (e != cudaSuccess) return nq_set_error(NQ_ERR_CUDA, "matrix row kernel launch failed: %s", cudaGetErrorString(e));
  return NQ_OK;
}
