// sketch.cu — K1+K2: rolling canonical k-mer hash + one-permutation bucket-min sketch, and K2b:
// densification.  Replaces Index::compute_sketch / sketch_densification and the helpers under
// them (/root/reference/src/niqki_index.cpp:114-123, 211-236, 240-273, 277-358).
//
// Scan kernel layout: one CTA per span (= a whole entry, or a slice of a long one).  The CTA keeps
// the 2^S-cell sketch in shared memory (u32, 0xFFFFFFFF = empty = int32 -1), every thread streams
// its own contiguous run of the span straight from HBM with 16-byte loads (the kernel is
// integer-ALU bound: ~1 byte of DRAM traffic per ~45 integer ops), and the per-CTA sketch is
// merged into the entry's sketch in HBM with atomicMin — bucket-min is associative and
// commutative, so slicing is exact (SURVEY.md App. A3 "Chunking").
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "device_common.cuh"
#include "internal.h"
#include "sketch_common.cuh"

namespace nq {

// char -> (forward code, complement code pre-shifted to its place in the reverse strand).
// nuc2int (:114-123): C,G,T -> 1,2,3, anything else 0.  nuc2intrc (:211-221): A,C,G -> 3,2,1,
// anything else 0.  Upper case only; every other byte is 0 on BOTH strands.
__device__ __forceinline__ uint32_t fw_code(uint32_t c) { return c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 0u; }
__device__ __forceinline__ uint32_t rv_code(uint32_t c) { return c == 'A' ? 3u : c == 'C' ? 2u : c == 'G' ? 1u : 0u; }

// str2numstrand (:255-273) accepts both cases; returns 4 for a foreign byte.
__device__ __forceinline__ uint32_t seed_code(uint32_t c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
  }
}

// DEF = the default parameter set K=31, W=12, H=4 (M=8, mask_M=255, maxrem=15): masks and shifts
// become immediates (kmask_lo = 2^32-1 needs no AND at all).  S stays a run-time value.
template <bool SMEM, int NT, bool RC_HI, bool SMALL_REM, bool DEF>
__global__ void __launch_bounds__(NT, 1) sketch_scan_kernel(const uint8_t* __restrict__ bases,
                                                         const uint64_t* __restrict__ offsets,
                                                         const Span* __restrict__ spans, uint32_t* gsk,
                                                         DevParams P) {
  __shared__ uint2 lut[256];                       // {fw, rv << (rc_shift or rc_shift-32)}; static => immediate address
  extern __shared__ __align__(16) uint32_t ssk[];  // [F] when SMEM

  uint64_t A, B, E0;
  uint32_t entry;
  if (spans) {
    const Span s = spans[blockIdx.x];
    A = s.kb; B = s.ke; E0 = s.e0; entry = s.entry;
  } else {
    entry = blockIdx.x;
    E0 = offsets[entry];
    const uint64_t E1 = offsets[entry + 1];
    A = E0;
    B = (E1 - E0 > P.K) ? E1 - P.K : E0;  // L-K k-mers: the one starting at L-K is skipped (:342, B2)
  }
  if (A >= B) return;

  constexpr bool rc_hi = RC_HI || DEF;  // 2K-2 >= 32: the complement code enters the high word
  const uint32_t rc_shift = DEF ? 60u : P.rc_shift;
  for (uint32_t c = threadIdx.x; c < 256; c += NT) {
    const uint32_t rv = rv_code(c);
    lut[c] = make_uint2(fw_code(c), rc_hi ? rv << (rc_shift - 32) : rv << rc_shift);
  }
  uint32_t* grow = gsk + (size_t)entry * P.F;
  if (SMEM)
    for (uint32_t i = threadIdx.x; i < P.F; i += NT) ssk[i] = kEmpty;
  else
    for (uint32_t i = threadIdx.x; i < P.F * (P.filter & 0xFFu) / 32; i += NT) ssk[i] = 0xFFFFFFFFu;  // filter: "anything may still win"
  __syncthreads();
  SketchSink<SMEM> sink{SMEM ? ssk : grow, reinterpret_cast<uint8_t*>(ssk), P.filter & 0xFFu, (P.filter & 0xFFu) ? P.W - (P.filter & 0xFFu) : 0u, P.filter >> 8};

  // this thread's run of k-mer starts [lo, hi): 16-byte aligned slices of the span
  const uint64_t base = A & ~15ull;
  uint64_t R = ((B - base + NT - 1) / NT + 15) & ~15ull;
  if (R < 64) R = 64;  // keeps every run but the first clear of the K-1 seed characters
  const uint64_t lo = max(A, base + (uint64_t)threadIdx.x * R);
  const uint64_t hi = min(B, base + (uint64_t)(threadIdx.x + 1) * R);

  if (lo < hi) {
    const uint32_t K = DEF ? 31u : P.K;
    uint32_t flo = 0, fhi = 0, rlo = 0, rhi = 0;
    const uint32_t kmask_lo = DEF ? 0xFFFFFFFFu : (uint32_t)P.kmask;
    const uint32_t kmask_hi = DEF ? 0x3FFFFFFFu : (uint32_t)(P.kmask >> 32);

    auto roll = [&](uint32_t fw, uint32_t rvlo, uint32_t rvhi) {
      fhi = __funnelshift_l(flo, fhi, 2) & kmask_hi;  // f = ((f<<2)+code) % 4^K      (:225-229)
      flo = DEF ? flo * 4u + fw : ((flo << 2) | fw) & kmask_lo;
      rlo = __funnelshift_r(rlo, rhi, 2) | rvlo;      // r = (r>>2) + (ccode<<(2K-2)) (:233-236)
      rhi = (rhi >> 2) | rvhi;
    };
    auto roll_char = [&](uint32_t c) {
      const uint2 e = lut[c];
      if (rc_hi) roll(e.x, 0u, e.y); else roll(e.x, e.y, 0u);
    };
    const uint32_t bshift = 32 - P.S;
    const uint32_t mask_M = DEF ? 255u : P.mask_M, maxrem = DEF ? 15u : P.maxrem, M = DEF ? 8u : P.M;
    auto emit = [&]() {
      constexpr uint32_t RCh = (uint32_t)(kRevC >> 32), RCl = (uint32_t)kRevC;
      constexpr uint32_t UCh = (uint32_t)(kUnrevC >> 32), UCl = (uint32_t)kUnrevC;
      const bool f_lt = (((uint64_t)fhi << 32) | flo) < (((uint64_t)rhi << 32) | rlo);  // canon = min(f, r) (:345)
      const uint32_t chi = f_lt ? fhi : rhi, clo = f_lt ? flo : rlo;
      const uint32_t t = clo ^ chi;                              // first fold, shared by both hashes
      // bucket = unrevhash64(canon) >> (64-S): only the high word of the second product (:347)
      const uint2 u1 = mul64c(chi, t, UCh, UCl);
      const uint32_t b = mul64c_hi_chain(u1.y, u1.x ^ u1.y, UCh, UCl) >> bshift;
      // fingerprint of revhash64(canon) (:346, :277-287): h = (hi2, lo2 ^ hi2)
      const uint2 r1 = mul64c(chi, t, RCh, RCl);
      const uint32_t t3 = r1.x ^ r1.y;
      const uint32_t hh = mul64c_hi_chain(r1.y, t3, RCh, RCl), hl = (t3 * RCl) ^ hh;
      if (DEF) {
        // rem = max(0, 15 - clz(hh)) = max(0, bfind(hh) - 16); bfind(0) = -1     (:280-286)
        // The shared sketch holds fp + 4096 (= max(bfind,16) << 8 instead of max(bfind-16,0) << 8:
        // a monotone offset, so the minimum is the same cell); the offset comes off at the flush.
        int msb;
        asm("bfind.u32 %0, %1;" : "=r"(msb) : "r"(hh));
        if (SMEM) sink.update(b, (hl & 255u) + ((uint32_t)max(msb, 16) << 8));
        else sink.update(b, (hl & 255u) + ((uint32_t)max(msb - 16, 0) << 8));
      } else {
        sink.update(b, fingerprint32<SMALL_REM>(hh, hl, mask_M, maxrem, M));  // :348
      }
    };

    // ---- warm-up over the K-1 characters before the first k-mer end
    if (lo == E0) {
      // record start: str2numstrand + rcb (:340-341).  Case-insensitive; one foreign byte zeroes
      // the whole seed (B4), which is the same as K-1 'A's on both strands.
      bool ok = true;
      for (uint32_t j = 0; j + 1 < K; ++j) ok = ok && (seed_code(bases[lo + j]) < 4);
      for (uint32_t j = 0; j + 1 < K; ++j) {
        const uint32_t code = ok ? seed_code(bases[lo + j]) : 0u;
        const uint64_t rv = (uint64_t)(3u - code) << rc_shift;
        roll(code, (uint32_t)rv, (uint32_t)(rv >> 32));
      }
    } else {
      for (uint32_t j = 0; j + 1 < K; ++j) roll_char(bases[lo + j]);
    }

    uint64_t pos = lo + K - 1;        // next character to consume; it ends the k-mer starting at pos-K+1
    const uint64_t end = hi + K - 1;  // one past the last character this thread consumes
    while (pos < end && (pos & 15)) {
      roll_char(bases[pos]);
      emit();
      ++pos;
    }
    if (pos + 16 <= end) {
      uint4 cur = __ldg(reinterpret_cast<const uint4*>(bases + pos));
      while (pos + 16 <= end) {
        const uint64_t nxt = pos + 16;
        uint4 pre = cur;
        if (nxt + 16 <= end) pre = __ldg(reinterpret_cast<const uint4*>(bases + nxt));  // prefetch
        const uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            roll_char((w[i] >> (8 * j)) & 0xFFu);
            emit();
          }
        }
        cur = pre;
        pos = nxt;
      }
    }
    while (pos < end) {
      roll_char(bases[pos]);
      emit();
      ++pos;
    }
  }

  if (SMEM) {
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < P.F; i += NT) {
      const uint32_t v = ssk[i];
      if (v != kEmpty) atomicMin(&grow[i], DEF ? v - 4096u : v);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Densification (:313-331).  The reference scans cells in index order and lets every non-empty
// cell i try to copy itself to t = hash_family(v_i, step) % F if that cell is empty.  Within a
// pass no new value appears, all holders of one value share one target, and the first holder in
// scan order decides, so a pass is equivalent to: every empty target takes the value of the
// LOWEST-INDEX cell aiming at it (verified against the reference).  That is an
// atomicMin of the source index per target: phase A posts 0x80000000|index into empty targets,
// phase B replaces the winner's index by its value.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kTent = 0x80000000u;

template <bool SMEM, int NT>
__global__ void __launch_bounds__(NT) densify_kernel(uint32_t* gsk, DevParams P, uint32_t* flags) {
  extern __shared__ __align__(16) uint32_t smem[];
  __shared__ uint32_t s_count;
  const uint32_t F = P.F, Fmask = F - 1;
  uint32_t* grow = gsk + (size_t)blockIdx.x * F;
  uint32_t* sk = SMEM ? smem : grow;

  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  uint32_t my_empty = 0;
  for (uint32_t i = threadIdx.x; i < F; i += NT) {
    const uint32_t v = grow[i];
    if (SMEM) sk[i] = v;
    my_empty += (v == kEmpty);
  }
  if (my_empty) atomicAdd(&s_count, my_empty);
  __syncthreads();
  uint32_t empty = s_count;
  __syncthreads();
  if (empty == 0) return;
  if (empty == F) {  // nothing was sketched (len <= K): callers of the reference drop the entry
    if (flags && threadIdx.x == 0) flags[blockIdx.x] |= NQ_ENTRY_SKIPPED;
    return;
  }

  uint32_t step = 0, idle = 0;
  // targets repeat with period F in `step`, so F passes without a fill means the reference's
  // loop would spin forever (seen on entries with 1-2 k-mers); stop and flag instead.
  while (empty != 0 && idle < F) {
    if (threadIdx.x == 0) s_count = 0;
    for (uint32_t i = threadIdx.x; i < F; i += NT) {
      const uint32_t v = sk[i];
      if (v < kTent) {
        const uint32_t t = (family_a(v, Fmask) + step * family_b(v, Fmask)) & Fmask;
        if (sk[t] >= kTent) atomicMin(&sk[t], kTent | i);
      }
    }
    __syncthreads();
    uint32_t filled = 0;
    for (uint32_t i = threadIdx.x; i < F; i += NT) {
      const uint32_t x = sk[i];
      if (x >= kTent && x != kEmpty) {
        sk[i] = sk[x & ~kTent];
        ++filled;
      }
    }
    if (filled) atomicAdd(&s_count, filled);
    __syncthreads();
    const uint32_t got = s_count;
    __syncthreads();
    empty -= got;
    idle = got ? 0 : idle + 1;
    ++step;
  }
  if (SMEM)
    for (uint32_t i = threadIdx.x; i < F; i += NT) grow[i] = sk[i];
  if (empty != 0 && flags && threadIdx.x == 0) flags[blockIdx.x] |= NQ_ENTRY_DENSIFY_STALLED;
}

// ------------------------------------------------------------------------------------------
// Short entries (--indexlines / --querylines: one read per entry, small S): one WARP per entry,
// sketch + densification fused, everything in the warp's slice of shared memory.  The CTA-per-entry
// kernels above leave 126 of 128 threads idle on a 150 bp read and pay two launches' worth of
// global sketch traffic; here lane l rolls its own run of ~nk/32 k-mers from scratch — the K-1
// characters before a run are re-read by the lane, with the record-seed rule (str2numstrand,
// :255-273: both cases accepted, one foreign byte zeroes the whole seed, B4) applied by POSITION
// (characters 0..K-2 of the entry) so that a run may start anywhere — and densification (:313-331,
// same lowest-index-source rule as densify_kernel) runs on the warp with the two hash_family
// terms of every cell cached beside it (they are functions of the cell's value, copied along).
// ------------------------------------------------------------------------------------------
// LISTED: densification walks a dense list of the initially non-empty cells {value, cached hash terms, lowest
// index holding the value so far} instead of all F cells: every holder of a value aims at the same target, so
// only the lowest-index holder matters (it is what the reference's in-order scan lets win), and a value that
// fills a cell just lowers its entry's index if the new cell lies below it.  ~95 entries instead of 256 cells
// per pass for a 150 bp read at S=8, and no resolve sweep: the winner of a target recognises its own mark.
template <int NW, bool LISTED>
__global__ void __launch_bounds__(NW * 32) sketch_reads_kernel(const uint8_t* __restrict__ bases,
                                                               const uint64_t* __restrict__ offsets, uint64_t n,
                                                               uint32_t* __restrict__ gsk, DevParams P,
                                                               uint32_t* __restrict__ flags) {
  extern __shared__ __align__(16) uint32_t smem[];  // per warp: sk[F] | ab[F] ({family_a, family_b} as two u16) [| lv[F] | lidx[F]]
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t F = P.F, Fmask = F - 1, K = P.K;
  uint32_t* sk = smem + (size_t)warp * (LISTED ? 4 : 2) * F;
  uint32_t* ab = sk + F;      // cell-indexed (cell form) or entry-indexed (listed form)
  uint32_t* lv = ab + F;      // listed form: value of entry e
  uint32_t* lidx = lv + F;    // listed form: lowest index of a cell holding that value
  constexpr unsigned kFull = 0xFFFFFFFFu;
  for (uint64_t entry = (uint64_t)blockIdx.x * NW + warp; entry < n; entry += (uint64_t)gridDim.x * NW) {
    const uint64_t E0 = offsets[entry], L = offsets[entry + 1] - E0;
    uint32_t* grow = gsk + entry * F;
    for (uint32_t i = lane; i < F; i += 32) sk[i] = kEmpty;
    __syncwarp();
    if (L > K) {
      const uint32_t nk = (uint32_t)(L - K);  // k-mers 0..L-K-1: the last one is skipped (:342, B2)
      const uint8_t* e = bases + E0;
      // seed rule: any foreign byte among the first K-1 characters zeroes all of them
      const bool seed_ok = __all_sync(kFull, lane + 1 >= K || seed_code(e[lane]) < 4);
      const uint32_t run = (nk + 31) / 32, lo = lane * run, hi = min(nk, lo + run);
      if (lo < hi) {
        uint64_t f = 0, r = 0;
        auto roll_at = [&](uint32_t x) {
          const uint32_t c = e[x];
          uint32_t fw, rv;
          if (x + 1 < K) { fw = seed_ok ? seed_code(c) : 0u; rv = 3u - fw; }  // seed region (:255-273, :240-250)
          else { fw = fw_code(c); rv = rv_code(c); }                          // rolled characters (:114-123, :211-221)
          f = ((f << 2) + fw) & P.kmask;                 // :225-229
          r = (r >> 2) + ((uint64_t)rv << P.rc_shift);   // :233-236
        };
        for (uint32_t j = 0; j + 1 < K; ++j) roll_at(lo + j);
        for (uint32_t p = lo; p < hi; ++p) {
          roll_at(p + K - 1);
          const uint64_t canon = f < r ? f : r;                                            // :345
          const uint32_t b = (uint32_t)(unrevhash64(canon) >> (64 - P.S));                 // :347
          atomicMin(&sk[b], fingerprint(revhash64(canon), P.mask_M, P.maxrem, P.M));       // :346, :348-355
        }
      }
    }
    __syncwarp();
    // ---- densification
    uint32_t my_empty = 0, n_ent = 0;
    for (uint32_t i0 = 0; i0 < F; i0 += 32) {
      const uint32_t i = i0 + lane;
      const bool in = i < F;  // F < 32 when S < 5
      const uint32_t v = in ? sk[i] : kEmpty;
      const bool full = in && v != kEmpty;
      if (in && !full) ++my_empty;
      if (LISTED) {
        const unsigned bal = __ballot_sync(kFull, full);
        if (full) {
          const uint32_t e = n_ent + __popc(bal & ((1u << lane) - 1));
          lv[e] = v;
          lidx[e] = i;
          ab[e] = family_a(v, Fmask) | (family_b(v, Fmask) << 16);
        }
        n_ent += __popc(bal);
      } else if (full) {
        ab[i] = family_a(v, Fmask) | (family_b(v, Fmask) << 16);
      }
    }
    uint32_t empty = __reduce_add_sync(kFull, my_empty);
    uint32_t fl = 0;
    if (empty == F) {
      fl = NQ_ENTRY_SKIPPED;  // nothing was sketched (len <= K)
    } else if (empty && LISTED) {
      __syncwarp();
      uint32_t step = 0, idle = 0;
      while (empty != 0 && idle < F) {
        for (uint32_t e = lane; e < n_ent; e += 32) {  // every entry marks its target with its lowest holder's index
          const uint32_t h = ab[e];
          const uint32_t t = ((h & 0xFFFFu) + step * (h >> 16)) & Fmask;
          if (sk[t] >= kTent) atomicMin(&sk[t], kTent | lidx[e]);
        }
        __syncwarp();
        uint32_t filled = 0;
        for (uint32_t e = lane; e < n_ent; e += 32) {  // the entry whose mark survived fills the cell
          const uint32_t h = ab[e];
          const uint32_t t = ((h & 0xFFFFu) + step * (h >> 16)) & Fmask;
          const uint32_t mine = lidx[e];
          if (sk[t] == (kTent | mine)) {
            sk[t] = lv[e];
            if (t < mine) lidx[e] = t;
            ++filled;
          }
        }
        const uint32_t got = __reduce_add_sync(kFull, filled);
        __syncwarp();
        empty -= got;
        idle = got ? 0 : idle + 1;
        ++step;
      }
      if (empty) fl = NQ_ENTRY_DENSIFY_STALLED;
    } else if (empty) {
      __syncwarp();
      uint32_t step = 0, idle = 0;
      while (empty != 0 && idle < F) {
        for (uint32_t i = lane; i < F; i += 32) {
          const uint32_t v = sk[i];
          if (v < kTent) {
            const uint32_t h = ab[i];
            const uint32_t t = ((h & 0xFFFFu) + step * (h >> 16)) & Fmask;
            if (sk[t] >= kTent) atomicMin(&sk[t], kTent | i);
          }
        }
        __syncwarp();
        uint32_t filled = 0;
        for (uint32_t i = lane; i < F; i += 32) {
          const uint32_t x = sk[i];
          if (x >= kTent && x != kEmpty) {
            const uint32_t src = x & ~kTent;
            sk[i] = sk[src];
            ab[i] = ab[src];
            ++filled;
          }
        }
        const uint32_t got = __reduce_add_sync(kFull, filled);
        __syncwarp();
        empty -= got;
        idle = got ? 0 : idle + 1;
        ++step;
      }
      if (empty) fl = NQ_ENTRY_DENSIFY_STALLED;
    }
    __syncwarp();
    for (uint32_t i = lane; i < F; i += 32) grow[i] = sk[i];
    if (flags && lane == 0) flags[entry] = fl;
    __syncwarp();
  }
}

static DevParams make_dev_params(const nq_params* p) {
  DevParams d;
  d.K = p->K; d.S = p->S; d.W = p->W; d.M = p->M; d.F = p->F;
  d.mask_M = p->mask_M; d.maxrem = p->maxrem; d.range = (uint32_t)p->range;
  d.kmask = (1ull << (2 * p->K)) - 1;
  d.rc_shift = 2 * p->K - 2;
  d.filter = 0;
  return d;
}

}  // namespace nq

using namespace nq;

DevParams nq_make_dev_params(const nq_params* p) { return make_dev_params(p); }

template <bool SMEM, int NT, bool RC_HI, bool SMALL_REM, bool DEF>
static int launch_scan_t(nq_ctx* ctx, const DevParams& P, const uint8_t* d_bases, const uint64_t* d_offsets,
                         const Span* d_spans, uint64_t nblocks, uint32_t* d_sk) {
  const size_t smem = SMEM ? (size_t)P.F * 4 : (size_t)P.F * (P.filter & 0xFFu) / 8;  // sketch, or the coarse filter of the global form
  auto kern = sketch_scan_kernel<SMEM, NT, RC_HI, SMALL_REM, DEF>;
  NQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NqTimer timer(ctx, NQK_SCAN);
  kern<<<(unsigned)nblocks, NT, smem, ctx->stream>>>(d_bases, d_offsets, d_spans, d_sk, P);
  NQ_CHECK_LAUNCH(ctx);
  return NQ_OK;
}

template <bool SMEM, int NT>
static int launch_scan(nq_ctx* ctx, const DevParams& P, const uint8_t* d_bases, const uint64_t* d_offsets,
                       const Span* d_spans, uint64_t nblocks, uint32_t* d_sk) {
  const bool rc_hi = P.rc_shift >= 32, small_rem = P.maxrem <= 32;
  if (P.K == 31 && P.M == 8 && P.mask_M == 255 && P.maxrem == 15)  // the default K/W/H (any S)
    return launch_scan_t<SMEM, NT, true, true, true>(ctx, P, d_bases, d_offsets, d_spans, nblocks, d_sk);
  if (rc_hi) return small_rem ? launch_scan_t<SMEM, NT, true, true, false>(ctx, P, d_bases, d_offsets, d_spans, nblocks, d_sk)
                              : launch_scan_t<SMEM, NT, true, false, false>(ctx, P, d_bases, d_offsets, d_spans, nblocks, d_sk);
  return small_rem ? launch_scan_t<SMEM, NT, false, true, false>(ctx, P, d_bases, d_offsets, d_spans, nblocks, d_sk)
                   : launch_scan_t<SMEM, NT, false, false, false>(ctx, P, d_bases, d_offsets, d_spans, nblocks, d_sk);
}

int nq_launch_sketch(nq_ctx* ctx, const nq_params* p, const char* d_bases, uint64_t bases_capacity,
                     const uint64_t* h_offsets, uint64_t n, const uint32_t* h_rec_entry, uint64_t n_entries,
                     int32_t* d_sketches, uint32_t* d_flags) {
  NQ_TRY(nq_params_check(p));
  if (!h_rec_entry) n_entries = n;
  if (n_entries == 0) return NQ_OK;
  if (n >= (1ull << 31) || n_entries >= (1ull << 31))
    return nq_set_error(NQ_ERR_INVALID, "too many entries in one batch: %llu", (unsigned long long)n);
  if (h_rec_entry) {
    bool identity = n == n_entries;
    for (uint64_t e = 0; e < n; ++e) {
      if (h_rec_entry[e] >= n_entries) return nq_set_error(NQ_ERR_INVALID, "record %llu maps to entry %u >= %llu",
                                                           (unsigned long long)e, h_rec_entry[e], (unsigned long long)n_entries);
      identity = identity && h_rec_entry[e] == e;
    }
    if (identity) h_rec_entry = nullptr;  // one record per entry (lines mode): the fused short-entry kernel applies
  }
  if ((reinterpret_cast<uintptr_t>(d_bases) & 15) != 0)
    return nq_set_error(NQ_ERR_INVALID, "d_bases must be 16-byte aligned");
  if (bases_capacity < h_offsets[n])
    return nq_set_error(NQ_ERR_INVALID, "bases_capacity %llu < offsets[n] %llu", (unsigned long long)bases_capacity,
                        (unsigned long long)h_offsets[n]);
  const DevParams P = make_dev_params(p);
  const size_t cells = (size_t)n_entries * P.F;
  NQ_CUDA(cudaMemsetAsync(d_sketches, 0xFF, cells * sizeof(int32_t), ctx->stream));
  if (d_flags) NQ_CUDA(cudaMemsetAsync(d_flags, 0, n_entries * sizeof(uint32_t), ctx->stream));

  // spans: slice entries only when there are too few CTAs to fill the machine
  uint64_t total_k = 0, longest = 0;
  for (uint64_t e = 0; e < n; ++e) {
    const uint64_t len = h_offsets[e + 1] - h_offsets[e];
    if (len > p->K) {
      total_k += len - p->K;
      longest = std::max(longest, len - p->K);
    }
  }
  // short entries, small sketches (lines mode): warp-per-entry fused kernel, no global sketch traffic
  if (!h_rec_entry && longest + p->K <= 4096 && P.S <= 11 && P.K <= 32) {
    static const char* env = nq_tuning_env("NQ_READS_KERNEL");  // "0": the CTA-per-entry kernels (measurement only)
    if (!(env && env[0] == '0')) {
      static const char* dens_env = nq_tuning_env("NQ_READS_DENSIFY");  // "cell": the all-cells densification (measurement only)
      const bool listed = !(dens_env && dens_env[0] == 'c');
      NqScratch s_off(ctx);
      NQ_TRY(nq_dmalloc(ctx, &s_off.p, (n + 1) * sizeof(uint64_t)));
      uint64_t* d_offsets = s_off.as<uint64_t>();
      NQ_CUDA(cudaMemcpyAsync(d_offsets, h_offsets, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
      auto launch = [&](auto kern, int nw, size_t smem) -> int {
        NQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 1;
        NQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nw * 32, smem));
        const unsigned grid = (unsigned)std::min<uint64_t>((n + nw - 1) / nw, (uint64_t)ctx->sm_count * std::max(per_sm, 1));
        NqTimer timer(ctx, NQK_SCAN);
        kern<<<grid, nw * 32, smem, ctx->stream>>>(reinterpret_cast<const uint8_t*>(d_bases), d_offsets, n,
                                                   reinterpret_cast<uint32_t*>(d_sketches), P, d_flags);
        return NQ_OK;
      };
      int lst;
      if (!listed) lst = launch(sketch_reads_kernel<8, false>, 8, (size_t)8 * 2 * P.F * 4);
      else if (P.F <= 1024) lst = launch(sketch_reads_kernel<8, true>, 8, (size_t)8 * 4 * P.F * 4);
      else lst = launch(sketch_reads_kernel<4, true>, 4, (size_t)4 * 4 * P.F * 4);
      if (lst != NQ_OK) return lst;
      ctx->launches++;
      const cudaError_t le = cudaPeekAtLastError();
      if (le != cudaSuccess) return nq_set_error(NQ_ERR_CUDA, "sketch_reads_kernel launch failed: %s", cudaGetErrorString(le));
      return NQ_OK;
    }
  }
  if (total_k == 0) return nq_launch_densify(ctx, p, d_sketches, n_entries, d_flags);
  uint64_t span_len = (total_k / ((uint64_t)ctx->sm_count * 8) + 1023) & ~1023ull;
  span_len = std::max<uint64_t>(span_len, 65536);
  const bool sliced = longest > span_len || h_rec_entry != nullptr;  // a record->row map needs spans

  NqScratch s_offsets(ctx), s_spans(ctx);  // returned on every path out
  uint64_t nblocks = n;
  NQ_TRY(nq_dmalloc(ctx, &s_offsets.p, (n + 1) * sizeof(uint64_t)));
  uint64_t* d_offsets = s_offsets.as<uint64_t>();
  NQ_TRY(nq_upload_small(ctx, d_offsets, h_offsets, (n + 1) * sizeof(uint64_t)));
  std::vector<Span> spans;
  if (sliced) {
    for (uint64_t e = 0; e < n; ++e) {
      const uint64_t e0 = h_offsets[e], len = h_offsets[e + 1] - e0;
      if (len <= p->K) continue;
      const uint64_t nk = len - p->K;
      const uint64_t parts = (nk + span_len - 1) / span_len;
      const uint64_t each = ((nk + parts - 1) / parts + 1023) & ~1023ull;
      for (uint64_t a = 0; a < nk; a += each)
        spans.push_back(Span{e0 + a, e0 + std::min(nk, a + each), e0, h_rec_entry ? h_rec_entry[e] : (uint32_t)e, 0});
    }
    nblocks = spans.size();
    NQ_TRY(nq_dmalloc(ctx, &s_spans.p, spans.size() * sizeof(Span)));
    NQ_TRY(nq_upload_small(ctx, s_spans.p, spans.data(), spans.size() * sizeof(Span)));  // through the pinned ring: no host stall
  }
  const Span* d_spans = s_spans.as<Span>();

  const uint8_t* b = reinterpret_cast<const uint8_t*>(d_bases);
  uint32_t* sk = reinterpret_cast<uint32_t*>(d_sketches);
  const bool fits = 2048 + 1024 + (size_t)P.F * 4 <= ctx->smem_optin;  // + static LUT + driver reserve
  const bool small = total_k / std::max<uint64_t>(nblocks, 1) < 16384;  // short entries: small CTAs
  int st;
  if (fits) {
    st = small ? launch_scan<true, 128>(ctx, P, b, d_offsets, d_spans, nblocks, sk)
               : launch_scan<true, 1024>(ctx, P, b, d_offsets, d_spans, nblocks, sk);
  } else {
    // the sketch stays in HBM/L2; a coarse filter (8 or 4 bits per cell) takes the shared memory instead
    DevParams PG = P;
    static const char* env = nq_tuning_env("NQ_SCAN_FILTER");  // "0": no filter (measurement only)
    const size_t room = ctx->smem_optin - 2048 - 1024;
    if (!small && !(env && env[0] == '0')) {
      if (P.W >= 8 && (size_t)P.F <= room) PG.filter = 8;
      else if (P.W >= 4 && (size_t)P.F / 2 <= room) PG.filter = 4;
    }
    // bit 8: read the cell before the atomic — only when no coarse filter stands in front of it
    static const char* gr = nq_tuning_env("NQ_SCAN_GREAD");  // "1" / "0": force (measurement only)
    if (gr ? gr[0] == '1' : PG.filter == 0) PG.filter |= 0x100u;
    st = small ? launch_scan<false, 128>(ctx, PG, b, d_offsets, d_spans, nblocks, sk)
               : launch_scan<false, 1024>(ctx, PG, b, d_offsets, d_spans, nblocks, sk);
  }
  NQ_TRY(st);
  return nq_launch_densify(ctx, p, d_sketches, n_entries, d_flags);
}

template <bool SMEM, int NT>
static int launch_densify(nq_ctx* ctx, const DevParams& P, uint32_t* d_sk, uint64_t n, uint32_t* d_flags) {
  const size_t smem = SMEM ? (size_t)P.F * 4 : 0;
  auto kern = densify_kernel<SMEM, NT>;
  NQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NqTimer timer(ctx, NQK_DENSIFY);
  kern<<<(unsigned)n, NT, smem, ctx->stream>>>(d_sk, P, d_flags);
  NQ_CHECK_LAUNCH(ctx);
  return NQ_OK;
}

int nq_launch_densify(nq_ctx* ctx, const nq_params* p, int32_t* d_sketches, uint64_t n, uint32_t* d_flags) {
  NQ_TRY(nq_params_check(p));
  if (n == 0) return NQ_OK;
  const DevParams P = make_dev_params(p);
  uint32_t* sk = reinterpret_cast<uint32_t*>(d_sketches);
  const bool fits = (size_t)P.F * 4 <= ctx->smem_optin;
  if (P.F <= 1024) return fits ? launch_densify<true, 64>(ctx, P, sk, n, d_flags) : launch_densify<false, 64>(ctx, P, sk, n, d_flags);
  if (P.F <= 8192) return fits ? launch_densify<true, 256>(ctx, P, sk, n, d_flags) : launch_densify<false, 256>(ctx, P, sk, n, d_flags);
  return fits ? launch_densify<true, 1024>(ctx, P, sk, n, d_flags) : launch_densify<false, 1024>(ctx, P, sk, n, d_flags);
}
