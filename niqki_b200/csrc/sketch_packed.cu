// sketch_packed.cu — K1+K2 on 2-bit packed sequence (the wire format of pack.cpp): the same
// rolling canonical k-mer hash + bucket-min sketch as sketch.cu
// (/root/reference/src/niqki_index.cpp:114-123, 211-236, 335-358), but the forward and reverse
// k-mers are not rolled base by base — every k-mer is a WINDOW of the packed stream:
//   f = sum c[p+j] * 4^(K-1-j)        = the 2K bits from base p on of the big-endian stream;
//   r = sum comp(c[p+j]) * 4^j        = the 2K bits from base p on of the little-endian stream of
//                                       complement codes (3 - code, 0 where the byte was not ACGT).
// A thread keeps three 32-bit words of either stream in registers and cuts 16 consecutive k-mers
// out of them with funnel shifts whose amounts are compile-time constants: two shifts and a mask
// per strand and base, no character look-up table, no byte extraction, no shared-memory loads.
// Per 16 bases one code word (+ its `other` mask where the 512-base block has one) is turned into
// the two stream words (~10 ALU ops).  Seeds, lower case and foreign bytes were resolved by the
// packer, so the kernel has no per-record special case.  Bucket-min, fingerprint and the merge of
// CTA sketches are those of sketch.cu.
#include <algorithm>
#include <vector>

#include "device_common.cuh"
#include "internal.h"
#include "pack.h"
#include "sketch_common.cuh"

namespace nq {

// 16 x 2-bit codes, base i at bits 2i  ->  base i at bits 30-2i (big-endian order of the forward strand)
__device__ __forceinline__ uint32_t pairs_reversed(uint32_t w) {
  const uint32_t t = __brev(w);  // pairs reversed, bits inside each pair swapped
  return ((t >> 1) & 0x55555555u) | ((t & 0x55555555u) << 1);
}
// 16 mask bits -> 16 bit pairs
__device__ __forceinline__ uint32_t pairs_of_bits(uint32_t m) {
  uint32_t x = m;
  x = (x | (x << 8)) & 0x00FF00FFu;
  x = (x | (x << 4)) & 0x0F0F0F0Fu;
  x = (x | (x << 2)) & 0x33333333u;
  x = (x | (x << 1)) & 0x55555555u;
  return x | (x << 1);
}

struct PackedIn {
  const uint32_t* codes;
  const uint32_t* blk;
  const uint16_t* pool;
  // word m of the two streams: big-endian forward codes, little-endian complement codes
  __device__ __forceinline__ void load(uint64_t m, uint32_t& be, uint32_t& comp) const {
    const uint32_t w = __ldg(codes + m);
    const uint32_t slot = __ldg(blk + (m >> 5));
    uint32_t c = ~w;
    if (slot != 0xFFFFFFFFu) c &= ~pairs_of_bits(__ldg(pool + (size_t)slot * 32 + (m & 31)));
    be = pairs_reversed(w);
    comp = c;
  }
};

template <bool SMEM, bool DEF, bool SMALL_REM>
struct KmerEmit {
  SketchSink<SMEM> sink;
  uint32_t bshift, mask_M, maxrem, M;
  __device__ __forceinline__ void operator()(uint32_t fhi, uint32_t flo, uint32_t rhi, uint32_t rlo) const {
    constexpr uint32_t RCh = (uint32_t)(kRevC >> 32), RCl = (uint32_t)kRevC;
    constexpr uint32_t UCh = (uint32_t)(kUnrevC >> 32), UCl = (uint32_t)kUnrevC;
    const bool f_lt = (((uint64_t)fhi << 32) | flo) < (((uint64_t)rhi << 32) | rlo);  // canon = min(f, r) (:345)
    const uint32_t chi = f_lt ? fhi : rhi, clo = f_lt ? flo : rlo;
    const uint32_t t = clo ^ chi;
    const uint2 u1 = mul64c(chi, t, UCh, UCl);
    const uint32_t b = mul64c_hi_chain(u1.y, u1.x ^ u1.y, UCh, UCl) >> bshift;  // bucket (:347)
    const uint2 r1 = mul64c(chi, t, RCh, RCl);
    const uint32_t t3 = r1.x ^ r1.y;
    const uint32_t hh = mul64c_hi_chain(r1.y, t3, RCh, RCl), hl = (t3 * RCl) ^ hh;  // revhash64 (:346)
    if (DEF) {
      int msb;  // same +4096 offset form as sketch.cu (taken off at the flush)
      asm("bfind.u32 %0, %1;" : "=r"(msb) : "r"(hh));
      if (SMEM) sink.update(b, (hl & 255u) + ((uint32_t)max(msb, 16) << 8));
      else sink.update(b, (hl & 255u) + ((uint32_t)max(msb - 16, 0) << 8));
    } else {
      sink.update(b, fingerprint32<SMALL_REM>(hh, hl, mask_M, maxrem, M));  // :348
    }
  }
};

template <bool SMEM, int NT, bool DEF, bool SMALL_REM>
__global__ void __launch_bounds__(NT, 1) sketch_scan_packed_kernel(PackedIn in, const uint64_t* __restrict__ offsets,
                                                                const Span* __restrict__ spans, uint32_t* gsk, DevParams P) {
  extern __shared__ __align__(16) uint32_t ssk[];  // [F] when SMEM, else the coarse filter

  uint64_t A, B;
  uint32_t entry;
  if (spans) {
    const Span s = spans[blockIdx.x];
    A = s.kb; B = s.ke; entry = s.entry;
  } else {
    entry = blockIdx.x;
    const uint64_t E0 = offsets[entry], E1 = offsets[entry + 1];
    A = E0;
    B = (E1 - E0 > P.K) ? E1 - P.K : E0;  // L-K k-mers: the one starting at L-K is skipped (:342, B2)
  }
  if (A >= B) return;
  A += kPackLead; B += kPackLead;  // positions in the packed stream

  uint32_t* grow = gsk + (size_t)entry * P.F;
  if (SMEM)
    for (uint32_t i = threadIdx.x; i < P.F; i += NT) ssk[i] = kEmpty;
  else
    for (uint32_t i = threadIdx.x; i < P.F * (P.filter & 0xFFu) / 32; i += NT) ssk[i] = 0xFFFFFFFFu;
  __syncthreads();
  KmerEmit<SMEM, DEF, SMALL_REM> emit{{SMEM ? ssk : grow, reinterpret_cast<uint8_t*>(ssk), P.filter & 0xFFu, (P.filter & 0xFFu) ? P.W - (P.filter & 0xFFu) : 0u, P.filter >> 8},
                                       32 - P.S, DEF ? 255u : P.mask_M, DEF ? 15u : P.maxrem, DEF ? 8u : P.M};

  // this thread's run of k-mer starts [lo, hi): whole 16-base words except at the ends of the span
  const uint64_t base = A & ~15ull;
  uint64_t R = ((B - base + NT - 1) / NT + 15) & ~15ull;
  if (R < 16) R = 16;
  const uint64_t lo = max(A, base + (uint64_t)threadIdx.x * R);
  const uint64_t hi = min(B, base + (uint64_t)(threadIdx.x + 1) * R);

  if (lo < hi) {
    const uint32_t K = DEF ? 31u : P.K;
    const uint32_t fsh = 64 - 2 * K;  // generic K: f = window >> fsh
    const uint32_t kmask_lo = (uint32_t)P.kmask, kmask_hi = (uint32_t)(P.kmask >> 32);
    // any k-mer start p, shift amounts at run time (ends of a span only)
    auto slow_kmer = [&](uint64_t p) {
      const uint64_t m = p >> 4;
      const uint32_t s = 2 * (uint32_t)(p & 15);
      uint32_t b0, b1, b2, c0, c1, c2;
      in.load(m, b0, c0); in.load(m + 1, b1, c1); in.load(m + 2, b2, c2);
      const uint32_t xhi = __funnelshift_l(b1, b0, s), xlo = __funnelshift_l(b2, b1, s);
      const uint32_t ylo = __funnelshift_r(c0, c1, s), yhi = __funnelshift_r(c1, c2, s);
      uint32_t fhi, flo;
      if (fsh < 32) { fhi = xhi >> fsh; flo = __funnelshift_r(xlo, xhi, fsh); }
      else { fhi = 0; flo = xhi >> ((fsh - 32) & 31); }
      emit(fhi, flo, yhi & kmask_hi, ylo & kmask_lo);
    };
    const uint64_t mf = (lo + 15) >> 4, ml = hi >> 4;  // full words [mf, ml)
    if (mf >= ml) {
      for (uint64_t p = lo; p < hi; ++p) slow_kmer(p);
    } else {
      for (uint64_t p = lo; p < (mf << 4); ++p) slow_kmer(p);
      uint32_t bm1, b0, b1, b2, c0, c1, c2, cx;
      in.load(mf - 1, bm1, cx);
      in.load(mf, b0, c0);
      in.load(mf + 1, b1, c1);
      in.load(mf + 2, b2, c2);
      for (uint64_t m = mf; m < ml; ++m) {
        uint32_t b3, c3;
        in.load(m + 3, b3, c3);  // next iteration's third word (the stream has zero words behind its end)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          uint32_t fhi, flo, rhi, rlo;
          if (DEF) {
            // K = 31: the 64-bit big-endian window from base p-1 on, its top base masked off
            if (j == 0) { fhi = __funnelshift_l(b0, bm1, 30); flo = __funnelshift_l(b1, b0, 30); }
            else if (j == 1) { fhi = b0; flo = b1; }
            else { fhi = __funnelshift_l(b1, b0, 2 * (j - 1)); flo = __funnelshift_l(b2, b1, 2 * (j - 1)); }
            fhi &= 0x3FFFFFFFu;
            if (j == 0) { rlo = c0; rhi = c1; }
            else { rlo = __funnelshift_r(c0, c1, 2 * j); rhi = __funnelshift_r(c1, c2, 2 * j); }
            rhi &= 0x3FFFFFFFu;
          } else {
            const uint32_t xhi = j ? __funnelshift_l(b1, b0, 2 * j) : b0, xlo = j ? __funnelshift_l(b2, b1, 2 * j) : b1;
            if (fsh < 32) { fhi = xhi >> fsh; flo = __funnelshift_r(xlo, xhi, fsh); }
            else { fhi = 0; flo = xhi >> ((fsh - 32) & 31); }
            rlo = (j ? __funnelshift_r(c0, c1, 2 * j) : c0) & kmask_lo;
            rhi = (j ? __funnelshift_r(c1, c2, 2 * j) : c1) & kmask_hi;
          }
          emit(fhi, flo, rhi, rlo);
        }
        bm1 = b0; b0 = b1; b1 = b2; b2 = b3;
        c0 = c1; c1 = c2; c2 = c3;
      }
      for (uint64_t p = max(lo, ml << 4); p < hi; ++p) slow_kmer(p);
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

}  // namespace nq

using namespace nq;

template <bool SMEM, int NT>
static int launch_packed(nq_ctx* ctx, const DevParams& P, const PackedIn& in, const uint64_t* d_offsets, const Span* d_spans,
                         uint64_t nblocks, uint32_t* d_sk) {
  const size_t smem = SMEM ? (size_t)P.F * 4 : (size_t)P.F * (P.filter & 0xFFu) / 8;
  const bool def = P.K == 31 && P.M == 8 && P.mask_M == 255 && P.maxrem == 15, small_rem = P.maxrem <= 32;
  NqTimer timer(ctx, NQK_SCAN);
#define NQ_LAUNCH(DEFV, SR)                                                                              \
  do {                                                                                                   \
    auto kern = sketch_scan_packed_kernel<SMEM, NT, DEFV, SR>;                                            \
    NQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
    kern<<<(unsigned)nblocks, NT, smem, ctx->stream>>>(in, d_offsets, d_spans, d_sk, P);                 \
  } while (0)
  if (def) NQ_LAUNCH(true, true);
  else if (small_rem) NQ_LAUNCH(false, true);
  else NQ_LAUNCH(false, false);
#undef NQ_LAUNCH
  NQ_CHECK_LAUNCH(ctx);
  return NQ_OK;
}

DevParams nq_make_dev_params(const nq_params* p);

// Packed counterpart of nq_launch_sketch (sketch.cu): same spans, same densification afterwards.
// `h_offsets` are the record boundaries in the ORIGINAL character stream (offsets[0] = 0).
int nq_launch_sketch_packed(nq_ctx* ctx, const nq_params* p, const uint32_t* d_codes, const uint32_t* d_blk, const uint16_t* d_pool,
                            const uint64_t* h_offsets, uint64_t n, const uint32_t* h_rec_entry, uint64_t n_entries,
                            int32_t* d_sketches, uint32_t* d_flags) {
  NQ_TRY(nq_params_check(p));
  if (!h_rec_entry) n_entries = n;
  if (n_entries == 0) return NQ_OK;
  if (n >= (1ull << 31) || n_entries >= (1ull << 31))
    return nq_set_error(NQ_ERR_INVALID, "too many entries in one batch: %llu", (unsigned long long)n);
  DevParams P = nq_make_dev_params(p);
  const size_t cells = (size_t)n_entries * P.F;
  NQ_CUDA(cudaMemsetAsync(d_sketches, 0xFF, cells * sizeof(int32_t), ctx->stream));
  if (d_flags) NQ_CUDA(cudaMemsetAsync(d_flags, 0, n_entries * sizeof(uint32_t), ctx->stream));
  uint64_t total_k = 0;
  for (uint64_t e = 0; e < n; ++e) {
    const uint64_t len = h_offsets[e + 1] - h_offsets[e];
    if (len > p->K) total_k += len - p->K;
  }
  if (total_k == 0) return nq_launch_densify(ctx, p, d_sketches, n_entries, d_flags);
  uint64_t span_len = (total_k / ((uint64_t)ctx->sm_count * 8) + 1023) & ~1023ull;
  span_len = std::max<uint64_t>(span_len, 65536);
  std::vector<Span> spans;
  for (uint64_t e = 0; e < n; ++e) {
    const uint64_t e0 = h_offsets[e], len = h_offsets[e + 1] - e0;
    if (len <= p->K) continue;
    const uint64_t nk = len - p->K;
    const uint64_t parts = (nk + span_len - 1) / span_len;
    const uint64_t each = ((nk + parts - 1) / parts + 1023) & ~1023ull;
    for (uint64_t a = 0; a < nk; a += each)
      spans.push_back(Span{e0 + a, e0 + std::min(nk, a + each), e0, h_rec_entry ? h_rec_entry[e] : (uint32_t)e, 0});
  }
  Span* d_spans = nullptr;
  NQ_TRY(nq_dmalloc(ctx, (void**)&d_spans, spans.size() * sizeof(Span)));
  {
    const int us = nq_upload_small(ctx, d_spans, spans.data(), spans.size() * sizeof(Span));  // pinned ring: no host stall
    if (us != NQ_OK) {
      nq_dfree(ctx, d_spans);
      return us;
    }
  }
  const PackedIn in{d_codes, d_blk, d_pool};
  uint32_t* sk = reinterpret_cast<uint32_t*>(d_sketches);
  const bool fits = 2048 + 1024 + (size_t)P.F * 4 <= ctx->smem_optin;
  const bool small = total_k / std::max<uint64_t>(spans.size(), 1) < 16384;
  int st;
  if (fits) {
    st = small ? launch_packed<true, 128>(ctx, P, in, nullptr, d_spans, spans.size(), sk)
               : launch_packed<true, 1024>(ctx, P, in, nullptr, d_spans, spans.size(), sk);
  } else {
    const size_t room = ctx->smem_optin - 2048 - 1024;
    if (!small) {
      if (P.W >= 8 && (size_t)P.F <= room) P.filter = 8;
      else if (P.W >= 4 && (size_t)P.F / 2 <= room) P.filter = 4;
    }
    if (P.filter == 0) P.filter |= 0x100u;  // no coarse filter: read the cell before the atomic (sketch_common.cuh)
    st = small ? launch_packed<false, 128>(ctx, P, in, nullptr, d_spans, spans.size(), sk)
               : launch_packed<false, 1024>(ctx, P, in, nullptr, d_spans, spans.size(), sk);
  }
  nq_dfree(ctx, d_spans);
  NQ_TRY(st);
  return nq_launch_densify(ctx, p, d_sketches, n_entries, d_flags);
}
