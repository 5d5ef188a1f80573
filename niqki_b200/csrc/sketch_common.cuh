// sketch_common.cuh — pieces shared by the sketch kernels (sketch.cu: ASCII input; sketch_packed.cu:
// 2-bit packed input).
#pragma once
#include "device_common.cuh"

namespace nq {

struct Span {
  uint64_t kb;     // first k-mer start of the span (byte offset into the bases buffer)
  uint64_t ke;     // one past the last k-mer start
  uint64_t e0;     // first byte of the entry
  uint32_t entry;  // sketch row
  uint32_t pad;
};

struct KmerState {
  uint64_t f, r;
};

template <bool SMEM>
struct SketchSink {
  uint32_t* sk;  // shared (SMEM) or global sketch row
  // global form only: coarse shared-memory filter, `fbits` (0, 4 or 8) bits per cell holding an upper
  // bound of the top bits (fp >> cshift) of the smallest fingerprint this CTA has SUBMITTED to the
  // cell.  A k-mer whose top bits are above it cannot lower the cell and skips the global access.
  // Updates are plain (racy) stores: a lost or stale update only leaves a bound that is too high,
  // i.e. a more permissive filter — every value ever stored is the top bits of a submitted fingerprint.
  uint8_t* filt;
  uint32_t fbits, cshift;
  uint32_t gread;  // global form: read the cell before the atomic (only without the coarse filter, see update())
  __device__ __forceinline__ void update(uint32_t b, uint32_t fp) const {
    // sketch[b] = min(sketch[b], fp) with empty = 0xFFFFFFFF (:350-355).
    if (SMEM) {
      // shared memory: a fire-and-forget ATOMS.MIN per k-mer is cheaper than read + compare +
      // conditional atomic (measured on B200: 563 vs 517 Gbases/s)
      atomicMin(&sk[b], fp);
    } else {
      if (fbits == 8) {
        const uint32_t c = min(fp >> cshift, 255u), cur = filt[b];
        if (c > cur) return;
        if (c < cur) filt[b] = (uint8_t)c;
      } else if (fbits == 4) {
        const uint32_t byte = filt[b >> 1], sh = (b & 1u) * 4u, cur = (byte >> sh) & 15u, c = min(fp >> cshift, 15u);
        if (c > cur) return;
        if (c < cur) filt[b >> 1] = (uint8_t)((byte & ~(15u << sh)) | (c << sh));
      }
      // global memory.  Without the coarse filter a plain read keeps the k-mers that cannot lower the cell
      // away from the L2 atomic unit (a stale read only makes that conservative: cells never increase).
      // Behind the coarse filter only ~20 % of the k-mers get here, and the read is what hurts: a dependent
      // L2 round trip in front of every survivor, three per 16 bases and warp — so they go out as
      // fire-and-forget RED.MIN instead.
      if (gread) {
        if (fp < sk[b]) atomicMin(&sk[b], fp);
      } else {
        atomicMin(&sk[b], fp);
      }
    }
  }
};

// (hi:lo) * (Ch:Cl) mod 2^64 as a chain of three multiply-adds with 32-bit addends (IMAD.WIDE,
// IMAD, IMAD): no zeroed register pair, no separate add.  PTX pins the association.
__device__ __forceinline__ uint2 mul64c(uint32_t hi, uint32_t lo, uint32_t Ch, uint32_t Cl) {
  uint32_t plo, phi;
  asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0, %1}, p;\n\t}" : "=r"(plo), "=r"(phi) : "r"(lo), "r"(Cl));
  asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(phi) : "r"(lo), "r"(Ch));
  asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(phi) : "r"(hi), "r"(Cl));
  return make_uint2(plo, phi);
}
// high word only: IMAD.HI, IMAD, IMAD
__device__ __forceinline__ uint32_t mul64c_hi_chain(uint32_t hi, uint32_t lo, uint32_t Ch, uint32_t Cl) {
  uint32_t r;
  asm("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(lo), "r"(Cl));
  asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r) : "r"(lo), "r"(Ch));
  asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r) : "r"(hi), "r"(Cl));
  return r;
}

}  // namespace nq
