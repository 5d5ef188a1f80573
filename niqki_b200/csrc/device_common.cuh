// device_common.cuh — device-side arithmetic of the NIQKI hot path (sm_100a).
// Reference lines are relative to /root/reference/.
#pragma once
#include <cstdint>

namespace nq {

constexpr uint32_t kEmpty = 0xFFFFFFFFu;  // int32 -1: an empty sketch cell (src/niqki_index.cpp:337)

// Scalar block handed to kernels by value (mirror of nq_params, pre-digested).
struct DevParams {
  uint32_t K, S, W, M;
  uint32_t F;
  uint32_t mask_M;
  uint32_t maxrem;
  uint32_t range;
  uint64_t kmask;     // 4^K - 1  (min %= offsetUpdatekmer, :228)
  uint32_t rc_shift;  // 2K-2     (:235)
  uint32_t filter;    // scan kernel, sketch in HBM: bits 0-7 = bits per cell of the coarse shared-memory filter (0, 4, 8), bit 8 = read the cell before the atomicMin
};

// x = ((x>>32)^x) * C  — one round of the xorshift-multiply mixers (:292-293, :301-302)
__device__ __forceinline__ uint64_t fold_mul(uint64_t x, uint64_t c) { return ((x >> 32) ^ x) * c; }

constexpr uint64_t kRevC = 0xD6E8FEB86659FD93ull;    // revhash64   (:291-296)
constexpr uint64_t kUnrevC = 0xCFEE444D8B59A89Bull;  // unrevhash64 (:300-305)

__device__ __forceinline__ uint64_t revhash64(uint64_t x) {
  x = fold_mul(fold_mul(x, kRevC), kRevC);
  return (x >> 32) ^ x;
}
__device__ __forceinline__ uint64_t unrevhash64(uint64_t x) {
  x = fold_mul(fold_mul(x, kUnrevC), kUnrevC);
  return (x >> 32) ^ x;
}
// ---- forms used by the scan kernel (same arithmetic, fewer instructions) -----------------------
// high word of (hi:lo)*(Ch:Cl) only: mulhi(lo,Cl) + lo*Ch + hi*Cl  -> 2 IMAD + 1 IMAD.HI
__device__ __forceinline__ uint32_t mul64c_hi(uint32_t hi, uint32_t lo, uint32_t Ch, uint32_t Cl) {
  return __umulhi(lo, Cl) + (lo * Ch + hi * Cl);
}
// High 32 bits of unrevhash64(x): the final fold leaves the high word untouched, and the bucket
// id (:347) only needs `hash >> (64-S)` with S <= 31.
__device__ __forceinline__ uint32_t unrevhash64_hi32(uint64_t x) {
  constexpr uint32_t Ch = (uint32_t)(kUnrevC >> 32), Cl = (uint32_t)kUnrevC;
  x = fold_mul(x, kUnrevC);
  const uint32_t hi = (uint32_t)(x >> 32), lo = (uint32_t)x ^ hi;
  return mul64c_hi(hi, lo, Ch, Cl);
}
// get_fingerprint (:277-287) from the two words of the hash.  With maxrem <= 32 only the high
// word's leading zeros matter: hi == 0 gives clz >= 32 >= maxrem, i.e. a zero HLL part either way.
template <bool SMALL_REM>
__device__ __forceinline__ uint32_t fingerprint32(uint32_t hi, uint32_t lo, uint32_t mask_M, uint32_t maxrem, uint32_t M) {
  int lz;
  if (SMALL_REM) lz = __clz((int)hi);
  else lz = hi ? __clz((int)hi) : 32 + __clz((int)lo);
  const int rem = max(0, (int)maxrem - lz);
  return (lo & mask_M) + ((uint32_t)rem << M);
}

// get_fingerprint (:277-287) with clz64(0) := 64 (bsr on 0 is UB in the reference; observed fp=0)
__device__ __forceinline__ uint32_t fingerprint(uint64_t h, uint32_t mask_M, uint32_t maxrem, uint32_t M) {
  const int lz = __clzll((long long)h);  // 64 for h == 0
  const int rem = max(0, (int)maxrem - lz);
  return ((uint32_t)h & mask_M) + ((uint32_t)rem << M);
}

// hash_family(x, step) % F for F a power of two (:308-310, :319): both terms can be reduced first.
__device__ __forceinline__ uint32_t family_a(uint32_t v, uint32_t Fmask) { return (uint32_t)unrevhash64(v) & Fmask; }
__device__ __forceinline__ uint32_t family_b(uint32_t v, uint32_t Fmask) { return (uint32_t)revhash64(v) & Fmask; }

// splitmix-style mixer of the synthetic generator
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

}  // namespace nq
