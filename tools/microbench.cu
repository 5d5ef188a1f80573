// microbench.cu — measures the primitives the K4a/K2 designs lean on (B200): shared-memory
// atomicAdd with random addresses, global RED with random addresses in an L2-resident table,
// __match_any_sync, and shared atomicMin after a filtering read.  Prints ops/s.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t xs(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

template <bool PACK>
__global__ void k_smem_atomic(uint32_t words, int iters, uint32_t* sink) {
  extern __shared__ uint32_t sm[];
  for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) sm[i] = 0;
  __syncthreads();
  uint32_t s = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
  for (int i = 0; i < iters; ++i) {
    const uint32_t r = xs(s);
    const uint32_t l = r % (PACK ? words * 2 : words);
    if (PACK) atomicAdd(&sm[l >> 1], 1u << ((l & 1) * 16)); else atomicAdd(&sm[l], 1u);
  }
  __syncthreads();
  if (threadIdx.x == 0) sink[blockIdx.x] = sm[0];
}

__global__ void k_red_global(uint32_t* tab, uint32_t words, int iters) {
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 1u;
  for (int i = 0; i < iters; ++i) atomicAdd(&tab[xs(s) % words], 1u);
}

__global__ void k_match(int iters, uint32_t* sink) {
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 1u, acc = 0;
  for (int i = 0; i < iters; ++i) acc += __popc(__match_any_sync(0xFFFFFFFFu, xs(s) & 4095u));
  if (acc == 0xDEADBEEF) sink[0] = acc;
}

__global__ void k_smem_min_filtered(uint32_t words, int iters, uint32_t* sink) {
  extern __shared__ uint32_t sm[];
  for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) sm[i] = 0xFFFFFFFFu;
  __syncthreads();
  uint32_t s = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
  for (int i = 0; i < iters; ++i) {
    const uint32_t b = xs(s) % words, v = xs(s) >> 20;
    if (v < sm[b]) atomicMin(&sm[b], v);
  }
  __syncthreads();
  if (threadIdx.x == 0) sink[blockIdx.x] = sm[0];
}

// INT32 issue rate of an SM (the roofline denominator of the sketch scan): chains of independent integer
// instructions, 8 per thread, timed with clock64() on the device so that the result is ops per clock
// and does not depend on the clock the box happens to run at.  MIX 0: LOP3 + IADD3 (ALU pipe),
// 1: IMAD (FMA pipe), 2: 3 IMAD : 4 ALU, close to the mix of the scan kernel's inner loop (15 : 19).
template <int MIX>
__global__ void __launch_bounds__(1024) k_int_rate(int iters, uint32_t seed, uint32_t* sink, unsigned long long* cycles) {
  uint32_t x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = seed + threadIdx.x * 8 + i;
  const uint32_t c = seed | 1u, d = seed * 3u + 7u;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MIX == 0) {           // 2 ops
        x[i] = (x[i] ^ d) + c;
      } else if (MIX == 1) {    // 1 op
        x[i] = x[i] * c + d;
      } else {                  // 7 instructions: 3 IMAD, SHF, 2 LOP3, IADD3 (checked in the SASS)
        x[i] = x[i] * c + d;
        x[i] ^= x[i] >> 3;
        x[i] = x[i] * d + c;
        x[i] = (x[i] & d) | c;
        x[i] = x[i] * c + x[(i + 1) & 7];
        x[i] = x[i] + __brev(d);
      }
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc ^= x[i];
  if (acc == 0x12345678u) sink[0] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <typename F>
static float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  printf("device %s, %d SMs, smem optin %zu\n", p.name, sms, p.sharedMemPerBlockOptin);
  uint32_t* sink; cudaMalloc(&sink, 1 << 20);
  const int iters = 4096;
  for (int nt : {256, 512, 1024}) {
    for (uint32_t words : {5000u, 50000u}) {
      const size_t smem = words * 4;
      cudaFuncSetAttribute(k_smem_atomic<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaFuncSetAttribute(k_smem_atomic<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      float ms = timeit([&] { k_smem_atomic<true><<<sms, nt, smem>>>(words, iters, sink); });
      printf("smem atomicAdd packed16  nt=%4d words=%6u : %8.1f Gops/s\n", nt, words, (double)sms * nt * iters / ms / 1e6);
      ms = timeit([&] { k_smem_atomic<false><<<sms, nt, smem>>>(words, iters, sink); });
      printf("smem atomicAdd u32       nt=%4d words=%6u : %8.1f Gops/s\n", nt, words, (double)sms * nt * iters / ms / 1e6);
    }
  }
  {
    const uint32_t words = 32768; const size_t smem = words * 4;
    cudaFuncSetAttribute(k_smem_min_filtered, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float ms = timeit([&] { k_smem_min_filtered<<<sms, 1024, smem>>>(words, iters, sink); });
    printf("smem filtered atomicMin  nt=1024 words=%6u : %8.1f Gops/s\n", words, (double)sms * 1024 * iters / ms / 1e6);
  }
  for (uint32_t words : {5u << 20, 64u << 20}) {  // 20 MB (L2 resident) and 256 MB
    uint32_t* tab; cudaMalloc(&tab, (size_t)words * 4); cudaMemset(tab, 0, (size_t)words * 4);
    float ms = timeit([&] { k_red_global<<<sms * 8, 256>>>(tab, words, 1024); });
    printf("global RED.ADD random    table=%4u MB      : %8.1f Gops/s\n", words >> 18, (double)sms * 8 * 256 * 1024 / ms / 1e6);
    cudaFree(tab);
  }
  {
    unsigned long long* cyc; cudaMalloc(&cyc, sms * 2 * sizeof(unsigned long long));
    unsigned long long h[2048];
    const int it = 2048;
    auto rate = [&](auto kern, int ops_per_elem, const char* what) {
      kern<<<sms * 2, 1024>>>(it, 12345u, sink, cyc); cudaDeviceSynchronize();
      kern<<<sms * 2, 1024>>>(it, 12345u, sink, cyc); cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, sms * 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
      double mean = 0; for (int i = 0; i < sms * 2; ++i) mean += (double)h[i]; mean /= sms * 2;
      // two 1024-thread CTAs share an SM for (about) the same `mean` cycles
      const double per_clk = 2.0 * 1024 * 8.0 * ops_per_elem * it / mean;
      printf("int32 %-28s : %6.1f ops/clk/SM\n", what, per_clk);
      return per_clk;
    };
    rate(k_int_rate<0>, 2, "LOP3+IADD3 (ALU pipe)");
    rate(k_int_rate<1>, 1, "IMAD (FMA pipe)");
    const double mix = rate(k_int_rate<2>, 7, "3 IMAD : 4 ALU (scan mix)");
    printf("{\"int32_ops_per_clk_per_sm\": %.2f, \"source\": \"tools/microbench k_int_rate<2>: 3 IMAD : 4 ALU instructions, 2 x 1024 threads per SM, clock64\"}\n", mix);
    cudaFree(cyc);
  }
  float ms = timeit([&] { k_match<<<sms * 8, 256>>>(1024, sink); });
  printf("__match_any_sync                              : %8.1f G lane-ops/s\n", (double)sms * 8 * 256 * 1024 / ms / 1e6);
  return 0;
}
