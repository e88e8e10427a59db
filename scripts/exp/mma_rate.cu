// Experiment: sustained tcgen05.mma rate (cycles per instruction) for M=128, K=16 bf16, N in {16..256}, operands in
// smem (SS mode, K-major SW128), one CTA per SM, all MMAs accumulate into the same TMEM tile.
// Variants: sbo=1024 (dense rows) vs sbo=1280 (halo-tile addressing); A start shifted by one row.
#include "../../climategan_b200/csrc/conv_tc.cu"
#include <vector>
#include <cstdio>
using namespace cgb;

__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}

template <int V>
__global__ void __launch_bounds__(128)
rate_kernel(int n, int iters, int sbo, int shift, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t a_addr = base;            // 64 KB region for A (garbage data is fine)
  const uint32_t b_addr = base + 65536;    // 32 KB for B
  const uint32_t bar = base + 65536 + 32768;
  const uint32_t tptr = bar + 8;
  volatile uint32_t* tptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tptr - raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // zero smem so no NaN slow paths
  for (int i = threadIdx.x; i < (65536 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (base - raw))[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tptr, 256);
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem_base = *tptr_gen;
  if (warp == 1) {
    const uint32_t idesc = make_idesc(n, false, false);
    const uint32_t hi_a = desc_hi((uint32_t)sbo), hi_b = desc_hi(1024u);
    const uint32_t elected = lane == 0;
    const uint32_t a_lo = desc_lo(a_addr + shift * 128, 16), b_lo = desc_lo(b_addr, 16);
    long long t0 = clock64();
    if (V == 0) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_elect(tmem_base, desc_join(a_lo + 2u * k, hi_a), desc_join(b_lo + 2u * k, hi_b), idesc, 1u, elected);
      }
      umma_commit_elect(bar, elected);
    } else {
      if (elect_one_sync()) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base, desc_join(a_lo + 2u * k, hi_a), desc_join(b_lo + 2u * k, hi_b), idesc, 1u);
        }
        umma_commit(bar);
      }
      __syncwarp();
    }
    mbar_wait(bar, 0);
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

int main() {
  long long* dO; cudaMalloc(&dO, 148 * 8);
  cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  const int iters = 2000;
  for (int ctas : {148})
    for (int cfg = 0; cfg < 2; ++cfg) {
      const int sbo = 1280, shift = 1;
      for (int n : {16, 48, 64, 96, 128, 160, 256}) {
        if (cfg == 0) rate_kernel<0><<<ctas, 128, 100 * 1024 + 2048>>>(n, iters, sbo, shift, dO);
        else rate_kernel<1><<<ctas, 128, 100 * 1024 + 2048>>>(n, iters, sbo, shift, dO);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<long long> h(ctas);
        cudaMemcpy(h.data(), dO, ctas * 8, cudaMemcpyDeviceToHost);
        long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
        printf("variant=%d ctas=%3d sbo=%4d shift=%d N=%3d : %.1f cycles/MMA (floor %d, smem-read bound %.0f)\n", cfg, ctas, sbo, shift, n,
               (double)mx / (iters * 4), n / 2, (4096.0 + n * 32) / 128);
      }
    }
  return 0;
}
