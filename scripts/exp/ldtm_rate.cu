// Experiment: sustained tcgen05.ld (TMEM -> registers) rate per SM, 32x32b shape, x16 / x32 / x64 per instruction,
// with 4 / 8 / 16 reading warps (1 / 2 / 4 per TMEM lane quarter) and 1 or 2 loads in flight before tcgen05.wait::ld.
// Also: the same loop followed by the bf16 pack + 128-byte-swizzled st.shared.v4 of the conv epilogue's phase 1.
// Question it answers: is the fprop epilogue (TMEM -> bf16 staging tile) bound by TMEM read bandwidth or by issue?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ldtm_rate scripts/exp/ldtm_rate.cu && ./ldtm_rate
#include "../../climategan_b200/csrc/conv_tc.cu"
#include <cstdio>
#include <vector>
using namespace cgb;
namespace cgb {   // the two symbols conv_tc.cu takes from api.cu
std::atomic<int64_t> g_launches{0};
void set_error(const char*, ...) {}
}

template <int X>
__device__ __forceinline__ void ldtm(uint32_t taddr, uint32_t* r);
template <>
__device__ __forceinline__ void ldtm<16>(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
template <>
__device__ __forceinline__ void ldtm<32>(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// mode 0: loads only (results xor-folded so they are not dead); mode 1: + pack to bf16 + swizzled st.shared.v4
template <int X, int MODE>
__global__ void __launch_bounds__(512) ldtm_kernel(int warps, int inflight, int iters, long long* out, uint32_t* sink) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t tptr = base + 65536;
  volatile uint32_t* tptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tptr - raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(tptr, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem_base = *tptr_gen;
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < warps) {
    const int q = warp & 3, g = warp >> 2, per_q = warps >> 2;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    const int row = q * 32 + lane;
    const uint32_t row_u32 = base + (uint32_t)row * 128u, sw16 = (uint32_t)(row & 7) << 4;
    const int nchunks = 256 / X;   // a 128 x 256 fp32 accumulator tile
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int c = g; c < nchunks; c += per_q * inflight) {
        uint32_t ra[X], rb[X];
        ldtm<X>(t_row + (uint32_t)(c * X), ra);
        if (inflight == 2 && c + per_q < nchunks) ldtm<X>(t_row + (uint32_t)((c + per_q) * X), rb);
        tmem_ld_wait();
        if (MODE == 0) {
#pragma unroll
          for (int j = 0; j < X; ++j) acc ^= ra[j];
          if (inflight == 2 && c + per_q < nchunks) {
#pragma unroll
            for (int j = 0; j < X; ++j) acc ^= rb[j];
          }
        } else {
#pragma unroll
          for (int h = 0; h < X / 8; ++h) {
            const uint32_t j = (uint32_t)(c * (X / 8) + h);
            sts128(row_u32 + (j >> 3) * 16384u + (((j & 7u) << 4) ^ sw16),
                   pack2<__nv_bfloat16>(__uint_as_float(ra[8 * h]), __uint_as_float(ra[8 * h + 1])),
                   pack2<__nv_bfloat16>(__uint_as_float(ra[8 * h + 2]), __uint_as_float(ra[8 * h + 3])),
                   pack2<__nv_bfloat16>(__uint_as_float(ra[8 * h + 4]), __uint_as_float(ra[8 * h + 5])),
                   pack2<__nv_bfloat16>(__uint_as_float(ra[8 * h + 6]), __uint_as_float(ra[8 * h + 7])));
          }
          if (inflight == 2 && c + per_q < nchunks) {
#pragma unroll
            for (int h = 0; h < X / 8; ++h) {
              const uint32_t j = (uint32_t)((c + per_q) * (X / 8) + h);
              sts128(row_u32 + (j >> 3) * 16384u + (((j & 7u) << 4) ^ sw16),
                     pack2<__nv_bfloat16>(__uint_as_float(rb[8 * h]), __uint_as_float(rb[8 * h + 1])),
                     pack2<__nv_bfloat16>(__uint_as_float(rb[8 * h + 2]), __uint_as_float(rb[8 * h + 3])),
                     pack2<__nv_bfloat16>(__uint_as_float(rb[8 * h + 4]), __uint_as_float(rb[8 * h + 5])),
                     pack2<__nv_bfloat16>(__uint_as_float(rb[8 * h + 6]), __uint_as_float(rb[8 * h + 7])));
          }
        }
        }
      }
    }
    t1 = clock64();
  }
  if (acc == 0x12345678u) sink[0] = acc;
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <int X, int MODE>
static void run(int warps, int inflight, long long* d_out, uint32_t* d_sink) {
  const int iters = 200;
  const size_t smem = 65536 + 1024 + 64;
  cudaFuncSetAttribute(ldtm_kernel<X, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  ldtm_kernel<X, MODE><<<148, 512, smem>>>(warps, inflight, iters, d_out, d_sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
  std::vector<long long> h(148);
  cudaMemcpy(h.data(), d_out, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (auto v : h) avg += (double)v;
  avg /= 148.0 * iters;
  printf("x%-3d %-22s warps=%2d inflight=%d : %8.0f cycles per 128x256 fp32 tile (128 KB) = %6.1f B/clk/SM\n", X,
         MODE ? "ld + pack + st.shared" : "ld only", warps, inflight, avg, 131072.0 / avg);
}

int main() {
  long long* d_out; uint32_t* d_sink;
  cudaMalloc(&d_out, 148 * sizeof(long long));
  cudaMalloc(&d_sink, 64);
  for (int warps : {4, 8, 16})
    for (int inflight : {1, 2}) {
      run<16, 0>(warps, inflight, d_out, d_sink);
      run<32, 0>(warps, inflight, d_out, d_sink);
      run<16, 1>(warps, inflight, d_out, d_sink);
      run<32, 1>(warps, inflight, d_out, d_sink);
    }
  return 0;
}
