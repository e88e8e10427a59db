// Experiment: can a K-major SWIZZLE_128B UMMA operand start at a 128-byte (one row) offset inside a TMA-written
// tile?  D[m][n] = sum_k A[m+s][k] * I[n][k] = A[m+s][n].  Tries base_offset = 0 and base_offset = s & 7.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -o shift_desc shift_desc.cu ../../climategan_b200/csrc/{api,ops,conv_simt}.cu
#include "../../climategan_b200/csrc/conv_tc.cu"
#include <vector>
#include <cstdio>
using namespace cgb;

__global__ void __launch_bounds__(128)
shift_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int shift_rows, int base_off,
             int mn_major, float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t a_addr = base;                 // 256 rows x 128 B = 32 KB
  const uint32_t b_addr = base + 32768;         // 64 rows x 128 B
  const uint32_t bar = base + 32768 + 8192;
  const uint32_t bar2 = bar + 8;
  const uint32_t tptr = bar + 16;
  volatile uint32_t* tptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tptr - raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tptr, 64);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem_base = *tptr_gen;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 32768 + 8192);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(a_addr), "l"(reinterpret_cast<uint64_t>(&tmA)), "r"(bar), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(b_addr), "l"(reinterpret_cast<uint64_t>(&tmB)), "r"(bar), "r"(0), "r"(0) : "memory");
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  if (threadIdx.x == 0) {
    if (!mn_major) {
      const uint32_t idesc = make_idesc(64, false, false);
      for (int k = 0; k < 4; ++k) {
        uint64_t ad = make_desc(a_addr + shift_rows * 128 + k * 32, 16, 1024) | ((uint64_t)(base_off & 7) << 49);
        uint64_t bd = make_desc(b_addr + k * 32, 16, 1024);
        umma_bf16(tmem_base, ad, bd, idesc, k > 0);
      }
    } else {
      // MN-major A: rows of the smem tile are K indices (pixels), 64 contiguous M elements per row; M=128 needs two
      // 64-wide column blocks: we only have one (64 ch) so rows m>=64 of D are garbage; K = 16 rows per step.
      // D[m][n] = sum_k A[k + s][m] * B[k][n] with B = rows of identity (MN-major B: B[k][n], 64 n per row).
      const uint32_t idesc = make_idesc(64, true, true);
      for (int k = 0; k < 4; ++k) {   // K = 64 "pixels"
        uint64_t ad = make_desc(a_addr + shift_rows * 128 + k * 2048, 16384, 1024) | ((uint64_t)(base_off & 7) << 49);
        uint64_t bd = make_desc(b_addr + k * 2048, 16384, 1024);
        umma_bf16(tmem_base, ad, bd, idesc, k > 0);
      }
    }
    umma_commit(bar2);
  }
  mbar_wait(bar2, 0);
  tc_fence_after();
  const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(t_row + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

int main() {
  const int ROWS = 256;
  std::vector<__nv_bfloat16> hA(ROWS * 64), hB(64 * 64);
  for (int r = 0; r < ROWS; ++r) for (int c = 0; c < 64; ++c) hA[r * 64 + c] = __float2bfloat16((float)((r * 7 + c * 3) % 251));
  for (int n = 0; n < 64; ++n) for (int k = 0; k < 64; ++k) hB[n * 64 + k] = __float2bfloat16(n == k ? 1.f : 0.f);
  __nv_bfloat16 *dA, *dB; float* dO;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dO, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmB;
  { cuuint64_t dims[2] = {64, (cuuint64_t)ROWS}; cuuint64_t str[1] = {128}; cuuint32_t box[2] = {64, 256}; cuuint32_t es[2] = {1, 1};
    if (!encode_map(&tmA, dA, 2, dims, str, box, es, "A")) { printf("encode A failed: %s\n", cgb_last_error()); return 1; } }
  { cuuint64_t dims[2] = {64, 64}; cuuint64_t str[1] = {128}; cuuint32_t box[2] = {64, 64}; cuuint32_t es[2] = {1, 1};
    if (!encode_map(&tmB, dB, 2, dims, str, box, es, "B")) { printf("encode B failed\n"); return 1; } }
  cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> hO(128 * 64);
  for (int mn = 0; mn < 2; ++mn)
    for (int s : {0, 1, 2, 3, 5, 8, 9, 17}) {
      for (int mode = 0; mode < 2; ++mode) {
        const int bo = mode == 0 ? 0 : (s & 7);
        if (mode == 1 && bo == 0) continue;
        cudaMemset(dO, 0, 128 * 64 * 4);
        shift_kernel<<<1, 128, 48 * 1024, 0>>>(tmA, tmB, s, bo, mn, dO);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mn=%d s=%d bo=%d: CUDA error %s\n", mn, s, bo, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0, first_bad = -1;
        const int mrows = mn ? 64 : 128;
        for (int m = 0; m < mrows; ++m) for (int n = 0; n < 64; ++n) {
          // K-major: D[m][n] = A[m+s][n].   MN-major: D[m][n] = sum_k A[k+s][m] * B[k][n] = A[n+s][m]
          float exp = mn ? __bfloat162float(hA[(n + s) * 64 + m]) : __bfloat162float(hA[(m + s) * 64 + n]);
          if (hO[m * 64 + n] != exp) { if (first_bad < 0) first_bad = m * 64 + n; ++bad; }
        }
        printf("%s-major shift=%2d base_offset=%d : %s (%d mismatches, first at m=%d n=%d)\n", mn ? "MN" : "K ", s, bo,
               bad ? "WRONG" : "OK", bad, first_bad / 64, first_bad % 64);
      }
    }
  return 0;
}
