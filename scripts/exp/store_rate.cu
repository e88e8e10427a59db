// Experiment: how fast can the SMs write a conv output tile stream to global memory?
// Output tensor [P = 51200 pixels][C channels] bf16 (the ResNet layer-3 1x1 conv: C = 1024; the SPADE mlp_shared: C = 128 at
// 3.3 M pixels is the same pattern, more rows).  Every CTA walks its 128-pixel x 256-channel tiles (n-tile fastest, like the
// streaming kernel) and writes each one from shared memory, nothing else:
//   mode 0: 4 TMA bulk tensor stores of {64 ch, 128 px} (128B swizzle) by ONE thread, wait_group.read 0 before the next tile
//   mode 1: the same 4 stores by FOUR threads (one per half), each waiting for its own group before its next store
//   mode 2: like 1, but the wait happens one tile later (2 staging buffers: wait_group.read 1)
//   mode 3: per-thread 16-byte st.global from the staging tile (512 threads, consecutive lanes = consecutive chunks of a pixel)
//   mode 4: per-thread st.global straight from registers, thread = pixel row, 32 B (16 channels) per store pair (the "no
//           staging" epilogue)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Iinclude -lcuda -o scripts/exp/store_rate scripts/exp/store_rate.cu
#include "../../climategan_b200/csrc/conv_tc.cu"
#include <cstdio>
#include <vector>
using namespace cgb;
namespace cgb {
std::atomic<int64_t> g_launches{0};
void set_error(const char*, ...) {}
}

__global__ void __launch_bounds__(512) store_kernel(const __grid_constant__ CUtensorMap tmY, __nv_bfloat16* y, int mode, int pix_tiles,
                                                   int n_tiles, int C) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 2 * 65536 / 4; i += 512) reinterpret_cast<uint32_t*>(gen)[i] = 0x3f803f80u;
  fence_proxy_async_smem();
  __syncthreads();
  const int total = pix_tiles * n_tiles;
  int lt = 0;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++lt) {
    const int nt = tile % n_tiles, pt = tile / n_tiles;
    const int cn0 = nt * 256, p0 = pt * 128;
    if (mode == 0) {
      if (threadIdx.x == 0) {
        tma_store_wait_read<0>();
        for (int hb = 0; hb < 4; ++hb) tma_store_4d(&tmY, base + hb * 16384u, cn0 + hb * 64, p0, 0, 0);
        tma_store_commit();
      }
    } else if (mode == 1) {
      if (lane == 0 && warp < 4) {
        tma_store_wait_read<0>();
        tma_store_4d(&tmY, base + warp * 16384u, cn0 + warp * 64, p0, 0, 0);
        tma_store_commit();
      }
    } else if (mode == 2) {
      if (lane == 0 && warp < 4) {
        tma_store_wait_read<1>();
        tma_store_4d(&tmY, base + (lt & 1) * 65536u + warp * 16384u, cn0 + warp * 64, p0, 0, 0);
        tma_store_commit();
      }
    } else if (mode == 3) {
      // 128 rows x 32 chunks of 16 B: thread t -> chunk t & 31 of rows (t >> 5) + 16 k
      const int c = threadIdx.x & 31;
      for (int r = threadIdx.x >> 5; r < 128; r += 16) {
        const uint4 v = *reinterpret_cast<const uint4*>(gen + r * 512 + c * 16);
        *reinterpret_cast<uint4*>(y + (size_t)(p0 + r) * C + cn0 + c * 8) = v;
      }
    } else {
      // thread = row (4 warps per lane quarter share the 16 chunks of 16 columns), 2 x 16 B per chunk
      const int q = warp & 3, g = warp >> 2, row = q * 32 + lane;
      const uint4 v = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
      for (int c = g; c < 16; c += 4) {
        __nv_bfloat16* dst = y + (size_t)(p0 + row) * C + cn0 + c * 16;
        *reinterpret_cast<uint4*>(dst) = v;
        *reinterpret_cast<uint4*>(dst + 8) = v;
      }
    }
  }
  if (mode <= 2 && lane == 0 && warp < 4) tma_store_wait_all();
}

int main() {
  const int P = 51200, C = 1024;
  __nv_bfloat16* y;
  cudaMalloc(&y, (size_t)P * C * 2);
  CUtensorMap tmY;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)P, 1, 1};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)P * C * 2, (cuuint64_t)P * C * 2};
  cuuint32_t box[4] = {64, 128, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if (!encode_map(&tmY, y, 4, dims, strides, box, estr, "y")) { printf("encode failed\n"); return 1; }
  const size_t smem = 2 * 65536 + 1024;
  cudaFuncSetAttribute(store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[5] = {"TMA store, 1 thread, 4 halves", "TMA store, 4 leaders", "TMA store, 4 leaders, 2 buffers", "st.global.v4 from staging (coalesced)",
                          "st.global.v4 from registers (row per thread)"};
  for (int mode = 0; mode < 5; ++mode) {
    for (int grid : {148, 296}) {
      if (grid == 296 && mode != 3 && mode != 4) continue;
      store_kernel<<<148, 512, smem>>>(tmY, y, mode, P / 128, C / 256, C);
      cudaDeviceSynchronize();
      cudaEventRecord(e0);
      const int reps = 20;
      for (int r = 0; r < reps; ++r) store_kernel<<<grid == 148 ? 148 : 148, 512, smem>>>(tmY, y, mode, P / 128, C / 256, C);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      ms /= reps;
      printf("mode %d %-46s: %7.1f us  %6.0f GB/s written  (%s)\n", mode, names[mode], ms * 1e3, (double)P * C * 2 / ms / 1e6,
             cudaGetErrorString(e));
      break;
    }
  }
  return 0;
}
