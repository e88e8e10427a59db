// tcgen05 engine: TMA-fed implicit-GEMM convolution on the 5th-gen tensor cores (sm_100a).
//
//   GEMM view (fprop, and dgrad of stride-1 convs run as an fprop over gy with flipped/transposed weights):
//     D[M = 128 output pixels][N = out channels] += A[M][K] * B[N][K]^T ,  K = taps x in-channels
//   * A operand: for every filter tap, ONE 4-D TMA box {64 ch, TW*s, TH*s, TN} (element strides s) of the
//     NHWC activation tensor, shifted by the tap offset.  Out-of-bounds coordinates (the zero padding, ragged
//     tile edges, channel tails) are zero-filled by TMA, so there is no padded copy and no predicate in the
//     loader (replaces ZeroPad2d + im2col of the reference's cuDNN path, climategan/blocks.py:66-71,117-144).
//   * B operand: 3-D TMA box {64 ch, 1 tap, BN out-channels} of the packed weights [co][tap][ci].
//   * Both land in shared memory in the 128-byte-swizzled K-major layout UMMA descriptors address directly.
//   * One elected thread issues tcgen05.mma (M=128, N=BN, K=16, bf16 x bf16 -> fp32) into a TMEM accumulator;
//     tcgen05.commit releases smem stages back to the TMA producer through mbarriers.
//   * 8 epilogue warps (2 per TMEM lane quarter) read the accumulator with tcgen05.ld.32x32b.x32, apply bias / activation /
//     residual add / activation-derivative mask (compile-time epilogue variants, EPI_*), stage the bf16 tile in shared memory
//     as 64-channel halves in the 128-byte swizzle and hand each half to the TMA unit (bulk tensor store; "rolling store").
//     The residual / mask operand itself arrives by TMA in the half the result is written to (struct EpiOperand).
//     BatchNorm statistics of the stored output are accumulated from the staging tile (epilogue_stats).
//   Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue.
//   Kernels: conv_tc_kernel (streaming: one A box + one B box per (tap, 64-channel block)), conv_tc2_kernel (the same as a
//   cta_group::2 CTA pair, half the weight tile per CTA), conv_tc_ws_kernel (weight-stationary: the CTA's weight slice resident,
//   one halo box per 64-channel block serves all taps through shifted descriptors).
//
//   wgrad:  D[M = in-channels][N = out-channels] += X^T[M][K] * G^T[N][K]^T , K = pixels.  Both operands are
//   "MN-major" for UMMA (the contraction index is the slow one in NHWC memory), loaded by the same shifted
//   TMA boxes; one TMEM accumulator per filter tap, fp32 red.global.add into gw at the end.
#include "common.cuh"
#include <cstdlib>
#include <cuda.h>
#include <mutex>
#include <string.h>
#include <type_traits>

namespace cgb {

// ------------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
// TMA store (shared::cta -> global, bulk async group), used by the copy-out of the streaming kernel's epilogue
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// elect.sync: one lane of a converged warp.  ptxas knows the guarded region is single-threaded, keeps the MMA
// operands in uniform registers and emits back-to-back UTCHMMA (measured: 44 cycles per M=128,N=48,K=16 MMA = the
// shared-memory operand-read bound (4096+32N)/128, vs 73 with a lane==0 predicate and ~130 with a lane==0 branch).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// Warp-uniform issue: every lane of the MMA warp computes the (uniform) descriptors, one lane (`elected`) issues.
// Keeping the operands provably warp-uniform lets ptxas hold them in uniform registers; branching on lane==0 around
// the address arithmetic instead makes it emit a VOTEU/BRA.U.ANY uniformisation loop per MMA (~130 cycles each).
__device__ __forceinline__ void umma_bf16_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate, uint32_t elected) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "setp.ne.b32 q, %5, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(elected)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar, uint32_t elected) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.b32 q, %1, 0;\n"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(bar), "r"(elected)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- 2-CTA (cta_group::2) forms: a CTA pair of one cluster shares one 256 x N accumulator tile ------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory offset in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads whose completion bytes are credited to a barrier of the LEADER CTA (mbar: a shared::cluster address)
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t mbar_cluster, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(mbar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t mbar_cluster, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(mbar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[256 x N] (+)= A[256 x 16] * B[N x 16]^T : rows 0..127 from the leader's smem / TMEM, rows 128..255 from the peer's;
// each CTA's smem holds N/2 rows of B
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair once the MMAs issued so far have completed
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}

// UMMA shared-memory descriptor, 128B swizzle (layout_type 2), descriptor version 1 (Blackwell).
//   K-major  operand: rows (M/N index) of 128 B, 8-row swizzle atoms 1024 B apart (SBO); LBO unused (1).
//   MN-major operand: rows are K indices of 128 B = 64 contiguous M/N elements; groups of 8 K-rows are SBO=1024 B
//                     apart; the next 64 M/N elements start LBO bytes further (one TMA box).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
// split form for tight issue loops: hi word is loop-invariant, lo word = start-address field (+LBO), so stepping the
// operand is one 32-bit add (no carry out of the 14-bit address field: smem < 256 KB)
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
}
__device__ __forceinline__ uint64_t desc_join(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | (uint64_t)lo; }

// packed-pair helpers for the two 16-bit storage types (bf16: training + inference; fp16: the inference-only --half mode)
template <typename T> struct Pk;
template <> struct Pk<__nv_bfloat16> {
  using T2 = __nv_bfloat162;
  static constexpr bool is_f16 = false;
  static __device__ __forceinline__ float2 to_f2(T2 v) { return __bfloat1622float2(v); }
  static __device__ __forceinline__ T2 zero2() { return __float2bfloat162_rn(0.f); }
};
template <> struct Pk<__half> {
  using T2 = __half2;
  static constexpr bool is_f16 = true;
  static __device__ __forceinline__ float2 to_f2(T2 v) { return __half22float2(v); }
  static __device__ __forceinline__ T2 zero2() { return __float2half2_rn(0.f); }
};

// instruction descriptor: (bf16 x bf16 | f16 x f16) -> fp32, M=128
__host__ __device__ inline uint32_t make_idesc(int n, bool a_mn_major, bool b_mn_major, bool f16 = false, int m = 128) {
  uint32_t d = 0;
  d |= 1u << 4;                          // c_format  F32
  d |= (f16 ? 0u : 1u) << 7;             // a_format  F16 = 0, BF16 = 1
  d |= (f16 ? 0u : 1u) << 10;            // b_format
  d |= (a_mn_major ? 1u : 0u) << 15;     // a_major
  d |= (b_mn_major ? 1u : 0u) << 16;     // b_major
  d |= (uint32_t)(n >> 3) << 17;         // N >> 3
  d |= (uint32_t)(m >> 4) << 24;         // M >> 4 (128: one CTA; 256: a cta_group::2 pair)
  return d;
}

// ------------------------------------------------------------------------------------------------------
// fprop / dgrad kernel
// ------------------------------------------------------------------------------------------------------
struct TcParams {
  int n, hout, wout, cout_s, cin_s;
  int kh, kw, dil, stride, pad_y, pad_x;
  int tw_log, th_log;          // tile = TN x TH x TW pixels, product 128 (powers of two)
  int tiles_x, tiles_y;
  int n_tiles, total_tiles;    // streaming kernel: total = pixel tiles * n_tiles, n-tile fastest
  int pix_tiles;               // weight-stationary kernel: pixel tiles (each CTA keeps one n-tile)
  int bn;                      // N tile (multiple of 16, <= 256)
  int kblocks;                 // ceil(cin_s / 64)
  int stages;
  int tmem_cols;               // >= 2*bn: two accumulator buffers
  int stage_pitch;             // bytes per staging row = bn*2 + 16
  int tma_store;               // streaming kernel: copy-out by TMA store from a 128B-swizzled staging tile (bn % 64 == 0)
  int staging_bufs;            // 1 or 2 staging tiles (2: the store of tile i overlaps the epilogue math of tile i+1)
  int ws_unroll;               // weight-stationary kernel: unrolled 3x3 MMA issue (CGB_WS_UNROLL, default on)
  int aux_tma;                 // the epilogue's second operand (residual / derivative mask) arrives by TMA in the staging tile
  int aux_bar_off;             // byte offset (from the aligned smem base) of its 8 mbarriers [staging buffer][half]
  int staging_tile_bytes;      // bytes of one staging tile: 128*stage_pitch, or ceil(bn/64) swizzled 16 KB halves (TMA store)
  int twh, thh;                // weight-stationary kernel: halo tile extent (pixels)
  int a_stage_bytes;           // weight-stationary kernel: bytes per halo stage (1024-aligned)
  int act;
  float slope;
  int dact;                    // derivative mask (dgrad) from mask_src
  int res_before_act;          // 1: act(acc + bias + residual), 0: act(acc + bias) + residual
  // streaming kernel, generalised taps: tap t reads the input at (oy*stride + tap_dy[t], ox*stride + tap_dx[t]) and
  // uses weight tap tap_w[t]; ntaps <= 64.  Output pixel (oy, ox) of the tile grid is written at
  // (oy*out_stride + out_off_y, ox*out_stride + out_off_x) of an [n, hfull, wfull, cout_s] tensor — this is how the
  // stride-s dgrad runs as s*s stride-1 sub-convolutions, one per output parity class.
  int ntaps;
  int out_stride, out_off_y, out_off_x, hfull, wfull;
  // per-channel sum / sum of squares of the stored output, accumulated by the epilogue (train-mode BatchNorm statistics of a
  // conv -> BN chain without a statistics pass over the tensor): a [2][cout_s] fp32 table in shared memory at byte offset
  // stats_off from the 1024-aligned base, flushed once per CTA to stats_out[blockIdx.x][2][cout_s]
  int stats, stats_off, stats_rows;   // stats_rows: rows of stats_out (CTA 0 zeroes the rows beyond the launch's grid)
  short tap_dy[64], tap_dx[64], tap_w[64];
  // debug (CGB_TC_TRACE=1): CTA 0 writes clock64() stamps [tile < 32][16] here — the per-role timeline of the pipeline
  unsigned long long* trace;
  // magic multipliers ceil(2^32 / d) for the epilogue's per-tile index decomposition (0: divide), see fdiv()
  uint32_t mg_nt, mg_tx, mg_ty, mg_txy;
};

// 2 epilogue warps per TMEM lane quarter.  (Round 1 used 4: its run-time-flag epilogue was a long dependent chain per warp and
// more warps meant more chains in flight.  The compile-time variants are issue-lean and keep two 32-column accumulator loads
// in flight per warp; 16 warps capped the kernel at 96 registers per thread (576 threads), which spilled the BatchNorm partial
// sums and the generic path — 8 warps leave 204.)
constexpr int EPI_WARPS = 8;
constexpr int EPI_PER_Q = EPI_WARPS / 4;
constexpr int TC_THREADS = 64 + 32 * EPI_WARPS;    // warp 0: TMA producer, warp 1: MMA issuer, then the epilogue warps
constexpr int EPI_THREADS = 32 * EPI_WARPS;
static_assert(EPI_PER_Q == 2, "the rolling store maps the 2 warps of a lane quarter onto the two 32-column halves of a 64-channel half");
constexpr int A_TILE_BYTES = 128 * 128;  // 128 pixels x 64 bf16

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); }
// x / d for launch-constant d: one IMAD.HI with the host's magic multiplier (exact while x * d < 2^32, checked on the host); the
// epilogue's 16 warps decompose the tile index once per tile — as real divisions that was ~250 warp instructions per tile and warp
__device__ __forceinline__ uint32_t fdiv(uint32_t x, uint32_t d, uint32_t m) {
  if (d == 1u) return x;
  return m ? __umulhi(x, m) : x / d;
}
static uint32_t magic_for(long long max_x, int d) {
  if (d <= 1 || max_x * (long long)d >= (1ll << 32)) return 0u;
  return (uint32_t)(((1ull << 32) + (unsigned long long)d - 1ull) / (unsigned long long)d);
}
__device__ __forceinline__ void tc_trace(unsigned long long* tr, int lt, int slot) {
  if (tr && blockIdx.x == 0 && lt < 32) tr[lt * 16 + slot] = (unsigned long long)clock64();
}

// ---- epilogue of one 128-pixel x bn tile, executed by the 8 epilogue warps --------------------------------
// phase 1: TMEM -> registers (two tcgen05.ld in flight) -> bias / activation / residual / mask -> bf16 -> staging row
// phase 2: coalesced 16-byte copy-out (consecutive threads write consecutive chunks of a pixel's channel vector)
template <typename T>
__device__ __forceinline__ void epi_chunk(const TcParams& p, const uint32_t* r, int c, int cn0, bool pix_ok,
                                          long long pix, float neg, bool mask_early, const float* __restrict__ bias,
                                          const T* __restrict__ residual,
                                          const T* __restrict__ mask_src, uint8_t* my_row, int sw = -1) {
  // sw < 0: padded row-major staging row (my_row + channel*2).  sw = row & 7: TMA-store staging — per 64-channel half a
  // [128 rows][128 B] tile in the 128-byte swizzle (16-byte chunk index XOR row & 7), my_row = tile base + row*128
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int ch = cn0 + c * 16 + h * 8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[h * 8 + j]);
    if (ch < p.cout_s) {
      if (bias) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + ch));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + ch + 4));
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      }
      if (pix_ok && residual && p.res_before_act) {
        float rr[8];
        Vec8<T>::load(residual + pix * p.cout_s + ch, rr);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += rr[j];
      }
      if (p.act == CGB_ACT_NONE) {
      } else if (p.act <= CGB_ACT_LRELU) {  // relu / lrelu (0 <= slope < 1): max(v, v*neg), 2 instructions
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], v[j] * neg);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = act_apply(v[j], p.act, p.slope);
      }
      if (pix_ok && residual && !p.res_before_act) {
        float rr[8];
        Vec8<T>::load(residual + pix * p.cout_s + ch, rr);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += rr[j];
      }
      if (pix_ok && mask_early) {
        float mm[8];
        Vec8<T>::load(mask_src + pix * p.cout_s + ch, mm);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= act_grad_from_out(mm[j], p.dact, p.slope);
      }
    }
    if (sw < 0) {
      Vec8<T>::store(reinterpret_cast<T*>(my_row + (size_t)(c * 16 + h * 8) * 2), v);
    } else {
      const int j = c * 2 + h;
      Vec8<T>::store(
          reinterpret_cast<T*>(my_row + (size_t)(j >> 3) * (128 * 128) + (size_t)(((j & 7) ^ sw) << 4)), v);
    }
  }
}

// ---- lean phase 1 (TMA-store staging, act in {none, relu, lrelu}) --------------------------------------------------------
// The generic epi_chunk above costs ~230 warp instructions per 16-column chunk (per-element activation switch, 64-bit generic
// addressing, per-8-channel re-tests of launch-uniform flags): with 4 epilogue warps per scheduler the epilogue of a
// 128 x 256 tile took ~4100 cycles against 2048 cycles of MMAs for K = 256 — the short-K convs were bound by the epilogue's
// instruction issue (CGB_TC_TRACE timeline, profiles/r02_epilogue_timeline.txt).  Here every launch-uniform decision is a
// warp-uniform branch per chunk, the operand that needs DRAM latency (residual or derivative mask: never both) is fetched
// before the accumulator wait, and the staging stores use 32-bit shared addresses.  Same arithmetic, same order, same bits.
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
enum { EPI_AUX_NONE = 0, EPI_AUX_RES_BEFORE = 1, EPI_AUX_RES_AFTER = 2, EPI_AUX_MASK = 3 };
// kernel-level epilogue variants (template parameter of the fprop / dgrad kernels)
// GENERIC: TMA-store tile with run-time flags (residual adds, ...); EXOTIC: tanh / sigmoid / selu epilogues (generic chunk code);
// NOTMA: the per-thread copy-out (strided parity-class dgrad, N tiles that are not whole 64-channel halves; the ReLU-masked dgrad too
// when CGB_AUX_TMA=0)
enum { EPI_GENERIC = 0, EPI_PLAIN = 1, EPI_BIAS = 2, EPI_BIAS_ACT = 3, EPI_MASK_RELU = 4, EPI_MASK_LRELU = 5, EPI_EXOTIC = 6, EPI_NOTMA = 7,
       EPI_RES = 8,   // + residual, no bias / activation: the ResNet bottleneck's conv1 dgrad with the skip gradient added in
       EPI_VARIANTS = 9 };
__host__ __device__ constexpr bool epi_may_aux(int epi) { return epi != EPI_PLAIN && epi != EPI_BIAS && epi != EPI_BIAS_ACT; }

template <typename T>
__device__ __forceinline__ void epi_fast_fetch(const TcParams& p, int ch0, bool pix_ok, long long pixoff, const T* __restrict__ aux_src,
                                               uint4* aux) {
  aux[0] = make_uint4(0u, 0u, 0u, 0u);
  aux[1] = make_uint4(0u, 0u, 0u, 0u);
  if (pix_ok) {
    if (ch0 < p.cout_s) aux[0] = __ldg(reinterpret_cast<const uint4*>(aux_src + pixoff + ch0));
    if (ch0 + 8 < p.cout_s) aux[1] = __ldg(reinterpret_cast<const uint4*>(aux_src + pixoff + ch0 + 8));
  }
}

// BIAS / ACT / AUX: 1 / 0 = known at compile time, -1 = decided at run time (the generic instantiation).  AUX adds two
// compile-time mask flavours to the EPI_AUX_* kinds: the launch-uniform branches of the run-time form cost more issue slots than
// the arithmetic (369 warp instructions per pair of chunks, most of them predicated-off bias adds and flag tests).
enum { EPI_AUX_MASK_RELU = 4, EPI_AUX_MASK_LRELU = 5 };
template <typename T, int BIAS, int ACT, int AUX>
__device__ __forceinline__ void epi_fast_apply(const TcParams& p, const uint32_t* r, int ch0, int j0, float neg, int aux_kind_rt,
                                               const float* __restrict__ bias, const uint4* aux, uint32_t row_u32,
                                               uint32_t sw16) {
  const bool has_bias = BIAS < 0 ? (bias != nullptr) : (BIAS != 0);
  const bool has_act = ACT < 0 ? (p.act != CGB_ACT_NONE) : (ACT != 0);
  const int aux_kind = AUX < 0 ? aux_kind_rt : AUX;
#pragma unroll
  for (int h = 0; h < 2; ++h) {   // one 16-byte staging chunk (8 channels) at a time: keeps the live set at 8 + 8 floats
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[8 * h + j]);
    if (has_bias) {
      if (ch0 + 8 * h < p.cout_s) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + ch0 + 8 * h));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + ch0 + 8 * h + 4));
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      }
    }
    float a[8];
    if (aux_kind != EPI_AUX_NONE) {
      const typename Pk<T>::T2* h2 = reinterpret_cast<const typename Pk<T>::T2*>(&aux[h]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = Pk<T>::to_f2(h2[i]);
        a[2 * i] = f.x; a[2 * i + 1] = f.y;
      }
    }
    if (aux_kind == EPI_AUX_RES_BEFORE) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += a[j];
    }
    if (has_act) {   // relu / lrelu (0 <= slope < 1): max(v, v*neg)
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], v[j] * neg);
    }
    if (aux_kind == EPI_AUX_RES_AFTER) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += a[j];
    }
    if (aux_kind == EPI_AUX_MASK_RELU || (aux_kind == EPI_AUX_MASK && p.dact == CGB_ACT_RELU)) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= (a[j] > 0.f ? 1.f : 0.f);
    } else if (aux_kind == EPI_AUX_MASK_LRELU || (aux_kind == EPI_AUX_MASK && p.dact == CGB_ACT_LRELU)) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= (a[j] > 0.f ? 1.f : p.slope);
    } else if (aux_kind == EPI_AUX_MASK) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= act_grad_from_out(a[j], p.dact, p.slope);
    }
    const uint32_t j = (uint32_t)(j0 + h);
    sts128(row_u32 + (j >> 3) * (128u * 128u) + (((j & 7u) << 4) ^ sw16), pack2<T>(v[0], v[1]), pack2<T>(v[2], v[3]),
           pack2<T>(v[4], v[5]), pack2<T>(v[6], v[7]));
  }
}

// ---- BatchNorm statistics from the staging tile -------------------------------------------------------------------
// After phase 1 the tile sits in shared memory exactly as it will be stored (bf16).  Lane = one 16-byte chunk column (8
// channels), warp = a row subset: every LDS.128 of a warp reads 32 consecutive chunks of ONE row (conflict-free in both
// staging layouts), so a thread owns its 8 channels for all its rows and no cross-lane reduction is needed.  The 16 running
// sums live in REGISTERS across the CTA's tiles (a persistent CTA keeps meeting the same N tile: n_tiles divides the grid
// for every BatchNorm'd conv of the path) and are folded into the CTA's shared-memory table only when the channel window
// moves and once at the end.  The table is laid out [moment][channel % 8][channel / 8], so the 32 lanes of a fold hit 32
// consecutive words.  (First version: 16 shared atomics per thread per tile at a stride of 8 words — an 8-way bank conflict
// times 16 warps; it cost 30 us per conv, more than the statistics pass it replaced.)
struct EpiStats {
  float s[8], q[8];
  int ch;   // first channel of the window the sums belong to, -1: none
};

__device__ __forceinline__ void epi_stats_flush(const TcParams& p, EpiStats& st, float* tab) {
  if (st.ch >= 0) {
    const int c8 = p.cout_s >> 3, v = st.ch >> 3;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(tab + j * c8 + v, st.s[j]);
      atomicAdd(tab + p.cout_s + j * c8 + v, st.q[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) st.s[j] = st.q[j] = 0.f;
  st.ch = -1;
}

template <typename T>
__device__ __forceinline__ void epilogue_stats(const TcParams& p, const uint8_t* staging_gen, bool tma, int ox0, int oy0, int n0,
                                               int cn0, float* tab, EpiStats& st, int warp, int lane) {
  const int ew = warp - 2;                 // 0..EPI_WARPS-1
  const int nchunk8 = p.bn >> 3;           // 16-byte chunk columns in the tile (<= 32)
  const int ch = cn0 + lane * 8;
  if (lane >= nchunk8 || ch >= p.cout_s) return;
  if (ch != st.ch) {
    epi_stats_flush(p, st, tab);
    st.ch = ch;
  }
  for (int r = ew; r < 128; r += EPI_WARPS) {
    const int tw2 = r & ((1 << p.tw_log) - 1);
    const int th2 = (r >> p.tw_log) & ((1 << p.th_log) - 1);
    const int tn2 = r >> (p.tw_log + p.th_log);
    if (ox0 + tw2 >= p.wout || oy0 + th2 >= p.hout || n0 + tn2 >= p.n) continue;
    const uint8_t* src = tma ? staging_gen + (size_t)(lane >> 3) * (128 * 128) + (size_t)r * 128 + (size_t)(((lane & 7) ^ (r & 7)) << 4)
                             : staging_gen + (size_t)r * p.stage_pitch + (size_t)lane * 16;
    const uint4 raw = *reinterpret_cast<const uint4*>(src);
    const typename Pk<T>::T2* h = reinterpret_cast<const typename Pk<T>::T2*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = Pk<T>::to_f2(h[i]);
      st.s[2 * i] += f.x; st.s[2 * i + 1] += f.y;
      st.q[2 * i] = fmaf(f.x, f.x, st.q[2 * i]); st.q[2 * i + 1] = fmaf(f.y, f.y, st.q[2 * i + 1]);
    }
  }
}

// ---- the epilogue's second operand (residual / derivative mask) by TMA, in place -------------------------------------------
// Read row-per-thread from global memory (epi_fast_fetch) the operand bounds every masked dgrad of the 640^2 layers: 8 warps x 8
// LDG.128 whose 32 lanes touch 32 different lines, ~2 us of DRAM latency per tile with 16 KB in flight per SM (48->128 gamma||beta
// dgrad: 1.6 ms against 0.6 ms for the same launch without a mask).  With p.aux_tma the operand's tile is loaded by the TMA unit
// INTO THE STAGING HALF THE RESULT WILL BE WRITTEN TO (same tensor map geometry and swizzle as the store, so every thread finds
// the operand of its 16-byte chunk exactly where it is about to write): the half's store leader requests the next tile's operand
// as soon as its bulk store of the current tile has finished reading the half, threads wait on the half's mbarrier, LDS, apply, STS.
// No extra shared memory (the 48->128 launch has none left), no global address arithmetic in the epilogue.
struct EpiOperand {
  int ox0, oy0, n0, cn0;    // the next tile of this CTA (n0 < 0: none)
  uint32_t aux_phase;       // parity bits of the operand barriers, bit = staging buffer * 4 + half
  int primed;               // the operand of the tile about to be processed has been requested
  int sbuf;                 // staging buffer of the current tile
  int ox2, oy2, n2;         // the tile after the next (n2 < 0: none / not tracked): its operand is prefetched into L2
};
__device__ __forceinline__ void tma_prefetch_l2_4d(const CUtensorMap* tm, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0),
               "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// ---- copy-out by the TMA unit, two 64-channel halves at a time ("rolling store") --------------------------------------------
// The 16 warps stage a PAIR of halves (each warp one 16-column chunk of its lane quarter per half), a barrier, the halves'
// leaders (lane 0 of epilogue warp h for half h — bulk groups are per thread) issue their bulk tensor stores, and everybody
// moves on to the next pair while the TMA unit reads this one: the stores overlap the staging of the next pair and, for the
// last pair, the MMAs of the next tile.  Image borders, the batch tail and the channel tail are clipped by the tensor map, so
// no thread computes a global address (unless a residual / derivative mask is read).
// BIAS / ACT / AUX are compile-time (-1: run time): measured with ncu, the run-time form executed ~1000 warp instructions per
// warp and tile (4 warps per scheduler -> ~4400 issue cycles per 128 x 256 tile against 2048 cycles of MMAs at K = 256), i.e.
// the short-K convs were bound by the epilogue's instruction issue (profiles/r02_epilogue_timeline.txt).
template <typename T, int BIAS, int ACT, int AUX, bool GENERIC_CHUNK>
__device__ __forceinline__ void epi_tma_tile(const TcParams& p, uint32_t tmem_acc, uint8_t* staging_gen, uint32_t staging_u32, int ox0,
                                             int oy0, int n0, int cn0, const float* __restrict__ bias,
                                             const T* __restrict__ residual, const T* __restrict__ mask_src, uint32_t tempty_bar,
                                             int warp, int lane, const CUtensorMap* tmY, float* stats_tab, EpiStats* est,
                                             bool tempty_is_cluster_addr, int trace_lt, EpiOperand* pf, const CUtensorMap* tmX,
                                             uint32_t aux_bar0) {
  const int q = warp & 3;              // TMEM lane quarter this warp may access
  const int g = (warp - 2) >> 2;       // 0..1: this warp's 32-column half of each 64-channel half
  const int row = q * 32 + lane;       // tile row == TMEM lane == pixel within the tile
  const int et = threadIdx.x - 64;
  const uint32_t t_row = tmem_acc + ((uint32_t)(q * 32) << 16);
  const int nchunks = p.bn >> 4;       // 16-column chunks in the tile
  const int nh = (p.bn + 63) >> 6;
  const int my_hb = (lane == 0 && (warp - 2) < nh) ? warp - 2 : -1;   // store leader of half my_hb (nh <= 4 <= EPI_WARPS)
  const int aux_kind_rt = residual ? (p.res_before_act ? EPI_AUX_RES_BEFORE : EPI_AUX_RES_AFTER) : (mask_src ? EPI_AUX_MASK : EPI_AUX_NONE);
  const bool any_aux = AUX < 0 ? (aux_kind_rt != EPI_AUX_NONE) : (AUX != EPI_AUX_NONE);
  const T* aux_src = residual ? residual : mask_src;
  const float neg = p.act == CGB_ACT_NONE ? 1.f : (p.act == CGB_ACT_RELU ? 0.f : p.slope);
  const bool aux_tma = !GENERIC_CHUNK && any_aux && p.aux_tma != 0 && pf != nullptr;
  bool pix_ok = false;
  long long pix = 0;
  if ((any_aux && !aux_tma) || GENERIC_CHUNK) {   // only the tiles that read a second tensor from global memory need this thread's pixel address
    const int tw_i = row & ((1 << p.tw_log) - 1);
    const int th_i = (row >> p.tw_log) & ((1 << p.th_log) - 1);
    const int tn_i = row >> (p.tw_log + p.th_log);
    const int ox = ox0 + tw_i, oy = oy0 + th_i, img = n0 + tn_i;
    pix_ok = (ox < p.wout) && (oy < p.hout) && (img < p.n);
    pix = ((long long)img * p.hfull + (oy * p.out_stride + p.out_off_y)) * p.wfull + (ox * p.out_stride + p.out_off_x);
  }
  const long long pixoff = pix * p.cout_s;
  const uint32_t row_u32 = staging_u32 + (uint32_t)row * 128u;
  const uint32_t sw16 = (uint32_t)(row & 7) << 4;
  uint8_t* my_row = staging_gen + (size_t)row * 128;
  const bool mask_early = mask_src != nullptr;
  auto wait_free = [&]() {   // this leader's store that last read the half about to be rewritten has finished reading it
    if (p.staging_bufs == 2) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
  };
  // one 32-column group = chunks c, c+1 (the tile's last group may hold one chunk: bn is a multiple of 16)
  auto process = [&](const uint32_t* r, const uint4* x, int c) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (c + k < nchunks) {
        if (!GENERIC_CHUNK)
          epi_fast_apply<T, BIAS, ACT, AUX>(p, r + 16 * k, cn0 + (c + k) * 16, (c + k) * 2, neg, aux_kind_rt, bias, x + 2 * k, row_u32, sw16);
        else
          epi_chunk(p, r + 16 * k, c + k, cn0, pix_ok, pix, neg, mask_early, bias, residual, mask_src, my_row, row & 7);
      }
    }
  };
  // leader of half h: request the operand tile of (tile coordinates) into staging buffer sb (host: every half of an aux_tma launch
  // starts inside the tensor; the tail channels of the last half are zero-filled by the tensor map)
  auto aux_issue = [&](int sb, uint32_t stg_u32, int h, int c_, int x_, int y_, int n_) {
    const uint32_t bar = aux_bar0 + 8u * (uint32_t)(sb * 4 + h);
    mbar_expect_tx(bar, 128u * 128u);
    tma_load_4d(stg_u32 + (uint32_t)h * (128u * 128u), tmX, bar, c_ + h * 64, x_, y_, n_);
  };
  if (my_hb >= 0 && my_hb < 2) wait_free();
  if (aux_tma && !pf->primed && my_hb >= 0) aux_issue(pf->sbuf, staging_u32, my_hb, cn0, ox0, oy0, n0);   // the CTA's first tile
  if (aux_tma && pf->n2 >= 0 && my_hb >= 0) tma_prefetch_l2_4d(tmX, cn0 + my_hb * 64, pf->ox2, pf->oy2, pf->n2);   // (same N tile)
  epi_bar_sync();
  if (et == 0) tc_trace(p.trace, trace_lt, 7);
  for (int hp = 0; hp < nh; hp += 2) {
    const int c0 = hp * 4 + g * 2, c1 = c0 + 4;   // first chunk of this warp's 32 columns in half hp / hp + 1
    const bool va = c0 < nchunks, vb = c1 < nchunks;
    uint32_t ra[32], rb[32];
    uint4 xa[4], xb[4];
    if (va) { if (c0 + 1 < nchunks) tmem_ld32(t_row + (uint32_t)(c0 * 16), ra); else tmem_ld16(t_row + (uint32_t)(c0 * 16), ra); }
    if (vb) { if (c1 + 1 < nchunks) tmem_ld32(t_row + (uint32_t)(c1 * 16), rb); else tmem_ld16(t_row + (uint32_t)(c1 * 16), rb); }
    if (!GENERIC_CHUNK && any_aux) {
      if (aux_tma) {   // the operand sits where this thread is about to write: halves hp, hp + 1 of this staging buffer
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          if (hp + hh < nh) {
            const uint32_t bit = (uint32_t)(pf->sbuf * 4 + hp + hh);
            mbar_wait(aux_bar0 + 8u * bit, (pf->aux_phase >> bit) & 1u);
            pf->aux_phase ^= 1u << bit;
          }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t ja = (uint32_t)((c0 + k) * 2 + h), jb = (uint32_t)((c1 + k) * 2 + h);
            if (c0 + k < nchunks) xa[2 * k + h] = lds128(row_u32 + (ja >> 3) * (128u * 128u) + (((ja & 7u) << 4) ^ sw16));
            if (c1 + k < nchunks) xb[2 * k + h] = lds128(row_u32 + (jb >> 3) * (128u * 128u) + (((jb & 7u) << 4) ^ sw16));
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (c0 + k < nchunks) epi_fast_fetch<T>(p, cn0 + (c0 + k) * 16, pix_ok, pixoff, aux_src, xa + 2 * k);
          if (c1 + k < nchunks) epi_fast_fetch<T>(p, cn0 + (c1 + k) * 16, pix_ok, pixoff, aux_src, xb + 2 * k);
        }
      }
    }
    tmem_ld_wait();
    if (hp + 2 >= nh) {   // this warp's last TMEM read of the tile: hand the accumulator buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0 && !tempty_is_cluster_addr) mbar_arrive(tempty_bar);
    }
    if (va) process(ra, xa, c0);
    if (vb) process(rb, xb, c1);
    fence_proxy_async_smem();   // this thread's staging writes -> visible to the async proxy
    if (my_hb >= hp + 2 && my_hb < hp + 4) wait_free();   // the next pair's halves are free
    epi_bar_sync();             // halves hp, hp+1 staged
    if (my_hb >= hp && my_hb < hp + 2) {
      if (cn0 + my_hb * 64 < p.cout_s) tma_store_4d(tmY, staging_u32 + (uint32_t)my_hb * (128u * 128u), cn0 + my_hb * 64, ox0, oy0, n0);
      tma_store_commit();
      if (aux_tma && pf->n0 >= 0) {   // the next tile's operand into the half it will be staged in, once that half has been read
        if (p.staging_bufs == 2) {
          tma_store_wait_read<1>();   // (the store of the tile before this one read the other buffer)
          aux_issue(pf->sbuf ^ 1, pf->sbuf ? staging_u32 - (uint32_t)p.staging_tile_bytes : staging_u32 + (uint32_t)p.staging_tile_bytes,
                    my_hb, pf->cn0, pf->ox0, pf->oy0, pf->n0);
        } else {
          tma_store_wait_read<0>();
          aux_issue(0, staging_u32, my_hb, pf->cn0, pf->ox0, pf->oy0, pf->n0);
        }
      }
    }
  }
  if (aux_tma) pf->primed = pf->n0 >= 0 ? 1 : 0;
  if (et == 0) tc_trace(p.trace, trace_lt, 9);
  // (2-CTA kernel: ONE remote arrive per CTA on the leader's barrier, after the last barrier — every warp's TMEM reads are
  //  done; sixteen ~500-cycle cluster arrives per tile and CTA showed up as a slow-down on the short-K 1x1 convs)
  if (et == 0 && tempty_is_cluster_addr) mbar_arrive_cluster(tempty_bar);
  if (stats_tab) epilogue_stats<T>(p, staging_gen, true, ox0, oy0, n0, cn0, stats_tab, *est, warp, lane);
}

template <typename T, int EPI>
__device__ __forceinline__ void epilogue_tile(const TcParams& p, uint32_t tmem_acc, uint8_t* staging_gen, int ox0, int oy0,
                                              int n0, int cn0, const float* __restrict__ bias,
                                              const T* __restrict__ residual,
                                              const T* __restrict__ mask_src, T* __restrict__ y,
                                              uint32_t tempty_bar, int warp, int lane, const CUtensorMap* tmY = nullptr,
                                              uint32_t staging_u32 = 0, float* stats_tab = nullptr, EpiStats* est = nullptr,
                                              bool tempty_is_cluster_addr = false, int trace_lt = 1 << 20,
                                              EpiOperand* pf = nullptr, const CUtensorMap* tmX = nullptr, uint32_t aux_bar0 = 0) {
  if constexpr (EPI != EPI_NOTMA) {
    // EPI: the launch's epilogue variant, chosen on the host (epi_variant_for) and compiled into the kernel
#define CGB_EPI_TILE(B, A, X, G)                                                                                                   \
  epi_tma_tile<T, B, A, X, G>(p, tmem_acc, staging_gen, staging_u32, ox0, oy0, n0, cn0, bias, residual, mask_src, tempty_bar, warp, \
                              lane, tmY, stats_tab, est, tempty_is_cluster_addr, trace_lt, pf, tmX, aux_bar0)
    if constexpr (EPI == EPI_PLAIN) CGB_EPI_TILE(0, 0, EPI_AUX_NONE, false);              // conv -> BatchNorm (ResNet), plain dgrad
    else if constexpr (EPI == EPI_BIAS) CGB_EPI_TILE(1, 0, EPI_AUX_NONE, false);          // gamma || beta, last layers
    else if constexpr (EPI == EPI_BIAS_ACT) CGB_EPI_TILE(1, 1, EPI_AUX_NONE, false);      // conv + bias + (leaky) ReLU
    else if constexpr (EPI == EPI_MASK_RELU) CGB_EPI_TILE(0, 0, EPI_AUX_MASK_RELU, false);    // dgrad through ReLU (CGB_MASK_TMA=1)
    else if constexpr (EPI == EPI_MASK_LRELU) CGB_EPI_TILE(0, 0, EPI_AUX_MASK_LRELU, false);  // dgrad through LeakyReLU (painter, D)
    else if constexpr (EPI == EPI_RES) CGB_EPI_TILE(0, 0, EPI_AUX_RES_AFTER, false);          // dgrad + skip gradient (ResNet bottleneck)
    else if constexpr (EPI == EPI_EXOTIC) CGB_EPI_TILE(-1, -1, -1, true);                 // tanh / sigmoid / selu: generic chunk code
    else CGB_EPI_TILE(-1, -1, -1, false);                                                 // run-time flags (residual adds, ...)
#undef CGB_EPI_TILE
    return;
  } else {
  const int q = warp & 3;              // TMEM lane quarter this warp may access
  const int half = (warp - 2) >> 2;    // EPI_PER_Q warps share a quarter: 16-column chunks interleaved among them
  const int row = q * 32 + lane;       // tile row == TMEM lane == pixel within the tile
  const int et = threadIdx.x - 64;     // 0..255 within the epilogue group
  const int tw_i = row & ((1 << p.tw_log) - 1);
  const int th_i = (row >> p.tw_log) & ((1 << p.th_log) - 1);
  const int tn_i = row >> (p.tw_log + p.th_log);
  const int ox = ox0 + tw_i, oy = oy0 + th_i, img = n0 + tn_i;
  const bool pix_ok = (ox < p.wout) && (oy < p.hout) && (img < p.n);
  const long long pix = ((long long)img * p.hfull + (oy * p.out_stride + p.out_off_y)) * p.wfull + (ox * p.out_stride + p.out_off_x);
  const bool mask_late = (mask_src != nullptr) && (p.dact == CGB_ACT_RELU) && (tmY == nullptr);  // 0/1 mask at the per-thread copy-out
  const bool mask_early = (mask_src != nullptr) && !mask_late;
  const float neg = p.act == CGB_ACT_NONE ? 1.f : (p.act == CGB_ACT_RELU ? 0.f : p.slope);

  const uint32_t t_row = tmem_acc + ((uint32_t)(q * 32) << 16);
  const int nchunks = p.bn >> 4;
  // ---- per-thread copy-out from a padded row-major staging tile (strided parity-class dgrad, N tiles that are not whole halves) --
  // phase-2 geometry of this thread (fixed for the launch): chunk column c of rows rb0 + i * rstep
  const int chunks_per_row = p.bn >> 3;                 // <= 32
  const int rows_per_iter = 32 / chunks_per_row;        // >= 1
  const int rsub = lane / chunks_per_row;
  const int c = lane - rsub * chunks_per_row;
  const int rstep = EPI_WARPS * rows_per_iter;
  const int rb0 = (warp - 2) * rows_per_iter + rsub;
  auto row_off = [&](int r2, int tx0, int ty0, int tn0, int ch2, bool& ok) -> long long {
    const int tw2 = r2 & ((1 << p.tw_log) - 1);
    const int th2 = (r2 >> p.tw_log) & ((1 << p.th_log) - 1);
    const int tn2 = r2 >> (p.tw_log + p.th_log);
    const int ox2 = tx0 + tw2, oy2 = ty0 + th2, img2 = tn0 + tn2;
    ok = r2 < 128 && rsub < rows_per_iter && ch2 < p.cout_s && ox2 < p.wout && oy2 < p.hout && img2 < p.n;
    return (((long long)img2 * p.hfull + (oy2 * p.out_stride + p.out_off_y)) * p.wfull + (ox2 * p.out_stride + p.out_off_x)) * p.cout_s + ch2;
  };
  epi_bar_sync();  // previous tile's copy-out has finished reading the staging buffer
  if (et == 0) tc_trace(p.trace, trace_lt, 7);
  uint8_t* my_row = staging_gen + (size_t)row * p.stage_pitch;
  // lean phase 1 for the common case of this path — the ReLU-masked dgrad (no bias, no activation, no residual; the 0/1 mask is
  // applied by phase 2): 32-column accumulator loads, pack, 16-byte staging stores with 32-bit shared addresses
  const bool lean = !bias && p.act == CGB_ACT_NONE && !residual && !mask_early;
  if (lean) {
    const uint32_t row_u32 = staging_u32 + (uint32_t)row * (uint32_t)p.stage_pitch;
    const int ngroups = (nchunks + 1) >> 1;   // 32-column groups (the last may hold one 16-column chunk)
    for (int ga = half; ga < ngroups; ga += 2 * EPI_PER_Q) {
      const int gb = ga + EPI_PER_Q;
      uint32_t ra[32], rb[32];
      const bool fa = ga * 2 + 1 < nchunks, vb = gb < ngroups, fb = vb && (gb * 2 + 1 < nchunks);
      if (fa) tmem_ld32(t_row + (uint32_t)(ga * 32), ra); else tmem_ld16(t_row + (uint32_t)(ga * 32), ra);
      if (vb) { if (fb) tmem_ld32(t_row + (uint32_t)(gb * 32), rb); else tmem_ld16(t_row + (uint32_t)(gb * 32), rb); }
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < 2 || fa)
          sts128(row_u32 + (uint32_t)(ga * 64 + j * 16), pack2<T>(__uint_as_float(ra[8 * j]), __uint_as_float(ra[8 * j + 1])),
                 pack2<T>(__uint_as_float(ra[8 * j + 2]), __uint_as_float(ra[8 * j + 3])),
                 pack2<T>(__uint_as_float(ra[8 * j + 4]), __uint_as_float(ra[8 * j + 5])),
                 pack2<T>(__uint_as_float(ra[8 * j + 6]), __uint_as_float(ra[8 * j + 7])));
      if (vb) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < 2 || fb)
            sts128(row_u32 + (uint32_t)(gb * 64 + j * 16), pack2<T>(__uint_as_float(rb[8 * j]), __uint_as_float(rb[8 * j + 1])),
                   pack2<T>(__uint_as_float(rb[8 * j + 2]), __uint_as_float(rb[8 * j + 3])),
                   pack2<T>(__uint_as_float(rb[8 * j + 4]), __uint_as_float(rb[8 * j + 5])),
                   pack2<T>(__uint_as_float(rb[8 * j + 6]), __uint_as_float(rb[8 * j + 7])));
      }
    }
  } else
  for (int c = half; c < nchunks; c += 2 * EPI_PER_Q) {
    uint32_t ra[16], rb[16];
    const bool two = (c + EPI_PER_Q) < nchunks;
    tmem_ld16(t_row + (uint32_t)(c * 16), ra);
    if (two) tmem_ld16(t_row + (uint32_t)((c + EPI_PER_Q) * 16), rb);
    tmem_ld_wait();
    epi_chunk(p, ra, c, cn0, pix_ok, pix, neg, mask_early, bias, residual, mask_src, my_row, -1);
    if (two) epi_chunk(p, rb, c + EPI_PER_Q, cn0, pix_ok, pix, neg, mask_early, bias, residual, mask_src, my_row, -1);
  }
  // accumulator buffer drained: hand it back to the MMA warp
  tc_fence_before();
  __syncwarp();
  if (lane == 0 && !tempty_is_cluster_addr) mbar_arrive(tempty_bar);
  if (et == 0) tc_trace(p.trace, trace_lt, 8);
  epi_bar_sync();  // staging complete
  if (et == 0) tc_trace(p.trace, trace_lt, 9);
  if (et == 0 && tempty_is_cluster_addr) mbar_arrive_cluster(tempty_bar);
  if (stats_tab) epilogue_stats<T>(p, staging_gen, false, ox0, oy0, n0, cn0, stats_tab, *est, warp, lane);
  // phase 2: lanes cover (rows_per_iter x chunks_per_row) 16-byte chunks; the row/chunk split of a lane is fixed, so
  // the only per-iteration work is the pixel address.  Consecutive lanes write consecutive chunks of a pixel and then
  // the next pixel: full 32-byte sectors, no read-modify-write.
  const int ch = cn0 + c * 8;
  // four rows per batch: the four mask loads (DRAM / L2 latency) are in flight together — one load -> multiply -> store chain
  // per row left ~8 serial global round trips per thread and tile (the 128->48 gamma||beta dgrad at 640^2: 1.2 ms -> see DESIGN)
  constexpr int PB = 4;
#pragma unroll 1
  for (int r0 = rb0; r0 < 128; r0 += PB * rstep) {
    long long off[PB];
    bool ok[PB];
    uint4 val[PB], mk[PB];
#pragma unroll
    for (int i = 0; i < PB; ++i) {
      const int r2 = r0 + i * rstep;
      off[i] = row_off(r2, ox0, oy0, n0, ch, ok[i]);
      if (ok[i]) {
        val[i] = *reinterpret_cast<const uint4*>(staging_gen + (size_t)r2 * p.stage_pitch + (size_t)c * 16);
        if (mask_late) mk[i] = __ldg(reinterpret_cast<const uint4*>(mask_src + off[i]));
      }
    }
#pragma unroll
    for (int i = 0; i < PB; ++i) {
      if (!ok[i]) continue;
      if (mask_late) {  // relu derivative: keep where the forward output was > 0 (packed bf16x2 compare + multiply)
        using T2 = typename Pk<T>::T2;
        const T2* mh = reinterpret_cast<const T2*>(&mk[i]);
        T2* vh = reinterpret_cast<T2*>(&val[i]);
        const T2 zero2 = Pk<T>::zero2();
#pragma unroll
        for (int j = 0; j < 4; ++j) vh[j] = __hmul2(vh[j], __hgt2(mh[j], zero2));
      }
      *reinterpret_cast<uint4*>(y + off[i]) = val[i];
    }
  }
  }   // EPI_NOTMA
}

// ------------------------------------------------------------------------------------------------------
// Streaming kernel (any geometry): per (tap, 64-channel block) one A box + one B box through a smem ring.
// Persistent: grid = min(#tiles, #SMs); CTA c processes tiles c, c+grid, ...
// smem: [stages x (A 16 KB | B bn*128 B)] [staging 128 x (bn*2+16) B] [barriers]
// ------------------------------------------------------------------------------------------------------
template <typename T, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX, const TcParams p, const float* __restrict__ bias, const T* __restrict__ residual,
               const T* __restrict__ mask_src, T* __restrict__ y, float* __restrict__ stats_out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // 1024-B alignment for the 128B swizzle atoms
  const uint32_t b_tile_bytes = (uint32_t)p.bn * 128u;
  const uint32_t stage_bytes = A_TILE_BYTES + b_tile_bytes;
  const uint32_t staging = base + (uint32_t)p.stages * stage_bytes;
  const uint32_t staging_tile = (uint32_t)p.staging_tile_bytes;
  const uint32_t bar_base = staging + staging_tile * (uint32_t)p.staging_bufs;
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(p.stages + s); };
  auto tfull_bar = [&](int b) { return bar_base + 16u * (uint32_t)p.stages + 8u * (uint32_t)b; };
  auto tempty_bar = [&](int b) { return bar_base + 16u * (uint32_t)p.stages + 16u + 8u * (uint32_t)b; };
  const uint32_t tmem_ptr_addr = bar_base + 16u * (uint32_t)p.stages + 32u;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw));
  uint8_t* staging_gen = smem_raw + (staging - raw);
  float* stats_tab = p.stats ? reinterpret_cast<float*>(smem_raw + (base - raw) + (uint32_t)p.stats_off) : nullptr;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int taps = p.ntaps;
  const int iters = taps * p.kblocks;
  const int tn_log = 7 - p.tw_log - p.th_log;
  if (stats_tab)
    for (int i = threadIdx.x; i < 2 * p.cout_s; i += TC_THREADS) stats_tab[i] = 0.f;   // visible after the __syncthreads below

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (p.tma_store) prefetch_tmap(&tmY);
    if (p.aux_tma) {
      prefetch_tmap(&tmX);
      for (int i = 0; i < 8; ++i) mbar_init(base + (uint32_t)p.aux_bar_off + 8u * (uint32_t)i, 1);
    }
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), EPI_WARPS);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_addr, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one_sync()) {
      int s = 0;
      uint32_t ph = 0;
      int lt = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
        const int nt = tile % p.n_tiles;
        const int pt = tile / p.n_tiles;
        const int tx = pt % p.tiles_x;
        const int ty = (pt / p.tiles_x) % p.tiles_y;
        const int tn = pt / (p.tiles_x * p.tiles_y);
        const int ox0 = tx << p.tw_log, oy0 = ty << p.th_log, n0 = tn << tn_log;
        const int cn0 = nt * p.bn;
        tc_trace(p.trace, lt, 0);
        for (int tap = 0; tap < taps; ++tap) {
          const int cx = ox0 * p.stride + p.tap_dx[tap];
          const int cy = oy0 * p.stride + p.tap_dy[tap];
          const int wtap = p.tap_w[tap];
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(empty_bar(s), ph ^ 1u);
            const uint32_t a_dst = base + (uint32_t)s * stage_bytes;
            mbar_expect_tx(full_bar(s), stage_bytes);
            tma_load_4d(a_dst, &tmA, full_bar(s), kb * 64, cx, cy, n0);
            tma_load_3d(a_dst + A_TILE_BYTES, &tmB, full_bar(s), kb * 64, wtap, cn0);
            if (++s == p.stages) { s = 0; ph ^= 1u; }
          }
        }
        tc_trace(p.trace, lt, 1);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = make_idesc(p.bn, false, false, Pk<T>::is_f16);
    const uint32_t hi1024 = desc_hi(1024u);
    const int ksteps_last = ((p.cin_s - (p.kblocks - 1) * 64) + 15) >> 4;
    int s = 0;
    uint32_t ph = 0;
    int lt = 0;  // local tile counter
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
      const int buf = lt & 1;
      const uint32_t bph = (uint32_t)(lt >> 1) & 1u;
      mbar_wait(tempty_bar(buf), bph ^ 1u);  // epilogue has drained this accumulator buffer
      tc_fence_after();
      if (lane == 0) tc_trace(p.trace, lt, 2);
      const uint32_t d_addr = tmem_base + (uint32_t)(buf * p.bn);
      int kb = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        if (lane == 0 && it == 0) tc_trace(p.trace, lt, 3);
        if (lane == 0 && it == iters - 1) tc_trace(p.trace, lt, 4);
        if (elect_one_sync()) {
          const int ksteps = (kb == p.kblocks - 1) ? ksteps_last : 4;
          const uint32_t a_addr = base + (uint32_t)s * stage_bytes;
          const uint32_t a_lo = desc_lo(a_addr, 16u), b_lo = desc_lo(a_addr + A_TILE_BYTES, 16u);
          uint32_t acc = it > 0 ? 1u : 0u;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (k < ksteps) {
              umma_bf16(d_addr, desc_join(a_lo + 2u * k, hi1024), desc_join(b_lo + 2u * k, hi1024), idesc, acc);
              acc = 1u;
            }
          }
          umma_commit(empty_bar(s));  // frees this smem stage once the MMAs above have read it
          if (it == iters - 1) umma_commit(tfull_bar(buf));
        }
        __syncwarp();
        if (++kb == p.kblocks) kb = 0;
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // ================= epilogue warps =================
    EpiStats est;
    est.ch = -1;
    epi_stats_flush(p, est, stats_tab);   // (zeroes the registers)
    EpiOperand pf;
    pf.n0 = -1;
    pf.aux_phase = 0u;
    pf.primed = 0;
    pf.sbuf = 0;
    pf.n2 = -1;
    int lt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
      const int buf = lt & 1;
      const uint32_t bph = (uint32_t)(lt >> 1) & 1u;
      const int pt = (int)fdiv((uint32_t)tile, (uint32_t)p.n_tiles, p.mg_nt);
      const int nt = tile - pt * p.n_tiles;
      const int py = (int)fdiv((uint32_t)pt, (uint32_t)p.tiles_x, p.mg_tx);
      const int tx = pt - py * p.tiles_x;
      const int tn = (int)fdiv((uint32_t)py, (uint32_t)p.tiles_y, p.mg_ty);
      const int ty = py - tn * p.tiles_y;
      if constexpr (epi_may_aux(EPI)) {   // the next tile of this CTA: its residual / mask operand is requested a tile ahead (EpiOperand)
        const int tile2 = tile + (int)gridDim.x;
        pf.n0 = -1;
        pf.sbuf = (p.staging_bufs == 2) ? (lt & 1) : 0;
        if ((mask_src || residual) && tile2 < p.total_tiles) {
          const int pt2 = (int)fdiv((uint32_t)tile2, (uint32_t)p.n_tiles, p.mg_nt);
          const int py2 = (int)fdiv((uint32_t)pt2, (uint32_t)p.tiles_x, p.mg_tx);
          const int tn2 = (int)fdiv((uint32_t)py2, (uint32_t)p.tiles_y, p.mg_ty);
          pf.cn0 = (tile2 - pt2 * p.n_tiles) * p.bn;
          pf.ox0 = (pt2 - py2 * p.tiles_x) << p.tw_log;
          pf.oy0 = (py2 - tn2 * p.tiles_y) << p.th_log;
          pf.n0 = tn2 << tn_log;
        }
      }
      if (threadIdx.x == 64) tc_trace(p.trace, lt, 5);
      mbar_wait(tfull_bar(buf), bph);
      tc_fence_after();
      if (threadIdx.x == 64) tc_trace(p.trace, lt, 6);
      const uint32_t sb = (p.staging_bufs == 2) ? (uint32_t)(lt & 1) * staging_tile : 0u;
      epilogue_tile<T, EPI>(p, tmem_base + (uint32_t)(buf * p.bn), staging_gen + sb, tx << p.tw_log, ty << p.th_log, tn << tn_log,
                    nt * p.bn, bias, residual, mask_src, y, tempty_bar(buf), warp, lane, p.tma_store ? &tmY : nullptr,
                    staging + sb, stats_tab, &est, false, lt, epi_may_aux(EPI) ? &pf : nullptr, &tmX, base + (uint32_t)p.aux_bar_off);
      if (threadIdx.x == 64) tc_trace(p.trace, lt, 10);
    }
    if (p.tma_store && lane == 0 && warp < 6) tma_store_wait_all();   // every store leader: its bulk stores have completed   // the issuing thread: every bulk store has completed
    if (stats_tab) {   // fold the registers, then write the CTA's table: one plain store per entry (a CTA without tiles writes zeros)
      epi_stats_flush(p, est, stats_tab);
      epi_bar_sync();
      float* out = stats_out + (size_t)blockIdx.x * 2 * p.cout_s;
      const int c8 = p.cout_s >> 3;
      for (int i = threadIdx.x - 64; i < 2 * p.cout_s; i += EPI_THREADS) {
        const int m = i >= p.cout_s ? 1 : 0, c = i - m * p.cout_s;
        out[i] = stats_tab[m * p.cout_s + (c & 7) * c8 + (c >> 3)];
      }
      if (blockIdx.x == 0) {   // rows no CTA owns (a launch smaller than the partial buffer) must read as zero
        const size_t lo = (size_t)gridDim.x * 2 * p.cout_s, hi = (size_t)p.stats_rows * 2 * p.cout_s;
        for (size_t i = lo + (threadIdx.x - 64); i < hi; i += EPI_THREADS) stats_out[i] = 0.f;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------
// Streaming kernel, CTA-PAIR form (tcgen05 cta_group::2): the two CTAs of a cluster own ONE 256-pixel x bn accumulator tile —
// CTA r the pixel tile 2q + r (rows r*128.. of the accumulator, in its own TMEM) — and each stages only HALF of the weight
// tile (bn/2 rows) per (tap, 64-channel block): the tensor cores of both SMs read both halves.  Per 256 output pixels the pair
// pulls 2 A tiles + ONE B tile through L2 instead of 2 + 2: the streaming kernel's 1x1 and dilated ResNet convs are bound by
// exactly that L2 -> SM traffic, and a stage shrinks from 16 KB + bn*128 B to 16 KB + bn*64 B (deeper
// ring).  Protocol (barriers live at the same offsets in both CTAs):
//   full[s]   (leader's is used): the leader's producer arms it with the bytes of BOTH CTAs' loads; every TMA load of the pair
//             (cp.async.bulk.tensor ... cta_group::2) credits its bytes to the leader's barrier;
//   empty[s]  (both): the leader's tcgen05.commit multicasts one arrive to each CTA when the MMAs that read stage s are done;
//   tfull[b]  (both): same multicast commit after the last MMA of a tile — each CTA's epilogue drains its own TMEM half;
//   tempty[b] (leader's is used): every epilogue warp of BOTH CTAs arrives on it (the peer's through mapa'd cluster addresses).
// The MMA is issued by the leader's warp 1 only; the peer's warp 1 just takes part in the paired TMEM allocation.
template <typename T, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX, const TcParams p, const float* __restrict__ bias, const T* __restrict__ residual,
                const T* __restrict__ mask_src, T* __restrict__ y, float* __restrict__ stats_out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t b_half_bytes = (uint32_t)p.bn * 64u;               // bn/2 rows of 128 B
  const uint32_t stage_bytes = A_TILE_BYTES + b_half_bytes;
  const uint32_t staging = base + (uint32_t)p.stages * stage_bytes;
  const uint32_t staging_tile = (uint32_t)p.staging_tile_bytes;
  const uint32_t bar_base = staging + staging_tile * (uint32_t)p.staging_bufs;
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(p.stages + s); };
  auto tfull_bar = [&](int b) { return bar_base + 16u * (uint32_t)p.stages + 8u * (uint32_t)b; };
  auto tempty_bar = [&](int b) { return bar_base + 16u * (uint32_t)p.stages + 16u + 8u * (uint32_t)b; };
  const uint32_t tmem_ptr_addr = bar_base + 16u * (uint32_t)p.stages + 32u;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw));
  uint8_t* staging_gen = smem_raw + (staging - raw);
  float* stats_tab = p.stats ? reinterpret_cast<float*>(smem_raw + (base - raw) + (uint32_t)p.stats_off) : nullptr;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int taps = p.ntaps;
  const int iters = taps * p.kblocks;
  const int tn_log = 7 - p.tw_log - p.th_log;
  const int pix_tiles = p.total_tiles / p.n_tiles;                 // pixel tiles; the pair walks them two at a time
  const int pair_tiles = ((pix_tiles + 1) >> 1) * p.n_tiles;       // (pixel-tile pair, n tile), n tile fastest
  if (stats_tab)
    for (int i = threadIdx.x; i < 2 * p.cout_s; i += TC_THREADS) stats_tab[i] = 0.f;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (p.tma_store) prefetch_tmap(&tmY);
    if (p.aux_tma) {
      prefetch_tmap(&tmX);
      for (int i = 0; i < 8; ++i) mbar_init(base + (uint32_t)p.aux_bar_off + 8u * (uint32_t)i, 1);
    }
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 2);               // one arrive per CTA of the pair (thread 0 of its epilogue group)
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(tmem_ptr_addr, (uint32_t)p.tmem_cols);
  tc_fence_before();
  cluster_sync_all();            // barrier inits and the allocation of BOTH CTAs are visible before any remote arrive / load
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  // decode pair tile -> this CTA's pixel tile (out of range: n0 >= n, every load zero-filled, every store clipped)
  auto decode = [&](int pt_tile, int& ox0, int& oy0, int& n0, int& cn0) {
    const int pq = (int)fdiv((uint32_t)pt_tile, (uint32_t)p.n_tiles, p.mg_nt);
    const int nt = pt_tile - pq * p.n_tiles;
    const int pt = pq * 2 + (int)rank;
    const int py = (int)fdiv((uint32_t)pt, (uint32_t)p.tiles_x, p.mg_tx);
    const int tx = pt - py * p.tiles_x;
    const int tn = (int)fdiv((uint32_t)py, (uint32_t)p.tiles_y, p.mg_ty);
    const int ty = py - tn * p.tiles_y;
    ox0 = tx << p.tw_log; oy0 = ty << p.th_log; n0 = tn << tn_log; cn0 = nt * p.bn;
  };

  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    if (elect_one_sync()) {
      int s = 0;
      uint32_t ph = 0;
      for (int t2 = pair; t2 < pair_tiles; t2 += npairs) {
        int ox0, oy0, n0, cn0;
        decode(t2, ox0, oy0, n0, cn0);
        for (int tap = 0; tap < taps; ++tap) {
          const int cx = ox0 * p.stride + p.tap_dx[tap];
          const int cy = oy0 * p.stride + p.tap_dy[tap];
          const int wtap = p.tap_w[tap];
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(empty_bar(s), ph ^ 1u);
            const uint32_t a_dst = base + (uint32_t)s * stage_bytes;
            const uint32_t lead_full = mapa_rank(full_bar(s), 0);
            if (leader) mbar_expect_tx(full_bar(s), 2u * stage_bytes);
            tma2_load_4d(a_dst, &tmA, lead_full, kb * 64, cx, cy, n0);
            tma2_load_3d(a_dst + A_TILE_BYTES, &tmB, lead_full, kb * 64, wtap, cn0 + (int)rank * (p.bn >> 1));
            if (++s == p.stages) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader only) =================
    if (leader) {
      const uint32_t idesc = make_idesc(p.bn, false, false, Pk<T>::is_f16, 256);
      const uint32_t hi1024 = desc_hi(1024u);
      const int ksteps_last = ((p.cin_s - (p.kblocks - 1) * 64) + 15) >> 4;
      int s = 0;
      uint32_t ph = 0;
      int lt = 0;
      for (int t2 = pair; t2 < pair_tiles; t2 += npairs, ++lt) {
        const int buf = lt & 1;
        const uint32_t bph = (uint32_t)(lt >> 1) & 1u;
        mbar_wait(tempty_bar(buf), bph ^ 1u);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + (uint32_t)(buf * p.bn);
        int kb = 0;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          if (elect_one_sync()) {
            const int ksteps = (kb == p.kblocks - 1) ? ksteps_last : 4;
            const uint32_t a_addr = base + (uint32_t)s * stage_bytes;
            const uint32_t a_lo = desc_lo(a_addr, 16u), b_lo = desc_lo(a_addr + A_TILE_BYTES, 16u);
            uint32_t acc = it > 0 ? 1u : 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (k < ksteps) {
                umma2_bf16(d_addr, desc_join(a_lo + 2u * k, hi1024), desc_join(b_lo + 2u * k, hi1024), idesc, acc);
                acc = 1u;
              }
            }
            umma2_commit_mc(empty_bar(s));
            if (it == iters - 1) umma2_commit_mc(tfull_bar(buf));
          }
          __syncwarp();
          if (++kb == p.kblocks) kb = 0;
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else {
    // ================= epilogue warps (both CTAs, each on its own 128 rows) =================
    EpiStats est;
    est.ch = -1;
    epi_stats_flush(p, est, stats_tab);
    EpiOperand pf;
    pf.n0 = -1;
    pf.aux_phase = 0u;
    pf.primed = 0;
    pf.sbuf = 0;
    pf.n2 = -1;
    int lt = 0;
    for (int t2 = pair; t2 < pair_tiles; t2 += npairs, ++lt) {
      const int buf = lt & 1;
      const uint32_t bph = (uint32_t)(lt >> 1) & 1u;
      int ox0, oy0, n0, cn0;
      decode(t2, ox0, oy0, n0, cn0);
      if constexpr (epi_may_aux(EPI)) {
        pf.n0 = -1;
        pf.sbuf = (p.staging_bufs == 2) ? (lt & 1) : 0;
        if ((mask_src || residual) && t2 + npairs < pair_tiles) decode(t2 + npairs, pf.ox0, pf.oy0, pf.n0, pf.cn0);
      }
      mbar_wait(tfull_bar(buf), bph);
      tc_fence_after();
      const uint32_t sb = (p.staging_bufs == 2) ? (uint32_t)(lt & 1) * staging_tile : 0u;
      epilogue_tile<T, EPI>(p, tmem_base + (uint32_t)(buf * p.bn), staging_gen + sb, ox0, oy0, n0, cn0, bias, residual, mask_src, y,
                       mapa_rank(tempty_bar(buf), 0), warp, lane, p.tma_store ? &tmY : nullptr, staging + sb, stats_tab, &est, true,
                       1 << 20, epi_may_aux(EPI) ? &pf : nullptr, &tmX, base + (uint32_t)p.aux_bar_off);
    }
    if (p.tma_store && lane == 0 && warp < 6) tma_store_wait_all();   // every store leader: its bulk stores have completed
    if (stats_tab) {
      epi_stats_flush(p, est, stats_tab);
      epi_bar_sync();
      float* out = stats_out + (size_t)blockIdx.x * 2 * p.cout_s;
      const int c8 = p.cout_s >> 3;
      for (int i = threadIdx.x - 64; i < 2 * p.cout_s; i += EPI_THREADS) {
        const int m = i >= p.cout_s ? 1 : 0, c = i - m * p.cout_s;
        out[i] = stats_tab[m * p.cout_s + (c & 7) * c8 + (c >> 3)];
      }
      if (blockIdx.x == 0) {
        const size_t lo = (size_t)gridDim.x * 2 * p.cout_s, hi = (size_t)p.stats_rows * 2 * p.cout_s;
        for (size_t i = lo + (threadIdx.x - 64); i < hi; i += EPI_THREADS) stats_out[i] = 0.f;
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();            // neither CTA frees TMEM / exits while the other may still signal its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------
// Weight-stationary halo kernel (stride-1 k x k convs on large maps — the SPADE gamma/beta convs and their dgrad).
// L2->SM bandwidth (~6.6 TB/s, about the HBM rate on this part) is what bounds the streaming kernel: it re-reads
// the activation tile once per filter tap and the weights once per pixel tile.  Here
//   * the CTA's weight slice [kblocks][taps][bn][64] is loaded ONCE into shared memory and stays resident,
//   * per 64-channel block ONE halo box (TH+dil*(kh-1)) x (TW+dil*(kw-1)) pixels is loaded, and every filter tap is
//     an UMMA descriptor into that same box: start address shifted by (dy*dil*TWh + dx*dil) rows of 128 B, 8-row
//     group stride SBO = TWh*128 B (the 128B swizzle is a function of the smem address, so row-shifted starts
//     read correctly — verified on hardware, scripts/exp/shift_desc.cu).
// Tile = 16 rows x 8 pixels (each 8-row swizzle group is one image-row segment).
// smem: [weights] [stages x halo tile] [staging] [barriers].  grid is a multiple of n_tiles; CTA c keeps n-tile c % n_tiles.
// ------------------------------------------------------------------------------------------------------
template <typename T, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_ws_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX, const TcParams p, const float* __restrict__ bias, const T* __restrict__ residual,
               const T* __restrict__ mask_src, T* __restrict__ y, float* __restrict__ stats_out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int taps = p.kh * p.kw;
  const uint32_t w_tap_bytes = (uint32_t)p.bn * 128u;
  const uint32_t w_bytes = (uint32_t)(p.kblocks * taps) * w_tap_bytes;
  const uint32_t a_base = base + w_bytes;
  const uint32_t staging = a_base + (uint32_t)p.stages * (uint32_t)p.a_stage_bytes;
  const uint32_t bar_base = staging + (uint32_t)p.staging_tile_bytes * (uint32_t)p.staging_bufs;
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(p.stages + s); };
  auto tfull_bar = [&](int b) { return bar_base + 16u * (uint32_t)p.stages + 8u * (uint32_t)b; };
  auto tempty_bar = [&](int b) { return bar_base + 16u * (uint32_t)p.stages + 16u + 8u * (uint32_t)b; };
  const uint32_t w_bar = bar_base + 16u * (uint32_t)p.stages + 32u;
  const uint32_t tmem_ptr_addr = w_bar + 8u;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw));
  uint8_t* staging_gen = smem_raw + (staging - raw);
  float* stats_tab = p.stats ? reinterpret_cast<float*>(smem_raw + (base - raw) + (uint32_t)p.stats_off) : nullptr;
  if (stats_tab)
    for (int i = threadIdx.x; i < 2 * p.cout_s; i += TC_THREADS) stats_tab[i] = 0.f;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nt = blockIdx.x % p.n_tiles;
  const int cn0 = nt * p.bn;
  const int pt0 = blockIdx.x / p.n_tiles;
  const int pt_step = gridDim.x / p.n_tiles;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (p.tma_store) prefetch_tmap(&tmY);
    if (p.aux_tma) {
      prefetch_tmap(&tmX);
      for (int i = 0; i < 8; ++i) mbar_init(base + (uint32_t)p.aux_bar_off + 8u * (uint32_t)i, 1);
    }
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), EPI_WARPS);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_addr, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    if (elect_one_sync()) {
      // resident weights: kblocks*taps boxes of {64 ch, 1 tap, bn rows}
      mbar_expect_tx(w_bar, w_bytes);
      for (int kb = 0; kb < p.kblocks; ++kb)
        for (int tap = 0; tap < taps; ++tap)
          tma_load_3d(base + (uint32_t)(kb * taps + tap) * w_tap_bytes, &tmB, w_bar, kb * 64, tap, cn0);
      const uint32_t halo_bytes = (uint32_t)(p.twh * p.thh) * 128u;
      int s = 0;
      uint32_t ph = 0;
      int lt = 0;
      for (int pt = pt0; pt < p.pix_tiles; pt += pt_step, ++lt) {
        const int tx = pt % p.tiles_x;
        const int ty = (pt / p.tiles_x) % p.tiles_y;
        const int img = pt / (p.tiles_x * p.tiles_y);
        const int cx = (tx << 3) - p.pad_x, cy = (ty << 4) - p.pad_y;
        tc_trace(p.trace, lt, 0);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_expect_tx(full_bar(s), halo_bytes);
          tma_load_4d(a_base + (uint32_t)s * (uint32_t)p.a_stage_bytes, &tmA, full_bar(s), kb * 64, cx, cy, img);
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
        tc_trace(p.trace, lt, 1);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(p.bn, false, false, Pk<T>::is_f16);
    const uint32_t hi_a = desc_hi((uint32_t)p.twh * 128u), hi_b = desc_hi(1024u);
    const int ksteps_last = ((p.cin_s - (p.kblocks - 1) * 64) + 15) >> 4;
    mbar_wait(w_bar, 0u);
    int s = 0;
    uint32_t ph = 0;
    int lt = 0;
    for (int pt = pt0; pt < p.pix_tiles; pt += pt_step, ++lt) {
      const int buf = lt & 1;
      const uint32_t bph = (uint32_t)(lt >> 1) & 1u;
      mbar_wait(tempty_bar(buf), bph ^ 1u);
      tc_fence_after();
      if (lane == 0) tc_trace(p.trace, lt, 2);
      const uint32_t d_addr = tmem_base + (uint32_t)(buf * p.bn);
      for (int kb = 0; kb < p.kblocks; ++kb) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        if (lane == 0 && kb == 0) tc_trace(p.trace, lt, 3);
        if (lane == 0 && kb == p.kblocks - 1) tc_trace(p.trace, lt, 4);
        if (elect_one_sync()) {
          const int ksteps = (kb == p.kblocks - 1) ? ksteps_last : 4;
          const uint32_t a_lo0 = desc_lo(a_base + (uint32_t)s * (uint32_t)p.a_stage_bytes, 16u);
          uint32_t b_lo = desc_lo(base + (uint32_t)(kb * taps) * w_tap_bytes, 16u);
          uint32_t acc = kb > 0 ? 1u : 0u;
          if (p.ws_unroll && p.kh == 3 && p.kw == 3) {
            // 3x3 (every user of this kernel on the hot path): taps and K steps unrolled, operands of the form
            // launch-constant + small compile-time multiples.  The generic loop below costs ~19 single-thread instructions per
            // MMA (six R2UR moves among them): measured with CGB_TC_TRACE, the N = 48 gamma||beta MMAs took 74 cycles each and
            // the N = 32 ones 113 against 44 / 40 of operand-read time — the issue loop, not the tensor pipe, was the bound.
            const uint32_t row16 = (uint32_t)(p.dil * p.twh) * 8u, col16 = (uint32_t)p.dil * 8u, wtap16 = w_tap_bytes >> 4;
            auto issue9 = [&](auto KS) {
#pragma unroll
              for (int t = 0; t < 9; ++t) {
                const uint32_t a_lo = a_lo0 + (uint32_t)(t / 3) * row16 + (uint32_t)(t % 3) * col16;
                const uint32_t b_t = b_lo + (uint32_t)t * wtap16;
#pragma unroll
                for (int k = 0; k < decltype(KS)::value; ++k) {
                  umma_bf16(d_addr, desc_join(a_lo + 2u * k, hi_a), desc_join(b_t + 2u * k, hi_b), idesc, acc);
                  acc = 1u;
                }
              }
            };
            if (ksteps == 4) issue9(std::integral_constant<int, 4>{});
            else if (ksteps == 3) issue9(std::integral_constant<int, 3>{});
            else if (ksteps == 2) issue9(std::integral_constant<int, 2>{});
            else issue9(std::integral_constant<int, 1>{});
          } else {
          for (int dy = 0; dy < p.kh; ++dy) {
            uint32_t a_lo = a_lo0 + (uint32_t)(dy * p.dil * p.twh) * 8u;
            for (int dx = 0; dx < p.kw; ++dx) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (k < ksteps) {
                  umma_bf16(d_addr, desc_join(a_lo + 2u * k, hi_a), desc_join(b_lo + 2u * k, hi_b), idesc, acc);
                  acc = 1u;
                }
              }
              a_lo += (uint32_t)p.dil * 8u;
              b_lo += w_tap_bytes >> 4;
            }
          }
          }
          umma_commit(empty_bar(s));
          if (kb == p.kblocks - 1) umma_commit(tfull_bar(buf));
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    EpiStats est;
    est.ch = -1;
    epi_stats_flush(p, est, stats_tab);
    EpiOperand pf;
    pf.n0 = -1;
    pf.aux_phase = 0u;
    pf.primed = 0;
    pf.sbuf = 0;
    pf.n2 = -1;
    int lt = 0;
    for (int pt = pt0; pt < p.pix_tiles; pt += pt_step, ++lt) {
      const int buf = lt & 1;
      const uint32_t bph = (uint32_t)(lt >> 1) & 1u;
      const int py = (int)fdiv((uint32_t)pt, (uint32_t)p.tiles_x, p.mg_tx);
      const int tx = pt - py * p.tiles_x;
      const int img = (int)fdiv((uint32_t)py, (uint32_t)p.tiles_y, p.mg_ty);
      const int ty = py - img * p.tiles_y;
      if (threadIdx.x == 64) tc_trace(p.trace, lt, 5);
      mbar_wait(tfull_bar(buf), bph);
      tc_fence_after();
      if (threadIdx.x == 64) tc_trace(p.trace, lt, 6);
      // two staging tiles when they fit (launch_fprop): the store of tile i overlaps the staging of tile i+1 — with one, every
      // tile of a small-channel conv (24->24 at 640^2: 684 cycles of MMAs) waited ~1500 cycles for the previous bulk store
      const uint32_t sb = (p.staging_bufs == 2) ? (uint32_t)(lt & 1) * (uint32_t)p.staging_tile_bytes : 0u;
      if constexpr (epi_may_aux(EPI)) {
        const int pt2 = pt + pt_step;
        pf.n0 = -1;
        pf.sbuf = (p.staging_bufs == 2) ? (lt & 1) : 0;
        if ((mask_src || residual) && pt2 < p.pix_tiles) {
          const int py2 = (int)fdiv((uint32_t)pt2, (uint32_t)p.tiles_x, p.mg_tx);
          const int img2 = (int)fdiv((uint32_t)py2, (uint32_t)p.tiles_y, p.mg_ty);
          pf.ox0 = (pt2 - py2 * p.tiles_x) << 3;
          pf.oy0 = (py2 - img2 * p.tiles_y) << 4;
          pf.n0 = img2;
          pf.cn0 = cn0;
        }
        const int pt3 = pt2 + pt_step;
        pf.n2 = -1;
        if (p.aux_tma > 1 && pt3 < p.pix_tiles) {   // single staging tile: the operand load is issued late, so warm L2 two tiles ahead
          const int py3 = (int)fdiv((uint32_t)pt3, (uint32_t)p.tiles_x, p.mg_tx);   // (48->128 dgrad at 640^2: 1.215 -> 1.114 ms)
          const int img3 = (int)fdiv((uint32_t)py3, (uint32_t)p.tiles_y, p.mg_ty);
          pf.ox2 = (pt3 - py3 * p.tiles_x) << 3;
          pf.oy2 = (py3 - img3 * p.tiles_y) << 4;
          pf.n2 = img3;
        }
      }
      epilogue_tile<T, EPI>(p, tmem_base + (uint32_t)(buf * p.bn), staging_gen + sb, tx << 3, ty << 4, img, cn0, bias, residual,
                    mask_src, y, tempty_bar(buf), warp, lane, p.tma_store ? &tmY : nullptr, staging + sb, stats_tab, &est, false, lt,
                    epi_may_aux(EPI) ? &pf : nullptr, &tmX, base + (uint32_t)p.aux_bar_off);
      if (threadIdx.x == 64) tc_trace(p.trace, lt, 10);
    }
    if (p.tma_store && lane == 0 && warp < 6) tma_store_wait_all();   // every store leader: its bulk stores have completed
    if (stats_tab) {
      epi_stats_flush(p, est, stats_tab);
      epi_bar_sync();
      float* out = stats_out + (size_t)blockIdx.x * 2 * p.cout_s;
      const int c8 = p.cout_s >> 3;
      for (int i = threadIdx.x - 64; i < 2 * p.cout_s; i += EPI_THREADS) {
        const int m = i >= p.cout_s ? 1 : 0, c = i - m * p.cout_s;
        out[i] = stats_tab[m * p.cout_s + (c & 7) * c8 + (c >> 3)];
      }
      if (blockIdx.x == 0) {   // rows no CTA owns (a launch smaller than the partial buffer) must read as zero
        const size_t lo = (size_t)gridDim.x * 2 * p.cout_s, hi = (size_t)p.stats_rows * 2 * p.cout_s;
        for (size_t i = lo + (threadIdx.x - 64); i < hi; i += EPI_THREADS) stats_out[i] = 0.f;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
    else
      cudaGetLastError();
  });
  return fn;
}

static bool encode_map(CUtensorMap* tm, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                       const cuuint32_t* box, const cuuint32_t* estr, const char* what, bool f16 = false) {
  EncodeTiledFn enc = get_encode();
  // cuTensorMapEncodeTiled is a DRIVER call: it needs a current context on the calling thread.  An autograd worker thread whose
  // first library call is a conv with cached packings reaches this point before any runtime call has bound the primary context
  // (CUDA_ERROR_INVALID_CONTEXT, found by tests/test_conv_skip.py) — bind it once per thread.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
  if (!enc) {
    set_error("tcgen05 engine: cuTensorMapEncodeTiled is unavailable in this driver");
    return false;
  }
  CUresult r = enc(tm, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("tcgen05 engine: cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
    return false;
  }
  return true;
}

static const size_t SMEM_LIMIT = 227 * 1024;

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// choose TW x TH x TN = 128 (powers of two) minimising the number of tiles; ties -> wider TW
static void pick_tile(int n, int h, int w, int stride, int* tw_log, int* th_log) {
  long long best = -1;
  int bw = 7, bh = 0;
  for (int a = 7; a >= 0; --a) {
    if (((1 << a) * stride) > 256) continue;
    for (int b = 7 - a; b >= 0; --b) {
      if (((1 << b) * stride) > 256) continue;
      const int c = 7 - a - b;
      const long long tiles = (long long)((w + (1 << a) - 1) >> a) * ((h + (1 << b) - 1) >> b) * ((n + (1 << c) - 1) >> c);
      if (best < 0 || tiles < best) {
        best = tiles;
        bw = a;
        bh = b;
      }
    }
  }
  *tw_log = bw;
  *th_log = bh;
}

// TMA-store copy-out of the epilogue (on unless CGB_TMA_STORE=0): the tile leaves shared memory as bulk tensor stores of
// 64-channel halves, so the N tile must be whole halves or the only N tile (the tensor map clips the channel tail)
static bool tma_store_enabled() {
  static const int v = getenv("CGB_TMA_STORE") ? atoi(getenv("CGB_TMA_STORE")) : 1;
  return v != 0;
}
// the epilogue's second operand by TMA into the staging tile (struct EpiOperand); CGB_AUX_TMA=0: row-per-thread global loads,
// and the ReLU-masked dgrad back on the per-thread copy-out
static bool aux_tma_enabled() {
  static const int v = getenv("CGB_AUX_TMA") ? atoi(getenv("CGB_AUX_TMA")) : 1;
  return v != 0;
}
static bool tma_store_ok(int bn, int n_tiles, int dact, const void* mask_src) {
  // without the operand-by-TMA epilogue (CGB_AUX_TMA=0) the ReLU-derivative mask stays at the per-thread copy-out, where its loads are
  // coalesced (consecutive lanes = consecutive chunks
  // of a pixel): read row-per-thread in phase 1 it cost 32 sectors in 32 lines per LDG (256->256 d2 dgrad: 54 -> 71 us)
  static const int mask_tma = getenv("CGB_MASK_TMA") ? atoi(getenv("CGB_MASK_TMA")) : 0;
  return tma_store_enabled() && (bn % 64 == 0 || n_tiles == 1) && (mask_tma || aux_tma_enabled() || !(mask_src && dact == CGB_ACT_RELU));
}
// second operand by TMA: a TMA-store launch with exactly one of residual / mask, an epilogue on the lean chunk code (no tanh /
// sigmoid / selu), no statistics pass over the staging tile, and every 64-channel half of every N tile starting inside the tensor
static bool aux_tma_ok(bool tma_store, int bn, int n_tiles, int cout_s, int act, const void* residual, const void* mask_src,
                       const float* stats_out) {
  if (!aux_tma_enabled() || !tma_store || stats_out || act > CGB_ACT_LRELU) return false;
  if ((residual != nullptr) == (mask_src != nullptr)) return false;
  const int last_cn0 = (n_tiles - 1) * bn, nh = (bn + 63) / 64;
  return last_cn0 + (nh - 1) * 64 < cout_s;
}
static int staging_tile_bytes_for(int bn, bool tma) {
  const int plain = 128 * (bn * 2 + 16);
  if (!tma) return plain;
  const int halves = ((bn + 63) / 64) * (128 * 128);
  return ((halves > plain ? halves : plain) + 1023) / 1024 * 1024;
}

static int pick_bn(int co) {
  // largest multiple of 16 <= 256 that tiles co with the least padded work
  const int co16 = (co + 15) / 16 * 16;
  if (co16 <= 256) return co16;
  int best_bn = 256;
  long long best_waste = -1;
  for (int bn = 256; bn >= 64; bn -= 16) {
    const int tiles = (co + bn - 1) / bn;
    const long long waste = (long long)tiles * bn - co;
    if (best_waste < 0 || waste < best_waste) {
      best_waste = waste;
      best_bn = bn;
    }
  }
  return best_bn;
}

bool conv_tc_supported(const cgb_conv_desc* d, int which) {
  // bf16: every operator; fp16 (the inference-only --half mode): fprop and dgrad (tcgen05 kind::f16 takes either operand
  // format); the fp16 weight gradient stays on the CUDA-core engine — nothing on the path trains in fp16
  if (d->dtype != CGB_BF16 && !(d->dtype == CGB_F16 && which != 2)) return false;
  if (d->pad_mode != CGB_PAD_ZERO && d->pad > 0) return false;
  if (which == 0) return d->stride <= 2;
  if (which == 1) {
    if (d->stride == 1) return d->pad <= d->dil * (d->kh - 1) && d->pad <= d->dil * (d->kw - 1);
    return d->stride == 2 && d->kh * d->kw <= 64;  // parity-class decomposition
  }
  return d->stride <= 2 && d->co <= 2048;  // wgrad
}

// debug only (CGB_TC_TRACE=1): per-role clock64 stamps of CTA 0, printed after a synchronisation — never on in production
static unsigned long long* g_trace_buf = nullptr;
static bool trace_on() {
  static const int v = getenv("CGB_TC_TRACE") ? atoi(getenv("CGB_TC_TRACE")) : 0;
  return v != 0;
}
static unsigned long long* trace_begin(cudaStream_t st) {
  if (!trace_on()) return nullptr;
  if (!g_trace_buf) cudaMallocManaged(&g_trace_buf, 32 * 16 * sizeof(unsigned long long));
  cudaStreamSynchronize(st);
  memset(g_trace_buf, 0, 32 * 16 * sizeof(unsigned long long));
  return g_trace_buf;
}
static void trace_end(const TcParams& p, cudaStream_t st) {
  if (!trace_on() || !g_trace_buf) return;
  cudaStreamSynchronize(st);
  const unsigned long long t0 = g_trace_buf[0];
  fprintf(stderr, "[tc_trace] bn=%d n_tiles=%d tiles=%d stages=%d iters=%d tma_store=%d staging_bufs=%d\n", p.bn, p.n_tiles,
          p.total_tiles ? p.total_tiles : p.pix_tiles, p.stages, (p.ntaps ? p.ntaps : 1) * p.kblocks, p.tma_store, p.staging_bufs);
  fprintf(stderr, "[tc_trace] tile: P.start P.issued | M.tempty M.full0 M.fullN | E.wait E.tfull E.stg_free E.ph1 E.bar E.done (cycles since start)\n");
  for (int t = 0; t < 32 && g_trace_buf[t * 16] != 0; ++t) {
    fprintf(stderr, "[tc_trace] %2d:", t);
    for (int k = 0; k <= 10; ++k)
      fprintf(stderr, " %7lld%s", (long long)(g_trace_buf[t * 16 + k] - t0), (k == 1 || k == 4) ? " |" : "");
    fprintf(stderr, "\n");
  }
}

// ---- epilogue variant of a launch (template parameter EPI of the kernels) ---------------------------------------------------
static int epi_variant_for(bool tma_store, bool f16, const float* bias, const void* residual, const void* mask_src, int act, int dact) {
  static const int on = getenv("CGB_EPI_VARIANTS") ? atoi(getenv("CGB_EPI_VARIANTS")) : 1;
  if (!tma_store) return EPI_NOTMA;
  if (act > CGB_ACT_LRELU || (residual && mask_src)) return EPI_EXOTIC;
  if (!on || f16) return EPI_GENERIC;   // (fp16: inference-only mode, run-time flags)
  if (residual) return (!bias && act == CGB_ACT_NONE && !mask_src) ? EPI_RES : EPI_GENERIC;
  if (mask_src) {
    if (bias || act != CGB_ACT_NONE) return EPI_GENERIC;
    return dact == CGB_ACT_RELU ? EPI_MASK_RELU : (dact == CGB_ACT_LRELU ? EPI_MASK_LRELU : EPI_GENERIC);
  }
  if (!bias) return act == CGB_ACT_NONE ? EPI_PLAIN : EPI_GENERIC;
  return act == CGB_ACT_NONE ? EPI_BIAS : EPI_BIAS_ACT;
}

typedef void (*TcKernelBf16)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const TcParams, const float*, const __nv_bfloat16*,
                             const __nv_bfloat16*, __nv_bfloat16*, float*);
typedef void (*TcKernelF16)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const TcParams, const float*, const __half*,
                            const __half*, __half*, float*);
#define CGB_KERNEL_TABLES(NAME, KERNEL)                                                  \
  static TcKernelBf16 NAME##_bf16(int epi) {                                             \
    switch (epi) {                                                                       \
      case EPI_PLAIN: return KERNEL<__nv_bfloat16, EPI_PLAIN>;                           \
      case EPI_BIAS: return KERNEL<__nv_bfloat16, EPI_BIAS>;                             \
      case EPI_BIAS_ACT: return KERNEL<__nv_bfloat16, EPI_BIAS_ACT>;                     \
      case EPI_MASK_RELU: return KERNEL<__nv_bfloat16, EPI_MASK_RELU>;                   \
      case EPI_MASK_LRELU: return KERNEL<__nv_bfloat16, EPI_MASK_LRELU>;                 \
      case EPI_EXOTIC: return KERNEL<__nv_bfloat16, EPI_EXOTIC>;                         \
      case EPI_NOTMA: return KERNEL<__nv_bfloat16, EPI_NOTMA>;                           \
      case EPI_RES: return KERNEL<__nv_bfloat16, EPI_RES>;                               \
      default: return KERNEL<__nv_bfloat16, EPI_GENERIC>;                                \
    }                                                                                    \
  }                                                                                      \
  static TcKernelF16 NAME##_f16(int epi) {   /* fp16: generic / exotic / per-thread */   \
    switch (epi) {                                                                       \
      case EPI_EXOTIC: return KERNEL<__half, EPI_EXOTIC>;                                \
      case EPI_NOTMA: return KERNEL<__half, EPI_NOTMA>;                                  \
      default: return KERNEL<__half, EPI_GENERIC>;                                       \
    }                                                                                    \
  }
CGB_KERNEL_TABLES(stream_kernel, conv_tc_kernel)
CGB_KERNEL_TABLES(pair_kernel, conv_tc2_kernel)
CGB_KERNEL_TABLES(ws_kernel, conv_tc_ws_kernel)
#undef CGB_KERNEL_TABLES

// ---- streaming launch with an explicit tap table and output mapping ---------------------------------------------
struct TapTable {
  int ntaps;
  short dy[64], dx[64], w[64];
};

static int launch_stream(const void* in, const void* w, void* out, int n, int hin, int win, int cin_s, int hgrid, int wgrid,
                         int cout_s, int wtaps_total, const TapTable& tt, int in_stride, int out_stride, int out_off_y,
                         int out_off_x, int hfull, int wfull, int act, float slope, const float* bias,
                         const void* residual, int dact, const void* mask_src, cudaStream_t st, int res_before_act = 0,
                         float* stats_out = nullptr, bool f16 = false) {
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.res_before_act = res_before_act;
  p.stats = stats_out ? 1 : 0;
  p.stats_rows = num_sms();
  const int stats_bytes = stats_out ? 2 * cout_s * 4 + 16 : 0;
  p.n = n; p.hout = hgrid; p.wout = wgrid; p.cout_s = cout_s; p.cin_s = cin_s;
  p.stride = in_stride;
  p.kblocks = (cin_s + 63) / 64;
  p.act = act; p.slope = slope; p.dact = dact;
  p.ntaps = tt.ntaps;
  for (int t = 0; t < tt.ntaps; ++t) { p.tap_dy[t] = tt.dy[t]; p.tap_dx[t] = tt.dx[t]; p.tap_w[t] = tt.w[t]; }
  p.out_stride = out_stride; p.out_off_y = out_off_y; p.out_off_x = out_off_x; p.hfull = hfull; p.wfull = wfull;
  pick_tile(n, hgrid, wgrid, in_stride, &p.tw_log, &p.th_log);
  const int tn_log = 7 - p.tw_log - p.th_log;
  p.tiles_x = (wgrid + (1 << p.tw_log) - 1) >> p.tw_log;
  p.tiles_y = (hgrid + (1 << p.th_log) - 1) >> p.th_log;
  const int tiles_n = (n + (1 << tn_log) - 1) >> tn_log;
  p.bn = pick_bn(cout_s);
  p.n_tiles = (cout_s + p.bn - 1) / p.bn;
  p.total_tiles = p.tiles_x * p.tiles_y * tiles_n * p.n_tiles;
  p.mg_nt = magic_for(p.total_tiles + 2, p.n_tiles);
  p.mg_tx = magic_for(p.total_tiles + 2, p.tiles_x);
  p.mg_ty = magic_for(p.total_tiles + 2, p.tiles_y);
  p.stage_pitch = p.bn * 2 + 16;
  // TMA-store copy-out: plain output mapping only (the strided parity-class dgrad keeps the per-thread copy-out)
  p.tma_store = (tma_store_ok(p.bn, p.n_tiles, dact, mask_src) && out_stride == 1 && out_off_y == 0 && out_off_x == 0 &&
                 hfull == hgrid && wfull == wgrid) ? 1 : 0;
  p.staging_bufs = (p.tma_store && p.bn <= 128) ? 2 : 1;
  p.staging_tile_bytes = staging_tile_bytes_for(p.bn, p.tma_store != 0);
  const int staging_bytes = p.staging_tile_bytes * p.staging_bufs;
  int cols = 32;
  while (cols < 2 * p.bn) cols <<= 1;
  p.tmem_cols = cols;
  const int stage_bytes = A_TILE_BYTES + p.bn * 128;
  int stages = (int)((SMEM_LIMIT - 1024 - 320 - staging_bytes - stats_bytes) / stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) stages = 2;
  p.stages = stages;
  p.stats_off = stages * stage_bytes + staging_bytes + 16 * stages + 64;
  p.aux_bar_off = p.stats_off + stats_bytes;
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[4] = {(cuuint64_t)cin_s, (cuuint64_t)win, (cuuint64_t)hin, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)cin_s * 2, (cuuint64_t)win * cin_s * 2, (cuuint64_t)hin * win * cin_s * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)((1 << p.tw_log) * in_stride), (cuuint32_t)((1 << p.th_log) * in_stride),
                         (cuuint32_t)(1 << tn_log)};
    cuuint32_t estr[4] = {1, (cuuint32_t)in_stride, (cuuint32_t)in_stride, 1};
    if (!encode_map(&tmA, in, 4, dims, strides, box, estr, "activations", f16)) return CGB_LAUNCH_FAILURE;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)cin_s, (cuuint64_t)wtaps_total, (cuuint64_t)cout_s};
    cuuint64_t strides[2] = {(cuuint64_t)cin_s * 2, (cuuint64_t)wtaps_total * cin_s * 2};
    cuuint32_t box[3] = {64, 1, (cuuint32_t)p.bn};
    cuuint32_t estr[3] = {1, 1, 1};
    if (!encode_map(&tmB, w, 3, dims, strides, box, estr, "weights", f16)) return CGB_LAUNCH_FAILURE;
  }
  CUtensorMap tmY = tmB;   // (any valid map when the TMA store is off: the kernel never touches it)
  CUtensorMap tmX = tmB;   // (likewise when the second operand does not come by TMA)
  p.aux_tma = aux_tma_ok(p.tma_store != 0, p.bn, p.n_tiles, cout_s, act, residual, mask_src, stats_out) ? 1 : 0;
  if (p.tma_store) {
    cuuint64_t dims[4] = {(cuuint64_t)cout_s, (cuuint64_t)wfull, (cuuint64_t)hfull, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)cout_s * 2, (cuuint64_t)wfull * cout_s * 2, (cuuint64_t)hfull * wfull * cout_s * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(1 << p.tw_log), (cuuint32_t)(1 << p.th_log), (cuuint32_t)(1 << tn_log)};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (!encode_map(&tmY, out, 4, dims, strides, box, estr, "output", f16)) return CGB_LAUNCH_FAILURE;
    if (p.aux_tma && !encode_map(&tmX, residual ? residual : mask_src, 4, dims, strides, box, estr, "epilogue operand", f16))
      return CGB_LAUNCH_FAILURE;
  }
  static std::once_flag attr_once;
  std::call_once(attr_once, [] {
    for (int e = 0; e < EPI_VARIANTS; ++e) {
      cudaFuncSetAttribute(stream_kernel_bf16(e), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
      cudaFuncSetAttribute(stream_kernel_f16(e), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
    }
  });
  // ---- CTA-pair form (cta_group::2): half a weight tile per CTA — for launches with enough pixel tiles to fill 74 pairs
  static const int tc2 = getenv("CGB_TC2") ? atoi(getenv("CGB_TC2")) : 1;
  {
    const int pix_tiles = p.total_tiles / p.n_tiles;
    const int pair_tiles = ((pix_tiles + 1) / 2) * p.n_tiles;
    // measured (scripts/exp/tc2_check.py, profiles/r02_tc2_pair_kernel.txt; with the compile-time epilogue variants:
    // gpurun_out/g24_*.txt): the pair form wins from K = 512 on (512->2048 1x1: 108 -> 104 us, 128->160 3x3: 77 -> 70, ASPP
    // 2048->256 d12: 386 -> 325) and loses at K = 256 (256->1024: 45 -> 53 us): CGB_TC2=1 takes it from tc2_min_k on, 2 always
    static const int tc2_min_k = getenv("CGB_TC2_MIN_K") ? atoi(getenv("CGB_TC2_MIN_K")) : 512;
    const bool k_ok = tc2 >= 2 || (long long)tt.ntaps * cin_s >= tc2_min_k;
    if (tc2 && k_ok && p.bn % 16 == 0 && pair_tiles >= num_sms() / 2) {
      const int stage2 = A_TILE_BYTES + p.bn * 64;
      int stages2 = (int)((SMEM_LIMIT - 1024 - 320 - staging_bytes - stats_bytes) / stage2);
      if (stages2 > 10) stages2 = 10;
      p.stages = stages2;
      p.stats_off = stages2 * stage2 + staging_bytes + 16 * stages2 + 64;
      p.aux_bar_off = p.stats_off + stats_bytes;
      // the pair's weight map delivers bn/2 rows per load
      {
        cuuint64_t dims[3] = {(cuuint64_t)cin_s, (cuuint64_t)wtaps_total, (cuuint64_t)cout_s};
        cuuint64_t strides[2] = {(cuuint64_t)cin_s * 2, (cuuint64_t)wtaps_total * cin_s * 2};
        cuuint32_t box[3] = {64, 1, (cuuint32_t)(p.bn / 2)};
        cuuint32_t estr[3] = {1, 1, 1};
        if (!encode_map(&tmB, w, 3, dims, strides, box, estr, "weights (half tile)", f16)) return CGB_LAUNCH_FAILURE;
      }
      const size_t smem2 = (size_t)stages2 * stage2 + staging_bytes + 16 * stages2 + 64 + stats_bytes + 64 + 1024;
      static std::once_flag attr2_once;
      std::call_once(attr2_once, [] {
        for (int e = 0; e < EPI_VARIANTS; ++e) {
          cudaFuncSetAttribute(pair_kernel_bf16(e), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
          cudaFuncSetAttribute(pair_kernel_f16(e), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
        }
      });
      int npairs = num_sms() / 2;
      if (npairs > pair_tiles) npairs = pair_tiles;
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3((unsigned)(2 * npairs));
      cfg.blockDim = dim3(TC_THREADS);
      cfg.dynamicSmemBytes = smem2;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      cudaError_t e;
      const int epi2 = epi_variant_for(p.tma_store != 0, f16, bias, residual, mask_src, act, dact);
      if (f16)
        e = cudaLaunchKernelEx(&cfg, pair_kernel_f16(epi2), tmA, tmB, tmY, tmX, p, bias, (const __half*)residual, (const __half*)mask_src,
                               (__half*)out, stats_out);
      else
        e = cudaLaunchKernelEx(&cfg, pair_kernel_bf16(epi2), tmA, tmB, tmY, tmX, p, bias, (const __nv_bfloat16*)residual,
                               (const __nv_bfloat16*)mask_src, (__nv_bfloat16*)out, stats_out);
      if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("conv_tc2: cluster launch failed: %s", cudaGetErrorString(e));
        return CGB_LAUNCH_FAILURE;
      }
      return after_launch("conv_tc2");
    }
  }
  const size_t smem = (size_t)stages * stage_bytes + staging_bytes + 16 * stages + 64 + stats_bytes + 64 + 1024;
  if (smem > SMEM_LIMIT) {
    set_error("tcgen05 engine: tile does not fit shared memory (bn=%d, stats=%d)", p.bn, p.stats);
    return CGB_UNSUPPORTED;
  }
  dim3 grid((unsigned)(p.total_tiles < num_sms() ? p.total_tiles : num_sms()));
  p.trace = trace_begin(st);   // debug only (CGB_TC_TRACE=1), nullptr otherwise
  const int epi = epi_variant_for(p.tma_store != 0, f16, bias, residual, mask_src, act, dact);
  if (f16)
    stream_kernel_f16(epi)<<<grid, TC_THREADS, smem, st>>>(tmA, tmB, tmY, tmX, p, bias, (const __half*)residual, (const __half*)mask_src,
                                                           (__half*)out, stats_out);
  else
    stream_kernel_bf16(epi)<<<grid, TC_THREADS, smem, st>>>(tmA, tmB, tmY, tmX, p, bias, (const __nv_bfloat16*)residual,
                                                            (const __nv_bfloat16*)mask_src, (__nv_bfloat16*)out, stats_out);
  trace_end(p, st);
  return after_launch("conv_tc");
}

// Launch an fprop on (in -> out).  in: [n,hin,win,cin_s], w: [cout_s][taps][cin_s], out: [n,hout,wout,cout_s]
static int launch_fprop(const void* in, const void* w, void* out, int n, int hin, int win, int cin_s, int hout, int wout,
                        int cout_s, int kh, int kw, int stride, int dil, int pad_y, int pad_x, int act, float slope,
                        const float* bias, const void* residual, int dact, const void* mask_src, cudaStream_t st,
                        int res_before_act = 0, float* stats_out = nullptr, bool f16 = false) {
  const int taps = kh * kw;
  const int stats_bytes = stats_out ? 2 * cout_s * 4 + 16 : 0;
  // ---- traffic estimate of the streaming configuration
  int s_tw_log, s_th_log;
  pick_tile(n, hout, wout, stride, &s_tw_log, &s_th_log);
  const int s_tn_log = 7 - s_tw_log - s_th_log;
  const int kblocks = (cin_s + 63) / 64;
  const int s_bn = pick_bn(cout_s);
  const double stream_bytes = (double)((wout + (1 << s_tw_log) - 1) >> s_tw_log) * ((hout + (1 << s_th_log) - 1) >> s_th_log) *
                              ((n + (1 << s_tn_log) - 1) >> s_tn_log) * ((cout_s + s_bn - 1) / s_bn) * taps * kblocks *
                              (A_TILE_BYTES + s_bn * 128.0);
  // ---- weight-stationary halo configuration (stride-1 k>1 convs on large maps)
  bool use_ws = false;
  int ws_bn = 0, ws_ntiles = 0, ws_stages = 0;
  const int twh = 8 + dil * (kw - 1), thh = 16 + dil * (kh - 1);
  const int a_stage = ((twh * thh * 128) + 1023) / 1024 * 1024;
  // 1x1 convs with many output channels (ResNet 256->1024) re-stream 128 KB of weights per 128-pixel tile in the streaming
  // kernel (L2 -> SM bound); holding a >= 128-channel weight slice resident makes them activation-bound instead
  static const int ws_1x1 = getenv("CGB_WS_1X1") ? atoi(getenv("CGB_WS_1X1")) : 0;
  if (stride == 1 && (taps > 1 || ws_1x1) && wout >= 24 && hout >= 32 && twh <= 256 && thh <= 256 && a_stage <= 48 * 1024) {
    for (int nt = 1; nt <= 8 && !use_ws; ++nt) {
      int bn = ((cout_s + nt - 1) / nt + 15) / 16 * 16;
      if (bn > 256) continue;
      const size_t wb = (size_t)kblocks * taps * bn * 128;
      const size_t fixed = wb + (size_t)staging_tile_bytes_for(bn, tma_store_ok(bn, nt, dact, mask_src)) + 1024 + 320 + stats_bytes;
      if (fixed + 2 * (size_t)a_stage > SMEM_LIMIT) continue;
      int stg = (int)((SMEM_LIMIT - fixed) / a_stage);
      if (stg > 6) stg = 6;
      // a resident-weight 1x1 needs a deep activation ring (each 16 KB stage is only 4 MMAs of work): prefer a narrower N
      // slice with >= ws_1x1_min_stages stages over the widest slice with two
      static const int ws_1x1_min_stages = getenv("CGB_WS_1X1_MIN_STAGES") ? atoi(getenv("CGB_WS_1X1_MIN_STAGES")) : 4;
      if (taps == 1 && ws_1x1 && stg < ws_1x1_min_stages && nt < 8) continue;
      const double ws_bytes = (double)((wout + 7) / 8) * ((hout + 15) / 16) * n * nt * kblocks * (double)(twh * thh * 128);
      // Measured on B200 (gpurun_out/ws_sweep.log -> profiles/r01_ws_vs_streaming.txt): the weight-stationary kernel only
      // wins for narrow outputs fed by few channels (gamma||beta 128->48 @640^2: 1.22 vs 1.50 ms; 24->24: 0.59 vs 2.03 ms).
      // As soon as the weights need several slices, or N >= 80, the streaming kernel is 1.3-3x faster (256->256 d2 @80^2:
      // 0.054 vs 0.153 ms; 128->160 @160^2: 0.143 vs 0.250 ms): narrow slices make the MMA itself inefficient (cost
      // max(N/2, (4096+32N)/128) cycles, scripts/exp/mma_rate.cu) and re-read the halo once per slice.
      static const int ws_max_co = getenv("CGB_WS_MAX_CO") ? atoi(getenv("CGB_WS_MAX_CO")) : 64;
      // ... and for wide outputs fed by a single 64-channel block (the dgrad of gamma||beta, 48->128: 1.82 vs 1.97 ms)
      if (ws_bytes < 0.8 * stream_bytes && (cout_s <= ws_max_co || kblocks == 1 || (taps == 1 && ws_1x1 && bn >= 128))) {
        use_ws = true; ws_bn = bn; ws_ntiles = nt; ws_stages = stg;
      }
      break;  // the smallest feasible n-split is the cheapest in activation re-reads
    }
  }
  if (!use_ws) {
    if (taps > 64) {
      set_error("tcgen05 engine: more than 64 filter taps");
      return CGB_UNSUPPORTED;
    }
    TapTable tt;
    tt.ntaps = taps;
    for (int t = 0; t < taps; ++t) {
      tt.dy[t] = (short)((t / kw) * dil - pad_y);
      tt.dx[t] = (short)((t % kw) * dil - pad_x);
      tt.w[t] = (short)t;
    }
    return launch_stream(in, w, out, n, hin, win, cin_s, hout, wout, cout_s, taps, tt, stride, 1, 0, 0, hout, wout, act, slope,
                         bias, residual, dact, mask_src, st, res_before_act, stats_out, f16);
  }

  TcParams p;
  memset(&p, 0, sizeof(p));
  p.res_before_act = res_before_act;
  p.n = n; p.hout = hout; p.wout = wout; p.cout_s = cout_s; p.cin_s = cin_s;
  p.kh = kh; p.kw = kw; p.dil = dil; p.stride = stride; p.pad_y = pad_y; p.pad_x = pad_x;
  p.kblocks = kblocks;
  p.act = act; p.slope = slope; p.dact = dact;
  p.out_stride = 1; p.hfull = hout; p.wfull = wout;
  p.stats = stats_out ? 1 : 0;
  p.stats_rows = num_sms();
  static std::once_flag attr_once;
  std::call_once(attr_once, [] {
    for (int e = 0; e < EPI_VARIANTS; ++e) {
      cudaFuncSetAttribute(ws_kernel_bf16(e), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
      cudaFuncSetAttribute(ws_kernel_f16(e), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
    }
  });
  p.tw_log = 3; p.th_log = 4;
  p.tiles_x = (wout + 7) / 8; p.tiles_y = (hout + 15) / 16;
  p.pix_tiles = p.tiles_x * p.tiles_y * n;
  p.mg_tx = magic_for(p.pix_tiles + 2, p.tiles_x);
  p.mg_ty = magic_for(p.pix_tiles + 2, p.tiles_y);
  p.bn = ws_bn; p.n_tiles = ws_ntiles; p.stages = ws_stages;
  p.twh = twh; p.thh = thh; p.a_stage_bytes = a_stage;
  p.stage_pitch = p.bn * 2 + 16;
  p.tma_store = tma_store_ok(p.bn, p.n_tiles, dact, mask_src) ? 1 : 0;
  p.staging_tile_bytes = staging_tile_bytes_for(p.bn, p.tma_store != 0);
  {
    // second staging tile if at least three activation stages remain (ws_stages was sized with one tile)
    static const int ws_stg2 = getenv("CGB_WS_STAGING2") ? atoi(getenv("CGB_WS_STAGING2")) : 1;
    const size_t fixed2 = (size_t)kblocks * taps * p.bn * 128 + 2 * (size_t)p.staging_tile_bytes + 1024 + 320 + stats_bytes;
    int stg2 = fixed2 < SMEM_LIMIT ? (int)((SMEM_LIMIT - fixed2) / a_stage) : 0;
    if (stg2 > 6) stg2 = 6;
    p.staging_bufs = (ws_stg2 && p.tma_store && stg2 >= 3) ? 2 : 1;
    if (p.staging_bufs == 2) p.stages = stg2;
  }
  static const int ws_unroll = getenv("CGB_WS_UNROLL") ? atoi(getenv("CGB_WS_UNROLL")) : 1;
  p.ws_unroll = ws_unroll;
  const int staging_bytes = p.staging_tile_bytes * p.staging_bufs;
  int cols = 32;
  while (cols < 2 * p.bn) cols <<= 1;
  p.tmem_cols = cols;
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[4] = {(cuuint64_t)cin_s, (cuuint64_t)win, (cuuint64_t)hin, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)cin_s * 2, (cuuint64_t)win * cin_s * 2, (cuuint64_t)hin * win * cin_s * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)twh, (cuuint32_t)thh, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (!encode_map(&tmA, in, 4, dims, strides, box, estr, "activations", f16)) return CGB_LAUNCH_FAILURE;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)cin_s, (cuuint64_t)taps, (cuuint64_t)cout_s};
    cuuint64_t strides[2] = {(cuuint64_t)cin_s * 2, (cuuint64_t)taps * cin_s * 2};
    cuuint32_t box[3] = {64, 1, (cuuint32_t)p.bn};
    cuuint32_t estr[3] = {1, 1, 1};
    if (!encode_map(&tmB, w, 3, dims, strides, box, estr, "weights", f16)) return CGB_LAUNCH_FAILURE;
  }
  CUtensorMap tmY = tmB;   // (any valid map when the TMA store is off: the kernel never touches it)
  CUtensorMap tmX = tmB;
  p.aux_tma = aux_tma_ok(p.tma_store != 0, p.bn, p.n_tiles, cout_s, act, residual, mask_src, stats_out) ? 1 : 0;
  if (p.tma_store) {
    cuuint64_t dims[4] = {(cuuint64_t)cout_s, (cuuint64_t)wout, (cuuint64_t)hout, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)cout_s * 2, (cuuint64_t)wout * cout_s * 2, (cuuint64_t)hout * wout * cout_s * 2};
    cuuint32_t box[4] = {64, 8, 16, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (!encode_map(&tmY, out, 4, dims, strides, box, estr, "output", f16)) return CGB_LAUNCH_FAILURE;
    if (p.aux_tma && !encode_map(&tmX, residual ? residual : mask_src, 4, dims, strides, box, estr, "epilogue operand", f16))
      return CGB_LAUNCH_FAILURE;
  }
  p.stats_off = (int)((size_t)p.kblocks * taps * p.bn * 128 + (size_t)p.stages * a_stage + staging_bytes + 16 * p.stages + 64);
  p.aux_bar_off = p.stats_off + stats_bytes;
  {
    static const int aux_l2 = getenv("CGB_AUX_L2") ? atoi(getenv("CGB_AUX_L2")) : 1;
    if (p.aux_tma && aux_l2 && p.staging_bufs == 1) p.aux_tma = 2;
  }
  const size_t smem = (size_t)p.stats_off + stats_bytes + 64 + 1024;
  int ctas = num_sms() / p.n_tiles * p.n_tiles;
  if (ctas > p.pix_tiles * p.n_tiles) ctas = p.pix_tiles * p.n_tiles;
  const int epi = epi_variant_for(p.tma_store != 0, f16, bias, residual, mask_src, act, dact);
  p.trace = trace_begin(st);
  if (f16)
    ws_kernel_f16(epi)<<<ctas, TC_THREADS, smem, st>>>(tmA, tmB, tmY, tmX, p, bias, (const __half*)residual, (const __half*)mask_src,
                                                       (__half*)out, stats_out);
  else
    ws_kernel_bf16(epi)<<<ctas, TC_THREADS, smem, st>>>(tmA, tmB, tmY, tmX, p, bias, (const __nv_bfloat16*)residual,
                                                        (const __nv_bfloat16*)mask_src, (__nv_bfloat16*)out, stats_out);
  trace_end(p, st);
  return after_launch("conv_tc_ws");
}

// stats_out (optional): [conv_tc_stats_rows()][2][co] fp32; row b receives CTA b's per-channel sum / sum of squares of the
// stored output, rows beyond the launch's grid are zeroed by CTA 0 (no memset launch), for cgb_bn_train_fwd_partials
int conv_tc_stats_rows() { return num_sms(); }
int conv_tc_fwd(const cgb_conv_desc* d, const void* x, const void* w, const float* bias, const void* residual, void* y,
                cudaStream_t st, float* stats_out) {
  return launch_fprop(x, w, y, d->n, d->hi, d->wi, d->ci, d->ho, d->wo, d->co, d->kh, d->kw, d->stride, d->dil, d->pad,
                      d->pad, d->act, d->slope, bias, residual, CGB_ACT_NONE, nullptr, st, d->res_before_act, stats_out,
                      d->dtype == CGB_F16);
}

// wt: dgrad packing [ci][taps][co] with the taps reversed (cgb_conv2d_pack_dgrad_weight)
int conv_tc_dgrad(const cgb_conv_desc* d, const void* gy, const void* wt, int dact, const void* mask_src, void* gx,
                  cudaStream_t st) {
  if (d->stride > 1) {
    // stride-s dgrad = s*s stride-1 sub-convolutions over gy, one per parity class (ry, rx) of the input pixel:
    //   gx[s*j+ry, s*i+rx] = sum over taps (dy,dx) with (ry+pad-dy) % s == 0 of gy[j + (ry+pad-dy)/s, ...] * W[dy][dx]
    // each written at the strided positions of gx by the epilogue.  wt is the dgrad packing [ci][T-1-t][co].
    const int s_ = d->stride, T = d->kh * d->kw;
    bool need_zero = false;
    for (int ry = 0; ry < s_; ++ry)
      for (int rx = 0; rx < s_; ++rx) {
        TapTable tt;
        tt.ntaps = 0;
        for (int dy = 0; dy < d->kh; ++dy) {
          if ((ry + d->pad - dy * d->dil) % s_ != 0) continue;
          for (int dx = 0; dx < d->kw; ++dx) {
            if ((rx + d->pad - dx * d->dil) % s_ != 0) continue;
            const int t = tt.ntaps++;
            // floor division is exact here (divisible); C division of a negative multiple of s is exact too
            tt.dy[t] = (short)((ry + d->pad - dy * d->dil) / s_);
            tt.dx[t] = (short)((rx + d->pad - dx * d->dil) / s_);
            tt.w[t] = (short)(T - 1 - (dy * d->kw + dx));
          }
        }
        if (tt.ntaps == 0) need_zero = true;
      }
    if (need_zero)
      cudaMemsetAsync(gx, 0, (size_t)d->n * d->hi * d->wi * d->ci * 2, st);
    for (int ry = 0; ry < s_; ++ry)
      for (int rx = 0; rx < s_; ++rx) {
        TapTable tt;
        tt.ntaps = 0;
        for (int dy = 0; dy < d->kh; ++dy) {
          if ((ry + d->pad - dy * d->dil) % s_ != 0) continue;
          for (int dx = 0; dx < d->kw; ++dx) {
            if ((rx + d->pad - dx * d->dil) % s_ != 0) continue;
            const int t = tt.ntaps++;
            tt.dy[t] = (short)((ry + d->pad - dy * d->dil) / s_);
            tt.dx[t] = (short)((rx + d->pad - dx * d->dil) / s_);
            tt.w[t] = (short)(T - 1 - (dy * d->kw + dx));
          }
        }
        const int hc = (d->hi - ry + s_ - 1) / s_, wc = (d->wi - rx + s_ - 1) / s_;
        if (tt.ntaps == 0 || hc <= 0 || wc <= 0) continue;
        int r = launch_stream(gy, wt, gx, d->n, d->ho, d->wo, d->co, hc, wc, d->ci, T, tt, 1, s_, ry, rx, d->hi, d->wi,
                              CGB_ACT_NONE, d->slope, nullptr, nullptr, dact, mask_src, st, 0, nullptr, d->dtype == CGB_F16);
        if (r) return r;
      }
    return CGB_OK;
  }
  return launch_fprop(gy, wt, gx, d->n, d->ho, d->wo, d->co, d->hi, d->wi, d->ci, d->kh, d->kw, 1, d->dil,
                      d->dil * (d->kh - 1) - d->pad, d->dil * (d->kw - 1) - d->pad, CGB_ACT_NONE, d->slope, nullptr, nullptr,
                      dact, mask_src, st, 0, nullptr, d->dtype == CGB_F16);
}

// ------------------------------------------------------------------------------------------------------
// wgrad kernel:  gw[co][tap][ci] += sum_pixels gy[pix][co] * x[pix shifted by tap][ci]
//   UMMA view: D_tap[M][N] += P^T[M][K] * Q^T[N][K]^T with K = 128 pixels per step; M is whichever of (ci, co)
//   is larger (fills the 128 TMEM lanes), N the other.  Both operands are MN-major in shared memory (a TMA box
//   row is one pixel = one K index holding 64 contiguous channels).  One TMEM accumulator per filter tap of the
//   CTA's tap group; the un-shifted operand (gy) is loaded once per pixel tile, the shifted one (x) once per tap.
// ------------------------------------------------------------------------------------------------------
struct WgParams {
  int n, ho, wo;               // output-pixel domain (the reduction runs over it)
  int kh, kw, dil, stride, pad;
  int tw_log, th_log, tiles_x, tiles_y, total_tiles, tiles_per_cta;
  int m_dim, n_dim;            // channel extents of the M-side / N-side operand
  int m_boxes, n_boxes;        // 64-channel TMA boxes per operand tile
  int bn;                      // UMMA N (multiple of 16)
  int taps_per_group, tap_groups;
  int x_is_m;                  // 1: M side = x (shifted per tap), N side = gy ; 0: M side = gy, N side = x
  int stages_s, stages_u;      // ring depths: shifted (x) tiles, un-shifted (gy) tiles
  int tmem_cols;
  long long sm, sn, st;        // gw index = m*sm + n*sn + tap*st
};

// 16-byte vector reduction into global memory (sm_90+): gw[0..3] += {a, b, c, d}
__device__ __forceinline__ void red_add_v4(float* dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(__uint_as_float(a)), "f"(__uint_as_float(b)),
               "f"(__uint_as_float(c)), "f"(__uint_as_float(d))
               : "memory");
}

constexpr int BOX_BYTES = 128 * 128;  // 128 pixels x 64 channels bf16
constexpr int WG_THREADS = 192;       // warp 0 producer, warp 1 MMA, warps 2..5 epilogue

__global__ void __launch_bounds__(WG_THREADS)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG, const WgParams p,
                float* __restrict__ gw) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int s_boxes = p.x_is_m ? p.m_boxes : p.n_boxes;  // boxes per shifted (x) tile
  const int u_boxes = p.x_is_m ? p.n_boxes : p.m_boxes;  // boxes per un-shifted (gy) tile
  const uint32_t s_bytes = (uint32_t)s_boxes * BOX_BYTES, u_bytes = (uint32_t)u_boxes * BOX_BYTES;  // allocation
  const uint32_t s_base = base;
  const uint32_t u_base = base + (uint32_t)p.stages_s * s_bytes;
  const uint32_t bar_base = u_base + (uint32_t)p.stages_u * u_bytes;
  auto s_full = [&](int i) { return bar_base + 8u * (uint32_t)i; };
  auto s_empty = [&](int i) { return bar_base + 8u * (uint32_t)(p.stages_s + i); };
  auto u_full = [&](int i) { return bar_base + 8u * (uint32_t)(2 * p.stages_s + i); };
  auto u_empty = [&](int i) { return bar_base + 8u * (uint32_t)(2 * p.stages_s + p.stages_u + i); };
  const uint32_t tmem_full_bar = bar_base + 16u * (uint32_t)(p.stages_s + p.stages_u);
  const uint32_t tmem_ptr_addr = tmem_full_bar + 8u;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * 128;
  const int grp = blockIdx.z % p.tap_groups;
  const int n0 = (blockIdx.z / p.tap_groups) * p.bn;
  const int taps = p.kh * p.kw;
  const int tap0 = grp * p.taps_per_group;
  int ntaps = taps - tap0;
  if (ntaps > p.taps_per_group) ntaps = p.taps_per_group;
  const int tile0 = blockIdx.x * p.tiles_per_cta;
  int tile1 = tile0 + p.tiles_per_cta;
  if (tile1 > p.total_tiles) tile1 = p.total_tiles;
  const int tn_log = 7 - p.tw_log - p.th_log;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmG);
    for (int i = 0; i < p.stages_s; ++i) { mbar_init(s_full(i), 1); mbar_init(s_empty(i), 1); }
    for (int i = 0; i < p.stages_u; ++i) { mbar_init(u_full(i), 1); mbar_init(u_empty(i), 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_addr, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  const int x_ch0 = p.x_is_m ? m0 : n0;   // first channel of the x / gy slices this CTA reduces
  const int g_ch0 = p.x_is_m ? n0 : m0;
  // the M-side tile always owns two 64-channel boxes of smem (UMMA reads 128 rows) but only the boxes that hold
  // real channels are loaded; accumulator rows beyond m_dim are never written out.
  const int m_load = (p.m_dim - m0 > 64) ? 2 : 1;
  const int s_load = p.x_is_m ? m_load : s_boxes;
  const int u_load = p.x_is_m ? u_boxes : m_load;

  if (warp == 0) {
    if (lane == 0) {
      int ss = 0, us = 0;
      uint32_t sph = 0, uph = 0;
      for (int tile = tile0; tile < tile1; ++tile) {
        const int tx = tile % p.tiles_x;
        const int ty = (tile / p.tiles_x) % p.tiles_y;
        const int tn = tile / (p.tiles_x * p.tiles_y);
        const int ox0 = tx << p.tw_log, oy0 = ty << p.th_log, img0 = tn << tn_log;
        mbar_wait(u_empty(us), uph ^ 1u);
        mbar_expect_tx(u_full(us), (uint32_t)u_load * BOX_BYTES);
        for (int b = 0; b < u_load; ++b)
          tma_load_4d(u_base + (uint32_t)us * u_bytes + (uint32_t)b * BOX_BYTES, &tmG, u_full(us), g_ch0 + b * 64, ox0, oy0, img0);
        if (++us == p.stages_u) { us = 0; uph ^= 1u; }
        for (int t = 0; t < ntaps; ++t) {
          const int tap = tap0 + t;
          const int dy = tap / p.kw, dx = tap - dy * p.kw;
          mbar_wait(s_empty(ss), sph ^ 1u);
          mbar_expect_tx(s_full(ss), (uint32_t)s_load * BOX_BYTES);
          for (int b = 0; b < s_load; ++b)
            tma_load_4d(s_base + (uint32_t)ss * s_bytes + (uint32_t)b * BOX_BYTES, &tmX, s_full(ss), x_ch0 + b * 64,
                        ox0 * p.stride - p.pad + dx * p.dil, oy0 * p.stride - p.pad + dy * p.dil, img0);
          if (++ss == p.stages_s) { ss = 0; sph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(p.bn, true, true);
    int ss = 0, us = 0;
    uint32_t sph = 0, uph = 0;
    for (int tile = tile0; tile < tile1; ++tile) {
      mbar_wait(u_full(us), uph);
      const uint32_t u_addr = u_base + (uint32_t)us * u_bytes;
      for (int t = 0; t < ntaps; ++t) {
        mbar_wait(s_full(ss), sph);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t s_addr = s_base + (uint32_t)ss * s_bytes;
          const uint32_t a_addr = p.x_is_m ? s_addr : u_addr;
          const uint32_t b_addr = p.x_is_m ? u_addr : s_addr;
          const uint32_t d_addr = tmem_base + (uint32_t)(t * p.bn);
          const uint32_t a_lo = desc_lo(a_addr, BOX_BYTES), b_lo = desc_lo(b_addr, BOX_BYTES);
          const uint32_t hi = desc_hi(1024u);
          uint32_t acc = tile > tile0 ? 1u : 0u;
#pragma unroll
          for (int k = 0; k < 8; ++k) {  // 8 x 16 pixels, 2048 B (16 rows) apart
            umma_bf16(d_addr, desc_join(a_lo + 128u * k, hi), desc_join(b_lo + 128u * k, hi), idesc, acc);
            acc = 1u;
          }
          umma_commit(s_empty(ss));
          if (t == ntaps - 1) {
            umma_commit(u_empty(us));
            if (tile == tile1 - 1) umma_commit(tmem_full_bar);
          }
        }
        __syncwarp();
        if (++ss == p.stages_s) { ss = 0; sph ^= 1u; }
      }
      if (++us == p.stages_u) { us = 0; uph ^= 1u; }
    }
  } else if (tile1 > tile0) {
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    const bool m_ok = m < p.m_dim;
    mbar_wait(tmem_full_bar, 0u);
    tc_fence_after();
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int t = 0; t < ntaps; ++t) {
      const long long tap_off = (long long)(tap0 + t) * p.st + (long long)m * p.sm;
      for (int c0 = 0; c0 < p.bn; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(t_row + (uint32_t)(t * p.bn + c0), r);
        tmem_ld_wait();
        if (!m_ok) continue;
        if (p.sn == 1) {
          // a lane's 16 accumulator columns are 16 consecutive floats of gw (M = output channels): four 16-byte vector
          // reductions instead of 16 scalar ones — the scalar form made this epilogue the kernel's bottleneck (60 % of the
          // warp samples in profiles/r01_wgrad_256to1024_k1_80x80_n8.txt).  n_dim, n0, c0 are multiples of 8: aligned.
          float* dst = gw + tap_off + (long long)(n0 + c0);
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            if (n0 + c0 + j < p.n_dim) red_add_v4(dst + j, r[j], r[j + 1], r[j + 2], r[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int nn = n0 + c0 + j;
            if (nn < p.n_dim) atomicAdd(gw + tap_off + (long long)nn * p.sn, __uint_as_float(r[j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------
// wgrad, CTA-PAIR form (cta_group::2; cluster (2,1,1) over two neighbouring M tiles — grid (m tiles, pixel splits, z)): the pair accumulates D_tap[256][bn] —
// CTA r the 128 M-side channels of ITS m tile, in its own TMEM — and each CTA stages only HALF of the N-side operand (bn/2
// channels = whole 64-channel boxes).  The N-side tensor is then pulled through L2 once per PAIR of m tiles instead of once
// per m tile: wgrad is the most L2-traffic-bound kernel of the engine (1024->256 1x1: 314 MB for 132 MB of operands).
// Same barrier protocol as conv_tc2_kernel: every load credits the LEADER's full barrier, the leader's commits multicast to
// both CTAs' empty / accumulator-ready barriers, the leader's warp 1 issues every MMA.
__global__ void __launch_bounds__(WG_THREADS)
wgrad_tc2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG, const WgParams p,
                 float* __restrict__ gw) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int nh_boxes = p.n_boxes >> 1;                                  // N-side boxes this CTA stages (its half)
  const int s_boxes = p.x_is_m ? p.m_boxes : nh_boxes;                  // boxes per shifted (x) tile, per CTA
  const int u_boxes = p.x_is_m ? nh_boxes : p.m_boxes;                  // boxes per un-shifted (gy) tile, per CTA
  const uint32_t s_bytes = (uint32_t)s_boxes * BOX_BYTES, u_bytes = (uint32_t)u_boxes * BOX_BYTES;
  const uint32_t s_base = base;
  const uint32_t u_base = base + (uint32_t)p.stages_s * s_bytes;
  const uint32_t bar_base = u_base + (uint32_t)p.stages_u * u_bytes;
  auto s_full = [&](int i) { return bar_base + 8u * (uint32_t)i; };
  auto s_empty = [&](int i) { return bar_base + 8u * (uint32_t)(p.stages_s + i); };
  auto u_full = [&](int i) { return bar_base + 8u * (uint32_t)(2 * p.stages_s + i); };
  auto u_empty = [&](int i) { return bar_base + 8u * (uint32_t)(2 * p.stages_s + p.stages_u + i); };
  const uint32_t tmem_full_bar = bar_base + 16u * (uint32_t)(p.stages_s + p.stages_u);
  const uint32_t tmem_ptr_addr = tmem_full_bar + 8u;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();            // == blockIdx.x & 1 (cluster (2,1,1): the m tiles run along grid x)
  const bool leader = rank == 0;
  const int m0 = blockIdx.x * 128;
  const int m_pair0 = (blockIdx.x & ~1) * 128;
  const int grp = blockIdx.z % p.tap_groups;
  const int n0 = (blockIdx.z / p.tap_groups) * p.bn;
  const int nh0 = n0 + (int)rank * (p.bn >> 1);       // first N-side channel this CTA stages
  const int taps = p.kh * p.kw;
  const int tap0 = grp * p.taps_per_group;
  int ntaps = taps - tap0;
  if (ntaps > p.taps_per_group) ntaps = p.taps_per_group;
  const int tile0 = blockIdx.y * p.tiles_per_cta;
  int tile1 = tile0 + p.tiles_per_cta;
  if (tile1 > p.total_tiles) tile1 = p.total_tiles;
  const int tn_log = 7 - p.tw_log - p.th_log;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmG);
    for (int i = 0; i < p.stages_s; ++i) { mbar_init(s_full(i), 1); mbar_init(s_empty(i), 1); }
    for (int i = 0; i < p.stages_u; ++i) { mbar_init(u_full(i), 1); mbar_init(u_empty(i), 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(tmem_ptr_addr, (uint32_t)p.tmem_cols);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  // M-side boxes that hold real channels, for this CTA and for the pair (the leader arms the barriers with the pair's bytes)
  auto m_load_of = [&](int mm0) { return (p.m_dim - mm0 > 64) ? 2 : 1; };
  const int m_load = m_load_of(m0);
  const int m_load_pair = m_load_of(m_pair0) + m_load_of(m_pair0 + 128);
  const int s_load = p.x_is_m ? m_load : nh_boxes, u_load = p.x_is_m ? nh_boxes : m_load;
  const int s_load_pair = p.x_is_m ? m_load_pair : 2 * nh_boxes, u_load_pair = p.x_is_m ? 2 * nh_boxes : m_load_pair;
  const int x_ch0 = p.x_is_m ? m0 : nh0;              // first channel of the x / gy slices THIS CTA stages
  const int g_ch0 = p.x_is_m ? nh0 : m0;

  if (warp == 0) {
    if (lane == 0) {
      int ss = 0, us = 0;
      uint32_t sph = 0, uph = 0;
      for (int tile = tile0; tile < tile1; ++tile) {
        const int tx = tile % p.tiles_x;
        const int ty = (tile / p.tiles_x) % p.tiles_y;
        const int tn = tile / (p.tiles_x * p.tiles_y);
        const int ox0 = tx << p.tw_log, oy0 = ty << p.th_log, img0 = tn << tn_log;
        mbar_wait(u_empty(us), uph ^ 1u);
        const uint32_t lead_u = mapa_rank(u_full(us), 0);
        if (leader) mbar_expect_tx(u_full(us), (uint32_t)u_load_pair * BOX_BYTES);
        for (int b = 0; b < u_load; ++b)
          tma2_load_4d(u_base + (uint32_t)us * u_bytes + (uint32_t)b * BOX_BYTES, &tmG, lead_u, g_ch0 + b * 64, ox0, oy0, img0);
        if (++us == p.stages_u) { us = 0; uph ^= 1u; }
        for (int t = 0; t < ntaps; ++t) {
          const int tap = tap0 + t;
          const int dy = tap / p.kw, dx = tap - dy * p.kw;
          mbar_wait(s_empty(ss), sph ^ 1u);
          const uint32_t lead_s = mapa_rank(s_full(ss), 0);
          if (leader) mbar_expect_tx(s_full(ss), (uint32_t)s_load_pair * BOX_BYTES);
          for (int b = 0; b < s_load; ++b)
            tma2_load_4d(s_base + (uint32_t)ss * s_bytes + (uint32_t)b * BOX_BYTES, &tmX, lead_s, x_ch0 + b * 64,
                         ox0 * p.stride - p.pad + dx * p.dil, oy0 * p.stride - p.pad + dy * p.dil, img0);
          if (++ss == p.stages_s) { ss = 0; sph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      const uint32_t idesc = make_idesc(p.bn, true, true, false, 256);
      int ss = 0, us = 0;
      uint32_t sph = 0, uph = 0;
      for (int tile = tile0; tile < tile1; ++tile) {
        mbar_wait(u_full(us), uph);
        const uint32_t u_addr = u_base + (uint32_t)us * u_bytes;
        for (int t = 0; t < ntaps; ++t) {
          mbar_wait(s_full(ss), sph);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t s_addr = s_base + (uint32_t)ss * s_bytes;
            const uint32_t a_addr = p.x_is_m ? s_addr : u_addr;
            const uint32_t b_addr = p.x_is_m ? u_addr : s_addr;
            const uint32_t d_addr = tmem_base + (uint32_t)(t * p.bn);
            const uint32_t a_lo = desc_lo(a_addr, BOX_BYTES), b_lo = desc_lo(b_addr, BOX_BYTES);
            const uint32_t hi = desc_hi(1024u);
            uint32_t acc = tile > tile0 ? 1u : 0u;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              umma2_bf16(d_addr, desc_join(a_lo + 128u * k, hi), desc_join(b_lo + 128u * k, hi), idesc, acc);
              acc = 1u;
            }
            umma2_commit_mc(s_empty(ss));
            if (t == ntaps - 1) {
              umma2_commit_mc(u_empty(us));
              if (tile == tile1 - 1) umma2_commit_mc(tmem_full_bar);
            }
          }
          __syncwarp();
          if (++ss == p.stages_s) { ss = 0; sph ^= 1u; }
        }
        if (++us == p.stages_u) { us = 0; uph ^= 1u; }
      }
    }
  } else if (tile1 > tile0) {
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    const bool m_ok = m < p.m_dim;
    mbar_wait(tmem_full_bar, 0u);
    tc_fence_after();
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int t = 0; t < ntaps; ++t) {
      const long long tap_off = (long long)(tap0 + t) * p.st + (long long)m * p.sm;
      for (int c0 = 0; c0 < p.bn; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(t_row + (uint32_t)(t * p.bn + c0), r);
        tmem_ld_wait();
        if (!m_ok) continue;
        if (p.sn == 1) {
          float* dst = gw + tap_off + (long long)(n0 + c0);
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            if (n0 + c0 + j < p.n_dim) red_add_v4(dst + j, r[j], r[j + 1], r[j + 2], r[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int nn = n0 + c0 + j;
            if (nn < p.n_dim) atomicAdd(gw + tap_off + (long long)nn * p.sn, __uint_as_float(r[j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------
// wgrad, halo variant (stride-1 k x k convs on large maps): the x tile is loaded ONCE per pixel tile as a halo box
// (16 + dil*(ndy-1)) x (8 + dil*(kw-1)) pixels and every tap of the CTA's dy-group is an MN-major UMMA descriptor into
// it (row-shifted start, 8-row group stride SBO = TWh*128 B) — instead of one shifted box per tap.  Tile = 16 rows x 8
// pixels of one image; K step = 16 pixels = two tile rows.
// ------------------------------------------------------------------------------------------------------
struct WgHaloParams {
  int n, ho, wo;
  int kh, kw, dil, pad;
  int tiles_x, tiles_y, total_tiles, tiles_per_cta;
  int m_dim, n_dim, bn, n_boxes;
  int rows_per_group, row_groups;  // dy rows per CTA, number of dy groups
  int x_is_m;
  int twh, thh;                    // halo extent for a full group
  int x_box_bytes;                 // 1024-aligned bytes of one 64-channel halo box
  int stages_x, stages_g;
  int tmem_cols;
  long long sm, sn, st;
};

__global__ void __launch_bounds__(WG_THREADS)
wgrad_tc_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG, const WgHaloParams p,
                     float* __restrict__ gw) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int x_boxes = p.x_is_m ? 2 : p.n_boxes;   // allocation (M side always owns two boxes)
  const int g_boxes = p.x_is_m ? p.n_boxes : 2;
  const uint32_t xs_bytes = (uint32_t)x_boxes * (uint32_t)p.x_box_bytes, gs_bytes = (uint32_t)g_boxes * BOX_BYTES;
  const uint32_t x_base = base;
  const uint32_t g_base = base + (uint32_t)p.stages_x * xs_bytes;
  const uint32_t bar_base = g_base + (uint32_t)p.stages_g * gs_bytes;
  auto x_full = [&](int i) { return bar_base + 8u * (uint32_t)i; };
  auto x_empty = [&](int i) { return bar_base + 8u * (uint32_t)(p.stages_x + i); };
  auto g_full = [&](int i) { return bar_base + 8u * (uint32_t)(2 * p.stages_x + i); };
  auto g_empty = [&](int i) { return bar_base + 8u * (uint32_t)(2 * p.stages_x + p.stages_g + i); };
  const uint32_t tmem_full_bar = bar_base + 16u * (uint32_t)(p.stages_x + p.stages_g);
  const uint32_t tmem_ptr_addr = tmem_full_bar + 8u;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * 128;
  const int grp = blockIdx.z % p.row_groups;
  const int n0 = (blockIdx.z / p.row_groups) * p.bn;
  const int dy0 = grp * p.rows_per_group;
  int ndy = p.kh - dy0;
  if (ndy > p.rows_per_group) ndy = p.rows_per_group;
  const int ntaps = ndy * p.kw;
  const int tile0 = blockIdx.x * p.tiles_per_cta;
  int tile1 = tile0 + p.tiles_per_cta;
  if (tile1 > p.total_tiles) tile1 = p.total_tiles;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmG);
    for (int i = 0; i < p.stages_x; ++i) { mbar_init(x_full(i), 1); mbar_init(x_empty(i), 1); }
    for (int i = 0; i < p.stages_g; ++i) { mbar_init(g_full(i), 1); mbar_init(g_empty(i), 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_addr, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  const int x_ch0 = p.x_is_m ? m0 : n0;
  const int g_ch0 = p.x_is_m ? n0 : m0;
  const int m_load = (p.m_dim - m0 > 64) ? 2 : 1;
  const int x_load = p.x_is_m ? m_load : x_boxes;
  const int g_load = p.x_is_m ? g_boxes : m_load;
  const uint32_t halo_bytes = (uint32_t)(p.twh * p.thh) * 128u;   // bytes one TMA box really delivers

  if (warp == 0) {
    if (elect_one_sync()) {
      int xs = 0, gs = 0;
      uint32_t xph = 0, gph = 0;
      for (int tile = tile0; tile < tile1; ++tile) {
        const int tx = tile % p.tiles_x;
        const int ty = (tile / p.tiles_x) % p.tiles_y;
        const int img = tile / (p.tiles_x * p.tiles_y);
        const int ox0 = tx << 3, oy0 = ty << 4;
        mbar_wait(g_empty(gs), gph ^ 1u);
        mbar_expect_tx(g_full(gs), (uint32_t)g_load * BOX_BYTES);
        for (int b = 0; b < g_load; ++b)
          tma_load_4d(g_base + (uint32_t)gs * gs_bytes + (uint32_t)b * BOX_BYTES, &tmG, g_full(gs), g_ch0 + b * 64, ox0, oy0, img);
        if (++gs == p.stages_g) { gs = 0; gph ^= 1u; }
        mbar_wait(x_empty(xs), xph ^ 1u);
        mbar_expect_tx(x_full(xs), (uint32_t)x_load * halo_bytes);
        for (int b = 0; b < x_load; ++b)
          tma_load_4d(x_base + (uint32_t)xs * xs_bytes + (uint32_t)b * (uint32_t)p.x_box_bytes, &tmX, x_full(xs), x_ch0 + b * 64,
                      ox0 - p.pad, oy0 - p.pad + dy0 * p.dil, img);
        if (++xs == p.stages_x) { xs = 0; xph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(p.bn, true, true);
    const uint32_t hi_x = desc_hi((uint32_t)p.twh * 128u), hi_g = desc_hi(1024u);
    const uint32_t kstep_x = (uint32_t)(2 * p.twh) * 8u;   // two halo rows per 16-pixel K step, in 16-byte units
    int xs = 0, gs = 0;
    uint32_t xph = 0, gph = 0;
    for (int tile = tile0; tile < tile1; ++tile) {
      mbar_wait(g_full(gs), gph);
      mbar_wait(x_full(xs), xph);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t x_lo0 = desc_lo(x_base + (uint32_t)xs * xs_bytes, (uint32_t)p.x_box_bytes);
        const uint32_t g_lo0 = desc_lo(g_base + (uint32_t)gs * gs_bytes, BOX_BYTES);
        const uint32_t acc0 = tile > tile0 ? 1u : 0u;
        int t = 0;
        for (int dy = 0; dy < ndy; ++dy) {
          for (int dx = 0; dx < p.kw; ++dx, ++t) {
            const uint32_t x_lo = x_lo0 + (uint32_t)((dy * p.dil) * p.twh + dx * p.dil) * 8u;
            const uint32_t d_addr = tmem_base + (uint32_t)(t * p.bn);
            uint32_t acc = acc0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint64_t xd = desc_join(x_lo + kstep_x * k, hi_x);
              const uint64_t gd = desc_join(g_lo0 + 128u * k, hi_g);
              if (p.x_is_m) umma_bf16(d_addr, xd, gd, idesc, acc);
              else umma_bf16(d_addr, gd, xd, idesc, acc);
              acc = 1u;
            }
          }
        }
        umma_commit(x_empty(xs));
        umma_commit(g_empty(gs));
        if (tile == tile1 - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
      if (++xs == p.stages_x) { xs = 0; xph ^= 1u; }
      if (++gs == p.stages_g) { gs = 0; gph ^= 1u; }
    }
  } else if (tile1 > tile0) {
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    const bool m_ok = m < p.m_dim;
    mbar_wait(tmem_full_bar, 0u);
    tc_fence_after();
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int t = 0; t < ntaps; ++t) {
      const long long tap_off = (long long)(dy0 * p.kw + t) * p.st + (long long)m * p.sm;
      for (int c0 = 0; c0 < p.bn; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(t_row + (uint32_t)(t * p.bn + c0), r);
        tmem_ld_wait();
        if (!m_ok) continue;
        if (p.sn == 1) {
          // a lane's 16 accumulator columns are 16 consecutive floats of gw (M = output channels): four 16-byte vector
          // reductions instead of 16 scalar ones — the scalar form made this epilogue the kernel's bottleneck (60 % of the
          // warp samples in profiles/r01_wgrad_256to1024_k1_80x80_n8.txt).  n_dim, n0, c0 are multiples of 8: aligned.
          float* dst = gw + tap_off + (long long)(n0 + c0);
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            if (n0 + c0 + j < p.n_dim) red_add_v4(dst + j, r[j], r[j + 1], r[j + 2], r[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int nn = n0 + c0 + j;
            if (nn < p.n_dim) atomicAdd(gw + tap_off + (long long)nn * p.sn, __uint_as_float(r[j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

static int launch_colsum(const cgb_conv_desc* d, const void* gy, float* gbias, cudaStream_t st);

// returns CGB_UNSUPPORTED when the halo variant does not apply (caller falls back to the per-tap kernel)
static int try_wgrad_halo(const cgb_conv_desc* d, const void* x, const void* gy, float* gw, cudaStream_t st) {
  const int taps = d->kh * d->kw;
  if (d->stride != 1 || taps == 1 || d->wo < 24 || d->ho < 32) return CGB_UNSUPPORTED;
  WgHaloParams p;
  memset(&p, 0, sizeof(p));
  p.n = d->n; p.ho = d->ho; p.wo = d->wo; p.kh = d->kh; p.kw = d->kw; p.dil = d->dil; p.pad = d->pad;
  p.x_is_m = d->ci >= d->co ? 1 : 0;
  p.m_dim = p.x_is_m ? d->ci : d->co;
  p.n_dim = p.x_is_m ? d->co : d->ci;
  p.bn = pick_bn(p.n_dim);
  // all-taps-resident N tile: with kh*kw*bn <= 512 TMEM columns one CTA holds every tap's accumulator, the x halo tile is loaded
  // once per pixel tile and gy once per N tile — the streaming form re-reads x once per tap and is L2 -> SM bound at ~6 TB/s
  // (256->256 d2 @80^2: 88 us for 43 us of MMAs).  The price is a narrow MMA (N = 48: 44 cycles instead of 24 per K step).
  static const int halo_bn = getenv("CGB_WG_HALO_BN") ? atoi(getenv("CGB_WG_HALO_BN")) : 0;
  if (halo_bn >= 16 && halo_bn % 16 == 0 && p.kh * p.kw * p.bn > 512 && p.kh * p.kw * halo_bn <= 512 && p.n_dim > halo_bn)
    p.bn = halo_bn;
  if (p.kw * p.bn > 512) return CGB_UNSUPPORTED;
  p.n_boxes = (p.bn + 63) / 64;
  p.rows_per_group = 512 / (p.kw * p.bn);
  if (p.rows_per_group > p.kh) p.rows_per_group = p.kh;
  p.row_groups = (p.kh + p.rows_per_group - 1) / p.rows_per_group;
  p.rows_per_group = (p.kh + p.row_groups - 1) / p.row_groups;
  p.twh = 8 + d->dil * (d->kw - 1);
  p.thh = 16 + d->dil * (p.rows_per_group - 1);
  if (p.twh > 256 || p.thh > 256) return CGB_UNSUPPORTED;
  p.x_box_bytes = (p.twh * p.thh * 128 + 1023) / 1024 * 1024;
  const int x_boxes = p.x_is_m ? 2 : p.n_boxes, g_boxes = p.x_is_m ? p.n_boxes : 2;
  const size_t xs = (size_t)x_boxes * p.x_box_bytes, gs = (size_t)g_boxes * BOX_BYTES;
  if (2 * (xs + gs) + 2048 > SMEM_LIMIT) return CGB_UNSUPPORTED;
  int stages = (int)((SMEM_LIMIT - 2048) / (xs + gs));
  if (stages > 4) stages = 4;
  p.stages_x = p.stages_g = stages;
  int cols = 32;
  while (cols < p.rows_per_group * p.kw * p.bn) cols <<= 1;
  p.tmem_cols = cols;
  if (p.x_is_m) { p.sm = 1; p.sn = (long long)taps * d->ci; } else { p.sm = (long long)taps * d->ci; p.sn = 1; }
  p.st = d->ci;
  p.tiles_x = (d->wo + 7) / 8; p.tiles_y = (d->ho + 15) / 16;
  p.total_tiles = p.tiles_x * p.tiles_y * d->n;
  const int m_tiles = (p.m_dim + 127) / 128;
  const int n_tiles = (p.n_dim + p.bn - 1) / p.bn;
  const int zdim = n_tiles * p.row_groups;
  int splits = num_sms() / (m_tiles * zdim);
  if (splits < 1) splits = 1;
  if (splits > p.total_tiles) splits = p.total_tiles;
  p.tiles_per_cta = (p.total_tiles + splits - 1) / splits;
  splits = (p.total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;

  CUtensorMap tmX, tmG;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->ci, (cuuint64_t)d->wi, (cuuint64_t)d->hi, (cuuint64_t)d->n};
    cuuint64_t strides[3] = {(cuuint64_t)d->ci * 2, (cuuint64_t)d->wi * d->ci * 2, (cuuint64_t)d->hi * d->wi * d->ci * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)p.twh, (cuuint32_t)p.thh, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (!encode_map(&tmX, x, 4, dims, strides, box, estr, "wgrad halo x")) return CGB_LAUNCH_FAILURE;
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->co, (cuuint64_t)d->wo, (cuuint64_t)d->ho, (cuuint64_t)d->n};
    cuuint64_t strides[3] = {(cuuint64_t)d->co * 2, (cuuint64_t)d->wo * d->co * 2, (cuuint64_t)d->ho * d->wo * d->co * 2};
    cuuint32_t box[4] = {64, 8, 16, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (!encode_map(&tmG, gy, 4, dims, strides, box, estr, "wgrad halo gy")) return CGB_LAUNCH_FAILURE;
  }
  const size_t smem = (size_t)stages * (xs + gs) + 32 * stages + 16 + 1024;
  static std::once_flag attr_once;
  std::call_once(attr_once, [] {
    cudaFuncSetAttribute(wgrad_tc_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
  });
  dim3 grid((unsigned)splits, (unsigned)m_tiles, (unsigned)zdim);
  wgrad_tc_halo_kernel<<<grid, WG_THREADS, smem, st>>>(tmX, tmG, p, gw);
  return after_launch("wgrad_tc_halo");
}

// per-channel sum over pixels (bias gradient): x [pixels, c] bf16 -> out[c] += sum
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, long long pixels, int c, long long px_per_cta) {
  extern __shared__ float sm[];
  const int cv = c >> 3;
  const int lanes = 256 / cv;
  const int tid = threadIdx.x;
  const int lane = tid / cv, v = tid - lane * cv;
  for (int i = tid; i < c; i += 256) sm[i] = 0.f;
  __syncthreads();
  if (lane < lanes) {
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
    const long long p0 = (long long)blockIdx.x * px_per_cta;
    long long p1 = p0 + px_per_cta;
    if (p1 > pixels) p1 = pixels;
    for (long long px = p0 + lane; px < p1; px += lanes) {
      float f[8];
      Vec8<__nv_bfloat16>::load(x + px * c + v * 8, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += f[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&sm[v * 8 + j], s[j]);
  }
  __syncthreads();
  for (int i = tid; i < c; i += 256) atomicAdd(out + i, sm[i]);
}

int conv_tc_wgrad(const cgb_conv_desc* d, const void* x, const void* gy, float* gw, float* gbias, cudaStream_t st) {
  {
    const int hs = try_wgrad_halo(d, x, gy, gw, st);
    if (hs == CGB_OK) return gbias ? launch_colsum(d, gy, gbias, st) : CGB_OK;
    if (hs != CGB_UNSUPPORTED) return hs;
  }
  WgParams p;
  p.n = d->n; p.ho = d->ho; p.wo = d->wo;
  p.kh = d->kh; p.kw = d->kw; p.dil = d->dil; p.stride = d->stride; p.pad = d->pad;
  pick_tile(d->n, d->ho, d->wo, d->stride, &p.tw_log, &p.th_log);
  const int tn_log = 7 - p.tw_log - p.th_log;
  p.tiles_x = (d->wo + (1 << p.tw_log) - 1) >> p.tw_log;
  p.tiles_y = (d->ho + (1 << p.th_log) - 1) >> p.th_log;
  const int tiles_n = (d->n + (1 << tn_log) - 1) >> tn_log;
  p.total_tiles = p.tiles_x * p.tiles_y * tiles_n;
  const int taps = d->kh * d->kw;
  // M side = x only when the output channels cannot fill an M tile: with M = output channels a lane's 16 accumulator columns are
  // 16 consecutive floats of gw (sn == 1) and the final reduction uses 16-byte vector reds; with M = x it is 16 scalar reds per
  // chunk, and that reduction is a fixed ~20 us per CTA (CGB_WG_XM=1: the round-1 rule ci >= co)
  static const int wg_xm = getenv("CGB_WG_XM") ? atoi(getenv("CGB_WG_XM")) : 0;
  p.x_is_m = (d->ci >= d->co && (wg_xm || d->co < 128)) ? 1 : 0;
  p.m_dim = p.x_is_m ? d->ci : d->co;
  p.n_dim = p.x_is_m ? d->co : d->ci;
  p.bn = pick_bn(p.n_dim);
  p.m_boxes = 2;
  p.n_boxes = (p.bn + 63) / 64;
  if (p.x_is_m) { p.sm = 1; p.sn = (long long)taps * d->ci; } else { p.sm = (long long)taps * d->ci; p.sn = 1; }
  p.st = d->ci;
  int tpg = 512 / p.bn;
  if (tpg > taps) tpg = taps;
  p.tap_groups = (taps + tpg - 1) / tpg;
  p.taps_per_group = (taps + p.tap_groups - 1) / p.tap_groups;
  int cols = 32;
  while (cols < p.taps_per_group * p.bn) cols <<= 1;
  p.tmem_cols = cols;
  const int s_boxes = p.x_is_m ? p.m_boxes : p.n_boxes;
  const int u_boxes = p.x_is_m ? p.n_boxes : p.m_boxes;
  p.stages_u = 2;
  int budget = 200 * 1024 - p.stages_u * u_boxes * BOX_BYTES;
  p.stages_s = budget / (s_boxes * BOX_BYTES);
  if (p.stages_s > 6) p.stages_s = 6;
  if (p.stages_s < 2) {
    set_error("tcgen05 wgrad: tile does not fit shared memory (s_boxes=%d u_boxes=%d)", s_boxes, u_boxes);
    return CGB_UNSUPPORTED;
  }
  const int m_tiles = (p.m_dim + 127) / 128;
  const int n_tiles = (p.n_dim + p.bn - 1) / p.bn;
  const int zdim = n_tiles * p.tap_groups;
  // pixel splits.  One CTA per SM (the accumulators own most of TMEM), so CTAs run in waves, and every CTA pays a fixed cost on
  // top of its tiles: pipeline fill plus the fp32 reduction of its whole accumulator into gw (~8 tile times).  Round 1 always
  // launched ~2 waves; one full wave is 20-28 % faster wherever m_tiles * zdim divides 148 well (256->256 d2 @80^2: 93 -> 71 us,
  // 2048->512 1x1: 176 -> 126), but halves the machine where it does not (ASPP 2048->256: 80 slots) — pick the split count that
  // minimises waves * (tiles per CTA + fixed cost).  CGB_WG_WAVES=n forces the old rule with n waves.
  static const int wg_waves = getenv("CGB_WG_WAVES") ? atoi(getenv("CGB_WG_WAVES")) : 0;
  const int slots = m_tiles * zdim;
  int splits = 1;
  if (wg_waves >= 1) {
    splits = (148 * wg_waves + slots - 1) / slots;
    if (wg_waves == 1) splits = 148 / slots > 0 ? 148 / slots : 1;
  } else {
    const int fixed = 8;
    long long best = -1;
    const int smax = (148 * 3 + slots - 1) / slots;
    for (int sp = 1; sp <= smax && sp <= p.total_tiles; ++sp) {
      const long long waves = ((long long)slots * sp + 147) / 148;
      const long long cost = waves * ((p.total_tiles + sp - 1) / sp + fixed);
      if (best < 0 || cost < best) { best = cost; splits = sp; }
    }
  }
  if (splits > p.total_tiles) splits = p.total_tiles;
  if (splits < 1) splits = 1;
  p.tiles_per_cta = (p.total_tiles + splits - 1) / splits;
  splits = (p.total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;

  CUtensorMap tmX, tmG;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->ci, (cuuint64_t)d->wi, (cuuint64_t)d->hi, (cuuint64_t)d->n};
    cuuint64_t strides[3] = {(cuuint64_t)d->ci * 2, (cuuint64_t)d->wi * d->ci * 2, (cuuint64_t)d->hi * d->wi * d->ci * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)((1 << p.tw_log) * d->stride), (cuuint32_t)((1 << p.th_log) * d->stride),
                         (cuuint32_t)(1 << tn_log)};
    cuuint32_t estr[4] = {1, (cuuint32_t)d->stride, (cuuint32_t)d->stride, 1};
    if (!encode_map(&tmX, x, 4, dims, strides, box, estr, "wgrad x")) return CGB_LAUNCH_FAILURE;
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->co, (cuuint64_t)d->wo, (cuuint64_t)d->ho, (cuuint64_t)d->n};
    cuuint64_t strides[3] = {(cuuint64_t)d->co * 2, (cuuint64_t)d->wo * d->co * 2, (cuuint64_t)d->ho * d->wo * d->co * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(1 << p.tw_log), (cuuint32_t)(1 << p.th_log), (cuuint32_t)(1 << tn_log)};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (!encode_map(&tmG, gy, 4, dims, strides, box, estr, "wgrad gy")) return CGB_LAUNCH_FAILURE;
  }
  // ---- CTA-pair form: two neighbouring m tiles share one pass over the N-side operand (whole 64-channel boxes per half)
  // measured (profiles/r02_tc2_pair_kernel.txt): 512->512 d4 235 -> 220 us, ASPP 438 -> 401 us, 256->256 d2 91 -> 89 us, the 1x1
  // classes unchanged (+-1 %); L2 -> SM bytes of 256->256 d2: 498 -> 367 MB per launch (ncu l1tex__m_xbar2l1tex_read_bytes)
  static const int tc2w = getenv("CGB_TC2_WGRAD") ? atoi(getenv("CGB_TC2_WGRAD")) : 1;
  if (tc2w && m_tiles % 2 == 0 && p.bn % 128 == 0) {
    const int nh_boxes = p.n_boxes / 2;
    const int s2 = p.x_is_m ? p.m_boxes : nh_boxes, u2 = p.x_is_m ? nh_boxes : p.m_boxes;
    WgParams q = p;
    q.stages_u = 2;
    int budget2 = 200 * 1024 - q.stages_u * u2 * BOX_BYTES;
    q.stages_s = budget2 / (s2 * BOX_BYTES);
    if (q.stages_s > 8) q.stages_s = 8;
    if (q.stages_s >= 2) {
      const size_t smem2 = (size_t)q.stages_s * s2 * BOX_BYTES + (size_t)q.stages_u * u2 * BOX_BYTES + 16 * (q.stages_s + q.stages_u) + 16 + 1024;
      static std::once_flag attr2_once;
      std::call_once(attr2_once, [] {
        cudaFuncSetAttribute(wgrad_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      });
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3((unsigned)m_tiles, (unsigned)splits, (unsigned)zdim);
      cfg.blockDim = dim3(WG_THREADS);
      cfg.dynamicSmemBytes = smem2;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      cudaError_t e = cudaLaunchKernelEx(&cfg, wgrad_tc2_kernel, tmX, tmG, q, gw);
      if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("wgrad_tc2: cluster launch failed: %s", cudaGetErrorString(e));
        return CGB_LAUNCH_FAILURE;
      }
      int s2r = after_launch("wgrad_tc2");
      if (s2r) return s2r;
      return gbias ? launch_colsum(d, gy, gbias, st) : CGB_OK;
    }
  }
  const size_t smem = (size_t)p.stages_s * s_boxes * BOX_BYTES + (size_t)p.stages_u * u_boxes * BOX_BYTES +
                      16 * (p.stages_s + p.stages_u) + 16 + 1024;
  static std::once_flag attr_once;
  std::call_once(attr_once, [] {
    cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  dim3 grid((unsigned)splits, (unsigned)m_tiles, (unsigned)zdim);
  wgrad_tc_kernel<<<grid, WG_THREADS, smem, st>>>(tmX, tmG, p, gw);
  int s = after_launch("wgrad_tc");
  if (s) return s;
  return gbias ? launch_colsum(d, gy, gbias, st) : CGB_OK;
}

static int launch_colsum(const cgb_conv_desc* d, const void* gy, float* gbias, cudaStream_t st) {
  const long long pixels = (long long)d->n * d->ho * d->wo;
  int ctas = (int)((pixels + 1023) / 1024);
  if (ctas > 148 * 8) ctas = 148 * 8;
  if (ctas < 1) ctas = 1;
  const long long ppc = (pixels + ctas - 1) / ctas;
  colsum_kernel<<<ctas, 256, d->co * sizeof(float), st>>>((const __nv_bfloat16*)gy, gbias, pixels, d->co, ppc);
  return after_launch("colsum");
}

// wt[ci][taps-1-t][co] = w[co][t][ci]
template <typename T>
__global__ void pack_dgrad_weight_kernel(const T* __restrict__ w, T* __restrict__ wt, int co, int taps, int ci) {
  const long long total = (long long)co * taps * ci;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(i % co);
    const long long r = i / co;
    const int t = (int)(r % taps);
    const int c = (int)(r / taps);
    wt[i] = w[((long long)o * taps + (taps - 1 - t)) * ci + c];
  }
}

int pack_dgrad_weight(const cgb_conv_desc* d, const void* w, void* wt, cudaStream_t st) {
  const int taps = d->kh * d->kw;
  const long long total = (long long)d->co * taps * d->ci;
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (d->dtype == CGB_F32)
    pack_dgrad_weight_kernel<float><<<grid, 256, 0, st>>>((const float*)w, (float*)wt, d->co, taps, d->ci);
  else if (d->dtype == CGB_F16)
    pack_dgrad_weight_kernel<__half><<<grid, 256, 0, st>>>((const __half*)w, (__half*)wt, d->co, taps, d->ci);
  else
    pack_dgrad_weight_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)w, (__nv_bfloat16*)wt, d->co, taps, d->ci);
  return after_launch("pack_dgrad_weight");
}

}  // namespace cgb
