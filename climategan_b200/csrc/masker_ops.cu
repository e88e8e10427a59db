// Training-path kernels of the Masker: train-mode BatchNorm (+ReLU/LeakyReLU, + the
// bottleneck's residual add) forward/backward, max-pool / bilinear / reflect-pad adjoints, dropout, and the masker's
// scalar losses (cross-entropy, entropy maps, MinEnt, TV, BCE, ground intersection, MiDaS scale-invariant gradient
// matching loss).  All HBM-bound: NHWC kernels are 16-byte vectorised over channels; NCHW fp32 loss kernels are
// coalesced over pixels.  Reductions: registers -> shared atomics -> one fp64 atomic per channel per CTA.
#include <climits>
#include "common.cuh"

namespace cgb {

static inline int grid_for(long long work, int block = 256) {
  long long g = (work + block - 1) / block;
  static const long long cap = 148LL * (getenv("CGB_FLAT_CTAS") ? atoi(getenv("CGB_FLAT_CTAS")) : 16);
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

#define DISPATCH_T(dtype, ...)                          \
  if ((dtype) == CGB_F32) {                             \
    using T = float;                                    \
    __VA_ARGS__                                         \
  } else if ((dtype) == CGB_BF16) {                     \
    using T = __nv_bfloat16;                            \
    __VA_ARGS__                                         \
  } else if ((dtype) == CGB_F16) {                      \
    using T = __half;                                   \
    __VA_ARGS__                                         \
  } else {                                              \
    set_error("unknown dtype %d", (int)(dtype));        \
    return CGB_BAD_ARG;                                 \
  }

// ---------------------------------------------------------------------------------------------------
// Train-mode BatchNorm passes.  All three are FLAT grid-stride kernels over the tensor's 16-byte channel vectors.  The grid is
// sized so that (gridDim * 256) is a multiple of cv = c/8: a thread then meets the SAME 8 channels on every iteration.
// Per-channel coefficients live in a shared-memory table built once per CTA (registers hold only the vectors in flight:
// <= 40 registers, 6-8 CTAs per SM).  History, measured on B200 (scripts/bench_hbm_kernels.py, profiles/r02_hbm_kernels.txt):
//   round 1: 256-pixel chunks, one load in flight, 200 CTAs on the 8x80x80 maps            -> 0.35-0.45 of the HBM roofline
//   first flat version: 4 vectors in flight held as fp32 (126-173 registers, 1-2 CTAs/SM)  -> 0.35-0.57 fwd, 0.24-0.34 bwd:
//   occupancy, not instruction-level parallelism, is what hides the latency here (spade_mod_fwd, 2048 threads/SM: 0.87).
constexpr int BN_U = 2;
constexpr long long BN_BWD_CAP = 148LL * 3;   // bn_apply_bwd_kernel: __launch_bounds__(256, 3) -> one resident wave

static inline long long gcd_ll(long long a, long long b) { while (b) { long long t = a % b; a = b; b = t; } return a; }

// CTAs for a flat pass over total_vec vectors, a multiple of cv / gcd(cv, 256) so that a thread's channel vector is fixed
static inline int bn_grid(long long total_vec, int cv, long long cap = 148LL * 8) {
  long long g = (total_vec + 256LL * BN_U - 1) / (256LL * BN_U);
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  const long long m = cv / gcd_ll(cv, 256);
  g = (g + m - 1) / m * m;
  return (int)g;
}

template <typename T> struct Raw8;                      // the 8-channel vector as loaded (no conversion until it is used)
template <> struct Raw8<float> { float4 a, b; };
template <> struct Raw8<__nv_bfloat16> { uint4 a; };
template <> struct Raw8<__half> { uint4 a; };
template <typename T>
__device__ __forceinline__ Raw8<T> ldraw(const T* p) { return *reinterpret_cast<const Raw8<T>*>(p); }
template <typename T>
__device__ __forceinline__ void cvt8(const Raw8<T>& r, float (&v)[8]) { Vec8<T>::load(reinterpret_cast<const T*>(&r), v); }

// y = act(x*A + B (+ residual)), A = rstd*w, B = b - mean*A  (per channel)
template <typename T>
__global__ void __launch_bounds__(256, 6)
bn_apply_fwd_kernel(const T* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ weight, const float* __restrict__ bias, const T* __restrict__ residual,
                    T* __restrict__ y, long long total_vec, int cv, float neg) {
  extern __shared__ float tab[];   // A[c], B[c]
  const int c = cv * 8;
  for (int i = threadIdx.x; i < c; i += 256) {
    const float a = rstd[i] * (weight ? weight[i] : 1.f);
    tab[i] = a;
    tab[c + i] = (bias ? bias[i] : 0.f) - mean[i] * a;
  }
  __syncthreads();
  const long long stride = (long long)gridDim.x * 256;
  const long long i0 = (long long)blockIdx.x * 256 + threadIdx.x;
  const int v = (int)(i0 % cv);
  const float* Ap = tab + v * 8;
  const float* Bp = tab + c + v * 8;
  for (long long i = i0; i < total_vec; i += stride * BN_U) {
    Raw8<T> xr[BN_U], rr[BN_U];
#pragma unroll
    for (int u = 0; u < BN_U; ++u) {
      const long long k = i + u * stride;
      if (k < total_vec) {
        xr[u] = ldraw<T>(x + k * 8);
        if (residual) rr[u] = ldraw<T>(residual + k * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < BN_U; ++u) {
      const long long k = i + u * stride;
      if (k < total_vec) {
        float xv[8], o[8];
        cvt8<T>(xr[u], xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(xv[j], Ap[j], Bp[j]);
        if (residual) {
          float r[8];
          cvt8<T>(rr[u], r);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += r[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = o[j] > 0.f ? o[j] : o[j] * neg;
        Vec8<T>::store(y + k * 8, o);
      }
    }
  }
}

// backward part 1: gpre = gy * act'(y); per-CTA partial sums of gpre and gpre * xhat (shared-memory table, then one plain
// store per (CTA, channel): no global atomics — the fp64 fold over CTAs is bn_bwd_reduce_kernel)
template <typename T>
__global__ void __launch_bounds__(256, 3)
bn_apply_bwd_kernel(const T* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                    const T* __restrict__ y, const T* __restrict__ gy, const T* __restrict__ gy2, T* __restrict__ gpre,
                    float* __restrict__ partial, long long total_vec, int cv, float neg, int has_act) {
  extern __shared__ float sm[];  // sums [2][8][cv] (conflict-free fold), then rs[c], nm[c]
  const int c = cv * 8;
  float* coef = sm + 2 * c;
  for (int i = threadIdx.x; i < 2 * c; i += 256) sm[i] = 0.f;
  for (int i = threadIdx.x; i < c; i += 256) {
    const float r = rstd[i];
    coef[i] = r;
    coef[c + i] = -mean[i] * r;
  }
  __syncthreads();
  const long long stride = (long long)gridDim.x * 256;
  const long long i0 = (long long)blockIdx.x * 256 + threadIdx.x;
  const int v = (int)(i0 % cv);
  const float* rs = coef + v * 8;
  const float* nm = coef + c + v * 8;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  for (long long i = i0; i < total_vec; i += stride * BN_U) {
    Raw8<T> xr[BN_U], gr[BN_U], yr[BN_U], g2r[BN_U];
#pragma unroll
    for (int u = 0; u < BN_U; ++u) {
      const long long k = i + u * stride;
      if (k < total_vec) {
        xr[u] = ldraw<T>(x + k * 8);
        gr[u] = ldraw<T>(gy + k * 8);
        if (gy2) g2r[u] = ldraw<T>(gy2 + k * 8);
        if (has_act) yr[u] = ldraw<T>(y + k * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < BN_U; ++u) {
      const long long k = i + u * stride;
      if (k < total_vec) {
        float o[8], t[8];
        cvt8<T>(gr[u], o);
        if (gy2) {   // the output fed two consumers: their gradients are summed here (fp32) instead of by a separate pass
          cvt8<T>(g2r[u], t);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += t[j];
        }
        if (has_act) {
          cvt8<T>(yr[u], t);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = t[j] > 0.f ? o[j] : o[j] * neg;
        }
        cvt8<T>(xr[u], t);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s1[j] += o[j];
          s2[j] = fmaf(o[j], fmaf(t[j], rs[j], nm[j]), s2[j]);
        }
        Vec8<T>::store(gpre + k * 8, o);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {   // [moment][j][v]: the lanes of a warp (consecutive v) hit consecutive words
    atomicAdd(&sm[j * cv + v], s1[j]);
    atomicAdd(&sm[c + j * cv + v], s2[j]);
  }
  __syncthreads();
  float* out = partial + (long long)blockIdx.x * 2 * c;
  for (int i = threadIdx.x; i < 2 * c; i += 256) {
    const int m = i >= c ? 1 : 0, ch = i - m * c;
    out[i] = sm[m * c + (ch & 7) * cv + (ch >> 3)];
  }
}

// sums[c][2] (fp64) = sum over chunks of the partials
// (only c / 32 CTAs run here — 8 for a 256-channel layer — so the fold is latency-bound on its row loop: the backward pass
//  launches one resident wave, BN_BWD_CAP = 3 CTAs per SM -> <= 444 partial rows instead of 1184, and 32 row groups per CTA:
//  14 us -> see DESIGN.md)
__global__ void __launch_bounds__(1024)
bn_bwd_reduce_kernel(const float* __restrict__ partial, double* __restrict__ sums, int c, int chunks) {
  __shared__ double sS[32][33], sQ[32][33];   // block (32 channels, 32 chunk groups)
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ch = blockIdx.x * 32 + tx;
  double S = 0.0, Q = 0.0;
  if (ch < c) {
    for (int k = ty; k < chunks; k += 32) {
      S += (double)partial[(long long)k * 2 * c + ch];
      Q += (double)partial[(long long)k * 2 * c + c + ch];
    }
  }
  sS[ty][tx] = S;
  sQ[ty][tx] = Q;
  __syncthreads();
  if (ty == 0 && ch < c) {
#pragma unroll
    for (int k = 1; k < 32; ++k) { S += sS[k][tx]; Q += sQ[k][tx]; }
    sums[ch * 2 + 0] = S;
    sums[ch * 2 + 1] = Q;
  }
}

// backward part 2: gx = w*rstd*(gpre - s0/M - xhat*s1/M) = A*g + B*x + C
template <typename T>
__global__ void __launch_bounds__(256, 5)
bn_bwd_finalize_kernel(const T* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                       const float* __restrict__ weight, const double* __restrict__ sums, const T* __restrict__ gpre,
                       T* __restrict__ gx, long long total_vec, int cv, double inv_npix) {
  extern __shared__ float tab[];   // A[c], B[c], C[c]
  const int c = cv * 8;
  for (int ch = threadIdx.x; ch < c; ch += 256) {
    const float w = weight ? weight[ch] : 1.f;
    const float rs = rstd[ch], mu = mean[ch];
    const float m1 = (float)(sums[ch * 2 + 0] * inv_npix);
    const float m2 = (float)(sums[ch * 2 + 1] * inv_npix);
    tab[ch] = w * rs;
    tab[c + ch] = -w * rs * rs * m2;
    tab[2 * c + ch] = w * (rs * rs * m2 * mu - rs * m1);
  }
  __syncthreads();
  const long long stride = (long long)gridDim.x * 256;
  const long long i0 = (long long)blockIdx.x * 256 + threadIdx.x;
  const int v = (int)(i0 % cv);
  const float* A = tab + v * 8;
  const float* B = tab + c + v * 8;
  const float* Cc = tab + 2 * c + v * 8;
  for (long long i = i0; i < total_vec; i += stride * BN_U) {
    Raw8<T> xr[BN_U], gr[BN_U];
#pragma unroll
    for (int u = 0; u < BN_U; ++u) {
      const long long k = i + u * stride;
      if (k < total_vec) {
        xr[u] = ldraw<T>(x + k * 8);
        gr[u] = ldraw<T>(gpre + k * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < BN_U; ++u) {
      const long long k = i + u * stride;
      if (k < total_vec) {
        float xv[8], gv[8], o[8];
        cvt8<T>(xr[u], xv);
        cvt8<T>(gr[u], gv);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(A[j], gv[j], fmaf(B[j], xv[j], Cc[j]));
        Vec8<T>::store(gx + k * 8, o);
      }
    }
  }
}

// mean / rstd (and the running-statistics update) from per-CTA partial sums [rows][2][c] — the statistics a tcgen05 conv's
// epilogue accumulated for its own output (cgb_conv2d_fwd_stats).  block (32 channels, 16 row groups), fp64 fold.
__global__ void __launch_bounds__(512)
bn_finalize_partials_kernel(const float* __restrict__ partial, int rows, int c, int c_logical, double inv_npix, double count,
                            float eps, float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ rmean,
                            float* __restrict__ rvar, float momentum, long long* __restrict__ num_batches_tracked) {
  __shared__ double sS[16][33], sQ[16][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ch = blockIdx.x * 32 + tx;
  double S = 0.0, Q = 0.0;
  if (ch < c) {
    for (int k = ty; k < rows; k += 16) {
      S += (double)partial[(long long)k * 2 * c + ch];
      Q += (double)partial[(long long)k * 2 * c + c + ch];
    }
  }
  sS[ty][tx] = S;
  sQ[ty][tx] = Q;
  __syncthreads();
  if (ty == 0 && ch < c) {
#pragma unroll
    for (int k = 1; k < 16; ++k) { S += sS[k][tx]; Q += sQ[k][tx]; }
    const double m = S * inv_npix;
    double var = Q * inv_npix - m * m;
    if (var < 0.0) var = 0.0;
    mean[ch] = (float)m;
    rstd[ch] = (float)(1.0 / sqrt(var + (double)eps));
    if (rmean && ch < c_logical) {
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      rmean[ch] = (1.f - momentum) * rmean[ch] + momentum * (float)m;
      rvar[ch] = (1.f - momentum) * rvar[ch] + momentum * (float)unbiased;
    }
  }
  if (blockIdx.x == 0 && tx == 0 && ty == 0 && num_batches_tracked) num_batches_tracked[0] += 1;
}

// running_mean = (1-mom)*running_mean + mom*mean ; running_var likewise with the UNBIASED batch variance
__global__ void bn_update_running_kernel(const float* __restrict__ mean, const float* __restrict__ rstd,
                                         float* __restrict__ rmean, float* __restrict__ rvar, int c, double count,
                                         float momentum, float eps, long long* __restrict__ num_batches_tracked) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && num_batches_tracked) num_batches_tracked[0] += 1;
  if (i >= c) return;
  const double rs = (double)rstd[i];
  double var = 1.0 / (rs * rs) - (double)eps;
  if (var < 0.0) var = 0.0;
  const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
  rmean[i] = (1.f - momentum) * rmean[i] + momentum * mean[i];
  rvar[i] = (1.f - momentum) * rvar[i] + momentum * (float)unbiased;
}

// ---------------------------------------------------------------------------------------------------
// nn.MaxPool2d(3, stride=2, padding=0, ceil_mode=True) backward, gather form: an input pixel belongs to <= 2x2 windows;
// it receives a window's gradient when it is that window's first maximum in row-major scan order (ATen's rule).
template <typename T>
__global__ void __launch_bounds__(256)
maxpool3s2_ceil_bwd_kernel(const T* __restrict__ x, const T* __restrict__ gy, T* __restrict__ gx, long long total_vec,
                           int hi, int wi, int ho, int wo, int c, int pad = 0) {
  const int cv = c >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / cv;
    const int v = (int)(i - pix * cv);
    const int ix = (int)(pix % wi);
    const long long t = pix / wi;
    const int iy = (int)(t % hi);
    const long long img = t / hi;
    float acc[8], me[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    Vec8<T>::load(x + pix * c + v * 8, me);
    // windows (oy, ox) covering the pixel: oy*2 - pad <= iy <= oy*2 - pad + 2
    const int oy_a = max(0, (iy + pad - 1) / 2), oy_b = min(ho - 1, (iy + pad) / 2);
    const int ox_a = max(0, (ix + pad - 1) / 2), ox_b = min(wo - 1, (ix + pad) / 2);
    for (int oy = oy_a; oy <= oy_b; ++oy) {
      if (!(oy * 2 - pad <= iy && iy <= oy * 2 - pad + 2)) continue;
      for (int ox = ox_a; ox <= ox_b; ++ox) {
        if (!(ox * 2 - pad <= ix && ix <= ox * 2 - pad + 2)) continue;
        // is (iy, ix) the first maximum of window (oy, ox)?
        bool first[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) first[j] = true;
        for (int dy = 0; dy < 3; ++dy)
          for (int dx = 0; dx < 3; ++dx) {
            const int sy = oy * 2 + dy - pad, sx = ox * 2 + dx - pad;
            if (sy < 0 || sx < 0 || sy >= hi || sx >= wi || (sy == iy && sx == ix)) continue;
            float f[8];
            Vec8<T>::load(x + ((img * hi + sy) * wi + sx) * c + v * 8, f);
            const bool before = (sy < iy) || (sy == iy && sx < ix);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (before ? (f[j] >= me[j]) : (f[j] > me[j])) first[j] = false;
            }
          }
        float g[8];
        Vec8<T>::load(gy + ((img * ho + oy) * wo + ox) * c + v * 8, g);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (first[j]) acc[j] += g[j];
      }
    }
    Vec8<T>::store(gx + pix * c + v * 8, acc);
  }
}

// The same operator, tiled (16-bit storage types): the gather form above re-derives every window's maximum for every pixel it
// covers (2.25 windows x 8 neighbour loads x 8 compares per pixel on average: 0.8 ms per launch on the ResNet stem's 8 x 320 x 320
// x 64 map, 1.6 ms per train step for 0.24 GB of traffic).  Here a CTA owns 16 x 16 input pixels x 4 channel vectors: it stages
// the 21 x 21 input patch that the covering windows read, every window's first-maximum position is found ONCE (4 bits per channel,
// ATen's row-major tie rule) together with its gradient vector, and each input pixel then gathers from the <= 4 windows that
// cover it.  Same accumulation order as the gather form: bit-identical results.
constexpr int MPT = 16;            // input pixels per tile side
constexpr int MPW = 10;            // windows per tile side (covers pad 0 and 1)
constexpr int MPS = 2 * MPW + 1;   // staged input rows / columns
constexpr int MPV = 4;             // channel vectors (of 8) per CTA

template <typename T>
__global__ void __launch_bounds__(256)
maxpool3s2_bwd_tiled_kernel(const T* __restrict__ x, const T* __restrict__ gy, T* __restrict__ gx, int hi, int wi, int ho, int wo,
                            int c, int pad, int vgroups) {
  __shared__ uint4 sx[MPS * MPS * MPV];      // staged input, raw 16-byte vectors (28 KB)
  __shared__ uint4 sg[MPW * MPW * MPV];      // gradient vector of every window of the tile
  __shared__ uint32_t sidx[MPW * MPW * MPV]; // first-maximum position (0..8) of every window, 4 bits per channel; 15: no window
  const int cv = c >> 3;
  const int vg = blockIdx.z % vgroups;
  const long long img = blockIdx.z / vgroups;
  const int v0 = vg * MPV;
  const int iy0 = blockIdx.y * MPT, ix0 = blockIdx.x * MPT;
  // first window that can cover row iy0: 2*oy - pad + 2 >= iy0  (floor division, may be -1)
  const int oyb = (iy0 + pad - 2 >= 0) ? (iy0 + pad - 2) >> 1 : -1;
  const int oxb = (ix0 + pad - 2 >= 0) ? (ix0 + pad - 2) >> 1 : -1;
  const int sy0 = 2 * oyb - pad, sx0 = 2 * oxb - pad;   // input coordinates of the staged patch's corner
  const float ninf = __int_as_float(0xff800000);
  for (int i = threadIdx.x; i < MPS * MPS * MPV; i += 256) {
    const int v = i % MPV;
    const int q = i / MPV;
    const int cx = q % MPS, cy = q / MPS;
    const int iy = sy0 + cy, ix = sx0 + cx;
    uint4 r;
    if (iy >= 0 && iy < hi && ix >= 0 && ix < wi && v0 + v < cv) {
      r = *reinterpret_cast<const uint4*>(x + ((img * hi + iy) * wi + ix) * c + (v0 + v) * 8);
    } else {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = ninf;
      Vec8<T>::store(reinterpret_cast<T*>(&r), f);
    }
    sx[i] = r;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < MPW * MPW * MPV; i += 256) {
    const int v = i % MPV;
    const int q = i / MPV;
    const int wx = q % MPW, wy = q / MPW;
    const int oy = oyb + wy, ox = oxb + wx;
    uint32_t word = 0xffffffffu;
    uint4 g = make_uint4(0u, 0u, 0u, 0u);
    if (oy >= 0 && oy < ho && ox >= 0 && ox < wo && v0 + v < cv) {
      float best[8];
      int bi[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { best[j] = ninf; bi[j] = 15; }
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        float f[8];
        Vec8<T>::load(reinterpret_cast<const T*>(&sx[((2 * wy + k / 3) * MPS + (2 * wx + k % 3)) * MPV + v]), f);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (f[j] > best[j]) { best[j] = f[j]; bi[j] = k; }   // strict: the first maximum in row-major order wins
      }
      word = 0u;
#pragma unroll
      for (int j = 0; j < 8; ++j) word |= (uint32_t)bi[j] << (4 * j);
      g = *reinterpret_cast<const uint4*>(gy + ((img * ho + oy) * wo + ox) * c + (v0 + v) * 8);
    }
    sidx[i] = word;
    sg[i] = g;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < MPT * MPT * MPV; i += 256) {
    const int v = i % MPV;
    const int q = i / MPV;
    const int px = q % MPT, py = q / MPT;
    const int iy = iy0 + py, ix = ix0 + px;
    if (iy >= hi || ix >= wi || v0 + v >= cv) continue;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const int oy_a = max(0, (iy + pad - 1) / 2), oy_b = min(ho - 1, (iy + pad) / 2);
    const int ox_a = max(0, (ix + pad - 1) / 2), ox_b = min(wo - 1, (ix + pad) / 2);
    for (int oy = oy_a; oy <= oy_b; ++oy) {
      const int dy = iy - (oy * 2 - pad);
      if (dy < 0 || dy > 2) continue;
      for (int ox = ox_a; ox <= ox_b; ++ox) {
        const int dx = ix - (ox * 2 - pad);
        if (dx < 0 || dx > 2) continue;
        const int w = ((oy - oyb) * MPW + (ox - oxb)) * MPV + v;
        const uint32_t word = sidx[w];
        const uint32_t k = (uint32_t)(dy * 3 + dx);
        float g[8];
        Vec8<T>::load(reinterpret_cast<const T*>(&sg[w]), g);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (((word >> (4 * j)) & 15u) == k) acc[j] += g[j];
      }
    }
    Vec8<T>::store(gx + ((img * hi + iy) * wi + ix) * c + (v0 + v) * 8, acc);
  }
}

template <typename T>
static bool maxpool3s2_bwd_tiled(const T* x, const T* gy, T* gx, int n, int hi, int wi, int ho, int wo, int c, int pad, cudaStream_t st) {
  static const int on = getenv("CGB_MAXPOOL_TILED") ? atoi(getenv("CGB_MAXPOOL_TILED")) : 1;
  if constexpr (sizeof(T) == 2) {
    if (!on || hi < 32 || wi < 32) return false;
    const int vgroups = (c / 8 + MPV - 1) / MPV;
    if ((long long)n * vgroups > 65535) return false;
    dim3 grid((wi + MPT - 1) / MPT, (hi + MPT - 1) / MPT, n * vgroups);
    maxpool3s2_bwd_tiled_kernel<T><<<grid, 256, 0, st>>>(x, gy, gx, hi, wi, ho, wo, c, pad, vgroups);
    return true;
  } else {
    return false;   // fp32 storage (the parity mode) keeps the gather form
  }
}

// ---------------------------------------------------------------------------------------------------
// bilinear resize backward (adjoint of resize_bilinear_kernel in ops.cu), gather form: deterministic, no atomics.
__device__ __forceinline__ float bil_src(int o, float s, int ac) { return ac ? s * o : fmaxf(s * (o + 0.5f) - 0.5f, 0.f); }

template <typename T>
__global__ void __launch_bounds__(256)
resize_bilinear_bwd_kernel(const T* __restrict__ gy, T* __restrict__ gx, long long total_vec, int hi, int wi, int ho, int wo,
                           int c, int ac) {
  const int cv = c >> 3;
  const float sh = ac ? (ho > 1 ? (float)(hi - 1) / (float)(ho - 1) : 0.f) : (float)hi / (float)ho;
  const float sw = ac ? (wo > 1 ? (float)(wi - 1) / (float)(wo - 1) : 0.f) : (float)wi / (float)wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / cv;
    const int v = (int)(i - pix * cv);
    const int ix = (int)(pix % wi);
    const long long t = pix / wi;
    const int iy = (int)(t % hi);
    const long long img = t / hi;
    // candidate output rows/cols whose source falls in (iy-1, iy+1)
    int oy0, oy1, ox0, ox1;
    if (sh > 0.f) {
      oy0 = max(0, (int)floorf(((float)iy - 1.f) / sh) - 1);
      oy1 = min(ho - 1, (int)ceilf(((float)iy + 1.f) / sh) + 1);
    } else { oy0 = 0; oy1 = ho - 1; }
    if (sw > 0.f) {
      ox0 = max(0, (int)floorf(((float)ix - 1.f) / sw) - 1);
      ox1 = min(wo - 1, (int)ceilf(((float)ix + 1.f) / sw) + 1);
    } else { ox0 = 0; ox1 = wo - 1; }
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int oy = oy0; oy <= oy1; ++oy) {
      const float fy = bil_src(oy, sh, ac);
      const int y0 = min((int)fy, hi - 1), y1 = min(y0 + 1, hi - 1);
      const float ly = fy - (float)y0;
      float wy = 0.f;
      if (y0 == iy) wy += 1.f - ly;
      if (y1 == iy) wy += ly;
      if (wy == 0.f) continue;
      for (int ox = ox0; ox <= ox1; ++ox) {
        const float fx = bil_src(ox, sw, ac);
        const int x0 = min((int)fx, wi - 1), x1 = min(x0 + 1, wi - 1);
        const float lx = fx - (float)x0;
        float wx = 0.f;
        if (x0 == ix) wx += 1.f - lx;
        if (x1 == ix) wx += lx;
        if (wx == 0.f) continue;
        float g[8];
        Vec8<T>::load(gy + ((img * ho + oy) * wo + ox) * c + v * 8, g);
        const float wgt = wy * wx;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(g[j], wgt, acc[j]);
      }
    }
    Vec8<T>::store(gx + pix * c + v * 8, acc);
  }
}

// ---------------------------------------------------------------------------------------------------
// nn.ReflectionPad2d(pad) forward (copy) and backward (fold), NHWC
template <typename T>
__global__ void __launch_bounds__(256)
reflect_pad_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long total_vec, int h, int w, int c, int pad) {
  const int cv = c >> 3;
  const int hp = h + 2 * pad, wp = w + 2 * pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / cv;
    const int v = (int)(i - pix * cv);
    const int ox = (int)(pix % wp);
    const long long t = pix / wp;
    const int oy = (int)(t % hp);
    const long long img = t / hp;
    const int sy = reflect_idx(oy - pad, h), sx = reflect_idx(ox - pad, w);
    const uint4* src = reinterpret_cast<const uint4*>(x + ((img * h + sy) * w + sx) * c + v * 8);
    uint4* dst = reinterpret_cast<uint4*>(y + pix * c + v * 8);
    dst[0] = src[0];
    if (sizeof(T) == 4) dst[1] = src[1];
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
reflect_pad_bwd_kernel(const T* __restrict__ gy, T* __restrict__ gx, long long total_vec, int h, int w, int c, int pad) {
  const int cv = c >> 3;
  const int hp = h + 2 * pad, wp = w + 2 * pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / cv;
    const int v = (int)(i - pix * cv);
    const int ix = (int)(pix % w);
    const long long t = pix / w;
    const int iy = (int)(t % h);
    const long long img = t / h;
    int ys[3], xs[3], ny = 0, nx = 0;
    ys[ny++] = iy + pad;
    if (iy >= 1 && iy <= pad) ys[ny++] = pad - iy;
    if (iy <= h - 2 && iy >= h - 1 - pad) ys[ny++] = pad + 2 * (h - 1) - iy;
    xs[nx++] = ix + pad;
    if (ix >= 1 && ix <= pad) xs[nx++] = pad - ix;
    if (ix <= w - 2 && ix >= w - 1 - pad) xs[nx++] = pad + 2 * (w - 1) - ix;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int a = 0; a < ny; ++a)
      for (int b = 0; b < nx; ++b) {
        float g[8];
        Vec8<T>::load(gy + ((img * hp + ys[a]) * wp + xs[b]) * c + v * 8, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += g[j];
      }
    Vec8<T>::store(gx + pix * c + v * 8, acc);
  }
}

// nn.ReplicationPad2d(pad), NHWC: forward is a clamped copy, backward a gather over the padding cells that replicate a border pixel
template <typename T>
__global__ void __launch_bounds__(256)
replicate_pad_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long total_vec, int h, int w, int c, int pad) {
  const int cv = c >> 3;
  const int hp = h + 2 * pad, wp = w + 2 * pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / cv;
    const int v = (int)(i - pix * cv);
    const int ox = (int)(pix % wp);
    const long long t = pix / wp;
    const int oy = (int)(t % hp);
    const long long img = t / hp;
    const int sy = min(max(oy - pad, 0), h - 1), sx = min(max(ox - pad, 0), w - 1);
    float f[8];
    Vec8<T>::load(x + ((img * h + sy) * w + sx) * c + v * 8, f);
    Vec8<T>::store(y + pix * c + v * 8, f);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
replicate_pad_bwd_kernel(const T* __restrict__ gy, T* __restrict__ gx, long long total_vec, int h, int w, int c, int pad) {
  const int cv = c >> 3;
  const int wp = w + 2 * pad, hp = h + 2 * pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / cv;
    const int v = (int)(i - pix * cv);
    const int ix = (int)(pix % w);
    const long long t = pix / w;
    const int iy = (int)(t % h);
    const long long img = t / h;
    // padded rows / columns that read this pixel
    const int y0 = iy == 0 ? 0 : iy + pad, y1 = iy == h - 1 ? hp - 1 : iy + pad;
    const int x0 = ix == 0 ? 0 : ix + pad, x1 = ix == w - 1 ? wp - 1 : ix + pad;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int yy = y0; yy <= y1; ++yy)
      for (int xx = x0; xx <= x1; ++xx) {
        float g[8];
        Vec8<T>::load(gy + ((img * hp + yy) * wp + xx) * c + v * 8, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += g[j];
      }
    Vec8<T>::store(gx + pix * c + v * 8, acc);
  }
}

// per-(sample, channel) affine + activation (see include/cgb200.h); grid (chunks, n): a thread owns one channel vector
template <typename T>
__global__ void __launch_bounds__(256)
affine_nc_fwd_kernel(const T* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift, T* __restrict__ y,
                     int hw, int c, int px_per_chunk, int act, float slope) {
  const int cv = c >> 3;
  const int lanes = 256 / cv;
  const int lane = threadIdx.x / cv, v = threadIdx.x - lane * cv;
  if (lane >= lanes) return;
  const int img = blockIdx.y;
  float A[8], B[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    A[j] = scale[(long long)img * c + v * 8 + j];
    B[j] = shift[(long long)img * c + v * 8 + j];
  }
  const int p0 = blockIdx.x * px_per_chunk;
  const int p1 = min(hw, p0 + px_per_chunk);
  const T* xb = x + ((long long)img * hw) * c + v * 8;
  T* yb = y + ((long long)img * hw) * c + v * 8;
  for (int p = p0 + lane; p < p1; p += lanes) {
    float f[8];
    Vec8<T>::load(xb + (long long)p * c, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = act_apply(fmaf(f[j], A[j], B[j]), act, slope);
    Vec8<T>::store(yb + (long long)p * c, f);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
affine_nc_bwd_kernel(const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ gy, const float* __restrict__ scale,
                     T* __restrict__ gx, double* __restrict__ sums, int hw, int c, int px_per_chunk, int act, float slope) {
  extern __shared__ float sm[];   // [2][c]
  const int cv = c >> 3;
  const int lanes = 256 / cv;
  const int lane = threadIdx.x / cv, v = threadIdx.x - lane * cv;
  const int img = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * c; i += 256) sm[i] = 0.f;
  __syncthreads();
  if (lane < lanes) {
    float A[8], s0[8], s1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      A[j] = scale[(long long)img * c + v * 8 + j];
      s0[j] = s1[j] = 0.f;
    }
    const int p0 = blockIdx.x * px_per_chunk;
    const int p1 = min(hw, p0 + px_per_chunk);
    const long long base = ((long long)img * hw) * c + v * 8;
    for (int p = p0 + lane; p < p1; p += lanes) {
      float xv[8], g[8], o[8];
      Vec8<T>::load(x + base + (long long)p * c, xv);
      Vec8<T>::load(gy + base + (long long)p * c, g);
      if (act != CGB_ACT_NONE) {
        float yv[8];
        Vec8<T>::load(y + base + (long long)p * c, yv);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] *= act_grad_from_out(yv[j], act, slope);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s0[j] += g[j];
        s1[j] = fmaf(g[j], xv[j], s1[j]);
        o[j] = g[j] * A[j];
      }
      Vec8<T>::store(gx + base + (long long)p * c, o);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&sm[v * 8 + j], s0[j]);
      atomicAdd(&sm[c + v * 8 + j], s1[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < c; i += 256) {
    atomicAdd(&sums[((long long)img * c + i) * 2 + 0], (double)sm[i]);
    atomicAdd(&sums[((long long)img * c + i) * 2 + 1], (double)sm[c + i]);
  }
}

// ---------------------------------------------------------------------------------------------------
// torch.mean(z, dim=1, keepdim=True) backward: gx[p, ch<c_logical] = gy[p,0]/c_logical
template <typename T>
__global__ void __launch_bounds__(256)
channel_mean_bwd_kernel(const T* __restrict__ gy, T* __restrict__ gx, long long total_vec, int cs, int c_logical) {
  const int cv = cs >> 3;
  const float inv = 1.f / (float)c_logical;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / cv;
    const int v = (int)(i - pix * cv);
    const float g = to_f<T>(gy[pix * 8]) * inv;
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (v * 8 + j) < c_logical ? g : 0.f;
    Vec8<T>::store(gx + pix * cs + v * 8, o);
  }
}

// dst[n,hw,c] = src[n,c] * scale   (GAP backward; bilinear upsampling of a 1x1 map)
template <typename T>
__global__ void __launch_bounds__(256)
broadcast_hw_kernel(const T* __restrict__ src, T* __restrict__ dst, long long total_vec, int hw, int c, float scale) {
  const int cv = c >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / cv;
    const int v = (int)(i - pix * cv);
    const long long img = pix / hw;
    float f[8];
    Vec8<T>::load(src + img * c + v * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= scale;
    Vec8<T>::store(dst + pix * c + v * 8, f);
  }
}

// ---------------------------------------------------------------------------------------------------
// dropout: y = x * keep / (1-p), keep from a counter-based hash of (seed, element index) — the same call with the same
// seed applied to the gradient is the backward.
__device__ __forceinline__ uint32_t hash32(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return (uint32_t)k;
}

template <typename T>
__global__ void __launch_bounds__(256)
dropout_kernel(const T* __restrict__ x, T* __restrict__ y, long long total_vec, float p, uint64_t seed,
               const unsigned long long* __restrict__ seed_dev) {
  if (seed_dev) seed = *seed_dev;   // this step's seed, read from device memory (captured CUDA graphs)
  const float scale = 1.f / (1.f - p);
  const uint32_t thr = (uint32_t)((double)p * 4294967296.0);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    Vec8<T>::load(x + i * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t r = hash32(seed * 0x9E3779B97F4A7C15ULL + (uint64_t)(i * 8 + j));
      f[j] = r >= thr ? f[j] * scale : 0.f;
    }
    Vec8<T>::store(y + i * 8, f);
  }
}

// =====================================================================================================
// NCHW fp32 loss kernels (thread = pixel; channel loop strided by hw -> coalesced across the warp)
// =====================================================================================================
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum -> one atomicAdd (fp32 target) per CTA
__device__ __forceinline__ void block_accumulate(float v, float* target, float scale) {
  __shared__ float red[32];
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = l < (blockDim.x >> 5) ? red[l] : 0.f;
    t = warp_sum(t);
    if (l == 0) atomicAdd(target, t * scale);
  }
  __syncthreads();
}
__device__ __forceinline__ void block_accumulate_d(double v, double* target) {
  __shared__ double redd[32];
  v = warp_sum_d(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) redd[w] = v;
  __syncthreads();
  if (w == 0) {
    double t = l < (blockDim.x >> 5) ? redd[l] : 0.0;
    t = warp_sum_d(t);
    if (l == 0) atomicAdd(target, t);
  }
  __syncthreads();
}

// softmax over dim 1
__global__ void __launch_bounds__(256)
softmax_nchw_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long npix, int c, int hw) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
    const long long img = p / hw;
    const long long base = img * c * hw + (p - img * hw);
    float mx = -3.4e38f;
    for (int k = 0; k < c; ++k) mx = fmaxf(mx, x[base + (long long)k * hw]);
    float s = 0.f;
    for (int k = 0; k < c; ++k) s += expf(x[base + (long long)k * hw] - mx);
    const float inv = 1.f / s;
    for (int k = 0; k < c; ++k) y[base + (long long)k * hw] = expf(x[base + (long long)k * hw] - mx) * inv;
  }
}
// gx = y * (gy - sum_k gy_k y_k)
__global__ void __launch_bounds__(256)
softmax_nchw_bwd_kernel(const float* __restrict__ y, const float* __restrict__ gy, float* __restrict__ gx, long long npix, int c,
                        int hw) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
    const long long img = p / hw;
    const long long base = img * c * hw + (p - img * hw);
    float dot = 0.f;
    for (int k = 0; k < c; ++k) dot = fmaf(gy[base + (long long)k * hw], y[base + (long long)k * hw], dot);
    for (int k = 0; k < c; ++k) {
      const long long o = base + (long long)k * hw;
      gx[o] = y[o] * (gy[o] - dot);
    }
  }
}

// nn.CrossEntropyLoss()(logits [n,c,h,w], target int64 [n,h,w]) mean over pixels, fused with its gradient
__global__ void __launch_bounds__(256)
cross_entropy_nchw_kernel(const float* __restrict__ x, const long long* __restrict__ target, float* __restrict__ loss,
                          float* __restrict__ gx, long long npix, int c, int hw, float scale) {
  float local = 0.f;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
    const long long img = p / hw;
    const long long base = img * c * hw + (p - img * hw);
    const int t = (int)target[p];
    float mx = -3.4e38f;
    for (int k = 0; k < c; ++k) mx = fmaxf(mx, x[base + (long long)k * hw]);
    float s = 0.f;
    for (int k = 0; k < c; ++k) s += expf(x[base + (long long)k * hw] - mx);
    const float lse = logf(s) + mx;
    local += lse - x[base + (long long)t * hw];
    if (gx) {
      const float inv = 1.f / s;
      for (int k = 0; k < c; ++k) {
        const long long o = base + (long long)k * hw;
        gx[o] = scale * (expf(x[o] - mx) * inv - (k == t ? 1.f : 0.f));
      }
    }
  }
  block_accumulate(local, loss, scale);
}

// prob_2_entropy (losses.py:466-471) [* depth broadcast over channels]: e = -p*log2(p+1e-30)/log2(c) * d
__device__ __forceinline__ float ent_f(float p, float inv_log2c) { return -p * log2f(p + 1e-30f) * inv_log2c; }
__device__ __forceinline__ float ent_df(float p, float inv_log2c) {
  return -(log2f(p + 1e-30f) + p / ((p + 1e-30f) * 0.6931471805599453f)) * inv_log2c;
}
__global__ void __launch_bounds__(256)
entropy_nchw_kernel(const float* __restrict__ p, const float* __restrict__ depth, const float* __restrict__ ge,
                    float* __restrict__ out, long long total, int c, int hw, float inv_log2c, int bwd) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long img = i / ((long long)c * hw);
    const long long px = i % hw;
    const float d = depth ? depth[img * hw + px] : 1.f;
    out[i] = bwd ? ge[i] * d * ent_df(p[i], inv_log2c) : d * ent_f(p[i], inv_log2c);
  }
}

// MinentLoss (losses.py:177-196).  pass 0: acc[0] += sum E (fp64).  pass 1 (after pass 0): loss and gradient.
//   v1: L = sum(E)/M,  M = n*h*w.          dL/dE = 1/M
//   v2: mu = sum(E)/M ; L = sum(E + lam*(E-mu)^2)/M ; dL/dE_i = (1 + 2 lam (E_i-mu))/M - 2 lam (S - N mu)/M^2,  S = sum E, N = numel
__global__ void __launch_bounds__(256)
minent_sum_kernel(const float* __restrict__ p, double* __restrict__ acc, long long total, float inv_log2c) {
  double local = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    local += (double)ent_f(p[i], inv_log2c);
  block_accumulate_d(local, acc);
}
__global__ void __launch_bounds__(256)
minent_final_kernel(const float* __restrict__ p, const double* __restrict__ acc, float* __restrict__ loss,
                    float* __restrict__ gp, long long total, double M, float inv_log2c, int version, float lam) {
  const double S = acc[0];
  const float mu = (float)(S / M);
  const float invM = (float)(1.0 / M);
  const float cterm = version == 2 ? (float)(2.0 * lam * (S - (double)total * (S / M)) / (M * M)) : 0.f;
  float local = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float e = ent_f(p[i], inv_log2c);
    float dLdE;
    if (version == 2) {
      local += e + lam * (e - mu) * (e - mu);
      dLdE = (1.f + 2.f * lam * (e - mu)) * invM - cterm;
    } else {
      local += e;
      dLdE = invM;
    }
    if (gp) gp[i] = dLdE * ent_df(p[i], inv_log2c);
  }
  block_accumulate(local, loss, invM);
}

// sigmoid(logits [n,1,hw]) -> prob [n,2,hw] = cat[p, 1-p]  (trainer.py:1532-1534) ; bwd: gl = (g0 - g1) p (1-p)
__global__ void __launch_bounds__(256)
sigmoid_pair_kernel(const float* __restrict__ logits, const float* __restrict__ gprob, float* __restrict__ out, long long npix,
                    int hw, int bwd) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
    const long long img = p / hw, px = p - img * hw;
    const float s = 1.f / (1.f + expf(-logits[p]));
    if (!bwd) {
      out[(img * 2 + 0) * hw + px] = s;
      out[(img * 2 + 1) * hw + px] = 1.f - s;
    } else {
      out[p] = (gprob[(img * 2 + 0) * hw + px] - gprob[(img * 2 + 1) * hw + px]) * s * (1.f - s);
    }
  }
}

// TVLoss (losses.py:140-171) on x [n,c,h,w]: 2*(sum dh^2/count_h + sum dw^2/count_w)/n, fused with its gradient
__global__ void __launch_bounds__(256)
tv_loss_kernel(const float* __restrict__ x, float* __restrict__ loss, float* __restrict__ gx, long long total, int h, int w,
               float sh, float sw) {
  float local = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % w);
    const int yy = (int)((i / w) % h);
    const float v = x[i];
    float g = 0.f;
    if (yy + 1 < h) { const float d = x[i + w] - v; local += sh * d * d; g -= 2.f * sh * d; }
    if (yy > 0) { const float d = v - x[i - w]; g += 2.f * sh * d; }
    if (xx + 1 < w) { const float d = x[i + 1] - v; local += sw * d * d; g -= 2.f * sw * d; }
    if (xx > 0) { const float d = v - x[i - 1]; g += 2.f * sw * d; }
    if (gx) gx[i] = g;
  }
  block_accumulate(local, loss, 1.f);
}

// nn.BCEWithLogitsLoss()(x, t) mean, fused with gradient: l = max(x,0) - x t + log(1+exp(-|x|)) ; dl/dx = sigmoid(x) - t
__global__ void __launch_bounds__(256)
bce_logits_kernel(const float* __restrict__ x, const float* __restrict__ t, float* __restrict__ loss, float* __restrict__ gx,
                  long long total, float scale) {
  float local = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i], tt = t[i];
    local += fmaxf(v, 0.f) - v * tt + log1pf(expf(-fabsf(v)));
    if (gx) gx[i] = scale * (1.f / (1.f + expf(-v)) - tt);
  }
  block_accumulate(local, loss, scale);
}

// GroundIntersectionLoss (losses.py:449-455): mean(1.0 * ((pseudo_ground - pred) > 0.5)) — piecewise constant, no gradient
__global__ void __launch_bounds__(256)
ground_intersection_kernel(const float* __restrict__ pred, const float* __restrict__ ground, float* __restrict__ loss,
                           long long total, float scale) {
  float local = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    local += (ground[i] - pred[i]) > 0.5f ? 1.f : 0.f;
  block_accumulate(local, loss, scale);
}

// ---------------------------------------------------------------------------------------------------
// SIGMLoss (losses.py:232-278; MiDaS scale-and-shift-invariant loss with Sobel gradient matching)
// stats[0]=median t (lower median, torch.median), [1]=s=mean|x-t|, [2]=#elements equal to t, [3]=sum sign(x-t)
__device__ __forceinline__ uint32_t f2key(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__global__ void __launch_bounds__(1024)
median_stats_kernel(const float* __restrict__ x, long long n, float* __restrict__ stats) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned int s_prefix, s_k;
  __shared__ float redf[64];
  const int tid = threadIdx.x;
  if (tid == 0) { s_prefix = 0; s_k = (unsigned int)((n - 1) / 2); }
  unsigned int mask = 0;
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    const unsigned int prefix = s_prefix;
    for (long long i = tid; i < n; i += 1024) {
      const uint32_t k = f2key(x[i]);
      if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 0xff], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned int kk = s_k, b = 0;
      for (; b < 256; ++b) {
        if (kk < hist[b]) break;
        kk -= hist[b];
      }
      s_k = kk;
      s_prefix = prefix | (b << shift);
    }
    mask |= 0xffu << shift;
    __syncthreads();
  }
  const float t = key2f(s_prefix);
  float sabs = 0.f, cnt = 0.f, ssign = 0.f;
  for (long long i = tid; i < n; i += 1024) {
    const float d = x[i] - t;
    sabs += fabsf(d);
    cnt += (x[i] == t) ? 1.f : 0.f;
    ssign += d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
  }
  float vals[3] = {sabs, cnt, ssign};
  for (int q = 0; q < 3; ++q) {
    float v = warp_sum(vals[q]);
    if ((tid & 31) == 0) redf[tid >> 5] = v;
    __syncthreads();
    if (tid < 32) {
      v = warp_sum(redf[tid]);
      if (tid == 0) vals[q] = v;
    }
    __syncthreads();
  }
  if (tid == 0) {
    stats[0] = t;
    stats[1] = vals[0] / (float)n;
    stats[2] = vals[1];
    stats[3] = vals[2];
  }
}

// R = (x-tp)/sp - (y-tt)/st ; G = coef*sign(R) ; loss += coef*sum|R|
__global__ void __launch_bounds__(256)
sigm_residual_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ sp,
                     const float* __restrict__ st, float* __restrict__ R, float* __restrict__ G, float* __restrict__ loss,
                     long long total, float coef) {
  const float tp = sp[0], ip = 1.f / sp[1], tt = st[0], it = 1.f / st[1];
  float local = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float r = (x[i] - tp) * ip - (y[i] - tt) * it;
    R[i] = r;
    local += fabsf(r);
    G[i] = coef * (r > 0.f ? 1.f : (r < 0.f ? -1.f : 0.f));
  }
  block_accumulate(local, loss, coef);
}

// one scale of the gradient-matching term: R_ = R[::f, ::f] (nearest, scale 1/f), valid 3x3 Sobel x/y
__global__ void __launch_bounds__(256)
sigm_sobel_kernel(const float* __restrict__ R, float* __restrict__ G, float* __restrict__ loss, int n, int h, int w, int f,
                  float coef) {
  const int hk = h / f, wk = w / f;
  const int oh = hk - 2, ow = wk - 2;
  const long long total = (long long)n * oh * ow;
  const float kx[9] = {1.f, 0.f, -1.f, 2.f, 0.f, -2.f, 1.f, 0.f, -1.f};
  const float ky[9] = {1.f, 2.f, 1.f, 0.f, 0.f, 0.f, -1.f, -2.f, -1.f};
  float local = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % ow);
    const int oy = (int)((i / ow) % oh);
    const long long img = i / ((long long)ow * oh);
    const float* base = R + img * h * w;
    float rx = 0.f, ry = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const float v = base[(long long)(oy + a) * f * w + (ox + b) * f];
        rx = fmaf(kx[a * 3 + b], v, rx);
        ry = fmaf(ky[a * 3 + b], v, ry);
      }
    local += fabsf(rx) + fabsf(ry);
    const float sx = rx > 0.f ? coef : (rx < 0.f ? -coef : 0.f);
    const float sy = ry > 0.f ? coef : (ry < 0.f ? -coef : 0.f);
    float* gb = G + img * h * w;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const float g = sx * kx[a * 3 + b] + sy * ky[a * 3 + b];
        if (g != 0.f) atomicAdd(&gb[(long long)(oy + a) * f * w + (ox + b) * f], g);
      }
  }
  block_accumulate(local, loss, coef);
}

// acc[0] += sum G ; acc[1] += sum G*p, p = (x-t)/s
__global__ void __launch_bounds__(256)
sigm_reduce_kernel(const float* __restrict__ x, const float* __restrict__ G, const float* __restrict__ sp,
                   double* __restrict__ acc, long long total) {
  const float tp = sp[0], ip = 1.f / sp[1];
  double a = 0.0, b = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    a += (double)G[i];
    b += (double)(G[i] * (x[i] - tp) * ip);
  }
  block_accumulate_d(a, acc);
  block_accumulate_d(b, acc + 1);
}

// gx_i = G_i/s - M_i*SG/s - (SGp/s) * (sign(x_i-t) - M_i*Ssign)/N,  M_i = [x_i==t]/cnt  (median: gradient shared among ties)
__global__ void __launch_bounds__(256)
sigm_grad_kernel(const float* __restrict__ x, const float* __restrict__ G, const float* __restrict__ sp,
                 const double* __restrict__ acc, float* __restrict__ gx, long long total) {
  const float t = sp[0], is = 1.f / sp[1], cnt = sp[2], ssign = sp[3];
  const float SG = (float)acc[0], SGp = (float)acc[1];
  const float invN = 1.f / (float)total;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float d = x[i] - t;
    const float M = (x[i] == t) ? 1.f / cnt : 0.f;
    const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    gx[i] = G[i] * is - M * SG * is - SGp * is * (sg - M * ssign) * invN;
  }
}

}  // namespace cgb

// =====================================================================================================
// C ABI
// =====================================================================================================
using namespace cgb;

#define REQ_C(c, what) CGB_REQUIRE((c) % 8 == 0 && (c) >= 8 && (c) <= 2048, what ": c=%d must be a multiple of 8 in [8,2048]", (int)(c))

extern "C" int cgb_bn_apply_fwd(const void* x, const float* mean, const float* rstd, const float* weight, const float* bias,
                                const void* residual, void* y, int32_t dtype, int64_t npix, int32_t c, int32_t act, float slope,
                                void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && mean && rstd && y && npix > 0, "bn_apply_fwd: bad arguments");
  CGB_REQUIRE(act == CGB_ACT_NONE || act == CGB_ACT_RELU || act == CGB_ACT_LRELU, "bn_apply_fwd: act must be none/relu/lrelu");
  REQ_C(c, "bn_apply_fwd");
  const float neg = act == CGB_ACT_NONE ? 1.f : (act == CGB_ACT_RELU ? 0.f : slope);
  const int cv = c / 8;
  const long long total = (long long)npix * cv;
  // CTAs per SM, measured (profiles/r02_bn_grid_size.txt; 8 was the default until the end of round 2): one resident wave or less beats
  // 1184 CTAs in 1.3-2 waves — plain pass 0.60 -> 0.75 of the HBM roof at 6 per SM, with a residual 0.70 -> 0.84 at 4
  static const int fwd_ctas = getenv("CGB_BN_FWD_CTAS") ? atoi(getenv("CGB_BN_FWD_CTAS")) : 0;
  const long long fwd_cap = 148LL * (fwd_ctas > 0 ? fwd_ctas : (residual ? 4 : 6));
  DISPATCH_T(dtype, bn_apply_fwd_kernel<T><<<bn_grid(total, cv, fwd_cap), 256, 2 * c * sizeof(float), (cudaStream_t)stream>>>(
                        (const T*)x, mean, rstd, weight, bias, (const T*)residual, (T*)y, total, cv, neg);)
  return after_launch("bn_apply_fwd");
}

extern "C" int64_t cgb_bn_bwd_ws_doubles(int64_t npix, int32_t c) {
  if (npix <= 0 || c <= 0) return 0;
  const int cv = c / 8;
  return 2 * (int64_t)c + (int64_t)bn_grid((long long)npix * cv, cv, BN_BWD_CAP) * c;   // sums[c][2] doubles, then grid * 2c fp32 partials
}

static int bn_apply_bwd_impl(const void* x, const float* mean, const float* rstd, const void* y, const void* gy, const void* gy2,
                             void* gpre, double* sums, int32_t dtype, int64_t npix, int32_t c, int32_t act, float slope, void* stream);

extern "C" int cgb_bn_apply_bwd(const void* x, const float* mean, const float* rstd, const void* y, const void* gy, void* gpre,
                                double* sums, int32_t dtype, int64_t npix, int32_t c, int32_t act, float slope, void* stream) {
  return bn_apply_bwd_impl(x, mean, rstd, y, gy, nullptr, gpre, sums, dtype, npix, c, act, slope, stream);
}

static int bn_apply_bwd_impl(const void* x, const float* mean, const float* rstd, const void* y, const void* gy, const void* gy2,
                             void* gpre, double* sums, int32_t dtype, int64_t npix, int32_t c, int32_t act, float slope, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && mean && rstd && gy && gpre && sums && npix > 0, "bn_apply_bwd: bad arguments");
  CGB_REQUIRE(act == CGB_ACT_NONE || y, "bn_apply_bwd: y is required when an activation is fused");
  REQ_C(c, "bn_apply_bwd");
  const float neg = act == CGB_ACT_NONE ? 1.f : (act == CGB_ACT_RELU ? 0.f : slope);
  cudaStream_t st = (cudaStream_t)stream;
  const int cv = c / 8;
  const long long total = (long long)npix * cv;
  const int grid = bn_grid(total, cv, BN_BWD_CAP);
  float* partial = reinterpret_cast<float*>(sums + 2 * (size_t)c);   // sums holds cgb_bn_bwd_ws_doubles(npix, c) doubles
  DISPATCH_T(dtype, bn_apply_bwd_kernel<T><<<grid, 256, 4 * c * sizeof(float), st>>>(
                        (const T*)x, mean, rstd, (const T*)y, (const T*)gy, (const T*)gy2, (T*)gpre, partial, total, cv, neg,
                        act != CGB_ACT_NONE);)
  int s = after_launch("bn_apply_bwd");
  if (s) return s;
  bn_bwd_reduce_kernel<<<(c + 31) / 32, dim3(32, 32), 0, st>>>(partial, sums, c, grid);
  return after_launch("bn_bwd_reduce");
}

extern "C" int cgb_bn_bwd_finalize(const void* x, const float* mean, const float* rstd, const float* weight, const double* sums,
                                   const void* gpre, void* gx, int32_t dtype, int64_t npix, int32_t c, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && mean && rstd && sums && gpre && gx && npix > 0, "bn_bwd_finalize: bad arguments");
  REQ_C(c, "bn_bwd_finalize");
  const int cv = c / 8;
  const long long total = (long long)npix * cv;
  static const long long fin_cap = 148LL * (getenv("CGB_BN_FIN_CTAS") ? atoi(getenv("CGB_BN_FIN_CTAS")) : 4);   // 0.70 -> 0.82 of the HBM roof (8 until the end of round 2)
  DISPATCH_T(dtype, bn_bwd_finalize_kernel<T><<<bn_grid(total, cv, fin_cap), 256, 3 * c * sizeof(float), (cudaStream_t)stream>>>(
                        (const T*)x, mean, rstd, weight, sums, (const T*)gpre, (T*)gx, total, cv, 1.0 / (double)npix);)
  return after_launch("bn_bwd_finalize");
}

extern "C" int cgb_bn_update_running(const float* mean, const float* rstd, float* running_mean, float* running_var, int32_t c,
                                     int64_t count, float momentum, float eps, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(mean && rstd && running_mean && running_var && c > 0, "bn_update_running: bad arguments");
  bn_update_running_kernel<<<(c + 255) / 256, 256, 0, (cudaStream_t)stream>>>(mean, rstd, running_mean, running_var, c,
                                                                             (double)count, momentum, eps, nullptr);
  return after_launch("bn_update_running");
}

// One call for the whole train-mode forward: statistics -> running-stat update (+ num_batches_tracked) -> normalise/affine/
// residual/activation.  Saves three Python->C round trips per BatchNorm (464 BatchNorm forwards per train step).
extern "C" int cgb_instnorm_stats(const void* x, int32_t dtype, int32_t n, int32_t hw, int32_t c, float eps, double* ws,
                                  float* mean, float* rstd, void* stream);
extern "C" int cgb_bn_train_fwd(const void* x, const float* weight, const float* bias, const void* residual, void* y, float* mean,
                                float* rstd, double* ws, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                                int32_t dtype, int64_t npix, int32_t c, int32_t c_logical, float momentum, float eps, int32_t act,
                                float slope, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(npix > 0 && npix < (1ll << 31), "bn_train_fwd: npix=%lld out of range", (long long)npix);
  int s = cgb_instnorm_stats(x, dtype, 1, (int32_t)npix, c, eps, ws, mean, rstd, stream);
  if (s) return s;
  if (running_mean && running_var) {
    bn_update_running_kernel<<<(c_logical + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        mean, rstd, running_mean, running_var, c_logical, (double)npix, momentum, eps, (long long*)num_batches_tracked);
    if ((s = after_launch("bn_update_running"))) return s;
  }
  return cgb_bn_apply_fwd(x, mean, rstd, weight, bias, residual, y, dtype, npix, c, act, slope, stream);
}

// Train-mode forward whose statistics were accumulated by the producing conv's epilogue (cgb_conv2d_fwd_stats):
// partial [rows][2][c] fp32 -> mean / rstd (+ running statistics, num_batches_tracked) -> normalise/affine/residual/activation.
// One pass over x instead of two.
extern "C" int cgb_bn_train_fwd_partials(const void* x, const float* partial, int32_t rows, const float* weight, const float* bias,
                                         const void* residual, void* y, float* mean, float* rstd, float* running_mean,
                                         float* running_var, int64_t* num_batches_tracked, int32_t dtype, int64_t npix, int32_t c,
                                         int32_t c_logical, float momentum, float eps, int32_t act, float slope, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && partial && y && mean && rstd && rows > 0 && npix > 0, "bn_train_fwd_partials: bad arguments");
  REQ_C(c, "bn_train_fwd_partials");
  const bool upd = running_mean && running_var;
  bn_finalize_partials_kernel<<<(c + 31) / 32, dim3(32, 16), 0, (cudaStream_t)stream>>>(
      partial, rows, c, c_logical, 1.0 / (double)npix, (double)npix, eps, mean, rstd, upd ? running_mean : nullptr,
      upd ? running_var : nullptr, momentum, upd ? (long long*)num_batches_tracked : nullptr);
  int s = after_launch("bn_finalize_partials");
  if (s) return s;
  return cgb_bn_apply_fwd(x, mean, rstd, weight, bias, residual, y, dtype, npix, c, act, slope, stream);
}

// backward in one call: part 1 (activation mask, sums, gpre) then part 2 (gx) when the input needs a gradient
extern "C" int cgb_bn_train_bwd(const void* x, const float* mean, const float* rstd, const float* weight, const void* y,
                                const void* gy, void* gpre, void* gx, double* sums, int32_t dtype, int64_t npix, int32_t c,
                                int32_t act, float slope, void* stream) {
  int s = cgb_bn_apply_bwd(x, mean, rstd, y, gy, gpre, sums, dtype, npix, c, act, slope, stream);
  if (s || !gx) return s;
  return cgb_bn_bwd_finalize(x, mean, rstd, weight, sums, gpre, gx, dtype, npix, c, stream);
}

extern "C" int cgb_bn_train_bwd2(const void* x, const float* mean, const float* rstd, const float* weight, const void* y,
                                 const void* gy, const void* gy2, void* gpre, void* gx, double* sums, int32_t dtype, int64_t npix,
                                 int32_t c, int32_t act, float slope, void* stream) {
  int s = bn_apply_bwd_impl(x, mean, rstd, y, gy, gy2, gpre, sums, dtype, npix, c, act, slope, stream);
  if (s || !gx) return s;
  return cgb_bn_bwd_finalize(x, mean, rstd, weight, sums, gpre, gx, dtype, npix, c, stream);
}

extern "C" int cgb_maxpool3s2_ceil_bwd(const void* x, const void* gy, void* gx, int32_t dtype, int32_t n, int32_t hi, int32_t wi,
                                       int32_t ho, int32_t wo, int32_t c, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && gy && gx && c % 8 == 0 && c >= 8, "maxpool3s2_ceil_bwd: bad arguments");
  const long long total = (long long)n * hi * wi * (c / 8);
  DISPATCH_T(dtype, if (!maxpool3s2_bwd_tiled<T>((const T*)x, (const T*)gy, (T*)gx, n, hi, wi, ho, wo, c, 0, (cudaStream_t)stream))
                        maxpool3s2_ceil_bwd_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(
                            (const T*)x, (const T*)gy, (T*)gx, total, hi, wi, ho, wo, c);)
  return after_launch("maxpool3s2_ceil_bwd");
}

extern "C" int cgb_maxpool3s2_bwd(const void* x, const void* gy, void* gx, int32_t dtype, int32_t n, int32_t hi, int32_t wi,
                                  int32_t ho, int32_t wo, int32_t c, int32_t pad, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && gy && gx && c % 8 == 0 && c >= 8 && (pad == 0 || pad == 1), "maxpool3s2_bwd: bad arguments");
  const long long total = (long long)n * hi * wi * (c / 8);
  DISPATCH_T(dtype, if (!maxpool3s2_bwd_tiled<T>((const T*)x, (const T*)gy, (T*)gx, n, hi, wi, ho, wo, c, pad, (cudaStream_t)stream))
                        maxpool3s2_ceil_bwd_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(
                            (const T*)x, (const T*)gy, (T*)gx, total, hi, wi, ho, wo, c, pad);)
  return after_launch("maxpool3s2_bwd");
}

extern "C" int cgb_resize_bilinear_bwd(const void* gy, void* gx, int32_t dtype, int32_t n, int32_t hi, int32_t wi, int32_t ho,
                                       int32_t wo, int32_t c, int32_t align_corners, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(gy && gx && c % 8 == 0 && c >= 8, "resize_bilinear_bwd: bad arguments");
  const long long total = (long long)n * hi * wi * (c / 8);
  DISPATCH_T(dtype, resize_bilinear_bwd_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(
                        (const T*)gy, (T*)gx, total, hi, wi, ho, wo, c, align_corners);)
  return after_launch("resize_bilinear_bwd");
}

extern "C" int cgb_reflect_pad_fwd(const void* x, void* y, int32_t dtype, int32_t n, int32_t h, int32_t w, int32_t c, int32_t pad,
                                   void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && y && c % 8 == 0 && c >= 8, "reflect_pad_fwd: bad arguments");
  CGB_REQUIRE(pad >= 0 && pad < h && pad < w, "reflect_pad_fwd: pad=%d must be smaller than the map (%dx%d)", pad, h, w);
  const long long total = (long long)n * (h + 2 * pad) * (w + 2 * pad) * (c / 8);
  DISPATCH_T(dtype, reflect_pad_fwd_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, total, h, w,
                                                                                                c, pad);)
  return after_launch("reflect_pad_fwd");
}

extern "C" int cgb_reflect_pad_bwd(const void* gy, void* gx, int32_t dtype, int32_t n, int32_t h, int32_t w, int32_t c,
                                   int32_t pad, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(gy && gx && c % 8 == 0 && c >= 8, "reflect_pad_bwd: bad arguments");
  CGB_REQUIRE(pad >= 0 && pad < h && pad < w, "reflect_pad_bwd: pad=%d must be smaller than the map (%dx%d)", pad, h, w);
  const long long total = (long long)n * h * w * (c / 8);
  DISPATCH_T(dtype, reflect_pad_bwd_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const T*)gy, (T*)gx, total, h,
                                                                                                w, c, pad);)
  return after_launch("reflect_pad_bwd");
}

extern "C" int cgb_replicate_pad_fwd(const void* x, void* y, int32_t dtype, int32_t n, int32_t h, int32_t w, int32_t c, int32_t pad,
                                     void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && y && c % 8 == 0 && c >= 8 && pad >= 0 && h > 0 && w > 0, "replicate_pad_fwd: bad arguments");
  const long long total = (long long)n * (h + 2 * pad) * (w + 2 * pad) * (c / 8);
  DISPATCH_T(dtype, replicate_pad_fwd_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, total, h, w, c, pad);)
  return after_launch("replicate_pad_fwd");
}

extern "C" int cgb_replicate_pad_bwd(const void* gy, void* gx, int32_t dtype, int32_t n, int32_t h, int32_t w, int32_t c,
                                     int32_t pad, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(gy && gx && c % 8 == 0 && c >= 8 && pad >= 0 && h > 0 && w > 0, "replicate_pad_bwd: bad arguments");
  const long long total = (long long)n * h * w * (c / 8);
  DISPATCH_T(dtype, replicate_pad_bwd_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const T*)gy, (T*)gx, total, h, w, c, pad);)
  return after_launch("replicate_pad_bwd");
}

static inline int affine_chunks(int n, int hw) {
  int want = (148 * 8 + n - 1) / n;
  const int maxc = (hw + 63) / 64;
  if (want > maxc) want = maxc;
  return want < 1 ? 1 : want;
}

extern "C" int cgb_affine_nc_fwd(const void* x, const float* scale, const float* shift, void* y, int32_t dtype, int32_t n,
                                 int32_t hw, int32_t c, int32_t act, float slope, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && scale && shift && y && n > 0 && hw > 0 && n <= 65535, "affine_nc_fwd: bad arguments");
  REQ_C(c, "affine_nc_fwd");
  CGB_REQUIRE(act >= CGB_ACT_NONE && act <= CGB_ACT_SELU, "affine_nc_fwd: unknown activation %d", act);
  const int chunks = affine_chunks(n, hw);
  const int ppc = (hw + chunks - 1) / chunks;
  dim3 grid((hw + ppc - 1) / ppc, n);
  DISPATCH_T(dtype, affine_nc_fwd_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)x, scale, shift, (T*)y, hw, c, ppc, act, slope);)
  return after_launch("affine_nc_fwd");
}

extern "C" int cgb_affine_nc_bwd(const void* x, const void* y, const void* gy, const float* scale, void* gx, double* sums,
                                 int32_t dtype, int32_t n, int32_t hw, int32_t c, int32_t act, float slope, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && gy && scale && gx && sums && n > 0 && hw > 0 && n <= 65535, "affine_nc_bwd: bad arguments");
  CGB_REQUIRE(act == CGB_ACT_NONE || y, "affine_nc_bwd: y is required when an activation is fused");
  REQ_C(c, "affine_nc_bwd");
  const int chunks = affine_chunks(n, hw);
  const int ppc = (hw + chunks - 1) / chunks;
  dim3 grid((hw + ppc - 1) / ppc, n);
  DISPATCH_T(dtype, affine_nc_bwd_kernel<T><<<grid, 256, 2 * c * sizeof(float), (cudaStream_t)stream>>>(
                        (const T*)x, (const T*)y, (const T*)gy, scale, (T*)gx, sums, hw, c, ppc, act, slope);)
  return after_launch("affine_nc_bwd");
}

extern "C" int cgb_channel_mean_bwd(const void* gy, void* gx, int32_t dtype, int64_t pixels, int32_t cs, int32_t c_logical,
                                    void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(gy && gx && cs % 8 == 0 && cs >= 8 && c_logical >= 1 && c_logical <= cs, "channel_mean_bwd: bad arguments");
  const long long total = pixels * (cs / 8);
  DISPATCH_T(dtype, channel_mean_bwd_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const T*)gy, (T*)gx, total, cs,
                                                                                                 c_logical);)
  return after_launch("channel_mean_bwd");
}

extern "C" int cgb_broadcast_hw(const void* src, void* dst, int32_t dtype, int32_t n, int32_t hw, int32_t c, float scale,
                                void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(src && dst && c % 8 == 0 && c >= 8, "broadcast_hw: bad arguments");
  const long long total = (long long)n * hw * (c / 8);
  DISPATCH_T(dtype, broadcast_hw_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const T*)src, (T*)dst, total, hw, c,
                                                                                             scale);)
  return after_launch("broadcast_hw");
}

extern "C" int cgb_dropout(const void* x, void* y, int32_t dtype, int64_t count, float p, uint64_t seed, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && y && count % 8 == 0 && p >= 0.f && p < 1.f, "dropout: bad arguments");
  const long long total = count / 8;
  DISPATCH_T(dtype, dropout_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, total, p, seed, nullptr);)
  return after_launch("dropout");
}

extern "C" int cgb_dropout_dev(const void* x, void* y, int32_t dtype, int64_t count, float p, const uint64_t* seed_dev, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && y && seed_dev && count % 8 == 0 && p >= 0.f && p < 1.f, "dropout_dev: bad arguments");
  const long long total = count / 8;
  DISPATCH_T(dtype, dropout_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(
                        (const T*)x, (T*)y, total, p, 0ULL, (const unsigned long long*)seed_dev);)
  return after_launch("dropout_dev");
}

extern "C" int cgb_softmax_nchw_fwd(const float* x, float* y, int32_t n, int32_t c, int32_t hw, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && y && n > 0 && c > 0 && hw > 0, "softmax_nchw_fwd: bad arguments");
  const long long npix = (long long)n * hw;
  softmax_nchw_fwd_kernel<<<grid_for(npix), 256, 0, (cudaStream_t)stream>>>(x, y, npix, c, hw);
  return after_launch("softmax_nchw_fwd");
}

extern "C" int cgb_softmax_nchw_bwd(const float* y, const float* gy, float* gx, int32_t n, int32_t c, int32_t hw, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(y && gy && gx && n > 0 && c > 0 && hw > 0, "softmax_nchw_bwd: bad arguments");
  const long long npix = (long long)n * hw;
  softmax_nchw_bwd_kernel<<<grid_for(npix), 256, 0, (cudaStream_t)stream>>>(y, gy, gx, npix, c, hw);
  return after_launch("softmax_nchw_bwd");
}

extern "C" int cgb_cross_entropy_nchw(const float* logits, const int64_t* target, float* loss, float* glogits, int32_t n, int32_t c,
                                      int32_t hw, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(logits && target && loss && n > 0 && c > 0 && hw > 0, "cross_entropy_nchw: bad arguments");
  const long long npix = (long long)n * hw;
  cross_entropy_nchw_kernel<<<grid_for(npix), 256, 0, (cudaStream_t)stream>>>(logits, (const long long*)target, loss, glogits, npix,
                                                                             c, hw, 1.f / (float)npix);
  return after_launch("cross_entropy_nchw");
}

// ---------------------------------------------------------------------------------------------------
// DADADepthLoss (losses.py:596-620): reverse Huber (berHu) on |pred - label| with the threshold c = 0.2 * max over the WHOLE
// batch — taken with .item() in the reference, so c is a constant of the graph:
//   loss = ( sum_{a<=c} a + sum_{a>c} (a^2 + c^2) / (2c) ) / count ;  dl/dpred = sign(d)/count (a <= c) | d/(c*count) (a > c)
// Depth maps are small (n x 160 x 160 at 640^2): one CTA, two passes over the data.
__global__ void __launch_bounds__(1024)
dada_depth_loss_kernel(const float* __restrict__ pred, const float* __restrict__ label, float* __restrict__ loss,
                       float* __restrict__ gpred, long long count) {
  __shared__ float sh[32];
  __shared__ float s_c;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float mx = 0.f;
  for (long long i = threadIdx.x; i < count; i += blockDim.x) mx = fmaxf(mx, fabsf(pred[i] - label[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) sh[warp] = mx;
  __syncthreads();
  if (warp == 0) {
    mx = lane < nw ? sh[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) s_c = 0.2f * mx;
  }
  __syncthreads();
  const float c = s_c;
  const float inv_n = 1.f / (float)count;
  float s = 0.f;
  for (long long i = threadIdx.x; i < count; i += blockDim.x) {
    const float d = pred[i] - label[i];
    const float a = fabsf(d);
    float g;
    if (a <= c) {
      s += a;
      g = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    } else {
      s += (a * a + c * c) / (2.f * c);
      g = d / c;
    }
    if (gpred) gpred[i] = g * inv_n;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __syncthreads();
  if (lane == 0) sh[warp] = s;
  __syncthreads();
  if (warp == 0) {
    s = lane < nw ? sh[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) atomicAdd(loss, s * inv_n);
  }
}

extern "C" int cgb_dada_depth_loss(const float* pred, const float* label, float* loss, float* gpred, int64_t count, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(pred && label && loss && count > 0, "dada_depth_loss: bad arguments");
  dada_depth_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(pred, label, loss, gpred, (long long)count);
  return after_launch("dada_depth_loss");
}

// ---------------------------------------------------------------------------------------------------
// Validation metrics (Trainer.eval_images trainer.py:1706-1799; accuracy / mIOU eval_metrics.py:68-130): both are functions of
// the confusion matrix of argmax(pred, dim=1) against the integer label, so one pass over the logits replaces the reference's
// .cpu() round trip and its Python loop over classes (two masked reductions + .item() per class).
//   conf[p][l] += 1 per pixel, p = argmax_c logits[n][c][pix] (first maximum; a NaN counts as the maximum: torch.argmax),
//   l = label if 0 <= label < c, else the extra column c (ignore index: counted in the prediction totals, never a match).
// HBM-bound: c floats + one int64 per pixel, class-major reads coalesced across the pixels of a warp.  Counts go to a
// shared-memory histogram (c (c+1) 32-bit cells, warp-aggregated with match.any) and are flushed once per CTA with 64-bit
// atomics.
__global__ void __launch_bounds__(256)
argmax_confusion_kernel(const float* __restrict__ logits, const int64_t* __restrict__ label, unsigned long long* __restrict__ conf,
                        long long* __restrict__ label_max, int c, long long hw, long long total) {
  extern __shared__ unsigned int sh_conf[];
  const int cells = c * (c + 1);
  for (int i = threadIdx.x; i < cells; i += blockDim.x) sh_conf[i] = 0u;
  __syncthreads();
  long long lmax = LLONG_MIN;
  const int lane = threadIdx.x & 31;
  // block-uniform trip count, so the warp-level match below always runs with the full mask
  for (long long base = (long long)blockIdx.x * blockDim.x; base < total; base += (long long)gridDim.x * blockDim.x) {
    const long long idx = base + threadIdx.x;
    int cell = -1;
    if (idx < total) {
      const long long n = idx / hw;
      const float* p = logits + n * c * hw + (idx - n * hw);
      float best = p[0];
      int bi = 0;
      for (int k = 1; k < c; ++k) {
        const float v = p[(long long)k * hw];
        if (best == best && (v > best || v != v)) {
          best = v;
          bi = k;
        }
      }
      const long long l = label[idx];
      lmax = l > lmax ? l : lmax;
      cell = bi * (c + 1) + ((l >= 0 && l < c) ? (int)l : c);
    }
    // neighbouring pixels mostly share (prediction, label): one shared-memory atomic per distinct cell of the warp
    const unsigned peers = __match_any_sync(0xffffffffu, cell);
    if (cell >= 0 && lane == __ffs(peers) - 1) atomicAdd(&sh_conf[cell], (unsigned)__popc(peers));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long long other = __shfl_xor_sync(0xffffffffu, lmax, o);
    lmax = other > lmax ? other : lmax;
  }
  if (lane == 0 && lmax != LLONG_MIN) atomicMax(label_max, lmax);
  __syncthreads();
  for (int i = threadIdx.x; i < cells; i += blockDim.x)
    if (sh_conf[i]) atomicAdd(&conf[i], (unsigned long long)sh_conf[i]);
}

extern "C" int cgb_argmax_confusion(const float* logits, const int64_t* label, int64_t* conf, int64_t* label_max, int32_t n, int32_t c,
                                    int64_t hw, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(logits && label && conf && label_max && n > 0 && hw > 0, "argmax_confusion: bad arguments");
  CGB_REQUIRE(c >= 1 && c <= 64, "argmax_confusion: 1 <= classes <= 64, got %d", c);
  const long long total = (long long)n * hw;
  // (a CTA counts at most total / gridDim.x + 256 pixels: 32-bit cells cannot wrap below 2^32 pixels per CTA)
  long long g = (total + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  argmax_confusion_kernel<<<(int)g, 256, (size_t)c * (c + 1) * sizeof(unsigned int), (cudaStream_t)stream>>>(
      logits, label, (unsigned long long*)conf, (long long*)label_max, c, (long long)hw, total);
  return after_launch("argmax_confusion");
}

extern "C" int cgb_entropy_nchw(const float* p, const float* depth, const float* ge, float* out, int32_t n, int32_t c, int32_t hw,
                                int32_t backward, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(p && out && n > 0 && c > 1 && hw > 0 && (!backward || ge), "entropy_nchw: bad arguments");
  const long long total = (long long)n * c * hw;
  entropy_nchw_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(p, depth, ge, out, total, c, hw,
                                                                        1.f / log2f((float)c), backward);
  return after_launch("entropy_nchw");
}

extern "C" int cgb_minent_loss(const float* p, float* loss, float* gp, double* acc, int32_t n, int32_t c, int32_t hw, int32_t version,
                               float lambda_var, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(p && loss && acc && n > 0 && c > 1 && hw > 0 && (version == 1 || version == 2), "minent_loss: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)n * c * hw;
  const float il = 1.f / log2f((float)c);
  cudaMemsetAsync(acc, 0, sizeof(double), st);
  minent_sum_kernel<<<grid_for(total), 256, 0, st>>>(p, acc, total, il);
  int s = after_launch("minent_sum");
  if (s) return s;
  minent_final_kernel<<<grid_for(total), 256, 0, st>>>(p, acc, loss, gp, total, (double)n * hw, il, version, lambda_var);
  return after_launch("minent_final");
}

extern "C" int cgb_sigmoid_pair(const float* logits, const float* gprob, float* out, int32_t n, int32_t hw, int32_t backward,
                                void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(logits && out && n > 0 && hw > 0 && (!backward || gprob), "sigmoid_pair: bad arguments");
  const long long npix = (long long)n * hw;
  sigmoid_pair_kernel<<<grid_for(npix), 256, 0, (cudaStream_t)stream>>>(logits, gprob, out, npix, hw, backward);
  return after_launch("sigmoid_pair");
}

extern "C" int cgb_tv_loss(const float* x, float* loss, float* gx, int32_t n, int32_t c, int32_t h, int32_t w, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && loss && n > 0 && c > 0 && h > 1 && w > 1, "tv_loss: bad arguments");
  const long long total = (long long)n * c * h * w;
  const float sh = 2.f / ((float)c * (h - 1) * w) / (float)n;
  const float sw = 2.f / ((float)c * h * (w - 1)) / (float)n;
  tv_loss_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, loss, gx, total, h, w, sh, sw);
  return after_launch("tv_loss");
}

extern "C" int cgb_bce_logits_loss(const float* x, const float* target, float* loss, float* gx, int64_t count, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && target && loss && count > 0, "bce_logits_loss: bad arguments");
  bce_logits_kernel<<<grid_for(count), 256, 0, (cudaStream_t)stream>>>(x, target, loss, gx, count, 1.f / (float)count);
  return after_launch("bce_logits_loss");
}

extern "C" int cgb_ground_intersection_loss(const float* pred, const float* ground, float* loss, int64_t count, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(pred && ground && loss && count > 0, "ground_intersection_loss: bad arguments");
  ground_intersection_kernel<<<grid_for(count), 256, 0, (cudaStream_t)stream>>>(pred, ground, loss, count, 1.f / (float)count);
  return after_launch("ground_intersection_loss");
}

extern "C" int cgb_sigm_loss(const float* pred, const float* target, float* loss, float* gpred, float* ws, int32_t n, int32_t h,
                             int32_t w, float gmweight, int32_t scales, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(pred && target && loss && gpred && ws && n > 0, "sigm_loss: bad arguments");
  CGB_REQUIRE(scales >= 1 && (h >> (scales - 1)) >= 3 && (w >> (scales - 1)) >= 3,
              "sigm_loss: %dx%d is too small for %d Sobel scales", h, w, scales);
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)n * h * w;
  // workspace layout (floats): [0,4) pred stats, [4,8) target stats, [8,12) two doubles, then R[total], G[total]
  float* sp = ws;
  float* stt = ws + 4;
  double* acc = reinterpret_cast<double*>(ws + 8);
  float* R = ws + 16;
  float* G = R + total;
  const float num_pix = (float)h * (float)w;
  median_stats_kernel<<<1, 1024, 0, st>>>(pred, total, sp);
  int s = after_launch("sigm_median(pred)");
  if (s) return s;
  median_stats_kernel<<<1, 1024, 0, st>>>(target, total, stt);
  if ((s = after_launch("sigm_median(target)"))) return s;
  cudaMemsetAsync(acc, 0, 2 * sizeof(double), st);
  sigm_residual_kernel<<<grid_for(total), 256, 0, st>>>(pred, target, sp, stt, R, G, loss, total, 0.5f / num_pix);
  if ((s = after_launch("sigm_residual"))) return s;
  for (int k = 0; k < scales; ++k) {
    const int f = 1 << k;
    const long long work = (long long)n * (h / f - 2) * (w / f - 2);
    // the reference's Sobel weights are expanded to [n,1,3,3], so conv2d returns n identical maps per image: factor n
    sigm_sobel_kernel<<<grid_for(work), 256, 0, st>>>(R, G, loss, n, h, w, f, gmweight * (float)n / num_pix);
    if ((s = after_launch("sigm_sobel"))) return s;
  }
  sigm_reduce_kernel<<<grid_for(total), 256, 0, st>>>(pred, G, sp, acc, total);
  if ((s = after_launch("sigm_reduce"))) return s;
  sigm_grad_kernel<<<grid_for(total), 256, 0, st>>>(pred, G, sp, acc, gpred, total);
  return after_launch("sigm_grad");
}
