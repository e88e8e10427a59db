// HBM-bound kernels of the hot path: instance-norm statistics, SPADE modulation (+backward),
// nearest resampling, layout conversion, compositing, L1 loss, spectral-norm power iteration.
// All are coalesced, 16/32-byte vectorised over the NHWC channel dimension; reductions use
// warp shuffles -> shared memory -> one fp64 atomic per (n,c) per CTA.
#include "common.cuh"
#include <vector>

namespace cgb {

// number of pixel-chunks per image for the reduction kernels
static inline int pick_chunks(int n, int hw, int cv) {
  // target >= 4 CTAs per SM overall, each CTA >= 256 pixels
  static const int per_sm = getenv("CGB_CHUNK_CTAS") ? atoi(getenv("CGB_CHUNK_CTAS")) : 4;   // measured: 8 -> 4 per SM: spade_mod_bwd 0.61 -> 0.68, in_stats 0.50 -> 0.54 of the HBM roof (profiles/r02_bn_grid_size.txt)
  int want = (148 * per_sm + n - 1) / n;
  int maxc = (hw + 255) / 256;
  if (want > maxc) want = maxc;
  if (want < 1) want = 1;
  return want;
}

// ---------------------------------------------------------------------------------------------------
// instance-norm statistics: x [n,hw,c] -> per-(n,c) sum and sum of squares.
// grid (chunks, n); block 256 threads = lanes x cv, cv = c/8 vector columns.  No atomics: every thread parks its 16 partial
// sums in shared memory, c threads fold the lanes and write ONE fp32 partial per (chunk, channel); the finalize kernel adds
// the chunks in fp64.  (The first version folded through shared and fp64 global atomics: with 200-1184 CTAs hitting the same
// 2c addresses the atomic tail cost more than the streaming pass itself, and made the result order-dependent.)
template <typename T>
__global__ void __launch_bounds__(256)
in_stats_kernel(const T* __restrict__ x, float* __restrict__ partial, int hw, int c, int px_per_chunk) {
  extern __shared__ float sm[];  // [256][16]
  const int cv = c >> 3;
  const int lanes = 256 / cv;
  const int tid = threadIdx.x;
  const int lane = tid / cv, v = tid - lane * cv;
  const int img = blockIdx.y;
  if (lane < lanes) {
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
    const int p0 = blockIdx.x * px_per_chunk;
    int p1 = p0 + px_per_chunk;
    if (p1 > hw) p1 = hw;
    const T* base = x + ((long long)img * hw) * c + v * 8;
    for (int p = p0 + lane; p < p1; p += lanes) {
      float f[8];
      Vec8<T>::load(base + (long long)p * c, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += f[j];
        q[j] = fmaf(f[j], f[j], q[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sm[tid * 16 + j] = s[j];
      sm[tid * 16 + 8 + j] = q[j];
    }
  }
  __syncthreads();
  float* out = partial + ((long long)img * gridDim.x + blockIdx.x) * 2 * c;
  for (int i = tid; i < c; i += 256) {
    const int vv = i >> 3, j = i & 7;
    float S = 0.f, Q = 0.f;
    for (int l = 0; l < lanes; ++l) {
      S += sm[(l * cv + vv) * 16 + j];
      Q += sm[(l * cv + vv) * 16 + 8 + j];
    }
    out[i] = S;
    out[c + i] = Q;
  }
}

// block (32 channels, 16 chunk groups): the chunk loop is split 16 ways and folded through shared memory, so a channel's 200
// partials cost ~13 dependent loads instead of 200 (the serial version took longer than the streaming pass it finishes)
__global__ void __launch_bounds__(512)
in_stats_finalize_kernel(const float* __restrict__ partial, float* __restrict__ mean, float* __restrict__ rstd, int c,
                         int chunks, double inv_hw, float eps) {
  __shared__ double sS[16][33], sQ[16][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int img = blockIdx.y;
  const int ch = blockIdx.x * 32 + tx;
  double S = 0.0, Q = 0.0;
  if (ch < c) {
    const float* p = partial + (long long)img * chunks * 2 * c;
    for (int k = ty; k < chunks; k += 16) {
      S += (double)p[(long long)k * 2 * c + ch];
      Q += (double)p[(long long)k * 2 * c + c + ch];
    }
  }
  sS[ty][tx] = S;
  sQ[ty][tx] = Q;
  __syncthreads();
  if (ty == 0 && ch < c) {
#pragma unroll
    for (int k = 1; k < 16; ++k) { S += sS[k][tx]; Q += sQ[k][tx]; }
    const double m = S * inv_hw;
    double var = Q * inv_hw - m * m;
    if (var < 0.0) var = 0.0;
    mean[(long long)img * c + ch] = (float)m;
    rstd[(long long)img * c + ch] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

// ---------------------------------------------------------------------------------------------------
// SPADE modulation forward.  grid (pixel chunks, n); a thread owns one 8-channel vector (its mean/rstd live in
// registers) and walks the chunk's pixels: 3 vector loads + 1 vector store per (pixel, vector), 2 FMAs per element.
template <typename T>
__global__ void __launch_bounds__(256)
spade_mod_fwd_kernel(const T* __restrict__ x, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const T* __restrict__ gb, T* __restrict__ out,
                     int hw, int c, int px_per_chunk, int act, float slope) {
  const int cv = c >> 3;
  const int lanes = 256 / cv;
  const int lane = threadIdx.x / cv, v = threadIdx.x - lane * cv;
  if (lane >= lanes) return;
  const int img = blockIdx.y;
  float rs[8], nm[8];  // xhat = x*rs + nm
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    rs[j] = rstd[(long long)img * c + v * 8 + j];
    nm[j] = -mean[(long long)img * c + v * 8 + j] * rs[j];
  }
  const float neg = act == CGB_ACT_NONE ? 1.f : (act == CGB_ACT_RELU ? 0.f : slope);
  const int p0 = blockIdx.x * px_per_chunk;
  const int p1 = min(hw, p0 + px_per_chunk);
  for (int p = p0 + lane; p < p1; p += lanes) {
    const long long pix = (long long)img * hw + p;
    float xv[8], g[8], b[8], o[8];
    Vec8<T>::load(x + pix * c + v * 8, xv);
    Vec8<T>::load(gb + pix * 2 * c + v * 8, g);
    Vec8<T>::load(gb + pix * 2 * c + c + v * 8, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = fmaf(xv[j], rs[j], nm[j]);
      const float t = fmaf(xh, 1.f + g[j], b[j]);
      o[j] = t > 0.f ? t : t * neg;
    }
    Vec8<T>::store(out + pix * c + v * 8, o);
  }
}

// SPADE modulation backward (part 1) — see cgb200.h
template <typename T>
__global__ void __launch_bounds__(256)
spade_mod_bwd_kernel(const T* __restrict__ x, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const T* __restrict__ gb,
                     const T* __restrict__ gout, T* __restrict__ ggb, T* __restrict__ gxhat,
                     double* __restrict__ sums, double* __restrict__ bsum, int hw, int c, int px_per_chunk, int act, float slope) {
  extern __shared__ float sm[];  // [2][c] (+ [2][c] column sums of ggb when bsum is given)
  const int cv = c >> 3;
  const int lanes = 256 / cv;
  const int tid = threadIdx.x;
  const int lane = tid / cv, v = tid - lane * cv;
  const int img = blockIdx.y;
  for (int i = tid; i < (bsum ? 4 : 2) * c; i += 256) sm[i] = 0.f;
  __syncthreads();
  if (lane < lanes) {
    float s1[8], s2[8], s3[8], s4[8], mu[8], rs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s1[j] = s2[j] = s3[j] = s4[j] = 0.f;
      mu[j] = mean[(long long)img * c + v * 8 + j];
      rs[j] = rstd[(long long)img * c + v * 8 + j];
    }
    const int p0 = blockIdx.x * px_per_chunk;
    int p1 = p0 + px_per_chunk;
    if (p1 > hw) p1 = hw;
    for (int p = p0 + lane; p < p1; p += lanes) {
      const long long pix = (long long)img * hw + p;
      float xv[8], g[8], b[8], go[8], gg[8], gbt[8], gxh[8];
      Vec8<T>::load(x + pix * c + v * 8, xv);
      Vec8<T>::load(gb + pix * 2 * c + v * 8, g);
      Vec8<T>::load(gb + pix * 2 * c + c + v * 8, b);
      Vec8<T>::load(gout + pix * c + v * 8, go);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (xv[j] - mu[j]) * rs[j];
        const float pre = fmaf(xh, 1.f + g[j], b[j]);
        float d = 1.f;
        if (act == CGB_ACT_LRELU) d = pre > 0.f ? 1.f : slope;
        else if (act == CGB_ACT_RELU) d = pre > 0.f ? 1.f : 0.f;
        const float gs = go[j] * d;
        gg[j] = gs * xh;
        gbt[j] = gs;
        gxh[j] = gs * (1.f + g[j]);
        s1[j] += gxh[j];
        s2[j] = fmaf(gxh[j], xh, s2[j]);
        s3[j] += gg[j];    // bias gradients of mlp_gamma / mlp_beta = column sums of ggb (used when bsum is given)
        s4[j] += gbt[j];
      }
      Vec8<T>::store(ggb + pix * 2 * c + v * 8, gg);
      Vec8<T>::store(ggb + pix * 2 * c + c + v * 8, gbt);
      Vec8<T>::store(gxhat + pix * c + v * 8, gxh);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&sm[v * 8 + j], s1[j]);
      atomicAdd(&sm[c + v * 8 + j], s2[j]);
    }
    if (bsum) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&sm[2 * c + v * 8 + j], s3[j]);
        atomicAdd(&sm[3 * c + v * 8 + j], s4[j]);
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < c; i += 256) {
    atomicAdd(&sums[((long long)img * c + i) * 2 + 0], (double)sm[i]);
    atomicAdd(&sums[((long long)img * c + i) * 2 + 1], (double)sm[c + i]);
  }
  if (bsum)
    for (int i = tid; i < 2 * c; i += 256) atomicAdd(&bsum[i], (double)sm[2 * c + i]);   // [gamma bias || beta bias], over all images
}

// gx = rstd*(g - m1 - xhat*m2) = A*g + B*x + C with per-(n,c) constants held in registers (same chunked mapping)
template <typename T>
__global__ void __launch_bounds__(256)
in_bwd_kernel(const T* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
              const double* __restrict__ sums, T* __restrict__ g, int hw, int c, int px_per_chunk) {
  const int cv = c >> 3;
  const int lanes = 256 / cv;
  const int lane = threadIdx.x / cv, v = threadIdx.x - lane * cv;
  if (lane >= lanes) return;
  const int img = blockIdx.y;
  const float inv_hw = 1.f / (float)hw;
  float A[8], B[8], Cc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const long long sc = (long long)img * c + v * 8 + j;
    const float rs = rstd[sc], mu = mean[sc];
    const float m1 = (float)sums[sc * 2 + 0] * inv_hw;
    const float m2 = (float)sums[sc * 2 + 1] * inv_hw;
    A[j] = rs;
    B[j] = -rs * rs * m2;
    Cc[j] = rs * rs * m2 * mu - rs * m1;
  }
  const int p0 = blockIdx.x * px_per_chunk;
  const int p1 = min(hw, p0 + px_per_chunk);
  for (int p = p0 + lane; p < p1; p += lanes) {
    const long long off = ((long long)img * hw + p) * c + v * 8;
    float xv[8], gv[8], o[8];
    Vec8<T>::load(x + off, xv);
    Vec8<T>::load(g + off, gv);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf(A[j], gv[j], fmaf(B[j], xv[j], Cc[j]));
    Vec8<T>::store(g + off, o);
  }
}

// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
resize_nearest_kernel(const T* __restrict__ x, T* __restrict__ y, long long total_vec, int hi, int wi,
                      int ho, int wo, int c) {
  const int cv = c >> 3;
  const float sh = (float)hi / (float)ho, sw = (float)wi / (float)wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / cv;
    const int v = (int)(i - pix * cv);
    const int ox = (int)(pix % wo);
    const long long t = pix / wo;
    const int oy = (int)(t % ho);
    const int img = (int)(t / ho);
    // ATen nearest: src = min(floor(dst * scale), in-1), scale = in/out in fp32
    int sy = (int)floorf(oy * sh);
    int sx = (int)floorf(ox * sw);
    if (sy > hi - 1) sy = hi - 1;
    if (sx > wi - 1) sx = wi - 1;
    const uint4* src = reinterpret_cast<const uint4*>(x + (((long long)img * hi + sy) * wi + sx) * c + v * 8);
    uint4* dst = reinterpret_cast<uint4*>(y + pix * c + v * 8);
    dst[0] = src[0];
    if (sizeof(T) == 4) dst[1] = src[1];
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
upsample_bwd_kernel(const T* __restrict__ gy, T* __restrict__ gx, long long total_vec, int hi, int wi,
                    int f, int c) {
  const int cv = c >> 3;
  const int ho = hi * f, wo = wi * f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / cv;
    const int v = (int)(i - pix * cv);
    const int ix = (int)(pix % wi);
    const long long t = pix / wi;
    const int iy = (int)(t % hi);
    const int img = (int)(t / hi);
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
    for (int a = 0; a < f; ++a)
      for (int b = 0; b < f; ++b) {
        float g[8];
        Vec8<T>::load(gy + (((long long)img * ho + iy * f + a) * wo + ix * f + b) * c + v * 8, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += g[j];
      }
    Vec8<T>::store(gx + pix * c + v * 8, s);
  }
}

// Adjoint of resize_nearest_kernel for ANY size ratio, gather form (deterministic, no atomics): an input pixel sums the
// output pixels whose ATen nearest source index is that pixel.  Candidate rows / columns come from inverting
// src = min(floor(dst * in/out), in-1) with one pixel of slack, then the forward formula itself decides membership.
__device__ __forceinline__ int nearest_src(int dst, float scale, int in) {
  int s = (int)floorf(dst * scale);
  return s > in - 1 ? in - 1 : s;
}
template <typename T>
__global__ void __launch_bounds__(256)
resize_nearest_bwd_kernel(const T* __restrict__ gy, T* __restrict__ gx, long long total_vec, int hi, int wi, int ho, int wo,
                          int c) {
  const int cv = c >> 3;
  const float sh = (float)hi / (float)ho, sw = (float)wi / (float)wo;
  const float rh = (float)ho / (float)hi, rw = (float)wo / (float)wi;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / cv;
    const int v = (int)(i - pix * cv);
    const int ix = (int)(pix % wi);
    const long long t = pix / wi;
    const int iy = (int)(t % hi);
    const int img = (int)(t / hi);
    int y0 = (int)floorf(iy * rh) - 1, y1 = (int)ceilf((iy + 1) * rh) + 1;
    int x0 = (int)floorf(ix * rw) - 1, x1 = (int)ceilf((ix + 1) * rw) + 1;
    if (y0 < 0) y0 = 0;
    if (x0 < 0) x0 = 0;
    if (y1 > ho - 1) y1 = ho - 1;
    if (x1 > wo - 1) x1 = wo - 1;
    if (iy == hi - 1) y1 = ho - 1;   // the clamp to in-1 folds every overshooting row / column onto the last one
    if (ix == wi - 1) x1 = wo - 1;
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
    for (int oy = y0; oy <= y1; ++oy) {
      if (nearest_src(oy, sh, hi) != iy) continue;
      for (int ox = x0; ox <= x1; ++ox) {
        if (nearest_src(ox, sw, wi) != ix) continue;
        float g[8];
        Vec8<T>::load(gy + (((long long)img * ho + oy) * wo + ox) * c + v * 8, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += g[j];
      }
    }
    Vec8<T>::store(gx + pix * c + v * 8, s);
  }
}

// ---------------------------------------------------------------------------------------------------
// NCHW fp32 <-> NHWC storage.  Tile transpose through shared memory: 32 pixels x cs channels.
template <typename T>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ x, T* __restrict__ y, int c, int hw, int cs) {
  __shared__ float tile[64 * 33];  // [64 channels][32 pixels + 1]
  const int img = blockIdx.y;
  const int p0 = blockIdx.x * 32;
  const int c0 = blockIdx.z * 64;
  const int cn = min(64, cs - c0);  // channels in this block (multiple of 8)
  for (int i = threadIdx.x; i < cn * 32; i += blockDim.x) {
    const int ch = i >> 5, p = i & 31;
    float v = 0.f;
    if (c0 + ch < c && p0 + p < hw) v = x[((long long)img * c + c0 + ch) * hw + p0 + p];
    tile[ch * 33 + p] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cn * 32; i += blockDim.x) {
    const int p = i / cn, ch = i - p * cn;
    if (p0 + p < hw) y[((long long)img * hw + p0 + p) * cs + c0 + ch] = from_f<T>(tile[ch * 33 + p]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const T* __restrict__ x, float* __restrict__ y, int c, int hw, int cs) {
  __shared__ float tile[64 * 33];
  const int img = blockIdx.y;
  const int p0 = blockIdx.x * 32;
  const int c0 = blockIdx.z * 64;
  const int cn = min(64, cs - c0);
  for (int i = threadIdx.x; i < cn * 32; i += blockDim.x) {
    const int p = i / cn, ch = i - p * cn;
    float v = 0.f;
    if (p0 + p < hw) v = to_f<T>(x[((long long)img * hw + p0 + p) * cs + c0 + ch]);
    tile[ch * 33 + p] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cn * 32; i += blockDim.x) {
    const int ch = i >> 5, p = i & 31;
    if (c0 + ch < c && p0 + p < hw) y[((long long)img * c + c0 + ch) * hw + p0 + p] = tile[ch * 33 + p];
  }
}

// Few-channel layout edges (images, masks, depth, seg logits: cs <= 32) — one thread per PIXEL: the c channel planes are read
// (written) coalesced over the pixel index and the pixel's channel vector is written (read) as 16-byte chunks.  The 64 x 32
// tile kernels above move one element per thread there (0.11 of the HBM roofline on an 8 x 3 x 640 x 640 image batch).
template <typename T, int CS>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_small_kernel(const float* __restrict__ x, T* __restrict__ y, int c, int hw, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long img = i / hw;
    const int p = (int)(i - img * hw);
    const float* src = x + img * c * hw + p;
#pragma unroll
    for (int v = 0; v < CS / 8; ++v) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (v * 8 + j < c) ? __ldg(src + (long long)(v * 8 + j) * hw) : 0.f;
      Vec8<T>::store(y + i * CS + v * 8, f);
    }
  }
}

template <typename T, int CS>
__global__ void __launch_bounds__(256)
nhwc_to_nchw_small_kernel(const T* __restrict__ x, float* __restrict__ y, int c, int hw, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long img = i / hw;
    const int p = (int)(i - img * hw);
    float* dst = y + img * c * hw + p;
#pragma unroll
    for (int v = 0; v < CS / 8; ++v) {
      if (v * 8 >= c) break;
      float f[8];
      Vec8<T>::load(x + i * CS + v * 8, f);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (v * 8 + j < c) dst[(long long)(v * 8 + j) * hw] = f[j];
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
act_bwd_kernel(const T* __restrict__ gy, const T* __restrict__ y, T* __restrict__ gx, long long count,
               int act, float slope) {
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < count;
       i += (long long)gridDim.x * blockDim.x * 8) {
    float g[8], yy[8], o[8];
    Vec8<T>::load(gy + i, g);
    Vec8<T>::load(y + i, yy);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = g[j] * act_grad_from_out(yy[j], act, slope);
    Vec8<T>::store(gx + i, o);
  }
}

// gx = gy * act'(y) and gbias[c] += sum over pixels of gx (fp32, before the rounding to the storage type): the bias gradient of the
// conv whose fused activation is being differentiated, so its weight-gradient launch needs no column-sum pass over gx
template <typename T>
__global__ void __launch_bounds__(256)
act_bwd_bias_kernel(const T* __restrict__ gy, const T* __restrict__ y, T* __restrict__ gx, float* __restrict__ gbias,
                    long long pixels, int c, long long px_per_cta, int act, float slope) {
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
      float g[8], yy[8], o[8];
      Vec8<T>::load(gy + px * c + v * 8, g);
      Vec8<T>::load(y + px * c + v * 8, yy);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j] = g[j] * act_grad_from_out(yy[j], act, slope);
        s[j] += o[j];
      }
      Vec8<T>::store(gx + px * c + v * 8, o);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&sm[v * 8 + j], s[j]);
  }
  __syncthreads();
  for (int i = tid; i < c; i += 256) atomicAdd(gbias + i, sm[i]);
}

template <typename T>
__global__ void __launch_bounds__(256)
act_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long count, int act, float slope) {
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < count;
       i += (long long)gridDim.x * blockDim.x * 8) {
    float v[8], o[8];
    Vec8<T>::load(x + i, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = act_apply(v[j], act, slope);
    Vec8<T>::store(y + i, o);
  }
}

// ---------------------------------------------------------------------------------------------------
// instance norm + activation (discriminator.py:120-133: conv -> InstanceNorm2d(affine=False) -> LeakyReLU(0.2))
template <typename T>
__global__ void __launch_bounds__(256)
in_apply_fwd_kernel(const T* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                    T* __restrict__ y, int hw, int c, int px_per_chunk, float neg) {
  const int cv = c >> 3;
  const int lanes = 256 / cv;
  const int lane = threadIdx.x / cv, v = threadIdx.x - lane * cv;
  if (lane >= lanes) return;
  const int img = blockIdx.y;
  float rs[8], nm[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    rs[j] = rstd[(long long)img * c + v * 8 + j];
    nm[j] = -mean[(long long)img * c + v * 8 + j] * rs[j];
  }
  const int p0 = blockIdx.x * px_per_chunk;
  const int p1 = min(hw, p0 + px_per_chunk);
  for (int p = p0 + lane; p < p1; p += lanes) {
    const long long off = ((long long)img * hw + p) * c + v * 8;
    float xv[8], o[8];
    Vec8<T>::load(x + off, xv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float t = fmaf(xv[j], rs[j], nm[j]);
      o[j] = fmaxf(t, t * neg);
    }
    Vec8<T>::store(y + off, o);
  }
}

// backward part 1: gxhat = gy * act'(xhat) (in place allowed), sums += (sum gxhat, sum gxhat*xhat)
template <typename T>
__global__ void __launch_bounds__(256)
in_apply_bwd_kernel(const T* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                    const T* __restrict__ gy, T* __restrict__ gxhat, double* __restrict__ sums, int hw, int c,
                    int px_per_chunk, float neg) {
  extern __shared__ float sm[];
  const int cv = c >> 3;
  const int lanes = 256 / cv;
  const int tid = threadIdx.x;
  const int lane = tid / cv, v = tid - lane * cv;
  const int img = blockIdx.y;
  for (int i = tid; i < 2 * c; i += 256) sm[i] = 0.f;
  __syncthreads();
  if (lane < lanes) {
    float s1[8], s2[8], rs[8], nm[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s1[j] = s2[j] = 0.f;
      rs[j] = rstd[(long long)img * c + v * 8 + j];
      nm[j] = -mean[(long long)img * c + v * 8 + j] * rs[j];
    }
    const int p0 = blockIdx.x * px_per_chunk;
    const int p1 = min(hw, p0 + px_per_chunk);
    for (int p = p0 + lane; p < p1; p += lanes) {
      const long long off = ((long long)img * hw + p) * c + v * 8;
      float xv[8], g[8], o[8];
      Vec8<T>::load(x + off, xv);
      Vec8<T>::load(gy + off, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = fmaf(xv[j], rs[j], nm[j]);
        o[j] = g[j] * (xh > 0.f ? 1.f : neg);
        s1[j] += o[j];
        s2[j] = fmaf(o[j], xh, s2[j]);
      }
      Vec8<T>::store(gxhat + off, o);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&sm[v * 8 + j], s1[j]);
      atomicAdd(&sm[c + v * 8 + j], s2[j]);
    }
  }
  __syncthreads();
  for (int i = tid; i < c; i += 256) {
    atomicAdd(&sums[((long long)img * c + i) * 2 + 0], (double)sm[i]);
    atomicAdd(&sums[((long long)img * c + i) * 2 + 1], (double)sm[c + i]);
  }
}

// nn.AvgPool2d(3, stride=2, padding=1, count_include_pad=False) (discriminator.py:223-225), NHWC
template <typename T, bool BWD>
__global__ void __launch_bounds__(256)
avgpool3s2_kernel(const T* __restrict__ src, T* __restrict__ dst, long long total_vec, int hi, int wi, int ho, int wo,
                  int c) {
  const int cv = c >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / cv;
    const int v = (int)(i - pix * cv);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (!BWD) {  // dst = pooled output pixel (oy, ox)
      const int ox = (int)(pix % wo);
      const long long t = pix / wo;
      const int oy = (int)(t % ho);
      const long long img = t / ho;
      int cnt = 0;
      for (int dy = 0; dy < 3; ++dy)
        for (int dx = 0; dx < 3; ++dx) {
          const int sy = oy * 2 - 1 + dy, sx = ox * 2 - 1 + dx;
          if (sy < 0 || sy >= hi || sx < 0 || sx >= wi) continue;
          float f[8];
          Vec8<T>::load(src + ((img * hi + sy) * wi + sx) * c + v * 8, f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += f[j];
          ++cnt;
        }
      const float inv = 1.f / (float)cnt;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] *= inv;
    } else {     // dst = input-gradient pixel (iy, ix); src = pooled-output gradient
      const int ix = (int)(pix % wi);
      const long long t = pix / wi;
      const int iy = (int)(t % hi);
      const long long img = t / hi;
      for (int oy = (iy + 1) / 2 - 1; oy <= (iy + 1) / 2; ++oy)
        for (int ox = (ix + 1) / 2 - 1; ox <= (ix + 1) / 2; ++ox) {
          if (oy < 0 || oy >= ho || ox < 0 || ox >= wo) continue;
          if (iy < oy * 2 - 1 || iy > oy * 2 + 1 || ix < ox * 2 - 1 || ix > ox * 2 + 1) continue;
          const int y0 = max(oy * 2 - 1, 0), y1 = min(oy * 2 + 1, hi - 1);
          const int x0 = max(ox * 2 - 1, 0), x1 = min(ox * 2 + 1, wi - 1);
          const float inv = 1.f / (float)((y1 - y0 + 1) * (x1 - x0 + 1));
          float f[8];
          Vec8<T>::load(src + ((img * ho + oy) * wo + ox) * c + v * 8, f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(f[j], inv, acc[j]);
        }
    }
    Vec8<T>::store(dst + pix * c + v * 8, acc);
  }
}

// mean-reduced losses against a constant target (GANLoss / HingeLoss, losses.py:13-83, :550-593) on fp32 arrays.
//   kind 0: BCE-with-logits(x, t)   1: MSE(x, t)   2: hinge D real  -mean(min(x-1,0))   3: hinge D fake -mean(min(-x-1,0))
//   kind 4: -mean(x) (hinge / WGAN generator term)
// loss[0] += scale * sum(l_i), gx_i = scale * dl_i/dx_i   (scale = weight / count)
__global__ void __launch_bounds__(256)
const_target_loss_kernel(const float* __restrict__ x, float* __restrict__ loss, float* __restrict__ gx, long long count,
                         int kind, float t, float scale, const float* __restrict__ t_dev) {
  if (t_dev) t = *t_dev;   // the target lives in device memory (a captured CUDA graph reads this step's draw from there)
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    float l, g;
    if (kind == 0) {
      l = fmaxf(v, 0.f) - v * t + log1pf(__expf(-fabsf(v)));
      g = 1.f / (1.f + __expf(-v)) - t;
    } else if (kind == 1) {
      l = (v - t) * (v - t);
      g = 2.f * (v - t);
    } else if (kind == 2) {
      l = -fminf(v - 1.f, 0.f);
      g = v < 1.f ? -1.f : 0.f;
    } else if (kind == 3) {
      l = -fminf(-v - 1.f, 0.f);
      g = v > -1.f ? 1.f : 0.f;
    } else {
      l = -v;
      g = -1.f;
    }
    s += l;
    if (gx) gx[i] = g * scale;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tt = 0.f;
    for (int i = 0; i < 8; ++i) tt += ws[i];
    atomicAdd(loss, tt * scale);
  }
}

// L1 between two storage tensors (FeatMatchLoss, losses.py:86-103): loss += scale*sum|a-b|, ga = scale*sign(a-b)
template <typename T>
__global__ void __launch_bounds__(256)
l1_storage_kernel(const T* __restrict__ a, const T* __restrict__ b, float* __restrict__ loss, T* __restrict__ ga,
                  long long count, float scale) {
  float s = 0.f;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < count;
       i += (long long)gridDim.x * blockDim.x * 8) {
    float x[8], y[8], g[8];
    Vec8<T>::load(a + i, x);
    Vec8<T>::load(b + i, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = x[j] - y[j];
      s += fabsf(d);
      g[j] = d > 0.f ? scale : (d < 0.f ? -scale : 0.f);
    }
    if (ga) Vec8<T>::store(ga + i, g);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tt = 0.f;
    for (int i = 0; i < 8; ++i) tt += ws[i];
    atomicAdd(loss, tt * scale);
  }
}

// ---------------------------------------------------------------------------------------------------
// ExtraAdam (climategan/optim.py:137-291) over flat fp32 arrays: one launch per parameter group.
//   m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g ; u = -step_size * m / (sqrt(v) + eps)
//   mode 0 (extrapolation, optim.py:153-170): if save_copy: c = p ;  p += u
//   mode 1 (step, optim.py:172-197):           p = c + u
__global__ void __launch_bounds__(256)
extra_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                  float* __restrict__ c, long long count, float b1, float b2, float eps, float wd, float step_size,
                  int mode, int save_copy) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i];
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float u = -step_size * mi / (sqrtf(vi) + eps);
    if (mode == 0) {
      if (save_copy) c[i] = pi;
      p[i] = pi + u;
    } else {
      p[i] = c[i] + u;
    }
  }
}

// nn.MaxPool2d(2, 2) of torchvision's VGG19 features (losses.py:304-336), NHWC; bwd routes to the first max
template <typename T>
__global__ void __launch_bounds__(256)
maxpool2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long total_vec, int hi, int wi, int c) {
  const int cv = c >> 3;
  const int ho = hi >> 1, wo = wi >> 1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / cv;
    const int v = (int)(i - pix * cv);
    const int ox = (int)(pix % wo);
    const long long t = pix / wo;
    const int oy = (int)(t % ho);
    const long long img = t / ho;
    float best[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) best[j] = -3.4e38f;
    for (int dy = 0; dy < 2; ++dy)
      for (int dx = 0; dx < 2; ++dx) {
        float f[8];
        Vec8<T>::load(x + ((img * hi + oy * 2 + dy) * wi + ox * 2 + dx) * c + v * 8, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) best[j] = fmaxf(best[j], f[j]);
      }
    Vec8<T>::store(y + pix * c + v * 8, best);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
maxpool2_bwd_kernel(const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ gy, T* __restrict__ gx,
                    long long total_vec, int hi, int wi, int c) {
  const int cv = c >> 3;
  const int ho = hi >> 1, wo = wi >> 1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / cv;  // pooled pixel
    const int v = (int)(i - pix * cv);
    const int ox = (int)(pix % wo);
    const long long t = pix / wo;
    const int oy = (int)(t % ho);
    const long long img = t / ho;
    float ym[8], g[8];
    bool taken[8];
    Vec8<T>::load(y + pix * c + v * 8, ym);
    Vec8<T>::load(gy + pix * c + v * 8, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) taken[j] = false;
    for (int dy = 0; dy < 2; ++dy)
      for (int dx = 0; dx < 2; ++dx) {
        const long long off = ((img * hi + oy * 2 + dy) * wi + ox * 2 + dx) * c + v * 8;
        float f[8], o[8];
        Vec8<T>::load(x + off, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const bool hit = !taken[j] && f[j] == ym[j];
          o[j] = hit ? g[j] : 0.f;
          taken[j] = taken[j] || hit;
        }
        Vec8<T>::store(gx + off, o);
      }
  }
}

// vgg_preprocess(img * m) (climategan/tutils.py:416-427 applied to fake*m / x*m, trainer.py:1281-1283):
// RGB->BGR, [-1,1] -> [0,255], subtract the caffe means; NCHW fp32 in, NHWC storage (8 ch) out.  bwd is the adjoint.
template <typename T>
__global__ void __launch_bounds__(256)
vgg_pre_fwd_kernel(const float* __restrict__ x, const float* __restrict__ m, T* __restrict__ y, int hw, long long total) {
  const float mean[3] = {103.939f, 116.779f, 123.680f};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long img = i / hw;
    const int p = (int)(i - img * hw);
    const float mk = m ? m[i] : 1.f;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) v[ch] = (x[(img * 3 + (2 - ch)) * hw + p] * mk + 1.f) * 127.5f - mean[ch];
    Vec8<T>::store(y + i * 8, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
vgg_pre_bwd_kernel(const T* __restrict__ gy, const float* __restrict__ m, float* __restrict__ gx, int hw, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long img = i / hw;
    const int p = (int)(i - img * hw);
    const float mk = m ? m[i] : 1.f;
    float g[8];
    Vec8<T>::load(gy + i * 8, g);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) gx[(img * 3 + (2 - ch)) * hw + p] = g[ch] * 127.5f * mk;
  }
}

// ---------------------------------------------------------------------------------------------------
// masker inference helpers (NHWC storage)
// nn.MaxPool2d(3, stride=2, padding=0, ceil_mode=True) (deeplab/resnetmulti_v2.py:76-78)
template <typename T>
__global__ void __launch_bounds__(256)
maxpool3s2_ceil_kernel(const T* __restrict__ x, T* __restrict__ y, long long total_vec, int hi, int wi, int ho, int wo, int c,
                       int pad = 0) {
  const int cv = c >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / cv;
    const int v = (int)(i - pix * cv);
    const int ox = (int)(pix % wo);
    const long long t = pix / wo;
    const int oy = (int)(t % ho);
    const long long img = t / ho;
    float best[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) best[j] = -3.4e38f;
    for (int dy = 0; dy < 3; ++dy)
      for (int dx = 0; dx < 3; ++dx) {
        const int sy = oy * 2 + dy - pad, sx = ox * 2 + dx - pad;
        if (sy < 0 || sx < 0 || sy >= hi || sx >= wi) continue;
        float f[8];
        Vec8<T>::load(x + ((img * hi + sy) * wi + sx) * c + v * 8, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) best[j] = fmaxf(best[j], f[j]);
      }
    Vec8<T>::store(y + pix * c + v * 8, best);
  }
}

// F.interpolate(mode="bilinear", align_corners=ac) (ATen area_pixel_compute_source_index semantics)
template <typename T>
__global__ void __launch_bounds__(256)
resize_bilinear_kernel(const T* __restrict__ x, T* __restrict__ y, long long total_vec, int hi, int wi, int ho, int wo,
                       int c, int ac) {
  const int cv = c >> 3;
  const float sh = ac ? (ho > 1 ? (float)(hi - 1) / (float)(ho - 1) : 0.f) : (float)hi / (float)ho;
  const float sw = ac ? (wo > 1 ? (float)(wi - 1) / (float)(wo - 1) : 0.f) : (float)wi / (float)wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / cv;
    const int v = (int)(i - pix * cv);
    const int ox = (int)(pix % wo);
    const long long t = pix / wo;
    const int oy = (int)(t % ho);
    const long long img = t / ho;
    float fy = ac ? sh * oy : fmaxf(sh * (oy + 0.5f) - 0.5f, 0.f);
    float fx = ac ? sw * ox : fmaxf(sw * (ox + 0.5f) - 0.5f, 0.f);
    const int y0 = min((int)fy, hi - 1), x0 = min((int)fx, wi - 1);
    const int y1 = min(y0 + 1, hi - 1), x1 = min(x0 + 1, wi - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    float a[8], b[8], cc[8], d[8], o[8];
    const T* base = x + img * hi * wi * c + v * 8;
    Vec8<T>::load(base + ((long long)y0 * wi + x0) * c, a);
    Vec8<T>::load(base + ((long long)y0 * wi + x1) * c, b);
    Vec8<T>::load(base + ((long long)y1 * wi + x0) * c, cc);
    Vec8<T>::load(base + ((long long)y1 * wi + x1) * c, d);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      o[j] = (1.f - ly) * ((1.f - lx) * a[j] + lx * b[j]) + ly * ((1.f - lx) * cc[j] + lx * d[j]);
    Vec8<T>::store(y + pix * c + v * 8, o);
  }
}

// F.interpolate(mode="bicubic", align_corners=False) (A = -0.75, clamped taps) — depth.py:144-149
__device__ __forceinline__ void cubic_coeffs(float t, float (&w)[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
  w[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  w[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
  w[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  w[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}

template <typename T>
__global__ void __launch_bounds__(256)
resize_bicubic_kernel(const T* __restrict__ x, T* __restrict__ y, long long total_vec, int hi, int wi, int ho, int wo, int c) {
  const int cv = c >> 3;
  const float sh = (float)hi / (float)ho, sw = (float)wi / (float)wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / cv;
    const int v = (int)(i - pix * cv);
    const int ox = (int)(pix % wo);
    const long long t = pix / wo;
    const int oy = (int)(t % ho);
    const long long img = t / ho;
    const float fy = sh * (oy + 0.5f) - 0.5f, fx = sw * (ox + 0.5f) - 0.5f;
    const int iy = (int)floorf(fy), ix = (int)floorf(fx);
    float wy[4], wx[4];
    cubic_coeffs(fy - (float)iy, wy);
    cubic_coeffs(fx - (float)ix, wx);
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.f;
    const T* base = x + img * hi * wi * c + v * 8;
    for (int a = 0; a < 4; ++a) {
      const int sy = min(max(iy - 1 + a, 0), hi - 1);
      for (int b = 0; b < 4; ++b) {
        const int sx = min(max(ix - 1 + b, 0), wi - 1);
        float f[8];
        Vec8<T>::load(base + ((long long)sy * wi + sx) * c, f);
        const float wgt = wy[a] * wx[b];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(wgt, f[j], o[j]);
      }
    }
    Vec8<T>::store(y + pix * c + v * 8, o);
  }
}

// torch.mean(x, dim=1, keepdim=True) over the c_logical real channels -> channel 0 of an 8-channel storage tensor
template <typename T>
__global__ void __launch_bounds__(256)
channel_mean_kernel(const T* __restrict__ x, T* __restrict__ y, long long pixels, int cs, int c_logical) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < pixels; p += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int v = 0; v < cs; v += 8) {
      float f[8];
      Vec8<T>::load(x + p * cs + v, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += (v + j < c_logical) ? f[j] : 0.f;
    }
    float o[8] = {s / (float)c_logical, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    Vec8<T>::store(y + p * 8, o);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
mul_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y, long long count) {
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < count; i += (long long)gridDim.x * blockDim.x * 8) {
    float x[8], z[8], o[8];
    Vec8<T>::load(a + i, x);
    Vec8<T>::load(b + i, z);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = x[j] * z[j];
    Vec8<T>::store(y + i, o);
  }
}

// per-sample min / max of channel 0 (tutils.normalize :567-575): mm[n] = {min, max}; caller initialises to {+inf,-inf}
__device__ __forceinline__ void atomic_min_f(float* addr, float v) {
  int* ai = reinterpret_cast<int*>(addr);
  int old = *ai;
  while (__int_as_float(old) > v) {
    const int assumed = old;
    old = atomicCAS(ai, assumed, __float_as_int(v));
    if (old == assumed) break;
  }
}
__device__ __forceinline__ void atomic_max_f(float* addr, float v) {
  int* ai = reinterpret_cast<int*>(addr);
  int old = *ai;
  while (__int_as_float(old) < v) {
    const int assumed = old;
    old = atomicCAS(ai, assumed, __float_as_int(v));
    if (old == assumed) break;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
minmax_c0_kernel(const T* __restrict__ x, float* __restrict__ mm, int hw, int cs) {
  const int img = blockIdx.y;
  float lo = 3.4e38f, hi = -3.4e38f;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
    const float v = to_f<T>(x[((long long)img * hw + p) * cs]);
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomic_min_f(mm + 2 * img, lo);
    atomic_max_f(mm + 2 * img + 1, hi);
  }
}

__global__ void mm_init_kernel(float* __restrict__ mm, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { mm[2 * i] = 3.4e38f; mm[2 * i + 1] = -3.4e38f; }
}

// OmniGenerator.make_m_cond (generator.py:196-230): cat[ normalize(d), softmax(s, dim=1), x resized ] -> 16-ch storage
//   d [n,hw,8] (channel 0), s [n,hw,ss] with ns classes, xr [n,hw,8] (3 channels, already bilinear-resized), mm per-sample min/max
template <typename T>
__global__ void __launch_bounds__(256)
m_cond_kernel(const T* __restrict__ d, const T* __restrict__ s, const T* __restrict__ xr, const float* __restrict__ mm,
              T* __restrict__ out, long long pixels, int hw, int ss, int ns, int cs_out) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < pixels; p += (long long)gridDim.x * blockDim.x) {
    const int img = (int)(p / hw);
    const float lo = mm[2 * img], hi = mm[2 * img + 1];
    float o[24];
#pragma unroll
    for (int j = 0; j < 24; ++j) o[j] = 0.f;
    o[0] = (to_f<T>(d[p * 8]) - lo) / (hi - lo);
    float mx = -3.4e38f;
    for (int k = 0; k < ns; ++k) mx = fmaxf(mx, to_f<T>(s[p * ss + k]));
    float sum = 0.f;
    for (int k = 0; k < ns; ++k) {
      const float e = __expf(to_f<T>(s[p * ss + k]) - mx);
      o[1 + k] = e;
      sum += e;
    }
    const float inv = 1.f / sum;
    for (int k = 0; k < ns; ++k) o[1 + k] *= inv;
    if (xr) {
      for (int k = 0; k < 3; ++k) o[1 + ns + k] = to_f<T>(xr[p * 8 + k]);
    }
    for (int k = 0; k < cs_out; ++k) out[p * cs_out + k] = from_f<T>(o[k]);
  }
}

// Adjoint of m_cond_kernel w.r.t. d and s (the conditioning is differentiable when gen.m.spade.detach is false,
// defaults.yaml:182; generator.py:217-219).  One CTA per sample (d / s are size/4 maps: 160x160 at 640x640).
//   y = (d - lo) / r, r = hi - lo (tutils.normalize :566-577):  gd_j = g_j / r - [j = argmin] S0 / r - ([j = argmax] - [j = argmin]) S1 / r
//   with S0 = sum_i g_i, S1 = sum_i g_i y_i; argmin / argmax = first occurrence in row-major order (torch.min / max with dim).
//   p = softmax(s): gs_k = p_k (g_k - sum_j g_j p_j), p read back from the forward's output.
template <typename T>
__global__ void __launch_bounds__(1024)
m_cond_bwd_kernel(const T* __restrict__ d, const T* __restrict__ out, const float* __restrict__ mm, const T* __restrict__ gout,
                  T* __restrict__ gd, T* __restrict__ gs, int hw, int ss, int ns, int cs_out) {
  const int img = blockIdx.x;
  const long long base = (long long)img * hw;
  const float lo = mm[2 * img], hi = mm[2 * img + 1];
  const float inv_r = 1.f / (hi - lo);
  float s0 = 0.f, s1 = 0.f;
  int imin = 0x7fffffff, imax = 0x7fffffff;
  for (int p = threadIdx.x; p < hw; p += blockDim.x) {
    const float g = to_f<T>(gout[(base + p) * cs_out]);
    const float v = to_f<T>(d[(base + p) * 8]);
    s0 += g;
    s1 += g * ((v - lo) * inv_r);
    if (v == lo && p < imin) imin = p;
    if (v == hi && p < imax) imax = p;
  }
  __shared__ float sh0[32], sh1[32];
  __shared__ int shmin[32], shmax[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    imin = min(imin, __shfl_xor_sync(0xffffffffu, imin, o));
    imax = min(imax, __shfl_xor_sync(0xffffffffu, imax, o));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) { sh0[warp] = s0; sh1[warp] = s1; shmin[warp] = imin; shmax[warp] = imax; }
  __syncthreads();
  if (warp == 0) {
    s0 = lane < nw ? sh0[lane] : 0.f;
    s1 = lane < nw ? sh1[lane] : 0.f;
    imin = lane < nw ? shmin[lane] : 0x7fffffff;
    imax = lane < nw ? shmax[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      imin = min(imin, __shfl_xor_sync(0xffffffffu, imin, o));
      imax = min(imax, __shfl_xor_sync(0xffffffffu, imax, o));
    }
    if (lane == 0) { sh0[0] = s0; sh1[0] = s1; shmin[0] = imin; shmax[0] = imax; }
  }
  __syncthreads();
  s0 = sh0[0]; s1 = sh1[0]; imin = shmin[0]; imax = shmax[0];
  for (int p = threadIdx.x; p < hw; p += blockDim.x) {
    const long long q = base + p;
    float g = to_f<T>(gout[q * cs_out]) * inv_r;
    if (p == imin) g += (s1 - s0) * inv_r;
    if (p == imax) g -= s1 * inv_r;
    gd[q * 8] = from_f<T>(g);
#pragma unroll
    for (int k = 1; k < 8; ++k) gd[q * 8 + k] = from_f<T>(0.f);
    float dot = 0.f;
    for (int k = 0; k < ns; ++k) dot += to_f<T>(gout[q * cs_out + 1 + k]) * to_f<T>(out[q * cs_out + 1 + k]);
    for (int k = 0; k < ss; ++k) {
      float v = 0.f;
      if (k < ns) v = to_f<T>(out[q * cs_out + 1 + k]) * (to_f<T>(gout[q * cs_out + 1 + k]) - dot);
      gs[q * ss + k] = from_f<T>(v);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// im2col of a few-channel tensor (the 3-channel SPADE conditioning): y[n,oy,ox, tap*c + ch] = x[n,oy+dy*dil-pad,
// ox+dx*dil-pad, ch], zero outside; lets SPADE.mlp_shared (norms.py:164-166) run as a K=32 1x1 GEMM on the
// tensor cores instead of 9 mostly-empty 64-channel K blocks.  One thread per (pixel, 8 output channels).
template <typename T>
__global__ void __launch_bounds__(256)
im2col_kernel(const T* __restrict__ x, T* __restrict__ y, long long total_vec, int h, int w, int cs_in, int c,
              int k, int pad, int dil, int cs_out, int stride = 1, int ho = 0, int wo = 0) {
  const int ov = cs_out >> 3;
  if (ho == 0) { ho = h; wo = w; }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / ov;
    const int v = (int)(i - pix * ov);
    const int ox = (int)(pix % wo) * stride;
    const long long t = pix / wo;
    const int oy = (int)(t % ho) * stride;
    const long long img = t / ho;
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int oc = v * 8 + j;
      const int tap = oc / c;
      const int ch = oc - tap * c;
      float val = 0.f;
      if (tap < k * k) {
        const int dy = tap / k, dx = tap - dy * k;
        const int sy = oy + dy * dil - pad, sx = ox + dx * dil - pad;
        if (sy >= 0 && sy < h && sx >= 0 && sx < w) val = to_f<T>(x[((img * h + sy) * w + sx) * cs_in + ch]);
      }
      o[j] = val;
    }
    Vec8<T>::store(y + pix * cs_out + v * 8, o);
  }
}

// ---------------------------------------------------------------------------------------------------
// Adjoint of im2col_kernel (gather form, no atomics): gx[n,y,x,ch] = sum over the taps (dy,dx) and output pixels (oy,ox) with
// oy*stride + dy*dil - pad == y, ox*stride + dx*dil - pad == x of g[n,oy,ox, tap*c + ch]; channels >= c of gx are zero.  Lets a
// first-layer conv on an image that NEEDS a data gradient (the discriminator under the generator loss: D(fake) -> G,
// discriminator.py:122, trainer.py:1421-1440) take the im2col + K = round8(k*k*c) GEMM route of the ResNet stem.
// One thread per input pixel (the 8-channel storage vector).
template <typename T>
__global__ void __launch_bounds__(256)
col2im_kernel(const T* __restrict__ g, T* __restrict__ gx, long long total_pix, int h, int w, int cs_in, int c, int k, int pad,
              int dil, int stride, int ho, int wo, int cs_col) {
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < total_pix; pix += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(pix % w);
    const long long t = pix / w;
    const int y = (int)(t % h);
    const long long img = t / h;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int dy = 0; dy < k; ++dy) {
      const int ny = y + pad - dy * dil;
      if (ny < 0 || ny % stride) continue;
      const int oy = ny / stride;
      if (oy >= ho) continue;
      for (int dx = 0; dx < k; ++dx) {
        const int nx = x + pad - dx * dil;
        if (nx < 0 || nx % stride) continue;
        const int ox = nx / stride;
        if (ox >= wo) continue;
        const T* src = g + ((img * ho + oy) * wo + ox) * cs_col + (dy * k + dx) * c;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < c) acc[j] += to_f<T>(src[j]);
      }
    }
    for (int v = 0; v < cs_in; v += 8) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = v == 0 ? acc[j] : 0.f;
      Vec8<T>::store(gx + pix * cs_in + v, o);
    }
  }
}

// im2col for image-like inputs (one 16-byte vector per pixel, C <= 4 logical channels) and small windows: ONE thread per output
// pixel loads each tap's vector once and writes the whole patch row — compile-time C and K make the tap -> output shuffle pure
// register moves.  (The generic kernel does one scalar load per output element: 8 scalar loads per 16 bytes written; at 640^2
// the SPADE conditioning patches and the discriminators' first layers made it 4.4 ms of the train step.)
template <typename T, int C, int K>
__global__ void __launch_bounds__(128)
im2col_pix_kernel(const T* __restrict__ x, T* __restrict__ y, long long total_pix, int h, int w, int pad, int stride, int ho, int wo,
                  int cs_out) {
  constexpr int KK = K * K * C;
  constexpr int NV = (KK + 7) / 8;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < total_pix; pix += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(pix % wo) * stride;
    const long long t = pix / wo;
    const int oy = (int)(t % ho) * stride;
    const long long img = t / ho;
    float o[NV * 8];
#pragma unroll
    for (int i = KK; i < NV * 8; ++i) o[i] = 0.f;
#pragma unroll
    for (int dy = 0; dy < K; ++dy) {
      const int sy = oy + dy - pad;
#pragma unroll
      for (int dx = 0; dx < K; ++dx) {
        const int sx = ox + dx - pad;
        float vec[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) vec[q] = 0.f;
        if (sy >= 0 && sy < h && sx >= 0 && sx < w) Vec8<T>::load(x + ((img * h + sy) * w + sx) * 8, vec);
#pragma unroll
        for (int ch = 0; ch < C; ++ch) o[(dy * K + dx) * C + ch] = vec[ch];
      }
    }
    T* dst = y + pix * cs_out;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float ov[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) ov[j] = o[v * 8 + j];
      Vec8<T>::store(dst + v * 8, ov);
    }
    for (int v = NV * 8; v < cs_out; v += 8) {   // (cs_out == round8(KK) in practice: nothing left)
      float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      Vec8<T>::store(dst + v, z);
    }
  }
}

template <typename T>
static bool launch_im2col_pix(const T* x, T* y, int n, int h, int w, int c, int k, int pad, int stride, int ho, int wo, int cs_out,
                              cudaStream_t st) {
  const long long total = (long long)n * ho * wo;
  const int grid = (int)((total + 127) / 128 < 148LL * 32 ? (total + 127) / 128 : 148LL * 32);
#define CGB_IM2COL_CASE(C_, K_)                                                                                              \
  if (c == C_ && k == K_) {                                                                                                 \
    im2col_pix_kernel<T, C_, K_><<<grid, 128, 0, st>>>(x, y, total, h, w, pad, stride, ho, wo, cs_out);                     \
    return true;                                                                                                           \
  }
  CGB_IM2COL_CASE(3, 3)
  CGB_IM2COL_CASE(3, 4)
  CGB_IM2COL_CASE(4, 4)
  CGB_IM2COL_CASE(2, 4)
  CGB_IM2COL_CASE(1, 4)
  CGB_IM2COL_CASE(4, 3)
  CGB_IM2COL_CASE(3, 7)   // the ResNet stem (147 -> 152 channels): 152 patch values live in registers
#undef CGB_IM2COL_CASE
  return false;
}

// ---------------------------------------------------------------------------------------------------
// compositing (generator.py:279-297)
template <typename T>
__global__ void __launch_bounds__(256)
mask_cond_kernel(const float* __restrict__ x, const float* __restrict__ m, T* __restrict__ cond, int hw,
                 int cs, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int img = (int)(i / hw);
    const int p = (int)(i - (long long)img * hw);
    const float k = 1.f - m[i];
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) v[ch] = x[((long long)img * 3 + ch) * hw + p] * k;
    T* dst = cond + i * cs;
    Vec8<T>::store(dst, v);
    for (int ch = 8; ch < cs; ++ch) dst[ch] = from_f<T>(0.f);
  }
}

__global__ void __launch_bounds__(256)
paste_fwd_kernel(const float* __restrict__ x, const float* __restrict__ m, const float* __restrict__ fake,
                 float* __restrict__ out, int hw, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pl = i / hw;  // img*3+ch
    const int p = (int)(i - pl * hw);
    const float mm = m[(pl / 3) * hw + p];
    out[i] = x[i] * (1.f - mm) + fake[i] * mm;
  }
}

__global__ void __launch_bounds__(256)
paste_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ m, float* __restrict__ gfake,
                 int hw, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pl = i / hw;
    const int p = (int)(i - pl * hw);
    gfake[i] = gout[i] * m[(pl / 3) * hw + p];
  }
}

// adjoints w.r.t. the MASK (painter loss for the masker, trainer.py:1618-1651: the mask is the masker's prediction):
//   cond = x (1 - m)                 ->  gm = - sum_c x_c gcond_c
//   out  = x (1 - m) + fake m        ->  gm =   sum_c gout_c (fake_c - x_c)
template <typename T>
__global__ void __launch_bounds__(256)
mask_cond_bwd_kernel(const float* __restrict__ x, const T* __restrict__ gcond, float* __restrict__ gm, int hw, int cs,
                     long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int img = (int)(i / hw);
    const int p = (int)(i - (long long)img * hw);
    float g[8];
    Vec8<T>::load(gcond + i * cs, g);
    float s = 0.f;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) s += x[((long long)img * 3 + ch) * hw + p] * g[ch];
    gm[i] = -s;
  }
}

__global__ void __launch_bounds__(256)
paste_bwd_mask_kernel(const float* __restrict__ gout, const float* __restrict__ x, const float* __restrict__ fake,
                      float* __restrict__ gm, int hw, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long img = i / hw;
    const int p = (int)(i - img * hw);
    float s = 0.f;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const long long j = (img * 3 + ch) * hw + p;
      s += gout[j] * (fake[j] - x[j]);
    }
    gm[i] = s;
  }
}

__global__ void __launch_bounds__(256)
l1_loss_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ loss,
               float* __restrict__ ga, long long count, float scale) {
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    s += fabsf(d);
    if (ga) ga[i] = d > 0.f ? scale : (d < 0.f ? -scale : 0.f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += ws[i];
    atomicAdd(loss, t * scale);
  }
}

// ---------------------------------------------------------------------------------------------------
// Spectral norm power iteration (norms.py:100-112): single CTA of 1024 threads, W [rows, cols] fp32.

// Spectral-norm power iteration (norms.py:100-112) as three small multi-CTA kernels (W is up to 640 x 5760):
//   K1: v_raw += W[rows-slice]^T u          grid (col chunks of 256, row slices of 64), atomics into zeroed v
//   K2: u_raw  = W (v_raw / (|v_raw|+eps))  one warp per row; every CTA recomputes |v_raw| (<= 23 KB read)
//   K3: v = v_raw/(|v_raw|+eps) ; u = u_raw/(|u_raw|+eps) ; sigma = |u_raw|^2/(|u_raw|+eps)      one CTA
__global__ void __launch_bounds__(256)
sn_wtu_kernel(const float* __restrict__ w, const float* __restrict__ u, float* __restrict__ v, int rows, int cols) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  const int r0 = blockIdx.y * 64;
  const int r1 = min(rows, r0 + 64);
  __shared__ float us[64];
  if (threadIdx.x < 64 && r0 + threadIdx.x < rows) us[threadIdx.x] = u[r0 + threadIdx.x];
  __syncthreads();
  if (c >= cols) return;
  float s = 0.f;
  for (int r = r0; r < r1; ++r) s = fmaf(w[(long long)r * cols + c], us[r - r0], s);
  atomicAdd(v + c, s);
}

__device__ __forceinline__ float block_sum_256(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < 8; ++i) t += red[i];
  return t;
}

__global__ void __launch_bounds__(256)
sn_wv_kernel(const float* __restrict__ w, const float* __restrict__ v, float* __restrict__ u, int rows, int cols) {
  __shared__ float red[8];
  float nv = 0.f;
  for (int c = threadIdx.x; c < cols; c += 256) nv = fmaf(v[c], v[c], nv);
  nv = block_sum_256(nv, red);
  const float inv_v = 1.f / (sqrtf(nv) + 1e-12f);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= rows) return;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s = fmaf(w[(long long)r * cols + c], v[c], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) u[r] = s * inv_v;
}

__global__ void __launch_bounds__(256)
sn_finalize_kernel(float* __restrict__ u, float* __restrict__ v, float* __restrict__ sigma, int rows, int cols) {
  __shared__ float red[8];
  float nv = 0.f, nu = 0.f;
  for (int c = threadIdx.x; c < cols; c += 256) nv = fmaf(v[c], v[c], nv);
  nv = block_sum_256(nv, red);
  for (int r = threadIdx.x; r < rows; r += 256) nu = fmaf(u[r], u[r], nu);
  nu = block_sum_256(nu, red);
  const float inv_v = 1.f / (sqrtf(nv) + 1e-12f);
  const float inv_u = 1.f / (sqrtf(nu) + 1e-12f);
  for (int c = threadIdx.x; c < cols; c += 256) v[c] *= inv_v;
  for (int r = threadIdx.x; r < rows; r += 256) u[r] *= inv_u;
  if (threadIdx.x == 0) sigma[0] = nu * inv_u;  // u_new . (W v_new) = |Wv|^2 / (|Wv| + eps)
}

}  // namespace cgb

// =====================================================================================================
// C ABI
// =====================================================================================================
using namespace cgb;

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

extern "C" int64_t cgb_instnorm_ws_doubles(int32_t n, int32_t hw, int32_t c) {
  if (n <= 0 || hw <= 0 || c <= 0) return 0;
  const int chunks = pick_chunks(n, hw, c / 8);
  const int ppc = (hw + chunks - 1) / chunks;
  return (int64_t)n * ((hw + ppc - 1) / ppc) * c;   // n * chunks * 2c floats
}

extern "C" int cgb_instnorm_stats(const void* x, int32_t dtype, int32_t n, int32_t hw, int32_t c,
                                  float eps, double* ws, float* mean, float* rstd, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && ws && mean && rstd, "instnorm_stats: null pointer");
  CGB_REQUIRE(c % 8 == 0 && c >= 8 && c <= 2048, "instnorm_stats: c=%d must be a multiple of 8 in [8,2048]", c);
  CGB_REQUIRE(n > 0 && hw > 0, "instnorm_stats: empty input");
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = pick_chunks(n, hw, c / 8);
  const int ppc = (hw + chunks - 1) / chunks;
  const int gx = (hw + ppc - 1) / ppc;
  dim3 grid(gx, n);
  float* partial = reinterpret_cast<float*>(ws);   // cgb_instnorm_ws_doubles(n, hw, c) doubles = n*gx*2c floats
  DISPATCH_T(dtype, in_stats_kernel<T><<<grid, 256, 256 * 16 * sizeof(float), st>>>((const T*)x, partial, hw, c, ppc);)
  int s = after_launch("in_stats");
  if (s) return s;
  in_stats_finalize_kernel<<<dim3((c + 31) / 32, n), dim3(32, 16), 0, st>>>(partial, mean, rstd, c, gx, 1.0 / (double)hw, eps);
  return after_launch("in_stats_finalize");
}

extern "C" int cgb_spade_modulate_fwd(const void* x, const float* mean, const float* rstd,
                                      const void* gb, void* out, int32_t dtype, int32_t n, int32_t hw,
                                      int32_t c, int32_t act, float slope, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && mean && rstd && gb && out, "spade_modulate_fwd: null pointer");
  CGB_REQUIRE(c % 8 == 0 && c >= 8, "spade_modulate_fwd: c=%d must be a multiple of 8", c);
  CGB_REQUIRE(c <= 2048, "spade_modulate_fwd: c=%d too large", c);
  CGB_REQUIRE(act == CGB_ACT_NONE || act == CGB_ACT_RELU || act == CGB_ACT_LRELU, "spade_modulate_fwd: act must be none/relu/lrelu");
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = pick_chunks(n, hw, c / 8);
  const int ppc = (hw + chunks - 1) / chunks;
  dim3 grid((hw + ppc - 1) / ppc, n);
  DISPATCH_T(dtype, spade_mod_fwd_kernel<T><<<grid, 256, 0, st>>>((const T*)x, mean, rstd, (const T*)gb, (T*)out, hw, c,
                                                                 ppc, act, slope);)
  return after_launch("spade_mod_fwd");
}

static int spade_modulate_bwd_impl(const void* x, const float* mean, const float* rstd, const void* gb, const void* gout, void* ggb,
                                   void* gxhat, double* sums, double* bsum, int32_t dtype, int32_t n, int32_t hw, int32_t c,
                                   int32_t act, float slope, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && mean && rstd && gb && gout && ggb && gxhat && sums, "spade_modulate_bwd: null pointer");
  CGB_REQUIRE(c % 8 == 0 && c >= 8 && c <= 2048, "spade_modulate_bwd: c=%d must be a multiple of 8 in [8,2048]", c);
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = pick_chunks(n, hw, c / 8);
  const int ppc = (hw + chunks - 1) / chunks;
  dim3 grid((hw + ppc - 1) / ppc, n);
  DISPATCH_T(dtype, spade_mod_bwd_kernel<T><<<grid, 256, (bsum ? 4 : 2) * c * sizeof(float), st>>>(
                        (const T*)x, mean, rstd, (const T*)gb, (const T*)gout, (T*)ggb, (T*)gxhat, sums, bsum,
                        hw, c, ppc, act, slope);)
  return after_launch("spade_mod_bwd");
}

extern "C" int cgb_spade_modulate_bwd(const void* x, const float* mean, const float* rstd,
                                      const void* gb, const void* gout, void* ggb, void* gxhat,
                                      double* sums, int32_t dtype, int32_t n, int32_t hw, int32_t c,
                                      int32_t act, float slope, void* stream) {
  return spade_modulate_bwd_impl(x, mean, rstd, gb, gout, ggb, gxhat, sums, nullptr, dtype, n, hw, c, act, slope, stream);
}

extern "C" int cgb_spade_modulate_bwd_bias(const void* x, const float* mean, const float* rstd,
                                           const void* gb, const void* gout, void* ggb, void* gxhat,
                                           double* sums, double* bsum, int32_t dtype, int32_t n, int32_t hw, int32_t c,
                                           int32_t act, float slope, void* stream) {
  CGB_REQUIRE(bsum, "spade_modulate_bwd_bias: null bsum");
  return spade_modulate_bwd_impl(x, mean, rstd, gb, gout, ggb, gxhat, sums, bsum, dtype, n, hw, c, act, slope, stream);
}

extern "C" int cgb_instnorm_bwd(const void* x, const float* mean, const float* rstd, const double* sums,
                                void* g, int32_t dtype, int32_t n, int32_t hw, int32_t c, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && mean && rstd && sums && g, "instnorm_bwd: null pointer");
  CGB_REQUIRE(c % 8 == 0 && c >= 8, "instnorm_bwd: c=%d must be a multiple of 8", c);
  CGB_REQUIRE(c <= 2048, "instnorm_bwd: c=%d too large", c);
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = pick_chunks(n, hw, c / 8);
  const int ppc = (hw + chunks - 1) / chunks;
  dim3 grid((hw + ppc - 1) / ppc, n);
  DISPATCH_T(dtype, in_bwd_kernel<T><<<grid, 256, 0, st>>>((const T*)x, mean, rstd, sums, (T*)g, hw, c, ppc);)
  return after_launch("in_bwd");
}

extern "C" int cgb_resize_nearest_fwd(const void* x, void* y, int32_t dtype, int32_t n, int32_t hi,
                                      int32_t wi, int32_t ho, int32_t wo, int32_t c, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && y, "resize_nearest: null pointer");
  CGB_REQUIRE(c % 8 == 0 && c >= 8, "resize_nearest: c=%d must be a multiple of 8", c);
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)n * ho * wo * (c / 8);
  DISPATCH_T(dtype, resize_nearest_kernel<T><<<grid_for(total), 256, 0, st>>>((const T*)x, (T*)y, total, hi,
                                                                             wi, ho, wo, c);)
  return after_launch("resize_nearest");
}

extern "C" int cgb_resize_nearest_bwd(const void* gy, void* gx, int32_t dtype, int32_t n, int32_t hi, int32_t wi,
                                      int32_t ho, int32_t wo, int32_t c, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(gy && gx, "resize_nearest_bwd: null pointer");
  CGB_REQUIRE(c % 8 == 0 && c >= 8 && hi > 0 && wi > 0 && ho > 0 && wo > 0, "resize_nearest_bwd: bad shape (c=%d)", c);
  const long long total = (long long)n * hi * wi * (c / 8);
  DISPATCH_T(dtype, resize_nearest_bwd_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const T*)gy, (T*)gx, total,
                                                                                                  hi, wi, ho, wo, c);)
  return after_launch("resize_nearest_bwd");
}

extern "C" int cgb_upsample_nearest_bwd(const void* gy, void* gx, int32_t dtype, int32_t n, int32_t hi,
                                        int32_t wi, int32_t f, int32_t c, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(gy && gx, "upsample_bwd: null pointer");
  CGB_REQUIRE(c % 8 == 0 && c >= 8 && f >= 1, "upsample_bwd: bad c=%d or f=%d", c, f);
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)n * hi * wi * (c / 8);
  DISPATCH_T(dtype, upsample_bwd_kernel<T><<<grid_for(total), 256, 0, st>>>((const T*)gy, (T*)gx, total, hi,
                                                                           wi, f, c);)
  return after_launch("upsample_bwd");
}

extern "C" int cgb_im2col(const void* x, void* y, int32_t dtype, int32_t n, int32_t h, int32_t w, int32_t cs_in,
                          int32_t c, int32_t k, int32_t pad, int32_t dil, int32_t cs_out, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && y, "im2col: null pointer");
  CGB_REQUIRE(cs_in % 8 == 0 && cs_out % 8 == 0 && c >= 1 && c <= cs_in && k * k * c <= cs_out,
              "im2col: bad channels cs_in=%d c=%d k=%d cs_out=%d", cs_in, c, k, cs_out);
  cudaStream_t st = (cudaStream_t)stream;
  if (cs_in == 8 && dil == 1 && (dtype == CGB_BF16 || dtype == CGB_F16)) {   // image-like input: the per-pixel kernel
    const bool done = dtype == CGB_BF16
                          ? launch_im2col_pix((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, h, w, c, k, pad, 1, h, w, cs_out, st)
                          : launch_im2col_pix((const __half*)x, (__half*)y, n, h, w, c, k, pad, 1, h, w, cs_out, st);
    if (done) return after_launch("im2col_pix");
  }
  const long long total = (long long)n * h * w * (cs_out / 8);
  DISPATCH_T(dtype, im2col_kernel<T><<<grid_for(total), 256, 0, st>>>((const T*)x, (T*)y, total, h, w, cs_in, c, k,
                                                                     pad, dil, cs_out);)
  return after_launch("im2col");
}

extern "C" int cgb_im2col_strided(const void* x, void* y, int32_t dtype, int32_t n, int32_t h, int32_t w, int32_t cs_in,
                                  int32_t c, int32_t k, int32_t pad, int32_t dil, int32_t stride, int32_t cs_out, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && y, "im2col_strided: null pointer");
  CGB_REQUIRE(cs_in % 8 == 0 && cs_out % 8 == 0 && c >= 1 && c <= cs_in && k * k * c <= cs_out && stride >= 1,
              "im2col_strided: bad arguments cs_in=%d c=%d k=%d cs_out=%d stride=%d", cs_in, c, k, cs_out, stride);
  const int ho = (h + 2 * pad - dil * (k - 1) - 1) / stride + 1, wo = (w + 2 * pad - dil * (k - 1) - 1) / stride + 1;
  if (cs_in == 8 && dil == 1 && (dtype == CGB_BF16 || dtype == CGB_F16)) {
    const bool done = dtype == CGB_BF16 ? launch_im2col_pix((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, h, w, c, k, pad, stride, ho,
                                                            wo, cs_out, (cudaStream_t)stream)
                                        : launch_im2col_pix((const __half*)x, (__half*)y, n, h, w, c, k, pad, stride, ho, wo, cs_out,
                                                            (cudaStream_t)stream);
    if (done) return after_launch("im2col_pix");
  }
  const long long total = (long long)n * ho * wo * (cs_out / 8);
  DISPATCH_T(dtype, im2col_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, total, h, w, cs_in, c,
                                                                                       k, pad, dil, cs_out, stride, ho, wo);)
  return after_launch("im2col_strided");
}

extern "C" int cgb_col2im_strided(const void* g, void* gx, int32_t dtype, int32_t n, int32_t h, int32_t w, int32_t cs_in, int32_t c,
                                  int32_t k, int32_t pad, int32_t dil, int32_t stride, int32_t cs_col, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(g && gx, "col2im_strided: null pointer");
  CGB_REQUIRE(cs_in % 8 == 0 && cs_col % 8 == 0 && c >= 1 && c <= 8 && c <= cs_in && k * k * c <= cs_col && stride >= 1,
              "col2im_strided: bad arguments cs_in=%d c=%d (<= 8) k=%d cs_col=%d stride=%d", cs_in, c, k, cs_col, stride);
  const int ho = (h + 2 * pad - dil * (k - 1) - 1) / stride + 1, wo = (w + 2 * pad - dil * (k - 1) - 1) / stride + 1;
  const long long total = (long long)n * h * w;
  DISPATCH_T(dtype, col2im_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const T*)g, (T*)gx, total, h, w, cs_in, c, k,
                                                                                       pad, dil, stride, ho, wo, cs_col);)
  return after_launch("col2im_strided");
}

extern "C" int cgb_nchw_to_nhwc(const float* x, void* y, int32_t dtype, int32_t n, int32_t c, int32_t hw,
                                int32_t cs, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && y, "nchw_to_nhwc: null pointer");
  CGB_REQUIRE(cs % 8 == 0 && cs >= c && n <= 65535, "nchw_to_nhwc: bad cs=%d for c=%d", cs, c);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((hw + 31) / 32, n, (cs + 63) / 64);
  if (cs <= 32) {
    const long long total = (long long)n * hw;
    const int g1 = grid_for(total);
    DISPATCH_T(dtype,
               if (cs == 8) nchw_to_nhwc_small_kernel<T, 8><<<g1, 256, 0, st>>>(x, (T*)y, c, hw, total);
               else if (cs == 16) nchw_to_nhwc_small_kernel<T, 16><<<g1, 256, 0, st>>>(x, (T*)y, c, hw, total);
               else if (cs == 24) nchw_to_nhwc_small_kernel<T, 24><<<g1, 256, 0, st>>>(x, (T*)y, c, hw, total);
               else nchw_to_nhwc_small_kernel<T, 32><<<g1, 256, 0, st>>>(x, (T*)y, c, hw, total);)
    return after_launch("nchw_to_nhwc_small");
  }
  DISPATCH_T(dtype, nchw_to_nhwc_kernel<T><<<grid, 256, 0, st>>>(x, (T*)y, c, hw, cs);)
  return after_launch("nchw_to_nhwc");
}

extern "C" int cgb_nhwc_to_nchw(const void* x, float* y, int32_t dtype, int32_t n, int32_t c, int32_t hw,
                                int32_t cs, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && y, "nhwc_to_nchw: null pointer");
  CGB_REQUIRE(cs % 8 == 0 && cs >= c && n <= 65535, "nhwc_to_nchw: bad cs=%d for c=%d", cs, c);
  cudaStream_t st = (cudaStream_t)stream;
  if (cs <= 32) {
    const long long total = (long long)n * hw;
    const int g1 = grid_for(total);
    DISPATCH_T(dtype,
               if (cs == 8) nhwc_to_nchw_small_kernel<T, 8><<<g1, 256, 0, st>>>((const T*)x, y, c, hw, total);
               else if (cs == 16) nhwc_to_nchw_small_kernel<T, 16><<<g1, 256, 0, st>>>((const T*)x, y, c, hw, total);
               else if (cs == 24) nhwc_to_nchw_small_kernel<T, 24><<<g1, 256, 0, st>>>((const T*)x, y, c, hw, total);
               else nhwc_to_nchw_small_kernel<T, 32><<<g1, 256, 0, st>>>((const T*)x, y, c, hw, total);)
    return after_launch("nhwc_to_nchw_small");
  }
  dim3 grid((hw + 31) / 32, n, (cs + 63) / 64);
  DISPATCH_T(dtype, nhwc_to_nchw_kernel<T><<<grid, 256, 0, st>>>((const T*)x, y, c, hw, cs);)
  return after_launch("nhwc_to_nchw");
}

extern "C" int cgb_act_bwd(const void* gy, const void* y, void* gx, int32_t dtype, int64_t count,
                           int32_t act, float slope, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(gy && y && gx, "act_bwd: null pointer");
  CGB_REQUIRE(count % 8 == 0, "act_bwd: count must be a multiple of 8");
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_T(dtype, act_bwd_kernel<T><<<grid_for(count / 8), 256, 0, st>>>((const T*)gy, (const T*)y, (T*)gx,
                                                                          count, act, slope);)
  return after_launch("act_bwd");
}

extern "C" int cgb_act_bwd_bias(const void* gy, const void* y, void* gx, float* gbias, int32_t dtype, int64_t npix, int32_t c,
                                int32_t act, float slope, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(gy && y && gx && gbias, "act_bwd_bias: null pointer");
  CGB_REQUIRE(c % 8 == 0 && c >= 8 && c <= 2048 && npix > 0, "act_bwd_bias: c=%d must be a multiple of 8 in [8,2048]", c);
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(gbias, 0, (size_t)c * sizeof(float), st);
  long long ctas = (npix * (c / 8) + 256LL * 8 - 1) / (256LL * 8);   // ~8 channel vectors per thread
  if (ctas > 148 * 8) ctas = 148 * 8;
  if (ctas < 1) ctas = 1;
  const long long ppc = (npix + ctas - 1) / ctas;
  DISPATCH_T(dtype, act_bwd_bias_kernel<T><<<(unsigned)((npix + ppc - 1) / ppc), 256, c * sizeof(float), st>>>(
                        (const T*)gy, (const T*)y, (T*)gx, gbias, (long long)npix, c, ppc, act, slope);)
  return after_launch("act_bwd_bias");
}

extern "C" int cgb_act_fwd(const void* x, void* y, int32_t dtype, int64_t count, int32_t act, float slope,
                           void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && y, "act_fwd: null pointer");
  CGB_REQUIRE(count % 8 == 0, "act_fwd: count must be a multiple of 8");
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_T(dtype, act_fwd_kernel<T><<<grid_for(count / 8), 256, 0, st>>>((const T*)x, (T*)y, count, act, slope);)
  return after_launch("act_fwd");
}

extern "C" int cgb_mask_cond(const float* x, const float* m, void* cond, int32_t dtype, int32_t n,
                             int32_t hw, int32_t cs, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && m && cond, "mask_cond: null pointer");
  CGB_REQUIRE(cs % 8 == 0 && cs >= 8, "mask_cond: bad cs=%d", cs);
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)n * hw;
  DISPATCH_T(dtype, mask_cond_kernel<T><<<grid_for(total), 256, 0, st>>>(x, m, (T*)cond, hw, cs, total);)
  return after_launch("mask_cond");
}

extern "C" int cgb_paste_fwd(const float* x, const float* m, const float* fake, float* out, int32_t n,
                             int32_t hw, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && m && fake && out, "paste_fwd: null pointer");
  const long long total = (long long)n * 3 * hw;
  paste_fwd_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, m, fake, out, hw, total);
  return after_launch("paste_fwd");
}

extern "C" int cgb_paste_bwd(const float* gout, const float* m, float* gfake, int32_t n, int32_t hw,
                             void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(gout && m && gfake, "paste_bwd: null pointer");
  const long long total = (long long)n * 3 * hw;
  paste_bwd_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(gout, m, gfake, hw, total);
  return after_launch("paste_bwd");
}

// ---------------------------------------------------------------------------------------------------
// Forward weight packing in ONE launch: OIHW fp32 (the nn.Parameter / the spectrally-normalised weight) -> [cos][kh*kw][cis]
// storage dtype, zero in the channel padding.  Replaces torch.zeros + permute/reshape copy + slice assignment (3 launches per
// weight, ~400 weights re-packed after every optimiser update: the host-side profile, scripts/host_profile_dryrun.py).
// One thread per 8 consecutive input channels of one (co, tap): reads are strided by kh*kw floats (weights are small and
// L2-resident), writes are 16-byte vectors.
template <typename T>
__global__ void __launch_bounds__(256)
pack_weight_kernel(const float* __restrict__ w, T* __restrict__ wp, long long total_vec, int o, int i, int taps, int cis) {
  const int iv = cis >> 3;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total_vec;
       idx += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(idx % iv);
    const long long r = idx / iv;
    const int t = (int)(r % taps);
    const int co = (int)(r / taps);
    float vals[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ci = v * 8 + j;
      vals[j] = (co < o && ci < i) ? w[((long long)co * i + ci) * taps + t] : 0.f;
    }
    Vec8<T>::store(wp + idx * 8, vals);
  }
}

// the forward packing AND the dgrad packing wt[cis][taps-1-t][cos] = w[co][t][ci] (cgb_conv2d_pack_dgrad_weight's layout) in one
// launch: the first total_fwd vectors belong to wp (8 consecutive ci), the rest to wt (8 consecutive co)
template <typename T>
__global__ void __launch_bounds__(256)
pack_weight_dual_kernel(const float* __restrict__ w, T* __restrict__ wp, T* __restrict__ wt, long long total_fwd, long long total_all,
                        int o, int i, int taps, int cos, int cis) {
  const int iv = cis >> 3, ov = cos >> 3;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total_all; idx += (long long)gridDim.x * blockDim.x) {
    float vals[8];
    if (idx < total_fwd) {
      const int v = (int)(idx % iv);
      const long long r = idx / iv;
      const int t = (int)(r % taps);
      const int co = (int)(r / taps);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ci = v * 8 + j;
        vals[j] = (co < o && ci < i) ? w[((long long)co * i + ci) * taps + t] : 0.f;
      }
      Vec8<T>::store(wp + idx * 8, vals);
    } else {
      const long long k = idx - total_fwd;
      const int v = (int)(k % ov);
      const long long r = k / ov;
      const int tr = (int)(r % taps);        // position in wt = taps-1-t
      const int ci = (int)(r / taps);
      const int t = taps - 1 - tr;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int co = v * 8 + j;
        vals[j] = (co < o && ci < i) ? w[((long long)co * i + ci) * taps + t] : 0.f;
      }
      Vec8<T>::store(wt + k * 8, vals);
    }
  }
}

extern "C" int cgb_pack_weight_dual(const float* w, void* wp, void* wt, int32_t dtype, int32_t o, int32_t i, int32_t taps,
                                    int32_t cos, int32_t cis, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(w && wp && wt, "pack_weight_dual: null pointer");
  CGB_REQUIRE(o >= 1 && i >= 1 && taps >= 1 && cos >= o && cis >= i && cis % 8 == 0 && cos % 8 == 0,
              "pack_weight_dual: bad shape o=%d i=%d cos=%d cis=%d", o, i, cos, cis);
  const long long total_fwd = (long long)cos * taps * (cis / 8);
  const long long total_all = total_fwd + (long long)cis * taps * (cos / 8);
  DISPATCH_T(dtype, pack_weight_dual_kernel<T><<<grid_for(total_all), 256, 0, (cudaStream_t)stream>>>(w, (T*)wp, (T*)wt, total_fwd,
                                                                                                   total_all, o, i, taps, cos, cis);)
  return after_launch("pack_weight_dual");
}

extern "C" int cgb_pack_weight(const float* w, void* wp, int32_t dtype, int32_t o, int32_t i, int32_t taps, int32_t cos,
                               int32_t cis, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(w && wp, "pack_weight: null pointer");
  CGB_REQUIRE(o >= 1 && i >= 1 && taps >= 1 && cos >= o && cis >= i && cis % 8 == 0, "pack_weight: bad shape o=%d i=%d cos=%d cis=%d",
              o, i, cos, cis);
  const long long total = (long long)cos * taps * (cis / 8);
  DISPATCH_T(dtype, pack_weight_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(w, (T*)wp, total, o, i, taps, cis);)
  return after_launch("pack_weight");
}

// ---------------------------------------------------------------------------------------------------
// DiffTransforms (climategan/transforms.py:505-626; gen.p.diff_aug, trainer.py:1079-1081, 1319-1321): differentiable augmentation
// of the images the painter discriminator sees — brightness, contrast, saturation jitter, integer translation with zero fill,
// cutout — fused into one pass over the image (the reference runs ~25 ATen kernels with two advanced-index gathers).
// NCHW fp32, c <= 8.  params[n][8] = {b, cf, sf, tx, ty, ox, oy, -}: per-sample draws made by the caller ON THE DEVICE:
//   v1 = x + b ; v2 = (v1 - mean_all(v1)) cf + mean_all(v1) ; v3 = (v2 - mean_ch(v2)) sf + mean_ch(v2)          (:505-533)
//   y(i, j) = v3(i + tx, j + ty), 0 outside the image ; y = 0 inside the cutout box                               (:546-606)
// (jitter off: b = 0, cf = sf = 1; translation off: tx = ty = 0; cutout off: cut_h = 0.)  One thread per pixel, the c channel
// planes read / written coalesced over the pixel index.  sums: per-sample fp64 totals from diff_aug_sum_kernel.
struct DiffAugGeom {
  int c, h, w, cut_h, cut_w;
};

__device__ __forceinline__ bool diffaug_cut(const float* __restrict__ pr, const DiffAugGeom& g, int i, int j) {
  if (g.cut_h <= 0) return false;
  const int a = (int)pr[5] - g.cut_h / 2, b = (int)pr[6] - g.cut_w / 2;   // the reference clamps the box's indices into the image
  return i >= max(a, 0) && i <= min(a + g.cut_h - 1, g.h - 1) && j >= max(b, 0) && j <= min(b + g.cut_w - 1, g.w - 1);
}

// mode 0: sums[n] += sum of t (the mean the contrast jitter is taken around);
// mode 1: t is the output gradient: sums[n] += sum over the output pixels that are outside the cutout and whose source pixel is
//         inside the image (the total that flows back through mean_all).
__global__ void __launch_bounds__(256)
diff_aug_sum_kernel(const float* __restrict__ t, const float* __restrict__ params, double* __restrict__ sums, DiffAugGeom g, int mode) {
  const int n = blockIdx.y;
  const float* pr = params + n * 8;
  const int tx = (int)pr[3], ty = (int)pr[4];
  const long long hw = (long long)g.h * g.w;
  const float* base = t + (long long)n * g.c * hw;
  double local = 0.0;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < hw; pix += (long long)gridDim.x * blockDim.x) {
    if (mode == 1) {
      const int i = (int)(pix / g.w), j = (int)(pix - (long long)i * g.w);
      const int si = i + tx, sj = j + ty;
      if (si < 0 || si >= g.h || sj < 0 || sj >= g.w || diffaug_cut(pr, g, i, j)) continue;
    }
    float s = 0.f;
    for (int ch = 0; ch < g.c; ++ch) s += base[ch * hw + pix];
    local += (double)s;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  __shared__ double sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) tot += sh[k];
    atomicAdd(&sums[n], tot);
  }
}

__global__ void __launch_bounds__(256)
diff_aug_fwd_kernel(const float* __restrict__ x, const float* __restrict__ params, const double* __restrict__ sums,
                    float* __restrict__ y, DiffAugGeom g) {
  const int n = blockIdx.y;
  const float* pr = params + n * 8;
  const float b = pr[0], cf = pr[1], sf = pr[2];
  const int tx = (int)pr[3], ty = (int)pr[4];
  const long long hw = (long long)g.h * g.w;
  const float mean_b = (float)(sums[n] / (double)(g.c * hw)) + b;
  const float* xb = x + (long long)n * g.c * hw;
  float* yb = y + (long long)n * g.c * hw;
  const float inv_c = 1.f / (float)g.c;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < hw; pix += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(pix / g.w), j = (int)(pix - (long long)i * g.w);
    const int si = i + tx, sj = j + ty;
    const bool live = si >= 0 && si < g.h && sj >= 0 && sj < g.w && !diffaug_cut(pr, g, i, j);
    float v[8];
    float mch = 0.f;
    if (live) {
      const long long src = (long long)si * g.w + sj;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {   // (fixed trip count + predicate: v[] stays in registers)
        if (ch < g.c) {
          const float v1 = xb[ch * hw + src] + b;
          v[ch] = (v1 - mean_b) * cf + mean_b;
          mch += v[ch];
        }
      }
      mch *= inv_c;
    }
#pragma unroll
    for (int ch = 0; ch < 8; ++ch)
      if (ch < g.c) yb[ch * hw + pix] = live ? (v[ch] - mch) * sf + mch : 0.f;
  }
}

// adjoint w.r.t. x.  gsums[n] = diff_aug_sum_kernel(mode 1) of gy.
__global__ void __launch_bounds__(256)
diff_aug_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ params, const double* __restrict__ gsums,
                    float* __restrict__ gx, DiffAugGeom g) {
  const int n = blockIdx.y;
  const float* pr = params + n * 8;
  const float cf = pr[1], sf = pr[2];
  const int tx = (int)pr[3], ty = (int)pr[4];
  const long long hw = (long long)g.h * g.w;
  const float through_mean = (1.f - cf) * (float)(gsums[n] / (double)(g.c * hw));
  const float* gyb = gy + (long long)n * g.c * hw;
  float* gxb = gx + (long long)n * g.c * hw;
  const float inv_c = 1.f / (float)g.c;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < hw; pix += (long long)gridDim.x * blockDim.x) {
    const int a = (int)(pix / g.w), bcol = (int)(pix - (long long)a * g.w);   // a source pixel; it feeds output (a - tx, bcol - ty)
    const int i = a - tx, j = bcol - ty;
    const bool live = i >= 0 && i < g.h && j >= 0 && j < g.w && !diffaug_cut(pr, g, i, j);
    float g3[8];
    float gch = 0.f;
    if (live) {
      const long long dst = (long long)i * g.w + j;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        if (ch < g.c) {
          g3[ch] = gyb[ch * hw + dst];
          gch += g3[ch];
        }
      }
    }
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      if (ch < g.c) {
        const float g2 = live ? sf * g3[ch] + (1.f - sf) * inv_c * gch : 0.f;
        gxb[ch * hw + pix] = cf * g2 + through_mean;
      }
    }
  }
}

static int diff_aug_check(const void* a, const void* b, const void* c, const void* d, int n, int ch, int h, int w, int cut_h,
                          int cut_w) {
  CGB_REQUIRE(a && b && c && d, "diff_aug: null pointer");
  CGB_REQUIRE(n > 0 && n <= 65535 && ch >= 1 && ch <= 8 && h > 0 && w > 0 && cut_h >= 0 && cut_w >= 0 && (cut_h > 0) == (cut_w > 0),
              "diff_aug: bad shape n=%d c=%d h=%d w=%d cut=%dx%d", n, ch, h, w, cut_h, cut_w);
  return CGB_OK;
}

extern "C" int cgb_diff_aug_sum(const float* t, const float* params, double* sums, int32_t n, int32_t c, int32_t h, int32_t w,
                                int32_t cut_h, int32_t cut_w, int32_t mode, void* stream) {
  CGB_CHECK_DEVICE();
  if (int s = diff_aug_check(t, params, sums, sums, n, c, h, w, cut_h, cut_w)) return s;
  CGB_REQUIRE(mode == 0 || mode == 1, "diff_aug_sum: mode %d", mode);
  int gx = grid_for((long long)h * w);
  if (gx > 148 * 2) gx = 148 * 2;
  diff_aug_sum_kernel<<<dim3(gx, n), 256, 0, (cudaStream_t)stream>>>(t, params, sums, DiffAugGeom{c, h, w, cut_h, cut_w}, mode);
  return after_launch("diff_aug_sum");
}

extern "C" int cgb_diff_aug_fwd(const float* x, const float* params, const double* sums, float* y, int32_t n, int32_t c, int32_t h,
                                int32_t w, int32_t cut_h, int32_t cut_w, void* stream) {
  CGB_CHECK_DEVICE();
  if (int s = diff_aug_check(x, params, sums, y, n, c, h, w, cut_h, cut_w)) return s;
  int gx = grid_for((long long)h * w);
  if (gx > 148 * 4) gx = 148 * 4;
  diff_aug_fwd_kernel<<<dim3(gx, n), 256, 0, (cudaStream_t)stream>>>(x, params, sums, y, DiffAugGeom{c, h, w, cut_h, cut_w});
  return after_launch("diff_aug_fwd");
}

extern "C" int cgb_diff_aug_bwd(const float* gy, const float* params, const double* gsums, float* gx_out, int32_t n, int32_t c,
                                int32_t h, int32_t w, int32_t cut_h, int32_t cut_w, void* stream) {
  CGB_CHECK_DEVICE();
  if (int s = diff_aug_check(gy, params, gsums, gx_out, n, c, h, w, cut_h, cut_w)) return s;
  int gx = grid_for((long long)h * w);
  if (gx > 148 * 4) gx = 148 * 4;
  diff_aug_bwd_kernel<<<dim3(gx, n), 256, 0, (cudaStream_t)stream>>>(gy, params, gsums, gx_out, DiffAugGeom{c, h, w, cut_h, cut_w});
  return after_launch("diff_aug_bwd");
}

extern "C" int cgb_mask_cond_bwd(const float* x, const void* gcond, float* gm, int32_t dtype, int32_t n, int32_t hw, int32_t cs,
                                 void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && gcond && gm && cs >= 8 && cs % 8 == 0, "mask_cond_bwd: bad arguments");
  const long long total = (long long)n * hw;
  DISPATCH_T(dtype, mask_cond_bwd_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, (const T*)gcond, gm, hw, cs,
                                                                                             total);)
  return after_launch("mask_cond_bwd");
}

extern "C" int cgb_paste_bwd_mask(const float* gout, const float* x, const float* fake, float* gm, int32_t n, int32_t hw,
                                  void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(gout && x && fake && gm, "paste_bwd_mask: null pointer");
  const long long total = (long long)n * hw;
  paste_bwd_mask_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(gout, x, fake, gm, hw, total);
  return after_launch("paste_bwd_mask");
}

extern "C" int cgb_l1_loss(const float* a, const float* b, float* loss, float* ga, int64_t count,
                           float scale, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(a && b && loss, "l1_loss: null pointer");
  l1_loss_kernel<<<grid_for(count), 256, 0, (cudaStream_t)stream>>>(a, b, loss, ga, count, scale);
  return after_launch("l1_loss");
}

extern "C" int cgb_spectral_power_iter(const float* w, float* u, float* v, float* sigma, int32_t rows,
                                       int32_t cols, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(w && u && v && sigma, "spectral_power_iter: null pointer");
  CGB_REQUIRE(rows > 0 && cols > 0, "spectral_power_iter: empty matrix");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(v, 0, sizeof(float) * (size_t)cols, st);
  dim3 g1((cols + 255) / 256, (rows + 63) / 64);
  sn_wtu_kernel<<<g1, 256, 0, st>>>(w, u, v, rows, cols);
  int s1 = after_launch("sn_wtu");
  if (s1) return s1;
  sn_wv_kernel<<<(rows + 7) / 8, 256, 0, st>>>(w, v, u, rows, cols);
  s1 = after_launch("sn_wv");
  if (s1) return s1;
  sn_finalize_kernel<<<1, 256, 0, st>>>(u, v, sigma, rows, cols);
  return after_launch("sn_finalize");
}

extern "C" int cgb_instnorm_apply_fwd(const void* x, const float* mean, const float* rstd, void* y, int32_t dtype,
                                      int32_t n, int32_t hw, int32_t c, int32_t act, float slope, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && mean && rstd && y, "instnorm_apply_fwd: null pointer");
  CGB_REQUIRE(c % 8 == 0 && c >= 8 && c <= 2048, "instnorm_apply_fwd: bad c=%d", c);
  CGB_REQUIRE(act == CGB_ACT_NONE || act == CGB_ACT_RELU || act == CGB_ACT_LRELU, "instnorm_apply_fwd: act must be none/relu/lrelu");
  cudaStream_t st = (cudaStream_t)stream;
  const float neg = act == CGB_ACT_NONE ? 1.f : (act == CGB_ACT_RELU ? 0.f : slope);
  const int chunks = pick_chunks(n, hw, c / 8);
  const int ppc = (hw + chunks - 1) / chunks;
  dim3 grid((hw + ppc - 1) / ppc, n);
  DISPATCH_T(dtype, in_apply_fwd_kernel<T><<<grid, 256, 0, st>>>((const T*)x, mean, rstd, (T*)y, hw, c, ppc, neg);)
  return after_launch("in_apply_fwd");
}

extern "C" int cgb_instnorm_apply_bwd(const void* x, const float* mean, const float* rstd, const void* gy, void* gxhat,
                                      double* sums, int32_t dtype, int32_t n, int32_t hw, int32_t c, int32_t act,
                                      float slope, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && mean && rstd && gy && gxhat && sums, "instnorm_apply_bwd: null pointer");
  CGB_REQUIRE(c % 8 == 0 && c >= 8 && c <= 2048, "instnorm_apply_bwd: bad c=%d", c);
  cudaStream_t st = (cudaStream_t)stream;
  const float neg = act == CGB_ACT_NONE ? 1.f : (act == CGB_ACT_RELU ? 0.f : slope);
  const int chunks = pick_chunks(n, hw, c / 8);
  const int ppc = (hw + chunks - 1) / chunks;
  dim3 grid((hw + ppc - 1) / ppc, n);
  DISPATCH_T(dtype, in_apply_bwd_kernel<T><<<grid, 256, 2 * c * sizeof(float), st>>>((const T*)x, mean, rstd, (const T*)gy,
                                                                                     (T*)gxhat, sums, hw, c, ppc, neg);)
  return after_launch("in_apply_bwd");
}

extern "C" int cgb_avgpool3s2_fwd(const void* x, void* y, int32_t dtype, int32_t n, int32_t hi, int32_t wi, int32_t c,
                                  void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && y, "avgpool3s2_fwd: null pointer");
  CGB_REQUIRE(c % 8 == 0 && c >= 8, "avgpool3s2_fwd: bad c=%d", c);
  const int ho = (hi + 2 - 3) / 2 + 1, wo = (wi + 2 - 3) / 2 + 1;
  const long long total = (long long)n * ho * wo * (c / 8);
  DISPATCH_T(dtype, (avgpool3s2_kernel<T, false><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, total,
                                                                                                  hi, wi, ho, wo, c));)
  return after_launch("avgpool3s2_fwd");
}

extern "C" int cgb_avgpool3s2_bwd(const void* gy, void* gx, int32_t dtype, int32_t n, int32_t hi, int32_t wi, int32_t c,
                                  void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(gy && gx, "avgpool3s2_bwd: null pointer");
  CGB_REQUIRE(c % 8 == 0 && c >= 8, "avgpool3s2_bwd: bad c=%d", c);
  const int ho = (hi + 2 - 3) / 2 + 1, wo = (wi + 2 - 3) / 2 + 1;
  const long long total = (long long)n * hi * wi * (c / 8);
  DISPATCH_T(dtype, (avgpool3s2_kernel<T, true><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const T*)gy, (T*)gx, total,
                                                                                                 hi, wi, ho, wo, c));)
  return after_launch("avgpool3s2_bwd");
}

extern "C" int cgb_const_target_loss(const float* x, float* loss, float* gx, int64_t count, int32_t kind, float target,
                                     float scale, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && loss, "const_target_loss: null pointer");
  CGB_REQUIRE(kind >= 0 && kind <= 4 && count > 0, "const_target_loss: bad kind %d or count", kind);
  const_target_loss_kernel<<<grid_for(count), 256, 0, (cudaStream_t)stream>>>(x, loss, gx, count, kind, target, scale, nullptr);
  return after_launch("const_target_loss");
}

extern "C" int cgb_const_target_loss_dev(const float* x, float* loss, float* gx, int64_t count, int32_t kind,
                                         const float* target_dev, float scale, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && loss && target_dev, "const_target_loss_dev: null pointer");
  CGB_REQUIRE(kind >= 0 && kind <= 4 && count > 0, "const_target_loss_dev: bad kind %d or count", kind);
  const_target_loss_kernel<<<grid_for(count), 256, 0, (cudaStream_t)stream>>>(x, loss, gx, count, kind, 0.f, scale, target_dev);
  return after_launch("const_target_loss_dev");
}

extern "C" int cgb_l1_loss_storage(const void* a, const void* b, float* loss, void* ga, int32_t dtype, int64_t count,
                                   float scale, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(a && b && loss, "l1_loss_storage: null pointer");
  CGB_REQUIRE(count % 8 == 0 && count > 0, "l1_loss_storage: count must be a positive multiple of 8");
  DISPATCH_T(dtype, l1_storage_kernel<T><<<grid_for(count / 8), 256, 0, (cudaStream_t)stream>>>((const T*)a, (const T*)b, loss,
                                                                                               (T*)ga, count, scale);)
  return after_launch("l1_loss_storage");
}

extern "C" int cgb_extra_adam(float* p, const float* g, float* m, float* v, float* c, int64_t count, float lr, float beta1,
                              float beta2, float eps, float weight_decay, int32_t step, int32_t mode, int32_t save_copy,
                              void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(p && g && m && v && c, "extra_adam: null pointer");
  CGB_REQUIRE(count > 0 && step >= 1 && (mode == 0 || mode == 1), "extra_adam: bad count/step/mode");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr * sqrt(bc2) / bc1);
  extra_adam_kernel<<<grid_for(count), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, c, count, beta1, beta2, eps, weight_decay,
                                                                       step_size, mode, save_copy);
  return after_launch("extra_adam");
}

extern "C" int cgb_maxpool2_fwd(const void* x, void* y, int32_t dtype, int32_t n, int32_t hi, int32_t wi, int32_t c,
                                void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && y, "maxpool2_fwd: null pointer");
  CGB_REQUIRE(c % 8 == 0 && c >= 8 && hi >= 2 && wi >= 2, "maxpool2_fwd: bad shape");
  const long long total = (long long)n * (hi / 2) * (wi / 2) * (c / 8);
  DISPATCH_T(dtype, maxpool2_fwd_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, total, hi, wi, c);)
  return after_launch("maxpool2_fwd");
}

extern "C" int cgb_maxpool2_bwd(const void* x, const void* y, const void* gy, void* gx, int32_t dtype, int32_t n,
                                int32_t hi, int32_t wi, int32_t c, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && y && gy && gx, "maxpool2_bwd: null pointer");
  CGB_REQUIRE(c % 8 == 0 && c >= 8 && hi % 2 == 0 && wi % 2 == 0, "maxpool2_bwd: needs even spatial size");
  const long long total = (long long)n * (hi / 2) * (wi / 2) * (c / 8);
  DISPATCH_T(dtype, maxpool2_bwd_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const T*)x, (const T*)y, (const T*)gy,
                                                                                             (T*)gx, total, hi, wi, c);)
  return after_launch("maxpool2_bwd");
}

extern "C" int cgb_vgg_preprocess_fwd(const float* x, const float* m, void* y, int32_t dtype, int32_t n, int32_t hw,
                                      void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && y, "vgg_preprocess_fwd: null pointer");
  const long long total = (long long)n * hw;
  DISPATCH_T(dtype, vgg_pre_fwd_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, m, (T*)y, hw, total);)
  return after_launch("vgg_preprocess_fwd");
}

extern "C" int cgb_vgg_preprocess_bwd(const void* gy, const float* m, float* gx, int32_t dtype, int32_t n, int32_t hw,
                                      void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(gy && gx, "vgg_preprocess_bwd: null pointer");
  const long long total = (long long)n * hw;
  DISPATCH_T(dtype, vgg_pre_bwd_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const T*)gy, m, gx, hw, total);)
  return after_launch("vgg_preprocess_bwd");
}

extern "C" int cgb_maxpool3s2_ceil_fwd(const void* x, void* y, int32_t dtype, int32_t n, int32_t hi, int32_t wi, int32_t ho,
                                       int32_t wo, int32_t c, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && y && c % 8 == 0 && c >= 8, "maxpool3s2_ceil_fwd: bad arguments");
  const long long total = (long long)n * ho * wo * (c / 8);
  DISPATCH_T(dtype, maxpool3s2_ceil_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, total, hi, wi,
                                                                                                ho, wo, c);)
  return after_launch("maxpool3s2_ceil");
}

extern "C" int cgb_maxpool3s2_fwd(const void* x, void* y, int32_t dtype, int32_t n, int32_t hi, int32_t wi, int32_t ho, int32_t wo,
                                  int32_t c, int32_t pad, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && y && c % 8 == 0 && c >= 8 && (pad == 0 || pad == 1), "maxpool3s2_fwd: bad arguments");
  const long long total = (long long)n * ho * wo * (c / 8);
  DISPATCH_T(dtype, maxpool3s2_ceil_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, total, hi, wi,
                                                                                                ho, wo, c, pad);)
  return after_launch("maxpool3s2");
}

extern "C" int cgb_resize_bilinear_fwd(const void* x, void* y, int32_t dtype, int32_t n, int32_t hi, int32_t wi, int32_t ho,
                                       int32_t wo, int32_t c, int32_t align_corners, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && y && c % 8 == 0 && c >= 8, "resize_bilinear_fwd: bad arguments");
  const long long total = (long long)n * ho * wo * (c / 8);
  DISPATCH_T(dtype, resize_bilinear_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, total, hi, wi,
                                                                                                ho, wo, c, align_corners);)
  return after_launch("resize_bilinear");
}

extern "C" int cgb_resize_bicubic_fwd(const void* x, void* y, int32_t dtype, int32_t n, int32_t hi, int32_t wi, int32_t ho,
                                      int32_t wo, int32_t c, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && y && c % 8 == 0 && c >= 8, "resize_bicubic_fwd: bad arguments");
  const long long total = (long long)n * ho * wo * (c / 8);
  DISPATCH_T(dtype, resize_bicubic_kernel<T><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, total, hi, wi,
                                                                                               ho, wo, c);)
  return after_launch("resize_bicubic");
}

extern "C" int cgb_channel_mean(const void* x, void* y, int32_t dtype, int64_t pixels, int32_t cs, int32_t c_logical,
                                void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && y && cs % 8 == 0 && c_logical >= 1 && c_logical <= cs, "channel_mean: bad arguments");
  DISPATCH_T(dtype, channel_mean_kernel<T><<<grid_for(pixels), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, pixels, cs,
                                                                                              c_logical);)
  return after_launch("channel_mean");
}

extern "C" int cgb_mul(const void* a, const void* b, void* y, int32_t dtype, int64_t count, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(a && b && y && count % 8 == 0, "mul: bad arguments");
  DISPATCH_T(dtype, mul_kernel<T><<<grid_for(count / 8), 256, 0, (cudaStream_t)stream>>>((const T*)a, (const T*)b, (T*)y, count);)
  return after_launch("mul");
}

extern "C" int cgb_make_m_cond(const void* d, const void* s, const void* xr, float* mm, void* out, int32_t dtype, int32_t n,
                               int32_t hw, int32_t ss, int32_t ns, int32_t cs_out, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(d && s && mm && out, "make_m_cond: null pointer");
  CGB_REQUIRE(ns >= 1 && ns <= ss && 1 + ns + (xr ? 3 : 0) <= cs_out && cs_out <= 24, "make_m_cond: bad channel counts");
  cudaStream_t st = (cudaStream_t)stream;
  mm_init_kernel<<<(n + 255) / 256, 256, 0, st>>>(mm, n);   // mm <- {+inf, -inf} per sample, stream-ordered (no host sync)
  dim3 g1((hw + 255) / 256 < 64 ? (hw + 255) / 256 : 64, n);
  DISPATCH_T(dtype, minmax_c0_kernel<T><<<g1, 256, 0, st>>>((const T*)d, mm, hw, 8);)
  int r = after_launch("minmax_c0");
  if (r) return r;
  const long long pixels = (long long)n * hw;
  DISPATCH_T(dtype, m_cond_kernel<T><<<grid_for(pixels), 256, 0, st>>>((const T*)d, (const T*)s, (const T*)xr, mm, (T*)out, pixels,
                                                                      hw, ss, ns, cs_out);)
  return after_launch("m_cond");
}

extern "C" int cgb_make_m_cond_bwd(const void* d, const void* out, const float* mm, const void* gout, void* gd, void* gs,
                                   int32_t dtype, int32_t n, int32_t hw, int32_t ss, int32_t ns, int32_t cs_out, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(d && out && mm && gout && gd && gs, "make_m_cond_bwd: null pointer");
  CGB_REQUIRE(ns >= 1 && ns <= ss && 1 + ns <= cs_out && cs_out <= 24 && n > 0 && hw > 0, "make_m_cond_bwd: bad channel counts");
  DISPATCH_T(dtype, m_cond_bwd_kernel<T><<<n, 1024, 0, (cudaStream_t)stream>>>((const T*)d, (const T*)out, mm, (const T*)gout,
                                                                               (T*)gd, (T*)gs, hw, ss, ns, cs_out);)
  return after_launch("m_cond_bwd");
}
