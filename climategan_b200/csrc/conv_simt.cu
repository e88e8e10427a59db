// CUDA-core implicit-GEMM convolution: the generic engine of libcgb200.
// Handles every geometry (stride, dilation, zero/reflect pad, any storage channel count that is a
// multiple of 8) in fp32 or bf16 storage with fp32 accumulation.  The tcgen05 engine (conv_tc.cu)
// takes over the shapes it supports; this file is what the rest fall back to ON THE GPU — there is
// no host path.
//
// Reference behaviour restated here: nn.Conv2d as used by Conv2dBlock (climategan/blocks.py:117-144,
// with ReflectionPad2d/ZeroPad2d :66-71), SPADE (climategan/norms.py:164-171) and SPADEResnetBlock
// (climategan/blocks.py:349-353).
#include "common.cuh"

namespace cgb {

struct ConvP {
  int n, hi, wi, ci, ho, wo, co, kh, kw, stride, dil, pad, pad_mode;
  int act;
  float slope;
  int dact;
  int res_before_act;
};

static ConvP make_p(const cgb_conv_desc* d) {
  ConvP p;
  p.n = d->n; p.hi = d->hi; p.wi = d->wi; p.ci = d->ci;
  p.ho = d->ho; p.wo = d->wo; p.co = d->co;
  p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.dil = d->dil; p.pad = d->pad;
  p.pad_mode = d->pad_mode; p.act = d->act; p.slope = d->slope; p.dact = 0; p.res_before_act = d->res_before_act;
  return p;
}

constexpr int BM = 64, BN = 64, BK = 16;

// MODE 0: fprop  (M = n*ho*wo output pixels, N = co, K = taps*ci, A gathered from x)
// MODE 1: dgrad  (M = n*hi*wi input pixels,  N = ci, K = taps*co, A gathered from gy)
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
conv_simt_kernel(ConvP p, const T* __restrict__ A, const T* __restrict__ W,
                 const float* __restrict__ bias, const T* __restrict__ residual,
                 const T* __restrict__ mask_src, T* __restrict__ Y) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];

  const int Mh = MODE == 0 ? p.ho : p.hi;
  const int Mw = MODE == 0 ? p.wo : p.wi;
  const int Sh = MODE == 0 ? p.hi : p.ho;  // source (gathered) tensor dims
  const int Sw = MODE == 0 ? p.wi : p.wo;
  const long long M = (long long)p.n * Mh * Mw;
  const int Nn = MODE == 0 ? p.co : p.ci;
  const int Kc = MODE == 0 ? p.ci : p.co;
  const int taps = p.kh * p.kw;
  const int cpt = Kc >> 3;  // 8-channel chunks per tap
  const int kchunks = taps * cpt;

  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;

  // ---- loader roles
  const bool is_a = tid < 128;
  const int lt = is_a ? tid : tid - 128;
  // A loader: one pixel, one 8-chunk
  const int a_px = lt >> 1, a_kv = lt & 1;
  int a_img = 0, a_y = 0, a_x = 0;
  const long long a_m = m0 + a_px;
  const bool a_ok = a_m < M;
  if (is_a && a_ok) {
    a_img = (int)(a_m / ((long long)Mh * Mw));
    int r = (int)(a_m - (long long)a_img * Mh * Mw);
    a_y = r / Mw;
    a_x = r - a_y * Mw;
  }
  // B loader fprop: one out-channel, one 8-chunk.  dgrad: one k (co), one 8-vector of ci.
  const int b_n = lt >> 1, b_kv = lt & 1;   // fprop
  const int b_k = lt >> 3, b_nv = lt & 7;   // dgrad

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0c = 0; k0c < kchunks; k0c += 2) {
    if (is_a) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
      const int kc = k0c + a_kv;
      if (a_ok && kc < kchunks) {
        const int tap = kc / cpt;
        const int c8 = kc - tap * cpt;
        const int dy = tap / p.kw, dx = tap - dy * p.kw;
        int sy, sx;
        bool ok = true;
        if (MODE == 0) {
          sy = a_y * p.stride - p.pad + dy * p.dil;
          sx = a_x * p.stride - p.pad + dx * p.dil;
          if (p.pad_mode == CGB_PAD_REFLECT) {
            sy = reflect_idx(sy, Sh);
            sx = reflect_idx(sx, Sw);
          } else {
            ok = (sy >= 0) && (sy < Sh) && (sx >= 0) && (sx < Sw);
          }
        } else {
          int ty_ = a_y + p.pad - dy * p.dil;
          int tx_ = a_x + p.pad - dx * p.dil;
          ok = (ty_ >= 0) && (tx_ >= 0);
          sy = ty_ / p.stride;
          sx = tx_ / p.stride;
          ok = ok && (sy * p.stride == ty_) && (sx * p.stride == tx_) && (sy < Sh) && (sx < Sw);
        }
        if (ok) Vec8<T>::load(A + (((long long)a_img * Sh + sy) * Sw + sx) * Kc + c8 * 8, v);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) As[a_kv * 8 + j][a_px] = v[j];
    } else {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
      if (MODE == 0) {
        const int kc = k0c + b_kv;
        const int nn = n0 + b_n;
        if (nn < Nn && kc < kchunks) {
          // W[co][tap][ci]: chunk kc enumerates (tap, c8) in exactly that order
          Vec8<T>::load(W + (long long)nn * taps * Kc + (long long)kc * 8, v);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) Bs[b_kv * 8 + j][b_n] = v[j];
      } else {
        const int kc = k0c + (b_k >> 3);
        const int nn = n0 + b_nv * 8;
        if (nn < Nn && kc < kchunks) {
          const int tap = kc / cpt;
          const int co = (kc - tap * cpt) * 8 + (b_k & 7);
          Vec8<T>::load(W + ((long long)co * taps + tap) * Nn + nn, v);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) Bs[b_k][b_nv * 8 + j] = v[j];
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue
  const int nn = n0 + tx * 4;
  if (nn >= Nn) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j];
    const long long off = m * Nn + nn;
    if (MODE == 0) {
      if (bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += bias[nn + j];
      }
      if (residual && p.res_before_act) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += to_f<T>(residual[off + j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = act_apply(v[j], p.act, p.slope);
      if (residual && !p.res_before_act) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += to_f<T>(residual[off + j]);
      }
    } else {
      if (mask_src) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          v[j] *= act_grad_from_out(to_f<T>(mask_src[off + j]), p.dact, p.slope);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) Y[off + j] = from_f<T>(v[j]);
  }
}

// wgrad: gw[co][tap][ci] += sum_{output pixels} gy[m,co] * x[src(m,tap),ci]
// grid = (ci tiles, co tiles, taps*splits); reduction over this CTA's pixel range, fp32 atomics at the end.
template <typename T>
__global__ void __launch_bounds__(256)
wgrad_simt_kernel(ConvP p, const T* __restrict__ X, const T* __restrict__ GY, float* __restrict__ GW,
                  float* __restrict__ GB, int splits, long long px_per_split) {
  __shared__ float Gs[BK][BM + 4];
  __shared__ float Xs[BK][BN + 4];
  const int taps = p.kh * p.kw;
  const int tap = blockIdx.z / splits;
  const int split = blockIdx.z - tap * splits;
  const int dy = tap / p.kw, dx = tap - dy * p.kw;
  const long long M = (long long)p.n * p.ho * p.wo;
  const long long m_begin = (long long)split * px_per_split;
  long long m_end = m_begin + px_per_split;
  if (m_end > M) m_end = M;
  const int ci0 = blockIdx.x * BN;
  const int co0 = blockIdx.y * BM;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const bool is_g = tid < 128;
  const int lt = is_g ? tid : tid - 128;
  const int l_k = lt >> 3, l_v = lt & 7;
  const bool do_bias = (GB != nullptr) && tap == 0 && blockIdx.x == 0 && tx == 0;

  float acc[4][4];
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long mb = m_begin; mb < m_end; mb += BK) {
    const long long m = mb + l_k;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (is_g) {
      const int co = co0 + l_v * 8;
      if (m < m_end && co < p.co) Vec8<T>::load(GY + m * p.co + co, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) Gs[l_k][l_v * 8 + j] = v[j];
    } else {
      const int ci = ci0 + l_v * 8;
      if (m < m_end && ci < p.ci) {
        const int img = (int)(m / ((long long)p.ho * p.wo));
        const int r = (int)(m - (long long)img * p.ho * p.wo);
        const int oy = r / p.wo, ox = r - oy * p.wo;
        int sy = oy * p.stride - p.pad + dy * p.dil;
        int sx = ox * p.stride - p.pad + dx * p.dil;
        bool ok = true;
        if (p.pad_mode == CGB_PAD_REFLECT) {
          sy = reflect_idx(sy, p.hi);
          sx = reflect_idx(sx, p.wi);
        } else {
          ok = (sy >= 0) && (sy < p.hi) && (sx >= 0) && (sx < p.wi);
        }
        if (ok) Vec8<T>::load(X + (((long long)img * p.hi + sy) * p.wi + sx) * p.ci + ci, v);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) Xs[l_k][l_v * 8 + j] = v[j];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&Gs[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Xs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      if (do_bias) {
#pragma unroll
        for (int i = 0; i < 4; ++i) bsum[i] += a[i];
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= p.co) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + tx * 4 + j;
      if (ci < p.ci) atomicAdd(GW + ((long long)co * taps + tap) * p.ci + ci, acc[i][j]);
    }
    if (do_bias) atomicAdd(GB + co, bsum[i]);
  }
}

// ---- host launchers -----------------------------------------------------------------------------------
int conv_simt_fwd(const cgb_conv_desc* d, const void* x, const void* w, const float* bias,
                  const void* residual, void* y, cudaStream_t st) {
  ConvP p = make_p(d);
  const long long M = (long long)p.n * p.ho * p.wo;
  dim3 grid(ceil_div(M, BM), ceil_div(p.co, BN));
  if (d->dtype == CGB_F32)
    conv_simt_kernel<float, 0><<<grid, 256, 0, st>>>(p, (const float*)x, (const float*)w, bias,
                                                     (const float*)residual, nullptr, (float*)y);
  else if (d->dtype == CGB_F16)
    conv_simt_kernel<__half, 0><<<grid, 256, 0, st>>>(p, (const __half*)x, (const __half*)w, bias, (const __half*)residual,
                                                      nullptr, (__half*)y);
  else
    conv_simt_kernel<__nv_bfloat16, 0><<<grid, 256, 0, st>>>(
        p, (const __nv_bfloat16*)x, (const __nv_bfloat16*)w, bias, (const __nv_bfloat16*)residual,
        nullptr, (__nv_bfloat16*)y);
  return after_launch("conv_simt_fwd");
}

int conv_simt_dgrad(const cgb_conv_desc* d, const void* gy, const void* w, int dact,
                    const void* mask_src, void* gx, cudaStream_t st) {
  ConvP p = make_p(d);
  p.dact = dact;
  const long long M = (long long)p.n * p.hi * p.wi;
  dim3 grid(ceil_div(M, BM), ceil_div(p.ci, BN));
  if (d->dtype == CGB_F32)
    conv_simt_kernel<float, 1><<<grid, 256, 0, st>>>(p, (const float*)gy, (const float*)w, nullptr,
                                                     nullptr, (const float*)mask_src, (float*)gx);
  else if (d->dtype == CGB_F16)
    conv_simt_kernel<__half, 1><<<grid, 256, 0, st>>>(p, (const __half*)gy, (const __half*)w, nullptr, nullptr,
                                                      (const __half*)mask_src, (__half*)gx);
  else
    conv_simt_kernel<__nv_bfloat16, 1><<<grid, 256, 0, st>>>(
        p, (const __nv_bfloat16*)gy, (const __nv_bfloat16*)w, nullptr, nullptr,
        (const __nv_bfloat16*)mask_src, (__nv_bfloat16*)gx);
  return after_launch("conv_simt_dgrad");
}

int conv_simt_wgrad(const cgb_conv_desc* d, const void* x, const void* gy, float* gw, float* gbias,
                    cudaStream_t st) {
  ConvP p = make_p(d);
  const long long M = (long long)p.n * p.ho * p.wo;
  const int taps = p.kh * p.kw;
  const int tiles = ceil_div(p.ci, BN) * ceil_div(p.co, BM) * taps;
  // aim for ~8 waves of 148 SMs, at least 256 pixels per split
  int splits = (148 * 8 + tiles - 1) / tiles;
  long long max_splits = (M + 255) / 256;
  if (splits > max_splits) splits = (int)max_splits;
  if (splits < 1) splits = 1;
  if ((long long)taps * splits > 65535) splits = 65535 / taps;
  long long pps = (M + splits - 1) / splits;
  pps = (pps + BK - 1) / BK * BK;
  splits = (int)((M + pps - 1) / pps);
  dim3 grid(ceil_div(p.ci, BN), ceil_div(p.co, BM), taps * splits);
  if (d->dtype == CGB_F32)
    wgrad_simt_kernel<float><<<grid, 256, 0, st>>>(p, (const float*)x, (const float*)gy, gw, gbias,
                                                   splits, pps);
  else if (d->dtype == CGB_F16)
    wgrad_simt_kernel<__half><<<grid, 256, 0, st>>>(p, (const __half*)x, (const __half*)gy, gw, gbias, splits, pps);
  else
    wgrad_simt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p, (const __nv_bfloat16*)x,
                                                           (const __nv_bfloat16*)gy, gw, gbias,
                                                           splits, pps);
  return after_launch("conv_simt_wgrad");
}

}  // namespace cgb
