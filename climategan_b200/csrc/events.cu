// Inference compositing of Trainer.infer_all: wildfire (climategan/fire.py:68-127), smog
// (trainer.py:1879-1939, HazeRD model) and the numpy/uint8 output edge (trainer.py:312-327).  All HBM-bound, NCHW fp32
// images at the API edge (what the reference passes around), one thread per pixel, coalesced over the pixel index.
#include "common.cuh"

namespace cgb {

static inline int grid_for(long long work, int block = 256) {
  long long g = (work + block - 1) / block;
  static const long long cap = 148LL * (getenv("CGB_FLAT_CTAS") ? atoi(getenv("CGB_FLAT_CTAS")) : 16);
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
  int* a = reinterpret_cast<int*>(addr);
  int old = *a;
  while (__int_as_float(old) > v) {
    const int assumed = old;
    old = atomicCAS(a, assumed, __float_as_int(v));
    if (old == assumed) break;
  }
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  int* a = reinterpret_cast<int*>(addr);
  int old = *a;
  while (__int_as_float(old) < v) {
    const int assumed = old;
    old = atomicCAS(a, assumed, __float_as_int(v));
    if (old == assumed) break;
  }
}

// per-sample min / max of x[n][count]  (tutils.normalize :567-576)
__global__ void minmax_init_kernel(float* mm, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { mm[2 * i] = 3.4e38f; mm[2 * i + 1] = -3.4e38f; }
}
__global__ void __launch_bounds__(256)
minmax_kernel(const float* __restrict__ x, float* __restrict__ mm, long long count) {
  const int img = blockIdx.y;
  const float* p = x + (long long)img * count;
  float lo = 3.4e38f, hi = -3.4e38f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const float v = p[i];
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomic_min_float(&mm[2 * img], lo);
    atomic_max_float(&mm[2 * img + 1], hi);
  }
}

__device__ __forceinline__ float trunc_u8(float v) { return floorf(fminf(fmaxf(v, 0.f), 255.f)); }

// fire.py:80-87: normalize(x, 0, 255) ; R += 40, G -= 10, B -= 20 ; clamp ; to uint8 ; + sum of the uint8 grayscale per image
__global__ void __launch_bounds__(256)
fire_tone_kernel(const float* __restrict__ x, const float* __restrict__ mm, float* __restrict__ out, double* __restrict__ gray_sum,
                 int hw) {
  const int img = blockIdx.y;
  const float lo = mm[2 * img], den = mm[2 * img + 1] - lo;
  const float* p = x + (long long)img * 3 * hw;
  float* o = out + (long long)img * 3 * hw;
  double local = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
    const float r = trunc_u8(255.f * ((p[i] - lo) / den) + 40.f);
    const float g = trunc_u8(255.f * ((p[hw + i] - lo) / den) - 10.f);
    const float b = trunc_u8(255.f * ((p[2 * hw + i] - lo) / den) - 20.f);
    o[i] = r; o[hw + i] = g; o[2 * hw + i] = b;
    local += (double)floorf(0.2989f * r + 0.587f * g + 0.114f * b);   // rgb_to_grayscale(...).to(uint8)
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) local += __shfl_xor_sync(0xffffffffu, local, s);
  if ((threadIdx.x & 31) == 0) atomicAdd(&gray_sum[img], local);
}

// adjust_contrast(img, c) then adjust_brightness(img, b) on uint8 (torchvision _blend: clamp + truncation), in place
__global__ void __launch_bounds__(256)
fire_contrast_brightness_kernel(float* __restrict__ img, const double* __restrict__ gray_sum, int hw, float contrast,
                                float brightness) {
  const int im = blockIdx.y;
  const float mean = (float)(gray_sum[im] / (double)hw);
  float* p = img + (long long)im * 3 * hw;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * hw; i += gridDim.x * blockDim.x) {
    float v = trunc_u8(contrast * p[i] + (1.f - contrast) * mean);
    p[i] = trunc_u8(brightness * v);
  }
}

// retrieve_sky_mask (tutils.py:579-597): argmax over classes == sky_idx ; fire.py:95-97 crops the bottom third
__global__ void __launch_bounds__(256)
sky_mask_kernel(const float* __restrict__ seg, float* __restrict__ out, int c, int hs, int ws, int sky_idx, int crop_row, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int hw = hs * ws;
    const long long img = i / hw;
    const int px = (int)(i - img * hw);
    const float* p = seg + img * c * hw + px;
    int best = 0;
    float bv = p[0];
    for (int k = 1; k < c; ++k) {
      const float v = p[(long long)k * hw];
      if (v > bv) { bv = v; best = k; }
    }
    out[i] = (best == sky_idx && (px / ws) < crop_row) ? 1.f : 0.f;
  }
}

// F.interpolate(mode="nearest") of single-channel planes [n,hi,wi] -> [n,ho,wo]
__global__ void __launch_bounds__(256)
plane_nearest_kernel(const float* __restrict__ x, float* __restrict__ y, int hi, int wi, int ho, int wo, long long total) {
  const float sh = (float)hi / (float)ho, sw = (float)wi / (float)wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % wo);
    const int oy = (int)((i / wo) % ho);
    const long long img = i / ((long long)wo * ho);
    const int sy = min((int)floorf(oy * sh), hi - 1), sx = min((int)floorf(ox * sw), wi - 1);
    y[i] = x[(img * hi + sy) * wi + sx];
  }
}

// increase_sky_mask (fire.py:15-47) on a binary mask = binary dilation by a (2rx+1) x (2ry+1) box, one axis per launch
__global__ void __launch_bounds__(256)
box_dilate_kernel(const float* __restrict__ x, float* __restrict__ y, int h, int w, int radius, int horizontal, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int px = (int)(i % w);
    const int py = (int)((i / w) % h);
    const float* plane = x + (i / ((long long)w * h)) * (long long)w * h;
    float acc = 0.f;
    if (horizontal) {
      const int a = max(0, px - radius), b = min(w - 1, px + radius);
      for (int k = a; k <= b; ++k) acc += plane[(long long)py * w + k];
    } else {
      const int a = max(0, py - radius), b = min(h - 1, py + radius);
      for (int k = a; k <= b; ++k) acc += plane[(long long)k * w + px];
    }
    y[i] = acc >= 1.f ? 1.f : acc;
  }
}

// kornia filter2d with a normalised Gaussian (fire.py:101-111): separable, reflect border, one axis per launch
__global__ void __launch_bounds__(256)
gauss_blur_kernel(const float* __restrict__ x, float* __restrict__ y, int h, int w, int ksize, float sigma, int horizontal,
                  long long total) {
  extern __shared__ float wgt[];  // [ksize]
  __shared__ float norm;
  const int half = ksize / 2;
  for (int k = threadIdx.x; k < ksize; k += blockDim.x) {
    float d = (float)(k - half);
    if ((ksize & 1) == 0) d += 0.5f;
    wgt[k] = expf(-d * d / (2.f * sigma * sigma));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < ksize; ++k) s += wgt[k];
    norm = 1.f / s;
  }
  __syncthreads();
  const float nrm = norm;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int px = (int)(i % w);
    const int py = (int)((i / w) % h);
    const float* plane = x + (i / ((long long)w * h)) * (long long)w * h;
    float acc = 0.f;
    if (horizontal) {
      for (int k = 0; k < ksize; ++k) acc = fmaf(wgt[k], plane[(long long)py * w + reflect_idx(px + k - half, w)], acc);
    } else {
      for (int k = 0; k < ksize; ++k) acc = fmaf(wgt[k], plane[(long long)reflect_idx(py + k - half, h) * w + px], acc);
    }
    y[i] = acc * nrm;
  }
}

// fire.py:113-125: paste the orange filter through the blurred sky mask, truncate to uint8, brightness, dummy corner pixels
__global__ void __launch_bounds__(256)
fire_paste_kernel(const float* __restrict__ img, const float* __restrict__ sky, float* __restrict__ out, int h, int w, float fr,
                  float fg, float fb, float transparency, float brightness) {
  const int im = blockIdx.y;
  const int hw = h * w;
  const float* p = img + (long long)im * 3 * hw;
  const float* s = sky + (long long)im * hw;
  float* o = out + (long long)im * 3 * hw;
  const float f[3] = {fr, fg, fb};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
    const float m = transparency / 255.f * s[i];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = m * f[c] + (1.f - m) * p[c * hw + i];
      v = trunc_u8(v);                 // .to(torch.uint8)
      v = trunc_u8(brightness * v);    // adjust_brightness(., 0.8)
      if (i == 0) v = 255.f;           // "dummy pixels to fool scaling and preserve range"
      if (i == hw - 1) v = 0.f;
      o[c * hw + i] = v;
    }
  }
}

// compute_smog (trainer.py:1879-1939): HazeRD transmission model on the linearised image + yellow filter
__device__ __forceinline__ float smog_depth(float d, float dlo, float dden) {
  const float dn = 0.3f + 0.7f * ((d - dlo) / dden);   // normalize(d, 0.3, 1.0)
  const float r = 1.f / dn;
  const float rlo = 1.f / (0.3f + 0.7f), rhi = 1.f / 0.3f;   // per-sample min / max of 1/dn
  return 0.1f + 0.9f * ((r - rlo) / (rhi - rlo));      // normalize(1/d, 0.1, 1)
}
__global__ void __launch_bounds__(256)
smog_kernel(const float* __restrict__ x, const float* __restrict__ mmx, const float* __restrict__ d, const float* __restrict__ mmd,
            float* __restrict__ out, int h, int w, int hd, int wd, float airlight, float beta, float alpha, float yr, float yg,
            float yb) {
  const int im = blockIdx.y;
  const int hw = h * w;
  const float xlo = mmx[2 * im], xden = mmx[2 * im + 1] - xlo;
  const float dlo = mmd[2 * im], dden = mmd[2 * im + 1] - dlo;
  const float sh = h > 1 ? (float)(hd - 1) / (float)(h - 1) : 0.f, sw = w > 1 ? (float)(wd - 1) / (float)(w - 1) : 0.f;
  const float* dp = d + (long long)im * hd * wd;
  const float yel[3] = {yr, yg, yb};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
    const int px = i % w, py = i / w;
    const float fy = sh * py, fx = sw * px;
    const int y0 = min((int)fy, hd - 1), x0 = min((int)fx, wd - 1);
    const int y1 = min(y0 + 1, hd - 1), x1 = min(x0 + 1, wd - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float d00 = smog_depth(dp[y0 * wd + x0], dlo, dden), d01 = smog_depth(dp[y0 * wd + x1], dlo, dden);
    const float d10 = smog_depth(dp[y1 * wd + x0], dlo, dden), d11 = smog_depth(dp[y1 * wd + x1], dlo, dden);
    const float dd = (1.f - ly) * ((1.f - lx) * d00 + lx * d01) + ly * ((1.f - lx) * d10 + lx * d11);
    const float t = expf(-beta * dd);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = (x[((long long)im * 3 + c) * hw + i] - xlo) / xden;                       // normalize(x)
      const float lin = v <= 0.04045f ? v / 12.92f : powf((v + 0.055f) / 1.055f, 2.4f);         // srgb2lrgb
      const float sm = t * lin + (1.f - t) * airlight;
      const float srgb = sm <= 0.0031308f ? 12.92f * sm : 1.055f * powf(sm, 1.f / 2.4f) - 0.055f;   // lrgb2srgb
      out[((long long)im * 3 + c) * hw + i] = srgb * (1.f - alpha) + yel[c] * alpha;
    }
  }
}

// trainer.py:312-327: normalize(t) per sample -> NHWC -> (t * 255).astype(uint8)
__global__ void __launch_bounds__(256)
to_uint8_nhwc_kernel(const float* __restrict__ x, const float* __restrict__ mm, uint8_t* __restrict__ out, int hw) {
  const int im = blockIdx.y;
  const float lo = mm[2 * im], den = mm[2 * im + 1] - lo;
  const float* p = x + (long long)im * 3 * hw;
  uint8_t* o = out + (long long)im * 3 * hw;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = ((p[c * hw + i] - lo) / den) * 255.f;
      o[i * 3 + c] = (uint8_t)fminf(fmaxf(v, 0.f), 255.f);
    }
  }
}

// ((mask > bin_value) * 255).astype(uint8)   (trainer.py:330-332)
__global__ void __launch_bounds__(256)
mask_to_uint8_kernel(const float* __restrict__ m, uint8_t* __restrict__ out, float bin_value, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    out[i] = m[i] > bin_value ? 255 : 0;
}

// rand_perlin_2d (tutils.py:648-686): gradients from (res0+1) x (res1+1) random angles, quintic fade, sqrt(2) scale.
// grid coordinate of pixel i along an axis = fmod(float(i * (res / size)), 1) exactly as torch.arange(0, res, res/size) % 1
// (ATen evaluates start + i*step in double and rounds to float); cell = i / (size / res) (repeat_interleave of the gradients).
__global__ void __launch_bounds__(256)
perlin_kernel(const float* __restrict__ angles, float* __restrict__ out, int h, int w, int res0, int res1) {
  const double step0 = (double)res0 / (double)h, step1 = (double)res1 / (double)w;
  const int d0 = h / res0, d1 = w / res1;
  const long long total = (long long)h * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int px = (int)(i % w), py = (int)(i / w);
    const float gy = fmodf((float)((double)py * step0), 1.f), gx = fmodf((float)((double)px * step1), 1.f);
    const int cy = min(py / d0, res0 - 1), cx = min(px / d1, res1 - 1);
    float n[2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const float ang = angles[(cy + a) * (res1 + 1) + (cx + b)];
        n[a][b] = (gy - (float)a) * cosf(ang) + (gx - (float)b) * sinf(ang);
      }
    const float t0 = gy * gy * gy * (gy * (gy * 6.f - 15.f) + 10.f);
    const float t1 = gx * gx * gx * (gx * (gx * 6.f - 15.f) + 10.f);
    const float l0 = n[0][0] + t0 * (n[1][0] - n[0][0]);   // lerp(n00, n10, t[...,0])
    const float l1 = n[0][1] + t0 * (n[1][1] - n[0][1]);   // lerp(n01, n11, t[...,0])
    out[i] = 1.4142135623730951f * (l0 + t1 * (l1 - l0));
  }
}

// paint_cloudy's conditioning image (generator.py:318-325, tutils.mix_noise :689-694): sky = argmax(bilinear(s)) == sky_idx ;
// y = sky * (weight * (noise - min noise) + (1 - weight) * x) + (1 - sky) * x
__global__ void __launch_bounds__(256)
cloudy_mix_kernel(const float* __restrict__ x, const float* __restrict__ seg, const float* __restrict__ noise,
                  const float* __restrict__ mm_noise, float* __restrict__ out, int h, int w, int c, int hs, int ws, int sky_idx,
                  float weight) {
  const int im = blockIdx.y;
  const int hw = h * w;
  const float sh = (float)hs / (float)h, sw = (float)ws / (float)w;
  const float nmin = mm_noise[0];
  const float* sp = seg + (long long)im * c * hs * ws;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
    const int px = i % w, py = i / w;
    const float fy = fmaxf(sh * (py + 0.5f) - 0.5f, 0.f), fx = fmaxf(sw * (px + 0.5f) - 0.5f, 0.f);   // align_corners=False
    const int y0 = min((int)fy, hs - 1), x0 = min((int)fx, ws - 1);
    const int y1 = min(y0 + 1, hs - 1), x1 = min(x0 + 1, ws - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    int best = 0;
    float bv = -3.4e38f;
    for (int k = 0; k < c; ++k) {
      const float* q = sp + (long long)k * hs * ws;
      const float v = (1.f - ly) * ((1.f - lx) * q[y0 * ws + x0] + lx * q[y0 * ws + x1]) +
                      ly * ((1.f - lx) * q[y1 * ws + x0] + lx * q[y1 * ws + x1]);
      if (v > bv) { bv = v; best = k; }
    }
    const float sky = best == sky_idx ? 1.f : 0.f;
    const float nz = noise[i] - nmin;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const long long o = ((long long)im * 3 + ch) * hw + i;
      const float xv = x[o];
      out[o] = sky * (weight * nz + (1.f - weight) * xv) + (1.f - sky) * xv;
    }
  }
}

}  // namespace cgb

using namespace cgb;

extern "C" int cgb_minmax_per_sample(const float* x, float* mm, int32_t n, int64_t count, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && mm && n > 0 && count > 0, "minmax_per_sample: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  minmax_init_kernel<<<(n + 255) / 256, 256, 0, st>>>(mm, n);
  int s = after_launch("minmax_init");
  if (s) return s;
  int gx = grid_for(count);
  if (gx > 148 * 4) gx = 148 * 4;
  minmax_kernel<<<dim3(gx, n), 256, 0, st>>>(x, mm, count);
  return after_launch("minmax");
}

extern "C" int cgb_fire_tone(const float* x, const float* mm, float* out, double* gray_sum, int32_t n, int32_t hw, float contrast,
                             float brightness, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && mm && out && gray_sum && n > 0 && hw > 0, "fire_tone: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(gray_sum, 0, sizeof(double) * n, st);
  int gx = grid_for(hw);
  if (gx > 148 * 2) gx = 148 * 2;
  fire_tone_kernel<<<dim3(gx, n), 256, 0, st>>>(x, mm, out, gray_sum, hw);
  int s = after_launch("fire_tone");
  if (s) return s;
  fire_contrast_brightness_kernel<<<dim3(gx, n), 256, 0, st>>>(out, gray_sum, hw, contrast, brightness);
  return after_launch("fire_contrast_brightness");
}

extern "C" int cgb_sky_mask(const float* seg, float* out, int32_t n, int32_t c, int32_t hs, int32_t ws, int32_t sky_idx,
                            int32_t crop_bottom, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(seg && out && n > 0 && c > 0 && hs > 0 && ws > 0, "sky_mask: bad arguments");
  const long long total = (long long)n * hs * ws;
  sky_mask_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(seg, out, c, hs, ws, sky_idx, crop_bottom ? 2 * hs / 3 : hs,
                                                                    total);
  return after_launch("sky_mask");
}

extern "C" int cgb_plane_resize_nearest(const float* x, float* y, int32_t n, int32_t hi, int32_t wi, int32_t ho, int32_t wo,
                                        void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && y && n > 0, "plane_resize_nearest: bad arguments");
  const long long total = (long long)n * ho * wo;
  plane_nearest_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, y, hi, wi, ho, wo, total);
  return after_launch("plane_resize_nearest");
}

extern "C" int cgb_box_dilate(const float* x, float* tmp, float* y, int32_t n, int32_t h, int32_t w, int32_t radius_w,
                              int32_t radius_h, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && tmp && y && n > 0 && radius_w >= 0 && radius_h >= 0, "box_dilate: bad arguments");
  const long long total = (long long)n * h * w;
  cudaStream_t st = (cudaStream_t)stream;
  box_dilate_kernel<<<grid_for(total), 256, 0, st>>>(x, tmp, h, w, radius_w, 1, total);
  int s = after_launch("box_dilate(w)");
  if (s) return s;
  box_dilate_kernel<<<grid_for(total), 256, 0, st>>>(tmp, y, h, w, radius_h, 0, total);
  return after_launch("box_dilate(h)");
}

extern "C" int cgb_gauss_blur(const float* x, float* tmp, float* y, int32_t n, int32_t h, int32_t w, int32_t ksize, float sigma,
                              void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && tmp && y && n > 0 && ksize >= 1 && ksize <= 4096 && sigma > 0.f, "gauss_blur: bad arguments");
  CGB_REQUIRE(ksize / 2 < h && ksize / 2 < w, "gauss_blur: reflect border needs kernel_size/2 (%d) < image size (%dx%d)", ksize / 2, h, w);
  const long long total = (long long)n * h * w;
  cudaStream_t st = (cudaStream_t)stream;
  // the reference's dense 2-D kernel is the outer product of two normalised 1-D Gaussians: two 1-D passes (rows, then columns)
  gauss_blur_kernel<<<grid_for(total), 256, ksize * sizeof(float), st>>>(x, tmp, h, w, ksize, sigma, 1, total);
  int s = after_launch("gauss_blur(w)");
  if (s) return s;
  gauss_blur_kernel<<<grid_for(total), 256, ksize * sizeof(float), st>>>(tmp, y, h, w, ksize, sigma, 0, total);
  return after_launch("gauss_blur(h)");
}

extern "C" int cgb_fire_paste(const float* img, const float* sky, float* out, int32_t n, int32_t h, int32_t w, float fr, float fg,
                              float fb, float transparency, float brightness, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(img && sky && out && n > 0, "fire_paste: bad arguments");
  int gx = grid_for((long long)h * w);
  if (gx > 148 * 2) gx = 148 * 2;
  fire_paste_kernel<<<dim3(gx, n), 256, 0, (cudaStream_t)stream>>>(img, sky, out, h, w, fr, fg, fb, transparency, brightness);
  return after_launch("fire_paste");
}

extern "C" int cgb_smog(const float* x, const float* mmx, const float* d, const float* mmd, float* out, int32_t n, int32_t h,
                        int32_t w, int32_t hd, int32_t wd, float airlight, float beta, float alpha, float yr, float yg, float yb,
                        void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && mmx && d && mmd && out && n > 0, "smog: bad arguments");
  int gx = grid_for((long long)h * w);
  if (gx > 148 * 2) gx = 148 * 2;
  smog_kernel<<<dim3(gx, n), 256, 0, (cudaStream_t)stream>>>(x, mmx, d, mmd, out, h, w, hd, wd, airlight, beta, alpha, yr, yg, yb);
  return after_launch("smog");
}

// ---------------------------------------------------------------------------------------------------
// Input edge of apply_events.py (resize_and_crop :211-241, to_m1_p1 :179-195; transforms.PrepareInference :292-360): a uint8 HWC
// photograph of any size -> anti-aliased bilinear resize (triangle filter widened by the down-scale factor: what
// F.interpolate(mode="bilinear", antialias=True, align_corners=False) computes) to (rh, rw) -> crop of size (th, tw) at
// (top, left) -> optional truncation to uint8 (the reference quantises the resized image, :231) -> (v/255 - 0.5)*2 written
// into image slot `img` of an NCHW fp32 batch.  One thread per output pixel; the taps of a pixel are read as 3-byte HWC
// triples (neighbouring threads read neighbouring source pixels).  The CPU did this in the reference (skimage).
__global__ void __launch_bounds__(256)
resize_crop_u8_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int h, int w, int rh, int rw, int top, int left,
                      int th, int tw, int quantize) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= th * tw) return;
  const int oy = idx / tw, ox = idx - oy * tw;
  const float sy = (float)h / (float)rh, sx = (float)w / (float)rw;
  const float supy = fmaxf(sy, 1.f), supx = fmaxf(sx, 1.f);
  const float cy = ((float)(oy + top) + 0.5f) * sy, cx = ((float)(ox + left) + 0.5f) * sx;   // centre in source pixel-edge units
  const int y0 = max((int)(cy - supy + 0.5f), 0), y1 = min((int)(cy + supy + 0.5f), h);
  const int x0 = max((int)(cx - supx + 0.5f), 0), x1 = min((int)(cx + supx + 0.5f), w);
  float acc[3] = {0.f, 0.f, 0.f}, wsum = 0.f;
  for (int yy = y0; yy < y1; ++yy) {
    const float wy = fmaxf(0.f, 1.f - fabsf(((float)yy + 0.5f - cy) / supy));
    if (wy == 0.f) continue;
    const uint8_t* row = src + ((long long)yy * w) * 3;
    float racc[3] = {0.f, 0.f, 0.f}, rw_ = 0.f;
    for (int xx = x0; xx < x1; ++xx) {
      const float wx = fmaxf(0.f, 1.f - fabsf(((float)xx + 0.5f - cx) / supx));
      racc[0] = fmaf(wx, (float)row[xx * 3 + 0], racc[0]);
      racc[1] = fmaf(wx, (float)row[xx * 3 + 1], racc[1]);
      racc[2] = fmaf(wx, (float)row[xx * 3 + 2], racc[2]);
      rw_ += wx;
    }
    acc[0] = fmaf(wy, racc[0], acc[0]);
    acc[1] = fmaf(wy, racc[1], acc[1]);
    acc[2] = fmaf(wy, racc[2], acc[2]);
    wsum = fmaf(wy, rw_, wsum);
  }
  const float inv = wsum > 0.f ? 1.f / wsum : 0.f;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    float v = acc[ch] * inv;
    if (quantize) v = floorf(fminf(fmaxf(v, 0.f), 255.f));
    dst[(long long)ch * th * tw + idx] = (v * (1.f / 255.f) - 0.5f) * 2.f;
  }
}

extern "C" int cgb_resize_crop_u8(const uint8_t* src, float* dst, int32_t h, int32_t w, int32_t rh, int32_t rw, int32_t top,
                                  int32_t left, int32_t th, int32_t tw, int32_t quantize, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(src && dst && h > 0 && w > 0 && rh > 0 && rw > 0 && th > 0 && tw > 0, "resize_crop_u8: bad arguments");
  CGB_REQUIRE(top >= 0 && left >= 0 && top + th <= rh && left + tw <= rw, "resize_crop_u8: the crop [%d:%d, %d:%d] leaves the resized image %dx%d",
              top, top + th, left, left + tw, rh, rw);
  resize_crop_u8_kernel<<<(th * tw + 255) / 256, 256, 0, (cudaStream_t)stream>>>(src, dst, h, w, rh, rw, top, left, th, tw, quantize);
  return after_launch("resize_crop_u8");
}

extern "C" int cgb_to_uint8_nhwc(const float* x, const float* mm, uint8_t* out, int32_t n, int32_t hw, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && mm && out && n > 0 && hw > 0, "to_uint8_nhwc: bad arguments");
  int gx = grid_for(hw);
  if (gx > 148 * 2) gx = 148 * 2;
  to_uint8_nhwc_kernel<<<dim3(gx, n), 256, 0, (cudaStream_t)stream>>>(x, mm, out, hw);
  return after_launch("to_uint8_nhwc");
}

extern "C" int cgb_mask_to_uint8(const float* m, uint8_t* out, float bin_value, int64_t count, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(m && out && count > 0, "mask_to_uint8: bad arguments");
  mask_to_uint8_kernel<<<grid_for(count), 256, 0, (cudaStream_t)stream>>>(m, out, bin_value, count);
  return after_launch("mask_to_uint8");
}

extern "C" int cgb_perlin_noise(const float* angles, float* out, int32_t h, int32_t w, int32_t res0, int32_t res1, void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(angles && out && h > 0 && w > 0 && res0 > 0 && res1 > 0 && h >= res0 && w >= res1, "perlin_noise: bad arguments");
  perlin_kernel<<<grid_for((long long)h * w), 256, 0, (cudaStream_t)stream>>>(angles, out, h, w, res0, res1);
  return after_launch("perlin_noise");
}

extern "C" int cgb_cloudy_mix(const float* x, const float* seg, const float* noise, const float* mm_noise, float* out, int32_t n,
                              int32_t h, int32_t w, int32_t c, int32_t hs, int32_t ws, int32_t sky_idx, float weight,
                              void* stream) {
  CGB_CHECK_DEVICE();
  CGB_REQUIRE(x && seg && noise && mm_noise && out && n > 0 && c > 0, "cloudy_mix: bad arguments");
  int gx = grid_for((long long)h * w);
  if (gx > 148 * 2) gx = 148 * 2;
  cloudy_mix_kernel<<<dim3(gx, n), 256, 0, (cudaStream_t)stream>>>(x, seg, noise, mm_noise, out, h, w, c, hs, ws, sky_idx, weight);
  return after_launch("cloudy_mix");
}
