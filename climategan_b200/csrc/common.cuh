// Shared helpers for libcgb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/cgb200.h"

namespace cgb {

// ---- error / launch bookkeeping (defined in api.cu) ----------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
int check_device();  // CGB_OK or CGB_UNSUPPORTED_ARCH (cached per process)

#define CGB_REQUIRE(cond, ...)           \
  do {                                   \
    if (!(cond)) {                       \
      cgb::set_error(__VA_ARGS__);       \
      return CGB_BAD_ARG;                \
    }                                    \
  } while (0)

#define CGB_CHECK_DEVICE()                  \
  do {                                      \
    int _s = cgb::check_device();           \
    if (_s != CGB_OK) return _s;            \
  } while (0)

// call after every kernel launch
inline int after_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return CGB_LAUNCH_FAILURE;
  }
  return CGB_OK;
}

// ---- element access ---------------------------------------------------------------------------
template <typename T>
struct Vec8;  // 8 consecutive channels

template <>
struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};

template <>
struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 r = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = r;
  }
};

template <>
struct Vec8<__half> {   // fp16 storage: the inference-only --half mode (apply_events.py:467-468, trainer.py:263-264)
  static __device__ __forceinline__ void load(const __half* p, float (&v)[8]) {
    uint4 r = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __half22float2(h[i]);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__half* p, const float (&v)[8]) {
    uint4 r;
    __half2* h = reinterpret_cast<__half2*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = r;
  }
};

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <>
__device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }

template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// ---- activations ----------------------------------------------------------------------------------
__device__ __forceinline__ float act_apply(float v, int act, float slope) {
  switch (act) {
    case CGB_ACT_RELU: return v > 0.f ? v : 0.f;
    case CGB_ACT_LRELU: return v > 0.f ? v : v * slope;
    case CGB_ACT_TANH: return tanhf(v);
    case CGB_ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    case CGB_ACT_SELU: return 1.0507009873554804934193349852946f * (v > 0.f ? v : 1.6732632423543772848170429916717f * (__expf(v) - 1.f));
    default: return v;
  }
}
// derivative expressed through the activation OUTPUT y
__device__ __forceinline__ float act_grad_from_out(float y, int act, float slope) {
  switch (act) {
    case CGB_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case CGB_ACT_LRELU: return y > 0.f ? 1.f : slope;
    case CGB_ACT_TANH: return 1.f - y * y;
    case CGB_ACT_SIGMOID: return y * (1.f - y);
    // selu: y = s*x (x > 0) | s*a*(e^x - 1) (x <= 0)  ->  dy/dx = s | y + s*a
    case CGB_ACT_SELU: return y > 0.f ? 1.0507009873554804934193349852946f : y + 1.0507009873554804934193349852946f * 1.6732632423543772848170429916717f;
    default: return 1.f;
  }
}

// reflect index into [0,n) (nn.ReflectionPad2d semantics, pad < n)
__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace cgb
