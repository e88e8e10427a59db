// libcgb200 C ABI: library bookkeeping + convolution dispatch (SIMT vs tcgen05 engines).
#include "common.cuh"
#include <stdarg.h>
#include <string.h>
#include <vector>
#include <mutex>
#include <map>
#include <string>

namespace cgb {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_device() {
  static std::atomic<int> cached{-100};
  int c = cached.load();
  if (c != -100) {
    if (c != CGB_OK) set_error("libcgb200 needs an sm_100 (B200) device; no CPU fallback exists");
    return c;
  }
  int dev = 0, major = 0, minor = 0;
  int st = CGB_OK;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) {
    cudaGetLastError();
    st = CGB_UNSUPPORTED_ARCH;
  } else if (major != 10) {
    st = CGB_UNSUPPORTED_ARCH;
  }
  cached.store(st);
  if (st != CGB_OK) set_error("libcgb200 needs an sm_100 (B200) device; no CPU fallback exists");
  return st;
}

// ---- optional per-launch timing of the conv engines (bench.py roofline leg) ----------------------
// When enabled, every conv launch is bracketed by two CUDA events recorded on the launching stream;
// cgb_prof_dump() (after the caller synchronised) folds them into per-(op,engine,shape) totals.
struct ProfRec { cudaEvent_t a, b; cgb_conv_desc d; int which; int tc; };
static std::vector<ProfRec> g_prof;
static std::mutex g_prof_mu;
static std::atomic<int> g_prof_on{0};

static std::vector<cudaEvent_t> g_event_pool;
static cudaEvent_t take_event() {
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_event_pool.empty()) {
      cudaEvent_t e = g_event_pool.back();
      g_event_pool.pop_back();
      return e;
    }
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

struct ProfScope {
  bool on; ProfRec r; cudaStream_t st;
  ProfScope(const cgb_conv_desc* d, int which, bool tc, cudaStream_t s) : on(g_prof_on.load() != 0), st(s) {
    if (!on) return;
    r.d = *d; r.which = which; r.tc = tc ? 1 : 0;
    r.a = take_event(); r.b = take_event();   // pooled: creating two events per launch costs more than recording them
    cudaEventRecord(r.a, st);
  }
  ~ProfScope() {
    if (!on) return;
    cudaEventRecord(r.b, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(r);
  }
};

// engines
int conv_simt_fwd(const cgb_conv_desc*, const void*, const void*, const float*, const void*, void*, cudaStream_t);
int conv_simt_dgrad(const cgb_conv_desc*, const void*, const void*, int, const void*, void*, cudaStream_t);
int conv_simt_wgrad(const cgb_conv_desc*, const void*, const void*, float*, float*, cudaStream_t);
bool conv_tc_supported(const cgb_conv_desc*, int which);
int conv_tc_fwd(const cgb_conv_desc*, const void*, const void*, const float*, const void*, void*, cudaStream_t, float* stats_out = nullptr);
int conv_tc_stats_rows();
int conv_tc_dgrad(const cgb_conv_desc*, const void*, const void*, int, const void*, void*, cudaStream_t);
int conv_tc_wgrad(const cgb_conv_desc*, const void*, const void*, float*, float*, cudaStream_t);
int pack_dgrad_weight(const cgb_conv_desc*, const void*, void*, cudaStream_t);

static int validate(const cgb_conv_desc* d, const char* who) {
  if (!d) { set_error("%s: null descriptor", who); return CGB_BAD_ARG; }
  if (d->n <= 0 || d->hi <= 0 || d->wi <= 0 || d->ho <= 0 || d->wo <= 0) {
    set_error("%s: empty tensor (n=%d hi=%d wi=%d ho=%d wo=%d)", who, d->n, d->hi, d->wi, d->ho, d->wo);
    return CGB_BAD_ARG;
  }
  if (d->ci < 8 || d->co < 8 || d->ci % 8 || d->co % 8) {
    set_error("%s: storage channels must be multiples of 8 (ci=%d co=%d)", who, d->ci, d->co);
    return CGB_BAD_ARG;
  }
  if (d->kh < 1 || d->kw < 1 || d->stride < 1 || d->dil < 1 || d->pad < 0) {
    set_error("%s: bad geometry k=%dx%d stride=%d dil=%d pad=%d", who, d->kh, d->kw, d->stride, d->dil, d->pad);
    return CGB_BAD_ARG;
  }
  const int eh = (d->hi + 2 * d->pad - d->dil * (d->kh - 1) - 1) / d->stride + 1;
  const int ew = (d->wi + 2 * d->pad - d->dil * (d->kw - 1) - 1) / d->stride + 1;
  if (eh != d->ho || ew != d->wo) {
    set_error("%s: output size %dx%d does not match geometry (expected %dx%d)", who, d->ho, d->wo, eh, ew);
    return CGB_BAD_ARG;
  }
  if (d->pad_mode == CGB_PAD_REFLECT && (d->pad >= d->hi || d->pad >= d->wi)) {
    set_error("%s: reflect pad %d must be smaller than the input (%dx%d)", who, d->pad, d->hi, d->wi);
    return CGB_BAD_ARG;
  }
  if (d->pad_mode != CGB_PAD_ZERO && d->pad_mode != CGB_PAD_REFLECT) {
    set_error("%s: unknown pad_mode %d", who, d->pad_mode);
    return CGB_BAD_ARG;
  }
  if (d->dtype != CGB_F32 && d->dtype != CGB_BF16 && d->dtype != CGB_F16) {
    set_error("%s: unknown dtype %d", who, d->dtype);
    return CGB_BAD_ARG;
  }
  return CGB_OK;
}

static int pick_engine(const cgb_conv_desc* d, int which, const char* who, bool* use_tc) {
  const bool ok = conv_tc_supported(d, which);
  if (d->engine == CGB_ENGINE_TCGEN05 && !ok) {
    set_error("%s: shape does not qualify for the tcgen05 engine", who);
    return CGB_UNSUPPORTED;
  }
  *use_tc = ok && d->engine != CGB_ENGINE_SIMT;
  return CGB_OK;
}

}  // namespace cgb

using namespace cgb;

extern "C" const char* cgb_version(void) { return "cgb200 0.1 (sm_100a)"; }
extern "C" const char* cgb_last_error(void) { return g_err; }
extern "C" int cgb_device_ok(void) { return check_device() == CGB_OK ? 1 : 0; }
extern "C" int64_t cgb_launch_count(void) { return g_launches.load(); }
extern "C" void cgb_launch_count_reset(void) { g_launches.store(0); }

extern "C" int cgb_conv2d_uses_tcgen05(const cgb_conv_desc* d, int which) {
  if (!d || which < 0 || which > 2) return 0;
  if (d->engine == CGB_ENGINE_SIMT) return 0;
  return conv_tc_supported(d, which) ? 1 : 0;
}

extern "C" int cgb_conv2d_fwd(const cgb_conv_desc* d, const void* x, const void* w, const float* bias,
                              const void* residual, void* y, void* stream) {
  CGB_CHECK_DEVICE();
  int s = validate(d, "conv2d_fwd");
  if (s) return s;
  CGB_REQUIRE(x && w && y, "conv2d_fwd: null pointer");
  CGB_REQUIRE(!(residual && d->act != CGB_ACT_NONE && !d->res_before_act),
              "conv2d_fwd: residual after an activation requires act=none (set res_before_act for act(conv+residual))");
  bool tc = false;
  s = pick_engine(d, 0, "conv2d_fwd", &tc);
  if (s) return s;
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope ps(d, 0, tc, st);
  return tc ? conv_tc_fwd(d, x, w, bias, residual, y, st) : conv_simt_fwd(d, x, w, bias, residual, y, st);
}

// conv + per-channel statistics of its output in one launch (tcgen05 engine only): see include/cgb200.h
extern "C" int32_t cgb_conv2d_stats_rows(void) { return conv_tc_stats_rows(); }

extern "C" int cgb_conv2d_fwd_stats(const cgb_conv_desc* d, const void* x, const void* w, const float* bias,
                                    const void* residual, void* y, float* stats_partial, void* stream) {
  CGB_CHECK_DEVICE();
  int s = validate(d, "conv2d_fwd_stats");
  if (s) return s;
  CGB_REQUIRE(x && w && y && stats_partial, "conv2d_fwd_stats: null pointer");
  CGB_REQUIRE(!(residual && d->act != CGB_ACT_NONE && !d->res_before_act),
              "conv2d_fwd_stats: residual after an activation requires act=none");
  if (d->engine == CGB_ENGINE_SIMT || !conv_tc_supported(d, 0)) {
    set_error("conv2d_fwd_stats: the epilogue statistics exist in the tcgen05 engine only (check cgb_conv2d_uses_tcgen05)");
    return CGB_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope ps(d, 0, true, st);
  return conv_tc_fwd(d, x, w, bias, residual, y, st, stats_partial);
}

extern "C" int cgb_conv2d_pack_dgrad_weight(const cgb_conv_desc* d, const void* w, void* wt, void* stream) {
  CGB_CHECK_DEVICE();
  int s = validate(d, "conv2d_pack_dgrad_weight");
  if (s) return s;
  CGB_REQUIRE(w && wt, "conv2d_pack_dgrad_weight: null pointer");
  return pack_dgrad_weight(d, w, wt, (cudaStream_t)stream);
}

extern "C" int cgb_conv2d_dgrad(const cgb_conv_desc* d, const void* gy, const void* w, const void* wt, int32_t dact,
                                const void* mask_src, void* gx, void* stream) {
  CGB_CHECK_DEVICE();
  int s = validate(d, "conv2d_dgrad");
  if (s) return s;
  CGB_REQUIRE(gy && (w || wt) && gx, "conv2d_dgrad: null pointer");
  if (d->pad_mode == CGB_PAD_REFLECT && d->pad > 0) {
    set_error("conv2d_dgrad: reflect padding is not implemented for the data gradient");
    return CGB_UNSUPPORTED;
  }
  if (!mask_src) dact = CGB_ACT_NONE;
  bool tc = false;
  s = pick_engine(d, 1, "conv2d_dgrad", &tc);
  if (s) return s;
  if (tc && !wt) {
    if (d->engine == CGB_ENGINE_TCGEN05) {
      set_error("conv2d_dgrad: the tcgen05 engine needs the dgrad weight packing (wt)");
      return CGB_BAD_ARG;
    }
    tc = false;
  }
  CGB_REQUIRE(tc || w, "conv2d_dgrad: the SIMT engine needs the forward weight packing (w)");
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope ps(d, 1, tc, st);
  return tc ? conv_tc_dgrad(d, gy, wt, dact, mask_src, gx, st) : conv_simt_dgrad(d, gy, w, dact, mask_src, gx, st);
}

extern "C" int cgb_conv2d_wgrad(const cgb_conv_desc* d, const void* x, const void* gy, float* gw,
                                float* gbias, int32_t accumulate, void* stream) {
  CGB_CHECK_DEVICE();
  int s = validate(d, "conv2d_wgrad");
  if (s) return s;
  CGB_REQUIRE(x && gy && gw, "conv2d_wgrad: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) {
    cudaMemsetAsync(gw, 0, sizeof(float) * (size_t)d->co * d->kh * d->kw * d->ci, st);
    if (gbias) cudaMemsetAsync(gbias, 0, sizeof(float) * (size_t)d->co, st);
  }
  bool tc = false;
  s = pick_engine(d, 2, "conv2d_wgrad", &tc);
  if (s) return s;
  ProfScope ps(d, 2, tc, st);
  return tc ? conv_tc_wgrad(d, x, gy, gw, gbias, st) : conv_simt_wgrad(d, x, gy, gw, gbias, st);
}

// ---- profiler ABI ---------------------------------------------------------------------------------
extern "C" void cgb_prof_enable(int on) {
  g_prof_on.store(on ? 1 : 0);
}

// Folds all recorded launches into text lines "which engine n hi wi ci ho wo co kh kw stride dil count total_ms"
// written to buf (NUL terminated, truncated to cap).  Caller must have synchronised the device.
// Returns the number of distinct keys, clears the records.
extern "C" int cgb_prof_dump(char* buf, int64_t cap) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  std::map<std::string, std::pair<int, double>> agg;
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) { cudaGetLastError(); ms = 0.f; }
    g_event_pool.push_back(r.a); g_event_pool.push_back(r.b);
    char key[256];
    snprintf(key, sizeof(key), "%d %d %d %d %d %d %d %d %d %d %d %d %d", r.which, r.tc, r.d.n, r.d.hi, r.d.wi,
             r.d.ci, r.d.ho, r.d.wo, r.d.co, r.d.kh, r.d.kw, r.d.stride, r.d.dil);
    auto& e = agg[key];
    e.first += 1; e.second += ms;
  }
  g_prof.clear();
  std::string out;
  for (auto& kv : agg) {
    char line[384];
    snprintf(line, sizeof(line), "%s %d %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    out += line;
  }
  if (buf && cap > 0) {
    size_t nb = out.size() < (size_t)(cap - 1) ? out.size() : (size_t)(cap - 1);
    memcpy(buf, out.data(), nb);
    buf[nb] = 0;
  }
  return (int)agg.size();
}
