// Error reporting and version entry points of libveloxseg_sm100.
#include "vx_kernels.h"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace vx {

static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

static int g_pdl = 0;      // measured slower when on (see vx_common.cuh): opt-in
int pdl_enabled() { return g_pdl; }
void pdl_set(int on) { g_pdl = on ? 1 : 0; }

static int g_precision = 0;
int precision_mode() { return g_precision; }
void precision_set(int m) { g_precision = m ? 1 : 0; }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return VX_ERR_LAUNCH;
  }
  return VX_OK;
}

static thread_local const void* g_seed_dev = nullptr;
void set_seed_dev(const void* p) { g_seed_dev = p; }
const unsigned long long* get_seed_dev() { return (const unsigned long long*)g_seed_dev; }

static thread_local char g_scope[96] = "";
void prof_scope(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_scope, sizeof(g_scope), fmt, ap);
  va_end(ap);
}

#ifndef VX_EMU
// ---- weight-gradient side streams: one per (device, caller stream), created on first use
struct SideCtx { cudaStream_t side; cudaEvent_t fork, join, mid; bool forked; };
static std::mutex g_side_mu;
static std::map<std::pair<int, cudaStream_t>, SideCtx> g_side;
static int g_side_on = 1;
void side_set(int enabled) { g_side_on = enabled; }

static SideCtx* side_ctx(cudaStream_t main) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_side_mu);
  auto key = std::make_pair(dev, main);
  auto it = g_side.find(key);
  if (it == g_side.end()) {
    SideCtx c{};
    if (cudaStreamCreateWithFlags(&c.side, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    cudaEventCreateWithFlags(&c.fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c.join, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c.mid, cudaEventDisableTiming);
    c.forked = false;
    it = g_side.emplace(key, c).first;
  }
  return &it->second;
}

cudaStream_t side_fork(cudaStream_t main) {
  if (!g_side_on) return main;
  SideCtx* c = side_ctx(main);
  if (!c) return main;
  if (cudaEventRecord(c->fork, main) != cudaSuccess || cudaStreamWaitEvent(c->side, c->fork, 0) != cudaSuccess) {
    cudaGetLastError();
    return main;
  }
  c->forked = true;
  return c->side;
}

// main waits for everything enqueued on the side stream so far (the side stream stays forked)
void side_wait(cudaStream_t main) {
  if (!g_side_on) return;
  SideCtx* c = side_ctx(main);
  if (!c || !c->forked) return;
  cudaEventRecord(c->mid, c->side);
  cudaStreamWaitEvent(main, c->mid, 0);
}

void side_join(cudaStream_t main) {
  if (!g_side_on) return;
  SideCtx* c = side_ctx(main);
  if (!c || !c->forked) return;
  cudaEventRecord(c->join, c->side);
  cudaStreamWaitEvent(main, c->join, 0);
  c->forked = false;
}

struct ProfRec { std::string key; cudaEvent_t a, b; double bytes, flops; };
static thread_local double g_next_bytes = 0.0, g_next_flops = 0.0;
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
static int g_prof_on = 0;

static std::vector<cudaEvent_t> g_pool;   // events are created once and recycled (creation costs tens of microseconds)

static cudaEvent_t pool_get() {
  if (g_pool.empty()) { cudaEvent_t e; cudaEventCreate(&e); return e; }
  cudaEvent_t e = g_pool.back();
  g_pool.pop_back();
  return e;
}

int prof_begin(const char* kernel, cudaStream_t st) {
  if (!g_prof_on) return -1;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r;
  r.key = std::string(g_scope) + "|" + kernel;
  r.a = pool_get();
  r.b = pool_get();
  r.bytes = g_next_bytes;
  r.flops = g_next_flops;
  g_next_bytes = 0.0;
  g_next_flops = 0.0;
  cudaEventRecord(r.a, st);
  g_prof.push_back(r);
  return (int)g_prof.size() - 1;
}

void prof_bytes(double b) { g_next_bytes = b; }
void prof_flops(double f) { g_next_flops = f; }

void prof_end(int slot, cudaStream_t st) {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_prof[slot].b, st);
}
#endif

}  // namespace vx

#ifndef VX_EMU
extern "C" int vx_profile_enable(int on) {
  const int prev = vx::g_prof_on;
  if (on && !prev) {     // pre-create a pool so that the profiled pass does not pay for event creation
    std::lock_guard<std::mutex> lk(vx::g_prof_mu);
    while (vx::g_pool.size() < 8192) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) break; vx::g_pool.push_back(e); }
  }
  vx::g_prof_on = on;
  return prev;
}
extern "C" void vx_profile_reset(void) {
  std::lock_guard<std::mutex> lk(vx::g_prof_mu);
  for (auto& r : vx::g_prof) { vx::g_pool.push_back(r.a); vx::g_pool.push_back(r.b); }
  vx::g_prof.clear();
}
extern "C" size_t vx_profile_report(char* buf, size_t cap) {
  std::lock_guard<std::mutex> lk(vx::g_prof_mu);
  struct Agg { int n = 0; double ms = 0.0, bytes = 0.0, flops = 0.0; };
  std::map<std::string, Agg> agg;
  for (auto& r : vx::g_prof) {
    cudaEventSynchronize(r.b);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) ms = 0.f;
    auto& e = agg[r.key];
    e.n += 1; e.ms += ms; e.bytes += r.bytes; e.flops += r.flops;
  }
  std::string out;
  char line[128];
  for (auto& kv : agg) {
    snprintf(line, sizeof(line), "|%d|%.6f|%.0f|%.0f\n", kv.second.n, kv.second.ms, kv.second.bytes, kv.second.flops);
    out += kv.first + line;
  }
  if (buf && cap > 0) { const size_t n = out.size() < cap - 1 ? out.size() : cap - 1; memcpy(buf, out.data(), n); buf[n] = 0; }
  return out.size() + 1;
}
// One line per recorded launch, "scope|kernel|start_us|dur_us", start relative to the first record (device timeline of the
// profiled region across every stream the library launched on).
extern "C" size_t vx_profile_timeline(char* buf, size_t cap) {
  std::lock_guard<std::mutex> lk(vx::g_prof_mu);
  std::string out;
  char line[64];
  if (!vx::g_prof.empty()) {
    cudaEvent_t t0 = vx::g_prof[0].a;
    for (auto& r : vx::g_prof) {
      cudaEventSynchronize(r.b);
      float st = 0.f, du = 0.f;
      if (cudaEventElapsedTime(&st, t0, r.a) != cudaSuccess) { cudaGetLastError(); continue; }
      if (cudaEventElapsedTime(&du, r.a, r.b) != cudaSuccess) { cudaGetLastError(); continue; }
      snprintf(line, sizeof(line), "|%.3f|%.3f\n", st * 1e3, du * 1e3);
      out += r.key + line;
    }
  }
  if (buf && cap > 0) { const size_t n = out.size() < cap - 1 ? out.size() : cap - 1; memcpy(buf, out.data(), n); buf[n] = 0; }
  return out.size() + 1;
}
#else
extern "C" size_t vx_profile_timeline(char*, size_t) { return 0; }
namespace vx {
void prof_bytes(double) {}
void prof_flops(double) {}
cudaStream_t side_fork(cudaStream_t main) { return main; }
void side_wait(cudaStream_t) {}
void side_join(cudaStream_t) {}
void side_set(int) {}
}
extern "C" int vx_profile_enable(int) { return 0; }
extern "C" void vx_profile_reset(void) {}
extern "C" size_t vx_profile_report(char*, size_t) { return 0; }
#endif

extern "C" int vx_set_option(int option, int value) {
  if (option == VX_OPT_WGRAD_TC_MIN_S) { vx::pw_wgrad_tc_set(-1, value); return VX_OK; }
  if (option == VX_OPT_JLC_SMALL_MAX_S) { vx::jlc_set_small_max(value); return VX_OK; }
  if (option == VX_OPT_JLC_KS) { vx::jlc_set_ks(value); return VX_OK; }
  if (option == VX_OPT_JLC_SMALL_THREADS) { vx::jlc_set_small_threads(value); return VX_OK; }
  if (option == VX_OPT_FFN_TC) { vx::pw_ffn_tc_set(value ? 1 : 0); return VX_OK; }
  if (option == VX_OPT_ATTN_TC) { vx::pwa_attn_tc_set(value ? 1 : 0); return VX_OK; }
  if (option == VX_OPT_CONV3_TRACE) { vx::conv3_trace_set(value); return VX_OK; }
  if (option == VX_OPT_PRECISION) { vx::precision_set(value); return VX_OK; }
#ifndef VX_EMU
  if (option == VX_OPT_PDL) { vx::pdl_set(value); return VX_OK; }
#else
  if (option == VX_OPT_PDL) return VX_OK;
#endif
  if (option == VX_OPT_SIDE_WGRAD) { vx::side_set(value ? 1 : 0); return VX_OK; }
#ifndef VX_EMU
  if (option == VX_OPT_PW_TENSOR_CORES) { vx::pw_tc_set(value ? 1 : 0); vx::pw_wgrad_tc_set(value ? 1 : 0, -1); return VX_OK; }
  if (option == VX_OPT_PW_SMALL_MAX_S) { vx::pw_set_thresholds(value, -1); return VX_OK; }
  if (option == VX_OPT_PW_TC_MIN_S) { vx::pw_set_thresholds(-1, value); return VX_OK; }
  if (option == VX_OPT_JLC_TILE_FWD) { vx::jlc_force_tile(0, value >> 8, value & 255); return VX_OK; }
  if (option == 6) { vx::jlc_force_vx(value); return VX_OK; }     // voxels per thread of the JLC conv kernels (probe)
  if (option == VX_OPT_JLC_TILE_WGRAD) { vx::jlc_force_tile(1, value >> 8, value & 255); return VX_OK; }
#endif
  vx::set_error("vx_set_option: unknown option %d", option);
  return VX_ERR_BAD_DESC;
}

extern "C" int vx_version(void) { return 101; }
extern "C" uint64_t vx_launch_count(void) { return __atomic_load_n(&vx::g_launches, __ATOMIC_RELAXED); }
extern "C" const char* vx_last_error_string(void) { return vx::g_err; }

extern "C" int vx_conv3_trace(long long* out64, int n) { return vx::conv3_trace_read(out64, n); }
