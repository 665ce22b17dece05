// TEST-ONLY CPU shim for a small subset of CUDA C++ (see vx_common.cuh, VX_EMU).
//
// Purpose: let the kernel sources in veloxseg_b200/csrc compile with g++ so their indexing logic can be
// checked against the oracle inside the GPU-less build container.  One CTA runs at a time; every CUDA
// thread of the CTA is an OS thread; __syncthreads and warp shuffles are barriers.  This is developer
// tooling: it is not built by __graft_entry__.build(), not loaded by veloxseg_b200, and is not a CPU
// fallback of the product (the product fails loudly without its sm_100a library).
#pragma once
#include <barrier>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <tuple>
#include <utility>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __constant__ static

struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uint3e { unsigned x, y, z; };
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
struct alignas(16) longlong2 { long long x, y; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline int4 make_int4(int a, int b, int c, int d) { return int4{a, b, c, d}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
static inline cudaError_t cudaGetLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return 0; }
enum { cudaMemcpyDeviceToDevice = 3 };

namespace vx_emu {
extern thread_local uint3e t_threadIdx, t_blockIdx;
extern thread_local int t_lane_slot;           // thread index inside the CTA
extern dim3 g_blockDim, g_gridDim;
unsigned char* dyn_smem();
void cta_barrier();
uint32_t shfl(uint32_t bits, int src_lane);    // src_lane relative to the caller's warp
void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);

template <typename... KArgs, typename... Args>
void launch(void (*k)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
  auto tup = std::make_tuple(static_cast<KArgs>(args)...);
  run_grid(grid, block, smem, [&]() { std::apply(k, tup); });
}
}  // namespace vx_emu

#define threadIdx (vx_emu::t_threadIdx)
#define blockIdx (vx_emu::t_blockIdx)
#define blockDim (vx_emu::g_blockDim)
#define gridDim (vx_emu::g_gridDim)

static inline void __syncthreads() { vx_emu::cta_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline float __shfl_xor_sync(unsigned, float v, int m) {
  uint32_t b; memcpy(&b, &v, 4); b = vx_emu::shfl(b, (vx_emu::t_lane_slot & 31) ^ m); float r; memcpy(&r, &b, 4); return r;
}
static inline int __shfl_xor_sync(unsigned, int v, int m) {
  return (int)vx_emu::shfl((uint32_t)v, (vx_emu::t_lane_slot & 31) ^ m);
}
static inline float __shfl_sync(unsigned, float v, int src) {
  uint32_t b; memcpy(&b, &v, 4); b = vx_emu::shfl(b, src); float r; memcpy(&r, &b, 4); return r;
}
static inline float __shfl_down_sync(unsigned, float v, int d) {
  int src = (vx_emu::t_lane_slot & 31) + d; if (src > 31) src = vx_emu::t_lane_slot & 31;
  uint32_t b; memcpy(&b, &v, 4); b = vx_emu::shfl(b, src); float r; memcpy(&r, &b, 4); return r;
}
static inline float atomicAdd(float* p, float v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(p);
  uint32_t old = __atomic_load_n(u, __ATOMIC_RELAXED);
  for (;;) {
    float f; memcpy(&f, &old, 4); f += v; uint32_t nv; memcpy(&nv, &f, 4);
    if (__atomic_compare_exchange_n(u, &old, nv, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) { float r; memcpy(&r, &old, 4); return r; }
  }
}
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
template <typename T> static inline T __ldg(const T* p) { return *p; }
#define __expf(x) expf(x)
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
static inline int __float_as_int(float v) { int i; memcpy(&i, &v, 4); return i; }
#define INFINITY_F (__builtin_inff())
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
