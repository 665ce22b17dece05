// Shared device helpers for libveloxseg_sm100.  sm_100a only.
//
// VX_EMU is a *test-only* build mode (tools/emu): the same kernel sources are compiled by g++ against a
// thread-per-lane CPU shim so indexing logic can be checked in a container that has no GPU.  The product
// library is never built with VX_EMU and has no CPU path.
#pragma once
#ifdef VX_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include <stddef.h>
#include <math.h>

#include "../../include/veloxseg_abi.h"

#ifdef VX_EMU
#define VX_LAUNCH(kern, grid, block, smem, stream, ...)                                   \
  do { auto _k = kern; vx_emu::launch(_k, dim3(grid), dim3(block), (size_t)(smem), __VA_ARGS__); } while (0)
#define VX_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(vx_emu::dyn_smem())
#define VX_SET_SMEM(kern, bytes) do { } while (0)
#else
// Programmatic dependent launch (VX_OPT_PDL, default OFF): with the option on, every kernel is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization; a kernel starts with griddepcontrol.wait (returns once the PREVIOUS kernel
// of the stream has completed and flushed) + griddepcontrol.launch_dependents (the NEXT kernel is dispatched and becomes resident
// while this one runs) -- VX_PDL_ENTRY, the first statement of every __global__ function of this library; with the option off
// both instructions are no-ops.  The idea was to pay the dispatch latency of the ~400 serial sub-wave launches of a step once
// instead of per edge.  Measured on a B200 (profiles/r3a_pdl_ab.txt, same tree, 100 steps each): 6.94 ms/step with
// launch_dependents before wait, 6.86 ms with it after, against 6.63-6.66 ms without PDL (sliding window 26.5-28.0 vs 24.3 ms):
// the early-resident CTAs of the next kernel take SM slots from the kernels the forked streams run concurrently.  Kept as an
// A/B switch, results bit-identical either way.
#define VX_LAUNCH(kern, grid, block, smem, stream, ...)                                   \
  do { auto _k = kern; const int _pi = vx::prof_begin(#kern, (stream));                  \
       vx::launch_pdl(_k, dim3(grid), dim3(block), (size_t)(smem), (stream), __VA_ARGS__); vx::count_launch();              \
       vx::prof_end(_pi, (stream)); } while (0)
#define VX_DYN_SMEM(type, name)                                                           \
  extern __shared__ __align__(16) unsigned char vx_dsm_[];                                \
  type* name = reinterpret_cast<type*>(vx_dsm_)
#define VX_SET_SMEM(kern, bytes)                                                          \
  do { auto _k = kern; static size_t _have = 32 * 1024;   /* one static per call site = per kernel instantiation; the default \
       48 KB limit counts static shared memory too, so the opt-in starts well below it */ \
       if ((size_t)(bytes) > _have) {                                                     \
         cudaFuncSetAttribute(_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)); _have = (bytes); } } while (0)
#endif

#define VX_DEV __device__ __forceinline__

#ifdef VX_EMU
#define VX_PDL_ENTRY() do { } while (0)
#else
#define VX_PDL_ENTRY() do { asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); } while (0)
#endif

// 4-byte asynchronous global -> shared copy (LDGSTS); `ok == false` zero-fills.
#ifdef VX_EMU
static inline void vx_cp_async4(float* dst, const float* src, bool ok) { *dst = ok ? *src : 0.f; }
static inline void vx_cp_async8(float* dst, const float* src, bool ok) { dst[0] = ok ? src[0] : 0.f; dst[1] = ok ? src[1] : 0.f; }
static inline void vx_cp_async_commit() {}
static inline void vx_cp_async_wait_all() {}
#else
// 8-byte variant: both addresses 8-byte aligned
VX_DEV void vx_cp_async8(float* dst, const float* src, bool ok) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int n = ok ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(src), "r"(n) : "memory");
}
VX_DEV void vx_cp_async4(float* dst, const float* src, bool ok) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int n = ok ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(src), "r"(n) : "memory");
}
VX_DEV void vx_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
VX_DEV void vx_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
#endif

namespace vx {

constexpr int kSMs = 148;

#ifndef VX_EMU
int pdl_enabled();                                     // VX_OPT_PDL
void pdl_set(int on);
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*k)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled();
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k, static_cast<KArgs>(args)...);
}
#endif

void set_error(const char* fmt, ...);
void count_launch();
#ifndef VX_EMU
int prof_begin(const char* kernel, cudaStream_t st);   // -1 when profiling is off
void prof_end(int slot, cudaStream_t st);
#endif
void prof_bytes(double bytes);                         // algorithmic bytes of the NEXT launch (profiler only)
void prof_flops(double flops);                         // algorithmic flops of the NEXT launch (profiler only)
void prof_scope(const char* fmt, ...);                 // names the op whose kernels follow (thread-local)
int check_launch(const char* what);   // cudaGetLastError -> vx_status

// round-to-nearest-even to bfloat16, kept in an fp32 container (bf16 numerics mode; finite inputs)
VX_DEV float bf16_round(float x) {
#ifdef VX_EMU
  uint32_t u; memcpy(&u, &x, 4); u += 0x7FFFu + ((u >> 16) & 1u); u &= 0xFFFF0000u; float r; memcpy(&r, &u, 4); return r;
#else
  uint32_t u = __float_as_uint(x);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return __uint_as_float(u & 0xFFFF0000u);
#endif
}
// 0: fp32-accurate (3xTF32 on the tensor cores), 1: bf16 numerics (tensor-core operands and outputs rounded to bf16, one product)
int precision_mode();

// erf(u) and exp(-u^2) together: Abramowitz & Stegun 7.1.26, erf(u) = sign(u) (1 - (a1 t + ... + a5 t^5) exp(-u^2)),
// t = 1 / (1 + p |u|), absolute error <= 1.5e-7 (fp32 resolution near 1) -- 14 instructions instead of the 27 (erff) / 39 (erff +
// expf) per value of the library calls, which dominated the issue slots of the element-wise and epilogue code (profiles/
// r2m_pw_tc_L2.source.txt: 30 % of the samples).  The exponential is shared with the Gaussian density of GELU'.
VX_DEV float erf_exp(float u, float& e) {
  const float a = fabsf(u);
#ifdef VX_EMU
  const float t = 1.0f / fmaf(0.3275911f, a, 1.0f);
  e = expf(-a * a);
#else
  const float t = __fdividef(1.0f, fmaf(0.3275911f, a, 1.0f));
  e = __expf(-a * a);
#endif
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float r = fmaf(-p * t, e, 1.0f);
  return copysignf(r, u);
}
VX_DEV float gelu_f(float x) {
  float e;
  return 0.5f * x * (1.0f + erf_exp(x * 0.70710678118654752440f, e));
}
// d/dx GELU(x) = Phi(x) + x * phi(x)
VX_DEV float gelu_grad_f(float x) {
  float e;
  const float cdf = 0.5f * (1.0f + erf_exp(x * 0.70710678118654752440f, e));
  return fmaf(x * 0.39894228040143267794f, e, cdf);
}

VX_DEV float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
VX_DEV float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum of up to 32 warps; every thread gets the result.  `red` is >= 33 floats of shared memory.
VX_DEV float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) { t = warp_sum(t); if (lane == 0) red[32] = t; }
  __syncthreads();
  return red[32];
}

// Counter-based RNG for dropout masks: 32 random bits per (seed, site, element index) from an integer hash (two rounds of
// the "lowbias32" xorshift-multiply finaliser over the element index, keyed by a hash of seed and site).  Forward and backward
// regenerate the same keep / drop decision without storing a mask.  Round 1 used Philox-4x32-10 (145 instructions per 4
// elements, one full call per ELEMENT in the per-element paths): the mask generation was the largest single block of issue
// slots in the contraction kernels of levels 3-4 (profiles/r2n_step_stalls.txt: 2 500-3 600 instructions per warp for <= 20
// MFLOP).  A dropout mask needs independence across elements and sites, not cryptographic strength: ~10 instructions per element.
VX_DEV uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
VX_DEV uint32_t rng_key(uint64_t seed, uint32_t site) {
  return mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) + site * 0x9E3779B9u + 0x85EBCA6Bu));
}
VX_DEV uint32_t rng_word(uint32_t key, uint64_t idx) {
  const uint32_t lo = (uint32_t)idx, hi = (uint32_t)(idx >> 32);
  return mix32((lo * 0x9E3779B1u) ^ key ^ (hi * 0xC2B2AE35u));
}
#ifdef VX_EMU
#define VX_NOINLINE
#else
#define VX_NOINLINE __noinline__
#endif
// Out-of-line copies for kernels whose code size matters (a kernel runs once per thread: unrolled call sites of erff / expf
// style helpers multiply the SASS beyond the 32 KB instruction cache and the kernel becomes instruction-fetch bound --
// measured: 68 % of pw_kernel's stall samples were `no_instructions` at 200 KB of SASS).
// the four words of elements 4 ctr .. 4 ctr + 3 (same values as rng_word element by element)
VX_DEV uint4 rng4(uint64_t seed, uint64_t ctr, uint32_t site) {
  const uint32_t key = rng_key(seed, site);
  uint4 o;
  o.x = rng_word(key, 4 * ctr); o.y = rng_word(key, 4 * ctr + 1); o.z = rng_word(key, 4 * ctr + 2); o.w = rng_word(key, 4 * ctr + 3);
  return o;
}
static __device__ VX_NOINLINE float4 gelu_grad4_call(float4 x) {
  return make_float4(gelu_grad_f(x.x), gelu_grad_f(x.y), gelu_grad_f(x.z), gelu_grad_f(x.w));
}
static __device__ VX_NOINLINE float4 gelu4_call(float4 x) {
  return make_float4(gelu_f(x.x), gelu_f(x.y), gelu_f(x.z), gelu_f(x.w));
}
VX_DEV float keep_from_bits(uint32_t bits, float p, float inv_keep) {
  return ((float)(bits >> 8) * (1.0f / 16777216.0f) < p) ? 0.f : inv_keep;
}
// keep-scale for element `idx` of dropout site `site`: 0 (dropped) or 1/(1-p).
VX_DEV float dropout_scale(uint64_t seed, uint32_t site, uint64_t idx, float p, float inv_keep) {
  return keep_from_bits(rng_word(rng_key(seed, site), idx), p, inv_keep);
}
// keep-scales of the 4 consecutive elements idx .. idx+3 (same values as dropout_scale element by element)
VX_DEV void dropout_scale4(uint64_t seed, uint32_t site, uint64_t idx, float p, float inv_keep, float (&ms)[4]) {
  const uint32_t key = rng_key(seed, site);
#pragma unroll
  for (int i = 0; i < 4; ++i) ms[i] = keep_from_bits(rng_word(key, idx + i), p, inv_keep);
}

// mean / rstd from partial sums laid out [row][npart][2] (sum, sumsq); n = element count behind the full row.
VX_DEV void finalize_stats(const float* __restrict__ part, int row, int npart, float n, float eps,
                           float& mean, float& rstd) {
  double s = 0.0, q = 0.0;
  const float* p = part + (size_t)row * npart * 2;
  for (int i = 0; i < npart; ++i) { s += (double)p[2 * i]; q += (double)p[2 * i + 1]; }
  const double m = s / (double)n;
  double var = q / (double)n - m * m;
  if (var < 0.0) var = 0.0;
  mean = (float)m;
  rstd = (float)(1.0 / sqrt(var + (double)eps));
}

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace vx
