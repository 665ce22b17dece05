// Portable wrappers around the sm_100a tensor-core / async-copy instructions used by the implicit-GEMM convolution kernels
// (conv3_tc.cu): tcgen05.mma kind::tf32 (cta_group::1, M = 128, SWIZZLE_NONE K-major operands), tensor memory alloc / ld,
// mbarriers, tcgen05.commit and cp.async.bulk (global -> shared, mbarrier-tracked).
//
// Under VX_EMU (tools/emu, developer tooling only) the same calls run a DESCRIPTOR-LEVEL software model: an MMA decodes
// its shared-memory descriptors (start, LBO, SBO, major bits) and reads the emulated shared memory exactly where the
// hardware would, so operand layouts, shifted start addresses and TMEM column bookkeeping are checked by the CPU tests with
// the kernel source unchanged.  Layout rule modelled (CUTLASS cute/atom/mma_traits_sm100.hpp "INTERLEAVE", confirmed on a
// B200 with tools/bringup/tc_probe2.cu incl. start addresses that are only 16-byte aligned):
//   K-major operand, element (r, k):  start + (r/8)*SBO + (r%8)*16 + (k/4)*LBO + (k%4)*4      (k < 8 per MMA)
//   D (M = 128):  row m -> TMEM lane m, column = d_col + n.
// MN-major operands (idesc bits 15 / 16) are NOT usable with kind::tf32: on the hardware every such MMA returned zeros
// whatever the descriptor fields (tools/bringup/tc_probe2.cu / tc_probe3.cu, profiles/r2a_tc_probe2.txt); the model aborts on them.
#pragma once
#include "vx_common.cuh"

#ifdef VX_EMU
#include <map>
#include <mutex>
#include <thread>
#endif

namespace vx {
namespace tc {

// shared-memory matrix descriptor (SWIZZLE_NONE): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46
VX_DEV uint64_t desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor, kind::tf32: D fp32 (bit 4), A/B tf32 (bits 7, 10), A / B MN-major (bits 15 / 16), N >> 3 at bit 17,
// M >> 4 at bit 24
VX_DEV uint32_t idesc_tf32(int n, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}

#ifndef VX_EMU
// ------------------------------------------------------------------------------------------------- hardware
VX_DEV uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
VX_DEV void split(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  lo = x - hi;
}
VX_DEV void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
VX_DEV void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
VX_DEV void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
VX_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_addr(bar);
  uint32_t ok, spins = 0;
  do {
    if (++spins > (1u << 26)) __trap();      // a lost arrival must surface as a launch error, never as a hung GPU
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!ok);
}
VX_DEV void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
// one arrival + `bytes` of pending transactions (the bulk copies that follow complete them)
VX_DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
// cp.async.bulk: `bytes` (multiple of 16, both addresses 16-byte aligned) from global to shared, completion on `bar`
VX_DEV void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
// all MMAs issued so far by this thread arrive (once) on `bar` when they have completed
VX_DEV void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
VX_DEV void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
VX_DEV void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
VX_DEV void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// warp-collective (one full warp); the base address lands in *slot (shared memory)
VX_DEV void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
VX_DEV void tmem_dealloc(uint32_t base, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
// 16 consecutive columns of this thread's TMEM lane (warp w of the CTA reads lanes 32*(w%4) ..): taddr = base + (lane0 << 16) + col
VX_DEV void tmem_ld16(uint32_t taddr, float (&r)[16]) {
  uint32_t q[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]), "=r"(q[9]),
        "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 16; ++j) r[j] = __uint_as_float(q[j]);
}
VX_DEV void tmem_ld32(uint32_t taddr, float (&r)[32]) {
  uint32_t q[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]), "=r"(q[9]),
        "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]), "=r"(q[16]), "=r"(q[17]), "=r"(q[18]),
        "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]), "=r"(q[25]), "=r"(q[26]), "=r"(q[27]),
        "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 32; ++j) r[j] = __uint_as_float(q[j]);
}
// 8-column variants (pieces of 8 keep the softmax passes of the attention kernel at 64 registers)
VX_DEV void tmem_ld8(uint32_t taddr, float (&r)[8]) {
  uint32_t q[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = __uint_as_float(q[j]);
}
VX_DEV void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
               "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
// the ".ts" form: A (128 x 8 tf32) read from tensor memory -- row m = lane m, element k = column tmem_a + k
VX_DEV void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 16 consecutive 32-bit columns of this thread's TMEM lane (same addressing as tmem_ld16); complete after tmem_wait_st()
VX_DEV void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
VX_DEV void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
#define VX_TC_SHARED_BARS(name, n) __shared__ __align__(8) uint64_t name[n]
#define VX_TC_SHARED_SLOT(name) __shared__ uint32_t name

#else
// ------------------------------------------------------------------------------------------------- software model
inline unsigned char* emu_base() { return vx_emu::dyn_smem() - 1024; }      // shared-window address 1024 = start of dynamic smem
inline uint32_t smem_addr(const void* p) { return (uint32_t)((const unsigned char*)p - emu_base()); }
inline void split(float x, float& hi, float& lo) {
  uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&hi, &u, 4); lo = x - hi;
}
inline float emu_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r; }
struct EmuTmem { float v[128][512]; };
inline EmuTmem& emu_tmem() { static EmuTmem t; return t; }
inline float emu_operand(uint64_t d, int mn_major, int r, int k) {
  const uint32_t start = (uint32_t)(d & 0x3FFFu) << 4, lbo = (uint32_t)((d >> 16) & 0x3FFFu) << 4, sbo = (uint32_t)((d >> 32) & 0x3FFFu) << 4;
  const uint32_t off = mn_major ? start + (uint32_t)(r / 4) * sbo + (uint32_t)(r % 4) * 4 + (uint32_t)(k % 8) * 16
                                : start + (uint32_t)(r / 8) * sbo + (uint32_t)(r % 8) * 16 + (uint32_t)(k / 4) * lbo + (uint32_t)(k % 4) * 4;
  if (off < 1024 || off + 4 > 1024 + 256 * 1024) { fprintf(stderr, "vx_emu: MMA operand outside shared memory (offset %u)\n", off); abort(); }
  float x; memcpy(&x, emu_base() + off, 4);
  return emu_tf32(x);
}
inline void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  const int N = (int)((idesc >> 17) & 0x3Fu) << 3, a_mn = (idesc >> 15) & 1, b_mn = (idesc >> 16) & 1;
  if (a_mn || b_mn) { fprintf(stderr, "vx_emu: MN-major operands do not work with kind::tf32 on the hardware\n"); abort(); }
  const int col = (int)(tmem_d & 0xFFFFu);
  if (col + N > 512) { fprintf(stderr, "vx_emu: MMA writes past TMEM column 512\n"); abort(); }
  float a[128][8];
  for (int m = 0; m < 128; ++m)
    for (int k = 0; k < 8; ++k) a[m][k] = emu_operand(adesc, a_mn, m, k);
  for (int n = 0; n < N; ++n) {
    float b[8];
    for (int k = 0; k < 8; ++k) b[k] = emu_operand(bdesc, b_mn, n, k);
    for (int m = 0; m < 128; ++m) {
      float acc = accumulate ? emu_tmem().v[m][col + n] : 0.f;
      for (int k = 0; k < 8; ++k) acc += a[m][k] * b[k];
      emu_tmem().v[m][col + n] = acc;
    }
  }
}
struct EmuBar { int init = 0, pending = 0; long long tx = 0; unsigned phase = 0; };
inline std::mutex& emu_bar_mutex() { static std::mutex m; return m; }
inline std::map<const void*, EmuBar>& emu_bars() { static std::map<const void*, EmuBar> m; return m; }
inline void emu_bar_check(EmuBar& b) { if (b.pending == 0 && b.tx == 0) { ++b.phase; b.pending = b.init; } }
inline void mbar_init(uint64_t* bar, uint32_t count) {
  std::lock_guard<std::mutex> g(emu_bar_mutex());
  EmuBar& b = emu_bars()[bar]; b.init = b.pending = (int)count; b.tx = 0; b.phase = 0;
}
inline void mbar_init_fence() {}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (long spins = 0;; ++spins) {
    { std::lock_guard<std::mutex> g(emu_bar_mutex()); if ((emu_bars()[bar].phase & 1u) != parity) return; }
    if (spins > 20000000L) { fprintf(stderr, "vx_emu: mbarrier wait timed out\n"); abort(); }
    std::this_thread::yield();
  }
}
inline void mbar_arrive(uint64_t* bar) { std::lock_guard<std::mutex> g(emu_bar_mutex()); EmuBar& b = emu_bars()[bar]; --b.pending; emu_bar_check(b); }
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  std::lock_guard<std::mutex> g(emu_bar_mutex()); EmuBar& b = emu_bars()[bar]; b.tx += bytes; --b.pending; emu_bar_check(b);
}
inline void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  if ((bytes & 15u) || ((uintptr_t)src & 15u) || (smem_addr(dst) & 15u)) { fprintf(stderr, "vx_emu: misaligned bulk copy\n"); abort(); }
  memcpy(dst, src, bytes);
  std::lock_guard<std::mutex> g(emu_bar_mutex()); EmuBar& b = emu_bars()[bar]; b.tx -= bytes; emu_bar_check(b);
}
inline void commit(uint64_t* bar) { mbar_arrive(bar); }       // the model's MMAs complete at issue
inline void fence_async_smem() {}
inline void fence_before() {}
inline void fence_after() {}
inline void tmem_alloc(uint32_t* slot, uint32_t) { *slot = 0; }
inline void tmem_dealloc(uint32_t, uint32_t) {}
inline void tmem_ld16(uint32_t taddr, float (&r)[16]) {
  const int lane = (int)(taddr >> 16) + (vx_emu::t_lane_slot & 31), col = (int)(taddr & 0xFFFFu);
  for (int j = 0; j < 16; ++j) r[j] = emu_tmem().v[lane][col + j];
}
inline void tmem_ld32(uint32_t taddr, float (&r)[32]) {
  const int lane = (int)(taddr >> 16) + (vx_emu::t_lane_slot & 31), col = (int)(taddr & 0xFFFFu);
  for (int j = 0; j < 32; ++j) r[j] = emu_tmem().v[lane][col + j];
}
inline void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  const int N = (int)((idesc >> 17) & 0x3Fu) << 3, b_mn = (idesc >> 16) & 1;
  if (b_mn) { fprintf(stderr, "vx_emu: MN-major operands do not work with kind::tf32 on the hardware\n"); abort(); }
  const int col = (int)(tmem_d & 0xFFFFu), acol = (int)(tmem_a & 0xFFFFu);
  if (col + N > 512 || acol + 8 > 512) { fprintf(stderr, "vx_emu: MMA touches TMEM past column 512\n"); abort(); }
  if (acol < col + N && col < acol + 8) { fprintf(stderr, "vx_emu: MMA accumulator overlaps its TMEM A operand\n"); abort(); }
  for (int n = 0; n < N; ++n) {
    float b[8];
    for (int k = 0; k < 8; ++k) b[k] = emu_operand(bdesc, 0, n, k);
    for (int m = 0; m < 128; ++m) {
      float acc = accumulate ? emu_tmem().v[m][col + n] : 0.f;
      for (int k = 0; k < 8; ++k) acc += emu_tf32(emu_tmem().v[m][acol + k]) * b[k];
      emu_tmem().v[m][col + n] = acc;
    }
  }
}
inline void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  const int lane = (int)(taddr >> 16) + (vx_emu::t_lane_slot & 31), col = (int)(taddr & 0xFFFFu);
  for (int j = 0; j < 16; ++j) emu_tmem().v[lane][col + j] = v[j];
}
inline void tmem_ld8(uint32_t taddr, float (&r)[8]) {
  const int lane = (int)(taddr >> 16) + (vx_emu::t_lane_slot & 31), col = (int)(taddr & 0xFFFFu);
  for (int j = 0; j < 8; ++j) r[j] = emu_tmem().v[lane][col + j];
}
inline void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  const int lane = (int)(taddr >> 16) + (vx_emu::t_lane_slot & 31), col = (int)(taddr & 0xFFFFu);
  for (int j = 0; j < 8; ++j) emu_tmem().v[lane][col + j] = v[j];
}
inline void tmem_wait_st() {}
#define VX_TC_SHARED_BARS(name, n) static uint64_t name[n]
#define VX_TC_SHARED_SLOT(name) static uint32_t name
#endif

}  // namespace tc
}  // namespace vx
