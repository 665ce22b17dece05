// JLC grouped convolutions (k = 1, 3, 5; c_g = 4 or 8 channels per group) as ONE implicit GEMM on the 5th-generation tensor
// cores -- tcgen05.mma kind::tf32, accumulators in tensor memory, fp32-accurate through the 3-term split of vx_tc.cuh.
//
// STATUS: candidate, OFF by default (vx_set_option(VX_OPT_JLC_CONV_TC, 1)).  Written after the round's GPU budget was
// spent: the indexing / operand layouts are checked on the CPU shim (tests/test_emu_kernels.py::test_jlc_conv_tensor_core,
// software model of the MMA on the same shared-memory layout); it has NOT run on a B200 yet.  Hardware questions to
// settle first: operand descriptors whose start address is 16-byte (not 128-byte) aligned, and LBO values that are not a
// multiple of 128 B.  On paper both are inside the canonical K-major no-swizzle form, in 16-byte units
// ((8, n), 2) : ((1, SBO), LBO) with free SBO / LBO (CuTe's mma_traits_sm100.hpp, "UmmaDescriptor Major-K", INTERLEAVE):
// here ((8, 16), 2) : ((1, 8), LBO).  The SIMT kernels in jlc.cu remain the product path until this one is parity-green and measured.
//
// GEMM view per CTA = (batch b, group g, brick of ZR planes x TY rows x full width):
//   D[m = flat padded position (128 per M-block)][n = z5 co 0..3 | z3 co 0..3 | z1 co 0..3 | 4 zero columns]
//     = sum over taps t = (dz, dy, dx) in 5^3 and ci in 0..3 of   X[m + shift(t)][ci] * Wn[t][ci]
// The halo brick (PZ = ZR+4 planes, PY = TY+4 rows, PX = W+4 columns, zero outside the volume) is staged ONCE as
// [position][4 channels] 16-byte rows, hi and lo copies.  In the K-major SWIZZLE_NONE layout a core matrix is 8 rows x
// 16 B and with SBO = 128 B the 128 rows of an operand are 128 consecutive positions, so the operand of tap t is the same
// brick read from  base + 16 * (p0 + shift(t)),  shift(t) = (dz*PY + dy)*PX + dx  -- a descriptor per tap, no im2col.
// One k-step (K = 8) carries two taps x 4 channels: the k-halves are LBO = 16 * (shift(tB) - shift(tA)) bytes apart.
// The 25 (dz, dy) pairs are taken two at a time (12 pairs + 1 single whose second half meets zero weights), for each of
// the 5 dx: 65 k-steps x 3 MMAs (lo*hi, hi*lo, hi*hi).  Weight rows of z3 / z1 are zero for taps outside 3^3 / the centre.
// Halo positions inside an M-block are computed and discarded.  Every M-block owns 16 TMEM columns, so one thread
// issues all MMAs of the CTA up front (one commit per block) and the warps drain blocks as they complete: bias, store
// of z1 / z3 / z5 (lanes along x), InstanceNorm partial statistics.
//
// MODE 1 is the data gradient  dx = sum_k corr(gz_k, flipped W_k) + dO  on the same machinery: three passes (k = 5, 3, 1),
// each staging the brick of one branch gradient (K axis = the group's 4 output channels) and that branch's weights as
// B[n = ci][k = co] with the taps mirrored, accumulating into the same TMEM columns; a pass barrier separates the MMAs of
// one pass from the restaging of the brick for the next.  The k = 3 / k = 1 passes walk only their own 15 / 1 k-steps.
#include "vx_kernels.h"
#include "vx_tc.cuh"

#ifdef VX_EMU
#define __grid_constant__
#endif

namespace vx {

constexpr int JT_THREADS = 256;
constexpr int JT_G0 = 8;           // guard positions before the brick (the first block reads 2 positions in front of it)
constexpr int JT_G1 = 144;         // guard positions behind it: 2 + 127 padding rows of the last block + 8 (second k-half of a single tap)
constexpr int JT_MAX_BLK = 32;     // 512 TMEM columns / 16

#ifdef VX_EMU
static float g_emu_tmem_jt[128][512];
static inline void jt_split(float x, float& hi, float& lo) {
  uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&hi, &u, 4); lo = x - hi;
}
#else
VX_DEV void jt_split(float x, float& hi, float& lo) { split_tf32(x, hi, lo); }
#endif

// TMEM columns per M-block = N of the MMA, and the most k-steps any pass stages
__host__ __device__ constexpr int jt_ncol(int mode, int cg) { return mode == 0 && cg == 8 ? 32 : 16; }
__host__ __device__ constexpr int jt_maxsteps(int cg) { return cg == 4 ? 65 : 27; }

// Pass p of a launch -> kernel size walked (5 / 3 / 1), dz slab (-1 = every dz) and whether the brick is restaged.
//   forward:        c_g 4: one pass over all taps;            c_g 8: five dz slabs of the 5^3 kernel over one brick
//   data gradient:  c_g 4: k = 5, 3, 1 (a brick each);        c_g 8: five slabs of k = 5, then k = 3, then k = 1
__host__ __device__ constexpr int jt_npass(int mode, int cg) { return mode == 0 ? (cg == 4 ? 1 : 5) : (cg == 4 ? 3 : 7); }
VX_DEV void jt_pass(int mode, int cg, int p, int& kind, int& slab, bool& restage) {
  if (cg == 4) { kind = mode == 0 ? 5 : (p == 0 ? 5 : p == 1 ? 3 : 1); slab = -1; restage = true; return; }
  if (mode == 0 || p < 5) { kind = 5; slab = p; restage = p == 0; return; }
  kind = p == 5 ? 3 : 1; slab = -1; restage = true;
}
VX_DEV int jt_nsteps(int cg, int kind, int slab) {
  if (cg == 4) return kind == 5 ? 65 : kind == 3 ? 15 : 1;
  return kind == 5 ? (slab >= 0 ? 25 : 125) : kind == 3 ? 27 : 1;
}
// k-step s -> column offset dx (0..4) and the (dz, dy) indices (0..24) of its taps (jb = -1: no second tap).
//   c_g 4: two taps per k-step.  kind 5: 13 (dz, dy) pairs x 5 dx; kind 3: the 9 inner (dz, dy) ascending, 5 pairs x 3 dx.
//   c_g 8: one tap per k-step.
VX_DEV void jt_step(int cg, int kind, int slab, int s, int& dxi, int& ja, int& jb) {
  jb = -1;
  if (kind == 1) { dxi = 2; ja = 12; return; }
  if (cg == 4) {
    if (kind == 5) {
      dxi = s % 5;
      const int j = s / 5;
      ja = 2 * j;
      jb = 2 * j + 1 < 25 ? 2 * j + 1 : -1;
    } else {
      dxi = 1 + s % 3;
      const int q = s / 3, i0 = 2 * q, i1 = 2 * q + 1;         // inner list index i -> j = (1 + i / 3) * 5 + 1 + i % 3
      ja = (1 + i0 / 3) * 5 + 1 + i0 % 3;
      jb = i1 < 9 ? (1 + i1 / 3) * 5 + 1 + i1 % 3 : -1;
    }
  } else if (kind == 5) {
    dxi = s % 5;
    ja = slab >= 0 ? slab * 5 + s / 5 : s / 5;
  } else {
    dxi = 1 + s % 3;
    ja = (1 + s / 9) * 5 + 1 + (s / 3) % 3;
  }
}
VX_DEV int jt_shift(int j, int dxi, int PY, int PX) { return ((j / 5 - 2) * PY + (j % 5 - 2)) * PX + (dxi - 2); }

template <int MODE, int CG>      // MODE 0: forward (x -> z1, z3, z5 + statistics), 1: data gradient (gz1, gz3, gz5, dO -> dx)
__global__ void __launch_bounds__(JT_THREADS) jlc_conv_tc_kernel(const __grid_constant__ JlcTcArgs A) {
  constexpr int NCH = CG / 4;                       // [position][4 channels] arrays of the brick
  constexpr int NCOL = jt_ncol(MODE, CG);
  constexpr int BSTEP = NCOL * 8;                   // floats of one k-step of the weight operand: [n-groups][2 k-halves][8 n][4 k]
  constexpr int MAXSTEPS = jt_maxsteps(CG);
  constexpr int NPASS = jt_npass(MODE, CG);
  const int g = blockIdx.y, b = blockIdx.z, tile = blockIdx.x;
  const int ty_i = tile % A.nty, tz_i = tile / A.nty;
  const int z0 = tz_i * A.ZR, y0 = ty_i * A.TY;
  const int D = A.D, H = A.H, W = A.W, C = A.C;
  const int PX = W + 4, PY = A.TY + 4, PZ = A.ZR + 4;
  const int NPOS = PZ * PY * PX;
  const int p_first = (2 * PY + 2) * PX;              // first output row of the first output plane, column 0
  const int nblk = A.nblk;
  const size_t S = (size_t)D * H * W;
  const size_t BCS = (size_t)A.B * C * S;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  VX_DYN_SMEM(float, sm);
  const int NALL = JT_G0 + NPOS + JT_G1;
  float* Xhi = sm;                                    // [NCH][NALL][4]
  float* Xlo = Xhi + (size_t)NCH * NALL * 4;
  float* Bhi = Xlo + (size_t)NCH * NALL * 4;          // [MAXSTEPS][BSTEP]
  float* Blo = Bhi + MAXSTEPS * BSTEP;
  float* sst = Blo + MAXSTEPS * BSTEP;                // [3 branches][CG co][2]

#ifndef VX_EMU
  __shared__ __align__(8) uint64_t mbar[JT_MAX_BLK + 1];      // one per M-block + the pass barrier
  __shared__ uint32_t tmem_slot;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"((uint32_t)A.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int i = 0; i < nblk; ++i) mbar_init(smem_u32(&mbar[i]), 1);
    mbar_init(smem_u32(&mbar[JT_MAX_BLK]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t tmem = 0;
#endif
  if (tid < 3 * CG * 2) sst[tid] = 0.f;

  // ---- guards: zero, written once (the brick is [position][4 channels] per chunk, hi / lo; zero outside the volume)
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = tid; i < NCH * (JT_G0 + JT_G1); i += JT_THREADS) {
    const int ch = i / (JT_G0 + JT_G1), r = i % (JT_G0 + JT_G1);
    const int pos = ch * NALL + (r < JT_G0 ? r : NPOS + r);
    reinterpret_cast<float4*>(Xhi)[pos] = zero4;
    reinterpret_cast<float4*>(Xlo)[pos] = zero4;
  }

#pragma unroll 1
  for (int pass = 0; pass < NPASS; ++pass) {
    int kind, slab;
    bool restage;
    jt_pass(MODE, CG, pass, kind, slab, restage);
    const int nsteps = jt_nsteps(CG, kind, slab);
    if (pass > 0) {      // the previous pass's MMAs have read the brick and the weights
#ifndef VX_EMU
      mbar_wait(smem_u32(&mbar[JT_MAX_BLK]), (uint32_t)((pass - 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#else
      __syncthreads();
#endif
    }
    // ---- the brick of this pass
    if (restage) {
      const float* xg = MODE == 0 ? A.x + ((size_t)b * C + g * CG) * S
                                  : A.gz + (size_t)(kind == 5 ? 2 : kind == 3 ? 1 : 0) * BCS + ((size_t)b * C + g * CG) * S;
      for (int it = tid; it < NCH * NPOS; it += JT_THREADS) {
        const int ch = it / NPOS, idx = it % NPOS;
        const int px = idx % PX, py = (idx / PX) % PY, pz = idx / (PX * PY);
        const int gz = z0 + pz - 2, gy = y0 + py - 2, gx = px - 2;
        float4 hi = zero4, lo = zero4;
        if (gz >= 0 && gz < D && gy >= 0 && gy < H && gx >= 0 && gx < W) {
          const float* xc = xg + (size_t)(ch * 4) * S + ((size_t)gz * H + gy) * W + gx;
          jt_split(__ldg(xc), hi.x, lo.x);
          jt_split(__ldg(xc + S), hi.y, lo.y);
          jt_split(__ldg(xc + 2 * S), hi.z, lo.z);
          jt_split(__ldg(xc + 3 * S), hi.w, lo.w);
        }
        reinterpret_cast<float4*>(Xhi)[ch * NALL + JT_G0 + idx] = hi;
        reinterpret_cast<float4*>(Xlo)[ch * NALL + JT_G0 + idx] = lo;
      }
    }
    // ---- the weights: element (n, k) of k-step s at s*BSTEP + (n/8)*64 + (k/4)*32 + (n%8)*4 + k%4  (LBO 128 B, SBO 256 B)
    for (int e = tid; e < nsteps * BSTEP; e += JT_THREADS) {
      const int s = e / BSTEP, r = e % BSTEP, n = r >> 3, k = r & 7;
      int dxi, ja, jb;
      jt_step(CG, kind, slab, s, dxi, ja, jb);
      const int j = CG == 4 ? ((k >> 2) ? jb : ja) : ja;      // c_g 4: the k-half picks the tap; c_g 8: k is the channel
      const int kc = CG == 4 ? (k & 3) : k;
      float w = 0.f;
      if (MODE == 0) {       // columns z5 | z3 | z1, k = input channel
        if (j >= 0 && n < 3 * CG) {
          const int dzi = j / 5, dyi = j % 5, br = n / CG, co = g * CG + n % CG;
          if (br == 0) {
            w = __ldg(A.w5 + ((size_t)co * CG + kc) * 125 + (dzi * 5 + dyi) * 5 + dxi);
          } else if (br == 1) {
            if (dzi >= 1 && dzi <= 3 && dyi >= 1 && dyi <= 3 && dxi >= 1 && dxi <= 3)
              w = __ldg(A.w3 + ((size_t)co * CG + kc) * 27 + ((dzi - 1) * 3 + (dyi - 1)) * 3 + (dxi - 1));
          } else if (dzi == 2 && dyi == 2 && dxi == 2) {
            w = __ldg(A.w1 + (size_t)co * CG + kc);
          }
        }
      } else if (j >= 0 && n < CG) {      // columns = input channel n, k = output channel; taps mirrored about the centre
        const int P = kind >> 1;
        const int tz = P - (j / 5 - 2), ty = P - (j % 5 - 2), tx = P - (dxi - 2);      // inside [0, kind) by construction
        const float* wk = kind == 5 ? A.w5 : kind == 3 ? A.w3 : A.w1;
        w = __ldg(wk + ((size_t)(g * CG + kc) * CG + n) * (kind * kind * kind) + (tz * kind + ty) * kind + tx);
      }
      float hi, lo;
      jt_split(w, hi, lo);
      const int o = s * BSTEP + (n >> 3) * 64 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3);
      Bhi[o] = hi; Blo[o] = lo;
    }

#ifndef VX_EMU
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tmem = tmem_slot;
    if (tid == 0) {
      const uint32_t idesc = umma_idesc_tf32(NCOL);
      const uint32_t x_hi = smem_u32(Xhi), x_lo = smem_u32(Xlo), b_hi = smem_u32(Bhi), b_lo = smem_u32(Blo);
#pragma unroll 1
      for (int blk = 0; blk < nblk; ++blk) {
        const int p0 = JT_G0 + p_first + blk * 128;
#pragma unroll 1
        for (int s = 0; s < nsteps; ++s) {
          int dxi, ja, jb;
          jt_step(CG, kind, slab, s, dxi, ja, jb);
          const int sa = jt_shift(ja, dxi, PY, PX);
          // distance between the k-halves: c_g 4: the second tap (8 positions when it is absent and meets zero weights);
          // c_g 8: the next [position][4 channels] array
          const uint32_t lbo = CG == 8 ? (uint32_t)NALL * 16u : jb >= 0 ? (uint32_t)(jt_shift(jb, dxi, PY, PX) - sa) * 16u : 128u;
          const uint32_t ao = (uint32_t)(p0 + sa) * 16u, bo = (uint32_t)s * (BSTEP * 4);
          const uint64_t dah = umma_desc(x_hi + ao, lbo, 128u), dal = umma_desc(x_lo + ao, lbo, 128u);
          const uint64_t dbh = umma_desc(b_hi + bo, 128u, 256u), dbl = umma_desc(b_lo + bo, 128u, 256u);
          const uint32_t d = tmem + (uint32_t)(blk * NCOL);
          umma_tf32(d, dal, dbh, idesc, (pass > 0 || s > 0) ? 1u : 0u);
          umma_tf32(d, dah, dbl, idesc, 1u);
          umma_tf32(d, dah, dbh, idesc, 1u);
        }
        if (pass == NPASS - 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar[blk]))
                       : "memory");
      }
      if (pass < NPASS - 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         smem_u32(&mbar[JT_MAX_BLK]))
                     : "memory");
    }
    __syncwarp();
#else
    __syncthreads();
    if (tid == 0) {      // software model of the MMAs: same shared layout, same descriptor arithmetic
      for (int blk = 0; blk < nblk; ++blk) {
        const int p0 = JT_G0 + p_first + blk * 128;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < NCOL; ++n) {
            float acc = pass > 0 ? g_emu_tmem_jt[m][blk * NCOL + n] : 0.f;
            for (int s = 0; s < nsteps; ++s) {
              int dxi, ja, jb;
              jt_step(CG, kind, slab, s, dxi, ja, jb);
              const int sa = jt_shift(ja, dxi, PY, PX);
              const int lbo_pos = CG == 8 ? NALL : jb >= 0 ? jt_shift(jb, dxi, PY, PX) - sa : 8;
              for (int k = 0; k < 8; ++k) {
                const int ao = (p0 + sa + m + (k >> 2) * lbo_pos) * 4 + (k & 3);
                const int bo = s * BSTEP + (n >> 3) * 64 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3);
                acc += Xlo[ao] * Bhi[bo] + Xhi[ao] * Blo[bo] + Xhi[ao] * Bhi[bo];
              }
            }
            g_emu_tmem_jt[m][blk * NCOL + n] = acc;
          }
      }
    }
    __syncthreads();
#endif
  }

  // ---- drain: warp = (TMEM lane quadrant, parity of the blocks it takes); thread = one position of the block
  const int wq = warp & 3, wp = warp >> 2;
  float bias[3][CG];
#pragma unroll
  for (int c = 0; c < CG; ++c) {
    bias[0][c] = bias[1][c] = bias[2][c] = 0.f;
    if (MODE == 0) {      // the data-gradient launch carries no bias pointers
      bias[0][c] = __ldg(A.b1 + g * CG + c); bias[1][c] = __ldg(A.b3 + g * CG + c); bias[2][c] = __ldg(A.b5 + g * CG + c);
    }
  }
  float ssum[3][CG], ssq[3][CG];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int c = 0; c < CG; ++c) { ssum[k][c] = 0.f; ssq[k][c] = 0.f; }
#pragma unroll 1
  for (int blk = wp; blk < nblk; blk += 2) {
    float r[NCOL];
#ifndef VX_EMU
    mbar_wait(smem_u32(&mbar[blk]), 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int c0 = 0; c0 < NCOL; c0 += 16) {
      uint32_t q[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
            "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
          : "r"(tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(blk * NCOL + c0))
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 16; ++j) r[c0 + j] = __uint_as_float(q[j]);
    }
#else
    for (int j = 0; j < NCOL; ++j) r[j] = g_emu_tmem_jt[wq * 32 + lane][blk * NCOL + j];
#endif
    const int p = p_first + blk * 128 + wq * 32 + lane;
    const int px = p % PX, py = (p / PX) % PY, pz = p / (PX * PY);
    const int gz = z0 + pz - 2, gy = y0 + py - 2, gx = px - 2;
    const bool ok = px >= 2 && px < 2 + W && py >= 2 && py < 2 + A.TY && pz >= 2 && pz < 2 + A.ZR && gz < D && gy < H;
    if (ok && MODE == 1) {
      const size_t o = ((size_t)b * C + g * CG) * S + ((size_t)gz * H + gy) * W + gx;
#pragma unroll
      for (int c = 0; c < CG; ++c) A.dx[o + c * S] = r[c] + __ldg(A.dO + o + c * S);
    } else if (ok) {
      const size_t o = ((size_t)b * C + g * CG) * S + ((size_t)gz * H + gy) * W + gx;
#pragma unroll
      for (int c = 0; c < CG; ++c) {
        const float r1 = r[2 * CG + c] + bias[0][c], r3 = r[CG + c] + bias[1][c], r5 = r[c] + bias[2][c];
        A.z[o + c * S] = r1; A.z[BCS + o + c * S] = r3; A.z[2 * BCS + o + c * S] = r5;
        ssum[0][c] += r1; ssq[0][c] = fmaf(r1, r1, ssq[0][c]);
        ssum[1][c] += r3; ssq[1][c] = fmaf(r3, r3, ssq[1][c]);
        ssum[2][c] += r5; ssq[2][c] = fmaf(r5, r5, ssq[2][c]);
      }
    }
  }
  if (MODE == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int c = 0; c < CG; ++c) {
        const float s1 = warp_sum(ssum[k][c]), s2 = warp_sum(ssq[k][c]);
        if (lane == 0) { atomicAdd(sst + (k * CG + c) * 2, s1); atomicAdd(sst + (k * CG + c) * 2 + 1, s2); }
      }
  }
#ifndef VX_EMU
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
#endif
  __syncthreads();
  const int ntiles = gridDim.x;
  if (MODE == 0 && tid < 3 * CG) {
    const int k = tid / CG, c = tid % CG;
    const size_t row = (size_t)k * A.B * C + (size_t)b * C + g * CG + c;
    float* pp = A.part + (row * ntiles + tile) * 2;
    pp[0] = sst[tid * 2]; pp[1] = sst[tid * 2 + 1];
  }
#ifndef VX_EMU
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)A.tmem_cols) : "memory");
  }
#endif
}

static int g_jt_enabled = 0;
void jlc_tc_set(int enabled) { g_jt_enabled = enabled ? 1 : 0; }

static size_t jt_smem_bytes(int CG, int ZR, int TY, int W) {
  const size_t nall = JT_G0 + (size_t)(ZR + 4) * (TY + 4) * (W + 4) + JT_G1;
  // the forward launch has the wider weight tile; both launches of an op share the geometry
  return sizeof(float) * (2 * (CG / 4) * nall * 4 + 2 * (size_t)jt_maxsteps(CG) * jt_ncol(0, CG) * 8 + 3 * CG * 2);
}

// Picks the brick (ZR planes x TY rows x full width) of the tensor-core conv kernels; returns the number of bricks per
// (batch, group) = statistics partials per row, or 0 when the path is off or the shape does not qualify.
int jlc_tc_geo(int B, int groups, int CG, int D, int H, int W, JlcTcArgs& A) {
  if (!g_jt_enabled || (CG != 4 && CG != 8) || W + 4 > 128) return 0;
  const int ncol = jt_ncol(0, CG);
  const double mma_cycles = CG == 4 ? 195.0 * 16.0 : 375.0 * 20.0;      // per M-block (guess until measured)
  double best = -1.0;
  const int zr_c[] = {1, 2, 3, 4, 6, 8};
  for (int div = 1; div <= 4; ++div) {
    const int TY = cdiv(H, div);
    for (int ZR : zr_c) {
      if (ZR > D && ZR != 1) continue;
      const int PX = W + 4, PY = TY + 4;
      const long long nall = JT_G0 + (long long)(ZR + 4) * PY * PX + JT_G1;
      if (jt_smem_bytes(CG, ZR, TY, W) > 227 * 1024) continue;
      const int nblk = cdiv((long long)((ZR - 1) * PY + TY) * PX, 128);
      if (nblk > JT_MAX_BLK || nblk * ncol > 512) continue;
      if ((PY - 4) * PX > 16383 || nall > 16383) continue;          // LBO field: 14 bits of 16-byte units
      const int ntz = cdiv(D, ZR), nty = cdiv(H, TY);
      const long long ncta = (long long)ntz * nty * groups * B;
      const double cost = (double)cdiv(ncta, kSMs) * ((double)nblk * mma_cycles + (double)(CG / 4) * (ZR + 4) * PY * PX * 6.0);
      if (best < 0.0 || cost < best) {
        best = cost;
        A.ZR = ZR; A.TY = TY; A.ntz = ntz; A.nty = nty; A.nblk = nblk;
        A.tmem_cols = 32;
        while (A.tmem_cols < nblk * ncol) A.tmem_cols <<= 1;
      }
    }
  }
  return best >= 0.0 ? A.ntz * A.nty : 0;
}

template <int MODE>
static int jt_launch(const JlcTcArgs& A, int groups, cudaStream_t st, const char* what) {
  const int CG = A.C / groups;
  const size_t smem = jt_smem_bytes(CG, A.ZR, A.TY, A.W);
  const dim3 grid(A.ntz * A.nty, groups, A.B);
  if (CG == 4) {
    VX_SET_SMEM((jlc_conv_tc_kernel<MODE, 4>), smem);
    VX_LAUNCH((jlc_conv_tc_kernel<MODE, 4>), grid, dim3(JT_THREADS), smem, st, A);
  } else {
    VX_SET_SMEM((jlc_conv_tc_kernel<MODE, 8>), smem);
    VX_LAUNCH((jlc_conv_tc_kernel<MODE, 8>), grid, dim3(JT_THREADS), smem, st, A);
  }
  return check_launch(what);
}

int jlc_conv_tc_fwd(const JlcTcArgs& A, int groups, cudaStream_t st) { return jt_launch<0>(A, groups, st, "jlc_conv_tc_kernel<0>"); }
int jlc_conv_tc_dgrad(const JlcTcArgs& A, int groups, cudaStream_t st) { return jt_launch<1>(A, groups, st, "jlc_conv_tc_kernel<1>"); }

}  // namespace vx
