// JLC grouped convolutions (k = 1, 3, 5, c_g = 4 channels per group) as ONE implicit GEMM on the 5th-generation tensor
// cores -- tcgen05.mma kind::tf32, accumulators in tensor memory, fp32-accurate through the 3-term split of vx_tc.cuh.
//
// STATUS: candidate, OFF by default (vx_set_option(VX_OPT_JLC_CONV_TC, 1)).  Written after the round's GPU budget was
// spent: the indexing / operand layouts are checked on the CPU shim (tests/test_emu_kernels.py::test_jlc_conv_tensor_core,
// software model of the MMA on the same shared-memory layout); it has NOT run on a B200 yet.  Hardware questions to
// settle first: operand descriptors whose start address is 16-byte (not 128-byte) aligned, and LBO values that are not a
// multiple of 128 B.  The SIMT kernels in jlc.cu remain the product path until this one is parity-green and measured.
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
constexpr int JT_KSTEPS = 65;      // 13 (dz, dy) pairs x 5 dx
constexpr int JT_N = 16;           // z5 | z3 | z1 | zero
constexpr int JT_G0 = 8;           // guard positions before the brick (the first block reads 2 positions in front of it)
constexpr int JT_G1 = 144;         // guard positions behind it: 2 + 127 padding rows of the last block + 8 (second k-half of the single tap)
constexpr int JT_BSTEP = 128;      // floats of one k-step of the weight operand: [2 n-groups][2 k-halves][8 n][4 k]
constexpr int JT_MAX_BLK = 32;     // 512 TMEM columns / 16

#ifdef VX_EMU
static float g_emu_tmem_jt[128][512];
static inline void jt_split(float x, float& hi, float& lo) {
  uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&hi, &u, 4); lo = x - hi;
}
#else
VX_DEV void jt_split(float x, float& hi, float& lo) { split_tf32(x, hi, lo); }
#endif

// k-step s of a pass over a k x k x k kernel -> column offset dx (0..4) and the two (dz, dy) indices (0..24, -1 = absent).
// kind 5: 13 pairs x 5 dx; kind 3: the 9 inner (dz, dy) in ascending order, 5 pairs x 3 dx; kind 1: the centre.
VX_DEV int jt_nsteps(int kind) { return kind == 5 ? JT_KSTEPS : kind == 3 ? 15 : 1; }
VX_DEV void jt_step(int kind, int s, int& dxi, int& ja, int& jb) {
  if (kind == 5) {
    dxi = s % 5;
    const int j = s / 5;
    ja = 2 * j;
    jb = 2 * j + 1 < 25 ? 2 * j + 1 : -1;
  } else if (kind == 3) {
    dxi = 1 + s % 3;
    const int q = s / 3, i0 = 2 * q, i1 = 2 * q + 1;           // inner list index i -> j = (1 + i / 3) * 5 + 1 + i % 3
    ja = (1 + i0 / 3) * 5 + 1 + i0 % 3;
    jb = i1 < 9 ? (1 + i1 / 3) * 5 + 1 + i1 % 3 : -1;
  } else {
    dxi = 2; ja = 12; jb = -1;
  }
}
VX_DEV int jt_shift(int j, int dxi, int PY, int PX) { return ((j / 5 - 2) * PY + (j % 5 - 2)) * PX + (dxi - 2); }

template <int MODE>      // 0: forward (x -> z1, z3, z5 + statistics), 1: data gradient (gz1, gz3, gz5, dO -> dx)
__global__ void __launch_bounds__(JT_THREADS) jlc_conv_tc_kernel(const __grid_constant__ JlcTcArgs A) {
  const int g = blockIdx.y, b = blockIdx.z, tile = blockIdx.x;
  const int ty_i = tile % A.nty, tz_i = tile / A.nty;
  const int z0 = tz_i * A.ZR, y0 = ty_i * A.TY;
  const int D = A.D, H = A.H, W = A.W, C = A.C;
  const int PX = W + 4, PY = A.TY + 4, PZ = A.ZR + 4;
  const int NPOS = PZ * PY * PX;
  const int p_first = (2 * PY + 2) * PX;              // first output row of the first output plane, column 0
  const int nblk = A.nblk;
  const size_t S = (size_t)D * H * W;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  VX_DYN_SMEM(float, sm);
  const int NALL = JT_G0 + NPOS + JT_G1;
  float* Xhi = sm;                                    // [NALL][4]
  float* Xlo = Xhi + (size_t)NALL * 4;
  float* Bhi = Xlo + (size_t)NALL * 4;                // [65][128]
  float* Blo = Bhi + JT_KSTEPS * JT_BSTEP;
  float* sst = Blo + JT_KSTEPS * JT_BSTEP;            // [3 branches][4 co][2]

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
#endif
  if (tid < 24) sst[tid] = 0.f;

  // ---- the brick: [position][4 channels], hi / lo; zero outside the volume and in the guards
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = tid; i < JT_G0 + JT_G1; i += JT_THREADS) {
    const int pos = i < JT_G0 ? i : NPOS + i;
    reinterpret_cast<float4*>(Xhi)[pos] = zero4;
    reinterpret_cast<float4*>(Xlo)[pos] = zero4;
  }
  constexpr int NPASS = MODE == 0 ? 1 : 3;
  const size_t BCS = (size_t)A.B * C * S;
#ifndef VX_EMU
  uint32_t tmem = 0;
#endif
#pragma unroll 1
  for (int pass = 0; pass < NPASS; ++pass) {
    const int kind = MODE == 0 ? 5 : (pass == 0 ? 5 : pass == 1 ? 3 : 1);
    const int nsteps = jt_nsteps(kind);
    if (pass > 0) {      // the previous pass's MMAs have read the brick and the weights
#ifndef VX_EMU
      mbar_wait(smem_u32(&mbar[JT_MAX_BLK]), (uint32_t)((pass - 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#else
      __syncthreads();
#endif
    }
    // ---- the brick of this pass
    const float* xg = MODE == 0 ? A.x + ((size_t)b * C + g * 4) * S
                                : A.gz + (size_t)(kind == 5 ? 2 : kind == 3 ? 1 : 0) * BCS + ((size_t)b * C + g * 4) * S;
    for (int idx = tid; idx < NPOS; idx += JT_THREADS) {
      const int px = idx % PX, py = (idx / PX) % PY, pz = idx / (PX * PY);
      const int gz = z0 + pz - 2, gy = y0 + py - 2, gx = px - 2;
      float4 hi = zero4, lo = zero4;
      if (gz >= 0 && gz < D && gy >= 0 && gy < H && gx >= 0 && gx < W) {
        const size_t o = ((size_t)gz * H + gy) * W + gx;
        jt_split(__ldg(xg + o), hi.x, lo.x);
        jt_split(__ldg(xg + S + o), hi.y, lo.y);
        jt_split(__ldg(xg + 2 * S + o), hi.z, lo.z);
        jt_split(__ldg(xg + 3 * S + o), hi.w, lo.w);
      }
      reinterpret_cast<float4*>(Xhi)[JT_G0 + idx] = hi;
      reinterpret_cast<float4*>(Xlo)[JT_G0 + idx] = lo;
    }
    // ---- the weights: element (n, k) of k-step s at s*128 + (n/8)*64 + (k/4)*32 + (n%8)*4 + k%4  (LBO 128 B, SBO 256 B)
    for (int e = tid; e < nsteps * JT_BSTEP; e += JT_THREADS) {
      const int s = e >> 7, r = e & 127, n = r >> 3, k = r & 7;
      int dxi, ja, jb;
      jt_step(kind, s, dxi, ja, jb);
      const int j = (k >> 2) ? jb : ja, kc = k & 3;
      float w = 0.f;
      if (MODE == 0) {       // columns z5 | z3 | z1, k = input channel
        if (j >= 0 && n < 12) {
          const int dzi = j / 5, dyi = j % 5, br = n >> 2, co = g * 4 + (n & 3);
          if (br == 0) {
            w = __ldg(A.w5 + ((size_t)co * 4 + kc) * 125 + (dzi * 5 + dyi) * 5 + dxi);
          } else if (br == 1) {
            if (dzi >= 1 && dzi <= 3 && dyi >= 1 && dyi <= 3 && dxi >= 1 && dxi <= 3)
              w = __ldg(A.w3 + ((size_t)co * 4 + kc) * 27 + ((dzi - 1) * 3 + (dyi - 1)) * 3 + (dxi - 1));
          } else if (dzi == 2 && dyi == 2 && dxi == 2) {
            w = __ldg(A.w1 + (size_t)co * 4 + kc);
          }
        }
      } else if (j >= 0 && n < 4) {      // columns = input channel n, k = output channel; taps mirrored about the centre
        const int P = kind >> 1;
        const int tz = P - (j / 5 - 2), ty = P - (j % 5 - 2), tx = P - (dxi - 2);      // inside [0, kind) by construction
        const float* wk = kind == 5 ? A.w5 : kind == 3 ? A.w3 : A.w1;
        w = __ldg(wk + ((size_t)(g * 4 + kc) * 4 + n) * (kind * kind * kind) + (tz * kind + ty) * kind + tx);
      }
      float hi, lo;
      jt_split(w, hi, lo);
      const int o = s * JT_BSTEP + (n >> 3) * 64 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3);
      Bhi[o] = hi; Blo[o] = lo;
    }

#ifndef VX_EMU
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tmem = tmem_slot;
    if (tid == 0) {
      const uint32_t idesc = umma_idesc_tf32(JT_N);
      const uint32_t x_hi = smem_u32(Xhi), x_lo = smem_u32(Xlo), b_hi = smem_u32(Bhi), b_lo = smem_u32(Blo);
#pragma unroll 1
      for (int blk = 0; blk < nblk; ++blk) {
        const int p0 = JT_G0 + p_first + blk * 128;
#pragma unroll 1
        for (int s = 0; s < nsteps; ++s) {
          int dxi, ja, jb;
          jt_step(kind, s, dxi, ja, jb);
          const int sa = jt_shift(ja, dxi, PY, PX);
          const uint32_t lbo = jb >= 0 ? (uint32_t)(jt_shift(jb, dxi, PY, PX) - sa) * 16u : 128u;
          const uint32_t ao = (uint32_t)(p0 + sa) * 16u, bo = (uint32_t)s * (JT_BSTEP * 4);
          const uint64_t dah = umma_desc(x_hi + ao, lbo, 128u), dal = umma_desc(x_lo + ao, lbo, 128u);
          const uint64_t dbh = umma_desc(b_hi + bo, 128u, 256u), dbl = umma_desc(b_lo + bo, 128u, 256u);
          const uint32_t d = tmem + (uint32_t)(blk * JT_N);
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
          for (int n = 0; n < JT_N; ++n) {
            float acc = pass > 0 ? g_emu_tmem_jt[m][blk * JT_N + n] : 0.f;
            for (int s = 0; s < nsteps; ++s) {
              int dxi, ja, jb;
              jt_step(kind, s, dxi, ja, jb);
              const int sa = jt_shift(ja, dxi, PY, PX);
              const int lbo_pos = jb >= 0 ? jt_shift(jb, dxi, PY, PX) - sa : 8;
              for (int k = 0; k < 8; ++k) {
                const int ao = (p0 + sa + m + (k >> 2) * lbo_pos) * 4 + (k & 3);
                const int bo = s * JT_BSTEP + (n >> 3) * 64 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3);
                acc += Xlo[ao] * Bhi[bo] + Xhi[ao] * Blo[bo] + Xhi[ao] * Bhi[bo];
              }
            }
            g_emu_tmem_jt[m][blk * JT_N + n] = acc;
          }
      }
    }
    __syncthreads();
#endif
  }

  // ---- drain: warp = (TMEM lane quadrant, parity of the blocks it takes); thread = one position of the block
  const int wq = warp & 3, wp = warp >> 2;
  float bias[3][4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    bias[0][c] = bias[1][c] = bias[2][c] = 0.f;
    if (MODE == 0) {      // the data-gradient launch carries no bias pointers
      bias[0][c] = __ldg(A.b1 + g * 4 + c); bias[1][c] = __ldg(A.b3 + g * 4 + c); bias[2][c] = __ldg(A.b5 + g * 4 + c);
    }
  }
  float ssum[3][4], ssq[3][4];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int c = 0; c < 4; ++c) { ssum[k][c] = 0.f; ssq[k][c] = 0.f; }
#pragma unroll 1
  for (int blk = wp; blk < nblk; blk += 2) {
    float r[JT_N];
#ifndef VX_EMU
    mbar_wait(smem_u32(&mbar[blk]), 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t q[JT_N];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
          "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
        : "r"(tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(blk * JT_N))
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < JT_N; ++j) r[j] = __uint_as_float(q[j]);
#else
    for (int j = 0; j < JT_N; ++j) r[j] = g_emu_tmem_jt[wq * 32 + lane][blk * JT_N + j];
#endif
    const int p = p_first + blk * 128 + wq * 32 + lane;
    const int px = p % PX, py = (p / PX) % PY, pz = p / (PX * PY);
    const int gz = z0 + pz - 2, gy = y0 + py - 2, gx = px - 2;
    const bool ok = px >= 2 && px < 2 + W && py >= 2 && py < 2 + A.TY && pz >= 2 && pz < 2 + A.ZR && gz < D && gy < H;
    if (ok && MODE == 1) {
      const size_t o = ((size_t)b * C + g * 4) * S + ((size_t)gz * H + gy) * W + gx;
#pragma unroll
      for (int c = 0; c < 4; ++c) A.dx[o + c * S] = r[c] + __ldg(A.dO + o + c * S);
    } else if (ok) {
      const size_t o = ((size_t)b * C + g * 4) * S + ((size_t)gz * H + gy) * W + gx;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float r1 = r[8 + c] + bias[0][c], r3 = r[4 + c] + bias[1][c], r5 = r[c] + bias[2][c];
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
      for (int c = 0; c < 4; ++c) {
        const float s1 = warp_sum(ssum[k][c]), s2 = warp_sum(ssq[k][c]);
        if (lane == 0) { atomicAdd(sst + (k * 4 + c) * 2, s1); atomicAdd(sst + (k * 4 + c) * 2 + 1, s2); }
      }
  }
#ifndef VX_EMU
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
#endif
  __syncthreads();
  const int ntiles = gridDim.x;
  if (MODE == 0 && tid < 12) {
    const int k = tid >> 2, c = tid & 3;
    const size_t row = (size_t)k * A.B * C + (size_t)b * C + g * 4 + c;
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

static size_t jt_smem_bytes(int ZR, int TY, int W) {
  const size_t npos = (size_t)(ZR + 4) * (TY + 4) * (W + 4);
  return sizeof(float) * (2 * 4 * (JT_G0 + npos + JT_G1) + 2 * JT_KSTEPS * JT_BSTEP + 24);
}

// Picks the brick (ZR planes x TY rows x full width) of the tensor-core forward conv; returns 0 when the path is off or
// the shape does not qualify (c_g != 4, rows wider than one M-block).
int jlc_tc_geo(int B, int groups, int CG, int D, int H, int W, JlcTcArgs& A) {
  if (!g_jt_enabled || CG != 4 || W + 4 > 128) return 0;
  double best = -1.0;
  const int zr_c[] = {1, 2, 3, 4, 6, 8};
  for (int div = 1; div <= 4; ++div) {
    const int TY = cdiv(H, div);
    for (int ZR : zr_c) {
      if (ZR > D && ZR != 1) continue;
      const int PX = W + 4, PY = TY + 4;
      if (jt_smem_bytes(ZR, TY, W) > 227 * 1024) continue;
      const int nblk = cdiv((long long)((ZR - 1) * PY + TY) * PX, 128);
      if (nblk > JT_MAX_BLK) continue;
      if ((PY - 4) * PX > 16383) continue;                           // LBO field: 14 bits of 16-byte units
      const int ntz = cdiv(D, ZR), nty = cdiv(H, TY);
      const long long ncta = (long long)ntz * nty * groups * B;
      const double cost = (double)cdiv(ncta, kSMs) * ((double)nblk * 3200.0 + (double)(ZR + 4) * PY * PX * 6.0);
      if (best < 0.0 || cost < best) {
        best = cost;
        A.ZR = ZR; A.TY = TY; A.ntz = ntz; A.nty = nty; A.nblk = nblk;
        A.tmem_cols = 32;
        while (A.tmem_cols < nblk * JT_N) A.tmem_cols <<= 1;
      }
    }
  }
  return best >= 0.0 ? A.ntz * A.nty : 0;
}

int jlc_conv_tc_fwd(const JlcTcArgs& A, int groups, cudaStream_t st) {
  const size_t smem = jt_smem_bytes(A.ZR, A.TY, A.W);
  VX_SET_SMEM(jlc_conv_tc_kernel<0>, smem);
  VX_LAUNCH(jlc_conv_tc_kernel<0>, dim3(A.ntz * A.nty, groups, A.B), dim3(JT_THREADS), smem, st, A);
  return check_launch("jlc_conv_tc_kernel<0>");
}

int jlc_conv_tc_dgrad(const JlcTcArgs& A, int groups, cudaStream_t st) {
  const size_t smem = jt_smem_bytes(A.ZR, A.TY, A.W);
  VX_SET_SMEM(jlc_conv_tc_kernel<1>, smem);
  VX_LAUNCH(jlc_conv_tc_kernel<1>, dim3(A.ntz * A.nty, groups, A.B), dim3(JT_THREADS), smem, st, A);
  return check_launch("jlc_conv_tc_kernel<1>");
}

}  // namespace vx
