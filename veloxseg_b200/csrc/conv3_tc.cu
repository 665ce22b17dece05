// Dense 3x3x3 convolution with 16 input channels -- decoder.out_conv1 and the reconstruction out_conv
// (model/Decoder.py:73-76,150-153; 44-60 % of the forward MACs) -- forward, data gradient and
// weight gradient as implicit GEMMs on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with the 3-term hi/lo split
// (x_lo*w_hi + x_hi*w_lo + x_hi*w_hi, fp32 accumulation in tensor memory), i.e. fp32-accurate (~2e-6 relative), not TF32.
// PixelShuffle(4) (components/superpixel.py:15) and the bias are fused into the forward epilogue, the inverse shuffle into
// the operand staging of both backward kernels, so the (B, C_out, D, H, W) intermediate never exists in memory.
//
// Measured on a B200 (tools/bringup/tc_probe2.cu, profiles/r2a_tc_probe2.txt): a tcgen05.mma M = 128, K = 8 (tf32) costs
// max(64, N/2) cycles -- N below 128 wastes the tensor pipe -- and K-major SWIZZLE_NONE operand descriptors may start at any
// 16-byte address.  All three kernels are therefore shaped to N >= 128 per instruction:
//
//   forward   D[m = padded position][n = co tile of 128] : A = the halo brick of x staged ONCE as [4-channel chunk][position][4]
//             16-byte rows (hi and lo copies); tap (dz,dy,dx) is the same brick read from  base + 16 * (p0 + shift(tap)) -- one
//             descriptor per tap, no im2col.  B = weights, pre-split and pre-arranged by a prep kernel into the exact
//             shared-memory image and streamed per (dz,dy) group with cp.async.bulk into a 2-deep mbarrier ring.
//   wgrad     D[m = co][n = (tap, ci) 27 x 16 | 16 bias columns] over k = positions, 8 per step: A = dz rows (the inverse pixel
//             shuffle is a 4x4 register transpose of 16 consecutive floats of dy), B = im2col rows built in shared memory
//             from a raw x brick; persistent CTAs accumulate their share of the positions in one 128 x 448 TMEM tile,
//             partial tiles go to a workspace and a second kernel folds them in a fixed order (deterministic).
//   dgrad     D[m = padded position][n = (dy,dx, ci) 9 x 16] accumulated over dz planes with shifted descriptors and over
//             co in passes of 8 (brick and weight ring double-buffered); the 9 in-plane taps are column blocks whose rows
//             are then summed at shifted positions through shared memory (col2im restricted to one plane).
#include "vx_kernels.h"
#include "vx_tc2.cuh"

#ifdef VX_EMU
#define __grid_constant__
#endif

namespace vx {

constexpr int C3_THREADS = 256;
constexpr int C3_G0 = 8;             // guard positions in front of a brick (first block reads up to 1 position before it)
constexpr int C3_G1 = 136;           // behind it: 1 + the 127 padding rows of the last M-block, rounded up
constexpr int C3_MAXBLK = 4;

// Phase timestamps of CTA 0 (developer diagnostics, tools/conv3_phases.py): set through vx_set_option(VX_OPT_CONV3_TRACE, 1).
#ifndef VX_EMU
__device__ long long g_c3_trace[64];
static int g_c3_trace_on = 0;
#define C3_STAMP(on, slot) do { if ((on) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) g_c3_trace[slot] = clock64(); } while (0)
#else
#define C3_STAMP(on, slot) do { } while (0)
#endif
void conv3_trace_set(int on) {
#ifndef VX_EMU
  g_c3_trace_on = on;
#endif
}
int conv3_trace_read(long long* out, int n) {
#ifndef VX_EMU
  return cudaMemcpyFromSymbol(out, g_c3_trace, sizeof(long long) * (n < 64 ? n : 64)) == cudaSuccess ? 0 : -1;
#else
  return -1;
#endif
}

VX_DEV float4 c3_ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
VX_DEV void c3_split4(const float4 v, float4& hi, float4& lo) {
  tc::split(v.x, hi.x, lo.x); tc::split(v.y, hi.y, lo.y); tc::split(v.z, hi.z, lo.z); tc::split(v.w, hi.w, lo.w);
}
// prec 0: fp32-accurate hi / lo split; prec 1 (bf16 numerics): hi = bf16(v), lo unused
VX_DEV void c3_split4p(int prec, const float4 v, float4& hi, float4& lo) {
  if (prec) { hi = make_float4(bf16_round(v.x), bf16_round(v.y), bf16_round(v.z), bf16_round(v.w)); lo = make_float4(0.f, 0.f, 0.f, 0.f); }
  else c3_split4(v, hi, lo);
}

// =====================================================================================================================
// weight images (one launch per forward / backward: the weights change every optimiser step)
// =====================================================================================================================
// forward image  [co tile][group g = dz*3+dy][hi|lo][step s = dx*2 + ci octet][NT*8]: element (n, k) of a step at
//   (n/8)*64 + (k/4)*32 + (n%8)*4 + k%4  (K-major B operand: LBO 128 B between the k halves, SBO 256 B between 8-row groups)
__global__ void conv3_prep_fwd_kernel(const float* __restrict__ w, float* __restrict__ img, int Cout, int NT, int ntile, int prec) {
  VX_PDL_ENTRY();
  const int per_tile = 9 * 6 * NT * 8;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntile * per_tile) return;
  const int t = i / per_tile, r = i % per_tile;
  const int g = r / (6 * NT * 8), r2 = r % (6 * NT * 8), s = r2 / (NT * 8), e = r2 % (NT * 8), n = e >> 3, k = e & 7;
  const int co = t * NT + n, ci = (s & 1) * 8 + k, tap = g * 3 + (s >> 1);
  float hi = 0.f, lo = 0.f;
  if (co < Cout) {
    const float wv = __ldg(w + ((size_t)co * 16 + ci) * 27 + tap);
    if (prec) hi = bf16_round(wv); else tc::split(wv, hi, lo);
  }
  const size_t base = ((size_t)(t * 9 + g) * 2) * (6 * NT * 8);
  const int o = s * NT * 8 + (n >> 3) * 64 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3);
  img[base + o] = hi;
  img[base + 6 * NT * 8 + o] = lo;
}
// data-gradient image  [pass c = co octet][tz][hi|lo][144 x 8]: row n = (ty*3+tx)*16 + ci, k = co - 8c
__global__ void conv3_prep_dgrad_kernel(const float* __restrict__ w, float* __restrict__ img, int Cout, int prec) {
  VX_PDL_ENTRY();
  const int npass = Cout / 8;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npass * 3 * 144 * 8) return;
  const int c = i / (3 * 144 * 8), r = i % (3 * 144 * 8), tz = r / (144 * 8), e = r % (144 * 8), n = e >> 3, k = e & 7;
  const int tyx = n >> 4, ci = n & 15, co = c * 8 + k;
  float hi, lo = 0.f;
  const float wv = __ldg(w + ((size_t)co * 16 + ci) * 27 + tz * 9 + tyx);
  if (prec) hi = bf16_round(wv); else tc::split(wv, hi, lo);
  const size_t base = ((size_t)(c * 3 + tz) * 2) * (144 * 8);
  const int o = (n >> 3) * 64 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3);
  img[base + o] = hi;
  img[base + 144 * 8 + o] = lo;
}

// =====================================================================================================================
// forward
// =====================================================================================================================
struct Conv3FwdArgs {
  const float* x; const float* wimg; const float* bias; float* y;
  int B, Cout, D, H, W, shuffle;
  int NT, ZR, TY, ntz, nty, nblk, tmem_cols, trace, prec;
};

__global__ void __launch_bounds__(C3_THREADS) conv3_fwd_tc_kernel(const __grid_constant__ Conv3FwdArgs A) {
  VX_PDL_ENTRY();
  const int tile = blockIdx.x, nt_i = blockIdx.y, b = blockIdx.z;
  const int ty_i = tile % A.nty, tz_i = tile / A.nty;
  const int z0 = tz_i * A.ZR, y0 = ty_i * A.TY;
  const int D = A.D, H = A.H, W = A.W, NT = A.NT;
  const int PX = W + 2, PY = A.TY + 2, PZ = A.ZR + 2;
  const int NPOS = PZ * PY * PX, NALL = C3_G0 + NPOS + C3_G1;
  const int p_first = (PY + 1) * PX;                  // first output row of the first output plane, column 0
  const int nblk = A.nblk;
  const size_t S = (size_t)D * H * W;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int GF = 2 * 6 * NT * 8;                      // floats of one weight group (hi | lo)

  VX_DYN_SMEM(float, sm);
  float* Xhi = sm;                                    // [4 chunks][NALL][4]
  float* Xlo = Xhi + (size_t)16 * NALL;
  float* Wb = Xlo + (size_t)16 * NALL;                // [2 buffers][GF]
  VX_TC_SHARED_BARS(bars, 4 + C3_MAXBLK);             // full[2], empty[2], done[nblk]
  VX_TC_SHARED_SLOT(tmem_slot);
  uint64_t* full = bars; uint64_t* empty = bars + 2; uint64_t* done = bars + 4;

  C3_STAMP(A.trace, 0);
  if (warp == 0) tc::tmem_alloc(&tmem_slot, (uint32_t)A.tmem_cols);
  if (tid == 0) {
    for (int i = 0; i < 4 + C3_MAXBLK; ++i) tc::mbar_init(&bars[i], 1);
    tc::mbar_init_fence();
  }
  __syncthreads();
  const float* wimg = A.wimg + (size_t)nt_i * 9 * GF;
  if (tid == 32) {                                    // the first two weight groups travel while the brick is staged
    for (int g = 0; g < 2; ++g) {
      tc::mbar_expect_tx(&full[g], (uint32_t)GF * 4u);
      tc::bulk_g2s(Wb + (size_t)g * GF, wimg + (size_t)g * GF, (uint32_t)GF * 2u, &full[g]);
      tc::bulk_g2s(Wb + (size_t)g * GF + GF / 2, wimg + (size_t)g * GF + GF / 2, (uint32_t)GF * 2u, &full[g]);
    }
  }
  // ---- the brick: zero guards, [position][4 channels] per chunk, hi / lo, zero outside the volume
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = tid; i < 4 * (C3_G0 + C3_G1); i += C3_THREADS) {
    const int ch = i / (C3_G0 + C3_G1), r = i % (C3_G0 + C3_G1);
    const int pos = ch * NALL + (r < C3_G0 ? r : NPOS + r);
    reinterpret_cast<float4*>(Xhi)[pos] = zero4;
    reinterpret_cast<float4*>(Xlo)[pos] = zero4;
  }
  C3_STAMP(A.trace, 1);
  const float* xg = A.x + (size_t)b * 16 * S;
  constexpr int FU = 4;                                // items in flight per thread: 16 independent loads before the first use
#pragma unroll 1
  for (int base = tid; base < 4 * NPOS; base += FU * C3_THREADS) {
    float4 v[FU];
#pragma unroll
    for (int u = 0; u < FU; ++u) {
      const int it = base + u * C3_THREADS;
      v[u] = zero4;
      if (it < 4 * NPOS) {
        const int ch = it / NPOS, idx = it % NPOS;
        const int px = idx % PX, py = (idx / PX) % PY, pz = idx / (PX * PY);
        const int gz = z0 + pz - 1, gy = y0 + py - 1, gx = px - 1;
        if (gz >= 0 && gz < D && gy >= 0 && gy < H && gx >= 0 && gx < W) {
          const float* xc = xg + (size_t)(ch * 4) * S + ((size_t)gz * H + gy) * W + gx;
          v[u] = make_float4(__ldg(xc), __ldg(xc + S), __ldg(xc + 2 * S), __ldg(xc + 3 * S));
        }
      }
    }
#pragma unroll
    for (int u = 0; u < FU; ++u) {
      const int it = base + u * C3_THREADS;
      if (it < 4 * NPOS) {
        const int ch = it / NPOS, idx = it % NPOS;
        float4 hi, lo;
        c3_split4p(A.prec, v[u], hi, lo);
        reinterpret_cast<float4*>(Xhi)[ch * NALL + C3_G0 + idx] = hi;
        reinterpret_cast<float4*>(Xlo)[ch * NALL + C3_G0 + idx] = lo;
      }
    }
  }
  tc::fence_async_smem();
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem = tmem_slot;
  C3_STAMP(A.trace, 2);

  if (tid == 0) {
    // ---- MMA issue: 9 weight groups x nblk M-blocks x 6 k-steps x 3 terms
    const uint32_t idesc = tc::idesc_tf32(NT, 0, 0);
    const uint32_t xh = tc::smem_addr(Xhi), xl = tc::smem_addr(Xlo), wb = tc::smem_addr(Wb);
    const uint32_t lbo = (uint32_t)NALL * 16u;
#pragma unroll 1
    for (int g = 0; g < 9; ++g) {
      const int buf = g & 1;
      tc::mbar_wait(&full[buf], (uint32_t)((g >> 1) & 1));
      tc::fence_after();
      C3_STAMP(A.trace, 8 + 2 * g);
      const int dzs = g / 3 - 1, dys = g % 3 - 1;
      const uint32_t wg = wb + (uint32_t)buf * (uint32_t)GF * 4u;
#pragma unroll 1
      for (int blk = 0; blk < nblk; ++blk) {
        const int p0 = C3_G0 + p_first + blk * 128;
#pragma unroll 1
        for (int s = 0; s < 6; ++s) {
          const int shift = (dzs * PY + dys) * PX + (s >> 1) - 1;
          const uint32_t ao = (uint32_t)((s & 1) * 2 * NALL + p0 + shift) * 16u;
          const uint64_t ah = tc::desc(xh + ao, lbo, 128u), al = tc::desc(xl + ao, lbo, 128u);
          const uint64_t bh = tc::desc(wg + (uint32_t)s * (uint32_t)NT * 32u, 128u, 256u);
          const uint64_t bl = tc::desc(wg + (uint32_t)(6 + s) * (uint32_t)NT * 32u, 128u, 256u);
          const uint32_t d = tmem + (uint32_t)(blk * NT);
          if (A.prec) {                                // bf16 numerics: one product
            tc::mma_tf32(d, ah, bh, idesc, (g > 0 || s > 0) ? 1u : 0u);
          } else {
            tc::mma_tf32(d, al, bh, idesc, (g > 0 || s > 0) ? 1u : 0u);
            tc::mma_tf32(d, ah, bl, idesc, 1u);
            tc::mma_tf32(d, ah, bh, idesc, 1u);
          }
        }
        if (g == 8) tc::commit(&done[blk]);
      }
      if (g < 7) tc::commit(&empty[buf]);
      C3_STAMP(A.trace, 9 + 2 * g);
    }
  } else if (tid == 32) {
    // ---- weight producer: group g reuses the buffer of group g-2 once that group's MMAs have completed
#pragma unroll 1
    for (int g = 2; g < 9; ++g) {
      const int buf = g & 1;
      tc::mbar_wait(&empty[buf], (uint32_t)(((g >> 1) - 1) & 1));
      tc::mbar_expect_tx(&full[buf], (uint32_t)GF * 4u);
      tc::bulk_g2s(Wb + (size_t)buf * GF, wimg + (size_t)g * GF, (uint32_t)GF * 2u, &full[buf]);
      tc::bulk_g2s(Wb + (size_t)buf * GF + GF / 2, wimg + (size_t)g * GF + GF / 2, (uint32_t)GF * 2u, &full[buf]);
    }
  }
  __syncwarp();
  C3_STAMP(A.trace, 3);

  // ---- drain: warp = (TMEM lane quadrant, half of the channels); thread = one position of the block
  const int wq = warp & 3, wp = warp >> 2;
  const int chalf = NT / 2;                            // NT is a multiple of 32
#pragma unroll 1
  for (int blk = 0; blk < nblk; ++blk) {
    tc::mbar_wait(&done[blk], 0u);
    tc::fence_after();
    C3_STAMP(A.trace, 4 + (blk > 0));
    const int p = p_first + blk * 128 + wq * 32 + lane;
    const int px = p % PX, py = (p / PX) % PY, pz = p / (PX * PY);
    const int gz = z0 + pz - 1, gy = y0 + py - 1, gx = px - 1;
    const bool ok = px >= 1 && px < 1 + W && py >= 1 && py < 1 + A.TY && pz >= 1 && pz < 1 + A.ZR && gz < D && gy < H;
    const size_t o = (size_t)b * A.Cout * S + ((size_t)gz * H + gy) * W + gx;
#pragma unroll 1
    for (int c0 = wp * chalf; c0 < (wp + 1) * chalf; c0 += 16) {
      float r[16];
      tc::tmem_ld16(tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(blk * NT + c0), r);
      const int co0 = nt_i * NT + c0;
      if (ok && co0 < A.Cout) {
        if (A.bias) {
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] += __ldg(A.bias + co0 + j);
        }
        if (A.prec) {
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = bf16_round(r[j]);
        }
        if (A.shuffle == 4) {
          // PixelShuffle(4): channel ((cls*4 + i)*4 + j)*4 + k -> voxel (4gz + i, 4gy + j, 4gx + k) of class cls.  The 16
          // channels of a chunk share (cls, i); each j is four consecutive output voxels = one 16-byte store.
          const int cls = co0 >> 6, i = (co0 & 63) >> 4;
          const size_t W4 = (size_t)4 * W, H4 = (size_t)4 * H;
          float* orow = A.y + ((((size_t)b * (A.Cout >> 6) + cls) * (4 * D) + (4 * gz + i)) * H4 + 4 * gy) * W4 + 4 * gx;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(orow + j * W4) = make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) A.y[o + (size_t)(co0 + j) * S] = r[j];
        }
      }
    }
  }
  tc::fence_before();
  __syncthreads();
  C3_STAMP(A.trace, 6);
  if (warp == 0) tc::tmem_dealloc(tmem, (uint32_t)A.tmem_cols);
}

static size_t c3_fwd_smem(int NT, int ZR, int TY, int W) {
  const size_t nall = C3_G0 + (size_t)(ZR + 2) * (TY + 2) * (W + 2) + C3_G1;
  return sizeof(float) * (2 * 16 * nall + 2 * (size_t)2 * 6 * NT * 8);
}

static int c3_fwd_geo(const vx_conv_desc* d, Conv3FwdArgs& A) {
  A.NT = d->C_out <= 128 ? (d->C_out + 31) / 32 * 32 : 128;
  const int ntile = cdiv(d->C_out, A.NT), max_blk = 512 / A.NT < C3_MAXBLK ? 512 / A.NT : C3_MAXBLK;
  double best = -1.0;
  for (int TY = 1; TY <= d->H; ++TY)
    for (int ZR = 1; ZR <= 4 && ZR <= d->D; ++ZR) {
      const int PX = d->W + 2, PY = TY + 2;
      const long long nall = C3_G0 + (long long)(ZR + 2) * PY * PX + C3_G1;
      const int nblk = cdiv((long long)((ZR - 1) * PY + TY) * PX, 128);
      if (nblk > max_blk || nall > 16383 || c3_fwd_smem(A.NT, ZR, TY, d->W) > 227 * 1024) continue;
      const int ntz = cdiv(d->D, ZR), nty = cdiv(d->H, TY);
      const long long ncta = (long long)ntz * nty * ntile * d->B;
      // cycles: 162 MMAs of max(64, NT/2) + 3 per M-block, ~10 per staged brick position and chunk
      const double cost = (double)cdiv(ncta, kSMs) * ((double)nblk * 162.0 * 67.0 + 4.0 * (ZR + 2) * PY * PX * 10.0 + 4000.0);
      if (best < 0.0 || cost < best) { best = cost; A.ZR = ZR; A.TY = TY; A.ntz = ntz; A.nty = nty; A.nblk = nblk; }
    }
  if (best < 0.0) return 0;
  A.tmem_cols = 32;
  while (A.tmem_cols < A.nblk * A.NT) A.tmem_cols <<= 1;
  return ntile;
}

// =====================================================================================================================
// weight gradient
// =====================================================================================================================
constexpr int C3W_N = 448;           // 27 taps x 16 ci + 16 bias columns (only the first is used)
constexpr int C3W_STAGES = 3;
constexpr int C3W_RY = 8;            // rows of one unit

struct Conv3WgradArgs {
  const float* dy; const float* x; float* part;       // part [grid][Cout][448]
  int B, Cout, Ctot, co0, D, H, W, shuffle, prec;     // this launch: output channels co0 .. co0 + Cout (<= 128) of Ctot
  int nyg, units;                                     // y groups per plane, units = B * D * nyg
};

// raw x brick: [16 ci][3 planes][RY+2 rows][W+2], the per-channel pitch padded to an odd number of words so that the 16
// channels of an im2col row group land in 16 different banks
VX_DEV int c3w_cpitch(int W) { return (3 * (C3W_RY + 2) * (W + 2)) | 1; }

__global__ void __launch_bounds__(C3_THREADS) conv3_wgrad_tc_kernel(const __grid_constant__ Conv3WgradArgs A) {
  VX_PDL_ENTRY();
  const int D = A.D, H = A.H, W = A.W, Cout = A.Cout;
  const int PXW = W + 2, RYP = C3W_RY + 2;
  const int CP = c3w_cpitch(W);
  const size_t S = (size_t)D * H * W;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nks_row = (W + 7) / 8;

  VX_DYN_SMEM(float, sm);
  float* xs = sm;
  const int XS = 16 * CP;
  float* stg = xs + ((XS + 3) & ~3);                  // stages: A hi [2][128][4] | A lo | B hi [2][448][4] | B lo
  constexpr int SA = 2 * 128 * 4, SB = 2 * C3W_N * 4, STG = 2 * SA + 2 * SB;
  VX_TC_SHARED_BARS(bars, C3W_STAGES + 1);            // empty[stage], done
  VX_TC_SHARED_SLOT(tmem_slot);

  if (warp == 0) tc::tmem_alloc(&tmem_slot, 512u);
  if (tid == 0) {
    for (int i = 0; i <= C3W_STAGES; ++i) tc::mbar_init(&bars[i], 1);
    tc::mbar_init_fence();
  }
  // constant parts of the stages: A rows >= Cout are zero, the bias columns of B are (hi 1, lo 0)
  for (int i = tid; i < C3W_STAGES * STG; i += C3_THREADS) {
    const int r = i % STG;
    float v = 0.f;
    if (r >= 2 * SA && r < 2 * SA + SB) {             // B hi
      const int n = ((r - 2 * SA) >> 2) % C3W_N;
      if (n >= 432) v = 1.f;
    }
    stg[i] = v;
  }
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = tc::idesc_tf32(C3W_N / 2, 0, 0);

  // A item of this thread: one 16-byte load per k-step.  shuffle: (q = (cls,i,j), h, p) -> 4 channels of position x0+4h+p;
  // plain: (co, h) -> 4 positions of channel co.  Loaded one k-step ahead.
  const bool a_live = A.shuffle == 4 ? tid < (Cout >> 2) * 8 : tid < Cout * 2;
  auto load_a = [&](int b, int z, int gy, int x0) -> float4 {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!a_live) return v;
    if (A.shuffle == 4) {
      const int p = tid & 3, h = (tid >> 2) & 1, q = tid >> 3;
      const int cls = q >> 4, ii = (q >> 2) & 3, jj = q & 3;
      const int xx = x0 + 4 * h + p;
      if (xx < W)
        v = c3_ld4(A.dy + ((((size_t)b * (A.Ctot >> 6) + (A.co0 >> 6) + cls) * (4 * D) + (4 * z + ii)) * (4 * H) + 4 * gy + jj) * (size_t)(4 * W) + 4 * xx);
    } else {
      const int h = tid / Cout, co = tid % Cout;
      const int xb = x0 + 4 * h;
      const float* src = A.dy + ((size_t)b * A.Ctot + A.co0 + co) * S + ((size_t)z * H + gy) * W + xb;
      v.x = xb < W ? __ldg(src) : 0.f; v.y = xb + 1 < W ? __ldg(src + 1) : 0.f;
      v.z = xb + 2 < W ? __ldg(src + 2) : 0.f; v.w = xb + 3 < W ? __ldg(src + 3) : 0.f;
    }
    return v;
  };

  int t = 0;                                          // k-steps issued by this CTA
#pragma unroll 1
  for (int unit = blockIdx.x; unit < A.units; unit += gridDim.x) {
    const int yg = unit % A.nyg, z = (unit / A.nyg) % D, b = unit / (A.nyg * D);
    const int y0 = yg * C3W_RY;
    const int ny = min(C3W_RY, H - y0);
    float4 av = load_a(b, z, y0, 0);                  // first k-step of the unit
    __syncthreads();                                  // every thread has finished reading the previous unit's brick
    // ---- raw brick of the unit, zero outside the volume: items = (ci, plane, row), a row's W floats in batches
    {
      const int nrow = 16 * 3 * RYP;
      constexpr int XU = 4;
#pragma unroll 1
      for (int rb = tid; rb < nrow * PXW; rb += XU * C3_THREADS) {
        float v[XU];
#pragma unroll
        for (int u = 0; u < XU; ++u) {
          const int i = rb + u * C3_THREADS;
          v[u] = 0.f;
          if (i < nrow * PXW) {
            const int px = i % PXW, r = i / PXW, py = r % RYP, pz = (r / RYP) % 3, ci = r / (RYP * 3);
            const int gz = z + pz - 1, gy = y0 + py - 1, gx = px - 1;
            if (gz >= 0 && gz < D && gy >= 0 && gy < H && gx >= 0 && gx < W) v[u] = __ldg(A.x + ((size_t)b * 16 + ci) * S + ((size_t)gz * H + gy) * W + gx);
          }
        }
#pragma unroll
        for (int u = 0; u < XU; ++u) {
          const int i = rb + u * C3_THREADS;
          if (i < nrow * PXW) {
            const int px = i % PXW, r = i / PXW, ci = r / (RYP * 3), rr = r % (RYP * 3);
            xs[ci * CP + rr * PXW + px] = v[u];
          }
        }
      }
    }
    __syncthreads();
    const int nks = ny * nks_row;
#pragma unroll 1
    for (int ks = 0; ks < nks; ++ks, ++t) {
      const int yl = ks / nks_row, x0 = (ks % nks_row) * 8;
      const int buf = t % C3W_STAGES;
      if (t >= C3W_STAGES) {                          // the MMAs that read this stage have completed
        tc::mbar_wait(&bars[buf], (uint32_t)((t / C3W_STAGES - 1) & 1));
        tc::fence_after();
      }
      float* Ahi = stg + (size_t)buf * STG; float* Alo = Ahi + SA; float* Bhi = Alo + SA; float* Blo = Bhi + SB;
      // ---- A: dz[co][8 positions] (K-major rows of 4 positions), from the value loaded one step ahead
      if (a_live) {
        if (A.shuffle == 4) {
          const int p = tid & 3, h = (tid >> 2) & 1, q = tid >> 3;
          float4 hi, lo;
          c3_split4p(A.prec, av, hi, lo);
          float* ah = Ahi + (h * 128 + 4 * q) * 4 + p; float* al = Alo + (h * 128 + 4 * q) * 4 + p;
          ah[0] = hi.x; ah[4] = hi.y; ah[8] = hi.z; ah[12] = hi.w;
          al[0] = lo.x; al[4] = lo.y; al[8] = lo.z; al[12] = lo.w;
        } else {
          const int h = tid / Cout, co = tid % Cout;
          float4 hi, lo;
          c3_split4p(A.prec, av, hi, lo);
          reinterpret_cast<float4*>(Ahi)[h * 128 + co] = hi; reinterpret_cast<float4*>(Alo)[h * 128 + co] = lo;
        }
      }
      if (ks + 1 < nks) av = load_a(b, z, y0 + (ks + 1) / nks_row, ((ks + 1) % nks_row) * 8);
      // ---- B: im2col rows n = tap*16 + ci of the same 8 positions, from the raw brick
      for (int it = tid; it < 2 * 432; it += C3_THREADS) {
        const int h = it / 432, n = it % 432;
        const int tap = n >> 4, ci = n & 15;
        const int tz = tap / 9, tyy = (tap / 3) % 3, tx = tap % 3;
        const float* s0 = xs + ci * CP + (tz * RYP + (yl + tyy)) * PXW + x0 + 4 * h + tx;
        const int xb = x0 + 4 * h;                       // ragged last k-step of a row: positions >= W are zero on both operands
        float4 v;
        v.x = xb < W ? s0[0] : 0.f; v.y = xb + 1 < W ? s0[1] : 0.f; v.z = xb + 2 < W ? s0[2] : 0.f; v.w = xb + 3 < W ? s0[3] : 0.f;
        float4 hi, lo;
        c3_split4p(A.prec, v, hi, lo);
        reinterpret_cast<float4*>(Bhi)[h * C3W_N + n] = hi; reinterpret_cast<float4*>(Blo)[h * C3W_N + n] = lo;
      }
      tc::fence_async_smem();
      tc::fence_before();
      __syncthreads();
      tc::fence_after();
      if (tid == 0) {
        const uint32_t ah = tc::smem_addr(Ahi), al = tc::smem_addr(Alo), bh = tc::smem_addr(Bhi), bl = tc::smem_addr(Blo);
        const uint64_t dah = tc::desc(ah, 128u * 16u, 128u), dal = tc::desc(al, 128u * 16u, 128u);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t bo = (uint32_t)half * (C3W_N / 2) * 16u;
          const uint64_t dbh = tc::desc(bh + bo, (uint32_t)C3W_N * 16u, 128u), dbl = tc::desc(bl + bo, (uint32_t)C3W_N * 16u, 128u);
          const uint32_t d = tmem + (uint32_t)half * (C3W_N / 2);
          if (A.prec) {
            tc::mma_tf32(d, dah, dbh, idesc, t > 0 ? 1u : 0u);
          } else {
            tc::mma_tf32(d, dal, dbh, idesc, t > 0 ? 1u : 0u);
            tc::mma_tf32(d, dah, dbl, idesc, 1u);
            tc::mma_tf32(d, dah, dbh, idesc, 1u);
          }
        }
        tc::commit(&bars[buf]);
      }
    }
  }
  // ---- epilogue: the CTA's partial tile -> workspace
  if (tid == 0) tc::commit(&bars[C3W_STAGES]);
  __syncwarp();
  float* pout = A.part + (size_t)blockIdx.x * Cout * C3W_N;
  if (t > 0) {
    tc::mbar_wait(&bars[C3W_STAGES], 0u);
    tc::fence_after();
  }
  const int wq = warp & 3, wp = warp >> 2;
  const int row = wq * 32 + lane;
#pragma unroll 1
  for (int c0 = wp * (C3W_N / 2); c0 < (wp + 1) * (C3W_N / 2); c0 += 16) {
    float r[16];
    if (t > 0) {
      tc::tmem_ld16(tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)c0, r);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = 0.f;
    }
    if (row < Cout) {
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(pout + (size_t)row * C3W_N + c0 + j) = make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]);
    }
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512u);
}

// dW[co][ci][tap] = sum over partial tiles, in tile order (deterministic); db[co] from column 432
__global__ void conv3_wgrad_reduce_kernel(const float* __restrict__ part, int nparts, int Cout, float* __restrict__ dw, float* __restrict__ db) {
  VX_PDL_ENTRY();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * 433) return;
  const int co = i / 433, n = i % 433;
  const float* p = part + (size_t)co * C3W_N + n;
  const size_t st = (size_t)Cout * C3W_N;
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
  int k = 0;
  for (; k + 7 < nparts; k += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += p[(size_t)(k + j) * st];
  }
  for (; k < nparts; ++k) a[0] += p[(size_t)k * st];
  const float s = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  if (n < 432) dw[((size_t)co * 16 + (n & 15)) * 27 + (n >> 4)] = s;
  else if (db) db[co] = s;
}

static size_t c3_wgrad_smem(int W) {
  const int XS = 16 * ((3 * (C3W_RY + 2) * (W + 2)) | 1);
  return sizeof(float) * (((XS + 3) & ~3) + (size_t)C3W_STAGES * (2 * 2 * 128 * 4 + 2 * 2 * C3W_N * 4)) + 16;
}

// =====================================================================================================================
// data gradient
// =====================================================================================================================
struct Conv3DgradArgs {
  const float* dy; const float* wimg; float* dx;
  int B, Cout, D, H, W, shuffle, prec;
  int TY, nty, nblk;
};
constexpr int C3D_N = 144;
constexpr int C3D_WF = 3 * 2 * C3D_N * 8;             // floats of one pass of the weight image: [tz][hi|lo][144 x 8]

__global__ void __launch_bounds__(C3_THREADS) conv3_dgrad_tc_kernel(const __grid_constant__ Conv3DgradArgs A) {
  VX_PDL_ENTRY();
  const int ty_i = blockIdx.x % A.nty, z = blockIdx.x / A.nty, b = blockIdx.y;
  const int y0 = ty_i * A.TY;
  const int D = A.D, H = A.H, W = A.W, Cout = A.Cout;
  const int PX = W + 2, PY = A.TY + 2;
  const int P2 = PY * PX, NPOS = 3 * P2, NALL = C3_G0 + NPOS + C3_G1;
  const int nblk = A.nblk;                            // M-blocks covering the P2 positions of the middle plane
  const size_t S = (size_t)D * H * W;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int npass = Cout / 8;

  VX_DYN_SMEM(float, sm);
  const size_t BR = (size_t)2 * 2 * NALL * 4;         // floats of one brick buffer: hi [2 chunks][NALL][4] | lo
  float* brick = sm;                                  // [2 buffers][BR]
  float* Wb = brick + 2 * BR;                         // [2 buffers][C3D_WF]
  VX_TC_SHARED_BARS(bars, 5);                         // full[2], empty[2], done
  VX_TC_SHARED_SLOT(tmem_slot);
  uint64_t* full = bars; uint64_t* empty = bars + 2; uint64_t* done = bars + 4;

  if (warp == 0) tc::tmem_alloc(&tmem_slot, 512u);
  if (tid == 0) {
    for (int i = 0; i < 5; ++i) tc::mbar_init(&bars[i], 1);
    tc::mbar_init_fence();
  }
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = tid; i < 2 * 2 * 2 * (C3_G0 + C3_G1); i += C3_THREADS) {      // guards of both buffers, hi and lo, both chunks
    const int a = i / (C3_G0 + C3_G1), r = i % (C3_G0 + C3_G1);
    reinterpret_cast<float4*>(brick)[(size_t)a * NALL + (r < C3_G0 ? r : NPOS + r)] = zero4;
  }
  __syncthreads();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = tc::idesc_tf32(C3D_N, 0, 0);

  // The brick of pass c (dz channels 8c .. 8c+7 at the 3 x PY x PX padded positions, [chunk][position][4 channels]) is loaded
  // into registers one pass AHEAD: the loads of pass c+1 are in flight while pass c is fenced, synchronised and issued.
  constexpr int DU = 8;                               // items per thread (2 * NPOS <= 8 * 256, checked by the launcher)
  float4 v[DU];
  auto load_pass = [&](int c) {
#pragma unroll
    for (int u = 0; u < DU; ++u) {
      const int it = tid + u * C3_THREADS;
      v[u] = zero4;
      if (it < 2 * NPOS) {
        const int ch = it / NPOS, idx = it % NPOS;
        const int px = idx % PX, py = (idx / PX) % PY, pz = idx / P2;
        const int gz = z + pz - 1, gy = y0 + py - 1, gx = px - 1;
        if (gz >= 0 && gz < D && gy >= 0 && gy < H && gx >= 0 && gx < W) {
          const int co = 8 * c + 4 * ch;
          if (A.shuffle == 4) {
            const int cls = co >> 6, ii = (co >> 4) & 3, jj = (co >> 2) & 3;
            v[u] = c3_ld4(A.dy + ((((size_t)b * (Cout >> 6) + cls) * (4 * D) + (4 * gz + ii)) * (4 * H) + 4 * gy + jj) * (size_t)(4 * W) + 4 * gx);
          } else {
            const float* s0 = A.dy + ((size_t)b * Cout + co) * S + ((size_t)gz * H + gy) * W + gx;
            v[u] = make_float4(__ldg(s0), __ldg(s0 + S), __ldg(s0 + 2 * S), __ldg(s0 + 3 * S));
          }
        }
      }
    }
  };
  load_pass(0);
#pragma unroll 1
  for (int c = 0; c < npass; ++c) {
    const int buf = c & 1;
    if (c >= 2) {                                     // the MMAs of pass c-2 have read this brick buffer and weight buffer
      tc::mbar_wait(&empty[buf], (uint32_t)(((c >> 1) - 1) & 1));
      tc::fence_after();
    }
    if (tid == 32) {
      tc::mbar_expect_tx(&full[buf], (uint32_t)C3D_WF * 4u);
      tc::bulk_g2s(Wb + (size_t)buf * C3D_WF, A.wimg + (size_t)c * C3D_WF, (uint32_t)C3D_WF * 4u, &full[buf]);
    }
    float* Bh = brick + (size_t)buf * BR; float* Bl = Bh + BR / 2;
#pragma unroll
    for (int u = 0; u < DU; ++u) {
      const int it = tid + u * C3_THREADS;
      if (it < 2 * NPOS) {
        const int ch = it / NPOS, idx = it % NPOS;
        float4 hi, lo;
        c3_split4p(A.prec, v[u], hi, lo);
        reinterpret_cast<float4*>(Bh)[ch * NALL + C3_G0 + idx] = hi;
        reinterpret_cast<float4*>(Bl)[ch * NALL + C3_G0 + idx] = lo;
      }
    }
    if (c + 1 < npass) load_pass(c + 1);
    tc::fence_async_smem();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    if (tid == 0) {
      tc::mbar_wait(&full[buf], (uint32_t)((c >> 1) & 1));
      tc::fence_after();
      const uint32_t xh = tc::smem_addr(Bh), xl = tc::smem_addr(Bl), wb = tc::smem_addr(Wb + (size_t)buf * C3D_WF);
      const uint32_t lbo = (uint32_t)NALL * 16u;
#pragma unroll 1
      for (int blk = 0; blk < nblk; ++blk) {
#pragma unroll 1
        for (int tz = 0; tz < 3; ++tz) {
          // D row u' reads dz at u' - (tz-1) planes
          const uint32_t ao = (uint32_t)(C3_G0 + P2 + blk * 128 - (tz - 1) * P2) * 16u;
          const uint64_t ah = tc::desc(xh + ao, lbo, 128u), al = tc::desc(xl + ao, lbo, 128u);
          const uint64_t bh = tc::desc(wb + (uint32_t)(tz * 2) * (C3D_N * 32u), 128u, 256u);
          const uint64_t bl = tc::desc(wb + (uint32_t)(tz * 2 + 1) * (C3D_N * 32u), 128u, 256u);
          const uint32_t d = tmem + (uint32_t)(blk * C3D_N);
          if (A.prec) {
            tc::mma_tf32(d, ah, bh, idesc, (c > 0 || tz > 0) ? 1u : 0u);
          } else {
            tc::mma_tf32(d, al, bh, idesc, (c > 0 || tz > 0) ? 1u : 0u);
            tc::mma_tf32(d, ah, bl, idesc, 1u);
            tc::mma_tf32(d, ah, bh, idesc, 1u);
          }
        }
      }
      tc::commit(c == npass - 1 ? done : &empty[buf]);
    }
  }
  __syncwarp();
  tc::mbar_wait(done, 0u);
  tc::fence_after();
  __syncthreads();                                    // both brick buffers are free: they become the exchange tile

  // ---- in-plane col2im: dx[q][ci] = sum over (ty,tx) of D[q - (ty-1)*PX - (tx-1)][(ty,tx,ci)], one channel quad at a time
  float4* T = reinterpret_cast<float4*>(brick);       // [nblk*128 rows][9 taps]
  const int wq = warp & 3, wp = warp >> 2;
  const int nout = A.TY * W;
#pragma unroll 1
  for (int qd = 0; qd < 4; ++qd) {
    for (int blk = wp; blk < nblk; blk += 2) {
      const int row = blk * 128 + wq * 32 + lane;
      float r[16];
#pragma unroll 1
      for (int tp = 0; tp < 9; ++tp) {
        tc::tmem_ld16(tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(blk * C3D_N + tp * 16), r);
        T[row * 9 + tp] = qd == 0 ? make_float4(r[0], r[1], r[2], r[3]) : qd == 1 ? make_float4(r[4], r[5], r[6], r[7])
                        : qd == 2 ? make_float4(r[8], r[9], r[10], r[11]) : make_float4(r[12], r[13], r[14], r[15]);
      }
    }
    __syncthreads();
    for (int o = tid; o < nout; o += C3_THREADS) {
      const int xx = o % W, yl = o / W;
      const int gy = y0 + yl;
      if (gy >= H) continue;
      const int q = (yl + 1) * PX + xx + 1;           // row inside the middle plane
      float4 acc = zero4;
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) {
        const float4 v = T[(q - (tp / 3 - 1) * PX - (tp % 3 - 1)) * 9 + tp];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      float* dst = A.dx + ((size_t)b * 16 + qd * 4) * S + ((size_t)z * H + gy) * W + xx;
      if (A.prec) acc = make_float4(bf16_round(acc.x), bf16_round(acc.y), bf16_round(acc.z), bf16_round(acc.w));
      dst[0] = acc.x; dst[S] = acc.y; dst[2 * S] = acc.z; dst[3 * S] = acc.w;
    }
    __syncthreads();
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512u);
}

static size_t c3_dgrad_smem(int TY, int W) {
  const size_t nall = C3_G0 + (size_t)3 * (TY + 2) * (W + 2) + C3_G1;
  return sizeof(float) * (2 * (size_t)2 * 2 * nall * 4 + 2 * (size_t)C3D_WF);
}

static int c3_dgrad_geo(const vx_conv_desc* d, Conv3DgradArgs& A) {
  double best = -1.0;
  for (int TY = 1; TY <= d->H; ++TY) {
    const int P2 = (TY + 2) * (d->W + 2);
    const int nblk = cdiv(P2, 128);
    const long long nall = C3_G0 + 3LL * P2 + C3_G1;
    if (nblk * C3D_N > 512 || nall > 16383 || c3_dgrad_smem(TY, d->W) > 227 * 1024 || 2 * 3 * P2 > 8 * C3_THREADS) continue;
    if ((size_t)nblk * 128 * 9 * 16 > sizeof(float) * 2 * (size_t)2 * 2 * nall * 4) continue;      // exchange tile inside the brick buffers
    const int nty = cdiv(d->H, TY);
    const long long ncta = (long long)nty * d->D * d->B;
    const double cost = (double)cdiv(ncta, kSMs) * ((double)(d->C_out / 8) * fmax(nblk * 9 * 75.0, 3.0 * P2 * 2 / 256.0 * 40.0 + 600.0) + 6000.0);
    if (best < 0.0 || cost < best) { best = cost; A.TY = TY; A.nty = nty; A.nblk = nblk; }
  }
  return best >= 0.0;
}

// =====================================================================================================================
// host side
// =====================================================================================================================
static bool c3_supported(const vx_conv_desc* d) {
  return d && d->B > 0 && d->C_in == 16 && d->C_out > 0 && d->C_out % 16 == 0 && d->kernel == 3 && d->stride == 1 && d->pad == 1 &&
         !d->transposed && d->D > 0 && d->H > 0 && d->W > 0 && d->W + 2 <= 126 && (d->shuffle == 0 || (d->shuffle == 4 && d->C_out % 64 == 0));
}
static size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }
static size_t c3_img_fwd_bytes(const vx_conv_desc* d) {
  const int NT = d->C_out <= 128 ? (d->C_out + 31) / 32 * 32 : 128;
  return align256(sizeof(float) * (size_t)cdiv(d->C_out, NT) * 9 * 2 * 6 * NT * 8);
}
static size_t c3_img_dgrad_bytes(const vx_conv_desc* d) { return align256(sizeof(float) * (size_t)(d->C_out / 8) * C3D_WF); }
static int c3_wgrad_grid(const vx_conv_desc* d) {
  const int units = d->B * d->D * cdiv(d->H, C3W_RY);
  return units < kSMs ? units : kSMs;
}
static size_t c3_part_bytes(const vx_conv_desc* d) {
  return align256(sizeof(float) * (size_t)c3_wgrad_grid(d) * (d->C_out < 128 ? d->C_out : 128) * C3W_N);
}

bool conv3_tc_supported(const vx_conv_desc* d) { return c3_supported(d) && d->C_out % 8 == 0 && (d->C_out <= 128 || d->C_out % 128 == 0); }

size_t conv3_tc_workspace(const vx_conv_desc* d) {
  if (!conv3_tc_supported(d)) return 0;
  const size_t f = c3_img_fwd_bytes(d), bwd = c3_img_dgrad_bytes(d) + c3_part_bytes(d);
  return f > bwd ? f : bwd;
}

int conv3_tc_fwd(const vx_conv_desc* d, const float* x, const float* w, const float* bias, float* y, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!conv3_tc_supported(d)) { set_error("conv3: needs 16 input channels, k3 s1 p1, C_out %% 16 == 0 (<= 128 or a multiple of 128), W <= 124"); return VX_ERR_UNSUPPORTED; }
  if (!ws || ws_bytes < c3_img_fwd_bytes(d)) { set_error("conv3_fwd: workspace too small"); return VX_ERR_WORKSPACE; }
  if (((uintptr_t)y & 15) || ((uintptr_t)ws & 15)) { set_error("conv3_fwd: output / workspace not 16-byte aligned"); return VX_ERR_BAD_DESC; }
  Conv3FwdArgs A{};
  A.x = x; A.wimg = (const float*)ws; A.bias = bias; A.y = y; A.B = d->B; A.Cout = d->C_out; A.D = d->D; A.H = d->H; A.W = d->W; A.shuffle = d->shuffle;
#ifndef VX_EMU
  A.trace = g_c3_trace_on;
#endif
  A.prec = precision_mode();
  const int ntile = c3_fwd_geo(d, A);
  if (!ntile) { set_error("conv3_fwd: no brick fits"); return VX_ERR_UNSUPPORTED; }
  prof_scope("conv3_fwd B%d Co%d %dx%dx%d", d->B, d->C_out, d->D, d->H, d->W);
  const int nimg = ntile * 9 * 6 * A.NT * 8;
  VX_LAUNCH(conv3_prep_fwd_kernel, dim3(cdiv(nimg, 256)), dim3(256), 0, st, w, (float*)ws, d->C_out, A.NT, ntile, A.prec);
  const size_t smem = c3_fwd_smem(A.NT, A.ZR, A.TY, A.W);
  prof_bytes(4.0 * d->B * (16.0 + d->C_out) * d->D * d->H * d->W + 4.0 * 27 * 16 * d->C_out);
  prof_flops(2.0 * 27 * 16 * (double)d->C_out * d->B * d->D * d->H * d->W);
  VX_SET_SMEM(conv3_fwd_tc_kernel, smem);
  VX_LAUNCH(conv3_fwd_tc_kernel, dim3(A.ntz * A.nty, ntile, A.B), dim3(C3_THREADS), smem, st, A);
  return check_launch("conv3_fwd_tc_kernel");
}

int conv3_tc_bwd(const vx_conv_desc* d, const float* dy, const float* x, const float* w, float* dx, float* dw, float* db, void* ws,
                 size_t ws_bytes, cudaStream_t st) {
  if (!conv3_tc_supported(d)) { set_error("conv3: unsupported shape"); return VX_ERR_UNSUPPORTED; }
  if (!ws || ws_bytes < conv3_tc_workspace(d) || ((uintptr_t)ws & 15)) { set_error("conv3_bwd: workspace too small or misaligned"); return VX_ERR_WORKSPACE; }
  if ((uintptr_t)dy & 15) { set_error("conv3_bwd: dy not 16-byte aligned"); return VX_ERR_BAD_DESC; }
  prof_scope("conv3_bwd B%d Co%d %dx%dx%d", d->B, d->C_out, d->D, d->H, d->W);
  char* wsb = (char*)ws;
  const double act_bytes = 4.0 * d->B * (16.0 + d->C_out) * d->D * d->H * d->W;
  if (dw) {
    // weight gradient on the side stream (a leaf of the backward graph), M-tiles of 128 output channels
    SideJoin join(st);
    cudaStream_t sw = side_fork(st);
    const int ntile = cdiv(d->C_out, 128);
    for (int t = 0; t < ntile; ++t) {                 // M-tiles of 128 output channels (with the shuffle: 2 whole classes each)
      Conv3WgradArgs A{};
      const int co0 = t * 128, cn = d->C_out - co0 < 128 ? d->C_out - co0 : 128;
      A.dy = dy; A.x = x; A.part = (float*)(wsb + c3_img_dgrad_bytes(d));       // the tiles run back to back on one stream: one partial buffer
      A.B = d->B; A.Cout = cn; A.Ctot = d->C_out; A.co0 = co0; A.D = d->D; A.H = d->H; A.W = d->W; A.shuffle = d->shuffle;
      A.prec = precision_mode();
      A.nyg = cdiv(d->H, C3W_RY); A.units = d->B * d->D * A.nyg;
      const int grid = c3_wgrad_grid(d);
      const size_t smem = c3_wgrad_smem(d->W);
      if (smem > 227 * 1024) { set_error("conv3_bwd: row too wide for the weight-gradient brick"); return VX_ERR_UNSUPPORTED; }
      prof_bytes(act_bytes);
      prof_flops(2.0 * 27 * 16 * (double)cn * d->B * d->D * d->H * d->W);
      VX_SET_SMEM(conv3_wgrad_tc_kernel, smem);
      VX_LAUNCH(conv3_wgrad_tc_kernel, dim3(grid), dim3(C3_THREADS), smem, sw, A);
      VX_LAUNCH(conv3_wgrad_reduce_kernel, dim3(cdiv(cn * 433, 256)), dim3(256), 0, sw, A.part, grid, cn, dw + (size_t)co0 * 16 * 27, db ? db + co0 : nullptr);
    }
    const int rc = check_launch("conv3_wgrad_tc_kernel");
    if (rc != VX_OK) return rc;
  }
  if (dx) {
    Conv3DgradArgs A{};
    A.dy = dy; A.wimg = (const float*)wsb; A.dx = dx; A.B = d->B; A.Cout = d->C_out; A.D = d->D; A.H = d->H; A.W = d->W; A.shuffle = d->shuffle;
    A.prec = precision_mode();
    if (!c3_dgrad_geo(d, A)) { set_error("conv3_bwd: no data-gradient brick fits"); return VX_ERR_UNSUPPORTED; }
    const int nimg = (d->C_out / 8) * 3 * 144 * 8;
    VX_LAUNCH(conv3_prep_dgrad_kernel, dim3(cdiv(nimg, 256)), dim3(256), 0, st, w, (float*)wsb, d->C_out, A.prec);
    const size_t smem = c3_dgrad_smem(A.TY, A.W);
    prof_bytes(act_bytes);
    prof_flops(2.0 * 27 * 16 * (double)d->C_out * d->B * d->D * d->H * d->W);
    VX_SET_SMEM(conv3_dgrad_tc_kernel, smem);
    VX_LAUNCH(conv3_dgrad_tc_kernel, dim3(A.nty * d->D, d->B), dim3(C3_THREADS), smem, st, A);
    const int rc = check_launch("conv3_dgrad_tc_kernel");
    if (rc != VX_OK) return rc;
  }
  return VX_OK;
}

}  // namespace vx
