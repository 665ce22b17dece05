// Dense 3x3x3 convolution with 16 input channels (decoder.out_conv1 and the reconstruction out_conv, Decoder.py:73-76,
// 150-153 -- SURVEY.md section 8f row 1: 44-60 % of the forward MACs) as an implicit GEMM on the 5th-generation tensor
// cores: tcgen05.mma kind::tf32, one term (inputs and weights rounded to tf32, fp32 accumulation in tensor memory) -- the
// precision class of the library convolution it replaces when torch.backends.cudnn.allow_tf32 is on.
//
// STATUS: candidate, OFF by default (vx_set_option(VX_OPT_DENSE_CONV_TC, 1) / VX_DENSE_CONV_TC=1 in the Python layer).
// Written after the round's GPU budget was spent; layouts are checked on the CPU shim with a software model of the MMA
// (tests/test_emu_kernels.py::test_dense_conv_tensor_core); it has NOT run on a B200 yet.  Same open hardware questions as
// jlc_tc.cu (operand descriptors with 16-byte-aligned start addresses).  Forward only: the backward stays the library's.
//
//   z[b, co, p] = sum over taps t in 3^3 and ci in 0..15 of  w[co, ci, t] * x[b, ci, p + t - 1]  (+ bias[co])
// optionally stored through PixelShuffle(4) (superpixel.py:15): the (B, C_out, D, H, W) intermediate never reaches memory.
//
// GEMM view per CTA = (batch b, tile of <= 128 output channels, brick of ZR planes x TY rows x full width):
//   D[m = flat padded position][n = co]  with the shifted-descriptor addressing of jlc_tc.cu: the halo brick (halo 1)
// is staged once as four [position][4 channels] arrays of 16-byte rows; tap t is the brick read from
// base + 16 * (p0 + shift(t)); a k-step is one tap x 8 channels (two arrays, LBO = one array apart), 54 k-steps in three
// passes of one dz slab (the 18 x NT x 8 weight tile of a slab is <= 72 KB; all 54 would not fit beside the brick).
// An M-block owns NT TMEM columns (<= 4 blocks per CTA at NT = 128).  Drain: thread = position, the two warps of a TMEM
// lane quadrant split the channels, lanes run along x so every channel row is one contiguous store.
#include "vx_kernels.h"
#include "vx_tc.cuh"

#ifdef VX_EMU
#define __grid_constant__
#endif

namespace vx {

constexpr int DC_THREADS = 256;
constexpr int DC_G0 = 8;            // guard positions before the brick (the first block reads 1 position in front of it)
constexpr int DC_G1 = 136;          // behind it: 1 + the 127 padding rows of the last block, rounded up
constexpr int DC_STEPS = 18;        // k-steps of one dz slab: 9 taps x 2 channel pairs
constexpr int DC_MAX_BLK = 4;

#ifdef VX_EMU
static float g_emu_tmem_dc[128][512];
#endif

// round-to-nearest (ties away) to tf32: the tensor core reads the top 19 bits of an fp32 operand
VX_DEV float dc_tf32(float x) {
#ifdef VX_EMU
  uint32_t u; memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r;
#else
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
#endif
}

__global__ void __launch_bounds__(DC_THREADS) conv_dense_tc_kernel(const __grid_constant__ DenseConvArgs A) {
  const int tile = blockIdx.x, nt_i = blockIdx.y, b = blockIdx.z;
  const int ty_i = tile % A.nty, tz_i = tile / A.nty;
  const int z0 = tz_i * A.ZR, y0 = ty_i * A.TY;
  const int D = A.D, H = A.H, W = A.W, NT = A.NT;
  const int PX = W + 2, PY = A.TY + 2, PZ = A.ZR + 2;
  const int NPOS = PZ * PY * PX;
  const int p_first = (PY + 1) * PX;                  // first output row of the first output plane, column 0
  const int nblk = A.nblk;
  const int BSTEP = NT * 8;                           // floats of one k-step of the weight operand
  const size_t S = (size_t)D * H * W;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  VX_DYN_SMEM(float, sm);
  const int NALL = DC_G0 + NPOS + DC_G1;
  float* X = sm;                                      // [4 chunks][NALL][4]
  float* Bw = X + (size_t)4 * NALL * 4;               // [18][NT / 8][2 k-halves][8 n][4 k]

#ifndef VX_EMU
  __shared__ __align__(8) uint64_t mbar[DC_MAX_BLK + 1];
  __shared__ uint32_t tmem_slot;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"((uint32_t)A.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int i = 0; i <= DC_MAX_BLK; ++i) mbar_init(smem_u32(&mbar[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t tmem = 0;
#endif

  // ---- the brick: zero guards, then [position][4 channels] per chunk, rounded to tf32, zero outside the volume
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = tid; i < 4 * (DC_G0 + DC_G1); i += DC_THREADS) {
    const int ch = i / (DC_G0 + DC_G1), r = i % (DC_G0 + DC_G1);
    reinterpret_cast<float4*>(X)[ch * NALL + (r < DC_G0 ? r : NPOS + r)] = zero4;
  }
  const float* xg = A.x + (size_t)b * 16 * S;
  for (int it = tid; it < 4 * NPOS; it += DC_THREADS) {
    const int ch = it / NPOS, idx = it % NPOS;
    const int px = idx % PX, py = (idx / PX) % PY, pz = idx / (PX * PY);
    const int gz = z0 + pz - 1, gy = y0 + py - 1, gx = px - 1;
    float4 v = zero4;
    if (gz >= 0 && gz < D && gy >= 0 && gy < H && gx >= 0 && gx < W) {
      const float* xc = xg + (size_t)(ch * 4) * S + ((size_t)gz * H + gy) * W + gx;
      v = make_float4(dc_tf32(__ldg(xc)), dc_tf32(__ldg(xc + S)), dc_tf32(__ldg(xc + 2 * S)), dc_tf32(__ldg(xc + 3 * S)));
    }
    reinterpret_cast<float4*>(X)[ch * NALL + DC_G0 + idx] = v;
  }

#pragma unroll 1
  for (int pass = 0; pass < 3; ++pass) {              // pass = dz slab
    if (pass > 0) {                                   // the previous slab's MMAs have read the weights
#ifndef VX_EMU
      mbar_wait(smem_u32(&mbar[DC_MAX_BLK]), (uint32_t)((pass - 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#else
      __syncthreads();
#endif
    }
    // weights of the slab: k-step s = (dy * 3 + dx) * 2 + channel pair; element (n, k) at
    // s*BSTEP + (n/8)*64 + (k/4)*32 + (n%8)*4 + k%4   (LBO 128 B, SBO 256 B)
    for (int e = tid; e < DC_STEPS * BSTEP; e += DC_THREADS) {
      const int s = e / BSTEP, r = e % BSTEP, n = r >> 3, k = r & 7;
      const int tap = pass * 9 + (s >> 1), ci = (s & 1) * 8 + k, co = nt_i * NT + n;
      const float w = co < A.Cout ? dc_tf32(__ldg(A.w + ((size_t)co * 16 + ci) * 27 + tap)) : 0.f;
      Bw[s * BSTEP + (n >> 3) * 64 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3)] = w;
    }
#ifndef VX_EMU
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tmem = tmem_slot;
    if (tid == 0) {
      const uint32_t idesc = umma_idesc_tf32(NT);
      const uint32_t xb = smem_u32(X), bb = smem_u32(Bw);
      const uint32_t lbo = (uint32_t)NALL * 16u;      // the two k-halves of a k-step are consecutive channel arrays
#pragma unroll 1
      for (int blk = 0; blk < nblk; ++blk) {
        const int p0 = DC_G0 + p_first + blk * 128;
#pragma unroll 1
        for (int s = 0; s < DC_STEPS; ++s) {
          const int t9 = s >> 1, pair = s & 1;
          const int shift = ((pass - 1) * PY + (t9 / 3 - 1)) * PX + (t9 % 3 - 1);
          const uint32_t ao = (uint32_t)(pair * 2 * NALL + p0 + shift) * 16u;
          umma_tf32(tmem + (uint32_t)(blk * NT), umma_desc(xb + ao, lbo, 128u),
                    umma_desc(bb + (uint32_t)s * (uint32_t)(BSTEP * 4), 128u, 256u), idesc, (pass > 0 || s > 0) ? 1u : 0u);
        }
        if (pass == 2)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar[blk]))
                       : "memory");
      }
      if (pass < 2)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         smem_u32(&mbar[DC_MAX_BLK]))
                     : "memory");
    }
    __syncwarp();
#else
    __syncthreads();
    if (tid == 0) {      // software model of the MMAs on the same shared layout / descriptor arithmetic
      for (int blk = 0; blk < nblk; ++blk) {
        const int p0 = DC_G0 + p_first + blk * 128;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < NT; ++n) {
            float acc = pass > 0 ? g_emu_tmem_dc[m][blk * NT + n] : 0.f;
            for (int s = 0; s < DC_STEPS; ++s) {
              const int t9 = s >> 1, pair = s & 1;
              const int shift = ((pass - 1) * PY + (t9 / 3 - 1)) * PX + (t9 % 3 - 1);
              for (int k = 0; k < 8; ++k) {
                const int ao = ((pair * 2 + (k >> 2)) * NALL + p0 + shift + m) * 4 + (k & 3);
                acc += X[ao] * Bw[s * BSTEP + (n >> 3) * 64 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3)];
              }
            }
            g_emu_tmem_dc[m][blk * NT + n] = acc;
          }
      }
    }
    __syncthreads();
#endif
  }

  // ---- drain: warp = (TMEM lane quadrant, half of the channels); thread = one position of the block
  const int wq = warp & 3, wp = warp >> 2;
  const int chalf = NT / 2;                            // NT is a multiple of 32
#pragma unroll 1
  for (int blk = 0; blk < nblk; ++blk) {
#ifndef VX_EMU
    mbar_wait(smem_u32(&mbar[blk]), 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#endif
    const int p = p_first + blk * 128 + wq * 32 + lane;
    const int px = p % PX, py = (p / PX) % PY, pz = p / (PX * PY);
    const int gz = z0 + pz - 1, gy = y0 + py - 1, gx = px - 1;
    const bool ok = px >= 1 && px < 1 + W && py >= 1 && py < 1 + A.TY && pz >= 1 && pz < 1 + A.ZR && gz < D && gy < H;
    const size_t o = (size_t)b * A.Cout * S + ((size_t)gz * H + gy) * W + gx;
#pragma unroll 1
    for (int c0 = wp * chalf; c0 < (wp + 1) * chalf; c0 += 16) {
      float r[16];
#ifndef VX_EMU
      uint32_t q[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
            "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
          : "r"(tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(blk * NT + c0))
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = __uint_as_float(q[j]);
#else
      for (int j = 0; j < 16; ++j) r[j] = g_emu_tmem_dc[wq * 32 + lane][blk * NT + c0 + j];
#endif
      const int co0 = nt_i * NT + c0;
      if (ok && co0 < A.Cout) {
        if (A.bias) {
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] += __ldg(A.bias + co0 + j);
        }
        if (A.shuffle == 4) {
          // PixelShuffle(4): channel ((cls * 4 + i) * 4 + j) * 4 + k -> voxel (4 gz + i, 4 gy + j, 4 gx + k) of class cls.
          // The 16 channels of a chunk share (cls, i); each j is four consecutive output voxels = one 16-byte store, and
          // the lanes of a warp (consecutive gx) write one contiguous row segment.
          const int cls = co0 >> 6, i = (co0 & 63) >> 4;
          const size_t W4 = (size_t)4 * W, H4 = (size_t)4 * H;
          float* orow = A.z + ((((size_t)b * (A.Cout >> 6) + cls) * (4 * D) + (4 * gz + i)) * H4 + 4 * gy) * W4 + 4 * gx;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(orow + j * W4) = make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) A.z[o + (size_t)(co0 + j) * S] = r[j];
        }
      }
    }
  }
#ifndef VX_EMU
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)A.tmem_cols) : "memory");
  }
#endif
}

static int g_dc_enabled = 0;
void dense_conv_tc_set(int enabled) { g_dc_enabled = enabled ? 1 : 0; }

static size_t dc_smem_bytes(int NT, int ZR, int TY, int W) {
  const size_t nall = DC_G0 + (size_t)(ZR + 2) * (TY + 2) * (W + 2) + DC_G1;
  return sizeof(float) * (4 * nall * 4 + (size_t)DC_STEPS * NT * 8);
}

int dense_conv_tc_fwd(const vx_dense_conv_desc* d, const float* x, const float* w, const float* bias, float* z, cudaStream_t st) {
  if (!g_dc_enabled) { set_error("dense_conv: the candidate kernel is off (VX_OPT_DENSE_CONV_TC)"); return VX_ERR_UNSUPPORTED; }
  if (!d || d->B <= 0 || d->C_in != 16 || d->C_out <= 0 || d->C_out % 16 || d->D <= 0 || d->H <= 0 || d->W <= 0 || d->W + 2 > 128) {
    set_error("dense_conv: needs 16 input channels, output channels in multiples of 16, width <= 126");
    return VX_ERR_UNSUPPORTED;
  }
  if (d->shuffle != 0 && (d->shuffle != 4 || d->C_out % 64)) {
    set_error("dense_conv: fused pixel shuffle needs scale 4 and output channels in multiples of 64");
    return VX_ERR_UNSUPPORTED;
  }
  if (((uintptr_t)z & 15) != 0) { set_error("dense_conv: output not 16-byte aligned"); return VX_ERR_BAD_DESC; }
  DenseConvArgs A{};
  A.x = x; A.w = w; A.bias = bias; A.shuffle = d->shuffle; A.z = z; A.B = d->B; A.Cout = d->C_out; A.D = d->D; A.H = d->H; A.W = d->W;
  A.NT = d->C_out <= 128 ? (d->C_out + 31) / 32 * 32 : 128;
  const int ntile = cdiv(d->C_out, A.NT), max_blk = 512 / A.NT < DC_MAX_BLK ? 512 / A.NT : DC_MAX_BLK;
  double best = -1.0;
  const int zr_c[] = {1, 2, 3, 4};
  for (int TY = 1; TY <= d->H; ++TY)
    for (int ZR : zr_c) {
      if (ZR > d->D && ZR != 1) continue;
      const int PX = d->W + 2, PY = TY + 2;
      const long long nall = DC_G0 + (long long)(ZR + 2) * PY * PX + DC_G1;
      const int nblk = cdiv((long long)((ZR - 1) * PY + TY) * PX, 128);
      if (nblk > max_blk || nall > 16383 || dc_smem_bytes(A.NT, ZR, TY, d->W) > 227 * 1024) continue;
      const int ntz = cdiv(d->D, ZR), nty = cdiv(d->H, TY);
      const long long ncta = (long long)ntz * nty * ntile * d->B;
      const double cost = (double)cdiv(ncta, kSMs) * ((double)nblk * 54.0 * 0.6 * A.NT + 4.0 * (ZR + 2) * PY * PX * 6.0);
      if (best < 0.0 || cost < best) { best = cost; A.ZR = ZR; A.TY = TY; A.ntz = ntz; A.nty = nty; A.nblk = nblk; }
    }
  if (best < 0.0) { set_error("dense_conv: no brick fits"); return VX_ERR_UNSUPPORTED; }
  A.tmem_cols = 32;
  while (A.tmem_cols < A.nblk * A.NT) A.tmem_cols <<= 1;
  const size_t smem = dc_smem_bytes(A.NT, A.ZR, A.TY, A.W);
  prof_scope("dense_conv_fwd B%d Co%d %dx%dx%d", d->B, d->C_out, d->D, d->H, d->W);
  prof_bytes(4.0 * d->B * (16.0 + d->C_out) * d->D * d->H * d->W + 4.0 * 27 * 16 * d->C_out);
  VX_SET_SMEM(conv_dense_tc_kernel, smem);
  VX_LAUNCH(conv_dense_tc_kernel, dim3(A.ntz * A.nty, ntile, A.B), dim3(DC_THREADS), smem, st, A);
  return check_launch("conv_dense_tc_kernel");
}

}  // namespace vx

using namespace vx;

extern "C" int vx_dense_conv_fwd(const vx_dense_conv_desc* d, const void* const* in, void* const* out, vx_stream_t stream) {
  if (!in || !out || !in[0] || !in[1] || !out[0]) { set_error("dense_conv_fwd: null pointer"); return VX_ERR_BAD_DESC; }
  return dense_conv_tc_fwd(d, (const float*)in[0], (const float*)in[1], (const float*)in[2], (float*)out[0], (cudaStream_t)stream);
}
