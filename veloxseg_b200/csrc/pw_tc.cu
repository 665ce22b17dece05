// Channel contraction (1x1x1 conv) on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with the accumulator in
// tensor memory, fp32-accurate through a 3-term split (3xTF32).
//
//   Y[b, co, v] = epi( sum_ci W[co, ci] * pro(X[b, ci, v]) + bias[co] )          (same contract as pw_kernel)
//
// GEMM view per CTA:  D[m = voxel (128)][n = out channel (<= 256)] = A[m][k = in channel] * B[k][n]
//   A = X^T tile.  X is NCDHW (voxels contiguous per channel), i.e. MN-major; kind::tf32 with an MN-major SWIZZLE_NONE
//       operand returns zeros on this part (tools/bringup/tc_micro.cu), so the tile is transposed on the way through
//       registers (where the prologue -- InstanceNorm / LayerNorm affine, GELU, dropout -- is applied anyway) into the
//       K-major no-swizzle canonical layout: 8 x 16 B core matrices (8 voxels x 4 channels), the two halves of a k-step
//       128 B apart (LBO), groups of 8 voxels 256 B apart (SBO), groups of 8 channels 4 KB apart (one descriptor per
//       k-step).  Lanes rotate which of their 4 voxels they store so that every store instruction is bank-conflict-free.
//   B = W as (n, k), same K-major layout (LBO 128 B, SBO 256 B).
//   Every fp32 operand x is split as x = hi + lo with hi = the top 19 bits (exact in tf32) and lo = x - hi (exact in
//   fp32, <= 13 significant bits): D = A_lo B_hi + A_hi B_lo + A_hi B_hi drops only the lo*lo term (2^-20 relative), so
//   the result meets the fp32 parity bar; the tensor pipe has two orders of magnitude of headroom at these shapes.
// One elected thread issues the K/8 x 3 MMAs and commits them to an mbarrier; the 4 warps then read their 32 TMEM lanes
// (tcgen05.ld 32x32b: thread = voxel, registers = output channels), apply the epilogue and store: for a fixed channel
// the 32 lanes of a warp write 128 contiguous bytes.
//
// Used for the level-1/2 problems (S >= 512 voxels, S % 4 == 0, K <= 256, N <= 256); everything else stays on the
// SIMT kernels in pointwise.cu.
#include "vx_kernels.h"
#include "vx_tc.cuh"

#ifndef VX_EMU

namespace vx {

constexpr int TC_THREADS = 256;      // two warps per TMEM lane quadrant: they alternate channel groups / column groups

struct PwTcShape { int Kpad, Npad, tmem_cols; };

__global__ void __launch_bounds__(TC_THREADS) pw_tc_kernel(const __grid_constant__ PwBatch batch,
                                                           const __grid_constant__ PwTcShape shp) {
  const int pi = blockIdx.z / batch.B, b = blockIdx.z % batch.B;
  const PwProblem& P = batch.p[pi];
  const int S = batch.S, Ci = P.Ci, Co = P.Co;
  const int Kpad = shp.Kpad, Npad = shp.Npad;
  const int v0 = blockIdx.x * TC_M;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wq = warp & 3, wp = warp >> 2;          // voxel quadrant (32 rows of the tile), half of the CTA

  VX_DYN_SMEM(float, sm);
  float* A_hi = sm;                                  // [Kpad/8][16 voxel groups][2 k-halves][8 voxels][4 k]
  float* A_lo = A_hi + (size_t)Kpad * TC_M;
  float* B_hi = A_lo + (size_t)Kpad * TC_M;          // [Kpad/8][Npad/8][2 k-halves][8 n][4 k]
  float* B_lo = B_hi + (size_t)Kpad * Npad;
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_slot;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"((uint32_t)shp.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(smem_u32(&mbar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  const bool pro_drop = P.pro == PRO_DROPOUT || P.pro == PRO_GELU_DROPOUT;
  const float pinv = pro_drop ? 1.0f / (1.0f - P.pro_drop_p) : 1.f;
  const uint64_t soff = batch.seed_dev ? (uint64_t)__ldg(batch.seed_dev) : 0;

  // ---- A: a warp covers 8 channels x 4 chunks of 4 voxels per step (64 contiguous bytes per channel from global).
  // lane bits [1:0] = k % 4, [2] = chunk % 2 select the bank; bits [4:3] = r pick (k / 4, chunk / 2), which do not, and
  // rotate the voxel a lane stores in step j, so the 32 lanes of a store hit 32 distinct banks.
  {
    const int r = lane >> 3;
    const int k = (lane & 3) + 4 * (r & 1);
    const int cw = ((lane >> 2) & 1) + 2 * (r >> 1);          // chunk within the warp's 4
    const int ngroups = Kpad >> 3;
    constexpr int U = 2;                                      // channel groups per batch of loads
    const int gh = (ngroups + 1) >> 1, gbeg = wp * gh, gend = min(ngroups, gbeg + gh);   // this half's channel groups
#pragma unroll 1
    for (int g0 = gbeg; g0 < gend; g0 += U) {
      float x[U][2][4];
      float pa[U], pc[U];
      bool live[U];
      // phase 1: every global load of this batch
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int ci = (g0 + u) * 8 + k;
        live[u] = (g0 + u) < gend && ci < Ci;
        const float* xrow = P.src[0].ptr;
        if (live[u]) {
          int c = ci, sidx = 0;
          while (sidx < P.nsrc - 1 && c >= P.src[sidx].C) { c -= P.src[sidx].C; ++sidx; }
          xrow = P.src[sidx].ptr + ((size_t)b * P.src[sidx].C + c) * S;
        }
        pa[u] = 1.f; pc[u] = 0.f;
        if (P.pro == PRO_AFFINE && live[u]) {
          const int q = b * P.pro_bstride + ci;
          pa[u] = __ldg(P.pro_a + q); pc[u] = __ldg(P.pro_c + q);
        }
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
          const int v = v0 + (wq * 4 + cw + rep * 16) * 4;
#pragma unroll
          for (int i = 0; i < 4; ++i) x[u][rep][i] = 0.f;
          if (live[u]) {
            if (v + 3 < S) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(xrow + v));
              x[u][rep][0] = t.x; x[u][rep][1] = t.y; x[u][rep][2] = t.z; x[u][rep][3] = t.w;
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) if (v + i < S) x[u][rep][i] = __ldg(xrow + v + i);
            }
          }
        }
      }
      // phase 2: prologue, split, transposing stores
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (g0 + u >= gend) break;
        const int ci = (g0 + u) * 8 + k;
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
          const int chunk = wq * 4 + cw + rep * 16;
          const int v = v0 + chunk * 4;
          float* xx = x[u][rep];
          if (live[u]) {
            if (P.pro == PRO_AFFINE) {
#pragma unroll
              for (int i = 0; i < 4; ++i) xx[i] = fmaf(xx[i], pa[u], pc[u]);
            } else if (P.pro != PRO_NONE) {
              // the 4 voxels of a chunk are one Philox block (v % 4 == 0 and S % 4 == 0 on this path): one call, not four
              if (P.pro == PRO_GELU || P.pro == PRO_GELU_DROPOUT) {
                const float4 gq = gelu4_call(make_float4(xx[0], xx[1], xx[2], xx[3]));
                xx[0] = gq.x; xx[1] = gq.y; xx[2] = gq.z; xx[3] = gq.w;
              }
              if (pro_drop) {
                float ms[4];
                dropout_scale4(P.pro_seed + soff, P.pro_site, ((uint64_t)b * Ci + ci) * (uint64_t)S + v, P.pro_drop_p, pinv, ms);
#pragma unroll
                for (int i = 0; i < 4; ++i) xx[i] *= ms[i];
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) if (v + i >= S) xx[i] = 0.f;      // rows past the tensor stay zero
          }
          // element (m = chunk * 4 + e, k): float offset (m / 8) * 64 + (k / 4) * 32 + (m % 8) * 4 + k % 4
          const int base = (g0 + u) * (TC_M * 8) + (chunk >> 1) * 64 + (k >> 2) * 32 + (chunk & 1) * 16 + (k & 3);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int e = (j + r) & 3;
            const float xe = e == 0 ? xx[0] : e == 1 ? xx[1] : e == 2 ? xx[2] : xx[3];
            if (batch.prec) {
              A_hi[base + e * 4] = bf16_round(xe);
            } else {
              float h, l;
              split_tf32(xe, h, l);
              A_hi[base + e * 4] = h;
              A_lo[base + e * 4] = l;
            }
          }
        }
      }
    }
  }
  // ---- B: weights (n, k) -> K-major core matrices.  16-byte loads along the contiguous axis of W (rows are 16-byte
  // aligned and 4-element granular on this path, checked by pw_tc_forward), two per thread in flight.
  {
    const int Q = P.transposed ? (Npad >> 2) : (Kpad >> 2);      // float4 per row of the operand as stored
    const int total = (P.transposed ? Kpad : Npad) * Q;
#pragma unroll 1
    for (int base = tid; base < total; base += 2 * TC_THREADS) {
      float4 w[2];
      int row[2], q4[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int idx = base + u * TC_THREADS;
        row[u] = idx / Q; q4[u] = (idx - row[u] * Q) * 4;
        w[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < total) {
          const int n = P.transposed ? q4[u] : row[u], kk = P.transposed ? row[u] : q4[u];
          if (n < Co && kk < Ci) {
            int ld;
            w[u] = __ldg(reinterpret_cast<const float4*>(pw_w_row(P, n, kk, ld)));
          }
        } else {
          row[u] = -1;
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (row[u] < 0) continue;
        const float ww[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
        float h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (batch.prec) { h[i] = bf16_round(ww[i]); l[i] = 0.f; } else split_tf32(ww[i], h[i], l[i]);
        }
        if (!P.transposed) {        // (n = row, k = q4..q4+3): four consecutive floats of one core-matrix row
          const int n = row[u], kk = q4[u];
          const int off = (kk >> 3) * (Npad * 8) + (n >> 3) * 64 + ((kk & 7) >> 2) * 32 + (n & 7) * 4;
          *reinterpret_cast<float4*>(B_hi + off) = make_float4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<float4*>(B_lo + off) = make_float4(l[0], l[1], l[2], l[3]);
        } else {                    // (k = row, n = q4..q4+3): four rows of one core matrix, same k
          const int kk = row[u], n = q4[u];
          const int off = (kk >> 3) * (Npad * 8) + (n >> 3) * 64 + ((kk & 7) >> 2) * 32 + (n & 7) * 4 + (kk & 3);
#pragma unroll
          for (int i = 0; i < 4; ++i) { B_hi[off + i * 4] = h[i]; B_lo[off + i * 4] = l[i]; }
        }
      }
    }
  }
  // generic-proxy smem writes -> visible to the tensor-core (async) proxy; TMEM address -> visible to all threads
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;

  if (tid == 0) {
    const uint32_t idesc = umma_idesc_tf32(Npad);
    const uint32_t a_hi = smem_u32(A_hi), a_lo = smem_u32(A_lo), b_hi = smem_u32(B_hi), b_lo = smem_u32(B_lo);
    const uint32_t a_lbo = 128u, a_sbo = 256u, b_lbo = 128u, b_sbo = 256u;
    const int ngroups = Kpad >> 3;
    for (int g8 = 0; g8 < ngroups; ++g8) {
      const uint32_t ao = (uint32_t)g8 * (TC_M * 8 * 4), bo = (uint32_t)g8 * (uint32_t)(Npad * 8 * 4);
      const uint64_t dah = umma_desc(a_hi + ao, a_lbo, a_sbo), dal = umma_desc(a_lo + ao, a_lbo, a_sbo);
      const uint64_t dbh = umma_desc(b_hi + bo, b_lbo, b_sbo), dbl = umma_desc(b_lo + bo, b_lbo, b_sbo);
      if (batch.prec) {                               // bf16 numerics: operands are exact in tf32, one product
        umma_tf32(tmem, dah, dbh, idesc, g8 > 0 ? 1u : 0u);
      } else {
        umma_tf32(tmem, dal, dbh, idesc, g8 > 0 ? 1u : 0u);
        umma_tf32(tmem, dah, dbl, idesc, 1u);
        umma_tf32(tmem, dah, dbh, idesc, 1u);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar))
                 : "memory");
  }
  mbar_wait(smem_u32(&mbar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- epilogue: thread = voxel row (TMEM lane), 8 output channels per tcgen05.ld.  Per batch of 8: pointers and every
  // global read first, then the arithmetic, then the stores (as far as the compiler knows the stores may alias the
  // residual tensors, so the order has to be explicit for the loads to overlap).
  {
    const int v = v0 + wq * 32 + lane;
    const bool vok = v < S;
    const float dinv = P.drop_p > 0.f ? 1.0f / (1.0f - P.drop_p) : 1.f;
    const bool heavy = P.act == 1 || P.mulgrad || P.drop_p > 0.f;
    const uint32_t trow = tmem + ((uint32_t)(wq * 32) << 16);
#pragma unroll 1
    for (int c0 = wp * 8; c0 < Npad; c0 += 16) {
      uint32_t r[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(trow + (uint32_t)c0)
                   : "memory");
      // the 8 channels of a batch share one output segment (segment sizes are multiples of 8 on this path)
      float* obase;
      const float* bptr = nullptr;
      if (!P.transposed) {
        int seg = 0, seg_off = 0;
        while (seg < P.nseg - 1 && c0 >= seg_off + P.seg[seg].n) { seg_off += P.seg[seg].n; ++seg; }
        obase = P.seg[seg].out + ((size_t)b * P.seg[seg].n + (c0 - seg_off)) * S + v;
        if (P.seg[seg].bias) bptr = P.seg[seg].bias + (c0 - seg_off);
      } else {
        obase = P.seg[0].out + ((size_t)b * Co + c0) * S + v;
      }
      const size_t lbase = ((size_t)b * Co + c0) * S + v;      // index in the logical (B, Co, S) tensor
      // Optional epilogue streams: the pointer tests are CTA-uniform, so each stream is one predictable branch around 8
      // independent loads (not 8 x 3 predicated selects); `nlive` = channels of this batch that exist for this voxel.
      const int nlive = vok ? (Co - c0 < 8 ? Co - c0 : 8) : 0;
      float bias[8], mg[8], r1[8], r2[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { bias[j] = 0.f; mg[j] = 0.f; r1[j] = 0.f; r2[j] = 0.f; }
      if (bptr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) if (c0 + j < Co) bias[j] = __ldg(bptr + j);
      }
      if (P.mulgrad) {
        const float* q = P.mulgrad + lbase;
#pragma unroll
        for (int j = 0; j < 8; ++j) if (j < nlive) mg[j] = __ldg(q + (size_t)j * S);
      }
      if (P.res) {
        const float* q = P.res + lbase;
#pragma unroll
        for (int j = 0; j < 8; ++j) if (j < nlive) r1[j] = __ldg(q + (size_t)j * S);
      }
      if (P.res2) {
        const float* q = P.res2 + lbase;
#pragma unroll
        for (int j = 0; j < 8; ++j) if (j < nlive) r2[j] = __ldg(q + (size_t)j * S);
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = __uint_as_float(r[j]) + bias[j];
      if (heavy) {
        // dropout: lanes 4q..4q+3 hold 4 consecutive voxels = one Philox block per channel.  Lane (lane & 3) = i computes the
        // blocks of channels i and i + 4 and the group exchanges the keep-scales by shuffle: 2 Philox calls per lane, not 8.
        float keep[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) keep[j] = 1.f;
        if (P.drop_p > 0.f) {
          const int li = lane & 3;
          const size_t vb = lbase - (size_t)li;                      // index of the group's first voxel, channel c0
          float mine[2][4];
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2)
            dropout_scale4(P.seed + soff, P.site, vb + (size_t)(li + 4 * h2) * S, P.drop_p, dinv, mine[h2]);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            // channel j's block lives in lane (j & 3) of the group, half j >> 2; this lane needs component li of it
            float got = 1.f;
#pragma unroll
            for (int comp = 0; comp < 4; ++comp) {
              const float cand = __shfl_sync(0xffffffffu, mine[j >> 2][comp], (lane & ~3) | (j & 3));
              if (comp == li) got = cand;
            }
            keep[j] = got;
          }
        }
        if (P.act == 1) {          // out-of-line 4-wide copies: the unrolled erff / expf would be ~10 KB of SASS
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const float4 gq = gelu4_call(make_float4(y[4 * h2], y[4 * h2 + 1], y[4 * h2 + 2], y[4 * h2 + 3]));
            y[4 * h2] = gq.x; y[4 * h2 + 1] = gq.y; y[4 * h2 + 2] = gq.z; y[4 * h2 + 3] = gq.w;
          }
        }
        if (P.mulgrad) {
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const float4 gq = gelu_grad4_call(make_float4(mg[4 * h2], mg[4 * h2 + 1], mg[4 * h2 + 2], mg[4 * h2 + 3]));
            y[4 * h2] *= gq.x; y[4 * h2 + 1] *= gq.y; y[4 * h2 + 2] *= gq.z; y[4 * h2 + 3] *= gq.w;
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] *= keep[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = fmaf(P.res_scale, r1[j], y[j]) + r2[j];      // r1 / r2 are zero without their stream
      if (batch.prec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = bf16_round(y[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) if (j < nlive) obase[(size_t)j * S] = y[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)shp.tmem_cols) : "memory");
  }
}

static int g_tc_enabled = 1;
void pw_tc_set(int enabled) { g_tc_enabled = enabled; }

// Returns VX_OK when the batch was launched on the tensor-core kernel, 1 when it does not qualify (caller falls back to
// the SIMT kernel), or a negative vx_status.
int pw_tc_forward(const PwBatch& batch, cudaStream_t stream) {
  if (!g_tc_enabled) return 1;
  const int S = batch.S;
  if (S < 512 || (S & 3)) return 1;
  int Kmax = 0, Nmax = 0;
  for (int i = 0; i < batch.nprob; ++i) {
    const PwProblem& P = batch.p[i];
    Kmax = P.Ci > Kmax ? P.Ci : Kmax;
    Nmax = P.Co > Nmax ? P.Co : Nmax;
    for (int s = 0; s < P.nsrc; ++s)
      if (((uintptr_t)P.src[s].ptr & 15) || ((size_t)P.src[s].C * S) % 4) return 1;
    // weights: 16-byte rows; forward orientation: output segments in multiples of 8 channels, K in multiples of 4
    int off = 0;
    for (int s = 0; s < P.nseg; ++s) {
      if (((uintptr_t)P.seg[s].W & 15) || (P.seg[s].ld & 3)) return 1;
      if (!P.transposed && s < P.nseg - 1 && (P.seg[s].n & 7)) return 1;
      if (P.transposed && (P.seg[s].n & 3)) return 1;          // input-channel segments: float4 along co never straddles
      off += P.seg[s].n;
    }
    if ((P.Ci & 3) || (P.transposed && (P.Co & 3))) return 1;
  }
  PwTcShape shp{};
  shp.Kpad = (Kmax + 7) & ~7;
  shp.Npad = (Nmax + 15) & ~15;
  if (shp.Kpad > 256 || shp.Npad > 256) return 1;
  const size_t smem = sizeof(float) * 2 * (size_t)shp.Kpad * (TC_M + shp.Npad);
  if (smem > 200 * 1024) return 1;
  shp.tmem_cols = 32;
  while (shp.tmem_cols < shp.Npad) shp.tmem_cols <<= 1;
  VX_SET_SMEM(pw_tc_kernel, smem);
  PwBatch launch = batch;
  launch.seed_dev = get_seed_dev();
  launch.prec = precision_mode();
  dim3 grid(cdiv(S, TC_M), 1, batch.nprob * batch.B);
  VX_LAUNCH(pw_tc_kernel, grid, dim3(TC_THREADS), smem, stream, launch, shp);
  return check_launch("pw_tc_kernel");
}

}  // namespace vx

#endif  // VX_EMU
