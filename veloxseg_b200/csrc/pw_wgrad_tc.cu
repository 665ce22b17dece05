// Weight gradient of the channel contractions on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulator in
// tensor memory, fp32-accurate through the 3-term split of vx_tc.cuh / pw_tc.cu).
//
//   dW[co, ci] += sum_{b,v} ypro(dY[b, co, v]) * xpro(X[b, ci, v]);     db[co] += sum ypro(dY[b, co, v])
//
// GEMM view per CTA:  D[m = co][n = ci] = A[m][k = voxel] * B[n][k]  with the reduction axis = voxels.  Both operands are
// NCDHW rows (voxels contiguous), i.e. already K-major: 16 bytes of global memory (4 voxels of one channel) are one row of
// an 8 x 16 B core matrix, so staging is a float4 load, the prologue (dropout mask on dY; InstanceNorm/LayerNorm affine,
// GELU, dropout on X), the hi/lo split and two 16-byte shared stores -- no transposition.  db comes from an extra
// all-ones B row.
//
// Shared layout of a chunk (64 voxels = 8 k-steps), per operand and per hi/lo half: k-step block g at g * BLK floats;
// inside a block row r, k -> (r / 8) * 64 + (k / 4) * 32 + (r % 8) * 4 + k % 4 (LBO 128 B, SBO 256 B).  BLK is an odd
// multiple of 64 floats so that the 32 lanes of a staging store (8 rows x 2 k-halves x 2 k-steps) hit 32 distinct banks.
// The MMA is issued with M = 128 although only Co <= 128 rows are staged: D rows are independent, the rows past Co read
// whatever follows in shared memory and are never read back from TMEM.
//
// Split-K: a CTA walks every gridDim.x-th (batch, chunk) pair, accumulates in TMEM, and folds its Co x (Ci + 1) tile
// into global memory with fp32 atomics (the caller zeroes dW / db).
#include "vx_kernels.h"
#include "vx_tc.cuh"

#ifdef VX_EMU
#define __grid_constant__
#endif

namespace vx {

constexpr int WT_THREADS = 256;
constexpr int WT_U = 6;              // float4 loads in flight per thread and staging round
constexpr int WT_KC = 64;            // voxels per chunk
constexpr int WT_KSTEPS = WT_KC / 8;

struct WgTcShape { int RGA, RGB, BLKA, BLKB, N, tmem_cols; };

#ifdef VX_EMU
static float g_emu_tmem[128][256];
static inline void split_tf32(float x, float& hi, float& lo) {
  uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&hi, &u, 4); lo = x - hi;
}
#endif

__global__ void __launch_bounds__(WT_THREADS) pw_wgrad_tc_kernel(const __grid_constant__ WgBatch batch,
                                                                const __grid_constant__ WgTcShape shp) {
  VX_PDL_ENTRY();
  const WgProblem& P = batch.p[blockIdx.y];
  const int S = batch.S, B = batch.B;
  const int Co = P.Co, Ci = P.Ci;
  const int nB = Ci + (P.db ? 1 : 0);                 // B rows that carry data (the last one is all ones)
  const int rgA = (Co + 7) >> 3, rgB = (nB + 7) >> 3; // row groups this problem stages (<= shp.RGA / shp.RGB)
  const int BLKA = shp.BLKA, BLKB = shp.BLKB;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  VX_DYN_SMEM(float, sm);
  float* A_hi = sm;
  float* A_lo = A_hi + (size_t)WT_KSTEPS * BLKA;
  float* B_hi = A_lo + (size_t)WT_KSTEPS * BLKA;
  float* B_lo = B_hi + (size_t)WT_KSTEPS * BLKB;
#ifndef VX_EMU
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
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
#endif

  const float yinv = P.y_drop_p > 0.f ? 1.0f / (1.0f - P.y_drop_p) : 1.f;
  const float xinv = P.x_drop_p > 0.f ? 1.0f / (1.0f - P.x_drop_p) : 1.f;
  const uint64_t soff = batch.seed_dev ? (uint64_t)__ldg(batch.seed_dev) : 0;
  const int nchunks = (S + WT_KC - 1) / WT_KC;

  // item i of an operand: 8 rows x 4 voxel quads per warp step -> (row, quad); returns the shared offset inside the
  // operand (floats) and the global coordinates
  auto decode = [&](int i, int BLK, int& row, int& quad, int& soff_f) {
    const int r_lo = i & 7, q_lo = (i >> 3) & 3, j = i >> 5;
    const int qg = j & 3, rg = j >> 2;
    row = rg * 8 + r_lo;
    quad = qg * 4 + q_lo;
    soff_f = (quad >> 1) * BLK + rg * 64 + (q_lo & 1) * 32 + r_lo * 4;
  };

  // One staging round = WT_U items per thread over the concatenated (A rows, B rows) item list: every global load of the
  // round is issued before the first one is consumed.  When a chunk fits in one round (all level-1 shapes) the loads of the
  // CTA's NEXT chunk are issued right after the MMAs of the current one, so their latency overlaps the tensor-core work.
  const int nA = rgA * 8 * (WT_KC / 4), nAB = nA + rgB * 8 * (WT_KC / 4);
  const int rounds = (nAB + WT_U * WT_THREADS - 1) / (WT_U * WT_THREADS);
  float4 v[WT_U];
  int so[WT_U], rw[WT_U], gvv[WT_U];       // shared offset (< 0: no item; B items carry bit 30), row / channel, first voxel
  bool ok[WT_U];
  auto load_round = [&](int b, int vbase, int rd) {
#pragma unroll
    for (int u = 0; u < WT_U; ++u) {
      const int i = (rd * WT_U + u) * WT_THREADS + tid;
      so[u] = -1; ok[u] = false; rw[u] = 0; gvv[u] = 0;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i >= nAB) continue;
      int quad = 0;
      if (i < nA) {
        decode(i, BLKA, rw[u], quad, so[u]);
        gvv[u] = vbase + quad * 4;
        ok[u] = rw[u] < Co && gvv[u] < S;
        if (ok[u]) v[u] = __ldg(reinterpret_cast<const float4*>(P.dY + ((size_t)b * Co + rw[u]) * S + gvv[u]));
      } else {
        decode(i - nA, BLKB, rw[u], quad, so[u]);
        so[u] |= 1 << 30;
        gvv[u] = vbase + quad * 4;
        ok[u] = rw[u] < Ci && gvv[u] < S;
        if (ok[u]) {
          int c = rw[u], s2 = 0;
          while (s2 < P.nsrc - 1 && c >= P.src[s2].C) { c -= P.src[s2].C; ++s2; }
          v[u] = __ldg(reinterpret_cast<const float4*>(P.src[s2].ptr + ((size_t)b * P.src[s2].C + c) * S + gvv[u]));
        } else if (rw[u] == Ci && P.db && gvv[u] < S) {
          v[u] = make_float4(1.f, 1.f, 1.f, 1.f);
        }
      }
    }
  };
  auto store_round = [&](int b) {
#pragma unroll
    for (int u = 0; u < WT_U; ++u) {
      if (so[u] < 0) continue;
      const bool isB = (so[u] >> 30) & 1;
      const int off = so[u] & ~(1 << 30);
      float x[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
      if (ok[u]) {
        if (!isB) {
          if (P.y_drop_p > 0.f) {
            float ms[4];
            dropout_scale4(P.y_seed + soff, P.y_site, ((uint64_t)b * Co + rw[u]) * (uint64_t)S + gvv[u], P.y_drop_p, yinv, ms);
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] *= ms[e];
          }
        } else if (P.xpro == PRO_AFFINE) {
          const int k = b * P.x_bstride + rw[u];
          const float a = __ldg(P.xa + k), c = __ldg(P.xc + k);
#pragma unroll
          for (int e = 0; e < 4; ++e) x[e] = fmaf(x[e], a, c);
        } else if (P.xpro == PRO_GELU || P.xpro == PRO_GELU_DROPOUT) {
          const float4 gq = gelu4_call(make_float4(x[0], x[1], x[2], x[3]));
          x[0] = gq.x; x[1] = gq.y; x[2] = gq.z; x[3] = gq.w;
          if (P.xpro == PRO_GELU_DROPOUT) {
            float ms[4];
            dropout_scale4(P.x_seed + soff, P.x_site, ((uint64_t)b * Ci + rw[u]) * (uint64_t)S + gvv[u], P.x_drop_p, xinv, ms);
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] *= ms[e];
          }
        }
      }
      float h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (batch.prec) { h[e] = bf16_round(x[e]); l[e] = 0.f; } else split_tf32(x[e], h[e], l[e]);
      }
      *reinterpret_cast<float4*>((isB ? B_hi : A_hi) + off) = make_float4(h[0], h[1], h[2], h[3]);
      if (!batch.prec) *reinterpret_cast<float4*>((isB ? B_lo : A_lo) + off) = make_float4(l[0], l[1], l[2], l[3]);
    }
  };

  int it = 0;
  const int total_ck = B * nchunks;
  bool prefetched = false;
  if (rounds == 1 && (int)blockIdx.x < total_ck) {
    load_round(blockIdx.x / nchunks, (blockIdx.x % nchunks) * WT_KC, 0);
    prefetched = true;
  }
  for (int ck = blockIdx.x; ck < total_ck; ck += gridDim.x, ++it) {
    const int b = ck / nchunks, vbase = (ck % nchunks) * WT_KC;
#ifndef VX_EMU
    if (it > 0) mbar_wait(smem_u32(&mbar), (uint32_t)((it - 1) & 1));     // the previous chunk's MMAs have read the tiles
#else
    __syncthreads();
#endif
    for (int rd = 0; rd < rounds; ++rd) {
      if (!(prefetched && rd == 0)) load_round(b, vbase, rd);
      store_round(b);
    }
    prefetched = false;
#ifndef VX_EMU
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t idesc = umma_idesc_tf32(shp.N);
      const uint32_t a_hi = smem_u32(A_hi), a_lo = smem_u32(A_lo), b_hi = smem_u32(B_hi), b_lo = smem_u32(B_lo);
#pragma unroll 1
      for (int g8 = 0; g8 < WT_KSTEPS; ++g8) {
        const uint32_t ao = (uint32_t)(g8 * BLKA * 4), bo = (uint32_t)(g8 * BLKB * 4);
        const uint64_t dah = umma_desc(a_hi + ao, 128u, 256u), dal = umma_desc(a_lo + ao, 128u, 256u);
        const uint64_t dbh = umma_desc(b_hi + bo, 128u, 256u), dbl = umma_desc(b_lo + bo, 128u, 256u);
        if (batch.prec) {
          umma_tf32(tmem, dah, dbh, idesc, (it > 0 || g8 > 0) ? 1u : 0u);
        } else {
          umma_tf32(tmem, dal, dbh, idesc, (it > 0 || g8 > 0) ? 1u : 0u);
          umma_tf32(tmem, dah, dbl, idesc, 1u);
          umma_tf32(tmem, dah, dbh, idesc, 1u);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar))
                   : "memory");
    }
#else
    __syncthreads();
    if (tid == 0) {      // software model of the MMA on the same shared layout (rows past the staged groups are skipped)
      for (int m = 0; m < rgA * 8; ++m)
        for (int n = 0; n < rgB * 8; ++n) {
          float acc = it > 0 ? g_emu_tmem[m][n] : 0.f;
          for (int g8 = 0; g8 < WT_KSTEPS; ++g8)
            for (int k = 0; k < 8; ++k) {
              const int ao = g8 * BLKA + (m >> 3) * 64 + (k >> 2) * 32 + (m & 7) * 4 + (k & 3);
              const int bo = g8 * BLKB + (n >> 3) * 64 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3);
              acc += batch.prec ? A_hi[ao] * B_hi[bo] : A_lo[ao] * B_hi[bo] + A_hi[ao] * B_lo[bo] + A_hi[ao] * B_hi[bo];
            }
          g_emu_tmem[m][n] = acc;
        }
    }
#endif
    if (rounds == 1 && ck + (int)gridDim.x < total_ck) {
      const int nk = ck + gridDim.x;
      load_round(nk / nchunks, (nk % nchunks) * WT_KC, 0);
      prefetched = true;
    }
  }

  // ---- fold the accumulator tile into global memory
  if (it > 0) {
#ifndef VX_EMU
    mbar_wait(smem_u32(&mbar), (uint32_t)((it - 1) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#else
    __syncthreads();
#endif
    // a warp reads the 32 TMEM lanes of its quadrant (warp % 4); the two warps of a quadrant split the column groups
    const int quad = warp & 3, co = quad * 32 + lane;
    if (quad * 32 < Co) {
#pragma unroll 1
      for (int c0 = (warp >> 2) * 8; c0 < rgB * 8; c0 += 16) {
        float r[8];
#ifndef VX_EMU
        uint32_t q[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7])
                     : "r"(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0)
                     : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __uint_as_float(q[j]);
#else
        for (int j = 0; j < 8; ++j) r[j] = co < rgA * 8 ? g_emu_tmem[co][c0 + j] : 0.f;
#endif
        if (co < Co) {
          // 16-byte vector reductions where a quad of columns lies inside the row (4x fewer L2 atomic operations: the fold of
          // ~300 CTAs onto one Co x Ci tile was 10 of the kernel's 30-40 us, profiles/r4a_wgrad_tc_L*.source.txt)
          const bool vec = (P.ld & 3) == 0 && ((uintptr_t)P.dW & 15) == 0;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int ci0 = c0 + 4 * h;
            float* dst = P.dW + (size_t)co * P.ld + ci0;
            if (vec && ci0 + 3 < Ci) {
#ifndef VX_EMU
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(r[4 * h]), "f"(r[4 * h + 1]), "f"(r[4 * h + 2]),
                           "f"(r[4 * h + 3])
                           : "memory");
#else
              for (int j = 0; j < 4; ++j) atomicAdd(dst + j, r[4 * h + j]);
#endif
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int ci = ci0 + j;
                if (ci < Ci) atomicAdd(dst + j, r[4 * h + j]);
                else if (ci == Ci && P.db) atomicAdd(P.db + co, r[4 * h + j]);
              }
            }
          }
        }
      }
    }
  }
#ifndef VX_EMU
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)shp.tmem_cols) : "memory");
  }
#endif
}

static int g_wt_enabled = 1, g_wt_min_s = 512;
void pw_wgrad_tc_set(int enabled, int min_s) {
  if (enabled >= 0) g_wt_enabled = enabled;
  if (min_s >= 0) g_wt_min_s = min_s;
}

// Returns VX_OK when the batch ran on the tensor-core kernel, 1 when it does not qualify (the caller uses the SIMT
// kernel), or a negative vx_status.
int pw_wgrad_tc(const WgBatch& batch, cudaStream_t stream) {
  if (!g_wt_enabled) return 1;
  const int S = batch.S;
  if (S < g_wt_min_s || (S & 3)) return 1;
  int maxCo = 0, maxNB = 0;
  double bytes = 0.0, flops = 0.0;
  for (int i = 0; i < batch.nprob; ++i) {
    const WgProblem& P = batch.p[i];
    int cs = 0;
    for (int s = 0; s < P.nsrc; ++s) {
      cs += P.src[s].C;
      if ((uintptr_t)P.src[s].ptr & 15) return 1;
    }
    if (cs != P.Ci) { set_error("pw_wgrad: source channels %d != Ci %d", cs, P.Ci); return VX_ERR_BAD_DESC; }
    if ((uintptr_t)P.dY & 15) return 1;
    if (P.xpro != PRO_NONE && P.xpro != PRO_AFFINE && P.xpro != PRO_GELU && P.xpro != PRO_GELU_DROPOUT) return 1;
    maxCo = P.Co > maxCo ? P.Co : maxCo;
    const int nb = P.Ci + (P.db ? 1 : 0);
    maxNB = nb > maxNB ? nb : maxNB;
    bytes += 4.0 * batch.B * S * (P.Ci + P.Co) + 4.0 * P.Ci * P.Co;
    flops += 2.0 * batch.B * S * P.Ci * P.Co;
  }
  if (maxCo > 128) return 1;
  WgTcShape shp{};
  shp.RGA = (maxCo + 7) >> 3;
  shp.N = (maxNB + 15) & ~15;
  if (shp.N > 256) return 1;
  shp.RGB = shp.N >> 3;
  shp.BLKA = (shp.RGA | 1) * 64;            // odd multiple of 64 floats
  shp.BLKB = (shp.RGB | 1) * 64;
  shp.tmem_cols = 32;
  while (shp.tmem_cols < shp.N) shp.tmem_cols <<= 1;
  // A_lo's last k-step is read 16 row groups (1024 floats) deep: the B tiles that follow must cover the overrun
  size_t fl = (size_t)2 * WT_KSTEPS * (shp.BLKA + shp.BLKB);
  const size_t tail = (size_t)2 * WT_KSTEPS * shp.BLKB;
  if (tail < 1024) fl += 1024 - tail;
  const size_t smem = fl * sizeof(float);
  if (smem > 200 * 1024) return 1;          // <= 113 KB keeps two CTAs per SM (staging of one overlaps the MMAs of the other)
  const int nchunks = batch.B * cdiv(S, WT_KC);
  // CTAs per SM over all problems of the batch (split-K factor); VX_WGRAD_TC_CTAS_PER_SM is the tuning probe
  static const int per_sm = [] { const char* e = getenv("VX_WGRAD_TC_CTAS_PER_SM"); const int v = e ? atoi(e) : 2; return v >= 1 && v <= 8 ? v : 2; }();
  int nsplit = (per_sm * kSMs) / batch.nprob;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > nchunks) nsplit = nchunks;
  VX_SET_SMEM(pw_wgrad_tc_kernel, smem);
  WgBatch launch = batch;
  launch.seed_dev = get_seed_dev();
  launch.prec = precision_mode();
  prof_bytes(bytes);
  prof_flops(flops);
  VX_LAUNCH(pw_wgrad_tc_kernel, dim3(nsplit, batch.nprob), dim3(WT_THREADS), smem, stream, launch, shp);
  return check_launch("pw_wgrad_tc_kernel");
}

}  // namespace vx
