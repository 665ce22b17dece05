// Channel contraction (1x1x1 conv) on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with the accumulator AND the
// activation operand in tensor memory, fp32-accurate through a 3-term split (3xTF32).
//
//   Y[b, co, v] = epi( sum_ci W[co, ci] * pro(X[b, ci, v]) + bias[co] )          (same contract as pw_kernel)
//
// GEMM view per CTA:  D[m = voxel (128)][n = out channel (<= 256)] = A[m][k = in channel] * B[k][n]
//   A = X^T tile.  X is NCDHW (voxels contiguous per channel).  The tile never passes through shared memory: thread =
//       voxel (TMEM lane), it reads its K channel values with warp-coalesced 128-byte loads, applies the prologue
//       (InstanceNorm / LayerNorm affine, GELU, dropout), splits each value into hi + lo and writes both straight into
//       TENSOR MEMORY with tcgen05.st (row m of the MMA's A operand = TMEM lane m, element k = one 32-bit column;
//       confirmed on a B200 by tools/bringup/tc_probe4.cu, profiles/r2i_tc_probe4_tmem_a.txt).  The MMA then takes A from
//       TMEM (the ".ts" form) -- no transposition, no shared-memory staging, no proxy fence for A.
//   B = W as (n, k) in shared memory, K-major no-swizzle canonical layout: 8 x 16 B core matrices, the two halves of a
//       k-step 128 B apart (LBO), groups of 8 rows 256 B apart (SBO), one descriptor per k-step.
//   Every fp32 operand x is split as x = hi + lo with hi = the top 19 bits (exact in tf32) and lo = x - hi (exact in
//   fp32, <= 13 significant bits): D = A_lo B_hi + A_hi B_lo + A_hi B_hi drops only the lo*lo term (2^-20 relative), so
//   the result meets the fp32 parity bar; the tensor pipe has two orders of magnitude of headroom at these shapes.
// TMEM columns of a CTA: D [0, Npad) | A_hi [Npad, Npad + Kpad) | A_lo [Npad + Kpad, Npad + 2 Kpad), rounded up to a power
// of two (128 at level 1, 256 at level 2): 2-4 CTAs share an SM, which is what hides the load -> st -> MMA -> ld -> store
// chain of one tile behind the others (the whole problem is one wave: 432 tiles at level 1).
// One elected thread issues the K/8 x 3 MMAs and commits them to an mbarrier; the warps then read their 32 TMEM lanes
// (tcgen05.ld 32x32b: thread = voxel, registers = output channels), apply the epilogue and store: for a fixed channel
// the 32 lanes of a warp write 128 contiguous bytes.
//
// Used for the level-1/2 problems (S >= 512 voxels, S % 4 == 0, K <= 128, N <= 256); everything else stays on the
// SIMT kernels in pointwise.cu.
#include "vx_kernels.h"
#include "vx_tc.cuh"

#ifndef VX_EMU

namespace vx {

constexpr int TC_THREADS = 256;      // two warps per TMEM lane quadrant: they alternate channel groups / column groups

struct PwTcShape { int Kpad, Npad, tmem_cols; };

// keep-scales (0 or 1 / (1 - p)) of the 8 channels c0 .. c0 + 7 at this thread's voxel.  `idx` = element index of (c0, voxel),
// channels S elements apart.
VX_DEV void drop_keep8(uint64_t seed, uint32_t site, size_t idx, size_t S, int lane, float p, float inv_keep, float (&keep)[8]) {
  (void)lane;
  const uint32_t key = rng_key(seed, site);
#pragma unroll
  for (int j = 0; j < 8; ++j) keep[j] = keep_from_bits(rng_word(key, idx + (size_t)j * S), p, inv_keep);
}

// The kernel body runs ONCE per thread (one 128-voxel tile per CTA), so its cost is first of all the number of distinct
// instructions a warp executes: straight-line code is fetched from L2 at 10-20 cycles per instruction (ncu: stall_no_instruction
// 9-11 warps per issue on the unrolled version, 130 KB of SASS, 1 700-4 000 instructions per warp, 20-25 us per launch --
// profiles/r2l_pw_tc_L1/L2.digest.txt).  Hence every phase is a ROLLED loop over 8-channel groups with a small body that hits
// the instruction cache from its second iteration on, and memory latency is hidden by prefetching exactly one iteration ahead
// (two register sets, rotated) plus the other CTAs of the SM, not by unrolling.
__global__ void __launch_bounds__(TC_THREADS, 3) pw_tc_kernel(const __grid_constant__ PwBatch batch,
                                                              const __grid_constant__ PwTcShape shp) {
  VX_PDL_ENTRY();
  const int pi = blockIdx.z / batch.B, b = blockIdx.z % batch.B;
  const PwProblem& P = batch.p[pi];
  const int S = batch.S, Ci = P.Ci, Co = P.Co;
  const int Kpad = shp.Kpad, Npad = shp.Npad;
  const int v0 = blockIdx.x * TC_M;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wq = warp & 3, wp = warp >> 2;          // voxel quadrant (32 rows of the tile), half of the CTA

  VX_DYN_SMEM(float, sm);
  float* B_hi = sm;                                  // [Kpad/8][Npad/8][2 k-halves][8 n][4 k]
  float* B_lo = B_hi + (size_t)Kpad * Npad;
  float* bias_s = B_lo + (size_t)Kpad * Npad;        // [Npad]  bias of every output column (0 where there is none)
  float* proa_s = bias_s + Npad;                     // [Kpad]  PRO_AFFINE scale / shift of this batch item (1 / 0 otherwise)
  float* proc_s = proa_s + Kpad;
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
  const bool pro_gelu = P.pro == PRO_GELU || P.pro == PRO_GELU_DROPOUT;
  const float pinv = pro_drop ? 1.0f / (1.0f - P.pro_drop_p) : 1.f;
  const uint64_t soff = batch.seed_dev ? (uint64_t)__ldg(batch.seed_dev) : 0;

  // ---- A, first loads: thread = voxel row (TMEM lane 32 wq + lane); the two warps of a lane quadrant split the channel
  // groups.  Per group of 8 channels 8 loads: for a fixed channel the 32 lanes read 128 contiguous bytes.
  const int v = v0 + wq * 32 + lane;
  const bool vok = v < S;
  const int ngroups = Kpad >> 3;
  const int gh = (ngroups + 1) >> 1, gbeg = wp * gh, gend = min(ngroups, gbeg + gh);   // this half's channel groups
  float xn[8];                                                // the group loaded one iteration ahead
  auto load_group = [&](int g) {
    const int c0 = g * 8;
    // the 8 channels of a group come from one source (source widths are multiples of 8 on this path)
    int c = c0, sidx = 0;
    while (sidx < P.nsrc - 1 && c >= P.src[sidx].C) { c -= P.src[sidx].C; ++sidx; }
    const float* xcol = P.src[sidx].ptr + ((size_t)b * P.src[sidx].C + c) * S + v;
    const int nlive = (g < gend && vok) ? (Ci - c0 < 8 ? Ci - c0 : 8) : 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) xn[j] = j < nlive ? __ldg(xcol + (size_t)j * S) : 0.f;
  };
  load_group(gbeg);

  // ---- B: weights (n, k) -> K-major core matrices.  16-byte loads along the contiguous axis of W (rows are 16-byte
  // aligned and 4-element granular on this path, checked by pw_tc_forward), two per thread in flight.
  {
    const int Q = P.transposed ? (Npad >> 2) : (Kpad >> 2);      // float4 per row of the operand as stored
    const int total = (P.transposed ? Kpad : Npad) * Q;
    constexpr int WU = 2;
#pragma unroll 1
    for (int base = tid; base < total; base += WU * TC_THREADS) {
      float4 w[WU];
      int row[WU], q4[WU];
#pragma unroll
      for (int u = 0; u < WU; ++u) {
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
      for (int u = 0; u < WU; ++u) {
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
    // bias per output column and the prologue's per-channel affine: read once per CTA, served from shared memory later
    for (int n = tid; n < Npad; n += TC_THREADS) {
      float bv = 0.f;
      if (!P.transposed && n < Co) {
        int seg = 0, seg_off = 0;
        while (seg < P.nseg - 1 && n >= seg_off + P.seg[seg].n) { seg_off += P.seg[seg].n; ++seg; }
        if (P.seg[seg].bias) bv = __ldg(P.seg[seg].bias + (n - seg_off));
      }
      bias_s[n] = bv;
    }
    for (int k = tid; k < Kpad; k += TC_THREADS) {
      const bool aff = P.pro == PRO_AFFINE && k < Ci;
      const int q = b * P.pro_bstride + k;
      proa_s[k] = aff ? __ldg(P.pro_a + q) : 1.f;
      proc_s[k] = aff ? __ldg(P.pro_c + q) : 0.f;
    }
  }
  // shared-memory weight image -> visible to the tensor-core (async) proxy; TMEM base address -> visible to all threads
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t trow = tmem + ((uint32_t)(wq * 32) << 16);

  // ---- A: prologue, hi / lo split, two tcgen05.st of 8 columns per channel group; the next group's loads are in flight
  {
    const uint32_t a_hi = trow + (uint32_t)Npad, a_lo = a_hi + (uint32_t)Kpad;
#pragma unroll 1
    for (int g = gbeg; g < gend; ++g) {
      float xx[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) xx[j] = xn[j];
      if (g + 1 < gend) load_group(g + 1);
      const int c0 = g * 8;
      if (pro_gelu) {
#pragma unroll 1
        for (int h2 = 0; h2 < 8; h2 += 4) {
          float4 gq = h2 ? make_float4(xx[4], xx[5], xx[6], xx[7]) : make_float4(xx[0], xx[1], xx[2], xx[3]);
          gq = gelu4_call(gq);
          if (h2) { xx[4] = gq.x; xx[5] = gq.y; xx[6] = gq.z; xx[7] = gq.w; } else { xx[0] = gq.x; xx[1] = gq.y; xx[2] = gq.z; xx[3] = gq.w; }
        }
      }
      if (pro_drop) {
        float keep[8];
        drop_keep8(P.pro_seed + soff, P.pro_site, ((size_t)b * Ci + c0) * (size_t)S + v, (size_t)S, lane, P.pro_drop_p, pinv, keep);
#pragma unroll
        for (int j = 0; j < 8; ++j) xx[j] *= keep[j];
      }
      float h[8], l[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        // scale / shift are 1 / 0 unless PRO_AFFINE; rows past the tensor stay zero (the affine shift is not)
        const float t = vok ? fmaf(xx[j], proa_s[c0 + j], proc_s[c0 + j]) : 0.f;
        if (batch.prec) { h[j] = bf16_round(t); l[j] = 0.f; } else split_tf32(t, h[j], l[j]);
      }
      tmem_st8(a_hi + (uint32_t)c0, h);
      if (!batch.prec) tmem_st8(a_lo + (uint32_t)c0, l);
    }
  }

  // ---- epilogue operands: thread = voxel row, column batches of 8 output channels c0 = wp * 8 + 16 i.  The reads of the
  // first batch (GELU' argument, residuals) are issued here, before the MMAs are even launched.
  const int nbt = Npad > wp * 8 ? (Npad - wp * 8 + 15) >> 4 : 0;       // this thread's column batches
  float e1n[8], e2n[8];                                                // GELU' argument; res_scale * res + res2 (one batch ahead)
  auto epi_load = [&](int i) {
    const int c0 = wp * 8 + 16 * i;
    const int nlive = (i < nbt && vok) ? (Co - c0 < 8 ? Co - c0 : 8) : 0;
    const size_t lbase = ((size_t)b * Co + c0) * S + v;              // index in the logical (B, Co, S) tensor
#pragma unroll
    for (int j = 0; j < 8; ++j) { e1n[j] = 0.f; e2n[j] = 0.f; }
    if (P.mulgrad) {
      const float* q = P.mulgrad + lbase;
#pragma unroll
      for (int j = 0; j < 8; ++j) if (j < nlive) e1n[j] = __ldg(q + (size_t)j * S);
    }
    if (P.res) {
      const float* q = P.res + lbase;
#pragma unroll
      for (int j = 0; j < 8; ++j) if (j < nlive) e2n[j] = P.res_scale * __ldg(q + (size_t)j * S);
    }
    if (P.res2) {
      const float* q = P.res2 + lbase;
#pragma unroll
      for (int j = 0; j < 8; ++j) if (j < nlive) e2n[j] += __ldg(q + (size_t)j * S);
    }
  };
  epi_load(0);

  // the A tile is complete in tensor memory once every warp has waited for its stores
  tmem_wait_st();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (tid == 0) {
    const uint32_t idesc = umma_idesc_tf32(Npad);
    const uint32_t b_hi = smem_u32(B_hi), b_lo = smem_u32(B_lo);
    const uint32_t ta_hi = tmem + (uint32_t)Npad, ta_lo = ta_hi + (uint32_t)Kpad;      // lane 0 of the CTA's columns
    const uint32_t b_lbo = 128u, b_sbo = 256u;
#pragma unroll 1
    for (int g8 = 0; g8 < ngroups; ++g8) {
      const uint32_t bo = (uint32_t)g8 * (uint32_t)(Npad * 8 * 4);
      const uint64_t dbh = umma_desc(b_hi + bo, b_lbo, b_sbo), dbl = umma_desc(b_lo + bo, b_lbo, b_sbo);
      const uint32_t ah = ta_hi + (uint32_t)(g8 * 8), al = ta_lo + (uint32_t)(g8 * 8);
      if (batch.prec) {                               // bf16 numerics: operands are exact in tf32, one product
        umma_tf32_ts(tmem, ah, dbh, idesc, g8 > 0 ? 1u : 0u);
      } else {
        umma_tf32_ts(tmem, al, dbh, idesc, g8 > 0 ? 1u : 0u);
        umma_tf32_ts(tmem, ah, dbl, idesc, 1u);
        umma_tf32_ts(tmem, ah, dbh, idesc, 1u);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar))
                 : "memory");
  }
  mbar_wait(smem_u32(&mbar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- epilogue: one column batch per iteration (tcgen05.ld of 8 columns), the next batch's operands in flight
  {
    const float dinv = P.drop_p > 0.f ? 1.0f / (1.0f - P.drop_p) : 1.f;
#pragma unroll 1
    for (int i = 0; i < nbt; ++i) {
      const int c0 = wp * 8 + 16 * i;
      uint32_t r[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(trow + (uint32_t)c0)
                   : "memory");
      float e1[8], e2[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { e1[j] = e1n[j]; e2[j] = e2n[j]; }
      if (i + 1 < nbt) epi_load(i + 1);
      // the 8 channels of a batch share one output segment (segment sizes are multiples of 8 on this path)
      float* obase;
      if (!P.transposed) {
        int seg = 0, seg_off = 0;
        while (seg < P.nseg - 1 && c0 >= seg_off + P.seg[seg].n) { seg_off += P.seg[seg].n; ++seg; }
        obase = P.seg[seg].out + ((size_t)b * P.seg[seg].n + (c0 - seg_off)) * S + v;
      } else {
        obase = P.seg[0].out + ((size_t)b * Co + c0) * S + v;
      }
      const int nlive = vok ? (Co - c0 < 8 ? Co - c0 : 8) : 0;
      tmem_wait_ld();
      float y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = __uint_as_float(r[j]) + bias_s[c0 + j];
      if (P.act == 1 || P.mulgrad) {          // out-of-line 4-wide copies, called from a rolled loop: one call site each
#pragma unroll 1
        for (int h2 = 0; h2 < 8; h2 += 4) {
          float4 yq = h2 ? make_float4(y[4], y[5], y[6], y[7]) : make_float4(y[0], y[1], y[2], y[3]);
          if (P.act == 1) yq = gelu4_call(yq);
          if (P.mulgrad) {
            const float4 gq = gelu_grad4_call(h2 ? make_float4(e1[4], e1[5], e1[6], e1[7]) : make_float4(e1[0], e1[1], e1[2], e1[3]));
            yq.x *= gq.x; yq.y *= gq.y; yq.z *= gq.z; yq.w *= gq.w;
          }
          if (h2) { y[4] = yq.x; y[5] = yq.y; y[6] = yq.z; y[7] = yq.w; } else { y[0] = yq.x; y[1] = yq.y; y[2] = yq.z; y[3] = yq.w; }
        }
      }
      if (P.drop_p > 0.f) {                                   // CTA-uniform branch; warp-collective inside
        float keep[8];
        drop_keep8(P.seed + soff, P.site, ((size_t)b * Co + c0) * (size_t)S + v, (size_t)S, lane, P.drop_p, dinv, keep);
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] *= keep[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] += e2[j];              // res_scale * res + res2 (zero without those streams)
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
    for (int s = 0; s < P.nsrc; ++s) {
      if (((uintptr_t)P.src[s].ptr & 3) || P.src[s].C <= 0) return 1;
      if (s < P.nsrc - 1 && (P.src[s].C & 7)) return 1;       // a group of 8 input channels never straddles two sources
    }
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
  if (shp.Npad > 256 || shp.Npad + 2 * shp.Kpad > 512) return 1;       // D | A_hi | A_lo share the CTA's TMEM columns
  const size_t smem = sizeof(float) * (2 * (size_t)shp.Kpad * shp.Npad + shp.Npad + 2 * shp.Kpad);
  if (smem > 200 * 1024) return 1;
  shp.tmem_cols = 32;
  while (shp.tmem_cols < shp.Npad + 2 * shp.Kpad) shp.tmem_cols <<= 1;
  PwBatch launch = batch;
  launch.seed_dev = get_seed_dev();
  launch.prec = precision_mode();
  dim3 grid(cdiv(S, TC_M), 1, batch.nprob * batch.B);
  VX_SET_SMEM(pw_tc_kernel, smem);
  VX_LAUNCH(pw_tc_kernel, grid, dim3(TC_THREADS), smem, stream, launch, shp);
  return check_launch("pw_tc_kernel");
}

}  // namespace vx

#endif  // VX_EMU
