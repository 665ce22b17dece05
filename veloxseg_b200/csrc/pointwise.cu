// Channel contractions (1x1x1 convs) and the norm kernels they are paired with.  fp32, NCDHW.
//
// pw_kernel    : voxel-parallel.  One CTA = 256 consecutive voxels x 16 output channels; a thread owns two voxels
//                and 16 accumulators each, weights are broadcast from shared memory as float4.  HBM side: every
//                load/store is a 128-byte-coalesced run along the voxel axis.
// pw_wgrad     : split-K over voxel chunks; chunk tiles are transposed through shared memory and reduced as
//                4x4 register outer products, partial results are folded with fp32 atomics.
#include <cstdlib>
#include "vx_kernels.h"

#ifdef VX_EMU
#define __grid_constant__
#endif

namespace vx {

constexpr int PW_THREADS = 128;
constexpr int PW_TS = 64;    // voxels per CTA tile
constexpr int PW_TC = 64;    // output channels per CTA tile
constexpr int PW_KC = 16;    // input channels per staged chunk

VX_DEV float pw_prologue(const PwProblem& P, float x, int b, int cg, int v, int S, float pinv, uint64_t soff) {
  if (P.pro == PRO_AFFINE) {
    const int k = b * P.pro_bstride + cg;
    return fmaf(x, __ldg(P.pro_a + k), __ldg(P.pro_c + k));
  }
  if (P.pro == PRO_NONE) return x;
  return pw_pro_heavy(P.pro, x, P.pro_seed + soff, P.pro_site, ((uint64_t)b * P.Ci + cg) * (uint64_t)S + v, P.pro_drop_p, pinv);
}

// Y[b, co, v] = epi( sum_ci W[co, ci] * pro(X[b, ci, v]) + bias[co] ) as a register-tiled GEMM: CTA tile = 64 voxels x
// 64 output channels, K staged 16 channels at a time through double-buffered shared memory with cp.async (the copies
// of chunk k+1 fly while chunk k is multiplied); a thread owns 4 voxels x 8 channels (32 accumulators, 3 LDS.128 per
// 32 FMA).  Warp = 16 voxel groups x 2 channel groups: X reads are two 128-B wavefronts, W reads are broadcasts.
// Code size is a first-class constraint here (a kernel runs once per thread, so SASS beyond the 32 KB instruction
// cache is fetch-bound): staging and epilogue are ROLLED loops with a single copy of the address / prologue /
// epilogue logic, and the epilogue reads the accumulator tile back from shared memory.
__global__ void __launch_bounds__(PW_THREADS) pw_kernel(const __grid_constant__ PwBatch batch) {
  VX_PDL_ENTRY();
  const int pi = blockIdx.z / batch.B, b = blockIdx.z % batch.B;
  const PwProblem& P = batch.p[pi];
  const int S = batch.S, Ci = P.Ci, Co = P.Co;
  const int co0 = blockIdx.y * PW_TC, v0 = blockIdx.x * PW_TS;
  if (co0 >= Co) return;
  __align__(16) __shared__ float smem[2 * PW_KC * PW_TS + 2 * PW_KC * PW_TC];
  float (*Xs)[PW_KC][PW_TS] = reinterpret_cast<float (*)[PW_KC][PW_TS]>(smem);
  float (*Ws)[PW_KC][PW_TC] = reinterpret_cast<float (*)[PW_KC][PW_TC]>(smem + 2 * PW_KC * PW_TS);
  const int tid = threadIdx.x;
  const int vg = tid & 15, cgp = tid >> 4;            // voxel group (4 voxels), channel group (8 channels)
  const bool pro_drop = P.pro == PRO_DROPOUT || P.pro == PRO_GELU_DROPOUT;
  const float pinv = pro_drop ? 1.0f / (1.0f - P.pro_drop_p) : 1.f;
  const uint64_t soff = batch.seed_dev ? (uint64_t)__ldg(batch.seed_dev) : 0;

  float acc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;

  // staging maps: X chunk 16 x 64 -> 8 elements per thread (voxel fastest: coalesced), W chunk 16 x 64 likewise
  auto stage = [&](int buf, int k0) {
#pragma unroll 1
    for (int r = 0; r < 8; ++r) {
      const int e = r * PW_THREADS + tid;
      const int v = e & (PW_TS - 1), k = e >> 6;
      const bool ok = k0 + k < Ci && v0 + v < S;
      vx_cp_async4(&Xs[buf][k][v], ok ? pw_x_ptr(P, b, k0 + k, v0 + v, S) : P.src[0].ptr, ok);
    }
#pragma unroll 1
    for (int r = 0; r < 8; ++r) {
      const int e = r * PW_THREADS + tid;
      // forward orientation reads W[co][ci] (ci contiguous): k fastest; transposed reads W[ci][co]: co fastest
      const int k = P.transposed ? (e >> 6) : (e & (PW_KC - 1));
      const int c = P.transposed ? (e & (PW_TC - 1)) : (e >> 4);
      const bool ok = k0 + k < Ci && co0 + c < Co;
      vx_cp_async4(&Ws[buf][k][c], ok ? pw_w_ptr(P, co0 + c, k0 + k) : P.seg[0].W, ok);
    }
    vx_cp_async_commit();
  };
  auto prologue = [&](int buf, int k0) {     // in place, on the elements this thread staged
    if (P.pro == PRO_NONE) return;
#pragma unroll 1
    for (int r = 0; r < 8; ++r) {
      const int e = r * PW_THREADS + tid;
      const int v = e & (PW_TS - 1), k = e >> 6;
      if (k0 + k < Ci && v0 + v < S) Xs[buf][k][v] = pw_prologue(P, Xs[buf][k][v], b, k0 + k, v0 + v, S, pinv, soff);
    }
  };

  const int nchunk = (Ci + PW_KC - 1) / PW_KC;
  stage(0, 0);
  vx_cp_async_wait_all();
  prologue(0, 0);
  __syncthreads();
  for (int ch = 0; ch < nchunk; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < nchunk) stage(buf ^ 1, (ch + 1) * PW_KC);
#pragma unroll
    for (int k = 0; k < PW_KC; ++k) {
      const float4 x = *reinterpret_cast<const float4*>(&Xs[buf][k][vg * 4]);
      const float4 wa = *reinterpret_cast<const float4*>(&Ws[buf][k][cgp * 8]);
      const float4 wb = *reinterpret_cast<const float4*>(&Ws[buf][k][cgp * 8 + 4]);
      const float xs[4] = {x.x, x.y, x.z, x.w};
      const float ws[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(ws[j], xs[i], acc[j][i]);
    }
    if (ch + 1 < nchunk) {
      vx_cp_async_wait_all();
      prologue(buf ^ 1, (ch + 1) * PW_KC);
    }
    __syncthreads();
  }

  // accumulator tile -> shared memory [co][v] (the staging buffers are free now), then a rolled epilogue
  float (*Cs)[PW_TS] = reinterpret_cast<float (*)[PW_TS]>(smem);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(&Cs[cgp * 8 + j][vg * 4]) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
  __syncthreads();
  const float dinv = P.drop_p > 0.f ? 1.0f / (1.0f - P.drop_p) : 1.f;
#pragma unroll 1
  for (int r = 0; r < (PW_TC * PW_TS / 4) / PW_THREADS; ++r) {
    const int it = r * PW_THREADS + tid;
    const int col = it >> 4, vq = it & 15;
    const int co = co0 + col, vb = v0 + vq * 4;
    if (co >= Co || vb >= S) continue;
    float* outp;
    float bias = 0.f;
    if (!P.transposed) {
      int seg = 0, seg_off = 0;
      while (co >= seg_off + P.seg[seg].n) { seg_off += P.seg[seg].n; ++seg; }
      outp = P.seg[seg].out + ((size_t)b * P.seg[seg].n + (co - seg_off)) * S;
      if (P.seg[seg].bias) bias = __ldg(P.seg[seg].bias + co - seg_off);
    } else {
      outp = P.seg[0].out + ((size_t)b * Co + co) * S;
    }
    const size_t lbase = ((size_t)b * Co + co) * S;   // index in the logical (B, Co, S) tensor
    const float4 a4 = *reinterpret_cast<const float4*>(&Cs[col][vq * 4]);
    float y[4] = {a4.x + bias, a4.y + bias, a4.z + bias, a4.w + bias};
    if (P.act == 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) y[i] = gelu_f(y[i]);
    }
    if (P.mulgrad) {
#pragma unroll
      for (int i = 0; i < 4; ++i) y[i] *= (vb + i < S) ? gelu_grad_f(__ldg(P.mulgrad + lbase + vb + i)) : 0.f;
    }
    if (P.drop_p > 0.f) {
      float ms[4];
      dropout_scale4(P.seed + soff, P.site, lbase + vb, P.drop_p, dinv, ms);
#pragma unroll
      for (int i = 0; i < 4; ++i) y[i] *= ms[i];
    }
    if (P.res) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (vb + i < S) y[i] = fmaf(P.res_scale, __ldg(P.res + lbase + vb + i), y[i]);
    }
    if (P.res2) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (vb + i < S) y[i] += __ldg(P.res2 + lbase + vb + i);
    }
    if (((S & 3) == 0) && (vb + 3 < S)) {
      *reinterpret_cast<float4*>(outp + vb) = make_float4(y[0], y[1], y[2], y[3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (vb + i < S) outp[vb + i] = y[i];
    }
  }
}

// Small-voxel variant (levels 3-4: S = 216 / 27).  pw_kernel's 64 x 64 CTA tile leaves 8-32 CTAs with one warp per
// scheduler walking a serial K loop -- pure instruction latency (85 us for 3.5 MFLOP at level 4).  Here the batch and
// voxel axes are flattened (f = b * S + v) and a warp owns 32 flattened voxels x 4 output channels with the whole K loop
// in registers: 4x-16x more warps, no shared memory, no barriers.  X reads are coalesced along the voxel axis, weight
// reads are warp-uniform 16-byte loads (the four rows co..co+3, four k at a time) served by L1.
//
// The K axis is additionally split over the KS = blockDim.y warps of a CTA (8-channel steps dealt round-robin), so that
// a level-4 problem (K = 256..512) is 8 short latency chains instead of one long one; partial sums meet in shared memory.
constexpr int PWS_MAX_KS = 8;

__global__ void __launch_bounds__(32 * PWS_MAX_KS) pw_small_kernel(const __grid_constant__ PwBatch batch) {
  VX_PDL_ENTRY();
  const PwProblem& P = batch.p[blockIdx.z];
  const int S = batch.S, Co = P.Co;
  const int lane = threadIdx.x, ks = threadIdx.y, KS = blockDim.y;
  const int f = blockIdx.x * 32 + lane;
  const int co0 = blockIdx.y * 4;
  __shared__ float part[PWS_MAX_KS][4][32];
  if (co0 >= Co) return;
  const bool ok = f < batch.B * S;
  const int b = ok ? f / S : 0, v = ok ? f % S : 0;
  const bool pro_drop = P.pro == PRO_DROPOUT || P.pro == PRO_GELU_DROPOUT;
  const float pinv = pro_drop ? 1.0f / (1.0f - P.pro_drop_p) : 1.f;
  const uint64_t soff = batch.seed_dev ? (uint64_t)__ldg(batch.seed_dev) : 0;
  const bool quad = co0 + 3 < Co;

  // forward orientation: the four weight rows of this warp (a row never crosses an output segment)
  const float* rp[4] = {nullptr, nullptr, nullptr, nullptr};
  // fewer than 4 live rows (the 2-class heads): the missing rows alias the last live one, their accumulators are never stored
  bool rows_vec = !P.transposed;
  if (!P.transposed) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int ld;
      rp[j] = pw_w_row(P, co0 + j < Co ? co0 + j : Co - 1, 0, ld);
      rows_vec = rows_vec && (((uintptr_t)rp[j] & 15) == 0) && ((ld & 3) == 0);
    }
  }

  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int cg0 = 0;                                            // channel index of the source's first channel in the concatenated input
  // One 8-channel block = 8 coalesced X loads + 8 warp-uniform 16-byte weight loads, then 32 FMAs.  The kernel is a latency
  // chain (one or two warps per scheduler), so a warp takes its blocks (ks, ks + KS, ...) two at a time: the 32 loads of both
  // are in flight before the first FMA -- half as many dependent memory round trips as one block per iteration (the K = 256-512
  // contractions of level 4 were 8 round trips per warp, 15-22 us per launch for 7-14 MFLOP; profiles/r3g kernel table).
  auto load_block = [&](const float* xp, int c, int cg, const float* w0, int ld, float (&x)[8], float4 (&w)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = ok ? __ldg(xp + (size_t)(c + i) * S) : 0.f;
    if (!P.transposed) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {      // w[2j], w[2j+1] = row j, k 0..3 and 4..7
        w[2 * j] = __ldg(reinterpret_cast<const float4*>(rp[j] + cg));
        w[2 * j + 1] = __ldg(reinterpret_cast<const float4*>(rp[j] + cg + 4));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = __ldg(reinterpret_cast<const float4*>(w0 + (size_t)i * ld));   // row k = i
    }
  };
  auto fma_block = [&](int cg, float (&x)[8], const float4 (&w)[8]) {
    if (P.pro == PRO_AFFINE) {
      const int q = b * P.pro_bstride + cg;
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], __ldg(P.pro_a + q + i), __ldg(P.pro_c + q + i));
    } else if (P.pro != PRO_NONE) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        x[i] = pw_pro_heavy(P.pro, x[i], P.pro_seed + soff, P.pro_site, ((uint64_t)b * P.Ci + cg + i) * (uint64_t)S + v,
                            P.pro_drop_p, pinv);
    }
    if (!P.transposed) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 a = w[2 * j], q = w[2 * j + 1];
        acc[j] = fmaf(a.x, x[0], acc[j]); acc[j] = fmaf(a.y, x[1], acc[j]); acc[j] = fmaf(a.z, x[2], acc[j]);
        acc[j] = fmaf(a.w, x[3], acc[j]); acc[j] = fmaf(q.x, x[4], acc[j]); acc[j] = fmaf(q.y, x[5], acc[j]);
        acc[j] = fmaf(q.z, x[6], acc[j]); acc[j] = fmaf(q.w, x[7], acc[j]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[0] = fmaf(w[i].x, x[i], acc[0]); acc[1] = fmaf(w[i].y, x[i], acc[1]);
        acc[2] = fmaf(w[i].z, x[i], acc[2]); acc[3] = fmaf(w[i].w, x[i], acc[3]);
      }
    }
  };
  // weight rows of block `blk` of the current source are vector-loadable (uniform over the CTA)
  auto block_vec = [&](int cg, const float*& w0, int& ld) -> bool {
    if (!P.transposed) { w0 = nullptr; ld = 0; return rows_vec && (cg & 3) == 0; }
    int ld7;
    w0 = pw_w_row(P, co0, cg, ld);
    return quad && (((uintptr_t)w0 & 15) == 0) && ((ld & 3) == 0) && pw_w_row(P, co0, cg + 7, ld7) == w0 + (size_t)7 * ld;
  };
#pragma unroll 1
  for (int s = 0; s < P.nsrc; ++s) {
    const int Cs = P.src[s].C;
    const float* xp = P.src[s].ptr + (size_t)b * Cs * S + v;
    // vector part: leading 8-channel blocks whose weights are vector-loadable (the first block that is not ends it for everyone)
    int nvec = 0;
    {
      const float* w0; int ld;
      while ((nvec + 1) * 8 <= Cs && block_vec(cg0 + nvec * 8, w0, ld)) ++nvec;
    }
#pragma unroll 1
    for (int blk = ks; blk < nvec; blk += 2 * KS) {
      const int blk2 = blk + KS;
      const bool two = blk2 < nvec;
      float xa[8], xb[8];
      float4 wa[8], wb[8];
      const float* w0a; const float* w0b = nullptr;
      int lda, ldb = 0;
      block_vec(cg0 + blk * 8, w0a, lda);
      load_block(xp, blk * 8, cg0 + blk * 8, w0a, lda, xa, wa);
      if (two) {
        block_vec(cg0 + blk2 * 8, w0b, ldb);
        load_block(xp, blk2 * 8, cg0 + blk2 * 8, w0b, ldb, xb, wb);
      }
      fma_block(cg0 + blk * 8, xa, wa);
      if (two) fma_block(cg0 + blk2 * 8, xb, wb);
    }
    // scalar remainder of this source, dealt to the warps channel by channel
#pragma unroll 1
    for (int c = nvec * 8 + ks; c < Cs; c += KS) {
      const int cg = cg0 + c;
      float x = ok ? __ldg(xp + (size_t)c * S) : 0.f;
      if (P.pro != PRO_NONE) x = pw_prologue(P, x, b, cg, v, S, pinv, soff);
#pragma unroll
      for (int j = 0; j < 4; ++j) if (co0 + j < Co) acc[j] = fmaf(__ldg(pw_w_ptr(P, co0 + j, cg)), x, acc[j]);
    }
    cg0 += Cs;
  }
  if (KS > 1) {
#pragma unroll
    for (int j = 0; j < 4; ++j) part[ks][j][lane] = acc[j];
    __syncthreads();
    if (ks != 0) return;
#pragma unroll 1
    for (int k2 = 1; k2 < KS; ++k2)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] += part[k2][j][lane];
  }
  if (!ok) return;
  // epilogue: gather every input first, then compute, then store
  const float dinv = P.drop_p > 0.f ? 1.0f / (1.0f - P.drop_p) : 1.f;
  const bool heavy = P.act == 1 || P.mulgrad || P.drop_p > 0.f;
  float* outp[4];
  size_t li[4];
  float bias[4], mg[4], r1[4], r2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int co = co0 + j < Co ? co0 + j : co0;
    bias[j] = 0.f;
    if (!P.transposed) {
      int seg = 0, seg_off = 0;
      while (co >= seg_off + P.seg[seg].n) { seg_off += P.seg[seg].n; ++seg; }
      outp[j] = P.seg[seg].out + ((size_t)b * P.seg[seg].n + (co - seg_off)) * S + v;
      if (P.seg[seg].bias) bias[j] = __ldg(P.seg[seg].bias + co - seg_off);
    } else {
      outp[j] = P.seg[0].out + ((size_t)b * Co + co) * S + v;
    }
    li[j] = ((size_t)b * Co + co) * S + v;
    mg[j] = P.mulgrad ? __ldg(P.mulgrad + li[j]) : 0.f;
    r1[j] = P.res ? __ldg(P.res + li[j]) : 0.f;
    r2[j] = P.res2 ? __ldg(P.res2 + li[j]) : 0.f;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (co0 + j >= Co) break;
    float y = acc[j] + bias[j];
    if (heavy) y = pw_epi_heavy(y, P.act, P.mulgrad != nullptr, mg[j], P.drop_p, P.seed + soff, P.site, li[j], dinv);
    if (P.res) y = fmaf(P.res_scale, r1[j], y);
    if (P.res2) y += r2[j];
    *outp[j] = y;
  }
}

static int g_small_max_s = 512, g_tc_min_s = 1024;
void pw_set_thresholds(int small_max_s, int tc_min_s) {
  if (small_max_s >= 0) g_small_max_s = small_max_s;
  if (tc_min_s >= 0) g_tc_min_s = tc_min_s;
}

int pw_forward(const PwBatch& batch, cudaStream_t stream) {
  int maxCo = 0;
  double bytes = 0.0, flops = 0.0;
  for (int i = 0; i < batch.nprob; ++i) {
    const PwProblem& P = batch.p[i];
    maxCo = P.Co > maxCo ? P.Co : maxCo;
    int cs = 0;
    for (int s = 0; s < P.nsrc; ++s) cs += P.src[s].C;
    if (cs != P.Ci) { set_error("pw_forward: source channels %d != Ci %d", cs, P.Ci); return VX_ERR_BAD_DESC; }
    if (!P.transposed && P.nseg > 1 && (P.res || P.mulgrad || P.drop_p > 0.f || P.res2)) {
      set_error("pw_forward: epilogue tensors need a single output segment"); return VX_ERR_BAD_DESC;
    }
    // algorithmic bytes: X in, Y out, epilogue tensors in, weights
    bytes += 4.0 * batch.B * batch.S * (P.Ci + P.Co * (1.0 + (P.res ? 1 : 0) + (P.res2 ? 1 : 0) + (P.mulgrad ? 1 : 0))) +
             4.0 * P.Ci * P.Co;
    flops += 2.0 * batch.B * batch.S * P.Ci * P.Co;
  }
  if (batch.nprob <= 0 || batch.B <= 0 || batch.S <= 0) return VX_OK;
#ifndef VX_EMU
  // tcgen05 kernel from 1024 voxels per sample (measured at the level-2 shapes, B = 4: JLC 129 -> 115 us, PWA block
  // 184 -> 157 us forward); concatenated multi-source inputs (the modal mixer: one problem, 56 tiles) stay on the SIMT
  // kernel below 4096 voxels (27 vs 35 us).
  bool multi_src = false;
  for (int i = 0; i < batch.nprob; ++i) multi_src = multi_src || batch.p[i].nsrc > 1;
  if (batch.S >= g_tc_min_s && batch.S >= g_small_max_s && !(multi_src && batch.S < 4096)) {
    prof_bytes(bytes);
    prof_flops(flops);
    const int rc = pw_tc_forward(batch, stream);
    if (rc <= 0) return rc;
  }
#endif
  PwBatch launch = batch;
  launch.seed_dev = get_seed_dev();
  prof_bytes(bytes);
  prof_flops(flops);
  if (batch.S < g_small_max_s) {
    int maxCi = 0;
    for (int i = 0; i < batch.nprob; ++i) maxCi = batch.p[i].Ci > maxCi ? batch.p[i].Ci : maxCi;
    int KS = 1;
    while (KS < PWS_MAX_KS && maxCi / (8 * KS * 2) >= 2) KS *= 2;      // >= 2 eight-channel steps per warp
    dim3 grid(cdiv((long long)batch.B * batch.S, 32), cdiv(maxCo, 4), batch.nprob);
    VX_LAUNCH(pw_small_kernel, grid, dim3(32, KS), 0, stream, launch);
    return check_launch("pw_small_kernel");
  }
  dim3 grid(cdiv(batch.S, PW_TS), cdiv(maxCo, PW_TC), batch.nprob * batch.B);
  VX_LAUNCH(pw_kernel, grid, dim3(PW_THREADS), 0, stream, launch);
  return check_launch("pw_kernel");
}

// ---------------------------------------------------------------------------------------------------
// Fused two-layer MLP of the small levels (S < 512: levels 3-4).  CTA = FF_VT consecutive flattened voxels (f = b S + v) x
// every channel; 256 threads.  Phase 1 stages the prologued input tile [C][FF_VT]; phase 2: thread = hidden channel (x a
// share of the tile's voxels when eC < 256), whole K loop in registers, weights as 16-byte loads along the row; hpre goes to
// global memory, GELU(hpre) (x mask) to shared memory; phase 3: thread = (output channel, slice of the hidden channels), the
// slices meet in shared memory; epilogue: bias, dropout, residual.  One launch and no global round trip for the hidden
// activation instead of two latency-bound pw_small_kernel launches (12-16 us each for 7-28 MFLOP).
// ---------------------------------------------------------------------------------------------------
constexpr int FF_VT = 8, FF_THREADS = 256, FF_MAX_C = 128, FF_MAX_E = 256;

// Row walk of the forward phases: acc[i] += W[row][c] * v[c][i0 + i] over c in [c0, c1) (c0 a multiple of 4), 8 16-byte weight
// loads (32 columns) in flight.  NV (voxels per thread) is a template parameter so that no accumulator slot is predicated.
template <int NV>
VX_DEV void ffn_row_walk(const float* __restrict__ wrow, int c0, int c1, const float (*v)[FF_VT], int i0, float (&acc)[NV]) {
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;
  int c = c0;
#pragma unroll 1
  for (; c + 32 <= c1; c += 32) {
    float4 w[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) w[u] = __ldg(reinterpret_cast<const float4*>(wrow + c + 4 * u));
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float ww[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] = fmaf(ww[k], v[c + 4 * u + k][i0 + i], acc[i]);
    }
  }
#pragma unroll 1
  for (; c + 4 <= c1; c += 4) {
    const float4 w = __ldg(reinterpret_cast<const float4*>(wrow + c));
    const float ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int i = 0; i < NV; ++i) acc[i] = fmaf(ww[k], v[c + k][i0 + i], acc[i]);
  }
  for (; c < c1; ++c) {
    const float w = __ldg(wrow + c);
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = fmaf(w, v[c][i0 + i], acc[i]);
  }
}

template <int NV>
VX_DEV void ffn_fwd_hidden(const FfnProblem& P, int S, uint64_t soff, const float (*xs)[FF_VT], float (*hs)[FF_VT], const int* vb,
                           const int* vv, int j, int i0) {
  float acc[NV];
  ffn_row_walk<NV>(P.W1 + (size_t)j * P.C, 0, P.C, xs, i0, acc);
  const float bj = P.b1 ? __ldg(P.b1 + j) : 0.f;
  const float minv = P.mid_drop_p > 0.f ? 1.0f / (1.0f - P.mid_drop_p) : 1.f;
  const uint32_t mkey = P.mid_drop_p > 0.f ? rng_key(P.mid_seed + soff, P.mid_site) : 0u;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int ii = i0 + i;
    float h = 0.f;
    if (vb[ii] >= 0) {
      const float pre = acc[i] + bj;
      const size_t idx = ((size_t)vb[ii] * P.eC + j) * S + vv[ii];
      P.hpre[idx] = pre;
      h = gelu_f(pre);
      if (P.mid_drop_p > 0.f) h *= keep_from_bits(rng_word(mkey, idx), P.mid_drop_p, minv);
    }
    hs[j][ii] = h;
  }
}

__global__ void __launch_bounds__(FF_THREADS, 2) pw_ffn_small_kernel(const __grid_constant__ FfnBatch batch) {
  VX_PDL_ENTRY();
  const FfnProblem& P = batch.p[blockIdx.y];
  const int S = batch.S, C = P.C, eC = P.eC, tid = threadIdx.x;
  const int f0 = blockIdx.x * FF_VT, total = batch.B * S;
  __align__(16) __shared__ float xs[FF_MAX_C][FF_VT];
  __align__(16) __shared__ float hs[FF_MAX_E][FF_VT];
  __align__(16) __shared__ float red[FF_THREADS][FF_VT];
  __shared__ int vb[FF_VT], vv[FF_VT];
  const uint64_t soff = batch.seed_dev ? (uint64_t)__ldg(batch.seed_dev) : 0;
  if (tid < FF_VT) {
    const int f = f0 + tid;
    vb[tid] = f < total ? f / S : -1;
    vv[tid] = f < total ? f % S : 0;
  }
  __syncthreads();
  // ---- phase 1: xs[c][i] = a x + c
  for (int e = tid; e < C * FF_VT; e += FF_THREADS) {
    const int c = e / FF_VT, i = e % FF_VT;
    float v = 0.f;
    if (vb[i] >= 0) {
      v = __ldg(P.x + ((size_t)vb[i] * C + c) * S + vv[i]);
      if (P.pro_a) { const int q = vb[i] * P.pro_bstride + c; v = fmaf(v, __ldg(P.pro_a + q), __ldg(P.pro_c + q)); }
    }
    xs[c][i] = v;
  }
  __syncthreads();
  // ---- phase 2: hidden channel j, voxel share [i0, i0 + nv)
  {
    const int nsp = eC <= 64 ? 4 : (eC <= 128 ? 2 : 1);       // threads per hidden channel
    const int j = tid / nsp, sh = tid % nsp;
    if (j < eC) {
      switch (nsp) {
        case 4: ffn_fwd_hidden<2>(P, S, soff, xs, hs, vb, vv, j, sh * 2); break;
        case 2: ffn_fwd_hidden<4>(P, S, soff, xs, hs, vb, vv, j, sh * 4); break;
        default: ffn_fwd_hidden<8>(P, S, soff, xs, hs, vb, vv, j, 0); break;
      }
    }
  }
  __syncthreads();
  // ---- phase 3: output channel c, hidden slice `part`
  {
    const int parts = FF_THREADS / C;                       // C is a power of two <= 128 on this path: 2, 4, 8, ...
    const int c = tid % C, part = tid / C;
    const int per = ((eC + parts - 1) / parts + 3) & ~3;    // slices start on 16-byte boundaries of the weight row; late ones may be empty
    const int j0 = part * per < eC ? part * per : eC, j1 = j0 + per < eC ? j0 + per : eC;
    float acc[FF_VT];
    ffn_row_walk<FF_VT>(P.W2 + (size_t)c * eC, j0, j1, hs, 0, acc);
#pragma unroll
    for (int i = 0; i < FF_VT; ++i) red[tid][i] = acc[i];
    __syncthreads();
    // epilogue: item = (c, voxel); the slices fold in fixed order
    const float dinv = P.drop_p > 0.f ? 1.0f / (1.0f - P.drop_p) : 1.f;
    const uint32_t okey = P.drop_p > 0.f ? rng_key(P.seed + soff, P.site) : 0u;
    for (int e = tid; e < C * FF_VT; e += FF_THREADS) {
      const int cc = e / FF_VT, i = e % FF_VT;
      if (vb[i] < 0) continue;
      float y = P.b2 ? __ldg(P.b2 + cc) : 0.f;
      for (int pp = 0; pp < parts; ++pp) y += red[pp * C + cc][i];
      const size_t idx = ((size_t)vb[i] * C + cc) * S + vv[i];
      if (P.drop_p > 0.f) y *= keep_from_bits(rng_word(okey, idx), P.drop_p, dinv);
      if (P.res) y = fmaf(P.res_scale, __ldg(P.res + idx), y);
      P.y[idx] = y;
    }
  }
}

static int g_ffn_fused = 1;
void pw_ffn_set(int on) { g_ffn_fused = on; }

int pw_ffn_small(const FfnBatch& batch, cudaStream_t stream) {
  if (!g_ffn_fused || batch.nprob <= 0 || batch.S >= g_small_max_s) return 1;
  double bytes = 0.0, flops = 0.0;
  for (int i = 0; i < batch.nprob; ++i) {
    const FfnProblem& P = batch.p[i];
    // C a power of two in [8, 128] (phase 3 deals 256 threads to C channels), eC <= 256, 16-byte weight rows
    if (P.C < 8 || P.C > FF_MAX_C || (P.C & (P.C - 1)) || P.eC > FF_MAX_E || (P.eC & 3) || ((uintptr_t)P.W1 & 15) || ((uintptr_t)P.W2 & 15)) return 1;
    bytes += 4.0 * batch.B * batch.S * (2.0 * P.C + P.eC + (P.res ? P.C : 0)) + 8.0 * P.C * P.eC;
    flops += 4.0 * batch.B * batch.S * P.C * P.eC;
  }
  FfnBatch launch = batch;
  launch.seed_dev = get_seed_dev();
  prof_bytes(bytes);
  prof_flops(flops);
  VX_LAUNCH(pw_ffn_small_kernel, dim3(cdiv((long long)batch.B * batch.S, FF_VT), batch.nprob), dim3(FF_THREADS), 0, stream, launch);
  return check_launch("pw_ffn_small_kernel");
}

// Data gradient of the fused MLP (same tiling).  Both contractions use the weights TRANSPOSED, so a thread walks a COLUMN of
// the stored matrix: threads own 4 adjacent columns (one 16-byte load per row, adjacent threads adjacent quads -- fully
// coalesced) and the rows of the reduction are split over thread groups that meet in shared memory.
//   phase 2: dh[j][i] = (sum_c W2[c][j] dym[c][i]) GELU'(hpre[j][i]) mask1       thread = (j quad, share of the tile's voxels)
//   phase 3: dx[c][i] = sum_j W1[j][c] dh[j][i]                                  thread = (c quad, slice of the hidden rows)
// Column-quad walk shared by both phases: acc[k][i] += W[r][col0 + k] * v[r][i0 + i] over rows [r0, r1), 8 rows of weights in
// flight.  NV (voxels per thread) is a template parameter: the accumulators are 4 NV registers and no slot is predicated.
template <int NV>
VX_DEV void ffn_col_walk(const float* __restrict__ wcol, int ld, int r0, int r1, const float (*v)[FF_VT], int i0, float (&acc)[4][NV]) {
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[k][i] = 0.f;
  int r = r0;
#pragma unroll 1
  for (; r + 8 <= r1; r += 8) {
    float4 w[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) w[u] = __ldg(reinterpret_cast<const float4*>(wcol + (size_t)(r + u) * ld));
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float ww[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float x = v[r + u][i0 + i];
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k][i] = fmaf(ww[k], x, acc[k][i]);
      }
    }
  }
  for (; r < r1; ++r) {
    const float4 w = __ldg(reinterpret_cast<const float4*>(wcol + (size_t)r * ld));
    const float ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float x = v[r][i0 + i];
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k][i] = fmaf(ww[k], x, acc[k][i]);
    }
  }
}

template <int NV>
VX_DEV void ffn_bwd_hidden(const FfnBwdProblem& P, int S, uint64_t soff, const float (*ds)[FF_VT], float (*hs)[FF_VT], const int* vb,
                           const int* vv, int jq, int i0) {
  float acc[4][NV];
  ffn_col_walk<NV>(P.W2 + 4 * jq, P.eC, 0, P.C, ds, i0, acc);
  const float minv = P.mid_drop_p > 0.f ? 1.0f / (1.0f - P.mid_drop_p) : 1.f;
  const uint32_t mkey = P.mid_drop_p > 0.f ? rng_key(P.mid_seed + soff, P.mid_site) : 0u;
  // every hpre load first (dead voxels read element 0), then the arithmetic: one round trip instead of 4 NV
  float hp[4][NV];
  size_t idx[4][NV];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int ii = i0 + i;
      idx[k][i] = vb[ii] >= 0 ? ((size_t)vb[ii] * P.eC + 4 * jq + k) * S + vv[ii] : 0;
      hp[k][i] = __ldg(P.hpre + idx[k][i]);
    }
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int ii = i0 + i;
      float g = acc[k][i] * gelu_grad_f(hp[k][i]);
      if (P.mid_drop_p > 0.f) g *= keep_from_bits(rng_word(mkey, idx[k][i]), P.mid_drop_p, minv);
      if (vb[ii] >= 0) P.dh[idx[k][i]] = g; else g = 0.f;
      hs[4 * jq + k][ii] = g;
    }
}

template <int NV>
VX_DEV void ffn_bwd_input(const FfnBwdProblem& P, const float (*hs)[FF_VT], float (*red)[FF_VT], int cq, int part, int parts, int i0) {
  const int per = (P.eC + parts - 1) / parts;
  const int j0 = part * per < P.eC ? part * per : P.eC, j1 = j0 + per < P.eC ? j0 + per : P.eC;
  float acc[4][NV];
  ffn_col_walk<NV>(P.W1 + 4 * cq, P.C, j0, j1, hs, i0, acc);
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int i = 0; i < NV; ++i) red[part * P.C + 4 * cq + k][i0 + i] = acc[k][i];
}

__global__ void __launch_bounds__(FF_THREADS, 2) pw_ffn_small_bwd_kernel(const __grid_constant__ FfnBwdBatch batch) {
  VX_PDL_ENTRY();
  const FfnBwdProblem& P = batch.p[blockIdx.y];
  const int S = batch.S, C = P.C, eC = P.eC, tid = threadIdx.x;
  const int f0 = blockIdx.x * FF_VT, total = batch.B * S;
  __align__(16) __shared__ float ds[FF_MAX_C][FF_VT];           // dy * mask2
  __align__(16) __shared__ float hs[FF_MAX_E][FF_VT];           // dh
  __align__(16) __shared__ float red[8 * FF_MAX_C][FF_VT];      // phase-3 slices: [part][c][voxel], parts * C <= 1024
  __shared__ int vb[FF_VT], vv[FF_VT];
  const uint64_t soff = batch.seed_dev ? (uint64_t)__ldg(batch.seed_dev) : 0;
  if (tid < FF_VT) {
    const int f = f0 + tid;
    vb[tid] = f < total ? f / S : -1;
    vv[tid] = f < total ? f % S : 0;
  }
  __syncthreads();
  {
    const float oinv = P.out_drop_p > 0.f ? 1.0f / (1.0f - P.out_drop_p) : 1.f;
    const uint32_t okey = P.out_drop_p > 0.f ? rng_key(P.out_seed + soff, P.out_site) : 0u;
    for (int e = tid; e < C * FF_VT; e += FF_THREADS) {
      const int c = e / FF_VT, i = e % FF_VT;
      float v = 0.f;
      if (vb[i] >= 0) {
        const size_t idx = ((size_t)vb[i] * C + c) * S + vv[i];
        v = __ldg(P.dy + idx);
        if (P.out_drop_p > 0.f) v *= keep_from_bits(rng_word(okey, idx), P.out_drop_p, oinv);
      }
      ds[c][i] = v;
    }
  }
  __syncthreads();
  // ---- phase 2
  {
    const int Q = eC >> 2;                                  // column quads of W2 (<= 64, a power of two)
    const int nsp = FF_THREADS / Q < FF_VT ? FF_THREADS / Q : FF_VT;      // voxel shares
    const int jq = tid % Q, vs = tid / Q;
    if (vs < nsp) {
      switch (nsp) {
        case 8: ffn_bwd_hidden<1>(P, S, soff, ds, hs, vb, vv, jq, vs); break;
        case 4: ffn_bwd_hidden<2>(P, S, soff, ds, hs, vb, vv, jq, vs * 2); break;
        default: ffn_bwd_hidden<4>(P, S, soff, ds, hs, vb, vv, jq, vs * 4); break;      // eC = 512 would land here; FF_MAX_E is 256
      }
    }
  }
  __syncthreads();
  // ---- phase 3
  const int parts = 8;                                      // slices of the hidden rows; parts * C <= 1024 rows of `red`
  {
    const int Q = C >> 2;                                   // column quads of W1 (2 ... 32)
    const int nsp = FF_THREADS / (Q * parts) < FF_VT ? FF_THREADS / (Q * parts) : FF_VT;      // voxel shares (1 ... 8)
    const int cq = tid % Q, part = (tid / Q) % parts, vs = tid / (Q * parts);
    if (vs < nsp) {
      switch (nsp) {
        case 8: ffn_bwd_input<1>(P, hs, red, cq, part, parts, vs); break;
        case 4: ffn_bwd_input<2>(P, hs, red, cq, part, parts, vs * 2); break;
        case 2: ffn_bwd_input<4>(P, hs, red, cq, part, parts, vs * 4); break;
        default: ffn_bwd_input<8>(P, hs, red, cq, part, parts, 0); break;
      }
    }
  }
  __syncthreads();
  for (int e = tid; e < C * FF_VT; e += FF_THREADS) {
    const int cc = e / FF_VT, i = e % FF_VT;
    if (vb[i] < 0) continue;
    float y = 0.f;
#pragma unroll
    for (int pp = 0; pp < parts; ++pp) y += red[pp * C + cc][i];
    P.dx[((size_t)vb[i] * C + cc) * S + vv[i]] = y;
  }
}

int pw_ffn_small_bwd(const FfnBwdBatch& batch, cudaStream_t stream) {
  if (!g_ffn_fused || batch.nprob <= 0 || batch.S >= g_small_max_s) return 1;
  double bytes = 0.0, flops = 0.0;
  for (int i = 0; i < batch.nprob; ++i) {
    const FfnBwdProblem& P = batch.p[i];
    // C a power of two in [8, 128] (rows of 8 in flight), eC <= 256 and a multiple of 4, eC / 4 a divisor of 256, 16-byte rows
    if (P.C < 8 || P.C > FF_MAX_C || (P.C & (P.C - 1)) || P.eC > FF_MAX_E || (P.eC & 3) || (FF_THREADS % (P.eC >> 2)) ||
        ((uintptr_t)P.W1 & 15) || ((uintptr_t)P.W2 & 15)) return 1;
    bytes += 4.0 * batch.B * batch.S * (2.0 * P.C + 2.0 * P.eC) + 8.0 * P.C * P.eC;
    flops += 4.0 * batch.B * batch.S * P.C * P.eC;
  }
  FfnBwdBatch launch = batch;
  launch.seed_dev = get_seed_dev();
  prof_bytes(bytes);
  prof_flops(flops);
  VX_LAUNCH(pw_ffn_small_bwd_kernel, dim3(cdiv((long long)batch.B * batch.S, FF_VT), batch.nprob), dim3(FF_THREADS), 0, stream, launch);
  return check_launch("pw_ffn_small_bwd_kernel");
}

// ---------------------------------------------------------------------------------------------------
// weight gradient
// ---------------------------------------------------------------------------------------------------
constexpr int WG_THREADS = 256;

// dW = dY X^T is a small-output GEMM whose reduction axis is (batch, voxel).  A CTA owns a block of 4x4 output tiles
// (all output channels x a strip of input channels) and walks its share of the (batch, voxel-chunk) list: blockIdx.x
// splits that list only as far as needed to fill the machine, so problems with many channels and few voxels (levels 3-4)
// run without any atomics, and large-voxel problems (level 1) issue grid.x atomics per element.  Per chunk the dY and X
// tiles are transposed into shared memory ([voxel][channel], rows padded to 4*odd words so the float4 reads of 4
// consecutive channels are conflict-free); when the tile block has fewer tiles than threads the voxels of a chunk are
// split between thread groups and folded through shared memory at the end.
__global__ void __launch_bounds__(WG_THREADS) pw_wgrad_kernel(const __grid_constant__ WgBatch batch, int TV, int nK) {
  VX_PDL_ENTRY();
  const WgProblem& P = batch.p[blockIdx.z];
  const int S = batch.S, B = batch.B;
  const int Co = P.Co, Ci = P.Ci;
  const int Co4 = (Co + 3) & ~3, Ci4 = (Ci + 1 + 3) & ~3;    // +1: the all-ones row that yields db
  const int nCo4 = Co4 >> 2, nCi4 = Ci4 >> 2;
  const int WCI = max(1, WG_THREADS / nCo4);                  // input-channel tiles per CTA
  const int ci4_0 = blockIdx.y * WCI;
  if (ci4_0 >= nCi4) return;
  const int width = min(WCI, nCi4 - ci4_0);
  const int nt = nCo4 * width;                                // tiles of this CTA (<= WG_THREADS)
  int G = 1;
  while (G * 2 * nt <= WG_THREADS && (TV / (G * 2)) >= 4) G *= 2;
  const int span = TV / G;
  const int CoP = Co4 + 4, CiP = WCI * 4 + 4;
  VX_DYN_SMEM(float, sm);
  float* sY = sm;                     // [TV][CoP]
  float* sX = sm + (size_t)TV * CoP;  // [TV][CiP]   (only this CTA's input-channel strip)
  const int tid = threadIdx.x;
  const bool active = tid < nt * G;
  const int tile = active ? tid % nt : 0, grp = active ? tid / nt : 0;
  const int co4 = tile % nCo4, ci4l = tile / nCo4;
  const int cbeg = ci4_0 * 4, ccount = width * 4;             // global input-channel range [cbeg, cbeg + ccount)
  const float yinv = P.y_drop_p > 0.f ? 1.0f / (1.0f - P.y_drop_p) : 1.f;
  const float xinv = P.x_drop_p > 0.f ? 1.0f / (1.0f - P.x_drop_p) : 1.f;
  const int nchunks = (S + TV - 1) / TV;
  const uint64_t soff = batch.seed_dev ? (uint64_t)__ldg(batch.seed_dev) : 0;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int ck = blockIdx.x; ck < B * nchunks; ck += nK) {
    const int b = ck / nchunks, vbase = (ck % nchunks) * TV;
    __syncthreads();
    // staging in rounds of WG_U elements per thread: every load of a round is issued before the first prologue (the rolled
    // one-load-per-iteration form made a thread wait out ~80 L2 round trips per chunk at levels 3-4); TV is a power of two
    constexpr int WG_U = 4;
    const int tvs = __ffs(TV) - 1;
#pragma unroll 1
    for (int idx0 = tid; idx0 < Co4 * TV; idx0 += WG_U * WG_THREADS) {
      float val[WG_U];
      size_t gi[WG_U];
      bool ok[WG_U];
#pragma unroll
      for (int u = 0; u < WG_U; ++u) {
        const int idx = idx0 + u * WG_THREADS;
        const int co = idx >> tvs, gv = vbase + (idx & (TV - 1));
        ok[u] = idx < Co4 * TV && co < Co && gv < S;
        gi[u] = ok[u] ? ((size_t)b * Co + co) * S + gv : 0;
        val[u] = ok[u] ? __ldg(P.dY + gi[u]) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < WG_U; ++u) {
        const int idx = idx0 + u * WG_THREADS;
        if (idx >= Co4 * TV) continue;
        if (ok[u] && P.y_drop_p > 0.f) val[u] *= dropout_scale(P.y_seed + soff, P.y_site, gi[u], P.y_drop_p, yinv);
        sY[(idx & (TV - 1)) * CoP + (idx >> tvs)] = val[u];
      }
    }
#pragma unroll 1
    for (int idx0 = tid; idx0 < ccount * TV; idx0 += WG_U * WG_THREADS) {
      float val[WG_U];
      bool ok[WG_U];
#pragma unroll
      for (int u = 0; u < WG_U; ++u) {
        const int idx = idx0 + u * WG_THREADS;
        const int cg = cbeg + (idx >> tvs), gv = vbase + (idx & (TV - 1));
        ok[u] = idx < ccount * TV && gv < S && cg < Ci;
        val[u] = 0.f;
        if (ok[u]) {
          int c = cg, s2 = 0;
          while (s2 < P.nsrc - 1 && c >= P.src[s2].C) { c -= P.src[s2].C; ++s2; }
          val[u] = __ldg(P.src[s2].ptr + ((size_t)b * P.src[s2].C + c) * S + gv);
        } else if (idx < ccount * TV && gv < S && cg == Ci) {
          val[u] = 1.f;
        }
      }
#pragma unroll
      for (int u = 0; u < WG_U; ++u) {
        const int idx = idx0 + u * WG_THREADS;
        if (idx >= ccount * TV) continue;
        const int cl = idx >> tvs, v = idx & (TV - 1);
        if (ok[u]) {
          const int cg = cbeg + cl, gv = vbase + v;
          if (P.xpro == PRO_AFFINE) {
            const int k = b * P.x_bstride + cg;
            val[u] = fmaf(val[u], __ldg(P.xa + k), __ldg(P.xc + k));
          } else if (P.xpro == PRO_GELU) {
            val[u] = gelu_f(val[u]);
          } else if (P.xpro == PRO_GELU_DROPOUT) {
            val[u] = gelu_f(val[u]) * dropout_scale(P.x_seed + soff, P.x_site, ((uint64_t)b * Ci + cg) * (uint64_t)S + gv, P.x_drop_p, xinv);
          }
        }
        sX[v * CiP + cl] = val[u];
      }
    }
    __syncthreads();
    if (active) {
      const float* py = sY + (size_t)grp * span * CoP + co4 * 4;
      const float* px = sX + (size_t)grp * span * CiP + ci4l * 4;
#pragma unroll 4
      for (int v = 0; v < span; ++v) {
        const float4 y = *reinterpret_cast<const float4*>(py + (size_t)v * CoP);
        const float4 x = *reinterpret_cast<const float4*>(px + (size_t)v * CiP);
        const float yy[4] = {y.x, y.y, y.z, y.w};
        const float xx[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(yy[i], xx[j], acc[i][j]);
      }
    }
  }
  // fold the G voxel slices through shared memory, then one global update per element
  __syncthreads();
  float* fold = sm;                  // [nt][16]
  if (G > 1) {
    for (int i = tid; i < nt * 16; i += WG_THREADS) fold[i] = 0.f;
    __syncthreads();
    if (active) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(fold + tile * 16 + i * 4 + j, acc[i][j]);
    }
  } else if (active) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) fold[tile * 16 + i * 4 + j] = acc[i][j];
  }
  __syncthreads();
#pragma unroll 1
  for (int e = tid; e < nt * 16; e += WG_THREADS) {
    const int tl = e >> 4, i = (e >> 2) & 3, j = e & 3;
    const int co = (tl % nCo4) * 4 + i, ci = cbeg + (tl / nCo4) * 4 + j;
    if (co >= Co) continue;
    if (ci < Ci) atomicAdd(P.dW + (size_t)co * P.ld + ci, fold[e]);
    else if (ci == Ci && P.db) atomicAdd(P.db + co, fold[e]);
  }
}

int pw_wgrad(const WgBatch& batch, cudaStream_t main_stream) {
  if (batch.nprob <= 0) return VX_OK;
  cudaStream_t stream = side_fork(main_stream);      // joined by the SideJoin of the calling backward op
  {
    const int rc = pw_wgrad_tc(batch, stream);
    if (rc <= 0) return rc;
  }
  int maxrow = 0, maxblocks = 1, maxfold = 0;
  for (int i = 0; i < batch.nprob; ++i) {
    const WgProblem& P = batch.p[i];
    int cs = 0;
    for (int s = 0; s < P.nsrc; ++s) cs += P.src[s].C;
    if (cs != P.Ci) { set_error("pw_wgrad: source channels %d != Ci %d", cs, P.Ci); return VX_ERR_BAD_DESC; }
    const int Co4 = (P.Co + 3) & ~3, nCo4 = Co4 >> 2, nCi4 = ((P.Ci + 4) & ~3) >> 2;
    if (nCo4 > WG_THREADS) { set_error("pw_wgrad: %d output channels not supported", P.Co); return VX_ERR_UNSUPPORTED; }
    const int WCI = WG_THREADS / nCo4 > 0 ? WG_THREADS / nCo4 : 1;
    const int row = (Co4 + 4) + (WCI * 4 + 4);
    maxrow = row > maxrow ? row : maxrow;
    const int blocks = cdiv(nCi4, WCI);
    maxblocks = blocks > maxblocks ? blocks : maxblocks;
    const int fold = WG_THREADS * 16;
    maxfold = fold > maxfold ? fold : maxfold;
  }
  int TV = 128;
  while (TV > 16 && (size_t)TV * maxrow * sizeof(float) > 64 * 1024) TV >>= 1;
  while (TV > 16 && TV / 2 >= batch.S) TV >>= 1;
  size_t smem = (size_t)TV * maxrow * sizeof(float);
  if (smem < (size_t)maxfold * sizeof(float)) smem = (size_t)maxfold * sizeof(float);
  if (smem > 200 * 1024) { set_error("pw_wgrad: channel count too large"); return VX_ERR_UNSUPPORTED; }
  VX_SET_SMEM(pw_wgrad_kernel, smem);
  // split the (batch, chunk) list only as far as needed to put ~2 CTAs on every SM
  const int nchunks = batch.B * cdiv(batch.S, TV);
  int nK = (2 * kSMs) / (maxblocks * batch.nprob);
  if (nK < 1) nK = 1;
  if (nK > nchunks) nK = nchunks;
  dim3 grid(nK, maxblocks, batch.nprob);
  WgBatch launch = batch;
  launch.seed_dev = get_seed_dev();
  {
    double bytes = 0.0, flops = 0.0;
    for (int i = 0; i < batch.nprob; ++i) {
      bytes += 4.0 * batch.B * batch.S * (batch.p[i].Ci + batch.p[i].Co) + 4.0 * batch.p[i].Ci * batch.p[i].Co;
      flops += 2.0 * batch.B * batch.S * batch.p[i].Ci * batch.p[i].Co;
    }
    prof_bytes(bytes);
    prof_flops(flops);
  }
  VX_LAUNCH(pw_wgrad_kernel, grid, dim3(WG_THREADS), smem, stream, launch, TV, nK);
  return check_launch("pw_wgrad_kernel");
}

// ---------------------------------------------------------------------------------------------------
// InstanceNorm over rows (one CTA per (b,c) row; rows are <= 16 K elements on this path, L1/L2 resident)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) inorm_rows_fwd_kernel(const float* __restrict__ x, const float* __restrict__ addend,
                                                             float* __restrict__ y, float* __restrict__ stats, int S,
                                                             float eps) {
  VX_PDL_ENTRY();
  __shared__ float red[33];
  const size_t base = (size_t)blockIdx.x * S;
  float s = 0.f;
  for (int i = threadIdx.x; i < S; i += blockDim.x) s += x[base + i];
  const float mean = block_sum(s, red) / (float)S;
  float q = 0.f;
  for (int i = threadIdx.x; i < S; i += blockDim.x) { const float d = x[base + i] - mean; q = fmaf(d, d, q); }
  const float var = block_sum(q, red) / (float)S;
  const float rstd = 1.0f / sqrtf(var + eps);
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    float v = (x[base + i] - mean) * rstd;
    if (addend) v += addend[base + i];
    y[base + i] = v;
  }
  if (threadIdx.x == 0 && stats) { stats[2 * blockIdx.x] = mean; stats[2 * blockIdx.x + 1] = rstd; }
}

// Register-cached variant for rows of up to 4 float4 per thread (S <= 16 K at 1024 threads, S % 4 == 0): the row is read
// from global memory once instead of three times, 16 bytes per access, and a level-1 row (13 824 voxels) gets 864-1024
// threads instead of 256.  Same arithmetic (mean, then centred variance) and a fixed reduction order.
constexpr int IN_Q = 4;
__global__ void __launch_bounds__(1024) inorm_rows_fwd_cached_kernel(const float* __restrict__ x, const float* __restrict__ addend,
                                                                    float* __restrict__ y, float* __restrict__ stats, int S,
                                                                    float eps) {
  VX_PDL_ENTRY();
  __shared__ float red[33];
  const size_t base = (size_t)blockIdx.x * S;
  const int nq = S >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x + base);
  float4 v[IN_Q];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < IN_Q; ++j) {
    const int i = threadIdx.x + j * blockDim.x;
    v[j] = i < nq ? __ldg(x4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  const float mean = block_sum(s, red) / (float)S;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < IN_Q; ++j) {
    const int i = threadIdx.x + j * blockDim.x;
    if (i < nq) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float var = block_sum(q, red) / (float)S;
  const float rstd = 1.0f / sqrtf(var + eps);
  float4* y4 = reinterpret_cast<float4*>(y + base);
  const float4* a4 = addend ? reinterpret_cast<const float4*>(addend + base) : nullptr;
#pragma unroll
  for (int j = 0; j < IN_Q; ++j) {
    const int i = threadIdx.x + j * blockDim.x;
    if (i < nq) {
      float4 o = make_float4((v[j].x - mean) * rstd, (v[j].y - mean) * rstd, (v[j].z - mean) * rstd, (v[j].w - mean) * rstd);
      if (a4) { const float4 a = __ldg(a4 + i); o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w; }
      y4[i] = o;
    }
  }
  if (threadIdx.x == 0 && stats) { stats[2 * blockIdx.x] = mean; stats[2 * blockIdx.x + 1] = rstd; }
}

static int inorm_cached_threads(const void* a, const void* b, const void* c, const void* d, int S) {
  if ((S & 3) || S < 1024 || S > IN_Q * 4 * 1024) return 0;
  if ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)c) | ((uintptr_t)d)) & 15) return 0;
  int t = ((S / 4 + IN_Q - 1) / IN_Q + 31) & ~31;
  return t < 128 ? 128 : t;
}

int inorm_rows_fwd(const float* x, const float* addend, float* y, float* stats, int rows, int S, float eps,
                   cudaStream_t stream) {
  if (rows <= 0) return VX_OK;
  prof_bytes(4.0 * rows * (double)S * (addend ? 3 : 2));
  if (const int t = inorm_cached_threads(x, addend, y, nullptr, S)) {
    VX_LAUNCH(inorm_rows_fwd_cached_kernel, dim3(rows), dim3(t), 0, stream, x, addend, y, stats, S, eps);
    return check_launch("inorm_rows_fwd_cached_kernel");
  }
  VX_LAUNCH(inorm_rows_fwd_kernel, dim3(rows), dim3(256), 0, stream, x, addend, y, stats, S, eps);
  return check_launch("inorm_rows_fwd_kernel");
}

__global__ void __launch_bounds__(256) inorm_rows_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                             const float* __restrict__ stats,
                                                             const float* __restrict__ dx_add, float* __restrict__ dx,
                                                             int S, float* __restrict__ db, int C) {
  VX_PDL_ENTRY();
  __shared__ float red[33];
  const size_t base = (size_t)blockIdx.x * S;
  const float mean = stats[2 * blockIdx.x], rstd = stats[2 * blockIdx.x + 1];
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    const float g = dy[base + i];
    s1 += g;
    s2 = fmaf(g, (x[base + i] - mean) * rstd, s2);
  }
  const float m1 = block_sum(s1, red) / (float)S;
  const float m2 = block_sum(s2, red) / (float)S;
  float bs = 0.f;
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    const float xh = (x[base + i] - mean) * rstd;
    float v = rstd * (dy[base + i] - m1 - xh * m2);
    bs += v;
    if (dx_add) v += dx_add[base + i];
    dx[base + i] = v;
  }
  if (db) {        // gradient of a per-channel bias in front of the norm: sum of this row's dx
    bs = block_sum(bs, red);
    if (threadIdx.x == 0) atomicAdd(db + blockIdx.x % C, bs);
  }
}

__global__ void __launch_bounds__(1024) inorm_rows_bwd_cached_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                    const float* __restrict__ stats,
                                                                    const float* __restrict__ dx_add, float* __restrict__ dx,
                                                                    int S, float* __restrict__ db, int C) {
  VX_PDL_ENTRY();
  __shared__ float red[33];
  const size_t base = (size_t)blockIdx.x * S;
  const int nq = S >> 2;
  const float mean = stats[2 * blockIdx.x], rstd = stats[2 * blockIdx.x + 1];
  const float4* g4 = reinterpret_cast<const float4*>(dy + base);
  const float4* x4 = reinterpret_cast<const float4*>(x + base);
  float4 g[IN_Q], h[IN_Q];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < IN_Q; ++j) {
    const int i = threadIdx.x + j * blockDim.x;
    g[j] = make_float4(0.f, 0.f, 0.f, 0.f); h[j] = g[j];
    if (i < nq) {
      g[j] = __ldg(g4 + i);
      const float4 xv = __ldg(x4 + i);
      h[j] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
    }
    s1 += (g[j].x + g[j].y) + (g[j].z + g[j].w);
    s2 += (g[j].x * h[j].x + g[j].y * h[j].y) + (g[j].z * h[j].z + g[j].w * h[j].w);
  }
  const float m1 = block_sum(s1, red) / (float)S;
  const float m2 = block_sum(s2, red) / (float)S;
  float4* o4 = reinterpret_cast<float4*>(dx + base);
  const float4* a4 = dx_add ? reinterpret_cast<const float4*>(dx_add + base) : nullptr;
  float bs = 0.f;
#pragma unroll
  for (int j = 0; j < IN_Q; ++j) {
    const int i = threadIdx.x + j * blockDim.x;
    if (i < nq) {
      float4 o = make_float4(rstd * (g[j].x - m1 - h[j].x * m2), rstd * (g[j].y - m1 - h[j].y * m2),
                             rstd * (g[j].z - m1 - h[j].z * m2), rstd * (g[j].w - m1 - h[j].w * m2));
      bs += (o.x + o.y) + (o.z + o.w);
      if (a4) { const float4 a = __ldg(a4 + i); o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w; }
      o4[i] = o;
    }
  }
  if (db) {
    bs = block_sum(bs, red);
    if (threadIdx.x == 0) atomicAdd(db + blockIdx.x % C, bs);
  }
}

int inorm_rows_bwd(const float* dy, const float* x, const float* stats, const float* dx_add, float* dx, int rows,
                   int S, cudaStream_t stream, float* db, int C) {
  if (rows <= 0) return VX_OK;
  if (db) {
    ZeroList zl;
    zl.add(db, C);
    const int rc = zero_many(zl, stream);
    if (rc != VX_OK) return rc;
  }
  prof_bytes(4.0 * rows * (double)S * (dx_add ? 4 : 3));
  if (const int t = inorm_cached_threads(dy, x, dx_add, dx, S)) {
    VX_LAUNCH(inorm_rows_bwd_cached_kernel, dim3(rows), dim3(t), 0, stream, dy, x, stats, dx_add, dx, S, db, C);
    return check_launch("inorm_rows_bwd_cached_kernel");
  }
  VX_LAUNCH(inorm_rows_bwd_kernel, dim3(rows), dim3(256), 0, stream, dy, x, stats, dx_add, dx, S, db, C);
  return check_launch("inorm_rows_bwd_kernel");
}

__global__ void stats_to_affine_kernel(const float* __restrict__ stats, float* __restrict__ a, float* __restrict__ c,
                                       int rows) {
  VX_PDL_ENTRY();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows) { const float m = stats[2 * i], r = stats[2 * i + 1]; a[i] = r; c[i] = -m * r; }
}

int stats_to_affine(const float* stats, float* a, float* c, int rows, cudaStream_t stream) {
  VX_LAUNCH(stats_to_affine_kernel, dim3(cdiv(rows, 128)), dim3(128), 0, stream, stats, a, c, rows);
  return check_launch("stats_to_affine_kernel");
}

// ---------------------------------------------------------------------------------------------------
// zero_many: blockIdx.y = buffer, blockIdx.x strides its elements (16-byte stores where the buffer allows)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) zero_many_kernel(const __grid_constant__ ZeroList Z) {
  VX_PDL_ENTRY();
  float* p = Z.ptr[blockIdx.y];
  const unsigned n = Z.n[blockIdx.y];
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
  if ((((uintptr_t)p) & 15) == 0) {
    const unsigned n4 = n >> 2;
    float4* p4 = reinterpret_cast<float4*>(p);
    for (unsigned i = tid; i < n4; i += nthr) p4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (unsigned i = (n4 << 2) + tid; i < n; i += nthr) p[i] = 0.f;
  } else {
    for (unsigned i = tid; i < n; i += nthr) p[i] = 0.f;
  }
}

int zero_many(const ZeroList& z, cudaStream_t stream) {
  if (z.overflow) { set_error("zero_many: more than %d buffers", VX_ZERO_MAX); return VX_ERR_BAD_DESC; }
  if (z.count == 0) return VX_OK;
  unsigned nmax = 0;
  for (int i = 0; i < z.count; ++i) nmax = z.n[i] > nmax ? z.n[i] : nmax;
  int gx = cdiv(nmax, 256 * 4 * 4);          // <= 4 float4 per thread for the largest buffer
  if (gx < 1) gx = 1;
  if (gx > 64) gx = 64;
  VX_LAUNCH(zero_many_kernel, dim3(gx, z.count), dim3(256), 0, stream, z);
  return check_launch("zero_many_kernel");
}

// ---------------------------------------------------------------------------------------------------
// channel-first LayerNorm (thread per voxel, channel loop strides by S so every warp access is coalesced)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ln_fwd_kernel(const __grid_constant__ LnBatch L) {
  VX_PDL_ENTRY();
  const int t = blockIdx.z, b = blockIdx.y;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= L.S) return;
  const int C = L.C, S = L.S;
  const float* x = L.x[t] + (size_t)b * C * S + v;
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += x[(size_t)c * S];
  const float mean = s / (float)C;
  float q = 0.f;
  for (int c = 0; c < C; ++c) { const float d = x[(size_t)c * S] - mean; q = fmaf(d, d, q); }
  const float rstd = 1.0f / sqrtf(q / (float)C + L.eps);
  float* xh = L.xhat[t] + (size_t)b * C * S + v;
  for (int c = 0; c < C; ++c) xh[(size_t)c * S] = (x[(size_t)c * S] - mean) * rstd;
  L.rstd[t][(size_t)b * S + v] = rstd;
}

// Few voxels, many channels (PatchMerging at levels 3-4: 216 / 27 voxels x 256 / 512 channels): the thread-per-voxel kernel
// is 4-8 CTAs walking 3 x C strided loads each.  Here a CTA is 32 voxels (lanes) x 8 warps that split the channel axis and
// meet in shared memory; same two-pass arithmetic.
__global__ void __launch_bounds__(256) ln_fwd_wide_kernel(const __grid_constant__ LnBatch L) {
  VX_PDL_ENTRY();
  const int t = blockIdx.z, b = blockIdx.y;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int v = blockIdx.x * 32 + lane;
  const int C = L.C, S = L.S;
  const bool ok = v < S;
  __shared__ float red[8][32];
  const float* x = L.x[t] + (size_t)b * C * S + (ok ? v : 0);
  float s = 0.f;
  for (int c = w; c < C; c += 8) s += ok ? __ldg(x + (size_t)c * S) : 0.f;
  red[w][lane] = s;
  __syncthreads();
  s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += red[i][lane];
  const float mean = s / (float)C;
  __syncthreads();
  float q = 0.f;
  for (int c = w; c < C; c += 8) { const float d = ok ? __ldg(x + (size_t)c * S) - mean : 0.f; q = fmaf(d, d, q); }
  red[w][lane] = q;
  __syncthreads();
  q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) q += red[i][lane];
  if (!ok) return;
  const float rstd = 1.0f / sqrtf(q / (float)C + L.eps);
  float* xh = L.xhat[t] + (size_t)b * C * S + v;
  for (int c = w; c < C; c += 8) xh[(size_t)c * S] = (__ldg(x + (size_t)c * S) - mean) * rstd;
  if (w == 0) L.rstd[t][(size_t)b * S + v] = rstd;
}

int ln_forward(const LnBatch& L, cudaStream_t stream) {
  if (L.n <= 0) return VX_OK;
  prof_bytes(4.0 * L.n * L.B * (double)L.S * (2.0 * L.C + 1));
  if (L.S < 1024 && L.C >= 64) {
    VX_LAUNCH(ln_fwd_wide_kernel, dim3(cdiv(L.S, 32), L.B, L.n), dim3(256), 0, stream, L);
    return check_launch("ln_fwd_wide_kernel");
  }
  VX_LAUNCH(ln_fwd_kernel, dim3(cdiv(L.S, 128), L.B, L.n), dim3(128), 0, stream, L);
  return check_launch("ln_fwd_kernel");
}

// Backward of the channel-first LayerNorm.  CTA = 32 consecutive voxels (lanes) x 8 warps striding the channel axis:
// every global access is a coalesced run along the voxel axis, a channel is owned by exactly one warp of the CTA (its
// dgamma / dbeta partial is one warp reduction + one global atomic), and the per-voxel means are combined through
// shared memory.
__global__ void __launch_bounds__(256) ln_bwd_kernel(const __grid_constant__ LnBwdBatch L, int groups) {
  VX_PDL_ENTRY();
  // `groups` consecutive blocks of 32 voxels per CTA: the dgamma / dbeta partial sums of a warp's channels stay in registers
  // across them (when a warp owns <= LN_NC channels), so the global atomics -- 2 per channel and CTA onto the same 2 C
  // addresses from every CTA -- shrink by that factor.
  constexpr int LN_NC = 8;
  const int t = blockIdx.z, b = blockIdx.y;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int C = L.C, S = L.S;
  __shared__ float red[2][8][32];
  const float* gamma = L.gamma[t];
  const bool keep = (C + 7) / 8 <= LN_NC;
  float ag[LN_NC], ab[LN_NC];
#pragma unroll
  for (int i = 0; i < LN_NC; ++i) { ag[i] = 0.f; ab[i] = 0.f; }
  for (int gi = 0; gi < groups; ++gi) {
    const int v = (blockIdx.x * groups + gi) * 32 + lane;
    if ((blockIdx.x * groups + gi) * 32 >= S) break;
    const bool ok = v < S;
    const size_t off = (size_t)b * C * S + (ok ? v : 0);
    const float* dout = L.dout[t] + off;
    const float* xh = L.xhat[t] + off;
    float m1 = 0.f, m2 = 0.f;
    int n = 0;
    for (int c = w; c < C; c += 8, ++n) {
      const float d = ok ? __ldg(dout + (size_t)c * S) : 0.f;
      const float h = ok ? __ldg(xh + (size_t)c * S) : 0.f;
      const float g = __ldg(gamma + c) * d;
      m1 += g;
      m2 = fmaf(g, h, m2);
      if (keep) {
#pragma unroll
        for (int i = 0; i < LN_NC; ++i) if (i == n) { ag[i] = fmaf(d, h, ag[i]); ab[i] += d; }
      } else {
        const float sg = warp_sum(d * h), sb = warp_sum(d);
        if (lane == 0) {
          if (L.dgamma[t]) atomicAdd(L.dgamma[t] + c, sg);
          if (L.dbeta[t]) atomicAdd(L.dbeta[t] + c, sb);
        }
      }
    }
    __syncthreads();                 // red[] of the previous group has been consumed
    red[0][w][lane] = m1;
    red[1][w][lane] = m2;
    __syncthreads();
    m1 = 0.f; m2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { m1 += red[0][i][lane]; m2 += red[1][i][lane]; }
    if (ok) {
      m1 /= (float)C; m2 /= (float)C;
      const float rstd = L.rstd[t][(size_t)b * S + v];
      float* dx = L.dx[t] + off;
      const float* add = L.dx_add[t] ? L.dx_add[t] + off : nullptr;
      for (int c = w; c < C; c += 8) {
        const float g = __ldg(gamma + c) * __ldg(dout + (size_t)c * S);
        float r = rstd * (g - m1 - __ldg(xh + (size_t)c * S) * m2);
        if (add) r = fmaf(L.dx_add_scale, __ldg(add + (size_t)c * S), r);
        dx[(size_t)c * S] = r;
      }
    }
  }
  if (keep) {
    int n = 0;
    for (int c = w; c < C; c += 8, ++n) {
      float sg = 0.f, sb = 0.f;
#pragma unroll
      for (int i = 0; i < LN_NC; ++i) if (i == n) { sg = ag[i]; sb = ab[i]; }
      sg = warp_sum(sg); sb = warp_sum(sb);
      if (lane == 0) {
        if (L.dgamma[t]) atomicAdd(L.dgamma[t] + c, sg);
        if (L.dbeta[t]) atomicAdd(L.dbeta[t] + c, sb);
      }
    }
  }
}

// Narrow-channel variant (C = 8 / 16: level 1 of every configuration, where S is largest).  The wide kernel above spreads the
// channels of a voxel over 8 warps and meets in shared memory: two barriers and two dependent load rounds per 32-voxel
// block, walked serially -- ncu: 11 warps per issue stalled on the long scoreboard, 16 % issue-active, 28 us per launch
// (profiles/r2n_step_stalls.txt).  Here a thread owns ALL channels of its voxel: 2 C independent coalesced loads in flight,
// the per-voxel means are plain register sums (no shared memory, no barrier), dgamma / dbeta partials stay in registers over
// the warp's voxel blocks and leave through one warp reduction per channel at the end.
template <int C>
__global__ void __launch_bounds__(256) ln_bwd_narrow_kernel(const __grid_constant__ LnBwdBatch L, int bpw) {
  VX_PDL_ENTRY();
  const int t = blockIdx.z, b = blockIdx.y;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int S = L.S;
  __shared__ float sg[8][2 * C];
  __shared__ float gam[C];
  if (threadIdx.x < C) gam[threadIdx.x] = __ldg(L.gamma[t] + threadIdx.x);
  __syncthreads();
  float ag[C], ab[C];
#pragma unroll
  for (int c = 0; c < C; ++c) { ag[c] = 0.f; ab[c] = 0.f; }
  const int nblk = (S + 31) >> 5;
  const int blk0 = (blockIdx.x * 8 + w) * bpw;
#pragma unroll 1
  for (int j = 0; j < bpw; ++j) {
    const int blk = blk0 + j;
    if (blk >= nblk) break;                                   // warp-uniform
    const int v = blk * 32 + lane;
    const bool ok = v < S;
    const size_t off = (size_t)b * C * S + (ok ? v : 0);
    const float* dout = L.dout[t] + off;
    const float* xh = L.xhat[t] + off;
    float d[C], h[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { d[c] = ok ? __ldg(dout + (size_t)c * S) : 0.f; h[c] = ok ? __ldg(xh + (size_t)c * S) : 0.f; }
    const float rstd = ok ? L.rstd[t][(size_t)b * S + v] : 0.f;
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float g = gam[c] * d[c];
      m1 += g; m2 = fmaf(g, h[c], m2);
      ag[c] = fmaf(d[c], h[c], ag[c]); ab[c] += d[c];
    }
    m1 *= 1.0f / (float)C; m2 *= 1.0f / (float)C;
    if (ok) {
      float* dx = L.dx[t] + off;
      const float* add = L.dx_add[t] ? L.dx_add[t] + off : nullptr;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float r = rstd * (gam[c] * d[c] - m1 - h[c] * m2);
        if (add) r = fmaf(L.dx_add_scale, __ldg(add + (size_t)c * S), r);
        dx[(size_t)c * S] = r;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float a = warp_sum(ag[c]), bb = warp_sum(ab[c]);
    if (lane == 0) { sg[w][c] = a; sg[w][C + c] = bb; }
  }
  __syncthreads();
  if (threadIdx.x < 2 * C) {
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) sum += sg[k][threadIdx.x];
    float* dst = threadIdx.x < C ? L.dgamma[t] : L.dbeta[t];
    if (dst) atomicAdd(dst + (threadIdx.x < C ? threadIdx.x : threadIdx.x - C), sum);
  }
}

int ln_backward(const LnBwdBatch& L, cudaStream_t stream) {
  if (L.n <= 0) return VX_OK;
  prof_bytes(4.0 * L.n * L.B * (double)L.S * (3.0 * L.C + 1 + (L.dx_add[0] ? L.C : 0)));
  static const bool narrow_on = !(getenv("VX_LN_NARROW") && atoi(getenv("VX_LN_NARROW")) == 0);      // A/B probe
  if (narrow_on && (L.C == 8 || L.C == 16) && L.S >= 1024) {
    const int nblk = cdiv(L.S, 32);
    // 32-voxel blocks per warp: as many as leave >= one CTA per SM (VX_LN_BPW overrides for probes; the block count makes no
    // measurable difference between 112 and 432 CTAs -- isolated: 13.3 us against 30.4 us for the wide kernel at level 1,
    // profiles/r2z_ln_narrow.digest.txt / r2z_ln_wide.digest.txt)
    int bpw = 1;
    while (bpw < 8 && (long long)cdiv(nblk, 8 * bpw * 2) * L.B * L.n >= kSMs) bpw *= 2;
    if (const char* e = getenv("VX_LN_BPW")) bpw = atoi(e) > 0 ? atoi(e) : bpw;
    const dim3 grid(cdiv(nblk, 8 * bpw), L.B, L.n);
    if (L.C == 16) VX_LAUNCH(ln_bwd_narrow_kernel<16>, grid, dim3(256), 0, stream, L, bpw);
    else VX_LAUNCH(ln_bwd_narrow_kernel<8>, grid, dim3(256), 0, stream, L, bpw);
    return check_launch("ln_bwd_narrow_kernel");
  }
  int groups = 1;      // 32-voxel blocks per CTA: as many as leave >= ~2 CTAs per SM
  while (groups < 8 && (long long)cdiv(L.S, 32 * groups * 2) * L.B * L.n >= 2 * kSMs) groups *= 2;
  VX_LAUNCH(ln_bwd_kernel, dim3(cdiv(L.S, 32 * groups), L.B, L.n), dim3(256), 0, stream, L, groups);
  return check_launch("ln_bwd_kernel");
}

}  // namespace vx
