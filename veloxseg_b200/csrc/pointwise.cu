// Channel contractions (1x1x1 convs) and the norm kernels they are paired with.  fp32, NCDHW.
//
// pw_kernel    : voxel-parallel.  One CTA = 256 consecutive voxels x 16 output channels; a thread owns two voxels
//                and 16 accumulators each, weights are broadcast from shared memory as float4.  HBM side: every
//                load/store is a 128-byte-coalesced run along the voxel axis.
// pw_wgrad     : split-K over voxel chunks; chunk tiles are transposed through shared memory and reduced as
//                4x4 register outer products, partial results are folded with fp32 atomics.
#include "vx_kernels.h"

#ifdef VX_EMU
#define __grid_constant__
#endif

namespace vx {

constexpr int PW_THREADS = 128;
constexpr int PW_TV = 256;
constexpr int PW_CO = 16;

VX_DEV float pw_weight(const PwProblem& P, int co, int ci) {
  int off = 0;
  if (!P.transposed) {
    for (int s = 0; s < P.nseg; ++s) {
      if (co < off + P.seg[s].n) return P.seg[s].W[(size_t)(co - off) * P.seg[s].ld + ci];
      off += P.seg[s].n;
    }
  } else {
    for (int s = 0; s < P.nseg; ++s) {
      if (ci < off + P.seg[s].n) return P.seg[s].W[(size_t)(ci - off) * P.seg[s].ld + co];
      off += P.seg[s].n;
    }
  }
  return 0.f;
}

VX_DEV float pw_bias(const PwProblem& P, int co) {
  if (P.transposed) return 0.f;
  int off = 0;
  for (int s = 0; s < P.nseg; ++s) {
    if (co < off + P.seg[s].n) return P.seg[s].bias ? P.seg[s].bias[co - off] : 0.f;
    off += P.seg[s].n;
  }
  return 0.f;
}

__global__ void __launch_bounds__(PW_THREADS) pw_kernel(const __grid_constant__ PwBatch batch) {
  const int pi = blockIdx.z / batch.B, b = blockIdx.z % batch.B;
  const PwProblem& P = batch.p[pi];
  const int S = batch.S, Ci = P.Ci, Co = P.Co;
  const int co0 = blockIdx.y * PW_CO;
  if (co0 >= Co) return;
  VX_DYN_SMEM(float, sm);
  float* Ws = sm;              // [Ci][16], prologue-folded
  float* bf = Ws + Ci * PW_CO; // [16] folded bias
  const int tid = threadIdx.x;

  for (int idx = tid; idx < Ci * PW_CO; idx += PW_THREADS) {
    const int ci = idx / PW_CO, j = idx % PW_CO, co = co0 + j;
    float w = co < Co ? pw_weight(P, co, ci) : 0.f;
    if (P.pro == PRO_AFFINE) w *= P.pro_a[b * P.pro_bstride + ci];
    Ws[idx] = w;
  }
  if (tid < PW_CO) {
    const int co = co0 + tid;
    float bb = 0.f;
    if (co < Co) {
      bb = pw_bias(P, co);
      if (P.pro == PRO_AFFINE)
        for (int ci = 0; ci < Ci; ++ci) bb += pw_weight(P, co, ci) * P.pro_c[b * P.pro_bstride + ci];
    }
    bf[tid] = bb;
  }
  __syncthreads();

  const int v0 = blockIdx.x * PW_TV + tid, v1 = v0 + PW_THREADS;
  const bool ok0 = v0 < S, ok1 = v1 < S;
  float acc0[PW_CO], acc1[PW_CO];
#pragma unroll
  for (int j = 0; j < PW_CO; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; }

  const bool pro_drop = P.pro == PRO_DROPOUT || P.pro == PRO_GELU_DROPOUT;
  const float pinv = pro_drop ? 1.0f / (1.0f - P.pro_drop_p) : 1.f;
  int cg = 0;
  for (int s = 0; s < P.nsrc; ++s) {
    const int Cs = P.src[s].C;
    const float* xp = P.src[s].ptr + (size_t)b * Cs * S;
    for (int c = 0; c < Cs; ++c, ++cg) {
      float x0 = ok0 ? __ldg(xp + (size_t)c * S + v0) : 0.f;
      float x1 = ok1 ? __ldg(xp + (size_t)c * S + v1) : 0.f;
      if (P.pro == PRO_GELU || P.pro == PRO_GELU_DROPOUT) { x0 = gelu_f(x0); x1 = gelu_f(x1); }
      if (pro_drop) {
        const uint64_t base = ((uint64_t)b * Ci + cg) * (uint64_t)S;
        x0 *= dropout_scale(P.pro_seed, P.pro_site, base + v0, P.pro_drop_p, pinv);
        x1 *= dropout_scale(P.pro_seed, P.pro_site, base + v1, P.pro_drop_p, pinv);
      }
      const float4* w4 = reinterpret_cast<const float4*>(Ws + cg * PW_CO);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 w = w4[q];
        acc0[4 * q + 0] = fmaf(w.x, x0, acc0[4 * q + 0]); acc1[4 * q + 0] = fmaf(w.x, x1, acc1[4 * q + 0]);
        acc0[4 * q + 1] = fmaf(w.y, x0, acc0[4 * q + 1]); acc1[4 * q + 1] = fmaf(w.y, x1, acc1[4 * q + 1]);
        acc0[4 * q + 2] = fmaf(w.z, x0, acc0[4 * q + 2]); acc1[4 * q + 2] = fmaf(w.z, x1, acc1[4 * q + 2]);
        acc0[4 * q + 3] = fmaf(w.w, x0, acc0[4 * q + 3]); acc1[4 * q + 3] = fmaf(w.w, x1, acc1[4 * q + 3]);
      }
    }
  }

  const float dinv = P.drop_p > 0.f ? 1.0f / (1.0f - P.drop_p) : 1.f;
  // output segment bookkeeping (forward orientation may write several tensors)
  int seg = 0, seg_off = 0;
#pragma unroll
  for (int j = 0; j < PW_CO; ++j) {
    const int co = co0 + j;
    if (co >= Co) break;
    float* outp;
    if (!P.transposed) {
      while (co >= seg_off + P.seg[seg].n) { seg_off += P.seg[seg].n; ++seg; }
      outp = P.seg[seg].out + ((size_t)b * P.seg[seg].n + (co - seg_off)) * S;
    } else {
      outp = P.seg[0].out + ((size_t)b * Co + co) * S;
    }
    const size_t lbase = ((size_t)b * Co + co) * S;   // index in the logical (B, Co, S) tensor
    float y0 = acc0[j] + bf[j], y1 = acc1[j] + bf[j];
    if (P.act == 1) { y0 = gelu_f(y0); y1 = gelu_f(y1); }
    if (P.mulgrad) {
      if (ok0) y0 *= gelu_grad_f(P.mulgrad[lbase + v0]);
      if (ok1) y1 *= gelu_grad_f(P.mulgrad[lbase + v1]);
    }
    if (P.drop_p > 0.f) {
      y0 *= dropout_scale(P.seed, P.site, lbase + v0, P.drop_p, dinv);
      y1 *= dropout_scale(P.seed, P.site, lbase + v1, P.drop_p, dinv);
    }
    if (P.res) {
      if (ok0) y0 = fmaf(P.res_scale, P.res[lbase + v0], y0);
      if (ok1) y1 = fmaf(P.res_scale, P.res[lbase + v1], y1);
    }
    if (P.res2) {
      if (ok0) y0 += P.res2[lbase + v0];
      if (ok1) y1 += P.res2[lbase + v1];
    }
    if (ok0) outp[v0] = y0;
    if (ok1) outp[v1] = y1;
  }
}

int pw_forward(const PwBatch& batch, cudaStream_t stream) {
  int maxCo = 0, maxCi = 0;
  for (int i = 0; i < batch.nprob; ++i) {
    const PwProblem& P = batch.p[i];
    maxCo = P.Co > maxCo ? P.Co : maxCo;
    maxCi = P.Ci > maxCi ? P.Ci : maxCi;
    int cs = 0;
    for (int s = 0; s < P.nsrc; ++s) cs += P.src[s].C;
    if (cs != P.Ci) { set_error("pw_forward: source channels %d != Ci %d", cs, P.Ci); return VX_ERR_BAD_DESC; }
    if (!P.transposed && P.nseg > 1 && (P.res || P.mulgrad || P.drop_p > 0.f || P.res2)) {
      set_error("pw_forward: epilogue tensors need a single output segment"); return VX_ERR_BAD_DESC;
    }
  }
  if (batch.nprob <= 0 || batch.B <= 0 || batch.S <= 0) return VX_OK;
  const size_t smem = (size_t)(maxCi * PW_CO + PW_CO) * sizeof(float);
  VX_SET_SMEM(pw_kernel, smem);
  dim3 grid(cdiv(batch.S, PW_TV), cdiv(maxCo, PW_CO), batch.nprob * batch.B);
  VX_LAUNCH(pw_kernel, grid, dim3(PW_THREADS), smem, stream, batch);
  return check_launch("pw_kernel");
}

// ---------------------------------------------------------------------------------------------------
// weight gradient
// ---------------------------------------------------------------------------------------------------
constexpr int WG_THREADS = 256;

__global__ void __launch_bounds__(WG_THREADS) pw_wgrad_kernel(const __grid_constant__ WgBatch batch, int TV) {
  const WgProblem& P = batch.p[blockIdx.z];
  const int b = blockIdx.y, S = batch.S;
  const int vbase = blockIdx.x * TV;
  const int Co = P.Co, Ci = P.Ci;
  const int Co4 = (Co + 3) & ~3, Ci4 = (Ci + 1 + 3) & ~3;    // +1: the all-ones row that yields db
  const int CoP = Co4 + 4, CiP = Ci4 + 4;
  VX_DYN_SMEM(float, sm);
  float* sY = sm;                     // [TV][CoP]
  float* sX = sm + (size_t)TV * CoP;  // [TV][CiP]
  const int tid = threadIdx.x;

  const float yinv = P.y_drop_p > 0.f ? 1.0f / (1.0f - P.y_drop_p) : 1.f;
  for (int idx = tid; idx < Co4 * TV; idx += WG_THREADS) {
    const int co = idx / TV, v = idx % TV, gv = vbase + v;
    float val = 0.f;
    if (co < Co && gv < S) {
      const size_t gi = ((size_t)b * Co + co) * S + gv;
      val = __ldg(P.dY + gi);
      if (P.y_drop_p > 0.f) val *= dropout_scale(P.y_seed, P.y_site, gi, P.y_drop_p, yinv);
    }
    sY[v * CoP + co] = val;
  }
  {
    int cg0 = 0;
    for (int s = 0; s < P.nsrc; ++s) {
      const int Cs = P.src[s].C;
      const float* xp = P.src[s].ptr + (size_t)b * Cs * S;
      for (int idx = tid; idx < Cs * TV; idx += WG_THREADS) {
        const int c = idx / TV, v = idx % TV, gv = vbase + v;
        float val = 0.f;
        if (gv < S) {
          val = __ldg(xp + (size_t)c * S + gv);
          if (P.xpro == PRO_AFFINE) {
            const int k = b * P.x_bstride + cg0 + c;
            val = fmaf(val, P.xa[k], P.xc[k]);
          } else if (P.xpro == PRO_GELU) {
            val = gelu_f(val);
          } else if (P.xpro == PRO_GELU_DROPOUT) {
            val = gelu_f(val) * dropout_scale(P.x_seed, P.x_site, ((uint64_t)b * Ci + cg0 + c) * (uint64_t)S + gv, P.x_drop_p,
                                              1.0f / (1.0f - P.x_drop_p));
          }
        }
        sX[v * CiP + cg0 + c] = val;
      }
      cg0 += Cs;
    }
    for (int idx = tid; idx < (Ci4 - Ci) * TV; idx += WG_THREADS) {
      const int c = Ci + idx / TV, v = idx % TV;
      sX[v * CiP + c] = (c == Ci && vbase + v < S) ? 1.f : 0.f;
    }
  }
  __syncthreads();

  const int nCo4 = Co4 >> 2, nCi4 = Ci4 >> 2, ntiles = nCo4 * nCi4;
  // When there are fewer 4x4 tiles than threads, split the voxel range between thread groups.
  int G = 1;
  while (G * 2 * ntiles <= WG_THREADS && (TV / (G * 2)) >= 8) G *= 2;
  const int span = TV / G;
  for (int t = tid; t < ntiles * G; t += WG_THREADS) {
    const int tile = t % ntiles, grp = t / ntiles;
    const int co4 = tile % nCo4, ci4 = tile / nCo4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const float* py = sY + (size_t)grp * span * CoP + co4 * 4;
    const float* px = sX + (size_t)grp * span * CiP + ci4 * 4;
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
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int co = co4 * 4 + i;
      if (co >= Co) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ci = ci4 * 4 + j;
        if (ci < Ci) atomicAdd(P.dW + (size_t)co * P.ld + ci, acc[i][j]);
        else if (ci == Ci && P.db) atomicAdd(P.db + co, acc[i][j]);
      }
    }
  }
}

int pw_wgrad(const WgBatch& batch, cudaStream_t stream) {
  if (batch.nprob <= 0) return VX_OK;
  int maxsum = 0;
  for (int i = 0; i < batch.nprob; ++i) {
    const WgProblem& P = batch.p[i];
    int cs = 0;
    for (int s = 0; s < P.nsrc; ++s) cs += P.src[s].C;
    if (cs != P.Ci) { set_error("pw_wgrad: source channels %d != Ci %d", cs, P.Ci); return VX_ERR_BAD_DESC; }
    const int tot = ((P.Co + 3) & ~3) + 4 + ((P.Ci + 4) & ~3) + 4;
    maxsum = tot > maxsum ? tot : maxsum;
  }
  int TV = 512;
  while (TV > 32 && (size_t)TV * maxsum * sizeof(float) > 96 * 1024) TV >>= 1;
  while (TV > 32 && TV / 2 >= batch.S) TV >>= 1;
  const size_t smem = (size_t)TV * maxsum * sizeof(float);
  if (smem > 200 * 1024) { set_error("pw_wgrad: channel count too large (%d)", maxsum); return VX_ERR_UNSUPPORTED; }
  VX_SET_SMEM(pw_wgrad_kernel, smem);
  dim3 grid(cdiv(batch.S, TV), batch.B, batch.nprob);
  VX_LAUNCH(pw_wgrad_kernel, grid, dim3(WG_THREADS), smem, stream, batch, TV);
  return check_launch("pw_wgrad_kernel");
}

// ---------------------------------------------------------------------------------------------------
// InstanceNorm over rows (one CTA per (b,c) row; rows are <= 16 K elements on this path, L1/L2 resident)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) inorm_rows_fwd_kernel(const float* __restrict__ x, const float* __restrict__ addend,
                                                             float* __restrict__ y, float* __restrict__ stats, int S,
                                                             float eps) {
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

int inorm_rows_fwd(const float* x, const float* addend, float* y, float* stats, int rows, int S, float eps,
                   cudaStream_t stream) {
  if (rows <= 0) return VX_OK;
  VX_LAUNCH(inorm_rows_fwd_kernel, dim3(rows), dim3(256), 0, stream, x, addend, y, stats, S, eps);
  return check_launch("inorm_rows_fwd_kernel");
}

__global__ void __launch_bounds__(256) inorm_rows_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                             const float* __restrict__ stats,
                                                             const float* __restrict__ dx_add, float* __restrict__ dx,
                                                             int S) {
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
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    const float xh = (x[base + i] - mean) * rstd;
    float v = rstd * (dy[base + i] - m1 - xh * m2);
    if (dx_add) v += dx_add[base + i];
    dx[base + i] = v;
  }
}

int inorm_rows_bwd(const float* dy, const float* x, const float* stats, const float* dx_add, float* dx, int rows,
                   int S, cudaStream_t stream) {
  if (rows <= 0) return VX_OK;
  VX_LAUNCH(inorm_rows_bwd_kernel, dim3(rows), dim3(256), 0, stream, dy, x, stats, dx_add, dx, S);
  return check_launch("inorm_rows_bwd_kernel");
}

__global__ void stats_to_affine_kernel(const float* __restrict__ stats, float* __restrict__ a, float* __restrict__ c,
                                       int rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows) { const float m = stats[2 * i], r = stats[2 * i + 1]; a[i] = r; c[i] = -m * r; }
}

int stats_to_affine(const float* stats, float* a, float* c, int rows, cudaStream_t stream) {
  VX_LAUNCH(stats_to_affine_kernel, dim3(cdiv(rows, 128)), dim3(128), 0, stream, stats, a, c, rows);
  return check_launch("stats_to_affine_kernel");
}

// ---------------------------------------------------------------------------------------------------
// channel-first LayerNorm (thread per voxel, channel loop strides by S so every warp access is coalesced)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ln_fwd_kernel(const __grid_constant__ LnBatch L) {
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

int ln_forward(const LnBatch& L, cudaStream_t stream) {
  if (L.n <= 0) return VX_OK;
  VX_LAUNCH(ln_fwd_kernel, dim3(cdiv(L.S, 128), L.B, L.n), dim3(128), 0, stream, L);
  return check_launch("ln_fwd_kernel");
}

__global__ void __launch_bounds__(128) ln_bwd_kernel(const __grid_constant__ LnBwdBatch L) {
  const int t = blockIdx.z, b = blockIdx.y;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  const int C = L.C, S = L.S;
  const bool ok = v < S;
  VX_DYN_SMEM(float, acc);   // [2C]: dgamma, dbeta partials of this CTA
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const size_t off = (size_t)b * C * S + (ok ? v : 0);
  const float* dout = L.dout[t] + off;
  const float* xh = L.xhat[t] + off;
  const float* gamma = L.gamma[t];
  const int lane = threadIdx.x & 31;
  float m1 = 0.f, m2 = 0.f;
  for (int c = 0; c < C; ++c) {
    const float d = ok ? dout[(size_t)c * S] : 0.f;
    const float h = ok ? xh[(size_t)c * S] : 0.f;
    const float g = gamma[c] * d;
    m1 += g;
    m2 = fmaf(g, h, m2);
    const float sg = warp_sum(d * h), sb = warp_sum(d);
    if (lane == 0) { atomicAdd(acc + c, sg); atomicAdd(acc + C + c, sb); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    if (L.dgamma[t]) atomicAdd(L.dgamma[t] + i, acc[i]);
    if (L.dbeta[t]) atomicAdd(L.dbeta[t] + i, acc[C + i]);
  }
  if (!ok) return;
  m1 /= (float)C; m2 /= (float)C;
  const float rstd = L.rstd[t][(size_t)b * S + v];
  float* dx = L.dx[t] + off;
  const float* add = L.dx_add[t] ? L.dx_add[t] + off : nullptr;
  for (int c = 0; c < C; ++c) {
    const float g = gamma[c] * dout[(size_t)c * S];
    float r = rstd * (g - m1 - xh[(size_t)c * S] * m2);
    if (add) r = fmaf(L.dx_add_scale, add[(size_t)c * S], r);
    dx[(size_t)c * S] = r;
  }
}

int ln_backward(const LnBwdBatch& L, cudaStream_t stream) {
  if (L.n <= 0) return VX_OK;
  const size_t smem = (size_t)2 * L.C * sizeof(float);
  VX_LAUNCH(ln_bwd_kernel, dim3(cdiv(L.S, 128), L.B, L.n), dim3(128), smem, stream, L);
  return check_launch("ln_bwd_kernel");
}

}  // namespace vx
