// Deep-supervision segmentation loss: for every deep output i, CrossEntropy + MONAI DiceLoss
// (include_background=False, to_onehot_y=True, softmax=True, smooth 1e-5, mean over (B, C-1)), summed with weights w_i.
// Reference: utils/loss.py:30-48 (per-output CE + Dice), utils/runtime.py:125-144 (weights).  The reference runs ~12
// element-wise / reduction kernels per output and direction over (B, C, D, H, W) logits; here one pass computes the
// softmax once per voxel and feeds both terms, partial sums are reduced in a fixed order (deterministic), and the
// backward pass recomputes the softmax instead of saving probabilities.
#include "vx_kernels.h"

#ifdef VX_EMU
#define __grid_constant__
#endif

namespace vx {

constexpr int SL_MAX_OUT = 8, SL_MAX_C = 8, SL_THREADS = 256;

struct SegLossArgs {
  const float* logits[SL_MAX_OUT];
  float* dlogits[SL_MAX_OUT];
  const long long* labels;     // (B, 1, S)
  float* part;                 // (n_out, B, nblk, 1 + 3 * C): ce, then (I, P, T) per class
  float* sums;                 // (n_out, B, 1 + 3 * C)
  float* loss;                 // (1)
  const float* dloss;          // (1)
  float w[SL_MAX_OUT];
  int n_out, B, C, S, nblk;
};

template <int C>
__global__ void __launch_bounds__(SL_THREADS) segloss_partial_kernel(const __grid_constant__ SegLossArgs A) {
  VX_PDL_ENTRY();
  __shared__ float red[33];
  const int i = blockIdx.z, b = blockIdx.y, blk = blockIdx.x;
  const float* lg = A.logits[i] + (size_t)b * C * A.S;
  const long long* lab = A.labels + (size_t)b * A.S;
  float ce = 0.f, I[C], P[C], T[C];
#pragma unroll
  for (int c = 0; c < C; ++c) { I[c] = 0.f; P[c] = 0.f; T[c] = 0.f; }
  auto voxel = [&](const float (&l)[C], int y) {
    float m = l[0];
#pragma unroll
    for (int c = 1; c < C; ++c) m = fmaxf(m, l[c]);
    float e[C], Z = 0.f, ly = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { e[c] = expf(l[c] - m); Z += e[c]; if (c == y) ly = l[c]; }
    const float inv = 1.0f / Z;
    ce += m + logf(Z) - ly;
#pragma unroll
    for (int c = 1; c < C; ++c) {
      const float p = e[c] * inv, t = (c == y) ? 1.f : 0.f;
      I[c] += p * t; P[c] += p; T[c] += t;
    }
  };
  // 4 voxels per iteration (16-byte loads of every class plane, labels as two 16-byte loads): four independent
  // exp / log chains in flight per thread instead of one
  const bool vec = (A.S & 3) == 0 && (((uintptr_t)lg | (uintptr_t)lab) & 15) == 0;
  if (vec) {
    for (int v = (blk * SL_THREADS + threadIdx.x) * 4; v < A.S; v += gridDim.x * SL_THREADS * 4) {
      float4 q[C];
#pragma unroll
      for (int c = 0; c < C; ++c) q[c] = __ldg(reinterpret_cast<const float4*>(lg + (size_t)c * A.S + v));
      const longlong2 y01 = __ldg(reinterpret_cast<const longlong2*>(lab + v));
      const longlong2 y23 = __ldg(reinterpret_cast<const longlong2*>(lab + v + 2));
      float l0[C], l1[C], l2[C], l3[C];
#pragma unroll
      for (int c = 0; c < C; ++c) { l0[c] = q[c].x; l1[c] = q[c].y; l2[c] = q[c].z; l3[c] = q[c].w; }
      voxel(l0, (int)y01.x); voxel(l1, (int)y01.y); voxel(l2, (int)y23.x); voxel(l3, (int)y23.y);
    }
  } else {
    for (int v = blk * SL_THREADS + threadIdx.x; v < A.S; v += gridDim.x * SL_THREADS) {
      float l[C];
#pragma unroll
      for (int c = 0; c < C; ++c) l[c] = __ldg(lg + (size_t)c * A.S + v);
      voxel(l, (int)__ldg(lab + v));
    }
  }
  float* out = A.part + (((size_t)i * A.B + b) * A.nblk + blk) * (1 + 3 * C);
  ce = block_sum(ce, red);
  if (threadIdx.x == 0) out[0] = ce;
#pragma unroll
  for (int c = 1; c < C; ++c) {
    const float a = block_sum(I[c], red), p = block_sum(P[c], red), t = block_sum(T[c], red);
    if (threadIdx.x == 0) { out[1 + 3 * c] = a; out[2 + 3 * c] = p; out[3 + 3 * c] = t; }
  }
}

// one CTA: fixed-order reduction of the partials, then the scalar loss
__global__ void __launch_bounds__(SL_THREADS) segloss_finalize_kernel(const __grid_constant__ SegLossArgs A) {
  VX_PDL_ENTRY();
  __shared__ float sloss[SL_THREADS];
  const int C = A.C, W = 1 + 3 * C;
  float acc = 0.f;
  for (int r = threadIdx.x; r < A.n_out * A.B; r += SL_THREADS) {
    const int i = r / A.B;
    const float* p = A.part + (size_t)r * A.nblk * W;
    float* s = A.sums + (size_t)r * W;
    for (int k = 0; k < W; ++k) {
      if (k >= 1 && k < 4) { s[k] = 0.f; continue; }      // background class is excluded
      double t = 0.0;
      for (int q = 0; q < A.nblk; ++q) t += (double)p[(size_t)q * W + k];
      s[k] = (float)t;
    }
    float dice = 0.f;
    for (int c = 1; c < C; ++c) dice += 1.0f - (2.0f * s[1 + 3 * c] + 1e-5f) / (s[2 + 3 * c] + s[3 + 3 * c] + 1e-5f);
    acc += A.w[i] * (s[0] / ((float)A.B * (float)A.S) + dice / ((float)A.B * (float)(C - 1)));
  }
  sloss[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < SL_THREADS; ++k) t += sloss[k];
    A.loss[0] = t;
  }
}

// dlogit_k = w_i dL [ (p_k - t_k) / (B S) + p_k (g_k - sum_c p_c g_c) ],  g_c = -(2 t_c den_c - (2 I_c + eps)) / den_c^2 / (B (C-1))
template <int C>
__global__ void __launch_bounds__(SL_THREADS) segloss_bwd_kernel(const __grid_constant__ SegLossArgs A) {
  VX_PDL_ENTRY();
  const int i = blockIdx.z, b = blockIdx.y;
  const float* lg = A.logits[i] + (size_t)b * C * A.S;
  float* dl = A.dlogits[i] + (size_t)b * C * A.S;
  const long long* lab = A.labels + (size_t)b * A.S;
  const float* s = A.sums + ((size_t)i * A.B + b) * (1 + 3 * C);
  const float up = A.w[i] * __ldg(A.dloss);
  const float kce = up / ((float)A.B * (float)A.S), kd = up / ((float)A.B * (float)(C - 1));
  float num[C], rden[C];
#pragma unroll
  for (int c = 1; c < C; ++c) {
    const float den = s[2 + 3 * c] + s[3 + 3 * c] + 1e-5f;
    rden[c] = 1.0f / den;
    num[c] = (2.0f * s[1 + 3 * c] + 1e-5f) * rden[c] * rden[c];       // (2I + eps) / den^2
  }
  auto voxel = [&](const float (&l)[C], int y, float (&d)[C]) {
    float m = l[0];
#pragma unroll
    for (int c = 1; c < C; ++c) m = fmaxf(m, l[c]);
    float p[C], Z = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { p[c] = expf(l[c] - m); Z += p[c]; }
    const float inv = 1.0f / Z;
    float g[C], sg = 0.f;
    g[0] = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) p[c] *= inv;
#pragma unroll
    for (int c = 1; c < C; ++c) {
      g[c] = kd * (num[c] - ((c == y) ? 2.0f * rden[c] : 0.f));
      sg = fmaf(p[c], g[c], sg);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) d[c] = kce * (p[c] - ((c == y) ? 1.f : 0.f)) + p[c] * (g[c] - sg);
  };
  const bool vec = (A.S & 3) == 0 && (((uintptr_t)lg | (uintptr_t)lab | (uintptr_t)dl) & 15) == 0;
  if (vec) {
    for (int v = (blockIdx.x * SL_THREADS + threadIdx.x) * 4; v < A.S; v += gridDim.x * SL_THREADS * 4) {
      float4 q[C];
#pragma unroll
      for (int c = 0; c < C; ++c) q[c] = __ldg(reinterpret_cast<const float4*>(lg + (size_t)c * A.S + v));
      const longlong2 y01 = __ldg(reinterpret_cast<const longlong2*>(lab + v));
      const longlong2 y23 = __ldg(reinterpret_cast<const longlong2*>(lab + v + 2));
      float l0[C], l1[C], l2[C], l3[C], d0[C], d1[C], d2[C], d3[C];
#pragma unroll
      for (int c = 0; c < C; ++c) { l0[c] = q[c].x; l1[c] = q[c].y; l2[c] = q[c].z; l3[c] = q[c].w; }
      voxel(l0, (int)y01.x, d0); voxel(l1, (int)y01.y, d1); voxel(l2, (int)y23.x, d2); voxel(l3, (int)y23.y, d3);
#pragma unroll
      for (int c = 0; c < C; ++c) *reinterpret_cast<float4*>(dl + (size_t)c * A.S + v) = make_float4(d0[c], d1[c], d2[c], d3[c]);
    }
  } else {
    for (int v = blockIdx.x * SL_THREADS + threadIdx.x; v < A.S; v += gridDim.x * SL_THREADS) {
      float l[C], d[C];
#pragma unroll
      for (int c = 0; c < C; ++c) l[c] = __ldg(lg + (size_t)c * A.S + v);
      voxel(l, (int)__ldg(lab + v), d);
#pragma unroll
      for (int c = 0; c < C; ++c) dl[(size_t)c * A.S + v] = d[c];
    }
  }
}

static int segloss_args(const vx_segloss_desc* d, SegLossArgs& A) {
  if (!d || d->n_out <= 0 || d->n_out > SL_MAX_OUT || d->B <= 0 || d->S <= 0 || d->C < 2 || d->C > SL_MAX_C) {
    set_error("segloss: bad descriptor"); return VX_ERR_BAD_DESC;
  }
  if (d->C != 2 && d->C != 3 && d->C != 4) { set_error("segloss: %d classes not instantiated (2, 3, 4)", d->C); return VX_ERR_UNSUPPORTED; }
  A.n_out = d->n_out; A.B = d->B; A.C = d->C; A.S = d->S;
  int nblk = cdiv(d->S, SL_THREADS * 8);
  const int cap = cdiv(4 * kSMs, d->B * d->n_out);
  if (nblk > cap) nblk = cap;
  if (nblk < 1) nblk = 1;
  A.nblk = nblk;
  for (int i = 0; i < d->n_out; ++i) A.w[i] = d->weights[i];
  return VX_OK;
}

}  // namespace vx

using namespace vx;

extern "C" size_t vx_segloss_workspace(const vx_segloss_desc* d) {
  SegLossArgs A{};
  if (segloss_args(d, A) != VX_OK) return 0;
  return sizeof(float) * (size_t)A.n_out * A.B * A.nblk * (1 + 3 * A.C);
}

extern "C" int vx_segloss_fwd(const vx_segloss_desc* d, const void* const* in, void* const* out, void* workspace,
                              size_t workspace_bytes, vx_stream_t stream) {
  SegLossArgs A{};
  int rc = segloss_args(d, A);
  if (rc != VX_OK) return rc;
  if (!workspace || workspace_bytes < vx_segloss_workspace(d)) { set_error("segloss_fwd: workspace too small"); return VX_ERR_WORKSPACE; }
  prof_scope("segloss_fwd n%d B%d C%d S%d", d->n_out, d->B, d->C, d->S);
  for (int i = 0; i < A.n_out; ++i) A.logits[i] = (const float*)in[i];
  A.labels = (const long long*)in[A.n_out];
  A.loss = (float*)out[0];
  A.sums = (float*)out[1];
  A.part = (float*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(A.nblk, A.B, A.n_out);
  prof_bytes((double)A.n_out * A.B * A.S * (4.0 * A.C + 8.0));
  if (A.C == 2) VX_LAUNCH(segloss_partial_kernel<2>, grid, dim3(SL_THREADS), 0, st, A);
  else if (A.C == 3) VX_LAUNCH(segloss_partial_kernel<3>, grid, dim3(SL_THREADS), 0, st, A);
  else VX_LAUNCH(segloss_partial_kernel<4>, grid, dim3(SL_THREADS), 0, st, A);
  rc = check_launch("segloss_partial_kernel");
  if (rc != VX_OK) return rc;
  VX_LAUNCH(segloss_finalize_kernel, dim3(1), dim3(SL_THREADS), 0, st, A);
  return check_launch("segloss_finalize_kernel");
}

extern "C" int vx_segloss_bwd(const vx_segloss_desc* d, const void* const* in, void* const* out, vx_stream_t stream) {
  SegLossArgs A{};
  int rc = segloss_args(d, A);
  if (rc != VX_OK) return rc;
  prof_scope("segloss_bwd n%d B%d C%d S%d", d->n_out, d->B, d->C, d->S);
  A.dloss = (const float*)in[0];
  for (int i = 0; i < A.n_out; ++i) { A.logits[i] = (const float*)in[1 + i]; A.dlogits[i] = (float*)out[i]; }
  A.labels = (const long long*)in[1 + A.n_out];
  A.sums = (float*)in[2 + A.n_out];
  int nblk = cdiv(A.S, SL_THREADS * 4);
  const int cap = cdiv(8 * kSMs, A.B * A.n_out);
  if (nblk > cap) nblk = cap;
  dim3 grid(nblk, A.B, A.n_out);
  cudaStream_t st = (cudaStream_t)stream;
  prof_bytes((double)A.n_out * A.B * A.S * (8.0 * A.C + 8.0));
  if (A.C == 2) VX_LAUNCH(segloss_bwd_kernel<2>, grid, dim3(SL_THREADS), 0, st, A);
  else if (A.C == 3) VX_LAUNCH(segloss_bwd_kernel<3>, grid, dim3(SL_THREADS), 0, st, A);
  else VX_LAUNCH(segloss_bwd_kernel<4>, grid, dim3(SL_THREADS), 0, st, A);
  return check_launch("segloss_bwd_kernel");
}
