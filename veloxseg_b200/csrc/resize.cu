// Trilinear resize with align_corners=True (reference: VeloxSeg.scale_prediction, model/VeloxSeg.py:177-184, applied to
// every deep-supervision output in training, VeloxSeg.py:200-202).  fp32, NCDHW.
//
// Forward : thread per output voxel, 8 taps from the (tiny, cache-resident) source.
// Backward: the adjoint is separable.  Three 1-D passes (W, then H, then D), each a gather over the output positions
//           whose interpolation support contains the input sample: deterministic, no atomics.  ATen's backward
//           scatters with atomicAdd; for a 3^3 -> 96^3 resize that is 884 736 atomics onto 27 addresses per plane.
#include "vx_kernels.h"

#ifdef VX_EMU
#define __grid_constant__
#endif

namespace vx {

VX_DEV void lerp_ac(int p, int n, int out, int& i0, int& i1, float& w1) {
  // F.interpolate(align_corners=True): src = p * (n-1)/(out-1)
  if (out <= 1 || n <= 1) { i0 = 0; i1 = 0; w1 = 0.f; return; }
  const float scale = (float)(n - 1) / (float)(out - 1);
  const float src = scale * (float)p;
  i0 = (int)src;
  if (i0 > n - 1) i0 = n - 1;
  i1 = i0 < n - 1 ? i0 + 1 : i0;
  w1 = src - (float)i0;
}

struct ResizeArgs { const float* x; float* y; int planes, d, h, w, D, H, W; };

__global__ void __launch_bounds__(256) resize_fwd_kernel(const __grid_constant__ ResizeArgs A) {
  const long long total = (long long)A.planes * A.D * A.H * A.W;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(e % A.W), Y = (int)((e / A.W) % A.H), Z = (int)((e / ((long long)A.W * A.H)) % A.D);
    const long long pl = e / ((long long)A.W * A.H * A.D);
    int a0, a1, b0, b1, c0, c1;
    float wa, wb, wc;
    lerp_ac(Z, A.d, A.D, a0, a1, wa);
    lerp_ac(Y, A.h, A.H, b0, b1, wb);
    lerp_ac(X, A.w, A.W, c0, c1, wc);
    const float* p = A.x + pl * A.d * A.h * A.w;
#define VX_R(a, b, c) __ldg(p + ((size_t)(a) * A.h + (b)) * A.w + (c))
    const float ua = 1.f - wa, ub = 1.f - wb, uc = 1.f - wc;
    A.y[e] = ua * (ub * (uc * VX_R(a0, b0, c0) + wc * VX_R(a0, b0, c1)) + wb * (uc * VX_R(a0, b1, c0) + wc * VX_R(a0, b1, c1))) +
             wa * (ub * (uc * VX_R(a1, b0, c0) + wc * VX_R(a1, b0, c1)) + wb * (uc * VX_R(a1, b1, c0) + wc * VX_R(a1, b1, c1)));
#undef VX_R
  }
}

// out[o, j, i] = sum_p weight(j <- p) * in[o, p, i]      in: (outer, P, inner)   out: (outer, n, inner)
__global__ void __launch_bounds__(256) resize_adjoint1d_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                               long long outer, int P, int n, int inner) {
  const long long total = outer * n * inner;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e % inner), j = (int)((e / inner) % n);
    const long long o = e / ((long long)inner * n);
    int lo = 0, hi = P - 1;
    if (n > 1 && P > 1) {      // positions whose source coordinate lies in (j-1, j+1)
      const float inv = (float)(P - 1) / (float)(n - 1);
      lo = (int)floorf((float)(j - 1) * inv) - 1;
      hi = (int)ceilf((float)(j + 1) * inv) + 1;
      if (lo < 0) lo = 0;
      if (hi > P - 1) hi = P - 1;
    }
    const float* src = in + (o * P) * inner + i;
    float acc = 0.f;
    for (int p = lo; p <= hi; ++p) {
      int i0, i1; float w1;
      lerp_ac(p, n, P, i0, i1, w1);
      float wgt = 0.f;
      if (i0 == j) wgt += 1.f - w1;
      if (i1 == j) wgt += w1;
      if (wgt != 0.f) acc = fmaf(wgt, __ldg(src + (size_t)p * inner), acc);
    }
    out[e] = acc;
  }
}

static int launch_adjoint(const float* in, float* out, long long outer, int P, int n, int inner, cudaStream_t st) {
  const long long total = outer * n * inner;
  int blocks = cdiv(total, 256);
  if (blocks > kSMs * 32) blocks = kSMs * 32;
  if (blocks < 1) blocks = 1;
  VX_LAUNCH(resize_adjoint1d_kernel, dim3(blocks), dim3(256), 0, st, in, out, outer, P, n, inner);
  return check_launch("resize_adjoint1d_kernel");
}

}  // namespace vx

using namespace vx;

static int resize_check(const vx_resize_desc* d) {
  if (!d || d->planes <= 0 || d->d <= 0 || d->h <= 0 || d->w <= 0 || d->D <= 0 || d->H <= 0 || d->W <= 0) {
    set_error("resize: bad descriptor");
    return VX_ERR_BAD_DESC;
  }
  return VX_OK;
}

extern "C" size_t vx_resize_workspace(const vx_resize_desc* d) {
  if (resize_check(d) != VX_OK) return 0;
  // backward temporaries: (planes, D, H, w) and (planes, D, h, w)
  const size_t t1 = (size_t)d->planes * d->D * d->H * d->w, t2 = (size_t)d->planes * d->D * d->h * d->w;
  return ((t1 * 4 + 255) & ~(size_t)255) + ((t2 * 4 + 255) & ~(size_t)255);
}

extern "C" int vx_resize_trilinear_fwd(const vx_resize_desc* d, const void* const* in, void* const* out, vx_stream_t stream) {
  int rc = resize_check(d);
  if (rc != VX_OK) return rc;
  prof_scope("resize_fwd P%d %dx%dx%d->%dx%dx%d", d->planes, d->d, d->h, d->w, d->D, d->H, d->W);
  ResizeArgs A{(const float*)in[0], (float*)out[0], d->planes, d->d, d->h, d->w, d->D, d->H, d->W};
  const long long total = (long long)d->planes * d->D * d->H * d->W;
  int blocks = cdiv(total, 256);
  if (blocks > kSMs * 32) blocks = kSMs * 32;
  VX_LAUNCH(resize_fwd_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, A);
  return check_launch("resize_fwd_kernel");
}

extern "C" int vx_resize_trilinear_bwd(const vx_resize_desc* d, const void* const* in, void* const* out, void* workspace,
                                       size_t workspace_bytes, vx_stream_t stream) {
  int rc = resize_check(d);
  if (rc != VX_OK) return rc;
  const size_t need = vx_resize_workspace(d);
  if (!workspace || workspace_bytes < need) { set_error("resize_bwd: workspace %zu < %zu", workspace_bytes, need); return VX_ERR_WORKSPACE; }
  prof_scope("resize_bwd P%d %dx%dx%d->%dx%dx%d", d->planes, d->d, d->h, d->w, d->D, d->H, d->W);
  cudaStream_t st = (cudaStream_t)stream;
  const float* dy = (const float*)in[0];
  float* dx = (float*)out[0];
  const size_t t1 = (size_t)d->planes * d->D * d->H * d->w;
  float* tmp1 = (float*)workspace;
  float* tmp2 = (float*)((char*)workspace + ((t1 * 4 + 255) & ~(size_t)255));
  rc = launch_adjoint(dy, tmp1, (long long)d->planes * d->D * d->H, d->W, d->w, 1, st);              // W -> w
  if (rc != VX_OK) return rc;
  rc = launch_adjoint(tmp1, tmp2, (long long)d->planes * d->D, d->H, d->h, d->w, st);                // H -> h
  if (rc != VX_OK) return rc;
  return launch_adjoint(tmp2, dx, (long long)d->planes, d->D, d->d, d->h * d->w, st);                // D -> d
}
