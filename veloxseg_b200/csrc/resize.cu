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
  VX_PDL_ENTRY();
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

// Forward, W % 4 == 0: CTA = (8 output rows, Z, plane); the x-axis taps (index, weight) are tabulated once per CTA, a thread
// produces 4 consecutive outputs of one row and stores them as one float4 -- the per-voxel kernel above spends ~150
// instructions per output on index decoding and three float divisions; this one is bound by the 4-byte-per-voxel write.
constexpr int RS_ROWS = 8;
__global__ void __launch_bounds__(256) resize_fwd_rows_kernel(const __grid_constant__ ResizeArgs A) {
  VX_PDL_ENTRY();
  VX_DYN_SMEM(float, sm);
  int* c0t = reinterpret_cast<int*>(sm);       // [W]
  float* wct = sm + A.W;                       // [W]
  for (int X = threadIdx.x; X < A.W; X += blockDim.x) {
    int c0, c1; float wc;
    lerp_ac(X, A.w, A.W, c0, c1, wc);
    c0t[X] = c0 | (c1 << 16);
    wct[X] = wc;
  }
  __syncthreads();
  const int nq = A.W >> 2;
  const int Z = blockIdx.y, pl = blockIdx.z;
  int a0, a1; float wa;
  lerp_ac(Z, A.d, A.D, a0, a1, wa);
  const float* p = A.x + (size_t)pl * A.d * A.h * A.w;
  for (int it = threadIdx.x; it < RS_ROWS * nq; it += blockDim.x) {
    const int Y = blockIdx.x * RS_ROWS + it / nq, xq = it % nq;
    if (Y >= A.H) continue;
    int b0, b1; float wb;
    lerp_ac(Y, A.h, A.H, b0, b1, wb);
    const float* r00 = p + ((size_t)a0 * A.h + b0) * A.w;
    const float* r01 = p + ((size_t)a0 * A.h + b1) * A.w;
    const float* r10 = p + ((size_t)a1 * A.h + b0) * A.w;
    const float* r11 = p + ((size_t)a1 * A.h + b1) * A.w;
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int X = xq * 4 + i;
      const int cc = c0t[X], c0 = cc & 0xffff, c1 = cc >> 16;
      const float wc = wct[X], uc = 1.f - wc;
      // same association as the per-voxel kernel: x first, then y, then z
      const float v00 = uc * __ldg(r00 + c0) + wc * __ldg(r00 + c1), v01 = uc * __ldg(r01 + c0) + wc * __ldg(r01 + c1);
      const float v10 = uc * __ldg(r10 + c0) + wc * __ldg(r10 + c1), v11 = uc * __ldg(r11 + c0) + wc * __ldg(r11 + c1);
      o[i] = (1.f - wa) * ((1.f - wb) * v00 + wb * v01) + wa * ((1.f - wb) * v10 + wb * v11);
    }
    *reinterpret_cast<float4*>(A.y + (((size_t)pl * A.D + Z) * A.H + Y) * A.W + xq * 4) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// First adjoint pass (W -> w, inner == 1) for long rows: persistent CTAs; a block of 64 rows is staged in shared memory with
// coalesced loads (row stride P + 1: conflict-free for threads that own different rows), the weight table is built once
// per CTA, a thread owns one (row, j) output.  The generic kernel below reads a row per thread: 96-float strides.
constexpr int RA_ROWS = 64;
__global__ void __launch_bounds__(256) resize_adjoint_rows_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                                  long long rows, int P, int n) {
  VX_PDL_ENTRY();
  VX_DYN_SMEM(float, sm);
  float* wt = sm;                          // [n][P]
  int* lohi = reinterpret_cast<int*>(sm + (size_t)n * P);      // [n] lo | hi << 16
  float* tile = sm + (size_t)n * P + n;    // [RA_ROWS][P + 1]
  for (int e = threadIdx.x; e < n * P; e += blockDim.x) {
    const int j = e / P, p = e % P;
    int i0, i1; float w1;
    lerp_ac(p, n, P, i0, i1, w1);
    float wgt = 0.f;
    if (i0 == j) wgt += 1.f - w1;
    if (i1 == j) wgt += w1;
    wt[e] = wgt;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    int lo = P, hi = -1;
    for (int p = 0; p < P; ++p) if (wt[j * P + p] != 0.f) { lo = p < lo ? p : lo; hi = p; }
    lohi[j] = lo | (hi << 16);
  }
  __syncthreads();
  const long long nblk = (rows + RA_ROWS - 1) / RA_ROWS;
  for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const long long r0 = blk * RA_ROWS;
    const int nr = (int)((rows - r0) < RA_ROWS ? (rows - r0) : RA_ROWS);
    __syncthreads();
    for (int e = threadIdx.x; e < nr * P; e += blockDim.x) {
      const int r = e / P, p = e - r * P;
      vx_cp_async4(tile + r * (P + 1) + p, in + (r0 + r) * P + p, true);
    }
    vx_cp_async_commit();
    vx_cp_async_wait_all();
    __syncthreads();
    for (int e = threadIdx.x; e < nr * n; e += blockDim.x) {
      const int r = e / n, j = e - r * n;
      const int lh = lohi[j], lo = lh & 0xffff, hi = lh >> 16;
      const float* row = tile + r * (P + 1);
      const float* w = wt + j * P;
      float acc = 0.f;
      for (int p = lo; p <= hi; ++p) acc = fmaf(w[p], row[p], acc);
      out[(r0 + r) * n + j] = acc;
    }
  }
}

// out[o, j, i] = sum_p weight(j <- p) * in[o, p, i]      in: (outer, P, inner)   out: (outer, n, inner)
__global__ void __launch_bounds__(256) resize_adjoint1d_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                               long long outer, int P, int n, int inner) {
  VX_PDL_ENTRY();
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

// The same reduction with a warp per output element (lanes stride the P axis): the later passes have only a few thousand
// outputs (216 for the last pass of a 3^3 source), each a 96-term sum -- one thread per output is a long serial chain.
__global__ void __launch_bounds__(256) resize_adjoint1d_warp_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                                    long long outer, int P, int n, int inner) {
  VX_PDL_ENTRY();
  const long long total = outer * n * inner;
  const int lane = threadIdx.x & 31;
  for (long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < total; e += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int i = (int)(e % inner), j = (int)((e / inner) % n);
    const long long o = e / ((long long)inner * n);
    int lo = 0, hi = P - 1;
    if (n > 1 && P > 1) {
      const float inv = (float)(P - 1) / (float)(n - 1);
      lo = (int)floorf((float)(j - 1) * inv) - 1;
      hi = (int)ceilf((float)(j + 1) * inv) + 1;
      if (lo < 0) lo = 0;
      if (hi > P - 1) hi = P - 1;
    }
    const float* src = in + (o * P) * inner + i;
    float acc = 0.f;
    for (int p = lo + lane; p <= hi; p += 32) {
      int i0, i1; float w1;
      lerp_ac(p, n, P, i0, i1, w1);
      float wgt = 0.f;
      if (i0 == j) wgt += 1.f - w1;
      if (i1 == j) wgt += w1;
      if (wgt != 0.f) acc = fmaf(wgt, __ldg(src + (size_t)p * inner), acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[e] = acc;
  }
}

static int launch_adjoint(const float* in, float* out, long long outer, int P, int n, int inner, cudaStream_t st) {
  if (inner == 1 && P >= 32 && P <= 512 && n <= 64 && outer >= 4096) {
    const size_t smem = sizeof(float) * ((size_t)n * P + n + (size_t)RA_ROWS * (P + 1));
    if (smem <= 160 * 1024) {
      VX_SET_SMEM(resize_adjoint_rows_kernel, smem);
      long long nblk = (outer + RA_ROWS - 1) / RA_ROWS;
      const int grid = (int)(nblk < kSMs * 4 ? nblk : kSMs * 4);
      VX_LAUNCH(resize_adjoint_rows_kernel, dim3(grid), dim3(256), smem, st, in, out, outer, P, n);
      return check_launch("resize_adjoint_rows_kernel");
    }
  }
  const long long total = outer * n * inner;
  if (total <= (long long)kSMs * 64 * 8 && P >= 32) {      // few outputs, long sums: a warp each
    int blocks = cdiv(total * 32, 256);
    if (blocks < 1) blocks = 1;
    VX_LAUNCH(resize_adjoint1d_warp_kernel, dim3(blocks), dim3(256), 0, st, in, out, outer, P, n, inner);
    return check_launch("resize_adjoint1d_warp_kernel");
  }
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
  prof_bytes(4.0 * ((double)total + (double)d->planes * d->d * d->h * d->w));
  if ((d->W & 3) == 0 && d->W <= 4096 && d->w < 32768 && (((uintptr_t)out[0]) & 15) == 0 && d->D <= 65535 && d->planes <= 65535) {
    const size_t smem = sizeof(float) * 2 * (size_t)d->W;
    VX_LAUNCH(resize_fwd_rows_kernel, dim3(cdiv(d->H, RS_ROWS), d->D, d->planes), dim3(256), smem, (cudaStream_t)stream, A);
    return check_launch("resize_fwd_rows_kernel");
  }
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
