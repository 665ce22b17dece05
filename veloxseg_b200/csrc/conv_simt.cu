// Strided and transposed convolutions of the glue layers in fp32 on the CUDA cores:
//   DownConv.down   Conv3d k = 2p-1, stride p, pad p-1   model/components/conv_blocks.py:10-17   (p = 4: k7 s4; p = 2: k3 s2)
//   UpConv.up       ConvTranspose3d k = stride = 2        model/components/conv_blocks.py:31-35
// Three generic kernels cover forward, data gradient and weight gradient of both (a transposed convolution's forward is a
// strided convolution's data gradient and vice versa; their weight gradients have the same form with the roles of the two
// activations exchanged).  Weights are addressed as W(o, i, t) = w[o*so + i*si + t] so that both torch layouts fit:
// Conv3d (C_out, C_in, k^3) and ConvTranspose3d (C_in, C_out, k^3).
//
//   strided    Y[b, o, v]  = sum_{i, t} W(o, i, t) * X[b, i, s*v + t - pad]                                (v coarse, X fine)
//   scatter    Y[b, i, q]  = sum_{o, t : (q + pad - t) % s == 0} W(o, i, t) * X[b, o, (q + pad - t) / s]   (q fine, X coarse)
//   wgrad      dW(o, i, t) = sum_{b, v} G[b, o, v] * F[b, i, s*v + t - pad]                                (G coarse, F fine)
// These layers are 11 % of the forward MACs (down1 alone 9 %); levels 2-4 are a few MFLOP each and latency-bound, so the kernels
// favour many small CTAs over tiling: a thread owns one output voxel (or one weight tap) and a register tile of channels,
// weights / coarse gradients are staged in shared memory in the order the inner loop walks them.
#include "vx_kernels.h"

namespace vx {

struct ConvGeo {
  int B, Co, Ci;                 // channels of the coarse (o) and fine (i) side
  int d, h, w;                   // coarse extent
  int D, H, W;                   // fine extent
  int k, s, pad;
  int so, si;                    // weight strides of o and i (taps are contiguous)
};

// ---------------------------------------------------------------------------------------------------------------------
// strided: thread = one coarse voxel x CT output channels; weights of the CTA's channel tile staged as [i*k3 + t][CT]
// ---------------------------------------------------------------------------------------------------------------------
template <int CT>
__global__ void __launch_bounds__(128) conv_strided_kernel(ConvGeo G, const float* __restrict__ X, const float* __restrict__ Wt,
                                                           const float* __restrict__ bias, float* __restrict__ Y, int kchunk) {
  VX_DYN_SMEM(float, ws);                              // [kchunk][CT]
  const int k = G.k, k2 = k * k, k3 = k2 * k;
  const int sv = G.d * G.h * G.w;
  const size_t SV = (size_t)G.D * G.H * G.W;
  const int o0 = blockIdx.y * CT;
  const long long gv = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = gv < (long long)G.B * sv;
  const int b = live ? (int)(gv / sv) : 0, v = live ? (int)(gv % sv) : 0;
  const int vx_ = v % G.w, vy = (v / G.w) % G.h, vz = v / (G.w * G.h);
  const int z0 = vz * G.s - G.pad, y0 = vy * G.s - G.pad, x0 = vx_ * G.s - G.pad;
  float acc[CT];
#pragma unroll
  for (int c = 0; c < CT; ++c) acc[c] = 0.f;
  const int K = G.Ci * k3;
  for (int kc = 0; kc < K; kc += kchunk) {
    const int kn = min(kchunk, K - kc);
    __syncthreads();
    for (int e = threadIdx.x; e < kn * CT; e += blockDim.x) {
      const int c = e / kn, kk = e % kn;               // kk fastest: contiguous reads of one output channel's taps
      const int i = (kc + kk) / k3, t = (kc + kk) % k3;
      ws[kk * CT + c] = (o0 + c < G.Co) ? __ldg(Wt + (size_t)(o0 + c) * G.so + (size_t)i * G.si + t) : 0.f;
    }
    __syncthreads();
    if (!live) continue;
    int i = kc / k3, t = kc % k3;
    int tz = t / k2, ty = (t / k) % k, tx = t % k;
    const float* xb = X + ((size_t)b * G.Ci + i) * SV;
    for (int kk = 0; kk < kn; ++kk) {
      const int z = z0 + tz, y = y0 + ty, x = x0 + tx;
      float xv = 0.f;
      if (z >= 0 && z < G.D && y >= 0 && y < G.H && x >= 0 && x < G.W) xv = __ldg(xb + ((size_t)z * G.H + y) * G.W + x);
      const float* wr = ws + kk * CT;
#pragma unroll
      for (int c = 0; c < CT; c += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(wr + c);
        acc[c] = fmaf(w4.x, xv, acc[c]); acc[c + 1] = fmaf(w4.y, xv, acc[c + 1]);
        acc[c + 2] = fmaf(w4.z, xv, acc[c + 2]); acc[c + 3] = fmaf(w4.w, xv, acc[c + 3]);
      }
      if (++tx == k) { tx = 0; if (++ty == k) { ty = 0; if (++tz == k) { tz = 0; xb += SV; } } }
    }
  }
  if (!live) return;
#pragma unroll
  for (int c = 0; c < CT; ++c)
    if (o0 + c < G.Co) Y[((size_t)b * G.Co + o0 + c) * sv + v] = acc[c] + (bias ? __ldg(bias + o0 + c) : 0.f);
}

// ---------------------------------------------------------------------------------------------------------------------
// scatter form (transposed convolution / data gradient of a strided one): thread = one fine voxel x CT fine channels;
// weights of the channel tile staged as [o][t][CT] in chunks of coarse channels
// ---------------------------------------------------------------------------------------------------------------------
template <int CT>
__global__ void __launch_bounds__(128) conv_scatter_kernel(ConvGeo G, const float* __restrict__ X, const float* __restrict__ Wt,
                                                           const float* __restrict__ bias, float* __restrict__ Y, int ochunk) {
  VX_DYN_SMEM(float, ws);                              // [ochunk][k3][CT]
  const int k = G.k, k2 = k * k, k3 = k2 * k, s = G.s;
  const int sv = G.d * G.h * G.w;
  const size_t SV = (size_t)G.D * G.H * G.W;
  const int i0 = blockIdx.y * CT;
  const long long gq = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = gq < (long long)G.B * (long long)SV;
  const int b = live ? (int)(gq / (long long)SV) : 0;
  const int q = live ? (int)(gq % (long long)SV) : 0;
  const int qx = q % G.W, qy = (q / G.W) % G.H, qz = q / (G.W * G.H);
  // per axis: taps t = t0, t0 + s, ... < k with coarse index (q + pad - t) / s inside the coarse extent
  int tz0 = (qz + G.pad) % s, ty0 = (qy + G.pad) % s, tx0 = (qx + G.pad) % s;
  float acc[CT];
#pragma unroll
  for (int c = 0; c < CT; ++c) acc[c] = 0.f;
  for (int oc = 0; oc < G.Co; oc += ochunk) {
    const int on = min(ochunk, G.Co - oc);
    __syncthreads();
    for (int e = threadIdx.x; e < on * k3 * CT; e += blockDim.x) {
      const int t = e % k3, c = (e / k3) % CT, o = e / (k3 * CT);      // taps fastest: contiguous reads
      ws[(o * k3 + t) * CT + c] = (i0 + c < G.Ci) ? __ldg(Wt + (size_t)(oc + o) * G.so + (size_t)(i0 + c) * G.si + t) : 0.f;
    }
    __syncthreads();
    if (!live) continue;
    for (int o = 0; o < on; ++o) {
      const float* xo = X + ((size_t)b * G.Co + oc + o) * sv;
      for (int tz = tz0; tz < k; tz += s) {
        const int vz = (qz + G.pad - tz) / s;
        if (qz + G.pad - tz < 0 || vz >= G.d) continue;
        for (int ty = ty0; ty < k; ty += s) {
          const int vy = (qy + G.pad - ty) / s;
          if (qy + G.pad - ty < 0 || vy >= G.h) continue;
          for (int tx = tx0; tx < k; tx += s) {
            const int vx_ = (qx + G.pad - tx) / s;
            if (qx + G.pad - tx < 0 || vx_ >= G.w) continue;
            const float xv = __ldg(xo + ((size_t)vz * G.h + vy) * G.w + vx_);
            const float* wr = ws + (o * k3 + (tz * k + ty) * k + tx) * CT;
#pragma unroll
            for (int c = 0; c < CT; c += 4) {
              const float4 w4 = *reinterpret_cast<const float4*>(wr + c);
              acc[c] = fmaf(w4.x, xv, acc[c]); acc[c + 1] = fmaf(w4.y, xv, acc[c + 1]);
              acc[c + 2] = fmaf(w4.z, xv, acc[c + 2]); acc[c + 3] = fmaf(w4.w, xv, acc[c + 3]);
            }
          }
        }
      }
    }
  }
  if (!live) return;
#pragma unroll
  for (int c = 0; c < CT; ++c)
    if (i0 + c < G.Ci) Y[((size_t)b * G.Ci + i0 + c) * SV + q] = acc[c] + (bias ? __ldg(bias + i0 + c) : 0.f);
}

// ---------------------------------------------------------------------------------------------------------------------
// weight gradient: CTA = (block of (i, t) pairs, tile of 16 coarse channels, chunk of coarse voxels); the chunk's coarse
// gradients are staged as [v][16]; thread = one (i, t) pair with 16 accumulators, atomically added into dW (zeroed by the
// launcher).  db[o] (optional) = sum of G over (b, v), from the CTAs of the first pair block.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int CW_CT = 16;
__global__ void __launch_bounds__(128) conv_wgrad_kernel(ConvGeo G, const float* __restrict__ Gc, const float* __restrict__ F,
                                                         float* __restrict__ dW, float* __restrict__ db, int vchunk) {
  VX_DYN_SMEM(float, gs);                              // [vchunk][16]
  const int k = G.k, k2 = k * k, k3 = k2 * k;
  const int sv = G.d * G.h * G.w;
  const size_t SV = (size_t)G.D * G.H * G.W;
  const int npair = G.Ci * k3;
  const int pr = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = pr < npair;
  const int i = live ? pr / k3 : 0, t = live ? pr % k3 : 0;
  const int tz = t / k2 - G.pad, ty = (t / k) % k - G.pad, tx = t % k - G.pad;
  const int o0 = blockIdx.y * CW_CT;
  const long long total = (long long)G.B * sv;
  const long long c0 = (long long)blockIdx.z * vchunk;
  const int nv = (int)min((long long)vchunk, total - c0);
  for (int e = threadIdx.x; e < vchunk * CW_CT; e += blockDim.x) {
    const int c = e / vchunk, vv = e % vchunk;         // voxels fastest: coalesced reads
    float g = 0.f;
    if (vv < nv && o0 + c < G.Co) {
      const long long gv = c0 + vv;
      g = __ldg(Gc + ((size_t)(gv / sv) * G.Co + o0 + c) * sv + (size_t)(gv % sv));
    }
    gs[vv * CW_CT + c] = g;
  }
  __syncthreads();
  if (db && blockIdx.x == 0 && threadIdx.x < CW_CT && o0 + threadIdx.x < G.Co) {
    float sacc = 0.f;
    for (int vv = 0; vv < nv; ++vv) sacc += gs[vv * CW_CT + threadIdx.x];
    atomicAdd(db + o0 + threadIdx.x, sacc);
  }
  if (!live) return;
  float acc[CW_CT];
#pragma unroll
  for (int c = 0; c < CW_CT; ++c) acc[c] = 0.f;
  long long gv = c0;
  int b = (int)(gv / sv), v = (int)(gv % sv);
  int vx_ = v % G.w, vy = (v / G.w) % G.h, vz = v / (G.w * G.h);
  for (int vv = 0; vv < nv; ++vv) {
    const int z = vz * G.s + tz, y = vy * G.s + ty, x = vx_ * G.s + tx;
    if (z >= 0 && z < G.D && y >= 0 && y < G.H && x >= 0 && x < G.W) {
      const float fv = __ldg(F + ((size_t)b * G.Ci + i) * SV + ((size_t)z * G.H + y) * G.W + x);
      const float* gr = gs + vv * CW_CT;
#pragma unroll
      for (int c = 0; c < CW_CT; c += 4) {
        const float4 g4 = *reinterpret_cast<const float4*>(gr + c);
        acc[c] = fmaf(g4.x, fv, acc[c]); acc[c + 1] = fmaf(g4.y, fv, acc[c + 1]);
        acc[c + 2] = fmaf(g4.z, fv, acc[c + 2]); acc[c + 3] = fmaf(g4.w, fv, acc[c + 3]);
      }
    }
    if (++vx_ == G.w) { vx_ = 0; if (++vy == G.h) { vy = 0; if (++vz == G.d) { vz = 0; ++b; } } }
  }
#pragma unroll
  for (int c = 0; c < CW_CT; ++c)
    if (o0 + c < G.Co) atomicAdd(dW + (size_t)(o0 + c) * G.so + (size_t)i * G.si + t, acc[c]);
}

// out[c] = sum over (b, s) of X[b, c, s]  (bias gradient of a transposed convolution): one CTA per channel, fixed fold order
__global__ void __launch_bounds__(256) channel_sum_kernel(const float* __restrict__ X, float* __restrict__ out, int B, int C, int S) {
  __shared__ float red[33];
  const int c = blockIdx.x;
  float acc = 0.f;
  for (int b = 0; b < B; ++b) {
    const float* p = X + ((size_t)b * C + c) * S;
    for (int i = threadIdx.x; i < S; i += blockDim.x) acc += __ldg(p + i);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) out[c] = acc;
}

// ---------------------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------------------
static int pick_ct(long long threads_at_16) { return threads_at_16 >= 16384 ? 16 : threads_at_16 >= 2048 ? 8 : 4; }

int conv_strided(const ConvGeo& G, const float* X, const float* Wt, const float* bias, float* Y, cudaStream_t st) {
  const long long nvox = (long long)G.B * G.d * G.h * G.w;
  int ct = pick_ct(nvox);
  while (ct > 4 && ct > G.Co) ct >>= 1;
  const int K = G.Ci * G.k * G.k * G.k;
  int kchunk = (40 * 1024 / 4) / ct;
  if (kchunk > K) kchunk = K;
  const size_t smem = sizeof(float) * (size_t)kchunk * ct;
  const dim3 grid(cdiv(nvox, 128), cdiv(G.Co, ct));
  prof_bytes(4.0 * ((double)G.B * G.Ci * G.D * G.H * G.W + (double)nvox * G.Co + (double)G.Co * K));
  if (ct == 16) { VX_LAUNCH(conv_strided_kernel<16>, grid, dim3(128), smem, st, G, X, Wt, bias, Y, kchunk); }
  else if (ct == 8) { VX_LAUNCH(conv_strided_kernel<8>, grid, dim3(128), smem, st, G, X, Wt, bias, Y, kchunk); }
  else { VX_LAUNCH(conv_strided_kernel<4>, grid, dim3(128), smem, st, G, X, Wt, bias, Y, kchunk); }
  return check_launch("conv_strided_kernel");
}

int conv_scatter(const ConvGeo& G, const float* X, const float* Wt, const float* bias, float* Y, cudaStream_t st) {
  const long long nfine = (long long)G.B * G.D * G.H * G.W;
  int ct = pick_ct(nfine);
  while (ct > 4 && ct > G.Ci) ct >>= 1;
  const int k3 = G.k * G.k * G.k;
  int ochunk = (40 * 1024 / 4) / (k3 * ct);
  if (ochunk < 1) ochunk = 1;
  if (ochunk > G.Co) ochunk = G.Co;
  const size_t smem = sizeof(float) * (size_t)ochunk * k3 * ct;
  const dim3 grid(cdiv(nfine, 128), cdiv(G.Ci, ct));
  prof_bytes(4.0 * ((double)nfine * G.Ci + (double)G.B * G.Co * G.d * G.h * G.w + (double)G.Co * G.Ci * k3));
  if (ct == 16) { VX_LAUNCH(conv_scatter_kernel<16>, grid, dim3(128), smem, st, G, X, Wt, bias, Y, ochunk); }
  else if (ct == 8) { VX_LAUNCH(conv_scatter_kernel<8>, grid, dim3(128), smem, st, G, X, Wt, bias, Y, ochunk); }
  else { VX_LAUNCH(conv_scatter_kernel<4>, grid, dim3(128), smem, st, G, X, Wt, bias, Y, ochunk); }
  return check_launch("conv_scatter_kernel");
}

// dW (and db) must be zeroed by the caller (zero_many)
int conv_wgrad(const ConvGeo& G, const float* Gc, const float* F, float* dW, float* db, cudaStream_t st) {
  const long long nvox = (long long)G.B * G.d * G.h * G.w;
  const int npair = G.Ci * G.k * G.k * G.k;
  const int gx = cdiv(npair, 128), gy = cdiv(G.Co, CW_CT);
  // enough voxel chunks for ~4 CTAs per SM, chunks of 64 .. 512 voxels
  long long want = (4LL * kSMs + (long long)gx * gy - 1) / ((long long)gx * gy);
  if (want < 1) want = 1;
  long long vchunk = (nvox + want - 1) / want;
  if (vchunk < 64) vchunk = 64;
  if (vchunk > 512) vchunk = 512;
  const int gz = cdiv(nvox, vchunk);
  const size_t smem = sizeof(float) * (size_t)vchunk * CW_CT;
  prof_bytes(4.0 * ((double)G.B * G.Ci * G.D * G.H * G.W + (double)nvox * G.Co + (double)G.Co * npair));
  VX_LAUNCH(conv_wgrad_kernel, dim3(gx, gy, gz), dim3(128), smem, st, G, Gc, F, dW, db, (int)vchunk);
  return check_launch("conv_wgrad_kernel");
}

// =====================================================================================================================
// C ABI: vx_conv_* dispatch (conv3_tc.cu for the dense 3x3x3 convolution, the kernels above otherwise)
// =====================================================================================================================
bool conv3_tc_supported(const vx_conv_desc* d);
size_t conv3_tc_workspace(const vx_conv_desc* d);
int conv3_tc_fwd(const vx_conv_desc* d, const float* x, const float* w, const float* bias, float* y, void* ws, size_t ws_bytes, cudaStream_t st);
int conv3_tc_bwd(const vx_conv_desc* d, const float* dy, const float* x, const float* w, float* dx, float* dw, float* db, void* ws,
                 size_t ws_bytes, cudaStream_t st);

static bool simt_geo(const vx_conv_desc* d, ConvGeo& G) {
  if (!d || d->B <= 0 || d->C_in <= 0 || d->C_out <= 0 || d->D <= 0 || d->H <= 0 || d->W <= 0 || d->kernel <= 0 || d->stride <= 0 ||
      d->pad < 0 || d->shuffle != 0)
    return false;
  const int k3 = d->kernel * d->kernel * d->kernel;
  G.B = d->B; G.k = d->kernel; G.s = d->stride; G.pad = d->pad;
  if (!d->transposed) {
    // coarse = output (C_out), fine = input (C_in); weight (C_out, C_in, k3)
    if (d->D + 2 * d->pad < d->kernel || d->H + 2 * d->pad < d->kernel || d->W + 2 * d->pad < d->kernel) return false;
    G.Co = d->C_out; G.Ci = d->C_in;
    G.D = d->D; G.H = d->H; G.W = d->W;
    G.d = (d->D + 2 * d->pad - d->kernel) / d->stride + 1; G.h = (d->H + 2 * d->pad - d->kernel) / d->stride + 1;
    G.w = (d->W + 2 * d->pad - d->kernel) / d->stride + 1;
    G.so = d->C_in * k3; G.si = k3;
  } else {
    // ConvTranspose3d k == s, pad 0: coarse = input (C_in), fine = output (C_out); weight (C_in, C_out, k3)
    if (d->kernel != d->stride || d->pad != 0) return false;
    G.Co = d->C_in; G.Ci = d->C_out;
    G.d = d->D; G.h = d->H; G.w = d->W;
    G.D = d->D * d->stride; G.H = d->H * d->stride; G.W = d->W * d->stride;
    G.so = d->C_out * k3; G.si = k3;
  }
  return true;
}

}  // namespace vx

using namespace vx;

extern "C" size_t vx_conv_workspace(const vx_conv_desc* d) { return conv3_tc_supported(d) ? conv3_tc_workspace(d) : 0; }

extern "C" int vx_conv_fwd(const vx_conv_desc* d, const void* const* in, void* const* out, void* ws, size_t ws_bytes, vx_stream_t stream) {
  if (!d || !in || !out || !in[0] || !in[1] || !out[0]) { set_error("conv_fwd: null pointer"); return VX_ERR_BAD_DESC; }
  cudaStream_t st = (cudaStream_t)stream;
  const float* x = (const float*)in[0]; const float* w = (const float*)in[1]; const float* bias = (const float*)in[2];
  if (conv3_tc_supported(d)) return conv3_tc_fwd(d, x, w, bias, (float*)out[0], ws, ws_bytes, st);
  ConvGeo G{};
  if (!simt_geo(d, G)) { set_error("conv_fwd: unsupported geometry (k%d s%d p%d transposed %d shuffle %d)", d->kernel, d->stride, d->pad, d->transposed, d->shuffle); return VX_ERR_UNSUPPORTED; }
  prof_scope("conv_fwd %s B%d %d->%d k%d s%d %dx%dx%d", d->transposed ? "T" : "S", d->B, d->C_in, d->C_out, d->kernel, d->stride, d->D, d->H, d->W);
  return d->transposed ? conv_scatter(G, x, w, bias, (float*)out[0], st) : conv_strided(G, x, w, bias, (float*)out[0], st);
}

extern "C" int vx_conv_bwd(const vx_conv_desc* d, const void* const* in, void* const* out, void* ws, size_t ws_bytes, vx_stream_t stream) {
  if (!d || !in || !out || !in[0] || !in[1] || !in[2] || !out[1]) { set_error("conv_bwd: null pointer"); return VX_ERR_BAD_DESC; }
  cudaStream_t st = (cudaStream_t)stream;
  const float* dy = (const float*)in[0]; const float* x = (const float*)in[1]; const float* w = (const float*)in[2];
  float* dx = (float*)out[0]; float* dw = (float*)out[1]; float* db = (float*)out[2];
  if (conv3_tc_supported(d)) return conv3_tc_bwd(d, dy, x, w, dx, dw, db, ws, ws_bytes, st);
  ConvGeo G{};
  if (!simt_geo(d, G)) { set_error("conv_bwd: unsupported geometry"); return VX_ERR_UNSUPPORTED; }
  prof_scope("conv_bwd %s B%d %d->%d k%d s%d %dx%dx%d", d->transposed ? "T" : "S", d->B, d->C_in, d->C_out, d->kernel, d->stride, d->D, d->H, d->W);
  const int k3 = d->kernel * d->kernel * d->kernel;
  int rc;
  {
    SideJoin join(st);                                 // weight gradient on the side stream (a leaf of the backward graph)
    cudaStream_t sw = side_fork(st);
    ZeroList zl;
    zl.add(dw, (size_t)d->C_in * d->C_out * k3);
    if (db && !d->transposed) zl.add(db, (size_t)d->C_out);
    rc = zero_many(zl, sw);
    if (rc != VX_OK) return rc;
    if (!d->transposed) {
      rc = conv_wgrad(G, dy, x, dw, db, sw);           // coarse gradient = dy (C_out), fine = x (C_in)
    } else {
      // dW(cin, cout, t) = sum x[cin][v] * dy[cout][s v + t]: coarse = x, fine = dy; the bias gradient is over dy (fine side)
      rc = conv_wgrad(G, x, dy, dw, nullptr, sw);
      if (rc == VX_OK && db) {
        VX_LAUNCH(channel_sum_kernel, dim3(d->C_out), dim3(256), 0, sw, dy, db, d->B, d->C_out, G.D * G.H * G.W);
        rc = check_launch("channel_sum_kernel");
      }
    }
    if (rc != VX_OK) return rc;
  }
  if (dx) {
    rc = d->transposed ? conv_strided(G, dy, w, nullptr, dx, st)       // dx[cin][v] = sum W(cin, cout, t) dy[cout][s v + t]
                       : conv_scatter(G, dy, w, nullptr, dx, st);      // dx[ci][q] = sum W(co, ci, t) dy[co][(q + pad - t)/s]
    if (rc != VX_OK) return rc;
  }
  return VX_OK;
}
