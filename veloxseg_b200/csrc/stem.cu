// Network-input stem and optimiser step: the two ends of the training step.
//
// patch_embed   monai PatchEmbed as used by Encoder.py:150-156: Conv3d k = s = patch (4), C_in -> C_out (16), applied to one
//               modality's channels of the network input.  The library convolution needs the modality slice as its own
//               contiguous tensor and (tf32 path) converts the full-resolution input NCHW -> NHWC in forward and again in the
//               weight-gradient pass: ~130 us per modality at the very end of backward.  Here the input is read where it
//               lies (channel offset into (B, C_total, D, H, W)), a patch row is one 16-byte load, and because the input
//               is the network input there is no data gradient.
// adamw         torch.optim.AdamW (utils/runtime.py optimizer: lr, weight_decay, betas (0.9, 0.999), eps 1e-8) over every
//               parameter tensor in ONE launch through a device pointer table (the multi-tensor library kernels are 10
//               launches of 22-39 CTAs, ~290 us of the step's tail).
#include "vx_kernels.h"

#ifdef VX_EMU
#define __grid_constant__
#endif

namespace vx {

// ---------------------------------------------------------------------------------------------------
// patch embed forward: thread = output voxel, C_out (<= 32) accumulators; weights [tap][co] in shared memory
// ---------------------------------------------------------------------------------------------------
constexpr int PE_MAX_CO = 32;

struct PeArgs {
  const float* x; const float* w; const float* bias; float* y;        // fwd
  const float* dy; float* dw; float* db;                              // bwd
  int B, Ct, c_off, Ci, Co, p, D, H, W, d, h, w_;                     // input extent (D,H,W), output extent (d,h,w_)
};

template <int CO>
__global__ void __launch_bounds__(256, 2) patch_embed_fwd_kernel(const __grid_constant__ PeArgs A) {
  VX_PDL_ENTRY();
  VX_DYN_SMEM(float, ws);                         // [Ci * p^3][CO] (channels past Co are zero)
  const int p = A.p, p3 = p * p * p, K = A.Ci * p3, Co = A.Co;
  for (int i = threadIdx.x; i < K * CO; i += blockDim.x) {
    const int k = i / CO, co = i - k * CO;
    ws[i] = co < Co ? __ldg(A.w + (size_t)co * K + k) : 0.f;      // w is (Co, Ci, p, p, p) = (Co, K)
  }
  __syncthreads();
  const int s = A.d * A.h * A.w_;
  const long long total = (long long)A.B * s;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int o = (int)(e % s), b = (int)(e / s);
  const int ox = o % A.w_, oy = (o / A.w_) % A.h, oz = o / (A.w_ * A.h);
  float acc[CO];
#pragma unroll
  for (int c = 0; c < CO; ++c) acc[c] = 0.f;
  const size_t S = (size_t)A.D * A.H * A.W;
  const bool vec = p == 4 && (A.W & 3) == 0 && (((uintptr_t)A.x) & 15) == 0;
  for (int ci = 0; ci < A.Ci; ++ci) {
    const float* xb = A.x + ((size_t)b * A.Ct + A.c_off + ci) * S + ((size_t)(oz * p) * A.H + oy * p) * A.W + ox * p;
    if (vec) {
      // p = 4: the 16 rows of the patch as 16 independent 16-byte loads issued together (one load per rolled (tz, ty) iteration
      // was a chain of 16 memory round trips per thread: 23 us for a kernel that moves 21 MB)
      float4 q[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) q[r] = __ldg(reinterpret_cast<const float4*>(xb + ((size_t)(r >> 2) * A.H + (r & 3)) * A.W));
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const float4* wr = reinterpret_cast<const float4*>(ws + (size_t)(ci * 16 + r) * 4 * CO);
        const float xv[4] = {q[r].x, q[r].y, q[r].z, q[r].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
          for (int c4 = 0; c4 < CO / 4; ++c4) {
            const float4 w4 = wr[i * (CO / 4) + c4];
            acc[4 * c4] = fmaf(w4.x, xv[i], acc[4 * c4]); acc[4 * c4 + 1] = fmaf(w4.y, xv[i], acc[4 * c4 + 1]);
            acc[4 * c4 + 2] = fmaf(w4.z, xv[i], acc[4 * c4 + 2]); acc[4 * c4 + 3] = fmaf(w4.w, xv[i], acc[4 * c4 + 3]);
          }
        }
      }
      continue;
    }
    for (int tz = 0; tz < p; ++tz)
      for (int ty = 0; ty < p; ++ty) {
        const float* xr = xb + ((size_t)tz * A.H + ty) * A.W;
        const float4* wr = reinterpret_cast<const float4*>(ws + (size_t)((ci * p + tz) * p + ty) * p * CO);
        float xv[4];
        int ntx = p;
        if (vec) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(xr));
          xv[0] = q.x; xv[1] = q.y; xv[2] = q.z; xv[3] = q.w;
          ntx = 4;
        }
        for (int tx0 = 0; tx0 < ntx; tx0 += 4) {
          if (!vec) {
#pragma unroll
            for (int i = 0; i < 4; ++i) xv[i] = tx0 + i < p ? __ldg(xr + tx0 + i) : 0.f;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (tx0 + i >= p) break;
#pragma unroll
            for (int c4 = 0; c4 < CO / 4; ++c4) {
              const float4 w4 = wr[(tx0 + i) * (CO / 4) + c4];
              acc[4 * c4] = fmaf(w4.x, xv[i], acc[4 * c4]); acc[4 * c4 + 1] = fmaf(w4.y, xv[i], acc[4 * c4 + 1]);
              acc[4 * c4 + 2] = fmaf(w4.z, xv[i], acc[4 * c4 + 2]); acc[4 * c4 + 3] = fmaf(w4.w, xv[i], acc[4 * c4 + 3]);
            }
          }
        }
      }
  }
#pragma unroll
  for (int c = 0; c < CO; ++c)
    if (c < Co) A.y[((size_t)b * Co + c) * s + o] = acc[c] + (A.bias ? __ldg(A.bias + c) : 0.f);
}

// ---------------------------------------------------------------------------------------------------
// patch embed weight gradient: dW[co][ci][t] += sum_{b,o} dy[b,co,o] x[b,ci,patch(o)+t];  db[co] += sum dy
// CTA = chunk of PE_VC output voxels of one batch item; dy chunk and the patches of one input channel in shared memory
// ([voxel][tap]: the lanes of a warp are consecutive taps); thread = (tap, 4 output channels), 256 threads cover 64 taps x
// 16 channels per pass; partial sums folded with one atomic per weight and CTA.
// ---------------------------------------------------------------------------------------------------
constexpr int PE_VC = 128;

__global__ void __launch_bounds__(256) patch_embed_wgrad_kernel(const __grid_constant__ PeArgs A) {
  VX_PDL_ENTRY();
  VX_DYN_SMEM(float, sm);
  const int p = A.p, p3 = p * p * p, Co = A.Co, Co4 = (Co + 3) & ~3;
  float* dys = sm;                               // [PE_VC][Co4]
  float* xs = sm + (size_t)PE_VC * Co4;          // [PE_VC][p3]
  const int s = A.d * A.h * A.w_;
  const int nchunk = (s + PE_VC - 1) / PE_VC;
  const int b = blockIdx.x / nchunk, o0 = (blockIdx.x % nchunk) * PE_VC;
  const int nv = min(PE_VC, s - o0);
  const int tid = threadIdx.x, nthr = blockDim.x;
  const size_t S = (size_t)A.D * A.H * A.W;
  for (int i = tid; i < PE_VC * Co4; i += nthr) {
    const int c = i / PE_VC, v = i - c * PE_VC;   // voxel fastest: coalesced reads of dy
    dys[v * Co4 + c] = (c < Co && v < nv) ? __ldg(A.dy + ((size_t)b * Co + c) * s + o0 + v) : 0.f;
  }
  __syncthreads();
  if (tid < Co) {                                 // bias gradient of the chunk
    float t = 0.f;
    for (int v = 0; v < nv; ++v) t += dys[v * Co4 + tid];
    atomicAdd(A.db + tid, t);
  }
  const int nq = Co4 >> 2;                        // channel quads
  for (int ci = 0; ci < A.Ci; ++ci) {
    __syncthreads();
    const float* xc = A.x + ((size_t)b * A.Ct + A.c_off + ci) * S;
    for (int i = tid; i < PE_VC * p * p; i += nthr) {      // item = (voxel, tz, ty): one patch row of p contiguous floats
      const int v = i % PE_VC, r = i / PE_VC, ty = r % p, tz = r / p;
      float* dst = xs + (size_t)v * p3 + (tz * p + ty) * p;
      if (v < nv) {
        const int o = o0 + v, ox = o % A.w_, oy = (o / A.w_) % A.h, oz = o / (A.w_ * A.h);
        const float* xr = xc + ((size_t)(oz * p + tz) * A.H + oy * p + ty) * A.W + ox * p;
        for (int tx = 0; tx < p; ++tx) dst[tx] = __ldg(xr + tx);
      } else {
        for (int tx = 0; tx < p; ++tx) dst[tx] = 0.f;
      }
    }
    __syncthreads();
    for (int item = tid; item < p3 * nq; item += nthr) {   // (tap, channel quad)
      const int t = item % p3, q = item / p3;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int v = 0; v < nv; ++v) {
        const float xv = xs[(size_t)v * p3 + t];
        const float4 g = *reinterpret_cast<const float4*>(dys + v * Co4 + q * 4);
        a0 = fmaf(g.x, xv, a0); a1 = fmaf(g.y, xv, a1); a2 = fmaf(g.z, xv, a2); a3 = fmaf(g.w, xv, a3);
      }
      const float av[4] = {a0, a1, a2, a3};
      for (int j = 0; j < 4; ++j) {
        const int co = q * 4 + j;
        if (co < Co) atomicAdd(A.dw + ((size_t)co * A.Ci + ci) * p3 + t, av[j]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// bias + 3-D pixel shuffle: y[b, c, d s + s1, h s + s2, w s + s3] = z[b, ((c s + s1) s + s2) s + s3, d, h, w] + bias[channel]
// (superpixel.py:15 behind out_conv1 / the reconstruction out_conv, Decoder.py:73-76,150-153).  The library convolution runs
// without its bias; this one pass replaces its 28 MB bias-add pass and the 28 MB permuted copy.  Thread = (b, c, s1, s2, d,
// h, w): s channel planes are read with the lanes along w, the s outputs are contiguous (one 16-byte store for s = 4).
// Backward: the inverse permutation plus the bias gradient (block-reduced, one atomic per channel and CTA).
// ---------------------------------------------------------------------------------------------------
struct PsArgs { const float* src; const float* bias; float* dst; float* db; int B, C, s, d, h, w; };
constexpr int PS_MAX_S = 4;

__global__ void __launch_bounds__(256) pixel_shuffle_fwd_kernel(const __grid_constant__ PsArgs A) {
  VX_PDL_ENTRY();
  const int s = A.s, s2n = s * s, dhw = A.d * A.h * A.w;
  const int cs = blockIdx.y;                         // (c, s1, s2)
  const int c = cs / s2n, s1 = (cs / s) % s, s2 = cs % s;
  const unsigned total = (unsigned)A.B * (unsigned)dhw;      // 32-bit index arithmetic (64-bit divisions are ~100 instructions each)
  const int H = A.h * s, W = A.w * s, D = A.d * s;
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int b = (int)(e / (unsigned)dhw), v = (int)(e - (unsigned)b * (unsigned)dhw);
    const int x = v % A.w, y = (v / A.w) % A.h, z = v / (A.w * A.h);
    const int ch0 = cs * s;
    float o[PS_MAX_S];
#pragma unroll
    for (int s3 = 0; s3 < PS_MAX_S; ++s3)
      if (s3 < s) o[s3] = __ldg(A.src + ((size_t)b * A.C * s2n * s + ch0 + s3) * dhw + v) + (A.bias ? __ldg(A.bias + ch0 + s3) : 0.f);
    float* q = A.dst + ((((size_t)b * A.C + c) * D + z * s + s1) * H + y * s + s2) * W + (size_t)x * s;
    if (s == 4 && (((uintptr_t)q) & 15) == 0) {
      *reinterpret_cast<float4*>(q) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int s3 = 0; s3 < PS_MAX_S; ++s3) if (s3 < s) q[s3] = o[s3];
    }
  }
}

__global__ void __launch_bounds__(256) pixel_shuffle_bwd_kernel(const __grid_constant__ PsArgs A) {
  VX_PDL_ENTRY();
  __shared__ float red[33];
  const int s = A.s, s2n = s * s, dhw = A.d * A.h * A.w;
  const int cs = blockIdx.y;
  const int c = cs / s2n, s1 = (cs / s) % s, s2 = cs % s;
  const unsigned total = (unsigned)A.B * (unsigned)dhw;      // 32-bit index arithmetic (64-bit divisions are ~100 instructions each)
  const int H = A.h * s, W = A.w * s, D = A.d * s;
  float acc[PS_MAX_S];
#pragma unroll
  for (int s3 = 0; s3 < PS_MAX_S; ++s3) acc[s3] = 0.f;
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int b = (int)(e / (unsigned)dhw), v = (int)(e - (unsigned)b * (unsigned)dhw);
    const int x = v % A.w, y = (v / A.w) % A.h, z = v / (A.w * A.h);
    const float* q = A.src + ((((size_t)b * A.C + c) * D + z * s + s1) * H + y * s + s2) * W + (size_t)x * s;
    float g[PS_MAX_S];
    if (s == 4 && (((uintptr_t)q) & 15) == 0) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(q));
      g[0] = t.x; g[1] = t.y; g[2] = t.z; g[3] = t.w;
    } else {
#pragma unroll
      for (int s3 = 0; s3 < PS_MAX_S; ++s3) g[s3] = s3 < s ? __ldg(q + s3) : 0.f;
    }
#pragma unroll
    for (int s3 = 0; s3 < PS_MAX_S; ++s3)
      if (s3 < s) { A.dst[((size_t)b * A.C * s2n * s + cs * s + s3) * dhw + v] = g[s3]; acc[s3] += g[s3]; }
  }
  if (A.db) {
#pragma unroll
    for (int s3 = 0; s3 < PS_MAX_S; ++s3) {
      const float t = block_sum(acc[s3], red);
      if (threadIdx.x == 0 && s3 < s) atomicAdd(A.db + cs * s + s3, t);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// AdamW
// ---------------------------------------------------------------------------------------------------
struct AdamArgs {
  const long long* table;      // [n][4]: param ptr, grad ptr (0 = skip), offset into the flat state, element count
  const int* chunks;           // [nchunk][2]: tensor index, first element of the chunk
  float* m; float* v;          // flat first / second moment
  float* step;                 // (1) step counter, advanced by adamw_tick_kernel
  float lr, beta1, beta2, eps, wd, gscale;
  const float* hyper;          // device (lr, wd) or null
};
constexpr int AD_CHUNK = 1024;

__global__ void adamw_tick_kernel(float* step) {
  VX_PDL_ENTRY(); step[0] += 1.f; }

__global__ void __launch_bounds__(256) adamw_kernel(const __grid_constant__ AdamArgs A) {
  VX_PDL_ENTRY();
  const int t = A.chunks[2 * blockIdx.x], start = A.chunks[2 * blockIdx.x + 1];
  const long long* e = A.table + 4 * (size_t)t;
  float* p = reinterpret_cast<float*>(e[0]);
  const float* g = reinterpret_cast<const float*>(e[1]);
  const long long off = e[2];
  const int n = (int)e[3];
  if (!g || start >= n) return;
  const float step = A.step[0];
  const float lr = A.hyper ? A.hyper[0] : A.lr, wd = A.hyper ? A.hyper[1] : A.wd;
  const float bc1 = 1.f - powf(A.beta1, step), bc2 = 1.f - powf(A.beta2, step);
  const float step_size = lr / bc1, rbc2s = 1.f / sqrtf(bc2), decay = 1.f - lr * wd;
  const int end = min(n, start + AD_CHUNK);
  for (int i = start + threadIdx.x; i < end; i += blockDim.x) {
    const float gi = g[i] * A.gscale;
    float pi = p[i] * decay;
    const float mi = A.m[off + i] + (gi - A.m[off + i]) * (1.f - A.beta1);        // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = A.v[off + i] * A.beta2 + (1.f - A.beta2) * gi * gi;
    A.m[off + i] = mi; A.v[off + i] = vi;
    const float denom = sqrtf(vi) * rbc2s + A.eps;
    pi -= step_size * (mi / denom);
    p[i] = pi;
  }
}

}  // namespace vx

using namespace vx;

static int pe_args(const vx_patch_embed_desc* d, PeArgs& A) {
  if (!d || d->B <= 0 || d->C_in <= 0 || d->C_out <= 0 || d->patch <= 0 || d->c_in_off < 0 ||
      d->c_in_off + d->C_in > d->C_in_total || d->D % d->patch || d->H % d->patch || d->W % d->patch) {
    set_error("patch_embed: bad descriptor"); return VX_ERR_BAD_DESC;
  }
  if (d->C_out > PE_MAX_CO) { set_error("patch_embed: %d output channels > %d", d->C_out, PE_MAX_CO); return VX_ERR_UNSUPPORTED; }
  A.B = d->B; A.Ct = d->C_in_total; A.c_off = d->c_in_off; A.Ci = d->C_in; A.Co = d->C_out; A.p = d->patch;
  A.D = d->D; A.H = d->H; A.W = d->W; A.d = d->D / d->patch; A.h = d->H / d->patch; A.w_ = d->W / d->patch;
  return VX_OK;
}

extern "C" int vx_patch_embed_fwd(const vx_patch_embed_desc* d, const void* const* in, void* const* out, vx_stream_t stream) {
  PeArgs A{};
  int rc = pe_args(d, A);
  if (rc != VX_OK) return rc;
  prof_scope("patch_embed_fwd B%d Ci%d Co%d %dx%dx%d", d->B, d->C_in, d->C_out, d->D, d->H, d->W);
  A.x = (const float*)in[0]; A.w = (const float*)in[1]; A.bias = (const float*)in[2]; A.y = (float*)out[0];
  const int CO = A.Co <= 16 ? 16 : 32;
  const size_t smem = sizeof(float) * (size_t)A.Ci * A.p * A.p * A.p * CO;
  if (smem > 160 * 1024) { set_error("patch_embed: weights do not fit shared memory"); return VX_ERR_UNSUPPORTED; }
  const long long total = (long long)A.B * A.d * A.h * A.w_;
  prof_bytes(4.0 * ((double)A.B * A.Ci * A.D * A.H * A.W + (double)total * A.Co));
  if (CO == 16) {
    VX_SET_SMEM(patch_embed_fwd_kernel<16>, smem);
    VX_LAUNCH(patch_embed_fwd_kernel<16>, dim3(cdiv(total, 256)), dim3(256), smem, (cudaStream_t)stream, A);
  } else {
    VX_SET_SMEM(patch_embed_fwd_kernel<32>, smem);
    VX_LAUNCH(patch_embed_fwd_kernel<32>, dim3(cdiv(total, 256)), dim3(256), smem, (cudaStream_t)stream, A);
  }
  return check_launch("patch_embed_fwd_kernel");
}

extern "C" int vx_patch_embed_bwd(const vx_patch_embed_desc* d, const void* const* in, void* const* out, vx_stream_t stream) {
  PeArgs A{};
  int rc = pe_args(d, A);
  if (rc != VX_OK) return rc;
  prof_scope("patch_embed_bwd B%d Ci%d Co%d %dx%dx%d", d->B, d->C_in, d->C_out, d->D, d->H, d->W);
  A.dy = (const float*)in[0]; A.x = (const float*)in[1]; A.dw = (float*)out[0]; A.db = (float*)out[1];
  cudaStream_t st = (cudaStream_t)stream;
  const int p3 = A.p * A.p * A.p, Co4 = (A.Co + 3) & ~3;
  {
    ZeroList zl;
    zl.add(A.dw, (size_t)A.Co * A.Ci * p3); zl.add(A.db, A.Co);
    rc = zero_many(zl, st);
    if (rc != VX_OK) return rc;
  }
  const size_t smem = sizeof(float) * (size_t)PE_VC * (Co4 + p3);
  if (smem > 160 * 1024) { set_error("patch_embed_bwd: patch too large"); return VX_ERR_UNSUPPORTED; }
  const int s = A.d * A.h * A.w_;
  prof_bytes(4.0 * ((double)A.B * A.Ci * A.D * A.H * A.W + (double)A.B * s * A.Co));
  VX_SET_SMEM(patch_embed_wgrad_kernel, smem);
  VX_LAUNCH(patch_embed_wgrad_kernel, dim3(A.B * cdiv(s, PE_VC)), dim3(256), smem, st, A);
  return check_launch("patch_embed_wgrad_kernel");
}

static int ps_args(const vx_pixel_shuffle_desc* d, PsArgs& A) {
  if (!d || d->B <= 0 || d->C <= 0 || d->scale <= 0 || d->scale > PS_MAX_S || d->d <= 0 || d->h <= 0 || d->w <= 0) {
    set_error("pixel_shuffle: bad descriptor (scale <= %d)", PS_MAX_S); return VX_ERR_BAD_DESC;
  }
  A.B = d->B; A.C = d->C; A.s = d->scale; A.d = d->d; A.h = d->h; A.w = d->w;
  if ((long long)d->B * d->d * d->h * d->w >= (1LL << 31)) { set_error("pixel_shuffle: more than 2^31 voxels"); return VX_ERR_UNSUPPORTED; }
  return VX_OK;
}

static dim3 ps_grid(const PsArgs& A) {
  const long long total = (long long)A.B * A.d * A.h * A.w;
  int gx = cdiv(total, 256 * 2);
  if (gx < 1) gx = 1;
  if (gx > 64) gx = 64;
  return dim3(gx, A.C * A.s * A.s);
}

extern "C" int vx_pixel_shuffle_fwd(const vx_pixel_shuffle_desc* d, const void* const* in, void* const* out, vx_stream_t stream) {
  PsArgs A{};
  int rc = ps_args(d, A);
  if (rc != VX_OK) return rc;
  prof_scope("pixel_shuffle_fwd B%d C%d s%d %dx%dx%d", d->B, d->C, d->scale, d->d, d->h, d->w);
  A.src = (const float*)in[0]; A.bias = (const float*)in[1]; A.dst = (float*)out[0];
  prof_bytes(8.0 * d->B * d->C * (double)d->scale * d->scale * d->scale * d->d * d->h * d->w);
  VX_LAUNCH(pixel_shuffle_fwd_kernel, ps_grid(A), dim3(256), 0, (cudaStream_t)stream, A);
  return check_launch("pixel_shuffle_fwd_kernel");
}

extern "C" int vx_pixel_shuffle_bwd(const vx_pixel_shuffle_desc* d, const void* const* in, void* const* out, vx_stream_t stream) {
  PsArgs A{};
  int rc = ps_args(d, A);
  if (rc != VX_OK) return rc;
  prof_scope("pixel_shuffle_bwd B%d C%d s%d %dx%dx%d", d->B, d->C, d->scale, d->d, d->h, d->w);
  A.src = (const float*)in[0]; A.dst = (float*)out[0]; A.db = (float*)out[1];
  cudaStream_t st = (cudaStream_t)stream;
  if (A.db) {
    ZeroList zl;
    zl.add(A.db, (size_t)d->C * d->scale * d->scale * d->scale);
    rc = zero_many(zl, st);
    if (rc != VX_OK) return rc;
  }
  prof_bytes(8.0 * d->B * d->C * (double)d->scale * d->scale * d->scale * d->d * d->h * d->w);
  VX_LAUNCH(pixel_shuffle_bwd_kernel, ps_grid(A), dim3(256), 0, st, A);
  return check_launch("pixel_shuffle_bwd_kernel");
}

extern "C" int vx_adamw_step(const vx_adamw_desc* d, const void* const* in, void* const* out, vx_stream_t stream) {
  if (!d || d->n_chunks < 0) { set_error("adamw: bad descriptor"); return VX_ERR_BAD_DESC; }
  if (d->n_chunks == 0) return VX_OK;
  prof_scope("adamw chunks%d", d->n_chunks);
  AdamArgs A{};
  A.table = (const long long*)in[0]; A.chunks = (const int*)in[1];
  A.m = (float*)out[0]; A.v = (float*)out[1]; A.step = (float*)out[2];
  A.lr = d->lr; A.beta1 = d->beta1; A.beta2 = d->beta2; A.eps = d->eps; A.wd = d->weight_decay;
  A.gscale = d->grad_scale != 0.f ? d->grad_scale : 1.f;
  A.hyper = d->hyper_on_device ? (const float*)in[2] : nullptr;
  if (d->hyper_on_device && !A.hyper) { set_error("adamw: hyper_on_device without in[2]"); return VX_ERR_BAD_DESC; }
  cudaStream_t st = (cudaStream_t)stream;
  VX_LAUNCH(adamw_tick_kernel, dim3(1), dim3(1), 0, st, A.step);
  int rc = check_launch("adamw_tick_kernel");
  if (rc != VX_OK) return rc;
  VX_LAUNCH(adamw_kernel, dim3(d->n_chunks), dim3(256), 0, st, A);
  return check_launch("adamw_kernel");
}
