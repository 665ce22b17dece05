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
  int flip;                      // strided kernel only: read tap k^3-1-t (a stride-1 data gradient as a convolution)
};

// ---------------------------------------------------------------------------------------------------------------------
// strided: CTA = 32 coarse voxels (lanes) x CT output channels x KS slices of the reduction (warps).  The reduction walks
// ROWS of the kernel, r = (i, tz, ty): one bounds test and one address per row, then KW taps along x with KW predicated
// loads in flight.  Weights of the CTA's channel tile are staged per chunk of rows as [row][tx][CT]; the KS partial sums
// meet in shared memory (fixed order).
// ---------------------------------------------------------------------------------------------------------------------
template <int CT, int KW>      // KW = kernel width known at compile time, 0 = runtime
__global__ void __launch_bounds__(256, 2) conv_strided_kernel(ConvGeo G, const float* __restrict__ X, const float* __restrict__ Wt,
                                                           const float* __restrict__ bias, float* __restrict__ Y, int rchunk) {
  VX_PDL_ENTRY();
  VX_DYN_SMEM(float, ws);                              // [rchunk][k][CT], reused as [KS][32][CT] for the final fold
  const int k = KW ? KW : G.k, k2 = k * k, k3 = k2 * k;
  const int lane = threadIdx.x, ks = threadIdx.y, KS = blockDim.y, vw = threadIdx.z, VW = blockDim.z;
  const int tid = (vw * KS + ks) * 32 + lane, nthr = 32 * KS * VW;
  const int sv = G.d * G.h * G.w;
  const size_t SV = (size_t)G.D * G.H * G.W;
  const int o0 = blockIdx.y * CT;
  const long long gv = ((long long)blockIdx.x * VW + vw) * 32 + lane;
  const bool live = gv < (long long)G.B * sv;
  const int b = live ? (int)(gv / sv) : 0, v = live ? (int)(gv % sv) : 0;
  const int vx_ = v % G.w, vy = (v / G.w) % G.h, vz = v / (G.w * G.h);
  const int z0 = vz * G.s - G.pad, y0 = vy * G.s - G.pad, x0 = vx_ * G.s - G.pad;
  const float* xb = X + (size_t)b * G.Ci * SV;
  float acc[CT];
#pragma unroll
  for (int c = 0; c < CT; ++c) acc[c] = 0.f;
  const int R = G.Ci * k2;
  for (int rc = 0; rc < R; rc += rchunk) {
    const int rn = min(rchunk, R - rc);
    __syncthreads();
#pragma unroll 4
    for (int e = tid; e < rn * k * CT; e += nthr) {      // 4 independent loads in flight per thread (17 % of the kernel's samples rolled)
      const int c = e / (rn * k), rt = e % (rn * k);   // (row, tx) fastest: contiguous reads of one output channel's taps
      const int r = rc + rt / k, tx = rt % k;
      const int i = r / k2, tzy = r % k2;
      const int tap = G.flip ? k3 - 1 - (tzy * k + tx) : tzy * k + tx;
      ws[rt * CT + c] = (o0 + c < G.Co) ? __ldg(Wt + (size_t)(o0 + c) * G.so + (size_t)i * G.si + tap) : 0.f;
    }
    __syncthreads();
    if (!live) continue;
#pragma unroll 4
    for (int rr = ks; rr < rn; rr += KS) {
      const int r = rc + rr;
      const int i = r / k2, tz = (r % k2) / k, ty = r % k;
      const int z = z0 + tz, y = y0 + ty;
      const bool rok = z >= 0 && z < G.D && y >= 0 && y < G.H;
      if (!KW && !rok) continue;
      const float* xr = xb + (size_t)i * SV + ((size_t)z * G.H + y) * G.W + x0;
      const float* wr = ws + rr * k * CT;
      if (KW) {
        // branch-free rows (an out-of-bounds row loads nothing and adds zeros): the loads of the unrolled rows can be issued
        // together -- with a `continue` here ptxas kept each row's loads behind the previous row's FMAs
        float xv[KW ? KW : 1];
#pragma unroll
        for (int tx = 0; tx < KW; ++tx) xv[tx] = (rok && x0 + tx >= 0 && x0 + tx < G.W) ? __ldg(xr + tx) : 0.f;
#pragma unroll
        for (int tx = 0; tx < KW; ++tx) {
#pragma unroll
          for (int c = 0; c < CT; c += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(wr + tx * CT + c);
            acc[c] = fmaf(w4.x, xv[tx], acc[c]); acc[c + 1] = fmaf(w4.y, xv[tx], acc[c + 1]);
            acc[c + 2] = fmaf(w4.z, xv[tx], acc[c + 2]); acc[c + 3] = fmaf(w4.w, xv[tx], acc[c + 3]);
          }
        }
      } else {
        for (int tx = 0; tx < k; ++tx) {
          const float xv = (x0 + tx >= 0 && x0 + tx < G.W) ? __ldg(xr + tx) : 0.f;
#pragma unroll
          for (int c = 0; c < CT; c += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(wr + tx * CT + c);
            acc[c] = fmaf(w4.x, xv, acc[c]); acc[c + 1] = fmaf(w4.y, xv, acc[c + 1]);
            acc[c + 2] = fmaf(w4.z, xv, acc[c + 2]); acc[c + 3] = fmaf(w4.w, xv, acc[c + 3]);
          }
        }
      }
    }
  }
  if (KS > 1) {                                        // fold the KS slices in slice order
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CT; ++c) ws[tid * CT + c] = acc[c];
    __syncthreads();
    if (ks != 0) return;
    for (int j = 1; j < KS; ++j)
#pragma unroll
      for (int c = 0; c < CT; ++c) acc[c] += ws[(tid + j * 32) * CT + c];
  }
  if (!live) return;
#pragma unroll
  for (int c = 0; c < CT; ++c)
    if (o0 + c < G.Co) Y[((size_t)b * G.Co + o0 + c) * sv + v] = acc[c] + (bias ? __ldg(bias + o0 + c) : 0.f);
}

// ---------------------------------------------------------------------------------------------------------------------
// scatter form (transposed convolution / data gradient of a strided one): CTA = 32 fine voxels (lanes) x CT fine channels x
// KS slices of the coarse channels (warps).  The taps that reach a fine voxel do not depend on the channel: their coarse
// offsets and tap indices (at most ceil(k/s)^3 <= 8) are listed once per thread; weights of the channel tile are staged per
// chunk of coarse channels as [o][t][CT].
// ---------------------------------------------------------------------------------------------------------------------
template <int CT>
__global__ void __launch_bounds__(256) conv_scatter_kernel(ConvGeo G, const float* __restrict__ X, const float* __restrict__ Wt,
                                                           const float* __restrict__ bias, float* __restrict__ Y, int ochunk) {
  VX_PDL_ENTRY();
  VX_DYN_SMEM(float, ws);                              // [ochunk][k3][CT], reused for the final fold
  const int k = G.k, k2 = k * k, k3 = k2 * k, s = G.s;
  const int lane = threadIdx.x, ks = threadIdx.y, KS = blockDim.y, vw = threadIdx.z, VW = blockDim.z;
  const int tid = (vw * KS + ks) * 32 + lane, nthr = 32 * KS * VW;
  const int sv = G.d * G.h * G.w;
  const size_t SV = (size_t)G.D * G.H * G.W;
  const int i0 = blockIdx.y * CT;
  const long long gq = ((long long)blockIdx.x * VW + vw) * 32 + lane;
  const bool live = gq < (long long)G.B * (long long)SV;
  const int b = live ? (int)(gq / (long long)SV) : 0;
  const int q = live ? (int)(gq % (long long)SV) : 0;
  const int qx = q % G.W, qy = (q / G.W) % G.H, qz = q / (G.W * G.H);
  int toff[8], ttap[8], ntap = 0;
  if (live) {
    for (int tz = (qz + G.pad) % s; tz < k; tz += s) {
      const int vz = (qz + G.pad - tz) / s;
      if (qz + G.pad - tz < 0 || vz >= G.d) continue;
      for (int ty = (qy + G.pad) % s; ty < k; ty += s) {
        const int vy = (qy + G.pad - ty) / s;
        if (qy + G.pad - ty < 0 || vy >= G.h) continue;
        for (int tx = (qx + G.pad) % s; tx < k; tx += s) {
          const int vx_ = (qx + G.pad - tx) / s;
          if (qx + G.pad - tx < 0 || vx_ >= G.w || ntap >= 8) continue;
          toff[ntap] = (vz * G.h + vy) * G.w + vx_; ttap[ntap] = ((tz * k + ty) * k + tx) * CT; ++ntap;
        }
      }
    }
  }
  float acc[CT];
#pragma unroll
  for (int c = 0; c < CT; ++c) acc[c] = 0.f;
  const float* xbase = X + (size_t)b * G.Co * sv;
  for (int oc = 0; oc < G.Co; oc += ochunk) {
    const int on = min(ochunk, G.Co - oc);
    __syncthreads();
    for (int e = tid; e < on * k3 * CT; e += nthr) {
      const int t = e % k3, c = (e / k3) % CT, o = e / (k3 * CT);      // taps fastest: contiguous reads
      ws[(o * k3 + t) * CT + c] = (i0 + c < G.Ci) ? __ldg(Wt + (size_t)(oc + o) * G.so + (size_t)(i0 + c) * G.si + t) : 0.f;
    }
    __syncthreads();
    for (int o = ks; o < on; o += KS) {
      const float* xo = xbase + (size_t)(oc + o) * sv;
      const float* wo = ws + o * k3 * CT;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < ntap) {
          const float xv = __ldg(xo + toff[j]);
          const float* wr = wo + ttap[j];
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
  if (KS > 1) {
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CT; ++c) ws[tid * CT + c] = acc[c];
    __syncthreads();
    if (ks != 0) return;
    for (int j = 1; j < KS; ++j)
#pragma unroll
      for (int c = 0; c < CT; ++c) acc[c] += ws[(tid + j * 32) * CT + c];
  }
  if (!live) return;
#pragma unroll
  for (int c = 0; c < CT; ++c)
    if (i0 + c < G.Ci) Y[((size_t)b * G.Ci + i0 + c) * SV + q] = acc[c] + (bias ? __ldg(bias + i0 + c) : 0.f);
}

// ---------------------------------------------------------------------------------------------------------------------
// weight gradient: CTA = (block of (i, t) pairs, tile of 16 coarse channels, chunk of coarse voxels); the chunk's coarse
// gradients are staged as [v][16] together with each voxel's fine-grid origin; thread = one (i, t) pair with 16 accumulators,
// four voxels in flight, atomically added into dW (zeroed by the launcher).  db[o] (optional) = sum of G over (b, v), from
// the CTAs of the first pair block.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int CW_CT = 16;
__global__ void __launch_bounds__(128, 4) conv_wgrad_kernel(ConvGeo G, const float* __restrict__ Gc, const float* __restrict__ F,
                                                         float* __restrict__ dW, float* __restrict__ db, int vchunk) {
  VX_PDL_ENTRY();
  VX_DYN_SMEM(float, gs);                              // [vchunk][16] gradients, then [vchunk] int4 (batch offset, z, y, x origins)
  int4* vo = reinterpret_cast<int4*>(gs + (size_t)vchunk * CW_CT);
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
  for (int vv = threadIdx.x; vv < vchunk; vv += blockDim.x) {
    const long long gv = c0 + (vv < nv ? vv : 0);
    const int b = (int)(gv / sv), v = (int)(gv % sv);
    vo[vv] = make_int4(b, (v / (G.w * G.h)) * G.s, ((v / G.w) % G.h) * G.s, (v % G.w) * G.s);
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
  const float* Fi = F + (size_t)i * SV;
  constexpr int CW_U = 8;                              // gathered fine-grid loads in flight per thread (4: 162 us for the k7 s4 stem)
#pragma unroll 1
  for (int v0 = 0; v0 < nv; v0 += CW_U) {
    float fv[CW_U];
#pragma unroll
    for (int u = 0; u < CW_U; ++u) {
      fv[u] = 0.f;
      if (v0 + u < nv) {
        const int4 o = vo[v0 + u];
        const int z = o.y + tz, y = o.z + ty, x = o.w + tx;
        if (z >= 0 && z < G.D && y >= 0 && y < G.H && x >= 0 && x < G.W) fv[u] = __ldg(Fi + (size_t)o.x * G.Ci * SV + ((size_t)z * G.H + y) * G.W + x);
      }
    }
#pragma unroll
    for (int u = 0; u < CW_U; ++u) {
      const float* gr = gs + (v0 + u) * CW_CT;         // rows beyond nv hold zeros
#pragma unroll
      for (int c = 0; c < CW_CT; c += 4) {
        const float4 g4 = *reinterpret_cast<const float4*>(gr + c);
        acc[c] = fmaf(g4.x, fv[u], acc[c]); acc[c + 1] = fmaf(g4.y, fv[u], acc[c + 1]);
        acc[c + 2] = fmaf(g4.z, fv[u], acc[c + 2]); acc[c + 3] = fmaf(g4.w, fv[u], acc[c + 3]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < CW_CT; ++c)
    if (o0 + c < G.Co) atomicAdd(dW + (size_t)(o0 + c) * G.so + (size_t)i * G.si + t, acc[c]);
}

// out[c] = sum over (b, s) of X[b, c, s]  (bias gradient of a transposed convolution): one CTA per channel, fixed fold order
__global__ void __launch_bounds__(256) channel_sum_kernel(const float* __restrict__ X, float* __restrict__ out, int B, int C, int S) {
  VX_PDL_ENTRY();
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
// channel tile and reduction slices: aim at >= ~64 K threads in flight, at most 8 slices and >= 8 reduction steps per slice.
// More slices do not pay: every CTA stages the whole weight tile, and at 256 K / 512 K threads the k7 s4 stem measured 113 us
// against 74 us (profiles/r3v_conv_target.txt; VX_CONV_TARGET_THREADS is the probe)
static long long conv_target_threads() {
  static long long t = [] { const char* e = getenv("VX_CONV_TARGET_THREADS"); return e ? atoll(e) : 65536LL; }();
  return t;
}
static void pick_tile_ks(long long nvox, int chans, int steps, int& ct, int& ks) {
  static const int ct_max = [] { const char* e = getenv("VX_CONV_CT_MAX"); return e ? atoi(e) : 16; }();      // probe
  ct = chans >= 16 ? 16 : chans >= 8 ? 8 : 4;
  if (ct > ct_max) ct = ct_max;
  while (ct > 4 && nvox * cdiv(chans, ct) < 32768) ct >>= 1;
  ks = 1;
  while (ks < 8 && nvox * cdiv(chans, ct) * ks < conv_target_threads() && steps / (2 * ks) >= 8) ks <<= 1;
}

template <int KW>
static void launch_strided(int ct, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const ConvGeo& G, const float* X, const float* Wt,
                           const float* bias, float* Y, int rchunk) {
  if (ct == 16) { VX_SET_SMEM((conv_strided_kernel<16, KW>), smem); VX_LAUNCH((conv_strided_kernel<16, KW>), grid, block, smem, st, G, X, Wt, bias, Y, rchunk); }
  else if (ct == 8) { VX_SET_SMEM((conv_strided_kernel<8, KW>), smem); VX_LAUNCH((conv_strided_kernel<8, KW>), grid, block, smem, st, G, X, Wt, bias, Y, rchunk); }
  else { VX_SET_SMEM((conv_strided_kernel<4, KW>), smem); VX_LAUNCH((conv_strided_kernel<4, KW>), grid, block, smem, st, G, X, Wt, bias, Y, rchunk); }
}

int conv_strided(const ConvGeo& G, const float* X, const float* Wt, const float* bias, float* Y, cudaStream_t st) {
  const long long nvox = (long long)G.B * G.d * G.h * G.w;
  const int R = G.Ci * G.k * G.k;
  int ct, ks;
  pick_tile_ks(nvox, G.Co, R, ct, ks);
  int rchunk = (46 * 1024 / 4) / (G.k * ct);             // the k7 s4 stem (98 rows x 7 x 16 channels = 43.9 KB) in one chunk
  if (rchunk > R) rchunk = R;
  if (rchunk < 1) rchunk = 1;
  size_t smem = sizeof(float) * (size_t)rchunk * G.k * ct;
  const int vw = 8 / ks;                               // 256 threads: KS reduction slices x VW voxel groups share the staged weights
  const size_t fold = sizeof(float) * (size_t)256 * ct;
  if (smem < fold) smem = fold;
  const dim3 grid(cdiv(nvox, 32 * vw), cdiv(G.Co, ct)), block(32, ks, vw);
  prof_bytes(4.0 * ((double)G.B * G.Ci * G.D * G.H * G.W + (double)nvox * G.Co + (double)G.Co * R * G.k));
  prof_flops(2.0 * (double)nvox * G.Co * R * G.k);
  if (G.k == 7) launch_strided<7>(ct, grid, block, smem, st, G, X, Wt, bias, Y, rchunk);
  else if (G.k == 3) launch_strided<3>(ct, grid, block, smem, st, G, X, Wt, bias, Y, rchunk);
  else if (G.k == 2) launch_strided<2>(ct, grid, block, smem, st, G, X, Wt, bias, Y, rchunk);
  else launch_strided<0>(ct, grid, block, smem, st, G, X, Wt, bias, Y, rchunk);
  return check_launch("conv_strided_kernel");
}

int conv_scatter(const ConvGeo& G, const float* X, const float* Wt, const float* bias, float* Y, cudaStream_t st) {
  const long long nfine = (long long)G.B * G.D * G.H * G.W;
  const int per = (G.k + G.s - 1) / G.s;
  if (per * per * per > 8) { set_error("conv: more than 8 taps reach a voxel (k%d s%d)", G.k, G.s); return VX_ERR_UNSUPPORTED; }
  int ct, ks;
  pick_tile_ks(nfine, G.Ci, G.Co, ct, ks);
  const int k3 = G.k * G.k * G.k;
  int ochunk = (40 * 1024 / 4) / (k3 * ct);
  if (ochunk < 1) ochunk = 1;
  if (ochunk > G.Co) ochunk = G.Co;
  size_t smem = sizeof(float) * (size_t)ochunk * k3 * ct;
  const int vw = 8 / ks;
  const size_t fold = sizeof(float) * (size_t)256 * ct;
  if (smem < fold) smem = fold;
  const dim3 grid(cdiv(nfine, 32 * vw), cdiv(G.Ci, ct)), block(32, ks, vw);
  prof_bytes(4.0 * ((double)nfine * G.Ci + (double)G.B * G.Co * G.d * G.h * G.w + (double)G.Co * G.Ci * k3));
  prof_flops(2.0 * (double)G.B * G.d * G.h * G.w * G.Co * G.Ci * k3);
  if (ct == 16) { VX_SET_SMEM(conv_scatter_kernel<16>, smem); VX_LAUNCH(conv_scatter_kernel<16>, grid, block, smem, st, G, X, Wt, bias, Y, ochunk); }
  else if (ct == 8) { VX_SET_SMEM(conv_scatter_kernel<8>, smem); VX_LAUNCH(conv_scatter_kernel<8>, grid, block, smem, st, G, X, Wt, bias, Y, ochunk); }
  else { VX_SET_SMEM(conv_scatter_kernel<4>, smem); VX_LAUNCH(conv_scatter_kernel<4>, grid, block, smem, st, G, X, Wt, bias, Y, ochunk); }
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
  vchunk = (vchunk + 7) & ~7LL;                        // whole rounds of CW_U voxels: the rows past nv are staged as zeros
  const int gz = cdiv(nvox, vchunk);
  const size_t smem = sizeof(float) * (size_t)vchunk * CW_CT + sizeof(int) * 4 * (size_t)vchunk;
  prof_bytes(4.0 * ((double)G.B * G.Ci * G.D * G.H * G.W + (double)nvox * G.Co + (double)G.Co * npair));
  prof_flops(2.0 * (double)nvox * G.Co * npair);
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

// Routes through the channel-contraction kernels (pointwise.cu / pw_tc.cu / pw_wgrad_tc.cu, tcgen05 3xTF32 from 512-1024 voxels):
//   1x1x1 convolution (deep-supervision heads):   y = W x + b          -- a contraction as it stands
//   ConvTranspose3d k = s = 2 without bias:        Y'[(co,t)] = W^T x   -- a contraction to 8 C_out channels followed by the
//        depth-to-space pass of stem.cu (pixel_shuffle with scale 2: channel index co*8 + t is exactly torch's (C_in, C_out, 2,2,2)
//        weight layout flattened); backward = space-to-depth of dy, then the two contractions
static bool pw_route(const vx_conv_desc* d, const float* bias) {
  if (!d || d->shuffle != 0) return false;
  if (!d->transposed) return d->kernel == 1 && d->stride == 1 && d->pad == 0;
  return d->kernel == 2 && d->stride == 2 && d->pad == 0 && bias == nullptr;
}
static size_t pw_route_ws(const vx_conv_desc* d) {
  if (!d || !d->transposed || d->kernel != 2 || d->stride != 2 || d->pad != 0 || d->shuffle != 0) return 0;
  return ((size_t)sizeof(float) * d->B * 8 * d->C_out * d->D * d->H * d->W + 255) & ~(size_t)255;
}

static int pw_conv_fwd(const vx_conv_desc* d, const float* x, const float* w, const float* bias, float* y, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int S = d->D * d->H * d->W;
  set_seed_dev(nullptr);
  PwBatch pb{}; pb.nprob = 1; pb.B = d->B; pb.S = S;
  PwProblem& p = pb.p[0];
  p.src[0] = PwSrc{x, d->C_in}; p.nsrc = 1; p.Ci = d->C_in; p.nseg = 1;
  if (!d->transposed) {
    p.seg[0] = PwSeg{w, bias, d->C_in, d->C_out, y}; p.Co = d->C_out;
    return pw_forward(pb, st);
  }
  if (!ws || ws_bytes < pw_route_ws(d)) { set_error("conv_fwd: workspace too small"); return VX_ERR_WORKSPACE; }
  float* yp = (float*)ws;                              // (B, 8 C_out, S)
  p.seg[0] = PwSeg{w, nullptr, 8 * d->C_out, d->C_in, yp}; p.Co = 8 * d->C_out; p.transposed = 1;
  int rc = pw_forward(pb, st);
  if (rc != VX_OK) return rc;
  vx_pixel_shuffle_desc ps{d->B, d->C_out, 2, d->D, d->H, d->W};
  const void* pin[2] = {yp, nullptr};
  void* pout[1] = {y};
  return vx_pixel_shuffle_fwd(&ps, pin, pout, (vx_stream_t)st);
}

static int pw_conv_bwd(const vx_conv_desc* d, const float* dy, const float* x, const float* w, float* dx, float* dw, float* db, void* ws,
                       size_t ws_bytes, cudaStream_t st) {
  const int S = d->D * d->H * d->W;
  set_seed_dev(nullptr);
  SideJoin side_guard(st);
  const float* g = dy;                                 // gradient at the contraction's output: (B, C_out, S) or (B, 8 C_out, S)
  int Cg = d->C_out;
  if (d->transposed) {
    if (!ws || ws_bytes < pw_route_ws(d)) { set_error("conv_bwd: workspace too small"); return VX_ERR_WORKSPACE; }
    vx_pixel_shuffle_desc ps{d->B, d->C_out, 2, d->D, d->H, d->W};
    const void* pin[1] = {dy};
    void* pout[2] = {ws, nullptr};
    const int rc = vx_pixel_shuffle_bwd(&ps, pin, pout, (vx_stream_t)st);
    if (rc != VX_OK) return rc;
    g = (const float*)ws; Cg = 8 * d->C_out;
  }
  // accumulated by the weight-gradient kernel on the side stream only: zeroed there
  { ZeroList zl; zl.add(dw, (size_t)d->C_in * Cg); if (db && !d->transposed) zl.add(db, d->C_out); const int rc = zero_many(zl, side_fork(st)); if (rc != VX_OK) return rc; }
  WgBatch wb{}; wb.nprob = 1; wb.B = d->B; wb.S = S;
  WgProblem& q = wb.p[0];
  if (!d->transposed) {      // dW (C_out, C_in) = dy x^T, db = sum dy
    q.dY = g; q.Co = d->C_out; q.src[0] = PwSrc{x, d->C_in}; q.nsrc = 1; q.Ci = d->C_in; q.dW = dw; q.ld = d->C_in; q.db = db;
  } else {                   // dW (C_in, 8 C_out) = x dY'^T
    q.dY = x; q.Co = d->C_in; q.src[0] = PwSrc{g, Cg}; q.nsrc = 1; q.Ci = Cg; q.dW = dw; q.ld = Cg; q.db = nullptr;
  }
  int rc = pw_wgrad(wb, st);
  if (rc != VX_OK) return rc;
  if (d->transposed && db) {
    VX_LAUNCH(channel_sum_kernel, dim3(d->C_out), dim3(256), 0, st, dy, db, d->B, d->C_out, 8 * S);
    rc = check_launch("channel_sum_kernel");
    if (rc != VX_OK) return rc;
  }
  if (dx) {
    PwBatch pb{}; pb.nprob = 1; pb.B = d->B; pb.S = S;
    PwProblem& p = pb.p[0];
    p.src[0] = PwSrc{g, Cg}; p.nsrc = 1; p.Ci = Cg; p.nseg = 1; p.Co = d->C_in;
    if (!d->transposed) { p.seg[0] = PwSeg{w, nullptr, d->C_in, d->C_out, dx}; p.transposed = 1; }      // dx = W^T dy
    else p.seg[0] = PwSeg{w, nullptr, Cg, d->C_in, dx};                                                   // dx = W dY'
    rc = pw_forward(pb, st);
    if (rc != VX_OK) return rc;
  }
  return VX_OK;
}

}  // namespace vx

using namespace vx;

extern "C" size_t vx_conv_workspace(const vx_conv_desc* d) { return conv3_tc_supported(d) ? conv3_tc_workspace(d) : pw_route_ws(d); }

extern "C" int vx_conv_fwd(const vx_conv_desc* d, const void* const* in, void* const* out, void* ws, size_t ws_bytes, vx_stream_t stream) {
  if (!d || !in || !out || !in[0] || !in[1] || !out[0]) { set_error("conv_fwd: null pointer"); return VX_ERR_BAD_DESC; }
  cudaStream_t st = (cudaStream_t)stream;
  const float* x = (const float*)in[0]; const float* w = (const float*)in[1]; const float* bias = (const float*)in[2];
  if (conv3_tc_supported(d)) return conv3_tc_fwd(d, x, w, bias, (float*)out[0], ws, ws_bytes, st);
  ConvGeo G{};
  if (!simt_geo(d, G)) { set_error("conv_fwd: unsupported geometry (k%d s%d p%d transposed %d shuffle %d)", d->kernel, d->stride, d->pad, d->transposed, d->shuffle); return VX_ERR_UNSUPPORTED; }
  prof_scope("conv_fwd %s B%d %d->%d k%d s%d %dx%dx%d", d->transposed ? "T" : "S", d->B, d->C_in, d->C_out, d->kernel, d->stride, d->D, d->H, d->W);
  if (pw_route(d, bias)) return pw_conv_fwd(d, x, w, bias, (float*)out[0], ws, ws_bytes, st);
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
  if (pw_route(d, nullptr) && (!d->transposed || ws)) return pw_conv_bwd(d, dy, x, w, dx, dw, db, ws, ws_bytes, st);
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
    if (!d->transposed && d->stride == 1) {
      // dx[ci][q] = sum W(co, ci, t) dy[co][q + pad - t] = a stride-1 convolution of dy with the taps reversed, channels exchanged
      ConvGeo F = G;
      F.Co = G.Ci; F.Ci = G.Co; F.so = G.si; F.si = G.so; F.flip = 1; F.pad = G.k - 1 - G.pad;
      F.d = G.D; F.h = G.H; F.w = G.W; F.D = G.d; F.H = G.h; F.W = G.w;
      rc = conv_strided(F, dy, w, nullptr, dx, st);
    } else {
      rc = d->transposed ? conv_strided(G, dy, w, nullptr, dx, st)     // dx[cin][v] = sum W(cin, cout, t) dy[cout][s v + t]
                         : conv_scatter(G, dy, w, nullptr, dx, st);    // dx[ci][q] = sum W(co, ci, t) dy[co][(q + pad - t)/s]
    }
    if (rc != VX_OK) return rc;
  }
  return VX_OK;
}
