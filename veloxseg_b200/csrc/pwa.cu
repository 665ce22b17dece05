// PWA block (reference: model/components/PWA.py:329-379 + 433-439, attention_utils.py:29-71) for sm_100a, fp32.
//
// Forward pipeline (M modality streams processed together, one launch per stage):
//   ln_forward        xhat = LN(x)                                    (pointwise.cu)
//   pw_forward        q,k,v = [Wq;Wk;Wv] (g*xhat + b) + bias          (one pass over xhat for all three)
//   pwa_gather        window partition + small-window max-pool -> tokens (B, h, Ns, M*l, c), arg-max kept
//   pwa_attn_fwd      per window: S = QK^T/sqrt(c) + bias, online softmax, (dropout), O = PV; scores never leave
//                     the SM (K/V of the window live in shared memory, one thread owns a query row)
//   pwa_scatter       per-window trilinear (align_corners) up-sampling of tokens back to NCDHW
//   pw_forward        y = 2x + Drop(Wmix a + b)                       (double residual of PWA.py:377,436)
//   ln_forward, pw_forward x2   z = y + Drop(W2 Drop(GELU(W1 LN(y) + b1)) + b2)
// Backward mirrors it; softmax is recomputed from the saved row log-sum-exp (no L x L tensor is ever stored).
#include "vx_kernels.h"
#include "vx_tc2.cuh"

#ifdef VX_EMU
#define __grid_constant__
#endif

#define VX_TRY(expr) do { int _rc = (expr); if (_rc != VX_OK) return _rc; } while (0)

namespace vx {
static inline size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }

struct PwaGeo {
  int B, M, D, H, W, S, heads, nb;
  int big[VX_MAX_SCALES][3], small[VX_MAX_SCALES][3];
  int Nw[VX_MAX_SCALES][3];   // windows per axis
  int Noff[VX_MAX_SCALES];    // first window index of the scale
  int vol[VX_MAX_SCALES];     // voxels per small window
  int Ns, n[3], l, L;
};

static int make_geo(const vx_pwa_desc* d, PwaGeo& G) {
  if (!d || d->B <= 0 || d->M <= 0 || d->M > VX_MAX_MODAL || d->C <= 0 || d->heads <= 0 || d->n_scales <= 0 ||
      d->n_scales > VX_MAX_SCALES) { set_error("pwa: bad descriptor"); return VX_ERR_BAD_DESC; }
  G.B = d->B; G.M = d->M; G.D = d->D; G.H = d->H; G.W = d->W; G.S = d->D * d->H * d->W;
  G.heads = d->heads; G.nb = d->n_scales;
  const int dims[3] = {d->D, d->H, d->W};
  int off = 0;
  for (int j = 0; j < G.nb; ++j) {
    int vol = 1;
    for (int a = 0; a < 3; ++a) {
      G.big[j][a] = d->big[j][a]; G.small[j][a] = d->small[j][a];
      if (G.big[j][a] <= 0 || G.small[j][a] <= 0 || G.big[j][a] % G.small[j][a] || dims[a] % G.big[j][a]) {
        set_error("pwa: window %d axis %d (big %d small %d) does not tile extent %d", j, a, G.big[j][a], G.small[j][a], dims[a]);
        return VX_ERR_BAD_DESC;
      }
      G.Nw[j][a] = dims[a] / G.big[j][a];
      const int na = G.big[j][a] / G.small[j][a];
      if (j == 0) G.n[a] = na;
      else if (G.n[a] != na) { set_error("pwa: scales disagree on tokens per window"); return VX_ERR_BAD_DESC; }
      vol *= G.small[j][a];
    }
    G.vol[j] = vol;
    G.Noff[j] = off;
    off += G.Nw[j][0] * G.Nw[j][1] * G.Nw[j][2];
  }
  G.Ns = off;
  G.l = G.n[0] * G.n[1] * G.n[2];
  G.L = G.l * G.M;
  if (d->c_qk % (G.nb * G.heads) || d->c_v % (G.nb * G.heads)) { set_error("pwa: channels do not split into scales x heads"); return VX_ERR_BAD_DESC; }
  // the gather / scatter kernels decode flat indices in 32-bit arithmetic
  const long long cmax = d->c_qk > d->c_v ? d->c_qk : d->c_v;
  if ((long long)G.B * G.S * cmax * (G.nb > 1 ? G.nb : 1) >= (1LL << 31)) { set_error("pwa: more than 2^31 elements per tensor"); return VX_ERR_UNSUPPORTED; }
  return VX_OK;
}

// ---------------------------------------------------------------------------------------------------
// gather: tokens + arg-max.  grid (blocks, scale, kind*M); small windows of >= 32 voxels use a warp per element.
// ---------------------------------------------------------------------------------------------------
struct GatherArgs {
  const float* src[3][VX_MAX_MODAL];   // [kind][m] (B, Ct, S)
  float* tok[3];                       // (B, h, Ns, L, cper)
  int* arg[3];
  int Ct[3], cper[3];
  int nkind;
};


// Tokens are enumerated by their position (tz, ty, tx) in the token grid of the scale (extent / small window per axis): the
// window index and the token coordinate inside the window of every grid position are tabulated per axis in shared memory
// once per CTA, so a token costs two divisions (flat index -> tz, ty, tx) instead of ~10 (round 1 decoded a 64-bit flat index
// with a handful of 64-bit divisions per token, and the warp mode three more per VOXEL).
// Thread mode (small windows of < 32 voxels): a thread owns one token and walks its cper channels; reads are runs along x
// (the token's small window, next to its neighbour's), a token row (cper floats / indices) is written as 16-byte vectors.
// Warp mode (>= 32 voxels per small window): a warp per token, lanes over the window, channels in turn.
// grid: x = (b, head) planes x token chunks, y = scale, z = kind * M + modality.
constexpr int GA_MAX_AXIS = 3 * 160;
__global__ void __launch_bounds__(256) pwa_gather_kernel(const __grid_constant__ PwaGeo G, const __grid_constant__ GatherArgs A,
                                                         int chunks) {
  VX_PDL_ENTRY();
  const int j = blockIdx.y;
  const int kind = blockIdx.z / G.M, m = blockIdx.z % G.M;
  const int cper = A.cper[kind], Ct = A.Ct[kind];
  const int plane = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
  const int b = plane / G.heads, head = plane % G.heads;
  const bool warp_mode = G.vol[j] >= 32;
  const int lane = threadIdx.x & 31;
  const int s0 = G.small[j][0], s1 = G.small[j][1], s2 = G.small[j][2];
  const int T0 = G.D / s0, T1 = G.H / s1, T2 = G.W / s2;            // token grid of this scale
  __shared__ int t_w[GA_MAX_AXIS], t_a[GA_MAX_AXIS];
  {
    const int toff[3] = {0, T0, T0 + T1};
    for (int i = threadIdx.x; i < T0 + T1 + T2; i += blockDim.x) {
      const int ax = i < T0 ? 0 : (i < T0 + T1 ? 1 : 2);
      const int p = i - toff[ax];
      const int w = p / G.n[ax];
      t_w[i] = w; t_a[i] = p - w * G.n[ax];
    }
  }
  __syncthreads();
  const unsigned ntok = (unsigned)(T0 * T1 * T2), T12 = (unsigned)(T1 * T2);
  const float* p0 = A.src[kind][m] + ((size_t)b * Ct + (size_t)(j * G.heads + head) * cper) * G.S;
  const size_t obase = ((size_t)b * G.heads + head) * G.Ns + G.Noff[j];
  auto token_out = [&](unsigned u, int& idx0) -> size_t {
    const unsigned tz = u / T12, rem = u - tz * T12;
    const unsigned ty = rem / (unsigned)T2, tx = rem - ty * (unsigned)T2;
    const int iz = (int)tz, iy = T0 + (int)ty, ix = T0 + T1 + (int)tx;
    const int Nloc = (t_w[iz] * G.Nw[j][1] + t_w[iy]) * G.Nw[j][2] + t_w[ix];
    const int t = (t_a[iz] * G.n[1] + t_a[iy]) * G.n[2] + t_a[ix];
    idx0 = ((int)tz * s0 * G.H + (int)ty * s1) * G.W + (int)tx * s2;
    return ((obase + Nloc) * G.L + (size_t)m * G.l + t) * cper;
  };
  if (!warp_mode) {
    for (unsigned u = chunk * blockDim.x + threadIdx.x; u < ntok; u += chunks * blockDim.x) {
      int idx0;
      const size_t o = token_out(u, idx0);
      for (int c4 = 0; c4 < cper; c4 += 4) {
        float bv[4];
        int bi[4];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const float* p = p0 + (size_t)(c4 + cc) * G.S;
          float best = -INFINITY;
          int bidx = idx0;
          if (c4 + cc < cper) {
            for (int dz = 0; dz < s0; ++dz)
              for (int dy = 0; dy < s1; ++dy)
                for (int dx = 0; dx < s2; ++dx) {
                  const int idx = idx0 + (dz * G.H + dy) * G.W + dx;
                  const float v = __ldg(p + idx);
                  if (v > best) { best = v; bidx = idx; }
                }
          }
          bv[cc] = best; bi[cc] = bidx;
        }
        if (c4 + 3 < cper) {
          *reinterpret_cast<float4*>(A.tok[kind] + o + c4) = make_float4(bv[0], bv[1], bv[2], bv[3]);
          if (A.arg[kind]) *reinterpret_cast<int4*>(A.arg[kind] + o + c4) = make_int4(bi[0], bi[1], bi[2], bi[3]);
        } else {
          for (int cc = 0; cc < 4 && c4 + cc < cper; ++cc) {
            A.tok[kind][o + c4 + cc] = bv[cc];
            if (A.arg[kind]) A.arg[kind][o + c4 + cc] = bi[cc];
          }
        }
      }
    }
    return;
  }
  // warp mode: lane q of the window -> (dz, dy, dx); shifts when the small window is a power of two per axis (the usual case)
  const bool pow2 = !(s1 & (s1 - 1)) && !(s2 & (s2 - 1));
  const int l2 = 31 - __clz(s2 > 0 ? s2 : 1), l1 = 31 - __clz(s1 > 0 ? s1 : 1);
  const int warps = blockDim.x >> 5, wid = threadIdx.x >> 5;
  // a warp per (token, channel): the window walk of one channel is the serial chain, so channels go to different warps
  for (unsigned uc = chunk * warps + wid; uc < ntok * (unsigned)cper; uc += chunks * warps) {
    const unsigned u = uc / (unsigned)cper;
    const int c = (int)(uc - u * (unsigned)cper);
    int idx0;
    const size_t o = token_out(u, idx0);
    const float* p = p0 + (size_t)c * G.S;
    float best = -INFINITY;
    int myidx = 0x7fffffff;
    for (int q = lane; q < G.vol[j]; q += 32) {
      int dx, dy, dz;
      if (pow2) { dx = q & (s2 - 1); dy = (q >> l2) & (s1 - 1); dz = q >> (l2 + l1); }
      else { dx = q % s2; dy = (q / s2) % s1; dz = q / (s2 * s1); }
      const int idx = idx0 + (dz * G.H + dy) * G.W + dx;
      const float v = __ldg(p + idx);
      if (v > best) { best = v; myidx = idx; }
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, sft);
      const int oi = __shfl_xor_sync(0xffffffffu, myidx, sft);
      if (ov > best || (ov == best && oi < myidx)) { best = ov; myidx = oi; }
    }
    if (lane == 0) {
      A.tok[kind][o + c] = best;
      if (A.arg[kind]) A.arg[kind][o + c] = myidx;
    }
  }
}

static int launch_gather(const PwaGeo& G, const GatherArgs& A, cudaStream_t st) {
  // chunks of the token list per (b, head) plane: sized for the busiest scale (thread mode: a thread per token, warp mode: a
  // warp per token), capped so that the whole grid stays around 16 CTAs per SM
  long long maxwork = 1;
  for (int j = 0; j < G.nb; ++j) {
    const long long ntok = (long long)(G.D / G.small[j][0]) * (G.H / G.small[j][1]) * (G.W / G.small[j][2]);
    if ((G.D / G.small[j][0]) + (G.H / G.small[j][1]) + (G.W / G.small[j][2]) > GA_MAX_AXIS) { set_error("pwa gather: extent beyond the kernel's tables"); return VX_ERR_UNSUPPORTED; }
    long long cmax = 1;
    for (int k = 0; k < A.nkind; ++k) cmax = A.cper[k] > cmax ? A.cper[k] : cmax;
    const long long work = G.vol[j] >= 32 ? ntok * cmax * 32 : ntok;      // warp mode: a warp per (token, channel)
    maxwork = work > maxwork ? work : maxwork;
  }
  const int planes = G.B * G.heads;
  int chunks = cdiv(maxwork, 256);
  const long long cap = (16LL * kSMs) / ((long long)planes * G.nb * A.nkind * G.M);
  if (chunks > cap) chunks = (int)(cap > 1 ? cap : 1);
  if (chunks < 1) chunks = 1;
  VX_LAUNCH(pwa_gather_kernel, dim3(chunks * planes, G.nb, A.nkind * G.M), dim3(256), 0, st, G, A, chunks);
  return check_launch("pwa_gather_kernel");
}

constexpr int GB_MAX_AXIS = 3 * 160;       // D + H + W entries of the per-axis tables
// backward of the gather: full-resolution gradient; a voxel receives its token's gradient iff it was the arg-max.
struct GatherBwdArgs {
  const float* dtok[3]; const int* arg[3];
  float* dst[3][VX_MAX_MODAL];   // (B, Ct, S)
  int Ct[3], cper[3];
};

__global__ void __launch_bounds__(256) pwa_gather_bwd_kernel(const __grid_constant__ PwaGeo G, const __grid_constant__ GatherBwdArgs A) {
  VX_PDL_ENTRY();
  // CTA = one (batch, scale, head) x a slab of voxels; thread = voxel: one 16-byte read of the token's gradient row (and
  // arg-max row) per 4 channels, channel planes written with the lanes along x.  Window index and token coordinate of every
  // position along each axis are tabulated in shared memory once per CTA (two divisions per voxel instead of ~15).
  const int kind = blockIdx.z / G.M, m = blockIdx.z % G.M;
  const int cper = A.cper[kind], Ct = A.Ct[kind];
  const int head = blockIdx.y % G.heads, j = (blockIdx.y / G.heads) % G.nb, b = blockIdx.y / (G.heads * G.nb);
  __shared__ int t_w[GB_MAX_AXIS], t_t[GB_MAX_AXIS];
  const int aoff[3] = {0, G.D, G.D + G.H};
  for (int i = threadIdx.x; i < G.D + G.H + G.W; i += blockDim.x) {
    const int ax = i < G.D ? 0 : (i < G.D + G.H ? 1 : 2);
    const int p = i - aoff[ax];
    const int big = G.big[j][ax];
    const int w = p / big;
    t_w[i] = w; t_t[i] = (p - w * big) / G.small[j][ax];
  }
  __syncthreads();
  float* dst = A.dst[kind][m];
  const unsigned HW = (unsigned)(G.H * G.W);
  for (unsigned uidx = blockIdx.x * blockDim.x + threadIdx.x; uidx < (unsigned)G.S; uidx += gridDim.x * blockDim.x) {
    const int idx = (int)uidx;
    const unsigned z = uidx / HW, rem = uidx - z * HW;
    const unsigned y = rem / (unsigned)G.W, x = rem - y * (unsigned)G.W;
    const int iz = (int)z, iy = G.D + (int)y, ix = G.D + G.H + (int)x;
    const int Nloc = (t_w[iz] * G.Nw[j][1] + t_w[iy]) * G.Nw[j][2] + t_w[ix];
    const int t = (t_t[iz] * G.n[1] + t_t[iy]) * G.n[2] + t_t[ix];
    const size_t o = ((((size_t)b * G.heads + head) * G.Ns + G.Noff[j] + Nloc) * G.L + (size_t)m * G.l + t) * cper;
    float* dp = dst + ((size_t)b * Ct + (size_t)(j * G.heads + head) * cper) * G.S + idx;
    const bool all = G.vol[j] == 1;
    for (int c4 = 0; c4 < cper; c4 += 4) {
      float g[4] = {0.f, 0.f, 0.f, 0.f};
      if (c4 + 3 < cper) {
        const float4 gv = __ldg(reinterpret_cast<const float4*>(A.dtok[kind] + o + c4));
        g[0] = gv.x; g[1] = gv.y; g[2] = gv.z; g[3] = gv.w;
        if (!all) {
          const int4 av = __ldg(reinterpret_cast<const int4*>(A.arg[kind] + o + c4));
          if (av.x != idx) g[0] = 0.f;
          if (av.y != idx) g[1] = 0.f;
          if (av.z != idx) g[2] = 0.f;
          if (av.w != idx) g[3] = 0.f;
        }
      } else {
        for (int q = 0; q < 4 && c4 + q < cper; ++q)
          g[q] = (all || A.arg[kind][o + c4 + q] == idx) ? A.dtok[kind][o + c4 + q] : 0.f;
      }
      for (int q = 0; q < 4 && c4 + q < cper; ++q) dp[(size_t)(c4 + q) * G.S] = g[q];
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// relative-position bias:  biasT[h][tk][tq] = table[index[tq][tk]][h]      (attention_utils.py:120-125)
// ---------------------------------------------------------------------------------------------------
__global__ void pwa_bias_kernel(const float* __restrict__ table, const long long* __restrict__ index,
                                float* __restrict__ biasT, float* __restrict__ biasN, int heads, int l) {
  VX_PDL_ENTRY();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= heads * l * l) return;
  const int tq = e % l, tk = (e / l) % l, h = e / (l * l);
  const float v = table[(size_t)index[(size_t)tq * l + tk] * heads + h];
  biasT[e] = v;
  if (biasN) biasN[((size_t)h * l + tq) * l + tk] = v;      // query-major copy: coalesced for the key-row phase of the backward
}

// dtable[index[tq][tk]][h] += dbias[h][tq][tk]     (the gradient buffer is query-major, unlike biasT)
// l*l contributions fold onto prod(2n-1) table rows (35 per row at level 2): a CTA histograms its share in shared memory
// and flushes one atomic per row, instead of l*l contended global atomics.
__global__ void __launch_bounds__(256) pwa_bias_bwd_kernel(const float* __restrict__ dbias, const long long* __restrict__ index,
                                                           float* __restrict__ dtable, int heads, int l, int rows) {
  VX_PDL_ENTRY();
  VX_DYN_SMEM(float, hist);
  const int h = blockIdx.y;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) hist[i] = 0.f;
  __syncthreads();
  const int n = l * l;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x)
    atomicAdd(hist + (int)index[e], dbias[(size_t)h * n + e]);
  __syncthreads();
  for (int i = threadIdx.x; i < rows; i += blockDim.x) {
    const float v = hist[i];
    if (v != 0.f) atomicAdd(dtable + (size_t)i * heads + h, v);
  }
}

// ---------------------------------------------------------------------------------------------------
// attention
// ---------------------------------------------------------------------------------------------------
struct AttnArgs {
  const float* Q; const float* K; const float* V; const float* biasT; const float* biasN;
  float* O; float* lse;
  // backward
  const float* dO; float* dQ; float* dK; float* dV; float* dbiasT;
  int B, heads, Ns, L, l;
  int smem_bias, wpc;      // backward: bias gradient accumulated in shared memory over wpc windows per CTA
  float scale, drop_p;
  uint64_t seed;
  const unsigned long long* seed_dev;
};

constexpr int ATT_TS = 4;            // threads per attention row (the key / query loop is dealt to them in turns)
constexpr int ATT_MAX_THREADS = 512;
constexpr uint32_t ATT_SITE = 7;

// keep-scales of 4 consecutive keys of one (window,row)
VX_DEV void attn_drop4(const AttnArgs& A, size_t row, int k4, float inv_keep, float (&ms)[4]) {
  const int nk4 = (A.L + 3) >> 2;
  const uint64_t soff = A.seed_dev ? (uint64_t)__ldg(A.seed_dev) : 0;
  const uint4 r = rng4(A.seed + soff, (uint64_t)row * nk4 + k4, ATT_SITE);
  const uint32_t bits[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) ms[i] = ((float)(bits[i] >> 8) * (1.0f / 16777216.0f) < A.drop_p) ? 0.f : inv_keep;
}

// keep-scale of key k of one (window,row): the same word attn_drop4 hands out for it
VX_DEV float attn_drop1(const AttnArgs& A, uint32_t key, size_t row, int k, float inv_keep) {
  const int nk4 = (A.L + 3) >> 2;
  return keep_from_bits(rng_word(key, ((uint64_t)row * nk4 + (k >> 2)) * 4 + (k & 3)), A.drop_p, inv_keep);
}

// K / V / Q / dO tiles in shared memory are [row][C] with 8 floats of padding after every 4 rows: the ATT_TS threads of a
// query row walk keys 4 apart (one quad each), which without the skew is a 4-way bank conflict on every K / V read
// (profiles/r1c_pwa_L2.digest.txt: 2.3e7 conflicts, L1 at 76 % of peak).  Row stride in floats: C, quad stride: 4 C + 8.
VX_DEV int att_row(int k, int C) { return k * C + (k >> 2) * 8; }
VX_DEV int att_rows_floats(int L, int C) { return L * C + ((L + 3) >> 2) * 8; }
// exp for the softmax terms: ex2.approx on the pre-scaled argument (relative error ~1e-6 for |x| < 20, far inside the
// fp32 parity bar); the precise expf is ~15 instructions of a ~40-instruction inner loop.
VX_DEV float att_exp(float x) {
#ifdef VX_EMU
  return expf(x);
#else
  return __expf(x);
#endif
}

VX_DEV void att_stage(float* dst, const float* src, int L, int C, int tid, int nthr) {
  for (int i = tid; i < L * C; i += nthr) { const int k = i / C; vx_cp_async4(dst + att_row(k, C) + (i - k * C), src + i, true); }
}
VX_DEV void att_stage_flat(float* dst, const float* src, int n, int tid, int nthr) {
  for (int i = tid; i < n; i += nthr) vx_cp_async4(dst + i, src + i, true);
}

// dbias[tq][tk0 .. tk0+3] += v (16-byte vector reduction when the row is 4-element granular)
VX_DEV void att_red4(float* p, const float (&v)[4]) {
#ifndef VX_EMU
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
#else
  for (int i = 0; i < 4; ++i) atomicAdd(p + i, v[i]);
#endif
}

// Forward: CTA = one window x a block of query rows; K / V of the window in shared memory; ATT_TS threads share a query
// row (each takes every ATT_TS-th quad of keys with its own online-softmax state, merged with two shuffle rounds).
template <int CQ, int CV>
__global__ void __launch_bounds__(ATT_MAX_THREADS) pwa_attn_fwd_kernel(const __grid_constant__ AttnArgs A) {
  VX_PDL_ENTRY();
  const int N = blockIdx.y, bh = blockIdx.z, head = bh % A.heads;
  const int L = A.L, l = A.l;
  const int tid = threadIdx.x, nthr = blockDim.x;
  VX_DYN_SMEM(float, sm);
  float* Ks = sm;                            // [L][CQ] (padded, att_row)
  float* Vs = sm + att_rows_floats(L, CQ);   // [L][CV]
  const size_t wbase = (size_t)bh * A.Ns + N;
  att_stage(Ks, A.K + wbase * L * CQ, L, CQ, tid, nthr);
  att_stage(Vs, A.V + wbase * L * CV, L, CV, tid, nthr);
  vx_cp_async_commit();
  const int rows = nthr / ATT_TS;
  const int i_raw = blockIdx.x * rows + tid / ATT_TS, t = tid % ATT_TS;
  const bool live = i_raw < L;
  const int i = live ? i_raw : L - 1;
  float q[CQ];
#pragma unroll
  for (int c = 0; c < CQ; ++c) q[c] = __ldg(A.Q + (wbase * L + i) * CQ + c) * A.scale;
  vx_cp_async_wait_all();
  __syncthreads();
  const float* bT = A.biasT + (size_t)head * l * l + (i % l);
  const bool drop = A.drop_p > 0.f;
  const float inv_keep = drop ? 1.0f / (1.0f - A.drop_p) : 1.f;
  const size_t row = wbase * L + i;
  float mx = -INFINITY, ssum = 0.f;
  float acc[CV];
#pragma unroll
  for (int c = 0; c < CV; ++c) acc[c] = 0.f;
  const int nk4 = (L + 3) >> 2;
  for (int qd = t; qd < nk4; qd += ATT_TS) {
    const int k0 = qd * 4;
    const int tk0 = k0 % l;
    float ms[4] = {1.f, 1.f, 1.f, 1.f};
    if (drop) attn_drop4(A, row, qd, inv_keep, ms);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int k = k0 + kk;
      if (k >= L) break;
      const float* kr = Ks + att_row(k, CQ);
      const float* vr = Vs + att_row(k, CV);
      int tk = tk0 + kk;
      while (tk >= l) tk -= l;
      float s = __ldg(bT + (size_t)tk * l);
#pragma unroll
      for (int c = 0; c < CQ; ++c) s = fmaf(q[c], kr[c], s);
      if (s > mx) {
        const float corr = att_exp(mx - s);
        ssum *= corr;
#pragma unroll
        for (int c = 0; c < CV; ++c) acc[c] *= corr;
        mx = s;
      }
      const float p = att_exp(s - mx);
      ssum += p;
      const float pm = p * ms[kk];
#pragma unroll
      for (int c = 0; c < CV; ++c) acc[c] = fmaf(pm, vr[c], acc[c]);
    }
  }
  // merge the ATT_TS partial softmax states of the row
#pragma unroll
  for (int o = 1; o < ATT_TS; o <<= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, mx, o);
    const float s2 = __shfl_xor_sync(0xffffffffu, ssum, o);
    const float mn = fmaxf(mx, m2);
    const float c1 = mx == -INFINITY ? 0.f : att_exp(mx - mn), c2 = m2 == -INFINITY ? 0.f : att_exp(m2 - mn);
    ssum = ssum * c1 + s2 * c2;
#pragma unroll
    for (int c = 0; c < CV; ++c) {
      const float a2 = __shfl_xor_sync(0xffffffffu, acc[c], o);
      acc[c] = acc[c] * c1 + a2 * c2;
    }
    mx = mn;
  }
  if (live && t == 0) {
    const float inv = 1.0f / ssum;
#pragma unroll
    for (int c = 0; c < CV; ++c) A.O[row * CV + c] = acc[c] * inv;
    A.lse[row] = mx + logf(ssum);
  }
}

// ---------------------------------------------------------------------------------------------------
// Forward on the tensor cores (PWA.py:308-327 at the levels whose windows are GEMM-sized: L >= 128, L % 16 == 0, 8 channels per
// head -- level 2 of the three configurations, L = 432).  CTA = one window x 128 query rows, 128 threads, thread = query row =
// TMEM lane.  3xTF32 throughout (fp32-accurate scores: the backward recomputes P in fp32 from the saved log-sum-exp):
//   S = Q K^T       A = [Q_hi | Q_lo | Q_hi] from shared memory (K-major), B = [K_hi | K_hi | K_lo]: 3 k-steps of one N = chunk MMA;
//   softmax         two rolled passes over the chunk's TMEM columns in pieces of 16: bias + running max (scores written back),
//                   then p = exp(s - m): row sum, dropout mask, hi / lo split written to TMEM -- P never visits shared memory;
//   O += P V        the ".ts" form: A = P_hi / P_lo from TMEM, B = [V_hi | V_lo] as ONE N = 16 operand (8 + 8 channels), so the
//                   N = 16 floor of the instruction carries the hi / lo products instead of padding: 2 MMAs per 8 keys;
//                   the chunk's 16 columns come back to registers where the online-softmax rescaling lives.
// Keys are processed in chunks of ATC_CHUNK = 112 (TMEM: S / P_hi 112 + P_lo 112 + O 2 x 16 = the 256 allocated columns, so
// two CTAs share an SM: one CTA's softmax runs under the other's MMAs).
// ---------------------------------------------------------------------------------------------------
constexpr int ATC_CHUNK = 112, ATC_ROWS = 128, ATC_TS = 4, ATC_THREADS = ATC_ROWS * ATC_TS, ATC_COLS = 256, ATC_C = 8;
constexpr uint32_t ATC_PLO = ATC_CHUNK, ATC_O = 2 * ATC_CHUNK;

// K-major SWIZZLE_NONE image of an (rows x 8) operand: 8-row groups 256 B apart, the two 16-byte K halves 128 B apart
VX_DEV int atc_kmajor(int r, int k) { return (r >> 3) * 64 + (k >> 2) * 32 + (r & 7) * 4 + (k & 3); }      // in floats

// bias of 8 consecutive keys (from key k of the window, k % 8 == 0, l % 8 == 0) of one query token's row of the query-major table
VX_DEV void atc_bias8(const float* __restrict__ brow, int k, int l, float (&b)[8]) {
  int tk = k;
  while (tk >= l) tk -= l;
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(brow + tk)), b1 = __ldg(reinterpret_cast<const float4*>(brow + tk) + 1);
  b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
}
// 2^x (the scores are kept in log2 units: the 1/ln 2 factor rides on the query scaling and the bias FMA)
VX_DEV float atc_exp2(float x) {
#ifdef VX_EMU
  return exp2f(x);
#else
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}

// ATC_TS threads per query row: warp w works on TMEM lanes 32 (w % 4) .. + 31 (the hardware's lane window of a warp) and takes
// every ATC_TS-th piece of 8 key columns, share = w / 4.  Per chunk: partial row maxima meet in shared memory, the partial row
// sums only at the end (every share uses the same reference maximum).
__global__ void __launch_bounds__(ATC_THREADS, 2) pwa_attn_fwd_tc_kernel(const __grid_constant__ AttnArgs A) {
  VX_PDL_ENTRY();
  const int N = blockIdx.y, bh = blockIdx.z, head = bh % A.heads;
  const int L = A.L, l = A.l;
  const int tid = threadIdx.x, warp = tid >> 5, lg = warp & 3, sh = warp >> 2;
  const int rt = lg * 32 + (tid & 31);       // row inside the tile = TMEM lane
  VX_DYN_SMEM(float, sm);
  float* Qhi = sm;                          // [128 x 8] K-major
  float* Qlo = Qhi + ATC_ROWS * ATC_C;
  float* red = Qlo + ATC_ROWS * ATC_C;      // [ATC_TS][128] partial maxima / sums
  float* Khi = red + ATC_TS * ATC_ROWS;     // [L x 8] K-major (rows = keys)
  float* Klo = Khi + (size_t)L * ATC_C;
  float* Vt = Klo + (size_t)L * ATC_C;      // [L / 8 k-steps][16 rows: V_hi channels | V_lo channels][8 keys] K-major
  VX_TC_SHARED_BARS(bars, 1);               // scores of the chunk (and the output columns of the one before) ready
  VX_TC_SHARED_SLOT(tmem_slot);
  if (warp == 0) tc::tmem_alloc(&tmem_slot, (uint32_t)ATC_COLS);
  if (tid == 0) {
    tc::mbar_init(&bars[0], 1);
    tc::mbar_init_fence();
  }
  const size_t wbase = (size_t)bh * A.Ns + N;
  const int i_raw = blockIdx.x * ATC_ROWS + rt;
  const bool live = i_raw < L;
  const int i = live ? i_raw : L - 1;
  // ---- operand staging
  if (sh == 0) {
    const float4* qp = reinterpret_cast<const float4*>(A.Q + (wbase * L + i) * ATC_C);
    float4 q[2] = {__ldg(qp), __ldg(qp + 1)};
    if (!live) q[0] = q[1] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float qs = A.scale * 1.4426950408889634f;      // scores in log2 units
      const float v[4] = {q[h].x * qs, q[h].y * qs, q[h].z * qs, q[h].w * qs};
      float hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) tc::split(v[j], hi[j], lo[j]);
      const int o = atc_kmajor(rt, 4 * h);
      *reinterpret_cast<float4*>(Qhi + o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<float4*>(Qlo + o) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
  {
    // item = (key, half of the 8 channels)
    const float* Kg = A.K + wbase * L * ATC_C;
    const float* Vg = A.V + wbase * L * ATC_C;
    for (int e = tid; e < 2 * L; e += ATC_THREADS) {
      const int r = e >> 1, h = e & 1;
      const float4 k4 = __ldg(reinterpret_cast<const float4*>(Kg + (size_t)r * ATC_C) + h);
      const float4 v4 = __ldg(reinterpret_cast<const float4*>(Vg + (size_t)r * ATC_C) + h);
      const float kv[4] = {k4.x, k4.y, k4.z, k4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
      float hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) tc::split(kv[j], hi[j], lo[j]);
      const int o = atc_kmajor(r, 4 * h);
      *reinterpret_cast<float4*>(Khi + o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<float4*>(Klo + o) = make_float4(lo[0], lo[1], lo[2], lo[3]);
      // V transposed: operand row n = channel (hi) / 8 + channel (lo), k = key inside the k-step
      float* vt = Vt + (size_t)(r >> 3) * 128;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float vh, vl;
        tc::split(vv[j], vh, vl);
        vt[atc_kmajor(4 * h + j, r & 7)] = vh;
        vt[atc_kmajor(8 + 4 * h + j, r & 7)] = vl;
      }
    }
  }
  tc::fence_async_smem();
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t trow = tmem + ((uint32_t)(lg * 32) << 16);       // this thread's lane, column 0
  const uint32_t qh = tc::smem_addr(Qhi), ql = tc::smem_addr(Qlo), kh = tc::smem_addr(Khi), kl = tc::smem_addr(Klo), vt = tc::smem_addr(Vt);

  const float* brow = A.biasN + ((size_t)head * l + (i % l)) * l;
  const bool drop = A.drop_p > 0.f;
  const float inv_keep = drop ? 1.0f / (1.0f - A.drop_p) : 1.f;
  const size_t row = wbase * L + i;
  constexpr float LOG2E = 1.4426950408889634f;
  float mx = -INFINITY, ssum = 0.f;          // log2 units; ssum: this share's part of the row sum
  float corr = 0.f;                          // rescaling of the chunks before the current one
  float acc[ATC_C];                          // share 0 only
#pragma unroll
  for (int c = 0; c < ATC_C; ++c) acc[c] = 0.f;

  // thread 0 issues, in order: scores of chunk 0 | P V of chunk 0, scores of chunk 1 | ... (tcgen05.mma executes in issue order,
  // so the score MMA that overwrites the P_hi columns may follow the P V MMAs that read them without a wait in between);
  // one commit per group: "scores of chunk n ready" also says "output columns of chunk n - 1 ready".
  auto issue_scores = [&](int k0) {
    const int n = L - k0 < ATC_CHUNK ? L - k0 : ATC_CHUNK;
    const uint32_t idesc = tc::idesc_tf32(n, 0, 0);
    const uint32_t ko = (uint32_t)(k0 >> 3) * 256u;
    const uint64_t aqh = tc::desc(qh, 128u, 256u), aql = tc::desc(ql, 128u, 256u);
    const uint64_t bkh = tc::desc(kh + ko, 128u, 256u), bkl = tc::desc(kl + ko, 128u, 256u);
    tc::mma_tf32(tmem, aql, bkh, idesc, 0u);
    tc::mma_tf32(tmem, aqh, bkl, idesc, 1u);
    tc::mma_tf32(tmem, aqh, bkh, idesc, 1u);
  };
  auto take_output = [&]() {                 // acc = acc * corr + O  (share 0, after the commit that covers the P V MMAs)
    float o[32];
    tc::tmem_ld32(trow + ATC_O, o);
#pragma unroll
    for (int c = 0; c < ATC_C; ++c) acc[c] = fmaf(acc[c], corr, (o[c] + o[8 + c]) + (o[16 + c] + o[24 + c]));
  };
  if (tid == 0) { issue_scores(0); tc::commit(&bars[0]); }

  int chunk_i = 0;
#pragma unroll 1
  for (int k0 = 0; k0 < L; k0 += ATC_CHUNK, ++chunk_i) {
    const int n = L - k0 < ATC_CHUNK ? L - k0 : ATC_CHUNK;           // a multiple of 16
    const int np = n >> 3;
    // the bias of the first piece travels under the score MMA; inside the loop the next piece's under the current one
    float bn[8];
    if (sh < np) atc_bias8(brow, k0 + 8 * sh, l, bn);
    if (warp == 0) tc::mbar_wait(&bars[0], (uint32_t)(chunk_i & 1));      // one warp polls, the others sleep at the barrier
    __syncthreads();
    tc::fence_after();
    if (sh == 0 && chunk_i > 0) take_output();
    // pass 1: bias, maximum; the biased scores go back to their columns
    float cm = -INFINITY;
#pragma unroll 1
    for (int p = sh; p < np; p += ATC_TS) {
      float bc[8], sv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) bc[j] = bn[j];
      if (p + ATC_TS < np) atc_bias8(brow, k0 + 8 * (p + ATC_TS), l, bn);
      tc::tmem_ld8(trow + (uint32_t)(8 * p), sv);
#pragma unroll
      for (int j = 0; j < 8; ++j) { sv[j] = fmaf(bc[j], LOG2E, sv[j]); cm = fmaxf(cm, sv[j]); }
      tc::tmem_st8(trow + (uint32_t)(8 * p), sv);
    }
    red[sh * ATC_ROWS + rt] = cm;
    tc::tmem_wait_st();
    __syncthreads();
#pragma unroll
    for (int t = 0; t < ATC_TS; ++t) cm = fmaxf(cm, red[t * ATC_ROWS + rt]);
    const float mn = fmaxf(mx, cm);
    corr = mx == -INFINITY ? 0.f : atc_exp2(mx - mn);
    mx = mn;
    ssum *= corr;
    // pass 2: probabilities; hi over the scores, lo beside them
#pragma unroll 1
    for (int p = sh; p < np; p += ATC_TS) {
      float sv[8], pl[8];
      tc::tmem_ld8(trow + (uint32_t)(8 * p), sv);
#pragma unroll
      for (int q4 = 0; q4 < 2; ++q4) {
        float ms[4] = {1.f, 1.f, 1.f, 1.f};
        if (drop) attn_drop4(A, row, (k0 + 8 * p) / 4 + q4, inv_keep, ms);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float pr = atc_exp2(sv[4 * q4 + j] - mn);
          ssum += pr;
          tc::split(drop ? pr * ms[j] : pr, sv[4 * q4 + j], pl[4 * q4 + j]);
        }
      }
      tc::tmem_st8(trow + (uint32_t)(8 * p), sv);
      tc::tmem_st8(trow + ATC_PLO + (uint32_t)(8 * p), pl);
    }
    tc::tmem_wait_st();
    tc::fence_before();
    __syncthreads();          // P complete (and share 0 has taken the previous output columns)
    if (tid == 0) {
      tc::fence_after();
      const uint32_t idesc = tc::idesc_tf32(16, 0, 0);
      for (int s = 0; s < np; ++s) {
        const uint64_t bv = tc::desc(vt + (uint32_t)((k0 >> 3) + s) * 512u, 128u, 256u);
        // two accumulators (P_hi V, P_lo V): consecutive MMAs are independent, the dependent chains half as long
        tc::mma_tf32_ts(tmem + ATC_O, tmem + (uint32_t)(8 * s), bv, idesc, s > 0 ? 1u : 0u);
        tc::mma_tf32_ts(tmem + ATC_O + 16u, tmem + ATC_PLO + (uint32_t)(8 * s), bv, idesc, s > 0 ? 1u : 0u);
      }
      if (k0 + ATC_CHUNK < L) issue_scores(k0 + ATC_CHUNK);
      tc::commit(&bars[0]);
    }
  }
  if (warp == 0) tc::mbar_wait(&bars[0], (uint32_t)(chunk_i & 1));
  __syncthreads();
  tc::fence_after();
  if (sh == 0) take_output();
  __syncthreads();
  red[sh * ATC_ROWS + rt] = ssum;
  __syncthreads();
  if (live && sh == 0) {
    float tot = 0.f;
#pragma unroll
    for (int t = 0; t < ATC_TS; ++t) tot += red[t * ATC_ROWS + rt];
    const float inv = 1.0f / tot;
    float4* op = reinterpret_cast<float4*>(A.O + row * ATC_C);
    op[0] = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
    op[1] = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
    A.lse[row] = mx * 0.6931471805599453f + logf(tot);
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, (uint32_t)ATC_COLS);
}

static int g_attn_tc = 1;
void pwa_attn_tc_set(int on) { g_attn_tc = on; }
// windows of at least one 128-row tile, whole 16-key MMA columns, pieces of 8 keys that never straddle a modality block
static bool attn_tc_eligible(int L, int l, int cq, int cv) {
  return g_attn_tc && cq == ATC_C && cv == ATC_C && L >= 128 && L % 16 == 0 && l % 8 == 0;
}

// Backward: same CTA shape.  Phase A: a query row (ATT_TS threads, keys dealt in quads) -> dQ and the bias gradient;
// phase B: a key row (queries dealt in turns) -> dK, dV.  P is recomputed from the saved row log-sum-exp.
// Bias gradient (dbias is laid out [head][tq][tk]; L*L score gradients per window fold onto l*l entries):
//  * l % 4 == 0 (level 2): a thread sums the M modality blocks of its key quad and issues one 16-byte reduction;
//  * small tables (l*l <= 2048, every other level): the CTA accumulates in shared memory over A.wpc consecutive windows
//    and flushes l*l coalesced atomics once -- 8-30x fewer L2 atomics than one per score element.
template <int CQ, int CV>
__global__ void __launch_bounds__(ATT_MAX_THREADS) pwa_attn_bwd_kernel(const __grid_constant__ AttnArgs A) {
  VX_PDL_ENTRY();
  const int bh = blockIdx.z, head = bh % A.heads;
  const int L = A.L, l = A.l;
  const int tid = threadIdx.x, nthr = blockDim.x;
  VX_DYN_SMEM(float, sm);
  const int fq = att_rows_floats(L, CQ), fv = att_rows_floats(L, CV);
  float* Ks = sm;                 // [L][CQ] (padded, att_row)
  float* Vs = Ks + fq;            // [L][CV]
  float* Qs = Vs + fv;            // [L][CQ]  (pre-scaled)
  float* dOs = Qs + fq;           // [L][CV]
  float* lses = dOs + fv;         // [L]
  float* Dv = lses + L;           // [L]
  float* sdb = A.smem_bias ? Dv + L : nullptr;    // [L][L]: score gradients of the current window (small-table path)
  const bool drop = A.drop_p > 0.f;
  const float inv_keep = drop ? 1.0f / (1.0f - A.drop_p) : 1.f;
  const int rows = nthr / ATT_TS;
  const int r_raw = blockIdx.x * rows + tid / ATT_TS, t = tid % ATT_TS;
  const bool live = r_raw < L;
  const int r = live ? r_raw : L - 1;
  const float* bTh = A.biasT + (size_t)head * l * l;
  const int nk4 = (L + 3) >> 2;
  const bool vec = (l & 3) == 0 && !sdb;
  const int tq = r % l;
  float* dbrow = A.dbiasT + ((size_t)head * l + tq) * l;       // [head][tq][tk]
  constexpr int ATT_NE = 8;                       // bias-table entries owned by a thread on the small-table path
  float accE[ATT_NE];
#pragma unroll
  for (int n = 0; n < ATT_NE; ++n) accE[n] = 0.f;

  const int N1 = min(A.Ns, (int)(blockIdx.y + 1) * A.wpc);
  for (int N = blockIdx.y * A.wpc; N < N1; ++N) {
    const size_t wbase = (size_t)bh * A.Ns + N;
    __syncthreads();
    att_stage(Ks, A.K + wbase * L * CQ, L, CQ, tid, nthr);
    att_stage(Qs, A.Q + wbase * L * CQ, L, CQ, tid, nthr);
    att_stage(Vs, A.V + wbase * L * CV, L, CV, tid, nthr);
    att_stage(dOs, A.dO + wbase * L * CV, L, CV, tid, nthr);
    att_stage_flat(lses, A.lse + wbase * L, L, tid, nthr);
    vx_cp_async_commit();
    for (int i = tid; i < L; i += nthr) {
      float d = 0.f;
      for (int c = 0; c < CV; ++c) d = fmaf(__ldg(A.dO + (wbase * L + i) * CV + c), __ldg(A.O + (wbase * L + i) * CV + c), d);
      Dv[i] = d;
    }
    vx_cp_async_wait_all();
    __syncthreads();
    for (int i = tid; i < fq; i += nthr) Qs[i] *= A.scale;      // (padding included: harmless)
    __syncthreads();

    // phase A: query row r -> dQ_r and the bias gradient of its row
    {
      float q[CQ], dq[CQ], dO[CV];
#pragma unroll
      for (int c = 0; c < CQ; ++c) { q[c] = Qs[att_row(r, CQ) + c]; dq[c] = 0.f; }
#pragma unroll
      for (int c = 0; c < CV; ++c) dO[c] = dOs[att_row(r, CV) + c];
      const float lse = lses[r], Dr = Dv[r];
      const size_t row = wbase * L + r;
      // score gradient of key k (and its contribution to dq)
      auto key = [&](int k, int tk, float mscale) -> float {
        const float* kr = Ks + att_row(k, CQ);
        const float* vr = Vs + att_row(k, CV);
        float s = __ldg(bTh + (size_t)tk * l + tq);
#pragma unroll
        for (int c = 0; c < CQ; ++c) s = fmaf(q[c], kr[c], s);
        const float p = att_exp(s - lse);
        float dp = 0.f;
#pragma unroll
        for (int c = 0; c < CV; ++c) dp = fmaf(dO[c], vr[c], dp);
        const float ds = p * (dp * mscale - Dr);
#pragma unroll
        for (int c = 0; c < CQ; ++c) dq[c] = fmaf(ds, kr[c], dq[c]);
        return ds;
      };
      if (vec) {
        const int lq = l >> 2, M = L / l;
        for (int qd = t; qd < lq; qd += ATT_TS) {
          float dsum[4] = {0.f, 0.f, 0.f, 0.f};
          for (int j = 0; j < M; ++j) {
            const int k0 = j * l + qd * 4;
            float ms[4] = {1.f, 1.f, 1.f, 1.f};
            if (drop) attn_drop4(A, row, k0 >> 2, inv_keep, ms);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) dsum[kk] += key(k0 + kk, qd * 4 + kk, ms[kk]);
          }
          if (live) att_red4(dbrow + qd * 4, dsum);
        }
      } else {
        for (int qd = t; qd < nk4; qd += ATT_TS) {
          const int k0 = qd * 4;
          const int tk0 = k0 % l;
          float ms[4] = {1.f, 1.f, 1.f, 1.f};
          if (drop) attn_drop4(A, row, qd, inv_keep, ms);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const int k = k0 + kk;
            if (k >= L) break;
            int tk = tk0 + kk;
            while (tk >= l) tk -= l;
            const float ds = key(k, tk, ms[kk]);
            if (live) { if (sdb) sdb[r * L + k] = ds; else atomicAdd(dbrow + tk, ds); }
          }
        }
      }
#pragma unroll
      for (int o = 1; o < ATT_TS; o <<= 1)
#pragma unroll
        for (int c = 0; c < CQ; ++c) dq[c] += __shfl_xor_sync(0xffffffffu, dq[c], o);
      if (live && t == 0) {
#pragma unroll
        for (int c = 0; c < CQ; ++c) A.dQ[row * CQ + c] = dq[c] * A.scale;
      }
    }
    if (sdb) {      // fold the window's L x L score gradients onto the l x l entries this thread owns (no atomics)
      __syncthreads();
      const int M = L / l;
#pragma unroll
      for (int n = 0; n < ATT_NE; ++n) {
        const int e = tid + n * nthr;
        if (e < l * l) {
          const int eq = e / l, ek = e - eq * l;
          float a = 0.f;
          for (int i = 0; i < M; ++i)
            for (int j = 0; j < M; ++j) a += sdb[(i * l + eq) * L + j * l + ek];
          accE[n] += a;
        }
      }
    }
    // phase B: key row r -> dK_r, dV_r
    {
      float kx[CQ], dk[CQ], v[CV], dv[CV];
#pragma unroll
      for (int c = 0; c < CQ; ++c) { kx[c] = Ks[att_row(r, CQ) + c]; dk[c] = 0.f; }
#pragma unroll
      for (int c = 0; c < CV; ++c) { v[c] = Vs[att_row(r, CV) + c]; dv[c] = 0.f; }
      const int tk = r % l;
      const float* bN = A.biasN + (size_t)head * l * l + tk;       // [tq][tk]: the lanes of a warp are consecutive keys
      int iq = t % l;                                              // i % l, kept incrementally
      // Dropout: one hashed word per (query i, key r) element -- cheaper than generating quads and exchanging them by shuffle
      const uint32_t dkey = drop ? rng_key(A.seed + (A.seed_dev ? (uint64_t)__ldg(A.seed_dev) : 0), ATT_SITE) : 0u;
#pragma unroll 1
      for (int i0 = t; i0 < L; i0 += 4 * ATT_TS) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + ATT_TS * u;
          const float msk = (drop && i < L) ? attn_drop1(A, dkey, wbase * L + i, r_raw, inv_keep) : 1.f;
          if (i < L) {
            const float* qr = Qs + att_row(i, CQ);
            const float* gr = dOs + att_row(i, CV);
            float s = __ldg(bN + (size_t)iq * l);
#pragma unroll
            for (int c = 0; c < CQ; ++c) s = fmaf(qr[c], kx[c], s);
            const float p = att_exp(s - lses[i]);
            float dp = 0.f;
#pragma unroll
            for (int c = 0; c < CV; ++c) dp = fmaf(gr[c], v[c], dp);
            const float pm = p * msk;
#pragma unroll
            for (int c = 0; c < CV; ++c) dv[c] = fmaf(pm, gr[c], dv[c]);
            const float ds = p * (dp * msk - Dv[i]);
#pragma unroll
            for (int c = 0; c < CQ; ++c) dk[c] = fmaf(ds, qr[c], dk[c]);   // Qs carries the 1/sqrt(c) scale
          }
          iq += ATT_TS;
          while (iq >= l) iq -= l;
        }
      }
#pragma unroll
      for (int o = 1; o < ATT_TS; o <<= 1) {
#pragma unroll
        for (int c = 0; c < CQ; ++c) dk[c] += __shfl_xor_sync(0xffffffffu, dk[c], o);
#pragma unroll
        for (int c = 0; c < CV; ++c) dv[c] += __shfl_xor_sync(0xffffffffu, dv[c], o);
      }
      if (live && t == 0) {
        const size_t row = wbase * L + r;
#pragma unroll
        for (int c = 0; c < CQ; ++c) A.dK[row * CQ + c] = dk[c];
#pragma unroll
        for (int c = 0; c < CV; ++c) A.dV[row * CV + c] = dv[c];
      }
    }
  }
  if (sdb) {
    float* gdb = A.dbiasT + (size_t)head * l * l;
#pragma unroll
    for (int n = 0; n < ATT_NE; ++n) {
      const int e = tid + n * nthr;
      if (e < l * l) atomicAdd(gdb + e, accE[n]);
    }
  }
}

template <int CQ, int CV>
static int launch_attn(const AttnArgs& A, bool bwd, cudaStream_t st) {
  // rows per CTA: up to 128 (512 threads); short windows get one CTA of ceil32(L * ATT_TS) threads
  int threads = ((A.L * ATT_TS + 31) / 32) * 32;
  if (threads > ATT_MAX_THREADS) threads = ATT_MAX_THREADS;
  const int rows = threads / ATT_TS;
  dim3 grid(cdiv(A.L, rows), A.Ns, A.B * A.heads);
  if (!bwd) {
    const size_t pad = (size_t)((A.L + 3) / 4) * 8;
    const size_t smem = sizeof(float) * ((size_t)A.L * (CQ + CV) + 2 * pad);
    VX_SET_SMEM((pwa_attn_fwd_kernel<CQ, CV>), smem);
    VX_LAUNCH((pwa_attn_fwd_kernel<CQ, CV>), grid, dim3(threads), smem, st, A);
    return check_launch("pwa_attn_fwd_kernel");
  }
  AttnArgs Ab = A;
  // small-table path: the CTA sees whole windows, a thread owns <= 8 table entries, the L x L buffer stays <= 64 KB
  Ab.smem_bias = (A.l * A.l <= 2048 && grid.x == 1 && A.l * A.l <= 8 * threads && (size_t)A.L * A.L <= 16384) ? 1 : 0;
  Ab.wpc = 1;
  if (Ab.smem_bias) {
    const long long ctas = (long long)A.Ns * A.B * A.heads;
    Ab.wpc = (int)((ctas + 2 * kSMs - 1) / (2 * kSMs));      // ceil: at most one wave of 2 CTAs per SM
    if (Ab.wpc < 1) Ab.wpc = 1;
    if (Ab.wpc > 16) Ab.wpc = 16;
    grid.y = cdiv(A.Ns, Ab.wpc);
  }
  const size_t smem = sizeof(float) * ((size_t)A.L * (2 * CQ + 2 * CV + 2) + 4 * (size_t)((A.L + 3) / 4) * 8 +
                                       (Ab.smem_bias ? (size_t)A.L * A.L : 0));
  VX_SET_SMEM((pwa_attn_bwd_kernel<CQ, CV>), smem);
  VX_LAUNCH((pwa_attn_bwd_kernel<CQ, CV>), grid, dim3(threads), smem, st, Ab);
  return check_launch("pwa_attn_bwd_kernel");
}

static int dispatch_attn(const AttnArgs& A, int cq, int cv, bool bwd, cudaStream_t st) {
  if (!bwd && A.biasN && attn_tc_eligible(A.L, A.l, cq, cv)) {
    const size_t smem = sizeof(float) * (2 * ATC_ROWS * ATC_C + ATC_TS * ATC_ROWS + (size_t)A.L * ATC_C * 2 + (size_t)A.L * 16);
    if (smem <= 100 * 1024) {
      VX_SET_SMEM(pwa_attn_fwd_tc_kernel, smem);
      prof_flops(4.0 * (double)A.B * A.heads * A.Ns * (double)A.L * A.L * ATC_C);
      VX_LAUNCH(pwa_attn_fwd_tc_kernel, dim3(cdiv(A.L, ATC_ROWS), A.Ns, A.B * A.heads), dim3(ATC_THREADS), smem, st, A);
      return check_launch("pwa_attn_fwd_tc_kernel");
    }
  }
#define VX_ATT(a, b) if (cq == a && cv == b) return launch_attn<a, b>(A, bwd, st)
  VX_ATT(4, 4); VX_ATT(4, 8); VX_ATT(8, 8); VX_ATT(8, 16); VX_ATT(16, 16); VX_ATT(16, 32);
#undef VX_ATT
  set_error("pwa: per-head dims (%d, %d) not instantiated", cq, cv);
  return VX_ERR_UNSUPPORTED;
}

// ---------------------------------------------------------------------------------------------------
// scatter: tokens -> NCDHW with per-window trilinear up-sampling (align_corners=True), PWA.py:177-200
// ---------------------------------------------------------------------------------------------------
VX_DEV void lerp_coef(int p, int n, int out, int& i0, int& i1, float& w1) {
  // F.interpolate(align_corners=True): src = p * (n-1)/(out-1)
  if (out <= 1 || n <= 1) { i0 = 0; i1 = 0; w1 = 0.f; return; }
  const float scale = (float)(n - 1) / (float)(out - 1);
  const float src = scale * (float)p;
  i0 = (int)src;
  if (i0 > n - 1) i0 = n - 1;
  i1 = i0 < n - 1 ? i0 + 1 : i0;
  w1 = src - (float)i0;
}

struct ScatterArgs {
  const float* tok;               // (B, h, Ns, L, cper)
  float* dst[VX_MAX_MODAL];       // (B, Ct, S)
  int Ct, cper;
};

// CTA = one (b, channel) plane x a slab of voxels; blockIdx.z = modality.  The scale j, head and per-head channel are uniform
// per CTA, and everything that depends on one coordinate only -- window index, token index inside the window, the two
// interpolation taps and their weight (align_corners=True inside each big window) -- is tabulated per axis in shared memory
// once per CTA: a voxel then costs two divisions (flat index -> z, y, x) instead of ~25 (profiles/r2n_step_stalls.txt:
// pwa_scatter_kernel was issue-bound at 72 % issue-active).
constexpr int SC_MAX_AXIS = 3 * 160;       // D + H + W table entries (the launcher falls back to one entry per voxel axis beyond)
__global__ void __launch_bounds__(256) pwa_scatter_kernel(const __grid_constant__ PwaGeo G, const __grid_constant__ ScatterArgs A) {
  VX_PDL_ENTRY();
  const int m = blockIdx.z;
  const int cper = A.cper, Ct = A.Ct;
  const int b = blockIdx.y / Ct, ch = blockIdx.y % Ct;
  const int c = ch % cper, head = (ch / cper) % G.heads, j = ch / (cper * G.heads);
  __shared__ int t_w[SC_MAX_AXIS], t_i0[SC_MAX_AXIS], t_i1[SC_MAX_AXIS];
  __shared__ float t_w1[SC_MAX_AXIS];
  const int ext[3] = {G.D, G.H, G.W};
  const int aoff[3] = {0, G.D, G.D + G.H};
  const bool nearest = G.vol[j] == 1;
  for (int i = threadIdx.x; i < G.D + G.H + G.W; i += blockDim.x) {
    const int ax = i < G.D ? 0 : (i < G.D + G.H ? 1 : 2);
    const int p = i - aoff[ax];
    const int big = G.big[j][ax];
    const int w = p / big, r = p - w * big;
    t_w[i] = w;
    if (nearest) { t_i0[i] = r; t_i1[i] = r; t_w1[i] = 0.f; }
    else { int i0, i1; float w1; lerp_coef(r, G.n[ax], big, i0, i1, w1); t_i0[i] = i0; t_i1[i] = i1; t_w1[i] = w1; }
  }
  (void)ext;
  __syncthreads();
  float* dst = A.dst[m] + ((size_t)b * Ct + ch) * G.S;
  const float* tok0 = A.tok + (((size_t)b * G.heads + head) * G.Ns + G.Noff[j]) * (size_t)G.L * cper + (size_t)m * G.l * cper + c;
  const int n1 = G.n[1], n2 = G.n[2], Nw1 = G.Nw[j][1], Nw2 = G.Nw[j][2];
  const size_t wstride = (size_t)G.L * cper;                     // tokens of one window
  const unsigned HW = (unsigned)(G.H * G.W);
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < (unsigned)G.S; idx += gridDim.x * blockDim.x) {
    const unsigned z = idx / HW, rem = idx - z * HW;
    const unsigned y = rem / (unsigned)G.W, x = rem - y * (unsigned)G.W;
    const int iz = (int)z, iy = G.D + (int)y, ix = G.D + G.H + (int)x;
    const int Nloc = (t_w[iz] * Nw1 + t_w[iy]) * Nw2 + t_w[ix];
    const float* tp = tok0 + (size_t)Nloc * wstride;
    float val;
    if (nearest) {
      const int t = (t_i0[iz] * n1 + t_i0[iy]) * n2 + t_i0[ix];
      val = __ldg(tp + (size_t)t * cper);
    } else {
      const int a0 = t_i0[iz], a1 = t_i1[iz], b0 = t_i0[iy], b1 = t_i1[iy], c0 = t_i0[ix], c1 = t_i1[ix];
      const float wa = t_w1[iz], wb = t_w1[iy], wc = t_w1[ix];
#define VX_T(a, bq, cq_) __ldg(tp + (size_t)(((a) * n1 + (bq)) * n2 + (cq_)) * cper)
      const float v000 = VX_T(a0, b0, c0), v001 = VX_T(a0, b0, c1), v010 = VX_T(a0, b1, c0), v011 = VX_T(a0, b1, c1);
      const float v100 = VX_T(a1, b0, c0), v101 = VX_T(a1, b0, c1), v110 = VX_T(a1, b1, c0), v111 = VX_T(a1, b1, c1);
#undef VX_T
      const float ua = 1.f - wa, ub = 1.f - wb, uc = 1.f - wc;
      val = ua * (ub * (uc * v000 + wc * v001) + wb * (uc * v010 + wc * v011)) +
            wa * (ub * (uc * v100 + wc * v101) + wb * (uc * v110 + wc * v111));
    }
    dst[idx] = val;
  }
}

// adjoint: dtok[token] = sum over the window's voxels of weight * dA.  Thread per token element for small
// windows, warp per token element (lanes along the contiguous axis) when the big window is >= 8 wide.
struct ScatterBwdArgs {
  const float* src[VX_MAX_MODAL];   // dA (B, Ct, S)
  float* dtok;                      // (B, h, Ns, L, cper)
  int Ct, cper;
};

VX_DEV float lerp_weight(int p, int n, int out, int a) {
  int i0, i1; float w1;
  lerp_coef(p, n, out, i0, i1, w1);
  float w = 0.f;
  if (i0 == a) w += 1.f - w1;
  if (i1 == a) w += w1;
  return w;
}

// One CTA per (channel, batch, modality): the channel's gradient volume is staged in shared memory and reduced along x,
// then y, then z with per-axis weight tables (the trilinear adjoint is separable inside a big window), so every voxel is
// read once and the work per token no longer grows with the window volume.
__global__ void __launch_bounds__(256) pwa_scatter_bwd_kernel(const __grid_constant__ PwaGeo G, const __grid_constant__ ScatterBwdArgs A) {
  VX_PDL_ENTRY();
  const int ch = blockIdx.x, b = blockIdx.y, m = blockIdx.z;
  const int cper = A.cper, Ct = A.Ct;
  const int c = ch % cper, head = (ch / cper) % G.heads, j = ch / (cper * G.heads);
  const int D = G.D, H = G.H, W = G.W, S = G.S;
  const int B0 = G.big[j][0], B1 = G.big[j][1], B2 = G.big[j][2];
  const int n0 = G.n[0], n1 = G.n[1], n2 = G.n[2];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const float* src = A.src[m] + ((size_t)b * Ct + ch) * S;
  float* dtok = A.dtok + ((((size_t)b * G.heads + head) * G.Ns + G.Noff[j]) * G.L + (size_t)m * G.l) * cper + c;
  const size_t wstride = (size_t)G.L * cper;               // between windows of the scale
  if (G.vol[j] == 1) {                                       // identity scale: a permuted copy
    for (int idx = tid; idx < S; idx += nthr) {
      const int x = idx % W, y = (idx / W) % H, z = idx / (W * H);
      const int Nloc = ((z / B0) * G.Nw[j][1] + y / B1) * G.Nw[j][2] + x / B2;
      const int t = ((z % B0) * n1 + (y % B1)) * n2 + (x % B2);
      dtok[(size_t)Nloc * wstride + (size_t)t * cper] = __ldg(src + idx);
    }
    return;
  }
  const int W2 = G.Nw[j][2] * n2, H1 = G.Nw[j][1] * n1, D0 = G.Nw[j][0] * n0;   // token-grid extents of the volume
  VX_DYN_SMEM(float, sm);
  float* vin = sm;                           // [D][H][W]
  float* t1 = vin + S;                       // [D][H][W2]
  float* t2 = t1 + (size_t)D * H * W2;       // [D][H1][W2]
  float* wX = t2 + (size_t)D * H1 * W2;      // [B2][n2]
  float* wY = wX + B2 * n2;                  // [B1][n1]
  float* wZ = wY + B1 * n1;                  // [B0][n0]
  for (int i = tid; i < S; i += nthr) vx_cp_async4(vin + i, src + i, true);
  vx_cp_async_commit();
  for (int i = tid; i < B2 * n2; i += nthr) wX[i] = lerp_weight(i / n2, n2, B2, i % n2);
  for (int i = tid; i < B1 * n1; i += nthr) wY[i] = lerp_weight(i / n1, n1, B1, i % n1);
  for (int i = tid; i < B0 * n0; i += nthr) wZ[i] = lerp_weight(i / n0, n0, B0, i % n0);
  vx_cp_async_wait_all();
  __syncthreads();
  for (int o = tid; o < D * H * W2; o += nthr) {
    const int xo = o % W2, zy = o / W2;
    const int wx = xo / n2, cc = xo % n2;
    const float* p = vin + (size_t)zy * W + wx * B2;
    float acc = 0.f;
    for (int px = 0; px < B2; ++px) acc = fmaf(wX[px * n2 + cc], p[px], acc);
    t1[o] = acc;
  }
  __syncthreads();
  for (int o = tid; o < D * H1 * W2; o += nthr) {
    const int xo = o % W2, yo = (o / W2) % H1, z = o / (W2 * H1);
    const int wy = yo / n1, bb = yo % n1;
    const float* p = t1 + ((size_t)z * H + wy * B1) * W2 + xo;
    float acc = 0.f;
    for (int py = 0; py < B1; ++py) acc = fmaf(wY[py * n1 + bb], p[(size_t)py * W2], acc);
    t2[o] = acc;
  }
  __syncthreads();
  for (int o = tid; o < D0 * H1 * W2; o += nthr) {
    const int xo = o % W2, yo = (o / W2) % H1, zo = o / (W2 * H1);
    const int wz = zo / n0, a = zo % n0;
    const float* p = t2 + ((size_t)(wz * B0) * H1 + yo) * W2 + xo;
    float acc = 0.f;
    for (int pz = 0; pz < B0; ++pz) acc = fmaf(wZ[pz * n0 + a], p[(size_t)pz * H1 * W2], acc);
    const int Nloc = (wz * G.Nw[j][1] + yo / n1) * G.Nw[j][2] + xo / n2;
    const int t = (a * n1 + yo % n1) * n2 + xo % n2;
    dtok[(size_t)Nloc * wstride + (size_t)t * cper] = acc;
  }
}

// ---------------------------------------------------------------------------------------------------
// buffer layouts
// ---------------------------------------------------------------------------------------------------
enum { SV_XHAT1 = 0, SV_RSTD1, SV_A, SV_Y, SV_XHAT2, SV_RSTD2, SV_HPRE, SV_QT, SV_KT, SV_VT, SV_ARGQ, SV_ARGK, SV_ARGV,
       SV_LSE, SV_OT, SV_COUNT };

struct PwaLayout {
  PwaGeo G;
  int C, cqk, cv, cq_h, cv_h, eC;
  size_t saved[SV_COUNT];
  size_t tokq, tokv;                       // element counts
  // workspace offsets (bytes)
  size_t off_qkv, off_biasT, off_biasN, off_dh, off_dln2, off_dy, off_dA, off_dOt, off_dQt, off_dKt, off_dVt, off_dqkv, off_dln1,
      off_dbiasT, total;
};

static int pwa_layout(const vx_pwa_desc* d, PwaLayout& P) {
  VX_TRY(make_geo(d, P.G));
  const PwaGeo& G = P.G;
  P.C = d->C; P.cqk = d->c_qk; P.cv = d->c_v; P.eC = d->ffn_expansion * d->C;
  P.cq_h = d->c_qk / (G.nb * G.heads); P.cv_h = d->c_v / (G.nb * G.heads);
  if (d->ffn_expansion <= 0) { set_error("pwa: bad ffn expansion"); return VX_ERR_BAD_DESC; }
  const size_t MB = (size_t)G.M * G.B, S = G.S;
  P.tokq = (size_t)G.B * G.heads * G.Ns * G.L * P.cq_h;
  P.tokv = (size_t)G.B * G.heads * G.Ns * G.L * P.cv_h;
  const size_t rows = (size_t)G.B * G.heads * G.Ns * G.L;
  P.saved[SV_XHAT1] = 4 * MB * P.C * S;  P.saved[SV_RSTD1] = 4 * MB * S;
  P.saved[SV_A] = 4 * MB * P.cv * S;     P.saved[SV_Y] = 4 * MB * P.C * S;
  P.saved[SV_XHAT2] = 4 * MB * P.C * S;  P.saved[SV_RSTD2] = 4 * MB * S;
  P.saved[SV_HPRE] = 4 * MB * P.eC * S;
  P.saved[SV_QT] = 4 * P.tokq; P.saved[SV_KT] = 4 * P.tokq; P.saved[SV_VT] = 4 * P.tokv;
  P.saved[SV_ARGQ] = 4 * P.tokq; P.saved[SV_ARGK] = 4 * P.tokq; P.saved[SV_ARGV] = 4 * P.tokv;
  P.saved[SV_LSE] = 4 * rows; P.saved[SV_OT] = 4 * P.tokv;
  size_t off = 0;
  const size_t cqkv = (size_t)2 * P.cqk + P.cv;
  P.off_qkv = off;    off += align256(4 * MB * cqkv * S);
  P.off_biasT = off;  off += align256(4 * (size_t)G.heads * G.l * G.l);
  P.off_biasN = off;  off += align256(4 * (size_t)G.heads * G.l * G.l);
  P.off_dh = off;     off += align256(4 * MB * P.eC * S);
  P.off_dln2 = off;   off += align256(4 * MB * P.C * S);
  P.off_dy = off;     off += align256(4 * MB * P.C * S);
  P.off_dA = off;     off += align256(4 * MB * P.cv * S);
  P.off_dOt = off;    off += align256(4 * P.tokv);
  P.off_dQt = off;    off += align256(4 * P.tokq);
  P.off_dKt = off;    off += align256(4 * P.tokq);
  P.off_dVt = off;    off += align256(4 * P.tokv);
  P.off_dqkv = off;   off += align256(4 * MB * cqkv * S);
  P.off_dln1 = off;   off += align256(4 * MB * P.C * S);
  P.off_dbiasT = off; off += align256(4 * (size_t)G.heads * G.l * G.l);
  P.total = off;
  return VX_OK;
}

enum { PP_LN1W = 0, PP_LN1B, PP_WQ, PP_BQ, PP_WK, PP_BK, PP_WV, PP_BV, PP_WMIX, PP_BMIX, PP_LN2W, PP_LN2B, PP_W1, PP_B1,
       PP_W2, PP_B2, PP_COUNT };
constexpr uint32_t SITE_MIX = 3, SITE_FFN1 = 4, SITE_FFN2 = 5;

}  // namespace vx

using namespace vx;

extern "C" int vx_pwa_saved_layout(const vx_pwa_desc* d, vx_pwa_saved* layout) {
  PwaLayout P;
  VX_TRY(pwa_layout(d, P));
  layout->n_saved = SV_COUNT;
  for (int i = 0; i < SV_COUNT; ++i) layout->saved_bytes[i] = P.saved[i];
  return VX_OK;
}

extern "C" size_t vx_pwa_workspace(const vx_pwa_desc* d) {
  PwaLayout P;
  if (pwa_layout(d, P) != VX_OK) return 0;
  return P.total;
}

extern "C" int vx_pwa_gather(const vx_pwa_desc* d, int32_t channels_total, const void* x, void* tokens, void* argmax,
                             vx_stream_t stream) {
  PwaGeo G;
  VX_TRY(make_geo(d, G));
  if (channels_total % (G.nb * G.heads)) { set_error("pwa_gather: channels do not split"); return VX_ERR_BAD_DESC; }
  G.M = 1; G.L = G.l;
  GatherArgs A{};
  A.nkind = 1; A.src[0][0] = (const float*)x; A.tok[0] = (float*)tokens; A.arg[0] = (int*)argmax;
  A.Ct[0] = channels_total; A.cper[0] = channels_total / (G.nb * G.heads);
  return launch_gather(G, A, (cudaStream_t)stream);
}

extern "C" int vx_pwa_block_fwd(const vx_pwa_desc* d, const void* const* in, void* const* out, void* workspace,
                                size_t workspace_bytes, vx_stream_t stream) {
  PwaLayout P;
  VX_TRY(pwa_layout(d, P));
  prof_scope("pwa_fwd B%d M%d C%d S%d L%d", d->B, d->M, d->C, P.G.S, P.G.L);
  set_seed_dev(d->seed_offset);
  if (!workspace || workspace_bytes < P.total) { set_error("pwa_fwd: workspace %zu < %zu", workspace_bytes, P.total); return VX_ERR_WORKSPACE; }
  const PwaGeo& G = P.G;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  const int M = G.M, B = G.B, S = G.S, C = P.C;
  const size_t BS = (size_t)B * S;
  auto X = [&](int m) { return (const float*)in[m]; };
  auto PRM = [&](int m, int k) { return (const float*)in[M + m * PP_COUNT + k]; };
  const float* table = (const float*)in[M + M * PP_COUNT];
  const long long* index = (const long long*)in[M + M * PP_COUNT + 1];
  auto Z = [&](int m) { return (float*)out[m]; };
  auto SV = [&](int k) { return (float*)out[M + k]; };
  const bool train = d->training != 0;
  const float attn_p = train ? d->attn_drop : 0.f, proj_p = train ? d->proj_drop : 0.f;

  float* qkv = (float*)(ws + P.off_qkv);
  float* biasT = (float*)(ws + P.off_biasT);
  const size_t cqkv = (size_t)2 * P.cqk + P.cv;
  auto Qf = [&](int m) { return qkv + (size_t)m * cqkv * BS; };
  auto Kf = [&](int m) { return Qf(m) + (size_t)P.cqk * BS; };
  auto Vf = [&](int m) { return Qf(m) + (size_t)2 * P.cqk * BS; };

  // LN1
  LnBatch L1{}; L1.n = M; L1.B = B; L1.C = C; L1.S = S; L1.eps = d->ln_eps;
  for (int m = 0; m < M; ++m) { L1.x[m] = X(m); L1.xhat[m] = SV(SV_XHAT1) + (size_t)m * C * BS; L1.rstd[m] = SV(SV_RSTD1) + (size_t)m * BS; }
  VX_TRY(ln_forward(L1, st));
  // q, k, v
  {
    PwBatch pb{}; pb.nprob = M; pb.B = B; pb.S = S;
    for (int m = 0; m < M; ++m) {
      PwProblem& p = pb.p[m];
      p.src[0] = PwSrc{L1.xhat[m], C}; p.nsrc = 1; p.Ci = C;
      p.seg[0] = PwSeg{PRM(m, PP_WQ), PRM(m, PP_BQ), C, P.cqk, Qf(m)};
      p.seg[1] = PwSeg{PRM(m, PP_WK), PRM(m, PP_BK), C, P.cqk, Kf(m)};
      p.seg[2] = PwSeg{PRM(m, PP_WV), PRM(m, PP_BV), C, P.cv, Vf(m)};
      p.nseg = 3; p.Co = (int)cqkv;
      p.pro = PRO_AFFINE; p.pro_a = PRM(m, PP_LN1W); p.pro_c = PRM(m, PP_LN1B); p.pro_bstride = 0;
    }
    VX_TRY(pw_forward(pb, st));
  }
  // tokens
  {
    GatherArgs A{}; A.nkind = 3;
    for (int m = 0; m < M; ++m) { A.src[0][m] = Qf(m); A.src[1][m] = Kf(m); A.src[2][m] = Vf(m); }
    A.tok[0] = SV(SV_QT); A.tok[1] = SV(SV_KT); A.tok[2] = SV(SV_VT);
    A.arg[0] = (int*)SV(SV_ARGQ); A.arg[1] = (int*)SV(SV_ARGK); A.arg[2] = (int*)SV(SV_ARGV);
    A.Ct[0] = A.Ct[1] = P.cqk; A.Ct[2] = P.cv; A.cper[0] = A.cper[1] = P.cq_h; A.cper[2] = P.cv_h;
    VX_TRY(launch_gather(G, A, st));
  }
  // attention
  {
    const int nb = G.heads * G.l * G.l;
    // the query-major copy feeds the tensor-core kernel (a thread = query row reads 8 consecutive keys with two 16-byte loads)
    float* biasN = attn_tc_eligible(G.L, G.l, P.cq_h, P.cv_h) ? (float*)(ws + P.off_biasN) : nullptr;
    VX_LAUNCH(pwa_bias_kernel, dim3(cdiv(nb, 256)), dim3(256), 0, st, table, index, biasT, biasN, G.heads, G.l);
    VX_TRY(check_launch("pwa_bias_kernel"));
    AttnArgs A{};
    A.Q = SV(SV_QT); A.K = SV(SV_KT); A.V = SV(SV_VT); A.biasT = biasT; A.biasN = biasN; A.O = SV(SV_OT); A.lse = SV(SV_LSE);
    A.B = B; A.heads = G.heads; A.Ns = G.Ns; A.L = G.L; A.l = G.l;
    A.scale = 1.0f / sqrtf((float)P.cq_h); A.drop_p = attn_p; A.seed = d->seed; A.seed_dev = get_seed_dev();
    VX_TRY(dispatch_attn(A, P.cq_h, P.cv_h, false, st));
  }
  // scatter
  {
    ScatterArgs A{}; A.tok = SV(SV_OT); A.Ct = P.cv; A.cper = P.cv_h;
    for (int m = 0; m < M; ++m) A.dst[m] = SV(SV_A) + (size_t)m * P.cv * BS;
    if (G.D + G.H + G.W > SC_MAX_AXIS || (long long)B * P.cv > 65535) { set_error("pwa scatter: extent / plane count beyond the kernel's tables"); return VX_ERR_UNSUPPORTED; }
    int blocks = cdiv(S, 256 * 2);                     // plane slabs: ~2 voxels per thread, at least ~2 CTAs per SM overall
    while (blocks > 1 && (long long)blocks * B * P.cv * M > 8LL * kSMs) blocks = (blocks + 1) / 2;
    VX_LAUNCH(pwa_scatter_kernel, dim3(blocks, B * P.cv, M), dim3(256), 0, st, G, A);
    VX_TRY(check_launch("pwa_scatter_kernel"));
  }
  // y = 2x + Drop(Wmix a + b)
  {
    PwBatch pb{}; pb.nprob = M; pb.B = B; pb.S = S;
    for (int m = 0; m < M; ++m) {
      PwProblem& p = pb.p[m];
      p.src[0] = PwSrc{SV(SV_A) + (size_t)m * P.cv * BS, P.cv}; p.nsrc = 1; p.Ci = P.cv;
      p.seg[0] = PwSeg{PRM(m, PP_WMIX), PRM(m, PP_BMIX), P.cv, C, SV(SV_Y) + (size_t)m * C * BS}; p.nseg = 1; p.Co = C;
      if (proj_p > 0.f) { p.drop_p = proj_p; p.seed = d->seed + 0x1000 * (m + 1); p.site = SITE_MIX; }
      p.res = X(m); p.res_scale = 2.f;
    }
    VX_TRY(pw_forward(pb, st));
  }
  // LN2, FFN
  LnBatch L2{}; L2.n = M; L2.B = B; L2.C = C; L2.S = S; L2.eps = d->ln_eps;
  for (int m = 0; m < M; ++m) { L2.x[m] = SV(SV_Y) + (size_t)m * C * BS; L2.xhat[m] = SV(SV_XHAT2) + (size_t)m * C * BS; L2.rstd[m] = SV(SV_RSTD2) + (size_t)m * BS; }
  VX_TRY(ln_forward(L2, st));
  {
    // small levels: W1 -> GELU -> dropout -> W2 -> dropout -> residual in one launch
    FfnBatch fb{}; fb.nprob = M; fb.B = B; fb.S = S;
    for (int m = 0; m < M; ++m) {
      FfnProblem& f = fb.p[m];
      f.x = L2.xhat[m]; f.C = C; f.pro_a = PRM(m, PP_LN2W); f.pro_c = PRM(m, PP_LN2B); f.pro_bstride = 0;
      f.W1 = PRM(m, PP_W1); f.b1 = PRM(m, PP_B1); f.eC = P.eC; f.hpre = SV(SV_HPRE) + (size_t)m * P.eC * BS;
      f.W2 = PRM(m, PP_W2); f.b2 = PRM(m, PP_B2);
      if (proj_p > 0.f) {
        f.mid_drop_p = proj_p; f.mid_seed = d->seed + 0x1000 * (m + 1); f.mid_site = SITE_FFN1;
        f.drop_p = proj_p; f.seed = d->seed + 0x1000 * (m + 1); f.site = SITE_FFN2;
      }
      f.res = SV(SV_Y) + (size_t)m * C * BS; f.res_scale = 1.f; f.y = Z(m);
    }
    int rc = pw_ffn_small(fb, st);
    if (rc == 1) rc = pw_ffn_tc(fb, st);
    if (rc < 0) return rc;
    if (rc == 0) return VX_OK;
  }
  {
    PwBatch pb{}; pb.nprob = M; pb.B = B; pb.S = S;
    for (int m = 0; m < M; ++m) {
      PwProblem& p = pb.p[m];
      p.src[0] = PwSrc{L2.xhat[m], C}; p.nsrc = 1; p.Ci = C;
      p.seg[0] = PwSeg{PRM(m, PP_W1), PRM(m, PP_B1), C, P.eC, SV(SV_HPRE) + (size_t)m * P.eC * BS}; p.nseg = 1; p.Co = P.eC;
      p.pro = PRO_AFFINE; p.pro_a = PRM(m, PP_LN2W); p.pro_c = PRM(m, PP_LN2B); p.pro_bstride = 0;
    }
    VX_TRY(pw_forward(pb, st));
  }
  {
    PwBatch pb{}; pb.nprob = M; pb.B = B; pb.S = S;
    for (int m = 0; m < M; ++m) {
      PwProblem& p = pb.p[m];
      p.src[0] = PwSrc{SV(SV_HPRE) + (size_t)m * P.eC * BS, P.eC}; p.nsrc = 1; p.Ci = P.eC;
      p.seg[0] = PwSeg{PRM(m, PP_W2), PRM(m, PP_B2), P.eC, C, Z(m)}; p.nseg = 1; p.Co = C;
      p.pro = PRO_GELU;
      if (proj_p > 0.f) {
        p.pro = PRO_GELU_DROPOUT; p.pro_drop_p = proj_p; p.pro_seed = d->seed + 0x1000 * (m + 1); p.pro_site = SITE_FFN1;
        p.drop_p = proj_p; p.seed = d->seed + 0x1000 * (m + 1); p.site = SITE_FFN2;
      }
      p.res = SV(SV_Y) + (size_t)m * C * BS; p.res_scale = 1.f;
    }
    VX_TRY(pw_forward(pb, st));
  }
  return VX_OK;
}

extern "C" int vx_pwa_block_bwd(const vx_pwa_desc* d, const void* const* in, void* const* out, void* workspace,
                                size_t workspace_bytes, vx_stream_t stream) {
  PwaLayout P;
  VX_TRY(pwa_layout(d, P));
  prof_scope("pwa_bwd B%d M%d C%d S%d L%d", d->B, d->M, d->C, P.G.S, P.G.L);
  set_seed_dev(d->seed_offset);
  if (!workspace || workspace_bytes < P.total) { set_error("pwa_bwd: workspace %zu < %zu", workspace_bytes, P.total); return VX_ERR_WORKSPACE; }
  const PwaGeo& G = P.G;
  cudaStream_t st = (cudaStream_t)stream;
  SideJoin side_guard(st);
  char* ws = (char*)workspace;
  const int M = G.M, B = G.B, S = G.S, C = P.C;
  const size_t BS = (size_t)B * S;
  auto DZ = [&](int m) { return (const float*)in[m]; };
  auto PRM = [&](int m, int k) { return (const float*)in[2 * M + m * PP_COUNT + k]; };
  const int base = 2 * M + M * PP_COUNT;
  const float* table = (const float*)in[base];
  const long long* index = (const long long*)in[base + 1];
  auto SV = [&](int k) { return (const float*)in[base + 2 + k]; };
  auto DX = [&](int m) { return (float*)out[m]; };
  auto DP = [&](int m, int k) { return (float*)out[M + m * PP_COUNT + k]; };
  float* dtable = (float*)out[M + M * PP_COUNT];
  const bool train = d->training != 0;
  const float attn_p = train ? d->attn_drop : 0.f, proj_p = train ? d->proj_drop : 0.f;
  const size_t cqkv = (size_t)2 * P.cqk + P.cv;

  float* biasT = (float*)(ws + P.off_biasT);
  float* dh = (float*)(ws + P.off_dh);
  float* dln2 = (float*)(ws + P.off_dln2);
  float* dy = (float*)(ws + P.off_dy);
  float* dA = (float*)(ws + P.off_dA);
  float* dOt = (float*)(ws + P.off_dOt);
  float* dQt = (float*)(ws + P.off_dQt);
  float* dKt = (float*)(ws + P.off_dKt);
  float* dVt = (float*)(ws + P.off_dVt);
  float* dqkv = (float*)(ws + P.off_dqkv);
  float* dln1 = (float*)(ws + P.off_dln1);
  float* dbiasT = (float*)(ws + P.off_dbiasT);

  // zero everything that is accumulated with atomics
  const size_t psz[PP_COUNT] = {(size_t)C, (size_t)C, (size_t)P.cqk * C, (size_t)P.cqk, (size_t)P.cqk * C, (size_t)P.cqk,
                                (size_t)P.cv * C, (size_t)P.cv, (size_t)C * P.cv, (size_t)C, (size_t)C, (size_t)C,
                                (size_t)P.eC * C, (size_t)P.eC, (size_t)C * P.eC, (size_t)C};
  const int nbias = G.heads * G.l * G.l;
  {
    ZeroList zl;
    for (int m = 0; m < M; ++m)
      for (int k = 0; k < PP_COUNT; ++k) zl.add(DP(m, k), psz[k]);
    zl.add(dbiasT, nbias);
    // table rows = prod(2n-1)
    const size_t rows = (size_t)(2 * G.n[0] - 1) * (2 * G.n[1] - 1) * (2 * G.n[2] - 1);
    zl.add(dtable, rows * G.heads);
    VX_TRY(zero_many(zl, side_fork(st)));      // side stream: the main stream catches up with it after its first contraction
  }
  float* biasN = (float*)(ws + P.off_biasN);
  {
    // the dense bias is first needed by the attention backward, a dozen launches down the main stream: built on the side
    // stream, behind the zeroing (the main stream joins both at side_wait below)
    cudaStream_t sb = side_fork(st);
    VX_LAUNCH(pwa_bias_kernel, dim3(cdiv(nbias, 256)), dim3(256), 0, sb, table, index, biasT, biasN, G.heads, G.l);
    VX_TRY(check_launch("pwa_bias_kernel"));
  }

  auto seedm = [&](int m) { return d->seed + 0x1000 * (uint64_t)(m + 1); };

  // ---- FFN backward
  int ffn_fused = 0;
  {
    // dh = (W2^T (dz*m2)) * GELU'(hpre) * m1 and dln2 = W1^T dh in one launch on the small levels (it writes no zeroed buffer)
    FfnBwdBatch fb{}; fb.nprob = M; fb.B = B; fb.S = S;
    for (int m = 0; m < M; ++m) {
      FfnBwdProblem& f = fb.p[m];
      f.dy = DZ(m); f.C = C; f.W2 = PRM(m, PP_W2); f.eC = P.eC; f.hpre = SV(SV_HPRE) + (size_t)m * P.eC * BS;
      f.dh = dh + (size_t)m * P.eC * BS; f.W1 = PRM(m, PP_W1); f.dx = dln2 + (size_t)m * C * BS;
      if (proj_p > 0.f) {
        f.out_drop_p = proj_p; f.out_seed = seedm(m); f.out_site = SITE_FFN2;
        f.mid_drop_p = proj_p; f.mid_seed = seedm(m); f.mid_site = SITE_FFN1;
      }
    }
    int rc = pw_ffn_small_bwd(fb, st);
    if (rc == 1) rc = pw_ffn_tc_bwd(fb, st);
    if (rc != VX_OK && rc != 1) return rc;
    ffn_fused = rc == VX_OK;
  }
  if (!ffn_fused) {
    PwBatch pb{}; pb.nprob = M; pb.B = B; pb.S = S;     // dh = (W2^T (dz*m2)) * GELU'(hpre) * m1
    for (int m = 0; m < M; ++m) {
      PwProblem& p = pb.p[m];
      p.src[0] = PwSrc{DZ(m), C}; p.nsrc = 1; p.Ci = C;
      p.seg[0] = PwSeg{PRM(m, PP_W2), nullptr, P.eC, C, dh + (size_t)m * P.eC * BS}; p.nseg = 1; p.Co = P.eC; p.transposed = 1;
      p.mulgrad = SV(SV_HPRE) + (size_t)m * P.eC * BS;
      if (proj_p > 0.f) {
        p.pro = PRO_DROPOUT; p.pro_drop_p = proj_p; p.pro_seed = seedm(m); p.pro_site = SITE_FFN2;
        p.drop_p = proj_p; p.seed = seedm(m); p.site = SITE_FFN1;
      }
    }
    VX_TRY(pw_forward(pb, st));
  }
  side_wait(st);       // every gradient buffer is zero from here on (nothing above touches one)
  if (!ffn_fused) {
    PwBatch pb{}; pb.nprob = M; pb.B = B; pb.S = S;     // dln2 = W1^T dh
    for (int m = 0; m < M; ++m) {
      PwProblem& p = pb.p[m];
      p.src[0] = PwSrc{dh + (size_t)m * P.eC * BS, P.eC}; p.nsrc = 1; p.Ci = P.eC;
      p.seg[0] = PwSeg{PRM(m, PP_W1), nullptr, C, P.eC, dln2 + (size_t)m * C * BS}; p.nseg = 1; p.Co = C; p.transposed = 1;
    }
    VX_TRY(pw_forward(pb, st));
  }
  {
    WgBatch wb{}; wb.nprob = 2 * M; wb.B = B; wb.S = S;
    for (int m = 0; m < M; ++m) {
      WgProblem& a = wb.p[2 * m];        // dW2 = (dz*m2) (GELU(hpre)*m1)^T
      a.dY = DZ(m); a.Co = C; a.src[0] = PwSrc{SV(SV_HPRE) + (size_t)m * P.eC * BS, P.eC}; a.nsrc = 1; a.Ci = P.eC;
      a.xpro = PRO_GELU;
      if (proj_p > 0.f) {
        a.xpro = PRO_GELU_DROPOUT; a.x_drop_p = proj_p; a.x_seed = seedm(m); a.x_site = SITE_FFN1;
        a.y_drop_p = proj_p; a.y_seed = seedm(m); a.y_site = SITE_FFN2;
      }
      a.dW = DP(m, PP_W2); a.ld = P.eC; a.db = DP(m, PP_B2);
      WgProblem& b1 = wb.p[2 * m + 1];   // dW1 = dh (g2*yhat + b2)^T
      b1.dY = dh + (size_t)m * P.eC * BS; b1.Co = P.eC; b1.src[0] = PwSrc{SV(SV_XHAT2) + (size_t)m * C * BS, C}; b1.nsrc = 1; b1.Ci = C;
      b1.xpro = PRO_AFFINE; b1.xa = PRM(m, PP_LN2W); b1.xc = PRM(m, PP_LN2B); b1.x_bstride = 0;
      b1.dW = DP(m, PP_W1); b1.ld = C; b1.db = DP(m, PP_B1);
    }
    VX_TRY(pw_wgrad(wb, st));
  }
  {
    LnBwdBatch Lb{}; Lb.n = M; Lb.B = B; Lb.C = C; Lb.S = S; Lb.dx_add_scale = 1.f;   // dy = dz + LN2_bwd(dln2)
    for (int m = 0; m < M; ++m) {
      Lb.dout[m] = dln2 + (size_t)m * C * BS; Lb.xhat[m] = SV(SV_XHAT2) + (size_t)m * C * BS; Lb.rstd[m] = SV(SV_RSTD2) + (size_t)m * BS;
      Lb.gamma[m] = PRM(m, PP_LN2W); Lb.dx_add[m] = DZ(m); Lb.dx[m] = dy + (size_t)m * C * BS;
      Lb.dgamma[m] = DP(m, PP_LN2W); Lb.dbeta[m] = DP(m, PP_LN2B);
    }
    VX_TRY(ln_backward(Lb, st));
  }
  // ---- mix backward
  {
    PwBatch pb{}; pb.nprob = M; pb.B = B; pb.S = S;     // dA = Wmix^T (dy*mask)
    for (int m = 0; m < M; ++m) {
      PwProblem& p = pb.p[m];
      p.src[0] = PwSrc{dy + (size_t)m * C * BS, C}; p.nsrc = 1; p.Ci = C;
      p.seg[0] = PwSeg{PRM(m, PP_WMIX), nullptr, P.cv, C, dA + (size_t)m * P.cv * BS}; p.nseg = 1; p.Co = P.cv; p.transposed = 1;
      if (proj_p > 0.f) { p.pro = PRO_DROPOUT; p.pro_drop_p = proj_p; p.pro_seed = seedm(m); p.pro_site = SITE_MIX; }
    }
    VX_TRY(pw_forward(pb, st));
    WgBatch wb{}; wb.nprob = M; wb.B = B; wb.S = S;
    for (int m = 0; m < M; ++m) {
      WgProblem& a = wb.p[m];
      a.dY = dy + (size_t)m * C * BS; a.Co = C; a.src[0] = PwSrc{SV(SV_A) + (size_t)m * P.cv * BS, P.cv}; a.nsrc = 1; a.Ci = P.cv;
      if (proj_p > 0.f) { a.y_drop_p = proj_p; a.y_seed = seedm(m); a.y_site = SITE_MIX; }
      a.dW = DP(m, PP_WMIX); a.ld = P.cv; a.db = DP(m, PP_BMIX);
    }
    VX_TRY(pw_wgrad(wb, st));
  }
  // ---- scatter adjoint
  {
    ScatterBwdArgs A{}; A.dtok = dOt; A.Ct = P.cv; A.cper = P.cv_h;
    for (int m = 0; m < M; ++m) A.src[m] = dA + (size_t)m * P.cv * BS;
    // shared memory: the volume, its x- and (x, y)-reduced copies (largest for the smallest pooled window) and the tables
    size_t fl = 0;
    for (int j = 0; j < G.nb; ++j) {
      if (G.vol[j] == 1) continue;
      const size_t W2 = (size_t)G.Nw[j][2] * G.n[2], H1 = (size_t)G.Nw[j][1] * G.n[1];
      const size_t f = (size_t)S + (size_t)G.D * G.H * W2 + (size_t)G.D * H1 * W2 + (size_t)G.big[j][2] * G.n[2] +
                       (size_t)G.big[j][1] * G.n[1] + (size_t)G.big[j][0] * G.n[0];
      fl = f > fl ? f : fl;
    }
    const size_t smem = fl * sizeof(float);
    if (smem > 200 * 1024) { set_error("pwa_bwd: level of %d voxels too large for the scatter adjoint", S); return VX_ERR_UNSUPPORTED; }
    VX_SET_SMEM(pwa_scatter_bwd_kernel, smem);
    VX_LAUNCH(pwa_scatter_bwd_kernel, dim3(P.cv, B, M), dim3(256), smem, st, G, A);
    VX_TRY(check_launch("pwa_scatter_bwd_kernel"));
  }
  // ---- attention backward
  {
    AttnArgs A{};
    A.Q = SV(SV_QT); A.K = SV(SV_KT); A.V = SV(SV_VT); A.biasT = biasT; A.biasN = biasN; A.O = (float*)SV(SV_OT); A.lse = (float*)SV(SV_LSE);
    A.dO = dOt; A.dQ = dQt; A.dK = dKt; A.dV = dVt; A.dbiasT = dbiasT;
    A.B = B; A.heads = G.heads; A.Ns = G.Ns; A.L = G.L; A.l = G.l;
    A.scale = 1.0f / sqrtf((float)P.cq_h); A.drop_p = attn_p; A.seed = d->seed; A.seed_dev = get_seed_dev();
    VX_TRY(dispatch_attn(A, P.cq_h, P.cv_h, true, st));
    const int trows = (2 * G.n[0] - 1) * (2 * G.n[1] - 1) * (2 * G.n[2] - 1);
    int parts = cdiv((long long)G.l * G.l, 256 * 8);           // ~8 contributions per thread
    if (parts > 32) parts = 32;
    VX_SET_SMEM(pwa_bias_bwd_kernel, sizeof(float) * (size_t)trows);
    VX_LAUNCH(pwa_bias_bwd_kernel, dim3(parts, G.heads), dim3(256), sizeof(float) * (size_t)trows, st, (const float*)dbiasT, index,
              dtable, G.heads, G.l, trows);
    VX_TRY(check_launch("pwa_bias_bwd_kernel"));
  }
  // ---- gather adjoint -> full-resolution dq, dk, dv
  auto dQf = [&](int m) { return dqkv + (size_t)m * cqkv * BS; };
  auto dKf = [&](int m) { return dQf(m) + (size_t)P.cqk * BS; };
  auto dVf = [&](int m) { return dQf(m) + (size_t)2 * P.cqk * BS; };
  {
    GatherBwdArgs A{};
    A.dtok[0] = dQt; A.dtok[1] = dKt; A.dtok[2] = dVt;
    A.arg[0] = (const int*)SV(SV_ARGQ); A.arg[1] = (const int*)SV(SV_ARGK); A.arg[2] = (const int*)SV(SV_ARGV);
    for (int m = 0; m < M; ++m) { A.dst[0][m] = dQf(m); A.dst[1][m] = dKf(m); A.dst[2][m] = dVf(m); }
    A.Ct[0] = A.Ct[1] = P.cqk; A.Ct[2] = P.cv; A.cper[0] = A.cper[1] = P.cq_h; A.cper[2] = P.cv_h;
    if (G.D + G.H + G.W > GB_MAX_AXIS) { set_error("pwa gather bwd: extent beyond the kernel's tables"); return VX_ERR_UNSUPPORTED; }
    int blocks = cdiv(S, 256);                          // one thread per voxel of a (batch, scale, head) plane
    while (blocks > 1 && (long long)blocks * B * G.nb * G.heads * 3 * M > 16LL * kSMs) blocks = (blocks + 1) / 2;
    VX_LAUNCH(pwa_gather_bwd_kernel, dim3(blocks, B * G.nb * G.heads, 3 * M), dim3(256), 0, st, G, A);
    VX_TRY(check_launch("pwa_gather_bwd_kernel"));
  }
  // ---- projection backward
  {
    PwBatch pb{}; pb.nprob = M; pb.B = B; pb.S = S;     // dln1 = Wq^T dq + Wk^T dk + Wv^T dv
    for (int m = 0; m < M; ++m) {
      PwProblem& p = pb.p[m];
      p.src[0] = PwSrc{dQf(m), P.cqk}; p.src[1] = PwSrc{dKf(m), P.cqk}; p.src[2] = PwSrc{dVf(m), P.cv}; p.nsrc = 3; p.Ci = (int)cqkv;
      p.seg[0] = PwSeg{PRM(m, PP_WQ), nullptr, C, P.cqk, dln1 + (size_t)m * C * BS};
      p.seg[1] = PwSeg{PRM(m, PP_WK), nullptr, C, P.cqk, nullptr};
      p.seg[2] = PwSeg{PRM(m, PP_WV), nullptr, C, P.cv, nullptr};
      p.nseg = 3; p.Co = C; p.transposed = 1;
    }
    VX_TRY(pw_forward(pb, st));
    WgBatch wb{}; wb.nprob = 3 * M; wb.B = B; wb.S = S;
    for (int m = 0; m < M; ++m) {
      const float* dsrc[3] = {dQf(m), dKf(m), dVf(m)};
      const int co[3] = {P.cqk, P.cqk, P.cv};
      const int wi[3] = {PP_WQ, PP_WK, PP_WV};
      for (int k = 0; k < 3; ++k) {
        WgProblem& a = wb.p[3 * m + k];
        a.dY = dsrc[k]; a.Co = co[k]; a.src[0] = PwSrc{SV(SV_XHAT1) + (size_t)m * C * BS, C}; a.nsrc = 1; a.Ci = C;
        a.xpro = PRO_AFFINE; a.xa = PRM(m, PP_LN1W); a.xc = PRM(m, PP_LN1B); a.x_bstride = 0;
        a.dW = DP(m, wi[k]); a.ld = C; a.db = DP(m, wi[k] + 1);
      }
    }
    VX_TRY(pw_wgrad(wb, st));
    LnBwdBatch Lb{}; Lb.n = M; Lb.B = B; Lb.C = C; Lb.S = S; Lb.dx_add_scale = 2.f;   // dx = LN1_bwd(dln1) + 2 dy
    for (int m = 0; m < M; ++m) {
      Lb.dout[m] = dln1 + (size_t)m * C * BS; Lb.xhat[m] = SV(SV_XHAT1) + (size_t)m * C * BS; Lb.rstd[m] = SV(SV_RSTD1) + (size_t)m * BS;
      Lb.gamma[m] = PRM(m, PP_LN1W); Lb.dx_add[m] = dy + (size_t)m * C * BS; Lb.dx[m] = DX(m);
      Lb.dgamma[m] = DP(m, PP_LN1W); Lb.dbeta[m] = DP(m, PP_LN1B);
    }
    VX_TRY(ln_backward(Lb, st));
  }
  return VX_OK;
}
