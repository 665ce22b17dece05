// Modal mixer, InstanceNorm, SDKT Gram / loss, LayerNorm + 1x1 (PatchMerging tail).  fp32, sm_100a.
#include "vx_kernels.h"

#ifdef VX_EMU
#define __grid_constant__
#endif

#define VX_TRY(expr) do { int _rc = (expr); if (_rc != VX_OK) return _rc; } while (0)

namespace vx {
static inline size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }

// ---------------------------------------------------------------------------------------------------
// Gram:  G[b] = X X^T / (C S).  Split-S partials (one CTA per 512-voxel chunk), then a fixed-order reduce.
// ---------------------------------------------------------------------------------------------------
constexpr int GRAM_TV = 512;

__global__ void __launch_bounds__(256) gram_partial_kernel(const float* __restrict__ x, float* __restrict__ part, int C,
                                                           int S) {
  VX_PDL_ENTRY();
  VX_DYN_SMEM(float, sx);   // [GRAM_TV][C+1]
  const int b = blockIdx.y, ck = blockIdx.x, CP = C + 1;
  const int v0 = ck * GRAM_TV;
  for (int idx = threadIdx.x; idx < C * GRAM_TV; idx += blockDim.x) {
    const int c = idx / GRAM_TV, v = idx % GRAM_TV;
    sx[v * CP + c] = (v0 + v < S) ? __ldg(x + ((size_t)b * C + c) * S + v0 + v) : 0.f;
  }
  __syncthreads();
  for (int p = threadIdx.x; p < C * C; p += blockDim.x) {
    const int m = p / C, n = p % C;
    float acc = 0.f;
    for (int v = 0; v < GRAM_TV; ++v) acc = fmaf(sx[v * CP + m], sx[v * CP + n], acc);
    part[((size_t)b * gridDim.x + ck) * C * C + p] = acc;
  }
}

__global__ void gram_reduce_kernel(const float* __restrict__ part, float* __restrict__ G, int CC, int nchunk,
                                   float scale) {
  VX_PDL_ENTRY();
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= CC) return;
  float acc = 0.f;
  for (int k = 0; k < nchunk; ++k) acc += part[((size_t)b * nchunk + k) * CC + p];
  G[(size_t)b * CC + p] = acc * scale;
}

// dx[b,c,s] = sum_n (dG[b,c,n] + dG[b,n,c]) x[b,n,s] / (C S)
template <int CT>
__global__ void __launch_bounds__(128) gram_bwd_kernel(const float* __restrict__ dG, const float* __restrict__ x,
                                                       float* __restrict__ dx, int C, int S, float scale) {
  VX_PDL_ENTRY();
  VX_DYN_SMEM(float, w);   // [C][C] symmetrised
  const int b = blockIdx.y;
  for (int p = threadIdx.x; p < C * C; p += blockDim.x) {
    const int m = p / C, n = p % C;
    w[p] = (dG[(size_t)b * C * C + m * C + n] + dG[(size_t)b * C * C + n * C + m]) * scale;
  }
  __syncthreads();
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= S) return;
  const float* xp = x + (size_t)b * C * S + v;
  float* dp = dx + (size_t)b * C * S + v;
  if (CT > 0) {
    float xr[CT > 0 ? CT : 1];
#pragma unroll
    for (int n = 0; n < CT; ++n) xr[n] = __ldg(xp + (size_t)n * S);
#pragma unroll
    for (int c = 0; c < CT; ++c) {
      float acc = 0.f;
#pragma unroll
      for (int n = 0; n < CT; ++n) acc = fmaf(w[c * CT + n], xr[n], acc);
      dp[(size_t)c * S] = acc;
    }
  } else {
    for (int c = 0; c < C; ++c) {
      float acc = 0.f;
      for (int n = 0; n < C; ++n) acc = fmaf(w[c * C + n], __ldg(xp + (size_t)n * S), acc);
      dp[(size_t)c * S] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// SDKT feature loss on Grams (single CTA: n_elem is B*C*C ~ 1 K)
// ---------------------------------------------------------------------------------------------------
struct SdktArgs { const float* gs; const float* gt[VX_MAX_MODAL]; float* dgt[VX_MAX_MODAL]; float* dgs; const float* dloss; float* loss; int n; int T; };

__global__ void __launch_bounds__(256) sdkt_loss_fwd_kernel(const __grid_constant__ SdktArgs A) {
  VX_PDL_ENTRY();
  __shared__ float red[33];
  float s = 0.f;
  for (int i = threadIdx.x; i < A.n; i += blockDim.x) {
    const float g = A.gs[i];
    for (int t = 0; t < A.T; ++t) { const float d = g - A.gt[t][i]; s = fmaf(d, d, s); }
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) A.loss[0] = s / ((float)A.n * (float)A.T);
}

__global__ void __launch_bounds__(256) sdkt_loss_bwd_kernel(const __grid_constant__ SdktArgs A) {
  VX_PDL_ENTRY();
  const float k = 2.0f * A.dloss[0] / ((float)A.n * (float)A.T);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += gridDim.x * blockDim.x) {
    const float g = A.gs[i];
    float acc = 0.f;
    for (int t = 0; t < A.T; ++t) {
      const float d = (g - A.gt[t][i]) * k;
      acc += d;
      A.dgt[t][i] = -d;
    }
    A.dgs[i] = acc;
  }
}

}  // namespace vx

using namespace vx;

// ---------------------------------------------------------------------------------------------------
// modal mixer
// ---------------------------------------------------------------------------------------------------
static int mixer_check(const vx_mixer_desc* d, int& K) {
  if (!d || d->B <= 0 || d->S <= 0 || d->n_streams <= 0 || d->n_streams > VX_MAX_MODAL || d->C_out <= 0) {
    set_error("mixer: bad descriptor"); return VX_ERR_BAD_DESC;
  }
  K = 0;
  for (int m = 0; m < d->n_streams; ++m) {
    if (d->stream_ch[m] <= 0) { set_error("mixer: bad stream channels"); return VX_ERR_BAD_DESC; }
    K += d->stream_ch[m];
  }
  return VX_OK;
}

extern "C" size_t vx_mixer_workspace(const vx_mixer_desc* d) {
  int K;
  if (mixer_check(d, K) != VX_OK) return 0;
  return align256(sizeof(float) * (size_t)d->B * d->C_out * d->S);
}

extern "C" int vx_mixer_fwd(const vx_mixer_desc* d, const void* const* in, void* const* out, void* workspace,
                            size_t workspace_bytes, vx_stream_t stream) {
  int K;
  VX_TRY(mixer_check(d, K));
  set_seed_dev(nullptr);
  prof_scope("mixer_fwd B%d K%d N%d S%d", d->B, K, d->C_out, d->S);
  (void)workspace; (void)workspace_bytes;
  cudaStream_t st = (cudaStream_t)stream;
  const int M = d->n_streams;
  const float* W = (const float*)in[M];
  const float* bias = (const float*)in[M + 1];
  const float* addend = d->has_addend ? (const float*)in[M + 2] : nullptr;
  float* y = (float*)out[0];
  float* t = (float*)out[1];
  float* stats = (float*)out[2];
  PwBatch pb{}; pb.nprob = 1; pb.B = d->B; pb.S = d->S;
  PwProblem& p = pb.p[0];
  for (int m = 0; m < M; ++m) p.src[m] = PwSrc{(const float*)in[m], d->stream_ch[m]};
  p.nsrc = M; p.Ci = K;
  p.seg[0] = PwSeg{W, bias, K, d->C_out, t}; p.nseg = 1; p.Co = d->C_out;
  VX_TRY(pw_forward(pb, st));
  return inorm_rows_fwd(t, addend, y, stats, d->B * d->C_out, d->S, d->eps, st);
}

extern "C" int vx_mixer_bwd(const vx_mixer_desc* d, const void* const* in, void* const* out, void* workspace,
                            size_t workspace_bytes, vx_stream_t stream) {
  int K;
  VX_TRY(mixer_check(d, K));
  set_seed_dev(nullptr);
  prof_scope("mixer_bwd B%d K%d N%d S%d", d->B, K, d->C_out, d->S);
  const size_t need = vx_mixer_workspace(d);
  if (!workspace || workspace_bytes < need) { set_error("mixer_bwd: workspace %zu < %zu", workspace_bytes, need); return VX_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  SideJoin side_guard(st);
  const int M = d->n_streams, Co = d->C_out;
  const float* dy = (const float*)in[0];
  const float* W = (const float*)in[M + 1];
  const float* t = (const float*)in[M + 2];
  const float* stats = (const float*)in[M + 3];
  float* dW = (float*)out[M];
  float* db = (float*)out[M + 1];
  float* dt = (float*)workspace;
  // dW / db are accumulated by the weight-gradient kernel on the side stream only: they are zeroed there
  { ZeroList zl; zl.add(dW, (size_t)Co * K); zl.add(db, Co); VX_TRY(zero_many(zl, side_fork(st))); }
  VX_TRY(inorm_rows_bwd(dy, t, stats, nullptr, dt, d->B * Co, d->S, st));
  WgBatch wb{}; wb.nprob = 1; wb.B = d->B; wb.S = d->S;
  WgProblem& w = wb.p[0];
  w.dY = dt; w.Co = Co;
  for (int m = 0; m < M; ++m) w.src[m] = PwSrc{(const float*)in[1 + m], d->stream_ch[m]};
  w.nsrc = M; w.Ci = K; w.dW = dW; w.ld = K; w.db = db;
  VX_TRY(pw_wgrad(wb, st));
  PwBatch pb{}; pb.nprob = M; pb.B = d->B; pb.S = d->S;
  int off = 0;
  for (int m = 0; m < M; ++m) {
    PwProblem& p = pb.p[m];
    p.src[0] = PwSrc{dt, Co}; p.nsrc = 1; p.Ci = Co;
    p.seg[0] = PwSeg{W + off, nullptr, K, Co, (float*)out[m]}; p.nseg = 1; p.Co = d->stream_ch[m]; p.transposed = 1;
    off += d->stream_ch[m];
  }
  return pw_forward(pb, st);
}

// ---------------------------------------------------------------------------------------------------
// InstanceNorm
// ---------------------------------------------------------------------------------------------------
extern "C" int vx_inorm_fwd(const vx_inorm_desc* d, const void* const* in, void* const* out, vx_stream_t stream) {
  if (!d || d->rows <= 0 || d->S <= 0) { set_error("inorm: bad descriptor"); return VX_ERR_BAD_DESC; }
  prof_scope("inorm_fwd R%d S%d", d->rows, d->S);
  return inorm_rows_fwd((const float*)in[0], d->has_addend ? (const float*)in[1] : nullptr, (float*)out[0],
                        (float*)out[1], d->rows, d->S, d->eps, (cudaStream_t)stream);
}

extern "C" int vx_inorm_bwd(const vx_inorm_desc* d, const void* const* in, void* const* out, vx_stream_t stream) {
  if (!d || d->rows <= 0 || d->S <= 0) { set_error("inorm: bad descriptor"); return VX_ERR_BAD_DESC; }
  prof_scope("inorm_bwd R%d S%d", d->rows, d->S);
  if (d->bias_channels < 0 || (d->bias_channels > 0 && d->rows % d->bias_channels)) { set_error("inorm_bwd: bad bias_channels"); return VX_ERR_BAD_DESC; }
  return inorm_rows_bwd((const float*)in[0], (const float*)in[1], (const float*)in[2], nullptr, (float*)out[0], d->rows,
                        d->S, (cudaStream_t)stream, d->bias_channels > 0 ? (float*)out[1] : nullptr, d->bias_channels);
}

// ---------------------------------------------------------------------------------------------------
// SDKT
// ---------------------------------------------------------------------------------------------------
extern "C" size_t vx_gram_workspace(const vx_gram_desc* d) {
  if (!d || d->B <= 0 || d->C <= 0 || d->S <= 0) return 0;
  return align256(sizeof(float) * (size_t)d->B * cdiv(d->S, GRAM_TV) * d->C * d->C);
}

extern "C" int vx_gram_fwd(const vx_gram_desc* d, const void* const* in, void* const* out, void* workspace,
                           size_t workspace_bytes, vx_stream_t stream) {
  const size_t need = vx_gram_workspace(d);
  if (!need || d->C > 64) { set_error("gram: bad descriptor"); return VX_ERR_BAD_DESC; }
  if (!workspace || workspace_bytes < need) { set_error("gram_fwd: workspace %zu < %zu", workspace_bytes, need); return VX_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  prof_scope("gram_fwd B%d C%d S%d", d->B, d->C, d->S);
  const int nchunk = cdiv(d->S, GRAM_TV), CC = d->C * d->C;
  const size_t smem = sizeof(float) * (size_t)GRAM_TV * (d->C + 1);
  VX_SET_SMEM(gram_partial_kernel, smem);
  VX_LAUNCH(gram_partial_kernel, dim3(nchunk, d->B), dim3(256), smem, st, (const float*)in[0], (float*)workspace, d->C, d->S);
  VX_TRY(check_launch("gram_partial_kernel"));
  VX_LAUNCH(gram_reduce_kernel, dim3(cdiv(CC, 128), d->B), dim3(128), 0, st, (const float*)workspace, (float*)out[0], CC,
            nchunk, 1.0f / ((float)d->C * (float)d->S));
  return check_launch("gram_reduce_kernel");
}

extern "C" int vx_gram_bwd(const vx_gram_desc* d, const void* const* in, void* const* out, vx_stream_t stream) {
  if (!d || d->B <= 0 || d->C <= 0 || d->C > 64 || d->S <= 0) { set_error("gram: bad descriptor"); return VX_ERR_BAD_DESC; }
  cudaStream_t st = (cudaStream_t)stream;
  prof_scope("gram_bwd B%d C%d S%d", d->B, d->C, d->S);
  const size_t smem = sizeof(float) * (size_t)d->C * d->C;
  const float scale = 1.0f / ((float)d->C * (float)d->S);
  dim3 grid(cdiv(d->S, 128), d->B);
  if (d->C == 16) VX_LAUNCH(gram_bwd_kernel<16>, grid, dim3(128), smem, st, (const float*)in[0], (const float*)in[1], (float*)out[0], d->C, d->S, scale);
  else if (d->C == 32) VX_LAUNCH(gram_bwd_kernel<32>, grid, dim3(128), smem, st, (const float*)in[0], (const float*)in[1], (float*)out[0], d->C, d->S, scale);
  else VX_LAUNCH(gram_bwd_kernel<0>, grid, dim3(128), smem, st, (const float*)in[0], (const float*)in[1], (float*)out[0], d->C, d->S, scale);
  return check_launch("gram_bwd_kernel");
}

static int sdkt_args(const vx_sdkt_loss_desc* d, SdktArgs& A) {
  if (!d || d->n_elem <= 0 || d->n_teachers <= 0 || d->n_teachers > VX_MAX_MODAL) { set_error("sdkt_loss: bad descriptor"); return VX_ERR_BAD_DESC; }
  A = SdktArgs{};
  A.n = d->n_elem; A.T = d->n_teachers;
  return VX_OK;
}

extern "C" int vx_sdkt_loss_fwd(const vx_sdkt_loss_desc* d, const void* const* in, void* const* out, vx_stream_t stream) {
  SdktArgs A;
  VX_TRY(sdkt_args(d, A));
  A.gs = (const float*)in[0];
  for (int t = 0; t < A.T; ++t) A.gt[t] = (const float*)in[1 + t];
  A.loss = (float*)out[0];
  prof_scope("sdkt_loss_fwd");
  VX_LAUNCH(sdkt_loss_fwd_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, A);
  return check_launch("sdkt_loss_fwd_kernel");
}

extern "C" int vx_sdkt_loss_bwd(const vx_sdkt_loss_desc* d, const void* const* in, void* const* out, vx_stream_t stream) {
  SdktArgs A;
  VX_TRY(sdkt_args(d, A));
  A.dloss = (const float*)in[0];
  A.gs = (const float*)in[1];
  for (int t = 0; t < A.T; ++t) { A.gt[t] = (const float*)in[2 + t]; A.dgt[t] = (float*)out[1 + t]; }
  A.dgs = (float*)out[0];
  prof_scope("sdkt_loss_bwd");
  VX_LAUNCH(sdkt_loss_bwd_kernel, dim3(cdiv(A.n, 256)), dim3(256), 0, (cudaStream_t)stream, A);
  return check_launch("sdkt_loss_bwd_kernel");
}

// ---------------------------------------------------------------------------------------------------
// LayerNorm(channels_first) + 1x1 conv without bias (PatchMerging tail)
// ---------------------------------------------------------------------------------------------------
extern "C" size_t vx_lnpw_workspace(const vx_lnpw_desc* d) {
  if (!d || d->B <= 0 || d->C_in <= 0 || d->C_out <= 0 || d->S <= 0) return 0;
  return align256(sizeof(float) * (size_t)d->B * d->C_in * d->S);
}

extern "C" int vx_lnpw_fwd(const vx_lnpw_desc* d, const void* const* in, void* const* out, void* workspace,
                           size_t workspace_bytes, vx_stream_t stream) {
  if (!vx_lnpw_workspace(d)) { set_error("lnpw: bad descriptor"); return VX_ERR_BAD_DESC; }
  (void)workspace; (void)workspace_bytes;
  set_seed_dev(nullptr);
  prof_scope("lnpw_fwd B%d Ci%d Co%d S%d", d->B, d->C_in, d->C_out, d->S);
  cudaStream_t st = (cudaStream_t)stream;
  float* y = (float*)out[0];
  float* xhat = (float*)out[1];
  float* rstd = (float*)out[2];
  LnBatch L{}; L.n = 1; L.B = d->B; L.C = d->C_in; L.S = d->S; L.eps = d->eps;
  L.x[0] = (const float*)in[0]; L.xhat[0] = xhat; L.rstd[0] = rstd;
  VX_TRY(ln_forward(L, st));
  PwBatch pb{}; pb.nprob = 1; pb.B = d->B; pb.S = d->S;
  PwProblem& p = pb.p[0];
  p.src[0] = PwSrc{xhat, d->C_in}; p.nsrc = 1; p.Ci = d->C_in;
  p.seg[0] = PwSeg{(const float*)in[3], nullptr, d->C_in, d->C_out, y}; p.nseg = 1; p.Co = d->C_out;
  p.pro = PRO_AFFINE; p.pro_a = (const float*)in[1]; p.pro_c = (const float*)in[2]; p.pro_bstride = 0;
  return pw_forward(pb, st);
}

extern "C" int vx_lnpw_bwd(const vx_lnpw_desc* d, const void* const* in, void* const* out, void* workspace,
                           size_t workspace_bytes, vx_stream_t stream) {
  const size_t need = vx_lnpw_workspace(d);
  if (!need) { set_error("lnpw: bad descriptor"); return VX_ERR_BAD_DESC; }
  if (!workspace || workspace_bytes < need) { set_error("lnpw_bwd: workspace %zu < %zu", workspace_bytes, need); return VX_ERR_WORKSPACE; }
  set_seed_dev(nullptr);
  prof_scope("lnpw_bwd B%d Ci%d Co%d S%d", d->B, d->C_in, d->C_out, d->S);
  cudaStream_t st = (cudaStream_t)stream;
  SideJoin side_guard(st);
  const float* dy = (const float*)in[0];
  const float* xhat = (const float*)in[1];
  const float* rstd = (const float*)in[2];
  const float* gamma = (const float*)in[3];
  const float* W = (const float*)in[4];
  float* dx = (float*)out[0];
  float* dgamma = (float*)out[1];
  float* dbeta = (float*)out[2];
  float* dW = (float*)out[3];
  float* dln = (float*)workspace;
  // zeroed on the side stream (dW is accumulated there; dgamma / dbeta by ln_backward on the main stream after side_wait)
  { ZeroList zl; zl.add(dgamma, d->C_in); zl.add(dbeta, d->C_in); zl.add(dW, (size_t)d->C_in * d->C_out); VX_TRY(zero_many(zl, side_fork(st))); }
  PwBatch pb{}; pb.nprob = 1; pb.B = d->B; pb.S = d->S;
  PwProblem& p = pb.p[0];
  p.src[0] = PwSrc{dy, d->C_out}; p.nsrc = 1; p.Ci = d->C_out;
  p.seg[0] = PwSeg{W, nullptr, d->C_in, d->C_out, dln}; p.nseg = 1; p.Co = d->C_in; p.transposed = 1;
  VX_TRY(pw_forward(pb, st));
  // dW = dy (gamma*xhat + beta)^T
  WgBatch wb{}; wb.nprob = 1; wb.B = d->B; wb.S = d->S;
  WgProblem& w = wb.p[0];
  w.dY = dy; w.Co = d->C_out; w.src[0] = PwSrc{xhat, d->C_in}; w.nsrc = 1; w.Ci = d->C_in;
  w.xpro = PRO_AFFINE; w.xa = gamma; w.xc = (const float*)in[5]; w.x_bstride = 0;
  w.dW = dW; w.ld = d->C_in; w.db = nullptr;
  VX_TRY(pw_wgrad(wb, st));
  LnBwdBatch B{}; B.n = 1; B.B = d->B; B.C = d->C_in; B.S = d->S;
  B.dout[0] = dln; B.xhat[0] = xhat; B.rstd[0] = rstd; B.gamma[0] = gamma; B.dx[0] = dx;
  B.dgamma[0] = dgamma; B.dbeta[0] = dbeta; B.dx_add_scale = 0.f;
  side_wait(st);
  return ln_backward(B, st);
}
