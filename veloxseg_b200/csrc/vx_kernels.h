// Internal host-side launchers shared by the op translation units of libveloxseg_sm100.
#pragma once
#include "vx_common.cuh"

namespace vx {

// ---------------------------------------------------------------------------------------------------
// Channel contraction ("1x1x1 conv") over NCDHW:  Y[b,co,s] = epi( sum_ci W[co,ci] * pro(X[b,ci,s]) + bias[co] )
// ---------------------------------------------------------------------------------------------------
enum { PRO_NONE = 0, PRO_AFFINE = 1, PRO_GELU = 2, PRO_DROPOUT = 3, PRO_GELU_DROPOUT = 4 };

struct PwSrc { const float* ptr; int C; };                 // input segment (B, C, S)
struct PwSeg { const float* W; const float* bias; int ld; int n; float* out; };

struct PwProblem {
  PwSrc src[4]; int nsrc; int Ci;
  // forward orientation: segments split the OUTPUT channels; seg.W is (n, Ci) with row stride ld, each segment
  // writes its own tensor seg.out (B, n, S).
  // transposed orientation: segments split the INPUT channels (same order as src); seg.W is (n, Co) with row
  // stride ld and the logical weight is W^T; the single output is seg[0].out (B, Co, S).
  PwSeg seg[3]; int nseg; int Co;
  int transposed;
  int pro;                        // PRO_*
  const float* pro_a; const float* pro_c; int pro_bstride;   // PRO_AFFINE: x*a[b*bstride+ci] + c[b*bstride+ci]
  float pro_drop_p; uint64_t pro_seed; uint32_t pro_site;    // PRO_DROPOUT / PRO_GELU_DROPOUT (mask indexed like the input)
  int act;                        // 1: GELU on the output
  const float* mulgrad;           // if set: out *= GELU'(mulgrad[b,co,s])
  float drop_p; uint64_t seed; uint32_t site;                // dropout on (acc + bias), before the residual
  const float* res; float res_scale;                         // out += res_scale * res[b,co,s]
  const float* res2;                                         // out += res2[b,co,s]
};

struct PwBatch { PwProblem p[VX_MAX_MODAL]; int nprob; int B; int S; const unsigned long long* seed_dev; int prec; };

int pw_forward(const PwBatch& batch, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------------
// Two-layer channel MLP of the small levels in one launch (JLC channel_conv, conv_blocks.py:62-69; PWA FFN,
// attention_utils.py:45-71):   hpre = W1 (a x + c) + b1;   y = res_scale * res + Drop2( W2 Drop1(GELU(hpre)) + b2 )
// hpre is written out (the backward pass needs it), the hidden activation never leaves shared memory.
// ---------------------------------------------------------------------------------------------------
struct FfnProblem {
  const float* x; int C;                                         // (B, C, S)
  const float* pro_a; const float* pro_c; int pro_bstride;       // prologue affine per (b * bstride + c); null = identity
  const float* W1; const float* b1; int eC;                      // (eC, C), (eC)
  float* hpre;                                                   // (B, eC, S)
  const float* W2; const float* b2;                              // (C, eC), (C)
  float mid_drop_p; uint64_t mid_seed; uint32_t mid_site;        // dropout on GELU(hpre), mask indexed like hpre
  float drop_p; uint64_t seed; uint32_t site;                    // dropout on the output, mask indexed like y
  const float* res; float res_scale;                             // y += res_scale * res
  float* y;                                                      // (B, C, S)
};
struct FfnBatch { FfnProblem p[VX_MAX_MODAL]; int nprob; int B; int S; const unsigned long long* seed_dev; };
// VX_OK when launched, 1 when the shape is outside the kernel's range (the caller runs the two contractions instead)
int pw_ffn_small(const FfnBatch& batch, cudaStream_t stream);
// Data gradient of the same MLP in one launch:  dh = (W2^T (dy * mask2)) * GELU'(hpre) * mask1  (written out: the weight
// gradients read it);  dx = W1^T dh  (the gradient at the prologue's output: the caller owns the norm backward).
struct FfnBwdProblem {
  const float* dy; int C;                                        // (B, C, S)
  float out_drop_p; uint64_t out_seed; uint32_t out_site;        // mask of the output dropout, indexed like dy
  const float* W2; int eC;                                       // (C, eC)
  const float* hpre;                                             // (B, eC, S)
  float mid_drop_p; uint64_t mid_seed; uint32_t mid_site;        // mask on GELU(hpre), indexed like hpre
  float* dh;                                                     // (B, eC, S)
  const float* W1;                                               // (eC, C)
  float* dx;                                                     // (B, C, S)
};
struct FfnBwdBatch { FfnBwdProblem p[VX_MAX_MODAL]; int nprob; int B; int S; const unsigned long long* seed_dev; };
int pw_ffn_small_bwd(const FfnBwdBatch& batch, cudaStream_t stream);
// the same two ops for S >= 1024 on tcgen05 (pw_ffn_tc.cu): the first contraction's accumulator becomes the second's A operand
// inside tensor memory.  Same return convention (1 = not eligible: bf16 mode, channel counts off the MMA grid, option off).
int pw_ffn_tc(const FfnBatch& batch, cudaStream_t stream);
int pw_ffn_tc_bwd(const FfnBwdBatch& batch, cudaStream_t stream);
void pw_ffn_tc_set(int on);

#if defined(__CUDACC__) || defined(VX_EMU)
// Address of one input element of the (possibly multi-source) X operand.
VX_DEV const float* pw_x_ptr(const PwProblem& P, int b, int cg, int v, int S) {
  int c = cg, s = 0;
  while (s < P.nsrc - 1 && c >= P.src[s].C) { c -= P.src[s].C; ++s; }
  return P.src[s].ptr + ((size_t)b * P.src[s].C + c) * S + v;
}
VX_DEV const float* pw_w_ptr(const PwProblem& P, int co, int ci) {
  int off = 0;
  if (!P.transposed) {
    for (int s = 0; s < P.nseg; ++s) {
      if (co < off + P.seg[s].n) return P.seg[s].W + (size_t)(co - off) * P.seg[s].ld + ci;
      off += P.seg[s].n;
    }
  } else {
    for (int s = 0; s < P.nseg; ++s) {
      if (ci < off + P.seg[s].n) return P.seg[s].W + (size_t)(ci - off) * P.seg[s].ld + co;
      off += P.seg[s].n;
    }
  }
  return P.seg[0].W;
}
#endif

#if defined(__CUDACC__) || defined(VX_EMU)
// Out-of-line pieces of the contraction prologue / epilogue.  GELU, GELU' and the dropout mask
// are tens to hundreds of instructions each; inlined at every unrolled use they push a kernel far beyond the instruction
// caches and it becomes fetch-bound (measured: pw_tc_kernel at 94 KB of SASS stalled 3-5 warps per issue on
// `no_instruction`).  One shared copy per kernel instead.
static __device__ VX_NOINLINE float pw_pro_heavy(int pro, float x, uint64_t seed, uint32_t site, uint64_t idx, float p, float pinv) {
  if (pro == PRO_GELU || pro == PRO_GELU_DROPOUT) x = gelu_f(x);
  if (pro == PRO_DROPOUT || pro == PRO_GELU_DROPOUT) x *= dropout_scale(seed, site, idx, p, pinv);
  return x;
}
static __device__ VX_NOINLINE float pw_epi_heavy(float y, int act, bool has_mg, float mg, float drop_p, uint64_t seed, uint32_t site,
                                                 uint64_t idx, float dinv) {
  if (act == 1) y = gelu_f(y);
  if (has_mg) y *= gelu_grad_f(mg);
  if (drop_p > 0.f) y *= dropout_scale(seed, site, idx, drop_p, dinv);
  return y;
}
// Row of the weight operand that multiplies input channel ci for output channels co.. (transposed orientation) or the row
// of output channel co starting at input channel ci (forward orientation), with the row stride of its segment.
VX_DEV const float* pw_w_row(const PwProblem& P, int co, int ci, int& ld) {
  int off = 0;
  if (!P.transposed) {
    for (int s = 0; s < P.nseg; ++s) {
      if (co < off + P.seg[s].n) { ld = P.seg[s].ld; return P.seg[s].W + (size_t)(co - off) * P.seg[s].ld + ci; }
      off += P.seg[s].n;
    }
  } else {
    for (int s = 0; s < P.nseg; ++s) {
      if (ci < off + P.seg[s].n) { ld = P.seg[s].ld; return P.seg[s].W + (size_t)(ci - off) * P.seg[s].ld + co; }
      off += P.seg[s].n;
    }
  }
  ld = P.seg[0].ld;
  return P.seg[0].W;
}
#endif

// Tensor-core (tcgen05, 3xTF32) variant of pw_forward for the large-voxel problems: returns 0 when launched, 1 when the
// batch does not qualify (the caller then uses the SIMT kernels), negative on error.  pw_tc.cu.
int pw_tc_forward(const PwBatch& batch, cudaStream_t stream);
void pw_tc_set(int enabled);
void jlc_force_vx(int vx);
void pwa_attn_tc_set(int on);
void jlc_set_small_threads(int t);
void jlc_set_ks(int ks);                                  // tuning probe: reduction slices of the level-1/2 conv kernels
void jlc_set_small_max(int s);                            // tuning probe: voxel threshold of the small-volume conv kernels
void jlc_force_tile(int kind, int tz, int ty);            // tuning probe: 0 = automatic
void pw_set_thresholds(int small_max_s, int tc_min_s);   // tuning probes; -1 keeps a value


// dW[co, ci] += sum_{b,s} ypro(dY[b,co,s]) * xpro(X[b,ci,s]);   db[co] += sum ypro(dY)
struct WgProblem {
  const float* dY; int Co;
  PwSrc src[4]; int nsrc; int Ci;
  int xpro; const float* xa; const float* xc; int x_bstride;  // PRO_NONE / PRO_AFFINE / PRO_GELU[_DROPOUT] on X
  float x_drop_p; uint64_t x_seed; uint32_t x_site;
  float y_drop_p; uint64_t y_seed; uint32_t y_site;           // dropout mask on dY (0 disables)
  float* dW; int ld;                                          // accumulated with atomics: caller zeroes
  float* db;                                                  // may be null
};
struct WgBatch { WgProblem p[VX_MAX_MODAL * 3]; int nprob; int B; int S; const unsigned long long* seed_dev; int prec; };

// Device pointer whose value the dropout kernels add to their seeds (graph-replay-safe masks).  Thread-local: set by the
// op entry point, picked up by pw_forward / pw_wgrad.
void set_seed_dev(const void* p);
const unsigned long long* get_seed_dev();
int pw_wgrad(const WgBatch& batch, cudaStream_t stream);
// Tensor-core (tcgen05, 3xTF32) variant for the large-voxel problems: 0 = launched, 1 = does not qualify.  pw_wgrad_tc.cu.
int pw_wgrad_tc(const WgBatch& batch, cudaStream_t stream);
void pw_wgrad_tc_set(int enabled, int min_s);             // -1 keeps a value

// ---------------------------------------------------------------------------------------------------
// Weight-gradient side stream.  Inside a backward op the weight gradients are leaves (nothing in the op consumes them)
// while the data gradients form the chain the next op waits for, and every kernel here is a sub-wave grid.  side_fork()
// returns a library-owned stream that has been made to wait for everything enqueued on `main` so far; the weight-gradient
// launchers run there.  SideJoin (one per backward entry point) makes `main` wait for it before the op returns, so callers
// see ordinary single-stream semantics; under CUDA-graph capture the fork/join become parallel graph branches.
// ---------------------------------------------------------------------------------------------------
cudaStream_t side_fork(cudaStream_t main);
void side_wait(cudaStream_t main);      // main waits for what the side stream has been given so far (stays forked)
void side_join(cudaStream_t main);
void side_set(int enabled);
struct SideJoin {
  cudaStream_t st;
  explicit SideJoin(cudaStream_t s) : st(s) {}
  ~SideJoin() { side_join(st); }
};

// ---------------------------------------------------------------------------------------------------
// One launch that zeroes every atomically-accumulated gradient buffer of an op (the reference's autograd allocates them
// zeroed; a cudaMemsetAsync per buffer is ~12-34 extra graph nodes per backward op).
// ---------------------------------------------------------------------------------------------------
constexpr int VX_ZERO_MAX = 72;     // >= VX_MAX_MODAL * 16 + 2 (PWA block backward)
struct ZeroList {
  float* ptr[VX_ZERO_MAX]; unsigned n[VX_ZERO_MAX]; int count;
  ZeroList() : count(0) {}
  void add(void* p, size_t nfloats) { if (p && nfloats && count < VX_ZERO_MAX) { ptr[count] = (float*)p; n[count] = (unsigned)nfloats; ++count; } else if (p && nfloats) overflow = 1; }
  int overflow = 0;
};
int zero_many(const ZeroList& z, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------------
// norms
// ---------------------------------------------------------------------------------------------------
// InstanceNorm rows: y = (x-mean)*rstd [+ addend]; stats[row] = (mean, rstd)
int inorm_rows_fwd(const float* x, const float* addend, float* y, float* stats, int rows, int S, float eps,
                   cudaStream_t stream);
// dx = rstd*(dy - mean(dy) - xhat*mean(dy*xhat)) [+ dx_add]
// db (optional, C floats): += per-channel sums of dx (row r belongs to channel r % C); zeroed by the call
int inorm_rows_bwd(const float* dy, const float* x, const float* stats, const float* dx_add, float* dx, int rows,
                   int S, cudaStream_t stream, float* db = nullptr, int C = 0);
// per-(b,c) affine (a = rstd, c = -mean*rstd) from (mean, rstd) stats
int stats_to_affine(const float* stats, float* a, float* c, int rows, cudaStream_t stream);

// channel-first LayerNorm core: xhat = (x-mean_c)*rstd ; rstd (B,S).  Batched over up to VX_MAX_MODAL tensors.
struct LnBatch {
  const float* x[VX_MAX_MODAL]; float* xhat[VX_MAX_MODAL]; float* rstd[VX_MAX_MODAL];
  int n; int B; int C; int S; float eps;
};
int ln_forward(const LnBatch& b, cudaStream_t stream);
// g = gamma*dout ; dx = rstd*(g - mean_c(g) - xhat*mean_c(g*xhat)) + dx_add ; dgamma += sum dout*xhat ; dbeta += sum dout
struct LnBwdBatch {
  const float* dout[VX_MAX_MODAL]; const float* xhat[VX_MAX_MODAL]; const float* rstd[VX_MAX_MODAL];
  const float* gamma[VX_MAX_MODAL]; const float* dx_add[VX_MAX_MODAL]; float dx_add_scale;
  float* dx[VX_MAX_MODAL]; float* dgamma[VX_MAX_MODAL]; float* dbeta[VX_MAX_MODAL];
  int n; int B; int C; int S;
};
int ln_backward(const LnBwdBatch& b, cudaStream_t stream);

// Convolutions of the glue layers: conv3_tc.cu (dense 3x3x3, tcgen05) and conv_simt.cu (strided / transposed); C ABI vx_conv_*.
void conv3_trace_set(int on);
int conv3_trace_read(long long* out, int n);
bool conv3_tc_supported(const vx_conv_desc* d);
size_t conv3_tc_workspace(const vx_conv_desc* d);

}  // namespace vx
