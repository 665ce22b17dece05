// The two-layer MLP of the level-1/2 blocks (S >= 1024) as ONE tcgen05 kernel per direction:
//   forward   y  = res_scale * res + Dropout(W2 (Dropout(GELU(W1 (a x + c) + b1))) + b2)     JLC conv_blocks.py:66-68 (x = IN(o)),
//                                                                                            PWA attention_utils.py:45-71 (x = LN2(y))
//   backward  dh = (W2^T (dy * mask2)) * GELU'(hpre) * mask1 ;  dx = W1^T dh
// Two pw_tc_kernel launches did this before, with the hidden activation (3 C channels) written by the first and read back by the
// second.  Here the first contraction's accumulator is turned into the second contraction's A operand inside tensor memory:
// CTA = 128 voxels of one batch item, thread = voxel = TMEM lane (two threads per lane, dealing 8-column pieces in turns):
//   stage 0   input channels -> prologue (IN / LN affine, or the output-dropout mask of the backward) -> hi / lo -> tcgen05.st
//   MMA 1     D1[128 x eC] = A1 (TMEM, ".ts" form) x B1 (weights, K-major shared memory), 3xTF32
//   stage 1   D1 -> (+b1, hpre out, GELU, mask | x GELU'(hpre), mask, dh out) -> hi over D1, lo beside it
//   MMA 2     D2[128 x C] = A2 (TMEM) x B2, accumulator in the columns A1 occupied
//   stage 2   D2 -> (+b2, mask, residual) -> y | dx
// TMEM columns: 2 eC + 2 C (128 at level 1, 256 at level 2).  The backward reads the same weight matrices transposed, which is a
// different scatter into the K-major shared-memory image and nothing else.  Rolled loops over 8-column pieces throughout (the
// instruction-fetch finding of pw_tc.cu).  fp32 mode only: in bf16 mode the callers keep the two-launch path.
#include "vx_kernels.h"
#include "vx_tc2.cuh"

#ifdef VX_EMU
#define __grid_constant__
#endif

namespace vx {

constexpr int FT_THREADS = 256, FT_ROWS = 128;

struct FfnTcProblem {
  const float* in; int C;                                       // stage-0 input (x | dy): (B, C, S)
  const float* pro_a; const float* pro_c; int pro_bstride;      // forward prologue affine per (b * bstride + c); null = identity
  float in_drop_p; uint64_t in_seed; uint32_t in_site;          // backward: mask on dy (indexed like dy)
  const float* Wa; const float* ba; int eC;                     // stage-1 weights: forward W1 (eC, C) + b1; backward W2 (C, eC) read transposed
  float* hid_out; const float* hid_in;                          // forward: hpre out; backward: dh out, hpre in
  float mid_drop_p; uint64_t mid_seed; uint32_t mid_site;       // mask on the hidden activation (indexed like hpre)
  const float* Wb; const float* bb;                             // stage-2 weights: forward W2 (C, eC) + b2; backward W1 (eC, C) read transposed
  float out_drop_p; uint64_t out_seed; uint32_t out_site;       // forward: mask on the output (indexed like y)
  const float* res; float res_scale;
  float* out;                                                   // y | dx: (B, C, S)
};
struct FfnTcBatch { FfnTcProblem p[VX_MAX_MODAL]; int nprob, B, S, bwd; const unsigned long long* seed_dev; };

// K-major SWIZZLE_NONE image of an (N x K) operand, one block of N x 8 per k-step: 8-row groups 256 B apart, the two 16-byte
// K halves 128 B apart (descriptor LBO = 128, SBO = 256, start = block of the k-step)
VX_DEV int ft_kmajor(int N, int n, int k) { return (k >> 3) * N * 8 + (n >> 3) * 64 + ((k >> 2) & 1) * 32 + (n & 7) * 4 + (k & 3); }

// weights (N x K) -> hi / lo images, four elements of the contiguous axis per step (N, K multiples of 16, 16-byte rows).
// `transposed`: element (n, k) is W[k * N + n] (a quad = 4 consecutive n: four 4-byte stores), else W[n * K + k] (a quad = 4
// consecutive k = one 16-byte store into the core matrix row)
VX_DEV void ft_stage_weights(float* hi, float* lo, const float* __restrict__ W, int N, int K, int transposed, int tid) {
  const int nq = (N * K) >> 2;
#pragma unroll 2
  for (int q = tid; q < nq; q += FT_THREADS) {
    const float4 w = __ldg(reinterpret_cast<const float4*>(W) + q);
    const float wv[4] = {w.x, w.y, w.z, w.w};
    float h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) tc::split(wv[j], h[j], l[j]);
    const int e = q << 2;
    if (transposed) {
      const int n = e % N, k = e / N;
      const int o = ft_kmajor(N, n, k);              // n .. n + 3 stay inside one 8-row group: 4 floats apart
#pragma unroll
      for (int j = 0; j < 4; ++j) { hi[o + 4 * j] = h[j]; lo[o + 4 * j] = l[j]; }
    } else {
      const int n = e / K, k = e % K;
      const int o = ft_kmajor(N, n, k);
      *reinterpret_cast<float4*>(hi + o) = make_float4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<float4*>(lo + o) = make_float4(l[0], l[1], l[2], l[3]);
    }
  }
}

__global__ void __launch_bounds__(FT_THREADS, 3) pw_ffn_tc_kernel(const __grid_constant__ FfnTcBatch batch) {
  VX_PDL_ENTRY();
  const FfnTcProblem& P = batch.p[blockIdx.z];
  const int S = batch.S, C = P.C, eC = P.eC, b = blockIdx.y, bwd = batch.bwd;
  const int tid = threadIdx.x, warp = tid >> 5, lg = warp & 3, sh = warp >> 2;
  const int v = blockIdx.x * FT_ROWS + lg * 32 + (tid & 31);
  const bool live = v < S;
  VX_DYN_SMEM(float, sm);
  float* B1h = sm;                          // [C / 8][eC x 8]
  float* B1l = B1h + (size_t)eC * C;
  float* B2h = B1l + (size_t)eC * C;        // [eC / 8][C x 8]
  float* B2l = B2h + (size_t)eC * C;
  VX_TC_SHARED_BARS(bars, 1);
  VX_TC_SHARED_SLOT(tmem_slot);
  const uint32_t ncols = 2u * eC + 2u * C <= 128u ? 128u : (2u * eC + 2u * C <= 256u ? 256u : 512u);
  if (warp == 0) tc::tmem_alloc(&tmem_slot, ncols);
  if (tid == 0) {
    tc::mbar_init(&bars[0], 1);
    tc::mbar_init_fence();
  }
  const uint64_t soff = batch.seed_dev ? (uint64_t)__ldg(batch.seed_dev) : 0;
  ft_stage_weights(B1h, B1l, P.Wa, eC, C, bwd, tid);
  ft_stage_weights(B2h, B2l, P.Wb, C, eC, bwd, tid);
  tc::fence_async_smem();
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t trow = tmem + ((uint32_t)(lg * 32) << 16);
  const uint32_t A1 = 2u * eC, D2 = 2u * eC;            // stage-0 operand (hi | lo), later the second accumulator

  // ---- stage 0
  {
    const float iinv = P.in_drop_p > 0.f ? 1.0f / (1.0f - P.in_drop_p) : 1.f;
    const uint32_t ikey = P.in_drop_p > 0.f ? rng_key(P.in_seed + soff, P.in_site) : 0u;
#pragma unroll 1
    for (int g = sh; g < (C >> 3); g += 2) {
      float hi[8], lo[8];
      const size_t idx0 = ((size_t)b * C + 8 * g) * S + v;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float x = live ? __ldg(P.in + idx0 + (size_t)j * S) : 0.f;
        if (P.pro_a) { const int q = b * P.pro_bstride + 8 * g + j; x = live ? fmaf(x, __ldg(P.pro_a + q), __ldg(P.pro_c + q)) : 0.f; }
        if (P.in_drop_p > 0.f) x *= keep_from_bits(rng_word(ikey, idx0 + (size_t)j * S), P.in_drop_p, iinv);
        tc::split(x, hi[j], lo[j]);
      }
      tc::tmem_st8(trow + A1 + (uint32_t)(8 * g), hi);
      tc::tmem_st8(trow + A1 + (uint32_t)(C + 8 * g), lo);
    }
  }
  tc::tmem_wait_st();
  tc::fence_before();
  __syncthreads();
  if (tid == 0) {
    tc::fence_after();
    const uint32_t idesc = tc::idesc_tf32(eC, 0, 0);
    const uint32_t bh = tc::smem_addr(B1h), bl = tc::smem_addr(B1l);
    for (int s = 0; s < (C >> 3); ++s) {
      const uint64_t dh = tc::desc(bh + (uint32_t)s * (uint32_t)eC * 32u, 128u, 256u), dl = tc::desc(bl + (uint32_t)s * (uint32_t)eC * 32u, 128u, 256u);
      tc::mma_tf32_ts(tmem, tmem + A1 + (uint32_t)(C + 8 * s), dh, idesc, s > 0 ? 1u : 0u);
      tc::mma_tf32_ts(tmem, tmem + A1 + (uint32_t)(8 * s), dl, idesc, 1u);
      tc::mma_tf32_ts(tmem, tmem + A1 + (uint32_t)(8 * s), dh, idesc, 1u);
    }
    tc::commit(&bars[0]);
  }
  if (warp == 0) tc::mbar_wait(&bars[0], 0u);
  __syncthreads();
  tc::fence_after();

  // ---- stage 1: the hidden activation, hi over the accumulator, lo beside it
  {
    const float minv = P.mid_drop_p > 0.f ? 1.0f / (1.0f - P.mid_drop_p) : 1.f;
    const uint32_t mkey = P.mid_drop_p > 0.f ? rng_key(P.mid_seed + soff, P.mid_site) : 0u;
#pragma unroll 1
    for (int p = sh; p < (eC >> 3); p += 2) {
      float d[8], lo[8], hp[8];
      const size_t idx0 = ((size_t)b * eC + 8 * p) * S + v;
      if (bwd) {
#pragma unroll
        for (int j = 0; j < 8; ++j) hp[j] = live ? __ldg(P.hid_in + idx0 + (size_t)j * S) : 0.f;
      }
      tc::tmem_ld8(trow + (uint32_t)(8 * p), d);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float h;
        if (bwd) {
          h = d[j] * gelu_grad_f(hp[j]);
        } else {
          const float pre = d[j] + (P.ba ? __ldg(P.ba + 8 * p + j) : 0.f);
          if (live) P.hid_out[idx0 + (size_t)j * S] = pre;
          h = gelu_f(pre);
        }
        if (P.mid_drop_p > 0.f) h *= keep_from_bits(rng_word(mkey, idx0 + (size_t)j * S), P.mid_drop_p, minv);
        if (bwd && live) P.hid_out[idx0 + (size_t)j * S] = h;
        if (!live) h = 0.f;
        tc::split(h, d[j], lo[j]);
      }
      tc::tmem_st8(trow + (uint32_t)(8 * p), d);
      tc::tmem_st8(trow + (uint32_t)(eC + 8 * p), lo);
    }
  }
  tc::tmem_wait_st();
  tc::fence_before();
  __syncthreads();
  if (tid == 0) {
    tc::fence_after();
    const uint32_t idesc = tc::idesc_tf32(C, 0, 0);
    const uint32_t bh = tc::smem_addr(B2h), bl = tc::smem_addr(B2l);
    for (int s = 0; s < (eC >> 3); ++s) {
      const uint64_t dh = tc::desc(bh + (uint32_t)s * (uint32_t)C * 32u, 128u, 256u), dl = tc::desc(bl + (uint32_t)s * (uint32_t)C * 32u, 128u, 256u);
      tc::mma_tf32_ts(tmem + D2, tmem + (uint32_t)(eC + 8 * s), dh, idesc, s > 0 ? 1u : 0u);
      tc::mma_tf32_ts(tmem + D2, tmem + (uint32_t)(8 * s), dl, idesc, 1u);
      tc::mma_tf32_ts(tmem + D2, tmem + (uint32_t)(8 * s), dh, idesc, 1u);
    }
    tc::commit(&bars[0]);
  }
  if (warp == 0) tc::mbar_wait(&bars[0], 1u);
  __syncthreads();
  tc::fence_after();

  // ---- stage 2
  {
    const float oinv = P.out_drop_p > 0.f ? 1.0f / (1.0f - P.out_drop_p) : 1.f;
    const uint32_t okey = P.out_drop_p > 0.f ? rng_key(P.out_seed + soff, P.out_site) : 0u;
#pragma unroll 1
    for (int p = sh; p < (C >> 3); p += 2) {
      float d[8], r[8];
      const size_t idx0 = ((size_t)b * C + 8 * p) * S + v;
      if (P.res) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = live ? __ldg(P.res + idx0 + (size_t)j * S) : 0.f;
      }
      tc::tmem_ld8(trow + D2 + (uint32_t)(8 * p), d);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float y = d[j] + (P.bb ? __ldg(P.bb + 8 * p + j) : 0.f);
        if (P.out_drop_p > 0.f) y *= keep_from_bits(rng_word(okey, idx0 + (size_t)j * S), P.out_drop_p, oinv);
        if (P.res) y = fmaf(P.res_scale, r[j], y);
        if (live) P.out[idx0 + (size_t)j * S] = y;
      }
    }
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, ncols);
}

static int g_ffn_tc = 1;
void pw_ffn_tc_set(int on) { g_ffn_tc = on; }

static bool ffn_tc_shape_ok(int S, int C, int eC, const float* Wa, const float* Wb) {
  return S >= 1024 && (C & 15) == 0 && (eC & 15) == 0 && C <= 128 && eC <= 256 && 2 * eC + 2 * C <= 512 && !((uintptr_t)Wa & 15) &&
         !((uintptr_t)Wb & 15);
}

static int launch_ffn_tc(FfnTcBatch& T, double bytes, double flops, cudaStream_t stream) {
  size_t smem = 0;                     // hi / lo images of both weight matrices of the largest problem
  for (int i = 0; i < T.nprob; ++i) {
    const size_t s = sizeof(float) * 4 * (size_t)T.p[i].C * T.p[i].eC;
    smem = s > smem ? s : smem;
  }
  if (smem > 160 * 1024) return 1;
  T.seed_dev = get_seed_dev();
  prof_bytes(bytes);
  prof_flops(flops);
  VX_SET_SMEM(pw_ffn_tc_kernel, smem);
  VX_LAUNCH(pw_ffn_tc_kernel, dim3(cdiv(T.S, FT_ROWS), T.B, T.nprob), dim3(FT_THREADS), smem, stream, T);
  return check_launch("pw_ffn_tc_kernel");
}

// VX_OK when launched, 1 when the batch does not qualify (the caller runs the two contractions), negative on error
int pw_ffn_tc(const FfnBatch& batch, cudaStream_t stream) {
  if (!g_ffn_tc || precision_mode() != 0 || batch.nprob <= 0) return 1;
  FfnTcBatch T{};
  T.nprob = batch.nprob; T.B = batch.B; T.S = batch.S; T.bwd = 0;
  double bytes = 0.0, flops = 0.0;
  for (int i = 0; i < batch.nprob; ++i) {
    const FfnProblem& F = batch.p[i];
    if (!ffn_tc_shape_ok(batch.S, F.C, F.eC, F.W1, F.W2)) return 1;
    FfnTcProblem& P = T.p[i];
    P.in = F.x; P.C = F.C; P.pro_a = F.pro_a; P.pro_c = F.pro_c; P.pro_bstride = F.pro_bstride;
    P.Wa = F.W1; P.ba = F.b1; P.eC = F.eC; P.hid_out = F.hpre; P.hid_in = nullptr;
    P.mid_drop_p = F.mid_drop_p; P.mid_seed = F.mid_seed; P.mid_site = F.mid_site;
    P.Wb = F.W2; P.bb = F.b2; P.out_drop_p = F.drop_p; P.out_seed = F.seed; P.out_site = F.site;
    P.res = F.res; P.res_scale = F.res_scale; P.out = F.y;
    bytes += 4.0 * batch.B * batch.S * (2.0 * F.C + F.eC + (F.res ? F.C : 0)) + 8.0 * F.C * F.eC;
    flops += 4.0 * batch.B * batch.S * F.C * F.eC;
  }
  return launch_ffn_tc(T, bytes, flops, stream);
}

int pw_ffn_tc_bwd(const FfnBwdBatch& batch, cudaStream_t stream) {
  if (!g_ffn_tc || precision_mode() != 0 || batch.nprob <= 0) return 1;
  FfnTcBatch T{};
  T.nprob = batch.nprob; T.B = batch.B; T.S = batch.S; T.bwd = 1;
  double bytes = 0.0, flops = 0.0;
  for (int i = 0; i < batch.nprob; ++i) {
    const FfnBwdProblem& F = batch.p[i];
    if (!ffn_tc_shape_ok(batch.S, F.C, F.eC, F.W2, F.W1)) return 1;
    FfnTcProblem& P = T.p[i];
    P.in = F.dy; P.C = F.C; P.in_drop_p = F.out_drop_p; P.in_seed = F.out_seed; P.in_site = F.out_site;
    P.Wa = F.W2; P.eC = F.eC; P.hid_out = F.dh; P.hid_in = F.hpre;
    P.mid_drop_p = F.mid_drop_p; P.mid_seed = F.mid_seed; P.mid_site = F.mid_site;
    P.Wb = F.W1; P.out = F.dx;
    bytes += 4.0 * batch.B * batch.S * (2.0 * F.C + 2.0 * F.eC) + 8.0 * F.C * F.eC;
    flops += 4.0 * batch.B * batch.S * F.C * F.eC;
  }
  return launch_ffn_tc(T, bytes, flops, stream);
}

}  // namespace vx
