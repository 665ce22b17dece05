// JLC block (reference: model/components/conv_blocks.py:41-75) for sm_100a, fp32.
//
//   o = x + sum_{k in 1,3,5} GELU(IN(gconv_k(x) + b_k));      y = o + Dropout(W2 GELU(W1 IN(o) + b1) + b2)
//
// Kernels
//   jlc_conv_fwd_kernel   one shared 5^3-halo tile of x per (b, group, spatial tile) feeds all three branches: the
//                         k=3 / k=1 taps are sub-cubes of the k=5 footprint, so x is staged once and each tap row
//                         is read from shared memory once for up to three accumulator sets.  A thread owns VX
//                         consecutive voxels x 4 output channels x 3 branches in registers; weights are float4
//                         broadcasts.  Epilogue: bias, coalesced stores of the raw conv outputs, and per-tile
//                         (sum, sumsq) partials for the InstanceNorm that follows (deterministic two-level reduce).
//   jlc_combine_kernel    o = x + sum_k GELU(IN(z_k)), with partial stats of o.
//   jlc_conv_dgrad_kernel transposed (flipped-weight) correlation of the three branch gradients into dx.
//   jlc_conv_wgrad_kernel per-tile weight-gradient partials, thread = (ci, dz, dy) x 4 output channels with the
//                         5 (or 3, 1) dx taps in registers; folded with fp32 atomics.
//   jlc_bwd_abc_kernel    element-wise InstanceNorm / GELU backward: one CTA per (b, channel) row, three block reductions.
// The channel_conv FFN runs on the generic channel-contraction kernels in pointwise.cu.
#include "vx_kernels.h"

#ifdef VX_EMU
#define __grid_constant__
#endif

namespace vx {

struct ConvTile { int TZ, TY, TX, ntz, nty, ntx, VX, threads; };

static int g_force_vx = 0;
void jlc_force_vx(int vx) { g_force_vx = vx; }
static int g_force_tz[2] = {0, 0}, g_force_ty[2] = {0, 0};     // tuning probes (vx_set_option): [fwd/dgrad, wgrad]
void jlc_force_tile(int kind, int tz, int ty) { g_force_tz[kind] = tz; g_force_ty[kind] = ty; }
static int g_small_max_s = 512;     // volumes up to this many voxels use the small-volume conv kernels (0 disables)
void jlc_set_small_max(int s) { g_small_max_s = s; }
static bool jlc_use_small(int D, int H, int W) { return g_small_max_s > 0 && D * H * W <= g_small_max_s; }

// Tile choice, from a sweep on B200 at the per-level shapes of the three reference configs (tools/gpu_jlc_tiles.sh,
// profiles/r1_jlc_tile_sweep.txt).  What the sweep showed:
//  * fwd / dgrad: 4 voxels per thread (not 8) -- twice the threads for the same tile, ~90 instead of 128 registers;
//    the best tile is the largest that still gives the grid >= 96 CTAs, with whole warps per channel block (the stats
//    epilogue then uses warp reductions instead of per-thread shared atomics); below 32 threads a CTA is pure latency.
//  * wgrad (fixed 256 threads, tile-size-independent register use): it wants CTAs: the largest tile that still gives
//    >= 256 of them, never smaller than 4 (z, y) positions.
static ConvTile pick_tile(int B, int groups, int CG, int D, int H, int W, int max_threads, size_t max_smem_floats,
                          int smem_kind, int min_ctas = 96) {
  // smem_kind 0: fwd/dgrad (4 * halo tile), 1: wgrad (4 * halo tile + 3 * CG * tile)
  ConvTile best{};
  long long best_key = -1;
  const int VX = g_force_vx ? g_force_vx : 4;
  const int TX = ((W < 32 ? W : 32) + VX - 1) / VX * VX;
  const int cand[] = {8, 6, 4, 3, 2, 1};
  for (int tz : cand) {
    if (tz > D && tz != 1) continue;
    if (g_force_tz[smem_kind] && tz != g_force_tz[smem_kind]) continue;
    for (int ty : cand) {
      if (ty > H && ty != 1) continue;
      if (g_force_ty[smem_kind] && ty != g_force_ty[smem_kind]) continue;
      const int npos = tz * ty * (TX / VX);
      const int threads = npos * (CG / 4);
      if (smem_kind == 0 && (threads > max_threads || threads < 1)) continue;
      const int TXP = (TX + 4 + 3) & ~3;
      size_t fl = (size_t)4 * (tz + 4) * (ty + 4) * TXP;
      if (smem_kind == 1) fl += (size_t)3 * CG * tz * ty * TX;
      if (fl > max_smem_floats) continue;
      const int ntz = cdiv(D, tz), nty = cdiv(H, ty), ntx = cdiv(W, TX);
      const long long ncta = (long long)ntz * nty * ntx * groups * B;
      const int waste = (ntz * tz - D) * H + (nty * ty - H) * D;        // padded rows: tie-break towards exact tilings
      long long key;
      if (smem_kind == 0) {
        // min_ctas: 96 for the forward kernel (48 accumulators per thread: it wants the larger tile), 132 = 0.9 x 148 SMs for
        // the data gradient (16 accumulators: one CTA per SM, the reduction slices supply the threads) -- measured with
        // tools/gpu_r2_call17.sh (profiles/r2s_jlc_ks_sweep.txt): level 2 forward 64 us at 8x4 / KS 2 vs 76 us at 4x4 / KS 4,
        // data gradient 42 us at 4x4 / KS 4 vs 60 us at 8x4 / KS 2.
        const int enough_threads = threads >= 32, whole_warps = npos % 32 == 0, enough_ctas = ncta >= min_ctas;
        key = ((((long long)enough_threads * 2 + whole_warps) * 2 + enough_ctas) << 40) +
              ((long long)(enough_ctas ? threads : (ncta < 4096 ? ncta : 4096)) << 20) + (1 << 19) - waste * 64 + tz;
      } else {
        const int area = tz * ty, big_enough = area >= 4, enough_ctas = ncta >= 256;
        key = (((long long)big_enough * 2 + enough_ctas) << 40) +
              ((long long)(enough_ctas ? area : (ncta < 4096 ? ncta : 4096)) << 20) + (1 << 19) - waste * 64 + tz;
      }
      if (key > best_key) { best_key = key; best = ConvTile{tz, ty, TX, ntz, nty, ntx, VX, threads}; }
    }
  }
  return best;
}

// Stages `nch` channels of a zero-padded halo tile, layout [ch][HZ][HY][TXP], from NCDHW rows: tile origin (z0, y0, x0), halo
// PZ in z / y and XH (0 or 2) in x (TXP even).  16 lanes walk one row in 8-byte chunks, the (thread / 16) groups walk the
// (channel, z, y) rows with incremental indices: no per-element integer division -- the flat-index decoding this replaces
// was 14 % of the forward kernel's instructions at level 1 (profiles/r2o_jlc_conv_L1.source.txt).  Rows that are not
// 8-byte aligned in global memory (odd W or odd x0) take 4-byte copies.
VX_DEV void stage_halo_rows(float* xs, const float* __restrict__ src, size_t ch_stride, int nch, int z0, int y0, int x0, int PZ,
                            int HZ, int HY, int TXP, int D, int H, int W, int XH = 2) {
  const bool wide = ((W | x0 | TXP) & 1) == 0 && (((uintptr_t)src | (ch_stride * sizeof(float))) & 7) == 0;
  // lanes per row: the smallest power of two that covers a row's copies (16 for the 24-wide level-1 tiles, 8 for the 10-wide
  // padded rows of the small-volume kernels: with 16 there, 11 of 16 lanes idled and the kernels lost 15-40 %)
  const int ncopy = wide ? TXP >> 1 : TXP;
  int lpr = 32;
  while (lpr > 1 && (lpr >> 1) >= ncopy) lpr >>= 1;
  const int l = threadIdx.x % lpr, sub = threadIdx.x / lpr, nsub = blockDim.x / lpr;
  if (sub >= nsub) return;
  int hy = sub % HY, rz = (sub / HY) % HZ, ch = sub / (HY * HZ);      // three divisions per thread, none per element
  for (int r = sub; r < nch * HZ * HY; r += nsub) {
    const int gz = z0 + rz - PZ, gy = y0 + hy - PZ;
    const bool rowok = gz >= 0 && gz < D && gy >= 0 && gy < H;
    const float* grow = src + (size_t)ch * ch_stride + ((size_t)(rowok ? gz : 0) * H + (rowok ? gy : 0)) * W;
    float* srow = xs + (size_t)r * TXP;
    if (wide) {
      for (int c = l; 2 * c < TXP; c += lpr) {
        const int gx = x0 + 2 * c - XH;
        const bool ok = rowok && gx >= 0 && gx < W;       // W and gx even: the pair is inside or outside together
        vx_cp_async8(srow + 2 * c, ok ? grow + gx : src, ok);
      }
    } else {
      for (int c = l; c < TXP; c += lpr) {
        const int gx = x0 + c - XH;
        const bool ok = rowok && gx >= 0 && gx < W;
        vx_cp_async4(srow + c, ok ? grow + gx : src, ok);
      }
    }
    hy += nsub;
    while (hy >= HY) { hy -= HY; ++rz; }
    while (rz >= HZ) { rz -= HZ; ++ch; }
  }
}

struct ConvFwdArgs {
  const float* x; const float* w1; const float* b1; const float* w3; const float* b3; const float* w5; const float* b5;
  float* z;      // (3, B, C, S): branch k=1, 3, 5
  float* part;   // (3, B*C, ntiles, 2)
  int* cnt;      // (B*C) arrival counters of jlc_combine_kernel's chunks: zeroed here, by the first tile of every (b, group)
  int B, C, D, H, W;
  ConvTile t;
  int uniform_warps;
};

template <int CG, int VX>
__global__ void __launch_bounds__(256) jlc_conv_fwd_kernel(const __grid_constant__ ConvFwdArgs A) {
  VX_PDL_ENTRY();
  constexpr int NCB = CG / 4;
  const int g = blockIdx.y, b = blockIdx.z;
  const int TZ = A.t.TZ, TY = A.t.TY, TX = A.t.TX;
  const int tile = blockIdx.x;
  const int tx_i = tile % A.t.ntx, ty_i = (tile / A.t.ntx) % A.t.nty, tz_i = tile / (A.t.ntx * A.t.nty);
  const int z0 = tz_i * TZ, y0 = ty_i * TY, x0 = tx_i * TX;
  const int HZ = TZ + 4, HY = TY + 4, TXP = (TX + 4 + 3) & ~3;
  const int D = A.D, H = A.H, W = A.W, C = A.C;
  const size_t S = (size_t)D * H * W;

  VX_DYN_SMEM(float, sm);
  float* xs = sm;                                   // [4][HZ][HY][TXP]
  float* ws5 = xs + (size_t)4 * HZ * HY * TXP;      // [4][125][CG]
  float* ws3 = ws5 + 4 * 125 * CG;                  // [4][27][CG]
  float* ws1 = ws3 + 4 * 27 * CG;                   // [4][CG]
  float* sst = ws1 + 4 * CG;                        // [3][CG][2]

  const int tid = threadIdx.x, nthr = blockDim.x;
  const int NXQ = TX / VX;
  const int NPOS = TZ * TY * NXQ;
  const int pos = tid % NPOS, cb = tid / NPOS;
  const int xq = pos % NXQ, ty = (pos / NXQ) % TY, tz = pos / (NXQ * TY);

  float a5[VX][4], a3[VX][4], a1[VX][4];
#pragma unroll
  for (int v = 0; v < VX; ++v)
#pragma unroll
    for (int c = 0; c < 4; ++c) { a5[v][c] = 0.f; a3[v][c] = 0.f; a1[v][c] = 0.f; }

  for (int i = tid; i < 3 * CG * 2; i += nthr) sst[i] = 0.f;
  if (tile == 0 && tid < CG && A.cnt) A.cnt[b * C + g * CG + tid] = 0;

  for (int chunk = 0; chunk < NCB; ++chunk) {
    if (chunk > 0) __syncthreads();
    // stage 4 input channels of the group with a 2-voxel zero halo
    const int cin0 = g * CG + chunk * 4;
    stage_halo_rows(xs, A.x + ((size_t)b * C + cin0) * S, S, 4, z0, y0, x0, 2, HZ, HY, TXP, D, H, W);
    // weights of this ci-chunk, transposed to [ci][tap][co]
    for (int idx = tid; idx < CG * 4 * 125; idx += nthr) {
      const int tap = idx % 125, ci = (idx / 125) % 4, co = idx / 500;
      vx_cp_async4(ws5 + ((ci * 125 + tap) * CG + co), A.w5 + ((size_t)(g * CG + co) * CG + chunk * 4 + ci) * 125 + tap, true);
    }
    for (int idx = tid; idx < CG * 4 * 27; idx += nthr) {
      const int tap = idx % 27, ci = (idx / 27) % 4, co = idx / 108;
      vx_cp_async4(ws3 + ((ci * 27 + tap) * CG + co), A.w3 + ((size_t)(g * CG + co) * CG + chunk * 4 + ci) * 27 + tap, true);
    }
    for (int idx = tid; idx < CG * 4; idx += nthr) {
      const int ci = idx % 4, co = idx / 4;
      vx_cp_async4(ws1 + (ci * CG + co), A.w1 + (size_t)(g * CG + co) * CG + chunk * 4 + ci, true);
    }
    vx_cp_async_commit();
    vx_cp_async_wait_all();
    __syncthreads();

    if (cb < NCB) {
      // (ci, dz) stay rolled: the unrolled (dy, dx, v) body is ~20 KB of SASS and must stay instruction-cache resident
#pragma unroll 1
      for (int ci = 0; ci < 4; ++ci) {
#pragma unroll 1
        for (int dz = 0; dz < 5; ++dz) {
          const float* xrow = xs + ((size_t)(ci * HZ + tz + dz) * HY + ty) * TXP + xq * VX;
          const bool mid_z = dz >= 1 && dz <= 3;
#pragma unroll 1
          for (int dy = 0; dy < 5; ++dy) {
            float xr[VX + 4];
#pragma unroll
            for (int q = 0; q < (VX + 4) / 4; ++q) {
              const float4 t4 = *reinterpret_cast<const float4*>(xrow + dy * TXP + 4 * q);
              xr[4 * q] = t4.x; xr[4 * q + 1] = t4.y; xr[4 * q + 2] = t4.z; xr[4 * q + 3] = t4.w;
            }
            const float* w5p = ws5 + (size_t)(ci * 125 + (dz * 5 + dy) * 5) * CG + cb * 4;
#pragma unroll
            for (int dx = 0; dx < 5; ++dx) {
              const float4 w = *reinterpret_cast<const float4*>(w5p + dx * CG);
#pragma unroll
              for (int v = 0; v < VX; ++v) {
                a5[v][0] = fmaf(w.x, xr[v + dx], a5[v][0]); a5[v][1] = fmaf(w.y, xr[v + dx], a5[v][1]);
                a5[v][2] = fmaf(w.z, xr[v + dx], a5[v][2]); a5[v][3] = fmaf(w.w, xr[v + dx], a5[v][3]);
              }
            }
            if (dy >= 1 && dy <= 3) {
              if (mid_z) {
                const float* w3p = ws3 + (size_t)(ci * 27 + ((dz - 1) * 3 + (dy - 1)) * 3) * CG + cb * 4;
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                  const float4 w = *reinterpret_cast<const float4*>(w3p + dx * CG);
#pragma unroll
                  for (int v = 0; v < VX; ++v) {
                    a3[v][0] = fmaf(w.x, xr[v + 1 + dx], a3[v][0]); a3[v][1] = fmaf(w.y, xr[v + 1 + dx], a3[v][1]);
                    a3[v][2] = fmaf(w.z, xr[v + 1 + dx], a3[v][2]); a3[v][3] = fmaf(w.w, xr[v + 1 + dx], a3[v][3]);
                  }
                }
              }
              if (dy == 2 && dz == 2) {
                const float4 w = *reinterpret_cast<const float4*>(ws1 + ci * CG + cb * 4);
#pragma unroll
                for (int v = 0; v < VX; ++v) {
                  a1[v][0] = fmaf(w.x, xr[v + 2], a1[v][0]); a1[v][1] = fmaf(w.y, xr[v + 2], a1[v][1]);
                  a1[v][2] = fmaf(w.z, xr[v + 2], a1[v][2]); a1[v][3] = fmaf(w.w, xr[v + 2], a1[v][3]);
                }
              }
            }
          }
        }
      }
    }
  }

  // epilogue: bias, store, stats
  const int gz = z0 + tz, gy = y0 + ty, gx0 = x0 + xq * VX;
  const bool row_ok = cb < NCB && gz < D && gy < H;
  const size_t BCS = (size_t)A.B * C * S;
  const int lane = tid & 31;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int co = cb < NCB ? g * CG + cb * 4 + c : 0;
    const float bias[3] = {__ldg(A.b1 + co), __ldg(A.b3 + co), __ldg(A.b5 + co)};
    float s[3] = {0.f, 0.f, 0.f}, q[3] = {0.f, 0.f, 0.f};
    if (row_ok) {
      const size_t o = ((size_t)b * C + co) * S + ((size_t)gz * H + gy) * W + gx0;
#pragma unroll
      for (int v = 0; v < VX; ++v) {
        if (gx0 + v < W) {
          const float r1 = a1[v][c] + bias[0], r3 = a3[v][c] + bias[1], r5 = a5[v][c] + bias[2];
          A.z[o + v] = r1; A.z[BCS + o + v] = r3; A.z[2 * BCS + o + v] = r5;
          s[0] += r1; q[0] = fmaf(r1, r1, q[0]);
          s[1] += r3; q[1] = fmaf(r3, r3, q[1]);
          s[2] += r5; q[2] = fmaf(r5, r5, q[2]);
        }
      }
    }
    if (A.uniform_warps) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float ss = warp_sum(s[k]), qq = warp_sum(q[k]);
        if (lane == 0 && cb < NCB) {
          atomicAdd(sst + ((k * CG) + cb * 4 + c) * 2, ss);
          atomicAdd(sst + ((k * CG) + cb * 4 + c) * 2 + 1, qq);
        }
      }
    } else if (row_ok) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        atomicAdd(sst + ((k * CG) + cb * 4 + c) * 2, s[k]);
        atomicAdd(sst + ((k * CG) + cb * 4 + c) * 2 + 1, q[k]);
      }
    }
  }
  __syncthreads();
  const int ntiles = gridDim.x;
  for (int i = tid; i < 3 * CG; i += nthr) {
    const int k = i / CG, c = i % CG;
    const size_t row = (size_t)k * A.B * C + (size_t)b * C + g * CG + c;
    float* p = A.part + (row * ntiles + tile) * 2;
    p[0] = sst[i * 2]; p[1] = sst[i * 2 + 1];
  }
}

// ---------------------------------------------------------------------------------------------------
// Reduction-split variant of the tiled forward kernel (levels 1-2).  The tiled kernel above has exactly one thread per
// (4 voxels x 4 output channels): 1 728 warps at level 1 and 432 at level 2 for B = 4 -- 12 and 3 warps per SM walking a
// serial (ci, dz, dy) chain (ncu: 17 % / 9 % warps active, 56 % / 46 % issue-active, profiles/r2o_jlc_conv_L*.digest.txt).
// Here the whole group (all CG input channels and their weights) is staged once and the CG input channels are dealt to KS
// thread slices: KS times the warps, a chain KS times shorter, no re-staging per 4-channel chunk; the slices' 48
// accumulators meet in shared memory (fixed order: slice 0 adds slices 1 .. KS-1) before the unchanged epilogue.
// ---------------------------------------------------------------------------------------------------
// CTA size limit / CTAs per SM the register budget is cut for: c_g = 4 (level 1, 288 CTAs at B = 4) wants two CTAs of 384
// threads per SM (24 warps), c_g = 8 (level 2, 96 CTAs) one CTA of up to 576 threads.
template <int CG> struct JlcKs { static constexpr int MAXT = CG == 4 ? 384 : 576, MINB = CG == 4 ? 2 : 1; };

template <int CG, int VX>
__global__ void __launch_bounds__(JlcKs<CG>::MAXT, JlcKs<CG>::MINB) jlc_conv_fwd_ks_kernel(const __grid_constant__ ConvFwdArgs A, int KS) {
  VX_PDL_ENTRY();
  constexpr int NCB = CG / 4;
  const int g = blockIdx.y, b = blockIdx.z;
  const int TZ = A.t.TZ, TY = A.t.TY, TX = A.t.TX;
  const int tile = blockIdx.x;
  const int tx_i = tile % A.t.ntx, ty_i = (tile / A.t.ntx) % A.t.nty, tz_i = tile / (A.t.ntx * A.t.nty);
  const int z0 = tz_i * TZ, y0 = ty_i * TY, x0 = tx_i * TX;
  const int HZ = TZ + 4, HY = TY + 4, TXP = (TX + 4 + 3) & ~3;
  const int D = A.D, H = A.H, W = A.W, C = A.C;
  const size_t S = (size_t)D * H * W;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int NXQ = TX / VX;
  const int NPOS = TZ * TY * NXQ, NT = NPOS * NCB;
  const int kz = tid / NT, t2 = tid % NT;               // reduction slice, thread inside the slice
  const int pos = t2 % NPOS, cb = t2 / NPOS;
  const int xq = pos % NXQ, ty = (pos / NXQ) % TY, tz = pos / (NXQ * TY);
  const int XV = HZ * HY * TXP;                          // one staged channel

  VX_DYN_SMEM(float, sm);
  float* xs = sm;                                   // [CG][HZ][HY][TXP]
  float* ws5 = xs + (size_t)CG * XV;                // [CG ci][125][CG co]
  float* ws3 = ws5 + CG * 125 * CG;                 // [CG][27][CG]
  float* ws1 = ws3 + CG * 27 * CG;                  // [CG][CG]
  const size_t stage_fl = (size_t)CG * XV + (size_t)CG * 153 * CG, red_fl = (size_t)(KS - 1) * 12 * VX * NT;
  float* sst = sm + (stage_fl > red_fl ? stage_fl : red_fl);      // [3][CG][2]
  float* red = sm;                                  // [(KS - 1)][12 VX][NT]: reuses the staging area after the compute phase

  for (int i = tid; i < 3 * CG * 2; i += nthr) sst[i] = 0.f;
  if (tile == 0 && tid < CG && A.cnt) A.cnt[b * C + g * CG + tid] = 0;
  stage_halo_rows(xs, A.x + ((size_t)b * C + g * CG) * S, S, CG, z0, y0, x0, 2, HZ, HY, TXP, D, H, W);
  for (int idx = tid; idx < CG * CG * 125; idx += nthr) {      // weights transposed to [ci][tap][co]
    const int tap = idx % 125, ci = (idx / 125) % CG, co = idx / (125 * CG);
    vx_cp_async4(ws5 + ((ci * 125 + tap) * CG + co), A.w5 + ((size_t)(g * CG + co) * CG + ci) * 125 + tap, true);
  }
  for (int idx = tid; idx < CG * CG * 27; idx += nthr) {
    const int tap = idx % 27, ci = (idx / 27) % CG, co = idx / (27 * CG);
    vx_cp_async4(ws3 + ((ci * 27 + tap) * CG + co), A.w3 + ((size_t)(g * CG + co) * CG + ci) * 27 + tap, true);
  }
  for (int idx = tid; idx < CG * CG; idx += nthr) {
    const int ci = idx % CG, co = idx / CG;
    vx_cp_async4(ws1 + (ci * CG + co), A.w1 + (size_t)(g * CG + co) * CG + ci, true);
  }
  vx_cp_async_commit();
  vx_cp_async_wait_all();
  __syncthreads();

  float a5[VX][4], a3[VX][4], a1[VX][4];
#pragma unroll
  for (int v = 0; v < VX; ++v)
#pragma unroll
    for (int c = 0; c < 4; ++c) { a5[v][c] = 0.f; a3[v][c] = 0.f; a1[v][c] = 0.f; }

  if (kz < KS) {
    const int per = CG / KS;
    // (ci, dz, dy) stay rolled: the unrolled (dx, v) body must stay instruction-cache resident
#pragma unroll 1
    for (int ci = kz * per; ci < (kz + 1) * per; ++ci) {
#pragma unroll 1
      for (int dz = 0; dz < 5; ++dz) {
        const float* xrow = xs + ((size_t)(ci * HZ + tz + dz) * HY + ty) * TXP + xq * VX;
        const bool mid_z = dz >= 1 && dz <= 3;
#pragma unroll 1
        for (int dy = 0; dy < 5; ++dy) {
          float xr[VX + 4];
#pragma unroll
          for (int q = 0; q < (VX + 4) / 4; ++q) {
            const float4 t4 = *reinterpret_cast<const float4*>(xrow + dy * TXP + 4 * q);
            xr[4 * q] = t4.x; xr[4 * q + 1] = t4.y; xr[4 * q + 2] = t4.z; xr[4 * q + 3] = t4.w;
          }
          const float* w5p = ws5 + (size_t)(ci * 125 + (dz * 5 + dy) * 5) * CG + cb * 4;
#pragma unroll
          for (int dx = 0; dx < 5; ++dx) {
            const float4 w = *reinterpret_cast<const float4*>(w5p + dx * CG);
#pragma unroll
            for (int v = 0; v < VX; ++v) {
              a5[v][0] = fmaf(w.x, xr[v + dx], a5[v][0]); a5[v][1] = fmaf(w.y, xr[v + dx], a5[v][1]);
              a5[v][2] = fmaf(w.z, xr[v + dx], a5[v][2]); a5[v][3] = fmaf(w.w, xr[v + dx], a5[v][3]);
            }
          }
          if (dy >= 1 && dy <= 3) {
            if (mid_z) {
              const float* w3p = ws3 + (size_t)(ci * 27 + ((dz - 1) * 3 + (dy - 1)) * 3) * CG + cb * 4;
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
                const float4 w = *reinterpret_cast<const float4*>(w3p + dx * CG);
#pragma unroll
                for (int v = 0; v < VX; ++v) {
                  a3[v][0] = fmaf(w.x, xr[v + 1 + dx], a3[v][0]); a3[v][1] = fmaf(w.y, xr[v + 1 + dx], a3[v][1]);
                  a3[v][2] = fmaf(w.z, xr[v + 1 + dx], a3[v][2]); a3[v][3] = fmaf(w.w, xr[v + 1 + dx], a3[v][3]);
                }
              }
            }
            if (dy == 2 && dz == 2) {
              const float4 w = *reinterpret_cast<const float4*>(ws1 + ci * CG + cb * 4);
#pragma unroll
              for (int v = 0; v < VX; ++v) {
                a1[v][0] = fmaf(w.x, xr[v + 2], a1[v][0]); a1[v][1] = fmaf(w.y, xr[v + 2], a1[v][1]);
                a1[v][2] = fmaf(w.z, xr[v + 2], a1[v][2]); a1[v][3] = fmaf(w.w, xr[v + 2], a1[v][3]);
              }
            }
          }
        }
      }
    }
  }
  // ---- the slices meet: slice r > 0 parks its accumulators, slice 0 adds them in slice order
  __syncthreads();                                  // the staged tile and weights are dead from here on
  if (kz > 0 && kz < KS) {
    float* r = red + (size_t)(kz - 1) * 12 * VX * NT + t2;
#pragma unroll
    for (int v = 0; v < VX; ++v)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        r[(size_t)((0 * VX + v) * 4 + c) * NT] = a1[v][c];
        r[(size_t)((1 * VX + v) * 4 + c) * NT] = a3[v][c];
        r[(size_t)((2 * VX + v) * 4 + c) * NT] = a5[v][c];
      }
  }
  __syncthreads();
  const bool lead = kz == 0;
  if (lead) {
    for (int s2 = 1; s2 < KS; ++s2) {
      const float* r = red + (size_t)(s2 - 1) * 12 * VX * NT + t2;
#pragma unroll
      for (int v = 0; v < VX; ++v)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          a1[v][c] += r[(size_t)((0 * VX + v) * 4 + c) * NT];
          a3[v][c] += r[(size_t)((1 * VX + v) * 4 + c) * NT];
          a5[v][c] += r[(size_t)((2 * VX + v) * 4 + c) * NT];
        }
    }
  }

  // epilogue (slice 0): bias, store, stats
  const int gz = z0 + tz, gy = y0 + ty, gx0 = x0 + xq * VX;
  const bool row_ok = lead && gz < D && gy < H;
  const size_t BCS = (size_t)A.B * C * S;
  const int lane = tid & 31;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int co = g * CG + cb * 4 + c;
    const float bias[3] = {__ldg(A.b1 + co), __ldg(A.b3 + co), __ldg(A.b5 + co)};
    float s[3] = {0.f, 0.f, 0.f}, q[3] = {0.f, 0.f, 0.f};
    if (row_ok) {
      const size_t o = ((size_t)b * C + co) * S + ((size_t)gz * H + gy) * W + gx0;
#pragma unroll
      for (int v = 0; v < VX; ++v) {
        if (gx0 + v < W) {
          const float r1 = a1[v][c] + bias[0], r3 = a3[v][c] + bias[1], r5 = a5[v][c] + bias[2];
          A.z[o + v] = r1; A.z[BCS + o + v] = r3; A.z[2 * BCS + o + v] = r5;
          s[0] += r1; q[0] = fmaf(r1, r1, q[0]);
          s[1] += r3; q[1] = fmaf(r3, r3, q[1]);
          s[2] += r5; q[2] = fmaf(r5, r5, q[2]);
        }
      }
    }
    if (A.uniform_warps) {                          // NPOS % 32 == 0: a warp is uniform in (slice, channel block)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float ss = warp_sum(s[k]), qq = warp_sum(q[k]);
        if (lane == 0 && lead) {
          atomicAdd(sst + ((k * CG) + cb * 4 + c) * 2, ss);
          atomicAdd(sst + ((k * CG) + cb * 4 + c) * 2 + 1, qq);
        }
      }
    } else if (row_ok) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        atomicAdd(sst + ((k * CG) + cb * 4 + c) * 2, s[k]);
        atomicAdd(sst + ((k * CG) + cb * 4 + c) * 2 + 1, q[k]);
      }
    }
  }
  __syncthreads();
  const int ntiles = gridDim.x;
  for (int i = tid; i < 3 * CG; i += nthr) {
    const int k = i / CG, c = i % CG;
    const size_t row = (size_t)k * A.B * C + (size_t)b * C + g * CG + c;
    float* p = A.part + (row * ntiles + tile) * 2;
    p[0] = sst[i * 2]; p[1] = sst[i * 2 + 1];
  }
}

// ---------------------------------------------------------------------------------------------------
// Small-volume variants (levels 3-4: S = 216 / 27, Hecktor 256 / 32).  The tiled kernels above leave 32-128 CTAs of
// 36-64 threads walking a serial (ci, tap) loop -- 75-250 us of pure latency for a few MFLOP.  Here the whole padded
// volume of a group lives in shared memory and the work is split three ways: CTA = (batch, group, 4 output channels,
// voxel slab <= 128), thread = (voxel, slice of the reduction channels) -- warps are uniform in the slice, so every
// weight read is a 16-byte broadcast -- and the slices meet in shared memory before the bias / store / stats epilogue.
// ---------------------------------------------------------------------------------------------------
struct SmallGeo { int nsplit, VP, KS, Dp, Hp, Wp, PV; };

static int g_small_threads = 512;      // CTA size of the small-volume kernels (VX_OPT_JLC_SMALL_THREADS: 256 / 512 / 1024)
void jlc_set_small_threads(int t) { if (t == 256 || t == 512 || t == 1024) g_small_threads = t; }

static SmallGeo small_geo(int CG, int D, int H, int W) {
  SmallGeo g{};
  const int S = D * H * W;
  g.nsplit = (S + 127) / 128;
  const int Vc = (S + g.nsplit - 1) / g.nsplit;
  g.VP = (Vc + 31) & ~31;
  g.KS = g_small_threads / g.VP;         // reduction slices: the serial (channel, tap) walk of a thread shrinks with them
  if (g.KS > CG) g.KS = CG;
  while (CG % g.KS) --g.KS;
  g.Dp = D + 4; g.Hp = H + 4; g.Wp = W + 4;
  g.PV = g.Dp * g.Hp * g.Wp;
  return g;
}

template <int CG>
__global__ void __launch_bounds__(1024) jlc_conv_small_fwd_kernel(const __grid_constant__ ConvFwdArgs A,
                                                                 const __grid_constant__ SmallGeo G) {
  VX_PDL_ENTRY();
  constexpr int NQ = CG / 4;
  const int split = blockIdx.x, g = blockIdx.y / NQ, q = blockIdx.y % NQ, b = blockIdx.z;
  const int D = A.D, H = A.H, W = A.W, C = A.C;
  const int S = D * H * W;
  const int Hp = G.Hp, Wp = G.Wp, PV = G.PV, VP = G.VP, KS = G.KS;
  VX_DYN_SMEM(float, sm);
  float* xs = sm;                               // [CG][Dp][Hp][Wp], zero halo of 2
  float* ws = xs + (size_t)CG * PV;             // [CG][153][4]: taps 0..124 k=5, 125..151 k=3, 152 k=1
  float* red = ws + (size_t)CG * 153 * 4;       // [KS][12][VP]
  __shared__ float sst[12 * 4 * 2];             // (sum, sumsq) per (branch * 4 + c, warp of the voxel slab): fixed-order fold
  const int tid = threadIdx.x, nthr = blockDim.x;

  stage_halo_rows(xs, A.x + ((size_t)b * C + g * CG) * S, (size_t)S, CG, 0, 0, 0, 2, G.Dp, Hp, Wp, D, H, W);
  const int co0 = g * CG + q * 4;
  if (split == 0 && tid < 4 && A.cnt) A.cnt[b * C + co0 + tid] = 0;
  for (int idx = tid; idx < 4 * CG * 125; idx += nthr) {
    const int tap = idx % 125, ci = (idx / 125) % CG, c = idx / (125 * CG);
    vx_cp_async4(ws + ((ci * 153 + tap) * 4 + c), A.w5 + ((size_t)(co0 + c) * CG + ci) * 125 + tap, true);
  }
  for (int idx = tid; idx < 4 * CG * 27; idx += nthr) {
    const int tap = idx % 27, ci = (idx / 27) % CG, c = idx / (27 * CG);
    vx_cp_async4(ws + ((ci * 153 + 125 + tap) * 4 + c), A.w3 + ((size_t)(co0 + c) * CG + ci) * 27 + tap, true);
  }
  for (int idx = tid; idx < 4 * CG; idx += nthr) {
    const int ci = idx % CG, c = idx / CG;
    vx_cp_async4(ws + ((ci * 153 + 152) * 4 + c), A.w1 + (size_t)(co0 + c) * CG + ci, true);
  }
  vx_cp_async_commit();
  vx_cp_async_wait_all();
  __syncthreads();

  const int vl = tid % VP, sl = tid / VP;
  const int Vc = (S + G.nsplit - 1) / G.nsplit;
  const int vox = split * Vc + vl;
  const bool live = sl < KS && vl < Vc && vox < S;
  float a5[4] = {0.f, 0.f, 0.f, 0.f}, a3[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
  if (live) {
    const int x = vox % W, y = (vox / W) % H, z = vox / (W * H);
    const int per = CG / KS;
#pragma unroll 1
    for (int ci = sl * per; ci < (sl + 1) * per; ++ci) {
      const float* xb = xs + (size_t)ci * PV + ((size_t)z * Hp + y) * Wp + x;
      const float* wb = ws + (size_t)ci * 153 * 4;
#pragma unroll 1
      for (int dz = 0; dz < 5; ++dz) {
        const bool mid_z = dz >= 1 && dz <= 3;
#pragma unroll 1
        for (int dy = 0; dy < 5; ++dy) {
          const float* xr = xb + ((size_t)dz * Hp + dy) * Wp;
          float xv[5];
#pragma unroll
          for (int dx = 0; dx < 5; ++dx) xv[dx] = xr[dx];
          const float* w5p = wb + (size_t)((dz * 5 + dy) * 5) * 4;
#pragma unroll
          for (int dx = 0; dx < 5; ++dx) {
            const float4 w = *reinterpret_cast<const float4*>(w5p + dx * 4);
            a5[0] = fmaf(w.x, xv[dx], a5[0]); a5[1] = fmaf(w.y, xv[dx], a5[1]);
            a5[2] = fmaf(w.z, xv[dx], a5[2]); a5[3] = fmaf(w.w, xv[dx], a5[3]);
          }
          if (mid_z && dy >= 1 && dy <= 3) {
            const float* w3p = wb + (size_t)(125 + ((dz - 1) * 3 + (dy - 1)) * 3) * 4;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              const float4 w = *reinterpret_cast<const float4*>(w3p + dx * 4);
              a3[0] = fmaf(w.x, xv[dx + 1], a3[0]); a3[1] = fmaf(w.y, xv[dx + 1], a3[1]);
              a3[2] = fmaf(w.z, xv[dx + 1], a3[2]); a3[3] = fmaf(w.w, xv[dx + 1], a3[3]);
            }
            if (dz == 2 && dy == 2) {
              const float4 w = *reinterpret_cast<const float4*>(wb + 152 * 4);
              a1[0] = fmaf(w.x, xv[2], a1[0]); a1[1] = fmaf(w.y, xv[2], a1[1]);
              a1[2] = fmaf(w.z, xv[2], a1[2]); a1[3] = fmaf(w.w, xv[2], a1[3]);
            }
          }
        }
      }
    }
  }
  if (sl < KS) {
    float* r = red + (size_t)sl * 12 * VP + vl;
#pragma unroll
    for (int c = 0; c < 4; ++c) { r[(size_t)c * VP] = a1[c]; r[(size_t)(4 + c) * VP] = a3[c]; r[(size_t)(8 + c) * VP] = a5[c]; }
  }
  __syncthreads();
  // epilogue: item = (j = branch * 4 + c, voxel); VP is a multiple of 32, so a warp shares j
  const size_t BCS = (size_t)A.B * C * S;
  const int lane = tid & 31;
  for (int it0 = 0; it0 < 12 * VP; it0 += nthr) {
    const int it = it0 + tid;
    const bool in = it < 12 * VP;
    const int j = in ? it / VP : 0, v = in ? it % VP : 0;
    const int gv = split * Vc + v;
    const int k = j >> 2, co = co0 + (j & 3);
    float val = 0.f;
    const bool ok = in && v < Vc && gv < S;
    if (ok) {
      for (int s2 = 0; s2 < KS; ++s2) val += red[((size_t)s2 * 12 + j) * VP + v];
      val += __ldg((k == 0 ? A.b1 : (k == 1 ? A.b3 : A.b5)) + co);
      A.z[(size_t)k * BCS + ((size_t)b * C + co) * S + gv] = val;
    }
    const float ss = warp_sum(val), qq = warp_sum(val * val);
    if (lane == 0 && in) { sst[(j * 4 + (v >> 5)) * 2] = ss; sst[(j * 4 + (v >> 5)) * 2 + 1] = qq; }
  }
  __syncthreads();
  if (tid < 12) {
    const int k = tid >> 2, co = co0 + (tid & 3);
    const size_t row = (size_t)k * A.B * C + (size_t)b * C + co;
    float ss = 0.f, qq = 0.f;
    for (int w = 0; w < (VP >> 5); ++w) { ss += sst[(tid * 4 + w) * 2]; qq += sst[(tid * 4 + w) * 2 + 1]; }
    float* p = A.part + (row * G.nsplit + split) * 2;
    p[0] = ss; p[1] = qq;
  }
}

// o = x + sum_k GELU((z_k - mean_k) * rstd_k) with both InstanceNorm statistics passes folded in (no finalize launches):
//  * every CTA (row, chunk) folds the convolution's per-tile (sum, sumsq) partials of its row into (mean, rstd) itself
//    (ntiles <= a few dozen numbers, fixed order); chunk 0 records them in `stats` for the backward pass;
//  * the CTA that arrives last at the row's counter (zeroed by the convolution kernel) folds the chunks' partials of o, again
//    in fixed order, into stats[3 rows + row] and the (a, c) affine the FFN's first contraction applies as its prologue.
__global__ void __launch_bounds__(256) jlc_combine_kernel(const float* __restrict__ x, const float* __restrict__ z,
                                                          const float* __restrict__ part_z, int ntiles, float* __restrict__ stats,
                                                          float* __restrict__ o, float* __restrict__ part_o, int* __restrict__ cnt,
                                                          float* __restrict__ aff_a, float* __restrict__ aff_c, int rows, int S,
                                                          int chunk, float eps) {
  VX_PDL_ENTRY();
  __shared__ float red[33];
  __shared__ float mr[6];
  const int row = blockIdx.y, ck = blockIdx.x, nchunk = gridDim.x;
  const size_t RS = (size_t)rows * S;
  if (threadIdx.x < 3) {
    float mean, rstd;
    finalize_stats(part_z, threadIdx.x * rows + row, ntiles, (float)S, eps, mean, rstd);
    mr[2 * threadIdx.x] = mean; mr[2 * threadIdx.x + 1] = rstd;
    if (ck == 0) { stats[2 * (threadIdx.x * rows + row)] = mean; stats[2 * (threadIdx.x * rows + row) + 1] = rstd; }
  }
  __syncthreads();
  float m[3], r[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { m[k] = mr[2 * k]; r[k] = mr[2 * k + 1]; }
  const size_t base = (size_t)row * S;
  const int lo = ck * chunk, hi = min(S, lo + chunk);
  float s = 0.f, q = 0.f;
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    float v = x[base + i];
#pragma unroll
    for (int k = 0; k < 3; ++k) v += gelu_f((z[k * RS + base + i] - m[k]) * r[k]);
    o[base + i] = v;
    s += v; q = fmaf(v, v, q);
  }
  s = block_sum(s, red);
  q = block_sum(q, red);
  if (threadIdx.x == 0) {
    float* p = part_o + ((size_t)row * nchunk + ck) * 2;
    p[0] = s; p[1] = q;
    bool last = true;
    if (nchunk > 1) {
      __threadfence();                                   // the partial is visible before the arrival is
      last = atomicAdd(cnt + row, 1) == nchunk - 1;
      if (last) __threadfence();
    }
    if (last) {
      float mean, rstd;
      finalize_stats(part_o, row, nchunk, (float)S, eps, mean, rstd);
      stats[2 * (3 * rows + row)] = mean; stats[2 * (3 * rows + row) + 1] = rstd;
      aff_a[row] = rstd; aff_c[row] = -mean * rstd;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// backward element-wise stages
// ---------------------------------------------------------------------------------------------------
// a + b + c in one launch (round 2): a CTA owns one whole (b, channel) row, so the three dependent row reductions
// (sum dohat / dohat ohat -> dO and the six branch sums -> gz) are block reductions in a fixed order instead of three launches
// with global atomics in between.  The row is re-read from L1 / L2 between the phases (the element-wise work is a few hundred
// KB per CTA); the sums no longer depend on the arrival order of atomics, i.e. the backward is run-to-run deterministic here.
template <int N>
VX_DEV void block_sum_n(float (&v)[N], float* red /* [N][33] */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < N; ++k) v[k] = warp_sum(v[k]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) red[k * 33 + w] = v[k];
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      float t = lane < nw ? red[k * 33 + lane] : 0.f;
      t = warp_sum(t);
      if (lane == 0) red[k * 33 + 32] = t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) v[k] = red[k * 33 + 32];
}

__global__ void __launch_bounds__(1024) jlc_bwd_abc_kernel(const float* __restrict__ dy, const float* __restrict__ dohat,
                                                           const float* __restrict__ o, const float* __restrict__ z,
                                                           const float* __restrict__ stats, float* __restrict__ dO,
                                                           float* __restrict__ gz, int rows, int S) {
  VX_PDL_ENTRY();
  __shared__ float red[6 * 33];
  const int row = blockIdx.x;
  const size_t RS = (size_t)rows * S, base = (size_t)row * S;
  const float mo = stats[2 * (3 * rows + row)], ro = stats[2 * (3 * rows + row) + 1];
  float m[3], r[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { m[k] = stats[2 * (k * rows + row)]; r[k] = stats[2 * (k * rows + row) + 1]; }
  const float invS = 1.0f / (float)S;
  // a: (sum dohat, sum dohat * ohat)
  float a[2] = {0.f, 0.f};
#pragma unroll 4
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    const float g = dohat[base + i];
    a[0] += g; a[1] = fmaf(g, (o[base + i] - mo) * ro, a[1]);
  }
  block_sum_n<2>(a, red);
  const float m1 = a[0] * invS, m2 = a[1] * invS;
  // b: dO = dy + rstd_o (dohat - m1 - ohat m2);  (sum g_k, sum g_k zhat_k), g_k = dO * GELU'(zhat_k)
  float sb[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    const float oh = (o[base + i] - mo) * ro;
    const float d = dy[base + i] + ro * (dohat[base + i] - m1 - oh * m2);
    dO[base + i] = d;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float zh = (z[k * RS + base + i] - m[k]) * r[k];
      const float g = d * gelu_grad_f(zh);
      sb[2 * k] += g; sb[2 * k + 1] = fmaf(g, zh, sb[2 * k + 1]);
    }
  }
  block_sum_n<6>(sb, red);
  // c: gz_k = rstd_k (g_k - mean(g_k) - zhat_k mean(g_k zhat_k)); every thread re-reads the dO values it wrote itself
#pragma unroll 2
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    const float d = dO[base + i];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float zh = (z[k * RS + base + i] - m[k]) * r[k];
      const float g = d * gelu_grad_f(zh);
      gz[k * RS + base + i] = r[k] * (g - sb[2 * k] * invS - zh * sb[2 * k + 1] * invS);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// conv dgrad:  dx[ci] = dO[ci] + sum_k sum_co sum_t w_k[co][ci][K-1-t] gz_k[co][u + t - p]
// The three branches are processed one after the other through the same shared tile (halo 2, 1, 0).
// ---------------------------------------------------------------------------------------------------
struct ConvDgradArgs {
  const float* gz;   // (3, B, C, S)
  const float* dO;   // (B, C, S)
  const float* w1; const float* w3; const float* w5;
  float* dx;
  int B, C, D, H, W;
  ConvTile t;
};

template <int CG, int VX, int K>
VX_DEV void dgrad_branch(const ConvDgradArgs& A, const float* __restrict__ gzk, const float* __restrict__ wk, float* xs,
                         float* ws, float (&acc)[VX][4], int g, int b, int z0, int y0, int x0, int tz, int ty, int xq,
                         int cb) {
  constexpr int NCB = CG / 4, P = K / 2, K3 = K * K * K;
  const int TZ = A.t.TZ, TY = A.t.TY, TX = A.t.TX;
  const int HZ = TZ + 2 * P, HY = TY + 2 * P, TXP = (TX + 4 + 3) & ~3;   // x keeps the 2-voxel halo for alignment
  const int D = A.D, H = A.H, W = A.W, C = A.C;
  const size_t S = (size_t)D * H * W;
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int chunk = 0; chunk < NCB; ++chunk) {
    __syncthreads();
    const int c0 = g * CG + chunk * 4;     // 4 "input" channels of this correlation = conv output channels
    stage_halo_rows(xs, gzk + ((size_t)b * C + c0) * S, S, 4, z0, y0, x0, P, HZ, HY, TXP, D, H, W);
    // ws[co_local(4)][tap'][ci_out(CG)] = w[co][ci][K3-1-tap']
    for (int idx = tid; idx < 4 * CG * K3; idx += nthr) {
      const int tap = idx % K3, cio = (idx / K3) % CG, col = idx / (K3 * CG);
      vx_cp_async4(ws + ((col * K3 + (K3 - 1 - tap)) * CG + cio), wk + ((size_t)(c0 + col) * CG + cio) * K3 + tap, true);
    }
    vx_cp_async_commit();
    vx_cp_async_wait_all();
    __syncthreads();
    if (cb < NCB) {
#pragma unroll 1
      for (int ci = 0; ci < 4; ++ci) {
#pragma unroll 1
        for (int dz = 0; dz < K; ++dz) {
          const float* xrow = xs + ((size_t)(ci * HZ + tz + dz) * HY + ty) * TXP + xq * VX;
#pragma unroll 1
          for (int dy = 0; dy < K; ++dy) {
            float xr[VX + 4];
#pragma unroll
            for (int q = 0; q < (VX + 4) / 4; ++q) {
              const float4 t4 = *reinterpret_cast<const float4*>(xrow + dy * TXP + 4 * q);
              xr[4 * q] = t4.x; xr[4 * q + 1] = t4.y; xr[4 * q + 2] = t4.z; xr[4 * q + 3] = t4.w;
            }
            const float* wp = ws + (size_t)(ci * K3 + (dz * K + dy) * K) * CG + cb * 4;
#pragma unroll
            for (int dx = 0; dx < K; ++dx) {
              const float4 w = *reinterpret_cast<const float4*>(wp + dx * CG);
#pragma unroll
              for (int v = 0; v < VX; ++v) {
                const float xv = xr[v + dx + 2 - P];
                acc[v][0] = fmaf(w.x, xv, acc[v][0]); acc[v][1] = fmaf(w.y, xv, acc[v][1]);
                acc[v][2] = fmaf(w.z, xv, acc[v][2]); acc[v][3] = fmaf(w.w, xv, acc[v][3]);
              }
            }
          }
        }
      }
    }
  }
}

template <int CG, int VX>
__global__ void __launch_bounds__(256) jlc_conv_dgrad_kernel(const __grid_constant__ ConvDgradArgs A) {
  VX_PDL_ENTRY();
  constexpr int NCB = CG / 4;
  const int g = blockIdx.y, b = blockIdx.z;
  const int TZ = A.t.TZ, TY = A.t.TY, TX = A.t.TX;
  const int tile = blockIdx.x;
  const int tx_i = tile % A.t.ntx, ty_i = (tile / A.t.ntx) % A.t.nty, tz_i = tile / (A.t.ntx * A.t.nty);
  const int z0 = tz_i * TZ, y0 = ty_i * TY, x0 = tx_i * TX;
  const int TXP = (TX + 4 + 3) & ~3;
  VX_DYN_SMEM(float, sm);
  float* xs = sm;
  float* ws = sm + (size_t)4 * (TZ + 4) * (TY + 4) * TXP;   // [4][125][CG] (largest branch)
  const int tid = threadIdx.x;
  const int NXQ = TX / VX, NPOS = TZ * TY * NXQ;
  const int pos = tid % NPOS, cb = tid / NPOS;
  const int xq = pos % NXQ, ty = (pos / NXQ) % TY, tz = pos / (NXQ * TY);
  const size_t S = (size_t)A.D * A.H * A.W;
  const size_t BCS = (size_t)A.B * A.C * S;

  float acc[VX][4];
#pragma unroll
  for (int v = 0; v < VX; ++v)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[v][c] = 0.f;

  dgrad_branch<CG, VX, 5>(A, A.gz + 2 * BCS, A.w5, xs, ws, acc, g, b, z0, y0, x0, tz, ty, xq, cb);
  dgrad_branch<CG, VX, 3>(A, A.gz + BCS, A.w3, xs, ws, acc, g, b, z0, y0, x0, tz, ty, xq, cb);
  dgrad_branch<CG, VX, 1>(A, A.gz, A.w1, xs, ws, acc, g, b, z0, y0, x0, tz, ty, xq, cb);

  const int gz = z0 + tz, gy = y0 + ty, gx0 = x0 + xq * VX;
  if (cb < NCB && gz < A.D && gy < A.H) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int ci = g * CG + cb * 4 + c;
      const size_t o = ((size_t)b * A.C + ci) * S + ((size_t)gz * A.H + gy) * A.W + gx0;
      // every residual load before the first store (dx and dO may alias as far as the compiler knows: interleaved, each load
      // waited for the store before it)
      float r[VX];
#pragma unroll
      for (int v = 0; v < VX; ++v) r[v] = gx0 + v < A.W ? A.dO[o + v] : 0.f;
#pragma unroll
      for (int v = 0; v < VX; ++v)
        if (gx0 + v < A.W) A.dx[o + v] = acc[v][c] + r[v];
    }
  }
}

// Reduction-split data gradient (same idea as jlc_conv_fwd_ks_kernel): per branch the gradients of ALL CG output channels
// of the group and the flipped weights are staged once (three staging phases instead of 3 * CG / 4) and the CG reduction
// channels are dealt to KS thread slices; the 4 VX accumulators of the slices meet in shared memory.
template <int CG, int VX, int K>
VX_DEV void dgrad_branch_ks(const ConvDgradArgs& A, const float* __restrict__ gzk, const float* __restrict__ wk, float* xs,
                            float* ws, float (&acc)[VX][4], int g, int b, int z0, int y0, int x0, int tz, int ty, int xq,
                            int cb, int kz, int KS) {
  constexpr int P = K / 2, K3 = K * K * K;
  const int TZ = A.t.TZ, TY = A.t.TY, TX = A.t.TX;
  const int HZ = TZ + 2 * P, HY = TY + 2 * P, TXP = (TX + 4 + 3) & ~3;   // x keeps the 2-voxel halo for alignment
  const int D = A.D, H = A.H, W = A.W, C = A.C;
  const size_t S = (size_t)D * H * W;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int XV = HZ * HY * TXP;
  __syncthreads();                                   // the previous branch has been consumed
  stage_halo_rows(xs, gzk + ((size_t)b * C + g * CG) * S, S, CG, z0, y0, x0, P, HZ, HY, TXP, D, H, W);
  // ws[co][tap'][ci] = w[co][ci][K3 - 1 - tap']
  for (int idx = tid; idx < CG * CG * K3; idx += nthr) {
    const int tap = idx % K3, cio = (idx / K3) % CG, col = idx / (K3 * CG);
    vx_cp_async4(ws + ((col * K3 + (K3 - 1 - tap)) * CG + cio), wk + ((size_t)(g * CG + col) * CG + cio) * K3 + tap, true);
  }
  vx_cp_async_commit();
  vx_cp_async_wait_all();
  __syncthreads();
  if (kz < KS) {
    const int per = CG / KS;
#pragma unroll 1
    for (int co = kz * per; co < (kz + 1) * per; ++co) {
#pragma unroll 1
      for (int dz = 0; dz < K; ++dz) {
        const float* xrow = xs + ((size_t)(co * HZ + tz + dz) * HY + ty) * TXP + xq * VX;
#pragma unroll 1
        for (int dy = 0; dy < K; ++dy) {
          float xr[VX + 4];
#pragma unroll
          for (int q = 0; q < (VX + 4) / 4; ++q) {
            const float4 t4 = *reinterpret_cast<const float4*>(xrow + dy * TXP + 4 * q);
            xr[4 * q] = t4.x; xr[4 * q + 1] = t4.y; xr[4 * q + 2] = t4.z; xr[4 * q + 3] = t4.w;
          }
          const float* wp = ws + (size_t)(co * K3 + (dz * K + dy) * K) * CG + cb * 4;
#pragma unroll
          for (int dx = 0; dx < K; ++dx) {
            const float4 w = *reinterpret_cast<const float4*>(wp + dx * CG);
#pragma unroll
            for (int v = 0; v < VX; ++v) {
              const float xv = xr[v + dx + 2 - P];
              acc[v][0] = fmaf(w.x, xv, acc[v][0]); acc[v][1] = fmaf(w.y, xv, acc[v][1]);
              acc[v][2] = fmaf(w.z, xv, acc[v][2]); acc[v][3] = fmaf(w.w, xv, acc[v][3]);
            }
          }
        }
      }
    }
  }
}

template <int CG, int VX>
__global__ void __launch_bounds__(JlcKs<CG>::MAXT, JlcKs<CG>::MINB) jlc_conv_dgrad_ks_kernel(const __grid_constant__ ConvDgradArgs A, int KS) {
  VX_PDL_ENTRY();
  constexpr int NCB = CG / 4;
  const int g = blockIdx.y, b = blockIdx.z;
  const int TZ = A.t.TZ, TY = A.t.TY, TX = A.t.TX;
  const int tile = blockIdx.x;
  const int tx_i = tile % A.t.ntx, ty_i = (tile / A.t.ntx) % A.t.nty, tz_i = tile / (A.t.ntx * A.t.nty);
  const int z0 = tz_i * TZ, y0 = ty_i * TY, x0 = tx_i * TX;
  const int TXP = (TX + 4 + 3) & ~3;
  VX_DYN_SMEM(float, sm);
  float* xs = sm;                                            // [CG][TZ + 4][TY + 4][TXP] (largest branch)
  float* ws = sm + (size_t)CG * (TZ + 4) * (TY + 4) * TXP;   // [CG][125][CG]
  const int tid = threadIdx.x;
  const int NXQ = TX / VX, NPOS = TZ * TY * NXQ, NT = NPOS * NCB;
  const int kz = tid / NT, t2 = tid % NT;
  const int pos = t2 % NPOS, cb = t2 / NPOS;
  const int xq = pos % NXQ, ty = (pos / NXQ) % TY, tz = pos / (NXQ * TY);
  const size_t S = (size_t)A.D * A.H * A.W;
  const size_t BCS = (size_t)A.B * A.C * S;

  float acc[VX][4];
#pragma unroll
  for (int v = 0; v < VX; ++v)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[v][c] = 0.f;

  dgrad_branch_ks<CG, VX, 5>(A, A.gz + 2 * BCS, A.w5, xs, ws, acc, g, b, z0, y0, x0, tz, ty, xq, cb, kz, KS);
  dgrad_branch_ks<CG, VX, 3>(A, A.gz + BCS, A.w3, xs, ws, acc, g, b, z0, y0, x0, tz, ty, xq, cb, kz, KS);
  dgrad_branch_ks<CG, VX, 1>(A, A.gz, A.w1, xs, ws, acc, g, b, z0, y0, x0, tz, ty, xq, cb, kz, KS);

  __syncthreads();
  float* red = sm;                                           // [(KS - 1)][4 VX][NT]
  if (kz > 0 && kz < KS) {
    float* r = red + (size_t)(kz - 1) * 4 * VX * NT + t2;
#pragma unroll
    for (int v = 0; v < VX; ++v)
#pragma unroll
      for (int c = 0; c < 4; ++c) r[(size_t)(v * 4 + c) * NT] = acc[v][c];
  }
  __syncthreads();
  if (kz != 0) return;
  for (int s2 = 1; s2 < KS; ++s2) {
    const float* r = red + (size_t)(s2 - 1) * 4 * VX * NT + t2;
#pragma unroll
    for (int v = 0; v < VX; ++v)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[v][c] += r[(size_t)(v * 4 + c) * NT];
  }
  const int gz = z0 + tz, gy = y0 + ty, gx0 = x0 + xq * VX;
  if (gz < A.D && gy < A.H) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int ci = g * CG + cb * 4 + c;
      const size_t o = ((size_t)b * A.C + ci) * S + ((size_t)gz * A.H + gy) * A.W + gx0;
      // every residual load before the first store (dx and dO may alias as far as the compiler knows: interleaved, each load
      // waited for the store before it)
      float r[VX];
#pragma unroll
      for (int v = 0; v < VX; ++v) r[v] = gx0 + v < A.W ? A.dO[o + v] : 0.f;
#pragma unroll
      for (int v = 0; v < VX; ++v)
        if (gx0 + v < A.W) A.dx[o + v] = acc[v][c] + r[v];
    }
  }
}

// Small-volume dgrad (same decomposition as jlc_conv_small_fwd_kernel): CTA = (batch, group, 4 input channels, voxel
// slab), thread = (voxel, slice of the group's output channels); the three branch gradients sit in shared memory as
// padded volumes (k=5 and k=3 share the halo-2 layout, k=1 is read at the centre), weights are staged flipped so the
// loop is a plain correlation.
template <int CG>
__global__ void __launch_bounds__(1024) jlc_conv_small_dgrad_kernel(const __grid_constant__ ConvDgradArgs A,
                                                                   const __grid_constant__ SmallGeo G) {
  VX_PDL_ENTRY();
  constexpr int NQ = CG / 4;
  const int split = blockIdx.x, g = blockIdx.y / NQ, q = blockIdx.y % NQ, b = blockIdx.z;
  const int D = A.D, H = A.H, W = A.W, C = A.C;
  const int S = D * H * W;
  const int Hp = G.Hp, Wp = G.Wp, PV = G.PV, VP = G.VP, KS = G.KS;
  const size_t BCS = (size_t)A.B * C * S;
  VX_DYN_SMEM(float, sm);
  float* gs = sm;                               // [3][CG][Dp][Hp][Wp]  (k = 1, 3, 5 like A.gz)
  float* ws = gs + (size_t)3 * CG * PV;         // [CG co][153][4 ci]: flipped taps, 0..124 k=5, 125..151 k=3, 152 k=1
  float* red = ws + (size_t)CG * 153 * 4;       // [KS][4][VP]
  const int tid = threadIdx.x, nthr = blockDim.x;

  for (int k = 0; k < 3; ++k)
    stage_halo_rows(gs + (size_t)k * CG * PV, A.gz + (size_t)k * BCS + ((size_t)b * C + g * CG) * S, (size_t)S, CG, 0, 0, 0, 2, G.Dp, Hp, Wp,
                    D, H, W);
  const int ci0 = q * 4;                         // input channels (within the group) this CTA produces
  for (int idx = tid; idx < CG * 4 * 125; idx += nthr) {
    const int tap = idx % 125, c = (idx / 125) % 4, co = idx / 500;
    vx_cp_async4(ws + ((co * 153 + (124 - tap)) * 4 + c), A.w5 + ((size_t)(g * CG + co) * CG + ci0 + c) * 125 + tap, true);
  }
  for (int idx = tid; idx < CG * 4 * 27; idx += nthr) {
    const int tap = idx % 27, c = (idx / 27) % 4, co = idx / 108;
    vx_cp_async4(ws + ((co * 153 + 125 + (26 - tap)) * 4 + c), A.w3 + ((size_t)(g * CG + co) * CG + ci0 + c) * 27 + tap, true);
  }
  for (int idx = tid; idx < CG * 4; idx += nthr) {
    const int c = idx % 4, co = idx / 4;
    vx_cp_async4(ws + ((co * 153 + 152) * 4 + c), A.w1 + (size_t)(g * CG + co) * CG + ci0 + c, true);
  }
  vx_cp_async_commit();
  vx_cp_async_wait_all();
  __syncthreads();

  const int vl = tid % VP, sl = tid / VP;
  const int Vc = (S + G.nsplit - 1) / G.nsplit;
  const int vox = split * Vc + vl;
  const bool live = sl < KS && vl < Vc && vox < S;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (live) {
    const int x = vox % W, y = (vox / W) % H, z = vox / (W * H);
    const int per = CG / KS;
#pragma unroll 1
    for (int co = sl * per; co < (sl + 1) * per; ++co) {
      const size_t vo = ((size_t)z * Hp + y) * Wp + x;
      const float* g1 = gs + (size_t)co * PV + vo;
      const float* g3 = gs + (size_t)(CG + co) * PV + vo;
      const float* g5 = gs + (size_t)(2 * CG + co) * PV + vo;
      const float* wb = ws + (size_t)co * 153 * 4;
#pragma unroll 1
      for (int dz = 0; dz < 5; ++dz) {
        const bool mid_z = dz >= 1 && dz <= 3;
#pragma unroll 1
        for (int dy = 0; dy < 5; ++dy) {
          const size_t ro = ((size_t)dz * Hp + dy) * Wp;
          const float* w5p = wb + (size_t)((dz * 5 + dy) * 5) * 4;
#pragma unroll
          for (int dx = 0; dx < 5; ++dx) {
            const float xv = g5[ro + dx];
            const float4 w = *reinterpret_cast<const float4*>(w5p + dx * 4);
            acc[0] = fmaf(w.x, xv, acc[0]); acc[1] = fmaf(w.y, xv, acc[1]);
            acc[2] = fmaf(w.z, xv, acc[2]); acc[3] = fmaf(w.w, xv, acc[3]);
          }
          if (mid_z && dy >= 1 && dy <= 3) {
            const float* w3p = wb + (size_t)(125 + ((dz - 1) * 3 + (dy - 1)) * 3) * 4;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              const float xv = g3[ro + dx + 1];
              const float4 w = *reinterpret_cast<const float4*>(w3p + dx * 4);
              acc[0] = fmaf(w.x, xv, acc[0]); acc[1] = fmaf(w.y, xv, acc[1]);
              acc[2] = fmaf(w.z, xv, acc[2]); acc[3] = fmaf(w.w, xv, acc[3]);
            }
            if (dz == 2 && dy == 2) {
              const float xv = g1[ro + 2];
              const float4 w = *reinterpret_cast<const float4*>(wb + 152 * 4);
              acc[0] = fmaf(w.x, xv, acc[0]); acc[1] = fmaf(w.y, xv, acc[1]);
              acc[2] = fmaf(w.z, xv, acc[2]); acc[3] = fmaf(w.w, xv, acc[3]);
            }
          }
        }
      }
    }
  }
  if (sl < KS) {
#pragma unroll
    for (int c = 0; c < 4; ++c) red[((size_t)sl * 4 + c) * VP + vl] = acc[c];
  }
  __syncthreads();
  for (int it = tid; it < 4 * VP; it += nthr) {
    const int c = it / VP, v = it % VP;
    const int gv = split * Vc + v;
    if (v >= Vc || gv >= S) continue;
    float val = 0.f;
    for (int s2 = 0; s2 < KS; ++s2) val += red[((size_t)s2 * 4 + c) * VP + v];
    const size_t o = ((size_t)b * C + g * CG + ci0 + c) * S + gv;
    A.dx[o] = val + A.dO[o];
  }
}

// ---------------------------------------------------------------------------------------------------
// conv wgrad:  dw_k[co][ci][t] += sum_{b,u} gz_k[b,co,u] * x[b,ci,u + t - p];   db_k[co] += sum gz_k[b,co,u]
// ---------------------------------------------------------------------------------------------------
struct ConvWgradArgs {
  const float* x; const float* gz;
  float* dw1; float* db1; float* dw3; float* db3; float* dw5; float* db5;
  int B, C, D, H, W;
  ConvTile t;
};

template <int CG, int K>
VX_DEV void wgrad_items(const ConvWgradArgs& A, const float* xs, const float* gsk, float* __restrict__ dwk, int g,
                        int chunk, int item0, int nitems_before, int HZ, int HY, int TXP) {
  // item = (ci in chunk (4), dz, dy, co-block) x z-split; accumulators: K dx taps x 4 output channels.  With few items
  // (CG = 4: 140 per chunk) every item is dealt to VS threads that take a share of the tile's z planes each.
  constexpr int NCB = CG / 4, P = K / 2, K3 = K * K * K, VS = CG == 4 ? 2 : 1;
  const int TZ = A.t.TZ, TY = A.t.TY, TX = A.t.TX;
  const int nitems = 4 * K * K * NCB * VS;
  for (int it0 = item0 - nitems_before; it0 < nitems; it0 += blockDim.x) {
    if (it0 < 0) continue;
    const int zs = it0 % VS, it = it0 / VS;
    const int zlo = zs * TZ / VS, zhi = (zs + 1) * TZ / VS;
    const int cb = it % NCB, dy = (it / NCB) % K, dz = (it / (NCB * K)) % K, ci = it / (NCB * K * K);
    float acc[K][4];
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
    for (int z = zlo; z < zhi; ++z) {
      for (int y = 0; y < TY; ++y) {
        const float* xrow = xs + ((size_t)(ci * HZ + z + dz + 2 - P) * HY + (y + dy + 2 - P)) * TXP;
        const float* grow = gsk + ((size_t)(cb * 4) * TZ + z) * TY * TX + (size_t)y * TX;
        for (int x = 0; x < TX; x += 4) {
          const float4 xa = *reinterpret_cast<const float4*>(xrow + x);
          const float4 xb = *reinterpret_cast<const float4*>(xrow + x + 4);
          const float xr[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 gv = *reinterpret_cast<const float4*>(grow + (size_t)c * TZ * TY * TX + x);
            const float gg[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
            for (int i = 0; i < K; ++i)
#pragma unroll
              for (int v = 0; v < 4; ++v) acc[i][c] = fmaf(gg[v], xr[v + i + 2 - P], acc[i][c]);
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int co = g * CG + cb * 4 + c;
#pragma unroll
      for (int i = 0; i < K; ++i)
        atomicAdd(dwk + ((size_t)co * CG + chunk * 4 + ci) * K3 + (dz * K + dy) * K + i, acc[i][c]);
    }
  }
}

constexpr int JW_THREADS = 320;      // 280 (item, z-split) tasks per ci-chunk for every CG: one balanced round

template <int CG>
__global__ void __launch_bounds__(JW_THREADS) jlc_conv_wgrad_kernel(const __grid_constant__ ConvWgradArgs A) {
  VX_PDL_ENTRY();
  constexpr int NCB = CG / 4;
  const int g = blockIdx.y, b = blockIdx.z;
  const int TZ = A.t.TZ, TY = A.t.TY, TX = A.t.TX;
  const int tile = blockIdx.x;
  const int tx_i = tile % A.t.ntx, ty_i = (tile / A.t.ntx) % A.t.nty, tz_i = tile / (A.t.ntx * A.t.nty);
  const int z0 = tz_i * TZ, y0 = ty_i * TY, x0 = tx_i * TX;
  const int HZ = TZ + 4, HY = TY + 4, TXP = (TX + 4 + 3) & ~3;
  const int D = A.D, H = A.H, W = A.W, C = A.C;
  const size_t S = (size_t)D * H * W;
  const size_t BCS = (size_t)A.B * C * S;
  VX_DYN_SMEM(float, sm);
  float* xs = sm;                                // [4][HZ][HY][TXP], x halo tile of one ci-chunk
  float* gs = sm + (size_t)4 * HZ * HY * TXP;    // [3][CG][TZ][TY][TX]
  __shared__ float sdb[3 * 16];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int tvol = TZ * TY * TX;

  if (tid < 3 * 16) sdb[tid] = 0.f;
  for (int k = 0; k < 3; ++k)                      // [k][co][TZ][TY][TX], no halo
    stage_halo_rows(gs + (size_t)k * CG * tvol, A.gz + k * BCS + ((size_t)b * C + g * CG) * S, S, CG, z0, y0, x0, 0, TZ, TY, TX, D, H, W, 0);
  vx_cp_async_commit();
  vx_cp_async_wait_all();
  __syncthreads();
  // db partials: one warp-strided pass per (k, co)
  for (int r = tid >> 5; r < 3 * CG; r += nthr >> 5) {
    float s = 0.f;
    for (int i = tid & 31; i < tvol; i += 32) s += gs[(size_t)r * tvol + i];
    s = warp_sum(s);
    if ((tid & 31) == 0) sdb[r] = s;
  }

  for (int chunk = 0; chunk < NCB; ++chunk) {
    __syncthreads();
    const int cin0 = g * CG + chunk * 4;
    stage_halo_rows(xs, A.x + ((size_t)b * C + cin0) * S, S, 4, z0, y0, x0, 2, HZ, HY, TXP, D, H, W);
    vx_cp_async_commit();
    vx_cp_async_wait_all();
    __syncthreads();
    // a flat item list: k=5 items first (heaviest), then k=3, then k=1
    constexpr int VS = CG == 4 ? 2 : 1;
    const int n5 = 4 * 25 * NCB * VS, n3 = 4 * 9 * NCB * VS;
    wgrad_items<CG, 5>(A, xs, gs + (size_t)2 * CG * tvol, A.dw5, g, chunk, tid, 0, HZ, HY, TXP);
    wgrad_items<CG, 3>(A, xs, gs + (size_t)1 * CG * tvol, A.dw3, g, chunk, tid, n5, HZ, HY, TXP);
    wgrad_items<CG, 1>(A, xs, gs, A.dw1, g, chunk, tid, n5 + n3, HZ, HY, TXP);
  }
  __syncthreads();
  if (tid < 3 * CG) {
    const int k = tid / CG, co = g * CG + tid % CG;
    float* db = k == 0 ? A.db1 : (k == 1 ? A.db3 : A.db5);
    atomicAdd(db + co, sdb[tid]);
  }
}

// ---------------------------------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------------------------------
static inline size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }

struct JlcLayout {
  size_t S, BCS, rows;
  ConvTile tf, td, tw;      // forward, data gradient, weight gradient
  int small, ntiles_f;      // small-volume conv kernels; stats partials per row of the forward conv
  int nchunk, chunk;
  size_t off_part_z, off_part_o, off_a, off_c, off_cnt;              // forward scratch
  size_t off_dh, off_dohat, off_dO, off_gz, off_acc, off_acc2;      // backward scratch
  size_t total;
};

static int jlc_layout(const vx_jlc_desc* d, JlcLayout& L) {
  if (!d || d->B <= 0 || d->C <= 0 || d->groups <= 0 || d->C % d->groups) { set_error("jlc: bad descriptor"); return VX_ERR_BAD_DESC; }
  const int CG = d->C / d->groups;
  if (CG != 4 && CG != 8 && CG != 16) { set_error("jlc: channels per group %d not in {4,8,16}", CG); return VX_ERR_UNSUPPORTED; }
  if (d->expansion <= 0 || d->expansion * d->C > 4096) { set_error("jlc: bad expansion"); return VX_ERR_BAD_DESC; }
  L.S = (size_t)d->D * d->H * d->W;
  L.rows = (size_t)d->B * d->C;
  L.BCS = L.rows * L.S;
  L.tf = pick_tile(d->B, d->groups, CG, d->D, d->H, d->W, 256, 40 * 1024, 0);
  L.td = pick_tile(d->B, d->groups, CG, d->D, d->H, d->W, 256, 40 * 1024, 0, 132);
  L.tw = pick_tile(d->B, d->groups, CG, d->D, d->H, d->W, 256, 44 * 1024, 1);
  if (L.tf.threads == 0 || L.tw.TZ == 0) { set_error("jlc: no tile fits"); return VX_ERR_UNSUPPORTED; }
  L.chunk = 2048;
  L.nchunk = cdiv((long long)L.S, L.chunk);
  L.small = jlc_use_small(d->D, d->H, d->W) &&
            sizeof(float) * ((size_t)3 * CG * (d->D + 4) * (d->H + 4) * (d->W + 4) + (size_t)CG * 153 * 4 + 4096) <= 200 * 1024;
  L.ntiles_f = L.small ? small_geo(CG, d->D, d->H, d->W).nsplit : L.tf.ntz * L.tf.nty * L.tf.ntx;
  const int ntiles = L.ntiles_f;
  size_t off = 0;
  L.off_part_z = off; off += align256(sizeof(float) * 3 * L.rows * ntiles * 2);
  L.off_part_o = off; off += align256(sizeof(float) * L.rows * L.nchunk * 2);
  L.off_a = off; off += align256(sizeof(float) * L.rows);
  L.off_c = off; off += align256(sizeof(float) * L.rows);
  L.off_cnt = off; off += align256(sizeof(int) * L.rows);
  L.off_dh = off; off += align256(sizeof(float) * (size_t)d->B * d->expansion * d->C * L.S);
  L.off_dohat = off; off += align256(sizeof(float) * L.BCS);
  L.off_dO = off; off += align256(sizeof(float) * L.BCS);
  L.off_gz = off; off += align256(sizeof(float) * 3 * L.BCS);
  L.off_acc = off; off += align256(sizeof(float) * L.rows * 2);
  L.off_acc2 = off; off += align256(sizeof(float) * 3 * L.rows * 2);
  L.total = off;
  return VX_OK;
}

// reduction slices of the K-split kernels: the largest divisor of CG in {8, 4, 2} that keeps the CTA within JLC_KS_THREADS
static int g_jlc_ks = -1;           // -1: automatic, 0 / 1: K-split kernels off, 2 / 4 / 8: forced (vx_set_option probe)
void jlc_set_ks(int ks) { g_jlc_ks = ks; }
static int pick_ks(int CG, int nt, int max_threads) {
  if (g_jlc_ks == 0 || g_jlc_ks == 1) return 1;
  const int cand[] = {8, 4, 2};
  for (int k : cand) {
    if (g_jlc_ks > 1 && k != g_jlc_ks) continue;
    if (CG % k == 0 && nt * k <= max_threads) return k;
  }
  return 1;
}

template <int CG>
static int launch_conv_fwd(const ConvFwdArgs& A, int groups, cudaStream_t st) {
  const ConvTile& t = A.t;
  const int TXP = (t.TX + 4 + 3) & ~3;
  dim3 grid(t.ntz * t.nty * t.ntx, groups, A.B);
  if (t.VX == 4) {
    const int KS = pick_ks(CG, t.threads, JlcKs<CG>::MAXT);
    const size_t stage = (size_t)CG * (t.TZ + 4) * (t.TY + 4) * TXP + (size_t)CG * 153 * CG, red = (size_t)(KS - 1) * 48 * t.threads;
    const size_t smem_ks = sizeof(float) * ((stage > red ? stage : red) + 3 * CG * 2);
    if (KS > 1 && smem_ks <= 200 * 1024) {
      VX_SET_SMEM((jlc_conv_fwd_ks_kernel<CG, 4>), smem_ks);
      VX_LAUNCH((jlc_conv_fwd_ks_kernel<CG, 4>), grid, dim3(t.threads * KS), smem_ks, st, A, KS);
      return check_launch("jlc_conv_fwd_ks_kernel");
    }
  }
  const size_t smem = sizeof(float) * ((size_t)4 * (t.TZ + 4) * (t.TY + 4) * TXP + 4 * 153 * CG + 3 * CG * 2);
  if (t.VX == 8) {
    VX_SET_SMEM((jlc_conv_fwd_kernel<CG, 8>), smem);
    VX_LAUNCH((jlc_conv_fwd_kernel<CG, 8>), grid, dim3(t.threads), smem, st, A);
  } else {
    VX_SET_SMEM((jlc_conv_fwd_kernel<CG, 4>), smem);
    VX_LAUNCH((jlc_conv_fwd_kernel<CG, 4>), grid, dim3(t.threads), smem, st, A);
  }
  return check_launch("jlc_conv_fwd_kernel");
}

template <int CG>
static int launch_conv_dgrad(const ConvDgradArgs& A, int groups, cudaStream_t st) {
  const ConvTile& t = A.t;
  const int TXP = (t.TX + 4 + 3) & ~3;
  dim3 grid(t.ntz * t.nty * t.ntx, groups, A.B);
  if (t.VX == 4) {
    const int KS = pick_ks(CG, t.threads, JlcKs<CG>::MAXT);
    const size_t stage = (size_t)CG * (t.TZ + 4) * (t.TY + 4) * TXP + (size_t)CG * 125 * CG, red = (size_t)(KS - 1) * 16 * t.threads;
    const size_t smem_ks = sizeof(float) * (stage > red ? stage : red);
    if (KS > 1 && smem_ks <= 200 * 1024) {
      VX_SET_SMEM((jlc_conv_dgrad_ks_kernel<CG, 4>), smem_ks);
      VX_LAUNCH((jlc_conv_dgrad_ks_kernel<CG, 4>), grid, dim3(t.threads * KS), smem_ks, st, A, KS);
      return check_launch("jlc_conv_dgrad_ks_kernel");
    }
  }
  const size_t smem = sizeof(float) * ((size_t)4 * (t.TZ + 4) * (t.TY + 4) * TXP + 4 * 125 * CG);
  if (t.VX == 8) {
    VX_SET_SMEM((jlc_conv_dgrad_kernel<CG, 8>), smem);
    VX_LAUNCH((jlc_conv_dgrad_kernel<CG, 8>), grid, dim3(t.threads), smem, st, A);
  } else {
    VX_SET_SMEM((jlc_conv_dgrad_kernel<CG, 4>), smem);
    VX_LAUNCH((jlc_conv_dgrad_kernel<CG, 4>), grid, dim3(t.threads), smem, st, A);
  }
  return check_launch("jlc_conv_dgrad_kernel");
}

// threads of a small-volume CTA: the slab x slices actually used, at least 256 (the staging loops are dealt to all of them)
static int small_block(const SmallGeo& G) { const int t = G.VP * G.KS; return t < 256 ? 256 : t; }

template <int CG>
static int launch_conv_small_fwd(const ConvFwdArgs& A, int groups, cudaStream_t st) {
  const SmallGeo G = small_geo(CG, A.D, A.H, A.W);
  const size_t smem = sizeof(float) * ((size_t)CG * G.PV + (size_t)CG * 153 * 4 + (size_t)G.KS * 12 * G.VP);
  if (smem > 200 * 1024) { set_error("jlc small fwd: volume too large for shared memory"); return VX_ERR_UNSUPPORTED; }
  VX_SET_SMEM((jlc_conv_small_fwd_kernel<CG>), smem);
  VX_LAUNCH((jlc_conv_small_fwd_kernel<CG>), dim3(G.nsplit, groups * (CG / 4), A.B), dim3(small_block(G)), smem, st, A, G);
  return check_launch("jlc_conv_small_fwd_kernel");
}

template <int CG>
static int launch_conv_small_dgrad(const ConvDgradArgs& A, int groups, cudaStream_t st) {
  const SmallGeo G = small_geo(CG, A.D, A.H, A.W);
  const size_t smem = sizeof(float) * ((size_t)3 * CG * G.PV + (size_t)CG * 153 * 4 + (size_t)G.KS * 4 * G.VP);
  if (smem > 200 * 1024) { set_error("jlc small dgrad: volume too large for shared memory"); return VX_ERR_UNSUPPORTED; }
  VX_SET_SMEM((jlc_conv_small_dgrad_kernel<CG>), smem);
  VX_LAUNCH((jlc_conv_small_dgrad_kernel<CG>), dim3(G.nsplit, groups * (CG / 4), A.B), dim3(small_block(G)), smem, st, A, G);
  return check_launch("jlc_conv_small_dgrad_kernel");
}

template <int CG>
static int launch_conv_wgrad(const ConvWgradArgs& A, int groups, cudaStream_t st) {
  const ConvTile& t = A.t;
  const int TXP = (t.TX + 4 + 3) & ~3;
  const size_t smem = sizeof(float) * ((size_t)4 * (t.TZ + 4) * (t.TY + 4) * TXP + (size_t)3 * CG * t.TZ * t.TY * t.TX);
  dim3 grid(t.ntz * t.nty * t.ntx, groups, A.B);
  VX_SET_SMEM((jlc_conv_wgrad_kernel<CG>), smem);
  VX_LAUNCH((jlc_conv_wgrad_kernel<CG>), grid, dim3(JW_THREADS), smem, st, A);
  return check_launch("jlc_conv_wgrad_kernel");
}

#define VX_TRY(expr) do { int _rc = (expr); if (_rc != VX_OK) return _rc; } while (0)

}  // namespace vx

using namespace vx;

extern "C" size_t vx_jlc_workspace(const vx_jlc_desc* d) {
  JlcLayout L;
  if (jlc_layout(d, L) != VX_OK) return 0;
  return L.total;
}

extern "C" int vx_jlc_fwd(const vx_jlc_desc* d, const void* const* in, void* const* out, void* workspace,
                          size_t workspace_bytes, vx_stream_t stream) {
  JlcLayout L;
  VX_TRY(jlc_layout(d, L));
  prof_scope("jlc_fwd B%d C%d S%d", d->B, d->C, d->D * d->H * d->W);
  set_seed_dev(d->seed_offset);
  if (!workspace || workspace_bytes < L.total) { set_error("jlc_fwd: workspace %zu < %zu", workspace_bytes, L.total); return VX_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  const int CG = d->C / d->groups, eC = d->expansion * d->C;
  const float* x = (const float*)in[0];
  float* y = (float*)out[0];
  float* z = (float*)out[1];
  float* o = (float*)out[2];
  float* hpre = (float*)out[3];
  float* stats = (float*)out[4];
  float* part_z = (float*)(ws + L.off_part_z);
  float* part_o = (float*)(ws + L.off_part_o);
  float* aff_a = (float*)(ws + L.off_a);
  float* aff_c = (float*)(ws + L.off_c);
  const int rows = (int)L.rows, S = (int)L.S;

  ConvFwdArgs A{};
  A.x = x; A.w1 = (const float*)in[1]; A.b1 = (const float*)in[2]; A.w3 = (const float*)in[3]; A.b3 = (const float*)in[4];
  A.w5 = (const float*)in[5]; A.b5 = (const float*)in[6];
  A.z = z; A.part = part_z; A.cnt = (int*)(ws + L.off_cnt); A.B = d->B; A.C = d->C; A.D = d->D; A.H = d->H; A.W = d->W; A.t = L.tf;
  const int npos = L.tf.TZ * L.tf.TY * (L.tf.TX / L.tf.VX);
  A.uniform_warps = (npos % 32 == 0) ? 1 : 0;
  prof_bytes(4.0 * sizeof(float) * (double)L.BCS);       // x in, z1 z3 z5 out
  prof_flops(2.0 * 153.0 * CG * (double)L.BCS);
  if (L.small) {
    if (CG == 4) VX_TRY(launch_conv_small_fwd<4>(A, d->groups, st));
    else if (CG == 8) VX_TRY(launch_conv_small_fwd<8>(A, d->groups, st));
    else VX_TRY(launch_conv_small_fwd<16>(A, d->groups, st));
  } else if (CG == 4) VX_TRY(launch_conv_fwd<4>(A, d->groups, st));
  else if (CG == 8) VX_TRY(launch_conv_fwd<8>(A, d->groups, st));
  else VX_TRY(launch_conv_fwd<16>(A, d->groups, st));

  const int ntiles = L.ntiles_f;
  prof_bytes(5.0 * sizeof(float) * (double)L.BCS);       // x, z1 z3 z5 in, o out
  VX_LAUNCH(jlc_combine_kernel, dim3(L.nchunk, rows), dim3(256), 0, st, x, (const float*)z, (const float*)part_z, ntiles, stats, o,
            part_o, (int*)(ws + L.off_cnt), aff_a, aff_c, rows, S, L.chunk, d->eps);
  VX_TRY(check_launch("jlc_combine_kernel"));

  // small levels: both contractions in one launch (the hidden activation stays in shared memory)
  {
    FfnBatch fb{};
    fb.nprob = 1; fb.B = d->B; fb.S = S;
    FfnProblem& f = fb.p[0];
    f.x = o; f.C = d->C; f.pro_a = aff_a; f.pro_c = aff_c; f.pro_bstride = d->C;
    f.W1 = (const float*)in[7]; f.b1 = (const float*)in[8]; f.eC = eC; f.hpre = hpre;
    f.W2 = (const float*)in[9]; f.b2 = (const float*)in[10];
    if (d->training && d->drop_p > 0.f) { f.drop_p = d->drop_p; f.seed = d->seed; f.site = 1; }
    f.res = o; f.res_scale = 1.f; f.y = y;
    int rc = pw_ffn_small(fb, st);
    if (rc == 1) rc = pw_ffn_tc(fb, st);
    if (rc <= 0) return rc;
  }
  // hpre = W1 IN(o) + b1
  PwBatch pb{};
  pb.nprob = 1; pb.B = d->B; pb.S = S;
  PwProblem& p1 = pb.p[0];
  p1.src[0] = PwSrc{o, d->C}; p1.nsrc = 1; p1.Ci = d->C;
  p1.seg[0] = PwSeg{(const float*)in[7], (const float*)in[8], d->C, eC, hpre}; p1.nseg = 1; p1.Co = eC;
  p1.pro = PRO_AFFINE; p1.pro_a = aff_a; p1.pro_c = aff_c; p1.pro_bstride = d->C;
  VX_TRY(pw_forward(pb, st));
  // y = o + Dropout(W2 GELU(hpre) + b2)
  PwBatch pb2{};
  pb2.nprob = 1; pb2.B = d->B; pb2.S = S;
  PwProblem& p2 = pb2.p[0];
  p2.src[0] = PwSrc{hpre, eC}; p2.nsrc = 1; p2.Ci = eC;
  p2.seg[0] = PwSeg{(const float*)in[9], (const float*)in[10], eC, d->C, y}; p2.nseg = 1; p2.Co = d->C;
  p2.pro = PRO_GELU;
  if (d->training && d->drop_p > 0.f) { p2.drop_p = d->drop_p; p2.seed = d->seed; p2.site = 1; }
  p2.res = o; p2.res_scale = 1.f;
  VX_TRY(pw_forward(pb2, st));
  return VX_OK;
}

extern "C" int vx_jlc_bwd(const vx_jlc_desc* d, const void* const* in, void* const* out, void* workspace,
                          size_t workspace_bytes, vx_stream_t stream) {
  JlcLayout L;
  VX_TRY(jlc_layout(d, L));
  prof_scope("jlc_bwd B%d C%d S%d", d->B, d->C, d->D * d->H * d->W);
  set_seed_dev(d->seed_offset);
  if (!workspace || workspace_bytes < L.total) { set_error("jlc_bwd: workspace %zu < %zu", workspace_bytes, L.total); return VX_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  SideJoin side_guard(st);
  char* ws = (char*)workspace;
  const int CG = d->C / d->groups, eC = d->expansion * d->C, C = d->C;
  const int rows = (int)L.rows, S = (int)L.S;
  const float* dy = (const float*)in[0];
  const float* x = (const float*)in[1];
  const float* z = (const float*)in[2];
  const float* o = (const float*)in[3];
  const float* hpre = (const float*)in[4];
  const float* stats = (const float*)in[5];
  const float* w1 = (const float*)in[6];
  const float* w3 = (const float*)in[7];
  const float* w5 = (const float*)in[8];
  const float* fw1 = (const float*)in[9];
  const float* fw2 = (const float*)in[11];
  float* dx = (float*)out[0];
  float* dw1 = (float*)out[1]; float* db1 = (float*)out[2];
  float* dw3 = (float*)out[3]; float* db3 = (float*)out[4];
  float* dw5 = (float*)out[5]; float* db5 = (float*)out[6];
  float* dfw1 = (float*)out[7]; float* dfb1 = (float*)out[8];
  float* dfw2 = (float*)out[9]; float* dfb2 = (float*)out[10];
  float* aff_a = (float*)(ws + L.off_a);
  float* aff_c = (float*)(ws + L.off_c);
  float* dh = (float*)(ws + L.off_dh);
  float* dohat = (float*)(ws + L.off_dohat);
  float* dO = (float*)(ws + L.off_dO);
  float* gz = (float*)(ws + L.off_gz);
  const float* stats_o = stats + (size_t)2 * 3 * rows;
  const bool drop = d->training && d->drop_p > 0.f;

  // The zeroing of every atomically accumulated buffer and the (a, c) affine of IN(o) run on the side stream from the start
  // of the op: their consumers are the weight-gradient kernels (side stream) and the statistics kernels four launches down
  // the main stream (side_wait below), so the data-gradient chain starts with the first contraction instead of two
  // latency-bound helper launches.  (Every consumer of these buffers runs on the side stream.)
  {
    cudaStream_t sz = side_fork(st);
    ZeroList zl;
    zl.add(dw1, (size_t)C * CG); zl.add(dw3, (size_t)C * CG * 27); zl.add(dw5, (size_t)C * CG * 125);
    zl.add(db1, C); zl.add(db3, C); zl.add(db5, C);
    zl.add(dfw1, (size_t)eC * C); zl.add(dfb1, eC); zl.add(dfw2, (size_t)eC * C); zl.add(dfb2, C);
    VX_TRY(zero_many(zl, sz));
    VX_TRY(stats_to_affine(stats_o, aff_a, aff_c, rows, sz));
  }

  // dh = (W2^T (dy * mask)) * GELU'(hpre);  dohat = W1^T dh  -- one launch on the small levels
  int ffn_fused = 0;
  {
    FfnBwdBatch fb{}; fb.nprob = 1; fb.B = d->B; fb.S = S;
    FfnBwdProblem& f = fb.p[0];
    f.dy = dy; f.C = C; f.W2 = fw2; f.eC = eC; f.hpre = hpre; f.dh = dh; f.W1 = fw1; f.dx = dohat;
    if (drop) { f.out_drop_p = d->drop_p; f.out_seed = d->seed; f.out_site = 1; }
    int rc = pw_ffn_small_bwd(fb, st);
    if (rc == 1) rc = pw_ffn_tc_bwd(fb, st);
    if (rc != VX_OK && rc != 1) return rc;
    ffn_fused = rc == VX_OK;
  }
  if (!ffn_fused) {
    PwBatch pb{}; pb.nprob = 1; pb.B = d->B; pb.S = S;
    PwProblem& p = pb.p[0];
    p.src[0] = PwSrc{dy, C}; p.nsrc = 1; p.Ci = C;
    p.seg[0] = PwSeg{fw2, nullptr, eC, C, dh}; p.nseg = 1; p.Co = eC; p.transposed = 1;
    if (drop) { p.pro = PRO_DROPOUT; p.pro_drop_p = d->drop_p; p.pro_seed = d->seed; p.pro_site = 1; }
    p.mulgrad = hpre;
    VX_TRY(pw_forward(pb, st));
  }
  // dW2 = (dy*mask) GELU(hpre)^T ; dW1 = dh IN(o)^T   (side stream: overlaps everything below)
  {
    WgBatch wb{}; wb.nprob = 2; wb.B = d->B; wb.S = S;
    WgProblem& a = wb.p[0];
    a.dY = dy; a.Co = C; a.src[0] = PwSrc{hpre, eC}; a.nsrc = 1; a.Ci = eC; a.xpro = PRO_GELU;
    if (drop) { a.y_drop_p = d->drop_p; a.y_seed = d->seed; a.y_site = 1; }
    a.dW = dfw2; a.ld = eC; a.db = dfb2;
    WgProblem& b2 = wb.p[1];
    b2.dY = dh; b2.Co = eC; b2.src[0] = PwSrc{o, C}; b2.nsrc = 1; b2.Ci = C; b2.xpro = PRO_AFFINE;
    b2.xa = aff_a; b2.xc = aff_c; b2.x_bstride = C;
    b2.dW = dfw1; b2.ld = C; b2.db = dfb1;
    VX_TRY(pw_wgrad(wb, st));
  }
  // dohat = W1^T dh
  if (!ffn_fused) {
    PwBatch pb{}; pb.nprob = 1; pb.B = d->B; pb.S = S;
    PwProblem& p = pb.p[0];
    p.src[0] = PwSrc{dh, eC}; p.nsrc = 1; p.Ci = eC;
    p.seg[0] = PwSeg{fw1, nullptr, C, eC, dohat}; p.nseg = 1; p.Co = C; p.transposed = 1;
    VX_TRY(pw_forward(pb, st));
  }
  {
    // one CTA per row: 32 threads per ~4 elements, at most 1024
    int T = ((S + 3) / 4 + 31) & ~31;
    T = T < 32 ? 32 : (T > 1024 ? 1024 : T);
    prof_bytes(10.0 * sizeof(float) * (double)L.BCS);    // dy, dohat, o, z(3) in, dO + gz(3) out
    VX_LAUNCH(jlc_bwd_abc_kernel, dim3(rows), dim3(T), 0, st, dy, (const float*)dohat, o, z, stats, dO, gz, rows, S);
    VX_TRY(check_launch("jlc_bwd_abc_kernel"));
  }

  // weight gradients first, on the side stream (forked here, after gz is complete): dgrad then runs beside them
  cudaStream_t st_w = side_fork(st);
  ConvWgradArgs Wg{};
  Wg.x = x; Wg.gz = gz; Wg.dw1 = dw1; Wg.db1 = db1; Wg.dw3 = dw3; Wg.db3 = db3; Wg.dw5 = dw5; Wg.db5 = db5;
  Wg.B = d->B; Wg.C = C; Wg.D = d->D; Wg.H = d->H; Wg.W = d->W; Wg.t = L.tw;
  prof_bytes(4.0 * sizeof(float) * (double)L.BCS);       // x, gz(3) in (weight gradients are KBs)
  prof_flops(2.0 * 153.0 * (C / d->groups) * (double)L.BCS);
  if (CG == 4) VX_TRY(launch_conv_wgrad<4>(Wg, d->groups, st_w));
  else if (CG == 8) VX_TRY(launch_conv_wgrad<8>(Wg, d->groups, st_w));
  else VX_TRY(launch_conv_wgrad<16>(Wg, d->groups, st_w));

  ConvDgradArgs G{};
  G.gz = gz; G.dO = dO; G.w1 = w1; G.w3 = w3; G.w5 = w5; G.dx = dx;
  G.B = d->B; G.C = C; G.D = d->D; G.H = d->H; G.W = d->W; G.t = L.td;
  prof_bytes(5.0 * sizeof(float) * (double)L.BCS);       // gz(3), dO in, dx out
  prof_flops(2.0 * 153.0 * (C / d->groups) * (double)L.BCS);
  if (L.small) {
    if (CG == 4) VX_TRY(launch_conv_small_dgrad<4>(G, d->groups, st));
    else if (CG == 8) VX_TRY(launch_conv_small_dgrad<8>(G, d->groups, st));
    else VX_TRY(launch_conv_small_dgrad<16>(G, d->groups, st));
  } else if (CG == 4) VX_TRY(launch_conv_dgrad<4>(G, d->groups, st));
  else if (CG == 8) VX_TRY(launch_conv_dgrad<8>(G, d->groups, st));
  else VX_TRY(launch_conv_dgrad<16>(G, d->groups, st));

  return VX_OK;
}
