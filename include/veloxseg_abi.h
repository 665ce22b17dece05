/* libveloxseg_sm100 — C ABI of the VeloxSeg hot path for NVIDIA B200 (sm_100a).
 *
 * The reference (JinPLu/VeloxSeg) has no FFI boundary: its hot path is the Python nn.Module surface of
 * model/components/ (SURVEY.md §8b).  These entry points are what a torch custom-op wrapper binds for that
 * path; every function names the reference code it replaces (path:line into the reference repository).
 *
 * Conventions
 *  - all tensors are device pointers to fp32, NCDHW contiguous, owned by the caller;
 *  - `in` / `out` are arrays of device pointers in the order documented per function;
 *  - the library never allocates, never synchronises and never throws: it enqueues kernels on `stream`
 *    and returns 0 or a negative vx_status (text via vx_last_error_string());
 *  - `workspace` is caller-allocated scratch of at least vx_<op>_workspace(desc) bytes, 256-B aligned;
 *  - there is no CPU implementation behind this ABI.
 */
#ifndef VELOXSEG_ABI_H_
#define VELOXSEG_ABI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vx_stream_t; /* cudaStream_t */

typedef enum {
  VX_OK = 0,
  VX_ERR_BAD_DESC = -1,
  VX_ERR_WORKSPACE = -2,
  VX_ERR_LAUNCH = -3,
  VX_ERR_UNSUPPORTED = -4
} vx_status;

#define VX_MAX_MODAL 4  /* modalities per PWA block / streams per mixer */
#define VX_MAX_SCALES 6 /* paired-window scales per PWA level          */

int vx_version(void);
const char* vx_last_error_string(void);
/* Process-wide switches (diagnostics / A-B measurements; defaults give the production path).
 *   VX_OPT_PW_TENSOR_CORES  1 (default): 1x1 contractions with >= 1024 voxels (weight gradients: >= 512) run on the tcgen05
 *                           3xTF32 kernels;
 *                           0: every contraction uses the fp32 SIMT kernels.
 *   VX_OPT_PW_SMALL_MAX_S   voxel count below which the warp-per-(32 voxels x 4 channels) kernel is used (tuning probe).
 *   VX_OPT_PW_TC_MIN_S      voxel count from which the tensor-core kernel is used (tuning probe).
 *   VX_OPT_JLC_TILE_FWD / _WGRAD   force the (z, y) extent of the JLC conv tile: value = tz << 8 | ty, 0 = automatic.
 *   VX_OPT_JLC_SMALL_MAX_S  volumes of up to this many voxels use the small-volume JLC conv kernels (default 512, 0 = never).
 *   VX_OPT_WGRAD_TC_MIN_S   voxel count from which weight gradients of the 1x1 contractions run on the tcgen05 kernel
 *                           (default 512; VX_OPT_PW_TENSOR_CORES = 0 switches it off together with the forward kernel).
 *   VX_OPT_SIDE_WGRAD       1 (default): inside a backward op the weight-gradient kernels run on a library-owned side
 *                           stream that forks from and joins back into the caller's stream before the op returns.
 *   VX_OPT_PRECISION        0 (default): fp32 mode -- every kernel fp32-accurate (tensor-core contractions 3xTF32);
 *                           1: bf16 mode (north_star's second precision mode) -- the tensor-core contractions (1x1 convs of
 *                           levels 1-2 and their weight gradients, dense 3x3x3 convs forward / data / weight gradient) round
 *                           both operands to bfloat16 (RNE), issue ONE product with fp32 accumulation and round the stored
 *                           activation to bfloat16; statistics, norms, softmax, losses, optimiser stay fp32.  Storage stays fp32.
 *   VX_OPT_JLC_KS           reduction slices of the level-1/2 JLC convolution kernels: -1 (default) automatic, 0 / 1 off (one
 *                           thread per 4 voxels x 4 channels, 4-channel staging chunks), 2 / 4 / 8 forced (A/B probe).
 *   VX_OPT_PDL              0 (default): plain stream-ordered launches.  1: every kernel is launched with programmatic dependent
 *                           launch allowed (cudaLaunchAttributeProgrammaticStreamSerialization; every kernel starts with
 *                           griddepcontrol.wait + griddepcontrol.launch_dependents), also as programmatic edges inside a captured
 *                           CUDA graph.  Measured 3-5 % SLOWER on the train step (profiles/r3a_pdl_ab.txt): A/B switch only.
 *   VX_OPT_ATTN_TC          1 (default): the attention forward of windows with L >= 128 tokens (L % 16 == 0, 8 channels per head)
 *                           runs on the tcgen05 kernel (QK^T and PV as 3xTF32 MMAs, P in tensor memory); 0: fp32 SIMT kernel.
 *   VX_OPT_JLC_SMALL_THREADS  CTA size of the small-volume JLC convolution kernels: 256 / 512 (default) / 1024 (tuning probe: more
 *                           threads = more reduction slices = a shorter serial tap walk per thread).
 *   VX_OPT_FFN_TC           1 (default): the two-layer MLP of the level-1/2 JLC and PWA blocks (S >= 1024, fp32 mode) runs as one
 *                           tcgen05 kernel per direction (hidden activation handed from the first accumulator to the second MMA inside
 *                           tensor memory); 0: two contraction launches.
 *   VX_OPT_CONV3_TRACE      0 (default).  1: CTA 0 of the dense-convolution forward kernel records clock64() at its phase
 *                           boundaries; vx_conv3_trace() copies the 64 stamps out (developer diagnostics, tools/conv3_phases.py). */
enum { VX_OPT_PW_TENSOR_CORES = 1, VX_OPT_PW_SMALL_MAX_S = 2, VX_OPT_PW_TC_MIN_S = 3, VX_OPT_JLC_TILE_FWD = 4,
       VX_OPT_JLC_TILE_WGRAD = 5, VX_OPT_JLC_SMALL_MAX_S = 8, VX_OPT_WGRAD_TC_MIN_S = 9, VX_OPT_SIDE_WGRAD = 10,
       VX_OPT_CONV3_TRACE = 13, VX_OPT_PRECISION = 14, VX_OPT_JLC_KS = 15, VX_OPT_PDL = 16, VX_OPT_ATTN_TC = 17, VX_OPT_JLC_SMALL_THREADS = 18, VX_OPT_FFN_TC = 19 };
int vx_set_option(int option, int value);
int vx_conv3_trace(long long* out64, int n);
/* Measurement helpers (bench.py): kind 0 = fp32 FMA throughput probe (2 * 8 * 32 * iters * 148 * 8 * 256 flops per launch,
 * `scratch` = one device float), kind 1 = one empty kernel (calibrates the per-launch overhead of the event profiler). */
int vx_microbench(int kind, int iters, void* scratch, vx_stream_t stream);
/* Strided block copy between a pinned host volume and its device copy (either direction; both sides share `pitch_bytes`):
 * `rows` rows of `width_bytes` each, `pitch_bytes` apart -- one call per (slab of planes x slab of rows) of a channel, so that the
 * sliding-window driver (utils/inference_petct.py:214-230 through veloxseg_b200.inference.sliding_window_labels) can upload the
 * volume in the order its windows need it.  kind: 0 = host to device, 1 = device to host.  Asynchronous on `stream`. */
int vx_copy_block_async(void* dst, const void* src, size_t pitch_bytes, size_t width_bytes, size_t rows, int kind, vx_stream_t stream);
/* number of kernels this library has enqueued since it was loaded (all threads, all streams) */
uint64_t vx_launch_count(void);
/* Per-kernel timing with CUDA events on the launching stream (diagnostics for bench.py; off by default).
 * vx_profile_enable(1) brackets every subsequent launch with an event pair; vx_profile_report() synchronises those
 * events and writes one line per (scope, kernel): "scope|kernel|launches|total_ms|algorithmic_bytes|algorithmic_flops\n"; returns the bytes needed. */
int vx_profile_enable(int on);
void vx_profile_reset(void);
size_t vx_profile_report(char* buf, size_t cap);
/* one line per recorded launch: "scope|kernel|start_us|dur_us" (device timeline over all streams used) */
size_t vx_profile_timeline(char* buf, size_t cap);

/* ---------------------------------------------------------------------------------------------------
 * JLC block — replaces JLC.forward, model/components/conv_blocks.py:41-75 (and autograd's backward of it).
 *   o = x + sum_{k in {1,3,5}} GELU(IN(gconv_k(x) + b_k));   y = o + Dropout(W2 GELU(W1 IN(o) + b1) + b2)
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t B, C, D, H, W;   /* x is (B, C, D, H, W)                                   */
  int32_t groups;          /* conv groups; c_g = C / groups in {4, 8, 16}            */
  int32_t expansion;       /* e: channel_conv hidden width = e * C                   */
  float eps;               /* InstanceNorm eps (1e-5)                                */
  float drop_p;            /* dropout on the channel_conv output (0 disables)        */
  int32_t training;        /* dropout is applied only when training != 0            */
  uint64_t seed;           /* counter-based RNG key for this call's dropout mask     */
  const uint64_t* seed_offset; /* optional DEVICE pointer: the kernels add *seed_offset to `seed`, so a captured
                              CUDA graph draws a fresh mask on every replay (NULL: no offset)            */
} vx_jlc_desc;

/* fwd  in : x, w1,b1, w3,b3, w5,b5 (grouped conv weights (C, c_g, k,k,k) + bias (C)), fw1 (eC, C), fb1 (eC),
 *           fw2 (C, eC), fb2 (C)                                               [11 pointers]
 *      out: y (B,C,S), z (3,B,C,S) raw conv outputs, o (B,C,S), hpre (B,eC,S), stats (4, B*C, 2) = (mean, rstd)
 *           of z1,z3,z5,o                                                       [5 pointers]
 * bwd  in : dy, x, z, o, hpre, stats, w1,w3,w5, fw1, fb1, fw2                    [12 pointers]
 *      out: dx, dw1,db1, dw3,db3, dw5,db5, dfw1,dfb1, dfw2,dfb2                  [11 pointers]         */
size_t vx_jlc_workspace(const vx_jlc_desc* d);
int vx_jlc_fwd(const vx_jlc_desc* d, const void* const* in, void* const* out, void* workspace,
               size_t workspace_bytes, vx_stream_t stream);
int vx_jlc_bwd(const vx_jlc_desc* d, const void* const* in, void* const* out, void* workspace,
               size_t workspace_bytes, vx_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * Modal mixer — replaces `attn2conv_i(torch.cat(attn_i, 1))` and the `+` that follows, model/Encoder.py:334-337,
 * 344-361; also serves the identical RC adapters `enc2rc_i(concat(attn, enc))`, model/Decoder.py:54-57,80-83.
 *   y = [addend +] IN(W . cat_m(a_m) + b)
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t B, S;                     /* S = D*H*W                                           */
  int32_t n_streams;                /* M                                                   */
  int32_t stream_ch[VX_MAX_MODAL];  /* channels of each stream; K = sum                    */
  int32_t C_out;                    /* N                                                   */
  int32_t has_addend;               /* 1: y = addend + IN(...)                             */
  float eps;
} vx_mixer_desc;

/* fwd  in : a_0..a_{M-1}, W (C_out, K), b (C_out), addend or NULL      out: y, t (B,C_out,S) pre-norm, stats (B*C_out,2)
 * bwd  in : dy, a_0..a_{M-1}, W, t, stats                               out: da_0..da_{M-1}, dW, db
 *      (the addend's gradient is dy itself; the caller routes it)                                          */
size_t vx_mixer_workspace(const vx_mixer_desc* d);
int vx_mixer_fwd(const vx_mixer_desc* d, const void* const* in, void* const* out, void* workspace,
                 size_t workspace_bytes, vx_stream_t stream);
int vx_mixer_bwd(const vx_mixer_desc* d, const void* const* in, void* const* out, void* workspace,
                 size_t workspace_bytes, vx_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * InstanceNorm3d(affine=False) — the norm behind DownConv / UpConv, model/components/conv_blocks.py:19-21,37-39,
 * model/components/common_function.py:62-66.   y = IN(x) [+ addend]
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t rows;  /* B*C */
  int32_t S;
  float eps;
  int32_t has_addend;
  int32_t bias_channels;  /* bwd only, 0 = off: out[1] = db (bias_channels floats) = per-channel sums of dx -- the gradient of
                             a per-channel bias added in front of the norm (conv bias of DownConv / UpConv: it cancels in
                             the forward value, so the conv runs without it and its gradient comes from here)           */
} vx_inorm_desc;
/* fwd in: x, addend|NULL   out: y, stats (rows, 2)         bwd in: dy, x, stats   out: dx [, db] */
int vx_inorm_fwd(const vx_inorm_desc* d, const void* const* in, void* const* out, vx_stream_t stream);
int vx_inorm_bwd(const vx_inorm_desc* d, const void* const* in, void* const* out, vx_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * PWA block — replaces Paired_Windows_TransformerBlock.forward, model/components/PWA.py:433-439, i.e.
 * MultiModal_Paired_Windows_Attention.forward (PWA.py:329-379: LN, Q/K/V 1x1, window_gathering_3d :106-140,
 * attention_operation :308-327, window_scattering_3d :177-200, mix 1x1, residual) plus the LN + FFN tail
 * (attention_utils.py:29-71).  Window geometry is the output of get_window_sizes (PWA.py:56-85).
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t B, M, C;                 /* modalities, channels per modality                     */
  int32_t D, H, W;                 /* spatial extent of this level (reference names h,w,d)   */
  int32_t heads, n_scales;
  int32_t big[VX_MAX_SCALES][3];   /* big window per scale                                   */
  int32_t small[VX_MAX_SCALES][3]; /* small (pooled) window per scale                        */
  int32_t c_qk, c_v;               /* TOTAL q/k and v channels (= n_scales*heads*per-head)   */
  int32_t ffn_expansion;
  float ln_eps;                    /* 1e-6                                                   */
  float attn_drop, proj_drop;      /* dropout on softmax weights / on mix + FFN outputs      */
  int32_t training;
  uint64_t seed;
  const uint64_t* seed_offset;     /* optional device pointer added to `seed` by the kernels (see vx_jlc_desc) */
} vx_pwa_desc;

/* Per-modality parameter block, in this order (16 pointers):
 *   ln1_w, ln1_b, wq, bq, wk, bk, wv, bv, wmix, bmix, ln2_w, ln2_b, w1, b1, w2, b2
 * fwd  in : x_0..x_{M-1}, params_0 .. params_{M-1}, bias_table ((2n-1)^3, heads), rel_index (l, l) int64
 *      out: z_0..z_{M-1} (block outputs), then the saved-for-backward buffers listed by vx_pwa_saved_layout()
 * bwd  in : dz_0..dz_{M-1}, x_0.., params.., bias_table, rel_index, saved buffers
 *      out: dx_0..dx_{M-1}, dparams_0 .. dparams_{M-1} (same 16-slot order), dbias_table                    */
typedef struct {
  int32_t n_saved;
  size_t saved_bytes[24];
} vx_pwa_saved;
int vx_pwa_saved_layout(const vx_pwa_desc* d, vx_pwa_saved* layout);
size_t vx_pwa_workspace(const vx_pwa_desc* d);
int vx_pwa_block_fwd(const vx_pwa_desc* d, const void* const* in, void* const* out, void* workspace,
                     size_t workspace_bytes, vx_stream_t stream);
int vx_pwa_block_bwd(const vx_pwa_desc* d, const void* const* in, void* const* out, void* workspace,
                     size_t workspace_bytes, vx_stream_t stream);

/* Integer window partition only (bit-exact check of PWA.py:106-140): tokens (B, heads, Ns, l, c) gathered from
 * x (B, n_scales*heads*c, D,H,W) with the max-pool arg-max voxel index per token element. */
int vx_pwa_gather(const vx_pwa_desc* d, int32_t channels_total, const void* x, void* tokens, void* argmax,
                  vx_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * SDKT — Gram matrix, model/components/common_function.py:8-14, and the feature loss, utils/loss.py:58-64.
 *   G[b,m,n] = sum_s x[b,m,s] x[b,n,s] / (C*S);      L = sum_t mean((G_s - G_t)^2) / T
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t B, C, S;
} vx_gram_desc;
size_t vx_gram_workspace(const vx_gram_desc* d);
/* fwd in: x   out: G (B,C,C)          bwd in: dG, x   out: dx */
int vx_gram_fwd(const vx_gram_desc* d, const void* const* in, void* const* out, void* workspace,
                size_t workspace_bytes, vx_stream_t stream);
int vx_gram_bwd(const vx_gram_desc* d, const void* const* in, void* const* out, vx_stream_t stream);

typedef struct {
  int32_t n_elem;     /* B*C*C       */
  int32_t n_teachers; /* T <= VX_MAX_MODAL */
} vx_sdkt_loss_desc;
/* fwd in: G_s, G_t0..   out: loss (1)       bwd in: dloss (1), G_s, G_t0..   out: dG_s, dG_t0.. */
int vx_sdkt_loss_fwd(const vx_sdkt_loss_desc* d, const void* const* in, void* const* out, vx_stream_t stream);
int vx_sdkt_loss_bwd(const vx_sdkt_loss_desc* d, const void* const* in, void* const* out, vx_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * Channel-first LayerNorm + 1x1 conv — the PatchMerging tail, model/components/attention_utils.py:163-168
 * (the 2x2x2 space-to-depth gather stays a view/cat on the caller side).    y = W . LN(x)     (no bias)
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t B, C_in, C_out, S;
  float eps;
} vx_lnpw_desc;
/* fwd in: x, ln_w, ln_b, W   out: y, xhat (B,C_in,S), rstd (B,S)
 * bwd in: dy, xhat, rstd, ln_w, W, ln_b   out: dx, dln_w, dln_b, dW                                        */
size_t vx_lnpw_workspace(const vx_lnpw_desc* d);
int vx_lnpw_fwd(const vx_lnpw_desc* d, const void* const* in, void* const* out, void* workspace,
                size_t workspace_bytes, vx_stream_t stream);
int vx_lnpw_bwd(const vx_lnpw_desc* d, const void* const* in, void* const* out, void* workspace,
                size_t workspace_bytes, vx_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * Trilinear resize, align_corners=True — VeloxSeg.scale_prediction, model/VeloxSeg.py:177-184 (applied to the four
 * deep-supervision outputs in training, VeloxSeg.py:200-202).   y (planes, D,H,W) = resize(x (planes, d,h,w))
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t planes;      /* B * C               */
  int32_t d, h, w;     /* source extent       */
  int32_t D, H, W;     /* destination extent  */
} vx_resize_desc;
size_t vx_resize_workspace(const vx_resize_desc* d);
/* fwd in: x   out: y            bwd in: dy   out: dx (needs the workspace) */
int vx_resize_trilinear_fwd(const vx_resize_desc* d, const void* const* in, void* const* out, vx_stream_t stream);
int vx_resize_trilinear_bwd(const vx_resize_desc* d, const void* const* in, void* const* out, void* workspace,
                            size_t workspace_bytes, vx_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * Deep-supervision segmentation loss -- utils/loss.py:30-48 (CrossEntropy + MONAI DiceLoss(include_background=False,
 * to_onehot_y=True, softmax=True) per deep output), weights utils/runtime.py:125-144.  
 *   L = sum_i w_i [ CE(logits_i, y) + mean_{b, c>=1} (1 - (2 I + 1e-5) / (P + T + 1e-5)) ]
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t n_out;       /* deep outputs (<= 8), all already at label resolution    */
  int32_t B, C, S;     /* batch, classes (2..4), voxels                            */
  float weights[8];    /* w_i (already normalised)                                 */
} vx_segloss_desc;
size_t vx_segloss_workspace(const vx_segloss_desc* d);
/* fwd in: logits_0..logits_{n-1} (B,C,S), labels (B,1,S) int64     out: loss (1), sums (n, B, 1 + 3C)
 * bwd in: dloss (1), logits_0.., labels, sums                       out: dlogits_0..dlogits_{n-1}           */
int vx_segloss_fwd(const vx_segloss_desc* d, const void* const* in, void* const* out, void* workspace,
                   size_t workspace_bytes, vx_stream_t stream);
int vx_segloss_bwd(const vx_segloss_desc* d, const void* const* in, void* const* out, vx_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * Network-input stem -- monai PatchEmbed as used by model/Encoder.py:150-156 (Conv3d k = s = patch on one modality's
 * channels of the input).  The input is addressed inside the full (B, C_in_total, D, H, W)
 * tensor; there is no data gradient (it is the network input).
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t B, C_in_total, c_in_off, C_in, C_out, patch, D, H, W;
} vx_patch_embed_desc;
/* fwd in: x (B,C_in_total,D,H,W), w (C_out,C_in,p,p,p), bias (C_out) or NULL    out: y (B,C_out,D/p,H/p,W/p)
 * bwd in: dy, x                                                                 out: dw, db                  */
int vx_patch_embed_fwd(const vx_patch_embed_desc* d, const void* const* in, void* const* out, vx_stream_t stream);
int vx_patch_embed_bwd(const vx_patch_embed_desc* d, const void* const* in, void* const* out, vx_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * Bias + 3-D PixelShuffle -- model/components/superpixel.py:15 behind `out_conv1` / the reconstruction `out_conv`
 * (model/Decoder.py:73-76,150-153): y[b, c, d s + s1, h s + s2, w s + s3] = z[b, ((c s + s1) s + s2) s + s3, d, h, w] + bias.
 * The convolution in front runs without its bias (the layout half).
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t B, C, scale, d, h, w;    /* C = channels after the shuffle; z has C * scale^3 channels, extent (d, h, w) */
} vx_pixel_shuffle_desc;
/* fwd in: z, bias (C * scale^3) or NULL   out: y (B, C, d s, h s, w s)          bwd in: dy   out: dz, db or NULL */
int vx_pixel_shuffle_fwd(const vx_pixel_shuffle_desc* d, const void* const* in, void* const* out, vx_stream_t stream);
int vx_pixel_shuffle_bwd(const vx_pixel_shuffle_desc* d, const void* const* in, void* const* out, vx_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * Convolutions of the glue layers around the hot path, all in fp32 accuracy:
 *   kernel 3, stride 1, pad 1, C_in 16   decoder.out_conv1 / reconstruction out_conv, model/Decoder.py:73-76,150-153
 *                                        -> tcgen05 implicit GEMMs (3xTF32 = fp32-accurate), csrc/conv3_tc.cu; with
 *                                        shuffle = 4 the PixelShuffle(4) of components/superpixel.py:15 and the bias are part
 *                                        of the kernel (y is (B, C_out/64, 4D, 4H, 4W); dy of that shape in backward)
 *   kernel 2p-1, stride p, pad p-1       DownConv.down, model/components/conv_blocks.py:10-17 (p = 4: k7 s4, p = 2: k3 s2)
 *   transposed, kernel = stride = 2      UpConv.up (ConvTranspose3d), model/components/conv_blocks.py:31-35
 * (D, H, W) is the INPUT extent.  Weights in torch layout: (C_out, C_in, k, k, k), transposed: (C_in, C_out, k, k, k).
 *   fwd  in[0] x   in[1] w   in[2] bias or NULL                     out[0] y
 *   bwd  in[0] dy  in[1] x   in[2] w      out[0] dx or NULL (no data gradient wanted)  out[1] dw  out[2] db or NULL
 * Any other geometry returns VX_ERR_UNSUPPORTED (there is no library fallback behind this ABI).
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t B, C_in, C_out, D, H, W;
  int32_t kernel, stride, pad;
  int32_t transposed;              /* 1: ConvTranspose3d with kernel == stride, pad 0 */
  int32_t shuffle;                 /* dense k3 only: 0, or 4 = PixelShuffle(4) fused (C_out % 64 == 0) */
} vx_conv_desc;
size_t vx_conv_workspace(const vx_conv_desc* d);
int vx_conv_fwd(const vx_conv_desc* d, const void* const* in, void* const* out, void* workspace, size_t workspace_bytes,
                vx_stream_t stream);
int vx_conv_bwd(const vx_conv_desc* d, const void* const* in, void* const* out, void* workspace, size_t workspace_bytes,
                vx_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * AdamW step over every parameter tensor in one launch (torch.optim.AdamW semantics: decoupled weight decay, bias
 * correction; the reference's optimiser, utils/runtime.py).  SURVEY.md section 8f row 4.
 *   in[0]  int64 table (n_tensors, 4): parameter pointer, gradient pointer (0 = no gradient: skipped), offset of the tensor
 *          in the flat moment buffers, element count
 *   in[1]  int32 (n_chunks, 2): tensor index, first element -- one CTA per 1024-element chunk
 *   in[2]  (only when hyper_on_device != 0) 2 floats on the device: learning rate, weight decay -- read by the kernel at run
 *          time, so a learning-rate schedule takes effect inside a replayed CUDA graph (utils/train_*.py schedulers)
 *   out[0], out[1]  flat exp_avg / exp_avg_sq (fp32);  out[2]  step counter (1 float, advanced by the call)
 *   grad_scale  the gradient is multiplied by it on the way in (1 / world_size after an all-reduce(sum)); 0 means 1
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t n_chunks;
  float lr, beta1, beta2, eps, weight_decay;
  float grad_scale;
  int32_t hyper_on_device;
} vx_adamw_desc;
int vx_adamw_step(const vx_adamw_desc* d, const void* const* in, void* const* out, vx_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VELOXSEG_ABI_H_ */
