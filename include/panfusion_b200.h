/*
 * panfusion_b200 — C ABI of the B200 (sm_100a) denoise hot path of PanFusion.
 *
 * Every entry point takes plain device pointers, sizes and a CUDA stream (cudaStream_t passed as void*).
 * The caller (PyTorch on the host side) owns every buffer; the library keeps no hidden state except a
 * thread-local error string. All functions return 0 on success or a negative PF_ERR_* code; launches are
 * asynchronous on the given stream and CUDA-graph capturable (no allocation, no host sync inside).
 *
 * Each declaration cites the reference interface (file:line under the upstream repo) it replaces.
 * Tensors are row-major; "tokens" layout means [N*H*W, C] channels-last, "NCHW" is the reference's layout.
 */
#ifndef PANFUSION_B200_H
#define PANFUSION_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PF_OK 0
#define PF_ERR_INVALID (-1)     /* bad argument (message in pf_last_error) */
#define PF_ERR_CUDA (-2)        /* CUDA runtime / driver failure */
#define PF_ERR_UNSUPPORTED (-3) /* shape / dtype not supported by the sm_100a kernels */

typedef enum { PF_F32 = 0, PF_F16 = 1, PF_BF16 = 2 } pf_dtype;

/* error convention: reference raises Python exceptions (utils/pano.py:85,
 * external/Perspective_and_Equirectangular/utils.py:15); here: return code + message. */
const char* pf_last_error(void);
int pf_version(void);
/* returns 0 iff the current device is compute capability 10.x (B200); the library has no other path */
int pf_check_device(void);

/* ------------------------------------------------------------------------------------------------
 * Spherical resampling (K1/K2): external/Perspective_and_Equirectangular/e2p.py:54-76, p2e.py:52-77
 * The sampling grid is computed in-kernel (fp64 per output pixel) from the camera record; no grid is
 * stored. Sampling follows kornia.remap -> F.grid_sample(align_corners=True, padding_mode='zeros').
 *
 * Camera record: PF_CAM_DOUBLES doubles, device memory, one per batch element (cam_stride = 1) or a single
 * record broadcast over the batch (cam_stride = 0):
 *   [0:9)  R1  row-major   (e2p: Rodrigues(z * rad(theta));   p2e: inv(R1))
 *   [9:18) R2  row-major   (e2p: Rodrigues(R1 y * rad(-phi)); p2e: inv(R2))
 *   [18] w_len = tan(rad(wfov/2))   [19] h_len = tan(rad(hfov/2))
 * mode: 0 = bilinear, 1 = nearest (round-half-even).
 * ------------------------------------------------------------------------------------------------ */
#define PF_CAM_DOUBLES 20

/* e2p(e_img[B,C,He,We]) -> [B,C,h,w]   (e2p.py:54-76; grid math e2p.py:9-51) */
int pf_e2p(const void* src, void* dst, int dtype, int B, int C, int He, int We, int h, int w,
           const double* cams, int cam_stride, int mode, void* stream);

/* e2p of B cameras over B / src_repeat source panoramas: src[B/src_repeat, C, He, We], camera b reads source b / src_repeat
 * (PanFusion.init_noise, models/pano/PanFusion.py:30-43, expands ONE panorama to its m views before e2p; feature-map
 * warps of a CFG batch read 2 panoramas from 16 cameras). Same result as pf_e2p on the expanded tensor; the source is read
 * from HBM once. */
int pf_e2p_shared(const void* src, void* dst, int dtype, int B, int src_repeat, int C, int He, int We, int h, int w,
                  const double* cams, int cam_stride, int mode, void* stream);

/* py360convert.e2p (external/py360convert/e2p.py:6-43, utils.py:104-132,231-243): the dataset's pixel-space convention
 * (utils/pano.py:160-161 `Equirectangular.to_perspective`, dataset/PanoDataset.py:138) — channels-last src[H, W, C] (uint8 when
 * is_u8, else fp32) -> dst[num_cams, h, w, C]; half-pixel centres (uv2coor), longitude wrap-around, pole rows padded with the
 * first / last row rolled by W/2, scipy 'wrap' boundaries, float64 grid math, integers rounded half up.
 * Camera record: PF_CAM360_DOUBLES doubles = Rx, Ry, Ri (row-major; rotation_matrix(v,[1,0,0]), rotation_matrix(-yaw,[0,1,0]),
 * rotation_matrix(in_rot, z Rx Ry)), tan(h_fov/2), tan(v_fov/2). mode: 0 bilinear, 1 nearest; anything else is
 * PF_ERR_UNSUPPORTED like the reference's NotImplementedError('unknown mode'). */
#define PF_CAM360_DOUBLES 29
int pf_e2p_py360(const void* src, void* dst, int is_u8, int H, int W, int C, int h, int w, const double* cams,
                 int num_cams, int mode, void* stream);

/* p2e(p_img[B,C,hp,wp]) -> equi[B,C,He,We] (already multiplied by mask), mask[B,1,He,We] uint8 (may be NULL)
 * (p2e.py:52-77; grid math p2e.py:9-49) */
int pf_p2e(const void* src, void* dst, uint8_t* mask, int dtype, int B, int C, int hp, int wp, int He, int We,
           const double* cams, int cam_stride, int mode, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Tap-GEMM (tcgen05 / TMEM / TMA): the one dense-contraction engine behind nn.Linear (K8, K10) and the
 * 3x3 / 1x1 convolutions (K9, K11) of the UNet walk in models/pano/MVGenModel.py:85-295.
 *
 *   acc[m, n] = sum_{t < num_taps} sum_{k < Kc} A[m + tap_off[t], k] * B[n, t*Kc + k]      (fp32 accumulate)
 *   v        = act(acc + bias[n] + rowbias[group(m), n])           (GEGLU: v = a * gelu(g), see below)
 *   out[row(m), n] = v + residual[row(m), n]
 *
 * A is a row-major [a_rows, a_ld] 16-bit matrix (rows outside [0, a_rows) read as zero), B the packed
 * weight [N, num_taps*Kc]. A 3x3 convolution is 9 taps over a zero-haloed channels-last image
 * ("padded-flat" layout); the M-space -> output-row map drops halo rows:
 *   map_mode 0: row(m) = m, group(m) = m / rows_per_group
 *   map_mode 1: m -> (img, i, j) with i = (m / Wm) % Hm, j = m % Wm; valid iff i0 <= i < i0+Hout and
 *               j0 <= j < j0+Wout; row(m) = (img*Hout + i-i0)*Wout + (j-j0); group(m) = img
 * act: PF_ACT_GEGLU expects B (and bias) packed so that every block_n-wide column tile holds block_n/2 value
 * columns followed by their block_n/2 gate columns; it writes N/2 output columns.
 * Constraints: Kc % 64 == 0, N % block_n == 0, block_n in {64,128,160,256}, a_ld/b_ld % 8 == 0.
 * ------------------------------------------------------------------------------------------------ */
enum { PF_ACT_NONE = 0, PF_ACT_SILU = 1, PF_ACT_GELU = 2, PF_ACT_GEGLU = 3 };
#define PF_MAX_TAPS 16

typedef struct pf_gemm_args {
  const void* A;
  int64_t a_rows;
  int32_t a_ld;
  const void* B;
  int32_t b_ld;
  int32_t dtype; /* PF_F16 | PF_BF16 (A and B) */
  int32_t M, N, Kc, num_taps;
  int32_t tap_off[PF_MAX_TAPS];
  int32_t block_n; /* 0 = auto */
  void* out;
  int32_t out_ld;
  int32_t out_dtype; /* PF_F32 or same as dtype */
  const float* bias;
  const float* rowbias;
  int32_t rowbias_ld;
  int32_t rows_per_group;
  const void* residual;
  int32_t res_ld;
  int32_t res_dtype;
  int32_t act;
  int32_t map_mode, Hm, Wm, i0, j0, Hout, Wout;
  /* split-K (long-K, few-tile problems: the 8x8 / 16x16 level convolutions): k_splits > 1 partitions the K-slabs over
   * k_splits CTAs per tile; partial accumulators go to splitk_ws (fp32, k_splits * M * N elements, caller-owned) and
   * a second kernel reduces them in a fixed order and applies the epilogue. 0 / 1 = off. */
  int32_t k_splits;
  float* splitk_ws;
  /* LayerNorm fused around the GEMM (diffusers BasicTransformerBlock norm1/2/3 -> to_q|k|v / to_q / GEGLU proj, and
   * models/modules/transformer.py:159-160 norm2 -> ff): instead of a LayerNorm kernel between two linears,
   *   PRODUCER (row_stats_out != NULL): besides `out`, writes per output row the partial (sum, sum of squares) of the
   *     final fp32 values of every column slot: row_stats_out[m][slot][2], slots = pf_gemm_row_stats_slots(args).
   *   CONSUMER (ln_stats != NULL): A is the UN-normalised tensor, B holds gamma-scaled weights W'[n,k] = gamma[k] W[n,k],
   *     ln_colsum[n] = sum_k W'[n,k] (of the 16-bit rounded W'), bias[n] = sum_k beta[k] W[n,k] + b[n]; the epilogue applies
   *     acc <- rstd[m] * (acc - mean[m] * ln_colsum[n]) with mean / rstd from ln_stats[m][ln_slots][2] over K = Kc*num_taps
   *     elements, biased variance, eps = ln_eps — algebraically LayerNorm(A) W^T + b with the normalised tensor never
   *     stored. Supported with the plain row map and 16-bit output, or with the GEGLU epilogue. */
  /* map_mode 1 with an output SCATTER (0 / 1 = off): the valid pixel (i, j) of image img is written to row
   *   ((img*Hout + i-i0)*out_sy + out_a) * (Wout*out_sx) + (j-j0)*out_sx + out_b
   * i.e. phase (out_a, out_b) of an image up-sampled by (out_sy, out_sx). Upsample2D's nearest-x2 followed by a 3x3
   * convolution (MVGenModel.py:272-277) is exactly four 2x2 convolutions of the ORIGINAL image, one per output phase, with
   * summed weights: 16 instead of 36 tap-GEMMs per input pixel and no up-sampled copy. */
  int32_t out_sy, out_sx, out_a, out_b;
  float* row_stats_out;
  const float* ln_stats;
  int32_t ln_slots;
  const float* ln_colsum;
  float ln_eps;
} pf_gemm_args;

int pf_gemm_taps(const pf_gemm_args* args, void* stream);
/* number of column slots a producer with these args writes per row of row_stats_out: 2 per column tile, and a producer's
 * tile width is a function of N alone (requests / tuning are ignored), so the statistics are bit-identical for any M */
int pf_gemm_row_stats_slots(const pf_gemm_args* args);
/* suggested k_splits for this problem (1 = do not split); only M, N, Kc, num_taps, act, map_mode, dtypes are read */
int pf_gemm_splitk_plan(const pf_gemm_args* args);
/* block_n the auto-tuner would pick for this N (used by the host-side weight packer for GEGLU) */
int pf_gemm_pick_block_n(int N, int act);

/* ------------------------------------------------------------------------------------------------
 * Flash attention forward (tcgen05): out = softmax(q k^T * scale + bias) v, fp32 softmax / accumulation.
 * Replaces xformers.ops.memory_efficient_attention at models/modules/transformer.py:71 (EPPA: head_dim 32,
 * additive fp32 bias shared by every head — transformer.py:68 materialises it per head, this never does) and the
 * bmm-softmax-bmm of the diffusers attention blocks walked at models/pano/MVGenModel.py:104,116,185,190,227,241
 * (head_dim 64, self and 77-token text cross attention).
 * q/k/v: 16-bit, element (b, l, h, d) at ptr[b*bstride + l*ld + h*head_dim + d] (so fused QKV buffers work
 * in place); out: [B, Lq, out_ld] same dtype. bias: fp32 [*, Lq, bias_ld], batch b reads bias + b*bias_bstride
 * (0 = one table for all batches), or NULL. Any Lq / Lk >= 1 (ragged tiles are masked).
 * ------------------------------------------------------------------------------------------------ */
typedef struct pf_fmha_args {
  const void* q;
  const void* k;
  const void* v;
  void* out;
  int32_t dtype;
  int32_t B, H, Lq, Lk, head_dim;
  int32_t q_ld, k_ld, v_ld, out_ld;
  int64_t q_bstride, k_bstride, v_bstride;
  float scale;
  const float* bias;
  int64_t bias_bstride;
  int32_t bias_ld;
  /* optional block-sparsity hint for the bias: flags[b][ceil(Lq/128)][ceil(Lk/64)] bytes, non-zero = every entry of
   * that 128x64 bias tile is exactly -1 (no geometric correspondence, SURVEY.md A.4: ~85% of the EPPA tiles), so the
   * kernel substitutes the constant instead of reading the tile. NULL = read everything. Build with
   * pf_bias_tile_flags. */
  const uint8_t* bias_flags;
  int64_t flags_bstride;
  int32_t flags_ld;
  /* tile-PACKED bias (the resident form of the EPPA tables): `bias` is then the store [n_live][128 * 64] fp32 (lane-interleaved tiles) built by
   * pf_bias_tile_pack and bias_tile_off[bias batch][ceil(Lq/128)][ceil(Lk/64)] (row stride flags_ld, batch stride
   * flags_bstride, in tiles) holds each tile's index in it, or -1 for a tile that is entirely -1 (no correspondence).
   * bias_ld / bias_bstride / bias_flags are unused. The reference materialises the dense tensor PER HEAD
   * (models/modules/transformer.py:68); this keeps ~15 % of ONE copy. */
  const int32_t* bias_tile_off;
} pf_fmha_args;

int pf_fmha_fwd(const pf_fmha_args* args, void* stream);
/* flags[g][qt][kt] = 1 iff bias[g][qt*128 .. , kt*64 ..] is entirely == -1.0f (tiles clipped at Lq / Lk) */
int pf_bias_tile_flags(const float* bias, int G, int Lq, int Lk, int bias_ld, int64_t bias_bstride, uint8_t* flags,
                       void* stream);
/* exclusive scan of the live (flag == 0) tiles: tile_off[t] = index among the live tiles or -1; n_live[0] = their number */
int pf_bias_tile_scan(const uint8_t* flags, int num_tiles, int32_t* tile_off, int32_t* n_live, void* stream);
/* packed[tile_off[g][qt][kt]] <- the dense 128 x 64 tile (zero outside [Lq, Lk]); constant tiles are skipped. Inside a
 * tile the 8192 floats are LANE-INTERLEAVED for the attention kernel (thread = query row): element (r, c) sits at
 * (((r / 32) * 16 + c / 4) * 32 + r % 32) * 4 + c % 4, so one 16-byte load of a warp covers 512 contiguous bytes. */
int pf_bias_tile_pack(const float* bias, int G, int Lq, int Lk, int bias_ld, int64_t bias_bstride, const int32_t* tile_off,
                      float* packed, void* stream);

/* ------------------------------------------------------------------------------------------------
 * GroupNorm statistics of a channels-last image x[N, H, W, C] (row stride ld), computed over the image
 * circularly extended by `circ` columns on each side — the reference normalises the padded tensor
 * (models/pano/MVGenModel.py:110-115 wraps each panorama ResnetBlock2D in utils/pano.py:74-105 pad/unpad), so
 * columns {0..circ-1, W-circ..W-1} count twice. ws: scratch of pf_groupnorm_ws_floats(N, groups) floats.
 * counters: N ints that are ZERO on entry (the kernel restores them to zero; concurrent calls need distinct slots):
 * the last CTA of each image reduces the partial sums in a fixed order, so the result is deterministic.
 * mean_rstd: [N, groups, 2] fp32 (mean, 1/sqrt(var + eps)), biased variance like torch.nn.GroupNorm.
 * ------------------------------------------------------------------------------------------------ */
int pf_groupnorm_ws_floats(int N, int groups);
int pf_groupnorm_stats(const void* x, int dtype, int N, int H, int W, int C, int ld, int groups, int circ,
                       float eps, float* ws, int* counters, float* mean_rstd, void* stream);

/* GroupNorm-apply (+SiLU) fused with building the tap-GEMM A operand (replaces norm+nonlinearity of diffusers
 * ResnetBlock2D / Transformer2DModel.norm, pad_pano/unpad_pano utils/pano.py:74-105, Upsample2D's nearest x2,
 * Downsample2D's stride-2 gather; call sites MVGenModel.py:98-277).
 *   source image S = x circularly extended by `circ` columns per side, then nearest-upsampled by `up` (1|2);
 *   phases == 1: out[N, Hu + 2*halo, Wu + 2*halo, C], zero halo (halo = 1 for a 3x3 conv input, 0 = plain apply)
 *   phases == 4: out[4][N][Hu/2 + 1][Wu/2 + 1][C], phase (py,px) element (i,j) = zero-padded S at (2i+py, 2j+px)
 * mean_rstd == NULL skips the normalisation; act: PF_ACT_NONE | PF_ACT_SILU. */
int pf_conv_prep(const void* x, void* out, int dtype, int N, int H, int W, int C, int ld, const float* mean_rstd,
                 const float* gamma, const float* beta, int groups, int act, int circ, int up, int phases, int halo,
                 void* stream);

/* pf_groupnorm_stats + pf_conv_prep in ONE launch (same semantics, same reference call sites), optionally over the channel
 * concatenation of two tensors (torch.cat([hidden, skip], dim=1) at MVGenModel.py:223,231,246,254):
 *   x = cat(x1[N*H*W, C1], x2[N*H*W, C2]) (x2 may be NULL); if cat_out != NULL the raw concatenation [N*H*W, C1+C2] is also
 *   written (the ResnetBlock2D shortcut convolution reads it);
 *   statistics over the image circularly extended by circ_stats columns (duplicated columns count twice), applied with
 *   gamma / beta (+ SiLU) while building the conv_prep layout (circ, up, phases, halo as in pf_conv_prep).
 * The kernel holds a per-image barrier between the statistics and the apply phase (the source is re-read from L2), so its
 * grid is capped at 148 CTAs of <= 512 threads — two such launches (the two UNet branches' streams) are always co-resident.
 * schedule: 1 = that fused launch, 2 = two launches (statistics with one CTA per slab, then the apply pass: no barrier, many
 * more CTAs per image — measured faster at every batch size of the denoise step), 0 = default (two launches unless
 * PF_GN_FUSED_MIN_N says otherwise). Same bits either way.
 * ws: pf_gn_prep_ws_floats(N, groups) floats of scratch; sync: 3*N ints that are ZERO on entry (restored to zero by the
 * kernel; concurrent launches need distinct slots). The partition of every sum depends on H*W only, never on N: results are
 * bit-identical for any batch size. */
int pf_gn_prep_ws_floats(int N, int groups);
int pf_gn_prep(const void* x1, int ld1, int C1, const void* x2, int ld2, int C2, void* cat_out, void* out, int dtype,
               int N, int H, int W, int groups, float eps, const float* gamma, const float* beta, int act,
               int circ_stats, int circ, int up, int phases, int halo, int schedule, float* ws, int* sync, void* stream);

/* out[t, :] = LayerNorm(x[t, :] + pe[t % pe_rows, :]) * gamma + beta (pe fp32, may be NULL);
 * models/modules/transformer.py:157-160 (EPPA norm1 on x + query_pe / context, norm2) and the diffusers
 * BasicTransformerBlock norm1/2/3. */
int pf_layernorm(const void* x, int ldx, void* out, int ldo, int dtype, int T, int C, const float* pe, int pe_rows,
                 const float* gamma, const float* beta, float eps, void* stream);

/* conv_in (MVGenModel.py:85-91): NCHW fp32 latent [N,Cin,H,W] -> tokens [N*H*W, Cout] 16-bit; 3x3 pad 1;
 * circ != 0 wraps columns (== pad_pano(1) -> conv -> unpad_pano(1)). w fp32 [Cout,Cin,3,3], bias fp32 [Cout].
 * act = PF_ACT_SILU applies SiLU to the result: the first convolution of the ControlNet conditioning embedding on the
 * 3-channel layout image (diffusers ControlNetConditioningEmbedding [3P], consumed at MVGenModel.py:66-83). */
int pf_conv_in(const float* x, const float* w, const float* bias, void* out, int dtype, int N, int Cin, int H, int W,
               int Cout, int circ, int act, void* stream);

/* conv_out (MVGenModel.py:279-295) over the PREPARED tensor: xp = pf_conv_prep(conv_norm_out statistics of the
 * un-padded tensor as the reference does at :288, SiLU, circ, halo = 1) of shape [N, H+2, W+2*circ+2, C] 16-bit
 * -> NCHW fp32 [N, Cout<=4, H, W]; circ = 1 for the panorama (pad_pano(1) -> conv -> unpad_pano(1)). */
int pf_conv_out(const void* xp, int dtype, const float* w, const float* bias, float* out, int N, int H, int W, int C,
                int Cout, int circ, void* stream);

/* strided 2-D copy of 16-bit rows (skip concatenation, torch.cat at MVGenModel.py:223,231,246,254) */
int pf_copy2d(const void* src, int src_ld, void* dst, int dst_ld, long long rows, int cols, void* stream);

/* pad_pano (utils/pano.py:74-99): out[r, j] = x[r, (j - pad) mod W] for j in [0, W + 2*pad) — circular padding of the
 * longitude axis of a contiguous tensor whose leading dims are flattened into `rows`; any pad >= 1 (F.pad's circular
 * mode limits pad <= W, this does not). elem_bytes in {1, 2, 4, 8}. */
int pf_pad_pano(const void* x, void* out, int elem_bytes, long long rows, int W, int pad, void* stream);

/* out[r, :cols] = softmax(scale * s[r, :cols]) — fp32 logits [rows, ld] -> 16-bit probabilities [rows, ldo].
 * The VAE mid-block attention (diffusers AutoencoderKL [3P], called through decode_latent, PanoGenerator.py:272-278)
 * has ONE head of width 512, outside the flash kernel's head sizes: it runs as pf_gemm_taps (Q K^T, fp32 out) ->
 * pf_softmax_rows -> pf_gemm_taps (P V). */
int pf_softmax_rows(const float* s, long long ld, void* out, long long ldo, int dtype, long long rows, int cols,
                    float scale, void* stream);

/* tensor_to_image (models/modules/utils.py:9-15): x fp32 [n, C, H, W] in [-1, 1] -> uint8 [n, H, W, C] =
 * round(clamp(x / 2 + 0.5, 0, 1) * 255), round-half-to-even like torch.round. */
int pf_tensor_to_image(const float* x, unsigned char* out, long long n, int C, int H, int W, void* stream);

/* diffusers Timesteps(dim, flip_sin_to_cos=True, freq_shift=0) (MVGenModel.py:55,59): t fp32 [n] -> [n, dim] */
int pf_timestep_embed(const float* t, void* out, int dtype, int n, int dim, void* stream);

/* CLIP text embeddings (transformers CLIPTextEmbeddings [3P] behind PanoGenerator.encode_text, models/pano/PanoGenerator.py:
 * 197-211): out[t, :] = tok_emb[ids[t], :] + pos_emb[t % L, :] (fp32 tables -> 16-bit tokens) and row_stats[t] = (sum, sum of
 * squares, 0, 0) of the stored row — the two-slot statistics the first encoder layer's fused LayerNorm consumes
 * (pf_gemm_args.ln_stats). ids int64 [T]; an id outside [0, vocab) traps (torch raises IndexError). */
int pf_embed_tokens(const long long* ids, const float* tok_emb, const float* pos_emb, void* out, int dtype,
                    float* row_stats, int T, int L, int C, int vocab, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Forward half of the training step (models/pano/PanFusion.py:64-98). The backward of the UNets, of
 * EPPA and of the LoRA adapters, the optimizer and the gradient all-reduce are NOT built.
 * pf_add_noise: diffusers SchedulerMixin.add_noise [3P] as called at PanFusion.py:84-85 —
 *   out[b, :] = sqrt(abar[t[b]]) * x0[b, :] + sqrt(1 - abar[t[b]]) * noise[b, :]   (fp32 [B, per_sample], t int64 [B],
 *   abar = alphas_cumprod fp32 [num_train_timesteps]; a timestep outside the table traps).
 * pf_mse_loss: torch.nn.functional.mse_loss(a, b) with mean reduction (PanFusion.py:92-93), out[0] = sum((a-b)^2) / n,
 *   summed in a launch-independent order (fixed chunks, partials added in index order in fp64). ws: pf_mse_loss_ws_floats()
 *   floats; counter: one zero-initialised int32, re-armed by the kernel.
 * ------------------------------------------------------------------------------------------------ */
/* Latent sampling at the end of the VAE encoder (diffusers DiagonalGaussianDistribution.sample [3P] + the scaling of
 * PanoGenerator.encode_image, models/pano/PanoGenerator.py:218-224): moments = channels-last rows of width ld holding
 * [mean (L) | logvar (L) | ...] for N images of HW pixels; eps, out NCHW fp32 [N, L, HW];
 * out = (mean + exp(0.5 * clamp(logvar, -30, 20)) * eps) * scale. */
int pf_gaussian_sample(const float* moments, int ld, const float* eps, float* out, int N, int L, int HW, float scale,
                       void* stream);
int pf_add_noise(const float* x0, const float* noise, float* out, const long long* t, const float* alphas_cumprod,
                 int num_train_timesteps, int B, long long per_sample, void* stream);
int pf_mse_loss_ws_floats(void);
int pf_mse_loss(const float* a, const float* b, long long n, float* ws, int* counter, float* out, void* stream);

/* Classifier-free-guidance combine + DDIM update (+ roll of the result by `roll` columns):
 *   e = eps[0:count] + guidance * (eps[count:2count] - eps[0:count])        (PanoGenerator.py:253-262)
 *   out[.., (col+roll) % W] = sqrt(a_prev) * (x - sqrt(1-a_t) e) / sqrt(a_t) + sqrt(1-a_prev) e   (DDIM, eta 0)
 * replaces combine_cls_free_guide_pred + DDIMScheduler.step (PanFusion.py:159-162) + torch.roll
 * (PanoGenerator.py:264-269) of the following step. x, out fp32 [count] viewed as rows of W. */
int pf_cfg_ddim_step(const float* x, const float* eps, float* out, long long count, int W, int roll, float guidance,
                     float alpha_t, float alpha_prev, void* stream);
/* same update with the two coefficients read from DEVICE memory: coef[0] = sqrt(a_prev/a_t),
 * coef[1] = sqrt(1-a_prev) - coef[0]*sqrt(1-a_t) — lets one captured CUDA graph serve every denoising step. */
int pf_cfg_ddim_step_dev(const float* x, const float* eps, float* out, long long count, int W, int roll,
                         float guidance, const float* coef, void* stream);

/* ------------------------------------------------------------------------------------------------
 * EPPA geometry tables (models/pano/utils.py:10-106, models/modules/transformer.py:185-201).
 * pf_eppa_tables: correspondence bias for both attention directions of V views (groups of m views per batch
 * element) straight from the camera records (device, PF_CAM_DOUBLES each; e2p and p2e flavours):
 *   bias1[V/m][eh*ew][m*ph*pw]  query = pano pixel, keys = view pixels   (modules.py:46)
 *   bias2[V/m][m*ph*pw][eh*ew]  query = view pixel, keys = pano pixels   (modules.py:53)
 * blur5: HOST pointer to the 5 taps of the sigma-1 gaussian. ws_idx / ws_w: scratch of 4*V*(ph*pw+eh*ew) each.
 * pf_eppa_pe: SphericalPE tables, fp32: pers_pe[V*ph*pw][4*n_freqs], equi_pe[eh*ew][4*n_freqs].
 * ------------------------------------------------------------------------------------------------ */
int pf_eppa_tables(const double* cams_e2p, const double* cams_p2e, int V, int m, int ph, int pw, int eh, int ew,
                   const float* blur5, int* ws_idx, float* ws_w, float* bias1, float* bias2, void* stream);
int pf_eppa_pe(const double* cams_e2p, int V, int ph, int pw, int eh, int ew, const float* freq_bands, int n_freqs,
               float* pers_pe, float* equi_pe, void* stream);

/* ------------------------------------------------------------------------------------------------
 * View-sharded step: device-initiated all-gather over NVLink peer mappings, capturable in a CUDA graph.
 * The reference has no model parallelism (Lightning DDP over prompts, main.py:63); this is the exchange the view partition
 * needs: models/pano/modules.py:44-48 lets every panorama query attend to the keys / values of ALL views, so each rank
 * publishes the projected K|V of its local views before every EPPA block (and the eps outputs at the end of the step).
 *   local         this rank's slice (slice_bytes, multiple of 16)
 *   peer_data     DEVICE array [nranks] of pointers: every rank's receive buffer [nranks][slice_bytes] for this call site,
 *                 mapped into this process (CUDA IPC / peer access); entry `rank` is the own buffer
 *   peer_flags    DEVICE array [nranks] of pointers to every rank's nranks flag words (zero-initialised, uint32)
 *   my_flags      == peer_flags[rank];  state: 2 zero-initialised uint32 of this rank (epoch, CTA counter)
 * On return (stream order) the own receive buffer holds every rank's slice in rank order. One kernel: push to all peers,
 * publish the epoch, wait for all peers; a peer that never arrives makes the kernel trap after ~15 s instead of hanging. */
/* Receive buffers of the device-side collectives: cudaMalloc'ed (zero-filled) by pf_comm_alloc, exported as a 64-byte CUDA IPC
 * handle, opened by the peers ON THEIR device (cudaIpcOpenMemHandle with lazy peer access — NVLink on an HGX board). */
int pf_comm_alloc(long long bytes, void** ptr);
int pf_comm_free(void* ptr);
int pf_ipc_export(const void* ptr, unsigned char handle[64]);
int pf_ipc_open(const unsigned char handle[64], void** ptr);
int pf_ipc_close(void* ptr);
/* let kernels of the CURRENT device store into memory of `peer_device` (cudaDeviceEnablePeerAccess; idempotent) — needed once
 * per peer before pf_allgather_views pushes into IPC-mapped buffers that live on the other GPUs of the node */
int pf_enable_peer_access(int peer_device);
int pf_allgather_views(const void* local, long long slice_bytes, void* const* peer_data, unsigned int* const* peer_flags,
                       unsigned int* my_flags, unsigned int* state, int rank, int nranks, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PANFUSION_B200_H */
