/* posetraj_b200 — C ABI of the B200-native PoseTraj denoising hot path.
 *
 * This is the drop-in boundary.  The reference has no FFI layer: the path is
 * plain Python (`ControlNetSDVModel.forward`, `UNetSpatioTemporalConditionControlNetModel.forward`,
 * `EulerDiscreteScheduler.step`, the loop in `StableVideoDiffusionPipelineControlNet.__call__`).
 * The Python mirror of those classes (posetraj_b200/*.py) lowers every forward onto the entry
 * points below; INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types.
 *   - Every device buffer (inputs, outputs, workspaces) is owned by the caller; the library never
 *     allocates or frees device memory and keeps no reference past the call.
 *   - All launches are stream-ordered on the `stream` argument (a cudaStream_t passed as void*).
 *   - Return value: 0 on success, otherwise a cudaError_t-compatible code (argument errors return
 *     cudaErrorInvalidValue = 1); `pt_last_error()` returns a thread-local message.
 *   - Activations are NHWC / token-major bf16 (`[rows, C]`, row = ((b*F+f)*H + y)*W + x) unless stated.
 */
#ifndef POSETRAJ_B200_H_
#define POSETRAJ_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------ */
/* library                                                                                    */
/* ------------------------------------------------------------------------------------------ */
const char* pt_last_error(void);
int pt_version(void);
/* 1 when kernels are launched with programmatic dependent launch (opt-in: library built with -DPT_ENABLE_PDL and
 * PT_PDL=1 in the environment; 0 in the default build) */
int pt_pdl(void);
/* number of kernels launched by this library in this process since load (bench.py's gpu_launches) */
int64_t pt_launch_count(void);
/* sizeof() of a struct declared in this header, by name (-1 if unknown): ABI self-check for bindings */
int pt_sizeof(const char* name);

/* 128-byte opaque TMA descriptor (CUtensorMap); must be 64-byte aligned in host memory. */
typedef struct PtTensorMap {
  uint64_t opaque[16];
} PtTensorMap;

/* Encode a bf16 tiled tensor map with SWIZZLE_128B and zero out-of-bounds fill.
 *   rank 2 or 3; dims[0] is the contiguous dimension; strides_bytes[i] is the stride of dims[i+1].
 *   box[0] must be 64 (one 128-byte swizzle row). */
int pt_tensormap_encode_bf16(PtTensorMap* out, const void* base, int rank, const uint64_t* dims,
                             const uint64_t* strides_bytes, const uint32_t* box);

/* ------------------------------------------------------------------------------------------ */
/* P7 + P8: fused CFG combine + v-prediction Euler step + next-step model input               */
/* replaces pipeline/pipeline_stable_video_diffusion_controlnet.py:532-537,567-572 and        */
/* utils/scheduling_euler_discrete_karras_fix.py:264-288,418-528                              */
/* ------------------------------------------------------------------------------------------ */
typedef struct PtCfgEulerArgs {
  const void* noise_pred;   /* [2, F, HW, ldp] bf16 token-major (row 0 uncond, row 1 cond), or fp32 NCHW if pred_nchw_f32 */
  int32_t pred_ld;          /* channel stride of noise_pred rows (token-major mode) */
  int32_t pred_nchw_f32;    /* 1: noise_pred is [2,F,C,H,W] fp32 contiguous (reference layout) */
  float* latents;           /* [F, C, H, W] fp32, updated in place (the reference's `latents`, batch 1) */
  const float* guidance;    /* [F] per-frame guidance scale */
  const float* sigmas;      /* [steps+1] device Karras sigma table (sigmas[steps] = 0) */
  const int32_t* step_index;/* device scalar: current step i (read); the kernel does not modify it */
  int32_t F, C, H, W;
  /* optional fused producer of the next step's model input:
   * next_in[b, f, y, x, 0:C] = latents_new / sqrt(sigma_next^2 + 1), next_in[..., C:2C] = image_latents[b,f]
   * written token-major bf16 in the zero-haloed conv layout (see PtGemmArgs map_mode 1), ld = next_ld */
  void* next_in;            /* may be NULL */
  const float* image_latents; /* [2, F, C, H, W] fp32 (NCHW) */
  int32_t next_ld;
  int32_t next_padded;      /* 1: rows laid out with one zero column/row per image ((H+1)*(W+1) rows/image) */
  int32_t mode;             /* 0: full step (update latents, next_in uses sigma[i+1]); 1: only build next_in for sigma[i] */
  int32_t single_pred;      /* 1: noise_pred holds ONE already-combined prediction [F,...] (plain scheduler.step) */
  int32_t row_begin;        /* next_in receives the model-input rows [row_begin, row_begin + row_count) of the CFG pair, */
  int32_t row_count;        /* packed from row 0 of next_in (CFG-branch sharding); row_count 0 means both rows */
} PtCfgEulerArgs;
int pt_cfg_euler_step(const PtCfgEulerArgs* a, void* stream);
/* *step_index += 1 (stream-ordered, so a captured CUDA graph of one step can be replayed) */
int pt_step_advance(int32_t* step_index, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* tcgen05 GEMM / implicit-GEMM convolution                                                   */
/*   D[m, n] = epilogue( sum_t sum_k A[m + shift_t, k] * Wt[n, t*K + k] )                     */
/* covers every dense contraction of P2/P3/P5: Linear, 1x1 conv, 3x3 conv (9 row-shifted taps */
/* over the zero-haloed NHWC layout), temporal (3,1,1) conv (3 taps shifted by H*W rows).     */
/* ------------------------------------------------------------------------------------------ */
enum { PT_DT_BF16 = 0, PT_DT_F32 = 1 };

typedef struct PtGemmArgs {
  const PtTensorMap* tmap_a0; /* rank-3 {K0, rows_per_batch, batches}, box {64,128,1} */
  const PtTensorMap* tmap_a1; /* optional second K-range (channel concat); NULL if unused */
  const PtTensorMap* tmap_b;  /* rank-2 {num_taps*(K0+K1), N_rows}, box {64, block_n/2} */
  int32_t rows_per_batch;     /* A/accumulator row space per batch */
  int32_t batches;
  int32_t n_out;              /* number of output columns (GEGLU: the gated width, = half of W rows) */
  int32_t k0_chunks;          /* K0 / 64 */
  int32_t k1_chunks;          /* K1 / 64 (0 if no second source) */
  int32_t num_taps;           /* 1, 3 or 9 */
  int32_t tap_shift[9];       /* row shift of each tap in the A row space */
  int32_t block_n;            /* accumulator tile width: multiple of 32, 32..256 (GEGLU: 64/128/192/256) */
  int32_t geglu;              /* 1: tile = [block_n/2 value rows | block_n/2 gate rows], out = v*gelu(g) */
  int32_t gate_row_offset;    /* row offset of the gate half inside Wt / bias (GEGLU only) */
  /* epilogue: val = acc_scale*(acc + bias[n] + rowvec[g(m), n]) + res1_scale*res1[m,n] + res2_scale*res2[m,n] */
  const float* bias;          /* [>= n rows of Wt] fp32 or NULL */
  const float* rowvec;        /* fp32 [groups, rowvec_ld] or NULL */
  int32_t rowvec_ld;
  int32_t rowvec_mode;        /* 0 none; 1: g = orow / rv_a; 2: g = ((orow / rv_a) * rv_b + orow % rv_mod + rv_off) % rv_c
                               * (rv_mod = 0 means rv_b, rv_off = 0: the unsharded case; a pixel-sharded temporal block
                               * passes its local pixel count in rv_mod and its first global pixel in rv_off) */
  int32_t rv_a, rv_b, rv_c;
  float acc_scale;
  const void* res1;           /* bf16 [out rows, res_ld] or NULL */
  const void* res2;
  float res1_scale, res2_scale;
  int32_t res_ld;
  /* output */
  void* out;
  int32_t out_ld;
  int32_t out_dtype;          /* PT_DT_BF16 / PT_DT_F32 */
  /* optional second output: out2[m,n] = val + aux_scale * aux[m,n]  (ControlNet residual injected into the
   * UNet skip tensor, models/unet_spatio_temporal_condition_controlnet.py:451-459); aux NULL: out2 = val */
  void* out2;
  const void* aux;
  float aux_scale;
  /* row mapping accumulator row -> output row:
   *   0: identity (orow = batch*rows_per_batch + r)
   *   1: zero-haloed image space -> compact: r = img*(pH1*pW1) + y*pW1 + x, valid iff y < pH1-1, x < pW1-1,
   *      y % ostride == 0, x % ostride == 0; orow = (img*oH + y/ostride)*oW + x/ostride (see out_halo) */
  int32_t map_mode;
  int32_t pW1, pH1, ostride, oW, oH;
  int32_t out_halo;           /* map_mode 1 only: write the zero-haloed layout of the (strided) output image instead
                               * of the compact one: orow = (img*(oH+1) + y/ostride)*(oW+1) + x/ostride; halo rows
                               * are never written (the caller zeroes the buffer once) */
  int32_t act_silu;           /* activation after acc_scale and before the residual terms: 0 none, 1 SiLU (cond-embedding
                               * convs, models/controlnet_sdv.py:103-109), 2 GELU (exact erf; CLIP ViT-H MLP),
                               * 3 quick-GELU x*sigmoid(1.702x) (CLIP ViT-L style configs) */
  int32_t cta_pair;           /* 1: run as clusters of two CTAs sharing 256 x block_n tiles (tcgen05 cta_group::2);
                               * tmap_b's box is block_n/2 rows in both modes */
  int32_t rv_mod, rv_off;     /* see rowvec_mode 2 */
  /* Fused all-to-all of the frame-sharded plan (SURVEY.md 8e "Frames"): instead of writing `out`, the epilogue
   * scatters every finished output row straight into the buffer of the rank that owns it in the OTHER sharding
   * (frames <-> pixels), over NVLink peer memory — the GEMM and the exchange are one kernel, no pack / NCCL / unpack.
   * The local output row is r = (b*sc_J + j)*sc_S + s; rank q owns [sc_start[q], sc_start[q] + sc_count[q]) of the
   * split axis.
   *   1: split the INNER axis s (frame layout -> pixel layout):
   *        row on q = (b*sc_kept_total + sc_kept_off + j)*sc_count[q] + (s - sc_start[q])
   *   2: split the MIDDLE axis j (pixel layout -> frame layout):
   *        row on q = (b*sc_count[q] + (j - sc_start[q]))*sc_kept_total + sc_kept_off + s
   * sc_peer[q] is the (peer-mapped) base of rank q's destination, row stride out_ld; bf16 outputs with n_out % 8 == 0
   * only; `out` / `out2` are ignored.  The caller orders consumers behind a cross-rank barrier. */
  int32_t scatter_mode;
  int32_t sc_world;           /* <= 8 */
  int32_t sc_J, sc_S;
  int32_t sc_kept_off, sc_kept_total;
  int32_t sc_start[8], sc_count[8];
  void* sc_peer[8];
  /* optional DEVICE scalar multiplied into acc_scale when the kernel runs (NULL: none).  The ControlNet zero-convs
   * read `conditioning_scale` (models/controlnet_sdv.py:641-643) through it, so that a captured CUDA graph of the
   * denoise step picks up the value of the current call at replay instead of the one baked in at capture. */
  const float* acc_scale_ptr;
} PtGemmArgs;
int pt_gemm(const PtGemmArgs* a, void* stream);
/* debug only: int64[8*64] device buffer receiving clock64 stamps of CTA 0 of later pt_gemm launches (NULL: off) */
void pt_gemm_set_trace(void* buf);

/* ------------------------------------------------------------------------------------------ */
/* Fused GEGLU feed-forward (diffusers FeedForward = GEGLU proj + Linear; BasicTransformerBlock.ff,           */
/* TemporalBasicTransformerBlock.ff_in / .ff as run by models/modified_svd.py:70-74,100-107):                */
/*   out = acc_scale * (GEGLU(x W1^T + b1) W2^T + b2) + res1_scale * res1 + res2_scale * res2                 */
/* in ONE kernel: the [rows, 4C] hidden activations stay in TMEM / shared memory (CTA pairs, 256-row tiles,   */
/* hidden dimension walked in chunks of 64).  For C <= 320 (the output accumulator must fit TMEM next to the */
/* hidden one): level 0 of the SVD UNet.                                                                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct PtMlpArgs {
  const PtTensorMap* tmap_x;   /* x  bf16 [rows, C]:  rank-2 {C, rows},        box {64, 128} */
  const PtTensorMap* tmap_w1;  /* W1 bf16 [2*hidden, C] (value rows, then gate rows): rank-2 {C, 2*hidden}, box {64, 64} */
  const PtTensorMap* tmap_w2;  /* W2 bf16 [C, hidden]: rank-2 {hidden, C},     box {64, C/4} */
  int32_t rows, C, hidden;     /* hidden == 4*C; C a multiple of 64, <= 320 */
  const float* bias1;          /* [2*hidden] */
  const float* bias2;          /* [C] */
  float acc_scale;
  const void* res1;            /* bf16 [rows, res_ld] or NULL */
  const void* res2;
  float res1_scale, res2_scale;
  int32_t res_ld;
  void* out;                   /* bf16 [rows, out_ld] */
  int32_t out_ld;
  void* trace;                 /* debug only: int64[2*8*64] clock stamps of CTA 0, or NULL */
} PtMlpArgs;
int pt_mlp_geglu(const PtMlpArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* GroupNorm(32) (+SiLU): diffusers ResnetBlock2D.norm1/2, TemporalResnetBlock.norm1/2 (5-D    */
/* statistics), TransformerSpatioTemporalModel.norm, conv_norm_out                             */
/* (models/unet_spatio_temporal_condition_controlnet.py:237-238,494-495)    */
/* ------------------------------------------------------------------------------------------ */
typedef struct PtGroupNormArgs {
  const void* x0;           /* bf16 [rows, ld0], channels [0, c0) */
  const void* x1;           /* optional bf16 [rows, ld1], channels [c0, c0+c1): the un-materialised torch.cat */
  int32_t c0, c1, ld0, ld1;
  int32_t rows_per_stat;    /* rows sharing statistics: H*W (per frame) or F*H*W (TemporalResnetBlock) */
  int32_t num_stat;         /* number of statistics groups: B*F or B */
  void* stats;              /* workspace of pt_groupnorm_workspace_bytes() bytes, ZERO-INITIALISED ONCE by the caller (it
                             * holds rendezvous counters the kernel re-arms): per-CTA fp64 partial sums folded in a
                             * fixed order (no floating-point atomics: two runs of the same input are bit-identical).
                             * One launch: statistics and normalisation share a co-resident grid. */
  const float* gamma;       /* [c0+c1] */
  const float* beta;
  float eps;
  int32_t silu;
  void* out;                /* bf16 [out rows, out_ld] */
  int32_t out_ld;
  int32_t halo;             /* 1: write the zero-haloed image layout ((H+1)*(W+1) rows per image, pads zeroed) */
  int32_t H, W;
  /* Sharded statistics (frame/pixel-sharded TemporalResnetBlock norms: the statistics group spans ranks):
   *   mode 0: statistics + normalisation in one launch (default)
   *   mode 1: statistics only — writes fp64 {sum, sum of squares} per (group, norm group) to `sums` [num_stat,32,2]
   *   mode 2: normalisation only — reads (all-reduced) `sums`, `count` = elements per (statistics, norm) group */
  int32_t mode;
  double* sums;
  double count;
  /* mode 2 without an all-reduce: when n_peers > 0 the statistics are the sum, in rank order (identical bits on every
   * rank), of the n_peers (<= 8) per-rank `sums` arrays sums_peers[q] — peer-mapped memory read over NVLink; the caller
   * puts a cross-rank barrier between the mode-1 launches and this one */
  int32_t n_peers;
  const double* sums_peers[8];
} PtGroupNormArgs;
int pt_groupnorm(const PtGroupNormArgs* a, void* stream);
/* bytes of PtGroupNormArgs.stats needed for this problem (-1 on bad arguments) */
int64_t pt_groupnorm_workspace_bytes(int32_t num_stat, int32_t rows_per_stat, int32_t channels);

/* LayerNorm (BasicTransformerBlock / TemporalBasicTransformerBlock norms, eps 1e-5).  Optional fused
 * `hidden_states + emb` of models/modified_svd.py:196-197: addvec[frame] is added first and the sum is also
 * written to sum_out. */
typedef struct PtLayerNormArgs {
  const void* x;            /* bf16 [rows, ld] */
  int32_t ld;
  const float* gamma;
  const float* beta;
  float eps;
  void* out;                /* bf16 [rows, out_ld] */
  int32_t out_ld;
  int32_t rows, C;
  const float* addvec;      /* optional fp32 [F, C]; frame = (row / hw) % F */
  int32_t hw, F;
  void* sum_out;            /* optional bf16 [rows, out_ld] = x + addvec */
} PtLayerNormArgs;
int pt_layernorm(const PtLayerNormArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* attention                                                                                  */
/* ------------------------------------------------------------------------------------------ */
/* Spatial self-attention (BasicTransformerBlock.attn1): per image, S = H*W tokens, head_dim 64, non-causal,
 * scale 64^-0.5.  qkv is the fused projection [n_img*S, 3C] = (Q | K | V); tmap_qkv is a rank-3 tensor map
 * {3C, S, n_img} with box {64, 128, 1} over it. */
typedef struct PtAttnSpatialArgs {
  const PtTensorMap* tmap_qkv;
  void* out;                /* bf16 [n_img*S, out_ld] */
  int32_t out_ld;
  int32_t S, heads, C, n_img;
  float* lse;               /* optional fp32 [n_img, heads, S]: log2(sum_j exp2(scale*log2e*s_ij)) per query row, kept for
                             * the training backward (pt_attention_spatial_bwd); NULL at inference */
} PtAttnSpatialArgs;
int pt_attention_spatial(const PtAttnSpatialArgs* a, void* stream);

/* Temporal self-attention (TemporalBasicTransformerBlock.attn1, models/modified_svd.py:64-81): for each
 * (batch, pixel, head) attention over the F frames; rows are (b*F + f)*HW + s. */
typedef struct PtAttnTemporalArgs {
  const void* qkv;          /* bf16 [B*F*HW, ld] = (Q | K | V) */
  int32_t ld;
  void* out;                /* bf16 [B*F*HW, out_ld] */
  int32_t out_ld;
  int32_t B, F, HW, heads, C;
} PtAttnTemporalArgs;
int pt_attention_temporal(const PtAttnTemporalArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* small / layout kernels                                                                     */
/* ------------------------------------------------------------------------------------------ */
/* out[m, n] (+)= act_out( sum_k act_in(in[m, k]) * W[n, k] + bias[n] ), fp32 in/out, bf16 weights (tiny M):
 * TimestepEmbedding MLPs, time_emb_proj(silu(emb)), 1-token cross-attention to_v/to_out, cc_projection camera columns */
typedef struct PtSmallLinearArgs {
  const float* in;
  int32_t in_ld;
  const void* w;            /* bf16 [N, w_ld] */
  int32_t w_ld;
  const float* bias;
  float* out;
  int32_t out_ld;
  int32_t M, N, K;
  int32_t act_in_silu, act_out_silu, accumulate;
} PtSmallLinearArgs;
int pt_small_linear(const PtSmallLinearArgs* a, void* stream);

/* Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): out[m] = [cos(t_m f_k) | sin(t_m f_k)].
 * If t is NULL every row uses t = 0.25*ln(sigmas[*step_index]) (utils/scheduling_euler_discrete_karras_fix.py:344-347) */
typedef struct PtSinCosArgs {
  const float* t;
  const float* sigmas;
  const int32_t* step_index;
  float* out;
  int32_t out_ld;
  int32_t M, dim;
} PtSinCosArgs;
int pt_timestep_sincos(const PtSinCosArgs* a, void* stream);

/* nearest-neighbour 2x (diffusers Upsample2D) from compact [n,H,W,C] to [n,2H,2W,C], optionally zero-haloed;
 * scale == 1 is a plain copy into the zero-haloed layout (input of the stride-2 Downsample2D conv) */
typedef struct PtUpsampleArgs {
  const void* x;
  int32_t ld;
  void* out;
  int32_t out_ld;
  int32_t n, H, W, C, halo;
  int32_t scale;            /* 1 or 2 (0 means 2) */
} PtUpsampleArgs;
int pt_upsample2x(const PtUpsampleArgs* a, void* stream);

/* direct 3x3 conv (+SiLU), pad 1, stride 1|2, Cin <= 32, Cout in {16, 32}: the narrow head of
 * ControlNetConditioningEmbeddingSVD (models/controlnet_sdv.py:84-109) */
typedef struct PtConvDirectArgs {
  const void* x;            /* fp32 NCHW [n,Cin,H,W] if in_nchw_f32 else bf16 NHWC compact [n*H*W, in_ld] */
  int32_t in_nchw_f32, in_ld;
  const float* w;           /* fp32 [3][3][Cin][Cout] */
  const float* bias;
  void* out;                /* bf16 NHWC [rows, out_ld], compact or zero-haloed (halo never written) */
  int32_t out_ld, out_halo;
  int32_t n, H, W, Cin, Cout, stride, silu;
} PtConvDirectArgs;
int pt_conv3x3_direct(const PtConvDirectArgs* a, void* stream);

/* reference tensors are NCHW ([B*F, C, H, W], fp32 or bf16); the kernels work on token-major bf16 */
typedef struct PtLayoutArgs {
  const void* nchw_const_unused; /* reserved (keeps the struct layout stable) */
  void* nchw;               /* NCHW side (source of pt_nchw_to_tokens, destination of pt_tokens_to_nchw) */
  void* tokens;             /* bf16 [rows, ld] side */
  int32_t n, C, H, W, ld, halo, nchw_f32;
} PtLayoutArgs;
int pt_nchw_to_tokens(const PtLayoutArgs* a, void* stream);
int pt_tokens_to_nchw(const PtLayoutArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Row-block copy between differently ordered token layouts (frame-sharded <-> pixel-sharded activations around a */
/* temporal sub-block): for i < n_blocks copies `rows[i]` rows of `cols` bf16 from src row src_row[i] to dst row   */
/* dst_row[i].  The three int32 tables live on the device.                                                       */
/* ------------------------------------------------------------------------------------------ */
typedef struct PtRowBlockCopyArgs {
  const void* src;
  void* dst;
  int32_t src_ld, dst_ld, cols;
  const int32_t* src_row;
  const int32_t* dst_row;
  const int32_t* rows;
  int32_t n_blocks;
} PtRowBlockCopyArgs;
int pt_row_block_copy(const PtRowBlockCopyArgs* a, void* stream);

/* out = x + scale * y on bf16 [rows, cols] (skip_i = h_i + m_i * r_i when the injection cannot ride a GEMM epilogue) */
typedef struct PtAxpyArgs {
  const void* x;
  const void* y;
  void* out;
  int32_t ld_x, ld_y, ld_out, rows, cols;
  float scale;
} PtAxpyArgs;
int pt_axpy_bf16(const PtAxpyArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* VAE (pipeline/pipeline_stable_video_diffusion_controlnet.py:174-195     */
/* `_encode_vae_image`, :225-251 `decode_latents`): diffusers AutoencoderKLTemporalDecoder.     */
/* Its convolutions, norms and linears run on pt_gemm / pt_groupnorm; these two are the rest.  */
/* ------------------------------------------------------------------------------------------ */
/* out[r, :cols] = softmax(in[r, :cols]) — fp32 logits of the single-head (head_dim = C) mid-block attention,
 * bf16 probabilities (the A operand of the P V GEMM); columns >= cols are left untouched */
typedef struct PtSoftmaxArgs {
  const float* in;
  void* out;                /* bf16 */
  int32_t rows, cols, ld_in, ld_out;
} PtSoftmaxArgs;
int pt_softmax_rows(const PtSoftmaxArgs* a, void* stream);

/* TemporalDecoder.time_conv_out: Conv3d(C, C, (3,1,1), padding (1,0,0)) over the frames of each video, on the
 * token-major fp32 output of conv_out, written as the caller's NCHW fp32 frames [B*F, C, H, W] */
typedef struct PtTimeConvArgs {
  const float* in;          /* [B*F*HW, ld] */
  int32_t ld;
  const float* w;           /* [C, C, 3] */
  const float* bias;        /* [C] */
  float* out;               /* [B*F, C, H*W] */
  int32_t B, F, HW, C;
} PtTimeConvArgs;
int pt_time_conv3(const PtTimeConvArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Image-conditioning branch (pipeline/pipeline_stable_video_diffusion_     */
/* controlnet.py:145-172 `_encode_image`, :602-712 `_resize_with_antialiasing`).               */
/* ------------------------------------------------------------------------------------------ */
/* One pass of the separable Gaussian pre-filter with reflect padding (`_gaussian_blur2d` / `_filter2d`, :652-712):
 * out[p, y, x] = sum_i w[i] * in[p, y, x + i - (k-1)/2]  (axis 0)   or the same along y (axis 1) */
typedef struct PtBlurArgs {
  const float* in;          /* [planes, H, W] fp32 */
  float* out;               /* [planes, H, W] fp32 */
  const float* w;           /* device [k] */
  int32_t planes, H, W, k, axis;
} PtBlurArgs;
int pt_blur_reflect(const PtBlurArgs* a, void* stream);

/* torch.nn.functional.interpolate(mode="bicubic", align_corners=True) (:631) of one image's planes to S x S, written as
 * fp32 planes [C, S, S] (out_f32, optional) and/or as the bf16 im2col rows of CLIP's patch embedding (out_patches,
 * optional): row = patch (py * S/P + px), column = c*P*P + ky*P + kx, columns C*P*P .. ld-1 untouched (zero them once) */
typedef struct PtBicubicArgs {
  const float* in;          /* [C, H, W] fp32 */
  int32_t C, H, W, S, P;
  float* out_f32;
  void* out_patches;        /* bf16 [ (S/P)^2, ld ] */
  int32_t ld;
} PtBicubicArgs;
int pt_bicubic_resize(const PtBicubicArgs* a, void* stream);

/* Multi-head self-attention for short sequences and any head_dim <= 128 (CLIP vision tower: 257 tokens, 16 heads of
 * 80): qkv rows are [q | k | v] of width 3C, head h owns columns h*hd .. (h+1)*hd of each third; fp32 softmax */
typedef struct PtAttnSmallArgs {
  const void* qkv;          /* bf16 [S, 3C] */
  int32_t ld;
  void* out;                /* bf16 [S, C] */
  int32_t out_ld;
  int32_t S, heads, head_dim;
} PtAttnSmallArgs;
int pt_attention_small(const PtAttnSmallArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* R1: trajectory maps (scripts/run_inference_vipseg_json_repro.py:438-449, utils/dataset.py:  */
/* 741-766): per frame transition k a black canvas with, per track in order, cv2.line(p_k,     */
/* p_k+1, BGR (0,0,255), thickness 3) and cv2.circle(p_k+1, 3, BGR (0,255,0), filled); then    */
/* BGR->RGB, a black last frame and the pipeline's preprocess (x/255*2-1, :500).  Bit-exact    */
/* against OpenCV's integer rasterisation.                                                     */
/* ------------------------------------------------------------------------------------------ */
typedef struct PtRasterArgs {
  const int32_t* tracks;    /* device [K, F, 2] (x, y) pixel coordinates, already rescaled and int()-truncated */
  int32_t K, F, H, W;
  void* order;              /* workspace of pt_rasterize_workspace_bytes(F, H, W) bytes */
  float* out_f32;           /* optional [F, 3, H, W] fp32 RGB in [-1, 1] (the pipeline's controlnet_condition) */
  uint8_t* out_u8;          /* optional [F, H, W, 3] uint8 RGB (the PIL images the reference builds) */
  int32_t swap_per_track;   /* 1: reproduce utils/dataset.py:741-766, where cv2.cvtColor(BGR2RGB) sits INSIDE the track loop:
                             * the canvas channels swap after every track, so a line drawn by track k ends up red iff
                             * (K - k) is odd, blue otherwise (discs are green either way) */
} PtRasterArgs;
int pt_rasterize_tracks(const PtRasterArgs* a, void* stream);
int64_t pt_rasterize_workspace_bytes(int32_t F, int32_t H, int32_t W);

/* ------------------------------------------------------------------------------------------ */
/* Training step of BASELINE configs[3]: backward / loss / optimizer kernels.             */
/* Reference: scripts/train_svd_traj_VIPSeg_14_cam_concat.py:1404-1475 (autograd through ControlNet + frozen    */
/* UNet, EDM-weighted MSE :1423-1436, AdamW :1472).  Formulas: oracle/backward.py (checked against autograd).   */
/* dgrad of a linear / conv layer is pt_gemm itself with W^T and negated tap shifts.                            */
/* ------------------------------------------------------------------------------------------ */
typedef struct PtEdmLossArgs {
  const void* pred;          /* bf16 tokens [B*F*HW, pred_ld], C columns: the UNet's noise prediction */
  int32_t pred_ld;
  const float* noisy;        /* fp32, element (b, f, c, pix) at b*sample_stride + f*frame_stride + c*HW + pix */
  const float* target;
  int64_t sample_stride, frame_stride;
  const float* sigmas;       /* [B] */
  int32_t B, F, C, HW;
  float weight;              /* 1 for the main loss, 0.5 for the spatial pass (:1462) */
  void* dpred;               /* bf16 tokens [B*F*HW, dpred_ld] = weight * d loss / d pred, or NULL */
  int32_t dpred_ld;
  void* workspace;           /* pt_edm_loss_workspace_bytes() */
  float* loss;               /* [1] */
  int32_t accumulate;        /* loss += instead of = */
} PtEdmLossArgs;
int pt_edm_loss(const PtEdmLossArgs* a, void* stream);
int64_t pt_edm_loss_workspace_bytes(void);

typedef struct PtGroupNormBwdArgs {
  const void* x0;            /* forward input, as PtGroupNormArgs */
  const void* x1;
  int32_t c0, c1, ld0, ld1;
  const void* dout;          /* bf16 gradient of the forward output, in the forward output's layout (halo: (H+1)*(W+1) rows/image) */
  int32_t dout_ld, halo, H, W;
  const float* gamma;
  const float* beta;
  float eps;
  int32_t silu, rows_per_stat, num_stat;
  void* dx0;                 /* bf16 [rows, dld0] */
  void* dx1;                 /* bf16 [rows, dld1] or NULL */
  int32_t dld0, dld1;
  void* workspace;           /* pt_groupnorm_bwd_workspace_bytes(num_stat, c0 + c1) */
  float* dgb_out;            /* fp32 [2*(c0+c1)] = dgamma | dbeta, or NULL (frozen layer) */
  int32_t accumulate_dgb;
} PtGroupNormBwdArgs;
int pt_groupnorm_bwd(const PtGroupNormBwdArgs* a, void* stream);
int64_t pt_groupnorm_bwd_workspace_bytes(int32_t num_stat, int32_t channels);

typedef struct PtLayerNormBwdArgs {
  const void* x;             /* bf16 [rows, ld]: forward input */
  int32_t ld;
  const void* dout;
  int32_t dout_ld;
  const float* gamma;
  float eps;
  int32_t rows, C;
  const float* addvec;       /* as PtLayerNormArgs (the forward normalised x + addvec[frame]) or NULL */
  int32_t hw, F;
  void* dx;                  /* bf16 [rows, dx_ld] */
  int32_t dx_ld;
  int32_t accumulate_dx;
  float* partials;           /* fp32 [n_blocks][2*C] scratch, or NULL when dgamma / dbeta are not wanted */
  int32_t n_blocks;          /* grid size (>= 1) */
  float* dgb_out;            /* fp32 [2*C] = dgamma | dbeta, or NULL */
  int32_t accumulate_dgb;
} PtLayerNormBwdArgs;
int pt_layernorm_bwd(const PtLayerNormBwdArgs* a, void* stream);

/* GEGLU as a separate pass: h bf16 [rows, 2*hidden] (value | gate) -> out bf16 [rows, hidden]; backward -> dh */
int pt_geglu_fwd(const void* h, int32_t ld, void* out, int32_t out_ld, int64_t rows, int32_t hidden, void* stream);
int pt_geglu_bwd(const void* h, int32_t ld, const void* dout, int32_t dout_ld, void* dh, int32_t dh_ld, int64_t rows,
                 int32_t hidden, void* stream);

typedef struct PtColsumArgs {
  const void* x;             /* bf16 [rows, ld] (halo: zero-haloed image rows, only real pixels are summed) */
  int32_t ld, halo, H, W;
  int64_t rows_per_group;    /* logical (un-haloed) rows per group */
  int32_t groups, C;
  float scale;
  float* out;                /* fp32 [groups, C] */
  int32_t accumulate;
} PtColsumArgs;
int pt_colsum(const PtColsumArgs* a, void* stream);
/* out[i] (+)= scale * sum_b partials[b*n + i], b = 0..nb-1 in order */
int pt_reduce_partials(const float* partials, int32_t nb, int64_t n, float scale, float* out, int32_t accumulate, void* stream);
/* out[0] (+)= scale * sum a[r,c]*b[r,c]; workspace: 8 KiB */
int pt_dot_bf16(const void* a, int32_t lda, const void* b, int32_t ldb, int64_t rows, int32_t cols, float scale, float* out,
                int32_t accumulate, void* workspace, void* stream);
/* bf16 [rows, cols] (row stride ld_in) -> [cols, rows] (row stride ld_out) */
int pt_transpose_bf16(const void* in, int32_t ld_in, void* out, int32_t ld_out, int32_t rows, int32_t cols, void* stream);

typedef struct PtAdamWArgs {
  float* master;             /* fp32 parameters */
  const float* grad;         /* fp32 gradients (after the data-parallel all-reduce) */
  float* m;
  float* v;
  void* work;                /* optional bf16 copy the kernels read */
  int64_t n;
  float lr, beta1, beta2, eps, weight_decay, grad_scale;
  int32_t step;              /* 1-based */
} PtAdamWArgs;
int pt_adamw(const PtAdamWArgs* a, void* stream);

/* Weight gradient of a linear / implicit-GEMM conv layer (oracle/backward.py conv_rows_wgrad):                  */
/*   dW[n, t*K + k] = sum_r dD[r, n] * A[r + shift_t, k]                                                         */
/* tcgen05, both operands MN-major, i.e. read in the layout the tensors already have: tap = row shift of the TMA           */
/* coordinate of the input (out-of-range rows zero-filled like the forward).  Rows are split over `splits` CTAs per      */
/* output tile (128 n x up to 256 k); fp32 partial tiles are folded in order by pt_reduce_partials.                       */
typedef struct PtWgradArgs {
  const PtTensorMap* tmap_dt;   /* dD bf16 [rows, N] (the gradient itself, no transpose): rank-2 {N, rows}, box {64, 64} */
  const PtTensorMap* tmap_a;    /* layer input bf16 [rows, K]: rank-2 {K, rows}, box {64, 64} */
  int32_t rows, N, K, num_taps;
  int32_t tap_shift[9];
  int32_t splits;
  float* partials;              /* fp32 [splits][N][num_taps*K] */
} PtWgradArgs;
int pt_wgrad(const PtWgradArgs* a, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Whole-network training step: the remaining backward kernels (train_attn.cu, train_misc.cu). */
/* Reference: autograd through diffusers' Attention / Upsample2D / Downsample2D / TimestepEmbedding and the     */
/* conditioning embedding (models/controlnet_sdv.py:95-116) inside `accelerator.backward(loss)`,               */
/* scripts/train_svd_traj_VIPSeg_14_cam_concat.py:1470.                                                        */
/* ------------------------------------------------------------------------------------------ */
/* delta[img, head, s] = sum_d dO[row, head*64 + d] * O[row, head*64 + d]  (rows = n_img * S) */
int pt_attention_delta(const void* out, int32_t out_ld, const void* dout, int32_t dout_ld, float* delta, int64_t rows,
                       int32_t S, int32_t heads, void* stream);

typedef struct PtAttnSpatialBwdArgs {
  const void* qkv;          /* bf16 [n_img*S, ld] = (Q | K | V), the forward's input */
  int32_t ld;
  const void* dout;         /* bf16 [n_img*S, dout_ld]: gradient of the attention output */
  int32_t dout_ld;
  const float* lse;         /* [n_img, heads, S] from pt_attention_spatial (PtAttnSpatialArgs.lse) */
  const float* delta;       /* [n_img, heads, S] from pt_attention_delta */
  void* dqkv;               /* bf16 [n_img*S, dld] = (dQ | dK | dV) */
  int32_t dld;
  int32_t S, heads, C, n_img;
  /* optional (both or neither): rank-3 tensor maps {3C | C, S, n_img}, box {64, 128, 1}, over qkv and dout.  With them and
   * S >= 256 the dK / dV kernel runs on tcgen05 (S^T, dP^T, dV, dK in TMEM); without, on mma.sync. */
  const PtTensorMap* tmap_qkv;
  const PtTensorMap* tmap_dout;
} PtAttnSpatialBwdArgs;
int pt_attention_spatial_bwd(const PtAttnSpatialBwdArgs* a, void* stream);

typedef struct PtAttnTemporalBwdArgs {
  const void* qkv;          /* bf16 [B*F*HW, ld] = (Q | K | V), rows (b*F + f)*HW + s */
  int32_t ld;
  const void* dout;         /* bf16 [B*F*HW, dout_ld] */
  int32_t dout_ld;
  void* dqkv;               /* bf16 [B*F*HW, dld] */
  int32_t dld;
  int32_t B, F, HW, heads, C;
} PtAttnTemporalBwdArgs;
int pt_attention_temporal_bwd(const PtAttnTemporalBwdArgs* a, void* stream);

/* backward of pt_upsample2x (same struct: `x`/`ld` = the compact gradient to WRITE, `out`/`out_ld` = the gradient of the
 * forward output to READ): sum over the scale x scale children; scale 1 + halo strips the zero halo */
int pt_upsample2x_bwd(const PtUpsampleArgs* a, void* stream);
/* gradient of a stride-2 conv output ([n, ceil(H/2), ceil(W/2)] rows, zero-haloed when src_halo) -> zero-haloed
 * [n, H+1, W+1] rows with the values at the even pixels and zeros elsewhere (every row is written) */
int pt_dilate2x(const void* src, int32_t src_ld, int32_t src_halo, void* dst, int32_t dst_ld, int32_t n, int32_t H, int32_t W,
                int32_t C, void* stream);
/* zero the halo rows (y == H or x == W) of a zero-haloed [n, H+1, W+1] buffer */
int pt_zero_halo(void* x, int32_t ld, int32_t n, int32_t H, int32_t W, int32_t C, void* stream);
/* SiLU as a separate pass (bf16 [rows, cols]); backward: dx = dy * silu'(x) */
int pt_silu_fwd(const void* x, int32_t ld, void* out, int32_t out_ld, int64_t rows, int32_t cols, void* stream);
int pt_silu_bwd(const void* x, int32_t ld, const void* dy, int32_t dy_ld, void* dx, int32_t dx_ld, int64_t rows, int32_t cols,
                void* stream);

/* backward of pt_small_linear without output activation: y = W act(x) + b */
typedef struct PtSmallLinearBwdArgs {
  const float* x;           /* fp32 [M, x_ld]: the forward input (before act_in) */
  int32_t x_ld;
  const void* w;            /* bf16 [N, w_ld] */
  int32_t w_ld;
  const float* dy;          /* fp32 [M, dy_ld] */
  int32_t dy_ld;
  int32_t M, N, K, act_in_silu;
  float* dx;                /* fp32 [M, dx_ld] or NULL */
  int32_t dx_ld, accumulate_dx;
  float* dw;                /* fp32 [N, K] contiguous or NULL */
  float* db;                /* fp32 [N] or NULL (written with dw) */
  int32_t accumulate_w;
  void* dx_workspace;       /* pt_small_linear_bwd_workspace_bytes(M, N, K) when dx is wanted */
} PtSmallLinearBwdArgs;
int pt_small_linear_bwd(const PtSmallLinearBwdArgs* a, void* stream);
int64_t pt_small_linear_bwd_workspace_bytes(int32_t M, int32_t N, int32_t K);

/* out[g, c] (+)= scale * sum over rows r with group(r) == g of x[r, c]; group(r) =
 *   mode 1: r / ga            mode 2: ((r / ga) * gb + r % gb) % gc  (PtGemmArgs.rowvec_mode 1 / 2)
 *   mode 3: (r / ga) % gc     (frame of row (b*F + f)*HW + s with ga = HW, gc = F)
 * deterministic (two stages, fixed order); mode 2 is vectorised for gc <= 4, every other shape has a scalar fallback */
typedef struct PtColsumGroupedArgs {
  const void* x;            /* bf16 [rows, ld] */
  int32_t ld;
  int64_t rows;
  int32_t C, groups, mode, ga, gb, gc;
  float scale;
  float* out;               /* fp32 [groups, out_ld] */
  int32_t out_ld, accumulate;
  void* workspace;          /* pt_colsum_grouped_workspace_bytes(rows, groups, C) */
} PtColsumGroupedArgs;
int pt_colsum_grouped(const PtColsumGroupedArgs* a, void* stream);
int64_t pt_colsum_grouped_workspace_bytes(int64_t rows, int32_t groups, int32_t C);

#ifdef __cplusplus
}
#endif
#endif /* POSETRAJ_B200_H_ */
