/* faithdiff_b200.h — C-ABI of libfaithdiff_b200.so (hand-written sm_100a kernels for FaithDiff's tiled
 * latent-denoising hot path).  Plain pointers + sizes, no torch types.  All pointers are DEVICE pointers unless
 * a name ends in _host.  Every entry point returns 0 on success; non-zero → fd_last_error() describes it.
 * `stream` is a cudaStream_t passed as void*.  Nothing allocates device memory; work buffers are passed in.
 *
 * The reference (tpetkovich/FaithDiff-dev) is pure Python over diffusers/PyTorch library kernels — it has no
 * FFI.  Each entry point therefore cites the reference *Python call site(s)* whose GPU work it replaces
 * (paths relative to the reference root); INTEGRATION.md shows the ctypes binding and where the reference's
 * modules would call it.
 */
#ifndef FAITHDIFF_B200_H
#define FAITHDIFF_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* fd_last_error(void);
int fd_version(void);
/* number of kernel launches issued through this library since load (bench.py's gpu_launches counter) */
int64_t fd_launch_count(void);

/* ---------------------------------------------------------------------------------------------------------
 * Tensor-core implicit GEMM (tcgen05.mma, TMEM accumulators, TMA operand staging, persistent, warp-specialised)
 *   out[M, N] = epilogue( alpha * A[M, K] · B[N, K]^T )
 * mode 0  A = up to two row-major bf16 matrices concatenated along K           (Linear layers, 1x1 convs)
 * mode 1  A = implicit im2col of up to two NHWC bf16 images concatenated along C, 3x3 stride 1  (conv3x3)
 * mode 2  A = implicit im2col of one NHWC bf16 image, 3x3 stride 2             (Downsample2D)
 * mode 3  one output phase (a, b) = (pad_lo >> 1, pad_lo & 1) of conv3x3(nearest-2x-upsample(x)) — Upsample2D, diffusers
 *         `F.interpolate(scale_factor=2, mode="nearest")` + conv reached from unet_2d_condition_vae_extension.py:661-684 and
 *         the VAE decoder: out[2i+a, 2j+b] sees only a 2x2 neighbourhood of x, so the 4x tensor is never materialised and the
 *         contraction is 4*C deep instead of 9*C (B = the phase's pre-summed taps [N, 4*C]); out = the [batch, 2H, 2W, N] image
 * B = weights bf16 [N, taps*(C0+C1)] (tap-major, channel-minor; K contiguous).
 * Replaces: diffusers Conv2d/Linear reached from FaithDiff/models/unet_2d_condition_vae_extension.py:571-579,
 * :611-690 (conv_in…conv_out, ResnetBlock2D, Transformer2DModel projections, GEGLU feed-forward,
 * ControlNetConditioningEmbedding :213-219, ResidualAttentionBlock :238-260) and the Encoder / AutoencoderKL
 * conv stacks (:122-148; pipeline_FaithDiff_tlc.py:1272,1438).
 * --------------------------------------------------------------------------------------------------------- */
typedef struct fd_gemm_desc {
  const void* a[2];      /* A sources (bf16); a[1] NULL if single source                                  */
  int64_t a_cols[2];     /* K columns (mode 0) / channels (mode 1,2) taken from each source               */
  int64_t a_ld[2];       /* row pitch (mode 0) / pixel pitch (mode 1,2) in elements                       */
  const void* b;         /* weights bf16 [N, K_total]                                                     */
  int64_t M, N;          /* mode 0: rows; mode 1,2: M is derived (batch*out_h*out_w), N = out channels    */
  int32_t mode;
  int32_t batch, in_h, in_w; /* conv input geometry                                                       */
  int32_t pad_lo;        /* zero padding at top/left (1: pad=1; 0: the VAE's F.pad(0,1,0,1) + pad 0)      */
  int32_t act;           /* 0 none, 1 SiLU, 2 GEGLU (B rows interleaved h0,g0,h1,g1… → N/2 outputs), 3 QuickGELU, */
                         /* 4 GELU (erf form; OpenCLIP-bigG text encoder MLP), 5 LeakyReLU(0.2) (BSRNet)        */
  int32_t bias_along_m;  /* bias indexed by output row instead of column                                  */
  int32_t out_fp32;      /* out dtype: 0 bf16, 1 fp32                                                     */
  const float* bias;     /* [N] (or [M]) fp32, or NULL                                                    */
  const float* rowbias;  /* [M / rows_per_batch, N] fp32 added per sample (timestep embedding), or NULL   */
  int64_t rows_per_batch;
  int64_t ld_rowbias;    /* row pitch of rowbias in floats (0 → N)                                        */
  const void* residual;  /* bf16 [M, ld_res] added after the activation, or NULL                          */
  int64_t ld_res;
  void* out;             /* [M, ld_out]                                                                   */
  int64_t ld_out;
  float alpha;
  int32_t block_n;       /* 0 = auto; low 12 bits: N tile, multiple of 32 in [32, 256]; tuning hooks: bit 12    */
                         /* forces CTA-pair (cta_group::2) tiles, bit 13 forces single-CTA tiles               */
} fd_gemm_desc;
int fd_gemm(const fd_gemm_desc* d, void* stream);

/* LayerNorm folded into the GEMMs around it (opt-in: FD_LN_FOLD=1 in the Python layer; not part of the default path).
 * BasicTransformerBlock's norm1 / norm2 / norm3 each feed exactly one Linear (fused QKV, cross-attention Q, GEGLU FF1);
 * with W' = bf16(W * gamma), colsum[n] = sum_k W'[n,k] and bias' = bias + W beta,
 *     LayerNorm(x) W^T + bias = rstd_r * (x W'^T) - rstd_r * mean_r * colsum + bias'
 * so the consumer reads the raw residual stream and applies a per-row affine in its epilogue, and the row statistics come
 * from the epilogue of the GEMM that produced x (attention-out / FF2 / proj_in): stats_out [M, stats_out_ld] float2 slots
 * (sum, sum of squares of the bf16-rounded outputs) per (N tile, chunk half); fd_gemm_ln writes the number of slots it used
 * to stats_out_n, which the consumer passes as stats_in_n.  A call either produces (stats_out) or consumes (stats_in). */
typedef struct fd_gemm_ln_desc {
  const void* stats_in;  /* float2 [M, stats_in_ld], or NULL                                             */
  int64_t stats_in_ld;
  int32_t stats_in_n;    /* slots per row to sum                                                         */
  int32_t pad0_;
  const float* colsum;   /* [N] fp32                                                                     */
  float inv_c;           /* 1 / (LayerNorm width = K of this GEMM)                                       */
  float eps;
  void* stats_out;       /* float2 [M, stats_out_ld], or NULL                                            */
  int64_t stats_out_ld;  /* >= 2 * ceil(N / block_n); 2 * ceil(N / 32) always suffices                   */
  int32_t stats_out_n;   /* OUT: slots written per row                                                   */
  int32_t pad1_;
  /* GroupNorm statistics from the producer: per (32-row epilogue box, column pair) sum / sum of squares of the bf16-rounded
   * outputs, so that fd_groupnorm_nhwc_from_stats normalises in ONE streaming pass.  Box r of tile t = rows [128 t + 32 r, +32)
   * (mode 0) / the r-th 32-pixel quarter of the t-th pixel patch (conv modes; a patch never crosses an image).              */
  void* colstats_out;    /* float2 [>= colstats_boxes, colstats_ld], or NULL                             */
  int64_t colstats_ld;   /* >= N / 2                                                                     */
  int32_t colstats_boxes;            /* OUT: box rows written = 4 * number of 128-row tiles              */
  int32_t colstats_boxes_per_image;  /* OUT: conv modes: 4 * tiles per image; mode 0: 0                  */
} fd_gemm_ln_desc;
int fd_gemm_ln(const fd_gemm_desc* d, fd_gemm_ln_desc* ln, void* stream);
int32_t fd_sizeof_gemm_ln_desc(void);
int fd_sizeof_gemm_desc(void);   /* lets a foreign-language binding verify its struct mirror */

/* ---------------------------------------------------------------------------------------------------------
 * Flash-style attention, head_dim 64, bf16, non-causal, no mask (tcgen05 QK^T and PV, online softmax in fp32).
 *   q: rows [batch*lq], k/v: rows [batch*lk]; head h occupies columns [h*64, h*64+64) from q/k/v + pitch ld*.
 * Replaces: F.scaled_dot_product_attention inside diffusers AttnProcessor2_0 for attn1 (self) and attn2
 * (cross, lk = 77) of every BasicTransformerBlock reached from unet_2d_condition_vae_extension.py:617-684.
 * --------------------------------------------------------------------------------------------------------- */
int fd_attention_d64(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                     int64_t ldo, int32_t batch, int32_t heads, int32_t lq, int32_t lk, float scale, void* stream);
/* Same, causal (key j is visible to query i iff j <= i), lk <= 128.
 * Replaces: the causal self-attention of the CLIP text encoders (transformers CLIPAttention with the causal mask built in
 * CLIPTextTransformer.forward) that pipeline_FaithDiff_tlc.py:605 (`text_encoder(text_input_ids, output_hidden_states=True)`)
 * runs for encode_prompt — SURVEY §8f rank 1.                                                                          */
int fd_attention_d64_causal(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                            int64_t ldo, int32_t batch, int32_t heads, int32_t lq, int32_t lk, float scale, void* stream);
/* Single-head attention with head_dim 512 — the mid-block attention of the FaithDiff LQ Encoder and of the VAE
 * (unet_2d_condition_vae_extension.py:96-106; diffusers Attention(heads=1, dim_head=512) inside UNetMidBlock2D): flash-style,
 * S and P never leave the SM.  q/k/v/out: bf16 row views [batch*l, 512] with pitches ld* (elements, multiples of 8); softmax over
 * the l keys of the same sample, scale = 512^-0.5 at the call sites.  Replaces the q·k^T GEMM -> fd_softmax_rows -> p·v GEMM
 * sequence (which wrote the [l, l] fp32 score matrix to HBM) of round 1.                                                        */
int fd_attention_d512(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out, int64_t ldo,
                      int32_t batch, int32_t l, float scale, void* stream);


/* row softmax: p[r, :] = softmax(alpha * s[r, :]) ; s fp32 [rows, ld_s], p bf16 [rows, ld_p]
 * (single-head d=512 attention of the VAE / LQ-encoder mid block, unet_…extension.py:96-106)              */
int fd_softmax_rows(const float* s, int64_t ld_s, void* p, int64_t ld_p, int64_t rows, int64_t cols, float alpha,
                    void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Bandwidth-bound kernels (NHWC bf16 activations)
 * --------------------------------------------------------------------------------------------------------- */
/* GroupNorm(+SiLU) over the channel-concatenation of up to two NHWC sources → one NHWC output [B,HW,C0+C1].
 * work: fp32 scratch of fd_groupnorm_work_floats(batch, groups, c0+c1) elements.
 * Replaces: ResnetBlock2D.norm1/norm2 + SiLU, Transformer2DModel.norm, conv_norm_out (…extension.py:688-689),
 * ControlNetConditioningEmbedding.norm_in (:215), torch.cat of skip connections in the up blocks.          */
int64_t fd_groupnorm_work_floats(int32_t batch, int32_t groups, int32_t channels);
int fd_groupnorm_nhwc(const void* x0, int32_t c0, const void* x1, int32_t c1, const float* gamma, const float* beta,
                      void* out, int32_t batch, int64_t hw, int32_t groups, float eps, int32_t silu, float* work,
                      void* stream);
/* Same result (same statistics; SiLU through tanh.approx, relative error 2^-11 of the sigmoid) from ONE launch: every CTA
 * reduces one pixel chunk of one sample, arrives on the sample's counter in `sync`, waits for the sample's other chunks,
 * combines the partial sums itself and normalises its chunk, re-reading it from L2 (norm.cu) — the path the UNet / VAE use.
 * sync: fd_groupnorm_sync_bytes() bytes of device memory, zeroed ONCE by the caller and then left alone (the kernel
 * recycles its counters); one buffer per stream that may run GroupNorm concurrently.  The grid must be co-resident: do not
 * run two of these concurrently on one device.  Falls back to the three-launch fd_groupnorm_nhwc for shapes it does not
 * cover (more than 2560 channels or 64 samples), for sync == NULL, or with FD_GN_FUSED=0.                          */
int64_t fd_groupnorm_sync_bytes(void);
int fd_groupnorm_nhwc_fused(const void* x0, int32_t c0, const void* x1, int32_t c1, const float* gamma, const float* beta,
                            void* out, int32_t batch, int64_t hw, int32_t groups, float eps, int32_t silu, float* work,
                            void* sync, void* stream);
/* GroupNorm(+SiLU) in ONE streaming pass over x, from statistics its producer left behind (fd_gemm_ln with colstats_out):
 * stats_k = float2 [segments_k * seg_stride_k, ld_k] per (32-row box, channel pair) of source k; sample b owns box rows
 * [s * seg_stride_k + b * boxes_per_sample_k, + boxes_per_sample_k) of every segment s (segments > 1: the four output phases of a
 * mode-3 upsampling conv).  A tiny finalize launch (one CTA per sample and group, fp64 combine in box order) writes the per-channel
 * scale / shift into work (>= 2 * batch * (c0 + c1) floats); the apply launch reads x once and writes y once.  Same call sites as
 * fd_groupnorm_nhwc (diffusers GroupNorm in ResnetBlock2D / Transformer2DModel / conv_norm_out).                              */
int fd_groupnorm_nhwc_from_stats(const void* x0, int32_t c0, const void* stats0, int64_t ld0, int32_t boxes_per_sample0,
                                 int32_t segments0, int64_t seg_stride0, const void* x1, int32_t c1, const void* stats1, int64_t ld1,
                                 int32_t boxes_per_sample1, int32_t segments1, int64_t seg_stride1, const float* gamma,
                                 const float* beta, void* out, int32_t batch, int64_t hw, int32_t groups, float eps, int32_t silu,
                                 float* work, void* stream);

/* LayerNorm over the last dim (affine), bf16 in/out. Replaces BasicTransformerBlock.norm1/2/3 and the
 * FaithDiff LayerNorm (…extension.py:229-235).                                                             */
int fd_layernorm(const void* x, const float* gamma, const float* beta, void* out, int64_t rows, int32_t cols, float eps,
                 void* stream);
/* out[r, :cols] = alpha * a[r, :cols] + beta * b[r, :cols] on bf16 row views with independent pitches (elements; cols and
 * pitches multiples of 8, 16-byte aligned) — the `out * 0.2 + x` of RRDB (bsrnet_arch.py:72) on channel-slice views.   */
int fd_axpby_bf16(const void* a, int64_t lda, const void* b, int64_t ldb, void* out, int64_t ldo, int64_t rows, int32_t cols,
                  float alpha, float beta, void* stream);
/* nearest-neighbour resize NHWC (Upsample2D before its conv: oh=2h, ow=2w, or the `upsample_size` path) */
int fd_upsample_nearest_nhwc(const void* x, void* out, int32_t batch, int32_t h, int32_t w, int32_t oh, int32_t ow,
                             int32_t c, void* stream);
/* NCHW <-> NHWC layout changes at the module boundary (the reference's forward signatures are NCHW). in_dtype /
 * out_dtype: 0 bf16, 1 fp16, 2 fp32.                                                                       */
int fd_nchw_to_nhwc(const void* x, int32_t in_dtype, void* out_bf16, int32_t batch, int32_t c, int64_t hw, void* stream);
/* same with the channel axis zero-padded to ldc (the UNet's conv_in, 4 -> 320 channels, runs on the tensor cores over an
 * 8-channel NHWC image: one 64-wide K block per tap, zero-filled past channel 8)                                              */
int fd_nchw_to_nhwc_padded(const void* x, int32_t in_dtype, void* out_bf16, int32_t batch, int32_t c, int32_t ldc, int64_t hw,
                           void* stream);
int fd_nhwc_to_nchw(const void* x_bf16, void* out, int32_t out_dtype, int32_t batch, int32_t c, int64_t hw, void* stream);
/* same from a channel-padded NHWC buffer: pixels ldc elements apart, the first c channels are emitted (the tensor-core path of the
 * few-output-channel convs — UNet conv_out 320 -> 4, VAE decoder conv_out 128 -> 3 — computes 32 padded channels)              */
int fd_nhwc_to_nchw_strided(const void* x_nhwc_bf16, int64_t ldc, void* out, int32_t out_dtype, int32_t batch, int32_t c, int64_t hw,
                            void* stream);
/* direct 3x3 convs (pad 1) with a tiny channel count on one side (conv_in 4→C, conv_out C→4, VAE 3→128 / 128→3):
 * w bf16 [cout, 9*cin] (tap-major).  cin <= 8: x per in_mode (0 NHWC bf16, 1 NCHW bf16, 2 NCHW fp32, 3 NCHW fp16),
 * out NHWC bf16 (out_mode 0).  cout <= 8: x NHWC bf16 (in_mode 0), out per out_mode (0 NHWC bf16, 1 NCHW bf16,
 * 2 NCHW fp32, 3 NCHW fp16).                                                                                */
int fd_conv3x3_small(const void* x, int32_t in_mode, const void* w, const float* bias, void* out, int32_t out_mode,
                     int64_t out_ld, int32_t batch, int32_t h, int32_t wd, int32_t cin, int32_t cout, void* stream);
/* 1x1 conv with <= 8 channels on both sides, NCHW, dtype 0 bf16 / 1 fp16 / 2 fp32 (quant_conv / post_quant_conv,
 * with the 1/scaling_factor latent scaling folded in as in_scale): out = W (in_scale * x) + b                  */
int fd_conv1x1_small(const void* x, int32_t in_dtype, const float* w, const float* bias, void* out, int32_t out_dtype,
                     int32_t batch, int32_t cin, int32_t cout, int64_t hw, float in_scale, void* stream);
/* tiny-M linear: out[m, n] = act(sum_k x[m,k] w[n,k] + bias[n]); x fp32 [M,K], w bf16 [N,K], out fp32; act 0/1(SiLU)
 * (time / micro-conditioning embedding MLPs and the 17 time_emb_proj, …extension.py:543-564)               */
int fd_linear_smallm(const float* x, const void* w, const float* bias, const float* addend_or_null, float* out, int32_t m,
                     int32_t n, int32_t k, int32_t act_in_silu, int32_t act_out_silu, void* stream);
/* sinusoidal timestep embedding (flip_sin_to_cos=True, shift 0): row i = [cos(v_i f), sin(v_i f)], fp32, written
 * at out + (i / inner) * outer_ld + (i % inner) * dim  (inner = 6, outer_ld = 2816 packs the SDXL time_ids).   */
int fd_timestep_embedding(const float* values, float* out, int32_t n, int32_t dim, int32_t inner, int64_t outer_ld,
                          void* stream);
/* FaithDiff information-transformer attention across the BATCH axis (nn.MultiheadAttention with
 * batch_first=False fed (B, HW, E), …extension.py:242,576-577): qkv bf16 [G, L, HW, 3E] → out bf16 [G, L, HW, E];
 * softmax over the L (= reference batch, <= 4) axis per pixel and head; G independent groups (batched tiles).  */
int fd_batch_axis_attention(const void* qkv, void* out, int32_t groups, int32_t seq, int64_t hw, int32_t embed,
                            int32_t heads, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Tiled-latent kernels (pipeline_FaithDiff_tlc.py:101-252, :1324-1409)
 * --------------------------------------------------------------------------------------------------------- */
/* grids(): gather T tiles [T, C, k1, k2] (NCHW, any of bf16/fp16/fp32 via elem_bytes) from a canvas [1,C,h,w]. */
int fd_tile_gather(const void* canvas, void* tiles, const int32_t* origins_ij, int32_t ntiles, int32_t c, int32_t h,
                   int32_t w, int32_t k1, int32_t k2, int32_t elem_bytes, void* stream);
/* grids_inverse(): canvas[1,C,h,w] fp32 = sum_t fp32(fp64(tile_t) * w64) / sum_t w  in row-major tile order,
 * bit-exact with the reference's fp16/bf16 × fp64 → fp32 accumulation. tile_dtype: 0 bf16, 1 fp16, 2 fp32.    */
int fd_tile_blend(const void* tiles, int32_t tile_dtype, const double* weights, const int32_t* origins_ij,
                  int32_t ntiles, int32_t c, int32_t h, int32_t w, int32_t k1, int32_t k2, float* canvas_f32,
                  void* canvas_lowp, int32_t lowp_dtype, void* stream);
/* fused per-step tail: CFG combine + Euler step per tile (diffusers rounding order) + Gaussian blend + cast
 * + re-tile + scale_model_input for the next step.  eps: UNet outputs NCHW [T,2,C,k1,k2] (uncond, text),
 * x_tiles: current latents [T,C,k1,k2] of x_dtype.  Writes canvas_f32, canvas_lowp, next_tiles (x_dtype), next_scaled (dtype).
 * dtype = the model dtype (eps, and the rounding of the CFG arithmetic, pipeline…:1365); x_dtype = dtype reproduces the
 * reference's per-step cast of the latents to the model dtype (:1373 `.to(model_output.dtype)`, :1409) bit for bit; x_dtype = 2
 * keeps the 2 MB latent state in fp32 between steps — the pipeline's default in bf16, whose 8-bit mantissa otherwise lets the
 * iterated state random-walk away from the reference (fp16 has three more bits).                                              */
int fd_cfg_euler_blend(const void* eps, const void* x_tiles, int32_t dtype, int32_t x_dtype, const double* weights,
                       const int32_t* origins_ij, int32_t ntiles, int32_t c, int32_t h, int32_t w, int32_t k1,
                       int32_t k2, float guidance, float sigma, float sigma_next, float next_scale_div,
                       float* canvas_f32, void* canvas_lowp, void* next_tiles, void* next_scaled, void* stream);
/* accumulate-only variant for the multi-GPU path: adds this rank's tiles (weighted; with eps != NULL the CFG +
 * Euler update is applied first, as in fd_cfg_euler_blend) into an fp32 numerator canvas; after the NCCL
 * all-reduce fd_canvas_finalize divides by the static count map (= this call on all-ones fp32 tiles).        */
int fd_tile_accumulate(const void* tiles, const void* eps_or_null, int32_t dtype, const double* weights,
                       const int32_t* origins_ij, int32_t ntiles, int32_t c, int32_t h, int32_t w, int32_t k1, int32_t k2,
                       float guidance, float sigma, float sigma_next, float* canvas_num, void* stream);
/* finer multi-GPU split into (tile, CFG-branch) units: CFG + Euler are linear in eps, so each unit u deposits
 * w * (unit_coef[u] * eps_u + unit_addx[u] * x_tile) into the fp32 numerator (coef = (1-g)*dt or g*dt; addx = 1 for the
 * uncond unit of a tile).  eps_units [nunits,C,k1,k2]; x_tiles [T,C,k1,k2]; unit_tile maps a unit to its tile.            */
int fd_unit_accumulate(const void* x_tiles, int32_t x_dtype, const void* eps_units, int32_t dtype, const double* weights,
                       const int32_t* origins_ij, const int32_t* unit_tile, const float* unit_coef, const int32_t* unit_addx,
                       int32_t nunits, int32_t c, int32_t h, int32_t w, int32_t k1, int32_t k2, float* canvas_num, void* stream);
int fd_canvas_finalize(const float* canvas_num, const float* count, float* canvas_f32, void* canvas_lowp,
                       int32_t lowp_dtype, int64_t n, void* stream);
/* the N-rank step tail as three launches (the 1-rank tail is one, fd_cfg_euler_blend): fd_unit_numerator = fd_unit_accumulate
 * that OVERWRITES canvas_num with this rank's contribution (no zero-fill pass); after the exchange fd_canvas_finalize_retile
 * divides by the count map, casts, re-tiles (grids() of the next step, pipeline_FaithDiff_tlc.py:1326) and applies
 * scale_model_input (:1336) — canvas_f32 / canvas_lowp optional, next_tiles / next_scaled [T,C,k1,k2] of `dtype`.          */
int fd_unit_numerator(const void* x_tiles, int32_t x_dtype, const void* eps_units, int32_t dtype, const double* weights,
                      const int32_t* origins_ij, const int32_t* unit_tile, const float* unit_coef, const int32_t* unit_addx,
                      int32_t nunits, int32_t c, int32_t h, int32_t w, int32_t k1, int32_t k2, float* canvas_num, void* stream);
int fd_canvas_finalize_retile(const float* canvas_num, const float* count, float* canvas_f32, void* canvas_lowp, int32_t dtype,
                              int32_t x_dtype, const int32_t* origins_ij, int32_t ntiles, int32_t c, int32_t h, int32_t w, int32_t k1, int32_t k2,
                              float next_scale_div, void* next_tiles, void* next_scaled, void* stream);
/* The same exchange as ONE kernel over NVLink peer memory (opt-in, FD_P2P_EXCHANGE=1; peer_exchange.cu): every rank's numerator
 * canvas lives in a region all peers have mapped through CUDA IPC ([canvas parity 0][canvas parity 1][flags],
 * fd_peer_region_bytes(n)); the kernel publishes this rank's step flag to all peers, waits for theirs, sums the canvases in rank
 * order with P2P loads and divides by the count map — the all-reduce and fd_canvas_finalize in one launch, deterministic and
 * identical on all ranks.  regions: host array of `world` device pointers (own region + IPC-opened peers) in rank order;
 * step: a counter != 0 that increases by one per scheduler step; parity = step & 1 selects the canvas slot.               */
int64_t fd_peer_region_bytes(int64_t n_floats);
int fd_peer_alloc(void** ptr, int64_t bytes);
int fd_peer_free(void* ptr);
int fd_peer_ipc_handle(void* ptr, void* handle64);
int fd_peer_ipc_open(const void* handle64, void** ptr);
int fd_peer_ipc_close(void* ptr);
int fd_canvas_exchange_finalize(const void* const* regions, int32_t rank, int32_t world, int32_t parity, uint32_t step, int64_t n,
                                const float* count, float* canvas_f32, void* canvas_lowp, int32_t lowp_dtype, void* stream);
/* EulerDiscreteScheduler.scale_model_input (pipeline…:1336): out = x / sqrt(sigma^2+1), rounded to dtype         */
int fd_scale_div(const void* x, int32_t x_dtype, void* out, int32_t dtype, float div, int64_t n, void* stream);
/* VAE / LQ-encoder seam blend (Encoder.blend_v/blend_h, …extension.py:150-162): in-place linear cross-fade of
 * the first `extent` rows (vertical=1) or columns of b with the last `extent` of a; NCHW, dtype 0/1/2.        */
int fd_seam_blend(const void* a, void* b, int32_t dtype, int32_t c, int32_t ah, int32_t aw, int32_t bh, int32_t bw,
                  int32_t extent, int32_t vertical, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Image pre/post and colour fix
 * --------------------------------------------------------------------------------------------------------- */
/* VaeImageProcessor.preprocess: u8 HWC → NHWC bf16 in [-1,1] ((x/255)*2-1 in fp32) (pipeline…:1246)         */
int fd_image_preprocess(const uint8_t* rgb_u8, void* out_nhwc_bf16, void* out_nchw_f32_or_null, int32_t h, int32_t w,
                        void* stream);
/* VaeImageProcessor.postprocess: NHWC bf16 (c=3) → u8 HWC: round(clamp(x/2+0.5,0,1)*255) (pipeline…:1451)   */
int fd_image_postprocess(const void* x, int32_t x_is_nhwc_bf16, uint8_t* rgb_u8, int32_t h, int32_t w, void* stream);
/* PIL-compatible separable resample of interleaved u8 images, bit-identical to Pillow's ImagingResample (8bpc):
 * the caller's `image.resize((w*scale, h*scale), Image.LANCZOS)` (test_wo_llava.py:78-81, test.py:104-108) without the
 * host round trip.  Two passes with a u8 intermediate, exactly as Pillow orders them (horizontal over the source rows
 * the vertical pass needs, then vertical).  bounds: int32 [out, 2] = (first source index, tap count); taps: int32
 * [out, ksize] 22-bit fixed point — both from the host (faithdiff_dev_b200.image_process.resample_taps).
 * horizontal: in [rows, in_w, channels] -> out [rows, out_w, channels]; vertical: in [*, row_bytes] -> out [out_h, row_bytes]. */
int fd_resample_u8_horizontal(const uint8_t* in, int32_t in_w, int64_t rows, uint8_t* out, int32_t out_w, int32_t channels,
                              const int32_t* bounds, const int32_t* taps, int32_t ksize, void* stream);
int fd_resample_u8_vertical(const uint8_t* in, int64_t row_bytes, uint8_t* out, int32_t out_h, const int32_t* bounds,
                            const int32_t* taps, int32_t ksize, void* stream);
/* utils/color_fix.py:14-27 adain_color_fix and :29-42 wavelet_color_fix on u8 HWC images (fp32 math).
 * work: fp32 scratch of fd_colorfix_work_floats(h, w) elements.                                              */
int64_t fd_colorfix_work_floats(int32_t h, int32_t w);
int fd_adain_color_fix(const uint8_t* target, const uint8_t* source, uint8_t* out, int32_t h, int32_t w, float* work,
                       void* stream);
int fd_wavelet_color_fix(const uint8_t* target, const uint8_t* source, uint8_t* out, int32_t h, int32_t w, float* work,
                         void* stream);
/* PNG encoder: the caller's `out_image.save(path)` (test_wo_llava.py:97, test.py:130; PIL's PngImagePlugin + host zlib there) on
 * the device.  image: u8 [h, w, channels] with `row_stride_bytes` between rows (a cropped view is fine), channels 1 / 2 / 3 / 4 =
 * grey / grey+alpha / RGB / RGBA, 8 bits.  Writes a COMPLETE PNG file (signature, IHDR, IDAT chunks with their CRCs, IEND) into
 * `out`; result[0] = file bytes, result[1] = 1 if `out_capacity` was too small (nothing usable written) — device memory, read
 * it after synchronising the stream.  Adaptive per-row filtering (minimum sum of absolute values), one dynamic-Huffman deflate
 * block per 16 KB of filtered data with run matches at distance 1, byte-aligned between blocks; lossless, decodes with any PNG
 * reader (the test oracle is PIL + zlib).  filtered: scratch of fd_png_filtered_bytes; work: fd_png_work_words 32-bit words,
 * 8-byte aligned; out: 4-byte aligned, fd_png_file_capacity bytes always suffice.  Five launches, no host synchronisation. */
int64_t fd_png_filtered_bytes(int32_t h, int32_t w, int32_t channels);
int64_t fd_png_work_words(int32_t h, int32_t w, int32_t channels);
int64_t fd_png_file_capacity(int32_t h, int32_t w, int32_t channels);
int fd_png_encode(const uint8_t* image, int32_t h, int32_t w, int32_t channels, int64_t row_stride_bytes, uint8_t* filtered,
                  uint32_t* work, uint8_t* out, int64_t out_capacity, int64_t* result, void* stream);

#ifdef __cplusplus
}
#endif
#endif
