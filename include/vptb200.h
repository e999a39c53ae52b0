/* vptb200 -- C ABI of the B200-native JiT/DiT NF4-QLoRA block kernels.
 *
 * The reference (p1atdev/vision-pt) is pure Python and has no FFI of its own; on this path it reaches native code
 * through third-party wheels.  Each entry point below names the reference call it replaces (paths under
 * /root/reference).  All pointers are DEVICE pointers unless noted, every call is asynchronous on `stream`
 * (a cudaStream_t), never allocates or frees, and never synchronises the host.  Return value: 0 = ok, non-zero =
 * error (message from vpt_last_error(), thread-local).  bf16 activations, fp32 statistics.
 *
 * Build: vision_pt_b200/csrc/build.py -> vision_pt_b200/libvptb200.so (nvcc, sm_100a only).
 */
#ifndef VPTB200_H_
#define VPTB200_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vpt_stream_t; /* cudaStream_t */

enum { VPT_BF16 = 0, VPT_F16 = 1, VPT_F32 = 2 };

const char* vpt_last_error(void);
int vpt_abi_version(void);

/* ------------------------------------------------------------------------------------------------ NF4 weight format
 * The tensors bitsandbytes stores for one Linear4bit weight (quant_type nf4, blocksize 64, compress_statistics):
 * `weight` (packed), `weight.absmax` (uint8), `weight.nested_absmax`, `weight.nested_quant_map`, `weight.quant_map`
 * and the nested_offset scalar of `weight.quant_state.bitsandbytes__nf4`
 * (src/modules/quant/bnb.py:76-129, src/modules/quant/functional.py:361-368). */
typedef struct {
  const uint8_t* packed;        /* [(N*K+1)/2]  high nibble = even element of the flattened [N,K] weight */
  const uint8_t* qabsmax;       /* [N*K/64] */
  const float* nested_absmax;   /* [ceil(N*K/64/256)] */
  const float* nested_code;     /* [256] */
  const float* code;            /* [16] */
  float offset;
  int32_t N, K;                 /* out_features, in_features */
  /* ragged in_features (K % 64 != 0): row-aligned copy made once by vpt_nf4_repack; NULL / 0 otherwise */
  const uint8_t* packed_rows;   /* [N, K_pad/2] */
  const float* absmax_f32;      /* [ceil(N*K/64)] decoded statistics */
  int32_t K_pad;                /* K rounded up to a multiple of 64 */
} vpt_nf4_weight;

/* bitsandbytes.functional.dequantize_4bit (called from bnb.nn.Linear4bit.forward -> MatMul4Bit, inherited by
 * BnbLinear4bit, src/modules/quant/bnb.py:37).  out: n elements of out_dtype.  Bit-exact. */
int vpt_nf4_dequant(const vpt_nf4_weight* w, int64_t n, int out_dtype, void* out, vpt_stream_t stream);

/* Load-time repack of a weight whose in_features is not a multiple of 64 (the 2730 / 3413 wide SwiGLU hidden of JiT-L / -H):
 * bitsandbytes packs the flattened tensor, so rows are not byte- or block-aligned.  Produces row-aligned codes and the
 * decoded per-block statistics; dequantised values stay bit-identical.  No reference counterpart (layout only). */
int vpt_nf4_repack(const vpt_nf4_weight* w, uint8_t* packed_rows, float* absmax_f32, int32_t K_pad, vpt_stream_t stream);

/* bitsandbytes.functional.quantize_4bit(A, quant_type="nf4") with compress_statistics=True, as called by
 * quantize_state_dict (src/modules/quant/functional.py:362-368) and Params4bit._quantize on .cuda()
 * (src/modules/quant/bnb.py:122-129).  n % 64 == 0.  absmax_ws: [n/64] fp32 workspace.  offset_out: 1 fp32 (device). */
int vpt_nf4_quantize(const void* w, int w_dtype, int64_t n, const float* nested_code, uint8_t* packed,
                     uint8_t* qabsmax, float* nested_absmax, float* offset_out, float* absmax_ws, vpt_stream_t stream);

/* ------------------------------------------------------------------------------------------------ NF4 + LoRA linear
 * LoRALinear.forward over a BnbLinear4bit base (src/modules/peft/lora.py:92-104):
 *   y = x W^T + bias + Ts lora_up^T (+ residual),   Ts = bf16(scale * x lora_down^T),   scale = alpha / rank
 * and the activation gradient of the same (MatMul4Bit.backward + autograd of the LoRA branch):
 *   dx = dy W + dTs lora_down (+ residual),         dTs = bf16(scale * dy lora_up)
 * rank is 16 (pad smaller ranks with zero rows/columns).  K % 64 == 0, a repacked weight, or w_scratch given.  x/dy/y/dx are row-major with the given
 * leading dimensions (multiples of 8 elements).  `side` receives Ts^T / dTs^T ([16, ld_side] bf16, ld_side >= M and a
 * multiple of 8: the layout vpt_lora_grad_batch reads by TMA) for the parameter gradients. */
typedef struct {
  vpt_nf4_weight w;
  const void* w_bf16;          /* optional: [N,K] bf16 weight used instead of the NF4 tensors (unquantised Linear) */
  const void* bias;            /* [N] bf16 or NULL (forward only) */
  const void* lora_down;       /* [16,K] bf16 (row pitch ld_lora_down) or NULL (LoRA disabled) */
  int64_t ld_lora_down;        /* multiple of 8 elements; columns >= K zero */
  const void* lora_up;         /* [N,16] bf16 */
  float scale;
  const void* in;              /* fwd: x [M,K];  bwd: dy [M,N] */
  int64_t ld_in;
  void* out;                   /* fwd: y [M,N];  bwd: dx [M,K] */
  int64_t ld_out;
  const void* residual;        /* optional, same shape as out */
  int64_t ld_res;
  void* side;                  /* [16, ld_side] bf16 or NULL */
  int32_t M;
  int32_t tile_n;              /* 0 = auto (128 or 192) */
  /* Optional caller-owned workspace of vpt_linear_scratch_bytes(N, K) bytes, 16-byte aligned.  When given with an NF4
   * weight, the weight is dequantised ONCE per call into it (bit-identical values; it stays L2-resident) -- [N, K] for the
   * forward, transposed [K, N] for the backward -- and the CTA-pair tcgen05 main loop reads it by TMA, instead of every
   * CTA re-dequantising its weight tile per 128-row block of x.  This is the large-M (training) path; NULL selects
   * the per-stage prologue dequantiser (small M). */
  void* w_scratch;
  int64_t ld_scratch;          /* ABI 4: row pitch of w_bf16 in elements (0 = in_features); must be a multiple of 8 */
  int64_t scratch_bytes;       /* capacity of w_scratch */
  int64_t ld_side;             /* row pitch of `side` in elements */
  int32_t reuse_scratch;       /* non-zero: w_scratch already holds this weight in this direction's layout (the previous
                                * call on the stream used the same weight and direction): skip the dequantisation */
  /* ABI 5: SwiGLU (src/models/jit/denoiser.py:498-506 and its autograd) fused into the epilogue of the large-M route.
   *   epilogue = 0  out = in W^T + b (+ residual)                                                       (every linear)
   *   epilogue = 1  forward call of w_2:  residual = g = w_1(x) [M,N];  out2 = u = w_2(x) + b,  out = silu(g) * u
   *   epilogue = 2  backward call of w_3: residual = g, in2 = u [M,K];  with da = dy W:  out = dg = da u silu'(g),
   *                 out2 = du = da silu(g)      (da itself is not written)
   * Modes 1 and 2 need w_scratch (the CTA-pair kernel); rounding points are those of the unfused kernels. */
  int32_t epilogue;
  const void* in2;             /* [M, n_out] bf16, pitch ld_in2 (mode 2) */
  int64_t ld_in2;
  void* out2;                  /* [M, n_out] bf16, pitch ld_out2 (modes 1, 2) */
  int64_t ld_out2;
  /* ABI 6: several linears over the SAME input as one forward call (q | k | v, src/models/jit/denoiser.py:351-363).  w_bf16
   * (or the filled w_scratch) holds their [N_i, K] weights stacked to [n_sections * N_i, K] (w.N = the total), bias and
   * lora_up are stacked the same way, lora_down holds 16 rows per section [16 * n_sections, K], side gets 16 rows per section
   * [16 * n_sections, ld_side]; section i's output columns get section i's LoRA pair.  N_i must be a multiple of 128.
   * 0 / 1 = one linear.  Forward calls with epilogue 0 only. */
  int32_t n_sections;
} vpt_linear_args;

int64_t vpt_linear_scratch_bytes(int32_t N, int32_t K);
/* bytes ONE direction needs (what vpt_nf4_dequant_batch writes; a slot filled by it may be this small) */
int64_t vpt_linear_scratch_bytes_dir(int32_t N, int32_t K, int32_t transposed);

/* Fills the workspaces of up to 8 linears in ONE launch (the seven NF4 weights of a transformer block): slot i gets what
 * vpt_nf4lora_linear_fwd (transposed = 0) or _bwd_dx (transposed = 1; lora_down / lora_up given when the linear has
 * LoRA) would dequantise into w_scratch itself; the linear calls then pass that slot with reuse_scratch = 1.  Bit-exact
 * like vpt_nf4_dequant. */
typedef struct {
  vpt_nf4_weight w;
  void* w_scratch;             /* vpt_linear_scratch_bytes(N, K) bytes, 16-byte aligned */
  int64_t scratch_bytes;
  const void* lora_down;       /* [16,K] bf16 pitch ld_lora_down, or NULL */
  int64_t ld_lora_down;
  const void* lora_up;         /* [N,16] bf16 */
} vpt_nf4_dequant_item;
int vpt_nf4_dequant_batch(const vpt_nf4_dequant_item* items, int32_t n_items, int32_t transposed, vpt_stream_t stream);
int vpt_nf4lora_linear_fwd(const vpt_linear_args* a, vpt_stream_t stream);
int vpt_nf4lora_linear_bwd_dx(const vpt_linear_args* a, vpt_stream_t stream);

/* autograd of lora_down / lora_up (src/modules/peft/lora.py:100-104), batched:  out_i += src^T small_i, fp32, i < nsmall.
 *   lora_up.weight.grad   [N,16]: src = dy [M,N], small = Ts  (side of the forward call),  transposed = 0
 *   lora_down.weight.grad [16,K]: src = x  [M,K], small = dTs (side of the backward call), transposed = 1 (ld_out = K)
 * small_t[i] are side tensors in their [16, ld_small] layout.  Linears sharing their input (q/k/v, w_1/w_2) are one item
 * with nsmall = 3 / 2 so the shared activation is read once.  At most 16 items per call (one transformer block). */
typedef struct {
  const void* src;
  int64_t ld_src;
  int32_t M, P;
  int32_t nsmall;              /* 1..3 */
  const void* small_t[3];
  int64_t ld_small;
  float* out[3];
  int32_t transposed;
  int64_t ld_out;
} vpt_lora_grad_item;
int vpt_lora_grad_batch(const vpt_lora_grad_item* items, int32_t n_items, vpt_stream_t stream);

/* ------------------------------------------------------------------------------------------------ attention
 * scaled_dot_product_attention(q, k, v, mask=key padding) (src/modules/attention.py:98-129) as used by
 * Attention.forward (src/models/jit/denoiser.py:351-397); the bool key-padding mask is given as seqlens_k[b] = number
 * of leading valid keys (NULL = all).  Tensors are (batch, token, head, head_dim) with element strides (sb, sl, sh);
 * both [B,H,L,hd] and [B,L,H,hd] memory layouts are accepted.  head_dim 64 (JiT-B/L, SDXL) and 80 (JiT-H) run the tcgen05
 * kernels; 32 / 96 / 128 run CUDA-core kernels. */
typedef struct {
  const void* ptr;
  int64_t sb, sl, sh;
} vpt_attn_tensor;
int vpt_attn_fwd(const vpt_attn_tensor* q, const vpt_attn_tensor* k, const vpt_attn_tensor* v, const vpt_attn_tensor* o,
                 int32_t B, int32_t H, int32_t Lq, int32_t Lk, int32_t head_dim, const int32_t* seqlens_k, float scale,
                 float* lse2 /* [B,H,Lq rounded up to 128] */, vpt_stream_t stream);
/* dq is fp32 and must be zero on entry; lse2 as written by vpt_attn_fwd; delta_ws: [B,H,Lq rounded up to 128] fp32
 * workspace. */
int vpt_attn_bwd(const vpt_attn_tensor* q, const vpt_attn_tensor* k, const vpt_attn_tensor* v, const vpt_attn_tensor* o,
                 const vpt_attn_tensor* d_o, const vpt_attn_tensor* dq_f32, const vpt_attn_tensor* dk,
                 const vpt_attn_tensor* dv, int32_t B, int32_t H, int32_t Lq, int32_t Lk, int32_t head_dim,
                 const int32_t* seqlens_k, float scale, const float* lse2, float* delta_ws, vpt_stream_t stream);

/* ------------------------------------------------------------------------------------------------ norms etc.
 * FP32RMSNorm.forward (src/modules/norm.py:20-27). rstd_out may be NULL. w may be NULL (no affine). */
int vpt_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd_out, int64_t rows, int32_t D, int64_t ldx,
                    int64_t ldy, float eps, vpt_stream_t stream);
/* dx = rmsnorm_bwd(dy) (+ dres); dw (fp32 [D], accumulated) may be NULL; rstd may be NULL (recomputed). */
int vpt_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, const void* dres, void* dx,
                    float* dw, int64_t rows, int32_t D, int64_t ld, float eps, vpt_stream_t stream);
/* q_norm/k_norm + apply_rope (src/models/jit/denoiser.py:98-111,365-373) on [tokens, H, head_dim]; cos_sin:
 * [L, head_dim/2, 2] fp32; head_dim 64, 80, 96 or 128 */
int vpt_qknorm_rope_fwd(const void* x, const void* w, const float* cos_sin, void* y, int64_t tokens, int32_t H,
                        int32_t L, int32_t head_dim, int64_t ldx, int64_t ldy, float eps, vpt_stream_t stream);
int vpt_qknorm_rope_bwd(const void* dy, int32_t dy_is_f32, const void* x, const void* w, const float* cos_sin,
                        void* dx, float* dw, int64_t tokens, int32_t H, int32_t L, int32_t head_dim, int64_t lddy,
                        int64_t ldx, int64_t lddx, float eps, vpt_stream_t stream);
/* SwiGLU.forward gate: a = silu(g) * u (src/models/jit/denoiser.py:502) */
int vpt_swiglu_fwd(const void* g, const void* u, void* a, int64_t rows, int32_t F, int64_t ldg, int64_t ldu,
                   int64_t lda, vpt_stream_t stream);
int vpt_swiglu_bwd(const void* da, const void* g, const void* u, void* dg, void* du, int64_t rows, int32_t F,
                   int64_t ldda, int64_t ldg, int64_t ldu, int64_t lddg, int64_t lddu, vpt_stream_t stream);
/* AdaLayerNormZero modulate: LN(x)*(1+scale[b])+shift[b] (src/models/cogview4/denoiser.py:182-187,
 * src/modules/norm.py:75-83); x [B*L, D] contiguous, scale/shift [B, D]. mean/rstd: [B*L] fp32 saved for backward. */
int vpt_ln_modulate_fwd(const void* x, const void* scale, const void* shift, void* y, float* mean, float* rstd,
                        int64_t rows, int32_t L, int32_t D, float eps, vpt_stream_t stream);
int vpt_ln_modulate_bwd(const void* dy, const void* x, const void* scale, const float* mean, const float* rstd,
                        void* dx, float* dscale, float* dshift, int64_t rows, int32_t L, int32_t D,
                        vpt_stream_t stream);
/* x + h * gate[b] (src/models/cogview4/denoiser.py:401-420) */
int vpt_gate_residual_fwd(const void* x, const void* h, const void* gate, void* y, int64_t rows, int32_t L, int32_t D,
                          vpt_stream_t stream);
int vpt_gate_residual_bwd(const void* dy, const void* h, const void* gate, void* dh, float* dgate, int64_t rows,
                          int32_t L, int32_t D, vpt_stream_t stream);
/* patchify / unpatchify (src/modules/patch.py:17-115; JiT._unpatchify src/models/jit/denoiser.py:828-860).
 * order 0 = (c,py,px), order 1 = (py,px,c); 2-byte elements. */
int vpt_patchify(const void* img, void* patches, int32_t B, int32_t C, int32_t H, int32_t W, int32_t p, int32_t order,
                 vpt_stream_t stream);
int vpt_unpatchify(const void* patches, void* img, int32_t B, int32_t C, int32_t H, int32_t W, int32_t p,
                   int32_t order, vpt_stream_t stream);


/* JiT.forward's per-block context-token refresh (src/models/jit/denoiser.py:1092-1113: the original context tokens are
 * re-appended in front of every block >= context_start_block and the block's outputs for them dropped) with the slots kept
 * in the token buffer: dst[r, 0:row_bytes] = src[r, 0:row_bytes] for rows a pitch apart; src NULL = zero fill (their
 * gradient).  Everything a multiple of 16 bytes. */
int vpt_copy_rows(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t rows, int64_t row_bytes,
                  vpt_stream_t stream);

/* ------------------------------------------------------------------------------------------------ other block families
 * The blocks either side of the JiT block that share its linears / attention.  bf16 tensors,
 * fp32 arithmetic, the reference's bf16 rounding points.
 *
 * vpt_layernorm_*: nn.LayerNorm / FP32LayerNorm WITH affine parameters (src/models/sdxl/denoiser.py:248-250,
 *   src/modules/norm.py:9-17): y = bf16((x - mean) rstd w + b); w / b may be NULL; mean / rstd [rows] fp32 are saved for the
 *   backward, which gives dx and (optionally, fp32 [D], accumulated) dw / db.
 * vpt_gated_act_*: a = bf16(bf16(act(gate)) h) and its backward; kind 0 = SiLU (SwiGLU), 1 = GELU erf (GeGLU,
 *   src/models/sdxl/denoiser.py:175-186: h, gate = the two halves of one projection, hence the row pitches), 2 = GELU tanh.
 * vpt_act_*: y = bf16(act(x)), dx = dy act'(x) over n contiguous elements (CogView4 FeedForward,
 *   src/models/cogview4/denoiser.py:312-343: kind 2).
 * vpt_rope_half: apply_rotary_emb of CogView4 (src/models/cogview4/denoiser.py:203-218) on token-major [B, L, H, hd]:
 *   tokens l >= l0 rotated by table row l - l0 (cos / sin fp32 [S, hd]), others copied; inverse != 0 is the backward.
 * vpt_pope_*: apply_pope (src/models/jit/extension/pope.py:6-38): y[.., 2i] = softplus(x_i) cos(phi_i), y[.., 2i+1] = .. sin(phi_i),
 *   phi = table angle (cos_sin fp32 [L, d, 2]) + learned bias (fp32 [H, d] or NULL); x [B, L, H, d] -> y [B, L, H, 2d].
 * vpt_token_gather: TREAD routing (train/jit/class_to_image_tread.py:73-118): scatter == 0: dst[b, j] = src[b, idx[j]]
 *   (src [B, L_full, D] -> dst [B, n, D]); scatter != 0: dst[b, idx[j]] = src[b, j] (src [B, n, D] -> dst [B, L_full, D]). */
int vpt_layernorm_fwd(const void* x, const void* w, const void* b, void* y, float* mean, float* rstd, int64_t rows, int32_t D,
                      float eps, vpt_stream_t stream);
int vpt_layernorm_bwd(const void* dy, const void* x, const void* w, const float* mean, const float* rstd, void* dx, float* dw,
                      float* db, int64_t rows, int32_t D, vpt_stream_t stream);
int vpt_gated_act_fwd(const void* h, const void* gate, void* a, int64_t rows, int32_t F, int64_t ldh, int64_t ldg, int64_t lda,
                      int32_t kind, vpt_stream_t stream);
int vpt_gated_act_bwd(const void* da, const void* h, const void* gate, void* dh, void* dgate, int64_t rows, int32_t F,
                      int64_t ldda, int64_t ldh, int64_t ldg, int64_t lddh, int64_t lddg, int32_t kind, vpt_stream_t stream);
int vpt_act_fwd(const void* x, void* y, int64_t n, int32_t kind, vpt_stream_t stream);
int vpt_act_bwd(const void* dy, const void* x, void* dx, int64_t n, int32_t kind, vpt_stream_t stream);
int vpt_rope_half(const void* x, const float* cosv, const float* sinv, void* y, int64_t tokens, int32_t L, int32_t H,
                  int32_t head_dim, int32_t l0, int64_t ldx, int64_t ldy, int32_t inverse, vpt_stream_t stream);
int vpt_pope_fwd(const void* x, const float* cos_sin, const float* bias, void* y, int64_t tokens, int32_t L, int32_t H, int32_t d,
                 int64_t ldx, int64_t ldy, vpt_stream_t stream);
int vpt_pope_bwd(const void* dy, const void* x, const float* cos_sin, const float* bias, void* dx, int64_t tokens, int32_t L,
                 int32_t H, int32_t d, int64_t lddy, int64_t ldx, int64_t lddx, vpt_stream_t stream);
int vpt_token_gather(const void* src, const int64_t* idx, void* dst, int32_t B, int64_t L_full, int64_t n, int32_t D,
                     int32_t scatter, vpt_stream_t stream);


/* ------------------------------------------------------------------------------------------------ optimiser / loss
 * accelerator.clip_grad_norm_ + optimizer.step + zero_grad (src/models/for_training.py:98-109,
 * src/trainer/common.py:382-388) over the flat LoRA buffers: param bf16 [n], grad / exp_avg / exp_avg_sq fp32 [n].
 * vpt_grad_sumsq: out[0] += sum((g*scale)^2), out zero on entry.  vpt_adamw_step: torch.optim.AdamW update with
 * g <- g * grad_scale * min(1, max_norm / (sqrt(*sumsq) + 1e-6)) (sumsq NULL = no clipping); *step is the 1-based
 * step number (device scalar, so that a captured graph replays); zero_grad != 0 clears grad afterwards. */
int vpt_grad_sumsq(const float* g, int64_t n, float scale, float* out, vpt_stream_t stream);
int vpt_adamw_step(void* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, float grad_scale, const float* sumsq, float max_norm,
                   const float* step, int32_t zero_grad, vpt_stream_t stream);
/* schedulefree.RAdamScheduleFree.step (the optimiser the shipped YAMLs name: configs/jit/x-loss/config.yml:75;
 * schedulefree 1.4.1 -- absent from this image, restated from the published algorithm) over the same flat buffers, with
 * the same clipping rule as vpt_adamw_step.  param = the y sequence (bf16), z = base sequence (fp32, initialised to
 * the parameters), exp_avg_sq fp32; sched = double[4] {steps done, lr_max, weight_sum, scheduled_lr} (zero-initialised,
 * advanced on the device in double precision, as the package's Python scalars are, so that a captured graph replays),
 * coef = float[8] scratch.  Two launches.
 * vpt_radam_schedulefree_swap: optimizer.eval() (to_eval != 0: p.lerp_(z, 1 - 1/beta1)) / optimizer.train(). */
int vpt_radam_schedulefree_step(void* param, float* grad, float* z, float* exp_avg_sq, int64_t n, double lr, double beta1,
                                double beta2, float eps, float weight_decay, double r, double weight_lr_power,
                                int32_t silent_sgd_phase, float grad_scale, const float* sumsq, float max_norm,
                                double* sched, float* coef, int32_t zero_grad, vpt_stream_t stream);
int vpt_radam_schedulefree_swap(void* param, const float* z, int64_t n, float beta1, int32_t to_eval, vpt_stream_t stream);
/* treat_loss, model_pred "image" (train/jit/class_to_image.py:106-139): mode 0 = MSE(pred, clean), mode 1 = MSE of
 * the velocities (image_to_velocity, src/models/jit/pipeline.py:253-260) with timestep [batch] fp32.  pred bf16,
 * clean/noisy of in_dtype (VPT_BF16/F16/F32); loss_out[0] += mean (zero on entry); dpred (bf16, may be NULL) = dloss/dpred. */
int vpt_flow_loss(const void* pred, const void* clean, const void* noisy, int in_dtype, const float* timestep,
                  int64_t batch, int64_t per_sample, int32_t mode, float clamp_eps, float* loss_out, void* dpred,
                  vpt_stream_t stream);

/* prepare_scaled_noised_latents (src/modules/loss/flow_match.py:60-74) in one pass: noisy = t*x + (1-t)*(randn*noise_scale)
 * (clean_at_zero: the two weights swapped) with the rounding of each of the reference's five elementwise ops in the tensors'
 * dtype (VPT_BF16/F16/F32), so the result is bit-identical to the op-by-op form.  latents / randn / noisy: [batch, per_sample]
 * of `dtype`; timestep [batch] fp32 (rounded to `dtype` first, as `timestep.to(latents.dtype)` does); noisy_bf16 (may be
 * NULL): the bf16 copy the denoiser consumes (ABI 7). */
int vpt_noise_mix(const void* latents, const void* randn, int dtype, const float* timestep, int64_t batch, int64_t per_sample,
                  float noise_scale, int32_t clean_at_zero, void* noisy, void* noisy_bf16, vpt_stream_t stream);
/* y = bf16(x * bf16(scalar[0])), x / y bf16 [n], scalar fp32 on the device: the upstream gradient of the scalar loss applied
 * to vpt_flow_loss' dpred (autograd's `dpred * dloss`) (ABI 7). */
int vpt_scale_by_scalar(const void* x, const float* scalar, void* y, int64_t n, vpt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VPTB200_H_ */
