// extern "C" entry points declared in include/vptb200.h.  Thin argument checking + kernel launches.
#include "../../include/vptb200.h"

#include <stdlib.h>

#include "attn_launch.cuh"
#include "blocks_ext.cuh"
#include "elementwise.cuh"
#include "gemm_launch.cuh"
#include "gemm_pair_launch.cuh"
#include "lora_grad.cuh"
#include "nf4.cuh"
#include "optim.cuh"

using namespace vpt;

static inline cudaStream_t S(vpt_stream_t s) { return static_cast<cudaStream_t>(s); }
static inline unsigned blocks_for(long n, int per_block, long cap = 1 << 20) {
  long b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return static_cast<unsigned>(b);
}

extern "C" const char* vpt_last_error(void) { return last_error().c_str(); }
extern "C" int vpt_abi_version(void) { return 7; }

// ---------------------------------------------------------------------------------------------------- NF4
extern "C" int vpt_nf4_dequant(const vpt_nf4_weight* w, int64_t n, int out_dtype, void* out, vpt_stream_t stream) {
  VPT_REQUIRE(w && out && n > 0, "vpt_nf4_dequant: bad arguments");
  const unsigned grid = blocks_for((n + 7) / 8, 256, 148 * 16);
  switch (out_dtype) {
    case VPT_BF16:
      nf4_dequant_kernel<kDtBf16><<<grid, 256, 0, S(stream)>>>(w->packed, w->qabsmax, w->nested_absmax, w->nested_code, w->code, w->offset, out, n);
      break;
    case VPT_F16:
      nf4_dequant_kernel<kDtF16><<<grid, 256, 0, S(stream)>>>(w->packed, w->qabsmax, w->nested_absmax, w->nested_code, w->code, w->offset, out, n);
      break;
    case VPT_F32:
      nf4_dequant_kernel<kDtF32><<<grid, 256, 0, S(stream)>>>(w->packed, w->qabsmax, w->nested_absmax, w->nested_code, w->code, w->offset, out, n);
      break;
    default:
      return fail("vpt_nf4_dequant: unknown out_dtype");
  }
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int vpt_nf4_repack(const vpt_nf4_weight* w, uint8_t* packed_rows, float* absmax_f32, int32_t K_pad,
                              vpt_stream_t stream) {
  VPT_REQUIRE(w && packed_rows && absmax_f32 && w->packed && w->qabsmax, "vpt_nf4_repack: null pointer");
  VPT_REQUIRE(K_pad % 64 == 0 && K_pad >= w->K && K_pad - w->K < 64, "vpt_nf4_repack: K_pad must be K rounded up to 64");
  const long nb = (static_cast<long>(w->N) * w->K + 63) / 64;
  nf4_absmax_kernel<<<blocks_for(nb, 256), 256, 0, S(stream)>>>(w->qabsmax, w->nested_absmax, w->nested_code, w->offset, absmax_f32, nb);
  nf4_repack_rows_kernel<<<blocks_for(static_cast<long>(w->N) * (K_pad / 2), 256, 148 * 32), 256, 0, S(stream)>>>(w->packed, packed_rows, w->N, w->K,
                                                                                                                    K_pad);
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int vpt_nf4_quantize(const void* w, int w_dtype, int64_t n, const float* nested_code, uint8_t* packed,
                                uint8_t* qabsmax, float* nested_absmax, float* offset_out, float* absmax_ws,
                                vpt_stream_t stream) {
  VPT_REQUIRE(w && packed && qabsmax && nested_absmax && offset_out && absmax_ws, "vpt_nf4_quantize: null pointer");
  VPT_REQUIRE(n > 0 && n % 64 == 0, "vpt_nf4_quantize: element count must be a multiple of 64");
  const long nb = n / 64;
  const unsigned grid = blocks_for(nb * 8, 256);
  switch (w_dtype) {
    case VPT_BF16: nf4_quantize_kernel<kDtBf16><<<grid, 256, 0, S(stream)>>>(w, packed, absmax_ws, nb); break;
    case VPT_F16: nf4_quantize_kernel<kDtF16><<<grid, 256, 0, S(stream)>>>(w, packed, absmax_ws, nb); break;
    case VPT_F32: nf4_quantize_kernel<kDtF32><<<grid, 256, 0, S(stream)>>>(w, packed, absmax_ws, nb); break;
    default: return fail("vpt_nf4_quantize: unknown dtype");
  }
  nf4_mean_kernel<<<1, 1024, 0, S(stream)>>>(absmax_ws, nb, offset_out);
  nf4_nested_quantize_kernel<<<static_cast<unsigned>((nb + 255) / 256), 256, 0, S(stream)>>>(absmax_ws, offset_out, nested_code, qabsmax,
                                                                                              nested_absmax, nb);
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------- linear
extern "C" int64_t vpt_linear_scratch_bytes(int32_t N, int32_t K) {
  const int64_t ldk = (K + 7) / 8 * 8, ldn = (N + 7) / 8 * 8;
  const int64_t fwd = static_cast<int64_t>(N) * ldk;
  const int64_t bwd = static_cast<int64_t>(K) * ldn + 16 * ldn + static_cast<int64_t>(K) * 16;
  return 2 * (fwd > bwd ? fwd : bwd) + 256;
}
// What ONE direction of a weight needs in its workspace slot: [N, ldk] for the forward; [K, ldn] + the transposed LoRA
// copies for the backward.  Forward slots of this size are gap-free for N % 128 == 0, so the slots of q | k | v laid out one
// after the other ARE the stacked [3N, K] weight of the sectioned forward call (vpt_linear_args.n_sections).
extern "C" int64_t vpt_linear_scratch_bytes_dir(int32_t N, int32_t K, int32_t transposed) {
  const int64_t ldk = (K + 7) / 8 * 8, ldn = (N + 7) / 8 * 8;
  const int64_t e = transposed ? static_cast<int64_t>(K) * ldn + 16 * ldn + static_cast<int64_t>(K) * 16 : static_cast<int64_t>(N) * ldk;
  return 2 * e;
}
extern "C" int vpt_nf4_dequant_batch(const vpt_nf4_dequant_item* items, int32_t n_items, int32_t transposed, vpt_stream_t stream) {
  VPT_REQUIRE(items != nullptr && n_items > 0 && n_items <= kDqMaxItems, "vpt_nf4_dequant_batch: 1..8 items");
  DequantBatch bp{};
  int ctas = 0;
  for (int i = 0; i < n_items; ++i) {
    const vpt_nf4_dequant_item& s = items[i];
    const int N = s.w.N, K = s.w.K;
    VPT_REQUIRE(N > 0 && K > 0 && s.w.packed && s.w.qabsmax && s.w.nested_absmax && s.w.nested_code && s.w.code && s.w_scratch,
                "vpt_nf4_dequant_batch: NF4 tensors / workspace missing");
    VPT_REQUIRE((reinterpret_cast<uintptr_t>(s.w_scratch) & 15) == 0 && s.scratch_bytes >= vpt_linear_scratch_bytes_dir(N, K, transposed),
                "vpt_nf4_dequant_batch: w_scratch must be 16-byte aligned and hold vpt_linear_scratch_bytes_dir(N, K, transposed) bytes");
    DequantItem& d = bp.items[i];
    d.packed = s.w.packed; d.qabsmax = s.w.qabsmax; d.nested_absmax = s.w.nested_absmax; d.nested_code = s.w.nested_code;
    d.code = s.w.code; d.offset = s.w.offset; d.N = N; d.K = K;
    d.out = static_cast<__nv_bfloat16*>(s.w_scratch);
    d.transposed = transposed ? 1 : 0;
    d.ld = transposed ? (N + 7) / 8 * 8 : (K + 7) / 8 * 8;
    d.tiles_k = (K + 255) / 256;                  // a CTA walks 4 consecutive 64-wide k-tiles of one 64-row band
    d.num_tiles = ((N + 63) / 64) * d.tiles_k;
    d.cta_begin = ctas;
    const bool lora = transposed && s.lora_down != nullptr;
    if (lora) VPT_REQUIRE(s.lora_up != nullptr && s.ld_lora_down >= K, "vpt_nf4_dequant_batch: bad LoRA arguments");
    d.up = lora ? static_cast<const __nv_bfloat16*>(s.lora_up) : nullptr;
    d.down = lora ? static_cast<const __nv_bfloat16*>(s.lora_down) : nullptr;
    d.ldd = s.ld_lora_down;
    ctas += d.num_tiles + (lora ? 4 : 0);
  }
  bp.n_items = n_items;
  VPT_CUDA_OK(launch_pdl(nf4_dequant_batch_kernel, dim3(ctas), dim3(256), 0, S(stream), bp));
  return 0;
}

static int linear_common(const vpt_linear_args* a, bool bwd, cudaStream_t stream) {
  VPT_REQUIRE(a && a->in && a->out, "vpt_nf4lora_linear: null pointer");
  const int N = a->w.N, K = a->w.K;
  VPT_REQUIRE(N > 0 && K > 0 && a->M > 0, "vpt_nf4lora_linear: bad shape");
  VPT_REQUIRE(a->ld_in % 8 == 0 && a->ld_out % 8 == 0, "vpt_nf4lora_linear: leading dimensions must be multiples of 8");
  VPT_REQUIRE(a->side == nullptr || (a->ld_side >= a->M && a->ld_side % 8 == 0), "vpt_nf4lora_linear: side needs ld_side >= M, a multiple of 8");
  const bool via_scratch = a->w_bf16 == nullptr && a->w_scratch != nullptr;
  // row pitch of a bf16 weight (ld_scratch doubles as that since ABI 4): a ragged in_features (3413, 2730) is served
  // from a copy whose rows are padded to a multiple of 8 elements -- TMA zero-fills past K
  const long ld_wb = (a->w_bf16 != nullptr && a->ld_scratch > 0) ? a->ld_scratch : K;
  if (a->w_bf16 != nullptr) VPT_REQUIRE(ld_wb >= K && ld_wb % 8 == 0, "vpt_nf4lora_linear: a bf16 weight needs a row pitch (ld_scratch, 0 = in_features) that is a multiple of 8");
  const bool lora = a->lora_down != nullptr;
  if (lora) VPT_REQUIRE(a->lora_up != nullptr && a->ld_lora_down % 8 == 0 && a->ld_lora_down >= K, "vpt_nf4lora_linear: bad LoRA arguments");
  static const bool use_pairs = getenv("VPT_NO_PAIR") == nullptr;   // A/B switch for profiling the 1-CTA kernel
  if (use_pairs && (via_scratch || (a->w_bf16 != nullptr && !bwd))) {
    // Large-M route: CTA-pair kernel over a K-major bf16 weight.  An NF4 weight is dequantised once per call into the
    // caller's L2-resident workspace -- as [N, K] for the forward, TRANSPOSED [K, N] (plus the two transposed LoRA
    // matrices) for the backward, so that both directions run the same kernel.
    PairLaunch g{};
    g.act = a->in; g.lda = static_cast<int>(a->ld_in);
    g.out = a->out; g.ldd = static_cast<int>(a->ld_out);
    g.bn = a->tile_n;
    g.p.M = a->M; g.p.NO = bwd ? K : N; g.p.R = bwd ? N : K;
    g.p.bias = bwd ? nullptr : static_cast<const __nv_bfloat16*>(a->bias);
    g.p.residual = static_cast<const __nv_bfloat16*>(a->residual); g.p.ldr = static_cast<int>(a->ld_res);
    g.p.scale = a->scale;
    g.p.side = static_cast<__nv_bfloat16*>(a->side);
    g.p.ld_side = static_cast<long>(a->ld_side);
    g.p.sec_n = 0;
    if (a->n_sections > 1) {
      VPT_REQUIRE(!bwd && a->epilogue == 0 && N % a->n_sections == 0 && (N / a->n_sections) % 128 == 0,
                  "vpt_nf4lora_linear: n_sections is a forward option; every section must be a multiple of 128 columns wide");
      g.p.sec_n = N / a->n_sections;
    }
    g.epi = a->epilogue;
    g.out2 = a->out2; g.ldd2 = static_cast<int>(a->ld_out2);
    g.in2 = a->in2; g.ldr2 = static_cast<int>(a->ld_in2);
    if (a->epilogue != 0) {
      // (a dense bf16 weight has no backward on this route: its caller passes the transposed weight to the forward entry
      // point, so mode 2 is accepted there)
      VPT_REQUIRE((a->epilogue == 1 && !bwd) || (a->epilogue == 2 && (bwd || a->w_bf16 != nullptr)),
                  "vpt_nf4lora_linear: epilogue 1 is a forward mode, 2 a backward mode");
      VPT_REQUIRE(a->residual && a->out2 && a->ld_out2 % 8 == 0 && a->ld_res % 8 == 0 && (a->epilogue == 1 || (a->in2 && a->ld_in2 % 8 == 0)),
                  "vpt_nf4lora_linear: fused SwiGLU epilogue needs residual (g), out2 (and in2 = u in mode 2) with pitches that are multiples of 8");
    }
    if (via_scratch) {
      VPT_REQUIRE(a->w.packed && a->w.qabsmax && a->w.nested_absmax && a->w.nested_code && a->w.code, "vpt_nf4lora_linear: NF4 tensors missing");
      VPT_REQUIRE((reinterpret_cast<uintptr_t>(a->w_scratch) & 15) == 0 &&
                      a->scratch_bytes >= (a->reuse_scratch ? vpt_linear_scratch_bytes_dir(N, K, bwd) : vpt_linear_scratch_bytes(N, K)),
                  "vpt_nf4lora_linear: w_scratch must be 16-byte aligned and hold vpt_linear_scratch_bytes(N, K) bytes (a pre-filled slot: _dir)");
      __nv_bfloat16* ws = static_cast<__nv_bfloat16*>(a->w_scratch);
      const long n = static_cast<long>(N) * K;
      if (!bwd) {
        const long ldk = (K + 7) / 8 * 8;
        if (!a->reuse_scratch)
          nf4_dequant_pitched_kernel<<<blocks_for((n + 7) / 8, 256, 148 * 16), 256, 0, stream>>>(
              a->w.packed, a->w.qabsmax, a->w.nested_absmax, a->w.nested_code, a->w.code, a->w.offset, ws, n, K, ldk);
        g.w = ws; g.ldw = ldk;
        g.p_rows = a->lora_down; g.ldp = a->ld_lora_down;
        g.p.q_rows = static_cast<const __nv_bfloat16*>(a->lora_up);
      } else {
        const long ldn = (N + 7) / 8 * 8;
        __nv_bfloat16* upT = ws + static_cast<long>(K) * ldn;       // [16, ldn]
        __nv_bfloat16* downT = upT + 16 * ldn;                      // [K, 16]
        const int tiles_k = (K + 63) / 64, tiles = ((N + 63) / 64) * tiles_k;
        if (!a->reuse_scratch)
          nf4_dequant_transposed_kernel<<<tiles + (lora ? 4 : 0), 256, 0, stream>>>(
            a->w.packed, a->w.qabsmax, a->w.nested_absmax, a->w.nested_code, a->w.code, a->w.offset, ws, N, K, ldn, tiles_k,
            tiles, static_cast<const __nv_bfloat16*>(a->lora_up), upT, ldn, static_cast<const __nv_bfloat16*>(a->lora_down),
            static_cast<long>(a->ld_lora_down), downT);
        g.w = ws; g.ldw = ldn;
        g.p_rows = lora ? upT : nullptr; g.ldp = ldn;
        g.p.q_rows = downT;
      }
      VPT_CUDA_OK(cudaGetLastError());
    } else {
      g.w = a->w_bf16; g.ldw = ld_wb;
      g.p_rows = a->lora_down; g.ldp = a->ld_lora_down;
      g.p.q_rows = static_cast<const __nv_bfloat16*>(a->lora_up);
    }
    return launch_pair(g, stream);
  }
  VPT_REQUIRE(a->epilogue == 0 && a->n_sections <= 1, "vpt_nf4lora_linear: the fused SwiGLU epilogues and n_sections exist on the large-M (CTA-pair) route only");
  const void* w_dense = a->w_bf16;
  long ldw = a->w_bf16 != nullptr ? ld_wb : 0;
  if (via_scratch) {
    const long n = static_cast<long>(N) * K;
    ldw = (K + 7) / 8 * 8;
    nf4_dequant_pitched_kernel<<<blocks_for((n + 7) / 8, 256, 148 * 16), 256, 0, stream>>>(
        a->w.packed, a->w.qabsmax, a->w.nested_absmax, a->w.nested_code, a->w.code, a->w.offset,
        static_cast<__nv_bfloat16*>(a->w_scratch), n, K, ldw);
    VPT_CUDA_OK(cudaGetLastError());
    w_dense = a->w_scratch;
  }
  const bool nf4 = w_dense == nullptr;
  const bool ragged = nf4 && a->w.packed_rows != nullptr;
  if (nf4) {
    VPT_REQUIRE(a->w.nested_code && a->w.code, "vpt_nf4lora_linear: NF4 code tables missing");
    if (ragged) {
      VPT_REQUIRE(a->w.absmax_f32 && a->w.K_pad % 64 == 0 && a->w.K_pad >= K, "vpt_nf4lora_linear: bad repacked weight");
    } else {
      VPT_REQUIRE(a->w.packed && a->w.qabsmax && a->w.nested_absmax, "vpt_nf4lora_linear: NF4 tensors missing");
      VPT_REQUIRE(K % 64 == 0, "vpt_nf4lora_linear: in_features must be a multiple of 64 (repack ragged weights with vpt_nf4_repack)");
    }
  }
  GemmLaunch g{};
  g.bwd = bwd; g.nf4 = nf4; g.lora = lora; g.bn = a->tile_n;
  g.act = a->in; g.lda = static_cast<int>(a->ld_in); g.w_bf16 = w_dense; g.ldw = ldw;
  g.p.M = a->M; g.p.NO = bwd ? K : N; g.p.R = bwd ? N : K;
  g.p.D = static_cast<__nv_bfloat16*>(a->out); g.p.ldd = static_cast<int>(a->ld_out);
  g.p.bias = bwd ? nullptr : static_cast<const __nv_bfloat16*>(a->bias);
  g.p.residual = static_cast<const __nv_bfloat16*>(a->residual); g.p.ldr = static_cast<int>(a->ld_res);
  g.p.w = Nf4Weight{a->w.packed, a->w.qabsmax, a->w.nested_absmax, a->w.nested_code, a->w.code, a->w.offset, N, K,
                    a->w.packed_rows, a->w.absmax_f32, a->w.K_pad};
  g.p.lora_down = static_cast<const __nv_bfloat16*>(a->lora_down);
  g.p.ld_down = static_cast<int>(a->ld_lora_down);
  g.p.lora_up = static_cast<const __nv_bfloat16*>(a->lora_up);
  g.p.scale = a->scale;
  g.p.side = static_cast<__nv_bfloat16*>(a->side);
  g.p.ld_side = static_cast<long>(a->ld_side);
  return launch_gemm(g, stream);
}
extern "C" int vpt_nf4lora_linear_fwd(const vpt_linear_args* a, vpt_stream_t stream) { return linear_common(a, false, S(stream)); }
extern "C" int vpt_nf4lora_linear_bwd_dx(const vpt_linear_args* a, vpt_stream_t stream) { return linear_common(a, true, S(stream)); }

extern "C" int vpt_lora_grad_batch(const vpt_lora_grad_item* items, int32_t n_items, vpt_stream_t stream) {
  VPT_REQUIRE(items != nullptr && n_items > 0 && n_items <= kLgMaxItems, "vpt_lora_grad_batch: 1..16 items");
  LoraGradDesc d[kLgMaxItems];
  for (int i = 0; i < n_items; ++i) {
    const vpt_lora_grad_item& s = items[i];
    VPT_REQUIRE(s.src && s.M > 0 && s.P > 0 && s.nsmall >= 1 && s.nsmall <= 3, "vpt_lora_grad_batch: bad item");
    d[i].src = s.src; d[i].ld_src = s.ld_src; d[i].M = s.M; d[i].P = s.P; d[i].nsmall = s.nsmall;
    for (int j = 0; j < 3; ++j) {
      d[i].small_t[j] = s.small_t[j];
      d[i].out[j] = s.out[j];
    }
    d[i].ld_small = s.ld_small; d[i].transposed = s.transposed; d[i].ld_out = s.ld_out;
  }
  return launch_lora_grad_batch(d, n_items, S(stream));
}

// ---------------------------------------------------------------------------------------------------- attention
static inline AttnTensor AT(const vpt_attn_tensor* t) { return AttnTensor{t->ptr, static_cast<long>(t->sb), static_cast<long>(t->sl), static_cast<long>(t->sh)}; }
extern "C" int vpt_attn_fwd(const vpt_attn_tensor* q, const vpt_attn_tensor* k, const vpt_attn_tensor* v,
                            const vpt_attn_tensor* o, int32_t B, int32_t H, int32_t Lq, int32_t Lk, int32_t head_dim,
                            const int32_t* seqlens_k, float scale, float* lse2, vpt_stream_t stream) {
  VPT_REQUIRE(q && k && v && o && lse2 && B > 0 && H > 0 && Lq > 0 && Lk > 0, "vpt_attn_fwd: bad arguments");
  if (head_dim == 64) return launch_attn_fwd(AT(q), AT(k), AT(v), AT(o), B, H, Lq, Lk, seqlens_k, scale, lse2, S(stream));
  // head_dim 80 (JiT-H): the same tcgen05 kernels, templated on the head dimension
  if (head_dim == 80)
    return launch_attn_fwd(AT(q), AT(k), AT(v), AT(o), B, H, Lq, Lk, seqlens_k, scale, lse2, S(stream), 80);
  return launch_attn_simple_fwd(AT(q), AT(k), AT(v), AT(o), B, H, Lq, Lk, head_dim, seqlens_k, scale, lse2, S(stream));
}
extern "C" int vpt_attn_bwd(const vpt_attn_tensor* q, const vpt_attn_tensor* k, const vpt_attn_tensor* v,
                            const vpt_attn_tensor* o, const vpt_attn_tensor* d_o, const vpt_attn_tensor* dq_f32,
                            const vpt_attn_tensor* dk, const vpt_attn_tensor* dv, int32_t B, int32_t H, int32_t Lq,
                            int32_t Lk, int32_t head_dim, const int32_t* seqlens_k, float scale, const float* lse2,
                            float* delta_ws, vpt_stream_t stream) {
  VPT_REQUIRE(q && k && v && o && d_o && dq_f32 && dk && dv && lse2 && delta_ws, "vpt_attn_bwd: null pointer");
  if (head_dim == 64 || head_dim == 80)
    return launch_attn_bwd(AT(q), AT(k), AT(v), AT(o), AT(d_o), AT(dq_f32), AT(dk), AT(dv), B, H, Lq, Lk, seqlens_k, scale, lse2,
                           delta_ws, S(stream), head_dim);
  return launch_attn_simple_bwd(AT(q), AT(k), AT(v), AT(o), AT(d_o), AT(dq_f32), AT(dk), AT(dv), B, H, Lq, Lk, head_dim, seqlens_k,
                                scale, lse2, delta_ws, S(stream));
}

// ---------------------------------------------------------------------------------------------------- elementwise
#define BF(p) static_cast<const __nv_bfloat16*>(p)
#define BFM(p) static_cast<__nv_bfloat16*>(p)
extern "C" int vpt_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd_out, int64_t rows, int32_t D, int64_t ldx,
                               int64_t ldy, float eps, vpt_stream_t stream) {
  VPT_REQUIRE(x && y && rows > 0 && D > 0 && D % 8 == 0 && D <= 2048 && ldx % 8 == 0 && ldy % 8 == 0, "vpt_rmsnorm_fwd: bad arguments");
  VPT_CUDA_OK(launch_pdl(rmsnorm_fwd_kernel, dim3(blocks_for(rows, kEwThreads / 32, 1L << 30)), dim3(kEwThreads), 0, S(stream), BF(x), BF(w), BFM(y), rstd_out, static_cast<long>(rows), D, static_cast<long>(ldx), static_cast<long>(ldy), eps));
  return 0;
}
extern "C" int vpt_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, const void* dres, void* dx,
                               float* dw, int64_t rows, int32_t D, int64_t ld, float eps, vpt_stream_t stream) {
  VPT_REQUIRE(dy && x && dx && rows > 0 && D % 8 == 0 && D <= 2048 && ld % 8 == 0, "vpt_rmsnorm_bwd: bad arguments");
  const unsigned grid = blocks_for(rows, kEwThreads / 32, 1L << 30);
  const int nch = (D / 8 + 31) / 32;
#define VPT_RMS_BWD(CH) VPT_CUDA_OK(launch_pdl(rmsnorm_bwd_kernel<CH>, dim3(grid), dim3(kEwThreads), 0, S(stream), BF(dy), BF(x), BF(w), rstd, BF(dres), BFM(dx), dw, static_cast<long>(rows), D, static_cast<long>(ld), eps))
  if (nch <= 1) VPT_RMS_BWD(1);
  else if (nch <= 3) VPT_RMS_BWD(3);
  else if (nch <= 4) VPT_RMS_BWD(4);
  else if (nch <= 5) VPT_RMS_BWD(5);
  else VPT_RMS_BWD(8);
#undef VPT_RMS_BWD
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int vpt_qknorm_rope_fwd(const void* x, const void* w, const float* cos_sin, void* y, int64_t tokens, int32_t H,
                                   int32_t L, int32_t head_dim, int64_t ldx, int64_t ldy, float eps, vpt_stream_t stream) {
  VPT_REQUIRE(x && w && cos_sin && y && tokens > 0 && H > 0 && L > 0 && ldx % 8 == 0 && ldy % 8 == 0, "vpt_qknorm_rope_fwd: bad arguments");
  VPT_REQUIRE(head_dim == 64 || head_dim == 80 || head_dim == 96 || head_dim == 128, "vpt_qknorm_rope_fwd: head_dim must be 64, 80, 96 or 128");
  const int hp = head_dim == 64 ? (H % 4 == 0 ? 4 : (H % 2 == 0 ? 2 : 1)) : (H % 2 == 0 ? 2 : 1);
  const int gl = head_dim == 64 ? 8 : 16;
  const dim3 grid(blocks_for(tokens * (H / hp) * gl, 256, 1L << 30));
#define VPT_QK_FWD(HP, HD) VPT_CUDA_OK(launch_pdl(qknorm_rope_fwd_kernel<HP, HD>, grid, dim3(256), 0, S(stream), BF(x), BF(w), cos_sin, BFM(y), tokens, H, L, ldx, ldy, eps))
#define VPT_QK_FWD_HD(HD) do { if (hp == 2) VPT_QK_FWD(2, HD); else VPT_QK_FWD(1, HD); } while (0)
  if (head_dim == 64) {
    if (hp == 4) VPT_QK_FWD(4, 64); else VPT_QK_FWD_HD(64);
  } else if (head_dim == 80) VPT_QK_FWD_HD(80);
  else if (head_dim == 96) VPT_QK_FWD_HD(96);
  else VPT_QK_FWD_HD(128);
#undef VPT_QK_FWD_HD
#undef VPT_QK_FWD
  return 0;
}
extern "C" int vpt_qknorm_rope_bwd(const void* dy, int32_t dy_is_f32, const void* x, const void* w, const float* cos_sin,
                                   void* dx, float* dw, int64_t tokens, int32_t H, int32_t L, int32_t head_dim, int64_t lddy,
                                   int64_t ldx, int64_t lddx, float eps, vpt_stream_t stream) {
  VPT_REQUIRE(dy && x && w && cos_sin && dx && tokens > 0 && H > 0 && L > 0, "vpt_qknorm_rope_bwd: bad arguments");
  VPT_REQUIRE(head_dim == 64 || head_dim == 80 || head_dim == 96 || head_dim == 128, "vpt_qknorm_rope_bwd: head_dim must be 64, 80, 96 or 128");
  const int hp = H % 2 == 0 ? 2 : 1;            // 4 heads per lane group cost occupancy here (measured slower)
  const int gl = head_dim == 64 ? 8 : 16;
  const dim3 grid(blocks_for(tokens * (H / hp) * gl, 256, 1L << 30));
#define VPT_QK_BWD(F32, HP, HD) VPT_CUDA_OK(launch_pdl(qknorm_rope_bwd_kernel<F32, HP, HD>, grid, dim3(256), 0, S(stream), dy, BF(x), BF(w), cos_sin, BFM(dx), dw, tokens, H, L, lddy, ldx, lddx, eps))
#define VPT_QK_BWD_HD(HD) do { \
    if (dy_is_f32) { if (hp == 2) VPT_QK_BWD(true, 2, HD); else VPT_QK_BWD(true, 1, HD); } \
    else { if (hp == 2) VPT_QK_BWD(false, 2, HD); else VPT_QK_BWD(false, 1, HD); } } while (0)
  if (head_dim == 64) VPT_QK_BWD_HD(64);
  else if (head_dim == 80) VPT_QK_BWD_HD(80);
  else if (head_dim == 96) VPT_QK_BWD_HD(96);
  else VPT_QK_BWD_HD(128);
#undef VPT_QK_BWD_HD
#undef VPT_QK_BWD
  return 0;
}
extern "C" int vpt_swiglu_fwd(const void* g, const void* u, void* a, int64_t rows, int32_t F, int64_t ldg, int64_t ldu,
                              int64_t lda, vpt_stream_t stream) {
  const int64_t f8 = (F + 7) / 8 * 8;
  VPT_REQUIRE(g && u && a && rows > 0 && F > 0 && ldg % 8 == 0 && ldu % 8 == 0 && lda % 8 == 0 && ldg >= f8 && ldu >= f8 && lda >= f8,
              "vpt_swiglu_fwd: bad arguments (row pitches must be multiples of 8 and cover F rounded up to 8)");
  VPT_CUDA_OK(launch_pdl(swiglu_fwd_kernel, dim3(blocks_for(rows * (f8 / 8), 256, 148 * 32)), dim3(256), 0, S(stream), BF(g), BF(u), BFM(a), rows, F, ldg, ldu, lda));
  return 0;
}
extern "C" int vpt_swiglu_bwd(const void* da, const void* g, const void* u, void* dg, void* du, int64_t rows, int32_t F,
                              int64_t ldda, int64_t ldg, int64_t ldu, int64_t lddg, int64_t lddu, vpt_stream_t stream) {
  const int64_t f8 = (F + 7) / 8 * 8;
  VPT_REQUIRE(da && g && u && dg && du && rows > 0 && F > 0 && ldda >= f8 && ldg >= f8 && ldu >= f8 && lddg >= f8 && lddu >= f8 &&
                  (ldda | ldg | ldu | lddg | lddu) % 8 == 0,
              "vpt_swiglu_bwd: bad arguments");
  VPT_CUDA_OK(launch_pdl(swiglu_bwd_kernel, dim3(blocks_for(rows * (f8 / 8), 256, 148 * 32)), dim3(256), 0, S(stream), BF(da), BF(g), BF(u), BFM(dg), BFM(du), rows, F, ldda, ldg, ldu, lddg, lddu));
  return 0;
}
// chunks per lane of the row-per-warp LayerNorm kernels: D <= 1024, 2048, 4096
#define VPT_LN_DISPATCH(D_, CALL) do { if ((D_) <= 1024) { CALL(4); } else if ((D_) <= 2048) { CALL(8); } else { CALL(16); } } while (0)
extern "C" int vpt_ln_modulate_fwd(const void* x, const void* scale, const void* shift, void* y, float* mean, float* rstd,
                                   int64_t rows, int32_t L, int32_t D, float eps, vpt_stream_t stream) {
  VPT_REQUIRE(x && scale && shift && y && rows > 0 && L > 0 && D % 8 == 0 && D <= 4096, "vpt_ln_modulate_fwd: bad arguments");
#define VPT_CALL(CH) VPT_CUDA_OK(launch_pdl(ln_modulate_fwd_kernel<CH>, dim3(blocks_for(rows, kEwThreads / 32, 1L << 30)), dim3(kEwThreads), 0, S(stream), BF(x), BF(scale), BF(shift), BFM(y), mean, rstd, static_cast<long>(rows), L, D, eps))
  VPT_LN_DISPATCH(D, VPT_CALL);
#undef VPT_CALL
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int vpt_ln_modulate_bwd(const void* dy, const void* x, const void* scale, const float* mean, const float* rstd,
                                   void* dx, float* dscale, float* dshift, int64_t rows, int32_t L, int32_t D,
                                   vpt_stream_t stream) {
  VPT_REQUIRE(dy && x && scale && mean && rstd && dx && rows > 0 && D % 8 == 0 && D <= 4096, "vpt_ln_modulate_bwd: bad arguments");
  VPT_REQUIRE((dscale == nullptr) == (dshift == nullptr), "vpt_ln_modulate_bwd: dscale and dshift go together");
  static const int wide_min = getenv("VPT_LN_BWD_WIDE_MIN_D") ? atoi(getenv("VPT_LN_BWD_WIDE_MIN_D")) : 2049;   // A/B switch
  if (D >= wide_min) {
    VPT_CUDA_OK(launch_pdl(ln_modulate_bwd_wide_kernel, dim3(blocks_for(rows, kEwThreads / 32, 1L << 30)), dim3(kEwThreads), 0, S(stream), BF(dy), BF(x),
                           BF(scale), mean, rstd, BFM(dx), dscale, dshift, static_cast<long>(rows), L, D, 0));
    return 0;
  }
#define VPT_CALL(CH) VPT_CUDA_OK(launch_pdl(ln_modulate_bwd_kernel<CH>, dim3(blocks_for(rows, kEwThreads / 32, 1L << 30)), dim3(kEwThreads), 0, S(stream), BF(dy), BF(x), BF(scale), mean, rstd, BFM(dx), dscale, dshift, static_cast<long>(rows), L, D))
  VPT_LN_DISPATCH(D, VPT_CALL);
#undef VPT_CALL
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int vpt_gate_residual_fwd(const void* x, const void* h, const void* gate, void* y, int64_t rows, int32_t L,
                                     int32_t D, vpt_stream_t stream) {
  VPT_REQUIRE(x && h && gate && y && rows > 0 && D % 8 == 0, "vpt_gate_residual_fwd: bad arguments");
  gate_residual_fwd_kernel<<<blocks_for(rows * (D / 8), 256, 148 * 32), 256, 0, S(stream)>>>(BF(x), BF(h), BF(gate), BFM(y), rows, L, D);
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int vpt_gate_residual_bwd(const void* dy, const void* h, const void* gate, void* dh, float* dgate, int64_t rows,
                                     int32_t L, int32_t D, vpt_stream_t stream) {
  VPT_REQUIRE(dy && h && gate && dh && rows > 0 && D % 8 == 0, "vpt_gate_residual_bwd: bad arguments");
  gate_residual_bwd_kernel<<<blocks_for(rows * (D / 8), 256, 148 * 32), 256, 0, S(stream)>>>(BF(dy), BF(h), BF(gate), BFM(dh), dgate, rows, L, D);
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}
// ---- block families either side of the JiT block (SDXL, CogView4, JiT extensions): blocks_ext.cuh
extern "C" int vpt_layernorm_fwd(const void* x, const void* w, const void* b, void* y, float* mean, float* rstd, int64_t rows,
                                 int32_t D, float eps, vpt_stream_t stream) {
  VPT_REQUIRE(x && y && rows > 0 && D > 0 && D % 8 == 0 && D <= 4096 && ((mean == nullptr) == (rstd == nullptr)), "vpt_layernorm_fwd: bad arguments");
#define VPT_CALL(CH) VPT_CUDA_OK(launch_pdl(layernorm_fwd_kernel<CH>, dim3(blocks_for(rows, kEwThreads / 32, 1L << 30)), dim3(kEwThreads), 0, S(stream), BF(x), BF(w), BF(b), BFM(y), mean, rstd, static_cast<long>(rows), D, eps))
  VPT_LN_DISPATCH(D, VPT_CALL);
#undef VPT_CALL
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int vpt_layernorm_bwd(const void* dy, const void* x, const void* w, const float* mean, const float* rstd, void* dx,
                                 float* dw, float* db, int64_t rows, int32_t D, vpt_stream_t stream) {
  VPT_REQUIRE(dy && x && mean && rstd && dx && rows > 0 && D % 8 == 0 && D <= 4096, "vpt_layernorm_bwd: bad arguments");
  static const int wide_min = getenv("VPT_LN_BWD_WIDE_MIN_D") ? atoi(getenv("VPT_LN_BWD_WIDE_MIN_D")) : 2049;
  if (D >= wide_min && (D > 2048 || (dw == nullptr) == (db == nullptr))) {   // dw / db accumulate like dscale / dshift with one "sample" (b = 0)
    VPT_REQUIRE((dw == nullptr) == (db == nullptr), "vpt_layernorm_bwd: rows wider than 2048 take dw and db together");
    VPT_CUDA_OK(launch_pdl(ln_modulate_bwd_wide_kernel, dim3(blocks_for(rows, kEwThreads / 32, 1L << 30)), dim3(kEwThreads), 0, S(stream), BF(dy), BF(x),
                           BF(w), mean, rstd, BFM(dx), dw, db, static_cast<long>(rows), 1, D, 1));
    return 0;
  }
#define VPT_CALL(CH) layernorm_bwd_kernel<CH><<<blocks_for(rows, kEwThreads / 32, 1L << 30), kEwThreads, 0, S(stream)>>>(BF(dy), BF(x), BF(w), mean, rstd, BFM(dx), dw, db, rows, D)
  VPT_LN_DISPATCH(D, VPT_CALL);
#undef VPT_CALL
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int vpt_gated_act_fwd(const void* h, const void* gate, void* a, int64_t rows, int32_t F, int64_t ldh, int64_t ldg,
                                 int64_t lda, int32_t kind, vpt_stream_t stream) {
  const int64_t f8 = (F + 7) / 8 * 8;
  VPT_REQUIRE(h && gate && a && rows > 0 && F > 0 && kind >= 0 && kind <= 2 && (ldh | ldg | lda) % 8 == 0 && ldh >= f8 && ldg >= f8 && lda >= f8,
              "vpt_gated_act_fwd: bad arguments (row pitches must be multiples of 8 and cover F rounded up to 8)");
  VPT_REQUIRE(((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(gate) | reinterpret_cast<uintptr_t>(a)) & 15) == 0, "vpt_gated_act_fwd: 16-byte alignment");
  gated_act_fwd_kernel<<<blocks_for(rows * (f8 / 8), 256, 148 * 32), 256, 0, S(stream)>>>(BF(h), BF(gate), BFM(a), rows, F, ldh, ldg, lda, kind);
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int vpt_gated_act_bwd(const void* da, const void* h, const void* gate, void* dh, void* dgate, int64_t rows, int32_t F,
                                 int64_t ldda, int64_t ldh, int64_t ldg, int64_t lddh, int64_t lddg, int32_t kind, vpt_stream_t stream) {
  const int64_t f8 = (F + 7) / 8 * 8;
  VPT_REQUIRE(da && h && gate && dh && dgate && rows > 0 && F > 0 && kind >= 0 && kind <= 2 && (ldda | ldh | ldg | lddh | lddg) % 8 == 0 &&
                  ldda >= f8 && ldh >= f8 && ldg >= f8 && lddh >= f8 && lddg >= f8, "vpt_gated_act_bwd: bad arguments");
  VPT_REQUIRE(((reinterpret_cast<uintptr_t>(da) | reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(gate) | reinterpret_cast<uintptr_t>(dh) |
                reinterpret_cast<uintptr_t>(dgate)) & 15) == 0, "vpt_gated_act_bwd: 16-byte alignment");
  gated_act_bwd_kernel<<<blocks_for(rows * (f8 / 8), 256, 148 * 32), 256, 0, S(stream)>>>(BF(da), BF(h), BF(gate), BFM(dh), BFM(dgate), rows, F, ldda, ldh, ldg, lddh, lddg, kind);
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int vpt_act_fwd(const void* x, void* y, int64_t n, int32_t kind, vpt_stream_t stream) {
  VPT_REQUIRE(x && y && n > 0 && n % 8 == 0 && kind >= 0 && kind <= 2, "vpt_act_fwd: bad arguments (n must be a multiple of 8)");
  act_fwd_kernel<<<blocks_for(n / 8, 256, 148 * 32), 256, 0, S(stream)>>>(BF(x), BFM(y), n / 8, kind);
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int vpt_act_bwd(const void* dy, const void* x, void* dx, int64_t n, int32_t kind, vpt_stream_t stream) {
  VPT_REQUIRE(dy && x && dx && n > 0 && n % 8 == 0 && kind >= 0 && kind <= 2, "vpt_act_bwd: bad arguments");
  act_bwd_kernel<<<blocks_for(n / 8, 256, 148 * 32), 256, 0, S(stream)>>>(BF(dy), BF(x), BFM(dx), n / 8, kind);
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int vpt_rope_half(const void* x, const float* cosv, const float* sinv, void* y, int64_t tokens, int32_t L, int32_t H,
                             int32_t head_dim, int32_t l0, int64_t ldx, int64_t ldy, int32_t inverse, vpt_stream_t stream) {
  VPT_REQUIRE(x && cosv && sinv && y && tokens > 0 && L > 0 && H > 0 && head_dim % 16 == 0 && l0 >= 0 && ldx % 8 == 0 && ldy % 8 == 0,
              "vpt_rope_half: bad arguments (head_dim must be a multiple of 16)");
  rope_half_kernel<<<blocks_for(tokens * H * (head_dim / 16), 256, 148 * 32), 256, 0, S(stream)>>>(BF(x), cosv, sinv, BFM(y), tokens, L, H, head_dim, l0, ldx, ldy, inverse);
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int vpt_pope_fwd(const void* x, const float* cos_sin, const float* bias, void* y, int64_t tokens, int32_t L, int32_t H,
                            int32_t d, int64_t ldx, int64_t ldy, vpt_stream_t stream) {
  VPT_REQUIRE(x && cos_sin && y && tokens > 0 && L > 0 && H > 0 && d % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0, "vpt_pope_fwd: bad arguments");
  pope_fwd_kernel<<<blocks_for(tokens * H * (d / 8), 256, 148 * 32), 256, 0, S(stream)>>>(BF(x), cos_sin, bias, BFM(y), tokens, L, H, d, ldx, ldy);
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int vpt_pope_bwd(const void* dy, const void* x, const float* cos_sin, const float* bias, void* dx, int64_t tokens,
                            int32_t L, int32_t H, int32_t d, int64_t lddy, int64_t ldx, int64_t lddx, vpt_stream_t stream) {
  VPT_REQUIRE(dy && x && cos_sin && dx && tokens > 0 && L > 0 && H > 0 && d % 8 == 0 && (lddy | ldx | lddx) % 8 == 0, "vpt_pope_bwd: bad arguments");
  pope_bwd_kernel<<<blocks_for(tokens * H * (d / 8), 256, 148 * 32), 256, 0, S(stream)>>>(BF(dy), BF(x), cos_sin, bias, BFM(dx), tokens, L, H, d, lddy, ldx, lddx);
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int vpt_token_gather(const void* src, const int64_t* idx, void* dst, int32_t B, int64_t L_full, int64_t n, int32_t D,
                                int32_t scatter, vpt_stream_t stream) {
  VPT_REQUIRE(src && idx && dst && B > 0 && L_full > 0 && n > 0 && n <= L_full && D % 8 == 0, "vpt_token_gather: bad arguments");
  VPT_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0, "vpt_token_gather: 16-byte alignment");
  token_gather_kernel<<<blocks_for(static_cast<long>(B) * n * (D / 8), 256, 148 * 32), 256, 0, S(stream)>>>(
      BF(src), reinterpret_cast<const long*>(idx), BFM(dst), B, static_cast<long>(L_full), static_cast<long>(n), D, scatter);
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}

static int patch_common(const void* src, void* dst, int B, int C, int H, int W, int p, int order, bool to_patches, cudaStream_t s) {
  VPT_REQUIRE(src && dst && B > 0 && C > 0 && p > 0 && H % p == 0 && W % p == 0 && (order == 0 || order == 1), "patchify: bad arguments");
  const long total = static_cast<long>(B) * C * H * W;
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  if (p % 8 == 0 && C >= 1 && C <= 4 && aligned) {
    const unsigned g = blocks_for(static_cast<long>(B) * H * (W / 8), 256, 148 * 32);
    const uint16_t* s_ = static_cast<const uint16_t*>(src);
    uint16_t* d_ = static_cast<uint16_t*>(dst);
#define VPT_PATCH(TP, CC) patch_permute_vec_kernel<TP, CC><<<g, 256, 0, s>>>(s_, d_, B, H, W, p, order)
    if (to_patches) {
      if (C == 1) VPT_PATCH(true, 1); else if (C == 2) VPT_PATCH(true, 2); else if (C == 3) VPT_PATCH(true, 3); else VPT_PATCH(true, 4);
    } else {
      if (C == 1) VPT_PATCH(false, 1); else if (C == 2) VPT_PATCH(false, 2); else if (C == 3) VPT_PATCH(false, 3); else VPT_PATCH(false, 4);
    }
#undef VPT_PATCH
    VPT_CUDA_OK(cudaGetLastError());
    return 0;
  }
  const unsigned grid = blocks_for(total, 256, 148 * 64);
  if (to_patches)
    patch_permute_kernel<true><<<grid, 256, 0, s>>>(static_cast<const uint16_t*>(src), static_cast<uint16_t*>(dst), B, C, H, W, p, order);
  else
    patch_permute_kernel<false><<<grid, 256, 0, s>>>(static_cast<const uint16_t*>(src), static_cast<uint16_t*>(dst), B, C, H, W, p, order);
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int vpt_patchify(const void* img, void* patches, int32_t B, int32_t C, int32_t H, int32_t W, int32_t p, int32_t order,
                            vpt_stream_t stream) {
  return patch_common(img, patches, B, C, H, W, p, order, true, S(stream));
}
extern "C" int vpt_unpatchify(const void* patches, void* img, int32_t B, int32_t C, int32_t H, int32_t W, int32_t p,
                              int32_t order, vpt_stream_t stream) {
  return patch_common(patches, img, B, C, H, W, p, order, false, S(stream));
}

extern "C" int vpt_copy_rows(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t rows, int64_t row_bytes,
                             vpt_stream_t stream) {
  VPT_REQUIRE(dst && rows > 0 && row_bytes > 0, "vpt_copy_rows: bad arguments");
  VPT_REQUIRE(((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | static_cast<uintptr_t>(dst_pitch) |
                static_cast<uintptr_t>(src_pitch) | static_cast<uintptr_t>(row_bytes)) & 15) == 0,
              "vpt_copy_rows: pointers, pitches and the row length must be multiples of 16 bytes");
  const long vec = row_bytes / 16;
  VPT_CUDA_OK(launch_pdl(copy_rows_kernel, dim3(blocks_for(rows * vec, 256, 148 * 8)), dim3(256), 0, S(stream),
                         static_cast<uint8_t*>(dst), static_cast<long>(dst_pitch), static_cast<const uint8_t*>(src),
                         static_cast<long>(src_pitch), static_cast<long>(rows), vec));
  return 0;
}

// ------------------------------------------------------------------------------------------------ optimiser / loss
extern "C" int vpt_grad_sumsq(const float* g, int64_t n, float scale, float* out, vpt_stream_t stream) {
  VPT_REQUIRE(g && out && n > 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0, "vpt_grad_sumsq: bad arguments");
  grad_sumsq_kernel<<<blocks_for(n / 4 + 1, 256, 148 * 8), 256, 0, S(stream)>>>(g, n, scale, out);
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int vpt_adamw_step(void* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                              float beta2, float eps, float weight_decay, float grad_scale, const float* sumsq,
                              float max_norm, const float* step, int32_t zero_grad, vpt_stream_t stream) {
  VPT_REQUIRE(param && grad && exp_avg && exp_avg_sq && step && n > 0, "vpt_adamw_step: bad arguments");
  AdamWParams a{BFM(param), grad, exp_avg, exp_avg_sq, static_cast<long>(n), lr, beta1, beta2, eps, weight_decay,
                grad_scale, sumsq, max_norm, step, zero_grad};
  adamw_kernel<<<blocks_for(n, 256, 148 * 8), 256, 0, S(stream)>>>(a);
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int vpt_radam_schedulefree_step(void* param, float* grad, float* z, float* exp_avg_sq, int64_t n, double lr,
                                           double beta1, double beta2, float eps, float weight_decay, double r,
                                           double weight_lr_power, int32_t silent_sgd_phase, float grad_scale,
                                           const float* sumsq, float max_norm, double* sched, float* coef,
                                           int32_t zero_grad, vpt_stream_t stream) {
  VPT_REQUIRE(param && grad && z && exp_avg_sq && sched && coef && n > 0, "vpt_radam_schedulefree_step: bad arguments");
  radam_sf_advance_kernel<<<1, 32, 0, S(stream)>>>(sched, coef, lr, beta1, beta2, r, weight_lr_power, silent_sgd_phase);
  VPT_CUDA_OK(cudaGetLastError());
  RAdamSFParams a{BFM(param), grad, z, exp_avg_sq, static_cast<long>(n), static_cast<float>(beta2), eps, weight_decay, grad_scale, sumsq, max_norm,
                  coef, zero_grad};
  radam_sf_kernel<<<blocks_for(n, 256, 148 * 8), 256, 0, S(stream)>>>(a);
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int vpt_radam_schedulefree_swap(void* param, const float* z, int64_t n, float beta1, int32_t to_eval,
                                           vpt_stream_t stream) {
  VPT_REQUIRE(param && z && n > 0 && beta1 > 0.f, "vpt_radam_schedulefree_swap: bad arguments");
  const float w = to_eval ? 1.f - 1.f / beta1 : 1.f - beta1;
  radam_sf_swap_kernel<<<blocks_for(n, 256, 148 * 8), 256, 0, S(stream)>>>(BFM(param), z, static_cast<long>(n), w);
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int vpt_noise_mix(const void* latents, const void* randn, int dtype, const float* timestep, int64_t batch,
                             int64_t per_sample, float noise_scale, int32_t clean_at_zero, void* noisy, void* noisy_bf16,
                             vpt_stream_t stream) {
  VPT_REQUIRE(latents && randn && timestep && noisy && batch > 0 && per_sample > 0 && dtype >= 0 && dtype <= 2,
              "vpt_noise_mix: bad arguments");
  const long total = static_cast<long>(batch) * per_sample;
  noise_mix_kernel<<<blocks_for(total / 8 + 1, 256, 148 * 8), 256, 0, S(stream)>>>(latents, randn, dtype, timestep, per_sample, total,
                                                                                   noise_scale, clean_at_zero, noisy, BFM(noisy_bf16));
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int vpt_scale_by_scalar(const void* x, const float* scalar, void* y, int64_t n, vpt_stream_t stream) {
  VPT_REQUIRE(x && scalar && y && n > 0, "vpt_scale_by_scalar: bad arguments");
  scale_by_scalar_kernel<<<blocks_for(n / 8 + 1, 256, 148 * 8), 256, 0, S(stream)>>>(BF(x), scalar, BFM(y), static_cast<long>(n));
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int vpt_flow_loss(const void* pred, const void* clean, const void* noisy, int in_dtype, const float* timestep,
                             int64_t batch, int64_t per_sample, int32_t mode, float clamp_eps, float* loss_out, void* dpred,
                             vpt_stream_t stream) {
  VPT_REQUIRE(pred && clean && loss_out && batch > 0 && per_sample > 0 && (mode == 0 || (mode == 1 && noisy && timestep)) &&
                  in_dtype >= 0 && in_dtype <= 2, "vpt_flow_loss: bad arguments");
  const long total = static_cast<long>(batch) * per_sample;
  flow_loss_kernel<<<blocks_for(total / 8 + 1, 256, 148 * 8), 256, 0, S(stream)>>>(BF(pred), clean, noisy, in_dtype, timestep, per_sample, total, mode,
                                                                       clamp_eps, loss_out, BFM(dpred));
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}
