// Host launchers for the attention kernels.
#pragma once
#include "attention.cuh"
#include "attention_bwd.cuh"
#include "attention_simple.cuh"
#include "host.cuh"

namespace vpt {

struct AttnTensor {
  const void* ptr;
  long sb, sl, sh;   // element strides of (batch, token, head); head_dim is contiguous
};

inline int make_attn_tmap(CUtensorMap* m, const AttnTensor& t, int B, int H, int L, int box_rows = kAttnTile, int hd = kAttnHD) {
  // head_dim 80: the 64-column box at column 64 is zero-filled past column 80
  const uint64_t dims[4] = {static_cast<uint64_t>(hd), static_cast<uint64_t>(L), static_cast<uint64_t>(H),
                            static_cast<uint64_t>(B)};
  const uint64_t strides[3] = {static_cast<uint64_t>(t.sl) * 2, static_cast<uint64_t>(t.sh) * 2,
                               static_cast<uint64_t>(t.sb) * 2};
  const uint32_t box[4] = {kAttnHD, static_cast<uint32_t>(box_rows), 1, 1};
  return make_tmap_bf16_4d(m, t.ptr, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int HD>
inline int launch_attn_fwd_t(const AttnTensor& q, const AttnTensor& k, const AttnTensor& v, const AttnTensor& o, int B,
                             int H, int Lq, int Lk, const int* seqlens_k, float scale, float* lse2, cudaStream_t stream) {
  using Smem = AttnFwdSmemT<HD>;
  if ((o.sl | o.sh | o.sb) & 7) return fail("attention output strides must be multiples of 8 elements");
  CUtensorMap tq, tk, tv, to;
  if (make_attn_tmap(&tq, q, B, H, Lq, kAttnTile, HD) || make_attn_tmap(&tk, k, B, H, Lk, kAttnTile, HD) ||
      make_attn_tmap(&tv, v, B, H, Lk, kAttnTile, HD) || make_attn_tmap(&to, o, B, H, Lq, 32, HD))
    return 1;
  AttnFwdParams p{};
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk;
  p.seqlens_k = seqlens_k;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.o = static_cast<__nv_bfloat16*>(const_cast<void*>(o.ptr));
  p.o_sb = o.sb; p.o_sl = o.sl; p.o_sh = o.sh;
  p.lse2 = lse2;
  static bool attr = false;
  if (!attr) {
    VPT_CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem::kTotal));
    attr = true;
  }
  const int items = ((Lq + kAttnTile - 1) / kAttnTile) * H * B;
  const int ctas = (items + Smem::kStreams - 1) / Smem::kStreams;   // two item streams per CTA
  const int slots = sm_count();
  VPT_CUDA_OK(launch_pdl(attn_fwd_kernel<HD>, dim3(ctas < slots ? ctas : slots), dim3(640), Smem::kTotal, stream, tq, tk, tv, to, p));
  return 0;
}
inline int launch_attn_fwd(const AttnTensor& q, const AttnTensor& k, const AttnTensor& v, const AttnTensor& o, int B,
                           int H, int Lq, int Lk, const int* seqlens_k, float scale, float* lse2,
                           cudaStream_t stream, int head_dim = 64) {
  if (head_dim == 80) return launch_attn_fwd_t<80>(q, k, v, o, B, H, Lq, Lk, seqlens_k, scale, lse2, stream);
  return launch_attn_fwd_t<64>(q, k, v, o, B, H, Lq, Lk, seqlens_k, scale, lse2, stream);
}

// dq_f32 must be zero on entry (fp32, same (b, l, h) element strides as given); delta is a [B,H,Lq] fp32 workspace.
template <int HD>
inline int launch_attn_bwd_t(const AttnTensor& q, const AttnTensor& k, const AttnTensor& v, const AttnTensor& o,
                             const AttnTensor& d_o, const AttnTensor& dq_f32, const AttnTensor& dk, const AttnTensor& dv,
                             int B, int H, int Lq, int Lk, const int* seqlens_k, float scale, const float* lse2,
                             float* delta, cudaStream_t stream) {
  using Smem = AttnBwd2SmemT<HD>;
  CUtensorMap tq, tk, tv, tdo;
  // Q / dO travel as 64-query sub-tiles (a ring of sub-tile slots), K / V as 128-key tiles
  if (make_attn_tmap(&tq, q, B, H, Lq, 64, HD) || make_attn_tmap(&tk, k, B, H, Lk, kAttnTile, HD) ||
      make_attn_tmap(&tv, v, B, H, Lk, kAttnTile, HD) || make_attn_tmap(&tdo, d_o, B, H, Lq, 64, HD))
    return 1;
  AttnDeltaParams dp{};
  dp.B = B; dp.H = H; dp.Lq = Lq;
  dp.o = static_cast<const __nv_bfloat16*>(o.ptr);
  dp.d_o = static_cast<const __nv_bfloat16*>(d_o.ptr);
  dp.o_sb = o.sb; dp.o_sl = o.sl; dp.o_sh = o.sh;
  dp.do_sb = d_o.sb; dp.do_sl = d_o.sl; dp.do_sh = d_o.sh;
  dp.delta = delta;
  dp.Lq_pad = (Lq + kAttnTile - 1) / kAttnTile * kAttnTile;
  dp.scale = scale;
  const long groups = static_cast<long>(B) * H * dp.Lq_pad;
  if (HD == 64) {
    if (B > 65535) return 1;   // batch entries ride on gridDim.y
    VPT_CUDA_OK(launch_pdl(attn_bwd_delta_kernel, dim3(static_cast<unsigned>(dp.Lq_pad / kDeltaRows), static_cast<unsigned>(B)), dim3(256), 0, stream, dp));
  } else {
    VPT_CUDA_OK(launch_pdl(attn_bwd_delta_row_kernel<HD>, dim3(static_cast<unsigned>((groups + 255) / 256)), dim3(256), 0, stream, dp));
  }

  CUtensorMap tdq;
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(HD), static_cast<uint64_t>(Lq), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
    const uint64_t strides[3] = {static_cast<uint64_t>(dq_f32.sl) * 4, static_cast<uint64_t>(dq_f32.sh) * 4, static_cast<uint64_t>(dq_f32.sb) * 4};
    const uint32_t box[4] = {32, 32, 1, 1};
    if (make_tmap_f32_4d(&tdq, dq_f32.ptr, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  AttnBwd2Params p{};
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk;
  p.seqlens_k = seqlens_k;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse2 = lse2;
  p.delta = delta;
  p.dq = static_cast<float*>(const_cast<void*>(dq_f32.ptr));
  p.dq_sb = dq_f32.sb; p.dq_sl = dq_f32.sl; p.dq_sh = dq_f32.sh;
  CUtensorMap tdk, tdv;
  if (make_attn_tmap(&tdk, dk, B, H, Lk, 32, HD) || make_attn_tmap(&tdv, dv, B, H, Lk, 32, HD)) return 1;
  p.nk = (Lk + kAttnTile - 1) / kAttnTile;
  p.nq = (Lq + kAttnTile - 1) / kAttnTile;
  p.num_items = B * H * p.nk;
  static bool attr = false;
  if (!attr) {
    VPT_CUDA_OK(cudaFuncSetAttribute(attn_bwd2_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem::kTotal));
    attr = true;
  }
  const int ctas = p.num_items < sm_count() ? p.num_items : sm_count();
  VPT_CUDA_OK(launch_pdl(attn_bwd2_kernel<HD>, dim3(ctas), dim3(512), Smem::kTotal, stream, tq, tk, tv, tdo, tdq, tdk, tdv, p));
  return 0;
}
inline int launch_attn_bwd(const AttnTensor& q, const AttnTensor& k, const AttnTensor& v, const AttnTensor& o,
                           const AttnTensor& d_o, const AttnTensor& dq_f32, const AttnTensor& dk, const AttnTensor& dv,
                           int B, int H, int Lq, int Lk, const int* seqlens_k, float scale, const float* lse2,
                           float* delta, cudaStream_t stream, int head_dim = 64) {
  if (head_dim == 80)
    return launch_attn_bwd_t<80>(q, k, v, o, d_o, dq_f32, dk, dv, B, H, Lq, Lk, seqlens_k, scale, lse2, delta, stream);
  return launch_attn_bwd_t<64>(q, k, v, o, d_o, dq_f32, dk, dv, B, H, Lq, Lk, seqlens_k, scale, lse2, delta, stream);
}

// ---- head_dim 32 / 96 / 128: CUDA-core kernels (attention_simple.cuh)
inline SimpleAttnTensor SAT(const AttnTensor& t) { return SimpleAttnTensor{static_cast<const __nv_bfloat16*>(t.ptr), t.sb, t.sl, t.sh}; }

template <int HD>
int launch_attn_simple_fwd_t(const SimpleAttnParams& p, cudaStream_t stream) {
  dim3 grid((p.Lq + kSaRows - 1) / kSaRows, p.H, p.B);
  attn_simple_fwd_kernel<HD><<<grid, kSaRows, 0, stream>>>(p);
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}
template <int HD>
int launch_attn_simple_bwd_t(const SimpleAttnParams& p, float* delta, cudaStream_t stream) {
  const long total = static_cast<long>(p.B) * p.H * p.Lq_pad;
  attn_simple_delta_kernel<HD><<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(p, delta);
  dim3 gq((p.Lq + kSaRows - 1) / kSaRows, p.H, p.B), gk((p.Lk + kSaRows - 1) / kSaRows, p.H, p.B);
  attn_simple_bwd_dq_kernel<HD><<<gq, kSaRows, 0, stream>>>(p);
  attn_simple_bwd_dkv_kernel<HD, true><<<gk, kSaRows, 0, stream>>>(p);
  attn_simple_bwd_dkv_kernel<HD, false><<<gk, kSaRows, 0, stream>>>(p);
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}
#define VPT_SIMPLE_HD(CALL)                                                                   \
  switch (head_dim) {                                                                         \
    case 32: return CALL(32);                                                                 \
    case 96: return CALL(96);                                                                 \
    case 128: return CALL(128);                                                               \
    default: return fail("attention: head_dim must be 64 / 80 (tcgen05 kernels) or 32 / 96 / 128 (CUDA-core kernels)"); \
  }

inline int launch_attn_simple_fwd(const AttnTensor& q, const AttnTensor& k, const AttnTensor& v, const AttnTensor& o, int B,
                                  int H, int Lq, int Lk, int head_dim, const int* seqlens_k, float scale, float* lse2,
                                  cudaStream_t stream) {
  SimpleAttnParams p{};
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk; p.Lq_pad = (Lq + 127) / 128 * 128;
  p.seqlens_k = seqlens_k; p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  p.q = SAT(q); p.k = SAT(k); p.v = SAT(v); p.o = SAT(o);
  p.lse2 = lse2;
#define VPT_CALL(HD) launch_attn_simple_fwd_t<HD>(p, stream)
  VPT_SIMPLE_HD(VPT_CALL)
#undef VPT_CALL
}
inline int launch_attn_simple_bwd(const AttnTensor& q, const AttnTensor& k, const AttnTensor& v, const AttnTensor& o,
                                  const AttnTensor& d_o, const AttnTensor& dq_f32, const AttnTensor& dk, const AttnTensor& dv,
                                  int B, int H, int Lq, int Lk, int head_dim, const int* seqlens_k, float scale,
                                  const float* lse2, float* delta, cudaStream_t stream) {
  SimpleAttnParams p{};
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk; p.Lq_pad = (Lq + 127) / 128 * 128;
  p.seqlens_k = seqlens_k; p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  p.q = SAT(q); p.k = SAT(k); p.v = SAT(v); p.o = SAT(o); p.d_o = SAT(d_o);
  p.lse2 = const_cast<float*>(lse2); p.delta = delta;
  p.dq = static_cast<float*>(const_cast<void*>(dq_f32.ptr)); p.dq_sb = dq_f32.sb; p.dq_sl = dq_f32.sl; p.dq_sh = dq_f32.sh;
  p.dk = static_cast<__nv_bfloat16*>(const_cast<void*>(dk.ptr)); p.dk_sb = dk.sb; p.dk_sl = dk.sl; p.dk_sh = dk.sh;
  p.dv = static_cast<__nv_bfloat16*>(const_cast<void*>(dv.ptr)); p.dv_sb = dv.sb; p.dv_sl = dv.sl; p.dv_sh = dv.sh;
#define VPT_CALL(HD) launch_attn_simple_bwd_t<HD>(p, delta, stream)
  VPT_SIMPLE_HD(VPT_CALL)
#undef VPT_CALL
}
#undef VPT_SIMPLE_HD

}  // namespace vpt
