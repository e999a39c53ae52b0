// HBM-bound kernels of the block families either side of the JiT block: affine LayerNorm,
// GeGLU / plain activations, CogView4's half-split rotary embedding on the image tokens, PoPE, and the token
// gather / scatter of TREAD routing.  Same conventions as elementwise.cuh: 16-byte accesses, fp32 arithmetic, the
// reference's bf16 rounding points.
//
// Reference semantics (file:line under /root/reference):
//   nn.LayerNorm / FP32LayerNorm (affine)    src/models/sdxl/denoiser.py:248-250; src/modules/norm.py:9-17
//   GeGLU                                    src/models/sdxl/denoiser.py:175-186        h * gelu(gate), erf form
//   FeedForward activation (gelu tanh)       src/models/cogview4/denoiser.py:312-343
//   apply_rotary_emb                         src/models/cogview4/denoiser.py:203-218
//   apply_pope                               src/models/jit/extension/pope.py:6-38
//   keep_and_route_tokens / re-insertion     train/jit/class_to_image_tread.py:73-118
#pragma once
#include "elementwise.cuh"

namespace vpt {

// ------------------------------------------------------------------------------------------- affine LayerNorm
// y = bf16( (x - mean) * rstd * w + b )  (one rounding, like torch's LayerNorm on a bf16 tensor); w, b [D] or nullptr
template <int kCh>
__global__ void __launch_bounds__(kEwThreads)
layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w, const __nv_bfloat16* __restrict__ b,
                     __nv_bfloat16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, long rows, int D,
                     float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * (kEwThreads / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nch = D >> 3;
  uint4 v[kCh];
#pragma unroll
  for (int i = 0; i < kCh; ++i) {
    const int c = lane + i * 32;
    v[i] = make_uint4(0, 0, 0, 0);
    if (c < nch) v[i] = ld_stream(x + row * D + c * 8);
  }
  // one reduction round for mean and variance: shifted moments (see ln_modulate_fwd_kernel)
  const float x0 = __shfl_sync(0xffffffffu, ew_lo(v[0].x), 0);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < kCh; ++i) {
    const int c = lane + i * 32;
    if (c < nch) {
      float f[8];
      ew_unpack8(v[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = f[e] - x0;
        s1 += d;
        s2 += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  const float m1 = s1 / static_cast<float>(D);
  const float mean = x0 + m1;
  const float rstd = rsqrtf(fmaxf(s2 / static_cast<float>(D) - m1 * m1, 0.f) + eps);
  if (lane == 0 && mean_out != nullptr) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
#pragma unroll
  for (int i = 0; i < kCh; ++i) {
    const int c = lane + i * 32;
    if (c < nch) {
      float f[8], fw[8], fb[8], o[8];
      ew_unpack8(v[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) { fw[e] = 1.f; fb[e] = 0.f; }
      if (w != nullptr) ew_unpack8(*reinterpret_cast<const uint4*>(w + c * 8), fw);
      if (b != nullptr) ew_unpack8(*reinterpret_cast<const uint4*>(b + c * 8), fb);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (f[e] - mean) * rstd * fw[e] + fb[e];
      st_stream(y + row * D + c * 8, ew_pack8(o));
    }
  }
}
// dx = rstd * (dn - mean(dn) - n * mean(dn * n)),  dn = dy * w;  optional dw[D] += sum dy * n, db[D] += sum dy (fp32 atomics)
template <int kCh>
__global__ void __launch_bounds__(kEwThreads)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                     const float* __restrict__ mean_in, const float* __restrict__ rstd_in, __nv_bfloat16* __restrict__ dx,
                     float* __restrict__ dw, float* __restrict__ db, long rows, int D) {
  const int lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * (kEwThreads / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nch = D >> 3;
  uint4 vx[kCh], vg[kCh];
#pragma unroll
  for (int i = 0; i < kCh; ++i) {                  // every load of the row is in flight before the first use
    const int c = lane + i * 32;
    vx[i] = vg[i] = make_uint4(0, 0, 0, 0);
    if (c < nch) {
      vx[i] = ld_stream(x + row * D + c * 8);
      vg[i] = ld_stream(dy + row * D + c * 8);
    }
  }
  const float mean = mean_in[row], rstd = rstd_in[row];
  float sa = 0.f, sb = 0.f;
#pragma unroll
  for (int i = 0; i < kCh; ++i) {
    const int c = lane + i * 32;
    if (c < nch) {
      float fx[8], fg[8], fw[8];
      ew_unpack8(vx[i], fx);
      ew_unpack8(vg[i], fg);
#pragma unroll
      for (int e = 0; e < 8; ++e) fw[e] = 1.f;
      if (w != nullptr) ew_unpack8(*reinterpret_cast<const uint4*>(w + c * 8), fw);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float n = (fx[e] - mean) * rstd;
        const float dn = fg[e] * fw[e];
        sa += dn;
        sb += dn * n;
        if (dw != nullptr) atomicAdd(dw + c * 8 + e, fg[e] * n);
        if (db != nullptr) atomicAdd(db + c * 8 + e, fg[e]);
      }
    }
  }
  sa = warp_sum(sa) / static_cast<float>(D);
  sb = warp_sum(sb) / static_cast<float>(D);
#pragma unroll
  for (int i = 0; i < kCh; ++i) {
    const int c = lane + i * 32;
    if (c < nch) {
      float fx[8], fg[8], fw[8], o[8];
      ew_unpack8(vx[i], fx);
      ew_unpack8(vg[i], fg);
#pragma unroll
      for (int e = 0; e < 8; ++e) fw[e] = 1.f;
      if (w != nullptr) ew_unpack8(*reinterpret_cast<const uint4*>(w + c * 8), fw);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float n = (fx[e] - mean) * rstd;
        o[e] = rstd * (fg[e] * fw[e] - sa - n * sb);
      }
      st_stream(dx + row * D + c * 8, ew_pack8(o));
    }
  }
}

// ------------------------------------------------------------------------------------------- activations
// kind: 0 = SiLU, 1 = GELU (erf form, F.gelu default), 2 = GELU (tanh form, "gelu_pytorch_tanh")
__device__ __forceinline__ float act_value(float x, int kind) {
  if (kind == 0) return x / (1.f + __expf(-x));
  if (kind == 1) return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.f + tanhf(u));
}
__device__ __forceinline__ float act_deriv(float x, int kind) {
  if (kind == 0) {
    const float s = 1.f / (1.f + __expf(-x));
    return s * (1.f + x * (1.f - s));
  }
  if (kind == 1) return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  const float t = tanhf(u);
  return 0.5f * (1.f + t) + 0.5f * x * (1.f - t * t) * 0.7978845608028654f * (1.f + 3.f * 0.044715f * x * x);
}
// a = bf16( bf16(act(gate)) * h ): GeGLU (kind 1: h, gate = halves of one [rows, 2F] projection) or SwiGLU (kind 0)
__global__ void __launch_bounds__(256)
gated_act_fwd_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ gate, __nv_bfloat16* __restrict__ a,
                     long rows, int F, long ldh, long ldg, long lda, int kind) {
  const int nch = (F + 7) >> 3;
  const long total = rows * nch;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / nch;
    const int c = static_cast<int>(i % nch);
    float fh[8], fg[8], o[8];
    ew_unpack8(ld_stream(h + r * ldh + c * 8), fh);
    ew_unpack8(ld_stream(gate + r * ldg + c * 8), fg);
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      float aa = act_value(fg[e], kind), ab = act_value(fg[e + 1], kind);
      ew_round2(aa, ab);
      o[e] = aa * fh[e];
      o[e + 1] = ab * fh[e + 1];
    }
    st_stream(a + r * lda + c * 8, ew_pack8(o));
  }
}
// dh = da * act(gate),  dgate = da * h * act'(gate)
__global__ void __launch_bounds__(256)
gated_act_bwd_kernel(const __nv_bfloat16* __restrict__ da, const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ gate,
                     __nv_bfloat16* __restrict__ dh, __nv_bfloat16* __restrict__ dgate, long rows, int F, long ldda, long ldh,
                     long ldg, long lddh, long lddg, int kind) {
  const int nch = (F + 7) >> 3;
  const long total = rows * nch;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / nch;
    const int c = static_cast<int>(i % nch);
    float fa[8], fh[8], fg[8], oh[8], og[8];
    ew_unpack8(ld_stream(da + r * ldda + c * 8), fa);
    ew_unpack8(ld_stream(h + r * ldh + c * 8), fh);
    ew_unpack8(ld_stream(gate + r * ldg + c * 8), fg);
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      float aa = act_value(fg[e], kind), ab = act_value(fg[e + 1], kind);
      ew_round2(aa, ab);
      oh[e] = fa[e] * aa;
      oh[e + 1] = fa[e + 1] * ab;
      og[e] = fa[e] * fh[e] * act_deriv(fg[e], kind);
      og[e + 1] = fa[e + 1] * fh[e + 1] * act_deriv(fg[e + 1], kind);
    }
    st_stream(dh + r * lddh + c * 8, ew_pack8(oh));
    st_stream(dgate + r * lddg + c * 8, ew_pack8(og));
  }
}
// y = bf16(act(x));  backward: dx = dy * act'(x)
__global__ void __launch_bounds__(256)
act_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long n8, int kind) {
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += static_cast<long>(gridDim.x) * blockDim.x) {
    float f[8];
    ew_unpack8(ld_stream(x + i * 8), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = act_value(f[e], kind);
    st_stream(y + i * 8, ew_pack8(f));
  }
}
__global__ void __launch_bounds__(256)
act_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ dx, long n8,
               int kind) {
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += static_cast<long>(gridDim.x) * blockDim.x) {
    float f[8], g[8];
    ew_unpack8(ld_stream(x + i * 8), f);
    ew_unpack8(ld_stream(dy + i * 8), g);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = g[e] * act_deriv(f[e], kind);
    st_stream(dx + i * 8, ew_pack8(f));
  }
}

// ------------------------------------------------------------------------------------------- half-split rotary (CogView4)
// x [B, L, H, hd] token-major; tokens l >= l0 are rotated with table row (l - l0):
//   y[i] = x[i] cos[i] - x[i + hd/2] sin[i],  y[i + hd/2] = x[i + hd/2] cos[i + hd/2] + x[i] sin[i + hd/2]      (i < hd/2)
// cos, sin fp32 [S, hd].  inverse = true applies the transposed rotation (the backward pass).  Tokens l < l0 are copied.
__global__ void __launch_bounds__(256)
rope_half_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ cosv, const float* __restrict__ sinv,
                 __nv_bfloat16* __restrict__ y, long tokens, int L, int H, int hd, int l0, long ldx, long ldy, int inverse) {
  const int half8 = hd >> 4;                         // 16-byte chunks per half head
  const long total = tokens * H * half8;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % half8);
    const long th = i / half8;
    const int hh = static_cast<int>(th % H);
    const long t = th / H;
    const int l = static_cast<int>(t % L);
    const __nv_bfloat16* src = x + t * ldx + hh * hd + c * 8;
    __nv_bfloat16* dst = y + t * ldy + hh * hd + c * 8;
    const uint4 va = ld_stream(src), vb = ld_stream(src + hd / 2);
    if (l < l0) {
      st_stream(dst, va);
      st_stream(dst + hd / 2, vb);
      continue;
    }
    float a[8], b[8], oa[8], ob[8];
    ew_unpack8(va, a);
    ew_unpack8(vb, b);
    const float* cr = cosv + static_cast<long>(l - l0) * hd + c * 8;
    const float* sr = sinv + static_cast<long>(l - l0) * hd + c * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float c1 = cr[e], c2 = cr[e + hd / 2], s1 = sr[e], s2 = sr[e + hd / 2];
      if (!inverse) {
        oa[e] = a[e] * c1 - b[e] * s1;
        ob[e] = b[e] * c2 + a[e] * s2;
      } else {
        oa[e] = a[e] * c1 + b[e] * s2;
        ob[e] = b[e] * c2 - a[e] * s1;
      }
    }
    st_stream(dst, ew_pack8(oa));
    st_stream(dst + hd / 2, ew_pack8(ob));
  }
}

// ------------------------------------------------------------------------------------------- PoPE
// x [B, L, H, d] token-major (any row pitch), table fp32 [L, d, 2] = (cos, sin) of the position angle, bias fp32 [H, d] or
// nullptr (learned phase offset): m = softplus(x), phi = angle + bias;  y [B, L, H, 2d]: y[2i] = m cos(phi), y[2i+1] = m sin(phi)
__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(__expf(x)); }
__global__ void __launch_bounds__(256)
pope_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ cos_sin, const float* __restrict__ bias,
                __nv_bfloat16* __restrict__ y, long tokens, int L, int H, int d, long ldx, long ldy) {
  const int nch = d >> 3;
  const long total = tokens * H * nch;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % nch);
    const long th = i / nch;
    const int hh = static_cast<int>(th % H);
    const long t = th / H;
    const int l = static_cast<int>(t % L);
    float f[8], o0[8], o1[8];
    ew_unpack8(ld_stream(x + t * ldx + hh * d + c * 8), f);
    const float* cs = cos_sin + (static_cast<long>(l) * d + c * 8) * 2;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float cr = cs[2 * e], sr = cs[2 * e + 1];
      if (bias != nullptr) {
        float sb, cb;
        sincosf(bias[hh * d + c * 8 + e], &sb, &cb);
        const float c2 = cr * cb - sr * sb, s2 = sr * cb + cr * sb;
        cr = c2;
        sr = s2;
      }
      const float m = softplus_f(f[e]);
      const int j = 2 * e;                              // interleaved (re, im) pairs
      (j < 8 ? o0[j] : o1[j - 8]) = m * cr;
      (j + 1 < 8 ? o0[j + 1] : o1[j + 1 - 8]) = m * sr;
    }
    __nv_bfloat16* dst = y + t * ldy + hh * 2 * d + c * 16;
    st_stream(dst, ew_pack8(o0));
    st_stream(dst + 8, ew_pack8(o1));
  }
}
// dx = sigmoid(x) * (dy[2i] cos(phi) + dy[2i+1] sin(phi))
__global__ void __launch_bounds__(256)
pope_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x, const float* __restrict__ cos_sin,
                const float* __restrict__ bias, __nv_bfloat16* __restrict__ dx, long tokens, int L, int H, int d, long lddy,
                long ldx, long lddx) {
  const int nch = d >> 3;
  const long total = tokens * H * nch;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % nch);
    const long th = i / nch;
    const int hh = static_cast<int>(th % H);
    const long t = th / H;
    const int l = static_cast<int>(t % L);
    float f[8], g0[8], g1[8], o[8];
    ew_unpack8(ld_stream(x + t * ldx + hh * d + c * 8), f);
    const __nv_bfloat16* src = dy + t * lddy + hh * 2 * d + c * 16;
    ew_unpack8(ld_stream(src), g0);
    ew_unpack8(ld_stream(src + 8), g1);
    const float* cs = cos_sin + (static_cast<long>(l) * d + c * 8) * 2;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float cr = cs[2 * e], sr = cs[2 * e + 1];
      if (bias != nullptr) {
        float sb, cb;
        sincosf(bias[hh * d + c * 8 + e], &sb, &cb);
        const float c2 = cr * cb - sr * sb, s2 = sr * cb + cr * sb;
        cr = c2;
        sr = s2;
      }
      const int j = 2 * e;
      const float gre = j < 8 ? g0[j] : g1[j - 8];
      const float gim = j + 1 < 8 ? g0[j + 1] : g1[j + 1 - 8];
      o[e] = (gre * cr + gim * sr) / (1.f + __expf(-f[e]));
    }
    st_stream(dx + t * lddx + hh * d + c * 8, ew_pack8(o));
  }
}

// ------------------------------------------------------------------------------------------- token gather / scatter (TREAD)
// gather : dst[b, j, :] = src[b, idx[j], :]                    src [B, Ls, D], dst [B, n, D]
// scatter: dst[b, idx[j], :] = src[b, j, :]                    src [B, n, D],  dst [B, Ld, D]  (indices unique)
// D % 8 == 0; idx int64 [n], shared by the batch (one permutation per step, like torch.randperm in the reference).
__global__ void __launch_bounds__(256)
token_gather_kernel(const __nv_bfloat16* __restrict__ src, const long* __restrict__ idx, __nv_bfloat16* __restrict__ dst, int B,
                    long Ls, long n, int D, int scatter) {
  const int nch = D >> 3;
  const long total = static_cast<long>(B) * n * nch;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % nch);
    const long bj = i / nch;
    const long j = bj % n;
    const long b = bj / n;
    const long k = idx[j];
    if (!scatter) st_stream(dst + (b * n + j) * D + c * 8, ld_stream(src + (b * Ls + k) * D + c * 8));
    else st_stream(dst + (b * Ls + k) * D + c * 8, ld_stream(src + (b * n + j) * D + c * 8));
  }
}

}  // namespace vpt
