// HBM-bound kernels of the JiT / DiT block: RMSNorm, QK-norm + RoPE, SwiGLU, LayerNorm + adaLN modulate,
// gate-residual, patchify / unpatchify.  All are one pass over the activation with 16-byte accesses; rows are
// handled by a warp (hidden width) or by an 8-lane group (one 64-wide head), statistics in fp32.
//
// Reference semantics (file:line under /root/reference):
//   FP32RMSNorm / FP32LayerNorm            src/modules/norm.py:9-27
//   q_norm/k_norm + apply_rope             src/models/jit/denoiser.py:98-111, 365-373
//   SwiGLU gate                            src/models/jit/denoiser.py:498-506
//   adaLN modulate / gate                  src/models/cogview4/denoiser.py:182-187, 401-420; src/modules/norm.py:75-83
//   patchify / unpatchify                  src/modules/patch.py:17-115; src/models/jit/denoiser.py:828-860
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "sm100.cuh"

namespace vpt {

__device__ __forceinline__ float ew_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float ew_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t ew_pack(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float ew_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
// Two values rounded to bf16 and back through ONE full-rate bf16x2 pack (F2FP) + two ALU ops.  ew_round costs a quarter-rate
// F2F per element, which put the kernels with several rounding points per element (adaLN modulate: 3) at ~0.5 of the HBM rate.
__device__ __forceinline__ void ew_round2(float& a, float& b) {
  const uint32_t p = ew_pack(a, b);
  a = ew_lo(p);
  b = ew_hi(p);
}
__device__ __forceinline__ void ew_unpack8(const uint4& v, float (&f)[8]) {
  f[0] = ew_lo(v.x); f[1] = ew_hi(v.x); f[2] = ew_lo(v.y); f[3] = ew_hi(v.y);
  f[4] = ew_lo(v.z); f[5] = ew_hi(v.z); f[6] = ew_lo(v.w); f[7] = ew_hi(v.w);
}
__device__ __forceinline__ uint4 ew_pack8(const float (&f)[8]) {
  return make_uint4(ew_pack(f[0], f[1]), ew_pack(f[2], f[3]), ew_pack(f[4], f[5]), ew_pack(f[6], f[7]));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint4 ld_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

constexpr int kEwMaxChunks = 8;   // 16B chunks per lane -> rows up to 2048 bf16 elements
constexpr int kEwThreads = 256;   // 8 rows per CTA

// ------------------------------------------------------------------------------------------- RMSNorm
// y = bf16( (x * rstd) * w ),  rstd = rsqrt(mean(x^2) + eps)   [rows, D], D % 8 == 0, D <= 2048
__global__ void __launch_bounds__(kEwThreads)
rmsnorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ y,
                   float* __restrict__ rstd_out, long rows, int D, long ldx, long ldy, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * (kEwThreads / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nch = D >> 3;
  uint4 v[kEwMaxChunks];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kEwMaxChunks; ++i) {
    const int c = lane + i * 32;
    if (c < nch) {
      v[i] = ld_stream(x + row * ldx + c * 8);
      float f[8];
      ew_unpack8(v[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) ss += f[e] * f[e];
    }
  }
  ss = warp_sum(ss);
  const float rstd = rsqrtf(ss / static_cast<float>(D) + eps);
  if (lane == 0 && rstd_out != nullptr) rstd_out[row] = rstd;
#pragma unroll
  for (int i = 0; i < kEwMaxChunks; ++i) {
    const int c = lane + i * 32;
    if (c < nch) {
      float f[8], g[8];
      ew_unpack8(v[i], f);
      if (w != nullptr) {
        ew_unpack8(*reinterpret_cast<const uint4*>(w + c * 8), g);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = (f[e] * rstd) * g[e];
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = f[e] * rstd;
      }
      st_stream(y + row * ldy + c * 8, ew_pack8(f));
    }
  }
}

// dx = w*dy*rstd - x * rstd^3 * mean(w*dy*x)  (+ dres);  optional dw[D] += sum_rows dy * x * rstd (fp32 atomics)
// kCh = 16-byte chunks per lane (D <= 256 * kCh): all three input rows (x, dy, dres) are requested before the first use,
// so a warp has 3 * kCh independent 16-byte loads in flight instead of stalling once per phase.
template <int kCh>
__global__ void __launch_bounds__(kEwThreads)
rmsnorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                   const __nv_bfloat16* __restrict__ w, const float* __restrict__ rstd_in,
                   const __nv_bfloat16* __restrict__ dres, __nv_bfloat16* __restrict__ dx, float* __restrict__ dw,
                   long rows, int D, long ld, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * (kEwThreads / 32) + (threadIdx.x >> 5);
  const int nch = D >> 3;
  const bool live = row < rows;
  uint4 vx[kCh], vg[kCh], vr[kCh];
#pragma unroll
  for (int i = 0; i < kCh; ++i) {
    const int c = lane + i * 32;
    vx[i] = vg[i] = vr[i] = make_uint4(0, 0, 0, 0);
    if (live && c < nch) {
      vx[i] = ld_stream(x + row * ld + c * 8);
      vg[i] = ld_stream(dy + row * ld + c * 8);
      if (dres != nullptr) vr[i] = ld_stream(dres + row * ld + c * 8);
    }
  }
  const float rstd_given = (live && rstd_in != nullptr) ? rstd_in[row] : 0.f;
  float dot = 0.f, ss = 0.f;
#pragma unroll
  for (int i = 0; i < kCh; ++i) {
    const int c = lane + i * 32;
    if (live && c < nch) {
      float fx[8], fg[8], fw[8];
      ew_unpack8(vx[i], fx);
      ew_unpack8(vg[i], fg);
      if (w != nullptr) ew_unpack8(*reinterpret_cast<const uint4*>(w + c * 8), fw);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        dot += fg[e] * (w != nullptr ? fw[e] : 1.f) * fx[e];
        ss += fx[e] * fx[e];
      }
    }
  }
  dot = warp_sum(dot);
  ss = warp_sum(ss);
  const float rstd = !live ? 0.f : (rstd_in != nullptr ? rstd_given : rsqrtf(ss / static_cast<float>(D) + eps));
  const float coef = dot * rstd * rstd * rstd / static_cast<float>(D);
#pragma unroll
  for (int i = 0; i < kCh; ++i) {
    const int c = lane + i * 32;
    if (live && c < nch) {
      float fx[8], fg[8], fw[8], fr[8], o[8];
      ew_unpack8(vx[i], fx);
      ew_unpack8(vg[i], fg);
      ew_unpack8(vr[i], fr);
      if (w != nullptr) ew_unpack8(*reinterpret_cast<const uint4*>(w + c * 8), fw);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = fg[e] * (w != nullptr ? fw[e] : 1.f) * rstd - fx[e] * coef + fr[e];
      st_stream(dx + row * ld + c * 8, ew_pack8(o));
      if (dw != nullptr) {
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(dw + c * 8 + e, fg[e] * fx[e] * rstd);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------- QK-norm + RoPE
// x: [tokens = B*L, H, HD] (row pitch ld elements); per head: n = bf16(rmsnorm(x) * w); rotate interleaved pairs
// (n[2i], n[2i+1]) by (cos, sin)[l, i]; l = token % L.  One GL-lane group per (token, HP consecutive heads), 8 elements per
// lane (GL = 8 for head_dim 64, 16 with the upper lanes idle for 80 / 96 / 128): the rotation table and the norm weight are
// fetched once for the HP heads and their 16-byte loads are all in flight together.
template <int HP, int HD>
__global__ void __launch_bounds__(256)
qknorm_rope_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                       const float* __restrict__ cs /* [L, HD/2, 2] */, __nv_bfloat16* __restrict__ y, long tokens, int H,
                       int L, long ldx, long ldy, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int GL = HD <= 64 ? 8 : 16;
  constexpr int GS = GL == 8 ? 3 : 4;
  const int hg = H / HP;
  const long g = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) >> GS;
  const int sub = threadIdx.x & (GL - 1);
  const bool lane_on = sub * 8 < HD;
  const bool live = g < tokens * hg;
  const bool act = live && lane_on;
  const long tok = live ? g / hg : 0;
  const int h0 = live ? static_cast<int>(g % hg) * HP : 0;
  uint4 v[HP];
#pragma unroll
  for (int j = 0; j < HP; ++j) v[j] = act ? ld_stream(x + tok * ldx + (h0 + j) * HD + sub * 8) : make_uint4(0, 0, 0, 0);
  float fw[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float c[4] = {0.f, 0.f, 0.f, 0.f}, s[4] = {0.f, 0.f, 0.f, 0.f};
  if (lane_on) {
    ew_unpack8(*reinterpret_cast<const uint4*>(w + sub * 8), fw);
    const float4* t = reinterpret_cast<const float4*>(cs + (static_cast<long>(tok % L) * (HD / 2) + sub * 4) * 2);
    const float4 t0 = t[0], t1 = t[1];
    c[0] = t0.x; c[1] = t0.z; c[2] = t1.x; c[3] = t1.z;
    s[0] = t0.y; s[1] = t0.w; s[2] = t1.y; s[3] = t1.w;
  }
#pragma unroll
  for (int j = 0; j < HP; ++j) {
    float f[8];
    ew_unpack8(v[j], f);
    float ss = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) ss += f[e] * f[e];
#pragma unroll
    for (int o = 1; o < GL; o <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss * (1.f / HD) + eps);
    float o[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float a = ew_round((f[2 * i] * rstd) * fw[2 * i]);       // the reference rounds to bf16 between norm and rope
      const float b = ew_round((f[2 * i + 1] * rstd) * fw[2 * i + 1]);
      o[2 * i] = a * c[i] - b * s[i];
      o[2 * i + 1] = a * s[i] + b * c[i];
    }
    if (act) st_stream(y + tok * ldy + (h0 + j) * HD + sub * 8, ew_pack8(o));
  }
}

// dy is fp32 (the attention dQ accumulator) or bf16.  dn = R^T dy ; dx = rmsnorm_bwd(dn) with weight w.
template <bool kDyF32, int HP, int HD>
__global__ void __launch_bounds__(256)
qknorm_rope_bwd_kernel(const void* __restrict__ dy_, const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                       const float* __restrict__ cs, __nv_bfloat16* __restrict__ dx, float* __restrict__ dw, long tokens,
                       int H, int L, long lddy, long ldx, long lddx, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int GL = HD <= 64 ? 8 : 16;
  constexpr int GS = GL == 8 ? 3 : 4;
  const int hg = H / HP;
  const long g = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) >> GS;
  const int sub = threadIdx.x & (GL - 1);
  const bool lane_on = sub * 8 < HD;
  const bool live = g < tokens * hg;
  const bool act = live && lane_on;
  const long tok = live ? g / hg : 0;
  const int h0 = live ? static_cast<int>(g % hg) * HP : 0;
  uint4 vx[HP];
  float d[HP][8];
#pragma unroll
  for (int j = 0; j < HP; ++j) {
    vx[j] = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int e = 0; e < 8; ++e) d[j][e] = 0.f;
    if (act) {
      vx[j] = ld_stream(x + tok * ldx + (h0 + j) * HD + sub * 8);
      if (kDyF32) {
        const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(dy_) + tok * lddy + (h0 + j) * HD + sub * 8);
        const float4 a = p[0], b = p[1];
        d[j][0] = a.x; d[j][1] = a.y; d[j][2] = a.z; d[j][3] = a.w; d[j][4] = b.x; d[j][5] = b.y; d[j][6] = b.z; d[j][7] = b.w;
      } else {
        ew_unpack8(ld_stream(static_cast<const __nv_bfloat16*>(dy_) + tok * lddy + (h0 + j) * HD + sub * 8), d[j]);
      }
    }
  }
  float fw[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float c[4] = {0.f, 0.f, 0.f, 0.f}, s[4] = {0.f, 0.f, 0.f, 0.f};
  if (lane_on) {
    ew_unpack8(*reinterpret_cast<const uint4*>(w + sub * 8), fw);
    const float4* t = reinterpret_cast<const float4*>(cs + (static_cast<long>(tok % L) * (HD / 2) + sub * 4) * 2);
    const float4 t0 = t[0], t1 = t[1];
    c[0] = t0.x; c[1] = t0.z; c[2] = t1.x; c[3] = t1.z;
    s[0] = t0.y; s[1] = t0.w; s[2] = t1.y; s[3] = t1.w;
  }
  float dwacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < HP; ++j) {
    float f[8], dn[8];
    ew_unpack8(vx[j], f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      dn[2 * i] = d[j][2 * i] * c[i] + d[j][2 * i + 1] * s[i];
      dn[2 * i + 1] = -d[j][2 * i] * s[i] + d[j][2 * i + 1] * c[i];
    }
    float ss = 0.f, dot = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      ss += f[e] * f[e];
      dot += dn[e] * fw[e] * f[e];
    }
#pragma unroll
    for (int o = 1; o < GL; o <<= 1) {
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
      dot += __shfl_xor_sync(0xffffffffu, dot, o);
    }
    const float rstd = rsqrtf(ss * (1.f / HD) + eps);
    const float coef = dot * rstd * rstd * rstd * (1.f / HD);
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      o[e] = dn[e] * fw[e] * rstd - f[e] * coef;
      dwacc[e] += dn[e] * f[e] * rstd;
    }
    if (act) st_stream(dx + tok * lddx + (h0 + j) * HD + sub * 8, ew_pack8(o));
  }
  if (act && dw != nullptr) {
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(dw + sub * 8 + e, dwacc[e]);
  }
}

// ------------------------------------------------------------------------------------------- SwiGLU gate
// a = bf16( bf16(silu(g)) * u )     [rows, F]; every row pitch >= F rounded up to 8 (the ragged tail chunk is computed
// on the padding too and lands in padding)
__global__ void __launch_bounds__(256)
swiglu_fwd_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ u, __nv_bfloat16* __restrict__ a,
                  long rows, int F, long ldg, long ldu, long lda) {
  pdl_launch_dependents();
  pdl_wait();
  const int nch = (F + 7) >> 3;
  const long total = rows * nch;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / nch;
    const int c = static_cast<int>(i % nch);
    float fg[8], fu[8], o[8];
    ew_unpack8(ld_stream(g + r * ldg + c * 8), fg);
    ew_unpack8(ld_stream(u + r * ldu + c * 8), fu);
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      float sa_ = fg[e] / (1.f + __expf(-fg[e])), sb_ = fg[e + 1] / (1.f + __expf(-fg[e + 1]));
      ew_round2(sa_, sb_);
      o[e] = sa_ * fu[e];
      o[e + 1] = sb_ * fu[e + 1];
    }
    st_stream(a + r * lda + c * 8, ew_pack8(o));
  }
}
// dg = da * u * silu'(g),  du = da * silu(g)
__global__ void __launch_bounds__(256)
swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ da, const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ u,
                  __nv_bfloat16* __restrict__ dg, __nv_bfloat16* __restrict__ du, long rows, int F, long ldda, long ldg,
                  long ldu, long lddg, long lddu) {
  pdl_launch_dependents();
  pdl_wait();
  const int nch = (F + 7) >> 3;
  const long total = rows * nch;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / nch;
    const int c = static_cast<int>(i % nch);
    float fa[8], fg[8], fu[8], og[8], ou[8];
    ew_unpack8(ld_stream(da + r * ldda + c * 8), fa);
    ew_unpack8(ld_stream(g + r * ldg + c * 8), fg);
    ew_unpack8(ld_stream(u + r * ldu + c * 8), fu);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float sg = 1.f / (1.f + __expf(-fg[e]));
      const float silu = fg[e] * sg;
      ou[e] = fa[e] * silu;
      og[e] = fa[e] * fu[e] * (sg * (1.f + fg[e] * (1.f - sg)));
    }
    st_stream(dg + r * lddg + c * 8, ew_pack8(og));
    st_stream(du + r * lddu + c * 8, ew_pack8(ou));
  }
}

// ------------------------------------------------------------------------------------------- LayerNorm + adaLN
// n = bf16(LN_noaffine(x)); y = bf16( bf16(n * bf16(1 + scale[b])) + shift[b] );  x [B*L, D], scale/shift [B, D]
// kCh = 16-byte chunks per lane: D <= 256 * kCh (CogView4's hidden width 4096 needs 16)
template <int kCh>
__global__ void __launch_bounds__(kEwThreads)
ln_modulate_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ scale,
                       const __nv_bfloat16* __restrict__ shift, __nv_bfloat16* __restrict__ y, float* __restrict__ mean_out,
                       float* __restrict__ rstd_out, long rows, int L, int D, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * (kEwThreads / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const long b = row / L;
  const int nch = D >> 3;
  uint4 v[kCh];
#pragma unroll
  for (int i = 0; i < kCh; ++i) {
    const int c = lane + i * 32;
    v[i] = make_uint4(0, 0, 0, 0);
    if (c < nch) v[i] = ld_stream(x + row * D + c * 8);
  }
  // mean and variance in ONE reduction round: moments of (x - x0), x0 = the row's first element (a shift keeps the
  // single-pass form free of cancellation when |mean| >> std); two dependent reductions cost the row a second latency chain
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
      float f[8], sc[8], sh[8], o[8];
      ew_unpack8(v[i], f);
      ew_unpack8(*reinterpret_cast<const uint4*>(scale + b * D + c * 8), sc);
      ew_unpack8(*reinterpret_cast<const uint4*>(shift + b * D + c * 8), sh);
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        float na = (f[e] - mean) * rstd, nb = (f[e + 1] - mean) * rstd;
        float sa_ = 1.f + sc[e], sb_ = 1.f + sc[e + 1];
        ew_round2(na, nb);
        ew_round2(sa_, sb_);
        float ta = na * sa_, tb = nb * sb_;
        ew_round2(ta, tb);
        o[e] = ta + sh[e];
        o[e + 1] = tb + sh[e + 1];
      }
      st_stream(y + row * D + c * 8, ew_pack8(o));
    }
  }
}
// dx = LN_bwd(dy * (1 + scale));  dscale[b] += sum_l dy * n;  dshift[b] += sum_l dy   (fp32 atomics, [B, D])
// All loads of the row (x, dy) are requested before the first use: 2 * kCh independent 16-byte loads in flight per lane
// (interleaving load and use left the D = 4096 rows at 0.34 of the HBM rate, profiles/r2h_membound.txt).
template <int kCh>
__global__ void __launch_bounds__(kEwThreads)
ln_modulate_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                       const __nv_bfloat16* __restrict__ scale, const float* __restrict__ mean_in,
                       const float* __restrict__ rstd_in, __nv_bfloat16* __restrict__ dx, float* __restrict__ dscale,
                       float* __restrict__ dshift, long rows, int L, int D) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * (kEwThreads / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const long b = row / L;
  const int nch = D >> 3;
  uint4 vx[kCh], vg[kCh];
#pragma unroll
  for (int i = 0; i < kCh; ++i) {
    const int c = lane + i * 32;
    vx[i] = vg[i] = make_uint4(0, 0, 0, 0);
    if (c < nch) {
      vx[i] = ld_stream(x + row * D + c * 8);
      vg[i] = ld_stream(dy + row * D + c * 8);
    }
  }
  const float mean = mean_in[row], rstd = rstd_in[row];
  float sa = 0.f, sb = 0.f;   // sum(dn), sum(dn * n)
#pragma unroll
  for (int i = 0; i < kCh; ++i) {
    const int c = lane + i * 32;
    if (c < nch) {
      float fx[8], fg[8], sc[8];
      ew_unpack8(vx[i], fx);
      ew_unpack8(vg[i], fg);
      ew_unpack8(*reinterpret_cast<const uint4*>(scale + b * D + c * 8), sc);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float n = (fx[e] - mean) * rstd;
        const float dn = fg[e] * (1.f + sc[e]);
        sa += dn;
        sb += dn * n;
        if (dscale != nullptr) {
          atomicAdd(dscale + b * D + c * 8 + e, fg[e] * n);
          atomicAdd(dshift + b * D + c * 8 + e, fg[e]);
        }
      }
    }
  }
  sa = warp_sum(sa) / static_cast<float>(D);
  sb = warp_sum(sb) / static_cast<float>(D);
#pragma unroll
  for (int i = 0; i < kCh; ++i) {
    const int c = lane + i * 32;
    if (c < nch) {
      float fx[8], fg[8], sc[8], o[8];
      ew_unpack8(vx[i], fx);
      ew_unpack8(vg[i], fg);
      ew_unpack8(*reinterpret_cast<const uint4*>(scale + b * D + c * 8), sc);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float n = (fx[e] - mean) * rstd;
        o[e] = rstd * (fg[e] * (1.f + sc[e]) - sa - n * sb);
      }
      st_stream(dx + row * D + c * 8, ew_pack8(o));
    }
  }
}

// Wide rows (D > 2048): the row is NOT cached in registers -- pass 1 reads x, dy for the two row sums, pass 2 reads them again
// (from L2: the row was just there) and writes dx.  ~40 registers instead of ~200, so six times as many rows are in flight per
// SM and one warp's arithmetic hides under the others' loads; HBM traffic is unchanged (profiles/r2_membound.json).
__global__ void __launch_bounds__(kEwThreads)
ln_modulate_bwd_wide_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                            const __nv_bfloat16* __restrict__ scale, const float* __restrict__ mean_in,
                            const float* __restrict__ rstd_in, __nv_bfloat16* __restrict__ dx, float* __restrict__ dscale,
                            float* __restrict__ dshift, long rows, int L, int D, int affine) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * (kEwThreads / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const long b = affine ? 0 : row / L;             // affine: `scale` is the LayerNorm weight [D] (or nullptr), dn = dy * w
  const int nch = D >> 3;
  const float mean = mean_in[row], rstd = rstd_in[row];
  float sa = 0.f, sb = 0.f;
  for (int c0 = lane; c0 < nch; c0 += 128) {       // four independent 16-byte loads of each input per trip
    uint4 vx[4], vg[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + u * 32;
      vx[u] = vg[u] = make_uint4(0, 0, 0, 0);
      if (c < nch) {
        vx[u] = *reinterpret_cast<const uint4*>(x + row * D + c * 8);
        vg[u] = *reinterpret_cast<const uint4*>(dy + row * D + c * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + u * 32;
      if (c < nch) {
        float fx[8], fg[8], sc[8];
        ew_unpack8(vx[u], fx);
        ew_unpack8(vg[u], fg);
#pragma unroll
        for (int e = 0; e < 8; ++e) sc[e] = affine ? 1.f : 0.f;
        if (scale != nullptr) ew_unpack8(*reinterpret_cast<const uint4*>(scale + b * D + c * 8), sc);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float n = (fx[e] - mean) * rstd;
          const float dn = fg[e] * (affine ? sc[e] : 1.f + sc[e]);
          sa += dn;
          sb += dn * n;
          if (dscale != nullptr) {
            atomicAdd(dscale + b * D + c * 8 + e, fg[e] * n);
            atomicAdd(dshift + b * D + c * 8 + e, fg[e]);
          }
        }
      }
    }
  }
  sa = warp_sum(sa) / static_cast<float>(D);
  sb = warp_sum(sb) / static_cast<float>(D);
  for (int c0 = lane; c0 < nch; c0 += 128) {
    uint4 vx[4], vg[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + u * 32;
      vx[u] = vg[u] = make_uint4(0, 0, 0, 0);
      if (c < nch) {
        vx[u] = *reinterpret_cast<const uint4*>(x + row * D + c * 8);
        vg[u] = *reinterpret_cast<const uint4*>(dy + row * D + c * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + u * 32;
      if (c < nch) {
        float fx[8], fg[8], sc[8], o[8];
        ew_unpack8(vx[u], fx);
        ew_unpack8(vg[u], fg);
#pragma unroll
        for (int e = 0; e < 8; ++e) sc[e] = affine ? 1.f : 0.f;
        if (scale != nullptr) ew_unpack8(*reinterpret_cast<const uint4*>(scale + b * D + c * 8), sc);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float n = (fx[e] - mean) * rstd;
          o[e] = rstd * (fg[e] * (affine ? sc[e] : 1.f + sc[e]) - sa - n * sb);
        }
        st_stream(dx + row * D + c * 8, ew_pack8(o));
      }
    }
  }
}

// y = bf16( x + bf16(h * gate[b]) )
__global__ void __launch_bounds__(256)
gate_residual_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ h,
                         const __nv_bfloat16* __restrict__ gate, __nv_bfloat16* __restrict__ y, long rows, int L, int D) {
  const int nch = D >> 3;
  const long total = rows * nch;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / nch;
    const int c = static_cast<int>(i % nch);
    float fx[8], fh[8], fg[8], o[8];
    ew_unpack8(ld_stream(x + r * D + c * 8), fx);
    ew_unpack8(ld_stream(h + r * D + c * 8), fh);
    ew_unpack8(*reinterpret_cast<const uint4*>(gate + (r / L) * D + c * 8), fg);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = fx[e] + ew_round(fh[e] * fg[e]);
    st_stream(y + r * D + c * 8, ew_pack8(o));
  }
}
// dh = dy * gate[b];  dgate[b] += sum_l dy * h  (fp32 atomics);  dx = dy is the caller's alias
__global__ void __launch_bounds__(kEwThreads)
gate_residual_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ h,
                         const __nv_bfloat16* __restrict__ gate, __nv_bfloat16* __restrict__ dh, float* __restrict__ dgate,
                         long rows, int L, int D) {
  const int nch = D >> 3;
  const long total = rows * nch;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / nch;
    const int c = static_cast<int>(i % nch);
    const long b = r / L;
    float fd[8], fh[8], fg[8], o[8];
    ew_unpack8(ld_stream(dy + r * D + c * 8), fd);
    ew_unpack8(ld_stream(h + r * D + c * 8), fh);
    ew_unpack8(*reinterpret_cast<const uint4*>(gate + b * D + c * 8), fg);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      o[e] = fd[e] * fg[e];
      if (dgate != nullptr) atomicAdd(dgate + b * D + c * 8 + e, fd[e] * fh[e]);
    }
    st_stream(dh + r * D + c * 8, ew_pack8(o));
  }
}

// ------------------------------------------------------------------------------------------- patchify
// img [B, C, Himg, Wimg]  <->  patches [B, (Himg/p)*(Wimg/p), C*p*p]
//   order 0: patch vector ordered (c, py, px)   src/modules/patch.py:39-54 and the conv patch-embed weight
//   order 1: patch vector ordered (py, px, c)   JiT._unpatchify, src/models/jit/denoiser.py:845-858
// One thread per patch element of 2 bytes x `vec` contiguous pixels along px when order 0 (vec = gcd(p, 8)).
template <bool kToPatches>
__global__ void __launch_bounds__(256)
patch_permute_kernel(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, int B, int C, int Himg, int Wimg, int p,
                     int order) {
  const int hp = Himg / p, wp = Wimg / p;
  const long total = static_cast<long>(B) * C * Himg * Wimg;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    // i enumerates the image in [B, C, Himg, Wimg] order (coalesced on the image side)
    const int xw = static_cast<int>(i % Wimg);
    const int yh = static_cast<int>((i / Wimg) % Himg);
    const int c = static_cast<int>((i / (static_cast<long>(Wimg) * Himg)) % C);
    const long b = i / (static_cast<long>(Wimg) * Himg * C);
    const int py = yh % p, px = xw % p;
    const long n = static_cast<long>(yh / p) * wp + xw / p;
    const long e = order == 0 ? (static_cast<long>(c) * p + py) * p + px : (static_cast<long>(py) * p + px) * C + c;
    const long j = (b * hp * wp + n) * (static_cast<long>(C) * p * p) + e;
    if (kToPatches) dst[j] = src[i]; else dst[i] = src[j];
  }
}

// Vector form for p % 8 == 0 (the JiT / CogView patch sizes after packing) and C <= 4: one thread moves the 8 pixels x..x+7
// of one image row for ALL channels -- C 16-byte accesses on the image side, and on the patch side either C separate
// 16-byte runs (order 0: c outermost) or one contiguous 16*C-byte run with the channels interleaved (order 1: c innermost).
template <bool kToPatches, int C>
__global__ void __launch_bounds__(256)
patch_permute_vec_kernel(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, int B, int Himg, int Wimg, int p, int order) {
  const int wp = Wimg / p, hp = Himg / p;
  const int w8 = Wimg >> 3;
  const long total = static_cast<long>(B) * Himg * w8;
  const long plane = static_cast<long>(Himg) * Wimg;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % w8) << 3;
    const int yh = static_cast<int>((i / w8) % Himg);
    const long b = i / (static_cast<long>(w8) * Himg);
    const int py = yh % p, px = x % p;
    const long n = static_cast<long>(yh / p) * wp + x / p;
    const long pbase = (b * hp * wp + n) * (static_cast<long>(C) * p * p);
    const long ibase = b * C * plane + static_cast<long>(yh) * Wimg + x;
    const uint16_t* img_c = kToPatches ? src : dst;
    uint16_t* img_m = const_cast<uint16_t*>(img_c);
    const uint16_t* pat_c = kToPatches ? dst : src;
    uint16_t* pat_m = const_cast<uint16_t*>(pat_c);
    if (order == 0) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const long pj = pbase + (static_cast<long>(c) * p + py) * p + px;
        if (kToPatches) *reinterpret_cast<uint4*>(pat_m + pj) = ld_stream(img_c + ibase + c * plane);
        else st_stream(img_m + ibase + c * plane, *reinterpret_cast<const uint4*>(pat_c + pj));
      }
    } else {
      const long pj = pbase + (static_cast<long>(py) * p + px) * C;      // 8 pixels x C channels contiguous
      uint16_t v[8 * C];
      if (kToPatches) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const uint4 q = ld_stream(img_c + ibase + c * plane);
          const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e * C + c] = static_cast<uint16_t>(w[e >> 1] >> (16 * (e & 1)));
        }
#pragma unroll
        for (int k = 0; k < C; ++k)
          *reinterpret_cast<uint4*>(pat_m + pj + k * 8) =
              make_uint4(v[k * 8 + 0] | (uint32_t(v[k * 8 + 1]) << 16), v[k * 8 + 2] | (uint32_t(v[k * 8 + 3]) << 16),
                         v[k * 8 + 4] | (uint32_t(v[k * 8 + 5]) << 16), v[k * 8 + 6] | (uint32_t(v[k * 8 + 7]) << 16));
      } else {
#pragma unroll
        for (int k = 0; k < C; ++k) {
          const uint4 q = *reinterpret_cast<const uint4*>(pat_c + pj + k * 8);
          const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) v[k * 8 + e] = static_cast<uint16_t>(w[e >> 1] >> (16 * (e & 1)));
        }
#pragma unroll
        for (int c = 0; c < C; ++c)
          st_stream(img_m + ibase + c * plane,
                    make_uint4(v[0 * C + c] | (uint32_t(v[1 * C + c]) << 16), v[2 * C + c] | (uint32_t(v[3 * C + c]) << 16),
                               v[4 * C + c] | (uint32_t(v[5 * C + c]) << 16), v[6 * C + c] | (uint32_t(v[7 * C + c]) << 16)));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- strided row copy / fill
// dst[r, 0:row_bytes] = src[r, 0:row_bytes] (or 0 when src == nullptr) for `rows` rows a pitch apart: the per-block refresh
// of JiT's context-token slots (reference denoiser.py:1092-1113) and the zeroing of their gradient.  16-byte accesses;
// torch's generic strided copy moves these 2-byte elements one at a time (15.6 us for 6 MB, profiles/r1n_*).
__global__ void __launch_bounds__(256)
copy_rows_kernel(uint8_t* __restrict__ dst, long dst_pitch, const uint8_t* __restrict__ src, long src_pitch, long rows,
                 long row_vec) {
  pdl_launch_dependents();
  pdl_wait();
  const long total = rows * row_vec;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / row_vec, c = i - r * row_vec;
    const uint4 v = src != nullptr ? __ldg(reinterpret_cast<const uint4*>(src + r * src_pitch) + c) : make_uint4(0u, 0u, 0u, 0u);
    reinterpret_cast<uint4*>(dst + r * dst_pitch)[c] = v;
  }
}

}  // namespace vpt
