// Attention forward / backward for the head dimensions the tcgen05 kernels (attention.cuh, attention_bwd.cuh: 64 and 80)
// do not cover: 32, 96, 128.  CUDA-core kernels, correctness first.  Same semantics:
// softmax(q k^T * scale + key-length mask) v, non-causal, bf16 in / out, fp32 arithmetic, lse2 / delta with the row pitch
// rounded up to 128 (src/modules/attention.py:98-129, src/models/jit/denoiser.py:351-397 of the reference).
//
// One thread owns one row (a query row in the forward and in dQ, a key row in dK / dV): its vectors stay in registers, the
// rows of the other operand pass through shared memory in tiles of 32 and are read with broadcast loads, so there are no
// cross-lane reductions at all.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace vpt {

struct SimpleAttnTensor {
  const __nv_bfloat16* ptr;
  long sb, sl, sh;
};

struct SimpleAttnParams {
  int B, H, Lq, Lk, Lq_pad;
  const int* seqlens_k;
  float scale, scale_log2;
  SimpleAttnTensor q, k, v, o, d_o;
  float* lse2;             // [B, H, Lq_pad]
  const float* delta;      // [B, H, Lq_pad]  delta * scale
  float* dq;               // fp32 accumulator (written, not accumulated, here), strides dq_s*
  long dq_sb, dq_sl, dq_sh;
  __nv_bfloat16 *dk, *dv;
  long dk_sb, dk_sl, dk_sh, dv_sb, dv_sl, dv_sh;
};

constexpr int kSaRows = 128;   // rows per CTA (one per thread)
constexpr int kSaTile = 32;    // rows of the other operand per shared-memory tile

__device__ __forceinline__ float sa_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float sa_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// cooperative load of `rows` x HD bf16 (row r at base + r * stride) into smem [kSaTile][HD], zero beyond `valid`
template <int HD>
__device__ __forceinline__ void sa_load_tile(uint4* dst, const __nv_bfloat16* base, long stride, int valid) {
  constexpr int kCh = HD / 8;
  for (int i = threadIdx.x; i < kSaTile * kCh; i += kSaRows) {
    const int r = i / kCh, c = i % kCh;
    dst[i] = r < valid ? *reinterpret_cast<const uint4*>(base + r * stride + c * 8) : make_uint4(0, 0, 0, 0);
  }
}
// dot of a packed-bf16 register row with a shared-memory row (broadcast reads)
template <int HD>
__device__ __forceinline__ float sa_dot(const uint32_t (&a)[HD / 2], const uint4* row) {
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < HD / 8; ++c) {
    const uint4 w = row[c];
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) acc += sa_lo(a[c * 4 + e]) * sa_lo(ww[e]) + sa_hi(a[c * 4 + e]) * sa_hi(ww[e]);
  }
  return acc;
}
// acc[0..HD) += w * row
template <int HD>
__device__ __forceinline__ void sa_axpy(float (&acc)[HD], float w, const uint4* row) {
#pragma unroll
  for (int c = 0; c < HD / 8; ++c) {
    const uint4 v = row[c];
    const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      acc[c * 8 + 2 * e] += w * sa_lo(vv[e]);
      acc[c * 8 + 2 * e + 1] += w * sa_hi(vv[e]);
    }
  }
}
template <int HD>
__device__ __forceinline__ void sa_load_row(uint32_t (&a)[HD / 2], const __nv_bfloat16* p, bool live) {
#pragma unroll
  for (int c = 0; c < HD / 8; ++c) {
    const uint4 w = live ? *reinterpret_cast<const uint4*>(p + c * 8) : make_uint4(0, 0, 0, 0);
    a[c * 4] = w.x; a[c * 4 + 1] = w.y; a[c * 4 + 2] = w.z; a[c * 4 + 3] = w.w;
  }
}
template <int HD>
__device__ __forceinline__ void sa_store_row_bf16(__nv_bfloat16* p, const float (&acc)[HD], float s) {
#pragma unroll
  for (int c = 0; c < HD / 8; ++c) {
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      __nv_bfloat162 h = __floats2bfloat162_rn(acc[c * 8 + 2 * e] * s, acc[c * 8 + 2 * e + 1] * s);
      w[e] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p + c * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ------------------------------------------------------------------------------------------------ forward
template <int HD>
__global__ void __launch_bounds__(kSaRows)
attn_simple_fwd_kernel(const SimpleAttnParams p) {
  __shared__ uint4 sK[kSaTile * HD / 8], sV[kSaTile * HD / 8];
  const int h = blockIdx.y, b = blockIdx.z;
  const int qi = blockIdx.x * kSaRows + threadIdx.x;
  const bool live = qi < p.Lq;
  int klen = p.seqlens_k ? p.seqlens_k[b] : p.Lk;
  klen = klen < p.Lk ? klen : p.Lk;
  uint32_t q[HD / 2];
  sa_load_row<HD>(q, p.q.ptr + b * p.q.sb + static_cast<long>(qi) * p.q.sl + h * p.q.sh, live);
  float o[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) o[d] = 0.f;
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < klen; k0 += kSaTile) {
    const int valid = min(kSaTile, klen - k0);
    __syncthreads();
    sa_load_tile<HD>(sK, p.k.ptr + b * p.k.sb + static_cast<long>(k0) * p.k.sl + h * p.k.sh, p.k.sl, valid);
    sa_load_tile<HD>(sV, p.v.ptr + b * p.v.sb + static_cast<long>(k0) * p.v.sl + h * p.v.sh, p.v.sl, valid);
    __syncthreads();
    for (int j = 0; j < valid; ++j) {
      const float s = sa_dot<HD>(q, sK + j * (HD / 8)) * p.scale_log2;
      const float m_new = fmaxf(m, s);
      const float alpha = exp2f(m - m_new);          // first key: exp2(-inf) = 0
      const float pj = exp2f(s - m_new);
      l = l * alpha + pj;
#pragma unroll
      for (int d = 0; d < HD; ++d) o[d] *= alpha;
      sa_axpy<HD>(o, pj, sV + j * (HD / 8));
      m = m_new;
    }
  }
  if (qi < p.Lq_pad && blockIdx.x * kSaRows + threadIdx.x < p.Lq_pad)
    p.lse2[(static_cast<long>(b) * p.H + h) * p.Lq_pad + qi] = live ? (l > 0.f ? m + log2f(l) : -INFINITY) : INFINITY;
  if (live)
    sa_store_row_bf16<HD>(const_cast<__nv_bfloat16*>(p.o.ptr) + b * p.o.sb + static_cast<long>(qi) * p.o.sl + h * p.o.sh, o,
                          l > 0.f ? 1.f / l : 0.f);
}

// delta[b,h,q] = scale * sum_d dO * O   (row pitch Lq_pad, zero in the padding)
template <int HD>
__global__ void __launch_bounds__(256)
attn_simple_delta_kernel(const SimpleAttnParams p, float* delta) {
  const long gid = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long total = static_cast<long>(p.B) * p.H * p.Lq_pad;
  if (gid >= total) return;
  const int qi = gid % p.Lq_pad;
  const int h = (gid / p.Lq_pad) % p.H;
  const int b = gid / (static_cast<long>(p.Lq_pad) * p.H);
  float acc = 0.f;
  if (qi < p.Lq) {
    const __nv_bfloat16* o = p.o.ptr + b * p.o.sb + static_cast<long>(qi) * p.o.sl + h * p.o.sh;
    const __nv_bfloat16* g = p.d_o.ptr + b * p.d_o.sb + static_cast<long>(qi) * p.d_o.sl + h * p.d_o.sh;
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      const uint4 a = *reinterpret_cast<const uint4*>(o + c * 8), w = *reinterpret_cast<const uint4*>(g + c * 8);
      const uint32_t aa[4] = {a.x, a.y, a.z, a.w}, ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) acc += sa_lo(aa[e]) * sa_lo(ww[e]) + sa_hi(aa[e]) * sa_hi(ww[e]);
    }
  }
  delta[gid] = acc * p.scale;
}

// ------------------------------------------------------------------------------------------------ backward: dQ
// thread = query row; dq = sum_k dS[q,k] K[k],  dS = P (dP - delta) scale,  P = exp2(S c - lse),  dP = dO . V[k]
template <int HD>
__global__ void __launch_bounds__(kSaRows)
attn_simple_bwd_dq_kernel(const SimpleAttnParams p) {
  __shared__ uint4 sK[kSaTile * HD / 8], sV[kSaTile * HD / 8];
  const int h = blockIdx.y, b = blockIdx.z;
  const int qi = blockIdx.x * kSaRows + threadIdx.x;
  const bool live = qi < p.Lq;
  int klen = p.seqlens_k ? p.seqlens_k[b] : p.Lk;
  klen = klen < p.Lk ? klen : p.Lk;
  uint32_t q[HD / 2], g[HD / 2];
  sa_load_row<HD>(q, p.q.ptr + b * p.q.sb + static_cast<long>(qi) * p.q.sl + h * p.q.sh, live);
  sa_load_row<HD>(g, p.d_o.ptr + b * p.d_o.sb + static_cast<long>(qi) * p.d_o.sl + h * p.d_o.sh, live);
  const long sidx = (static_cast<long>(b) * p.H + h) * p.Lq_pad + qi;
  const float lse = live ? p.lse2[sidx] : INFINITY;
  const float dl = live ? p.delta[sidx] : 0.f;
  float acc[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) acc[d] = 0.f;
  for (int k0 = 0; k0 < klen; k0 += kSaTile) {
    const int valid = min(kSaTile, klen - k0);
    __syncthreads();
    sa_load_tile<HD>(sK, p.k.ptr + b * p.k.sb + static_cast<long>(k0) * p.k.sl + h * p.k.sh, p.k.sl, valid);
    sa_load_tile<HD>(sV, p.v.ptr + b * p.v.sb + static_cast<long>(k0) * p.v.sl + h * p.v.sh, p.v.sl, valid);
    __syncthreads();
    for (int j = 0; j < valid; ++j) {
      const float pj = exp2f(sa_dot<HD>(q, sK + j * (HD / 8)) * p.scale_log2 - lse);
      const float ds = pj * (sa_dot<HD>(g, sV + j * (HD / 8)) * p.scale - dl);
      sa_axpy<HD>(acc, ds, sK + j * (HD / 8));
    }
  }
  if (live) {
    float* dst = p.dq + b * p.dq_sb + static_cast<long>(qi) * p.dq_sl + h * p.dq_sh;
#pragma unroll
    for (int d = 0; d < HD; d += 4) *reinterpret_cast<float4*>(dst + d) = make_float4(acc[d], acc[d + 1], acc[d + 2], acc[d + 3]);
  }
}

// ------------------------------------------------------------------------------------------------ backward: dK / dV
// thread = key row.  kDV: dv = sum_q P[q,k] dO[q];  else: dk = sum_q dS[q,k] Q[q]
template <int HD, bool kDV>
__global__ void __launch_bounds__(kSaRows)
attn_simple_bwd_dkv_kernel(const SimpleAttnParams p) {
  __shared__ uint4 sQ[kSaTile * HD / 8], sG[kSaTile * HD / 8];
  __shared__ float sLse[kSaTile], sDl[kSaTile];
  const int h = blockIdx.y, b = blockIdx.z;
  const int ki = blockIdx.x * kSaRows + threadIdx.x;
  int klen = p.seqlens_k ? p.seqlens_k[b] : p.Lk;
  klen = klen < p.Lk ? klen : p.Lk;
  const bool in_range = ki < p.Lk;
  const bool live = ki < klen;                       // padded keys get exactly zero gradient
  uint32_t kr[HD / 2], vr[HD / 2];
  sa_load_row<HD>(kr, p.k.ptr + b * p.k.sb + static_cast<long>(ki) * p.k.sl + h * p.k.sh, live);
  sa_load_row<HD>(vr, p.v.ptr + b * p.v.sb + static_cast<long>(ki) * p.v.sl + h * p.v.sh, live && !kDV);
  float acc[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) acc[d] = 0.f;
  const long sbase = (static_cast<long>(b) * p.H + h) * p.Lq_pad;
  for (int q0 = 0; q0 < p.Lq; q0 += kSaTile) {
    const int valid = min(kSaTile, p.Lq - q0);
    __syncthreads();
    sa_load_tile<HD>(sQ, p.q.ptr + b * p.q.sb + static_cast<long>(q0) * p.q.sl + h * p.q.sh, p.q.sl, valid);
    sa_load_tile<HD>(sG, p.d_o.ptr + b * p.d_o.sb + static_cast<long>(q0) * p.d_o.sl + h * p.d_o.sh, p.d_o.sl, valid);
    if (threadIdx.x < kSaTile) {
      sLse[threadIdx.x] = threadIdx.x < valid ? p.lse2[sbase + q0 + threadIdx.x] : INFINITY;
      sDl[threadIdx.x] = threadIdx.x < valid ? p.delta[sbase + q0 + threadIdx.x] : 0.f;
    }
    __syncthreads();
    if (live) {
      for (int j = 0; j < valid; ++j) {
        const float pj = exp2f(sa_dot<HD>(kr, sQ + j * (HD / 8)) * p.scale_log2 - sLse[j]);
        if (kDV) {
          sa_axpy<HD>(acc, pj, sG + j * (HD / 8));
        } else {
          const float ds = pj * (sa_dot<HD>(vr, sG + j * (HD / 8)) * p.scale - sDl[j]);
          sa_axpy<HD>(acc, ds, sQ + j * (HD / 8));
        }
      }
    }
  }
  if (in_range) {
    __nv_bfloat16* dst = kDV ? p.dv + b * p.dv_sb + static_cast<long>(ki) * p.dv_sl + h * p.dv_sh
                             : p.dk + b * p.dk_sb + static_cast<long>(ki) * p.dk_sl + h * p.dk_sh;
    sa_store_row_bf16<HD>(dst, acc, 1.f);
  }
}

}  // namespace vpt
