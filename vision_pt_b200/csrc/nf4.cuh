// Standalone NF4 kernels: bit-exact dequantisation (what bitsandbytes' dequantize_blockwise + dequantize_4bit produce
// for `BnbLinear4bit`, /root/reference/src/modules/quant/bnb.py:37-129) and setup-time quantisation
// (quantize_4bit with compress_statistics=True, /root/reference/src/modules/quant/functional.py:342-371).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vpt {

enum { kDtBf16 = 0, kDtF16 = 1, kDtF32 = 2 };

// absmax[i] = fl32( fl32(nested_code[q[i]] * nested_absmax[i / 256]) + offset )
__global__ void nf4_absmax_kernel(const uint8_t* __restrict__ q, const float* __restrict__ nested_absmax,
                                  const float* __restrict__ nested_code, float offset, float* __restrict__ absmax, long nb) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < nb) absmax[i] = __fadd_rn(__fmul_rn(nested_code[q[i]], nested_absmax[i >> 8]), offset);
}

template <int kDt>
__device__ __forceinline__ void nf4_store2(void* out, long idx, float a, float b) {
  if (kDt == kDtBf16) {
    reinterpret_cast<__nv_bfloat162*>(out)[idx >> 1] = __floats2bfloat162_rn(a, b);
  } else if (kDt == kDtF16) {
    reinterpret_cast<__half2*>(out)[idx >> 1] = __floats2half2_rn(a, b);
  } else {
    reinterpret_cast<float2*>(out)[idx >> 1] = make_float2(a, b);
  }
}

// One thread per packed byte pair group: 4 bytes -> 8 weights.  n = number of weights (even).
template <int kDt>
__global__ void __launch_bounds__(256)
nf4_dequant_kernel(const uint8_t* __restrict__ packed, const uint8_t* __restrict__ qabsmax,
                   const float* __restrict__ nested_absmax, const float* __restrict__ nested_code,
                   const float* __restrict__ code, float offset, void* __restrict__ out, long n) {
  __shared__ float s_code[16];
  __shared__ float s_ncode[256];
  if (threadIdx.x < 16) s_code[threadIdx.x] = code[threadIdx.x];
  s_ncode[threadIdx.x] = nested_code[threadIdx.x];
  __syncthreads();
  const long nbytes = (n + 1) >> 1;
  for (long byte0 = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; byte0 < nbytes;
       byte0 += static_cast<long>(gridDim.x) * blockDim.x * 4) {
    const long e0 = byte0 * 2;                       // 8 weights, all in one 64-block (e0 % 8 == 0)
    const long blk = e0 >> 6;
    const float am = __fadd_rn(__fmul_rn(s_ncode[qabsmax[blk]], nested_absmax[blk >> 8]), offset);
    uint32_t w;
    if (byte0 + 4 <= nbytes) {
      w = *reinterpret_cast<const uint32_t*>(packed + byte0);
    } else {
      w = 0;
      for (int b = 0; byte0 + b < nbytes; ++b) w |= static_cast<uint32_t>(packed[byte0 + b]) << (8 * b);
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const long e = e0 + 2 * b;
      if (e >= n) break;
      const uint32_t byte = (w >> (8 * b)) & 0xffu;
      const float hi = __fmul_rn(s_code[byte >> 4], am);
      const float lo = __fmul_rn(s_code[byte & 15u], am);
      if (e + 1 < n) {
        nf4_store2<kDt>(out, e, hi, lo);
      } else {
        if (kDt == kDtBf16) reinterpret_cast<__nv_bfloat16*>(out)[e] = __float2bfloat16_rn(hi);
        else if (kDt == kDtF16) reinterpret_cast<__half*>(out)[e] = __float2half_rn(hi);
        else reinterpret_cast<float*>(out)[e] = hi;
      }
    }
  }
}

// Dequantise the FLATTENED bitsandbytes layout into a row-pitched bf16 matrix out[n * ld + k] (ld >= K, ld % 8 == 0):
// the operand the TMA-fed GEMM reads.  Values are the same bits as nf4_dequant_kernel<bf16>; only the addressing
// differs, so ragged in_features (2730, 3413: rows that start mid-byte / mid-block) need no repacked copy.
// One thread = one 32-bit word of codes = 8 weights of one 64-block.
__global__ void __launch_bounds__(256)
nf4_dequant_pitched_kernel(const uint8_t* __restrict__ packed, const uint8_t* __restrict__ qabsmax,
                           const float* __restrict__ nested_absmax, const float* __restrict__ nested_code,
                           const float* __restrict__ code, float offset, __nv_bfloat16* __restrict__ out, long n, int K,
                           long ld) {
  __shared__ float s_code[16];
  __shared__ float s_ncode[256];
  if (threadIdx.x < 16) s_code[threadIdx.x] = code[threadIdx.x];
  s_ncode[threadIdx.x] = nested_code[threadIdx.x];
  __syncthreads();
  const long nbytes = (n + 1) >> 1;
  for (long byte0 = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; byte0 < nbytes;
       byte0 += static_cast<long>(gridDim.x) * blockDim.x * 4) {
    const long e0 = byte0 * 2;
    const long blk = e0 >> 6;
    const float am = __fadd_rn(__fmul_rn(s_ncode[qabsmax[blk]], nested_absmax[blk >> 8]), offset);
    uint32_t w;
    if (byte0 + 4 <= nbytes) {
      w = *reinterpret_cast<const uint32_t*>(packed + byte0);
    } else {
      w = 0;
      for (int b = 0; byte0 + b < nbytes; ++b) w |= static_cast<uint32_t>(packed[byte0 + b]) << (8 * b);
    }
    uint32_t o[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const uint32_t byte = (w >> (8 * b)) & 0xffu;
      const __nv_bfloat162 h = __floats2bfloat162_rn(__fmul_rn(s_code[byte >> 4], am), __fmul_rn(s_code[byte & 15u], am));
      o[b] = *reinterpret_cast<const uint32_t*>(&h);
    }
    const long row = e0 / K;
    const int col = static_cast<int>(e0 - row * K);
    __nv_bfloat16* dst = out + row * ld + col;
    if (col + 8 <= K && e0 + 8 <= n && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
    } else {
      long r = row;
      int c = col;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (e0 + e >= n) break;
        const uint32_t v = o[e >> 1];
        reinterpret_cast<unsigned short*>(out)[r * ld + c] = (e & 1) ? static_cast<unsigned short>(v >> 16) : static_cast<unsigned short>(v & 0xffffu);
        if (++c == K) {
          c = 0;
          ++r;
        }
      }
    }
  }
}

// Dequantise into the TRANSPOSED bf16 matrix out[k * ld + n] (ld >= N rounded up to 8): the operand of the backward
// GEMM dX = dY * W, which then reads the weight K-major exactly like the forward.  Same bits as nf4_dequant_kernel<bf16>.
// One CTA = one 64 (n) x 64 (k) tile, transposed through shared memory.  The block after the last tile transposes the two
// LoRA matrices the backward kernel wants row-major in the other direction:
//   upT[r * ldu + n] = up[n * 16 + r]       downT[k * 16 + r] = down[r * ldd + k]
__global__ void __launch_bounds__(256)
nf4_dequant_transposed_kernel(const uint8_t* __restrict__ packed, const uint8_t* __restrict__ qabsmax,
                              const float* __restrict__ nested_absmax, const float* __restrict__ nested_code,
                              const float* __restrict__ code, float offset, __nv_bfloat16* __restrict__ out, int N, int K,
                              long ld, int tiles_k, int num_tiles, const __nv_bfloat16* __restrict__ up,
                              __nv_bfloat16* __restrict__ upT, long ldu, const __nv_bfloat16* __restrict__ down, long ldd,
                              __nv_bfloat16* __restrict__ downT) {
  if (static_cast<int>(blockIdx.x) >= num_tiles) {
    if (up == nullptr) return;
    const int t0 = (blockIdx.x - num_tiles) * blockDim.x + threadIdx.x;
    const int stride = (gridDim.x - num_tiles) * blockDim.x;
    for (int i = t0; i < 16 * static_cast<int>(ldu); i += stride) {
      const int r = i / static_cast<int>(ldu), n = i % static_cast<int>(ldu);
      upT[i] = n < N ? up[static_cast<long>(n) * 16 + r] : __float2bfloat16_rn(0.f);
    }
    for (int i = t0; i < K * 16; i += stride) {
      const int k = i >> 4, r = i & 15;
      downT[i] = down[static_cast<long>(r) * ldd + k];
    }
    return;
  }
  __shared__ float s_code[16];
  __shared__ float s_ncode[256];
  __shared__ unsigned short tile[64][66];
  if (threadIdx.x < 16) s_code[threadIdx.x] = code[threadIdx.x];
  s_ncode[threadIdx.x] = nested_code[threadIdx.x];
  __syncthreads();
  const int n0 = (blockIdx.x / tiles_k) * 64, k0 = (blockIdx.x % tiles_k) * 64;
  const bool aligned = (K & 7) == 0;
  for (int i = threadIdx.x; i < 512; i += 256) {
    const int r = i >> 3, g = i & 7;
    const int n = n0 + r, k = k0 + g * 8;
    unsigned short v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0;
    if (n < N && k < K) {
      const long flat = static_cast<long>(n) * K + k;
      if (aligned) {                               // 8 codes = one 32-bit word inside one 64-block
        const long blk = flat >> 6;
        const float am = __fadd_rn(__fmul_rn(s_ncode[qabsmax[blk]], nested_absmax[blk >> 8]), offset);
        const uint32_t w = *reinterpret_cast<const uint32_t*>(packed + (flat >> 1));
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const uint32_t byte = (w >> (8 * b)) & 0xffu;
          v[2 * b] = __bfloat16_as_ushort(__float2bfloat16_rn(__fmul_rn(s_code[byte >> 4], am)));
          v[2 * b + 1] = __bfloat16_as_ushort(__float2bfloat16_rn(__fmul_rn(s_code[byte & 15u], am)));
        }
      } else {
        for (int e = 0; e < 8 && k + e < K; ++e) {
          const long f = flat + e;
          const long blk = f >> 6;
          const float am = __fadd_rn(__fmul_rn(s_ncode[qabsmax[blk]], nested_absmax[blk >> 8]), offset);
          const uint32_t byte = packed[f >> 1];
          const uint32_t nib = (f & 1) ? (byte & 15u) : (byte >> 4);
          v[e] = __bfloat16_as_ushort(__float2bfloat16_rn(__fmul_rn(s_code[nib], am)));
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) tile[r][g * 8 + e] = v[e];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += 256) {
    const int kk = i >> 3, g = i & 7;
    const int k = k0 + kk, n = n0 + g * 8;
    if (k >= K || n >= ld) continue;
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      o[e] = static_cast<uint32_t>(tile[g * 8 + 2 * e][kk]) | (static_cast<uint32_t>(tile[g * 8 + 2 * e + 1][kk]) << 16);
    *reinterpret_cast<uint4*>(out + static_cast<long>(k) * ld + n) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// Batched variant: every NF4 weight of a transformer block in ONE launch (7 per direction), each into its own workspace
// slot in the layout the CTA-pair GEMM reads ([N, ldk] for the forward, transposed [K, ldn] + the two transposed LoRA
// matrices for the backward).  Same arithmetic, hence the same bits, as nf4_dequant_kernel<bf16>.  One launch instead of
// seven removes six launch gaps / tails per block and direction (profiles/r1f: 168 dequant launches = 1.07 ms per step).
constexpr int kDqMaxItems = 8;
struct DequantItem {
  const uint8_t* packed;
  const uint8_t* qabsmax;
  const float* nested_absmax;
  const float* nested_code;
  const float* code;
  float offset;
  int N, K;
  __nv_bfloat16* out;              // [N, ld] or, transposed, [K, ld]
  long ld;
  int transposed;
  int tiles_k, cta_begin, num_tiles;
  const __nv_bfloat16* up;         // transposed only: lora_up [N,16] -> upT [16, ld] at out + K*ld
  const __nv_bfloat16* down;       //                  lora_down [16,K] pitch ldd -> downT [K,16] after upT
  long ldd;
};
struct DequantBatch {
  DequantItem items[kDqMaxItems];
  int n_items;
};

__global__ void __launch_bounds__(256)
nf4_dequant_batch_kernel(const __grid_constant__ DequantBatch bp) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  int ii = 0;
  while (ii + 1 < bp.n_items && static_cast<int>(blockIdx.x) >= bp.items[ii + 1].cta_begin) ++ii;
  const DequantItem& d = bp.items[ii];
  const int local = blockIdx.x - d.cta_begin;
  const int N = d.N, K = d.K;
  const long ld = d.ld;
  if (local >= d.num_tiles) {                       // the item's trailing blocks: LoRA transposes
    if (d.up == nullptr) return;
    __nv_bfloat16* upT = d.out + static_cast<long>(K) * ld;
    __nv_bfloat16* downT = upT + 16 * ld;
    const int t0 = (local - d.num_tiles) * blockDim.x + threadIdx.x;
    const int stride = 4 * blockDim.x;
    for (int i = t0; i < 16 * static_cast<int>(ld); i += stride) {
      const int r = i / static_cast<int>(ld), n = i % static_cast<int>(ld);
      upT[i] = n < N ? d.up[static_cast<long>(n) * 16 + r] : __float2bfloat16_rn(0.f);
    }
    for (int i = t0; i < K * 16; i += stride) {
      const int k = i >> 4, r = i & 15;
      downT[i] = d.down[static_cast<long>(r) * d.ldd + k];
    }
    return;
  }
  __shared__ float s_code[16];
  __shared__ float s_ncode[256];
  __shared__ unsigned short tile[64][66];
  if (threadIdx.x < 16) s_code[threadIdx.x] = d.code[threadIdx.x];
  s_ncode[threadIdx.x] = d.nested_code[threadIdx.x];
  __syncthreads();
  const int n0 = (local / d.tiles_k) * 64;
  const bool aligned = (K & 7) == 0;
  for (int kt = 0; kt < 4; ++kt) {
  const int k0 = (local % d.tiles_k) * 256 + kt * 64;
  if (k0 >= K) break;
  for (int i = threadIdx.x; i < 512; i += 256) {
    const int r = i >> 3, g = i & 7;
    const int n = n0 + r, k = k0 + g * 8;
    unsigned short v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0;
    if (n < N && k < K) {
      const long flat = static_cast<long>(n) * K + k;
      if (aligned) {                               // 8 codes = one 32-bit word inside one 64-block
        const long blk = flat >> 6;
        const float am = __fadd_rn(__fmul_rn(s_ncode[d.qabsmax[blk]], d.nested_absmax[blk >> 8]), d.offset);
        const uint32_t w = *reinterpret_cast<const uint32_t*>(d.packed + (flat >> 1));
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const uint32_t byte = (w >> (8 * b)) & 0xffu;
          v[2 * b] = __bfloat16_as_ushort(__float2bfloat16_rn(__fmul_rn(s_code[byte >> 4], am)));
          v[2 * b + 1] = __bfloat16_as_ushort(__float2bfloat16_rn(__fmul_rn(s_code[byte & 15u], am)));
        }
      } else {
        for (int e = 0; e < 8 && k + e < K; ++e) {
          const long f = flat + e;
          const long blk = f >> 6;
          const float am = __fadd_rn(__fmul_rn(s_ncode[d.qabsmax[blk]], d.nested_absmax[blk >> 8]), d.offset);
          const uint32_t byte = d.packed[f >> 1];
          const uint32_t nib = (f & 1) ? (byte & 15u) : (byte >> 4);
          v[e] = __bfloat16_as_ushort(__float2bfloat16_rn(__fmul_rn(s_code[nib], am)));
        }
      }
    }
    if (!d.transposed) {
      // row-major slot: columns k .. k+7 of row n (ld is a multiple of 8 >= K, so a full 16-byte store stays in the row)
      if (n < N && k < ld)
        *reinterpret_cast<uint4*>(d.out + static_cast<long>(n) * ld + k) =
            make_uint4(v[0] | (uint32_t(v[1]) << 16), v[2] | (uint32_t(v[3]) << 16), v[4] | (uint32_t(v[5]) << 16),
                       v[6] | (uint32_t(v[7]) << 16));
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) tile[r][g * 8 + e] = v[e];
    }
  }
  if (!d.transposed) continue;
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += 256) {
    const int kk = i >> 3, g = i & 7;
    const int k = k0 + kk, n = n0 + g * 8;
    if (k >= K || n >= ld) continue;
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      o[e] = static_cast<uint32_t>(tile[g * 8 + 2 * e][kk]) | (static_cast<uint32_t>(tile[g * 8 + 2 * e + 1][kk]) << 16);
    *reinterpret_cast<uint4*>(d.out + static_cast<long>(k) * ld + n) = make_uint4(o[0], o[1], o[2], o[3]);
  }
  __syncthreads();                                  // the tile buffer is reused by the next k-tile
  }
}

// Load-time repack for ragged in_features (K % 64 != 0, e.g. the SwiGLU hidden 2730 / 3413 of JiT-L / -H): bitsandbytes
// packs the FLATTENED [N,K] weight, so rows start at arbitrary nibbles and 64-blocks straddle rows.  The GEMM producers
// want 16-byte aligned rows: codes are re-packed row by row with pitch K_pad/2 bytes (K_pad = K rounded up to 64, padding
// = code 7 = 0.0).  The statistics stay indexed by the original flat position, so dequantised values are bit-identical.
__global__ void __launch_bounds__(256)
nf4_repack_rows_kernel(const uint8_t* __restrict__ packed, uint8_t* __restrict__ rows, int N, int K, int K_pad) {
  const long total = static_cast<long>(N) * (K_pad / 2);
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long n = i / (K_pad / 2);
    const int k = static_cast<int>(i % (K_pad / 2)) * 2;
    uint32_t c[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      if (k + e < K) {
        const long f = n * K + k + e;
        const uint8_t b = packed[f >> 1];
        c[e] = (f & 1) ? (b & 15u) : (b >> 4);
      } else {
        c[e] = 7u;
      }
    }
    rows[i] = static_cast<uint8_t>((c[0] << 4) | c[1]);
  }
}

// ---------------------------------------------------------------------------------------------- quantise
template <int kDt>
__device__ __forceinline__ float nf4_load(const void* w, long i) {
  if (kDt == kDtBf16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(w)[i]);
  if (kDt == kDtF16) return __half2float(reinterpret_cast<const __half*>(w)[i]);
  return reinterpret_cast<const float*>(w)[i];
}
// dQuantizeNF4: strict ">" against the midpoints of neighbouring code values
__device__ __forceinline__ uint32_t nf4_code_of(float x) {
  const float t[15] = {-0.8480964004993439f, -0.6106329262256622f, -0.4599952697753906f, -0.33967943489551544f,
                       -0.23460740596055984f, -0.13791173323988914f, -0.045525018125772476f, 0.03979014977812767f,
                       0.1202552504837513f, 0.2035212516784668f, 0.2920137718319893f, 0.3893125355243683f,
                       0.5016634166240692f, 0.6427869200706482f, 0.8614784181118011f};
  uint32_t c = 0;
#pragma unroll
  for (int i = 0; i < 15; ++i) c += x > t[i] ? 1u : 0u;
  return c;
}
// 8 lanes per 64-weight block: absmax, then 8 codes per lane -> 4 packed bytes.  n % 64 == 0.
template <int kDt>
__global__ void __launch_bounds__(256)
nf4_quantize_kernel(const void* __restrict__ w, uint8_t* __restrict__ packed, float* __restrict__ absmax, long nblocks) {
  const long g = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 3;
  const int sub = threadIdx.x & 7;
  const bool live = g < nblocks;
  float v[8];
  float m = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    v[e] = live ? nf4_load<kDt>(w, g * 64 + sub * 8 + e) : 0.f;
    m = fmaxf(m, fabsf(v[e]));
  }
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
  if (!live) return;
  if (sub == 0) absmax[g] = m;
  const float inv = __fdiv_rn(1.0f, m);
  uint32_t out = 0;
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const uint32_t hi = nf4_code_of(__fmul_rn(v[2 * b], inv)), lo = nf4_code_of(__fmul_rn(v[2 * b + 1], inv));
    out |= ((hi << 4) | lo) << (8 * b);
  }
  *reinterpret_cast<uint32_t*>(packed + g * 32 + sub * 4) = out;
}
// offset = mean(absmax) with a fixed summation order (single CTA, double accumulation)
__global__ void __launch_bounds__(1024) nf4_mean_kernel(const float* __restrict__ absmax, long nb, float* __restrict__ offset) {
  __shared__ double s[1024];
  double acc = 0.0;
  for (long i = threadIdx.x; i < nb; i += 1024) acc += static_cast<double>(absmax[i]);
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *offset = static_cast<float>(s[0] / static_cast<double>(nb));
}
// dQuantize<0>: binary search in the sorted 256-entry code, then the midpoint rule
__device__ __forceinline__ uint8_t nf4_nearest8(const float* code, float x) {
  int pivot = 127, upper_pivot = 255, lower_pivot = 0;
  float lower = -1.0f, upper = 1.0f, val = code[pivot];
  for (int i = 64; i > 0; i >>= 1) {
    if (x > val) {
      lower_pivot = pivot;
      lower = val;
      pivot += i;
    } else {
      upper_pivot = pivot;
      upper = val;
      pivot -= i;
    }
    val = code[pivot];
  }
  if (upper_pivot == 255) upper = code[upper_pivot];
  if (lower_pivot == 0) lower = code[lower_pivot];
  if (x > val) return x > (upper + val) * 0.5f ? upper_pivot : pivot;
  return x < (lower + val) * 0.5f ? lower_pivot : pivot;
}
// one CTA of 256 threads per nested block of 256 statistics
__global__ void __launch_bounds__(256)
nf4_nested_quantize_kernel(const float* __restrict__ absmax, const float* __restrict__ offset_p,
                           const float* __restrict__ nested_code, uint8_t* __restrict__ q, float* __restrict__ nested_absmax,
                           long nb) {
  __shared__ float s_code[256];
  __shared__ float s_red[256];
  s_code[threadIdx.x] = nested_code[threadIdx.x];
  const long i = static_cast<long>(blockIdx.x) * 256 + threadIdx.x;
  const float v = i < nb ? __fsub_rn(absmax[i], *offset_p) : 0.f;
  s_red[threadIdx.x] = fabsf(v);
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s_red[threadIdx.x] = fmaxf(s_red[threadIdx.x], s_red[threadIdx.x + o]);
    __syncthreads();
  }
  const float m = s_red[0];
  if (threadIdx.x == 0) nested_absmax[blockIdx.x] = m;
  if (i < nb) q[i] = nf4_nearest8(s_code, __fmul_rn(v, __fdiv_rn(1.0f, m)));
}

}  // namespace vpt
