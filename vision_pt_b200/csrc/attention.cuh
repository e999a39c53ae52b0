// Attention forward / backward for sm_100a (tcgen05 + TMEM + TMA), head_dim 64, non-causal, per-sample valid key
// length (the reference's key-padding mask on trailing context tokens: src/models/jit/denoiser.py:375-381,
// src/modules/attention.py:98-129 with mask [B,1,1,Lk] -> here `seqlens_k[b]`).
//
// q/k/v/o are addressed as 4-D tensors (hd, L, H, B) with arbitrary strides for L, H, B (hd contiguous), so both the
// reference's [B,H,L,hd] layout and the token-major [B,L,H,hd] layout the fused block uses are accepted without copies.
//
// Forward: CTA = (128-query tile, head, sample).  S = Q K^T lives in TMEM (128 cols), one softmax thread per query
// row, P is written as a bf16 A-operand tile in shared memory, O accumulates in TMEM (64 cols) with online rescale.
// Backward: attention_bwd.cuh.
#pragma once
#include "sm100.cuh"

namespace vpt {

constexpr int kAttnHD = 64;
constexpr int kAttnTile = 128;

struct AttnFwdParams {
  int B, H, Lq, Lk;
  const int* seqlens_k;        // [B] or nullptr (= Lk)
  float scale_log2;            // softmax scale * log2(e)
  __nv_bfloat16* o;            // element strides below
  long o_sb, o_sl, o_sh;
  float* lse2;                 // [B, H, Lq rounded up to 128]  log2-domain logsumexp of scale*S; +inf in the padding
};

struct AttnFwdSmem {
  static constexpr int kQ = 0;
  static constexpr int kK = 16384;            // 2 stages
  static constexpr int kV = kK + 2 * 16384;   // 2 stages
  static constexpr int kP = kV + 2 * 16384;   // 2 K-atoms of [128 x 64]
  static constexpr int kBars = kP + 32768;
  static constexpr int kNumBars = 8;
  static constexpr int kTmemSlot = kBars + kNumBars * 8;
  static constexpr int kTotal = kTmemSlot + 16;   // no alignment slack: two CTAs must fit in one SM's 228 KB
};

__global__ void __launch_bounds__(192, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnFwdParams p) {
  using S = AttnFwdSmem;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // 128B-swizzled tiles need a 1024B-aligned base
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBars);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::kTmemSlot);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kAttnTile, h = blockIdx.y, b = blockIdx.z;
  pdl_wait();                                    // seqlens_k may come from the kernel just before
  int klen = p.seqlens_k ? p.seqlens_k[b] : p.Lk;
  klen = klen < p.Lk ? klen : p.Lk;
  const int nblk = (klen + kAttnTile - 1) / kAttnTile;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tO = tmem_base + 128;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      mbar_arrive_expect_tx(q_full, 16384);
      tma_load_4d(&tmQ, q_full, smem + S::kQ, 0, q0, h, b);
      for (int j = 0; j < nblk; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 32768);
        tma_load_4d(&tmK, &kv_full[s], smem + S::kK + s * 16384, 0, j * kAttnTile, h, b);
        tma_load_4d(&tmV, &kv_full[s], smem + S::kV + s * 16384, 0, j * kAttnTile, h, b);
      }
    }
  } else if (warp == 1) {
    // converged warp, one elected lane issues (operands stay in uniform registers)
    constexpr uint32_t kIdS = umma_idesc_bf16(128, 128, 0, 0);   // S = Q K^T   (both K-major)
    constexpr uint32_t kIdO = umma_idesc_bf16(128, 64, 0, 1);    // O += P V    (V is MN-major)
    const uint32_t smem_base = smem_u32(smem);
    const uint64_t dK_ = umma_smem_desc(0, 16, 1024, kLayoutSW128);
    const uint64_t dMN = umma_smem_desc(0, 8192, 1024, kLayoutSW128);
    const uint64_t qd = dK_ + ((smem_base + S::kQ) >> 4), pd = dK_ + ((smem_base + S::kP) >> 4);
    mbar_wait(q_full, 0);
    for (int j = 0; j < nblk; ++j) {
      const int s = j & 1;
      const uint64_t kd = dK_ + ((smem_base + S::kK + s * 16384) >> 4), vd = dMN + ((smem_base + S::kV + s * 16384) >> 4);
      mbar_wait(&kv_full[s], (j >> 1) & 1);
      tc_fence_after_sync();
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tS, qd + 2 * k, kd + 2 * k, kIdS, k != 0);
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_full, j & 1);
      tc_fence_after_sync();
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_ss(tO, pd + (k >> 2) * 1024 + (k & 3) * 2, vd + k * 128, kIdO, (j | k) != 0);
        umma_commit(&kv_empty[s]);
        umma_commit(o_full);
      }
      __syncwarp();
    }
  } else {
    // softmax warps 2..5: TMEM lane quarter = warp % 4
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t p_row = smem_u32(smem + S::kP) + row * 128;
    const int rin = row & 7;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nblk; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after_sync();
      const int kbase = j * kAttnTile;
      // pass 1: block max (log2 domain)
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tS + lane_off + c * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float s = (kbase + c * 32 + i < klen) ? __uint_as_float(v[i]) * p.scale_log2 : -INFINITY;
          mx = fmaxf(mx, s);
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = fast_exp2(m_run - m_use);   // m_run = -inf -> 0
      // P(j-1) must have been consumed, and O must be final for block j-1, before they are touched
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after_sync();
      }
      // pass 2: p = exp2(s - m), row sum, bf16 P tile
      float psum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tS + lane_off + c * 32, v);
        tmem_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int key = kbase + c * 32 + 2 * i;
          const float p0 = key < klen ? fast_exp2(__uint_as_float(v[2 * i]) * p.scale_log2 - m_use) : 0.f;
          const float p1 = key + 1 < klen ? fast_exp2(__uint_as_float(v[2 * i + 1]) * p.scale_log2 - m_use) : 0.f;
          psum += p0 + p1;
          pk[i] = pack_bf16x2(p0, p1);
        }
        // keys c*32 .. c*32+31 -> K-atom c/2, 16B chunks (c%2)*4 .. +3 of this row
        const uint32_t base = p_row + (c >> 1) * 16384;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t chunk = static_cast<uint32_t>(((c & 1) * 4 + g) ^ rin);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + chunk * 16), "r"(pk[4 * g]),
                       "r"(pk[4 * g + 1]), "r"(pk[4 * g + 2]), "r"(pk[4 * g + 3])
                       : "memory");
        }
      }
      l_run = l_run * alpha + psum;
      m_run = m_new;
      if (j > 0 && !__all_sync(0xffffffffu, alpha == 1.f)) {
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(tO + lane_off + c * 32, v);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st32(tO + lane_off + c * 32, v);
        }
        tmem_wait_st();
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(p_full);
    }
    // epilogue
    if (nblk > 0) {
      mbar_wait(o_full, (nblk - 1) & 1);
      tc_fence_after_sync();
    }
    const int q = q0 + row;
    const float inv_l = l_run > 0.f ? 1.f / l_run : 0.f;
    __nv_bfloat16* orow = p.o + b * p.o_sb + static_cast<long>(q) * p.o_sl + h * p.o_sh;
    // row pitch = Lq rounded up to 128 so that the backward can fetch a tile's 128 statistics with one bulk copy;
    // padded queries get +inf (their recomputed probabilities are exp2(-inf) = 0)
    p.lse2[(static_cast<long>(b) * p.H + h) * (static_cast<long>(gridDim.x) * kAttnTile) + q] =
        q < p.Lq ? ((l_run > 0.f) ? m_run + log2f(l_run) : -INFINITY) : INFINITY;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      if (nblk > 0) {                       // CTA-uniform: the .sync.aligned TMEM load stays warp-convergent
        tmem_ld32(tO + lane_off + c * 32, v);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0;
      }
      if (q < p.Lq) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o4;
          o4.x = pack_bf16x2(__uint_as_float(v[g * 8 + 0]) * inv_l, __uint_as_float(v[g * 8 + 1]) * inv_l);
          o4.y = pack_bf16x2(__uint_as_float(v[g * 8 + 2]) * inv_l, __uint_as_float(v[g * 8 + 3]) * inv_l);
          o4.z = pack_bf16x2(__uint_as_float(v[g * 8 + 4]) * inv_l, __uint_as_float(v[g * 8 + 5]) * inv_l);
          o4.w = pack_bf16x2(__uint_as_float(v[g * 8 + 6]) * inv_l, __uint_as_float(v[g * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = o4;
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

// ---------------------------------------------------------------------------------------------- backward pre-pass
// delta[b,h,q] = scale * sum_d dO[b,q,h,d] * O[b,q,h,d]   (fp32, row pitch Lq rounded up to 128, zero in the padding)
struct AttnDeltaParams {
  int B, H, Lq, Lq_pad;
  float scale;
  const __nv_bfloat16 *o, *d_o;
  long o_sb, o_sl, o_sh, do_sb, do_sl, do_sh;
  float* delta;                // [B, H, Lq_pad]
};
__global__ void attn_bwd_delta_kernel(const AttnDeltaParams p) {
  pdl_launch_dependents();
  pdl_wait();
  // one 8-lane group per (b, h, q): 8 lanes x 8 elements = 64
  const long gid = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 3;
  const int sub = threadIdx.x & 7;
  const long total = static_cast<long>(p.B) * p.H * p.Lq_pad;
  float acc = 0.f;
  if (gid < total) {
    const int q = gid % p.Lq_pad;
    const int h = (gid / p.Lq_pad) % p.H;
    const int b = gid / (static_cast<long>(p.Lq_pad) * p.H);
    if (q < p.Lq) {
      const uint4 a = *reinterpret_cast<const uint4*>(p.o + b * p.o_sb + static_cast<long>(q) * p.o_sl + h * p.o_sh + sub * 8);
      const uint4 g = *reinterpret_cast<const uint4*>(p.d_o + b * p.do_sb + static_cast<long>(q) * p.do_sl + h * p.do_sh + sub * 8);
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) acc += bf16lo(aw[i]) * bf16lo(gw[i]) + bf16hi(aw[i]) * bf16hi(gw[i]);
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (gid < total && sub == 0) p.delta[gid] = acc * p.scale;
}

}  // namespace vpt
