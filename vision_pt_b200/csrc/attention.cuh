// Attention forward / backward for sm_100a (tcgen05 + TMEM + TMA), head_dim 64, non-causal, per-sample valid key
// length (the reference's key-padding mask on trailing context tokens: src/models/jit/denoiser.py:375-381,
// src/modules/attention.py:98-129 with mask [B,1,1,Lk] -> here `seqlens_k[b]`).
//
// q/k/v/o are addressed as 4-D tensors (hd, L, H, B) with arbitrary strides for L, H, B (hd contiguous), so both the
// reference's [B,H,L,hd] layout and the token-major [B,L,H,hd] layout the fused block uses are accepted without copies.
//
// Forward: CTA = (128-query tile, head, sample).  S = Q K^T lives in TMEM (128 cols), one softmax thread per query
// row, P is written as a bf16 A-operand tile in shared memory, O accumulates in TMEM (64 cols) with online rescale.
// Backward: CTA = (128-key tile, head, sample), loops over query tiles.  S^T = K Q^T and dP^T = V dO^T in TMEM,
// one thread per key row builds P^T and dS^T (bf16, shared memory); dV += P^T dO, dK += dS^T Q, dQ = dS K.
// dQ tiles are reduced across key tiles with fp32 vector reductions into a zero-initialised fp32 buffer.
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
  float* lse2;                 // [B, H, Lq]  log2-domain logsumexp of scale*S
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
    if (q < p.Lq) p.lse2[(static_cast<long>(b) * p.H + h) * p.Lq + q] = (l_run > 0.f) ? m_run + log2f(l_run) : -INFINITY;
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

// ---------------------------------------------------------------------------------------------- backward
// delta[b,h,q] = sum_d dO[b,q,h,d] * O[b,q,h,d]   (fp32)
struct AttnDeltaParams {
  int B, H, Lq;
  const __nv_bfloat16 *o, *d_o;
  long o_sb, o_sl, o_sh, do_sb, do_sl, do_sh;
  float* delta;                // [B, H, Lq]
};
__global__ void attn_bwd_delta_kernel(const AttnDeltaParams p) {
  // one 8-lane group per (b, h, q): 8 lanes x 8 elements = 64
  const long gid = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 3;
  const int sub = threadIdx.x & 7;
  const long total = static_cast<long>(p.B) * p.H * p.Lq;
  float acc = 0.f;
  if (gid < total) {
    const int q = gid % p.Lq;
    const int h = (gid / p.Lq) % p.H;
    const int b = gid / (static_cast<long>(p.Lq) * p.H);
    const uint4 a = *reinterpret_cast<const uint4*>(p.o + b * p.o_sb + static_cast<long>(q) * p.o_sl + h * p.o_sh + sub * 8);
    const uint4 g = *reinterpret_cast<const uint4*>(p.d_o + b * p.do_sb + static_cast<long>(q) * p.do_sl + h * p.do_sh + sub * 8);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) acc += bf16lo(aw[i]) * bf16lo(gw[i]) + bf16hi(aw[i]) * bf16hi(gw[i]);
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (gid < total && sub == 0) p.delta[gid] = acc;
}

struct AttnBwdParams {
  int B, H, Lq, Lk;
  const int* seqlens_k;
  float scale_log2, scale;
  const float* lse2;           // [B, H, Lq]
  const float* delta;          // [B, H, Lq]
  float* dq;                   // fp32, zero-initialised by the caller; element strides below
  long dq_sb, dq_sl, dq_sh;
  __nv_bfloat16 *dk, *dv;
  long dk_sb, dk_sl, dk_sh, dv_sb, dv_sl, dv_sh;
};

struct AttnBwdSmem {
  static constexpr int kK = 0;
  static constexpr int kV = 16384;
  static constexpr int kQ = 32768;                 // 2 stages
  static constexpr int kDO = kQ + 2 * 16384;       // 2 stages
  static constexpr int kPT = kDO + 2 * 16384;      // [128 keys x 128 queries] bf16, 2 K-atoms
  static constexpr int kDST = kPT + 32768;
  static constexpr int kStats = kDST + 32768;      // 2 stages x (lse2[128], delta[128]) fp32
  static constexpr int kBars = kStats + 2 * 1024;
  static constexpr int kNumBars = 9;
  static constexpr int kTmemSlot = kBars + kNumBars * 8;
  static constexpr int kTotal = kTmemSlot + 16 + 1024;
};

// tmQ / tmDO / tmK / tmV: 4-D maps (hd, L, H, B), box {64, 128, 1, 1}, SWIZZLE_128B
__global__ void __launch_bounds__(192, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                const AttnBwdParams p) {
  using S = AttnBwdSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBars);
  uint64_t* kv_full = bars;
  uint64_t* qdo_full = bars + 1;   // [2]
  uint64_t* qdo_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* mma2_done = bars + 7;
  uint64_t* stats_free = bars + 8; // unused slot kept for alignment of the layout
  (void)stats_free;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::kTmemSlot);
  float* s_stats = reinterpret_cast<float*>(smem + S::kStats);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * kAttnTile, h = blockIdx.y, b = blockIdx.z;
  int klen = p.seqlens_k ? p.seqlens_k[b] : p.Lk;
  klen = klen < p.Lk ? klen : p.Lk;
  const int nq = (p.Lq + kAttnTile - 1) / kAttnTile;
  const bool active = k0 < klen;      // a key tile that is entirely padding only writes zero gradients

  if (threadIdx.x == 0) {
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&qdo_full[s], 1);
      mbar_init(&qdo_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(mma2_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tST = tmem_base, tDPT = tmem_base + 128, tDV = tmem_base + 256, tDK = tmem_base + 320,
                 tDQ = tmem_base + 384;

  if (warp == 0) {
    if (lane == 0 && active) {
      mbar_arrive_expect_tx(kv_full, 32768);
      tma_load_4d(&tmK, kv_full, smem + S::kK, 0, k0, h, b);
      tma_load_4d(&tmV, kv_full, smem + S::kV, 0, k0, h, b);
      for (int i = 0; i < nq; ++i) {
        const int s = i & 1;
        mbar_wait(&qdo_empty[s], ((i >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&qdo_full[s], 32768);
        tma_load_4d(&tmQ, &qdo_full[s], smem + S::kQ + s * 16384, 0, i * kAttnTile, h, b);
        tma_load_4d(&tmDO, &qdo_full[s], smem + S::kDO + s * 16384, 0, i * kAttnTile, h, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && active) {
      constexpr uint32_t kIdKK = umma_idesc_bf16(128, 128, 0, 0);  // S^T = K Q^T, dP^T = V dO^T (both K-major)
      constexpr uint32_t kIdKM = umma_idesc_bf16(128, 64, 0, 1);   // dV += P^T dO, dK += dS^T Q (B MN-major)
      constexpr uint32_t kIdMM = umma_idesc_bf16(128, 64, 1, 1);   // dQ = dS K (A = dS^T tile viewed MN-major, B MN-major)
      const uint32_t sK = smem_u32(smem + S::kK), sV = smem_u32(smem + S::kV);
      const uint32_t sPT = smem_u32(smem + S::kPT), sDST = smem_u32(smem + S::kDST);
      mbar_wait(kv_full, 0);
      for (int i = 0; i < nq; ++i) {
        const int s = i & 1;
        const uint32_t sQ = smem_u32(smem + S::kQ + s * 16384), sDO = smem_u32(smem + S::kDO + s * 16384);
        mbar_wait(&qdo_full[s], (i >> 1) & 1);
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss(tST, umma_smem_desc(sK + k * 32, 16, 1024, kLayoutSW128),
                  umma_smem_desc(sQ + k * 32, 16, 1024, kLayoutSW128), kIdKK, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss(tDPT, umma_smem_desc(sV + k * 32, 16, 1024, kLayoutSW128),
                  umma_smem_desc(sDO + k * 32, 16, 1024, kLayoutSW128), kIdKK, k != 0);
        umma_commit(s_full);
        mbar_wait(p_full, i & 1);
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < 8; ++k) {   // reduction over the 128 queries of this tile
          const uint64_t a_pt = umma_smem_desc(sPT + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024, kLayoutSW128);
          const uint64_t a_ds = umma_smem_desc(sDST + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024, kLayoutSW128);
          umma_ss(tDV, a_pt, umma_smem_desc(sDO + k * 2048, 8192, 1024, kLayoutSW128), kIdKM, (i | k) != 0);
          umma_ss(tDK, a_ds, umma_smem_desc(sQ + k * 2048, 8192, 1024, kLayoutSW128), kIdKM, (i | k) != 0);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)     // reduction over the 128 keys of this CTA
          umma_ss(tDQ, umma_smem_desc(sDST + k * 2048, 16384, 1024, kLayoutSW128),
                  umma_smem_desc(sK + k * 2048, 8192, 1024, kLayoutSW128), kIdMM, k != 0);
        umma_commit(&qdo_empty[s]);
        umma_commit(mma2_done);
      }
    }
  } else {
    const int qd = warp & 3;
    const int row = qd * 32 + lane;                  // key row of S^T / query row of dQ
    const int st_tid = threadIdx.x - 64;             // 0..127
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const int key = k0 + row;
    const bool key_ok = key < klen;
    const int rin = row & 7;
    const uint32_t pt_row = smem_u32(smem + S::kPT) + row * 128;
    const uint32_t dst_row = smem_u32(smem + S::kDST) + row * 128;
    if (active) {
      for (int i = 0; i < nq; ++i) {
        const int qbase = i * kAttnTile;
        float* st = s_stats + (i & 1) * 256;
        {
          const int q = qbase + st_tid;
          const long sidx = (static_cast<long>(b) * p.H + h) * p.Lq + q;
          st[st_tid] = q < p.Lq ? p.lse2[sidx] : INFINITY;     // +inf -> p = exp2(-inf) = 0 for padded queries
          st[128 + st_tid] = q < p.Lq ? p.delta[sidx] : 0.f;
        }
        named_bar_sync(1, 128);
        mbar_wait(s_full, i & 1);
        tc_fence_after_sync();
        if (i > 0) {
          // previous pair's MMAs are done: P^T / dS^T may be overwritten, dQ(i-1) can be drained
          mbar_wait(mma2_done, (i - 1) & 1);
          tc_fence_after_sync();
          const int q = qbase - kAttnTile + row;
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld32(tDQ + lane_off + c * 32, v);
            tmem_wait_ld();
            if (q < p.Lq) {
              float* dst = p.dq + b * p.dq_sb + static_cast<long>(q) * p.dq_sl + h * p.dq_sh + c * 32;
#pragma unroll
              for (int g = 0; g < 8; ++g)
                red_add_v4_f32(dst + g * 4, __uint_as_float(v[g * 4]), __uint_as_float(v[g * 4 + 1]),
                               __uint_as_float(v[g * 4 + 2]), __uint_as_float(v[g * 4 + 3]));
            }
          }
        }
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t sv[32], dv[32];
          tmem_ld32(tST + lane_off + c * 32, sv);
          tmem_ld32(tDPT + lane_off + c * 32, dv);
          tmem_wait_ld();
          uint32_t pk[16], dk[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int c0 = c * 32 + 2 * e;
            float p0 = fast_exp2(__uint_as_float(sv[2 * e]) * p.scale_log2 - st[c0]);
            float p1 = fast_exp2(__uint_as_float(sv[2 * e + 1]) * p.scale_log2 - st[c0 + 1]);
            if (!key_ok) {
              p0 = 0.f;
              p1 = 0.f;
            }
            const float d0 = p0 * (__uint_as_float(dv[2 * e]) - st[128 + c0]) * p.scale;
            const float d1 = p1 * (__uint_as_float(dv[2 * e + 1]) - st[128 + c0 + 1]) * p.scale;
            pk[e] = pack_bf16x2(p0, p1);
            dk[e] = pack_bf16x2(d0, d1);
          }
          const uint32_t off = (c >> 1) * 16384;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t chunk = static_cast<uint32_t>(((c & 1) * 4 + g) ^ rin) * 16;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pt_row + off + chunk), "r"(pk[4 * g]),
                         "r"(pk[4 * g + 1]), "r"(pk[4 * g + 2]), "r"(pk[4 * g + 3])
                         : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst_row + off + chunk), "r"(dk[4 * g]),
                         "r"(dk[4 * g + 1]), "r"(dk[4 * g + 2]), "r"(dk[4 * g + 3])
                         : "memory");
          }
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        mbar_arrive(p_full);
      }
      // last dQ tile, then dK / dV
      mbar_wait(mma2_done, (nq - 1) & 1);
      tc_fence_after_sync();
      {
        const int q = (nq - 1) * kAttnTile + row;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(tDQ + lane_off + c * 32, v);
          tmem_wait_ld();
          if (q < p.Lq) {
            float* dst = p.dq + b * p.dq_sb + static_cast<long>(q) * p.dq_sl + h * p.dq_sh + c * 32;
#pragma unroll
            for (int g = 0; g < 8; ++g)
              red_add_v4_f32(dst + g * 4, __uint_as_float(v[g * 4]), __uint_as_float(v[g * 4 + 1]),
                             __uint_as_float(v[g * 4 + 2]), __uint_as_float(v[g * 4 + 3]));
          }
        }
      }
    }
#pragma unroll 1
    for (int t = 0; t < 2; ++t) {
      __nv_bfloat16* base = t == 0 ? p.dv + b * p.dv_sb + static_cast<long>(key) * p.dv_sl + h * p.dv_sh
                                   : p.dk + b * p.dk_sb + static_cast<long>(key) * p.dk_sl + h * p.dk_sh;
      const uint32_t tsrc = t == 0 ? tDV : tDK;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        if (active) {
          tmem_ld32(tsrc + lane_off + c * 32, v);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = 0;
        }
        if (key < p.Lk) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 o4;
            o4.x = pack_bf16x2(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]));
            o4.y = pack_bf16x2(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3]));
            o4.z = pack_bf16x2(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
            o4.w = pack_bf16x2(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]));
            *reinterpret_cast<uint4*>(base + c * 32 + g * 8) = o4;
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace vpt
