// Attention backward for sm_100a (tcgen05 + TMEM + TMA), head_dim 64, non-causal, per-sample valid key length.
// Autograd of scaled_dot_product_attention as Attention.forward calls it (src/models/jit/denoiser.py:351-397,
// src/modules/attention.py:98-129 of the reference).
//
// Persistent CTAs (one per SM, 16 warps) walk work items (128-key tile, head, sample) and, inside an item, the query
// tiles.  Every 128-query tile is two 64-query sub-tiles A and B that ping-pong between the tensor core and two softmax
// warpgroups, so the exp / dS arithmetic of one sub-tile hides behind the MMAs of the other:
//
//   tensor core    S^T_X = K Q_X^T, dP^T_X = V dO_X^T   (TMEM, 64 columns each)                          X = A, B
//   warpgroup X    P^T_X = exp2(S^T_X c - lse), dS^T_X = P^T_X (dP^T_X - delta) scale  -> bf16, shared memory
//   tensor core    dV += P^T_X dO_X, dK += dS^T_X Q_X   (TMEM accumulators over the whole item),  dQ = dS K (per tile)
//   drain warps    dQ tile -> fp32 shared-memory slabs -> TMA reduce-add (dQ is summed across key tiles in global
//                  memory; per-thread red.global measured ~10k cycles per tile on the LSU, profiles/r1e_attn_bwd.txt)
//
// Warp roles: 0 TMA producer (K/V double buffered across items; Q/dO and the tile's lse / delta*scale rows in two
// stages), 1 MMA issuer, 2 TMEM allocator, 4-7 warpgroup A, 8-11 warpgroup B, 12-15 drain (dQ per tile, dV / dK per item).
// The last query tile is trimmed to a multiple of 16 queries (UMMA N / K granularity).
#pragma once
#include "sm100.cuh"

namespace vpt {

struct AttnBwd2Params {
  int B, H, Lq, Lk;
  const int* seqlens_k;
  float scale_log2, scale;
  const float* lse2;           // [B, H, nq * 128]  (+inf in the padding)
  const float* delta;          // [B, H, nq * 128]  delta * scale (0 in the padding)
  float* dq;                   // fp32, zero-initialised by the caller; element strides below
  long dq_sb, dq_sl, dq_sh;
  int nk, nq, num_items;       // key tiles, query tiles, B * H * nk
};

struct AttnBwd2Smem {
  static constexpr int kK = 0;                       // 2 buffers x 16 KB
  static constexpr int kV = 32768;                   // 2 buffers x 16 KB
  static constexpr int kQ = 65536;                   // 2 stages
  static constexpr int kDO = kQ + 2 * 16384;         // 2 stages
  static constexpr int kPT = kDO + 2 * 16384;        // [128 keys x 128 queries] bf16 = 2 K-atoms (A, B)
  static constexpr int kDST = kPT + 32768;
  static constexpr int kStats = kDST + 32768;        // 2 stages x (lse2[128], delta*scale[128]) fp32
  static constexpr int kDQ = kStats + 2 * 1024;      // 4 drain warps x 2 slabs of [32 queries x 32 fp32], 128B-swizzled
  static constexpr int kBars = kDQ + 4 * 2 * 4096;
  // kv_full[2] kv_empty[2] qdo_full[2] qdo_empty[2] s_full[2] p_full[2] mma2_done dq_free dkv_free
  static constexpr int kNumBars = 15;
  static constexpr int kTmemSlot = kBars + kNumBars * 8;
  static constexpr int kTotal = kTmemSlot + 16;      // no alignment slack: the dynamic segment starts 1024B-aligned
};

__device__ __forceinline__ int attn_round16(int x) { return (x + 15) & ~15; }

// tmQ / tmDO / tmK / tmV: 4-D bf16 maps (hd, L, H, B), box {64, 128, 1, 1}, SWIZZLE_128B
// tmDQ: 4-D fp32 map (hd, Lq, H, B) of the dQ accumulator, box {32, 32, 1, 1}, SWIZZLE_128B
// tmDK / tmDV: 4-D bf16 maps (hd, Lk, H, B) of the outputs, box {64, 32, 1, 1}, SWIZZLE_128B
__global__ void __launch_bounds__(512, 1)
attn_bwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                 const __grid_constant__ CUtensorMap tmDQ, const __grid_constant__ CUtensorMap tmDK,
                 const __grid_constant__ CUtensorMap tmDV, const AttnBwd2Params p) {
  using S = AttnBwd2Smem;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // 128B-swizzled tiles need a 1024B-aligned base
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBars);
  uint64_t* kv_full = bars;          // [2]
  uint64_t* kv_empty = bars + 2;     // [2]
  uint64_t* qdo_full = bars + 4;     // [2]  TMA bytes + the statistics warp
  uint64_t* qdo_empty = bars + 6;    // [2]
  uint64_t* s_full = bars + 8;       // [2]  A, B
  uint64_t* p_full = bars + 10;      // [2]  A, B
  uint64_t* mma2_done = bars + 12;
  uint64_t* dq_free = bars + 13;
  uint64_t* dkv_free = bars + 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::kTmemSlot);
  float* s_stats = reinterpret_cast<float*>(smem + S::kStats);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq = p.nq;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
      mbar_init(&qdo_full[s], 1);
      mbar_init(&qdo_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 128);
    }
    mbar_init(mma2_done, 1);
    mbar_init(dq_free, 4);
    mbar_init(dkv_free, 4);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();
  // columns: S^T_A 0, S^T_B 64, dP^T_A 128, dP^T_B 192, dV 256, dK 320, dQ 384
  const uint32_t tST = tmem_base, tDPT = tmem_base + 128, tDV = tmem_base + 256, tDK = tmem_base + 320,
                 tDQ = tmem_base + 384;

  // queries of tile i that sub-tile X covers, rounded up to the UMMA granularity (0 = sub-tile absent)
  auto sub_n = [&](int i, int X) {
    const int valid = min(128, p.Lq - i * 128) - X * 64;
    return valid <= 0 ? 0 : min(64, attn_round16(valid));
  };

  if (warp == 0) {
    // ============================================================ TMA producer
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      tma_prefetch_desc(&tmDO);
      uint32_t f = 0, itn = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++itn) {
        const int kt = item % p.nk, h = (item / p.nk) % p.H, b = item / (p.nk * p.H);
        const uint32_t kb = itn & 1;
        mbar_wait(&kv_empty[kb], ((itn >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[kb], 32768);
        tma_load_4d(&tmK, &kv_full[kb], smem + S::kK + kb * 16384, 0, kt * 128, h, b);
        tma_load_4d(&tmV, &kv_full[kb], smem + S::kV + kb * 16384, 0, kt * 128, h, b);
        for (int i = 0; i < nq; ++i, ++f) {
          const uint32_t s = f & 1;
          mbar_wait(&qdo_empty[s], ((f >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&qdo_full[s], 32768 + 1024);
          tma_load_4d(&tmQ, &qdo_full[s], smem + S::kQ + s * 16384, 0, i * 128, h, b);
          tma_load_4d(&tmDO, &qdo_full[s], smem + S::kDO + s * 16384, 0, i * 128, h, b);
          const long srow = (static_cast<long>(b) * p.H + h) * (static_cast<long>(nq) * 128) + i * 128;
          bulk_load_1d(smem + S::kStats + s * 1024, p.lse2 + srow, 512, &qdo_full[s]);
          bulk_load_1d(smem + S::kStats + s * 1024 + 512, p.delta + srow, 512, &qdo_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer (converged warp, one elected lane issues)
    {
      constexpr uint32_t kIdKM64 = umma_idesc_bf16(128, 64, 0, 1);   // dV += P^T dO, dK += dS^T Q (B MN-major)
      constexpr uint32_t kIdMM = umma_idesc_bf16(128, 64, 1, 1);     // dQ = dS K (A = dS^T viewed MN-major, B MN-major)
      const uint32_t smem_base = smem_u32(smem);
      const uint64_t dK_ = umma_smem_desc(0, 16, 1024, kLayoutSW128);       // K-major operand, + (addr >> 4)
      const uint64_t dMN = umma_smem_desc(0, 8192, 1024, kLayoutSW128);     // MN-major [64-wide blocks 8 KB apart]
      const uint64_t dMNq = umma_smem_desc(0, 16384, 1024, kLayoutSW128);   // dS^T viewed MN-major (query atoms 16 KB apart)
      const uint64_t aPT = dK_ + ((smem_base + S::kPT) >> 4), aDST = dK_ + ((smem_base + S::kDST) >> 4);
      const uint64_t aDSTq = dMNq + ((smem_base + S::kDST) >> 4);
      const int my_items = (p.num_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
      const uint32_t total = static_cast<uint32_t>(my_items) * nq;
      const int n_last[2] = {sub_n(nq - 1, 0), sub_n(nq - 1, 1)};
      // S^T_X / dP^T_X of the flat tile with item counter itn, query tile i, Q/dO stage s (phase parity sp)
      auto mma1 = [&](uint32_t itn, int i, uint32_t s, uint32_t sp, int X) {
        const int n = i == nq - 1 ? n_last[X] : 64;
        if (n == 0) return;
        const uint32_t kb = itn & 1;
        mbar_wait(&kv_full[kb], (itn >> 1) & 1);
        mbar_wait(&qdo_full[s], sp);
        tc_fence_after_sync();
        const uint64_t kd = dK_ + ((smem_base + S::kK + kb * 16384) >> 4), vd = dK_ + ((smem_base + S::kV + kb * 16384) >> 4);
        const uint64_t qd_ = dK_ + ((smem_base + S::kQ + s * 16384 + X * 8192) >> 4);
        const uint64_t dod = dK_ + ((smem_base + S::kDO + s * 16384 + X * 8192) >> 4);
        const uint32_t idesc = umma_idesc_bf16(128, n, 0, 0);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tST + X * 64, kd + 2 * k, qd_ + 2 * k, idesc, k != 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tDPT + X * 64, vd + 2 * k, dod + 2 * k, idesc, k != 0);
          umma_commit(&s_full[X]);
        }
        __syncwarp();
      };
      uint32_t cnt[2] = {0, 0};                      // sub-tiles of each kind handed to the warpgroups so far
      if (total > 0) {
        mma1(0, 0, 0, 0, 0);
        mma1(0, 0, 0, 0, 1);
      }
      uint32_t itn = 0;
      int i = 0;
      for (uint32_t g = 0; g < total; ++g) {
        const uint32_t s = g & 1;
        const uint32_t kb = itn & 1;
        const uint64_t qmn = dMN + ((smem_base + S::kQ + s * 16384) >> 4), domn = dMN + ((smem_base + S::kDO + s * 16384) >> 4);
        const uint64_t kmn = dMN + ((smem_base + S::kK + kb * 16384) >> 4);
        // next flat tile
        const bool has_next = g + 1 < total;
        const int ni = i + 1 == nq ? 0 : i + 1;
        const uint32_t nitn = i + 1 == nq ? itn + 1 : itn;
#pragma unroll 1
        for (int X = 0; X < 2; ++X) {
          const int n = i == nq - 1 ? n_last[X] : 64;
          if (n > 0) {
            mbar_wait(&p_full[X], cnt[X] & 1);
            ++cnt[X];
            if (i == 0 && X == 0 && itn > 0) mbar_wait(dkv_free, (itn - 1) & 1);   // previous item's dV / dK are out
            tc_fence_after_sync();
            if (elect_one_sync()) {
              for (int k = 0; k < n / 16; ++k) {       // reduction over this sub-tile's queries
                const uint32_t acc = (i | X | k) != 0;
                umma_ss(tDV, aPT + X * 1024 + 2 * k, domn + (X * 4 + k) * 128, kIdKM64, acc);
                umma_ss(tDK, aDST + X * 1024 + 2 * k, qmn + (X * 4 + k) * 128, kIdKM64, acc);
              }
            }
            __syncwarp();
          }
          if (X == 1) {
            if (g > 0) {
              mbar_wait(dq_free, (g - 1) & 1);
              tc_fence_after_sync();
            }
            if (elect_one_sync()) {
#pragma unroll
              for (int k = 0; k < 8; ++k)             // reduction over the 128 keys of this item
                umma_ss(tDQ, aDSTq + k * 128, kmn + k * 128, kIdMM, k != 0);
              umma_commit(&qdo_empty[s]);
              if (i == nq - 1) umma_commit(&kv_empty[kb]);
              umma_commit(mma2_done);
            }
            __syncwarp();
          }
          if (has_next) mma1(nitn, ni, (g + 1) & 1, ((g + 1) >> 1) & 1, X);   // the other warpgroup keeps computing meanwhile
        }
        i = ni;
        itn = nitn;
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ============================================================ softmax warpgroups A (warps 4-7) and B (8-11)
    const int X = (warp - 4) >> 2;
    const int qd = warp & 3;
    const int row = qd * 32 + lane;                  // key row of S^T
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const int rin = row & 7;
    const uint32_t pt_row = smem_u32(smem + S::kPT) + X * 16384 + row * 128;
    const uint32_t dst_row = smem_u32(smem + S::kDST) + X * 16384 + row * 128;
    uint32_t f = 0, cx = 0, itn = 0;
    auto klen_of = [&](int item) {
      if (p.seqlens_k == nullptr || item >= p.num_items) return p.Lk;
      const int kl = __ldg(p.seqlens_k + item / (p.nk * p.H));
      return kl < p.Lk ? kl : p.Lk;
    };
    int klen_next = klen_of(blockIdx.x);
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++itn) {
      const int kt = item % p.nk, h = (item / p.nk) % p.H, b = item / (p.nk * p.H);
      const int klen = klen_next;
      klen_next = klen_of(item + gridDim.x);       // fetched a whole item ahead of its first use
      const int key = kt * 128 + row;
      const bool key_ok = key < klen;
      for (int i = 0; i < nq; ++i, ++f) {
        const int n = sub_n(i, X);
        if (n == 0) {
          // absent sub-tile: still observe every phase of mma2_done (a parity wait is only unambiguous for a waiter
          // that is at most one phase behind)
          if (f > 0) mbar_wait(mma2_done, (f - 1) & 1);
          continue;
        }
        const uint32_t s = f & 1;
        const float* st = s_stats + s * 256 + X * 64;
        mbar_wait(&s_full[X], cx & 1);
        ++cx;
        mbar_wait(&qdo_full[s], (f >> 1) & 1);       // already complete: acquires the statistics warp's writes
        tc_fence_after_sync();
        // P^T / dS^T of the previous tile must have been consumed (dV, dK and dQ MMAs) before they are overwritten; that
        // commit precedes the one that released S^T of this tile, so this wait does not stall
        if (f > 0) mbar_wait(mma2_done, (f - 1) & 1);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (c * 32 < n) {                            // columns >= n are never read by the dV / dK MMAs
            uint32_t sv[32], dv[32];
            tmem_ld32(tST + X * 64 + lane_off + c * 32, sv);
            tmem_ld32(tDPT + X * 64 + lane_off + c * 32, dv);
            tmem_wait_ld();
            uint32_t pk[16], dk[16];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int c0 = c * 32 + 4 * e;
              const float4 ls = *reinterpret_cast<const float4*>(st + c0);         // broadcast reads: 4 queries per LDS
              const float4 dl = *reinterpret_cast<const float4*>(st + 128 + c0);
              const float p0 = fast_exp2(fmaf(__uint_as_float(sv[4 * e]), p.scale_log2, -ls.x));
              const float p1 = fast_exp2(fmaf(__uint_as_float(sv[4 * e + 1]), p.scale_log2, -ls.y));
              const float p2 = fast_exp2(fmaf(__uint_as_float(sv[4 * e + 2]), p.scale_log2, -ls.z));
              const float p3 = fast_exp2(fmaf(__uint_as_float(sv[4 * e + 3]), p.scale_log2, -ls.w));
              const float d0 = p0 * fmaf(__uint_as_float(dv[4 * e]), p.scale, -dl.x);
              const float d1 = p1 * fmaf(__uint_as_float(dv[4 * e + 1]), p.scale, -dl.y);
              const float d2 = p2 * fmaf(__uint_as_float(dv[4 * e + 2]), p.scale, -dl.z);
              const float d3 = p3 * fmaf(__uint_as_float(dv[4 * e + 3]), p.scale, -dl.w);
              pk[2 * e] = key_ok ? pack_bf16x2(p0, p1) : 0u;   // padded keys: P = dS = 0 (exactly zero gradient)
              pk[2 * e + 1] = key_ok ? pack_bf16x2(p2, p3) : 0u;
              dk[2 * e] = key_ok ? pack_bf16x2(d0, d1) : 0u;
              dk[2 * e + 1] = key_ok ? pack_bf16x2(d2, d3) : 0u;
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint32_t chunk = static_cast<uint32_t>((c * 4 + g) ^ rin) * 16;
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pt_row + chunk), "r"(pk[4 * g]), "r"(pk[4 * g + 1]),
                           "r"(pk[4 * g + 2]), "r"(pk[4 * g + 3])
                           : "memory");
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst_row + chunk), "r"(dk[4 * g]), "r"(dk[4 * g + 1]),
                           "r"(dk[4 * g + 2]), "r"(dk[4 * g + 3])
                           : "memory");
            }
          }
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        mbar_arrive(&p_full[X]);
      }
    }
  } else if (warp >= 12) {
    // ============================================================ dQ drain
    const int qd = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    uint8_t* slabs = smem + S::kDQ + qd * 8192;
    const uint32_t srow = smem_u32(slabs) + lane * 128;
    if (warp == 12 && lane == 0) {
      tma_prefetch_desc(&tmDQ);
      tma_prefetch_desc(&tmDK);
      tma_prefetch_desc(&tmDV);
    }
    uint32_t f = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int h = (item / p.nk) % p.H, b = item / (p.nk * p.H);
      for (int i = 0; i < nq; ++i, ++f) {
        mbar_wait(mma2_done, f & 1);
        tc_fence_after_sync();
        if (lane == 0) tma_store_wait_read<0>();     // the previous tile's reductions have read the slabs
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(tDQ + lane_off + c * 32, v);
          tmem_wait_ld();
#pragma unroll
          for (int g = 0; g < 8; ++g)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + c * 4096 + ((g ^ (lane & 7)) * 16)), "r"(v[4 * g]),
                         "r"(v[4 * g + 1]), "r"(v[4 * g + 2]), "r"(v[4 * g + 3])
                         : "memory");
        }
        tc_fence_before_sync();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(dq_free);                      // TMEM is read: the next dQ MMA may overwrite it
          if (i * 128 + qd * 32 < p.Lq) {            // rows past Lq inside the box are clipped by the TMA unit
            tma_reduce_add_4d(&tmDQ, slabs, 0, i * 128 + qd * 32, h, b);
            tma_reduce_add_4d(&tmDQ, slabs + 4096, 32, i * 128 + qd * 32, h, b);
          }
          tma_store_commit();
        }
      }
      // item finished (the wait above covered its last MMAs): dV, dK -> bf16 slabs -> TMA store, so that the softmax
      // warpgroups go straight on to the next item
      if (lane == 0) tma_store_wait_read<0>();
      __syncwarp();
#pragma unroll
      for (int t = 0; t < 2; ++t) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32((t == 0 ? tDV : tDK) + lane_off + c * 32, v);
          tmem_wait_ld();
#pragma unroll
          for (int g = 0; g < 4; ++g)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + t * 4096 + (((c * 4 + g) ^ (lane & 7)) * 16)),
                         "r"(pack_bf16x2(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]))),
                         "r"(pack_bf16x2(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3]))),
                         "r"(pack_bf16x2(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]))),
                         "r"(pack_bf16x2(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7])))
                         : "memory");
        }
      }
      tc_fence_before_sync();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(dkv_free);                       // the accumulators are read: the next item may overwrite them
        const int kt = item % p.nk;
        if (kt * 128 + qd * 32 < p.Lk) {             // rows past Lk inside the box are clipped by the TMA unit
          tma_store_4d(&tmDV, slabs, 0, kt * 128 + qd * 32, h, b);
          tma_store_4d(&tmDK, slabs + 4096, 0, kt * 128 + qd * 32, h, b);
        }
        tma_store_commit();
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace vpt
