// Attention backward for sm_100a (tcgen05 + TMEM + TMA), head_dim 64, non-causal, per-sample valid key length.
// Autograd of scaled_dot_product_attention as Attention.forward calls it (src/models/jit/denoiser.py:351-397,
// src/modules/attention.py:98-129 of the reference).
//
// Persistent CTAs (one per SM, 16 warps) walk work items (128-key tile, head, sample) and, inside an item, the query
// tiles.  Every 128-query tile is two 64-query sub-tiles A and B, each owned by one softmax warpgroup:
//
//   tensor core    S^T_X = K Q_X^T, dP^T_X = V dO_X^T   (TMEM, 64 columns each)                          X = A, B
//   warpgroup X    P^T_X = exp2(S^T_X c - lse) -> bf16 back into TMEM (A operand of the dV MMA);
//                  dS^T_X = P^T_X (dP^T_X - delta) scale -> bf16, shared memory (double buffered by tile parity: dQ of
//                  tile g reads dS^T of both sub-tiles while the warpgroups already write tile g+1)
//   tensor core    dV += P^T_X dO_X, dK += dS^T_X Q_X   (TMEM accumulators over the whole item),  dQ = dS K (per tile)
//   drain warps    dQ tile -> fp32 shared-memory slab -> TMA reduce-add (dQ is summed across key tiles in global
//                  memory; per-thread red.global measured ~10k cycles per tile on the LSU, profiles/r1e_attn_bwd.txt)
//
// The tensor core runs ahead of the warpgroups: S^T_X / dP^T_X of the NEXT tile are issued as soon as warpgroup X has the
// current ones in registers, so a warpgroup never waits for its scores behind the other warpgroup's MMAs.  Four things
// make that possible (each measured, profiles/r1n_attn_bwd.txt, tools/probes/umma_rate.cu):
//   * two issuing warps - tcgen05.mma issue blocks while the pipe is busy, so one warp issuing everything kept the
//     scores of the next tile stuck behind its own blocked dV / dK / dQ issue;
//   * the Q / dO sub-tiles (with their lse / delta*scale rows) live in one ring of kSlots sub-tile slots, released right
//     after the dV / dK MMAs that read them last, not a whole tile later after dQ;
//   * K / V of the next item are fetched by their own warp a whole item ahead;
//   * P^T in TMEM and a second dS^T buffer, so no warpgroup waits for the dQ MMA of the previous tile.
//
// Warp roles: 0 Q/dO producer (ring), 2 TMEM allocator + K/V producer (double buffered, one item ahead), 1 issues
// S^T / dP^T, 3 issues dV / dK / dQ, 4-7 warpgroup A, 8-11 warpgroup B, 12-15 drain (dQ per tile, dV / dK per item).
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

// head_dim 64: the layout described above.  head_dim 80 (JiT-H): every operand tile is two 128-byte-swizzled 64-column
// blocks (the second zero-filled past column 80 by the TMA unit), which doubles K / V / Q / dO in shared memory -- so K / V
// are single-buffered, the Q / dO ring has 3 slots, dS^T one buffer, and (TMEM: 2 * 64 + 2 * 64 + 3 * 80 = 496 columns) P^T is
// written over the S^T columns it came from, which ties the next S^T_X to the dV MMA of the current one.
template <int HD>
struct AttnBwd2SmemT {
  static constexpr int kBlk = (HD + 63) / 64;         // 64-column blocks per operand row
  static constexpr int kSlots = HD == 64 ? 5 : 3;     // Q / dO ring: sub-tile slots of [64 queries x HD] bf16 each
  static constexpr int kKVBufs = HD == 64 ? 2 : 1;
  static constexpr int kDSTBufs = HD == 64 ? 2 : 1;
  static constexpr bool kPInPlace = HD != 64;         // P^T_X over S^T_X columns [0, 32)
  static constexpr int kSub = kBlk * 8192;            // one Q or dO sub-tile
  static constexpr int kKVTile = kBlk * 16384;        // one K or V tile
  static constexpr int kK = 0;
  static constexpr int kV = kKVBufs * kKVTile;
  static constexpr int kQ = 2 * kKVBufs * kKVTile;
  static constexpr int kDO = kQ + kSlots * kSub;
  static constexpr int kDST = kDO + kSlots * kSub;    // kDSTBufs x [128 keys x 128 queries] bf16 = 2 K-atoms (A, B) each
  static constexpr int kDQ = kDST + kDSTBufs * 32768; // 4 drain warps x one slab of [32 x 128 B], 128B-swizzled
  static constexpr int kStats = kDQ + 4 * 4096;       // per slot: lse2[64], delta*scale[64] fp32
  static constexpr int kBars = kStats + kSlots * 512;
  // kv_full[2] kv_empty[2] s_full[2] p_full[2] mma2_done[2] st_free[2] dvdk_done[2] dq_free dkv_free
  // qdo_full[kSlots] qdo_empty[kSlots]
  static constexpr int kNumBars = 16 + 2 * kSlots;
  static constexpr int kTmemSlot = kBars + kNumBars * 8;
  static constexpr int kTotal = kTmemSlot + 16;       // no alignment slack: the dynamic segment starts 1024B-aligned
  static_assert(kTotal <= 232448, "shared memory");
};
using AttnBwd2Smem = AttnBwd2SmemT<64>;

__device__ __forceinline__ int attn_round16(int x) { return (x + 15) & ~15; }

// Position in the Q / dO ring; every role steps it once per PRESENT sub-tile, in the order A(i), B(i), A(i+1), ...
template <int kSlots>
struct AttnRingT {
  uint32_t slot = 0, phase = 0;
  __device__ __forceinline__ void next() {
    if (++slot == kSlots) {
      slot = 0;
      phase ^= 1;
    }
  }
};

// tmQ / tmDO: 4-D bf16 maps (hd, L, H, B), box {64, 64, 1, 1}; tmK / tmV: box {64, 128, 1, 1}; all SWIZZLE_128B
// tmDQ: 4-D fp32 map (hd, Lq, H, B) of the dQ accumulator, box {32, 32, 1, 1}, SWIZZLE_128B
// tmDK / tmDV: 4-D bf16 maps (hd, Lk, H, B) of the outputs, box {64, 32, 1, 1}, SWIZZLE_128B
template <int HD>
__global__ void __launch_bounds__(512, 1)
attn_bwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                 const __grid_constant__ CUtensorMap tmDQ, const __grid_constant__ CUtensorMap tmDK,
                 const __grid_constant__ CUtensorMap tmDV, const AttnBwd2Params p) {
  using S = AttnBwd2SmemT<HD>;
  using AttnRing = AttnRingT<S::kSlots>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // 128B-swizzled tiles need a 1024B-aligned base
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBars);
  uint64_t* kv_full = bars;          // [2]
  uint64_t* kv_empty = bars + 2;     // [2]
  uint64_t* s_full = bars + 4;       // [2]  A, B
  uint64_t* p_full = bars + 6;       // [2]  A, B
  uint64_t* mma2_done = bars + 8;    // [2] by tile parity: dV, dK, dQ of that tile are done
  uint64_t* st_free = bars + 10;     // [2] warpgroup X has S^T_X / dP^T_X in registers
  uint64_t* dvdk_done = bars + 12;   // [2] the dV / dK MMAs of warpgroup X's last sub-tile have read P^T_X
  uint64_t* dq_free = bars + 14;
  uint64_t* dkv_free = bars + 15;
  uint64_t* qdo_full = bars + 16;               // [kSlots]  Q, dO and the statistics rows of the sub-tile have landed
  uint64_t* qdo_empty = bars + 16 + S::kSlots;  // [kSlots]  the dV / dK MMAs of that sub-tile are done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::kTmemSlot);
  float* s_stats = reinterpret_cast<float*>(smem + S::kStats);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq = p.nq;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 128);
      mbar_init(&mma2_done[s], 1);
      mbar_init(&st_free[s], 128);
      mbar_init(&dvdk_done[s], 1);
    }
    for (int s = 0; s < S::kSlots; ++s) {
      mbar_init(&qdo_full[s], 1);
      mbar_init(&qdo_empty[s], 1);
    }
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
  // columns (head_dim 64): S^T_A 0, S^T_B 64, dP^T_A 128, dP^T_B 192, dV 256, dK 320, dQ 384, P^T_A 448, P^T_B 480 (bf16
  // pairs); head_dim 80: dV 256, dK 336, dQ 416 and P^T_X over S^T_X
  const uint32_t tST = tmem_base, tDPT = tmem_base + 128, tDV = tmem_base + 256, tDK = tDV + HD, tDQ = tDK + HD;
  auto tPT_of = [&](int X) { return S::kPInPlace ? tST + X * 64 : tmem_base + 448 + X * 32; };

  // queries of tile i that sub-tile X covers, rounded up to the UMMA granularity (0 = sub-tile absent)
  auto sub_n = [&](int i, int X) {
    const int valid = min(128, p.Lq - i * 128) - X * 64;
    return valid <= 0 ? 0 : min(64, attn_round16(valid));
  };

  if (warp == 0) {
    // ============================================================ Q / dO producer
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmDO);
      AttnRing r;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int h = (item / p.nk) % p.H, b = item / (p.nk * p.H);
        const long stat_row = (static_cast<long>(b) * p.H + h) * (static_cast<long>(nq) * 128);
        for (int i = 0; i < nq; ++i) {
#pragma unroll 1
          for (int X = 0; X < 2; ++X) {
            if (sub_n(i, X) == 0) continue;
            mbar_wait(&qdo_empty[r.slot], r.phase ^ 1);
            uint64_t* full = &qdo_full[r.slot];
            mbar_arrive_expect_tx(full, 2 * S::kSub + 512);
#pragma unroll
            for (int blk = 0; blk < S::kBlk; ++blk) {
              tma_load_4d(&tmQ, full, smem + S::kQ + r.slot * S::kSub + blk * 8192, blk * 64, i * 128 + X * 64, h, b);
              tma_load_4d(&tmDO, full, smem + S::kDO + r.slot * S::kSub + blk * 8192, blk * 64, i * 128 + X * 64, h, b);
            }
            const long srow = stat_row + i * 128 + X * 64;
            bulk_load_1d(smem + S::kStats + r.slot * 512, p.lse2 + srow, 256, full);
            bulk_load_1d(smem + S::kStats + r.slot * 512 + 256, p.delta + srow, 256, full);
            r.next();
          }
        }
      }
    }
  } else if (warp == 2) {
    // ============================================================ K / V producer
    // Its own warp: the buffer of item n+1 frees at the end of item n-1, and waiting for that inside the Q / dO producer
    // would hold back the Q / dO loads of item n (measured: ~3000 cycles per item).
    if (lane == 0) {
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      uint32_t n = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++n) {
        const int kt = item % p.nk, h = (item / p.nk) % p.H, b = item / (p.nk * p.H);
        const uint32_t kb = n % S::kKVBufs;
        mbar_wait(&kv_empty[kb], ((n / S::kKVBufs) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[kb], 2 * S::kKVTile);
#pragma unroll
        for (int blk = 0; blk < S::kBlk; ++blk) {
          tma_load_4d(&tmK, &kv_full[kb], smem + S::kK + kb * S::kKVTile + blk * 16384, blk * 64, kt * 128, h, b);
          tma_load_4d(&tmV, &kv_full[kb], smem + S::kV + kb * S::kKVTile + blk * 16384, blk * 64, kt * 128, h, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == 3) {
    // ============================================================ MMA issuers (converged warps, one elected lane issues)
    constexpr uint32_t kIdKM64 = umma_idesc_bf16(128, HD, 0, 1);   // dV += P^T dO, dK += dS^T Q (B MN-major, N = head_dim)
    constexpr uint32_t kIdMM = umma_idesc_bf16(128, HD, 1, 1);     // dQ = dS K (A = dS^T viewed MN-major, B MN-major)
    const uint32_t smem_base = smem_u32(smem);
    const uint64_t dK_ = umma_smem_desc(0, 16, 1024, kLayoutSW128);       // K-major operand, + (addr >> 4)
    const uint64_t dMN = umma_smem_desc(0, 8192, 1024, kLayoutSW128);     // MN-major Q / dO sub-tile: 64-column blocks 8 KB apart
    const uint64_t dMNkv = umma_smem_desc(0, 16384, 1024, kLayoutSW128);  // MN-major K tile: 64-column blocks 16 KB apart
    const uint64_t dMNq = umma_smem_desc(0, 16384, 1024, kLayoutSW128);   // dS^T viewed MN-major (query atoms 16 KB apart)
    const uint64_t aDST0 = dK_ + ((smem_base + S::kDST) >> 4);     // + db * 2048 (32 KB buffers) + X * 1024 + 2 k
    const uint64_t aDSTq0 = dMNq + ((smem_base + S::kDST) >> 4);
    const int my_items = (p.num_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const uint32_t total = static_cast<uint32_t>(my_items) * nq;
    const int n_last[2] = {sub_n(nq - 1, 0), sub_n(nq - 1, 1)};
    AttnRing r;
    PROF_DECL(8)
    if (warp == 1) {
      // ---- S^T_X = K Q_X^T, dP^T_X = V dO_X^T, as soon as warpgroup X has the previous pair in registers
      uint32_t issued[2] = {0, 0};
      uint32_t itn = 0;
      int i = 0;
      for (uint32_t g = 0; g < total; ++g) {
        const uint32_t kb = itn % S::kKVBufs;
        const uint64_t kd = dK_ + ((smem_base + S::kK + kb * S::kKVTile) >> 4), vd = dK_ + ((smem_base + S::kV + kb * S::kKVTile) >> 4);
        PROF(0)
        if (i == 0) mbar_wait(&kv_full[kb], (itn / S::kKVBufs) & 1);
        PROF(4)
#pragma unroll 1
        for (int X = 0; X < 2; ++X) {
          const int n = i == nq - 1 ? n_last[X] : 64;
          if (n == 0) continue;
          PROF(0)
          // S^T_X / dP^T_X are single-buffered: free once warpgroup X has them in registers -- or, with P^T_X written over
          // S^T_X, once the dV MMA that reads it is done
          if (issued[X] > 0) mbar_wait(S::kPInPlace ? &dvdk_done[X] : &st_free[X], (issued[X] - 1) & 1);
          ++issued[X];
          PROF(1)
          mbar_wait(&qdo_full[r.slot], r.phase);
          PROF(2)
          tc_fence_after_sync();
          const uint64_t qd_ = dK_ + ((smem_base + S::kQ + r.slot * S::kSub) >> 4);
          const uint64_t dod = dK_ + ((smem_base + S::kDO + r.slot * S::kSub) >> 4);
          const uint32_t idesc = umma_idesc_bf16(128, n, 0, 0);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss(tST + X * 64, kd + 2 * k, qd_ + 2 * k, idesc, k != 0);
#pragma unroll
            for (int k = 0; k < (HD - 64) / 16; ++k)      // head-dim columns 64..: second blocks (+16 KB of K / V, +8 KB of Q / dO)
              umma_ss(tST + X * 64, kd + 1024 + 2 * k, qd_ + 512 + 2 * k, idesc, 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss(tDPT + X * 64, vd + 2 * k, dod + 2 * k, idesc, k != 0);
#pragma unroll
            for (int k = 0; k < (HD - 64) / 16; ++k)
              umma_ss(tDPT + X * 64, vd + 1024 + 2 * k, dod + 512 + 2 * k, idesc, 1);
            umma_commit(&s_full[X]);
          }
          __syncwarp();
          r.next();
          PROF(3)
        }
        if (++i == nq) {
          i = 0;
          ++itn;
        }
      }
#ifdef VPT_BWD_PROF
      PROF(0)
      if (blockIdx.x == 0 && lane == 0)
        printf("mma S : other %lld st_free %lld qdo_full %lld issue %lld kv_full %lld\n", prof_[0], prof_[1], prof_[2], prof_[3],
               prof_[4]);
#endif
    } else {
      // ---- dV += P^T_X dO_X, dK += dS^T_X Q_X per sub-tile; dQ = dS K per tile
      uint32_t cnt[2] = {0, 0};                      // sub-tiles of each kind taken from the warpgroups so far
      uint32_t itn = 0;
      int i = 0;
      for (uint32_t g = 0; g < total; ++g) {
        const uint32_t kb = itn % S::kKVBufs;
        const uint64_t kmn = dMNkv + ((smem_base + S::kK + kb * S::kKVTile) >> 4);
        const uint32_t db = g % S::kDSTBufs;          // dS^T buffer of this tile
        PROF(0)
        if (i == 0) mbar_wait(&kv_full[kb], (itn / S::kKVBufs) & 1);   // long complete (the scores came from it): acquires K for dQ
        PROF(6)
#pragma unroll 1
        for (int X = 0; X < 2; ++X) {
          const int n = i == nq - 1 ? n_last[X] : 64;
          if (n > 0) {
            PROF(0)
            mbar_wait(&p_full[X], cnt[X] & 1);
            ++cnt[X];
            PROF(1)
            if (i == 0 && X == 0 && itn > 0) mbar_wait(dkv_free, (itn - 1) & 1);   // previous item's dV / dK are out
            PROF(2)
            tc_fence_after_sync();
            const uint64_t qmn = dMN + ((smem_base + S::kQ + r.slot * S::kSub) >> 4), domn = dMN + ((smem_base + S::kDO + r.slot * S::kSub) >> 4);
            const uint32_t tPT = tPT_of(X);
            if (elect_one_sync()) {
              if (n == 64) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {          // reduction over this sub-tile's queries
                  const uint32_t acc = (i | X | k) != 0;
                  umma_ts(tDV, tPT + k * 8, domn + k * 128, kIdKM64, acc);      // A = P^T_X in TMEM
                  umma_ss(tDK, aDST0 + db * 2048 + X * 1024 + 2 * k, qmn + k * 128, kIdKM64, acc);
                }
              } else {
                for (int k = 0; k < n / 16; ++k) {
                  const uint32_t acc = (i | X | k) != 0;
                  umma_ts(tDV, tPT + k * 8, domn + k * 128, kIdKM64, acc);
                  umma_ss(tDK, aDST0 + db * 2048 + X * 1024 + 2 * k, qmn + k * 128, kIdKM64, acc);
                }
              }
              umma_commit(&dvdk_done[X]);
              umma_commit(&qdo_empty[r.slot]);         // Q_X / dO_X are not read again (dQ reads K)
            }
            __syncwarp();
            r.next();
            PROF(3)
          }
          if (X == 1) {
            PROF(0)
            if (g > 0) {
              mbar_wait(dq_free, (g - 1) & 1);
              tc_fence_after_sync();
            }
            PROF(4)
            if (elect_one_sync()) {
#pragma unroll
              for (int k = 0; k < 8; ++k)             // reduction over the 128 keys of this item
                umma_ss(tDQ, aDSTq0 + db * 2048 + k * 128, kmn + k * 128, kIdMM, k != 0);
              if (i == nq - 1) umma_commit(&kv_empty[kb]);
              umma_commit(&mma2_done[db]);
            }
            __syncwarp();
            PROF(5)
          }
        }
        if (++i == nq) {
          i = 0;
          ++itn;
        }
      }
#ifdef VPT_BWD_PROF
      PROF(0)
      if (blockIdx.x == 0 && lane == 0)
        printf("mma G : other %lld p_full %lld dkv_free %lld dV/dK issue %lld dq_free %lld dQ issue %lld kv_full %lld\n", prof_[0],
               prof_[1], prof_[2], prof_[3], prof_[4], prof_[5], prof_[6]);
#endif
    }
  } else if (warp >= 4 && warp < 12) {
    // ============================================================ softmax warpgroups A (warps 4-7) and B (8-11)
    const int X = (warp - 4) >> 2;
    const int qd = warp & 3;
    const int row = qd * 32 + lane;                  // key row of S^T
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const int rin = row & 7;
    const uint32_t dst_row0 = smem_u32(smem + S::kDST) + X * 16384 + row * 128;   // + (tile parity) * 32768
    uint32_t f = 0, cx = 0;
    AttnRing r;
    auto klen_of = [&](int item) {
      if (p.seqlens_k == nullptr || item >= p.num_items) return p.Lk;
      const int kl = __ldg(p.seqlens_k + item / (p.nk * p.H));
      return kl < p.Lk ? kl : p.Lk;
    };
    int klen_next = klen_of(blockIdx.x);
    PROF_DECL(8)
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int kt = item % p.nk;
      const int klen = klen_next;
      klen_next = klen_of(item + gridDim.x);       // fetched a whole item ahead of its first use
      const int key = kt * 128 + row;
      const bool key_ok = key < klen;
      for (int i = 0; i < nq; ++i, ++f) {
        const int n = sub_n(i, X);
        const uint32_t db = f % S::kDSTBufs, dst_use = f / S::kDSTBufs;     // buffer and how often it has been used before
        if (X == 1) r.next();                        // sub-tile A of this tile (always present) sits before ours in the ring
        if (n == 0) {
          // absent sub-tile: still observe this buffer's mma2_done phase (a parity wait is only unambiguous for a waiter
          // that is at most one phase behind)
          if (dst_use >= 1) mbar_wait(&mma2_done[db], (dst_use - 1) & 1);
          continue;
        }
        const float* st = s_stats + r.slot * 128;
        PROF(0)
        mbar_wait(&s_full[X], cx & 1);
        PROF(1)
        mbar_wait(&qdo_full[r.slot], r.phase);       // complete since the scores were issued: acquires the statistics rows
        tc_fence_after_sync();
        const uint32_t dst_row = dst_row0 + db * 32768;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (c * 32 < n) {                            // columns >= n are never read by the dV / dK MMAs
            uint32_t sv[32], dv[32];
            tmem_ld32(tST + X * 64 + lane_off + c * 32, sv);
            tmem_ld32(tDPT + X * 64 + lane_off + c * 32, dv);
            tmem_wait_ld();
            PROF(2)
            if ((c + 1) * 32 >= n) {                   // the whole sub-tile is in registers: the tensor core may refill S^T / dP^T
              tc_fence_before_sync();
              mbar_arrive(&st_free[X]);
            }
            uint32_t pk[16], dk[16];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int c0 = c * 32 + 4 * e;
              const float4 ls = *reinterpret_cast<const float4*>(st + c0);         // broadcast reads: 4 queries per LDS
              const float4 dl = *reinterpret_cast<const float4*>(st + 64 + c0);
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
            PROF(3)
            if (c == 0) {
              // P^T_X (TMEM) was last read by the dV MMAs of this warpgroup's previous sub-tile; dS^T[db] (shared) by the
              // dK / dQ MMAs of the tile two back.  Neither wait involves the other warpgroup's current tile.  (Holding
              // both chunks' P^T in registers to store them last did not pay: the wait moved into a slower schedule.)
              if (cx > 0) mbar_wait(&dvdk_done[X], (cx - 1) & 1);
              if (dst_use >= 1) mbar_wait(&mma2_done[db], (dst_use - 1) & 1);
              tc_fence_after_sync();
              PROF(4)
            }
            // 32 queries = 16 columns of bf16 pairs (in place at head_dim 80: inside S^T columns that are already read)
            tmem_st16(tPT_of(X) + lane_off + c * 16, pk);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint32_t chunk = static_cast<uint32_t>((c * 4 + g) ^ rin) * 16;
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst_row + chunk), "r"(dk[4 * g]), "r"(dk[4 * g + 1]),
                           "r"(dk[4 * g + 2]), "r"(dk[4 * g + 3])
                           : "memory");
            }
          }
        }
        ++cx;
        tmem_wait_st();
        fence_proxy_async_smem();
        tc_fence_before_sync();
        mbar_arrive(&p_full[X]);
        r.next();
        if (X == 0 && sub_n(i, 1) > 0) r.next();     // sub-tile B of this tile follows ours in the ring
        PROF(5)
      }
    }
#ifdef VPT_BWD_PROF
    PROF(0)
    if (blockIdx.x == 0 && lane == 0 && qd == 0)
      printf("wg%d   : other %lld s_full %lld tmem_ld %lld math %lld dvdk/mma2_done %lld st+fence %lld\n", X, prof_[0], prof_[1],
             prof_[2], prof_[3], prof_[4], prof_[5]);
#endif
  } else if (warp >= 12) {
    // ============================================================ drain: dQ per tile, dV / dK per item
    // One 4 KB slab per warp ([32 rows x 128 B], 128B-swizzled), reused for every 32-column piece: the ring of Q / dO
    // slots needs the shared memory more than the drain needs overlap (dq_free / dkv_free are waited on for < 5 % of the
    // issuing warp's time).
    const int qd = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    uint8_t* slab = smem + S::kDQ + qd * 4096;
    const uint32_t srow = smem_u32(slab) + lane * 128;
    if (warp == 12 && lane == 0) {
      tma_prefetch_desc(&tmDQ);
      tma_prefetch_desc(&tmDK);
      tma_prefetch_desc(&tmDV);
    }
    uint32_t f = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int h = (item / p.nk) % p.H, b = item / (p.nk * p.H);
      for (int i = 0; i < nq; ++i, ++f) {
        mbar_wait(&mma2_done[f % S::kDSTBufs], (f / S::kDSTBufs) & 1);
        tc_fence_after_sync();
        if (i == nq - 1) {
          // item finished (the wait above covered its last MMAs).  dV / dK go first: the first MMA of the next item waits
          // for these accumulators, its dQ MMA only comes a whole tile later.  bf16 slab -> TMA store, 64 columns at a time
          // (head_dim 80: the second piece holds 16 columns; the store's box is clipped at the tensor's edge).
          const int kt = item % p.nk;
#pragma unroll
          for (int t = 0; t < 2; ++t) {
#pragma unroll
            for (int pc = 0; pc < S::kBlk; ++pc) {
              uint32_t va[32], vb[32];
              const uint32_t tacc = (t == 0 ? tDV : tDK) + lane_off + pc * 64;
              if (pc == 0) {
                tmem_ld32(tacc, va);
                tmem_ld32(tacc + 32, vb);
              } else {
                tmem_ld16(tacc, *reinterpret_cast<uint32_t(*)[16]>(&va[0]));
              }
              tmem_wait_ld();
              if (t == 1 && pc == S::kBlk - 1) {
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(dkv_free);  // the accumulators are read: the next item may overwrite them
              }
              if (lane == 0) tma_store_wait_read<0>();   // the slab's previous TMA read is done
              __syncwarp();
#pragma unroll
              for (int g = 0; g < (pc == 0 ? 8 : 2); ++g) {
                const uint32_t* v = g < 4 ? &va[g * 8] : &vb[(g - 4) * 8];
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((g ^ (lane & 7)) * 16)),
                             "r"(pack_bf16x2(__uint_as_float(v[0]), __uint_as_float(v[1]))),
                             "r"(pack_bf16x2(__uint_as_float(v[2]), __uint_as_float(v[3]))),
                             "r"(pack_bf16x2(__uint_as_float(v[4]), __uint_as_float(v[5]))),
                             "r"(pack_bf16x2(__uint_as_float(v[6]), __uint_as_float(v[7])))
                             : "memory");
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                if (kt * 128 + qd * 32 < p.Lk)           // rows past Lk inside the box are clipped by the TMA unit
                  tma_store_4d(t == 0 ? &tmDV : &tmDK, slab, pc * 64, kt * 128 + qd * 32, h, b);
                tma_store_commit();
              }
            }
          }
        }
        constexpr int kPieces = (HD + 31) / 32;      // dQ leaves in 32-column fp32 pieces (head_dim 80: the last holds 16)
#pragma unroll
        for (int c = 0; c < kPieces; ++c) {
          uint32_t v[32];
          if (c * 32 + 32 <= HD) {
            tmem_ld32(tDQ + lane_off + c * 32, v);
          } else {
            tmem_ld16(tDQ + lane_off + c * 32, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
#pragma unroll
            for (int e = 16; e < 32; ++e) v[e] = 0u;
          }
          tmem_wait_ld();
          if (c == kPieces - 1) {
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(dq_free);     // TMEM is read: the next dQ MMA may overwrite it
          }
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
#pragma unroll
          for (int g = 0; g < 8; ++g)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((g ^ (lane & 7)) * 16)), "r"(v[4 * g]),
                         "r"(v[4 * g + 1]), "r"(v[4 * g + 2]), "r"(v[4 * g + 3])
                         : "memory");
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (i * 128 + qd * 32 < p.Lq)              // rows past Lq (and columns past head_dim) are clipped by the TMA unit
              tma_reduce_add_4d(&tmDQ, slab, c * 32, i * 128 + qd * 32, h, b);
            tma_store_commit();
          }
        }
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace vpt
