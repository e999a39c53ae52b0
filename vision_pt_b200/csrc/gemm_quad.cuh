// LoRA GEMM on clusters of FOUR CTAs = two CTA pairs (tcgen05 cta_group::2) that share the weight tile by TMA multicast.
// Same math, warp roles and epilogue as gemm_pair.cuh (read that file first); what changes is the operand traffic:
//
//   pair p (ranks 2p, 2p+1) computes rows [512 q + 256 p, +256) of the SAME 192-column tile as the other pair, so both
//   pairs need the same B rows.  B-half h (rank & 1) lands in the CTAs {h, h+2}: each of the two loads a part of it
//   (56 + 48 rows, or 56 + 32 + the 16 LoRA rows) and multicasts it to both.  Per k-step an SM pulls 16 KB of A and
//   6.5 KB of B out of L2 instead of 16 + 13 KB: 55 instead of 72 B per SM-cycle.  The pair kernel measured 60 % tensor-pipe
//   activity at exactly the L2 rate that 72 B/cycle implies (and BN = 128, 89 B/cycle, was 1.24x slower: L2-bound).
//
// Plain (Hopper-style) TMA: every byte that lands in a CTA is credited to THAT CTA's local_full barrier; the follower of a
// pair forwards "my stage is complete" to its leader with one remote arrive per stage (warp 2), and the MMA issuer waits
// for its own and the follower's.  A stage is free again when BOTH pairs have consumed it: each leader's tcgen05.commit
// multicasts to the empty barrier (count 2) of all four CTAs.
#pragma once
#include "gemm_pair.cuh"

namespace vpt {

constexpr int kQuadStagesMax = 6;

// extra barriers of the quad kernel live after the pair kernel's: local_full[stages], peer_full[stages]
template <int BN, bool kLoRA>
struct QuadSmem : PairSmem<BN, kLoRA> {
  using P = PairSmem<BN, kLoRA>;
  static constexpr int kOffBars2 = P::kOffTmemSlot + 16;
  static constexpr int kTotal = kOffBars2 + 2 * P::kStages * 8 + 1024;
};

// Multicast form of tma_load_2d: the tile lands at the same offset in every CTA of `mask` and each of them gets the
// bytes credited to the barrier at `bar`'s offset in ITS OWN shared memory.
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* m, uint64_t* bar, void* smem_dst, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1)
      : "memory");
}
// tcgen05.commit arriving on the barrier at this offset in the CTAs of `mask`
__device__ __forceinline__ void umma_commit_mask(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
      "h"(mask)
      : "memory");
}

// tmA : A [M, R], box {64, 128};  tmB56 / tmB48 / tmB32: Bw [NO, R] with 56 / 48 / 32-row boxes;  tmP: P [16, R], box {64, 16};
// tmD / tmR: D and residual [M, NO], box {64, 32}.  All SWIZZLE_128B.
template <int BN, bool kLoRA>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(kPairThreads, 1)
gemm_quad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB56,
                 const __grid_constant__ CUtensorMap tmB48, const __grid_constant__ CUtensorMap tmB32,
                 const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmD,
                 const __grid_constant__ CUtensorMap tmR, const PairParams p) {
  using S = QuadSmem<BN, kLoRA>;
  static_assert(BN == 192, "the row split of the multicast loads is written for BN = 192");
  static_assert(BN % 64 == 0 && BN >= 64 && S::kNT <= 256, "unsupported BN");
  constexpr int kStages = S::kStages;
  constexpr uint32_t kIdescMain = umma_idesc_bf16(256, S::kNT, 0, 0);
  constexpr uint32_t kIdescLora = umma_idesc_bf16(256, BN, 0, 0);
  constexpr uint32_t kStageTx = S::kABytes + S::kNH * 128;     // bytes each CTA's loads credit per stage

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOffBars);
  uint64_t* bars2 = reinterpret_cast<uint64_t*>(smem + S::kOffBars2);
  uint64_t* local_full = bars2;                // [kStages] per CTA: every byte that lands in this CTA (A, both B parts)
  uint64_t* peer_full = bars2 + kStages;       // [kStages] leader's: the follower's stage is complete (remote arrive)
  uint64_t* empty = bars + kStages;            // [kStages] per CTA, count 2: both pairs' leaders commit to all four CTAs
  uint64_t* tmem_full = bars + 2 * kStages;    // [2] per CTA: main loop of a tile finished
  uint64_t* tmem_full2 = tmem_full + 2;        // [2] per CTA: rank-16 update finished
  uint64_t* tmem_empty = tmem_full2 + 2;       // [2] leader's: 8 epilogue warps (both CTAs) drained the buffer
  uint64_t* ts_full = tmem_empty + 2;          // [1] leader's: 8 epilogue warps staged Ts / Q
  uint64_t* res_full = ts_full + 1;            // [4 warps][kSlabs] residual slab landed (per epilogue warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::kOffTmemSlot);
  float* s_bias = reinterpret_cast<float*>(smem + S::kOffBias);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t pair = rank >> 1, half = rank & 1;
  const bool leader = half == 0;
  const uint32_t leader_rank = rank & ~1u;
  const uint16_t pair_mask = static_cast<uint16_t>(3u << (2 * pair));
  const int cluster_id = blockIdx.x >> 2;
  const int num_clusters = gridDim.x >> 2;
  const int num_m_quads = (p.num_m_pairs + 1) / 2;
  const int num_tiles = num_m_quads * p.num_n_tiles;
  const int ksteps = (p.R + 63) / 64;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&local_full[s], 1);
      mbar_init(&peer_full[s], 1);
      mbar_init(&empty[s], 2);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_full2[b], 1);
      mbar_init(&tmem_empty[b], 8);
    }
    mbar_init(ts_full, 8);
    for (int i = 0; i < 4 * S::kSlabs; ++i) mbar_init(&res_full[i], 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB56);
    tma_prefetch_desc(&tmB48);
    if (kLoRA && !leader) {
      tma_prefetch_desc(&tmB32);
      tma_prefetch_desc(&tmP);
    }
    tma_prefetch_desc(&tmD);
    if (p.residual != nullptr) tma_prefetch_desc(&tmR);
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();                           // the peer's barriers exist before anything remote touches them
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();                                   // set-up above overlapped the previous kernel; its data is visible now

  if (warp == 0) {
    // ============================================================ TMA producer (all four CTAs)
    if (lane == 0) {
      uint32_t it = 0;
      const uint16_t mc = static_cast<uint16_t>((1u << half) | (1u << (half + 2)));   // the two CTAs that hold B-half `half`
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m0 = (tile / p.num_n_tiles) * 512 + static_cast<int>(pair) * 256 + static_cast<int>(half) * 128;
        const int o0 = (tile % p.num_n_tiles) * BN;
        for (int ks = 0; ks < ksteps; ++ks, ++it) {
          const int s = it % kStages;
          mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);          // both pairs are done with this stage, in all four CTAs
          uint8_t* sa = smem + s * S::kStageBytes;
          uint8_t* sb = sa + S::kABytes;
          mbar_arrive_expect_tx(&local_full[s], kStageTx);
          tma_load_2d(&tmA, &local_full[s], sa, ks * 64, m0);
          const int r0 = o0 + static_cast<int>(half) * S::kNH;     // first weight row of this B-half
          if (kLoRA) {
            // half 0: 104 weight rows = 56 (pair 0) + 48 (pair 1); half 1: 88 weight rows = 56 + 32, then the 16 LoRA rows
            if (pair == 0) {
              tma_load_2d_mc(&tmB56, &local_full[s], sb, ks * 64, r0, mc);
            } else if (half == 0) {
              tma_load_2d_mc(&tmB48, &local_full[s], sb + 56 * 128, ks * 64, r0 + 56, mc);
            } else {
              tma_load_2d_mc(&tmB32, &local_full[s], sb + 56 * 128, ks * 64, r0 + 56, mc);
              tma_load_2d_mc(&tmP, &local_full[s], sb + 88 * 128, ks * 64, 0, mc);
            }
          } else {
            tma_load_2d_mc(&tmB48, &local_full[s], sb + static_cast<int>(pair) * 48 * 128, ks * 64, r0 + static_cast<int>(pair) * 48, mc);
          }
        }
      }
    }
  } else if (warp == 2) {
    // ============================================================ follower: forward "my stage landed" to the pair's leader
    if (!leader && lane == 0) {
      uint32_t it = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        for (int ks = 0; ks < ksteps; ++ks, ++it) {
          const int s = it % kStages;
          mbar_wait(&local_full[s], (it / kStages) & 1);
          mbar_arrive_cluster(mapa_shared(smem_u32(&peer_full[s]), leader_rank));
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer (leader CTA; converged warp, one elected lane)
    if (leader) {
      uint32_t it = 0, lt = 0;
      bool lora_pending = false;
      uint32_t pend_buf = 0, pend_parity = 0;
      const uint32_t smem_base = smem_u32(smem);
      const uint64_t ts_desc = umma_smem_desc(smem_base + S::kOffTs, 128, 256, kLayoutNone);
      const uint64_t q_desc = umma_smem_desc(smem_base + S::kOffQ, 128, 256, kLayoutNone);
      const uint64_t sw_desc = umma_smem_desc(0, 16, 1024, kLayoutSW128);   // + (address >> 4) per operand
      auto issue_lora = [&]() {
        tc_fence_after_sync();
        if (elect_one_sync()) {
          umma_ss_pair(tmem_base + pend_buf * 256, ts_desc, q_desc, kIdescLora, 1);
          umma_commit_mask(&tmem_full2[pend_buf], pair_mask);
        }
        __syncwarp();
        lora_pending = false;
      };
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++lt) {
        const uint32_t buf = lt & 1;
        const uint32_t d_tmem = tmem_base + buf * 256;
        mbar_wait(&tmem_empty[buf], ((lt >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        for (int ks = 0; ks < ksteps; ++ks, ++it) {
          if (kLoRA && lora_pending && __all_sync(0xffffffffu, mbar_test_wait(ts_full, pend_parity))) issue_lora();
          const int s = it % kStages;
          mbar_wait(&local_full[s], (it / kStages) & 1);
          mbar_wait(&peer_full[s], (it / kStages) & 1);
          tc_fence_after_sync();
          const uint32_t sa = smem_base + s * S::kStageBytes;
          const uint64_t adesc = sw_desc + (sa >> 4);
          const uint64_t bdesc = sw_desc + ((sa + S::kABytes) >> 4);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, kIdescMain, (ks | k) != 0);
            umma_commit_mask(&empty[s], 0xF);
          }
          __syncwarp();
        }
        if (kLoRA && lora_pending) {             // the previous tile's update must precede this tile's hand-over
          mbar_wait(ts_full, pend_parity);
          issue_lora();
        }
        if (elect_one_sync()) umma_commit_mask(&tmem_full[buf], pair_mask);
        __syncwarp();
        if (kLoRA) {
          lora_pending = true;
          pend_buf = buf;
          pend_parity = lt & 1;
        }
      }
      if (kLoRA && lora_pending) {
        mbar_wait(ts_full, pend_parity);
        issue_lora();
      }
    }
  } else if (warp >= kPairEpiWarp0) {
    // ============================================================ epilogue (both CTAs)
    const int q = warp - kPairEpiWarp0;          // TMEM lane quarter == warp % 4
    const int row = q * 32 + lane;
    const int et = threadIdx.x - kPairEpiWarp0 * 32;
    const uint32_t ts_bar = mapa_shared(smem_u32(ts_full), leader_rank);
    uint8_t* slab0 = smem + S::kOffOut + q * S::kSlabs * S::kSlabBytes;
    uint64_t* my_res = res_full + q * S::kSlabs;
    constexpr int kChunks = BN / 64;             // 64-column output chunks per tile
    const bool has_res = p.residual != nullptr;
    uint32_t lt = 0, chunk_no = 0;
    // Residual: chunk c of this warp's chunk sequence is fetched by TMA into slab c % kSlabs two chunks ahead of its use;
    // the epilogue adds the accumulators in place and stores the slab.  (A per-thread row read of the residual stalled the
    // epilogue on global-load latency: +55 us on a 768 -> 2048 call, profiles/r1c_pair_v1.txt.)
    auto prefetch_res = [&](uint32_t c) {
      const int t = cluster_id + static_cast<int>(c / kChunks) * num_clusters;
      if (t >= num_tiles) return;
      const int g = static_cast<int>(c % kChunks);
      const int rm0 = (t / p.num_n_tiles) * 512 + static_cast<int>(pair) * 256 + static_cast<int>(half) * 128 + q * 32;
      const int ro0 = (t % p.num_n_tiles) * BN + g * 64;
      const uint32_t sl = c % S::kSlabs;
      mbar_arrive_expect_tx(&my_res[sl], S::kSlabBytes);
      tma_load_2d(&tmR, &my_res[sl], slab0 + sl * S::kSlabBytes, ro0, rm0);
    };
    if (has_res && lane == 0) {
      prefetch_res(0);
      prefetch_res(1);
    }
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++lt) {
      const int m0 = (tile / p.num_n_tiles) * 512 + static_cast<int>(pair) * 256 + static_cast<int>(half) * 128;
      const int nt = tile % p.num_n_tiles;
      const int o0 = nt * BN;
      const uint32_t buf = lt & 1;
      const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * 256;
      const int m = m0 + row;

      for (int i = et; i < BN; i += 128)
        s_bias[i] = (p.bias != nullptr && o0 + i < p.NO) ? __bfloat162float(p.bias[o0 + i]) : 0.f;
      if (kLoRA) {
        // this CTA's half of Q: rows o0 + rank*BN/2 + r  -> K-major no-swizzle: (r/8)*256 + c*128 + (r%8)*16
        const uint32_t q_s = smem_u32(smem + S::kOffQ);
        const int qrow0 = o0 + static_cast<int>(half) * (BN / 2);
        for (int i = et; i < BN; i += 128) {       // BN/2 rows x 2 chunks
          const int r = i >> 1, c = i & 1;
          uint4 v = make_uint4(0, 0, 0, 0);
          if (qrow0 + r < p.NO) v = *reinterpret_cast<const uint4*>(p.q_rows + static_cast<size_t>(qrow0 + r) * kPairRank + c * 8);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(q_s + (r >> 3) * 256 + c * 128 + (r & 7) * 16),
                       "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                       : "memory");
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");

      mbar_wait(&tmem_full[buf], (lt >> 1) & 1);
      tc_fence_after_sync();
      if (kLoRA) {
        uint32_t t[16];
        tmem_ld16(t_lane + BN, t);
        tmem_wait_ld();
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          pk[i] = pack_bf16x2(__uint_as_float(t[2 * i]) * p.scale, __uint_as_float(t[2 * i + 1]) * p.scale);
        const uint32_t ts_s = smem_u32(smem + S::kOffTs) + (row >> 3) * 256 + (row & 7) * 16;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ts_s), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ts_s + 128), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
        if (nt == 0 && m < p.M && p.side != nullptr) {
          unsigned short* dst = reinterpret_cast<unsigned short*>(p.side) + m;     // lanes = consecutive m: coalesced
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            dst[static_cast<size_t>(2 * i) * p.ld_side] = static_cast<unsigned short>(pk[i] & 0xffffu);
            dst[static_cast<size_t>(2 * i + 1) * p.ld_side] = static_cast<unsigned short>(pk[i] >> 16);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(ts_bar);
        mbar_wait(&tmem_full2[buf], (lt >> 1) & 1);
        tc_fence_after_sync();
      }
#pragma unroll 1
      for (int g = 0; g < kChunks; ++g, ++chunk_no) {
        const uint32_t sl = chunk_no % S::kSlabs;
        uint8_t* slab = slab0 + sl * S::kSlabBytes;
        if (lane == 0) {
          tma_store_wait_read<1>();                // the slab of chunk_no - 2 (= chunk_no + 2 mod 4) is free again
          if (has_res) prefetch_res(chunk_no + 2);
        }
        __syncwarp();
        if (has_res) mbar_wait(&my_res[sl], (chunk_no / S::kSlabs) & 1);
        const uint32_t srow = smem_u32(slab) + lane * 128;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = g * 2 + h;
          uint32_t v[32];
          tmem_ld32(t_lane + c * 32, v);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t chunk = static_cast<uint32_t>((h * 4 + j) ^ (lane & 7));
            uint32_t rr[4] = {0, 0, 0, 0};
            if (has_res)
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]) : "r"(srow + chunk * 16));
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a = __uint_as_float(v[j * 8 + 2 * e]) + s_bias[c * 32 + j * 8 + 2 * e] + bf16lo(rr[e]);
              const float b = __uint_as_float(v[j * 8 + 2 * e + 1]) + s_bias[c * 32 + j * 8 + 2 * e + 1] + bf16hi(rr[e]);
              o[e] = pack_bf16x2(a, b);
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + chunk * 16), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
          }
        }
        if (g == kChunks - 1) {
          // every TMEM read of this tile has retired: hand the accumulator buffer back before the last store goes out
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[buf]), leader_rank));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmD, slab, o0 + g * 64, m0 + q * 32);
          tma_store_commit();
        }
      }
      // s_bias / Q / Ts are rewritten for the next tile only after every epilogue thread is done with this one
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
}

}  // namespace vpt
