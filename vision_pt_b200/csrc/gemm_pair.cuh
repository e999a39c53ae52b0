// LoRA GEMM on CTA pairs (tcgen05 cta_group::2): the large-M (training) route of the fused NF4-LoRA linear.
//
//   D[M, NO] = A[M, R] * Bw[NO, R]^T + bias + Ts * Q[NO, 16]^T (+ residual),   Ts = bf16(scale * A * P[16, R]^T)
//
// forward :  A = x,  Bw = dequantised weight [N, K],            P = lora_down [16, K],   Q = lora_up [N, 16]
// backward:  A = dy, Bw = dequantised weight transposed [K, N], P = lora_up^T [16, N],   Q = lora_down^T [K, 16]
// (src/modules/quant/bnb.py:37-129 + src/modules/peft/lora.py:92-104 of the reference and their autograd).  The weight
// arrives as the bf16 workspace the NF4 dequantiser filled for this call (nf4.cuh), so every operand is K-major and the
// two directions are the same kernel.
//
// Why pairs: with one CTA per 128x192 tile the main loop pulls 104 B per SM-cycle out of L2 and measured 49 % tensor-pipe
// activity at 48 % L2 throughput (profiles/r1_gemm_1cta_ncu.txt).  A pair computes a 256 x BN tile: each CTA loads its own 128
// rows of A and HALF of the B rows; the tensor cores of both SMs read both halves.  L2 bytes per FLOP drop by 1.45x.
//
// Per CTA, 12 warps: warp 0 TMA producer, warp 1 MMA issuer (leader CTA only) + TMEM allocator, warps 4-11 epilogue
// (TMEM -> registers -> 128B-swizzled smem slabs -> TMA store): TWO warps per TMEM lane quarter, each taking 32 of a
// chunk's 64 columns, so an epilogue that does real work per element (the SwiGLU modes below) still hides behind the next
// tile's main loop.  Accumulators are double buffered in TMEM (2 x 256 columns); the 16 LoRA columns ride in the same UMMA
// (N = BN + 16), the rank-16 update is one more K=16 UMMA issued while the next tile's main loop is already running.
//
// Epilogue modes (kEpi), all computed in fp32 on the accumulators with the reference's bf16 rounding points:
//   0  D = acc + bias (+ residual)                                          every linear
//   1  SwiGLU forward in the w_2 GEMM (reference SwiGLU.forward, jit/denoiser.py:498-506): in1 = g = w_1(x),
//      u = bf16(acc + bias);  D = bf16(bf16(silu(g)) * u),  D2 = u           -- swiglu_fwd_kernel disappears
//   2  SwiGLU backward in the w_3 backward-dX GEMM: in1 = g, in2 = u, da = bf16(acc);
//      D = dg = da * u * silu'(g),  D2 = du = da * silu(g)                   -- swiglu_bwd_kernel and the da round trip disappear
#pragma once
#include "sm100.cuh"

namespace vpt {

constexpr int kPairThreads = 384;
constexpr int kPairEpiWarp0 = 4;
constexpr int kPairEpiWarps = 8;
constexpr int kPairRank = 16;

struct PairParams {
  int M, NO, R;
  const __nv_bfloat16* bias;       // [NO] or nullptr
  const __nv_bfloat16* residual;   // in1: residual (mode 0) / g (modes 1, 2): [M, NO] pitch ldr, or nullptr
  int ldr;
  const __nv_bfloat16* q_rows;     // [NO, 16] second-phase LoRA operand
  float scale;
  __nv_bfloat16* side;             // Ts^T [16, ld_side] or nullptr (the layout lora_grad reads by TMA)
  long ld_side;
  int num_m_pairs, num_n_tiles;
  // Several linears that share their input as ONE GEMM (q | k | v: reference Attention.forward, jit/denoiser.py:351-363):
  // the output columns are `sec_n`-wide sections, each with its own LoRA pair -- P holds 16 rows per section, a tile takes
  // the rows of the section it lies in (tiles never straddle sections: BN divides sec_n), `side` holds 16 rows per section.
  // 0 = one section.
  int sec_n;
};

template <int BN, bool kLoRA, int kEpi = 0>
struct PairSmem {
  static constexpr int kNT = BN + (kLoRA ? kPairRank : 0);   // UMMA N
  static constexpr int kNH = kNT / 2;                         // B rows held by each CTA
  static constexpr int kABytes = 128 * 128;
  static constexpr int kBBytes = (kNH * 128 + 1023) / 1024 * 1024;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSlabBytes = 32 * 128;                 // [32 rows x 64 cols] bf16 slab of one lane quarter (two warps)
#ifndef VPT_PAIR_MAX_STAGES
#define VPT_PAIR_MAX_STAGES 6
#endif
  // slab rings per lane quarter: ring 0 = in1 -> D (in place), ring 1 = in2 / D2 of the SwiGLU modes.  An input is fetched
  // kSlabs0 - 2 chunks ahead of its use (a slab lives from its prefetch to the end of its store: distance + 2 chunks).
  // Mode 1's ring 1 only carries an output (u), so two slabs do; that leaves ring 0 its four slabs = distance 2, like
  // mode 0.  Mode 2 has inputs in both rings: 3 + 3 (distance 1) keeps four pipeline stages, 4 + 4 would leave three.
#ifndef VPT_PAIR_M2_SLABS
#define VPT_PAIR_M2_SLABS 3
#endif
  static constexpr int kRings = kEpi == 0 ? 1 : 2;
  static constexpr int kSlabs0 = kEpi == 2 ? VPT_PAIR_M2_SLABS : 4;
  static constexpr int kSlabs1 = kEpi == 0 ? 0 : (kEpi == 1 ? 2 : VPT_PAIR_M2_SLABS);
  static constexpr int kOutBytes = 4 * (kSlabs0 + kSlabs1) * kSlabBytes;
  static constexpr int kTsBytes = 128 * 32;
  static constexpr int kQBytes = (BN / 2) * 32;
  static constexpr int kBiasBytes = BN * 4;
  static constexpr int kFixed = kOutBytes + kTsBytes + kQBytes + kBiasBytes + 512 + 1024;
  static constexpr int kStagesRaw = (227 * 1024 - kFixed) / kStageBytes;
  static constexpr int kStages = kStagesRaw > VPT_PAIR_MAX_STAGES ? VPT_PAIR_MAX_STAGES : kStagesRaw;
  static constexpr int kOffOut = kStages * kStageBytes;
  static constexpr int kOffTs = kOffOut + kOutBytes;
  static constexpr int kOffQ = kOffTs + kTsBytes;
  static constexpr int kOffBias = kOffQ + kQBytes;
  static constexpr int kOffBars = kOffBias + kBiasBytes;
  static constexpr int kNumBars = 2 * kStages + 7 + 4 * (kSlabs0 + kSlabs1);
  static constexpr int kOffTmemSlot = kOffBars + kNumBars * 8;
  static constexpr int kTotal = kOffTmemSlot + 16 + 1024;
  static_assert(kStages >= 3, "pipeline too shallow");
  static_assert(kNumBars * 8 <= 512, "barrier area");
};

// sigmoid on the special-function unit: ex2 + rcp, both approximate to ~1 ulp of fp32 -- invisible after the bf16 rounding
// that follows.  (A true divide and __float2bfloat16_rn per element put 4 quarter-rate F2F / MUFU instructions on every
// element and made the K = 768 GEMMs epilogue-bound; rounding goes through the full-rate bf16x2 pack instead.)
__device__ __forceinline__ float fast_sigmoid(float x) {
  float r;
  const float e = fast_exp2(x * -1.4426950408889634f);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return r;
}

// tmA : A [M, R],           box {64, 128}
// tmB0: Bw [NO, R],         box {64, kNH}            rows o0 .. o0 + kNH            (CTA 0)
// tmB1: Bw [NO, R],         box {64, kNH - 16 | kNH} rows o0 + kNH .. o0 + BN       (CTA 1)
// tmP : P [16, R],          box {64, 16}             appended below CTA 1's weight rows
// tmD : D [M, NO],          box {64, 32}             (store)
// tmR : in1 [M, NO],        box {64, 32}             (prefetched into ring 0)
// tmD2: D2 [M, NO],         box {64, 32}             (store, modes 1 and 2)
// tmR2: in2 [M, NO],        box {64, 32}             (prefetched into ring 1, mode 2)      all SWIZZLE_128B
template <int BN, bool kLoRA, int kEpi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB0,
                 const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmP,
                 const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmR,
                 const __grid_constant__ CUtensorMap tmD2, const __grid_constant__ CUtensorMap tmR2,
                 const PairParams p) {
  using S = PairSmem<BN, kLoRA, kEpi>;
  static_assert(BN % 64 == 0 && BN >= 64 && S::kNT <= 256, "unsupported BN");
  constexpr int kStages = S::kStages;
  constexpr uint32_t kIdescMain = umma_idesc_bf16(256, S::kNT, 0, 0);
  constexpr uint32_t kIdescLora = umma_idesc_bf16(256, BN, 0, 0);
  constexpr uint32_t kStageTx = S::kABytes + S::kNH * 128;     // bytes each CTA's loads credit per stage

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOffBars);
  uint64_t* full = bars;                       // [kStages] leader's copy is the live one (count 2 + both CTAs' bytes)
  uint64_t* empty = bars + kStages;            // [kStages] per CTA, arrived by the multicast commit
  uint64_t* tmem_full = bars + 2 * kStages;    // [2] per CTA: main loop of a tile finished
  uint64_t* tmem_full2 = tmem_full + 2;        // [2] per CTA: rank-16 update finished
  uint64_t* tmem_empty = tmem_full2 + 2;       // [2] leader's: 16 epilogue warps (both CTAs) drained the buffer
  uint64_t* ts_full = tmem_empty + 2;          // [1] leader's: 8 staging warps (both CTAs) staged Ts / Q
  uint64_t* res_full = ts_full + 1;            // [4 quarters][kSlabs0 + kSlabs1] prefetched input slab landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::kOffTmemSlot);
  float* s_bias = reinterpret_cast<float*>(smem + S::kOffBias);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int num_tiles = p.num_m_pairs * p.num_n_tiles;
  const int ksteps = (p.R + 63) / 64;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 2);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_full2[b], 1);
      mbar_init(&tmem_empty[b], 2 * kPairEpiWarps);
    }
    mbar_init(ts_full, 8);
    for (int i = 0; i < 4 * (S::kSlabs0 + S::kSlabs1); ++i) mbar_init(&res_full[i], 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(leader ? &tmB0 : &tmB1);
    if (kLoRA && !leader) tma_prefetch_desc(&tmP);
    tma_prefetch_desc(&tmD);
    if (p.residual != nullptr) tma_prefetch_desc(&tmR);
    if (kEpi != 0) tma_prefetch_desc(&tmD2);
    if (kEpi == 2) tma_prefetch_desc(&tmR2);
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
    // ============================================================ TMA producer (both CTAs)
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m0 = (tile / p.num_n_tiles) * 256 + static_cast<int>(rank) * 128;
        const int o0 = (tile % p.num_n_tiles) * BN;
        for (int ks = 0; ks < ksteps; ++ks, ++it) {
          const int s = it % kStages;
          mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);
          const uint32_t bar = mapa_shared(smem_u32(&full[s]), 0);
          uint8_t* sa = smem + s * S::kStageBytes;
          uint8_t* sb = sa + S::kABytes;
          if (leader) {
            mbar_arrive_expect_tx(&full[s], 2 * kStageTx);
            tma_load_2d_pair(&tmA, bar, sa, ks * 64, m0);
            tma_load_2d_pair(&tmB0, bar, sb, ks * 64, o0);
          } else {
            mbar_arrive_cluster(bar);
            tma_load_2d_pair(&tmA, bar, sa, ks * 64, m0);
            tma_load_2d_pair(&tmB1, bar, sb, ks * 64, o0 + S::kNH);
            if (kLoRA) tma_load_2d_pair(&tmP, bar, sb + (S::kNH - kPairRank) * 128, ks * 64, p.sec_n > 0 ? (o0 / p.sec_n) * kPairRank : 0);
          }
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
          umma_commit_pair(&tmem_full2[pend_buf]);
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
          mbar_wait(&full[s], (it / kStages) & 1);
          tc_fence_after_sync();
          const uint32_t sa = smem_base + s * S::kStageBytes;
          const uint64_t adesc = sw_desc + (sa >> 4);
          const uint64_t bdesc = sw_desc + ((sa + S::kABytes) >> 4);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, kIdescMain, (ks | k) != 0);
            umma_commit_pair(&empty[s]);
          }
          __syncwarp();
        }
        if (kLoRA && lora_pending) {             // the previous tile's update must precede this tile's hand-over
          mbar_wait(ts_full, pend_parity);
          issue_lora();
        }
        if (elect_one_sync()) umma_commit_pair(&tmem_full[buf]);
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
    // ============================================================ epilogue (both CTAs): two warps per TMEM lane quarter
    const int ew = warp - kPairEpiWarp0;
    const int q = ew & 3;                        // TMEM lane quarter == warp % 4
    const int half = ew >> 2;                    // which 32 of a chunk's 64 columns
    const bool pair_lead = half == 0;            // issues this quarter's TMA traffic and stages its Ts rows
    const int row = q * 32 + lane;
    const int et = threadIdx.x - kPairEpiWarp0 * 32;
    const uint32_t ts_bar = mapa_shared(smem_u32(ts_full), 0);
    uint8_t* ring0 = smem + S::kOffOut + q * (S::kSlabs0 + S::kSlabs1) * S::kSlabBytes;
    uint8_t* ring1 = ring0 + S::kSlabs0 * S::kSlabBytes;                     // only when kRings == 2
    uint64_t* res0 = res_full + q * (S::kSlabs0 + S::kSlabs1);
    uint64_t* res1 = res0 + S::kSlabs0;
    constexpr int kChunks = BN / 64;             // 64-column output chunks per tile
    constexpr int kAhead = S::kSlabs0 - 2;       // input prefetch distance in chunks
    constexpr int kS1 = S::kSlabs1 > 0 ? S::kSlabs1 : 1;
    const bool has_in1 = kEpi != 0 || p.residual != nullptr;
    constexpr bool kIn2 = kEpi == 2;
    constexpr bool kOut2 = kEpi != 0;
    uint32_t lt = 0, chunk_no = 0;
    // Inputs: chunk c of this quarter's chunk sequence is fetched by TMA into slab c % kSlabs kAhead chunks ahead of its use;
    // the epilogue combines it with the accumulators in place and stores the slab.  (A per-thread row read of the residual
    // stalled the epilogue on global-load latency: +55 us on a 768 -> 2048 call, profiles/r1c_pair_v1.txt.)
    auto prefetch_in = [&](uint32_t c) {
      const int t = cluster_id + static_cast<int>(c / kChunks) * num_clusters;
      if (t >= num_tiles) return;
      const int g = static_cast<int>(c % kChunks);
      const int rm0 = (t / p.num_n_tiles) * 256 + static_cast<int>(rank) * 128 + q * 32;
      const int ro0 = (t % p.num_n_tiles) * BN + g * 64;
      const uint32_t sl = c % S::kSlabs0;
      mbar_arrive_expect_tx(&res0[sl], S::kSlabBytes);
      tma_load_2d(&tmR, &res0[sl], ring0 + sl * S::kSlabBytes, ro0, rm0);
      if (kIn2) {
        const uint32_t sl1 = c % kS1;
        mbar_arrive_expect_tx(&res1[sl1], S::kSlabBytes);
        tma_load_2d(&tmR2, &res1[sl1], ring1 + sl1 * S::kSlabBytes, ro0, rm0);
      }
    };
    if (has_in1 && pair_lead && lane == 0) {
#pragma unroll
      for (int c = 0; c < kAhead; ++c) prefetch_in(c);
    }
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++lt) {
      const int m0 = (tile / p.num_n_tiles) * 256 + static_cast<int>(rank) * 128;
      const int nt = tile % p.num_n_tiles;
      const int o0 = nt * BN;
      const uint32_t buf = lt & 1;
      const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * 256;
      const int m = m0 + row;

      for (int i = et; i < BN; i += 32 * kPairEpiWarps)
        s_bias[i] = (p.bias != nullptr && o0 + i < p.NO) ? __bfloat162float(p.bias[o0 + i]) : 0.f;
      if (kLoRA) {
        // this CTA's half of Q: rows o0 + rank*BN/2 + r  -> K-major no-swizzle: (r/8)*256 + c*128 + (r%8)*16
        const uint32_t q_s = smem_u32(smem + S::kOffQ);
        const int qrow0 = o0 + static_cast<int>(rank) * (BN / 2);
        for (int i = et; i < BN; i += 32 * kPairEpiWarps) {       // BN/2 rows x 2 chunks
          const int r = i >> 1, c = i & 1;
          uint4 v = make_uint4(0, 0, 0, 0);
          if (qrow0 + r < p.NO) v = *reinterpret_cast<const uint4*>(p.q_rows + static_cast<size_t>(qrow0 + r) * kPairRank + c * 8);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(q_s + (r >> 3) * 256 + c * 128 + (r & 7) * 16),
                       "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                       : "memory");
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");

      mbar_wait(&tmem_full[buf], (lt >> 1) & 1);
      tc_fence_after_sync();
      if (kLoRA) {
        if (pair_lead) {
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
          const int sec = p.sec_n > 0 ? o0 / p.sec_n : 0;
          const bool first_of_section = p.sec_n > 0 ? (o0 - sec * p.sec_n == 0) : (nt == 0);
          if (first_of_section && m < p.M && p.side != nullptr) {
            unsigned short* dst = reinterpret_cast<unsigned short*>(p.side) + static_cast<size_t>(sec) * kPairRank * p.ld_side + m;   // lanes = consecutive m: coalesced
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
        }
        mbar_wait(&tmem_full2[buf], (lt >> 1) & 1);
        tc_fence_after_sync();
      }
#pragma unroll 1
      for (int g = 0; g < kChunks; ++g, ++chunk_no) {
        const uint32_t sl = chunk_no % S::kSlabs0;
        const uint32_t sl1 = chunk_no % kS1;
        uint8_t* slab = ring0 + sl * S::kSlabBytes;
        uint8_t* slab2 = ring1 + sl1 * S::kSlabBytes;
        if (pair_lead && lane == 0) {
          tma_store_wait_read<1>();                // the slabs of chunk_no - 2 (= chunk_no + kAhead mod kSlabs0) are free again
          if (has_in1) prefetch_in(chunk_no + kAhead);
        }
        __syncwarp();
        // a two-slab ring is rewritten two chunks after its store was issued: the partner warp may only write once the
        // leader has seen that store drain (deeper rings get this from the previous chunk's barrier)
        if (kOut2 && S::kSlabs1 < 3) asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
        if (has_in1) mbar_wait(&res0[sl], (chunk_no / S::kSlabs0) & 1);
        if (kIn2) mbar_wait(&res1[sl1], (chunk_no / kS1) & 1);
        const uint32_t srow = smem_u32(slab) + lane * 128;
        const uint32_t srow2 = smem_u32(slab2) + lane * 128;
        const int c = g * 2 + half;                // this warp's 32-column group of the tile
        uint32_t v[32];
        tmem_ld32(t_lane + c * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t chunk = static_cast<uint32_t>((half * 4 + j) ^ (lane & 7));
          uint32_t rr[4] = {0, 0, 0, 0}, r2[4] = {0, 0, 0, 0};
          if (has_in1)
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]) : "r"(srow + chunk * 16));
          if (kIn2)
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r2[0]), "=r"(r2[1]), "=r"(r2[2]), "=r"(r2[3]) : "r"(srow2 + chunk * 16));
          uint32_t o[4], o2[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float acc_a = __uint_as_float(v[j * 8 + 2 * e]) + s_bias[c * 32 + j * 8 + 2 * e];
            const float acc_b = __uint_as_float(v[j * 8 + 2 * e + 1]) + s_bias[c * 32 + j * 8 + 2 * e + 1];
            if (kEpi == 0) {
              o[e] = pack_bf16x2(acc_a + bf16lo(rr[e]), acc_b + bf16hi(rr[e]));
            } else if (kEpi == 1) {
              // u = bf16(acc); a = bf16( bf16(silu(g)) * u )
              const float ga = bf16lo(rr[e]), gb = bf16hi(rr[e]);
              const uint32_t u2 = pack_bf16x2(acc_a, acc_b);
              const uint32_t s2 = pack_bf16x2(ga * fast_sigmoid(ga), gb * fast_sigmoid(gb));
              o[e] = pack_bf16x2(bf16lo(s2) * bf16lo(u2), bf16hi(s2) * bf16hi(u2));
              o2[e] = u2;
            } else {
              // da = bf16(acc); dg = da * u * silu'(g), du = da * silu(g)
              const float ga = bf16lo(rr[e]), gb = bf16hi(rr[e]);
              const float ua = bf16lo(r2[e]), ub = bf16hi(r2[e]);
              const uint32_t d2 = pack_bf16x2(acc_a, acc_b);
              const float da = bf16lo(d2), db = bf16hi(d2);
              const float sga = fast_sigmoid(ga), sgb = fast_sigmoid(gb);
              o[e] = pack_bf16x2(da * ua * (sga * (1.f + ga * (1.f - sga))), db * ub * (sgb * (1.f + gb * (1.f - sgb))));
              o2[e] = pack_bf16x2(da * (ga * sga), db * (gb * sgb));
            }
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + chunk * 16), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
          if (kOut2)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow2 + chunk * 16), "r"(o2[0]), "r"(o2[1]), "r"(o2[2]), "r"(o2[3]) : "memory");
        }
        if (g == kChunks - 1) {
          // every TMEM read of this tile by this warp has retired: hand the accumulator buffer back before the last store
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[buf]), 0));
        }
        fence_proxy_async_smem();
        // both warps of the quarter have written their halves of the slab(s)
        asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
        if (pair_lead && lane == 0) {
          tma_store_2d(&tmD, slab, o0 + g * 64, m0 + q * 32);
          if (kOut2) tma_store_2d(&tmD2, slab2, o0 + g * 64, m0 + q * 32);
          tma_store_commit();
        }
      }
      // s_bias / Q / Ts are rewritten for the next tile only after every epilogue thread is done with this one
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    if (pair_lead && lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
}

}  // namespace vpt
