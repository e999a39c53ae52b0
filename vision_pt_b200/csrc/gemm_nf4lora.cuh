// Fused NF4-dequant + LoRA GEMM for sm_100a (tcgen05 / TMEM / TMA), forward and backward-dX.
//
// Replaces, for one Linear of the reference, the chain
//   bitsandbytes dequantize_blockwise + dequantize_4bit + F.linear      (src/modules/quant/bnb.py:37-129, inherited
//                                                                         bnb.nn.Linear4bit.forward -> MatMul4Bit)
//   LoRALinear.forward: lora_down, lora_up, *alpha/rank, +               (src/modules/peft/lora.py:92-104)
// and the autograd of both for the activation gradient.
//
// One persistent CTA per SM, 16 warps:
//   warp 0      TMA producer: activation tiles (and lora_down rows / bf16 weights when those come by TMA)
//   warp 1      tcgen05.mma issuer (one lane)
//   warp 2      TMEM allocator
//   warps 4-7   epilogue: TMEM -> registers -> global; also builds the rank-16 LoRA operand between the two MMA phases
//   warps 8-15  NF4 dequant producers: packed nibbles + double-quant absmax -> bf16 -> 128B-swizzled smem B tile
//
// MMA view (both directions):  D[M, NO] = A[M, R] * B,  accumulators in TMEM (2 x 256 columns, double buffered)
//   fwd : A = X  [M,K],  B rows = weight rows n (K-major, 64 k per stage)          NO = N, R = K
//   bwd : A = dY [M,N],  B rows = weight rows n (MN-major, 64 n per stage)         NO = K, R = N
// LoRA is folded in as 16 extra accumulator columns in the main loop (T = X*A_down^T  resp.  dT = dY*B_up) and one
// extra K=16 MMA per tile ( += Ts * B_up^T  resp.  += dTs * A_down ), Ts = bf16(scale * T).
#pragma once
#include "sm100.cuh"

namespace vpt {

constexpr int kBM = 128;       // rows of D per tile (UMMA M)
constexpr int kBK = 64;        // reduction elements per pipeline stage (one 128B swizzle row of bf16)
constexpr int kRank = 16;      // LoRA rank handled by the fused path
constexpr int kGemmThreads = 512;
constexpr int kDeqWarps = 8;
constexpr int kDeqThreads = kDeqWarps * 32;
constexpr int kEpiWarp0 = 4;
constexpr int kDeqWarp0 = 8;

struct Nf4Weight {
  const uint8_t* packed;        // [(N*K+1)/2] two codes per byte, high nibble = even element
  const uint8_t* qabsmax;       // [N*K/64] 8-bit codes of (absmax - offset)
  const float* nested_absmax;   // [ceil(N*K/64/256)]
  const float* nested_code;     // [256]
  const float* code;            // [16]
  float offset;
  int N, K;                     // out_features, in_features
  // ragged layout (K % 64 != 0), produced once by vpt_nf4_repack: row-aligned codes + decoded statistics
  const uint8_t* packed_rows;   // [N, K_pad / 2]
  const float* absmax_f32;      // [ceil(N*K/64)] fl32(fl32(code2[q]*nested) + offset)
  int K_pad;
};

struct GemmParams {
  int M, NO, R;                 // D[M,NO] = A[M,R] * B
  __nv_bfloat16* D;
  int ldd;
  const __nv_bfloat16* bias;    // [NO] or nullptr
  const __nv_bfloat16* residual;  // [M, NO] (pitch ldr) added in the epilogue, or nullptr
  int ldr;
  Nf4Weight w;
  const __nv_bfloat16* lora_down;  // [16, K] with row pitch ld_down (multiple of 8 elements, zero padded)
  int ld_down;
  const __nv_bfloat16* lora_up;    // [N, 16]
  float scale;                  // alpha / rank
  __nv_bfloat16* side;          // [16, ld_side] (transposed): fwd Ts = bf16(scale * X A_down^T); bwd dTs = bf16(scale * dY B_up)
  long ld_side;
  int num_m_tiles, num_n_tiles;
  // descriptor stride overrides (bytes, 0 = default); only the bring-up probe sets them
  uint32_t dbg_b_lbo, dbg_b_sbo, dbg_q_lbo, dbg_q_sbo, dbg_ts_lbo, dbg_ts_sbo;
};

template <int BN, bool kBwd, bool kLoRA>
struct GemmSmem {
  static constexpr int kStages = 4;
  static constexpr int kABytes = kBM * 128;
  static constexpr int kBBytes = kBwd ? (BN / 64 + (kLoRA ? 1 : 0)) * 8192 : ((BN + (kLoRA ? kRank : 0)) * 128 + 1023) / 1024 * 1024;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTsBytes = kBM * 32;         // [128 x 16] bf16, K-major no-swizzle
  static constexpr int kQBytes = BN * 32;           // second-phase LoRA operand
  static constexpr int kBiasBytes = BN * 4;
  static constexpr int kCodeBytes = (16 + 256) * 4; // NF4 code + nested code tables
  static constexpr int kOffTs = kStages * kStageBytes;
  static constexpr int kOffQ = kOffTs + kTsBytes;
  static constexpr int kOffBias = kOffQ + kQBytes;
  static constexpr int kOffCode = (kOffBias + kBiasBytes + 63) / 64 * 64;
  static constexpr int kOffBars = kOffCode + kCodeBytes;
  static constexpr int kNumBars = 2 * kStages + 2 + 2 + 2 + 1;
  static constexpr int kOffTmemSlot = kOffBars + kNumBars * 8;
  static constexpr int kTotal = kOffTmemSlot + 16 + 1024;  // + slack for 1024B alignment of the base
};

// Dequantises 32 consecutive weights (one uint4 of packed codes) and writes them as four 16B chunks of a
// 128B-swizzled row.  w = bf16_rn( fl32( code[nibble] * absmax ) ): the same two roundings as bitsandbytes'
// kDequantizeBlockwise<bf16, NF4>.
__device__ __forceinline__ float lds_f32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
}
// `code_saddr` = shared address of the 16-entry fp32 code table, 64-byte aligned so that "| base" replaces "+ base".
__device__ __forceinline__ void nf4_dequant32_to_swizzled(const uint4& pk, float am, uint32_t code_saddr,
                                                          uint32_t row_saddr, int half, int row_in_atom) {
  const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t x = w[i];
    uint32_t o[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      // byte b holds elements (2b: high nibble, 2b+1: low nibble); table offset = nibble * 4
      const uint32_t a_hi = ((x >> (8 * b + 2)) & 0x3cu) | code_saddr;
      const uint32_t a_lo = (b == 0 ? ((x << 2) & 0x3cu) : ((x >> (8 * b - 2)) & 0x3cu)) | code_saddr;
      const float hi = __fmul_rn(lds_f32(a_hi), am);   // even element
      const float lo = __fmul_rn(lds_f32(a_lo), am);   // odd element
      o[b] = pack_bf16x2(hi, lo);
    }
    const uint32_t chunk = static_cast<uint32_t>((half * 4 + i) ^ row_in_atom);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_saddr + chunk * 16u), "r"(o[0]), "r"(o[1]),
                 "r"(o[2]), "r"(o[3])
                 : "memory");
  }
}

// Ragged rows: the 32 elements may straddle one 64-block boundary of the flattened weight; element e uses am_a when
// e < split, am_b otherwise (same products, hence the same bits, as the flat layout).
__device__ __forceinline__ void nf4_dequant32_split_to_swizzled(const uint4& pk, float am_a, float am_b, int split,
                                                                uint32_t code_saddr, uint32_t row_saddr, int half,
                                                                int row_in_atom) {
  const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t x = w[i];
    uint32_t o[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int e = i * 8 + b * 2;
      const uint32_t a_hi = ((x >> (8 * b + 2)) & 0x3cu) | code_saddr;
      const uint32_t a_lo = (b == 0 ? ((x << 2) & 0x3cu) : ((x >> (8 * b - 2)) & 0x3cu)) | code_saddr;
      const float hi = __fmul_rn(lds_f32(a_hi), e < split ? am_a : am_b);
      const float lo = __fmul_rn(lds_f32(a_lo), e + 1 < split ? am_a : am_b);
      o[b] = pack_bf16x2(hi, lo);
    }
    const uint32_t chunk = static_cast<uint32_t>((half * 4 + i) ^ row_in_atom);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_saddr + chunk * 16u), "r"(o[0]), "r"(o[1]),
                 "r"(o[2]), "r"(o[3])
                 : "memory");
  }
}

__device__ __forceinline__ void st_shared_zero16(uint32_t saddr) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(saddr), "r"(0u) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t saddr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// tmA : activations [M, R] row-major, box {64, 128}, SWIZZLE_128B
// tmB : (kNF4 == false) bf16 weight [N, K] row-major; fwd box {64, BN}, bwd box {64, 64}; SWIZZLE_128B
// tmP : (fwd, kLoRA) lora_down [16, K], box {64, 16}, SWIZZLE_128B
template <int BN, bool kBwd, bool kNF4, bool kLoRA, bool kRagged = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_nf4lora_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmP, const GemmParams p) {
  using S = GemmSmem<BN, kBwd, kLoRA>;
  static_assert(BN % 32 == 0 && BN >= 64 && BN + (kLoRA ? kRank : 0) <= 256, "unsupported BN");
  static_assert(!kBwd || BN % 64 == 0, "bwd tiles are built from 64-wide MN blocks");
  constexpr int kStages = S::kStages;
  constexpr int kUmmaN = BN + (kLoRA ? kRank : 0);
  constexpr uint32_t kIdescMain = umma_idesc_bf16(kBM, kUmmaN, 0, kBwd ? 1 : 0);
  constexpr uint32_t kIdescLora = umma_idesc_bf16(kBM, BN, 0, kBwd ? 1 : 0);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOffBars);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;       // [2] main loop of a tile finished
  uint64_t* tmem_full2 = tmem_full + 2;           // [2] LoRA MMA finished
  uint64_t* tmem_empty = tmem_full2 + 2;          // [2] epilogue drained the buffer
  uint64_t* ts_full = tmem_empty + 2;             // [1] Ts/Q operands are in smem
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::kOffTmemSlot);
  float* s_code = reinterpret_cast<float*>(smem + S::kOffCode);
  float* s_bias = reinterpret_cast<float*>(smem + S::kOffBias);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int ksteps = (p.R + kBK - 1) / kBK;
  const bool deq_active = kNF4 || (kBwd && kLoRA);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1 + (deq_active ? kDeqWarps : 0));
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_full2[b], 1);
      mbar_init(&tmem_empty[b], 4);
    }
    mbar_init(ts_full, 128);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    if (!kNF4) tma_prefetch_desc(&tmB);
    if (!kBwd && kLoRA) tma_prefetch_desc(&tmP);
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  if (kNF4) {
    for (int i = threadIdx.x; i < 16 + 256; i += kGemmThreads)
      s_code[i] = i < 16 ? p.w.code[i] : p.w.nested_code[i - 16];
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================================================ TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.num_n_tiles) * kBM;
        const int o0 = (tile % p.num_n_tiles) * BN;
        for (int ks = 0; ks < ksteps; ++ks, ++it) {
          const int s = it % kStages;
          mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);
          uint8_t* sa = smem + s * S::kStageBytes;
          uint8_t* sb = sa + S::kABytes;
          uint32_t bytes = S::kABytes;
          if (!kNF4) bytes += BN * 128;
          if (!kBwd && kLoRA) bytes += kRank * 128;
          mbar_arrive_expect_tx(&full[s], bytes);
          tma_load_2d(&tmA, &full[s], sa, ks * kBK, m0);
          if (!kNF4) {
            if (!kBwd) {
              tma_load_2d(&tmB, &full[s], sb, ks * kBK, o0);
            } else {
              for (int j = 0; j < BN / 64; ++j) tma_load_2d(&tmB, &full[s], sb + j * 8192, o0 + 64 * j, ks * kBK);
            }
          }
          if (!kBwd && kLoRA) tma_load_2d(&tmP, &full[s], sb + BN * 128, ks * kBK, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer
    if (lane == 0) {
      uint32_t it = 0, lt = 0;
      const uint32_t b_lbo = p.dbg_b_lbo ? p.dbg_b_lbo : 8192u, b_sbo = p.dbg_b_sbo ? p.dbg_b_sbo : 1024u;
      const uint32_t q_lbo = p.dbg_q_lbo ? p.dbg_q_lbo : 128u, q_sbo = p.dbg_q_sbo ? p.dbg_q_sbo : 256u;
      const uint32_t ts_lbo = p.dbg_ts_lbo ? p.dbg_ts_lbo : 128u, ts_sbo = p.dbg_ts_sbo ? p.dbg_ts_sbo : 256u;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
        const uint32_t buf = lt & 1;
        const uint32_t d_tmem = tmem_base + buf * 256;
        mbar_wait(&tmem_empty[buf], ((lt >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        for (int ks = 0; ks < ksteps; ++ks, ++it) {
          const int s = it % kStages;
          mbar_wait(&full[s], (it / kStages) & 1);
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(smem + s * S::kStageBytes);
          const uint32_t sb = sa + S::kABytes;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t adesc = umma_smem_desc(sa + k * 32, 16, 1024, kLayoutSW128);
            const uint64_t bdesc = kBwd ? umma_smem_desc(sb + k * 2048, b_lbo, b_sbo, kLayoutSW128)
                                        : umma_smem_desc(sb + k * 32, 16, 1024, kLayoutSW128);
            umma_ss(d_tmem, adesc, bdesc, kIdescMain, (ks | k) != 0);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&tmem_full[buf]);
        if (kLoRA) {
          mbar_wait(ts_full, lt & 1);
          tc_fence_after_sync();
          const uint64_t adesc = umma_smem_desc(smem_u32(smem + S::kOffTs), ts_lbo, ts_sbo, kLayoutNone);
          // fwd: Q = lora_up rows [BN x 16], K-major none (LBO = k-chunk stride, SBO = 8-row group stride)
          // bwd: Q = lora_down cols [16 x BN], MN-major none (LBO = k-group stride, SBO = 8-column group stride)
          const uint64_t bdesc = umma_smem_desc(smem_u32(smem + S::kOffQ), q_lbo, q_sbo, kLayoutNone);
          umma_ss(d_tmem, adesc, bdesc, kIdescLora, 1);
          umma_commit(&tmem_full2[buf]);
        }
      }
    }
  } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + 4) {
    // ============================================================ epilogue
    const int q = warp - kEpiWarp0;          // TMEM lane quarter == warp % 4
    const int row = q * 32 + lane;           // row of the tile owned by this thread
    const int et = threadIdx.x - kEpiWarp0 * 32;
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int m0 = (tile / p.num_n_tiles) * kBM;
      const int nt = tile % p.num_n_tiles;
      const int o0 = nt * BN;
      const uint32_t buf = lt & 1;
      const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * 256;
      const int m = m0 + row;

      // stage bias and the second-phase LoRA operand while the main loop runs
      for (int i = et; i < BN; i += 128)
        s_bias[i] = (p.bias != nullptr && o0 + i < p.NO) ? __bfloat162float(p.bias[o0 + i]) : 0.f;
      if (kLoRA) {
        const uint32_t q_s = smem_u32(smem + S::kOffQ);
        if (!kBwd) {
          // lora_up[o0 + r, 0:16] -> K-major no-swizzle: (r/8)*256 + c*128 + (r%8)*16
          for (int i = et; i < BN * 2; i += 128) {
            const int r = i >> 1, c = i & 1;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (o0 + r < p.NO) v = *reinterpret_cast<const uint4*>(p.lora_up + static_cast<size_t>(o0 + r) * kRank + c * 8);
            st_shared_v4(q_s + (r >> 3) * 256 + c * 128 + (r & 7) * 16, v);
          }
        } else {
          // lora_down[j, o0 + 8*oc .. +7] -> MN-major no-swizzle: oc*256 + (j/8)*128 + (j%8)*16
          for (int i = et; i < (BN / 8) * kRank; i += 128) {
            const int j = i / (BN / 8), oc = i % (BN / 8);
            uint4 v = make_uint4(0, 0, 0, 0);
            if (o0 + oc * 8 < p.NO) v = *reinterpret_cast<const uint4*>(p.lora_down + static_cast<size_t>(j) * p.ld_down + o0 + oc * 8);
            st_shared_v4(q_s + oc * 256 + (j >> 3) * 128 + (j & 7) * 16, v);
          }
        }
      }
      named_bar_sync(1, 128);  // s_bias is read by other threads than the ones that staged it

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
        st_shared_v4(ts_s, make_uint4(pk[0], pk[1], pk[2], pk[3]));
        st_shared_v4(ts_s + 128, make_uint4(pk[4], pk[5], pk[6], pk[7]));
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
        mbar_arrive(ts_full);
        mbar_wait(&tmem_full2[buf], (lt >> 1) & 1);
        tc_fence_after_sync();
      }
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(t_lane + c * 32, v);
        tmem_wait_ld();
        if (m < p.M) {
          __nv_bfloat16* drow = p.D + static_cast<size_t>(m) * p.ldd + o0 + c * 32;
          const __nv_bfloat16* rrow = p.residual ? p.residual + static_cast<size_t>(m) * p.ldr + o0 + c * 32 : nullptr;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = o0 + c * 32 + g * 8;
            uint32_t rr[4] = {0, 0, 0, 0};
            if (rrow != nullptr) {
              if (col + 8 <= p.NO) {
                const uint4 t4 = *reinterpret_cast<const uint4*>(rrow + g * 8);
                rr[0] = t4.x; rr[1] = t4.y; rr[2] = t4.z; rr[3] = t4.w;
              } else {
                for (int e = 0; e < 8 && col + e < p.NO; ++e)
                  rr[e >> 1] |= static_cast<uint32_t>(reinterpret_cast<const unsigned short*>(rrow)[g * 8 + e]) << (16 * (e & 1));
              }
            }
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a = __uint_as_float(v[g * 8 + 2 * e]) + s_bias[c * 32 + g * 8 + 2 * e] + bf16lo(rr[e]);
              const float b = __uint_as_float(v[g * 8 + 2 * e + 1]) + s_bias[c * 32 + g * 8 + 2 * e + 1] + bf16hi(rr[e]);
              o[e] = pack_bf16x2(a, b);
            }
            if (col + 8 <= p.NO) {
              *reinterpret_cast<uint4*>(drow + g * 8) = make_uint4(o[0], o[1], o[2], o[3]);
            } else {
              for (int e = 0; e < 8 && col + e < p.NO; ++e) {
                const uint32_t wv = o[e >> 1];
                const unsigned short hv = (e & 1) ? static_cast<unsigned short>(wv >> 16) : static_cast<unsigned short>(wv & 0xffff);
                reinterpret_cast<unsigned short*>(drow)[g * 8 + e] = hv;
              }
            }
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
      // s_bias / Q / Ts are rewritten for the next tile only after every epilogue thread is done with this one
      named_bar_sync(1, 128);
    }
  } else if (warp >= kDeqWarp0 && deq_active) {
    // ============================================================ weight producers
    const int dt = threadIdx.x - kDeqWarp0 * 32;
    constexpr int kTasks = 2 * BN;                       // 32-element half rows per stage
    constexpr int kPerThread = (kTasks + kDeqThreads - 1) / kDeqThreads;
    const int K = p.w.K;
    uint4 pk[kPerThread];
    uint32_t qa[kPerThread];      // flat layout: absmax code;        ragged: elements before the block boundary
    float nest[kPerThread];       // flat layout: nested absmax;      ragged: absmax of the first block
    float amb[kPerThread];        //                                   ragged: absmax of the second block
    bool valid[kPerThread];
    const long nblocks = (static_cast<long>(p.w.N) * K + 63) >> 6;

    // (tile, kstep) -> global coordinates of this thread's tasks
    auto prefetch = [&](int tile, int ks) {
      const int o0 = (tile % p.num_n_tiles) * BN;
#pragma unroll
      for (int u = 0; u < kPerThread; ++u) {
        const int t = dt + u * kDeqThreads;
        valid[u] = false;
        pk[u] = make_uint4(0, 0, 0, 0);
        qa[u] = 0;
        nest[u] = 0.f;
        amb[u] = 0.f;
        if (!kNF4 || t >= kTasks) continue;
        int wrow, wcol;                                  // weight row (n) and first column (k) of the 32 elements
        if (!kBwd) {
          wrow = o0 + (t >> 1);
          wcol = ks * kBK + (t & 1) * 32;
        } else {
          wrow = ks * kBK + ((t & 127) >> 1);
          wcol = o0 + (t >> 7) * 64 + (t & 1) * 32;
        }
        if (kRagged) {
          if (wrow < p.w.N && wcol < p.w.K_pad) {
            const long flat = static_cast<long>(wrow) * K + wcol;
            const long blk = flat >> 6;
            valid[u] = true;
            pk[u] = __ldg(reinterpret_cast<const uint4*>(p.w.packed_rows + (static_cast<size_t>(wrow) * p.w.K_pad + wcol) / 2));
            qa[u] = 64u - static_cast<uint32_t>(flat & 63);
            nest[u] = __ldg(p.w.absmax_f32 + (blk < nblocks ? blk : nblocks - 1));
            amb[u] = __ldg(p.w.absmax_f32 + (blk + 1 < nblocks ? blk + 1 : nblocks - 1));
          }
        } else if (wrow < p.w.N && wcol < K) {
          const size_t flat = static_cast<size_t>(wrow) * K + wcol;
          valid[u] = true;
          pk[u] = __ldg(reinterpret_cast<const uint4*>(p.w.packed + (flat >> 1)));
          const size_t blk = flat >> 6;
          qa[u] = __ldg(p.w.qabsmax + blk);
          nest[u] = __ldg(p.w.nested_absmax + (blk >> 8));
        }
      }
    };

    const uint32_t code_s = smem_u32(s_code);
    uint32_t it = 0;
    int tile = blockIdx.x;
    int ks = 0;
    if (tile < num_tiles) prefetch(tile, 0);
    while (tile < num_tiles) {
      const int s = it % kStages;
      // current task data -> locals, then start the loads of the next (tile, kstep)
      uint4 cpk[kPerThread];
      float cam[kPerThread], camb[kPerThread];
      int csplit[kPerThread];
      bool cvalid[kPerThread];
#pragma unroll
      for (int u = 0; u < kPerThread; ++u) {
        cpk[u] = pk[u];
        cvalid[u] = valid[u];
        if (kRagged) {
          cam[u] = nest[u];
          camb[u] = amb[u];
          csplit[u] = static_cast<int>(qa[u]);
        } else {
          // double-quant decode, two separately rounded fp32 ops exactly like dequantize_blockwise followed by "+= offset"
          cam[u] = valid[u] ? __fadd_rn(__fmul_rn(lds_f32(code_s + 64 + qa[u] * 4), nest[u]), p.w.offset) : 0.f;
          camb[u] = 0.f;
          csplit[u] = 64;
        }
      }
      const int cur_tile = tile, cur_ks = ks;
      if (++ks == ksteps) {
        ks = 0;
        tile += gridDim.x;
      }
      if (tile < num_tiles) prefetch(tile, ks);

      mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);
      const uint32_t sb = smem_u32(smem + s * S::kStageBytes + S::kABytes);
      if (kNF4) {
#pragma unroll
        for (int u = 0; u < kPerThread; ++u) {
          const int t = dt + u * kDeqThreads;
          if (t >= kTasks) continue;
          const int half = t & 1;
          uint32_t row_s;
          int rin;
          if (!kBwd) {
            const int r = t >> 1;
            row_s = sb + r * 128;
            rin = r & 7;
          } else {
            const int r = (t & 127) >> 1;
            row_s = sb + (t >> 7) * 8192 + r * 128;
            rin = r & 7;
          }
          if (cvalid[u]) {
            if (kRagged) nf4_dequant32_split_to_swizzled(cpk[u], cam[u], camb[u], csplit[u], code_s, row_s, half, rin);
            else nf4_dequant32_to_swizzled(cpk[u], cam[u], code_s, row_s, half, rin);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) st_shared_zero16(row_s + (((half * 4 + i) ^ rin) * 16));
          }
        }
      }
      if (kBwd && kLoRA) {
        // partial MN block: row r (weight row n = ks*64 + r) holds lora_up[n, 0:16] in logical chunks 0 and 1
        if (dt < 128) {
          const int r = dt >> 1, c = dt & 1;
          const int n = cur_ks * kBK + r;
          uint4 v = make_uint4(0, 0, 0, 0);
          if (n < p.w.N) v = __ldg(reinterpret_cast<const uint4*>(p.lora_up + static_cast<size_t>(n) * kRank + c * 8));
          st_shared_v4(sb + (BN / 64) * 8192 + r * 128 + ((c ^ (r & 7)) * 16), v);
        }
      }
      (void)cur_tile;
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
      ++it;
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace vpt
