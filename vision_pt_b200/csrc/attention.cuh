// Attention forward / backward for sm_100a (tcgen05 + TMEM + TMA), head_dim 64, non-causal, per-sample valid key
// length (the reference's key-padding mask on trailing context tokens: src/models/jit/denoiser.py:375-381,
// src/modules/attention.py:98-129 with mask [B,1,1,Lk] -> here `seqlens_k[b]`).
//
// q/k/v/o are addressed as 4-D tensors (hd, L, H, B) with arbitrary strides for L, H, B (hd contiguous), so both the
// reference's [B,H,L,hd] layout and the token-major [B,L,H,hd] layout the fused block uses are accepted without copies.
//
// Forward: attn_fwd_kernel below.
// Backward: attention_bwd.cuh.
#pragma once
#include "sm100.cuh"

namespace vpt {

constexpr int kAttnHD = 64;
constexpr int kAttnTile = 128;
#ifndef VPT_ATTN_POLY_EVERY
#define VPT_ATTN_POLY_EVERY 2
#endif
constexpr int kAttnPolyEvery = VPT_ATTN_POLY_EVERY;   // one exponential in 2 * kAttnPolyEvery goes to the FMA pipe

struct AttnFwdParams {
  int B, H, Lq, Lk;
  const int* seqlens_k;        // [B] or nullptr (= Lk)
  float scale_log2;            // softmax scale * log2(e)
  __nv_bfloat16* o;            // element strides below
  long o_sb, o_sl, o_sh;
  float* lse2;                 // [B, H, Lq rounded up to 128]  log2-domain logsumexp of scale*S; +inf in the padding
};

// 2^x on the FMA pipe (Cody-Waite split + degree-3 polynomial on [-0.5, 0.5], max relative error 7.7e-5: far inside the
// bf16 rounding of P).  The forward's exponentials are bound by the MUFU (16 ex2 / clock / SM); moving one in four to the
// FMA / ALU pipes balances the two (x must be <= ~100: here x <= 8 by the lazy-rescale rule).
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;                // 1.5 * 2^23: rint(x) lands in the low mantissa bits
  const float f = x - (t - 12582912.f);          // in [-0.5, 0.5]
  float r = fmaf(f, 0.05508868396282196f, 0.24260404706001282f);
  r = fmaf(r, f, 0.6932762265205383f);
  r = fmaf(r, f, 0.9999289512634277f);
  return __int_as_float(__float_as_int(r) + (__float_as_int(t) << 23));   // * 2^rint(x)
}

template <int HD>
struct AttnFwdSmemT {
  static constexpr int kBlk = (HD + 63) / 64;          // 64-column (128-byte, swizzled) blocks per row: 80 -> 2, the second
  static constexpr int kTile = kBlk * 16384;           // zero-filled past the head dimension by the TMA unit
  static constexpr int kStreams = 2;
  // per stream (X = 0, 1): Q[kQBufs] (items), K[kKVStages], V[kKVStages] (key blocks).  head_dim 64: all double-buffered
  // (6 x 16 KB); head_dim 80: single-buffered (3 x 32 KB) -- a stream then waits ~1.5 us for every tile it loads, and the
  // other stream fills that time
  static constexpr int kQBufs = HD == 64 ? 2 : 1;
  static constexpr int kKVStages = HD == 64 ? 2 : 1;
  static constexpr bool kPInPlace = HD != 64;          // TMEM: 2 x (128 + 80 + 64) > 512, so P goes over the S columns it came from
  static constexpr int kStream = (kQBufs + 2 * kKVStages) * kTile;
  static constexpr int kQ = 0, kK = kQBufs * kTile, kV = (kQBufs + kKVStages) * kTile;
  static constexpr int kOut = kStreams * kStream;   // 8 softmax warps x one slab of [32 rows x 128 B] (output tile -> TMA store)
  static constexpr int kXchg = kOut + 8 * 4096;   // per stream: 2 column halves x 128 rows, fp32 row maxima (row sums at item end)
  static constexpr int kBars = kXchg + 2 * 1024;
  // per stream: q_full[2] q_empty[2] kv_full[2] kv_empty[2] s_full p_full (256 arrivals) pv_done  (11)
  static constexpr int kBarsPerStream = 11;
  static constexpr int kTmemSlot = kBars + 2 * kBarsPerStream * 8;
  static constexpr int kTotal = kTmemSlot + 16;   // no alignment slack: the dynamic segment starts 1024B-aligned
  static_assert(kTotal <= 232448, "shared memory");
};
using AttnFwdSmem = AttnFwdSmemT<64>;

// Persistent CTAs (one per SM) run TWO independent streams, each walking its own work items (128-query tile, head,
// sample) with its own softmax warpgroup, MMA-issuing warp, TMA producer warp, Q / K / V buffers and TMEM columns:
//
//   tensor core    S = Q K_j^T  (128 queries x <=128 keys, fp32, TMEM)
//   2 warpgroups   each takes one half of the block's key columns: pass 1 row maximum (exchanged between the halves
//                  through shared memory), pass 2 P = exp2(S c - m) -> bf16 into the stream's P columns of TMEM (A operand
//                  of the next MMA, never through shared memory), partial row sums in registers
//   tensor core    O += P V_j (A from TMEM), then S of the next key block (or of the stream's next item)
//   warpgroup      end of item: O / l -> bf16 -> 128B-swizzled slab -> TMA store (tmO: box {64, 32}), lse2
//
// The streams never exchange anything, so one stream's exponentials (the MUFU is the binding unit at head_dim 64:
// 128 x 128 ex2 per key block = 1024 cycles of the SM's 16 ex2 / clock, against 512 tensor cycles) hide the other
// stream's MMAs, epilogue and item turn-over; two issuing warps because tcgen05.mma issue blocks while the pipe is busy
// (tools/probes/umma_rate.cu).  The previous form (one item per CTA, key halves as streams, merge at the end) spent 43 % of
// its time in per-item merge / turn-over at 2.6 key blocks per item (profiles/r1l_attn_fwd_phases.txt).
// A stream rescales its accumulator only when its maximum grows by more than 2^8 (P stays <= 256: exact in the fp32
// sums, safe in bf16).
// Eight softmax warps per stream (two per TMEM lane quarter): with four, a scheduler held two softmax warps and the
// tcgen05.ld -> ex2 -> tcgen05.st chains were latency-bound (MUFU 30 % busy, profiles/r1o_attn_fwd_ncu.txt).
// Warps: 0 / 2 TMA producers of stream A / B (2 also allocates TMEM), 1 / 3 MMA issuers, 4-11 softmax A, 12-19 softmax B.
// head_dim 80 (JiT-H): the same kernel on two-block tiles -- QK^T runs a fifth k-step over the second block, O += P V uses
// N = 80 across both blocks (MN-major operand, leading-block stride = one 16 KB block) -- with single-buffered Q / K / V
// (three 32 KB tiles per stream), each half's P written over its own S columns (64 half .. 64 half + 31), and the output
// written straight from registers.
template <int HD>
__global__ void __launch_bounds__(640, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const AttnFwdParams p) {
  using S = AttnFwdSmemT<HD>;
  constexpr int kTile = S::kTile;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // 128B-swizzled tiles need a 1024B-aligned base
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::kTmemSlot);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int X = warp < 4 ? (warp >> 1) : ((warp - 4) >> 3);          // stream of this warp
  uint8_t* sm = smem + (X < S::kStreams ? X : 0) * S::kStream;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBars) + X * S::kBarsPerStream;
  uint64_t* q_full = bars;          // [2]
  uint64_t* q_empty = bars + 2;     // [2]
  uint64_t* kv_full = bars + 4;     // [2]
  uint64_t* kv_empty = bars + 6;    // [2]
  uint64_t* s_full = bars + 8;      // S_j is complete (and with it every earlier MMA of the stream)
  uint64_t* p_full = bars + 9;      // P_j is in TMEM
  uint64_t* pv_done = bars + 10;    // the item's last O += P V is complete

  const int nqt = (p.Lq + kAttnTile - 1) / kAttnTile;
  const int num_items = nqt * p.H * p.B;
  const int first_item = blockIdx.x * S::kStreams + X, item_stride = gridDim.x * S::kStreams;
  const bool idle = X >= S::kStreams;
  auto klen_of = [&](int item) {
    if (item >= num_items) return 0;
    if (p.seqlens_k == nullptr) return p.Lk;
    const int kl = __ldg(p.seqlens_k + item / (nqt * p.H));
    return kl < p.Lk ? kl : p.Lk;
  };

  if (threadIdx.x == 0) {
    uint64_t* all = reinterpret_cast<uint64_t*>(smem + S::kBars);
    for (int x = 0; x < 2; ++x) {
      uint64_t* bx = all + x * S::kBarsPerStream;
      for (int i = 0; i < 9; ++i) mbar_init(&bx[i], 1);
      mbar_init(&bx[9], 256);
      mbar_init(&bx[10], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // per stream: S 0..127, O 128..128+HD-1, P (64 columns of bf16 pairs) behind it; stream 1 starts at column 256
  const uint32_t tS = tmem_base + X * 256, tO = tS + 128, tP = tO + 64;    // tP: head_dim 64 only
  pdl_launch_dependents();
  pdl_wait();

  if (idle) {
    // nothing: joins the final barrier below
  } else if (warp == 0 || warp == 2) {
    // ============================================================ TMA producer of stream X
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      uint32_t it = 0, jj = 0;
      for (int item = first_item; item < num_items; item += item_stride, ++it) {
        const int qt = item % nqt, h = (item / nqt) % p.H, b = item / (nqt * p.H);
        const int nblk = (klen_of(item) + kAttnTile - 1) / kAttnTile;
        const uint32_t qb = it % S::kQBufs;
        mbar_wait(&q_empty[qb], ((it / S::kQBufs) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[qb], kTile);
#pragma unroll
        for (int blk = 0; blk < S::kBlk; ++blk)
          tma_load_4d(&tmQ, &q_full[qb], sm + S::kQ + qb * kTile + blk * 16384, blk * 64, qt * kAttnTile, h, b);
        for (int j = 0; j < nblk; ++j, ++jj) {
          const uint32_t s = jj % S::kKVStages;
          mbar_wait(&kv_empty[s], ((jj / S::kKVStages) & 1) ^ 1);
          mbar_arrive_expect_tx(&kv_full[s], 2 * kTile);
#pragma unroll
          for (int blk = 0; blk < S::kBlk; ++blk) {
            tma_load_4d(&tmK, &kv_full[s], sm + S::kK + s * kTile + blk * 16384, blk * 64, j * kAttnTile, h, b);
            tma_load_4d(&tmV, &kv_full[s], sm + S::kV + s * kTile + blk * 16384, blk * 64, j * kAttnTile, h, b);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == 3) {
    // ============================================================ MMA issuer of stream X (converged warp, one elected lane)
    constexpr uint32_t kIdO = umma_idesc_bf16(128, HD, 0, 1);    // O += P V  (V is MN-major)
    const uint32_t sbase = smem_u32(sm);
    const uint64_t dK_ = umma_smem_desc(0, 16, 1024, kLayoutSW128);
    const uint64_t dMN = umma_smem_desc(0, 16384, 1024, kLayoutSW128);   // MN-major: 64-column blocks 16 KB apart
    uint32_t it = 0, jj = 0, pc = 0;
    // TMEM columns of the 16 keys of k-step k of P: its own 64 columns, or (in place) the first 32 columns of the S half
    // the keys belong to
    auto p_cols = [&](int k) { return S::kPInPlace ? tS + (k >> 2) * 64 + (k & 3) * 8 : tP + k * 8; };
    PROF_DECL(8)
    for (int item = first_item; item < num_items; item += item_stride, ++it) {
      const int kl = klen_of(item);
      const int nblk = (kl + kAttnTile - 1) / kAttnTile;
      const uint32_t qb = it % S::kQBufs;
      const uint64_t qd = dK_ + ((sbase + S::kQ + qb * kTile) >> 4);
      // S = Q K_j^T (the softmax halves have read the previous S: p_full)
      auto issue_s = [&](int j, uint32_t jn) {
        const uint32_t s = jn % S::kKVStages;
        const int valid = min(128, kl - j * 128);
        const int n = (valid + 15) & ~15;
        PROF(0)
        mbar_wait(&kv_full[s], (jn / S::kKVStages) & 1);
        PROF(1)
        tc_fence_after_sync();
        const uint64_t kd = dK_ + ((sbase + S::kK + s * kTile) >> 4);
        const uint32_t idesc = umma_idesc_bf16(128, n, 0, 0);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tS, qd + 2 * k, kd + 2 * k, idesc, k != 0);
#pragma unroll
          for (int k = 0; k < (HD - 64) / 16; ++k)        // head-dim columns 64.. : the tiles' second block (+16 KB)
            umma_ss(tS, qd + 1024 + 2 * k, kd + 1024 + 2 * k, idesc, 1);
          umma_commit(s_full);
        }
        __syncwarp();
        PROF(2)
      };
      PROF(0)
      mbar_wait(&q_full[qb], (it / S::kQBufs) & 1);
      PROF(3)
      if (nblk > 0) issue_s(0, jj);
      for (int j = 0; j < nblk; ++j, ++jj) {
        const uint32_t s = jj % S::kKVStages;
        const uint64_t vd = dMN + ((sbase + S::kV + s * kTile) >> 4);
        const int valid = min(128, kl - j * 128);
        const int n = (valid + 15) & ~15;
        PROF(0)
        mbar_wait(p_full, pc & 1);
        PROF(4)
        ++pc;
        tc_fence_after_sync();
        if (elect_one_sync()) {
          if (n == 128) {
#pragma unroll
            for (int k = 0; k < 8; ++k) umma_ts(tO, p_cols(k), vd + k * 128, kIdO, (j | k) != 0);
          } else {
            for (int k = 0; k < n / 16; ++k) umma_ts(tO, p_cols(k), vd + k * 128, kIdO, (j | k) != 0);
          }
          umma_commit(&kv_empty[s]);
        }
        __syncwarp();
        PROF(5)
        if (j + 1 < nblk) issue_s(j + 1, jj + 1);
      }
      if (elect_one_sync()) {
        umma_commit(pv_done);
        umma_commit(&q_empty[qb]);     // every S MMA of the item is complete
      }
      __syncwarp();
    }
#ifdef VPT_BWD_PROF
    PROF(0)
    if (blockIdx.x == 0 && lane == 0)
      printf("fwd mma%d: other %lld kv_full %lld S issue %lld q_full %lld p_full %lld PV issue %lld\n", X, prof_[0], prof_[1], prof_[2],
             prof_[3], prof_[4], prof_[5]);
#endif
  } else {
    // ============================================================ softmax warps of stream X (two column halves)
    const int half = ((warp - 4) >> 2) & 1;       // key columns 64 half .. 64 half + 63 of every block
    const int qd = warp & 3;                      // TMEM lane quarter
    const int row = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    uint8_t* out_slab = smem + S::kOut + (X * 4 + qd) * 4096;       // used by half 0 (the epilogue half)
    const uint32_t out_row = smem_u32(out_slab) + lane * 128;
    float* xchg = reinterpret_cast<float*>(smem + S::kXchg + X * 1024);   // [half][128]
    if (lane == 0 && half == 0) tma_prefetch_desc(&tmO);
    uint32_t cx = 0, it = 0;
    PROF_DECL(8)
    for (int item = first_item; item < num_items; item += item_stride, ++it) {
      const int qt = item % nqt, h = (item / nqt) % p.H, b = item / (nqt * p.H);
      const int klen = klen_of(item);
      const int nblk = (klen + kAttnTile - 1) / kAttnTile;
      float m_ref = -INFINITY, l_run = 0.f;       // l_run: this half's part of the row sum
      for (int j = 0; j < nblk; ++j) {
        const int nvalid = min(128, klen - j * 128);          // < 128 only in the last block
        const int nch = (nvalid + 31) >> 5;                   // 32-column chunks that hold keys
        const int c_begin = 2 * half, c_end = min(nch, 2 * half + 2);
        PROF(0)
        mbar_wait(s_full, cx & 1);
        PROF(1)
        ++cx;
        tc_fence_after_sync();
        // ---- pass 1: row maximum over this half's columns
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = c_begin; c < c_end; ++c) {
          uint32_t v[32];
          tmem_ld32(tS + lane_off + c * 32, v);
          tmem_wait_ld();
          const int lim = nvalid - c * 32;
          if (lim < 32) {                            // only the item's last chunk
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i >= lim) v[i] = 0xff800000u;   // -inf: padded / masked keys
          }
          float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            m0 = fmaxf(m0, __uint_as_float(v[i]));
            m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
          }
          mx = fmaxf(mx, fmaxf(m0, m1));
        }
        // the halves agree on ONE reference maximum through shared memory.  One buffer is enough: this half's next write to
        // its slot comes after s_full of the next block, which the issuer commits only after p_full of this block, which
        // the other half arrives on after it has read the slot.
        mx *= p.scale_log2;                         // scale > 0: the maximum commutes with the scaling
        xchg[half * 128 + row] = mx;
        named_bar_sync(1 + X, 256);
        mx = fmaxf(mx, xchg[(half ^ 1) * 128 + row]);
        PROF(2)
        // lazy rescale: keep the reference maximum unless the true one outgrew it by 2^8
        const bool grow = mx > m_ref + 8.f;
        const float m_new = grow ? mx : m_ref;
        if (j > 0 && __any_sync(0xffffffffu, grow)) {
          const float alpha = grow ? fast_exp2(m_ref - m_new) : 1.f;
          l_run *= alpha;
          if (half == 0) {
            // O is final for block j-1: s_full of block j was committed after that MMA.  One half rescales it; the other's
            // P store and this one's are both behind p_full (256 arrivals) before the next accumulation.
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
              uint32_t o[32];
              tmem_ld32(tO + lane_off + c * 32, o);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st32(tO + lane_off + c * 32, o);
            }
            if (HD > 64) {
              uint32_t o[16];
              tmem_ld16(tO + lane_off + 64, o);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st16(tO + lane_off + 64, o);
            }
          }
        }
        PROF(3)
        m_ref = m_new;
        const float neg_m = -m_ref;
        // ---- pass 2: P = exp2(S c - m) -> bf16 into the P columns
        float ps0 = 0.f, ps1 = 0.f;
#pragma unroll 1
        for (int c = c_begin; c < c_end; ++c) {
          uint32_t v[32];
          tmem_ld32(tS + lane_off + c * 32, v);
          tmem_wait_ld();
          const int lim = nvalid - c * 32;
          if (lim < 32) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i >= lim) v[i] = 0xff800000u;   // exp2(-inf) = 0 on both paths
          }
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float x0 = fmaf(__uint_as_float(v[2 * e]), p.scale_log2, neg_m);
            const float x1 = fmaf(__uint_as_float(v[2 * e + 1]), p.scale_log2, neg_m);
            const float p0 = fast_exp2(x0);
            const float p1 = (e % kAttnPolyEvery == kAttnPolyEvery - 1) ? poly_exp2(x1) : fast_exp2(x1);
            ps0 += p0;
            ps1 += p1;
            pk[e] = pack_bf16x2(p0, p1);
          }
          // in place: this half's chunk c' = c - 2 half lands in columns 64 half + 16 c' .. + 15, inside S columns the
          // half has already read (chunk 0: its own first 32; chunk 1: still chunk 0's range)
          tmem_st16((S::kPInPlace ? tS + 64 * half + (c - 2 * half) * 16 : tP + c * 16) + lane_off, pk);
        }
        l_run += ps0 + ps1;
        tmem_wait_st();
        tc_fence_before_sync();
        mbar_arrive(p_full);                        // also: this half has read its S columns (the next S may be issued)
        PROF(4)
      }
      // ---- end of item: the halves add their row sums; half 0 writes O / l -> bf16 -> slab -> TMA store
      PROF(0)
      mbar_wait(pv_done, it & 1);
      PROF(5)
      tc_fence_after_sync();
      if (nblk > 0) {
        // half 1 hands its row sum over in half 0's slot: half 1 read that slot before its own last pass 2, half 0 writes
        // it again only after this read (program order), and half 1 reads it again only after the next block's barrier
        if (half == 1) xchg[row] = l_run;
        named_bar_sync(1 + X, 256);
        if (half == 0) l_run += xchg[row];
      }
      if (half == 0) {
        const float inv_l = l_run > 0.f ? 1.f / l_run : 0.f;
        const int q = qt * kAttnTile + row;
        // row pitch = Lq rounded up to 128 (one bulk copy per tile in the backward); +inf for padded queries
        p.lse2[(static_cast<long>(b) * p.H + h) * (static_cast<long>(nqt) * kAttnTile) + q] =
            q < p.Lq ? (l_run > 0.f ? m_ref + log2f(l_run) : -INFINITY) : INFINITY;
        if constexpr (HD == 64) {
          if (lane == 0) tma_store_wait_read<0>();     // the previous item's store has read the slab
          __syncwarp();
  #pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            if (nblk > 0) {                            // stream-uniform: the .sync.aligned load stays convergent
              tmem_ld32(tO + lane_off + c * 32, o);
              tmem_wait_ld();
            } else {
  #pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = 0u;
            }
  #pragma unroll
            for (int g = 0; g < 4; ++g)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(out_row + (((c * 4 + g) ^ (lane & 7)) * 16)),
                           "r"(pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l)),
                           "r"(pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l)),
                           "r"(pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l)),
                           "r"(pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l))
                           : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (qt * kAttnTile + qd * 32 < p.Lq)       // rows past Lq inside the box are clipped by the TMA unit
              tma_store_4d(&tmO, out_slab, 0, qt * kAttnTile + qd * 32, h, b);
            tma_store_commit();
          }
        } else {
          // head_dim 80: 160-byte rows straight from registers (no room for a slab beside six 32 KB tiles)
          __nv_bfloat16* orow = p.o + b * p.o_sb + static_cast<long>(q) * p.o_sl + h * p.o_sh;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            uint32_t o[32];
            if (nblk > 0) {
              if (c < 2) {
                tmem_ld32(tO + lane_off + c * 32, o);
              } else {
                tmem_ld16(tO + lane_off + 64, *reinterpret_cast<uint32_t(*)[16]>(&o[0]));
              }
              tmem_wait_ld();
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = 0u;
            }
            if (q < p.Lq) {
#pragma unroll
              for (int g = 0; g < (c < 2 ? 4 : 2); ++g) {
                uint32_t w[4];
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  w[e] = pack_bf16x2(__uint_as_float(o[g * 8 + 2 * e]) * inv_l, __uint_as_float(o[g * 8 + 2 * e + 1]) * inv_l);
                *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = make_uint4(w[0], w[1], w[2], w[3]);
              }
            }
          }
        }
        // the accumulator is read: order those loads before this half's next p_full arrival, after which the issuer
        // overwrites O (accumulate = 0 on the next item's first block)
        tc_fence_before_sync();
      }
      PROF(6)
    }
#ifdef VPT_BWD_PROF
    PROF(0)
    if (blockIdx.x == 0 && lane == 0 && qd == 0)
      printf("fwd wg%d.%d: other %lld s_full %lld pass1 %lld rescale %lld pass2 %lld pv_done %lld epilogue %lld\n", X, half, prof_[0], prof_[1],
             prof_[2], prof_[3], prof_[4], prof_[5], prof_[6]);
#endif
    if (lane == 0 && half == 0) tma_store_wait_all<0>();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
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
// head_dim 64.  CTA = 8 consecutive query rows of one batch entry (blockIdx.y); 8 lanes x 8 elements = one head of one row.
// The (row, head) items run heads-fastest, i.e. in the memory order of a [B, L, H*64] tensor, two items per lane group and
// pass so that four 16-byte loads per thread are in flight; no 64-bit division anywhere.  CTA size does not matter
// (8 rows: 16.8 us, 32 rows: 16.0 us at the JiT-B shape; the (b, h, q)-ordered form with 64-bit index math took 18.8 us).
constexpr int kDeltaRows = 8;
static_assert(kDeltaRows % 4 == 0, "the pass count must be uniform over a warp (four lane groups)");
__global__ void __launch_bounds__(256) attn_bwd_delta_kernel(const AttnDeltaParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y, q0 = blockIdx.x * kDeltaRows;
  const int grp = threadIdx.x >> 3, sub = threadIdx.x & 7;
  const int items = kDeltaRows * p.H;
  const __nv_bfloat16* ob = p.o + b * p.o_sb + sub * 8;
  const __nv_bfloat16* gb = p.d_o + b * p.do_sb + sub * 8;
  float* drow = p.delta + static_cast<long>(b) * p.H * p.Lq_pad;
  for (int i0 = grp; i0 < items; i0 += 64) {     // items is a multiple of 4 lane groups: the trip count is uniform over a warp
    uint4 a[2], g[2];
    int q[2], h[2];
    bool on[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = i0 + 32 * u;
      const int ql = i / p.H;
      h[u] = i - ql * p.H;
      q[u] = q0 + ql;
      on[u] = i < items;
      a[u] = g[u] = make_uint4(0, 0, 0, 0);
      if (on[u] && q[u] < p.Lq) {
        a[u] = __ldg(reinterpret_cast<const uint4*>(ob + static_cast<long>(q[u]) * p.o_sl + h[u] * p.o_sh));
        g[u] = __ldg(reinterpret_cast<const uint4*>(gb + static_cast<long>(q[u]) * p.do_sl + h[u] * p.do_sh));
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint32_t aw[4] = {a[u].x, a[u].y, a[u].z, a[u].w}, gw[4] = {g[u].x, g[u].y, g[u].z, g[u].w};
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) acc += bf16lo(aw[i]) * bf16lo(gw[i]) + bf16hi(aw[i]) * bf16hi(gw[i]);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if (on[u] && sub == 0) drow[static_cast<long>(h[u]) * p.Lq_pad + q[u]] = acc * p.scale;   // rows >= Lq: zero
    }
  }
}

// the same for any head_dim that is a multiple of 8: one thread per (b, h, q) row
template <int HD>
__global__ void __launch_bounds__(256) attn_bwd_delta_row_kernel(const AttnDeltaParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const long gid = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long total = static_cast<long>(p.B) * p.H * p.Lq_pad;
  if (gid >= total) return;
  const int q = gid % p.Lq_pad;
  const int h = (gid / p.Lq_pad) % p.H;
  const int b = gid / (static_cast<long>(p.Lq_pad) * p.H);
  float acc = 0.f;
  if (q < p.Lq) {
    const __nv_bfloat16* o = p.o + b * p.o_sb + static_cast<long>(q) * p.o_sl + h * p.o_sh;
    const __nv_bfloat16* g = p.d_o + b * p.do_sb + static_cast<long>(q) * p.do_sl + h * p.do_sh;
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(o + c * 8)), w = __ldg(reinterpret_cast<const uint4*>(g + c * 8));
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) acc += bf16lo(aw[i]) * bf16lo(gw[i]) + bf16hi(aw[i]) * bf16hi(gw[i]);
    }
  }
  p.delta[gid] = acc * p.scale;
}

}  // namespace vpt
