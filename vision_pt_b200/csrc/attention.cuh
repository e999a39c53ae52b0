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
  static constexpr int kP = kV + 2 * 16384;   // P_A, P_B: [128 queries x 64 keys] bf16 each
  static constexpr int kStat = kP + 32768;    // (m, l) of both streams per query row, [2 item parities][2 streams][128] float2
  static constexpr int kBars = kStat + 4096;
  static constexpr int kNumBars = 15;         // q_full q_empty kv_full[2] kv_empty[2] s_full[2] p_full[2] pv_done[2] o_free s_free[2]
  static constexpr int kTmemSlot = kBars + kNumBars * 8;
  static constexpr int kTotal = kTmemSlot + 16;   // no alignment slack: two CTAs must fit in one SM's 228 KB
};

// Persistent CTAs (one per SM) walk work items (128-query tile, head, sample).  Every 128-key block is split into two
// 64-key halves A and B, each an independent online-softmax stream with its own warpgroup, running maximum, row sum and O
// accumulator (S_A, S_B, O_A, O_B: 64 TMEM columns each); the tensor core ping-pongs between the streams, and the two
// partial results are merged at the end of the item (O = (O_A w_A + O_B w_B) / (l_A w_A + l_B w_B)).  A stream rescales
// its accumulator only when its maximum grows by more than 2^8 (P stays <= 256: exact in the fp32 sums, safe in bf16).
// Sequences here are short (330 tokens = 2.6 key blocks per item), so the loop over items matters: the next item's Q / K / V
// are already in flight while the current one finishes, instead of one CTA launch + pipeline fill per 2.6 blocks.
// Warps: 0 TMA producer, 1 MMA issuer + TMEM allocator, 2-5 stream A, 6-9 stream B.
__global__ void __launch_bounds__(320, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnFwdParams p) {
  using S = AttnFwdSmem;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // 128B-swizzled tiles need a 1024B-aligned base
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBars);
  uint64_t* q_full = bars;
  uint64_t* q_empty = bars + 1;
  uint64_t* kv_full = bars + 2;    // [2]
  uint64_t* kv_empty = bars + 4;   // [2]
  uint64_t* s_full = bars + 6;     // [2] A, B
  uint64_t* p_full = bars + 8;     // [2]
  uint64_t* pv_done = bars + 10;   // [2]
  uint64_t* o_free = bars + 12;    // the merge has read O_A / O_B: the next item may overwrite them
  uint64_t* s_free = bars + 13;    // [2] the stream has S_X in registers: the next S_X may be computed while it works
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::kTmemSlot);
  float2* s_stat = reinterpret_cast<float2*>(smem + S::kStat);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqt = (p.Lq + kAttnTile - 1) / kAttnTile;
  const int num_items = nqt * p.H * p.B;
  auto klen_of = [&](int item) {
    if (item >= num_items) return 0;
    if (p.seqlens_k == nullptr) return p.Lk;
    const int kl = __ldg(p.seqlens_k + item / (nqt * p.H));
    return kl < p.Lk ? kl : p.Lk;
  };
  // keys of block j that stream X covers, rounded up to the UMMA granularity (0 = none)
  auto sub_n = [](int klen, int j, int X) {
    const int valid = min(128, klen - j * 128) - X * 64;
    return valid <= 0 ? 0 : min(64, (valid + 15) & ~15);
  };

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 128);
      mbar_init(&pv_done[s], 1);
      mbar_init(&s_free[s], 128);
    }
    mbar_init(o_free, 8);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tO = tmem_base + 128;   // S_A 0, S_B 64, O_A 128, O_B 192
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      uint32_t jj = 0, itn = 0;
      int klen = klen_of(blockIdx.x);
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++itn) {
        const int qt = item % nqt, h = (item / nqt) % p.H, b = item / (nqt * p.H);
        const int nblk = (klen + kAttnTile - 1) / kAttnTile;
        klen = klen_of(item + gridDim.x);
        mbar_wait(q_empty, (itn & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, 16384);
        tma_load_4d(&tmQ, q_full, smem + S::kQ, 0, qt * kAttnTile, h, b);
        for (int j = 0; j < nblk; ++j, ++jj) {
          const uint32_t s = jj & 1;
          mbar_wait(&kv_empty[s], ((jj >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&kv_full[s], 32768);
          tma_load_4d(&tmK, &kv_full[s], smem + S::kK + s * 16384, 0, j * kAttnTile, h, b);
          tma_load_4d(&tmV, &kv_full[s], smem + S::kV + s * 16384, 0, j * kAttnTile, h, b);
        }
      }
    }
  } else if (warp == 1) {
    // converged warp, one elected lane issues (operands stay in uniform registers)
    constexpr uint32_t kIdO = umma_idesc_bf16(128, 64, 0, 1);    // O_X += P_X V_X  (V is MN-major)
    const uint32_t smem_base = smem_u32(smem);
    const uint64_t dK_ = umma_smem_desc(0, 16, 1024, kLayoutSW128);
    const uint64_t dMN = umma_smem_desc(0, 8192, 1024, kLayoutSW128);
    const uint64_t qd = dK_ + ((smem_base + S::kQ) >> 4);
    uint32_t jj = 0, itn = 0;
    uint32_t cnt[2] = {0, 0};                     // P tiles consumed per stream
    uint32_t s_issued[2] = {0, 0};                // S_X MMAs issued per stream (each is answered by one s_free phase)
    int klen = klen_of(blockIdx.x);
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++itn) {
      const int nblk = (klen + kAttnTile - 1) / kAttnTile;
      const int kl = klen;
      klen = klen_of(item + gridDim.x);
      // S_X = Q K_X^T for block j.  It overwrites the stream's single S buffer, which is free as soon as the stream has
      // LOADED the previous S_X into registers (s_free) -- long before its exponentials are done -- so the next block's
      // scores are already waiting when the stream comes back for them.
      auto issue_s = [&](int j, uint32_t s, uint32_t sp, int X) {
        const int n = sub_n(kl, j, X);
        if (n == 0) return;
        if (s_issued[X] > 0) mbar_wait(&s_free[X], (s_issued[X] - 1) & 1);
        ++s_issued[X];
        mbar_wait(&kv_full[s], sp);
        tc_fence_after_sync();
        const uint64_t kd = dK_ + ((smem_base + S::kK + s * 16384 + X * 8192) >> 4);
        const uint32_t idesc = umma_idesc_bf16(128, n, 0, 0);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tS + X * 64, qd + 2 * k, kd + 2 * k, idesc, k != 0);
          umma_commit(&s_full[X]);
        }
        __syncwarp();
      };
      mbar_wait(q_full, itn & 1);
      if (nblk > 0) {
        issue_s(0, jj & 1, (jj >> 1) & 1, 0);
        issue_s(0, jj & 1, (jj >> 1) & 1, 1);
      }
      if (nblk <= 1) {                             // every S MMA of this item is issued: Q may be replaced
        if (elect_one_sync()) umma_commit(q_empty);
        __syncwarp();
      }
      bool first[2] = {true, true};
      for (int j = 0; j < nblk; ++j, ++jj) {
        const uint32_t s = jj & 1;
        const uint64_t vd = dMN + ((smem_base + S::kV + s * 16384) >> 4);
        if (j + 1 < nblk) {
          issue_s(j + 1, (jj + 1) & 1, ((jj + 1) >> 1) & 1, 0);
          issue_s(j + 1, (jj + 1) & 1, ((jj + 1) >> 1) & 1, 1);
          if (j + 2 == nblk) {
            if (elect_one_sync()) umma_commit(q_empty);
            __syncwarp();
          }
        }
#pragma unroll
        for (int X = 0; X < 2; ++X) {
          const int n = sub_n(kl, j, X);
          if (n > 0) {
            mbar_wait(&p_full[X], cnt[X] & 1);
            ++cnt[X];
            if (first[X] && itn > 0 && X == 0) mbar_wait(o_free, (itn - 1) & 1);   // previous item's merge has read O
            tc_fence_after_sync();
            const uint64_t pd = dK_ + ((smem_base + S::kP + X * 16384) >> 4);
            const uint32_t acc0 = first[X] ? 0u : 1u;
            first[X] = false;
            if (elect_one_sync()) {
              for (int k = 0; k < n / 16; ++k) umma_ss(tO + X * 64, pd + 2 * k, vd + (X * 4 + k) * 128, kIdO, acc0 | (k != 0));
              umma_commit(&pv_done[X]);
            }
            __syncwarp();
          }
          if (X == 1) {
            if (elect_one_sync()) umma_commit(&kv_empty[s]);
            __syncwarp();
          }
        }
      }
    }
  } else {
    // ============================================================ softmax streams: A = warps 2-5, B = warps 6-9
    const int X = (warp - 2) >> 2;
    const int qd = warp & 3;                      // TMEM lane quarter
    const int row = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t p_row = smem_u32(smem + S::kP) + X * 16384 + row * 128;
    const int rin = row & 7;
    uint32_t cx = 0;
    int klen_next = klen_of(blockIdx.x), klen_next2 = klen_of(blockIdx.x + gridDim.x);
    uint32_t itn_s = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++itn_s) {
      const int qt = item % nqt, h = (item / nqt) % p.H, b = item / (nqt * p.H);
      const int klen = klen_next;
      klen_next = klen_next2;
      klen_next2 = klen_of(item + 2 * gridDim.x);   // fetched two items ahead of its first use
      const int nblk = (klen + kAttnTile - 1) / kAttnTile;
      float m_ref = -INFINITY, l_run = 0.f;
      bool first = true;
      for (int j = 0; j < nblk; ++j) {
        const int n = sub_n(klen, j, X);
        if (n == 0) continue;
        mbar_wait(&s_full[X], cx & 1);
        tc_fence_after_sync();
        const int nvalid = min(64, klen - j * 128 - X * 64);     // < n only in the last block
        uint32_t v[64];
        tmem_ld32(tS + X * 64 + lane_off, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        if (n > 32) tmem_ld32(tS + X * 64 + lane_off + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        tmem_wait_ld();
        tc_fence_before_sync();
        mbar_arrive(&s_free[X]);                    // the scores are in registers: the tensor core may write the next S_X
        if (nvalid < 64) {
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (i >= nvalid) v[i] = 0xff800000u;   // -inf: padded / masked keys
        }
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 64; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        mx *= p.scale_log2;                         // scale > 0: the maximum commutes with the scaling
        // lazy rescale: keep the reference maximum unless the true one outgrew it by 2^8
        const bool grow = mx > m_ref + 8.f;
        const float m_new = grow ? mx : m_ref;
        if (cx > 0) {
          // P_X of the previous block (possibly the previous item's) is consumed before it is overwritten, and O_X is
          // final for it before it is rescaled
          mbar_wait(&pv_done[X], (cx - 1) & 1);
          tc_fence_after_sync();
        }
        if (!first && __any_sync(0xffffffffu, grow)) {
          const float alpha = grow ? fast_exp2(m_ref - m_new) : 1.f;
          l_run *= alpha;
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld32(tO + X * 64 + lane_off + c * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(tO + X * 64 + lane_off + c * 32, o);
          }
          tmem_wait_st();
        }
        first = false;
        m_ref = m_new;
        const float neg_m = -m_ref;
        float psum = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float p0 = fast_exp2(fmaf(__uint_as_float(v[g * 8 + 2 * e]), p.scale_log2, neg_m));
            const float p1 = fast_exp2(fmaf(__uint_as_float(v[g * 8 + 2 * e + 1]), p.scale_log2, neg_m));
            psum += p0 + p1;
            pk[e] = pack_bf16x2(p0, p1);
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_row + ((g ^ rin) * 16)), "r"(pk[0]), "r"(pk[1]),
                       "r"(pk[2]), "r"(pk[3])
                       : "memory");
        }
        l_run += psum;
        ++cx;
        fence_proxy_async_smem();
        tc_fence_before_sync();
        mbar_arrive(&p_full[X]);
      }
      // ---- merge the two streams
      if (!first) {
        mbar_wait(&pv_done[X], (cx - 1) & 1);
        tc_fence_after_sync();
      }
      // one barrier per item: the statistics are double-buffered by item parity (a thread is never more than one
      // item ahead of another, since every item has this exchange)
      float2* stat = s_stat + (itn_s & 1) * 256;
      stat[X * 128 + row] = make_float2(m_ref, l_run);
      named_bar_sync(1, 256);
      const float2 sa = stat[row], sb = stat[128 + row];
      const float m = fmaxf(sa.x, sb.x);
      const float wa = sa.y > 0.f ? fast_exp2(sa.x - m) : 0.f;     // a stream without keys has l = 0
      const float wb = sb.y > 0.f ? fast_exp2(sb.x - m) : 0.f;
      const float l = sa.y * wa + sb.y * wb;
      const float inv_l = l > 0.f ? 1.f / l : 0.f;
      const int q = qt * kAttnTile + row;
      if (X == 0)   // row pitch = Lq rounded up to 128 (one bulk copy per tile in the backward); +inf for padded queries
        p.lse2[(static_cast<long>(b) * p.H + h) * (static_cast<long>(nqt) * kAttnTile) + q] =
            q < p.Lq ? (l > 0.f ? m + log2f(l) : -INFINITY) : INFINITY;
      // stream X writes output columns X*32 .. X*32+31, reading that slice of both accumulators
      const bool has_a = klen > 0, has_b = klen > 64;               // CTA-uniform: the .sync.aligned loads stay convergent
      uint32_t oa[32], ob[32];
      if (has_a) tmem_ld32(tO + lane_off + X * 32, oa);
      if (has_b) tmem_ld32(tO + 64 + lane_off + X * 32, ob);
      tmem_wait_ld();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);
      if (q < p.Lq) {
        __nv_bfloat16* orow = p.o + b * p.o_sb + static_cast<long>(q) * p.o_sl + h * p.o_sh + X * 32;
        const float fa = wa * inv_l, fb = wb * inv_l;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float x0 = 0.f, x1 = 0.f;
            if (has_a) {
              x0 = __uint_as_float(oa[g * 8 + 2 * e]) * fa;
              x1 = __uint_as_float(oa[g * 8 + 2 * e + 1]) * fa;
            }
            if (has_b) {
              x0 = fmaf(__uint_as_float(ob[g * 8 + 2 * e]), fb, x0);
              x1 = fmaf(__uint_as_float(ob[g * 8 + 2 * e + 1]), fb, x1);
            }
            w[e] = pack_bf16x2(x0, x1);
          }
          *reinterpret_cast<uint4*>(orow + g * 8) = make_uint4(w[0], w[1], w[2], w[3]);
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
