// LoRA parameter gradients: skinny tcgen05 reductions over the token dimension, batched over several linears per launch.
//     Out_i[P, 16] += Src[M, P]^T * Small_i[M, 16]       i < nsmall <= 3   (fp32 accumulation, split over M across CTAs)
//   dB   = dY^T * Ts      Src = dY [M, N], Small = Ts  -> lora_up.weight.grad   [N, 16]
//   dA^T = X^T  * dTs     Src = X  [M, K], Small = dTs -> lora_down.weight.grad [16, K] (written transposed)
// This is the autograd of LoRALinear.forward (/root/reference/src/modules/peft/lora.py:100-104) for lora_down / lora_up.
// Linears that share their input (to_q / to_k / to_v read h1, w_1 / w_2 read h2) are one item with nsmall = 3 / 2, so the
// shared activation is read once.
//
// The kernel is HBM-bound (reads Src once).  Both operands arrive by TMA: Src tiles as MN-major A (its rows are the
// reduction dimension), the side tensors in the [16, M] layout the GEMM epilogues write them in, as K-major B rows.
// One launch covers every item of a transformer block's backward (<= 16 items), so the ~2 us a CTA spends on barrier
// / TMEM / descriptor setup overlaps other CTAs' streaming instead of serialising 14 small launches.
#pragma once
#include "host.cuh"
#include "sm100.cuh"

namespace vpt {

constexpr int kLgMaxItems = 16;
constexpr int kLgMaxMaps = 40;
constexpr int kLgMaxSmall = 3;

struct LoraGradItem {
  int M, P, nsmall, transposed;   // transposed 0: out[p * 16 + r]   1: out[r * ldo + p]
  int ldo;
  int map_src, map_small[kLgMaxSmall];
  float* out[kLgMaxSmall];
  int cta_begin, ptiles, rows_per_cta;
};

struct alignas(64) LoraGradBatch {
  CUtensorMap maps[kLgMaxMaps];
  LoraGradItem items[kLgMaxItems];
  int n_items;
};

struct LoraGradSmem {
  static constexpr int kStages = 4;
  static constexpr int kA = 16384, kB = kLgMaxSmall * 2048;
  static constexpr int kStage = kA + kB;
  static constexpr int kBars = kStages * kStage;
  static constexpr int kTmemSlot = kBars + (2 * kStages + 1) * 8;
  static constexpr int kTotal = kTmemSlot + 16 + 1024;
};

// maps[map_src]      : Src [M, P] row-major, box {64 (p), 64 (m)}, SWIZZLE_128B
// maps[map_small[i]] : Small_i^T [16, M] row-major, box {64 (m), 16 (r)}, SWIZZLE_128B
__global__ void __launch_bounds__(256, 2)
lora_grad_kernel(const __grid_constant__ LoraGradBatch bp) {
  using S = LoraGradSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::kBars);
  uint64_t* empty = full + S::kStages;
  uint64_t* acc_full = empty + S::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::kTmemSlot);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int ii = 0;
  while (ii + 1 < bp.n_items && static_cast<int>(blockIdx.x) >= bp.items[ii + 1].cta_begin) ++ii;
  const LoraGradItem& it_ = bp.items[ii];
  const int local = blockIdx.x - it_.cta_begin;
  const int p0 = (local % it_.ptiles) * 128;
  const int m_begin = (local / it_.ptiles) * it_.rows_per_cta;
  const int m_end = min(it_.M, m_begin + it_.rows_per_cta);
  const int nsteps = (m_end - m_begin + 63) / 64;
  const int ns = it_.nsmall;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 64);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      const CUtensorMap* tmS = &bp.maps[it_.map_src];
      tma_prefetch_desc(tmS);
      for (int it = 0; it < nsteps; ++it) {
        const int s = it % S::kStages;
        mbar_wait(&empty[s], ((it / S::kStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[s], S::kA + ns * 2048);
        uint8_t* sa = smem + s * S::kStage;
        const int m = m_begin + it * 64;
        tma_load_2d(tmS, &full[s], sa, p0, m);
        tma_load_2d(tmS, &full[s], sa + 8192, p0 + 64, m);
        for (int i = 0; i < ns; ++i) tma_load_2d(&bp.maps[it_.map_small[i]], &full[s], sa + S::kA + i * 2048, m, 0);
      }
    }
  } else if (warp == 1) {
    // converged warp, one elected lane issues
    const uint32_t idesc = umma_idesc_bf16(128, 16 * ns, 1, 0);   // A = Src tile viewed MN-major, B = Small^T rows K-major
    const uint32_t smem_base = smem_u32(smem);
    const uint64_t dMN = umma_smem_desc(0, 8192, 1024, kLayoutSW128), dK_ = umma_smem_desc(0, 16, 1024, kLayoutSW128);
    for (int it = 0; it < nsteps; ++it) {
      const int s = it % S::kStages;
      mbar_wait(&full[s], (it / S::kStages) & 1);
      tc_fence_after_sync();
      const uint64_t ad = dMN + ((smem_base + s * S::kStage) >> 4), bd = dK_ + ((smem_base + s * S::kStage + S::kA) >> 4);
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem_base, ad + k * 128, bd + 2 * k, idesc, (it | k) != 0);
        umma_commit(&empty[s]);
      }
      __syncwarp();
    }
    if (elect_one_sync()) umma_commit(acc_full);
    __syncwarp();
  } else if (warp >= 4) {
    const int qd = warp & 3;
    const int prow = p0 + qd * 32 + lane;
    if (nsteps > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after_sync();
      for (int i = 0; i < ns; ++i) {
        uint32_t v[16];
        tmem_ld16(tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + i * 16, v);
        tmem_wait_ld();
        if (prow < it_.P) {
          float* out = it_.out[i];
          if (it_.transposed) {
#pragma unroll
            for (int r = 0; r < 16; ++r) atomicAdd(out + static_cast<size_t>(r) * it_.ldo + prow, __uint_as_float(v[r]));
          } else {
            float* dst = out + static_cast<size_t>(prow) * 16;
#pragma unroll
            for (int g = 0; g < 4; ++g)
              red_add_v4_f32(dst + g * 4, __uint_as_float(v[g * 4]), __uint_as_float(v[g * 4 + 1]),
                             __uint_as_float(v[g * 4 + 2]), __uint_as_float(v[g * 4 + 3]));
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 64);
}

// Host-side description of one item (device pointers).
struct LoraGradDesc {
  const void* src;          // [M, P] bf16, pitch ld_src
  long ld_src;
  int M, P;
  int nsmall;
  const void* small_t[kLgMaxSmall];   // each [16, ld_small] bf16: the side tensor of a fused linear call
  long ld_small;
  float* out[kLgMaxSmall];
  int transposed;
  long ld_out;
};

inline int launch_lora_grad_batch(const LoraGradDesc* d, int n, cudaStream_t stream) {
  if (n <= 0) return 0;
  if (n > kLgMaxItems) return fail("vpt_lora_grad_batch: at most 16 items per call");
  static thread_local LoraGradBatch bp;      // 7 KB: too large for the stack of a ctypes callback thread to be comfortable
  int nmaps = 0, ctas = 0;
  const int target = 4 * 2 * (sm_count() > 0 ? sm_count() : 148);   // ~4 waves of 2 CTAs/SM over the whole batch
  long total_tiles = 0;
  for (int i = 0; i < n; ++i) total_tiles += static_cast<long>((d[i].P + 127) / 128) * ((d[i].M + 63) / 64);
  // rows per CTA: the same for every item, sized so that the batch has about `target` CTAs, at least 8 steps each
  long steps = (total_tiles + target - 1) / target;
  if (steps < 8) steps = 8;
  for (int i = 0; i < n; ++i) {
    const LoraGradDesc& s = d[i];
    if (s.nsmall < 1 || s.nsmall > kLgMaxSmall) return fail("vpt_lora_grad_batch: nsmall must be 1..3");
    if (s.ld_src % 8 != 0 || s.ld_small % 8 != 0 || s.ld_small < s.M) return fail("vpt_lora_grad_batch: bad leading dimension");
    if (nmaps + 1 + s.nsmall > kLgMaxMaps) return fail("vpt_lora_grad_batch: too many tensor maps");
    LoraGradItem& it = bp.items[i];
    it.M = s.M; it.P = s.P; it.nsmall = s.nsmall; it.transposed = s.transposed; it.ldo = static_cast<int>(s.ld_out);
    it.map_src = nmaps;
    if (make_tmap_bf16_2d(&bp.maps[nmaps++], s.src, s.P, s.M, static_cast<uint64_t>(s.ld_src) * 2, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    for (int j = 0; j < kLgMaxSmall; ++j) {
      it.map_small[j] = 0;
      it.out[j] = nullptr;
    }
    for (int j = 0; j < s.nsmall; ++j) {
      if (s.small_t[j] == nullptr || s.out[j] == nullptr) return fail("vpt_lora_grad_batch: null pointer");
      it.map_small[j] = nmaps;
      it.out[j] = s.out[j];
      if (make_tmap_bf16_2d(&bp.maps[nmaps++], s.small_t[j], s.M, 16, static_cast<uint64_t>(s.ld_small) * 2, 64, 16, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    }
    it.ptiles = (s.P + 127) / 128;
    it.rows_per_cta = static_cast<int>(steps) * 64;
    it.cta_begin = ctas;
    ctas += it.ptiles * ((s.M + it.rows_per_cta - 1) / it.rows_per_cta);
  }
  bp.n_items = n;
  static bool attr = false;
  if (!attr) {
    VPT_CUDA_OK(cudaFuncSetAttribute(lora_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LoraGradSmem::kTotal));
    attr = true;
  }
  VPT_CUDA_OK(launch_pdl(lora_grad_kernel, dim3(ctas), dim3(256), LoraGradSmem::kTotal, stream, bp));
  return 0;
}

}  // namespace vpt
