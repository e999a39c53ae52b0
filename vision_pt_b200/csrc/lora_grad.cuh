// LoRA parameter gradients as one skinny tcgen05 reduction over the token dimension:
//     Out[P, 16] += Src[M, P]^T * Small[M, 16]          (fp32 accumulation, split over M across CTAs)
//   dB = dY^T * Ts        Src = dY [M, N], Small = Ts  [M, 16] -> lora_up.weight.grad   [N, 16]
//   dA^T = X^T * dTs      Src = X  [M, K], Small = dTs [M, 16] -> lora_down.weight.grad [16, K] (written transposed)
// This is the autograd of LoRALinear.forward (/root/reference/src/modules/peft/lora.py:100-104) for lora_down / lora_up.
// The kernel is HBM-bound (reads Src once); both operands are MN-major views of row-major tiles.
#pragma once
#include "host.cuh"
#include "sm100.cuh"

namespace vpt {

struct LoraGradParams {
  int M, P;
  const __nv_bfloat16* small;   // [M, 16]
  float* out;                   // fp32
  int transposed;               // 0: out[p * 16 + r]   1: out[r * ldo + p]
  int ldo;
  int rows_per_cta;             // multiple of 64
};

struct LoraGradSmem {
  static constexpr int kStages = 4;
  static constexpr int kA = 16384, kB = 8192;
  static constexpr int kStage = kA + kB;
  static constexpr int kBars = kStages * kStage;
  static constexpr int kTmemSlot = kBars + (2 * kStages + 1) * 8;
  static constexpr int kTotal = kTmemSlot + 16 + 1024;
};

// tmS: Src [M, P] row-major, box {64 (p), 64 (m)}, SWIZZLE_128B
__global__ void __launch_bounds__(256, 1)
lora_grad_kernel(const __grid_constant__ CUtensorMap tmS, const LoraGradParams p) {
  using S = LoraGradSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::kBars);
  uint64_t* empty = full + S::kStages;
  uint64_t* acc_full = empty + S::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::kTmemSlot);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p0 = blockIdx.x * 128;
  const int m_begin = blockIdx.y * p.rows_per_cta;
  const int m_end = min(p.M, m_begin + p.rows_per_cta);
  const int nsteps = (m_end - m_begin + 63) / 64;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S::kStages; ++s) {
      mbar_init(&full[s], 2);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 32);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < nsteps; ++it) {
        const int s = it % S::kStages;
        mbar_wait(&empty[s], ((it / S::kStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[s], S::kA);
        uint8_t* sa = smem + s * S::kStage;
        tma_load_2d(&tmS, &full[s], sa, p0, m_begin + it * 64);
        tma_load_2d(&tmS, &full[s], sa + 8192, p0 + 64, m_begin + it * 64);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t kIdesc = umma_idesc_bf16(128, 16, 1, 1);
      for (int it = 0; it < nsteps; ++it) {
        const int s = it % S::kStages;
        mbar_wait(&full[s], (it / S::kStages) & 1);
        tc_fence_after_sync();
        const uint32_t sa = smem_u32(smem + s * S::kStage), sb = sa + S::kA;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss(tmem_base, umma_smem_desc(sa + k * 2048, 8192, 1024, kLayoutSW128),
                  umma_smem_desc(sb + k * 2048, 8192, 1024, kLayoutSW128), kIdesc, (it | k) != 0);
        umma_commit(&empty[s]);
      }
      umma_commit(acc_full);
    }
  } else if (warp == 3) {
    // Small rows -> first two 16B chunks (swizzled) of a 128B-pitch MN-major block
    for (int it = 0; it < nsteps; ++it) {
      const int s = it % S::kStages;
      mbar_wait(&empty[s], ((it / S::kStages) & 1) ^ 1);
      const uint32_t sb = smem_u32(smem + s * S::kStage + S::kA);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = lane + u * 32;            // 0..127 = 64 rows x 2 chunks
        const int r = idx >> 1, c = idx & 1;
        const int m = m_begin + it * 64 + r;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (m < m_end) v = __ldg(reinterpret_cast<const uint4*>(p.small + static_cast<size_t>(m) * 16 + c * 8));
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + r * 128 + ((c ^ (r & 7)) * 16)), "r"(v.x),
                     "r"(v.y), "r"(v.z), "r"(v.w)
                     : "memory");
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
    }
  } else if (warp >= 4) {
    const int qd = warp & 3;
    const int prow = p0 + qd * 32 + lane;
    if (nsteps > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after_sync();
      uint32_t v[16];
      tmem_ld16(tmem_base + (static_cast<uint32_t>(qd * 32) << 16), v);
      tmem_wait_ld();
      if (prow < p.P) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          float* dst = p.transposed ? p.out + static_cast<size_t>(r) * p.ldo + prow : p.out + static_cast<size_t>(prow) * 16 + r;
          atomicAdd(dst, __uint_as_float(v[r]));
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 32);
}

// The rows of Src outside [m_begin, m_end) that a 64-row TMA box may touch belong to the next CTA's range: they are
// multiplied by zero rows of Small (the copier zero-fills m >= m_end), so every row is counted exactly once.
inline int launch_lora_grad(const void* src, int lds, const void* small, float* out, int M, int P, int transposed, int ldo,
                            cudaStream_t stream) {
  CUtensorMap tm;
  if (make_tmap_bf16_2d(&tm, src, P, M, static_cast<uint64_t>(lds) * 2, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  LoraGradParams p{};
  p.M = M; p.P = P;
  p.small = static_cast<const __nv_bfloat16*>(small);
  p.out = out; p.transposed = transposed; p.ldo = ldo;
  const int ptiles = (P + 127) / 128;
  int splits = (2 * sm_count() + ptiles - 1) / ptiles;
  int rows = ((M + splits - 1) / splits + 63) / 64 * 64;
  if (rows < 64) rows = 64;
  splits = (M + rows - 1) / rows;
  p.rows_per_cta = rows;
  static bool attr = false;
  if (!attr) {
    VPT_CUDA_OK(cudaFuncSetAttribute(lora_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LoraGradSmem::kTotal));
    attr = true;
  }
  lora_grad_kernel<<<dim3(ptiles, splits), 256, LoraGradSmem::kTotal, stream>>>(tm, p);
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace vpt
