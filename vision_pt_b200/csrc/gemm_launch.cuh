// Host launcher for gemm_nf4lora_kernel: picks the tile width, encodes the tensor maps, launches one persistent
// CTA per SM.
#pragma once
#include "gemm_nf4lora.cuh"
#include "host.cuh"

namespace vpt {

struct GemmLaunch {
  bool bwd, nf4, lora;
  int bn;                        // 0 = choose
  const void* act;               // A operand [M, R] bf16
  int lda;
  const void* w_bf16;            // [N, K] bf16 when !nf4
  long ldw;                      // row pitch of w_bf16 in elements (0 = K)
  GemmParams p;                  // M, NO, R, D, ldd, bias, w, lora_*, scale, side filled by the caller
  int max_ctas;                  // 0 = all SMs
};

template <int BN, bool kBwd, bool kNF4, bool kLoRA, bool kRagged>
int launch_gemm_t(const GemmLaunch& g, cudaStream_t stream) {
  using S = GemmSmem<BN, kBwd, kLoRA>;
  GemmParams p = g.p;
  p.num_m_tiles = (p.M + kBM - 1) / kBM;
  p.num_n_tiles = (p.NO + BN - 1) / BN;
  CUtensorMap tmA, tmB, tmP;
  if (make_tmap_bf16_2d(&tmA, g.act, p.R, p.M, static_cast<uint64_t>(g.lda) * 2, 64, kBM, CU_TENSOR_MAP_SWIZZLE_128B))
    return 1;
  tmB = tmA;
  tmP = tmA;
  if (!kNF4) {
    if (make_tmap_bf16_2d(&tmB, g.w_bf16, p.w.K, p.w.N, static_cast<uint64_t>(g.ldw > 0 ? g.ldw : p.w.K) * 2, 64, kBwd ? 64 : BN,
                          CU_TENSOR_MAP_SWIZZLE_128B))
      return 1;
  }
  if (!kBwd && kLoRA) {
    if (make_tmap_bf16_2d(&tmP, p.lora_down, p.w.K, kRank, static_cast<uint64_t>(p.ld_down) * 2, 64, kRank,
                          CU_TENSOR_MAP_SWIZZLE_128B))
      return 1;
  }
  auto kern = gemm_nf4lora_kernel<BN, kBwd, kNF4, kLoRA, kRagged>;
  static bool attr_set = false;
  if (!attr_set) {
    VPT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    attr_set = true;
  }
  int ctas = p.num_m_tiles * p.num_n_tiles;
  const int cap = g.max_ctas > 0 ? g.max_ctas : sm_count();
  if (ctas > cap) ctas = cap;
  kern<<<ctas, kGemmThreads, S::kTotal, stream>>>(tmA, tmB, tmP, p);
  VPT_CUDA_OK(cudaGetLastError());
  return 0;
}

// Tile width: the one that wastes the fewest tensor-core cycles over whole waves of the persistent grid.
inline int choose_bn(int M, int NO, bool lora) {
  const int sms = sm_count() > 0 ? sm_count() : 148;
  const int cands[2] = {192, 128};
  int best = 128;
  double best_cost = 1e30;
  for (int c = 0; c < 2; ++c) {
    const int bn = cands[c];
    const long tiles = static_cast<long>((M + kBM - 1) / kBM) * ((NO + bn - 1) / bn);
    const long rounds = (tiles + sms - 1) / sms;
    const double cost = static_cast<double>(rounds) * (bn + (lora ? kRank : 0));   // ~ MMA cycles per k-step
    if (cost < best_cost) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

template <bool kBwd, bool kNF4, bool kLoRA>
int launch_gemm_bn(const GemmLaunch& g, cudaStream_t stream) {
  const int bn = g.bn > 0 ? g.bn : choose_bn(g.p.M, g.p.NO, kLoRA);
  const bool ragged = kNF4 && g.p.w.packed_rows != nullptr;
  if (kNF4 && ragged) {
    if (bn == 192) return launch_gemm_t<192, kBwd, kNF4, kLoRA, kNF4>(g, stream);
    if (bn == 128) return launch_gemm_t<128, kBwd, kNF4, kLoRA, kNF4>(g, stream);
  } else {
    if (bn == 192) return launch_gemm_t<192, kBwd, kNF4, kLoRA, false>(g, stream);
    if (bn == 128) return launch_gemm_t<128, kBwd, kNF4, kLoRA, false>(g, stream);
  }
  return fail("unsupported tile width");
}

inline int launch_gemm(const GemmLaunch& g, cudaStream_t stream) {
  const int key = (g.bwd ? 4 : 0) | (g.nf4 ? 2 : 0) | (g.lora ? 1 : 0);
  switch (key) {
    case 0: return launch_gemm_bn<false, false, false>(g, stream);
    case 1: return launch_gemm_bn<false, false, true>(g, stream);
    case 2: return launch_gemm_bn<false, true, false>(g, stream);
    case 3: return launch_gemm_bn<false, true, true>(g, stream);
    case 4: return launch_gemm_bn<true, false, false>(g, stream);
    case 5: return launch_gemm_bn<true, false, true>(g, stream);
    case 6: return launch_gemm_bn<true, true, false>(g, stream);
    default: return launch_gemm_bn<true, true, true>(g, stream);
  }
}

}  // namespace vpt
