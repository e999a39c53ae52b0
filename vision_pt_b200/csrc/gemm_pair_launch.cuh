// Host launcher of gemm_pair_kernel: tensor maps, tile width, one CTA pair per TPC.
#pragma once
#include "gemm_pair.cuh"
#include <stdlib.h>

#include "host.cuh"

namespace vpt {

struct PairLaunch {
  const void* act;        // A [M, R] bf16, pitch lda
  int lda;
  const void* w;          // Bw [NO, R] bf16, pitch ldw
  long ldw;
  const void* p_rows;     // P [16, R] bf16, pitch ldp (nullptr = no LoRA)
  long ldp;
  void* out;              // D [M, NO], pitch ldd
  int ldd;
  int bn;                 // 0 = choose
  int epi;                // epilogue mode (gemm_pair.cuh): 0 plain, 1 SwiGLU forward, 2 SwiGLU backward
  void* out2;             // D2 [M, NO], pitch ldd2 (modes 1, 2)
  int ldd2;
  const void* in2;        // in2 [M, NO], pitch ldr2 (mode 2)
  int ldr2;
  PairParams p;           // M, NO, R, bias, residual (in1), ldr, q_rows, scale, side
};

template <int BN, bool kLoRA, int kEpi>
int launch_pair_t(const PairLaunch& g, cudaStream_t stream) {
  using S = PairSmem<BN, kLoRA, kEpi>;
  PairParams p = g.p;
  p.num_m_pairs = (p.M + 255) / 256;
  p.num_n_tiles = (p.NO + BN - 1) / BN;
  CUtensorMap tmA, tmB0, tmB1, tmP, tmD, tmR, tmD2, tmR2;
  const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
  if (make_tmap_bf16_2d(&tmA, g.act, p.R, p.M, static_cast<uint64_t>(g.lda) * 2, 64, 128, sw)) return 1;
  if (make_tmap_bf16_2d(&tmB0, g.w, p.R, p.NO, static_cast<uint64_t>(g.ldw) * 2, 64, S::kNH, sw)) return 1;
  if (make_tmap_bf16_2d(&tmB1, g.w, p.R, p.NO, static_cast<uint64_t>(g.ldw) * 2, 64, kLoRA ? S::kNH - kPairRank : S::kNH, sw)) return 1;
  tmP = tmA;
  const int n_sec = p.sec_n > 0 ? (p.NO + p.sec_n - 1) / p.sec_n : 1;
  if (kLoRA && make_tmap_bf16_2d(&tmP, g.p_rows, p.R, static_cast<uint64_t>(kPairRank) * n_sec, static_cast<uint64_t>(g.ldp) * 2, 64, kPairRank, sw)) return 1;
  if (make_tmap_bf16_2d(&tmD, g.out, p.NO, p.M, static_cast<uint64_t>(g.ldd) * 2, 64, 32, sw)) return 1;
  tmR = tmD;
  if (p.residual != nullptr && make_tmap_bf16_2d(&tmR, p.residual, p.NO, p.M, static_cast<uint64_t>(p.ldr) * 2, 64, 32, sw)) return 1;
  tmD2 = tmD;
  tmR2 = tmD;
  if (kEpi != 0 && make_tmap_bf16_2d(&tmD2, g.out2, p.NO, p.M, static_cast<uint64_t>(g.ldd2) * 2, 64, 32, sw)) return 1;
  if (kEpi == 2 && make_tmap_bf16_2d(&tmR2, g.in2, p.NO, p.M, static_cast<uint64_t>(g.ldr2) * 2, 64, 32, sw)) return 1;
  auto kern = gemm_pair_kernel<BN, kLoRA, kEpi>;
  static bool attr_set = false;
  if (!attr_set) {
    VPT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    attr_set = true;
    if (getenv("VPT_DEBUG") != nullptr) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * (sm_count() / 2));
      cfg.blockDim = dim3(kPairThreads);
      cfg.dynamicSmemBytes = S::kTotal;
      int nclusters = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg);
      fprintf(stderr, "[vpt] gemm_pair<%d,%d,%d>: smem %d B, %d stages, max active clusters %d (%s)\n", BN, int(kLoRA), kEpi, S::kTotal,
              S::kStages, nclusters, cudaGetErrorString(e));
    }
  }
  int pairs = p.num_m_pairs * p.num_n_tiles;
  const int cap = sm_count() / 2;
  if (pairs > cap) pairs = cap;
  VPT_CUDA_OK(launch_pdl(kern, dim3(2 * pairs), dim3(kPairThreads), S::kTotal, stream, tmA, tmB0, tmB1, tmP, tmD, tmR, tmD2, tmR2, p));
  return 0;
}

// Tile width: whole waves of the 74 pairs x UMMA N, with the narrower tile charged for its extra L2 traffic.
inline int choose_pair_bn(int M, int NO, bool lora, int sec_n = 0) {
  const int pairs = (sm_count() > 0 ? sm_count() : 148) / 2;
  const int cands[2] = {192, 128};
  int best = 0;
  double best_cost = 1e30;
  for (int c = 0; c < 2; ++c) {
    const int bn = cands[c];
    if (sec_n > 0 && sec_n % bn != 0) continue;          // a tile must lie inside one section
    const long tiles = static_cast<long>((M + 255) / 256) * ((NO + bn - 1) / bn);
    const long rounds = (tiles + pairs - 1) / pairs;
    const double cost = static_cast<double>(rounds) * (bn + (lora ? kPairRank : 0)) * (bn == 128 ? 1.12 : 1.0);
    if (cost < best_cost) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

inline int launch_pair(const PairLaunch& g, cudaStream_t stream) {
  const bool lora = g.p_rows != nullptr;
  const int bn = g.bn > 0 ? g.bn : choose_pair_bn(g.p.M, g.p.NO, lora, g.p.sec_n);
  if (bn != 192 && bn != 128) return fail("unsupported tile width (sections must be multiples of 128)");
  if (g.p.sec_n > 0 && (g.p.sec_n % bn != 0 || g.p.NO % g.p.sec_n != 0)) return fail("sections: the section width must divide the output width and be a multiple of the tile width");
  if (g.epi != 0) {
    if (g.p.residual == nullptr || g.out2 == nullptr || (g.epi == 2 && g.in2 == nullptr))
      return fail("fused SwiGLU epilogue: in1 (residual) / out2 / in2 missing");
  }
#define VPT_PAIR_CASE(BN_, LORA_, EPI_) \
  if (bn == BN_ && lora == LORA_ && g.epi == EPI_) return launch_pair_t<BN_, LORA_, EPI_>(g, stream);
  VPT_PAIR_CASE(192, true, 0) VPT_PAIR_CASE(128, true, 0) VPT_PAIR_CASE(192, false, 0) VPT_PAIR_CASE(128, false, 0)
  VPT_PAIR_CASE(192, true, 1) VPT_PAIR_CASE(128, true, 1) VPT_PAIR_CASE(192, false, 1) VPT_PAIR_CASE(128, false, 1)
  VPT_PAIR_CASE(192, true, 2) VPT_PAIR_CASE(128, true, 2) VPT_PAIR_CASE(192, false, 2) VPT_PAIR_CASE(128, false, 2)
#undef VPT_PAIR_CASE
  return fail("unsupported epilogue mode");
}

}  // namespace vpt
