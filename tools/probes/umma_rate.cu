// Cycles per tcgen05.mma for the operand shapes the attention kernels use (one issuing warp per SM, all SMs busy).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I vision_pt_b200/csrc -o build/umma_rate tools/probes/umma_rate.cu
#include <cstdio>
#include "sm100.cuh"
using namespace vpt;

// mode: 0 SS K-major/K-major, 1 SS A K-major B MN-major, 2 TS B MN-major, 3 SS A MN-major B MN-major, 4 TS B K-major
template <int mode, int N, int nacc>
__global__ void __launch_bounds__(128, 1) rate_kernel(int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm0 = slot;
  const uint32_t tm = tm0;
  if (warp == 1) {
    const uint32_t base = smem_u32(smem);
    const uint64_t dK_ = umma_smem_desc(0, 16, 1024, kLayoutSW128);
    const uint64_t dMN = umma_smem_desc(0, 8192, 1024, kLayoutSW128);
    const uint64_t dMNq = umma_smem_desc(0, 16384, 1024, kLayoutSW128);
    const uint64_t a_k = dK_ + (base >> 4), b_k = dK_ + ((base + 65536) >> 4);
    const uint64_t a_mn = dMNq + (base >> 4), b_mn = dMN + ((base + 65536) >> 4);
    constexpr uint32_t id = umma_idesc_bf16(128, N, mode == 3, mode == 1 || mode == 2 || mode == 3);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 3; ++rep) {
      t0 = clock64();
      if (elect_one_sync()) {
        for (int r0 = 0; r0 < reps; r0 += 8) {
#pragma unroll
         for (int r = 0; r < 8; ++r) {
          const int k = r & 3;
          const uint32_t tm = tm0 + (r % nacc) * N;   // independent accumulators, round robin
          if (mode == 0) umma_ss(tm, a_k + 2 * k, b_k + 2 * k, id, 1);
          else if (mode == 1) umma_ss(tm, a_k + 2 * k, b_mn + k * 128, id, 1);
          else if (mode == 2) umma_ts(tm, tm0 + 256 + k * 8, b_mn + k * 128, id, 1);
          else if (mode == 3) umma_ss(tm, a_mn + k * 128, b_mn + k * 128, id, 1);
          else umma_ts(tm, tm0 + 256 + k * 8, b_k + 2 * k, id, 1);
         }
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, rep & 1);
      t1 = clock64();
    }
    if (blockIdx.x == 0 && threadIdx.x == 32) *out = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// The attention-backward issue pattern of one 128 x 128 tile, alone on the SM (no softmax / drain warps, no TMA):
// variant 0: as the kernel issues it; 1: dV and dK not interleaved (4 TS then 4 SS); 2: all SS (P^T in shared memory);
// 3: S / dP as one N=128 MMA per k-step
template <int variant, int hammer>
__global__ void __launch_bounds__(384, 1) tile_pattern_kernel(int tiles, long long* out, int random_data) {
  __shared__ volatile int done_flag;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[5];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { for (int i = 0; i < 5; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); done_flag = 0; }
  {
    // operands: zeros, or random bf16 in (-2, 2) (sign / exponent 0x3f.. / random mantissa), like real activations
    uint32_t* w = reinterpret_cast<uint32_t*>(smem);
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) {
      uint32_t h = (i + 1) * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      w[i] = random_data ? ((h & 0x807f807fu) | 0x3f003f00u | ((h >> 3) & 0x00800080u)) : 0u;
    }
    fence_proxy_async_smem();
  }
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = slot;
  if (warp == 1) {
    const uint32_t base = smem_u32(smem);
    const uint64_t dK_ = umma_smem_desc(0, 16, 1024, kLayoutSW128);
    const uint64_t dMN = umma_smem_desc(0, 8192, 1024, kLayoutSW128);
    const uint64_t dMNq = umma_smem_desc(0, 16384, 1024, kLayoutSW128);
    const uint64_t kd = dK_ + (base >> 4), vd = dK_ + ((base + 16384) >> 4), qd = dK_ + ((base + 32768) >> 4), dod = dK_ + ((base + 49152) >> 4);
    const uint64_t pt = dK_ + ((base + 65536) >> 4), dst = dK_ + ((base + 98304) >> 4), dstq = dMNq + ((base + 98304) >> 4);
    const uint64_t qmn = dMN + ((base + 32768) >> 4), domn = dMN + ((base + 49152) >> 4), kmn = dMN + (base >> 4);
    constexpr uint32_t idS = umma_idesc_bf16(128, variant == 3 ? 128 : 64, 0, 0), idKM = umma_idesc_bf16(128, 64, 0, 1), idMM = umma_idesc_bf16(128, 64, 1, 1);
    const uint32_t tST = tm, tDPT = tm + 128, tDV = tm + 256, tDK = tm + 320, tDQ = tm + 384, tPT = tm + 448;
    long long t0 = clock64();
    long long issue = 0;
    for (int t = 0; t < tiles; ++t) {
      const long long ti0 = clock64();
      if (variant >= 4) {
#pragma unroll 1
        for (int X = 0; X < 2; ++X) {
          tc_fence_after_sync();
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss(tST + X * 64, kd + 2 * k, qd + X * 512 + 2 * k, idS, k != 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss(tDPT + X * 64, vd + 2 * k, dod + X * 512 + 2 * k, idS, k != 0);
            umma_commit(&bar[X]);
            if (variant == 5) umma_commit(&bar[X]);
          }
          __syncwarp();
          tc_fence_after_sync();
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_ts(tDV, tPT + X * 32 + k * 8, domn + (X * 4 + k) * 128, idKM, 1);
              umma_ss(tDK, dst + X * 1024 + 2 * k, qmn + (X * 4 + k) * 128, idKM, 1);
            }
            umma_commit(&bar[2]);
            if (variant == 5) { umma_commit(&bar[2]); umma_commit(&bar[2]); }
          }
          __syncwarp();
        }
        tc_fence_after_sync();
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 8; ++k) umma_ss(tDQ, dstq + k * 128, kmn + k * 128, idMM, k != 0);
          umma_commit(&bar[3 + (t & 1)]);
        }
      } else
      if (elect_one_sync()) {
        if (variant == 3) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tST, kd + 2 * k, qd + 2 * k, idS, k != 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tDPT, vd + 2 * k, dod + 2 * k, idS, k != 0);
          umma_commit(&bar[0]);
        } else {
#pragma unroll
          for (int X = 0; X < 2; ++X) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss(tST + X * 64, kd + 2 * k, qd + X * 512 + 2 * k, idS, k != 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss(tDPT + X * 64, vd + 2 * k, dod + X * 512 + 2 * k, idS, k != 0);
            umma_commit(&bar[X]);
          }
        }
#pragma unroll
        for (int X = 0; X < 2; ++X) {
          if (variant == 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ts(tDV, tPT + X * 32 + k * 8, domn + (X * 4 + k) * 128, idKM, 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss(tDK, dst + X * 1024 + 2 * k, qmn + (X * 4 + k) * 128, idKM, 1);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (variant == 2) umma_ss(tDV, pt + X * 1024 + 2 * k, domn + (X * 4 + k) * 128, idKM, 1);
              else umma_ts(tDV, tPT + X * 32 + k * 8, domn + (X * 4 + k) * 128, idKM, 1);
              umma_ss(tDK, dst + X * 1024 + 2 * k, qmn + (X * 4 + k) * 128, idKM, 1);
            }
          }
          umma_commit(&bar[2]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_ss(tDQ, dstq + k * 128, kmn + k * 128, idMM, k != 0);
        umma_commit(&bar[3 + (t & 1)]);
      }
      __syncwarp();
      issue += clock64() - ti0;
      if (t > 0) mbar_wait(&bar[3 + ((t - 1) & 1)], ((t - 1) >> 1) & 1);   // at most two tiles in flight
    }
    mbar_wait(&bar[3 + ((tiles - 1) & 1)], ((tiles - 1) >> 1) & 1);
    done_flag = 1;
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 32) { out[0] = t1 - t0; out[1] = issue; }
  }
  if (warp >= 4 && hammer != 0) {
    // background traffic of the kind the softmax / drain warps generate
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t row = smem_u32(smem) + 131072 + (warp - 4) * 4096 + lane * 128;
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    uint32_t acc = 0;
    while (!done_flag) {
      if (hammer == 1) {          // 8 x STS.128 per thread, swizzled (conflict free), then ~idle arithmetic
#pragma unroll
        for (int g = 0; g < 8; ++g)
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(row + ((g ^ (lane & 7)) * 16)), "r"(acc) : "memory");
        acc += 1;
      } else if (hammer == 2) {   // TMEM reads of the accumulator columns (values unused)
        uint32_t v[32];
        tmem_ld32(tm + lane_off + ((acc & 7) * 32), v);
        tmem_wait_ld();
        acc += v[0] & 1 ? 1 : 1;
      }
      if (hammer == 3) {          // a long straight-line body (about 24 KB of SASS per pass): instruction-cache pressure
        float x = __uint_as_float(acc | 0x3f800000u);
#pragma unroll
        for (int u = 0; u < 1536; ++u) x = fmaf(x, 1.0001f + u * 1e-7f, 0.5f);
        acc += __float_as_uint(x) & 1;
      }
      if (hammer == 1) __nanosleep(0);
    }
    if (acc == 0xffffffffu) *out = acc;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  
  const char* names[] = {"SS A K-major, B K-major", "SS A K-major, B MN-major", "TS B MN-major", "SS A MN-major, B MN-major", "TS B K-major"};
  const int reps = 512;
#define RUN(mode, N, nacc) { cudaFuncSetAttribute(rate_kernel<mode, N, nacc>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); rate_kernel<mode, N, nacc><<<148, 128, 200 * 1024>>>(reps, d); long long c = 0; \
    cudaError_t e = cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost); \
    printf("%-28s M=128 N=%3d K=16 accumulators %d: %6.1f cyc/MMA (floor %d)%s\n", names[mode], N, nacc, double(c) / reps, N / 2, e ? cudaGetErrorString(e) : ""); }
#define RUNM(mode) RUN(mode, 64, 1) RUN(mode, 64, 2) RUN(mode, 64, 4) RUN(mode, 128, 1) RUN(mode, 128, 2) RUN(mode, 256, 1)
  RUNM(0) RUNM(1) RUNM(2) RUNM(3) RUNM(4)
#define PAT(v, name) PATH(v, 0, name)
#define PATX(v, name) PATH(v, 0, name) PATH(v, 1, name " + 8 warps STS.128") PATH(v, 2, name " + 8 warps tcgen05.ld") PATH(v, 3, name " + 8 warps long FMA body")
#define PAT_UNUSED(v, name) PATH(v, 0, name) PATH(v, 1, name " + 8 warps STS.128") PATH(v, 2, name " + 8 warps tcgen05.ld") PATH(v, 3, name " + 8 warps long FMA body")
#define PATH(v, hm, name) { cudaFuncSetAttribute(tile_pattern_kernel<v, hm>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
    tile_pattern_kernel<v, hm><<<148, 384, 200 * 1024>>>(48, d, rnd); long long c[2] = {0, 0}; cudaError_t e = cudaMemcpy(c, d, 16, cudaMemcpyDeviceToHost); \
    printf("backward tile pattern, %-64s: %7.0f cyc/tile (issue %5.0f)%s\n", name, double(c[0]) / 48, double(c[1]) / 48, e ? cudaGetErrorString(e) : ""); }
  for (int rnd = 0; rnd < 2; ++rnd) {
  printf("-- shared-memory operands: %s\n", rnd ? "random bf16" : "zeros");
  PAT(0, "as issued (dV TS / dK SS interleaved)") PAT(1, "dV x4 TS then dK x4 SS") PAT(2, "all SS (P^T in shared memory)") PAT(3, "S, dP as N=128") PAT(4, "as issued, fenced groups") PAT(5, "fenced groups, 11 commits per tile")
  }
  return 0;
}
