// Host-side helpers shared by the C-ABI entry points: error reporting, tensor-map (TMA descriptor) encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

namespace vpt {

inline std::string& last_error() {
  static thread_local std::string e;
  return e;
}
inline int fail(const char* what, const char* detail = "") {
  last_error() = std::string(what) + (detail[0] ? ": " : "") + detail;
  return 1;
}
#define VPT_CUDA_OK(expr)                                                              \
  do {                                                                                 \
    cudaError_t e__ = (expr);                                                          \
    if (e__ != cudaSuccess) return ::vpt::fail(#expr, cudaGetErrorString(e__));        \
  } while (0)
#define VPT_REQUIRE(cond, msg)                 \
  do {                                         \
    if (!(cond)) return ::vpt::fail(msg, #cond); \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  // cuTensorMapEncodeTiled is a driver call and wants a current context: on a thread whose first CUDA call this is
  // (autograd's backward thread running one of these ops first) the runtime has not bound the primary context yet and
  // the encode fails with CUDA_ERROR_INVALID_CONTEXT -- bind it once per thread.
  static thread_local bool bound = (cudaFree(nullptr), true);
  (void)bound;
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Row-major 2-D bf16 tensor [outer, inner] with row pitch `pitch_bytes`; out-of-bounds elements read as zero.
inline int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                             uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return fail("cuTensorMapEncodeTiled unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (pitch_bytes & 15) != 0)
    return fail("TMA needs a 16-byte aligned base and row pitch");
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[64];
    snprintf(buf, sizeof buf, "CUresult %d", static_cast<int>(r));
    return fail("cuTensorMapEncodeTiled", buf);
  }
  return 0;
}

// Row-major 3-D bf16 tensor [d2, d1, d0] (d0 innermost) with byte strides for d1 and d2.
inline int make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                             uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2,
                             CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return fail("cuTensorMapEncodeTiled unavailable");
  cuuint64_t gdim[3] = {d0, d1, d2};
  cuuint64_t gstride[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[64];
    snprintf(buf, sizeof buf, "CUresult %d", static_cast<int>(r));
    return fail("cuTensorMapEncodeTiled(3d)", buf);
  }
  return 0;
}

// 4-D bf16 tensor, d0 innermost (contiguous); byte strides for d1..d3.
inline int make_tmap_bf16_4d(CUtensorMap* out, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                             const uint32_t box[4], CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return fail("cuTensorMapEncodeTiled unavailable");
  for (int i = 0; i < 3; ++i)
    if (strides_bytes[i] & 15) return fail("TMA needs 16-byte aligned strides");
  if (reinterpret_cast<uintptr_t>(base) & 15) return fail("TMA needs a 16-byte aligned base");
  cuuint64_t gdim[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t gstride[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t b[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstride, b, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[64];
    snprintf(buf, sizeof buf, "CUresult %d", static_cast<int>(r));
    return fail("cuTensorMapEncodeTiled(4d)", buf);
  }
  return 0;
}

// 4-D fp32 tensor, d0 innermost (contiguous); byte strides for d1..d3.
inline int make_tmap_f32_4d(CUtensorMap* out, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                            const uint32_t box[4], CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return fail("cuTensorMapEncodeTiled unavailable");
  for (int i = 0; i < 3; ++i)
    if (strides_bytes[i] & 15) return fail("TMA needs 16-byte aligned strides");
  if (reinterpret_cast<uintptr_t>(base) & 15) return fail("TMA needs a 16-byte aligned base");
  cuuint64_t gdim[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t gstride[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t b[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), gdim, gstride, b, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[64];
    snprintf(buf, sizeof buf, "CUresult %d", static_cast<int>(r));
    return fail("cuTensorMapEncodeTiled(4d f32)", buf);
  }
  return 0;
}

// Launch with programmatic stream serialisation (see pdl_wait / pdl_launch_dependents in sm100.cuh).  VPT_NO_PDL=1 in the
// environment turns the attribute off (A/B measurements).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  static const bool enabled = getenv("VPT_NO_PDL") == nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = enabled ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

}  // namespace vpt
