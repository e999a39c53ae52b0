// Optimiser step over the flat LoRA parameter buffer: gradient-norm clipping + AdamW in one pass.
//
// Reference semantics (file:line under /root/reference):
//   clip_grad_norm_ after backward         src/models/for_training.py:98-109 (accelerator.clip_grad_norm_ -> torch
//                                          clip_grad_norm_: coef = min(1, max_norm / (total_norm + 1e-6)))
//   optimizer.step / zero_grad             src/trainer/common.py:382-388 (torch.optim.AdamW semantics, decoupled decay)
// Only the LoRA matrices train, so all parameters live in ONE contiguous bf16 buffer with an fp32 gradient buffer of
// the same length (the lora_grad kernels accumulate into it; with data parallelism it is the all-reduce payload).
// Both kernels are HBM-bound streaming passes; nothing here synchronises the host: the step number and the squared
// gradient norm are read from device memory so that a captured CUDA graph replays correctly.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vpt {

// out[0] += sum((g * scale)^2)   (out must be zero on entry)
__global__ void __launch_bounds__(256)
grad_sumsq_kernel(const float* __restrict__ g, long n, float scale, float* __restrict__ out) {
  float acc = 0.f;
  const long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(g4 + i);
    acc += (v.x * scale) * (v.x * scale) + (v.y * scale) * (v.y * scale) + (v.z * scale) * (v.z * scale) + (v.w * scale) * (v.w * scale);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = g[(n4 << 2) + threadIdx.x] * scale;
    acc += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += s[i];
    atomicAdd(out, t);
  }
}

struct AdamWParams {
  __nv_bfloat16* p;        // [n] parameters (updated in place)
  float* g;                // [n] gradients; zeroed after use when zero_grad != 0
  float* m;                // [n] first moment
  float* v;                // [n] second moment
  long n;
  float lr, beta1, beta2, eps, weight_decay;
  float grad_scale;        // 1 / world_size for a SUM all-reduce, 1 otherwise
  const float* sumsq;      // device scalar: squared norm of (g * grad_scale); nullptr = no clipping
  float max_norm;
  const float* step;       // device scalar: 1-based step number of THIS update
  int zero_grad;
};

__global__ void __launch_bounds__(256)
adamw_kernel(const AdamWParams a) {
  float coef = a.grad_scale;
  if (a.sumsq != nullptr) {
    const float norm = sqrtf(__ldg(a.sumsq));
    coef *= fminf(1.f, a.max_norm / (norm + 1e-6f));
  }
  const float t = __ldg(a.step);
  const float bc1 = 1.f - powf(a.beta1, t);
  const float bc2 = 1.f - powf(a.beta2, t);
  const float step_size = a.lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const float decay = 1.f - a.lr * a.weight_decay;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float g = a.g[i] * coef;
    const float m = a.beta1 * a.m[i] + (1.f - a.beta1) * g;
    const float v = a.beta2 * a.v[i] + (1.f - a.beta2) * g * g;
    a.m[i] = m;
    a.v[i] = v;
    float p = __bfloat162float(a.p[i]) * decay;
    p -= step_size * m / (sqrtf(v) * inv_sqrt_bc2 + a.eps);
    a.p[i] = __float2bfloat16_rn(p);
    if (a.zero_grad) a.g[i] = 0.f;
  }
}

// ---- schedulefree.RAdamScheduleFree (the optimiser of the shipped YAMLs: configs/jit/x-loss/config.yml:75; package
// schedulefree 1.4.1, uv.lock:3428 -- not in this image, so the update below restates the published algorithm of
// radam_schedulefree.py and is checked against this repo's CPU restatement of it only: parity unpinned).
// The parameters are the "y" sequence (what the model trains with), z is the base sequence (fp32 here; the package
// keeps it in the parameter dtype), exp_avg_sq the second moment.  Per step, on the device so that a graph replays:
//   sched = {k, lr_max, weight_sum, scheduled_lr}  ->  coef = {lr_t, ckp1, adaptive_y_lr, bias_correction2, adam}
__global__ void radam_sf_advance_kernel(double* sched, float* coef, double lr, double beta1, double beta2, double r,
                                        double weight_lr_power, int silent_sgd_phase) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double step = sched[0] + 1.0;
  const double beta2_t = pow(beta2, step);
  const double bc2 = 1.0 - beta2_t;
  const double rho_inf = 2.0 / (1.0 - beta2) - 1.0;
  const double rho_t = rho_inf - 2.0 * step * beta2_t / bc2;
  const bool adam = rho_t > 4.0;
  const double rect = adam ? sqrt((rho_t - 4.0) * (rho_t - 2.0) * rho_inf / ((rho_inf - 4.0) * (rho_inf - 2.0) * rho_t))
                           : (silent_sgd_phase ? 0.0 : 1.0);
  const double lr_t = lr * rect;
  const double lr_max = fmax(lr_t, sched[1]);
  const double weight = pow(step, r) * pow(lr_max, weight_lr_power);
  const double weight_sum = sched[2] + weight;
  const double ckp1 = weight_sum > 0.0 ? weight / weight_sum : 0.0;
  sched[0] = step;
  sched[1] = lr_max;
  sched[2] = weight_sum;
  sched[3] = lr_t;                                   // param_group["scheduled_lr"] (logged by src/trainer/common.py:499-506)
  coef[0] = static_cast<float>(lr_t);
  coef[1] = static_cast<float>(ckp1);
  coef[2] = static_cast<float>(lr_t * (beta1 * (1.0 - ckp1) - 1.0));
  coef[3] = static_cast<float>(bc2);
  coef[4] = adam ? 1.f : 0.f;
}

struct RAdamSFParams {
  __nv_bfloat16* y;        // [n] parameters (the y sequence), updated in place
  float* g;                // [n] gradients; zeroed after use when zero_grad != 0
  float* z;                // [n] base sequence
  float* v;                // [n] second moment
  long n;
  float beta2, eps, weight_decay;
  float grad_scale;
  const float* sumsq;
  float max_norm;
  const float* coef;       // written by radam_sf_advance_kernel just before
  int zero_grad;
};

__global__ void __launch_bounds__(256)
radam_sf_kernel(const RAdamSFParams a) {
  float clip = a.grad_scale;
  if (a.sumsq != nullptr) clip *= fminf(1.f, a.max_norm / (sqrtf(__ldg(a.sumsq)) + 1e-6f));
  const float lr_t = __ldg(a.coef), ckp1 = __ldg(a.coef + 1), y_lr = __ldg(a.coef + 2), bc2 = __ldg(a.coef + 3);
  const bool adam = __ldg(a.coef + 4) != 0.f;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float g = a.g[i] * clip;
    const float v = a.beta2 * a.v[i] + (1.f - a.beta2) * g * g;
    a.v[i] = v;
    float y = __bfloat162float(a.y[i]);
    float gn = adam ? g / (sqrtf(v / bc2) + a.eps) : g;
    gn += a.weight_decay * y;                       // weight decay is evaluated at y
    const float z = a.z[i];
    y = y + ckp1 * (z - y);                         // y.lerp_(z, ckp1)
    y += y_lr * gn;
    a.y[i] = __float2bfloat16_rn(y);
    a.z[i] = z - lr_t * gn;
    if (a.zero_grad) a.g[i] = 0.f;
  }
}

// optimizer.eval() / .train() of the package: parameters y <-> x = the averaged iterate (what is evaluated and saved)
//   to_eval != 0: p.lerp_(z, 1 - 1/beta1);  else: p.lerp_(z, 1 - beta1)
__global__ void __launch_bounds__(256)
radam_sf_swap_kernel(__nv_bfloat16* p, const float* z, long n, float weight) {
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float y = __bfloat162float(p[i]);
    p[i] = __float2bfloat16_rn(y + weight * (z[i] - y));
  }
}

}  // namespace vpt

// ------------------------------------------------------------------------------------------- flow-matching loss
// treat_loss (train/jit/class_to_image.py:106-164) for model_pred == "image", forward and gradient in one pass:
//   mode 0 (loss_target "image")    : loss = mean((pred - clean)^2)
//   mode 1 (loss_target "velocity") : loss = mean(((pred - noisy)/d - (clean - noisy)/d)^2),  d = max(1 - t[b], clamp_eps)
//                                     (JiT pipeline image_to_velocity, src/models/jit/pipeline.py:253-260)
// loss_out[0] += partial sums (zero on entry); dpred = d loss / d pred (bf16), to be scaled by the upstream gradient.
namespace vpt {

__device__ __forceinline__ float loss_load(const void* p, int dtype, long i) {
  if (dtype == 0) return __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]);
  if (dtype == 1) return __half2float(static_cast<const __half*>(p)[i]);
  return static_cast<const float*>(p)[i];
}

// eight consecutive elements of a bf16 / fp16 / fp32 tensor as floats (i8 = index of the group of eight)
__device__ __forceinline__ void loss_load8(const void* p, int dtype, long i8, float (&f)[8]) {
  if (dtype == 2) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p) + 2 * i8), b = __ldg(reinterpret_cast<const float4*>(p) + 2 * i8 + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    return;
  }
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p) + i8);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (dtype == 0) {
      f[2 * k] = __uint_as_float(w[k] << 16);
      f[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
    } else {
      const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
      f[2 * k] = h.x;
      f[2 * k + 1] = h.y;
    }
  }
}

__global__ void __launch_bounds__(256)
flow_loss_kernel(const __nv_bfloat16* __restrict__ pred, const void* __restrict__ clean, const void* __restrict__ noisy,
                 int in_dtype, const float* __restrict__ timestep, long per_sample, long total, int mode, float clamp_eps,
                 float* __restrict__ loss_out, __nv_bfloat16* __restrict__ dpred) {
  const float inv_n = 1.f / static_cast<float>(total);
  float acc = 0.f;
  const long tid = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x, nthreads = static_cast<long>(gridDim.x) * blockDim.x;
  const bool aligned = ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(clean) | reinterpret_cast<uintptr_t>(noisy) |
                         reinterpret_cast<uintptr_t>(dpred)) & 15) == 0;
  if ((per_sample & 7) == 0 && aligned) {
    // HBM-bound streaming pass: 16-byte accesses, eight elements of one sample per thread and iteration
    for (long i8 = tid; i8 < (total >> 3); i8 += nthreads) {
      float pf[8], cf[8], zf[8];
      loss_load8(pred, 0, i8, pf);
      loss_load8(clean, in_dtype, i8, cf);
      float d = 1.f, w = 1.f;
      if (mode == 1) {
        loss_load8(noisy, in_dtype, i8, zf);
        d = fmaxf(1.f - __ldg(timestep + (i8 << 3) / per_sample), clamp_eps);
        w = 1.f / d;
      }
      uint32_t out[4];
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        float df[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          df[e] = mode == 1 ? (pf[k + e] - zf[k + e]) / d - (cf[k + e] - zf[k + e]) / d : pf[k + e] - cf[k + e];
          acc += df[e] * df[e];
        }
        const __nv_bfloat162 o2 = __floats2bfloat162_rn(2.f * df[0] * w * inv_n, 2.f * df[1] * w * inv_n);
        out[k >> 1] = *reinterpret_cast<const uint32_t*>(&o2);
      }
      if (dpred != nullptr) reinterpret_cast<uint4*>(dpred)[i8] = make_uint4(out[0], out[1], out[2], out[3]);
    }
  } else {
    for (long i = tid; i < total; i += nthreads) {
      const float p = __bfloat162float(pred[i]);
      const float c = loss_load(clean, in_dtype, i);
      float diff, w = 1.f;
      if (mode == 1) {
        const float d = fmaxf(1.f - __ldg(timestep + i / per_sample), clamp_eps);
        const float z = loss_load(noisy, in_dtype, i);
        diff = (p - z) / d - (c - z) / d;
        w = 1.f / d;
      } else {
        diff = p - c;
      }
      acc += diff * diff;
      if (dpred != nullptr) dpred[i] = __float2bfloat16_rn(2.f * diff * w * inv_n);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += s[i];
    atomicAdd(loss_out, t * inv_n);
  }
}


// ------------------------------------------------------------------------------------------------ noised latents
// prepare_scaled_noised_latents (reference src/modules/loss/flow_match.py:60-74) as ONE pass with the rounding points of its
// five elementwise ops in the tensors' dtype T (bf16 / fp16 / fp32):
//   n = T(randn * noise_scale);  tv = T(timestep[b]);  noisy = T( T(tv * x) + T( T(1 - tv) * n ) )     (clean_at_zero swaps
// the two weights).  Besides `noisy` (T) it writes the bf16 copy the denoiser takes.  __fmul_rn / __fadd_rn: never
// contracted into an FMA, so the fp32 variant rounds like the separate kernels too.
__device__ __forceinline__ float round_as(float x, int dtype) {
  if (dtype == 0) return __bfloat162float(__float2bfloat16_rn(x));
  if (dtype == 1) return __half2float(__float2half_rn(x));
  return x;
}
__device__ __forceinline__ float noised_value(float x, float z, float tv, float omt, float noise_scale, int dtype, int clean_at_zero) {
  const float n = round_as(__fmul_rn(z, noise_scale), dtype);
  const float wx = clean_at_zero ? omt : tv, wn = clean_at_zero ? tv : omt;
  return round_as(__fadd_rn(round_as(__fmul_rn(wx, x), dtype), round_as(__fmul_rn(wn, n), dtype)), dtype);
}
__device__ __forceinline__ void noised_store(void* p, int dtype, long i, float v) {
  if (dtype == 0) static_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else if (dtype == 1) static_cast<__half*>(p)[i] = __float2half_rn(v);
  else static_cast<float*>(p)[i] = v;
}

__global__ void __launch_bounds__(256)
noise_mix_kernel(const void* __restrict__ latents, const void* __restrict__ randn, int dtype, const float* __restrict__ timestep,
                 long per_sample, long total, float noise_scale, int clean_at_zero, void* __restrict__ noisy,
                 __nv_bfloat16* __restrict__ noisy_bf16) {
  const long tid = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x, nthreads = static_cast<long>(gridDim.x) * blockDim.x;
  const bool aligned = ((reinterpret_cast<uintptr_t>(latents) | reinterpret_cast<uintptr_t>(randn) | reinterpret_cast<uintptr_t>(noisy) |
                         reinterpret_cast<uintptr_t>(noisy_bf16)) & 15) == 0;
  if (dtype != 2 && (per_sample & 7) == 0 && aligned) {
    for (long i8 = tid; i8 < (total >> 3); i8 += nthreads) {
      float xf[8], zf[8], o[8];
      loss_load8(latents, dtype, i8, xf);
      loss_load8(randn, dtype, i8, zf);
      const float tv = round_as(__ldg(timestep + (i8 << 3) / per_sample), dtype);
      const float omt = round_as(__fadd_rn(1.f, -tv), dtype);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = noised_value(xf[e], zf[e], tv, omt, noise_scale, dtype, clean_at_zero);
      uint32_t w[4], wb[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162 b2 = __floats2bfloat162_rn(o[2 * k], o[2 * k + 1]);
        wb[k] = *reinterpret_cast<const uint32_t*>(&b2);
        if (dtype == 1) {
          const __half2 h2 = __floats2half2_rn(o[2 * k], o[2 * k + 1]);
          w[k] = *reinterpret_cast<const uint32_t*>(&h2);
        } else {
          w[k] = wb[k];
        }
      }
      reinterpret_cast<uint4*>(noisy)[i8] = make_uint4(w[0], w[1], w[2], w[3]);
      if (noisy_bf16 != nullptr) reinterpret_cast<uint4*>(noisy_bf16)[i8] = make_uint4(wb[0], wb[1], wb[2], wb[3]);
    }
  } else {
    for (long i = tid; i < total; i += nthreads) {
      const float tv = round_as(__ldg(timestep + i / per_sample), dtype);
      const float omt = round_as(__fadd_rn(1.f, -tv), dtype);
      const float v = noised_value(loss_load(latents, dtype, i), loss_load(randn, dtype, i), tv, omt, noise_scale, dtype, clean_at_zero);
      noised_store(noisy, dtype, i, v);
      if (noisy_bf16 != nullptr) noisy_bf16[i] = __float2bfloat16_rn(v);
    }
  }
}

// y = bf16( x * bf16(s[0]) ) for a device scalar s: the upstream gradient of a scalar loss applied to d loss / d pred
// (16-byte accesses; a broadcast multiply by a 0-dim CUDA tensor runs torch's strided kernel at a third of the bandwidth)
__global__ void __launch_bounds__(256)
scale_by_scalar_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ s, __nv_bfloat16* __restrict__ y, long n) {
  const float sc = __bfloat162float(__float2bfloat16_rn(__ldg(s)));
  const long tid = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x, nthreads = static_cast<long>(gridDim.x) * blockDim.x;
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  const long n8 = aligned ? n >> 3 : 0;
  for (long i8 = tid; i8 < n8; i8 += nthreads) {
    float f[8];
    loss_load8(x, 0, i8, f);
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat162 b2 = __floats2bfloat162_rn(__fmul_rn(f[2 * k], sc), __fmul_rn(f[2 * k + 1], sc));
      w[k] = *reinterpret_cast<const uint32_t*>(&b2);
    }
    reinterpret_cast<uint4*>(y)[i8] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  for (long i = (n8 << 3) + tid; i < n; i += nthreads) y[i] = __float2bfloat16_rn(__fmul_rn(__bfloat162float(x[i]), sc));
}

}  // namespace vpt
