// Bring-up probe for the attention kernels: small problem vs a host double-precision reference, then timing at the
// JiT-B / JiT-L bench shapes.
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "attn_launch.cuh"

using namespace vpt;
static float frand() { return (rand() / (float)RAND_MAX) * 2.f - 1.f; }
#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                       \
    }                                                                                \
  } while (0)

static int check(int B, int H, int Lq, int Lk, const std::vector<int>& seqlens) {
  const int D = H * 64;
  const size_t nq = (size_t)B * Lq * D, nk = (size_t)B * Lk * D;
  std::vector<__nv_bfloat16> q(nq), k(nk), v(nk), d_o(nq);
  for (auto& x : q) x = __float2bfloat16_rn(frand() * 2.f);
  for (auto& x : k) x = __float2bfloat16_rn(frand() * 2.f);
  for (auto& x : v) x = __float2bfloat16_rn(frand());
  for (auto& x : d_o) x = __float2bfloat16_rn(frand());
  const float scale = 0.125f;
  __nv_bfloat16 *dq_, *dk_, *dv_, *ddo, *dout, *ddk, *ddv;
  float *dlse, *ddelta, *ddq;
  int* dseq;
  CK(cudaMalloc(&dq_, nq * 2)); CK(cudaMalloc(&dk_, nk * 2)); CK(cudaMalloc(&dv_, nk * 2)); CK(cudaMalloc(&ddo, nq * 2));
  CK(cudaMalloc(&dout, nq * 2)); CK(cudaMalloc(&ddk, nk * 2)); CK(cudaMalloc(&ddv, nk * 2));
  CK(cudaMalloc(&dlse, (size_t)B * H * Lq * 4)); CK(cudaMalloc(&ddelta, (size_t)B * H * Lq * 4)); CK(cudaMalloc(&ddq, nq * 4));
  CK(cudaMalloc(&dseq, B * 4));
  CK(cudaMemcpy(dq_, q.data(), nq * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dk_, k.data(), nk * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dv_, v.data(), nk * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(ddo, d_o.data(), nq * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dseq, seqlens.data(), B * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xff, nq * 2)); CK(cudaMemset(ddk, 0xff, nk * 2)); CK(cudaMemset(ddv, 0xff, nk * 2));
  CK(cudaMemset(ddq, 0, nq * 4));
  AttnTensor tq{dq_, (long)Lq * D, D, 64}, tk{dk_, (long)Lk * D, D, 64}, tv{dv_, (long)Lk * D, D, 64};
  AttnTensor to{dout, (long)Lq * D, D, 64}, tdo{ddo, (long)Lq * D, D, 64}, tdq{ddq, (long)Lq * D, D, 64};
  AttnTensor tdk{ddk, (long)Lk * D, D, 64}, tdv{ddv, (long)Lk * D, D, 64};
  if (launch_attn_fwd(tq, tk, tv, to, B, H, Lq, Lk, dseq, scale, dlse, 0)) { printf("fwd launch: %s\n", last_error().c_str()); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("fwd kernel failed: %s\n", cudaGetErrorString(e)); exit(3); }
  if (launch_attn_bwd(tq, tk, tv, to, tdo, tdq, tdk, tdv, B, H, Lq, Lk, dseq, scale, dlse, ddelta, 0)) { printf("bwd launch: %s\n", last_error().c_str()); return 1; }
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("bwd kernel failed: %s\n", cudaGetErrorString(e)); exit(3); }
  std::vector<__nv_bfloat16> out(nq), gk(nk), gv(nk);
  std::vector<float> gq(nq), lse((size_t)B * H * Lq);
  CK(cudaMemcpy(out.data(), dout, nq * 2, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(gk.data(), ddk, nk * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(gv.data(), ddv, nk * 2, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(gq.data(), ddq, nq * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(lse.data(), dlse, lse.size() * 4, cudaMemcpyDeviceToHost));
  // host reference
  double eo = 0, eq = 0, ek = 0, ev = 0, el = 0, mo = 0, mq = 0, mk = 0, mv = 0;
  std::vector<double> rdk(nk, 0.0), rdv(nk, 0.0);
  for (int b = 0; b < B; ++b)
    for (int h = 0; h < H; ++h) {
      const int kl = seqlens[b];
      std::vector<double> P((size_t)Lq * kl);
      for (int i = 0; i < Lq; ++i) {
        const __nv_bfloat16* qi = &q[((size_t)b * Lq + i) * D + h * 64];
        double mx = -1e30;
        std::vector<double> s(kl);
        for (int j = 0; j < kl; ++j) {
          const __nv_bfloat16* kj = &k[((size_t)b * Lk + j) * D + h * 64];
          double a = 0;
          for (int d = 0; d < 64; ++d) a += (double)__bfloat162float(qi[d]) * __bfloat162float(kj[d]);
          s[j] = a * scale;
          mx = fmax(mx, s[j]);
        }
        double l = 0;
        for (int j = 0; j < kl; ++j) { s[j] = exp(s[j] - mx); l += s[j]; }
        double o[64] = {0};
        for (int j = 0; j < kl; ++j) {
          P[(size_t)i * kl + j] = s[j] / l;
          const __nv_bfloat16* vj = &v[((size_t)b * Lk + j) * D + h * 64];
          for (int d = 0; d < 64; ++d) o[d] += s[j] / l * __bfloat162float(vj[d]);
        }
        const double lse_ref = (mx + log(l)) * 1.4426950408889634;
        el = fmax(el, fabs(lse_ref - lse[((size_t)b * H + h) * Lq + i]));
        const __nv_bfloat16* doi = &d_o[((size_t)b * Lq + i) * D + h * 64];
        double delta = 0;
        for (int d = 0; d < 64; ++d) {
          const double got = __bfloat162float(out[((size_t)b * Lq + i) * D + h * 64 + d]);
          eo = fmax(eo, fabs(got - o[d])); mo = fmax(mo, fabs(o[d]));
          delta += o[d] * __bfloat162float(doi[d]);
        }
        double dqv[64] = {0};
        for (int j = 0; j < kl; ++j) {
          const __nv_bfloat16* vj = &v[((size_t)b * Lk + j) * D + h * 64];
          const __nv_bfloat16* kj = &k[((size_t)b * Lk + j) * D + h * 64];
          double dp = 0;
          for (int d = 0; d < 64; ++d) dp += (double)__bfloat162float(doi[d]) * __bfloat162float(vj[d]);
          const double p = P[(size_t)i * kl + j];
          const double ds = p * (dp - delta) * scale;
          for (int d = 0; d < 64; ++d) {
            dqv[d] += ds * __bfloat162float(kj[d]);
            rdk[((size_t)b * Lk + j) * D + h * 64 + d] += ds * __bfloat162float(qi[d]);
            rdv[((size_t)b * Lk + j) * D + h * 64 + d] += p * __bfloat162float(doi[d]);
          }
        }
        for (int d = 0; d < 64; ++d) {
          eq = fmax(eq, fabs(gq[((size_t)b * Lq + i) * D + h * 64 + d] - dqv[d])); mq = fmax(mq, fabs(dqv[d]));
        }
      }
    }
  for (size_t i = 0; i < nk; ++i) {
    ek = fmax(ek, fabs(__bfloat162float(gk[i]) - rdk[i])); mk = fmax(mk, fabs(rdk[i]));
    ev = fmax(ev, fabs(__bfloat162float(gv[i]) - rdv[i])); mv = fmax(mv, fabs(rdv[i]));
  }
  const bool ok = eo < 0.02 * mo + 1e-3 && eq < 0.03 * mq + 1e-3 && ek < 0.03 * mk + 1e-3 && ev < 0.03 * mv + 1e-3 && el < 0.01;
  printf("B=%d H=%d Lq=%d Lk=%d : %s  O err %.4f/%.2f  lse err %.5f  dQ %.4f/%.2f  dK %.4f/%.2f  dV %.4f/%.2f\n", B, H, Lq, Lk,
         ok ? "OK  " : "FAIL", eo, mo, el, eq, mq, ek, mk, ev, mv);
  cudaFree(dq_); cudaFree(dk_); cudaFree(dv_); cudaFree(ddo); cudaFree(dout); cudaFree(ddk); cudaFree(ddv);
  cudaFree(dlse); cudaFree(ddelta); cudaFree(ddq); cudaFree(dseq);
  return ok ? 0 : 1;
}

static void bench(int B, int H, int L) {
  const int D = H * 64;
  const size_t n = (size_t)B * L * D;
  __nv_bfloat16 *q, *k, *v, *o, *d_o, *dk, *dv;
  float *lse, *delta, *dq;
  CK(cudaMalloc(&q, n * 2)); CK(cudaMalloc(&k, n * 2)); CK(cudaMalloc(&v, n * 2)); CK(cudaMalloc(&o, n * 2));
  CK(cudaMalloc(&d_o, n * 2)); CK(cudaMalloc(&dk, n * 2)); CK(cudaMalloc(&dv, n * 2)); CK(cudaMalloc(&dq, n * 4));
  CK(cudaMalloc(&lse, (size_t)B * H * L * 4)); CK(cudaMalloc(&delta, (size_t)B * H * L * 4));
  std::vector<__nv_bfloat16> h(n);
  for (auto& x : h) x = __float2bfloat16_rn(frand());
  for (auto p : {q, k, v, d_o}) CK(cudaMemcpy(p, h.data(), n * 2, cudaMemcpyHostToDevice));
  AttnTensor tq{q, (long)L * D, D, 64}, tk{k, (long)L * D, D, 64}, tv{v, (long)L * D, D, 64}, to{o, (long)L * D, D, 64};
  AttnTensor tdo{d_o, (long)L * D, D, 64}, tdq{dq, (long)L * D, D, 64}, tdk{dk, (long)L * D, D, 64}, tdv{dv, (long)L * D, D, 64};
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int iters = 10;
  for (int i = 0; i < 3; ++i) launch_attn_fwd(tq, tk, tv, to, B, H, L, L, nullptr, 0.125f, lse, 0);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) launch_attn_fwd(tq, tk, tv, to, B, H, L, L, nullptr, 0.125f, lse, 0);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  const double fl = 4.0 * B * H * (double)L * L * 64;
  printf("B=%d H=%d L=%d fwd: %.1f us  %.1f TF\n", B, H, L, ms * 1000 / iters, fl / (ms * 1e-3 / iters) * 1e-12);
  for (int i = 0; i < 3; ++i) { cudaMemsetAsync(dq, 0, n * 4); launch_attn_bwd(tq, tk, tv, to, tdo, tdq, tdk, tdv, B, H, L, L, nullptr, 0.125f, lse, delta, 0); }
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) { cudaMemsetAsync(dq, 0, n * 4); launch_attn_bwd(tq, tk, tv, to, tdo, tdq, tdk, tdv, B, H, L, L, nullptr, 0.125f, lse, delta, 0); }
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("B=%d H=%d L=%d bwd (memset + delta + main): %.1f us  %.1f TF\n", B, H, L, ms * 1000 / iters, 2.5 * fl / (ms * 1e-3 / iters) * 1e-12);
  cudaFree(q); cudaFree(k); cudaFree(v); cudaFree(o); cudaFree(d_o); cudaFree(dk); cudaFree(dv); cudaFree(dq); cudaFree(lse); cudaFree(delta);
}

int main(int argc, char** argv) {
  srand(7);
  int fails = 0;
  fails += check(2, 2, 330, 330, {330, 281});
  fails += check(1, 3, 128, 128, {128});
  fails += check(2, 1, 200, 77, {77, 50});
  fails += check(1, 2, 520, 520, {515});
  printf(fails == 0 ? "ATTN PROBE PASS\n" : "ATTN PROBE FAIL (%d)\n", fails);
  if (argc > 1) {
    bench(64, 12, 330);
    bench(64, 16, 330);
    bench(16, 16, 1100);
  }
  return fails;
}
