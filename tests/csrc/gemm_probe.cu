// Bring-up probe for gemm_nf4lora_kernel: runs every (direction x weight source x LoRA x tile width) variant on
// small ragged shapes and checks against a host fp32 computation.  Standalone binary (no torch):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I vision_pt_b200/csrc tests/csrc/gemm_probe.cu -o ...
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "gemm_launch.cuh"

using namespace vpt;

static float bf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
static float frand() { return (rand() / (float)RAND_MAX) * 2.f - 1.f; }

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e = (x);                                                               \
    if (e != cudaSuccess) {                                                            \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);   \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

template <class T>
T* upload(const std::vector<T>& h) {
  T* d;
  CK(cudaMalloc(&d, h.size() * sizeof(T) + 64));
  CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

struct Problem {
  int M, N, K;
  std::vector<__nv_bfloat16> X, dY, Wdq, down, up, bias;
  std::vector<uint8_t> packed, qabs;
  std::vector<float> nested, ncode, code;
  float offset, scale;
  std::vector<float> Yref, Tsref, dXref, dTsref;
};

static const float kNF4[16] = {-1.0f, -0.6961928009986877f, -0.5250730514526367f, -0.39491748809814453f,
                               -0.28444138169288635f, -0.18477343022823334f, -0.09105003625154495f, 0.0f,
                               0.07958029955625534f, 0.16093020141124725f, 0.24611230194568634f, 0.33791524171829224f,
                               0.44070982933044434f, 0.5626170039176941f, 0.7229568362236023f, 1.0f};

static void build(Problem& P, int M, int N, int K) {
  P.M = M; P.N = N; P.K = K;
  P.scale = 0.25f;
  const size_t nw = (size_t)N * K;
  std::vector<float> W(nw);
  for (auto& w : W) w = frand() * 0.05f;
  // blockwise absmax, then an 8-bit code for (absmax - mean) with a linear 256-entry table (any table works: the
  // kernel takes the tables as data)
  const size_t nb = nw / 64;
  std::vector<float> am(nb);
  double mean = 0;
  for (size_t b = 0; b < nb; ++b) {
    float m = 0;
    for (int i = 0; i < 64; ++i) m = fmaxf(m, fabsf(W[b * 64 + i]));
    am[b] = m;
    mean += m;
  }
  P.offset = (float)(mean / nb);
  P.ncode.resize(256);
  for (int i = 0; i < 256; ++i) P.ncode[i] = (i - 127.5f) / 127.5f;
  const size_t nnb = (nb + 255) / 256;
  P.nested.resize(nnb);
  P.qabs.resize(nb);
  std::vector<float> amdq(nb);
  for (size_t g = 0; g < nnb; ++g) {
    float m = 0;
    for (size_t b = g * 256; b < nb && b < (g + 1) * 256; ++b) m = fmaxf(m, fabsf(am[b] - P.offset));
    P.nested[g] = m > 0 ? m : 1.f;
    for (size_t b = g * 256; b < nb && b < (g + 1) * 256; ++b) {
      float x = (am[b] - P.offset) / P.nested[g];
      int q = (int)lrintf(x * 127.5f + 127.5f);
      q = q < 0 ? 0 : (q > 255 ? 255 : q);
      P.qabs[b] = (uint8_t)q;
      volatile float t = P.ncode[q] * P.nested[g];
      volatile float u = t + P.offset;
      amdq[b] = u;
    }
  }
  P.code.assign(kNF4, kNF4 + 16);
  P.packed.resize(nw / 2);
  P.Wdq.resize(nw);
  for (size_t i = 0; i < nw; ++i) {
    float x = W[i] / amdq[i / 64];
    int best = 0;
    float bd = 1e9f;
    for (int c = 0; c < 16; ++c) {
      float d = fabsf(x - kNF4[c]);
      if (d < bd) { bd = d; best = c; }
    }
    if (i & 1) P.packed[i / 2] |= (uint8_t)best; else P.packed[i / 2] = (uint8_t)(best << 4);
    volatile float prod = kNF4[best] * amdq[i / 64];
    P.Wdq[i] = __float2bfloat16_rn(prod);
  }
  P.X.resize((size_t)M * K);
  for (auto& x : P.X) x = __float2bfloat16_rn(frand());
  P.dY.resize((size_t)M * N);
  for (auto& x : P.dY) x = __float2bfloat16_rn(frand());
  P.down.resize((size_t)16 * K);
  for (auto& x : P.down) x = __float2bfloat16_rn(frand() * 0.1f);
  P.up.resize((size_t)N * 16);
  for (auto& x : P.up) x = __float2bfloat16_rn(frand() * 0.1f);
  P.bias.resize(N);
  for (auto& x : P.bias) x = __float2bfloat16_rn(frand());

  std::vector<float> Xf(P.X.size()), dYf(P.dY.size()), Wf(nw), dn(P.down.size()), upf(P.up.size());
  for (size_t i = 0; i < Xf.size(); ++i) Xf[i] = __bfloat162float(P.X[i]);
  for (size_t i = 0; i < dYf.size(); ++i) dYf[i] = __bfloat162float(P.dY[i]);
  for (size_t i = 0; i < nw; ++i) Wf[i] = __bfloat162float(P.Wdq[i]);
  for (size_t i = 0; i < dn.size(); ++i) dn[i] = __bfloat162float(P.down[i]);
  for (size_t i = 0; i < upf.size(); ++i) upf[i] = __bfloat162float(P.up[i]);

  P.Tsref.assign((size_t)M * 16, 0);
  P.dTsref.assign((size_t)M * 16, 0);
  for (int m = 0; m < M; ++m)
    for (int j = 0; j < 16; ++j) {
      double a = 0, b = 0;
      for (int k = 0; k < K; ++k) a += (double)Xf[(size_t)m * K + k] * dn[(size_t)j * K + k];
      for (int n = 0; n < N; ++n) b += (double)dYf[(size_t)m * N + n] * upf[(size_t)n * 16 + j];
      P.Tsref[(size_t)m * 16 + j] = bf((float)a * P.scale);
      P.dTsref[(size_t)m * 16 + j] = bf((float)b * P.scale);
    }
  // base parts
  P.Yref.assign((size_t)M * N, 0);
  P.dXref.assign((size_t)M * K, 0);
  for (int m = 0; m < M; ++m) {
    for (int n = 0; n < N; ++n) {
      double a = 0;
      for (int k = 0; k < K; ++k) a += (double)Xf[(size_t)m * K + k] * Wf[(size_t)n * K + k];
      P.Yref[(size_t)m * N + n] = (float)a;
    }
    for (int k = 0; k < K; ++k) {
      double a = 0;
      for (int n = 0; n < N; ++n) a += (double)dYf[(size_t)m * N + n] * Wf[(size_t)n * K + k];
      P.dXref[(size_t)m * K + k] = (float)a;
    }
  }
}

struct Dev {
  __nv_bfloat16 *X, *dY, *Wdq, *down, *up, *bias, *out, *side;
  uint8_t *packed, *qabs;
  float *nested, *ncode, *code;
};

static bool run_case(const Problem& P, const Dev& d, bool bwd, bool nf4, bool lora, int bn, int max_ctas,
                     const uint32_t* dbg /*6 or null*/, bool verbose) {
  const int M = P.M, N = P.N, K = P.K;
  const int NO = bwd ? K : N, R = bwd ? N : K;
  GemmLaunch g{};
  g.bwd = bwd; g.nf4 = nf4; g.lora = lora; g.bn = bn;
  g.act = bwd ? d.dY : d.X;
  g.lda = R;
  g.w_bf16 = d.Wdq;
  g.max_ctas = max_ctas;
  g.p.M = M; g.p.NO = NO; g.p.R = R;
  g.p.D = d.out; g.p.ldd = NO;
  g.p.bias = bwd ? nullptr : d.bias;
  g.p.w = Nf4Weight{d.packed, d.qabs, d.nested, d.ncode, d.code, P.offset, N, K, nullptr, nullptr, 0};
  g.p.ld_down = K;
  g.p.lora_down = d.down; g.p.lora_up = d.up; g.p.scale = P.scale; g.p.side = d.side;
  if (dbg) {
    g.p.dbg_b_lbo = dbg[0]; g.p.dbg_b_sbo = dbg[1]; g.p.dbg_q_lbo = dbg[2];
    g.p.dbg_q_sbo = dbg[3]; g.p.dbg_ts_lbo = dbg[4]; g.p.dbg_ts_sbo = dbg[5];
  }
  CK(cudaMemset(d.out, 0xff, (size_t)M * NO * 2));
  CK(cudaMemset(d.side, 0xff, (size_t)M * 16 * 2));
  if (launch_gemm(g, 0)) {
    printf("  launch failed: %s\n", last_error().c_str());
    return false;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("  kernel failed: %s\n", cudaGetErrorString(e));
    exit(3);   // context is gone after a trap
  }
  std::vector<__nv_bfloat16> out((size_t)M * NO), side((size_t)M * 16);
  CK(cudaMemcpy(out.data(), d.out, out.size() * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(side.data(), d.side, side.size() * 2, cudaMemcpyDeviceToHost));
  // reference with the LoRA term built from the reference side tensor
  const std::vector<float>& base = bwd ? P.dXref : P.Yref;
  const std::vector<float>& sref = bwd ? P.dTsref : P.Tsref;
  double max_err = 0, max_ref = 0, side_err = 0;
  long bad = 0;
  for (int m = 0; m < M; ++m)
    for (int o = 0; o < NO; ++o) {
      double r = base[(size_t)m * NO + o];
      if (!bwd) r += __bfloat162float(P.bias[o]);
      if (lora) {
        for (int j = 0; j < 16; ++j) {
          const float q = bwd ? __bfloat162float(P.down[(size_t)j * K + o]) : __bfloat162float(P.up[(size_t)o * 16 + j]);
          r += (double)sref[(size_t)m * 16 + j] * q;
        }
      }
      const double got = __bfloat162float(out[(size_t)m * NO + o]);
      const double err = fabs(got - r);
      if (!(err <= 0.02 * fabs(r) + 0.02)) {
        if (verbose && bad < 6) printf("    mismatch m=%d o=%d got=%f ref=%f\n", m, o, got, r);
        ++bad;
      }
      if (err > max_err || err != err) max_err = err;
      if (fabs(r) > max_ref) max_ref = fabs(r);
    }
  long side_bad = 0;
  if (lora)
    for (size_t i = 0; i < side.size(); ++i) {
      const double got = __bfloat162float(side[i]), r = sref[i];
      const double err = fabs(got - r);
      if (!(err <= 0.02 * fabs(r) + 0.01)) ++side_bad;
      if (err > side_err) side_err = err;
    }
  const bool ok = bad == 0 && side_bad == 0;
  printf("  %s %s %s BN=%d ctas=%d : %s  max_err=%.4f (max_ref=%.2f) bad=%ld side_err=%.4f side_bad=%ld\n",
         bwd ? "bwd" : "fwd", nf4 ? "nf4 " : "bf16", lora ? "lora" : "----", bn, max_ctas, ok ? "OK  " : "FAIL",
         max_err, max_ref, bad, side_err, side_bad);
  return ok;
}


// ---------------------------------------------------------------------------------------------- timing mode
static void fill_random(void* d, size_t bytes) {
  std::vector<uint32_t> h(bytes / 4 + 1);
  for (auto& x : h) x = (uint32_t)rand() * 2654435761u + (uint32_t)rand();
  CK(cudaMemcpy(d, h.data(), bytes, cudaMemcpyHostToDevice));
}
static void fill_bf16(__nv_bfloat16* d, size_t n, float amp) {
  std::vector<__nv_bfloat16> h(n);
  for (auto& x : h) x = __float2bfloat16_rn(frand() * amp);
  CK(cudaMemcpy(d, h.data(), n * 2, cudaMemcpyHostToDevice));
}
static int bench_main(int only_k, int only_n, int only_bwd, int only_nf4, int only_lora, int only_bn, int iters_arg) {
  const int M = 21120;
  const int shapes[][2] = {{768, 768}, {768, 2048}, {2048, 768}, {1024, 1024}, {1024, 2752}, {2752, 1024}};  // {K, N}
  const int maxd = 2752;
  __nv_bfloat16 *X, *dY, *W, *down, *up, *bias, *out, *side;
  uint8_t *packed, *qabs;
  float *nested, *ncode, *code;
  CK(cudaMalloc(&X, (size_t)M * maxd * 2)); CK(cudaMalloc(&dY, (size_t)M * maxd * 2));
  CK(cudaMalloc(&out, (size_t)M * maxd * 2)); CK(cudaMalloc(&side, (size_t)M * 16 * 2));
  CK(cudaMalloc(&W, (size_t)maxd * maxd * 2)); CK(cudaMalloc(&down, 16 * maxd * 2)); CK(cudaMalloc(&up, 16 * maxd * 2));
  CK(cudaMalloc(&bias, maxd * 2)); CK(cudaMalloc(&packed, (size_t)maxd * maxd / 2)); CK(cudaMalloc(&qabs, (size_t)maxd * maxd / 64));
  CK(cudaMalloc(&nested, 4 * ((size_t)maxd * maxd / 64 / 256 + 1))); CK(cudaMalloc(&ncode, 1024)); CK(cudaMalloc(&code, 64));
  fill_bf16(X, (size_t)M * maxd, 1.f); fill_bf16(dY, (size_t)M * maxd, 1.f); fill_bf16(W, (size_t)maxd * maxd, 0.05f);
  fill_bf16(down, 16 * maxd, 0.1f); fill_bf16(up, 16 * maxd, 0.1f); fill_bf16(bias, maxd, 1.f);
  fill_random(packed, (size_t)maxd * maxd / 2); fill_random(qabs, (size_t)maxd * maxd / 64);
  {
    std::vector<float> h((size_t)maxd * maxd / 64 / 256 + 1, 0.01f), nc(256), c(kNF4, kNF4 + 16);
    for (int i = 0; i < 256; ++i) nc[i] = (i - 127.5f) / 127.5f;
    CK(cudaMemcpy(nested, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ncode, nc.data(), 1024, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(code, c.data(), 64, cudaMemcpyHostToDevice));
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  printf("M=%d; time per launch (us) and TFLOP/s (2*M*K*N + LoRA 2*M*16*(K+N))\n", M);
  for (auto& sh : shapes) {
    const int K = sh[0], N = sh[1];
    if (only_k > 0 && (K != only_k || N != only_n)) continue;
    for (int bwd = 0; bwd < 2; ++bwd)
      for (int nf4 = 0; nf4 < 2; ++nf4)
        for (int lora = 0; lora < 2; ++lora)
          for (int bn : {128, 192}) {
            if (only_k > 0 && (bwd != only_bwd || nf4 != only_nf4 || lora != only_lora || bn != only_bn)) continue;
            const int NO = bwd ? K : N, R = bwd ? N : K;
            GemmLaunch g{};
            g.bwd = bwd; g.nf4 = nf4; g.lora = lora; g.bn = bn;
            g.act = bwd ? dY : X; g.lda = R; g.w_bf16 = W;
            g.p.M = M; g.p.NO = NO; g.p.R = R; g.p.D = out; g.p.ldd = NO; g.p.bias = bwd ? nullptr : bias;
            g.p.w = Nf4Weight{packed, qabs, nested, ncode, code, 0.02f, N, K, nullptr, nullptr, 0};
            g.p.ld_down = K;
            g.p.lora_down = down; g.p.lora_up = up; g.p.scale = 0.0625f; g.p.side = side;
            for (int i = 0; i < 3; ++i) if (launch_gemm(g, 0)) { printf("launch failed %s\n", last_error().c_str()); return 1; }
            CK(cudaDeviceSynchronize());
            const int iters = iters_arg;
            CK(cudaEventRecord(e0));
            for (int i = 0; i < iters; ++i) launch_gemm(g, 0);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            const double us = ms * 1000.0 / iters;
            const double fl = 2.0 * M * K * N + (lora ? 2.0 * M * 16 * (K + N) : 0.0);
            printf("  K=%4d N=%4d %s %s %s BN=%d : %8.1f us  %7.1f TF\n", K, N, bwd ? "bwd" : "fwd", nf4 ? "nf4 " : "bf16",
                   lora ? "lora" : "----", bn, us, fl / us * 1e-6);
          }
  }
  return 0;
}

int main(int argc, char** argv) {
  srand(1234);
  if (argc > 1 && strcmp(argv[1], "bench") == 0) return bench_main(0, 0, 0, 0, 0, 0, 20);
  if (argc > 8 && strcmp(argv[1], "one") == 0)
    return bench_main(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7]), atoi(argv[8]));
  int dev_count = 0;
  CK(cudaGetDeviceCount(&dev_count));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);

  Problem P;
  build(P, 300, 448, 320);   // ragged M (3 tiles, last partial), N and K multiples of 64
  Dev d;
  d.X = upload(P.X); d.dY = upload(P.dY); d.Wdq = upload(P.Wdq); d.down = upload(P.down); d.up = upload(P.up);
  d.bias = upload(P.bias); d.packed = upload(P.packed); d.qabs = upload(P.qabs); d.nested = upload(P.nested);
  d.ncode = upload(P.ncode); d.code = upload(P.code);
  CK(cudaMalloc(&d.out, (size_t)P.M * 512 * 2));
  CK(cudaMalloc(&d.side, (size_t)P.M * 16 * 2));

  int fails = 0;
  const uint32_t swap_all[6] = {1024, 8192, 256, 128, 256, 128};
  for (int bwd = 0; bwd < 2; ++bwd)
    for (int lora = 0; lora < 2; ++lora)
      for (int nf4 = 0; nf4 < 2; ++nf4)
        for (int bn : {128, 192}) {
          bool ok = run_case(P, d, bwd, nf4, lora, bn, 0, nullptr, true);
          if (!ok) {
            ++fails;
            printf("   retry with swapped LBO/SBO:\n");
            run_case(P, d, bwd, nf4, lora, bn, 0, swap_all, false);
            const uint32_t swap_b[6] = {1024, 8192, 0, 0, 0, 0};
            const uint32_t swap_q[6] = {0, 0, 256, 128, 0, 0};
            const uint32_t swap_t[6] = {0, 0, 0, 0, 256, 128};
            const uint32_t swap_qt[6] = {0, 0, 256, 128, 256, 128};
            if (bwd) run_case(P, d, bwd, nf4, lora, bn, 0, swap_b, false);
            if (lora) {
              run_case(P, d, bwd, nf4, lora, bn, 0, swap_q, false);
              run_case(P, d, bwd, nf4, lora, bn, 0, swap_t, false);
              run_case(P, d, bwd, nf4, lora, bn, 0, swap_qt, false);
            }
          }
        }
  // persistence: few CTAs looping over many tiles (exercises the accumulator double buffering and phase logic)
  printf("persistent (2 CTAs):\n");
  for (int bwd = 0; bwd < 2; ++bwd)
    if (!run_case(P, d, bwd, true, true, 128, 2, nullptr, true)) ++fails;
  printf(fails == 0 ? "PROBE PASS\n" : "PROBE FAIL (%d)\n", fails);
  return fails == 0 ? 0 : 1;
}
