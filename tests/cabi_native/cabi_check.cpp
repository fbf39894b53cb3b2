// cabi_check.cpp -- TEST PROGRAM: drives libspinn_b200.so through include/spinn_b200.h from plain C++
// (CUDA runtime only, no Python, no torch), the way a non-Python host would bind it, and checks the
// results against straightforward double-precision loops written here (Gaussian kernel closed forms,
// SURVEY.md 7.1).  Built by __graft_entry__.build(); run by tests/test_gpu_cabi_native.py.
//
//   cabi_check [N M_free M_fixed]      exit code 0 = all checks passed
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/spinn_b200.h"

#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      std::fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 2;                                                                          \
    }                                                                                    \
  } while (0)

static unsigned long long rng_state = 0x9E3779B97F4A7C15ull;
static double urand() {  // xorshift64*, uniform in [0, 1)
  rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27;
  return (double)((rng_state * 2685821657736338717ull) >> 11) / 9007199254740992.0;
}
static double nrand() { return std::sqrt(-2.0 * std::log(1.0 - urand())) * std::cos(6.283185307179586 * urand()); }

template <class T>
static T* to_device(const std::vector<T>& v) {
  T* p = nullptr;
  if (cudaMalloc(&p, v.size() * sizeof(T) + 16) != cudaSuccess) return nullptr;
  cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  return p;
}

// Two metrics against the double-precision loop: max-norm (|a-b|_max / |b|_max) and the parity
// metric of tests/helpers.py, max |a-b| / (|b| + 1e-2 |b|_max).  Random clouds with random-sign
// weights put many elements near zero crossings, where a few ulp of the largest TERMS of the fp32
// sum show up as a few 1e-5 in the second metric (for any fp32 evaluation, the reference's included).
struct Err { double maxnorm, mixed; };
static Err errors(const std::vector<float>& a, const std::vector<double>& b) {
  double scale = 0.0;
  Err e = {0.0, 0.0};
  for (double v : b) scale = std::fmax(scale, std::fabs(v));
  for (size_t i = 0; i < b.size(); ++i) {
    const double d = std::fabs((double)a[i] - b[i]);
    e.maxnorm = std::fmax(e.maxnorm, d / (scale + 1e-300));
    e.mixed = std::fmax(e.mixed, d / (std::fabs(b[i]) + 1e-2 * scale + 1e-300));
  }
  return e;
}
static int report(const char* name, const std::vector<float>& a, const std::vector<double>& b) {
  const Err e = errors(a, b);
  std::printf("%-5s maxnorm %.2e  mixed %.2e\n", name, e.maxnorm, e.mixed);
  return (e.maxnorm > 5e-6) + (e.mixed > 1e-4);
}

int main(int argc, char** argv) {
  const int N = argc > 3 ? std::atoi(argv[1]) : 1001;
  const int Mf = argc > 3 ? std::atoi(argv[2]) : 50;
  const int Mx = argc > 3 ? std::atoi(argv[3]) : 27;
  const int M = Mf + Mx, K = 1, NJ = 5;
  if (spinn_abi_version() != SPINN_ABI_VERSION) { std::fprintf(stderr, "ABI version mismatch\n"); return 1; }
  if (spinn_num_jets(2, 2) != NJ) { std::fprintf(stderr, "spinn_num_jets\n"); return 1; }

  std::vector<float> x(N), y(N), xc(M), yc(M), h(M), W(M), bias(1), g((size_t)NJ * N);
  for (int i = 0; i < N; ++i) { x[i] = (float)urand(); y[i] = (float)urand(); }
  for (int j = 0; j < M; ++j) {
    xc[j] = (float)urand(); yc[j] = (float)urand();
    h[j] = (float)((1.0 + 0.5 * urand()) / std::sqrt((double)M));
    W[j] = (float)(0.3 * nrand());
  }
  bias[0] = 0.1f;
  for (auto& v : g) v = (float)nrand();

  // ---- double-precision evaluation of the same definition ----
  std::vector<double> J((size_t)NJ * N, 0.0), dxc(M, 0.0), dyc(M, 0.0), dh(M, 0.0), dW(M, 0.0), db(1, 0.0);
  for (int i = 0; i < N; ++i) {
    for (int j = 0; j < M; ++j) {
      const double r = 1.0 / (double)h[j], rho = r * r;
      const double X = (double)x[i] - (double)xc[j], Y = (double)y[i] - (double)yc[j];
      const double a = rho * X, b = rho * Y;
      const double p = std::exp(-0.5 * rho * (X * X + Y * Y));
      const double d0 = p, d1 = -a * p, d2 = -b * p, d3 = (a * a - rho) * p, d4 = (b * b - rho) * p, d5 = a * b * p;
      const double d6 = a * (3.0 * rho - a * a) * p, d9 = b * (3.0 * rho - b * b) * p;   // xxx, yyy
      const double d7 = -a * (b * b - rho) * p, d8 = -(a * a - rho) * b * p;             // xyy, xxy
      const double w = W[j];
      J[0 * (size_t)N + i] += w * d0; J[1 * (size_t)N + i] += w * d1; J[2 * (size_t)N + i] += w * d2;
      J[3 * (size_t)N + i] += w * d3; J[4 * (size_t)N + i] += w * d4;
      const double g0 = g[i], g1 = g[(size_t)N + i], g2 = g[2 * (size_t)N + i], g3 = g[3 * (size_t)N + i],
                   g4 = g[4 * (size_t)N + i];
      dW[j] += g0 * d0 + g1 * d1 + g2 * d2 + g3 * d3 + g4 * d4;
      // d/dc = -d/dx of every jet; d/dh by Euler scaling: -(1/h)(|alpha| phi_alpha + X phi_alpha,x + Y phi_alpha,y)
      const double cx = g0 * d1 + g1 * d3 + g2 * d5 + g3 * d6 + g4 * d7;
      const double cy = g0 * d2 + g1 * d5 + g2 * d4 + g3 * d8 + g4 * d9;
      const double ord = g1 * d1 + g2 * d2 + 2.0 * (g3 * d3 + g4 * d4);
      dxc[j] += -w * cx; dyc[j] += -w * cy;
      dh[j] += -w * r * (ord + X * cx + Y * cy);
    }
    J[i] += bias[0];
    db[0] += g[i];
  }

  // ---- the library ----
  float *dx = to_device(x), *dy = to_device(y), *dxcf = to_device(xc), *dycf = to_device(yc), *dhh = to_device(h),
        *dWw = to_device(W), *dbias = to_device(bias), *dg = to_device(g);
  if (!dx || !dy || !dxcf || !dycf || !dhh || !dWw || !dbias || !dg) { std::fprintf(stderr, "cudaMalloc failed\n"); return 2; }
  float *djets, *o_xc, *o_yc, *o_h, *o_W, *o_b;
  CK(cudaMalloc(&djets, sizeof(float) * NJ * (size_t)N));
  CK(cudaMalloc(&o_xc, sizeof(float) * (Mf + 1))); CK(cudaMalloc(&o_yc, sizeof(float) * (Mf + 1)));
  CK(cudaMalloc(&o_h, sizeof(float) * M)); CK(cudaMalloc(&o_W, sizeof(float) * M)); CK(cudaMalloc(&o_b, sizeof(float)));
  const size_t wf = spinn2d_fwd_workspace_bytes(N, M, K), wb = spinn2d_bwd_workspace_bytes(N, M, K);
  void *wsf, *wsb;
  CK(cudaMalloc(&wsf, wf + 16)); CK(cudaMalloc(&wsb, wb + 16));
  cudaStream_t s;
  CK(cudaStreamCreate(&s));
  // free nodes = the first Mf entries, fixed nodes = the rest of the same arrays
  int rc = spinn2d_jets_fwd(SPINN_KIND_GAUSSIAN, 2, N, Mf, Mx, K, dx, dy, dxcf, dycf, dxcf + Mf, dycf + Mf, dhh, 0,
                            dWw, dbias, djets, wsf, wf, s);
  if (rc != SPINN_OK) { std::fprintf(stderr, "fwd rc=%d: %s\n", rc, spinn_last_error()); return 1; }
  rc = spinn2d_jets_bwd(SPINN_KIND_GAUSSIAN, 2, 0x1F, N, Mf, Mx, K, dx, dy, dxcf, dycf, dxcf + Mf, dycf + Mf, dhh, 0,
                        dWw, dg, o_xc, o_yc, o_h, o_W, o_b, wsb, wb, s);
  if (rc != SPINN_OK) { std::fprintf(stderr, "bwd rc=%d: %s\n", rc, spinn_last_error()); return 1; }
  CK(cudaStreamSynchronize(s));
  std::vector<float> jets((size_t)NJ * N), gxc(Mf), gyc(Mf), gh(M), gW(M), gb(1);
  CK(cudaMemcpy(jets.data(), djets, sizeof(float) * jets.size(), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(gxc.data(), o_xc, sizeof(float) * Mf, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(gyc.data(), o_yc, sizeof(float) * Mf, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(gh.data(), o_h, sizeof(float) * M, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(gW.data(), o_W, sizeof(float) * M, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(gb.data(), o_b, sizeof(float), cudaMemcpyDeviceToHost));

  int bad = 0;
  const char* jn[NJ] = {"u", "u_x", "u_y", "u_xx", "u_yy"};
  for (int a = 0; a < NJ; ++a) {
    std::vector<float> got(jets.begin() + (size_t)a * N, jets.begin() + (size_t)(a + 1) * N);
    std::vector<double> want(J.begin() + (size_t)a * N, J.begin() + (size_t)(a + 1) * N);
    bad += report(jn[a], got, want);
  }
  dxc.resize(Mf); dyc.resize(Mf);
  struct { const char* n; const std::vector<float>* g; const std::vector<double>* w; } rows[] = {
      {"d_xc", &gxc, &dxc}, {"d_yc", &gyc, &dyc}, {"d_h", &gh, &dh}, {"d_W", &gW, &dW}, {"d_b", &gb, &db}};
  for (auto& r : rows) bad += report(r.n, *r.g, *r.w);
  // ---- error behaviour: bad arguments return a negative code and a message, nothing is launched ----
  rc = spinn2d_jets_fwd(SPINN_KIND_GAUSSIAN, 2, -1, Mf, Mx, K, dx, dy, dxcf, dycf, dxcf + Mf, dycf + Mf, dhh, 0, dWw,
                        dbias, djets, wsf, wf, s);
  if (rc >= 0 || !spinn_last_error()[0]) { std::printf("negative N accepted\n"); ++bad; }
  rc = spinn2d_jets_bwd(SPINN_KIND_GAUSSIAN, 2, 0x1F, N, Mf, Mx, K, dx, dy, dxcf, dycf, dxcf + Mf, dycf + Mf, dhh, 0, dWw,
                        dg, o_xc, o_yc, o_h, o_W, o_b, wsb, 16, s);
  if (rc != SPINN_E_WORKSPACE) { std::printf("short workspace accepted (rc=%d)\n", rc); ++bad; }
  std::printf(bad ? "FAILED (%d)\n" : "cabi_check OK\n", bad);
  return bad ? 1 : 0;
}
