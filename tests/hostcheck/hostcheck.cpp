// hostcheck.cpp -- TEST INFRASTRUCTURE ONLY (never loaded by the spinn_b200 package).
// Runs the per-(point,node) formulas of spinn_b200/csrc/spinn_math.cuh on the CPU,
// with the same loop structure as the CUDA kernels (node pairs / point pairs in the
// two lanes of a float2, two-level accumulation), so that the algebra can be checked
// against the oracle in the build container, which has no GPU.
#include "../../spinn_b200/csrc/spinn_math.cuh"
#include <vector>

using namespace spinn;

template <int MASK>
static void hc_masked_impl(int N, int M, const float* x, const float* y, const float* g, const float* c,
                           const float* d, const float* h, const float* W, float* dc, float* dd, float* dh,
                           float* dW) {
  int N2 = (N + 1) / 2;
  for (int j = 0; j < M; ++j) {
    G2BwdNode n = g2_bwd_node(c[j], d[j], h[j]);
    float2 acc[4];
    for (int a = 0; a < 4; ++a) acc[a] = f2(0.f, 0.f);
    for (int p = 0; p < N2; ++p) {
      int i0 = 2 * p, i1 = 2 * p + 1;
      bool v = i1 < N;
      float2 gg[5];
      for (int a = 0; a < 5; ++a) gg[a] = f2(g[(size_t)a * N + i0], v ? g[(size_t)a * N + i1] : 0.f);
      g2_bwd_pair_mask<MASK>(f2(x[i0], v ? x[i1] : x[i0]), f2(y[i0], v ? y[i1] : y[i0]), gg[0], gg[1], gg[2],
                             gg[3], gg[4], f2dup(n.nc), f2dup(n.nd), f2dup(n.rp), f2dup(n.nrt), f2dup(n.rho),
                             f2dup(n.nk1), f2dup(n.k2p), acc);
    }
    g2_bwd_finish(h[j], W[j], acc[0].x + acc[0].y, acc[1].x + acc[1].y, acc[2].x + acc[2].y,
                  acc[3].x + acc[3].y, &dW[j], &dc[j], &dd[j], &dh[j]);
  }
}

extern "C" {

// fast forward emulation: jets [6][N]
void hc_g2_fwd(int N, int M, const float* x, const float* y, const float* c, const float* d,
               const float* h, const float* W, float bias, float* jets) {
  int npairs = (M + 1) / 2;
  std::vector<G2FwdNode> nodes(2 * npairs);
  for (int j = 0; j < 2 * npairs; ++j)
    nodes[j] = j < M ? g2_fwd_node(c[j], d[j], h[j], W[j]) : g2_fwd_node(0.f, 0.f, 1.f, 0.f);
  for (int i = 0; i < N; ++i) {
    float2 acc[6];
    for (int a = 0; a < 6; ++a) acc[a] = f2(0.f, 0.f);
    for (int p = 0; p < npairs; ++p) {
      const G2FwdNode &n0 = nodes[2 * p], &n1 = nodes[2 * p + 1];
      g2_fwd_pair<true>(f2dup(x[i]), f2dup(y[i]), f2(n0.nc, n1.nc), f2(n0.nd, n1.nd), f2(n0.rp, n1.rp),
                        f2(n0.w0, n1.w0), f2(n0.w1, n1.w1), f2(n0.w2, n1.w2), acc);
    }
    for (int a = 0; a < 6; ++a) jets[(size_t)a * N + i] = acc[a].x + acc[a].y + (a == 0 ? bias : 0.f);
  }
}

// mask-specialised general backward (g2_bwd_pair_mask<MASK>): g [5][N] with the ABSENT rows holding
// NaN (they must never enter the arithmetic); outputs dc, dd, dh, dW [M]
void hc_g2_bwd_masked(int mask, int N, int M, const float* x, const float* y, const float* g, const float* c,
                      const float* d, const float* h, const float* W, float* dc, float* dd, float* dh,
                      float* dW) {
  switch (mask) {
    case 0x06: hc_masked_impl<0x06>(N, M, x, y, g, c, d, h, W, dc, dd, dh, dW); break;
    case 0x0C: hc_masked_impl<0x0C>(N, M, x, y, g, c, d, h, W, dc, dd, dh, dW); break;
    case 0x0F: hc_masked_impl<0x0F>(N, M, x, y, g, c, d, h, W, dc, dd, dh, dW); break;
    default: hc_masked_impl<0x1F>(N, M, x, y, g, c, d, h, W, dc, dd, dh, dW); break;
  }
}

// fast backward emulation: g [5][N]; outputs dc, dd, dh, dW [M]
void hc_g2_bwd(int N, int M, const float* x, const float* y, const float* g, const float* c,
               const float* d, const float* h, const float* W, float* dc, float* dd, float* dh,
               float* dW) {
  int N2 = (N + 1) / 2;
  for (int j = 0; j < M; ++j) {
    G2BwdNode n = g2_bwd_node(c[j], d[j], h[j]);
    float tot[4] = {0, 0, 0, 0};
    for (int t0 = 0; t0 < N2; t0 += 256) {
      float2 acc[4];
      for (int a = 0; a < 4; ++a) acc[a] = f2(0.f, 0.f);
      int t1 = t0 + 256 < N2 ? t0 + 256 : N2;
      for (int p = t0; p < t1; ++p) {
        int i0 = 2 * p, i1 = 2 * p + 1;
        bool v = i1 < N;
        float2 gg[5];
        for (int a = 0; a < 5; ++a) gg[a] = f2(g[(size_t)a * N + i0], v ? g[(size_t)a * N + i1] : 0.f);
        g2_bwd_pair(f2(x[i0], v ? x[i1] : x[i0]), f2(y[i0], v ? y[i1] : y[i0]), gg[0], gg[1], gg[2],
                    gg[3], gg[4], f2dup(n.nc), f2dup(n.nd), f2dup(n.rp), f2dup(n.nrt), f2dup(n.rho),
                    f2dup(n.nk1), f2dup(n.k2p), acc);
      }
      for (int a = 0; a < 4; ++a) tot[a] += acc[a].x + acc[a].y;
    }
    g2_bwd_finish(h[j], W[j], tot[0], tot[1], tot[2], tot[3], &dW[j], &dc[j], &dd[j], &dh[j]);
  }
}

// structured-upstream backward emulation: mode 1 (only g3,g4 given as g[0..1][N]) and
// mode 2 (only g0 given as g[0][N]); outputs dc, dd, dh, dW [M]
void hc_g2_bwd_mode(int mode, int N, int M, const float* x, const float* y, const float* g,
                    const float* c, const float* d, const float* h, const float* W, float* dc,
                    float* dd, float* dh, float* dW) {
  int N2 = (N + 1) / 2;
  for (int j = 0; j < M; ++j) {
    G2BwdNode n = g2_bwd_node(c[j], d[j], h[j]);
    float tot[5] = {0, 0, 0, 0, 0};
    float2 acc[5];
    for (int a = 0; a < 5; ++a) acc[a] = f2(0.f, 0.f);
    for (int p = 0; p < N2; ++p) {
      int i0 = 2 * p, i1 = 2 * p + 1;
      bool v = i1 < N;
      float2 x2 = f2(x[i0], v ? x[i1] : x[i0]), y2 = f2(y[i0], v ? y[i1] : y[i0]);
      if (mode == 1) {
        const G2LaplPoint ca = g2_lapl_point(g[i0], g[(size_t)N + i0]);
        const G2LaplPoint cb = v ? g2_lapl_point(g[i1], g[(size_t)N + i1]) : g2_lapl_point(0.f, 0.f);
        g2_bwd_pair_lapl(x2, y2, f2(ca.D, cb.D), f2(ca.g4, cb.g4), f2(ca.m, cb.m), f2(ca.k3, cb.k3),
                         f2(ca.k4, cb.k4), f2dup(n.nc), f2dup(n.nd), f2dup(n.rp), acc);
      } else {
        g2_bwd_pair_u(x2, y2, f2(g[i0], v ? g[i1] : 0.f), f2dup(n.nc), f2dup(n.nd), f2dup(n.rp), acc);
      }
    }
    for (int a = 0; a < 5; ++a) tot[a] = acc[a].x + acc[a].y;
    if (mode == 1) g2_bwd_finish_m<BWD_LAPL>(h[j], W[j], tot, &dW[j], &dc[j], &dd[j], &dh[j]);
    else g2_bwd_finish_m<BWD_U>(h[j], W[j], tot, &dW[j], &dc[j], &dd[j], &dh[j]);
  }
}

}  // extern "C"

template <int DIM, int KIND, int K, int LEVEL>
static void gen_fwd(int N, int M, const float* x, const float* y, const float* c, const float* d,
                    const float* h, const float* W, const float* bias, float* jets) {
  constexpr int NJ = njets(DIM, LEVEL);
  for (int i = 0; i < N; ++i) {
    float acc[NJ][K];
    for (int a = 0; a < NJ; ++a) for (int k = 0; k < K; ++k) acc[a][k] = 0.f;
    for (int j = 0; j < M; ++j) {
      float Wk[K];
      for (int k = 0; k < K; ++k) Wk[k] = W[(size_t)k * M + j];
      generic_fwd_pair<DIM, KIND, K, LEVEL>(x[i] - c[j], DIM == 2 ? y[i] - d[j] : 0.f, 1.0f / h[j], Wk, acc);
    }
    for (int a = 0; a < NJ; ++a) for (int k = 0; k < K; ++k)
      jets[((size_t)a * N + i) * K + k] = acc[a][k] + ((a == 0 && bias) ? bias[k] : 0.f);
  }
}

template <int DIM, int KIND, int K, int LEVEL>
static void gen_bwd(int N, int M, const float* x, const float* y, const float* g, const float* c,
                    const float* d, const float* h, const float* W, float* out /*[NACC][M]*/) {
  constexpr int NG = njets(DIM, LEVEL);
  constexpr int NACC = DIM + 1 + K;
  std::vector<float> gp(NG * K);
  for (int j = 0; j < M; ++j) {
    float acc[NACC];
    for (int a = 0; a < NACC; ++a) acc[a] = 0.f;
    float Wk[K];
    for (int k = 0; k < K; ++k) Wk[k] = W[(size_t)k * M + j];
    for (int i = 0; i < N; ++i) {
      for (int a = 0; a < NG; ++a) for (int k = 0; k < K; ++k) gp[a * K + k] = g[((size_t)a * N + i) * K + k];
      generic_bwd_pair<DIM, KIND, K, LEVEL>(x[i] - c[j], DIM == 2 ? y[i] - d[j] : 0.f, 1.0f / h[j], Wk,
                                            gp.data(), acc);
    }
    for (int a = 0; a < NACC; ++a) out[(size_t)a * M + j] = acc[a];
  }
}

template <int DIM, int KIND, int K, int LEVEL>
static void gen_fwd2(int N, int M, const float* x, const float* y, const float* c, const float* d,
                     const float* h, const float* W, const float* bias, float* jets) {
  constexpr int NJ = njets(DIM, LEVEL);
  for (int i = 0; i < N; i += 2) {
    bool v1 = i + 1 < N;
    F2 px(x[i], v1 ? x[i + 1] : 0.f), py(DIM == 2 ? y[i] : 0.f, (DIM == 2 && v1) ? y[i + 1] : 0.f);
    F2 acc[NJ][K];
    for (int a = 0; a < NJ; ++a) for (int k = 0; k < K; ++k) acc[a][k] = F2(0.f);
    for (int j = 0; j < M; ++j) {
      float Wk[K];
      for (int k = 0; k < K; ++k) Wk[k] = W[(size_t)k * M + j];
      generic_fwd_pair<DIM, KIND, K, LEVEL>(px - c[j], DIM == 2 ? py - d[j] : F2(0.f), 1.0f / h[j], Wk, acc);
    }
    for (int a = 0; a < NJ; ++a) for (int k = 0; k < K; ++k) {
      float b = (a == 0 && bias) ? bias[k] : 0.f;
      jets[((size_t)a * N + i) * K + k] = acc[a][k].v.x + b;
      if (v1) jets[((size_t)a * N + i + 1) * K + k] = acc[a][k].v.y + b;
    }
  }
}

template <int DIM, int KIND, int K, int LEVEL>
static void gen_bwd2(int N, int M, const float* x, const float* y, const float* g, const float* c,
                     const float* d, const float* h, const float* W, float* out /*[NACC][M]*/) {
  constexpr int NG = njets(DIM, LEVEL);
  constexpr int NACC = DIM + 1 + K;
  for (int j = 0; j < M; ++j) {
    F2 acc[NACC];
    for (int a = 0; a < NACC; ++a) acc[a] = F2(0.f);
    float Wk[K];
    for (int k = 0; k < K; ++k) Wk[k] = W[(size_t)k * M + j];
    for (int i = 0; i < N; i += 2) {
      bool v1 = i + 1 < N;
      F2 gp[NG * K];
      for (int a = 0; a < NG; ++a) for (int k = 0; k < K; ++k)
        gp[a * K + k] = F2(g[((size_t)a * N + i) * K + k], v1 ? g[((size_t)a * N + i + 1) * K + k] : 0.f);
      F2 X = F2(x[i], v1 ? x[i + 1] : 0.f) - c[j];
      F2 Y = DIM == 2 ? F2(y[i], v1 ? y[i + 1] : 0.f) - d[j] : F2(0.f);
      generic_bwd_pair<DIM, KIND, K, LEVEL>(X, Y, 1.0f / h[j], Wk, gp, acc);
    }
    for (int a = 0; a < NACC; ++a) out[(size_t)a * M + j] = vlane_sum(acc[a]);
  }
}

extern "C" {

#define HC_DISPATCH_LEVEL(FN, DIM, KIND, K, ...)                                     \
  switch (level) {                                                                   \
    case 0: FN<DIM, KIND, K, 0>(__VA_ARGS__); break;                                 \
    case 1: FN<DIM, KIND, K, 1>(__VA_ARGS__); break;                                 \
    case 2: FN<DIM, KIND, K, 2>(__VA_ARGS__); break;                                 \
    default: FN<DIM, KIND, K, (DIM == 2 ? 3 : 2)>(__VA_ARGS__); break;               \
  }
#define HC_DISPATCH_K(FN, DIM, KIND, ...)                                            \
  switch (K) {                                                                       \
    case 1: HC_DISPATCH_LEVEL(FN, DIM, KIND, 1, __VA_ARGS__) break;                  \
    case 2: HC_DISPATCH_LEVEL(FN, DIM, KIND, 2, __VA_ARGS__) break;                  \
    default: HC_DISPATCH_LEVEL(FN, DIM, KIND, 3, __VA_ARGS__) break;                 \
  }
#define HC_DISPATCH(FN, ...)                                                         \
  if (dim == 2) {                                                                    \
    if (kind == 0) { HC_DISPATCH_K(FN, 2, GAUSSIAN, __VA_ARGS__) }                   \
    else { HC_DISPATCH_K(FN, 2, SOFTPLUS, __VA_ARGS__) }                             \
  } else {                                                                           \
    if (kind == 0) { HC_DISPATCH_K(FN, 1, GAUSSIAN, __VA_ARGS__) }                   \
    else { HC_DISPATCH_K(FN, 1, SOFTPLUS, __VA_ARGS__) }                             \
  }

void hc_generic_fwd(int dim, int kind, int level, int K, int N, int M, const float* x, const float* y,
                    const float* c, const float* d, const float* h, const float* W, const float* bias,
                    float* jets) {
  HC_DISPATCH(gen_fwd, N, M, x, y, c, d, h, W, bias, jets)
}

void hc_generic_bwd(int dim, int kind, int level, int K, int N, int M, const float* x, const float* y,
                    const float* g, const float* c, const float* d, const float* h, const float* W,
                    float* out) {
  HC_DISPATCH(gen_bwd, N, M, x, y, g, c, d, h, W, out)
}

// the same, two points per step in packed lanes (what the CUDA kernels do)
void hc_generic_fwd2(int dim, int kind, int level, int K, int N, int M, const float* x, const float* y,
                     const float* c, const float* d, const float* h, const float* W, const float* bias,
                     float* jets) {
  HC_DISPATCH(gen_fwd2, N, M, x, y, c, d, h, W, bias, jets)
}

void hc_generic_bwd2(int dim, int kind, int level, int K, int N, int M, const float* x, const float* y,
                     const float* g, const float* c, const float* d, const float* h, const float* W,
                     float* out) {
  HC_DISPATCH(gen_bwd2, N, M, x, y, g, c, d, h, W, out)
}

}  // extern "C"

// ---- softplus-hat fast path (s2_fwd_pair / s2_bwd_pair<MASK>): same loop structure as the kernels ----
template <int MASK>
static void hc_s2_bwd_impl(int N, int M, const float* x, const float* y, const float* g, const float* c,
                           const float* d, const float* h, const float* W, float* dc, float* dd, float* dh,
                           float* dW) {
  // lanes = two adjacent NODES (as in the kernel); points one at a time, broadcast
  for (int j = 0; j < M; j += 2) {
    const bool v1 = j + 1 < M;
    const S2BwdNode n0 = s2_bwd_node(c[j], d[j], h[j]);
    const S2BwdNode n1 = v1 ? s2_bwd_node(c[j + 1], d[j + 1], h[j + 1]) : s2_bwd_node(0.f, 0.f, 1.f);
    float2 acc[5];
    for (int a = 0; a < 5; ++a) acc[a] = f2(0.f, 0.f);
    for (int i = 0; i < N; ++i) {
      float2 gg[5];
      for (int a = 0; a < 5; ++a) gg[a] = f2dup(g[(size_t)a * N + i]);
      s2_bwd_pair<MASK>(f2dup(x[i]), f2dup(y[i]), gg[0], gg[1], gg[2], gg[3], gg[4], f2(n0.nc, n1.nc),
                        f2(n0.nd, n1.nd), f2(n0.c1, n1.c1), f2(n0.cr, n1.cr), f2(n0.ncr, n1.ncr),
                        f2(n0.nrho2, n1.nrho2), acc);
    }
    float a0[5], a1[5];
    for (int a = 0; a < 5; ++a) { a0[a] = acc[a].x; a1[a] = acc[a].y; }
    s2_bwd_finish(h[j], W[j], a0, &dW[j], &dc[j], &dd[j], &dh[j]);
    if (v1) s2_bwd_finish(h[j + 1], W[j + 1], a1, &dW[j + 1], &dc[j + 1], &dd[j + 1], &dh[j + 1]);
  }
}

extern "C" {

// jets [6][N]
void hc_s2_fwd(int N, int M, const float* x, const float* y, const float* c, const float* d,
               const float* h, const float* W, float bias, float* jets) {
  int npairs = (M + 1) / 2;
  std::vector<S2FwdNode> nodes(2 * npairs);
  for (int j = 0; j < 2 * npairs; ++j)
    nodes[j] = j < M ? s2_fwd_node(c[j], d[j], h[j], W[j]) : s2_fwd_node(0.f, 0.f, 1.f, 0.f);
  for (int i = 0; i < N; ++i) {
    float2 acc[6];
    for (int a = 0; a < 6; ++a) acc[a] = f2(0.f, 0.f);
    for (int p = 0; p < npairs; ++p) {
      const S2FwdNode &n0 = nodes[2 * p], &n1 = nodes[2 * p + 1];
      s2_fwd_pair<true>(f2dup(x[i]), f2dup(y[i]), f2(n0.nc, n1.nc), f2(n0.nd, n1.nd), f2(n0.c1, n1.c1),
                        f2(n0.w0, n1.w0), f2(n0.w1, n1.w1), f2(n0.w2, n1.w2), acc);
    }
    for (int a = 0; a < 6; ++a) jets[(size_t)a * N + i] = acc[a].x + acc[a].y + (a == 0 ? bias : 0.f);
  }
}

// g [5][N] with NaN in the rows absent from `mask`; outputs dc, dd, dh, dW [M]
void hc_s2_bwd(int mask, int N, int M, const float* x, const float* y, const float* g, const float* c,
               const float* d, const float* h, const float* W, float* dc, float* dd, float* dh, float* dW) {
  switch (mask) {
    case 0x01: hc_s2_bwd_impl<0x01>(N, M, x, y, g, c, d, h, W, dc, dd, dh, dW); break;
    case 0x18: hc_s2_bwd_impl<0x18>(N, M, x, y, g, c, d, h, W, dc, dd, dh, dW); break;
    case 0x07: hc_s2_bwd_impl<0x07>(N, M, x, y, g, c, d, h, W, dc, dd, dh, dW); break;
    case 0x06: hc_s2_bwd_impl<0x06>(N, M, x, y, g, c, d, h, W, dc, dd, dh, dW); break;
    case 0x0C: hc_s2_bwd_impl<0x0C>(N, M, x, y, g, c, d, h, W, dc, dd, dh, dW); break;
    case 0x10: hc_s2_bwd_impl<0x10>(N, M, x, y, g, c, d, h, W, dc, dd, dh, dW); break;
    default: hc_s2_bwd_impl<0x1F>(N, M, x, y, g, c, d, h, W, dc, dd, dh, dW); break;
  }
}

}  // extern "C"

// ---- residual epilogues of the problem scripts (ode3 forcing, cavity point residual + dLoss/dJets) ----
extern "C" {
void hc_ode3_forcing(int n, const float* x, float Kc, float* f) {
  for (int i = 0; i < n; ++i) f[i] = ode3_forcing(x[i], Kc);
}
// J, g: [5][n][3]; loss: [n]
void hc_cavity_residual(int n, const float* J, float nu, float* g, float* loss) {
  for (int i = 0; i < n; ++i) {
    float Jp[5][3], gp[5][3];
    for (int a = 0; a < 5; ++a)
      for (int k = 0; k < 3; ++k) Jp[a][k] = J[((size_t)a * n + i) * 3 + k];
    loss[i] = cavity_residual_point(Jp, nu, gp);
    for (int a = 0; a < 5; ++a)
      for (int k = 0; k < 3; ++k) g[((size_t)a * n + i) * 3 + k] = gp[a][k];
  }
}
}  // extern "C"
