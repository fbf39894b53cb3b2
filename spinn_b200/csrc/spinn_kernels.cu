// spinn_kernels.cu -- sm_100a kernels + C ABI of the SPINN hot path.
//
// What the kernels replace in the reference (paths relative to /root/reference):
//   forward   : SPINN2D.forward / SPINN1D.forward (code/spinn2d.py:169-179,
//               code/spinn1d.py:147-156) and the three nested autograd.grad
//               calls of _compute_derivatives (code/pde2d_base.py:46-62,
//               code/ode_base.py:32-42) -> ONE fused kernel producing all jets;
//   backward  : loss.backward() through that graph (code/common.py:242-248)
//               -> ONE fused kernel producing per-node parameter gradients as
//               per-block partial sums + a small deterministic second stage
//               (no atomics anywhere).
// The (N x M) intermediates the reference materialises never exist here: points
// stream from HBM, nodes live in shared memory (forward) or registers (backward).
//
// Two code paths:
//   * FAST   2-D Gaussian, K=1, level 2: packed f32x2 arithmetic (FFMA2), the
//            BASELINE headline configuration (4M points x 4096 nodes);
//   * GENERIC every other (dim, kernel, K<=4, level): scalar arithmetic built on
//            phi_derivs_* (same tiling, same reduction).
#include <cuda_runtime.h>
#include <atomic>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/spinn_b200.h"
#include "spinn_host.h"
#include "spinn_math.cuh"

using namespace spinn;

// -----------------------------------------------------------------------------
// TMA 1-D bulk copies (cp.async.bulk, SASS UBLKCP) + mbarrier: contiguous record tiles move
// global -> shared memory asynchronously, issued by one thread, completion tracked by the
// barrier's transaction count.  Source/destination 16-byte aligned, size a multiple of 16.
// -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MBAR_DONE;\n"
      "bra MBAR_WAIT;\n"
      "MBAR_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
constexpr unsigned TMA_CHUNK = 32768;      // bytes per bulk copy

// -----------------------------------------------------------------------------
// host-side utilities
// -----------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// shared with spinn_sparse.cu (spinn_host.h)
int spinn_fail_msg(int code, const char* msg) { return fail(code, "%s", msg); }

#define CUDA_TRY(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess)                                                          \
      return fail(SPINN_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                  __FILE__, __LINE__);                                              \
  } while (0)

#define LAUNCH_CHECK(name)                                                          \
  do {                                                                              \
    cudaError_t _e = cudaPeekAtLastError();                                         \
    if (_e != cudaSuccess)                                                          \
      return fail(SPINN_E_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
  } while (0)

static int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

int spinn_sm_count() { return sm_count(); }

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int round_up_i(int v, int a) { return (v + a - 1) / a * a; }
static inline int ceil_div_i(long long a, long long b) { return (int)((a + b - 1) / b); }

// -----------------------------------------------------------------------------
// node packing
// SoA rows (each Mp long): 2-D: c, d, h, r, W_0..W_{K-1};  1-D: c, h, r, W_0..
// padding nodes: c=d=0, h=r=1, W=0 (contribute exactly zero).
// -----------------------------------------------------------------------------
template <int DIM>
__global__ void pack_nodes_soa_kernel(int M, int Mfree, int Mp, int K,
                                      const float* __restrict__ xf, const float* __restrict__ yf,
                                      const float* __restrict__ xfix, const float* __restrict__ yfix,
                                      const float* __restrict__ h, int hs,
                                      const float* __restrict__ W, float* __restrict__ out) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= Mp) return;
  float c = 0.f, d = 0.f, hh = 1.f;
  if (j < M) {
    c = j < Mfree ? xf[j] : xfix[j - Mfree];
    if (DIM == 2) d = j < Mfree ? yf[j] : yfix[j - Mfree];
    hh = hs ? h[0] : h[j];
  }
  int row = 0;
  out[(size_t)(row++) * Mp + j] = c;
  if (DIM == 2) out[(size_t)(row++) * Mp + j] = d;
  out[(size_t)(row++) * Mp + j] = hh;
  out[(size_t)(row++) * Mp + j] = 1.0f / hh;
  for (int k = 0; k < K; ++k) out[(size_t)(row++) * Mp + j] = j < M ? W[(size_t)k * M + j] : 0.f;
}

// fast forward records: one per NODE PAIR (j, j+1), three float4:
//   Gaussian     {nc_j, nc_j1, nd_j, nd_j1} {rp_j, rp_j1, w0_j, w0_j1} {w1_j, w1_j1, w2_j, w2_j1}
//   softplus-hat {nc.., nd..} {c1.., w0..} {w1.., w2..}                       (S2FwdNode)
template <int KIND>
__global__ void pack_nodes_k1fwd_kernel(int M, int Mfree, int npairs,
                                        const float* __restrict__ xf, const float* __restrict__ yf,
                                        const float* __restrict__ xfix, const float* __restrict__ yfix,
                                        const float* __restrict__ h, int hs,
                                        const float* __restrict__ W, float4* __restrict__ recs) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  float v[2][6];
#pragma unroll
  for (int l = 0; l < 2; ++l) {
    int j = 2 * p + l;
    float c = 0.f, d = 0.f, hh = 1.f, w = 0.f;             // padding node: contributes exactly zero
    if (j < M) {
      c = j < Mfree ? xf[j] : xfix[j - Mfree];
      d = j < Mfree ? yf[j] : yfix[j - Mfree];
      hh = hs ? h[0] : h[j];
      w = W[j];
    }
    if (KIND == GAUSSIAN) {
      const G2FwdNode n = g2_fwd_node(c, d, hh, w);
      v[l][0] = n.nc; v[l][1] = n.nd; v[l][2] = n.rp; v[l][3] = n.w0; v[l][4] = n.w1; v[l][5] = n.w2;
    } else {
      const S2FwdNode n = s2_fwd_node(c, d, hh, w);
      v[l][0] = n.nc; v[l][1] = n.nd; v[l][2] = n.c1; v[l][3] = n.w0; v[l][4] = n.w1; v[l][5] = n.w2;
    }
  }
  recs[3 * (size_t)p + 0] = make_float4(v[0][0], v[1][0], v[0][1], v[1][1]);
  recs[3 * (size_t)p + 1] = make_float4(v[0][2], v[1][2], v[0][3], v[1][3]);
  recs[3 * (size_t)p + 2] = make_float4(v[0][4], v[1][4], v[0][5], v[1][5]);
}

// fast backward records: one per POINT PAIR (i, i+1), bwd_nrec(MODE) float4 each:
//   BWD_ALL : {x_i,x_i1,y_i,y_i1} {g0,g0',g1,g1'} {g2,g2',g3,g3'} {g4,g4',0,0}
//   BWD_LAPL: {x..,y..} {D,D',g4,g4'} {m,m',k3,k3'} {k4,k4',0,0}   (g2_lapl_point)
//   BWD_U   : {x..,y..} {g0,g0',0,0}
// rows of g whose bit in `mask` is clear are never read (they count as zero).
template <int MODE>
__global__ void pack_points_g2bwd_kernel(int N, int N2, const float* __restrict__ x,
                                         const float* __restrict__ y, const float* __restrict__ g,
                                         int mask, float4* __restrict__ recs) {
  constexpr int NREC = bwd_nrec(MODE);
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N2) return;
  int i0 = 2 * p, i1 = 2 * p + 1;
  bool v1 = i1 < N;
  size_t n = (size_t)N;
  float x0 = x[i0], y0 = y[i0];
  float x1 = v1 ? x[i1] : x0, y1 = v1 ? y[i1] : y0;
  float a[5], b[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const bool need = bwd_general(MODE) || (MODE == BWD_LAPL && k >= 3) || (MODE == BWD_U && k == 0);
    const bool have = need && ((mask >> k) & 1);
    a[k] = have ? g[k * n + i0] : 0.f;
    b[k] = (have && v1) ? g[k * n + i1] : 0.f;
  }
  float4* r = recs + (size_t)NREC * p;
  r[0] = make_float4(x0, x1, y0, y1);
  if (bwd_general(MODE)) {
    r[1] = make_float4(a[0], b[0], a[1], b[1]);
    r[2] = make_float4(a[2], b[2], a[3], b[3]);
    r[3 % NREC] = make_float4(a[4], b[4], 0.f, 0.f);
  } else if (MODE == BWD_LAPL) {
    const G2LaplPoint ca = g2_lapl_point(a[3], a[4]), cb = g2_lapl_point(b[3], b[4]);
    r[1] = make_float4(ca.D, cb.D, ca.g4, cb.g4);
    r[2 % NREC] = make_float4(ca.m, cb.m, ca.k3, cb.k3);
    r[3 % NREC] = make_float4(ca.k4, cb.k4, 0.f, 0.f);
  } else {
    r[1] = make_float4(a[0], b[0], 0.f, 0.f);
  }
}

// -----------------------------------------------------------------------------
// tuning knobs (production defaults; SPINN_TUNE="fwd=<id>,bwd=<id>,waves=<n>" overrides
// them at call time -- used by tools/tune.py to pick the defaults on the box)
// -----------------------------------------------------------------------------
struct FwdCfg { int p, threads, minb, unroll; };
struct BwdCfg { int q, threads, minb, tile; };
static const FwdCfg FWD_CFGS[] = {
    {4, 256, 2, 2},   // 0
    {2, 256, 2, 4},   // 1  (alias of 3)
    {1, 512, 2, 4},   // 2
    {2, 256, 2, 4},   // 3  (default: best on B200 at 4M and 512K points, tools/tune.py)
};
static const BwdCfg BWD_CFGS[] = {
    {2, 128, 4, 256},  // 0
    {1, 128, 8, 256},  // 1
    {2, 256, 2, 256},  // 2  (default)
    {4, 128, 2, 256},  // 3
    {1, 64, 8, 256},   // 4  (64-node tile for small node sets)
    {4, 256, 2, 256},  // 5  (1024-node tile; two node-pair streams per thread in the structured modes)
    {2, 256, 4, 256},  // 6  (as 2 with 64 registers/thread, 4 CTAs/SM)
    {4, 128, 3, 256},  // 7
};
// softplus-hat fast backward: node tile = q*threads (lanes are node pairs, q even)
static const BwdCfg S2_BWD_CFGS[] = {
    {2, 256, 2, 256},  // 0  512-node tile (default)
    {2, 128, 4, 256},  // 1  256-node tile
    {2, 64, 8, 256},   // 2  128-node tile
    {2, 32, 16, 256},  // 3  64-node tile
};
constexpr int N_S2_BWD_CFGS = sizeof(S2_BWD_CFGS) / sizeof(S2_BWD_CFGS[0]);
constexpr int N_FWD_CFGS = sizeof(FWD_CFGS) / sizeof(FWD_CFGS[0]);
constexpr int N_BWD_CFGS = sizeof(BWD_CFGS) / sizeof(BWD_CFGS[0]);
constexpr int FWD_DEFAULT = 3, BWD_DEFAULT = 2, WAVES_DEFAULT = 2;   // picked with tools/tune.py on B200

// The knobs live in three atomics: parsed from SPINN_TUNE ONCE when the library is loaded, changed
// afterwards only through spinn_set_tuning() (tools/tune.py).  The entry points never touch the
// environment.
struct Tune { int fwd, bwd, waves; bool bwd_forced; };
static std::atomic<int> g_tune_fwd{FWD_DEFAULT}, g_tune_bwd{-1}, g_tune_waves{WAVES_DEFAULT};

static void parse_tuning(const char* e) {
  int fwd = FWD_DEFAULT, bwd = -1, waves = WAVES_DEFAULT;
  if (e) {
    const char* q;
    if ((q = strstr(e, "fwd="))) fwd = atoi(q + 4);
    if ((q = strstr(e, "bwd="))) bwd = atoi(q + 4);
    if ((q = strstr(e, "waves="))) waves = atoi(q + 6);
  }
  if (fwd < 0 || fwd >= N_FWD_CFGS) fwd = FWD_DEFAULT;
  if (bwd >= N_BWD_CFGS) bwd = -1;
  if (waves < 1 || waves > 64) waves = WAVES_DEFAULT;
  g_tune_fwd.store(fwd); g_tune_bwd.store(bwd); g_tune_waves.store(waves);
}
namespace { struct TuneInit { TuneInit() { parse_tuning(getenv("SPINN_TUNE")); } } g_tune_init; }

static Tune get_tune() {
  Tune t;
  t.fwd = g_tune_fwd.load(std::memory_order_relaxed);
  const int b = g_tune_bwd.load(std::memory_order_relaxed);
  t.bwd_forced = b >= 0;
  t.bwd = b >= 0 ? b : BWD_DEFAULT;
  t.waves = g_tune_waves.load(std::memory_order_relaxed);
  return t;
}

// cudaFuncSetAttribute once per (kernel instantiation, device); safe with concurrent host threads
struct OnceFlags {
  std::atomic<int> st[64];
  OnceFlags() { for (auto& v : st) v.store(0); }
};
template <class F>
static cudaError_t opt_in_smem_once(OnceFlags& flags, int dev, F kfn, int bytes) {
  if (dev < 0 || dev >= 64) return cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (flags.st[dev].load(std::memory_order_acquire) == 1) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) flags.st[dev].store(1, std::memory_order_release);   // idempotent: a race repeats the call
  return e;
}

// -----------------------------------------------------------------------------
// FAST forward: 2-D Gaussian, K=1.  Persistent CTAs; the node-pair records (the whole
// node set when it fits: 48 B per pair, 96 KB at M=4096) are staged in shared memory
// ONCE per CTA; each WARP then walks warp-tiles of 32*P consecutive points (thread =
// P consecutive points in registers, coalesced vector loads/stores).  Packed f32x2
// math where the two lanes are the two nodes of a pair; records are broadcast LDS.128.
// -----------------------------------------------------------------------------
template <int P>
__device__ __forceinline__ void load_points(const float* __restrict__ p, long long i0, int N, bool vec,
                                            float* out) {
  if (vec && i0 + P - 1 < N) {
    if (P % 4 == 0) {
#pragma unroll
      for (int k = 0; k < P / 4; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(p + i0 + 4 * k);
        out[(4 * k) % P] = v.x; out[(4 * k + 1) % P] = v.y; out[(4 * k + 2) % P] = v.z; out[(4 * k + 3) % P] = v.w;
      }
      return;
    }
    if (P == 2) {
      const float2 v = *reinterpret_cast<const float2*>(p + i0);
      out[0] = v.x; out[1 % P] = v.y;
      return;
    }
  }
#pragma unroll
  for (int q = 0; q < P; ++q) out[q] = (i0 + q < N) ? p[i0 + q] : 0.f;
}

template <int P>
__device__ __forceinline__ void store_points(float* __restrict__ p, long long i0, int N, bool vec,
                                             const float* v) {
  if (vec && i0 + P - 1 < N) {
    if (P % 4 == 0) {
#pragma unroll
      for (int k = 0; k < P / 4; ++k)
        *reinterpret_cast<float4*>(p + i0 + 4 * k) =
            make_float4(v[(4 * k) % P], v[(4 * k + 1) % P], v[(4 * k + 2) % P], v[(4 * k + 3) % P]);
      return;
    }
    if (P == 2) {
      *reinterpret_cast<float2*>(p + i0) = make_float2(v[0], v[1 % P]);
      return;
    }
  }
#pragma unroll
  for (int q = 0; q < P; ++q)
    if (i0 + q < N) p[i0 + q] = v[q];
}

template <int KIND, int P, bool MIXED, int UNROLL>
__device__ __forceinline__ void k1_fwd_inner(const float4* __restrict__ sm4, int n,
                                             const float2* x2, const float2* y2,
                                             float2 (*acc)[MIXED ? 6 : 5]) {
#pragma unroll UNROLL
  for (int p = 0; p < n; ++p) {
    const float4 A = sm4[3 * p + 0], B = sm4[3 * p + 1], C = sm4[3 * p + 2];
    const float2 nc = f2(A.x, A.y), nd = f2(A.z, A.w), rp = f2(B.x, B.y);
    const float2 w0 = f2(B.z, B.w), w1 = f2(C.x, C.y), w2 = f2(C.z, C.w);
#pragma unroll
    for (int q = 0; q < P; ++q) {
      if (KIND == GAUSSIAN) g2_fwd_pair<MIXED>(x2[q], y2[q], nc, nd, rp, w0, w1, w2, acc[q]);
      else s2_fwd_pair<MIXED>(x2[q], y2[q], nc, nd, rp, w0, w1, w2, acc[q]);
    }
  }
}

// Residual epilogue of the stock Poisson residual, run by the forward kernel on the jets it has just
// formed: res = scale (u_xx + u_yy + amp sin(kx x) sin(ky y) + f0), the loss
// partial sum res^2 / n_total, and dLoss/du_xx = dLoss/du_yy = g = 2 scale res / n_total written
// straight into the backward's point-pair record (layouts: poisson_residual_pack_kernel).
struct FwdEpilogue {
  float scale, amp, kx, ky, f0, inv_n;
  float4* recs;          // [4 * ceil(N/2)] backward point-pair records
  float* lpart;          // one loss partial per warp of the grid
};

template <int KIND>
__device__ __forceinline__ void write_point_record(float4* __restrict__ r, float xa, float xb, float ya, float yb,
                                                   float ga, float gb) {
  r[0] = make_float4(xa, xb, ya, yb);
  if (KIND == GAUSSIAN) {
    const G2LaplPoint ca = g2_lapl_point(ga, ga), cb = g2_lapl_point(gb, gb);
    r[1] = make_float4(ca.D, cb.D, ca.g4, cb.g4);
    r[2] = make_float4(ca.m, cb.m, ca.k3, cb.k3);
    r[3] = make_float4(ca.k4, cb.k4, 0.f, 0.f);
  } else {
    r[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    r[2] = make_float4(0.f, 0.f, ga, gb);
    r[3] = make_float4(ga, gb, 0.f, 0.f);
  }
}

// The epilogue of one warp-tile, fed from global memory (the thread re-reads the u_xx, u_yy it has
// stored itself: L2 hits).  It runs in a second walk over the warp's tiles AFTER the jets loop:
// placed inside that loop it changed nothing in the inner loop's instruction count but perturbed
// ptxas' schedule of it enough to cost 4 % of the kernel (profiles/r02c: 8.08 vs 7.76 ms).
// Returns this thread's sum of res^2.
template <int KIND, int P>
__device__ __forceinline__ float fwd_epilogue_tile(long long i0, int lim, int N, const float* __restrict__ x,
                                                const float* __restrict__ y, const float* jets,
                                                const FwdEpilogue& ep) {
  const int lane = threadIdx.x & 31;
  float px[P], py[P], g[P], s = 0.f;
#pragma unroll
  for (int q = 0; q < P; ++q) {
    const bool v = i0 + q < lim;
    px[q] = v ? x[i0 + q] : 0.f;
    py[q] = v ? y[i0 + q] : 0.f;
    const float lap = v ? jets[3 * (size_t)N + i0 + q] + jets[4 * (size_t)N + i0 + q] : 0.f;
    const float res = v ? ep.scale * (lap + fmaf(ep.amp * sinf(ep.kx * px[q]), sinf(ep.ky * py[q]), ep.f0)) : 0.f;
    s = fmaf(res, res, s);
    g[q] = 2.f * ep.scale * res * ep.inv_n;
  }
  if (P == 1) {
    // one point per thread: lanes (2k, 2k+1) form the pair (tile bases are multiples of 32)
    const float xb = __shfl_down_sync(0xffffffffu, px[0], 1), yb = __shfl_down_sync(0xffffffffu, py[0], 1);
    const float gb = __shfl_down_sync(0xffffffffu, g[0], 1);
    if ((lane & 1) == 0 && i0 < lim) {
      const bool v1 = i0 + 1 < lim;
      write_point_record<KIND>(ep.recs + 4 * (size_t)(i0 >> 1), px[0], v1 ? xb : px[0], py[0], v1 ? yb : py[0],
                               g[0], v1 ? gb : 0.f);
    }
  } else {
#pragma unroll
    for (int q = 0; q < P; q += 2) {
      if (i0 + q < lim) {
        const bool v1 = i0 + q + 1 < lim;
        write_point_record<KIND>(ep.recs + 4 * (size_t)((i0 + q) >> 1), px[q], v1 ? px[(q + 1) % P] : px[q], py[q],
                                 v1 ? py[(q + 1) % P] : py[q], g[q], v1 ? g[(q + 1) % P] : 0.f);
      }
    }
  }
  return s;
}

// One span of the point set, walked in warp-tiles of 32*P consecutive points: tile `wt` of the span
// starts at point `base + wt*32*P`; the span ends at `end`.  The whole node table is in shared memory.
template <int KIND, int P, bool MIXED, int THREADS, int UNROLL>
__device__ __forceinline__ void k1_fwd_span(long long base, long long end, int N, const float* __restrict__ x,
                                            const float* __restrict__ y, const float4* __restrict__ sm4, int npairs,
                                            float b0, float* __restrict__ jets, bool vec, long long gw,
                                            long long nwarps) {
  constexpr int NJ = MIXED ? 6 : 5;
  const int lane = threadIdx.x & 31;
  const long long nwt = (end - base + 32 * P - 1) / (32 * P);
  const long long iters = (nwt + nwarps - 1) / nwarps;
  for (long long it = 0; it < iters; ++it) {
    const long long wt = gw + it * nwarps;
    if (wt >= nwt) break;                          // warp-uniform; no block-level syncs on this path
    const long long i0 = base + (wt * 32 + lane) * P;
    const int lim = (int)end;                      // points at or beyond `end` belong to another span
    float px[P], py[P];
    load_points<P>(x, i0, lim, vec, px);
    load_points<P>(y, i0, lim, vec, py);
    float2 x2[P], y2[P], acc[P][NJ];
#pragma unroll
    for (int q = 0; q < P; ++q) {
      x2[q] = f2dup(px[q]);
      y2[q] = f2dup(py[q]);
#pragma unroll
      for (int a = 0; a < NJ; ++a) acc[q][a] = f2(0.f, 0.f);
    }
    k1_fwd_inner<KIND, P, MIXED, UNROLL>(sm4, npairs, x2, y2, acc);
#pragma unroll
    for (int a = 0; a < NJ; ++a) {
      float out[P];
#pragma unroll
      for (int q = 0; q < P; ++q) out[q] = acc[q][a].x + acc[q][a].y + (a == 0 ? b0 : 0.f);
      // `lim` (not N) bounds the stores of this span; the jets rows are N long
      store_points<P>(jets + (size_t)a * N, i0, lim, vec, out);
    }
  }
}

// The same walk when the node table does not fit shared memory (more than 4096 nodes): every
// warp-tile of points streams ALL node-record tiles through a 2-stage ring filled by TMA bulk copies
// -- tile g+1 is in flight while tile g is consumed; one block-wide barrier per tile hands the
// consumed stage back to the producer (thread 0), so all warps of the CTA run the same number of
// iterations (inactive ones on dummy points).
template <int KIND, int P, bool MIXED, int THREADS, int UNROLL>
__device__ __forceinline__ void k1_fwd_span_multi(int N, const float* __restrict__ x, const float* __restrict__ y,
                                                  const float4* __restrict__ recs, float4* __restrict__ sm4,
                                                  int npairs, int tile_pairs, float b0, float* __restrict__ jets,
                                                  bool vec, long long gw, long long nwarps, uint64_t* fullb) {
  constexpr int NJ = MIXED ? 6 : 5;
  const int lane = threadIdx.x & 31;
  const long long nwt = ((long long)N + 32 * P - 1) / (32 * P);
  const long long iters = (nwt + nwarps - 1) / nwarps;
  const int ntile = (npairs + tile_pairs - 1) / tile_pairs;
  const long long total = iters * ntile;
  long long g = 0;
  auto issue = [&](long long gi) {                 // thread 0 only
    const int t = (int)(gi % ntile), st = (int)(gi & 1);
    const int n = min(tile_pairs, npairs - t * tile_pairs);
    const unsigned bytes = 3u * (unsigned)n * (unsigned)sizeof(float4);
    mbar_expect_tx(&fullb[st], bytes);
    for (unsigned off = 0; off < bytes; off += TMA_CHUNK)
      tma_bulk_g2s(reinterpret_cast<char*>(sm4 + (size_t)st * 3 * tile_pairs) + off,
                   reinterpret_cast<const char*>(recs + 3 * (size_t)t * tile_pairs) + off,
                   min(TMA_CHUNK, bytes - off), &fullb[st]);
  };
  if (threadIdx.x == 0 && total > 0) issue(0);
  for (long long it = 0; it < iters; ++it) {
    const long long wt = gw + it * nwarps;
    const bool active = wt < nwt;                  // warp-uniform
    const long long i0 = (wt * 32 + lane) * P;
    float px[P], py[P];
    load_points<P>(x, active ? i0 : (long long)N, N, vec, px);
    load_points<P>(y, active ? i0 : (long long)N, N, vec, py);
    float2 x2[P], y2[P], acc[P][NJ];
#pragma unroll
    for (int q = 0; q < P; ++q) {
      x2[q] = f2dup(px[q]);
      y2[q] = f2dup(py[q]);
#pragma unroll
      for (int a = 0; a < NJ; ++a) acc[q][a] = f2(0.f, 0.f);
    }
    for (int t = 0; t < ntile; ++t, ++g) {
      const int st = (int)(g & 1);
      const int n = min(tile_pairs, npairs - t * tile_pairs);
      if (threadIdx.x == 0 && g + 1 < total) issue(g + 1);        // stage st^1 was released by the barrier below
      mbar_wait(&fullb[st], (unsigned)((g >> 1) & 1));
      k1_fwd_inner<KIND, P, MIXED, UNROLL>(sm4 + (size_t)st * 3 * tile_pairs, n, x2, y2, acc);
      __syncthreads();                                            // everyone is done reading stage st
    }
    if (active) {
#pragma unroll
      for (int a = 0; a < NJ; ++a) {
        float out[P];
#pragma unroll
        for (int q = 0; q < P; ++q) out[q] = acc[q][a].x + acc[q][a].y + (a == 0 ? b0 : 0.f);
        store_points<P>(jets + (size_t)a * N, i0, N, vec, out);
      }
    }
  }
}

// Second walk over the same warp-tiles, after the jets of ALL of this warp's tiles are stored: the
// residual epilogue (kept out of the jets loop so that the hot loop's code is the plain kernel's).
template <int KIND, int P>
__device__ __forceinline__ float k1_fwd_epilogue_span(long long base, long long end, int N,
                                                      const float* __restrict__ x, const float* __restrict__ y,
                                                      const float* jets, long long gw, long long nwarps,
                                                      const FwdEpilogue& ep) {
  const int lane = threadIdx.x & 31;
  const long long nwt = (end - base + 32 * P - 1) / (32 * P);
  float s = 0.f;
  for (long long wt = gw; wt < nwt; wt += nwarps)
    s += fwd_epilogue_tile<KIND, P>(base + (wt * 32 + lane) * P, (int)end, N, x, y, jets, ep);
  return s;
}

// Persistent CTAs.  Main span: as many FULL rounds (every warp one tile of 32*P points) as fit;
// the remainder -- less than one round -- is walked with one point per thread (tiles of 32 points), so
// that the last, partial round keeps (nearly) every warp of every SM busy at half the work instead of
// leaving half of the warps idle (8-GPU shards: 3.46 rounds -> 3 rounds + one half-size round).
template <int KIND, int P, bool MIXED, int THREADS, int UNROLL, bool EPI, bool MULTI>
__device__ __forceinline__ void k1_fwd_body(int N, const float* __restrict__ x, const float* __restrict__ y,
                                            const float4* __restrict__ recs, int npairs, int tile_pairs,
                                            const float* __restrict__ bias, float* __restrict__ jets,
                                            int vec_ok, float4* sm4, const FwdEpilogue& ep) {
  constexpr int WPC = THREADS / 32;
  // slot-major warp rank: warps 0-3 of every CTA first, then warps 4-7, ... so that a partial
  // round of warp-tiles is spread over all SMs / sub-partitions instead of leaving whole CTAs idle
  const int lw = threadIdx.x >> 5;
  const long long gw = (long long)(lw >> 2) * ((long long)gridDim.x * 4) + (long long)blockIdx.x * 4 + (lw & 3);
  const long long nwarps = (long long)gridDim.x * WPC;
  constexpr bool single = !MULTI;                  // MULTI: node table larger than shared memory (own kernel)
  const bool vec = vec_ok != 0;
  const float b0 = bias ? bias[0] : 0.f;
  float lsum = 0.f;                                // EPI: this thread's share of sum res^2

  if (single) {
    // the whole node-record table (48 B per node pair, 96 KB at M=4096) arrives by TMA bulk copies
    __shared__ __align__(8) uint64_t bar;
    const unsigned bytes = 3u * (unsigned)npairs * (unsigned)sizeof(float4);
    if (threadIdx.x == 0) {
      mbar_init(&bar, 1);
      mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(&bar, bytes);
      for (unsigned off = 0; off < bytes; off += TMA_CHUNK)
        tma_bulk_g2s(reinterpret_cast<char*>(sm4) + off, reinterpret_cast<const char*>(recs) + off,
                     min(TMA_CHUNK, bytes - off), &bar);
    }
    mbar_wait(&bar, 0);
    const long long round_pts = nwarps * 32 * P;
    const long long main_end = P == 1 ? (long long)N : ((long long)N / round_pts) * round_pts;
    k1_fwd_span<KIND, P, MIXED, THREADS, UNROLL>(0, main_end, N, x, y, sm4, npairs, b0, jets, vec, gw, nwarps);
    if (P > 1 && main_end < N)
      k1_fwd_span<KIND, 1, MIXED, THREADS, UNROLL>(main_end, N, N, x, y, sm4, npairs, b0, jets, vec, gw, nwarps);
    if (EPI) {
      lsum = k1_fwd_epilogue_span<KIND, P>(0, main_end, N, x, y, jets, gw, nwarps, ep);
      if (P > 1 && main_end < N) lsum += k1_fwd_epilogue_span<KIND, 1>(main_end, N, N, x, y, jets, gw, nwarps, ep);
    }
  } else {
    __shared__ __align__(8) uint64_t fullb[2];
    if (threadIdx.x == 0) {
      mbar_init(&fullb[0], 1);
      mbar_init(&fullb[1], 1);
      mbar_fence_init();
    }
    __syncthreads();
    k1_fwd_span_multi<KIND, P, MIXED, THREADS, UNROLL>(N, x, y, recs, sm4, npairs, tile_pairs, b0, jets, vec, gw,
                                                       nwarps, fullb);
    if (EPI) lsum = k1_fwd_epilogue_span<KIND, P>(0, N, N, x, y, jets, gw, nwarps, ep);
  }
  if (EPI) {                                       // one loss partial per warp, every warp writes its slot
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    if ((threadIdx.x & 31) == 0) ep.lpart[gw] = lsum * ep.inv_n;
  }
}

// EPI: the forward also runs the Poisson residual epilogue (FwdEpilogue) on its jets
template <int P, bool MIXED, int THREADS, int MINB, int UNROLL, bool EPI, bool MULTI>
__global__ void __launch_bounds__(THREADS, MINB)
g2k1_fwd_kernel(int N, const float* __restrict__ x, const float* __restrict__ y,
                const float4* __restrict__ recs, int npairs, int tile_pairs,
                const float* __restrict__ bias, float* __restrict__ jets, int vec_ok, const FwdEpilogue ep) {
  extern __shared__ float4 sm4[];
  k1_fwd_body<GAUSSIAN, P, MIXED, THREADS, UNROLL, EPI, MULTI>(N, x, y, recs, npairs, tile_pairs, bias, jets, vec_ok, sm4, ep);
}

// the same persistent, TMA-staged forward with the softplus-hat pair function (spinn_math.cuh, FAST PATH 2)
template <int P, bool MIXED, int THREADS, int MINB, int UNROLL, bool EPI, bool MULTI>
__global__ void __launch_bounds__(THREADS, MINB)
s2k1_fwd_kernel(int N, const float* __restrict__ x, const float* __restrict__ y,
                const float4* __restrict__ recs, int npairs, int tile_pairs,
                const float* __restrict__ bias, float* __restrict__ jets, int vec_ok, const FwdEpilogue ep) {
  extern __shared__ float4 sm4[];
  k1_fwd_body<SOFTPLUS, P, MIXED, THREADS, UNROLL, EPI, MULTI>(N, x, y, recs, npairs, tile_pairs, bias, jets, vec_ok, sm4, ep);
}

// -----------------------------------------------------------------------------
// FAST backward: 2-D Gaussian, K=1, level 2.  Thread = Q nodes (registers), point
// pair records broadcast from shared memory, packed lanes = the two points of a
// pair.  grid = (node groups, point chunks); every block writes its partial sums
// to partials[chunk][row][node]; finalize_kernel reduces over chunks in order.
// Two-level accumulation (per tile, then per chunk) keeps fp32 round-off at the
// level of the reference's pairwise torch.sum.
// -----------------------------------------------------------------------------
template <int Q, int THREADS, int MINB, int TILE, int MODE>
__global__ void __launch_bounds__(THREADS, MINB)
g2k1_bwd_kernel(int N2, int pairs_per_chunk, const float4* __restrict__ ptrecs,
                const float* __restrict__ soa, int Mp, float* __restrict__ partials) {
  constexpr int NREC = bwd_nrec(MODE), NACC = bwd_nacc(MODE);
  // Lane assignment.  General mode: the two lanes of a packed instruction are the two POINTS of a
  // record pair (node constants are scalar-broadcast operands).  Structured modes with an even
  // number of nodes per thread: the two lanes are two adjacent NODES and the per-point constants
  // (5 of them in the (u_xx,u_yy) mode) are the scalar-broadcast operands -- an FFMA2 with three
  // full register-pair operands issues at half rate (tools/microbench.py probe 3), and this
  // assignment leaves only the three accumulate-products in that class.
  constexpr bool LN = (MODE == BWD_LAPL || MODE == BWD_U) && (Q % 2 == 0);
  constexpr int QS = LN ? Q / 2 : Q;             // instruction streams per thread
  // ring of point-record tiles filled by TMA bulk copies: tiles t+1.. are in flight while
  // tile t is consumed; `full[s]` flips when the bytes of stage s have landed
  constexpr int NS = NREC <= 3 ? 3 : 2;          // ring depth (static shared memory <= 48 KB)
  __shared__ __align__(16) float4 ring[NS][NREC * TILE];
  __shared__ __align__(8) uint64_t full[NS];
  const int jblock = blockIdx.x * (Q * THREADS);
  float2 nc[QS], nd[QS], rp[QS], nrt[QS], rho[QS], nk1[QS], k2p[QS];
  float hq[Q], wq[Q];
#pragma unroll
  for (int q = 0; q < QS; ++q) {
    if (LN) {                                    // nodes j, j+1 in the two lanes
      const int j = jblock + 2 * (q * THREADS + threadIdx.x);
      const float2 c = *reinterpret_cast<const float2*>(soa + j);
      const float2 d = *reinterpret_cast<const float2*>(soa + (size_t)Mp + j);
      const float2 h = *reinterpret_cast<const float2*>(soa + 2 * (size_t)Mp + j);
      const float2 w = *reinterpret_cast<const float2*>(soa + 4 * (size_t)Mp + j);
      hq[(2 * q) % Q] = h.x; hq[(2 * q + 1) % Q] = h.y;
      wq[(2 * q) % Q] = w.x; wq[(2 * q + 1) % Q] = w.y;
      const G2BwdNode n0 = g2_bwd_node(c.x, d.x, h.x), n1 = g2_bwd_node(c.y, d.y, h.y);
      nc[q] = f2(n0.nc, n1.nc); nd[q] = f2(n0.nd, n1.nd); rp[q] = f2(n0.rp, n1.rp);
      nrt[q] = f2(n0.nrt, n1.nrt); rho[q] = f2(n0.rho, n1.rho); nk1[q] = f2(n0.nk1, n1.nk1);
      k2p[q] = f2(n0.k2p, n1.k2p);
    } else {
      const int j = jblock + threadIdx.x + q * THREADS;
      const float c = soa[j], d = soa[(size_t)Mp + j];
      hq[q % Q] = soa[2 * (size_t)Mp + j];
      wq[q % Q] = soa[4 * (size_t)Mp + j];
      const G2BwdNode n = g2_bwd_node(c, d, hq[q % Q]);
      nc[q] = f2dup(n.nc); nd[q] = f2dup(n.nd); rp[q] = f2dup(n.rp); nrt[q] = f2dup(n.nrt);
      rho[q] = f2dup(n.rho); nk1[q] = f2dup(n.nk1); k2p[q] = f2dup(n.k2p);
    }
  }
  float2 tot[QS][NACC];                          // LN: per-lane (= per-node) sums; else .x only
#pragma unroll
  for (int q = 0; q < QS; ++q)
#pragma unroll
    for (int a = 0; a < NACC; ++a) tot[q][a] = f2(0.f, 0.f);

  const int p_begin = blockIdx.y * pairs_per_chunk;
  const int p_end = min(N2, p_begin + pairs_per_chunk);
  const int ntiles = p_end > p_begin ? (p_end - p_begin + TILE - 1) / TILE : 0;
  auto issue = [&](int t) {                    // thread 0 only
    const int t0 = p_begin + t * TILE;
    const unsigned bytes = (unsigned)(min(TILE, p_end - t0) * NREC) * (unsigned)sizeof(float4);
    mbar_expect_tx(&full[t % NS], bytes);
    tma_bulk_g2s(ring[t % NS], ptrecs + NREC * (size_t)t0, bytes, &full[t % NS]);
  };
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NS; ++i) mbar_init(&full[i], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int i = 0; i < NS - 1 && i < ntiles; ++i) issue(i);
  for (int t = 0; t < ntiles; ++t) {
    const int t0 = p_begin + t * TILE;
    const int n = min(TILE, p_end - t0);
    const float4* sm4 = ring[t % NS];
    // stage (t+NS-1)%NS was consumed in iteration t-1 (block-wide sync at its end): refill it now
    if (threadIdx.x == 0 && t + NS - 1 < ntiles) issue(t + NS - 1);
    mbar_wait(&full[t % NS], (unsigned)((t / NS) & 1));
    float2 acc[QS][NACC];
#pragma unroll
    for (int q = 0; q < QS; ++q)
#pragma unroll
      for (int a = 0; a < NACC; ++a) acc[q][a] = f2(0.f, 0.f);
#pragma unroll 2
    for (int p = 0; p < n; ++p) {
      const float4 A = sm4[NREC * p + 0], B = sm4[NREC * p + 1];
      const float2 x2 = f2(A.x, A.y), y2 = f2(A.z, A.w);
      if (bwd_general(MODE)) {
        const float4 C = sm4[NREC * p + 2 % NREC], D = sm4[NREC * p + 3 % NREC];
        const float2 g0 = f2(B.x, B.y), g1 = f2(B.z, B.w), g2 = f2(C.x, C.y), g3 = f2(C.z, C.w);
        const float2 g4 = f2(D.x, D.y);
#pragma unroll
        for (int q = 0; q < QS; ++q) {
          if (MODE == BWD_ALL)
            g2_bwd_pair(x2, y2, g0, g1, g2, g3, g4, nc[q], nd[q], rp[q], nrt[q], rho[q], nk1[q],
                                k2p[q], acc[q]);
          else
            g2_bwd_pair_mask<(MODE >= BWD_MASKED ? MODE - BWD_MASKED : 0x1F)>(
                x2, y2, g0, g1, g2, g3, g4, nc[q], nd[q], rp[q], nrt[q], rho[q], nk1[q], k2p[q], acc[q]);
        }
      } else if (MODE == BWD_LAPL) {
        const float4 C = sm4[NREC * p + 2 % NREC];
        const float2 k4 = *reinterpret_cast<const float2*>(&sm4[NREC * p + 3 % NREC]);
        if (LN) {
#pragma unroll
          for (int q = 0; q < QS; ++q) {
            g2_bwd_pair_lapl(f2dup(A.x), f2dup(A.z), f2dup(B.x), f2dup(B.z), f2dup(C.x), f2dup(C.z),
                             f2dup(k4.x), nc[q], nd[q], rp[q], acc[q]);
            g2_bwd_pair_lapl(f2dup(A.y), f2dup(A.w), f2dup(B.y), f2dup(B.w), f2dup(C.y), f2dup(C.w),
                             f2dup(k4.y), nc[q], nd[q], rp[q], acc[q]);
          }
        } else {
#pragma unroll
          for (int q = 0; q < QS; ++q)
            g2_bwd_pair_lapl(x2, y2, f2(B.x, B.y), f2(B.z, B.w), f2(C.x, C.y), f2(C.z, C.w), k4, nc[q],
                             nd[q], rp[q], acc[q]);
        }
      } else {
        if (LN) {
#pragma unroll
          for (int q = 0; q < QS; ++q) {
            g2_bwd_pair_u(f2dup(A.x), f2dup(A.z), f2dup(B.x), nc[q], nd[q], rp[q], acc[q]);
            g2_bwd_pair_u(f2dup(A.y), f2dup(A.w), f2dup(B.y), nc[q], nd[q], rp[q], acc[q]);
          }
        } else {
          const float2 g0 = f2(B.x, B.y);
#pragma unroll
          for (int q = 0; q < QS; ++q) g2_bwd_pair_u(x2, y2, g0, nc[q], nd[q], rp[q], acc[q]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < QS; ++q)
#pragma unroll
      for (int a = 0; a < NACC; ++a) {
        if (LN) { tot[q][a].x += acc[q][a].x; tot[q][a].y += acc[q][a].y; }
        else tot[q][a].x += acc[q][a].x + acc[q][a].y;
      }
    __syncthreads();                                   // everyone is done reading this stage
  }
  // rows: 0 dc, 1 dd, 2 dh, 3 dW
  float* out = partials + (size_t)blockIdx.y * 4 * Mp;
#pragma unroll
  for (int q = 0; q < QS; ++q) {
    if (LN) {
      const int j = jblock + 2 * (q * THREADS + threadIdx.x);
      float t0[NACC], t1[NACC];
#pragma unroll
      for (int a = 0; a < NACC; ++a) { t0[a] = tot[q][a].x; t1[a] = tot[q][a].y; }
      float2 dW, dc, dd, dh;
      g2_bwd_finish_m<MODE>(hq[(2 * q) % Q], wq[(2 * q) % Q], t0, &dW.x, &dc.x, &dd.x, &dh.x);
      g2_bwd_finish_m<MODE>(hq[(2 * q + 1) % Q], wq[(2 * q + 1) % Q], t1, &dW.y, &dc.y, &dd.y, &dh.y);
      *reinterpret_cast<float2*>(out + j) = dc;
      *reinterpret_cast<float2*>(out + (size_t)Mp + j) = dd;
      *reinterpret_cast<float2*>(out + 2 * (size_t)Mp + j) = dh;
      *reinterpret_cast<float2*>(out + 3 * (size_t)Mp + j) = dW;
    } else {
      const int j = jblock + threadIdx.x + q * THREADS;
      float t0[NACC];
#pragma unroll
      for (int a = 0; a < NACC; ++a) t0[a] = tot[q][a].x;
      float dW, dc, dd, dh;
      g2_bwd_finish_m<MODE>(hq[q % Q], wq[q % Q], t0, &dW, &dc, &dd, &dh);
      out[j] = dc;
      out[(size_t)Mp + j] = dd;
      out[2 * (size_t)Mp + j] = dh;
      out[3 * (size_t)Mp + j] = dW;
    }
  }
}

// -----------------------------------------------------------------------------
// FAST backward, softplus-hat: 2-D, K=1, level <= 2.  Same skeleton as g2k1_bwd_kernel (thread =
// Q nodes in registers, point-pair record tiles through a TMA bulk-copy ring, partial sums per
// (chunk, node), no atomics); the two lanes of every packed instruction are two adjacent NODES, the
// point's coordinates and upstream gradients are scalar-broadcast operands.  Records use the
// general layout {x,x',y,y'} {g0,g0',g1,g1'} {g2,g2',g3,g3'} {g4,g4',0,0}.  MASK: upstream rows
// present (compile-time pruned, s2_bwd_pair<MASK>).
// -----------------------------------------------------------------------------
template <int Q, int THREADS, int MINB, int TILE, int MASK>
__global__ void __launch_bounds__(THREADS, MINB)
s2k1_bwd_kernel(int N2, int pairs_per_chunk, const float4* __restrict__ ptrecs,
                const float* __restrict__ soa, int Mp, float* __restrict__ partials) {
  static_assert(Q % 2 == 0, "lanes are node pairs");
  constexpr int NREC = 4, NACC = 5, QS = Q / 2, NS = 2;
  __shared__ __align__(16) float4 ring[NS][NREC * TILE];
  __shared__ __align__(8) uint64_t full[NS];
  const int jblock = blockIdx.x * (Q * THREADS);
  float2 nc[QS], nd[QS], c1[QS], cr[QS], ncr[QS], nrho2[QS];
  float hq[Q], wq[Q];
#pragma unroll
  for (int q = 0; q < QS; ++q) {
    const int j = jblock + 2 * (q * THREADS + threadIdx.x);
    const float2 c = *reinterpret_cast<const float2*>(soa + j);
    const float2 d = *reinterpret_cast<const float2*>(soa + (size_t)Mp + j);
    const float2 h = *reinterpret_cast<const float2*>(soa + 2 * (size_t)Mp + j);
    const float2 w = *reinterpret_cast<const float2*>(soa + 4 * (size_t)Mp + j);
    hq[2 * q] = h.x; hq[2 * q + 1] = h.y;
    wq[2 * q] = w.x; wq[2 * q + 1] = w.y;
    const S2BwdNode n0 = s2_bwd_node(c.x, d.x, h.x), n1 = s2_bwd_node(c.y, d.y, h.y);
    nc[q] = f2(n0.nc, n1.nc); nd[q] = f2(n0.nd, n1.nd); c1[q] = f2(n0.c1, n1.c1);
    cr[q] = f2(n0.cr, n1.cr); ncr[q] = f2(n0.ncr, n1.ncr); nrho2[q] = f2(n0.nrho2, n1.nrho2);
  }
  float2 tot[QS][NACC];
#pragma unroll
  for (int q = 0; q < QS; ++q)
#pragma unroll
    for (int a = 0; a < NACC; ++a) tot[q][a] = f2(0.f, 0.f);

  const int p_begin = blockIdx.y * pairs_per_chunk;
  const int p_end = min(N2, p_begin + pairs_per_chunk);
  const int ntiles = p_end > p_begin ? (p_end - p_begin + TILE - 1) / TILE : 0;
  auto issue = [&](int t) {                    // thread 0 only
    const int t0 = p_begin + t * TILE;
    const unsigned bytes = (unsigned)(min(TILE, p_end - t0) * NREC) * (unsigned)sizeof(float4);
    mbar_expect_tx(&full[t % NS], bytes);
    tma_bulk_g2s(ring[t % NS], ptrecs + NREC * (size_t)t0, bytes, &full[t % NS]);
  };
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NS; ++i) mbar_init(&full[i], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int i = 0; i < NS - 1 && i < ntiles; ++i) issue(i);
  for (int t = 0; t < ntiles; ++t) {
    const int t0 = p_begin + t * TILE;
    const int n = min(TILE, p_end - t0);
    const float4* sm4 = ring[t % NS];
    if (threadIdx.x == 0 && t + NS - 1 < ntiles) issue(t + NS - 1);
    mbar_wait(&full[t % NS], (unsigned)((t / NS) & 1));
    float2 acc[QS][NACC];
#pragma unroll
    for (int q = 0; q < QS; ++q)
#pragma unroll
      for (int a = 0; a < NACC; ++a) acc[q][a] = f2(0.f, 0.f);
#pragma unroll 1
    for (int p = 0; p < n; ++p) {
      const float4 A = sm4[NREC * p + 0];
      float4 B = make_float4(0.f, 0.f, 0.f, 0.f), C = B, D = B;
      if (MASK & 0x03) B = sm4[NREC * p + 1];
      if (MASK & 0x0C) C = sm4[NREC * p + 2];
      if (MASK & 0x10) D = sm4[NREC * p + 3];
#pragma unroll
      for (int q = 0; q < QS; ++q) {
        s2_bwd_pair<MASK>(f2dup(A.x), f2dup(A.z), f2dup(B.x), f2dup(B.z), f2dup(C.x), f2dup(C.z), f2dup(D.x),
                          nc[q], nd[q], c1[q], cr[q], ncr[q], nrho2[q], acc[q]);
        s2_bwd_pair<MASK>(f2dup(A.y), f2dup(A.w), f2dup(B.y), f2dup(B.w), f2dup(C.y), f2dup(C.w), f2dup(D.y),
                          nc[q], nd[q], c1[q], cr[q], ncr[q], nrho2[q], acc[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < QS; ++q)
#pragma unroll
      for (int a = 0; a < NACC; ++a) { tot[q][a].x += acc[q][a].x; tot[q][a].y += acc[q][a].y; }
    __syncthreads();                                   // everyone is done reading this stage
  }
  // rows: 0 dc, 1 dd, 2 dh, 3 dW
  float* out = partials + (size_t)blockIdx.y * 4 * Mp;
#pragma unroll
  for (int q = 0; q < QS; ++q) {
    const int j = jblock + 2 * (q * THREADS + threadIdx.x);
    float t0[NACC], t1[NACC];
#pragma unroll
    for (int a = 0; a < NACC; ++a) { t0[a] = tot[q][a].x; t1[a] = tot[q][a].y; }
    float2 dW, dc, dd, dh;
    s2_bwd_finish(hq[2 * q], wq[2 * q], t0, &dW.x, &dc.x, &dd.x, &dh.x);
    s2_bwd_finish(hq[2 * q + 1], wq[2 * q + 1], t1, &dW.y, &dc.y, &dd.y, &dh.y);
    *reinterpret_cast<float2*>(out + j) = dc;
    *reinterpret_cast<float2*>(out + (size_t)Mp + j) = dd;
    *reinterpret_cast<float2*>(out + 2 * (size_t)Mp + j) = dh;
    *reinterpret_cast<float2*>(out + 3 * (size_t)Mp + j) = dW;
  }
}

// -----------------------------------------------------------------------------
// GENERIC forward: thread = one point, nodes tiled through shared memory (SoA).
// -----------------------------------------------------------------------------
constexpr int GEN_THREADS = 128;
constexpr int GEN_TILE_NODES = 512;
constexpr int GEN_TILE_POINTS = 256;

template <int DIM, int KIND, int K, int LEVEL>
__global__ void __launch_bounds__(GEN_THREADS)
fwd_generic_kernel(int N, int Mp, const float* __restrict__ x, const float* __restrict__ y,
                   const float* __restrict__ soa, const float* __restrict__ bias,
                   float* __restrict__ jets) {
  // thread = TWO consecutive points in the two lanes of packed f32x2 arithmetic (F2)
  constexpr int NJ = njets(DIM, LEVEL);
  constexpr int NROW = DIM + 1 + K;              // c,(d),r,W_k  (the h row is skipped)
  constexpr int TN = GEN_TILE_NODES;
  __shared__ float sm[NROW * TN];
  const long long i0 = 2LL * ((long long)blockIdx.x * blockDim.x + threadIdx.x);
  const bool v0 = i0 < N, v1 = i0 + 1 < N;
  const F2 px(v0 ? x[i0] : 0.f, v1 ? x[i0 + 1] : 0.f);
  const F2 py((DIM == 2 && v0) ? y[i0] : 0.f, (DIM == 2 && v1) ? y[i0 + 1] : 0.f);
  F2 acc[NJ][K];
#pragma unroll
  for (int a = 0; a < NJ; ++a)
#pragma unroll
    for (int k = 0; k < K; ++k) acc[a][k] = F2(0.f);

  for (int t0 = 0; t0 < Mp; t0 += TN) {
    const int n = min(TN, Mp - t0);
    __syncthreads();
    for (int e = threadIdx.x; e < NROW * n; e += blockDim.x) {
      const int row = e / n, col = e - row * n;
      const int srow = row < DIM ? row : row + 1;   // skip h
      sm[row * TN + col] = soa[(size_t)srow * Mp + t0 + col];
    }
    __syncthreads();
#pragma unroll 2
    for (int j = 0; j < n; ++j) {
      float Wk[K];
#pragma unroll
      for (int k = 0; k < K; ++k) Wk[k] = sm[(DIM + 1 + k) * TN + j];
      if (DIM == 2)
        generic_fwd_pair<DIM, KIND, K, LEVEL>(px - sm[j], py - sm[TN + j], sm[2 * TN + j], Wk, acc);
      else
        generic_fwd_pair<DIM, KIND, K, LEVEL>(px - sm[j], F2(0.f), sm[TN + j], Wk, acc);
    }
  }
#pragma unroll
  for (int a = 0; a < NJ; ++a)
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float b = (a == 0 && bias) ? bias[k] : 0.f;
      if (v0) jets[((size_t)a * N + i0) * K + k] = acc[a][k].v.x + b;
      if (v1) jets[((size_t)a * N + i0 + 1) * K + k] = acc[a][k].v.y + b;
    }
}

// -----------------------------------------------------------------------------
// GENERIC backward: thread = one node, points tiled through shared memory, two
// points per step in the lanes of packed arithmetic.
// partial rows: 0 dc, (1 dd), DIM dh, DIM+1+k dW_k.
// -----------------------------------------------------------------------------
template <int DIM, int KIND, int K, int LEVEL>
__global__ void __launch_bounds__(GEN_THREADS)
bwd_generic_kernel(int N, int Mp, int pts_per_chunk, const float* __restrict__ x,
                   const float* __restrict__ y, const float* __restrict__ gj, int mask,
                   const float* __restrict__ soa, float* __restrict__ partials) {
  constexpr int NG = njets(DIM, LEVEL);
  constexpr int TP = GEN_TILE_POINTS;            // even
  constexpr int NACC = DIM + 1 + K;
  __shared__ __align__(16) float spx[TP];
  __shared__ __align__(16) float spy[DIM == 2 ? TP : 2];
  __shared__ __align__(16) float sg[NG * K * TP];
  const int j = blockIdx.x * GEN_THREADS + threadIdx.x;     // < Mp (Mp multiple of GEN_THREADS)
  int row = 0;
  const float c = soa[(size_t)(row++) * Mp + j];
  const float dn = DIM == 2 ? soa[(size_t)(row++) * Mp + j] : 0.f;
  row++;  // h
  const float r = soa[(size_t)(row++) * Mp + j];
  float Wk[K];
#pragma unroll
  for (int k = 0; k < K; ++k) Wk[k] = soa[(size_t)(row++) * Mp + j];

  float tot[NACC];
#pragma unroll
  for (int a = 0; a < NACC; ++a) tot[a] = 0.f;

  const long long i_begin = (long long)blockIdx.y * pts_per_chunk;
  const long long i_end = min((long long)N, i_begin + pts_per_chunk);
  for (long long t0 = i_begin; t0 < i_end; t0 += TP) {
    const int n = (int)min((long long)TP, i_end - t0);
    const int n_even = (n + 1) & ~1;
    __syncthreads();
    for (int e = threadIdx.x; e < n_even; e += GEN_THREADS) {
      const bool v = e < n;                       // odd tail: a zero-gradient dummy point
      spx[e] = v ? x[t0 + e] : 0.f;
      if (DIM == 2) spy[e] = v ? y[t0 + e] : 0.f;
#pragma unroll
      for (int a = 0; a < NG; ++a)
#pragma unroll
        for (int k = 0; k < K; ++k)
          sg[(a * K + k) * TP + e] = (v && ((mask >> a) & 1)) ? gj[((size_t)a * N + t0 + e) * K + k] : 0.f;
    }
    __syncthreads();
    F2 acc[NACC];
#pragma unroll
    for (int a = 0; a < NACC; ++a) acc[a] = F2(0.f);
    for (int i = 0; i < n_even; i += 2) {
      const F2 X = F2(*reinterpret_cast<const float2*>(&spx[i])) - c;
      const F2 Y = DIM == 2 ? F2(*reinterpret_cast<const float2*>(&spy[DIM == 2 ? i : 0])) - dn : F2(0.f);
      F2 g[NG * K];
#pragma unroll
      for (int q = 0; q < NG * K; ++q) g[q] = F2(*reinterpret_cast<const float2*>(&sg[q * TP + i]));
      generic_bwd_pair<DIM, KIND, K, LEVEL>(X, Y, r, Wk, g, acc);
    }
#pragma unroll
    for (int a = 0; a < NACC; ++a) tot[a] += vlane_sum(acc[a]);
  }
  float* out = partials + (size_t)blockIdx.y * NACC * Mp;
#pragma unroll
  for (int a = 0; a < NACC; ++a) out[(size_t)a * Mp + j] = tot[a];
}

// -----------------------------------------------------------------------------
// second stage: bias gradient partials, chunk reduction, scalar-h reduction
// -----------------------------------------------------------------------------
constexpr int BIAS_BLOCKS = 1024;
constexpr int LOSS_BLOCKS = 8192;   // >= warps of the largest forward grid (2 CTAs x 16 warps x SMs)

__global__ void __launch_bounds__(256)
bias_partial_kernel(int N, int K, const float* __restrict__ g0, float* __restrict__ bpart) {
  __shared__ float red[SPINN_MAX_K][8];
  float s[SPINN_MAX_K] = {0.f, 0.f, 0.f, 0.f};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N;
       i += (long long)gridDim.x * blockDim.x)
    for (int k = 0; k < K; ++k) s[k] += g0[(size_t)i * K + k];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = 0; k < K; ++k) {
    float v = s[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[threadIdx.x][w];
    bpart[blockIdx.x * K + threadIdx.x] = v;
  }
}

// grid (ceil(M/32), nacc + 1), block (32, 8): thread (tx, ty) sums chunks ty, ty+8, ... of one
// (row, node) in fp64, the 8 partial sums are combined in a fixed order (deterministic); the
// extra grid row reduces the bias partials.
__global__ void __launch_bounds__(256)
finalize_kernel(int dim, int K, int M, int Mfree, int Mp, int C, const float* __restrict__ partials,
                const float* __restrict__ bpart, int nb, float* __restrict__ d_xc,
                float* __restrict__ d_yc, float* __restrict__ d_h, float* __restrict__ d_W,
                float* __restrict__ d_bias, const float* __restrict__ lpart, int nlp,
                float* __restrict__ loss) {
  __shared__ double red[8][33];
  const int nacc = dim + 1 + K;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int a = blockIdx.y;
  if (a == nacc + 1) {                   // loss row (fused interior step): loss = sum of block partials
    if (blockIdx.x != 0 || loss == nullptr) return;
    const int t = ty * 32 + tx;
    double s = 0.0;
    for (int b = t; b < nlp; b += 256) s += (double)lpart[b];
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0) {
      double v = 0.0;
      for (int r = 0; r < 8; ++r) v += red[r][tx];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (tx == 0) loss[0] = (float)v;
    }
    return;
  }
  if (a == nacc) {                       // bias row
    if (blockIdx.x != 0 || d_bias == nullptr) return;
    const int t = ty * 32 + tx;
    for (int k = 0; k < K; ++k) {
      double s = 0.0;
      for (int b = t; b < nb; b += 256) s += (double)bpart[b * K + k];
      __syncthreads();
      red[ty][tx] = s;
      __syncthreads();
      if (ty == 0) {
        double v = 0.0;
        for (int r = 0; r < 8; ++r) v += red[r][tx];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (tx == 0) d_bias[k] = (float)v;
      }
    }
    return;
  }
  const int j = blockIdx.x * 32 + tx;
  double s = 0.0;
  if (j < M)
    for (int c = ty; c < C; c += 8) s += (double)partials[((size_t)c * nacc + a) * Mp + j];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && j < M) {
    double v = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r) v += red[r][tx];
    const float f = (float)v;
    if (a == 0) { if (j < Mfree) d_xc[j] = f; }
    else if (dim == 2 && a == 1) { if (j < Mfree) d_yc[j] = f; }
    else if (a == dim) d_h[j] = f;
    else d_W[(size_t)(a - dim - 1) * M + j] = f;
  }
}

__global__ void __launch_bounds__(256)
sum_to_scalar_kernel(int n, const float* __restrict__ v, float* __restrict__ out) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += (double)v[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)red[0];
}

// -----------------------------------------------------------------------------
// Poisson residual epilogue (K=1): res, loss partials, upstream jets gradients
// -----------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
poisson_residual_kernel(int N, float inv_n, const float* __restrict__ x, const float* __restrict__ y,
                        const float* __restrict__ jets, float scale, float amp, float kx, float ky,
                        float f0, float* __restrict__ gj, float* __restrict__ lpart) {
  __shared__ float red[8];
  const size_t n = (size_t)N;
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N;
       i += (long long)gridDim.x * blockDim.x) {
    const float f = fmaf(amp * sinf(kx * x[i]), sinf(ky * y[i]), f0);
    const float res = scale * (jets[3 * n + i] + jets[4 * n + i] + f);
    s = fmaf(res, res, s);
    const float g = 2.f * scale * res * inv_n;
    gj[i] = 0.f; gj[n + i] = 0.f; gj[2 * n + i] = 0.f;
    gj[3 * n + i] = g; gj[4 * n + i] = g;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[w];
    lpart[blockIdx.x] = v * inv_n;
  }
}

// ode3 residual epilogue (1-D, K=1): jets [3][N] -> gjets row 2 (u_xx) only (rows 0, 1 are absent:
// present_mask 0x4), loss partials
__global__ void __launch_bounds__(256)
ode3_residual_kernel(int N, float inv_n, float Kc, const float* __restrict__ x, const float* __restrict__ jets,
                     float* __restrict__ gj, float* __restrict__ lpart) {
  __shared__ float red[8];
  const size_t n = (size_t)N;
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N;
       i += (long long)gridDim.x * blockDim.x) {
    const float res = jets[2 * n + i] + ode3_forcing(x[i], Kc);
    s = fmaf(res, res, s);
    gj[2 * n + i] = 2.f * res * inv_n;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[w];
    lpart[blockIdx.x] = v * inv_n;
  }
}

// cavity residual epilogue (2-D, K=3: u, v, p): jets [5][N][3] -> gjets [5][N][3], loss partials
__global__ void __launch_bounds__(256)
cavity_residual_kernel(int N, float nu, const float* __restrict__ jets, float* __restrict__ gj,
                       float* __restrict__ lpart) {
  __shared__ float red[8];
  const size_t n = (size_t)N;
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N;
       i += (long long)gridDim.x * blockDim.x) {
    float J[5][3], g[5][3];
#pragma unroll
    for (int a = 0; a < 5; ++a)
#pragma unroll
      for (int k = 0; k < 3; ++k) J[a][k] = jets[(a * n + i) * 3 + k];
    s += cavity_residual_point(J, nu, g);
#pragma unroll
    for (int a = 0; a < 5; ++a)
#pragma unroll
      for (int k = 0; k < 3; ++k) gj[(a * n + i) * 3 + k] = g[a][k];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[w];
    lpart[blockIdx.x] = v;
  }
}

__global__ void __launch_bounds__(256)
add_partials_kernel(int n, const float* __restrict__ part, float* __restrict__ out) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += (double)part[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] += (float)red[0];
}

// -----------------------------------------------------------------------------
// Fused interior step (stock Poisson residual): node packing for BOTH passes in one kernel (the
// residual epilogue itself runs inside the forward kernel: FwdEpilogue).
// -----------------------------------------------------------------------------
template <int KIND>
__global__ void pack_nodes_step_kernel(int M, int Mfree, int npairs, int Mp,
                                       const float* __restrict__ xf, const float* __restrict__ yf,
                                       const float* __restrict__ xfix, const float* __restrict__ yfix,
                                       const float* __restrict__ h, int hs, const float* __restrict__ W,
                                       float4* __restrict__ recs, float* __restrict__ soa) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  auto node = [&](int j, float& c, float& d, float& hh, float& w) {
    c = 0.f; d = 0.f; hh = 1.f; w = 0.f;
    if (j < M) {
      c = j < Mfree ? xf[j] : xfix[j - Mfree];
      d = j < Mfree ? yf[j] : yfix[j - Mfree];
      hh = hs ? h[0] : h[j];
      w = W[j];
    }
  };
  if (t < npairs) {                      // forward record of node pair t
    float v[2][6];
#pragma unroll
    for (int l = 0; l < 2; ++l) {
      float c, d, hh, w;
      node(2 * t + l, c, d, hh, w);
      if (KIND == GAUSSIAN) {
        const G2FwdNode n = g2_fwd_node(c, d, hh, w);
        v[l][0] = n.nc; v[l][1] = n.nd; v[l][2] = n.rp; v[l][3] = n.w0; v[l][4] = n.w1; v[l][5] = n.w2;
      } else {
        const S2FwdNode n = s2_fwd_node(c, d, hh, w);
        v[l][0] = n.nc; v[l][1] = n.nd; v[l][2] = n.c1; v[l][3] = n.w0; v[l][4] = n.w1; v[l][5] = n.w2;
      }
    }
    recs[3 * (size_t)t + 0] = make_float4(v[0][0], v[1][0], v[0][1], v[1][1]);
    recs[3 * (size_t)t + 1] = make_float4(v[0][2], v[1][2], v[0][3], v[1][3]);
    recs[3 * (size_t)t + 2] = make_float4(v[0][4], v[1][4], v[0][5], v[1][5]);
  }
  if (t < Mp) {                          // backward SoA column t: c, d, h, 1/h, W
    float c, d, hh, w;
    node(t, c, d, hh, w);
    soa[t] = c;
    soa[(size_t)Mp + t] = d;
    soa[2 * (size_t)Mp + t] = hh;
    soa[3 * (size_t)Mp + t] = 1.0f / hh;
    soa[4 * (size_t)Mp + t] = w;
  }
}

// -----------------------------------------------------------------------------
// fast forward launcher (one instantiation per kernel kind / tile shape)
// -----------------------------------------------------------------------------
struct FastFwdArgs {
  int N; const float* x; const float* y; const float4* recs; int npairs, tile_pairs;
  const float* bias; float* jets; int vec_ok; size_t smem; int sms, dev; cudaStream_t s;
  const FwdEpilogue* epi;      // non-null: run the residual epilogue (level 2 only)
  int* nwarps_out;             // EPI: number of loss partials written (= warps of the grid)
};

template <int KIND, int P, int T, int B, int U, bool MIXED, bool EPI, bool MULTI>
static int launch_fast_fwd_k(const FastFwdArgs& a) {
  const long long nwt = ((long long)a.N + 32 * P - 1) / (32 * P);
  long long blocks = (nwt + (T / 32) - 1) / (T / 32);
  if (blocks > (long long)a.sms * B) blocks = (long long)a.sms * B;
  const FwdEpilogue ep = EPI ? *a.epi : FwdEpilogue{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, nullptr, nullptr};
  if (EPI && a.nwarps_out) *a.nwarps_out = (int)blocks * (T / 32);
  static OnceFlags once;
  if (KIND == GAUSSIAN) {
    auto kfn = g2k1_fwd_kernel<P, MIXED, T, B, U, EPI, MULTI>;
    CUDA_TRY(opt_in_smem_once(once, a.dev, kfn, 98304));
    kfn<<<(int)blocks, T, a.smem, a.s>>>(a.N, a.x, a.y, a.recs, a.npairs, a.tile_pairs, a.bias, a.jets, a.vec_ok, ep);
  } else {
    auto kfn = s2k1_fwd_kernel<P, MIXED, T, B, U, EPI, MULTI>;
    CUDA_TRY(opt_in_smem_once(once, a.dev, kfn, 98304));
    kfn<<<(int)blocks, T, a.smem, a.s>>>(a.N, a.x, a.y, a.recs, a.npairs, a.tile_pairs, a.bias, a.jets, a.vec_ok, ep);
  }
  LAUNCH_CHECK(KIND == GAUSSIAN ? "g2k1_fwd_kernel" : "s2k1_fwd_kernel");
  return SPINN_OK;
}
template <int KIND, int P, int T, int B, int U, bool MIXED, bool EPI>
static int launch_fast_fwd_e(const FastFwdArgs& a) {
  // the multi-tile kernel exists for the default tile shape only (node sets > 4096 are rare)
  if (a.npairs > a.tile_pairs) return launch_fast_fwd_k<KIND, 2, 256, 2, 2, MIXED, EPI, true>(a);
  return launch_fast_fwd_k<KIND, P, T, B, U, MIXED, EPI, false>(a);
}
template <int KIND, int P, int T, int B, int U>
static int launch_fast_fwd(bool mixed, const FastFwdArgs& a) {
  if (a.epi) return launch_fast_fwd_e<KIND, P, T, B, U, false, true>(a);
  return mixed ? launch_fast_fwd_e<KIND, P, T, B, U, true, false>(a) : launch_fast_fwd_e<KIND, P, T, B, U, false, false>(a);
}

// -----------------------------------------------------------------------------
// planning
// -----------------------------------------------------------------------------
struct FwdPlan {
  bool fast;
  int Mp;          // generic: padded node count (multiple of GEN_THREADS)
  int npairs;      // fast: node pairs
  size_t ws_bytes;
};

static bool fwd_is_fast(int dim, int kind, int level, int K) {
  return dim == 2 && K == 1 && (level == 2 || level == 3);
}

static FwdPlan plan_fwd(int dim, int kind, int level, int M, int K) {
  FwdPlan p;
  p.fast = fwd_is_fast(dim, kind, level, K);
  p.Mp = round_up_i(M, GEN_THREADS);
  p.npairs = (M + 1) / 2;
  const size_t gen = (size_t)(dim + 2 + K) * p.Mp * sizeof(float);
  const size_t fast = (size_t)3 * p.npairs * sizeof(float4);
  p.ws_bytes = align_up(gen > fast ? gen : fast, 256);
  return p;
}

static size_t fwd_ws_bytes_any(int dim, int M, int K) {
  // upper bound over kind/level so that the caller can size once per (M, K)
  FwdPlan a = plan_fwd(dim, SPINN_KIND_GAUSSIAN, 2, M, K);
  return a.ws_bytes;
}

struct BwdPlan {
  bool fast;
  int Mp;              // padded node count
  int node_groups;
  int chunks;          // point chunks (grid.y)
  int per_chunk;       // points (generic) or point pairs (fast) per chunk
  int N2;              // fast: point pairs
  int variant;         // fast: BWD_CFGS index
  size_t off_soa, off_pts, off_part, off_bias, off_htmp, off_loss, ws_bytes;
};

static bool bwd_is_fast(int dim, int kind, int level, int K) {
  (void)kind;
  return dim == 2 && K == 1 && level <= 2;
}

// effective upstream mask: bit a set <=> row a of gjets is present; 0 on input means "all"
static int effective_mask(int dim, int level, int present_mask) {
  const int all = (1 << njets(dim, level)) - 1;
  return present_mask == 0 ? all : (present_mask & all);
}

static int bwd_mode_for(int mask) {
  if (mask == 0x01) return BWD_U;
  if ((mask & ~0x18) == 0) return BWD_LAPL;
  // residual patterns of the packaged space-time problems: advection (u_x,u_t), heat (u_t,u_xx),
  // Burgers (u,u_x,u_t,u_xx)
  if (mask == 0x06 || mask == 0x0C || mask == 0x0F) return BWD_MASKED + mask;
  return BWD_ALL;
}

// chunking: enough (node group x point chunk) blocks for `waves` waves of resident CTAs
static int chunks_for(int sms, int waves, int groups, long long units, int tile, int blocks_per_sm,
                      int* per_chunk) {
  long long target = (long long)sms * blocks_per_sm * waves;
  int c = (int)((target + groups - 1) / groups);
  if (c < 1) c = 1;
  if (c > 65535) c = 65535;                    // grid.y limit
  // chunk length: whole smem tiles for large point sets; for small ones a finer granularity
  // (the tile loop handles partial tiles) so that a few hundred points still spread over many CTAs
  long long per = (units + c - 1) / c;
  const int gran = per >= 4LL * tile ? tile : 32;
  per = (per + gran - 1) / gran * gran;
  if (per < gran) per = gran;
  *per_chunk = (int)per;
  c = (int)((units + per - 1) / per);
  return c < 1 ? 1 : c;
}

static BwdPlan plan_bwd(int dim, int kind, int level, int N, int M, int K, bool worst, int mode = BWD_ALL) {
  BwdPlan p;
  p.fast = bwd_is_fast(dim, kind, level, K);
  const Tune tune = get_tune();
  const int sms = sm_count();
  const int nacc = dim + 1 + K;
  p.N2 = (N + 1) / 2;
  p.variant = tune.bwd;
  // generic path
  const int Mp_gen = round_up_i(M, GEN_THREADS);
  int pc_gen = 0;
  const int c_gen = chunks_for(sms, 2, Mp_gen / GEN_THREADS, N, GEN_TILE_POINTS, 4, &pc_gen);
  // fast path: the node tile of a CTA is Q*threads nodes; pick the variant that wastes the fewest
  // padded node slots (relative kernel cost measured with tools/tune.py: 512-node tile 1.00,
  // 256-node tile 1.015, 128-node tile 1.06, 64-node tile ~1.10), unless SPINN_TUNE pins one
  // The structured modes also have a 1024-node tile (two node-pair streams per thread, 3 % faster
  // than the 512-node tile at M = 4096).
  const bool sp = kind == SPINN_KIND_SOFTPLUS;
  if (sp) {
    const double cost[N_S2_BWD_CFGS] = {1.0, 1.015, 1.06, 1.10};
    double best = 1e300;
    p.variant = 0;
    for (int i = 0; i < N_S2_BWD_CFGS; ++i) {
      const BwdCfg c = S2_BWD_CFGS[i];
      const double w = (double)round_up_i(M, c.q * c.threads) * cost[i];
      if (w < best) { best = w; p.variant = i; }
    }
  } else if (!tune.bwd_forced) {
    const int cand[5] = {2, 0, 1, 4, 5};
    const double cost[5] = {1.0, 1.015, 1.06, 1.10, 0.97};
    double best = 1e300;
    for (int i = 0; i < ((mode == BWD_LAPL || mode == BWD_U) ? 5 : 4); ++i) {
      const BwdCfg c = BWD_CFGS[cand[i]];
      const double w = (double)round_up_i(M, c.q * c.threads) * cost[i];
      if (w < best) { best = w; p.variant = cand[i]; }
    }
  }
  const BwdCfg cfg = sp ? S2_BWD_CFGS[p.variant] : BWD_CFGS[p.variant];
  const int group = cfg.q * cfg.threads;
  const int Mp_fast = round_up_i(M, group);
  int pc_fast = 0;
  const int c_fast = chunks_for(sms, tune.waves, Mp_fast / group, p.N2, cfg.tile, cfg.minb, &pc_fast);
  if (p.fast) { p.Mp = Mp_fast; p.node_groups = Mp_fast / group; p.chunks = c_fast; p.per_chunk = pc_fast; }
  else { p.Mp = Mp_gen; p.node_groups = Mp_gen / GEN_THREADS; p.chunks = c_gen; p.per_chunk = pc_gen; }
  size_t part_floats = (size_t)p.chunks * nacc * p.Mp;
  size_t Mp_w = p.Mp;
  if (worst) {
    part_floats = (size_t)c_gen * nacc * Mp_gen;
    Mp_w = Mp_gen;
    for (int v = 0; v < N_BWD_CFGS + N_S2_BWD_CFGS; ++v) {
      const BwdCfg c = v < N_BWD_CFGS ? BWD_CFGS[v] : S2_BWD_CFGS[v - N_BWD_CFGS];
      const int g = c.q * c.threads, mp = round_up_i(M, g);
      int pc = 0;
      const int waves = tune.waves > 8 ? tune.waves : 8;
      const int ch = chunks_for(sms, waves, mp / g, p.N2, c.tile, c.minb, &pc);
      const size_t f = (size_t)ch * nacc * mp;
      if (f > part_floats) part_floats = f;
      if ((size_t)mp > Mp_w) Mp_w = mp;
    }
  }
  size_t off = 0;
  p.off_soa = off;  off += align_up((size_t)(dim + 2 + K) * Mp_w * sizeof(float), 256);
  p.off_pts = off;  off += (worst || p.fast) ? align_up((size_t)4 * p.N2 * sizeof(float4), 256) : 0;
  p.off_part = off; off += align_up(part_floats * sizeof(float), 256);
  p.off_bias = off; off += align_up((size_t)BIAS_BLOCKS * SPINN_MAX_K * sizeof(float), 256);
  p.off_htmp = off; off += align_up(Mp_w * sizeof(float), 256);
  p.off_loss = off; off += align_up((size_t)LOSS_BLOCKS * sizeof(float), 256);
  p.ws_bytes = off;
  return p;
}

// -----------------------------------------------------------------------------
// dispatch helpers
// -----------------------------------------------------------------------------
template <int DIM, int KIND, int K>
static void launch_fwd_generic_level(int level, int N, int Mp, const float* x, const float* y,
                                     const float* soa, const float* bias, float* jets,
                                     cudaStream_t s) {
  dim3 grid(ceil_div_i(N, 2 * GEN_THREADS)), block(GEN_THREADS);
  switch (level) {
    case 0: fwd_generic_kernel<DIM, KIND, K, 0><<<grid, block, 0, s>>>(N, Mp, x, y, soa, bias, jets); break;
    case 1: fwd_generic_kernel<DIM, KIND, K, 1><<<grid, block, 0, s>>>(N, Mp, x, y, soa, bias, jets); break;
    case 2: fwd_generic_kernel<DIM, KIND, K, 2><<<grid, block, 0, s>>>(N, Mp, x, y, soa, bias, jets); break;
    default:
      if (DIM == 2)
        fwd_generic_kernel<DIM, KIND, K, (DIM == 2 ? 3 : 2)><<<grid, block, 0, s>>>(N, Mp, x, y, soa, bias, jets);
      break;
  }
}

template <int DIM, int KIND>
static void launch_fwd_generic_k(int K, int level, int N, int Mp, const float* x, const float* y,
                                 const float* soa, const float* bias, float* jets, cudaStream_t s) {
  switch (K) {
    case 1: launch_fwd_generic_level<DIM, KIND, 1>(level, N, Mp, x, y, soa, bias, jets, s); break;
    case 2: launch_fwd_generic_level<DIM, KIND, 2>(level, N, Mp, x, y, soa, bias, jets, s); break;
    case 3: launch_fwd_generic_level<DIM, KIND, 3>(level, N, Mp, x, y, soa, bias, jets, s); break;
    default: launch_fwd_generic_level<DIM, KIND, 4>(level, N, Mp, x, y, soa, bias, jets, s); break;
  }
}

template <int DIM, int KIND, int K>
static void launch_bwd_generic_level(int level, dim3 grid, int N, int Mp, int per_chunk,
                                     const float* x, const float* y, const float* gj, int mask,
                                     const float* soa, float* partials, cudaStream_t s) {
  dim3 block(GEN_THREADS);
  switch (level) {
    case 0: bwd_generic_kernel<DIM, KIND, K, 0><<<grid, block, 0, s>>>(N, Mp, per_chunk, x, y, gj, mask, soa, partials); break;
    case 1: bwd_generic_kernel<DIM, KIND, K, 1><<<grid, block, 0, s>>>(N, Mp, per_chunk, x, y, gj, mask, soa, partials); break;
    case 2: bwd_generic_kernel<DIM, KIND, K, 2><<<grid, block, 0, s>>>(N, Mp, per_chunk, x, y, gj, mask, soa, partials); break;
    default:
      if (DIM == 2)
        bwd_generic_kernel<DIM, KIND, K, (DIM == 2 ? 3 : 2)><<<grid, block, 0, s>>>(N, Mp, per_chunk, x, y, gj, mask, soa, partials);
      break;
  }
}

template <int DIM, int KIND>
static void launch_bwd_generic_k(int K, int level, dim3 grid, int N, int Mp, int per_chunk,
                                 const float* x, const float* y, const float* gj, int mask,
                                 const float* soa, float* partials, cudaStream_t s) {
  switch (K) {
    case 1: launch_bwd_generic_level<DIM, KIND, 1>(level, grid, N, Mp, per_chunk, x, y, gj, mask, soa, partials, s); break;
    case 2: launch_bwd_generic_level<DIM, KIND, 2>(level, grid, N, Mp, per_chunk, x, y, gj, mask, soa, partials, s); break;
    case 3: launch_bwd_generic_level<DIM, KIND, 3>(level, grid, N, Mp, per_chunk, x, y, gj, mask, soa, partials, s); break;
    default: launch_bwd_generic_level<DIM, KIND, 4>(level, grid, N, Mp, per_chunk, x, y, gj, mask, soa, partials, s); break;
  }
}

static int check_common(int dim, int kind, int level, int N, int M_free, int M_fixed, int K) {
  if (kind != SPINN_KIND_GAUSSIAN && kind != SPINN_KIND_SOFTPLUS)
    return fail(SPINN_E_BADARG, "kind must be 0 (gaussian) or 1 (softplus-hat), got %d", kind);
  if (spinn_num_jets(dim, level) < 0)
    return fail(SPINN_E_BADARG, "invalid level %d for dim %d", level, dim);
  if (N < 0 || M_free < 0 || M_fixed < 0 || M_free + M_fixed <= 0)
    return fail(SPINN_E_BADARG, "bad sizes N=%d M_free=%d M_fixed=%d", N, M_free, M_fixed);
  if (K < 1) return fail(SPINN_E_BADARG, "K must be >= 1, got %d", K);
  if (K > SPINN_MAX_K)
    return fail(SPINN_E_UNSUPPORTED, "K=%d outputs not supported (max %d)", K, SPINN_MAX_K);
  return SPINN_OK;
}

// -----------------------------------------------------------------------------
// forward implementation shared by the 1-D / 2-D entry points
// -----------------------------------------------------------------------------
static int jets_fwd_impl(int dim, int kind, int level, int N, int M_free, int M_fixed, int K,
                         const float* x, const float* y, const float* xf, const float* yf,
                         const float* xfix, const float* yfix, const float* h, int hs,
                         const float* W, const float* bias, float* jets, void* ws, size_t ws_bytes,
                         void* stream, bool nodes_packed = false, const FwdEpilogue* epi = nullptr,
                         int* nwarps_out = nullptr) {
  int rc = check_common(dim, kind, level, N, M_free, M_fixed, K);
  if (rc) return rc;
  if (N == 0) return SPINN_OK;
  const int M = M_free + M_fixed;
  if (!x || (dim == 2 && !y) || !h || !W || !jets || (M_free && !xf) || (M_fixed && !xfix) ||
      (dim == 2 && ((M_free && !yf) || (M_fixed && !yfix))))
    return fail(SPINN_E_BADARG, "null pointer argument in jets_fwd");
  cudaStream_t s = (cudaStream_t)stream;
  const FwdPlan p = plan_fwd(dim, kind, level, M, K);
  if (!ws || ws_bytes < p.ws_bytes)
    return fail(SPINN_E_WORKSPACE, "forward workspace too small: %zu < %zu", ws_bytes, p.ws_bytes);

  if (p.fast) {
    float4* recs = (float4*)ws;
    if (nodes_packed) {
    } else if (kind == SPINN_KIND_GAUSSIAN)
      pack_nodes_k1fwd_kernel<GAUSSIAN><<<ceil_div_i(p.npairs, 128), 128, 0, s>>>(M, M_free, p.npairs, xf, yf,
                                                                                  xfix, yfix, h, hs, W, recs);
    else
      pack_nodes_k1fwd_kernel<SOFTPLUS><<<ceil_div_i(p.npairs, 128), 128, 0, s>>>(M, M_free, p.npairs, xf, yf,
                                                                                  xfix, yfix, h, hs, W, recs);
    LAUNCH_CHECK("pack_nodes_k1fwd_kernel");
    // node table: one stage when it fits 96 KB of shared memory (<= 2048 node pairs), else a 2-stage
    // ring of 1024-pair tiles (same 96 KB)
    const bool one_stage = p.npairs <= 2048;
    const int tile_pairs = one_stage ? p.npairs : 1024;
    const size_t smem = (size_t)3 * tile_pairs * sizeof(float4) * (one_stage ? 1 : 2);
    const int vec_ok = (((uintptr_t)x | (uintptr_t)y | (uintptr_t)jets) % 16 == 0) && (N % 4 == 0);
    const bool mixed = level == 3;
    const Tune tune = get_tune();
    const int sms = sm_count();
    int cur_dev = 0;
    if (cudaGetDevice(&cur_dev) != cudaSuccess) cur_dev = -1;
    // small point sets: one point per thread so that every SM gets work
    const bool small = (long long)N < (long long)32 * 4 * 8 * sms;
    const int v = small ? 2 : tune.fwd;
    const bool sp = kind == SPINN_KIND_SOFTPLUS;
    const FastFwdArgs fa = {N, x, y, recs, p.npairs, tile_pairs, bias, jets, vec_ok, smem, sms, cur_dev, s,
                            epi, nwarps_out};
    int rcl;
    switch (v) {
      case 0: rcl = sp ? launch_fast_fwd<SOFTPLUS, 4, 256, 2, 1>(mixed, fa) : launch_fast_fwd<GAUSSIAN, 4, 256, 2, 2>(mixed, fa); break;
      case 2: rcl = sp ? launch_fast_fwd<SOFTPLUS, 1, 512, 2, 2>(mixed, fa) : launch_fast_fwd<GAUSSIAN, 1, 512, 2, 4>(mixed, fa); break;
      default: rcl = sp ? launch_fast_fwd<SOFTPLUS, 2, 256, 2, 2>(mixed, fa) : launch_fast_fwd<GAUSSIAN, 2, 256, 2, 4>(mixed, fa); break;
    }
    if (rcl) return rcl;
    LAUNCH_CHECK("g2k1_fwd_kernel");
    return SPINN_OK;
  }

  float* soa = (float*)ws;
  if (dim == 2)
    pack_nodes_soa_kernel<2><<<ceil_div_i(p.Mp, 128), 128, 0, s>>>(M, M_free, p.Mp, K, xf, yf, xfix, yfix, h, hs, W, soa);
  else
    pack_nodes_soa_kernel<1><<<ceil_div_i(p.Mp, 128), 128, 0, s>>>(M, M_free, p.Mp, K, xf, nullptr, xfix, nullptr, h, hs, W, soa);
  LAUNCH_CHECK("pack_nodes_soa_kernel");
  if (dim == 2) {
    if (kind == SPINN_KIND_GAUSSIAN) launch_fwd_generic_k<2, GAUSSIAN>(K, level, N, p.Mp, x, y, soa, bias, jets, s);
    else launch_fwd_generic_k<2, SOFTPLUS>(K, level, N, p.Mp, x, y, soa, bias, jets, s);
  } else {
    if (kind == SPINN_KIND_GAUSSIAN) launch_fwd_generic_k<1, GAUSSIAN>(K, level, N, p.Mp, x, nullptr, soa, bias, jets, s);
    else launch_fwd_generic_k<1, SOFTPLUS>(K, level, N, p.Mp, x, nullptr, soa, bias, jets, s);
  }
  LAUNCH_CHECK("fwd_generic_kernel");
  return SPINN_OK;
}

// fused interior step: nodes and point records are already in the workspace, and the second stage
// also reduces the loss partials
struct BwdFused { int nlp; float* loss; };

static int jets_bwd_impl(int dim, int kind, int level, int present_mask, int N, int M_free,
                         int M_fixed, int K, const float* x, const float* y, const float* xf, const float* yf,
                         const float* xfix, const float* yfix, const float* h, int hs,
                         const float* W, const float* gj, float* d_xc, float* d_yc, float* d_h,
                         float* d_W, float* d_bias, void* ws, size_t ws_bytes, void* stream,
                         const BwdFused* fused = nullptr) {
  int rc = check_common(dim, kind, level, N, M_free, M_fixed, K);
  if (rc) return rc;
  const int M = M_free + M_fixed;
  if (!h || !W || !d_h || !d_W || (M_free && (!xf || !d_xc)) || (M_fixed && !xfix) ||
      (dim == 2 && ((M_free && (!yf || !d_yc)) || (M_fixed && !yfix))))
    return fail(SPINN_E_BADARG, "null pointer argument in jets_bwd");
  if (N > 0 && !fused && (!x || (dim == 2 && !y) || !gj))
    return fail(SPINN_E_BADARG, "null point/gradient pointer in jets_bwd");
  cudaStream_t s = (cudaStream_t)stream;
  if (present_mask < 0) return fail(SPINN_E_BADARG, "present_mask must be >= 0");
  const BwdPlan p = plan_bwd(dim, kind, level, N, M, K, false, bwd_mode_for(effective_mask(dim, level, present_mask)));
  if (!ws || ws_bytes < p.ws_bytes)
    return fail(SPINN_E_WORKSPACE, "backward workspace too small: %zu < %zu", ws_bytes, p.ws_bytes);
  char* base = (char*)ws;
  float* soa = (float*)(base + p.off_soa);
  float* partials = (float*)(base + p.off_part);
  float* bpart = (float*)(base + p.off_bias);
  float* htmp = (float*)(base + p.off_htmp);
  const int nacc = dim + 1 + K;
  if (present_mask < 0) return fail(SPINN_E_BADARG, "present_mask must be >= 0");
  const int mask = effective_mask(dim, level, present_mask);

  if (N == 0 || mask == 0) {  // empty point set or no upstream gradient: all gradients are zero
    if (M_free) CUDA_TRY(cudaMemsetAsync(d_xc, 0, sizeof(float) * M_free, s));
    if (dim == 2 && M_free) CUDA_TRY(cudaMemsetAsync(d_yc, 0, sizeof(float) * M_free, s));
    CUDA_TRY(cudaMemsetAsync(d_h, 0, sizeof(float) * (hs ? 1 : M), s));
    CUDA_TRY(cudaMemsetAsync(d_W, 0, sizeof(float) * (size_t)K * M, s));
    if (d_bias) CUDA_TRY(cudaMemsetAsync(d_bias, 0, sizeof(float) * K, s));
    return SPINN_OK;
  }

  if (fused) {
  } else if (dim == 2)
    pack_nodes_soa_kernel<2><<<ceil_div_i(p.Mp, 128), 128, 0, s>>>(M, M_free, p.Mp, K, xf, yf, xfix, yfix, h, hs, W, soa);
  else
    pack_nodes_soa_kernel<1><<<ceil_div_i(p.Mp, 128), 128, 0, s>>>(M, M_free, p.Mp, K, xf, nullptr, xfix, nullptr, h, hs, W, soa);
  LAUNCH_CHECK("pack_nodes_soa_kernel");

  dim3 grid(p.node_groups, p.chunks);
  if (p.fast && kind == SPINN_KIND_SOFTPLUS) {
    float4* ptrecs = (float4*)(base + p.off_pts);
    const int pgrid = ceil_div_i(p.N2, 256);
    if (!fused) pack_points_g2bwd_kernel<BWD_ALL><<<pgrid, 256, 0, s>>>(N, p.N2, x, y, gj, mask, ptrecs);
    LAUNCH_CHECK("pack_points_g2bwd_kernel");
    const int sm = (mask == 0x01) ? 0x01 : ((mask & ~0x18) == 0 ? 0x18 : 0x1F);
#define SPINN_LAUNCH_S2_BWD(Q_, T_, B_, TILE_)                                                                 \
  do {                                                                                                        \
    if (sm == 0x01)                                                                                           \
      s2k1_bwd_kernel<Q_, T_, B_, TILE_, 0x01><<<grid, T_, 0, s>>>(p.N2, p.per_chunk, ptrecs, soa, p.Mp, partials); \
    else if (sm == 0x18)                                                                                      \
      s2k1_bwd_kernel<Q_, T_, B_, TILE_, 0x18><<<grid, T_, 0, s>>>(p.N2, p.per_chunk, ptrecs, soa, p.Mp, partials); \
    else                                                                                                      \
      s2k1_bwd_kernel<Q_, T_, B_, TILE_, 0x1F><<<grid, T_, 0, s>>>(p.N2, p.per_chunk, ptrecs, soa, p.Mp, partials); \
  } while (0)
    switch (p.variant) {
      case 1: SPINN_LAUNCH_S2_BWD(2, 128, 4, 256); break;
      case 2: SPINN_LAUNCH_S2_BWD(2, 64, 8, 256); break;
      case 3: SPINN_LAUNCH_S2_BWD(2, 32, 16, 256); break;
      default: SPINN_LAUNCH_S2_BWD(2, 256, 2, 256); break;
    }
#undef SPINN_LAUNCH_S2_BWD
    LAUNCH_CHECK("s2k1_bwd_kernel");
  } else if (p.fast) {
    float4* ptrecs = (float4*)(base + p.off_pts);
    const int mode = bwd_mode_for(mask);
    const int pgrid = ceil_div_i(p.N2, 256);
    if (fused) { }
    else if (bwd_general(mode)) pack_points_g2bwd_kernel<BWD_ALL><<<pgrid, 256, 0, s>>>(N, p.N2, x, y, gj, mask, ptrecs);
    else if (mode == BWD_LAPL) pack_points_g2bwd_kernel<BWD_LAPL><<<pgrid, 256, 0, s>>>(N, p.N2, x, y, gj, mask, ptrecs);
    else pack_points_g2bwd_kernel<BWD_U><<<pgrid, 256, 0, s>>>(N, p.N2, x, y, gj, mask, ptrecs);
    LAUNCH_CHECK("pack_points_g2bwd_kernel");
#define SPINN_LAUNCH_FAST_BWD(Q_, T_, B_, TILE_)                                                              \
  do {                                                                                                        \
    if (mode == BWD_ALL)                                                                                      \
      g2k1_bwd_kernel<Q_, T_, B_, TILE_, BWD_ALL><<<grid, T_, 0, s>>>(p.N2, p.per_chunk, ptrecs, soa, p.Mp, partials);  \
    else if (mode == BWD_LAPL)                                                                                \
      g2k1_bwd_kernel<Q_, T_, B_, TILE_, BWD_LAPL><<<grid, T_, 0, s>>>(p.N2, p.per_chunk, ptrecs, soa, p.Mp, partials); \
    else if (mode == BWD_MASKED + 0x06)                                                                       \
      g2k1_bwd_kernel<Q_, T_, B_, TILE_, BWD_MASKED + 0x06><<<grid, T_, 0, s>>>(p.N2, p.per_chunk, ptrecs, soa, p.Mp, partials); \
    else if (mode == BWD_MASKED + 0x0C)                                                                       \
      g2k1_bwd_kernel<Q_, T_, B_, TILE_, BWD_MASKED + 0x0C><<<grid, T_, 0, s>>>(p.N2, p.per_chunk, ptrecs, soa, p.Mp, partials); \
    else if (mode == BWD_MASKED + 0x0F)                                                                       \
      g2k1_bwd_kernel<Q_, T_, B_, TILE_, BWD_MASKED + 0x0F><<<grid, T_, 0, s>>>(p.N2, p.per_chunk, ptrecs, soa, p.Mp, partials); \
    else                                                                                                      \
      g2k1_bwd_kernel<Q_, T_, B_, TILE_, BWD_U><<<grid, T_, 0, s>>>(p.N2, p.per_chunk, ptrecs, soa, p.Mp, partials);    \
  } while (0)
    switch (p.variant) {
      case 0: SPINN_LAUNCH_FAST_BWD(2, 128, 4, 256); break;
      case 1: SPINN_LAUNCH_FAST_BWD(1, 128, 8, 256); break;
      case 3: SPINN_LAUNCH_FAST_BWD(4, 128, 2, 256); break;
      case 4: SPINN_LAUNCH_FAST_BWD(1, 64, 8, 256); break;
      case 5: SPINN_LAUNCH_FAST_BWD(4, 256, 2, 256); break;
      case 6: SPINN_LAUNCH_FAST_BWD(2, 256, 4, 256); break;
      case 7: SPINN_LAUNCH_FAST_BWD(4, 128, 3, 256); break;
      default: SPINN_LAUNCH_FAST_BWD(2, 256, 2, 256); break;
    }
#undef SPINN_LAUNCH_FAST_BWD
    LAUNCH_CHECK("g2k1_bwd_kernel");
  } else {
    if (dim == 2) {
      if (kind == SPINN_KIND_GAUSSIAN) launch_bwd_generic_k<2, GAUSSIAN>(K, level, grid, N, p.Mp, p.per_chunk, x, y, gj, mask, soa, partials, s);
      else launch_bwd_generic_k<2, SOFTPLUS>(K, level, grid, N, p.Mp, p.per_chunk, x, y, gj, mask, soa, partials, s);
    } else {
      if (kind == SPINN_KIND_GAUSSIAN) launch_bwd_generic_k<1, GAUSSIAN>(K, level, grid, N, p.Mp, p.per_chunk, x, nullptr, gj, mask, soa, partials, s);
      else launch_bwd_generic_k<1, SOFTPLUS>(K, level, grid, N, p.Mp, p.per_chunk, x, nullptr, gj, mask, soa, partials, s);
    }
    LAUNCH_CHECK("bwd_generic_kernel");
  }
  int nb = 0;
  if (d_bias && (mask & 1)) {
    nb = ceil_div_i(N, 256) < BIAS_BLOCKS ? ceil_div_i(N, 256) : BIAS_BLOCKS;
    bias_partial_kernel<<<nb, 256, 0, s>>>(N, K, gj, bpart);
    LAUNCH_CHECK("bias_partial_kernel");
  }
  finalize_kernel<<<dim3(ceil_div_i(M, 32), nacc + (fused ? 2 : 1)), dim3(32, 8), 0, s>>>(
      dim, K, M, M_free, p.Mp, p.chunks, partials, bpart, nb, d_xc, d_yc, hs ? htmp : d_h, d_W, d_bias,
      fused ? (const float*)(base + p.off_loss) : nullptr, fused ? fused->nlp : 0, fused ? fused->loss : nullptr);
  LAUNCH_CHECK("finalize_kernel");
  if (hs) {
    sum_to_scalar_kernel<<<1, 256, 0, s>>>(M, htmp, d_h);
    LAUNCH_CHECK("sum_to_scalar_kernel");
  }
  (void)nacc;
  return SPINN_OK;
}

// -----------------------------------------------------------------------------
// C ABI
// -----------------------------------------------------------------------------
extern "C" {

int spinn_abi_version(void) { return SPINN_ABI_VERSION; }
void spinn_set_tuning(const char* spec) { parse_tuning(spec); }
const char* spinn_last_error(void) { return g_err; }

int spinn_num_jets(int dim, int level) {
  if (dim == 2 && level >= 0 && level <= 3) return njets(2, level);
  if (dim == 1 && level >= 0 && level <= 2) return njets(1, level);
  return -1;
}

size_t spinn2d_fwd_workspace_bytes(int N, int M, int K) { (void)N; return fwd_ws_bytes_any(2, M, K); }
size_t spinn1d_fwd_workspace_bytes(int N, int M, int K) { (void)N; return fwd_ws_bytes_any(1, M, K); }
size_t spinn2d_bwd_workspace_bytes(int N, int M, int K) {
  return plan_bwd(2, SPINN_KIND_GAUSSIAN, 2, N, M, K, true).ws_bytes;
}
size_t spinn1d_bwd_workspace_bytes(int N, int M, int K) {
  return plan_bwd(1, SPINN_KIND_GAUSSIAN, 2, N, M, K, true).ws_bytes;
}

int spinn2d_jets_fwd(int kind, int level, int N, int M_free, int M_fixed, int K, const float* x,
                     const float* y, const float* xc_free, const float* yc_free,
                     const float* xc_fixed, const float* yc_fixed, const float* h, int h_is_scalar,
                     const float* W, const float* bias, float* jets, void* workspace,
                     size_t workspace_bytes, void* stream) {
  return jets_fwd_impl(2, kind, level, N, M_free, M_fixed, K, x, y, xc_free, yc_free, xc_fixed,
                       yc_fixed, h, h_is_scalar, W, bias, jets, workspace, workspace_bytes, stream);
}

int spinn1d_jets_fwd(int kind, int level, int N, int M_free, int M_fixed, int K, const float* x,
                     const float* xc_free, const float* xc_fixed, const float* h, int h_is_scalar,
                     const float* W, const float* bias, float* jets, void* workspace,
                     size_t workspace_bytes, void* stream) {
  return jets_fwd_impl(1, kind, level, N, M_free, M_fixed, K, x, nullptr, xc_free, nullptr, xc_fixed,
                       nullptr, h, h_is_scalar, W, bias, jets, workspace, workspace_bytes, stream);
}

int spinn2d_jets_bwd(int kind, int level, int present_mask, int N, int M_free, int M_fixed, int K, const float* x,
                     const float* y, const float* xc_free, const float* yc_free,
                     const float* xc_fixed, const float* yc_fixed, const float* h, int h_is_scalar,
                     const float* W, const float* gjets, float* d_xc, float* d_yc, float* d_h,
                     float* d_W, float* d_bias, void* workspace, size_t workspace_bytes,
                     void* stream) {
  return jets_bwd_impl(2, kind, level, present_mask, N, M_free, M_fixed, K, x, y, xc_free, yc_free, xc_fixed,
                       yc_fixed, h, h_is_scalar, W, gjets, d_xc, d_yc, d_h, d_W, d_bias, workspace,
                       workspace_bytes, stream);
}

int spinn1d_jets_bwd(int kind, int level, int present_mask, int N, int M_free, int M_fixed, int K, const float* x,
                     const float* xc_free, const float* xc_fixed, const float* h, int h_is_scalar,
                     const float* W, const float* gjets, float* d_xc, float* d_h, float* d_W,
                     float* d_bias, void* workspace, size_t workspace_bytes, void* stream) {
  return jets_bwd_impl(1, kind, level, present_mask, N, M_free, M_fixed, K, x, nullptr, xc_free, nullptr, xc_fixed,
                       nullptr, h, h_is_scalar, W, gjets, d_xc, nullptr, d_h, d_W, d_bias, workspace,
                       workspace_bytes, stream);
}

size_t spinn2d_poisson_residual_workspace_bytes(int N) {
  (void)N;
  return align_up((size_t)1024 * sizeof(float), 256);
}

int spinn2d_poisson_residual(int N, double n_total, const float* x, const float* y,
                             const float* jets, float scale, float amp, float kx, float ky, float f0,
                             float* gjets, float* loss, void* workspace, size_t workspace_bytes,
                             void* stream) {
  if (N < 0 || n_total <= 0.0) return fail(SPINN_E_BADARG, "bad N/n_total");
  if (N == 0) return SPINN_OK;
  if (!x || !y || !jets || !gjets || !loss) return fail(SPINN_E_BADARG, "null pointer in poisson_residual");
  if (!workspace || workspace_bytes < spinn2d_poisson_residual_workspace_bytes(N))
    return fail(SPINN_E_WORKSPACE, "residual workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  int blocks = ceil_div_i(N, 256 * 4);
  if (blocks > 1024) blocks = 1024;
  if (blocks < 1) blocks = 1;
  float* lpart = (float*)workspace;
  poisson_residual_kernel<<<blocks, 256, 0, s>>>(N, (float)(1.0 / n_total), x, y, jets, scale, amp, kx,
                                                 ky, f0, gjets, lpart);
  LAUNCH_CHECK("poisson_residual_kernel");
  add_partials_kernel<<<1, 256, 0, s>>>(blocks, lpart, loss);
  LAUNCH_CHECK("add_partials_kernel");
  return SPINN_OK;
}

int spinn1d_ode3_residual(int N, double n_total, const float* x, const float* jets, float Kc, float* gjets,
                          float* loss, void* workspace, size_t workspace_bytes, void* stream) {
  if (N < 0 || n_total <= 0.0 || !(Kc > 0.f)) return fail(SPINN_E_BADARG, "bad N/n_total/K");
  if (N == 0) return SPINN_OK;
  if (!x || !jets || !gjets || !loss) return fail(SPINN_E_BADARG, "null pointer in ode3_residual");
  if (!workspace || workspace_bytes < spinn2d_poisson_residual_workspace_bytes(N))
    return fail(SPINN_E_WORKSPACE, "residual workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  int blocks = ceil_div_i(N, 256);
  if (blocks > 1024) blocks = 1024;
  float* lpart = (float*)workspace;
  ode3_residual_kernel<<<blocks, 256, 0, s>>>(N, (float)(1.0 / n_total), Kc, x, jets, gjets, lpart);
  LAUNCH_CHECK("ode3_residual_kernel");
  add_partials_kernel<<<1, 256, 0, s>>>(blocks, lpart, loss);
  LAUNCH_CHECK("add_partials_kernel");
  return SPINN_OK;
}

int spinn2d_cavity_residual(int N, float nu, const float* jets, float* gjets, float* loss, void* workspace,
                            size_t workspace_bytes, void* stream) {
  if (N < 0) return fail(SPINN_E_BADARG, "bad N");
  if (N == 0) return SPINN_OK;
  if (!jets || !gjets || !loss) return fail(SPINN_E_BADARG, "null pointer in cavity_residual");
  if (!workspace || workspace_bytes < spinn2d_poisson_residual_workspace_bytes(N))
    return fail(SPINN_E_WORKSPACE, "residual workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  int blocks = ceil_div_i(N, 256);
  if (blocks > 1024) blocks = 1024;
  float* lpart = (float*)workspace;
  cavity_residual_kernel<<<blocks, 256, 0, s>>>(N, nu, jets, gjets, lpart);
  LAUNCH_CHECK("cavity_residual_kernel");
  add_partials_kernel<<<1, 256, 0, s>>>(blocks, lpart, loss);
  LAUNCH_CHECK("add_partials_kernel");
  return SPINN_OK;
}

int spinn2d_poisson_interior_step(int kind, int N, double n_total, int M_free, int M_fixed, const float* x,
                                  const float* y, const float* xc_free, const float* yc_free,
                                  const float* xc_fixed, const float* yc_fixed, const float* h,
                                  int h_is_scalar, const float* W, const float* bias, float scale, float amp,
                                  float kx, float ky, float f0, float* jets, float* loss, float* d_xc,
                                  float* d_yc, float* d_h, float* d_W, float* d_bias, void* ws_fwd,
                                  size_t ws_fwd_bytes, void* ws_bwd, size_t ws_bwd_bytes, void* stream) {
  int rc = check_common(2, kind, 2, N, M_free, M_fixed, 1);
  if (rc) return rc;
  if (N <= 0 || n_total <= 0.0) return fail(SPINN_E_BADARG, "interior step needs N > 0 and n_total > 0");
  if (!x || !y || !jets || !loss || !h || !W || !d_h || !d_W || (M_free && (!xc_free || !yc_free || !d_xc || !d_yc)) ||
      (M_fixed && (!xc_fixed || !yc_fixed)))
    return fail(SPINN_E_BADARG, "null pointer argument in poisson_interior_step");
  const int M = M_free + M_fixed;
  cudaStream_t s = (cudaStream_t)stream;
  const FwdPlan fp = plan_fwd(2, kind, 2, M, 1);
  const BwdPlan bp = plan_bwd(2, kind, 2, N, M, 1, false, BWD_LAPL);
  if (!ws_fwd || ws_fwd_bytes < fp.ws_bytes || !ws_bwd || ws_bwd_bytes < bp.ws_bytes)
    return fail(SPINN_E_WORKSPACE, "interior step workspace too small (fwd %zu < %zu or bwd %zu < %zu)",
                ws_fwd_bytes, fp.ws_bytes, ws_bwd_bytes, bp.ws_bytes);
  char* bb = (char*)ws_bwd;
  // 1. node constants of both passes
  const int nthreads = fp.npairs > bp.Mp ? fp.npairs : bp.Mp;
  if (kind == SPINN_KIND_GAUSSIAN)
    pack_nodes_step_kernel<GAUSSIAN><<<ceil_div_i(nthreads, 128), 128, 0, s>>>(
        M, M_free, fp.npairs, bp.Mp, xc_free, yc_free, xc_fixed, yc_fixed, h, h_is_scalar, W, (float4*)ws_fwd,
        (float*)(bb + bp.off_soa));
  else
    pack_nodes_step_kernel<SOFTPLUS><<<ceil_div_i(nthreads, 128), 128, 0, s>>>(
        M, M_free, fp.npairs, bp.Mp, xc_free, yc_free, xc_fixed, yc_fixed, h, h_is_scalar, W, (float4*)ws_fwd,
        (float*)(bb + bp.off_soa));
  LAUNCH_CHECK("pack_nodes_step_kernel");
  // 2. forward jets + residual epilogue: loss partials (one per warp) and the backward's point-pair records
  const FwdEpilogue ep = {scale, amp, kx, ky, f0, (float)(1.0 / n_total), (float4*)(bb + bp.off_pts),
                          (float*)(bb + bp.off_loss)};
  int nlp = 0;
  rc = jets_fwd_impl(2, kind, 2, N, M_free, M_fixed, 1, x, y, xc_free, yc_free, xc_fixed, yc_fixed, h, h_is_scalar,
                     W, bias, jets, ws_fwd, ws_fwd_bytes, stream, true, &ep, &nlp);
  if (rc) return rc;
  if (nlp <= 0 || nlp > LOSS_BLOCKS) return fail(SPINN_E_CUDA, "unexpected forward grid (%d warps)", nlp);
  // 3.-4. backward on the packed records + second stage (gradients, bias gradient = 0, loss)
  const BwdFused fused = {nlp, loss};
  return jets_bwd_impl(2, kind, 2, 0x18, N, M_free, M_fixed, 1, x, y, xc_free, yc_free, xc_fixed, yc_fixed, h,
                       h_is_scalar, W, nullptr, d_xc, d_yc, d_h, d_W, d_bias, ws_bwd, ws_bwd_bytes, stream, &fused);
}

}  // extern "C"
