/* spinn_b200.h -- C ABI of the B200-native SPINN hot path (libspinn_b200.so).
 *
 * The reference (nn4pde/SPINN) is pure Python/PyTorch and has no FFI: its
 * "operator" for this path is the duck-typed nn_cls protocol
 *   SPINN2D.forward(x, y)   /root/reference/code/spinn2d.py:169-179
 *   SPINN1D.forward(x)      /root/reference/code/spinn1d.py:147-156
 * differentiated by the problem classes with torch.autograd.grad
 *   RegularPDE._compute_derivatives  /root/reference/code/pde2d_base.py:46-62
 *   BasicODE._compute_derivatives    /root/reference/code/ode_base.py:32-42
 * and back-propagated by Optimizer.closure  /root/reference/code/common.py:242-248.
 * The entry points below are what a binding for that path calls instead: one
 * fused "jets" forward (u and its analytic spatial derivatives at N points
 * against M nodes) and one fused backward (parameter gradients of
 * L = sum g*J).  INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to contiguous fp32, owned by the caller
 *    (PyTorch tensors in practice); nothing is allocated or freed inside;
 *  - calls only ENQUEUE work on `stream` (a cudaStream_t passed as void*) and
 *    never synchronise; they are re-entrant (autograd calls the backward from
 *    its own device thread);
 *  - return value 0 on success, <0 on error (SPINN_E_*); spinn_last_error()
 *    returns a thread-local message for the last failure;
 *  - kind: 0 gaussian (spinn2d.py:191-192, spinn1d.py:176-177),
 *          1 softplus-hat (spinn2d.py:195-205, spinn1d.py:180-189);
 *  - level selects which jets are produced / which upstream gradients exist:
 *      2-D: 0:{u} 1:{u,ux,uy} 2:{u,ux,uy,uxx,uyy} 3:{u,ux,uy,uxx,uyy,uxy}
 *      1-D: 0:{u} 1:{u,ux}    2:{u,ux,uxx}
 *  - jets / gjets layout: [n_jets][N][K] (for K=1 every jet is a plain N-vector);
 *  - centres come as two arrays, free (learnable, spinn2d.py:109-110) and fixed
 *    (spinn2d.py:112-113); node j<M_free is free, the rest fixed, M=M_free+M_fixed;
 *  - h: [M], or [1] with h_is_scalar=1 (--fixed-h, spinn2d.py:115-117);
 *  - W: [K][M] row-major (nn.Linear weight, spinn2d.py:164); bias: [K] or NULL;
 *  - 1 <= K <= SPINN_MAX_K.
 */
#ifndef SPINN_B200_H
#define SPINN_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPINN_ABI_VERSION 2
#define SPINN_MAX_K 4

#define SPINN_KIND_GAUSSIAN 0
#define SPINN_KIND_SOFTPLUS 1

#define SPINN_OK 0
#define SPINN_E_BADARG (-1)
#define SPINN_E_WORKSPACE (-2)
#define SPINN_E_CUDA (-3)
#define SPINN_E_UNSUPPORTED (-4)

int spinn_abi_version(void);
const char* spinn_last_error(void);

/* number of jets for (dim, level), or <0 if the pair is invalid */
int spinn_num_jets(int dim, int level);

/* ---- 2-D: replaces SPINN2D.forward + 3x ag.grad (spinn2d.py:169-179, pde2d_base.py:46-62) */
size_t spinn2d_fwd_workspace_bytes(int N, int M, int K);
int spinn2d_jets_fwd(int kind, int level, int N, int M_free, int M_fixed, int K,
                     const float* x, const float* y,
                     const float* xc_free, const float* yc_free,
                     const float* xc_fixed, const float* yc_fixed,
                     const float* h, int h_is_scalar,
                     const float* W, const float* bias,
                     float* jets, void* workspace, size_t workspace_bytes, void* stream);

/* ---- 2-D backward: replaces loss.backward() through that graph (common.py:242-248).
 * d_xc/d_yc: [M_free]; d_h: [M] or [1]; d_W: [K][M]; d_bias: [K] or NULL.  Outputs are
 * overwritten (not accumulated).
 * present_mask: bit a set <=> row a of gjets holds an upstream gradient; rows whose bit is
 * clear count as zero and are NEVER READ (autograd's `None` for jets the loss does not use);
 * 0 means "all rows of the level".  The structure is exploited: e.g. a Laplacian-type
 * residual (only u_xx,u_yy: mask 0x18) or a Dirichlet penalty (only u: mask 0x01) run
 * specialised per-pair code with fewer instructions -- results are those of the full formula. */
size_t spinn2d_bwd_workspace_bytes(int N, int M, int K);
int spinn2d_jets_bwd(int kind, int level, int present_mask, int N, int M_free, int M_fixed, int K,
                     const float* x, const float* y,
                     const float* xc_free, const float* yc_free,
                     const float* xc_fixed, const float* yc_fixed,
                     const float* h, int h_is_scalar,
                     const float* W, const float* gjets,
                     float* d_xc, float* d_yc, float* d_h, float* d_W, float* d_bias,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---- 1-D twins: SPINN1D.forward + 2x ag.grad (spinn1d.py:147-156, ode_base.py:32-42) */
size_t spinn1d_fwd_workspace_bytes(int N, int M, int K);
int spinn1d_jets_fwd(int kind, int level, int N, int M_free, int M_fixed, int K,
                     const float* x, const float* xc_free, const float* xc_fixed,
                     const float* h, int h_is_scalar,
                     const float* W, const float* bias,
                     float* jets, void* workspace, size_t workspace_bytes, void* stream);
size_t spinn1d_bwd_workspace_bytes(int N, int M, int K);
int spinn1d_jets_bwd(int kind, int level, int present_mask, int N, int M_free, int M_fixed, int K,
                     const float* x, const float* xc_free, const float* xc_fixed,
                     const float* h, int h_is_scalar,
                     const float* W, const float* gjets,
                     float* d_xc, float* d_h, float* d_W, float* d_bias,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---- fused residual epilogue for the stock Poisson residual
 * res = scale*(u_xx + u_yy + f(x,y)),  f = amp*sin(kx*x)*sin(ky*y) + f0
 * (poisson2d_sine.py:16-18: scale .1, amp 20pi^2, kx 2pi, ky 4pi;
 *  poisson2d_irreg_dom.py:146-147: scale 1, amp 0, f0 1) and loss = sum(res^2)/n_total
 * (pde2d_base.py:139-141).  Reads jets [5][N], writes gjets [5][N] = dLoss/dJ and adds
 * the loss into *loss (one float, caller zeroes it).  K=1 only. */
int spinn2d_poisson_residual(int N, double n_total, const float* x, const float* y,
                             const float* jets, float scale, float amp, float kx, float ky,
                             float f0, float* gjets, float* loss, void* workspace,
                             size_t workspace_bytes, void* stream);
size_t spinn2d_poisson_residual_workspace_bytes(int N);

/* ---- residual epilogues of the other BASELINE problem scripts (same contract as
 * spinn2d_poisson_residual: *loss is ADDED to, workspace sized by
 * spinn2d_poisson_residual_workspace_bytes):
 *  ode3  (ode3.py:12-16, ode_base.py:83-92): res = u_xx + f(x), f the script's forcing with K = Kc
 *        (0.01), loss = sum(res^2)/n_total; jets [3][N] -> gjets [3][N] of which ONLY row 2 (u_xx)
 *        is written: pass present_mask 0x4 to spinn1d_jets_bwd.
 *  cavity (cavity.py:84-100): K = 3 outputs (u, v, p), jets/gjets [5][N][3];
 *        loss = sum_i (u_x+v_y)^2 + (u u_x + v u_y + p_x - nu lap u)^2 + (u v_x + v v_y + p_y - nu lap v)^2
 *        (a sum over points, not a mean); all 15 entries of gjets are written. */
int spinn1d_ode3_residual(int N, double n_total, const float* x, const float* jets, float Kc,
                          float* gjets, float* loss, void* workspace, size_t workspace_bytes, void* stream);
int spinn2d_cavity_residual(int N, float nu, const float* jets, float* gjets, float* loss,
                            void* workspace, size_t workspace_bytes, void* stream);

/* ---- the whole interior step of the stock Poisson residual in ONE call (K = 1):
 * what RegularPDE.interior_loss + loss.backward() compute (pde2d_base.py:132-141, common.py:242-248)
 * for res = scale*(u_xx + u_yy + amp*sin(kx x) sin(ky y) + f0), loss = sum(res^2)/n_total.
 * Four launches: node constants of both passes; forward jets WITH the residual epilogue (loss
 * partials and the backward's point-pair records are written by the forward kernel itself -- the
 * upstream gradient never exists as an [5][N] array); structured backward; second stage
 * (gradients + loss).  jets [5][N] and *loss are overwritten;
 * gradients as in spinn2d_jets_bwd (d_bias, if given, receives 0: the residual does not depend on u).
 * ws_fwd / ws_bwd: sized by spinn2d_fwd_workspace_bytes / spinn2d_bwd_workspace_bytes. */
int spinn2d_poisson_interior_step(int kind, int N, double n_total, int M_free, int M_fixed,
                                  const float* x, const float* y,
                                  const float* xc_free, const float* yc_free,
                                  const float* xc_fixed, const float* yc_fixed,
                                  const float* h, int h_is_scalar, const float* W, const float* bias,
                                  float scale, float amp, float kx, float ky, float f0,
                                  float* jets, float* loss,
                                  float* d_xc, float* d_yc, float* d_h, float* d_W, float* d_bias,
                                  void* ws_fwd, size_t ws_fwd_bytes, void* ws_bwd, size_t ws_bwd_bytes,
                                  void* stream);

/* ---- compact-support ("cut-off") evaluation, 2-D Gaussian kernel, K = 1  (OPT-IN; SURVEY 8f-4,
 * manuscript :1089-1095).  Same per-pair formulas as the dense calls, but (32-node block) x
 * (64-point group) combinations whose bounding boxes are farther apart than cutoff * h_max are
 * skipped, so results differ from the dense ones only by the Gaussian tail beyond `cutoff`
 * widths (relative size exp(-cutoff^2/2)).  node_perm: int32[M] visiting order of the nodes
 * (spatially sorted, e.g. Morton; NULL = given order) -- it affects speed only, gradients come
 * back in the caller's node order.  Consecutive points should be spatially close (grids are).
 * stats: optional device counter; receives the number of point-node pairs actually evaluated
 * (the dense count is N*M) so that skipped work can be reported as skipped.  The forward keeps
 * all node records in shared memory (M up to ~4500); larger M returns SPINN_E_UNSUPPORTED. */
size_t spinn2d_sparse_fwd_workspace_bytes(int N, int M);
size_t spinn2d_sparse_bwd_workspace_bytes(int N, int M);
int spinn2d_jets_fwd_sparse(int kind, int level, int N, int M_free, int M_fixed, int K,
                            const float* x, const float* y,
                            const float* xc_free, const float* yc_free,
                            const float* xc_fixed, const float* yc_fixed,
                            const float* h, int h_is_scalar, const float* W, const float* bias,
                            const int* node_perm, float cutoff, float* jets,
                            unsigned long long* stats, void* workspace, size_t workspace_bytes,
                            void* stream);
int spinn2d_jets_bwd_sparse(int kind, int level, int present_mask, int N, int M_free, int M_fixed, int K,
                            const float* x, const float* y,
                            const float* xc_free, const float* yc_free,
                            const float* xc_fixed, const float* yc_fixed,
                            const float* h, int h_is_scalar, const float* W, const float* gjets,
                            const int* node_perm, float cutoff,
                            float* d_xc, float* d_yc, float* d_h, float* d_W, float* d_bias,
                            unsigned long long* stats, void* workspace, size_t workspace_bytes,
                            void* stream);

/* ---- development hook (tools/tune.py): select kernel tile-shape variants, "fwd=<id>,bwd=<id>,waves=<n>";
 * NULL restores the defaults.  The same string is read ONCE from the environment variable SPINN_TUNE
 * when the library is loaded; the compute entry points never read the environment. */
void spinn_set_tuning(const char* spec);

/* ---- instruction-throughput probes used to measure the roofline denominators on the box
 * (spinn_b200/csrc/microbench.cu).  Every thread runs 8 independent float2 accumulator chains
 * for `iters` iterations with operands held in registers loaded from `in` (>= 64 floats), so
 * nothing folds into immediates.  *lane_ops receives the number of per-lane operations issued;
 * timing is the caller's (CUDA events on `stream`).  which:
 *   0 FFMA a=a*m+d (shared regs)   1 FFMA a=b*c+a (3 distinct regs)   2 FFMA2 a=a*m+d
 *   3 FFMA2 a=b*c+a (3 distinct register pairs)   4 FFMA2 with a scalar-broadcast operand
 *   5 FMUL2+FADD2   6 FFMA with immediate   7 half FFMA2 / half FFMA   8 FADD   9 FMUL
 *   10 MUFU.EX2   11/13/14 16 FFMA2 (broadcast operand) + 0/2/4 MUFU.EX2   12 FFMA2 with pairwise-reused
 *   operands   15 FMUL2 a=a*m   16 FADD2 a=a+d   17 FMUL2 a=a*b   18 FADD2 a=a+b
 *   19 the Gaussian forward's instruction mix (7 FFMA2 : 4 FMUL2 : 3 FADD2 : 2 MUFU.EX2) */
int spinn_microbench(int which, int iters, int blocks, int threads, const float* in, float* sink,
                     double* lane_ops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPINN_B200_H */
