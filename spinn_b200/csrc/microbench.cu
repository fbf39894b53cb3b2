// microbench.cu -- instruction-throughput probes used to establish the FP32 / MUFU
// roofline denominators on the box (spinn_microbench).  Not on the product path.
#include <cuda_runtime.h>
#include "../../include/spinn_b200.h"
#include "spinn_math.cuh"
using namespace spinn;

// Every probe: 8 independent accumulator chains per thread (float2 each), operands
// come from REGISTERS loaded from global memory (so ptxas cannot fold them into
// immediates / constant-bank operands), `iters` loop iterations, 16 lane-FMAs per
// chain-step pair as documented per probe.
template <int WHICH>
__global__ void __launch_bounds__(256)
probe_kernel(int iters, const float* __restrict__ in, float* sink) {
  float2 a[8], b[8], c[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    a[k] = f2(in[(threadIdx.x + k) & 63], in[(threadIdx.x + 2 * k + 1) & 63]);
    b[k] = f2(in[(k + 3) & 63] + 1e-6f * threadIdx.x, in[(k + 5) & 63]);
    c[k] = f2(in[(k + 7) & 63], in[(k + 11) & 63] - 1e-6f * threadIdx.x);
  }
  const float2 m = f2(in[1], in[2]), d = f2(in[3], in[4]);
  for (int it = 0; it < iters; ++it) {
    if (WHICH == 0) {            // scalar FFMA, accumulate form, shared reg operands: a = a*m + d
#pragma unroll
      for (int k = 0; k < 8; ++k) { a[k].x = fmaf(a[k].x, m.x, d.x); a[k].y = fmaf(a[k].y, m.y, d.y); }
    } else if (WHICH == 1) {     // scalar FFMA, 3 distinct regs per instr: a = b*c + a
#pragma unroll
      for (int k = 0; k < 8; ++k) { a[k].x = fmaf(b[k].x, c[k].x, a[k].x); a[k].y = fmaf(b[k].y, c[k].y, a[k].y); }
    } else if (WHICH == 2) {     // FFMA2 shared operands: a = a*m + d
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] = f2fma(a[k], m, d);
    } else if (WHICH == 3) {     // FFMA2 3 distinct pairs: a = b*c + a
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] = f2fma(b[k], c[k], a[k]);
    } else if (WHICH == 4) {     // FFMA2 with scalar-broadcast operand (as the kernels use): a = b*m.x + a
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] = f2fma(b[k], f2dup(m.x), a[k]);
    } else if (WHICH == 5) {     // 8 FMUL2 + 8 FADD2 on INDEPENDENT chains (a *= m; b += d), so that
                                 // ptxas cannot contract a pair into one FFMA2 (SASS checked: tools/microbench.py)
#pragma unroll
      for (int k = 0; k < 8; ++k) { a[k] = f2mul(a[k], m); b[k] = f2add(b[k], d); }
    } else if (WHICH == 6) {     // scalar FFMA with immediate addend: a = a*m + 1e-9
#pragma unroll
      for (int k = 0; k < 8; ++k) { a[k].x = fmaf(a[k].x, m.x, 1e-9f); a[k].y = fmaf(a[k].y, m.y, 1e-9f); }
    } else if (WHICH == 7) {     // half FFMA2 (distinct) + half scalar FFMA (distinct): does mixing help?
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        a[k] = f2fma(b[k], c[k], a[k]);
        a[k + 1].x = fmaf(b[k + 1].x, c[k + 1].x, a[k + 1].x);
        a[k + 1].y = fmaf(b[k + 1].y, c[k + 1].y, a[k + 1].y);
      }
    } else if (WHICH == 8) {     // scalar FADD 2 distinct regs
#pragma unroll
      for (int k = 0; k < 8; ++k) { a[k].x = a[k].x + b[k].x; a[k].y = a[k].y + b[k].y; }
    } else if (WHICH == 9) {     // scalar FMUL 2 distinct regs
#pragma unroll
      for (int k = 0; k < 8; ++k) { a[k].x = a[k].x * b[k].x; a[k].y = a[k].y * b[k].y; }
    } else if (WHICH == 10) {    // MUFU.EX2
#pragma unroll
      for (int k = 0; k < 8; ++k) { a[k].x = ex2f_fast(a[k].x); a[k].y = ex2f_fast(a[k].y); }
    } else if (WHICH == 11 || WHICH == 13 || WHICH == 14) {
      // 16 FFMA2 (broadcast operand form) + 0 / 2 / 4 MUFU.EX2 per iteration: what does one
      // MUFU warp-instruction cost the FMA pipe?  (the kernels run 2 MUFU per 14..20 FFMA2)
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = f2fma(b[k], f2dup(m.x), a[k]);
      if (WHICH >= 13) { c[0].x = ex2f_fast(-c[0].x); c[0].y = ex2f_fast(-c[0].y); }
      if (WHICH == 14) { c[1].x = ex2f_fast(-c[1].x); c[1].y = ex2f_fast(-c[1].y); }
    } else if (WHICH == 15) {    // FMUL2 only, shared operand: a = a*m
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] = f2mul(a[k], m);
    } else if (WHICH == 16) {    // FADD2 only, shared operand: a = a+d
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] = f2add(a[k], d);
    } else if (WHICH == 17) {    // FMUL2 only, two distinct pairs: a = a*b
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] = f2mul(a[k], b[k]);
    } else if (WHICH == 18) {    // FADD2 only, two distinct pairs: a = a+b
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] = f2add(a[k], b[k]);
    } else if (WHICH == 19) {    // the forward kernel's own instruction mix per node pair and point, on
                                 // independent register chains: 7 FFMA2 : 4 FMUL2 : 3 FADD2 : 2 MUFU.EX2
      a[0] = f2add(a[0], d); a[1] = f2add(a[1], d);                 // X, Y
      a[2] = f2mul(a[2], m); a[3] = f2mul(a[3], m);                 // xi, eta
      a[4] = f2fma(a[4], m, d); a[5] = f2fma(a[5], m, d);           // hx, hy
      a[6] = f2add(a[6], d);                                        // t
      c[0].x = ex2f_fast(-c[0].x); c[0].y = ex2f_fast(-c[0].y);     // E
      b[0] = f2fma(b[1], m, b[0]);                                  // u
      a[7] = f2mul(a[7], m);                                        // g
      b[2] = f2fma(b[3], c[2], b[2]); b[4] = f2fma(b[3], c[3], b[4]);   // ux, uy
      c[4] = f2mul(c[4], m);                                        // e
      b[5] = f2fma(b[6], c[5], b[5]); b[7] = f2fma(b[6], c[6], b[7]);   // uxx, uyy
    } else if (WHICH == 12) {    // FFMA2 where b, c reused across two consecutive instrs (reuse cache)
#pragma unroll
      for (int k = 0; k < 8; k += 2) { a[k] = f2fma(b[k], c[k], a[k]); a[k + 1] = f2fma(b[k], c[k], a[k + 1]); }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k].x + a[k].y + b[k].x;
  if (WHICH >= 11) s += c[0].x + c[0].y + c[1].x + c[1].y;     // keep the MUFU results live
  if (WHICH == 5 || WHICH == 19) {
#pragma unroll
    for (int k = 0; k < 8; ++k) s += b[k].y + c[k].x + c[k].y;
  }
  if (s == 123.456f) sink[0] = s;
}

extern "C" int spinn_microbench(int which, int iters, int blocks, int threads, const float* in,
                                 float* sink, double* lane_ops, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const double lanes = (double)blocks * threads * iters;
  double per_iter = 16.0;
  switch (which) {
    case 0: probe_kernel<0><<<blocks, threads, 0, s>>>(iters, in, sink); break;
    case 1: probe_kernel<1><<<blocks, threads, 0, s>>>(iters, in, sink); break;
    case 2: probe_kernel<2><<<blocks, threads, 0, s>>>(iters, in, sink); break;
    case 3: probe_kernel<3><<<blocks, threads, 0, s>>>(iters, in, sink); break;
    case 4: probe_kernel<4><<<blocks, threads, 0, s>>>(iters, in, sink); break;
    case 5: probe_kernel<5><<<blocks, threads, 0, s>>>(iters, in, sink); per_iter = 32.0; break;
    case 6: probe_kernel<6><<<blocks, threads, 0, s>>>(iters, in, sink); break;
    case 7: probe_kernel<7><<<blocks, threads, 0, s>>>(iters, in, sink); break;
    case 8: probe_kernel<8><<<blocks, threads, 0, s>>>(iters, in, sink); break;
    case 9: probe_kernel<9><<<blocks, threads, 0, s>>>(iters, in, sink); break;
    case 10: probe_kernel<10><<<blocks, threads, 0, s>>>(iters, in, sink); break;
    case 11: probe_kernel<11><<<blocks, threads, 0, s>>>(iters, in, sink); per_iter = 32.0; break;
    case 12: probe_kernel<12><<<blocks, threads, 0, s>>>(iters, in, sink); break;
    case 13: probe_kernel<13><<<blocks, threads, 0, s>>>(iters, in, sink); per_iter = 32.0; break;
    case 14: probe_kernel<14><<<blocks, threads, 0, s>>>(iters, in, sink); per_iter = 32.0; break;
    case 15: probe_kernel<15><<<blocks, threads, 0, s>>>(iters, in, sink); break;
    case 16: probe_kernel<16><<<blocks, threads, 0, s>>>(iters, in, sink); break;
    case 17: probe_kernel<17><<<blocks, threads, 0, s>>>(iters, in, sink); break;
    case 18: probe_kernel<18><<<blocks, threads, 0, s>>>(iters, in, sink); break;
    case 19: probe_kernel<19><<<blocks, threads, 0, s>>>(iters, in, sink); per_iter = 28.0; break;  /* 14 packed instr */
    default: return -1;
  }
  if (lane_ops) *lane_ops = lanes * per_iter;
  return cudaPeekAtLastError() == cudaSuccess ? 0 : -3;
}
