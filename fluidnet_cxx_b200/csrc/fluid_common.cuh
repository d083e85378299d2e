// fluid_common.cuh -- shared device helpers of the sm_100a fluid kernels.
//
// Arithmetic contract (see DESIGN.md "bit-exactness"): the reference evaluates
// every tensor op with its own fp32 rounding (ATen), so the translation units
// that include this header are compiled with -fmad=false: a*b+c stays FMUL+FADD.
// Divisions and sqrt are IEEE (nvcc defaults -prec-div/-prec-sqrt = true).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace fnx {

// Manta cell types, fp32 encoded (pytorch/lib/fluid/cell_type.py:5-14)
constexpr float kFluid = 1.0f;
constexpr float kObstacle = 2.0f;
constexpr float kEmpty = 4.0f;
constexpr float kOutflow = 16.0f;

constexpr float kHitMargin = 1e-5f;  // calc_line_trace.cpp:7
constexpr float kEpsilon = 1e-12f;   // calc_line_trace.cpp:8

struct Grid {
  int B, D, H, W;
  int sy;        // row stride   = W
  long long sz;  // plane stride = H*W
  long long n;   // cells per batch item = D*H*W
  int row0, row1;  // rows of the flattened (D*H) row space the one-thread-per-cell kernels process
                   // (whole grid by default; a slab window in the domain-decomposed step)
  int own0, own1;  // rows that count in a residual sum (whole grid by default; a slab's OWNED rows otherwise)
};

__host__ __device__ inline Grid make_grid(int B, int D, int H, int W) {
  Grid g;
  g.B = B; g.D = D; g.H = H; g.W = W;
  g.sy = W;
  g.sz = (long long)H * W;
  g.n = (long long)D * H * W;
  g.row0 = 0; g.row1 = D * H;
  g.own0 = 0; g.own1 = D * H;
  return g;
}

// border ring of width 1 (fluids_init.cpp:313-320, velocity_divergence.py:46-59)
template <bool Z>
__device__ __forceinline__ bool is_border(const Grid& g, int k, int j, int i) {
  bool m = (i < 1) | (i > g.W - 2) | (j < 1) | (j > g.H - 2);
  if (Z) m = m | (k < 1) | (k > g.D - 2);
  return m;
}

__device__ __forceinline__ float clamp01(float x) { return x < 0.f ? 0.f : (x > 1.f ? 1.f : x); }

// torch.clamp(x, lo, hi) == min(max(x, lo), hi)  (hi wins when lo > hi)
__device__ __forceinline__ long long clampll(long long x, long long lo, long long hi) {
  x = x < lo ? lo : x;
  return x > hi ? hi : x;
}

// float -> int64 truncation toward zero (ATen .toType(kLong)); 32-bit fast path
__device__ __forceinline__ long long trunc_ll(float x) {
  return (fabsf(x) < 1.0e9f) ? (long long)__float2int_rz(x) : __float2ll_rz(x);
}

// float -> int32 as x86 cvttss2si does it (ATen .toType(kInt) on the CPU path):
// out-of-range and NaN give INT_MIN ("integer indefinite").
__device__ __forceinline__ long long trunc_i32_x86(float x) {
  if (!(x > -2147483904.0f && x < 2147483648.0f)) return -2147483648LL;
  return (long long)__float2int_rz(x);
}

// clampll(trunc_ll(x), 0, hi) for EVERY float x, in 32-bit arithmetic: cvt.rzi.s32.f32 saturates
// (NaN -> 0, +-big -> INT_MAX / INT_MIN) and each saturated value lands on the same side of [0, hi]
// as the 64-bit truncation (NaN -> LLONG_MIN -> 0) does.
__device__ __forceinline__ int trunc_clamp0(float x, int hi) {
  int r = __float2int_rz(x);
  r = r < 0 ? 0 : r;
  return r > hi ? hi : r;
}
// clampll(trunc_i32_x86(x), 0, hi): the x86 conversion gives INT_MIN outside int32 and for NaN -> 0
__device__ __forceinline__ int trunc_x86_clamp0(float x, int hi) {
  if (!(x > -2147483904.0f && x < 2147483648.0f)) return hi < 0 ? hi : 0;
  int r = __float2int_rz(x);
  r = r < 0 ? 0 : r;
  return r > hi ? hi : r;
}

// at::min / at::max semantics of the oracle (strict compare, first operand kept on ties/NaN)
__device__ __forceinline__ float min_t(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float max_t(float a, float b) { return (b > a) ? b : a; }

}  // namespace fnx
