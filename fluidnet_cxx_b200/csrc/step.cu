// step.cu -- fused entry points behind lib.simulate (pytorch/lib/simulate.py:28-171) and the
// global reduction of the FluidNet wrapper (_ScaleNet, pytorch/lib/model.py:8-23).
// Compiled with -fmad=false.
#include <cuda_runtime.h>

#include "../../include/fluidstep.h"
#include "advect_cells.cuh"
#include "advect_device.cuh"
#include "fluid_common.cuh"
#include "host_util.h"
#include "stencil_device.cuh"
#include "step2d.h"

#include <stdlib.h>

namespace fnx {

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
#ifndef FNX_STEP2D_MIN_CELLS
#define FNX_STEP2D_MIN_CELLS (1LL << 20)
#endif

// FNX_STEP2D=0 routes the 2-D step through the generic one-thread-per-cell kernels below (the 3-D
// path), kept as the cross-check of the fused 2-D kernels (tests/test_gpu_parity.py)
static bool use_clean() {
  const char* e = getenv("FNX_STEP2D_CLEAN");
  return !(e && e[0] == '0');
}
// The fused 2-D kernels pay a fixed latency (a tile of 2 048 cells per CTA, two dependent advection kernels, a
// 32-row march per thread in the forces kernel): below ~1 M cells the grid is a handful of tiles per SM and the
// one-thread-per-cell kernels are faster (measured: 512^2 ScaleNet step 0.597 vs 0.621 ms, 128^2 Jacobi-28 step
// 0.103 vs 0.124 ms).  FNX_STEP2D=1 / 0 forces one or the other; held-row (slab) arrays always take the fused ones.
static bool use_step2d(long long cells = (1LL << 40)) {
  const char* e = getenv("FNX_STEP2D");
  if (e && e[0] == '0') return false;
  if (e && e[0] == '1') return true;
  return cells >= FNX_STEP2D_MIN_CELLS;
}

// ---- unbiased std (model.py:18) ----------------------------------------------------------
// pass 1: per-batch sum -> mean; pass 2: sum (x-mean)^2; both in double (torch accumulates the
// fp32 CPU std in double as well), one atomic per block.
__global__ void __launch_bounds__(256)
    k_sum(const float* __restrict__ x, size_t count, double* __restrict__ out, const double* __restrict__ sum_in,
          int pass) {
  const int b = blockIdx.y;
  const float* xb = x + (size_t)b * count;
  double mean = 0.0;
  if (pass == 1) mean = sum_in[b] / (double)count;
  double acc = 0.0;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < count; q += (size_t)gridDim.x * blockDim.x) {
    const double v = (double)__ldg(xb + q);
    if (pass == 0) acc += v;
    else { const double d = v - mean; acc += d * d; }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  __shared__ double wsum[8];
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++) t += wsum[w];
    atomicAdd(out + b, t);
  }
}

__global__ void k_std_finish(const double* __restrict__ ssq, size_t count, int B, float threshold,
                             float* __restrict__ scale) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float s = (float)sqrt(ssq[b] / (double)(count - 1));
  scale[b] = s < threshold ? threshold : s;  // torch.clamp(std, min=threshold); NaN stays NaN
}

// ---- fused step kernels (simulate.py:28-171) ---------------------------------------------------
struct StepMasks {
  const float* UBC;
  const float* UBCInv;
  const float* rBC;
  const float* rBCInv;
  const unsigned char* rows;  // per (b, d*H+row): bit0 = U masks differ from identity, bit1 = density masks
};
struct StepForces {
  int use_buoyancy, use_gravity, wall_bcs, density_passes;
  float bstrength[3], gforce[3];
  float rho_star;
};

// one block per grid row: does this row's mask differ from (InvMask = 1, BC = 0) anywhere?
__global__ void __launch_bounds__(128)
    k_mask_rows(Grid g, int nc, const float* __restrict__ UBC, const float* __restrict__ UBCInv,
                const float* __restrict__ rBC, const float* __restrict__ rBCInv, unsigned char* __restrict__ rows) {
  const int row = blockIdx.x;  // over B*D*H
  const int b = row / (g.D * g.H), r = row - b * (g.D * g.H);
  int uact = 0, ract = 0;
  for (int i = threadIdx.x; i < g.W; i += blockDim.x) {
    if (UBC)
      for (int c = 0; c < nc; c++) {
        const long long o = ((long long)b * nc + c) * g.n + (long long)r * g.W + i;
        uact |= (__ldg(UBC + o) != 0.f) | (__ldg(UBCInv + o) != 1.f);
      }
    if (rBC) {
      const long long o = (long long)b * g.n + (long long)r * g.W + i;
      ract |= (__ldg(rBC + o) != 0.f) | (__ldg(rBCInv + o) != 1.f);
    }
  }
  const int anyu = __syncthreads_or(uact);  // (returns a predicate, not a bitwise OR)
  const int anyr = __syncthreads_or(ract);
  if (threadIdx.x == 0) rows[row] = (unsigned char)((anyu ? 1 : 0) | (anyr ? 2 : 0));
}

// x*InvMask + BC for velocity component c / density, skipping rows whose masks are the identity.
// NOTE: x*1 + 0 == x bit-for-bit except -0 -> +0 and NaN payloads, which compare equal.
__device__ __forceinline__ float cv_u(const StepMasks& m, const Grid& g, int nc, int b, int row, int c,
                                      long long o, float x) {
  if (!m.UBC) return x;
  if (m.rows && !(m.rows[b * g.D * g.H + row] & 1)) return x;
  const long long q = ((long long)b * nc + c) * g.n + o;
  return const_vals_apply(x, __ldg(m.UBCInv + q), __ldg(m.UBC + q));
}
__device__ __forceinline__ float cv_r(const StepMasks& m, const Grid& g, int b, int row, long long o, float x,
                                      int passes) {
  if (!m.rBC || passes <= 0) return x;
  if (m.rows && !(m.rows[b * g.D * g.H + row] & 2)) return x;
  const long long q = (long long)b * g.n + o;
  const float inv = __ldg(m.rBCInv + q), bc = __ldg(m.rBC + q);
  for (int t = 0; t < passes; t++) x = const_vals_apply(x, inv, bc);
  return x;
}

template <bool Z>
__global__ void __launch_bounds__(kBX* kBY)
    k_step_advect_fwd(Grid g, float mdt, const float* __restrict__ rho, const float* __restrict__ U,
                      const float* __restrict__ flags, int sample_outside, float* __restrict__ rho_fwd,
                      int* __restrict__ fidx, float* __restrict__ U_fwd) {
  constexpr int NC = Z ? 3 : 2;
  CellIdx c;
  if (!cell_of(g, c)) return;
  const long long bo = (long long)c.b * g.n;
  int idx;
  rho_fwd[bo + c.o] = scalar_fwd_cell<Z>(g, c, mdt, rho + bo, U + bo * NC, flags + bo, sample_outside, true, &idx);
  fidx[bo + c.o] = idx;
  float out[3];
  vel_fwd_cell<Z>(g, c, mdt, U + bo * NC, U + bo * NC, flags + bo, out);
#pragma unroll
  for (int a = 0; a < NC; a++) U_fwd[bo * NC + (long long)a * g.n + c.o] = out[a];
}

// MacCormack backward pass + correction + clamp of both fields, then setConstVals (simulate.py:96)
template <bool Z>
__global__ void __launch_bounds__(kBX* kBY)
    k_step_advect_bwd(Grid g, float dt, float hs, const float* __restrict__ rho, const float* __restrict__ U,
                      const float* __restrict__ flags, int sample_outside, const float* __restrict__ rho_fwd,
                      const int* __restrict__ fidx, const float* __restrict__ U_fwd, StepMasks m,
                      float* __restrict__ rho1, float* __restrict__ U1) {
  constexpr int NC = Z ? 3 : 2;
  CellIdx c;
  if (!cell_of(g, c)) return;
  const long long bo = (long long)c.b * g.n;
  const int row = c.k * g.H + c.j;
  float r = scalar_bwd_cell<Z>(g, c, dt, hs, rho + bo, U + bo * NC, flags + bo, sample_outside, rho_fwd + bo, fidx + bo);
  rho1[bo + c.o] = cv_r(m, g, c.b, row, c.o, r, 1);
  float out[3];
  vel_bwd_cell<Z>(g, c, dt, hs, U + bo * NC, U + bo * NC, flags + bo, U_fwd + bo * NC, out);
#pragma unroll
  for (int a = 0; a < NC; a++) U1[bo * NC + (long long)a * g.n + c.o] = cv_u(m, g, NC, c.b, row, a, c.o, out[a]);
}

// velocity component `comp` of cell (k,j,i) after addBuoyancy -> addGravity -> [setWallBcs] -> setConstVals
// (simulate.py:98-133), computed from the post-advection fields rho1 / U1.
template <bool Z>
__device__ __forceinline__ float forced_vel(const Grid& g, int b, int k, int j, int i, long long o, int comp,
                                            const float* __restrict__ rho1, const float* __restrict__ U1,
                                            const float* __restrict__ flags, const StepMasks& m,
                                            const StepForces& f) {
  constexpr int NC = Z ? 3 : 2;
  float u = __ldg(U1 + (long long)comp * g.n + o);
  const float fc = __ldg(flags + o);
  const int idx = comp == 0 ? i : (comp == 1 ? j : k);
  const long long on = o - nb_off(g, comp);
  const float fn = idx > 0 ? __ldg(flags + on) : fc;
  if (!is_border<Z>(g, k, j, i)) {
    if (f.use_buoyancy) u = buoyancy_apply(u, fc, fn, __ldg(rho1 + o), __ldg(rho1 + on), f.bstrength[comp], f.rho_star);
    if (f.use_gravity) u = gravity_apply(u, fc, fn, f.gforce[comp]);
  }
  if (f.wall_bcs) u = wall_bcs_apply(u, fc, fn);
  return cv_u(m, g, NC, b, k * g.H + j, comp, o, u);
}

// forces + wall BCs + setConstVals + divergence (simulate.py:98-145) in one pass
template <bool Z>
__global__ void __launch_bounds__(kBX* kBY)
    k_step_forces_div(Grid g, const float* __restrict__ rho1, const float* __restrict__ U1,
                      const float* __restrict__ flags, StepMasks m, StepForces f, float* __restrict__ rho_out,
                      float* __restrict__ U_out, float* __restrict__ div) {
  constexpr int NC = Z ? 3 : 2;
  CellIdx c;
  if (!cell_of(g, c)) return;
  const long long bo = (long long)c.b * g.n;
  rho1 += bo; flags += bo; U1 += bo * NC;
  float u[3];
#pragma unroll
  for (int a = 0; a < NC; a++) {
    u[a] = forced_vel<Z>(g, c.b, c.k, c.j, c.i, c.o, a, rho1, U1, flags, m, f);
    U_out[bo * NC + (long long)a * g.n + c.o] = u[a];
  }
  rho_out[bo + c.o] = cv_r(m, g, c.b, c.k * g.H + c.j, c.o, __ldg(rho1 + c.o), f.density_passes);
  if (div) {
    float v = 0.f;
    if (!is_border<Z>(g, c.k, c.j, c.i)) {
      const float ux = forced_vel<Z>(g, c.b, c.k, c.j, c.i + 1, c.o + 1, 0, rho1, U1, flags, m, f);
      const float uy = forced_vel<Z>(g, c.b, c.k, c.j + 1, c.i, c.o + g.sy, 1, rho1, U1, flags, m, f);
      v = u[0] - ux + u[1] - uy;
      if (Z) {
        const float uz = forced_vel<Z>(g, c.b, c.k + 1, c.j, c.i, c.o + g.sz, 2, rho1, U1, flags, m, f);
        v = v + (u[2] - uz);
      }
    }
    if (__ldg(flags + c.o) == kObstacle) v = 0.f;
    div[bo + c.o] = v;
  }
}

// velocityUpdate -> [setWallBcs] -> setConstVals (simulate.py:154-168), in place on U
template <bool Z>
__global__ void __launch_bounds__(kBX* kBY)
    k_step_project(Grid g, const float* __restrict__ p, float* __restrict__ U, const float* __restrict__ flags,
                   StepMasks m, int wall_bcs) {
  constexpr int NC = Z ? 3 : 2;
  CellIdx c;
  if (!cell_of(g, c)) return;
  const long long bo = (long long)c.b * g.n;
  flags += bo; p += bo; U += bo * NC;
  const float fc = __ldg(flags + c.o);
  const bool interior = !is_border<Z>(g, c.k, c.j, c.i);
  const float P = __ldg(p + c.o);
  const int idx[3] = {c.i, c.j, c.k};
#pragma unroll
  for (int a = 0; a < NC; a++) {
    const long long on = c.o - nb_off(g, a);
    const float fn = idx[a] > 0 ? __ldg(flags + on) : fc;
    float u = U[(long long)a * g.n + c.o];
    if (interior) u = velocity_update_apply(u, fc, fn, P, __ldg(p + on));
    if (wall_bcs) u = wall_bcs_apply(u, fc, fn);
    U[(long long)a * g.n + c.o] = cv_u(m, g, NC, c.b, c.k * g.H + c.j, a, c.o, u);
  }
}

}  // namespace fnx

using namespace fnx;

#define FNX_TRY(call)            \
  do {                           \
    int e_ = (call);             \
    if (e_ != FNX_OK) return e_; \
  } while (0)

#define FNX_CUDA_TRY(who, call)                                                      \
  do {                                                                               \
    cudaError_t e_ = (call);                                                         \
    if (e_ != cudaSuccess) return fnx_set_error(FNX_ERR_CUDA, "%s: %s", who, cudaGetErrorString(e_)); \
  } while (0)

extern "C" {

size_t fnx_scale_std_workspace(int B) { return align256(2 * (size_t)B * sizeof(double)); }

int fnx_scale_std(const float* x, size_t count_per_batch, int B, float threshold, float* scale, void* workspace,
                  size_t workspace_bytes, void* stream) {
  if (B < 1 || count_per_batch < 2) return fnx_set_error(FNX_ERR_ARG, "scale_std: need B>=1 and >=2 elements");
  if (!workspace || workspace_bytes < fnx_scale_std_workspace(B))
    return fnx_set_error(FNX_ERR_WORKSPACE, "scale_std: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  double* sum = (double*)workspace;
  double* ssq = sum + B;
  FNX_CUDA_TRY("scale_std", cudaMemsetAsync(sum, 0, 2 * (size_t)B * sizeof(double), st));
  size_t want = (count_per_batch + 256 * 8 - 1) / (256 * 8);
  unsigned nblk = (unsigned)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
  dim3 grid(nblk, B);
  k_sum<<<grid, 256, 0, st>>>(x, count_per_batch, sum, nullptr, 0);
  k_sum<<<grid, 256, 0, st>>>(x, count_per_batch, ssq, sum, 1);
  k_std_finish<<<(B + 63) / 64, 64, 0, st>>>(ssq, count_per_batch, B, threshold, scale);
  fnx_count_launches(3);
  FNX_CUDA_TRY("scale_std", cudaGetLastError());
  return FNX_OK;
}

// ---- fused step ------------------------------------------------------------------------------
static inline unsigned char* ws_take(char*& ws, size_t bytes) {
  unsigned char* p = (unsigned char*)ws;
  ws += align256(bytes);
  return p;
}

size_t fnx_step_workspace(int B, int D, int H, int W, int is3d) {
  const size_t n = (size_t)B * D * H * W;
  const int nc = is3d ? 3 : 2;
  size_t s = 0;
  s += align256(n * sizeof(float));        // rho_fwd
  s += align256(n * sizeof(int));          // traced cell index
  s += align256(n * nc * sizeof(float));   // U_fwd
  s += align256(n * sizeof(float));        // rho after advection + BCs
  s += align256(n * nc * sizeof(float));   // U after advection + BCs
  s += align256(n * sizeof(float));        // div (fnx_step_jacobi)
  s += align256(fnx_jacobi_workspace(B, D, H, W, 1));
  return s;
}

int fnx_mask_rows(const float* UBC, const float* UBCInvMask, const float* densityBC, const float* densityBCInvMask,
                  unsigned char* rows, int B, int D, int H, int W, int is3d, void* stream) {
  if (B < 1 || D < 1 || H < 1 || W < 1) return fnx_set_error(FNX_ERR_ARG, "mask_rows: bad grid");
  Grid g = make_grid(B, D, H, W);
  const int nrows = B * D * H;
  const bool ubc = UBC && UBCInvMask, rbc = densityBC && densityBCInvMask;
  k_mask_rows<<<nrows, 128, 0, (cudaStream_t)stream>>>(g, is3d ? 3 : 2, ubc ? UBC : nullptr, ubc ? UBCInvMask : nullptr,
                                                       rbc ? densityBC : nullptr, rbc ? densityBCInvMask : nullptr, rows);
  fnx_count_launches(1);
  FNX_CUDA_TRY("mask_rows", cudaGetLastError());
  return FNX_OK;
}

int fnx_step_advect_forces_div(const fnx_step_params* prm, const float* density_in, const float* U_in,
                               const float* flags, const float* UBC, const float* UBCInvMask,
                               const float* densityBC, const float* densityBCInvMask, const unsigned char* mask_rows,
                               float* density_out, float* U_out, float* div, int B, int D, int H, int W, int is3d,
                               void* workspace, size_t workspace_bytes, void* stream) {
  if (!prm) return fnx_set_error(FNX_ERR_ARG, "step: null params");
  if (B < 1 || D < 1 || H < 2 || W < 2 || (is3d && D < 2) || (!is3d && D != 1) || (long long)D * H * W >= (1LL << 31))
    return fnx_set_error(FNX_ERR_ARG, "step: unsupported grid B=%d D=%d H=%d W=%d is3d=%d", B, D, H, W, is3d);
  const int Hheld = (prm->held_row_end > prm->held_row_begin) ? prm->held_row_end - prm->held_row_begin : H;
  if (!workspace || workspace_bytes < fnx_step_workspace(B, D, Hheld, W, is3d))
    return fnx_set_error(FNX_ERR_WORKSPACE, "step: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)B * D * Hheld * W;
  const int nc = is3d ? 3 : 2;
  char* ws = (char*)workspace;
  float* rho_fwd = (float*)ws_take(ws, n * sizeof(float));
  int* fidx = (int*)ws_take(ws, n * sizeof(int));
  float* U_fwd = (float*)ws_take(ws, n * nc * sizeof(float));
  float* rho1 = (float*)ws_take(ws, n * sizeof(float));
  float* U1 = (float*)ws_take(ws, n * nc * sizeof(float));
  Grid g = make_grid(B, D, H, W);
  if (prm->row_end > prm->row_begin) {  // slab window (domain-decomposed step): only these rows are computed
    if (prm->row_begin < 0 || prm->row_end > D * H) return fnx_set_error(FNX_ERR_ARG, "step: row window out of range");
    g.row0 = prm->row_begin; g.row1 = prm->row_end;
  }
  const bool held = prm->held_row_end > prm->held_row_begin;   // arrays hold these rows only
  if (held && (is3d || !fnx_step2d_supported(H, W)))
    return fnx_set_error(FNX_ERR_ARG, "step: row-window arrays (held_row_*) are a 2-D feature of the fused kernels");
  if (!is3d && fnx_step2d_supported(H, W) && (held || use_step2d((long long)(g.row1 - g.row0) * W))) {
    // 2-D: fused advection (forward pass staged in shared memory) + tiled forces/divergence (step2d.cu)
    fnx_step2d_win w;
    w.H = H; w.W = W; w.row0 = g.row0; w.row1 = g.row1;
    w.ya0 = held ? prm->held_row_begin : 0; w.ya1 = held ? prm->held_row_end : H;
    FNX_TRY(fnx_step2d_check_window(w));
    const bool ubc2 = UBC && UBCInvMask, rbc2 = densityBC && densityBCInvMask;
    fnx_step2d_masks mk;
    mk.UBC = ubc2 ? UBC : nullptr; mk.UBCInv = ubc2 ? UBCInvMask : nullptr;
    mk.rBC = rbc2 ? densityBC : nullptr; mk.rBCInv = rbc2 ? densityBCInvMask : nullptr;
    mk.rows = mask_rows;
    float* rho_mid = rbc2 ? rho_fwd : nullptr;   // one-pass density of the masked rows (what addBuoyancy reads)
    // the interior fast path (FNX_STEP2D_CLEAN=0 switches it off: generic fused kernel only) keeps its list of
    // declined tiles in the slot of the traced-index array the fused kernels do not need
    int* tile_ws = use_clean() && fnx_step2d_tile_ws_ints(w, B) * sizeof(int) <= n * sizeof(int) ? fidx : nullptr;
    FNX_TRY(fnx_step2d_advect(w, prm->dt, prm->maccormack_strength, prm->sample_outside_fluid, density_in, U_in, flags,
                              mk, prm->density_const_passes, density_out, rho_mid, U1, B, tile_ws, st));
    FNX_TRY(fnx_step2d_forces_div(w, prm, density_out, rho_mid, U1, flags, mk, U_out, div, B, st));
    FNX_CUDA_TRY("step", cudaGetLastError());
    return FNX_OK;
  }
  StepMasks m;
  const bool ubc = UBC && UBCInvMask, rbc = densityBC && densityBCInvMask;
  m.UBC = ubc ? UBC : nullptr; m.UBCInv = ubc ? UBCInvMask : nullptr;
  m.rBC = rbc ? densityBC : nullptr; m.rBCInv = rbc ? densityBCInvMask : nullptr;
  m.rows = mask_rows;
  StepForces f;
  f.use_buoyancy = prm->use_buoyancy; f.use_gravity = prm->use_gravity;
  for (int a = 0; a < 3; a++) {
    f.bstrength[a] = prm->buoyancy3[a] * prm->dt;  // gravity*dt, one fp32 product (source_terms.py:70)
    f.gforce[a] = prm->gravity3[a] * prm->dt;
  }
  f.rho_star = prm->rho_star;
  f.wall_bcs = prm->apply_wall_bcs;
  f.density_passes = prm->density_const_passes;
  const float hs = prm->maccormack_strength * 0.5f;
  dim3 gr = cell_grid(g), bl = cell_block();
  if (is3d) {
    k_step_advect_fwd<true><<<gr, bl, 0, st>>>(g, -prm->dt, density_in, U_in, flags, prm->sample_outside_fluid, rho_fwd, fidx, U_fwd);
    k_step_advect_bwd<true><<<gr, bl, 0, st>>>(g, prm->dt, hs, density_in, U_in, flags, prm->sample_outside_fluid, rho_fwd, fidx, U_fwd, m, rho1, U1);
    k_step_forces_div<true><<<gr, bl, 0, st>>>(g, rho1, U1, flags, m, f, density_out, U_out, div);
  } else {
    k_step_advect_fwd<false><<<gr, bl, 0, st>>>(g, -prm->dt, density_in, U_in, flags, prm->sample_outside_fluid, rho_fwd, fidx, U_fwd);
    k_step_advect_bwd<false><<<gr, bl, 0, st>>>(g, prm->dt, hs, density_in, U_in, flags, prm->sample_outside_fluid, rho_fwd, fidx, U_fwd, m, rho1, U1);
    k_step_forces_div<false><<<gr, bl, 0, st>>>(g, rho1, U1, flags, m, f, density_out, U_out, div);
  }
  fnx_count_launches(3);
  FNX_CUDA_TRY("step", cudaGetLastError());
  return FNX_OK;
}

int fnx_step_project_bcs(const float* pressure, float* U, const float* flags, const float* UBC,
                         const float* UBCInvMask, const unsigned char* mask_rows, int apply_wall_bcs, int B, int D,
                         int H, int W, int is3d, void* stream) {
  return fnx_step_project_bcs_rows(pressure, U, flags, UBC, UBCInvMask, mask_rows, apply_wall_bcs, B, D, H, W, is3d, 0,
                                   0, stream);
}

int fnx_step_project_bcs_rows(const float* pressure, float* U, const float* flags, const float* UBC,
                              const float* UBCInvMask, const unsigned char* mask_rows, int apply_wall_bcs, int B,
                              int D, int H, int W, int is3d, int row_begin, int row_end, void* stream) {
  return fnx_step_project_bcs_held(pressure, U, flags, UBC, UBCInvMask, mask_rows, apply_wall_bcs, B, D, H, W, is3d,
                                   row_begin, row_end, 0, 0, stream);
}

int fnx_step_project_bcs_held(const float* pressure, float* U, const float* flags, const float* UBC,
                              const float* UBCInvMask, const unsigned char* mask_rows, int apply_wall_bcs, int B,
                              int D, int H, int W, int is3d, int row_begin, int row_end, int held_row_begin,
                              int held_row_end, void* stream) {
  if (B < 1 || D < 1 || H < 2 || W < 2 || (is3d && D < 2) || (!is3d && D != 1))
    return fnx_set_error(FNX_ERR_ARG, "step: unsupported grid B=%d D=%d H=%d W=%d is3d=%d", B, D, H, W, is3d);
  Grid g = make_grid(B, D, H, W);
  if (row_end > row_begin) {
    if (row_begin < 0 || row_end > D * H) return fnx_set_error(FNX_ERR_ARG, "step: row window out of range");
    g.row0 = row_begin; g.row1 = row_end;
  }
  const bool held = held_row_end > held_row_begin;
  if (held && (is3d || !fnx_step2d_supported(H, W)))
    return fnx_set_error(FNX_ERR_ARG, "step: row-window arrays (held_row_*) are a 2-D feature of the fused kernels");
  if (!is3d && fnx_step2d_supported(H, W) && (held || use_step2d((long long)(g.row1 - g.row0) * W))) {
    fnx_step2d_win w;
    w.H = H; w.W = W; w.row0 = g.row0; w.row1 = g.row1;
    w.ya0 = held ? held_row_begin : 0; w.ya1 = held ? held_row_end : H;
    if (!(0 <= w.ya0 && w.ya0 <= w.row0 && w.row1 <= w.ya1 && w.ya1 <= H) || (w.ya0 > 0 && w.row0 - w.ya0 < 1))
      return fnx_set_error(FNX_ERR_ARG, "step: bad row window");
    const bool ubc2 = UBC && UBCInvMask;
    fnx_step2d_masks mk;
    mk.UBC = ubc2 ? UBC : nullptr; mk.UBCInv = ubc2 ? UBCInvMask : nullptr;
    mk.rBC = nullptr; mk.rBCInv = nullptr; mk.rows = mask_rows;
    FNX_TRY(fnx_step2d_project(w, pressure, U, flags, mk, apply_wall_bcs, B, (cudaStream_t)stream));
    FNX_CUDA_TRY("step", cudaGetLastError());
    return FNX_OK;
  }
  StepMasks m;
  const bool ubc = UBC && UBCInvMask;
  m.UBC = ubc ? UBC : nullptr; m.UBCInv = ubc ? UBCInvMask : nullptr;
  m.rBC = nullptr; m.rBCInv = nullptr; m.rows = mask_rows;
  if (is3d) k_step_project<true><<<cell_grid(g), cell_block(), 0, (cudaStream_t)stream>>>(g, pressure, U, flags, m, apply_wall_bcs);
  else k_step_project<false><<<cell_grid(g), cell_block(), 0, (cudaStream_t)stream>>>(g, pressure, U, flags, m, apply_wall_bcs);
  fnx_count_launches(1);
  FNX_CUDA_TRY("step", cudaGetLastError());
  return FNX_OK;
}

int fnx_step_jacobi(const fnx_step_params* prm, const float* density_in, const float* U_in, const float* flags,
                    const float* UBC, const float* UBCInvMask, const float* densityBC, const float* densityBCInvMask,
                    const unsigned char* mask_rows, float* density_out, float* U_out, float* p, float* residual, int B,
                    int D, int H, int W, int is3d, void* workspace, size_t workspace_bytes, void* stream) {
  if (!prm) return fnx_set_error(FNX_ERR_ARG, "step: null params");
  if (!workspace || workspace_bytes < fnx_step_workspace(B, D, H, W, is3d))
    return fnx_set_error(FNX_ERR_WORKSPACE, "step: workspace too small");
  const size_t n = (size_t)B * D * H * W;
  const int nc = is3d ? 3 : 2;
  char* ws = (char*)workspace;
  ws += 2 * align256(n * sizeof(float)) + align256(n * sizeof(int)) + 2 * align256(n * nc * sizeof(float));
  float* div = (float*)ws_take(ws, n * sizeof(float));
  void* ws_j = ws;
  const size_t ws_j_bytes = align256(fnx_jacobi_workspace(B, D, H, W, 1));
  fnx_step_params q = *prm;
  q.row_begin = q.row_end = 0;
  q.apply_wall_bcs = 1;
  q.density_const_passes = 2;  // simulate.py:133 and :168 both re-apply the density BC
  FNX_TRY(fnx_step_advect_forces_div(&q, density_in, U_in, flags, UBC, UBCInvMask, densityBC, densityBCInvMask,
                                     mask_rows, density_out, U_out, div, B, D, H, W, is3d, workspace, workspace_bytes,
                                     stream));
  FNX_TRY(fnx_solve_linear_system_jacobi(flags, div, p, residual, B, D, H, W, is3d, 0.f, prm->jacobi_iters, nullptr,
                                         ws_j, ws_j_bytes, stream));
  FNX_TRY(fnx_step_project_bcs(p, U_out, flags, UBC, UBCInvMask, mask_rows, 1, B, D, H, W, is3d, stream));
  return FNX_OK;
}

}  // extern "C"
