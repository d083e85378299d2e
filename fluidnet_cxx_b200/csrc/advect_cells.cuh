// advect_cells.cuh -- thread->cell mapping and the per-cell bodies of the four advection passes
// (scalar fwd / bwd+correct+clamp, MAC fwd / bwd+correct+clamp).  Shared by the per-op kernels
// (stencils.cu) and the fused step kernels (step.cu).
#pragma once
#include "advect_device.cuh"
#include "fluid_common.cuh"
#include "stencil_device.cuh"

namespace fnx {

// thread -> cell mapping shared by all one-cell-per-thread kernels:
// threadIdx.x = W (coalesced), threadIdx.y = row; blockIdx.x = linear tile index, column tiles
// fastest (consecutive blocks touch consecutive memory; D*H row blocks can exceed the 65535 limit
// of gridDim.y, so rows are not a grid dimension of their own), blockIdx.z = batch
constexpr int kBX = 64, kBY = 4;

struct CellIdx {
  int b, k, j, i;
  long long o;  // offset inside one (D,H,W) volume
};

__device__ __forceinline__ bool cell_of(const Grid& g, CellIdx& c) {
  const unsigned wtiles = (unsigned)(g.W + kBX - 1) / kBX;
  const unsigned rb = blockIdx.x / wtiles, wt = blockIdx.x - rb * wtiles;
  c.i = wt * blockDim.x + threadIdx.x;
  int row = g.row0 + rb * blockDim.y + threadIdx.y;
  c.b = blockIdx.z;
  if (c.i >= g.W || row >= g.row1) return false;
  c.k = row / g.H;
  c.j = row - c.k * g.H;
  c.o = (long long)row * g.W + c.i;
  return true;
}

static inline dim3 cell_grid(const Grid& g) {
  return dim3(((g.row1 - g.row0 + kBY - 1) / kBY) * ((g.W + kBX - 1) / kBX), 1, g.B);
}
static inline dim3 cell_block() { return dim3(kBX, kBY, 1); }

// =====================================================================================
// advectScalar (fluids_init.cpp:265-382)
// =====================================================================================
// pass 1 (SemiLagrangeEulerFluidNetSavePos :69-133): fwd value + the cell index of the traced
// position (all MacCormackClampFluidNet :224-263 needs of it).
// (pointers already offset to the cell's batch item; returns the forward value, *fidx_out the
// cell index of the traced position)
template <bool Z>
__device__ __forceinline__ float scalar_fwd_cell(const Grid& g, const CellIdx& c, float mdt,
                                                 const float* __restrict__ src, const float* __restrict__ U,
                                                 const float* __restrict__ flags, int sample_outside,
                                                 bool want_idx, int* fidx_out) {
  constexpr int NA = Z ? 3 : 2;
  float val;
  long long idx = c.o;
  if (is_border<Z>(g, c.k, c.j, c.i)) {
    val = 0.f;
  } else if (__ldg(flags + c.o) != kFluid) {
    val = __ldg(src + c.o);  // don't advect solid geometry
  } else {
    float pos[3] = {(float)c.i + 0.5f, (float)c.j + 0.5f, (float)c.k + 0.5f};
    float vel[3], delta[3], back[3];
    centered_vel<Z>(g, U, c.o, vel);
#pragma unroll
    for (int a = 0; a < NA; a++) delta[a] = mdt * vel[a];
    line_trace<NA>(g, flags, pos, delta, back);
    val = sample_outside ? sample_field<Z>(g, src, back) : sample_with_fluid<Z>(g, src, flags, back);
    if (want_idx) {
      const int i0 = trunc_clamp0(back[0], g.W - 1);
      const int j0 = trunc_clamp0(back[1], g.H - 1);
      const int k0 = Z ? trunc_clamp0(back[2], g.D - 1) : 0;
      idx = (k0 * g.H + j0) * g.W + i0;
    }
  }
  *fidx_out = (int)idx;
  return val;
}

// pass 2-4: backward trace on `fwd`, MacCormackCorrect :135-148, clamp :154-263
template <bool Z>
__device__ __forceinline__ float scalar_bwd_cell(const Grid& g, const CellIdx& c, float dt, float half_strength,
                                                 const float* __restrict__ src, const float* __restrict__ U,
                                                 const float* __restrict__ flags, int sample_outside,
                                                 const float* __restrict__ fwd, const int* __restrict__ fidx) {
  constexpr int NA = Z ? 3 : 2;
  const bool border = is_border<Z>(g, c.k, c.j, c.i);
  const bool fluid = __ldg(flags + c.o) == kFluid;
  const float fw = __ldg(fwd + c.o);
  float v = fw;
  if (fluid) {
    float bwd = 0.f;  // border cells of the backward pass are zeroed (:354-363)
    if (!border) {
      float pos[3] = {(float)c.i + 0.5f, (float)c.j + 0.5f, (float)c.k + 0.5f};
      float vel[3], delta[3], back[3];
      centered_vel<Z>(g, U, c.o, vel);
#pragma unroll
      for (int a = 0; a < NA; a++) delta[a] = dt * vel[a];
      line_trace<NA>(g, flags, pos, delta, back);
      bwd = sample_outside ? sample_field<Z>(g, fwd, back) : sample_with_fluid<Z>(g, fwd, flags, back);
    }
    v = fw + half_strength * (__ldg(src + c.o) - bwd);
  }
  if (!border) {
    // getClampBounds :154-222: 3x3(x3) neighbourhood of the forward-traced cell, fluid cells only
    const int idx = __ldg(fidx + c.o);
    const int k0 = Z ? (int)(idx / g.sz) : 0;
    const int rem = (int)(idx - k0 * g.sz);
    const int j0 = rem / g.W, i0 = rem - j0 * g.W;
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    bool any = false;
#pragma unroll
    for (int dk = (Z ? -1 : 0); dk <= (Z ? 1 : 0); dk++) {
      const int kk = k0 + dk;
      if (Z && (kk < 0 || kk >= g.D)) continue;
#pragma unroll
      for (int dj = -1; dj <= 1; dj++) {
        const int jj = j0 + dj;
        if (jj < 0 || jj >= g.H) continue;
#pragma unroll
        for (int di = -1; di <= 1; di++) {
          const int ii = i0 + di;
          if (ii < 0 || ii >= g.W) continue;
          const int q = (kk * g.H + jj) * g.W + ii;
          if (sample_outside || __ldg(flags + q) == kFluid) {
            const float s = __ldg(src + q);
            mn = min_t(mn, s);
            mx = max_t(mx, s);
            any = true;
          }
        }
      }
    }
    v = any ? max_t(mn, min_t(mx, v)) : fw;
  }
  return v;
}

// =====================================================================================
// advectVel (fluids_init.cpp:656-807)
// =====================================================================================
// SemiLagrangeEulerFluidNetMAC :388-451 (no line trace, Q2; solid-cell quirk Q1)
template <bool Z>
__device__ __forceinline__ void vel_fwd_cell(const Grid& g, const CellIdx& c, float mdt,
                                             const float* __restrict__ orig, const float* __restrict__ U,
                                             const float* __restrict__ flags, float* out) {
  constexpr int NA = Z ? 3 : 2, NC = Z ? 3 : 2;
  out[0] = out[1] = out[2] = 0.f;
  if (!is_border<Z>(g, c.k, c.j, c.i)) {
    if (__ldg(flags + c.o) != kFluid) {
      if (!Z) { out[0] = __ldg(orig + g.n + c.o); out[1] = 0.f; }  // Q1
      else {
#pragma unroll
        for (int a = 0; a < NC; a++) out[a] = __ldg(orig + a * g.n + c.o);
      }
    } else {
      const float pos[3] = {(float)c.i + 0.5f, (float)c.j + 0.5f, (float)c.k + 0.5f};
#pragma unroll
      for (int comp = 0; comp < NC; comp++) {
        float v[3], p[3];
        mac_vel<Z>(g, U, comp, c.o, v);
#pragma unroll
        for (int a = 0; a < NA; a++) p[a] = pos[a] + v[a] * mdt;
        out[comp] = sample_field<Z>(g, orig + comp * g.n, p);
      }
    }
  }
}

// backward pass on `fwd`, MacCormackCorrectMAC :453-498, MacCormackClampMAC :500-654
template <bool Z>
__device__ __forceinline__ void vel_bwd_cell(const Grid& g, const CellIdx& c, float dt, float half_strength,
                                             const float* __restrict__ orig, const float* __restrict__ U,
                                             const float* __restrict__ flags, const float* __restrict__ fwd,
                                             float* out) {
  constexpr int NA = Z ? 3 : 2, NC = Z ? 3 : 2;
  out[0] = out[1] = out[2] = 0.f;
  if (is_border<Z>(g, c.k, c.j, c.i)) return;
  const bool solid = __ldg(flags + c.o) != kFluid;
  const float pos[3] = {(float)c.i + 0.5f, (float)c.j + 0.5f, (float)c.k + 0.5f};
  const float posi[3] = {(float)c.i, (float)c.j, (float)c.k};
  const int idx[3] = {c.i, c.j, c.k};
#pragma unroll
  for (int comp = 0; comp < NC; comp++) {
    float vel[3];
    mac_vel<Z>(g, U, comp, c.o, vel);
    const float fw = __ldg(fwd + comp * g.n + c.o);
    // correction skipped when the cell or its lower neighbour along `comp` is not fluid
    bool skip = solid;
    if (!skip && idx[comp] > 0 && __ldg(flags + c.o - nb_off(g, comp)) != kFluid) skip = true;
    float v = fw;
    if (!skip) {
      float p[3];
#pragma unroll
      for (int a = 0; a < NA; a++) p[a] = pos[a] + vel[a] * dt;
      const float bwd = sample_field<Z>(g, fwd + comp * g.n, p);
      v = fw + half_strength * (__ldg(orig + comp * g.n + c.o) - bwd);
    }
    // doClampComponentMAC: min/max of orig over the 2x2(x2) blocks at trunc(pos -/+ vel*dt), Q5
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    const float* oc = orig + comp * g.n;
#pragma unroll
    for (int l = 0; l < 2; l++) {
      const int hi[3] = {g.W - 2, g.H - 2, g.D - 2};
      int q[3] = {0, 0, 0};
#pragma unroll
      for (int a = 0; a < NA; a++) {
        const float va = vel[a] * dt;
        q[a] = trunc_x86_clamp0(l == 0 ? posi[a] - va : posi[a] + va, hi[a]);
      }
      const float* b0 = oc + (q[2] * g.H + q[1]) * g.W + q[0];
      // same visiting order as the reference: (j0,i0) (j0,i0+1) (j0+1,i0) (j0+1,i0+1)
      float s;
      s = __ldg(b0); mn = min_t(mn, s); mx = max_t(mx, s);
      s = __ldg(b0 + 1); mn = min_t(mn, s); mx = max_t(mx, s);
      s = __ldg(b0 + g.sy); mn = min_t(mn, s); mx = max_t(mx, s);
      s = __ldg(b0 + g.sy + 1); mn = min_t(mn, s); mx = max_t(mx, s);
      if (Z) {
        const float* b1 = b0 + g.sz;
        s = __ldg(b1); mn = min_t(mn, s); mx = max_t(mx, s);
        s = __ldg(b1 + 1); mn = min_t(mn, s); mx = max_t(mx, s);
        s = __ldg(b1 + g.sy); mn = min_t(mn, s); mx = max_t(mx, s);
        s = __ldg(b1 + g.sy + 1); mn = min_t(mn, s); mx = max_t(mx, s);
      }
    }
    out[comp] = max_t(min_t(v, mx), mn);
  }
}


}  // namespace fnx
