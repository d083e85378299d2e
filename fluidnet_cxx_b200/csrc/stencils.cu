// stencils.cu -- sm_100a kernels + C-ABI for the per-op entry points of the fluid path
// (advection, forces, wall BCs, divergence, Jacobi, velocity update).  See include/fluidstep.h.
// Compiled with -fmad=false (bit-exact operation order).
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "../../include/fluidstep.h"
#include "advect_cells.cuh"
#include "advect_device.cuh"
#include "fluid_common.cuh"
#include "host_util.h"
#include "stencil_device.cuh"

namespace fnx {

// =====================================================================================
// advection kernels (per-cell bodies in advect_cells.cuh)
// =====================================================================================
template <bool Z>
__global__ void __launch_bounds__(kBX* kBY)
    k_advect_scalar_fwd(Grid g, float mdt, const float* __restrict__ src, const float* __restrict__ U,
                        const float* __restrict__ flags, int sample_outside, float* __restrict__ fwd,
                        int* __restrict__ fidx) {
  constexpr int NC = Z ? 3 : 2;
  CellIdx c;
  if (!cell_of(g, c)) return;
  const long long bo = (long long)c.b * g.n;
  int idx;
  fwd[bo + c.o] = scalar_fwd_cell<Z>(g, c, mdt, src + bo, U + bo * NC, flags + bo, sample_outside, fidx != nullptr, &idx);
  if (fidx) fidx[bo + c.o] = idx;
}

template <bool Z>
__global__ void __launch_bounds__(kBX* kBY)
    k_advect_scalar_bwd(Grid g, float dt, float half_strength, const float* __restrict__ src,
                        const float* __restrict__ U, const float* __restrict__ flags, int sample_outside,
                        const float* __restrict__ fwd, const int* __restrict__ fidx, float* __restrict__ dst) {
  constexpr int NC = Z ? 3 : 2;
  CellIdx c;
  if (!cell_of(g, c)) return;
  const long long bo = (long long)c.b * g.n;
  dst[bo + c.o] = scalar_bwd_cell<Z>(g, c, dt, half_strength, src + bo, U + bo * NC, flags + bo, sample_outside,
                                     fwd + bo, fidx + bo);
}

template <bool Z>
__global__ void __launch_bounds__(kBX* kBY)
    k_advect_vel_fwd(Grid g, float mdt, const float* __restrict__ orig, const float* __restrict__ U,
                     const float* __restrict__ flags, float* __restrict__ fwd) {
  constexpr int NC = Z ? 3 : 2;
  CellIdx c;
  if (!cell_of(g, c)) return;
  const long long bo = (long long)c.b * g.n;
  float out[3];
  vel_fwd_cell<Z>(g, c, mdt, orig + bo * NC, U + bo * NC, flags + bo, out);
#pragma unroll
  for (int a = 0; a < NC; a++) fwd[bo * NC + (long long)a * g.n + c.o] = out[a];
}

template <bool Z>
__global__ void __launch_bounds__(kBX* kBY)
    k_advect_vel_bwd(Grid g, float dt, float half_strength, const float* __restrict__ orig,
                     const float* __restrict__ U, const float* __restrict__ flags,
                     const float* __restrict__ fwd, float* __restrict__ dst) {
  constexpr int NC = Z ? 3 : 2;
  CellIdx c;
  if (!cell_of(g, c)) return;
  const long long bo = (long long)c.b * g.n;
  float out[3];
  vel_bwd_cell<Z>(g, c, dt, half_strength, orig + bo * NC, U + bo * NC, flags + bo, fwd + bo * NC, out);
#pragma unroll
  for (int a = 0; a < NC; a++) dst[bo * NC + (long long)a * g.n + c.o] = out[a];
}

// =====================================================================================
// small stencils (per-op kernels of the lib.fluid surface)
// =====================================================================================
template <bool Z>
__global__ void __launch_bounds__(kBX* kBY)
    k_add_buoyancy(Grid g, float* __restrict__ U, const float* __restrict__ flags,
                   const float* __restrict__ density, float3 strength, float rho_star) {
  constexpr int NC = Z ? 3 : 2;
  CellIdx c;
  if (!cell_of(g, c)) return;
  if (is_border<Z>(g, c.k, c.j, c.i)) return;
  flags += c.b * g.n; density += c.b * g.n; U += (long long)c.b * NC * g.n;
  const float fc = __ldg(flags + c.o);
  if (fc != kFluid) return;
  const float rc = __ldg(density + c.o);
  const float st[3] = {strength.x, strength.y, strength.z};
#pragma unroll
  for (int a = 0; a < NC; a++) {
    const long long on = c.o - nb_off(g, a);
    float* u = U + a * g.n + c.o;
    *u = buoyancy_apply(*u, fc, __ldg(flags + on), rc, __ldg(density + on), st[a], rho_star);
  }
}

template <bool Z>
__global__ void __launch_bounds__(kBX* kBY)
    k_add_gravity(Grid g, float* __restrict__ U, const float* __restrict__ flags, float3 force) {
  constexpr int NC = Z ? 3 : 2;
  CellIdx c;
  if (!cell_of(g, c)) return;
  if (is_border<Z>(g, c.k, c.j, c.i)) return;
  flags += c.b * g.n; U += (long long)c.b * NC * g.n;
  const float fc = __ldg(flags + c.o);
  const float fo[3] = {force.x, force.y, force.z};
#pragma unroll
  for (int a = 0; a < NC; a++) {
    float* u = U + a * g.n + c.o;
    *u = gravity_apply(*u, fc, __ldg(flags + c.o - nb_off(g, a)), fo[a]);
  }
}

// viscosity.py:7-70 addViscosity (2-D; the reference's 3-D branch uses an undefined mask): interior
// component c = mask_c * (u + s*((((u_E + u_N) + u_W) + u_SW) - 4u)), mask_c = cell and its lower
// neighbour along c are Fluid; u_SW = U(i-1, j-1) as written at :68.  Reads `Uin`, writes `Uout`
// (the reference evaluates the whole right-hand side before assigning).
__global__ void __launch_bounds__(kBX* kBY)
    k_add_viscosity(Grid g, const float* __restrict__ Uin, float* __restrict__ Uout, const float* __restrict__ flags,
                    float s) {
  CellIdx c;
  if (!cell_of(g, c)) return;
  flags += c.b * g.n; Uin += (long long)c.b * 2 * g.n; Uout += (long long)c.b * 2 * g.n;
  if (is_border<false>(g, c.k, c.j, c.i)) {
#pragma unroll
    for (int a = 0; a < 2; a++) Uout[a * g.n + c.o] = __ldg(Uin + a * g.n + c.o);
    return;
  }
  const bool fl = __ldg(flags + c.o) == kFluid;
  const float m[2] = {(fl && __ldg(flags + c.o - 1) == kFluid) ? 1.f : 0.f,
                      (fl && __ldg(flags + c.o - g.sy) == kFluid) ? 1.f : 0.f};
#pragma unroll
  for (int a = 0; a < 2; a++) {
    const float* q = Uin + a * g.n + c.o;
    const float lap = (((__ldg(q + 1) + __ldg(q + g.sy)) + __ldg(q - 1)) + __ldg(q - g.sy - 1)) - (4.f * __ldg(q));
    Uout[a * g.n + c.o] = m[a] * (__ldg(q) + s * lap);
  }
}

// advection.py:9-12 correctScalar: src += (t*src)*div on Fluid cells, t = (float)(dt*0.5)
__global__ void __launch_bounds__(256)
    k_correct_scalar(float* __restrict__ src, const float* __restrict__ div, const float* __restrict__ flags, float t,
                     size_t count) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= count) return;
  if (__ldg(flags + q) == kFluid) {
    const float v = src[q];
    src[q] = v + (t * v) * __ldg(div + q);
  }
}

template <bool Z>
__global__ void __launch_bounds__(kBX* kBY)
    k_set_wall_bcs(Grid g, float* __restrict__ U, const float* __restrict__ flags) {
  constexpr int NC = Z ? 3 : 2;
  CellIdx c;
  if (!cell_of(g, c)) return;
  flags += c.b * g.n; U += (long long)c.b * NC * g.n;
  const float fc = __ldg(flags + c.o);
  const int idx[3] = {c.i, c.j, c.k};
#pragma unroll
  for (int a = 0; a < NC; a++) {
    const float fn = idx[a] <= 0 ? fc : __ldg(flags + c.o - nb_off(g, a));
    float* u = U + a * g.n + c.o;
    *u = wall_bcs_apply(*u, fc, fn);
  }
}

template <bool Z>
__global__ void __launch_bounds__(kBX* kBY)
    k_velocity_divergence(Grid g, const float* __restrict__ U, const float* __restrict__ flags,
                          float* __restrict__ div) {
  constexpr int NC = Z ? 3 : 2;
  CellIdx c;
  if (!cell_of(g, c)) return;
  flags += c.b * g.n; U += (long long)c.b * NC * g.n; div += c.b * g.n;
  float v = 0.f;
  if (!is_border<Z>(g, c.k, c.j, c.i)) {
    // (u_i - u_{i+1}) + v_j - v_{j+1} left to right (velocity_divergence.py:61-69)
    v = __ldg(U + c.o) - __ldg(U + c.o + 1) + __ldg(U + g.n + c.o) - __ldg(U + g.n + c.o + g.sy);
    if (Z) v = v + (__ldg(U + 2 * g.n + c.o) - __ldg(U + 2 * g.n + c.o + g.sz));
  }
  if (__ldg(flags + c.o) == kObstacle) v = 0.f;
  div[c.o] = v;
}

template <bool Z>
__global__ void __launch_bounds__(kBX* kBY)
    k_velocity_update(Grid g, const float* __restrict__ p, float* __restrict__ U,
                      const float* __restrict__ flags) {
  constexpr int NC = Z ? 3 : 2;
  CellIdx c;
  if (!cell_of(g, c)) return;
  if (is_border<Z>(g, c.k, c.j, c.i)) return;
  flags += c.b * g.n; p += c.b * g.n; U += (long long)c.b * NC * g.n;
  const float fc = __ldg(flags + c.o), P = __ldg(p + c.o);
#pragma unroll
  for (int a = 0; a < NC; a++) {
    const long long on = c.o - nb_off(g, a);
    float* u = U + a * g.n + c.o;
    *u = velocity_update_apply(*u, fc, __ldg(flags + on), P, __ldg(p + on));
  }
}

__global__ void k_flags_to_occupancy(const float* __restrict__ flags, float* __restrict__ occ,
                                     size_t count) {
  size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < count) occ[q] = occupancy_of(__ldg(flags + q));
}

__global__ void k_set_const_vals(float* __restrict__ x, const float* __restrict__ inv_mask,
                                 const float* __restrict__ bc, size_t count) {
  size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < count) x[q] = const_vals_apply(x[q], __ldg(inv_mask + q), __ldg(bc + q));
}

// The output step of the drivers (plume.py:238-263 plots, :330-423 VTK dump) as ONE pass: divergence
// (velocity_divergence.py:4-74), cell-centred velocity (grid.py:7-30) and its norm, the centred density and
// pressure gradients of the VTK block (plume.py:343-359, 2-D as there), pressure; Obstacle cells of the velocity,
// norm and pressure planes filled with NaN when `mask` (the drivers' numpy masked_array .filled(nan)).
// out: (B, FNX_OUTPUT_PLANES, D, H, W).  The reference forms these with ~60 tensor ops and 9 device->host copies;
// here one kernel writes one buffer that goes to the host in one copy.
template <bool Z>
__global__ void __launch_bounds__(kBX* kBY)
    k_output_fields(Grid g, const float* __restrict__ U, const float* __restrict__ flags, const float* __restrict__ rho,
                    const float* __restrict__ p, float* __restrict__ out, int mask) {
  constexpr int NC = Z ? 3 : 2;
  CellIdx c;
  if (!cell_of(g, c)) return;
  U += (long long)c.b * NC * g.n; flags += c.b * g.n; out += (long long)c.b * FNX_OUTPUT_PLANES * g.n;
  if (rho) rho += c.b * g.n;
  if (p) p += c.b * g.n;
  const float f = __ldg(flags + c.o);
  const float ux = __ldg(U + c.o), uy = __ldg(U + g.n + c.o), uz = Z ? __ldg(U + 2 * g.n + c.o) : 0.f;
  const float ux1 = c.i < g.W - 1 ? __ldg(U + c.o + 1) : 0.f;
  const float uy1 = c.j < g.H - 1 ? __ldg(U + g.n + c.o + g.sy) : 0.f;
  const float uz1 = (Z && c.k < g.D - 1) ? __ldg(U + 2 * g.n + c.o + g.sz) : 0.f;
  float dv = 0.f;
  if (!is_border<Z>(g, c.k, c.j, c.i)) {
    dv = ux - ux1 + uy - uy1;
    if (Z) dv = dv + (uz - uz1);
  }
  if (f == kObstacle) dv = 0.f;
  float cx = 0.f, cy = 0.f, cz = 0.f;
  if (c.i < g.W - 1) cx = 0.5f * (ux + ux1);
  if (c.j < g.H - 1) cy = 0.5f * (uy + uy1);
  if (Z && c.k < g.D - 1) cz = 0.5f * (uz + uz1);
  float nrm = sqrtf(cx * cx + cy * cy + cz * cz);
  // centred gradients: getCentered of the face differences of the (H-2) x (W-2) interior, placed at [1, H-1) x [1, W-1)
  float grx = 0.f, gry = 0.f, gpx = 0.f, gpy = 0.f;
  if (!Z && c.j >= 1 && c.j <= g.H - 2 && c.i >= 1 && c.i <= g.W - 2) {
    if (c.i < g.W - 2) {
      if (rho) grx = 0.5f * ((__ldg(rho + c.o) - __ldg(rho + c.o - 1)) + (__ldg(rho + c.o + 1) - __ldg(rho + c.o)));
      if (p) gpx = 0.5f * ((__ldg(p + c.o) - __ldg(p + c.o - 1)) + (__ldg(p + c.o + 1) - __ldg(p + c.o)));
    }
    if (c.j < g.H - 2) {
      if (rho) gry = 0.5f * ((__ldg(rho + c.o) - __ldg(rho + c.o - g.sy)) + (__ldg(rho + c.o + g.sy) - __ldg(rho + c.o)));
      if (p) gpy = 0.5f * ((__ldg(p + c.o) - __ldg(p + c.o - g.sy)) + (__ldg(p + c.o + g.sy) - __ldg(p + c.o)));
    }
  }
  float pm = p ? __ldg(p + c.o) : 0.f;
  if (mask && f == kObstacle) {
    const float qnan = __int_as_float(0x7fc00000);
    cx = cy = cz = nrm = pm = qnan;
  }
  out[c.o] = dv;
  out[1 * g.n + c.o] = cx; out[2 * g.n + c.o] = cy; out[3 * g.n + c.o] = cz; out[4 * g.n + c.o] = nrm;
  out[5 * g.n + c.o] = grx; out[6 * g.n + c.o] = gry; out[7 * g.n + c.o] = gpx; out[8 * g.n + c.o] = gpy;
  out[9 * g.n + c.o] = pm;
}

template <bool Z>
__global__ void __launch_bounds__(kBX* kBY) k_empty_domain(Grid g, float* __restrict__ flags, int bnd) {
  CellIdx c;
  if (!cell_of(g, c)) return;
  bool m = (c.i < bnd) | (c.i > g.W - 1 - bnd) | (c.j < bnd) | (c.j > g.H - 1 - bnd);
  if (Z) m = m | (c.k < bnd) | (c.k > g.D - 1 - bnd);
  flags[c.b * g.n + c.o] = m ? kObstacle : kFluid;
}

// grid.py:7-30 getCentered (output helper: last row/column/plane stay 0)
template <bool Z>
__global__ void __launch_bounds__(kBX* kBY)
    k_get_centered(Grid g, const float* __restrict__ U, float* __restrict__ out) {
  constexpr int NC = Z ? 3 : 2;
  CellIdx c;
  if (!cell_of(g, c)) return;
  U += (long long)c.b * NC * g.n; out += (long long)c.b * 3 * g.n;
  float x = 0.f, y = 0.f, z = 0.f;
  if (c.i < g.W - 1) x = 0.5f * (__ldg(U + c.o) + __ldg(U + c.o + 1));
  if (c.j < g.H - 1) y = 0.5f * (__ldg(U + g.n + c.o) + __ldg(U + g.n + c.o + g.sy));
  if (Z && c.k < g.D - 1) z = 0.5f * (__ldg(U + 2 * g.n + c.o) + __ldg(U + 2 * g.n + c.o + g.sz));
  out[c.o] = x; out[g.n + c.o] = y; out[2 * g.n + c.o] = z;
}

// =====================================================================================
// Jacobi (fluids_init.cpp:809-1004), one iteration per launch (generic path: p_tol > 0, 3-D, ...)
// =====================================================================================
struct JacobiCtrl {
  int done;        // set when residual < p_tol
  int iters;       // iterations executed so far
  float residual;  // max_b ||p - p_prev||_2 of the last executed iteration
  int pad;
};

template <bool Z, bool FIRST, bool RESID>
__global__ void __launch_bounds__(kBX* kBY)
    k_jacobi_iter(Grid g, const float* __restrict__ flags, const float* __restrict__ div,
                  const float* __restrict__ prev, float* __restrict__ cur, double* __restrict__ ssq,
                  const JacobiCtrl* __restrict__ ctrl) {
  if (ctrl && ctrl->done) return;
  CellIdx c;
  const bool valid = cell_of(g, c);
  float d2 = 0.f;
  if (valid) {
    flags += c.b * g.n; div += c.b * g.n; cur += c.b * g.n;
    if (!FIRST) prev += c.b * g.n;
    float pn = 0.f;
    const float pC = FIRST ? 0.f : __ldg(prev + c.o);
    if (!(is_border<Z>(g, c.k, c.j, c.i) || __ldg(flags + c.o) == kObstacle)) {
      float s;
      if (FIRST) {
        s = __ldg(div + c.o);  // p0 = 0: every neighbour term is 0
      } else {
        // Neumann: an Obstacle neighbour contributes the centre value (:895-943)
        const float p1 = __ldg(flags + c.o - 1) == kObstacle ? pC : __ldg(prev + c.o - 1);
        const float p2 = __ldg(flags + c.o + 1) == kObstacle ? pC : __ldg(prev + c.o + 1);
        const float p3 = __ldg(flags + c.o - g.sy) == kObstacle ? pC : __ldg(prev + c.o - g.sy);
        const float p4 = __ldg(flags + c.o + g.sy) == kObstacle ? pC : __ldg(prev + c.o + g.sy);
        s = p1 + p2 + p3 + p4;
        if (Z) {
          const float p5 = __ldg(flags + c.o - g.sz) == kObstacle ? pC : __ldg(prev + c.o - g.sz);
          const float p6 = __ldg(flags + c.o + g.sz) == kObstacle ? pC : __ldg(prev + c.o + g.sz);
          s = s + p5 + p6;
        }
        s = s + __ldg(div + c.o);
      }
      pn = Z ? s / 6.f : s * 0.25f;  // /4 is exact as *0.25
    }
    cur[c.o] = pn;
    const float d = pn - pC;
    const int row = c.k * g.H + c.j;
    d2 = (row >= g.own0 && row < g.own1) ? d * d : 0.f;
  }
  if (RESID) {
    // block reduction of sum((p - p_prev)^2) in double, one atomic per block
    double acc = (double)d2;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    __shared__ double wsum[kBX * kBY / 32];
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if ((tid & 31) == 0) wsum[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < kBX * kBY / 32; w++) t += wsum[w];
      atomicAdd(ssq + blockIdx.z, t);
    }
  }
}

// ---- 3-D fixed-count path: neighbour masks precomputed once per solve, 4 cells per thread --------
// The flags never change during a solve: one byte per cell replaces the 7 fp32 flag loads of every
// iteration.  bit0 = p is pinned to 0 (border ring or Obstacle), bits 1..6 = the -x, +x, -y, +y, -z,
// +z neighbour is an Obstacle (Neumann: it contributes the centre value, fluids_init.cpp:895-943).
__global__ void __launch_bounds__(256)
    k_jacobi3d_mask(Grid g, const float* __restrict__ flags, unsigned char* __restrict__ mask) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long long)g.B * g.n) return;
  const long long o = q % g.n;
  const int i = (int)(o % g.W), j = (int)((o / g.W) % g.H), k = (int)(o / g.sz);
  const float* f = flags + (q - o);
  unsigned m = 0;
  if (is_border<true>(g, k, j, i) || __ldg(f + o) == kObstacle) {
    m = 1;
  } else {
    m |= (__ldg(f + o - 1) == kObstacle) << 1;
    m |= (__ldg(f + o + 1) == kObstacle) << 2;
    m |= (__ldg(f + o - g.sy) == kObstacle) << 3;
    m |= (__ldg(f + o + g.sy) == kObstacle) << 4;
    m |= (__ldg(f + o - g.sz) == kObstacle) << 5;
    m |= (__ldg(f + o + g.sz) == kObstacle) << 6;
  }
  mask[q] = (unsigned char)m;
}

// one iteration, a thread owns 4 consecutive x cells (W % 4 == 0): float4 loads of the centre row and of
// the four y / z neighbour rows, two scalar loads for the x neighbours of the group's ends.  Per-cell
// arithmetic = k_jacobi_iter<true>: ((((p1+p2)+p3)+p4)+p5)+p6, + div, / 6.
template <bool FIRST, bool RESID>
__global__ void __launch_bounds__(256)
    k_jacobi3d_vec(Grid g, const unsigned char* __restrict__ mask, const float* __restrict__ div,
                   const float* __restrict__ prev, float* __restrict__ cur, double* __restrict__ ssq) {
  const int w4 = g.W >> 2;
  const long long groups_per_batch = (long long)(g.row1 - g.row0) * w4;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  float d2 = 0.f;
  if (t < groups_per_batch) {
    const int row = g.row0 + (int)(t / w4), i0 = (int)(t % w4) * 4;
    const long long o = (long long)row * g.W + i0;
    mask += (long long)b * g.n; div += (long long)b * g.n; cur += (long long)b * g.n;
    if (!FIRST) prev += (long long)b * g.n;
    const uchar4 m4 = *reinterpret_cast<const uchar4*>(mask + o);
    const unsigned mm[4] = {m4.x, m4.y, m4.z, m4.w};
    float pc[4] = {0.f, 0.f, 0.f, 0.f}, pn[4] = {0.f, 0.f, 0.f, 0.f};
    if (!FIRST) {
      const float4 c4 = __ldg(reinterpret_cast<const float4*>(prev + o));
      pc[0] = c4.x; pc[1] = c4.y; pc[2] = c4.z; pc[3] = c4.w;
    }
    if (!(mm[0] & mm[1] & mm[2] & mm[3] & 1u)) {  // at least one free cell: an interior row and plane
      const float4 d4 = __ldg(reinterpret_cast<const float4*>(div + o));
      const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
      if (FIRST) {
#pragma unroll
        for (int c = 0; c < 4; c++) pn[c] = (mm[c] & 1u) ? 0.f : dv[c] / 6.f;  // p0 = 0: only div remains
      } else {
        const float4 u4 = __ldg(reinterpret_cast<const float4*>(prev + o - g.sy));
        const float4 s4 = __ldg(reinterpret_cast<const float4*>(prev + o + g.sy));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(prev + o - g.sz));
        const float4 f4 = __ldg(reinterpret_cast<const float4*>(prev + o + g.sz));
        const float pl = i0 > 0 ? __ldg(prev + o - 1) : 0.f, pr = i0 + 4 < g.W ? __ldg(prev + o + 4) : 0.f;
        const float xm[4] = {pl, pc[0], pc[1], pc[2]}, xp[4] = {pc[1], pc[2], pc[3], pr};
        const float ym[4] = {u4.x, u4.y, u4.z, u4.w}, yp[4] = {s4.x, s4.y, s4.z, s4.w};
        const float zm[4] = {b4.x, b4.y, b4.z, b4.w}, zp[4] = {f4.x, f4.y, f4.z, f4.w};
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const unsigned m = mm[c];
          if (m & 1u) continue;
          const float pC = pc[c];
          const float p1 = (m & 2u) ? pC : xm[c], p2 = (m & 4u) ? pC : xp[c];
          const float p3 = (m & 8u) ? pC : ym[c], p4 = (m & 16u) ? pC : yp[c];
          const float p5 = (m & 32u) ? pC : zm[c], p6 = (m & 64u) ? pC : zp[c];
          float sum = p1 + p2 + p3 + p4;
          sum = sum + p5 + p6;
          sum = sum + dv[c];
          pn[c] = sum / 6.f;
        }
      }
    }
    *reinterpret_cast<float4*>(cur + o) = make_float4(pn[0], pn[1], pn[2], pn[3]);
    if (RESID) {
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const float d = pn[c] - pc[c];
        d2 += d * d;  // (the scalar kernel sums one square per thread; the total is formed in double below)
      }
    }
  }
  if (RESID) {
    double acc = (double)d2;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    __shared__ double wsum[8];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tt = 0.0;
#pragma unroll
      for (int w = 0; w < 8; w++) tt += wsum[w];
      atomicAdd(ssq + b, tt);
    }
  }
}

// one thread: fold ssq[b] into the residual, test the tolerance, re-arm ssq
__global__ void k_jacobi_ctrl(JacobiCtrl* ctrl, double* ssq, int B, float p_tol, int iter_index,
                              float* residual_out) {
  if (ctrl->done) return;
  float r = 0.f;
  for (int b = 0; b < B; b++) {
    float rb = (float)sqrt(ssq[b]);
    if (rb > r) r = rb;
    ssq[b] = 0.0;
  }
  ctrl->residual = r;
  ctrl->iters = iter_index + 1;
  if (residual_out) *residual_out = r;
  if (r < p_tol) ctrl->done = 1;
}

}  // namespace fnx

// =====================================================================================
// C-ABI
// =====================================================================================
using namespace fnx;

static int check_grid(int B, int D, int H, int W, int is3d, const char* who) {
  if (B < 1 || D < 1 || H < 2 || W < 2 || (is3d && D < 2) || (!is3d && D != 1)) {
    return fnx_set_error(FNX_ERR_ARG, "%s: unsupported grid B=%d D=%d H=%d W=%d is3d=%d", who, B, D, H, W, is3d);
  }
  if ((long long)D * H * W >= (1LL << 31)) return fnx_set_error(FNX_ERR_ARG, "%s: grid too large for int32 cell index", who);
  return FNX_OK;
}

#define FNX_LAUNCH_CHECK(who, nlaunch)                                              \
  do {                                                                              \
    fnx_count_launches(nlaunch);                                                    \
    cudaError_t e_ = cudaGetLastError();                                            \
    if (e_ != cudaSuccess) return fnx_set_error(FNX_ERR_CUDA, "%s: %s", who, cudaGetErrorString(e_)); \
  } while (0)

#define FNX_DISPATCH_3D(is3d, KERNEL, ...)            \
  do {                                                \
    if (is3d) KERNEL<true> __VA_ARGS__;               \
    else KERNEL<false> __VA_ARGS__;                   \
  } while (0)

template <bool Z>
static void launch_jacobi_iter(const Grid& g, bool first, bool resid, const float* flags, const float* div,
                               const float* prev, float* cur, double* ssq, const JacobiCtrl* ctrl,
                               cudaStream_t st) {
  dim3 gr = cell_grid(g), bl = cell_block();
  if (first) {
    if (resid) k_jacobi_iter<Z, true, true><<<gr, bl, 0, st>>>(g, flags, div, prev, cur, ssq, ctrl);
    else k_jacobi_iter<Z, true, false><<<gr, bl, 0, st>>>(g, flags, div, prev, cur, ssq, ctrl);
  } else {
    if (resid) k_jacobi_iter<Z, false, true><<<gr, bl, 0, st>>>(g, flags, div, prev, cur, ssq, ctrl);
    else k_jacobi_iter<Z, false, false><<<gr, bl, 0, st>>>(g, flags, div, prev, cur, ssq, ctrl);
  }
}

int fnx_jacobi_2d_blocked(const float* flags, const float* div, const float* p_init, float* p, float* scratch,
                          double* ssq, int B, int H, int W, int max_iter, int row0, int row1, void* tile_ws,
                          size_t tile_ws_bytes, cudaStream_t st);  // jacobi_blocked.cu
int fnx_jacobi_2d_blocked_held(const float* flags, const float* div, const float* p_init, float* p, float* scratch,
                               double* ssq, int B, int H, int W, int max_iter, int row0, int row1, int ya0, int ya1,
                               void* tile_ws, size_t tile_ws_bytes, cudaStream_t st);  // jacobi_blocked.cu
size_t fnx_jacobi_2d_tilemask_bytes(int B, int H, int W);  // jacobi_blocked.cu
int fnx_jacobi_2d_blocked_masks(const float* flags, const float* div, const float* p_init, float* p, float* scratch,
                                int B, int H, int W, int max_iter, int row0, int row1, int ya0, int ya1, void* tile_ws,
                                size_t tile_ws_bytes, int mask_mode, cudaStream_t st);  // jacobi_blocked.cu


// 3-D fixed iteration count: mask once, then the 4-cells-per-thread kernel (needs W % 4 == 0 and 16-byte
// aligned fields; otherwise the caller falls back to k_jacobi_iter).  p_init == nullptr: start from 0.
static bool jacobi3d_vec_ok(const Grid& g, const float* a, const float* b, const float* c, const float* d) {
  return (g.W % 4 == 0) && ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d) & 15) == 0);
}

static int jacobi3d_vec_run(const Grid& g, const float* flags, const float* div, const float* p_init, float* p,
                            float* scratch, unsigned char* mask, double* ssq, int iters, cudaStream_t st) {
  const long long total = (long long)g.B * g.n;
  k_jacobi3d_mask<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g, flags, mask);
  const long long groups = (long long)(g.row1 - g.row0) * (g.W >> 2);
  dim3 grid((unsigned)((groups + 255) / 256), g.B);
  auto wbuf = [&](int it) { return ((iters - 1 - it) % 2 == 0) ? p : scratch; };
  for (int it = 0; it < iters; it++) {
    const float* prev = it == 0 ? p_init : wbuf(it - 1);
    const bool first = prev == nullptr, resid = ssq != nullptr && it == iters - 1;
    if (first && resid) k_jacobi3d_vec<true, true><<<grid, 256, 0, st>>>(g, mask, div, prev, wbuf(it), ssq);
    else if (first) k_jacobi3d_vec<true, false><<<grid, 256, 0, st>>>(g, mask, div, prev, wbuf(it), ssq);
    else if (resid) k_jacobi3d_vec<false, true><<<grid, 256, 0, st>>>(g, mask, div, prev, wbuf(it), ssq);
    else k_jacobi3d_vec<false, false><<<grid, 256, 0, st>>>(g, mask, div, prev, wbuf(it), ssq);
  }
  fnx_count_launches(iters + 1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fnx_set_error(FNX_ERR_CUDA, "jacobi3d: %s", cudaGetErrorString(e));
  return FNX_OK;
}

extern "C" {

size_t fnx_advect_scalar_workspace(int B, int D, int H, int W) {
  size_t n = (size_t)B * D * H * W;
  return n * sizeof(float) + n * sizeof(int);
}

int fnx_advect_scalar(float dt, const float* src, const float* U, const float* flags, float* dst, int B,
                      int D, int H, int W, int is3d, int method, int boundary_width,
                      int sample_outside_fluid, float maccormack_strength, void* workspace,
                      size_t workspace_bytes, void* stream) {
  if (int e = check_grid(B, D, H, W, is3d, "advect_scalar")) return e;
  if (boundary_width != 1) return fnx_set_error(FNX_ERR_ARG, "advect_scalar: boundary_width must be 1 (Q4)");
  if (method != FNX_METHOD_EULER && method != FNX_METHOD_MACCORMACK)
    return fnx_set_error(FNX_ERR_ARG, "advect_scalar: No defined method for MacCormackClamp");
  cudaStream_t st = (cudaStream_t)stream;
  Grid g = make_grid(B, D, H, W);
  if (method == FNX_METHOD_EULER) {
    FNX_DISPATCH_3D(is3d, k_advect_scalar_fwd, <<<cell_grid(g), cell_block(), 0, st>>>(
        g, -dt, src, U, flags, sample_outside_fluid, dst, nullptr));
    FNX_LAUNCH_CHECK("advect_scalar", 1);
    return FNX_OK;
  }
  if (workspace_bytes < fnx_advect_scalar_workspace(B, D, H, W) || !workspace)
    return fnx_set_error(FNX_ERR_WORKSPACE, "advect_scalar: workspace too small");
  float* fwd = (float*)workspace;
  int* fidx = (int*)(fwd + (size_t)B * g.n);
  FNX_DISPATCH_3D(is3d, k_advect_scalar_fwd, <<<cell_grid(g), cell_block(), 0, st>>>(
      g, -dt, src, U, flags, sample_outside_fluid, fwd, fidx));
  FNX_DISPATCH_3D(is3d, k_advect_scalar_bwd, <<<cell_grid(g), cell_block(), 0, st>>>(
      g, dt, maccormack_strength * 0.5f, src, U, flags, sample_outside_fluid, fwd, fidx, dst));
  FNX_LAUNCH_CHECK("advect_scalar", 2);
  return FNX_OK;
}

size_t fnx_advect_vel_workspace(int B, int D, int H, int W, int is3d) {
  return (size_t)B * D * H * W * (is3d ? 3 : 2) * sizeof(float);
}

int fnx_advect_vel(float dt, const float* orig, const float* U, const float* flags, float* dst, int B,
                   int D, int H, int W, int is3d, int method, int boundary_width,
                   float maccormack_strength, void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_grid(B, D, H, W, is3d, "advect_vel")) return e;
  if (boundary_width != 1) return fnx_set_error(FNX_ERR_ARG, "advect_vel: boundary_width must be 1 (Q4)");
  if (method != FNX_METHOD_EULER && method != FNX_METHOD_MACCORMACK)
    return fnx_set_error(FNX_ERR_ARG, "advect_vel: No defined method for MacCormackClamp");
  cudaStream_t st = (cudaStream_t)stream;
  Grid g = make_grid(B, D, H, W);
  if (method == FNX_METHOD_EULER) {
    FNX_DISPATCH_3D(is3d, k_advect_vel_fwd, <<<cell_grid(g), cell_block(), 0, st>>>(g, -dt, orig, U, flags, dst));
    FNX_LAUNCH_CHECK("advect_vel", 1);
    return FNX_OK;
  }
  if (workspace_bytes < fnx_advect_vel_workspace(B, D, H, W, is3d) || !workspace)
    return fnx_set_error(FNX_ERR_WORKSPACE, "advect_vel: workspace too small");
  float* fwd = (float*)workspace;
  FNX_DISPATCH_3D(is3d, k_advect_vel_fwd, <<<cell_grid(g), cell_block(), 0, st>>>(g, -dt, orig, U, flags, fwd));
  FNX_DISPATCH_3D(is3d, k_advect_vel_bwd, <<<cell_grid(g), cell_block(), 0, st>>>(
      g, dt, maccormack_strength * 0.5f, orig, U, flags, fwd, dst));
  FNX_LAUNCH_CHECK("advect_vel", 2);
  return FNX_OK;
}

int fnx_velocity_divergence(const float* U, const float* flags, float* div, int B, int D, int H, int W,
                            int is3d, void* stream) {
  if (int e = check_grid(B, D, H, W, is3d, "velocity_divergence")) return e;
  Grid g = make_grid(B, D, H, W);
  FNX_DISPATCH_3D(is3d, k_velocity_divergence, <<<cell_grid(g), cell_block(), 0, (cudaStream_t)stream>>>(g, U, flags, div));
  FNX_LAUNCH_CHECK("velocity_divergence", 1);
  return FNX_OK;
}

int fnx_output_fields(const float* U, const float* flags, const float* density, const float* pressure, float* out,
                      int B, int D, int H, int W, int is3d, int mask_obstacles, void* stream) {
  if (int e = check_grid(B, D, H, W, is3d, "output_fields")) return e;
  if (!U || !flags || !out) return fnx_set_error(FNX_ERR_ARG, "output_fields: U, flags and out are required");
  Grid g = make_grid(B, D, H, W);
  FNX_DISPATCH_3D(is3d, k_output_fields, <<<cell_grid(g), cell_block(), 0, (cudaStream_t)stream>>>(g, U, flags, density, pressure, out, mask_obstacles));
  FNX_LAUNCH_CHECK("output_fields", 1);
  return FNX_OK;
}

int fnx_velocity_update(const float* pressure, float* U, const float* flags, int B, int D, int H, int W,
                        int is3d, void* stream) {
  if (int e = check_grid(B, D, H, W, is3d, "velocity_update")) return e;
  Grid g = make_grid(B, D, H, W);
  FNX_DISPATCH_3D(is3d, k_velocity_update, <<<cell_grid(g), cell_block(), 0, (cudaStream_t)stream>>>(g, pressure, U, flags));
  FNX_LAUNCH_CHECK("velocity_update", 1);
  return FNX_OK;
}

int fnx_set_wall_bcs(float* U, const float* flags, int B, int D, int H, int W, int is3d, void* stream) {
  if (int e = check_grid(B, D, H, W, is3d, "set_wall_bcs")) return e;
  Grid g = make_grid(B, D, H, W);
  FNX_DISPATCH_3D(is3d, k_set_wall_bcs, <<<cell_grid(g), cell_block(), 0, (cudaStream_t)stream>>>(g, U, flags));
  FNX_LAUNCH_CHECK("set_wall_bcs", 1);
  return FNX_OK;
}

int fnx_add_buoyancy(float* U, const float* flags, const float* density, const float* gravity3,
                     float rho_star, float dt, int B, int D, int H, int W, int is3d, void* stream) {
  if (int e = check_grid(B, D, H, W, is3d, "add_buoyancy")) return e;
  Grid g = make_grid(B, D, H, W);
  // strength = gravity * dt as one fp32 product per component (source_terms.py:70)
  float3 s = make_float3(gravity3[0] * dt, gravity3[1] * dt, gravity3[2] * dt);
  FNX_DISPATCH_3D(is3d, k_add_buoyancy, <<<cell_grid(g), cell_block(), 0, (cudaStream_t)stream>>>(g, U, flags, density, s, rho_star));
  FNX_LAUNCH_CHECK("add_buoyancy", 1);
  return FNX_OK;
}

int fnx_add_gravity(float* U, const float* flags, const float* gravity3, float dt, int B, int D, int H,
                    int W, int is3d, void* stream) {
  if (int e = check_grid(B, D, H, W, is3d, "add_gravity")) return e;
  Grid g = make_grid(B, D, H, W);
  float3 f = make_float3(gravity3[0] * dt, gravity3[1] * dt, gravity3[2] * dt);
  FNX_DISPATCH_3D(is3d, k_add_gravity, <<<cell_grid(g), cell_block(), 0, (cudaStream_t)stream>>>(g, U, flags, f));
  FNX_LAUNCH_CHECK("add_gravity", 1);
  return FNX_OK;
}

int fnx_add_viscosity(float* U, const float* flags, double dt, double viscosity, int B, int H, int W, void* workspace,
                      size_t workspace_bytes, void* stream) {
  if (int e = check_grid(B, 1, H, W, 0, "add_viscosity")) return e;
  const size_t bytes = (size_t)B * 2 * H * W * sizeof(float);
  if (!workspace || workspace_bytes < bytes) return fnx_set_error(FNX_ERR_WORKSPACE, "add_viscosity: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  // out of place into the workspace (every term is the pre-update field), then back
  Grid g = make_grid(B, 1, H, W);
  k_add_viscosity<<<cell_grid(g), cell_block(), 0, st>>>(g, U, (float*)workspace, flags, (float)(dt * viscosity));
  FNX_LAUNCH_CHECK("add_viscosity", 1);
  cudaError_t ce = cudaMemcpyAsync(U, workspace, bytes, cudaMemcpyDeviceToDevice, st);
  if (ce != cudaSuccess) return fnx_set_error(FNX_ERR_CUDA, "add_viscosity: %s", cudaGetErrorString(ce));
  return FNX_OK;
}

int fnx_correct_scalar(float* src, const float* div, const float* flags, double dt, size_t count, void* stream) {
  if (count == 0) return FNX_OK;
  k_correct_scalar<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, div, flags, (float)(dt * 0.5),
                                                                                   count);
  FNX_LAUNCH_CHECK("correct_scalar", 1);
  return FNX_OK;
}

int fnx_flags_to_occupancy(const float* flags, float* occupancy, size_t count, void* stream) {
  if (count == 0) return FNX_OK;
  k_flags_to_occupancy<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(flags, occupancy, count);
  FNX_LAUNCH_CHECK("flags_to_occupancy", 1);
  return FNX_OK;
}

int fnx_set_const_vals(float* x, const float* inv_mask, const float* bc, size_t count, void* stream) {
  if (count == 0) return FNX_OK;
  k_set_const_vals<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, inv_mask, bc, count);
  FNX_LAUNCH_CHECK("set_const_vals", 1);
  return FNX_OK;
}

int fnx_empty_domain(float* flags, int B, int D, int H, int W, int is3d, int bnd, void* stream) {
  if (int e = check_grid(B, D, H, W, is3d, "empty_domain")) return e;
  if (bnd < 1) return fnx_set_error(FNX_ERR_ARG, "empty_domain: Boundary width must be greater than zero!");
  Grid g = make_grid(B, D, H, W);
  FNX_DISPATCH_3D(is3d, k_empty_domain, <<<cell_grid(g), cell_block(), 0, (cudaStream_t)stream>>>(g, flags, bnd));
  FNX_LAUNCH_CHECK("empty_domain", 1);
  return FNX_OK;
}

int fnx_get_centered(const float* U, float* out, int B, int D, int H, int W, int is3d, void* stream) {
  if (int e = check_grid(B, D, H, W, is3d, "get_centered")) return e;
  Grid g = make_grid(B, D, H, W);
  FNX_DISPATCH_3D(is3d, k_get_centered, <<<cell_grid(g), cell_block(), 0, (cudaStream_t)stream>>>(g, U, out));
  FNX_LAUNCH_CHECK("get_centered", 1);
  return FNX_OK;
}

// ---- Jacobi ---------------------------------------------------------------------------
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t fnx_jacobi_workspace(int B, int D, int H, int W, int max_iter) {
  (void)max_iter;
  size_t n = (size_t)B * D * H * W;
  return align256(n * sizeof(float)) + align256(sizeof(JacobiCtrl)) + align256((size_t)B * sizeof(double)) +
         (D > 1 ? align256(n)                                      // 3-D: one neighbour-mask byte per cell
                : align256(fnx_jacobi_2d_tilemask_bytes(B, H, W)));  // 2-D: per-thread tile masks of the blocked kernel
}

int fnx_solve_linear_system_jacobi(const float* flags, const float* div, float* p, float* residual, int B,
                                   int D, int H, int W, int is3d, float p_tol, int max_iter, int* iters_run,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_grid(B, D, H, W, is3d, "solve_linear_system")) return e;
  if (max_iter < 1) return fnx_set_error(FNX_ERR_ARG, "solve_linear_system: At least 1 iteration of the solver is needed.");
  if (!workspace || workspace_bytes < fnx_jacobi_workspace(B, D, H, W, max_iter))
    return fnx_set_error(FNX_ERR_WORKSPACE, "solve_linear_system: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  Grid g = make_grid(B, D, H, W);
  const size_t n = (size_t)B * g.n;
  char* ws = (char*)workspace;
  float* scratch = (float*)ws;
  JacobiCtrl* ctrl = (JacobiCtrl*)(ws + align256(n * sizeof(float)));
  double* ssq = (double*)((char*)ctrl + align256(sizeof(JacobiCtrl)));
  cudaError_t ce = cudaMemsetAsync(ctrl, 0, align256(sizeof(JacobiCtrl)) + align256((size_t)B * sizeof(double)), st);
  if (ce != cudaSuccess) return fnx_set_error(FNX_ERR_CUDA, "solve_linear_system: %s", cudaGetErrorString(ce));

  const bool tol = p_tol > 0.f;
  if (!tol && !is3d) {
    // fixed iteration count, 2-D: temporally blocked shared-memory kernel
    void* tmask = (char*)ssq + align256((size_t)B * sizeof(double));
    // residual == NULL (the fused step: simulate.py:150-153 discards it): no residual pass in the last launch
    int e = fnx_jacobi_2d_blocked(flags, div, nullptr, p, scratch, residual ? ssq : nullptr, B, H, W, max_iter, 0, 0, tmask,
                                  fnx_jacobi_2d_tilemask_bytes(B, H, W), st);
    if (e) return e;
    if (residual) {
      k_jacobi_ctrl<<<1, 1, 0, st>>>(ctrl, ssq, B, p_tol, max_iter - 1, residual);
      FNX_LAUNCH_CHECK("solve_linear_system", 1);
    }
    if (iters_run) *iters_run = max_iter;
    return FNX_OK;
  }
  if (!tol && is3d && jacobi3d_vec_ok(g, flags, div, p, scratch)) {
    // fixed iteration count, 3-D: precomputed neighbour masks + 4 cells per thread
    unsigned char* mask = (unsigned char*)((char*)ssq + align256((size_t)B * sizeof(double)));
    int e = jacobi3d_vec_run(g, flags, div, nullptr, p, scratch, mask, residual ? ssq : nullptr, max_iter, st);
    if (e) return e;
    if (residual) {
      k_jacobi_ctrl<<<1, 1, 0, st>>>(ctrl, ssq, B, p_tol, max_iter - 1, residual);
      FNX_LAUNCH_CHECK("solve_linear_system", 1);
    }
    if (iters_run) *iters_run = max_iter;
    return FNX_OK;
  }
  // buffer written by iteration `it`; the last possible iteration lands in p
  auto wbuf = [&](int it) { return ((max_iter - 1 - it) % 2 == 0) ? p : scratch; };
  int executed = max_iter;
  for (int it = 0; it < max_iter; it++) {
    const bool resid = tol || it == max_iter - 1;
    float* cur = wbuf(it);
    const float* prev = it == 0 ? nullptr : wbuf(it - 1);
    if (is3d) launch_jacobi_iter<true>(g, it == 0, resid, flags, div, prev, cur, ssq, tol ? ctrl : nullptr, st);
    else launch_jacobi_iter<false>(g, it == 0, resid, flags, div, prev, cur, ssq, tol ? ctrl : nullptr, st);
    if (resid) k_jacobi_ctrl<<<1, 1, 0, st>>>(ctrl, ssq, B, p_tol, it, residual);
    fnx_count_launches(resid ? 2 : 1);
    if (tol && ((it & 7) == 7 || it == max_iter - 1)) {
      JacobiCtrl h;
      ce = cudaMemcpyAsync(&h, ctrl, sizeof(h), cudaMemcpyDeviceToHost, st);
      if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
      if (ce != cudaSuccess) return fnx_set_error(FNX_ERR_CUDA, "solve_linear_system: %s", cudaGetErrorString(ce));
      if (h.done || it == max_iter - 1) { executed = h.iters; break; }
    }
  }
  FNX_LAUNCH_CHECK("solve_linear_system", 0);
  if (tol && wbuf(executed - 1) != p) {
    ce = cudaMemcpyAsync(p, scratch, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (ce != cudaSuccess) return fnx_set_error(FNX_ERR_CUDA, "solve_linear_system: %s", cudaGetErrorString(ce));
  }
  if (iters_run) *iters_run = executed;
  return FNX_OK;
}

int fnx_jacobi_iterate(const float* flags, const float* div, const float* p_init, float* p, int B, int D, int H, int W,
                       int is3d, int iters, int row_begin, int row_end, void* workspace, size_t workspace_bytes,
                       void* stream) {
  if (int e = check_grid(B, D, H, W, is3d, "jacobi_iterate")) return e;
  if (iters < 1) return fnx_set_error(FNX_ERR_ARG, "jacobi_iterate: At least 1 iteration of the solver is needed.");
  if (p_init == p) return fnx_set_error(FNX_ERR_ARG, "jacobi_iterate: p_init must not alias p");
  if (!workspace || workspace_bytes < fnx_jacobi_workspace(B, D, H, W, iters))
    return fnx_set_error(FNX_ERR_WORKSPACE, "jacobi_iterate: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  Grid g = make_grid(B, D, H, W);
  float* scratch = (float*)workspace;
  if (row_end > row_begin) {
    if (row_begin < 0 || row_end > D * H) return fnx_set_error(FNX_ERR_ARG, "jacobi_iterate: row window out of range");
    g.row0 = row_begin; g.row1 = row_end;
  }
  if (!is3d) {
    const size_t nn = (size_t)B * g.n;
    void* tmask = (char*)workspace + align256(nn * sizeof(float)) + align256(sizeof(JacobiCtrl)) +
                  align256((size_t)B * sizeof(double));
    return fnx_jacobi_2d_blocked(flags, div, p_init, p, scratch, nullptr, B, H, W, iters, row_begin, row_end, tmask,
                                 fnx_jacobi_2d_tilemask_bytes(B, H, W), st);
  }
  if (jacobi3d_vec_ok(g, flags, div, p, scratch) && (((uintptr_t)p_init) & 15) == 0) {
    const size_t n = (size_t)B * g.n;
    unsigned char* mask = (unsigned char*)workspace + align256(n * sizeof(float)) + align256(sizeof(JacobiCtrl)) +
                          align256((size_t)B * sizeof(double));
    return jacobi3d_vec_run(g, flags, div, p_init, p, scratch, mask, nullptr, iters, st);
  }
  auto wbuf = [&](int it) { return ((iters - 1 - it) % 2 == 0) ? p : scratch; };
  for (int it = 0; it < iters; it++) {
    const float* prev = it == 0 ? p_init : wbuf(it - 1);
    launch_jacobi_iter<true>(g, prev == nullptr, false, flags, div, prev, wbuf(it), nullptr, nullptr, st);
    fnx_count_launches(1);
  }
  FNX_LAUNCH_CHECK("jacobi_iterate", 0);
  return FNX_OK;
}

// Residual-terminated Jacobi across slabs (fluids_init.cpp:958-990 evaluates max_b ||p - p_prev||_2 after EVERY
// iteration): `iters` single-iteration launches continued from p_init, rows [row_begin, row_end) written, and
// ssq[it * B + b] = sum over the rows [own_begin, own_end) of (p_it - p_{it-1})^2 -- this rank's share of iteration
// it's squared residual.  The caller all-reduces (sum) the iters x B vector, takes sqrt and the max over b, and finds
// the iteration the single-GPU solver stops at; nothing here synchronises.
int fnx_jacobi_iterate_resid(const float* flags, const float* div, const float* p_init, float* p, int B, int D, int H,
                             int W, int is3d, int iters, int row_begin, int row_end, int own_begin, int own_end,
                             double* ssq, void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_grid(B, D, H, W, is3d, "jacobi_iterate_resid")) return e;
  if (iters < 1) return fnx_set_error(FNX_ERR_ARG, "jacobi_iterate_resid: At least 1 iteration of the solver is needed.");
  if (p_init == p) return fnx_set_error(FNX_ERR_ARG, "jacobi_iterate_resid: p_init must not alias p");
  if (!ssq) return fnx_set_error(FNX_ERR_ARG, "jacobi_iterate_resid: ssq is NULL");
  const size_t n = (size_t)B * D * H * W;
  if (!workspace || workspace_bytes < n * sizeof(float))
    return fnx_set_error(FNX_ERR_WORKSPACE, "jacobi_iterate_resid: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  Grid g = make_grid(B, D, H, W);
  if (row_end > row_begin) {
    if (row_begin < 0 || row_end > D * H) return fnx_set_error(FNX_ERR_ARG, "jacobi_iterate_resid: row window out of range");
    g.row0 = row_begin; g.row1 = row_end;
  }
  if (own_end > own_begin) { g.own0 = own_begin; g.own1 = own_end; }
  cudaError_t ce = cudaMemsetAsync(ssq, 0, (size_t)iters * B * sizeof(double), st);
  if (ce != cudaSuccess) return fnx_set_error(FNX_ERR_CUDA, "jacobi_iterate_resid: %s", cudaGetErrorString(ce));
  float* scratch = (float*)workspace;
  auto wbuf = [&](int it) { return ((iters - 1 - it) % 2 == 0) ? p : scratch; };
  for (int it = 0; it < iters; it++) {
    const float* prev = it == 0 ? p_init : wbuf(it - 1);
    if (is3d) launch_jacobi_iter<true>(g, prev == nullptr, true, flags, div, prev, wbuf(it), ssq + (size_t)it * B, nullptr, st);
    else launch_jacobi_iter<false>(g, prev == nullptr, true, flags, div, prev, wbuf(it), ssq + (size_t)it * B, nullptr, st);
  }
  FNX_LAUNCH_CHECK("jacobi_iterate_resid", iters);
  return FNX_OK;
}

// 2-D, arrays holding rows [held_row_begin, held_row_end) only (a slab of the domain-decomposed step):
// `iters` iterations continued from p_init (NULL = from p = 0), rows [row_begin, row_end) written.
// More than 8 iterations ping-pong through `workspace` (held rows x W floats).
int fnx_jacobi_iterate_held(const float* flags, const float* div, const float* p_init, float* p, int B, int H, int W,
                            int iters, int row_begin, int row_end, int held_row_begin, int held_row_end,
                            void* workspace, size_t workspace_bytes, void* stream) {
  if (B < 1 || H < 2 || W < 2) return fnx_set_error(FNX_ERR_ARG, "jacobi_iterate_held: bad grid");
  if (iters < 1) return fnx_set_error(FNX_ERR_ARG, "jacobi_iterate_held: At least 1 iteration of the solver is needed.");
  if (p_init == p) return fnx_set_error(FNX_ERR_ARG, "jacobi_iterate_held: p_init must not alias p");
  const size_t n = (size_t)B * (held_row_end - held_row_begin) * W;
  if (iters > 8 && (!workspace || workspace_bytes < n * sizeof(float)))
    return fnx_set_error(FNX_ERR_WORKSPACE, "jacobi_iterate_held: workspace too small");
  return fnx_jacobi_2d_blocked_held(flags, div, p_init, p, (float*)workspace, nullptr, B, H, W, iters, row_begin, row_end,
                                    held_row_begin, held_row_end, nullptr, 0, (cudaStream_t)stream);
}

// Tile masks of one launch geometry for fnx_jacobi_iterate_held_masked (the flags are static during a simulation:
// a slab stepper computes them once per launch shape instead of decoding the flags in every launch).
size_t fnx_jacobi_tilemask_bytes(int B, int rows, int W) { return fnx_jacobi_2d_tilemask_bytes(B, rows, W); }

int fnx_jacobi_tilemask_held(const float* flags, int B, int H, int W, int row_begin, int row_end, int held_row_begin,
                             int held_row_end, void* masks, size_t mask_bytes, void* stream) {
  return fnx_jacobi_2d_blocked_masks(flags, flags, nullptr, nullptr, nullptr, B, H, W, 1, row_begin, row_end, held_row_begin,
                                     held_row_end, masks, mask_bytes, 1, (cudaStream_t)stream);
}

// one launch (iters <= 8) with masks computed by fnx_jacobi_tilemask_held for the SAME (row_begin, row_end, held rows)
int fnx_jacobi_iterate_held_masked(const float* flags, const float* div, const float* p_init, float* p, int B, int H, int W,
                                   int iters, int row_begin, int row_end, int held_row_begin, int held_row_end,
                                   const void* masks, size_t mask_bytes, void* stream) {
  if (iters < 1 || iters > 8) return fnx_set_error(FNX_ERR_ARG, "jacobi_iterate_held_masked: 1..8 iterations per call");
  if (p_init == p) return fnx_set_error(FNX_ERR_ARG, "jacobi_iterate_held_masked: p_init must not alias p");
  return fnx_jacobi_2d_blocked_masks(flags, div, p_init, p, nullptr, B, H, W, iters, row_begin, row_end, held_row_begin,
                                     held_row_end, const_cast<void*>(masks), mask_bytes, 2, (cudaStream_t)stream);
}

}  // extern "C"
