// step2d.cu -- the 2-D stencil stages of lib.simulate (pytorch/lib/simulate.py:28-171) as three
// window-aware sm_100a kernels.  Compiled with -fmad=false (bit-exact vs the reference's ATen path).
//
//   k2_advect      advectScalar + advectVelocity (MacCormack, fluids_init.cpp:265-382, 656-807) +
//                  setConstVals (simulate.py:96) in ONE kernel: the forward (semi-Lagrangian) pass
//                  of a 64x32 output tile and of a 2-cell apron around it is computed into SHARED
//                  MEMORY (rho_fwd, U_fwd, the traced cell index), then the backward pass /
//                  correction / clamp of the tile samples it from there.  The forward fields never
//                  travel through HBM (the two-kernel version wrote and re-read 16 B/cell of them
//                  and re-read every input).  A backward sample that leaves the staged apron
//                  (|u| dt > 1 cell) recomputes the forward value of the cells it needs from global
//                  memory through an out-of-line copy of the same per-cell function: any velocity is
//                  handled, with identical results.
//   k2_forces_div  addBuoyancy / addGravity -> [setWallBcs] -> setConstVals -> velocityDivergence
//                  (simulate.py:98-145): the forced velocity of a tile + one row / column is computed
//                  once into shared memory, the divergence reads its neighbours from there.
//   k2_project     velocityUpdate -> [setWallBcs] -> setConstVals (simulate.py:154-168), in place.
//
// Per-cell arithmetic, operation order and quirks are those of the per-op
// device functions (advect_device.cuh, advect_cells.cuh, stencil_device.cuh) the op-level entry points
// and the 3-D path keep using; tests/test_gpu_parity.py holds the two implementations equal bit for
// bit, and both equal to the reference's golden outputs.
//
// Windows.  Every array argument may hold only rows [ya0, ya1) of the global H x W grid (a slab of
// the domain-decomposed step): the kernels index with GLOBAL coordinates through "virtual base"
// pointers (base - ya0*W), channel c of a multi-channel field lives `cs` elements after channel
// c-1, and every data-dependent row index is clamped to the rows that exist.  With ya0 = 0, ya1 = H
// the clamps are the reference's own clamps to the grid.
#include <cuda_runtime.h>

#include "../../include/fluidstep.h"
#include "fluid_common.cuh"
#include "host_util.h"
#include "stencil_device.cuh"
#include "step2d.h"

namespace fnx {
namespace s2 {

// ---- advection ------------------------------------------------------------------------------------
constexpr int TX = 64, TY = 32;          // output tile
constexpr int AP = 2;                    // apron of forward values around it
constexpr int SW = TX + 2 * AP;
constexpr int NT = 256;                  // threads: 64 x 4

struct Adv {
  int H, W;
  int row0, row1;        // rows written
  int ycl, ych;          // clamp of a bilinear sample's lower row:   [max(0,ya0), min(H-2, ya1-2)]
  int yrl, yrh;          // clamp of a cell row index:                [max(0,ya0), min(H-1, ya1-1)]
  int yf0, yf1;          // rows whose forward value can be computed (all +-1 neighbours exist)
  float dt, mdt, hs;
  int sample_outside;
  const float* rho;      // virtual bases of this batch item
  const float* u0;
  const float* u1;
  const float* fl;
};

struct Tap {
  int x0, y0;
  float s0, s1, t0, t1;
};

// grid.cpp:28-52 (Q3): weights clamped to [0,1] before the index clamp; positions beyond 1e9 and NaN
// take the int64 truncation of the ATen path (out of line: never on the hot path)
__device__ __noinline__ void taps_big(const Adv& a, float px, float py, Tap& t) {
  const long long ix = trunc_ll(px), iy = trunc_ll(py);
  const float s1 = px - (float)ix, t1 = py - (float)iy;
  const float s0 = 1.f - s1, t0 = 1.f - t1;
  t.x0 = (int)clampll(ix, 0, a.W - 2);
  t.y0 = (int)clampll(clampll(iy, 0, a.H - 2), a.ycl, a.ych);
  t.s1 = clamp01(s1); t.t1 = clamp01(t1);
  t.s0 = clamp01(s0); t.t0 = clamp01(t0);
}

__device__ __forceinline__ Tap make_tap(const Adv& a, float x, float y) {
  Tap t;
  const float px = x - 0.5f, py = y - 0.5f;
  if ((fabsf(px) < 1.0e9f) & (fabsf(py) < 1.0e9f)) {
    const int ix = __float2int_rz(px), iy = __float2int_rz(py);
    const float s1 = px - (float)ix, t1 = py - (float)iy;
    const float s0 = 1.f - s1, t0 = 1.f - t1;
    int x0 = ix < 0 ? 0 : ix, y0 = iy < a.ycl ? a.ycl : iy;
    t.x0 = x0 > a.W - 2 ? a.W - 2 : x0;
    t.y0 = y0 > a.ych ? a.ych : y0;
    t.s1 = clamp01(s1); t.t1 = clamp01(t1);
    t.s0 = clamp01(s0); t.t0 = clamp01(t0);
  } else {
    taps_big(a, px, py, t);
  }
  return t;
}

// grid.cpp:54-75: (Ia*t0 + Ib*t1)*s0 + (Ic*t0 + Id*t1)*s1 with a=(y0,x0) b=(y0+1,x0) c=(y0,x0+1) d=(y0+1,x0+1)
__device__ __forceinline__ float bilerp(float Ia, float Ib, float Ic, float Id, const Tap& t) {
  return (Ia * t.t0 + Ib * t.t1) * t.s0 + (Ic * t.t0 + Id * t.t1) * t.s1;
}

// grid.cpp:78-96
__device__ __forceinline__ void mix1(float va, bool fa, float vb, bool fb, float wa, float wb, float& v, bool& fl) {
  if (!fa && !fb) { v = 0.f; fl = false; }
  else if (!fa) { v = vb; fl = true; }
  else if (!fb) { v = va; fl = true; }
  else { v = va * wa + vb * wb; fl = true; }
}

// grid.cpp:118-269 interpolWithFluid on four corner values + their fluid bits; all-fluid (the common
// case) is the plain formula, no fluid corner falls back to it as well
__device__ __forceinline__ float bilerp_fluid(float Ia, float Ib, float Ic, float Id, bool fa, bool fb, bool fc,
                                              bool fd, const Tap& t) {
  if (fa & fb & fc & fd) return bilerp(Ia, Ib, Ic, Id, t);
  float ab, cd, v;
  bool fab, fcd, fv;
  mix1(Ia, fa, Ib, fb, t.t0, t.t1, ab, fab);
  mix1(Ic, fc, Id, fd, t.t0, t.t1, cd, fcd);
  mix1(ab, fab, cd, fcd, t.s0, t.s1, v, fv);
  return fv ? v : bilerp(Ia, Ib, Ic, Id, t);
}

__device__ __forceinline__ float sample_g(const Adv& a, const float* __restrict__ f, float x, float y) {
  const Tap t = make_tap(a, x, y);
  const float* p = f + t.y0 * a.W + t.x0;
  return bilerp(__ldg(p), __ldg(p + a.W), __ldg(p + 1), __ldg(p + a.W + 1), t);
}

__device__ __forceinline__ float sample_g_fluid(const Adv& a, const float* __restrict__ f, float x, float y) {
  const Tap t = make_tap(a, x, y);
  const int o = t.y0 * a.W + t.x0;
  const float* p = f + o;
  const float* q = a.fl + o;
  return bilerp_fluid(__ldg(p), __ldg(p + a.W), __ldg(p + 1), __ldg(p + a.W + 1), __ldg(q) == kFluid,
                      __ldg(q + a.W) == kFluid, __ldg(q + 1) == kFluid, __ldg(q + a.W + 1) == kFluid, t);
}

// ---- line trace, 2-D (calc_line_trace.cpp:259-424) --------------------------------------------------
__device__ __forceinline__ bool ood(const Adv& a, float x, float y) {  // :16-27
  return x <= 0.f || x >= (float)a.W || y <= 0.f || y >= (float)a.H;
}
__device__ __forceinline__ bool blocked(const Adv& a, float x, float y) {  // :33-64
  if (ood(a, x, y)) return false;
  int r = __float2int_rz(y);
  r = r < a.yrl ? a.yrl : (r > a.yrh ? a.yrh : r);
  return __ldg(a.fl + r * a.W + __float2int_rz(x)) != kFluid;
}

// the step that leaves the grid or enters a non-fluid cell (:323-412); always ends the trace.
// (px,py) trace start, (bx,by) current position (in/out), (nx,ny) the offending step, (dx,dy) direction
__device__ __noinline__ void trace_stop(const Adv& a, float px, float py, float dirx, float diry, float nx, float ny,
                                        float& bx, float& by) {
  if (ood(a, nx, ny)) {
    // calcRayBorderIntersection from the trace's START position (:175-257, Q11)
    float min_step = CUDART_INF_F;
    const float nxt[2] = {nx, ny}, pos[2] = {px, py}, dimf[2] = {(float)a.W, (float)a.H};
#pragma unroll
    for (int k = 0; k < 2; k++)
      if (nxt[k] <= kHitMargin) {
        const float d = nxt[k] - pos[k];
        if (fabsf(d) >= kEpsilon) min_step = min_t(min_step, (kHitMargin - pos[k]) / d);
      }
#pragma unroll
    for (int k = 0; k < 2; k++) {
      const float lim = dimf[k] - kHitMargin;
      if (nxt[k] >= lim) {
        const float d = nxt[k] - pos[k];
        if (fabsf(d) >= kEpsilon) min_step = min_t(min_step, (lim - pos[k]) / d);
      }
    }
    const bool hit = (min_step >= 0.f) && (min_step < CUDART_INF_F);
    float ix = nx, iy = ny;  // clampToDomain is a no-op in the reference (Q9)
    if (hit) {
      ix = min_step * (nx - px) + px;
      iy = min_step * (ny - py) + py;
    }
    if (!blocked(a, ix, iy)) { bx = ix; by = iy; return; }
    nx = ix; ny = iy;
  }
  if (blocked(a, nx, ny)) {
    bool stopped = false;
    for (int count = 0; count < 4; count++) {
      if (!blocked(a, nx, ny)) break;
      // HitBoundingBox as the ATen code evaluates it (:73-149, Q10) on the cell's inflated box
      const float o[2] = {bx, by}, dir[2] = {dirx, diry}, nn[2] = {nx, ny};
      float minB[2], maxB[2], cand[2], maxT[2], coord[2];
      bool mid[2], inside = true;
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const float ctr = (float)__float2int_rz(nn[k]) + 0.5f;
        minB[k] = ctr - 0.5f - kHitMargin;
        maxB[k] = ctr + 0.5f + kHitMargin;
        const bool lt = o[k] < minB[k], gt = o[k] > maxB[k];
        mid[k] = (o[k] >= minB[k]) && (o[k] <= maxB[k]);
        cand[k] = 0.f;
        if (lt) cand[k] = minB[k];
        if (gt) cand[k] = maxB[k];
        if (lt || gt) inside = false;
      }
      const bool outside = !inside;
#pragma unroll
      for (int k = 0; k < 2; k++) {
        maxT[k] = 0.f;
        if (outside && !mid[k] && dir[k] != 0.f) maxT[k] = (cand[k] - o[k]) / dir[k];
        if ((outside && mid[k]) || dir[k] == 0.f) maxT[k] = -1.f;
      }
      const int wp = maxT[1] > maxT[0] ? 1 : 0;  // argmax keeps the first maximum
      const float T = wp ? maxT[1] : maxT[0];
      bool ret = !(T < 0.f && outside);
      const float err_tol = 1e-6f;
#pragma unroll
      for (int k = 0; k < 2; k++) coord[k] = (k == wp) ? cand[k] : (o[k] + T * dir[k]);
#pragma unroll
      for (int k = 0; k < 2; k++)
        if (k != wp && (coord[k] < minB[k] - err_tol || coord[k] > maxB[k] + err_tol)) ret = false;
      if (!ret) { stopped = true; break; }
      nx = coord[0]; ny = coord[1];
    }
    if (!stopped) { bx = nx; by = ny; }
    return;
  }
  // unreachable from trace2 (it only calls with an offending step); kept for completeness
  bx = nx; by = ny;
}

// trace from the centre (px,py) of an INTERIOR FLUID cell (so the start is neither outside the grid
// nor blocked, :283-288) along (dx,dy); returns the end position
__device__ __forceinline__ void trace2(const Adv& a, float px, float py, float dx, float dy, float& bx, float& by) {
  bx = px; by = py;
  float acc = dx * dx;  // at::norm(2): acc += x*x in fp32, then sqrt
  acc = acc + dy * dy;
  const float length = sqrtf(acc);
  if (length <= kEpsilon) return;
  if (!(length < CUDART_INF_F)) return;   // infinite displacement: the reference's march never ends (advect_device.cuh)
  const float dirx = dx / length, diry = dy / length;
  float cur = 0.f;
  while (cur < length - kHitMargin) {  // :310-314
    const float rem = length - cur;
    const float step = rem < 1.f ? rem : 1.f;
    const float nx = bx + dirx * step, ny = by + diry * step;
    if (ood(a, nx, ny) || blocked(a, nx, ny)) {
      trace_stop(a, px, py, dirx, diry, nx, ny, bx, by);
      return;
    }
    bx = nx; by = ny;
    cur = cur + step;
  }
}

struct Fwd {
  float rho, u0, u1;
  unsigned idx;  // traced cell (j0 << 16 | i0)
};

// forward pass of cell (j,i): SemiLagrangeEulerFluidNetSavePos (:69-133) + SemiLagrangeEulerFluidNetMAC
// (:388-451; no line trace Q2, solid-cell channel mix-up Q1)
__device__ __forceinline__ Fwd fwd_cell(const Adv& a, int j, int i) {
  Fwd f;
  const int c = j * a.W + i;
  f.idx = ((unsigned)j << 16) | (unsigned)i;
  const bool border = (i < 1) | (i > a.W - 2) | (j < 1) | (j > a.H - 2);
  if (border) {
    f.rho = 0.f; f.u0 = 0.f; f.u1 = 0.f;
    return f;
  }
  if (__ldg(a.fl + c) != kFluid) {
    f.rho = __ldg(a.rho + c);  // don't advect solid geometry
    f.u0 = __ldg(a.u1 + c);    // Q1
    f.u1 = 0.f;
    return f;
  }
  const int W = a.W;
  const float px = (float)i + 0.5f, py = (float)j + 0.5f;
  const float ux = __ldg(a.u0 + c), uxr = __ldg(a.u0 + c + 1), uxd = __ldg(a.u0 + c - W), uxdr = __ldg(a.u0 + c - W + 1);
  const float vy = __ldg(a.u1 + c), vyl = __ldg(a.u1 + c - 1), vyu = __ldg(a.u1 + c + W), vyul = __ldg(a.u1 + c + W - 1);
  // scalar: centred velocity (grid.cpp:300-309), line trace, fluid-aware sample
  {
    const float cx = 0.5f * (ux + uxr), cy = 0.5f * (vy + vyu);
    float bx, by;
    trace2(a, px, py, a.mdt * cx, a.mdt * cy, bx, by);
    f.rho = a.sample_outside ? sample_g(a, a.rho, bx, by) : sample_g_fluid(a, a.rho, bx, by);
    const int i0 = trunc_clamp0(bx, W - 1);
    int j0 = trunc_clamp0(by, a.H - 1);
    j0 = j0 < a.yrl ? a.yrl : (j0 > a.yrh ? a.yrh : j0);
    f.idx = ((unsigned)j0 << 16) | (unsigned)i0;
  }
  // velocity: face-centred velocities (grid.cpp:314-402), sum order ((a+b)+c)+d
  {
    const float vx_y = 0.25f * (((vy + vyl) + vyu) + vyul);  // v at the x face
    f.u0 = sample_g(a, a.u0, px + ux * a.mdt, py + vx_y * a.mdt);
    const float vy_x = 0.25f * (((ux + uxd) + uxr) + uxdr);  // u at the y face
    f.u1 = sample_g(a, a.u1, px + vy_x * a.mdt, py + vy * a.mdt);
  }
  return f;
}

__device__ __noinline__ void fwd_cell_far(const Adv& a, int j, int i, Fwd& f) { f = fwd_cell(a, j, i); }

template <int TYY>
struct Smem {   // forward values of a TX x TYY tile + apron
  static constexpr int SHH = TYY + 2 * AP;
  float rho[SHH][SW];
  float u0[SHH][SW];
  float u1[SHH][SW];
  unsigned idx[SHH][SW];
};

// forward value `which` (0 rho, 1 u0, 2 u1) at the four corners of a tap: from the staged apron, or
// recomputed when the sample leaves it
template <int WHICH, int TYY>
__device__ __forceinline__ void fwd_corners(const Adv& a, const Smem<TYY>& s, int ty0, int tx0, const Tap& t, float& Ia,
                                            float& Ib, float& Ic, float& Id) {
  const int ly = t.y0 - (ty0 - AP), lx = t.x0 - (tx0 - AP);
  if ((unsigned)ly < (unsigned)(Smem<TYY>::SHH - 1) && (unsigned)lx < (unsigned)(SW - 1) && t.y0 >= a.yf0 && t.y0 + 1 < a.yf1) {
    const float(*f)[SW] = WHICH == 0 ? s.rho : (WHICH == 1 ? s.u0 : s.u1);
    Ia = f[ly][lx]; Ib = f[ly + 1][lx]; Ic = f[ly][lx + 1]; Id = f[ly + 1][lx + 1];
    return;
  }
  Fwd f;
  float v[4];
#pragma unroll 1
  for (int k = 0; k < 4; k++) {
    const int jj = t.y0 + (k & 1), ii = t.x0 + (k >> 1);
    if (jj >= a.yf0 && jj < a.yf1) fwd_cell_far(a, jj, ii, f);
    else { f.rho = 0.f; f.u0 = 0.f; f.u1 = 0.f; }   // rows at the edge of a slab's memory: stale-halo territory
    v[k] = WHICH == 0 ? f.rho : (WHICH == 1 ? f.u0 : f.u1);
  }
  Ia = v[0]; Ib = v[1]; Ic = v[2]; Id = v[3];
}

struct Masks {   // setConstVals (simulate.py:4-26); virtual bases, NULL = no mask
  const float* u0bc; const float* u0inv; const float* u1bc; const float* u1inv;
  const float* rbc; const float* rinv;
  const unsigned char* rows;   // per row: bit0 = U masks differ from identity, bit1 = density masks (NULL: test every cell)
};

// TYY = rows per tile (32 on large grids; 16 / 8 on small ones, where a launch is a handful of tiles per SM
// and its duration is the latency of ONE tile).  With a tile list (written by k2_advect_clean) only the tiles
// that kernel declined are done here -- a few per cent of a plume grid, the ring along the walls.
template <int TYY>
__global__ void __launch_bounds__(NT, 3)
    k2_advect(const __grid_constant__ Adv a, const __grid_constant__ Masks m, int rho_passes, float* __restrict__ rho_out, float* __restrict__ rho_mid,
              float* __restrict__ u0_out, float* __restrict__ u1_out, int tiles_x, const int* __restrict__ tile_count,
              const int* __restrict__ tile_list) {
  __shared__ Smem<TYY> s;
  constexpr int SH_ = Smem<TYY>::SHH;
  int tile = blockIdx.x;
  if (tile_list) {
    if (tile >= *tile_count) return;
    tile = tile_list[tile];
  }
  const int tile_y = tile / tiles_x, tile_x = tile - tile_y * tiles_x;
  const int tx0 = tile_x * TX, ty0 = a.row0 + tile_y * TYY;
  const int W = a.W, H = a.H;
  if (ty0 >= a.row1) return;

  // ---- phase 1: forward pass of the tile and its apron into shared memory ----
  for (int e = threadIdx.x; e < SW * SH_; e += NT) {
    const int ly = e / SW, lx = e - ly * SW;
    const int j = ty0 - AP + ly, i = tx0 - AP + lx;
    if (i < 0 || i >= W || j < a.yf0 || j >= a.yf1) continue;  // never sampled (samples clamp to existing rows)
    const Fwd f = fwd_cell(a, j, i);
    s.rho[ly][lx] = f.rho; s.u0[ly][lx] = f.u0; s.u1[ly][lx] = f.u1; s.idx[ly][lx] = f.idx;
  }
  __syncthreads();

  // ---- phase 2: backward pass + MacCormack correction + clamp + setConstVals of the tile ----
  const int lxo = threadIdx.x & (TX - 1), lyo = threadIdx.x / TX;
  const int i = tx0 + lxo;
  if (i >= W) return;
#pragma unroll 1
  for (int r = lyo; r < TYY; r += NT / TX) {
    const int j = ty0 + r;
    if (j >= a.row1) break;
    const int c = j * W + i;
    const int ly = r + AP, lx = lxo + AP;
    const bool border = (i < 1) | (i > W - 2) | (j < 1) | (j > H - 2);
    const float fc = __ldg(a.fl + c);
    const bool fluid = fc == kFluid;
    float rho_v, u0_v = 0.f, u1_v = 0.f;
    // ---------------- scalar (fluids_init.cpp:322-382) ----------------
    {
      const float fw = s.rho[ly][lx];
      float v = fw;
      float ux = 0.f, uxr = 0.f, vy = 0.f, vyu = 0.f;
      if (fluid) {
        float bwd = 0.f;  // border cells of the backward pass are zeroed (:354-363)
        if (!border) {
          ux = __ldg(a.u0 + c); uxr = __ldg(a.u0 + c + 1); vy = __ldg(a.u1 + c); vyu = __ldg(a.u1 + c + W);
          const float cx = 0.5f * (ux + uxr), cy = 0.5f * (vy + vyu);
          float bx, by;
          trace2(a, (float)i + 0.5f, (float)j + 0.5f, a.dt * cx, a.dt * cy, bx, by);
          const Tap t = make_tap(a, bx, by);
          float Ia, Ib, Ic, Id;
          fwd_corners<0, TYY>(a, s, ty0, tx0, t, Ia, Ib, Ic, Id);
          if (a.sample_outside) {
            bwd = bilerp(Ia, Ib, Ic, Id, t);
          } else {
            const float* q = a.fl + t.y0 * W + t.x0;
            bwd = bilerp_fluid(Ia, Ib, Ic, Id, __ldg(q) == kFluid, __ldg(q + W) == kFluid, __ldg(q + 1) == kFluid,
                               __ldg(q + W + 1) == kFluid, t);
          }
        }
        v = fw + a.hs * (__ldg(a.rho + c) - bwd);  // MacCormackCorrect :135-148
      }
      if (!border) {
        // getClampBounds :154-222: 3x3 neighbourhood of the forward-traced cell, fluid cells only
        const unsigned id = s.idx[ly][lx];
        const int j0 = (int)(id >> 16), i0 = (int)(id & 0xffffu);
        float mn = CUDART_INF_F, mx = -CUDART_INF_F;
        bool any = false;
        const int jlo = a.yrl, jhi = a.yrh;
        if (i0 >= 1 && i0 <= W - 2 && j0 - 1 >= jlo && j0 + 1 <= jhi) {
          const float* sp = a.rho + (j0 - 1) * W + (i0 - 1);
          const float* fp = a.fl + (j0 - 1) * W + (i0 - 1);
#pragma unroll
          for (int dj = 0; dj < 3; dj++)
#pragma unroll
            for (int di = 0; di < 3; di++) {
              if (a.sample_outside || __ldg(fp + dj * W + di) == kFluid) {
                const float sv = __ldg(sp + dj * W + di);
                mn = min_t(mn, sv); mx = max_t(mx, sv);
                any = true;
              }
            }
        } else {
          for (int dj = -1; dj <= 1; dj++) {
            const int jj = j0 + dj;
            if (jj < 0 || jj >= H || jj < jlo || jj > jhi) continue;
            for (int di = -1; di <= 1; di++) {
              const int ii = i0 + di;
              if (ii < 0 || ii >= W) continue;
              const int q = jj * W + ii;
              if (a.sample_outside || __ldg(a.fl + q) == kFluid) {
                const float sv = __ldg(a.rho + q);
                mn = min_t(mn, sv); mx = max_t(mx, sv);
                any = true;
              }
            }
          }
        }
        v = any ? max_t(mn, min_t(mx, v)) : fw;
      }
      rho_v = v;
      // ---------------- velocity (fluids_init.cpp:700-807) ----------------
      if (!border) {
        if (!fluid) { ux = __ldg(a.u0 + c); uxr = __ldg(a.u0 + c + 1); vy = __ldg(a.u1 + c); vyu = __ldg(a.u1 + c + W); }
        const float uxd = __ldg(a.u0 + c - W), uxdr = __ldg(a.u0 + c - W + 1);
        const float vyl = __ldg(a.u1 + c - 1), vyul = __ldg(a.u1 + c + W - 1);
        const float px = (float)i + 0.5f, py = (float)j + 0.5f;
        const float fi = (float)i, fj = (float)j;
        const bool solid = !fluid;
#pragma unroll
        for (int comp = 0; comp < 2; comp++) {
          float velx, vely;
          if (comp == 0) { velx = ux; vely = 0.25f * (((vy + vyl) + vyu) + vyul); }
          else { velx = 0.25f * (((ux + uxd) + uxr) + uxdr); vely = vy; }
          const float* oc = comp == 0 ? a.u0 : a.u1;
          const float fw = comp == 0 ? s.u0[ly][lx] : s.u1[ly][lx];
          // correction skipped when the cell or its lower neighbour along `comp` is not fluid (:470-487)
          bool skip = solid;
          if (!skip && __ldg(a.fl + c - (comp == 0 ? 1 : W)) != kFluid) skip = true;   // interior: index > 0
          float v = fw;
          const float vdx = velx * a.dt, vdy = vely * a.dt;
          if (!skip) {
            const Tap t = make_tap(a, px + vdx, py + vdy);
            float Ia, Ib, Ic, Id;
            if (comp == 0) fwd_corners<1, TYY>(a, s, ty0, tx0, t, Ia, Ib, Ic, Id);
            else fwd_corners<2, TYY>(a, s, ty0, tx0, t, Ia, Ib, Ic, Id);
            const float bwd = bilerp(Ia, Ib, Ic, Id, t);
            v = fw + a.hs * ((comp == 0 ? ux : vy) - bwd);
          }
          // doClampComponentMAC :500-614: min/max of orig over the 2x2 blocks at trunc(pos -/+ vel*dt) (Q5)
          float mn = CUDART_INF_F, mx = -CUDART_INF_F;
#pragma unroll
          for (int l = 0; l < 2; l++) {
            const int q0 = trunc_x86_clamp0(l == 0 ? fi - vdx : fi + vdx, W - 2);
            int q1 = trunc_x86_clamp0(l == 0 ? fj - vdy : fj + vdy, H - 2);
            q1 = q1 < a.ycl ? a.ycl : (q1 > a.ych ? a.ych : q1);
            const float* b0 = oc + q1 * W + q0;
            float sv;
            sv = __ldg(b0); mn = min_t(mn, sv); mx = max_t(mx, sv);
            sv = __ldg(b0 + 1); mn = min_t(mn, sv); mx = max_t(mx, sv);
            sv = __ldg(b0 + W); mn = min_t(mn, sv); mx = max_t(mx, sv);
            sv = __ldg(b0 + W + 1); mn = min_t(mn, sv); mx = max_t(mx, sv);
          }
          v = max_t(min_t(v, mx), mn);
          if (comp == 0) u0_v = v; else u1_v = v;
        }
      }
    }
    // ---------------- setConstVals (simulate.py:96) ----------------
    const unsigned char rb = m.rows ? m.rows[j] : (unsigned char)3;
    if (m.u0bc && (rb & 1)) {
      u0_v = const_vals_apply(u0_v, __ldg(m.u0inv + c), __ldg(m.u0bc + c));
      u1_v = const_vals_apply(u1_v, __ldg(m.u1inv + c), __ldg(m.u1bc + c));
    }
    u0_out[c] = u0_v;
    u1_out[c] = u1_v;
    if (m.rbc && (rb & 2)) {
      const float inv = __ldg(m.rinv + c), bc = __ldg(m.rbc + c);
      rho_v = const_vals_apply(rho_v, inv, bc);
      if (rho_mid) rho_mid[c] = rho_v;        // the density addBuoyancy reads (one pass)
      for (int t = 0; t < rho_passes; t++) rho_v = const_vals_apply(rho_v, inv, bc);   // simulate.py:133,168
    }
    rho_out[c] = rho_v;
  }
}

// ---- advection, interior fast path --------------------------------------------------------------------
// A tile is CLEAN when, on the tile and the two cells around it, every cell is Fluid, no cell computed
// is a border cell, every row is held in memory, no value is NaN and |u| dt <= 0.95 cells.  Then (with
// exactly the same arithmetic as the generic path, which it is tested against bit for bit):
//   * no border / flag test can fire and every fluid-aware sample is the plain bilinear formula;
//   * a line trace never leaves the grid nor meets a blocked cell: it is its free-step loop;
//   * every back-traced position stays within one cell of its start, so every index clamp, the weight
//     clamp to [0,1] and the |x| < 1e9 guard are identities and the apron of forward values is ONE cell;
//   * the min / max of the MacCormack clamps run over values known not to be NaN, where the
//     reference's compare-and-select equals fminf / fmaxf (up to the sign of a zero, which the parity
//     convention of this repo -- and every consumer of these fields -- treats as equal).
// Tiles that are not clean are appended to a list and done by the generic kernel.
constexpr int CA = 1;                               // apron of the clean path
constexpr int CSW = TX + 2 * CA;
constexpr int CR = 2;                               // scanned margin
template <int TYY>
struct SmemC {
  static constexpr int CSH = TYY + 2 * CA;
  float rho[CSH][CSW];
  float u0[CSH][CSW];
  float u1[CSH][CSW];
  unsigned idx[CSH][CSW];
};

__device__ __forceinline__ void trace_free(float px, float py, float dx, float dy, float& bx, float& by) {
  bx = px; by = py;
  float acc = dx * dx;
  acc = acc + dy * dy;
  const float length = sqrtf(acc);
  if (length <= kEpsilon) return;
  const float dirx = dx / length, diry = dy / length;
  float cur = 0.f;
  while (cur < length - kHitMargin) {
    const float rem = length - cur;
    const float step = rem < 1.f ? rem : 1.f;
    bx = bx + dirx * step; by = by + diry * step;
    cur = cur + step;
  }
}

struct TapC {
  int ix, iy;
  float s0, s1, t0, t1;
};
__device__ __forceinline__ TapC tap_c(float x, float y) {
  TapC t;
  const float px = x - 0.5f, py = y - 0.5f;
  t.ix = __float2int_rz(px); t.iy = __float2int_rz(py);
  t.s1 = px - (float)t.ix; t.t1 = py - (float)t.iy;
  t.s0 = 1.f - t.s1; t.t0 = 1.f - t.t1;
  return t;
}
__device__ __forceinline__ float bilerp_c(float Ia, float Ib, float Ic, float Id, const TapC& t) {
  return (Ia * t.t0 + Ib * t.t1) * t.s0 + (Ic * t.t0 + Id * t.t1) * t.s1;
}
__device__ __forceinline__ float gather_g(const float* __restrict__ f, int W, float x, float y) {
  const TapC t = tap_c(x, y);
  const float* p = f + t.iy * W + t.ix;
  return bilerp_c(__ldg(p), __ldg(p + W), __ldg(p + 1), __ldg(p + W + 1), t);
}

template <int TYY>
__global__ void __launch_bounds__(NT, 4)
    k2_advect_clean(const __grid_constant__ Adv a, const __grid_constant__ Masks m, int rho_passes, int ya0, int ya1,
                    float* __restrict__ rho_out, float* __restrict__ rho_mid, float* __restrict__ u0_out,
                    float* __restrict__ u1_out, int tiles_x, int* __restrict__ slow_count, int* __restrict__ slow_list) {
  __shared__ SmemC<TYY> s;
  constexpr int CSH = SmemC<TYY>::CSH;
  const int tile_y = blockIdx.x / tiles_x, tile_x = blockIdx.x - tile_y * tiles_x;
  const int tx0 = tile_x * TX, ty0 = a.row0 + tile_y * TYY;
  const int W = a.W, H = a.H;

  // ---- classification ----
  int ok = (tx0 - CA >= 1) && (tx0 + TX - 1 + CA <= W - 2) && (ty0 - CA >= 1) && (ty0 + TYY - 1 + CA <= H - 2) &&
           (ty0 - CR >= ya0) && (ty0 + TYY + CR <= ya1);
  if (ok) {
    const float adt = fabsf(a.dt);
    constexpr int RW = TX + 2 * CR, RH = TYY + 2 * CR;
    for (int e = threadIdx.x; e < RW * RH; e += NT) {
      const int ly = e / RW, lx = e - ly * RW;
      const int c = (ty0 - CR + ly) * W + (tx0 - CR + lx);
      const float f = __ldg(a.fl + c), u = __ldg(a.u0 + c), v = __ldg(a.u1 + c), r = __ldg(a.rho + c);
      ok &= (f == kFluid) & (fabsf(u) * adt <= 0.95f) & (fabsf(v) * adt <= 0.95f) & (r == r);
    }
  }
  ok = __syncthreads_and(ok);
  if (!ok) {
    if (threadIdx.x == 0) slow_list[atomicAdd(slow_count, 1)] = blockIdx.x;
    return;
  }

  // ---- phase 1: forward pass of the tile + one cell around it ----
  for (int e = threadIdx.x; e < CSW * CSH; e += NT) {
    const int ly = e / CSW, lx = e - ly * CSW;
    const int j = ty0 - CA + ly, i = tx0 - CA + lx;
    const int c = j * W + i;
    const float px = (float)i + 0.5f, py = (float)j + 0.5f;
    const float ux = __ldg(a.u0 + c), uxr = __ldg(a.u0 + c + 1), uxd = __ldg(a.u0 + c - W), uxdr = __ldg(a.u0 + c - W + 1);
    const float vy = __ldg(a.u1 + c), vyl = __ldg(a.u1 + c - 1), vyu = __ldg(a.u1 + c + W), vyul = __ldg(a.u1 + c + W - 1);
    const float cx = 0.5f * (ux + uxr), cy = 0.5f * (vy + vyu);
    float bx, by;
    trace_free(px, py, a.mdt * cx, a.mdt * cy, bx, by);
    s.rho[ly][lx] = gather_g(a.rho, W, bx, by);
    s.idx[ly][lx] = ((unsigned)__float2int_rz(by) << 16) | (unsigned)__float2int_rz(bx);
    const float vx_y = 0.25f * (((vy + vyl) + vyu) + vyul);
    s.u0[ly][lx] = gather_g(a.u0, W, px + ux * a.mdt, py + vx_y * a.mdt);
    const float vy_x = 0.25f * (((ux + uxd) + uxr) + uxdr);
    s.u1[ly][lx] = gather_g(a.u1, W, px + vy_x * a.mdt, py + vy * a.mdt);
  }
  __syncthreads();

  // ---- phase 2: backward pass + correction + clamp + setConstVals ----
  const int lxo = threadIdx.x & (TX - 1), lyo = threadIdx.x / TX;
  const int i = tx0 + lxo;
  const int lbase_x = tx0 - CA, lbase_y = ty0 - CA;
#pragma unroll 1
  for (int r = lyo; r < TYY; r += NT / TX) {
    const int j = ty0 + r;
    if (j >= a.row1) break;
    const int c = j * W + i;
    const int ly = r + CA, lx = lxo + CA;
    const float px = (float)i + 0.5f, py = (float)j + 0.5f;
    const float ux = __ldg(a.u0 + c), uxr = __ldg(a.u0 + c + 1), uxd = __ldg(a.u0 + c - W), uxdr = __ldg(a.u0 + c - W + 1);
    const float vy = __ldg(a.u1 + c), vyl = __ldg(a.u1 + c - 1), vyu = __ldg(a.u1 + c + W), vyul = __ldg(a.u1 + c + W - 1);
    float rho_v, u0_v, u1_v;
    {  // scalar
      const float fw = s.rho[ly][lx];
      const float cx = 0.5f * (ux + uxr), cy = 0.5f * (vy + vyu);
      float bx, by;
      trace_free(px, py, a.dt * cx, a.dt * cy, bx, by);
      const TapC t = tap_c(bx, by);
      const float* f = &s.rho[t.iy - lbase_y][t.ix - lbase_x];
      const float bwd = bilerp_c(f[0], f[CSW], f[1], f[CSW + 1], t);
      float v = fw + a.hs * (__ldg(a.rho + c) - bwd);
      const unsigned id = s.idx[ly][lx];
      const float* sp = a.rho + ((int)(id >> 16) - 1) * W + ((int)(id & 0xffffu) - 1);
      float mn = CUDART_INF_F, mx = -CUDART_INF_F;
#pragma unroll
      for (int dj = 0; dj < 3; dj++)
#pragma unroll
        for (int di = 0; di < 3; di++) {
          const float sv = __ldg(sp + dj * W + di);
          mn = fminf(mn, sv); mx = fmaxf(mx, sv);
        }
      rho_v = max_t(mn, min_t(mx, v));
    }
    {  // velocity
      const float fi = (float)i, fj = (float)j;
#pragma unroll
      for (int comp = 0; comp < 2; comp++) {
        float velx, vely;
        if (comp == 0) { velx = ux; vely = 0.25f * (((vy + vyl) + vyu) + vyul); }
        else { velx = 0.25f * (((ux + uxd) + uxr) + uxdr); vely = vy; }
        const float* oc = comp == 0 ? a.u0 : a.u1;
        const float fw = comp == 0 ? s.u0[ly][lx] : s.u1[ly][lx];
        const float vdx = velx * a.dt, vdy = vely * a.dt;
        const TapC t = tap_c(px + vdx, py + vdy);
        const float* f = comp == 0 ? &s.u0[t.iy - lbase_y][t.ix - lbase_x] : &s.u1[t.iy - lbase_y][t.ix - lbase_x];
        const float bwd = bilerp_c(f[0], f[CSW], f[1], f[CSW + 1], t);
        float v = fw + a.hs * ((comp == 0 ? ux : vy) - bwd);
        float mn = CUDART_INF_F, mx = -CUDART_INF_F;
#pragma unroll
        for (int l = 0; l < 2; l++) {
          const int q0 = __float2int_rz(l == 0 ? fi - vdx : fi + vdx);
          const int q1 = __float2int_rz(l == 0 ? fj - vdy : fj + vdy);
          const float* b0 = oc + q1 * W + q0;
          const float s00 = __ldg(b0), s01 = __ldg(b0 + 1), s10 = __ldg(b0 + W), s11 = __ldg(b0 + W + 1);
          mn = fminf(fminf(mn, s00), fminf(s01, fminf(s10, s11)));
          mx = fmaxf(fmaxf(mx, s00), fmaxf(s01, fmaxf(s10, s11)));
        }
        v = max_t(min_t(v, mx), mn);
        if (comp == 0) u0_v = v; else u1_v = v;
      }
    }
    const unsigned char rb = m.rows ? m.rows[j] : (unsigned char)3;
    if (m.u0bc && (rb & 1)) {
      u0_v = const_vals_apply(u0_v, __ldg(m.u0inv + c), __ldg(m.u0bc + c));
      u1_v = const_vals_apply(u1_v, __ldg(m.u1inv + c), __ldg(m.u1bc + c));
    }
    u0_out[c] = u0_v;
    u1_out[c] = u1_v;
    if (m.rbc && (rb & 2)) {
      const float inv = __ldg(m.rinv + c), bc = __ldg(m.rbc + c);
      rho_v = const_vals_apply(rho_v, inv, bc);
      if (rho_mid) rho_mid[c] = rho_v;
      for (int t = 0; t < rho_passes; t++) rho_v = const_vals_apply(rho_v, inv, bc);
    }
    rho_out[c] = rho_v;
  }
}

// ---- forces + BCs + divergence ----------------------------------------------------------------------
// One thread per column marching DOWN a strip of FR rows: the forced velocity (addBuoyancy ->
// addGravity -> [setWallBcs] -> setConstVals, simulate.py:98-133) of every face is computed ONCE; the
// divergence takes the right neighbour's x face from the next lane (warp shuffle; a warp owns 31 output
// columns + 1 overlap lane) and the upper neighbour's y face from the previous iteration (a register);
// flags / density of the lower row are loaded once and carried to the next iteration.
constexpr int FW = 31, FWARPS = 4, FR = 32;   // output columns per warp, warps per CTA, rows per strip
struct Frc {
  int H, W, row0, row1;
  int use_buoyancy, use_gravity, wall_bcs;
  float bs0, bs1, gf0, gf1, rho_star;
};

// one velocity component of one cell: u after the forces, wall BCs and imposed values.  fc / fn flags of
// the cell and of its lower neighbour along the component (the cell itself at index 0, Q13), rc / rn the
// densities addBuoyancy sees there
__device__ __forceinline__ float forced1(const Frc& g, float u, float fc, float fn, float rc, float rn, bool border,
                                         float bs, float gf, bool masked, float inv, float bc) {
  if (!border) {
    if (g.use_buoyancy) u = buoyancy_apply(u, fc, fn, rc, rn, bs, g.rho_star);
    if (g.use_gravity) u = gravity_apply(u, fc, fn, gf);
  }
  if (g.wall_bcs) u = wall_bcs_apply(u, fc, fn);
  if (masked) u = const_vals_apply(u, inv, bc);
  return u;
}

// MASKED = false: no row of the strip (nor the row below / above it) carries imposed values, so the row
// byte map, the mask loads and the choice between the one-pass and the final density drop out (the
// plume inlet touches 4 rows of the grid); chosen per CTA at run time.
template <bool MASKED>
__device__ __forceinline__ void forces_strip(const Frc& g, const Masks& m, const float* __restrict__ rho,
                                             const float* __restrict__ rho_mid, const float* __restrict__ u0,
                                             const float* __restrict__ u1, const float* __restrict__ fl,
                                             float* __restrict__ u0_out, float* __restrict__ u1_out,
                                             float* __restrict__ div, int i, int jb, int jt, int lane) {
  const int W = g.W, H = g.H;
  const bool in = i < W;                                      // lanes beyond the grid only feed shuffles
  const bool out = in && lane < FW;
  const bool cborder = (i < 1) | (i > W - 2);
  const unsigned full = 0xffffffffu;
  auto rows_of = [&](int j) -> unsigned { return MASKED ? (m.rows ? (unsigned)m.rows[j] : 3u) : 0u; };
  auto rho_of = [&](int j) -> const float* { return (MASKED && rho_mid && (rows_of(j) & 2u)) ? rho_mid : rho; };
  const int ic = in ? i : W - 1;                              // clamped column: every lane loads something valid

  // the row above the strip only contributes its y face (it exists unless the strip ends at the grid's top)
  const bool top = jt < H;
  int j = top ? jt : jt - 1;
  float fc = __ldg(fl + j * W + ic), rc = __ldg(rho_of(j) + j * W + ic);
  float fv_up = 0.f;
  constexpr int U = 4;                                        // rows in flight per thread
  while (j >= jb) {
    // loads of up to U rows first (independent of the arithmetic below)
    float au[U], av[U], fbv[U], rbv[U];
#pragma unroll
    for (int k = 0; k < U; k++) {
      const int jj = j - k;
      const int c = (jj < jb ? jb : jj) * W + ic;
      au[k] = __ldg(u0 + c);
      av[k] = __ldg(u1 + c);
      const int cb = jj > 0 ? c - W : c;
      fbv[k] = __ldg(fl + cb);
      rbv[k] = __ldg(rho_of(jj > 0 ? jj - 1 : 0) + cb);
    }
#pragma unroll
    for (int k = 0; k < U; k++) {
      if (j < jb) break;
      const int c = j * W + ic;
      const bool face_only = j == jt;                         // first iteration when `top`
      const float fb = j > 0 ? fbv[k] : fc, rb = j > 0 ? rbv[k] : rc;
      // left neighbour from the previous lane; lane 0 reads it (column 0 is its own neighbour, Q13)
      float fl_l = __shfl_up_sync(full, fc, 1), r_l = __shfl_up_sync(full, rc, 1);
      if (lane == 0) {
        fl_l = fc; r_l = rc;
        if (i > 0 && in) {
          fl_l = __ldg(fl + c - 1);
          r_l = __ldg(rho_of(j) + c - 1);
        }
      }
      const bool masked = MASKED && m.u0bc != nullptr && (rows_of(j) & 1u);
      const bool border = cborder | (j < 1) | (j > H - 2);
      float i0 = 0.f, b0 = 0.f, i1 = 0.f, b1 = 0.f;
      if (masked) {
        i0 = __ldg(m.u0inv + c); b0 = __ldg(m.u0bc + c);
        i1 = __ldg(m.u1inv + c); b1 = __ldg(m.u1bc + c);
      }
      const float b = forced1(g, av[k], fc, j > 0 ? fb : fc, rc, rb, border, g.bs1, g.gf1, masked, i1, b1);
      const float a = forced1(g, au[k], fc, fl_l, rc, r_l, border, g.bs0, g.gf0, masked, i0, b0);
      const float a_right = __shfl_down_sync(full, a, 1);
      if (!face_only && out) {
        u0_out[c] = a;
        u1_out[c] = b;
        if (div) {
          float d = 0.f;
          if (!border) d = a - a_right + b - fv_up;           // velocity_divergence.py:61-66 (Q15)
          if (fc == kObstacle) d = 0.f;
          div[c] = d;
        }
      }
      fv_up = b;
      fc = fb; rc = rb;
      j--;
    }
  }
}

// The common strip: no imposed values on its rows, every row it touches (jb-1 .. jt) inside the grid and
// jt - jb a multiple of 4: no row test, no mask, no clamp left in the loop; running row pointers.
__device__ __forceinline__ void forces_strip_fast(const Frc& g, const float* __restrict__ rho, const float* __restrict__ u0,
                                                  const float* __restrict__ u1, const float* __restrict__ fl,
                                                  float* __restrict__ u0_out, float* __restrict__ u1_out,
                                                  float* __restrict__ div, int i, int jb, int jt, int lane) {
  const int W = g.W;
  const bool in = i < W;
  const bool out = in && lane < FW;
  const bool border = (i < 1) | (i > W - 2);                  // rows jb .. jt-1 are interior rows
  const unsigned full = 0xffffffffu;
  const int ic = in ? i : W - 1;
  const bool lead = lane == 0 && in && i > 0;                 // lane 0 reads its left neighbour itself
  const float* pf = fl + jt * W + ic;
  const float* pr = rho + jt * W + ic;
  const float* pa = u0 + jt * W + ic;
  const float* pb = u1 + jt * W + ic;
  float* oa = u0_out + jt * W + ic;
  float* ob = u1_out + jt * W + ic;
  float* od = div ? div + jt * W + ic : nullptr;
  // the row above the strip: its y face only
  float fc = __ldg(pf), rc = __ldg(pr);
  float fb = __ldg(pf - W), rb = __ldg(pr - W);
  float fv_up = forced1(g, __ldg(pb), fc, fb, rc, rb, border | (jt > g.H - 2), g.bs1, g.gf1, false, 0.f, 0.f);
  fc = fb; rc = rb;
  for (int j = jt - 1; j >= jb; j -= 4) {
    pf -= 4 * W; pr -= 4 * W; pa -= 4 * W; pb -= 4 * W; oa -= 4 * W; ob -= 4 * W;
    if (od) od -= 4 * W;
    // rows j, j-1, j-2, j-3 live at +3W, +2W, +W, 0 of the moved pointers; their lower rows one W below
    float au[4], av[4], fbv[4], rbv[4], fl_l0[4], r_l0[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int o = (3 - k) * W;
      au[k] = __ldg(pa + o);
      av[k] = __ldg(pb + o);
      fbv[k] = __ldg(pf + o - W);
      rbv[k] = __ldg(pr + o - W);
      fl_l0[k] = 0.f; r_l0[k] = 0.f;
      if (lead) { fl_l0[k] = __ldg(pf + o - 1); r_l0[k] = __ldg(pr + o - 1); }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int o = (3 - k) * W;
      float fl_l = __shfl_up_sync(full, fc, 1), r_l = __shfl_up_sync(full, rc, 1);
      if (lane == 0) { fl_l = lead ? fl_l0[k] : fc; r_l = lead ? r_l0[k] : rc; }
      const float b = forced1(g, av[k], fc, fbv[k], rc, rbv[k], border, g.bs1, g.gf1, false, 0.f, 0.f);
      const float a = forced1(g, au[k], fc, fl_l, rc, r_l, border, g.bs0, g.gf0, false, 0.f, 0.f);
      const float a_right = __shfl_down_sync(full, a, 1);
      if (out) {
        oa[o] = a;
        ob[o] = b;
        if (od) {
          float d = 0.f;
          if (!border) d = a - a_right + b - fv_up;           // velocity_divergence.py:61-66 (Q15)
          if (fc == kObstacle) d = 0.f;
          od[o] = d;
        }
      }
      fv_up = b;
      fc = fbv[k]; rc = rbv[k];
    }
  }
}

__global__ void __launch_bounds__(32 * FWARPS)
    k2_forces_div(const __grid_constant__ Frc g, const __grid_constant__ Masks m, const float* __restrict__ rho,
                  const float* __restrict__ rho_mid, const float* __restrict__ u0, const float* __restrict__ u1,
                  const float* __restrict__ fl, float* __restrict__ u0_out, float* __restrict__ u1_out,
                  float* __restrict__ div, int tiles_x) {
  const int tile_y = blockIdx.x / tiles_x, tile_x = blockIdx.x - tile_y * tiles_x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = (tile_x * FWARPS + w) * FW + lane;
  const int jb = g.row0 + tile_y * FR;                        // lowest row of the strip
  int jt = jb + FR;                                           // one above its highest row
  jt = jt > g.row1 ? g.row1 : jt;
  // does any row this strip touches (jb-1 .. jt) carry imposed values?
  int any = 0;
  if (m.u0bc || rho_mid) {
    if (!m.rows) any = 1;
    else
      for (int r = jb - 1 + (int)threadIdx.x; r <= jt; r += 32 * FWARPS)
        if (r >= 0 && r < g.H && m.rows[r]) any = 1;
  }
  any = __syncthreads_or(any);
  if (any) forces_strip<true>(g, m, rho, rho_mid, u0, u1, fl, u0_out, u1_out, div, i, jb, jt, lane);
  else if (jb >= 1 && jt <= g.H - 1 && ((jt - jb) & 3) == 0 && jt > jb)
    forces_strip_fast(g, rho, u0, u1, fl, u0_out, u1_out, div, i, jb, jt, lane);
  else forces_strip<false>(g, m, rho, rho_mid, u0, u1, fl, u0_out, u1_out, div, i, jb, jt, lane);
}

// ---- velocityUpdate + [setWallBcs] + setConstVals, in place -------------------------------------------
__global__ void __launch_bounds__(256)
    k2_project(int H, int W, int row0, int row1, const float* __restrict__ p, float* __restrict__ u0,
               float* __restrict__ u1, const float* __restrict__ fl, Masks m, int wall_bcs) {
  const int i = blockIdx.x * 128 + (threadIdx.x & 127);
  const int j = row0 + blockIdx.y * 2 + (threadIdx.x >> 7);
  if (i >= W || j >= row1) return;
  const int c = j * W + i;
  const float fc = __ldg(fl + c);
  const bool interior = !((i < 1) | (i > W - 2) | (j < 1) | (j > H - 2));
  const float P = __ldg(p + c);
  const float fl_l = i > 0 ? __ldg(fl + c - 1) : fc, fl_d = j > 0 ? __ldg(fl + c - W) : fc;
  float a = u0[c], b = u1[c];
  if (interior) {
    a = velocity_update_apply(a, fc, fl_l, P, __ldg(p + c - 1));
    b = velocity_update_apply(b, fc, fl_d, P, __ldg(p + c - W));
  }
  if (wall_bcs) {
    a = wall_bcs_apply(a, fc, fl_l);
    b = wall_bcs_apply(b, fc, fl_d);
  }
  const unsigned char rb = m.rows ? m.rows[j] : (unsigned char)3;
  if (m.u0bc && (rb & 1)) {
    a = const_vals_apply(a, __ldg(m.u0inv + c), __ldg(m.u0bc + c));
    b = const_vals_apply(b, __ldg(m.u1inv + c), __ldg(m.u1bc + c));
  }
  u0[c] = a;
  u1[c] = b;
}

// four consecutive cells of a row per thread (16-byte loads / stores); W % 4 == 0 and 16-byte aligned rows
__global__ void __launch_bounds__(128)
    k2_project4(int H, int W, int row0, int row1, const float* __restrict__ p, float* __restrict__ u0,
                float* __restrict__ u1, const float* __restrict__ fl, Masks m, int wall_bcs) {
  const int i0 = (blockIdx.x * 128 + threadIdx.x) * 4;
  const int j = row0 + blockIdx.y;
  if (i0 >= W || j >= row1) return;
  const int c = j * W + i0;
  const float4 P4 = __ldg(reinterpret_cast<const float4*>(p + c));
  const float4 F4 = __ldg(reinterpret_cast<const float4*>(fl + c));
  float4 A4 = *reinterpret_cast<const float4*>(u0 + c);
  float4 B4 = *reinterpret_cast<const float4*>(u1 + c);
  float4 Pd4 = P4, Fd4 = F4;                 // row below (the cell itself at j = 0, Q13)
  if (j > 0) {
    Pd4 = __ldg(reinterpret_cast<const float4*>(p + c - W));
    Fd4 = __ldg(reinterpret_cast<const float4*>(fl + c - W));
  }
  float pl = 0.f, fll = F4.x;                // left neighbour of the first cell
  if (i0 > 0) { pl = __ldg(p + c - 1); fll = __ldg(fl + c - 1); }
  const float P[4] = {P4.x, P4.y, P4.z, P4.w}, F[4] = {F4.x, F4.y, F4.z, F4.w};
  const float Pd[4] = {Pd4.x, Pd4.y, Pd4.z, Pd4.w}, Fd[4] = {Fd4.x, Fd4.y, Fd4.z, Fd4.w};
  float a[4] = {A4.x, A4.y, A4.z, A4.w}, b[4] = {B4.x, B4.y, B4.z, B4.w};
  const bool rborder = (j < 1) | (j > H - 2);
  const unsigned char rb = m.rows ? m.rows[j] : (unsigned char)3;
  const bool masked = m.u0bc && (rb & 1);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int i = i0 + k;
    const float fc = F[k];
    const float fl_l = k == 0 ? fll : F[k - 1], p_l = k == 0 ? pl : P[k - 1];
    const float fl_d = j > 0 ? Fd[k] : fc;
    if (!(rborder | (i < 1) | (i > W - 2))) {
      a[k] = velocity_update_apply(a[k], fc, fl_l, P[k], p_l);
      b[k] = velocity_update_apply(b[k], fc, fl_d, P[k], Pd[k]);
    }
    if (wall_bcs) {
      a[k] = wall_bcs_apply(a[k], fc, fl_l);
      b[k] = wall_bcs_apply(b[k], fc, fl_d);
    }
    if (masked) {
      a[k] = const_vals_apply(a[k], __ldg(m.u0inv + c + k), __ldg(m.u0bc + c + k));
      b[k] = const_vals_apply(b[k], __ldg(m.u1inv + c + k), __ldg(m.u1bc + c + k));
    }
  }
  *reinterpret_cast<float4*>(u0 + c) = make_float4(a[0], a[1], a[2], a[3]);
  *reinterpret_cast<float4*>(u1 + c) = make_float4(b[0], b[1], b[2], b[3]);
}

}  // namespace s2
}  // namespace fnx

using namespace fnx;
using namespace fnx::s2;

namespace {
inline const float* vb(const float* p, long long off) { return p ? p - off : nullptr; }
inline float* vb(float* p, long long off) { return p ? p - off : nullptr; }
}  // namespace

bool fnx_step2d_supported(int H, int W) { return H >= 4 && W >= 4 && H < 65536 && W < 65536; }

size_t fnx_step2d_tile_ws_ints(const fnx_step2d_win& w, int B) {
  const int tiles_x = (w.W + TX - 1) / TX, tiles_y = (w.row1 - w.row0 + 7) / 8;   // the smallest tile height
  return (size_t)B * ((size_t)tiles_x * tiles_y + 1);
}

int fnx_step2d_check_window(const fnx_step2d_win& w) {
  if (w.H < 4 || w.W < 4) return fnx_set_error(FNX_ERR_ARG, "step2d: grid too small");
  if (!(0 <= w.ya0 && w.ya0 <= w.row0 && w.row0 < w.row1 && w.row1 <= w.ya1 && w.ya1 <= w.H))
    return fnx_set_error(FNX_ERR_ARG, "step2d: bad window rows [%d,%d) in memory rows [%d,%d) of %d", w.row0, w.row1,
                         w.ya0, w.ya1, w.H);
  // a computed row reads its +-1 neighbours and the forward apron reads +-1 around +-2: three existing
  // rows beyond the window on every interior side
  if ((w.ya0 > 0 && w.row0 - w.ya0 < 3) || (w.ya1 < w.H && w.ya1 - w.row1 < 3))
    return fnx_set_error(FNX_ERR_ARG, "step2d: a slab needs >= 3 rows of memory beyond the computed window");
  return FNX_OK;
}

int fnx_step2d_advect(const fnx_step2d_win& w, float dt, float maccormack_strength, int sample_outside,
                      const float* rho, const float* U, const float* flags, const fnx_step2d_masks& mk, int rho_passes,
                      float* rho_out, float* rho_mid, float* U_out, int B, int* tile_ws, cudaStream_t st) {
  const long long off = (long long)w.ya0 * w.W;
  const long long plane = (long long)(w.ya1 - w.ya0) * w.W;   // elements per channel / batch item of a 1-channel field
  // tile height: 32 rows, or 16 / 8 on small grids (fewer than ~8 / ~2 tiles of 32 rows per SM slot)
  const int tiles_x = (w.W + TX - 1) / TX;
  const int n32 = tiles_x * ((w.row1 - w.row0 + 31) / 32);
  const int th = n32 > 1184 ? 32 : (n32 > 296 ? 16 : 8);
  const int tiles_y = (w.row1 - w.row0 + th - 1) / th;
  const int ntiles = tiles_x * tiles_y;
  for (int b = 0; b < B; b++) {
    Adv a;
    a.H = w.H; a.W = w.W; a.row0 = w.row0; a.row1 = w.row1;
    a.ycl = w.ya0; a.ych = (w.H - 2 < w.ya1 - 2) ? w.H - 2 : w.ya1 - 2;
    a.yrl = w.ya0; a.yrh = (w.H - 1 < w.ya1 - 1) ? w.H - 1 : w.ya1 - 1;
    a.yf0 = w.ya0 == 0 ? 0 : w.ya0 + 1;
    a.yf1 = w.ya1 == w.H ? w.H : w.ya1 - 1;
    a.dt = dt; a.mdt = -dt; a.hs = maccormack_strength * 0.5f;
    a.sample_outside = sample_outside;
    a.rho = vb(rho + b * plane, off);
    a.u0 = vb(U + (2 * b) * plane, off);
    a.u1 = vb(U + (2 * b + 1) * plane, off);
    a.fl = vb(flags + b * plane, off);
    Masks m;
    m.u0bc = vb(mk.UBC ? mk.UBC + (2 * b) * plane : nullptr, off);
    m.u1bc = vb(mk.UBC ? mk.UBC + (2 * b + 1) * plane : nullptr, off);
    m.u0inv = vb(mk.UBCInv ? mk.UBCInv + (2 * b) * plane : nullptr, off);
    m.u1inv = vb(mk.UBCInv ? mk.UBCInv + (2 * b + 1) * plane : nullptr, off);
    m.rbc = vb(mk.rBC ? mk.rBC + b * plane : nullptr, off);
    m.rinv = vb(mk.rBCInv ? mk.rBCInv + b * plane : nullptr, off);
    m.rows = mk.rows ? mk.rows + (long long)b * (w.ya1 - w.ya0) - w.ya0 : nullptr;
    float* ro = vb(rho_out + b * plane, off);
    float* rm = vb(rho_mid ? rho_mid + b * plane : nullptr, off);
    float* uo0 = vb(U_out + (2 * b) * plane, off);
    float* uo1 = vb(U_out + (2 * b + 1) * plane, off);
#define FNX_ADV_LAUNCH(T)                                                                                              \
  do {                                                                                                               \
    if (tile_ws) {                                                                                                   \
      k2_advect_clean<T><<<ntiles, NT, 0, st>>>(a, m, rho_passes, w.ya0, w.ya1, ro, rm, uo0, uo1, tiles_x, count, list); \
      k2_advect<T><<<ntiles, NT, 0, st>>>(a, m, rho_passes, ro, rm, uo0, uo1, tiles_x, count, list);                 \
    } else {                                                                                                         \
      k2_advect<T><<<ntiles, NT, 0, st>>>(a, m, rho_passes, ro, rm, uo0, uo1, tiles_x, nullptr, nullptr);            \
    }                                                                                                                \
  } while (0)
    // interior fast path first; the tiles it declines come back in a list for the generic kernel
    int* count = tile_ws ? tile_ws + (size_t)b * (ntiles + 1) : nullptr;
    int* list = tile_ws ? count + 1 : nullptr;
    if (tile_ws && cudaMemsetAsync(count, 0, sizeof(int), st) != cudaSuccess)
      return fnx_set_error(FNX_ERR_CUDA, "step2d: memset failed");
    if (th == 32) FNX_ADV_LAUNCH(32);
    else if (th == 16) FNX_ADV_LAUNCH(16);
    else FNX_ADV_LAUNCH(8);
#undef FNX_ADV_LAUNCH
  }
  fnx_count_launches(tile_ws ? 2 * B : B);
  return FNX_OK;
}

int fnx_step2d_forces_div(const fnx_step2d_win& w, const fnx_step_params* prm, const float* rho, const float* rho_mid,
                          const float* U, const float* flags, const fnx_step2d_masks& mk, float* U_out, float* div, int B,
                          cudaStream_t st) {
  const long long off = (long long)w.ya0 * w.W;
  const long long plane = (long long)(w.ya1 - w.ya0) * w.W;
  const int tiles_x = (w.W + FW * FWARPS - 1) / (FW * FWARPS), tiles_y = (w.row1 - w.row0 + FR - 1) / FR;
  Frc g;
  g.H = w.H; g.W = w.W; g.row0 = w.row0; g.row1 = w.row1;
  g.use_buoyancy = prm->use_buoyancy; g.use_gravity = prm->use_gravity; g.wall_bcs = prm->apply_wall_bcs;
  g.bs0 = prm->buoyancy3[0] * prm->dt; g.bs1 = prm->buoyancy3[1] * prm->dt;   // gravity*dt, one fp32 product
  g.gf0 = prm->gravity3[0] * prm->dt; g.gf1 = prm->gravity3[1] * prm->dt;
  g.rho_star = prm->rho_star;
  for (int b = 0; b < B; b++) {
    Masks m;
    m.u0bc = vb(mk.UBC ? mk.UBC + (2 * b) * plane : nullptr, off);
    m.u1bc = vb(mk.UBC ? mk.UBC + (2 * b + 1) * plane : nullptr, off);
    m.u0inv = vb(mk.UBCInv ? mk.UBCInv + (2 * b) * plane : nullptr, off);
    m.u1inv = vb(mk.UBCInv ? mk.UBCInv + (2 * b + 1) * plane : nullptr, off);
    m.rbc = nullptr; m.rinv = nullptr;
    m.rows = mk.rows ? mk.rows + (long long)b * (w.ya1 - w.ya0) - w.ya0 : nullptr;
    k2_forces_div<<<tiles_x * tiles_y, 32 * FWARPS, 0, st>>>(
        g, m, vb(rho + b * plane, off), vb(rho_mid ? rho_mid + b * plane : nullptr, off), vb(U + (2 * b) * plane, off),
        vb(U + (2 * b + 1) * plane, off), vb(flags + b * plane, off), vb(U_out + (2 * b) * plane, off),
        vb(U_out + (2 * b + 1) * plane, off), vb(div ? div + b * plane : nullptr, off), tiles_x);
  }
  fnx_count_launches(B);
  return FNX_OK;
}

int fnx_step2d_project(const fnx_step2d_win& w, const float* p, float* U, const float* flags, const fnx_step2d_masks& mk,
                       int wall_bcs, int B, cudaStream_t st) {
  const long long off = (long long)w.ya0 * w.W;
  const long long plane = (long long)(w.ya1 - w.ya0) * w.W;
  dim3 grid((w.W + 127) / 128, (w.row1 - w.row0 + 1) / 2);
  for (int b = 0; b < B; b++) {
    Masks m;
    m.u0bc = vb(mk.UBC ? mk.UBC + (2 * b) * plane : nullptr, off);
    m.u1bc = vb(mk.UBC ? mk.UBC + (2 * b + 1) * plane : nullptr, off);
    m.u0inv = vb(mk.UBCInv ? mk.UBCInv + (2 * b) * plane : nullptr, off);
    m.u1inv = vb(mk.UBCInv ? mk.UBCInv + (2 * b + 1) * plane : nullptr, off);
    m.rbc = nullptr; m.rinv = nullptr;
    m.rows = mk.rows ? mk.rows + (long long)b * (w.ya1 - w.ya0) - w.ya0 : nullptr;
    const float* pp = vb(p + b * plane, off);
    float* pu0 = vb(U + (2 * b) * plane, off);
    float* pu1 = vb(U + (2 * b + 1) * plane, off);
    const float* pfl = vb(flags + b * plane, off);
    const bool vec = (w.W % 4 == 0) && ((((uintptr_t)pp | (uintptr_t)pu0 | (uintptr_t)pu1 | (uintptr_t)pfl) & 15) == 0) &&
                     (w.row1 - w.row0) <= 65535;
    if (vec) {
      dim3 g4((w.W / 4 + 127) / 128, w.row1 - w.row0);
      k2_project4<<<g4, 128, 0, st>>>(w.H, w.W, w.row0, w.row1, pp, pu0, pu1, pfl, m, wall_bcs);
    } else {
      k2_project<<<grid, 256, 0, st>>>(w.H, w.W, w.row0, w.row1, pp, pu0, pu1, pfl, m, wall_bcs);
    }
  }
  fnx_count_launches(B);
  return FNX_OK;
}
