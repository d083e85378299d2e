// advect_device.cuh -- per-cell device functions of the semi-Lagrangian / MacCormack
// advection (one thread = one cell).  Semantics follow the reference's ATen path:
//   interpolation        pytorch/lib/fluid/cpp/grid.cpp:13-76, 118-269, 448-511
//   MAC averages         grid.cpp:274-446
//   line trace           calc_line_trace.cpp:16-64, 73-149, 154-257, 259-424
// including its quirks.  NA = number of active axes
// (2 or 3): in 2-D every z term of the reference is neutral (pos.z = 0.5, delta.z = 0),
// so it is dropped.
#pragma once
#include "fluid_common.cuh"

namespace fnx {

template <bool Z>
struct Taps {
  long long o;  // offset of corner (z0, y0, x0)
  float s0, s1, t0, t1, f0, f1;
};

// grid.cpp:28-52: p = pos - 0.5; i0 = trunc(p); weights clamped to [0,1] BEFORE the index clamp,
// so an index clamped at the upper edge keeps its fractional weight (Q3).
template <bool Z>
__device__ __forceinline__ Taps<Z> make_taps(const Grid& g, const float* pos) {
  Taps<Z> t;
  float px = pos[0] - 0.5f, py = pos[1] - 0.5f;
  // every position that can index a real grid takes the 32-bit path; the int64 code below it is the
  // same arithmetic for the (garbage-velocity) positions beyond 1e9 and for NaN
  bool small = (fabsf(px) < 1.0e9f) & (fabsf(py) < 1.0e9f);
  if (Z) small = small & (fabsf(pos[2] - 0.5f) < 1.0e9f);
  if (small) {
    const int ix = __float2int_rz(px), iy = __float2int_rz(py);
    const float s1 = px - (float)ix, t1 = py - (float)iy;
    const float s0 = 1.f - s1, t0 = 1.f - t1;
    int x0 = ix < 0 ? 0 : ix, y0 = iy < 0 ? 0 : iy, z0 = 0;
    x0 = x0 > g.W - 2 ? g.W - 2 : x0;
    y0 = y0 > g.H - 2 ? g.H - 2 : y0;
    t.s1 = clamp01(s1); t.t1 = clamp01(t1);
    t.s0 = clamp01(s0); t.t0 = clamp01(t0);
    t.f0 = 1.f; t.f1 = 0.f;
    if (Z) {
      const float pz = pos[2] - 0.5f;
      const int iz = __float2int_rz(pz);
      const float f1 = pz - (float)iz, f0 = 1.f - f1;
      z0 = iz < 0 ? 0 : iz;
      z0 = z0 > g.D - 2 ? g.D - 2 : z0;
      t.f1 = clamp01(f1); t.f0 = clamp01(f0);
    }
    t.o = (z0 * g.H + y0) * g.W + x0;  // < 2^31 cells per batch item (checked at the entry points)
    return t;
  }
  long long ix = trunc_ll(px), iy = trunc_ll(py);
  float s1 = px - (float)ix, t1 = py - (float)iy;
  float s0 = 1.f - s1, t0 = 1.f - t1;
  long long x0 = clampll(ix, 0, g.W - 2), y0 = clampll(iy, 0, g.H - 2), z0 = 0;
  t.s1 = clamp01(s1); t.t1 = clamp01(t1);
  t.s0 = clamp01(s0); t.t0 = clamp01(t0);
  t.f0 = 1.f; t.f1 = 0.f;
  if (Z) {
    float pz = pos[2] - 0.5f;
    long long iz = trunc_ll(pz);
    float f1 = pz - (float)iz, f0 = 1.f - f1;
    z0 = clampll(iz, 0, g.D - 2);
    t.f1 = clamp01(f1); t.f0 = clamp01(f0);
  }
  t.o = (z0 * g.H + y0) * g.W + x0;
  return t;
}

// bi/tri-linear sample, grid.cpp:54-75 (interpol) == :491-510 (interpolComponent)
template <bool Z>
__device__ __forceinline__ float sample_field(const Grid& g, const float* __restrict__ f,
                                              const float* pos) {
  Taps<Z> t = make_taps<Z>(g, pos);
  const float* p0 = f + t.o;
  float Ia = __ldg(p0), Ib = __ldg(p0 + g.sy), Ic = __ldg(p0 + 1), Id = __ldg(p0 + g.sy + 1);
  float lo = (Ia * t.t0 + Ib * t.t1) * t.s0 + (Ic * t.t0 + Id * t.t1) * t.s1;
  if (!Z) return lo;
  const float* p1 = p0 + g.sz;
  float Ie = __ldg(p1), If = __ldg(p1 + g.sy), Ig = __ldg(p1 + 1), Ih = __ldg(p1 + g.sy + 1);
  float hi = (Ie * t.t0 + If * t.t1) * t.s0 + (Ig * t.t0 + Ih * t.t1) * t.s1;
  return lo * t.f0 + hi * t.f1;
}

// grid.cpp:78-96 interpol1DWithFluid: drop the non-fluid end(s) of a 1-D lerp
__device__ __forceinline__ void mix_fluid(float va, bool fa, float vb, bool fb, float wa, float wb,
                                          float& v, bool& fl) {
  if (!fa && !fb) { v = 0.f; fl = false; }
  else if (!fa) { v = vb; fl = true; }
  else if (!fb) { v = va; fl = true; }
  else { v = va * wa + vb * wb; fl = true; }
}

// grid.cpp:118-269 interpolWithFluid; falls back to the plain sample when no corner is fluid
template <bool Z>
__device__ __forceinline__ float sample_with_fluid(const Grid& g, const float* __restrict__ f,
                                                   const float* __restrict__ flags,
                                                   const float* pos) {
  Taps<Z> t = make_taps<Z>(g, pos);
  const float* p0 = f + t.o;
  const float* q0 = flags + t.o;
  float ab, cd, v;
  bool fab, fcd, fv;
  mix_fluid(__ldg(p0), __ldg(q0) == kFluid, __ldg(p0 + g.sy), __ldg(q0 + g.sy) == kFluid, t.t0, t.t1, ab, fab);
  mix_fluid(__ldg(p0 + 1), __ldg(q0 + 1) == kFluid, __ldg(p0 + g.sy + 1), __ldg(q0 + g.sy + 1) == kFluid,
            t.t0, t.t1, cd, fcd);
  mix_fluid(ab, fab, cd, fcd, t.s0, t.s1, v, fv);
  if (Z) {
    const float* p1 = p0 + g.sz;
    const float* q1 = q0 + g.sz;
    float ef, gh, w, lo = v;
    bool fef, fgh, fw, flo = fv;
    mix_fluid(__ldg(p1), __ldg(q1) == kFluid, __ldg(p1 + g.sy), __ldg(q1 + g.sy) == kFluid, t.t0, t.t1, ef, fef);
    mix_fluid(__ldg(p1 + 1), __ldg(q1 + 1) == kFluid, __ldg(p1 + g.sy + 1), __ldg(q1 + g.sy + 1) == kFluid,
              t.t0, t.t1, gh, fgh);
    mix_fluid(ef, fef, gh, fgh, t.s0, t.s1, w, fw);
    mix_fluid(lo, flo, w, fw, t.f0, t.f1, v, fv);
  }
  if (!fv) return sample_field<Z>(g, f, pos);
  return v;
}

// ---- velocity averages (grid.cpp:274-446) -------------------------------------
// cell-centred velocity, getCentered :300-309
template <bool Z>
__device__ __forceinline__ void centered_vel(const Grid& g, const float* __restrict__ U, long long o,
                                             float* c) {
  c[0] = 0.5f * (__ldg(U + o) + __ldg(U + o + 1));
  c[1] = 0.5f * (__ldg(U + g.n + o) + __ldg(U + g.n + o + g.sy));
  if (Z) c[2] = 0.5f * (__ldg(U + 2 * g.n + o) + __ldg(U + 2 * g.n + o + g.sz));
}

// velocity at the face centre of component `comp`, getAtMACX/Y/Z :314-446
// (sum order ((a+b)+c)+d as written in the reference)
template <bool Z>
__device__ __forceinline__ void mac_vel(const Grid& g, const float* __restrict__ U, int comp,
                                        long long o, float* v) {
  const float* U0 = U;
  const float* U1 = U + g.n;
  const float* U2 = U + 2 * g.n;
  const long long sy = g.sy, sz = g.sz;
  if (comp == 0) {
    v[0] = __ldg(U0 + o);
    v[1] = 0.25f * (((__ldg(U1 + o) + __ldg(U1 + o - 1)) + __ldg(U1 + o + sy)) + __ldg(U1 + o + sy - 1));
    if (Z) v[2] = 0.25f * (((__ldg(U2 + o) + __ldg(U2 + o - 1)) + __ldg(U2 + o + sz)) + __ldg(U2 + o + sz - 1));
  } else if (comp == 1) {
    v[0] = 0.25f * (((__ldg(U0 + o) + __ldg(U0 + o - sy)) + __ldg(U0 + o + 1)) + __ldg(U0 + o - sy + 1));
    v[1] = __ldg(U1 + o);
    if (Z) v[2] = 0.25f * (((__ldg(U2 + o) + __ldg(U2 + o - sy)) + __ldg(U2 + o + sz)) + __ldg(U2 + o + sz - sy));
  } else {
    v[0] = 0.25f * (((__ldg(U0 + o) + __ldg(U0 + o - sz)) + __ldg(U0 + o + 1)) + __ldg(U0 + o - sz + 1));
    v[1] = 0.25f * (((__ldg(U1 + o) + __ldg(U1 + o - sz)) + __ldg(U1 + o + sy)) + __ldg(U1 + o - sz + sy));
    v[2] = __ldg(U2 + o);
  }
}

// ---- line trace (calc_line_trace.cpp) --------------------------------------------
template <int NA>
__device__ __forceinline__ bool out_of_domain(const Grid& g, const float* q) {  // :16-27
  bool m = q[0] <= 0.f || q[0] >= (float)g.W || q[1] <= 0.f || q[1] >= (float)g.H;
  if (NA == 3) m = m || q[2] <= 0.f || q[2] >= (float)g.D;
  return m;
}

template <int NA>
__device__ __forceinline__ bool blocked_cell(const Grid& g, const float* __restrict__ flags,
                                             const float* q) {  // :33-64
  if (out_of_domain<NA>(g, q)) return false;
  long long o = (long long)__float2int_rz(q[1]) * g.W + __float2int_rz(q[0]);
  if (NA == 3) o += (long long)__float2int_rz(q[2]) * g.sz;
  return __ldg(flags + o) != kFluid;
}

// :175-257 calcRayBorderIntersection, evaluated from the trace's START position (Q11)
template <int NA>
__device__ __forceinline__ bool ray_border(const Grid& g, const float* pos, const float* next,
                                           float* ipos) {
  float min_step = CUDART_INF_F;
  const float dimf[3] = {(float)g.W, (float)g.H, (float)g.D};
#pragma unroll
  for (int a = 0; a < NA; a++) {
    if (next[a] <= kHitMargin) {
      float d = next[a] - pos[a];
      if (fabsf(d) >= kEpsilon) min_step = min_t(min_step, (kHitMargin - pos[a]) / d);
    }
  }
#pragma unroll
  for (int a = 0; a < NA; a++) {
    float lim = dimf[a] - kHitMargin;
    if (next[a] >= lim) {
      float d = next[a] - pos[a];
      if (fabsf(d) >= kEpsilon) min_step = min_t(min_step, (lim - pos[a]) / d);
    }
  }
  bool hit = (min_step >= 0.f) && (min_step < CUDART_INF_F);
#pragma unroll
  for (int a = 0; a < NA; a++) ipos[a] = hit ? (min_step * (next[a] - pos[a]) + pos[a]) : 0.f;
  return hit;
}

// :73-149 HitBoundingBox exactly as the ATen code evaluates it (Q10)
template <int NA>
__device__ __forceinline__ bool hit_bounding_box(const float* minB, const float* maxB,
                                                 const float* o, const float* dir, float* coord) {
  bool mid[NA], inside = true;
  float cand[NA], maxT[NA];
#pragma unroll
  for (int a = 0; a < NA; a++) {
    bool lt = o[a] < minB[a], gt = o[a] > maxB[a];
    mid[a] = (o[a] >= minB[a]) && (o[a] <= maxB[a]);
    cand[a] = 0.f;
    if (lt) cand[a] = minB[a];
    if (gt) cand[a] = maxB[a];
    if (lt || gt) inside = false;
  }
  const bool outside = !inside;
#pragma unroll
  for (int a = 0; a < NA; a++) {
    maxT[a] = 0.f;
    if (outside && !mid[a] && dir[a] != 0.f) maxT[a] = (cand[a] - o[a]) / dir[a];
    if ((outside && mid[a]) || dir[a] == 0.f) maxT[a] = -1.f;
  }
  int wp = 0;
#pragma unroll
  for (int a = 1; a < NA; a++)
    if (maxT[a] > maxT[wp]) wp = a;  // argmax keeps the first maximum
  float T = 0.f;
#pragma unroll
  for (int a = 0; a < NA; a++)
    if (a == wp) T = maxT[a];
  bool ret = !(T < 0.f && outside);
  const float err_tol = 1e-6f;
#pragma unroll
  for (int a = 0; a < NA; a++) coord[a] = (a == wp) ? cand[a] : (o[a] + T * dir[a]);
#pragma unroll
  for (int a = 0; a < NA; a++)
    if (a != wp && (coord[a] < minB[a] - err_tol || coord[a] > maxB[a] + err_tol)) ret = false;
  return ret;
}

// :259-424 calcLineTrace for one cell: march `pos` along `delta` in <= 1-cell steps, stopping
// hit_margin before the domain border or the first non-fluid cell.
template <int NA>
__device__ __forceinline__ void line_trace(const Grid& g, const float* __restrict__ flags,
                                           const float* pos, const float* delta, float* new_pos) {
#pragma unroll
  for (int a = 0; a < NA; a++) new_pos[a] = pos[a];
  if (out_of_domain<NA>(g, pos) || blocked_cell<NA>(g, flags, pos)) return;
  // at::norm(2, dim=1): acc += x*x in fp32, then sqrt
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < NA; a++) acc = acc + delta[a] * delta[a];
  const float length = sqrtf(acc);
  if (length <= kEpsilon) return;
  // An infinite displacement (|delta| beyond 1.8e19: garbage input) has direction delta / inf = 0: the march
  // below would never move and never end -- the reference's loop (calc_line_trace.cpp:310) does not end
  // either, so there is no result to match; keep the start position instead of hanging the GPU.
  if (!(length < CUDART_INF_F)) return;
  float dir[NA], next[NA];
#pragma unroll
  for (int a = 0; a < NA; a++) dir[a] = delta[a] / length;
  float cur_length = 0.f;
  while (cur_length < length - kHitMargin) {  // :310-314
    const float rem = length - cur_length;
    const float cur_step = rem < 1.f ? rem : 1.f;
#pragma unroll
    for (int a = 0; a < NA; a++) next[a] = new_pos[a] + dir[a] * cur_step;
    // case 1: the step leaves the grid (:323-361)
    if (out_of_domain<NA>(g, next)) {
      float ipos[NA];
      bool hit = ray_border<NA>(g, pos, next, ipos);
      if (!hit) {  // clampToDomain is a no-op in the reference (Q9)
#pragma unroll
        for (int a = 0; a < NA; a++) ipos[a] = next[a];
      }
      if (!blocked_cell<NA>(g, flags, ipos)) {
#pragma unroll
        for (int a = 0; a < NA; a++) new_pos[a] = ipos[a];
        return;
      }
#pragma unroll
      for (int a = 0; a < NA; a++) next[a] = ipos[a];
    }
    // case 2: the step enters a blocked cell (:363-412): back off to the cell's inflated box
    if (blocked_cell<NA>(g, flags, next)) {
      bool stopped = false;
      for (int count = 0; count < 4; count++) {
        if (!blocked_cell<NA>(g, flags, next)) break;
        float bmin[NA], bmax[NA], ipos[NA];
#pragma unroll
        for (int a = 0; a < NA; a++) {
          float ctr = (float)__float2int_rz(next[a]) + 0.5f;
          bmin[a] = ctr - 0.5f - kHitMargin;
          bmax[a] = ctr + 0.5f + kHitMargin;
        }
        if (!hit_bounding_box<NA>(bmin, bmax, new_pos, dir, ipos)) { stopped = true; break; }
#pragma unroll
        for (int a = 0; a < NA; a++) next[a] = ipos[a];
      }
      if (!stopped) {
#pragma unroll
        for (int a = 0; a < NA; a++) new_pos[a] = next[a];
      }
      return;
    }
    // case 3: free step (:415-420)
#pragma unroll
    for (int a = 0; a < NA; a++) new_pos[a] = next[a];
    cur_length = cur_length + cur_step;
  }
}

}  // namespace fnx
