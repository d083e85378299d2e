// stencil_device.cuh -- per-cell device functions of the 1-neighbour stencils.
// One function = the value ONE cell gets from ONE reference op, so the same code
// serves the per-op kernels (lib.fluid surface) and the fused step kernels.
//   addBuoyancy        pytorch/lib/fluid/source_terms.py:6-116      (Q14)
//   addGravity         source_terms.py:122-219
//   setWallBcs         set_wall_bcs.py:4-86                         (Q13)
//   velocityDivergence velocity_divergence.py:4-74                  (Q15)
//   velocityUpdate     velocity_update.py:6-162                     (Q12)
//   Jacobi             cpp/fluids_init.cpp:858-1003                 (Q16)
//   setConstVals       ../simulate.py:4-26
#pragma once
#include "fluid_common.cuh"

namespace fnx {

// offset of the lower neighbour along axis c
__device__ __forceinline__ long long nb_off(const Grid& g, int c) {
  return c == 0 ? 1LL : (c == 1 ? (long long)g.sy : g.sz);
}

// buoyancy increment for component c at an INTERIOR cell; fc/fn = flags of the cell and its
// lower neighbour along c, rc/rn the densities; strength_c = gravity[c]*dt (source_terms.py:70-75)
__device__ __forceinline__ float buoyancy_apply(float u, float fc, float fn, float rc, float rn,
                                                float strength_c, float rho_star) {
  if (fc != kFluid || fn != kFluid) return u;
  float factor = strength_c * (0.5f * (rc + rn) - rho_star);
  return u + factor;
}

// gravity increment (source_terms.py:176-216): cell Fluid|Empty, neighbour Fluid, or Empty with
// a Fluid cell
__device__ __forceinline__ float gravity_apply(float u, float fc, float fn, float force_c) {
  bool cf = fc == kFluid, ce = fc == kEmpty;
  if (!cf && !ce) return u;
  if (fn == kFluid || (fn == kEmpty && cf)) return u + force_c;
  return u;
}

// setWallBcs for component c; fn = flag of the lower neighbour, or of the cell itself when the
// cell's index along c is 0 (the Python version lost the i<=0 guard, Q13)
__device__ __forceinline__ float wall_bcs_apply(float u, float fc, float fn) {
  bool cf = fc == kFluid, co = fc == kObstacle;
  if (!cf && !co) return u;
  if (fn == kObstacle || (co && fn == kFluid)) return 0.f;
  return u;
}

// velocityUpdate for component c at an INTERIOR cell: the sum of masked products the reference
// forms (velocity_update.py:143-149), kept literal so NaN/Inf propagate identically
__device__ __forceinline__ float velocity_update_apply(float u, float fc, float fn, float P,
                                                       float Pn) {
  bool cf = fc == kFluid;
  bool ce = (fc == kEmpty) && (fc != kOutflow);
  float m1 = (cf && fn == kFluid) ? 1.f : 0.f;
  float m2 = (cf && fn == kEmpty) ? 1.f : 0.f;
  float m3 = (ce && fn == kFluid) ? 1.f : 0.f;
  float m4 = (ce && fn == kEmpty) ? 1.f : 0.f;
  return m1 * (u - (P - Pn)) + m2 * (u - P) + m3 * (u + Pn) + m4 * 0.f;
}

// setConstVals: x*inv_mask + bc
__device__ __forceinline__ float const_vals_apply(float x, float inv_mask, float bc) {
  return x * inv_mask + bc;
}

// flagsToOccupancy
__device__ __forceinline__ float occupancy_of(float f) {
  return f == kFluid ? 0.f : (f == kObstacle ? 1.f : f);
}

}  // namespace fnx
