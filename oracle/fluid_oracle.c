/*
 * fluid_oracle.c -- CPU restatement (plain C99, one loop body per cell) of the
 * per-timestep fluid path of jolibrain/fluidnet_cxx.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker: it is compiled to
 * oracle/_build/libfluid_oracle.so and may be loaded only by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg.  The product
 * (fluidnet_cxx_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED for 2-D -- tests/test_oracle_golden.py checks every
 * function below bit-for-bit (fp32) against outputs of the reference's own ATen
 * CPU path (patched build, oracle/build_ref.py) stored in tests/golden/.
 * UNPINNED for 3-D: the reference asserts 3-D off (advection.py:58,108) and its
 * 3-D branches carry shadowing bugs; the 3-D branches here state
 * the intended per-cell semantics (2-D rules extended to z, D=1 reduces to the
 * pinned 2-D case).
 *
 * All arithmetic is fp32 in the reference's operation order; compile with
 * -ffp-contract=off so no FMA is formed (ATen evaluates each tensor op with a
 * separate rounding).  Layout: (B, C, D, H, W) contiguous, flags are fp32.
 *
 * Reference files (under /root/reference/pytorch/lib/fluid/):
 *   cpp/fluids_init.cpp, cpp/grid.cpp, cpp/calc_line_trace.cpp,
 *   velocity_divergence.py, velocity_update.py, set_wall_bcs.py, source_terms.py,
 *   flags_to_occupancy.py, ../simulate.py, ../multi_scale_net.py, ../model.py
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TYPE_FLUID 1.0f    /* cell_type.py:7  */
#define TYPE_OBSTACLE 2.0f /* cell_type.py:8  */
#define TYPE_EMPTY 4.0f    /* cell_type.py:9  */
#define TYPE_OUTFLOW 16.0f /* cell_type.py:11 */

static const float hit_margin = 1e-5f; /* calc_line_trace.cpp:7 */
static const float epsilon = 1e-12f;   /* calc_line_trace.cpp:8 */

typedef struct {
  int B, D, H, W, is3d;
} grid_t;

/* coverage counters (tests assert the golden cases really reach the rare branches):
   [0] traces run  [1] case-1 border exits  [2] case-2 blocked-cell hits  [3] ray-box misses
   [4] ray-box origin inside box (Q10)  [5] unit steps taken  [6] fluid-aware fallbacks to interpol */
static long trace_stats[8];
void orc_trace_stats(long *out, int reset) {
  for (int q = 0; q < 8; q++) { out[q] = trace_stats[q]; if (reset) trace_stats[q] = 0; }
}

static inline size_t nvox(const grid_t *g) { return (size_t)g->D * g->H * g->W; }
static inline size_t at(const grid_t *g, int k, int j, int i) {
  return ((size_t)k * g->H + j) * g->W + i;
}
static inline int is_border(const grid_t *g, int k, int j, int i, int bnd) {
  /* fluids_init.cpp:313-320 */
  int m = (i < bnd) || (i > g->W - 1 - bnd) || (j < bnd) || (j > g->H - 1 - bnd);
  if (g->is3d) m = m || (k < bnd) || (k > g->D - 1 - bnd);
  return m;
}
static inline float clamp01(float x) { return x < 0.f ? 0.f : (x > 1.f ? 1.f : x); }
static inline long clampl(long x, long lo, long hi) {
  /* at::clamp with lo > hi yields hi (torch semantics: min(max(x,lo),hi)) */
  if (x < lo) x = lo;
  if (x > hi) x = hi;
  return x;
}
static inline float fminf_t(float a, float b) { return (b < a) ? b : a; }
static inline float fmaxf_t(float a, float b) { return (b > a) ? b : a; }

/* ------------------------------------------------------------------------ */
/* grid.cpp:13-76 interpol / :448-511 interpolComponent (same arithmetic)     */
/* ------------------------------------------------------------------------ */
typedef struct {
  long x0, y0, z0;
  float s0, s1, t0, t1, f0, f1;
} tap_t;

static inline tap_t make_tap(const grid_t *g, const float pos[3]) {
  tap_t t;
  float px = pos[0] - 0.5f, py = pos[1] - 0.5f, pz = pos[2] - 0.5f; /* grid.cpp:28 */
  long ix = (long)px, iy = (long)py, iz = (long)pz;                 /* trunc, :31 */
  float s1 = px - (float)ix, t1 = py - (float)iy, f1 = pz - (float)iz;
  float s0 = 1.f - s1, t0 = 1.f - t1, f0 = 1.f - f1; /* :36-38 before the clamps */
  t.x0 = clampl(ix, 0, g->W - 2);                    /* :40-42 */
  t.y0 = clampl(iy, 0, g->H - 2);
  t.z0 = clampl(iz, 0, g->D - 2); /* D==1 -> -1, wraps to plane 0 (Q3) */
  if (t.z0 < 0) t.z0 += g->D;
  t.s1 = clamp01(s1); t.t1 = clamp01(t1); t.f1 = clamp01(f1);
  t.s0 = clamp01(s0); t.t0 = clamp01(t0); t.f0 = clamp01(f0);
  return t;
}

static inline float interp_field(const grid_t *g, const float *f, const float pos[3]) {
  tap_t t = make_tap(g, pos);
  const float *p0 = f + at(g, (int)t.z0, (int)t.y0, (int)t.x0);
  size_t sy = g->W;
  float Ia = p0[0], Ib = p0[sy], Ic = p0[1], Id = p0[sy + 1];
  float lo = (Ia * t.t0 + Ib * t.t1) * t.s0 + (Ic * t.t0 + Id * t.t1) * t.s1;
  if (!g->is3d) return lo; /* grid.cpp:74 */
  const float *p1 = p0 + (size_t)g->H * g->W;
  float Ie = p1[0], If = p1[sy], Ig = p1[1], Ih = p1[sy + 1];
  float hi = (Ie * t.t0 + If * t.t1) * t.s0 + (Ig * t.t0 + Ih * t.t1) * t.s1;
  return lo * t.f0 + hi * t.f1; /* grid.cpp:66-67 */
}

/* grid.cpp:78-96 interpol1DWithFluid */
static inline void mix1d(float va, int fa, float vb, int fb, float wa, float wb, float *v, int *fl) {
  if (!fa && !fb) { *v = 0.f; *fl = 0; }
  else if (!fa) { *v = vb; *fl = 1; }
  else if (!fb) { *v = va; *fl = 1; }
  else { *v = va * wa + vb * wb; *fl = 1; }
}

/* grid.cpp:118-269 interpolWithFluid */
static inline float interp_with_fluid(const grid_t *g, const float *f, const float *flags,
                                      const float pos[3]) {
  tap_t t = make_tap(g, pos);
  size_t o = at(g, (int)t.z0, (int)t.y0, (int)t.x0), sy = g->W;
  float ab, cd, v;
  int fab, fcd, fv;
  mix1d(f[o], flags[o] == TYPE_FLUID, f[o + sy], flags[o + sy] == TYPE_FLUID, t.t0, t.t1, &ab, &fab);
  mix1d(f[o + 1], flags[o + 1] == TYPE_FLUID, f[o + sy + 1], flags[o + sy + 1] == TYPE_FLUID, t.t0,
        t.t1, &cd, &fcd);
  mix1d(ab, fab, cd, fcd, t.s0, t.s1, &v, &fv);
  if (g->is3d) {
    size_t o1 = o + (size_t)g->H * g->W;
    float ef, gh, w, lo = v;
    int fef, fgh, fw, flo = fv;
    mix1d(f[o1], flags[o1] == TYPE_FLUID, f[o1 + sy], flags[o1 + sy] == TYPE_FLUID, t.t0, t.t1, &ef, &fef);
    /* intended semantics: corners g,h test their own flags (reference grid.cpp:204-205
       reads the x0 column -- a 3-D-only defect) */
    mix1d(f[o1 + 1], flags[o1 + 1] == TYPE_FLUID, f[o1 + sy + 1], flags[o1 + sy + 1] == TYPE_FLUID,
          t.t0, t.t1, &gh, &fgh);
    mix1d(ef, fef, gh, fgh, t.s0, t.s1, &w, &fw);
    mix1d(lo, flo, w, fw, t.f0, t.f1, &v, &fv);
  }
  if (!fv) { trace_stats[6]++; return interp_field(g, f, pos); } /* grid.cpp:264-265 */
  return v;
}

/* ------------------------------------------------------------------------ */
/* calc_line_trace.cpp                                                        */
/* ------------------------------------------------------------------------ */
static inline int out_of_domain(const grid_t *g, const float q[3]) { /* :16-27 */
  return q[0] <= 0.f || q[0] >= (float)g->W || q[1] <= 0.f || q[1] >= (float)g->H || q[2] <= 0.f ||
         q[2] >= (float)g->D;
}
static inline int blocked_cell(const grid_t *g, const float *flags, const float q[3]) { /* :33-64 */
  if (out_of_domain(g, q)) return 0;
  long ix = (long)q[0], iy = (long)q[1], iz = (long)q[2];
  return flags[at(g, (int)iz, (int)iy, (int)ix)] != TYPE_FLUID;
}

/* :175-257 calcRayBorderIntersection (called with the trace's START pos, Q11) */
static int ray_border(const grid_t *g, const float pos[3], const float next[3], float ipos[3]) {
  float min_step = INFINITY;
  float dimf[3] = {(float)g->W, (float)g->H, (float)g->D};
  for (int a = 0; a < 3; a++) { /* left, front, bottom faces, :206-224 */
    if (next[a] <= hit_margin) {
      float d = next[a] - pos[a];
      if (fabsf(d) >= epsilon) {
        float st = (hit_margin - pos[a]) / d;
        min_step = fminf_t(min_step, st);
      }
    }
  }
  for (int a = 0; a < 3; a++) { /* right, back, upper faces, :231-249 */
    float lim = dimf[a] - hit_margin; /* (int64 - float) evaluated in float */
    if (next[a] >= lim) {
      float d = next[a] - pos[a];
      if (fabsf(d) >= epsilon) {
        float st = (lim - pos[a]) / d;
        min_step = fminf_t(min_step, st);
      }
    }
  }
  int hit = (min_step >= 0.f) && (min_step < INFINITY);
  for (int a = 0; a < 3; a++) ipos[a] = hit ? (min_step * (next[a] - pos[a]) + pos[a]) : 0.f;
  return hit;
}

/* :73-149 HitBoundingBox as the ATen code evaluates it (Q10) */
static int hit_bounding_box(const float minB[3], const float maxB[3], const float o[3],
                            const float dir[3], float coord[3]) {
  int mid[3], inside = 1;
  float cand[3], maxT[3];
  for (int a = 0; a < 3; a++) {
    int lt = o[a] < minB[a], gt = o[a] > maxB[a];
    mid[a] = (o[a] >= minB[a]) && (o[a] <= maxB[a]);
    cand[a] = 0.f;
    if (lt) cand[a] = minB[a];
    if (gt) cand[a] = maxB[a];
    if (lt || gt) inside = 0;
  }
  int outside = !inside;
  if (inside) trace_stats[4]++;
  for (int a = 0; a < 3; a++) {
    maxT[a] = 0.f;
    if (outside && !mid[a] && dir[a] != 0.f) maxT[a] = (cand[a] - o[a]) / dir[a];
    if ((outside && mid[a]) || dir[a] == 0.f) maxT[a] = -1.f;
  }
  int wp = 0;
  for (int a = 1; a < 3; a++)
    if (maxT[a] > maxT[wp]) wp = a; /* argmax: first maximal index */
  float T = maxT[wp];
  int ret = 1;
  if (T < 0.f && outside) ret = 0;
  const float err_tol = 1e-6f;
  for (int a = 0; a < 3; a++) coord[a] = (a == wp) ? cand[a] : (o[a] + T * dir[a]);
  for (int a = 0; a < 3; a++)
    if (a != wp && (coord[a] < minB[a] - err_tol || coord[a] > maxB[a] + err_tol)) ret = 0;
  return ret;
}

/* at::norm(2, dim=1) over the 3 displacement channels.  ATen's fp32 CPU reduction
   over a strided size-3 dim accumulates acc += x*x in float and takes sqrtf
   (pinned empirically by tests/test_oracle_golden.py). */
static inline float norm3(const float d[3]) {
  float acc = 0.f;
  acc = acc + d[0] * d[0];
  acc = acc + d[1] * d[1];
  acc = acc + d[2] * d[2];
  return sqrtf(acc);
}

/* :259-424 calcLineTrace, one cell.  Returns number of internal errors the
   reference would have asserted on (0 in valid runs). */
static int line_trace(const grid_t *g, const float *flags, const float pos[3], const float delta[3],
                      int do_trace, float new_pos[3]) {
  int errors = 0;
  if (!do_trace) { /* :264-267 */
    for (int a = 0; a < 3; a++) new_pos[a] = pos[a] + delta[a];
    return 0;
  }
  int cont = 1;
  trace_stats[0]++;
  if (out_of_domain(g, pos)) cont = 0;
  if (blocked_cell(g, flags, pos)) cont = 0;
  for (int a = 0; a < 3; a++) new_pos[a] = pos[a];
  float length = norm3(delta);
  if (length <= epsilon) cont = 0;
  /* an infinite displacement has direction delta/inf = 0: the reference's march (:310) never moves and never
   * ends, so there is no reference result; the CUDA path keeps the start position -- do the same here */
  if (!(length < INFINITY)) cont = 0;
  float dt[3] = {0.f, 0.f, 0.f};
  if (cont)
    for (int a = 0; a < 3; a++) dt[a] = delta[a] / length;
  float cur_length = 0.f, next[3] = {0.f, 0.f, 0.f};
  while (cont) {
    if (cur_length >= length - hit_margin) break; /* :310-314 */
    float rem = length - cur_length;
    float cur_step = rem < 1.f ? rem : 1.f; /* at::min(length-cur, 1) */
    for (int a = 0; a < 3; a++) next[a] = new_pos[a] + dt[a] * cur_step;
    /* case 1: next exits the grid, :323-361 */
    if (out_of_domain(g, next)) {
      float ipos[3];
      trace_stats[1]++;
      int hit = ray_border(g, pos, next, ipos);
      if (!hit) { /* clampToDomain is a no-op (Q9) */
        for (int a = 0; a < 3; a++) ipos[a] = next[a];
      }
      if (out_of_domain(g, ipos)) errors++; /* reference: "case 1 exited bounds!" */
      if (!blocked_cell(g, flags, ipos)) {
        for (int a = 0; a < 3; a++) new_pos[a] = ipos[a];
        cont = 0;
        break;
      }
      for (int a = 0; a < 3; a++) next[a] = ipos[a];
    }
    /* case 2: next enters a blocked cell, :363-412 */
    if (blocked_cell(g, flags, next)) {
      int count_mask = 1, stopped = 0;
      trace_stats[2]++;
      for (int count = 0; count <= 4; count++) {
        if (!blocked_cell(g, flags, next)) count_mask = 0;
        if (!count_mask) break;
        if (count >= 4) { errors++; break; } /* "Cannot find non-geometry point" */
        float ctr[3], bmin[3], bmax[3], ipos[3];
        for (int a = 0; a < 3; a++) {
          ctr[a] = (float)(long)next[a] + 0.5f;
          bmin[a] = ctr[a] - 0.5f - hit_margin; /* :158-159 */
          bmax[a] = ctr[a] + 0.5f + hit_margin;
        }
        int hit = hit_bounding_box(bmin, bmax, new_pos, dt, ipos);
        if (!hit) { stopped = 1; count_mask = 0; trace_stats[3]++; break; }
        for (int a = 0; a < 3; a++) next[a] = ipos[a];
      }
      if (!stopped)
        for (int a = 0; a < 3; a++) new_pos[a] = next[a];
      cont = 0;
      break;
    }
    /* otherwise advance, :415-420 */
    for (int a = 0; a < 3; a++) new_pos[a] = next[a];
    cur_length = cur_length + cur_step;
    trace_stats[5]++;
  }
  return errors;
}

/* ------------------------------------------------------------------------ */
/* velocity averages, grid.cpp:274-446                                        */
/* ------------------------------------------------------------------------ */
static inline void get_centered(const grid_t *g, const float *U, int k, int j, int i, float c[3]) {
  size_t n = nvox(g), o = at(g, k, j, i);
  c[0] = 0.5f * (U[o] + U[o + 1]);
  c[1] = 0.5f * (U[n + o] + U[n + o + g->W]);
  c[2] = g->is3d ? 0.5f * (U[2 * n + o] + U[2 * n + o + (size_t)g->H * g->W]) : 0.f;
}
static inline void get_at_mac(const grid_t *g, const float *U, int comp, int k, int j, int i,
                              float v[3]) {
  size_t n = nvox(g), o = at(g, k, j, i), sy = g->W, sz = (size_t)g->H * g->W;
  const float *U0 = U, *U1 = U + n, *U2 = U + 2 * n;
  if (comp == 0) { /* getAtMACX :314-357 */
    v[0] = U0[o];
    v[1] = 0.25f * (((U1[o] + U1[o - 1]) + U1[o + sy]) + U1[o + sy - 1]);
    v[2] = g->is3d ? 0.25f * (((U2[o] + U2[o - 1]) + U2[o + sz]) + U2[o + sz - 1]) : 0.f;
  } else if (comp == 1) { /* getAtMACY :359-402 */
    v[0] = 0.25f * (((U0[o] + U0[o - sy]) + U0[o + 1]) + U0[o - sy + 1]);
    v[1] = U1[o];
    v[2] = g->is3d ? 0.25f * (((U2[o] + U2[o - sy]) + U2[o + sz]) + U2[o + sz - sy]) : 0.f;
  } else { /* getAtMACZ :404-446 (3-D only) */
    v[0] = 0.25f * (((U0[o] + U0[o - sz]) + U0[o + 1]) + U0[o - sz + 1]);
    v[1] = 0.25f * (((U1[o] + U1[o - sz]) + U1[o + sy]) + U1[o - sz + sy]);
    v[2] = U2[o];
  }
}

/* ------------------------------------------------------------------------ */
/* advectScalar, fluids_init.cpp:265-382                                      */
/* ------------------------------------------------------------------------ */
/* one semi-Lagrangian pass (SemiLagrangeEulerFluidNetSavePos :69-133); border cells get 0 */
static int semi_lagrange_scalar(const grid_t *g, const float *flags, const float *U,
                                const float *src, float dt, int sample_outside, float *out,
                                float *out_pos /* 3*n or NULL */) {
  size_t n = nvox(g);
  int errors = 0;
  for (int k = 0; k < g->D; k++)
    for (int j = 0; j < g->H; j++)
      for (int i = 0; i < g->W; i++) {
        size_t o = at(g, k, j, i);
        float start[3] = {(float)i + 0.5f, (float)j + 0.5f, (float)k + 0.5f};
        if (is_border(g, k, j, i, 1)) {
          out[o] = 0.f;
          if (out_pos) { out_pos[o] = start[0]; out_pos[n + o] = start[1]; out_pos[2 * n + o] = start[2]; }
          continue;
        }
        if (flags[o] != TYPE_FLUID) { /* don't advect solid geometry */
          out[o] = src[o];
          if (out_pos) { out_pos[o] = start[0]; out_pos[n + o] = start[1]; out_pos[2 * n + o] = start[2]; }
          continue;
        }
        float c[3], delta[3], back[3];
        get_centered(g, U, k, j, i, c);
        for (int a = 0; a < 3; a++) delta[a] = (-dt) * c[a];
        errors += line_trace(g, flags, start, delta, 1, back);
        out[o] = sample_outside ? interp_field(g, src, back) : interp_with_fluid(g, src, flags, back);
        if (out_pos) { out_pos[o] = back[0]; out_pos[n + o] = back[1]; out_pos[2 * n + o] = back[2]; }
      }
  return errors;
}

/* method: 0 = eulerFluidNet, 1 = maccormackFluidNet (advect_type.cpp:5-16) */
int orc_advect_scalar(float dt, const float *src, const float *U, const float *flags, int B, int D,
                      int H, int W, int is3d, int method, int sample_outside, float strength,
                      float *dst) {
  grid_t g = {B, D, H, W, is3d};
  size_t n = nvox(&g);
  int nc = is3d ? 3 : 2, errors = 0;
  float *fwd = (float *)malloc(n * sizeof(float)), *bwd = (float *)malloc(n * sizeof(float));
  float *fpos = (float *)malloc(3 * n * sizeof(float));
  for (int b = 0; b < B; b++) {
    const float *s = src + b * n, *u = U + (size_t)b * nc * n, *f = flags + b * n;
    float *d = dst + b * n;
    if (method == 0) {
      errors += semi_lagrange_scalar(&g, f, u, s, dt, sample_outside, d, NULL);
      continue;
    }
    errors += semi_lagrange_scalar(&g, f, u, s, dt, sample_outside, fwd, fpos);
    errors += semi_lagrange_scalar(&g, f, u, fwd, -dt, sample_outside, bwd, NULL);
    float t = strength * 0.5f; /* :145 scalar product first */
    for (int k = 0; k < D; k++)
      for (int j = 0; j < H; j++)
        for (int i = 0; i < W; i++) {
          size_t o = at(&g, k, j, i);
          /* MacCormackCorrect :135-148 (all cells, incl. border) */
          float v = fwd[o];
          if (f[o] == TYPE_FLUID) v = fwd[o] + t * (s[o] - bwd[o]);
          if (!is_border(&g, k, j, i, 1)) { /* MacCormackClampFluidNet :224-263 + getClampBounds :154-222 */
            long i0 = clampl((long)fpos[o], 0, W - 1);
            long j0 = clampl((long)fpos[n + o], 0, H - 1);
            /* reference: k0 = 0 because src has one channel (:177); 3-D intent: trunc(pos.z) */
            long k0 = is3d ? clampl((long)fpos[2 * n + o], 0, D - 1) : 0;
            float mn = INFINITY, mx = -INFINITY;
            int ncells = 0;
            for (int dk = -1; dk <= 1; dk++)
              for (int dj = -1; dj <= 1; dj++)
                for (int di = -1; di <= 1; di++) {
                  long ii = i0 + di, jj = j0 + dj, kk = k0 + dk;
                  if (kk < 0 || kk >= D || jj < 0 || jj >= H || ii < 0 || ii >= W) continue;
                  size_t q = at(&g, (int)kk, (int)jj, (int)ii);
                  if (f[q] == TYPE_FLUID || sample_outside) {
                    mn = fminf_t(mn, s[q]);
                    mx = fmaxf_t(mx, s[q]);
                    ncells++;
                  }
                }
            v = (ncells >= 1) ? fmaxf_t(mn, fminf_t(mx, v)) : fwd[o];
          }
          d[o] = v;
        }
  }
  free(fwd); free(bwd); free(fpos);
  return errors;
}

/* ------------------------------------------------------------------------ */
/* advectVel, fluids_init.cpp:656-807                                         */
/* ------------------------------------------------------------------------ */
/* SemiLagrangeEulerFluidNetMAC :388-451 ; border cells get 0 */
static void semi_lagrange_mac(const grid_t *g, const float *flags, const float *U, const float *src,
                              float dt, float *out) {
  size_t n = nvox(g);
  int nc = g->is3d ? 3 : 2;
  for (int k = 0; k < g->D; k++)
    for (int j = 0; j < g->H; j++)
      for (int i = 0; i < g->W; i++) {
        size_t o = at(g, k, j, i);
        if (is_border(g, k, j, i, 1)) {
          for (int c = 0; c < nc; c++) out[c * n + o] = 0.f;
          continue;
        }
        if (flags[o] != TYPE_FLUID) {
          if (!g->is3d) { /* Q1: second scatter writes channel 1 into channel 0 (:413-416) */
            out[o] = src[n + o];
            out[n + o] = 0.f;
          } else { /* 3-D intent: solid cells keep src */
            for (int c = 0; c < nc; c++) out[c * n + o] = src[c * n + o];
          }
          continue;
        }
        float pos[3] = {(float)i + 0.5f, (float)j + 0.5f, (float)k + 0.5f};
        for (int c = 0; c < nc; c++) {
          float v[3], p[3];
          get_at_mac(g, U, c, k, j, i, v);
          for (int a = 0; a < 3; a++) p[a] = pos[a] + v[a] * (-dt); /* no line trace (Q2) */
          out[c * n + o] = interp_field(g, src + c * n, p);
        }
      }
}

int orc_advect_vel(float dt, const float *orig, const float *U, const float *flags, int B, int D,
                   int H, int W, int is3d, int method, float strength, float *dst) {
  grid_t g = {B, D, H, W, is3d};
  size_t n = nvox(&g);
  int nc = is3d ? 3 : 2;
  float *fwd = (float *)malloc(nc * n * sizeof(float)), *bwd = (float *)malloc(nc * n * sizeof(float));
  for (int b = 0; b < B; b++) {
    const float *og = orig + (size_t)b * nc * n, *u = U + (size_t)b * nc * n, *f = flags + b * n;
    float *d = dst + (size_t)b * nc * n;
    if (method == 0) {
      semi_lagrange_mac(&g, f, u, og, dt, d);
      continue;
    }
    semi_lagrange_mac(&g, f, u, og, dt, fwd);
    semi_lagrange_mac(&g, f, u, fwd, -dt, bwd);
    float t = strength * 0.5f; /* :495 */
    for (int k = 0; k < D; k++)
      for (int j = 0; j < H; j++)
        for (int i = 0; i < W; i++) {
          size_t o = at(&g, k, j, i);
          if (is_border(&g, k, j, i, 1)) {
            for (int c = 0; c < nc; c++) d[c * n + o] = 0.f;
            continue;
          }
          int solid = f[o] != TYPE_FLUID;
          for (int c = 0; c < nc; c++) {
            /* MacCormackCorrectMAC :453-498 */
            int skip = solid;
            if (c == 0 && i > 0 && f[o - 1] != TYPE_FLUID) skip = 1;
            if (c == 1 && j > 0 && f[o - W] != TYPE_FLUID) skip = 1;
            if (c == 2 && k > 0 && f[o - (size_t)H * W] != TYPE_FLUID) skip = 1;
            float fw = fwd[c * n + o];
            float v = skip ? fw : fw + t * (og[c * n + o] - bwd[c * n + o]);
            /* doClampComponentMAC :500-614 */
            float vel[3], posf[3] = {(float)i, (float)j, (float)k};
            get_at_mac(&g, u, c, k, j, i, vel);
            float mn = INFINITY, mx = -INFINITY;
            for (int l = 0; l < 2; l++) {
              long p[3];
              for (int a = 0; a < 3; a++) {
                float va = vel[a] * dt;
                float q = l == 0 ? posf[a] - va : posf[a] + va;
                p[a] = (long)(int32_t)q; /* toType(kInt): trunc toward zero */
              }
              long i0 = clampl(p[0], 0, W - 2), j0 = clampl(p[1], 0, H - 2);
              long k0 = clampl(p[2], 0, is3d ? D - 2 : 0);
              long k1 = is3d ? k0 + 1 : k0;
              const float *oc = og + c * n;
              float s[8] = {oc[at(&g, k0, j0, i0)],     oc[at(&g, k0, j0, i0 + 1)],
                            oc[at(&g, k0, j0 + 1, i0)], oc[at(&g, k0, j0 + 1, i0 + 1)],
                            oc[at(&g, k1, j0, i0)],     oc[at(&g, k1, j0, i0 + 1)],
                            oc[at(&g, k1, j0 + 1, i0)], oc[at(&g, k1, j0 + 1, i0 + 1)]};
              for (int q = 0; q < (is3d ? 8 : 4); q++) {
                mn = fminf_t(mn, s[q]);
                mx = fmaxf_t(mx, s[q]);
              }
            }
            d[c * n + o] = fmaxf_t(fminf_t(v, mx), mn);
          }
        }
  }
  free(fwd); free(bwd);
  return 0;
}

/* ------------------------------------------------------------------------ */
/* source_terms.py:6-116 addBuoyancy (Q14) ; :122-219 addGravity              */
/* ------------------------------------------------------------------------ */
void orc_add_buoyancy(float *U, const float *flags, const float *density, const float gravity[3],
                      float rho_star, float dt, int B, int D, int H, int W, int is3d) {
  grid_t g = {B, D, H, W, is3d};
  size_t n = nvox(&g);
  int nc = is3d ? 3 : 2;
  float strength[3] = {gravity[0] * dt, gravity[1] * dt, gravity[2] * dt};
  size_t nb[3] = {1, (size_t)W, (size_t)H * W};
  for (int b = 0; b < B; b++) {
    float *u = U + (size_t)b * nc * n;
    const float *f = flags + b * n, *r = density + b * n;
    for (int k = 0; k < D; k++)
      for (int j = 0; j < H; j++)
        for (int i = 0; i < W; i++) {
          size_t o = at(&g, k, j, i);
          if (is_border(&g, k, j, i, 1) || f[o] != TYPE_FLUID) continue;
          for (int c = 0; c < nc; c++) {
            if (f[o - nb[c]] != TYPE_FLUID) continue;
            float factor = strength[c] * (0.5f * (r[o] + r[o - nb[c]]) - rho_star);
            u[c * n + o] = u[c * n + o] + factor;
          }
        }
  }
}

void orc_add_gravity(float *U, const float *flags, const float gravity[3], float dt, int B, int D,
                     int H, int W, int is3d) {
  /* source_terms.py:176-216: interior cells that are Fluid or Empty; component c forced when the
     lower neighbour is Fluid, or when it is Empty and the cell itself is Fluid */
  grid_t g = {B, D, H, W, is3d};
  size_t n = nvox(&g);
  int nc = is3d ? 3 : 2;
  float force[3] = {gravity[0] * dt, gravity[1] * dt, gravity[2] * dt};
  size_t nb[3] = {1, (size_t)W, (size_t)H * W};
  for (int b = 0; b < B; b++) {
    float *u = U + (size_t)b * nc * n;
    const float *f = flags + b * n;
    for (int k = 0; k < D; k++)
      for (int j = 0; j < H; j++)
        for (int i = 0; i < W; i++) {
          size_t o = at(&g, k, j, i);
          if (is_border(&g, k, j, i, 1)) continue;
          int cf = f[o] == TYPE_FLUID, ce = f[o] == TYPE_EMPTY;
          if (!cf && !ce) continue;
          for (int c = 0; c < nc; c++) {
            float fn = f[o - nb[c]];
            if (fn == TYPE_FLUID || (fn == TYPE_EMPTY && cf)) u[c * n + o] = u[c * n + o] + force[c];
          }
        }
  }
}

/* viscosity.py:7-70 addViscosity (2-D only; the 3-D branch of the reference references an undefined
   mask).  In place on the interior; every right-hand side term is read from the field BEFORE the
   update (the reference evaluates the whole expression, then assigns).  Component c is kept only
   where the cell and its lower neighbour along c are both Fluid, else it becomes 0.  The fourth
   neighbour term is U(i-1, j-1) as written at :68 (not U(i, j-1)).  s = (float)(dt * viscosity) with
   the product formed in double (two Python floats). */
void orc_add_viscosity(float *U, const float *flags, double dt, double viscosity, int B, int H, int W) {
  size_t n = (size_t)H * W;
  float s = (float)(dt * viscosity);
  float *old = (float *)malloc(2 * n * sizeof(float));
  for (int b = 0; b < B; b++) {
    float *u = U + (size_t)b * 2 * n;
    const float *f = flags + (size_t)b * n;
    memcpy(old, u, 2 * n * sizeof(float));
    for (int j = 1; j < H - 1; j++)
      for (int i = 1; i < W - 1; i++) {
        size_t o = (size_t)j * W + i;
        int fl = f[o] == TYPE_FLUID;
        float m[2] = {(fl && f[o - 1] == TYPE_FLUID) ? 1.f : 0.f, (fl && f[o - W] == TYPE_FLUID) ? 1.f : 0.f};
        for (int c = 0; c < 2; c++) {
          const float *q = old + c * n;
          float lap = (((q[o + 1] + q[o + W]) + q[o - 1]) + q[o - W - 1]) - (4.f * q[o]);
          u[c * n + o] = m[c] * (q[o] + s * lap);
        }
      }
  }
  free(old);
}

/* advection.py:9-12 correctScalar: src += dt*0.5*src*div on Fluid cells, in place;
   t = (float)(dt*0.5) (double product), then ((t*src)*div) in fp32 */
void orc_correct_scalar(float *src, const float *div, const float *flags, double dt, size_t count) {
  float t = (float)(dt * 0.5);
  for (size_t q = 0; q < count; q++)
    if (flags[q] == TYPE_FLUID) src[q] = src[q] + (t * src[q]) * div[q];
}

/* set_wall_bcs.py:4-86 (Q13) */
void orc_set_wall_bcs(float *U, const float *flags, int B, int D, int H, int W, int is3d) {
  grid_t g = {B, D, H, W, is3d};
  size_t n = nvox(&g);
  int nc = is3d ? 3 : 2;
  for (int b = 0; b < B; b++) {
    float *u = U + (size_t)b * nc * n;
    const float *f = flags + b * n;
    for (int k = 0; k < D; k++)
      for (int j = 0; j < H; j++)
        for (int i = 0; i < W; i++) {
          size_t o = at(&g, k, j, i);
          int cf = f[o] == TYPE_FLUID, co = f[o] == TYPE_OBSTACLE;
          if (!cf && !co) continue;
          int idx[3] = {i, j, k};
          size_t nb[3] = {1, (size_t)W, (size_t)H * W};
          for (int c = 0; c < nc; c++) {
            float fn = idx[c] <= 0 ? f[o] : f[o - nb[c]]; /* index 0 tests the cell itself */
            if (fn == TYPE_OBSTACLE || (co && fn == TYPE_FLUID)) u[c * n + o] = 0.f;
          }
        }
  }
}

/* velocity_divergence.py:4-74 (Q15) */
void orc_velocity_divergence(const float *U, const float *flags, float *div, int B, int D, int H,
                             int W, int is3d) {
  grid_t g = {B, D, H, W, is3d};
  size_t n = nvox(&g);
  int nc = is3d ? 3 : 2;
  for (int b = 0; b < B; b++) {
    const float *u = U + (size_t)b * nc * n, *f = flags + b * n;
    float *dv = div + b * n;
    for (int k = 0; k < D; k++)
      for (int j = 0; j < H; j++)
        for (int i = 0; i < W; i++) {
          size_t o = at(&g, k, j, i);
          float v = 0.f;
          if (!is_border(&g, k, j, i, 1)) {
            v = u[o] - u[o + 1] + u[n + o] - u[n + o + W];
            if (is3d) v = v + (u[2 * n + o] - u[2 * n + o + (size_t)H * W]);
          }
          if (f[o] == TYPE_OBSTACLE) v = 0.f;
          dv[o] = v;
        }
  }
}

/* velocity_update.py:6-162 (Q12): the sum-of-masked-products the reference forms */
void orc_velocity_update(const float *p, float *U, const float *flags, int B, int D, int H, int W,
                         int is3d) {
  grid_t g = {B, D, H, W, is3d};
  size_t n = nvox(&g);
  int nc = is3d ? 3 : 2;
  size_t nb[3] = {1, (size_t)W, (size_t)H * W};
  for (int b = 0; b < B; b++) {
    float *u = U + (size_t)b * nc * n;
    const float *f = flags + b * n, *pr = p + b * n;
    for (int k = 0; k < D; k++)
      for (int j = 0; j < H; j++)
        for (int i = 0; i < W; i++) {
          if (is_border(&g, k, j, i, 1)) continue;
          size_t o = at(&g, k, j, i);
          int cf = f[o] == TYPE_FLUID;
          int ce = (f[o] == TYPE_EMPTY) && (f[o] != TYPE_OUTFLOW);
          for (int c = 0; c < nc; c++) {
            float fn = f[o - nb[c]];
            float m1 = (cf && fn == TYPE_FLUID) ? 1.f : 0.f;
            float m2 = (cf && fn == TYPE_EMPTY) ? 1.f : 0.f;
            float m3 = (ce && fn == TYPE_FLUID) ? 1.f : 0.f;
            float m4 = (ce && fn == TYPE_EMPTY) ? 1.f : 0.f;
            float uc = u[c * n + o], P = pr[o], Pm = pr[o - nb[c]];
            u[c * n + o] = m1 * (uc - (P - Pm)) + m2 * (uc - P) + m3 * (uc + Pm) + m4 * 0.f;
          }
        }
  }
}

/* fluids_init.cpp:809-1004 solveLinearSystemJacobi (Q16).  Returns iterations run. */
int orc_jacobi(const float *flags, const float *div, int B, int D, int H, int W, int is3d,
               float p_tol, int max_iter, float *p_out, float *residual_out) {
  grid_t g = {B, D, H, W, is3d};
  size_t n = nvox(&g), tot = (size_t)B * n;
  float *pa = p_out, *pb = (float *)calloc(tot, sizeof(float));
  memset(pa, 0, tot * sizeof(float));
  float *cur = pa, *prev = pb;
  float denom = is3d ? 6.f : 4.f, residual = 0.f;
  int iter = 0;
  while (1) {
    for (int b = 0; b < B; b++) {
      const float *f = flags + b * n, *dv = div + b * n, *pp = prev + b * n;
      float *pc = cur + b * n;
      for (int k = 0; k < D; k++)
        for (int j = 0; j < H; j++)
          for (int i = 0; i < W; i++) {
            size_t o = at(&g, k, j, i);
            if (is_border(&g, k, j, i, 1) || f[o] == TYPE_OBSTACLE) { pc[o] = 0.f; continue; }
            float pC = pp[o];
            float p1 = f[o - 1] == TYPE_OBSTACLE ? pC : pp[o - 1];
            float p2 = f[o + 1] == TYPE_OBSTACLE ? pC : pp[o + 1];
            float p3 = f[o - W] == TYPE_OBSTACLE ? pC : pp[o - W];
            float p4 = f[o + W] == TYPE_OBSTACLE ? pC : pp[o + W];
            float p5 = 0.f, p6 = 0.f;
            if (is3d) { /* intent: z neighbours get the same Neumann rule (:935-943 are shadowed) */
              size_t sz = (size_t)H * W;
              p5 = f[o - sz] == TYPE_OBSTACLE ? pC : pp[o - sz];
              p6 = f[o + sz] == TYPE_OBSTACLE ? pC : pp[o + sz];
            }
            pc[o] = (p1 + p2 + p3 + p4 + p5 + p6 + dv[o]) / denom;
          }
    }
    /* residual = max_b || p - p_prev ||_2 over the two buffers (:966-972) */
    residual = 0.f;
    for (int b = 0; b < B; b++) {
      double acc = 0.0;
      for (size_t q = 0; q < n; q++) {
        float dd = pa[b * n + q] - pb[b * n + q];
        acc += (double)dd * (double)dd;
      }
      float r = (float)sqrt(acc);
      if (r > residual) residual = r;
    }
    if (residual < p_tol) break;
    iter++;
    if (iter >= max_iter) break;
    float *tmp = cur; cur = prev; prev = tmp;
  }
  if (cur == pb) memcpy(pa, pb, tot * sizeof(float));
  free(pb);
  if (residual_out) *residual_out = residual;
  return iter;
}

/* simulate.py:4-26 setConstVals: x = x*invmask + bc */
void orc_set_const_vals(float *x, const float *inv_mask, const float *bc, size_t count) {
  for (size_t q = 0; q < count; q++) x[q] = x[q] * inv_mask[q] + bc[q];
}

/* flags_to_occupancy.py:6-19 */
void orc_flags_to_occupancy(const float *flags, float *occ, size_t count) {
  for (size_t q = 0; q < count; q++) {
    float f = flags[q];
    occ[q] = f == TYPE_FLUID ? 0.f : (f == TYPE_OBSTACLE ? 1.f : f);
  }
}

/* ------------------------------------------------------------------------ */
/* CNN pieces (multi_scale_net.py:101-127, model.py:8-23).  The arithmetic    */
/* lives in PyTorch (not vendored); these restate the published definitions   */
/* with double accumulation and are pinned against torch CPU fp32 outputs.    */
/* ------------------------------------------------------------------------ */
/* nn.Conv2d, stride 1, zero padding k/2, NCHW, optional ReLU */
void orc_conv2d(const float *x, const float *w, const float *bias, float *y, int N, int Cin, int H,
                int W, int Cout, int K, int relu) {
  int pad = K / 2;
  for (int nimg = 0; nimg < N; nimg++)
    for (int co = 0; co < Cout; co++)
      for (int j = 0; j < H; j++)
        for (int i = 0; i < W; i++) {
          double acc = bias ? bias[co] : 0.0;
          for (int ci = 0; ci < Cin; ci++)
            for (int kj = 0; kj < K; kj++) {
              int jj = j + kj - pad;
              if (jj < 0 || jj >= H) continue;
              for (int ki = 0; ki < K; ki++) {
                int ii = i + ki - pad;
                if (ii < 0 || ii >= W) continue;
                acc += (double)x[(((size_t)nimg * Cin + ci) * H + jj) * W + ii] *
                       (double)w[(((size_t)co * Cin + ci) * K + kj) * K + ki];
              }
            }
          float v = (float)acc;
          if (relu && v < 0.f) v = 0.f;
          y[(((size_t)nimg * Cout + co) * H + j) * W + i] = v;
        }
}

/* F.interpolate(mode='bilinear', align_corners=False) to (Ho, Wo), NCHW (Q18) */
void orc_resize_bilinear(const float *x, float *y, int NC, int H, int W, int Ho, int Wo) {
  float sh = (float)H / (float)Ho, sw = (float)W / (float)Wo;
  for (int c = 0; c < NC; c++)
    for (int j = 0; j < Ho; j++) {
      float fy = sh * ((float)j + 0.5f) - 0.5f;
      if (fy < 0.f) fy = 0.f;
      int y0 = (int)fy;
      int y1 = y0 + (y0 < H - 1 ? 1 : 0);
      float ly = fy - (float)y0, hy = 1.f - ly;
      for (int i = 0; i < Wo; i++) {
        float fx = sw * ((float)i + 0.5f) - 0.5f;
        if (fx < 0.f) fx = 0.f;
        int x0 = (int)fx;
        int x1 = x0 + (x0 < W - 1 ? 1 : 0);
        float lx = fx - (float)x0, hx = 1.f - lx;
        const float *p = x + (size_t)c * H * W;
        y[((size_t)c * Ho + j) * Wo + i] = hy * (hx * p[(size_t)y0 * W + x0] + lx * p[(size_t)y0 * W + x1]) +
                                           ly * (hx * p[(size_t)y1 * W + x0] + lx * p[(size_t)y1 * W + x1]);
      }
    }
}

/* torch.std(x.view(B,-1), dim=1) with Bessel's correction, double accumulation (Q17) */
void orc_std_unbiased(const float *x, int B, size_t count, float *out) {
  for (int b = 0; b < B; b++) {
    double s = 0.0;
    for (size_t q = 0; q < count; q++) s += x[b * count + q];
    double mean = s / (double)count, ss = 0.0;
    for (size_t q = 0; q < count; q++) {
      double d = x[b * count + q] - mean;
      ss += d * d;
    }
    out[b] = (float)sqrt(ss / (double)(count - 1));
  }
}
