// jacobi_blocked.cu -- temporally blocked, register-resident 2-D Jacobi pressure iterations (sm_100a).
//
// Reference: solveLinearSystemJacobi, pytorch/lib/fluid/cpp/fluids_init.cpp:858-1003 (Q16):
//   p_new = ((((pL + pR) + pU) + pD) + div) / 4 on interior non-Obstacle cells, 0 elsewhere;
//   an Obstacle neighbour contributes the centre value (Neumann).
// The reference runs one whole-grid pass (≈490 ATen ops) per iteration.  Here one launch
// advances a 128 x (8*NW) tile by up to HALO = 8 iterations:
//   * a warp owns 8 rows x 128 columns; a lane owns 8 rows x 4 consecutive columns and keeps
//     p (32 values, as 16 packed fp32 pairs) and the Neumann/fixed bit masks in REGISTERS for
//     the whole launch; the divergence tile sits in shared memory (one LDS.128 per row and sweep);
//   * left/right neighbours come from the adjacent lanes by warp shuffle, up/down neighbours
//     inside the strip are registers; only the strip's first/last row is exchanged with the
//     neighbouring warps through a small double-buffered shared-memory array (one
//     __syncthreads per iteration);
//   * the tile is loaded with an 8-cell halo and only the inner (128-16) x (8*NW-16) cells,
//     whose dependency cone stays inside the tile, are written back (halo = exactly 2 lanes
//     and 1 warp per side, so every store is an aligned float4);
//   * warps whose cells touch no Obstacle/border cell run a branch-free sweep on fp32 PAIRS
//     (add.rn.f32x2 / mul.rn.f32x2 = SASS FADD2 / FMUL2, same rounding as the scalar ops):
//     2 FADD2 + 0.5 FMUL2 + 0.5 SHFL per cell-iteration;
//   * the first and last warp strip of a tile (all halo) skip the rows no later iteration needs.
// HBM traffic per iteration drops from 16 B/cell to (12*overlap + 4)/8 B/cell.  The per-cell
// arithmetic and its order are those of the one-iteration kernel (stencils.cu), so results are
// bit-identical to it (tests/test_gpu_parity.py::test_full_size_properties).  -fmad=false.
#include <cuda_runtime.h>
#include <stdlib.h>

#include "../../include/fluidstep.h"
#include "fluid_common.cuh"
#include "host_util.h"

namespace fnx {

constexpr int JB_R = 8;      // rows per warp strip
constexpr int JB_C = 4;      // columns per lane
constexpr int JB_TW = 128;   // tile width  = 32 lanes * 4
constexpr int JB_HALO = 8;   // halo cells = max iterations per launch

struct Row4 {
  float v[JB_C];
};

// sm_100a packed fp32: add.rn.f32x2 / mul.rn.f32x2 (SASS FADD2 / FMUL2) do two IEEE round-to-nearest operations on
// a 64-bit register pair in ONE issue slot.  The fp32 pipe rate is unchanged (tools/f32x2_probe.cu: 36.9 vs 36.2
// TFLOP/s, 0 rounding mismatches in 2 M random pairs incl. denormals / infinities / NaNs), but this kernel is bound
// by issue slots and latency, not by the pipe (36 % busy): pairing the cells halves its arithmetic instructions.
typedef unsigned long long pair_t;  // two fp32 values: lo = first, hi = second
__device__ __forceinline__ pair_t pk(float lo, float hi) {
  pair_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk(pair_t r, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r)); }
__device__ __forceinline__ pair_t add2(pair_t a, pair_t b) {
  pair_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ pair_t mul2(pair_t a, pair_t b) {
  pair_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// A lane's 4 consecutive cells of one row, held as the pairs a = (c0, c2), b = (c1, c3): the left neighbours of
// (c1, c3) are then exactly a, the right neighbours of (c0, c2) exactly b, and only (left, c1) / (c2, right) have to
// be assembled around the two shuffled halo values.  The divergence tile and the exchanged rows in shared memory
// use the same (c0, c2, c1, c3) order, so a 128-bit LDS delivers two ready pairs.
struct RowP {
  pair_t a, b;
};
__device__ __forceinline__ Row4 unpack_row(const RowP& r) {
  Row4 o;
  upk(r.a, o.v[0], o.v[2]);
  upk(r.b, o.v[1], o.v[3]);
  return o;
}
__device__ __forceinline__ RowP pack_row(const Row4& r) {
  RowP o;
  o.a = pk(r.v[0], r.v[2]);
  o.b = pk(r.v[1], r.v[3]);
  return o;
}
__device__ __forceinline__ RowP load_rowp(const ulonglong2* q) {
  const ulonglong2 t = *q;
  RowP o;
  o.a = t.x; o.b = t.y;
  return o;
}

// one Jacobi sweep over the lane's 8x4 patch.  `p` holds the old values and receives the new.
// up_row / dn_row: the neighbouring strips' adjacent rows in shared memory (NULL at the tile edge: any value does,
// the row lies in the discarded halo); the row below is fetched when the last row needs it, not before.
// SLOW applies the Neumann/fixed masks, RESID accumulates |p - p_prev|^2: both run cell by cell; the plain sweep
// runs on pairs.  The per-cell operation order is the one-iteration kernel's either way:
// ((((pL + pR) + pU) + pD) + div) * 0.25.
// EDGE: the tile's first / last warp strip is all halo, and iteration t of n only needs its rows within
// n - 1 - t of the tile's inner region: rows outside [r_lo, r_hi) are skipped (their stale values are never read
// by a row that is still needed).
template <bool SLOW, bool RESID, bool EDGE, bool PACKED>
__device__ __forceinline__ void jacobi_sweep(RowP (&p)[JB_R], const ulonglong2* __restrict__ sdv,
                                             const ulonglong2* __restrict__ up_row, const ulonglong2* __restrict__ dn_row,
                                             unsigned Lb, unsigned Rb, unsigned Ub,
                                             unsigned Db, unsigned fixedb, float& acc, int r_lo, int r_hi) {
  RowP prevp = p[0];
  if (up_row) prevp = load_rowp(up_row);
  const pair_t quarter = pk(0.25f, 0.25f);
#pragma unroll
  for (int rr = 0; rr < JB_R; rr++) {
    const RowP curp = p[rr];
    if (EDGE && (rr < r_lo || rr >= r_hi)) {  // warp-uniform
      prevp = curp;
      continue;
    }
    RowP downp;
    if (rr < JB_R - 1) {
      downp = p[rr + 1];
    } else {
      downp = curp;
      if (dn_row) downp = load_rowp(dn_row);
    }
    const ulonglong2 dvp = sdv[rr * (JB_TW / 4)];  // this lane's 4 divergence values of row rr: (d0, d2), (d1, d3)
    const Row4 cur = unpack_row(curp);
    const float left = __shfl_up_sync(0xffffffffu, cur.v[JB_C - 1], 1);
    const float right = __shfl_down_sync(0xffffffffu, cur.v[0], 1);
    if (PACKED && !SLOW && !RESID) {
      const pair_t la = pk(left, cur.v[1]), rb = pk(cur.v[2], right);
      const pair_t sa = add2(add2(add2(add2(la, curp.b), prevp.a), downp.a), dvp.x);
      const pair_t sb = add2(add2(add2(add2(curp.a, rb), prevp.b), downp.b), dvp.y);
      prevp = curp;
      p[rr].a = mul2(sa, quarter);
      p[rr].b = mul2(sb, quarter);
      continue;
    }
    const Row4 prev = unpack_row(prevp), down = unpack_row(downp);
    float dvr[JB_C];
    upk(dvp.x, dvr[0], dvr[2]);
    upk(dvp.y, dvr[1], dvr[3]);
    Row4 nw;
#pragma unroll
    for (int c = 0; c < JB_C; c++) {
      float p1 = c == 0 ? left : cur.v[c - 1];
      float p2 = c == JB_C - 1 ? right : cur.v[c + 1];
      float p3 = prev.v[c];
      float p4 = down.v[c];
      if (SLOW) {
        const unsigned bit = 1u << (rr * JB_C + c);
        if (Lb & bit) p1 = cur.v[c];
        if (Rb & bit) p2 = cur.v[c];
        if (Ub & bit) p3 = cur.v[c];
        if (Db & bit) p4 = cur.v[c];
      }
      float pn = (p1 + p2 + p3 + p4 + dvr[c]) * 0.25f;
      if (SLOW) {
        if (fixedb & (1u << (rr * JB_C + c))) pn = 0.f;
      }
      if (RESID) {
        const float d = pn - cur.v[c];
        acc = acc + d * d;  // per-lane fp32 partial over 32 cells, folded in double below
      }
      nw.v[c] = pn;
    }
    prevp = curp;
    p[rr] = pack_row(nw);
  }
}

// The divergence tile lives in SHARED memory (read once per sweep with one LDS.128 per row), not in
// registers: with p (32 values) alone a thread needs < 85 registers, so TWO 12-warp CTAs (or three 8-warp
// ones) are resident per SM = 24 warps (the version that also kept div in registers ran one 16-warp
// CTA at 128 registers: 25 % occupancy, issue slots half empty).
// The flags never change during a solve and every launch of a solve has the same tile geometry: the five mask
// words of a thread (Neumann L/R/U/D and "fixed" bits of its 8x4 cells) are computed ONCE per solve by
// k_jacobi2d_tilemask and re-loaded (20 bytes per thread) by the launches (PRE = true), instead of decoding 32
// fp32 flags per thread in every launch -- a third of a launch's instructions and a quarter of its traffic.
template <int NW>
__device__ __forceinline__ void tile_masks(int H, int W, int ya0, int ya1, int gx0, int gy0, int vec_ok,
                                           const float* __restrict__ flags, unsigned (&xob)[NW][32], int w, int lane,
                                           unsigned& Lb, unsigned& Rb, unsigned& Ub, unsigned& Db, unsigned& fixedb) {
  unsigned obw = 0;
  fixedb = 0;
  const bool xvec = vec_ok && gx0 >= 0 && gx0 + JB_C <= W;
#pragma unroll
  for (int rr = 0; rr < JB_R; rr++) {
    const int gy = gy0 + rr;
    const bool yin = gy >= ya0 && gy < ya1;
    float f[JB_C];
    if (xvec && yin) {
      const float4 f4 = __ldg(reinterpret_cast<const float4*>(flags + (long long)gy * W + gx0));
      f[0] = f4.x; f[1] = f4.y; f[2] = f4.z; f[3] = f4.w;
    } else {
#pragma unroll
      for (int c = 0; c < JB_C; c++) {
        const int gx = gx0 + c;
        const bool inb = yin && gx >= 0 && gx < W;
        f[c] = inb ? __ldg(flags + (long long)gy * W + gx) : -1.f;  // -1: outside the domain
      }
    }
#pragma unroll
    for (int c = 0; c < JB_C; c++) {
      const int gx = gx0 + c;
      const bool ob = f[c] == kObstacle;
      const bool border = (gx < 1) | (gx > W - 2) | (gy < 1) | (gy > H - 2);  // also covers outside
      if (ob) obw |= 1u << (rr * JB_C + c);
      if (ob || border) fixedb |= 1u << (rr * JB_C + c);
    }
  }
  // Neumann masks: which neighbours are Obstacle cells
  xob[w][lane] = obw;
  __syncthreads();
  const unsigned obl = __shfl_up_sync(0xffffffffu, obw, 1), obr = __shfl_down_sync(0xffffffffu, obw, 1);
  const unsigned obu = w > 0 ? xob[w - 1][lane] : 0u, obd = w < NW - 1 ? xob[w + 1][lane] : 0u;
  Lb = (obw << 1) & 0xEEEEEEEEu; Rb = (obw >> 1) & 0x77777777u;
  if (lane > 0) Lb |= (obl >> 3) & 0x11111111u;
  if (lane < 31) Rb |= (obr << 3) & 0x88888888u;
  Ub = (obw << JB_C) | (obu >> (JB_C * (JB_R - 1)));
  Db = (obw >> JB_C) | (obd << (JB_C * (JB_R - 1)));
}

template <int NW>
__global__ void __launch_bounds__(NW * 32)
    k_jacobi2d_tilemask(int H, int W, int row0, int ya0, int ya1, int vec_ok, const float* __restrict__ flags,
                        uint4* __restrict__ m4, unsigned* __restrict__ m1) {
  constexpr int TH = NW * JB_R;
  constexpr int OW = JB_TW - 2 * JB_HALO, OH = TH - 2 * JB_HALO;
  __shared__ unsigned xob[NW][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int gx0 = blockIdx.x * OW - JB_HALO + lane * JB_C;
  const int gy0 = row0 + blockIdx.y * OH - JB_HALO + w * JB_R;
  flags += (long long)blockIdx.z * (ya1 - ya0) * W - (long long)ya0 * W;
  unsigned Lb, Rb, Ub, Db, fixedb;
  tile_masks<NW>(H, W, ya0, ya1, gx0, gy0, vec_ok, flags, xob, w, lane, Lb, Rb, Ub, Db, fixedb);
  const size_t t = ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * (NW * 32) + threadIdx.x;
  m4[t] = make_uint4(Lb, Rb, Ub, Db);
  m1[t] = fixedb;
}

template <int NW, bool FIRST, bool RESID, bool PRE, bool PACKED>
__global__ void __launch_bounds__(NW * 32, (NW <= 8 ? 3 : 2))
    k_jacobi2d_blocked(int H, int W, int row0, int row1, int ya0, int ya1, int iters, int vec_ok, const float* __restrict__ flags,
                       const float* __restrict__ div, const float* __restrict__ prev,
                       float* __restrict__ cur, double* __restrict__ ssq, const uint4* __restrict__ m4,
                       const unsigned* __restrict__ m1) {
  constexpr int TH = NW * JB_R;
  extern __shared__ __align__(16) float sdv_all[];      // [TH][JB_TW] divergence tile
  __shared__ __align__(16) float xch[2][NW][2][JB_TW];  // [parity][warp][top|bottom][column]
  __shared__ unsigned xob[NW][32];
  __shared__ double wsum[NW];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  constexpr int OW = JB_TW - 2 * JB_HALO, OH = TH - 2 * JB_HALO;
  const int gx0 = blockIdx.x * OW - JB_HALO + lane * JB_C;
  const int gy0 = row0 + blockIdx.y * OH - JB_HALO + w * JB_R;  // [row0, row1): rows this launch writes
  // the arrays hold rows [ya0, ya1) of the grid (a slab; 0, H = everything) and are addressed through
  // virtual bases with global row indices; rows that are not held read as "outside" (they lie in the halo
  // of a slab, whose values never reach the rows written)
  const long long boff = (long long)blockIdx.z * (ya1 - ya0) * W - (long long)ya0 * W;
  flags += boff; div += boff; cur += boff;
  if (!FIRST) prev += boff;

  RowP p[JB_R];
  // this lane's slot of the divergence tile: one 16-byte entry per row holding (d0, d2), (d1, d3)
  ulonglong2* const sdv = reinterpret_cast<ulonglong2*>(sdv_all) + (w * JB_R) * (JB_TW / 4) + lane;
  const RowP zero_row = {pk(0.f, 0.f), pk(0.f, 0.f)};
  const bool xvec = vec_ok && gx0 >= 0 && gx0 + JB_C <= W;
  // the whole 8 x 128 strip of this warp inside the held rows and the grid (all but the tiles on the rim):
  // 16 unconditional 128-bit loads through two row pointers
  const bool strip_in = __all_sync(0xffffffffu, xvec && gy0 >= ya0 && gy0 + JB_R <= ya1);
  if (strip_in) {
    const long long o = (long long)gy0 * W + gx0;
    const float4* dp = reinterpret_cast<const float4*>(div + o);
    const float4* pp = FIRST ? nullptr : reinterpret_cast<const float4*>(prev + o);
    const int w4 = W >> 2;
#pragma unroll
    for (int rr = 0; rr < JB_R; rr++) {
      const float4 d4 = __ldg(dp + rr * w4);
      sdv[rr * (JB_TW / 4)] = make_ulonglong2(pk(d4.x, d4.z), pk(d4.y, d4.w));
      if (!FIRST) {
        const float4 p4 = __ldg(pp + rr * w4);
        p[rr].a = pk(p4.x, p4.z); p[rr].b = pk(p4.y, p4.w);
      } else {
        p[rr] = zero_row;
      }
    }
  } else {
#pragma unroll
  for (int rr = 0; rr < JB_R; rr++) {
    const int gy = gy0 + rr;
    const bool yin = gy >= ya0 && gy < ya1;
    if (xvec && yin) {
      const long long o = (long long)gy * W + gx0;
      const float4 d4 = __ldg(reinterpret_cast<const float4*>(div + o));
      sdv[rr * (JB_TW / 4)] = make_ulonglong2(pk(d4.x, d4.z), pk(d4.y, d4.w));
      if (!FIRST) {
        const float4 p4 = __ldg(reinterpret_cast<const float4*>(prev + o));
        p[rr].a = pk(p4.x, p4.z); p[rr].b = pk(p4.y, p4.w);
      }
    } else {
      float dvv[JB_C];
      Row4 pv;
#pragma unroll
      for (int c = 0; c < JB_C; c++) {
        const int gx = gx0 + c;
        const bool inb = yin && gx >= 0 && gx < W;
        const long long o = (long long)gy * W + gx;
        dvv[c] = inb ? __ldg(div + o) : 0.f;
        pv.v[c] = (!FIRST && inb) ? __ldg(prev + o) : 0.f;
      }
      sdv[rr * (JB_TW / 4)] = make_ulonglong2(pk(dvv[0], dvv[2]), pk(dvv[1], dvv[3]));
      if (!FIRST) p[rr] = pack_row(pv);
    }
    if (FIRST) p[rr] = zero_row;
  }
  }
  unsigned Lb, Rb, Ub, Db, fixedb;
  if (PRE) {
    const size_t t = ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * (NW * 32) + threadIdx.x;
    const uint4 q = __ldg(m4 + t);
    Lb = q.x; Rb = q.y; Ub = q.z; Db = q.w;
    fixedb = __ldg(m1 + t);
  } else {
    tile_masks<NW>(H, W, ya0, ya1, gx0, gy0, vec_ok, flags, xob, w, lane, Lb, Rb, Ub, Db, fixedb);
  }
  const bool slow = __any_sync(0xffffffffu, (Lb | Rb | Ub | Db | fixedb) != 0u);

  float acc = 0.f;
  for (int t = 0; t < iters; t++) {
    // publish the strip's first and last row (old values), fetch the neighbours'
    float* mine = &xch[t & 1][w][0][lane * JB_C];
    *reinterpret_cast<ulonglong2*>(mine) = make_ulonglong2(p[0].a, p[0].b);
    *reinterpret_cast<ulonglong2*>(mine + JB_TW) = make_ulonglong2(p[JB_R - 1].a, p[JB_R - 1].b);
    // (one __syncthreads per iteration: neighbour-pair named barriers -- bar.sync id, 64, even boundaries first --
    // were measured and lost, 138 vs 100 us per launch at 4096^2: two barrier instructions per iteration and all 16
    // hardware barriers reserved per CTA cost more than the looser coupling saves)
    __syncthreads();
    const ulonglong2* up_row = w > 0 ? reinterpret_cast<const ulonglong2*>(&xch[t & 1][w - 1][1][lane * JB_C]) : nullptr;
    const ulonglong2* dn_row = w < NW - 1 ? reinterpret_cast<const ulonglong2*>(&xch[t & 1][w + 1][0][lane * JB_C]) : nullptr;
    if (RESID) acc = 0.f;  // only the last iteration's |p - p_prev|^2 survives
    if (slow) {
      jacobi_sweep<true, RESID, false, PACKED>(p, sdv, up_row, dn_row, Lb, Rb, Ub, Db, fixedb, acc, 0, JB_R);
    } else if (w == 0 || w == NW - 1) {
      // halo strips: iteration t (0-based) of `iters` is needed JB_HALO - (iters - 1 - t) rows into the tile only
      int skip = JB_HALO + 1 - iters + t;
      skip = skip < 0 ? 0 : (skip > JB_R ? JB_R : skip);
      const int r_lo = w == 0 ? skip : 0, r_hi = w == 0 ? JB_R : JB_R - skip;
      jacobi_sweep<false, RESID, true, PACKED>(p, sdv, up_row, dn_row, Lb, Rb, Ub, Db, fixedb, acc, r_lo, r_hi);
    } else {
      jacobi_sweep<false, RESID, false, PACKED>(p, sdv, up_row, dn_row, Lb, Rb, Ub, Db, fixedb, acc, 0, JB_R);
    }
  }

  // write back the cells whose dependency cone stayed inside the tile: warps 1..NW-2, lanes 2..29
  const bool owner = w >= 1 && w <= NW - 2 && lane >= 2 && lane <= 29;
  if (owner && vec_ok && gx0 + JB_C <= W && gy0 + JB_R <= row1) {
    float4* op = reinterpret_cast<float4*>(cur + (long long)gy0 * W + gx0);
    const int w4 = W >> 2;
#pragma unroll
    for (int rr = 0; rr < JB_R; rr++) {
      const Row4 o4 = unpack_row(p[rr]);
      op[rr * w4] = make_float4(o4.v[0], o4.v[1], o4.v[2], o4.v[3]);
    }
  } else if (owner) {
#pragma unroll
    for (int rr = 0; rr < JB_R; rr++) {
      const int gy = gy0 + rr;
      if (gy >= row1) break;
      const long long o = (long long)gy * W + gx0;
      const Row4 o4 = unpack_row(p[rr]);
      if (vec_ok && gx0 + JB_C <= W) {
        *reinterpret_cast<float4*>(cur + o) = make_float4(o4.v[0], o4.v[1], o4.v[2], o4.v[3]);
      } else {
#pragma unroll
        for (int c = 0; c < JB_C; c++)
          if (gx0 + c < W) cur[o + c] = o4.v[c];
      }
    }
  }
  if (RESID) {
    // the residual only counts cells this lane owns AND that exist (fixed/outside cells give 0)
    double a = owner ? (double)acc : 0.0;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
    if (lane == 0) wsum[w] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tsum = 0.0;
#pragma unroll
      for (int q = 0; q < NW; q++) tsum += wsum[q];
      atomicAdd(ssq + blockIdx.z, tsum);
    }
  }
}

}  // namespace fnx

using namespace fnx;

// iterations fused per launch: the halo is 8 cells, FNX_JACOBI_T (1..8) lowers it for tuning
static int jacobi_block_iters() {
  static int t = 0;
  if (t == 0) {
    const char* e = getenv("FNX_JACOBI_T");
    t = e ? atoi(e) : JB_HALO;
    if (t < 1) t = 1;
    if (t > JB_HALO) t = JB_HALO;
  }
  return t;
}

template <int NW, bool PRE, bool PACKED>
static void launch_blocked_v(bool first, bool resid, dim3 grid, cudaStream_t st, int H, int W, int row0, int row1,
                           int ya0, int ya1, int iters, int vec_ok,
                           const float* flags, const float* div, const float* prev, float* cur, double* ssq,
                           const uint4* m4, const unsigned* m1) {
  const int threads = NW * 32;
  constexpr size_t dsm = (size_t)NW * JB_R * JB_TW * sizeof(float);   // the divergence tile
  static bool attr_done = false;   // per instantiation; > 48 KB of dynamic shared memory must be opted in
  if (!attr_done) {
    cudaFuncSetAttribute(k_jacobi2d_blocked<NW, true, true, PRE, PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
    cudaFuncSetAttribute(k_jacobi2d_blocked<NW, true, false, PRE, PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
    cudaFuncSetAttribute(k_jacobi2d_blocked<NW, false, true, PRE, PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
    cudaFuncSetAttribute(k_jacobi2d_blocked<NW, false, false, PRE, PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
    attr_done = true;
  }
  if (first && resid) k_jacobi2d_blocked<NW, true, true, PRE, PACKED><<<grid, threads, dsm, st>>>(H, W, row0, row1, ya0, ya1, iters, vec_ok, flags, div, prev, cur, ssq, m4, m1);
  else if (first) k_jacobi2d_blocked<NW, true, false, PRE, PACKED><<<grid, threads, dsm, st>>>(H, W, row0, row1, ya0, ya1, iters, vec_ok, flags, div, prev, cur, ssq, m4, m1);
  else if (resid) k_jacobi2d_blocked<NW, false, true, PRE, PACKED><<<grid, threads, dsm, st>>>(H, W, row0, row1, ya0, ya1, iters, vec_ok, flags, div, prev, cur, ssq, m4, m1);
  else k_jacobi2d_blocked<NW, false, false, PRE, PACKED><<<grid, threads, dsm, st>>>(H, W, row0, row1, ya0, ya1, iters, vec_ok, flags, div, prev, cur, ssq, m4, m1);
}

// The sweep on fp32 PAIRS (FADD2 / FMUL2) issues a quarter fewer instructions; FNX_JACOBI_PACKED=0 selects the
// scalar sweep (same results bit for bit) for A/B measurements.
template <int NW, bool PRE>
static void launch_blocked(bool packed, bool first, bool resid, dim3 grid, cudaStream_t st, int H, int W, int row0, int row1,
                           int ya0, int ya1, int iters, int vec_ok,
                           const float* flags, const float* div, const float* prev, float* cur, double* ssq,
                           const uint4* m4, const unsigned* m1) {
  if (packed) launch_blocked_v<NW, PRE, true>(first, resid, grid, st, H, W, row0, row1, ya0, ya1, iters, vec_ok, flags, div, prev, cur, ssq, m4, m1);
  else launch_blocked_v<NW, PRE, false>(first, resid, grid, st, H, W, row0, row1, ya0, ya1, iters, vec_ok, flags, div, prev, cur, ssq, m4, m1);
}

// bytes of tile-mask scratch that make a multi-launch solve on a (B, H, W) grid skip the per-launch flag decode
size_t fnx_jacobi_2d_tilemask_bytes(int B, int H, int W) {
  constexpr int OW = JB_TW - 2 * JB_HALO;
  const size_t tx = (size_t)(W + OW - 1) / OW;
  const size_t a = (size_t)((H + 47) / 48) * 256, b = (size_t)((H + 79) / 80) * 384;   // threads per tile column, NW = 8 / 12
  return (size_t)B * tx * (a > b ? a : b) * 20 + 512;
}

int fnx_jacobi_2d_blocked_held(const float* flags, const float* div, const float* p_init, float* p, float* scratch,
                               double* ssq, int B, int H, int W, int max_iter, int row0, int row1, int ya0, int ya1,
                               void* tile_ws, size_t tile_ws_bytes, cudaStream_t st);

int fnx_jacobi_2d_blocked(const float* flags, const float* div, const float* p_init, float* p, float* scratch,
                          double* ssq, int B, int H, int W, int max_iter, int row0, int row1, void* tile_ws,
                          size_t tile_ws_bytes, cudaStream_t st) {
  return fnx_jacobi_2d_blocked_held(flags, div, p_init, p, scratch, ssq, B, H, W, max_iter, row0, row1, 0, H, tile_ws,
                                    tile_ws_bytes, st);
}

static int jacobi_2d_blocked_impl(const float* flags, const float* div, const float* p_init, float* p, float* scratch,
                                  double* ssq, int B, int H, int W, int max_iter, int row0, int row1, int ya0, int ya1,
                                  void* tile_ws, size_t tile_ws_bytes, int mask_mode, cudaStream_t st);

int fnx_jacobi_2d_blocked_held(const float* flags, const float* div, const float* p_init, float* p, float* scratch,
                               double* ssq, int B, int H, int W, int max_iter, int row0, int row1, int ya0, int ya1,
                               void* tile_ws, size_t tile_ws_bytes, cudaStream_t st) {
  return jacobi_2d_blocked_impl(flags, div, p_init, p, scratch, ssq, B, H, W, max_iter, row0, row1, ya0, ya1, tile_ws,
                                tile_ws_bytes, 0, st);
}

// mask_mode 1: only compute the tile masks of this launch geometry into tile_ws (no iteration);
// mask_mode 2: tile_ws already holds them (computed by a mask_mode-1 call with the same geometry): use, do not recompute
int fnx_jacobi_2d_blocked_masks(const float* flags, const float* div, const float* p_init, float* p, float* scratch,
                                int B, int H, int W, int max_iter, int row0, int row1, int ya0, int ya1, void* tile_ws,
                                size_t tile_ws_bytes, int mask_mode, cudaStream_t st) {
  return jacobi_2d_blocked_impl(flags, div, p_init, p, scratch, nullptr, B, H, W, max_iter, row0, row1, ya0, ya1, tile_ws,
                                tile_ws_bytes, mask_mode, st);
}

static int jacobi_2d_blocked_impl(const float* flags, const float* div, const float* p_init, float* p, float* scratch,
                                  double* ssq, int B, int H, int W, int max_iter, int row0, int row1, int ya0, int ya1,
                                  void* tile_ws, size_t tile_ws_bytes, int mask_mode, cudaStream_t st) {
  if (row1 <= row0) { row0 = 0; row1 = H; }
  if (!(0 <= ya0 && ya0 <= row0 && row1 <= ya1 && ya1 <= H))
    return fnx_set_error(FNX_ERR_ARG, "jacobi_2d_blocked: rows [%d,%d) not inside the held rows [%d,%d) of %d", row0, row1,
                         ya0, ya1, H);
  const int T = jacobi_block_iters();
  const int nL = (max_iter + T - 1) / T;
  auto wbuf = [&](int l) { return ((nL - 1 - l) % 2 == 0) ? p : scratch; };
  // float4 path: rows 16-byte aligned in every buffer
  const int vec_ok = (W % 4 == 0) && ((((uintptr_t)flags | (uintptr_t)div | (uintptr_t)p | (uintptr_t)scratch | (uintptr_t)p_init) & 15) == 0);
  // Tile height: 12-warp tiles (96 rows, 2 CTAs per SM) waste less halo work, 8-warp tiles (64 rows, 3 CTAs per
  // SM) quantise better: a launch costs about (waves of CTAs) x (rows of a tile), and a slab of a strong-scaled
  // grid is only one or two waves deep -- pick the cheaper shape for this launch.
  static const char* force = getenv("FNX_JACOBI_NW");
  constexpr int OW = JB_TW - 2 * JB_HALO;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
  }
  auto cost = [&](int nw) {
    const long long ctas = (long long)((W + OW - 1) / OW) * ((row1 - row0 + nw * JB_R - 2 * JB_HALO - 1) / (nw * JB_R - 2 * JB_HALO)) * B;
    const long long slots = (long long)sms * (nw <= 8 ? 3 : 2);
    return ((ctas + slots - 1) / slots) * (long long)(nw * JB_R);
  };
  bool tall = cost(12) < cost(8);
  if (force) tall = atoi(force) >= 12;
  const int oh = (tall ? 12 : 8) * JB_R - 2 * JB_HALO;
  dim3 grid((W + OW - 1) / OW, (row1 - row0 + oh - 1) / oh, B);
  // packed (FADD2 / FMUL2) sweep: measured faster or equal at every launch size (tools/jacobi_launch_time.py:
  // 546 / 1058 / 2082 / 4096 rows x 4096: 49 / 62 / 96 / 162 us against 51 / 65 / 98 / 166 us scalar)
  static const char* pk_env = getenv("FNX_JACOBI_PACKED");
  const bool packed = pk_env ? atoi(pk_env) != 0 : true;
  // tile masks once per solve (worth it from the second launch on)
  const size_t nthr = (size_t)grid.x * grid.y * grid.z * (tall ? 384 : 256);
  static const char* nopre = getenv("FNX_JACOBI_NO_TILEMASK");
  const bool room = tile_ws && tile_ws_bytes >= nthr * 20 + 256;
  if (mask_mode != 0 && !room) return fnx_set_error(FNX_ERR_WORKSPACE, "jacobi_2d_blocked: tile-mask buffer too small");
  const bool pre = mask_mode != 0 || (nL >= 2 && room && !(nopre && nopre[0] == '1'));
  uint4* m4 = nullptr;
  unsigned* m1 = nullptr;
  if (pre) {
    m4 = reinterpret_cast<uint4*>(((uintptr_t)tile_ws + 15) & ~(uintptr_t)15);
    m1 = reinterpret_cast<unsigned*>(m4 + nthr);
    if (mask_mode != 2) {
      if (tall) k_jacobi2d_tilemask<12><<<grid, 384, 0, st>>>(H, W, row0, ya0, ya1, vec_ok, flags, m4, m1);
      else k_jacobi2d_tilemask<8><<<grid, 256, 0, st>>>(H, W, row0, ya0, ya1, vec_ok, flags, m4, m1);
      fnx_count_launches(1);
    }
    if (mask_mode == 1) {
      cudaError_t e1 = cudaGetLastError();
      if (e1 != cudaSuccess) return fnx_set_error(FNX_ERR_CUDA, "jacobi_2d_blocked: %s", cudaGetErrorString(e1));
      return FNX_OK;
    }
  }
  int done = 0;
  for (int l = 0; l < nL; l++) {
    const int iters = (max_iter - done) < T ? (max_iter - done) : T;
    // p_init == nullptr: start from p = 0 (the reference); else continue from p_init (must not alias p / scratch)
    const bool first = l == 0 && p_init == nullptr, resid = l == nL - 1 && ssq != nullptr;
    const float* prev = l == 0 ? p_init : wbuf(l - 1);
    float* cur = wbuf(l);
    if (tall && pre) launch_blocked<12, true>(packed, first, resid, grid, st, H, W, row0, row1, ya0, ya1, iters, vec_ok, flags, div, prev, cur, ssq, m4, m1);
    else if (tall) launch_blocked<12, false>(packed, first, resid, grid, st, H, W, row0, row1, ya0, ya1, iters, vec_ok, flags, div, prev, cur, ssq, m4, m1);
    else if (pre) launch_blocked<8, true>(packed, first, resid, grid, st, H, W, row0, row1, ya0, ya1, iters, vec_ok, flags, div, prev, cur, ssq, m4, m1);
    else launch_blocked<8, false>(packed, first, resid, grid, st, H, W, row0, row1, ya0, ya1, iters, vec_ok, flags, div, prev, cur, ssq, m4, m1);
    done += iters;
    fnx_count_launches(1);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fnx_set_error(FNX_ERR_CUDA, "jacobi_2d_blocked: %s", cudaGetErrorString(e));
  return FNX_OK;
}
