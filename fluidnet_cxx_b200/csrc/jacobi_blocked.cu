// jacobi_blocked.cu -- temporally blocked 2-D Jacobi pressure iterations (sm_100a).
//
// Reference: solveLinearSystemJacobi, pytorch/lib/fluid/cpp/fluids_init.cpp:858-1003 (Q16):
//   p_new = ((((pL + pR) + pU) + pD) + div) / 4 on interior non-Obstacle cells, 0 elsewhere,
//   an Obstacle neighbour contributes the centre value (Neumann).
// The reference runs one whole-grid pass (≈490 ATen ops) per iteration.  Here one launch
// advances a (TW x TH) shared-memory tile by `iters` iterations: the tile is loaded once with an
// `iters`-cell halo, iterated in shared memory (ping-pong), and only the inner
// (TW-2*iters) x (TH-2*iters) cells -- whose dependency cone lies inside the tile -- are written
// back.  HBM traffic per iteration drops from 16 B/cell to ~ (12*overlap + 4)/iters B/cell.
// Each thread owns a vertical strip of R cells for the whole launch and keeps their div values
// and Neumann/fixed masks in registers; only p lives in shared memory.  Per-cell arithmetic and
// summation order are exactly those of the one-iteration kernel (stencils.cu), so results are
// bit-identical to it.  Compiled with -fmad=false.
#include <cuda_runtime.h>
#include <stdlib.h>

#include "../../include/fluidstep.h"
#include "fluid_common.cuh"
#include "host_util.h"

namespace fnx {

template <int TW, int TH, int R>
struct JacobiTile {
  static constexpr int kThreads = TW * (TH / R);
  static constexpr size_t kSmem = 2 * TW * TH * sizeof(float) + TW * TH;
};

template <int TW, int TH, int R, bool FIRST, bool RESID>
__global__ void __launch_bounds__(TW*(TH / R), 2)
    k_jacobi2d_blocked(int H, int W, int iters, const float* __restrict__ flags,
                       const float* __restrict__ div, const float* __restrict__ prev,
                       float* __restrict__ cur, double* __restrict__ ssq) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* pbuf = reinterpret_cast<float*>(smem_raw);                 // [2][TH][TW]
  unsigned char* obst = smem_raw + 2 * TW * TH * sizeof(float);     // [TH][TW]
  const int tid = threadIdx.x;
  const int x = tid % TW, r0 = (tid / TW) * R;
  const int ow = TW - 2 * iters, oh = TH - 2 * iters;
  const int gx = blockIdx.x * ow - iters + x;
  const int gy0 = blockIdx.y * oh - iters + r0;
  const long long boff = (long long)blockIdx.z * H * W;
  flags += boff; div += boff; cur += boff;
  if (!FIRST) prev += boff;

  float dv[R];
  unsigned fixedb = 0, Lb = 0, Rb = 0, Ub = 0, Db = 0;
  const bool xin = gx >= 0 && gx < W;
#pragma unroll
  for (int rr = 0; rr < R; rr++) {
    const int gy = gy0 + rr;
    const bool inb = xin && gy >= 0 && gy < H;
    const long long o = (long long)gy * W + gx;
    float f = kFluid, d = 0.f, p0 = 0.f;
    if (inb) {
      f = __ldg(flags + o);
      d = __ldg(div + o);
      if (!FIRST) p0 = __ldg(prev + o);
    }
    const bool ob = inb && f == kObstacle;
    const bool border = (gx < 1) | (gx > W - 2) | (gy < 1) | (gy > H - 2);
    if (!inb || border || ob) fixedb |= 1u << rr;
    dv[rr] = d;
    pbuf[(r0 + rr) * TW + x] = p0;
    obst[(r0 + rr) * TW + x] = ob ? 1 : 0;
  }
  __syncthreads();
#pragma unroll
  for (int rr = 0; rr < R; rr++) {
    const int r = r0 + rr;
    if (x > 0 && obst[r * TW + x - 1]) Lb |= 1u << rr;
    if (x < TW - 1 && obst[r * TW + x + 1]) Rb |= 1u << rr;
    if (r > 0 && obst[(r - 1) * TW + x]) Ub |= 1u << rr;
    if (r < TH - 1 && obst[(r + 1) * TW + x]) Db |= 1u << rr;
  }

  const bool xcomp = x > 0 && x < TW - 1;
  for (int t = 0; t < iters; t++) {
    const float* src = pbuf + (t & 1) * TW * TH;
    float* dst = pbuf + ((t + 1) & 1) * TW * TH;
    float up = r0 > 0 ? src[(r0 - 1) * TW + x] : 0.f;
    float c = src[r0 * TW + x];
#pragma unroll
    for (int rr = 0; rr < R; rr++) {
      const int r = r0 + rr;
      const float down = r < TH - 1 ? src[(r + 1) * TW + x] : 0.f;
      if (xcomp && r > 0 && r < TH - 1) {
        const float l = src[r * TW + x - 1], rt = src[r * TW + x + 1];
        const float p1 = (Lb >> rr & 1u) ? c : l;
        const float p2 = (Rb >> rr & 1u) ? c : rt;
        const float p3 = (Ub >> rr & 1u) ? c : up;
        const float p4 = (Db >> rr & 1u) ? c : down;
        float pn = (p1 + p2 + p3 + p4 + dv[rr]) * 0.25f;
        if (fixedb >> rr & 1u) pn = 0.f;
        dst[r * TW + x] = pn;
      }
      up = c;
      c = down;
    }
    __syncthreads();
  }
  // write back the cells whose dependency cone stayed inside the tile
  const float* fin = pbuf + (iters & 1) * TW * TH;
  const float* prv = pbuf + ((iters - 1) & 1) * TW * TH;
  double acc = 0.0;
  if (x >= iters && x < TW - iters && xin) {
    const int rlo = r0 > iters ? r0 : iters;
    int rhi = r0 + R < TH - iters ? r0 + R : TH - iters;
    const int gybase = blockIdx.y * oh - iters;
    if (rhi > H - gybase) rhi = H - gybase;
    for (int r = rlo; r < rhi; r++) {
      const float v = fin[r * TW + x];
      cur[(long long)(gybase + r) * W + gx] = v;
      if (RESID) {
        const float dd = v - prv[r * TW + x];
        acc += (double)(dd * dd);
      }
    }
  }
  if (RESID) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    __shared__ double wsum[TW * (TH / R) / 32];
    if ((tid & 31) == 0) wsum[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
      double tsum = 0.0;
#pragma unroll
      for (int w = 0; w < TW * (TH / R) / 32; w++) tsum += wsum[w];
      atomicAdd(ssq + blockIdx.z, tsum);
    }
  }
}

}  // namespace fnx

using namespace fnx;

// iterations fused per launch (halo width).  8 balances redundant halo work against HBM traffic
// for the 128x64 tile (DESIGN.md "Jacobi"); FNX_JACOBI_T overrides for tuning.
static int jacobi_block_iters() {
  static int t = 0;
  if (t == 0) {
    const char* e = getenv("FNX_JACOBI_T");
    t = e ? atoi(e) : 8;
    if (t < 1) t = 1;
    if (t > 24) t = 24;
  }
  return t;
}

int fnx_jacobi_2d_blocked(const float* flags, const float* div, float* p, float* scratch, double* ssq,
                          int B, int H, int W, int max_iter, cudaStream_t st) {
  constexpr int TW = 128, TH = 64, R = 16;
  using Tile = JacobiTile<TW, TH, R>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaSuccess;
    auto set = [&](const void* fn) {
      if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tile::kSmem);
    };
    set((const void*)k_jacobi2d_blocked<TW, TH, R, true, true>);
    set((const void*)k_jacobi2d_blocked<TW, TH, R, true, false>);
    set((const void*)k_jacobi2d_blocked<TW, TH, R, false, true>);
    set((const void*)k_jacobi2d_blocked<TW, TH, R, false, false>);
    if (e != cudaSuccess) return fnx_set_error(FNX_ERR_CUDA, "jacobi_2d_blocked: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int T = jacobi_block_iters();
  const int nL = (max_iter + T - 1) / T;
  auto wbuf = [&](int l) { return ((nL - 1 - l) % 2 == 0) ? p : scratch; };
  int done = 0;
  for (int l = 0; l < nL; l++) {
    const int iters = (max_iter - done) < T ? (max_iter - done) : T;
    const int ow = TW - 2 * iters, oh = TH - 2 * iters;
    dim3 grid((W + ow - 1) / ow, (H + oh - 1) / oh, B);
    const bool first = l == 0, resid = l == nL - 1;
    const float* prev = first ? nullptr : wbuf(l - 1);
    float* cur = wbuf(l);
    if (first && resid) k_jacobi2d_blocked<TW, TH, R, true, true><<<grid, Tile::kThreads, Tile::kSmem, st>>>(H, W, iters, flags, div, prev, cur, ssq);
    else if (first) k_jacobi2d_blocked<TW, TH, R, true, false><<<grid, Tile::kThreads, Tile::kSmem, st>>>(H, W, iters, flags, div, prev, cur, ssq);
    else if (resid) k_jacobi2d_blocked<TW, TH, R, false, true><<<grid, Tile::kThreads, Tile::kSmem, st>>>(H, W, iters, flags, div, prev, cur, ssq);
    else k_jacobi2d_blocked<TW, TH, R, false, false><<<grid, Tile::kThreads, Tile::kSmem, st>>>(H, W, iters, flags, div, prev, cur, ssq);
    done += iters;
    fnx_count_launches(1);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fnx_set_error(FNX_ERR_CUDA, "jacobi_2d_blocked: %s", cudaGetErrorString(e));
  return FNX_OK;
}
